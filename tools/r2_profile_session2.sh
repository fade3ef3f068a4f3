#!/bin/bash
# Round-2 measurement session (one B200): GPU suite, bench line, launch list of one ungraphed
# step, DRAM bytes of every kernel of that step, ncu --set full of the step's kernels.
#   /usr/local/graft/bin/gpurun --timeout 1700 -- 'bash tools/r2_profile_session2.sh v6'
set -u
TAG=${1:-v6}
O=gpurun_out
mkdir -p $O
if [ "${2:-tests}" = "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/r02_gpu_tests_$TAG.log
  tail -3 $O/r02_gpu_tests_$TAG.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 --no-s5 > $O/r02_bench_$TAG.json 2> $O/r02_bench_$TAG.err
tail -c 300 $O/r02_bench_$TAG.json; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $O/r02_launches_step_${TAG}_nograph.csv python tools/prof_step.py 4096 1 0 > /dev/null 2>&1
python tools/summarize_launches.py $O/r02_launches_step_${TAG}_nograph.csv > $O/r02_launches_step_${TAG}_nograph_summary.txt 2>&1
head -12 $O/r02_launches_step_${TAG}_nograph_summary.txt
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --profile-from-start off --csv --log-file $O/r02_step_dram_${TAG}.csv python tools/prof_step.py 4096 1 0 > /dev/null 2>&1
python tools/ncu_step_dram.py $O/r02_step_dram_${TAG}.csv > $O/r02_step_dram_${TAG}_summary.txt 2>&1
tail -2 $O/r02_step_dram_${TAG}_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'k_smooth2|k_resid_restrict|k_zsmooth|k_restrict|k_adv|k_map_vec|orthogradient|celltocorner|k_resid_sumsq|k_reduce1' -c 70 \
  -f -o $O/r02_full_$TAG python tools/prof_step.py 4096 1 0 > /dev/null 2>&1
ncu -i $O/r02_full_$TAG.ncu-rep --page raw --csv > $O/r02_ncu_full_${TAG}.csv 2>/dev/null
ls -la $O/r02_full_$TAG.ncu-rep
[ $(stat -c %s $O/r02_full_$TAG.ncu-rep) -gt 30000000 ] && rm -f $O/r02_full_$TAG.ncu-rep
python tools/ncu_dram_table.py $O/r02_ncu_full_${TAG}.csv | head -40
