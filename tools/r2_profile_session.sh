#!/bin/bash
# Round-2 measurement session on one B200: bench line, launch list of one step (no graphs, so
# that the solve's kernels are listed), ncu --set full of the step's kernels.
#   /usr/local/graft/bin/gpurun --timeout 1700 -- 'bash tools/r2_profile_session.sh v1'
set -u
TAG=${1:-v1}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_$TAG.json 2> gpurun_out/r02_bench_$TAG.err
tail -c 400 gpurun_out/r02_bench_$TAG.json; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r02_launches_step_${TAG}_nograph.csv python tools/prof_step.py 4096 1 0 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_step_${TAG}_nograph.csv > gpurun_out/r02_launches_step_${TAG}_nograph_summary.txt 2>&1
head -30 gpurun_out/r02_launches_step_${TAG}_nograph_summary.txt
if [ "${2:-full}" = "full" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_smooth2|k_resid_restrict|k_restrict|k_adv|k_map_vec|orthogradient|celltocorner' -c 110 \
    -f -o gpurun_out/r02_full_a python tools/prof_step.py 4096 1 0 > /dev/null 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_resid_sumsq|k_reduce1|k_mg_ctail' -c 8 \
    -f -o gpurun_out/r02_full_b python tools/prof_step.py 4096 1 0 > /dev/null 2>&1
  for f in a b; do
    ncu -i gpurun_out/r02_full_$f.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_${TAG}_$f.csv 2>/dev/null
    ls -la gpurun_out/r02_full_$f.ncu-rep
    # keep the report itself only when it is small enough to travel
    [ $(stat -c %s gpurun_out/r02_full_$f.ncu-rep) -gt 25000000 ] && rm -f gpurun_out/r02_full_$f.ncu-rep
  done
fi
