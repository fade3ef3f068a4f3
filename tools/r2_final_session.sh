#!/bin/bash
# final single-GPU session of round 2: GPU suite, smoke, the three bench lines, launch list and
# DRAM bytes of one ungraphed step (tag = $1)
set -u
TAG=${1:-v8}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/r02_gpu_tests_$TAG.log
tail -3 $O/r02_gpu_tests_$TAG.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r02_bench_$TAG.json 2> $O/r02_bench_$TAG.err
tail -c 300 $O/r02_bench_$TAG.json; echo
for c in rb vk; do
  timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu > $O/r02_bench_${c}_$TAG.json 2> $O/r02_bench_${c}_$TAG.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $O/r02_launches_step_${TAG}_nograph.csv python tools/prof_step.py 4096 1 0 > /dev/null 2>&1
python tools/summarize_launches.py $O/r02_launches_step_${TAG}_nograph.csv > $O/r02_launches_step_${TAG}_nograph_summary.txt 2>&1
head -14 $O/r02_launches_step_${TAG}_nograph_summary.txt
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --profile-from-start off --csv --log-file $O/r02_step_dram_${TAG}.csv python tools/prof_step.py 4096 1 0 > /dev/null 2>&1
python tools/ncu_step_dram.py $O/r02_step_dram_${TAG}.csv > $O/r02_step_dram_${TAG}_summary.txt 2>&1
tail -2 $O/r02_step_dram_${TAG}_summary.txt
