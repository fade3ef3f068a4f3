#!/bin/bash
# First GPU session of round 2: everything that was written at the end of round 1 without GPU
# access.  Usage (from the repo root):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/r2_first_gpu_session.sh'
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 1300 -- 'bash tools/r2_first_gpu_session.sh slabs'
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 1500 -- 'bash tools/r2_first_gpu_session.sh scaling'
set -u
mkdir -p gpurun_out
if [ "${1:-}" = "slabs" ]; then
  # needs --gpus 2: decomposition invariance, the no-slip Rayleigh-Benard channel included
  F2D_TEST_UNVERIFIED_SLABS=1 timeout 1200 python -m pytest tests/test_gpu_slabs.py -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/slab_tests.log
  exit 0
fi
if [ "${1:-}" = "scaling" ]; then
  # SURVEY 8d case S5: strong scaling of the 16384^2 Euler step (1-GPU run = the denominator)
  for N in 1 2 4 8; do
    if [ $N = 1 ]; then
      timeout 900 python bench.py --strong --n 16384 --steps 10 --warmup 3 --no-cpu > gpurun_out/s5_strong_1.json 2> gpurun_out/s5_strong_1.err
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2950$N \
        bench.py --gpus $N --strong --n 16384 --steps 10 --warmup 3 --no-cpu > gpurun_out/s5_strong_$N.json 2> gpurun_out/s5_strong_$N.err
    fi
    tail -c 600 gpurun_out/s5_strong_$N.json
  done
  # how many levels stay distributed (weak scaling line, 8 GPUs): every distributed level costs
  # one neighbour synchronisation per operator application, a gathered one costs redundant work
  for C in 1048576 4194304 16777216; do
    F2D_SLAB_MIN_CELLS=$C timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
      --master-port 29611 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/weak8_mincells_$C.json 2> gpurun_out/weak8_mincells_$C.err
    tail -c 300 gpurun_out/weak8_mincells_$C.json
  done
  exit 0
fi
# 1. the late tests alone (line relaxation, BoussinesqTS, QG diagnosed), verbose, under a timeout
timeout 900 python -m pytest tests/test_gpu_zz_late.py -m gpu -q 2>&1 | tail -60 | tee gpurun_out/late_tests.log
# 2. the whole GPU suite
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee gpurun_out/gpu_tests.log
# 3. how slow is the line relaxation (doublediffusion at its own size, 128 x 256)
timeout 300 python - <<'PY' 2>&1 | tail -3 | tee gpurun_out/tridiag_time.log
import sys, time, tempfile, io, contextlib
sys.path.insert(0, 'tests/golden')
import torch, cases, fluid2d_b200
api = fluid2d_b200.api()
for relax in ('default', 'tridiagonal'):
    with contextlib.redirect_stdout(io.StringIO()):
        f = cases.dbldiff(api, tempfile.mkdtemp(), 128, relaxation=relax)
        cases.run_steps(f, (2,))
        torch.cuda.synchronize(); t0 = time.time()
        cases.run_steps(f, (10,))
        torch.cuda.synchronize()
    print(relax, '%.2f ms/step' % ((time.time()-t0)*100))
PY
