export F2D_SLAB_MIN_CELLS=1500
for geom in perio xchannel obstacle; do
  echo "=== $geom"
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/debug_slab_mg.py 128 64 two_vcycle,two_vcycle,solve $geom 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -6
done
