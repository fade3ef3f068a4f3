set -u
O=gpurun_out; mkdir -p $O
for k in 3 1 4; do python tools/prof_stored_op.py 256 1024 2 $k 2>&1 | tail -2; done
python tools/prof_stored_op.py 256 1024 1 3 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o $O/r02_stored_op python tools/prof_stored_op.py 256 1024 2 3 > /dev/null 2>&1
ncu -i $O/r02_stored_op.ncu-rep --page raw --csv > $O/r02_stored_op_raw.csv 2>/dev/null
ncu -i $O/r02_stored_op.ncu-rep --page source --csv --print-source sass > $O/r02_stored_op_source.csv 2>/dev/null
ls -la $O/r02_stored_op*
