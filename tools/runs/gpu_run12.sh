#!/bin/bash
# Round-2 run 12 on one B200: secp256k1 scalar-multiplication throughput (configs[2] shape, device buffers) and an ncu --set full capture
# of the variable-base launch (one report per call: the merge-back limit of gpurun_out is 64 MiB).
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
python tools/ec_throughput.py 20 2>&1 | tail -1 > $O/r02_ec_throughput.json
cat $O/r02_ec_throughput.json
ncu --set full --clock-control none --import-source on -k regex:k_secp_mul -s 3 -c 1 -o $O/r02_secp_mul_var -f python tools/ec_throughput.py 18 > $O/r02_ncu_ec.log 2>&1
tail -2 $O/r02_ncu_ec.log
ls -la $O/r02_secp_mul_*.ncu-rep
