#!/bin/bash
# Round-2 run 22 on one B200: modexp kernel at 4 blocks per SM (128 registers, 8 B of spills) against 3 blocks (136 registers); the new
# N-adic default shape in the bench loop.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
: > $O/r02_modexp_minb.log
for rep in 1 2; do
  for v in default mb4; do
    if [ $v = default ]; then unset TECDSA_B200_LIB; else export TECDSA_B200_LIB=$PWD/multi-party-ecdsa_b200/variants/libtecdsa_b200_$v.so; fi
    for cfg in "2048 4" "4096 8" "1024 4"; do set -- $cfg; echo "variant=$v bits=$1 tpi=$2" >> $O/r02_modexp_minb.log; python tools/prof_modexp.py $1 $2 65536 2 2>&1 | tail -1 >> $O/r02_modexp_minb.log; done
  done
done
unset TECDSA_B200_LIB
python tools/offline_throughput.py 8192 2>&1 | tail -2 >> $O/r02_modexp_minb.log
cat $O/r02_modexp_minb.log
