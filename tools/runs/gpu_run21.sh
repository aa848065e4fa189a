#!/bin/bash
# Round-2 run 21 on one B200: lane-group shapes of the p-adic kernel (modulus p^2, q^2): 4 lanes x 8 limbs (default) against 2 lanes x 16 limbs
# and against 4 lanes with 4 blocks per SM, through TECDSA_NADIC_SHAPE, on the 8192-session batch.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
: > $O/r02_nadic_shape2.log
for rep in 1 2; do
  for shape in "8,1,4,4" "8,4,4,4" "8,4,4,1"; do
    echo "TECDSA_NADIC_SHAPE=$shape rep=$rep" >> $O/r02_nadic_shape2.log
    TECDSA_NADIC_SHAPE=$shape python tools/offline_throughput.py 8192 2>&1 | tail -1 >> $O/r02_nadic_shape2.log
  done
done
cat $O/r02_nadic_shape2.log
