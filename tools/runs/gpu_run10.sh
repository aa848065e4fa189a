#!/bin/bash
# Round-2 run 10 on one B200: parity of the section 8(f) rank-4 entry points (Lindell-17, zk_pdl, GG18), then an A/B of the N-adic
# pass loop unrolled over 1 / 2 / 4 lanes per trip (libraries built from the same sources with -DNADIC_GROUP_UNROLL=n).
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_other_protocols.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -60 > $O/r02_t10.log
tail -60 $O/r02_t10.log
# the unroll A/B of this script's first version (libraries built with -DNADIC_GROUP_UNROLL=1|2|4 under multi-party-ecdsa_b200/variants/,
# selected with TECDSA_B200_LIB) is recorded in profiles/r02_nadic_unroll_ab.md
