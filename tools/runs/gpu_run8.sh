#!/bin/bash
# Round-2 evidence run 8 on one B200: persistent modexp check, ncu --set full of the dominant N-adic launch and of the persistent
# modexp kernel, then the bench lines (own arm with cpu_baseline, reference arm).  Outputs: gpurun_out/.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_modexp_gpu.py -m gpu -q --tb=short -x -p no:cacheprovider 2>&1 | tail -5 > $O/r02_t8.log
tail -3 $O/r02_t8.log
python tools/prof_modexp.py 2048 4 65536 3 2>&1 | tail -3 > $O/r02_modexp_persistent.log
python tools/prof_modexp.py 4096 8 65536 2 2>&1 | tail -1 >> $O/r02_modexp_persistent.log
python tools/prof_modexp.py 1024 4 65536 2 2>&1 | tail -1 >> $O/r02_modexp_persistent.log
cat $O/r02_modexp_persistent.log
TECDSA_SPLIT=0 ncu --set full --clock-control none --import-source on -k regex:nadic_jobs_kernel -s 8 -c 1 -o $O/r02_nadic64 -f \
    python tools/offline_throughput.py 4096 > $O/r02_ncu_nadic64.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:modexp_kernel -s 1 -c 1 -o $O/r02_modexp_persistent -f \
    python tools/prof_modexp.py 2048 4 65536 2 > $O/r02_ncu_modexp_persistent.log 2>&1
python bench.py --impl reference > $O/r02_ref_n1.json 2> $O/r02_ref_n1.err
python bench.py > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err
tail -c 3000 $O/r02_ref_n1.json; echo; tail -c 6000 $O/r02_bench_n1.json
ls -la $O/*.ncu-rep
