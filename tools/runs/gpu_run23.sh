#!/bin/bash
# Round-2 run 23 on one B200: final bench lines with the 4-blocks-per-SM defaults, launch list of the bench command, ncu --set full of the
# dominant N-adic launch in its new shape.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
python bench.py --impl reference > $O/r02_ref_n1.json 2> $O/r02_ref_n1.err
python bench.py > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err
head -c 400 $O/r02_bench_n1.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r02_bench_under_ncu.json 2> $O/r02_bench_under_ncu.err
TECDSA_SPLIT=0 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:nadic_jobs_kernel<.int.64' -s 8 -c 1 \
    -o $O/r02_nadic64_minb4 -f python tools/offline_throughput.py 4096 > $O/r02_ncu_nadic64.log 2>&1
ls -la $O/r02_nadic64_minb4.ncu-rep
python -m pytest tests -m gpu -q --tb=short -x -p no:cacheprovider 2>&1 | tail -4 > $O/r02_t23.log
tail -3 $O/r02_t23.log
