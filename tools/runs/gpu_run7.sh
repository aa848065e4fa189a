#!/bin/bash
# Round-2 evidence run on one B200 (through gpurun): parity suite, launch list of the bench command, ncu --set full captures of the
# two headline kernels, lane-group sweep of the modexp kernel, compute-sanitizer over every kernel family.  Outputs: gpurun_out/.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q --tb=short -x -p no:cacheprovider 2>&1 | tail -12 > $O/r02_t7.log
tail -4 $O/r02_t7.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r02_bench_under_ncu.json 2> $O/r02_bench_under_ncu.err
ncu --set full --clock-control none --import-source on -k regex:nadic_jobs_kernel -s 6 -c 1 -o $O/r02_nadic \
    python tools/offline_throughput.py 2048 > $O/r02_ncu_nadic.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:modexp_kernel -s 1 -c 1 -o $O/r02_modexp \
    python tools/prof_modexp.py 2048 4 65536 2 > $O/r02_ncu_modexp.log 2>&1
: > $O/r02_tpi_sweep.log
for cfg in "2048 4" "2048 8" "2048 16" "2048 32" "1024 4" "1024 8" "1024 16" "4096 8" "4096 16" "4096 32"; do
    set -- $cfg
    echo "bits=$1 tpi=$2" >> $O/r02_tpi_sweep.log
    python tools/prof_modexp.py $1 $2 65536 2 2>&1 | tail -1 >> $O/r02_tpi_sweep.log
done
: > $O/r02_san_summary.log
for tool in memcheck racecheck synccheck; do
    timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_small.py > $O/r02_san_$tool.log 2>&1
    echo "$tool rc=$?" >> $O/r02_san_summary.log
done
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_small.py split > $O/r02_san_racecheck_split.log 2>&1
echo "racecheck split rc=$?" >> $O/r02_san_summary.log
cat $O/r02_san_summary.log
ls -la $O/*.ncu-rep
