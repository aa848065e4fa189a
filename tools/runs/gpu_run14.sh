#!/bin/bash
# Round-2 run 14 on TWO B200s of one box (gpurun --gpus 2): the library's NCCL gather against torch.distributed, then bench.py at N = 2.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_nccl_gather_gpu.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -15 > $O/r02_t14.log
tail -15 $O/r02_t14.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline \
    > $O/r02_bench_n2.json 2> $O/r02_bench_n2.err
tail -c 1800 $O/r02_bench_n2.json; echo; tail -3 $O/r02_bench_n2.err
