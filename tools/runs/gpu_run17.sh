#!/bin/bash
# Round-2 run 17 on N GPUs of one box (gpurun --gpus N): the library's NCCL gather test and both bench arms as the driver launches them.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out
N=$(nvidia-smi -L | wc -l)
mkdir -p $O
python -m pytest tests/test_nccl_gather_gpu.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -5 > $O/r02_t17.log
tail -3 $O/r02_t17.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --impl reference --gpus $N --steps 3 --warmup 1 \
    > $O/r02_ref_n$N.json 2> $O/r02_ref_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline \
    > $O/r02_bench_n$N.json 2> $O/r02_bench_n$N.err
tail -c 400 $O/r02_ref_n$N.json; echo; head -c 600 $O/r02_bench_n$N.json; echo; tail -3 $O/r02_bench_n$N.err
