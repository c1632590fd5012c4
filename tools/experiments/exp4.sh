#!/bin/bash
# 2-GPU bring-up: NCCL parity, then the bench at N=2
out=gpurun_out; mkdir -p $out
n=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29631 tests/mgpu_worker.py > $out/mgpu_parity_$n.log 2>&1; echo "parity rc=$?"
grep -v "^P1\|^M1\|^@" $out/mgpu_parity_$n.log | tail -25
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus $n --steps 10 --warmup 3 > $out/bench_n$n.log 2>&1; echo "bench rc=$?"
grep -v "^P1\|^M1\|^@" $out/bench_n$n.log | tail -8
