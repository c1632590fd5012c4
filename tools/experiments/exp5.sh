#!/bin/bash
out=gpurun_out; mkdir -p $out
n=${1:-2}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29631 tests/mgpu_worker.py > $out/mgpu_parity_$n.log 2>&1; echo "parity rc=$?"
grep -v "^P1\|^M1\|^@\|^\*\*\*\|OMP_NUM\|^$" $out/mgpu_parity_$n.log | tail -25
for mode in fused nccl; do
OFFTB_EXCHANGE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus $n --steps 10 --warmup 3 --no-e2e > $out/bench_n${n}_$mode.log 2>&1; echo "bench $mode rc=$?"
grep -v "^P1\|^M1\|^@\|^\*\*\*\|OMP_NUM\|^$" $out/bench_n${n}_$mode.log | tail -4
done
