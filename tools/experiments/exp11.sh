#!/bin/bash
out=gpurun_out; mkdir -p $out
n=${1:-4}
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29631 tests/mgpu_worker.py > $out/mgpu_parity_$n.log 2>&1; echo "parity rc=$?"
grep "ok \|FAIL\|MGPU\|Error\|error" $out/mgpu_parity_$n.log | grep -v "^\[rank[1-9]" | tail -20
