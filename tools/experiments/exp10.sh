#!/bin/bash
out=gpurun_out; mkdir -p $out
n=${1:-4}
timeout 600 python -m pytest tests/test_dropin.py tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -15
grep -h "ok \|FAIL\|MGPU" $out/*.log 2>/dev/null | tail -0
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus $n --steps 10 --warmup 3 > $out/bench_n${n}_final.log 2>&1; echo "bench rc=$?"
grep '^{"metric' $out/bench_n${n}_final.log
