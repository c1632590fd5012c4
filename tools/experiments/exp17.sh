#!/bin/bash
# phase launches: 1-GPU parity (classic + emulated ranks), then 2-GPU parity and bench with and without phase launches
out=gpurun_out; mkdir -p $out
n=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 200 $TR --master-port 29631 tests/mgpu_worker.py > $out/mgpu_parity_$n.log 2>&1; echo "parity rc=$?"
grep "FAIL\|MGPU\|tuning\|rror" $out/mgpu_parity_$n.log | grep -v "^\[rank[1-9]" | tail -6
for pl in 1 0; do
OFFTB_PHASE_LAUNCH=$pl timeout 120 $TR --master-port 2964$pl bench.py --gpus $n --steps 10 --warmup 3 --no-e2e > $out/bench_n${n}_pl$pl.log 2>&1; echo "bench phase_launch=$pl rc=$?"
grep '^{"metric' $out/bench_n${n}_pl$pl.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('   ', d['value'], d['ms_per_step'], d['gpu_launches'], d['parseval_rel_err'], {k:(v['ms_per_step']) for k,v in d['roofline']['passes'].items()})
"
grep -i "error\|trap\|fail" $out/bench_n${n}_pl$pl.log | head -3
done
