#!/bin/bash
out=gpurun_out; mkdir -p $out
n=${1:-2}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29631 tests/mgpu_worker.py > $out/mgpu_parity_$n.log 2>&1; echo "parity rc=$?"
grep "MGPU\|FAIL" $out/mgpu_parity_$n.log
for cz in 4 8; do for T2 in 0 64; do
OFFTB_CZ=$cz timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus $n --steps 10 --warmup 3 --no-e2e --T2 $T2 > $out/bench_n${n}_cz$cz_$T2.log 2>&1; echo "bench cz=$cz T2=$T2 rc=$?"
grep '^{"metric' $out/bench_n${n}_cz$cz_$T2.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['value'], d['ms_per_step'], d['config']['tunables'], {k:(v['ms_per_step'],v['GBps']) for k,v in d['roofline']['passes'].items()})
"
done; done
