#!/bin/bash
out=gpurun_out; mkdir -p $out
n=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
timeout 100 python tools/kbench.py 512 64 --clogs -1 2>&1 | grep -v "^P1\|^M1"
timeout 100 python tools/kbench.py 1024 64 --clogs -1 2>&1 | grep -v "^P1\|^M1"
timeout 200 $TR --master-port 29631 tests/mgpu_worker.py > $out/mgpu_parity_$n.log 2>&1; echo "parity rc=$?"
grep "FAIL\|MGPU\|rror" $out/mgpu_parity_$n.log | grep -v "^\[rank[1-9]" | tail -4
for cfg in "1 64 3" "1 32 3" "1 16 3" "0 64 3"; do set -- $cfg
OFFTB_PHASE_LAUNCH=$1 timeout 120 $TR --master-port 29640 bench.py --gpus $n --steps 10 --warmup 3 --no-e2e --T2 $2 --W2 $3 > $out/bench_n${n}_pl.log 2>&1; echo "bench phase_launch=$1 T2=$2 W2=$3 rc=$?"
grep '^{"metric' $out/bench_n${n}_pl.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('   ', d['value'], d['ms_per_step'], d['gpu_launches'], d['parseval_rel_err'], {k:(v['ms_per_step']) for k,v in d['roofline']['passes'].items()})
"
done
