#!/bin/bash
# lean phase launches built as a variant library (tools/experiments/phase_launch.patch applied to a copy of csrc): parity + A/B against the product library
out=gpurun_out; mkdir -p $out
n=${1:-2}
export OFFTB_LIB=${PHASE_LIB:?path of the variant library built from the patched copy}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "plan_matches or fixture or raw_maps" 2>&1 | tail -2
timeout 200 $TR --master-port 29631 tests/mgpu_worker.py > $out/mgpu_parity_$n.log 2>&1; echo "parity rc=$?"
grep "FAIL\|MGPU\|rror" $out/mgpu_parity_$n.log | grep -v "^\[rank[1-9]" | tail -4
for cfg in "1 64 3" "1 32 3" "1 16 3" "0 64 3"; do set -- $cfg
OFFTB_PHASE_LAUNCH=$1 timeout 120 $TR --master-port 29640 bench.py --gpus $n --steps 10 --warmup 3 --no-e2e --T2 $2 --W2 $3 > $out/bench_n${n}_pl.log 2>$out/bench_n${n}_pl.err; echo "bench phase_launch=$1 T2=$2 W2=$3 rc=$?"
grep '^{"metric' $out/bench_n${n}_pl.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('   ', d['value'], d['ms_per_step'], d['ms_min'], d['gpu_launches'], d['parseval_rel_err'], {k:(v['ms_per_step']) for k,v in d['roofline']['passes'].items()})
"
done
