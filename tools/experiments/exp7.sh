#!/bin/bash
out=gpurun_out; mkdir -p $out
n=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29631 tests/mgpu_worker.py > $out/mgpu_parity_$n.log 2>&1; echo "parity rc=$?"
grep "MGPU\|FAIL" $out/mgpu_parity_$n.log
run() {
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus $n --steps 10 --warmup 3 --no-e2e $BARGS > $out/bench_n${n}_$tag.log 2>&1; echo "bench $tag ($* $BARGS) rc=$?"
  grep '^{"metric' $out/bench_n${n}_$tag.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('   ', d['value'], d['ms_per_step'], {k:(v['ms_per_step'],v['GBps']) for k,v in d['roofline']['passes'].items()})
"
}
BARGS="--T2 64 --S 1" run ov0_s1 OFFTB_OVERLAP=0
BARGS="--T2 64 --S 1" run ov1_s1 OFFTB_OVERLAP=1
BARGS="--T2 64 --S 0" run ov0_s0 OFFTB_OVERLAP=0
BARGS="--T2 64 --S 0" run ov1_s0 OFFTB_OVERLAP=1
BARGS="--T2 32 --S 0 --W2 3" run ov1_s0_t32 OFFTB_OVERLAP=1
BARGS="--T2 64 --S 0" run ov1_s0_c12 OFFTB_OVERLAP=1 OFFTB_CAP_W=1 OFFTB_CAP_R=2
