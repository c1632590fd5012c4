#!/bin/bash
out=gpurun_out; mkdir -p $out
n=${1:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29631 tests/mgpu_worker.py > $out/mgpu_parity_$n.log 2>&1; echo "parity rc=$?"
grep "MGPU\|FAIL\|ok " $out/mgpu_parity_$n.log | tail -12
run() {
  tag=$1; shift
  env "$@" timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus $n --steps 10 --warmup 3 --no-e2e $BARGS > $out/bench_n${n}_$tag.log 2>&1; echo "bench $tag ($* $BARGS) rc=$?"
  grep '^{"metric' $out/bench_n${n}_$tag.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('   ', d['value'], d['ms_per_step'], d['ms_min'], {k:(v['ms_per_step'],v['GBps']) for k,v in d['roofline']['passes'].items()})
"
}
BARGS="--T2 64 --S 0" run ov0 OFFTB_OVERLAP=0
BARGS="--T2 64 --S 0" run ov50 OFFTB_WRITER_SHARE=50
BARGS="--T2 64 --S 0" run ov30 OFFTB_WRITER_SHARE=30
BARGS="--T2 32 --S 0 --W2 3" run ov50t32 OFFTB_WRITER_SHARE=50
BARGS="--T2 64 --S 1" run ov50s1 OFFTB_WRITER_SHARE=50
BARGS="--T2 64 --S 0" run nccl OFFTB_EXCHANGE=nccl
