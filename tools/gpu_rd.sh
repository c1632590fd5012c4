#!/bin/bash
n=${1:-2}; tag=${2:-r02}
out=gpurun_out; mkdir -p $out
export OFFTB_FLAG_TIMEOUT_S=20
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
run() { name=$1; shift; xargs=$1; shift
  env "$@" timeout 300 $TR --master-port 29632 bench.py --gpus $n --steps 10 --warmup 3 --no-gate --no-e2e --no-extra $xargs > $out/rd_n${n}_${name}_$tag.log 2>&1
  grep '^{"metric' $out/rd_n${n}_${name}_$tag.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']; print('$name', 'ms', d['ms_per_step'], 'min', d['ms_min'], {k:v['ms_per_step'] for k,v in r['passes'].items()}, (d.get('exchange') or {}).get('GBps_per_direction'), 'parity', d['parity']['rel_l2'])
"; grep -E "Error|error|timed out" $out/rd_n${n}_${name}_$tag.log | tail -2; }
run rd0 "" OFFTB_READER_DEPTH=0
run rdauto "" A=1
run rd2 "" OFFTB_READER_DEPTH=2
run rdauto_s40 "" OFFTB_WRITER_SHARE=40
