#!/bin/bash
# T2 / W2 / writer-share variants of the headline multi-GPU configuration (short runs: no gate, no e2e, no extras)
n=${1:-8}; tag=${2:-r02}
out=gpurun_out; mkdir -p $out
export OFFTB_FLAG_TIMEOUT_S=30
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
run() { # name args env...
  name=$1; shift; xargs=$1; shift
  env "$@" timeout 300 $TR --master-port 29632 bench.py --gpus $n --steps 10 --warmup 3 --no-gate --no-e2e --no-extra $xargs > $out/sweep_n${n}_${name}_$tag.log 2>&1
  grep '^{"metric' $out/sweep_n${n}_${name}_$tag.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']; print('$name', 'ms', d['ms_per_step'], 'min', d['ms_min'], {k:v['ms_per_step'] for k,v in r['passes'].items()}, (d.get('exchange') or {}).get('GBps_per_direction'), 'parity', d['parity']['rel_l2'])
"
}
run t32w3 "--T2 32 --W2 3" A=1
run t128w3 "--T2 128 --W2 3" A=1
run t64w2 "--T2 64 --W2 2" A=1
run t64w3s60 "--T2 64 --W2 3" OFFTB_WRITER_SHARE=60
run t32w5 "--T2 32 --W2 5" A=1
