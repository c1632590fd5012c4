#!/bin/bash
# 2-GPU check of the half-width strided tiles (2048-point transforms on both sides of the exchange) and of the extra-config path
n=${1:-2}; tag=${2:-r02}
out=gpurun_out; mkdir -p $out
export OFFTB_FLAG_TIMEOUT_S=20
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
run_bench() { # name extra-args env...
  name=$1; shift; xargs=$1; shift
  env "$@" timeout 600 $TR --master-port 29632 bench.py --gpus $n --steps 8 --warmup 3 $xargs > $out/bench_n${n}_${name}_$tag.log 2>&1; echo "bench $name rc=$?"
  grep '^{"metric' $out/bench_n${n}_${name}_$tag.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']; print('  ms', d['ms_per_step'], 'min', d['ms_min'], {k:(v['ms_per_step'],v['GBps']) for k,v in r['passes'].items()}, (d.get('exchange') or {}).get('GBps_per_direction'), 'parity', d.get('parity',{}).get('rel_l2'), d.get('invalid'))
    for e in d.get('extra_configs') or []: print('  extra', json.dumps(e)[:600])
"
  grep -E "Error|error|timed out|Traceback" $out/bench_n${n}_${name}_$tag.log | tail -3
}
G="--grid 2048x2048x64 --T2 8 --W2 3 --no-gate --no-e2e --no-extra"
run_bench wide "$G" OFFTB_NARROW=0
run_bench narrow "$G" OFFTB_NARROW=1
run_bench narrow65 "$G" OFFTB_NARROW=1 OFFTB_WRITER_SHARE=65
run_bench extratest "--no-e2e --no-gate --extra-test --steps 3" A=1
