#!/bin/bash
# multi-GPU session: the full bench line (parity gate included), writer-share variants, the reference arm.  usage: bash tools/gpu_multi2.sh N tag
n=${1:-2}; tag=${2:-r02}
out=gpurun_out; mkdir -p $out
export OFFTB_FLAG_TIMEOUT_S=20
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
run_bench() { # name extra-args env...
  name=$1; shift; xargs=$1; shift
  env "$@" timeout 600 $TR --master-port 29632 bench.py --gpus $n --steps 10 --warmup 3 $xargs > $out/bench_n${n}_${name}_$tag.log 2>&1; echo "bench $name rc=$?"
  grep '^{"metric' $out/bench_n${n}_${name}_$tag.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']; print('  ms', d['ms_per_step'], 'min', d['ms_min'], {k:(v['ms_per_step'],v['GBps']) for k,v in r['passes'].items()}, (d.get('exchange') or {}).get('GBps_per_direction'), 'parity', d.get('parity',{}).get('rel_l2'), [ (c['process_grid'], c['forward_vs_numpy'], c['round_trip']) for c in (d.get('parity',{}).get('cases') or [])], 'e2e', (d.get('e2e') or {}).get('ms_per_step'), d.get('invalid'))
"
  grep -E "Error|error|timed out|Traceback" $out/bench_n${n}_${name}_$tag.log | tail -3
}
run_bench full "" A=1
run_bench sh60 "--no-e2e --no-gate" OFFTB_WRITER_SHARE=60
run_bench sh70 "--no-e2e --no-gate" OFFTB_WRITER_SHARE=70
run_bench sh40 "--no-e2e --no-gate" OFFTB_WRITER_SHARE=40
run_bench ov0pdl "--no-e2e --no-gate" OFFTB_OVERLAP=0
timeout 900 python bench.py --impl reference --gpus $n --steps 2 --warmup 1 > $out/bench_ref_n${n}_$tag.log 2>&1; echo "ref rc=$?"; cut -c1-700 $out/bench_ref_n${n}_$tag.log
free -g | head -2; nproc
