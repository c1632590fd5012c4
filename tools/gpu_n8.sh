#!/bin/bash
# the 8-GPU session: multi-process parity, then the full bench line (parity gate, e2e, configs[3] and configs[4] as extra_configs)
n=${1:-8}; tag=${2:-r02}
out=gpurun_out; mkdir -p $out
export OFFTB_FLAG_TIMEOUT_S=30
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29631 tests/mgpu_worker.py > $out/mgpu_parity_${n}_$tag.log 2>&1; echo "parity rc=$?"
grep -E "^ok|^FAIL|MGPU" $out/mgpu_parity_${n}_$tag.log | cut -c1-220
timeout 900 $TR --master-port 29632 bench.py --gpus $n --steps 20 --warmup 3 > $out/bench_n${n}_$tag.log 2>&1; echo "bench rc=$?"
grep '^{"metric' $out/bench_n${n}_$tag.log > $out/bench_n${n}_$tag.json
python - <<PY
import json
d=json.loads(open("$out/bench_n${n}_$tag.json").read())
r=d['roofline']
print('ms', d['ms_per_step'], 'min', d['ms_min'], 'GFLOP/s', d['value'], {k:(v['ms_per_step'],v['GBps']) for k,v in r['passes'].items()})
print('exchange', d.get('exchange'))
print('parity', d['parity']['rel_l2'], [(c['process_grid'], c['S'], c['forward_vs_numpy'], c['round_trip']) for c in d['parity']['cases'] or []])
print('e2e', d.get('e2e'))
for e in d.get('extra_configs') or []: print('extra', json.dumps(e)[:900])
print('invalid', d.get('invalid'))
PY
grep -E "Error|error|timed out|Traceback" $out/bench_n${n}_$tag.log | tail -5
nvidia-smi topo -m 2>/dev/null | head -12; nproc; free -g | head -2
