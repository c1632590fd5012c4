#!/bin/bash
# the 8-GPU session of a round: multi-process parity, the bench line, the T/W sweep of configs[2], configs[3] and configs[4]
out=gpurun_out; mkdir -p $out
n=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29631 tests/mgpu_worker.py > $out/mgpu_parity_$n.log 2>&1; echo "parity rc=$?"
grep "FAIL\|MGPU\|tuning" $out/mgpu_parity_$n.log | tail -5
timeout 240 $TR --master-port 29632 bench.py --gpus $n --steps 20 --warmup 3 > $out/bench_n${n}_final.log 2>&1; echo "bench rc=$?"
grep '^{"metric' $out/bench_n${n}_final.log | cut -c1-400
timeout 240 $TR --master-port 29633 tools/run_config.py --grid 1024 --oned 1 --sweep --steps 3 > $out/cfg3_sweep_$n.log 2>&1; echo "cfg3 sweep rc=$?"
grep '^{' $out/cfg3_sweep_$n.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('main', d['main']); print('best', d.get('best')); print(sorted([(s['ms_min'],s['T'],s['W']) for s in d['sweep'] if 'ms_min' in s])[:8])
"
timeout 300 $TR --master-port 29634 tools/run_config.py --grid 2048 --p1 2 --steps 2 > $out/cfg4_pencil2048_$n.log 2>&1; echo "cfg4 rc=$?"; grep '^{' $out/cfg4_pencil2048_$n.log
timeout 200 $TR --master-port 29635 tools/run_config.py --grid 2048x1024x512 --bits 32 --oned 1 --tune 12 --steps 3 > $out/cfg5_tune_$n.log 2>&1; echo "cfg5 rc=$?"; grep '^{\|@ BEST' $out/cfg5_tune_$n.log | cut -c1-700
