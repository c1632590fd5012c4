#!/bin/bash
# one multi-GPU session: multi-process parity in every exchange mode, then bench lines.  usage (under gpurun --gpus N): bash tools/gpu_multi.sh N tag
n=${1:-2}; tag=${2:-r02}
out=gpurun_out; mkdir -p $out
export OFFTB_FLAG_TIMEOUT_S=20
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
run_parity() { # name env...
  name=$1; shift
  env "$@" timeout 400 $TR --master-port 29631 tests/mgpu_worker.py > $out/mgpu_parity_${n}_${name}_$tag.log 2>&1; echo "parity $name rc=$?"
  grep -E "FAIL|MGPU|tuning|Error|error" $out/mgpu_parity_${n}_${name}_$tag.log | tail -4
}
run_bench() { # name env...
  name=$1; shift
  env "$@" timeout 400 $TR --master-port 29632 bench.py --gpus $n --steps 10 --warmup 3 --no-e2e > $out/bench_n${n}_${name}_$tag.log 2>&1; echo "bench $name rc=$?"
  grep '^{"metric' $out/bench_n${n}_${name}_$tag.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']; print('  ms', d['ms_per_step'], 'min', d['ms_min'], {k:(v['ms_per_step'],v['GBps']) for k,v in r['passes'].items()}, (r.get('exchange') or d.get('exchange') or {}).get('GBps_per_direction'), d.get('parity'))
"
  grep -E "Error|error|timed out" $out/bench_n${n}_${name}_$tag.log | tail -3
}
run_parity default A=1
run_bench default A=1
run_bench nobulk OFFTB_BULK=0
run_bench nopdl OFFTB_PDL=0
run_bench old OFFTB_BULK=0 OFFTB_PDL=0
run_parity direct OFFTB_BULK=0 OFFTB_PDL=0 OFFTB_OVERLAP=0
run_parity nccl OFFTB_EXCHANGE=nccl
