#!/bin/bash
out=gpurun_out; mkdir -p $out
n=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29631 tests/mgpu_worker.py > $out/mgpu_parity_$n.log 2>&1; echo "parity rc=$?"
grep "ok \|FAIL\|MGPU" $out/mgpu_parity_$n.log | tail -12
timeout 200 $TR --master-port 29633 tools/run_config.py --grid 512 --p1 2 --steps 3 > $out/cfg_pencil512_$n.log 2>&1; echo "pencil rc=$?"; grep '^{' $out/cfg_pencil512_$n.log
timeout 300 $TR --master-port 29634 tools/run_config.py --grid 512x256x128 --bits 32 --oned 1 --sweep --steps 3 > $out/cfg_sweep_$n.log 2>&1; echo "sweep rc=$?"; grep '^{' $out/cfg_sweep_$n.log | cut -c1-1500
timeout 200 $TR --master-port 29635 tools/run_config.py --grid 512x256x128 --bits 32 --oned 1 --tune 10 --steps 3 > $out/cfg_tune_$n.log 2>&1; echo "tune rc=$?"; grep '^{\|@ BEST' $out/cfg_tune_$n.log | cut -c1-600
