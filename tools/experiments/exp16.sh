#!/bin/bash
# asymmetric overlap (SM partition) on 2 GPUs: slab 1 x p with 2048-point y rows; plus the parity worker
out=gpurun_out; mkdir -p $out
n=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29631 tests/mgpu_worker.py > $out/mgpu_parity_$n.log 2>&1; echo "parity rc=$?"
grep "FAIL\|MGPU\|tuning" $out/mgpu_parity_$n.log | tail -4
for ov in 0 1; do
OFFTB_OVERLAP=$ov timeout 100 $TR --master-port 2964$ov tools/run_config.py --grid 64x2048x2048 --p1 1 --oned 1 --steps 3 > $out/cfg_asym_ov$ov.log 2>&1; echo "asym overlap=$ov rc=$?"; grep '^{' $out/cfg_asym_ov$ov.log
done
OFFTB_WRITER_SM_SHARE=40 timeout 100 $TR --master-port 29643 tools/run_config.py --grid 64x2048x2048 --p1 1 --oned 1 --steps 3 > $out/cfg_asym_sh40.log 2>&1; echo "asym share40 rc=$?"; grep '^{' $out/cfg_asym_sh40.log
timeout 100 $TR --master-port 29644 tools/run_config.py --grid 2048x2048x64 --oned 1 --steps 3 > $out/cfg_asym_px1.log 2>&1; echo "asym px1 rc=$?"; grep '^{' $out/cfg_asym_px1.log
