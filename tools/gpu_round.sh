#!/bin/bash
# one GPU session: parity tests, the bench line, the ncu launch list and one full capture of the bench's kernels.
# usage (under gpurun): bash tools/gpu_round.sh <tag>
set -u
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu_$tag.log
tail -3 $out/pytest_gpu_$tag.log
python __graft_entry__.py smoke > $out/smoke_$tag.log 2>&1; tail -1 $out/smoke_$tag.log
python bench.py --steps 20 --warmup 3 > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"; cat $out/bench_$tag.json
python tools/kbench.py 512 64 > $out/kbench_512_$tag.log 2>&1; tail -30 $out/kbench_512_$tag.log
BENCH_SHORT="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
$BENCH_SHORT > $out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/launches_$tag.csv $BENCH_SHORT > $out/ncu_launches_$tag.log 2>&1
$BENCH_SHORT > $out/plain2_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fft_kernel -s 9 -c 3 -o $out/prof_bench_$tag $BENCH_SHORT > $out/ncu_full_$tag.log 2>&1
ls -la $out | tail -20
