#!/bin/bash
# one 1-GPU session: parity tests, smoke, bench line, per-pass kernel rates.  usage (under gpurun): bash tools/gpu_n1.sh <tag>
set -u
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
export OFFTB_FLAG_TIMEOUT_S=20
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"
tail -5 $out/pytest_gpu_$tag.log
timeout 120 python __graft_entry__.py smoke > $out/smoke_$tag.log 2>&1; tail -1 $out/smoke_$tag.log
timeout 600 python bench.py --steps 20 --warmup 3 > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"; cut -c1-1500 $out/bench_$tag.json
timeout 300 python tools/kbench.py 512 64 --clogs -1 --plan > $out/kbench_512_$tag.log 2>&1; tail -12 $out/kbench_512_$tag.log
timeout 300 python tools/kbench.py 1024 64 --clogs -1 > $out/kbench_1024_$tag.log 2>&1; tail -8 $out/kbench_1024_$tag.log
