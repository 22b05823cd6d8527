#!/bin/bash
# One gpurun call of round 2: smoke, GPU parity tests, bench lines.  usage: scripts/gpu_r2.sh <tag> [tests|notests] [extra]
TAG=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; RC=$?; echo "smoke rc=$RC" >> gpurun_out/${TAG}_smoke.log
tail -4 gpurun_out/${TAG}_smoke.log
if [ $RC -ne 0 ]; then echo "smoke failed: trying the v3 parse to localise"; ZLB_PARSE=3 timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2; fi
if [ "$2" != "notests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
  tail -15 gpurun_out/${TAG}_pytest.log
fi
ZLB_V4_TRACE=1 timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json; grep "v4 phases" gpurun_out/${TAG}_bench.err | tail -2; tail -3 gpurun_out/${TAG}_bench.err
if [ "$3" == "e4" ]; then
  ZLB_V4_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --level 4 --no-decode > gpurun_out/${TAG}_bench_e4.json 2> gpurun_out/${TAG}_bench_e4.err; echo "bench e4 rc=$?"
  cat gpurun_out/${TAG}_bench_e4.json | cut -c1-2500; grep "v4 phases" gpurun_out/${TAG}_bench_e4.err | tail -1; tail -3 gpurun_out/${TAG}_bench_e4.err
fi
