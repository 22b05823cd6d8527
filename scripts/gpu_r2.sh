#!/bin/bash
# One gpurun call of round 2: smoke, GPU parity tests, bench lines.  usage: scripts/gpu_r2.sh <tag> [tests|notests|TESTFILES] [e4] [mixed]
TAG=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; RC=$?; echo "smoke rc=$RC" >> gpurun_out/${TAG}_smoke.log
tail -4 gpurun_out/${TAG}_smoke.log | cut -c1-400
if [ "$2" == "tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
  tail -15 gpurun_out/${TAG}_pytest.log
elif [ "$2" != "notests" ] && [ -n "$2" ]; then
  timeout 1500 python -m pytest $2 -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
  tail -15 gpurun_out/${TAG}_pytest.log
fi
ZLB_V4_TRACE=1 timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json | cut -c1-3000; grep "v4 phases" gpurun_out/${TAG}_bench.err | tail -1; tail -2 gpurun_out/${TAG}_bench.err | cut -c1-300
if [ "$3" == "e4" ] || [ "$4" == "e4" ]; then
  ZLB_V4_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --level 4 --no-decode > gpurun_out/${TAG}_bench_e4.json 2> gpurun_out/${TAG}_bench_e4.err; echo "bench e4 rc=$?"
  cat gpurun_out/${TAG}_bench_e4.json | cut -c1-2500; grep "v4 phases" gpurun_out/${TAG}_bench_e4.err | tail -1
fi
if [ "$3" == "mixed" ] || [ "$4" == "mixed" ]; then
  ZLB_V4_TRACE=1 timeout 900 python bench.py --steps 2 --warmup 3 --level 2 --corpus mixed --size-mb 200 --no-decode > gpurun_out/${TAG}_bench_mixed.json 2> gpurun_out/${TAG}_bench_mixed.err; echo "bench mixed rc=$?"
  cat gpurun_out/${TAG}_bench_mixed.json | cut -c1-2500; grep "v4 phases" gpurun_out/${TAG}_bench_mixed.err | tail -1
fi
