#!/bin/bash
# Smoke + bench with the shipped library, then the same bench with the profiling build (libzling_prof.so, built with
# -DZL_V4_PROFILE=1: per-warp timers inside the parse kernel).  usage: scripts/gpu_prof.sh <tag> [e4]
TAG=${1:-prof}
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
tail -2 gpurun_out/${TAG}_smoke.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
ZLB_V4_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-decode > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/${TAG}_bench.json; python -c "import json;d=json.load(open('gpurun_out/${TAG}_bench.json'));print(d['kernel_ms'], d['parse_counters'])"; grep "v4 phases" gpurun_out/${TAG}_bench.err | tail -1
if [ "$2" == "e4" ]; then
  ZLB_V4_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --level 4 --no-decode > gpurun_out/${TAG}_bench_e4.json 2> gpurun_out/${TAG}_bench_e4.err; echo "bench e4 rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/${TAG}_bench_e4.json'));print(d['value'], d['kernel_ms'], d['parse_counters'])"; grep "v4 phases" gpurun_out/${TAG}_bench_e4.err | tail -1
fi
if [ -f libzling_b200/libzling_prof.so ]; then
  cp libzling_b200/libzling_prof.so libzling_b200/libzling.so
  ZLB_V4_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 2 --no-decode --skip-parity > gpurun_out/${TAG}_prof.json 2> gpurun_out/${TAG}_prof.err; echo "prof rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/${TAG}_prof.json'));print(d['value'], d['kernel_ms'])"; grep "v4 phases" gpurun_out/${TAG}_prof.err | tail -1
fi
