#!/bin/bash
# last call of the round: A/B of the producer count (window length), then full validation of the winner
mkdir -p gpurun_out
val() { python -c "import sys,json; d=json.loads(open('$1').read().strip().splitlines()[-1]); print(d['kernel_ms']['parse'])"; }
timeout 200 python bench.py --steps 2 --warmup 1 --no-decode > gpurun_out/fin_A.json 2> gpurun_out/fin_A.err; echo "A rc=$? parse_ms=$(val gpurun_out/fin_A.json)"
ZL_V3_PROD=256 python -m libzling_b200.build > /dev/null 2>&1
timeout 200 python bench.py --steps 2 --warmup 1 --no-decode > gpurun_out/fin_B.json 2> gpurun_out/fin_B.err; echo "B rc=$? parse_ms=$(val gpurun_out/fin_B.json)"
WIN=$(python -c "
import json
def v(p):
    try: return json.loads(open(p).read().strip().splitlines()[-1])['kernel_ms']['parse']
    except Exception: return 1e30
print('B' if v('gpurun_out/fin_B.json') < v('gpurun_out/fin_A.json') else 'A')")
echo "winner=$WIN"
if [ "$WIN" == "A" ]; then python -m libzling_b200.build > /dev/null 2>&1; fi
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fin_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/fin_smoke.log | cut -c1-200
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/fin_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/fin_pytest.log
timeout 120 python bench.py --steps 3 --warmup 3 > gpurun_out/fin_bench.json 2> gpurun_out/fin_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/fin_bench.json
