#!/bin/bash
# A/B experiments on the 100 MB e0 workload: each line prints parse ms and resolver cycles per token
mkdir -p gpurun_out
run() { echo "== $1"; env $1 ZLB_V3_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 1 2>gpurun_out/exp.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); p = d['parse_counters']
print('value', d['value'], 'exact', d['config']['bit_exact_vs_cpu_reference'], 'kernel_ms', d['kernel_ms'], 'cyc/token resolve %.0f spec-per-window %.0f' % (p['cyc_resolve'] / p['tokens'], p['cyc_spec'] / max(p['windows'], 1)))"; grep "^v3:" gpurun_out/exp.err | tail -1; }
for v in "$@"; do run "$v"; done
