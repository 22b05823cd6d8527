#!/bin/bash
# The remaining lines of the round: configs[3] on ONE GPU (the N = 1 point of the strong-scaling series), the batch API at one
# stream per SM, the e0..e4 table through the reference's unmodified CLI.  usage: scripts/gpu_last.sh <tag>
TAG=${1:-r2l}
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "cpu", d["cpu_baseline"]["value"], d.get("kernel_ms"), (d.get("parse_counters") or {}).get("reparsed_blocks"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
}
timeout 900 python bench.py --corpus mixed --size-mb 1000 --level 2 --steps 2 --warmup 2 --no-decode > gpurun_out/${TAG}_stream1g_n1.json 2> gpurun_out/${TAG}_stream1g_n1.err; echo "1g n1 rc=$?"; show gpurun_out/${TAG}_stream1g_n1.json
timeout 600 python scripts/cli_table.py 100 gpurun_out/${TAG}_cli_table.md > gpurun_out/${TAG}_cli_table.log 2>&1; echo "cli table rc=$?"; cat gpurun_out/${TAG}_cli_table.md
timeout 900 python bench.py --streams 148 --size-mb 16.7 --steps 2 --warmup 2 > gpurun_out/${TAG}_batch148_enc.json 2> gpurun_out/${TAG}_batch148_enc.err; echo "batch encode rc=$?"; show gpurun_out/${TAG}_batch148_enc.json; tail -1 gpurun_out/${TAG}_batch148_enc.err | cut -c1-300
timeout 900 python bench.py --mode decode --streams 148 --size-mb 16.7 --steps 2 --warmup 2 > gpurun_out/${TAG}_batch148_dec.json 2> gpurun_out/${TAG}_batch148_dec.err; echo "batch decode rc=$?"; show gpurun_out/${TAG}_batch148_dec.json; tail -1 gpurun_out/${TAG}_batch148_dec.err | cut -c1-300
