#!/bin/bash
# The bench lines besides the default one: every level e1..e3 on the 100 MB workload (e0/e4/mixed come from gpu_r2.sh), full-size decode,
# and the batch API (many independent streams per call).  usage: scripts/gpu_extras.sh <tag> [levels] [decode] [batch]
TAG=${1:-r2x}; shift
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "cpu", d["cpu_baseline"]["value"], d.get("kernel_ms"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
}
for what in "$@"; do
  case $what in
    levels) for LV in 1 2 3; do
              timeout 600 python bench.py --steps 2 --warmup 3 --level $LV --no-decode > gpurun_out/${TAG}_bench_e${LV}.json 2> gpurun_out/${TAG}_bench_e${LV}.err; echo "e$LV rc=$?"; show gpurun_out/${TAG}_bench_e${LV}.json
            done ;;
    decode) timeout 900 python bench.py --mode decode --steps 2 --warmup 1 > gpurun_out/${TAG}_decode.json 2> gpurun_out/${TAG}_decode.err; echo "decode rc=$?"; show gpurun_out/${TAG}_decode.json; tail -2 gpurun_out/${TAG}_decode.err | cut -c1-300 ;;
    batch)  timeout 900 python bench.py --streams 32 --size-mb 16.7 --steps 2 --warmup 2 > gpurun_out/${TAG}_batch_enc.json 2> gpurun_out/${TAG}_batch_enc.err; echo "batch encode rc=$?"; show gpurun_out/${TAG}_batch_enc.json; tail -2 gpurun_out/${TAG}_batch_enc.err | cut -c1-300
            timeout 900 python bench.py --mode decode --streams 48 --size-mb 16.7 --steps 2 --warmup 2 > gpurun_out/${TAG}_batch_dec.json 2> gpurun_out/${TAG}_batch_dec.err; echo "batch decode rc=$?"; show gpurun_out/${TAG}_batch_dec.json; tail -2 gpurun_out/${TAG}_batch_dec.err | cut -c1-300 ;;
  esac
done
