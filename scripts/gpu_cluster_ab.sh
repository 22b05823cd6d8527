#!/bin/bash
# A/B of the parse kernel's cluster size (helpers per block).  usage: scripts/gpu_cluster_ab.sh <tag> "<cluster sizes>" "<levels>"
TAG=${1:-r2}; CLS=${2:-"8 16"}; LVS=${3:-"0 4"}
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log | cut -c1-200
for LV in $LVS; do for CL in $CLS; do
  ZLB_PARSE_CLUSTER=$CL ZLB_V4_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-decode --level $LV > gpurun_out/${TAG}_cl${CL}_e${LV}.json 2> gpurun_out/${TAG}_cl${CL}_e${LV}.err; echo "cl=$CL level=$LV rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_cl${CL}_e${LV}.json').read().strip().splitlines()[-1]); p=d['parse_counters']; w=p['windows']
    print('cl=${CL} e${LV} value',d['value'],'kernel_ms',d['kernel_ms'],'per window spec %.0f rounds %.0f final %.0f'%(p['cyc_spec']/w,p['cyc_resolve']/w,p['cyc_final']/w))
except Exception as e: print('cl=${CL}', e)
PY
  grep "v4 phases" gpurun_out/${TAG}_cl${CL}_e${LV}.err | tail -1 | cut -c60-330; tail -1 gpurun_out/${TAG}_cl${CL}_e${LV}.err | cut -c1-300
done; done
