#!/bin/bash
# A/B of the parse kernel's cluster size (helpers per block) on the default workload.  usage: scripts/gpu_cluster_ab.sh <tag>
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log | cut -c1-300
for CL in 8; do
  ZLB_PARSE_CLUSTER=$CL ZLB_V4_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-decode > gpurun_out/${TAG}_cl${CL}.json 2> gpurun_out/${TAG}_cl${CL}.err; echo "cl=$CL rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_cl${CL}.json').read().strip().splitlines()[-1]); p=d['parse_counters']; w=p['windows']
    print('cl=${CL} value',d['value'],'kernel_ms',d['kernel_ms'],'per window spec %.0f rounds %.0f final %.0f'%(p['cyc_spec']/w,p['cyc_resolve']/w,p['cyc_final']/w))
except Exception as e: print('cl=${CL}', e)
PY
  grep "v4 phases" gpurun_out/${TAG}_cl${CL}.err | tail -1 | cut -c60-400; tail -2 gpurun_out/${TAG}_cl${CL}.err | cut -c1-300
done
ZLB_PARSE_CLUSTER=8 timeout 600 python bench.py --steps 2 --warmup 3 --no-decode --level 4 > gpurun_out/${TAG}_cl4_e4.json 2> gpurun_out/${TAG}_cl4_e4.err; echo "e4 cl=4 rc=$?"; cut -c1-400 gpurun_out/${TAG}_cl4_e4.json; tail -2 gpurun_out/${TAG}_cl4_e4.err | cut -c1-300
ZLB_PARSE_CLUSTER=4 timeout 600 python bench.py --steps 2 --warmup 3 --no-decode --level 4 > gpurun_out/${TAG}_cl4b_e4.json 2> gpurun_out/${TAG}_cl4b_e4.err; echo "e4 cl=4 rc=$?"; cut -c1-120 gpurun_out/${TAG}_cl4b_e4.json
python - <<PY
import json
for f in ('gpurun_out/${TAG}_cl4_e4.json','gpurun_out/${TAG}_cl4b_e4.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); p=d['parse_counters']; w=p['windows']
    print(f, d['value'], d['kernel_ms'], 'per window spec %.0f rounds %.0f final %.0f'%(p['cyc_spec']/w,p['cyc_resolve']/w,p['cyc_final']/w))
PY
