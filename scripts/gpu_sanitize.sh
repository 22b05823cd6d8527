#!/bin/bash
# compute-sanitizer over the smoke workload (1 MiB at e0 and e4 + decode): memcheck, then racecheck (shared-memory hazards inside a
# CTA; DSMEM traffic between the CTAs of a cluster is ordered by cluster barriers and not seen by the tool).  usage: scripts/gpu_sanitize.sh <tag>
TAG=${1:-san}
mkdir -p gpurun_out
for TOOL in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $TOOL --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_${TOOL}.log 2>&1; echo "$TOOL rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|Error|hazard" gpurun_out/${TAG}_${TOOL}.log | head -12 | cut -c1-220
done
