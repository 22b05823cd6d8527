#!/bin/bash
# ncu --set full capture of the parse kernel on ONE 16 MiB block (the kernel is one CTA per block: one block is representative),
# with source correlation.  usage: scripts/gpu_ncu_parse.sh <tag> [level]
TAG=${1:-r2}; LEVEL=${2:-0}
mkdir -p gpurun_out
timeout 800 ncu --set full --clock-control none --import-source on -k regex:zl_rolz_parse_v4 -c 1 -f -o gpurun_out/${TAG}_parse_v4_e${LEVEL} \
    python bench.py --steps 1 --warmup 0 --size-mb 16.7 --level ${LEVEL} --no-decode > gpurun_out/${TAG}_ncu_parse.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/${TAG}_ncu_parse.log; ls -la gpurun_out/${TAG}_parse_v4_e${LEVEL}.ncu-rep
