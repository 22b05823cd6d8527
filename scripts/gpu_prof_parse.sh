#!/bin/bash
# source-level counters of the parse kernel on a 2-block workload (cheap: two ncu sections, one launch)
TAG=${1:-prof}
mkdir -p gpurun_out
timeout 900 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --clock-control none --import-source on -k regex:zl_rolz_parse_v3 -c 1 \
    -o gpurun_out/${TAG}_parse_v3 -f python bench.py --size-mb 33 --steps 1 --warmup 0 --skip-parity > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
ls -la gpurun_out | tail -4
