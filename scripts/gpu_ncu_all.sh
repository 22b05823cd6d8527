#!/bin/bash
# ncu --set full capture of every kernel of the pipeline (encode + decode) on a ONE-block workload, plus the launch list of a bench
# step.  usage: scripts/gpu_ncu_all.sh <tag>
TAG=${1:-r2}
mkdir -p gpurun_out
sha256sum libzling_b200/libzling.so | cut -c1-16 > gpurun_out/${TAG}_lib_sha16.txt
# launch list (device time per launch) of the default bench workload, 1 warm-up + 2 steps
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-decode > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "launch list rc=$?"
# full capture: the first launch of every kernel whose name starts with zl_ (the bench's parity encode + decode leg launch them all)
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:^zl_ -c 14 -f -o gpurun_out/${TAG}_all_kernels \
    python bench.py --steps 1 --warmup 0 --size-mb 16.7 > gpurun_out/${TAG}_ncu_all.log 2>&1; echo "ncu all rc=$?"
tail -3 gpurun_out/${TAG}_ncu_all.log; ls -la gpurun_out/${TAG}_all_kernels.ncu-rep
