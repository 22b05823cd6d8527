#!/bin/bash
# ncu captures of the final library: (1) launch list of a bench step, (2) --set full of the parse kernel on the DEFAULT 100 MB workload
# (the figure bench.py reports as roofline.traffic), (3) --set full of every kernel of the pipeline (encode + decode) on a ONE-block
# workload.  usage: scripts/gpu_ncu_all.sh <tag>
TAG=${1:-r2}
mkdir -p gpurun_out
echo "$(sha256sum libzling_b200/libzling.so | cut -c1-16):$(python -c 'from libzling_b200 import build; print(build.source_sha16())')" > gpurun_out/${TAG}_lib_sha16.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-decode > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^zl_rolz_parse -c 1 -f -o gpurun_out/${TAG}_parse_100mb \
    python bench.py --steps 1 --warmup 0 --no-decode --skip-parity > gpurun_out/${TAG}_ncu_parse.log 2>&1; echo "ncu parse rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:^zl_ -c 14 -f -o gpurun_out/${TAG}_all_kernels \
    python bench.py --steps 1 --warmup 0 --size-mb 16.7 > gpurun_out/${TAG}_ncu_all.log 2>&1; echo "ncu all rc=$?"
tail -2 gpurun_out/${TAG}_ncu_all.log | cut -c1-200; ls -la gpurun_out/${TAG}_*.ncu-rep
