#!/bin/bash
# One gpurun call: GPU parity tests, a bench line, the ncu launch list and one full ncu capture of the parse kernel.
# usage: scripts/gpu_round.sh <tag> [quick]
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; RC=$?; echo "smoke rc=$RC" >> gpurun_out/${TAG}_smoke.log
tail -4 gpurun_out/${TAG}_smoke.log
if [ $RC -ne 0 ]; then echo "smoke failed: trying the v2 parse / v1 mtf to localise"; ZLB_PARSE=2 timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2; ZLB_MTF=1 timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2; exit 1; fi
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
if [ "$2" == "big" ]; then
  timeout 900 python bench.py --steps 2 --warmup 3 --size-mb 1000 --level 2 --corpus mixed > gpurun_out/${TAG}_bench_1g_e2.json 2> gpurun_out/${TAG}_bench_1g_e2.err; echo "bench 1g rc=$?"
  cat gpurun_out/${TAG}_bench_1g_e2.json | cut -c1-1500; tail -3 gpurun_out/${TAG}_bench_1g_e2.err
  timeout 600 python bench.py --steps 2 --warmup 3 --level 4 > gpurun_out/${TAG}_bench_e4.json 2> gpurun_out/${TAG}_bench_e4.err; echo "bench e4 rc=$?"
  cat gpurun_out/${TAG}_bench_e4.json | cut -c1-1500; tail -3 gpurun_out/${TAG}_bench_e4.err
fi
if [ "$2" == "full" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${TAG}_ncu_bench.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:zl_rolz_parse_v3 -c 1 -o gpurun_out/${TAG}_parse_v3 -f python bench.py --steps 1 --warmup 0 --skip-parity > gpurun_out/${TAG}_ncu_full.log 2>&1
  ls -la gpurun_out | tail -8
fi
