#!/bin/bash
# Multi-GPU call of round 2 (gpurun --gpus N).  usage: scripts/gpu_r2_multi.sh <tag> <N> [tests] [bench] [ref] [s200] [big]
#   tests  sharded-stream GPU tests        bench  default bench (one stream per GPU) at N      ref   reference arm at N
#   s200   ONE 200 MB mixed e2 stream over N GPUs                 big    ONE 1 GB mixed e2 stream over N GPUs
TAG=${1:-r2m}; N=${2:-2}; shift 2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
env | grep -i nccl > gpurun_out/${TAG}_nccl_env.txt
run() { timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:3}" > gpurun_out/${TAG}_$2.json 2> gpurun_out/${TAG}_$2.err; echo "$2 rc=$?"; tail -1 gpurun_out/${TAG}_$2.json | cut -c1-1800; tail -2 gpurun_out/${TAG}_$2.err | cut -c1-300; }
for what in "$@"; do
  case $what in
    tests) timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/${TAG}_pytest_sharded.log 2>&1; echo "pytest sharded rc=$?"; tail -5 gpurun_out/${TAG}_pytest_sharded.log ;;
    bench) run 29511 bench_n${N} --steps 3 --warmup 3 --no-decode ;;
    ref)   timeout 600 python bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/${TAG}_ref_n${N}.json 2> gpurun_out/${TAG}_ref_n${N}.err; tail -1 gpurun_out/${TAG}_ref_n${N}.json | cut -c1-600 ;;
    s200)  run 29512 stream200_n${N} --steps 2 --warmup 3 --no-decode --shard stream --corpus mixed --size-mb 200 --level 2 ;;
    big)   run 29513 stream1g_n${N} --steps 2 --warmup 3 --no-decode --shard stream --corpus mixed --size-mb 1000 --level 2 ;;
  esac
done
ls gpurun_out | grep nccl | head -12
