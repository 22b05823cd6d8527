#!/bin/bash
# N-GPU checks (run under `gpurun --gpus N`): the bench contract at N ranks + the in-stream sharded path
N=${1:-2}; TAG=${2:-multi}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench rc=$?"
cut -c1-900 gpurun_out/${TAG}_bench_n$N.json; tail -3 gpurun_out/${TAG}_bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/${TAG}_ref_n$N.json 2> gpurun_out/${TAG}_ref_n$N.err; echo "ref rc=$?"
cut -c1-600 gpurun_out/${TAG}_ref_n$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/sharded_stream_check.py > gpurun_out/${TAG}_sharded_n$N.json 2> gpurun_out/${TAG}_sharded_n$N.err; echo "sharded rc=$?"
cat gpurun_out/${TAG}_sharded_n$N.json; tail -3 gpurun_out/${TAG}_sharded_n$N.err
