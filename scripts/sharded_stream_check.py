#!/usr/bin/env python
"""One stream over N GPUs (torchrun, NCCL): contiguous block ranges per rank, carried state handed rank -> rank with
send/recv, one gather of the framed outputs (libzling_b200/sharded.py).  Rank 0 compares the gathered stream with the
CPU checker byte for byte and prints one JSON line.   torchrun --nproc-per-node N scripts/sharded_stream_check.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist

import libzling_b200
from libzling_b200 import corpus, sharded


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nbytes = int(float(os.environ.get("ZL_STREAM_MB", "150")) * 1e6)
    level = int(os.environ.get("ZL_LEVEL", "2"))
    data = corpus.mixed(nbytes, seed=21)                       # identical on every rank
    lo, hi = sharded.block_ranges(nbytes, world)[rank]
    ctx = libzling_b200.Context(device=local, max_blocks=max(1, (hi - lo + libzling_b200.BLOCK - 1) // libzling_b200.BLOCK))
    times = []
    out = None
    for _ in range(2):
        enc = libzling_b200.Encoder(ctx, level)
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        out = sharded.encode_stream(enc, data, rank, world, dist, device="cuda")
        torch.cuda.synchronize(); dist.barrier()
        times.append(time.perf_counter() - t0)
        enc.close()
    if rank == 0:
        from _libs import Oracle
        want = Oracle().encode(data, level)
        print(json.dumps({"check": "one stream over %d GPUs" % world, "bytes": nbytes, "level": level, "bit_exact": out == want,
                          "compressed": len(out), "seconds": round(min(times), 3), "MBps": round(nbytes / 1e6 / min(times), 2)}), flush=True)
        if out != want:
            raise SystemExit(1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
