"""ONE stream over several GPUs through the C ABI (zlb_comm_*, zlb_encode_stream_sharded, zlb_encode_blocks_gathered): the
carried state (MTF tables + level, src/libzling_lz.h:105, src/libzling.cpp:185,261-266) moves GPU -> GPU by ncclSend/ncclRecv,
one NCCL gather brings the framed ranges to rank 0, and the bytes must equal the single-process reference stream.
Needs >= 2 GPUs (skipped otherwise): one process per GPU, NCCL over 127.0.0.1."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

pytestmark = pytest.mark.gpu


def _stream(nbytes, seed):
    """text | random | text ..., so that the level feedback flips across a rank boundary as well"""
    from libzling_b200 import corpus
    rng = np.random.default_rng(seed)
    parts, left = [], nbytes
    while left > 0:
        k = min(left, int(rng.integers(3 << 20, 9 << 20)))
        parts.append(corpus.enwik8_shaped(k, seed=int(rng.integers(1, 1000))) if len(parts) % 2 == 0 else rng.integers(0, 256, k, dtype=np.uint8))
        left -= k
    return np.concatenate(parts)


def _worker(rank, world, port, nbytes, level, mode, q):
    import torch
    import torch.distributed as dist
    import libzling_b200
    from libzling_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        def bcast(idb):
            t = torch.from_numpy(idb.copy()).cuda()
            dist.broadcast(t, src=0)
            return t.cpu().numpy()
        L = libzling_b200.load()
        if mode == "stream":
            data = _stream(nbytes, 11)
            lo, hi = sharded.block_ranges(data.size, world)[rank]
            ctx = libzling_b200.Context(device=rank, max_blocks=max(1, (hi - lo + libzling_b200.BLOCK - 1) // libzling_b200.BLOCK))
            comm = libzling_b200.Comm(ctx, rank, world, bcast)
            out = np.empty(L.zlb_encode_bound(data.size), dtype=np.uint8) if rank == 0 else None
            enc = libzling_b200.Encoder(ctx, level)
            n = comm.encode_stream(enc, data[lo:hi], out=out)
            enc.close()
            if rank == 0:
                q.put(("stream", bytes(out[:n])))
        else:
            data = _stream(nbytes, 20 + rank)
            ctx = libzling_b200.Context(device=rank, max_blocks=(data.size + libzling_b200.BLOCK - 1) // libzling_b200.BLOCK)
            comm = libzling_b200.Comm(ctx, rank, world, bcast)
            out = np.empty(world * L.zlb_encode_bound(data.size), dtype=np.uint8) if rank == 0 else None
            enc = libzling_b200.Encoder(ctx, level)
            n, sizes = comm.encode_gathered(enc, data, out=out)
            enc.close()
            if rank == 0:
                q.put(("streams", bytes(out[:n]), sizes))
        comm.close()
        ctx.close()
    finally:
        dist.destroy_process_group()


def _run(world, nbytes, level, mode):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_worker, args=(r, world, port, nbytes, level, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return res


def _world():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    return min(n, 4)


@pytest.mark.parametrize("level", [2, 0])
def test_one_stream_over_gpus_is_bit_exact(oracle, level):
    world = _world()
    nbytes = 5 * 16777216 + 12345                      # 6 blocks over 2..4 ranks: uneven ranges, a short last block
    res = _run(world, nbytes, level, "stream")
    assert res[1] == oracle.encode(_stream(nbytes, 11), level)


def test_independent_streams_gathered(oracle):
    world = _world()
    nbytes = 16777216 + 54321
    _, blob, sizes = _run(world, nbytes, 1, "streams")
    at = 0
    for r in range(world):
        want = oracle.encode(_stream(nbytes, 20 + r), 1)
        assert sizes[r] == len(want) and blob[at:at + sizes[r]] == want, r
        at += sizes[r]
    assert at == len(blob)
