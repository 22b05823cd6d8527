"""N > 1 host logic on CPU: world_size-2 and -3 gloo runs of libzling_b200.sharded with a CPU stand-in encoder built on
the oracle (range encode with carried state, oracle/zling_oracle.c zo_encode_range).  Checks that the carry hand-off +
single gather reproduce the single-process stream byte for byte — the same orchestration code drives the GPU encoder."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from libzling_b200 import sharded  # noqa: E402


class OracleRangeEncoder:
    """submit / set_state / complete / get_state on the CPU oracle (test stand-in for libzling_b200.Encoder)"""

    def __init__(self, level):
        from _libs import Oracle
        self.lib = Oracle().lib
        self.lib.zo_encode_range.restype = C.c_longlong
        self.lib.zo_encode_range.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
        self.level = level
        init = np.array(self.lib.zo_table_mtfinit()[:256], dtype=np.uint8)
        self.state = np.concatenate([np.tile(init, 256), np.frombuffer(np.int32(level).tobytes(), dtype=np.uint8)]).copy()
        self.pending = None

    def submit(self, a):
        self.pending = np.ascontiguousarray(a, dtype=np.uint8)

    def set_state(self, s):
        self.state = np.ascontiguousarray(s, dtype=np.uint8).copy()

    def get_state(self):
        return self.state.copy()

    def complete(self):
        a = self.pending
        out = np.zeros(a.size + a.size // 8 + 4096, dtype=np.uint8)
        n = self.lib.zo_encode_range(a.ctypes.data, a.size, self.level, out.ctypes.data, out.size, self.state.ctypes.data)
        assert 0 <= n <= out.size
        return out[:n].tobytes()


def _stream(nbytes, seed):
    """text | random | text, so that the level feedback flips across a range boundary as well"""
    from libzling_b200 import corpus
    rng = np.random.default_rng(seed)
    parts, left = [], nbytes
    while left > 0:
        k = min(left, int(rng.integers(1 << 20, 6 << 20)))
        parts.append(corpus.enwik8_shaped(k, seed=int(rng.integers(1, 1000))) if len(parts) % 2 == 0 else rng.integers(0, 256, k, dtype=np.uint8))
        left -= k
    return np.concatenate(parts)[:nbytes]


def _worker(rank, world, port, nbytes, level, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        data = _stream(nbytes, seed=5)
        enc = OracleRangeEncoder(level)
        out = sharded.encode_stream(enc, data, rank, world, dist)
        if rank == 0:
            from _libs import Oracle
            want = Oracle().encode(data, level)
            q.put((out == want, len(out), len(want)))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,nbytes,level", [(2, 3 * 16777216 + 12345, 2), (3, 2 * 16777216 + 1, 0)])
def test_block_ranges_over_gloo_reproduce_the_stream(world, nbytes, level):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nbytes, level, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    same, got, want = q.get(timeout=10)
    assert same, (got, want)


def test_block_ranges_cover_the_stream():
    for nbytes in (0, 1, 16777216, 16777217, 5 * 16777216 - 3, 1000000000):
        for world in (1, 2, 3, 8):
            r = sharded.block_ranges(nbytes, world)
            assert r[0][0] == 0 and r[-1][1] == nbytes
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert all(lo % 16777216 == 0 for lo, hi in r if hi > lo)
