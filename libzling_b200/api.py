"""ctypes binding of the C ABI (include/zlb.h) — what a Python user of the reference would call.

The reference has no Python layer; this mirrors its C++ surface one to one so that tests read like the
reference's usage (README.md:40-60: Encode(inputter, outputter, level) / Decode(inputter, outputter)):

    ctx = Context(device=0, max_blocks=8)
    z   = ctx.encode(data, level=2)        # baidu::zling::Encode  (src/libzling.cpp:174)
    raw = ctx.decode(z)                    # baidu::zling::Decode  (src/libzling.cpp:293)

There is no CPU fallback: a missing library or a missing CUDA device raises RuntimeError.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

BLOCK = 16777216
_u8p = C.POINTER(C.c_uint8)


class Stats(C.Structure):
    _fields_ = [("ms_total", C.c_double), ("ms_parse", C.c_double), ("ms_mtf", C.c_double), ("ms_huff_build", C.c_double),
                ("ms_pack", C.c_double), ("ms_h2d", C.c_double), ("ms_d2h", C.c_double), ("launches", C.c_uint32),
                ("parse_launches", C.c_uint32), ("reparsed_blocks", C.c_uint32), ("tokens", C.c_uint64), ("subblocks", C.c_uint64),
                ("slow_main", C.c_uint64), ("slow_lazy", C.c_uint64), ("window_hits", C.c_uint64), ("windows", C.c_uint64),
                ("cyc_spec", C.c_uint64), ("cyc_resolve", C.c_uint64), ("general_path", C.c_uint64), ("cyc_total", C.c_uint64), ("flagged", C.c_uint64),
                ("rounds", C.c_uint64), ("cyc_final", C.c_uint64), ("cyc_orbit", C.c_uint64), ("cyc_rank", C.c_uint64), ("cyc_decide", C.c_uint64)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class ShardStats(C.Structure):
    _fields_ = [("ms_wait_carry", C.c_double), ("ms_gather", C.c_double), ("local_bytes", C.c_uint64), ("total_bytes", C.c_uint64), ("spec_reparsed_blocks", C.c_uint64)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class StreamIO(C.Structure):
    _fields_ = [("inp", C.c_void_p), ("n", C.c_size_t), ("out", C.c_void_p), ("out_cap", C.c_size_t), ("out_len", C.c_size_t)]


class SubBlock(C.Structure):
    _fields_ = [(k, C.c_uint32) for k in ("tok_begin", "tok_end", "enc_begin", "enc_end", "rlen", "level", "olen", "bits_lo")]


EXPORTS = ["zlb_device_count", "zlb_create", "zlb_destroy", "zlb_max_blocks", "zlb_last_error", "zlb_version",
           "zlb_host_alloc", "zlb_host_free", "zlb_encode_bound", "zlb_encoder_begin", "zlb_encoder_end",
           "zlb_encode_blocks", "zlb_encode_blocks_device", "zlb_encode_submit", "zlb_encode_complete", "zlb_encoder_get_state", "zlb_encoder_set_state",
           "zlb_decoder_begin", "zlb_decoder_end", "zlb_decode_blocks", "zlb_get_stats", "zlb_debug_tokens",
           "zlb_debug_subblocks", "zlb_debug_huff_tables",
           "zlb_comm_get_unique_id", "zlb_comm_create", "zlb_comm_destroy", "zlb_comm_get_stats", "zlb_encode_stream_sharded", "zlb_encode_blocks_gathered", "zlb_gather_packed",
           "zlb_encode_batch", "zlb_decode_batch"]

_lib = None


def lib_path():
    return _build.LIB


def load():
    """load libzling.so (built in-tree by libzling_b200.build); raises if it is missing — never falls back"""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError("libzling_b200: %s not built (run `python -m libzling_b200.build`); there is no CPU fallback" % path)
    L = C.CDLL(path)
    L.zlb_create.restype = C.c_void_p
    L.zlb_create.argtypes = [C.c_int, C.c_int]
    L.zlb_destroy.argtypes = [C.c_void_p]
    L.zlb_max_blocks.argtypes = [C.c_void_p]
    L.zlb_last_error.restype = C.c_char_p
    L.zlb_version.restype = C.c_char_p
    L.zlb_host_alloc.restype = C.c_void_p
    L.zlb_host_alloc.argtypes = [C.c_size_t]
    L.zlb_host_free.argtypes = [C.c_void_p]
    L.zlb_encode_bound.restype = C.c_size_t
    L.zlb_encode_bound.argtypes = [C.c_size_t]
    L.zlb_encoder_begin.restype = C.c_void_p
    L.zlb_encoder_begin.argtypes = [C.c_void_p, C.c_int]
    L.zlb_encoder_end.argtypes = [C.c_void_p]
    for f in (L.zlb_encode_blocks, L.zlb_encode_blocks_device):
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.zlb_encode_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.zlb_encode_complete.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.zlb_encoder_get_state.argtypes = [C.c_void_p, _u8p]
    L.zlb_encoder_set_state.argtypes = [C.c_void_p, _u8p]
    L.zlb_decoder_begin.restype = C.c_void_p
    L.zlb_decoder_begin.argtypes = [C.c_void_p]
    L.zlb_decoder_end.argtypes = [C.c_void_p]
    L.zlb_decode_blocks.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.zlb_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    L.zlb_debug_tokens.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint32), C.c_size_t, C.POINTER(C.c_size_t)]
    L.zlb_debug_subblocks.argtypes = [C.c_void_p, C.c_int, C.POINTER(SubBlock), C.c_size_t, C.POINTER(C.c_size_t)]
    L.zlb_debug_huff_tables.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.c_int, C.c_int, C.c_int, _u8p, C.POINTER(C.c_uint16)]
    L.zlb_comm_get_unique_id.argtypes = [_u8p]
    L.zlb_comm_create.restype = C.c_void_p
    L.zlb_comm_create.argtypes = [C.c_void_p, C.c_int, C.c_int, _u8p]
    L.zlb_comm_destroy.argtypes = [C.c_void_p]
    L.zlb_comm_get_stats.argtypes = [C.c_void_p, C.POINTER(ShardStats)]
    L.zlb_encode_stream_sharded.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.zlb_encode_blocks_gathered.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_uint64)]
    L.zlb_encode_batch.argtypes = [C.c_void_p, C.c_int, C.POINTER(StreamIO), C.c_int]
    L.zlb_decode_batch.argtypes = [C.c_void_p, C.POINTER(StreamIO), C.c_int]
    L.zlb_gather_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_uint64)]
    _lib = L
    return L


class ZlingError(RuntimeError):
    pass


class FormatError(ZlingError, ValueError):
    """malformed compressed stream — where the reference throws std::runtime_error (src/libzling.cpp:316-407)"""


def _check(rc):
    if rc < 0:
        msg = load().zlb_last_error().decode()
        raise (FormatError if rc == -6 else ZlingError)("zlb error %d: %s" % (rc, msg))


def _as_u8(data):
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
    return np.frombuffer(bytes(data), dtype=np.uint8)


class PinnedBuffer:
    """page-locked host memory exposed as a numpy uint8 array"""

    def __init__(self, nbytes):
        L = load()
        self._p = L.zlb_host_alloc(nbytes)
        if not self._p:
            raise ZlingError(L.zlb_last_error().decode())
        self.array = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(self._p))

    def close(self):
        if self._p:
            self.array = None
            load().zlb_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """one GPU's block pipeline (zlb_ctx).  max_blocks 16 MiB blocks are processed per device call."""

    def __init__(self, device=0, max_blocks=8):
        L = load()
        if L.zlb_device_count() <= 0:
            raise ZlingError("libzling_b200: no CUDA device visible; this library has no CPU path")
        self._h = L.zlb_create(device, max_blocks)
        if not self._h:
            raise ZlingError(L.zlb_last_error().decode())
        self.device, self.max_blocks = device, max_blocks

    def close(self):
        if self._h:
            load().zlb_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- whole-stream helpers (the shape of baidu::zling::Encode / Decode over in-memory buffers) -------
    def encode(self, data, level=0):
        enc = Encoder(self, level)
        try:
            a = _as_u8(data)
            parts = []
            step = self.max_blocks * BLOCK
            for off in range(0, a.size, step):
                parts.append(enc.encode_blocks(a[off:off + step]))
            return b"".join(parts)
        finally:
            enc.close()

    def decode(self, data, size_hint=None):
        dec = Decoder(self)
        try:
            a = _as_u8(data)
            parts, at = [], 0
            while at < a.size:
                used, raw = dec.decode_blocks(a[at:])
                parts.append(raw)
                at += used
            return b"".join(parts)
        finally:
            dec.close()

    # -- batch: many independent streams in one call (zlb_encode_batch / zlb_decode_batch) -------------------
    def _batch(self, streams, caps, call):
        ins = [_as_u8(x) for x in streams]
        outs = [np.empty(max(int(cap), 1), dtype=np.uint8) for cap in caps]
        arr = (StreamIO * len(ins))()
        for i, (a, o) in enumerate(zip(ins, outs)):
            arr[i].inp, arr[i].n, arr[i].out, arr[i].out_cap, arr[i].out_len = a.ctypes.data, a.size, o.ctypes.data, o.size, 0
        _check(call(arr, len(ins)))
        return [bytes(o[:arr[i].out_len]) for i, o in enumerate(outs)]

    def encode_batch(self, streams, level=0):
        """whole streams in, framed streams out; all blocks of all streams share one pass of the pipeline"""
        L = load()
        return self._batch(streams, [L.zlb_encode_bound(len(_as_u8(x))) for x in streams], lambda arr, n: L.zlb_encode_batch(self._h, level, arr, n))

    def decode_batch(self, streams, raw_sizes):
        """framed streams in, raw streams out (raw_sizes: upper bounds of the decoded sizes); one decode chain per stream"""
        L = load()
        return self._batch(streams, raw_sizes, lambda arr, n: L.zlb_decode_batch(self._h, arr, n))

    def stats(self):
        s = Stats()
        _check(load().zlb_get_stats(self._h, C.byref(s)))
        return s.asdict()

    # -- intermediates of the last encode call (parity tests) -------------------------------------------
    def debug_tokens(self, blk):
        L = load()
        n = C.c_size_t(0)
        _check(L.zlb_debug_tokens(self._h, blk, None, 0, C.byref(n)))
        out = np.zeros(max(n.value, 1), dtype=np.uint32)
        _check(L.zlb_debug_tokens(self._h, blk, out.ctypes.data_as(C.POINTER(C.c_uint32)), out.size, C.byref(n)))
        return out[:n.value]

    def debug_subblocks(self, blk):
        L = load()
        arr = (SubBlock * 80)()
        n = C.c_size_t(0)
        _check(L.zlb_debug_subblocks(self._h, blk, arr, 80, C.byref(n)))
        return [{k: getattr(arr[i], k) for k, _ in SubBlock._fields_} for i in range(n.value)]

    def debug_huff_tables(self, freq, cap):
        f = np.ascontiguousarray(freq, dtype=np.uint32)
        nt, ns = f.shape
        lens = np.zeros((nt, ns), dtype=np.uint8)
        codes = np.zeros((nt, ns), dtype=np.uint16)
        _check(load().zlb_debug_huff_tables(self._h, f.ctypes.data_as(C.POINTER(C.c_uint32)), nt, ns, cap,
                                            lens.ctypes.data_as(_u8p), codes.ctypes.data_as(C.POINTER(C.c_uint16))))
        return lens, codes


class Encoder:
    """one stream being encoded (zlb_encoder): carries the MTF tables and the level-feedback flag across calls"""

    def __init__(self, ctx, level=0):
        L = load()
        self.ctx = ctx
        self._h = L.zlb_encoder_begin(ctx._h, level)
        if not self._h:
            raise ZlingError(L.zlb_last_error().decode())
        self._out = None

    def close(self):
        if self._h:
            load().zlb_encoder_end(self._h)
            self._h = None

    def encode_blocks(self, data, out=None):
        """host buffers in, framed bytes out (copies included).  `data`: whole 16 MiB blocks except at stream end"""
        a = _as_u8(data)
        L = load()
        cap = L.zlb_encode_bound(a.size)
        if out is None:
            out = np.empty(cap, dtype=np.uint8)
        n = C.c_size_t(0)
        _check(L.zlb_encode_blocks(self._h, a.ctypes.data, a.size, out.ctypes.data, out.size, C.byref(n)))
        return out[:n.value].tobytes()

    def encode_blocks_into(self, a, out):
        """like encode_blocks but writes into a caller buffer (e.g. pinned) and returns the byte count"""
        n = C.c_size_t(0)
        _check(load().zlb_encode_blocks(self._h, a.ctypes.data, a.size, out.ctypes.data, out.size, C.byref(n)))
        return n.value

    def encode_blocks_device(self, d_in_ptr, nbytes, d_out_ptr, out_cap):
        """device pointers (e.g. torch tensors' data_ptr()); returns the byte count left in d_out"""
        n = C.c_size_t(0)
        _check(load().zlb_encode_blocks_device(self._h, d_in_ptr, nbytes, d_out_ptr, out_cap, C.byref(n)))
        return n.value

    def submit(self, data):
        """split form, step 1: H2D + parse launch of a block range; returns immediately (see include/zlb.h)"""
        a = _as_u8(data)
        self._pending = a                     # keep the host buffer alive until complete()
        _check(load().zlb_encode_submit(self._h, a.ctypes.data, a.size))

    def complete(self):
        """split form, step 2: MTF + Huffman + framing of the submitted range with the state installed by now"""
        L = load()
        out = np.empty(L.zlb_encode_bound(self._pending.size), dtype=np.uint8)
        n = C.c_size_t(0)
        _check(L.zlb_encode_complete(self._h, out.ctypes.data, out.size, C.byref(n)))
        self._pending = None
        return out[:n.value].tobytes()

    def get_state(self):
        s = np.zeros(65540, dtype=np.uint8)
        _check(load().zlb_encoder_get_state(self._h, s.ctypes.data_as(_u8p)))
        return s

    def set_state(self, s):
        s = np.ascontiguousarray(s, dtype=np.uint8)
        assert s.size == 65540
        _check(load().zlb_encoder_set_state(self._h, s.ctypes.data_as(_u8p)))


class Comm:
    """NCCL communicator of the ranks that share ONE stream (zlb_comm): one process per GPU.  `bcast(id_bytes)` must
    return rank 0's 128-byte id on every rank (e.g. a torch.distributed broadcast, an MPI bcast, a file)."""

    def __init__(self, ctx, rank, world, bcast):
        L = load()
        idb = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            _check(L.zlb_comm_get_unique_id(idb.ctypes.data_as(_u8p)))
        idb = np.ascontiguousarray(bcast(idb), dtype=np.uint8)
        self.ctx, self.rank, self.world = ctx, rank, world
        self._h = L.zlb_comm_create(ctx._h, rank, world, idb.ctypes.data_as(_u8p))
        if not self._h:
            raise ZlingError(L.zlb_last_error().decode())

    def close(self):
        if self._h:
            load().zlb_comm_destroy(self._h)
            self._h = None

    def stats(self):
        s = ShardStats()
        _check(load().zlb_comm_get_stats(self._h, C.byref(s)))
        return s.asdict()

    def encode_stream(self, enc, shard, out=None, device_ptr=None):
        """this rank's block range of ONE stream (host array, or a device pointer + shard = byte count); on rank 0 `out`
        (a host uint8 array) receives the whole framed stream.  Returns the byte count (total on rank 0, local elsewhere)."""
        n = C.c_size_t(0)
        if device_ptr is not None:
            ptr, size, on_dev = device_ptr, int(shard), 1
        else:
            a = _as_u8(shard)
            self._keep = a
            ptr, size, on_dev = a.ctypes.data, a.size, 0
        _check(load().zlb_encode_stream_sharded(enc._h, self._h, ptr, size, on_dev, out.ctypes.data if out is not None else None,
                                                out.size if out is not None else 0, C.byref(n)))
        return n.value

    def encode_gathered(self, enc, data, out=None, device_ptr=None):
        """independent streams, one per rank: encode this rank's stream and gather all framed streams on rank 0 (one NCCL
        gather, device to device); returns (bytes, sizes) — bytes = total on rank 0, local elsewhere"""
        n = C.c_size_t(0)
        sizes = (C.c_uint64 * self.world)()
        if device_ptr is not None:
            ptr, size, on_dev = device_ptr, int(data), 1
        else:
            a = _as_u8(data)
            self._keep = a
            ptr, size, on_dev = a.ctypes.data, a.size, 0
        _check(load().zlb_encode_blocks_gathered(enc._h, self._h, ptr, size, on_dev, out.ctypes.data if out is not None else None,
                                                 out.size if out is not None else 0, C.byref(n), sizes))
        return n.value, list(sizes)

    def gather_packed(self, d_ptr, nbytes, out=None):
        """independent streams per rank: one gather of the packed outputs (device memory) to rank 0; returns (total, sizes)"""
        n = C.c_size_t(0)
        sizes = (C.c_uint64 * self.world)()
        _check(load().zlb_gather_packed(self._h, d_ptr, nbytes, out.ctypes.data if out is not None else None,
                                        out.size if out is not None else 0, C.byref(n), sizes))
        return n.value, list(sizes)


class Decoder:
    def __init__(self, ctx):
        L = load()
        self.ctx = ctx
        self._h = L.zlb_decoder_begin(ctx._h)
        if not self._h:
            raise ZlingError(L.zlb_last_error().decode())

    def close(self):
        if self._h:
            load().zlb_decoder_end(self._h)
            self._h = None

    def decode_blocks(self, data):
        """decodes as many complete blocks as fit (<= max_blocks); returns (compressed bytes used, raw bytes)"""
        a = _as_u8(data)
        out = np.empty(self.ctx.max_blocks * BLOCK, dtype=np.uint8)
        used, n = C.c_size_t(0), C.c_size_t(0)
        _check(load().zlb_decode_blocks(self._h, a.ctypes.data, a.size, C.byref(used), out.ctypes.data, out.size, C.byref(n)))
        return used.value, out[:n.value].tobytes()
