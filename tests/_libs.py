"""ctypes loaders for the TEST-ONLY checkers: oracle/libzling_oracle.so (C restatement) and, when it was built
in the container that has /root/reference, oracle/_ref/libzling_ref.so (the unmodified reference)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
u8p = C.POINTER(C.c_uint8)


def _ptr(a, t=C.c_uint8):
    return a.ctypes.data_as(C.POINTER(t))


def as_u8(data):
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data, dtype=np.uint8)
    return np.frombuffer(bytes(data), dtype=np.uint8)


def bound(n):
    """generous upper bound of the compressed size of n bytes (273+12+1 per <=~131k-byte sub-block + 1/blk)"""
    return int(n * 1.01) + (n // 100000 + 2) * 300 + 64


class _Codec:
    """whole-stream encode/decode through one of the checker libraries (same call shape for both)"""

    def __init__(self, lib, enc, dec):
        self.lib = lib
        self._enc = getattr(lib, enc)
        self._dec = getattr(lib, dec)
        self._enc.restype = C.c_longlong
        self._enc.argtypes = [u8p, C.c_size_t, C.c_int, u8p, C.c_size_t]
        self._dec.restype = C.c_longlong
        self._dec.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t]

    def encode(self, data, level=0):
        a = as_u8(data)
        out = np.empty(bound(a.size), dtype=np.uint8)
        n = self._enc(_ptr(a), a.size, level, _ptr(out), out.size)
        if n < 0:
            raise RuntimeError("encode failed rc=%d" % n)
        return out[:n].tobytes()

    def decode(self, data, cap):
        a = as_u8(data)
        out = np.empty(max(cap, 1), dtype=np.uint8)
        n = self._dec(_ptr(a), a.size, _ptr(out), cap)
        if n < 0:
            raise ValueError("decode failed rc=%d" % n)
        return out[:n].tobytes()


class Oracle(_Codec):
    def __init__(self):
        path = os.path.join(ORACLE_DIR, "libzling_oracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)
        lib = C.CDLL(path)
        super().__init__(lib, "zo_encode", "zo_decode")
        lib.zo_rolz_new.restype = C.c_void_p
        lib.zo_rolz_free.argtypes = [C.c_void_p]
        lib.zo_rolz_reset.argtypes = [C.c_void_p]
        lib.zo_rolz_encode.argtypes = [C.c_void_p, C.c_int, u8p, C.POINTER(C.c_uint16), C.c_int, C.c_int, C.POINTER(C.c_int)]
        lib.zo_rolz_trace.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), u8p, C.c_int]
        lib.zo_rolz_trace_count.argtypes = [C.c_void_p]
        lib.zo_rolz_get_mtf.argtypes = [C.c_void_p, u8p]
        lib.zo_make_length_table.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_int, C.c_int]
        lib.zo_make_encode_table.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint16), C.c_int, C.c_int]
        lib.zo_huff_encode_subblock.argtypes = [C.POINTER(C.c_uint16), C.c_int, u8p]
        for f, t in (("zo_table_mtfinit", C.c_uint8), ("zo_table_mtfnext", C.c_uint8), ("zo_table_idx_code", C.c_uint8),
                     ("zo_table_idx_base", C.c_uint16), ("zo_table_idx_bits", C.c_uint8)):
            getattr(lib, f).restype = C.POINTER(t)

    def table(self, name, n):
        return np.array(getattr(self.lib, "zo_table_" + name)()[:n])

    def length_table(self, freq, cap):
        f = np.ascontiguousarray(freq, dtype=np.uint32)
        out = np.zeros(f.size, dtype=np.uint32)
        self.lib.zo_make_length_table(_ptr(f, C.c_uint32), _ptr(out, C.c_uint32), f.size, cap)
        return out

    def encode_table(self, lens, cap):
        l = np.ascontiguousarray(lens, dtype=np.uint32)
        out = np.zeros(l.size, dtype=np.uint16)
        self.lib.zo_make_encode_table(_ptr(l, C.c_uint32), _ptr(out, C.c_uint16), l.size, cap)
        return out

    def huff_payload(self, syms):
        s = np.ascontiguousarray(syms, dtype=np.uint16)
        out = np.zeros(393216 + 300, dtype=np.uint8)
        n = self.lib.zo_huff_encode_subblock(_ptr(s, C.c_uint16), s.size, _ptr(out))
        return out[:n].tobytes()

    def parse_block(self, block, level, trace=False, levels=None):
        """tokenise ONE block (<=16 MiB) at a fixed level (or per-sub-block `levels`): list of dicts with
        encpos, syms (u16 array, literals MTF-ranked) and — with trace — token positions / raw literal bytes"""
        return _parse_block(self.lib, "zo", block, level, trace, levels)


class Ref(_Codec):
    """the unmodified reference, only where oracle/_ref was built (container with /root/reference)"""

    def __init__(self):
        path = os.path.join(ORACLE_DIR, "_ref", "libzling_ref.so")
        if not os.path.exists(path):
            if not os.path.isdir("/root/reference/src"):
                raise FileNotFoundError("oracle/_ref not built and /root/reference absent")
            subprocess.check_call(["make", "-C", ORACLE_DIR, "ref"], stdout=subprocess.DEVNULL)
        lib = C.CDLL(path)
        super().__init__(lib, "zref_encode", "zref_decode")
        lib.zref_rolz_new.restype = C.c_void_p
        lib.zref_rolz_free.argtypes = [C.c_void_p]
        lib.zref_rolz_reset.argtypes = [C.c_void_p]
        lib.zref_rolz_encode.argtypes = [C.c_void_p, C.c_int, u8p, C.POINTER(C.c_uint16), C.c_int, C.c_int, C.POINTER(C.c_int)]
        lib.zref_make_length_table.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_int, C.c_int]
        lib.zref_make_encode_table.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint16), C.c_int, C.c_int]

    def length_table(self, freq, cap):
        f = np.ascontiguousarray(freq, dtype=np.uint32)
        out = np.zeros(f.size + 1, dtype=np.uint32)
        self.lib.zref_make_length_table(_ptr(f, C.c_uint32), _ptr(out, C.c_uint32), f.size, cap)
        return out[:f.size]

    def encode_table(self, lens, cap):
        l = np.ascontiguousarray(lens, dtype=np.uint32)
        out = np.zeros(l.size, dtype=np.uint16)
        self.lib.zref_make_encode_table(_ptr(l, C.c_uint32), _ptr(out, C.c_uint16), l.size, cap)
        return out

    def parse_block(self, block, level, trace=False, levels=None):
        assert not trace
        return _parse_block(self.lib, "zref", block, level, False, levels)


def _parse_block(lib, prefix, block, level, trace, levels):
    a = as_u8(block)
    assert a.size <= 16777216
    pad = np.concatenate([a, np.zeros(300, dtype=np.uint8)])
    new, free, enc = (getattr(lib, prefix + "_rolz_" + s) for s in ("new", "free", "encode"))
    h = C.c_void_p(new())
    encpos = C.c_int(0)
    syms = np.zeros(262144 + 300, dtype=np.uint16)
    tp = np.zeros(262144, dtype=np.uint32)
    tr = np.zeros(262144, dtype=np.uint8)
    out = []
    try:
        while encpos.value < a.size:
            lv = level if levels is None else levels[min(len(out), len(levels) - 1)]
            if trace:
                lib.zo_rolz_trace(h, _ptr(tp, C.c_uint32), _ptr(tr), tp.size)
            rlen = enc(h, lv, _ptr(pad), _ptr(syms, C.c_uint16), a.size, 262144, C.byref(encpos))
            rec = {"encpos": encpos.value, "syms": syms[:rlen].copy(), "level": lv}
            if trace:
                nt = lib.zo_rolz_trace_count(h)
                rec["tok_pos"] = tp[:nt].copy()
                rec["tok_raw"] = tr[:nt].copy()
            out.append(rec)
    finally:
        free(h)
    return out


def have_ref():
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libzling_ref.so")) or os.path.isdir("/root/reference/src")


def walk_container(stream):
    """split a zling stream into [(block_index, encpos, rlen, olen, payload_bytes)] (src/libzling.cpp:200,269-278)"""
    s = bytes(stream)
    at, blk, out = 0, 0, []
    while at < len(s):
        flag = s[at]; at += 1
        if flag == 0:
            blk += 1
            continue
        assert flag == 1
        encpos, rlen, olen = (int.from_bytes(s[at + 4 * k: at + 4 * k + 4], "big") for k in range(3))
        at += 12
        out.append((blk, encpos, rlen, olen, s[at:at + olen]))
        at += olen
    return out
