"""Pins the C restatement (oracle/zling_oracle.c) to the UNMODIFIED reference compiled here (oracle/_ref):
whole streams at e0-e4, per-sub-block symbol buffers, Huffman tables.  Skipped where /root/reference and a
prebuilt oracle/_ref are both absent."""
import numpy as np
import pytest

from _inputs import small_cases, block_boundary_cases

SMALL = small_cases()


@pytest.mark.parametrize("level", [0, 1, 2, 3, 4])
def test_streams_small(oracle, ref, level):
    for name, data in SMALL:
        a, b = oracle.encode(data, level), ref.encode(data, level)
        assert a == b, (name, level, len(a), len(b))
        assert oracle.decode(a, len(data)) == data, name
        assert ref.decode(a, len(data)) == data, name


@pytest.mark.parametrize("level", [0, 2, 4])
def test_streams_block_boundaries(oracle, ref, level):
    for name, data in block_boundary_cases():
        a, b = oracle.encode(data, level), ref.encode(data, level)
        assert a == b, (name, level)
        assert oracle.decode(a, len(data)) == data, name


def test_symbol_buffers(oracle, ref):
    for name, data in SMALL:
        if len(data) < 4:
            continue
        for level in (0, 3, 4):
            pa, pb = oracle.parse_block(data, level), ref.parse_block(data, level)
            assert len(pa) == len(pb), name
            for x, y in zip(pa, pb):
                assert x["encpos"] == y["encpos"] and np.array_equal(x["syms"], y["syms"]), (name, level)


def test_huffman_tables_random(oracle, ref):
    rng = np.random.default_rng(123)
    for t in range(4000):
        n, cap = (514, 15) if t % 2 == 0 else (32, 8)
        kind = t % 5
        if kind == 0:
            f = rng.integers(0, 50, size=n)
        elif kind == 1:
            f = (rng.pareto(0.7, size=n) * 3).astype(np.int64)        # skewed: forces the rescale path
        elif kind == 2:
            f = np.zeros(n, dtype=np.int64); k = rng.integers(1, 6); f[rng.choice(n, size=k, replace=False)] = rng.integers(1, 100000, size=k)
        elif kind == 3:
            f = 2 ** rng.integers(0, 18, size=n) * (rng.random(n) < 0.5)  # fibonacci-like depth blow-ups
        else:
            f = rng.integers(0, 3, size=n) * rng.integers(0, 262144 // n, size=n)
        f = np.minimum(f, 262144).astype(np.uint32)
        la, lb = oracle.length_table(f, cap), ref.length_table(f, cap)
        assert np.array_equal(la, lb), (t, kind)
        assert la.max(initial=0) <= cap
        assert np.array_equal(oracle.encode_table(la, cap), ref.encode_table(lb, cap))
