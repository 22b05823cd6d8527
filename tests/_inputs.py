"""Named, seeded test inputs: the edge cases SURVEY.md §8c lists plus compressible/periodic/adversarial data
(the reference's own fuzz test only feeds uniform random bytes, test/fuzzy/libzling_fuzzy.py:20-42)."""
import numpy as np

from libzling_b200 import corpus

BLOCK = 16777216


def _rng(seed):
    return np.random.default_rng(seed)


def small_cases():
    """(name, bytes) — each encodes in well under a second on the CPU oracle"""
    r = _rng(7)
    text = corpus.enwik8_shaped(1 << 20, seed=3).tobytes()
    words = corpus.ascii_words(1 << 20, seed=1).tobytes()
    cases = [
        ("empty", b""), ("one", b"a"), ("two", b"ab"), ("three", b"abc"),
        ("len275", text[:275]), ("len276", text[:276]), ("len277", text[:277]), ("len300", text[:300]),
        ("a1000", b"a" * 1000), ("zeros100k", b"\0" * 100000), ("bytes256x64", bytes(range(256)) * 64),
        ("abcd5000", b"abcd" * 5000), ("hello", b"hello " * 7 + b"hello\n"),
        ("period2", b"xy" * 40000), ("period3", b"xyz" * 30000), ("period5", b"abcde" * 20000),
        ("period7_long", bytes(range(7)) * 60000),
        ("zero_words", b"\0\0" * 5000 + b"ab" * 100 + b"\0" * 700),
        ("text64k", text[:65536]), ("text1m", text), ("words1m", words),
        ("random64k", r.integers(0, 256, size=65536, dtype=np.uint8).tobytes()),
        ("random600k", r.integers(0, 256, size=600000, dtype=np.uint8).tobytes()),
        ("lowentropy", r.integers(0, 4, size=400000, dtype=np.uint8).tobytes()),
        ("two_symbols_runs", np.repeat(r.integers(0, 2, size=20000, dtype=np.uint8) + 65, r.integers(1, 40, size=20000)).tobytes()),
        # > 4096 inserts into one context: ring wrap + stale hash heads (App. A.5 / B.7)
        ("ringwrap", (b"e" + r.integers(97, 123, size=3, dtype=np.uint8).tobytes()) * 1 + b"".join(
            b"e" + bytes(r.integers(97, 101, size=int(r.integers(1, 6)), dtype=np.uint8)) for _ in range(60000))),
        ("text_random_text", text[:400000] + r.integers(0, 256, size=700000, dtype=np.uint8).tobytes() + text[400000:900000]),
        ("long_matches", (text[:5000] * 40) + text[5000:9000] + (text[100:3000] * 30)),
        ("binary_records", corpus.mixed(1 << 20, seed=11).tobytes()),
    ]
    return cases


def block_boundary_cases():
    """inputs that cross 16 MiB block boundaries (MTF carry, level carry): (name, bytes)"""
    r = _rng(9)
    base = corpus.enwik8_shaped(BLOCK + 70000, seed=5)
    rnd = r.integers(0, 256, size=900000, dtype=np.uint8)
    mixed_tail = np.concatenate([base[:BLOCK - 500000], rnd, base[BLOCK - 500000:BLOCK - 500000 + 1200000]])
    return [
        ("blk_exact", base[:BLOCK].tobytes()),
        ("blk_minus1", base[:BLOCK - 1].tobytes()),
        ("blk_plus1", base[:BLOCK + 1].tobytes()),
        ("blk_plus_tail", base.tobytes()),
        ("random_across_boundary", mixed_tail.tobytes()),
    ]
