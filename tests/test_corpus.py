"""the synthetic stand-ins must stay inside the acceptance band SURVEY.md §8d gives for 'enwik8-shaped'"""
import zlib

import numpy as np

from libzling_b200 import corpus


def test_enwik8_shaped_band(oracle):
    a = corpus.enwik8_shaped(6 << 20, seed=8)
    assert a.size == 6 << 20
    assert 4.9 <= corpus.order0_entropy(a) <= 5.3
    assert 0.33 <= len(zlib.compress(a.tobytes(), 6)) / a.size <= 0.41
    assert 0.28 <= len(oracle.encode(a, 0)) / a.size <= 0.35
    assert 0.27 <= len(oracle.encode(a, 4)) / a.size <= 0.33
    assert 0.02 <= float((a >= 128).mean()) <= 0.05
    assert np.array_equal(a[:100000], corpus.enwik8_shaped(100000, seed=8))   # deterministic, prefix-stable


def test_ascii_words_roundtrip_config1(oracle):
    """BASELINE.json configs[0]: 1 MB synthetic ASCII, e0, CPU reference encode->decode roundtrip"""
    a = corpus.ascii_words(1 << 20, seed=1)
    assert a.max() < 128
    z = oracle.encode(a, 0)
    assert oracle.decode(z, a.size) == a.tobytes()


def test_mixed_has_all_segment_kinds(oracle):
    a = corpus.mixed(24 << 20, seed=4)
    z = oracle.encode(a, 2)
    assert 0.2 < len(z) / a.size < 0.9
    assert oracle.decode(z, a.size) == a.tobytes()
