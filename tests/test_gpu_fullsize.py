"""GPU parity at BASELINE.json's sizes (one GPU): the 100 MB enwik8-shaped stream at e0 / e4 and a 100 MB slice of the
mixed text+binary+random stream at e2 (level feedback across block boundaries), bit-exact against the CPU checker, plus
the size-independent property encode -> decode == identity; and the split submit / set_state / complete path that
one stream sharded over several GPUs uses (two block ranges on one GPU here)."""
import numpy as np
import pytest

import libzling_b200
from libzling_b200 import corpus, sharded

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx6():
    c = libzling_b200.Context(device=0, max_blocks=6)
    yield c
    c.close()


@pytest.mark.parametrize("level", [0, 4])
def test_enwik8_shaped_100mb_bit_exact(ctx6, oracle, level):
    data = corpus.enwik8_shaped(100_000_000, seed=8)
    z = ctx6.encode(data, level)
    assert z == oracle.encode(data, level)
    if level == 0:
        assert ctx6.decode(z) == data.tobytes()


def test_mixed_100mb_e2_bit_exact_and_round_trip(ctx6, oracle):
    data = corpus.mixed(100_000_000, seed=4)
    z = ctx6.encode(data, 2)
    assert z == oracle.encode(data, 2)
    assert ctx6.stats()["reparsed_blocks"] >= 0
    assert ctx6.decode(z) == data.tobytes()


def test_split_submit_complete_matches_one_call(ctx6, oracle):
    """rank 0 owns blocks [0, 2), rank 1 owns blocks [2, 4): rank 1 submits first (parse with the guessed level), then
    receives rank 0's carried state, completes; concatenation == the single-stream encoding"""
    data = corpus.mixed(3 * 16777216 + 54321, seed=11)
    want = oracle.encode(data, 2)
    (lo0, hi0), (lo1, hi1) = sharded.block_ranges(data.size, 2)
    other = libzling_b200.Context(device=0, max_blocks=2)
    try:
        e0 = libzling_b200.Encoder(ctx6, 2)
        e1 = libzling_b200.Encoder(other, 2)
        e1.submit(data[lo1:hi1])
        e0.submit(data[lo0:hi0])
        out0 = e0.complete()
        e1.set_state(e0.get_state())
        out1 = e1.complete()
        e0.close(); e1.close()
    finally:
        other.close()
    assert out0 + out1 == want


def test_split_submit_with_wrong_level_guess_reparses(ctx6, oracle):
    """the range boundary falls inside incompressible data: the carried level is 0, not the requested one"""
    rng = np.random.default_rng(3)
    data = np.concatenate([corpus.enwik8_shaped(16777216 - 300000, seed=5), rng.integers(0, 256, 600000, dtype=np.uint8),
                           corpus.enwik8_shaped(2000000, seed=6)])
    want = oracle.encode(data, 3)
    e0 = libzling_b200.Encoder(ctx6, 3)
    out0 = e0.encode_blocks(data[:16777216])
    state = e0.get_state()
    e0.close()
    assert int(np.frombuffer(state[65536:].tobytes(), dtype=np.int32)[0]) == 0
    e1 = libzling_b200.Encoder(ctx6, 3)
    e1.submit(data[16777216:])
    e1.set_state(state)
    out1 = e1.complete()
    e1.close()
    assert out0 + out1 == want
