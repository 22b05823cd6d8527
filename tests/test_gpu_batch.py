"""Multi-stream batch API (zlb_encode_batch / zlb_decode_batch): many independent streams share ONE pass of
the block pipeline (one CTA per 16 MiB block, one MTF chain per (context, stream)); every stream's bytes must equal what the
reference produces for it alone."""
import numpy as np
import pytest

import libzling_b200
from _inputs import small_cases, block_boundary_cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = libzling_b200.Context(device=0, max_blocks=12)
    yield c
    c.close()


@pytest.mark.parametrize("level", [0, 2, 4])
def test_batch_encode_matches_solo_reference(ctx, oracle, level):
    cases = dict(small_cases())
    names = ["text1m", "empty", "random600k", "text_random_text", "one", "binary_records", "long_matches", "zero_words", "lowentropy", "ringwrap"]
    streams = [cases[n] for n in names]
    out = ctx.encode_batch(streams, level)
    for n, z in zip(names, out):
        assert z == oracle.encode(cases[n], level), (n, level)
    back = ctx.decode_batch(out, [len(cases[n]) + 16 for n in names])
    for n, raw in zip(names, back):
        assert raw == cases[n], (n, level)


def test_batch_with_multi_block_streams(ctx, oracle):
    """streams longer than one block inside a batch: the MTF tables and the level are carried inside each stream only"""
    big = dict(block_boundary_cases())
    cases = dict(small_cases())
    streams = [big["blk_plus_tail"], cases["text1m"], big["random_across_boundary"], cases["text_random_text"]]
    out = ctx.encode_batch(streams, 2)
    for s, z in zip(streams, out):
        assert z == oracle.encode(s, 2)
    back = ctx.decode_batch(out, [len(s) + 16 for s in streams])
    assert back == [bytes(s) for s in streams]


def test_batch_rejects_too_many_blocks(ctx):
    blk = libzling_b200.BLOCK
    with pytest.raises(libzling_b200.ZlingError):
        ctx.encode_batch([np.zeros(blk + 1, dtype=np.uint8)] * 7, 0)      # 14 blocks > max_blocks 12
