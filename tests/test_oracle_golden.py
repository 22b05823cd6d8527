"""oracle (C restatement) against the committed golden vectors that the unmodified reference produced
(tests/golden/make_golden.py).  Runs everywhere — this is what pins the oracle on machines without /root/reference."""
import hashlib
import json
import os

import pytest

from _inputs import small_cases, block_boundary_cases
from _libs import walk_container

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(GOLD, "golden.json")) as f:
    GOLDEN = json.load(f)


def _check(oracle, cases, levels):
    for name, data in cases:
        g = GOLDEN[name]
        assert g["size"] == len(data) and g["md5"] == hashlib.md5(data).hexdigest(), "input generator drifted: " + name
        for level in levels:
            z = oracle.encode(data, level)
            gl = g["levels"][str(level)]
            assert len(z) == gl["size"], (name, level)
            assert hashlib.md5(z).hexdigest() == gl["md5"], (name, level)
            assert [[b, e, r, o] for (b, e, r, o, _) in walk_container(z)][:64] == gl["subblocks"]
            assert oracle.decode(z, len(data)) == data


def test_oracle_small_all_levels(oracle):
    _check(oracle, small_cases(), range(5))


def test_oracle_block_boundaries(oracle):
    _check(oracle, block_boundary_cases(), (0, 2, 4))


@pytest.mark.parametrize("fname", sorted(f for f in os.listdir(GOLD) if f.endswith(".zl")))
def test_oracle_decodes_reference_streams(oracle, fname):
    name = fname.split(".")[0]
    data = dict(small_cases())[name]
    with open(os.path.join(GOLD, fname), "rb") as f:
        z = f.read()
    assert oracle.decode(z, len(data)) == data


def test_malformed_streams_rejected(oracle):
    z = bytearray(oracle.encode(b"hello " * 50, 0))
    bad = bytes([2]) + bytes(z[1:])                      # invalid flag (src/libzling.cpp:315-317)
    with pytest.raises(ValueError):
        oracle.decode(bad, 1000)
    bad = bytearray(z); bad[5:9] = (300000).to_bytes(4, "big")   # rlen > 262144 (src/libzling.cpp:326-328)
    with pytest.raises(ValueError):
        oracle.decode(bytes(bad), 1000)
    bad = bytearray(z); bad[1:5] = (5).to_bytes(4, "big")         # encpos mismatch -> lz decode fails (:406-408)
    with pytest.raises(ValueError):
        oracle.decode(bytes(bad), 1000)
