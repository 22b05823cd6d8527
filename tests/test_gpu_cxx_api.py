"""The C++ drop-in API (baidu::zling::Encode/Decode with File{In,Out}putter and an ActionHandler) on the GPU:
a small program written against include/libzling/libzling.h, linked with libzling.so, must produce the oracle's
bytes and round-trip.  (On the container with /root/reference, test_abi_cpu additionally links the reference's
own CLI against this library.)"""
import os
import subprocess

import pytest

import libzling_b200
from _inputs import small_cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cli(tmp_path_factory):
    d = tmp_path_factory.mktemp("cxx")
    exe = d / "zl_cli"
    libdir = os.path.dirname(libzling_b200.lib_path())
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cxx", "zl_cli.cpp"),
                           "-o", str(exe), "-L", libdir, "-lzling", "-Wl,-rpath," + libdir])
    return str(exe), d


def test_cxx_encode_decode(cli, oracle):
    exe, d = cli
    cases = dict(small_cases())
    for name in ("empty", "one", "hello", "text1m", "text_random_text"):
        data = cases[name]
        src, z, back = d / "in.bin", d / "out.zl", d / "back.bin"
        src.write_bytes(data)
        for level in (0, 3):
            r = subprocess.run([exe, "e%d" % level, str(src), str(z)], capture_output=True)
            assert r.returncode == 0, r.stderr
            assert z.read_bytes() == oracle.encode(data, level), (name, level)
            # handler bookkeeping printed by the CLI: one OnProcess per 16 MiB block, after its bytes were written
            assert b"blocks=%d" % ((len(data) + (1 << 24) - 1) >> 24) in r.stderr, r.stderr
            r = subprocess.run([exe, "d", str(z), str(back)], capture_output=True)
            assert r.returncode == 0, r.stderr
            assert back.read_bytes() == data, (name, level)


def test_cxx_streaming_pipeline(cli, oracle):
    """an input of several batches (ZLING_B200_BLOCKS=2: batches of two blocks) goes through the streaming pipeline of the driver — the
    next batch is read while the current one is parsed, frames are written while the next one runs — and must still be the
    reference's bytes, with OnProcess called once per block in order; also with an input that ends exactly at a batch boundary"""
    import numpy as np
    from libzling_b200 import corpus
    exe, d = cli
    env = dict(os.environ, ZLING_B200_BLOCKS="2")
    for nbytes, level in ((4 * (1 << 24) + 12345, 0), (4 * (1 << 24), 2)):
        rng = np.random.default_rng(5)
        data = np.concatenate([corpus.enwik8_shaped(nbytes - (3 << 20), seed=21), rng.integers(0, 256, 3 << 20, dtype=np.uint8)]).tobytes()
        src, z, back = d / "big.bin", d / "big.zl", d / "big.out"
        src.write_bytes(data)
        r = subprocess.run([exe, "e%d" % level, str(src), str(z)], capture_output=True, env=env)
        assert r.returncode == 0, r.stderr
        assert z.read_bytes() == oracle.encode(data, level), (nbytes, level)
        assert b"blocks=%d" % ((len(data) + (1 << 24) - 1) >> 24) in r.stderr, r.stderr
        r = subprocess.run([exe, "d", str(z), str(back)], capture_output=True, env=env)
        assert r.returncode == 0, r.stderr
        assert back.read_bytes() == data


def test_cxx_decode_malformed_throws(cli, oracle):
    exe, d = cli
    z = bytearray(oracle.encode(b"hello " * 50, 0))
    z[0] = 7
    (d / "bad.zl").write_bytes(bytes(z))
    r = subprocess.run([exe, "d", str(d / "bad.zl"), str(d / "bad.out")], capture_output=True)
    assert r.returncode == 3 and b"invalid encflag" in r.stderr


def test_cxx_api_is_reentrant(tmp_path_factory):
    """two and four threads inside ONE process call Encode/Decode concurrently (the reference is re-entrant: every call owns its
    EncodeResource/DecodeResource, src/libzling.cpp:108-163); results must equal the solo runs"""
    d = tmp_path_factory.mktemp("cxxthr")
    exe = d / "zl_threads"
    libdir = os.path.dirname(libzling_b200.lib_path())
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-pthread", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cxx", "zl_threads.cpp"),
                           "-o", str(exe), "-L", libdir, "-lzling", "-Wl,-rpath," + libdir])
    data = dict(small_cases())["text_random_text"] + dict(small_cases())["text1m"]
    (d / "in.bin").write_bytes(data)
    for n in (2, 4):
        r = subprocess.run([str(exe), str(d / "in.bin"), str(n)], capture_output=True)
        assert r.returncode == 0, r.stderr
