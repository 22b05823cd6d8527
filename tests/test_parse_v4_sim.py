"""Host replay of the v4 parse kernel's phases (tests/cxx/parse_v4_sim.cu) against the oracle's tokeniser: the
speculate / iterate-to-a-fixed-point / finalize algorithm of libzling_b200/csrc/zl_parse_v4.cuh is scalar host+device
code, so its logic is checked here on the CPU; the -m gpu parity tests then have the kernel's synchronisation and its
parallel forms of the ordered passes left to prove.  Also a seeded fuzz with match-heavy generators (the reference's
own fuzz script feeds uniform random bytes only, test/fuzzy/libzling_fuzzy.py:20-42)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from _inputs import small_cases

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def sim(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    d = tmp_path_factory.mktemp("v4sim")
    obj = str(d / "oracle.o")
    exe = str(d / "parse_v4_sim")
    subprocess.check_call(["gcc", "-std=c11", "-O2", "-c", os.path.join(ROOT, "oracle", "zling_oracle.c"), "-o", obj])
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                           os.path.join(HERE, "cxx", "parse_v4_sim.cu"), obj])
    return exe, d


@pytest.mark.parametrize("level", [0, 2, 4])
def test_v4_phases_match_oracle(sim, level):
    exe, d = sim
    cases = dict(small_cases())
    for name in ("empty", "one", "three", "len275", "len277", "text64k", "random64k", "ringwrap", "zero_words", "period3",
                 "two_symbols_runs", "zeros100k", "long_matches", "lowentropy"):
        path = str(d / (name + ".bin"))
        with open(path, "wb") as f:
            f.write(bytes(cases[name]))
        r = subprocess.run([exe, path, str(level)], capture_output=True, text=True)
        assert r.returncode == 0, (name, level, r.stdout, r.stderr)


def test_v4_level_switches_inside_a_block(sim):
    """level feedback: sub-blocks of one block parsed at different levels (plan digits per sub-block, or predicted)"""
    exe, d = sim
    cases = dict(small_cases())
    path = str(d / "trt.bin")
    with open(path, "wb") as f:
        f.write(bytes(cases["text_random_text"]))
    for level, plan in ((2, "2202"), (4, "40404"), (3, "03030"), (2, "auto")):
        r = subprocess.run([exe, path, str(level), plan], capture_output=True, text=True)
        assert r.returncode == 0, (level, plan, r.stdout, r.stderr)


def _case(rng, words):
    kind = int(rng.integers(0, 7))
    n = int(rng.integers(300, 120000))
    if kind == 0:
        return rng.integers(0, int(rng.integers(2, 20)), n).astype(np.uint8) + 65
    if kind == 1:
        p = rng.integers(0, 256, int(rng.integers(1, 40))).astype(np.uint8)
        a = np.tile(p, n // len(p) + 1)[:n].copy()
        k = int(rng.integers(0, n // 50 + 1))
        a[rng.integers(0, n, k)] = rng.integers(0, 256, k)
        return a
    if kind == 2:
        blob = b" ".join(words[int(i)] for i in rng.integers(0, int(rng.integers(5, len(words))), n // 3 + 2))
        return np.frombuffer(blob[:n], dtype=np.uint8).copy()
    if kind == 3:
        return rng.integers(0, 256, n).astype(np.uint8)
    if kind == 4:
        return np.repeat(rng.integers(0, 4, n // 20 + 2).astype(np.uint8), rng.integers(1, 60, n // 20 + 2))[:n].copy()
    if kind == 5:
        a = rng.integers(0, 256, n).astype(np.uint8)
        for _ in range(n // 100):
            ln = int(rng.integers(4, 300))
            s, d = int(rng.integers(0, n - ln)), int(rng.integers(0, n - ln))
            a[d:d + ln] = a[s:s + ln]
        return a
    a = np.empty(n, dtype=np.uint8)
    a[0::2] = 32
    a[1::2] = rng.integers(97, 101, len(a[1::2]))
    return a


def test_v4_seeded_fuzz_matches_checker(sim):
    exe, d = sim
    rng = np.random.default_rng(20261018)
    words = [bytes(rng.integers(97, 123, int(rng.integers(2, 9))).astype(np.uint8)) for _ in range(300)]
    path = str(d / "case.bin")
    for i in range(int(os.environ.get("ZL_FUZZ_CASES", "40"))):
        _case(rng, words).tofile(path)
        level = int(rng.integers(0, 5))
        plan = [[], ["auto"], ["".join(str(int(rng.integers(0, level + 1))) for _ in range(3))]][int(rng.integers(0, 3))]
        r = subprocess.run([exe, path, str(level)] + plan, capture_output=True, text=True)
        assert r.returncode == 0, (i, level, plan, r.stdout[-400:], r.stderr[-400:])
