"""Host replay of the v3 parse kernel's phases (tests/cxx/parse_v3_sim.cu) against the oracle's tokeniser: the
speculate / resolve / apply algorithm of libzling_b200/csrc/zl_parse_v3.cuh is scalar host+device code, so its
logic is checked here on the CPU; the -m gpu parity tests then only have the kernel's synchronisation left to prove."""
import os
import shutil
import subprocess

import pytest

from _inputs import small_cases

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def sim(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    d = tmp_path_factory.mktemp("v3sim")
    obj = str(d / "oracle.o")
    exe = str(d / "parse_v3_sim")
    subprocess.check_call(["gcc", "-std=c11", "-O2", "-c", os.path.join(ROOT, "oracle", "zling_oracle.c"), "-o", obj])
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                           os.path.join(HERE, "cxx", "parse_v3_sim.cu"), obj])
    return exe, d


@pytest.mark.parametrize("level", [0, 2, 4])
def test_v3_phases_match_oracle(sim, level):
    exe, d = sim
    cases = dict(small_cases())
    for name in ("empty", "one", "three", "len275", "len277", "text64k", "random64k", "ringwrap", "zero_words", "period3", "two_symbols_runs"):
        path = str(d / (name + ".bin"))
        with open(path, "wb") as f:
            f.write(bytes(cases[name]))
        for order in (0, 1):
            r = subprocess.run([exe, path, str(level), str(order)], capture_output=True, text=True)
            assert r.returncode == 0, (name, level, order, r.stdout, r.stderr)


def test_v3_level_switches_inside_a_block(sim):
    """level feedback: sub-blocks of one block parsed at different levels (plan digits per sub-block)"""
    exe, d = sim
    cases = dict(small_cases())
    path = str(d / "trt.bin")
    with open(path, "wb") as f:
        f.write(bytes(cases["text_random_text"]))
    for level, plan in ((2, "2202"), (4, "40404"), (3, "03030")):
        r = subprocess.run([exe, path, str(level), "0", plan], capture_output=True, text=True)
        assert r.returncode == 0, (level, plan, r.stdout, r.stderr)
