"""CPU-side checks of the boundary: the C-ABI library builds for sm_100a without a GPU, loads, exports every
symbol include/zlb.h declares, keeps the reference's C++ symbol names, and refuses to run without a device."""
import os
import re
import subprocess

import pytest

import libzling_b200
from libzling_b200 import build as zbuild

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    zbuild.build()
    return libzling_b200.load()


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "zlb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(zlb_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(libzling_b200.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_cxx_symbols_match_reference_abi():
    """mangled names a binary linked against the reference's libzling.so needs"""
    out = subprocess.check_output(["nm", "-D", "--defined-only", libzling_b200.lib_path()]).decode()
    for sym in ("_ZN5baidu5zling6EncodeEPNS0_8InputterEPNS0_9OutputterEPNS0_13ActionHandlerEi",
                "_ZN5baidu5zling6DecodeEPNS0_8InputterEPNS0_9OutputterEPNS0_13ActionHandlerE",
                "_ZN5baidu5zling8Inputter7GetCharEv", "_ZN5baidu5zling8Inputter9GetUInt32Ev",
                "_ZN5baidu5zling9Outputter7PutCharEi", "_ZN5baidu5zling9Outputter9PutUInt32Ej",
                "_ZN5baidu5zling12FileInputter7GetDataEPhm", "_ZN5baidu5zling12FileInputter5IsEndEv",
                "_ZN5baidu5zling12FileInputter5IsErrEv", "_ZN5baidu5zling12FileInputter12GetInputSizeEv",
                "_ZN5baidu5zling13FileOutputter7PutDataEPhm", "_ZN5baidu5zling13FileOutputter5IsErrEv",
                "_ZN5baidu5zling13FileOutputter13GetOutputSizeEv",
                "_ZTVN5baidu5zling12FileInputterE", "_ZTIN5baidu5zling13FileOutputterE"):
        assert sym in out, sym


def test_sass_is_sm100a():
    out = subprocess.check_output(["cuobjdump", "-lelf", libzling_b200.lib_path()]).decode()
    assert "sm_100a" in out


@pytest.mark.skipif(not os.path.exists("/root/reference/demo/zling.cpp"), reason="reference tree not on this machine")
def test_reference_demo_links_unmodified(tmp_path, lib):
    """acceptance of the drop-in boundary: the reference CLI compiles against OUR headers and links OUR library"""
    exe = tmp_path / "zling_demo"
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-I", os.path.join(ROOT, "include"), "/root/reference/demo/zling.cpp",
                           "-o", str(exe), "-L", os.path.dirname(libzling_b200.lib_path()), "-lzling",
                           "-Wl,-rpath," + os.path.dirname(libzling_b200.lib_path())])
    assert exe.exists()


def test_no_cpu_fallback(lib):
    if lib.zlb_device_count() > 0:
        pytest.skip("a GPU is visible here")
    with pytest.raises(libzling_b200.ZlingError):
        libzling_b200.Context()
    assert lib.zlb_create(0, 1) is None
    assert b"no CUDA device" in lib.zlb_last_error()


def test_product_never_touches_oracle():
    """nothing under libzling_b200/ may import, link or load oracle/ (tier rule 3)"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "libzling_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in txt.lower() or f == "corpus.py", os.path.join(dirpath, f)
