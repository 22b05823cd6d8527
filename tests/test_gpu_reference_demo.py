"""the reference's own CLI (demo/zling.cpp), compiled UNMODIFIED against this repo's headers and
linked with this repo's libzling.so (oracle/_ref/zling_demo_b200, built by __graft_entry__.build() where the reference
tree exists), run on the GPU: its output must be the reference's bytes, and it must decode its own output."""
import os
import subprocess

import pytest

from _inputs import small_cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "oracle", "_ref", "zling_demo_b200")


@pytest.mark.skipif(not os.path.exists(DEMO), reason="oracle/_ref/zling_demo_b200 not built (needs the reference tree at build time)")
def test_unmodified_reference_cli_on_the_gpu_library(tmp_path, oracle):
    cases = dict(small_cases())
    for name, level in (("text1m", 0), ("text_random_text", 4), ("hello", 2)):
        data = cases[name]
        src, z, back = tmp_path / "in.bin", tmp_path / "out.zl", tmp_path / "back.bin"
        src.write_bytes(data)
        r = subprocess.run([DEMO, "e%d" % level, str(src), str(z)], capture_output=True, timeout=300)
        assert r.returncode == 0, r.stderr
        assert z.read_bytes() == oracle.encode(data, level), (name, level)
        r = subprocess.run([DEMO, "d", str(z), str(back)], capture_output=True, timeout=300)
        assert r.returncode == 0, r.stderr
        assert back.read_bytes() == data, (name, level)
