#!/usr/bin/env python
"""e0..e4 table through the reference's own CLI (demo/zling.cpp, UNMODIFIED): once linked with the reference library
(oracle/_ref/zling_demo), once compiled against this repo's headers and linked with the GPU library (oracle/_ref/zling_demo_b200;
both built by __graft_entry__.build() where the reference tree exists).  Same columns as the reference's benchmark script: encode
time, decode time, compressed size, PASS = the decoded file equals the input; plus whether the two CLIs wrote identical bytes.
Wall clock of the whole process (CUDA context creation and page-locked allocations included for the GPU CLI).

    python scripts/cli_table.py [size_mb] [out.md]"""
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libzling_b200 import corpus  # noqa: E402

sizes = [float(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "100").split(",")]     # the first size runs e0..e4, the others e0 only
out_md = sys.argv[2] if len(sys.argv) > 2 else None
ref, gpu = os.path.join(ROOT, "oracle", "_ref", "zling_demo"), os.path.join(ROOT, "oracle", "_ref", "zling_demo_b200")
for exe in (ref, gpu):
    if not os.path.exists(exe):
        raise SystemExit("missing %s (run __graft_entry__.build() where /root/reference exists)" % exe)
tmp = "/tmp/zl_cli_table"
os.makedirs(tmp, exist_ok=True)
src = os.path.join(tmp, "in.bin")


def run(exe, mode, a, b):
    """returns (wall seconds of the process, seconds the CLI itself reports: its clock starts when the library's OnInit runs)"""
    t0 = time.perf_counter()
    r = subprocess.run([exe, mode, a, b], capture_output=True, text=True)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise SystemExit("%s %s failed: %s" % (exe, mode, r.stderr[-300:]))
    m = re.findall(r"time=([0-9.]+) sec", r.stderr)
    return dt, (float(m[-1]) if m else float("nan"))


def same(a, b):
    return subprocess.run(["cmp", "-s", a, b]).returncode == 0


lines = ["| file | level | CLI | encode s (process wall) | encode s (CLI's own clock) | MB/s (own clock) | decode s (wall) | decode s (own clock) | size | PASS | bytes identical to the reference CLI's |",
         "|---|---|---|---|---|---|---|---|---|---|---|"]
for si, size_mb in enumerate(sizes):
    corpus.enwik8_shaped(int(size_mb * 1e6), seed=8).tofile(src)
    if si == 0:
        run(gpu, "e0", src, os.path.join(tmp, "warm.zl"))          # first touch of the driver on a fresh box, not reported
    for level in (range(5) if si == 0 else (0,)):
        outs = {}
        for name, exe in (("reference (CPU)", ref), ("this repo (B200)", gpu)):
            z, back = os.path.join(tmp, "%s.zl" % name[:4]), os.path.join(tmp, "%s.out" % name[:4])
            te, te_own = run(exe, "e%d" % level, src, z)
            td, td_own = run(exe, "d", z, back)
            outs[name] = z
            ident = "-" if exe == ref else ("yes" if same(z, outs["reference (CPU)"]) else "NO")
            lines.append("| %.0f MB | e%d | %s | %.2f | %.2f | %.1f | %.2f | %.2f | %d | %s | %s |" % (
                size_mb, level, name, te, te_own, size_mb / te_own, td, td_own, os.path.getsize(z), "PASS" if same(back, src) else "FAIL", ident))
txt = "\n".join(lines) + "\n"
print(txt)
if out_md:
    with open(out_md, "w") as f:
        f.write("enwik8-shaped files; process wall clock per CLI invocation (CUDA start-up of the process included: about 1.2 s on this box) and the time the\n"
                "CLI itself prints (demo/zling.cpp: its clock starts in OnInit)\n\n" + txt)
