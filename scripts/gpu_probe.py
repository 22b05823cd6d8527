"""quick GPU timing / parity probe: python scripts/gpu_probe.py [MiB ...]"""
import sys
import time

sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import libzling_b200
from libzling_b200 import corpus
from _libs import Oracle

sizes = [int(a) for a in sys.argv[1:]] or [1, 4]
o = Oracle()
ctx = libzling_b200.Context(0, 2)
for mib in sizes:
    data = corpus.enwik8_shaped(mib << 20, seed=3).tobytes()
    for lv in (0, 2, 4):
        t = time.time(); z = ctx.encode(data, lv); dt = time.time() - t
        st = ctx.stats()
        ok = z == o.encode(data, lv)
        print("n=%dMiB lv=%d ok=%s t=%.3fs parse=%.1fms mtf=%.1fms build=%.1fms pack=%.2fms tokens=%d slow_main=%d slow_lazy=%d winhits=%d windows=%d cyc_spec=%.1fM cyc_res=%.1fM general=%d" % (
            mib, lv, ok, dt, st["ms_parse"], st["ms_mtf"], st["ms_huff_build"], st["ms_pack"], st["tokens"], st["slow_main"], st["slow_lazy"],
            st["window_hits"], st["windows"], st["cyc_spec"] / 1e6, st["cyc_resolve"] / 1e6, st["general_path"]), flush=True)
        if not ok:
            want = o.parse_block(data, lv)
            subs = ctx.debug_subblocks(0)
            print("   subs gpu:", [(s["enc_end"], s["rlen"]) for s in subs][:4], " oracle:", [(w["encpos"], w["syms"].size) for w in want][:4])
        t = time.time(); r = ctx.decode(z); dt = time.time() - t
        print("   decode ok=%s t=%.3fs" % (r == data, dt), flush=True)
