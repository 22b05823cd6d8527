import time, sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, libzling_b200
from libzling_b200 import corpus
from _libs import Oracle
o = Oracle()
ctx = libzling_b200.Context(0, 2)
for n in (1 << 20, 4 << 20):
    data = corpus.enwik8_shaped(n, seed=3).tobytes()
    for lv in (0, 4):
        t = time.time(); z = ctx.encode(data, lv); dt = time.time() - t
        st = ctx.stats()
        print("n=%d lv=%d ok=%s t=%.3fs parse=%.1fms mtf=%.1fms build=%.1fms pack=%.2fms tokens=%d" % (n, lv, z == o.encode(data, lv), dt, st["ms_parse"], st["ms_mtf"], st["ms_huff_build"], st["ms_pack"], st["tokens"]), flush=True)
        t = time.time(); r = ctx.decode(z); dt = time.time() - t
        print("   decode ok=%s t=%.3fs" % (r == data, dt), flush=True)
