"""where does the start-up of a process that uses the library go? (run on a GPU box)"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
t0 = time.perf_counter()
import libzling_b200  # noqa: E402
L = libzling_b200.load()
t1 = time.perf_counter(); print("import + dlopen           %.3f s" % (t1 - t0))
L.zlb_host_alloc.restype = C.c_void_p
p = L.zlb_host_alloc(C.c_size_t(4096))
t2 = time.perf_counter(); print("first CUDA call (init)    %.3f s" % (t2 - t1))
c1 = libzling_b200.Context(device=0, max_blocks=1)
t3 = time.perf_counter(); print("zlb_create(1 block)       %.3f s" % (t3 - t2))
c8 = libzling_b200.Context(device=0, max_blocks=8)
t4 = time.perf_counter(); print("zlb_create(8 blocks)      %.3f s" % (t4 - t3))
b = libzling_b200.PinnedBuffer(134217728)
t5 = time.perf_counter(); print("page-locked 134 MB        %.3f s" % (t5 - t4))
