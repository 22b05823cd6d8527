"""libzling_b200 — Blackwell-native (sm_100a) ROLZ+Huffman block pipeline behind libzling's API.

  csrc/        CUDA kernels, host engine, C ABI (include/zlb.h), C++ drop-in API (include/libzling/*.h)
  build.py     in-tree nvcc build of libzling.so
  api.py       ctypes binding of the C ABI
  corpus.py    seeded synthetic inputs of BASELINE.json's configs
"""
from .api import Context, Encoder, Decoder, Comm, PinnedBuffer, ZlingError, FormatError, load, lib_path, EXPORTS, BLOCK  # noqa: F401
