#!/bin/bash
# builds the shipped library and, next to it, the profiling variant (libzling_prof.so, -DZL_V4_PROFILE=1; not loaded unless copied over)
cd "$(dirname "$0")/.." || exit 1
python -m libzling_b200.build | tail -1
(cd libzling_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=default -shared -cudart static -DZL_V4_PROFILE=1 -o ../libzling_prof.so zl_engine.cu zl_api.cpp 2>&1 | grep -i "error")
ls -la libzling_b200/*.so
