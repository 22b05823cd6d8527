// libzling/libzling_inc.h — common includes of the public zling API (B200 build).
// Same role as the reference's src/libzling_inc.h:38-58 (installed as libzling/libzling_inc.h by
// build/CMakeLists.txt:11-14); only standard headers, so user code that relied on them keeps compiling.
#ifndef LIBZLING_B200_INC_H
#define LIBZLING_B200_INC_H

#include <stdint.h>
#include <inttypes.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <queue>
#include <stdexcept>
#include <utility>
#include <vector>

#endif  // LIBZLING_B200_INC_H
