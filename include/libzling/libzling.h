// libzling/libzling.h — public entry points of zling, B200 build.
//
// Drop-in for the reference's src/libzling.h:44-45: same signatures, same mangled names, same return / throw
// contract (0 ok, -1 if the inputter or outputter reports an error; std::runtime_error on a malformed
// stream).  Behind them the 16 MiB block pipeline runs as sm_100a CUDA kernels through the C ABI in zlb.h;
// there is no CPU path: without a CUDA device both functions throw std::runtime_error.
// One deliberate difference: an out-of-range `level` makes the reference spin forever
// (src/libzling_lz.cpp:136 + src/libzling.cpp:199); here Encode returns -1 without touching the streams.
#ifndef LIBZLING_B200_H
#define LIBZLING_B200_H

#include "libzling_inc.h"
#include "libzling_utils.h"

namespace baidu {
namespace zling {

int Encode(Inputter* inputter, Outputter* outputter, ActionHandler* action_handler = NULL, int level = 0);
int Decode(Inputter* inputter, Outputter* outputter, ActionHandler* action_handler = NULL);

}  // namespace zling
}  // namespace baidu
#endif  // LIBZLING_B200_H
