/*
 * The interface declared in this header re-states the public API of zling (richox/libzling, BSD 3-clause).  The notice of the
 * interface's origin is retained as the licence asks:
 *
 * zling:
 *  light-weight lossless data compression utility.
 *
 * Copyright (C) 2012-2013 by Zhang Li <zhangli10 at baidu.com>
 * All rights reserved.
 *
 * Redistribution and use in source and binary forms, with or without
 * modification, are permitted provided that the following conditions
 * are met:
 * 1. Redistributions of source code must retain the above copyright
 *    notice, this list of conditions and the following disclaimer.
 * 2. Redistributions in binary form must reproduce the above copyright
 *    notice, this list of conditions and the following disclaimer in the
 *    documentation and/or other materials provided with the distribution.
 * 3. Neither the name of the project nor the names of its contributors
 *    may be used to endorse or promote products derived from this software
 *    without specific prior written permission.
 *
 * THIS SOFTWARE IS PROVIDED BY THE PROJECT AND CONTRIBUTORS ``AS IS'' AND
 * ANY EXPRESS OR IMPLIED WARRANTIES, INCLUDING, BUT NOT LIMITED TO, THE
 * IMPLIED WARRANTIES OF MERCHANTABILITY AND FITNESS FOR A PARTICULAR PURPOSE
 * ARE DISCLAIMED.  IN NO EVENT SHALL THE PROJECT OR CONTRIBUTORS BE LIABLE
 * FOR ANY DIRECT, INDIRECT, INCIDENTAL, SPECIAL, EXEMPLARY, OR CONSEQUENTIAL
 * DAMAGES (INCLUDING, BUT NOT LIMITED TO, PROCUREMENT OF SUBSTITUTE GOODS
 * OR SERVICES; LOSS OF USE, DATA, OR PROFITS; OR BUSINESS INTERRUPTION)
 * HOWEVER CAUSED AND ON ANY THEORY OF LIABILITY, WHETHER IN CONTRACT, STRICT
 * LIABILITY, OR TORT (INCLUDING NEGLIGENCE OR OTHERWISE) ARISING IN ANY WAY
 * OUT OF THE USE OF THIS SOFTWARE, EVEN IF ADVISED OF THE POSSIBILITY OF
 * SUCH DAMAGE.
 */
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
