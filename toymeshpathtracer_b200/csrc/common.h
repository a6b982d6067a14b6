// common.h -- shared between the CUDA translation unit (kernels.cu) and the plain C++ host
// glue (host.cpp): error reporting and the launch counter behind the C ABI.
#pragma once
#include <stdint.h>

#include "../../include/tmpt.h"

namespace tmpt {
// printf-style; stores the message for tmpt_last_error() (thread-local) and returns `status`
int fail(int status, const char* fmt, ...);
void count_launch(uint64_t n = 1);
}  // namespace tmpt
