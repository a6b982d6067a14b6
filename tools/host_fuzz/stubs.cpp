// Stubs for the entry points csrc/host.cpp calls into kernels.cu, so that host.cpp alone can be built with AddressSanitizer +
// UBSan and fuzzed on the CPU (tools/host_fuzz/run.sh).  TEST TOOLING: never linked into libtmpt.so.
#include <cstdarg>
#include <cstdio>
#include <string>
#include "../../toymeshpathtracer_b200/csrc/common.h"
static thread_local std::string g_err;
namespace tmpt {
int fail(int status, const char* fmt, ...) { char b[512]; va_list a; va_start(a, fmt); vsnprintf(b, sizeof b, fmt, a); va_end(a); g_err = b; return status; }
void count_launch(uint64_t) {}
}
extern "C" {
const char* tmpt_last_error(void) { return g_err.c_str(); }
int tmpt_scene_create(const float*, int, int, unsigned, tmpt_scene**) { return TMPT_ERR_CUDA; }
void tmpt_scene_destroy(tmpt_scene*) {}
int tmpt_render_multi(tmpt_scene* const*, int, const tmpt_camera*, int, int, int, uint8_t*, uint64_t*, double*) { return TMPT_ERR_CUDA; }
int tmpt_device_count(void) { return 0; }
}
