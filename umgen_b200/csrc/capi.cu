// C-ABI plumbing shared by every translation unit of libumgen_sm100.
#include <stdarg.h>

#include "common.cuh"
#include "../../include/umgen.h"

namespace umgen {
static thread_local char g_err[512] = "";
int64_t g_launches = 0;
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace umgen

extern "C" int umgen_abi_version(void) { return UMGEN_ABI_VERSION; }
extern "C" const char* umgen_last_error(void) { return umgen::g_err; }
extern "C" int64_t umgen_launch_count(void) { return umgen::g_launches; }
