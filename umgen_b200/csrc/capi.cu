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

namespace umgen {
int preload_tar(); int preload_gemm(); int preload_attn(); int preload_decode(); int preload_decode_cluster(); int preload_vq();
}
// Load every kernel of the library on the current device now (see the note on lazy module loading in the translation units).
extern "C" int umgen_preload(void) {
    using namespace umgen;
    if (int rc = preload_tar()) return rc;
    if (int rc = preload_gemm()) return rc;
    if (int rc = preload_attn()) return rc;
    if (int rc = preload_decode()) return rc;
    if (int rc = preload_decode_cluster()) return rc;
    return preload_vq();
}
