// umgen_preload: force-load every kernel of libumgen_sm100 (kept out of capi.cu so that the tools library can link the error plumbing alone).
#include "common.cuh"
#include "../../include/umgen.h"

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
