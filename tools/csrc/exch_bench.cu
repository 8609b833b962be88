// Debug microbenchmark: cost of one all-to-all exchange of a 768-value vector between all CTAs through
// L2 with tagged LL lines, optionally while every SM streams weights from HBM with bulk copies.
// Not part of the product path; used to size the decode kernel's phase structure.
#include "../../umgen_b200/csrc/common.cuh"
#include "../../include/umgen.h"

namespace umgen {
constexpr int XB_THREADS = 512;

// flavour 0: relaxed.gpu (strong)   1: .cg (L2 only, weak)   2: ld.cv / st.wt   3: ld.volatile / st.volatile (= strong.sys)
template <int F> __device__ __forceinline__ uint4 xb_ld(const float* p) {
    uint4 r;
    if (F == 0) asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    if (F == 1) asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    if (F == 2) asm volatile("ld.global.cv.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    if (F == 3) asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
template <int F> __device__ __forceinline__ void xb_st(float* p, float v, uint32_t tag) {
    if (F == 0) asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(tag) : "memory");
    if (F == 1) asm volatile("st.global.cg.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(tag) : "memory");
    if (F == 2) asm volatile("st.global.wt.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(tag) : "memory");
    if (F == 3) asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}

// one-way latency: CTA 0 and CTA 1 bounce a tagged line; reports ns per round trip
template <int F> __global__ void pingpong_kernel(float* buf, int iters, unsigned long long* out_ns) {
    if (threadIdx.x != 0 || blockIdx.x > 1) return;
    const int me = blockIdx.x;
    float* mine = buf + me * 64, *other = buf + (1 - me) * 64;
    unsigned long long t0 = globaltimer_ns();
    for (int it = 1; it <= iters; ++it) {
        if (me == 0) {
            xb_st<F>(other, 1.f, it); xb_st<F>(other + 2, 1.f, it);
            uint4 v; do { v = xb_ld<F>(mine); } while (v.y < (uint32_t)it);
        } else {
            uint4 v; do { v = xb_ld<F>(mine); } while (v.y < (uint32_t)it);
            xb_st<F>(other, 1.f, it); xb_st<F>(other + 2, 1.f, it);
        }
    }
    if (me == 0) out_ns[0] = globaltimer_ns() - t0;
}

// variant 0: LL pull with krep replicas.  variant 1: release/acquire flag per writer CTA + plain data.
template <int F> __global__ void __launch_bounds__(XB_THREADS, 1)
exch_bench_kernel(float* buf, uint32_t* flags, int iters, int krep, int variant, const uint8_t* stream_src, int stream_bytes, int nvals,
                  unsigned long long* out_ns) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[2];
    __shared__ float sink[32];
    const int tid = threadIdx.x, cta = blockIdx.x, G = gridDim.x;
    const int warp = tid >> 5;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
    __syncthreads();
    const int r0 = nvals * cta / G, r1 = nvals * (cta + 1) / G;
    const int rep = cta % krep;
    const int vstride = 2 * nvals;
    unsigned long long t0 = 0;
    float accum = 0.f;
    if (warp == 15) {           // streaming warp: keeps two bulk copies of stream_bytes/2 in flight per iteration
        if (tid % 32 == 0 && stream_bytes > 0) {
            const int half = (stream_bytes / 2) & ~15;
            uint32_t ph[2] = {0, 0};
            size_t off = (size_t)cta * stream_bytes;
            for (int it = 0; it < iters; ++it) {
                for (int b = 0; b < 2; ++b) {
                    if (it > 0) { while (!mbar_try_wait(&bars[b], ph[b])) {} ph[b] ^= 1; }
                    mbar_arrive_expect_tx(&bars[b], half);
                    bulk_g2s(smem + b * half, stream_src + off + (size_t)b * half, half, &bars[b]);
                }
                off += (size_t)G * stream_bytes;
                if (off + (size_t)G * stream_bytes > (size_t)1 << 30) off = (size_t)cta * stream_bytes;
            }
            for (int b = 0; b < 2; ++b) while (!mbar_try_wait(&bars[b], ph[b])) {}
        }
        return;
    }
    if (tid == 0) t0 = globaltimer_ns();
    for (int it = 1; it <= iters; ++it) {
        // publish my values
        if (variant == 0) {
            for (int r = r0 + tid; r < r1; r += 480)
                for (int k = 0; k < krep; ++k) xb_st<F>(buf + k * vstride + 2 * r, (float)(r + it), (uint32_t)it);
        } else {
            for (int r = r0 + tid; r < r1; r += 480) buf[r] = (float)(r + it);
            asm volatile("bar.sync 1, 480;" ::: "memory");
            if (tid == 0) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + cta * 32), "r"((uint32_t)it) : "memory"); }
        }
        // gather everything
        float s = 0.f;
        if (variant == 0) {
            for (int line = tid; line < nvals / 2; line += 480) {
                uint4 v;
                do { v = xb_ld<F>(buf + rep * vstride + 4 * line); } while (v.y < (uint32_t)it || v.w < (uint32_t)it);   // a writer may run one round ahead
                s += __uint_as_float(v.x) + __uint_as_float(v.z);
            }
        } else {
            if (tid < G) { while (ld_acquire_gpu(flags + tid * 32) < (uint32_t)it) {} }
            asm volatile("bar.sync 1, 480;" ::: "memory");
            for (int i = tid; i < nvals; i += 480) s += __ldcg(buf + i);
        }
        accum += s;
        asm volatile("bar.sync 1, 480;" ::: "memory");
    }
    if (tid % 32 == 0) sink[warp] = accum;
    if (tid == 0) { out_ns[cta] = globaltimer_ns() - t0; if (accum == 123.456f) out_ns[cta] = 0; }
}
}  // namespace umgen

extern "C" int umgen_debug_pingpong(void* buf_f, int iters, int flavour, void* out_ns_u64, void* stream_v) {
    using namespace umgen;
    void* args[] = {&buf_f, &iters, &out_ns_u64};
    void* fn = flavour == 0 ? (void*)pingpong_kernel<0> : flavour == 1 ? (void*)pingpong_kernel<1> : flavour == 2 ? (void*)pingpong_kernel<2> : (void*)pingpong_kernel<3>;
    UMGEN_CUDA_OK(cudaLaunchCooperativeKernel(fn, dim3(2), dim3(32), args, 0, (cudaStream_t)stream_v));
    return 0;
}

extern "C" int umgen_debug_exchange_bench(void* buf_f, void* flags_u32, int iters, int krep, int variant, const void* stream_src,
                                          int stream_bytes, int nvals, void* out_ns_u64, void* stream_v) {
    using namespace umgen;
    const int flavour = variant >> 4;
    variant &= 15;
    void* fn = flavour == 0 ? (void*)exch_bench_kernel<0> : flavour == 1 ? (void*)exch_bench_kernel<1> : flavour == 2 ? (void*)exch_bench_kernel<2> : (void*)exch_bench_kernel<3>;
    int dev = 0, sms = 0;
    UMGEN_CUDA_OK(cudaGetDevice(&dev));
    UMGEN_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t smem = 200 * 1024;
    UMGEN_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void* args[] = {&buf_f, &flags_u32, &iters, &krep, &variant, &stream_src, &stream_bytes, &nvals, &out_ns_u64};
    UMGEN_CUDA_OK(cudaLaunchCooperativeKernel(fn, dim3(sms), dim3(XB_THREADS), args, smem, (cudaStream_t)stream_v));
    return 0;
}
