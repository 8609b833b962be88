// Debug microbenchmark (not part of the product path): how fast can ONE SM pull a contiguous stream HBM -> shared memory with 1-D bulk
// copies (cp.async.bulk + mbarrier complete_tx) when only a few SMs of the chip are active?  This is the number that bounds a decode
// partitioning that keeps all exchanges inside one thread-block cluster (DSMEM only, no L2 hops): 510 MB of weights per step divided
// by (cluster size x per-SM rate).
//   grid = n_clusters x cluster_size CTAs, each CTA streams `bytes_per_cta` from its own region in chunks of `chunk` bytes through a
//   ring of `nst` stages; one producer thread issues, one consumer warp waits on the full barrier, reads 16 bytes per lane of the stage
//   (so the data is really consumed) and releases it.
// out[0] = cycles of the slowest CTA, out[1] = cycles of CTA 0, out[2] = max active clusters reported by the occupancy API
#include "../../umgen_b200/csrc/common.cuh"
#include "../../include/umgen.h"

namespace umgen {
namespace {
constexpr int SB_THREADS = 64;
constexpr int SB_MAX_STAGES = 32;

__global__ void __launch_bounds__(SB_THREADS, 1) stream_bench_kernel(const uint8_t* __restrict__ src, long long bytes_per_cta, int chunk, int nst,
                                                                    long long* out, uint32_t* sink, int pf_dist) {
    extern __shared__ __align__(128) uint8_t sb_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(sb_raw);
    uint64_t* empty = full + SB_MAX_STAGES;
    uint8_t* ring = sb_raw + 1024;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < nst; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_fence_init();
    }
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    const uint8_t* mine = src + (size_t)blockIdx.x * (size_t)bytes_per_cta;
    const int n = (int)(bytes_per_cta / chunk);
    const long long t0 = clock64();
    uint32_t acc = 0;
    if (tid == 0) {
        for (int k = 0; k < n; ++k) {
            const int s = k % nst;
            if (k >= nst) { while (!mbar_try_wait(&empty[s], ((k / nst) - 1) & 1)) {} }
            if (pf_dist > 0 && k + pf_dist < n)      // warm L2 pf_dist chunks ahead: the bulk copy then pays L2 latency, not HBM latency
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(mine + (size_t)(k + pf_dist) * chunk), "r"(chunk) : "memory");
            mbar_arrive_expect_tx(&full[s], (uint32_t)chunk);
            bulk_g2s(ring + (size_t)s * chunk, mine + (size_t)k * chunk, (uint32_t)chunk, &full[s]);
        }
    } else if (tid >= 32) {
        const int lane = tid - 32;
        for (int k = 0; k < n; ++k) {
            const int s = k % nst;
            while (!mbar_try_wait(&full[s], (k / nst) & 1)) {}
            acc += *reinterpret_cast<const uint32_t*>(ring + (size_t)s * chunk + lane * 16);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
    }
    const long long t1 = clock64();
    if (tid == 32) {
        atomicMax((unsigned long long*)&out[0], (unsigned long long)(t1 - t0));
        if (blockIdx.x == 0) out[1] = t1 - t0;
    }
    if (acc == 0x12345678u) sink[0] = acc;
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
}  // namespace
}  // namespace umgen

extern "C" int umgen_debug_stream_bench(const void* src, int64_t bytes_per_cta, int chunk, int nst, int cluster_size, int n_clusters,
                                        void* out_i64, void* sink_u32, int pf_dist, void* stream_v) {
    using namespace umgen;
    if (nst < 1 || nst > SB_MAX_STAGES || chunk % 16 != 0 || bytes_per_cta % chunk != 0) { set_error("stream bench: bad arguments"); return -1; }
    const size_t smem = 1024 + (size_t)nst * chunk;
    if (smem > 227 * 1024) { set_error("stream bench: ring too large"); return -1; }
    UMGEN_CUDA_OK(cudaFuncSetAttribute(stream_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (cluster_size > 8) UMGEN_CUDA_OK(cudaFuncSetAttribute(stream_bench_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(cluster_size * n_clusters);
    cfg.blockDim = dim3(SB_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream_v;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster_size; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int ncl = 0;
    UMGEN_CUDA_OK(cudaOccupancyMaxActiveClusters(&ncl, (const void*)stream_bench_kernel, &cfg));
    long long* out = (long long*)out_i64;
    UMGEN_CUDA_OK(cudaMemsetAsync(out, 0, 4 * sizeof(long long), cfg.stream));
    long long ncl_ll = ncl;
    UMGEN_CUDA_OK(cudaMemcpyAsync(out + 2, &ncl_ll, sizeof(long long), cudaMemcpyHostToDevice, cfg.stream));
    if (ncl < n_clusters) { set_error("stream bench: only %d clusters of %d CTAs fit at once", ncl, cluster_size); return -3; }
    const uint8_t* s = (const uint8_t*)src;
    long long bpc = bytes_per_cta;
    uint32_t* sink = (uint32_t*)sink_u32;
    void* args[] = {&s, &bpc, &chunk, &nst, &out, &sink, &pf_dist};
    UMGEN_CUDA_OK(cudaLaunchKernelExC(&cfg, (const void*)stream_bench_kernel, args));
    return 0;
}
