// Latency microbenchmarks of the instructions the decode kernel's serial path is made of (one warp, dependent chains):
// mma.sync.m16n8k16 (HMMA.16816.F32), shfl.sync, ld.shared, bar.sync over 12 warps, ex2, fp16 split.  Results in cycles per operation.
#include "../../umgen_b200/csrc/common.cuh"

namespace umgen {
__device__ __forceinline__ void mma16816(float (&d)[4], uint4 a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
__global__ void __launch_bounds__(416, 1) lat_bench_kernel(long long* out, float* sink) {
    __shared__ float sm[4096];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 4096; i += blockDim.x) sm[i] = (float)((i * 37) & 1023) * 4.0f;      // pointer-chase table: byte offsets
    __syncthreads();
    constexpr int N = 256;
    long long t0, t1;
    float acc = 0.f;
    if (warp == 0) {
        // (0) dependent HMMA chain
        float d[4] = {0.f, 0.f, 0.f, 0.f};
        uint4 a = make_uint4(0x3c003c00u + lane, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
        t0 = clock64();
#pragma unroll 16
        for (int i = 0; i < N; ++i) mma16816(d, a, 0x3c003c00u, 0x3c003c00u);
        t1 = clock64();
        acc += d[0] + d[1] + d[2] + d[3];
        if (lane == 0) out[0] = (t1 - t0) / N;
        // (1) 4 independent HMMA chains (throughput per MMA with ILP 4)
        float e[4][4] = {};
        t0 = clock64();
#pragma unroll 4
        for (int i = 0; i < N / 4; ++i) {
#pragma unroll
            for (int k = 0; k < 4; ++k) mma16816(e[k], a, 0x3c003c00u, 0x3c003c00u);
        }
        t1 = clock64();
        for (int k = 0; k < 4; ++k) acc += e[k][0] + e[k][3];
        if (lane == 0) out[1] = (t1 - t0) / N;
        // (2) dependent shfl chain
        float v = (float)lane;
        t0 = clock64();
#pragma unroll 16
        for (int i = 0; i < N; ++i) v = __shfl_xor_sync(0xffffffffu, v, 1) + 1.0f;
        t1 = clock64();
        acc += v;
        if (lane == 0) out[2] = (t1 - t0) / N;
        // (3) dependent ld.shared chain (pointer chase)
        int idx = lane;
        t0 = clock64();
#pragma unroll 16
        for (int i = 0; i < N; ++i) idx = (int)sm[idx & 4095] >> 2;
        t1 = clock64();
        acc += (float)idx;
        if (lane == 0) out[3] = (t1 - t0) / N;
        // (4) dependent FFMA chain
        float f = 1.0f + lane;
        t0 = clock64();
#pragma unroll 16
        for (int i = 0; i < N; ++i) f = fmaf(f, 1.0001f, 0.5f);
        t1 = clock64();
        acc += f;
        if (lane == 0) out[4] = (t1 - t0) / N;
        // (5) dependent ex2 chain
        float g = 0.5f;
        t0 = clock64();
#pragma unroll 16
        for (int i = 0; i < N; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(g));
        t1 = clock64();
        acc += g;
        if (lane == 0) out[5] = (t1 - t0) / N;
    }
    __syncthreads();
    // (6) bar.sync over 12 warps (named barrier 1, 384 threads), back to back
    if (warp < 12) {
        t0 = clock64();
        for (int i = 0; i < N; ++i) asm volatile("bar.sync 1, 384;" ::: "memory");
        t1 = clock64();
        if (tid == 0) out[6] = (t1 - t0) / N;
        // (7) bar.sync + one smem store/load round (the LayerNorm reduction pattern)
        t0 = clock64();
        float s = (float)tid;
        for (int i = 0; i < N; ++i) {
            if (lane == 0) sm[warp] = s;
            asm volatile("bar.sync 1, 384;" ::: "memory");
            float ts = 0.f;
#pragma unroll
            for (int w = 0; w < 12; ++w) ts += sm[w];
            s = ts * 0.01f;
            asm volatile("bar.sync 1, 384;" ::: "memory");
        }
        t1 = clock64();
        acc += s;
        if (tid == 0) out[7] = (t1 - t0) / N;
    }
    if (acc == 123456.789f) sink[0] = acc;
}
}  // namespace umgen

extern "C" int umgen_tools_lat_bench(void* out_i64, void* sink_f, void* stream) {
    umgen::lat_bench_kernel<<<1, 416, 0, (cudaStream_t)stream>>>((long long*)out_i64, (float*)sink_f);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
