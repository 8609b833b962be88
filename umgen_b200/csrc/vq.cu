// VQ pixel decoders (map_vae / image_var): the kernels around the tcgen05 GEMM that turn token grids into pixels.
// Reference: tokenizer/vq_model.py:87-101 (decode / decode_code / indices_to_quant), tokenizer/vq_modules.py:14-176,
// 293-415 (swish, GroupNorm32 eps 1e-6, Upsample, ResnetBlock, AttnBlock, Decoder), tools/decode_map.py:25-30 (to_rgb).
// Activations are channels-last fp16 [B, H, W, C]; 3x3 convolutions over >= 64 channels are implicit GEMMs (umgen_conv3x3_f16 in
// gemm_sm100.cu: the taps are fetched from the activation by 4-D TMA boxes), the two 16-channel ones go through the im2col matrix below, 1x1
// convolutions are plain GEMMs; fp32 accumulation throughout (the reference runs them under fp16 autocast), GroupNorm statistics are fp32.
#include "common.cuh"
#include "../../include/umgen.h"

namespace umgen {
extern int64_t g_launches;

// quant[b, h, w, :] = table[idx[b, h, w]]  (EmbeddingEMA.forward, quantize.py:341-342), 16 channels
__global__ void vq_gather_kernel(const int* __restrict__ idx, const float* __restrict__ table, __half* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 16) return;
    out[i] = __float2half_rn(__ldg(table + (size_t)idx[i >> 4] * 16 + (i & 15)));
}

// im2col for a 3x3 / stride 1 / pad 1 convolution over NHWC fp16, optionally fused with the nearest 2x upsample
// that precedes it (Upsample.forward, vq_modules.py:34-40).  Row (b, y, x) of A holds the taps in (ky, kx, c)
// order, zero padded to k_pad columns.  One thread moves 8 channels (16 bytes).
__global__ void im2col3x3_kernel(const __half* __restrict__ in, __half* __restrict__ A, int B, int H, int W, int Cin, int k_pad, int up) {
    const int chunks = k_pad / 8;
    const long long total = (long long)B * H * W * chunks;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int ch = (int)(gid % chunks);
    const long long row = gid / chunks;
    const int x = (int)(row % W), y = (int)((row / W) % H), b = (int)(row / ((long long)W * H));
    uint4 v = make_uint4(0, 0, 0, 0);
    const int col = ch * 8;
    if (col < 9 * Cin) {
        const int tap = col / Cin, c = col - tap * Cin;
        const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
            const int Hs = H >> up, Ws = W >> up;
            v = *reinterpret_cast<const uint4*>(in + (((size_t)b * Hs + (yy >> up)) * Ws + (xx >> up)) * Cin + c);
        }
    }
    *reinterpret_cast<uint4*>(A + (size_t)row * k_pad + col) = v;
}

// nearest 2x upsample (Upsample.forward, vq_modules.py:34-40: F.interpolate(scale_factor=2.0, mode="nearest")) over NHWC fp16; one thread
// writes 16 bytes.  The implicit-GEMM convolution that follows reads the upsampled activation (4 x the input bytes, against 36 x for its im2col).
__global__ void upsample2x_kernel(const __half* __restrict__ in, __half* __restrict__ out, int B, int H, int W, int C) {
    const int chunks = C / 8;
    const long long total = (long long)B * 2 * H * 2 * W * chunks;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int ch = (int)(gid % chunks);
    const long long px = gid / chunks;
    const int x = (int)(px % (2 * W)), y = (int)((px / (2 * W)) % (2 * H)), b = (int)(px / ((long long)4 * W * H));
    reinterpret_cast<uint4*>(out)[gid] = __ldg(reinterpret_cast<const uint4*>(in + (((size_t)b * H + (y >> 1)) * W + (x >> 1)) * C) + ch);
}

// GroupNorm(32 groups, eps 1e-6, affine) statistics over NHWC fp16.
// Slab version (default): a CTA reads a slab of pixels (gn_slab_px) of one image with coalesced 16-byte loads (thread t always holds the same 8 channels), reduces
// its slab to 32 x (sum, sum of squares) in a fixed order and writes them to part[b][slab][g]; the last CTA of an image to finish (ticket counter)
// adds the slabs up in slab order and writes mean / rstd, so the statistics do not depend on which CTA ran when.
static inline int gn_slab_px(int64_t HW) { return HW >= 32768 ? 1024 : HW >= 4096 ? 256 : 64; }      // 8 .. 128 slabs per image for the decoders' 512 .. 131072 pixels
__global__ void __launch_bounds__(256) gn_stats_slab_kernel(const __half* __restrict__ x, float* __restrict__ stats, float* __restrict__ part,
                                                            unsigned int* __restrict__ ticket, int HW, int C, int slab_px) {
    const int slab = blockIdx.x, n_slabs = gridDim.x, b = blockIdx.y;
    const int tpp = C / 8;                                  // threads per pixel: 16, 32 or 64
    const int j = threadIdx.x % tpp, prow = threadIdx.x / tpp, pstep = 256 / tpp;
    const int p0 = slab * slab_px, p1 = min(HW, p0 + slab_px);
    const __half* base = x + ((size_t)b * HW) * C + 8 * j;
    float sa = 0.f, qa = 0.f, sb = 0.f, qb = 0.f;           // channels 8j..8j+3 and 8j+4..8j+7 (a group is 4, 8 or 16 channels wide)
    auto add = [&](const uint4& v) {
        const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&v.x)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
        const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&v.z)), f3 = __half22float2(*reinterpret_cast<const __half2*>(&v.w));
        sa += (f0.x + f0.y) + (f1.x + f1.y);
        qa = fmaf(f0.x, f0.x, fmaf(f0.y, f0.y, fmaf(f1.x, f1.x, fmaf(f1.y, f1.y, qa))));
        sb += (f2.x + f2.y) + (f3.x + f3.y);
        qb = fmaf(f2.x, f2.x, fmaf(f2.y, f2.y, fmaf(f3.x, f3.x, fmaf(f3.y, f3.y, qb))));
    };
    int px = p0 + prow;
    for (; px + 3 * pstep < p1; px += 4 * pstep) {          // four 16-byte loads in flight per thread: the pass is a pure stream
        const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(base + (size_t)px * C));
        const uint4 v1 = __ldg(reinterpret_cast<const uint4*>(base + (size_t)(px + pstep) * C));
        const uint4 v2 = __ldg(reinterpret_cast<const uint4*>(base + (size_t)(px + 2 * pstep) * C));
        const uint4 v3 = __ldg(reinterpret_cast<const uint4*>(base + (size_t)(px + 3 * pstep) * C));
        add(v0); add(v1); add(v2); add(v3);
    }
    for (; px < p1; px += pstep) add(__ldg(reinterpret_cast<const uint4*>(base + (size_t)px * C)));
    __shared__ float4 acc[256];
    __shared__ float2 quad[128];                            // per 4-channel run: C / 4 <= 128 of them
    __shared__ bool last;
    acc[threadIdx.x] = make_float4(sa, qa, sb, qb);
    __syncthreads();
    const int n_quads = C / 4;
    if (threadIdx.x < n_quads) {
        const int jj = threadIdx.x >> 1, hi = threadIdx.x & 1;
        float s = 0.f, q = 0.f;
        for (int r = 0; r < pstep; ++r) {
            const float4 a = acc[r * tpp + jj];
            s += hi ? a.z : a.x;
            q += hi ? a.w : a.y;
        }
        quad[threadIdx.x] = make_float2(s, q);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int qpg = n_quads / 32;                       // 1, 2 or 4 runs per group
        float s = 0.f, q = 0.f;
        for (int r = 0; r < qpg; ++r) { s += quad[threadIdx.x * qpg + r].x; q += quad[threadIdx.x * qpg + r].y; }
        float* dst = part + (((size_t)b * n_slabs + slab) * 32 + threadIdx.x) * 2;
        __stcg(dst, s);
        __stcg(dst + 1, q);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(ticket + b, 1u) == (unsigned)n_slabs - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (threadIdx.x < 32) {
        float s = 0.f, q = 0.f;
        const float* src = part + ((size_t)b * n_slabs * 32 + threadIdx.x) * 2;
        for (int k = 0; k < n_slabs; ++k) { s += __ldcg(src + (size_t)k * 64); q += __ldcg(src + (size_t)k * 64 + 1); }
        const float n = (float)HW * (float)(C / 32), mean = s / n;
        stats[(b * 32 + threadIdx.x) * 2] = mean;
        stats[(b * 32 + threadIdx.x) * 2 + 1] = rsqrtf(fmaxf(q / n - mean * mean, 0.f) + 1e-6f);
    }
    if (threadIdx.x == 0) ticket[b] = 0u;                   // ready for the next call on this stream
}

// one CTA per (batch, group): strided reads, kept for C not a multiple of 128 and as the cross-check of the slab kernel
__global__ void __launch_bounds__(256) gn_stats_kernel(const __half* __restrict__ x, float* __restrict__ stats, int HW, int C) {
    const int b = blockIdx.x / 32, g = blockIdx.x % 32, cpg = C / 32;
    const __half* base = x + (size_t)b * HW * C + g * cpg;
    float s = 0.f, q = 0.f;
    const int per_px = cpg / 2;                       // half2 loads (cpg is 4, 8 or 16)
    for (int i = threadIdx.x; i < HW * per_px; i += 256) {
        const int px = i / per_px, j = i - px * per_px;
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(base + (size_t)px * C + 2 * j));
        s += f.x + f.y;
        q = fmaf(f.x, f.x, fmaf(f.y, f.y, q));
    }
    __shared__ float rs[8], rq[8];
    s = warp_sum(s); q = warp_sum(q);
    if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = s; rq[threadIdx.x >> 5] = q; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float ts = 0.f, tq = 0.f;
        for (int i = 0; i < 8; ++i) { ts += rs[i]; tq += rq[i]; }
        const float n = (float)HW * cpg, mean = ts / n;
        stats[blockIdx.x * 2] = mean;
        stats[blockIdx.x * 2 + 1] = rsqrtf(fmaxf(tq / n - mean * mean, 0.f) + 1e-6f);
    }
}
// y = (x - mean) * rstd * gamma + beta, optionally followed by swish x * sigmoid(x) (vq_modules.py:14-16); one thread = 8 channels (16 bytes), which
// lie in one group (C >= 256) or two (C = 128: 4 channels per group)
__global__ void __launch_bounds__(256) gn_apply_kernel(const __half* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, __half* __restrict__ y, long long n8, int HW, int C, int swish) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // 16-byte index
    if (i >= n8) return;
    const int c = (int)((i * 8) % C);
    const int b = (int)((i * 8) / ((long long)HW * C));
    const int cpg = C / 32;
    const float* st0 = stats + (b * 32 + c / cpg) * 2;
    const float* st1 = stats + (b * 32 + (c + 4) / cpg) * 2;
    const float m0 = st0[0], r0 = st0[1], m1 = st1[0], r1 = st1[1];
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(x) + i);
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
    float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&v.x)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
    float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&v.z)), f3 = __half22float2(*reinterpret_cast<const __half2*>(&v.w));
    f0.x = (f0.x - m0) * r0 * g0.x + b0.x; f0.y = (f0.y - m0) * r0 * g0.y + b0.y;
    f1.x = (f1.x - m0) * r0 * g0.z + b0.z; f1.y = (f1.y - m0) * r0 * g0.w + b0.w;
    f2.x = (f2.x - m1) * r1 * g1.x + b1.x; f2.y = (f2.y - m1) * r1 * g1.y + b1.y;
    f3.x = (f3.x - m1) * r1 * g1.z + b1.z; f3.y = (f3.y - m1) * r1 * g1.w + b1.w;
    if (swish) {
        f0.x = f0.x / (1.0f + __expf(-f0.x)); f0.y = f0.y / (1.0f + __expf(-f0.y)); f1.x = f1.x / (1.0f + __expf(-f1.x)); f1.y = f1.y / (1.0f + __expf(-f1.y));
        f2.x = f2.x / (1.0f + __expf(-f2.x)); f2.y = f2.y / (1.0f + __expf(-f2.y)); f3.x = f3.x / (1.0f + __expf(-f3.x)); f3.y = f3.y / (1.0f + __expf(-f3.y));
    }
    const __half2 h0 = __floats2half2_rn(f0.x, f0.y), h1 = __floats2half2_rn(f1.x, f1.y), h2 = __floats2half2_rn(f2.x, f2.y), h3 = __floats2half2_rn(f3.x, f3.y);
    reinterpret_cast<uint4*>(y)[i] = make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1),
                                                *reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
}

// softmax over rows of fp32 scores (already scaled) -> fp16 probabilities (AttnBlock, vq_modules.py:161-163)
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s, __half* __restrict__ p, int n, float scale) {
    const float* row = s + (size_t)blockIdx.x * n;
    __shared__ float red[8];
    float m = -INFINITY;
    for (int i = threadIdx.x; i < n; i += 256) m = fmaxf(m, row[i] * scale);
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    m = red[0];
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    __syncthreads();
    float l = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) l += __expf(row[i] * scale - m);
    l = warp_sum(l);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = l;
    __syncthreads();
    l = 0.f;
    for (int i = 0; i < 8; ++i) l += red[i];
    const float inv = 1.0f / l;
    for (int i = threadIdx.x; i < n; i += 256) p[(size_t)blockIdx.x * n + i] = __float2half_rn(__expf(row[i] * scale - m) * inv);
}

// out[c][r] = in[r][c]  (fp16), 32x32 tiles
__global__ void transpose_h_kernel(const __half* __restrict__ in, __half* __restrict__ out, int rows, int cols) {
    __shared__ __half tile[32][33];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int r = r0 + j, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[j][threadIdx.x] = in[(size_t)r * cols + c];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int c = c0 + j, r = r0 + threadIdx.x;
        if (r < rows && c < cols) out[(size_t)c * rows + r] = tile[threadIdx.x][j];
    }
}

// final 3x3 convolution to a handful of channels (conv_out, vq_modules.py:384-387): NHWC fp16 in, NCHW fp32 out.
// weights fp32 [Cout][9][Cin] in (ky, kx, c) order.  One warp per output pixel.
__global__ void __launch_bounds__(256) conv_out_kernel(const __half* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                                                       float* __restrict__ out, int B, int H, int W, int Cin, int Cout) {
    const long long px = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (px >= (long long)B * H * W) return;
    const int x = (int)(px % W), y = (int)((px / W) % H), b = (int)(px / ((long long)W * H));
    float acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = 0.f;
    for (int tap = 0; tap < 9; ++tap) {
        const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
        const __half* src = in + (((size_t)b * H + yy) * W + xx) * Cin;
        for (int c = lane * 2; c < Cin; c += 64) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(src + c));
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                if (o < Cout) {
                    const float* wp = w + ((size_t)o * 9 + tap) * Cin + c;
                    acc[o] = fmaf(f.x, __ldg(wp), fmaf(f.y, __ldg(wp + 1), acc[o]));
                }
            }
        }
    }
#pragma unroll
    for (int o = 0; o < 8; ++o) {
        if (o < Cout) {
            const float v = warp_sum(acc[o]);
            if (lane == 0) out[(((size_t)b * Cout + o) * H + y) * W + x] = v + __ldg(bias + o);
        }
    }
}

// to_rgb (tools/decode_map.py:25-30): 1x1 projection Cin -> 3 with fixed weights, then min-max to [-1, 1] over the chunk
__device__ __forceinline__ unsigned int f2ord(float f) { unsigned int u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(unsigned int u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }
__global__ void rgb_project_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ out, unsigned int* __restrict__ mm,
                                   int B, int Cin, int HW) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float lo = INFINITY, hi = -INFINITY;
    if (i < (long long)B * HW) {
        const int b = (int)(i / HW), p = (int)(i % HW);
        for (int o = 0; o < 3; ++o) {
            float s = 0.f;
            for (int c = 0; c < Cin; ++c) s = fmaf(__ldg(w + o * Cin + c), x[((size_t)b * Cin + c) * HW + p], s);
            out[((size_t)b * 3 + o) * HW + p] = s;
            lo = fminf(lo, s); hi = fmaxf(hi, s);
        }
    }
    lo = -warp_max(-lo); hi = warp_max(hi);
    if ((threadIdx.x & 31) == 0 && hi >= lo) { atomicMin(mm, f2ord(lo)); atomicMax(mm + 1, f2ord(hi)); }
}
__global__ void rgb_normalize_kernel(float* __restrict__ x, const unsigned int* __restrict__ mm, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float lo = ord2f(mm[0]), hi = ord2f(mm[1]);
    x[i] = 2.0f * (x[i] - lo) / (hi - lo) - 1.0f;
}
__global__ void rgb_init_kernel(unsigned int* mm) { mm[0] = 0xffffffffu; mm[1] = 0u; }

}  // namespace umgen

using namespace umgen;
#define ST(s) ((cudaStream_t)(s))
#define LAUNCH_OK() do { UMGEN_CUDA_OK(cudaGetLastError()); g_launches += 1; } while (0)

extern "C" int umgen_vq_gather(const void* idx_i32, const void* table_f, void* out_h, int64_t n, void* stream) {
    vq_gather_kernel<<<(unsigned)((n * 16 + 255) / 256), 256, 0, ST(stream)>>>((const int*)idx_i32, (const float*)table_f, (__half*)out_h, (int)n);
    LAUNCH_OK();
    return 0;
}
extern "C" int umgen_im2col3x3(const void* in_h, void* a_h, int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t k_pad, int upsample, void* stream) {
    if (Cin % 8 != 0 || k_pad % 64 != 0 || k_pad < 9 * Cin || (upsample && ((H | W) & 1))) { set_error("im2col: bad shape"); return -1; }
    const long long total = B * H * W * (k_pad / 8);
    im2col3x3_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ST(stream)>>>((const __half*)in_h, (__half*)a_h, (int)B, (int)H, (int)W, (int)Cin,
                                                                             (int)k_pad, upsample ? 1 : 0);
    LAUNCH_OK();
    return 0;
}
extern "C" int umgen_groupnorm_nhwc(const void* x_h, const void* gamma_f, const void* beta_f, void* y_h, void* stats_f, int64_t B, int64_t HW,
                                    int64_t C, int swish, void* stream) {
    if (C % 128 != 0) { set_error("groupnorm: C must be a multiple of 128 (groups of >= 4 channels)"); return -1; }
    gn_stats_kernel<<<(unsigned)(B * 32), 256, 0, ST(stream)>>>((const __half*)x_h, (float*)stats_f, (int)HW, (int)C);
    LAUNCH_OK();
    const long long n8 = B * HW * C / 8;
    gn_apply_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, ST(stream)>>>((const __half*)x_h, (const float*)stats_f, (const float*)gamma_f,
                                                                         (const float*)beta_f, (__half*)y_h, n8, (int)HW, (int)C, swish);
    LAUNCH_OK();
    return 0;
}
extern "C" int64_t umgen_groupnorm_scratch_floats(int64_t B, int64_t HW) {
    const int64_t n_slabs = (HW + gn_slab_px(HW) - 1) / gn_slab_px(HW);
    return B * 64 + B * n_slabs * 64 + B;      // mean / rstd, per-slab partial sums, one ticket word per image
}
// same result layout as umgen_groupnorm_nhwc; scratch_f holds umgen_groupnorm_scratch_floats(B, HW) floats whose LAST B words (the tickets) the caller
// zeroes once (the kernel leaves them zero).  Coalesced statistics pass: see gn_stats_slab_kernel.
extern "C" int umgen_groupnorm_nhwc_slab(const void* x_h, const void* gamma_f, const void* beta_f, void* y_h, void* scratch_f, int64_t B, int64_t HW,
                                         int64_t C, int swish, void* stream) {
    if (C != 128 && C != 256 && C != 512) { set_error("groupnorm (slab): C must be 128, 256 or 512"); return -1; }
    if (B < 1 || B > 65535 || HW < 1) { set_error("groupnorm (slab): bad shape"); return -1; }
    const int slab_px = gn_slab_px(HW);
    const int64_t n_slabs = (HW + slab_px - 1) / slab_px;
    float* stats = (float*)scratch_f;
    float* part = stats + B * 64;
    unsigned int* ticket = (unsigned int*)(part + B * n_slabs * 64);
    gn_stats_slab_kernel<<<dim3((unsigned)n_slabs, (unsigned)B), 256, 0, ST(stream)>>>((const __half*)x_h, stats, part, ticket, (int)HW, (int)C, slab_px);
    LAUNCH_OK();
    const long long n8 = B * HW * C / 8;
    gn_apply_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, ST(stream)>>>((const __half*)x_h, (const float*)stats, (const float*)gamma_f,
                                                                         (const float*)beta_f, (__half*)y_h, n8, (int)HW, (int)C, swish);
    LAUNCH_OK();
    return 0;
}
extern "C" int umgen_upsample2x_nhwc(const void* in_h, void* out_h, int64_t B, int64_t H, int64_t W, int64_t C, void* stream) {
    if (C % 8 != 0 || B < 1 || H < 1 || W < 1) { set_error("upsample2x: C must be a multiple of 8"); return -1; }
    const long long total = B * 4 * H * W * (C / 8);
    upsample2x_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ST(stream)>>>((const __half*)in_h, (__half*)out_h, (int)B, (int)H, (int)W, (int)C);
    LAUNCH_OK();
    return 0;
}
extern "C" int umgen_softmax_rows(const void* s_f, void* p_h, int64_t rows, int64_t n, double scale, void* stream) {
    softmax_rows_kernel<<<(unsigned)rows, 256, 0, ST(stream)>>>((const float*)s_f, (__half*)p_h, (int)n, (float)scale);
    LAUNCH_OK();
    return 0;
}
extern "C" int umgen_transpose_f16(const void* in_h, void* out_h, int64_t rows, int64_t cols, void* stream) {
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
    transpose_h_kernel<<<grid, dim3(32, 8), 0, ST(stream)>>>((const __half*)in_h, (__half*)out_h, (int)rows, (int)cols);
    LAUNCH_OK();
    return 0;
}
extern "C" int umgen_conv_out3x3(const void* in_h, const void* w_f, const void* bias_f, void* out_f, int64_t B, int64_t H, int64_t W,
                                 int64_t Cin, int64_t Cout, void* stream) {
    if (Cout > 8 || Cin % 2 != 0) { set_error("conv_out: Cout <= 8"); return -1; }
    const long long px = B * H * W;
    conv_out_kernel<<<(unsigned)((px + 7) / 8), 256, 0, ST(stream)>>>((const __half*)in_h, (const float*)w_f, (const float*)bias_f, (float*)out_f,
                                                                     (int)B, (int)H, (int)W, (int)Cin, (int)Cout);
    LAUNCH_OK();
    return 0;
}
extern "C" int umgen_to_rgb(const void* x_f, const void* w_f, void* out_f, void* minmax_u32, int64_t B, int64_t Cin, int64_t HW, void* stream) {
    rgb_init_kernel<<<1, 1, 0, ST(stream)>>>((unsigned int*)minmax_u32);
    LAUNCH_OK();
    rgb_project_kernel<<<(unsigned)((B * HW + 255) / 256), 256, 0, ST(stream)>>>((const float*)x_f, (const float*)w_f, (float*)out_f,
                                                                                (unsigned int*)minmax_u32, (int)B, (int)Cin, (int)HW);
    LAUNCH_OK();
    const long long n = B * 3 * HW;
    rgb_normalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>((float*)out_f, (const unsigned int*)minmax_u32, n);
    LAUNCH_OK();
    return 0;
}

// Lazy module loading (the CUDA 12 default) loads a kernel on its first launch and that load waits for an idle device -- which never comes while the
// persistent decode kernel spins on a flag.  umgen_preload() (capi.cu) forces every kernel of the library to load up front.
#define UMGEN_PRELOAD(k) UMGEN_CUDA_OK(cudaFuncGetAttributes(&fa_, k))
namespace umgen {
int preload_vq() {
    cudaFuncAttributes fa_;
    UMGEN_PRELOAD(vq_gather_kernel); UMGEN_PRELOAD(im2col3x3_kernel); UMGEN_PRELOAD(gn_stats_kernel); UMGEN_PRELOAD(gn_apply_kernel);
    UMGEN_PRELOAD(softmax_rows_kernel); UMGEN_PRELOAD(transpose_h_kernel); UMGEN_PRELOAD(conv_out_kernel); UMGEN_PRELOAD(rgb_project_kernel);
    UMGEN_PRELOAD(rgb_normalize_kernel); UMGEN_PRELOAD(rgb_init_kernel); UMGEN_PRELOAD(upsample2x_kernel); UMGEN_PRELOAD(gn_stats_slab_kernel);
    return 0;
}
}  // namespace umgen
