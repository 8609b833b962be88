// TAR-side kernels (everything around the GEMMs): LayerNorm, input embedding, action-aware map warp,
// temporal / ego small attention, spatial flash attention, ego cross attention, row sampler, and the
// assembly of the OAR conditioning feature.  Reference: models/UMGen.py:310-354, 411-515, 634-872,
// 994-1024; models/module.py:26-37, 179-230, 296-375, 454-509, 630-706.
#include <cuda_bf16.h>

#include "decode_shared.cuh"      // block_topp_sample (the nucleus sampler of the decode kernels) for the ego head

namespace umgen {
extern int64_t g_launches;

// ------------------------------------------------------------------------------------------------
// LayerNorm over rows of 768 (weight only, eps 1e-5; module.py:26-37).  One warp per row.
// ------------------------------------------------------------------------------------------------
template <bool OUT_HALF>
__global__ void __launch_bounds__(256) ln_rows_kernel(const float* __restrict__ x, const float* __restrict__ w, void* __restrict__ out, int rows) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * C);
    float4 v[6];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) { v[i] = xr[lane + 32 * i]; s += v[i].x + v[i].y + v[i].z + v[i].w; }
    const float mean = warp_sum(s) * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
        q += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + 1e-5f);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(w) + lane + 32 * i);
        const float a = v[i].x * rstd * g.x, b = v[i].y * rstd * g.y, c2 = v[i].z * rstd * g.z, d = v[i].w * rstd * g.w;
        if (OUT_HALF) {
            __half2 h0 = __floats2half2_rn(a, b), h1 = __floats2half2_rn(c2, d);
            uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
            reinterpret_cast<uint2*>((__half*)out + (size_t)row * C)[lane + 32 * i] = pk;
        } else {
            reinterpret_cast<float4*>((float*)out + (size_t)row * C)[lane + 32 * i] = make_float4(a, b, c2, d);
        }
    }
}

// fp32 -> fp16 row cast (attention outputs that are already fp16 never need it; used for ego queries)
__global__ void cast_f16_kernel(const float* __restrict__ x, __half* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __float2half_rn(x[i]);
}

// ------------------------------------------------------------------------------------------------
// Map feature of every conditioning frame: emb_table[token] (+ grid-centre sinusoid)   (UMGen.py:448-458)
// ------------------------------------------------------------------------------------------------
__global__ void map_feature_kernel(const int* __restrict__ tok, const float* __restrict__ table, const float* __restrict__ grid_pos,
                                   float* __restrict__ out, int n_tok) {
    const int i = blockIdx.x;          // token index t*1024 + cell
    if (i >= n_tok) return;
    const float4* src = reinterpret_cast<const float4*>(table + (size_t)tok[i] * C);
    const float4* gp = grid_pos ? reinterpret_cast<const float4*>(grid_pos + (size_t)(i & 1023) * C) : nullptr;
    float4* dst = reinterpret_cast<float4*>(out + (size_t)i * C);
    for (int c4 = threadIdx.x; c4 < C / 4; c4 += blockDim.x) {
        float4 v = __ldg(src + c4);
        if (gp) { float4 g = __ldg(gp + c4); v.x += g.x; v.y += g.y; v.z += g.z; v.w += g.w; }
        dst[c4] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// Action-aware map alignment: affine_grid + grid_sample(bilinear, zeros, align_corners=False) of the
// 32x32 map feature by the frame's decoded ego action (UMGen.py:310-354).
// ------------------------------------------------------------------------------------------------
__global__ void map_warp_kernel(const float* __restrict__ feat, const int* __restrict__ pose_tok, const float* __restrict__ pose_lut,
                                float* __restrict__ out, int T) {
    const int t = blockIdx.x >> 10, cell = blockIdx.x & 1023;
    if (t >= T) return;
    const int h = cell >> 5, w = cell & 31;
    const float px = __ldg(pose_lut + __ldg(pose_tok + t * 3 + 0) * 3 + 0);
    const float py = __ldg(pose_lut + __ldg(pose_tok + t * 3 + 1) * 3 + 1);
    const float th = __ldg(pose_lut + __ldg(pose_tok + t * 3 + 2) * 3 + 2);
    const float dx = 2.0f * (px / 4.0f) / 32.0f, dy = 2.0f * (py / 4.0f) / 32.0f;
    const float cs = cosf(-th), sn = sinf(-th);
    const float xn = (2.0f * w + 1.0f) / 32.0f - 1.0f, yn = (2.0f * h + 1.0f) / 32.0f - 1.0f;
    const float gx = cs * xn + (-sn) * yn + (-dy);
    const float gy = sn * xn + cs * yn + (-dx);
    const float ix = ((gx + 1.0f) * 32.0f - 1.0f) * 0.5f, iy = ((gy + 1.0f) * 32.0f - 1.0f) * 0.5f;
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
    const float wx1 = ix - fx, wx0 = 1.0f - wx1, wy1 = iy - fy, wy0 = 1.0f - wy1;
    const bool vx0 = x0 >= 0 && x0 < 32, vx1 = x1 >= 0 && x1 < 32, vy0 = y0 >= 0 && y0 < 32, vy1 = y1 >= 0 && y1 < 32;
    const float* base = feat + (size_t)t * 1024 * C;
    const float4* p00 = reinterpret_cast<const float4*>(base + (size_t)(y0 * 32 + x0) * C);
    const float4* p01 = reinterpret_cast<const float4*>(base + (size_t)(y0 * 32 + x1) * C);
    const float4* p10 = reinterpret_cast<const float4*>(base + (size_t)(y1 * 32 + x0) * C);
    const float4* p11 = reinterpret_cast<const float4*>(base + (size_t)(y1 * 32 + x1) * C);
    float4* dst = reinterpret_cast<float4*>(out + ((size_t)t * 1024 + cell) * C);
    const float w00 = wy0 * wx0, w01 = wy0 * wx1, w10 = wy1 * wx0, w11 = wy1 * wx1;
    for (int c4 = threadIdx.x; c4 < C / 4; c4 += blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (vy0 && vx0) { float4 v = p00[c4]; acc.x += w00 * v.x; acc.y += w00 * v.y; acc.z += w00 * v.z; acc.w += w00 * v.w; }
        if (vy0 && vx1) { float4 v = p01[c4]; acc.x += w01 * v.x; acc.y += w01 * v.y; acc.z += w01 * v.z; acc.w += w01 * v.w; }
        if (vy1 && vx0) { float4 v = p10[c4]; acc.x += w10 * v.x; acc.y += w10 * v.y; acc.z += w10 * v.z; acc.w += w10 * v.w; }
        if (vy1 && vx1) { float4 v = p11[c4]; acc.x += w11 * v.x; acc.y += w11 * v.y; acc.z += w11 * v.z; acc.w += w11 * v.w; }
        dst[c4] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// Input sequence of one TAR pass: per-modality embedding + bos/eos + spe + tpe   (UMGen.py:438-515)
// ------------------------------------------------------------------------------------------------
struct EmbedArgs {
    const int *pose, *map, *bbox, *image;     // [T,3] [T,1024] [T,660] [T,512]
    const float *fpe, *img_table, *be, *axe, *spe, *tpe, *sp;   // sp = bbox3d_spatial_posi [1030,768]
    const float *map_feat, *map_warped;       // [T,1024,768]; warped may be null
    float* out;                               // [T, S, 768]
    int T, S, n_mods;                         // n_mods: 2 (pose,map) 3 (+bbox3d) 4 (+image)
    int t_offset;                             // index in the window of the first frame given (temporal position embedding)
};
__global__ void embed_sequence_kernel(const EmbedArgs a) {
    const int t = blockIdx.x / a.S, pos = blockIdx.x - t * a.S;
    const float* src = nullptr;
    const float* add1 = nullptr;     // second additive term (warped map / bbox sinusoid handled separately)
    int bx = -1, by = -1;
    if (pos < 5) {
        src = (pos == 0) ? a.axe : (pos == 4) ? a.axe + C : a.fpe + (size_t)a.pose[t * 3 + pos - 1] * C;
    } else if (pos < 1031) {
        const int i = pos - 5;
        if (i == 0) src = a.axe + 2 * C;
        else if (i == 1025) src = a.axe + 3 * C;
        else {
            src = a.map_feat + ((size_t)t * 1024 + (i - 1)) * C;
            if (a.map_warped) add1 = a.map_warped + ((size_t)t * 1024 + (i - 1)) * C;
        }
    } else if (pos < 1693) {
        const int i = pos - 1031;
        if (i == 0) src = a.axe + 4 * C;
        else if (i == 661) src = a.axe + 5 * C;
        else {
            const int* fr = a.bbox + (size_t)t * 660;
            src = a.be + (size_t)fr[i - 1] * C;
            const int slot = (i - 1) / 11;
            bx = fr[slot * 11];
            by = fr[slot * 11 + 1];
        }
    } else {
        const int i = pos - 1693;
        if (i == 0) src = a.axe + 6 * C;
        else if (i == 513) src = a.axe + 7 * C;
        else src = a.img_table + (size_t)a.image[t * 512 + i - 1] * C;
    }
    const float* spe = a.spe + (size_t)pos * C;
    const float* tpe = a.tpe + (size_t)(t + a.t_offset) * C;
    float* dst = a.out + ((size_t)t * a.S + pos) * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float v = __ldg(src + c);
        if (add1) v = add1[c] + v;
        if (bx >= 0) {     // bf16(sp[x] + sp[y]) : add_spatial_pos_emb works in bfloat16 (UMGen.py:418-423)
            const float s2 = __ldg(a.sp + (size_t)bx * C + c) + __ldg(a.sp + (size_t)by * C + c);
            v += __bfloat162float(__float2bfloat16_rn(s2));
        }
        dst[c] = v + __ldg(spe + c) + __ldg(tpe + c);
    }
}

// ------------------------------------------------------------------------------------------------
// Small attention (<= 32 keys) for the causal temporal attention over frames of one sequence position
// (module.py:342-345 with causal=True) and the 3-token ego self attention (module.py:669, causal=False).
// qkv fp16 [rows][2304] (q | k | v); token i of group g lives at row g * group_stride + i * tok_stride.
// One warp per (group, head); lane i owns query i.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) small_attn_kernel(const __half* __restrict__ qkv, __half* __restrict__ y, int n_groups, int n_tok,
                                                         long long group_stride, long long tok_stride, int causal, int q0) {
    __shared__ __half kv[4][2][32][HD];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gw = blockIdx.x * 4 + wid;
    if (gw >= n_groups * NH) return;
    const int g = gw / NH, h = gw - g * NH;
    const float scale = 0.14433756729740643f;
    // stage k and v rows: n_tok rows x 96 B each = 6 x 16 B
    for (int i = lane; i < n_tok * 12; i += 32) {
        const int tkn = i / 12, part = i - tkn * 12, which = part / 6, piece = part - which * 6;
        const __half* src = qkv + (size_t)(g * group_stride + tkn * tok_stride) * (3 * C) + (1 + which) * C + h * HD + piece * 8;
        *reinterpret_cast<uint4*>(&kv[wid][which][tkn][piece * 8]) = *reinterpret_cast<const uint4*>(src);
    }
    __syncwarp();
    if (lane >= q0 && lane < n_tok) {      // q0 > 0: only the last queries are wanted (the other rows of qkv serve as keys / values)
        float q[HD];
        const __half* qp = qkv + (size_t)(g * group_stride + lane * tok_stride) * (3 * C) + h * HD;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            uint4 w = *reinterpret_cast<const uint4*>(qp + i * 8);
            const __half2* h2 = reinterpret_cast<const __half2*>(&w);
#pragma unroll
            for (int e = 0; e < 4; ++e) { float2 f = __half22float2(h2[e]); q[i * 8 + 2 * e] = f.x * scale; q[i * 8 + 2 * e + 1] = f.y * scale; }
        }
        const int nk = causal ? lane + 1 : n_tok;
        float sc[32];
        float m = -INFINITY;
#pragma unroll 1
        for (int u = 0; u < nk; ++u) {
            float s = 0.f;
#pragma unroll
            for (int d = 0; d < HD; d += 2) {
                float2 kf = __half22float2(*reinterpret_cast<const __half2*>(&kv[wid][0][u][d]));
                s = fmaf(q[d], kf.x, s); s = fmaf(q[d + 1], kf.y, s);
            }
            sc[u] = s;
            m = fmaxf(m, s);
        }
        float o[HD];
#pragma unroll
        for (int d = 0; d < HD; ++d) o[d] = 0.f;
        float l = 0.f;
#pragma unroll 1
        for (int u = 0; u < nk; ++u) {
            const float pr = __expf(sc[u] - m);
            l += pr;
#pragma unroll
            for (int d = 0; d < HD; d += 2) {
                float2 vf = __half22float2(*reinterpret_cast<const __half2*>(&kv[wid][1][u][d]));
                o[d] = fmaf(pr, vf.x, o[d]); o[d + 1] = fmaf(pr, vf.y, o[d + 1]);
            }
        }
        const float inv = 1.0f / l;
        __half* yp = y + (size_t)(g * group_stride + lane * tok_stride) * C + h * HD;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            uint4 pk;
            __half2* h2 = reinterpret_cast<__half2*>(&pk);
#pragma unroll
            for (int e = 0; e < 4; ++e) h2[e] = __floats2half2_rn(o[i * 8 + 2 * e] * inv, o[i * 8 + 2 * e + 1] * inv);
            *reinterpret_cast<uint4*>(yp + i * 8) = pk;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Spatial (non-causal, within one frame) flash attention, head dim 48 (module.py:336-338, 349-351).
// mma.sync m16n8k16 fp16 with fp32 online softmax; 4 warps x 32 query rows per CTA, 64-key tiles
// double-buffered with cp.async.  q/k/v are column slices of the fused QKV activation [rows][2304].
// ------------------------------------------------------------------------------------------------
namespace fa {
constexpr int BQ = 128, BKV = 64, LDS = 56;     // smem row pitch 56 halves = 112 B (conflict-free ldmatrix)

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    const uint32_t s = smem_u32(smem);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// grid: (ceil(S / 128), 16 heads, T frames).  v2: every warp owns two 16-row query tiles, so each K / V fragment fetched from shared
// memory (ldmatrix) feeds two MMAs instead of one (v1 was bound by shared-memory wavefronts); the 64-key tile is consumed in two
// 32-key halves to keep the score / probability fragments in registers.
__global__ void __launch_bounds__(128) spatial_attn_kernel(const __half* __restrict__ qkv, __half* __restrict__ y, int S) {
    __shared__ __align__(16) __half sq[BQ][LDS];
    __shared__ __align__(16) __half sk[2][BKV][LDS];
    __shared__ __align__(16) __half sv[2][BKV][LDS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    const int q0 = blockIdx.x * BQ, h = blockIdx.y, t = blockIdx.z;
    const __half* base = qkv + (size_t)t * S * (3 * C) + h * HD;
    // stage Q (128 rows x 6 chunks of 16 B)
    for (int i = tid; i < BQ * 6; i += 128) {
        const int r = i / 6, ch = i - r * 6;
        cp_async16(&sq[r][ch * 8], base + (size_t)min(q0 + r, S - 1) * (3 * C) + ch * 8, q0 + r < S);
    }
    auto load_kv = [&](int buf, int k0) {
        for (int i = tid; i < BKV * 12; i += 128) {
            const int r = i / 12, part = i - r * 12, which = part / 6, ch = part - which * 6;
            const bool ok = k0 + r < S;
            const __half* src = base + (size_t)min(k0 + r, S - 1) * (3 * C) + (1 + which) * C + ch * 8;
            cp_async16(which ? &sv[buf][r][ch * 8] : &sk[buf][r][ch * 8], src, ok);
        }
    };
    load_kv(0, 0);
    cp_async_commit();
    const int n_tiles = (S + BKV - 1) / BKV;

    uint32_t qf[2][3][4];              // Q fragments: 2 row tiles x 3 k-steps of 16 dims
    float o[2][6][4];                  // output accumulators: 2 row tiles x 6 n-tiles of 8 dims
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int i = 0; i < 6; ++i) { o[mt][i][0] = o[mt][i][1] = o[mt][i][2] = o[mt][i][3] = 0.f; }
    float mrow[2][2] = {{-INFINITY, -INFINITY}, {-INFINITY, -INFINITY}}, lrow[2][2] = {{0.f, 0.f}, {0.f, 0.f}};     // [row tile][row g / g + 8]
    const float sl2 = 0.14433756729740643f * 1.4426950408889634f;

    for (int kt = 0; kt < n_tiles; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < n_tiles) load_kv(buf ^ 1, (kt + 1) * BKV);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        if (kt == 0) {
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int ks = 0; ks < 3; ++ks)
                    ldsm_x4(qf[mt][ks][0], qf[mt][ks][1], qf[mt][ks][2], qf[mt][ks][3], &sq[warp * 32 + mt * 16 + (lane & 15)][ks * 16 + (lane >> 4) * 8]);
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            // S = Q K^T : 4 n-tiles (32 keys) x 3 k-steps, both row tiles share the K fragments
            float sc[2][4][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) { sc[mt][nt][0] = sc[mt][nt][1] = sc[mt][nt][2] = sc[mt][nt][3] = 0.f; }
#pragma unroll
            for (int np = 0; np < 2; ++np) {          // pairs of key n-tiles (16 keys)
#pragma unroll
                for (int ks = 0; ks < 3; ++ks) {
                    uint32_t b0, b1, b2, b3;
                    // rows = keys half*32 + np*16 + (lane&7) + 8*(lane>>4), cols = dims ks*16 + 8*((lane>>3)&1)
                    ldsm_x4(b0, b1, b2, b3, &sk[buf][half * 32 + np * 16 + (lane & 7) + ((lane >> 4) << 3)][ks * 16 + (((lane >> 3) & 1) << 3)]);
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        mma16816(sc[mt][2 * np], qf[mt][ks], b0, b1);
                        mma16816(sc[mt][2 * np + 1], qf[mt][ks], b2, b3);
                    }
                }
            }
            // mask keys beyond S, online softmax (rows g = lane>>2 and g+8 of each row tile)
            const int kbase = kt * BKV + half * 32 + (lane & 3) * 2;
            uint32_t pf[2][2][4];            // P as A fragments: 2 row tiles x 2 k-steps of 16 keys
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const int kc = kbase + nt * 8;
                    if (kc >= S) { sc[mt][nt][0] = -INFINITY; sc[mt][nt][2] = -INFINITY; }
                    if (kc + 1 >= S) { sc[mt][nt][1] = -INFINITY; sc[mt][nt][3] = -INFINITY; }
                    mx0 = fmaxf(mx0, fmaxf(sc[mt][nt][0], sc[mt][nt][1]));
                    mx1 = fmaxf(mx1, fmaxf(sc[mt][nt][2], sc[mt][nt][3]));
                }
                mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
                mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
                // a 32-key half can be fully masked (keys >= S): keep the running maximum finite once it is
                const float mn0 = fmaxf(mrow[mt][0], mx0 * sl2), mn1 = fmaxf(mrow[mt][1], mx1 * sl2);
                const float c0 = (mn0 == -INFINITY) ? 1.f : exp2f(mrow[mt][0] - mn0), c1 = (mn1 == -INFINITY) ? 1.f : exp2f(mrow[mt][1] - mn1);
                const float e0 = (mn0 == -INFINITY) ? 0.f : mn0, e1 = (mn1 == -INFINITY) ? 0.f : mn1;
                mrow[mt][0] = mn0; mrow[mt][1] = mn1;
                float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const float p0 = exp2f(sc[mt][nt][0] * sl2 - e0), p1 = exp2f(sc[mt][nt][1] * sl2 - e0);
                    const float p2 = exp2f(sc[mt][nt][2] * sl2 - e1), p3 = exp2f(sc[mt][nt][3] * sl2 - e1);
                    rs0 += p0 + p1; rs1 += p2 + p3;
                    pf[mt][nt >> 1][(nt & 1) * 2 + 0] = pack_h2(p0, p1);
                    pf[mt][nt >> 1][(nt & 1) * 2 + 1] = pack_h2(p2, p3);
                }
                lrow[mt][0] = lrow[mt][0] * c0 + rs0; lrow[mt][1] = lrow[mt][1] * c1 + rs1;
#pragma unroll
                for (int i = 0; i < 6; ++i) { o[mt][i][0] *= c0; o[mt][i][1] *= c0; o[mt][i][2] *= c1; o[mt][i][3] *= c1; }
            }
            // O += P V : 2 k-steps (16 keys) x 6 n-tiles (8 dims); V fragments via transposed ldmatrix, shared by both row tiles
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
                for (int dp = 0; dp < 3; ++dp) {      // pairs of dim n-tiles (16 dims)
                    uint32_t b0, b1, b2, b3;
                    // rows = keys half*32 + ks*16 + (lane&7) + 8*((lane>>3)&1), cols = dims dp*16 + 8*(lane>>4)
                    ldsm_x4_t(b0, b1, b2, b3, &sv[buf][half * 32 + ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3)][dp * 16 + ((lane >> 4) << 3)]);
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        mma16816(o[mt][2 * dp], pf[mt][ks], b0, b1);
                        mma16816(o[mt][2 * dp + 1], pf[mt][ks], b2, b3);
                    }
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        float l0 = lrow[mt][0], l1 = lrow[mt][1];
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float i0 = 1.0f / l0, i1 = 1.0f / l1;
        const int r0 = q0 + warp * 32 + mt * 16 + (lane >> 2), r1 = r0 + 8;
        __half* yb = y + (size_t)t * S * C + h * HD + (lane & 3) * 2;
#pragma unroll
        for (int nt = 0; nt < 6; ++nt) {
            if (r0 < S) *reinterpret_cast<uint32_t*>(yb + (size_t)r0 * C + nt * 8) = pack_h2(o[mt][nt][0] * i0, o[mt][nt][1] * i0);
            if (r1 < S) *reinterpret_cast<uint32_t*>(yb + (size_t)r1 * C + nt * 8) = pack_h2(o[mt][nt][2] * i1, o[mt][nt][3] * i1);
        }
    }
}
}  // namespace fa

// ------------------------------------------------------------------------------------------------
// Ego cross attention: nq (<= 4) query rows against n_k scene rows, one CTA per head (module.py:482-509)
// q fp16 [nq][768], k/v fp16 [n_k][768], y fp16 [nq][768]; one CTA per (head, query)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cross_attn_kernel(const __half* __restrict__ q, const __half* __restrict__ k, const __half* __restrict__ v,
                                                         __half* __restrict__ y, int n_k) {
    __shared__ float qs[HD];
    __shared__ float red_m[8], red_l[8], red_o[8][HD];
    const int h = blockIdx.x, qi = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < HD) qs[tid] = __half2float(q[(size_t)qi * C + h * HD + tid]) * 0.14433756729740643f;
    __syncthreads();
    float m = -INFINITY, l = 0.f, o[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = 0.f;
    for (int j = tid; j < n_k; j += 256) {
        const __half* kp = k + (size_t)j * C + h * HD;
        const __half* vp = v + (size_t)j * C + h * HD;
        float s = 0.f;
#pragma unroll
        for (int c8 = 0; c8 < 6; ++c8) {
            uint4 wk = *reinterpret_cast<const uint4*>(kp + c8 * 8);
            const __half2* hk = reinterpret_cast<const __half2*>(&wk);
#pragma unroll
            for (int e = 0; e < 4; ++e) { float2 a = __half22float2(hk[e]); s = fmaf(qs[c8 * 8 + 2 * e], a.x, s); s = fmaf(qs[c8 * 8 + 2 * e + 1], a.y, s); }
        }
        const float mn = fmaxf(m, s), c = __expf(m - mn), pr = __expf(s - mn);
        l = l * c + pr;
#pragma unroll
        for (int c8 = 0; c8 < 6; ++c8) {
            uint4 wv = *reinterpret_cast<const uint4*>(vp + c8 * 8);
            const __half2* hv = reinterpret_cast<const __half2*>(&wv);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float2 b = __half22float2(hv[e]);
                o[c8 * 8 + 2 * e] = fmaf(pr, b.x, o[c8 * 8 + 2 * e] * c);
                o[c8 * 8 + 2 * e + 1] = fmaf(pr, b.y, o[c8 * 8 + 2 * e + 1] * c);
            }
        }
        m = mn;
    }
    // merge the 256 threads: warp shuffle tree, then across the 8 warps through shared memory
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float mo = __shfl_xor_sync(0xffffffffu, m, off), lo = __shfl_xor_sync(0xffffffffu, l, off);
        const float mn = fmaxf(m, mo);
        const float ca = (m > -INFINITY) ? __expf(m - mn) : 0.f, cb = (mo > -INFINITY) ? __expf(mo - mn) : 0.f;
        l = l * ca + lo * cb;
#pragma unroll
        for (int d = 0; d < HD; ++d) { const float oo = __shfl_xor_sync(0xffffffffu, o[d], off); o[d] = o[d] * ca + oo * cb; }
        m = mn;
    }
    if (lane == 0) {
        red_m[warp] = m; red_l[warp] = l;
#pragma unroll
        for (int d = 0; d < HD; ++d) red_o[warp][d] = o[d];
    }
    __syncthreads();
    if (tid < HD) {
        float mm = -INFINITY;
        for (int w = 0; w < 8; ++w) mm = fmaxf(mm, red_m[w]);
        float ll = 0.f, oo = 0.f;
        for (int w = 0; w < 8; ++w) {
            const float f = (red_m[w] > -INFINITY) ? __expf(red_m[w] - mm) : 0.f;
            ll += f * red_l[w];
            oo += f * red_o[w][tid];
        }
        y[(size_t)qi * C + h * HD + tid] = __float2half_rn(oo / ll);
    }
}

// ------------------------------------------------------------------------------------------------
// top-k sampling of logits rows (ego head; UMGen.py:899-913, 1001-1004).  One warp per row.
// ------------------------------------------------------------------------------------------------
__global__ void sample_rows_kernel(const float* __restrict__ logits, int V, int k, float inv_temp, uint64_t seed, uint32_t frame,
                                   int* __restrict__ out) {
    extern __shared__ float vals[];
    const int row = blockIdx.x, lane = threadIdx.x;
    for (int i = lane; i < V; i += 32) vals[i] = logits[(size_t)row * V + i];
    __syncwarp();
    float selv = -INFINITY;
    int seli = 0x7fffffff;
    for (int r = 0; r < k; ++r) {
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int i = lane; i < V; i += 32) { const float v = vals[i]; if (v > bv) { bv = v; bi = i; } }
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == r) { selv = bv; seli = bi; }
        if (lane == 0 && bi != 0x7fffffff) vals[bi] = -INFINITY;
        __syncwarp();
    }
    const float vmax = __shfl_sync(0xffffffffu, selv, 0);
    const float w = (lane < k && selv > -INFINITY) ? __expf((selv - vmax) * inv_temp) : 0.f;
    float cum = w;
    for (int o = 1; o < 32; o <<= 1) { const float tv = __shfl_up_sync(0xffffffffu, cum, o); if (lane >= o) cum += tv; }
    const float total = __shfl_sync(0xffffffffu, cum, 31);
    const float u = philox_uniform(seed, frame, 0x70000000u + row, 0u);
    const unsigned hit = __ballot_sync(0xffffffffu, (w > 0.f) && (cum > u * total));
    const int pick = hit ? (__ffs(hit) - 1) : 0;
    const int tok = __shfl_sync(0xffffffffu, seli, pick);
    if (lane == 0) out[row] = tok;
}

// sample_top_p (UMGen.py:915-965) on logits rows: one block of N_CONS threads per row, the decode kernels' block sampler
struct ToppSmem {
    float red[64];
    volatile int tok;
};
__global__ void __launch_bounds__(N_CONS) sample_rows_topp_kernel(const float* __restrict__ logits, int V, float p, float inv_temp, uint64_t seed,
                                                                   uint32_t frame, int* __restrict__ out) {
    __shared__ ToppSmem sm;
    const int row = blockIdx.x, tid = threadIdx.x;
    float v[TOPP_PER];
#pragma unroll
    for (int i = 0; i < TOPP_PER; ++i) {
        const int id = tid + i * N_CONS;
        v[i] = id < V ? logits[(size_t)row * V + id] : -INFINITY;
    }
    const float u = philox_uniform(seed, frame, 0x70000000u + row, 0u);
    const int pick = block_topp_sample(&sm, v, p, inv_temp, u, tid);
    if (tid == 0) out[row] = pick;
}

// ------------------------------------------------------------------------------------------------
// OAR conditioning feature of the last frame (UMGen.py:1496-1511): pose/image rows from TAR, map rows
// from map_tar plus the warped-map prior on content cells, bbox3d rows from box_tar.
// ------------------------------------------------------------------------------------------------
__global__ void assemble_tar_feat_kernel(const float* __restrict__ f_all, const float* __restrict__ f_map, const float* __restrict__ f_box,
                                         const float* __restrict__ warped_last, float* __restrict__ out, int row0) {
    const int pos = row0 + blockIdx.x;
    const float* src = f_all + (size_t)pos * C;
    const float* add = nullptr;
    if (pos >= 5 && pos < 1031) {
        src = f_map + (size_t)pos * C;
        if (pos >= 6 && pos < 1030) add = warped_last + (size_t)(pos - 6) * C;
    } else if (pos >= 1031 && pos < 1693) {
        src = f_box + (size_t)pos * C;
    }
    for (int c = threadIdx.x; c < C; c += blockDim.x) out[(size_t)pos * C + c] = src[c] + (add ? add[c] : 0.f);
}

}  // namespace umgen

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using namespace umgen;
#define ST(s) ((cudaStream_t)(s))

extern "C" int umgen_layernorm(const void* x_f, const void* w_f, void* out, int64_t rows, int out_half, void* stream) {
    if (!x_f || !w_f || !out || rows < 1) { set_error("layernorm: bad args"); return -1; }
    const int grid = (int)((rows + 7) / 8);
    if (out_half) ln_rows_kernel<true><<<grid, 256, 0, ST(stream)>>>((const float*)x_f, (const float*)w_f, out, (int)rows);
    else ln_rows_kernel<false><<<grid, 256, 0, ST(stream)>>>((const float*)x_f, (const float*)w_f, out, (int)rows);
    UMGEN_CUDA_OK(cudaGetLastError());
    g_launches += 1;
    return 0;
}
extern "C" int umgen_cast_f16(const void* x_f, void* out_h, int64_t n, void* stream) {
    cast_f16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>((const float*)x_f, (__half*)out_h, (size_t)n);
    UMGEN_CUDA_OK(cudaGetLastError());
    g_launches += 1;
    return 0;
}
extern "C" int umgen_map_feature(const void* tok_i32, const void* table_f, const void* grid_pos_f, void* out_f, int64_t n_tok, void* stream) {
    map_feature_kernel<<<(unsigned)n_tok, 192, 0, ST(stream)>>>((const int*)tok_i32, (const float*)table_f, (const float*)grid_pos_f, (float*)out_f, (int)n_tok);
    UMGEN_CUDA_OK(cudaGetLastError());
    g_launches += 1;
    return 0;
}
extern "C" int umgen_map_warp(const void* feat_f, const void* pose_tok_i32, const void* pose_lut_f, void* out_f, int64_t T, void* stream) {
    map_warp_kernel<<<(unsigned)(T * 1024), 192, 0, ST(stream)>>>((const float*)feat_f, (const int*)pose_tok_i32, (const float*)pose_lut_f, (float*)out_f, (int)T);
    UMGEN_CUDA_OK(cudaGetLastError());
    g_launches += 1;
    return 0;
}
extern "C" int umgen_embed_sequence(const UmgenEmbedArgs* a, void* stream) {
    if (!a || a->T < 1 || (a->n_mods != 2 && a->n_mods != 3 && a->n_mods != 4)) { set_error("embed: bad args"); return -1; }
    EmbedArgs e;
    e.pose = (const int*)a->pose_i32; e.map = (const int*)a->map_i32; e.bbox = (const int*)a->bbox_i32; e.image = (const int*)a->image_i32;
    e.fpe = (const float*)a->fpe_f; e.img_table = (const float*)a->img_table_f; e.be = (const float*)a->be_f; e.axe = (const float*)a->axe_f;
    e.spe = (const float*)a->spe_f; e.tpe = (const float*)a->tpe_f; e.sp = (const float*)a->spatial_f;
    e.map_feat = (const float*)a->map_feat_f; e.map_warped = (const float*)a->map_warped_f; e.out = (float*)a->out_f;
    e.T = (int)a->T; e.n_mods = (int)a->n_mods; e.t_offset = (int)a->t_offset;
    e.S = a->n_mods == 2 ? 1031 : (a->n_mods == 3 ? 1693 : 2207);
    embed_sequence_kernel<<<(unsigned)(e.T * e.S), 192, 0, ST(stream)>>>(e);
    UMGEN_CUDA_OK(cudaGetLastError());
    g_launches += 1;
    return 0;
}
extern "C" int umgen_small_attention_from(const void* qkv_h, void* y_h, int64_t n_groups, int64_t n_tok, int64_t group_stride, int64_t tok_stride,
                                          int causal, int64_t q0, void* stream) {
    if (n_tok < 1 || n_tok > 32 || q0 < 0 || q0 >= n_tok) { set_error("small attention handles 1..32 tokens (got %lld, first query %lld)", (long long)n_tok, (long long)q0); return -1; }
    const long long warps = n_groups * NH;
    small_attn_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, ST(stream)>>>((const __half*)qkv_h, (__half*)y_h, (int)n_groups, (int)n_tok,
                                                                           (long long)group_stride, (long long)tok_stride, causal, (int)q0);
    UMGEN_CUDA_OK(cudaGetLastError());
    g_launches += 1;
    return 0;
}
extern "C" int umgen_small_attention(const void* qkv_h, void* y_h, int64_t n_groups, int64_t n_tok, int64_t group_stride, int64_t tok_stride,
                                     int causal, void* stream) {
    return umgen_small_attention_from(qkv_h, y_h, n_groups, n_tok, group_stride, tok_stride, causal, 0, stream);
}
extern "C" int umgen_spatial_attention(const void* qkv_h, void* y_h, int64_t T, int64_t S, void* stream) {
    if (T < 1 || S < 1) { set_error("spatial attention: bad shape"); return -1; }
    dim3 grid((unsigned)((S + fa::BQ - 1) / fa::BQ), NH, (unsigned)T);
    fa::spatial_attn_kernel<<<grid, 128, 0, ST(stream)>>>((const __half*)qkv_h, (__half*)y_h, (int)S);
    UMGEN_CUDA_OK(cudaGetLastError());
    g_launches += 1;
    return 0;
}
extern "C" int umgen_cross_attention(const void* q_h, const void* k_h, const void* v_h, void* y_h, int64_t nq, int64_t n_k, void* stream) {
    if (nq < 1 || nq > 64 || n_k < 1) { set_error("cross attention: bad shape"); return -1; }
    cross_attn_kernel<<<dim3(NH, (unsigned)nq), 256, 0, ST(stream)>>>((const __half*)q_h, (const __half*)k_h, (const __half*)v_h, (__half*)y_h, (int)n_k);
    UMGEN_CUDA_OK(cudaGetLastError());
    g_launches += 1;
    return 0;
}
extern "C" int umgen_sample_rows(const void* logits_f, int64_t rows, int64_t V, int64_t top_k, double top_p, double temperature, uint64_t seed,
                                 int64_t frame_index, void* out_i32, void* stream) {
    if (!(temperature > 0)) { set_error("sample_rows: temperature must be > 0"); return -1; }
    if (top_p > 0) {
        if (V > TOPP_PER * N_CONS) { set_error("sample_rows: V <= %d in top-p mode", TOPP_PER * N_CONS); return -1; }
        sample_rows_topp_kernel<<<(unsigned)rows, N_CONS, 0, ST(stream)>>>((const float*)logits_f, (int)V, (float)top_p, (float)(1.0 / temperature), seed,
                                                                         (uint32_t)frame_index, (int*)out_i32);
        UMGEN_CUDA_OK(cudaGetLastError());
        g_launches += 1;
        return 0;
    }
    if (top_k < 1 || top_k > 32 || V > 12000) { set_error("sample_rows: top_k in [1,32], V <= 12000"); return -1; }
    sample_rows_kernel<<<(unsigned)rows, 32, (size_t)V * 4, ST(stream)>>>((const float*)logits_f, (int)V, (int)top_k, (float)(1.0 / temperature), seed,
                                                                         (uint32_t)frame_index, (int*)out_i32);
    UMGEN_CUDA_OK(cudaGetLastError());
    g_launches += 1;
    return 0;
}
extern "C" int umgen_assemble_tar_feat(const void* f_all, const void* f_map, const void* f_box, const void* warped_last, void* out, int64_t row0, int64_t row1,
                                       void* stream) {
    if (row0 < 0 || row1 > SEQ || row0 >= row1) { set_error("assemble_tar_feat: bad row range"); return -1; }
    assemble_tar_feat_kernel<<<(unsigned)(row1 - row0), 192, 0, ST(stream)>>>((const float*)f_all, (const float*)f_map, (const float*)f_box, (const float*)warped_last,
                                                                              (float*)out, (int)row0);
    UMGEN_CUDA_OK(cudaGetLastError());
    g_launches += 1;
    return 0;
}

// Lazy module loading (the CUDA 12 default) loads a kernel on its first launch and that load waits for an idle device -- which never comes while the
// persistent decode kernel spins on a flag.  umgen_preload() (capi.cu) forces every kernel of the library to load up front.
#define UMGEN_PRELOAD(k) UMGEN_CUDA_OK(cudaFuncGetAttributes(&fa_, k))
namespace umgen {
int preload_tar() {
    cudaFuncAttributes fa_;
    UMGEN_PRELOAD(ln_rows_kernel<true>); UMGEN_PRELOAD(ln_rows_kernel<false>); UMGEN_PRELOAD(cast_f16_kernel); UMGEN_PRELOAD(map_feature_kernel);
    UMGEN_PRELOAD(map_warp_kernel); UMGEN_PRELOAD(embed_sequence_kernel); UMGEN_PRELOAD(small_attn_kernel); UMGEN_PRELOAD(fa::spatial_attn_kernel);
    UMGEN_PRELOAD(cross_attn_kernel); UMGEN_PRELOAD(sample_rows_kernel); UMGEN_PRELOAD(sample_rows_topp_kernel); UMGEN_PRELOAD(assemble_tar_feat_kernel);
    return 0;
}
}  // namespace umgen
