// Spatial (non-causal, within one frame) flash attention of the TAR blocks on the 5th-generation tensor cores, head dim 48
// (reference models/module.py:218 flash_attn_func as called by BlockTAR.forward_func, module.py:336-338, 349-351).
//
// One CTA = one 128-row query tile of one (frame, head); two CTAs are resident per SM (256 TMEM columns, 64 KB of shared memory and 6 warps each).
// Both contractions are tcgen05.mma with the A operand in tensor memory:
//   S[128 x 64 keys] = Q . K^T   A = Q (fp16, written to TMEM once by the softmax threads), B = K tile in shared memory, K-major, 3 k-steps of 16 dims
//   O[128 x 48]     += P . V     A = P (fp16, written over the S columns by the softmax threads), B = V tile in shared memory, MN-major, 4 k-steps of 16 keys
// K / V tiles (64 keys x 64 dims: the 48 dims of the head + 16 unused ones, so that a row is one 128-byte swizzle atom) arrive by TMA
// (3-D tensor map [frame][row][column] over the fused qkv activation, 128-byte swizzle, rows beyond the frame zero-filled) into a 4-stage ring.
// The score accumulator is double buffered: S(j+2) is issued as soon as P(j) has been consumed, so the softmax warps never wait for the tensor
// cores and the kernel runs at the rate of its slowest pipe -- MUFU (exp2: 16 per clock and SM against 4 D = 192 tensor flops per score).
//   warps 0-3   softmax: thread = one query row; scores by tcgen05.ld, fp32 online softmax with exp2, P by tcgen05.st; O is rescaled in TMEM only
//               when a row's running maximum grew by more than 2^8 (any reference maximum is exact as long as nothing overflows)
//   warp 4      TMA producer (+ TMEM allocation)
//   warp 5      MMA issuer: S(0) S(1) | PV(j) S(j+2) ...; issued by one elect.sync lane of the converged warp -- under a plain `lane == 0` branch the
//               compiler wraps every tcgen05.mma in an election loop that costs ~70 clocks per instruction, more than these small MMAs take
#include <cuda.h>

#include "common.cuh"
#include "../../include/umgen.h"

namespace umgen {
extern int64_t g_launches;
namespace attn {

constexpr int BQ = 128, BKV = 64, NST = 4;
constexpr int ROW_B = 128;                         // bytes per staged K / V row (64 halves)
constexpr int TILE_B = BKV * ROW_B;                // 8 KB
constexpr int THREADS = 192;
constexpr int W_PROD = 4, W_MMA = 5;
constexpr uint32_t TMEM_COLS = 256;
// two score / probability buffers of 64 columns, O 48 columns, Q 24 columns
__device__ __forceinline__ constexpr uint32_t col_s(int b) { return 64u * b; }
constexpr uint32_t COL_O = 128, COL_Q = 192;
constexpr float SL2 = 0.14433756729740643f * 1.4426950408889634f;     // 1/sqrt(48) * log2(e)  (module.py:196-198)
constexpr float RESCALE_GAP = 8.0f;                // log2 units a row maximum may lag behind before O is rescaled

struct __align__(1024) Smem {
    uint8_t kv[NST][2][TILE_B];                    // [stage][K | V][64 rows x 128 B], 128-byte swizzle
    uint64_t kv_full[NST], kv_empty[NST];
    uint64_t q_ready, s_full[2], p_full[2], pv_done, o_full;
    uint32_t tmem_base;
};

struct Params {
    const __half* qkv;      // [T][S][2304]
    __half* y;              // [T][S][768]
    int S, n_tiles;
    float* dbg;             // optional: CTA (0,0,0) dumps S of the first key tile [128][64], then l [128], then O [128][48]
};

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(smem)),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}
// shared-memory operand descriptor, 128-byte swizzle, 8-row groups 1024 bytes apart (K-major: 8 rows of the N dimension; MN-major: 8 rows of K)
__device__ __forceinline__ uint64_t smem_desc(const void* smem, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(smem) >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;                // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                // SWIZZLE_128B
    return d;
}
// kind::f16, fp16 operands, fp32 accumulate, M = 128; b_mn = 1: B is MN-major (V tile: [key][dim], dims contiguous)
__device__ __forceinline__ uint32_t idesc(int n, int b_mn) {
    return (1u << 4) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
}
// D[tmem] (+)= A[tmem] . B[smem]
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t id, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(id), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// one lane of a converged warp (elect.sync): the form the compiler turns into straight-line UTCHMMA sequences
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred)::"memory");
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

#define R8(r, o) "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]), "=r"(r[o + 7])
#define W8(r, o) "r"(r[o + 0]), "r"(r[o + 1]), "r"(r[o + 2]), "r"(r[o + 3]), "r"(r[o + 4]), "r"(r[o + 5]), "r"(r[o + 6]), "r"(r[o + 7])
// 32 lanes x 32 / 16 columns of 32 bits: thread i of the warp gets columns taddr.col .. of lane taddr.lane + i (no wait)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, "
        "%27, %28, %29, %30, %31}, [%32];"
        : R8(r, 0), R8(r, 8), R8(r, 16), R8(r, 24)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : R8(r, 0), R8(r, 8)
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
                 W8(r, 0), W8(r, 8)
                 : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), W8(r, 0) : "memory");
}
__device__ __forceinline__ float ex2f(float x) {          // ex2(-inf) = 0
#ifdef ATTN_NOEXP      // experiment: how long is a tile without the MUFU work (wrong results)
    return x * 1e-3f;
#else
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#endif
}
#ifdef ATTN_PROF        // experiment: clock stamps of one softmax thread of CTA (1,0,0) into p.dbg (as long long): [tile 64][8]
#define PSTAMP(j, k) if (prof && (j) < 64) prof[(j) * 8 + (k)] = clock64();
#else
#define PSTAMP(j, k)
#endif
// 2^x on the FMA / ALU pipes: x = n + f with n = round(x), f in [-0.5, 0.5]; 2^f by a cubic (max relative error 7.5e-5, a sixth of the fp16 half-ulp of
// the probability it becomes), 2^n by adding n to the exponent field.  The kernel is bound by the 16 exp2 per clock of the MUFU pipe; every
// POLY_EVERY-th score takes this path instead and the two pipes share the work.
#ifndef ATTN_POLY_EVERY
#define ATTN_POLY_EVERY 4
#endif
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -126.0f);                                 // also turns the -inf of a masked key into ~1e-38 (0 in fp16)
    const float t = x + 12582912.0f;                       // 1.5 * 2^23: the low mantissa bits of t hold round(x)
    const float f = x - (t - 12582912.0f);
    float p = fmaf(f, 0.05517157167196274f, 0.2426111102104187f);
    p = fmaf(p, f, 0.6932610273361206f);
    p = fmaf(p, f, 0.9999280571937561f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// grid: (ceil(S / 128), 16 heads, T frames)
__global__ void __launch_bounds__(THREADS, 2) spatial_attn_tc_kernel(const __grid_constant__ CUtensorMap map_kv, const __grid_constant__ Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    Smem* sm = reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * BQ, h = blockIdx.y, t = blockIdx.z;
    const int S = p.S, n_tiles = p.n_tiles;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NST; ++i) { mbar_init(&sm->kv_full[i], 1); mbar_init(&sm->kv_empty[i], 1); }
        mbar_init(&sm->q_ready, 4); mbar_init(&sm->o_full, 1); mbar_init(&sm->pv_done, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&sm->s_full[i], 1); mbar_init(&sm->p_full[i], 4); }
        mbar_fence_init();
    }
    if (warp == W_PROD) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm->tmem_base)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm->tmem_base;

    if (warp == W_PROD) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_kv) : "memory");
            for (int j = 0; j < n_tiles; ++j) {
                const uint32_t s = j % NST, ph = (j / NST) & 1;
                mbar_wait(&sm->kv_empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&sm->kv_full[s], 2 * TILE_B);
                tma_load_3d(sm->kv[s][0], &map_kv, C + h * HD, j * BKV, t, &sm->kv_full[s]);
                tma_load_3d(sm->kv[s][1], &map_kv, 2 * C + h * HD, j * BKV, t, &sm->kv_full[s]);
            }
        }
    } else if (warp == W_MMA) {
        // ===================== MMA issuer: the whole warp walks the schedule, one elected lane issues =====================
        const uint32_t id_s = idesc(BKV, 0), id_o = idesc(HD, 1);
        auto issue_s = [&](int j) {                      // S(j) = Q K_j^T into buffer j & 1: 3 k-steps of 16 dims = 32 bytes inside the swizzle atom
            const uint32_t s = j % NST;
            mbar_wait(&sm->kv_full[s], (j / NST) & 1);
            tc_fence_after();
            const uint64_t dk = smem_desc(sm->kv[s][0], 0);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) umma_ts(tmem + col_s(j & 1), tmem + COL_Q + 8 * k, dk + 2 * k, id_s, k != 0);
                umma_commit(&sm->s_full[j & 1]);
            }
            __syncwarp();
        };
        mbar_wait(&sm->q_ready, 0);
        issue_s(0);
        if (n_tiles > 1) issue_s(1);
        for (int j = 0; j < n_tiles; ++j) {
            const uint32_t s = j % NST;
            mbar_wait(&sm->p_full[j & 1], (j >> 1) & 1);
            tc_fence_after();
            const uint64_t dv = smem_desc(sm->kv[s][1], BKV * ROW_B);
            if (elect_one()) {                           // O += P(j) V_j: 4 k-steps of 16 keys = two 8-row groups = 2048 bytes
#pragma unroll
                for (int k = 0; k < BKV / 16; ++k)
                    umma_ts(tmem + COL_O, tmem + col_s(j & 1) + 8 * k, dv + (uint64_t)(k * (16 * ROW_B >> 4)), id_o, (j | k) != 0);
                umma_commit(&sm->kv_empty[s]);           // K_j (read by S(j) two steps ago) and V_j are free once these retire
                umma_commit(&sm->pv_done);
            }
            __syncwarp();
            if (j + 2 < n_tiles) issue_s(j + 2);         // the score buffer P(j) lived in is free again (in-order tensor pipe)
        }
        if (elect_one()) umma_commit(&sm->o_full);
        __syncwarp();
    } else {
        // ===================== softmax warps: one query row per thread =====================
        const int quarter = warp & 3;                      // the TMEM lanes this warp may touch
        const int r = quarter * 32 + lane, row = q0 + r;
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        const uint32_t c_o = tmem + lane_base + COL_O, c_q = tmem + lane_base + COL_Q;
        const bool dump = p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
        {       // Q row -> TMEM (two halves per column: the A operand of S = Q K^T)
            uint32_t q[24];
            if (row < S) {
                const uint4* src = reinterpret_cast<const uint4*>(p.qkv + ((size_t)t * S + row) * (3 * C) + h * HD);
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    const uint4 v = __ldg(src + i);
                    q[4 * i] = v.x; q[4 * i + 1] = v.y; q[4 * i + 2] = v.z; q[4 * i + 3] = v.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 24; ++i) q[i] = 0u;
            }
            tmem_st8(c_q, q);
            tmem_st8(c_q + 8, q + 8);
            tmem_st8(c_q + 16, q + 16);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm->q_ready);
        }
#ifdef ATTN_PROF
        long long* prof = (p.dbg && r == 0 && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0) ? (long long*)p.dbg : nullptr;
#endif
        float m_run = -INFINITY, l_run = 0.f;
        for (int j = 0; j < n_tiles; ++j) {
            const uint32_t c_s = tmem + lane_base + col_s(j & 1);
            PSTAMP(j, 0)
            mbar_wait(&sm->s_full[j & 1], (j >> 1) & 1);
            PSTAMP(j, 1)
            tc_fence_after();
            uint32_t sc[BKV];
#pragma unroll
            for (int c = 0; c < BKV / 32; ++c) tmem_ld32(c_s + 32 * c, sc + 32 * c);
            tmem_wait_ld();
            PSTAMP(j, 2)
            const int nvalid = S - j * BKV;                // keys of this tile inside the frame (the rest is zero-filled)
            if (nvalid < BKV) {
#pragma unroll
                for (int k = 0; k < BKV; ++k)
                    if (k >= nvalid) sc[k] = 0xff800000u;  // -inf
            }
            if (dump && j == 0) {
                for (int k = 0; k < BKV; ++k) p.dbg[r * BKV + k] = __uint_as_float(sc[k]);
            }
            float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int k = 0; k < BKV; ++k) mx4[k & 3] = fmaxf(mx4[k & 3], __uint_as_float(sc[k]));
            const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
            // the reference maximum only moves when the row maximum outgrew it by RESCALE_GAP (p <= 2^8 fits fp16 with room to spare)
            const float m_tile = mx * SL2;
            const bool grow = m_tile > m_run + RESCALE_GAP;
            const float m_new = grow ? m_tile : m_run;
            if (__any_sync(0xffffffffu, grow) && j > 0) {
                const float alpha = ex2f(m_run - m_new);   // 1 for the rows that keep their maximum
                l_run *= alpha;
                mbar_wait(&sm->pv_done, (j - 1) & 1);       // O must hold P(j-1) V before it is touched (S(j) was issued ahead of that product)
                tc_fence_after();
                uint32_t o[HD];
#pragma unroll
                for (int c = 0; c < HD / 16; ++c) tmem_ld16(c_o + 16 * c, o + 16 * c);
                tmem_wait_ld();
#pragma unroll
                for (int d = 0; d < HD; ++d) o[d] = __float_as_uint(__uint_as_float(o[d]) * alpha);
#pragma unroll
                for (int c = 0; c < HD / 16; ++c) tmem_st16(c_o + 16 * c, o + 16 * c);
            }
            m_run = m_new;
            PSTAMP(j, 3)
            // p = 2^(s * SL2 - m), written as fp16 pairs over the first 32 columns of the score buffer (the A operand of O += P V)
            float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < BKV / 32; ++c) {
                uint32_t pk[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const float x0 = fmaf(__uint_as_float(sc[32 * c + 2 * k]), SL2, -m_run), x1 = fmaf(__uint_as_float(sc[32 * c + 2 * k + 1]), SL2, -m_run);
                    const float p0 = ex2f(x0);
                    const float p1 = (ATTN_POLY_EVERY > 0 && (2 * k + 1) % ATTN_POLY_EVERY == ATTN_POLY_EVERY - 1) ? ex2_poly(x1) : ex2f(x1);
                    ls[k & 3] += p0 + p1;
                    pk[k] = pack_h2(p0, p1);
                }
                tmem_st16(c_s + 16 * c, pk);
            }
            l_run += (ls[0] + ls[1]) + (ls[2] + ls[3]);
            PSTAMP(j, 4)
            tmem_wait_st();
            PSTAMP(j, 5)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm->p_full[j & 1]);
            PSTAMP(j, 6)
        }
        // O / l -> y[t][row][h * 48 ..]
        mbar_wait(&sm->o_full, 0);
        tc_fence_after();
        uint32_t o[HD];
#pragma unroll
        for (int c = 0; c < HD / 16; ++c) tmem_ld16(c_o + 16 * c, o + 16 * c);
        tmem_wait_ld();
        if (dump) {
            p.dbg[BQ * BKV + r] = l_run;
            for (int d = 0; d < HD; ++d) p.dbg[BQ * BKV + BQ + r * HD + d] = __uint_as_float(o[d]);
        }
        if (row < S) {
            const float inv = 1.0f / l_run;
            uint4* dst = reinterpret_cast<uint4*>(p.y + ((size_t)t * S + row) * C + h * HD);
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                uint4 v;
                v.x = pack_h2(__uint_as_float(o[8 * i]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
                v.y = pack_h2(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
                v.z = pack_h2(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
                v.w = pack_h2(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
                dst[i] = v;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_PROD) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

// ---- host -----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static int get_encode() {
    if (g_encode) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        set_error("cuTensorMapEncodeTiled not available (%s)", cudaGetErrorString(e));
        return -3;
    }
    g_encode = (EncodeTiledFn)fn;
    return 0;
}
// the fused qkv activation as [T frames][S rows][2304 columns] fp16; box = 64 columns x 128 rows of one frame
struct MapKey { const void* ptr; int64_t T, S; };
struct MapSlot { MapKey key; CUtensorMap map; };
static MapSlot g_maps[64];
static int g_nmaps = 0, g_next = 0;
static int get_map(const void* qkv, int64_t T, int64_t S, const CUtensorMap** out) {
    for (int i = 0; i < g_nmaps; ++i)
        if (g_maps[i].key.ptr == qkv && g_maps[i].key.T == T && g_maps[i].key.S == S) { *out = &g_maps[i].map; return 0; }
    if (int rc = get_encode()) return rc;
    MapSlot& s = g_maps[g_next];
    cuuint64_t dims[3] = {(cuuint64_t)(3 * C), (cuuint64_t)S, (cuuint64_t)T};
    cuuint64_t strides[2] = {(cuuint64_t)(3 * C) * 2, (cuuint64_t)S * (3 * C) * 2};
    cuuint32_t box[3] = {64, BKV, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(&s.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(qkv), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (attention) failed (%d) T=%lld S=%lld", (int)r, (long long)T, (long long)S); return -2; }
    s.key = MapKey{qkv, T, S};
    *out = &s.map;
    g_next = (g_next + 1) % 64;
    if (g_nmaps < 64) g_nmaps++;
    return 0;
}

}  // namespace attn

int preload_attn() {
    cudaFuncAttributes fa_;
    UMGEN_CUDA_OK(cudaFuncGetAttributes(&fa_, attn::spatial_attn_tc_kernel));
    return 0;
}
}  // namespace umgen

using namespace umgen;

extern "C" int umgen_spatial_attention_tc(const void* qkv_h, void* y_h, int64_t T, int64_t S, void* dbg_f, void* stream) {
    using namespace umgen::attn;
    if (!qkv_h || !y_h || T < 1 || S < 1) { set_error("spatial attention: bad arguments"); return -1; }
    if (((uintptr_t)qkv_h & 15) != 0) { set_error("spatial attention: qkv must be 16-byte aligned"); return -1; }
    static bool configured = false;
    constexpr int smem = (int)sizeof(Smem) + 1024;
    if (!configured) {
        UMGEN_CUDA_OK(cudaFuncSetAttribute(spatial_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    const CUtensorMap* map = nullptr;
    if (int rc = get_map(qkv_h, T, S, &map)) return rc;
    Params p;
    p.qkv = (const __half*)qkv_h; p.y = (__half*)y_h; p.S = (int)S; p.n_tiles = (int)((S + BKV - 1) / BKV); p.dbg = (float*)dbg_f;
    dim3 grid((unsigned)((S + BQ - 1) / BQ), NH, (unsigned)T);
    spatial_attn_tc_kernel<<<grid, THREADS, smem, (cudaStream_t)stream>>>(*map, p);
    {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            cudaFuncAttributes fa;
            cudaFuncGetAttributes(&fa, spatial_attn_tc_kernel);
            set_error("spatial attention launch: %s (threads %d, max threads %d, regs %d, static smem %zu, dynamic smem %d, max dynamic %d)", cudaGetErrorString(e),
                      THREADS, fa.maxThreadsPerBlock, fa.numRegs, fa.sharedSizeBytes, smem, fa.maxDynamicSharedSizeBytes);
            return -2;
        }
    }
    g_launches += 1;
    return 0;
}
