// Persistent warp-specialised tcgen05 GEMM for the TAR encoders:  D[M,N] = epilogue(A[M,K] . W[N,K]^T)
//
// Replaces the F.linear calls of BlockTAR / Decoder / GMLP under fp16 autocast (reference
// models/module.py:206,229,246-248,485-487,724-726): fp16 operands, fp32 accumulation in TMEM.
//   warp 0      TMA producer: A and W tiles (K-major, 128-byte swizzle) into a 4-stage shared-memory ring
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma (M=128, N=256, K=16) into one of two
//               TMEM accumulator stages; tcgen05.commit frees ring slots / publishes the accumulator
//   warps 2-5   epilogue: tcgen05.ld the accumulator 32 columns at a time (one row per lane), transpose the 32 x 32 fp32 chunk through a
//               padded shared-memory buffer so that a lane group owns 128 contiguous bytes of a row, apply the fused epilogue (bias, erf-GELU,
//               fp32 / fp16 residual) there and store coalesced; overlaps the next tile's MMAs.  (Round 1 stored straight from the row-per-lane
//               layout: 32 cache lines per store instruction, 16 k L1 wavefronts per 128 x 256 tile against 6 k cycles of MMAs -- the GEMMs with
//               16-bit outputs ran at the store rate.)  With K = 768 a tile's MMAs take ~7 k cycles and the epilogue of the previous tile has
//               to fit behind them: the epilogue kind is a template parameter (straight-line code), the staging buffer is addressed in the
//               shared window (STS / LDS, not generic stores), the next chunk's tcgen05.ld and the residual operands of the chunk two ahead
//               are in flight while a chunk is processed.
//
// Implicit-GEMM 3x3 convolution for the VQ pixel decoders (tokenizer/vq_modules.py:63-127, 293-415): the same kernel with the A operand
// fetched straight from the channels-last activation [B, H, W, C] through a 4-D tensor map.  An M tile is a box of 128 output pixels
// (box_w x box_h of one image), a K block is 64 channels of one of the nine taps: the producer asks for the box shifted by (dx, dy), the TMA
// unit zero-fills what falls outside the image (the convolution's padding) and writes the 128 rows x 128 bytes in the same swizzled layout a
// row-major [128][64] tile would have -- no im2col matrix ever exists in memory.
#include <cuda.h>

#include "common.cuh"
#include "../../include/umgen.h"

namespace umgen {
namespace gemm {

constexpr int BM = 128, BK = 64, STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;        // 16 KB; the W tile is BN x 64 halves (32 KB at BN = 256)
constexpr int THREADS = 192;
constexpr int EPI_PITCH = 36;

struct Params {
    int M, N, K;
    int epilogue;            // UMGEN_EPI_*
    const float* bias;       // [N] or null
    void* out;               // fp16 [M,N] (EPI 0/1/4) or fp32 [M,N] (EPI 2/3)
    int ldo;                 // row pitch of out in elements
    const __half* resid;     // EPI 4: fp16 residual [M,N]
    int ldr;
    // implicit 3x3 convolution (conv != 0): A is [B, cH, cW, C] fp16 behind a 4-D tensor map, K = 9 * C, M = B * cH * cW
    int conv, cH, cW, cchunks;      // cchunks = C / 64
    int n_out;                      // UMGEN_EPI_NCHW_F32: the first n_out columns go to out[b][column][y][x] (fp32), the others are padding
};

__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(smem)),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                     smem_u32(smem)),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// K-major operand tile with 128-byte swizzle: 8-row atoms of 128 B, 1024 B apart (SBO), descriptor version 1
__device__ __forceinline__ uint64_t umma_desc_sw128(const void* smem) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(smem) >> 4) & 0x3fff);
    d |= (uint64_t)(1024 >> 4) << 32;      // stride byte offset
    d |= (uint64_t)1 << 46;                // version (Blackwell)
    d |= (uint64_t)2 << 61;                // SWIZZLE_128B
    return d;
}
// kind::f16, A/B fp16 K-major, fp32 accumulate, M=128, N=bn
__device__ __forceinline__ uint32_t umma_idesc(int bn) {
    return (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, "
        "%27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, "
        "%27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
// one lane of a converged warp (elect.sync): the form the compiler turns into straight-line UTCHMMA sequences (a `lane == 0` branch gets an
// election loop around every tcgen05.mma)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred)::"memory");
    return pred != 0;
}
// exact-erf GELU (module.py:239) for 16-bit outputs: erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, three orders below the fp16 rounding
// of the result) -- half the instructions of erff, which made the c_fc epilogue (32 k GELUs per tile) slower than the tile's MMAs
__device__ __forceinline__ float gelu_erf_fast(float x) {
    const float z = fabsf(x) * 0.70710678118654752f;
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    p *= t;
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * z * z));
    const float erf_abs = fmaf(-p, e, 1.0f);
    return 0.5f * x + 0.5f * fabsf(x) * erf_abs;          // 0.5 x (1 + sign(x) erf|z|)
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int BN> struct __align__(1024) Smem {
    static constexpr int STAGE_BYTES = A_BYTES + BN * BK * 2;
    uint8_t stage[STAGES][STAGE_BYTES];      // [A 16 KB | W BN x 128 B], every tile 1024-B aligned
    float epi[4][32][EPI_PITCH];             // per epilogue warp: 32 rows x 32 fp32 columns, row pitch 36 floats (conflict-free 16-byte accesses both ways)
    uint64_t full[STAGES], empty[STAGES];
    uint64_t acc_full[2], acc_empty[2];
    uint32_t tmem_base;
};

template <int BN, int EPI> __global__ void __launch_bounds__(THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const __grid_constant__ Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    using SmemT = Smem<BN>;
    constexpr int STAGE_BYTES = SmemT::STAGE_BYTES;
    constexpr uint32_t TMEM_COLS = 2 * BN;      // two accumulator stages of BN fp32 columns
    SmemT* sm = reinterpret_cast<SmemT*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_m = (p.M + BM - 1) / BM, tiles_n = p.N / BN;
    const int n_tiles = tiles_m * tiles_n;
    const int kblocks = p.K / BK;

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&sm->full[i], 1); mbar_init(&sm->empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&sm->acc_full[i], 1); mbar_init(&sm->acc_empty[i], 4); }
        mbar_fence_init();
    }
    if (warp == 1) {       // TMEM allocation by one full warp
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm->tmem_base)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm->tmem_base;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int tm = tile / tiles_n, tn = tile - tm * tiles_n;
                // conv: first pixel of the tile's box (a tile never straddles two image rows partially nor two images: host-checked geometry)
                int cb = 0, cy = 0, cx = 0, tap = 0, cc = 0;
                if (p.conv) {
                    const int px = tm * BM, per_img = p.cH * p.cW;
                    cb = px / per_img;
                    const int r = px - cb * per_img;
                    cy = r / p.cW;
                    cx = r - cy * p.cW;
                }
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(&sm->empty[s], ph ^ 1);
                    mbar_arrive_expect_tx(&sm->full[s], STAGE_BYTES);
                    if (p.conv) {      // K block kb = channels [64 cc, 64 cc + 64) of tap (ky, kx); out-of-image rows / columns arrive as zeros
                        tma_load_4d(sm->stage[s], &map_a, cc * BK, cx + tap % 3 - 1, cy + tap / 3 - 1, cb, &sm->full[s]);
                        if (++cc == p.cchunks) { cc = 0; ++tap; }
                    } else {
                        tma_load_2d(sm->stage[s], &map_a, kb * BK, tm * BM, &sm->full[s]);
                    }
                    tma_load_2d(sm->stage[s] + A_BYTES, &map_w, kb * BK, tn * BN, &sm->full[s]);   // W tile: BN rows
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: the whole warp walks the schedule, one elected lane issues =====================
        const uint32_t idesc = umma_idesc(BN);
        uint32_t it = 0, tcount = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
            const uint32_t as = tcount & 1, aph = (tcount >> 1) & 1;
            mbar_wait(&sm->acc_empty[as], aph ^ 1);       // epilogue drained this accumulator stage
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + as * BN;
            for (int kb = 0; kb < kblocks; ++kb, ++it) {
                const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(&sm->full[s], ph);
                tc_fence_after();
                const uint64_t da = umma_desc_sw128(sm->stage[s]);
                const uint64_t db = umma_desc_sw128(sm->stage[s] + A_BYTES);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)       // +32 bytes (2 x 16-B units) per K=16 step inside the swizzle atom
                        umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                    umma_commit(&sm->empty[s]);             // ring slot reusable once these MMAs retire
                    if (kb + 1 == kblocks) umma_commit(&sm->acc_full[as]);      // accumulator complete
                }
                __syncwarp();
            }
        }
    } else {
        // ===================== epilogue warps (2..5) =====================
        constexpr bool HALF_OUT = EPI == UMGEN_EPI_BIAS_F16 || EPI == UMGEN_EPI_GELU_F16 || EPI == UMGEN_EPI_RESID_F16;
        constexpr bool HAS_RES = EPI == UMGEN_EPI_RESID_F32 || EPI == UMGEN_EPI_RESID_F16;
        constexpr int NCH = BN / 32;
        const int quarter = warp & 3;                       // TMEM lane quarter this warp may access
        const uint32_t stg = smem_u32(&sm->epi[quarter][0][0]);
        const int rsub = lane >> 3, cq = lane & 7;          // coalesced view of a 32 x 32 chunk: pass i covers rows 4i + rsub, lane group cq holds columns 4cq .. 4cq+3
        const uint32_t st_addr = stg + (uint32_t)lane * (EPI_PITCH * 4), ld_addr = stg + (uint32_t)rsub * (EPI_PITCH * 4) + (uint32_t)cq * 16;
        uint32_t tcount = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
            const int tm = tile / tiles_n, tn = tile - tm * tiles_n;
            const uint32_t as = tcount & 1, aph = (tcount >> 1) & 1;
            const int row0 = tm * BM + quarter * 32 + rsub;             // my row of pass 0
            const int col0 = tn * BN + 4 * cq;                          // my first column of chunk 0
            const int rows_left = p.M - row0;                           // pass i is inside the matrix when 4 i < rows_left
            // Residual operand of chunk cb, row pass i.  The epilogue is a stream of 16-byte loads whose latency nothing else hides: every operand is
            // re-loaded for the chunk two ahead the moment it has been consumed, so two chunks' loads stay in flight behind the work.
            auto load_res = [&](int cb, int i) -> float4 {
                if (!HAS_RES || cb >= NCH || 4 * i >= rows_left) return make_float4(0.f, 0.f, 0.f, 0.f);
                if (EPI == UMGEN_EPI_RESID_F32) return *reinterpret_cast<const float4*>((const float*)p.out + (size_t)(row0 + 4 * i) * p.ldo + col0 + cb * 32);
                const uint2 rv = *reinterpret_cast<const uint2*>(p.resid + (size_t)(row0 + 4 * i) * p.ldr + col0 + cb * 32);
                const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&rv.x)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&rv.y));
                return make_float4(f0.x, f0.y, f1.x, f1.y);
            };
            float4 res_a[HAS_RES ? 8 : 1], res_b[HAS_RES ? 8 : 1];
            if (HAS_RES) {       // the residual does not depend on the accumulator: on its way before the MMAs of this tile have finished
#pragma unroll
                for (int i = 0; i < (HAS_RES ? 8 : 1); ++i) { res_a[i] = load_res(0, i); res_b[i] = load_res(1, i); }
            }
            mbar_wait(&sm->acc_full[as], aph);
            tc_fence_after();
            const uint32_t tacc = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * BN;
            // r: row-per-lane accumulator chunk; processed = transposed through shared memory, epilogue applied, stored
            auto process = [&](int cb, const uint32_t (&r)[32], float4 (&res)[HAS_RES ? 8 : 1]) {
                float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + cb * 32));
#pragma unroll
                for (int v = 0; v < 8; ++v) sts128(st_addr + 16 * v, r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]);
                __syncwarp();
                float4 acc[8];         // all eight row passes first: 32 independent values for the math below, and the buffer is free again at once
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = lds128(ld_addr + (uint32_t)i * (4 * EPI_PITCH * 4));
                __syncwarp();          // the chunk has left the buffer before the next one overwrites it
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    acc[i].x += b4.x; acc[i].y += b4.y; acc[i].z += b4.z; acc[i].w += b4.w;
                    if (HAS_RES) {
                        acc[i].x += res[i].x; acc[i].y += res[i].y; acc[i].z += res[i].z; acc[i].w += res[i].w;
                        res[i] = load_res(cb + 2, i);          // (columns of another chunk: the stores below do not touch them)
                    }
                    if (EPI == UMGEN_EPI_GELU_F16) {
                        acc[i].x = gelu_erf_fast(acc[i].x); acc[i].y = gelu_erf_fast(acc[i].y); acc[i].z = gelu_erf_fast(acc[i].z); acc[i].w = gelu_erf_fast(acc[i].w);
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (4 * i < rows_left) {
                        const size_t row = (size_t)(row0 + 4 * i);
                        const int col = col0 + cb * 32;
                        if (HALF_OUT) {
                            const __half2 h0 = __floats2half2_rn(acc[i].x, acc[i].y), h1 = __floats2half2_rn(acc[i].z, acc[i].w);
                            *reinterpret_cast<uint2*>((__half*)p.out + row * p.ldo + col) =
                                make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
                        } else if (EPI == UMGEN_EPI_NCHW_F32) {      // conv_out: a handful of planes, [B][n_out][H][W] fp32
                            const int hw = p.cH * p.cW, bi = (int)row / hw, rem = (int)row - bi * hw;
                            float* o = (float*)p.out + (size_t)bi * p.n_out * hw + rem;
                            if (col < p.n_out) o[(size_t)col * hw] = acc[i].x;
                            if (col + 1 < p.n_out) o[(size_t)(col + 1) * hw] = acc[i].y;
                            if (col + 2 < p.n_out) o[(size_t)(col + 2) * hw] = acc[i].z;
                            if (col + 3 < p.n_out) o[(size_t)(col + 3) * hw] = acc[i].w;
                        } else {
                            *reinterpret_cast<float4*>((float*)p.out + row * p.ldo + col) = acc[i];
                        }
                    }
                }
            };
            // conv_out: only the chunks that hold real output columns (the others are zero padding of the weight matrix)
            const int nch = EPI == UMGEN_EPI_NCHW_F32 ? (p.n_out + 31) / 32 : NCH;
            uint32_t ra[32], rb[32];
            tmem_ld32_nowait(tacc, ra);
#pragma unroll 1
            for (int cb = 0; cb < nch; cb += 2) {
                tmem_ld_wait();
                if (cb + 1 < nch) tmem_ld32_nowait(tacc + (cb + 1) * 32, rb);
                process(cb, ra, res_a);
                if (cb + 1 < nch) {
                    tmem_ld_wait();
                    if (cb + 2 < nch) tmem_ld32_nowait(tacc + (cb + 2) * 32, ra);
                    process(cb + 1, rb, res_b);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm->acc_empty[as]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---- host ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
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
// row-major fp16 [rows][cols] with row pitch ld elements; box = [box_rows][64 cols], 128-byte swizzle
static int make_map(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu", (int)r, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld); return -2; }
    return 0;
}

// cuTensorMapEncodeTiled costs microseconds of host time and a frame issues ~3000 GEMMs over ~1500 distinct (pointer, shape) pairs (every weight matrix
// and a handful of activation buffers): a small hash table of encoded maps.  The caller gets a COPY: a later insertion may reuse the slot.
// channels-last fp16 activation [B][H][W][C] as a 4-D tensor (innermost first: C, W, H, B); box = 64 channels x box_w x box_h pixels of one
// image, 128-byte swizzle: in shared memory the box is 128 rows of 128 bytes, the layout of a [128][64] operand tile
static int make_map_conv(CUtensorMap* map, const void* ptr, uint64_t B, uint64_t H, uint64_t W, uint64_t Cc, uint32_t box_w, uint32_t box_h) {
    cuuint64_t dims[4] = {Cc, W, H, B};
    cuuint64_t strides[3] = {Cc * 2, W * Cc * 2, H * W * Cc * 2};
    cuuint32_t box[4] = {BK, box_w, box_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (4-D) failed (%d) B=%llu H=%llu W=%llu C=%llu box %u x %u", (int)r, (unsigned long long)B, (unsigned long long)H, (unsigned long long)W, (unsigned long long)Cc, box_w, box_h); return -2; }
    return 0;
}

// 2-D maps: (rows, cols, ld, box_rows), img_h = img_w = 0.  4-D conv maps: rows = B, cols = C, ld unused, box_rows = box_w, (img_h, img_w) = (H, W).
struct MapKey { const void* ptr; uint64_t rows, cols, ld; uint32_t box_rows; uint32_t img_h, img_w; };
struct MapSlot { MapKey key; CUtensorMap map; bool used; };
constexpr int MAP_SLOTS = 8192, MAP_PROBES = 8;
static MapSlot g_maps[MAP_SLOTS];
static int cached_map(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t img_h = 0, uint32_t img_w = 0) {
    uint64_t h = (uint64_t)(uintptr_t)ptr * 0x9E3779B97F4A7C15ull ^ (rows * 0xBF58476D1CE4E5B9ull) ^ (cols << 20) ^ (ld << 40) ^ box_rows ^ ((uint64_t)img_h << 13) ^ ((uint64_t)img_w << 27);
    h ^= h >> 29;
    const int base = (int)(h & (MAP_SLOTS - 1));
    for (int pr = 0; pr < MAP_PROBES; ++pr) {
        MapSlot& s = g_maps[(base + pr) & (MAP_SLOTS - 1)];
        if (s.used && s.key.ptr == ptr && s.key.rows == rows && s.key.cols == cols && s.key.ld == ld && s.key.box_rows == box_rows && s.key.img_h == img_h &&
            s.key.img_w == img_w) { *out = s.map; return 0; }
    }
    int victim = base;
    for (int pr = 0; pr < MAP_PROBES; ++pr)
        if (!g_maps[(base + pr) & (MAP_SLOTS - 1)].used) { victim = (base + pr) & (MAP_SLOTS - 1); break; }
    MapSlot& s = g_maps[victim];
    const int rc = img_w ? make_map_conv(&s.map, ptr, rows, img_h, img_w, cols, box_rows, BM / box_rows) : make_map(&s.map, ptr, rows, cols, ld, box_rows);
    if (rc) { s.used = false; return rc; }
    s.key = MapKey{ptr, rows, cols, ld, box_rows, img_h, img_w};
    s.used = true;
    *out = s.map;
    return 0;
}

}  // namespace gemm
namespace { int g_sms = 0; int g_sm_limit = 0; }
// CTAs the persistent GEMM may launch (0 = one per SM): the engine lowers it while the decode kernel holds 64 SMs beside a TAR pass
extern "C" int umgen_gemm_set_sm_limit(int n) { g_sm_limit = n > 0 ? n : 0; return 0; }
extern int64_t g_launches;
}  // namespace umgen

using namespace umgen;

template <int BN, int EPI> static int launch_gemm_epi(const CUtensorMap& ma, const CUtensorMap& mw, const umgen::gemm::Params& p, int sms, cudaStream_t st) {
    using namespace umgen::gemm;
    static bool configured = false;
    constexpr int smem = (int)sizeof(Smem<BN>) + 1024;
    if (!configured) {
        UMGEN_CUDA_OK(cudaFuncSetAttribute(gemm_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    const int n_tiles = ((p.M + BM - 1) / BM) * (p.N / BN);
    const int grid = n_tiles < sms ? n_tiles : sms;
    gemm_kernel<BN, EPI><<<grid, THREADS, smem, st>>>(ma, mw, p);
    UMGEN_CUDA_OK(cudaGetLastError());
    return 0;
}
template <int BN> static int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mw, const umgen::gemm::Params& p, int sms, cudaStream_t st) {
    switch (p.epilogue) {
        case UMGEN_EPI_BIAS_F16: return launch_gemm_epi<BN, UMGEN_EPI_BIAS_F16>(ma, mw, p, sms, st);
        case UMGEN_EPI_GELU_F16: return launch_gemm_epi<BN, UMGEN_EPI_GELU_F16>(ma, mw, p, sms, st);
        case UMGEN_EPI_RESID_F32: return launch_gemm_epi<BN, UMGEN_EPI_RESID_F32>(ma, mw, p, sms, st);
        case UMGEN_EPI_STORE_F32: return launch_gemm_epi<BN, UMGEN_EPI_STORE_F32>(ma, mw, p, sms, st);
        case UMGEN_EPI_RESID_F16: return launch_gemm_epi<BN, UMGEN_EPI_RESID_F16>(ma, mw, p, sms, st);
        case UMGEN_EPI_NCHW_F32: return launch_gemm_epi<BN, UMGEN_EPI_NCHW_F32>(ma, mw, p, sms, st);
    }
    umgen::set_error("gemm: bad epilogue %d", p.epilogue);
    return -1;
}

extern "C" int umgen_gemm_f16_ex(const void* a_h, int64_t lda, const void* w_h, const void* bias_f, void* out, int64_t ldo,
                                 const void* resid_h, int64_t ldr, int64_t M, int64_t N, int64_t K, int epilogue, void* stream_v) {
    using namespace umgen::gemm;
    if (!a_h || !w_h || !out) { set_error("gemm: null buffer"); return -1; }
    if (M < 1 || N % 128 != 0 || K % BK != 0 || K < BK) { set_error("gemm: need N %% 128 == 0 and K %% 64 == 0 (M=%lld N=%lld K=%lld)", (long long)M, (long long)N, (long long)K); return -1; }
    if (epilogue < 0 || epilogue > 4) { set_error("gemm: bad epilogue %d", epilogue); return -1; }
    if (epilogue == UMGEN_EPI_RESID_F16 && (!resid_h || ldr % 8 != 0)) { set_error("gemm: fp16 residual epilogue needs resid_h with a pitch multiple of 8"); return -1; }
    if (lda % 8 != 0 || ldo % 8 != 0) { set_error("gemm: row pitches must be multiples of 8 elements"); return -1; }
    if (((uintptr_t)out & 15) != 0 || (bias_f && ((uintptr_t)bias_f & 15) != 0)) { set_error("gemm: out and bias must be 16-byte aligned"); return -1; }
    if (int rc = get_encode()) return rc;
    const int bn = (N % 256 == 0) ? 256 : 128;
    CUtensorMap ma, mw;
    if (int rc = cached_map(&ma, a_h, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM)) return rc;
    if (int rc = cached_map(&mw, w_h, (uint64_t)N, (uint64_t)K, (uint64_t)K, bn)) return rc;
    if (g_sms == 0) {
        int dev = 0;
        UMGEN_CUDA_OK(cudaGetDevice(&dev));
        UMGEN_CUDA_OK(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    Params p;
    p.M = (int)M; p.N = (int)N; p.K = (int)K; p.epilogue = epilogue; p.bias = (const float*)bias_f; p.out = out; p.ldo = (int)ldo;
    p.resid = (const __half*)resid_h; p.ldr = (int)ldr;
    p.conv = 0; p.cH = p.cW = p.cchunks = 0; p.n_out = 0;
    const int rc = bn == 256 ? launch_gemm<256>(ma, mw, p, (g_sm_limit && g_sm_limit < g_sms) ? g_sm_limit : g_sms, (cudaStream_t)stream_v)
                             : launch_gemm<128>(ma, mw, p, (g_sm_limit && g_sm_limit < g_sms) ? g_sm_limit : g_sms, (cudaStream_t)stream_v);
    if (rc) return rc;
    g_launches += 1;
    return 0;
}

// out[(b,y,x), :] = epilogue(sum over the 3x3 taps and channels of x[b, y+ky-1, x+kx-1, c] * w[:, (ky,kx,c)]) -- nn.Conv2d(Cin, Cout, 3, 1, 1) on
// channels-last fp16 (vq_modules.py:63-66, 98-107, 322-326), zero padding by the TMA unit's out-of-bounds fill.
static int conv3x3_launch(const void* x_h, int64_t B, int64_t H, int64_t W, int64_t Cin, const void* w_h, const void* bias_f, void* out,
                          const void* resid_h, int64_t Cout, int epilogue, int64_t n_out, void* stream_v) {
    using namespace umgen::gemm;
    if (!x_h || !w_h || !out) { set_error("conv3x3: null buffer"); return -1; }
    if (epilogue == UMGEN_EPI_RESID_F16 && !resid_h) { set_error("conv3x3: residual epilogue without resid_h"); return -1; }
    if (B < 1 || H < 1 || W < 8 || Cin < BK || Cin % BK != 0 || Cout % 128 != 0 || Cout < 128) { set_error("conv3x3: need Cin %% 64 == 0, Cout %% 128 == 0, W >= 8 (B=%lld H=%lld W=%lld Cin=%lld Cout=%lld)", (long long)B, (long long)H, (long long)W, (long long)Cin, (long long)Cout); return -1; }
    // an M tile is box_w x box_h pixels of one image: whole rows when W <= 128, a 128-pixel run of one row otherwise
    const int64_t box_w = W >= BM ? BM : W;
    if (BM % box_w != 0 || W % box_w != 0 || H % (BM / box_w) != 0) { set_error("conv3x3: image %lld x %lld does not tile into boxes of 128 pixels", (long long)H, (long long)W); return -1; }
    if (B * H * W > 0x7fffffffll || B * H * W * Cout > 0x7fffffffffll) { set_error("conv3x3: too many pixels"); return -1; }
    if ((((uintptr_t)x_h | (uintptr_t)out | (uintptr_t)resid_h) & 15) != 0 || (bias_f && ((uintptr_t)bias_f & 15) != 0)) { set_error("conv3x3: buffers must be 16-byte aligned"); return -1; }
    if (int rc = get_encode()) return rc;
    const int bn = (Cout % 256 == 0) ? 256 : 128;
    const int64_t K = 9 * Cin;
    CUtensorMap ma, mw;
    if (int rc = cached_map(&ma, x_h, (uint64_t)B, (uint64_t)Cin, 0, (uint32_t)box_w, (uint32_t)H, (uint32_t)W)) return rc;
    if (int rc = cached_map(&mw, w_h, (uint64_t)Cout, (uint64_t)K, (uint64_t)K, bn)) return rc;
    if (g_sms == 0) {
        int dev = 0;
        UMGEN_CUDA_OK(cudaGetDevice(&dev));
        UMGEN_CUDA_OK(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    Params p;
    p.M = (int)(B * H * W); p.N = (int)Cout; p.K = (int)K; p.epilogue = epilogue; p.bias = (const float*)bias_f; p.out = out; p.ldo = (int)Cout;
    p.resid = (const __half*)resid_h; p.ldr = (int)Cout;
    p.conv = 1; p.cH = (int)H; p.cW = (int)W; p.cchunks = (int)(Cin / BK); p.n_out = (int)n_out;
    const int sms = (g_sm_limit && g_sm_limit < g_sms) ? g_sm_limit : g_sms;
    const int rc = bn == 256 ? launch_gemm<256>(ma, mw, p, sms, (cudaStream_t)stream_v) : launch_gemm<128>(ma, mw, p, sms, (cudaStream_t)stream_v);
    if (rc) return rc;
    g_launches += 1;
    return 0;
}
extern "C" int umgen_conv3x3_f16(const void* x_h, int64_t B, int64_t H, int64_t W, int64_t Cin, const void* w_h, const void* bias_f, void* out_h,
                                 const void* resid_h, int64_t Cout, int epilogue, void* stream_v) {
    if (epilogue != UMGEN_EPI_BIAS_F16 && epilogue != UMGEN_EPI_RESID_F16) { set_error("conv3x3: epilogue must be BIAS_F16 or RESID_F16 (got %d)", epilogue); return -1; }
    return conv3x3_launch(x_h, B, H, W, Cin, w_h, bias_f, out_h, resid_h, Cout, epilogue, 0, stream_v);
}
// conv_out (vq_modules.py:330-334, 413-414): a 3x3 convolution to n_out <= 32 planes.  The weight matrix is zero-padded to 128 rows so the same tiles
// apply (the decoders' last activation is fetched nine times whatever the width of the MMA: the padding costs no time); the epilogue writes the first
// n_out columns as fp32 planes [B][n_out][H][W].
extern "C" int umgen_conv3x3_nchw_f32(const void* x_h, int64_t B, int64_t H, int64_t W, int64_t Cin, const void* w_h, const void* bias_f, void* out_f,
                                      int64_t n_out, void* stream_v) {
    if (n_out < 1 || n_out > 32) { set_error("conv3x3_nchw: 1 <= n_out <= 32 (got %lld)", (long long)n_out); return -1; }
    return conv3x3_launch(x_h, B, H, W, Cin, w_h, bias_f, out_f, nullptr, 128, UMGEN_EPI_NCHW_F32, n_out, stream_v);
}

extern "C" int umgen_gemm_f16(const void* a_h, int64_t lda, const void* w_h, const void* bias_f, void* out, int64_t ldo, int64_t M,
                              int64_t N, int64_t K, int epilogue, void* stream_v) {
    return umgen_gemm_f16_ex(a_h, lda, w_h, bias_f, out, ldo, nullptr, 0, M, N, K, epilogue, stream_v);
}

// Lazy module loading (the CUDA 12 default) loads a kernel on its first launch and that load waits for an idle device -- which never comes while the
// persistent decode kernel spins on a flag.  umgen_preload() (capi.cu) forces every kernel of the library to load up front.
#define UMGEN_PRELOAD(k) UMGEN_CUDA_OK(cudaFuncGetAttributes(&fa_, k))
namespace umgen {
int preload_gemm() {
    cudaFuncAttributes fa_;
    UMGEN_PRELOAD((gemm::gemm_kernel<256, UMGEN_EPI_BIAS_F16>)); UMGEN_PRELOAD((gemm::gemm_kernel<256, UMGEN_EPI_GELU_F16>));
    UMGEN_PRELOAD((gemm::gemm_kernel<256, UMGEN_EPI_RESID_F32>)); UMGEN_PRELOAD((gemm::gemm_kernel<256, UMGEN_EPI_STORE_F32>));
    UMGEN_PRELOAD((gemm::gemm_kernel<256, UMGEN_EPI_RESID_F16>)); UMGEN_PRELOAD((gemm::gemm_kernel<256, UMGEN_EPI_NCHW_F32>));
    UMGEN_PRELOAD((gemm::gemm_kernel<128, UMGEN_EPI_BIAS_F16>)); UMGEN_PRELOAD((gemm::gemm_kernel<128, UMGEN_EPI_GELU_F16>));
    UMGEN_PRELOAD((gemm::gemm_kernel<128, UMGEN_EPI_RESID_F32>)); UMGEN_PRELOAD((gemm::gemm_kernel<128, UMGEN_EPI_STORE_F32>));
    UMGEN_PRELOAD((gemm::gemm_kernel<128, UMGEN_EPI_RESID_F16>)); UMGEN_PRELOAD((gemm::gemm_kernel<128, UMGEN_EPI_NCHW_F32>));
    return 0;
}
}  // namespace umgen
