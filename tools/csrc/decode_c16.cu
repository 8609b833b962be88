// One-cluster OAR decode kernel: ONE thread-block cluster of 16 CTAs runs every single-token step of a frame.  Same contract as decode.cu /
// decode_cluster.cu (reference models/UMGen.py:1151-1383, models/module.py:378-428).  Parity-tested like the other two kernels; NOT the
// default: on the B200s of this pool it measures 721 us/step against 594 us/step for the 64-CTA cluster kernel (DESIGN.md section 3.4 has the
// breakdown).  umgen_decode_frame picks it when the device cannot keep the 8 clusters of the cluster kernel resident, or with mode = 3.
//
// Idea.  The 64-CTA cluster kernel spends most of a layer on its two L2 hops and the skew of 64 CTAs they absorb, while the weights it streams
// need only ~13 GB/s per SM.  The streaming microbenchmark (tools/bench_stream.py, profiles/r1_stream_microbench.txt) shows that a single SM
// pulls 78-106 bytes per clock (150-200 GB/s) from HBM with 36 KB bulk copies, so 16 SMs alone sustain 2.4-2.7 TB/s -- and 16 CTAs fit in one
// cluster, where every exchange is a distributed-shared-memory store instead of an L2 round trip.
//
// Partitioning (CTA r of the cluster: 12 consumer warps, 1 producer warp, sender warps):
//   * attention head r lives entirely in CTA r: its 144 c_attn rows (q | k | v), its KV cache (16-key fragment tiles, K tile = [keys][dims],
//     V tile = [dims][keys]) and the softmax; nothing about attention crosses CTAs.
//   * c_proj is split along K by head: CTA r multiplies its head's 48 outputs into all 768 rows; the 16 partial vectors are reduce-scattered
//     over DSMEM (rank s sums rows [48 s, +48) in a fixed butterfly order, adds bias) and the sums all-gathered: 2 DSMEM hops.
//   * c_fc is split by rows (192 per CTA, GELU output stays in the CTA), the MLP c_proj along K (768 x 192 per CTA) with the same
//     reduce-scatter / all-gather: 4 DSMEM hops per layer, no L2 hop, no device-wide barrier.
//   * the residual vector lives in registers (thread t holds elements 2t, 2t+1; every CTA holds bit-identical values).
// Streaming.  Everything a CTA reads from HBM -- its 884 736 weight bytes per layer, its head's cache tiles, its 1/16 of the head matrix --
// arrives through a ring of four 36 KB slots fed by one producer thread with 1-D bulk copies (cp.async.bulk + mbarrier complete_tx).  Every
// stage is laid out as [warp 12][6 fragment blocks of 512 B] (weights) or 24 tiles / 24 rows (cache, head) so that all 12 consumer warps
// work on every stage as soon as it lands and release it with one mbarrier arrive each: compute follows the arrival of the bytes, the tail
// after the last byte of a matrix is 6 MMAs.
// What limits it (measured, cycle probes and per-warp timelines at step 1200): per SM the stream runs at ~64 B/clk inside the kernel (4 slots
// in flight against ~1800 cycles of loaded HBM latency; L2 prefetch does not help -- L2 hits are capped at 64 B/clk per SM), so the 885 KB of a
// layer alone take 14 k cycles; attention of a whole head on one SM is bound by the instruction issue of its 12 warps (8 k cycles at 1200 keys);
// and the four exchanges cost ~2 k cycles each although a DSMEM all-to-all takes 960 cycles in isolation (tools/bench_dsmem.py).
#include <stdlib.h>

#define UMGEN_CONS_WARPS 12
#include "../../umgen_b200/csrc/decode_shared.cuh"

namespace umgen {
namespace c16 {

constexpr int CL = 16;                       // CTAs in the cluster = attention heads
static_assert(CL == NH, "one head per CTA");
constexpr int XS = C / CL;                   // 48 residual rows summed by rank s
constexpr int LINES = XS / 2;                // 24 two-value lines per (rank, slice)
constexpr int QKV_R = 3 * HD;                // 144 c_attn rows per CTA: q | k | v of its head
constexpr int FC_R = FF / CL;                // 192 c_fc rows per CTA
constexpr uint32_t SLOT_BYTES = 72 * 512;    // 36 864: one stage = 72 A-fragment blocks = 24 cache tiles = 24 head rows
constexpr int NSLOT = 4;
constexpr int ST_QKV = 6, ST_PROJ = 2, ST_FC = 8, ST_PROJ2 = 8;
constexpr int ST_LAYER = ST_QKV + ST_PROJ + ST_FC + ST_PROJ2;       // 24 weight stages per layer
constexpr uint32_t CTA_LAYER_BYTES = ST_LAYER * SLOT_BYTES;        // 884 736
static_assert((size_t)CTA_LAYER_BYTES * CL == (size_t)UMGEN_OAR_LAYER_H * 2, "c16 packing");
constexpr uint32_t KV_TILE_BYTES = 16 * HD * 2;                    // 1536: 3 fragment blocks
constexpr int KV_TILES = UMGEN_KV_ROWS / 16;                       // 144 tiles per (layer, k|v, head)
constexpr int TPS = 24;                                            // cache tiles per stage
constexpr int MAX_KV_STAGES = 6;                                   // 138 tiles cover 2207 keys
static_assert(TPS * KV_TILE_BYTES == SLOT_BYTES && MAX_KV_STAGES * TPS * 16 >= SEQ && MAX_KV_STAGES * TPS <= KV_TILES, "cache staging");
constexpr int HEAD_ROWS = 24;
static_assert(HEAD_ROWS * C * 2 == SLOT_BYTES, "head staging");
constexpr int PART_STRIDE = 52;              // (m, l, o[48]) + pad
constexpr uint64_t TIMEOUT_NS = 10ull * 1000 * 1000 * 1000;
#ifndef UMGEN_C16_SENDERS
#define UMGEN_C16_SENDERS 2
#endif
constexpr int N_SEND_WARPS = UMGEN_C16_SENDERS;                // the 12 remote-store instructions of an exchange are spread over these warps (1: 765, 2: 721, 3: 731 us/step)
constexpr int NT = N_CONS + 32 + 32 * N_SEND_WARPS;      // 12 consumer warps + producer warp + sender warps = 16 warps
constexpr int SC_LOGIT = 0;                  // global scratch (floats): [8192 lines {value, tag}] AR logits (top-p mode only)
constexpr int SC_TOTAL = 2 * 8192;

// per-layer fp32 parameters staged in shared memory one layer ahead (cp.async): ln_1 | ln_2 | c_proj bias rows [48 r, +48) | c_attn bias of head r (q | k | v)
constexpr int PRM_LN1 = 0, PRM_LN2 = C, PRM_BPROJ = 2 * C, PRM_BQKV = 2 * C + XS, PRM_FLOATS = 2 * C + XS + QKV_R;
// offsets inside one layer of oar_f: ln_1[768] | c_attn.bias[2304] | c_proj.bias[768] | ln_2[768]
constexpr int F_LN1 = 0, F_BQKV = C, F_BPROJ = C + 3 * C, F_LN2 = C + 3 * C + C, LAYER_F = UMGEN_OAR_LAYER_F;

struct KParams {
    UmgenDecodeArgs a;
    const void* oar_c16_h;     // [n_layer][UMGEN_OAR_LAYER_H] fp16 in this kernel's stage order (umgen_tools_pack_oar_c16)
};

struct __align__(128) Smem {
    uint8_t ring[NSLOT * SLOT_BYTES];
    float xn[C];                     // normalised vector feeding the head GEMV
    uint2 xf[C / 16][8];             // normalised vector as mma B fragments: [k-step][lane 0..3 hi, 4..7 lo] = {b0, b1}
    uint2 yf[HD / 16][8];            // attention output of my head, same form
    uint2 hf[FC_R / 16][8];          // my slice of the MLP hidden vector, same form
    float lno[C];                    // ln_oar weight
    float prm[2][PRM_FLOATS];        // layer parameters, double buffered
    // exchange targets, written remotely as self-flagged 16-byte lines {v0, tag, v1, tag}
    uint4 rsl[CL][LINES];            // partial sums of my 48 rows from the 16 ranks (reduce-scatter)
    uint4 xl[CL][LINES];             // summed rows [48 r, +48) from rank r (all-gather)
    uint4 candl[CL][MAX_CAND];       // top-k candidates {value, tag, id, tag} of the 16 ranks
    float2 sums[LINES];              // my 48 summed rows between the reduce-scatter and the all-gather
    float out2[C];                   // my K-slice of a c_proj output before the reduce-scatter
    float pq[4][QKV_R];              // K-quarter partials of the c_attn rows
    float qv[HD];                    // q of my head (fp32)
    __half knew[HD];                 // k, v of this step at cache precision
    __half vnew[HD];
    float stage[1040];               // candidates (values | ids) / TAR-head row scratch (>= 1028)
    float acc[8192 / CL + 8];        // head logits of my slice
    float wpart[N_CONS_WARPS][PART_STRIDE];
    float red[64];
    float corners[MAX_BOX][8];
    int box_dropped[MAX_BOX];
    int recent[16];
    uint64_t full[NSLOT];
    uint64_t empty[NSLOT];
    volatile uint32_t kv_progress;   // layers (step * L + layer + 1) whose cache row is written and fenced
    uint64_t hs_rs, hs_ag;           // mbarriers (12 arrivals): the consumer warps handed their part of out2 / sums to the sender warp
    volatile int tok;
    int nbox;
};
static_assert(sizeof(Smem) + 128 <= 227 * 1024, "shared memory budget");
static_assert(2 * CL * MAX_CAND <= 1040, "stage holds the merged candidates");

extern __shared__ __align__(128) uint8_t smem_raw_c16[];
__device__ __forceinline__ Smem* SM() { return reinterpret_cast<Smem*>(smem_raw_c16); }

struct Ctx {
    const KParams* p;
    int* abort_flag;
    int* probe;
    long long probe_t0;
    int tid, warp, lane;
    int i;                    // rank in the cluster = my head
    uint32_t k;               // stages consumed (consumers) / issued (producer) so far
    bool rdy;                 // consumers: stage k is already known to have landed (stage_peek_next)
    uint32_t lc;              // layers completed so far
    uint32_t sbase, rbase, rstride;   // my shared window, rank 0's window in the cluster address space, window stride per rank
    uint64_t t_dead;
    bool dbg_local;           // debug (args.grid bit 1): no exchange waits -> wrong results, isolates the local work
    long long* tl;            // timeline row of this warp (UMGEN_DECODE_PROFILE == 3), lane 0 only
    long long tl_base;
};
#ifndef UMGEN_DECODE_PROFILE
#define UMGEN_DECODE_PROFILE 0
#endif
#ifndef UMGEN_C16_POLL_SLEEP
#define UMGEN_C16_POLL_SLEEP 0
#endif
#ifndef UMGEN_C16_PRM_LATE
#define UMGEN_C16_PRM_LATE 0
#endif
#ifndef UMGEN_PROBE_STEP
#define UMGEN_PROBE_STEP 1200
#endif
#ifndef UMGEN_PROBE_TID
#define UMGEN_PROBE_TID 0
#endif
#if UMGEN_DECODE_PROFILE == 1
#define PROBE(n) if (c.probe) { c.probe[n] = (int)(clock64() - c.probe_t0); }
#else
#define PROBE(n)
#endif
#if UMGEN_DECODE_PROFILE == 3       // timeline: every consumer warp of every CTA stamps its clock (relative to the cluster-wide start barrier)
#define STAMP(n) if (c.tl) { c.tl[n] = clock64() - c.tl_base; }
#else
#define STAMP(n)
#endif
#if UMGEN_DECODE_PROFILE == 2       // probes inside attention() instead of the layer-level ones
#define PROBE2(n) if (c.probe) { c.probe[n] = (int)(clock64() - c.probe_t0); }
#else
#define PROBE2(n)
#endif

__device__ __noinline__ bool check_abort_slow(int* abort_flag, uint64_t t_dead, uint32_t code) {
    if (*(volatile int*)abort_flag != 0) return true;
    if (globaltimer_ns() > t_dead) {
        atomicCAS(abort_flag, 0, 300 + (int)(code & 0xffff));       // a wait this long is a deadlock: every CTA drains
        return true;
    }
    return false;
}
__device__ __forceinline__ bool check_abort(Ctx& c, uint32_t& spins) {
    ++spins;
    if ((spins & 0x3ffu) == 0) return check_abort_slow(c.abort_flag, c.t_dead, c.lc);
    return false;
}
__device__ __forceinline__ void wait_mbar(Ctx& c, uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (check_abort(c, spins)) return;
    }
}

// ---- ring: fixed slots, stage k lives in slot k % 4; every consumer warp arrives once on the slot's empty barrier ------------------
__device__ __forceinline__ const uint8_t* stage_wait(Ctx& c) {
    const uint32_t slot = c.k & (NSLOT - 1);
    if (!c.rdy) wait_mbar(c, &SM()->full[slot], (c.k / NSLOT) & 1u);
    c.rdy = false;
    return SM()->ring + slot * SLOT_BYTES;
}
// try_wait of the NEXT stage, issued while the current one is being worked on so that its ~100 cycles of latency overlap with the MMAs.  A true
// answer stays true (the slot cannot be refilled before this warp releases it); a false one only means stage_wait will poll.
__device__ __forceinline__ void stage_peek_next(Ctx& c) {
    const uint32_t k = c.k + 1;
    c.rdy = mbar_try_wait(&SM()->full[k & (NSLOT - 1)], (k / NSLOT) & 1u);
}
__device__ __forceinline__ void stage_done(Ctx& c) {       // this warp no longer reads the stage
    __syncwarp();
    if (c.lane == 0) mbar_arrive(&SM()->empty[c.k & (NSLOT - 1)]);
    c.k++;
}
__device__ __forceinline__ void produce(Ctx& c, const void* src, uint32_t bytes) {
    Smem* sm = SM();
    const uint32_t slot = c.k & (NSLOT - 1), use = c.k / NSLOT;
    if (use > 0) wait_mbar(c, &sm->empty[slot], (use - 1) & 1u);
    mbar_arrive_expect_tx(&sm->full[slot], bytes);
    bulk_g2s(sm->ring + slot * SLOT_BYTES, src, bytes, &sm->full[slot]);
    c.k++;
}

// ---- DSMEM lines (see decode_cluster.cu: plain remote 16-byte stores, local polls, each 8-byte half carries its own tag) -------------
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ uint32_t remote(const Ctx& c, const void* p, uint32_t rank) { return c.rbase + rank * c.rstride + (smem_u32(p) - c.sbase); }
__device__ __forceinline__ void send_line(const Ctx& c, uint4* dst, uint32_t rank, float v0, float v1, uint32_t tag) {
    asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %2};" ::"r"(remote(c, dst, rank)), "r"(__float_as_uint(v0)), "r"(tag), "r"(__float_as_uint(v1))
                 : "memory");
}
__device__ __forceinline__ float2 wait_line(Ctx& c, const uint4* line, uint32_t tag) {
    const uint32_t addr = smem_u32(line);
    uint32_t spins = 0;
    uint4 r;
    while (true) {
        asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr) : "memory");
        if ((r.y == tag && r.w == tag) || c.dbg_local) break;
        if (check_abort(c, spins)) break;
#if UMGEN_C16_POLL_SLEEP
        __nanosleep(UMGEN_C16_POLL_SLEEP);
#endif
    }
    return make_float2(__uint_as_float(r.x), __uint_as_float(r.z));
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint4 ll_ld(const float* p) {
    uint4 r;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}

// ---- cp.async staging of the small per-layer parameters ---------------------------------------------
__device__ __forceinline__ void cp_async16(void* s, const void* g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(s)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_params(Ctx& c, const float* fl, float* dst) {
    if (c.tid < 192) cp_async16(dst + PRM_LN1 + 4 * c.tid, fl + F_LN1 + 4 * c.tid);
    else cp_async16(dst + PRM_LN2 + 4 * (c.tid - 192), fl + F_LN2 + 4 * (c.tid - 192));
    if (c.tid < XS / 4) cp_async16(dst + PRM_BPROJ + 4 * c.tid, fl + F_BPROJ + XS * c.i + 4 * c.tid);
    if (c.tid >= 32 && c.tid < 32 + QKV_R / 4) {      // q | k | v bias of head i: three runs of 48
        const int u = c.tid - 32, which = u / (HD / 4), e = u - which * (HD / 4);
        cp_async16(dst + PRM_BQKV + which * HD + 4 * e, fl + F_BQKV + which * C + c.i * HD + 4 * e);
    }
    cp_async_commit();
}

// ---- tensor-core GEMV pieces (same conventions as decode_cluster.cu) ------------------------------------
// D[16x8] += A[16x16] . B[16x8]: A = one packed 512-byte fragment block (lane l reads bytes [16 l, 16 l + 16)),
// B columns 0 / 1 = the fp16 hi / lo parts of the fp32 input vector, so D[:, 0] + D[:, 1] is the row's dot product
__device__ __forceinline__ void mma16816(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float shfl_idx_raw(float v, int src) {
    float r;
    asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=f"(r) : "f"(v), "r"(src));
    return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
    return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}
// hi / lo fp16 split of a pair of fp32 values: hi = rn(x), lo = rn(x - hi); x ~ hi + lo to ~22 bits
__device__ __forceinline__ void split_hilo(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    // packed conversions (cvt.rn.f16x2.f32 = F2FP, ALU pipe) -- the scalar F2F.F16.F32 runs on the slow conversion pipe, and a layer needs
    // ~100 of them per warp
    const __half2 h = __floats2half2_rn(x0, x1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// f[k-step][lane] = {b0, b1} for lanes 0..3 (column 0 = hi) and 4..7 (column 1 = lo).  Values (2u, 2u+1) of the vector go to k-step u / 8,
// lane u % 4 (+4 for lo), register (u % 8) / 4.
__device__ __forceinline__ void store_bfrag_pair(uint2* f, int u, float x0, float x1) {
    uint32_t hi, lo;
    split_hilo(x0, x1, hi, lo);
    uint32_t* w = reinterpret_cast<uint32_t*>(f + (u >> 3) * 8 + (u & 3)) + ((u & 7) >> 2);
    w[0] = hi;
    w[8] = lo;            // lane + 4: 4 uint2 further
}
__device__ __forceinline__ uint2 load_bfrag(const uint2* f, int ks, int lane) {
    uint2 b = make_uint2(0u, 0u);
    if (lane < 8) b = f[ks * 8 + lane];
    return b;
}
__device__ __forceinline__ const uint4& frag_at(const uint8_t* warp_part, int block) {      // warp_part already includes lane * 16
    return *reinterpret_cast<const uint4*>(warp_part + block * 512);
}
// byte offset of element (row, col) inside a 512-byte A-fragment block (see frag_pos in the packer)
__device__ __forceinline__ uint32_t frag_off(int row, int col) {
    const int lane = (row & 7) * 4 + ((col & 7) >> 1), reg = (row >> 3) + 2 * (col >> 3);
    return (uint32_t)(lane * 16 + reg * 4 + (col & 1) * 2);
}
// LayerNorm (module.py:26-37: weight only, eps 1e-5) of the residual vector, of which thread t holds elements 2t, 2t+1 in `v`.
// FRAG: the result goes to sm->xf as MMA B fragments, else to sm->xn as fp32.  gw = weight in shared memory.
template <bool FRAG, int PB = -1>
__device__ __forceinline__ void layer_norm(Ctx& c, float2 v, const float* gw) {
    Smem* sm = SM();
    float s = v.x + v.y, q = fmaf(v.x, v.x, v.y * v.y);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (c.lane == 0) { sm->red[c.warp] = s; sm->red[32 + c.warp] = q; }
    const float2 g = reinterpret_cast<const float2*>(gw)[c.tid];
    if (PB >= 0) { PROBE(PB) }
    cons_sync();
    float ts = 0.f, tq = 0.f;
#pragma unroll
    for (int w = 0; w < N_CONS_WARPS; ++w) { ts += sm->red[w]; tq += sm->red[32 + w]; }
    if (PB >= 0) { if (ts == 12345.f) tq += 1.f; PROBE(PB + 1) }
    const float mean = ts * (1.0f / C);
    const float var = fmaxf(tq * (1.0f / C) - mean * mean, 0.f);
    const float rstd = rsqrtf(var + 1e-5f);
    const float y0 = (v.x - mean) * rstd * g.x, y1 = (v.y - mean) * rstd * g.y;
    if (FRAG) store_bfrag_pair(&sm->xf[0][0], c.tid, y0, y1);
    else reinterpret_cast<float2*>(sm->xn)[c.tid] = make_float2(y0, y1);
    cons_sync();
}
struct XRegs {
    float4 a[3], b[3];
};
__device__ __forceinline__ XRegs load_x(const float* xn, int lane) {
    XRegs x;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float* xp = xn + k * 256 + lane * 8;
        x.a[k] = *reinterpret_cast<const float4*>(xp);
        x.b[k] = *reinterpret_cast<const float4*>(xp + 4);
    }
    return x;
}
// one row of 768 halves (shared memory) . x, result in every lane (head GEMV: row-major weights)
__device__ __forceinline__ float row_dot768(const uint8_t* wrow, const XRegs& x, int lane) {
    const uint4* wp = reinterpret_cast<const uint4*>(wrow) + lane;
    const uint4 w0 = wp[0], w1 = wp[32], w2 = wp[64];
    return warp_sum(dot8(w0, x.a[0], x.b[0]) + dot8(w1, x.a[1], x.b[1]) + dot8(w2, x.a[2], x.b[2]));
}

// Reduce-scatter + all-gather of a 768-vector of which every CTA holds a partial in sm->out2 (module.py:409-410).  hop = 0 after attention
// (adds the c_proj bias), 1 after the MLP.
//   sender warp:  line u (values 2u, 2u+1 of out2) -> rsl[me][u % 24] of rank u / 24
//   consumers:    thread 16 line + k polls rank k's partial of my rows 2 line, 2 line + 1; the 16 lanes of a group add them up by xor butterfly
//                 (the same order whatever the arrival order); lane 0 of the group leaves the sum in sm->sums[line]
//   sender warp:  sums[u % 24] -> xl[me][u % 24] of rank u / 24
//   consumers:    thread t picks up its own elements 2t, 2t+1 from xl[t / 24][t % 24]
// Remote stores are issued by the sender warps only (a remote store occupies its warp for ~100-250 cycles; the consumers go straight to
// polling).  The sender warps take part in no block barrier; consumers hand over through an mbarrier (one arrive per warp).
// Buffers are reused by every hop: a rank can only send hop h + 1 after it received all of hop h's all-gather, which every rank sends after
// polling all of hop h's partials -- so nobody overwrites a line that is still being waited for.
__device__ __forceinline__ void handoff(Ctx& c, uint64_t* bar) {       // this warp's writes to out2 / sums are done (arrive has release semantics)
    __syncwarp();
    if (c.lane == 0) mbar_arrive(bar);
}
__device__ __forceinline__ float2 residual_hop(Ctx& c, int hop, const float* bias48) {
    Smem* sm = SM();
    const uint32_t tag = 2u * c.lc + (uint32_t)hop + 1u;
    handoff(c, &sm->hs_rs);
    PROBE(hop ? 14 : 7)
    STAMP(hop ? 9 : 4)
    const int line = c.tid >> 4, k = c.tid & 15;
    float2 pv = wait_line(c, &sm->rsl[k][line], tag);
    PROBE(hop ? 15 : 8)
    STAMP(hop ? 10 : 5)
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        pv.x += __shfl_xor_sync(0xffffffffu, pv.x, o);
        pv.y += __shfl_xor_sync(0xffffffffu, pv.y, o);
    }
    if (bias48) { pv.x += bias48[2 * line]; pv.y += bias48[2 * line + 1]; }
    if (k == 0) sm->sums[line] = pv;
    handoff(c, &sm->hs_ag);
    STAMP(hop ? 13 : 12)
    const float2 res = wait_line(c, &sm->xl[c.tid / LINES][c.tid % LINES], tag);
    PROBE(hop ? 16 : 9)
    STAMP(hop ? 11 : 6)
    return res;
}
// sender warp: all remote stores of the two exchanges of every layer
__device__ __forceinline__ bool sender_wait(Ctx& c, uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (check_abort(c, spins)) return false;
    }
    return true;
}
__device__ __forceinline__ void sender_loop(Ctx& c, long long n_hops, int sidx) {
    Smem* sm = SM();
#pragma unroll 1
    for (long long h = 0; h < n_hops; ++h) {
        const uint32_t tag = (uint32_t)h + 1u, parity = (uint32_t)h & 1u;
        if (!sender_wait(c, &sm->hs_rs, parity)) return;
#pragma unroll
        for (int ii = 0; ii < N_CONS / 32 / N_SEND_WARPS; ++ii) {
            const int u = 32 * (sidx * (N_CONS / 32 / N_SEND_WARPS) + ii) + c.lane;
            float2 ov;
            asm volatile("ld.volatile.shared.v2.f32 {%0, %1}, [%2];" : "=f"(ov.x), "=f"(ov.y) : "r"(smem_u32(sm->out2 + 2 * u)) : "memory");
            send_line(c, &sm->rsl[c.i][u % LINES], (uint32_t)(u / LINES), ov.x, ov.y, tag);
        }
        if (!sender_wait(c, &sm->hs_ag, parity)) return;
#pragma unroll
        for (int ii = 0; ii < N_CONS / 32 / N_SEND_WARPS; ++ii) {
            const int u = 32 * (sidx * (N_CONS / 32 / N_SEND_WARPS) + ii) + c.lane;
            float2 sv;
            asm volatile("ld.volatile.shared.v2.f32 {%0, %1}, [%2];" : "=f"(sv.x), "=f"(sv.y) : "r"(smem_u32(sm->sums + u % LINES)) : "memory");
            send_line(c, &sm->xl[c.i][u % LINES], (uint32_t)(u / LINES), sv.x, sv.y, tag);
        }
    }
}

// Attention of my head at step j, layer l (module.py:214-227 with one query, causal): keys 0..j-1 from the cache, key j appended now.
// K stages then V stages of 24 tiles; warp w works on tiles w and w + 12 of every stage.  Two passes per warp (all scores, one max, then
// p and p.V without rescaling), the 12 warps' (m, l, o[48]) are merged through shared memory.
__device__ __forceinline__ void attention(Ctx& c, int l, int j) {
    const KParams& p = *c.p;
    Smem* sm = SM();
    const int total = j + 1;
    const int ntile = (total + 15) >> 4;
    const int nks = (ntile + TPS - 1) / TPS;
    const int tn = j >> 4, kk = j & 15;                // tile / key slot of the appended row
    const int tn_stage = tn / TPS, tn_ti = tn - tn_stage * TPS;
    const bool patcher = c.warp == (tn_ti % N_CONS_WARPS);
    const int g = c.lane >> 2, t = c.lane & 3;
    PROBE2(0)
    // q as B fragments (3 k-steps of 16 dims), pre-scaled by 1/sqrt(48) * log2(e) (module.py:196-198)
    uint32_t qb0[3], qb1[3];
    {
        const float qscale = 0.14433756729740643f * 1.4426950408889634f;
#pragma unroll
        for (int ds = 0; ds < 3; ++ds) {
            const float2 x01 = *reinterpret_cast<const float2*>(&sm->qv[16 * ds + 2 * t]), x89 = *reinterpret_cast<const float2*>(&sm->qv[16 * ds + 2 * t + 8]);
            uint32_t h01, l01, h89, l89;
            split_hilo(x01.x * qscale, x01.y * qscale, h01, l01);
            split_hilo(x89.x * qscale, x89.y * qscale, h89, l89);
            qb0[ds] = (g == 0) ? h01 : ((g == 1) ? l01 : 0u);
            qb1[ds] = (g == 0) ? h89 : ((g == 1) ? l89 : 0u);
        }
    }
    uint8_t* kg = (uint8_t*)p.a.kv_h + (((size_t)(l * 2 + 0) * NH + c.i) * KV_TILES + tn) * KV_TILE_BYTES;
    uint8_t* vg = (uint8_t*)p.a.kv_h + (((size_t)(l * 2 + 1) * NH + c.i) * KV_TILES + tn) * KV_TILE_BYTES;
    // Pass 1: scores of my tiles (warp w: tiles w and w + 12 of every stage), both tiles of a stage in flight together, one accumulator per
    // 16-dim slice so the three MMAs of a tile do not chain.  Tiles beyond ntile read stale (finite) ring bytes and are masked.
    float sa[2 * MAX_KV_STAGES], sb[2 * MAX_KV_STAGES];
    PROBE2(1)
#pragma unroll
    for (int s = 0; s < MAX_KV_STAGES; ++s) {
        sa[2 * s] = sa[2 * s + 1] = sb[2 * s] = sb[2 * s + 1] = -INFINITY;
        if (s < nks) {                                 // CTA-uniform
            uint8_t* base = const_cast<uint8_t*>(stage_wait(c));
            if (s < 4) { PROBE2(2 + 2 * s) }
            if (s == tn_stage && patcher) {
                // cache append (module.py:209-210): K[key kk][dims 2e, 2e+1] into the staged tile and into the cache
                if (c.lane < HD / 2) {
                    const int d = 2 * c.lane;
                    const uint32_t koff = (uint32_t)(d >> 4) * 512 + frag_off(kk, d & 15);
                    const __half2 k2 = *reinterpret_cast<const __half2*>(&sm->knew[d]);
                    *reinterpret_cast<__half2*>(base + (size_t)tn_ti * KV_TILE_BYTES + koff) = k2;
                    *reinterpret_cast<__half2*>(kg + koff) = k2;
                }
                __syncwarp();
            }
            const uint8_t* kt = base + (size_t)c.warp * KV_TILE_BYTES + c.lane * 16;
            uint4 ka[2][3];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int ds = 0; ds < 3; ++ds) ka[h][ds] = *reinterpret_cast<const uint4*>(kt + (size_t)h * (N_CONS_WARPS * KV_TILE_BYTES) + ds * 512);
            stage_peek_next(c);
            float sc[2][3][4];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int ds = 0; ds < 3; ++ds) {
                    sc[h][ds][0] = sc[h][ds][1] = sc[h][ds][2] = sc[h][ds][3] = 0.f;
                    mma16816(sc[h][ds], ka[h][ds], qb0[ds], qb1[ds]);
                }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int key = (s * TPS + c.warp + N_CONS_WARPS * h) * 16 + g;
                const float lo = (sc[h][0][0] + sc[h][0][1]) + (sc[h][1][0] + sc[h][1][1]) + (sc[h][2][0] + sc[h][2][1]);
                const float hi = (sc[h][0][2] + sc[h][0][3]) + (sc[h][1][2] + sc[h][1][3]) + (sc[h][2][2] + sc[h][2][3]);
                if (t == 0 && key < total) sa[2 * s + h] = lo;
                if (t == 0 && key + 8 < total) sb[2 * s + h] = hi;
            }
            if (s == tn_stage && patcher) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the slot's next bulk copy may overwrite the patched tile
            if (s < 4) { if (sa[2 * s] == 12345.f) sb[2 * s] = 0.f; PROBE2(3 + 2 * s) }
            stage_done(c);
        }
    }
    PROBE(21)
    float m_run = -INFINITY;
#pragma unroll
    for (int s = 0; s < 2 * MAX_KV_STAGES; ++s) m_run = fmaxf(m_run, fmaxf(sa[s], sb[s]));
#pragma unroll
    for (int o2 = 4; o2 < 32; o2 <<= 1) m_run = fmaxf(m_run, __shfl_xor_sync(0xffffffffu, m_run, o2));      // over the 8 lanes with my t
    m_run = __shfl_sync(0xffffffffu, m_run, 0);                // the t == 0 group holds the scores
    const float mref = (m_run == -INFINITY) ? 0.f : m_run;     // a warp without keys: every p is exp2(-inf) = 0
    // p of every tile as ready-made B fragments before the V stages arrive: lane (g, t) needs p[2t], p[2t+1] (b0) and p[2t+8], p[2t+9] (b1);
    // p[k] lives in lane 4 (k % 8).  (sa / sb are dead afterwards: the fragments reuse their registers.)
    float l_run = 0.f;
    uint32_t pf0[2 * MAX_KV_STAGES], pf1[2 * MAX_KV_STAGES];
#pragma unroll
    for (int u = 0; u < 2 * MAX_KV_STAGES; ++u) {
        pf0[u] = pf1[u] = 0u;
        if ((u >> 1) < nks) {       // CTA-uniform; raw shfl.sync below (a __shfl_sync under a branch the compiler cannot prove uniform costs a warp-sync call)
            const float pa = ex2_approx(sa[u] - mref), pb = ex2_approx(sb[u] - mref);       // ex2(-inf) = 0
            l_run += pa + pb;
            const float p0 = shfl_idx_raw(pa, 8 * t), p1 = shfl_idx_raw(pa, 8 * t + 4);
            const float p8 = shfl_idx_raw(pb, 8 * t), p9 = shfl_idx_raw(pb, 8 * t + 4);
            uint32_t h01, l01, h89, l89;
            split_hilo(p0, p1, h01, l01);
            split_hilo(p8, p9, h89, l89);
            // B columns 0, 2, 4, 6 = hi, 1, 3, 5, 7 = lo: only output columns 0 and 1 are read, the others may hold anything finite
            pf0[u] = (g & 1) ? l01 : h01;
            pf1[u] = (g & 1) ? l89 : h89;
        }
    }
    PROBE2(10)
    // Pass 2: o += V^T p, two independent accumulator sets (one per tile of the stage)
    float o[2][3][4];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int dt = 0; dt < 3; ++dt) { o[h][dt][0] = o[h][dt][1] = o[h][dt][2] = o[h][dt][3] = 0.f; }
#pragma unroll
    for (int s = 0; s < MAX_KV_STAGES; ++s) {
        if (s < nks) {
            uint8_t* base = const_cast<uint8_t*>(stage_wait(c));
            if (s < 4) { PROBE2(11 + 2 * s) }
            if (s == tn_stage && patcher) {
                // V^T[dims 2e, 2e+1][key kk]
                if (c.lane < HD / 2) {
                    const int d = 2 * c.lane;
                    const uint32_t voff0 = (uint32_t)(d >> 4) * 512 + frag_off(d & 15, kk), voff1 = (uint32_t)((d + 1) >> 4) * 512 + frag_off((d + 1) & 15, kk);
                    const __half v0 = sm->vnew[d], v1 = sm->vnew[d + 1];
                    uint8_t* vt = base + (size_t)tn_ti * KV_TILE_BYTES;
                    *reinterpret_cast<__half*>(vt + voff0) = v0;
                    *reinterpret_cast<__half*>(vg + voff0) = v0;
                    *reinterpret_cast<__half*>(vt + voff1) = v1;
                    *reinterpret_cast<__half*>(vg + voff1) = v1;
                }
                __syncwarp();
            }
            const uint8_t* vt = base + (size_t)c.warp * KV_TILE_BYTES + c.lane * 16;
            uint4 va[2][3];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int dt = 0; dt < 3; ++dt) va[h][dt] = *reinterpret_cast<const uint4*>(vt + (size_t)h * (N_CONS_WARPS * KV_TILE_BYTES) + dt * 512);
            stage_peek_next(c);
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int dt = 0; dt < 3; ++dt) mma16816(o[h][dt], va[h][dt], pf0[2 * s + h], pf1[2 * s + h]);
            if (s == tn_stage && patcher) {
                fence_proxy_async_global();            // my producer's later bulk copies (async proxy) must see the appended row
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            if (s < 4) { if (o[0][0][0] == 12345.f) l_run = 0.f; PROBE2(12 + 2 * s) }
            stage_done(c);
        }
    }
    PROBE(22)
#pragma unroll
    for (int o2 = 4; o2 < 32; o2 <<= 1) l_run += __shfl_xor_sync(0xffffffffu, l_run, o2);      // lane 0: sum over the t == 0 group
    if (c.lane == 0) { sm->wpart[c.warp][0] = m_run; sm->wpart[c.warp][1] = l_run; }
    if (t == 0) {
#pragma unroll
        for (int dt = 0; dt < 3; ++dt) {
            sm->wpart[c.warp][2 + dt * 16 + g] = (o[0][dt][0] + o[1][dt][0]) + (o[0][dt][1] + o[1][dt][1]);
            sm->wpart[c.warp][2 + dt * 16 + g + 8] = (o[0][dt][2] + o[1][dt][2]) + (o[0][dt][3] + o[1][dt][3]);
        }
    }
    cons_sync();
    PROBE(23)
    if (c.tid == 0) sm->kv_progress = c.lc + 1;        // the appended row is written and fenced (the patching warp passed the barrier)
    {       // thread 16 pr + w: warp w's share of outputs 2 pr, 2 pr + 1 (pr < 24); the 16 lanes of a group merge by butterfly
        const int pr = c.tid >> 4, w = c.tid & 15;
        float mw = -INFINITY, lw = 0.f;
        float2 ov = make_float2(0.f, 0.f);
        if (w < N_CONS_WARPS) {
            mw = sm->wpart[w][0];
            lw = sm->wpart[w][1];
            ov = *reinterpret_cast<const float2*>(&sm->wpart[w][2 + 2 * pr]);
        }
        float m = mw;
#pragma unroll
        for (int o2 = 1; o2 < 16; o2 <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o2));
        const float f = (mw > -INFINITY) ? exp2f(mw - m) : 0.f;
        float ls = f * lw, a0 = f * ov.x, a1 = f * ov.y;
#pragma unroll
        for (int o2 = 1; o2 < 16; o2 <<= 1) {
            ls += __shfl_xor_sync(0xffffffffu, ls, o2);
            a0 += __shfl_xor_sync(0xffffffffu, a0, o2);
            a1 += __shfl_xor_sync(0xffffffffu, a1, o2);
        }
        if (w == 0) {
            const float inv = 1.0f / ls;
            store_bfrag_pair(&sm->yf[0][0], pr, a0 * inv, a1 * inv);
        }
    }
    cons_sync();
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1) decode_c16_kernel(const __grid_constant__ KParams p) {
    Smem* sm = SM();
    const UmgenDecodeArgs& a = p.a;
    Ctx c;
    c.p = &p; c.abort_flag = (int*)a.status_i32;
    c.tid = threadIdx.x; c.warp = threadIdx.x >> 5; c.lane = threadIdx.x & 31;
    {
        uint32_t rk;
        asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rk));
        c.i = (int)rk;
    }
    c.k = 0; c.rdy = false; c.lc = 0; c.probe = nullptr; c.probe_t0 = 0;
    c.dbg_local = (a.grid & 2) != 0;
    const long long t_start = clock64();
    const bool dbg_same_layer = (a.grid & 1) != 0;      // debug: stream layer 0's matrices for every layer (L2-resident weights)
    c.t_dead = globaltimer_ns() + TIMEOUT_NS;
    const int L = (int)a.n_layer;
    const int n_steps = (int)a.n_steps;
    c.sbase = smem_u32(sm);
    c.rbase = mapa_u32(c.sbase, 0);
    c.rstride = mapa_u32(c.sbase, 1) - c.rbase;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSLOT; ++s) { mbar_init(&sm->full[s], 1); mbar_init(&sm->empty[s], N_CONS_WARPS); }
        sm->nbox = 0; sm->tok = 0; sm->kv_progress = 0;
        mbar_init(&sm->hs_rs, N_CONS_WARPS); mbar_init(&sm->hs_ag, N_CONS_WARPS);
        mbar_fence_init();
    }
    {       // no line may carry a valid tag before the first exchange
        uint4* z = &sm->rsl[0][0];
        constexpr int NZ = (sizeof(Smem::rsl) + sizeof(Smem::xl) + sizeof(Smem::candl)) / 16;
        static_assert(offsetof(Smem, xl) == offsetof(Smem, rsl) + sizeof(Smem::rsl) && offsetof(Smem, candl) == offsetof(Smem, xl) + sizeof(Smem::xl),
                      "line buffers are contiguous");
        for (int k = threadIdx.x; k < NZ; k += NT) z[k] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    cluster_sync_all();          // every CTA's line buffers are clean before anyone sends
    c.tl = nullptr;
    c.tl_base = clock64();

    const uint8_t* Wc = (const uint8_t*)p.oar_c16_h;
    const float* Fl = (const float*)a.oar_f;
    const __half* heads[3] = {(const __half*)a.head_map_h, (const __half*)a.head_bbox_h, (const __half*)a.head_img_h};
    const float* emb_tables[3] = {(const float*)a.map_table_f, (const float*)a.be_f, (const float*)a.img_table_f};

    if (c.warp > N_CONS_WARPS) {
        // ============================== sender warps ===========================================
        sender_loop(c, (long long)n_steps * L * 2, c.warp - (N_CONS_WARPS + 1));
    } else if (c.warp == N_CONS_WARPS) {
        // ============================== producer warp ==========================================
        if (c.lane == 0) {
#pragma unroll 1
            for (int j = 0; j < n_steps; ++j) {
                const int q = j + 1;
                const int ntile = (j + 16) >> 4;
                const int nks = (ntile + TPS - 1) / TPS;
#pragma unroll 1
                for (int l = 0; l < L; ++l) {
                    const uint8_t* wl = Wc + ((size_t)(dbg_same_layer ? 0 : l) * CL + c.i) * CTA_LAYER_BYTES;
#pragma unroll 1
                    for (int s = 0; s < ST_QKV; ++s) produce(c, wl + (size_t)s * SLOT_BYTES, SLOT_BYTES);
                    if (j > 0) {
                        const uint32_t need = (uint32_t)((j - 1) * L + l + 1);       // the row appended by step j - 1 in this layer
                        uint32_t spins = 0;
                        while (sm->kv_progress < need) {
                            if (check_abort(c, spins)) break;
                        }
                        // no proxy fence here: the appending warp fenced (fence.proxy.async.global) before the barrier that precedes the
                        // kv_progress store, and a fence in this thread would first drain the bulk copies it has in flight
                    }
#pragma unroll 1
                    for (int kv = 0; kv < 2; ++kv) {
                        const uint8_t* src = (const uint8_t*)a.kv_h + (((size_t)(l * 2 + kv) * NH + c.i) * KV_TILES) * KV_TILE_BYTES;
#pragma unroll 1
                        for (int s = 0; s < nks; ++s)
                            produce(c, src + (size_t)s * SLOT_BYTES, (uint32_t)min(TPS, ntile - s * TPS) * KV_TILE_BYTES);
                    }
#pragma unroll 1
                    for (int s = ST_QKV; s < ST_LAYER; ++s) produce(c, wl + (size_t)s * SLOT_BYTES, SLOT_BYTES);
                    if (*(volatile int*)c.abort_flag != 0) break;
                }
                if (needs_head(q)) {
                    const int mod = pos_mod(q);
                    const int V = vocab_of(mod);
                    const int r0 = (V * c.i) / CL, r1 = (V * (c.i + 1)) / CL;
#pragma unroll 1
                    for (int r = r0; r < r1; r += HEAD_ROWS)
                        produce(c, (const uint8_t*)heads[mod] + (size_t)r * (C * 2), (uint32_t)min(HEAD_ROWS, r1 - r) * C * 2);
                }
                if (*(volatile int*)c.abort_flag != 0) break;
            }
        }
    } else {
        // ============================== consumer warps =========================================
        const float* tar = (const float*)a.tar_feat_f;
        int* out_tokens = (int*)a.out_tokens_i32;
        int* picks = (int*)a.picks_i32;
        const int* pose_tok = (const int*)a.pose_tok_i32;
        const int* teacher = (const int*)a.teacher_i32;
        float* scratch = (float*)a.scratch_f;
        const int gq = c.lane >> 2;

        // ln_oar and the first layer's parameters -> shared memory; the zero columns of the fragment buffers stay zero
        if (c.tid < 192) cp_async16(sm->lno + 4 * c.tid, (const float*)a.ln_oar_f + 4 * c.tid);
        prefetch_params(c, Fl, sm->prm[0]);
        // The residual vector lives in registers: thread t of every CTA holds elements 2t, 2t+1 (all CTAs compute identical values).
        // Input of step 0: task embedding + TAR feature of index 0 (UMGen.py:1175,1215,1231)
        float2 x;
        {
            const float2 t0 = __ldg(reinterpret_cast<const float2*>(a.tske_f) + c.tid), t1 = __ldg(reinterpret_cast<const float2*>(tar) + c.tid);
            x = make_float2(t0.x + t1.x, t0.y + t1.y);
        }
        if (c.i == 0 && c.tid < 8) {
            const int qs[8] = {1, 5, 6, 1031, 1032, 1693, 1694, 2207};
            out_tokens[qs[c.tid] - 1] = forced_id(qs[c.tid]);
            picks[qs[c.tid] - 1] = forced_id(qs[c.tid]);
            if (c.tid < 3) { out_tokens[1 + c.tid] = pose_tok[c.tid]; picks[1 + c.tid] = pose_tok[c.tid]; }
        }
        cp_async_wait_all();
        cons_sync();

#pragma unroll 1
        for (int j = 0; j < n_steps; ++j) {
            const int q = j + 1;
            // TAR feature of the next position, fetched a whole step ahead of its use
            float2 tnext = make_float2(0.f, 0.f);
            if (j + 1 < SEQ) tnext = __ldg(reinterpret_cast<const float2*>(tar + (size_t)(j + 1) * C) + c.tid);
#pragma unroll 1
            for (int l = 0; l < L; ++l) {
#if UMGEN_DECODE_PROFILE
                c.probe = nullptr;
                if (c.tid == UMGEN_PROBE_TID && l == 1 && j == UMGEN_PROBE_STEP && (c.i == 0 || c.i == 11)) {
                    c.probe = (int*)a.status_i32 + (c.i == 0 ? 8 : 40);
                    c.probe_t0 = clock64();
                }
#endif
#if UMGEN_DECODE_PROFILE == 3
                c.tl = (a.debug_u64 && c.lane == 0 && l == 1 && j == UMGEN_PROBE_STEP) ? (long long*)a.debug_u64 + (c.i * N_CONS_WARPS + c.warp) * 16 : nullptr;
#endif
                STAMP(0)
                const float* prm = sm->prm[c.lc & 1u];
#if !UMGEN_C16_PRM_LATE
                // the next layer's parameters start their trip now (the buffer's last readers finished a layer ago)
                prefetch_params(c, Fl + (size_t)((l + 1 == L) ? 0 : l + 1) * LAYER_F, sm->prm[(c.lc + 1) & 1u]);
#endif

                // ---- LN1 -> q | k | v of my head (+bias) (module.py:206).  Stage s holds k-steps [8 s, 8 s + 8) of the 9 row tiles; warp (g3, i4)
                // multiplies tiles 3 g3 .. 3 g3 + 2 by k-steps 8 s + 2 i4, + 1; the four K-quarters of a row meet in sm->pq
                layer_norm<true, 17>(c, x, prm + PRM_LN1);
                PROBE(0)
                STAMP(1)
                {
                    const int g3 = c.warp >> 2, i4 = c.warp & 3;
                    float acc[3][4];
#pragma unroll
                    for (int m = 0; m < 3; ++m) { acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f; }
#pragma unroll
                    for (int s = 0; s < ST_QKV; ++s) {
                        const uint2 b0 = load_bfrag(&sm->xf[0][0], 8 * s + 2 * i4, c.lane), b1 = load_bfrag(&sm->xf[0][0], 8 * s + 2 * i4 + 1, c.lane);
                        const uint8_t* wp = stage_wait(c) + c.warp * 3072 + c.lane * 16;
                        uint4 af[6];
#pragma unroll
                        for (int m = 0; m < 6; ++m) af[m] = frag_at(wp, m);
                        stage_peek_next(c);
#pragma unroll
                        for (int m = 0; m < 3; ++m) mma16816(acc[m], af[m], b0.x, b0.y);
#pragma unroll
                        for (int m = 0; m < 3; ++m) mma16816(acc[m], af[3 + m], b1.x, b1.y);
                        stage_done(c);
                    }
                    if ((c.lane & 3) == 0) {
#pragma unroll
                        for (int m = 0; m < 3; ++m) {
                            sm->pq[i4][(3 * g3 + m) * 16 + gq] = acc[m][0] + acc[m][1];
                            sm->pq[i4][(3 * g3 + m) * 16 + gq + 8] = acc[m][2] + acc[m][3];
                        }
                    }
                    cons_sync();
                    if (c.tid < QKV_R / 2) {          // thread u: rows 2u, 2u+1 of q (u < 24) | k | v
                        const int r = 2 * c.tid;
                        const float v0 = ((sm->pq[0][r] + sm->pq[1][r]) + (sm->pq[2][r] + sm->pq[3][r])) + prm[PRM_BQKV + r];
                        const float v1 = ((sm->pq[0][r + 1] + sm->pq[1][r + 1]) + (sm->pq[2][r + 1] + sm->pq[3][r + 1])) + prm[PRM_BQKV + r + 1];
                        if (c.tid < HD / 2) *reinterpret_cast<float2*>(&sm->qv[r]) = make_float2(v0, v1);
                        else if (c.tid < HD) *reinterpret_cast<__half2*>(&sm->knew[r - HD]) = __floats2half2_rn(v0, v1);     // k, v live at cache precision
                        else *reinterpret_cast<__half2*>(&sm->vnew[r - 2 * HD]) = __floats2half2_rn(v0, v1);
                    }
                    cons_sync();
                }
                PROBE(1)
                STAMP(2)
                attention(c, l, j);
                PROBE(2)
                STAMP(3)
                // ---- c_proj split along K: all 768 rows x my head's 48 columns (module.py:227-229).  Stage s holds row tiles [24 s, +24) x 3 k-steps;
                // warp w takes tiles 24 s + 2 w, + 1
                {
                    uint2 yb[3];
#pragma unroll
                    for (int ks = 0; ks < 3; ++ks) yb[ks] = load_bfrag(&sm->yf[0][0], ks, c.lane);
#pragma unroll
                    for (int s = 0; s < ST_PROJ; ++s) {
                        const uint8_t* wp = stage_wait(c) + c.warp * 3072 + c.lane * 16;
                        uint4 af[6];
#pragma unroll
                        for (int m = 0; m < 6; ++m) af[m] = frag_at(wp, m);
                        stage_peek_next(c);
#pragma unroll
                        for (int tt = 0; tt < 2; ++tt) {
                            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                            for (int ks = 0; ks < 3; ++ks) mma16816(acc, af[tt * 3 + ks], yb[ks].x, yb[ks].y);
                            if ((c.lane & 3) == 0) {
                                const int row = (24 * s + 2 * c.warp + tt) * 16 + gq;
                                sm->out2[row] = acc[0] + acc[1];
                                sm->out2[row + 8] = acc[2] + acc[3];
                            }
                        }
                        stage_done(c);
                    }
                }
                PROBE(3)
                {       // residual (module.py:409)
                    const float2 s = residual_hop(c, 0, prm + PRM_BPROJ);
                    x.x += s.x;
                    x.y += s.y;
                }
                PROBE(4)
                // ---- LN2 -> my 192 rows of c_fc -> erf-GELU (module.py:245-247).  Stage s holds k-steps [6 s, +6) of the 12 row tiles; warp w owns tile w
                layer_norm<true, 19>(c, x, prm + PRM_LN2);
                PROBE(5)
                STAMP(7)
#if UMGEN_C16_PRM_LATE
                prefetch_params(c, Fl + (size_t)((l + 1 == L) ? 0 : l + 1) * LAYER_F, sm->prm[(c.lc + 1) & 1u]);
#endif
                {
                    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
                    for (int s = 0; s < ST_FC; ++s) {
                        uint2 b[6];
#pragma unroll
                        for (int ks = 0; ks < 6; ++ks) b[ks] = load_bfrag(&sm->xf[0][0], 6 * s + ks, c.lane);
                        const uint8_t* wp = stage_wait(c) + c.warp * 3072 + c.lane * 16;
                        uint4 af[6];
#pragma unroll
                        for (int m = 0; m < 6; ++m) af[m] = frag_at(wp, m);
                        stage_peek_next(c);
#pragma unroll
                        for (int ks = 0; ks < 6; ++ks) mma16816(acc[ks & 1], af[ks], b[ks].x, b[ks].y);
                        stage_done(c);
                    }
                    // lane 4 gq holds rows gq (acc0 + acc1) and gq + 8 (acc2 + acc3) of tile w
                    const float lo = gelu_erf((acc[0][0] + acc[1][0]) + (acc[0][1] + acc[1][1])), hi = gelu_erf((acc[0][2] + acc[1][2]) + (acc[0][3] + acc[1][3]));
                    const float lo1 = __shfl_down_sync(0xffffffffu, lo, 4), hi1 = __shfl_down_sync(0xffffffffu, hi, 4);
                    if ((c.lane & 7) == 0) {
                        store_bfrag_pair(&sm->hf[0][0], (c.warp * 16 + gq) >> 1, lo, lo1);
                        store_bfrag_pair(&sm->hf[0][0], (c.warp * 16 + gq + 8) >> 1, hi, hi1);
                    }
                    cons_sync();
                }
                PROBE(6)
                STAMP(8)
                // ---- MLP c_proj split along K (module.py:248): all 768 rows x my 192 columns.  Stages 2 pp, 2 pp + 1 hold k-steps [0, 6) and [6, 12) of
                // row tiles [12 pp, +12); warp w owns tile 12 pp + w
                {
                    uint2 hb[12];
#pragma unroll
                    for (int ks = 0; ks < 12; ++ks) hb[ks] = load_bfrag(&sm->hf[0][0], ks, c.lane);
#pragma unroll
                    for (int pp = 0; pp < ST_PROJ2 / 2; ++pp) {
                        float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
                        for (int hf2 = 0; hf2 < 2; ++hf2) {
                            const uint8_t* wp = stage_wait(c) + c.warp * 3072 + c.lane * 16;
                            uint4 af[6];
#pragma unroll
                            for (int m = 0; m < 6; ++m) af[m] = frag_at(wp, m);
                            stage_peek_next(c);
#pragma unroll
                            for (int ks = 0; ks < 6; ++ks) mma16816(acc[ks & 1], af[ks], hb[6 * hf2 + ks].x, hb[6 * hf2 + ks].y);
                            stage_done(c);
                        }
                        if ((c.lane & 3) == 0) {
                            const int row = (12 * pp + c.warp) * 16 + gq;
                            sm->out2[row] = (acc[0][0] + acc[1][0]) + (acc[0][1] + acc[1][1]);
                            sm->out2[row + 8] = (acc[0][2] + acc[1][2]) + (acc[0][3] + acc[1][3]);
                        }
                    }
                }
                PROBE(10)
                {         // residual (module.py:410)
                    const float2 s = residual_hop(c, 1, nullptr);
                    x.x += s.x;
                    x.y += s.y;
                }
                PROBE(11)
                cp_async_wait_all();                   // my share of the next layer's parameters has landed (made visible by the next barrier)
                c.lc++;
            }

            // ---- head + sampling (UMGen.py:1247-1250, 1046-1137)
            int tok;
            const int fid = forced_id(q);
            if (q <= 5) {
                tok = (fid >= 0) ? fid : __ldg(pose_tok + (q - 2));
            } else if (fid >= 0) {
                tok = fid;
            } else {
                const int mod = pos_mod(q);
                const int V = vocab_of(mod);
                const int k = (int)(mod == 0 ? a.top_k_map : (mod == 1 ? a.top_k_bbox : a.top_k_img));
                const int r0 = (V * c.i) / CL, r1 = (V * (c.i + 1)) / CL;
                const uint32_t mine = (uint32_t)q;     // tag of this step's candidate / logit lines
                layer_norm<false>(c, x, sm->lno);
                {
                    const XRegs xr = load_x(sm->xn, c.lane);
#pragma unroll 1
                    for (int r = r0; r < r1; r += HEAD_ROWS) {
                        const int nr = min(HEAD_ROWS, r1 - r);
                        const uint8_t* w = stage_wait(c);
                        stage_peek_next(c);
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int rr = c.warp + N_CONS_WARPS * h;
                            if (rr < nr) {
                                const float s = row_dot768(w + (size_t)rr * (C * 2), xr, c.lane);
                                if (c.lane == 0) sm->acc[r - r0 + rr] = s;
                            }
                        }
                        stage_done(c);
                    }
                    cons_sync();
                }
                if (a.logits_dump_f) {
                    float* dump = (float*)a.logits_dump_f + (size_t)(q - 1) * 8192;
                    for (int r = r0 + c.tid; r < r1; r += N_CONS) dump[r] = sm->acc[r - r0];
                    cons_sync();           // warp 0 overwrites acc while selecting
                }
                if (a.sample_topp) {
                    // ---- nucleus sampling: all-gather the logits through L2, every CTA samples identically (UMGen.py:915-965)
                    float* LG = scratch + SC_LOGIT;
                    for (int r = r0 + c.tid; r < r1; r += N_CONS)
                        asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(LG + 2 * r), "r"(__float_as_uint(sm->acc[r - r0])), "r"(mine) : "memory");
                    float v[TOPP_PER];
                    {
                        uint4 rr[(TOPP_PER + 1) / 2];
#pragma unroll
                        for (int t = 0; t < (TOPP_PER + 1) / 2; ++t) { const int line = c.tid + t * N_CONS; if (line < V / 2) rr[t] = ll_ld(LG + 4 * line); }
#pragma unroll
                        for (int t = 0; t < (TOPP_PER + 1) / 2; ++t) {
                            const int line = c.tid + t * N_CONS;
                            float a0 = -INFINITY, a1 = -INFINITY;
                            if (line < V / 2) {
                                uint32_t spins = 0;
                                while (!(rr[t].y == mine && rr[t].w == mine)) { if (check_abort(c, spins)) break; rr[t] = ll_ld(LG + 4 * line); }
                                a0 = __uint_as_float(rr[t].x); a1 = __uint_as_float(rr[t].z);
                            }
                            if (2 * t < TOPP_PER) v[2 * t] = a0;
                            if (2 * t + 1 < TOPP_PER) v[2 * t + 1] = a1;
                        }
                    }
                    const float pm = (float)(mod == 0 ? a.top_p_map : (mod == 1 ? a.top_p_bbox : a.top_p_img));
                    const float inv_t = 1.0f / (float)a.temperature;
                    const float u0 = philox_uniform(a.seed, (uint32_t)a.frame_index, (uint32_t)q, 0u);
                    int slot = block_topp_sample(sm, v, pm, inv_t, u0, c.tid);
                    // slot = tid' + s * N_CONS with s the thread-local position: id = 2 * (tid' + (s / 2) * N_CONS) + (s & 1)
                    int t = 2 * ((slot % N_CONS) + ((slot / N_CONS) >> 1) * N_CONS) + ((slot / N_CONS) & 1);
                    if (mod == 1) {
                        const int bidx = q - BBOX_FIRST_POS - 1;
                        const int prev = __ldg((const int*)a.prev_bbox_i32 + bidx);
                        const bool controlled = (a.control_mask >> ((q - BBOX_FIRST_POS) / 11)) & 1ull;
                        const float* row = (const float*)a.tar_bbox_logits_f + (size_t)bidx * 1028;
                        for (int pass = 0; pass < 2; ++pass) {
                            const bool go2 = pass == 0 ? controlled : (t == PAD_TOKEN && a.merge_ar_tar && prev != PAD_TOKEN);
                            if (!go2) continue;
#pragma unroll
                            for (int s = 0; s < TOPP_PER; ++s) {
                                const int id = c.tid + s * N_CONS;
                                v[s] = (id < 1028 && !(controlled && id == 1027)) ? __ldg(row + id) : -INFINITY;
                            }
                            const float uu = philox_uniform(a.seed, (uint32_t)a.frame_index, (uint32_t)q, 1u + pass);
                            t = block_topp_sample(sm, v, (float)a.top_p_bbox, inv_t, uu, c.tid);
                            if (pass == 1 && c.i == 0 && c.tid == 0) atomicAdd((int*)a.status_i32 + 2, 1);
                        }
                    }
                    bool wipe = false;
                    if (c.warp == 0) {
                        if (mod == 1) { t = bbox_rules(sm, a, c.lane, c.i, q, t, 0.f, true); wipe = (t & WIPE_BIT) != 0; t &= ~WIPE_BIT; }
                        if (c.lane == 0) {
                            if (wipe && c.i == 0)
                                for (int s = 1; s <= 10; ++s) out_tokens[q - 1 - s] = PAD_TOKEN;
                            if (wipe) for (int s = 1; s <= 10; ++s) sm->recent[(q - s) & 15] = PAD_TOKEN;
                            sm->tok = t;
                        }
                    }
                    cons_sync();
                    tok = sm->tok;
                } else {
                    if (c.warp == 0) {     // local top-k of my slice -> candidate lines {val, tag, id, tag} in every rank's shared memory
                        const int n = r1 - r0;
#pragma unroll 1
                        for (int r = 0; r < k; ++r) {
                            float bv = -INFINITY;
                            int bi = 0x7fffffff;
#pragma unroll 1
                            for (int s = c.lane; s < n; s += 32) {
                                const float vv = sm->acc[s];
                                if (vv > bv) { bv = vv; bi = s; }
                            }
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) {
                                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                            }
                            const int id = (bi == 0x7fffffff) ? 0x7fffffff : r0 + bi;
                            if (c.lane < CL) send_line(c, &sm->candl[c.i][r], (uint32_t)c.lane, bv, __int_as_float(id), mine);
                            if (c.lane == 0 && bi != 0x7fffffff) sm->acc[bi] = -INFINITY;
                            __syncwarp();
                        }
                    }
                    // every CTA merges all candidates and decides the token identically
                    const int ncand = CL * k;
                    float* candv = sm->stage;
                    int* candi = reinterpret_cast<int*>(sm->stage + CL * MAX_CAND);
                    if (c.tid < ncand) {
                        const int cta_s = c.tid / k;
                        const float2 cv = wait_line(c, &sm->candl[cta_s][c.tid - cta_s * k], mine);
                        candv[c.tid] = cv.x;
                        candi[c.tid] = __float_as_int(cv.y);
                    }
                    cons_sync();
                    if (c.warp == 0) {
                        const float u0 = philox_uniform(a.seed, (uint32_t)a.frame_index, (uint32_t)q, 0u);
                        int t = warp_topk_sample(candv, candi, ncand, k, 1.0f / (float)a.temperature, u0, c.lane);
                        bool wipe = false;
                        if (mod == 1) {
                            const float u2 = philox_uniform(a.seed, (uint32_t)a.frame_index, (uint32_t)q, 2u);
                            t = bbox_rules(sm, a, c.lane, c.i, q, t, u2, false);
                            wipe = (t & WIPE_BIT) != 0;
                            t &= ~WIPE_BIT;
                        }
                        if (c.lane == 0) {
                            if (wipe && c.i == 0)
                                for (int s = 1; s <= 10; ++s) out_tokens[q - 1 - s] = PAD_TOKEN;    // UMGen.py:1357-1365
                            if (wipe) for (int s = 1; s <= 10; ++s) sm->recent[(q - s) & 15] = PAD_TOKEN;
                            sm->tok = t;
                        }
                    }
                    cons_sync();
                    tok = sm->tok;
                }
            }
            if (c.dbg_local) tok = 0;          // free-running debug mode computes garbage: keep the table index in range
            int tok_used = tok;
            if (teacher != nullptr && q > 5 && fid < 0) tok_used = __ldg(teacher + (q - 1));
            if (c.tid == 0) {
                sm->recent[q & 15] = tok_used;
                if (c.i == 0 && q > 5) { out_tokens[q - 1] = tok_used; picks[q - 1] = tok; }
            }
            if (j == SEQ - 2) break;           // q = 2206 was the last sampled token; q = 2207 is forced

            // ---- the next input: embedding of the token + TAR feature of index j + 1 (UMGen.py:1046-1137, 1215-1231).
            // bos/eos -> axe, pose -> fouier_pe, map/image -> GMLP(codebook[tok]) (precomputed table), bbox3d -> be
            {
                const float* row;
                if (forced_id(q) >= 0) row = (const float*)a.axe_f + (size_t)forced_id(q) * C;
                else if (q <= 5) row = (const float*)a.fpe_f + (size_t)tok_used * C;
                else row = emb_tables[pos_mod(q)] + (size_t)tok_used * C;
                const float2 e = __ldg(reinterpret_cast<const float2*>(row) + c.tid);
                x = make_float2(e.x + tnext.x, e.y + tnext.y);
            }
            if (*(volatile int*)c.abort_flag != 0) break;
        }
        cp_async_wait_all();
        if (c.i == 0 && c.tid == 0) {
            int* st = (int*)a.status_i32;
            st[3] = n_steps;
            st[60] = (int)((clock64() - t_start) >> 10);      // kilo-cycles
        }
    }
    // nobody leaves while a peer may still write into its shared memory
    __syncwarp();
    cluster_sync_all();
}

}  // namespace c16
}  // namespace umgen

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
namespace umgen {
extern int64_t g_launches;

static cudaError_t c16_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attrs, cudaStream_t stream) {
    const size_t smem = sizeof(c16::Smem) + 128;
    cudaError_t e = cudaFuncSetAttribute(c16::decode_c16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(c16::decode_c16_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
    memset(cfg, 0, sizeof(*cfg));
    cfg->gridDim = dim3(c16::CL);
    cfg->blockDim = dim3(c16::NT);
    cfg->dynamicSmemBytes = smem;
    cfg->stream = stream;
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = c16::CL;
    attrs[0].val.clusterDim.y = 1;
    attrs[0].val.clusterDim.z = 1;
    cfg->attrs = attrs;
    cfg->numAttrs = 1;
    return cudaSuccess;
}

// number of 16-CTA clusters of the one-cluster decode kernel that can be resident at once on the current device (1 is needed)
int decode_c16_capacity() {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attrs[1];
    if (c16_config(&cfg, attrs, nullptr) != cudaSuccess) { cudaGetLastError(); return 0; }
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, (const void*)c16::decode_c16_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
int64_t decode_c16_scratch_floats() { return c16::SC_TOTAL; }

int decode_c16_launch(const UmgenDecodeArgs* args, const void* oar_c16_h, cudaStream_t stream) {
    if (!oar_c16_h) { set_error("the one-cluster decode kernel needs oar_c16_h (umgen_tools_pack_oar_c16)"); return -1; }
    if (decode_c16_capacity() < 1) { set_error("device cannot hold a cluster of 16 CTAs of the one-cluster decode kernel"); return -3; }
    c16::KParams kp;
    kp.a = *args;
    kp.oar_c16_h = oar_c16_h;
    UMGEN_CUDA_OK(cudaMemsetAsync(args->scratch_f, 0, c16::SC_TOTAL * sizeof(float), stream));
    UMGEN_CUDA_OK(cudaMemsetAsync(args->status_i32, 0, 96 * sizeof(int), stream));
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attrs[1];
    void* kargs[] = {&kp};
    UMGEN_CUDA_OK(c16_config(&cfg, attrs, stream));
    UMGEN_CUDA_OK(cudaLaunchKernelExC(&cfg, (const void*)c16::decode_c16_kernel, kargs));
    g_launches += 1;
    return 0;
}

// Element (row, col) of a 16x16 tile at position e (in halves) of its 512-byte mma.m16n8k16 A-fragment block:
// lane = e / 8, register = (e % 8) / 2, half = e % 2; row = lane / 4 + 8 (register & 1), col = 2 (lane % 4) + half + 8 (register / 2)
__device__ __forceinline__ void frag_pos16(int e, int& row, int& col) {
    const int lane = e >> 3, reg = (e & 7) >> 1, hp = e & 1;
    row = (lane >> 2) + 8 * (reg & 1);
    col = 2 * (lane & 3) + hp + 8 * (reg >> 1);
}
// oar_h [L][c_attn | c_proj | c_fc | mlp c_proj] (row-major) -> oar_c16_h [L][16 CTAs][24 stages][warp 12][6 blocks][256 halves]
// (layout documented in include/umgen.h)
__global__ void pack_c16_kernel(const __half* __restrict__ src, __half* __restrict__ dst, int n_layer) {
    const size_t per_cta = c16::CTA_LAYER_BYTES / 2;
    const size_t total = (size_t)n_layer * c16::CL * per_cta;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t l = idx / (c16::CL * per_cta);
        const size_t rem = idx - l * (c16::CL * per_cta);
        const int r = (int)(rem / per_cta);
        const int o = (int)(rem - (size_t)r * per_cta);
        const int st = o / 18432, o1 = o % 18432, w = o1 / 1536, b = (o1 % 1536) / 256;
        int row, col;
        frag_pos16(o1 % 256, row, col);
        const __half* wsrc = src + l * (size_t)UMGEN_OAR_LAYER_H;
        size_t s;
        if (st < 6) {                       // c_attn: k-steps [8 st, +8) x 9 row tiles; warp (g3, i4): tiles 3 g3 + m, k-steps 8 st + 2 i4 + b / 3
            const int g3 = w >> 2, i4 = w & 3, tile = 3 * g3 + b % 3, ks = 8 * st + 2 * i4 + b / 3;
            const int which = tile / 3, e48 = (tile % 3) * 16 + row;
            s = (size_t)(which * C + r * HD + e48) * C + ks * 16 + col;
        } else if (st < 8) {                // c_proj: row tiles [24 (st - 6), +24) x 3 k-steps; warp w: tiles + 2 w + b / 3, k-step b % 3
            const int tile = 24 * (st - 6) + 2 * w + b / 3, ks = b % 3;
            s = (size_t)3 * C * C + (size_t)(tile * 16 + row) * C + r * HD + ks * 16 + col;
        } else if (st < 16) {               // c_fc: k-steps [6 (st - 8), +6) x 12 row tiles; warp w: tile w, k-step + b
            const int ks = 6 * (st - 8) + b;
            s = (size_t)4 * C * C + (size_t)(r * c16::FC_R + w * 16 + row) * C + ks * 16 + col;
        } else {                            // mlp c_proj: row tiles [12 pp, +12), k-steps [6 half, +6); warp w: tile 12 pp + w, k-step 6 half + b
            const int pp = (st - 16) >> 1, hf2 = (st - 16) & 1, tile = 12 * pp + w, ks = 6 * hf2 + b;
            s = (size_t)4 * C * C + (size_t)FF * C + (size_t)(tile * 16 + row) * FF + r * c16::FC_R + ks * 16 + col;
        }
        dst[idx] = wsrc[s];
    }
}
}  // namespace umgen

using namespace umgen;

// Tools-library entry points of the one-cluster design study (not part of the product ABI in include/umgen.h)
extern "C" int umgen_tools_decode_c16_capacity(void) { return decode_c16_capacity(); }
extern "C" int64_t umgen_tools_decode_c16_scratch_floats(void) { return decode_c16_scratch_floats(); }
extern "C" int umgen_tools_decode_c16_frame(const UmgenDecodeArgs* args, const void* oar_c16_h, void* stream_v) {
    if (!args) { set_error("null args"); return -1; }
    return decode_c16_launch(args, oar_c16_h, (cudaStream_t)stream_v);
}

extern "C" int umgen_tools_pack_oar_c16(const void* oar_h, void* oar_c16_h, int64_t n_layer, void* stream_v) {
    if (!oar_h || !oar_c16_h || n_layer < 1) { set_error("bad arguments"); return -1; }
    pack_c16_kernel<<<1184, 256, 0, (cudaStream_t)stream_v>>>((const __half*)oar_h, (__half*)oar_c16_h, (int)n_layer);
    UMGEN_CUDA_OK(cudaGetLastError());
    g_launches += 1;
    return 0;
}
