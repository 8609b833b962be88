// Cluster OAR decode kernel: one persistent launch of 8 thread-block clusters x 8 CTAs runs every single-token step of a
// frame.  Same contract as decode.cu (reference models/UMGen.py:1151-1383, models/module.py:378-428).  decode.cu turned out
// to be bound by instruction issue (38 k warp-instructions per CTA and layer at IPC 0.3, profiles/), not by HBM or by its
// six L2 hops per layer, so this kernel is built around two ideas:
//
//   (1) every dense contraction runs on the tensor cores.  The matrices are packed by the host in mma.m16n8k16 A-fragment
//       order, so a lane fetches its fragment with one conflict-free 16-byte shared-memory load; the fp32 activation vector
//       is split into fp16 (hi, lo) and occupies columns 0 and 1 of the B operand, so the product keeps ~22 mantissa bits
//       and one MMA does 256 multiply-adds.
//   (2) the partitioning keeps exchanges head-local.  A cluster of 8 CTAs owns 2 attention heads.  Per head its CTAs compute
//       the 144 q|k|v rows (18 each) and all-gather them over distributed shared memory (st.async + mbarrier complete_tx);
//       attention is split-KV with cache row r owned by CTA r % 8 (a CTA only ever reads cache rows it wrote itself); the 8
//       partials are all-gathered over DSMEM and merged redundantly.  c_proj is split along K: the cluster multiplies its 96
//       attention outputs into all 768 rows (96 rows per CTA); the 8 per-cluster partial vectors cross the chip once through
//       tagged 16-byte lines in L2: every thread of every CTA polls the 8 partials of its own two rows and adds them in
//       cluster order (+ bias + residual) -- 8x the poll traffic of "one rank sums a slice and fans it out", but one exchange
//       less on the critical path (592 -> 476 us/step).  c_fc is split by rows (48 per CTA) so the GELU output never leaves
//       the CTA; the MLP c_proj is split along K (768 x 48 per CTA), reduce-scattered over DSMEM inside the cluster, and
//       crosses the chip once like c_proj: 2 L2 hops and 3 DSMEM hops per layer.
//
// Weights (221 184 B per CTA and layer, one contiguous run) and the CTA's cache rows are streamed HBM -> shared memory by a
// producer warp with 1-D bulk copies ahead of use.  8 clusters of 8 fit every B200 seen so far (umgen_decode_cluster_capacity
// reports 15 on this pool, so one head per cluster -- 16 clusters -- does not).
#include <stdlib.h>

#define UMGEN_CONS_WARPS 12
#include "decode_shared.cuh"

#ifndef UMGEN_SMALL_CODE
#define UMGEN_SMALL_CODE 0          // experiment: 1 = out-of-line copies of the big helpers (layer loop body < 32 KB, the L1.5 instruction cache)
#endif
#if UMGEN_SMALL_CODE
#define UMGEN_INLINE __noinline__
#else
#define UMGEN_INLINE __forceinline__
#endif

namespace umgen {
namespace cl {

constexpr int CL = 8;                        // CTAs per cluster = KV splits per head
constexpr int HPC = 2;                       // heads per cluster
constexpr int NCL = NH / HPC;                // 8 clusters
constexpr int GRID = NCL * CL;               // 64 CTAs
constexpr int XS = C / CL;                   // 96 residual rows summed by CTA rank i
constexpr int QKV_R = HPC * 3 * HD / CL;     // 36 q|k|v rows per CTA (2 heads x {q,k,v} x 6)
constexpr int FC_R = FF / GRID;              // 48 c_fc rows per CTA
constexpr int KV_TILES = 18;                 // 16-key tiles per (layer, k|v, head, owner): 8 owners x 288 rows >= 2207
constexpr uint32_t KV_TILE_BYTES = 16 * HD * 2;      // 1536: one tile = 3 fragment blocks of 512 B
static_assert(CL * KV_TILES * 16 * HD == UMGEN_KV_ROWS * HD, "cache geometry");
constexpr int KSTEPS = C / 16;               // 48 k-steps of 16 over the 768 inputs
constexpr int KS_PER_WARP = KSTEPS / N_CONS_WARPS;     // 4
static_assert(KSTEPS % N_CONS_WARPS == 0 && QKV_R == 36 && FC_R == 48, "geometry");
constexpr uint32_t B_QKV = QKV_R * C * 2;    // 55 296
constexpr uint32_t B_PROJ = XS * HPC * HD * 2;     // 18 432
constexpr uint32_t B_FC = FC_R * C * 2;      // 73 728
constexpr uint32_t B_PROJ2 = C * FC_R * 2;   // 73 728
constexpr uint32_t CTA_LAYER_BYTES = B_QKV + B_PROJ + B_FC + B_PROJ2;      // 221 184
static_assert((size_t)CTA_LAYER_BYTES * GRID == (size_t)UMGEN_OAR_LAYER_H * 2, "cluster packing");
constexpr uint32_t QKV_WARP_BYTES = KS_PER_WARP * (2 * 512 + 128);          // per warp: 4 k-steps x (2 full tiles + 4-row tile)
constexpr uint32_t FC_WARP_BYTES = KS_PER_WARP * 3 * 512;
static_assert(QKV_WARP_BYTES * N_CONS_WARPS == B_QKV && FC_WARP_BYTES * N_CONS_WARPS == B_FC, "fragment packing");

// Poll pacing of the L2 hops (cycles): every CTA polls 3072 lines per round and all 64 CTAs poll the same lines, which loads L2 with several times
// the bytes of the weight stream; a pause before the first round / between rounds trades detection latency for L2 bandwidth.
#ifndef UMGEN_POLL_DELAY
#define UMGEN_POLL_DELAY 0
#endif
#ifndef UMGEN_POLL_BACKOFF
#define UMGEN_POLL_BACKOFF 0
#endif
#ifndef UMGEN_MAX_SCENES
#define UMGEN_MAX_SCENES 3          // scenes one launch can decode in lockstep (umgen_decode_frames): they share every weight fragment and every exchange
#endif
constexpr int NB_MAX = UMGEN_MAX_SCENES;
static_assert(NB_MAX >= 1 && NB_MAX <= 4, "a scene takes two of the 8 columns of the MMA B operand");
constexpr int NSLOT = 16;
constexpr int HEAD_ROWS = 24;                // head rows per ring stage
constexpr int PART_VALS = 50;                // (m, l, o[48])
constexpr int PART_STRIDE = 52;
constexpr int CREP = 8;                      // replicas of the candidate lines (reader rank i polls replica i)
constexpr uint64_t TIMEOUT_NS = 10ull * 1000 * 1000 * 1000;
constexpr int TAR_LATE_ROW0 = UMGEN_TAR_LATE_ROW0;      // first tar_feat row covered by args.tar_ready_i32

// global scratch (floats): tagged 16-byte lines {v0, tag, v1, tag}; NB = scenes of the launch
constexpr int LINES_X = XS / 2;                                   // 48 lines per (cluster, rank) slice
constexpr int CANDV = GRID * MAX_CAND * 4;
template <int NB>
struct Scratch {
    static constexpr int GP = 0;                                  // [2][NB][NCL][CL][48][4] c_proj partials
    static constexpr int GR = GP + 2 * NB * NCL * CL * LINES_X * 4;   // [2][NB][NCL][CL][48][4] MLP c_proj partials
    static constexpr int CAND = GR + 2 * NB * NCL * CL * LINES_X * 4; // [NB][CREP][GRID][MAX_CAND][4]
    static constexpr int LOGIT = CAND + NB * CREP * CANDV;        // [NB][8192 values] (top-p mode only)
    static constexpr int TOTAL = LOGIT + NB * 2 * 8192;
};

// per-layer fp32 parameters staged in shared memory one layer ahead (cp.async): ln_1 | ln_2 | my c_proj bias rows | my c_attn bias rows
constexpr int PRM_LN1 = 0, PRM_LN2 = C, PRM_BPROJ = 2 * C, PRM_BQKV = 3 * C, PRM_FLOATS = 3 * C + 40;
// F offsets inside one layer of oar_f: ln_1[768] | c_attn.bias[2304] | c_proj.bias[768] | ln_2[768]
constexpr int F_LN1 = 0, F_BQKV = C, F_BPROJ = C + 3 * C, F_LN2 = C + 3 * C + C, LAYER_F = UMGEN_OAR_LAYER_F;

template <int NB>
struct KParamsT {
    UmgenDecodeArgs a[NB];           // one per scene; the weights, the sampling set-up, n_steps and prefix_len are those of a[0]
};

// what the sampler / rule templates of decode_shared.cuh need, per scene
struct SceneSm {
    float* stage;                    // the launch's one candidate / TAR-head row buffer (SmemRest::stage): the scenes' tokens are decided one after the other
    float red[64];
    float corners[MAX_BOX][8];
    int box_dropped[MAX_BOX];
    int recent[16];
    volatile int tok;
    int nbox;
};
static_assert(2 * GRID * MAX_CAND >= 1028, "stage doubles as the TAR-head row scratch");

// everything in shared memory except the ring
// first in shared memory whatever NB is: what the out-of-line slow paths of the waits need
struct __align__(16) WaitInfo {
    int* abort_flag;                 // status word of scene 0
    uint64_t t_dead;                 // globaltimer deadline of the launch: a wait that is still spinning then is a deadlock
    int dbg_local;                   // debug (args.grid bit 1): polls do not wait
};
template <int NB>
struct __align__(128) SmemRest {
    WaitInfo wi;
    float lno[C];                    // ln_oar weight
    float prm[PRM_FLOATS];           // layer parameters; the next layer's are fetched (cp.async) once LN2 has read the last of this layer's
    // vectors that feed an MMA, as B fragments: [k-step][lane 8 s + (0..3) hi, 8 s + (4..7) lo of scene s] = {b0, b1}
    uint2 xf[KSTEPS][8 * NB];        // normalised residual vector
    uint2 yf[HPC * HD / 16][8 * NB]; // merged attention output of my heads
    uint2 hf[FC_R / 16][8 * NB];     // my slice of the MLP hidden vector
    // exchange targets inside the cluster, written remotely as self-flagged 16-byte lines {v0, tag, v1, tag} (tag = layer count + 1)
    uint4 qkvl[NB][CL][QKV_R / 2];   // q | k_new | v_new rows of my heads as computed by each rank: [rank][(hh, {q,k,v}, pair)]
    uint4 partl[NB][HPC][CL][PART_VALS / 2];   // split-KV partials (m, l, o[48]) of the 8 ranks; the same memory then receives the MLP c_proj partials of
                                     // my 96 rows from the 8 ranks (reduce-scatter, rsl()): the two are never live together and carry different tags
    float out2[NB][C];               // my K-slice of the MLP c_proj output before the reduce-scatter; after the last layer: the normalised vector feeding the head
                                     // GEMV; during attention: the warps' split-KV partials (wpart())
    float pq[N_CONS_WARPS][NB][FC_R];      // per-warp K-slice partials of the c_attn / c_fc rows
    float acc[NB][136];              // head logits of my slice (8192 / 64 rows)
    float lnred[NB][64];             // LayerNorm statistics per warp
    SceneSm sc[NB];
    float stage[2 * GRID * MAX_CAND];      // candidates (values | ids) of the scene being decided / TAR-head row scratch (>= 1028)
    uint64_t full[NSLOT];
    uint64_t empty[NSLOT];
    uint32_t fl_off[NSLOT];
    uint32_t fl_bytes[NSLOT];
    volatile uint32_t kv_progress;   // layers (step * L + layer + 1) whose cache rows are written and fenced
    void* kv_ptr[NB];                // the scenes' caches (kernel parameters indexed by a runtime scene number would go through local memory)
};
// the ring takes what the 227 KB of a CTA leave: 184 / 164 / 144 KB for 1 / 2 / 3 scenes (-DUMGEN_RING_KB=n pins the one-scene ring)
template <int NB>
constexpr uint32_t ring_bytes() {
#ifdef UMGEN_RING_KB
    if (NB == 1) return (uint32_t)UMGEN_RING_KB * 1024u;
#endif
    return (uint32_t)((227 * 1024 - 128 - sizeof(SmemRest<NB>)) / 2048 * 2048);
}
static_assert(sizeof(WaitInfo) % 16 == 0 && offsetof(SmemRest<1>, lno) % 16 == 0 && offsetof(SmemRest<1>, prm) % 16 == 0 && offsetof(SmemRest<NB_MAX>, qkvl) % 16 == 0 &&
              offsetof(SmemRest<NB_MAX>, prm) % 16 == 0,
              "cp.async / 16-byte line alignment");
template <int NB>
struct __align__(128) SmemT : SmemRest<NB> {
    uint8_t ring[ring_bytes<NB>()];
};
static_assert(sizeof(SmemT<1>) + 128 <= 227 * 1024 && sizeof(SmemT<NB_MAX>) + 128 <= 227 * 1024, "shared memory budget");
// the c_fc and the MLP c_proj part of a layer (73 728 B each) are resident together (ring_next's reserve)
static_assert(ring_bytes<(NB_MAX < 3 ? NB_MAX : 3)>() >= 2 * 73728, "up to three scenes: the c_fc + MLP c_proj pair fits the ring");
static_assert(ring_bytes<NB_MAX>() >= 73728 + 55296, "the ring holds the largest stage beside the c_attn stage");

extern __shared__ __align__(128) uint8_t smem_raw_cl[];
template <int NB>
__device__ __forceinline__ SmemT<NB>* SM() { return reinterpret_cast<SmemT<NB>*>(smem_raw_cl); }
// buffers that share memory with another one whose lifetime does not overlap theirs (the ring needs every KB: 227 KB per CTA)
template <int NB>
__device__ __forceinline__ float* wpart(SmemT<NB>* sm, int s, int warp) {          // [NB][12 warps][PART_STRIDE] in out2
    static_assert(N_CONS_WARPS * PART_STRIDE <= C, "wpart fits out2");
    return sm->out2[s] + warp * PART_STRIDE;
}
template <int NB>
__device__ __forceinline__ uint4* rsl(SmemT<NB>* sm, int s, int rank) {            // [NB][CL][LINES_X] in partl
    static_assert(CL * LINES_X <= HPC * CL * (PART_VALS / 2), "rsl fits partl");
    return &sm->partl[s][0][0][0] + rank * LINES_X;
}
constexpr uint32_t RSL_TAG = 0x80000000u;      // reduce-scatter lines carry (layer tag | RSL_TAG): never equal to a partial line's tag

struct Ring {
    uint32_t head = 0, k = 0;
};
struct Stage {
    uint32_t off, slot, parity;
};
// reserve: bytes of the stage that follows this one and should be resident together with it (c_fc and MLP c_proj: when the second would have to
// wrap into the first, its copy could only start once the first is released -- measured as ~2 000 cycles of exposed wait per layer -- so the
// pair wraps together).  Ignored when the pair does not fit the ring at all.
template <uint32_t RING_BYTES>
__device__ __forceinline__ Stage ring_next(Ring& r, uint32_t bytes, uint32_t reserve = 0) {
    if (r.head + bytes > RING_BYTES || (bytes + reserve <= RING_BYTES && r.head + bytes + reserve > RING_BYTES)) r.head = 0;
    Stage s{r.head, r.k % NSLOT, (r.k / NSLOT) & 1u};
    r.head += bytes;
    r.k++;
    return s;
}

struct Ctx {
    const UmgenDecodeArgs* a; // the launch's scenes: a[0 .. NB)
    int* abort_flag;
    int* probe;
    long long probe_t0;
    int cta, tid, warp, lane;
    int h, i;                 // cluster index and rank in the cluster
    Ring ring;
    uint32_t lc;              // layers completed so far (tag of the DSMEM lines, parity of the L2 buffers and the parameter buffers)
    float* scratch;
    bool dbg_local;           // debug (args.grid bit 1): send only to myself, polls do not wait -> wrong results, isolates the exchange cost
    bool acct;                // thread 0 of CTA 0: account the cycles spent in each kind of wait (status[60..])
    long long acc_ring, acc_x, acc_poll, acc_attn, acc_head;
    long long* tl;            // timeline row of this warp (UMGEN_DECODE_PROFILE == 3), lane 0 only
    long long tl_base;
};
// -DUMGEN_DECODE_PROFILE=1 compiles in the cycle probes of one layer (status[8..], [40..]) and the wait accounting of CTA 0 / thread 0
// (status[60..63]).  Off by default: the per-layer path is latency-bound on its serial instruction count, every inlined probe costs time.
#ifndef UMGEN_DECODE_PROFILE
#define UMGEN_DECODE_PROFILE 0
#endif
// -DUMGEN_DECODE_DEBUG=1 compiles in the experiment knobs of args.grid (bit 1: free-running CTAs that send only to themselves and do not wait
// in polls -- wrong results, isolates the exchange cost).  Off by default: the flag costs a register and a test at every send and poll.
#ifndef UMGEN_DECODE_DEBUG
#define UMGEN_DECODE_DEBUG 0
#endif
#if UMGEN_DECODE_DEBUG
#define DBG_LOCAL(c) ((c).dbg_local)
#define DBG_LOCAL_SM() (wait_info()->dbg_local != 0)
#else
#define DBG_LOCAL(c) false
#define DBG_LOCAL_SM() false
#endif
#ifndef UMGEN_PROBE_TID
#define UMGEN_PROBE_TID 0           // the consumer thread whose view of the layer the probes record
#endif
#if UMGEN_DECODE_PROFILE == 3       // timeline: lane 0 of every consumer warp of every CTA stamps its clock (relative to the CTA's start)
#define STAMP(n) if (c.tl) { c.tl[n] = (long long)globaltimer_ns(); }      // one clock for the whole chip (clock64 is per SM / GPC)
#else
#define STAMP(n)
#endif
#if UMGEN_DECODE_PROFILE == 1
#define PROBE(k) if (c.probe) { c.probe[k] = (int)(clock64() - c.probe_t0); }
#else
#define PROBE(k)
#endif
#if UMGEN_DECODE_PROFILE == 1
#define ACCT_BEGIN() long long acct_t0 = 0; if (c.acct) acct_t0 = clock64();
#define ACCT_END(field) if (c.acct) c.field += clock64() - acct_t0;
#else
#define ACCT_BEGIN()
#define ACCT_END(field)
#endif

// Every wait gives up when the launch's deadline has passed (or another CTA already gave up) and raises the abort word the host checks.  The slow
// paths are out of line and take their context from shared memory: the layer loop has to stay inside the 32 KB instruction cache
// (a body beyond it costs this 64-CTA kernel ~25 %, measured), and there are ~20 wait sites per layer.
__device__ __forceinline__ WaitInfo* wait_info() { return reinterpret_cast<WaitInfo*>(smem_raw_cl); }
__device__ __noinline__ bool check_abort_slow() {
    WaitInfo* wi = wait_info();
    if (*(volatile int*)wi->abort_flag != 0) return true;
    if (globaltimer_ns() > wi->t_dead) {
        atomicCAS(wi->abort_flag, 0, 200);      // a wait this long is a deadlock: every CTA drains
        return true;
    }
    return false;
}
__device__ __forceinline__ bool check_abort(Ctx& c, uint32_t& spins) {
    ++spins;
    if ((spins & 0x3ffu) == 0) return check_abort_slow();
    return false;
}
__device__ __noinline__ void wait_mbar_slow(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (true) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if (((++spins) & 0x3ffu) == 0 && check_abort_slow()) return;
    }
}
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {      // non-blocking
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void wait_mbar(Ctx& c, uint64_t* bar, uint32_t parity) {
    if (!mbar_try_wait(bar, parity)) wait_mbar_slow(smem_u32(bar), parity);
}

// Register pressure: 13 warps leave 128 registers per thread, and every value the compiler keeps in local memory costs an L2 round trip when it
// is read back (the 28 KB of L1 left beside 227 KB of shared memory do not hold the spill slots of 13 warps): ~0.3 us each on the serial path
// of a layer.  The compiler likes to hoist per-thread index arithmetic (functions of tid, rank, cluster) out of the step / layer loops and then
// spills the results; REFRESH makes the inputs opaque again so that the arithmetic is redone where it is used (a few ALU instructions).
#ifndef UMGEN_NO_REFRESH
#define REFRESH(c)                                                                                   \
    do {                                                                                             \
        asm volatile("" : "+r"((c).tid), "+r"((c).h), "+r"((c).i));                                  \
        (c).warp = (c).tid >> 5;                                                                     \
        (c).lane = (c).tid & 31;                                                                     \
    } while (0)
#else
#define REFRESH(c)
#endif

// ---- DSMEM ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
// Exchanges inside the cluster use the same self-flagged lines as the L2 hops: the sender stores {v0, tag, v1, tag} straight into the
// receiver's shared memory (one st.shared::cluster.v4 per thread and scene), the receiver polls its own shared memory until both tags
// match.  No mbarrier, no sender-side barrier.  (st.async + complete_tx serialises
// every store on the receiver's mbarrier -- measured ~5 cycles per 4-byte store, 4 000 cycles for the 768-value reduce-scatter -- and
// store + barrier + release-arrive puts two remote trips back to back.)  Each 8-byte half carries its own tag, so a torn 16-byte store is harmless.
__device__ __forceinline__ uint32_t remote(const Ctx& c, const void* p, uint32_t rank) { return mapa_u32(smem_u32(p), rank); }
__device__ __forceinline__ void send_line(const Ctx& c, uint4* dst, uint32_t rank, float v0, float v1, uint32_t tag) {
    if (DBG_LOCAL(c) && (int)rank != c.i) return;
    asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %2};" ::"r"(remote(c, dst, rank)), "r"(__float_as_uint(v0)), "r"(tag), "r"(__float_as_uint(v1))
                 : "memory");
}
__device__ __forceinline__ uint4 lds_line(uint32_t addr) {
    uint4 r;
    asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr) : "memory");
    return r;
}
// poll N lines of my own shared memory (byte addresses a[]) until all carry `tag`; all loads of a round are in flight together
template <int N>
__device__ __forceinline__ void wait_lines(Ctx& c, const uint32_t (&a)[N], uint32_t tag, float2 (&out)[N]) {
    ACCT_BEGIN()
    uint32_t spins = 0;
    while (true) {
        uint4 r[N];
#pragma unroll
        for (int k = 0; k < N; ++k) r[k] = lds_line(a[k]);
        uint32_t bad = 0;
#pragma unroll
        for (int k = 0; k < N; ++k) {
            bad |= (r[k].y ^ tag) | (r[k].w ^ tag);
            out[k] = make_float2(__uint_as_float(r[k].x), __uint_as_float(r[k].z));
        }
        if (bad == 0 || DBG_LOCAL(c)) break;
        if (check_abort(c, spins)) break;
    }
#ifdef UMGEN_REREAD_SMEM      // experiment: is a remote 16-byte store ever seen half-written?
#pragma unroll
    for (int k = 0; k < N; ++k) { const uint4 r2 = lds_line(a[k]); out[k] = make_float2(__uint_as_float(r2.x), __uint_as_float(r2.z)); }
#endif
    ACCT_END(acc_x)
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// ---- tagged lines in L2 ----------------------------------------------------------------------------
__device__ __forceinline__ void ll_store2(float* base, int line, float v0, float v1, uint32_t tag) {
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(base + 4 * line), "r"(__float_as_uint(v0)), "r"(tag),
                 "r"(__float_as_uint(v1)), "r"(tag)
                 : "memory");
}
__device__ __forceinline__ uint4 ll_ld(const float* p) {
    uint4 r;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
// One L2 hop of a residual update: thread t polls the 8 clusters' partials of its OWN rows 2t, 2t+1 (published by rank t / 48 of every cluster)
// straight from L2 and adds them in cluster order.  8x the poll traffic of "one rank sums a slice and fans it out over DSMEM" (the round-1
// design), one exchange less on the critical path.  Scenes are polled one after the other (8 loads in flight; 16 lines = 64 registers would
// spill): the first scene's poll absorbs the wait for the slowest cluster, the others cost one L2 round trip each.  Out of line: two hops per
// layer and scene share one copy of the code (instruction cache, see check_abort_slow).
// src: line of cluster 0 for my rows; the other clusters' slots follow at CSTRIDE floats
constexpr size_t HOP_CSTRIDE = (size_t)CL * LINES_X * 4;      // floats between two clusters' slots
constexpr size_t HOP_SSTRIDE = (size_t)NCL * HOP_CSTRIDE;     // ... between two scenes
#ifdef UMGEN_HOP_INLINE
__device__ __forceinline__
#else
__device__ __noinline__
#endif
float2 hop_poll(const float* src, uint32_t tag) {
    uint4 v[NCL];
#if UMGEN_POLL_DELAY > 0
    { const long long t0 = clock64(); while (clock64() - t0 < UMGEN_POLL_DELAY) {} }
#endif
#pragma unroll
    for (int cc = 0; cc < NCL; ++cc) v[cc] = ll_ld(src + cc * HOP_CSTRIDE);
    uint32_t spins = 0;
    while (true) {
        uint32_t bad = 0;
#pragma unroll
        for (int cc = 0; cc < NCL; ++cc) bad |= (v[cc].y ^ tag) | (v[cc].w ^ tag);
        if (bad == 0 || DBG_LOCAL_SM()) break;
        if (((++spins) & 0x3ffu) == 0 && check_abort_slow()) break;
#if UMGEN_POLL_BACKOFF > 0
        { const long long t0 = clock64(); while (clock64() - t0 < UMGEN_POLL_BACKOFF) {} }
#endif
#pragma unroll
        for (int cc = 0; cc < NCL; ++cc)
            if ((v[cc].y ^ tag) | (v[cc].w ^ tag)) v[cc] = ll_ld(src + cc * HOP_CSTRIDE);
    }
#ifdef UMGEN_REREAD_L2
#pragma unroll
    for (int cc = 0; cc < NCL; ++cc) v[cc] = ll_ld(src + cc * HOP_CSTRIDE);
#endif
    float v0 = 0.f, v1 = 0.f;
#pragma unroll
    for (int cc = 0; cc < NCL; ++cc) { v0 += __uint_as_float(v[cc].x); v1 += __uint_as_float(v[cc].z); }
    return make_float2(v0, v1);
}
// buf: [2 (layer parity)][NB][NCL][CL][48 lines][4 floats]; thread t's lines are those of rank t / 48, line t % 48 = flat index t
template <int NB>
__device__ __forceinline__ const float* hop_src(const Ctx& c, const float* buf, int s) {
    return buf + ((size_t)(c.lc & 1u) * NB + s) * HOP_SSTRIDE + (size_t)c.tid * 4;
}
template <int NB>
__device__ __forceinline__ float* partial_slot(Ctx& c, int buf_off, int s) {
    return c.scratch + buf_off + ((((size_t)(c.lc & 1u) * NB + s) * NCL + c.h) * CL + c.i) * (LINES_X * 4);
}

// ---- ring ------------------------------------------------------------------------------------------
// Stages are released per warp: empty[] counts N_CONS_WARPS arrivals, lane 0 of every consumer warp arrives once the warp has read the stage
// for the last time (no block barrier on the release path).
template <int NB>
__device__ __forceinline__ const uint8_t* acquire(Ctx& c, uint32_t bytes, Stage& st, uint32_t reserve = 0) {
    st = ring_next<ring_bytes<NB>()>(c.ring, bytes, reserve);
    ACCT_BEGIN()
    wait_mbar(c, &SM<NB>()->full[st.slot], st.parity);
    ACCT_END(acc_ring)
    return SM<NB>()->ring + st.off;
}
template <int NB, int SITE = -1>
__device__ __forceinline__ void release(Ctx& c) {   // the warp's reads of the stage it acquired last are complete (stages are held one at a time)
#ifdef UMGEN_REL_SYNC      // debug: block barrier before the arrivals (of site UMGEN_REL_SYNC, or of every site when it is 99)
    if (UMGEN_REL_SYNC == 99 || UMGEN_REL_SYNC == SITE) cons_sync();
#endif
    __syncwarp();
    if (c.lane == 0) mbar_arrive(&SM<NB>()->empty[(c.ring.k - 1u) % NSLOT]);
}
template <int NB>
struct Producer {
    uint32_t tail = 0;   // oldest stage not known to be released
    uint32_t chunk = 0;  // > 0: a stage is fetched as bulk copies of at most `chunk` bytes, issued at least `gap` cycles apart (debug knobs, see decode_cluster_kernel)
    uint32_t gap = 0;
    bool tiny = false;   // debug (args.grid bit 2): fetch 16 bytes per stage only (wrong results, isolates the cost of the weight traffic)
    long long t_next = 0;
    __device__ __forceinline__ void copy(SmemT<NB>* sm, uint32_t off, const uint8_t* src, uint32_t bytes, uint64_t* bar) {
        if (chunk == 0) { bulk_g2s(sm->ring + off, src, bytes, bar); return; }
        for (uint32_t o = 0; o < bytes; o += chunk) {
            while (clock64() < t_next) {}
            bulk_g2s(sm->ring + off + o, src + o, min(chunk, bytes - o), bar);
            t_next = clock64() + gap;
        }
    }
    // one stage = `nsrc` runs of bytes / nsrc each (nsrc = 2: the K or V tiles of my two heads)
    __device__ __forceinline__ void issue(Ctx& cx, const void* src, uint32_t bytes, const uint8_t* const* srcs = nullptr, int nsrc = 1, uint32_t reserve = 0) {
        SmemT<NB>* sm = SM<NB>();
        Stage st = ring_next<ring_bytes<NB>()>(cx.ring, bytes, reserve);
        const uint32_t me = cx.ring.k - 1;
        // retire what the consumers have released meanwhile (non-blocking), so that the overlap scan below stays short
        while (tail < me && mbar_test_wait(&sm->empty[tail % NSLOT], (tail / NSLOT) & 1u)) tail++;
        while (true) {
            bool conflict = (me - tail) >= (uint32_t)NSLOT;
#pragma unroll 1
            for (uint32_t i = tail; i < me && !conflict; ++i) {
                uint32_t o = sm->fl_off[i % NSLOT], b = sm->fl_bytes[i % NSLOT];
                conflict = (st.off < o + b) && (o < st.off + bytes);
            }
            if (!conflict) break;
            wait_mbar(cx, &sm->empty[tail % NSLOT], (tail / NSLOT) & 1u);
            if (*(volatile int*)cx.abort_flag != 0) return;
            tail++;
        }
        sm->fl_off[me % NSLOT] = st.off;
        sm->fl_bytes[me % NSLOT] = bytes;
        if (tiny) {
            mbar_arrive_expect_tx(&sm->full[st.slot], 16);
            bulk_g2s(sm->ring + st.off, srcs ? srcs[0] : src, 16, &sm->full[st.slot]);
            return;
        }
        mbar_arrive_expect_tx(&sm->full[st.slot], bytes);
        if (srcs == nullptr) {
            copy(sm, st.off, (const uint8_t*)src, bytes, &sm->full[st.slot]);
        } else {
            for (int k = 0; k < nsrc; ++k) copy(sm, st.off + k * (bytes / nsrc), srcs[k], bytes / nsrc, &sm->full[st.slot]);
        }
    }
};
__device__ __forceinline__ void prefetch_l2(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// ---- cp.async staging of the small per-layer parameters ---------------------------------------------
__device__ __forceinline__ void cp_async16(void* s, const void* g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(s)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async4(void* s, const void* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(s)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// local c_attn row lr (0..35) of CTA (cluster h, rank i): head h*2 + lr/18, {q,k,v} = (lr%18)/6, element i*6 + lr%6 of the head
__device__ __forceinline__ int qkv_global_row(int h, int i, int lr) {
    const int hh = lr / 18, wr = lr - hh * 18, which = wr / 6, e = wr - which * 6;
    return which * C + (h * HPC + hh) * HD + i * 6 + e;
}
__device__ __forceinline__ void prefetch_params(Ctx& c, const float* fl, float* dst) {
    if (c.tid < 192) cp_async16(dst + PRM_LN1 + 4 * c.tid, fl + F_LN1 + 4 * c.tid);
    else cp_async16(dst + PRM_LN2 + 4 * (c.tid - 192), fl + F_LN2 + 4 * (c.tid - 192));
    if (c.tid < 192) cp_async16(dst + PRM_BPROJ + 4 * c.tid, fl + F_BPROJ + 4 * c.tid);
    if (c.tid >= 32 && c.tid < 32 + QKV_R) cp_async4(dst + PRM_BQKV + (c.tid - 32), fl + F_BQKV + qkv_global_row(c.h, c.i, c.tid - 32));
    cp_async_commit();
}

// ---- tensor-core GEMV pieces ---------------------------------------------------------------------------
// D[16x8] += A[16x16] . B[16x8]: A = one packed 512-byte fragment block (lane l reads bytes [16 l, 16 l + 16)),
// B columns 0 / 1 = the fp16 hi / lo parts of the fp32 input vector, so D[:, 0] + D[:, 1] is the row's dot product
__device__ __forceinline__ void mma16816(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
// raw shfl.sync for places where the compiler cannot prove warp uniformity (a __shfl_sync under such a branch costs a warp-sync call)
__device__ __forceinline__ float shfl_idx_raw(float v, int src) {
    float r;
    asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=f"(r) : "f"(v), "r"(src));
    return r;
}
__device__ __forceinline__ float ex2_approx(float x) {          // ex2(-inf) = 0
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
    return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}
// hi / lo fp16 split of a pair of fp32 values: hi = rn(x), lo = rn(x - hi); x ~ hi + lo to ~22 bits
__device__ __forceinline__ void split_hilo(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    // packed conversions (cvt.rn.f16x2.f32 = F2FP on the ALU pipe); the scalar F2F.F16.F32 runs on the slow conversion pipe
    const __half2 h = __floats2half2_rn(x0, x1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// Vectors that feed an MMA are kept in shared memory as ready-made B fragments, two columns per scene: f[k-step][lane] = {b0, b1} for lanes
// 8 s + 0..3 (column 2 s = hi of scene s) and 8 s + 4..7 (column 2 s + 1 = lo); lanes >= 8 NB hold zero columns and load nothing.  Values
// (2u, 2u+1) of a vector go to k-step u / 8, lane 8 s + u % 4 (+4 for lo), register (u % 8) / 4.  Column n of the product D[16x8] lands in the
// lanes with t = lane % 4 == n / 2: d[0] + d[1] (row g = lane / 4) and d[2] + d[3] (row g + 8) are the dot products of scene t.
template <int NB>
__device__ __forceinline__ void store_bfrag_pair(uint2* f, int s, int u, float x0, float x1) {
    uint32_t hi, lo;
    split_hilo(x0, x1, hi, lo);
    uint32_t* w = reinterpret_cast<uint32_t*>(f + (u >> 3) * (8 * NB) + 8 * s + (u & 3)) + ((u & 7) >> 2);
    w[0] = hi;
    w[8] = lo;            // lane + 4: 4 uint2 further
}
template <int NB>
__device__ __forceinline__ uint2 load_bfrag(const uint2* f, int ks, int lane) {
    uint2 b = make_uint2(0u, 0u);
    if (NB == 4 || lane < 8 * NB) b = f[ks * (8 * NB) + lane];
    return b;
}
// The per-scene parts of a layer are runtime loops (#pragma unroll 1), not unrolled copies: the layer loop must fit the instruction cache.
// Values that live in registers per scene are picked / updated with selects.
template <int NB>
__device__ __forceinline__ float2 pick(const float2 (&x)[NB], int s) {
    float2 r = x[0];
#pragma unroll
    for (int k = 1; k < NB; ++k) if (s == k) r = x[k];
    return r;
}
template <int NB>
__device__ __forceinline__ void put(float2 (&x)[NB], int s, float2 v) {
#pragma unroll
    for (int k = 0; k < NB; ++k) if (NB == 1 || s == k) x[k] = v;
}
// rows x 768 GEMV split along K over the 12 warps: warp w multiplies k-steps [4w, 4w+4) into NT full 16-row tiles (+ a 4-row tile
// when REM) and leaves its partial sums in sm->pq[w][scene][row].  `wp` = this warp's part of the matrix in fragment order:
// [k-step][tile][512 B] (+128 B for the 4-row tile)
template <int NB, int NT, bool REM>
__device__ __forceinline__ void gemv_ksplit(const Ctx& c, const uint8_t* wp, const uint2* xf) {
    SmemT<NB>* sm = SM<NB>();
    constexpr uint32_t KS_BYTES = NT * 512 + (REM ? 128 : 0);
    float acc[NT + 1][4];
#pragma unroll
    for (int m = 0; m <= NT; ++m) { acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < KS_PER_WARP; ++ks) {
        const uint2 b = load_bfrag<NB>(xf, c.warp * KS_PER_WARP + ks, c.lane);
#pragma unroll
        for (int m = 0; m < NT; ++m) {
            const uint4 a = *reinterpret_cast<const uint4*>(wp + ks * KS_BYTES + m * 512 + c.lane * 16);
            mma16816(acc[m], a, b.x, b.y);
        }
        if (REM) {       // rows 16 NT .. 16 NT + 3: lanes 0..15 hold {a0, a2}, rows g + 8 do not exist
            uint2 ar = make_uint2(0u, 0u);
            if (c.lane < 16) ar = *reinterpret_cast<const uint2*>(wp + ks * KS_BYTES + NT * 512 + c.lane * 8);
            mma16816(acc[NT], make_uint4(ar.x, 0u, ar.y, 0u), b.x, b.y);
        }
    }
    const int t = c.lane & 3;
    if (t < NB) {
        const int g = c.lane >> 2;
#pragma unroll
        for (int m = 0; m < NT; ++m) {
            sm->pq[c.warp][t][m * 16 + g] = acc[m][0] + acc[m][1];
            sm->pq[c.warp][t][m * 16 + g + 8] = acc[m][2] + acc[m][3];
        }
        if (REM && g < 4) sm->pq[c.warp][t][NT * 16 + g] = acc[NT][0] + acc[NT][1];
    }
}
template <int NB>
__device__ __forceinline__ float sum_pq(const SmemT<NB>* sm, int s, int row) {
    float r = 0.f;
#pragma unroll
    for (int w = 0; w < N_CONS_WARPS; ++w) r += sm->pq[w][s][row];
    return r;
}
// LayerNorm (module.py:26-37: weight only, eps 1e-5) of the residual vectors, of which thread t holds elements 2t, 2t+1 in v[scene].
// FRAG: the result goes to sm->xf as MMA B fragments, else to sm->out2 as fp32 (the head's input).  gw = weight in shared memory.  One block barrier for all scenes.
template <int NB, bool FRAG, int SB = -1>
__device__ UMGEN_INLINE void layer_norm(Ctx& c, const float2 (&v)[NB], const float* gw) {
    SmemT<NB>* sm = SM<NB>();
#pragma unroll 1
    for (int s = 0; s < NB; ++s) {
        const float2 vs = pick<NB>(v, s);
        float a = vs.x + vs.y, q = fmaf(vs.x, vs.x, vs.y * vs.y);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (c.lane == 0) { sm->lnred[s][c.warp] = a; sm->lnred[s][32 + c.warp] = q; }
    }
    const float2 g = reinterpret_cast<const float2*>(gw)[c.tid];
    if (SB >= 0) { STAMP(SB) }
    if (FRAG && SB < 0) { PROBE(19) }
    cons_sync();
    if (FRAG && SB < 0) { PROBE(20) }
#pragma unroll 1
    for (int s = 0; s < NB; ++s) {
        const float2 vs = pick<NB>(v, s);
        float ts = 0.f, tq = 0.f;
#pragma unroll
        for (int w = 0; w < N_CONS_WARPS; ++w) { ts += sm->lnred[s][w]; tq += sm->lnred[s][32 + w]; }
        if (SB >= 0 && s == 0) { if (ts == 12345.f) tq += 1.f; STAMP(SB + 1) }
        const float mean = ts * (1.0f / C);
        const float var = fmaxf(tq * (1.0f / C) - mean * mean, 0.f);
        const float rstd = rsqrtf(var + 1e-5f);
        const float y0 = (vs.x - mean) * rstd * g.x, y1 = (vs.y - mean) * rstd * g.y;
        if (FRAG) {
            // thread t holds elements 2t, 2t+1, i.e. warp w holds k-steps 4w .. 4w+3 of the fragment buffer -- exactly the k-steps warp w multiplies
            // in gemv_ksplit, so the fragments never cross a warp and need no block barrier
            store_bfrag_pair<NB>(&sm->xf[0][0], s, c.tid, y0, y1);
        } else {
            reinterpret_cast<float2*>(sm->out2[s])[c.tid] = make_float2(y0, y1);
        }
    }
    if (FRAG) __syncwarp(); else cons_sync();
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

// byte offset of element (row, col) inside a 512-byte A-fragment block (see frag_pos in the packer)
__device__ __forceinline__ uint32_t frag_off(int row, int col) {
    const int lane = (row & 7) * 4 + ((col & 7) >> 1), reg = (row >> 3) + 2 * (col >> 3);
    return (uint32_t)(lane * 16 + reg * 4 + (col & 1) * 2);
}
// split-KV attention of my two heads at step j, layer l (module.py:214-227 with one query, causal) on the tensor cores, one scene after the
// other: warps 0..5 work on head 0, warps 6..11 on head 1.  My cache rows are r % 8 == i, kept in 16-key tiles of 1536 B: K tile = 3 A-fragment
// blocks [keys 16][dims 16 ds ..] (scores = K q), V tile = 3 blocks [dims 16 dt ..][keys 16] (o = V^T p).  The K tiles and the V tiles of a scene
// are one ring stage each (head 0 | head 1), held one at a time and released per warp.  The row appended this step belongs to rank j % 8: the
// warp that owns its tile patches it into the staged tiles and writes it to the cache.  Sends (m, l, o[48]) of both heads to every rank.
template <int NB>
__device__ UMGEN_INLINE void attention(Ctx& c, int l, int j) {
    constexpr int WPH = N_CONS_WARPS / HPC;            // warps per head
    SmemT<NB>* sm = SM<NB>();
    const uint32_t dtag = c.lc + 1;
    const int hh = c.warp / WPH, wh = c.warp - hh * WPH, head = c.h * HPC + hh;
    const int cnt = (j + 7 - c.i) >> 3;
    const int own = ((j & 7) == c.i) ? 1 : 0;
    const int total = cnt + own;
    const int ntile = (total + 15) >> 4;
    const uint32_t stage_bytes = (uint32_t)HPC * (uint32_t)ntile * KV_TILE_BYTES;
    const int g = c.lane >> 2, t = c.lane & 3;
    // cache append (module.py:209-210; values already fp16-rounded): local row cnt = key cnt % 16 of tile cnt / 16.  The appending warp patches the
    // staged tiles (its scores need the new key); the copy to the cache in global memory comes after the partials are on their way.
    const bool appender = own && wh == ((cnt >> 4) % WPH);
    const int app_tile = cnt >> 4, app_kk = cnt & 15;
    constexpr int MAXT = KV_TILES / WPH;               // a warp owns at most 3 tiles: two passes instead of an online softmax
    static_assert(KV_TILES % WPH == 0, "tiles per warp");
    PROBE(3)
#pragma unroll 1
    for (int s = 0; s < NB; ++s) {
        // q as B fragments (3 k-steps of 16 dims), pre-scaled by 1/sqrt(48) * log2(e) (module.py:196-198)
        // elements 2e, 2e+1 (e < 24) of q / k / v (which = 0 / 1 / 2) of this head: line [rank e / 3][hh * 9 + which * 3 + e % 3]
        const uint4* ql = &sm->qkvl[s][0][hh * 9];
        uint32_t qb0[3], qb1[3];
        {
            const float qscale = 0.14433756729740643f * 1.4426950408889634f;
            uint32_t qa[6];
            float2 qv[6];
#pragma unroll
            for (int ds = 0; ds < 3; ++ds) {
                const int e0 = ds * 8 + t, e8 = e0 + 4;           // pairs (2 e0, 2 e0 + 1) and (2 e8, 2 e8 + 1) = dims 16 ds + 2t (+8)
                qa[2 * ds] = smem_u32(ql + (e0 / 3) * (QKV_R / 2) + e0 % 3);
                qa[2 * ds + 1] = smem_u32(ql + (e8 / 3) * (QKV_R / 2) + e8 % 3);
            }
            PROBE(24)
            wait_lines<6>(c, qa, dtag, qv);        // every lane polls (lanes with g >= 2 discard the values): no divergence around the loop
            PROBE(25)
#pragma unroll
            for (int ds = 0; ds < 3; ++ds) {
                const float2 x01 = qv[2 * ds], x89 = qv[2 * ds + 1];
                uint32_t h01, l01, h89, l89;
                split_hilo(x01.x * qscale, x01.y * qscale, h01, l01);
                split_hilo(x89.x * qscale, x89.y * qscale, h89, l89);
                qb0[ds] = (g == 0) ? h01 : ((g == 1) ? l01 : 0u);
                qb1[ds] = (g == 0) ? h89 : ((g == 1) ? l89 : 0u);
            }
        }
        __half2 app_k = __floats2half2_rn(0.f, 0.f);
        __half app_v0 = __float2half_rn(0.f), app_v1 = app_v0;
        if (appender && c.lane < HD / 2) {      // lane e: K[key kk][dims 2e, 2e+1] and V^T[dims 2e, 2e+1][key kk]
            const int e = c.lane;
            const uint32_t kva[2] = {smem_u32(ql + (e / 3) * (QKV_R / 2) + 3 + e % 3), smem_u32(ql + (e / 3) * (QKV_R / 2) + 6 + e % 3)};
            float2 kvv[2];
            wait_lines<2>(c, kva, dtag, kvv);
            app_k = __floats2half2_rn(kvv[0].x, kvv[0].y);
            app_v0 = __float2half_rn(kvv[1].x);
            app_v1 = __float2half_rn(kvv[1].y);
        }
        PROBE(26)
        // ---- scores: all tiles first (independent MMA chains); they live in the lanes with t == 0 (keys g and g + 8)
        float sa[MAXT], sb[MAXT];
        {
            Stage stk;
            const uint8_t* ks = nullptr;
            if (ntile > 0) {
                ks = acquire<NB>(c, stage_bytes, stk) + (size_t)hh * ntile * KV_TILE_BYTES;
                if (appender) {
                    if (c.lane < HD / 2) {
                        const int d = 2 * c.lane;
                        *reinterpret_cast<__half2*>(const_cast<uint8_t*>(ks) + (size_t)app_tile * KV_TILE_BYTES + (uint32_t)(d >> 4) * 512 + frag_off(app_kk, d & 15)) = app_k;
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // a later bulk copy may overwrite the patched tile
                    __syncwarp();
                }
            }
#pragma unroll
            for (int it = 0; it < MAXT; ++it) {
                const int tile = wh + it * WPH;
                sa[it] = -INFINITY; sb[it] = -INFINITY;
                if (tile < ntile) {                                    // warp-uniform
                    const uint8_t* kt = ks + (size_t)tile * KV_TILE_BYTES + c.lane * 16;
                    float sc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int ds = 0; ds < 3; ++ds) mma16816(sc, *reinterpret_cast<const uint4*>(kt + ds * 512), qb0[ds], qb1[ds]);
                    const int ka_i = tile * 16 + g;
                    if (t == 0 && ka_i < total) sa[it] = sc[0] + sc[1];
                    if (t == 0 && ka_i + 8 < total) sb[it] = sc[2] + sc[3];
                }
            }
            if (ntile > 0) release<NB, 1>(c);
        }
        PROBE(27)
        float m_run = -INFINITY;
#pragma unroll
        for (int it = 0; it < MAXT; ++it) m_run = fmaxf(m_run, fmaxf(sa[it], sb[it]));
#pragma unroll
        for (int o2 = 4; o2 < 32; o2 <<= 1) m_run = fmaxf(m_run, __shfl_xor_sync(0xffffffffu, m_run, o2));      // over the 8 lanes with my t
        m_run = __shfl_sync(0xffffffffu, m_run, 0);                // the t == 0 group holds the scores
        const float mref = (m_run == -INFINITY) ? 0.f : m_run;     // a warp without keys: every p is exp2(-inf) = 0
        PROBE(28)
        // ---- p and p.V per tile without rescaling
        float l_run = 0.f;
        float o[3][4];
#pragma unroll
        for (int dt = 0; dt < 3; ++dt) { o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f; }
        {
            Stage stv;
            const uint8_t* vs = nullptr;
            if (ntile > 0) {
                vs = acquire<NB>(c, stage_bytes, stv) + (size_t)hh * ntile * KV_TILE_BYTES;
                if (appender) {
                    if (c.lane < HD / 2) {
                        const int d = 2 * c.lane;
                        uint8_t* vt = const_cast<uint8_t*>(vs) + (size_t)app_tile * KV_TILE_BYTES;
                        *reinterpret_cast<__half*>(vt + (uint32_t)(d >> 4) * 512 + frag_off(d & 15, app_kk)) = app_v0;
                        *reinterpret_cast<__half*>(vt + (uint32_t)((d + 1) >> 4) * 512 + frag_off((d + 1) & 15, app_kk)) = app_v1;
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                }
            }
#pragma unroll
            for (int it = 0; it < MAXT; ++it) {
                const int tile = wh + it * WPH;
                if (tile < ntile) {
                    const uint8_t* vt = vs + (size_t)tile * KV_TILE_BYTES + c.lane * 16;
                    uint4 va[3];
#pragma unroll
                    for (int dt = 0; dt < 3; ++dt) va[dt] = *reinterpret_cast<const uint4*>(vt + dt * 512);
                    const float pa = ex2_approx(sa[it] - mref), pb = ex2_approx(sb[it] - mref);
                    l_run += pa + pb;
                    // p as a B fragment: lane (g, t) needs p[2t], p[2t+1] (b0) and p[2t+8], p[2t+9] (b1); p[k] lives in lane 4 (k % 8)
                    const float p0 = shfl_idx_raw(pa, 8 * t), p1 = shfl_idx_raw(pa, 8 * t + 4);        // tile < ntile is warp-uniform
                    const float p8 = shfl_idx_raw(pb, 8 * t), p9 = shfl_idx_raw(pb, 8 * t + 4);
                    uint32_t h01, l01, h89, l89;
                    split_hilo(p0, p1, h01, l01);
                    split_hilo(p8, p9, h89, l89);
                    const uint32_t pb0 = (g == 0) ? h01 : ((g == 1) ? l01 : 0u), pb1 = (g == 0) ? h89 : ((g == 1) ? l89 : 0u);
#pragma unroll
                    for (int dt = 0; dt < 3; ++dt) mma16816(o[dt], va[dt], pb0, pb1);
                }
            }
            if (ntile > 0) release<NB, 2>(c);
        }
        PROBE(29)
#pragma unroll
        for (int o2 = 4; o2 < 32; o2 <<= 1) l_run += __shfl_xor_sync(0xffffffffu, l_run, o2);      // lane 0: sum over the t == 0 group
        if (c.lane == 0) { wpart<NB>(sm, s, c.warp)[0] = m_run; wpart<NB>(sm, s, c.warp)[1] = l_run; }
        if (t == 0) {
#pragma unroll
            for (int dt = 0; dt < 3; ++dt) {
                wpart<NB>(sm, s, c.warp)[2 + dt * 16 + g] = o[dt][0] + o[dt][1];
                wpart<NB>(sm, s, c.warp)[2 + dt * 16 + g + 8] = o[dt][2] + o[dt][3];
            }
        }
        if (appender) {         // the new row -> my cache in global memory (read back by my own bulk copies from the next step on)
            uint8_t* kg = (uint8_t*)sm->kv_ptr[s] + ((((size_t)(l * 2 + 0) * NH + head) * CL + c.i) * KV_TILES + app_tile) * KV_TILE_BYTES;
            uint8_t* vg = (uint8_t*)sm->kv_ptr[s] + ((((size_t)(l * 2 + 1) * NH + head) * CL + c.i) * KV_TILES + app_tile) * KV_TILE_BYTES;
            if (c.lane < HD / 2) {
                const int d = 2 * c.lane;
                *reinterpret_cast<__half2*>(kg + (uint32_t)(d >> 4) * 512 + frag_off(app_kk, d & 15)) = app_k;
                *reinterpret_cast<__half*>(vg + (uint32_t)(d >> 4) * 512 + frag_off(d & 15, app_kk)) = app_v0;
                *reinterpret_cast<__half*>(vg + (uint32_t)((d + 1) >> 4) * 512 + frag_off((d + 1) & 15, app_kk)) = app_v1;
            }
        }
    }
    PROBE(30)
    cons_sync();
    PROBE(31)
    PROBE(4)
    REFRESH(c);
    // CTA partials = merge of each head's 6 warps.  Thread (head hm, rank r, line u >= 1) sends o[2u-2], o[2u-1] to rank r: 2 x 8 x 24 = 384 items,
    // one per thread and scene; the 16 threads with u == 1 also send line 0 = (m, l) (a second send, not a second pass over the warps' partials).
    {
        static_assert(HPC * CL * (PART_VALS / 2 - 1) == N_CONS, "one o-line per consumer thread");
        const int hm = c.tid / (N_CONS / HPC), rem = c.tid - hm * (N_CONS / HPC);
        const int r = rem / (PART_VALS / 2 - 1), u = 1 + rem - r * (PART_VALS / 2 - 1);
#pragma unroll 1
        for (int s = 0; s < NB; ++s) {
            float m = -INFINITY;
#pragma unroll
            for (int w = 0; w < WPH; ++w) m = fmaxf(m, wpart<NB>(sm, s, hm * WPH + w)[0]);
            float a0 = 0.f, a1 = 0.f, ls = 0.f;
#pragma unroll
            for (int w = 0; w < WPH; ++w) {
                const float2 ml = *reinterpret_cast<const float2*>(&wpart<NB>(sm, s, hm * WPH + w)[0]);
                const float f = (ml.x > -INFINITY) ? ex2_approx(ml.x - m) : 0.f;
                const float2 wv = *reinterpret_cast<const float2*>(&wpart<NB>(sm, s, hm * WPH + w)[2 * u]);
                a0 = fmaf(f, wv.x, a0);
                a1 = fmaf(f, wv.y, a1);
                ls = fmaf(f, ml.y, ls);
            }
            send_line(c, &sm->partl[s][hm][c.i][u], (uint32_t)r, a0, a1, dtag);
            if (u == 1) send_line(c, &sm->partl[s][hm][c.i][0], (uint32_t)r, m, ls, dtag);
        }
    }
    if (appender) {     // after the sends: this CTA is the one the whole cluster waits for
        fence_proxy_async_global();           // my later bulk copies (async proxy) must see the appended rows; kv_progress is published after the next barrier
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// the kernel: NB scenes decoded in lockstep (same position, same layer), two columns of every MMA's B operand per scene
// ------------------------------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(N_THREADS, 1) decode_cluster_kernel(const __grid_constant__ KParamsT<NB> p) {
    using SC = Scratch<NB>;
    SmemT<NB>* sm = SM<NB>();
    const UmgenDecodeArgs& a = p.a[0];            // weights, sampling set-up, n_steps, prefix_len: common to the scenes (checked by the host)
    Ctx c;
    c.a = p.a; c.abort_flag = (int*)a.status_i32;
    c.cta = blockIdx.x; c.tid = threadIdx.x; c.warp = threadIdx.x >> 5; c.lane = threadIdx.x & 31;
    {
        uint32_t rk, cid;
        asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rk));
        asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(cid));
        c.i = (int)rk; c.h = (int)cid;
    }
    c.lc = 0; c.probe = nullptr; c.probe_t0 = 0;
    c.dbg_local = (a.grid & 2) != 0;
    c.acct = (blockIdx.x == 0 && threadIdx.x == 0); c.acc_ring = 0; c.acc_x = 0; c.acc_poll = 0; c.acc_attn = 0; c.acc_head = 0;
    const long long t_start = clock64();
    const bool dbg_same_layer = (a.grid & 1) != 0;      // debug: stream layer 0's matrices for every layer (L2-resident weights)
    c.scratch = (float*)a.scratch_f;
    const int L = (int)a.n_layer;
    const int n_steps = (int)a.n_steps;
    const int g = c.h * CL + c.i;                 // CTA index in the packed weights

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSLOT; ++s) { mbar_init(&sm->full[s], 1); mbar_init(&sm->empty[s], N_CONS_WARPS); }
        sm->kv_progress = 0;
        sm->wi.abort_flag = (int*)a.status_i32;
        sm->wi.t_dead = globaltimer_ns() + TIMEOUT_NS;
        sm->wi.dbg_local = (a.grid & 2) != 0 ? 1 : 0;
#pragma unroll
        for (int s = 0; s < NB; ++s) { sm->sc[s].nbox = 0; sm->sc[s].tok = 0; sm->sc[s].stage = sm->stage; sm->kv_ptr[s] = p.a[s].kv_h; }
        mbar_fence_init();
    }
    {       // no line may carry a valid tag before the first exchange
        uint4* z = &sm->qkvl[0][0][0];
        constexpr int NZ = (sizeof(sm->qkvl) + sizeof(sm->partl)) / 16;
        static_assert(offsetof(SmemRest<NB>, partl) == offsetof(SmemRest<NB>, qkvl) + sizeof(sm->qkvl), "line buffers are contiguous");
        for (int k = threadIdx.x; k < NZ; k += N_THREADS) z[k] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    cluster_sync_all();          // every CTA's line buffers are clean before anyone sends
    c.tl = nullptr;
    c.tl_base = clock64();

    const uint8_t* Wc = (const uint8_t*)a.oar_cl_h;
    const float* Fl = (const float*)a.oar_f;
    const __half* heads[3] = {(const __half*)a.head_map_h, (const __half*)a.head_bbox_h, (const __half*)a.head_img_h};
    const float* emb_tables[3] = {(const float*)a.map_table_f, (const float*)a.be_f, (const float*)a.img_table_f};

    if (c.warp == N_CONS_WARPS) {
        // ============================== producer warp ==========================================
        if (c.lane == 0) {
            Producer<NB> pr;
            pr.tiny = (a.grid & 4) != 0;
            pr.chunk = (uint32_t)((a.grid >> 8) & 0xff) * 1024u;        // experiment knobs: args.grid bits 8..15 = chunk KB, bits 16..27 = cycles between chunks
            pr.gap = (uint32_t)((a.grid >> 16) & 0xfff);
#pragma unroll 1
            for (int j = 0; j < n_steps; ++j) {
                const int q = j + 1;
                const int cnt = (j + 7 - c.i) >> 3;
                const int ntile = (cnt + (((j & 7) == c.i) ? 1 : 0) + 15) >> 4;
#pragma unroll 1
                for (int l = 0; l < L; ++l) {
                    const uint8_t* wl = Wc + ((size_t)(dbg_same_layer ? 0 : l) * GRID + g) * CTA_LAYER_BYTES;
                    {       // the next layer's matrices start their trip HBM -> L2 now: the ring holds less than a layer, L2 takes the HBM latency and its jitter
                        const uint8_t* wn = Wc + ((size_t)((l + 1 == L) ? 0 : l + 1) * GRID + g) * CTA_LAYER_BYTES;
                        prefetch_l2(wn, B_QKV + B_PROJ);
                        prefetch_l2(wn + B_QKV + B_PROJ, B_FC);
                        prefetch_l2(wn + B_QKV + B_PROJ + B_FC, B_PROJ2);
                    }
#if UMGEN_DECODE_PROFILE == 1
                    const bool pprobe = (c.cta == 0 && l == 1 && j == 1200);
#define PPROBE(k) if (pprobe) ((int*)a.status_i32)[81 + (k)] = (int)(clock64() - t_start);
#else
#define PPROBE(k)
#endif
                    PPROBE(0)
                    pr.issue(c, wl, B_QKV);
                    PPROBE(1)
                    if (ntile > 0) {
                        if (cnt > 0) {
                            const uint32_t need = (uint32_t)((j - 1) * L + l + 1);       // my rows of step j - 1 in this layer
                            uint32_t spins = 0;
                            while (sm->kv_progress < need) {
                                if (check_abort(c, spins)) break;
                            }
                            fence_proxy_async_global();
                        }
                        // per scene its K tiles, then its V tiles (the order the consumers take them in); a stage = head 0 | head 1
#pragma unroll 1
                        for (int s = 0; s < NB; ++s)
#pragma unroll 1
                            for (int kv = 0; kv < 2; ++kv) {
                                const uint8_t* srcs[HPC];
#pragma unroll
                                for (int k = 0; k < HPC; ++k)
                                    srcs[k] = (const uint8_t*)sm->kv_ptr[s] + (((size_t)(l * 2 + kv) * NH + c.h * HPC + k) * CL + c.i) * (KV_TILES * KV_TILE_BYTES);
                                pr.issue(c, nullptr, (uint32_t)HPC * (uint32_t)ntile * KV_TILE_BYTES, srcs, HPC);
                            }
                    }
                    PPROBE(2)
                    const uint8_t* wp = wl + B_QKV;
                    pr.issue(c, wp, B_PROJ);
                    PPROBE(3)
                    wp += B_PROJ;
                    pr.issue(c, wp, B_FC, nullptr, 1, B_PROJ2);
                    PPROBE(4)
                    wp += B_FC;
                    pr.issue(c, wp, B_PROJ2);
                    PPROBE(5)
                }
                if (needs_head(q) && q > (int)a.prefix_len) {
                    const int mod = pos_mod(q);
                    const int V = vocab_of(mod);
                    const int r0 = (V * g) / GRID, r1 = (V * (g + 1)) / GRID;
#pragma unroll 1
                    for (int r = r0; r < r1; r += HEAD_ROWS)
                        pr.issue(c, (const uint8_t*)heads[mod] + (size_t)r * (C * 2), (uint32_t)min(HEAD_ROWS, r1 - r) * C * 2);
                }
                if (*(volatile int*)c.abort_flag != 0) break;
            }
        }
    } else {
        // ============================== consumer warps =========================================
        float* scratch = c.scratch;

        // ln_oar and the first layer's parameters -> shared memory; the zero columns of the fragment buffers stay zero
        if (c.tid < 192) cp_async16(sm->lno + 4 * c.tid, (const float*)a.ln_oar_f + 4 * c.tid);
        prefetch_params(c, Fl, sm->prm);
        // The residual vectors live in registers: thread t of every CTA holds elements 2t, 2t+1 of every scene (all CTAs compute identical values).
        // Input of step 0: task embedding + TAR feature of index 0 (UMGen.py:1175,1215,1231)
        float2 x[NB];
#pragma unroll
        for (int s = 0; s < NB; ++s) {
            const float2 t0 = __ldg(reinterpret_cast<const float2*>(a.tske_f) + c.tid), t1 = __ldg(reinterpret_cast<const float2*>(p.a[s].tar_feat_f) + c.tid);
            x[s] = make_float2(t0.x + t1.x, t0.y + t1.y);
        }
        if (c.cta == 0 && c.tid < 8) {
            const int qs[8] = {1, 5, 6, 1031, 1032, 1693, 1694, 2207};
#pragma unroll
            for (int s = 0; s < NB; ++s) {
                int* out_tokens = (int*)p.a[s].out_tokens_i32;
                int* picks = (int*)p.a[s].picks_i32;
                const int* pose_tok = (const int*)p.a[s].pose_tok_i32;
                out_tokens[qs[c.tid] - 1] = forced_id(qs[c.tid]);
                picks[qs[c.tid] - 1] = forced_id(qs[c.tid]);
                if (c.tid < 3) { out_tokens[1 + c.tid] = pose_tok[c.tid]; picks[1 + c.tid] = pose_tok[c.tid]; }
            }
        }
        cp_async_wait_all();
        cons_sync();

#pragma unroll 1
        for (int j = 0; j < n_steps; ++j) {
            const int q = j + 1;
            // TAR feature of the next position, fetched a whole step ahead of its use
            if (j + 1 == TAR_LATE_ROW0) {
                // the bbox3d rows of tar_feat and the TAR-head logits may be produced by kernels running beside this one (box_tar pass on the SMs
                // this kernel leaves free): wait for the host's signal before the first of them is read
                bool any = false;
#pragma unroll
                for (int s = 0; s < NB; ++s) any |= p.a[s].tar_ready_i32 != nullptr;
                if (any) {
                    if (c.tid == 0) {
#pragma unroll 1
                        for (int s = 0; s < NB; ++s) {
                            if (p.a[s].tar_ready_i32 == nullptr) continue;
                            uint32_t spins = 0;
                            while (ld_acquire_gpu((const uint32_t*)p.a[s].tar_ready_i32) != (uint32_t)p.a[s].tar_ready_value) {
                                if (check_abort(c, spins)) break;
                            }
                        }
                    }
                    cons_sync();
                }
            }

#pragma unroll 1
            for (int l = 0; l < L; ++l) {
#if UMGEN_DECODE_PROFILE
                c.probe = nullptr;
                if (c.tid == UMGEN_PROBE_TID && l == 1 && j == 1200 && (c.cta == 0 || c.cta == 37)) {
                    c.probe = (int*)a.status_i32 + (c.cta == 0 ? 8 : 40);
                    c.probe_t0 = clock64();
                    if (c.cta == 0) ((int*)a.status_i32)[80] = (int)(c.probe_t0 - t_start);      // same clock as the producer's stamps [81..86]
                }
#endif
#if UMGEN_DECODE_PROFILE == 3
                c.tl = (a.debug_u64 && c.lane == 0 && l == 1 && j == 1200) ? (long long*)a.debug_u64 + (c.cta * N_CONS_WARPS + c.warp) * 16 : nullptr;
#endif
                STAMP(0)
                REFRESH(c);
#ifdef UMGEN_PAD_INSTRS      // experiment: executed filler instructions in the layer loop (how sensitive is the kernel to the size of its loop body?)
                {
                    uint32_t d0 = c.tid, d1 = c.lane, d2 = c.warp, d3 = c.lc;
#pragma unroll
                    for (int k = 0; k < UMGEN_PAD_INSTRS / 4; ++k)
                        asm volatile("add.u32 %0, %0, 1;\n\tadd.u32 %1, %1, 1;\n\tadd.u32 %2, %2, 1;\n\tadd.u32 %3, %3, 1;" : "+r"(d0), "+r"(d1), "+r"(d2), "+r"(d3));
                    if ((d0 ^ d1 ^ d2 ^ d3) == 0xdeadbeefu) c.lc++;
                }
#endif
                const float* prm = sm->prm;

                const uint32_t dtag = c.lc + 1;        // tag of this layer's lines inside the cluster
                // ---- LN1 -> my 36 rows of c_attn (+bias) -> all-gather q|k|v of my heads (module.py:206)
                layer_norm<NB, true>(c, x, prm + PRM_LN1);
                PROBE(0)
                STAMP(1)
                {
                    Stage s0;
                    const uint8_t* w0 = acquire<NB>(c, B_QKV, s0);
                    PROBE(1)
                    gemv_ksplit<NB, 2, true>(c, w0 + (size_t)c.warp * QKV_WARP_BYTES, &sm->xf[0][0]);
                    PROBE(21)
                    release<NB>(c);
                    cons_sync();
                    PROBE(22)
                    PROBE(23)
                    REFRESH(c);
                    // item (scene s, rank r, rows 2 ln, 2 ln + 1): reduce the 12 K-slices, add the bias, send to rank r
                    constexpr int ITEMS = (QKV_R / 2) * CL;
#pragma unroll 1
                    for (int it0 = 0; it0 < NB * ITEMS; it0 += N_CONS) {
                        const int it = it0 + c.tid;
                        if (it < NB * ITEMS) {
                            const int s = it / ITEMS, rem = it - s * ITEMS, r = rem / (QKV_R / 2), ln = rem - r * (QKV_R / 2);
                            float v0 = sum_pq<NB>(sm, s, 2 * ln) + prm[PRM_BQKV + 2 * ln], v1 = sum_pq<NB>(sm, s, 2 * ln + 1) + prm[PRM_BQKV + 2 * ln + 1];
                            if ((ln % 9) >= 3) { v0 = __half2float(__float2half_rn(v0)); v1 = __half2float(__float2half_rn(v1)); }   // k, v live at cache precision
                            send_line(c, &sm->qkvl[s][c.i][ln], (uint32_t)r, v0, v1, dtag);
                        }
                    }
                }
                PROBE(2)
                STAMP(2)
                // ---- split-KV attention (each warp picks up q, the appending warp k and v, from the lines); partials all-gathered
#if UMGEN_DECODE_PROFILE == 1
                long long attn_t0 = 0;
                if (c.acct) attn_t0 = clock64();       // the attention path: cache tiles -> scores -> softmax -> P V -> partials merged (up to the c_proj input)
#endif
                REFRESH(c);
                int jl = j;
                asm volatile("" : "+r"(jl));           // (j & 7, j + 7, ... are layer-invariant: redone here, not kept in a spill slot)
                attention<NB>(c, l, jl);
                PROBE(5)
                STAMP(3)
                REFRESH(c);
                {       // thread u = 8 p + s: rank s's share of outputs 2p, 2p+1 (p < 48, same head); the 8 lanes of a group merge by butterfly
                    const int p2 = c.tid >> 3, rk = c.tid & 7, hh = p2 / (HD / 2), ln = 1 + (p2 - hh * (HD / 2));
#pragma unroll 1
                    for (int s = 0; s < NB; ++s) {
                        const uint32_t pa[2] = {smem_u32(&sm->partl[s][hh][rk][0]), smem_u32(&sm->partl[s][hh][rk][ln])};
                        float2 pv[2];
                        wait_lines<2>(c, pa, dtag, pv);
                        float m = pv[0].x;
#pragma unroll
                        for (int o = 1; o < 8; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                        const float f = (pv[0].x > -INFINITY) ? ex2_approx(pv[0].x - m) : 0.f;
                        float lsum = f * pv[0].y, o0 = f * pv[1].x, o1 = f * pv[1].y;
#pragma unroll
                        for (int o = 1; o < 8; o <<= 1) {
                            lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
                            o0 += __shfl_xor_sync(0xffffffffu, o0, o);
                            o1 += __shfl_xor_sync(0xffffffffu, o1, o);
                        }
                        if (rk == 0) {
                            const float inv = 1.0f / lsum;
                            store_bfrag_pair<NB>(&sm->yf[0][0], s, p2, o0 * inv, o1 * inv);
                        }
                    }
                }
                cons_sync();
#if UMGEN_DECODE_PROFILE == 1
                if (c.acct) c.acc_attn += clock64() - attn_t0;
#endif
                if (c.tid == 0) sm->kv_progress = c.lc + 1;       // after a barrier that follows attention(): the appended rows are written and fenced
                PROBE(13)
                STAMP(4)
                // ---- c_proj split along K: my 96 rows x my heads' 96 columns -> partial sums into L2 (module.py:227-229)
                const uint32_t tagP = 2u * c.lc + 1u;          // tags of the two L2 hops of this layer (each buffer sees every tag once)
                REFRESH(c);
                {
                    Stage st;
                    const uint8_t* w = acquire<NB>(c, B_PROJ, st);
                    PROBE(14)
                    if (c.warp < XS / 16) {            // warp w: rows [16 w, 16 w + 16), 6 k-steps; fragment blocks [tile][k-step][512 B]
                        float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
                        for (int ks = 0; ks < HPC * HD / 16; ++ks) {
                            const uint2 b = load_bfrag<NB>(&sm->yf[0][0], ks, c.lane);
                            const uint4 af = *reinterpret_cast<const uint4*>(w + ((size_t)(c.warp * (HPC * HD / 16) + ks) * 32 + c.lane) * 16);
                            mma16816(acc[ks & 1], af, b.x, b.y);
                        }
                        // lane (g, t) holds rows g (acc0 + acc1) and g + 8 (acc2 + acc3) of scene t; rows 2p, 2p+1 travel in one line
                        const float lo = (acc[0][0] + acc[1][0]) + (acc[0][1] + acc[1][1]), hi = (acc[0][2] + acc[1][2]) + (acc[0][3] + acc[1][3]);
                        const float lo1 = __shfl_down_sync(0xffffffffu, lo, 4), hi1 = __shfl_down_sync(0xffffffffu, hi, 4);
                        const int t = c.lane & 3;
                        if ((c.lane & 4) == 0 && t < NB) {
                            const int gq = c.lane >> 2;       // even row g of the tile
                            float* slot = partial_slot<NB>(c, SC::GP, t);
                            ll_store2(slot, (c.warp * 16 + gq) >> 1, lo, lo1, tagP);
                            ll_store2(slot, (c.warp * 16 + gq + 8) >> 1, hi, hi1, tagP);
                        }
                    }
                    // (A block barrier here, not only the per-warp arrivals: without it the warps that have no c_proj rows run ahead into the L2 poll
                    // and the kernel's results stop being reproducible run to run -- measured, tests/test_decode_gpu.py; round 1 had the same barrier.)
                    cons_sync();
                    release<NB, 3>(c);
                }
                PROBE(6)
                STAMP(5)
                // residual (module.py:409)
                REFRESH(c);
                {
                    const float2 bp = reinterpret_cast<const float2*>(prm + PRM_BPROJ)[c.tid];
                    ACCT_BEGIN()
#pragma unroll 1
                    for (int s = 0; s < NB; ++s) {
                        const float2 r = hop_poll(hop_src<NB>(c, scratch + SC::GP, s), tagP);
                        const float2 xs = pick<NB>(x, s);
                        put<NB>(x, s, make_float2(xs.x + (r.x + bp.x), xs.y + (r.y + bp.y)));
                    }
                    ACCT_END(acc_poll)
                    STAMP(7)
                }
                PROBE(7)
                // ---- LN2 -> my 48 rows of c_fc -> erf-GELU (module.py:245-247); the hidden slice stays in this CTA
#if UMGEN_SMALL_CODE
                layer_norm<NB, true>(c, x, prm + PRM_LN2);
#else
                layer_norm<NB, true, 14>(c, x, prm + PRM_LN2);
#endif
                // the next layer's parameters start their trip now: LN2 was the last reader of this layer's (every thread read its share before the
                // barrier inside layer_norm), and they are needed ~8 000 cycles from here
                prefetch_params(c, Fl + (size_t)((l + 1 == L) ? 0 : l + 1) * LAYER_F, sm->prm);
                PROBE(8)
                STAMP(8)
                REFRESH(c);
                {
                    Stage s0;
                    const uint8_t* w0 = acquire<NB>(c, B_FC, s0, B_PROJ2);
                    PROBE(15)
                    gemv_ksplit<NB, 3, false>(c, w0 + (size_t)c.warp * FC_WARP_BYTES, &sm->xf[0][0]);
                    release<NB>(c);
                    cons_sync();
#pragma unroll 1
                    for (int it0 = 0; it0 < NB * (FC_R / 2); it0 += N_CONS) {
                        const int it = it0 + c.tid;
                        if (it < NB * (FC_R / 2)) {
                            const int s = it / (FC_R / 2), u = it - s * (FC_R / 2);
                            store_bfrag_pair<NB>(&sm->hf[0][0], s, u, gelu_erf(sum_pq<NB>(sm, s, 2 * u)), gelu_erf(sum_pq<NB>(sm, s, 2 * u + 1)));
                        }
                    }
                    cons_sync();
                }
                PROBE(9)
                STAMP(9)
                // ---- MLP c_proj split along K (module.py:248): all 768 rows x my 48 columns, reduce-scattered in the cluster.
                // warp w: row tiles [4w, 4w+4), 3 k-steps; fragment blocks [tile][k-step][512 B]
                REFRESH(c);
                {
                    Stage s0;
                    const uint8_t* w0 = acquire<NB>(c, B_PROJ2, s0);
                    PROBE(16)
                    const uint8_t* wmine = w0 + (size_t)c.warp * (4 * 3 * 512);
                    uint2 b[3];
#pragma unroll
                    for (int ks = 0; ks < 3; ++ks) b[ks] = load_bfrag<NB>(&sm->hf[0][0], ks, c.lane);
                    float acc[4][4];
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f;
#pragma unroll
                        for (int ks = 0; ks < 3; ++ks) {
                            const uint4 af = *reinterpret_cast<const uint4*>(wmine + ((size_t)(m * 3 + ks) * 32 + c.lane) * 16);
                            mma16816(acc[m], af, b[ks].x, b[ks].y);
                        }
                    }
                    release<NB, 5>(c);
                    const int t = c.lane & 3;
                    if (t < NB) {                      // lane (g, t) holds rows g and g + 8 of each tile for scene t
                        const int gq = c.lane >> 2;
#pragma unroll
                        for (int m = 0; m < 4; ++m) {
                            sm->out2[t][(c.warp * 4 + m) * 16 + gq] = acc[m][0] + acc[m][1];
                            sm->out2[t][(c.warp * 4 + m) * 16 + gq + 8] = acc[m][2] + acc[m][3];
                        }
                    }
                    cons_sync();
                    PROBE(17)
                    // thread u sends rows 2u, 2u+1 (= rows 2 (u % 48) of rank u / 48's slice) to rank u / 48
#pragma unroll 1
                    for (int s = 0; s < NB; ++s) {
                        const float2 ov = reinterpret_cast<const float2*>(sm->out2[s])[c.tid];
                        send_line(c, rsl<NB>(sm, s, c.i) + c.tid % LINES_X, (uint32_t)(c.tid / LINES_X), ov.x, ov.y, dtag | RSL_TAG);
                    }
                }
                PROBE(18)
                STAMP(10)
                const uint32_t tagR = 2u * c.lc + 2u;
                REFRESH(c);
                {       // thread u = 8 line + k: rank k's partial of rows 2 line, 2 line + 1 of my slice; butterfly sum -> one line in L2
                    const int line = c.tid >> 3, k = c.tid & 7;
#pragma unroll 1
                    for (int s = 0; s < NB; ++s) {
                        const uint32_t ra[1] = {smem_u32(rsl<NB>(sm, s, k) + line)};
                        float2 pv[1];
                        wait_lines<1>(c, ra, dtag | RSL_TAG, pv);
#pragma unroll
                        for (int o = 1; o < 8; o <<= 1) {
                            pv[0].x += __shfl_xor_sync(0xffffffffu, pv[0].x, o);
                            pv[0].y += __shfl_xor_sync(0xffffffffu, pv[0].y, o);
                        }
                        if (k == 0) ll_store2(partial_slot<NB>(c, SC::GR, s), line, pv[0].x, pv[0].y, tagR);
                    }
                }
                PROBE(10)
                STAMP(11)
                REFRESH(c);
                {         // residual (module.py:410)
                    ACCT_BEGIN()
#pragma unroll 1
                    for (int s = 0; s < NB; ++s) {
                        const float2 r = hop_poll(hop_src<NB>(c, scratch + SC::GR, s), tagR);
                        const float2 xs = pick<NB>(x, s);
                        put<NB>(x, s, make_float2(xs.x + r.x, xs.y + r.y));
                    }
                    ACCT_END(acc_poll)
                    STAMP(13)
                }
                PROBE(11)
                cp_async_wait_all();                   // my share of the next layer's parameters has landed (made visible by the next barrier)
                c.lc++;
            }

            // ---- head + sampling (UMGen.py:1247-1250, 1046-1137), every scene at the same position q
#if UMGEN_DECODE_PROFILE == 1
            long long head_t0 = 0;
            if (c.acct) head_t0 = clock64();
#endif
            int tok[NB];
            const int fid = forced_id(q);
            if (q <= 5 || fid >= 0 || q <= (int)a.prefix_len) {
#pragma unroll
                for (int s = 0; s < NB; ++s) {
                    if (q <= 5) tok[s] = (fid >= 0) ? fid : __ldg((const int*)p.a[s].pose_tok_i32 + (q - 2));
                    else if (fid >= 0) tok[s] = fid;
                    else tok[s] = __ldg((const int*)p.a[s].teacher_i32 + (q - 1));      // given prefix (UMGen.py:1184-1201): no head, no sampling, no rule check
                }
            } else {
                const int mod = pos_mod(q);
                const int V = vocab_of(mod);
                const int k = (int)(mod == 0 ? a.top_k_map : (mod == 1 ? a.top_k_bbox : a.top_k_img));
                const int r0 = (V * g) / GRID, r1 = (V * (g + 1)) / GRID;
                const uint32_t mine = (uint32_t)q;           // tag of this step's candidate / logit lines
                layer_norm<NB, false>(c, x, sm->lno);
                {
                    XRegs xr[NB];
#pragma unroll
                    for (int s = 0; s < NB; ++s) xr[s] = load_x(sm->out2[s], c.lane);
#pragma unroll 1
                    for (int r = r0; r < r1; r += HEAD_ROWS) {
                        const int nr = min(HEAD_ROWS, r1 - r);
                        Stage st;
                        const uint8_t* w = acquire<NB>(c, (uint32_t)nr * C * 2, st);
#pragma unroll 1
                        for (int rr = c.warp; rr < nr; rr += N_CONS_WARPS) {        // one row of 768 halves (row-major) . x of every scene
                            const uint4* wp = reinterpret_cast<const uint4*>(w + (size_t)rr * (C * 2)) + c.lane;
                            const uint4 w0 = wp[0], w1 = wp[32], w2 = wp[64];
#pragma unroll
                            for (int s = 0; s < NB; ++s) {
                                const float d = warp_sum(dot8(w0, xr[s].a[0], xr[s].b[0]) + dot8(w1, xr[s].a[1], xr[s].b[1]) + dot8(w2, xr[s].a[2], xr[s].b[2]));
                                if (c.lane == 0) sm->acc[s][r - r0 + rr] = d;
                            }
                        }
                        release<NB>(c);
                    }
                }
                cons_sync();
                {
                    bool dumped = false;
#pragma unroll 1
                    for (int s = 0; s < NB; ++s) {
                        if (p.a[s].logits_dump_f) {
                            float* dump = (float*)p.a[s].logits_dump_f + (size_t)(q - 1) * 8192;
                            for (int r = r0 + c.tid; r < r1; r += N_CONS) dump[r] = sm->acc[s][r - r0];
                            dumped = true;
                        }
                    }
                    if (dumped) cons_sync();       // the selecting warps overwrite acc
                }
                if (a.sample_topp) {
                    // ---- nucleus sampling: all-gather the logits through L2, every CTA samples identically (UMGen.py:915-965); one scene after the other
#pragma unroll 1
                    for (int s = 0; s < NB; ++s) {
                        const UmgenDecodeArgs& as = p.a[s];
                        SceneSm* ss = &sm->sc[s];
                        float* LG = scratch + SC::LOGIT + (size_t)s * (2 * 8192);
                        for (int r = r0 + c.tid; r < r1; r += N_CONS)
                            asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(LG + 2 * r), "r"(__float_as_uint(sm->acc[s][r - r0])), "r"(mine) : "memory");
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
                        const float u0 = philox_uniform(as.seed, (uint32_t)as.frame_index, (uint32_t)q, 0u);
                        int slot = block_topp_sample(ss, v, pm, inv_t, u0, c.tid);
                        // slot = tid' + s * N_CONS with s the thread-local position: id = 2 * (tid' + (s / 2) * N_CONS) + (s & 1)
                        int t = 2 * ((slot % N_CONS) + ((slot / N_CONS) >> 1) * N_CONS) + ((slot / N_CONS) & 1);
                        if (mod == 1) {
                            const int bidx = q - BBOX_FIRST_POS - 1;
                            const int prev = __ldg((const int*)as.prev_bbox_i32 + bidx);
                            const bool controlled = (as.control_mask >> ((q - BBOX_FIRST_POS) / 11)) & 1ull;
                            const float* row = (const float*)as.tar_bbox_logits_f + (size_t)bidx * 1028;
                            for (int pass = 0; pass < 2; ++pass) {
                                const bool go2 = pass == 0 ? controlled : (t == PAD_TOKEN && a.merge_ar_tar && prev != PAD_TOKEN);
                                if (!go2) continue;
#pragma unroll
                                for (int e = 0; e < TOPP_PER; ++e) {
                                    const int id = c.tid + e * N_CONS;
                                    v[e] = (id < 1028 && !(controlled && id == 1027)) ? __ldcg(row + id) : -INFINITY;
                                }
                                const float uu = philox_uniform(as.seed, (uint32_t)as.frame_index, (uint32_t)q, 1u + pass);
                                t = block_topp_sample(ss, v, (float)a.top_p_bbox, inv_t, uu, c.tid);
                                if (pass == 1 && c.cta == 0 && c.tid == 0) atomicAdd((int*)as.status_i32 + 2, 1);
                            }
                        }
                        bool wipe = false;
                        if (c.warp == 0) {
                            if (mod == 1) { t = bbox_rules(ss, as, c.lane, c.cta, q, t, 0.f, true); wipe = (t & WIPE_BIT) != 0; t &= ~WIPE_BIT; }
                            if (c.lane == 0) {
                                if (wipe && c.cta == 0)
                                    for (int e = 1; e <= 10; ++e) ((int*)as.out_tokens_i32)[q - 1 - e] = PAD_TOKEN;
                                if (wipe) for (int e = 1; e <= 10; ++e) ss->recent[(q - e) & 15] = PAD_TOKEN;
                                ss->tok = t;
                            }
                        }
                        cons_sync();
                        tok[s] = ss->tok;
                    }
                } else {
                    if (c.warp < NB) {     // warp s: local top-k of scene s in my slice -> candidate lines {val, tag, id, tag}, one copy per reader rank
                        const int s = c.warp;
                        float* cline = scratch + SC::CAND + (size_t)s * (CREP * CANDV) + (size_t)g * MAX_CAND * 4;
                        const int n = r1 - r0;
#pragma unroll 1
                        for (int r = 0; r < k; ++r) {
                            float bv = -INFINITY;
                            int bi = 0x7fffffff;
#pragma unroll 1
                            for (int e = c.lane; e < n; e += 32) {
                                const float vv = sm->acc[s][e];
                                if (vv > bv) { bv = vv; bi = e; }
                            }
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) {
                                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                            }
                            const int id = (bi == 0x7fffffff) ? 0x7fffffff : r0 + bi;
                            if (c.lane < CREP) ll_store2(cline + c.lane * CANDV, r, bv, __int_as_float(id), mine);
                            if (c.lane == 0 && bi != 0x7fffffff) sm->acc[s][bi] = -INFINITY;
                            __syncwarp();
                        }
                    }
                    // every CTA merges all candidates and decides the tokens identically, one scene after the other (one candidate buffer)
                    const int ncand = GRID * k;
#pragma unroll 1
                    for (int s = 0; s < NB; ++s) {
                        constexpr int N = (GRID * MAX_CAND + N_CONS - 1) / N_CONS;    // 3
                        const float* cbase = scratch + SC::CAND + (size_t)s * (CREP * CANDV) + (size_t)c.i * CANDV;
                        float* candv = sm->stage;
                        int* candi = reinterpret_cast<int*>(sm->stage + GRID * MAX_CAND);
                        uint4 r[N];
                        int src[N];
#pragma unroll
                        for (int t = 0; t < N; ++t) {
                            const int e = c.tid + t * N_CONS;
                            src[t] = -1;
                            if (e < ncand) {
                                const int cta_s = e / k;
                                src[t] = cta_s * MAX_CAND + (e - cta_s * k);
                                r[t] = ll_ld(cbase + 4 * src[t]);
                            }
                        }
#pragma unroll
                        for (int t = 0; t < N; ++t) {
                            if (src[t] >= 0) {
                                uint32_t spins = 0;
                                while (!(r[t].y == mine && r[t].w == mine)) {
                                    if (check_abort(c, spins)) break;
                                    r[t] = ll_ld(cbase + 4 * src[t]);
                                }
                                candv[c.tid + t * N_CONS] = __uint_as_float(r[t].x);
                                candi[c.tid + t * N_CONS] = (int)r[t].z;
                            }
                        }
                        cons_sync();
                        if (c.warp == 0) {
                            const UmgenDecodeArgs& as = p.a[s];
                            SceneSm* ss = &sm->sc[s];
                            const float u0 = philox_uniform(as.seed, (uint32_t)as.frame_index, (uint32_t)q, 0u);
                            int t = warp_topk_sample(candv, candi, ncand, k, 1.0f / (float)a.temperature, u0, c.lane);
                            bool wipe = false;
                            if (mod == 1) {
                                const float u2 = philox_uniform(as.seed, (uint32_t)as.frame_index, (uint32_t)q, 2u);
                                t = bbox_rules(ss, as, c.lane, c.cta, q, t, u2, false);
                                wipe = (t & WIPE_BIT) != 0;
                                t &= ~WIPE_BIT;
                            }
                            if (c.lane == 0) {
                                if (wipe && c.cta == 0)
                                    for (int e = 1; e <= 10; ++e) ((int*)as.out_tokens_i32)[q - 1 - e] = PAD_TOKEN;    // UMGen.py:1357-1365
                                if (wipe) for (int e = 1; e <= 10; ++e) ss->recent[(q - e) & 15] = PAD_TOKEN;
                                ss->tok = t;
                            }
                        }
                        cons_sync();
                    }
#pragma unroll
                    for (int s = 0; s < NB; ++s) tok[s] = sm->sc[s].tok;
                }
            }
#if UMGEN_DECODE_PROFILE == 1
            if (c.acct) c.acc_head += clock64() - head_t0;
#endif
            int tok_used[NB];
#pragma unroll
            for (int s = 0; s < NB; ++s) {
                if (DBG_LOCAL(c) || (a.grid & 4)) tok[s] = 0;          // these debug modes compute garbage: keep the table index in range
                const int* teacher = (const int*)p.a[s].teacher_i32;
                tok_used[s] = tok[s];
                if (teacher != nullptr && q > 5 && fid < 0 && (a.prefix_len == 0 || q <= (int)a.prefix_len)) tok_used[s] = __ldg(teacher + (q - 1));
                if (c.tid == 0) {
                    sm->sc[s].recent[q & 15] = tok_used[s];
                    if (c.cta == 0 && q > 5) { ((int*)p.a[s].out_tokens_i32)[q - 1] = tok_used[s]; ((int*)p.a[s].picks_i32)[q - 1] = tok[s]; }
                }
            }
            if (j == SEQ - 2) break;           // q = 2206 was the last sampled token; q = 2207 is forced

            // ---- the next input: embedding of the token + TAR feature of index j + 1 (UMGen.py:1046-1137, 1215-1231).
            // bos/eos -> axe, pose -> fouier_pe, map/image -> GMLP(codebook[tok]) (precomputed table), bbox3d -> be
#pragma unroll
            for (int s = 0; s < NB; ++s) {
                const float* row;
                if (forced_id(q) >= 0) row = (const float*)a.axe_f + (size_t)forced_id(q) * C;
                else if (q <= 5) row = (const float*)a.fpe_f + (size_t)tok_used[s] * C;
                else row = emb_tables[pos_mod(q)] + (size_t)tok_used[s] * C;
                const float2 e = __ldg(reinterpret_cast<const float2*>(row) + c.tid);
                // TAR feature of the next position (L2-coherent load: late rows are written by other kernels while this one runs).  One exposed round trip
                // per step (~0.2 % of it); staging it ahead took 3 KB of shared memory per scene away from the ring.
                const float2 tn = (j + 1 < SEQ) ? __ldcg(reinterpret_cast<const float2*>((const float*)p.a[s].tar_feat_f + (size_t)(j + 1) * C) + c.tid) : make_float2(0.f, 0.f);
                x[s] = make_float2(e.x + tn.x, e.y + tn.y);
            }
            if (*(volatile int*)c.abort_flag != 0) break;
        }
        cp_async_wait_all();
        if (c.cta == 0 && c.tid == 0) {
            int* st = (int*)a.status_i32;
            st[3] = n_steps;
            st[60] = (int)((clock64() - t_start) >> 10);      // kilo-cycles: total; with UMGEN_DECODE_PROFILE also ring waits, DSMEM waits, L2 polls
            st[61] = (int)(c.acc_ring >> 10); st[62] = (int)(c.acc_x >> 10); st[63] = (int)(c.acc_poll >> 10);
            st[64] = (int)(c.acc_attn >> 10); st[65] = (int)(c.acc_head >> 10);      // UMGEN_DECODE_PROFILE: attention path / head + sampling, kilo-cycles
            for (int s = 1; s < NB; ++s) {      // the abort word is scene 0's; the other scenes report the same outcome
                int* so = (int*)p.a[s].status_i32;
                so[3] = n_steps;
                so[60] = st[60];
                const int ab = *(volatile int*)c.abort_flag;
                if (ab != 0) so[0] = ab;
            }
        }
    }
    // nobody leaves while a peer may still write into its shared memory
    __syncwarp();
    cluster_sync_all();
}

}  // namespace cl
}  // namespace umgen

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
namespace umgen {
extern int64_t g_launches;

template <int NB>
static cudaError_t cluster_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attrs, bool coop, cudaStream_t stream) {
    const size_t smem = sizeof(cl::SmemT<NB>) + 128;
    cudaError_t e = cudaFuncSetAttribute(cl::decode_cluster_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    memset(cfg, 0, sizeof(*cfg));
    cfg->gridDim = dim3(cl::GRID);
    cfg->blockDim = dim3(N_THREADS);
    cfg->dynamicSmemBytes = smem;
    cfg->stream = stream;
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = cl::CL;
    attrs[0].val.clusterDim.y = 1;
    attrs[0].val.clusterDim.z = 1;
    attrs[1].id = cudaLaunchAttributeCooperative;
    attrs[1].val.cooperative = 1;
    cfg->attrs = attrs;
    cfg->numAttrs = coop ? 2 : 1;
    return cudaSuccess;
}

// number of 8-CTA clusters of the decode kernel that can be resident at once on the current device (8 are needed)
template <int NB>
static int cluster_capacity_nb() {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attrs[2];
    if (cluster_config<NB>(&cfg, attrs, false, nullptr) != cudaSuccess) { cudaGetLastError(); return 0; }
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, (const void*)cl::decode_cluster_kernel<NB>, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
int decode_cluster_capacity() { return cluster_capacity_nb<1>(); }
int decode_cluster_need() { return cl::NCL; }
int decode_cluster_max_scenes() { return cl::NB_MAX; }
int64_t decode_cluster_scratch_floats() { return cl::Scratch<cl::NB_MAX>::TOTAL; }

template <int NB>
static int cluster_launch_nb(const UmgenDecodeArgs* args, cudaStream_t stream) {
    const int cap = cluster_capacity_nb<NB>();
    if (cap < cl::NCL) { set_error("device can hold only %d of the %d clusters of 8 CTAs the cluster decode kernel needs", cap, cl::NCL); return -3; }
    cl::KParamsT<NB> kp;
    for (int s = 0; s < NB; ++s) kp.a[s] = args[s];
    UMGEN_CUDA_OK(cudaMemsetAsync(args[0].scratch_f, 0, cl::Scratch<NB>::TOTAL * sizeof(float), stream));
    for (int s = 0; s < NB; ++s) UMGEN_CUDA_OK(cudaMemsetAsync(args[s].status_i32, 0, 96 * sizeof(int), stream));
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attrs[2];
    void* kargs[] = {&kp};
    const char* no_coop = getenv("UMGEN_DECODE_NO_COOP");      // profilers that cannot replay cooperative cluster launches
    UMGEN_CUDA_OK(cluster_config<NB>(&cfg, attrs, !(no_coop && no_coop[0] == '1'), stream));
    cudaError_t e = cudaLaunchKernelExC(&cfg, (const void*)cl::decode_cluster_kernel<NB>, kargs);
    if (e != cudaSuccess) {          // cooperative + cluster refused: all clusters still fit (checked above) on an idle device
        cudaGetLastError();
        UMGEN_CUDA_OK(cluster_config<NB>(&cfg, attrs, false, stream));
        UMGEN_CUDA_OK(cudaLaunchKernelExC(&cfg, (const void*)cl::decode_cluster_kernel<NB>, kargs));
    }
    g_launches += 1;
    return 0;
}

// args[0 .. n_scenes): scenes decoded in lockstep by one launch (the caller has checked that they agree on everything but the per-scene fields)
int decode_cluster_launch(const UmgenDecodeArgs* args, int n_scenes, cudaStream_t stream) {
    for (int s = 0; s < n_scenes; ++s)
        if (!args[s].oar_cl_h) { set_error("cluster decode kernel needs oar_cl_h (umgen_pack_oar_cluster)"); return -1; }
    switch (n_scenes) {
        case 1: return cluster_launch_nb<1>(args, stream);
#if UMGEN_MAX_SCENES >= 2
        case 2: return cluster_launch_nb<2>(args, stream);
#endif
#if UMGEN_MAX_SCENES >= 3
        case 3: return cluster_launch_nb<3>(args, stream);
#endif
#if UMGEN_MAX_SCENES >= 4
        case 4: return cluster_launch_nb<4>(args, stream);
#endif
        default: set_error("the cluster decode kernel takes 1..%d scenes per launch (got %d)", cl::NB_MAX, n_scenes); return -1;
    }
}

// Element (row, col) of a 16x16 tile at position e (in halves) of its 512-byte mma.m16n8k16 A-fragment block:
// lane = e / 8, register = (e % 8) / 2, half = e % 2; row = lane / 4 + 8 (register & 1), col = 2 (lane % 4) + half + 8 (register / 2)
__device__ __forceinline__ void frag_pos(int e, int& row, int& col) {
    const int lane = e >> 3, reg = (e & 7) >> 1, hp = e & 1;
    row = (lane >> 2) + 8 * (reg & 1);
    col = 2 * (lane & 3) + hp + 8 * (reg >> 1);
}
// oar_h [L][c_attn | c_proj | c_fc | mlp c_proj] (row-major) -> oar_cl_h [L][64 CTAs][c_attn part | c_proj part | c_fc part | mlp c_proj part],
// every part in A-fragment order (layout documented in include/umgen.h); CTA g = cluster * 8 + rank
__global__ void pack_cluster_kernel(const __half* __restrict__ src, __half* __restrict__ dst, int n_layer) {
    constexpr int QKV_H = cl::B_QKV / 2, PROJ_H = cl::B_PROJ / 2, FC_H = cl::B_FC / 2;
    const size_t per_cta = cl::CTA_LAYER_BYTES / 2;
    const size_t total = (size_t)n_layer * cl::GRID * per_cta;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t l = idx / (cl::GRID * per_cta);
        const size_t rem = idx - l * (cl::GRID * per_cta);
        const int g = (int)(rem / per_cta);
        int o = (int)(rem - (size_t)g * per_cta);
        const int cluster = g / cl::CL, i = g % cl::CL;
        const __half* w = src + l * (size_t)UMGEN_OAR_LAYER_H;
        size_t s;
        int row, col;
        if (o < QKV_H) {                    // [warp 12][k-step 4][tile0 256 | tile1 256 | 4-row tile 64] halves
            const int wq = o / 2304, o1 = o % 2304, ksl = o1 / 576, o2 = o1 % 576;
            if (o2 < 512) {
                frag_pos(o2 % 256, row, col);
                row += (o2 / 256) * 16;
            } else {                        // lanes 0..15 x {a0, a2}
                const int e = o2 - 512, lane = e >> 2, rr = (e & 3) >> 1, hp = e & 1;
                row = 32 + (lane >> 2);
                col = 2 * (lane & 3) + hp + 8 * rr;
            }
            col += (wq * cl::KS_PER_WARP + ksl) * 16;
            const int hh = row / 18, wr = row % 18, which = wr / 6, e6 = wr % 6;
            s = (size_t)(which * C + (cluster * cl::HPC + hh) * HD + i * 6 + e6) * C + col;
        } else if ((o -= QKV_H) < PROJ_H) { // [tile 6][k-step 6][256]
            const int mt = o / (6 * 256), ks = (o / 256) % 6;
            frag_pos(o % 256, row, col);
            row += mt * 16; col += ks * 16;
            const int hh = col / HD, d = col % HD;
            s = (size_t)3 * C * C + (size_t)(i * cl::XS + row) * C + (cluster * cl::HPC + hh) * HD + d;
        } else if ((o -= PROJ_H) < FC_H) {  // [warp 12][k-step 4][tile 3][256]
            const int wq = o / 3072, o1 = o % 3072, ksl = o1 / 768, mt = (o1 % 768) / 256;
            frag_pos(o1 % 256, row, col);
            row += mt * 16; col += (wq * cl::KS_PER_WARP + ksl) * 16;
            s = (size_t)4 * C * C + (size_t)(g * cl::FC_R + row) * C + col;
        } else {                            // [tile 48][k-step 3][256]
            o -= FC_H;
            const int mt = o / 768, ks = (o % 768) / 256;
            frag_pos(o % 256, row, col);
            row += mt * 16; col += ks * 16;
            s = (size_t)4 * C * C + (size_t)FF * C + (size_t)row * FF + g * cl::FC_R + col;
        }
        dst[idx] = w[s];
    }
}
}  // namespace umgen

using namespace umgen;

extern "C" int umgen_decode_cluster_capacity(void) { return decode_cluster_capacity(); }

extern "C" int umgen_pack_oar_cluster(const void* oar_h, void* oar_cl_h, int64_t n_layer, void* stream_v) {
    if (!oar_h || !oar_cl_h || n_layer < 1) { set_error("bad arguments"); return -1; }
    pack_cluster_kernel<<<1184, 256, 0, (cudaStream_t)stream_v>>>((const __half*)oar_h, (__half*)oar_cl_h, (int)n_layer);
    UMGEN_CUDA_OK(cudaGetLastError());
    g_launches += 1;
    return 0;
}

// Lazy module loading (the CUDA 12 default) loads a kernel on its first launch and that load waits for an idle device -- which never comes while the
// persistent decode kernel spins on a flag.  umgen_preload() (capi.cu) forces every kernel of the library to load up front.
#define UMGEN_PRELOAD(k) UMGEN_CUDA_OK(cudaFuncGetAttributes(&fa_, k))
namespace umgen {
int preload_decode_cluster() {
    cudaFuncAttributes fa_;
    UMGEN_PRELOAD(cl::decode_cluster_kernel<1>); UMGEN_PRELOAD(pack_cluster_kernel);
#if UMGEN_MAX_SCENES >= 2
    UMGEN_PRELOAD(cl::decode_cluster_kernel<2>);
#endif
#if UMGEN_MAX_SCENES >= 3
    UMGEN_PRELOAD(cl::decode_cluster_kernel<3>);
#endif
#if UMGEN_MAX_SCENES >= 4
    UMGEN_PRELOAD(cl::decode_cluster_kernel<4>);
#endif
    return 0;
}
}  // namespace umgen
