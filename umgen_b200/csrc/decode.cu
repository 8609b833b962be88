// Persistent OAR decode kernel: one cooperative launch runs every single-token step of a frame.
//
// Replaces UMGen.infer_oar_net / sample_next_token / rule_based_constraint (reference
// models/UMGen.py:1151-1383) and BlockOAR (models/module.py:378-428).  Design:
//   * grid = one CTA per SM, 16 consumer warps + 1 producer warp per CTA
//   * every GEMV is split by output rows across CTAs; the CTA's fp16 row slices for the next
//     ~1.2 layers are streamed HBM -> shared memory ahead of time by the producer warp with 1-D bulk
//     copies (cp.async.bulk + mbarrier) into a byte ring, so weight traffic never waits on activations
//   * the already-written part of the KV cache for the CTA's (head, split) is prefetched the same way;
//     attention is split-KV with an online-softmax partial per CTA, merged in the c_proj phase
//   * phases are separated by a device-wide barrier (one release-add + acquire-spin per CTA)
//   * head GEMV -> per-CTA top-k candidates -> every CTA redundantly merges, samples (Philox),
//     applies the bbox3d rules (TAR-head resample, control slots, collision wipe) and embeds the token
#include "decode_shared.cuh"

namespace umgen {

constexpr uint32_t RING_BYTES = 176 * 1024;
constexpr uint32_t MAX_STAGE = 36864;            // 6 rows of 3072 halves / 24 rows of 768 halves / 384 KV rows
constexpr int NSLOT = 8;
constexpr int KV_BLOCK = N_CONS_WARPS * 16;      // keys per attention block (one K stage + one V stage)
constexpr int MAX_ROWS = 160;                    // max rows of any GEMV slice per CTA (grid >= 64)
constexpr int MAX_GRID = 160;
constexpr int PART_STRIDE = 52;                  // (m, l, o[48]) padded
constexpr uint64_t TIMEOUT_NS = 20ull * 1000 * 1000 * 1000;

// offsets inside one packed layer
constexpr int OFF_QKV = 0;
constexpr int OFF_PROJ = 3 * C * C;
constexpr int OFF_FC = OFF_PROJ + C * C;
constexpr int OFF_PROJ2 = OFF_FC + FF * C;
constexpr int LAYER_H = OFF_PROJ2 + C * FF;
constexpr int F_LN1 = 0, F_BQKV = C, F_BPROJ = C + 3 * C, F_LN2 = C + 3 * C + C, LAYER_F = 3 * C + 3 * C;
static_assert(LAYER_H == UMGEN_OAR_LAYER_H && LAYER_F == UMGEN_OAR_LAYER_F, "packing");

// scratch layout (floats).  Every exchanged vector lives in 16-byte "LL lines" {v0, tag, v1, tag}: the
// epoch tag travels with the data (NCCL LL style), so a reader simply polls the line until both tags
// match -- no membar, no device-wide barrier on the critical path.
constexpr int MAX_SPLIT = 10;
constexpr int PART_VALS = 50;                      // (m, l, o[48])
constexpr int KREP = 16;                           // replicas of every all-to-all vector: reader r polls copy r % KREP,
                                                   // so one L2 line is shared by <= ceil(grid / KREP) pollers
constexpr int XV = 2 * C;                          // floats of one 768-value LL vector
constexpr int SC_XB = 0;                           // [KREP] residual after the MLP / next-step input
constexpr int SC_XA = SC_XB + KREP * XV;           // [KREP] residual after attention
constexpr int SC_Y = SC_XA + KREP * XV;            // [KREP] attention output (merged by the head leaders)
constexpr int SC_H = SC_Y + KREP * XV;             // [KREP] MLP hidden (3072 values)
constexpr int SC_CAND = SC_H + KREP * 2 * FF;      // [KREP][MAX_GRID][16] lines {val, tag, id, tag}
constexpr int CANDV = 4 * MAX_GRID * MAX_CAND;
constexpr int SC_Q = SC_CAND + KREP * CANDV;       // q of the current layer (read by the <= MAX_SPLIT CTAs of its head)
constexpr int SC_KVN = SC_Q + XV;                  // k_new | v_new (fp16-rounded), read by the head leader
constexpr int SC_PART = SC_KVN + 2 * XV;           // split-KV partials [16][MAX_SPLIT][50], read by the head leader
constexpr int SC_LOGIT = SC_PART + 2 * NH * MAX_SPLIT * PART_VALS;    // [KREP][8192 values] AR logits (top-p mode only)
constexpr int LOGITV = 2 * 8192;
constexpr int SC_KVFLAG = SC_LOGIT + KREP * LOGITV;                   // [16] steps whose KV rows are published
constexpr int SC_TOTAL = SC_KVFLAG + 64;
constexpr int STAGE_FLOATS = NH * MAX_SPLIT * PART_VALS;            // 8000 floats: partials / hidden / candidates

struct KParams {
    UmgenDecodeArgs a;
    int grid;
    int nsplit;
};

struct __align__(128) Smem {
    uint8_t ring[RING_BYTES];
    float xs[C];                  // normalised input of the current GEMV
    float xraw[C];                // raw residual vector last read (owner rows reused for the residual add)
    float stage[STAGE_FLOATS];    // MLP hidden (3072) | split partials (16*ns*50) | candidates
    float acc[MAX_ROWS * 3];      // raw dot products (row, k-chunk)
    float wpart[N_CONS_WARPS][PART_STRIDE];
    float pw[N_CONS_WARPS][16];   // softmax weights of the warp's 16 keys
    float red[64];
    float qs[HD], knew[HD], vnew[HD];
    float corners[MAX_BOX][8];    // decoded boxes of this frame (UMGen.py:1183,1338)
    int box_dropped[MAX_BOX];     // x >= 63 (misc.py:475-481)
    int recent[16];               // last tokens by position & 15
    float code[16];
    uint64_t full[NSLOT];
    uint64_t empty[NSLOT];
    uint32_t fl_off[NSLOT];       // producer bookkeeping of in-flight stages
    uint32_t fl_bytes[NSLOT];
    volatile int tok;             // token decided for the current position
    int nbox;
};

extern __shared__ __align__(128) uint8_t smem_raw[];
__device__ __forceinline__ Smem* SM() { return reinterpret_cast<Smem*>(smem_raw); }


__device__ __forceinline__ void row_slice(int rows, int cta, int grid, int& r0, int& r1) {
    r0 = (rows * cta) / grid;
    r1 = (rows * (cta + 1)) / grid;
}
// number of KV splits used when `nold` rows are already cached (fewer partials while the cache is short)
__device__ __forceinline__ int splits_for(int nold, int nsplit) { return max(1, min(nsplit, (nold + 191) / 192)); }
__device__ __forceinline__ void kv_range(int nold, int ns, int s, int& k0, int& k1) {
    int chunk = (nold + ns - 1) / ns;
    k0 = min(nold, s * chunk);
    k1 = min(nold, k0 + chunk);
}

struct Ring {
    uint32_t head = 0, k = 0;
};
struct Stage {
    uint32_t off, slot, parity;
};
__device__ __forceinline__ Stage ring_next(Ring& r, uint32_t bytes) {
    if (r.head + bytes > RING_BYTES) r.head = 0;
    Stage s{r.head, r.k % NSLOT, (r.k / NSLOT) & 1u};
    r.head += bytes;
    r.k++;
    return s;
}

// ------------------------------------------------------------------------------------------------
// abortable waits
// ------------------------------------------------------------------------------------------------
struct Ctx {
    const KParams* p;
    int* abort_flag;     // status[0]
    int* probe;          // non-null on the probing thread while the probed layer runs
    unsigned long long* tl;   // per-CTA timeline row (debug), non-null on thread 0 during the probed layer
    long long probe_t0;
    int cta, tid, warp, lane;
    Ring ring;
    uint32_t epoch;      // phases published so far (tag of the most recent phase's outputs)
    float* scratch;
};
#define PROBE(i) if (c.probe) { c.probe[i] = (int)(clock64() - c.probe_t0); }
#define TPROBE(i) if (c.tl) { c.tl[i] = globaltimer_ns(); }

__device__ __noinline__ bool check_abort_slow(int* abort_flag, uint32_t epoch) {
    if (*(volatile int*)abort_flag != 0) return true;
    // a wait that spins this long (~2^27 polls) is a deadlock: flag it so every CTA drains
    atomicCAS(abort_flag, 0, 100 + (int)(epoch & 0xffff));
    return true;
}
__device__ __forceinline__ bool check_abort(Ctx& c, uint32_t& spins) {
    ++spins;
    if ((spins & 0xfffu) == 0) {
        if (*(volatile int*)c.abort_flag != 0) return true;
        if (spins >= (1u << 27)) return check_abort_slow(c.abort_flag, c.epoch);
    }
    return false;
}
__device__ __forceinline__ void wait_mbar(Ctx& c, uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (check_abort(c, spins)) return;
    }
}

// ---- LL lines -----------------------------------------------------------------------------------
// Strong (relaxed.gpu) accesses on purpose: weak .cg/.cv polls compile to the same LDG.E.STRONG.GPU opcode but ptxas
// may hoist or merge them (observed: the spin loops dead-locked into the abort guard), so they buy nothing.
__device__ __forceinline__ void ll_store1(float* base, int idx, float v, uint32_t tag) {     // value idx -> 8 bytes
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(base + 2 * idx), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
__device__ __forceinline__ void ll_store2(float* base, int line, float v0, float v1, uint32_t tag) {
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(base + 4 * line), "r"(__float_as_uint(v0)), "r"(tag),
                 "r"(__float_as_uint(v1)), "r"(tag)
                 : "memory");
}
// value idx of an all-to-all vector: one 8-byte store per replica
__device__ __forceinline__ void ll_store1_rep(float* base, int stride, int idx, float v, uint32_t tag) {
#pragma unroll
    for (int k = 0; k < KREP; ++k) ll_store1(base + k * stride, idx, v, tag);
}
__device__ __forceinline__ uint4 ll_ld(const float* base, int line) {
    uint4 r;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(base + 4 * line) : "memory");
    return r;
}
// poll line `line` until both tags equal `tag`
__device__ __forceinline__ void ll_load2(Ctx& c, const float* base, int line, uint32_t tag, float& v0, float& v1) {
    uint32_t spins = 0;
    uint4 r;
    while (true) {
        r = ll_ld(base, line);
        if (r.y == tag && r.w == tag) break;
        if (check_abort(c, spins)) break;
    }
    v0 = __uint_as_float(r.x);
    v1 = __uint_as_float(r.z);
}
// Pipelined poll of up to N lines per thread: lines first + t*N_CONS for t < N (those < nlines);
// every load is in flight before the first tag is checked.  dst2 receives line i at float2 index i.
template <int N>
__device__ __forceinline__ void ll_read_lines(Ctx& c, const float* base, int nlines, uint32_t tag, float* dst) {
    uint4 r[N];
#pragma unroll
    for (int t = 0; t < N; ++t) {
        const int line = c.tid + t * N_CONS;
        if (line < nlines) r[t] = ll_ld(base, line);
    }
#pragma unroll
    for (int t = 0; t < N; ++t) {
        const int line = c.tid + t * N_CONS;
        if (line < nlines) {
            uint32_t spins = 0;
            while (!(r[t].y == tag && r[t].w == tag)) {
                if (check_abort(c, spins)) break;
                r[t] = ll_ld(base, line);
            }
            reinterpret_cast<float2*>(dst)[line] = make_float2(__uint_as_float(r[t].x), __uint_as_float(r[t].z));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// stage acquire / release (consumer) and issue (producer)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ const uint8_t* acquire(Ctx& c, uint32_t bytes, Stage& st) {
    st = ring_next(c.ring, bytes);
    wait_mbar(c, &SM()->full[st.slot], st.parity);
    return SM()->ring + st.off;
}
__device__ __forceinline__ void release(Ctx& c, const Stage& st) {   // caller synced the consumer warps
    if (c.tid == 0) mbar_arrive(&SM()->empty[st.slot]);
}

struct Producer {
    uint32_t tail = 0;   // oldest stage not known to be released
    __device__ __forceinline__ void issue(Ctx& cx, const void* src, uint32_t bytes) {
        Smem* sm = SM();
        Stage st = ring_next(cx.ring, bytes);
        const uint32_t me = cx.ring.k - 1;
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
        mbar_arrive_expect_tx(&sm->full[st.slot], bytes);
        bulk_g2s(sm->ring + st.off, src, bytes, &sm->full[st.slot]);
    }
    __device__ __forceinline__ void issue_rows(Ctx& cx, const uint8_t* base, uint32_t row_bytes, int r0, int r1) {
        const int per = (int)(MAX_STAGE / row_bytes);
#pragma unroll 1
        for (int r = r0; r < r1; r += per) {
            int nr = min(per, r1 - r);
            issue(cx, base + (size_t)r * row_bytes, (uint32_t)nr * row_bytes);
        }
    }
};

// wait until head `h`'s KV rows of every step < `step` are published (see publish in the consumer)
__device__ __forceinline__ void wait_kv_published(Ctx& c, int h, int step) {
    const uint32_t* flag = (const uint32_t*)(c.scratch + SC_KVFLAG) + h;
    uint32_t spins = 0;
    while (ld_acquire_gpu(flag) < (uint32_t)step) {
        if (check_abort(c, spins)) return;
    }
    fence_proxy_async_global();
}

// ------------------------------------------------------------------------------------------------
// consumer math
// ------------------------------------------------------------------------------------------------
// Read the 768-value vector `buf` (LL lines tagged `tag`) into sm->xraw and its LayerNorm
// (module.py:26-37: weight only, eps 1e-5) into sm->xs.  Single statistics pass (sum, sum of squares).
__device__ __forceinline__ void read_ln(Ctx& c, const float* buf, uint32_t tag, const float* w) {
    Smem* sm = SM();
    float v0 = 0.f, v1 = 0.f;
    const bool mine = c.tid < C / 2;
    float2 g = make_float2(0.f, 0.f);
    if (mine) {
        g = __ldg(reinterpret_cast<const float2*>(w) + c.tid);      // issued before the poll: off the critical path
        ll_load2(c, buf, c.tid, tag, v0, v1);
        reinterpret_cast<float2*>(sm->xraw)[c.tid] = make_float2(v0, v1);
    }
    float s = v0 + v1, q = fmaf(v0, v0, v1 * v1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (c.lane == 0) { sm->red[c.warp] = s; sm->red[32 + c.warp] = q; }
    cons_sync();
    float ts = 0.f, tq = 0.f;
#pragma unroll
    for (int i = 0; i < 12; ++i) { ts += sm->red[i]; tq += sm->red[32 + i]; }
    const float mean = ts * (1.0f / C);
    const float var = fmaxf(tq * (1.0f / C) - mean * mean, 0.f);
    const float rstd = rsqrtf(var + 1e-5f);
    if (mine) reinterpret_cast<float2*>(sm->xs)[c.tid] = make_float2((v0 - mean) * rstd * g.x, (v1 - mean) * rstd * g.y);
    cons_sync();
}

// acc[(row_off + r) ] = W[r][:] . xs  for r in [0, nr), K = 768, one warp per row (W in shared memory)
__device__ __forceinline__ void gemv768(const Ctx& c, const uint8_t* W, int nr, int row_off) {
    Smem* sm = SM();
    float4 xa[3], xb[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float* xp = sm->xs + i * 256 + c.lane * 8;
        xa[i] = *reinterpret_cast<const float4*>(xp);
        xb[i] = *reinterpret_cast<const float4*>(xp + 4);
    }
#pragma unroll 1
    for (int r = c.warp; r < nr; r += N_CONS_WARPS) {
        const uint4* wp = reinterpret_cast<const uint4*>(W + (size_t)r * (C * 2)) + c.lane;
        uint4 w0 = wp[0], w1 = wp[32], w2 = wp[64];
        float s = dot8(w0, xa[0], xb[0]) + dot8(w1, xa[1], xb[1]) + dot8(w2, xa[2], xb[2]);
        s = warp_sum(s);
        if (c.lane == 0) sm->acc[row_off + r] = s;
    }
}
// K = 3072 split in 3 chunks of 1024: acc[(row_off + r) * 3 + chunk]; input vector in sm->stage
__device__ __forceinline__ void gemv3072(const Ctx& c, const uint8_t* W, int nr, int row_off) {
    Smem* sm = SM();
#pragma unroll 1
    for (int u = c.warp; u < nr * 3; u += N_CONS_WARPS) {
        int r = u / 3, ch = u - r * 3;
        const uint4* wp = reinterpret_cast<const uint4*>(W + (size_t)r * (FF * 2) + ch * 2048) + c.lane;
        const float* xp = sm->stage + ch * 1024 + c.lane * 8;
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint4 w = wp[i * 32];
            s += dot8(w, *reinterpret_cast<const float4*>(xp + i * 256), *reinterpret_cast<const float4*>(xp + i * 256 + 4));
        }
        s = warp_sum(s);
        if (c.lane == 0) sm->acc[(row_off + r) * 3 + ch] = s;
    }
}

// stream rows [r0, r1) of a [rows][768] matrix through the ring and leave the dot products in sm->acc
__device__ __forceinline__ void gemv_slice768(Ctx& c, int r0, int r1) {
    const int per = MAX_STAGE / (C * 2);
#pragma unroll 1
    for (int r = r0; r < r1; r += per) {
        int nr = min(per, r1 - r);
        Stage st;
        const uint8_t* w = acquire(c, (uint32_t)nr * C * 2, st);
        gemv768(c, w, nr, r - r0);
        cons_sync();
        release(c, st);
    }
}
__device__ __forceinline__ void gemv_slice3072(Ctx& c, int r0, int r1) {
    const int per = MAX_STAGE / (FF * 2);
#pragma unroll 1
    for (int r = r0; r < r1; r += per) {
        int nr = min(per, r1 - r);
        Stage st;
        const uint8_t* w = acquire(c, (uint32_t)nr * FF * 2, st);
        gemv3072(c, w, nr, r - r0);
        cons_sync();
        release(c, st);
    }
}

// ---- split-KV attention of one (head, split): reference module.py:214-227 with causal=True, 1 query --
// Reads q (and, on split 0, the new k/v) of layer `layer` from the LL lines tagged `want`; split 0 also
// appends the new row to the cache (module.py:209-210).  Publishes (m, l, o[48]) tagged `mine`.
__device__ __forceinline__ void attention_phase(Ctx& c, int layer, int j, uint32_t want, uint32_t mine) {
    const KParams& p = *c.p;
    Smem* sm = SM();
    const int ns = splits_for(j, p.nsplit);
    const int h = c.cta / p.nsplit, s = c.cta - h * p.nsplit;
    if (c.cta >= NH * p.nsplit || s >= ns) return;
    int k0, k1;
    kv_range(j, ns, s, k0, k1);
    const int nk = k1 - k0;
    const int has_new = (s == 0) ? 1 : 0;
    const int total = nk + has_new;
    float* scratch = c.scratch;
    if (total == 0) return;

    // q_h (and k_new_h, v_new_h) -> shared memory
    if (c.tid < 24) {
        float a, b;
        ll_load2(c, scratch + SC_Q, h * 24 + c.tid, want, a, b);
        sm->qs[2 * c.tid] = a; sm->qs[2 * c.tid + 1] = b;
    } else if (has_new && c.tid >= 32 && c.tid < 32 + 48) {
        const int t = c.tid - 32, which = t / 24, i = t - which * 24;
        float a, b;
        ll_load2(c, scratch + SC_KVN, which * (C / 2) + h * 24 + i, want, a, b);
        float* dst = which ? sm->vnew : sm->knew;
        dst[2 * i] = a; dst[2 * i + 1] = b;
        __half* row = (__half*)p.a.kv_h + (((size_t)(layer * 2 + which) * NH + h) * SMAX + j) * HD;
        *reinterpret_cast<__half2*>(row + 2 * i) = __floats2half2_rn(a, b);   // exact: rounded by the writer
    }
    cons_sync();
    PROBE(10)

    // scores: 2 lanes per key, 24 dims each; q pre-scaled by scale * log2(e)
    const int sub = c.lane & 1, kslot = c.lane >> 1;
    const float qscale = 0.14433756729740643f * 1.4426950408889634f;   // 1/sqrt(48) (module.py:196-198)
    float qr[24];
#pragma unroll
    for (int i = 0; i < 24; ++i) qr[i] = sm->qs[sub * 24 + i] * qscale;
    // P.V: lane = (key group kg of 4 keys, dim group dp of 6 dims); reduced over kg at the very end
    const int kg = c.lane >> 3, dp = c.lane & 7;
    float m_run = -INFINITY, l_run = 0.f;
    float o[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

    const int nblocks = (total + KV_BLOCK - 1) / KV_BLOCK;
#pragma unroll 1
    for (int b = 0; b < nblocks; ++b) {
        const int kb = b * KV_BLOCK;
        const int staged = max(0, min(KV_BLOCK, nk - kb));       // keys of this block that come through the ring
        const int in_block = min(KV_BLOCK, total - kb);
        Stage stk, stv;
        const uint8_t* ks = nullptr;
        const uint8_t* vs = nullptr;
        if (staged > 0) ks = acquire(c, (uint32_t)staged * HD * 2, stk);
        const int li = c.warp * 16 + kslot;
        float sc = -INFINITY;
        if (li < in_block) {
            float a;
            if (li < staged) {
                const uint4* kp = reinterpret_cast<const uint4*>(ks + (size_t)li * (HD * 2)) + sub * 3;
                uint4 w0 = kp[0], w1 = kp[1], w2 = kp[2];
                a = dot8(w0, make_float4(qr[0], qr[1], qr[2], qr[3]), make_float4(qr[4], qr[5], qr[6], qr[7]));
                a += dot8(w1, make_float4(qr[8], qr[9], qr[10], qr[11]), make_float4(qr[12], qr[13], qr[14], qr[15]));
                a += dot8(w2, make_float4(qr[16], qr[17], qr[18], qr[19]), make_float4(qr[20], qr[21], qr[22], qr[23]));
            } else {   // the key appended this step
                a = 0.f;
#pragma unroll
                for (int i = 0; i < 24; ++i) a = fmaf(sm->knew[sub * 24 + i], qr[i], a);
            }
            sc = a;
        }
        float other = __shfl_xor_sync(0xffffffffu, sc, 1);
        sc = (li < in_block) ? sc + other : -INFINITY;
        const float m_blk = warp_max(sc);
        PROBE(11)
        if (staged > 0) vs = acquire(c, (uint32_t)staged * HD * 2, stv);
        if (m_blk > -INFINITY) {     // warp-uniform
            const float m_new = fmaxf(m_run, m_blk);
            const float corr = exp2f(m_run - m_new);
            const float pr = (li < in_block) ? exp2f(sc - m_new) : 0.f;
            if (sub == 0) sm->pw[c.warp][kslot] = pr;
            float psum = warp_sum(pr) * 0.5f;                      // each key counted by its 2 lanes
            l_run = l_run * corr + psum;
#pragma unroll
            for (int e = 0; e < 6; ++e) o[e] *= corr;
            __syncwarp();
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int i = kg * 4 + t;                    // key within the warp's 16
                const int lk = c.warp * 16 + i;
                if (lk < in_block) {
                    const float pi = sm->pw[c.warp][i];
                    float f[6];
                    if (lk < staged) {
                        const __half2* vp = reinterpret_cast<const __half2*>(vs + (size_t)lk * (HD * 2) + dp * 12);
                        float2 a0 = __half22float2(vp[0]), a1 = __half22float2(vp[1]), a2 = __half22float2(vp[2]);
                        f[0] = a0.x; f[1] = a0.y; f[2] = a1.x; f[3] = a1.y; f[4] = a2.x; f[5] = a2.y;
                    } else {
#pragma unroll
                        for (int e = 0; e < 6; ++e) f[e] = sm->vnew[dp * 6 + e];
                    }
#pragma unroll
                    for (int e = 0; e < 6; ++e) o[e] = fmaf(pi, f[e], o[e]);
                }
            }
            m_run = m_new;
        }
        PROBE(12)
        cons_sync();
        if (staged > 0) { release(c, stk); release(c, stv); }
    }
    // reduce the 4 key groups, then merge the 16 warps
#pragma unroll
    for (int e = 0; e < 6; ++e) {
        o[e] += __shfl_xor_sync(0xffffffffu, o[e], 8);
        o[e] += __shfl_xor_sync(0xffffffffu, o[e], 16);
    }
    if (c.lane == 0) { sm->wpart[c.warp][0] = m_run; sm->wpart[c.warp][1] = l_run; }
    if (c.lane < 8) {
#pragma unroll
        for (int e = 0; e < 6; ++e) sm->wpart[c.warp][2 + dp * 6 + e] = o[e];
    }
    cons_sync();
    PROBE(13)
    // CTA partial (m, l, o[48]) = merge of the warps
    float a0 = 0.f, a1 = 0.f;
    if (c.tid < PART_VALS / 2) {          // thread t owns values 2t, 2t+1
        float m = -INFINITY;
#pragma unroll
        for (int w = 0; w < N_CONS_WARPS; ++w) m = fmaxf(m, sm->wpart[w][0]);
#pragma unroll 5
        for (int w = 0; w < N_CONS_WARPS; ++w) {
            const float mw = sm->wpart[w][0];
            const float f = (mw > -INFINITY) ? exp2f(mw - m) : 0.f;
            a0 = fmaf(f, sm->wpart[w][2 * c.tid], a0);
            a1 = fmaf(f, sm->wpart[w][2 * c.tid + 1], a1);
        }
        if (c.tid == 0) a0 = m;           // slot 0 carries the running max itself, slot 1 the sum
        if (s != 0) ll_store2(scratch + SC_PART, (h * MAX_SPLIT + s) * (PART_VALS / 2) + c.tid, a0, a1, mine);
    }
    PROBE(14)
    if (s != 0) return;
    // ---- head leader: merge the ns partials and publish y_h (48 values) to every replica, tagged mine + 1
    float2* st2 = reinterpret_cast<float2*>(sm->stage);
    if (c.tid < PART_VALS / 2) st2[c.tid] = make_float2(a0, a1);
    {
        const int nl = (ns - 1) * (PART_VALS / 2);
        if (c.tid < nl) {
            const int sp = 1 + c.tid / (PART_VALS / 2), ln = c.tid % (PART_VALS / 2);
            float v0, v1;
            ll_load2(c, scratch + SC_PART, (h * MAX_SPLIT + sp) * (PART_VALS / 2) + ln, mine, v0, v1);
            st2[sp * (PART_VALS / 2) + ln] = make_float2(v0, v1);
        }
    }
    cons_sync();
    if (c.tid < HD) {
        const float* base = sm->stage;
        float m = -INFINITY;
#pragma unroll 1
        for (int sp = 0; sp < ns; ++sp) m = fmaxf(m, base[sp * PART_VALS]);
        float l = 0.f, o = 0.f;
#pragma unroll 1
        for (int sp = 0; sp < ns; ++sp) {
            const float f = exp2f(base[sp * PART_VALS] - m);
            l = fmaf(f, base[sp * PART_VALS + 1], l);
            o = fmaf(f, base[sp * PART_VALS + 2 + c.tid], o);
        }
        ll_store1_rep(scratch + SC_Y, XV, h * HD + c.tid, o / l, mine + 1);
    }
    PROBE(15)
}





// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(N_THREADS, 1) decode_frame_kernel(const __grid_constant__ KParams p) {
    Smem* sm = SM();
    const UmgenDecodeArgs& a = p.a;
    Ctx c;
    c.p = &p; c.abort_flag = (int*)a.status_i32;
    c.cta = blockIdx.x; c.tid = threadIdx.x; c.warp = threadIdx.x >> 5; c.lane = threadIdx.x & 31;
    c.epoch = 0; c.probe = nullptr; c.probe_t0 = 0; c.tl = nullptr;
    float* scratch = (float*)a.scratch_f;
    c.scratch = scratch;
    const int G = p.grid, L = (int)a.n_layer;
    const int n_steps = (int)a.n_steps;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSLOT; ++i) { mbar_init(&sm->full[i], 1); mbar_init(&sm->empty[i], 1); }
        sm->nbox = 0; sm->tok = 0;
        mbar_fence_init();
    }
    __syncthreads();

    const __half* Wl = (const __half*)a.oar_h;
    const float* Fl = (const float*)a.oar_f;
    const __half* heads[3] = {(const __half*)a.head_map_h, (const __half*)a.head_bbox_h, (const __half*)a.head_img_h};
    const float* emb_tables[3] = {(const float*)a.map_table_f, (const float*)a.be_f, (const float*)a.img_table_f};

    int rq0, rq1, rp0, rp1, rf0, rf1;
    row_slice(3 * C, c.cta, G, rq0, rq1);
    row_slice(C, c.cta, G, rp0, rp1);
    row_slice(FF, c.cta, G, rf0, rf1);       // c_fc rows; the GMLP c_fc uses the same slice
    const int att_h = c.cta / p.nsplit, att_s = c.cta - att_h * p.nsplit;
    const bool att_cta = c.cta < NH * p.nsplit;

    // ============================== producer warp ==============================================
    if (c.warp == N_CONS_WARPS) {
        if (c.lane != 0) return;
        Producer pr;
        #pragma unroll 1
        for (int j = 0; j < n_steps; ++j) {
            const int q = j + 1;       // position produced by this step
            const int ns = splits_for(j, p.nsplit);
            int k0 = 0, k1 = 0;
            const bool att_active = att_cta && att_s < ns;
            if (att_active) kv_range(j, ns, att_s, k0, k1);
            bool kv_ok = false;
#pragma unroll 1
            for (int l = 0; l < L; ++l) {
                const uint8_t* wl = (const uint8_t*)(Wl + (size_t)l * LAYER_H);
                pr.issue_rows(c, wl + (size_t)OFF_QKV * 2, C * 2, rq0, rq1);
                if (att_active && k1 > k0) {
                    if (!kv_ok) { wait_kv_published(c, att_h, j); kv_ok = true; }   // rows < j of every layer
                    const __half* kb = (const __half*)a.kv_h + ((size_t)(l * 2 + 0) * NH + att_h) * SMAX * HD;
                    const __half* vb = (const __half*)a.kv_h + ((size_t)(l * 2 + 1) * NH + att_h) * SMAX * HD;
                    const int nk = k1 - k0, has_new = (att_s == 0);
                    const int nblocks = (nk + has_new + KV_BLOCK - 1) / KV_BLOCK;
#pragma unroll 1
                    for (int b = 0; b < nblocks; ++b) {
                        int staged = max(0, min(KV_BLOCK, nk - b * KV_BLOCK));
                        if (staged > 0) {
                            pr.issue(c, kb + (size_t)(k0 + b * KV_BLOCK) * HD, (uint32_t)staged * HD * 2);
                            pr.issue(c, vb + (size_t)(k0 + b * KV_BLOCK) * HD, (uint32_t)staged * HD * 2);
                        }
                    }
                }
                pr.issue_rows(c, wl + (size_t)OFF_PROJ * 2, C * 2, rp0, rp1);
                pr.issue_rows(c, wl + (size_t)OFF_FC * 2, C * 2, rf0, rf1);
                pr.issue_rows(c, wl + (size_t)OFF_PROJ2 * 2, FF * 2, rp0, rp1);
            }
            if (needs_head(q) && q > (int)a.prefix_len) {
                const int mod = pos_mod(q);
                int r0, r1;
                row_slice(vocab_of(mod), c.cta, G, r0, r1);
                pr.issue_rows(c, (const uint8_t*)heads[mod], C * 2, r0, r1);
            }
            if (*(volatile int*)c.abort_flag != 0) return;
        }
        return;
    }

    // ============================== consumer warps =============================================
    const float* tar = (const float*)a.tar_feat_f;
    int* out_tokens = (int*)a.out_tokens_i32;
    int* picks = (int*)a.picks_i32;
    const int* pose_tok = (const int*)a.pose_tok_i32;
    const int* teacher = (const int*)a.teacher_i32;
    const int rep = c.cta % KREP;              // the replica this CTA polls
    float* XB = scratch + SC_XB;
    float* XA = scratch + SC_XA;
    float* HB = scratch + SC_H;
    float* YB = scratch + SC_Y;

    // input of step 0: task embedding + TAR feature of index 0 (UMGen.py:1175,1215,1231)
    c.epoch = 1;
    for (int r = rp0 + c.tid; r < rp1; r += N_CONS) ll_store1_rep(XB, XV, r, __ldg((const float*)a.tske_f + r) + __ldg(tar + r), c.epoch);
    if (c.cta == 0 && c.tid < 8) {
        const int qs[8] = {1, 5, 6, 1031, 1032, 1693, 1694, 2207};
        out_tokens[qs[c.tid] - 1] = forced_id(qs[c.tid]);
        picks[qs[c.tid] - 1] = forced_id(qs[c.tid]);
        if (c.tid < 3) { out_tokens[1 + c.tid] = pose_tok[c.tid]; picks[1 + c.tid] = pose_tok[c.tid]; }
    }

#pragma unroll 1
    for (int j = 0; j < n_steps; ++j) {
        const int q = j + 1;
#pragma unroll 1
        for (int l = 0; l < L; ++l) {
            const __half* wl = Wl + (size_t)l * LAYER_H;
            const float* fl = Fl + (size_t)l * LAYER_F;
            c.probe = nullptr;
            c.tl = (a.debug_u64 && c.tid == 0 && l == 1 && j == 1200) ? (unsigned long long*)a.debug_u64 + c.cta * 16 : nullptr;
            TPROBE(0)
            if (c.tid == 0 && l == 1 && j == 1200 && (c.cta == 0 || c.cta == 77)) {
                c.probe = (int*)a.status_i32 + (c.cta == 0 ? 8 : 40);
                c.probe_t0 = clock64();
            }
            // ---- phase 1: LN1 -> c_attn (+bias): q, k_new, v_new published (module.py:206)
            uint32_t want = c.epoch, mine = ++c.epoch;
            const float bias_qkv = (rq0 + c.tid < rq1) ? __ldg(fl + F_BQKV + rq0 + c.tid) : 0.f;   // rows per CTA <= 512
            const float bias_proj = (rp0 + c.tid < rp1) ? __ldg(fl + F_BPROJ + rp0 + c.tid) : 0.f;
            read_ln(c, XB + rep * XV, want, fl + F_LN1);
            PROBE(0)
            TPROBE(1)
            gemv_slice768(c, rq0, rq1);
#pragma unroll 1
            for (int r = rq0 + c.tid; r < rq1; r += N_CONS) {
                float v = sm->acc[r - rq0] + bias_qkv;
                if (r < C) ll_store1(scratch + SC_Q, r, v, mine);
                else ll_store1(scratch + SC_KVN, r - C, __half2float(__float2half_rn(v)), mine);   // cache precision
            }
            PROBE(1)
            TPROBE(2)
            // ---- phase 2: split-KV attention over rows 0..j (+ cache append by split 0)
            want = mine; mine = ++c.epoch;
            attention_phase(c, l, j, want, mine);
            PROBE(2)
            TPROBE(3)
            // ---- phase 3: merge partials -> c_proj (+bias) -> residual (module.py:227-229, 409)
            want = ++c.epoch;              // the leaders' merge phase (tag mine + 1 inside attention_phase)
            mine = ++c.epoch;
            if (c.tid < C / 2) {
                float v0, v1;
                ll_load2(c, YB + rep * XV, c.tid, want, v0, v1);
                reinterpret_cast<float2*>(sm->xs)[c.tid] = make_float2(v0, v1);
            }
            cons_sync();
            PROBE(3)
            TPROBE(4)
            gemv_slice768(c, rp0, rp1);
#pragma unroll 1
            for (int r = rp0 + c.tid; r < rp1; r += N_CONS)
                ll_store1_rep(XA, XV, r, sm->xraw[r] + sm->acc[r - rp0] + bias_proj, mine);
            PROBE(4)
            TPROBE(5)
            // ---- phase 4: LN2 -> c_fc -> erf-GELU (module.py:245-247)
            want = mine; mine = ++c.epoch;
            read_ln(c, XA + rep * XV, want, fl + F_LN2);
            PROBE(5)
            TPROBE(6)
            gemv_slice768(c, rf0, rf1);
#pragma unroll 1
            for (int r = rf0 + c.tid; r < rf1; r += N_CONS) ll_store1_rep(HB, 2 * FF, r, gelu_erf(sm->acc[r - rf0]), mine);
            PROBE(6)
            TPROBE(7)
            // ---- phase 5: mlp c_proj -> residual (module.py:248, 410)
            want = mine; mine = ++c.epoch;
            ll_read_lines<(FF / 2 + N_CONS - 1) / N_CONS>(c, HB + rep * 2 * FF, FF / 2, want, sm->stage);
            cons_sync();
            PROBE(7)
            TPROBE(8)
            gemv_slice3072(c, rp0, rp1);
#pragma unroll 1
            for (int r = rp0 + c.tid; r < rp1; r += N_CONS) {
                const float* ap = sm->acc + (r - rp0) * 3;
                ll_store1_rep(XB, XV, r, sm->xraw[r] + (ap[0] + ap[1] + ap[2]), mine);
            }
            PROBE(8)
            TPROBE(9)
        }
        // publish this step's KV rows of head att_h (written by split 0 in every layer's phase 2)
        if (att_cta && att_s == 0) {
            cons_sync();
            if (c.tid == 0) {
                __threadfence();
                fence_proxy_async_global();
                asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"((uint32_t*)(scratch + SC_KVFLAG) + att_h), "r"((uint32_t)(j + 1)) : "memory");
            }
        }

        // ---- head + sampling (UMGen.py:1247-1250, 1046-1137)
        int tok;
        const int fid = forced_id(q);
        if (q <= 5) {
            tok = (fid >= 0) ? fid : __ldg(pose_tok + (q - 2));
        } else if (fid >= 0) {
            tok = fid;
        } else if (q <= (int)a.prefix_len) {
            tok = __ldg(teacher + (q - 1));      // given prefix (UMGen.py:1184-1201): no head, no sampling, no rule check
        } else {
            const int mod = pos_mod(q);
            const int V = vocab_of(mod);
            const int k = (int)(mod == 0 ? a.top_k_map : (mod == 1 ? a.top_k_bbox : a.top_k_img));
            int r0, r1;
            row_slice(V, c.cta, G, r0, r1);
            uint32_t want = c.epoch, mine = ++c.epoch;
            read_ln(c, XB + rep * XV, want, (const float*)a.ln_oar_f);
            gemv_slice768(c, r0, r1);
            if (a.logits_dump_f) {
                float* dump = (float*)a.logits_dump_f + (size_t)(q - 1) * 8192;
                for (int r = r0 + c.tid; r < r1; r += N_CONS) dump[r] = sm->acc[r - r0];
                cons_sync();           // warp 0 overwrites acc while selecting
            }
            if (a.sample_topp) {
                // ---- nucleus sampling: all-gather the logits, every CTA samples identically (UMGen.py:915-965)
                float* LG = scratch + SC_LOGIT;
#pragma unroll 1
                for (int r = r0 + c.tid; r < r1; r += N_CONS) ll_store1_rep(LG, LOGITV, r, sm->acc[r - r0], mine);
                float v[TOPP_PER];
                {
                    uint4 rr[(TOPP_PER + 1) / 2];
#pragma unroll
                    for (int t = 0; t < (TOPP_PER + 1) / 2; ++t) { const int line = c.tid + t * N_CONS; if (line < V / 2) rr[t] = ll_ld(LG + rep * LOGITV, line); }
#pragma unroll
                    for (int t = 0; t < (TOPP_PER + 1) / 2; ++t) {
                        const int line = c.tid + t * N_CONS;
                        float a0 = -INFINITY, a1 = -INFINITY;
                        if (line < V / 2) {
                            uint32_t spins = 0;
                            while (!(rr[t].y == mine && rr[t].w == mine)) { if (check_abort(c, spins)) break; rr[t] = ll_ld(LG + rep * LOGITV, line); }
                            a0 = __uint_as_float(rr[t].x); a1 = __uint_as_float(rr[t].z);
                        }
                        // ids 2*line and 2*line+1 are held as two consecutive "slots" of this thread
                        if (2 * t < TOPP_PER) v[2 * t] = a0;
                        if (2 * t + 1 < TOPP_PER) v[2 * t + 1] = a1;
                    }
                }
                const float pm = (float)(mod == 0 ? a.top_p_map : (mod == 1 ? a.top_p_bbox : a.top_p_img));
                const float inv_t = 1.0f / (float)a.temperature;
                const float u0 = philox_uniform(a.seed, (uint32_t)a.frame_index, (uint32_t)q, 0u);
                int slot = block_topp_sample(sm, v, pm, inv_t, u0, c.tid);
                // slot = tid' + i * N_CONS with i the thread-local position: id = 2 * (tid' + (i / 2) * N_CONS) + (i & 1)
                int t = 2 * ((slot % N_CONS) + ((slot / N_CONS) >> 1) * N_CONS) + ((slot / N_CONS) & 1);
                if (mod == 1) {
                    const int bidx = q - BBOX_FIRST_POS - 1;
                    const int prev = __ldg((const int*)a.prev_bbox_i32 + bidx);
                    const bool controlled = (a.control_mask >> ((q - BBOX_FIRST_POS) / 11)) & 1ull;
                    const float* row = (const float*)a.tar_bbox_logits_f + (size_t)bidx * 1028;
                    for (int pass = 0; pass < 2; ++pass) {
                        const bool go = pass == 0 ? controlled : (t == PAD_TOKEN && a.merge_ar_tar && prev != PAD_TOKEN);
                        if (!go) continue;
#pragma unroll
                        for (int i = 0; i < TOPP_PER; ++i) {
                            const int id = c.tid + i * N_CONS;
                            v[i] = (id < 1028 && !(controlled && id == 1027)) ? __ldg(row + id) : -INFINITY;
                        }
                        const float uu = philox_uniform(a.seed, (uint32_t)a.frame_index, (uint32_t)q, 1u + pass);
                        t = block_topp_sample(sm, v, (float)a.top_p_bbox, inv_t, uu, c.tid);
                        if (pass == 1 && c.cta == 0 && c.tid == 0) atomicAdd((int*)a.status_i32 + 2, 1);
                    }
                }
                bool wipe = false;
                if (c.warp == 0) {
                    if (mod == 1) { t = bbox_rules(sm, a, c.lane, c.cta, q, t, 0.f, true); wipe = (t & WIPE_BIT) != 0; t &= ~WIPE_BIT; }
                    if (c.lane == 0) {
                        if (wipe && c.cta == 0)
                            for (int i = 1; i <= 10; ++i) out_tokens[q - 1 - i] = PAD_TOKEN;
                        if (wipe) for (int i = 1; i <= 10; ++i) sm->recent[(q - i) & 15] = PAD_TOKEN;
                        sm->tok = t;
                    }
                }
                cons_sync();
                tok = sm->tok;
            } else {
            if (c.warp == 0) {     // local top-k of this CTA's slice -> candidate lines {val, tag, id, tag}
                float* cl = scratch + SC_CAND + (size_t)c.cta * MAX_CAND * 4;   // + k * CANDV per replica
                const int n = r1 - r0;
#pragma unroll 1
                for (int r = 0; r < k; ++r) {
                    float bv = -INFINITY;
                    int bi = 0x7fffffff;
#pragma unroll 1
                    for (int i = c.lane; i < n; i += 32) {
                        float v = sm->acc[i];
                        if (v > bv) { bv = v; bi = i; }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                    }
                    if (c.lane == 0) {
                        const int id = (bi == 0x7fffffff) ? 0x7fffffff : r0 + bi;
#pragma unroll
                        for (int kk = 0; kk < KREP; ++kk) ll_store2(cl + kk * CANDV, r, bv, __int_as_float(id), mine);
                        if (bi != 0x7fffffff) sm->acc[bi] = -INFINITY;
                    }
                    __syncwarp();
                }
            }
            // every CTA merges all candidates and decides the token identically
            want = mine;
            const int ncand = G * k;
            float* candv = sm->stage;
            int* candi = reinterpret_cast<int*>(sm->stage + MAX_GRID * MAX_CAND);
#pragma unroll 1
            {
                constexpr int N = (MAX_GRID * MAX_CAND + N_CONS - 1) / N_CONS;    // 5
                uint4 r[N];
                int src[N];
#pragma unroll
                for (int t = 0; t < N; ++t) {
                    const int i = c.tid + t * N_CONS;
                    src[t] = -1;
                    if (i < ncand) {
                        const int cta_i = i / k;
                        src[t] = cta_i * MAX_CAND + (i - cta_i * k);
                        r[t] = ll_ld(scratch + SC_CAND + rep * CANDV, src[t]);
                    }
                }
#pragma unroll
                for (int t = 0; t < N; ++t) {
                    if (src[t] >= 0) {
                        uint32_t spins = 0;
                        while (!(r[t].y == want && r[t].w == want)) {
                            if (check_abort(c, spins)) break;
                            r[t] = ll_ld(scratch + SC_CAND + rep * CANDV, src[t]);
                        }
                        candv[c.tid + t * N_CONS] = __uint_as_float(r[t].x);
                        candi[c.tid + t * N_CONS] = (int)r[t].z;
                    }
                }
            }
            cons_sync();
            if (c.warp == 0) {
                const float u0 = philox_uniform(a.seed, (uint32_t)a.frame_index, (uint32_t)q, 0u);
                int t = warp_topk_sample(candv, candi, ncand, k, 1.0f / (float)a.temperature, u0, c.lane);
                bool wipe = false;
                if (mod == 1) {
                    const float u2 = philox_uniform(a.seed, (uint32_t)a.frame_index, (uint32_t)q, 2u);
                    t = bbox_rules(sm, a, c.lane, c.cta, q, t, u2, false);
                            wipe = (t & WIPE_BIT) != 0;
                            t &= ~WIPE_BIT;
                }
                if (c.lane == 0) {
                    if (wipe && c.cta == 0)
                        for (int i = 1; i <= 10; ++i) out_tokens[q - 1 - i] = PAD_TOKEN;    // UMGen.py:1357-1365
                    if (wipe) for (int i = 1; i <= 10; ++i) sm->recent[(q - i) & 15] = PAD_TOKEN;
                    sm->tok = t;
                }
            }
            cons_sync();
            tok = sm->tok;
            }
        }
        int tok_used = tok;
        if (teacher != nullptr && q > 5 && fid < 0 && (a.prefix_len == 0 || q <= (int)a.prefix_len)) tok_used = __ldg(teacher + (q - 1));
        if (c.tid == 0) {
            sm->recent[q & 15] = tok_used;
            if (c.cta == 0 && q > 5) { out_tokens[q - 1] = tok_used; picks[q - 1] = tok; }
        }
        if (j == SEQ - 2) break;           // q = 2206 was the last sampled token; q = 2207 is forced

        // ---- embed the token as the next input and add the TAR feature of index j+1 (UMGen.py:1215-1231)
        const float* tnext = tar + (size_t)(j + 1) * C;
        {
            // token embedding as appended to out_tokens (UMGen.py:1046-1137): bos/eos -> axe, pose -> fouier_pe,
            // map/image -> GMLP(codebook[tok]) (precomputed table), bbox3d -> be
            const float* row;
            if (forced_id(q) >= 0) row = (const float*)a.axe_f + (size_t)forced_id(q) * C;
            else if (q <= 5) row = (const float*)a.fpe_f + (size_t)tok_used * C;
            else row = emb_tables[pos_mod(q)] + (size_t)tok_used * C;
            const uint32_t mine = ++c.epoch;
            cons_sync();        // everyone is done with acc / xraw of the previous phase
            for (int r = rp0 + c.tid; r < rp1; r += N_CONS) ll_store1_rep(XB, XV, r, __ldg(row + r) + __ldg(tnext + r), mine);
        }
        if (*(volatile int*)c.abort_flag != 0) break;
    }
    if (c.cta == 0 && c.tid == 0) ((int*)a.status_i32)[3] = n_steps;
}

__global__ void tar_bbox_logits_kernel(const float* __restrict__ tar_feat, const __half* __restrict__ w, float* __restrict__ out) {
    // one CTA per bbox content row; 8 warps, warp per output column
    __shared__ float xs[C];
    const int i = blockIdx.x;
    const float* xr = tar_feat + (size_t)(1032 + i) * C;
    for (int t = threadIdx.x; t < C; t += blockDim.x) xs[t] = xr[t];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    float4 xa[3], xb[3];
    for (int k = 0; k < 3; ++k) {
        xa[k] = *reinterpret_cast<const float4*>(xs + k * 256 + lane * 8);
        xb[k] = *reinterpret_cast<const float4*>(xs + k * 256 + lane * 8 + 4);
    }
    for (int v = warp; v < 1028; v += nw) {
        const uint4* wp = reinterpret_cast<const uint4*>(w + (size_t)v * C) + lane;
        float s = dot8(__ldg(wp), xa[0], xb[0]) + dot8(__ldg(wp + 32), xa[1], xb[1]) + dot8(__ldg(wp + 64), xa[2], xb[2]);
        s = warp_sum(s);
        if (lane == 0) out[(size_t)i * 1028 + v] = s;
    }
}

}  // namespace umgen

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
namespace umgen {
extern int64_t g_launches;
int decode_cluster_capacity();                                           // decode_cluster.cu
int64_t decode_cluster_scratch_floats();
int decode_cluster_launch(const UmgenDecodeArgs* args, int n_scenes, cudaStream_t stream);
int decode_cluster_need();
int decode_cluster_max_scenes();
}
using namespace umgen;

extern "C" int64_t umgen_decode_scratch_floats(void) {
    const int64_t a = SC_TOTAL, b = decode_cluster_scratch_floats();
    return a > b ? a : b;
}

extern "C" int umgen_decode_max_scenes(void) { return decode_cluster_max_scenes(); }

static int check_decode_args(const UmgenDecodeArgs* args) {
    if (args->n_layer < 1 || args->n_layer > 256) { set_error("n_layer out of range: %lld", (long long)args->n_layer); return -1; }
    if (args->mode < 0 || args->mode > 2) { set_error("mode must be 0 (auto), 1 (L2-exchange kernel) or 2 (8-cluster kernel)"); return -1; }
    if (args->prefix_len < 0 || args->prefix_len > SEQ || (args->prefix_len > 5 && !args->teacher_i32)) { set_error("prefix_len needs teacher_i32 and must be in [0, 2207]"); return -1; }
    if (args->n_steps < 1 || args->n_steps > SEQ - 1) { set_error("n_steps must be in [1, 2206]"); return -1; }
    const int64_t ks[3] = {args->top_k_map, args->top_k_bbox, args->top_k_img};
    for (int i = 0; i < 3; ++i)
        if (ks[i] < 1 || ks[i] > MAX_CAND) { set_error("top_k must be in [1, 16] (got %lld)", (long long)ks[i]); return -1; }
    if (!(args->temperature > 0)) { set_error("temperature must be > 0"); return -1; }
    if (args->sample_topp && !(args->top_p_map > 0 && args->top_p_bbox > 0 && args->top_p_img > 0)) { set_error("top_p values must be > 0"); return -1; }
    if (!args->tar_bbox_logits_f && (args->merge_ar_tar || args->control_mask)) { set_error("tar_bbox_logits required"); return -1; }
    if (!args->kv_h || !args->scratch_f || !args->out_tokens_i32 || !args->picks_i32 || !args->status_i32 || !args->tar_feat_f) {
        set_error("null buffer"); return -1;
    }
    return 0;
}

// Several scenes decoded in lockstep by ONE launch of the 8-cluster kernel (the reference decodes one scene at a time,
// UMGen.py:907,1093): args[0 .. n_scenes) differ in the per-frame inputs, the state and the outputs only.
extern "C" int umgen_decode_frames(const UmgenDecodeArgs* args, int64_t n_scenes, void* stream_v) {
    if (!args) { set_error("null args"); return -1; }
    if (n_scenes == 1) return umgen_decode_frame(args, stream_v);
    if (n_scenes < 1 || n_scenes > decode_cluster_max_scenes()) { set_error("n_scenes must be in [1, %d] (got %lld)", decode_cluster_max_scenes(), (long long)n_scenes); return -1; }
    const UmgenDecodeArgs& a0 = args[0];
    for (int64_t s = 0; s < n_scenes; ++s) {
        const UmgenDecodeArgs& a = args[s];
        if (int rc = check_decode_args(&a)) return rc;
        if (a.mode == 1) { set_error("the L2-exchange kernel decodes one scene per launch"); return -1; }
        const bool same = a.n_layer == a0.n_layer && a.oar_f == a0.oar_f && a.ln_oar_f == a0.ln_oar_f && a.head_map_h == a0.head_map_h &&
                          a.head_bbox_h == a0.head_bbox_h && a.head_img_h == a0.head_img_h && a.map_table_f == a0.map_table_f && a.img_table_f == a0.img_table_f &&
                          a.be_f == a0.be_f && a.axe_f == a0.axe_f && a.tske_f == a0.tske_f && a.fpe_f == a0.fpe_f && a.box_lut_d == a0.box_lut_d &&
                          a.oar_cl_h == a0.oar_cl_h && a.prefix_len == a0.prefix_len && a.top_k_map == a0.top_k_map && a.top_k_bbox == a0.top_k_bbox &&
                          a.top_k_img == a0.top_k_img && a.sample_topp == a0.sample_topp && a.top_p_map == a0.top_p_map && a.top_p_bbox == a0.top_p_bbox &&
                          a.top_p_img == a0.top_p_img && a.temperature == a0.temperature && a.merge_ar_tar == a0.merge_ar_tar &&
                          a.rule_constrain == a0.rule_constrain && a.n_steps == a0.n_steps && a.scratch_f == a0.scratch_f && a.grid == a0.grid;
        if (!same) { set_error("scene %lld differs from scene 0 in the weights, the sampling set-up, prefix_len, n_steps or the scratch buffer", (long long)s); return -1; }
        for (int64_t o = 0; o < s; ++o)
            if (args[o].kv_h == a.kv_h || args[o].out_tokens_i32 == a.out_tokens_i32 || args[o].picks_i32 == a.picks_i32 || args[o].status_i32 == a.status_i32) {
                set_error("scenes %lld and %lld share a state / output buffer", (long long)o, (long long)s); return -1;
            }
    }
    if (!a0.oar_cl_h || decode_cluster_capacity() < decode_cluster_need()) { set_error("several scenes per launch need the 8-cluster kernel (oar_cl_h, 8 resident clusters)"); return -3; }
    return decode_cluster_launch(args, (int)n_scenes, (cudaStream_t)stream_v);
}

extern "C" int umgen_decode_frame(const UmgenDecodeArgs* args, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    if (!args) { set_error("null args"); return -1; }
    if (int rc = check_decode_args(args)) return rc;
    const bool cluster_pick = args->mode == 2 || (args->mode == 0 && args->oar_cl_h && decode_cluster_capacity() >= decode_cluster_need());
    if (args->tar_ready_i32 && !cluster_pick) { set_error("tar_ready_i32 is supported by the 8-cluster decode kernel only"); return -1; }
    if (args->mode == 2 || (args->mode == 0 && args->oar_cl_h && decode_cluster_capacity() >= decode_cluster_need())) return decode_cluster_launch(args, 1, stream);
    if (!args->oar_h) { set_error("the L2-exchange decode kernel needs oar_h"); return -1; }
    int dev = 0, sms = 0, coop = 0;
    UMGEN_CUDA_OK(cudaGetDevice(&dev));
    UMGEN_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    UMGEN_CUDA_OK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    if (!coop) { set_error("device lacks cooperative launch"); return -3; }
    const size_t smem = sizeof(Smem) + 128;
    UMGEN_CUDA_OK(cudaFuncSetAttribute(decode_frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    UMGEN_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, decode_frame_kernel, N_THREADS, smem));
    if (occ < 1) { set_error("decode kernel does not fit on an SM (smem %zu)", smem); return -3; }
    KParams kp;
    kp.a = *args;
    kp.grid = args->grid > 0 ? (int)args->grid : sms;
    if (kp.grid > sms * occ || kp.grid > MAX_GRID || kp.grid < 64) {
        set_error("grid %d not co-resident / supported (sms %d, occ %d, range 64..%d)", kp.grid, sms, occ, MAX_GRID);
        return -3;
    }
    kp.nsplit = kp.grid / NH;
    if (kp.nsplit > MAX_SPLIT) kp.nsplit = MAX_SPLIT;
    UMGEN_CUDA_OK(cudaMemsetAsync(args->scratch_f, 0, SC_TOTAL * sizeof(float), stream));
    UMGEN_CUDA_OK(cudaMemsetAsync(args->status_i32, 0, 96 * sizeof(int), stream));
    void* kargs[] = {&kp};
    UMGEN_CUDA_OK(cudaLaunchCooperativeKernel((void*)decode_frame_kernel, dim3(kp.grid), dim3(N_THREADS), kargs, smem, stream));
    g_launches += 1;
    return 0;
}

__global__ void signal_ready_kernel(int* flag, int value) {
    __threadfence();
    asm volatile("st.release.gpu.global.b32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}
extern "C" int umgen_signal_ready(void* flag_i32, int64_t value, void* stream_v) {
    if (!flag_i32) { set_error("null flag"); return -1; }
    signal_ready_kernel<<<1, 1, 0, (cudaStream_t)stream_v>>>((int*)flag_i32, (int)value);
    UMGEN_CUDA_OK(cudaGetLastError());
    g_launches += 1;
    return 0;
}

// Test entry for the bbox3d rule path's geometry: the same __device__ functions bbox_rules runs (box_corners, last_box_collides), one warp per case.
__global__ void check_collision_kernel(const double* __restrict__ boxes, const int* __restrict__ offsets, int n_cases, int* __restrict__ out) {
    __shared__ float corners[MAX_BOX][8];
    __shared__ int dropped[MAX_BOX];
    const int cs = blockIdx.x, lane = threadIdx.x;
    if (cs >= n_cases) return;
    const int b0 = offsets[cs], nb = offsets[cs + 1] - b0;
    if (nb > MAX_BOX) { if (lane == 0) out[cs] = -1; return; }
    for (int i = lane; i < nb; i += 32) {
        const double* b = boxes + (size_t)(b0 + i) * 10;
        box_corners(b[0], b[1], b[3], b[4], b[6], corners[i]);
        dropped[i] = (b[0] >= 63.0) ? 1 : 0;
    }
    __syncwarp();
    const bool hit = last_box_collides(corners, dropped, nb, lane);
    if (lane == 0) out[cs] = hit ? 1 : 0;
}
extern "C" int umgen_check_collision(const void* boxes_d, const void* offsets_i32, int64_t n_cases, void* out_i32, void* stream_v) {
    if (!boxes_d || !offsets_i32 || !out_i32 || n_cases < 1) { set_error("check_collision: bad args"); return -1; }
    check_collision_kernel<<<(unsigned)n_cases, 32, 0, (cudaStream_t)stream_v>>>((const double*)boxes_d, (const int*)offsets_i32, (int)n_cases, (int*)out_i32);
    UMGEN_CUDA_OK(cudaGetLastError());
    g_launches += 1;
    return 0;
}

extern "C" int umgen_tar_bbox_logits(const void* tar_feat_f, const void* head_tar_bbox_h, void* out_f, void* stream_v) {
    if (!tar_feat_f || !head_tar_bbox_h || !out_f) { set_error("null buffer"); return -1; }
    tar_bbox_logits_kernel<<<660, 256, 0, (cudaStream_t)stream_v>>>((const float*)tar_feat_f, (const __half*)head_tar_bbox_h, (float*)out_f);
    UMGEN_CUDA_OK(cudaGetLastError());
    g_launches += 1;
    return 0;
}

// Lazy module loading (the CUDA 12 default) loads a kernel on its first launch and that load waits for an idle device -- which never comes while the
// persistent decode kernel spins on a flag.  umgen_preload() (capi.cu) forces every kernel of the library to load up front.
#define UMGEN_PRELOAD(k) UMGEN_CUDA_OK(cudaFuncGetAttributes(&fa_, k))
namespace umgen {
int preload_decode() {
    cudaFuncAttributes fa_;
    UMGEN_PRELOAD(decode_frame_kernel); UMGEN_PRELOAD(tar_bbox_logits_kernel); UMGEN_PRELOAD(signal_ready_kernel); UMGEN_PRELOAD(check_collision_kernel);
    return 0;
}
}  // namespace umgen
