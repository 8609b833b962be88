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
#include "common.cuh"
#include "../../include/umgen.h"

namespace umgen {

constexpr int N_CONS_WARPS = 16;
constexpr int N_CONS = N_CONS_WARPS * 32;        // 512 consumer threads
constexpr int N_THREADS = N_CONS + 32;           // + producer warp
constexpr uint32_t RING_BYTES = 176 * 1024;
constexpr uint32_t MAX_STAGE = 36864;            // 6 rows of 3072 halves / 24 rows of 768 halves / 384 KV rows
constexpr int NSLOT = 8;
constexpr int KV_BLOCK = 256;                    // keys per attention block (one K stage + one V stage)
constexpr int MAX_ROWS = 160;                    // max rows of any GEMV slice per CTA (grid >= 64)
constexpr int MAX_CAND = 16;
constexpr int MAX_GRID = 160;
constexpr int MAX_BOX = 64;
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

// scratch layout (floats)
constexpr int SC_X = 0;                           // [768] residual stream
constexpr int SC_Q = SC_X + C;                    // [768]
constexpr int SC_H = SC_Q + C;                    // [3072]
constexpr int SC_PART = SC_H + FF;                // [16][MAX_SPLIT=10][52]
constexpr int MAX_SPLIT = 10;
constexpr int SC_CANDV = SC_PART + NH * MAX_SPLIT * PART_STRIDE;   // [MAX_GRID][16]
constexpr int SC_CANDI = SC_CANDV + MAX_GRID * MAX_CAND;            // [MAX_GRID][16] (int)
constexpr int SC_BAR = SC_CANDI + MAX_GRID * MAX_CAND;              // barrier counter (+ padding)
constexpr int SC_TOTAL = SC_BAR + 64;

struct KParams {
    UmgenDecodeArgs a;
    int grid;
    int nsplit;
};

struct __align__(128) Smem {
    uint8_t ring[RING_BYTES];
    float xs[C];                  // normalised input of the current GEMV
    float hs[FF];                 // MLP hidden / attention y
    float acc[MAX_ROWS * 3];      // raw dot products (row, k-chunk)
    float wpart[N_CONS_WARPS][PART_STRIDE];
    float candv[MAX_GRID * MAX_CAND];
    int candi[MAX_GRID * MAX_CAND];
    float red[32];
    float corners[MAX_BOX][8];    // decoded boxes of this frame (UMGen.py:1183,1338)
    int box_dropped[MAX_BOX];     // x >= 63 (misc.py:475-481)
    int recent[16];               // last tokens by position & 15
    float code[16];
    uint64_t full[NSLOT];
    uint64_t empty[NSLOT];
    uint32_t fl_off[NSLOT];       // producer bookkeeping of in-flight stages
    uint32_t fl_bytes[NSLOT];
    volatile uint32_t progress;   // device-wide barriers passed by this CTA's consumers
    volatile int tok;             // token decided for the current position
    int nbox;
    int dead;
};

// ------------------------------------------------------------------------------------------------
// static schedule helpers (identical on producer and consumer side)
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int forced_id(int q) {   // q: 1-indexed position; -1 if sampled
    switch (q) {
        case 1: return 0; case 5: return 1; case 6: return 2; case 1031: return 3;
        case 1032: return 4; case 1693: return 5; case 1694: return 6; case 2207: return 7;
        default: return -1;
    }
}
// 0 map, 1 bbox3d, 2 image, 3 pose (UMGen.py:986-992)
__host__ __device__ __forceinline__ int pos_mod(int q) { return q <= 5 ? 3 : (q <= 1031 ? 0 : (q <= 1693 ? 1 : 2)); }
__host__ __device__ __forceinline__ bool needs_head(int q) { return forced_id(q) < 0 && q > 5; }
__host__ __device__ __forceinline__ bool needs_gmlp(int q) { return needs_head(q) && pos_mod(q) != 1; }
__host__ __device__ __forceinline__ int vocab_of(int mod) { return mod == 1 ? 1028 : 8192; }

__device__ __forceinline__ void row_slice(int rows, int cta, int grid, int& r0, int& r1) {
    r0 = (int)(((long long)rows * cta) / grid);
    r1 = (int)(((long long)rows * (cta + 1)) / grid);
}
__device__ __forceinline__ void kv_range(int nold, int nsplit, int s, int& k0, int& k1) {
    int chunk = (nold + nsplit - 1) / nsplit;
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
    Smem* sm;
    int* abort_flag;     // status[0]
    uint64_t deadline;
    int cta, tid, warp, lane;
    Ring ring;
    uint32_t epoch;      // device-wide barriers passed
    uint32_t* bar;
};

__device__ __forceinline__ bool check_abort(Ctx& c, uint32_t& spins) {
    if ((++spins & 0x3ffu) == 0) {
        if (*(volatile int*)c.abort_flag != 0) return true;
        if (globaltimer_ns() > c.deadline) { atomicCAS(c.abort_flag, 0, 100 + (int)(c.epoch & 0xffff)); return true; }
    }
    return false;
}
__device__ __forceinline__ void wait_mbar(Ctx& c, uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (check_abort(c, spins)) return;
    }
}
__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, %0;" ::"n"(N_CONS) : "memory"); }

// device-wide barrier among the consumer threads of all CTAs
__device__ __forceinline__ void grid_barrier(Ctx& c) {
    cons_sync();
    c.epoch++;
    if (c.tid == 0) {
        __threadfence();
        fence_proxy_async_global();      // later bulk copies (async proxy) of other CTAs read our KV writes
        red_release_gpu_add(c.bar, 1u);
        const uint32_t target = c.epoch * (uint32_t)c.p->grid;
        uint32_t spins = 0;
        while (ld_acquire_gpu(c.bar) < target) {
            if (check_abort(c, spins)) break;
        }
        __threadfence();
        c.sm->progress = c.epoch;
    }
    cons_sync();
}

// ------------------------------------------------------------------------------------------------
// stage acquire / release (consumer) and issue (producer)
// ------------------------------------------------------------------------------------------------
// consumer: returns a pointer to `bytes` of data that mirror [src, src+bytes)
__device__ __forceinline__ const uint8_t* acquire(Ctx& c, const void* src, uint32_t bytes, Stage& st) {
    if (c.p->a.mode == 1) return (const uint8_t*)src;
    st = ring_next(c.ring, bytes);
    wait_mbar(c, &c.sm->full[st.slot], st.parity);
    return c.sm->ring + st.off;
}
// consumer: all consumer threads are past their last read of the stage (caller synced)
__device__ __forceinline__ void release(Ctx& c, const Stage& st) {
    if (c.p->a.mode == 1) return;
    if (c.tid == 0) mbar_arrive(&c.sm->empty[st.slot]);
}

struct Producer {
    Ctx* c;
    uint32_t tail = 0;   // oldest stage not known to be released
    __device__ void issue(const void* src, uint32_t bytes) {
        Ctx& cx = *c;
        Smem* sm = cx.sm;
        Stage st = ring_next(cx.ring, bytes);
        const uint32_t me = cx.ring.k - 1;
        while (true) {
            bool conflict = (me - tail) >= (uint32_t)NSLOT;
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
    // rows [r0, r1) of a row-major matrix, split into sub-stages of at most MAX_STAGE bytes
    __device__ void issue_rows(const uint8_t* base, uint32_t row_bytes, int r0, int r1) {
        const int per = (int)(MAX_STAGE / row_bytes);
        for (int r = r0; r < r1; r += per) {
            int nr = min(per, r1 - r);
            issue(base + (size_t)r * row_bytes, (uint32_t)nr * row_bytes);
        }
    }
    __device__ void wait_progress(uint32_t need) {
        Ctx& cx = *c;
        uint32_t spins = 0;
        while (cx.sm->progress < need) {
            __nanosleep(64);
            if (check_abort(cx, spins)) return;
        }
        __threadfence();
        fence_proxy_async_global();
    }
};

// ------------------------------------------------------------------------------------------------
// consumer math
// ------------------------------------------------------------------------------------------------
// LayerNorm (module.py:26-37: weight only, eps 1e-5) of the global vector x -> sm->xs
__device__ void layer_norm_to_smem(Ctx& c, const float* x, const float* w) {
    Smem* sm = c.sm;
    float4 v = make_float4(0, 0, 0, 0);
    if (c.tid < C / 4) v = __ldcg(reinterpret_cast<const float4*>(x) + c.tid);
    float s = warp_sum(v.x + v.y + v.z + v.w);
    if (c.lane == 0 && c.warp < 6) sm->red[c.warp] = s;
    cons_sync();
    float mean = (sm->red[0] + sm->red[1] + sm->red[2] + sm->red[3] + sm->red[4] + sm->red[5]) * (1.0f / C);
    float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
    float q = (c.tid < C / 4) ? (dx * dx + dy * dy + dz * dz + dw * dw) : 0.f;
    q = warp_sum(q);
    if (c.lane == 0 && c.warp < 6) sm->red[8 + c.warp] = q;
    cons_sync();
    float var = (sm->red[8] + sm->red[9] + sm->red[10] + sm->red[11] + sm->red[12] + sm->red[13]) * (1.0f / C);
    float rstd = rsqrtf(var + 1e-5f);
    if (c.tid < C / 4) {
        float4 g = __ldg(reinterpret_cast<const float4*>(w) + c.tid);
        reinterpret_cast<float4*>(sm->xs)[c.tid] = make_float4(dx * rstd * g.x, dy * rstd * g.y, dz * rstd * g.z, dw * rstd * g.w);
    }
    cons_sync();
}

// acc[(row_off + r) ] = W[r][:] . xs  for r in [0, nr), K = 768, one warp per row
__device__ __forceinline__ void gemv768(const Ctx& c, const uint8_t* W, int nr, const float* xs, float* acc, int row_off) {
    float4 xa[3], xb[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float* xp = xs + i * 256 + c.lane * 8;
        xa[i] = *reinterpret_cast<const float4*>(xp);
        xb[i] = *reinterpret_cast<const float4*>(xp + 4);
    }
    for (int r = c.warp; r < nr; r += N_CONS_WARPS) {
        const uint4* wp = reinterpret_cast<const uint4*>(W + (size_t)r * (C * 2)) + c.lane;
        uint4 w0 = wp[0], w1 = wp[32], w2 = wp[64];
        float s = dot8(w0, xa[0], xb[0]) + dot8(w1, xa[1], xb[1]) + dot8(w2, xa[2], xb[2]);
        s = warp_sum(s);
        if (c.lane == 0) acc[row_off + r] = s;
    }
}
// K = 3072 split in 3 chunks of 1024: acc[(row_off + r) * 3 + chunk]
__device__ __forceinline__ void gemv3072(const Ctx& c, const uint8_t* W, int nr, const float* hs, float* acc, int row_off) {
    for (int u = c.warp; u < nr * 3; u += N_CONS_WARPS) {
        int r = u / 3, ch = u - r * 3;
        const uint4* wp = reinterpret_cast<const uint4*>(W + (size_t)r * (FF * 2) + ch * 2048) + c.lane;
        const float* xp = hs + ch * 1024 + c.lane * 8;
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint4 w = wp[i * 32];
            s += dot8(w, *reinterpret_cast<const float4*>(xp + i * 256), *reinterpret_cast<const float4*>(xp + i * 256 + 4));
        }
        s = warp_sum(s);
        if (c.lane == 0) acc[(row_off + r) * 3 + ch] = s;
    }
}

// stream rows [r0, r1) of a [rows][768] matrix through the ring and leave the dot products in sm->acc
__device__ void gemv_slice768(Ctx& c, const __half* W, int r0, int r1) {
    const int per = MAX_STAGE / (C * 2);
    for (int r = r0; r < r1; r += per) {
        int nr = min(per, r1 - r);
        Stage st;
        const uint8_t* w = acquire(c, W + (size_t)r * C, (uint32_t)nr * C * 2, st);
        gemv768(c, w, nr, c.sm->xs, c.sm->acc, r - r0);
        cons_sync();
        release(c, st);
    }
}
__device__ void gemv_slice3072(Ctx& c, const __half* W, int r0, int r1) {
    const int per = MAX_STAGE / (FF * 2);
    for (int r = r0; r < r1; r += per) {
        int nr = min(per, r1 - r);
        Stage st;
        const uint8_t* w = acquire(c, W + (size_t)r * FF, (uint32_t)nr * FF * 2, st);
        gemv3072(c, w, nr, c.sm->hs, c.sm->acc, r - r0);
        cons_sync();
        release(c, st);
    }
}

// ---- split-KV attention of one (head, split): reference module.py:214-227 with causal=True, 1 query --
__device__ void attention_phase(Ctx& c, int layer, int j) {
    const KParams& p = *c.p;
    Smem* sm = c.sm;
    const int nsplit = p.nsplit;
    if (c.cta >= NH * nsplit) return;
    const int h = c.cta / nsplit, s = c.cta - h * nsplit;
    int k0, k1;
    kv_range(j, nsplit, s, k0, k1);
    const int nk = k1 - k0;
    const int has_new = (s == 0) ? 1 : 0;
    const int total = nk + has_new;
    float* scratch = (float*)p.a.scratch_f;
    float* part = scratch + SC_PART + (h * MAX_SPLIT + s) * PART_STRIDE;
    if (total == 0) return;     // readers derive validity from (j, s) themselves

    const __half* kbase = (const __half*)p.a.kv_h + ((size_t)(layer * 2 + 0) * NH + h) * SMAX * HD;
    const __half* vbase = (const __half*)p.a.kv_h + ((size_t)(layer * 2 + 1) * NH + h) * SMAX * HD;

    // query slice of this lane: 2 lanes per key, 24 dims each; pre-scaled by scale * log2(e)
    const int sub = c.lane & 1, kslot = c.lane >> 1;
    const float qscale = 0.14433756729740643f * 1.4426950408889634f;   // 1/sqrt(48) (module.py:196-198)
    float qr[24];
    {
        const float* qp = scratch + SC_Q + h * HD + sub * 24;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            float4 t = __ldcg(reinterpret_cast<const float4*>(qp) + i);
            qr[i * 4 + 0] = t.x * qscale; qr[i * 4 + 1] = t.y * qscale; qr[i * 4 + 2] = t.z * qscale; qr[i * 4 + 3] = t.w * qscale;
        }
    }
    float m_run = -INFINITY, l_run = 0.f, o0 = 0.f, o1 = 0.f;   // o0/o1: dims 2*lane, 2*lane+1 (lanes < 24)

    const int nblocks = (total + KV_BLOCK - 1) / KV_BLOCK;
    for (int b = 0; b < nblocks; ++b) {
        const int kb = b * KV_BLOCK;
        const int staged = max(0, min(KV_BLOCK, nk - kb));       // keys of this block that come through the ring
        const int in_block = min(KV_BLOCK, total - kb);
        Stage stk, stv;
        const uint8_t* ks = nullptr;
        const uint8_t* vs = nullptr;
        if (staged > 0) ks = acquire(c, kbase + (size_t)(k0 + kb) * HD, (uint32_t)staged * HD * 2, stk);
        // scores: this warp owns keys kb + it*256.. -> local index li = warp*16 + kslot
        const int li = c.warp * 16 + kslot;
        float sc = -INFINITY;
        if (li < in_block) {
            uint4 w0, w1, w2;
            if (li < staged) {
                const uint4* kp = reinterpret_cast<const uint4*>(ks + (size_t)li * (HD * 2)) + sub * 3;
                w0 = kp[0]; w1 = kp[1]; w2 = kp[2];
            } else {   // the key appended this step (row j), written to HBM in phase 1 by other CTAs
                const uint4* kp = reinterpret_cast<const uint4*>(kbase + (size_t)j * HD) + sub * 3;
                w0 = __ldcg(kp); w1 = __ldcg(kp + 1); w2 = __ldcg(kp + 2);
            }
            float a = dot8(w0, make_float4(qr[0], qr[1], qr[2], qr[3]), make_float4(qr[4], qr[5], qr[6], qr[7]));
            a += dot8(w1, make_float4(qr[8], qr[9], qr[10], qr[11]), make_float4(qr[12], qr[13], qr[14], qr[15]));
            a += dot8(w2, make_float4(qr[16], qr[17], qr[18], qr[19]), make_float4(qr[20], qr[21], qr[22], qr[23]));
            sc = a;
        }
        float other = __shfl_xor_sync(0xffffffffu, sc, 1);
        sc = (li < in_block) ? sc + other : -INFINITY;
        const float m_blk = warp_max(sc);
        if (staged > 0) vs = acquire(c, vbase + (size_t)(k0 + kb) * HD, (uint32_t)staged * HD * 2, stv);
        if (m_blk > -INFINITY) {     // warp-uniform
            const float m_new = fmaxf(m_run, m_blk);
            const float corr = exp2f(m_run - m_new);
            const float pr = (li < in_block) ? exp2f(sc - m_new) : 0.f;
            float psum = warp_sum(pr) * 0.5f;                      // each key counted by its 2 lanes
            l_run = l_run * corr + psum;
            o0 *= corr; o1 *= corr;
            const int nloc = min(16, in_block - c.warp * 16);
            for (int i = 0; i < nloc; ++i) {
                const float pi = __shfl_sync(0xffffffffu, pr, 2 * i);
                const int lk = c.warp * 16 + i;
                if (c.lane < 24) {
                    __half2 vv;
                    if (lk < staged) vv = *reinterpret_cast<const __half2*>(vs + (size_t)lk * (HD * 2) + c.lane * 4);
                    else {
                        unsigned int raw = __ldcg(reinterpret_cast<const unsigned int*>(vbase + (size_t)j * HD) + c.lane);
                        vv = *reinterpret_cast<__half2*>(&raw);
                    }
                    float2 vf = __half22float2(vv);
                    o0 = fmaf(pi, vf.x, o0); o1 = fmaf(pi, vf.y, o1);
                }
            }
            m_run = m_new;
        }
        cons_sync();
        if (staged > 0) { release(c, stk); release(c, stv); }
    }
    // merge the 16 warps
    if (c.lane == 0) { sm->wpart[c.warp][0] = m_run; sm->wpart[c.warp][1] = l_run; }
    if (c.lane < 24) { sm->wpart[c.warp][2 + 2 * c.lane] = o0; sm->wpart[c.warp][3 + 2 * c.lane] = o1; }
    cons_sync();
    if (c.tid < HD + 2) {
        float m = -INFINITY;
        for (int w = 0; w < N_CONS_WARPS; ++w) m = fmaxf(m, sm->wpart[w][0]);
        float accv = 0.f;
        for (int w = 0; w < N_CONS_WARPS; ++w) {
            float mw = sm->wpart[w][0];
            float f = (mw > -INFINITY) ? exp2f(mw - m) : 0.f;
            accv += f * (c.tid == 0 ? 0.f : sm->wpart[w][c.tid]);
        }
        part[c.tid] = (c.tid == 0) ? m : accv;       // [0]=m, [1]=l, [2..49]=o (unnormalised)
    }
}

// merge the split partials of all heads into sm->hs[0..767] (attention output y)
__device__ void combine_partials(Ctx& c, int j) {
    const KParams& p = *c.p;
    const float* scratch = (const float*)p.a.scratch_f;
    for (int t = c.tid; t < C; t += N_CONS) {
        const int h = t / HD, d = t - h * HD;
        float ms[MAX_SPLIT];
        float m = -INFINITY;
        for (int s = 0; s < p.nsplit; ++s) {
            int k0, k1;
            kv_range(j, p.nsplit, s, k0, k1);
            bool valid = (s == 0) || (k1 > k0);
            ms[s] = valid ? __ldcg(scratch + SC_PART + (h * MAX_SPLIT + s) * PART_STRIDE) : -INFINITY;
            m = fmaxf(m, ms[s]);
        }
        float l = 0.f, o = 0.f;
        for (int s = 0; s < p.nsplit; ++s) {
            if (ms[s] > -INFINITY) {
                const float* pp = scratch + SC_PART + (h * MAX_SPLIT + s) * PART_STRIDE;
                float f = exp2f(ms[s] - m);
                l += f * __ldcg(pp + 1);
                o += f * __ldcg(pp + 2 + d);
            }
        }
        c.sm->xs[t] = o / l;
    }
    cons_sync();
}

// ---- warp-level top-k pick among n (value, id) candidates held in shared memory ------------------
// Returns (in every lane) the sampled id.  topk (UMGen.py:899-913) + sfmx_temp_sampling (:967-974):
// keep the k largest, softmax(v / temp), inverse-CDF draw with uniform u.  k == 1 is the arg-max with
// the lowest id winning ties.  Values are destroyed.
__device__ int warp_topk_sample(float* vals, const int* ids, int n, int k, float inv_temp, float u, int lane) {
    float selv = -INFINITY;
    int seli = 0x7fffffff;
    for (int r = 0; r < k; ++r) {
        float bv = -INFINITY;
        int bi = 0x7fffffff, bp = -1;
        for (int i = lane; i < n; i += 32) {
            float v = vals[i];
            int id = ids ? ids[i] : i;
            if (v > bv || (v == bv && id < bi)) { bv = v; bi = id; bp = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            int op = __shfl_xor_sync(0xffffffffu, bp, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; bp = op; }
        }
        if (lane == r) { selv = bv; seli = bi; }
        if (lane == 0 && bp >= 0) vals[bp] = -INFINITY;
        __syncwarp();
    }
    const float vmax = __shfl_sync(0xffffffffu, selv, 0);
    float w = (lane < k && selv > -INFINITY) ? __expf((selv - vmax) * inv_temp) : 0.f;
    float cum = w;      // inclusive prefix over lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_up_sync(0xffffffffu, cum, o);
        if (lane >= o) cum += t;
    }
    const float total = __shfl_sync(0xffffffffu, cum, 31);
    const float target = u * total;
    unsigned hit = __ballot_sync(0xffffffffu, (w > 0.f) && (cum > target));
    int pick = hit ? (__ffs(hit) - 1) : 0;
    return __shfl_sync(0xffffffffu, seli, pick);
}

// ---- rotated-box collision (reference plugin/misc/misc.py:203-311), float32 corners --------------
__device__ __forceinline__ bool ccw_gt(const float* p, const float* q, const float* r) {
    return __fmul_rn(r[1] - p[1], q[0] - p[0]) > __fmul_rn(q[1] - p[1], r[0] - p[0]);
}
__device__ bool inside_all(const float* outer, const float* inner) {
    for (int l = 0; l < 4; ++l)
        for (int k = 0; k < 4; ++k) {
            const float* a = outer + 2 * k;
            const float* b = outer + 2 * ((k + 1) & 3);
            float vx = -(a[0] - b[0]), vy = -(a[1] - b[1]);
            float cross = __fmul_rn(vy, a[0] - inner[2 * l]);
            cross = __fsub_rn(cross, __fmul_rn(vx, a[1] - inner[2 * l + 1]));
            if (cross >= 0.f) return false;
        }
    return true;
}
__device__ bool pair_collides(const float* a, const float* b) {
    float axmin = fminf(fminf(a[0], a[2]), fminf(a[4], a[6])), axmax = fmaxf(fmaxf(a[0], a[2]), fmaxf(a[4], a[6]));
    float aymin = fminf(fminf(a[1], a[3]), fminf(a[5], a[7])), aymax = fmaxf(fmaxf(a[1], a[3]), fmaxf(a[5], a[7]));
    float bxmin = fminf(fminf(b[0], b[2]), fminf(b[4], b[6])), bxmax = fmaxf(fmaxf(b[0], b[2]), fmaxf(b[4], b[6]));
    float bymin = fminf(fminf(b[1], b[3]), fminf(b[5], b[7])), bymax = fmaxf(fmaxf(b[1], b[3]), fmaxf(b[5], b[7]));
    if (!(fminf(axmax, bxmax) - fmaxf(axmin, bxmin) > 0.f)) return false;
    if (!(fminf(aymax, bymax) - fmaxf(aymin, bymin) > 0.f)) return false;
    for (int k = 0; k < 4; ++k) {
        const float* A = a + 2 * k;
        const float* B = a + 2 * ((k + 1) & 3);
        for (int l = 0; l < 4; ++l) {
            const float* Cc = b + 2 * l;
            const float* Dd = b + 2 * ((l + 1) & 3);
            if (ccw_gt(A, Cc, Dd) != ccw_gt(B, Cc, Dd) && ccw_gt(A, B, Cc) != ccw_gt(A, B, Dd)) return true;
        }
    }
    if (inside_all(a, b)) return true;
    return inside_all(b, a);
}
// corners of (x, y, l, w, yaw) as bbox3d2bevcorners (misc.py:143-177) after check_collision negates yaw (:609)
__device__ void box_corners(double x, double y, double l, double w, double yaw, float* out) {
    const double ang = -yaw;
    const double s = sin(ang), co = cos(ang);
    const float ux[4] = {-0.5f, -0.5f, 0.5f, 0.5f}, uy[4] = {-0.5f, 0.5f, 0.5f, -0.5f};
    for (int i = 0; i < 4; ++i) {
        double cx = (double)ux[i] * l, cy = (double)uy[i] * w;
        // row-vector times rot_mat^T as laid out by np.transpose(rot_mat, (2, 1, 0)): [[cos, sin], [-sin, cos]]
        double rx = cx * co + cy * (-s);
        double ry = cx * s + cy * co;
        out[2 * i] = (float)(rx + x);
        out[2 * i + 1] = (float)(ry + y);
    }
}

// bbox3d post-processing of one sampled token by warp 0 (UMGen.py:1071-1129, 1275-1383).
// Returns the final token; may wipe the slot (ids rewritten by the caller through *wipe).
__device__ int bbox_rules(Ctx& c, int q, int tok, float u2, bool* wipe) {
    const KParams& p = *c.p;
    Smem* sm = c.sm;
    const int lane = c.lane;
    *wipe = false;
    const int bidx = q - BBOX_FIRST_POS - 1;
    const int prev = __ldg((const int*)p.a.prev_bbox_i32 + bidx);
    const int slot_of_q = (q - BBOX_FIRST_POS) / 11;
    const bool controlled = (p.a.control_mask >> slot_of_q) & 1ull;
    const float inv_temp = 1.0f / (float)p.a.temperature;
    int* status = (int*)p.a.status_i32;
    const bool resample_on_pad = p.a.merge_ar_tar && prev != PAD_TOKEN;
    if (controlled || (tok == PAD_TOKEN && resample_on_pad)) {
        const float* row = (const float*)p.a.tar_bbox_logits_f + (size_t)bidx * 1028;
        float* tmp = sm->candv;        // AR candidates are already consumed
        if (controlled) {              // UMGen.py:1083-1089: TAR head with <pad> masked
            for (int i = lane; i < 1028; i += 32) tmp[i] = (i == 1027) ? -INFINITY : __ldg(row + i);
            __syncwarp();
            const float u1 = philox_uniform(p.a.seed, (uint32_t)p.a.frame_index, (uint32_t)q, 1u);
            tok = warp_topk_sample(tmp, nullptr, 1028, (int)p.a.top_k_bbox, inv_temp, u1, lane);
        }
        if (tok == PAD_TOKEN && resample_on_pad) {                         // UMGen.py:1092-1104
            for (int i = lane; i < 1028; i += 32) tmp[i] = __ldg(row + i);
            __syncwarp();
            tok = warp_topk_sample(tmp, nullptr, 1028, (int)p.a.top_k_bbox, inv_temp, u2, lane);
            if (lane == 0 && c.cta == 0) atomicAdd(status + 2, 1);
        }
    }
    // rule_based_constraint at the slot's 11th token (UMGen.py:1295-1383)
    if (p.a.rule_constrain && tok != PAD_TOKEN && (q - BBOX_FIRST_POS) % 11 == 0) {
        const double* lut = (const double*)p.a.box_lut_d;
        int nb = sm->nbox;
        if (nb == 0) {
            if (lane == 0) { box_corners(0.0, 0.0, 5.176, 2.297, 0.0, sm->corners[0]); sm->box_dropped[0] = 0; }
            nb = 1;
        }
        if (lane == 0) {
            int t[10];
            for (int i = 0; i < 10; ++i) t[i] = sm->recent[(q - 10 + i) & 15];
            double x = lut[t[0] * 10 + 0], y = lut[t[1] * 10 + 1], l = lut[t[3] * 10 + 3], w = lut[t[4] * 10 + 4],
                   yaw = lut[t[6] * 10 + 6];
            box_corners(x, y, l, w, yaw, sm->corners[nb]);
            sm->box_dropped[nb] = (x >= 63.0) ? 1 : 0;
        }
        nb += 1;
        __syncwarp();
        // query = last kept box; collide against every kept box (including itself, which never hits)
        int qi = -1, kept = 0;
        for (int i = 0; i < nb; ++i) if (!sm->box_dropped[i]) { qi = i; kept++; }
        bool hit = false;
        if (kept > 1) {
            for (int i = lane; i < nb; i += 32)
                if (!sm->box_dropped[i] && pair_collides(sm->corners[i], sm->corners[qi])) hit = true;
        }
        hit = __any_sync(0xffffffffu, hit);
        const bool was_pad = (prev == PAD_TOKEN);
        if (was_pad && (hit || nb > 30)) {
            *wipe = true;
            tok = PAD_TOKEN;
            nb -= 1;
            if (lane == 0 && c.cta == 0) atomicAdd(status + 1, 1);
        }
        if (lane == 0) sm->nbox = nb;
        __syncwarp();
    }
    return tok;
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(N_THREADS, 1) decode_frame_kernel(const __grid_constant__ KParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    Smem* sm = reinterpret_cast<Smem*>(smem_raw);
    const UmgenDecodeArgs& a = p.a;
    Ctx c;
    c.p = &p; c.sm = sm; c.abort_flag = (int*)a.status_i32;
    c.deadline = globaltimer_ns() + TIMEOUT_NS;
    c.cta = blockIdx.x; c.tid = threadIdx.x; c.warp = threadIdx.x >> 5; c.lane = threadIdx.x & 31;
    c.epoch = 0;
    float* scratch = (float*)a.scratch_f;
    c.bar = (uint32_t*)(scratch + SC_BAR);
    const int G = p.grid, L = (int)a.n_layer;
    const int n_steps = (int)a.n_steps;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSLOT; ++i) { mbar_init(&sm->full[i], 1); mbar_init(&sm->empty[i], 1); }
        sm->progress = 0; sm->nbox = 0; sm->tok = 0; sm->dead = 0;
        mbar_fence_init();
    }
    __syncthreads();

    const __half* Wl = (const __half*)a.oar_h;
    const float* Fl = (const float*)a.oar_f;
    const __half* heads[3] = {(const __half*)a.head_map_h, (const __half*)a.head_bbox_h, (const __half*)a.head_img_h};
    const __half* gfc[3] = {(const __half*)a.map_fc_h, nullptr, (const __half*)a.img_fc_h};
    const __half* gproj[3] = {(const __half*)a.map_proj_h, nullptr, (const __half*)a.img_proj_h};
    const float* books[3] = {(const float*)a.map_codebook_f, nullptr, (const float*)a.img_codebook_f};

    int rq0, rq1, rp0, rp1, rf0, rf1, rg0, rg1;
    row_slice(3 * C, c.cta, G, rq0, rq1);
    row_slice(C, c.cta, G, rp0, rp1);
    row_slice(FF, c.cta, G, rf0, rf1);
    row_slice(FF, c.cta, G, rg0, rg1);     // GMLP c_fc rows (3072)

    // ============================== producer warp ==============================================
    if (c.warp == N_CONS_WARPS) {
        if (c.lane != 0 || a.mode == 1) return;
        Producer pr;
        pr.c = &c;
        uint32_t bars_before = 1;      // init barrier
        uint32_t bars_prev = 0;
        for (int j = 0; j < n_steps; ++j) {
            const int q = j + 1;       // position produced by this step
            int h = 0, s = 0, k0 = 0, k1 = 0;
            const bool att_active = c.cta < NH * p.nsplit;
            if (att_active) { h = c.cta / p.nsplit; s = c.cta - h * p.nsplit; kv_range(j, p.nsplit, s, k0, k1); }
            for (int l = 0; l < L; ++l) {
                const uint8_t* wl = (const uint8_t*)(Wl + (size_t)l * LAYER_H);
                pr.issue_rows(wl + (size_t)OFF_QKV * 2, C * 2, rq0, rq1);
                if (att_active && k1 > k0) {
                    // row j-1 of this layer was written in phase 1 of the previous step
                    pr.wait_progress(bars_prev + 5u * l + 1u);
                    const __half* kb = (const __half*)a.kv_h + ((size_t)(l * 2 + 0) * NH + h) * SMAX * HD;
                    const __half* vb = (const __half*)a.kv_h + ((size_t)(l * 2 + 1) * NH + h) * SMAX * HD;
                    const int nk = k1 - k0, has_new = (s == 0);
                    const int nblocks = (nk + has_new + KV_BLOCK - 1) / KV_BLOCK;
                    for (int b = 0; b < nblocks; ++b) {
                        int staged = max(0, min(KV_BLOCK, nk - b * KV_BLOCK));
                        if (staged > 0) {
                            pr.issue(kb + (size_t)(k0 + b * KV_BLOCK) * HD, (uint32_t)staged * HD * 2);
                            pr.issue(vb + (size_t)(k0 + b * KV_BLOCK) * HD, (uint32_t)staged * HD * 2);
                        }
                    }
                }
                pr.issue_rows(wl + (size_t)OFF_PROJ * 2, C * 2, rp0, rp1);
                pr.issue_rows(wl + (size_t)OFF_FC * 2, C * 2, rf0, rf1);
                pr.issue_rows(wl + (size_t)OFF_PROJ2 * 2, FF * 2, rp0, rp1);
            }
            uint32_t nb = 5u * L;
            if (needs_head(q)) {
                const int mod = pos_mod(q);
                int r0, r1;
                row_slice(vocab_of(mod), c.cta, G, r0, r1);
                pr.issue_rows((const uint8_t*)heads[mod], C * 2, r0, r1);
                nb += 1;
            }
            if (j == SEQ - 2) {
                // last step of the frame: no next input
            } else if (needs_gmlp(q)) {
                const int mod = pos_mod(q);
                if (rg1 > rg0) pr.issue((const uint8_t*)gfc[mod] + (size_t)rg0 * 32, (uint32_t)(rg1 - rg0) * 32);
                pr.issue_rows((const uint8_t*)gproj[mod], FF * 2, rp0, rp1);
                nb += 2;
            } else {
                nb += 1;
            }
            bars_prev = bars_before;
            bars_before += nb;
            if (*(volatile int*)c.abort_flag != 0) return;
        }
        return;
    }

    // ============================== consumer warps =============================================
    const float* tar = (const float*)a.tar_feat_f;
    float* x = scratch + SC_X;
    int* out_tokens = (int*)a.out_tokens_i32;
    int* picks = (int*)a.picks_i32;
    const int* pose_tok = (const int*)a.pose_tok_i32;
    const int* teacher = (const int*)a.teacher_i32;

    // input of step 0: task embedding + TAR feature of index 0 (UMGen.py:1175,1215,1231)
    for (int r = rp0 + c.tid; r < rp1; r += N_CONS) x[r] = __ldg((const float*)a.tske_f + r) + __ldg(tar + r);
    if (c.cta == 0 && c.tid < 8) {
        // given prefix and forced ids
        const int qs[8] = {1, 5, 6, 1031, 1032, 1693, 1694, 2207};
        out_tokens[qs[c.tid] - 1] = forced_id(qs[c.tid]);
        picks[qs[c.tid] - 1] = forced_id(qs[c.tid]);
        if (c.tid < 3) { out_tokens[1 + c.tid] = pose_tok[c.tid]; picks[1 + c.tid] = pose_tok[c.tid]; }
    }
    grid_barrier(c);

    for (int j = 0; j < n_steps; ++j) {
        const int q = j + 1;
        for (int l = 0; l < L; ++l) {
            const __half* wl = Wl + (size_t)l * LAYER_H;
            const float* fl = Fl + (size_t)l * LAYER_F;
            // ---- phase 1: LN1 -> c_attn (+bias) -> q to scratch, k/v appended to the cache (module.py:206-210)
            layer_norm_to_smem(c, x, fl + F_LN1);
            gemv_slice768(c, wl + OFF_QKV, rq0, rq1);
            for (int r = rq0 + c.tid; r < rq1; r += N_CONS) {
                float v = sm->acc[r - rq0] + __ldg(fl + F_BQKV + r);
                if (r < C) scratch[SC_Q + r] = v;
                else {
                    const int which = (r < 2 * C) ? 0 : 1;
                    const int cc = r - (which + 1) * C;
                    const int hh = cc / HD, d = cc - hh * HD;
                    __half* dst = (__half*)a.kv_h + (((size_t)(l * 2 + which) * NH + hh) * SMAX + j) * HD + d;
                    *dst = __float2half_rn(v);
                }
            }
            grid_barrier(c);
            // ---- phase 2: split-KV attention over rows 0..j
            attention_phase(c, l, j);
            grid_barrier(c);
            // ---- phase 3: merge partials -> c_proj (+bias) -> residual (module.py:227-229, 409)
            combine_partials(c, j);
            gemv_slice768(c, wl + OFF_PROJ, rp0, rp1);
            for (int r = rp0 + c.tid; r < rp1; r += N_CONS) x[r] = __ldcg(x + r) + sm->acc[r - rp0] + __ldg(fl + F_BPROJ + r);
            grid_barrier(c);
            // ---- phase 4: LN2 -> c_fc -> erf-GELU (module.py:245-247)
            layer_norm_to_smem(c, x, fl + F_LN2);
            gemv_slice768(c, wl + OFF_FC, rf0, rf1);
            for (int r = rf0 + c.tid; r < rf1; r += N_CONS) scratch[SC_H + r] = gelu_erf(sm->acc[r - rf0]);
            grid_barrier(c);
            // ---- phase 5: mlp c_proj -> residual (module.py:248, 410)
            for (int i = c.tid; i < FF / 4; i += N_CONS)
                reinterpret_cast<float4*>(sm->hs)[i] = __ldcg(reinterpret_cast<const float4*>(scratch + SC_H) + i);
            cons_sync();
            gemv_slice3072(c, wl + OFF_PROJ2, rp0, rp1);
            for (int r = rp0 + c.tid; r < rp1; r += N_CONS) {
                const float* ap = sm->acc + (r - rp0) * 3;
                x[r] = __ldcg(x + r) + (ap[0] + ap[1] + ap[2]);
            }
            grid_barrier(c);
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
            int r0, r1;
            row_slice(V, c.cta, G, r0, r1);
            layer_norm_to_smem(c, x, (const float*)a.ln_oar_f);
            gemv_slice768(c, heads[mod], r0, r1);
            if (a.logits_dump_f) {
                float* dump = (float*)a.logits_dump_f + (size_t)(q - 1) * 8192;
                for (int r = r0 + c.tid; r < r1; r += N_CONS) dump[r] = sm->acc[r - r0];
                cons_sync();           // warp 0 overwrites acc while selecting
            }
            if (c.warp == 0) {     // local top-k of this CTA's slice
                float* cv = scratch + SC_CANDV + c.cta * MAX_CAND;
                int* ci = (int*)(scratch + SC_CANDI) + c.cta * MAX_CAND;
                const int n = r1 - r0;
                for (int r = 0; r < k; ++r) {
                    float bv = -INFINITY;
                    int bi = 0x7fffffff;
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
                        cv[r] = bv;
                        ci[r] = (bi == 0x7fffffff) ? 0x7fffffff : r0 + bi;
                        if (bi != 0x7fffffff) sm->acc[bi] = -INFINITY;
                    }
                    __syncwarp();
                }
            }
            grid_barrier(c);
            // every CTA merges all candidates and decides the token identically
            const int ncand = G * k;
            for (int i = c.tid; i < ncand; i += N_CONS) {
                const int cta_i = i / k, r = i - cta_i * k;
                sm->candv[i] = __ldcg(scratch + SC_CANDV + cta_i * MAX_CAND + r);
                sm->candi[i] = __ldcg((const int*)(scratch + SC_CANDI) + cta_i * MAX_CAND + r);
            }
            cons_sync();
            if (c.warp == 0) {
                const float u0 = philox_uniform(a.seed, (uint32_t)a.frame_index, (uint32_t)q, 0u);
                int t = warp_topk_sample(sm->candv, sm->candi, ncand, k, 1.0f / (float)a.temperature, u0, c.lane);
                bool wipe = false;
                if (mod == 1) {
                    const float u2 = philox_uniform(a.seed, (uint32_t)a.frame_index, (uint32_t)q, 2u);
                    t = bbox_rules(c, q, t, u2, &wipe);
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
        int tok_used = tok;
        if (teacher != nullptr && q > 5 && fid < 0) tok_used = __ldg(teacher + (q - 1));
        if (c.tid == 0) {
            sm->recent[q & 15] = tok_used;
            if (c.cta == 0 && q > 5) { out_tokens[q - 1] = tok_used; picks[q - 1] = tok; }
        }
        if (j == SEQ - 2) break;           // q = 2206 was the last sampled token; q = 2207 is forced

        // ---- embed the token as the next input and add the TAR feature of index j+1 (UMGen.py:1215-1231)
        const float* tnext = tar + (size_t)(j + 1) * C;
        if (needs_gmlp(q)) {
            const int mod = pos_mod(q);
            if (c.tid < 16) sm->code[c.tid] = __ldg(books[mod] + (size_t)tok_used * 16 + c.tid);
            cons_sync();
            if (rg1 > rg0) {
                Stage st;
                const uint8_t* w = acquire(c, (const uint8_t*)gfc[mod] + (size_t)rg0 * 32, (uint32_t)(rg1 - rg0) * 32, st);
                if (c.tid < rg1 - rg0) {
                    const uint4* wp = reinterpret_cast<const uint4*>(w + (size_t)c.tid * 32);
                    uint4 w0 = wp[0], w1 = wp[1];
                    const float* cd = sm->code;
                    float s = dot8(w0, make_float4(cd[0], cd[1], cd[2], cd[3]), make_float4(cd[4], cd[5], cd[6], cd[7])) +
                              dot8(w1, make_float4(cd[8], cd[9], cd[10], cd[11]), make_float4(cd[12], cd[13], cd[14], cd[15]));
                    scratch[SC_H + rg0 + c.tid] = gelu_erf(s);
                }
                cons_sync();
                release(c, st);
            }
            grid_barrier(c);
            for (int i = c.tid; i < FF / 4; i += N_CONS)
                reinterpret_cast<float4*>(sm->hs)[i] = __ldcg(reinterpret_cast<const float4*>(scratch + SC_H) + i);
            cons_sync();
            gemv_slice3072(c, gproj[mod], rp0, rp1);
            for (int r = rp0 + c.tid; r < rp1; r += N_CONS) {
                const float* ap = sm->acc + (r - rp0) * 3;
                x[r] = (ap[0] + ap[1] + ap[2]) + __ldg(tnext + r);
            }
            grid_barrier(c);
        } else {
            const float* row;
            const int qn = q;   // embedding of token at position q
            if (forced_id(qn) >= 0) row = (const float*)a.axe_f + (size_t)forced_id(qn) * C;
            else if (qn <= 5) row = (const float*)a.fpe_f + (size_t)tok_used * C;
            else row = (const float*)a.be_f + (size_t)tok_used * C;
            for (int r = rp0 + c.tid; r < rp1; r += N_CONS) x[r] = __ldg(row + r) + __ldg(tnext + r);
            grid_barrier(c);
        }
        if (sm->dead || *(volatile int*)c.abort_flag != 0) break;
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
namespace umgen { extern int64_t g_launches; }
using namespace umgen;

extern "C" int64_t umgen_decode_scratch_floats(void) { return SC_TOTAL; }

extern "C" int umgen_decode_frame(const UmgenDecodeArgs* args, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    if (!args) { set_error("null args"); return -1; }
    if (args->n_layer < 1 || args->n_layer > 256) { set_error("n_layer out of range: %lld", (long long)args->n_layer); return -1; }
    if (args->n_steps < 1 || args->n_steps > SEQ - 1) { set_error("n_steps must be in [1, 2206]"); return -1; }
    const int64_t ks[3] = {args->top_k_map, args->top_k_bbox, args->top_k_img};
    for (int i = 0; i < 3; ++i)
        if (ks[i] < 1 || ks[i] > MAX_CAND) { set_error("top_k must be in [1, 16] (got %lld)", (long long)ks[i]); return -1; }
    if (!(args->temperature > 0)) { set_error("temperature must be > 0"); return -1; }
    if (!args->tar_bbox_logits_f && (args->merge_ar_tar || args->control_mask)) { set_error("tar_bbox_logits required"); return -1; }
    if (!args->kv_h || !args->scratch_f || !args->out_tokens_i32 || !args->picks_i32 || !args->status_i32 || !args->tar_feat_f) {
        set_error("null buffer"); return -1;
    }
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
    UMGEN_CUDA_OK(cudaMemsetAsync(args->status_i32, 0, 8 * sizeof(int), stream));
    void* kargs[] = {&kp};
    UMGEN_CUDA_OK(cudaLaunchCooperativeKernel((void*)decode_frame_kernel, dim3(kp.grid), dim3(N_THREADS), kargs, smem, stream));
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
