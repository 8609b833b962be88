// Pieces shared by the two OAR decode kernels (decode.cu: L2-exchange kernel, decode_cluster.cu: cluster kernel):
// the static schedule of a frame, the top-k / nucleus samplers and the bbox3d rule path
// (reference models/UMGen.py:899-992, 1071-1129, 1275-1383; plugin/misc/misc.py:143-311).
// The sampler / rule functions are templates over the kernel's shared-memory struct, which must provide
//   float red[64]; volatile int tok; int nbox; float corners[MAX_BOX][8]; int box_dropped[MAX_BOX]; int recent[16];
//   float stage[>= 1028];
#pragma once
#include "common.cuh"
#include "../../include/umgen.h"

namespace umgen {

#ifndef UMGEN_CONS_WARPS
#define UMGEN_CONS_WARPS 15                      // decode.cu: 15 consumer + 1 producer warp = 512 threads -> 128 registers each
#endif
constexpr int N_CONS_WARPS = UMGEN_CONS_WARPS;
constexpr int N_CONS = N_CONS_WARPS * 32;        // consumer threads
constexpr int N_THREADS = N_CONS + 32;           // + producer warp
constexpr int MAX_CAND = 16;
constexpr int MAX_BOX = 64;

__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, %0;" ::"n"(N_CONS) : "memory"); }

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
__host__ __device__ __forceinline__ int vocab_of(int mod) { return mod == 1 ? 1028 : 8192; }

// ---- warp-level top-k pick among n (value, id) candidates held in shared memory ------------------
// Returns (in every lane) the sampled id.  topk (UMGen.py:899-913) + sfmx_temp_sampling (:967-974):
// keep the k largest, softmax(v / temp), inverse-CDF draw with uniform u.  k == 1 is the arg-max with
// the lowest id winning ties.  Values are destroyed.
static __device__ __noinline__ int warp_topk_sample(float* vals, const int* ids, int n, int k, float inv_temp, float u, int lane) {
    float selv = -INFINITY;
    int seli = 0x7fffffff;
#pragma unroll 1
    for (int r = 0; r < k; ++r) {
        float bv = -INFINITY;
        int bi = 0x7fffffff, bp = -1;
#pragma unroll 1
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

// ---- block-level nucleus (top-p) sampler: sample_top_p (UMGen.py:915-965) ---------------------------------
// Every consumer thread holds up to TOPP_PER values (v[i] belongs to id tid + i * N_CONS, -inf if absent).
// softmax(v / temp); a token is kept iff the probability mass of strictly more likely tokens is <= p (the
// reference's `(cumsum - p_sorted) > p` mask on the descending sort); one draw from the renormalised kept
// set by inverse CDF with uniform u (order: thread-major, any fixed order is distributionally equivalent).
constexpr int TOPP_PER = (8192 + N_CONS - 1) / N_CONS;      // 18
__device__ __forceinline__ float block_sum(float x, float* red, int warp, int lane) {
    x = warp_sum(x);
    if (lane == 0) red[warp] = x;
    cons_sync();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < N_CONS_WARPS; ++w) t += red[w];
    cons_sync();
    return t;
}
template <class S>
__device__ __noinline__ int block_topp_sample(S* sm, float (&v)[TOPP_PER], float p, float inv_temp, float u, int tid) {
    const int warp = tid >> 5, lane = tid & 31;
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < TOPP_PER; ++i) m = fmaxf(m, v[i]);
    m = warp_max(m);
    if (lane == 0) sm->red[warp] = m;
    cons_sync();
    m = sm->red[0];
#pragma unroll
    for (int w = 1; w < N_CONS_WARPS; ++w) m = fmaxf(m, sm->red[w]);
    cons_sync();
    float e[TOPP_PER], z = 0.f;
#pragma unroll
    for (int i = 0; i < TOPP_PER; ++i) { e[i] = (v[i] > -INFINITY) ? __expf((v[i] - m) * inv_temp) : 0.f; z += e[i]; }
    const float Z = block_sum(z, sm->red, warp, lane);
    const float budget = p * Z;
    // smallest threshold (as a bit pattern) whose strictly-greater mass fits the budget
    uint32_t lo = 0u, hi = 0x3f800000u;          // e <= 1
    {
        float f0 = 0.f;
#pragma unroll
        for (int i = 0; i < TOPP_PER; ++i) f0 += (e[i] > 0.f) ? e[i] : 0.f;
        if (block_sum(f0, sm->red, warp, lane) <= budget) hi = 0u;      // p >= 1: everything with e > 0 ... keep all
    }
#pragma unroll 1
    while (hi > lo + 1u && hi != 0u) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        const float tau = __uint_as_float(mid);
        float f = 0.f;
#pragma unroll
        for (int i = 0; i < TOPP_PER; ++i) f += (e[i] > tau) ? e[i] : 0.f;
        if (block_sum(f, sm->red, warp, lane) <= budget) hi = mid; else lo = mid;
    }
    const float tau = __uint_as_float(hi);
    float ksum = 0.f;
#pragma unroll
    for (int i = 0; i < TOPP_PER; ++i) { if (!(e[i] >= tau && e[i] > 0.f)) e[i] = 0.f; ksum += e[i]; }
    // exclusive prefix of per-thread kept mass, thread-major order
    float incl = ksum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const float t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) sm->red[32 + warp] = incl;
    cons_sync();
    float base = 0.f, total = 0.f;
#pragma unroll
    for (int w = 0; w < N_CONS_WARPS; ++w) { if (w < warp) base += sm->red[32 + w]; total += sm->red[32 + w]; }
    const float target = fminf(u, 0.99999994f) * total;
    const float start = base + incl - ksum;
    if (tid == 0) sm->tok = -1;
    cons_sync();
    if (ksum > 0.f && target >= start && target < start + ksum) {
        float run = start;
        int pick = -1;
#pragma unroll
        for (int i = 0; i < TOPP_PER; ++i) {
            if (e[i] > 0.f && (pick < 0 || target >= run)) { pick = tid + i * N_CONS; run += e[i]; }
        }
        sm->tok = pick;
    }
    cons_sync();
    int tok = sm->tok;
    if (tok < 0) {      // rounding left the target past the last kept element: take the most likely token
        float best = -INFINITY; int bi = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < TOPP_PER; ++i) if (v[i] > best) { best = v[i]; bi = tid + i * N_CONS; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) { sm->red[warp] = best; reinterpret_cast<int*>(sm->red)[32 + warp] = bi; }
        cons_sync();
        if (tid == 0) {
            float b = -INFINITY; int id = 0;
            for (int w = 0; w < N_CONS_WARPS; ++w) { const float x = sm->red[w]; const int xi = reinterpret_cast<int*>(sm->red)[32 + w]; if (x > b || (x == b && xi < id)) { b = x; id = xi; } }
            sm->tok = id;
        }
        cons_sync();
        tok = sm->tok;
    }
    cons_sync();
    return tok;
}

// ---- rotated-box collision (reference plugin/misc/misc.py:203-311) --------------------------------------
// float32 on purpose: bbox3d2bevcorners computes the corners in float64 and returns them `.astype(np.float32)` (misc.py:177), the aligned
// boxes are float32 too (misc.py:192), so numba's box_collision_test does float32 arithmetic with separately rounded products (no
// contraction without fastmath) -- reproduced here with __fmul_rn / __fsub_rn.  Pinned on the device by umgen_check_collision over the
// 912 answers of tests/golden/collision.npz.
__device__ __forceinline__ bool ccw_gt(const float* p, const float* q, const float* r) {
    return __fmul_rn(r[1] - p[1], q[0] - p[0]) > __fmul_rn(q[1] - p[1], r[0] - p[0]);
}
static __device__ bool inside_all(const float* outer, const float* inner) {
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
static __device__ __noinline__ bool pair_collides(const float* a, const float* b) {
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
static __device__ __noinline__ void box_corners(double x, double y, double l, double w, double yaw, float* out) {
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

// BoxOverlap.check_collision(boxes, fliter=True) (misc.py:591-630) on the warp's box list: boxes with x >= 63 are dropped (fliter_and_map_object,
// misc.py:475-481), the query is the last kept box, tested against every kept box (itself included, which never hits).  Result in every lane.
static __device__ __forceinline__ bool last_box_collides(const float (*corners)[8], const int* dropped, int nb, int lane) {
    int qi = -1, kept = 0;
    for (int i = 0; i < nb; ++i) if (!dropped[i]) { qi = i; kept++; }
    bool hit = false;
    if (kept > 1) {
        for (int i = lane; i < nb; i += 32)
            if (!dropped[i] && pair_collides(corners[i], corners[qi])) hit = true;
    }
    return __any_sync(0xffffffffu, hit);
}

// bbox3d post-processing of one sampled token by warp 0 (UMGen.py:1071-1129, 1275-1383).
// Returns the final token, with WIPE_BIT set when the slot must be wiped (ids rewritten by the caller).
constexpr int WIPE_BIT = 1 << 30;     // set in bbox_rules' return value when the slot must be rewritten to <pad>
template <class S>
__device__ __noinline__ int bbox_rules(S* sm, const UmgenDecodeArgs& pa, int lane, int cta, int q, int tok, float u2, bool skip_resample) {
    struct { const UmgenDecodeArgs& a; } p{pa};
    struct { int cta; } c{cta};
    bool wipe_flag = false;
    bool* wipe = &wipe_flag;
    const int bidx = q - BBOX_FIRST_POS - 1;
    const int prev = __ldg((const int*)p.a.prev_bbox_i32 + bidx);
    const int slot_of_q = (q - BBOX_FIRST_POS) / 11;
    const bool controlled = (p.a.control_mask >> slot_of_q) & 1ull;
    const float inv_temp = 1.0f / (float)p.a.temperature;
    int* status = (int*)p.a.status_i32;
    const bool resample_on_pad = p.a.merge_ar_tar && prev != PAD_TOKEN;
    if (!skip_resample && (controlled || (tok == PAD_TOKEN && resample_on_pad))) {
        const float* row = (const float*)p.a.tar_bbox_logits_f + (size_t)bidx * 1028;
        float* tmp = sm->stage;        // AR candidates are already consumed
        if (controlled) {              // UMGen.py:1083-1089: TAR head with <pad> masked
            for (int i = lane; i < 1028; i += 32) tmp[i] = (i == 1027) ? -INFINITY : __ldcg(row + i);      // coherent: the rows may be written while the kernel runs (tar_ready)
            __syncwarp();
            const float u1 = philox_uniform(p.a.seed, (uint32_t)p.a.frame_index, (uint32_t)q, 1u);
            tok = warp_topk_sample(tmp, nullptr, 1028, (int)p.a.top_k_bbox, inv_temp, u1, lane);
        }
        if (tok == PAD_TOKEN && resample_on_pad) {                         // UMGen.py:1092-1104
            for (int i = lane; i < 1028; i += 32) tmp[i] = __ldcg(row + i);
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
        const bool hit = last_box_collides(sm->corners, sm->box_dropped, nb, lane);
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
    return tok | (wipe_flag ? WIPE_BIT : 0);
}

}  // namespace umgen
