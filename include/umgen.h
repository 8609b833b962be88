/* libumgen_sm100 -- C ABI of the B200-native UMGen next-scene decode engine.
 *
 * Drop-in boundary for the hot path of YanhaoWu/UMGen.  The reference has no
 * FFI of its own (it is pure PyTorch); each entry point below replaces the body of one reference
 * method and is bound from Python with ctypes (umgen_b200/capi.py; INTEGRATION.md shows the stub a
 * reference maintainer would add).  Paths cited are relative to /root/reference/projects.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is DEVICE memory unless the name ends in _host
 *   - the caller (PyTorch) owns every buffer; the library allocates nothing and keeps no state
 *   - return 0 on success, negative on error; umgen_last_error() returns a thread-local message
 *   - calls are stream-ordered on the cudaStream_t passed as `stream` (void* here); no host sync
 *     inside any call unless documented
 *   - matrices are row-major [out_features][in_features] fp16 ("h"), vectors/tables fp32 ("f")
 */
#ifndef UMGEN_H_
#define UMGEN_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UMGEN_ABI_VERSION 16

/* geometry (configs/UMGen_config_evaluation.py:27-38,284-290) */
#define UMGEN_TAR_LATE_ROW0 1031 /* first sequence position whose conditioning feature comes from the box_tar pass (bos of the bbox3d block) */
#define UMGEN_C 768
#define UMGEN_HEADS 16
#define UMGEN_HEAD_DIM 48
#define UMGEN_SEQ 2207
#define UMGEN_KV_ROWS 2304 /* rows allocated per (layer, k|v, head): 8 owners x 18 tiles x 16 keys for the cluster kernel, >= 2208 for the L2-exchange kernel */

/* fp16 elements of one packed OAR layer: c_attn[2304][768] | c_proj[768][768] | c_fc[3072][768] | mlp c_proj[768][3072] */
#define UMGEN_OAR_LAYER_H (2304 * 768 + 768 * 768 + 3072 * 768 + 768 * 3072)
/* fp32 elements of one packed OAR layer: ln_1[768] | c_attn.bias[2304] | c_proj.bias[768] | ln_2[768] */
#define UMGEN_OAR_LAYER_F (768 + 2304 + 768 + 768)

int umgen_abi_version(void);
const char* umgen_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
int64_t umgen_launch_count(void);
/* force every kernel of the library to be loaded on the current device (CUDA loads kernels lazily on first launch, and that load waits for an
 * idle device: it would deadlock against the persistent decode kernel while it spins on tar_ready_i32) */
int umgen_preload(void);

/* ------------------------------------------------------------------------------------------------
 * OAR decode of one frame: replaces UMGen.infer_oar_net + sample_next_token + rule_based_constraint
 * (models/UMGen.py:1151-1273, 1029-1139, 1275-1383) and the BlockOAR / CausalFlashAttention / MLP /
 * LayerNorm forward passes it drives (models/module.py:378-428, 179-230, 233-250, 26-37).
 * One persistent kernel runs all 2206 single-token steps of the frame.  Two kernels implement it:
 *   - the cluster kernel (csrc/decode_cluster.cu, default): 8 thread-block clusters x 8 CTAs, tensor-core GEMVs on fragment-packed
 *     weights, head-local exchanges over distributed shared memory, 2 L2 hops per layer; needs
 *     umgen_decode_cluster_capacity() >= 8 and oar_cl_h
 *   - the L2-exchange kernel (csrc/decode.cu): one CTA per SM, every exchange through tagged lines in L2; the fallback for
 *     devices that cannot keep 8 clusters of 8 CTAs resident
 * ---------------------------------------------------------------------------------------------- */
typedef struct UmgenDecodeArgs {
    /* ---- weights (packed once by the host, see umgen_b200/weights.py) ---- */
    int64_t n_layer;            /* OAR depth (36 for UMGen_Large) */
    const void* oar_h;          /* [n_layer][UMGEN_OAR_LAYER_H] fp16 */
    const void* oar_f;          /* [n_layer][UMGEN_OAR_LAYER_F] fp32 */
    const void* ln_oar_f;       /* [768] */
    const void* head_map_h;     /* head_ar_map   [8192][768] */
    const void* head_bbox_h;    /* head_ar_bbox3d[1028][768] */
    const void* head_img_h;     /* head_ar_img   [8192][768] */
    const void* map_table_f;    /* [8192][768] map_mlp_pre(map_codebook.weight): embedding of a map token (UMGen.py:1067-1068) */
    const void* img_table_f;    /* [8192][768] img_mlp_pre(img_codebook.weight) (UMGen.py:1135-1136) */
    const void* be_f;           /* transformer.be  [1028][768] */
    const void* axe_f;          /* transformer.axe [8][768] */
    const void* tske_f;         /* transformer.tske[task id] row, [768] */
    const void* fpe_f;          /* fouier_pe [1024][768] (bf16 values widened to fp32) */
    const void* box_lut_d;      /* [1028][10] float64: attribute token -> metres (tokenizer.py:679-687, normalize.py:189-229) */
    /* ---- per-frame inputs ---- */
    const void* tar_feat_f;        /* [2207][768] conditioning feature of the last frame (UMGen.py:1227-1231) */
    const void* tar_bbox_logits_f; /* [660][1028] head_tar_bbox3d(tar_feat[1032+i]) for bbox content i (UMGen.py:1087,1103); may be NULL if merge and control are off */
    const void* pose_tok_i32;      /* [3] pose tokens of the new frame (from the ego net or the control dict) */
    const void* prev_bbox_i32;     /* [660] bbox3d tokens of the last conditioning frame */
    const void* teacher_i32;       /* optional [2207] ids forced into the stream after each pick (parity tests, given prefix); NULL = free running */
    int64_t prefix_len;            /* positions 1..prefix_len (1-indexed, incl. bos/eos) are GIVEN (init_tokens of UMGen.py:1184-1201: pose, then map,
                                      then bbox3d): their ids come from teacher_i32, no head / sampling / rule check runs for them, and teacher_i32 is ignored
                                      beyond them.  0 = no given prefix (teacher_i32, if any, forces every position) */
    uint64_t control_mask;         /* bit s set = agent slot s is controlled (UMGen.py:1083-1089) */
    /* ---- sampling (UMGen.py:899-974) ---- */
    int64_t top_k_map, top_k_bbox, top_k_img; /* sample_method "topk": 1 = greedy; <= 16 */
    int64_t sample_topp;                       /* 1 = sample_method "topp" (UMGen.py:915-965), the top_k fields are then ignored */
    double top_p_map, top_p_bbox, top_p_img;   /* nucleus mass per modality (the reference passes topk_image = 16 as p for images, UMGen.py:1133) */
    double temperature;
    uint64_t seed;
    int64_t frame_index;
    int64_t merge_ar_tar;   /* config.merage_ar_tar */
    int64_t rule_constrain; /* config.rule_constrain */
    /* ---- state and scratch ---- */
    void* kv_h;        /* [n_layer][2][16][UMGEN_KV_ROWS][48] fp16 (kernel-private layout: the 8-cluster kernel keeps row r in owner r % 8's run of 16-key fragment tiles) */
    void* scratch_f;   /* >= umgen_decode_scratch_floats() fp32, zeroed by the call */
    /* ---- outputs ---- */
    void* out_tokens_i32;  /* [2207] ids of the frame (bos/eos positions hold the aux id) */
    void* picks_i32;       /* [2207] what the sampler itself chose at each position (== out unless teacher forced) */
    void* logits_dump_f;   /* optional [2207][8192] AR-head logits per position (row p-1), NULL to skip */
    void* status_i32;      /* [96]: [0] abort code (0 ok), [1] slots wiped by the rule check, [2] TAR-head resamples, [3] steps run; [8..59] debug cycle
                              probes; [60] kernel kilo-cycles (CTA 0); profiling builds (-DUMGEN_DECODE_PROFILE=1) also [61..63] ring / DSMEM / L2 wait,
                              [64] attention path, [65] head + sampling, all in kilo-cycles of CTA 0 / thread 0 */
    /* ---- execution ---- */
    int64_t n_steps;   /* number of decode steps to run (2206 = whole frame; fewer for tests) */
    int64_t mode;      /* 0 = 8-cluster kernel when oar_cl_h is given and the device can hold its 8 clusters, else the L2-exchange kernel;
                          1 = L2-exchange kernel; 2 = 8-cluster kernel (error if unavailable) */
    int64_t grid;      /* L2-exchange kernel only: CTAs to launch; 0 = one per SM */
    void* debug_u64;   /* L2-exchange kernel only: optional [grid][16] globaltimer stamps of one probed layer, NULL to skip */
    const void* oar_cl_h; /* [n_layer][UMGEN_OAR_LAYER_H] fp16: oar_h re-packed per CTA of the cluster kernel (umgen_pack_oar_cluster); may be NULL */
    /* ---- late conditioning rows (8-cluster kernel only; NULL = everything is ready at launch) ----
     * The decode of a frame needs the bbox3d rows of tar_feat (rows >= UMGEN_TAR_LATE_ROW0, from the box_tar pass, UMGEN.py:1497-1511) and
     * tar_bbox_logits_f only from step 1030 on, and the kernel occupies 64 of the 148 SMs: the host may launch it as soon as the other rows are
     * final, run the box_tar pass beside it on another stream, and then store tar_ready_value to *tar_ready_i32 with umgen_signal_ready. */
    const void* tar_ready_i32;
    int64_t tar_ready_value;
} UmgenDecodeArgs;

int64_t umgen_decode_scratch_floats(void);
int umgen_decode_frame(const UmgenDecodeArgs* args, void* stream);
/* Several scenes per launch (the reference's infer_oar_net is hard-wired to batch 1, UMGen.py:907,1093): args[0 .. n_scenes)
 * are decoded in lockstep by ONE launch of the 8-cluster kernel -- every scene at the same position and layer, two columns of each MMA's B
 * operand per scene, so the scenes share every weight fragment read from HBM and every exchange.  Per scene: tar_feat_f, tar_bbox_logits_f,
 * pose_tok_i32, prev_bbox_i32, teacher_i32, control_mask, seed, frame_index, kv_h, out_tokens_i32, picks_i32, logits_dump_f, status_i32,
 * tar_ready_*; everything else (weights, sampling set-up, prefix_len, n_steps, scratch_f) must be equal across scenes (checked).  The ids of a
 * scene are bit-identical to those of umgen_decode_frame on the same inputs.  n_scenes <= umgen_decode_max_scenes() (3 in this build: shared
 * memory per scene is what limits it); n_scenes == 1 is umgen_decode_frame. */
int umgen_decode_frames(const UmgenDecodeArgs* args, int64_t n_scenes, void* stream);
int umgen_decode_max_scenes(void);
/* how many 8-CTA clusters of the cluster decode kernel the current device can keep resident (8 are needed); no launch */
int umgen_decode_cluster_capacity(void);
/* oar_h [n_layer][UMGEN_OAR_LAYER_H] -> oar_cl_h (same size), the cluster kernel's layout.  CTA g = cluster * 8 + rank (cluster < 8, rank < 8)
 * owns heads 2 cluster + hh (hh < 2), the c_proj / mlp c_proj output rows [96 rank, +96) and the hidden units [48 g, +48).  Per layer and CTA
 * one contiguous run of 221 184 bytes, every 16x16 tile stored as a 512-byte mma.m16n8k16 A-fragment block: the half at position e of a
 * block is tile element (row, col) with lane = e / 8, reg = (e % 8) / 2, row = lane / 4 + 8 (reg & 1), col = 2 (lane % 4) + e % 2 + 8 (reg / 2).
 *   c_attn   [warp 12][k-step 4][tile 0 | tile 1 | 4-row tile]: local rows lr < 36 = (hh, {q,k,v}, e < 6) <-> c_attn row
 *            {q,k,v} * 768 + (2 cluster + hh) * 48 + 6 rank + e; columns 16 (4 warp + k-step) ..; the 4-row tile (rows 32..35) keeps only
 *            lanes 0..15 x {reg 0, reg 2} (128 bytes)
 *   c_proj   [tile 6][k-step 6]: rows 96 rank + 16 tile .., columns (2 cluster) * 48 + 16 k-step ..
 *   c_fc     [warp 12][k-step 4][tile 3]: rows 48 g + 16 tile .., columns 16 (4 warp + k-step) ..
 *   mlp c_proj [tile 48][k-step 3]: rows 16 tile .., columns 48 g + 16 k-step .. */
int umgen_pack_oar_cluster(const void* oar_h, void* oar_cl_h, int64_t n_layer, void* stream);

/* *flag_i32 = value with release semantics at device scope, stream-ordered after everything enqueued before it (see tar_ready_i32) */
int umgen_signal_ready(void* flag_i32, int64_t value, void* stream);

/* BoxOverlap.check_collision(boxes, fliter=True) (plugin/misc/misc.py:591-630) for n_cases box lists, computed by the same device functions the
 * decode kernels' rule path runs (rule_based_constraint, UMGen.py:1336-1345).  boxes_d: device float64 [total][10] (x,y,z,l,w,h,yaw,vx,vy,vz), the
 * boxes of case i are rows offsets_i32[i] .. offsets_i32[i+1]-1 (<= 64 per case); out_i32[i] = 1 if the last kept box collides, else 0. */
int umgen_check_collision(const void* boxes_d, const void* offsets_i32, int64_t n_cases, void* out_i32, void* stream);

/* head_tar_bbox3d over the 660 bbox content rows of tar_feat (UMGen.py:1087,1103):
 * out[i][v] = sum_c tar_feat[1032 + i][c] * w[v][c] */
int umgen_tar_bbox_logits(const void* tar_feat_f, const void* head_tar_bbox_h, void* out_f, void* stream);

/* ------------------------------------------------------------------------------------------------
 * TAR encoders (models/UMGen.py:634-872 driving models/module.py BlockTAR / Decoder).  The host
 * (umgen_b200/tar.py) sequences these per sub-block exactly as BlockTAR.forward_func does
 * (module.py:332-359): LayerNorm -> fused-QKV GEMM -> attention -> projection GEMM (+residual) ->
 * LayerNorm -> c_fc GEMM (GELU) -> c_proj GEMM (+residual).
 * ---------------------------------------------------------------------------------------------- */
#define UMGEN_EPI_BIAS_F16 0  /* out fp16 = acc + bias            (c_attn, q/k/v_attn: module.py:206,485-487) */
#define UMGEN_EPI_GELU_F16 1  /* out fp16 = gelu_erf(acc + bias)  (MLP c_fc: module.py:246-247) */
#define UMGEN_EPI_RESID_F32 2 /* out fp32 += acc + bias           (c_proj + residual: module.py:229,338,248) */
#define UMGEN_EPI_STORE_F32 3 /* out fp32 = acc + bias            (heads) */
#define UMGEN_EPI_RESID_F16 4 /* out fp16 = acc + bias + resid_h  (VQ decoder: conv + skip, vq_modules.py:127) */
#define UMGEN_EPI_NCHW_F32 5  /* umgen_conv3x3_nchw_f32 only: the first n_out columns + bias as fp32 planes [B][n_out][H][W] */

/* D[M,N] = epilogue(A[M,K] . W[N,K]^T): fp16 operands, fp32 accumulation on tcgen05 tensor cores fed by TMA.
 * N % 128 == 0, K % 64 == 0; lda/ldo = row pitches in elements; bias_f may be NULL. */
int umgen_gemm_f16(const void* a_h, int64_t lda, const void* w_h, const void* bias_f, void* out, int64_t ldo, int64_t M, int64_t N,
                   int64_t K, int epilogue, void* stream);
/* same with an fp16 residual operand (UMGEN_EPI_RESID_F16) and N % 128 == 0 allowed */
int umgen_gemm_f16_ex(const void* a_h, int64_t lda, const void* w_h, const void* bias_f, void* out, int64_t ldo, const void* resid_h,
                      int64_t ldr, int64_t M, int64_t N, int64_t K, int epilogue, void* stream);
/* LayerNorm(weight only, eps 1e-5) over rows of 768 (module.py:26-37); out fp16 if out_half else fp32 */
int umgen_layernorm(const void* x_f, const void* w_f, void* out, int64_t rows, int out_half, void* stream);
int umgen_cast_f16(const void* x_f, void* out_h, int64_t n, void* stream);
/* out[i] = table[tok[i]] (+ grid_pos[i % 1024]) : map token feature (UMGen.py:448-458); table = GMLP(codebook) */
int umgen_map_feature(const void* tok_i32, const void* table_f, const void* grid_pos_f, void* out_f, int64_t n_tok, void* stream);
/* affine_transform (UMGen.py:310-354): bilinear warp of [T,1024,768] by the decoded pose tokens [T,3] */
int umgen_map_warp(const void* feat_f, const void* pose_tok_i32, const void* pose_lut_f, void* out_f, int64_t T, void* stream);

typedef struct UmgenEmbedArgs {
    const void *pose_i32, *map_i32, *bbox_i32, *image_i32;   /* [T,3] [T,1024] [T,660] [T,512] */
    const void *fpe_f, *img_table_f, *be_f, *axe_f, *spe_f, *tpe_f, *spatial_f;
    const void *map_feat_f, *map_warped_f;                   /* [T,1024,768]; warped may be NULL */
    void* out_f;                                             /* [T, S, 768] with S = 1031 / 1693 / 2207 */
    int64_t T, n_mods;                                       /* n_mods: 2 pose+map, 3 +bbox3d, 4 +image */
    int64_t t_offset;                                        /* index in the window of the first of the T frames given (selects tpe rows t_offset ..) */
} UmgenEmbedArgs;
/* get_mod_emb_pre + add_bos_eos + add_pos_emb (UMGen.py:438-515) for one TAR pass */
int umgen_embed_sequence(const UmgenEmbedArgs* args, void* stream);

/* attention over <= 32 tokens per group, 16 heads x 48: token i of group g is row g*group_stride + i*tok_stride of
 * the fused qkv activation [rows][2304]; temporal attention (causal) and the 3-token ego self attention */
int umgen_small_attention(const void* qkv_h, void* y_h, int64_t n_groups, int64_t n_tok, int64_t group_stride, int64_t tok_stride,
                          int causal, void* stream);
/* the same for queries q0 .. n_tok-1 only (rows 0 .. q0-1 of qkv serve as keys / values; their y rows are not written): the last frame of a
 * window against the cached keys / values of the frames before it */
int umgen_small_attention_from(const void* qkv_h, void* y_h, int64_t n_groups, int64_t n_tok, int64_t group_stride, int64_t tok_stride,
                               int causal, int64_t q0, void* stream);
/* non-causal attention inside each of T frames of S tokens (module.py:336-338) on the fused qkv activation [T*S][2304] -> y [T*S][768].
 * tcgen05 kernel (csrc/attn_sm100.cu): Q K^T and P V as tcgen05.mma with Q, P (operand A) and S, O (accumulators) in tensor memory, K / V tiles
 * by TMA; dbg_f: NULL, or >= 128*128 + 128 + 128*48 floats that receive CTA (0,0,0)'s first score tile, row sums and unnormalised output */
int umgen_spatial_attention_tc(const void* qkv_h, void* y_h, int64_t T, int64_t S, void* dbg_f, void* stream);
/* the same contraction on mma.sync tensor-core instructions (csrc/tar.cu; the round-1 kernel, kept as the cross-check of the tcgen05 one) */
int umgen_spatial_attention(const void* qkv_h, void* y_h, int64_t T, int64_t S, void* stream);
/* FlashCrossAttention core (module.py:494-506): nq query rows against n_k key/value rows, all [*,768] fp16 */
int umgen_cross_attention(const void* q_h, const void* k_h, const void* v_h, void* y_h, int64_t nq, int64_t n_k, void* stream);
/* token_sampler on `rows` logit rows of width V (ego head, UMGen.py:1001-1004): top_p <= 0 -> topk + sfmx_temp_sampling (UMGen.py:899-913,
 * 967-974) with top_k; top_p > 0 -> sample_top_p (UMGen.py:915-965) with nucleus mass top_p (top_k ignored) */
int umgen_sample_rows(const void* logits_f, int64_t rows, int64_t V, int64_t top_k, double top_p, double temperature, uint64_t seed,
                      int64_t frame_index, void* out_i32, void* stream);
/* tar_emb assembly of _inference step 2 (UMGen.py:1496-1511) for the last frame: rows [row0, row1) of out [2207,768]
 * (rows 5..1030 from the map pass, UMGEN_TAR_LATE_ROW0..1692 from the box pass, the rest from the full pass) */
int umgen_assemble_tar_feat(const void* f_all, const void* f_map, const void* f_box, const void* warped_last, void* out, int64_t row0, int64_t row1,
                            void* stream);
/* CTAs the persistent GEMM may launch (0 = one per SM); lowered by the host while the decode kernel holds 64 SMs beside a TAR pass */
int umgen_gemm_set_sm_limit(int n);

/* ------------------------------------------------------------------------------------------------
 * VQ pixel decoders (tokenizer/vq_model.py:87-101, tokenizer/vq_modules.py:293-415, tools/decode_map.py:25-30).
 * Channels-last fp16 activations; 3x3 convolutions over >= 64 channels = umgen_conv3x3_f16 (implicit GEMM), the 16-channel
 * ones = umgen_im2col3x3 + umgen_gemm_f16_ex, 1x1 convolutions = umgen_gemm_f16_ex on the activation itself.  The host
 * (umgen_b200/vq.py) sequences them as Decoder.forward does.
 * ---------------------------------------------------------------------------------------------- */
/* nn.Conv2d(Cin, Cout, 3, stride 1, padding 1) (vq_modules.py:63-66 Upsample.conv, :98-107 ResnetBlock.conv1/conv2, :322-326 conv_in) as an
 * implicit GEMM on the tcgen05 kernel: x_h [B,H,W,Cin] fp16 channels-last, w_h [Cout][9*Cin] fp16 in (ky,kx,c) order, out_h [B*H*W][Cout] fp16.
 * The nine taps are fetched by 4-D TMA boxes straight from x_h (zero padding = the TMA unit's out-of-bounds fill); no im2col matrix.
 * epilogue: UMGEN_EPI_BIAS_F16, or UMGEN_EPI_RESID_F16 with resid_h [B*H*W][Cout] (the block's skip connection, vq_modules.py:127).
 * Needs Cin % 64 == 0, Cout % 128 == 0 and an image that tiles into boxes of 128 pixels (W % 128 == 0, or 128 % W == 0 and H % (128/W) == 0). */
int umgen_conv3x3_f16(const void* x_h, int64_t B, int64_t H, int64_t W, int64_t Cin, const void* w_h, const void* bias_f, void* out_h,
                      const void* resid_h, int64_t Cout, int epilogue, void* stream);
/* conv_out (vq_modules.py:330-334, 413-414): the same implicit GEMM for n_out <= 32 output channels; w_h [128][9*Cin] fp16 and bias_f [128] with
 * the rows >= n_out zero; out_f fp32 [B][n_out][H][W] (the layout Decoder.forward returns). */
int umgen_conv3x3_nchw_f32(const void* x_h, int64_t B, int64_t H, int64_t W, int64_t Cin, const void* w_h, const void* bias_f, void* out_f,
                           int64_t n_out, void* stream);
/* F.interpolate(scale_factor=2, mode="nearest") (vq_modules.py:34-40) over [B,H,W,C] fp16 -> [B,2H,2W,C] */
int umgen_upsample2x_nhwc(const void* in_h, void* out_h, int64_t B, int64_t H, int64_t W, int64_t C, void* stream);
/* out[i, 0:16] = table[idx[i]] : codebook lookup (quantize.py:341-342) */
int umgen_vq_gather(const void* idx_i32, const void* table_f, void* out_h, int64_t n, void* stream);
/* A[(b,y,x), (ky,kx,c)] for a 3x3/s1/p1 conv over [B,H>>up,W>>up,Cin]; upsample=1 folds the nearest 2x Upsample (vq_modules.py:34-40) */
int umgen_im2col3x3(const void* in_h, void* a_h, int64_t B, int64_t H, int64_t W, int64_t Cin, int64_t k_pad, int upsample, void* stream);
/* GroupNorm(32, eps 1e-6, affine) (+ swish) over [B,HW,C] fp16 (vq_modules.py:14-22); stats_f scratch [B*32*2] */
int umgen_groupnorm_nhwc(const void* x_h, const void* gamma_f, const void* beta_f, void* y_h, void* stats_f, int64_t B, int64_t HW, int64_t C,
                         int swish, void* stream);
/* The same GroupNorm with a coalesced statistics pass (pixel slabs, per-slab partial sums added up in slab order by the last CTA of an image:
 * deterministic).  C in {128, 256, 512}.  scratch_f: umgen_groupnorm_scratch_floats(B, HW) floats; the caller zeroes its last B words once. */
int64_t umgen_groupnorm_scratch_floats(int64_t B, int64_t HW);
int umgen_groupnorm_nhwc_slab(const void* x_h, const void* gamma_f, const void* beta_f, void* y_h, void* scratch_f, int64_t B, int64_t HW, int64_t C,
                              int swish, void* stream);
/* p = softmax(scale * s) over rows of length n (AttnBlock, vq_modules.py:160-163) */
int umgen_softmax_rows(const void* s_f, void* p_h, int64_t rows, int64_t n, double scale, void* stream);
int umgen_transpose_f16(const void* in_h, void* out_h, int64_t rows, int64_t cols, void* stream);
/* conv_out (vq_modules.py:384-387): 3x3 conv to <= 8 channels, NHWC fp16 in, NCHW fp32 out; w_f [Cout][9][Cin] */
int umgen_conv_out3x3(const void* in_h, const void* w_f, const void* bias_f, void* out_f, int64_t B, int64_t H, int64_t W, int64_t Cin,
                      int64_t Cout, void* stream);
/* to_rgb (tools/decode_map.py:25-30): 1x1 projection to 3 channels + min-max to [-1,1] over the whole chunk */
int umgen_to_rgb(const void* x_f, const void* w_f, void* out_f, void* minmax_u32, int64_t B, int64_t Cin, int64_t HW, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UMGEN_H_ */
