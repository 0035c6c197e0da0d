// Internal (C++) declarations of the kernel launchers.  The public boundary is include/vitdet_b200.h.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace vitdet {

// ------------------------------------------------------------------------------------------------
// Dense layer:  out[M, ldc] = epilogue( A[M,K] (lda) * W[N,K]^T (ldw) )
//   epilogue(acc)[m,n] = act(acc + bias[n] + pos[m % pos_period]) + resid[m,n]
// (the order matches the reference graph: Dense -> activation -> keras.layers.add, det.py:388-412,
//  and Dense -> add(position) for the patch embedding, det.py:297-307).
// Columns [N, round_up(N, vec)) of `out` are written as zero, vec = 8 (bf16) or 4 (f32).
// ------------------------------------------------------------------------------------------------
struct GemmDesc {
    const void* A = nullptr;   // bf16 (tensor-core path) or f32 (SIMT path), row-major, K contiguous
    int lda = 0;
    const void* W = nullptr;   // same element type as A; row n = output unit n, K contiguous
    int ldw = 0;
    int M = 0, N = 0, K = 0;
    const float* bias = nullptr;
    const float* pos = nullptr;
    int pos_period = 1;
    const float* resid = nullptr;   // f32, only with out_f32; may alias `out`
    int ldr = 0;
    void* out = nullptr;
    int ldc = 0;
    int out_f32 = 0;
    int act = 0;               // vitdet::Act
    int block_n = 0;           // 0 = choose
    // Optional fused LayerNormalization of the OUTPUT row (tensor-core kernel, f32 output, N <= 32 only: the
    // epilogue thread holds the whole row): ln_out[m, :] = LN(out[m, :]) * gamma + beta as bf16, pads zero.
    const float* ln_gamma = nullptr;
    const float* ln_beta = nullptr;
    float ln_eps = 1e-3f;
    void* ln_out = nullptr;    // bf16 [M, ln_ld]
    int ln_ld = 0;
    // fp32-accumulate mode on the tensor cores: every float32 value travels as TWO bf16 planes, hi = bf16(x) and
    // lo = bf16(x - hi) (|x - hi - lo| <= 2^-18 |x|), and the product is A_hi W_hi + A_hi W_lo + A_lo W_hi — three passes
    // over K through the same pipeline (the lo*lo term, 2^-18 of the result, is dropped).  split = 1: A_lo / W_lo are the
    // lo planes (same shapes and pitches as A / W).  out_split = 1 (bf16 output only): the result is stored as planes too,
    // out = hi, out_lo = lo.  precise = 1: exact expf / tanhf based activations instead of the ex2/rcp.approx forms.
    int split = 0;
    const void* A_lo = nullptr;
    const void* W_lo = nullptr;
    int out_split = 0;
    void* out_lo = nullptr;
    int precise = 0;
};

struct TcGemmPlan {
    CUtensorMap tmA, tmB;
    CUtensorMap tmC;           // bf16 outputs only: 32 x 32 boxes, 64B swizzle (TMA store from the epilogue)
    CUtensorMap tmA_lo, tmB_lo, tmC_lo;      // the lo planes of the split (fp32-accumulate) mode
    GemmDesc desc;
    int block_n = 0;
    int num_stages = 0;
    int n_tiles = 0;
    int num_tiles = 0;
    int grid = 0;
    size_t smem_bytes = 0;
    int pair = 0;              // 1: built for the CTA-pair kernel
};

int choose_block_n(int N);
int make_tmap_bf16_2d(CUtensorMap* map, const void* base, int rows, int cols, int ld, int box_rows);
int make_tmap_bf16_2d_ex(CUtensorMap* map, const void* base, int rows, int cols, int ld, int box_cols, int box_rows, int swizzle_bytes);
int tc_gemm_make_plan(TcGemmPlan* plan, const GemmDesc& d, int num_sms);
cudaError_t tc_gemm_launch(const TcGemmPlan& plan, cudaStream_t stream);
// CTA-pair (cta_group::2) variant for the large layers (gemm_tc2.cu): 256 x BN tile per 2-CTA cluster.
int tc2_gemm_make_plan(TcGemmPlan* plan, const GemmDesc& d, int num_sms);
cudaError_t tc2_gemm_launch(const TcGemmPlan& plan, cudaStream_t stream);

// ------------------------------------------------------------------------------------------------
// Fused tail of the encoder MLP pyramid (mlp_tail.cu): x += act(W2 act(W1 act(W0 A + b0) + b1) + b2) in place on the
// f32 residual stream x, optionally followed by LayerNorm -> bf16 (the next block's first LayerNorm).
// ------------------------------------------------------------------------------------------------
struct MlpTailDesc {
    const void* A = nullptr; int lda = 0;      // bf16 [M, K[0]]
    int M = 0;
    int N[3] = {0, 0, 0}, K[3] = {0, 0, 0};
    const void* W[3] = {nullptr, nullptr, nullptr}; int ldw[3] = {0, 0, 0};     // bf16 [N, K], K contiguous
    const float* bias[3] = {nullptr, nullptr, nullptr};
    float* x = nullptr; int ldx = 0;           // f32 [M, ldx]: residual in, result out
    const float* ln_gamma = nullptr; const float* ln_beta = nullptr; float ln_eps = 1e-3f;
    void* ln_out = nullptr; int ln_ld = 0;     // bf16 [M, ln_ld] or null
    int act = 0;
};
struct MlpTailPlan {
    CUtensorMap tmA, tmW[3];
    MlpTailDesc desc;
    int npad[3], kb[3];
    uint32_t off_w[3], off_act[2];
    size_t smem_bytes = 0;
    int num_tiles = 0, grid = 0;
    bool valid = false;
};
bool mlp_tail_supported(const int N[3], const int K[3]);
int mlp_tail_make_plan(MlpTailPlan* plan, const MlpTailDesc& d, int num_sms);
cudaError_t mlp_tail_launch(const MlpTailPlan& plan, cudaStream_t stream);

// f32 CUDA-core GEMM with the same epilogue (the fp32 parity mode); A, W, out are f32.
// Requires lda, ldw multiples of 4 and zero padding of A and W in columns [K, round_up(K,4)).
cudaError_t simt_gemm_launch(const GemmDesc& d, cudaStream_t stream);

// ------------------------------------------------------------------------------------------------
// Attention.  qkv: [B*T, ldq] with q of head h at columns [h*hp, h*hp+d), k at [(H+h)*hp, ...),
// v at [(2H+h)*hp, ...); ctx: [B*T, H*hp] (columns d..hp-1 of every head written as zero).
// ------------------------------------------------------------------------------------------------
struct AttnDesc {
    const void* qkv = nullptr;
    int ldq = 0;
    void* ctx = nullptr;
    int ldo = 0;
    int B = 0, T = 0, H = 0, d = 0, hp = 0;
    float scale = 1.f;          // 1/sqrt(key_dim), applied to q·k (Keras scales q after its bias)
};
struct AttnPlan {
    CUtensorMap tmQKV;
    AttnDesc desc;
};
int attn_bf16_make_plan(AttnPlan* plan, const AttnDesc& d);
cudaError_t attn_tc_launch(const AttnPlan& plan, cudaStream_t stream);     // tcgen05 kernel, one softmax thread per query row (attention_tc.cu)
#ifdef VITDET_EXPERIMENTS      // experiments/attention/: measured and dropped (profiles/r02_attention_analysis.md)
cudaError_t attn_tc8_launch(const AttnPlan& plan, cudaStream_t stream);
cudaError_t attn_tc3_launch(const AttnPlan& plan, cudaStream_t stream);
cudaError_t attn_tcp_launch(const AttnPlan& plan, int num_sms, cudaStream_t stream);
cudaError_t attn_tc8p_launch(const AttnPlan& plan, int num_sms, cudaStream_t stream);
cudaError_t attn_pp_launch(const AttnPlan& plan, int num_sms, cudaStream_t stream);
cudaError_t attn_sw_launch(const AttnPlan& plan, int num_sms, cudaStream_t stream);
#endif
cudaError_t attn_f32_launch(const AttnDesc& d, cudaStream_t stream);
// fp32-accumulate mode on the tensor cores (attention_tcs.cu): plan.desc.qkv / ctx are the hi planes, tm_lo the tensor map
// of the lo plane of qkv (same shape and pitch), ctx_lo the lo plane of the context rows.
cudaError_t attn_tcs_launch(const AttnPlan& plan, const CUtensorMap& tm_lo, void* ctx_lo, cudaStream_t stream);

// ------------------------------------------------------------------------------------------------
// Memory-bound row kernels
// ------------------------------------------------------------------------------------------------
// tf.image.extract_patches(sizes=strides=p, padding='SAME') + Reshape (det.py:195-197, 279-280).
// images f32 NHWC [B,H,W,3] -> patches [B*gh*gw, ldp]: element r*rp + c*3 + ch (rp >= 3p = run pitch of one
// patch row; rp = 3p is the reference's dense vector), zero padded.
cudaError_t patchify_launch(const void* images, int in_u8 /*1: uint8 pixels, normalised x/127.5-1 on the fly*/, int B, int H, int W, int p,
                            void* patches, int ldp, int rp, int out_f32, cudaStream_t stream);

// keras LayerNormalization(axis=-1), eps=1e-3, biased variance (det.py:353-357, 375-379).
cudaError_t layernorm_launch(const float* x, int ldx, const float* gamma, const float* beta, int M, int D,
                             float eps, void* y, int ldy, int out_f32, cudaStream_t stream);

// mlp_head first stage (det.py:454-463): Dense(D -> S) on every token of M / tokens images, stored as the slot matrix
// [images*S, ldo] (ldo >= tokens; ldo == tokens is the compact [B, T*S] buffer whose row-major view IS the reference's
// Reshape((S,-1))).
cudaError_t head_slots_launch(const float* x, int ldx, const float* w /*[S, D] f32*/, const float* bias,
                              int M, int D, int S, int tokens, int ldo, void* out, int out_f32, cudaStream_t stream);

// Final Dense(U -> 6) 'MLP_Head_no_Sigmoid' (det.py:489-493) fused with transform_predictions
// (det.py:586-647) and the 0.5/0.5 thresholds (det.py:2257-2283 / 1359-1384).
struct DecodeParams {
    float obj_thr = 0.5f, cls_thr = 0.5f;
    int strict = 1;             // 1: keep iff score >  thr (metric copy); 0: keep iff score >= thr (visualise copy)
    float img_h = 608.f, img_w = 608.f;
    int classes = 80;
    int apply_transform = 1;    // 0: rows are already transform_predictions output (det.py:1340-1341)
    float corner_scale = 1.f;   // enlarged_image_scale (det.py:2294-2297)
};
struct DecodeOut {
    float* logits = nullptr;    // [R, 6] raw (optional when decoding given logits)
    float* decoded = nullptr;   // [R, 6] conf, class in [0, classes-1], cx, cy, h, w (pixels)
    int32_t* class_id = nullptr;  // [R] round-half-even(class)
    float* class_conf = nullptr;  // [R] (0.5 - |class - id|) / 0.5
    uint8_t* keep = nullptr;    // [R]
    int32_t* corners = nullptr; // [R, 4] x0, y0, x1, y1 (int truncation then clip; det.py:2300-2325), optional
    float* packed = nullptr;    // [R, 13] decoded[6] | class id | class confidence | keep | corners[4], all as float32 (the gather record)
};
cudaError_t head_tail_launch(const void* h, int ldh, int in_f32, const float* w /*[6, U] f32*/, const float* bias,
                             int R, int U, const DecodeParams& dp, const DecodeOut& out, cudaStream_t stream);
cudaError_t decode_launch(const float* logits, int R, const DecodeParams& dp, const DecodeOut& out,
                          cudaStream_t stream);
// Input side (vision_transformer_utilities.py:435-447): uint8 [h, w, 3] -> resize_with_pad(th, tw) -> clip -> /127.5 - 1, f32 [th, tw, 3].
void resize_with_pad_geometry(int h, int w, int th, int tw, int* rh, int* rw, int* ph, int* pw);
cudaError_t preprocess_launch(const uint8_t* image, int h, int w, float* out, int th, int tw, cudaStream_t stream);
// iou_calculator (det.py:761-875): element-wise IoU of (cx, cy, h, w) boxes in the last four entries of `width`-wide rows.
cudaError_t iou_launch(const float* label, const float* pred, long long R, int width, float eps, float* iou,
                       cudaStream_t stream);

// Records the message vitdet_last_error() returns on the calling thread and hands `code` back (engine.cu).
int fail(int code, const char* fmt, ...);

}  // namespace vitdet
