// fp32-mode attention and the tensor-map plan shared with the tcgen05 kernel (attention_tc.cu).
//
// attn_f32_kernel: the exact-arithmetic parity mode of the MultiHeadAttention core (reference det.py:364-369):
// one thread per query, IEEE f32 products and sums, q scaled by 1/sqrt(d) after its bias as Keras does, K/V tiles
// staged in shared memory and read as warp-wide broadcasts, online softmax so that the (B, H, T, T) score tensor
// is never materialised.
#include "common.cuh"
#include "kernels.h"

namespace vitdet {

namespace {

// ------------------------------------------------------------------------------------------------
// f32 parity kernel: thread = query, 128 queries per CTA, 32-key tiles in shared memory.
// ------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(128)
attn_f32_kernel(const float* __restrict__ qkv, int ldq, float* __restrict__ ctx, int ldo, int T, int H, int hp,
                float scale) {
    constexpr int KT = 32;
    __shared__ __align__(16) float sK[KT][D];
    __shared__ __align__(16) float sV[KT][D];
    const int bh = blockIdx.y;
    const int b = bh / H, h = bh - b * H;
    const int q = blockIdx.x * 128 + threadIdx.x;
    const bool q_ok = q < T;
    const size_t row_base = static_cast<size_t>(b) * T;

    float qv[D], acc[D];
    {
        const float* qp = qkv + (row_base + (q_ok ? q : 0)) * ldq + h * hp;
#pragma unroll
        for (int i = 0; i < D; ++i) { qv[i] = qp[i] * scale; acc[i] = 0.f; }   // Keras scales q (after bias)
    }
    float m = -INFINITY, l = 0.f;
    for (int k0 = 0; k0 < T; k0 += KT) {
        __syncthreads();
        for (int e = threadIdx.x; e < KT * D; e += 128) {
            const int r = e / D, c = e - r * D;
            const int key = k0 + r;
            float kv = 0.f, vv = 0.f;
            if (key < T) {
                const float* rp = qkv + (row_base + key) * ldq;
                kv = rp[(H + h) * hp + c];
                vv = rp[(2 * H + h) * hp + c];
            }
            sK[r][c] = kv;
            sV[r][c] = vv;
        }
        __syncthreads();
        const int kn = (T - k0) < KT ? (T - k0) : KT;
        for (int r = 0; r < kn; ++r) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < D; ++i) s = fmaf(qv[i], sK[r][i], s);
            if (s > m) {
                const float c = expf(m - s);
                l *= c;
#pragma unroll
                for (int i = 0; i < D; ++i) acc[i] *= c;
                m = s;
            }
            const float pw = expf(s - m);
            l += pw;
#pragma unroll
            for (int i = 0; i < D; ++i) acc[i] = fmaf(pw, sV[r][i], acc[i]);
        }
    }
    if (q_ok) {
        float* op = ctx + (row_base + q) * ldo + h * hp;
        const float inv = 1.f / l;
#pragma unroll
        for (int i = 0; i < D; ++i) op[i] = acc[i] * inv;
        for (int i = D; i < hp; ++i) op[i] = 0.f;
    }
}

}  // namespace

// Tensor map over the fused qkv matrix [B*T, 3*H*hp] (bf16): boxes of 64 rows x 64 columns, 128B swizzle, placed at
// column head * hp (+ 64 for the second box of a head wider than 64); hp = key_dim rounded up to 8.  A box also carries the first columns of the next head
// (zeros past the end of the matrix); the kernel neutralises them (attention_tc.cu).
int attn_bf16_make_plan(AttnPlan* plan, const AttnDesc& d) {
    if (d.d <= 0 || d.hp < d.d || d.hp > 128 || (d.hp % 8)) return -20;      // one or two 64-column boxes per head (key_dim <= 128)
    if ((d.ldq % 8) || (d.ldo % 8) || d.ldq < 3 * d.H * d.hp || d.ldo < d.H * d.hp) return -21;
    if ((reinterpret_cast<uintptr_t>(d.qkv) & 15) || (reinterpret_cast<uintptr_t>(d.ctx) & 15)) return -22;
    plan->desc = d;
    return make_tmap_bf16_2d(&plan->tmQKV, d.qkv, d.B * d.T, 3 * d.H * d.hp, d.ldq, 64);
}

cudaError_t attn_f32_launch(const AttnDesc& d, cudaStream_t stream) {
    const float* qkv = static_cast<const float*>(d.qkv);
    float* ctx = static_cast<float*>(d.ctx);
    dim3 grid((d.T + 127) / 128, d.B * d.H);
#define VITDET_ATTN_F32(DD)                                                                              \
    case DD:                                                                                             \
        attn_f32_kernel<DD><<<grid, 128, 0, stream>>>(qkv, d.ldq, ctx, d.ldo, d.T, d.H, d.hp, d.scale);  \
        break;
    switch ((d.d + 7) / 8 * 8) {   // pad columns of every head slot are zero in qkv
        VITDET_ATTN_F32(8) VITDET_ATTN_F32(16) VITDET_ATTN_F32(24) VITDET_ATTN_F32(32) VITDET_ATTN_F32(40)
        VITDET_ATTN_F32(48) VITDET_ATTN_F32(56) VITDET_ATTN_F32(64)
        // wider heads (the query and the accumulator no longer fit the register file: slow, but the exact form exists)
        VITDET_ATTN_F32(72) VITDET_ATTN_F32(80) VITDET_ATTN_F32(88) VITDET_ATTN_F32(96) VITDET_ATTN_F32(104) VITDET_ATTN_F32(112)
        VITDET_ATTN_F32(120) VITDET_ATTN_F32(128)
        default: return cudaErrorInvalidValue;   // key_dim <= 128
    }
#undef VITDET_ATTN_F32
    return cudaGetLastError();
}

}  // namespace vitdet
