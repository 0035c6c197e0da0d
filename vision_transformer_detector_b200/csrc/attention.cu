// Multi-head self-attention core: softmax(q k^T / sqrt(d)) v per (image, head), without ever
// materialising the (B, H, T, T) score tensor that keras.layers.MultiHeadAttention builds
// (reference det.py:364-369; 53.7 MB per image in f32 at T = 1296).
//
// bf16 kernel: flash-style.  One CTA = 128 queries of one (image, head): 8 consumer warps x 16
// query rows + 1 producer warp.  The producer streams 64-key K and V tiles with TMA
// (cp.async.bulk.tensor, 128B swizzle) through a 3-stage mbarrier ring; consumers read them with
// ldmatrix (conflict-free thanks to the swizzle), do QK^T and PV on the tensor cores
// (mma.sync m16n8k16 bf16, f32 accumulate) and keep the online-softmax state (running max, running
// sum, un-normalised output) in registers.  exp is evaluated as ex2 with log2(e)/sqrt(d) folded
// into one FFMA per score.
//
// f32 kernel: the fp32 parity mode; one thread per query, exact f32 arithmetic, K/V tiles staged
// in shared memory and read as warp-wide broadcasts.
#include "common.cuh"
#include "kernels.h"

namespace vitdet {

namespace {

constexpr int kBQ = 128;         // queries per CTA
constexpr int kBKV = 64;         // keys per pipeline stage
constexpr int kHP = 64;          // head pitch in elements (one 128-byte swizzle row of bf16)
constexpr int kStages = 3;
constexpr int kConsumerWarps = 8;
constexpr int kAttnThreads = (kConsumerWarps + 1) * 32;
constexpr int kTileBytes = kBKV * kHP * 2;     // 8 KiB: one K or V tile, also half of the Q tile

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Byte offset of 16-byte chunk `chunk` of row `row` inside a 128B-swizzled tile (rows of 128 B,
// tile base 1024-aligned): the hardware XORs address bits [4,7) with bits [7,10).
__device__ __forceinline__ uint32_t swz(int row, int chunk) {
    return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4));
}

struct AttnArgs {
    __nv_bfloat16* ctx;
    int ldo;
    int T, H;
    float scale_log2;    // log2(e) / sqrt(key_dim)
};

// NK16 = ceil(d/16) k-steps of QK^T, ND8 = ceil(d/8) n-tiles of PV.
template <int NK16, int ND8>
__global__ void __launch_bounds__(kAttnThreads, 2)
attn_bf16_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnArgs p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * kStages + 1];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * kBQ;
    const int bh = blockIdx.y;
    const int b = bh / p.H, h = bh - b * p.H;
    const int row_base = b * p.T;          // first token row of this image in the [B*T, ld] matrices
    const int nkv = (p.T + kBKV - 1) / kBKV;

    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sQ = base;                          // 2 x 8 KiB (rows 0-63, 64-127)
    const uint32_t sKV = base + 2 * kTileBytes;        // stage s: K at +2*s*tile, V right after
    const uint32_t bar_full = smem_u32(&bars[0]);
    const uint32_t bar_empty = smem_u32(&bars[kStages]);
    const uint32_t bar_q = smem_u32(&bars[2 * kStages]);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, kConsumerWarps);
        }
        mbar_init(bar_q, 1);
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == kConsumerWarps) {
        // ------------------------------ TMA producer ------------------------------
        if (lane == 0) {
            tma_prefetch_desc(&tmQKV);
            mbar_arrive_expect_tx(bar_q, 2 * kTileBytes);
            tma_load_2d(sQ, &tmQKV, bar_q, h * kHP, row_base + q0);
            tma_load_2d(sQ + kTileBytes, &tmQKV, bar_q, h * kHP, row_base + q0 + 64);
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < nkv; ++i) {
                mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
                mbar_arrive_expect_tx(bar_full + 8 * stage, 2 * kTileBytes);
                const uint32_t dst = sKV + stage * 2 * kTileBytes;
                tma_load_2d(dst, &tmQKV, bar_full + 8 * stage, (p.H + h) * kHP, row_base + i * kBKV);
                tma_load_2d(dst + kTileBytes, &tmQKV, bar_full + 8 * stage, (2 * p.H + h) * kHP, row_base + i * kBKV);
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
        }
        return;
    }

    // ---------------------------------- consumers ----------------------------------
    const int g = lane >> 2, t4 = lane & 3;

    // Q fragments (A operand, 16 rows x 16 d per k-step), loaded once.
    uint32_t qf[NK16][4];
    mbar_wait(bar_q, 0);
    {
        const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;       // row within the 128-row Q tile
        const uint32_t qb = sQ + (r >> 6) * kTileBytes;
#pragma unroll
        for (int kk = 0; kk < NK16; ++kk)
            ldsm_x4(qb + swz(r & 63, 2 * kk + (lane >> 4)), qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
    }

    float oacc[ND8][4];
#pragma unroll
    for (int i = 0; i < ND8; ++i) { oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY};     // rows g and g+8
    float l_run[2] = {0.f, 0.f};

    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < nkv; ++i) {
        mbar_wait(bar_full + 8 * stage, phase);
        const uint32_t sK = sKV + stage * 2 * kTileBytes;
        const uint32_t sV = sK + kTileBytes;

        // ---- S = Q K^T : 16 x 64 per warp ----
        float sacc[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sacc[j][0] = sacc[j][1] = sacc[j][2] = sacc[j][3] = 0.f;
            const int krow = 8 * j + (lane & 7);
#pragma unroll
            for (int kk = 0; kk < NK16; kk += 2) {
                // four 8x8 matrices: chunks 2kk .. 2kk+3 of key rows 8j..8j+7 (two k-steps)
                uint32_t b0, b1, b2, b3;
                ldsm_x4(sK + swz(krow, 2 * kk + (lane >> 3)), b0, b1, b2, b3);
                mma_bf16_16816(sacc[j], qf[kk], b0, b1);
                if (kk + 1 < NK16) mma_bf16_16816(sacc[j], qf[kk + 1], b2, b3);
            }
        }

        // ---- online softmax ----
        const int kv0 = i * kBKV;
        const bool tail = kv0 + kBKV > p.T;
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float s = sacc[j][e] * p.scale_log2;
                if (tail && (kv0 + 8 * j + 2 * t4 + (e & 1)) >= p.T) s = -INFINITY;
                sacc[j][e] = s;
                mx[e >> 1] = fmaxf(mx[e >> 1], s);
            }
        }
        float corr[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float m_new = fmaxf(m_run[r], mx[r]);
            corr[r] = fast_exp2(m_run[r] - m_new);
            m_run[r] = m_new;
            l_run[r] *= corr[r];
        }
#pragma unroll
        for (int dn = 0; dn < ND8; ++dn) {
            oacc[dn][0] *= corr[0]; oacc[dn][1] *= corr[0];
            oacc[dn][2] *= corr[1]; oacc[dn][3] *= corr[1];
        }
        uint32_t pf[4][4];     // P as A fragments: 4 k-steps of 16 keys
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float p0 = fast_exp2(sacc[j][0] - m_run[0]);
            const float p1 = fast_exp2(sacc[j][1] - m_run[0]);
            const float p2 = fast_exp2(sacc[j][2] - m_run[1]);
            const float p3 = fast_exp2(sacc[j][3] - m_run[1]);
            l_run[0] += p0 + p1;
            l_run[1] += p2 + p3;
            pf[j >> 1][(j & 1) * 2 + 0] = pack_bf16x2(p0, p1);
            pf[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(p2, p3);
        }

        // ---- O += P V : 16 x d per warp ----
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const int vrow = 16 * kk + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
            for (int dn = 0; dn < ND8; dn += 2) {
                uint32_t b0, b1, b2, b3;
                ldsm_x4_t(sV + swz(vrow, dn + (lane >> 4)), b0, b1, b2, b3);
                mma_bf16_16816(oacc[dn], pf[kk], b0, b1);
                if (dn + 1 < ND8) mma_bf16_16816(oacc[dn + 1], pf[kk], b2, b3);
            }
        }

        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * stage);
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
    }

    // ---- finalise: O / l, stage the warp's 16 x 64 bf16 block in (its own rows of) the Q tile ----
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
    __syncwarp();      // all lanes have finished their ldmatrix reads of this warp's Q rows
    {
        const int r0 = warp * 16 + g, r1 = r0 + 8;
        const uint32_t ob0 = sQ + (r0 >> 6) * kTileBytes, ob1 = sQ + (r1 >> 6) * kTileBytes;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            uint32_t v0 = 0u, v1 = 0u;
            if (c < ND8) {
                v0 = pack_bf16x2(oacc[c < ND8 ? c : 0][0] * inv0, oacc[c < ND8 ? c : 0][1] * inv0);
                v1 = pack_bf16x2(oacc[c < ND8 ? c : 0][2] * inv1, oacc[c < ND8 ? c : 0][3] * inv1);
            }
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(ob0 + swz(r0 & 63, c) + 4 * t4), "r"(v0) : "memory");
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(ob1 + swz(r1 & 63, c) + 4 * t4), "r"(v1) : "memory");
        }
    }
    __syncwarp();
    // 16 rows x 8 chunks of 16 B: 4 per lane, row-contiguous 128 B segments in global memory.
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int idx = it * 32 + lane;
        const int r = warp * 16 + (idx >> 3), c = idx & 7;
        const int q = q0 + r;
        if (q < p.T) {
            uint4 v;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "r"(sQ + (r >> 6) * kTileBytes + swz(r & 63, c)));
            *reinterpret_cast<uint4*>(p.ctx + static_cast<size_t>(row_base + q) * p.ldo + h * kHP + c * 8) = v;
        }
    }
}

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

template <int NK16, int ND8>
cudaError_t launch_bf16(const AttnPlan& plan, const AttnArgs& a, dim3 grid, size_t smem, cudaStream_t st) {
    auto kern = attn_bf16_kernel<NK16, ND8>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    kern<<<grid, kAttnThreads, smem, st>>>(plan.tmQKV, a);
    return cudaGetLastError();
}

}  // namespace

int attn_bf16_make_plan(AttnPlan* plan, const AttnDesc& d) {
    if (d.hp != kHP || d.d <= 0 || d.d > kHP) return -20;       // key_dim > 64 not supported in this build
    if ((d.ldq % 8) || (d.ldo % 8) || d.ldq < 3 * d.H * d.hp || d.ldo < d.H * d.hp) return -21;
    if ((reinterpret_cast<uintptr_t>(d.qkv) & 15) || (reinterpret_cast<uintptr_t>(d.ctx) & 15)) return -22;
    plan->desc = d;
    return make_tmap_bf16_2d(&plan->tmQKV, d.qkv, d.B * d.T, 3 * d.H * d.hp, d.ldq, kBKV);
}

cudaError_t attn_bf16_launch(const AttnPlan& plan, cudaStream_t stream) {
    const AttnDesc& d = plan.desc;
    AttnArgs a;
    a.ctx = static_cast<__nv_bfloat16*>(d.ctx);
    a.ldo = d.ldo;
    a.T = d.T;
    a.H = d.H;
    a.scale_log2 = d.scale * 1.4426950408889634f;
    dim3 grid((d.T + kBQ - 1) / kBQ, d.B * d.H);
    const size_t smem = 1024 + 2 * kTileBytes + kStages * 2 * kTileBytes;
    switch ((d.d + 7) / 8) {
        case 1: return launch_bf16<1, 1>(plan, a, grid, smem, stream);
        case 2: return launch_bf16<1, 2>(plan, a, grid, smem, stream);
        case 3: return launch_bf16<2, 3>(plan, a, grid, smem, stream);
        case 4: return launch_bf16<2, 4>(plan, a, grid, smem, stream);
        case 5: return launch_bf16<3, 5>(plan, a, grid, smem, stream);
        case 6: return launch_bf16<3, 6>(plan, a, grid, smem, stream);
        case 7: return launch_bf16<4, 7>(plan, a, grid, smem, stream);
        case 8: return launch_bf16<4, 8>(plan, a, grid, smem, stream);
    }
    return cudaErrorInvalidValue;
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
        default: return cudaErrorInvalidValue;   // key_dim <= 64 in this build
    }
#undef VITDET_ATTN_F32
    return cudaGetLastError();
}

}  // namespace vitdet
