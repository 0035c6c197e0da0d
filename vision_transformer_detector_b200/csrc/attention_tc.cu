// Multi-head self-attention core on the 5th-gen tensor cores: softmax(q k^T / sqrt(d)) v per
// (image, head) without materialising the (B, H, T, T) score tensor of keras.layers.MultiHeadAttention
// (reference det.py:364-369).
//
// One CTA = 128 queries of one (image, head); two CTAs are resident per SM so that the softmax of one
// overlaps the MMAs of the other.  192 threads:
//   warp 0      TMA producer: Q once, then 128-key K and V tiles (cp.async.bulk.tensor, 128B swizzle)
//               through a 3-stage mbarrier ring.
//   warp 1      allocates 256 TMEM columns; one lane issues
//                 S = Q K^T      tcgen05.mma  M128 x N128 x K(16*ceil(d/16)), A and B K-major from smem
//                 O += P V       tcgen05.mma  M128 x N64  x K128, A = P from TENSOR MEMORY, B = V MN-major
//               QK^T of tile j+1 is issued before PV of tile j, so S(j+1) is ready when the softmax
//               warps finish tile j.
//   warps 2..5  softmax: thread = query row (TMEM lane).  Pass A reads S (tcgen05.ld) for the row
//               maximum — no shuffles, a thread owns its row; pass B re-reads S, p = ex2(s*c - m),
//               accumulates the row sum and writes P as packed bf16 back to TMEM (tcgen05.st), never
//               through shared memory.  The running maximum is only advanced (and O rescaled in TMEM)
//               when it grew by more than 2^8, so most tiles skip the correction.
// TMEM columns: S [0,128) f32 | P [128,192) bf16x2 | O [192,256) f32.
//
// The binding unit is the SFU, not the tensor pipe: per 128 x 128 score tile the CTA needs 16 384 ex2
// at 16/clk/SM = 1024 clk, against ~200 clk of MMA at head_dim 40.
#include "common.cuh"
#include "kernels.h"

namespace vitdet {

namespace {

constexpr int kQ = 128;            // queries per CTA (UMMA M)
constexpr int kKV = 128;           // keys per tile (UMMA N of QK^T, K of PV)
constexpr int kHP = 64;            // head pitch in elements (one 128-byte swizzle row of bf16)
constexpr int kStages = 3;
constexpr int kThreads = 192;
constexpr int kTileBytes = 128 * kHP * 2;     // 16 KiB: Q, one K tile or one V tile
constexpr int kHalfBytes = kTileBytes / 2;    // one TMA box: 64 rows
constexpr int kTmemCols = 256;
constexpr uint32_t kColS = 0, kColP = 128, kColO = 192;
constexpr float kRescaleThreshold = 8.f;      // log2 units: P stays <= 2^8 between rescales

struct AttnTcArgs {
    __nv_bfloat16* ctx;
    int ldo;
    int T, H;
    int k16;             // ceil(d / 16): K steps of the QK^T product
    float scale_log2;    // log2(e) / sqrt(key_dim)
};

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(kThreads, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnTcArgs p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * kStages + 5];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * kQ;
    const int bh = blockIdx.y;
    const int b = bh / p.H, h = bh - b * p.H;
    const int row_base = b * p.T;          // first token row of this image in the [B*T, ld] matrices
    const int nkv = (p.T + kKV - 1) / kKV;

    const uint32_t base = smem_u32(smem_raw);
    if ((base & 1023u) != 0u) __trap();
    const uint32_t sQ = base;
    const uint32_t sKV = base + kTileBytes;            // stage s: K at + 2*s*tile, V right after
    const uint32_t bar_full = smem_u32(&bars[0]);
    const uint32_t bar_empty = smem_u32(&bars[kStages]);
    const uint32_t bar_q = smem_u32(&bars[2 * kStages]);
    const uint32_t bar_s_full = smem_u32(&bars[2 * kStages + 1]);   // QK^T(j) complete           (MMA commit)
    const uint32_t bar_s_free = smem_u32(&bars[2 * kStages + 2]);   // softmax has read S(j)       (4 warps)
    const uint32_t bar_p_full = smem_u32(&bars[2 * kStages + 3]);   // P(j) (and rescaled O) in TMEM (4 warps)
    const uint32_t bar_pv_done = smem_u32(&bars[2 * kStages + 4]);  // PV(j) complete              (MMA commit)

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_q, 1);
        mbar_init(bar_s_full, 1);
        mbar_init(bar_s_free, 4);
        mbar_init(bar_p_full, 4);
        mbar_init(bar_pv_done, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(&tmem_base_s), kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------
        if (lane == 0) {
            tma_prefetch_desc(&tmQKV);
            mbar_arrive_expect_tx(bar_q, kTileBytes);
            tma_load_2d(sQ, &tmQKV, bar_q, h * kHP, row_base + q0);
            tma_load_2d(sQ + kHalfBytes, &tmQKV, bar_q, h * kHP, row_base + q0 + 64);
            int stage = 0;
            uint32_t phase = 0;
            for (int j = 0; j < nkv; ++j) {
                mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
                mbar_arrive_expect_tx(bar_full + 8 * stage, 2 * kTileBytes);
                const uint32_t dK = sKV + stage * 2 * kTileBytes, dV = dK + kTileBytes;
                const int r = row_base + j * kKV;
                tma_load_2d(dK, &tmQKV, bar_full + 8 * stage, (p.H + h) * kHP, r);
                tma_load_2d(dK + kHalfBytes, &tmQKV, bar_full + 8 * stage, (p.H + h) * kHP, r + 64);
                tma_load_2d(dV, &tmQKV, bar_full + 8 * stage, (2 * p.H + h) * kHP, r);
                tma_load_2d(dV + kHalfBytes, &tmQKV, bar_full + 8 * stage, (2 * p.H + h) * kHP, r + 64);
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issuer --------------------------------
        if (lane == 0) {
            const uint32_t idesc_qk = umma_idesc_bf16_f32(kQ, kKV);
            const uint32_t idesc_pv = umma_idesc_bf16_f32_bmn(kQ, kHP);
            const uint32_t tS = tmem_base + kColS, tP = tmem_base + kColP, tO = tmem_base + kColO;
            const uint64_t dq = umma_desc_sw128_kmajor(sQ);
            mbar_wait(bar_q, 0);
            int stage = 0;
            uint32_t phase = 0;
            for (int j = 0; j <= nkv; ++j) {
                if (j < nkv) {
                    // S(j) = Q K(j)^T
                    mbar_wait(bar_full + 8 * stage, phase);
                    if (j > 0) mbar_wait(bar_s_free, (j - 1) & 1);
                    tc_fence_after();
                    const uint64_t dk = umma_desc_sw128_kmajor(sKV + stage * 2 * kTileBytes);
                    for (int k = 0; k < p.k16; ++k) umma_bf16_ss(tS, dq + 2u * k, dk + 2u * k, idesc_qk, k != 0);
                    umma_commit(bar_s_full);
                }
                if (j > 0) {
                    // O += P(j-1) V(j-1)
                    const int ps = (j - 1) % kStages;
                    mbar_wait(bar_p_full, (j - 1) & 1);
                    tc_fence_after();
                    const uint64_t dv = umma_desc_sw128_mnmajor(sKV + ps * 2 * kTileBytes + kTileBytes);
                    // 16 keys per step: 8 packed columns of P, 16 rows (2048 B) of V
                    for (int k = 0; k < kKV / 16; ++k)
                        umma_bf16_ts(tO, tP + 8u * k, dv + static_cast<uint64_t>(128u * k), idesc_pv, (j > 1) || (k != 0));
                    umma_commit(bar_empty + 8 * ps);
                    umma_commit(bar_pv_done);
                }
                if (j < nkv) { if (++stage == kStages) { stage = 0; phase ^= 1u; } }
            }
        }
    } else {
        // ------------------------------ softmax -----------------------------------
        const int quad = warp & 3;                          // TMEM lane quadrant of this warp
        const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
        const uint32_t tS = tmem_base + lane_off + kColS, tP = tmem_base + lane_off + kColP,
                       tO = tmem_base + lane_off + kColO;
        float m_used = -INFINITY;      // running maximum in the scaled log2 domain
        float l = 0.f;                 // running sum of p
        for (int j = 0; j < nkv; ++j) {
            const int valid = min(kKV, p.T - j * kKV);      // keys of this tile that exist
            const int nch = (valid + 31) >> 5;              // 32-key chunks holding at least one valid key
            mbar_wait(bar_s_full, j & 1);
            tc_fence_after();

            // ---- pass A: row maximum ----
            float mx = -INFINITY;
            for (int c = 0; c < nch; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(tS + 32u * c, v);
                tmem_ld_wait();
                if (32 * c + 32 <= valid) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, (32 * c + i < valid) ? __uint_as_float(v[i]) : -INFINITY);
                }
            }
            const float m_new = fmaxf(m_used, mx * p.scale_log2);
            const bool grow = __any_sync(0xffffffffu, m_new > m_used + kRescaleThreshold);   // true on the first tile
            float alpha = 1.f;
            if (grow) {
                alpha = ex2f(m_used - m_new);       // 0 on the first tile (m_used = -inf)
                m_used = m_new;
                l *= alpha;
            }

            // P(j-1) and O must no longer be in use by PV(j-1)
            if (j > 0) {
                mbar_wait(bar_pv_done, (j - 1) & 1);
                tc_fence_after();
                if (grow) {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        uint32_t o[32];
                        tmem_ld_32x32(tO + 32u * c, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st_32x32_x32(tO + 32u * c, o);
                    }
                }
            }

            // ---- pass B: p = 2^(s*c - m), row sum, P -> TMEM as packed bf16 ----
            const float neg_m = -m_used;
            for (int c = 0; c < 4; ++c) {
                uint32_t pk[16];
                if (c < nch) {
                    uint32_t v[32];
                    tmem_ld_32x32(tS + 32u * c, v);
                    tmem_ld_wait();
                    if (c == nch - 1) {
                        // every S read of this tile is complete: QK^T(j+1) may overwrite S
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_s_free);
                    }
                    float e[32];
                    if (32 * c + 32 <= valid) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) e[i] = ex2f(fmaf(__uint_as_float(v[i]), p.scale_log2, neg_m));
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            e[i] = (32 * c + i < valid) ? ex2f(fmaf(__uint_as_float(v[i]), p.scale_log2, neg_m)) : 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < 32; i += 4) l += (e[i] + e[i + 1]) + (e[i + 2] + e[i + 3]);
#pragma unroll
                    for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(e[2 * i], e[2 * i + 1]);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) pk[i] = 0u;      // keys past the end of the image contribute nothing
                }
                tmem_st_32x32_x16(tP + 16u * c, pk);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_p_full);
        }

        // ---- finalise: O / l -> bf16 context rows ----
        mbar_wait(bar_pv_done, (nkv - 1) & 1);
        tc_fence_after();
        const int q = q0 + quad * 32 + lane;
        const float inv = 1.f / l;
        __nv_bfloat16* orow = p.ctx + static_cast<size_t>(row_base + (q < p.T ? q : 0)) * p.ldo + h * kHP;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld_32x32(tO + 32u * c, o);
            tmem_ld_wait();
            if (q < p.T) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 w;
                    w.x = pack_bf16x2(__uint_as_float(o[8 * g + 0]) * inv, __uint_as_float(o[8 * g + 1]) * inv);
                    w.y = pack_bf16x2(__uint_as_float(o[8 * g + 2]) * inv, __uint_as_float(o[8 * g + 3]) * inv);
                    w.z = pack_bf16x2(__uint_as_float(o[8 * g + 4]) * inv, __uint_as_float(o[8 * g + 5]) * inv);
                    w.w = pack_bf16x2(__uint_as_float(o[8 * g + 6]) * inv, __uint_as_float(o[8 * g + 7]) * inv);
                    *reinterpret_cast<uint4*>(orow + 32 * c + 8 * g) = w;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace

cudaError_t attn_tc_launch(const AttnPlan& plan, cudaStream_t stream) {
    const AttnDesc& d = plan.desc;
    AttnTcArgs a;
    a.ctx = static_cast<__nv_bfloat16*>(d.ctx);
    a.ldo = d.ldo;
    a.T = d.T;
    a.H = d.H;
    a.k16 = (d.d + 15) / 16;
    a.scale_log2 = d.scale * 1.4426950408889634f;
    const size_t smem = static_cast<size_t>(kTileBytes) * (1 + 2 * kStages);
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    dim3 grid((d.T + kQ - 1) / kQ, d.B * d.H);
    attn_tc_kernel<<<grid, kThreads, smem, stream>>>(plan.tmQKV, a);
    return cudaGetLastError();
}

}  // namespace vitdet
