// Multi-head self-attention core, second generation: softmax(q k^T / sqrt(d)) v per (image, head) on tcgen05 with the
// score row of a query SPLIT OVER TWO WARPS (keras.layers.MultiHeadAttention, reference det.py:364-369).
//
// Why: the first-generation kernel (attention_tc.cu) has one softmax thread per query row, i.e. four softmax warps per
// 128-query tile and — at two CTAs per SM, which is all the 512 TMEM columns allow — two softmax warps per SM
// sub-partition.  Its profile (profiles/r01e_ncu_attn_tc.md and the per-instruction samples) shows the binding unit,
// the SFU (one ex2 per score, 16/clk/SM), only 66 % busy: half of a softmax warp's time is the exponentials, the rest
// is TMEM load latency, the row maximum, barrier hand-offs and waiting for the slowest of its three sibling warps, and
// two warps per sub-partition cannot cover that.  Here the eight softmax warps of a CTA pair up per TMEM lane quadrant
// (warps w and w + 4 may both touch lanes [32 (w & 3), +32)): warp `half` of a pair owns keys [64 half, +64) of every
// 128-key tile.  Same TMEM footprint (S 128 | P 64 | O 64 columns), same MMA shapes, but four softmax warps per
// sub-partition, 64 instead of 128 live scores per thread (two CTAs of 320 threads fit the register file) and half the
// exposed latency per exponential.  The pair shares the row maximum through shared memory (one float per row and tile
// + a 64-thread named barrier) so that both halves of P(j) are scaled against the same reference; the row sums are
// combined once at the end.
//
//   warps 0..7  softmax (quad = w & 3, half = w >> 2)
//   warp  8     TMA producer: Q once, then 128-key K tiles through a 2-slot and V tiles through a 3-slot mbarrier ring
//               (K(j) is free again as soon as QK^T(j) has run, V(j) only after PV(j) a tile later: separate rings hold
//               what a 3-stage K+V ring would in 80 KB instead of 96 KB — with the 2 KB exchange buffer that is what lets
//               two CTAs share an SM)
//   warp  9     TMEM allocation; one elected lane issues
//                 S = Q K(j)^T   tcgen05.mma M128 x N128 x K(16 ceil(d/16)), both operands K-major from smem
//                 O += P V(j)    tcgen05.mma M128 x N(16 ceil(d/16)) x K128, A = P from TENSOR MEMORY, B = V MN-major
// Softmax warps whose 32 query rows all lie past the end of the image (three of four quadrants in the last query tile at
// T = 1296) only keep the barrier protocol going: no TMEM traffic, no exponentials.
#include "common.cuh"
#include "kernels.h"
#include "launch.h"

#include <cstdlib>

namespace vitdet {

namespace {

constexpr int kQ = 128;            // queries per CTA (UMMA M)
constexpr int kKV = 128;           // keys per tile (UMMA N of QK^T, K of PV)
constexpr int kHP = 64;            // head pitch in shared memory (one 128-byte swizzle row of bf16)
constexpr int kKSlots = 2, kVSlots = 3;
constexpr int kSoftmaxWarps = 8;
constexpr int kThreads = 32 * (kSoftmaxWarps + 2);
constexpr int kProducerWarp = kSoftmaxWarps, kMmaWarp = kSoftmaxWarps + 1;      // highest warp ids: favoured by the arbiter
constexpr int kQBytes = kQ * kHP * 2;         // 16 KiB
constexpr int kTileBytes = kKV * kHP * 2;     // 16 KiB: one K tile or one V tile = two TMA boxes of 64 rows
constexpr int kBoxBytes = 64 * kHP * 2;
constexpr int kTmemCols = 256;
constexpr uint32_t kColS = 0, kColP = 128, kColO = 192;   // S [0,128) f32 | P [128,192) bf16x2 | O [192,256) f32
constexpr float kRescaleThreshold = 8.f;      // log2 units: P stays <= 2^8 between rescales

struct AttnTc8Args {
    __nv_bfloat16* ctx;
    int ldo;
    int T, H;
    int hp;              // elements per head in qkv / ctx (key_dim rounded up to 8)
    int k16;             // ceil(d / 16): K steps of QK^T
    int n_pv;            // N of the PV product: 16 * k16 (VITDET_ATTN_PVN=64 forces the full 64-column V tile, A/B switch)
    float scale_log2;    // log2(e) / sqrt(key_dim)
};

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void pair_barrier(int quad) {       // the two warps of a TMEM lane quadrant
    asm volatile("bar.sync %0, 64;" ::"r"(quad + 1) : "memory");
}

struct TileBars { uint32_t s_free, pv_done, p_full; };

// One key tile of the online softmax for the calling thread's (query row, key half): NC = number of 32-key chunks of the
// half that hold at least one existing key (2 for a full tile, 0 when the half lies past the end of the image), MASK =
// the last of them is partial.  Static loops only, so that the 64 scores stay in registers.
template <int NC, bool MASK>
__device__ __forceinline__ void softmax_half_tile(uint32_t tS, uint32_t tP, uint32_t tO, const TileBars& b, float* xch_mine,
                                                  const float* xch_other, int lane, int quad, int half, int j, int valid_h,
                                                  int n_pv, float scale_log2, float& m_used, float& l) {
    uint32_t v[NC > 0 ? NC : 1][32];
#pragma unroll
    for (int c = 0; c < NC; ++c) tmem_ld_32x32(tS + 32u * c, v[c]);
    if (NC > 0) tmem_ld_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(b.s_free);           // S(j) is in registers (or not needed): QK^T(j+1) may overwrite it

    if (MASK) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (32 * (NC - 1) + i >= valid_h) v[NC - 1][i] = 0xff800000u;     // -inf: keys past the end of the image
    }
    float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            mx[0] = fmaxf(mx[0], __uint_as_float(v[c][i]));     mx[1] = fmaxf(mx[1], __uint_as_float(v[c][i + 1]));
            mx[2] = fmaxf(mx[2], __uint_as_float(v[c][i + 2])); mx[3] = fmaxf(mx[3], __uint_as_float(v[c][i + 3]));
        }
    // both halves of the row must scale P(j) against the same reference: exchange the half-row maxima
    *xch_mine = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
    pair_barrier(quad);
    const float m_row = fmaxf(*xch_mine, *xch_other);
    const float m_new = fmaxf(m_used, m_row * scale_log2);
    const bool grow = __any_sync(0xffffffffu, m_new > m_used + kRescaleThreshold);   // true on the first tile; same in both warps
    float alpha = 1.f;
    if (grow) {
        alpha = ex2f(m_used - m_new);       // 0 on the first tile (m_used = -inf)
        m_used = m_new;
        l *= alpha;
    }
    const float neg_m = -m_used;
    // p = 2^(s*c - m) in place; packed pairs overwrite the first half of each chunk's registers
    const uint64_t sc2 = f2_pack(scale_log2, scale_log2), nm2 = f2_pack(neg_m, neg_m);
    uint64_t sum2[2] = {0ull, 0ull};
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            float t0, t1, t2, t3;
            f2_unpack(f2_fma(f2_pack(__uint_as_float(v[c][i]), __uint_as_float(v[c][i + 1])), sc2, nm2), t0, t1);
            f2_unpack(f2_fma(f2_pack(__uint_as_float(v[c][i + 2]), __uint_as_float(v[c][i + 3])), sc2, nm2), t2, t3);
            const float e0 = ex2f(t0), e1 = ex2f(t1), e2 = ex2f(t2), e3 = ex2f(t3);      // ex2(-inf) = 0
            sum2[0] = f2_add(sum2[0], f2_pack(e0, e1));
            sum2[1] = f2_add(sum2[1], f2_pack(e2, e3));
            v[c][i / 2] = pack_bf16x2(e0, e1);
            v[c][i / 2 + 1] = pack_bf16x2(e2, e3);
        }
    {
        float s0, s1, s2, s3;
        f2_unpack(sum2[0], s0, s1);
        f2_unpack(sum2[1], s2, s3);
        l += (s0 + s1) + (s2 + s3);
    }

    // P and O must no longer be in use by PV(j-1)
    if (j > 0) {
        mbar_wait(b.pv_done, (j - 1) & 1);
        tc_fence_after();
        if (grow && 32 * half < n_pv) {         // this warp rescales O columns [32 half, +32)
            uint32_t o[32];
            tmem_ld_32x32(tO + 32u * half, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32_x32(tO + 32u * half, o);
        }
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = (c < NC) ? v[c < NC ? c : 0][i] : 0u;     // P = 0 for keys that do not exist
        tmem_st_32x32_x16(tP + 16u * c, pk);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(b.p_full);
}

__global__ void __launch_bounds__(kThreads, 2)
attn_tc8_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnTc8Args p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * (kKSlots + kVSlots) + 5];
    __shared__ uint32_t tmem_base_s;
    __shared__ float xch[2][2][kQ];        // [tile parity][key half][query row]: half-row maxima, and the row sums at the end

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    pdl_launch_dependents();
    const int q0 = blockIdx.x * kQ;
    const int bh = blockIdx.y;
    const int b = bh / p.H, h = bh - b * p.H;
    const int row_base = b * p.T;          // first token row of this image in the [B*T, ld] matrices
    const int nkv = (p.T + kKV - 1) / kKV;
    const int n_pv = p.n_pv;               // PV accumulator width: key_dim rounded up to the UMMA N granularity for M = 128

    const uint32_t base = smem_u32(smem_raw);
    if ((base & 1023u) != 0u) __trap();
    const uint32_t sQ = base;
    const uint32_t sK = base + kQBytes;                // K slot s at + s * tile
    const uint32_t sV = sK + kKSlots * kTileBytes;     // V slot s at + s * tile
    constexpr int kRing = kKSlots + kVSlots;
    const uint32_t bar_kfull = smem_u32(&bars[0]);
    const uint32_t bar_vfull = smem_u32(&bars[kKSlots]);
    const uint32_t bar_kempty = smem_u32(&bars[kRing]);
    const uint32_t bar_vempty = smem_u32(&bars[kRing + kKSlots]);
    const uint32_t bar_q = smem_u32(&bars[2 * kRing]);
    const uint32_t bar_s_full = smem_u32(&bars[2 * kRing + 1]);   // QK^T(j) complete                 (MMA commit)
    const uint32_t bar_s_free = smem_u32(&bars[2 * kRing + 2]);   // S(j) is in registers             (8 warps)
    const uint32_t bar_p_full = smem_u32(&bars[2 * kRing + 3]);   // P(j) (and rescaled O) in TMEM    (8 warps)
    const uint32_t bar_pv_done = smem_u32(&bars[2 * kRing + 4]);  // PV(j) complete                   (MMA commit)

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2 * kRing; ++s) mbar_init(bar_kfull + 8 * s, 1);
        mbar_init(bar_q, 1);
        mbar_init(bar_s_full, 1);
        mbar_init(bar_s_free, kSoftmaxWarps);
        mbar_init(bar_p_full, kSoftmaxWarps);
        mbar_init(bar_pv_done, 1);
        fence_mbar_init();
    }
    if (warp == kMmaWarp) {
        tmem_alloc(smem_u32(&tmem_base_s), kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    pdl_wait();       // set-up above overlapped the previous kernel; q/k/v are read from here on

    if (warp == kProducerWarp) {
        // ------------------------------ TMA producer ------------------------------
        if (lane == 0) {
            tma_prefetch_desc(&tmQKV);
            mbar_arrive_expect_tx(bar_q, kQBytes);
            tma_load_2d(sQ, &tmQKV, bar_q, h * p.hp, row_base + q0);
            tma_load_2d(sQ + kBoxBytes, &tmQKV, bar_q, h * p.hp, row_base + q0 + 64);
            int ks = 0, vs = 0;
            uint32_t kphase = 0, vphase = 0;
            for (int j = 0; j < nkv; ++j) {
                const int r = row_base + j * kKV;
                mbar_wait_relaxed(bar_kempty + 8 * ks, kphase ^ 1u);
                mbar_arrive_expect_tx(bar_kfull + 8 * ks, kTileBytes);
                tma_load_2d(sK + ks * kTileBytes, &tmQKV, bar_kfull + 8 * ks, (p.H + h) * p.hp, r);
                tma_load_2d(sK + ks * kTileBytes + kBoxBytes, &tmQKV, bar_kfull + 8 * ks, (p.H + h) * p.hp, r + 64);
                if (++ks == kKSlots) { ks = 0; kphase ^= 1u; }
                mbar_wait_relaxed(bar_vempty + 8 * vs, vphase ^ 1u);
                mbar_arrive_expect_tx(bar_vfull + 8 * vs, kTileBytes);
                tma_load_2d(sV + vs * kTileBytes, &tmQKV, bar_vfull + 8 * vs, (2 * p.H + h) * p.hp, r);
                tma_load_2d(sV + vs * kTileBytes + kBoxBytes, &tmQKV, bar_vfull + 8 * vs, (2 * p.H + h) * p.hp, r + 64);
                if (++vs == kVSlots) { vs = 0; vphase ^= 1u; }
            }
        }
    } else if (warp == kMmaWarp) {
        // ------------------------------ MMA issuer --------------------------------
        // The whole warp runs the loop (warp-uniform control flow); only the tcgen05 instructions are issued by one
        // elected lane.
        const uint32_t idesc_qk = umma_idesc_bf16_f32(kQ, kKV);
        const uint32_t idesc_pv = umma_idesc_bf16_f32_bmn(kQ, n_pv);
        const uint32_t tS = tmem_base + kColS, tP = tmem_base + kColP, tO = tmem_base + kColO;
        const uint64_t dq = umma_desc_sw128_kmajor(sQ);
        const uint64_t dk0 = umma_desc_sw128_kmajor(sK), dv0 = umma_desc_sw128_kmajor(sV);
        constexpr uint32_t kSlotStep = kTileBytes >> 4;     // descriptor address units (16 B)
        mbar_wait(bar_q, 0);
        // Heads are stored hp (< 64) columns apart, so the 64-column TMA boxes also carry the first columns of the
        // next head.  In QK^T only the columns below 16 * k16 take part: clearing Q's columns [hp, 16 * k16) once makes
        // their products vanish whatever K holds there; V's extra columns only produce columns of O that are never
        // stored.  16-byte chunk c of row r sits at chunk c ^ (r & 7) of the 128-byte swizzled row.
        if (p.hp < 16 * p.k16) {
            const int c_lo = p.hp >> 3, c_hi = 2 * p.k16;
            for (int r = lane; r < kQ; r += 32)
                for (int c = c_lo; c < c_hi; ++c)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sQ + r * 128 + ((c ^ (r & 7)) << 4)), "r"(0u) : "memory");
            fence_proxy_async_smem();
            __syncwarp();
        }
        int ks = 0, vs = 0;
        uint32_t kphase = 0, vphase = 0;
        for (int j = 0; j <= nkv; ++j) {
            if (j < nkv) {
                // S = Q K(j)^T; the softmax warps moved S(j-1) into registers before signalling s_free
                mbar_wait(bar_kfull + 8 * ks, kphase);
                if (j >= 1) mbar_wait(bar_s_free, (j - 1) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t dk = dk0 + static_cast<uint64_t>(ks * kSlotStep);
                    umma_bf16_ss(tS, dq, dk, idesc_qk, 0u);
                    if (p.k16 > 1) umma_bf16_ss(tS, dq + 2u, dk + 2u, idesc_qk, 1u);
                    if (p.k16 > 2) umma_bf16_ss(tS, dq + 4u, dk + 4u, idesc_qk, 1u);
                    if (p.k16 > 3) umma_bf16_ss(tS, dq + 6u, dk + 6u, idesc_qk, 1u);
                    umma_commit(bar_kempty + 8 * ks);
                    umma_commit(bar_s_full);
                }
                __syncwarp();
                if (++ks == kKSlots) { ks = 0; kphase ^= 1u; }
            }
            if (j > 0) {
                // O += P(j-1) V(j-1)
                mbar_wait(bar_vfull + 8 * vs, vphase);
                mbar_wait(bar_p_full, (j - 1) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t dv = dv0 + static_cast<uint64_t>(vs * kSlotStep);
                    // 16 keys per step: 8 packed columns of P, 16 rows (2048 B) of V
#pragma unroll
                    for (int k = 0; k < kKV / 16; ++k)
                        umma_bf16_ts(tO, tP + 8u * k, dv + static_cast<uint64_t>(128u * k), idesc_pv, (k != 0) ? 1u : (j > 1 ? 1u : 0u));
                    umma_commit(bar_vempty + 8 * vs);
                    umma_commit(bar_pv_done);
                }
                __syncwarp();
                if (++vs == kVSlots) { vs = 0; vphase ^= 1u; }
            }
        }
    } else {
        // ------------------------------ softmax -----------------------------------
        const int quad = warp & 3, half = warp >> 2;        // TMEM lane quadrant; which 64 keys of every tile
        const int r = quad * 32 + lane;                     // query row within the tile
        const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
        const uint32_t tS = tmem_base + lane_off + kColS + 64u * half;
        const uint32_t tP = tmem_base + lane_off + kColP + 32u * half;
        const uint32_t tO = tmem_base + lane_off + kColO;
        const TileBars tb{bar_s_free, bar_pv_done, bar_p_full};
        const bool rows_exist = q0 + quad * 32 < p.T;       // warp-uniform, and the same in both warps of the pair
        float m_used = -INFINITY;      // running maximum in the scaled log2 domain (identical in both warps of a pair)
        float l = 0.f;                 // running sum of p over this warp's key half
        for (int j = 0; j < nkv; ++j) {
            const int valid = min(kKV, p.T - j * kKV);      // keys of this tile that exist
            const int valid_h = max(0, min(64, valid - 64 * half));
            mbar_wait(bar_s_full, j & 1);
            tc_fence_after();
            if (!rows_exist) {
                // no query of this warp exists: keep the protocol going, touch nothing
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_s_free);
                if (j > 0) mbar_wait(bar_pv_done, (j - 1) & 1);      // phase j-1 of p_full is complete before the next arrive
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_p_full);
                continue;
            }
            float* xm = &xch[j & 1][half][r];
            const float* xo = &xch[j & 1][half ^ 1][r];
#define VITDET_TILE(NC, MASK) \
    softmax_half_tile<NC, MASK>(tS, tP, tO, tb, xm, xo, lane, quad, half, j, valid_h, n_pv, p.scale_log2, m_used, l)
            if (valid_h == 64) {
                VITDET_TILE(2, false);
            } else {
                // last tile of the image: only the chunks with existing keys are loaded and exponentiated; warp-uniform
                const bool partial = (valid_h & 31) != 0;
                switch ((valid_h + 31) >> 5) {
                    case 0: VITDET_TILE(0, false); break;
                    case 1: if (partial) VITDET_TILE(1, true); else VITDET_TILE(1, false); break;
                    default: VITDET_TILE(2, true); break;
                }
            }
#undef VITDET_TILE
        }

        if (rows_exist) {
            // ---- finalise: O / (l_half0 + l_half1) -> bf16 context rows; each warp stores its 32 columns ----
            mbar_wait(bar_pv_done, (nkv - 1) & 1);
            tc_fence_after();
            xch[nkv & 1][half][r] = l;
            pair_barrier(quad);
            const float inv = 1.f / (l + xch[nkv & 1][half ^ 1][r]);
            const int q = q0 + r;
            if (32 * half < p.hp) {
                uint32_t o[32];
                tmem_ld_32x32(tO + 32u * half, o);
                tmem_ld_wait();
                if (q < p.T) {
                    __nv_bfloat16* orow = p.ctx + static_cast<size_t>(row_base + q) * p.ldo + h * p.hp + 32 * half;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if (32 * half + 8 * g >= p.hp) break;          // the head holds hp columns
                        uint4 w;
                        w.x = pack_bf16x2(__uint_as_float(o[8 * g + 0]) * inv, __uint_as_float(o[8 * g + 1]) * inv);
                        w.y = pack_bf16x2(__uint_as_float(o[8 * g + 2]) * inv, __uint_as_float(o[8 * g + 3]) * inv);
                        w.z = pack_bf16x2(__uint_as_float(o[8 * g + 4]) * inv, __uint_as_float(o[8 * g + 5]) * inv);
                        w.w = pack_bf16x2(__uint_as_float(o[8 * g + 6]) * inv, __uint_as_float(o[8 * g + 7]) * inv);
                        *reinterpret_cast<uint4*>(orow + 8 * g) = w;
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace

cudaError_t attn_tc8_launch(const AttnPlan& plan, cudaStream_t stream) {
    const AttnDesc& d = plan.desc;
    AttnTc8Args a;
    a.ctx = static_cast<__nv_bfloat16*>(d.ctx);
    a.ldo = d.ldo;
    a.T = d.T;
    a.H = d.H;
    a.hp = d.hp;
    a.k16 = (d.d + 15) / 16;
    static int pvn = -1;
    if (pvn < 0) { const char* e = getenv("VITDET_ATTN_PVN"); pvn = e ? atoi(e) : 0; }
    a.n_pv = (pvn == 64) ? 64 : 16 * a.k16;
    a.scale_log2 = d.scale * 1.4426950408889634f;
    const size_t smem = static_cast<size_t>(kQBytes) + static_cast<size_t>(kKSlots + kVSlots) * kTileBytes;
    cudaError_t e = ensure_max_dynamic_smem(reinterpret_cast<const void*>(attn_tc8_kernel), static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    dim3 grid((d.T + kQ - 1) / kQ, d.B * d.H);
    return launch_kernel(attn_tc8_kernel, grid, dim3(kThreads), smem, stream, 1, plan.tmQKV, a);
}

}  // namespace vitdet
