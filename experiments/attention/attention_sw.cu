// Multi-head self-attention core: persistent kernel with a SOFTWARE-PIPELINED softmax warp — softmax(q k^T / sqrt(d)) v
// per (image, head) on tcgen05 without materialising the (B, H, T, T) score tensor of keras.layers.MultiHeadAttention
// (reference det.py:364-369).
//
// What bounds the earlier kernels (profiles/r02_attention_analysis.md): the SFU accepts one warp-wide ex2 every 8 clocks
// per SM sub-partition, but ONE warp gets at most one in every ~16 — so the SFU is only saturated while both softmax
// warps of a sub-partition (two CTAs per SM) are inside their exponential phase, and every clock a warp spends on
// anything else (TMEM load of the next scores, row maximum, P to TMEM, barrier hand-offs: ~900 of ~3000 clocks per
// tile) is SFU time lost.  More warps (split rows), explicit turn-taking and FMA-pipe exponentials were all measured and
// lost; what is left is to take the "anything else" off the warp's critical path.  A warp issues a MUFU every 16 clocks
// and has ~13 free issue slots in between: here the per-tile work is reordered so that those slots carry the next
// tile's work.  Per 32-key chunk c of tile g the warp
//     exponentiates chunk c of S(g) (registers) and stores that quarter of P(g) to TMEM,
//     reloads the freed registers with chunk c of S(g+1) (QK^T(g+1) ran while S(g) was being exponentiated),
//     folds chunk c-1 of S(g+1) — loaded one chunk earlier — into the next row maximum,
// so TMEM latency, the maximum and the stores overlap the exponentials of the same warp.  Everything around the softmax
// warps (persistent flat tile loop, Q double-buffered, split K / V rings, MMA issue order) is attention_tcp.cu's.
#include "common.cuh"
#include "kernels.h"
#include "launch.h"

#include <cstdlib>

namespace vitdet {

namespace {

constexpr int kQ = 128;            // queries per work item (UMMA M)
constexpr int kKV = 128;           // keys per tile (UMMA N of QK^T, K of PV)
constexpr int kHP = 64;            // head pitch in shared memory (one 128-byte swizzle row of bf16)
constexpr int kKSlots = 2, kVSlots = 3, kQSlots = 2;
constexpr int kThreads = 192;
constexpr int kProducerWarp = 4, kMmaWarp = 5;      // highest warp ids: favoured by the arbiter
constexpr int kQBytes = kQ * kHP * 2;         // 16 KiB
constexpr int kTileBytes = kKV * kHP * 2;     // 16 KiB: one K tile or one V tile = two TMA boxes of 64 rows
constexpr int kBoxBytes = 64 * kHP * 2;
constexpr int kTmemCols = 256;
constexpr uint32_t kColS = 0, kColP = 128, kColO = 192;   // S [0,128) f32 | P [128,192) bf16x2 | O [192,256) f32
constexpr float kRescaleThreshold = 8.f;      // log2 units: P stays <= 2^8 between rescales
constexpr int kRing = kKSlots + kVSlots + kQSlots;
constexpr int kNumBars = 2 * kRing + 4;
constexpr int kDefaultPoly = 0;

struct AttnSwArgs {
    __nv_bfloat16* ctx;
    int ldo;
    int T, H;
    int hp;              // elements per head in qkv / ctx (key_dim rounded up to 8)
    int k16;             // ceil(d / 16): K steps of QK^T, N / 16 of PV
    int nq;              // query tiles per (image, head)
    int n_items;         // B * H * nq
    float scale_log2;    // log2(e) / sqrt(key_dim)
};

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 2^t for a pair of scores on the FMA pipe (no SFU): t = n + f with n = round(t), f in [-0.5, 0.5];
// 2^f by a degree-3 minimax polynomial (relative error 7.5e-5, far below the bf16 rounding of P), 2^n by adding n to the
// exponent field.  r = t + 1.5 * 2^23 holds n in its low mantissa bits and (bits(r) << 23) is exactly n << 23 (the magic
// constant's low nine bits are zero), so the scaling is one integer multiply-add.  A softmax warp can issue one MUFU every
// ~16 cycles (profiles/r02_attention_analysis.md); these FMA-pipe instructions fill the issue slots in between.
__device__ __forceinline__ void exp2_pair_poly(uint64_t t2, float& e0, float& e1) {
    float t0, t1;
    f2_unpack(t2, t0, t1);
    t0 = fmaxf(t0, -126.f);             // 2^-126 stands in for 0 (masked keys, far-away scores): no exponent wrap-around
    t1 = fmaxf(t1, -126.f);
    const uint64_t t = f2_pack(t0, t1);
    const uint64_t magic = f2_pack(12582912.f, 12582912.f), neg_magic = f2_pack(-12582912.f, -12582912.f);
    const uint64_t r = f2_add(t, magic);                                   // n in the low mantissa bits
    const uint64_t f = f2_fma(f2_add(r, neg_magic), f2_pack(-1.f, -1.f), t);            // t - n
    uint64_t pl = f2_fma(f2_pack(0.055171460f, 0.055171460f), f, f2_pack(0.24261086f, 0.24261086f));
    pl = f2_fma(pl, f, f2_pack(0.69326097f, 0.69326097f));
    pl = f2_fma(pl, f, f2_pack(0.99992812f, 0.99992812f));
    float p0, p1, r0, r1;
    f2_unpack(pl, p0, p1);
    f2_unpack(r, r0, r1);
    e0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(r0) << 23));
    e1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(r1) << 23));
}


// exponentials of one 32-key chunk held in v[0..32): p = 2^(s*c - m) in place as packed bf16 pairs in pk[0..16), row sum
// accumulated in sum2.  P4 of every four score pairs take the polynomial path instead of the SFU.
template <int P4>
__device__ __forceinline__ void exp_chunk(const uint32_t (&v)[32], uint32_t (&pk)[16], uint64_t sc2, uint64_t nm2, uint64_t (&sum2)[2]) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        float t0, t1, t2, t3;
        const uint64_t ta = f2_fma(f2_pack(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sc2, nm2);
        const uint64_t tb = f2_fma(f2_pack(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])), sc2, nm2);
        float e0, e1, e2, e3;
        if (((i >> 1) & 3) < P4) exp2_pair_poly(ta, e0, e1);
        else { f2_unpack(ta, t0, t1); e0 = ex2f(t0); e1 = ex2f(t1); }      // ex2(-inf) = 0
        if ((((i >> 1) + 1) & 3) < P4) exp2_pair_poly(tb, e2, e3);
        else { f2_unpack(tb, t2, t3); e2 = ex2f(t2); e3 = ex2f(t3); }
        sum2[0] = f2_add(sum2[0], f2_pack(e0, e1));
        sum2[1] = f2_add(sum2[1], f2_pack(e2, e3));
        pk[i / 2] = pack_bf16x2(e0, e1);
        pk[i / 2 + 1] = pack_bf16x2(e2, e3);
    }
}

// masks the keys past the end of the image (chunk `c` of a tile with `valid` existing keys) and folds the chunk into mx
__device__ __forceinline__ void max_chunk(uint32_t (&v)[32], int c, int valid, float (&mx)[4]) {
    if (32 * c + 32 > valid) {          // warp-uniform: only the last chunk of an image's last tile
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (32 * c + i >= valid) v[i] = 0xff800000u;     // -inf
    }
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        mx[0] = fmaxf(mx[0], __uint_as_float(v[i]));     mx[1] = fmaxf(mx[1], __uint_as_float(v[i + 1]));
        mx[2] = fmaxf(mx[2], __uint_as_float(v[i + 2])); mx[3] = fmaxf(mx[3], __uint_as_float(v[i + 3]));
    }
}

template <int P4>
__global__ void __launch_bounds__(kThreads, 2)
attn_sw_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnSwArgs p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[kNumBars];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    pdl_launch_dependents();
    const int nkv = (p.T + kKV - 1) / kKV;
    const int my_items = (p.n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const int total = my_items * nkv;          // key tiles this CTA walks through, in order
    const int n_pv = 16 * p.k16;

    const uint32_t base = smem_u32(smem_raw);
    if ((base & 1023u) != 0u) __trap();
    const uint32_t sQ = base;                              // Q buffer b at + b * kQBytes
    const uint32_t sK = sQ + kQSlots * kQBytes;            // K slot s at + s * tile
    const uint32_t sV = sK + kKSlots * kTileBytes;         // V slot s at + s * tile
    const uint32_t bar_kfull = smem_u32(&bars[0]);
    const uint32_t bar_vfull = bar_kfull + 8 * kKSlots;
    const uint32_t bar_qfull = bar_vfull + 8 * kVSlots;
    const uint32_t bar_kempty = bar_kfull + 8 * kRing;
    const uint32_t bar_vempty = bar_kempty + 8 * kKSlots;
    const uint32_t bar_qempty = bar_vempty + 8 * kVSlots;
    const uint32_t bar_s_full = bar_kfull + 8 * (2 * kRing);     // QK^T(g) complete               (MMA commit)
    const uint32_t bar_s_free = bar_s_full + 8;                  // S(g) is in registers           (4 warps)
    const uint32_t bar_p_full = bar_s_full + 16;                 // P(g) (and rescaled O) in TMEM  (4 warps)
    const uint32_t bar_pv_done = bar_s_full + 24;                // PV(g) complete                 (MMA commit)

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2 * kRing; ++s) mbar_init(bar_kfull + 8 * s, 1);
        mbar_init(bar_s_full, 1);
        mbar_init(bar_s_free, 4);
        mbar_init(bar_p_full, 4);
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
            int ks = 0, vs = 0;
            uint32_t kphase = 0, vphase = 0;
            int it = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
                const int bh = item / p.nq, q0 = (item - bh * p.nq) * kQ;
                const int b = bh / p.H, h = bh - b * p.H;
                const int row_base = b * p.T;
                const int qb = it & 1;
                mbar_wait_relaxed(bar_qempty + 8 * qb, ((it >> 1) & 1) ^ 1u);
                mbar_arrive_expect_tx(bar_qfull + 8 * qb, kQBytes);
                tma_load_2d(sQ + qb * kQBytes, &tmQKV, bar_qfull + 8 * qb, h * p.hp, row_base + q0);
                tma_load_2d(sQ + qb * kQBytes + kBoxBytes, &tmQKV, bar_qfull + 8 * qb, h * p.hp, row_base + q0 + 64);
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
        }
    } else if (warp == kMmaWarp) {
        // ------------------------------ MMA issuer --------------------------------
        // The whole warp runs the loop (warp-uniform control flow); only the tcgen05 instructions are issued by one
        // elected lane.  Step g issues QK^T(g) and then PV(g-1): the first QK^T of the next work item goes out before
        // the last PV of the current one.
        const uint32_t idesc_qk = umma_idesc_bf16_f32(kQ, kKV);
        const uint32_t idesc_pv = umma_idesc_bf16_f32_bmn(kQ, n_pv);
        const uint32_t tS = tmem_base + kColS, tP = tmem_base + kColP, tO = tmem_base + kColO;
        const uint64_t dq0 = umma_desc_sw128_kmajor(sQ);
        const uint64_t dk0 = umma_desc_sw128_kmajor(sK), dv0 = umma_desc_sw128_kmajor(sV);
        constexpr uint32_t kSlotStep = kTileBytes >> 4, kQStep = kQBytes >> 4;     // descriptor address units (16 B)
        int ks = 0, vs = 0;
        uint32_t kphase = 0, vphase = 0;
        int jq = 0, itq = 0;        // (tile within item, item ordinal) of the next QK^T
        int jp = 0;                 // tile within item of the next PV
        // Step g issues QK^T(g) and then PV(g-2): the softmax warps release S(g-1) and publish P(g-2) at the same moment
        // (the end of their tile g-2), and the scores are needed first.
        for (int g = 0; g <= total + 1; ++g) {
            if (g < total) {
                const int qb = itq & 1;
                if (jq == 0) {
                    mbar_wait(bar_qfull + 8 * qb, (itq >> 1) & 1);
                    // Heads are stored hp (< 64) columns apart, so the 64-column TMA boxes also carry the first columns of
                    // the next head.  In QK^T only the columns below 16 * k16 take part: clearing Q's columns
                    // [hp, 16 * k16) makes their products vanish whatever K holds there; V's extra columns only produce
                    // columns of O that are never stored.  16-byte chunk c of row r sits at chunk c ^ (r & 7).
                    if (p.hp < 16 * p.k16) {
                        const int c_lo = p.hp >> 3, c_hi = 2 * p.k16;
                        const uint32_t q = sQ + qb * kQBytes;
                        for (int r = lane; r < kQ; r += 32)
                            for (int c = c_lo; c < c_hi; ++c)
                                asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(q + r * 128 + ((c ^ (r & 7)) << 4)), "r"(0u) : "memory");
                        fence_proxy_async_smem();
                        __syncwarp();
                    }
                }
                // S = Q K(g)^T; the softmax warps moved S(g-1) into registers before signalling s_free
                mbar_wait(bar_kfull + 8 * ks, kphase);
                if (g >= 1) mbar_wait(bar_s_free, (g - 1) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t dq = dq0 + static_cast<uint64_t>(qb * kQStep);
                    const uint64_t dk = dk0 + static_cast<uint64_t>(ks * kSlotStep);
                    umma_bf16_ss(tS, dq, dk, idesc_qk, 0u);
                    if (p.k16 > 1) umma_bf16_ss(tS, dq + 2u, dk + 2u, idesc_qk, 1u);
                    if (p.k16 > 2) umma_bf16_ss(tS, dq + 4u, dk + 4u, idesc_qk, 1u);
                    if (p.k16 > 3) umma_bf16_ss(tS, dq + 6u, dk + 6u, idesc_qk, 1u);
                    umma_commit(bar_kempty + 8 * ks);
                    if (jq == nkv - 1) umma_commit(bar_qempty + 8 * qb);      // last use of this item's Q
                    umma_commit(bar_s_full);
                }
                __syncwarp();
                if (++ks == kKSlots) { ks = 0; kphase ^= 1u; }
                if (++jq == nkv) { jq = 0; ++itq; }
            }
            if (g > 1) {
                // O (+)= P(g-2) V(g-2)
                mbar_wait(bar_vfull + 8 * vs, vphase);
                mbar_wait(bar_p_full, (g - 2) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t dv = dv0 + static_cast<uint64_t>(vs * kSlotStep);
                    // 16 keys per step: 8 packed columns of P, 16 rows (2048 B) of V
#pragma unroll
                    for (int k = 0; k < kKV / 16; ++k)
                        umma_bf16_ts(tO, tP + 8u * k, dv + static_cast<uint64_t>(128u * k), idesc_pv, (k != 0) ? 1u : (jp > 0 ? 1u : 0u));
                    umma_commit(bar_vempty + 8 * vs);
                    umma_commit(bar_pv_done);
                }
                __syncwarp();
                if (++vs == kVSlots) { vs = 0; vphase ^= 1u; }
                if (++jp == nkv) jp = 0;
            }
        }
    } else {
        // ------------------------------ softmax -----------------------------------
        const int quad = warp & 3;                          // TMEM lane quadrant of this warp
        const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
        const uint32_t tS = tmem_base + lane_off + kColS, tP = tmem_base + lane_off + kColP, tO = tmem_base + lane_off + kColO;
        float m_used = -INFINITY;      // running maximum in the scaled log2 domain
        float l = 0.f;                 // running sum of p
        int j = 0, item = blockIdx.x;
        uint32_t v[4][32];             // the scores of the tile being exponentiated; refilled chunk by chunk with the next tile's
        float alpha = 1.f;
        bool grow = false;
        auto valid_of = [&](int jj) { return min(kKV, p.T - jj * kKV); };
        // row statistics of the tile whose scores are in v (called once all its chunks went through max_chunk)
        auto stats = [&](const float (&mx)[4]) {
            const float m_new = fmaxf(m_used, fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * p.scale_log2);
            grow = __any_sync(0xffffffffu, m_new > m_used + kRescaleThreshold);   // true on the first tile of an item
            alpha = 1.f;
            if (grow) {
                alpha = ex2f(m_used - m_new);       // 0 on the first tile (m_used = -inf)
                m_used = m_new;
                l *= alpha;
            }
        };
        if (total > 0) {
            // ---- prologue: S(0) into registers, its row maximum ----
            const int valid = valid_of(0), nch = (valid + 31) >> 5;
            mbar_wait(bar_s_full, 0);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (c < nch) tmem_ld_32x32(tS + 32u * c, v[c]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_s_free);
            float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (c < nch) max_chunk(v[c], c, valid, mx);
            stats(mx);
        }
        for (int g = 0; g < total; ++g) {
            // here: v = S(g), (m_used, l, alpha, grow) already account for tile g's maximum
            const bool first = j == 0;
            const int nch = (valid_of(j) + 31) >> 5;                       // chunks of this tile that hold existing keys
            const bool have_next = g + 1 < total;
            const int jn = (j + 1 == nkv) ? 0 : j + 1;
            const int valid_n = valid_of(jn), nch_n = have_next ? (valid_n + 31) >> 5 : 0;
            const float neg_m = -m_used;
            const uint64_t sc2 = f2_pack(p.scale_log2, p.scale_log2), nm2 = f2_pack(neg_m, neg_m);
            uint64_t sum2[2] = {0ull, 0ull};
            float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t pk[16];
                if (c < nch) {
                    exp_chunk<P4>(v[c], pk, sc2, nm2, sum2);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) pk[i] = 0u;               // P = 0 for keys that do not exist
                }
                if (c == 0) {
                    // P and O must no longer be in use by PV(g-1): it was issued when this warp finished tile g-1 and has
                    // had the first chunk's exponentials to complete
                    if (g > 0) {
                        mbar_wait(bar_pv_done, (g - 1) & 1);
                        tc_fence_after();
                        if (grow && !first) {
#pragma unroll
                            for (int h2 = 0; h2 < 4; ++h2) {
                                if (16 * h2 >= n_pv) break;
                                uint32_t o[16];
                                tmem_ld_32x32_x16(tO + 16u * h2, o);
                                tmem_ld_wait();
#pragma unroll
                                for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                                tmem_st_32x32_x16(tO + 16u * h2, o);
                            }
                        }
                    }
                    if (have_next) { mbar_wait(bar_s_full, (g + 1) & 1); tc_fence_after(); }     // QK^T(g+1) ran during tile g-1's tail
                }
                tmem_st_32x32_x16(tP + 16u * c, pk);
                if (c > 0 && c - 1 < nch_n) {
                    tmem_ld_wait();                                        // chunk c-1 of S(g+1), loaded one chunk ago
                    max_chunk(v[c - 1], c - 1, valid_n, mx);
                }
                if (c < nch_n) tmem_ld_32x32(tS + 32u * c, v[c]);          // the registers of chunk c are free again
            }
            {
                float s0, s1, s2, s3;
                f2_unpack(sum2[0], s0, s1);
                f2_unpack(sum2[1], s2, s3);
                l += (s0 + s1) + (s2 + s3);
            }
            tmem_st_wait();
            if (nch_n > 0) tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(bar_p_full);                                   // P(g) (and the rescaled O) are in TMEM
                if (have_next) mbar_arrive(bar_s_free);                    // S(g+1) is in registers: QK^T(g+2) may overwrite it
            }
            if (3 < nch_n) max_chunk(v[3], 3, valid_n, mx);
            if (++j == nkv) {
                // ---- end of the work item: O / l -> bf16 context rows, statistics reset ----
                mbar_wait(bar_pv_done, g & 1);
                tc_fence_after();
                const int bh = item / p.nq, q0 = (item - bh * p.nq) * kQ;
                const int b = bh / p.H, h = bh - b * p.H;
                const int q = q0 + quad * 32 + lane;
                const float inv = 1.f / l;
                __nv_bfloat16* orow = p.ctx + static_cast<size_t>(b * p.T + (q < p.T ? q : 0)) * p.ldo + h * p.hp;
#pragma unroll
                for (int h2 = 0; h2 < 4; ++h2) {
                    if (16 * h2 >= p.hp) break;
                    uint32_t o[16];
                    tmem_ld_32x32_x16(tO + 16u * h2, o);
                    tmem_ld_wait();
                    if (q < p.T) {
#pragma unroll
                        for (int gq = 0; gq < 2; ++gq) {
                            if (16 * h2 + 8 * gq >= p.hp) break;          // the head holds hp columns
                            uint4 w;
                            w.x = pack_bf16x2(__uint_as_float(o[8 * gq + 0]) * inv, __uint_as_float(o[8 * gq + 1]) * inv);
                            w.y = pack_bf16x2(__uint_as_float(o[8 * gq + 2]) * inv, __uint_as_float(o[8 * gq + 3]) * inv);
                            w.z = pack_bf16x2(__uint_as_float(o[8 * gq + 4]) * inv, __uint_as_float(o[8 * gq + 5]) * inv);
                            w.w = pack_bf16x2(__uint_as_float(o[8 * gq + 6]) * inv, __uint_as_float(o[8 * gq + 7]) * inv);
                            *reinterpret_cast<uint4*>(orow + 16 * h2 + 8 * gq) = w;
                        }
                    }
                }
                tc_fence_before();         // the O reads above are ordered before this warp's next p_full arrive
                j = 0; item += gridDim.x;
                m_used = -INFINITY; l = 0.f;
            }
            if (have_next) stats(mx);      // tile g+1's maximum against the (possibly reset) running statistics
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

cudaError_t attn_sw_launch(const AttnPlan& plan, int num_sms, cudaStream_t stream) {
    const AttnDesc& d = plan.desc;
    AttnSwArgs a;
    a.ctx = static_cast<__nv_bfloat16*>(d.ctx);
    a.ldo = d.ldo;
    a.T = d.T;
    a.H = d.H;
    a.hp = d.hp;
    a.k16 = (d.d + 15) / 16;
    a.nq = (d.T + kQ - 1) / kQ;
    a.n_items = d.B * d.H * a.nq;
    a.scale_log2 = d.scale * 1.4426950408889634f;
    const size_t smem = static_cast<size_t>(kQSlots) * kQBytes + static_cast<size_t>(kKSlots + kVSlots) * kTileBytes;
    const int grid = a.n_items < 2 * num_sms ? a.n_items : 2 * num_sms;
    // share of the exponentials computed on the FMA pipe, in pairs per four pairs (VITDET_ATTN_POLY=0..4 overrides; A/B switch)
    static int p4 = -1;
    if (p4 < 0) { const char* e = getenv("VITDET_ATTN_POLY"); p4 = e ? atoi(e) : kDefaultPoly; if (p4 < 0 || p4 > 4) p4 = kDefaultPoly; }
#define VITDET_LAUNCH(P)                                                                                                  \
    {                                                                                                                     \
        cudaError_t e = ensure_max_dynamic_smem(reinterpret_cast<const void*>(attn_sw_kernel<P>), static_cast<int>(smem)); \
        if (e != cudaSuccess) return e;                                                                                   \
        return launch_kernel(attn_sw_kernel<P>, dim3(grid), dim3(kThreads), smem, stream, 1, plan.tmQKV, a);             \
    }
    switch (p4) {
        case 1: VITDET_LAUNCH(1)
        case 2: VITDET_LAUNCH(2)
        case 3: VITDET_LAUNCH(3)
        case 4: VITDET_LAUNCH(4)
        default: VITDET_LAUNCH(0)
    }
#undef VITDET_LAUNCH
}

}  // namespace vitdet
