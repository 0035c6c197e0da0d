// Evaluation metric on the device: MeanAveragePrecision of the reference (det.py:1268-2060) — update_state
// (det.py:1310-1862), result (det.py:1865-2049), reset_state (det.py:2052-2060) — consuming the decode output
// where it already lies in HBM.
//
// The reference walks batch x 80 classes x labels in eager Python (5-8 s per 8 images).  Here:
//   * the three state tensors live on the device; latest_positive_bboxes / labels_quantity_per_image are RINGS per
//     class (head[c] = physical slot of the most recent related image) instead of being shifted by one slot per
//     related image; vitdet_map_state() un-rotates them into the reference's layout;
//   * classes are independent, so update is one CTA per class; inside a class only the ring position depends on the
//     order of the batch, and that is a suffix count of "related" flags: image i of the batch lands r_i slots
//     behind the new head, r_i = number of related images after i.  Images with r_i >= L would be shifted out
//     again by the end of the batch, so they are never computed;
//   * one warp evaluates one (image, class) pair — scenario b / c / d of det.py:1497-1839 — with the slots spread
//     over the lanes; the label loop of scenario d is sequential (each match removes a prediction) with a warp
//     max-reduction per label;
//   * result sorts each class's L*K (confidence, IoU) pairs once (stable, descending: the lower flat index first on
//     ties, as tf.argsort does) and ten threads walk the sorted list, one per IoU threshold.
// All float arithmetic the reference does in float32 is done here with the same operations in the same order
// (__f*_rn intrinsics: no FMA contraction), so state and APs are bit-identical to a float32 restatement
// (what the parity tests compare against).  These kernels move a few hundred KB; they are latency-bound, not roofline material.
#include <cstdint>
#include <cstring>
#include <new>

#include "../../include/vitdet_b200.h"
#include "boxops.cuh"
#include "common.cuh"
#include "devbuf.h"
#include "kernels.h"
#include "launch.h"

namespace vitdet {

constexpr float kNoBox = -8.f;       // the reference's filler for "no object here" (det.py:1468-1476, 1627-1629)
constexpr int kNoClass = -1;
constexpr int kUpdateThreads = 128;
constexpr int kUpdateWarps = kUpdateThreads / 32;
constexpr int kResultThreads = 128;
constexpr int kIouThresholds = 10;   // tf.linspace(0.5, 0.95, num=10), det.py:1876

struct IouThresholds { float v[kIouThresholds]; };

// ------------------------------------------------------------------------------------------------
// K-m1: per slot — transform_predictions (optional), the positive rule, class ids, showed_up_classes.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
map_prepare_kernel(const float* __restrict__ y_true, const float* __restrict__ y_pred, int R, DecodeParams dp,
                   float* __restrict__ pred, float* __restrict__ pcls, int* __restrict__ pcat, int* __restrict__ lcat,
                   uint8_t* __restrict__ showed) {
    pdl_launch_dependents();
    pdl_wait();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    float l[6], dec[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) l[j] = y_pred[static_cast<size_t>(r) * 6 + j];
    if (dp.apply_transform) {
        transform_slot(l, dp, dec);                                  // det.py:1341-1342
    } else {
#pragma unroll
        for (int j = 0; j < 6; ++j) dec[j] = l[j];
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) pred[static_cast<size_t>(r) * 6 + j] = dec[j];
    const bool positive = (dec[0] > dp.obj_thr) && (class_confidence(dec[1]) > dp.cls_thr);   // det.py:1381-1384, 1461-1464
    const float id = rintf(dec[1]);
    int cat = kNoClass;
    if (positive && id >= 0.f && id < static_cast<float>(dp.classes)) cat = static_cast<int>(id);
    pcat[r] = cat;
    pcls[r] = positive ? dec[1] : kNoBox;                            // positives_one_pred[..., 1], det.py:1468-1470
    if (cat != kNoClass) showed[cat] = 1;                            // det.py:1386-1420

    // label side: isclose(label_class, category) (det.py:1487-1488) can hold for the nearest integer only
    const float lab = y_true[static_cast<size_t>(r) * 6 + 1];
    const double a = static_cast<double>(lab), c = rint(a);
    int lc = kNoClass;
    if (c >= 0.0 && c < static_cast<double>(dp.classes) && fabs(a - c) <= 1e-8 + 1e-5 * fabs(c)) lc = static_cast<int>(c);
    lcat[r] = lc;
    if (lab >= 0.f) {                                                // det.py:1352, cast to int32 at det.py:1398
        const int sc = static_cast<int>(lab);
        if (sc < dp.classes) showed[sc] = 1;
    }
}

// ------------------------------------------------------------------------------------------------
// K-m2: one CTA per class.
// ------------------------------------------------------------------------------------------------
struct MapUpdateArgs {
    const float* y_true;     // [B, S, 6]
    const float* pred;       // [B, S, 6] decoded
    const float* pcls;       // [B, S]    class value of positives, -8 elsewhere
    const int* pcat;         // [B, S]    class id of positives, -1 elsewhere
    const int* lcat;         // [B, S]    class id of labels, -1 elsewhere
    uint8_t* related;        // [C, B]    scratch
    float* bboxes;           // [C, L, K, 2] ring
    float* labels;           // [C, L]       ring
    int* head;               // [C]
    int B, S, L, K;
    float eps;
};

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// Writes the K (confidence, IoU) rows and the label count of one related image of class c.
__device__ void map_one_image(const MapUpdateArgs& a, int i, int c, float* cur, float* tmp, int* lslot, int* order,
                              float* ent, float* __restrict__ dst, float* __restrict__ lab_dst) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int S = a.S, K = a.K;
    const float* yt = a.y_true + static_cast<size_t>(i) * S * 6;
    const float* pr = a.pred + static_cast<size_t>(i) * S * 6;
    const float* pcl = a.pcls + static_cast<size_t>(i) * S;
    const int* pc = a.pcat + static_cast<size_t>(i) * S;
    const int* lc = a.lcat + static_cast<size_t>(i) * S;

    int nlab = 0, npred = 0;
    for (int s0 = 0; s0 < S; s0 += 32) {
        const int s = s0 + lane;
        nlab += __popc(__ballot_sync(0xffffffffu, s < S && lc[s] == c));
        npred += __popc(__ballot_sync(0xffffffffu, s < S && pc[s] == c));
    }
    if (lane == 0) *lab_dst = static_cast<float>(nlab);             // det.py:1529-1543

    if (npred == 0) {                                                // scenario b (det.py:1551-1555)
        for (int k = lane; k < 2 * K; k += 32) dst[k] = 0.f;
        return;
    }
    if (nlab == 0) {                                                 // scenario c (det.py:1559-1616)
        int n = 0;
        for (int s0 = 0; s0 < S; s0 += 32) {
            const int s = s0 + lane;
            const bool p = s < S && pc[s] == c;
            const unsigned m = __ballot_sync(0xffffffffu, p);
            if (p) tmp[n + __popc(m & lt)] = class_confidence(pcl[s]);
            n += __popc(m);
        }
        __syncwarp();
        if (npred < K) {                                             // slot order, zero padded at the end
            for (int k = lane; k < K; k += 32) { dst[2 * k] = k < npred ? tmp[k] : 0.f; dst[2 * k + 1] = 0.f; }
        } else {                                                     // K largest, descending
            for (int t = lane; t < npred; t += 32) {
                const float v = tmp[t];
                int rank = 0;
                for (int u = 0; u < npred; ++u) rank += (tmp[u] > v) || (tmp[u] == v && u < t);
                if (rank < K) { dst[2 * rank] = v; dst[2 * rank + 1] = 0.f; }
            }
        }
        return;
    }

    // scenario d (det.py:1620-1839)
    for (int s = lane; s < S; s += 32) {                             // bboxes_iou_pred, det.py:1627-1629
        const bool act = pc[s] == c;
#pragma unroll
        for (int j = 0; j < 4; ++j) cur[4 * s + j] = act ? pr[6 * s + 2 + j] : kNoBox;
    }
    {   // labels of the class sorted by area, ascending, lower slot first on ties (det.py:1632-1649)
        int n = 0;
        for (int s0 = 0; s0 < S; s0 += 32) {
            const int s = s0 + lane;
            const bool p = s < S && lc[s] == c;
            const unsigned m = __ballot_sync(0xffffffffu, p);
            if (p) {
                const int pos = n + __popc(m & lt);
                lslot[pos] = s;
                tmp[pos] = __fmul_rn(yt[6 * s + 5], yt[6 * s + 4]);
            }
            n += __popc(m);
        }
        __syncwarp();
        for (int t = lane; t < nlab; t += 32) {
            const float v = tmp[t];
            int rank = 0;
            for (int u = 0; u < nlab; ++u) rank += (tmp[u] < v) || (tmp[u] == v && u < t);
            order[rank] = lslot[t];
        }
        __syncwarp();
    }
    int fresh = 0;                                                   // new_bboxes_quantity
    for (int j = 0; j < nlab; ++j) {                                 // det.py:1661
        const int ls = order[j];
        const float lx = yt[6 * ls + 2], ly = yt[6 * ls + 3], lh = yt[6 * ls + 4], lw = yt[6 * ls + 5];
        float best = -INFINITY;
        for (int s = lane; s < S; s += 32) {
            const float v = iou_boxes(lx, ly, lh, lw, cur[4 * s], cur[4 * s + 1], cur[4 * s + 2], cur[4 * s + 3], a.eps);
            tmp[s] = v;
            best = fmaxf(best, v);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, off));
        if (best > 0.5f) {                                           // det.py:1686
            const float tol = __fadd_rn(1e-8f, __fmul_rn(1e-5f, fabsf(best)));   // isclose(ious, max), det.py:1691-1693
            int first = 0x7fffffff;
            for (int s = lane; s < S; s += 32) {
                if (fabsf(__fsub_rn(tmp[s], best)) <= tol) {
                    first = min(first, s);
#pragma unroll
                    for (int q = 0; q < 4; ++q) cur[4 * s + q] = kNoBox;        // det.py:1748-1750
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, off));
            if (lane == 0) { ent[2 * fresh] = class_confidence(pcl[first]); ent[2 * fresh + 1] = best; }
            ++fresh;
        }
        __syncwarp();
        if (fresh == K) break;                                       // det.py:1754-1756
    }
    int appended = 0;
    {   // predictions of the class that hit no label (det.py:1765-1839)
        int nleft = 0;
        for (int s0 = 0; s0 < S; s0 += 32) {
            const int s = s0 + lane;
            const bool p = s < S && cur[4 * s] >= 0.f && cur[4 * s + 1] >= 0.f && cur[4 * s + 2] >= 0.f && cur[4 * s + 3] >= 0.f;
            const unsigned m = __ballot_sync(0xffffffffu, p);
            if (p) tmp[nleft + __popc(m & lt)] = class_confidence(pcl[s]);
            nleft += __popc(m);
        }
        __syncwarp();
        if (nleft > 0 && fresh < K) {
            const int room = K - fresh;
            if (fresh + nleft > K) {
                for (int t = lane; t < nleft; t += 32) {
                    const float v = tmp[t];
                    int rank = 0;
                    for (int u = 0; u < nleft; ++u) rank += (tmp[u] > v) || (tmp[u] == v && u < t);
                    if (rank < room) { ent[2 * (fresh + rank)] = v; ent[2 * (fresh + rank) + 1] = 0.f; }
                }
                appended = room;
            } else {
                for (int t = lane; t < nleft; t += 32) { ent[2 * (fresh + t)] = tmp[t]; ent[2 * (fresh + t) + 1] = 0.f; }
                appended = nleft;
            }
        }
    }
    __syncwarp();
    const int total = fresh + appended, pad = K - total;             // rolling concat keeps the last K rows
    for (int k = lane; k < K; k += 32) {
        dst[2 * k] = k < pad ? 0.f : ent[2 * (k - pad)];
        dst[2 * k + 1] = k < pad ? 0.f : ent[2 * (k - pad) + 1];
    }
}

__global__ void __launch_bounds__(kUpdateThreads)
map_update_kernel(MapUpdateArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_related_total;
    pdl_launch_dependents();
    pdl_wait();
    const int c = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int B = a.B, S = a.S, L = a.L, K = a.K;
    int* s_img = reinterpret_cast<int*>(smem_raw);                   // [L] image at ring offset r
    unsigned char* wbase = smem_raw + static_cast<size_t>(L) * 4 + static_cast<size_t>(warp) * (7 * S + 2 * K) * 4;
    float* cur = reinterpret_cast<float*>(wbase);                    // [S, 4]
    float* tmp = cur + 4 * S;                                        // [S]
    int* lslot = reinterpret_cast<int*>(tmp + S);                    // [S]
    int* order = lslot + S;                                          // [S]
    float* ent = reinterpret_cast<float*>(order + S);                // [K, 2]

    // pass 1: does image i carry class c in its labels or its positives (scenarios b, c, d)?
    uint8_t* rel = a.related + static_cast<size_t>(c) * B;
    for (int i = tid; i < B; i += kUpdateThreads) {
        const int* pc = a.pcat + static_cast<size_t>(i) * S;
        const int* lc = a.lcat + static_cast<size_t>(i) * S;
        bool f = false;
        for (int s = 0; s < S; ++s) f |= (pc[s] == c) | (lc[s] == c);
        rel[i] = f ? 1 : 0;
    }
    __syncthreads();
    // ring offsets: r_i = related images after i; the last related image of the batch becomes slot 0
    if (warp == 0) {
        int running = 0;
        for (int base = ((B - 1) / 32) * 32; base >= 0; base -= 32) {
            const int i = base + lane;
            const bool f = i < B && rel[i];
            const unsigned m = __ballot_sync(0xffffffffu, f);
            const int r = running + __popc(lane == 31 ? 0u : (m >> (lane + 1)));
            if (f && r < L) s_img[r] = i;
            running += __popc(m);
        }
        if (lane == 0) s_related_total = running;
    }
    __syncthreads();
    const int total = s_related_total;
    if (total == 0) return;                                          // scenario a for the whole batch
    const int old_head = a.head[c];
    const int new_head = ((old_head - total % L) + L) % L;
    const int n = min(total, L);
    for (int r = warp; r < n; r += kUpdateWarps) {
        const int phys = (new_head + r) % L;
        map_one_image(a, s_img[r], c, cur, tmp, lslot, order, ent,
                      a.bboxes + (static_cast<size_t>(c) * L + phys) * K * 2, a.labels + static_cast<size_t>(c) * L + phys);
    }
    __syncthreads();
    if (tid == 0) a.head[c] = new_head;
}

// ------------------------------------------------------------------------------------------------
// K-m3: AP of one class at the ten IoU thresholds (det.py:1883-2022).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kResultThreads)
map_class_ap_kernel(const float* __restrict__ bboxes, const float* __restrict__ labels, const int* __restrict__ head,
                    const uint8_t* __restrict__ showed, int C, int L, int K, IouThresholds thr, float* __restrict__ ap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pdl_launch_dependents();
    pdl_wait();
    const int c = blockIdx.x, tid = threadIdx.x, n = L * K;
    if (!showed[c]) {
        if (tid < kIouThresholds) ap[tid * C + c] = 0.f;
        return;
    }
    float* conf = reinterpret_cast<float*>(smem_raw);
    float* iou = conf + n;
    float* sconf = iou + n;
    float* siou = sconf + n;
    const int h = head[c];
    for (int e = tid; e < n; e += kResultThreads) {                  // reshape(-1, 2) of the logical layout, det.py:1900-1901
        const int j = e / K, k = e - j * K, phys = (h + j) % L;
        const float* p = bboxes + ((static_cast<size_t>(c) * L + phys) * K + k) * 2;
        conf[e] = p[0];
        iou[e] = p[1];
    }
    __syncthreads();
    for (int e = tid; e < n; e += kResultThreads) {                  // argsort DESCENDING, det.py:1907-1915
        const float v = conf[e];
        int rank = 0;
        for (int u = 0; u < n; ++u) rank += (conf[u] > v) || (conf[u] == v && u < e);
        sconf[rank] = v;
        siou[rank] = iou[e];
    }
    __syncthreads();
    if (tid >= kIouThresholds) return;
    const float t = thr.v[tid];
    float labels_quantity = 0.f;                                     // det.py:1955-1956
    for (int j = 0; j < L; ++j) labels_quantity = __fadd_rn(labels_quantity, labels[static_cast<size_t>(c) * L + (h + j) % L]);
    // recall_precisions (det.py:1886, 1920-1951) is only ever read as sum(rp[i] + rp[i+1]); keep its last element
    // (`last`, still changing), the one before it (`prev`, final) and the running sum of finished pairs.
    float tp = 0.f, fp = 0.f, prev = 0.f, last = 1.f, edges = 0.f;
    int len = 1;
    for (int e = 0; e < n; ++e) {
        if (!(sconf[e] > 0.f)) continue;                             // empty rows, det.py:1927
        if (siou[e] > t) {
            tp = __fadd_rn(tp, 1.f);
            const float precision = __fdiv_rn(tp, __fadd_rn(tp, fp));
            if (len >= 2) edges = __fadd_rn(edges, __fadd_rn(prev, last));
            prev = last;
            last = precision;
            ++len;
        } else {
            fp = __fadd_rn(fp, 1.f);
            last = __fdiv_rn(tp, __fadd_rn(tp, fp));
        }
    }
    if (len >= 2) edges = __fadd_rn(edges, __fadd_rn(prev, last));
    float area = 0.f;
    if (labels_quantity > 0.f && len >= 2)                           // det.py:1962-1998
        area = __fdiv_rn(__fmul_rn(edges, __fdiv_rn(1.f, labels_quantity)), 2.f);
    ap[tid * C + c] = area;
}

// K-m4: mean over the classes that showed up, then over the ten thresholds (det.py:2024-2049).
// res[0] = mAP, res[1..10] = AP per IoU threshold.
__global__ void __launch_bounds__(32)
map_mean_kernel(const float* __restrict__ ap, const uint8_t* __restrict__ showed, int C, float* __restrict__ res) {
    pdl_launch_dependents();
    pdl_wait();
    const int t = threadIdx.x;
    float mean = 0.f;
    if (t < kIouThresholds) {
        float s = 0.f;
        int n = 0;
        for (int c = 0; c < C; ++c)
            if (showed[c]) { s = __fadd_rn(s, ap[t * C + c]); ++n; }
        if (n) mean = __fdiv_rn(s, static_cast<float>(n));
        res[1 + t] = mean;
    }
    float total = 0.f;
    for (int q = 0; q < kIouThresholds; ++q) total = __fadd_rn(total, __shfl_sync(0xffffffffu, mean, q));
    if (t == 0) res[0] = __fdiv_rn(total, static_cast<float>(kIouThresholds));
}

// un-rotates the rings into the reference's layout (slot 0 = most recent related image)
__global__ void __launch_bounds__(256)
map_export_kernel(const float* __restrict__ bboxes, const float* __restrict__ labels, const int* __restrict__ head,
                  int C, int L, int K, float* __restrict__ bboxes_out, float* __restrict__ labels_out) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = 2 * K;
    const long long total = static_cast<long long>(C) * L * row;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int q = static_cast<int>(e % row);
        const long long cj = e / row;
        const int j = static_cast<int>(cj % L), c = static_cast<int>(cj / L);
        const int phys = (head[c] + j) % L;
        bboxes_out[e] = bboxes[(static_cast<size_t>(c) * L + phys) * row + q];
        if (q == 0) labels_out[cj] = labels[static_cast<size_t>(c) * L + phys];
    }
}

// tf.linspace(0.5, 0.95, num=10) in float32: first = start, last = stop, middle = start + delta * i.
static IouThresholds make_thresholds() {
    IouThresholds t;
    volatile float start = 0.5f, stop = 0.95f;
    volatile float delta = (stop - start) / 9.f;
    t.v[0] = start;
    for (int i = 1; i < kIouThresholds - 1; ++i) {
        volatile float step = delta * static_cast<float>(i);
        volatile float v = start + step;
        t.v[i] = v;
    }
    t.v[kIouThresholds - 1] = stop;
    return t;
}

}  // namespace vitdet

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using namespace vitdet;

struct vitdet_map {
    int device = 0;
    int C = 0, L = 0, K = 0;
    DevBuf bboxes, labels, head, showed;          // state
    DevBuf ap, res, exp_bboxes, exp_labels;       // result / export staging
    DevBuf pred, pcls, pcat, lcat, related;       // per-update scratch, grown on demand
    DevBuf in_true, in_pred;                      // update_host staging
    IouThresholds thr;
    uint64_t launches = 0;
};

static int map_check_device(const vitdet_map* m, const char* what) {
    if (!m) return fail(VITDET_E_INVALID, "%s: null metric handle", what);
    int cur = -1;
    CU_TRY(cudaGetDevice(&cur));
    if (cur != m->device)
        return fail(VITDET_E_INVALID, "%s: the metric lives on device %d but the current CUDA device is %d", what, m->device, cur);
    return 0;
}

extern "C" {

int vitdet_map_create(int classes, int latest_related_images, int bboxes_per_image, vitdet_map** out) {
    if (!out) return fail(VITDET_E_INVALID, "map_create: null out pointer");
    *out = nullptr;
    if (classes <= 1 || latest_related_images <= 0 || bboxes_per_image <= 0)
        return fail(VITDET_E_INVALID, "map_create: classes must be > 1, latest_related_images and bboxes_per_image > 0");
    const long long n = static_cast<long long>(latest_related_images) * bboxes_per_image;
    if (n * 16 > 200 * 1024)
        return fail(VITDET_E_SHAPE, "map_create: latest_related_images * bboxes_per_image = %lld does not fit the result kernel's shared memory (max 12800)", n);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(VITDET_E_NO_DEVICE, "no CUDA device: this library has no CPU fallback");
    }
    vitdet_map* m = new (std::nothrow) vitdet_map();
    if (!m) return fail(VITDET_E_INVALID, "map_create: out of host memory");
    m->C = classes; m->L = latest_related_images; m->K = bboxes_per_image;
    m->thr = make_thresholds();
    int rc = 0;
    do {
        cudaError_t e = cudaGetDevice(&m->device);
        if (e != cudaSuccess) { rc = fail(VITDET_E_CUDA, "cudaGetDevice failed: %s", cudaGetErrorString(e)); break; }
        const size_t C = classes, L = latest_related_images, K = bboxes_per_image;
        if ((rc = m->bboxes.ensure(C * L * K * 2 * 4)) || (rc = m->labels.ensure(C * L * 4)) || (rc = m->head.ensure(C * 4)) ||
            (rc = m->showed.ensure(C)) || (rc = m->ap.ensure(kIouThresholds * C * 4)) || (rc = m->res.ensure((1 + kIouThresholds) * 4)) ||
            (rc = m->exp_bboxes.ensure(C * L * K * 2 * 4)) || (rc = m->exp_labels.ensure(C * L * 4)))
            break;
        e = cudaMemset(m->bboxes.p, 0, m->bboxes.bytes);
        if (e == cudaSuccess) e = cudaMemset(m->labels.p, 0, m->labels.bytes);
        if (e == cudaSuccess) e = cudaMemset(m->head.p, 0, m->head.bytes);
        if (e == cudaSuccess) e = cudaMemset(m->showed.p, 0, m->showed.bytes);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) rc = fail(VITDET_E_CUDA, "map_create: clearing the state failed: %s", cudaGetErrorString(e));
    } while (0);
    if (rc) { delete m; return rc; }
    *out = m;
    return 0;
}

void vitdet_map_destroy(vitdet_map* m) { delete m; }

int vitdet_map_reset(vitdet_map* m, void* stream) {
    RC_TRY(map_check_device(m, "map_reset"));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CU_TRY(cudaMemsetAsync(m->bboxes.p, 0, m->bboxes.bytes, st));
    CU_TRY(cudaMemsetAsync(m->labels.p, 0, m->labels.bytes, st));
    CU_TRY(cudaMemsetAsync(m->head.p, 0, m->head.bytes, st));
    CU_TRY(cudaMemsetAsync(m->showed.p, 0, m->showed.bytes, st));
    // the host-array entry points run on the legacy stream: without this a reset issued on a non-blocking stream could
    // still be in flight when the next vitdet_map_update_host starts
    CU_TRY(cudaStreamSynchronize(st));
    return 0;
}

int vitdet_map_update(vitdet_map* m, const float* y_true_dev, const float* y_pred_dev, int batch, int slots,
                      const vitdet_decode_params* p, void* stream) {
    RC_TRY(map_check_device(m, "map_update"));
    if (!y_true_dev || !y_pred_dev || !p || batch < 0 || slots <= 0) return fail(VITDET_E_INVALID, "map_update: bad arguments");
    if (p->classes != m->C) return fail(VITDET_E_SHAPE, "map_update: params.classes = %d but the metric was created for %d", p->classes, m->C);
    if (batch == 0) return 0;
    const size_t smem = static_cast<size_t>(m->L) * 4 + static_cast<size_t>(kUpdateWarps) * (7 * static_cast<size_t>(slots) + 2 * static_cast<size_t>(m->K)) * 4;
    if (smem > 200 * 1024) return fail(VITDET_E_SHAPE, "map_update: slots = %d is too large for the update kernel's shared memory", slots);
    if (static_cast<long long>(batch) * slots > (1ll << 30)) return fail(VITDET_E_SHAPE, "map_update: batch * slots too large");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int R = batch * slots;
    const size_t r = static_cast<size_t>(R);
    RC_TRY(m->pred.ensure(r * 24)); RC_TRY(m->pcls.ensure(r * 4)); RC_TRY(m->pcat.ensure(r * 4)); RC_TRY(m->lcat.ensure(r * 4));
    RC_TRY(m->related.ensure(static_cast<size_t>(m->C) * batch));
    DecodeParams dp;
    dp.obj_thr = p->objectness_threshold; dp.cls_thr = p->classification_threshold; dp.strict = 1;
    dp.img_h = p->image_h; dp.img_w = p->image_w; dp.classes = p->classes;
    dp.apply_transform = p->use_transform_predictions ? 1 : 0;
    CU_TRY(launch_kernel(map_prepare_kernel, dim3((R + 255) / 256), dim3(256), 0, st, 1, y_true_dev, y_pred_dev, R, dp,
                         m->pred.as<float>(), m->pcls.as<float>(), m->pcat.as<int>(), m->lcat.as<int>(), m->showed.as<uint8_t>()));
    MapUpdateArgs a;
    a.y_true = y_true_dev; a.pred = m->pred.as<float>(); a.pcls = m->pcls.as<float>(); a.pcat = m->pcat.as<int>(); a.lcat = m->lcat.as<int>();
    a.related = m->related.as<uint8_t>(); a.bboxes = m->bboxes.as<float>(); a.labels = m->labels.as<float>(); a.head = m->head.as<int>();
    a.B = batch; a.S = slots; a.L = m->L; a.K = m->K;
    a.eps = 1e-8f;   // Constants.EPSILON, det.py:24
    if (smem > 48 * 1024) CU_TRY(ensure_max_dynamic_smem(reinterpret_cast<const void*>(map_update_kernel), static_cast<int>(smem)));
    CU_TRY(launch_kernel(map_update_kernel, dim3(m->C), dim3(kUpdateThreads), smem, st, 1, a));
    m->launches += 2;
    return 0;
}

int vitdet_map_update_host(vitdet_map* m, const float* y_true_host, const float* y_pred_host, int batch, int slots,
                           const vitdet_decode_params* p) {
    RC_TRY(map_check_device(m, "map_update_host"));
    if (!y_true_host || !y_pred_host || batch < 0 || slots <= 0) return fail(VITDET_E_INVALID, "map_update_host: bad arguments");
    if (batch == 0) return 0;
    const size_t nb = static_cast<size_t>(batch) * slots * 24;
    RC_TRY(m->in_true.ensure(nb)); RC_TRY(m->in_pred.ensure(nb));
    CU_TRY(cudaMemcpy(m->in_true.p, y_true_host, nb, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(m->in_pred.p, y_pred_host, nb, cudaMemcpyHostToDevice));
    RC_TRY(vitdet_map_update(m, m->in_true.as<float>(), m->in_pred.as<float>(), batch, slots, p, nullptr));
    CU_TRY(cudaStreamSynchronize(nullptr));
    return 0;
}

int vitdet_map_result(vitdet_map* m, float* mean_ap_host, float* per_iou_host, float* per_class_host, void* stream) {
    RC_TRY(map_check_device(m, "map_result"));
    if (!mean_ap_host) return fail(VITDET_E_INVALID, "map_result: null output pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t smem = static_cast<size_t>(m->L) * m->K * 16;
    if (smem > 48 * 1024) CU_TRY(ensure_max_dynamic_smem(reinterpret_cast<const void*>(map_class_ap_kernel), static_cast<int>(smem)));
    CU_TRY(launch_kernel(map_class_ap_kernel, dim3(m->C), dim3(kResultThreads), smem, st, 1, m->bboxes.as<float>(), m->labels.as<float>(),
                         m->head.as<int>(), m->showed.as<uint8_t>(), m->C, m->L, m->K, m->thr, m->ap.as<float>()));
    CU_TRY(launch_kernel(map_mean_kernel, dim3(1), dim3(32), 0, st, 1, m->ap.as<float>(), m->showed.as<uint8_t>(), m->C, m->res.as<float>()));
    m->launches += 2;
    float res[1 + kIouThresholds];
    CU_TRY(cudaMemcpyAsync(res, m->res.p, sizeof(res), cudaMemcpyDeviceToHost, st));
    if (per_class_host)
        CU_TRY(cudaMemcpyAsync(per_class_host, m->ap.p, static_cast<size_t>(kIouThresholds) * m->C * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    *mean_ap_host = res[0];
    if (per_iou_host) memcpy(per_iou_host, res + 1, kIouThresholds * 4);
    return 0;
}

int vitdet_map_state(vitdet_map* m, float* bboxes_host, float* labels_host, uint8_t* showed_host, void* stream) {
    RC_TRY(map_check_device(m, "map_state"));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t nb = static_cast<size_t>(m->C) * m->L * m->K * 2, nl = static_cast<size_t>(m->C) * m->L;
    if (bboxes_host || labels_host) {
        const int blocks = static_cast<int>((nb + 255) / 256 < 1184 ? (nb + 255) / 256 : 1184);
        CU_TRY(launch_kernel(map_export_kernel, dim3(blocks), dim3(256), 0, st, 1, m->bboxes.as<float>(), m->labels.as<float>(), m->head.as<int>(),
                             m->C, m->L, m->K, m->exp_bboxes.as<float>(), m->exp_labels.as<float>()));
        m->launches += 1;
        if (bboxes_host) CU_TRY(cudaMemcpyAsync(bboxes_host, m->exp_bboxes.p, nb * 4, cudaMemcpyDeviceToHost, st));
        if (labels_host) CU_TRY(cudaMemcpyAsync(labels_host, m->exp_labels.p, nl * 4, cudaMemcpyDeviceToHost, st));
    }
    if (showed_host) CU_TRY(cudaMemcpyAsync(showed_host, m->showed.p, m->C, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    return 0;
}

int vitdet_map_iou_thresholds(const vitdet_map* m, float* out10) {
    if (!m || !out10) return fail(VITDET_E_INVALID, "map_iou_thresholds: null pointer");
    memcpy(out10, m->thr.v, sizeof(m->thr.v));
    return 0;
}

uint64_t vitdet_map_launch_count(const vitdet_map* m) { return m ? m->launches : 0; }

}  // extern "C"
