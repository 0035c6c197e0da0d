// Box / slot arithmetic shared by the decode kernels (rowops.cu) and the evaluation metric (metric.cu):
// transform_predictions (det.py:619-645), the class-confidence rule (det.py:1366-1376) and iou_calculator
// (det.py:761-875).  Every product and sum that the reference evaluates as a separate float32 op uses the
// __f*_rn intrinsics, so the compiler cannot contract them into FMAs and the results are bit-identical to a
// float32 evaluation of the reference's statements.
#pragma once

#include <cuda_runtime.h>

#include "kernels.h"

namespace vitdet {

__device__ __forceinline__ float sigmoid_f32(float x) { return 1.f / (1.f + expf(-x)); }

// tf.clip_by_value(x, 0, 1) = minimum(maximum(x, 0), 1); TF's maximum/minimum propagate NaN.
__device__ __forceinline__ float clip01_nan(float x) { return (x != x) ? x : fminf(fmaxf(x, 0.f), 1.f); }

// transform_predictions of one slot (det.py:619-645): sigmoid on all six, clip the box part to [0, 1], scale.
__device__ __forceinline__ void transform_slot(const float (&l)[6], const DecodeParams& dp, float (&dec)[6]) {
    float s[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) s[j] = sigmoid_f32(l[j]);
#pragma unroll
    for (int j = 2; j < 6; ++j) s[j] = clip01_nan(s[j]);
    dec[0] = s[0];
    dec[1] = s[1] * static_cast<float>(dp.classes - 1);
    dec[2] = s[2] * dp.img_w;   // center_x   (det.py:637)
    dec[3] = s[3] * dp.img_h;   // center_y   (det.py:638)
    dec[4] = s[4] * dp.img_h;   // bbox_height(det.py:639)
    dec[5] = s[5] * dp.img_w;   // bbox_width (det.py:640)
}

// (0.5 - |c - round_half_even(c)|) / 0.5   (det.py:1366-1376, 2279)
__device__ __forceinline__ float class_confidence(float classification) {
    const float err = fabsf(__fsub_rn(classification, rintf(classification)));
    return __fdiv_rn(__fsub_rn(0.5f, err), 0.5f);
}

__device__ __forceinline__ void sort2(float& a, float& b) { const float lo = fminf(a, b), hi = fmaxf(a, b); a = lo; b = hi; }

__device__ __forceinline__ float middle_extent(float a, float b, float c, float d) {
    sort2(a, b); sort2(c, d); sort2(a, c); sort2(b, d); sort2(b, c);      // a <= b <= c <= d
    return __fsub_rn(c, b);                                                // sorted[-2] - sorted[-3]
}

// iou_calculator for one pair of (cx, cy, h, w) boxes: edges, strict '<' / '>' overlap test, the two middle values
// of the four sorted edges per axis, I / (U + eps).
__device__ __forceinline__ float iou_boxes(float lx, float ly, float lh, float lw, float px, float py, float ph, float pw, float eps) {
    const float l_left = __fsub_rn(lx, __fmul_rn(lw, 0.5f)), l_right = __fadd_rn(lx, __fmul_rn(lw, 0.5f));
    const float p_left = __fsub_rn(px, __fmul_rn(pw, 0.5f)), p_right = __fadd_rn(px, __fmul_rn(pw, 0.5f));
    const float l_top = __fsub_rn(ly, __fmul_rn(lh, 0.5f)), l_bottom = __fadd_rn(ly, __fmul_rn(lh, 0.5f));
    const float p_top = __fsub_rn(py, __fmul_rn(ph, 0.5f)), p_bottom = __fadd_rn(py, __fmul_rn(ph, 0.5f));
    const bool hit = (l_left < p_right) && (l_right > p_left) && (l_top < p_bottom) && (l_bottom > p_top);
    float inter = 0.f;
    if (hit) inter = __fmul_rn(middle_extent(l_top, l_bottom, p_top, p_bottom), middle_extent(l_left, l_right, p_left, p_right));
    const float uni = __fsub_rn(__fadd_rn(__fmul_rn(pw, ph), __fmul_rn(lw, lh)), inter);
    return __fdiv_rn(inter, __fadd_rn(uni, eps));
}

}  // namespace vitdet
