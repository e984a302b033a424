// common.cuh - shared device helpers for the SSD box codec kernels (sm_100a).
//
// Everything parity-critical is compiled with --fmad=false (see build.py): the
// reference is numpy, which never fuses a multiply with an add, so neither may
// the kernels.  IoU and coordinate helpers follow
// /root/reference/localisation_part/bounding_box_utils/bounding_box_utils.py
// (cited per function).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ssdcodec.h"

namespace ssdc {

// ---------------------------------------------------------------------------
// Sort keys.  A candidate is identified by (score, anchor index); the reference
// picks `np.argmax(score)` = highest score, first (lowest anchor) on ties
// (ssd_output_decoder.py:85), so the canonical order is score descending, then
// anchor ascending.  Keys are built so that a plain *descending* comparison
// gives exactly that order.
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t ord32(float f) {
    uint32_t u = __builtin_bit_cast(uint32_t, f + 0.0f);   // -0.0 -> +0.0 (numpy compares them equal)
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float unord32(uint32_t u) {
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    return __builtin_bit_cast(float, u);
}
__host__ __device__ __forceinline__ uint64_t ord64(double f) {
    uint64_t u = __builtin_bit_cast(uint64_t, f + 0.0);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double unord64(uint64_t u) {
    u = (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
    return __builtin_bit_cast(double, u);
}

// float32 scores: one 64-bit word  [ord32(score) | ~anchor].
struct Key64 {
    uint64_t v;
    __host__ __device__ __forceinline__ static Key64 make(float score, uint32_t anchor) {
        Key64 k; k.v = ((uint64_t)ord32(score) << 32) | (uint64_t)(0xffffffffu - anchor); return k;
    }
    __host__ __device__ __forceinline__ static Key64 lowest() { Key64 k; k.v = 0; return k; }
    __host__ __device__ __forceinline__ uint32_t anchor() const { return 0xffffffffu - (uint32_t)v; }
    __host__ __device__ __forceinline__ double score() const { return (double)unord32((uint32_t)(v >> 32)); }
    __host__ __device__ __forceinline__ uint64_t score_bits() const { return v >> 32; }
};
// float64 scores (decode_detections_fast on y_encoded, round trip): two words.
struct Key128 {
    uint64_t hi, lo;
    __host__ __device__ __forceinline__ static Key128 make(double score, uint32_t anchor) {
        Key128 k; k.hi = ord64(score); k.lo = (uint64_t)(0xffffffffu - anchor); return k;
    }
    __host__ __device__ __forceinline__ static Key128 lowest() { Key128 k; k.hi = 0; k.lo = 0; return k; }
    __host__ __device__ __forceinline__ uint32_t anchor() const { return 0xffffffffu - (uint32_t)lo; }
    __host__ __device__ __forceinline__ double score() const { return unord64(hi); }
    __host__ __device__ __forceinline__ uint64_t score_bits() const { return hi; }
};

// `by_anchor`: order by anchor ascending only (decode_detections_fast without
// NMS keeps the boxes in anchor order, ssd_output_decoder.py:324-331).
__device__ __forceinline__ bool key_before(const Key64& a, const Key64& b, bool by_anchor) {
    if (by_anchor) return (uint32_t)a.v > (uint32_t)b.v;
    return a.v > b.v;
}
__device__ __forceinline__ bool key_before(const Key128& a, const Key128& b, bool by_anchor) {
    if (by_anchor) return a.lo > b.lo;
    return (a.hi > b.hi) || (a.hi == b.hi && a.lo > b.lo);
}

template <typename InT> struct KeyOf;
template <> struct KeyOf<float>  { typedef Key64  type; };
template <> struct KeyOf<double> { typedef Key128 type; };

__device__ __forceinline__ Key64 shfl_xor_key(const Key64& k, int m) {
    Key64 r; r.v = __shfl_xor_sync(0xffffffffu, k.v, m); return r;
}
__device__ __forceinline__ Key128 shfl_xor_key(const Key128& k, int m) {
    Key128 r; r.hi = __shfl_xor_sync(0xffffffffu, k.hi, m); r.lo = __shfl_xor_sync(0xffffffffu, k.lo, m); return r;
}

// In-register bitonic sort of 32 keys (one per lane) into canonical order
// (lane 0 first).
template <typename KeyT>
__device__ __forceinline__ KeyT warp_sort(KeyT k, bool by_anchor) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int j = size >> 1; j > 0; j >>= 1) {
            KeyT o = shfl_xor_key(k, j);
            bool first_half = (lane & j) == 0;            // this lane holds the lower index of the pair
            bool canonical = (lane & size) == 0;          // this sub-sequence is sorted in canonical order
            bool o_before = key_before(o, k, by_anchor);
            bool k_before = key_before(k, o, by_anchor);
            // lower index keeps the element that comes first (canonical) or last (reversed)
            bool take_other = (first_half == canonical) ? o_before : k_before;
            if (take_other) k = o;
        }
    }
    return k;
}

// Block-wide bitonic sort of N (power of two) keys at `s` (shared or global
// memory) into canonical order.  All threads of the block must call it.
template <typename KeyT>
__device__ __forceinline__ void block_bitonic_sort(KeyT* s, int N, bool by_anchor) {
    for (int k = 2; k <= N; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < (N >> 1); i += blockDim.x) {
                int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));
                int hi = lo | j;
                KeyT a = s[lo], b = s[hi];
                bool canonical = (lo & k) == 0;
                bool swap = canonical ? key_before(b, a, by_anchor) : key_before(a, b, by_anchor);
                if (swap) { s[lo] = b; s[hi] = a; }
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------
// numpy-faithful min / max / clamp (np.minimum / np.maximum propagate NaN).
// ---------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T np_min(T a, T b) { return (a < b || a != a) ? a : b; }
template <typename T> __device__ __forceinline__ T np_max(T a, T b) { return (a > b || a != a) ? a : b; }
template <typename T> __device__ __forceinline__ T np_relu(T x) { return (x > T(0)) ? x : ((x == x) ? T(0) : x); }

// A box in 'corners' order plus its area term of the union
// (bounding_box_utils.py:378-379: (xmax - xmin + d) * (ymax - ymin + d)).
template <typename T> struct Box {
    T x0, y0, x1, y1, area;
};
template <typename T>
__device__ __forceinline__ Box<T> make_box(T x0, T y0, T x1, T y1, T d) {
    Box<T> b; b.x0 = x0; b.y0 = y0; b.x1 = x1; b.y1 = y1;
    b.area = (x1 - x0 + d) * (y1 - y0 + d);
    return b;
}

// IoU exactly as bounding_box_utils.py:345 + :268-280 + :378-383 evaluate it:
// the intersection ignores `border_pixels` (d = 0 there), the union does not.
template <typename T>
__device__ __forceinline__ T iou_boxes(const Box<T>& a, const Box<T>& b) {
    T sx = np_relu(np_min(a.x1, b.x1) - np_max(a.x0, b.x0) + T(0));
    T sy = np_relu(np_min(a.y1, b.y1) - np_max(a.y0, b.y0) + T(0));
    T inter = sx * sy;
    T uni = a.area + b.area - inter;
    return inter / uni;
}

// Box given as 4 numbers in one of the three coordinate formats -> corner form, as `iou` does it
// (bounding_box_utils.py:334-336 -> :77-80 for centroids; index mapping :353-362 otherwise).
__device__ __forceinline__ void to_corners(const double* c4, int coords, double* x0, double* y0, double* x1, double* y1) {
    if (coords == SSDC_COORDS_CENTROIDS) {
        // bounding_box_utils.py:334-336 -> :77-80
        *x0 = c4[0] - c4[2] / 2.0; *y0 = c4[1] - c4[3] / 2.0;
        *x1 = c4[0] + c4[2] / 2.0; *y1 = c4[1] + c4[3] / 2.0;
    } else if (coords == SSDC_COORDS_MINMAX) {
        *x0 = c4[0]; *x1 = c4[1]; *y0 = c4[2]; *y1 = c4[3];
    } else {
        *x0 = c4[0]; *y0 = c4[1]; *x1 = c4[2]; *y1 = c4[3];
    }
}


// IoU of TensorFlow 1.x's NonMaxSuppression CPU kernel (float32), used by the
// Keras-layer contract (keras_layer_DecodeDetections.py:195-199).  Parity
// unpinned: TensorFlow is not available to execute.
__device__ __forceinline__ float iou_tf(const Box<float>& a, const Box<float>& b) {
    float ay0 = fminf(a.y0, a.y1), ax0 = fminf(a.x0, a.x1), ay1 = fmaxf(a.y0, a.y1), ax1 = fmaxf(a.x0, a.x1);
    float by0 = fminf(b.y0, b.y1), bx0 = fminf(b.x0, b.x1), by1 = fmaxf(b.y0, b.y1), bx1 = fmaxf(b.x0, b.x1);
    float area_a = (ay1 - ay0) * (ax1 - ax0);
    float area_b = (by1 - by0) * (bx1 - bx0);
    if (area_a <= 0.f || area_b <= 0.f) return 0.f;
    float iy0 = fmaxf(ay0, by0), ix0 = fmaxf(ax0, bx0), iy1 = fminf(ay1, by1), ix1 = fminf(ax1, bx1);
    float inter = fmaxf(iy1 - iy0, 0.f) * fmaxf(ix1 - ix0, 0.f);
    return inter / (area_a + area_b - inter);
}

// Correctly rounded float32 exp (float32(exp(float64))).  np.exp on float32 is
// SIMD dispatched, host dependent and up to 2 ulp off; the codec defines its
// result as the correctly rounded value ("exp").
__device__ __forceinline__ float exp_cr(float x) { return (float)exp((double)x); }
__device__ __forceinline__ double exp_cr(double x) { return exp(x); }

// Streaming 128-bit global load that does not allocate in L1.
__device__ __forceinline__ int4 ldg_stream(const int4* p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ int pow2_ceil(int n) {
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}

}  // namespace ssdc
