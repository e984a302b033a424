// voc.cu - VOC-style matching of detections to ground truth on the device (SURVEY section 8f, rank 1:
// the direct consumer of the decoder's output).
//
// Replaces the per-prediction Python loop of
//   Evaluator.match_predictions   /root/reference/localisation_part/eval_utils/average_precision_evaluator.py:570-777
//
// The reference sorts the predictions of a class by descending confidence and walks them one by one,
// keeping a per-image "already matched" table.  Only predictions of the same (class, image) interact, and
// only through "was this ground-truth box matched by an earlier prediction": a prediction is a true
// positive iff it is the FIRST one (in sorted order) whose best-overlapping box is that ground-truth box.
// That is an atomicMin over the sorted positions, so every prediction is handled by its own thread:
//   voc_key_kernel     : sort keys (confidence desc, original index asc = a stable sort)
//   voc_sort_kernel    : per class bitonic sort (shared memory, global memory for large classes)
//   voc_best_kernel    : per prediction: IoU with the image's boxes of that class (float64, the prediction's
//                        own area in float32 exactly like the reference's structured 'f4' array), first
//                        argmax, threshold; atomicMin of the sorted position on the matched box
//   voc_flag_kernel    : true / false positive flags, per class inclusive scans (cumulative TP / FP)
#include "common.cuh"
#include "ctx.cuh"
#include <math.h>
#include <algorithm>

namespace ssdc {

constexpr int VOC_SORT_SMEM_KEYS = 16384;      // 128 KB of Key64

__global__ void voc_key_kernel(const float* __restrict__ conf, const long long* __restrict__ cls_off, int n_classes,
                               Key64* __restrict__ keys) {
    const int c = blockIdx.y + 1;
    const long long p0 = cls_off[c], p1 = cls_off[c + 1];
    for (long long i = p0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p1; i += (long long)gridDim.x * blockDim.x)
        keys[i] = Key64::make(conf[i] + 0.0f, (uint32_t)(i - p0));      // (-0 sorts like +0, as in argsort(-conf))
}

// one CTA per class
__global__ void __launch_bounds__(1024)
voc_sort_kernel(Key64* __restrict__ keys, const long long* __restrict__ cls_off, Key64* __restrict__ scratch,
                const long long* __restrict__ scratch_off) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int c = blockIdx.x + 1;
    const long long p0 = cls_off[c];
    const int n = (int)(cls_off[c + 1] - p0);
    if (n <= 1) return;
    int N = 1;
    while (N < n) N <<= 1;
    Key64* s = (N <= VOC_SORT_SMEM_KEYS) ? reinterpret_cast<Key64*>(smem_raw) : scratch + scratch_off[c];
    Key64* g = keys + p0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) s[i] = (i < n) ? g[i] : Key64::lowest();
    __syncthreads();
    block_bitonic_sort(s, N, false);
    for (int i = threadIdx.x; i < n; i += blockDim.x) g[i] = s[i];
}

__device__ __forceinline__ bool better_ov(double v, int i, double bv, int bi) {
    const bool vn = v != v, bn = bv != bv;
    if (vn != bn) return vn;
    if (!vn && v != bv) return v > bv;
    return i < bi;
}

struct VocArgs {
    int n_classes, only_first, has_neutral;
    double thr, d;
};

// per sorted prediction: best ground-truth box of its class in its image
__global__ void voc_best_kernel(const Key64* __restrict__ keys, const long long* __restrict__ cls_off,
                                const int* __restrict__ pred_image, const float* __restrict__ pred_box,
                                const double* __restrict__ gt, const unsigned char* __restrict__ gt_neutral,
                                const long long* __restrict__ gt_img_off, VocArgs a,
                                int* __restrict__ match_gt /*global gt row or -1: fp, -2: neutral / ignored*/,
                                int* __restrict__ first_pos /*per gt row*/) {
    const int c = blockIdx.y + 1;
    const long long p0 = cls_off[c], p1 = cls_off[c + 1];
    const float df = (float)a.d;
    for (long long sp = p0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; sp < p1; sp += (long long)gridDim.x * blockDim.x) {
        if (a.only_first && sp != p0) { match_gt[sp] = -3; continue; }      // not evaluated (reference quirk)
        const long long src = p0 + keys[sp].anchor();
        const int img = pred_image[src];
        const float px0 = pred_box[4 * src], py0 = pred_box[4 * src + 1], px1 = pred_box[4 * src + 2], py1 = pred_box[4 * src + 3];
        // the prediction's area term is computed in float32 (the reference keeps the box in an 'f4' record)
        const double area_p = (double)((px1 - px0 + df) * (py1 - py0 + df));
        double bv = -INFINITY;
        int bi = 0x7fffffff, brow = -1, local = 0;
        for (long long r = gt_img_off[img]; r < gt_img_off[img + 1]; ++r) {
            const double* g = gt + 5 * r;
            if (g[0] != (double)c) continue;
            const double sx = np_relu(np_min(g[3], (double)px1) - np_max(g[1], (double)px0) + 0.0);
            const double sy = np_relu(np_min(g[4], (double)py1) - np_max(g[2], (double)py0) + 0.0);
            const double inter = sx * sy;
            const double area_g = (g[3] - g[1] + a.d) * (g[4] - g[2] + a.d);
            const double ov = inter / (area_g + area_p - inter);
            if (better_ov(ov, local, bv, bi)) { bv = ov; bi = local; brow = (int)r; }
            ++local;
        }
        int m;
        if (brow < 0) m = -1;                                   // no box of this class in the image: fp (:709-713)
        else if (bv < a.thr) m = -1;                            // :727-731
        else if (a.has_neutral && gt_neutral[brow]) m = -2;     // neutral box: neither tp nor fp
        else { m = brow; atomicMin(&first_pos[brow], (int)(sp - p0)); }
        match_gt[sp] = m;
    }
}

// flags + per class inclusive scans; one CTA per class
__global__ void __launch_bounds__(1024)
voc_flag_kernel(const long long* __restrict__ cls_off, const int* __restrict__ match_gt, const int* __restrict__ first_pos,
                const Key64* __restrict__ keys, int* __restrict__ order, int* __restrict__ tp, int* __restrict__ fp,
                int* __restrict__ ctp, int* __restrict__ cfp) {
    __shared__ int wsum_t[32], wsum_f[32];
    __shared__ int carry_t, carry_f;
    const int c = blockIdx.x + 1;
    const long long p0 = cls_off[c];
    const int n = (int)(cls_off[c + 1] - p0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { carry_t = 0; carry_f = 0; }
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        int t = 0, f = 0;
        if (i < n) {
            const int m = match_gt[p0 + i];
            if (m >= 0) { if (first_pos[m] == i) t = 1; else f = 1; }     // duplicate detection: fp (:753-757)
            else if (m == -1) f = 1;
            tp[p0 + i] = t; fp[p0 + i] = f;
            order[p0 + i] = (int)keys[p0 + i].anchor();
        }
        int xt = t, xf = f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int yt = __shfl_up_sync(0xffffffffu, xt, o), yf = __shfl_up_sync(0xffffffffu, xf, o);
            if (lane >= o) { xt += yt; xf += yf; }
        }
        if (lane == 31) { wsum_t[warp] = xt; wsum_f[warp] = xf; }
        __syncthreads();
        if (warp == 0) {
            int wt = wsum_t[lane], wf = wsum_f[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int yt = __shfl_up_sync(0xffffffffu, wt, o), yf = __shfl_up_sync(0xffffffffu, wf, o);
                if (lane >= o) { wt += yt; wf += yf; }
            }
            wsum_t[lane] = wt; wsum_f[lane] = wf;
        }
        __syncthreads();
        const int it = carry_t + (warp ? wsum_t[warp - 1] : 0) + xt;
        const int jf = carry_f + (warp ? wsum_f[warp - 1] : 0) + xf;
        if (i < n) { ctp[p0 + i] = it; cfp[p0 + i] = jf; }
        __syncthreads();
        if (tid == 1023) { carry_t = it; carry_f = jf; }
        __syncthreads();
    }
}

}  // namespace ssdc

using namespace ssdc;

extern "C" int ssdc_voc_match(ssdc_ctx* ctx, const int32_t* pred_image, const float* pred_conf, const float* pred_box,
                              const int64_t* pred_class_offsets, int n_classes,
                              const double* gt, const uint8_t* gt_neutral, const int64_t* gt_image_offsets,
                              int64_t n_images, double iou_threshold, int border_pixels, int only_first,
                              int32_t* out_order, int32_t* out_tp, int32_t* out_fp, int32_t* out_ctp, int32_t* out_cfp) {
    if (!ctx || !pred_class_offsets || !gt_image_offsets || n_classes < 1 || n_images < 0 ||
        border_pixels < 0 || border_pixels > 2) { set_error("ssdc_voc_match: bad argument"); return SSDC_ERR_ARG; }
    const int64_t P = pred_class_offsets[n_classes + 1];
    const int64_t G = gt_image_offsets[n_images];
    if (pred_class_offsets[1] != 0 || P < 0 || P > 0x7fffffff || G < 0 || G > 0x7fffffff) { set_error("ssdc_voc_match: bad offsets"); return SSDC_ERR_ARG; }
    if (P == 0) return SSDC_OK;
    if (!pred_image || !pred_conf || !pred_box || (G > 0 && !gt) || !out_order || !out_tp || !out_fp || !out_ctp || !out_cfp) {
        set_error("ssdc_voc_match: NULL buffer"); return SSDC_ERR_ARG;
    }
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevCtx& d = ctx->devs[0];
    SSDC_CUDA(cudaSetDevice(d.device));
    cudaStream_t st = d.stream;
    // device buffers (thin-op scratch): carve one arena
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    // global-memory sort scratch for classes above the shared-memory capacity
    std::vector<long long> soff(n_classes + 2, 0);
    long long stotal = 0;
    int max_c = 0;
    for (int c = 1; c <= n_classes; ++c) {
        const long long n = pred_class_offsets[c + 1] - pred_class_offsets[c];
        if (n < 0) { set_error("ssdc_voc_match: offsets must be non-decreasing"); return SSDC_ERR_ARG; }
        max_c = (int)std::max<long long>(max_c, n);
        soff[c] = stotal;
        if (n > VOC_SORT_SMEM_KEYS) { long long N = 1; while (N < n) N <<= 1; stotal += N; }
    }
    const size_t o_img = carve((size_t)P * 4), o_conf = carve((size_t)P * 4), o_box = carve((size_t)P * 16);
    const size_t o_coff = carve((size_t)(n_classes + 2) * 8), o_soff = carve((size_t)(n_classes + 2) * 8);
    const size_t o_gt = carve((size_t)std::max<int64_t>(G, 1) * 40), o_neu = carve((size_t)std::max<int64_t>(G, 1));
    const size_t o_goff = carve((size_t)(n_images + 1) * 8);
    const size_t o_keys = carve((size_t)P * 8), o_scr = carve((size_t)std::max<long long>(stotal, 1) * 8);
    const size_t o_match = carve((size_t)P * 4), o_first = carve((size_t)std::max<int64_t>(G, 1) * 4);
    const size_t o_order = carve((size_t)P * 4), o_tp = carve((size_t)P * 4), o_fp = carve((size_t)P * 4);
    const size_t o_ctp = carve((size_t)P * 4), o_cfp = carve((size_t)P * 4);
    SSDC_TRY(d.t0buf.ensure(off));
    char* base = d.t0buf.as<char>();
    SSDC_CUDA(cudaMemcpyAsync(base + o_img, pred_image, (size_t)P * 4, cudaMemcpyHostToDevice, st));
    SSDC_CUDA(cudaMemcpyAsync(base + o_conf, pred_conf, (size_t)P * 4, cudaMemcpyHostToDevice, st));
    SSDC_CUDA(cudaMemcpyAsync(base + o_box, pred_box, (size_t)P * 16, cudaMemcpyHostToDevice, st));
    SSDC_CUDA(cudaMemcpyAsync(base + o_coff, pred_class_offsets, (size_t)(n_classes + 2) * 8, cudaMemcpyHostToDevice, st));
    SSDC_CUDA(cudaMemcpyAsync(base + o_soff, soff.data(), (size_t)(n_classes + 2) * 8, cudaMemcpyHostToDevice, st));
    if (G > 0) SSDC_CUDA(cudaMemcpyAsync(base + o_gt, gt, (size_t)G * 40, cudaMemcpyHostToDevice, st));
    if (G > 0 && gt_neutral) SSDC_CUDA(cudaMemcpyAsync(base + o_neu, gt_neutral, (size_t)G, cudaMemcpyHostToDevice, st));
    SSDC_CUDA(cudaMemcpyAsync(base + o_goff, gt_image_offsets, (size_t)(n_images + 1) * 8, cudaMemcpyHostToDevice, st));
    SSDC_CUDA(cudaMemsetAsync(base + o_first, 0x7f, (size_t)std::max<int64_t>(G, 1) * 4, st));     // INT_MAX-ish
    // the host vectors above must outlive the async copies: synchronise before they go out of scope
    VocArgs a;
    a.n_classes = n_classes; a.only_first = only_first; a.has_neutral = gt_neutral ? 1 : 0;
    a.thr = iou_threshold;
    a.d = (border_pixels == SSDC_BORDER_INCLUDE) ? 1.0 : (border_pixels == SSDC_BORDER_EXCLUDE ? -1.0 : 0.0);
    const long long* coff = reinterpret_cast<const long long*>(base + o_coff);
    Key64* keys = reinterpret_cast<Key64*>(base + o_keys);
    const unsigned gx = (unsigned)std::min<int>((max_c + 255) / 256, 4 * d.sm_count);
    {
        LaunchScope ls(ctx, &d, SSDC_K_THIN);
        voc_key_kernel<<<dim3(std::max(gx, 1u), n_classes), 256, 0, st>>>(reinterpret_cast<const float*>(base + o_conf), coff, n_classes, keys);
        SSDC_TRY(check_launch("voc_key_kernel"));
    }
    {
        LaunchScope ls(ctx, &d, SSDC_K_THIN);
        const size_t smem = (size_t)std::min<long long>(VOC_SORT_SMEM_KEYS, [&] { long long N = 1; while (N < max_c) N <<= 1; return N; }()) * sizeof(Key64);
        SSDC_CUDA(cudaFuncSetAttribute(voc_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(VOC_SORT_SMEM_KEYS * sizeof(Key64))));
        voc_sort_kernel<<<n_classes, 1024, smem, st>>>(keys, coff, reinterpret_cast<Key64*>(base + o_scr), reinterpret_cast<const long long*>(base + o_soff));
        SSDC_TRY(check_launch("voc_sort_kernel"));
    }
    {
        LaunchScope ls(ctx, &d, SSDC_K_THIN);
        voc_best_kernel<<<dim3(std::max(gx, 1u), n_classes), 256, 0, st>>>(
            keys, coff, reinterpret_cast<const int*>(base + o_img), reinterpret_cast<const float*>(base + o_box),
            reinterpret_cast<const double*>(base + o_gt), reinterpret_cast<const unsigned char*>(base + o_neu),
            reinterpret_cast<const long long*>(base + o_goff), a, reinterpret_cast<int*>(base + o_match), reinterpret_cast<int*>(base + o_first));
        SSDC_TRY(check_launch("voc_best_kernel"));
    }
    {
        LaunchScope ls(ctx, &d, SSDC_K_THIN);
        voc_flag_kernel<<<n_classes, 1024, 0, st>>>(coff, reinterpret_cast<const int*>(base + o_match), reinterpret_cast<const int*>(base + o_first), keys,
                                                  reinterpret_cast<int*>(base + o_order), reinterpret_cast<int*>(base + o_tp), reinterpret_cast<int*>(base + o_fp),
                                                  reinterpret_cast<int*>(base + o_ctp), reinterpret_cast<int*>(base + o_cfp));
        SSDC_TRY(check_launch("voc_flag_kernel"));
    }
    SSDC_CUDA(cudaMemcpyAsync(out_order, base + o_order, (size_t)P * 4, cudaMemcpyDeviceToHost, st));
    SSDC_CUDA(cudaMemcpyAsync(out_tp, base + o_tp, (size_t)P * 4, cudaMemcpyDeviceToHost, st));
    SSDC_CUDA(cudaMemcpyAsync(out_fp, base + o_fp, (size_t)P * 4, cudaMemcpyDeviceToHost, st));
    SSDC_CUDA(cudaMemcpyAsync(out_ctp, base + o_ctp, (size_t)P * 4, cudaMemcpyDeviceToHost, st));
    SSDC_CUDA(cudaMemcpyAsync(out_cfp, base + o_cfp, (size_t)P * 4, cudaMemcpyDeviceToHost, st));
    SSDC_CUDA(cudaStreamSynchronize(st));
    return SSDC_OK;
}
