// evalprep.cu - the glue between the decoder and the evaluator / the augmentation chain.
//
//   apply_inverse_transforms      /root/reference/localisation_part/data_generator/object_detection_2d_misc_utils.py:22-73
//     with the inverters the reference's transformations return:
//       Resize                    .../data_generator/object_detection_2d_geometric_ops.py:75-79   (scale, np.round to 0 decimals)
//       RandomPatch / CropPad ..  .../data_generator/object_detection_2d_patch_sampling_ops.py:316-320  (translate), :577, :730 (identity)
//   result lists of the Evaluator .../eval_utils/average_precision_evaluator.py:402-422  (round(conf, n), round(coord, 1))
//   BoxFilter.__call__            .../data_generator/object_detection_2d_image_boxes_validation_utils.py:174-232
//
// Everything here is a few arithmetic operations per box; the point is that the decoder's device-resident rows reach
// `ssdc_voc_match` as flat arrays without a Python loop over every detection.
#include "common.cuh"
#include "ctx.cuh"
#include <math.h>

namespace ssdc {

// One inverter of one image: kind 0 identity, 1 scale + round half to even to 0 decimals (Resize), 2 translate.
// `a_y` acts on the ymin / ymax columns, `a_x` on xmin / xmax.
struct InvStep { double kind, a_y, a_x; };

__device__ __forceinline__ void apply_steps(double& xmin, double& ymin, double& xmax, double& ymax,
                                            const InvStep* __restrict__ steps, long long s0, long long s1) {
    for (long long s = s0; s < s1; ++s) {
        const InvStep st = steps[s];
        if (st.kind == 1.0) {            // geometric_ops.py:77-78: np.round(labels[:, cols] * (orig / out), decimals=0)
            ymin = rint(ymin * st.a_y); ymax = rint(ymax * st.a_y);
            xmin = rint(xmin * st.a_x); xmax = rint(xmax * st.a_x);
        } else if (st.kind == 2.0) {     // patch_sampling_ops.py:318-319: labels[:, cols] += patch_ymin / patch_xmin
            ymin = ymin + st.a_y; ymax = ymax + st.a_y;
            xmin = xmin + st.a_x; xmax = xmax + st.a_x;
        }
    }
}

// rows (n, width) in place; image of a row by binary search in the row offsets
__global__ void inverse_rows_kernel(double* __restrict__ rows, long long n, int width, const long long* __restrict__ row_off, int B,
                                    const InvStep* __restrict__ steps, const long long* __restrict__ step_off,
                                    int cx0, int cy0, int cx1, int cy1) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lo = 0, hi = B;                                   // last image whose first row is <= i
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (row_off[mid] <= i) lo = mid; else hi = mid; }
    double* r = rows + i * width;
    double xmin = r[cx0], ymin = r[cy0], xmax = r[cx1], ymax = r[cy1];
    apply_steps(xmin, ymin, xmax, ymax, steps, step_off[lo], step_off[lo + 1]);
    r[cx0] = xmin; r[cy0] = ymin; r[cx1] = xmax; r[cy1] = ymax;
}

// np.float64.__round__(d): rint(x * 10^d) / 10^d (numpy rounds half to even on the scaled value)
__device__ __forceinline__ double round_dec(double x, double p10) { return rint(x * p10) / p10; }

// The padded (B, K, 6) rows of an image-sweep decode -> packed evaluation records in (image, row) order
__global__ void eval_rows_kernel(const double* __restrict__ pad_rows, const int* __restrict__ counts, const long long* __restrict__ row_off,
                                 int K, int b0, const InvStep* __restrict__ steps, const long long* __restrict__ step_off,
                                 double conf_p10, int* __restrict__ o_img, int* __restrict__ o_cls, float* __restrict__ o_conf,
                                 float* __restrict__ o_box) {
    const int b = blockIdx.x;
    const int n = counts[b];
    const long long base = row_off[b];
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
        const double* p = pad_rows + ((size_t)b * K + r) * 6;
        double xmin = p[2], ymin = p[3], xmax = p[4], ymax = p[5];
        if (steps) apply_steps(xmin, ymin, xmax, ymax, steps, step_off[b0 + b], step_off[b0 + b + 1]);
        const long long o = base + r;
        o_img[o] = b0 + b;
        o_cls[o] = (int)p[0];
        // average_precision_evaluator.py:411-418 (then stored in an 'f4' record, :668-675)
        o_conf[o] = (float)(conf_p10 > 0.0 ? round_dec(p[1], conf_p10) : p[1]);
        o_box[4 * o + 0] = (float)round_dec(xmin, 10.0);
        o_box[4 * o + 1] = (float)round_dec(ymin, 10.0);
        o_box[4 * o + 2] = (float)round_dec(xmax, 10.0);
        o_box[4 * o + 3] = (float)round_dec(ymax, 10.0);
    }
}

__global__ void eval_scan_kernel(const int* __restrict__ counts, int B, long long* __restrict__ row_off) {
    // (B is a batch size: a single thread's loop is a few microseconds and needs no scratch)
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        long long acc = 0;
        for (int b = 0; b < B; ++b) { row_off[b] = acc; acc += counts[b]; }
        row_off[B] = acc;
    }
}

// BoxFilter.__call__ (image_boxes_validation_utils.py:174-232): one thread per box, all checks in float64 in the
// reference's operation order.  boxes (n, 4) = xmin, ymin, xmax, ymax; per-box image size (the batch form filters the
// boxes of many images in one launch).
struct BoxFilterArgs {
    int check_degenerate, check_min_area, check_overlap, criterion;      // criterion 0 center_point, 1 iou, 2 area
    double lower, upper, min_area, d;
};
__global__ void box_filter_kernel(const double* __restrict__ boxes, const double* __restrict__ img_hw, long long n,
                                  BoxFilterArgs a, unsigned char* __restrict__ keep) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double xmin = boxes[4 * i], ymin = boxes[4 * i + 1], xmax = boxes[4 * i + 2], ymax = boxes[4 * i + 3];
    const double H = img_hw[2 * i], W = img_hw[2 * i + 1];
    bool ok = true;
    if (a.check_degenerate) ok = ok && (xmax > xmin) && (ymax > ymin);                        // :186-188
    if (a.check_min_area) ok = ok && ((xmax - xmin) * (ymax - ymin) >= a.min_area);          // :190-192
    if (a.check_overlap) {
        if (a.criterion == 1) {
            // :203-207: iou(image, boxes, coords='corners', mode='element-wise', border_pixels)
            const Box<double> img = make_box<double>(0.0, 0.0, W, H, a.d);
            const Box<double> bx = make_box<double>(xmin, ymin, xmax, ymax, a.d);
            const double v = iou_boxes<double>(img, bx);
            ok = ok && (v > a.lower) && (v <= a.upper);
        } else if (a.criterion == 2) {
            // :208-226
            const double area = (xmax - xmin + a.d) * (ymax - ymin + a.d);
            const double cy0 = fmin(fmax(ymin, 0.0), H - 1.0), cy1 = fmin(fmax(ymax, 0.0), H - 1.0);
            const double cx0 = fmin(fmax(xmin, 0.0), W - 1.0), cx1 = fmin(fmax(xmax, 0.0), W - 1.0);
            const double inter = (cx1 - cx0 + a.d) * (cy1 - cy0 + a.d);
            const bool lo = (a.lower == 0.0) ? (inter > a.lower * area) : (inter >= a.lower * area);
            ok = ok && lo && (inter <= a.upper * area);
        } else {
            // :227-231
            const double cy = (ymin + ymax) / 2.0, cx = (xmin + xmax) / 2.0;
            ok = ok && (cy >= 0.0) && (cy <= H - 1.0) && (cx >= 0.0) && (cx <= W - 1.0);
        }
    }
    keep[i] = ok ? 1 : 0;
}

}  // namespace ssdc

using namespace ssdc;

extern "C" int ssdc_inverse_transform_rows(ssdc_ctx* ctx, double* rows, int64_t n, int width, const int64_t* row_offsets, int64_t B,
                                           const double* steps, const int64_t* step_offsets,
                                           int xmin_col, int ymin_col, int xmax_col, int ymax_col) {
    if (!ctx || n < 0 || B < 0 || width < 4 || (n > 0 && (!rows || !row_offsets || !step_offsets)) ||
        xmin_col < 0 || xmin_col >= width || ymin_col < 0 || ymin_col >= width || xmax_col < 0 || xmax_col >= width ||
        ymax_col < 0 || ymax_col >= width || B > 0x7fffffff) {
        set_error("ssdc_inverse_transform_rows: bad argument");
        return SSDC_ERR_ARG;
    }
    if (n == 0 || B == 0) return SSDC_OK;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevCtx& d = ctx->devs[0];
    SSDC_CUDA(cudaSetDevice(d.device));
    const int64_t n_steps = step_offsets[B];
    const size_t b_rows = (size_t)n * width * sizeof(double), b_ro = (size_t)(B + 1) * sizeof(long long),
                 b_st = (size_t)(n_steps > 0 ? n_steps : 1) * sizeof(InvStep);
    SSDC_TRY(d.t0buf.ensure(b_rows));
    SSDC_TRY(d.t1buf.ensure(b_ro));
    SSDC_TRY(d.t2buf.ensure(b_st));
    SSDC_TRY(d.t3buf.ensure(b_ro));
    cudaStream_t st = d.stream;
    SSDC_CUDA(cudaMemcpyAsync(d.t0buf.p, rows, b_rows, cudaMemcpyHostToDevice, st));
    SSDC_CUDA(cudaMemcpyAsync(d.t1buf.p, row_offsets, b_ro, cudaMemcpyHostToDevice, st));
    if (n_steps > 0) SSDC_CUDA(cudaMemcpyAsync(d.t2buf.p, steps, (size_t)n_steps * sizeof(InvStep), cudaMemcpyHostToDevice, st));
    SSDC_CUDA(cudaMemcpyAsync(d.t3buf.p, step_offsets, b_ro, cudaMemcpyHostToDevice, st));
    {
        LaunchScope ls(ctx, &d, SSDC_K_THIN);
        inverse_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d.t0buf.as<double>(), n, width, d.t1buf.as<long long>(), (int)B,
                                                                         d.t2buf.as<InvStep>(), d.t3buf.as<long long>(),
                                                                         xmin_col, ymin_col, xmax_col, ymax_col);
        SSDC_TRY(check_launch("inverse_rows_kernel"));
    }
    SSDC_CUDA(cudaMemcpyAsync(rows, d.t0buf.p, b_rows, cudaMemcpyDeviceToHost, st));
    SSDC_CUDA(cudaStreamSynchronize(st));
    return SSDC_OK;
}

extern "C" int ssdc_results_for_evaluation(ssdc_ctx* ctx, const double* steps, const int64_t* step_offsets, int round_conf_decimals,
                                           int32_t* out_image, int32_t* out_class, float* out_conf, float* out_box,
                                           int64_t capacity, int64_t* n_out) {
    if (!ctx || !n_out || capacity < 0 || (capacity > 0 && (!out_image || !out_class || !out_conf || !out_box)) || (steps && !step_offsets)) {
        set_error("ssdc_results_for_evaluation: bad argument");
        return SSDC_ERR_ARG;
    }
    std::lock_guard<std::mutex> lk(ctx->mu);
    *n_out = 0;
    int64_t B_all = 0;
    for (DevCtx& d : ctx->devs) {
        if (!d.job.valid || (d.job.B > 0 && !d.job.padded)) {
            set_error("ssdc_results_for_evaluation: the last decode left no device-resident padded result (needs a finite top_k, "
                      "float32 input, per-class or layer mode)");
            return SSDC_ERR_STATE;
        }
        B_all += d.job.B;
    }
    const int64_t n_steps = steps ? step_offsets[B_all] : 0;
    const double conf_p10 = round_conf_decimals >= 0 ? pow(10.0, (double)round_conf_decimals) : 0.0;
    int64_t pos = 0;
    for (DevCtx& d : ctx->devs) {
        const int64_t B = d.job.B;
        if (B == 0) continue;
        SSDC_CUDA(cudaSetDevice(d.device));
        SSDC_TRY(d.wait_sweeps());
        cudaStream_t st = d.stream;
        const int K = d.job.p.top_k;
        const size_t cap_rows = (size_t)B * K;
        // scratch: row offsets | steps | step offsets | out arrays
        SSDC_TRY(d.row_offset.ensure((size_t)(B + 1) * sizeof(long long)));
        SSDC_TRY(d.t2buf.ensure((size_t)(n_steps > 0 ? n_steps : 1) * sizeof(InvStep)));
        SSDC_TRY(d.t3buf.ensure((size_t)(B_all + 1) * sizeof(long long)));
        SSDC_TRY(d.t0buf.ensure(cap_rows * (2 * sizeof(int) + 5 * sizeof(float))));
        if (steps) {
            if (n_steps > 0) SSDC_CUDA(cudaMemcpyAsync(d.t2buf.p, steps, (size_t)n_steps * sizeof(InvStep), cudaMemcpyHostToDevice, st));
            SSDC_CUDA(cudaMemcpyAsync(d.t3buf.p, step_offsets, (size_t)(B_all + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
        }
        int* o_img = d.t0buf.as<int>();
        int* o_cls = o_img + cap_rows;
        float* o_conf = reinterpret_cast<float*>(o_cls + cap_rows);
        float* o_box = o_conf + cap_rows;
        {
            LaunchScope ls(ctx, &d, SSDC_K_MERGE);
            eval_scan_kernel<<<1, 32, 0, st>>>(d.out_count.as<int>(), (int)B, d.row_offset.as<long long>());
            SSDC_TRY(check_launch("eval_scan_kernel"));
        }
        d.job.scan_pending = false;                         // (row_offset now holds the packed offsets of this job)
        {
            LaunchScope ls(ctx, &d, SSDC_K_MERGE);
            eval_rows_kernel<<<(unsigned)B, 64, 0, st>>>(d.pad_rows.as<double>(), d.out_count.as<int>(), d.row_offset.as<long long>(), K,
                                                         (int)d.job.b0, steps ? d.t2buf.as<InvStep>() : nullptr, d.t3buf.as<long long>(),
                                                         conf_p10, o_img, o_cls, o_conf, o_box);
            SSDC_TRY(check_launch("eval_rows_kernel"));
        }
        long long total = 0;
        SSDC_CUDA(cudaMemcpyAsync(&total, d.row_offset.as<long long>() + B, sizeof(long long), cudaMemcpyDeviceToHost, st));
        SSDC_CUDA(cudaStreamSynchronize(st));
        if (pos + total > capacity) {
            set_error("ssdc_results_for_evaluation: output arrays hold %lld records, more are needed", (long long)capacity);
            return SSDC_ERR_CAPACITY;
        }
        if (total > 0) {
            SSDC_CUDA(cudaMemcpyAsync(out_image + pos, o_img, (size_t)total * sizeof(int), cudaMemcpyDeviceToHost, st));
            SSDC_CUDA(cudaMemcpyAsync(out_class + pos, o_cls, (size_t)total * sizeof(int), cudaMemcpyDeviceToHost, st));
            SSDC_CUDA(cudaMemcpyAsync(out_conf + pos, o_conf, (size_t)total * sizeof(float), cudaMemcpyDeviceToHost, st));
            SSDC_CUDA(cudaMemcpyAsync(out_box + 4 * pos, o_box, (size_t)total * 4 * sizeof(float), cudaMemcpyDeviceToHost, st));
            SSDC_CUDA(cudaStreamSynchronize(st));
        }
        pos += total;
    }
    *n_out = pos;
    return SSDC_OK;
}

extern "C" int ssdc_box_filter(ssdc_ctx* ctx, const double* boxes, const double* image_hw, int64_t n,
                               int check_degenerate, int check_min_area, int check_overlap, int criterion,
                               double lower, double upper, double min_area, int border_pixels, uint8_t* out_keep) {
    if (!ctx || n < 0 || (n > 0 && (!boxes || !image_hw || !out_keep)) || criterion < 0 || criterion > 2 || border_pixels < 0 || border_pixels > 2) {
        set_error("ssdc_box_filter: bad argument");
        return SSDC_ERR_ARG;
    }
    if (n == 0) return SSDC_OK;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevCtx& d = ctx->devs[0];
    SSDC_CUDA(cudaSetDevice(d.device));
    cudaStream_t st = d.stream;
    SSDC_TRY(d.t0buf.ensure((size_t)n * 4 * sizeof(double)));
    SSDC_TRY(d.t1buf.ensure((size_t)n * 2 * sizeof(double)));
    SSDC_TRY(d.t2buf.ensure((size_t)n));
    SSDC_CUDA(cudaMemcpyAsync(d.t0buf.p, boxes, (size_t)n * 4 * sizeof(double), cudaMemcpyHostToDevice, st));
    SSDC_CUDA(cudaMemcpyAsync(d.t1buf.p, image_hw, (size_t)n * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
    BoxFilterArgs a;
    a.check_degenerate = check_degenerate; a.check_min_area = check_min_area; a.check_overlap = check_overlap; a.criterion = criterion;
    a.lower = lower; a.upper = upper; a.min_area = min_area;
    a.d = (border_pixels == SSDC_BORDER_INCLUDE) ? 1.0 : (border_pixels == SSDC_BORDER_EXCLUDE ? -1.0 : 0.0);
    {
        LaunchScope ls(ctx, &d, SSDC_K_THIN);
        box_filter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d.t0buf.as<double>(), d.t1buf.as<double>(), n, a, d.t2buf.as<unsigned char>());
        SSDC_TRY(check_launch("box_filter_kernel"));
    }
    SSDC_CUDA(cudaMemcpyAsync(out_keep, d.t2buf.p, (size_t)n, cudaMemcpyDeviceToHost, st));
    SSDC_CUDA(cudaStreamSynchronize(st));
    return SSDC_OK;
}
