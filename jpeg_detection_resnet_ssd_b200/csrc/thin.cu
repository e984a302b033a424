// thin.cu - standalone entry points for the small public functions of the reference's
// bounding_box_utils / matching_utils modules, running on the device with the same device
// functions the decode / encode kernels use.
//
//   iou                     /root/reference/localisation_part/bounding_box_utils/bounding_box_utils.py:283-383
//   convert_coordinates     .../bounding_box_utils.py:24-87
//   match_bipartite_greedy  /root/reference/localisation_part/ssd_encoder_decoder/matching_utils.py:22-79
//   match_multi             .../matching_utils.py:81-116
#include "common.cuh"
#include "ctx.cuh"
#include <math.h>
#include <algorithm>

namespace ssdc {

__global__ void iou_kernel(const double* __restrict__ b1, long long m, const double* __restrict__ b2, long long n,
                           int coords, int mode, double d, int what, double* __restrict__ out) {
    const long long total = (mode == SSDC_IOU_OUTER) ? m * n : (m > n ? m : n);
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        long long i, j;
        if (mode == SSDC_IOU_OUTER) { i = e / n; j = e % n; }
        else { i = (m == 1) ? 0 : e; j = (n == 1) ? 0 : e; }      // numpy broadcasting of a single box
        double x0, y0, x1, y1;
        to_corners(b1 + 4 * i, coords, &x0, &y0, &x1, &y1);
        Box<double> a = make_box<double>(x0, y0, x1, y1, d);
        to_corners(b2 + 4 * j, coords, &x0, &y0, &x1, &y1);
        Box<double> b = make_box<double>(x0, y0, x1, y1, d);
        if (what == 0) {
            out[e] = iou_boxes<double>(a, b);
        } else {   // intersection_area: here the side lengths do get `+ d` (bounding_box_utils.py:212, :222)
            double sx = np_relu(np_min(a.x1, b.x1) - np_max(a.x0, b.x0) + d);
            double sy = np_relu(np_min(a.y1, b.y1) - np_max(a.y0, b.y0) + d);
            out[e] = sx * sy;
        }
    }
}

// convert_coordinates: arithmetic in the input dtype, result stored as float64
template <typename T>
__global__ void convert_kernel(const T* __restrict__ in, long long rows, int width, int start, int conv, double dd,
                               double* __restrict__ out) {
    const long long total = rows * width;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long r = e / width;
        const int c = (int)(e % width);
        const int k = c - start;
        if (k < 0 || k > 3) { out[e] = (double)in[e]; continue; }
        const T* t = in + r * width + start;
        const T d = (T)dd;
        T v;
        switch (conv) {
            case SSDC_CONV_MINMAX2CENTROIDS:
                v = (k == 0) ? (t[0] + t[1]) / T(2) : (k == 1) ? (t[2] + t[3]) / T(2) : (k == 2) ? (t[1] - t[0] + d) : (t[3] - t[2] + d);
                break;
            case SSDC_CONV_CENTROIDS2MINMAX:
                v = (k == 0) ? t[0] - t[2] / T(2) : (k == 1) ? t[0] + t[2] / T(2) : (k == 2) ? t[1] - t[3] / T(2) : t[1] + t[3] / T(2);
                break;
            case SSDC_CONV_CORNERS2CENTROIDS:
                v = (k == 0) ? (t[0] + t[2]) / T(2) : (k == 1) ? (t[1] + t[3]) / T(2) : (k == 2) ? (t[2] - t[0] + d) : (t[3] - t[1] + d);
                break;
            case SSDC_CONV_CENTROIDS2CORNERS:
                v = (k == 0) ? t[0] - t[2] / T(2) : (k == 1) ? t[1] - t[3] / T(2) : (k == 2) ? t[0] + t[2] / T(2) : t[1] + t[3] / T(2);
                break;
            default:   // minmax2corners / corners2minmax: swap the two middle coordinates
                v = (k == 1) ? t[2] : (k == 2) ? t[1] : t[k];
                break;
        }
        out[e] = (double)v;
    }
}

__device__ __forceinline__ bool better_w(double v, long long i, double bv, long long bi) {
    const bool vn = v != v, bn = bv != bv;
    if (vn != bn) return vn;
    if (!vn && v != bv) return v > bv;
    return i < bi;
}

// match_bipartite_greedy, literal: m rounds of (row argmax, argmax over rows, zero row and column)
// on a private copy `w` of the weight matrix.  One CTA; each warp owns rows r = warp, warp+nw, ...
__global__ void __launch_bounds__(1024)
bipartite_kernel(double* __restrict__ w, long long m, long long n, double* __restrict__ row_val,
                 long long* __restrict__ row_idx, long long* __restrict__ matches) {
    __shared__ long long s_g, s_a;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (long long r = threadIdx.x; r < m; r += blockDim.x) matches[r] = 0;
    for (long long round = 0; round < m; ++round) {
        for (long long r = warp; r < m; r += nw) {
            double bv = -INFINITY; long long bi = 0x7fffffffffffffffLL;
            for (long long a = lane; a < n; a += 32) {
                double v = w[r * n + a];
                if (better_w(v, a, bv, bi)) { bv = v; bi = a; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (better_w(ov, oi, bv, bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) { row_val[r] = bv; row_idx[r] = bi; }
        }
        __syncthreads();
        if (warp == 0) {
            double bv = -INFINITY; long long bg = 0x7fffffffffffffffLL;
            for (long long r = lane; r < m; r += 32) {
                double v = row_val[r];
                if (better_w(v, r, bv, bg)) { bv = v; bg = r; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                long long og = __shfl_xor_sync(0xffffffffu, bg, o);
                if (better_w(ov, og, bv, bg)) { bv = ov; bg = og; }
            }
            if (lane == 0) { s_g = bg; s_a = row_idx[bg]; matches[bg] = row_idx[bg]; }
        }
        __syncthreads();
        const long long gs = s_g, as = s_a;
        for (long long a = threadIdx.x; a < n; a += blockDim.x) w[gs * n + a] = 0.0;
        for (long long r = threadIdx.x; r < m; r += blockDim.x) w[r * n + as] = 0.0;
        __syncthreads();
    }
}

// match_multi: per column argmax over the rows; flags[a] = best >= threshold
__global__ void multi_kernel(const double* __restrict__ w, long long m, long long n, double thr,
                             long long* __restrict__ col_gt, unsigned char* __restrict__ flag) {
    for (long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x; a < n; a += (long long)gridDim.x * blockDim.x) {
        double bv = -INFINITY; long long bg = 0x7fffffffffffffffLL;
        for (long long r = 0; r < m; ++r) {
            double v = w[r * n + a];
            if (better_w(v, r, bv, bg)) { bv = v; bg = r; }
        }
        col_gt[a] = bg;
        flag[a] = (bv >= thr) ? 1 : 0;
    }
}

// ordered compaction of the flagged columns (np.nonzero keeps ascending order); one CTA
__global__ void __launch_bounds__(1024)
compact_kernel(const long long* __restrict__ col_gt, const unsigned char* __restrict__ flag, long long n,
               long long* __restrict__ out_gt, long long* __restrict__ out_anchor, long long* __restrict__ count) {
    __shared__ int warp_sums[32];
    __shared__ long long carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (long long base = 0; base < n; base += 1024) {
        const long long a = base + tid;
        const bool f = (a < n) && flag[a];
        const unsigned mk = __ballot_sync(0xffffffffu, f);
        if (lane == 0) warp_sums[warp] = __popc(mk);
        __syncthreads();
        if (warp == 0) {
            int v = warp_sums[lane], x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            warp_sums[lane] = x - v;        // exclusive
        }
        __syncthreads();
        const long long pos = carry + warp_sums[warp] + __popc(mk & ((1u << lane) - 1u));
        if (f) { out_gt[pos] = col_gt[a]; out_anchor[pos] = a; }
        __syncthreads();
        if (tid == 1023) carry = pos + (f ? 1 : 0);
        __syncthreads();
    }
    if (tid == 0) *count = carry;
}

}  // namespace ssdc

using namespace ssdc;

extern "C" {

static int iou_like(ssdc_ctx* ctx, const double* boxes1, int64_t m, const double* boxes2, int64_t n,
                    int coords, int mode, int border_pixels, int what, double* out) {
    if (!ctx || m < 0 || n < 0 || coords < 0 || coords > 2 || border_pixels < 0 || border_pixels > 2 ||
        (mode != SSDC_IOU_OUTER && mode != SSDC_IOU_ELEMENTWISE)) { set_error("ssdc_iou: bad argument"); return SSDC_ERR_ARG; }
    if (mode == SSDC_IOU_ELEMENTWISE && !(m == n || m == 1 || n == 1)) { set_error("ssdc_iou: element-wise shapes (%lld) and (%lld) do not broadcast", (long long)m, (long long)n); return SSDC_ERR_ARG; }
    const int64_t total = (mode == SSDC_IOU_OUTER) ? m * n : (m == 0 || n == 0 ? 0 : std::max(m, n));
    if (total == 0) return SSDC_OK;
    if (!boxes1 || !boxes2 || !out) { set_error("ssdc_iou: NULL buffer"); return SSDC_ERR_ARG; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevCtx& d = ctx->devs[0];
    SSDC_CUDA(cudaSetDevice(d.device));
    SSDC_TRY(d.t0buf.ensure((size_t)m * 4 * sizeof(double)));
    SSDC_TRY(d.t1buf.ensure((size_t)n * 4 * sizeof(double)));
    SSDC_TRY(d.t2buf.ensure((size_t)total * sizeof(double)));
    SSDC_CUDA(cudaMemcpyAsync(d.t0buf.p, boxes1, (size_t)m * 4 * sizeof(double), cudaMemcpyHostToDevice, d.stream));
    SSDC_CUDA(cudaMemcpyAsync(d.t1buf.p, boxes2, (size_t)n * 4 * sizeof(double), cudaMemcpyHostToDevice, d.stream));
    const double dd = (border_pixels == SSDC_BORDER_INCLUDE) ? 1.0 : (border_pixels == SSDC_BORDER_EXCLUDE ? -1.0 : 0.0);
    {
        LaunchScope ls(ctx, &d, SSDC_K_THIN);
        unsigned grid = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)d.sm_count * 16);
        iou_kernel<<<grid, 256, 0, d.stream>>>(d.t0buf.as<double>(), m, d.t1buf.as<double>(), n, coords, mode, dd, what, d.t2buf.as<double>());
        SSDC_TRY(check_launch("iou_kernel"));
    }
    SSDC_CUDA(cudaMemcpyAsync(out, d.t2buf.p, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost, d.stream));
    SSDC_CUDA(cudaStreamSynchronize(d.stream));
    return SSDC_OK;
}

int ssdc_iou(ssdc_ctx* ctx, const double* boxes1, int64_t m, const double* boxes2, int64_t n,
             int coords, int mode, int border_pixels, double* out) {
    return iou_like(ctx, boxes1, m, boxes2, n, coords, mode, border_pixels, 0, out);
}

int ssdc_intersection_area(ssdc_ctx* ctx, const double* boxes1, int64_t m, const double* boxes2, int64_t n,
                           int coords, int mode, int border_pixels, double* out) {
    return iou_like(ctx, boxes1, m, boxes2, n, coords, mode, border_pixels, 1, out);
}

int ssdc_convert_coordinates(ssdc_ctx* ctx, const void* in, int dtype, int64_t rows, int width,
                             int start, int conversion, int border_pixels, double* out) {
    if (!ctx || rows < 0 || width < 4 || start < 0 || start + 4 > width || conversion < 0 || conversion > 5 ||
        border_pixels < 0 || border_pixels > 2 || (dtype != SSDC_F32 && dtype != SSDC_F64)) {
        set_error("ssdc_convert_coordinates: bad argument"); return SSDC_ERR_ARG;
    }
    if (rows == 0) return SSDC_OK;
    if (!in || !out) { set_error("ssdc_convert_coordinates: NULL buffer"); return SSDC_ERR_ARG; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevCtx& d = ctx->devs[0];
    SSDC_CUDA(cudaSetDevice(d.device));
    const size_t elem = dtype == SSDC_F32 ? 4 : 8;
    const size_t total = (size_t)rows * width;
    SSDC_TRY(d.t0buf.ensure(total * elem));
    SSDC_TRY(d.t2buf.ensure(total * sizeof(double)));
    SSDC_CUDA(cudaMemcpyAsync(d.t0buf.p, in, total * elem, cudaMemcpyHostToDevice, d.stream));
    const double dd = (border_pixels == SSDC_BORDER_INCLUDE) ? 1.0 : (border_pixels == SSDC_BORDER_EXCLUDE ? -1.0 : 0.0);
    {
        LaunchScope ls(ctx, &d, SSDC_K_THIN);
        unsigned grid = (unsigned)std::min<size_t>((total + 255) / 256, (size_t)d.sm_count * 16);
        if (dtype == SSDC_F32) convert_kernel<float><<<grid, 256, 0, d.stream>>>(d.t0buf.as<float>(), rows, width, start, conversion, dd, d.t2buf.as<double>());
        else convert_kernel<double><<<grid, 256, 0, d.stream>>>(d.t0buf.as<double>(), rows, width, start, conversion, dd, d.t2buf.as<double>());
        SSDC_TRY(check_launch("convert_kernel"));
    }
    SSDC_CUDA(cudaMemcpyAsync(out, d.t2buf.p, total * sizeof(double), cudaMemcpyDeviceToHost, d.stream));
    SSDC_CUDA(cudaStreamSynchronize(d.stream));
    return SSDC_OK;
}

int ssdc_match_bipartite_greedy(ssdc_ctx* ctx, const double* weights, int64_t m, int64_t n, int64_t* out_matches) {
    if (!ctx || m < 0 || n < 0) { set_error("ssdc_match_bipartite_greedy: bad argument"); return SSDC_ERR_ARG; }
    if (m == 0) return SSDC_OK;
    if (n == 0 || !weights || !out_matches) { set_error("ssdc_match_bipartite_greedy: empty anchor axis / NULL buffer"); return SSDC_ERR_ARG; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevCtx& d = ctx->devs[0];
    SSDC_CUDA(cudaSetDevice(d.device));
    SSDC_TRY(d.t0buf.ensure((size_t)m * n * sizeof(double)));
    SSDC_TRY(d.t1buf.ensure((size_t)m * sizeof(double)));
    SSDC_TRY(d.t2buf.ensure((size_t)m * sizeof(long long)));
    SSDC_TRY(d.t3buf.ensure((size_t)m * sizeof(long long)));
    SSDC_CUDA(cudaMemcpyAsync(d.t0buf.p, weights, (size_t)m * n * sizeof(double), cudaMemcpyHostToDevice, d.stream));
    {
        LaunchScope ls(ctx, &d, SSDC_K_THIN);
        bipartite_kernel<<<1, 1024, 0, d.stream>>>(d.t0buf.as<double>(), m, n, d.t1buf.as<double>(), d.t2buf.as<long long>(), d.t3buf.as<long long>());
        SSDC_TRY(check_launch("bipartite_kernel"));
    }
    SSDC_CUDA(cudaMemcpyAsync(out_matches, d.t3buf.p, (size_t)m * sizeof(long long), cudaMemcpyDeviceToHost, d.stream));
    SSDC_CUDA(cudaStreamSynchronize(d.stream));
    return SSDC_OK;
}

int ssdc_match_multi(ssdc_ctx* ctx, const double* weights, int64_t m, int64_t n, double threshold,
                     int64_t* out_gt, int64_t* out_anchor, int64_t* n_matches) {
    if (!ctx || m <= 0 || n < 0 || !n_matches) { set_error("ssdc_match_multi: bad argument"); return SSDC_ERR_ARG; }
    *n_matches = 0;
    if (n == 0) return SSDC_OK;
    if (!weights || !out_gt || !out_anchor) { set_error("ssdc_match_multi: NULL buffer"); return SSDC_ERR_ARG; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevCtx& d = ctx->devs[0];
    SSDC_CUDA(cudaSetDevice(d.device));
    SSDC_TRY(d.t0buf.ensure((size_t)m * n * sizeof(double)));
    SSDC_TRY(d.t1buf.ensure((size_t)n * sizeof(long long) + (size_t)n));
    SSDC_TRY(d.t2buf.ensure((size_t)n * sizeof(long long)));
    SSDC_TRY(d.t3buf.ensure((size_t)(n + 1) * sizeof(long long)));
    SSDC_CUDA(cudaMemcpyAsync(d.t0buf.p, weights, (size_t)m * n * sizeof(double), cudaMemcpyHostToDevice, d.stream));
    long long* col_gt = d.t1buf.as<long long>();
    unsigned char* flag = reinterpret_cast<unsigned char*>(col_gt + n);
    long long* o_gt = d.t2buf.as<long long>();
    long long* o_an = d.t3buf.as<long long>();
    long long* cnt = o_an + n;
    {
        LaunchScope ls(ctx, &d, SSDC_K_THIN);
        unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)d.sm_count * 8);
        multi_kernel<<<grid, 256, 0, d.stream>>>(d.t0buf.as<double>(), m, n, threshold, col_gt, flag);
        SSDC_TRY(check_launch("multi_kernel"));
    }
    {
        LaunchScope ls(ctx, &d, SSDC_K_THIN);
        compact_kernel<<<1, 1024, 0, d.stream>>>(col_gt, flag, n, o_gt, o_an, cnt);
        SSDC_TRY(check_launch("compact_kernel"));
    }
    long long k = 0;
    SSDC_CUDA(cudaMemcpyAsync(&k, cnt, sizeof(long long), cudaMemcpyDeviceToHost, d.stream));
    SSDC_CUDA(cudaStreamSynchronize(d.stream));
    if (k > 0) {
        SSDC_CUDA(cudaMemcpyAsync(out_gt, o_gt, (size_t)k * sizeof(long long), cudaMemcpyDeviceToHost, d.stream));
        SSDC_CUDA(cudaMemcpyAsync(out_anchor, o_an, (size_t)k * sizeof(long long), cudaMemcpyDeviceToHost, d.stream));
        SSDC_CUDA(cudaStreamSynchronize(d.stream));
    }
    *n_matches = k;
    return SSDC_OK;
}

}  // extern "C"
