// encode.cu - ground truth -> training target encoding of the SSD box codec (sm_100a).
//
// Replaces the numpy body of
//   SSDInputEncoder.__call__        /root/reference/localisation_part/ssd_encoder_decoder/ssd_input_encoder.py:277-418
//   generate_encoding_template      .../ssd_input_encoder.py:550-611
//   match_bipartite_greedy          .../matching_utils.py:22-79
//   match_multi                     .../matching_utils.py:81-116
//   iou (outer product)             /root/reference/localisation_part/bounding_box_utils/bounding_box_utils.py:283-383
//
// Kernels:
//   anchor_prep_kernel : once per encoder: anchor corners + union area term, and the 12-double tail
//                        [offsets | anchor | variances] every unmatched row carries (literal formula,
//                        so 0/0 anchors still produce the reference's NaN).  The host then groups the
//                        anchors into shape classes (build_shape_classes) for the sparse path.
//   gt_prep_kernel     : per ground-truth row: normalise, convert to the target coordinate format,
//                        corner form for the IoU.
//   E3 template_tma_kernel : streams y_encoded out (TMA bulk stores of per-anchor-range tiles that are
//                        identical for every image) on a side stream, beside the matching.
//   Sparse path (anchor sets with <= 64 shape classes, <= 128 rows per image, positive thresholds):
//     seed_kernel      : per row: list level tau_r from the most promising shape class
//     E1 pair_kernel   : thread <-> anchor; (row, class) shape bounds, group bounding boxes, float32
//                        screens, exact float64 IoU for the rest; match_multi / neutral decision per
//                        anchor, candidate list per row
//     E2 greedy_kernel : CTA per image: the m rounds of match_bipartite_greedy literally (including the
//                        all-zero re-match quirk) on the lists; cooperative rescan when a list runs dry
//     E3' apply_list_kernel : patches the matched / neutral rows into the streamed template
//   General path (everything else; also the reference point the sparse path is tested against):
//     E1 rowbest_kernel: per (image, anchor chunk): best anchor of every GT row, np.argmax-faithful
//     E2 match_kernel  : one CTA per image: chunk partials + greedy rounds; a row whose best column was
//                        taken is rescanned in every chunk
//     E3 write_tma_kernel<false> : decide + assemble + one TMA bulk store per 128-row tile (no side stream)
//        write_tma_kernel<true>  : decide + patch after the template stream
//        write_kernel  : element-wise fallback for unaligned tiles
#include "common.cuh"
#include "ctx.cuh"
#include <math.h>
#include <algorithm>

struct ssdc_encoder {
    ssdc_ctx* ctx = nullptr;
    int64_t A = 0;
    ssdc_encode_params p;
    double variances[4];
    int64_t bad_image = -1;
    // per device of the context
    struct PerDev {
        ssdc::Buf anchor_box, anchor_tail, anchor_boxf;
        ssdc::Buf f_perm, f_box, f_boxf, f_grpcls, f_grpbox, f_cls, f_clsgoff;       // shape-class tables of the sparse path
    };
    std::vector<PerDev> dev;
    // sparse path (anchors fall into a few shape classes): see pairmatch_kernel
    bool fast_ok = false;
    int f_slots = 0, f_classes = 0;
};

namespace ssdc {

constexpr int E1_THREADS = 256;
constexpr int E1_PER_THREAD = 8;
constexpr int E1_CHUNK = E1_THREADS * E1_PER_THREAD;
constexpr int E3_THREADS = 128;
constexpr int E3_ROWS = 128;

struct GtPrep {
    double data[4];     // target coordinates in the encoder's `coords` format (written to y_encoded)
    Box<double> box;    // corner form + union area term, as `iou` sees it
    int cls;
    int regular;        // finite, positive area term
    float4 sf;          // float32 screening box (corners rounded outward; NaN when irregular)
    float w_up, h_up;   // upper bounds of the side lengths
    float area_lo;      // lower bound of the area term
    float pad;
};

struct EncArgs {
    int A, C, W, coords, background_id, multi, log_wh, chunks;
    double pos_thr, neg_thr, d;
    double img_h, img_w;
    int normalize;
};

// np.argmax-faithful "is (v, i) a better argmax than (bv, bi)": larger value wins, NaN beats
// everything, equal values (or two NaNs) resolve to the lower index.
__device__ __forceinline__ bool better(double v, int i, double bv, int bi) {
    const bool vn = v != v, bn = bv != bv;
    if (vn != bn) return vn;
    if (!vn && v != bv) return v > bv;
    return i < bi;
}

// target offsets of coordinates `t` relative to anchor `a` (ssd_input_encoder.py:396-410)
__device__ __forceinline__ void encode_offsets(const double* t, const double* a, const double* v,
                                               int coords, int log_wh, double* o) {
    if (coords == SSDC_COORDS_CENTROIDS) {
        o[0] = (t[0] - a[0]) / (a[2] * v[0]);
        o[1] = (t[1] - a[1]) / (a[3] * v[1]);
        double rw = t[2] / a[2], rh = t[3] / a[3];
        if (log_wh) {
            rw = (rw == 1.0) ? 0.0 : log(rw);
            rh = (rh == 1.0) ? 0.0 : log(rh);
        }
        o[2] = rw / v[2];
        o[3] = rh / v[3];
    } else if (coords == SSDC_COORDS_CORNERS) {
        double w = a[2] - a[0], h = a[3] - a[1];
        o[0] = (t[0] - a[0]) / w / v[0];
        o[1] = (t[1] - a[1]) / h / v[1];
        o[2] = (t[2] - a[2]) / w / v[2];
        o[3] = (t[3] - a[3]) / h / v[3];
    } else {
        double w = a[1] - a[0], h = a[3] - a[2];
        o[0] = (t[0] - a[0]) / w / v[0];
        o[1] = (t[1] - a[1]) / w / v[1];
        o[2] = (t[2] - a[2]) / h / v[2];
        o[3] = (t[3] - a[3]) / h / v[3];
    }
}

// ---------------------------------------------------------------------------
// Exact shortcuts for the ground-truth x anchor IoU (both boxes "regular": finite, area term > 0):
//   disjoint boxes                      => iou == +0 exactly;
//   iou <= min(a1, a2) / max(a1, a2)    (inter <= smaller area, union >= larger area), so a pair
//   whose area ratio is below a bound (with a 2^-40 guard for the roundings) is below that bound.
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool regular_d(const Box<double>& b) { return b.area > 0.0 && b.area < INFINITY; }
__device__ __forceinline__ bool disjoint_d(const Box<double>& a, const Box<double>& b) {
    return (a.x1 <= b.x0) || (b.x1 <= a.x0) || (a.y1 <= b.y0) || (b.y1 <= a.y0);
}
__device__ __forceinline__ bool ratio_below(const Box<double>& a, const Box<double>& b, double bound_lo) {
    const double mn = a.area < b.area ? a.area : b.area;
    const double mx = a.area < b.area ? b.area : a.area;
    return mn < bound_lo * mx;
}
constexpr double GUARD_LO = 1.0 - 0x1p-40;

// float32 screening box: the corners rounded OUTWARD, so that float-disjoint implies double-disjoint.
// An irregular box becomes NaN: every comparison fails and the pair takes the exact path.
__device__ __forceinline__ float4 screen_box(const Box<double>& b) {
    float4 f;
    if (!regular_d(b)) { f.x = f.y = f.z = f.w = __int_as_float(0x7fc00000); return f; }
    f.x = __double2float_rd(b.x0); f.y = __double2float_rd(b.y0);
    f.z = __double2float_ru(b.x1); f.w = __double2float_ru(b.y1);
    return f;
}
__device__ __forceinline__ bool screen_disjoint(const float4& a, const float4& b) {
    return (a.z <= b.x) || (b.z <= a.x) || (a.w <= b.y) || (b.w <= a.y);
}

__global__ void anchor_prep_kernel(const double* __restrict__ anchors, int A, int coords, double d, int log_wh,
                                   double v0, double v1, double v2, double v3,
                                   Box<double>* __restrict__ abox, double* __restrict__ tail, float4* __restrict__ aboxf) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= A) return;
    double c4[4] = {anchors[4 * a], anchors[4 * a + 1], anchors[4 * a + 2], anchors[4 * a + 3]};
    double x0, y0, x1, y1;
    to_corners(c4, coords, &x0, &y0, &x1, &y1);
    abox[a] = make_box<double>(x0, y0, x1, y1, d);
    aboxf[a] = screen_box(abox[a]);
    double v[4] = {v0, v1, v2, v3};
    double o[4];
    encode_offsets(c4, c4, v, coords, log_wh, o);      // an unmatched row encodes the anchor against itself
    double* t = tail + (size_t)a * 12;
    t[0] = o[0]; t[1] = o[1]; t[2] = o[2]; t[3] = o[3];
    t[4] = c4[0]; t[5] = c4[1]; t[6] = c4[2]; t[7] = c4[3];
    t[8] = v0; t[9] = v1; t[10] = v2; t[11] = v3;
}

// slot-ordered copies of the anchor boxes for the sparse path (padding slots: NaN screening box)
__global__ void permute_anchor_kernel(const int* __restrict__ perm, int S, const Box<double>* __restrict__ abox,
                                      const float4* __restrict__ aboxf, Box<double>* __restrict__ pbox, float4* __restrict__ pboxf) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const int a = perm[s];
    if (a >= 0) { pbox[s] = abox[a]; pboxf[s] = aboxf[a]; }
    else {
        Box<double> z; z.x0 = z.y0 = z.x1 = z.y1 = z.area = 0.0;
        pbox[s] = z;
        float4 n; n.x = n.y = n.z = n.w = __int_as_float(0x7fc00000);
        pboxf[s] = n;
    }
}

// One ground-truth row [class, xmin, ymin, xmax, ymax] -> normalised target-format box, corner form, screening data.
__device__ __forceinline__ GtPrep make_gtprep(const double* __restrict__ r, const EncArgs& g) {
    double xmin = r[1], ymin = r[2], xmax = r[3], ymax = r[4];
    if (g.normalize) {          // ssd_input_encoder.py:339-341
        ymin = ymin / g.img_h; ymax = ymax / g.img_h;
        xmin = xmin / g.img_w; xmax = xmax / g.img_w;
    }
    GtPrep p;
    if (g.coords == SSDC_COORDS_CENTROIDS) {          // :345 -> bounding_box_utils.py:72-75
        p.data[0] = (xmin + xmax) / 2.0;
        p.data[1] = (ymin + ymax) / 2.0;
        p.data[2] = xmax - xmin + g.d;
        p.data[3] = ymax - ymin + g.d;
    } else if (g.coords == SSDC_COORDS_MINMAX) {      // :347
        p.data[0] = xmin; p.data[1] = xmax; p.data[2] = ymin; p.data[3] = ymax;
    } else {
        p.data[0] = xmin; p.data[1] = ymin; p.data[2] = xmax; p.data[3] = ymax;
    }
    double x0, y0, x1, y1;
    to_corners(p.data, g.coords, &x0, &y0, &x1, &y1);
    p.box = make_box<double>(x0, y0, x1, y1, g.d);
    p.cls = (int)r[0];
    p.regular = regular_d(p.box) ? 1 : 0;
    p.sf = screen_box(p.box);
    p.w_up = __fsub_ru(p.sf.z, p.sf.x);
    p.h_up = __fsub_ru(p.sf.w, p.sf.y);
    p.area_lo = __double2float_rd(p.box.area);
    p.pad = 0.f;
    return p;
}

__global__ void gt_prep_kernel(const double* __restrict__ gt, int n, EncArgs g, GtPrep* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = make_gtprep(gt + (size_t)i * 5, g);
}

// ---------------------------------------------------------------------------
// E1: per GT row, best anchor inside one chunk of anchors
// ---------------------------------------------------------------------------

constexpr int E1_ROWS = 4;          // GT rows screened per barrier round

__global__ void __launch_bounds__(E1_THREADS, 4)
rowbest_kernel(const GtPrep* __restrict__ gtp, const long long* __restrict__ gt_off,
               const Box<double>* __restrict__ abox, const float4* __restrict__ aboxf, EncArgs g, int B, int chunk_hi,
               double* __restrict__ part_val, int* __restrict__ part_idx, int* __restrict__ irregular,
               unsigned long long* __restrict__ rowmax_bits) {
    // E1_ROWS GT rows per round: (A) every thread screens its anchors against the rows with cheap
    // float32 tests - disjointness of the outward-rounded boxes (then iou == +0 exactly) and the shape
    // bound against the best IoU already known for the row - and queues the surviving (row, anchor)
    // pairs in shared memory; (B) the survivors are evaluated exactly (float64) with all threads busy
    // and reduced per row with np.argmax semantics.  Two barriers per round.
    __shared__ unsigned short queue[E1_ROWS * E1_CHUNK];      // row << 13 | anchor inside the chunk
    __shared__ int s_qn[2];
    __shared__ double s_pv[E1_ROWS][E1_THREADS / 32];
    __shared__ int s_pi[E1_ROWS][E1_THREADS / 32];
    __shared__ int s_first;
    // chunk-major launch order, LAST chunk first: the large anchors of the late predictor layers give
    // large ground-truth boxes a high IoU early, which lets the small-anchor chunks skip almost all
    // of their pairs through the shape bound
    const int chunk = chunk_hi - (int)(blockIdx.x / B);
    const int b = blockIdx.x % B;
    const long long g0 = gt_off[b], g1 = gt_off[b + 1];
    const int m = (int)(g1 - g0);
    if (m == 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int a_base = chunk * E1_CHUNK;

    float4 af[E1_PER_THREAD];
    float aarea[E1_PER_THREAD];             // lower bound of the area term
    int first = 0x7fffffff;                 // lowest regular anchor index of this thread
    unsigned validm = 0;                    // which of this thread's anchors exist
#pragma unroll
    for (int k = 0; k < E1_PER_THREAD; ++k) {
        const int a = a_base + k * E1_THREADS + tid;
        if (a < g.A) validm |= 1u << k;
        aarea[k] = 0.f;
        af[k].x = af[k].y = af[k].z = af[k].w = __int_as_float(0x7fc00000);
        if (a < g.A) {
            af[k] = aboxf[a];
            if (af[k].x == af[k].x) {
                if (first == 0x7fffffff) first = a;
                aarea[k] = __double2float_rd(abox[a].area);
            }
        }
    }
    if (tid == 0) { s_first = 0x7fffffff; s_qn[0] = 0; s_qn[1] = 0; }
    __syncthreads();
    {
        const int wf = __reduce_min_sync(0xffffffffu, first);
        if (lane == 0 && wf != 0x7fffffff) atomicMin(&s_first, wf);
    }
    __syncthreads();
    const int first_reg = s_first;
    bool irr = false;
    int round = 0;
    for (int r0 = 0; r0 < m; r0 += E1_ROWS, ++round) {
        const int rn = min(E1_ROWS, m - r0);
        int* qn = &s_qn[round & 1];
        // ---- (A) screening
        unsigned keepm[E1_ROWS];
        int cnt = 0;
#pragma unroll
        for (int q = 0; q < E1_ROWS; ++q) {
            keepm[q] = 0;
            if (q < rn) {
                const GtPrep* gp = gtp + g0 + r0 + q;
                const float4 gf = gp->sf;
                const bool greg = gp->regular != 0;
                const double best = __longlong_as_double((long long)rowmax_bits[g0 + r0 + q]);
                const float cut = __fmul_rd(__double2float_rd(best), 1.0f - 0x1p-20f);
                const float gw = gp->w_up, gh = gp->h_up, garea = gp->area_lo;
                unsigned km = 0;
#pragma unroll
                for (int k = 0; k < E1_PER_THREAD; ++k) {
                    // (anchors past the end and irregular ones carry NaN screening boxes: never disjoint)
                    bool keep = !screen_disjoint(gf, af[k]);                      // disjoint: iou == +0 exactly
                    if (greg && cut > 0.f) {
                        // shape bound: inter <= min(w) * min(h) whatever the positions, so
                        // iou <= mwh / (a1 + a2 - mwh); below the row's known best => neither argmax nor tie
                        const float mwh = __fmul_ru(fminf(gw, __fsub_ru(af[k].z, af[k].x)), fminf(gh, __fsub_ru(af[k].w, af[k].y)));
                        const float den = __fsub_rd(__fadd_rd(garea, aarea[k]), mwh);
                        if (den > 0.f && mwh < __fmul_rd(cut, den) && af[k].x == af[k].x) keep = false;   // division-free
                    }
                    km |= (unsigned)keep << k;
                }
                keepm[q] = km & validm;
                cnt += __popc(keepm[q]);
            }
        }
        if (__any_sync(0xffffffffu, cnt != 0)) {
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            int base = 0;
            if (lane == 31) base = atomicAdd(qn, incl);
            base = __shfl_sync(0xffffffffu, base, 31) + incl - cnt;
#pragma unroll
            for (int q = 0; q < E1_ROWS; ++q)
                for (unsigned mm = keepm[q]; mm; mm &= mm - 1)
                    queue[base++] = (unsigned short)((q << 13) | ((__ffs(mm) - 1) * E1_THREADS + tid));
        }
        __syncthreads();
        // ---- (B) exact evaluation of the survivors
        const int nq = *qn;
        double bv[E1_ROWS];
        int bi[E1_ROWS];
#pragma unroll
        for (int q = 0; q < E1_ROWS; ++q) { bv[q] = -INFINITY; bi[q] = 0x7fffffff; }
        for (int i = tid; i < nq; i += E1_THREADS) {
            const unsigned e = queue[i];
            const int q = (int)(e >> 13);
            const int a = a_base + (int)(e & 0x1fffu);
            const GtPrep* gp = gtp + g0 + r0 + q;
            const Box<double> gb = gp->box;
            const Box<double> ab = abox[a];
            double s;
            if (gp->regular && regular_d(ab)) {
                if (disjoint_d(gb, ab)) s = 0.0;
                else s = iou_boxes<double>(gb, ab);
            } else {
                s = iou_boxes<double>(gb, ab);
                irr |= (s < 0.0);
            }
#pragma unroll
            for (int qq = 0; qq < E1_ROWS; ++qq)
                if (qq == q && better(s, a, bv[qq], bi[qq])) { bv[qq] = s; bi[qq] = a; }
        }
        if (nq > 0) {
#pragma unroll
            for (int q = 0; q < E1_ROWS; ++q) {
                if (__any_sync(0xffffffffu, bi[q] != 0x7fffffff)) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        double ov = __shfl_xor_sync(0xffffffffu, bv[q], o);
                        int oi = __shfl_xor_sync(0xffffffffu, bi[q], o);
                        if (better(ov, oi, bv[q], bi[q])) { bv[q] = ov; bi[q] = oi; }
                    }
                }
                if (lane == 0) { s_pv[q][warp] = bv[q]; s_pi[q][warp] = bi[q]; }
            }
        }
        __syncthreads();
        // ---- (C) combine with the +0 baseline of the screened-out (regular) pairs
        if (tid < rn) {
            const GtPrep* gp = gtp + g0 + r0 + tid;
            double rv = -INFINITY; int ri = 0x7fffffff;
            if (gp->regular && first_reg != 0x7fffffff) { rv = 0.0; ri = first_reg; }
            if (nq > 0) {
#pragma unroll
                for (int w = 0; w < E1_THREADS / 32; ++w)
                    if (better(s_pv[tid][w], s_pi[tid][w], rv, ri)) { rv = s_pv[tid][w]; ri = s_pi[tid][w]; }
            }
            part_val[(size_t)(g0 + r0 + tid) * g.chunks + chunk] = rv;
            part_idx[(size_t)(g0 + r0 + tid) * g.chunks + chunk] = ri;
            if (rv > 0.0) atomicMax(&rowmax_bits[g0 + r0 + tid], (unsigned long long)__double_as_longlong(rv));
        }
        if (tid == 0) *qn = 0;
    }
    if (__any_sync(0xffffffffu, irr) && lane == 0) atomicOr(&irregular[b], 1);
}

// ---------------------------------------------------------------------------
// E2: greedy bipartite rounds (matching_utils.py:63-77), one CTA per image
// ---------------------------------------------------------------------------
constexpr int E2_THREADS = 256;

// best column of ground-truth row `r` inside chunk `c`, with the taken columns zeroed
__device__ __forceinline__ void rescan_chunk(const Box<double>& gb, const Box<double>* __restrict__ abox, int A, int c,
                                             const unsigned* taken_bits, double* red_val, int* red_idx,
                                             double* out_val, int* out_idx) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool greg = regular_d(gb);
    double bv = -INFINITY; int bi = 0x7fffffff;
    for (int a = c * E1_CHUNK + tid; a < min(A, (c + 1) * E1_CHUNK); a += E2_THREADS) {
        double s;
        if ((taken_bits[a >> 5] >> (a & 31)) & 1u) s = 0.0;
        else {
            const Box<double> ab = abox[a];
            if (greg && regular_d(ab) && disjoint_d(gb, ab)) s = 0.0;
            else s = iou_boxes<double>(gb, ab);
        }
        if (better(s, a, bv, bi)) { bv = s; bi = a; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { red_val[warp] = bv; red_idx[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < E2_THREADS / 32; ++w)
            if (better(red_val[w], red_idx[w], bv, bi)) { bv = red_val[w]; bi = red_idx[w]; }
        *out_val = bv; *out_idx = bi;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(E2_THREADS)
match_kernel(const GtPrep* __restrict__ gtp, const long long* __restrict__ gt_off,
             const Box<double>* __restrict__ abox, EncArgs g,
             double* __restrict__ part_val, int* __restrict__ part_idx,
             const int* __restrict__ irregular, int* __restrict__ match) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double red_val[E2_THREADS / 32];
    __shared__ int red_idx[E2_THREADS / 32];
    __shared__ int s_asel;
    const int b = blockIdx.x;
    const long long g0 = gt_off[b];
    const int m = (int)(gt_off[b + 1] - g0);
    if (m == 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool irr = irregular[b] != 0;
    const int nbits = (g.A + 31) >> 5;
    // smem: row best value / index / done flag, bitmap of taken (zeroed) columns
    double* rb_val = reinterpret_cast<double*>(smem_raw);
    int* rb_idx = reinterpret_cast<int*>(rb_val + m);
    unsigned* taken_bits = reinterpret_cast<unsigned*>(rb_idx + m);
    unsigned char* done = reinterpret_cast<unsigned char*>(taken_bits + nbits);

    for (int i = tid; i < nbits; i += E2_THREADS) taken_bits[i] = 0;
    for (int r = tid; r < m; r += E2_THREADS) {
        double bv = -INFINITY; int bi = 0x7fffffff;
        for (int c = 0; c < g.chunks; ++c) {
            const double v = part_val[(size_t)(g0 + r) * g.chunks + c];
            const int i = part_idx[(size_t)(g0 + r) * g.chunks + c];
            if (better(v, i, bv, bi)) { bv = v; bi = i; }
        }
        rb_val[r] = bv; rb_idx[r] = bi; done[r] = 0;
        match[g0 + r] = 0;
    }
    __syncthreads();

    for (int round = 0; round < m; ++round) {
        if (warp == 0) {
            // ground_truth_index = np.argmax(overlaps): matched (zeroed) rows take part with weight 0
            double bv = -INFINITY; int bg = 0x7fffffff;
            for (int r = lane; r < m; r += 32) {
                const double v = rb_val[r];
                if (better(v, r, bv, bg)) { bv = v; bg = r; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                int og = __shfl_xor_sync(0xffffffffu, bg, o);
                if (better(ov, og, bv, bg)) { bv = ov; bg = og; }
            }
            if (lane == 0) {
                const int asel = rb_idx[bg];
                match[g0 + bg] = asel;
                rb_val[bg] = 0.0;          // weight_matrix[ground_truth_index] = 0 -> argmax 0, weight 0
                rb_idx[bg] = 0;
                done[bg] = 1;
                taken_bits[asel >> 5] |= 1u << (asel & 31);     // weight_matrix[:, anchor_index] = 0
                s_asel = asel;
            }
        }
        __syncthreads();
        const int asel = s_asel;
        // rows whose best column was just zeroed get a new best (everything, every round, for irregular inputs)
        for (int r = 0; r < m; ++r) {
            if (done[r]) continue;
            if (!irr && rb_idx[r] != asel) continue;
            const Box<double> gb = gtp[g0 + r].box;
            // every chunk: E1 pruned pairs against the row's best value of that time, so the stored partial
            // of a chunk is only guaranteed while that best stands - once it is gone, the runner-up may be a
            // pair E1 never evaluated
            for (int c = 0; c < g.chunks; ++c)
                rescan_chunk(gb, abox, g.A, c, taken_bits, red_val, red_idx,
                             &part_val[(size_t)(g0 + r) * g.chunks + c], &part_idx[(size_t)(g0 + r) * g.chunks + c]);
            if (tid == 0) {
                double bv = -INFINITY; int bi = 0x7fffffff;
                for (int c = 0; c < g.chunks; ++c) {
                    const double v = part_val[(size_t)(g0 + r) * g.chunks + c];
                    const int i = part_idx[(size_t)(g0 + r) * g.chunks + c];
                    if (better(v, i, bv, bi)) { bv = v; bi = i; }
                }
                rb_val[r] = bv; rb_idx[r] = bi;
            }
            __syncthreads();
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// E3: per-anchor matching + offsets + write-out
// ---------------------------------------------------------------------------
// Decision for one anchor: which ground-truth row it is matched to (-1: none) and whether it is
// neutral.  `sgt` / `smatch`: the image's GT boxes and bipartite matches in shared memory.
__device__ __forceinline__ void decide_anchor(int a, int m, const Box<double>* sgt, const float4* sgf, const int* smatch,
                                              bool all_gt_regular,
                                              const Box<double>* __restrict__ abox, const float4* __restrict__ aboxf,
                                              const EncArgs& g, bool prune, double prune_lo, int* gsel_out, bool* neutral_out) {
    int gsel = -1;
    bool neutral = false;
    if (m > 0) {
        // bipartite assignment: y_encoded[i, bipartite_matches, :-8] = labels_one_hot (last write wins)
        int bip = -1;
        for (int r = 0; r < m; ++r) if (smatch[r] == a) bip = r;
        // column of the similarity matrix; a bipartite column is all zeros (ssd_input_encoder.py:366)
        double cv = -INFINITY; int cg = 0x7fffffff;
        if (bip >= 0) { cv = 0.0; cg = 0; }
        else {
            const float4 af = aboxf[a];
            const bool areg = af.x == af.x;              // screen boxes of irregular anchors are NaN
            if (areg && all_gt_regular) {
                // regular pairs: iou >= 0, screened-out pairs exactly +0 => start at (0, row 0)
                cv = 0.0; cg = 0;
                Box<double> ab;
                bool have_ab = false;
                for (int r = 0; r < m; ++r) {
                    if (screen_disjoint(sgf[r], af)) continue;
                    if (!have_ab) { ab = abox[a]; have_ab = true; }
                    const Box<double> gb = sgt[r];
                    if (disjoint_d(gb, ab)) continue;
                    if (prune && ratio_below(gb, ab, prune_lo)) continue;        // below both thresholds
                    const double s = iou_boxes<double>(gb, ab);
                    if (better(s, r, cv, cg)) { cv = s; cg = r; }
                }
            } else {
                const Box<double> ab = abox[a];
                for (int r = 0; r < m; ++r) {
                    const double s = iou_boxes<double>(sgt[r], ab);
                    if (better(s, r, cv, cg)) { cv = s; cg = r; }
                }
            }
        }
        gsel = bip;
        bool zeroed = bip >= 0;
        if (g.multi && (cv >= g.pos_thr)) {      // matching_utils.py:109-114, ssd_input_encoder.py:375-381
            gsel = cg;
            zeroed = true;
        }
        const double rest = zeroed ? 0.0 : cv;   // np.amax of the column after the zeroing
        neutral = rest >= g.neg_thr;              // ssd_input_encoder.py:388-390
    }
    *gsel_out = gsel;
    *neutral_out = neutral;
}

struct RowMeta {
    double off[4];
    int cls;        // class column set to 1, or -1 for none (neutral background)
    int pad;
};

__global__ void __launch_bounds__(E3_THREADS)
write_kernel(const GtPrep* __restrict__ gtp, const long long* __restrict__ gt_off,
             const Box<double>* __restrict__ abox, const float4* __restrict__ aboxf, const double* __restrict__ tail,
             const int* __restrict__ match, EncArgs g, int tiles,
             double* __restrict__ y, double* __restrict__ y2, int* __restrict__ midx) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    RowMeta* meta = reinterpret_cast<RowMeta*>(smem_raw);               // E3_ROWS
    float4* sgf = reinterpret_cast<float4*>(meta + E3_ROWS);             // the image's GT boxes: screening form
    const int tile = blockIdx.x % tiles;
    const int b = blockIdx.x / tiles;
    const int a0 = tile * E3_ROWS;
    const int rows = min(E3_ROWS, g.A - a0);
    const long long g0 = gt_off[b];
    const int m = (int)(gt_off[b + 1] - g0);
    const int tid = threadIdx.x;

    Box<double>* sgt = reinterpret_cast<Box<double>*>(sgf + m);          // ... and exact form
    int* smatch = reinterpret_cast<int*>(sgt + m);
    bool greg_local = true;
    for (int r = tid; r < m; r += E3_THREADS) {
        const Box<double> gb = gtp[g0 + r].box;
        sgt[r] = gb; sgf[r] = screen_box(gb); smatch[r] = match[g0 + r];
        greg_local = greg_local && regular_d(gb);
    }
    const bool all_gt_regular = __syncthreads_and(greg_local) != 0;

    // pairs whose IoU is certainly below both thresholds never influence a decision (they can be
    // neither a match nor make the anchor neutral, nor be the argmax among pairs that do)
    const double thr_min = g.multi ? (g.pos_thr < g.neg_thr ? g.pos_thr : g.neg_thr) : g.neg_thr;
    const bool prune = thr_min > 0.0 && thr_min < INFINITY;
    const double prune_lo = thr_min * GUARD_LO;

    if (tid < rows) {
        const int a = a0 + tid;
        int gsel; bool neutral;
        decide_anchor(a, m, sgt, sgf, smatch, all_gt_regular, abox, aboxf, g, prune, prune_lo, &gsel, &neutral);
        RowMeta mt;
        const double* t = tail + (size_t)a * 12;
        if (gsel >= 0) {
            const GtPrep gp = gtp[g0 + gsel];
            double an[4] = {t[4], t[5], t[6], t[7]};
            double v[4] = {t[8], t[9], t[10], t[11]};
            encode_offsets(gp.data, an, v, g.coords, g.log_wh, mt.off);
            mt.cls = gp.cls;
            if (neutral && gp.cls == g.background_id) mt.cls = -1;
        } else {
            mt.off[0] = t[0]; mt.off[1] = t[1]; mt.off[2] = t[2]; mt.off[3] = t[3];
            mt.cls = neutral ? -1 : g.background_id;
        }
        mt.pad = 0;
        meta[tid] = mt;
        if (midx) midx[(size_t)b * g.A + a] = (gsel >= 0) ? gsel : (neutral ? -2 : -1);
    }
    __syncthreads();

    // stream the tile out: element e of the tile is (row e / W, column e % W); 16-byte stores when
    // the tile starts on a 16-byte boundary
    const int W = g.W, C = g.C;
    const int n = rows * W;
    const size_t base = ((size_t)b * g.A + a0) * W;
    double* out = y + base;
    double* out2 = y2 ? y2 + base : nullptr;
    const double* tl = tail + (size_t)a0 * 12;
    const float inv_w = 1.0f / (float)W;
    auto value = [&](int e) -> double {
        const int r = (int)(((float)e + 0.5f) * inv_w);
        const int c = e - r * W;
        if (c < C) return (c == meta[r].cls) ? 1.0 : 0.0;
        if (c < C + 4) return meta[r].off[c - C];
        return tl[(size_t)r * 12 + (c - C)];
    };
    auto value2 = [&](int e, double v) -> double {
        const int r = (int)(((float)e + 0.5f) * inv_w);
        const int c = e - r * W;
        return (c >= C && c < C + 4) ? 0.0 : v;
    };
    const bool vec = ((reinterpret_cast<uintptr_t>(out) & 15) == 0) && (!out2 || (reinterpret_cast<uintptr_t>(out2) & 15) == 0);
    if (vec) {
        const int npair = n >> 1;
        for (int i = tid; i < npair; i += E3_THREADS) {
            double2 v;
            v.x = value(2 * i);
            v.y = value(2 * i + 1);
            reinterpret_cast<double2*>(out)[i] = v;
            if (out2) {
                double2 w;
                w.x = value2(2 * i, v.x);
                w.y = value2(2 * i + 1, v.y);
                reinterpret_cast<double2*>(out2)[i] = w;
            }
        }
        if ((n & 1) && tid == 0) {
            const double v = value(n - 1);
            out[n - 1] = v;
            if (out2) out2[n - 1] = value2(n - 1, v);
        }
    } else {
        for (int e = tid; e < n; e += E3_THREADS) {
            const double v = value(e);
            out[e] = v;
            if (out2) out2[e] = value2(e, v);
        }
    }
}

// ---- E3, main path: rows are assembled in shared memory and the whole tile leaves with ONE TMA
// bulk store (cp.async.bulk shared -> global, SASS UBLKCP / UTMASTG class), no per-element index
// arithmetic.  Needs 16-byte aligned tiles (host checks; else write_kernel above).
__device__ __forceinline__ void tma_store_1d(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_commit_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ---- E3 template stream (overlapped path): the rows of unmatched anchors - all but a few hundred per
// image - depend on the anchor only, not on the image.  A CTA assembles the tile of one anchor range ONCE
// in shared memory and then issues one TMA bulk store of it per image: no per-element work at all, the
// copy engine streams `y_encoded` out while E1 / E2 (which this kernel does not depend on) run beside it
// on the main stream.  The matched / neutral rows are patched in afterwards (write_tma_kernel<true>).
constexpr int ET_ROWS = 64;
constexpr int ET_THREADS = 128;
__global__ void __launch_bounds__(ET_THREADS)
template_tma_kernel(const double* __restrict__ tail, EncArgs g, int tiles, int splits, int B,
                    double* __restrict__ y, double* __restrict__ y2) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int W = g.W, C = g.C;
    double* trow = reinterpret_cast<double*>(smem_raw);                 // ET_ROWS x W
    double* trow2 = trow + (size_t)ET_ROWS * W;                          // diagnostics copy (only with y2)
    const int tile = blockIdx.x % tiles, split = blockIdx.x / tiles;
    const int a0 = tile * ET_ROWS;
    const int rows = min(ET_ROWS, g.A - a0);
    const int b_lo = (int)((long long)B * split / splits), b_hi = (int)((long long)B * (split + 1) / splits);
    const int n = rows * W;
    for (int e = threadIdx.x; e < n; e += ET_THREADS) {
        const int r = e / W, c = e - r * W;
        const double v = (c < C) ? ((c == g.background_id) ? 1.0 : 0.0) : tail[(size_t)(a0 + r) * 12 + (c - C)];
        trow[e] = v;
        if (y2) trow2[e] = (c >= C && c < C + 4) ? 0.0 : v;             // ssd_input_encoder.py:412-416
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)((size_t)n * sizeof(double));
        // (measured and dropped: the L2 evict_first hint on these stores - B = 1024: 0.446 -> 0.463 ms per batch)
        for (int b = b_lo; b < b_hi; ++b) {
            const size_t base = ((size_t)b * g.A + a0) * W;
            tma_store_1d(y + base, trow, bytes);
            if (y2) tma_store_1d(y2 + base, trow2, bytes);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

// PATCH = false: the whole tile (every row) is assembled and stored.  PATCH = true (after
// template_tma_kernel): only rows that differ from the template - matched or neutral anchors - are
// written, class columns and offsets only.
template <bool PATCH>
__global__ void __launch_bounds__(E3_THREADS)
write_tma_kernel(const GtPrep* __restrict__ gtp, const long long* __restrict__ gt_off,
                 const Box<double>* __restrict__ abox, const float4* __restrict__ aboxf, const double* __restrict__ tail,
                 const int* __restrict__ match, EncArgs g, int tiles,
                 double* __restrict__ y, double* __restrict__ y2, int* __restrict__ midx) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int W = g.W, C = g.C;
    double* trow = reinterpret_cast<double*>(smem_raw);                                   // E3_ROWS x W
    float4* sgf = reinterpret_cast<float4*>(trow + (PATCH ? (size_t)0 : (size_t)E3_ROWS * W));
    const int tile = blockIdx.x % tiles;
    const int b = blockIdx.x / tiles;
    const int a0 = tile * E3_ROWS;
    const int rows = min(E3_ROWS, g.A - a0);
    const long long g0 = gt_off[b];
    const int m = (int)(gt_off[b + 1] - g0);
    const int tid = threadIdx.x;
    Box<double>* sgt = reinterpret_cast<Box<double>*>(sgf + m);          // ... and exact form
    int* smatch = reinterpret_cast<int*>(sgt + m);
    bool greg_local = true;
    for (int r = tid; r < m; r += E3_THREADS) {
        const Box<double> gb = gtp[g0 + r].box;
        sgt[r] = gb; sgf[r] = screen_box(gb); smatch[r] = match[g0 + r];
        greg_local = greg_local && regular_d(gb);
    }
    const bool all_gt_regular = __syncthreads_and(greg_local) != 0;

    const double thr_min = g.multi ? (g.pos_thr < g.neg_thr ? g.pos_thr : g.neg_thr) : g.neg_thr;
    const bool prune = thr_min > 0.0 && thr_min < INFINITY;
    const double prune_lo = thr_min * GUARD_LO;

    if (tid < rows) {
        const int a = a0 + tid;
        int gsel; bool neutral;
        decide_anchor(a, m, sgt, sgf, smatch, all_gt_regular, abox, aboxf, g, prune, prune_lo, &gsel, &neutral);
        double* row = trow + (size_t)tid * W;
        const double* t = tail + (size_t)a * 12;
        double off[4];
        int cls;
        if (gsel >= 0) {
            const GtPrep gp = gtp[g0 + gsel];
            double an[4] = {t[4], t[5], t[6], t[7]};
            double v[4] = {t[8], t[9], t[10], t[11]};
            encode_offsets(gp.data, an, v, g.coords, g.log_wh, off);
            cls = gp.cls;
            if (neutral && gp.cls == g.background_id) cls = -1;
        } else {
            off[0] = t[0]; off[1] = t[1]; off[2] = t[2]; off[3] = t[3];
            cls = neutral ? -1 : g.background_id;
        }
        if (PATCH) {
            if (gsel >= 0 || neutral) {
                double* dst = y + ((size_t)b * g.A + a) * W;
                for (int c = 0; c < C; ++c) dst[c] = (c == cls) ? 1.0 : 0.0;
#pragma unroll
                for (int k = 0; k < 4; ++k) dst[C + k] = off[k];
                if (y2) {
                    double* dst2 = y2 + ((size_t)b * g.A + a) * W;
                    for (int c = 0; c < C; ++c) dst2[c] = (c == cls) ? 1.0 : 0.0;
                }
            }
        } else {
            for (int c = 0; c < C; ++c) row[c] = 0.0;
            if (cls >= 0) row[cls] = 1.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) row[C + k] = off[k];
#pragma unroll
            for (int k = 4; k < 12; ++k) row[C + k] = t[k];
        }
        if (midx) midx[(size_t)b * g.A + a] = (gsel >= 0) ? gsel : (neutral ? -2 : -1);
    }
    if (PATCH) return;
    // generic-proxy writes to shared memory must be visible to the async proxy before the bulk store
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const size_t base = ((size_t)b * g.A + a0) * W;
    const uint32_t bytes = (uint32_t)((size_t)rows * W * sizeof(double));
    if (tid == 0) {
        tma_store_1d(y + base, trow, bytes);
        tma_store_commit_wait_read();
    }
    if (y2) {
        // diagnostics copy (ssd_input_encoder.py:412-416): same rows with the four offsets zeroed
        __syncthreads();
        if (tid < rows) {
            double* row = trow + (size_t)tid * W;
#pragma unroll
            for (int k = 0; k < 4; ++k) row[C + k] = 0.0;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            tma_store_1d(y2 + base, trow, bytes);
            tma_store_commit_wait_read();
        }
    }
}


// ===========================================================================
// Sparse path (E1 + E2 fused, then E3 patch).  Anchor sets of SSD models consist of a few SHAPE CLASSES
// (boxes per cell x predictor layers: 30 for SSD300); all anchors of a class share width, height and
// area up to rounding noise.  For a ground-truth row the position-independent bound
//     iou <= min(w) * min(h) / (area_gt + area_anchor - min(w) * min(h))
// is therefore one number per (row, class), and a whole class is skipped when that bound is below both
// the smallest threshold that matters (pos / neg IoU limits) and the row's best IoU found so far.
// The encoder keeps the anchors sorted by class, every class padded to whole warps, so the skip is a
// warp-uniform branch.  One CTA per image:
//   * slot loop: thread <-> anchor.  Per surviving (row, warp): float32 disjointness screen, exact
//     float64 IoU for overlapping pairs; the thread keeps its anchor's np.argmax over rows (match_multi,
//     neutral test), the warp folds its best pair into the row's running (value, first index) maximum.
//   * the m greedy rounds of match_bipartite_greedy on the row maxima (same literal semantics as
//     match_kernel); a row whose best column was taken is rescanned exhaustively.
//   * `cand[b, a]`: -1 unmatched, -2 neutral, r >= 0 matched to ground-truth row r (bipartite matches are
//     written last, in row order: last write wins, ssd_input_encoder.py:362).
// Everything skipped is provably below the value it is compared with, so the results are those of the
// exhaustive evaluation.
// ===========================================================================
constexpr int EF_MAX_M = 128;          // more ground-truth rows per image: general path
constexpr int EF_MAX_K = 64;           // more shape classes: general path

struct FastArgs {
    const int* perm;                   // slot -> anchor index (-1: padding)
    const Box<double>* pbox;           // slot -> anchor box (exact)
    const float4* pboxf;               // slot -> screening box (NaN for padding)
    const int* grp_cls;                // 32-slot group -> class
    const float4* grp_box;             // 32-slot group -> screening box around all its anchors
    const float4* cls;                 // class -> (w_up, h_up, area_lo, -)
    const int* cls_goff;               // class -> first group (K + 1 entries)
    int n_slots, K;
};

__device__ __forceinline__ float cut_of(double v) {
    return (v > 0.0) ? __fmul_rd(__double2float_rd(v), 1.0f - 0x1p-20f) : ((v != v) ? __int_as_float(0x7fc00000) : 0.f);
}

// np.argmax-faithful warp reduction: every lane ends with the best (value, index) pair of the warp
__device__ __forceinline__ void warp_best(double& bv, int& bi) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
}

// shape bound of (ground-truth row, class): iou <= u for every anchor of the class (+inf: no statement)
__device__ __forceinline__ float shape_bound(const GtPrep* gp, const float4 c) {
    if (!gp->regular) return INFINITY;
    const float mwh = __fmul_ru(fminf(gp->w_up, c.x), fminf(gp->h_up, c.y));
    const float den = __fsub_rd(__fadd_rd(gp->area_lo, c.z), mwh);
    return (den > 0.f) ? __fdiv_ru(mwh, den) : INFINITY;
}

// IoU of regular row r with the anchor in `slot`; pairs that are provably below `cut` (> 0: smaller than a
// threshold and than the row's known best) may come back as +0, which is below everything they are compared with
__device__ __forceinline__ double pair_iou(const Box<double>& gb, const float4& gf, float garea_lo, const float4& af, float aarea_lo,
                                           float cut, const Box<double>* __restrict__ pbox, int slot, Box<double>& ab, bool& have_ab) {
    if (screen_disjoint(gf, af)) return 0.0;                          // disjoint: +0 exactly
    if (cut > 0.f) {
        // float upper bound of the intersection over a lower bound of the union (screening boxes are rounded outward)
        const float iw = __fsub_ru(fminf(gf.z, af.z), fmaxf(gf.x, af.x));
        const float ih = __fsub_ru(fminf(gf.w, af.w), fmaxf(gf.y, af.y));
        const float in_up = __fmul_ru(iw, ih);
        const float den = __fsub_rd(__fadd_rd(garea_lo, aarea_lo), in_up);
        if (den > 0.f && in_up < __fmul_rd(cut, den)) return 0.0;
    }
    if (!have_ab) { ab = pbox[slot]; have_ab = true; }
    if (disjoint_d(gb, ab)) return 0.0;
    return iou_boxes<double>(gb, ab);
}

// ---- Row candidate lists.  For every ground-truth row the sparse path keeps the list of ALL anchors whose
// IoU with the row reaches a row-specific level tau_r > 0.  Whatever bipartite round asks "best anchor of row r
// among the columns not taken yet": if a listed anchor is still free, the best free listed one is the exact
// answer (every unlisted anchor is below tau_r, every taken column is zero), including np.argmax's
// first-index rule, because all anchors that tie at that value are listed too.  Only a row whose list
// ran dry (or overflowed, or has no level) is rescanned.
constexpr int EL_CAP = 128;                   // list entries per row
constexpr float EL_FRAC = 0.7f;               // tau_r = EL_FRAC * (best IoU inside the row's most promising class)

// Seed: one warp per ground-truth row evaluates the class with the largest shape bound (where the row's
// best anchor usually lives) exactly and derives tau_r from the best IoU found there.  Any tau_r > 0 is
// valid; this choice keeps the lists short (a few dozen entries) and almost never dry.
constexpr int ES_WARPS = 4;
__global__ void __launch_bounds__(ES_WARPS * 32)
seed_kernel(const double* __restrict__ gt_rows, EncArgs g, GtPrep* __restrict__ gtp, int n_gt, FastArgs f, float* __restrict__ gtau) {
    __shared__ int s_cut;
    __shared__ GtPrep s_gp;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = blockIdx.x;
    // (the row's preparation happens here, not in a launch of its own: thread 0 derives it, everybody uses it, and it is
    // left in global memory for the kernels that follow)
    if (threadIdx.x == 0) { s_gp = make_gtprep(gt_rows + (size_t)row * 5, g); gtp[row] = s_gp; s_cut = 0; }
    __syncthreads();
    const GtPrep* gp = &s_gp;
    if (!gp->regular) { if (threadIdx.x == 0) gtau[row] = 0.f; return; }
    const int K = f.K;
    const float u0 = (lane < K) ? shape_bound(gp, f.cls[lane]) : -1.f;
    const float u1 = (lane + 32 < K) ? shape_bound(gp, f.cls[lane + 32]) : -1.f;
    float bu = fmaxf(u0, u1);
    int bk = (u0 >= u1) ? lane : lane + 32;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ou = __shfl_xor_sync(0xffffffffu, bu, o);
        const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
        if (ou > bu || (ou == bu && ok < bk)) { bu = ou; bk = ok; }
    }
    const Box<double> gb = gp->box;
    const float4 gf = gp->sf;
    const float garea = gp->area_lo, carea = f.cls[bk].z;
    const int ge = f.cls_goff[bk + 1];
    __syncthreads();
    float cut = 0.f;
    for (int gb0 = f.cls_goff[bk]; gb0 < ge; gb0 += 32) {
        const int gl = gb0 + lane;
        const bool want = (gl < ge) && !screen_disjoint(gf, f.grp_box[gl]);
        unsigned gm = __ballot_sync(0xffffffffu, want);
        int nth = 0;
        while (gm) {
            const int gidx = gb0 + __ffs(gm) - 1;
            gm &= gm - 1;
            if ((nth++ % ES_WARPS) != warp) continue;           // the groups that meet the row, dealt to the warps in turn
            const int slot = gidx * 32 + lane;
            const int a = f.perm[slot];
            const float4 af = f.pboxf[slot];
            Box<double> ab = f.pbox[slot];
            bool have_ab = true;
            double sv = 0.0;
            if (a >= 0) sv = pair_iou(gb, gf, garea, af, carea, cut, f.pbox, slot, ab, have_ab);
            const int c = __reduce_max_sync(0xffffffffu, __float_as_int(cut_of(sv)));      // (non-negative floats order like ints)
            cut = fmaxf(cut, __int_as_float(c));
        }
    }
    if (lane == 0 && cut > 0.f) atomicMax(&s_cut, __float_as_int(cut));
    __syncthreads();
    if (threadIdx.x == 0) gtau[row] = __fmul_rd(__int_as_float(s_cut), EL_FRAC);
}

// E1 of the sparse path.  Grid (nblk, B): CTA (j, b) visits the 32-slot groups j, j + nblk, ... of image b,
// one group per warp and step, thread <-> anchor.  A (row, group) pair is worked on only if the class's
// shape bound reaches min(smallest IoU threshold, tau_r) and the group's bounding box meets the row; then
// the float32 disjointness screen and IoU bound, and the exact float64 IoU for what is left.  The thread
// keeps its anchor's np.argmax over the rows (match_multi + neutral test -> `cand`), pairs at or above tau_r
// are appended to the row's candidate list.  No block-level synchronisation after the prologue.
constexpr int EP_THREADS = 256;
__global__ void __launch_bounds__(EP_THREADS, 4)
pair_kernel(const GtPrep* __restrict__ gtp, const long long* __restrict__ gt_off, FastArgs f, EncArgs g,
            const float* __restrict__ gtau, int* __restrict__ cand, int* __restrict__ lcnt,
            double* __restrict__ lval, int* __restrict__ lidx, int* __restrict__ img_irr,
            int* __restrict__ plist, int* __restrict__ pcount) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int b = blockIdx.y, nblk = gridDim.x, blk = blockIdx.x;
    const long long g0 = gt_off[b];
    const int m = (int)(gt_off[b + 1] - g0);
    if (m == 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = EP_THREADS / 32;
    const int K = f.K;
    const int me = (m + 1) & ~1;                                           // (keeps the float4 array 16-byte aligned)
    Box<double>* sgt = reinterpret_cast<Box<double>*>(smem_raw);           // m
    float4* sgf = reinterpret_cast<float4*>(sgt + me);                     // m
    float* ub = reinterpret_cast<float*>(sgf + me);                        // m x K  shape bounds
    float* sga = ub + (size_t)me * K;                                      // m   lower bound of the area term
    float* stau = sga + me;                                                // m   list level (0: the row keeps no list)
    float* scut = stau + me;                                               // m   below this bound a pair cannot matter
    int* sreg = reinterpret_cast<int*>(scut + me);                         // m   row is regular
    float* scarea = reinterpret_cast<float*>(sreg + me);                   // K   lower bound of the class's area term
    unsigned* cneed = reinterpret_cast<unsigned*>(scarea + EF_MAX_K);      // K x 4  rows for which the class can matter (bit masks)
    const double thr_min = g.multi ? (g.pos_thr < g.neg_thr ? g.pos_thr : g.neg_thr) : g.neg_thr;
    const float thr_cut = cut_of(thr_min);
    for (int r = tid; r < m; r += EP_THREADS) {
        const GtPrep* gp = gtp + g0 + r;
        sgt[r] = gp->box; sgf[r] = gp->sf; sga[r] = gp->area_lo; sreg[r] = gp->regular;
        const float tau = gtau[g0 + r];
        stau[r] = tau;
        scut[r] = (tau > 0.f) ? fminf(thr_cut, __fmul_rd(tau, 1.0f - 0x1p-20f)) : thr_cut;
    }
    for (int e = tid; e < m * K; e += EP_THREADS) {
        const int r = e / K, k = e - r * K;
        ub[e] = shape_bound(gtp + g0 + r, f.cls[k]);
    }
    for (int k = tid; k < K; k += EP_THREADS) scarea[k] = f.cls[k].z;
    __syncthreads();
    // per class: the rows whose shape bound reaches a threshold or their list level (fixed for the whole image, so a
    // class nobody needs costs one shared-memory load per group)
    for (int e = tid; e < K * 4; e += EP_THREADS) {
        const int k = e >> 2, j = e & 3;
        unsigned mk = 0;
        for (int r = 32 * j; r < min(m, 32 * j + 32); ++r) mk |= (unsigned)(!(ub[r * K + k] < scut[r])) << (r & 31);
        cneed[e] = mk;
    }
    __syncthreads();

    bool irr = false;
    const int ngroups = f.n_slots >> 5;
    // (the loads of the next group are issued before the current one is worked on)
    const int gstep = NW * nblk;
    int gidx = warp * nblk + blk;
    int k_n = 0, a_n = -1;
    float4 gbox_n = make_float4(0.f, 0.f, 0.f, 0.f), af_n = gbox_n;
    if (gidx < ngroups) { k_n = f.grp_cls[gidx]; gbox_n = f.grp_box[gidx]; a_n = f.perm[gidx * 32 + lane]; af_n = f.pboxf[gidx * 32 + lane]; }
    for (; gidx < ngroups; gidx += gstep) {
        const int slot = gidx * 32 + lane;
        const int k = k_n;
        const float4 gbox = gbox_n;
        const int a = a_n;
        const float4 af = af_n;
        {
            const int gn = gidx + gstep;
            if (gn < ngroups) { k_n = f.grp_cls[gn]; gbox_n = f.grp_box[gn]; a_n = f.perm[gn * 32 + lane]; af_n = f.pboxf[gn * 32 + lane]; }
        }
        const unsigned* cm = cneed + 4 * k;
        if ((cm[0] | cm[1] | cm[2] | cm[3]) == 0u) continue;      // no row of this image can do anything with this class
        const float aarea = scarea[k];
        const bool live = a >= 0;
        Box<double> ab;
        bool have_ab = false;
        double cv = -INFINITY;
        int cg = 0x7fffffff;
        for (int r0 = 0; r0 < m; r0 += 32) {
            const unsigned want = cm[r0 >> 5];
            if (!want) continue;
            const int rl = r0 + lane;
            // the class can matter for the row (its shape bound reaches a threshold or the row's list level) and the row
            // touches this group of anchors at all (else every iou of the group is +0)
            bool need = (want >> lane) & 1u;
            if (need && sreg[rl] && screen_disjoint(sgf[rl], gbox)) need = false;
            unsigned rows = __ballot_sync(0xffffffffu, need);
            while (rows) {
                const int r = r0 + __ffs(rows) - 1;
                rows &= rows - 1;
                const bool rreg = sreg[r] != 0;
                double s = 0.0;
                if (live) {
                    if (rreg) s = pair_iou(sgt[r], sgf[r], sga[r], af, aarea, scut[r], f.pbox, slot, ab, have_ab);
                    else {
                        if (!have_ab) { ab = f.pbox[slot]; have_ab = true; }
                        s = iou_boxes<double>(sgt[r], ab);
                        irr |= (s < 0.0);
                    }
                    if (better(s, r, cv, cg)) { cv = s; cg = r; }
                }
                const float tau = stau[r];
                const bool hit = live && tau > 0.f && s >= (double)tau;
                const unsigned hm = __ballot_sync(0xffffffffu, hit);
                if (hm) {
                    int basep = 0;
                    if (lane == 0) basep = atomicAdd(&lcnt[g0 + r], __popc(hm));
                    basep = __shfl_sync(0xffffffffu, basep, 0) + __popc(hm & ((1u << lane) - 1u));
                    if (hit && basep < EL_CAP) {
                        lval[(size_t)(g0 + r) * EL_CAP + basep] = s;
                        lidx[(size_t)(g0 + r) * EL_CAP + basep] = a;
                    }
                }
            }
        }
        const bool pos = live && g.multi && (cv >= g.pos_thr);      // matching_utils.py:109-114
        const bool neu = live && (cv >= g.neg_thr);                  // ssd_input_encoder.py:388-390
        if (pos) cand[(size_t)b * g.A + a] = cg;
        else if (neu) cand[(size_t)b * g.A + a] = -2;
        // positions that differ from the template, for the patch kernel
        const unsigned wm = __ballot_sync(0xffffffffu, pos || neu);
        if (wm && plist) {
            int basep = 0;
            if (lane == 0) basep = atomicAdd(pcount, __popc(wm));
            basep = __shfl_sync(0xffffffffu, basep, 0);
            if (pos || neu) plist[basep + __popc(wm & ((1u << lane) - 1u))] = (int)((long long)b * g.A + a);
        }
    }
    if (__any_sync(0xffffffffu, irr) && lane == 0) atomicOr(&img_irr[b], 1);
}

// E2 of the sparse path: one CTA per image.  Warp 0 plays the m greedy rounds of match_bipartite_greedy
// (matching_utils.py:63-77, same literal semantics as match_kernel).  A row that lost its best column
// takes the best free entry of its candidate list (warp 0, on the spot); only when the list is dry are the
// other warps called in for a rescan: they visit the class with the largest shape bound first (it usually
// holds the new maximum), then every other group whose class bound reaches the running maximum of THIS
// scan and whose bounding box meets the row.
// `cand[b, a]`: the bipartite matches are written last, in row order (last write wins, ssd_input_encoder.py:362;
// a bipartite column is zeroed before match_multi and the neutral test, so it overrides the first pass).
constexpr int EG_WARPS = 8;

// best free entry of row `row`'s list (all lanes get it); (-inf, INT_MAX) if none
__device__ __forceinline__ void list_best(const double* __restrict__ lval, const int* __restrict__ lidx, long long row, int n,
                                          const int* taken, int ntaken, double& bv, int& bi) {
    const int lane = threadIdx.x & 31;
    bv = -INFINITY; bi = 0x7fffffff;
    for (int e = lane; e < n; e += 32) {
        const double v = lval[(size_t)row * EL_CAP + e];
        const int a = lidx[(size_t)row * EL_CAP + e];
        bool tk = false;
        for (int t = 0; t < ntaken; ++t) tk |= (taken[t] == a);
        if (!tk && better(v, a, bv, bi)) { bv = v; bi = a; }
    }
    warp_best(bv, bi);
}

__global__ void __launch_bounds__(EG_WARPS * 32)
greedy_kernel(const GtPrep* __restrict__ gtp, const long long* __restrict__ gt_off, FastArgs f, EncArgs g,
              const float* __restrict__ gtau, const int* __restrict__ lcnt, const double* __restrict__ lval, const int* __restrict__ lidx,
              const int* __restrict__ img_irr, int* __restrict__ cand, int* __restrict__ match,
              int* __restrict__ plist, int* __restrict__ pcount) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double red_val[EG_WARPS];
    __shared__ int red_idx[EG_WARPS];
    __shared__ int s_round, s_more, s_cut2, s_ln;
    __shared__ unsigned s_rescan[4];
    __shared__ float s_ub[EF_MAX_K];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    const long long g0 = gt_off[b];
    const int m = (int)(gt_off[b + 1] - g0);
    if (m == 0) return;
    const int K = f.K;
    const int ngroups = f.n_slots >> 5;
    // the small tables of the encoder in shared memory (the rescans chase them with dependent loads)
    float4* s_gbox = reinterpret_cast<float4*>(smem_raw);                  // ngroups
    float4* s_cls = s_gbox + ngroups;                                      // K
    double* rb_val = reinterpret_cast<double*>(s_cls + K);                 // m   row maximum ...
    int* rb_idx = reinterpret_cast<int*>(rb_val + m);                      // m   ... and its first index
    int* taken = rb_idx + m;                                               // m   zeroed columns
    int* done = taken + m;                                                 // m
    int* nlist = done + m;                                                 // m   usable list entries (0: none)
    int* s_match = nlist + m;                                              // m   matches (matching_utils.py:59-77)
    int* s_goff = s_match + m;                                             // K + 1
    int* s_gcls = s_goff + K + 1;                                          // ngroups
    int* s_list = s_gcls + ngroups;                                        // ngroups
    for (int i = tid; i < ngroups; i += EG_WARPS * 32) { s_gbox[i] = f.grp_box[i]; s_gcls[i] = f.grp_cls[i]; }
    for (int i = tid; i < K; i += EG_WARPS * 32) s_cls[i] = f.cls[i];
    for (int i = tid; i <= K; i += EG_WARPS * 32) s_goff[i] = f.cls_goff[i];
    const bool irrb = img_irr[b] != 0;
    if (tid < 4) s_rescan[tid] = 0;
    if (tid == 0) { s_round = 0; s_more = 0; }
    __syncthreads();
    // row maxima from the lists (the warps take rows in turn); rows without a usable list are rescanned
    for (int r = warp; r < m; r += EG_WARPS) {
        const int n = lcnt[g0 + r];
        const bool usable = !irrb && gtau[g0 + r] > 0.f && n <= EL_CAP;
        double bv; int bi;
        list_best(lval, lidx, g0 + r, usable ? n : 0, taken, 0, bv, bi);
        if (lane == 0) {
            nlist[r] = usable ? n : 0;
            rb_val[r] = bv; rb_idx[r] = bi; done[r] = 0;
            s_match[r] = 0;                // matches = np.zeros(num_ground_truth_boxes) (matching_utils.py:59)
            if (bi == 0x7fffffff) { atomicOr(&s_rescan[r >> 5], 1u << (r & 31)); s_more = 1; }
        }
    }
    __syncthreads();
    while (true) {
        if (s_more) {
            // ---- rescans, all warps
            const int ntaken = s_round;
            const unsigned rescan_rows[4] = {s_rescan[0], s_rescan[1], s_rescan[2], s_rescan[3]};   // (warp 0 rewrites them next phase)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                unsigned rows = rescan_rows[j];
                while (rows) {
                    const int r = 32 * j + __ffs(rows) - 1;
                    rows &= rows - 1;
                    const GtPrep* gp = gtp + g0 + r;
                    const bool rreg = gp->regular != 0;
                    const Box<double> gb = gp->box;
                    const float4 gf = gp->sf;
                    const float garea = gp->area_lo;
                    if (tid == 0) { s_cut2 = 0; s_ln = 0; }
                    const float u0 = (lane < K) ? shape_bound(gp, s_cls[lane]) : -1.f;
                    const float u1 = (lane + 32 < K) ? shape_bound(gp, s_cls[lane + 32]) : -1.f;
                    if (warp == 0) {
                        if (lane < K) s_ub[lane] = u0;
                        if (lane + 32 < K) s_ub[lane + 32] = u1;
                    }
                    // the class with the largest shape bound first: it usually holds the new maximum
                    float bu = fmaxf(u0, u1);
                    int bk = (u0 >= u1) ? lane : lane + 32;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const float ou = __shfl_xor_sync(0xffffffffu, bu, o);
                        const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
                        if (ou > bu || (ou == bu && ok < bk)) { bu = ou; bk = ok; }
                    }
                    __syncthreads();
                    double mv = rreg ? 0.0 : -INFINITY;
                    int mi = rreg ? 0 : 0x7fffffff;
                    // one group: exact maximum of the non-taken anchors, folded into (mv, mi) / s_cut2
                    auto visit = [&](int gidx, float carea) {
                        const int slot = gidx * 32 + lane;
                        const int a = f.perm[slot];
                        const float4 af = f.pboxf[slot];
                        Box<double> ab = f.pbox[slot];                          // (fetched with the others: one latency)
                        bool have_ab = true;
                        const bool live = a >= 0;
                        const float cut2 = __int_as_float(*reinterpret_cast<volatile int*>(&s_cut2));
                        double sv = 0.0;
                        if (live) {
                            bool tk = false;
                            for (int t = 0; t < ntaken; ++t) tk |= (taken[t] == a);
                            if (!tk) {
                                if (rreg) sv = pair_iou(gb, gf, garea, af, carea, cut2, f.pbox, slot, ab, have_ab);
                                else sv = iou_boxes<double>(gb, ab);
                            }
                        }
                        const bool contender = live && (rreg ? (sv > 0.0 && sv >= (double)cut2) : true);
                        if (__any_sync(0xffffffffu, contender)) {
                            double cbv = contender ? sv : -INFINITY;
                            int cbi = contender ? a : 0x7fffffff;
                            warp_best(cbv, cbi);
                            if (better(cbv, cbi, mv, mi)) {
                                mv = cbv; mi = cbi;
                                if (rreg && lane == 0) atomicMax(&s_cut2, __float_as_int(cut_of(cbv)));
                            }
                        }
                    };
                    {
                        const int ge = s_goff[bk + 1];
                        const float carea = s_cls[bk].z;
                        for (int gb0 = s_goff[bk]; gb0 < ge; gb0 += 32) {
                            const int gl = gb0 + lane;
                            bool want = gl < ge;
                            if (want && rreg) want = !screen_disjoint(gf, s_gbox[gl]);
                            unsigned gm = __ballot_sync(0xffffffffu, want);
                            int nth = 0;
                            while (gm) {
                                const int gidx = gb0 + __ffs(gm) - 1;
                                gm &= gm - 1;
                                if ((nth++ % EG_WARPS) == warp) visit(gidx, carea);
                            }
                        }
                    }
                    __syncthreads();
                    // every other group that can still matter
                    {
                        const float cut2 = __int_as_float(*reinterpret_cast<volatile int*>(&s_cut2));
                        for (int g1 = tid; g1 < ((ngroups + 31) & ~31); g1 += EG_WARPS * 32) {
                            bool want = false;
                            if (g1 < ngroups) {
                                const int kk = s_gcls[g1];
                                want = kk != bk;
                                if (want && rreg) want = !(s_ub[kk] < cut2) && !screen_disjoint(gf, s_gbox[g1]);
                            }
                            const unsigned wm = __ballot_sync(0xffffffffu, want);
                            int basep = 0;
                            if (lane == 0 && wm) basep = atomicAdd(&s_ln, __popc(wm));
                            basep = __shfl_sync(0xffffffffu, basep, 0);
                            if (want) s_list[basep + __popc(wm & ((1u << lane) - 1u))] = g1;
                        }
                    }
                    __syncthreads();
                    {
                        const int n = s_ln;
                        for (int i = warp; i < n; i += EG_WARPS) {
                            const int gidx = s_list[i];
                            const int kk = s_gcls[gidx];
                            if (rreg && s_ub[kk] < __int_as_float(*reinterpret_cast<volatile int*>(&s_cut2))) continue;
                            visit(gidx, s_cls[kk].z);
                        }
                    }
                    if (lane == 0) { red_val[warp] = mv; red_idx[warp] = mi; }
                    __syncthreads();
                    if (tid == 0) {
                        for (int w = 1; w < EG_WARPS; ++w)
                            if (better(red_val[w], red_idx[w], mv, mi)) { mv = red_val[w]; mi = red_idx[w]; }
                        rb_val[r] = mv; rb_idx[r] = mi;
                        nlist[r] = 0;                  // from now on the list says nothing about this row
                    }
                    __syncthreads();
                }
            }
        }
        // ---- rounds, warp 0, until the block is needed again
        if (warp == 0 && !irrb && m <= 32) {
            // Common case (regular image, at most 32 rows): lane <-> row, the row state lives in REGISTERS for the whole
            // run of rounds - a round is three REDUX, a shuffle and a ballot instead of a chain of shared-memory round
            // trips (measured with clock64: ~0.9 us per round through shared memory, 15 us for an 18-box image).  The
            // state goes back to shared memory when the block is needed (a dry list) or the rounds are over.
            int round = s_round;
            unsigned any = 0;
            const bool has = lane < m;
            double v = has ? rb_val[lane] : 0.0;
            int ix = has ? rb_idx[lane] : 0;
            bool dn = has ? (done[lane] != 0) : true;
            int mt = has ? s_match[lane] : 0;
            const int nl = has ? nlist[lane] : 0;
            while (round < m && !any) {
                // np.argmax over the rows' maxima: values are >= +0, so their bit patterns order like unsigned integers;
                // first row on ties
                const unsigned long long bits = has ? (unsigned long long)__double_as_longlong(v) : 0ull;
                const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
                const unsigned mh = __reduce_max_sync(0xffffffffu, has ? hi : 0u);
                const bool top = has && hi == mh;
                const unsigned ml = __reduce_max_sync(0xffffffffu, top ? lo : 0u);
                const int bg = (int)__reduce_min_sync(0xffffffffu, (top && lo == ml) ? (unsigned)lane : 0x7fffffffu);
                const int asel = __shfl_sync(0xffffffffu, ix, bg);
                if (lane == bg) { mt = asel; v = 0.0; ix = 0; dn = true; }      // matching_utils.py:71-77
                if (lane == 0) taken[round] = asel;
                __syncwarp();
                ++round;
                // rows whose best column was just zeroed take the best free entry of their list
                unsigned mk = __ballot_sync(0xffffffffu, has && !dn && ix == asel);
                unsigned dry = 0;
                while (mk) {
                    const int bit = __ffs(mk) - 1;
                    mk &= mk - 1;
                    const int nlb = __shfl_sync(0xffffffffu, nl, bit);
                    double lv; int li;
                    list_best(lval, lidx, g0 + bit, nlb, taken, round, lv, li);
                    if (li == 0x7fffffff) dry |= 1u << bit;
                    else if (lane == bit) { v = lv; ix = li; }
                }
                any |= dry;
            }
            if (has) { rb_val[lane] = v; rb_idx[lane] = ix; done[lane] = dn ? 1 : 0; s_match[lane] = mt; }
            if (lane == 0) { s_round = round; s_more = any ? 1 : 0; s_rescan[0] = any; s_rescan[1] = 0; s_rescan[2] = 0; s_rescan[3] = 0; }
        } else if (warp == 0) {
            int round = s_round;
            unsigned any = 0;
            while (round < m && !any) {
                double bv = -INFINITY; int bg = 0x7fffffff;
                for (int r = lane; r < m; r += 32) {
                    const double v = rb_val[r];
                    if (better(v, r, bv, bg)) { bv = v; bg = r; }
                }
                if (irrb) warp_best(bv, bg);
                else {
                    // regular input: the values are >= +0, so their bit patterns order like unsigned integers -
                    // three REDUX instead of five shuffle rounds (first row on ties: np.argmax)
                    const bool has = bg != 0x7fffffff;
                    const unsigned long long bits = has ? (unsigned long long)__double_as_longlong(bv) : 0ull;
                    const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
                    const unsigned mh = __reduce_max_sync(0xffffffffu, has ? hi : 0u);
                    const bool top = has && hi == mh;
                    const unsigned ml = __reduce_max_sync(0xffffffffu, top ? lo : 0u);
                    bg = (int)__reduce_min_sync(0xffffffffu, (top && lo == ml) ? (unsigned)bg : 0x7fffffffu);
                }
                const int asel = rb_idx[bg];
                __syncwarp();
                if (lane == 0) {
                    s_match[bg] = asel;
                    rb_val[bg] = 0.0;            // weight_matrix[ground_truth_index] = 0 -> argmax 0, weight 0
                    rb_idx[bg] = 0;
                    done[bg] = 1;
                    taken[round] = asel;         // weight_matrix[:, anchor_index] = 0
                }
                __syncwarp();
                ++round;
                // rows whose best column was just zeroed need a new maximum (every open row for irregular inputs)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (32 * j >= m) { if (lane == 0) s_rescan[j] = 0; continue; }
                    const int rr = 32 * j + lane;
                    const bool again = (rr < m) && !done[rr] && (irrb || rb_idx[rr] == asel);
                    unsigned mk = __ballot_sync(0xffffffffu, again);
                    unsigned dry = 0;
                    while (mk) {
                        const int bit = __ffs(mk) - 1;
                        mk &= mk - 1;
                        const int r = 32 * j + bit;
                        double lv; int li;
                        list_best(lval, lidx, g0 + r, nlist[r], taken, round, lv, li);
                        if (li == 0x7fffffff) dry |= 1u << bit;
                        else if (lane == 0) { rb_val[r] = lv; rb_idx[r] = li; }
                    }
                    if (lane == 0) s_rescan[j] = dry;
                    any |= dry;
                }
                __syncwarp();
            }
            if (lane == 0) { s_round = round; s_more = any ? 1 : 0; }
        }
        __syncthreads();
        if (!s_more) break;
    }
    // y_encoded[i, bipartite_matches, :-8] = labels_one_hot in row order: the LAST row that names an anchor wins
    __shared__ int s_base;
    if (tid == 0) s_base = plist ? atomicAdd(pcount, m) : 0;
    __syncthreads();
    for (int r = tid; r < m; r += EG_WARPS * 32) {
        const int a = s_match[r];
        bool last = true;
        for (int r2 = r + 1; r2 < m; ++r2) last = last && (s_match[r2] != a);
        const long long pos = (long long)b * g.A + a;
        if (last) cand[pos] = r;
        if (plist) plist[s_base + r] = (int)pos;
        match[g0 + r] = a;
    }
}

// E3 patch of the sparse path: rows whose `cand` entry is not -1 differ from the template
__device__ __forceinline__ void apply_one(int c, long long i, const GtPrep* __restrict__ gtp, const long long* __restrict__ gt_off,
                                          const double* __restrict__ tail, const EncArgs& g, double* __restrict__ y, double* __restrict__ y2) {
    const int W = g.W, C = g.C;
    const long long b = i / g.A;
    const int a = (int)(i - b * g.A);
    double* dst = y + (size_t)i * W;
    int cls = -1;
    if (c >= 0) {
        const GtPrep gp = gtp[gt_off[b] + c];
        const double* t = tail + (size_t)a * 12;
        double an[4] = {t[4], t[5], t[6], t[7]};
        double v[4] = {t[8], t[9], t[10], t[11]};
        double off[4];
        encode_offsets(gp.data, an, v, g.coords, g.log_wh, off);
#pragma unroll
        for (int k = 0; k < 4; ++k) dst[C + k] = off[k];
        cls = gp.cls;
    }
    // class vector: the template row is one-hot background; only the columns that change are written
    if (cls != g.background_id) {
        dst[g.background_id] = 0.0;
        if (cls >= 0) dst[cls] = 1.0;
        if (y2) {
            double* dst2 = y2 + (size_t)i * W;
            dst2[g.background_id] = 0.0;
            if (cls >= 0) dst2[cls] = 1.0;
        }
    }
}

// Each CTA sweeps windows of AP_WIN candidates: the few that are not -1 are queued in shared memory and
// then worked on by consecutive threads (the rows are long serial jobs - a warp with one such lane would
// otherwise idle 31 lanes for the whole job, one job after the other).
constexpr int AP_THREADS = 256;
constexpr int AP_WIN = AP_THREADS * 8;
__global__ void __launch_bounds__(AP_THREADS)
apply_kernel(const int* __restrict__ cand, const GtPrep* __restrict__ gtp, const long long* __restrict__ gt_off,
             const double* __restrict__ tail, EncArgs g, long long total, double* __restrict__ y, double* __restrict__ y2) {
    __shared__ int q_pos[AP_WIN];
    __shared__ int q_n;
    const int tid = threadIdx.x;
    const bool vec = (reinterpret_cast<uintptr_t>(cand) & 15) == 0;
    for (long long w0 = (long long)blockIdx.x * AP_WIN; w0 < total; w0 += (long long)gridDim.x * AP_WIN) {
        if (tid == 0) q_n = 0;
        __syncthreads();
        if (vec && w0 + AP_WIN <= total) {
            const int4* c4 = reinterpret_cast<const int4*>(cand + w0);
            const int4 u = c4[tid], v = c4[tid + AP_THREADS];
            if ((u.x & u.y & u.z & u.w) != -1) {
                if (u.x != -1) q_pos[atomicAdd(&q_n, 1)] = 4 * tid;
                if (u.y != -1) q_pos[atomicAdd(&q_n, 1)] = 4 * tid + 1;
                if (u.z != -1) q_pos[atomicAdd(&q_n, 1)] = 4 * tid + 2;
                if (u.w != -1) q_pos[atomicAdd(&q_n, 1)] = 4 * tid + 3;
            }
            if ((v.x & v.y & v.z & v.w) != -1) {
                const int o = 4 * (tid + AP_THREADS);
                if (v.x != -1) q_pos[atomicAdd(&q_n, 1)] = o;
                if (v.y != -1) q_pos[atomicAdd(&q_n, 1)] = o + 1;
                if (v.z != -1) q_pos[atomicAdd(&q_n, 1)] = o + 2;
                if (v.w != -1) q_pos[atomicAdd(&q_n, 1)] = o + 3;
            }
        } else {
            for (int e = tid; e < AP_WIN && w0 + e < total; e += AP_THREADS)
                if (cand[w0 + e] != -1) q_pos[atomicAdd(&q_n, 1)] = e;
        }
        __syncthreads();
        const int n = q_n;
        for (int e = tid; e < n; e += AP_THREADS) {
            const long long i = w0 + q_pos[e];
            apply_one(cand[i], i, gtp, gt_off, tail, g, y, y2);
        }
        __syncthreads();
    }
}

// the same over the list of positions pair_kernel / greedy_kernel recorded (duplicates are harmless: same row, same values)
__global__ void __launch_bounds__(128)
apply_list_kernel(const int* __restrict__ plist, const int* __restrict__ pcount, int* __restrict__ cand, int reset,
                  const GtPrep* __restrict__ gtp, const long long* __restrict__ gt_off,
                  const double* __restrict__ tail, EncArgs g, double* __restrict__ y, double* __restrict__ y2) {
    const int n = *pcount;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const long long i = plist[e];
        // (a position can be listed twice - multi match and bipartite match of the same anchor: the first visitor takes
        // the decision and, when the array is the library's own, leaves the entry clean for the next call)
        const int c = reset ? atomicExch(&cand[i], -1) : cand[i];
        if (c != -1) apply_one(c, i, gtp, gt_off, tail, g, y, y2);
    }
}

// ---- shape classes of the anchor set (host, once per encoder) ---------------------------------
static float f_up(double v) { float f = (float)v; if ((double)f < v) f = nextafterf(f, INFINITY); return f; }
static float f_down(double v) { float f = (float)v; if ((double)f > v) f = nextafterf(f, -INFINITY); return f; }

struct ShapeClasses {
    std::vector<int> perm;              // slot -> anchor (-1 padding), classes padded to multiples of 32
    std::vector<int> grp_cls;           // group of 32 slots -> class
    std::vector<float> grp_box;         // group -> x0, y0, x1, y1 of a box around all its anchors (rounded outward)
    std::vector<float> cls;             // K x 4: w_up, h_up, area_lo, 0
    std::vector<int> cls_goff;          // K + 1: first group of every class
    int K = 0;
};

// Anchors with (nearly) the same width and height form a class; classes are ordered by descending area so
// that the large anchors - few, and the likeliest best matches of large boxes - are visited first.
// Returns false when the anchor set does not qualify (improper boxes, too many classes).
static bool build_shape_classes(const std::vector<Box<double>>& ab, ShapeClasses* out) {
    const int A = (int)ab.size();
    std::vector<double> w(A), h(A);
    for (int a = 0; a < A; ++a) {
        const Box<double>& b = ab[a];
        w[a] = b.x1 - b.x0; h[a] = b.y1 - b.y0;
        if (!(w[a] > 0.0 && h[a] > 0.0 && b.area > 0.0 && std::isfinite(b.area) && std::isfinite(b.x0) && std::isfinite(b.y0) &&
              std::isfinite(b.x1) && std::isfinite(b.y1))) return false;
    }
    const double tol = 1e-6;
    std::vector<int> idx(A), cls_of(A, -1);
    for (int a = 0; a < A; ++a) idx[a] = a;
    std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return w[x] < w[y]; });
    int K = 0;
    for (int i0 = 0; i0 < A;) {
        int i1 = i0 + 1;
        while (i1 < A && w[idx[i1]] - w[idx[i1 - 1]] <= tol * w[idx[i1]]) ++i1;
        std::vector<int> sub(idx.begin() + i0, idx.begin() + i1);
        std::stable_sort(sub.begin(), sub.end(), [&](int x, int y) { return h[x] < h[y]; });
        for (size_t j0 = 0; j0 < sub.size();) {
            size_t j1 = j0 + 1;
            while (j1 < sub.size() && h[sub[j1]] - h[sub[j1 - 1]] <= tol * h[sub[j1]]) ++j1;
            if (K >= EF_MAX_K) return false;
            for (size_t j = j0; j < j1; ++j) cls_of[sub[j]] = K;
            ++K;
            j0 = j1;
        }
        i0 = i1;
    }
    std::vector<float> wu(K, 0.f), hu(K, 0.f), al(K, INFINITY);
    std::vector<int> count(K, 0);
    for (int a = 0; a < A; ++a) {
        const int k = cls_of[a];
        wu[k] = std::max(wu[k], f_up(w[a])); hu[k] = std::max(hu[k], f_up(h[a])); al[k] = std::min(al[k], f_down(ab[a].area));
        ++count[k];
    }
    std::vector<int> order(K);
    for (int k = 0; k < K; ++k) order[k] = k;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return al[x] > al[y]; });
    std::vector<int> rank(K);
    for (int i = 0; i < K; ++i) rank[order[i]] = i;
    std::vector<int> start(K + 1, 0);
    for (int i = 0; i < K; ++i) start[i + 1] = start[i] + ((count[order[i]] + 31) & ~31);
    out->K = K;
    out->perm.assign(start[K], -1);
    out->grp_cls.assign(start[K] / 32, 0);
    out->cls.assign((size_t)K * 4, 0.f);
    std::vector<int> fill(K, 0);
    for (int a = 0; a < A; ++a) {                       // ascending anchor index inside a class
        const int i = rank[cls_of[a]];
        out->perm[start[i] + fill[i]++] = a;
    }
    const int G = start[K] / 32;
    out->cls_goff.resize(K + 1);
    for (int i = 0; i <= K; ++i) out->cls_goff[i] = start[i] / 32;
    out->grp_box.assign((size_t)G * 4, 0.f);
    for (int gidx = 0; gidx < G; ++gidx) {
        float x0 = INFINITY, y0 = INFINITY, x1 = -INFINITY, y1 = -INFINITY;
        for (int l = 0; l < 32; ++l) {
            const int a = out->perm[gidx * 32 + l];
            if (a < 0) continue;
            x0 = std::min(x0, f_down(ab[a].x0)); y0 = std::min(y0, f_down(ab[a].y0));
            x1 = std::max(x1, f_up(ab[a].x1)); y1 = std::max(y1, f_up(ab[a].y1));
        }
        out->grp_box[4 * gidx] = x0; out->grp_box[4 * gidx + 1] = y0; out->grp_box[4 * gidx + 2] = x1; out->grp_box[4 * gidx + 3] = y1;
    }
    for (int i = 0; i < K; ++i) {
        for (int gidx = start[i] / 32; gidx < start[i + 1] / 32; ++gidx) out->grp_cls[gidx] = i;
        out->cls[4 * i] = wu[order[i]]; out->cls[4 * i + 1] = hu[order[i]]; out->cls[4 * i + 2] = al[order[i]];
    }
    return true;
}

// The per-call host data of an encode reaches the device in ONE copy: [image offsets | ground-truth rows | zeros].  The
// zeros initialise the small counters of the sparse path (list lengths per row, irregular-image flags, the patch
// position counter), so no memset launch is needed; the prepared rows (GtPrep) live behind them.
struct GtUpload {
    int64_t n_gt = 0;
    const long long* gt_off = nullptr;     // (B + 1) offsets relative to the shard
    const double* rows = nullptr;          // n_gt x 5
    int* zeros = nullptr;                  // n_gt + B + 1 ints, zero
    GtPrep* gtp = nullptr;                 // n_gt prepared rows (written on the device)
};
// Streams, events and scratch of one encode call: the device's own (synchronous host-output calls, profiling) or those of
// an encode lane (device-output calls, DevCtx::EncLane).
struct EncRes {
    cudaStream_t st, ts;
    cudaEvent_t ev_fork, ev_join;
    Buf* gt; Buf* partial; Buf* matches;
    size_t* cand_clean;
};
static int upload_gt(DevCtx* d, const EncRes& R, const double* gt, const int64_t* gt_offsets, int64_t b0, int64_t B, GtUpload* up) {
    const int64_t first = gt_offsets[b0], last = gt_offsets[b0 + B];
    const int64_t n = last - first;
    up->n_gt = n;
    // Staged in a pinned area of the event-guarded ring: a call that only enqueues work (on_device outputs) may return
    // before the copy ran, and neither the next call nor the caller's own reuse of `gt` may change the bytes it reads.
    const size_t off_bytes = (((size_t)(B + 1) * sizeof(long long)) + 63) & ~(size_t)63;
    const size_t gt_bytes = (((size_t)n * 5 * sizeof(double)) + 63) & ~(size_t)63;
    const size_t z_bytes = (((size_t)(n + B + 1) * sizeof(int)) + 63) & ~(size_t)63;
    const size_t copy_bytes = off_bytes + gt_bytes + z_bytes;
    void* hp = nullptr; int hs = 0;
    SSDC_TRY(d->stage_acquire(copy_bytes, &hp, &hs));
    char* h = reinterpret_cast<char*>(hp);
    long long* ho = reinterpret_cast<long long*>(h);
    for (int64_t i = 0; i <= B; ++i) ho[i] = gt_offsets[b0 + i] - first;
    if (n > 0) memcpy(h + off_bytes, gt + first * 5, (size_t)n * 5 * sizeof(double));
    memset(h + off_bytes + gt_bytes, 0, z_bytes);
    SSDC_TRY(R.gt->ensure(copy_bytes + (size_t)n * sizeof(GtPrep)));
    char* dv = R.gt->as<char>();
    SSDC_CUDA(cudaMemcpyAsync(dv, h, copy_bytes, cudaMemcpyHostToDevice, R.st));
    up->gt_off = reinterpret_cast<const long long*>(dv);
    up->rows = reinterpret_cast<const double*>(dv + off_bytes);
    up->zeros = reinterpret_cast<int*>(dv + off_bytes + gt_bytes);
    up->gtp = n > 0 ? reinterpret_cast<GtPrep*>(dv + copy_bytes) : nullptr;
    return d->stage_done(hs, R.st);
}

static int encode_body(ssdc_encoder* enc, int slot, const EncRes& R, const double* gt, const int64_t* gt_offsets, int64_t b0, int64_t B,
                       int max_m, double* y_dev, double* y2_dev, int* midx_dev);

// `lanes`: the call only enqueues work (device-resident outputs) and may run beside its neighbours on an encode lane.
int encode_dev(ssdc_encoder* enc, int slot, const double* gt, const int64_t* gt_offsets, int64_t b0, int64_t B,
               int max_m, double* y_dev, double* y2_dev, int* midx_dev, bool lanes) {
    ssdc_ctx* ctx = enc->ctx;
    DevCtx* d = &ctx->devs[slot];
    SSDC_CUDA(cudaSetDevice(d->device));
    if (B == 0) return SSDC_OK;
    int n_lanes = (int)ctx->opt[SSDC_OPT_ENC_LANES];
    if (n_lanes <= 0) n_lanes = 4;
    if (n_lanes > DevCtx::ENC_LANES) n_lanes = DevCtx::ENC_LANES;
    if (!lanes || ctx->profile || n_lanes == 1) {
        SSDC_TRY(d->wait_encodes());
        EncRes R = {d->stream, d->stream2, d->ev_fork, d->ev_join, &d->gt, &d->partial, &d->matches, &d->cand_clean};
        return encode_body(enc, slot, R, gt, gt_offsets, b0, B, max_m, y_dev, y2_dev, midx_dev);
    }
    SSDC_TRY(d->lanes_init());
    if (d->enc_next >= n_lanes) d->enc_next = 0;
    const int li = d->enc_next;
    d->enc_next = (li + 1) % n_lanes;
    DevCtx::EncLane& L = d->enc_lane[li];
    // behind everything the main stream holds so far (a reader of the output buffers, an earlier synchronous call) and
    // behind the sweeps in flight (they read their decode's input)
    // (an idle main stream has nothing to wait for: one query instead of an event record + wait per call)
    const cudaError_t q = cudaStreamQuery(d->stream);
    if (q == cudaErrorNotReady) {
        SSDC_CUDA(cudaEventRecord(d->ev_order, d->stream));
        SSDC_CUDA(cudaStreamWaitEvent(L.st, d->ev_order, 0));
    } else if (q != cudaSuccess) {
        set_error("cudaStreamQuery failed: %s", cudaGetErrorString(q));
        return SSDC_ERR_CUDA;
    }
    for (int k = 0; k < 2; ++k)
        if (d->sweep_pending[k]) SSDC_CUDA(cudaStreamWaitEvent(L.st, d->ev_sweep[k], 0));
    // two lanes never write the same bytes at the same time: a call whose outputs overlap those of a call in flight on
    // another lane runs behind it
    const size_t img_elems = (size_t)enc->A * (enc->p.n_classes + 12);
    const char* lo[3] = {reinterpret_cast<const char*>(y_dev), reinterpret_cast<const char*>(y2_dev), reinterpret_cast<const char*>(midx_dev)};
    const size_t len[3] = {(size_t)B * img_elems * sizeof(double), (size_t)B * img_elems * sizeof(double), (size_t)B * enc->A * sizeof(int)};
    for (int k = 0; k < DevCtx::ENC_LANES; ++k) {
        DevCtx::EncLane& O = d->enc_lane[k];
        if (k == li || !O.pending) continue;
        bool clash = false;
        for (int i = 0; i < 3 && !clash; ++i)
            for (int j = 0; j < 3 && !clash; ++j)
                clash = lo[i] && O.out_lo[j] && lo[i] < O.out_hi[j] && O.out_lo[j] < lo[i] + len[i];
        if (clash) SSDC_CUDA(cudaStreamWaitEvent(L.st, O.ev_done, 0));
    }
    EncRes R = {L.st, L.ts, L.ev_fork, L.ev_join, &L.gt, &L.partial, &L.matches, &L.cand_clean};
    const int r = encode_body(enc, slot, R, gt, gt_offsets, b0, B, max_m, y_dev, y2_dev, midx_dev);
    // (also after a failure: whatever was enqueued must be waited for)
    if (cudaEventRecord(L.ev_done, L.st) != cudaSuccess) { set_error("cudaEventRecord(encode lane) failed"); return SSDC_ERR_CUDA; }
    L.pending = true;
    for (int i = 0; i < 3; ++i) { L.out_lo[i] = lo[i]; L.out_hi[i] = lo[i] ? lo[i] + len[i] : nullptr; }
    // Scratch grows in every lane at once: the allocations (device-wide synchronisations) of a steady loop all land in
    // its first call instead of one per lane in the calls that follow.
    if (r == SSDC_OK) {
        for (int k = 0; k < n_lanes; ++k) {
            DevCtx::EncLane& O = d->enc_lane[k];
            if (k == li) continue;
            SSDC_TRY(O.gt.reserve(L.gt.cap));
            SSDC_TRY(O.partial.reserve(L.partial.cap));
            if (O.matches.cap < L.matches.cap) { SSDC_TRY(O.matches.reserve(L.matches.cap)); O.cand_clean = 0; }
        }
    }
    return r;
}

static int encode_body(ssdc_encoder* enc, int slot, const EncRes& R, const double* gt, const int64_t* gt_offsets, int64_t b0, int64_t B,
                       int max_m, double* y_dev, double* y2_dev, int* midx_dev) {
    ssdc_ctx* ctx = enc->ctx;
    DevCtx* d = &ctx->devs[slot];
    const ssdc_encode_params& p = enc->p;
    EncArgs g;
    memset(&g, 0, sizeof(g));
    g.A = (int)enc->A; g.C = p.n_classes; g.W = p.n_classes + 12; g.coords = p.coords;
    g.background_id = p.background_id; g.multi = p.matching_multi; g.log_wh = p.log_wh;
    g.chunks = (int)((enc->A + E1_CHUNK - 1) / E1_CHUNK);
    g.pos_thr = p.pos_iou_threshold; g.neg_thr = p.neg_iou_limit;
    g.d = (p.border_pixels == SSDC_BORDER_INCLUDE) ? 1.0 : (p.border_pixels == SSDC_BORDER_EXCLUDE ? -1.0 : 0.0);
    g.img_h = p.img_h; g.img_w = p.img_w; g.normalize = p.normalize;

    cudaStream_t st = R.st;
    const size_t row_bytes = (size_t)g.W * sizeof(double);
    const bool tma_ok = (reinterpret_cast<uintptr_t>(y_dev) % 16 == 0) && (!y2_dev || reinterpret_cast<uintptr_t>(y2_dev) % 16 == 0) &&
                        (((size_t)enc->A * row_bytes) % 16 == 0) && (((size_t)ET_ROWS * row_bytes) % 16 == 0) &&
                        ((((size_t)enc->A % ET_ROWS) * row_bytes) % 16 == 0);
    const bool no_overlap = ctx->opt[SSDC_OPT_ENC_NO_OVERLAP] != 0;
    const size_t smem_tpl = (size_t)ET_ROWS * row_bytes * (y2_dev ? 2 : 1);
    const bool overlap = tma_ok && !no_overlap && smem_tpl <= 64 * 1024;
#ifdef SSDC_TIMING_KNOBS
    const int dbg = getenv("SSDC_ENC_DBG") ? atoi(getenv("SSDC_ENC_DBG")) : 0;      // (timing experiments only; not in release builds)
#else
    constexpr int dbg = 0;
#endif
    bool forked = false;          // the template runs on the side stream: the patch waits for ev_join
    if (overlap && dbg != 1) {
        // E3 template stream: independent of the ground truth, so it starts first and runs beside E1 / E2.
        // (With per-launch profiling on, everything stays on the main stream so that each kernel is timed alone.)
        // (Measured and dropped: keeping a small batch's template on the call's own stream to save the four stream
        // operations of the fork / join - B = 32 on 4 lanes: 0.023 -> 0.029 ms per batch; the lanes are bound by the length
        // of the dependent chain on the device, not by the host.)
        cudaStream_t ts = ctx->profile ? st : R.ts;
        if (ts != st) {
            SSDC_CUDA(cudaEventRecord(R.ev_fork, st));
            SSDC_CUDA(cudaStreamWaitEvent(ts, R.ev_fork, 0));
        }
        const int tiles = (int)((enc->A + ET_ROWS - 1) / ET_ROWS);
        int splits = (2 * d->sm_count + tiles - 1) / tiles;
        if (splits > B) splits = (int)B;
        if (splits < 1) splits = 1;
        {
            LaunchScope ls(ctx, d, SSDC_K_ENC_WRITE);
            SSDC_TRY(ensure_dyn_smem(d->device, (const void*)template_tma_kernel, smem_tpl));
            template_tma_kernel<<<(unsigned)(tiles * splits), ET_THREADS, smem_tpl, ts>>>(enc->dev[slot].anchor_tail.as<double>(), g, tiles, splits, (int)B, y_dev, y2_dev);
            SSDC_TRY(check_launch("template_tma_kernel"));
        }
        if (ts != st) { SSDC_CUDA(cudaEventRecord(R.ev_join, ts)); forked = true; }
    }
    GtUpload up;
    SSDC_TRY(upload_gt(d, R, gt, gt_offsets, b0, B, &up));
    const int64_t n_gt = up.n_gt;
    const long long* gt_off = up.gt_off;
    const Box<double>* abox = enc->dev[slot].anchor_box.as<Box<double>>();
    const double* tail = enc->dev[slot].anchor_tail.as<double>();
    GtPrep* gtp = up.gtp;
    int* match = nullptr;

    // ---- sparse path: shape classes + fused matching, then a patch of the few rows that differ from the template
    const double thr_min = g.multi ? (g.pos_thr < g.neg_thr ? g.pos_thr : g.neg_thr) : g.neg_thr;
    const bool sparse = overlap && enc->fast_ok && max_m <= EF_MAX_M && thr_min > 0.0 && thr_min < INFINITY &&
                        (!g.multi || g.pos_thr == g.pos_thr) && g.neg_thr == g.neg_thr && ctx->opt[SSDC_OPT_ENC_GENERAL] == 0;
    if (sparse) {
        const long long total = (long long)B * enc->A;
        const bool use_plist = total < 0x7fffffffLL && ctx->opt[SSDC_OPT_ENC_DENSE_PATCH] == 0;
        int* cand = midx_dev;
        if (cand) {
            SSDC_CUDA(cudaMemsetAsync(cand, 0xff, (size_t)total * sizeof(int), st));
        } else {
            // The library's own decision array is kept all -1 BETWEEN calls: the patch kernel resets every entry it
            // consumes (the position list names exactly the entries that were written), so the 4 bytes per anchor are
            // only cleared once per allocation instead of once per call.
            const size_t need = (size_t)total * sizeof(int);
            if (need > R.matches->cap) *R.cand_clean = 0;
            SSDC_TRY(R.matches->ensure(need));
            cand = R.matches->as<int>();
            if (n_gt > 0) {
                if (*R.cand_clean < need) SSDC_CUDA(cudaMemsetAsync(cand, 0xff, R.matches->cap, st));
                *R.cand_clean = 0;          // (dirty until the patch kernel of this call has been enqueued)
            }
        }
        int* plist = nullptr;
        int* pcount = nullptr;
        if (n_gt > 0) {
            const ssdc_encoder::PerDev& pd = enc->dev[slot];
            FastArgs f;
            f.perm = pd.f_perm.as<int>(); f.pbox = pd.f_box.as<Box<double>>(); f.pboxf = pd.f_boxf.as<float4>();
            f.grp_cls = pd.f_grpcls.as<int>(); f.grp_box = pd.f_grpbox.as<float4>(); f.cls = pd.f_cls.as<float4>(); f.cls_goff = pd.f_clsgoff.as<int>(); f.n_slots = enc->f_slots; f.K = enc->f_classes;
            const int ngroups = f.n_slots / 32;
            // CTAs per image: enough CTAs to fill the device a few times over, at least one group per warp and step
            int nblk = (int)((6LL * d->sm_count + B - 1) / B);
#ifdef SSDC_TIMING_KNOBS
            if (const char* e = getenv("SSDC_ENC_NBLK")) nblk = atoi(e);
#endif
            if (nblk > ngroups / (EP_THREADS / 32)) nblk = ngroups / (EP_THREADS / 32);
            if (nblk > 16) nblk = 16;
            if (nblk < 1) nblk = 1;
            size_t off = 0;
            auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
            const size_t o_mt = carve((size_t)n_gt * sizeof(int));
            const size_t o_tau = carve((size_t)n_gt * sizeof(float));
            const size_t o_lv = carve((size_t)n_gt * EL_CAP * sizeof(double));
            const size_t o_li = carve((size_t)n_gt * EL_CAP * sizeof(int));
            const size_t o_pl = use_plist ? carve((size_t)(total + n_gt) * sizeof(int)) : 0;      // every anchor once + the bipartite matches
            SSDC_TRY(R.partial->ensure(off));
            char* base = R.partial->as<char>();
            match = reinterpret_cast<int*>(base + o_mt);
            int* lcnt = up.zeros;                       // list lengths, irregular flags and the position counter arrive
            int* img_irr = up.zeros + n_gt;             // zeroed with the upload
            float* gtau = reinterpret_cast<float*>(base + o_tau);
            double* lval = reinterpret_cast<double*>(base + o_lv);
            int* lidx = reinterpret_cast<int*>(base + o_li);
            pcount = img_irr + B;
            plist = use_plist ? reinterpret_cast<int*>(base + o_pl) : nullptr;
            {
                LaunchScope ls(ctx, d, SSDC_K_ENC_ROWBEST);
                seed_kernel<<<(unsigned)n_gt, ES_WARPS * 32, 0, st>>>(up.rows, g, gtp, (int)n_gt, f, gtau);
                SSDC_TRY(check_launch("seed_kernel"));
            }
            {
                LaunchScope ls(ctx, d, SSDC_K_ENC_ROWBEST);
                const size_t mm = ((size_t)max_m + 1) & ~(size_t)1;
                const size_t smem = mm * (sizeof(Box<double>) + sizeof(float4) + 3 * sizeof(float) + sizeof(int)) + mm * f.K * sizeof(float) + EF_MAX_K * (sizeof(float) + 4 * sizeof(unsigned)) + 32;
                dim3 grid((unsigned)nblk, (unsigned)B);
                SSDC_TRY(ensure_dyn_smem(d->device, (const void*)pair_kernel, smem));
                pair_kernel<<<grid, EP_THREADS, smem, st>>>(gtp, gt_off, f, g, gtau, cand, lcnt, lval, lidx, img_irr, plist, pcount);
                SSDC_TRY(check_launch("pair_kernel"));
            }
            {
                LaunchScope ls(ctx, d, SSDC_K_ENC_MATCH);
                const size_t smem = (size_t)(ngroups + f.K) * sizeof(float4) + (size_t)max_m * (sizeof(double) + 5 * sizeof(int)) + (f.K + 1 + 2 * (size_t)ngroups) * sizeof(int) + 16;
                SSDC_TRY(ensure_dyn_smem(d->device, (const void*)greedy_kernel, smem));
                greedy_kernel<<<(unsigned)B, EG_WARPS * 32, smem, st>>>(gtp, gt_off, f, g, gtau, lcnt, lval, lidx, img_irr, cand, match, plist, pcount);
                SSDC_TRY(check_launch("greedy_kernel"));
            }
        }
        if (forked && dbg != 1) SSDC_CUDA(cudaStreamWaitEvent(st, R.ev_join, 0));
        if (n_gt > 0 && dbg != 2) {
            LaunchScope ls(ctx, d, SSDC_K_ENC_PATCH);
            if (plist) {
                apply_list_kernel<<<(unsigned)(d->sm_count * 8), 128, 0, st>>>(plist, pcount, cand, midx_dev ? 0 : 1, gtp, gt_off, tail, g, y_dev, y2_dev);
                SSDC_TRY(check_launch("apply_list_kernel"));
                if (!midx_dev) *R.cand_clean = R.matches->cap;
            } else {
                long long blocks = (total + AP_WIN - 1) / AP_WIN;
                if (blocks > (long long)d->sm_count * 8) blocks = (long long)d->sm_count * 8;
                apply_kernel<<<(unsigned)blocks, AP_THREADS, 0, st>>>(cand, gtp, gt_off, tail, g, total, y_dev, y2_dev);
                SSDC_TRY(check_launch("apply_kernel"));
            }
        }
        return SSDC_OK;
    }
    if (n_gt > 0) {
        // scratch: partials (val, idx), row bests, taken columns, done flags, matches, irregular flags
        size_t off = 0;
        auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
        size_t o_pv = carve((size_t)n_gt * g.chunks * sizeof(double));
        size_t o_pi = carve((size_t)n_gt * g.chunks * sizeof(int));
        size_t o_mt = carve((size_t)n_gt * sizeof(int));
        size_t o_ir = carve((size_t)B * sizeof(int));
        size_t o_rm = carve((size_t)n_gt * sizeof(unsigned long long));
        SSDC_TRY(R.partial->ensure(off));
        char* base = R.partial->as<char>();
        double* part_val = reinterpret_cast<double*>(base + o_pv);
        int* part_idx = reinterpret_cast<int*>(base + o_pi);
        match = reinterpret_cast<int*>(base + o_mt);
        int* irregular = reinterpret_cast<int*>(base + o_ir);
        unsigned long long* rowmax_bits = reinterpret_cast<unsigned long long*>(base + o_rm);
        SSDC_CUDA(cudaMemsetAsync(irregular, 0, (size_t)B * sizeof(int), st));
        SSDC_CUDA(cudaMemsetAsync(rowmax_bits, 0, (size_t)n_gt * sizeof(unsigned long long), st));
        {
            LaunchScope ls(ctx, d, SSDC_K_ENC_ROWBEST);
            gt_prep_kernel<<<(unsigned)((n_gt + 127) / 128), 128, 0, st>>>(up.rows, (int)n_gt, g, gtp);
            SSDC_TRY(check_launch("gt_prep_kernel"));
        }
        {
            LaunchScope ls(ctx, d, SSDC_K_ENC_ROWBEST);
            // the last chunk (largest anchors) first, in its own launch: its results seed the per-row best
            // IoU that lets every other chunk discard most pairs with the shape bound
            rowbest_kernel<<<(unsigned)B, E1_THREADS, 0, st>>>(gtp, gt_off, abox, enc->dev[slot].anchor_boxf.as<float4>(), g, (int)B, g.chunks - 1, part_val, part_idx, irregular, rowmax_bits);
            SSDC_TRY(check_launch("rowbest_kernel"));
            if (g.chunks > 1) {
                ctx->launches.fetch_add(1, std::memory_order_relaxed);
                rowbest_kernel<<<(unsigned)(B * (g.chunks - 1)), E1_THREADS, 0, st>>>(gtp, gt_off, abox, enc->dev[slot].anchor_boxf.as<float4>(), g, (int)B, g.chunks - 2, part_val, part_idx, irregular, rowmax_bits);
                SSDC_TRY(check_launch("rowbest_kernel"));
            }
        }
        {
            LaunchScope ls(ctx, d, SSDC_K_ENC_MATCH);
            const size_t smem = (size_t)max_m * (sizeof(double) + sizeof(int) + 1) + (size_t)((enc->A + 31) / 32) * sizeof(unsigned) + 64;
            SSDC_CUDA(cudaFuncSetAttribute(match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            match_kernel<<<(unsigned)B, E2_THREADS, smem, st>>>(gtp, gt_off, abox, g, part_val, part_idx, irregular, match);
            SSDC_TRY(check_launch("match_kernel"));
        }
    }
    {
        const int tiles = (int)((enc->A + E3_ROWS - 1) / E3_ROWS);
        const size_t smem_gt = (size_t)max_m * (sizeof(Box<double>) + sizeof(float4) + sizeof(int)) + 16;
        if (overlap) {
            // patch the rows of matched / neutral anchors once the template has landed
            if (forked) SSDC_CUDA(cudaStreamWaitEvent(st, R.ev_join, 0));
            if (n_gt == 0 && !midx_dev) return SSDC_OK;
            LaunchScope ls(ctx, d, SSDC_K_ENC_PATCH);
            SSDC_CUDA(cudaFuncSetAttribute(write_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_gt));
            write_tma_kernel<true><<<(unsigned)(B * tiles), E3_THREADS, smem_gt, st>>>(gtp, gt_off, abox, enc->dev[slot].anchor_boxf.as<float4>(), tail, match, g, tiles, y_dev, y2_dev, midx_dev);
            SSDC_TRY(check_launch("write_tma_kernel<patch>"));
            return SSDC_OK;
        }
        const bool tma_ok3 = (reinterpret_cast<uintptr_t>(y_dev) % 16 == 0) && (!y2_dev || reinterpret_cast<uintptr_t>(y2_dev) % 16 == 0) &&
                             (((size_t)enc->A * row_bytes) % 16 == 0) && (((size_t)E3_ROWS * row_bytes) % 16 == 0) &&
                             ((((size_t)enc->A % E3_ROWS) * row_bytes) % 16 == 0);
        const size_t smem_tma = (size_t)E3_ROWS * row_bytes + smem_gt;
        LaunchScope ls(ctx, d, SSDC_K_ENC_WRITE);
        if (tma_ok3 && smem_tma <= 200 * 1024) {
            SSDC_CUDA(cudaFuncSetAttribute(write_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tma));
            write_tma_kernel<false><<<(unsigned)(B * tiles), E3_THREADS, smem_tma, st>>>(gtp, gt_off, abox, enc->dev[slot].anchor_boxf.as<float4>(), tail, match, g, tiles, y_dev, y2_dev, midx_dev);
            SSDC_TRY(check_launch("write_tma_kernel"));
            return SSDC_OK;
        }
        size_t smem = sizeof(RowMeta) * E3_ROWS + (size_t)max_m * (sizeof(Box<double>) + sizeof(float4) + sizeof(int)) + 16;
        SSDC_CUDA(cudaFuncSetAttribute(write_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        write_kernel<<<(unsigned)(B * tiles), E3_THREADS, smem, st>>>(gtp, gt_off, abox, enc->dev[slot].anchor_boxf.as<float4>(), tail, match, g, tiles, y_dev, y2_dev, midx_dev);
        SSDC_TRY(check_launch("write_kernel"));
    }
    return SSDC_OK;
}

}  // namespace ssdc

using namespace ssdc;

extern "C" {

int ssdc_encoder_create(ssdc_ctx* ctx, const double* anchors, int64_t A, const double* variances,
                        const ssdc_encode_params* p, ssdc_encoder** out) {
    if (!ctx || !anchors || !variances || !p || !out || A <= 0) { set_error("ssdc_encoder_create: bad argument"); return SSDC_ERR_ARG; }
    if (p->n_classes < 1 || p->coords < 0 || p->coords > 2 || p->border_pixels < 0 || p->border_pixels > 2 ||
        p->background_id < 0 || p->background_id >= p->n_classes) {
        set_error("ssdc_encoder_create: bad params"); return SSDC_ERR_ARG;
    }
    if (A > 0x7fffffff / (p->n_classes + 12)) { set_error("ssdc_encoder_create: too many anchors"); return SSDC_ERR_ARG; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    ssdc_encoder* enc = new ssdc_encoder();
    enc->ctx = ctx; enc->A = A; enc->p = *p;
    for (int i = 0; i < 4; ++i) enc->variances[i] = variances[i];
    enc->dev.resize(ctx->devs.size());
    const double dd = (p->border_pixels == SSDC_BORDER_INCLUDE) ? 1.0 : (p->border_pixels == SSDC_BORDER_EXCLUDE ? -1.0 : 0.0);
    ShapeClasses sc;
    for (size_t i = 0; i < ctx->devs.size(); ++i) {
        DevCtx& d = ctx->devs[i];
        int r = SSDC_OK;
        if (cudaSetDevice(d.device) != cudaSuccess) r = SSDC_ERR_CUDA;
        if (r == SSDC_OK) r = enc->dev[i].anchor_box.ensure((size_t)A * sizeof(Box<double>));
        if (r == SSDC_OK) r = enc->dev[i].anchor_tail.ensure((size_t)A * 12 * sizeof(double));
        if (r == SSDC_OK) r = enc->dev[i].anchor_boxf.ensure((size_t)A * sizeof(float4));
        if (r == SSDC_OK) r = d.t0buf.ensure((size_t)A * 4 * sizeof(double));
        if (r == SSDC_OK && cudaMemcpyAsync(d.t0buf.p, anchors, (size_t)A * 4 * sizeof(double), cudaMemcpyHostToDevice, d.stream) != cudaSuccess) r = SSDC_ERR_CUDA;
        if (r == SSDC_OK) {
            LaunchScope ls(ctx, &d, SSDC_K_THIN);
            anchor_prep_kernel<<<(unsigned)((A + 127) / 128), 128, 0, d.stream>>>(
                d.t0buf.as<double>(), (int)A, p->coords, dd, p->log_wh, variances[0], variances[1], variances[2], variances[3],
                enc->dev[i].anchor_box.as<Box<double>>(), enc->dev[i].anchor_tail.as<double>(), enc->dev[i].anchor_boxf.as<float4>());
            r = check_launch("anchor_prep_kernel");
        }
        if (r == SSDC_OK && cudaStreamSynchronize(d.stream) != cudaSuccess) { set_error("anchor_prep failed: %s", cudaGetErrorString(cudaGetLastError())); r = SSDC_ERR_CUDA; }
        if (r == SSDC_OK && i == 0) {
            // shape classes for the sparse path, from the boxes exactly as the kernels see them
            std::vector<Box<double>> hb((size_t)A);
            if (cudaMemcpy(hb.data(), enc->dev[0].anchor_box.p, (size_t)A * sizeof(Box<double>), cudaMemcpyDeviceToHost) != cudaSuccess) r = SSDC_ERR_CUDA;
            else enc->fast_ok = build_shape_classes(hb, &sc);
            if (enc->fast_ok) { enc->f_slots = (int)sc.perm.size(); enc->f_classes = sc.K; }
        }
        if (r == SSDC_OK && enc->fast_ok) {
            const size_t S = sc.perm.size();
            ssdc_encoder::PerDev& pd = enc->dev[i];
            if (r == SSDC_OK) r = pd.f_perm.ensure(S * sizeof(int));
            if (r == SSDC_OK) r = pd.f_box.ensure(S * sizeof(Box<double>));
            if (r == SSDC_OK) r = pd.f_boxf.ensure(S * sizeof(float4));
            if (r == SSDC_OK) r = pd.f_grpcls.ensure(sc.grp_cls.size() * sizeof(int));
            if (r == SSDC_OK) r = pd.f_grpbox.ensure(sc.grp_box.size() * sizeof(float));
            if (r == SSDC_OK) r = pd.f_clsgoff.ensure(sc.cls_goff.size() * sizeof(int));
            if (r == SSDC_OK) r = pd.f_cls.ensure(sc.cls.size() * sizeof(float));
            if (r == SSDC_OK &&
                (cudaMemcpy(pd.f_perm.p, sc.perm.data(), S * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess ||
                 cudaMemcpy(pd.f_grpcls.p, sc.grp_cls.data(), sc.grp_cls.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess ||
                 cudaMemcpy(pd.f_grpbox.p, sc.grp_box.data(), sc.grp_box.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
                 cudaMemcpy(pd.f_clsgoff.p, sc.cls_goff.data(), sc.cls_goff.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess ||
                 cudaMemcpy(pd.f_cls.p, sc.cls.data(), sc.cls.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess)) r = SSDC_ERR_CUDA;
            if (r == SSDC_OK) {
                LaunchScope ls(ctx, &d, SSDC_K_THIN);
                permute_anchor_kernel<<<(unsigned)((S + 127) / 128), 128, 0, d.stream>>>(
                    pd.f_perm.as<int>(), (int)S, pd.anchor_box.as<Box<double>>(), pd.anchor_boxf.as<float4>(), pd.f_box.as<Box<double>>(), pd.f_boxf.as<float4>());
                r = check_launch("permute_anchor_kernel");
                if (r == SSDC_OK && cudaStreamSynchronize(d.stream) != cudaSuccess) r = SSDC_ERR_CUDA;
            }
        }
        if (r != SSDC_OK) {
            for (auto& pd : enc->dev) {
                pd.anchor_box.release(); pd.anchor_tail.release(); pd.anchor_boxf.release();
                pd.f_perm.release(); pd.f_box.release(); pd.f_boxf.release(); pd.f_grpcls.release(); pd.f_grpbox.release(); pd.f_cls.release(); pd.f_clsgoff.release();
            }
            delete enc;
            return r;
        }
    }
    *out = enc;
    return SSDC_OK;
}

void ssdc_encoder_destroy(ssdc_encoder* enc) {
    if (!enc) return;
    std::lock_guard<std::mutex> lk(enc->ctx->mu);
    for (size_t i = 0; i < enc->dev.size(); ++i) {
        cudaSetDevice(enc->ctx->devs[i].device);
        enc->ctx->devs[i].wait_encodes();                 // (encodes in flight on the lanes read the tables freed below)
        cudaStreamSynchronize(enc->ctx->devs[i].stream);
        enc->dev[i].anchor_box.release();
        enc->dev[i].anchor_tail.release();
        enc->dev[i].anchor_boxf.release();
        enc->dev[i].f_perm.release(); enc->dev[i].f_box.release(); enc->dev[i].f_boxf.release();
        enc->dev[i].f_grpcls.release(); enc->dev[i].f_grpbox.release(); enc->dev[i].f_cls.release(); enc->dev[i].f_clsgoff.release();
    }
    delete enc;
}

int64_t ssdc_encoder_bad_image(const ssdc_encoder* enc) { return enc ? enc->bad_image : -1; }

int ssdc_encode(ssdc_encoder* enc, const double* gt, const int64_t* gt_offsets, int64_t B,
                int on_device, double* y_encoded, double* y_matched, int32_t* match_idx) {
    if (!enc || !gt_offsets || B < 0 || (B > 0 && !y_encoded)) { set_error("ssdc_encode: bad argument"); return SSDC_ERR_ARG; }
    ssdc_ctx* ctx = enc->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    enc->bad_image = -1;
    const int64_t n_total = gt_offsets[B] - gt_offsets[0];
    if (n_total > 0 && !gt) { set_error("ssdc_encode: gt is NULL"); return SSDC_ERR_ARG; }
    // host-side validation (ssd_input_encoder.py:333-336 and the class index used at :349)
    int max_m = 0;
    for (int64_t i = 0; i < B; ++i) {
        const int64_t m = gt_offsets[i + 1] - gt_offsets[i];
        if (m < 0) { set_error("ssdc_encode: gt_offsets must be non-decreasing"); return SSDC_ERR_ARG; }
        if (m > max_m) max_m = (int)m;
        for (int64_t r = gt_offsets[i]; r < gt_offsets[i + 1]; ++r) {
            const double* row = gt + r * 5;
            if (row[3] - row[1] <= 0 || row[4] - row[2] <= 0) {
                enc->bad_image = i;
                set_error("degenerate ground truth box in batch item %lld", (long long)i);
                return SSDC_ERR_DEGENERATE;
            }
            if (!(row[0] > -1.0 && row[0] < (double)enc->p.n_classes)) {
                enc->bad_image = i;
                set_error("class id %g of batch item %lld out of range [0, %d)", row[0], (long long)i, enc->p.n_classes);
                return SSDC_ERR_ARG;
            }
        }
    }
    if ((size_t)max_m * (sizeof(Box<double>) + sizeof(int)) > 160 * 1024) {
        set_error("ssdc_encode: more than %d ground-truth boxes in one image are not supported", (int)(160 * 1024 / 44));
        return SSDC_ERR_ARG;
    }
    const int n = (int)ctx->devs.size();
    const size_t img_elems = (size_t)enc->A * (enc->p.n_classes + 12);
    if (on_device) {
        if (n != 1) { set_error("ssdc_encode: device-resident output needs a single-device context"); return SSDC_ERR_ARG; }
        return encode_dev(enc, 0, gt, gt_offsets, 0, B, max_m, y_encoded, y_matched, match_idx, true);
    }
    for (int i = 0; i < n; ++i) {
        DevCtx& d = ctx->devs[i];
        int64_t per = (B + n - 1) / n;
        int64_t b0 = std::min<int64_t>(B, per * i), b1 = std::min<int64_t>(B, per * (i + 1));
        if (b1 == b0) continue;
        SSDC_CUDA(cudaSetDevice(d.device));
        SSDC_TRY(d.enc_out.ensure((size_t)(b1 - b0) * img_elems * sizeof(double)));
        if (y_matched) SSDC_TRY(d.enc_out2.ensure((size_t)(b1 - b0) * img_elems * sizeof(double)));
        if (match_idx) SSDC_TRY(d.enc_idx.ensure((size_t)(b1 - b0) * enc->A * sizeof(int)));
        SSDC_TRY(encode_dev(enc, i, gt, gt_offsets, b0, b1 - b0, max_m, d.enc_out.as<double>(),
                            y_matched ? d.enc_out2.as<double>() : nullptr, match_idx ? d.enc_idx.as<int>() : nullptr, false));
        SSDC_CUDA(cudaMemcpyAsync(y_encoded + (size_t)b0 * img_elems, d.enc_out.p, (size_t)(b1 - b0) * img_elems * sizeof(double), cudaMemcpyDeviceToHost, d.stream));
        if (y_matched)
            SSDC_CUDA(cudaMemcpyAsync(y_matched + (size_t)b0 * img_elems, d.enc_out2.p, (size_t)(b1 - b0) * img_elems * sizeof(double), cudaMemcpyDeviceToHost, d.stream));
        if (match_idx)
            SSDC_CUDA(cudaMemcpyAsync(match_idx + (size_t)b0 * enc->A, d.enc_idx.p, (size_t)(b1 - b0) * enc->A * sizeof(int), cudaMemcpyDeviceToHost, d.stream));
    }
    for (DevCtx& d : ctx->devs) {
        SSDC_CUDA(cudaSetDevice(d.device));
        SSDC_CUDA(cudaStreamSynchronize(d.stream));
    }
    return SSDC_OK;
}

int ssdc_encoding_template(ssdc_encoder* enc, int64_t B, double* out) {
    // generate_encoding_template (ssd_input_encoder.py:550-611): all-zero class vector, anchors in
    // both coordinate slots, variances.  Host-side tiling of construction-time data is not on the
    // hot path (the encode kernels never materialise it); it is produced from the device tail table
    // to keep a single source of truth.
    if (!enc || B < 0 || (B > 0 && !out)) { set_error("ssdc_encoding_template: bad argument"); return SSDC_ERR_ARG; }
    ssdc_ctx* ctx = enc->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevCtx& d = ctx->devs[0];
    SSDC_CUDA(cudaSetDevice(d.device));
    const int C = enc->p.n_classes, W = C + 12;
    std::vector<double> tail((size_t)enc->A * 12);
    SSDC_CUDA(cudaMemcpyAsync(tail.data(), enc->dev[0].anchor_tail.p, tail.size() * sizeof(double), cudaMemcpyDeviceToHost, d.stream));
    SSDC_CUDA(cudaStreamSynchronize(d.stream));
    for (int64_t b = 0; b < B; ++b) {
        for (int64_t a = 0; a < enc->A; ++a) {
            double* row = out + ((size_t)b * enc->A + a) * W;
            for (int c = 0; c < C; ++c) row[c] = 0.0;
            const double* t = tail.data() + (size_t)a * 12;
            for (int k = 0; k < 4; ++k) { row[C + k] = t[4 + k]; row[C + 4 + k] = t[4 + k]; row[C + 8 + k] = t[8 + k]; }
        }
    }
    return SSDC_OK;
}

}  // extern "C"
