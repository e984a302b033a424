// decode.cu - the decode + NMS pipeline of the SSD box codec (sm_100a).
//
// Replaces the numpy bodies of
//   decode_detections        /root/reference/localisation_part/ssd_encoder_decoder/ssd_output_decoder.py:111-226
//   decode_detections_fast   .../ssd_output_decoder.py:228-333
//   _greedy_nms/_greedy_nms2 .../ssd_output_decoder.py:77-109
//   DecodeDetections{,Fast}  /root/reference/localisation_part/keras_layers/keras_layer_DecodeDetections{,Fast}.py
//
// Kernels (one family per numpy stage; details in DESIGN.md section 3):
//   D1 decode_filter_tma_kernel : stream y_pred (B, A, C+12) once from HBM: warp-specialised persistent
//                             CTAs, TMA bulk copies of whole-row tiles into a shared-memory ring; per
//                             (anchor, class) confidence threshold, warp-aggregated compaction (image-sweep
//                             path: keys parked per warp in shared memory, one slot reservation per 32 keys);
//                             general path: anchor-offset decode of every anchor that produced a candidate
//                             (decode_filter_kernel: LDG fallback for unaligned inputs).
//   S  sweep_kernel         : hot configuration (finite top_k): per image radix-select + sort of the best
//                             candidates, anchor-offset decode of their boxes from y_pred, and ONE
//                             descending class-aware NMS sweep that stops at top_k;
//                             writes the final rows at (B, top_k, 6) stride.  scan_counts_kernel +
//                             sweep_pack_kernel pack them for the host copy at collect time.
//   general path (top_k='all', decode_detections_fast, float64 input, greedy_nms):
//      plan_kernel          : bins the segments by size into device-side work lists.
//   D2 sort_kernel          : segmented bitonic sort by (score desc, anchor asc), persistent CTAs.
//   D3 nms_kernel           : greedy NMS, one warp per segment, raw-corner screening + compacted exact
//                             pair decisions + bit-mask resolution; stops at the per-segment cap.
//   D4 count_scan_kernel +  : per image totals, exclusive scan to packed row offsets,
//      emit_merge_kernel      cross-class top-k as a warp k-way merge, row output.
#include "common.cuh"
#include "ctx.cuh"
#include <math.h>
#include <stdlib.h>
#include <sched.h>
#include <thread>
#include <condition_variable>

namespace ssdc {

constexpr int D1_THREADS = 256;
constexpr int NMS_WARPS = 4;
constexpr int NBINS = 5;          // 0: nms list, 1..4: sort bins
constexpr int CNT_LIST = 0;       // counters[0..4]  list sizes
constexpr int CNT_CURSOR = 8;     // counters[8..12] work cursors
constexpr int CNT_STAT_KEYS = 16, CNT_STAT_FLOORED = 17, CNT_STAT_FALLBACK = 18;      // image sweep statistics (ssdc_decode_stats)
constexpr int CNT_STAT_ERR = 19;      // first violated invariant of the image-sweep kernels (0: none); the offending access is skipped
constexpr int SORT_BYTES1 = 8 * 1024, SORT_BYTES2 = 32 * 1024, SORT_BYTES3 = 128 * 1024;   // shared-memory sort bins

template <typename T> struct alignas(16) SBox { T x0, y0, x1, y1; };

struct DecodeArgs {
    int A, C, W, NS, tiles, tile_rows;
    int floor_target;             // image sweep: D1 keeps at least this many best candidates per image complete (score floor)
    int have_hist;                // image sweep: D1 left per-image score histograms + floors (TMA kernel), else the sweep selects by radix passes
    int input_coords, log_wh, layer_assoc, ge;
    int do_nms, K, Kseg, always_sort;
    double iou_thr, sx, sy, d;
    int nseg;
    int sweep;      // per-image candidate lists with composite keys (image sweep path)
    int evict_first;   // D1's bulk copies carry the L2 evict_first hint
};

// ---------------------------------------------------------------------------
// D1: decode + threshold + compaction
// ---------------------------------------------------------------------------
template <typename InT>
__device__ __forceinline__ SBox<InT> decode_box(const InT* row, int C, const DecodeArgs& g) {
    // row[C..C+3] offsets, row[C+4..C+7] anchor, row[C+8..C+11] variances
    const InT o0 = row[C], o1 = row[C + 1], o2 = row[C + 2], o3 = row[C + 3];
    const InT a0 = row[C + 4], a1 = row[C + 5], a2 = row[C + 6], a3 = row[C + 7];
    const InT v0 = row[C + 8], v1 = row[C + 9], v2 = row[C + 10], v3 = row[C + 11];
    SBox<InT> b;
    if (g.input_coords == SSDC_COORDS_CENTROIDS) {
        // ssd_output_decoder.py:175-179 (+ bounding_box_utils.py:77-80)
        InT tw = o2 * v2, th = o3 * v3;
        if (g.log_wh) { tw = exp_cr(tw); th = exp_cr(th); }
        InT w = tw * a2, h = th * a3;
        InT cx, cy;
        if (g.layer_assoc) {   // keras_layer_DecodeDetections.py:124-125: (off * var) * size + centre
            cx = o0 * v0 * a2 + a0;
            cy = o1 * v1 * a3 + a1;
        } else {               // ssd_output_decoder.py:177-178: off * (var * size) + centre
            cx = o0 * (v0 * a2) + a0;
            cy = o1 * (v1 * a3) + a1;
        }
        InT hw = w / InT(2), hh = h / InT(2);
        b.x0 = cx - hw; b.y0 = cy - hh; b.x1 = cx + hw; b.y1 = cy + hh;
    } else if (g.input_coords == SSDC_COORDS_MINMAX) {
        // ssd_output_decoder.py:181-185: (xmin, xmax, ymin, ymax)
        InT aw = a1 - a0, ah = a3 - a2;
        InT p0 = o0 * v0 * aw + a0, p1 = o1 * v1 * aw + a1, p2 = o2 * v2 * ah + a2, p3 = o3 * v3 * ah + a3;
        b.x0 = p0; b.x1 = p1; b.y0 = p2; b.y1 = p3;
    } else {
        // ssd_output_decoder.py:187-190: (xmin, ymin, xmax, ymax)
        InT aw = a2 - a0, ah = a3 - a1;
        b.x0 = o0 * v0 * aw + a0; b.y0 = o1 * v1 * ah + a1; b.x1 = o2 * v2 * aw + a2; b.y1 = o3 * v3 * ah + a3;
    }
    return b;
}

// One tile = `rows` whole rows of y_pred staged in shared memory at `dst` (row stride W).  Thread t
// owns row t: it builds the bit mask of its classes that pass the confidence threshold, the warp
// aggregates the per-class counts with ballots, ONE warp-wide atomicAdd reserves the slots of all
// classes present in the warp (lane c reserves for class c), and the keys are written.  Warps never
// wait for one another, so there is no block-level synchronisation inside a tile.
// Deferred key write-out of the image-sweep path (TMA kernel).  A warp collects the composite keys of its tiles
// in its own shared-memory buffer of two halves.  When a half is full (or the image changes) the warp reserves
// the slots with ONE global atomicAdd and goes on filling the other half; the keys of a half leave for global
// memory when the half is needed again - several tiles later, when the atomic's round trip is long over.  So no
// warp waits for an atomic while it holds a pipeline stage, the atomics per image drop from one per warp-tile
// to one per D1_PEND_HALF keys, and the keys leave in contiguous runs.  (The CTAs own contiguous tile ranges,
// so a warp stays inside one image for ~35 tiles.)
constexpr int D1_PEND_HALF = 32;              // keys per half, two halves per warp
struct PendingKeys {
    unsigned long long* buf;       // this warp's 2 x D1_PEND_HALF keys in shared memory
    int cur;                       // half being filled
    int n_cur, b_cur;              // its keys and their image
    int n_pend, b_pend;            // the other half: keys waiting for their write-out, their image
    int base_pend;                 // lane 31: first reserved slot of the waiting half (result of the atomic)
    bool crowded;                  // the warp's previous tile: most rows had a candidate (process_tile_sweep walks instead of asking)
};
constexpr int D1_COOP_MAX = 10;                // rows with a candidate per warp-tile up to which the warp looks at them together
__device__ __forceinline__ void retire_pending(PendingKeys& pk, unsigned long long* __restrict__ keys, size_t img_stride) {
    if (pk.n_pend == 0) return;
    const int lane = threadIdx.x & 31;
    const int base = __shfl_sync(0xffffffffu, pk.base_pend, 31);
    unsigned long long* out = keys + (size_t)pk.b_pend * img_stride + base;
    const unsigned long long* src = pk.buf + (pk.cur ^ 1) * D1_PEND_HALF;
    if (lane < pk.n_pend && (size_t)(base + lane) < img_stride && base >= 0) out[lane] = src[lane];
    pk.n_pend = 0;
    __syncwarp();
}
// the half being filled becomes the waiting half (its slots are reserved now, written later)
__device__ __forceinline__ void close_current(PendingKeys& pk, int* __restrict__ seg_count, unsigned long long* __restrict__ keys,
                                              size_t img_stride) {
    if (pk.n_cur == 0) return;
    retire_pending(pk, keys, img_stride);
    if ((threadIdx.x & 31) == 31) pk.base_pend = atomicAdd(&seg_count[pk.b_cur], pk.n_cur);     // (nobody waits for it now)
    pk.n_pend = pk.n_cur; pk.b_pend = pk.b_cur;
    pk.cur ^= 1; pk.n_cur = 0;
}

// ---------------------------------------------------------------------------
// Speculative score floor of the image-sweep path.
//
// The sweep consumes the candidates of an image in descending score order and stops at top_k kept boxes, so of
// the (possibly hundreds of thousands of) candidates that pass the confidence threshold only a prefix is ever
// needed.  D1 therefore keeps, per CTA and image, a 256-bin histogram of the scores it has emitted so far
// (bins of 1/16 octave, monotone in the key order) and raises a FLOOR to the highest bin edge that still has
// `floor_target` emitted candidates at or above it: scores below the floor cannot be among the image's
// `floor_target` best and are not emitted.  Every floor ever published is a valid lower bound of the
// image's floor_target-th best score, so the set {score >= F} with F = the maximum published floor
// (atomicMax, g_floor) is complete in the key list, and the flushed histograms are complete for the bins >= F.
// The sweep trusts exactly that set; an image whose sweep runs dry inside it before top_k boxes are kept is
// rescanned without a floor (exact fallback, sweep_kernel).  Typical inputs never reach floor_target
// candidates per image, so nothing is dropped; dense inputs (low thresholds) shrink from 10^5 keys per image to
// a few times floor_target.
// ---------------------------------------------------------------------------
constexpr int FL_BINS = 256;
constexpr int FL_SHIFT = 19;                                         // ord32 >> 19: sign, exponent, 4 mantissa bits
constexpr unsigned FL_BASE = (0xBF800000u >> FL_SHIFT) - (FL_BINS - 1);    // bin 255 <-> scores in [1.0 * 2^(-1/16).., ...)
constexpr int FL_UPDATE = 256;                                       // histogram walk every FL_UPDATE emitted candidates
__host__ __device__ __forceinline__ int floor_bin_of_ord(unsigned ord) {
    const unsigned q = ord >> FL_SHIFT;
    return q <= FL_BASE ? 0 : (q - FL_BASE > (unsigned)(FL_BINS - 1) ? FL_BINS - 1 : (int)(q - FL_BASE));
}
// smallest ord32 value of bin q (q >= 1)
__host__ __device__ __forceinline__ unsigned floor_edge_ord(int q) { return (FL_BASE + (unsigned)q) << FL_SHIFT; }

struct FloorSlot {
    unsigned hist[FL_BINS];
    float thr_excl;                // a candidate needs score > thr_excl (the confidence threshold, or just below the floor's edge)
    int bin;                       // current floor bin (0: no floor)
    unsigned since;                // candidates histogrammed so far (monotone; update trigger)
    int image;                     // image the slot belongs to (-1: none)
};

// One warp re-derives the floor of its slot from the histogram (any snapshot of the monotone counters gives a valid bound).
__device__ __forceinline__ void floor_update(FloorSlot* fs, int b, int target, float thr, int* __restrict__ g_floor) {
    const int lane = threadIdx.x & 31;
    volatile unsigned* h = fs->hist;
    unsigned cnt[8];
    int mine = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) { cnt[q] = h[lane * 8 + q]; mine += (int)cnt[q]; }
    int suffix = mine;                                   // candidates in this lane's bins and above
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_down_sync(0xffffffffu, suffix, o);
        if (lane + o < 32) suffix += v;
    }
    const int above = suffix - mine;
    const bool has = above < target && suffix >= target;
    const unsigned hm = __ballot_sync(0xffffffffu, has);
    if (!hm) return;                                     // fewer than `target` candidates so far
    int qsel = 0;
    if (has) {
        int cum = above;
        qsel = lane * 8;
#pragma unroll
        for (int q = 7; q >= 0; --q) {
            cum += (int)cnt[q];
            if (cum >= target) { qsel = lane * 8 + q; break; }
        }
    }
    qsel = __shfl_sync(0xffffffffu, qsel, __ffs(hm) - 1);
    if (lane == 0) {
        const int seen = *reinterpret_cast<volatile int*>(&g_floor[b]);       // other CTAs working on the same image
        int nb = qsel > seen ? qsel : seen;
        if (nb > fs->bin) {
            const float edge_excl = unord32(floor_edge_ord(nb) - 1u);          // largest float below the bin's lower edge
            fs->bin = nb;
            *reinterpret_cast<volatile float*>(&fs->thr_excl) = edge_excl > thr ? edge_excl : thr;
            if (nb > seen) atomicMax(&g_floor[b], nb);
        }
    }
}

// Image-sweep variant of the tile filter (float32, per-class semantics, composite keys, one list per image).
// (Measured and dropped: filtering a crowded warp-tile twice - histogram first, raise the floor on the spot, emit only what
// is still above it.  It saves a few thousand keys per image at the start of every image and costs more instructions than
// those keys: SSD512 / conf 0.001: 1.39 M -> 1.57 M images/s without it, SSD300 dense 3.2 M -> 3.9 M.)
__device__ __forceinline__ void process_tile_sweep(const float* __restrict__ dst, int rows, int b, int a0,
                                                   const DecodeArgs& g, float thr, int* __restrict__ seg_count,
                                                   unsigned long long* __restrict__ gkeys, PendingKeys* pk,
                                                   FloorSlot* fs, int* __restrict__ g_floor) {
    const int W = g.W, NS = g.NS, A = g.A;
    const int tid = threadIdx.x, lane = tid & 31;
    const bool valid = tid < rows;
    const float* row = dst + (size_t)(valid ? tid : 0) * W;
    const int a = a0 + tid;
    // (the slot's threshold is raised concurrently by other warps: one lane reads it, so that the whole warp filters -
    // and branches - on the same value)
    const float t = fs ? __shfl_sync(0xffffffffu, *reinterpret_cast<volatile float*>(&fs->thr_excl), 0) : thr;
    int counted = 0;                          // candidates added to the histogram by this call (warp-uniform)
    bool crowded = pk ? pk->crowded : false;  // most rows of the warp's previous tile had a candidate (warp-uniform)
    // ssd_output_decoder.py:207-209, 32 classes per pass
    for (int c0 = 0; c0 < NS; c0 += 32) {
        const int nc = min(32, NS - c0);
        unsigned mask = 0;
        {
            // Rows with a candidate are usually rare (a few per warp-tile; fewer still once a score floor stands), so a row
            // first asks only WHETHER any of its classes passes: the maximum of its confidences (3-input FMNMX, half an
            // instruction per class; a NaN never wins, and `max > t` holds exactly when some non-NaN class is > t).  The
            // rows that pass are then looked at by the whole warp, lane <-> class: one shared-memory read, one compare and
            // one ballot give a row's class mask.  Where most rows pass (dense inputs before the floor bites) that costs
            // more than every lane walking its own row, so a warp whose last tile was crowded skips the question and walks.
            const float* cf = row + 1 + c0;
            int npass = 32;
            if (!crowded) {
                float m = -INFINITY;
                if (nc == 20) {                                   // (the VOC layout: straight-line code, 20 reads + 10 FMNMX3)
#pragma unroll
                    for (int c = 0; c < 20; c += 2) m = fmaxf(m, fmaxf(cf[c], cf[c + 1]));
                } else {
                    int c = 0;
#pragma unroll 1
                    for (; c + 4 <= nc; c += 4) m = fmaxf(fmaxf(m, fmaxf(cf[c], cf[c + 1])), fmaxf(cf[c + 2], cf[c + 3]));
#pragma unroll 1
                    for (; c < nc; ++c) m = fmaxf(m, cf[c]);
                }
                const unsigned pm = __ballot_sync(0xffffffffu, valid && m > t);
                if (pm == 0u) continue;
                npass = __popc(pm);
                if (npass <= D1_COOP_MAX) {
                    const float* wrow = dst + (size_t)(tid - lane) * W + 1 + c0 + (lane < nc ? lane : 0);      // this warp's first row, lane's class
                    for (unsigned rem = pm; rem; rem &= rem - 1) {
                        const int r = __ffs(rem) - 1;
                        const unsigned M = __ballot_sync(0xffffffffu, lane < nc && wrow[(size_t)r * W] > t);
                        if (lane == r) mask = M;
                    }
                }
            }
            if (npass > D1_COOP_MAX) {
                // (four classes per trip: one shared-memory read, one compare, one bit insert per class; the loop
                // bookkeeping must not double it)
                int c = 0;
#pragma unroll 1
                for (; c + 4 <= nc; c += 4) {
                    const float v0 = cf[c], v1 = cf[c + 1], v2 = cf[c + 2], v3 = cf[c + 3];
                    mask |= ((unsigned)(v0 > t) | ((unsigned)(v1 > t) << 1) | ((unsigned)(v2 > t) << 2) | ((unsigned)(v3 > t) << 3)) << c;
                }
                for (; c < nc; ++c) mask |= (unsigned)(cf[c] > t) << c;
                if (!valid) mask = 0;
                npass = __popc(__ballot_sync(0xffffffffu, mask != 0u));
            }
            crowded = npass > D1_COOP_MAX;
        }
        int total = __reduce_add_sync(0xffffffffu, __popc(mask));
        if (total == 0) continue;
        // keys carry the class: [ord32(score) | 255 - class | 2^24 - 1 - anchor]; one slot reservation per warp
        const int cnt = __popc(mask);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        unsigned long long* ck;
        const bool park = pk && total <= D1_PEND_HALF;
        if (park) {
            if (pk->n_cur && (pk->b_cur != b || pk->n_cur + total > D1_PEND_HALF)) close_current(*pk, seg_count, gkeys, (size_t)NS * A);
            ck = pk->buf + pk->cur * D1_PEND_HALF + pk->n_cur + (incl - cnt);
        } else {
            // a dense warp-tile (or no parking buffer): everything parked goes out first, then straight to global memory
            if (pk) { close_current(*pk, seg_count, gkeys, (size_t)NS * A); retire_pending(*pk, gkeys, (size_t)NS * A); }
            int base = 0;
            if (lane == 31) base = atomicAdd(&seg_count[b], incl);
            base = __shfl_sync(0xffffffffu, base, 31);
            ck = gkeys + (size_t)b * NS * A + base + (incl - cnt);
            if (base < 0 || (size_t)base + (size_t)incl > (size_t)NS * A) mask = 0;      // (capacity guard: cannot happen, never write out of bounds)
        }
        // (no box decode here: the sweep decodes the boxes of the candidates it actually visits from y_pred itself)
        for (unsigned mm = mask; mm; mm &= mm - 1) {
            const int c = __ffs(mm) - 1;
            const unsigned ord = ord32(row[1 + c0 + c]);
            *ck++ = ((unsigned long long)ord << 32) | ((unsigned long long)(0xffu - (unsigned)(c0 + c + 1)) << 24) |
                    (unsigned long long)(0xffffffu - (unsigned)a);
            if (fs) atomicAdd(&fs->hist[floor_bin_of_ord(ord)], 1u);
        }
        if (park) {
            __syncwarp();
            pk->n_cur += total; pk->b_cur = b;
        }
        counted += total;
    }
    if (pk) pk->crowded = crowded;
    if (fs && counted) {
        unsigned old = 0;
        if (lane == 0) old = atomicAdd(&fs->since, (unsigned)counted);
        old = __shfl_sync(0xffffffffu, old, 0);
        if (old / FL_UPDATE != (old + (unsigned)counted) / FL_UPDATE) floor_update(fs, b, g.floor_target, thr, g_floor);
    }
}

template <typename InT, bool FAST>
__device__ __forceinline__ void process_tile(const InT* __restrict__ dst, int rows, int b, int a0,
                                             const DecodeArgs& g, InT thr, int* __restrict__ seg_count,
                                             typename KeyOf<InT>::type* __restrict__ keys,
                                             SBox<InT>* __restrict__ boxes, int* __restrict__ aux_class) {
    typedef typename KeyOf<InT>::type KeyT;
    const int W = g.W, C = g.C, NS = g.NS, A = g.A;
    const int tid = threadIdx.x, lane = tid & 31;
    const bool valid = tid < rows;
    const InT* row = dst + (size_t)(valid ? tid : 0) * W;
    const int a = a0 + tid;
    const unsigned lt = (1u << lane) - 1u;

    if (FAST) {
        // ssd_output_decoder.py:292-293 (np.argmax = first maximum, NaN wins), :324-325
        InT best = row[0];
        bool nan = best != best;
        int cls = 0;
        for (int c = 1; c < C; ++c) {
            InT v = row[c];
            nan |= (v != v);
            if (v > best) { best = v; cls = c; }
        }
        const bool pass = valid && !nan && cls != 0 && (g.ge ? (best >= thr) : (best > thr));
        const unsigned m = __ballot_sync(0xffffffffu, pass);
        if (m) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&seg_count[b], __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (pass) {
                keys[(size_t)b * A + base + __popc(m & lt)] = KeyT::make(best, (uint32_t)a);
                aux_class[(size_t)b * A + a] = cls;
                boxes[(size_t)b * A + a] = decode_box<InT>(row, C, g);
            }
        }
        return;
    }

    // ssd_output_decoder.py:207-209, 32 classes per pass
    bool any = false;
    for (int c0 = 0; c0 < NS; c0 += 32) {
        const int nc = min(32, NS - c0);
        unsigned mask = 0;
        if (g.ge) {
            for (int c = 0; c < nc; ++c) mask |= (unsigned)(row[1 + c0 + c] >= thr) << c;
        } else {
            for (int c = 0; c < nc; ++c) mask |= (unsigned)(row[1 + c0 + c] > thr) << c;
        }
        if (!valid) mask = 0;
        const unsigned u = __reduce_or_sync(0xffffffffu, mask);
        if (!u) continue;
        any |= (mask != 0);
        const int total = __reduce_add_sync(0xffffffffu, __popc(mask));
        if (total <= 2 * __popc(u)) {
            // sparse: about one candidate per class present in the warp, aggregation would not save
            // atomics - every lane reserves its own slots
            for (unsigned mm = mask; mm; mm &= mm - 1) {
                const int c = __ffs(mm) - 1;
                const int pos = atomicAdd(&seg_count[(size_t)b * NS + c0 + c], 1);
                keys[((size_t)b * NS + c0 + c) * A + pos] = KeyT::make(row[1 + c0 + c], (uint32_t)a);
            }
            continue;
        }
        int mycnt = 0;
        for (unsigned uu = u; uu; uu &= uu - 1) {
            const int c = __ffs(uu) - 1;
            const unsigned m = __ballot_sync(0xffffffffu, (mask >> c) & 1u);
            if (lane == c) mycnt = __popc(m);
        }
        int mybase = 0;
        if ((u >> lane) & 1u) mybase = atomicAdd(&seg_count[(size_t)b * NS + c0 + lane], mycnt);
        for (unsigned uu = u; uu; uu &= uu - 1) {
            const int c = __ffs(uu) - 1;
            const unsigned m = __ballot_sync(0xffffffffu, (mask >> c) & 1u);
            const int base = __shfl_sync(0xffffffffu, mybase, c);
            if ((mask >> c) & 1u)
                keys[((size_t)b * NS + c0 + c) * A + base + __popc(m & lt)] = KeyT::make(row[1 + c0 + c], (uint32_t)a);
        }
    }
    if (any) boxes[(size_t)b * A + a] = decode_box<InT>(row, C, g);
}

// Fallback loader: 128-bit LDG -> STS staging of one tile per CTA (any alignment).
template <typename InT, bool FAST>
__global__ void __launch_bounds__(D1_THREADS)
decode_filter_kernel(const InT* __restrict__ y, DecodeArgs g, InT thr,
                     int* __restrict__ seg_count, typename KeyOf<InT>::type* __restrict__ keys,
                     SBox<InT>* __restrict__ boxes, int* __restrict__ aux_class) {
    constexpr int V = 16 / (int)sizeof(InT);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int W = g.W, A = g.A;
    InT* tile = reinterpret_cast<InT*>(smem_raw);
    const int tid = threadIdx.x;
    const int tile_id = blockIdx.x % g.tiles;
    const int b = blockIdx.x / g.tiles;
    const int a0 = tile_id * g.tile_rows;
    const int rows = min(g.tile_rows, A - a0);

    const InT* src = y + ((size_t)b * A + a0) * W;
    const int n = rows * W;
    const int mis = (int)((reinterpret_cast<uintptr_t>(src) / sizeof(InT)) % V);
    InT* dst = tile + mis;                       // smem offset keeps the global 16-byte phase
    const int head = min(n, (V - mis) % V);
    const int nvec = (n - head) / V;
    const int tail0 = head + nvec * V;
    if (tid < head) dst[tid] = src[tid];
    for (int e = tail0 + tid; e < n; e += D1_THREADS) dst[e] = src[e];
    {
        const int4* g4 = reinterpret_cast<const int4*>(src + head);
        int4* s4 = reinterpret_cast<int4*>(dst + head);
        constexpr int U = 8;
        for (int i0 = tid; i0 < nvec; i0 += D1_THREADS * U) {
            int4 r[U];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                int idx = i0 + k * D1_THREADS;
                if (idx < nvec) r[k] = ldg_stream(g4 + idx);
            }
#pragma unroll
            for (int k = 0; k < U; ++k) {
                int idx = i0 + k * D1_THREADS;
                if (idx < nvec) s4[idx] = r[k];
            }
        }
    }
    __syncthreads();
    if (sizeof(InT) == 4 && !FAST && g.sweep)
        process_tile_sweep(reinterpret_cast<const float*>(dst), rows, b, a0, g, (float)thr, seg_count,
                           reinterpret_cast<unsigned long long*>(keys), nullptr, nullptr, nullptr);
    else
        process_tile<InT, FAST>(dst, rows, b, a0, g, thr, seg_count, keys, boxes, aux_class);
}

// ---- TMA (cp.async.bulk) + mbarrier helpers ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
}
// 1-D bulk copy global -> shared through the TMA unit; completion is signalled on `bar`.
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ... with an L2 eviction-priority hint: the batch is read exactly once, so its lines should be the first to leave the L2
// (evict_first) instead of displacing the keys, counters and histograms the same kernel writes for the sweep.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_1d_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// Main loader: persistent, warp-specialised CTAs.  One producer warp keeps a ring of shared-memory
// stages filled with TMA bulk copies (UBLKCP) of whole-row tiles; eight consumer warps (32 rows each)
// filter a tile as soon as its `full` mbarrier flips and hand the stage back through an `empty`
// mbarrier.  There is no block-wide barrier: warps never wait for one another, only for data.
// Requires 16-byte aligned tile spans (host checks; else the LDG kernel above runs).
constexpr int D1_STAGES = 2;
constexpr int D1_TMA_THREADS = D1_THREADS + 32;

// Producer side of the score floor: adds a slot's histogram to the image's global one (complete for the bins at or
// above the image's final floor) and re-arms the slot for image `b`.  Called by the whole producer warp when no
// consumer can still be working on the slot's previous image.
__device__ __forceinline__ void floor_slot_flush(FloorSlot* fs, unsigned* __restrict__ g_hist) {
    const int lane = threadIdx.x & 31;
    if (fs->image >= 0) {
        unsigned* gh = g_hist + (size_t)fs->image * FL_BINS;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const unsigned v = fs->hist[lane * 8 + q];
            if (v) atomicAdd(&gh[lane * 8 + q], v);
        }
    }
    __syncwarp();
}
// `seen`: the image's published floor bin as lane 0 read it (a CTA that started the image earlier; any earlier snapshot is a
// valid bound, so the producer loads it one tile ahead and never waits for it in front of a copy it has to issue)
__device__ __forceinline__ void floor_slot_arm(FloorSlot* fs, int b, float thr, int seen) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int q = 0; q < 8; ++q) fs->hist[lane * 8 + q] = 0u;
    if (lane == 0) {
        const float edge_excl = seen > 0 ? unord32(floor_edge_ord(seen) - 1u) : thr;
        fs->bin = seen;
        fs->thr_excl = edge_excl > thr ? edge_excl : thr;
        fs->since = 0u;
        fs->image = b;
    }
    __syncwarp();
}

template <typename InT, bool FAST>
__global__ void __launch_bounds__(D1_TMA_THREADS)
decode_filter_tma_kernel(const InT* __restrict__ y, DecodeArgs g, InT thr, int total_tiles,
                         int* __restrict__ seg_count, typename KeyOf<InT>::type* __restrict__ keys,
                         SBox<InT>* __restrict__ boxes, int* __restrict__ aux_class,
                         int* __restrict__ g_floor, unsigned* __restrict__ g_hist) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[D1_STAGES];
    __shared__ __align__(8) uint64_t empty[D1_STAGES];
    __shared__ __align__(8) uint64_t done_bar;
    const int W = g.W, A = g.A;
    const size_t stage_bytes = (((size_t)g.tile_rows * W * sizeof(InT)) + 127) & ~(size_t)127;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ncw = (int)(blockDim.x >> 5) - 1;            // consumer warps (one row each lane); the last warp produces
    constexpr bool SWEEPABLE = (sizeof(InT) == 4) && !FAST;
    const bool sweep = SWEEPABLE && g.sweep;
    const bool floors = sweep && g.have_hist;             // score histograms + floor (needs >= D1_STAGES tiles per image, host-checked)
    unsigned long long* pend_base = reinterpret_cast<unsigned long long*>(smem_raw + (size_t)D1_STAGES * stage_bytes);
    FloorSlot* slots = reinterpret_cast<FloorSlot*>(pend_base + (size_t)ncw * 2 * D1_PEND_HALF);
    if (tid == 0) {
        for (int s = 0; s < D1_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], (uint32_t)ncw); }
        mbar_init(&done_bar, (uint32_t)ncw);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (floors) { slots[0].image = -1; slots[1].image = -1; }
    }
    __syncthreads();

    // every CTA owns a contiguous range of tiles (a warp then stays inside one image for many tiles, which is what
    // lets it batch its slot reservations and lets the CTA keep ONE score histogram per image)
    const int per = total_tiles / (int)gridDim.x, extra = total_tiles - per * (int)gridDim.x;
    const int t_begin = (int)blockIdx.x * per + min((int)blockIdx.x, extra);
    const int t_end = t_begin + per + ((int)blockIdx.x < extra ? 1 : 0);
    if (warp == ncw) {
        // ---- producer warp (lane 0 issues the copies; the whole warp serves the floor slots) ----
        int b = t_begin / g.tiles, tile_id = t_begin - b * g.tiles;
        int it = 0, armed = -1;
        const uint64_t pol = g.evict_first ? l2_policy_evict_first() : 0ull;
        int seen_next = (floors && lane == 0) ? *reinterpret_cast<const volatile int*>(&g_floor[b]) : 0;
        for (int t = t_begin; t < t_end; ++t, ++it, ++tile_id) {
            if (tile_id == g.tiles) { tile_id = 0; ++b; }
            const int s = it % D1_STAGES;
            if (it >= D1_STAGES) mbar_wait(&empty[s], (uint32_t)(((it / D1_STAGES) - 1) & 1));
            if (floors && b != armed) {
                // first tile of image b in this CTA.  Every consumer has released tile t - D1_STAGES, which lies in
                // image b - 1 or later (an image has >= D1_STAGES tiles, host-checked), so nobody touches slot b & 1
                // (image b - 2) any more; tile t itself cannot be consumed before its `full` barrier is armed below.
                FloorSlot* fs = &slots[b & 1];
                floor_slot_flush(fs, g_hist);
                floor_slot_arm(fs, b, (float)thr, seen_next);
                armed = b;
            }
            if (lane == 0) {
                const int a0 = tile_id * g.tile_rows;
                const int rows = min(g.tile_rows, A - a0);
                const uint32_t bytes = (uint32_t)((size_t)rows * W * sizeof(InT));
                mbar_expect_tx(&full[s], bytes);
                if (g.evict_first) tma_load_1d_hint(smem_raw + (size_t)s * stage_bytes, y + ((size_t)b * A + a0) * W, bytes, &full[s], pol);
                else tma_load_1d(smem_raw + (size_t)s * stage_bytes, y + ((size_t)b * A + a0) * W, bytes, &full[s]);
                // the next tile opens a new image: its published floor is on its way while this stage is being consumed
                if (floors && tile_id + 1 == g.tiles && t + 1 < t_end) seen_next = *reinterpret_cast<const volatile int*>(&g_floor[b + 1]);
            }
            __syncwarp();
        }
        if (floors) {
            mbar_wait(&done_bar, 0u);                         // all consumer warps are through their last tile
            floor_slot_flush(&slots[0], g_hist);
            floor_slot_flush(&slots[1], g_hist);
        }
        return;
    }
    // ---- consumer warps ----
    PendingKeys pend;
    pend.buf = pend_base + (size_t)warp * 2 * D1_PEND_HALF;
    pend.cur = 0; pend.n_cur = 0; pend.b_cur = 0; pend.n_pend = 0; pend.b_pend = 0; pend.base_pend = 0; pend.crowded = false;
    int b = t_begin / g.tiles, tile_id = t_begin - b * g.tiles;
    int it = 0;
    for (int t = t_begin; t < t_end; ++t, ++it, ++tile_id) {
        if (tile_id == g.tiles) { tile_id = 0; ++b; }
        const int s = it % D1_STAGES;
        mbar_wait(&full[s], (uint32_t)((it / D1_STAGES) & 1));
        const int a0 = tile_id * g.tile_rows;
        const int rows = min(g.tile_rows, A - a0);
        const InT* tile = reinterpret_cast<const InT*>(smem_raw + (size_t)s * stage_bytes);
        if (SWEEPABLE && sweep)
            process_tile_sweep(reinterpret_cast<const float*>(tile), rows, b, a0, g, (float)thr, seg_count,
                               reinterpret_cast<unsigned long long*>(keys), &pend, floors ? &slots[b & 1] : nullptr, g_floor);
        else
            process_tile<InT, FAST>(tile, rows, b, a0, g, thr, seg_count, keys, boxes, aux_class);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);             // this warp is done with stage s
    }
    if (sweep) {
        unsigned long long* gkeys = reinterpret_cast<unsigned long long*>(keys);
        close_current(pend, seg_count, gkeys, (size_t)g.NS * A);
        retire_pending(pend, gkeys, (size_t)g.NS * A);
        __syncwarp();
        if (lane == 0) mbar_arrive(&done_bar);
    }
}

// ---------------------------------------------------------------------------
// plan: bin the non-empty segments into work lists
// ---------------------------------------------------------------------------
__device__ __forceinline__ int warp_append(int* counter, bool want) {
    unsigned m = __ballot_sync(0xffffffffu, want);
    if (!m) return 0;
    int lane = threadIdx.x & 31;
    int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + __popc(m & ((1u << lane) - 1u));
}

__global__ void plan_kernel(const int* __restrict__ seg_count, int nseg, int* __restrict__ kept_count,
                            int* __restrict__ lists, int* __restrict__ counters, int sort_min, int n1, int n2, int n3) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    int n = (s < nseg) ? seg_count[s] : 0;
    if (s < nseg) kept_count[s] = 0;
    int bin = 0;
    if (n > n3) bin = 4; else if (n > n2) bin = 3; else if (n > n1) bin = 2; else if (n > sort_min) bin = 1;
    int pos = warp_append(&counters[CNT_LIST + 0], n > 0);
    if (n > 0) lists[pos] = s;
#pragma unroll
    for (int k = 1; k < NBINS; ++k) {
        int p = warp_append(&counters[CNT_LIST + k], bin == k);
        if (bin == k) lists[(size_t)k * nseg + p] = s;
    }
}

// ---------------------------------------------------------------------------
// D2: segmented sort (persistent CTAs over a work list)
// ---------------------------------------------------------------------------
template <typename KeyT, bool IN_SMEM>
__global__ void sort_kernel(KeyT* __restrict__ keys, const int* __restrict__ seg_count,
                            const int* __restrict__ list, int* __restrict__ counters, int bin,
                            DecodeArgs g, KeyT* __restrict__ scratch, int scratch_stride) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int s_idx;
    KeyT* s = IN_SMEM ? reinterpret_cast<KeyT*>(smem_raw) : scratch + (size_t)blockIdx.x * scratch_stride;
    const int total = counters[CNT_LIST + bin];
    for (;;) {
        if (threadIdx.x == 0) s_idx = atomicAdd(&counters[CNT_CURSOR + bin], 1);
        __syncthreads();
        const int idx = s_idx;
        __syncthreads();
        if (idx >= total) break;
        const int seg = list[idx];
        const int n = seg_count[seg];
        const int N = pow2_ceil(n);
        const bool by_anchor = !g.do_nms && (g.K <= 0 || n <= g.K);
        KeyT* gk = keys + (size_t)seg * g.A;
        for (int i = threadIdx.x; i < N; i += blockDim.x) s[i] = (i < n) ? gk[i] : KeyT::lowest();
        __syncthreads();
        block_bitonic_sort(s, N, by_anchor);
        for (int i = threadIdx.x; i < n; i += blockDim.x) gk[i] = s[i];
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// D3: greedy NMS, one warp per segment
// ---------------------------------------------------------------------------
// ---------------------------------------------------------------------------
// Pair decisions.  The reference keeps a box iff `iou <= iou_threshold`
// (ssd_output_decoder.py:91; NaN => dropped) with iou = RN(inter / union).
//
// `regular` box: x1 > x0, y1 > y0 and a finite positive area term.  For two regular boxes:
//   - disjoint (a zero side length): inter == 0, union = a1 + a2 > 0  =>  iou == +0 exactly, so the
//     pair is decided by `0 <= thr` alone.  Disjointness is tested on the RAW stored corners: the
//     scaling by a positive image size is monotone, so raw-disjoint implies scaled-disjoint;
//   - with p = thr * union: inter <= p (1 - e) implies RN(inter/union) < thr and inter >= p (1 + e)
//     implies RN(inter/union) > thr for e = 2^-48 (2^-20 in float32): every rounding contributes at
//     most half an ulp (2^-53 / 2^-24 relative), far inside the guard band.  Only pairs inside the
//     band take the exact division.
// ---------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ bool box_regular(const Box<T>& b) {
    return b.x1 > b.x0 && b.y1 > b.y0 && b.area > T(0) && b.area < T(INFINITY);
}
template <typename T>
__device__ __forceinline__ bool raw_disjoint(const SBox<T>& a, const SBox<T>& b) {
    return (a.x1 <= b.x0) || (b.x1 <= a.x0) || (a.y1 <= b.y0) || (b.y1 <= a.y0);
}
template <typename T> struct GuardBand;
template <> struct GuardBand<double> { static constexpr double lo = 1.0 - 0x1p-48, hi = 1.0 + 0x1p-48; };
template <> struct GuardBand<float> { static constexpr float lo = 1.0f - 0x1p-20f, hi = 1.0f + 0x1p-20f; };

template <typename StoreT, typename IouT>
__device__ __forceinline__ Box<IouT> scale_box(const SBox<StoreT>& s, IouT sx, IouT sy, IouT d) {
    return make_box<IouT>((IouT)s.x0 * sx, (IouT)s.y0 * sy, (IouT)s.x1 * sx, (IouT)s.y1 * sy, d);
}

// Does `kept` suppress `cand`?
template <typename IouT, bool TF>
__device__ __forceinline__ bool suppresses(const Box<IouT>& kept, const Box<IouT>& cand, IouT thr, bool thr_ok) {
    if (TF) {
        Box<float> a, b;
        a.x0 = (float)cand.x0; a.y0 = (float)cand.y0; a.x1 = (float)cand.x1; a.y1 = (float)cand.y1; a.area = 0.f;
        b.x0 = (float)kept.x0; b.y0 = (float)kept.y0; b.x1 = (float)kept.x1; b.y1 = (float)kept.y1; b.area = 0.f;
        return iou_tf(a, b) > (float)thr;
    }
    if (thr_ok && box_regular(kept) && box_regular(cand)) {
        const IouT sx = (cand.x1 < kept.x1 ? cand.x1 : kept.x1) - (cand.x0 > kept.x0 ? cand.x0 : kept.x0);
        const IouT sy = (cand.y1 < kept.y1 ? cand.y1 : kept.y1) - (cand.y0 > kept.y0 ? cand.y0 : kept.y0);
        if (!(sx > IouT(0)) || !(sy > IouT(0))) return false;       // iou == +0 <= thr
        const IouT inter = sx * sy;                                  // same value as iou_boxes computes
        const IouT uni = cand.area + kept.area - inter;
        const IouT p = thr * uni;
        if (inter <= p * GuardBand<IouT>::lo) return false;
        if (inter >= p * GuardBand<IouT>::hi) return true;
    }
    return !(iou_boxes<IouT>(cand, kept) <= thr);
}

// In-register bitonic sort of 32*R keys (key e lives in register e / 32 of lane e % 32).
template <typename KeyT, int R>
__device__ __forceinline__ void warp_sort_multi(KeyT (&k)[R], bool by_anchor) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int size = 2; size <= 32 * R; size <<= 1) {
#pragma unroll
        for (int jj = size >> 1; jj > 0; jj >>= 1) {
            if (jj >= 32) {
                const int rj = jj >> 5;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if ((r & rj) == 0) {
                        const int r2 = r | rj;
                        const bool canonical = (((r << 5) | lane) & size) == 0;
                        KeyT a = k[r], b = k[r2];
                        const bool swap = canonical ? key_before(b, a, by_anchor) : key_before(a, b, by_anchor);
                        if (swap) { k[r] = b; k[r2] = a; }
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    KeyT o = shfl_xor_key(k[r], jj);
                    const bool first_half = (lane & jj) == 0;
                    const bool canonical = (((r << 5) | lane) & size) == 0;
                    const bool take_other = (first_half == canonical) ? key_before(o, k[r], by_anchor)
                                                                      : key_before(k[r], o, by_anchor);
                    if (take_other) k[r] = o;
                }
            }
        }
    }
}

// Float32 evaluation of a pair of REGULAR boxes on their raw corners, valid when the union carries no border term
// (d == 0: the IoU is invariant under the scaling by the image size).  The float32 IoU differs from the real number the
// reference rounds (float64, relative error 1e-16) by < 2e-6 relative (a dozen roundings of 2^-24, no cancellation: the
// union is at least the larger area), so outside a 2e-5 band around the threshold the float64 decision is known.
// Returns 0: not suppressed, 1: suppressed, 2: inside the band (or out of the float32 range) - decide exactly.
__device__ __forceinline__ int pair_f32(const SBox<float>& a, const SBox<float>& b, float thr_lo, float thr_hi) {
    const float ix = fminf(a.x1, b.x1) - fmaxf(a.x0, b.x0);
    const float iy = fminf(a.y1, b.y1) - fmaxf(a.y0, b.y0);
    if (!(ix > 0.f) || !(iy > 0.f)) return 0;                        // disjoint: iou == +0 <= thr
    const float inter = ix * iy;
    const float uni = (a.x1 - a.x0) * (a.y1 - a.y0) + (b.x1 - b.x0) * (b.y1 - b.y0) - inter;
    if (!(uni > 1e-30f && uni < 1e30f)) return 2;
    if (inter < thr_lo * uni) return 0;
    if (inter > thr_hi * uni) return 1;
    return 2;
}

constexpr int NMS_SMEM_SORT = 128;              // segments of up to 128 candidates are sorted by the NMS warp itself
constexpr int NMS_REG_MAX = NMS_SMEM_SORT;
constexpr int NMS_QUEUE = 1024;                 // 16-bit pair entries: a full 32 x 32 block of pairs

// Warp-level bitonic sort of N (power of two, <= NMS_SMEM_SORT) keys in shared memory.
template <typename KeyT>
__device__ __forceinline__ void warp_smem_sort(KeyT* s, int N, bool by_anchor) {
    const int lane = threadIdx.x & 31;
#pragma unroll 1
    for (int k = 2; k <= N; k <<= 1) {
#pragma unroll 1
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll 1
            for (int i = lane; i < (N >> 1); i += 32) {
                const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));
                const int hi = lo | j;
                const KeyT a = s[lo], b = s[hi];
                const bool canonical = (lo & k) == 0;
                const bool swap = canonical ? key_before(b, a, by_anchor) : key_before(a, b, by_anchor);
                if (swap) { s[lo] = b; s[hi] = a; }
            }
            __syncwarp();
        }
    }
}

// The pair decision is called from two places; keeping one copy keeps the kernel inside the
// instruction cache.
template <typename StoreT, typename IouT, bool TF>
__device__ __noinline__ bool decide_pair(SBox<StoreT> kept, SBox<StoreT> cand, IouT sx, IouT sy, IouT d, IouT thr, bool thr_ok) {
    return suppresses<IouT, TF>(scale_box<StoreT, IouT>(kept, sx, sy, d), scale_box<StoreT, IouT>(cand, sx, sy, d), thr, thr_ok);
}

// D3.  One warp per segment.  Each step takes the next 32 candidates in canonical order:
//   (1) kept phase: lane k screens kept box k against the step's 32 candidates with a cheap
//       disjointness test on the raw corners (tight loop, no warp-wide synchronisation); only the
//       overlapping pairs are queued in shared memory and decided 32 at a time with all lanes busy
//       (these pairs are independent of one another);
//   (2) the same screening + batched decisions for the pairs inside the step give each candidate the
//       bit mask of the earlier candidates that would suppress it;
//   (3) the greedy order is then resolved with scalar bit operations, stopping at the segment cap.
template <typename StoreT, typename IouT, typename KeyT, bool TF>
__global__ void __launch_bounds__(NMS_WARPS * 32)
nms_kernel(KeyT* __restrict__ keys, const int* __restrict__ seg_count, int* __restrict__ kept_count,
           const int* __restrict__ list, int* __restrict__ counters,
           const SBox<StoreT>* __restrict__ boxes, DecodeArgs g, int KS) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // per-warp shared memory: raw corners of the kept boxes (cache) and of the step's candidates,
    // the sort buffer, the pair queue and the per-candidate suppression masks
    const size_t per_warp = (size_t)(KS + 32) * sizeof(SBox<StoreT>) + (size_t)NMS_SMEM_SORT * sizeof(KeyT) +
                            (size_t)NMS_QUEUE * sizeof(unsigned short) + 32 * sizeof(unsigned);
    unsigned char* base = smem_raw + (size_t)warp * per_warp;
    SBox<StoreT>* kraw = reinterpret_cast<SBox<StoreT>*>(base);
    SBox<StoreT>* craw = kraw + KS;
    KeyT* sbuf = reinterpret_cast<KeyT*>(craw + 32);
    unsigned* sup = reinterpret_cast<unsigned*>(sbuf + NMS_SMEM_SORT);
    unsigned short* queue = reinterpret_cast<unsigned short*>(sup + 32);

    const int total = counters[CNT_LIST + 0];
    const IouT sx = (IouT)g.sx, sy = (IouT)g.sy, d = (IouT)g.d, thr = (IouT)g.iou_thr;
    const unsigned lt = (1u << lane) - 1u;
    const bool thr_ok = thr > IouT(0) && thr < IouT(INFINITY);
    // the raw-corner screening is only exact for regular boxes, a usable threshold and positive scales
    const bool screen_ok = thr_ok && g.sx > 0.0 && g.sy > 0.0;

    // writes lane's set bits of `mask` as pairs (a << 5 | bit) behind one another; returns the total
    auto enqueue = [&](unsigned mask, unsigned a) -> int {
        const int np = __popc(mask);
        int start = np;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, start, o);
            if (lane >= o) start += v;
        }
        const int npairs = __shfl_sync(0xffffffffu, start, 31);
        start -= np;
        for (unsigned rem = mask; rem; rem &= rem - 1)
            queue[start++] = (unsigned short)((a << 5) | (unsigned)(__ffs(rem) - 1));
        __syncwarp();
        return npairs;
    };

    for (;;) {
        int idx = 0;
        if (lane == 0) idx = atomicAdd(&counters[CNT_CURSOR + 0], 1);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= total) break;
        const int seg = list[idx];
        const int b = seg / g.NS;
        const int n = seg_count[seg];
        const int cap = (g.Kseg > 0) ? min(g.Kseg, n) : n;
        KeyT* kp = keys + (size_t)seg * g.A;
        const SBox<StoreT>* bx = boxes + (size_t)b * g.A;
        const bool local_sort = n <= NMS_SMEM_SORT;
        const bool by_anchor = !g.do_nms && (g.K <= 0 || n <= g.K);

        if (!g.do_nms) {
            // `if iou_threshold:` falsy (ssd_output_decoder.py:326): every candidate is kept; small
            // segments still have to be put in order here (larger ones were sorted by sort_kernel)
            if (local_sort) {
                const int N = pow2_ceil(n);
                for (int i = lane; i < N; i += 32) sbuf[i] = (i < n) ? kp[i] : KeyT::lowest();
                __syncwarp();
                warp_smem_sort(sbuf, N, by_anchor);
                for (int i = lane; i < n; i += 32) kp[i] = sbuf[i];
            }
            if (lane == 0) kept_count[seg] = cap;
            continue;
        }

        // small segments are sorted here (no sort kernel, no write-back of the sorted keys):
        // <= 64 keys in registers (shuffle network), <= 128 in shared memory
        KeyT key0 = KeyT::lowest(), key1 = KeyT::lowest();
        if (n <= 32) {
            key0 = warp_sort((lane < n) ? kp[lane] : KeyT::lowest(), false);
        } else if (n <= 64) {
            KeyT two[2];
            two[0] = kp[lane];
            two[1] = (32 + lane < n) ? kp[32 + lane] : KeyT::lowest();
            warp_sort_multi<KeyT, 2>(two, false);
            key0 = two[0]; key1 = two[1];
        } else if (local_sort) {
            const int N = pow2_ceil(n);
            for (int i = lane; i < N; i += 32) sbuf[i] = (i < n) ? kp[i] : KeyT::lowest();
            __syncwarp();
            warp_smem_sort(sbuf, N, false);
        }

        int nkept = 0;
        bool screen = screen_ok;                 // cleared for good once an irregular box shows up
        for (int t0 = 0; t0 < n && nkept < cap; t0 += 32) {
            const int i = t0 + lane;
            const bool valid = i < n;
            KeyT key;
            if (n <= 64) key = (t0 == 0) ? key0 : key1;
            else if (local_sort) key = valid ? sbuf[i] : KeyT::lowest();
            else key = valid ? kp[i] : KeyT::lowest();
            SBox<StoreT> me;
            if (valid) me = bx[key.anchor()];
            else { me.x0 = me.y0 = StoreT(0); me.x1 = me.y1 = StoreT(1); }
            craw[lane] = me;
            if (screen) {
                const Box<IouT> mb = scale_box<StoreT, IouT>(me, sx, sy, d);
                if (__any_sync(0xffffffffu, valid && !box_regular(mb))) screen = false;
            }
            const unsigned vm = __ballot_sync(0xffffffffu, valid);
            __syncwarp();

            // ---- (1) against everything kept so far: lane <-> kept box ----
            unsigned dead = 0;                   // uniform: candidates of this step already suppressed
            for (int k0 = 0; k0 < nkept; k0 += 32) {
                const int k = k0 + lane;
                unsigned mask = 0;
                SBox<StoreT> kr;
                if (k < nkept) {
                    kr = (k < KS) ? kraw[k] : bx[kp[k].anchor()];
                    if (screen) {
#pragma unroll 8
                        for (int c = 0; c < 32; ++c) mask |= (unsigned)(!raw_disjoint(craw[c], kr)) << c;
                        mask &= vm & ~dead;
                    } else {
                        mask = vm & ~dead;
                    }
                }
                const int npairs = enqueue(mask, (unsigned)lane);
                for (int q0 = 0; q0 < npairs; q0 += 32) {
                    bool s = false;
                    unsigned c = 0;
                    if (q0 + lane < npairs) {
                        const unsigned pr = queue[q0 + lane];
                        c = pr & 31u;
                        const int kk = k0 + (int)(pr >> 5);
                        const SBox<StoreT> kb = (kk < KS) ? kraw[kk] : bx[kp[kk].anchor()];
                        s = decide_pair<StoreT, IouT, TF>(kb, craw[c], sx, sy, d, thr, thr_ok);
                    }
                    dead |= __reduce_or_sync(0xffffffffu, s ? (1u << c) : 0u);
                }
                __syncwarp();
                if ((dead | ~vm) == 0xffffffffu) break;          // every candidate of the step is suppressed
            }
            const bool alive = valid && !((dead >> lane) & 1u);
            const unsigned am = vm & ~dead;

            // ---- (2) pairs inside the step: which earlier candidates would suppress me ----
            unsigned ovl = 0;
            if (screen) {
#pragma unroll 8
                for (int j = 0; j < 32; ++j) ovl |= (unsigned)(!raw_disjoint(me, craw[j])) << j;
                ovl &= am & lt;
            } else {
                ovl = am & lt;
            }
            if (!alive) ovl = 0;
            sup[lane] = 0;
            const int npairs = enqueue(ovl, (unsigned)lane);
            for (int q0 = 0; q0 < npairs; q0 += 32) {
                if (q0 + lane < npairs) {
                    const unsigned pr = queue[q0 + lane];
                    const unsigned c = pr >> 5, j = pr & 31u;
                    if (decide_pair<StoreT, IouT, TF>(craw[j], craw[c], sx, sy, d, thr, thr_ok))
                        atomicOr(&sup[c], 1u << j);
                }
            }
            __syncwarp();

            // ---- (3) greedy resolution in canonical order, up to the cap ----
            // candidates nobody could suppress are kept outright; only the others are walked in order
            const unsigned mysup = sup[lane];
            const unsigned nz = __ballot_sync(0xffffffffu, mysup != 0u) & am;
            unsigned keptm = am & ~nz;
            for (unsigned rem = nz; rem; rem &= rem - 1) {
                const int c = __ffs(rem) - 1;
                const unsigned sc = __shfl_sync(0xffffffffu, mysup, c);
                if (!(sc & keptm)) keptm |= 1u << c;      // suppressors of c are all earlier than c
            }
            {   // stop at the cap: keep only the first (cap - nkept) of them
                const int room = cap - nkept;
                const bool mine = ((keptm >> lane) & 1u) && (__popc(keptm & lt) < room);
                keptm = __ballot_sync(0xffffffffu, mine);
            }
            if ((keptm >> lane) & 1u) {
                const int pos = nkept + __popc(keptm & lt);
                if (pos < KS) kraw[pos] = me;
                kp[pos] = key;
            }
            nkept += __popc(keptm);
            __syncwarp();
        }
        if (lane == 0) kept_count[seg] = nkept;
    }
}

// ---------------------------------------------------------------------------
// S: image sweep (decode_detections / DecodeDetections layer with a finite top_k; the hot configuration).
//
// Per-class greedy NMS followed by a cross-class top-k (ssd_output_decoder.py:205-221) equals one
// sweep over ALL candidates of the image in descending (score, then class, then anchor) order in
// which a candidate is only compared with kept boxes of its own class, stopped as soon as top_k
// boxes are kept: later candidates have lower scores, so they can neither enter the top-k nor
// influence a higher-scored decision.  Only the needed prefix of the candidate list is ever
// selected, sorted and examined.  One CTA per image:
//   select : D1 left a 256-bin score histogram of the image (exact for the bins at or above the image's score
//            floor): one walk over it yields the bin edge above which the next ~top_k candidates lie - no pass
//            over the keys.  (Radix passes over the keys remain as the fallback for degenerate score
//            distributions - hundreds of candidates inside one 1/16-octave bin - and for the LDG loader.)
//   compact: ONE pass over the image's keys moves that slice into shared memory;
//   sort   : every warp sorts a run of the slice in registers (shuffle network), the runs are merged by rank
//            (binary searches, one barrier) - no block-wide sorting network;
//   decode : anchor-offset decode of the slice's boxes from their y_pred rows (prefetched to L2 during the sort);
//   NMS    : panels of 128 candidates: thread <-> candidate computes (a) whether a kept box of its class
//            suppresses it and (b) the bit mask of the earlier same-class candidates of the panel that would -
//            per-class bit masks, raw-corner disjointness screening, exact division-free decisions (section 3.1
//            of DESIGN.md); ONE warp then resolves the greedy order of the panel with bit operations.
// An image whose trusted candidates (score >= floor) run out before top_k boxes are kept is rescanned from y_pred
// without a floor by the same CTA and swept again (exact fallback of D1's speculative score floor).
// ---------------------------------------------------------------------------
constexpr int SW_CHUNK = 512;           // candidates staged + sorted at a time
constexpr int SW_TARGET = 288;          // later chunks aim at >= this many (and <= SW_CHUNK)
constexpr int SW_KMAX = 256;            // largest top_k the sweep path handles
// (candidates resolved per NMS panel = threads of the CTA: 128 or 256)
constexpr int SW_KW = SW_KMAX / 32;

__device__ __forceinline__ int ck_cls(unsigned long long k) { return (int)(0xffu - (unsigned)((k >> 24) & 0xffu)); }
__device__ __forceinline__ unsigned ck_anchor(unsigned long long k) { return 0xffffffu - (unsigned)(k & 0xffffffu); }

// Sorts kA[0..cn) (cn <= SW_CHUNK) descending into kB.  Runs of 32*R keys are sorted by one warp each in registers,
// then every key finds its final position as the sum of its ranks in all runs (keys are unique).  Three barriers.
template <int R>
__device__ __forceinline__ void sort_runs(unsigned long long* kA) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long* run = kA + (size_t)warp * 32 * R;
    Key64 k[R];
#pragma unroll
    for (int r = 0; r < R; ++r) k[r].v = run[r * 32 + lane];
    warp_sort_multi<Key64, R>(k, false);
#pragma unroll
    for (int r = 0; r < R; ++r) run[r * 32 + lane] = k[r].v;
}
template <int SW_THREADS>
__device__ __forceinline__ void chunk_sort(unsigned long long* kA, unsigned long long* kB, int cn) {
    constexpr int SW_WARPS = SW_THREADS / 32;
    int S = 32 * SW_WARPS;
    while (S < cn) S <<= 1;
    const int L = S / SW_WARPS;                      // run length: 32, 64 or 128
    for (int i = cn + threadIdx.x; i < S; i += SW_THREADS) kA[i] = 0ull;       // (0 sorts last; no real key is 0)
    __syncthreads();
    if (L == 32) sort_runs<1>(kA);
    else if (L == 64) sort_runs<2>(kA);
    else sort_runs<4>(kA);
    __syncthreads();
    for (int e = threadIdx.x; e < S; e += SW_THREADS) {
        const unsigned long long key = kA[e];
        if (key == 0ull) continue;
        const int r = e / L;
        // number of keys of every other run that come before `key`: SW_WARPS - 1 independent binary searches, advanced
        // in lock step so that their shared-memory reads overlap
        int lo[SW_WARPS];
#pragma unroll
        for (int q = 0; q < SW_WARPS; ++q) lo[q] = 0;
        for (int half = L >> 1; half > 0; half >>= 1) {
#pragma unroll
            for (int q = 0; q < SW_WARPS; ++q)
                if (kA[q * L + lo[q] + half - 1] > key) lo[q] += half;          // (run q sorted descending, zero padded)
        }
        int rank = e - r * L;
#pragma unroll
        for (int q = 0; q < SW_WARPS; ++q) {
            const int cntq = lo[q] + (kA[q * L + lo[q]] > key ? 1 : 0);
            if (q != r) rank += cntq;
        }
        if (rank < S) kB[rank] = key;
    }
    __syncthreads();
}

// Sorts kA[0..cn) descending IN PLACE (kB is scratch) by distributing the keys over the 1/16-octave score bins of the
// histogram (a counting sort: the bins are monotone in the key order) and ranking every key inside its own bin.  A
// slice of a few hundred candidates near the top of a score distribution spreads over dozens of bins with a handful of
// keys each, so this is two shared-memory atomic passes, one warp scan and a short scan per key instead of a sorting
// network.  Returns false (nothing changed) when a bin holds more than SW_BIN_MAX keys - score ties, saturated
// confidences -; the caller then uses the sorting network.  `scratch`: 2 * FL_BINS words.
constexpr int SW_BIN_MAX = 48;
template <int SW_THREADS>
__device__ __forceinline__ bool bin_sort(unsigned long long* kA, unsigned long long* kB, int cn, unsigned* scratch, int* s_flag) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned* cnt = scratch;                 // keys per bin, then: first position of the bin
    unsigned* fill = scratch + FL_BINS;      // next free position of the bin during the scatter
    for (int i = tid; i < FL_BINS; i += SW_THREADS) cnt[i] = 0u;
    if (tid == 0) *s_flag = 0;
    __syncthreads();
    for (int e = tid; e < cn; e += SW_THREADS) atomicAdd(&cnt[floor_bin_of_ord((unsigned)(kA[e] >> 32))], 1u);
    __syncthreads();
    if (warp == 0) {
        unsigned c[8];
        unsigned mine = 0, mx = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) { c[q] = cnt[lane * 8 + q]; mine += c[q]; mx = c[q] > mx ? c[q] : mx; }
        unsigned suffix = mine;              // keys in this lane's bins and above
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_down_sync(0xffffffffu, suffix, o);
            if (lane + o < 32) suffix += v;
        }
        unsigned pos = suffix - mine;        // keys in higher bins: first position of this lane's top bin
#pragma unroll
        for (int q = 7; q >= 0; --q) { fill[lane * 8 + q] = pos; pos += c[q]; }
        if (__any_sync(0xffffffffu, mx > (unsigned)SW_BIN_MAX) && lane == 0) *s_flag = 1;
    }
    __syncthreads();
    if (*s_flag) return false;
    // (cnt keeps the populations; the start of a bin is recovered as fill-after-scatter minus its population)
    for (int e = tid; e < cn; e += SW_THREADS) {
        const unsigned long long k = kA[e];
        kB[atomicAdd(&fill[floor_bin_of_ord((unsigned)(k >> 32))], 1u)] = k;
    }
    __syncthreads();
    for (int e = tid; e < cn; e += SW_THREADS) {
        const unsigned long long k = kB[e];
        const int bin = floor_bin_of_ord((unsigned)(k >> 32));
        const int pop = (int)cnt[bin], s0 = (int)fill[bin] - pop;
        int before = 0;
        for (int j = 0; j < pop; ++j) before += kB[s0 + j] > k;
        kA[s0 + before] = k;
    }
    __syncthreads();
    return true;
}

template <typename IouT, bool TF, int SW_THREADS>
__global__ void __launch_bounds__(SW_THREADS, SW_THREADS == 256 ? 3 : 7)
sweep_kernel(unsigned long long* __restrict__ keys, const int* __restrict__ img_count, size_t img_stride,
             const float* __restrict__ y, DecodeArgs g, float conf_thr,
             int* __restrict__ g_floor, unsigned* __restrict__ g_hist, int* __restrict__ stats,
             double* __restrict__ pad_rows, int* __restrict__ pad_anchor, int* __restrict__ out_count) {
    constexpr int SW_WARPS = SW_THREADS / 32;
    constexpr int SW_PANEL = SW_THREADS;                         // candidates resolved per NMS panel
    constexpr int SW_PW = SW_PANEL / 32;
    extern __shared__ __align__(16) unsigned char sw_dyn[];      // cm[C][SW_PW] | km[C][SW_KW]
    __shared__ unsigned long long kA[SW_CHUNK];                  // slice as compacted (unsorted), sort scratch
    __shared__ unsigned long long kB[SW_CHUNK];                  // slice sorted descending
    __shared__ SBox<float> cbox[SW_CHUNK];                       // raw corners of the slice's boxes
    __shared__ unsigned long long kkey[SW_KMAX];                 // kept keys in keep order
    __shared__ SBox<float> kraw[SW_KMAX];                        // their raw corners
    __shared__ unsigned char kcls[SW_KMAX];
    __shared__ unsigned sup[SW_PW][SW_PANEL];                    // sup[w][i]: candidates 32w.. of the panel that suppress candidate i
    __shared__ unsigned hist[FL_BINS];
    __shared__ unsigned dead[SW_PW];
    __shared__ unsigned s_cnt, s_screen_off;
    __shared__ int s_nkept, s_done, s_bin;
    __shared__ unsigned long long s_tau;

    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = g.K, C = g.C;
    unsigned* cm = reinterpret_cast<unsigned*>(sw_dyn);          // per class: candidates of the panel
    unsigned* km = cm + (size_t)C * SW_PW;                        // per class: kept boxes (positions in keep order)
    unsigned long long* gk = keys + (size_t)b * img_stride;
    const float* yb = y + (size_t)b * g.A * g.W;                 // this image's rows of y_pred
    const IouT sx = (IouT)g.sx, sy = (IouT)g.sy, d = (IouT)g.d, thr = (IouT)g.iou_thr;
    const bool thr_ok = thr > IouT(0) && thr < IouT(INFINITY);
    const bool screen_ok = thr_ok && g.sx > 0.0 && g.sy > 0.0 && !TF;
    const bool f32_ok = screen_ok && g.d == 0.0 && g.iou_thr < 1e30;      // pair_f32 applies
    const float thr_lo = (float)g.iou_thr * (1.0f - 2e-5f), thr_hi = (float)g.iou_thr * (1.0f + 2e-5f);
    const unsigned lt = (1u << lane) - 1u;

    // The first reads of the kernel (count, floor, histogram, keys) do not depend on one another: the head of the image's
    // key list is copied into shared memory asynchronously (cp.async, no registers) while the count and the histogram are
    // on their way, so that the first compaction does not pay a second round trip.  Positions beyond the image's count
    // hold stale keys and are masked when the count is known.  The landing zone is `cbox`, which is idle until the slice's
    // boxes are decoded.
    constexpr int SW_PRE = SW_CHUNK * (int)sizeof(SBox<float>) / (int)sizeof(unsigned long long);      // 1024 keys
    unsigned long long* kpre = reinterpret_cast<unsigned long long*>(cbox);
    bool pre_valid = (img_stride & 1) == 0;                       // (16-byte alignment of every image's list)
    const int n_pre = (int)(img_stride < (size_t)SW_PRE ? img_stride & ~(size_t)1 : (size_t)SW_PRE);
    if (pre_valid) {
        for (int i = 2 * tid; i < n_pre; i += 2 * SW_THREADS)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(kpre + i)), "l"(gk + i) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    int n = img_count[b];
    int F = 0;                                                    // floor bin: keys of the bins >= F are complete
    bool use_hist = g.have_hist != 0;
    if (use_hist) {
        F = g_floor[b];
        for (int i = tid; i < FL_BINS; i += SW_THREADS) {
            unsigned* gh = g_hist + (size_t)b * FL_BINS + i;
            hist[i] = *gh;
            *gh = 0u;                                            // (left clean for the next decode)
        }
    }
    if (tid == 0) {
        s_screen_off = 0;
        // statistics of the decode (ssdc_decode_stats): keys D1 emitted, images whose score floor engaged
        atomicAdd(&stats[CNT_STAT_KEYS], n);
        if (F > 0) atomicAdd(&stats[CNT_STAT_FLOORED], 1);
    }
    __syncthreads();
    if (n == 0) { if (tid == 0) out_count[b] = 0; return; }

    for (int attempt = 0; attempt < 2; ++attempt) {
        if (attempt == 1) {
            // ---------------------------------------------------------------- exact fallback: rescan without a floor
            for (int i = tid; i < FL_BINS; i += SW_THREADS) hist[i] = 0u;
            if (tid == 0) { s_cnt = 0; atomicAdd(&stats[CNT_STAT_FALLBACK], 1); }
            __syncthreads();
            for (int a0 = 0; a0 < g.A; a0 += SW_THREADS) {
                const int a = a0 + tid;
                const float* row = yb + (size_t)(a < g.A ? a : 0) * g.W;
                for (int c0 = 0; c0 < g.NS; c0 += 32) {
                    const int nc = min(32, g.NS - c0);
                    unsigned mask = 0;
                    if (a < g.A)
                        for (int c = 0; c < nc; ++c) mask |= (unsigned)(row[1 + c0 + c] > conf_thr) << c;
                    if (!__any_sync(0xffffffffu, mask != 0u)) continue;
                    const int cnt = __popc(mask);
                    int incl = cnt;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int v = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += v;
                    }
                    unsigned base = 0;
                    if (lane == 31) base = atomicAdd(&s_cnt, (unsigned)incl);
                    base = __shfl_sync(0xffffffffu, base, 31);
                    unsigned long long* out = gk + base + (incl - cnt);
                    for (unsigned mm = mask; mm; mm &= mm - 1) {
                        const int c = __ffs(mm) - 1;
                        const unsigned ord = ord32(row[1 + c0 + c]);
                        *out++ = ((unsigned long long)ord << 32) | ((unsigned long long)(0xffu - (unsigned)(c0 + c + 1)) << 24) |
                                 (unsigned long long)(0xffffffu - (unsigned)a);
                        atomicAdd(&hist[floor_bin_of_ord(ord)], 1u);
                    }
                }
            }
            __syncthreads();
            n = (int)s_cnt;
            F = 0;
            use_hist = true;
        }
        const unsigned long long lo_key = F ? ((unsigned long long)floor_edge_ord(F) << 32) : 0ull;   // trusted keys: >= lo_key
        // per-class kept masks, kept count
        for (int i = tid; i < C * SW_KW; i += SW_THREADS) km[i] = 0u;
        if (tid == 0) s_nkept = 0;
        // trusted candidates not yet consumed
        int remaining = n;
        if (use_hist) {
            __syncthreads();
            int part = 0;
            for (int i = tid; i < FL_BINS; i += SW_THREADS) part += (i >= F) ? (int)hist[i] : 0;
            part = __reduce_add_sync(0xffffffffu, part);
            if (tid == 0) s_cnt = 0;
            __syncthreads();
            if (lane == 0) atomicAdd(&s_cnt, (unsigned)part);
            __syncthreads();
            remaining = (int)s_cnt;
        }
        __syncthreads();
        unsigned long long hi = ~0ull;              // keys >= hi are consumed
        int hi_bin = FL_BINS;                       // (histogram mode) bins >= hi_bin are consumed
        // First slice: a little more than top_k candidates - with little suppression that is all the sweep needs.
        int target = K + max(K >> 3, 8);
        if (target > SW_TARGET) target = SW_TARGET;
        while (remaining > 0) {
            // ------------------------------------------------------------ select the next slice [tau, hi)
            unsigned long long tau = lo_key;
            int tau_bin = F;
            if (use_hist) {
                if (warp == 0) {
                    // walk the bins [F, hi_bin) from the top: lane l owns bins [8l, 8l+8)
                    int cntq[8], mine = 0;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int bin = lane * 8 + q;
                        cntq[q] = (bin >= F && bin < hi_bin) ? (int)hist[bin] : 0;
                        mine += cntq[q];
                    }
                    int suffix = mine;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int v = __shfl_down_sync(0xffffffffu, suffix, o);
                        if (lane + o < 32) suffix += v;
                    }
                    const int above = suffix - mine;
                    const bool has = above < target && suffix >= target;
                    const unsigned hm = __ballot_sync(0xffffffffu, has);
                    if (hm ? has : (lane == 0)) {          // (no lane: fewer than `target` left -> everything down to F)
                        int sel = F, cum = hm ? above : suffix;
                        if (hm) {
#pragma unroll
                            for (int q = 7; q >= 0; --q) {
                                cum += cntq[q];
                                if (cum >= target) { sel = lane * 8 + q; break; }
                            }
                        }
                        s_bin = sel;
                        s_cnt = (unsigned)cum;              // candidates of the slice
                    }
                }
                __syncthreads();
                tau_bin = s_bin;
                const int cnt = (int)s_cnt;
                __syncthreads();
                if (cnt > SW_CHUNK) use_hist = false;       // too many candidates inside one bin: refine by radix passes
                else tau = tau_bin > 0 ? ((unsigned long long)floor_edge_ord(tau_bin) << 32) : 0ull;
                if (tau < lo_key) tau = lo_key;
            }
            if (!use_hist && remaining > SW_CHUNK) {
                // 8-bit radix passes over the image's keys in [lo_key, hi): the bucket holding the target-th best
                unsigned long long prefix = 0, pmask = 0;   // decided high bits of the threshold
                int above = 0;                              // candidates above the bucket being refined
                const int tgt = max(target, SW_TARGET);
                for (int shift = 56; shift >= 0; shift -= 8) {
                    unsigned* rh = &sup[0][0];              // 256 counters (sup is free during the selection)
                    for (int i = tid; i < 256; i += SW_THREADS) rh[i] = 0u;
                    __syncthreads();
                    for (int i0 = tid; i0 < n; i0 += 8 * SW_THREADS) {
                        unsigned long long k4[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int i = i0 + u * SW_THREADS;
                            k4[u] = (i < n) ? gk[i] : ~0ull;                 // (~0: never below `hi`)
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            if (k4[u] < hi && k4[u] >= lo_key && (k4[u] & pmask) == prefix) atomicAdd(&rh[(unsigned)(k4[u] >> shift) & 255u], 1u);
                    }
                    __syncthreads();
                    if (warp == 0) {
                        int mine = 0;
#pragma unroll
                        for (int q = 0; q < 8; ++q) mine += (int)rh[lane * 8 + q];
                        int suffix = mine;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int v = __shfl_down_sync(0xffffffffu, suffix, o);
                            if (lane + o < 32) suffix += v;
                        }
                        const int cum_before = above + suffix - mine;
                        const bool has = (cum_before < tgt) && (cum_before + mine >= tgt);
                        const unsigned hm = __ballot_sync(0xffffffffu, has);
                        if (hm ? has : (lane == 0)) {
                            int cum = cum_before, dsel = lane * 8;
                            for (int q = 7; q >= 0; --q) {
                                const int h = (int)rh[lane * 8 + q];
                                if (cum + h >= tgt || q == 0) { dsel = lane * 8 + q; break; }
                                cum += h;
                            }
                            s_cnt = (unsigned)cum;                        // strictly above the bucket
                            s_tau = prefix | ((unsigned long long)dsel << shift);
                            s_done = (cum + (int)rh[dsel] <= SW_CHUNK) || shift == 0;
                        }
                    }
                    __syncthreads();
                    prefix = s_tau;
                    pmask |= 0xffull << shift;
                    above = (int)s_cnt;
                    const int fin = s_done;
                    __syncthreads();
                    if (fin) break;
                }
                tau = prefix > lo_key ? prefix : lo_key;    // lower edge of the selected bucket
            }
            // ------------------------------------------------------------ compact {tau <= key < hi} into shared memory
            if (tid == 0) s_cnt = 0;
            if (pre_valid) asm volatile("cp.async.wait_all;" ::: "memory");
            __syncthreads();
            for (int i0 = 0; i0 < n; i0 += 8 * SW_THREADS) {
                unsigned long long k4[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = i0 + u * SW_THREADS + tid;
                    k4[u] = (i < n) ? ((pre_valid && i < n_pre) ? kpre[i] : gk[i]) : ~0ull;      // (~0: never below `hi`)
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const unsigned long long k = k4[u];
                    const bool take = k >= tau && k < hi;
                    const unsigned m = __ballot_sync(0xffffffffu, take);
                    if (m) {
                        unsigned base = 0;
                        if (lane == 0) base = atomicAdd(&s_cnt, (unsigned)__popc(m));
                        base = __shfl_sync(0xffffffffu, base, 0);
                        const unsigned pos = base + __popc(m & lt);
                        if (take && pos < SW_CHUNK) {
                            kA[pos] = k;
                            // the box is decoded from its y_pred row after the sort: pull the row towards L2 now
                            const float* rt = yb + (size_t)ck_anchor(k) * g.W + g.C;
                            asm volatile("prefetch.global.L2 [%0];" :: "l"(rt));
                            asm volatile("prefetch.global.L2 [%0];" :: "l"(rt + 11));
                        }
                    }
                }
            }
            __syncthreads();
            pre_valid = false;                      // (the key list may be rewritten by the rescan; later slices reload)
            int cn = (int)s_cnt;
            __syncthreads();
            if (cn > SW_CHUNK) {
                // (only possible when equal keys' multiplicity defeats the radix selection's last digit: cannot happen,
                // keys are unique; kept as a guard so that shared memory is never overrun)
                cn = SW_CHUNK;
            }
            if (cn == 0) {
                if (use_hist && tau_bin > F) { hi = tau; hi_bin = tau_bin; continue; }     // (empty bins above: keep walking)
                break;
            }
            // ------------------------------------------------------------ sort, decode the boxes
            unsigned long long* ks = kA;                   // the slice, sorted descending
            if (!bin_sort<SW_THREADS>(kA, kB, cn, &sup[0][0], &s_done)) { chunk_sort<SW_THREADS>(kA, kB, cn); ks = kB; }
            for (int i = tid; i < cn; i += SW_THREADS) {
                unsigned anc = ck_anchor(ks[i]);
                if (anc >= (unsigned)g.A || ck_cls(ks[i]) >= C || ck_cls(ks[i]) < 1) {       // invariant: every staged key is a real candidate
                    atomicCAS(&stats[CNT_STAT_ERR], 0, 1 + (ck_cls(ks[i]) >= C || ck_cls(ks[i]) < 1));
                    anc = 0; ks[i] = (ks[i] & 0xffffffff00000000ull) | (0xfeull << 24) | 0xffffffull;
                }
                const SBox<float> bx = decode_box<float>(yb + (size_t)anc * g.W, g.C, g);
                cbox[i] = bx;
                if (screen_ok && !box_regular(scale_box<float, IouT>(bx, sx, sy, d))) s_screen_off = 1;
            }
            __syncthreads();
            const bool screen = screen_ok && !s_screen_off;
            const bool f32 = f32_ok && !s_screen_off;
            // ------------------------------------------------------------ NMS panels
            for (int p0 = 0; p0 < cn; p0 += SW_PANEL) {
                const int P = min(SW_PANEL, cn - p0);
                const int nkept = s_nkept;
                if (nkept >= K) break;
                for (int i = tid; i < C * SW_PW; i += SW_THREADS) cm[i] = 0u;
                if (tid < SW_PW) dead[tid] = 0u;
                __syncthreads();
                for (int i = tid; i < P; i += SW_THREADS) atomicOr(&cm[(size_t)ck_cls(ks[p0 + i]) * SW_PW + (i >> 5)], 1u << (i & 31));
                __syncthreads();
                // Work units of 32 lanes: (row r of the panel, word w <= r of the panel) = which earlier candidates of word w
                // would suppress candidate 32 r + lane, and (row r, word of the kept list) = does a kept box suppress it.
                // Row r has r + 1 + kwords units - dealt to the warps round robin, every warp gets the same share (with
                // thread <-> candidate the last warp of the panel would do eight times the first one's work).
                {
                    const int R = (P + 31) >> 5;
                    const int kwords = (nkept + 31) >> 5;
                    int ubase = 0;                                  // units of the rows before r
                    for (int r = 0; r < R; ++r) {
                        const int i = (r << 5) + lane;
                        const bool act = i < P;
                        const int nu = r + 1 + kwords;              // units of row r; this warp takes ubase + w == warp (mod SW_WARPS)
                        const int w0 = (warp - ubase) & (SW_WARPS - 1);
                        ubase += nu;
                        if (!act) continue;
                        const int cls = ck_cls(ks[p0 + i]);
                        const SBox<float> me = cbox[p0 + i];
                        for (int w = w0; w < nu; w += SW_WARPS) {
                            if (w <= r) {
                                // earlier candidates of the panel with the same class, word w
                                unsigned rem = cm[(size_t)cls * SW_PW + w];
                                if (w == r) rem &= lt;
                                unsigned sw = 0;
                                for (; rem; rem &= rem - 1) {
                                    const int bit = __ffs(rem) - 1;
                                    const SBox<float> ob = cbox[p0 + (w << 5) + bit];
                                    if (f32) {
                                        const int rr = pair_f32(me, ob, thr_lo, thr_hi);
                                        if (rr == 0) continue;
                                        if (rr == 1) { sw |= 1u << bit; continue; }
                                    } else if (screen && raw_disjoint(me, ob)) continue;
                                    if (decide_pair<float, IouT, TF>(ob, me, sx, sy, d, thr, thr_ok)) sw |= 1u << bit;
                                }
                                sup[w][i] = sw;
                            } else {
                                // kept boxes of the same class, word w - r - 1 of the keep list
                                const int kw = w - r - 1;
                                bool is_dead = false;
                                for (unsigned rem = km[(size_t)cls * SW_KW + kw]; rem; rem &= rem - 1) {
                                    const int k = (kw << 5) + __ffs(rem) - 1;
                                    const SBox<float> kr = kraw[k];
                                    if (f32) {
                                        const int rr = pair_f32(me, kr, thr_lo, thr_hi);
                                        if (rr == 0) continue;
                                        if (rr == 1) { is_dead = true; break; }
                                    } else if (screen && raw_disjoint(me, kr)) continue;
                                    if (decide_pair<float, IouT, TF>(kr, me, sx, sy, d, thr, thr_ok)) { is_dead = true; break; }
                                }
                                if (is_dead) atomicOr(&dead[r], 1u << lane);
                            }
                        }
                    }
                }
                __syncthreads();
                if (warp == 0) {
                    // greedy resolution of the panel, 32 candidates at a time
                    unsigned keptw[SW_PW];
#pragma unroll
                    for (int w = 0; w < SW_PW; ++w) keptw[w] = 0u;
                    int nk = nkept;
#pragma unroll
                    for (int gi = 0; gi < SW_PW; ++gi) {
                        if (gi * 32 < P && nk < K) {
                            const int i = gi * 32 + lane;
                            const bool valid = i < P;
                            bool alive = valid && !((dead[gi] >> lane) & 1u);
                            unsigned mysup = 0;
                            if (alive) {
#pragma unroll
                                for (int w = 0; w < SW_PW; ++w)
                                    if (w < gi && (sup[w][i] & keptw[w])) alive = false;
                                if (alive) mysup = sup[gi][i];
                            }
                            const unsigned am = __ballot_sync(0xffffffffu, alive);
                            mysup &= am;
                            const unsigned nz = __ballot_sync(0xffffffffu, mysup != 0u);
                            unsigned keptm = am & ~nz;
                            for (unsigned rem = nz; rem; rem &= rem - 1) {
                                const int c = __ffs(rem) - 1;
                                const unsigned sc = __shfl_sync(0xffffffffu, mysup, c);
                                if (!(sc & keptm)) keptm |= 1u << c;
                            }
                            const int room = K - nk;
                            const bool mine = ((keptm >> lane) & 1u) && (__popc(keptm & lt) < room);
                            keptm = __ballot_sync(0xffffffffu, mine);
                            if (mine) {
                                const int pos = nk + __popc(keptm & lt);
                                const unsigned long long key = ks[p0 + i];
                                const int cls = ck_cls(key);
                                kraw[pos] = cbox[p0 + i];
                                kcls[pos] = (unsigned char)cls;
                                kkey[pos] = key;
                                atomicOr(&km[(size_t)cls * SW_KW + (pos >> 5)], 1u << (pos & 31));
                            }
                            keptw[gi] = keptm;
                            nk += __popc(keptm);
                        }
                    }
                    if (lane == 0) s_nkept = nk;
                }
                __syncthreads();
            }
            remaining -= cn;
            if (s_nkept >= K) break;
            hi = tau; hi_bin = tau_bin;
            target = SW_TARGET;
            if (tau <= lo_key) break;                   // every trusted candidate has been consumed
        }
        __syncthreads();
        // ran dry inside the trusted set although D1 dropped candidates below the floor: do it again on all of them
        if (!(s_nkept < K && F > 0)) break;
        __syncthreads();
    }

    // ------------------------------------------------------------ final order, rows written in place
    // The image's rows go to pad_rows[b, pos, :] = [class, conf, xmin, ymin, xmax, ymax] (float64, the layout of the
    // reference's DecodeDetections layer output without the zero padding); packing for the host copy happens at
    // collect time.
    const int nkept = s_nkept;
    const bool sweep_order = nkept >= K || g.always_sort;       // truncated (or layer mode): descending score, the sweep order
    for (int i = tid; i < nkept; i += SW_THREADS) {
        int pos = i;
        const int c = kcls[i];
        if (!sweep_order) {
            // nothing truncated: classes ascending, inside a class the keep order (:212-218)
            pos = 0;
            for (int j = 0; j < nkept; ++j) {
                const int cj = kcls[j];
                pos += (cj < c) || (cj == c && j < i);
            }
        }
        const unsigned long long k = kkey[i];
        const SBox<float> s = kraw[i];
        double* o = pad_rows + ((size_t)b * K + pos) * 6;
        o[0] = (double)c;
        o[1] = (double)unord32((uint32_t)(k >> 32));
        o[2] = (double)((IouT)s.x0 * sx);
        o[3] = (double)((IouT)s.y0 * sy);
        o[4] = (double)((IouT)s.x1 * sx);
        o[5] = (double)((IouT)s.y1 * sy);
        pad_anchor[(size_t)b * K + pos] = (int)ck_anchor(k);
    }
    if (tid == 0) out_count[b] = nkept;
}

// exclusive scan of per-image row counts -> packed row offsets (single CTA)
__global__ void __launch_bounds__(1024)
scan_counts_kernel(const int* __restrict__ out_count, int B, long long* __restrict__ row_offset) {
    __shared__ long long warp_sums[32];
    __shared__ long long carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < B; base += 1024) {
        const int b = base + tid;
        const long long c = (b < B) ? out_count[b] : 0;
        long long x = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long yv = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += yv;
        }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        if (warp == 0) {
            long long w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                long long yv = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += yv;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const long long before = carry + (warp ? warp_sums[warp - 1] : 0) + (x - c);
        if (b < B) row_offset[b] = before;
        __syncthreads();
        if (tid == 1023) carry = before + c;
        __syncthreads();
    }
    if (tid == 0) row_offset[B] = carry;
}

// ---------------------------------------------------------------------------
// D4: per-image totals + scan, cross-class top-k, row output
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
count_scan_kernel(const int* __restrict__ kept_count, int B, int NS, int K,
                  int* __restrict__ out_count, long long* __restrict__ row_offset) {
    __shared__ long long warp_sums[32];
    __shared__ long long carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < B; base += 1024) {
        const int b = base + tid;
        long long c = 0;
        if (b < B) {
            long long t = 0;
            for (int s = 0; s < NS; ++s) t += kept_count[(size_t)b * NS + s];
            c = (K > 0 && t > K) ? K : t;
            out_count[b] = (int)c;
        }
        long long x = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long yv = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += yv;
        }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        if (warp == 0) {
            long long w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                long long yv = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += yv;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const long long before = carry + (warp ? warp_sums[warp - 1] : 0) + (x - c);
        if (b < B) row_offset[b] = before;
        __syncthreads();
        if (tid == 1023) carry = before + c;
        __syncthreads();
    }
    if (tid == 0) row_offset[B] = carry;
}

template <typename KeyT> __device__ __forceinline__ double score_from_bits(uint64_t bits);
template <> __device__ __forceinline__ double score_from_bits<Key64>(uint64_t bits) { return (double)unord32((uint32_t)bits); }
template <> __device__ __forceinline__ double score_from_bits<Key128>(uint64_t bits) { return unord64(bits); }

template <typename StoreT, typename IouT>
__device__ __forceinline__ void write_row(double* __restrict__ rows, int* __restrict__ anchors, long long r,
                                          int cls, double score, uint32_t anchor,
                                          const SBox<StoreT>* __restrict__ bx, IouT sx, IouT sy) {
    SBox<StoreT> s = bx[anchor];
    double* o = rows + r * 6;
    o[0] = (double)cls;
    o[1] = score;
    o[2] = (double)((IouT)s.x0 * sx);
    o[3] = (double)((IouT)s.y0 * sy);
    o[4] = (double)((IouT)s.x1 * sx);
    o[5] = (double)((IouT)s.y1 * sy);
    anchors[r] = (int)anchor;
}

// image-sweep path, collect time: padded rows -> packed rows
__global__ void __launch_bounds__(128)
sweep_pack_kernel(const double* __restrict__ pad_rows, const int* __restrict__ pad_anchor, const int* __restrict__ out_count,
                  const long long* __restrict__ row_offset, int K, double* __restrict__ rows, int* __restrict__ anchors) {
    const int b = blockIdx.x;
    const int cnt = out_count[b];
    const long long off = row_offset[b];
    const double* src = pad_rows + (size_t)b * K * 6;
    for (int e = threadIdx.x; e < cnt * 6; e += 128) rows[off * 6 + e] = src[e];
    for (int r = threadIdx.x; r < cnt; r += 128) anchors[off + r] = pad_anchor[(size_t)b * K + r];
}

constexpr int EMIT_THREADS = 256;
constexpr int EMIT_SMEM_KEYS = 4096;

template <typename StoreT, typename IouT, typename KeyT>
__global__ void __launch_bounds__(EMIT_THREADS)
emit_kernel(const KeyT* __restrict__ keys, const int* __restrict__ kept_count,
            const int* __restrict__ out_count, const long long* __restrict__ row_offset,
            const SBox<StoreT>* __restrict__ boxes, const int* __restrict__ aux_class, DecodeArgs g,
            Key128* __restrict__ merge_scratch, int merge_stride,
            double* __restrict__ rows, int* __restrict__ anchors) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int b = blockIdx.x;
    const int NS = g.NS;
    const int cnt = out_count[b];
    if (cnt == 0) return;
    const long long off = row_offset[b];
    const SBox<StoreT>* bx = boxes + (size_t)b * g.A;
    const int* cls_of = aux_class ? aux_class + (size_t)b * g.A : nullptr;
    const IouT sx = (IouT)g.sx, sy = (IouT)g.sy;

    // total kept over the image's segments
    long long T = 0;
    for (int s = 0; s < NS; ++s) T += kept_count[(size_t)b * NS + s];

    if (T <= cnt && !g.always_sort) {
        // no truncation: classes ascending, inside a class the NMS keep order
        // (ssd_output_decoder.py:212-218)
        long long base = 0;
        for (int s = 0; s < NS; ++s) {
            const int k = kept_count[(size_t)b * NS + s];
            const KeyT* kp = keys + ((size_t)b * NS + s) * g.A;
            for (int r = threadIdx.x; r < k; r += EMIT_THREADS) {
                KeyT key = kp[r];
                uint32_t anchor = key.anchor();
                int cls = (NS > 1) ? (s + 1) : cls_of[anchor];
                write_row<StoreT, IouT>(rows, anchors, off + base + r, cls, key.score(), anchor, bx, sx, sy);
            }
            base += k;
        }
        return;
    }

    // truncation to top_k (ssd_output_decoder.py:219-221) or layer mode: order all kept boxes by
    // (score desc, class asc, anchor asc) and emit the first `cnt`.
    const int N = pow2_ceil((int)T);
    Key128* sk = (N <= EMIT_SMEM_KEYS) ? reinterpret_cast<Key128*>(smem_raw)
                                       : merge_scratch + (size_t)b * merge_stride;
    // segment prefix (NS may exceed the block size: computed serially by thread 0 in chunks)
    long long base = 0;
    for (int s = 0; s < NS; ++s) {
        const int k = kept_count[(size_t)b * NS + s];
        const KeyT* kp = keys + ((size_t)b * NS + s) * g.A;
        for (int r = threadIdx.x; r < k; r += EMIT_THREADS) {
            KeyT key = kp[r];
            uint32_t anchor = key.anchor();
            int cls = (NS > 1) ? (s + 1) : cls_of[anchor];
            Key128 ck;
            ck.hi = key.score_bits();
            ck.lo = ((uint64_t)(0xffffffffu - (uint32_t)cls) << 32) | (uint64_t)(0xffffffffu - anchor);
            sk[base + r] = ck;
        }
        base += k;
    }
    for (int i = (int)T + threadIdx.x; i < N; i += EMIT_THREADS) sk[i] = Key128::lowest();
    __syncthreads();
    block_bitonic_sort(sk, N, false);
    for (int r = threadIdx.x; r < cnt; r += EMIT_THREADS) {
        Key128 ck = sk[r];
        int cls = (int)(0xffffffffu - (uint32_t)(ck.lo >> 32));
        uint32_t anchor = 0xffffffffu - (uint32_t)ck.lo;
        write_row<StoreT, IouT>(rows, anchors, off + r, cls, score_from_bits<KeyT>(ck.hi), anchor, bx, sx, sy);
    }
}

// ---- D4 (main path): one warp per image, k-way merge of the per-class keep lists --------------
// Every class list is already in canonical order (score desc, anchor asc), so the cross-class
// top-k (ssd_output_decoder.py:219-221) is a k-way merge that stops after `cnt` rows: lane l owns
// the lists l, l+32, ..; each round a warp-wide max picks the next row.  Ties between classes go
// to the lower class id, then the lower anchor (composite key).  The lists are staged in shared
// memory when they fit.
constexpr int MERGE_Q_MAX = 4;        // lists per lane => up to 128 segments per image

template <typename KeyT> struct Comp;
template <> struct Comp<Key64> {       // [score bits 32 | ~class 8 | ~anchor 24]
    typedef uint64_t type;
    __device__ __forceinline__ static type make(const Key64& k, int cls) {
        return (k.v & 0xffffffff00000000ull) | ((uint64_t)(0xffu - (uint32_t)cls) << 24) | (uint64_t)(0xffffffu - k.anchor());
    }
    __device__ __forceinline__ static type lowest() { return 0; }
    __device__ __forceinline__ static bool gt(type a, type b) { return a > b; }
    __device__ __forceinline__ static bool eq(type a, type b) { return a == b; }
    __device__ __forceinline__ static type warp_max(type a) {
        // two 32-bit REDUX steps instead of a 64-bit shuffle tree
        const uint32_t hi = __reduce_max_sync(0xffffffffu, (uint32_t)(a >> 32));
        const uint32_t lo = __reduce_max_sync(0xffffffffu, ((uint32_t)(a >> 32) == hi) ? (uint32_t)a : 0u);
        return ((uint64_t)hi << 32) | lo;
    }
    __device__ __forceinline__ static int cls(type a) { return (int)(0xffu - (uint32_t)((a >> 24) & 0xffu)); }
    __device__ __forceinline__ static uint32_t anchor(type a) { return 0xffffffu - (uint32_t)(a & 0xffffffu); }
    __device__ __forceinline__ static double score(type a) { return (double)unord32((uint32_t)(a >> 32)); }
};
template <> struct Comp<Key128> {      // hi = score bits, lo = [~class 32 | ~anchor 32]
    typedef Key128 type;
    __device__ __forceinline__ static type make(const Key128& k, int cls) {
        Key128 c; c.hi = k.hi; c.lo = ((uint64_t)(0xffffffffu - (uint32_t)cls) << 32) | (uint64_t)(0xffffffffu - k.anchor()); return c;
    }
    __device__ __forceinline__ static type lowest() { return Key128::lowest(); }
    __device__ __forceinline__ static bool gt(const type& a, const type& b) { return (a.hi > b.hi) || (a.hi == b.hi && a.lo > b.lo); }
    __device__ __forceinline__ static bool eq(const type& a, const type& b) { return a.hi == b.hi && a.lo == b.lo; }
    __device__ __forceinline__ static type warp_max(type best) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const type other = shfl_xor_key(best, o);
            if (gt(other, best)) best = other;
        }
        return best;
    }
    __device__ __forceinline__ static int cls(const type& a) { return (int)(0xffffffffu - (uint32_t)(a.lo >> 32)); }
    __device__ __forceinline__ static uint32_t anchor(const type& a) { return 0xffffffffu - (uint32_t)a.lo; }
    __device__ __forceinline__ static double score(const type& a) { return unord64(a.hi); }
};

template <typename StoreT, typename IouT, typename KeyT, bool SMEM, int MERGE_Q>
__global__ void __launch_bounds__(32)
emit_merge_kernel(const KeyT* __restrict__ keys, const int* __restrict__ kept_count,
                  const int* __restrict__ out_count, const long long* __restrict__ row_offset,
                  const SBox<StoreT>* __restrict__ boxes, const int* __restrict__ aux_class, DecodeArgs g,
                  int Kcap, double* __restrict__ rows, int* __restrict__ anchors) {
    typedef Comp<KeyT> CK;
    typedef typename CK::type ck_t;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    KeyT* sk = reinterpret_cast<KeyT*>(smem_raw);
    const int b = blockIdx.x, lane = threadIdx.x;
    const int NS = g.NS;
    const int cnt = out_count[b];
    if (cnt == 0) return;
    const long long off = row_offset[b];
    const SBox<StoreT>* bx = boxes + (size_t)b * g.A;
    const int* cls_of = aux_class ? aux_class + (size_t)b * g.A : nullptr;
    const IouT sx = (IouT)g.sx, sy = (IouT)g.sy;
    const KeyT* kbase = keys + (size_t)b * NS * g.A;

    int len[MERGE_Q];
    int T = 0;
#pragma unroll
    for (int q = 0; q < MERGE_Q; ++q) {
        const int s = lane + 32 * q;
        len[q] = (s < NS) ? kept_count[(size_t)b * NS + s] : 0;
        T += len[q];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) T += __shfl_xor_sync(0xffffffffu, T, o);

    if ((T <= cnt && !g.always_sort) || NS == 1) {
        // no truncation: classes ascending, inside a class the NMS keep order (:212-218)
        int base = 0;
        for (int s = 0; s < NS; ++s) {
            int k = 0;
#pragma unroll
            for (int q = 0; q < MERGE_Q; ++q) if ((s >> 5) == q) k = len[q];
            k = __shfl_sync(0xffffffffu, k, s & 31);
            if (k > cnt - base) k = cnt - base;            // (NS == 1 and always_sort: list longer than top_k)
            const KeyT* kp = kbase + (size_t)s * g.A;
            for (int r = lane; r < k; r += 32) {
                const KeyT key = kp[r];
                const uint32_t anchor = key.anchor();
                const int cls = (NS > 1) ? (s + 1) : cls_of[anchor];
                write_row<StoreT, IouT>(rows, anchors, off + base + r, cls, key.score(), anchor, bx, sx, sy);
            }
            base += k;
        }
        return;
    }

    if (SMEM) {
        for (int s = 0; s < NS; ++s) {
            int k = 0;
#pragma unroll
            for (int q = 0; q < MERGE_Q; ++q) if ((s >> 5) == q) k = len[q];
            k = min(__shfl_sync(0xffffffffu, k, s & 31), Kcap);
            const KeyT* kp = kbase + (size_t)s * g.A;
            for (int r = lane; r < k; r += 32) sk[(size_t)s * Kcap + r] = kp[r];
        }
        __syncwarp();
    }
    auto fetch = [&](int s, int pos) -> KeyT {
        return SMEM ? sk[(size_t)s * Kcap + pos] : kbase[(size_t)s * g.A + pos];
    };
    int cur[MERGE_Q];
    ck_t head[MERGE_Q];
#pragma unroll
    for (int q = 0; q < MERGE_Q; ++q) {
        cur[q] = 0;
        head[q] = (len[q] > 0) ? CK::make(fetch(lane + 32 * q, 0), lane + 32 * q + 1) : CK::lowest();
    }
    ck_t outk = CK::lowest();
    for (int r = 0; r < cnt; ++r) {
        ck_t mine = head[0];
#pragma unroll
        for (int q = 1; q < MERGE_Q; ++q) if (CK::gt(head[q], mine)) mine = head[q];
        const ck_t best = CK::warp_max(mine);
        if ((r & 31) == lane) outk = best;
        if (CK::eq(mine, best)) {                       // keys are unique: exactly one lane advances
#pragma unroll
            for (int q = 0; q < MERGE_Q; ++q) {
                if (CK::eq(head[q], best)) {
                    ++cur[q];
                    head[q] = (cur[q] < len[q]) ? CK::make(fetch(lane + 32 * q, cur[q]), lane + 32 * q + 1) : CK::lowest();
                }
            }
        }
        if ((r & 31) == 31 || r == cnt - 1) {
            if (lane <= (r & 31))
                write_row<StoreT, IouT>(rows, anchors, off + (r & ~31) + lane, CK::cls(outk), CK::score(outk),
                                        CK::anchor(outk), bx, sx, sy);
        }
    }
}

// ---------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------
static float float_round_down(double x) {
    float f = (float)x;
    if ((double)f > x) f = nextafterf(f, -INFINITY);
    return f;
}
static float float_round_up(double x) {
    float f = (float)x;
    if ((double)f < x) f = nextafterf(f, INFINITY);
    return f;
}

struct IntLayout {
    size_t seg_count, floor, kept_count, lists, counters, total_ints;
};
static IntLayout int_layout(size_t nseg) {
    IntLayout L;
    L.counters = 0;                       // 32 ints, zeroed together with seg_count and floor
    L.seg_count = 32;
    L.floor = L.seg_count + nseg;         // image sweep: per-image score floor bins (first B entries)
    L.kept_count = L.floor + nseg;
    L.lists = L.kept_count + nseg;
    L.total_ints = L.lists + (size_t)NBINS * nseg;
    return L;
}

static bool d1_tma_ok(const void* y_dev, const DecodeArgs& g, size_t elem) {
    const size_t row_bytes = (size_t)g.W * elem;
    return (reinterpret_cast<uintptr_t>(y_dev) % 16 == 0) && (((size_t)g.A * row_bytes) % 16 == 0) &&
           (((size_t)g.tile_rows * row_bytes) % 16 == 0) && ((((size_t)g.A % g.tile_rows) * row_bytes) % 16 == 0);
}

template <typename InT>
static int launch_d1(ssdc_ctx* ctx, DevCtx* d, const InT* y_dev, const DecodeArgs& g, int64_t B, InT thr, bool fast,
                     int* seg_count, typename KeyOf<InT>::type* keys, SBox<InT>* boxes, int* aux,
                     int* g_floor = nullptr, unsigned* g_hist = nullptr, int max_ctas_per_sm = 4) {
    cudaStream_t st = d->stream;
    constexpr int V = 16 / (int)sizeof(InT);
    LaunchScope ls(ctx, d, SSDC_K_DECODE_FILTER);
    const size_t row_bytes = (size_t)g.W * sizeof(InT);
    if (d1_tma_ok(y_dev, g, sizeof(InT))) {
        const size_t stage_bytes = (((size_t)g.tile_rows * row_bytes) + 127) & ~(size_t)127;
        const int ncw = (g.tile_rows + 31) / 32;               // consumer warps: one row per thread
        const unsigned threads = (unsigned)(ncw + 1) * 32;
        const size_t smem = stage_bytes * D1_STAGES + (size_t)ncw * 2 * D1_PEND_HALF * sizeof(unsigned long long) +
                            2 * sizeof(FloorSlot);
        int ctas_per_sm = (int)((226 * 1024) / (smem + 1024 + 256));
        if (ctx->opt[SSDC_OPT_D1_CTAS] > 0) max_ctas_per_sm = (int)ctx->opt[SSDC_OPT_D1_CTAS];
        if (ctas_per_sm > max_ctas_per_sm) ctas_per_sm = max_ctas_per_sm;
        if (ctas_per_sm < 1) ctas_per_sm = 1;
#ifdef SSDC_TIMING_KNOBS
        if (const char* e = getenv("SSDC_D1_CTAS")) ctas_per_sm = atoi(e);      // (timing experiments only; not in release builds)
#endif
        const long long total_tiles = (long long)B * g.tiles;
        long long grid = (long long)d->sm_count * ctas_per_sm;
        if (grid > total_tiles) grid = total_tiles;
        if (fast) {
            SSDC_TRY(ensure_dyn_smem(d->device, (const void*)decode_filter_tma_kernel<InT, true>, smem));
            decode_filter_tma_kernel<InT, true><<<(unsigned)grid, threads, smem, st>>>(y_dev, g, thr, (int)total_tiles, seg_count, keys, boxes, aux, g_floor, g_hist);
        } else {
            SSDC_TRY(ensure_dyn_smem(d->device, (const void*)decode_filter_tma_kernel<InT, false>, smem));
            decode_filter_tma_kernel<InT, false><<<(unsigned)grid, threads, smem, st>>>(y_dev, g, thr, (int)total_tiles, seg_count, keys, boxes, aux, g_floor, g_hist);
        }
        SSDC_TRY(check_launch("decode_filter_tma_kernel"));
    } else {
        size_t smem = (((size_t)g.tile_rows * g.W + V) * sizeof(InT) + 15) & ~(size_t)15;
        dim3 grid((unsigned)((size_t)B * g.tiles));
        if (fast) {
            SSDC_CUDA(cudaFuncSetAttribute(decode_filter_kernel<InT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            decode_filter_kernel<InT, true><<<grid, D1_THREADS, smem, st>>>(y_dev, g, thr, seg_count, keys, boxes, aux);
        } else {
            SSDC_CUDA(cudaFuncSetAttribute(decode_filter_kernel<InT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            decode_filter_kernel<InT, false><<<grid, D1_THREADS, smem, st>>>(y_dev, g, thr, seg_count, keys, boxes, aux);
        }
        SSDC_TRY(check_launch("decode_filter_kernel"));
    }
    return SSDC_OK;
}

// ---------------------------------------------------------------------------
// Host input.  The batch is copied in chunks of whole images on the copy stream (stream2) and D1 of a chunk is
// enqueued on the main stream behind that chunk's copy: the filter of chunk k runs under the copy of chunk k + 1, the
// rest of the pipeline starts when the last chunk has been filtered.  A pageable source (a plain numpy array - what
// model.predict() returns) is staged through a ring of pinned buffers that a few host threads fill while the previous
// chunk is on the wire; a pinned source is copied from directly.
// ---------------------------------------------------------------------------
struct HostFeed {
    const char* src = nullptr;      // nullptr: the batch is already on the device
    size_t img_bytes = 0;
};

static int host_threads(const ssdc_ctx* ctx) {
    static int cached = 0;
    if (!cached) {
        int n = 0;
        if (const char* e = getenv("SSDC_STAGE_THREADS")) n = atoi(e);
        if (n <= 0) {
            cpu_set_t set;
            CPU_ZERO(&set);
            int cpus = sched_getaffinity(0, sizeof(set), &set) == 0 ? CPU_COUNT(&set) : (int)std::thread::hardware_concurrency();
            n = cpus / 2;
        }
        cached = n < 1 ? 1 : (n > 8 ? 8 : n);
    }
    const int nd = (int)ctx->devs.size();
    const int t = cached / (nd > 0 ? nd : 1);
    return t < 1 ? 1 : t;
}

// Host threads that stage pageable inputs into pinned memory.  They live as long as the process (started at the first
// pageable decode input, never joined: a worker only ever sleeps on the condition variable or copies bytes the caller is
// waiting for), so a small batch does not pay a thread start per call - B = 8: 1.5 ms per call with threads spawned per chunk.
class CopyPool {
public:
    static constexpr int MAX_WORKERS = 7;
    static CopyPool& get() { static CopyPool* p = new CopyPool(); return *p; }     // (intentionally never destroyed)
    void run(char* dst, const char* src, size_t n, int threads) {
        std::lock_guard<std::mutex> one(call_mu_);                               // one copy at a time (contexts on other threads wait)
        if (threads > MAX_WORKERS + 1) threads = MAX_WORKERS + 1;
        const size_t per = ((n / (size_t)threads) + 4095) & ~(size_t)4095;
        int used = 0;
        {
            std::lock_guard<std::mutex> lk(mu_);
            while ((int)workers_.size() < threads - 1) {
                const int id = (int)workers_.size();
                workers_.emplace_back([this, id] { work(id); });
                workers_.back().detach();
            }
            for (int t = 1; t < threads; ++t) {
                const size_t off = per * (size_t)t;
                if (off >= n) break;
                jobs_[t - 1] = Job{dst + off, src + off, n - off < per ? n - off : per};
                ++used;
            }
            for (int t = used; t < MAX_WORKERS; ++t) jobs_[t] = Job{nullptr, nullptr, 0};
            pending_ = used;
            ++gen_;
        }
        if (used) cv_work_.notify_all();
        memcpy(dst, src, per < n ? per : n);
        if (used) {
            std::unique_lock<std::mutex> lk(mu_);
            cv_done_.wait(lk, [this] { return pending_ == 0; });
        }
    }
private:
    struct Job { char* dst; const char* src; size_t n; };
    void work(int id) {
        unsigned long long seen = 0;
        {
            std::lock_guard<std::mutex> lk(mu_);
            seen = gen_ - 1;                   // (started inside run(), before the generation it was started for is published)
        }
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_work_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                j = jobs_[id];
            }
            if (j.n == 0) continue;
            memcpy(j.dst, j.src, j.n);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--pending_ == 0) cv_done_.notify_one();
            }
        }
    }
    std::mutex call_mu_, mu_;
    std::condition_variable cv_work_, cv_done_;
    std::vector<std::thread> workers_;
    Job jobs_[MAX_WORKERS] = {};
    unsigned long long gen_ = 0;
    int pending_ = 0;
};

static void parallel_memcpy(char* dst, const char* src, size_t n, int threads) {
    if (threads <= 1 || n < ((size_t)1 << 20)) { memcpy(dst, src, n); return; }
    CopyPool::get().run(dst, src, n, threads);
}

// Calls filter(b0, nb) for consecutive image ranges covering [0, B), each enqueued behind the arrival of its images.
template <typename F>
static int feed_chunks(ssdc_ctx* ctx, DevCtx* d, const HostFeed& feed, char* y_dev, int64_t B, F&& filter) {
    if (!feed.src) return filter((int64_t)0, B);
    int64_t chunk_mb = ctx->opt[SSDC_OPT_H2D_CHUNK_MB];
    if (chunk_mb == 0) chunk_mb = 64;
    const size_t total = (size_t)B * feed.img_bytes;
    cudaPointerAttributes attr;
    bool pageable = true;
    if (cudaPointerGetAttributes(&attr, feed.src) == cudaSuccess) pageable = attr.type == cudaMemoryTypeUnregistered;
    else cudaGetLastError();
    // images per chunk: whole images, chunk starts 16-byte aligned (the TMA loader needs it)
    int64_t per = B;
    if (chunk_mb > 0 && feed.img_bytes % 16 == 0) {
        size_t chunk_bytes = (size_t)chunk_mb << 20;
        // a small pageable batch is staged by ONE host thread (below): four chunks, so that the staging of chunk k + 1 runs
        // under the copy and the filter pass of chunk k.  (Larger batches keep whole 64 MB chunks: D1's score floor needs
        // CTAs that own several tiles of an image.)
        if (pageable && total < ((size_t)16 << 20)) chunk_bytes = total / 4 > ((size_t)1 << 20) ? total / 4 : ((size_t)1 << 20);
        per = (int64_t)(chunk_bytes / feed.img_bytes);
        if (per < 1) per = 1;
        if (per * 2 > B && !pageable) per = B;                 // (fewer than two chunks: one copy)
    }
    if (per >= B && !pageable) {
        SSDC_CUDA(cudaMemcpyAsync(y_dev, feed.src, total, cudaMemcpyHostToDevice, d->stream));
        return filter((int64_t)0, B);
    }
    if (per > B) per = B;
    // staging threads by chunk size: waking sleeping workers costs more than copying a few megabytes alone (measured on the
    // GPU host, 9.2 MB: 1.47 ms per call with 8 threads, 0.83 ms with one; 147 MB: 6.5 ms with 8, 11.4 ms with one)
    int threads = 1;
    if (pageable) {
        const size_t cb = (size_t)per * feed.img_bytes;
        const int t_max = host_threads(ctx);
        threads = cb >= ((size_t)48 << 20) ? t_max : cb >= ((size_t)24 << 20) ? 4 : cb >= ((size_t)12 << 20) ? 2 : 1;
        if (threads > t_max) threads = t_max;
    }
    cudaStream_t cs = d->stream2;
    int k = 0;
    for (int64_t b0 = 0; b0 < B; b0 += per, ++k) {
        const int64_t nb = B - b0 < per ? B - b0 : per;
        const size_t off = (size_t)b0 * feed.img_bytes, bytes = (size_t)nb * feed.img_bytes;
        const char* from = feed.src + off;
        if (pageable) {
            const int s = k % DevCtx::FEED_RING;
            if (d->feed_busy[s]) { SSDC_CUDA(cudaEventSynchronize(d->feed_ev[s])); d->feed_busy[s] = false; }
            SSDC_TRY(d->feed_buf[s].ensure((size_t)per * feed.img_bytes));
            parallel_memcpy(d->feed_buf[s].as<char>(), from, bytes, threads);
            from = d->feed_buf[s].as<char>();
            SSDC_CUDA(cudaMemcpyAsync(y_dev + off, from, bytes, cudaMemcpyHostToDevice, cs));
            if (!d->feed_ev[s]) SSDC_CUDA(cudaEventCreateWithFlags(&d->feed_ev[s], cudaEventDisableTiming));
            SSDC_CUDA(cudaEventRecord(d->feed_ev[s], cs));
            d->feed_busy[s] = true;
        } else {
            SSDC_CUDA(cudaMemcpyAsync(y_dev + off, from, bytes, cudaMemcpyHostToDevice, cs));
        }
        cudaEvent_t& ev = d->chunk_ev[k % 8];
        if (!ev) SSDC_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        SSDC_CUDA(cudaEventRecord(ev, cs));
        SSDC_CUDA(cudaStreamWaitEvent(d->stream, ev, 0));
        SSDC_TRY(filter(b0, nb));
    }
    return SSDC_OK;
}

// Image sweep path (decode_detections / DecodeDetections layer with a finite top_k, float32 input).
template <typename InT, typename IouT, bool TF>
static int run_sweep(ssdc_ctx* ctx, DevCtx* d, const InT* y_dev, const HostFeed& feed, DecodeArgs g, int64_t B, InT thr, bool pipe) {
    if (sizeof(InT) != 4) { set_error("internal: sweep path needs float32 input"); return SSDC_ERR_STATE; }
    typedef typename KeyOf<InT>::type KeyT;
    IntLayout L = int_layout((size_t)g.nseg);
    int* ints = d->ints.as<int>();
    int* img_count = ints + L.seg_count;            // one counter per image (first B entries)
    int* g_floor = ints + L.floor;                  // one floor bin per image (zeroed with the counters)
    cudaStream_t st = d->stream;
    // Score histograms + floor: only with the TMA loader, an image of >= D1_STAGES tiles and a positive target
    int64_t target = ctx->opt[SSDC_OPT_FLOOR_TARGET];
    // default: five times top_k, at least 1024 - the sweep reaches top_k kept boxes inside the trusted set unless NMS
    // suppresses more than ~4 of 5 candidates in an image with more candidates than that (then: exact rescan)
    if (target == 0) target = 5LL * g.K > 1024 ? 5LL * g.K : 1024;
    if (target < 2LL * g.K + 64) target = 2LL * g.K + 64;            // (never tighter than what one sweep slice asks for)
    if (target > 0x3fffffff) target = 0x3fffffff;
    g.floor_target = (int)target;
    g.have_hist = d1_tma_ok(y_dev, g, sizeof(InT)) && g.tiles >= D1_STAGES && ctx->opt[SSDC_OPT_FLOOR_TARGET] >= 0;
    if (g.have_hist) {
        const size_t hist_bytes = (size_t)B * FL_BINS * sizeof(unsigned);
        if (hist_bytes > d->hist.cap) d->hist_clean = false;
        SSDC_TRY(d->hist.ensure(hist_bytes));
        if (!d->hist_clean) SSDC_CUDA(cudaMemsetAsync(d->hist.p, 0, d->hist.cap, st));
        d->hist_clean = false;                      // (dirty until the sweep kernel, which leaves it zeroed, has been enqueued)
    }
    SSDC_TRY(feed_chunks(ctx, d, feed, reinterpret_cast<char*>(const_cast<InT*>(y_dev)), B, [&](int64_t b0, int64_t nb) -> int {
        return launch_d1<InT>(ctx, d, y_dev + (size_t)b0 * g.A * g.W, g, nb, thr, false, img_count + b0,
                              d->keys.as<KeyT>() + (size_t)b0 * g.NS * g.A, d->boxes.as<SBox<InT>>(), nullptr,
                              g_floor + b0, d->hist.as<unsigned>() + (size_t)b0 * FL_BINS);
    }));
    // Pipelined (device-resident input): the sweep goes to the low-priority side stream behind this D1 and may run beside
    // D1 of the NEXT decode; the scratch both touch is this decode's own bank.  D1 keeps its full occupancy (measured:
    // giving up a CTA per SM so that sweep CTAs fit beside it costs D1 20 % at B = 1024 - more than the overlap returns),
    // so what overlaps is the launch latency, D1's ramp and tail and the sweep's: 2 % at B = 1024, 10 % at B = 128,
    // 35 % at B = 32.
    cudaStream_t ns = pipe ? d->stream_nms : st;
    if (pipe) {
        SSDC_CUDA(cudaEventRecord(d->ev_d1[d->bank], st));
        SSDC_CUDA(cudaStreamWaitEvent(ns, d->ev_d1[d->bank], 0));
    }
    SSDC_TRY(d->pad_rows.ensure((size_t)B * g.K * 6 * sizeof(double)));
    SSDC_TRY(d->pad_anchor.ensure((size_t)B * g.K * sizeof(int)));
    {
        LaunchScope ls(ctx, d, SSDC_K_NMS);
        // one image per CTA: 256 threads while every image of the batch is resident at once, else 128 threads (7 CTAs per
        // SM: B = 1024 on 148 SMs is one wave)
        const bool narrow = B > 3LL * d->sm_count;
        const size_t dyn = (size_t)g.C * ((narrow ? 128 : 256) / 32 + SW_KW) * sizeof(unsigned);
        if (narrow) {
            SSDC_TRY(ensure_dyn_smem(d->device, (const void*)sweep_kernel<IouT, TF, 128>, dyn));
            sweep_kernel<IouT, TF, 128><<<(unsigned)B, 128, dyn, ns>>>(
                d->keys.as<unsigned long long>(), img_count, (size_t)g.NS * g.A, reinterpret_cast<const float*>(y_dev), g, (float)thr,
                g_floor, d->hist.as<unsigned>(), ints + L.counters, d->pad_rows.as<double>(), d->pad_anchor.as<int>(), d->out_count.as<int>());
        } else {
            SSDC_TRY(ensure_dyn_smem(d->device, (const void*)sweep_kernel<IouT, TF, 256>, dyn));
            sweep_kernel<IouT, TF, 256><<<(unsigned)B, 256, dyn, ns>>>(
                d->keys.as<unsigned long long>(), img_count, (size_t)g.NS * g.A, reinterpret_cast<const float*>(y_dev), g, (float)thr,
                g_floor, d->hist.as<unsigned>(), ints + L.counters, d->pad_rows.as<double>(), d->pad_anchor.as<int>(), d->out_count.as<int>());
        }
        SSDC_TRY(check_launch("sweep_kernel"));
    }
    if (pipe) {
        SSDC_CUDA(cudaEventRecord(d->ev_sweep[d->bank], ns));
        d->sweep_pending[d->bank] = true;
    }
    if (g.have_hist) d->hist_clean = true;
    // (the packed row offsets are only needed for the host copy: scan_counts_kernel runs at collect time)
    return SSDC_OK;
}

template <typename InT, typename IouT, bool TF>
static int run_pipeline(ssdc_ctx* ctx, DevCtx* d, const InT* y_dev, const HostFeed& feed, const DecodeArgs& g, int64_t B,
                        double conf_thresh, int cmp_f32_rn, bool pipe) {
    typedef typename KeyOf<InT>::type KeyT;
    const bool fast = (g.NS == 1);
    const size_t nseg = (size_t)g.nseg;
    IntLayout L = int_layout(nseg);
    int* ints = d->ints.as<int>();
    int* counters = ints + L.counters;
    int* seg_count = ints + L.seg_count;
    int* kept_count = ints + L.kept_count;
    int* lists = ints + L.lists;
    KeyT* keys = d->keys.as<KeyT>();
    SBox<InT>* boxes = d->boxes.as<SBox<InT>>();
    int* aux = fast ? d->aux_class.as<int>() : nullptr;
    cudaStream_t st = d->stream;

    // typed threshold (hard part 1): the reference compares float64(conf) with
    // the Python float for the centroid / minmax paths and float32 with float32(thr) for 'corners'
    // and the Keras layers.  For float32 inputs the float64 comparison is folded into an
    // equivalent float32 one: x > t  <=>  x > round_down_f32(t);  x >= t  <=>  x >= round_up_f32(t).
    InT thr;
    if (sizeof(InT) == 4) {
        if (cmp_f32_rn) thr = (InT)(float)conf_thresh;
        else thr = (InT)(g.ge ? float_round_up(conf_thresh) : float_round_down(conf_thresh));
    } else {
        thr = (InT)conf_thresh;
    }

    SSDC_CUDA(cudaMemsetAsync(ints, 0, (L.kept_count) * sizeof(int), st));
    if (g.sweep) return run_sweep<InT, IouT, TF>(ctx, d, y_dev, feed, g, B, thr, pipe);
    const int n1 = SORT_BYTES1 / (int)sizeof(KeyT), n2 = SORT_BYTES2 / (int)sizeof(KeyT), n3 = SORT_BYTES3 / (int)sizeof(KeyT);

    // D1
    SSDC_TRY(feed_chunks(ctx, d, feed, reinterpret_cast<char*>(const_cast<InT*>(y_dev)), B, [&](int64_t b0, int64_t nb) -> int {
        return launch_d1<InT>(ctx, d, y_dev + (size_t)b0 * g.A * g.W, g, nb, thr, fast, seg_count + (size_t)b0 * g.NS,
                              keys + (size_t)b0 * g.NS * g.A, boxes + (size_t)b0 * g.A, aux ? aux + (size_t)b0 * g.A : nullptr);
    }));
    // plan
    {
        LaunchScope ls(ctx, d, SSDC_K_PLAN);
        plan_kernel<<<(unsigned)((nseg + 255) / 256), 256, 0, st>>>(seg_count, (int)nseg, kept_count, lists, counters, NMS_REG_MAX, n1, n2, n3);
        SSDC_TRY(check_launch("plan_kernel"));
    }
    // D2: one persistent launch per size bin
    {
        const int sms = d->sm_count;
        struct BinCfg { int nmax, threads, ctas_per_sm; };
        const BinCfg cfg[4] = {{n1, 128, 8}, {n2, 512, 3}, {n3, 1024, 1}, {0, 1024, 1}};
        for (int k = 1; k <= 4; ++k) {
            const BinCfg& c = cfg[k - 1];
            if (k < 4 && (size_t)g.A <= (size_t)(k == 1 ? NMS_REG_MAX : cfg[k - 2].nmax)) continue;   // bin cannot occur
            if (k == 4 && g.A <= n3) continue;
            LaunchScope ls(ctx, d, SSDC_K_SORT);
            if (k < 4) {
                size_t smem = (size_t)c.nmax * sizeof(KeyT);
                SSDC_CUDA(cudaFuncSetAttribute(sort_kernel<KeyT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_BYTES3));
                int ctas = (int)((227 * 1024) / (smem + 1024));
                if (ctas > c.ctas_per_sm) ctas = c.ctas_per_sm;
                if (ctas < 1) ctas = 1;
                sort_kernel<KeyT, true><<<sms * ctas, c.threads, smem, st>>>(keys, seg_count, lists + (size_t)k * nseg, counters, k, g, nullptr, 0);
            } else {
                int stride = 1; while (stride < g.A) stride <<= 1;
                int grid = sms;
                if ((size_t)grid > nseg) grid = (int)nseg;
                SSDC_TRY(d->sort_scratch.ensure((size_t)grid * stride * sizeof(KeyT)));
                sort_kernel<KeyT, false><<<grid, c.threads, 0, st>>>(keys, seg_count, lists + (size_t)k * nseg, counters, k, g, d->sort_scratch.as<KeyT>(), stride);
            }
            SSDC_TRY(check_launch("sort_kernel"));
        }
    }
    // D3
    {
        int KS = 256;
        if (g.Kseg > 0 && g.Kseg < KS) KS = (g.Kseg + 31) & ~31;
        size_t smem = (size_t)NMS_WARPS * ((size_t)(KS + 32) * sizeof(SBox<InT>) + (size_t)NMS_SMEM_SORT * sizeof(KeyT) + (size_t)NMS_QUEUE * sizeof(unsigned short) + 32 * sizeof(unsigned));
        int ctas_per_sm = (int)((200 * 1024) / (smem + 1024));
        if (ctas_per_sm > 12) ctas_per_sm = 12;
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        size_t want = (nseg + NMS_WARPS - 1) / NMS_WARPS;
        size_t grid = (size_t)d->sm_count * ctas_per_sm;
        if (grid > want) grid = want;
        if (grid < 1) grid = 1;
        LaunchScope ls(ctx, d, SSDC_K_NMS);
        SSDC_CUDA(cudaFuncSetAttribute(nms_kernel<InT, IouT, KeyT, TF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        nms_kernel<InT, IouT, KeyT, TF><<<(unsigned)grid, NMS_WARPS * 32, smem, st>>>(
            keys, seg_count, kept_count, lists, counters, boxes, g, KS);
        SSDC_TRY(check_launch("nms_kernel"));
    }
    // D4a
    {
        LaunchScope ls(ctx, d, SSDC_K_MERGE);
        count_scan_kernel<<<1, 1024, 0, st>>>(kept_count, (int)B, g.NS, g.K, d->out_count.as<int>(), d->row_offset.as<long long>());
        SSDC_TRY(check_launch("count_scan_kernel"));
    }
    return SSDC_OK;
}

template <typename InT, typename IouT>
static int run_emit(ssdc_ctx* ctx, DevCtx* d, const DecodeArgs& g, int64_t B) {
    typedef typename KeyOf<InT>::type KeyT;
    IntLayout L = int_layout((size_t)g.nseg);
    int* ints = d->ints.as<int>();
    const bool fast = (g.NS == 1);
    // composite-key sort space: only needed when a truncating / always-sorting merge can exceed smem
    size_t per_image_max = (g.Kseg > 0) ? (size_t)g.NS * (size_t)min(g.Kseg, g.A) : (size_t)g.NS * g.A;
    int stride = 0;
    bool may_sort = g.always_sort || g.K > 0;
    if (may_sort && per_image_max > EMIT_SMEM_KEYS) {
        size_t s = 1; while (s < per_image_max) s <<= 1;
        stride = (int)s;
        SSDC_TRY(d->merge_scratch.ensure((size_t)B * s * sizeof(Key128)));
    }
    const int Kcap = (g.Kseg > 0) ? min(g.Kseg, g.A) : g.A;
    LaunchScope ls(ctx, d, SSDC_K_MERGE);
    if (g.sweep) {
        sweep_pack_kernel<<<(unsigned)B, 128, 0, d->stream>>>(
            d->pad_rows.as<double>(), d->pad_anchor.as<int>(), d->out_count.as<int>(), d->row_offset.as<long long>(), g.K,
            d->out_rows.as<double>(), d->out_anchor.as<int>());
        SSDC_TRY(check_launch("sweep_pack_kernel"));
        return SSDC_OK;
    }
    if (g.NS <= 32 * MERGE_Q_MAX && g.A < (1 << 24) && g.C <= 256) {
        // warp-per-image k-way merge
        const size_t list_bytes = may_sort ? (size_t)g.NS * Kcap * sizeof(KeyT) : 0;
        const bool use_smem = g.NS > 1 && list_bytes > 0 && list_bytes <= 100 * 1024;
        auto launch = [&](auto kern) -> int {
            if (use_smem) SSDC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)list_bytes));
            kern<<<(unsigned)B, 32, use_smem ? list_bytes : 0, d->stream>>>(
                d->keys.as<KeyT>(), ints + L.kept_count, d->out_count.as<int>(), d->row_offset.as<long long>(),
                d->boxes.as<SBox<InT>>(), fast ? d->aux_class.as<int>() : nullptr, g, Kcap,
                d->out_rows.as<double>(), d->out_anchor.as<int>());
            return SSDC_OK;
        };
        if (g.NS <= 32) {
            if (use_smem) SSDC_TRY(launch(emit_merge_kernel<InT, IouT, KeyT, true, 1>));
            else SSDC_TRY(launch(emit_merge_kernel<InT, IouT, KeyT, false, 1>));
        } else {
            if (use_smem) SSDC_TRY(launch(emit_merge_kernel<InT, IouT, KeyT, true, MERGE_Q_MAX>));
            else SSDC_TRY(launch(emit_merge_kernel<InT, IouT, KeyT, false, MERGE_Q_MAX>));
        }
        SSDC_TRY(check_launch("emit_merge_kernel"));
        return SSDC_OK;
    }
    size_t smem = (size_t)EMIT_SMEM_KEYS * sizeof(Key128);
    SSDC_CUDA(cudaFuncSetAttribute(emit_kernel<InT, IouT, KeyT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    emit_kernel<InT, IouT, KeyT><<<(unsigned)B, EMIT_THREADS, smem, d->stream>>>(
        d->keys.as<KeyT>(), ints + L.kept_count, d->out_count.as<int>(), d->row_offset.as<long long>(),
        d->boxes.as<SBox<InT>>(), fast ? d->aux_class.as<int>() : nullptr, g,
        d->merge_scratch.as<Key128>(), stride, d->out_rows.as<double>(), d->out_anchor.as<int>());
    SSDC_TRY(check_launch("emit_kernel"));
    return SSDC_OK;
}

static int build_args(const DecodeJob& job, bool no_sweep, DecodeArgs* out, int* iou_f32, int* tf, int* cmp_rn) {
    const ssdc_decode_params& p = job.p;
    DecodeArgs g;
    memset(&g, 0, sizeof(g));
    g.A = (int)job.A; g.C = job.C; g.W = job.C + 12;
    const bool layer = (p.mode == SSDC_MODE_LAYER || p.mode == SSDC_MODE_LAYER_FAST);
    const bool fast = (p.mode == SSDC_MODE_FAST || p.mode == SSDC_MODE_LAYER_FAST);
    g.NS = fast ? 1 : job.C - 1;
    g.nseg = (int)(job.B * g.NS);
    g.input_coords = p.input_coords;
    g.log_wh = p.log_wh;
    g.layer_assoc = layer ? 1 : 0;
    g.ge = (p.mode == SSDC_MODE_FAST) ? 1 : 0;      // `>=` only in decode_detections_fast (:325)
    g.do_nms = (fast && !layer) ? p.do_nms : 1;
    g.K = p.top_k > 0 ? p.top_k : 0;
    g.Kseg = layer ? p.nms_cap : g.K;
    g.always_sort = layer ? 1 : 0;
    g.iou_thr = p.iou_thresh;
    g.sx = p.normalize ? p.img_w : 1.0;
    g.sy = p.normalize ? p.img_h : 1.0;
    g.d = (p.border_pixels == SSDC_BORDER_INCLUDE) ? 1.0 : (p.border_pixels == SSDC_BORDER_EXCLUDE ? -1.0 : 0.0);
    // tile rows: as many whole rows as fit a comfortable shared-memory tile
    size_t elem = (job.dtype == SSDC_F32) ? 4 : 8;
    int rows = D1_THREADS;
    while (rows > 32 && ((size_t)rows * g.W + 4) * elem + 4096 > 100 * 1024) rows >>= 1;
    g.tile_rows = rows;
    g.tiles = (int)((job.A + rows - 1) / rows);
    // float32 input stays float32 end to end only where the reference never upcasts:
    // input_coords == 'corners' (ssd_output_decoder.py:186-190) and the Keras layers.
    // image sweep path: per-class semantics with a finite top_k that no per-class cap can undercut
    g.sweep = (job.dtype == SSDC_F32) && !fast && g.K > 0 && g.K <= SW_KMAX && job.C <= 256 && job.A < (1 << 24) &&
              (!layer || p.nms_cap >= p.top_k) && !no_sweep;
    *iou_f32 = (job.dtype == SSDC_F32) && (layer || p.input_coords == SSDC_COORDS_CORNERS);
    *tf = layer;
    *cmp_rn = *iou_f32;
    *out = g;
    return SSDC_OK;
}

int decode_submit_dev(ssdc_ctx* ctx, DevCtx* d, const void* y_pred, int dtype, int on_device,
                      int64_t b0, int64_t B, int64_t A, int C, const ssdc_decode_params* p) {
    SSDC_CUDA(cudaSetDevice(d->device));
    DecodeJob& job = d->job;
    job.valid = false;
    job.b0 = b0; job.B = B; job.A = A; job.C = C; job.dtype = dtype; job.p = *p; job.emitted = false;
    if (B == 0) { job.valid = true; job.NS = 0; return SSDC_OK; }
    DecodeArgs g; int iou_f32, tf, cmp_rn;
    SSDC_TRY(build_args(job, ctx->opt[SSDC_OPT_NO_SWEEP] != 0, &g, &iou_f32, &tf, &cmp_rn));
    g.evict_first = g.sweep && ctx->opt[SSDC_OPT_NO_L2_HINTS] == 0;
    if (ctx->opt[SSDC_OPT_D1_WARPS] >= 1 && ctx->opt[SSDC_OPT_D1_WARPS] * 32 < g.tile_rows) {      // (diagnosis: smaller tiles, fewer consumer warps per CTA)
        g.tile_rows = (int)ctx->opt[SSDC_OPT_D1_WARPS] * 32;
        g.tiles = (int)((job.A + g.tile_rows - 1) / g.tile_rows);
    }      // (measured on the image-sweep path: headline 0.229 -> 0.222 ms)
    job.NS = g.NS; job.iou_f32 = iou_f32;
    const size_t elem = (dtype == SSDC_F32) ? 4 : 8;
    const size_t in_bytes = (size_t)B * A * g.W * elem;
    const void* y_dev = y_pred;
    HostFeed feed;
    if (!on_device) {
        SSDC_TRY(d->y_in.ensure(in_bytes));
        y_dev = d->y_in.p;
        feed.src = reinterpret_cast<const char*>(y_pred);           // copied chunk by chunk, D1 behind every chunk (feed_chunks)
        feed.img_bytes = (size_t)A * g.W * elem;
    }
    // consecutive device-resident image-sweep decodes are pipelined: this one takes the other scratch bank and only waits
    // for the sweep that used it two decodes ago; everything else waits for every sweep still in flight (it shares the
    // current bank, or overwrites the staged input the sweep decodes its boxes from)
    const bool pipe = g.sweep && on_device && !ctx->profile && ctx->opt[SSDC_OPT_NO_PIPELINE] == 0;
    if (on_device) SSDC_TRY(d->wait_encodes());            // (the input may be what an encode still in flight on a lane writes)
    if (pipe) { d->swap_banks(); SSDC_TRY(d->wait_sweeps(d->bank)); }
    else SSDC_TRY(d->wait_sweeps());
    const size_t nseg = (size_t)g.nseg;
    IntLayout L = int_layout(nseg);
    SSDC_TRY(d->ints.ensure(L.total_ints * sizeof(int)));
    SSDC_TRY(d->keys.ensure(nseg * (size_t)A * (dtype == SSDC_F32 ? sizeof(Key64) : sizeof(Key128))));
    SSDC_TRY(d->boxes.ensure((size_t)B * A * 4 * elem));
    if (g.NS == 1) SSDC_TRY(d->aux_class.ensure((size_t)B * A * sizeof(int)));
    SSDC_TRY(d->out_count.ensure((size_t)B * sizeof(int)));
    SSDC_TRY(d->row_offset.ensure((size_t)(B + 1) * sizeof(long long)));

    int r;
    if (dtype == SSDC_F32) {
        if (!iou_f32) r = run_pipeline<float, double, false>(ctx, d, (const float*)y_dev, feed, g, B, p->conf_thresh, cmp_rn, pipe);
        else if (!tf) r = run_pipeline<float, float, false>(ctx, d, (const float*)y_dev, feed, g, B, p->conf_thresh, cmp_rn, pipe);
        else r = run_pipeline<float, float, true>(ctx, d, (const float*)y_dev, feed, g, B, p->conf_thresh, cmp_rn, pipe);
    } else {
        r = run_pipeline<double, double, false>(ctx, d, (const double*)y_dev, feed, g, B, p->conf_thresh, cmp_rn, pipe);
    }
    SSDC_TRY(r);

    job.scan_pending = g.sweep != 0;
    job.padded = g.sweep != 0;
    if (g.K > 0 && !g.sweep) {
        // bounded output: rows can be emitted right away without knowing the total
        job.out_capacity = B * (int64_t)g.K;
        SSDC_TRY(d->out_rows.ensure((size_t)job.out_capacity * 6 * sizeof(double)));
        SSDC_TRY(d->out_anchor.ensure((size_t)job.out_capacity * sizeof(int)));
        if (dtype == SSDC_F32) {
            if (!iou_f32) r = run_emit<float, double>(ctx, d, g, B);
            else r = run_emit<float, float>(ctx, d, g, B);
        } else r = run_emit<double, double>(ctx, d, g, B);
        SSDC_TRY(r);
        job.emitted = true;
    }
    job.valid = true;
    return SSDC_OK;
}

// Waits for the device, returns the number of result rows of this shard.
int decode_finish_dev(ssdc_ctx* ctx, DevCtx* d, int64_t* total_rows) {
    DecodeJob& job = d->job;
    *total_rows = 0;
    if (!job.valid) { set_error("ssdc_decode_collect without a submitted decode"); return SSDC_ERR_STATE; }
    if (job.B == 0) return SSDC_OK;
    SSDC_CUDA(cudaSetDevice(d->device));
    SSDC_TRY(d->wait_sweeps());
    if (job.scan_pending) {
        LaunchScope ls(ctx, d, SSDC_K_MERGE);
        scan_counts_kernel<<<1, 1024, 0, d->stream>>>(d->out_count.as<int>(), (int)job.B, d->row_offset.as<long long>());
        SSDC_TRY(check_launch("scan_counts_kernel"));
        job.scan_pending = false;
    }
    long long total = 0;
    SSDC_CUDA(cudaMemcpyAsync(&total, d->row_offset.as<long long>() + job.B, sizeof(long long), cudaMemcpyDeviceToHost, d->stream));
    int st3[4] = {0, 0, 0, 0};
    if (job.padded) SSDC_CUDA(cudaMemcpyAsync(st3, d->ints.as<int>() + CNT_STAT_KEYS, sizeof(st3), cudaMemcpyDeviceToHost, d->stream));
    SSDC_CUDA(cudaStreamSynchronize(d->stream));
    *total_rows = total;
    job.stat_keys = st3[0]; job.stat_floored = st3[1]; job.stat_fallback = st3[2];
    if (st3[3] != 0) { set_error("decode: internal invariant %d of the image sweep violated", st3[3]); return SSDC_ERR_STATE; }
    return SSDC_OK;
}

// top_k == 'all': the row buffer is sized once the total is known, then rows are emitted.
int decode_emit_all_dev(ssdc_ctx* ctx, DevCtx* d, int64_t total_rows) {
    DecodeJob& job = d->job;
    if (job.emitted || job.B == 0) return SSDC_OK;
    SSDC_CUDA(cudaSetDevice(d->device));
    DecodeArgs g; int iou_f32, tf, cmp_rn;
    SSDC_TRY(build_args(job, ctx->opt[SSDC_OPT_NO_SWEEP] != 0, &g, &iou_f32, &tf, &cmp_rn));
    job.out_capacity = total_rows;
    SSDC_TRY(d->out_rows.ensure((size_t)(total_rows + 1) * 6 * sizeof(double)));
    SSDC_TRY(d->out_anchor.ensure((size_t)(total_rows + 1) * sizeof(int)));
    int r;
    if (job.dtype == SSDC_F32) {
        if (!iou_f32) r = run_emit<float, double>(ctx, d, g, job.B);
        else r = run_emit<float, float>(ctx, d, g, job.B);
    } else r = run_emit<double, double>(ctx, d, g, job.B);
    SSDC_TRY(r);
    job.emitted = true;
    return SSDC_OK;
}

// ---------------------------------------------------------------------------
// standalone greedy NMS (greedy_nms / _greedy_nms / _greedy_nms2, ssd_output_decoder.py:27-109)
// ---------------------------------------------------------------------------
__global__ void nms_prep_kernel(const double* __restrict__ boxes, const double* __restrict__ scores, int n,
                                int coords, Key128* __restrict__ keys, SBox<double>* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double c4[4] = {boxes[4 * i], boxes[4 * i + 1], boxes[4 * i + 2], boxes[4 * i + 3]};
    SBox<double> b;
    to_corners(c4, coords, &b.x0, &b.y0, &b.x1, &b.y1);
    out[i] = b;
    keys[i] = Key128::make(scores[i], (uint32_t)i);
}

int greedy_nms_dev(ssdc_ctx* ctx, DevCtx* d, const double* boxes, const double* scores, int64_t n,
                   double iou_thresh, int coords, int border_pixels, int32_t* out_keep, int64_t* n_keep) {
    SSDC_CUDA(cudaSetDevice(d->device));
    *n_keep = 0;
    if (n == 0) return SSDC_OK;
    d->job.valid = false;            // shares the decode scratch: a pending decode is invalidated
    SSDC_TRY(d->wait_sweeps());
    cudaStream_t st = d->stream;
    DecodeArgs g;
    memset(&g, 0, sizeof(g));
    g.A = (int)n; g.C = 2; g.W = 14; g.NS = 1; g.nseg = 1; g.do_nms = 1; g.K = 0; g.Kseg = 0;
    g.iou_thr = iou_thresh; g.sx = 1.0; g.sy = 1.0;
    g.d = (border_pixels == SSDC_BORDER_INCLUDE) ? 1.0 : (border_pixels == SSDC_BORDER_EXCLUDE ? -1.0 : 0.0);
    IntLayout L = int_layout(1);
    SSDC_TRY(d->ints.ensure(L.total_ints * sizeof(int)));
    SSDC_TRY(d->keys.ensure((size_t)n * sizeof(Key128)));
    SSDC_TRY(d->boxes.ensure((size_t)n * sizeof(SBox<double>)));
    SSDC_TRY(d->t0buf.ensure((size_t)n * 4 * sizeof(double)));
    SSDC_TRY(d->t1buf.ensure((size_t)n * sizeof(double)));
    int* ints = d->ints.as<int>();
    int* counters = ints + L.counters;
    int* seg_count = ints + L.seg_count;
    int* kept_count = ints + L.kept_count;
    int* lists = ints + L.lists;
    Key128* keys = d->keys.as<Key128>();
    SSDC_CUDA(cudaMemcpyAsync(d->t0buf.p, boxes, (size_t)n * 4 * sizeof(double), cudaMemcpyHostToDevice, st));
    SSDC_CUDA(cudaMemcpyAsync(d->t1buf.p, scores, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
    SSDC_CUDA(cudaMemsetAsync(ints, 0, L.kept_count * sizeof(int), st));
    const int n32 = (int)n;
    SSDC_CUDA(cudaMemcpyAsync(seg_count, &n32, sizeof(int), cudaMemcpyHostToDevice, st));
    const int n1 = SORT_BYTES1 / (int)sizeof(Key128), n2 = SORT_BYTES2 / (int)sizeof(Key128), n3 = SORT_BYTES3 / (int)sizeof(Key128);
    {
        LaunchScope ls(ctx, d, SSDC_K_THIN);
        nms_prep_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d->t0buf.as<double>(), d->t1buf.as<double>(), n32, coords, keys, d->boxes.as<SBox<double>>());
        SSDC_TRY(check_launch("nms_prep_kernel"));
    }
    {
        LaunchScope ls(ctx, d, SSDC_K_PLAN);
        plan_kernel<<<1, 256, 0, st>>>(seg_count, 1, kept_count, lists, counters, NMS_REG_MAX, n1, n2, n3);
        SSDC_TRY(check_launch("plan_kernel"));
    }
    if (n > NMS_REG_MAX) {
        LaunchScope ls(ctx, d, SSDC_K_SORT);
        int bin = n > n3 ? 4 : (n > n2 ? 3 : (n > n1 ? 2 : 1));
        if (bin < 4) {
            size_t smem = (size_t)(bin == 1 ? SORT_BYTES1 : (bin == 2 ? SORT_BYTES2 : SORT_BYTES3));
            SSDC_CUDA(cudaFuncSetAttribute(sort_kernel<Key128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_BYTES3));
            sort_kernel<Key128, true><<<1, bin == 1 ? 128 : (bin == 2 ? 512 : 1024), smem, st>>>(keys, seg_count, lists + (size_t)bin, counters, bin, g, nullptr, 0);
        } else {
            int stride = 1; while (stride < g.A) stride <<= 1;
            SSDC_TRY(d->sort_scratch.ensure((size_t)stride * sizeof(Key128)));
            sort_kernel<Key128, false><<<1, 1024, 0, st>>>(keys, seg_count, lists + (size_t)bin, counters, bin, g, d->sort_scratch.as<Key128>(), stride);
        }
        SSDC_TRY(check_launch("sort_kernel"));
    }
    {
        const int KS = 256;
        size_t smem = (size_t)NMS_WARPS * ((size_t)(KS + 32) * sizeof(SBox<double>) + (size_t)NMS_SMEM_SORT * sizeof(Key128) + (size_t)NMS_QUEUE * sizeof(unsigned short) + 32 * sizeof(unsigned));
        LaunchScope ls(ctx, d, SSDC_K_NMS);
        SSDC_CUDA(cudaFuncSetAttribute(nms_kernel<double, double, Key128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        nms_kernel<double, double, Key128, false><<<1, NMS_WARPS * 32, smem, st>>>(keys, seg_count, kept_count, lists, counters, d->boxes.as<SBox<double>>(), g, KS);
        SSDC_TRY(check_launch("nms_kernel"));
    }
    int kept = 0;
    SSDC_CUDA(cudaMemcpyAsync(&kept, kept_count, sizeof(int), cudaMemcpyDeviceToHost, st));
    SSDC_CUDA(cudaStreamSynchronize(st));
    std::vector<Key128> hk((size_t)kept);
    if (kept > 0) {
        SSDC_CUDA(cudaMemcpyAsync(hk.data(), keys, (size_t)kept * sizeof(Key128), cudaMemcpyDeviceToHost, st));
        SSDC_CUDA(cudaStreamSynchronize(st));
    }
    for (int i = 0; i < kept; ++i) out_keep[i] = (int32_t)hk[i].anchor();
    *n_keep = kept;
    return SSDC_OK;
}

}  // namespace ssdc

extern "C" int ssdc_greedy_nms(ssdc_ctx* ctx, const double* boxes, const double* scores, int64_t n,
                               double iou_thresh, int coords, int border_pixels,
                               int32_t* out_keep, int64_t* n_keep) {
    using namespace ssdc;
    if (!ctx || n < 0 || !n_keep || (n > 0 && (!boxes || !scores || !out_keep)) || coords < 0 || coords > 2 ||
        border_pixels < 0 || border_pixels > 2 || n > 0x7fffffff) {
        set_error("ssdc_greedy_nms: bad argument");
        return SSDC_ERR_ARG;
    }
    std::lock_guard<std::mutex> lk(ctx->mu);
    return greedy_nms_dev(ctx, &ctx->devs[0], boxes, scores, n, iou_thresh, coords, border_pixels, out_keep, n_keep);
}
