// loss.cu - SSD multibox loss, forward pass, on the device (the consumer of the
// encoder's output; with `y_encoded` device-resident its 2.3 MB/image never cross PCIe).
//
// Replaces   SSDLoss.compute_loss   /root/reference/localisation_part/keras_loss_function/keras_ssd_loss.py:98-211
//            (smooth_L1_loss :53-76, log_loss :78-96)
//
// PARITY UNPINNED: the reference evaluates this in TensorFlow (tensorflow-gpu 1.8 / 1.14, not installable here).
// The arithmetic follows the TensorFlow graph op by op in float32 (the Keras placeholder dtype); only the
// reduction ORDER of `tf.reduce_sum` is unspecified there - sums are accumulated in float64 here and rounded
// once.  `tf.nn.top_k` takes the lower index first among equal values; so does the selection below.
//
// Kernels:
//   loss_box_tma_kernel  : streams y_true / y_pred once (warp-specialised TMA ring as in D1; loss_box_kernel: plain tile
//                          copies for unaligned inputs), thread = box:
//                          log loss, smooth L1, positive / negative masks; per-image partial sums, batch counters,
//                          per-box classification loss and negative loss for the mining step
//   loss_hist_kernel +   : 4 x 8-bit radix selection of the k-th largest negative loss over the whole batch
//   loss_pick_kernel       (hard negative mining, :161-191)
//   loss_eqcount_kernel, loss_quota_kernel, loss_negsum_kernel : per-image sums of the kept negatives; boxes
//                          that tie with the k-th value are admitted in flat index order
//   loss_final_kernel    : (class + alpha * loc) / max(1, n_positive) * batch_size   (:201-209)
#include "common.cuh"
#include "ctx.cuh"
#include <math.h>
#include <algorithm>

namespace ssdc {

constexpr int LS_ROWS = 128;

struct LossState {
    double n_pos;                      // sum of the positive masks over the batch (:143)
    unsigned long long n_nonzero;      // tf.count_nonzero(neg_class_loss_all) (:150)
    unsigned prefix, pmask;            // radix selection: decided high bits of the k-th largest key
    long long k_rem;                   // still to take inside the undecided bucket
    unsigned hist[256];
};

__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 ldg_stream_d2(const double2* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

// One tile of `rows` boxes staged in shared memory (row stride W): thread t evaluates box r0 + t.
template <typename TrueS>
__device__ __forceinline__ void loss_rows(const float* __restrict__ sp, const TrueS* __restrict__ stt, int tid, int rows, long long r0,
                                          long long n_boxes, int A, int C, int W, float* __restrict__ closs, float* __restrict__ nl,
                                          double* __restrict__ img_pos, double* __restrict__ img_loc, double& my_pos, unsigned long long& my_nz) {
    const int lane = tid & 31, warp = tid >> 5;
    float cl = 0.f, ll = 0.f, pos = 0.f, neg = 0.f;
    const long long i = r0 + tid;
    if (tid < rows) {
        const float* yp = sp + (size_t)tid * W;
        const TrueS* yt = stt + (size_t)tid * W;
        float acc = 0.f;
        for (int c = 0; c < C; ++c) {                                                  // :93-95
            // 0 * log(finite) adds an exact zero: the logarithm is only evaluated where it can matter (one class
            // of a one-hot target; a NaN / inf prediction still has to poison the sum like it does in TensorFlow)
            const float t = (float)yt[c], x = yp[c];
            if (t != 0.0f || !(fabsf(x) < INFINITY)) acc += t * logf(fmaxf(x, 1e-15f));
        }
        cl = -acc;
#pragma unroll
        for (int k = 0; k < 4; ++k) {                                                  // :72-75
            const float dlt = (float)yt[C + k] - yp[C + k];
            const float ab = fabsf(dlt);
            ll += (ab < 1.0f) ? (0.5f * (dlt * dlt)) : (ab - 0.5f);
        }
        neg = (float)yt[0];                                                            // :137
        pos = (C >= 2) ? (float)yt[1] : -INFINITY;
        for (int c = 2; c < C; ++c) pos = fmaxf(pos, (float)yt[c]);                    // :138
        const float nlv = cl * neg;                                                    // :149
        closs[i] = cl;
        nl[i] = nlv;
        my_pos += (double)pos;
        my_nz += (nlv != 0.0f) ? 1ull : 0ull;
    }
    // per-image sums: a warp's 32 boxes belong to one image except at an image boundary
    const long long b_first = __shfl_sync(0xffffffffu, i / A, 0);
    const long long i_last = min(r0 + warp * 32 + 31, n_boxes - 1);
    const bool one_image = (i_last / A) == b_first;
    const double pc = (tid < rows) ? (double)(cl * pos) : 0.0;                         // :147
    const double lc = (tid < rows) ? (double)(ll * pos) : 0.0;                         // :197
    if (one_image) {
        double a = pc, c2 = lc;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); c2 += __shfl_xor_sync(0xffffffffu, c2, o); }
        if (lane == 0 && r0 + warp * 32 < n_boxes) { atomicAdd(&img_pos[b_first], a); atomicAdd(&img_loc[b_first], c2); }
    } else if (tid < rows) {
        atomicAdd(&img_pos[i / A], pc); atomicAdd(&img_loc[i / A], lc);
    }
}

template <typename TrueT>
__global__ void __launch_bounds__(LS_ROWS)
loss_box_kernel(const TrueT* __restrict__ y_true, const float* __restrict__ y_pred, long long n_boxes, int A, int C, int W,
                float* __restrict__ closs, float* __restrict__ nl, double* __restrict__ img_pos, double* __restrict__ img_loc,
                LossState* __restrict__ st) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* sp = reinterpret_cast<float*>(smem_raw);             // LS_ROWS x W   predictions
    float* stt = sp + (size_t)LS_ROWS * W;                      // LS_ROWS x W   targets, float32 like the Keras placeholder
    __shared__ double red_pos[LS_ROWS / 32];
    __shared__ unsigned long long red_nz[LS_ROWS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double my_pos = 0.0;
    unsigned long long my_nz = 0;
    const long long n_tiles = (n_boxes + LS_ROWS - 1) / LS_ROWS;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long r0 = tile * LS_ROWS;
        const int rows = (int)min((long long)LS_ROWS, n_boxes - r0);
        const int n = rows * W;
        __syncthreads();
        {
            // (a tile starts at a multiple of LS_ROWS rows: 16-byte aligned for any W when the tensors are)
            const float* gp = y_pred + r0 * W;
            const TrueT* gt = y_true + r0 * W;
            const bool al = ((reinterpret_cast<uintptr_t>(gp) | reinterpret_cast<uintptr_t>(gt)) & 15) == 0;
            int done_p = 0, done_t = 0;
            if (al) {
                const int nv = n >> 2;
                const float4* g4 = reinterpret_cast<const float4*>(gp);
                for (int v = tid; v < nv; v += LS_ROWS) reinterpret_cast<float4*>(sp)[v] = ldg_stream_f4(g4 + v);
                done_p = nv << 2;
                if (sizeof(TrueT) == 4) {
                    const float4* t4 = reinterpret_cast<const float4*>(gt);
                    for (int v = tid; v < nv; v += LS_ROWS) reinterpret_cast<float4*>(stt)[v] = ldg_stream_f4(t4 + v);
                    done_t = nv << 2;
                } else {
                    const int nd = n >> 1;
                    const double2* t2 = reinterpret_cast<const double2*>(gt);
                    for (int v = tid; v < nd; v += LS_ROWS) {
                        const double2 t = ldg_stream_d2(t2 + v);
                        reinterpret_cast<float2*>(stt)[v] = make_float2((float)t.x, (float)t.y);
                    }
                    done_t = nd << 1;
                }
            }
            for (int e = done_p + tid; e < n; e += LS_ROWS) sp[e] = gp[e];
            for (int e = done_t + tid; e < n; e += LS_ROWS) stt[e] = (float)gt[e];
        }
        __syncthreads();
        loss_rows<float>(sp, stt, tid, rows, r0, n_boxes, A, C, W, closs, nl, img_pos, img_loc, my_pos, my_nz);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { my_pos += __shfl_xor_sync(0xffffffffu, my_pos, o); my_nz += __shfl_xor_sync(0xffffffffu, my_nz, o); }
    if (lane == 0) { red_pos[warp] = my_pos; red_nz[warp] = my_nz; }
    __syncthreads();
    if (tid == 0) {
        double p = 0.0; unsigned long long z = 0;
        for (int w = 0; w < LS_ROWS / 32; ++w) { p += red_pos[w]; z += red_nz[w]; }
        atomicAdd(&st->n_pos, p);
        atomicAdd(&st->n_nonzero, z);
    }
}

// ---- TMA variant of the box kernel: warp-specialised persistent CTAs, exactly the structure of D1
// (decode.cu): one producer warp keeps a 2-stage ring of tiles - LS_ROWS rows of y_pred and of y_true each, in
// their own dtypes - filled with cp.async.bulk + full / empty mbarriers; four consumer warps evaluate a tile as
// soon as it has landed.  Every CTA owns a contiguous range of tiles.  Needs 16-byte aligned tensors and
// n_boxes % 4 == 0 (else loss_box_kernel above).
__device__ __forceinline__ uint32_t ls_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ls_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(ls_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void ls_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(ls_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ls_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LS_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra LS_WAIT_DONE;\n"
        "bra LS_WAIT_LOOP;\n"
        "LS_WAIT_DONE:\n"
        "}\n" :: "r"(ls_smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
}
__device__ __forceinline__ void ls_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(ls_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ls_tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(ls_smem_u32(dst)), "l"(src), "r"(bytes), "r"(ls_smem_u32(bar)) : "memory");
}

constexpr int LS_STAGES = 2;
template <typename TrueT>
__global__ void __launch_bounds__(LS_ROWS + 32)
loss_box_tma_kernel(const TrueT* __restrict__ y_true, const float* __restrict__ y_pred, long long n_boxes, int A, int C, int W,
                    float* __restrict__ closs, float* __restrict__ nl, double* __restrict__ img_pos, double* __restrict__ img_loc,
                    LossState* __restrict__ st) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[LS_STAGES];
    __shared__ __align__(8) uint64_t empty[LS_STAGES];
    __shared__ double red_pos[LS_ROWS / 32];
    __shared__ unsigned long long red_nz[LS_ROWS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t pred_bytes = ((size_t)LS_ROWS * W * sizeof(float) + 127) & ~(size_t)127;
    const size_t true_bytes = ((size_t)LS_ROWS * W * sizeof(TrueT) + 127) & ~(size_t)127;
    const size_t stage_bytes = pred_bytes + true_bytes;
    if (tid == 0) {
        for (int s = 0; s < LS_STAGES; ++s) { ls_mbar_init(&full[s], 1); ls_mbar_init(&empty[s], LS_ROWS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long n_tiles = (n_boxes + LS_ROWS - 1) / LS_ROWS;
    const long long per = n_tiles / gridDim.x, extra = n_tiles - per * gridDim.x;
    const long long t_begin = (long long)blockIdx.x * per + min((long long)blockIdx.x, extra);
    const long long t_end = t_begin + per + ((long long)blockIdx.x < extra ? 1 : 0);
    if (warp == LS_ROWS / 32) {
        // ---- producer warp ----
        if (lane == 0) {
            int it = 0;
            for (long long t = t_begin; t < t_end; ++t, ++it) {
                const int s = it % LS_STAGES;
                if (it >= LS_STAGES) ls_mbar_wait(&empty[s], (uint32_t)(((it / LS_STAGES) - 1) & 1));
                const long long r0 = t * LS_ROWS;
                const int rows = (int)min((long long)LS_ROWS, n_boxes - r0);
                const uint32_t pb = (uint32_t)((size_t)rows * W * sizeof(float)), tb = (uint32_t)((size_t)rows * W * sizeof(TrueT));
                unsigned char* dst = smem_raw + (size_t)s * stage_bytes;
                ls_mbar_expect_tx(&full[s], pb + tb);
                ls_tma_load_1d(dst, y_pred + r0 * W, pb, &full[s]);
                ls_tma_load_1d(dst + pred_bytes, y_true + r0 * W, tb, &full[s]);
            }
        }
        return;
    }
    // ---- consumer warps ----
    double my_pos = 0.0;
    unsigned long long my_nz = 0;
    int it = 0;
    for (long long t = t_begin; t < t_end; ++t, ++it) {
        const int s = it % LS_STAGES;
        ls_mbar_wait(&full[s], (uint32_t)((it / LS_STAGES) & 1));
        const long long r0 = t * LS_ROWS;
        const int rows = (int)min((long long)LS_ROWS, n_boxes - r0);
        const unsigned char* src = smem_raw + (size_t)s * stage_bytes;
        loss_rows<TrueT>(reinterpret_cast<const float*>(src), reinterpret_cast<const TrueT*>(src + pred_bytes), tid, rows, r0,
                         n_boxes, A, C, W, closs, nl, img_pos, img_loc, my_pos, my_nz);
        __syncwarp();
        if (lane == 0) ls_mbar_arrive(&empty[s]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { my_pos += __shfl_xor_sync(0xffffffffu, my_pos, o); my_nz += __shfl_xor_sync(0xffffffffu, my_nz, o); }
    if (lane == 0) { red_pos[warp] = my_pos; red_nz[warp] = my_nz; }
    // (the producer warp has left: a named barrier over the consumer warps only)
    asm volatile("bar.sync 1, %0;" :: "n"(LS_ROWS) : "memory");
    if (tid == 0) {
        double p = 0.0; unsigned long long z = 0;
        for (int w = 0; w < LS_ROWS / 32; ++w) { p += red_pos[w]; z += red_nz[w]; }
        atomicAdd(&st->n_pos, p);
        atomicAdd(&st->n_nonzero, z);
    }
}

// histogram of the next 8 bits of the keys that match the decided prefix
__global__ void __launch_bounds__(256)
loss_hist_kernel(const float* __restrict__ nl, long long n, int shift, LossState* __restrict__ st) {
    __shared__ unsigned h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned prefix = st->prefix, pmask = st->pmask;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const unsigned k = ord32(nl[i]);
        if ((k & pmask) == prefix) atomicAdd(&h[(k >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], h[threadIdx.x]);
}

// the bucket that holds the k-th largest key: one warp walks the buckets from the top, lane l owns buckets [8l, 8l+8)
__global__ void __launch_bounds__(32) loss_pick_kernel(int shift, LossState* __restrict__ st) {
    const int lane = threadIdx.x;
    long long h[8];
    long long mine = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) { h[q] = st->hist[lane * 8 + q]; mine += h[q]; }
    long long suffix = mine;                                  // sum over lanes >= this lane
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long v = __shfl_down_sync(0xffffffffu, suffix, o);
        if (lane + o < 32) suffix += v;
    }
    const long long k = st->k_rem;
    const long long above = suffix - mine;                    // keys in higher buckets than this lane's
    const bool has = above < k && k <= above + mine;
    __syncwarp();
    if (has) {
        long long cum = above;
        int sel = lane * 8;
        for (int q = 7; q >= 0; --q) {
            if (k <= cum + h[q]) { sel = lane * 8 + q; break; }
            cum += h[q];
        }
        st->prefix |= (unsigned)sel << shift;
        st->pmask |= 0xffu << shift;
        st->k_rem = k - cum;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) st->hist[lane * 8 + q] = 0;
}

// boxes per image whose negative loss equals the k-th largest value
__global__ void __launch_bounds__(256)
loss_eqcount_kernel(const float* __restrict__ nl, int A, const LossState* __restrict__ st, int* __restrict__ eq) {
    __shared__ int s;
    if (threadIdx.x == 0) s = 0;
    __syncthreads();
    const unsigned tau = st->prefix;
    const float* p = nl + (size_t)blockIdx.x * A;
    int c = 0;
    for (int a = threadIdx.x; a < A; a += 256) c += (ord32(p[a]) == tau) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s, c);
    __syncthreads();
    if (threadIdx.x == 0) eq[blockIdx.x] = s;
}

// how many of an image's tied boxes are admitted: ties are taken in flat index order (tf.nn.top_k)
__global__ void __launch_bounds__(1024)
loss_quota_kernel(const int* __restrict__ eq, int B, const LossState* __restrict__ st, int* __restrict__ quota) {
    __shared__ long long wsum[32];
    __shared__ long long carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long need = st->k_rem;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < B; base += 1024) {
        const int b = base + tid;
        const long long c = (b < B) ? eq[b] : 0;
        long long x = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const long long v = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += v; }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            long long w = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const long long v = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += v; }
            wsum[lane] = w;
        }
        __syncthreads();
        const long long before = carry + (warp ? wsum[warp - 1] : 0) + (x - c);     // ties in earlier images
        if (b < B) {
            long long q = need - before;
            q = q < 0 ? 0 : (q > c ? c : q);
            quota[b] = (int)q;
        }
        __syncthreads();
        if (tid == 1023) carry = before + c;
        __syncthreads();
    }
}

// sum of the classification losses of the kept negatives of one image (:186-189)
__global__ void __launch_bounds__(256)
loss_negsum_kernel(const float* __restrict__ nl, const float* __restrict__ closs, int A, const LossState* __restrict__ st,
                   const int* __restrict__ eq, const int* __restrict__ quota, double* __restrict__ img_neg) {
    __shared__ int wcnt[8];
    __shared__ double wsum[8];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned tau = st->prefix;
    const int q = quota[blockIdx.x];
    const float* p = nl + (size_t)blockIdx.x * A;
    const float* cl = closs + (size_t)blockIdx.x * A;
    if (tid == 0) s_base = 0;
    __syncthreads();
    double sum = 0.0;
    if (eq[blockIdx.x] == q) {
        // every tie of this image is admitted (or there is none): no ranks needed
        for (int a = tid; a < A; a += 256) if (ord32(p[a]) >= tau) sum += (double)cl[a];
    } else
    for (int a0 = 0; a0 < A; a0 += 256) {
        const int a = a0 + tid;
        const unsigned k = (a < A) ? ord32(p[a]) : 0u;
        const bool tie = (a < A) && k == tau;
        // rank of this tie among the image's ties, in anchor order
        const unsigned m = __ballot_sync(0xffffffffu, tie);
        if (lane == 0) wcnt[warp] = __popc(m);
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < warp; ++w) before += wcnt[w];
        const int rank = before + __popc(m & ((1u << lane) - 1u));
        if (a < A && (k > tau || (tie && rank < q))) sum += (double)cl[a];
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += wcnt[w]; s_base += t; }
        __syncthreads();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) wsum[warp] = sum;
    __syncthreads();
    if (tid == 0) { double t = 0.0; for (int w = 0; w < 8; ++w) t += wsum[w]; img_neg[blockIdx.x] = t; }
}

__global__ void loss_final_kernel(const double* __restrict__ img_pos, const double* __restrict__ img_neg, const double* __restrict__ img_loc,
                                  const LossState* __restrict__ st, int B, float alpha, int use_neg, float* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float pos_cls = (float)img_pos[b];
    const float neg_cls = use_neg ? (float)img_neg[b] : 0.0f;                              // :191
    const float class_loss = pos_cls + neg_cls;                                            // :193
    const float loc_loss = (float)img_loc[b];
    const float n_positive = (float)st->n_pos;
    const float total = (class_loss + alpha * loc_loss) / fmaxf(1.0f, n_positive);         // :202
    out[b] = total * (float)B;                                                             // :207
}

}  // namespace ssdc

using namespace ssdc;

extern "C" int ssdc_ssd_loss(ssdc_ctx* ctx, const void* y_true, int dtype_true, const float* y_pred, int on_device,
                             int64_t B, int64_t A, int C, int neg_pos_ratio, int n_neg_min, double alpha, float* out_loss) {
    if (!ctx || !y_true || !y_pred || !out_loss || B <= 0 || A <= 0 || C < 1 || (dtype_true != SSDC_F32 && dtype_true != SSDC_F64)) {
        set_error("ssdc_ssd_loss: bad argument"); return SSDC_ERR_ARG;
    }
    if (B * A > 0x7fffffffLL) { set_error("ssdc_ssd_loss: batch too large"); return SSDC_ERR_ARG; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (on_device && ctx->devs.size() != 1) { set_error("ssdc_ssd_loss: device-resident input needs a single-device context"); return SSDC_ERR_ARG; }
    DevCtx& d = ctx->devs[0];
    SSDC_CUDA(cudaSetDevice(d.device));
    cudaStream_t st = d.stream;
    if (on_device) SSDC_TRY(d.wait_encodes());             // (y_true may be what an encode still in flight on a lane writes)
    const int W = C + 12;
    const long long n_boxes = B * A;
    const size_t true_bytes = (size_t)n_boxes * W * (dtype_true == SSDC_F32 ? 4 : 8), pred_bytes = (size_t)n_boxes * W * 4;
    const void* d_true = y_true;
    const float* d_pred = y_pred;
    if (!on_device) {
        SSDC_TRY(d.t0buf.ensure(true_bytes));
        SSDC_TRY(d.t1buf.ensure(pred_bytes));
        SSDC_CUDA(cudaMemcpyAsync(d.t0buf.p, y_true, true_bytes, cudaMemcpyHostToDevice, st));
        SSDC_CUDA(cudaMemcpyAsync(d.t1buf.p, y_pred, pred_bytes, cudaMemcpyHostToDevice, st));
        d_true = d.t0buf.p; d_pred = d.t1buf.as<float>();
    }
    // scratch: per-box losses, per-image sums, state
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_cl = carve((size_t)n_boxes * 4), o_nl = carve((size_t)n_boxes * 4);
    const size_t o_sum = carve((size_t)B * 3 * sizeof(double)), o_st = carve(sizeof(LossState));
    const size_t o_eq = carve((size_t)B * 4), o_q = carve((size_t)B * 4), o_out = carve((size_t)B * 4);
    SSDC_TRY(d.t2buf.ensure(off));
    char* base = d.t2buf.as<char>();
    float* closs = reinterpret_cast<float*>(base + o_cl);
    float* nl = reinterpret_cast<float*>(base + o_nl);
    double* img_pos = reinterpret_cast<double*>(base + o_sum);
    double* img_neg = img_pos + B;
    double* img_loc = img_neg + B;
    LossState* stt = reinterpret_cast<LossState*>(base + o_st);
    int* eq = reinterpret_cast<int*>(base + o_eq);
    int* quota = reinterpret_cast<int*>(base + o_q);
    float* d_out = reinterpret_cast<float*>(base + o_out);
    SSDC_CUDA(cudaMemsetAsync(base + o_sum, 0, (o_eq - o_sum), st));            // sums + state
    const bool tma_ok = ((reinterpret_cast<uintptr_t>(d_true) | reinterpret_cast<uintptr_t>(d_pred)) % 16 == 0) && (n_boxes % 4 == 0) &&
                        ctx->opt[SSDC_OPT_LOSS_NO_TMA] == 0;
    if (tma_ok) {
        LaunchScope ls(ctx, &d, SSDC_K_THIN);
        const size_t tsz = (dtype_true == SSDC_F32) ? 4 : 8;
        const size_t stage = (((size_t)LS_ROWS * W * 4 + 127) & ~(size_t)127) + (((size_t)LS_ROWS * W * tsz + 127) & ~(size_t)127);
        const size_t smem = stage * LS_STAGES;
        int ctas_per_sm = (int)((224 * 1024) / (smem + 2048));
        if (ctas_per_sm > 4) ctas_per_sm = 4;
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        const long long tiles = (n_boxes + LS_ROWS - 1) / LS_ROWS;
        const unsigned grid = (unsigned)std::min<long long>(tiles, (long long)d.sm_count * ctas_per_sm);
        if (dtype_true == SSDC_F32) {
            SSDC_CUDA(cudaFuncSetAttribute(loss_box_tma_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            loss_box_tma_kernel<float><<<grid, LS_ROWS + 32, smem, st>>>((const float*)d_true, d_pred, n_boxes, (int)A, C, W, closs, nl, img_pos, img_loc, stt);
        } else {
            SSDC_CUDA(cudaFuncSetAttribute(loss_box_tma_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            loss_box_tma_kernel<double><<<grid, LS_ROWS + 32, smem, st>>>((const double*)d_true, d_pred, n_boxes, (int)A, C, W, closs, nl, img_pos, img_loc, stt);
        }
        SSDC_TRY(check_launch("loss_box_tma_kernel"));
    } else {
        LaunchScope ls(ctx, &d, SSDC_K_THIN);
        const size_t smem = (size_t)LS_ROWS * W * 2 * sizeof(float);
        const long long tiles = (n_boxes + LS_ROWS - 1) / LS_ROWS;
        const unsigned grid = (unsigned)std::min<long long>(tiles, (long long)d.sm_count * 6);
        if (dtype_true == SSDC_F32) {
            SSDC_CUDA(cudaFuncSetAttribute(loss_box_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            loss_box_kernel<float><<<grid, LS_ROWS, smem, st>>>((const float*)d_true, d_pred, n_boxes, (int)A, C, W, closs, nl, img_pos, img_loc, stt);
        } else {
            SSDC_CUDA(cudaFuncSetAttribute(loss_box_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            loss_box_kernel<double><<<grid, LS_ROWS, smem, st>>>((const double*)d_true, d_pred, n_boxes, (int)A, C, W, closs, nl, img_pos, img_loc, stt);
        }
        SSDC_TRY(check_launch("loss_box_kernel"));
    }
    // k = min(max(neg_pos_ratio * int(n_positive), n_neg_min), n_neg_losses)   (:163)
    LossState hs;
    SSDC_CUDA(cudaMemcpyAsync(&hs, stt, sizeof(double) + sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    SSDC_CUDA(cudaStreamSynchronize(st));
    const float n_pos_f = (float)hs.n_pos;
    long long k = (long long)neg_pos_ratio * (long long)(int)n_pos_f;
    if (k < n_neg_min) k = n_neg_min;
    if (k > (long long)hs.n_nonzero) k = (long long)hs.n_nonzero;
    const int use_neg = (hs.n_nonzero != 0 && k > 0) ? 1 : 0;
    if (use_neg) {
        SSDC_CUDA(cudaMemcpyAsync(&stt->k_rem, &k, sizeof(long long), cudaMemcpyHostToDevice, st));
        const unsigned grid = (unsigned)std::min<long long>((n_boxes + 255) / 256, (long long)d.sm_count * 8);
        for (int shift = 24; shift >= 0; shift -= 8) {
            LaunchScope ls(ctx, &d, SSDC_K_THIN);
            loss_hist_kernel<<<grid, 256, 0, st>>>(nl, n_boxes, shift, stt);
            SSDC_TRY(check_launch("loss_hist_kernel"));
            ctx->launches.fetch_add(1, std::memory_order_relaxed);
            loss_pick_kernel<<<1, 32, 0, st>>>(shift, stt);
            SSDC_TRY(check_launch("loss_pick_kernel"));
        }
        LaunchScope ls(ctx, &d, SSDC_K_THIN);
        loss_eqcount_kernel<<<(unsigned)B, 256, 0, st>>>(nl, (int)A, stt, eq);
        SSDC_TRY(check_launch("loss_eqcount_kernel"));
        loss_quota_kernel<<<1, 1024, 0, st>>>(eq, (int)B, stt, quota);
        SSDC_TRY(check_launch("loss_quota_kernel"));
        loss_negsum_kernel<<<(unsigned)B, 256, 0, st>>>(nl, closs, (int)A, stt, eq, quota, img_neg);
        SSDC_TRY(check_launch("loss_negsum_kernel"));
        ctx->launches.fetch_add(2, std::memory_order_relaxed);
    }
    {
        LaunchScope ls(ctx, &d, SSDC_K_THIN);
        loss_final_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(img_pos, img_neg, img_loc, stt, (int)B, (float)alpha, use_neg, d_out);
        SSDC_TRY(check_launch("loss_final_kernel"));
    }
    SSDC_CUDA(cudaMemcpyAsync(out_loss, d_out, (size_t)B * sizeof(float), cudaMemcpyDeviceToHost, st));
    SSDC_CUDA(cudaStreamSynchronize(st));
    return SSDC_OK;
}
