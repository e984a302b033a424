// ctx.cu - C ABI glue: context lifetime, memory helpers, timing, decode submit/collect.
#include "ctx.cuh"
#include <map>
#include <algorithm>

namespace ssdc {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
        return SSDC_ERR_CUDA;
    }
    return SSDC_OK;
}

int ensure_dyn_smem(int device, const void* fn, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> seen;
    std::lock_guard<std::mutex> lk(mu);
    size_t& cur = seen[std::make_pair(device, fn)];
    if (bytes <= cur) return SSDC_OK;
    SSDC_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    cur = bytes;
    return SSDC_OK;
}

static void shard(int64_t B, int n, int i, int64_t* b0, int64_t* b1) {
    // contiguous slices of ceil(B / n) images
    int64_t per = (B + n - 1) / n;
    *b0 = std::min<int64_t>(B, per * i);
    *b1 = std::min<int64_t>(B, per * (i + 1));
}

}  // namespace ssdc

using namespace ssdc;

extern "C" {

int ssdc_version(void) { return SSDC_VERSION; }
const char* ssdc_last_error(void) { return g_err; }

int ssdc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int ssdc_init(const int* device_ids, int n_devices, ssdc_ctx** out) {
    if (!out) { set_error("ssdc_init: out is NULL"); return SSDC_ERR_ARG; }
    *out = nullptr;
    int avail = 0;
    cudaError_t e = cudaGetDeviceCount(&avail);
    if (e != cudaSuccess || avail == 0) {
        cudaGetLastError();
        set_error("ssdc_init: no CUDA device available (%s); libssdcodec has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return SSDC_ERR_NODEVICE;
    }
    int def = 0;
    if (!device_ids || n_devices <= 0) { device_ids = &def; n_devices = 1; }
    ssdc_ctx* ctx = new ssdc_ctx();
    memset(ctx->prof_ms, 0, sizeof(ctx->prof_ms));
    memset(ctx->prof_n, 0, sizeof(ctx->prof_n));
    ctx->devs.resize(n_devices);
    for (int i = 0; i < n_devices; ++i) {
        DevCtx& d = ctx->devs[i];
        d.device = device_ids[i];
        if (d.device < 0 || d.device >= avail) {
            set_error("ssdc_init: device id %d out of range (have %d)", d.device, avail);
            ssdc_destroy(ctx); return SSDC_ERR_ARG;
        }
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, d.device) != cudaSuccess) { set_error("cudaGetDeviceProperties failed"); ssdc_destroy(ctx); return SSDC_ERR_CUDA; }
        if (prop.major != 10) {
            set_error("ssdc_init: device %d is sm_%d%d; libssdcodec is built for sm_100a (B200) only", d.device, prop.major, prop.minor);
            ssdc_destroy(ctx); return SSDC_ERR_NODEVICE;
        }
        d.sm_count = prop.multiProcessorCount;
        int prio_lo = 0, prio_hi = 0;
        if (cudaSetDevice(d.device) != cudaSuccess ||
            cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi) != cudaSuccess ||
            cudaStreamCreateWithPriority(&d.stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
            cudaStreamCreateWithFlags(&d.stream2, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithPriority(&d.stream_nms, cudaStreamNonBlocking, prio_lo) != cudaSuccess ||
            cudaEventCreateWithFlags(&d.ev_d1[0], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&d.ev_d1[1], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&d.ev_sweep[0], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&d.ev_sweep[1], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&d.ev_fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&d.ev_join, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreate(&d.t0) != cudaSuccess || cudaEventCreate(&d.t1) != cudaSuccess) {
            set_error("ssdc_init: stream/event creation failed on device %d: %s", d.device, cudaGetErrorString(cudaGetLastError()));
            ssdc_destroy(ctx); return SSDC_ERR_CUDA;
        }
    }
    *out = ctx;
    return SSDC_OK;
}

void ssdc_destroy(ssdc_ctx* ctx) {
    if (!ctx) return;
    { std::lock_guard<std::mutex> lk(ctx->mu); }       // a call still running on another thread finishes first
    for (DevCtx& d : ctx->devs) {
        if (d.device < 0) continue;
        cudaSetDevice(d.device);
        if (d.stream2) cudaStreamSynchronize(d.stream2);
        if (d.stream_nms) cudaStreamSynchronize(d.stream_nms);
        if (d.stream) cudaStreamSynchronize(d.stream);
        d.lanes_release();
        Buf* sb[] = {&d.shadow.ints, &d.shadow.keys, &d.shadow.hist, &d.shadow.pad_rows, &d.shadow.pad_anchor, &d.shadow.out_count};
        for (Buf* b : sb) b->release();
        for (int k = 0; k < 2; ++k) { if (d.ev_d1[k]) cudaEventDestroy(d.ev_d1[k]); if (d.ev_sweep[k]) cudaEventDestroy(d.ev_sweep[k]); }
        if (d.stream_nms) cudaStreamDestroy(d.stream_nms);
        Buf* bufs[] = {&d.y_in, &d.ints, &d.keys, &d.boxes, &d.aux_class, &d.sort_scratch, &d.merge_scratch, &d.out_rows,
                       &d.out_anchor, &d.out_count, &d.row_offset, &d.hist, &d.pad_rows, &d.pad_anchor, &d.gt, &d.gt_off, &d.partial, &d.matches,
                       &d.enc_out, &d.enc_out2, &d.enc_idx, &d.enc_flags, &d.t0buf, &d.t1buf, &d.t2buf, &d.t3buf};
        for (Buf* b : bufs) b->release();
        for (int i = 0; i < DevCtx::H_RING; ++i) { d.h_ring[i].release(); if (d.h_ev[i]) cudaEventDestroy(d.h_ev[i]); }
        for (int i = 0; i < DevCtx::FEED_RING; ++i) { d.feed_buf[i].release(); if (d.feed_ev[i]) cudaEventDestroy(d.feed_ev[i]); }
        for (int i = 0; i < 8; ++i) if (d.chunk_ev[i]) cudaEventDestroy(d.chunk_ev[i]);
        if (d.t0) cudaEventDestroy(d.t0);
        if (d.t1) cudaEventDestroy(d.t1);
        if (d.ev_fork) cudaEventDestroy(d.ev_fork);
        if (d.ev_join) cudaEventDestroy(d.ev_join);
        if (d.stream2) cudaStreamDestroy(d.stream2);
        if (d.stream) cudaStreamDestroy(d.stream);
    }
    for (auto& pe : ctx->prof_pending) { cudaEventDestroy(pe.a); cudaEventDestroy(pe.b); }
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    delete ctx;
}

int ssdc_ctx_num_devices(const ssdc_ctx* ctx) { return ctx ? (int)ctx->devs.size() : 0; }

int ssdc_synchronize(ssdc_ctx* ctx) {
    if (!ctx) { set_error("ctx is NULL"); return SSDC_ERR_ARG; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (DevCtx& d : ctx->devs) {
        SSDC_CUDA(cudaSetDevice(d.device));
        SSDC_TRY(d.wait_encodes());
        SSDC_CUDA(cudaStreamSynchronize(d.stream));
        SSDC_CUDA(cudaStreamSynchronize(d.stream_nms));
        d.sweep_pending[0] = d.sweep_pending[1] = false;
    }
    return SSDC_OK;
}

int ssdc_set_option(ssdc_ctx* ctx, int option, int64_t value) {
    if (!ctx || option < 0 || option >= SSDC_OPT_COUNT) { set_error("ssdc_set_option: bad context / unknown option %d", option); return SSDC_ERR_ARG; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->opt[option] = value;
    return SSDC_OK;
}
int64_t ssdc_get_option(const ssdc_ctx* ctx, int option) {
    if (!ctx || option < 0 || option >= SSDC_OPT_COUNT) return 0;
    return ctx->opt[option];
}

int64_t ssdc_launch_count(const ssdc_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }

int ssdc_profile_enable(ssdc_ctx* ctx, int on) {
    if (!ctx) { set_error("ctx is NULL"); return SSDC_ERR_ARG; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->profile = on != 0;
    return SSDC_OK;
}

int ssdc_profile_read(ssdc_ctx* ctx, double* ms, int64_t* launches) {
    if (!ctx) { set_error("ctx is NULL"); return SSDC_ERR_ARG; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (DevCtx& d : ctx->devs) {
        SSDC_CUDA(cudaSetDevice(d.device));
        SSDC_CUDA(cudaStreamSynchronize(d.stream));
    }
    for (auto& pe : ctx->prof_pending) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, pe.a, pe.b) == cudaSuccess) {
            ctx->prof_ms[pe.family] += t;
            ctx->prof_n[pe.family] += 1;
        } else cudaGetLastError();
        ctx->ev_pool.push_back(pe.a);
        ctx->ev_pool.push_back(pe.b);
    }
    ctx->prof_pending.clear();
    for (int i = 0; i < SSDC_K_COUNT; ++i) {
        if (ms) ms[i] = ctx->prof_ms[i];
        if (launches) launches[i] = ctx->prof_n[i];
        ctx->prof_ms[i] = 0; ctx->prof_n[i] = 0;
    }
    return SSDC_OK;
}

int ssdc_timer_start(ssdc_ctx* ctx) {
    if (!ctx) { set_error("ctx is NULL"); return SSDC_ERR_ARG; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (DevCtx& d : ctx->devs) {
        SSDC_CUDA(cudaSetDevice(d.device));
        SSDC_TRY(d.wait_sweeps());                       // the span starts when the work in flight on the side streams has ended
        SSDC_TRY(d.wait_encodes());
        SSDC_CUDA(cudaEventRecord(d.t0, d.stream));
    }
    return SSDC_OK;
}

int ssdc_timer_stop(ssdc_ctx* ctx, double* elapsed_ms) {
    if (!ctx) { set_error("ctx is NULL"); return SSDC_ERR_ARG; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    double mx = 0.0;
    for (DevCtx& d : ctx->devs) {
        SSDC_CUDA(cudaSetDevice(d.device));
        SSDC_TRY(d.wait_sweeps());                       // the span ends when the sweeps on the side stream have ended too
        SSDC_TRY(d.wait_encodes());                      // ... and the encodes on their lanes
        SSDC_CUDA(cudaEventRecord(d.t1, d.stream));
    }
    for (DevCtx& d : ctx->devs) {
        SSDC_CUDA(cudaSetDevice(d.device));
        SSDC_CUDA(cudaEventSynchronize(d.t1));
        float t = 0.f;
        SSDC_CUDA(cudaEventElapsedTime(&t, d.t0, d.t1));
        mx = std::max(mx, (double)t);
    }
    if (elapsed_ms) *elapsed_ms = mx;
    return SSDC_OK;
}

static DevCtx* slot(ssdc_ctx* ctx, int dev_slot) {
    if (!ctx || dev_slot < 0 || dev_slot >= (int)ctx->devs.size()) { set_error("bad ctx / device slot %d", dev_slot); return nullptr; }
    return &ctx->devs[dev_slot];
}

int ssdc_dev_alloc(ssdc_ctx* ctx, int dev_slot, uint64_t bytes, void** out) {
    DevCtx* d = slot(ctx, dev_slot);
    if (!d || !out) return SSDC_ERR_ARG;
    SSDC_CUDA(cudaSetDevice(d->device));
    SSDC_CUDA(cudaMalloc(out, bytes ? bytes : 1));
    return SSDC_OK;
}
int ssdc_dev_free(ssdc_ctx* ctx, int dev_slot, void* p) {
    DevCtx* d = slot(ctx, dev_slot);
    if (!d) return SSDC_ERR_ARG;
    SSDC_CUDA(cudaSetDevice(d->device));
    SSDC_CUDA(cudaFree(p));
    return SSDC_OK;
}
int ssdc_host_alloc(uint64_t bytes, void** out) {
    if (!out) return SSDC_ERR_ARG;
    SSDC_CUDA(cudaMallocHost(out, bytes ? bytes : 1));
    return SSDC_OK;
}
int ssdc_host_free(void* p) {
    SSDC_CUDA(cudaFreeHost(p));
    return SSDC_OK;
}
int ssdc_memcpy_h2d(ssdc_ctx* ctx, int dev_slot, void* dst, const void* src, uint64_t bytes) {
    DevCtx* d = slot(ctx, dev_slot);
    if (!d) return SSDC_ERR_ARG;
    SSDC_CUDA(cudaSetDevice(d->device));
    SSDC_TRY(d->wait_encodes());
    SSDC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, d->stream));
    SSDC_CUDA(cudaStreamSynchronize(d->stream));
    return SSDC_OK;
}
int ssdc_memcpy_d2h(ssdc_ctx* ctx, int dev_slot, void* dst, const void* src, uint64_t bytes) {
    DevCtx* d = slot(ctx, dev_slot);
    if (!d) return SSDC_ERR_ARG;
    SSDC_CUDA(cudaSetDevice(d->device));
    SSDC_TRY(d->wait_encodes());
    SSDC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, d->stream));
    SSDC_CUDA(cudaStreamSynchronize(d->stream));
    return SSDC_OK;
}

// ---------------------------------------------------------------------------
// decoder
// ---------------------------------------------------------------------------
static int check_decode_args(const void* y, int dtype, int64_t B, int64_t A, int C, const ssdc_decode_params* p) {
    if (!p) { set_error("decode: params is NULL"); return SSDC_ERR_ARG; }
    if (B < 0 || A <= 0 || C < 2) { set_error("decode: bad shape B=%lld A=%lld C=%d", (long long)B, (long long)A, C); return SSDC_ERR_ARG; }
    if (B > 0 && !y) { set_error("decode: y_pred is NULL"); return SSDC_ERR_ARG; }
    if (dtype != SSDC_F32 && dtype != SSDC_F64) { set_error("decode: dtype must be SSDC_F32 or SSDC_F64"); return SSDC_ERR_ARG; }
    if (p->mode < SSDC_MODE_PER_CLASS || p->mode > SSDC_MODE_LAYER_FAST) { set_error("decode: bad mode %d", p->mode); return SSDC_ERR_ARG; }
    if (p->input_coords < 0 || p->input_coords > 2) { set_error("decode: bad input_coords %d", p->input_coords); return SSDC_ERR_ARG; }
    if (p->border_pixels < 0 || p->border_pixels > 2) { set_error("decode: bad border_pixels %d", p->border_pixels); return SSDC_ERR_ARG; }
    if ((p->mode == SSDC_MODE_LAYER || p->mode == SSDC_MODE_LAYER_FAST)) {
        if (dtype != SSDC_F32) { set_error("decode: layer modes take float32 input"); return SSDC_ERR_ARG; }
        if (p->input_coords != SSDC_COORDS_CENTROIDS) { set_error("decode: layer modes support 'centroids' only"); return SSDC_ERR_ARG; }
        if (p->nms_cap <= 0 || p->top_k <= 0) { set_error("decode: layer modes need nms_cap > 0 and top_k > 0"); return SSDC_ERR_ARG; }
    }
    if (A > 0x7fffffff / (C + 12) || B * (int64_t)(p->mode == SSDC_MODE_PER_CLASS || p->mode == SSDC_MODE_LAYER ? C - 1 : 1) > 0x7fffffff) {
        set_error("decode: problem too large for 32-bit segment indexing"); return SSDC_ERR_ARG;
    }
    return SSDC_OK;
}

int ssdc_decode_submit(ssdc_ctx* ctx, const void* y_pred, int dtype, int on_device,
                       int64_t B, int64_t A, int C, const ssdc_decode_params* p) {
    if (!ctx) { set_error("ctx is NULL"); return SSDC_ERR_ARG; }
    SSDC_TRY(check_decode_args(y_pred, dtype, B, A, C, p));
    std::lock_guard<std::mutex> lk(ctx->mu);
    const int n = (int)ctx->devs.size();
    if (on_device) {
        if (n != 1) { set_error("decode: device-resident input needs a single-device context"); return SSDC_ERR_ARG; }
        return decode_submit_dev(ctx, &ctx->devs[0], y_pred, dtype, 1, 0, B, A, C, p);
    }
    const size_t elem = (dtype == SSDC_F32) ? 4 : 8;
    const size_t img_bytes = (size_t)A * (C + 12) * elem;
    for (int i = 0; i < n; ++i) {
        int64_t b0, b1;
        shard(B, n, i, &b0, &b1);
        const char* src = reinterpret_cast<const char*>(y_pred) + (size_t)b0 * img_bytes;
        SSDC_TRY(decode_submit_dev(ctx, &ctx->devs[i], src, dtype, 0, b0, b1 - b0, A, C, p));
    }
    return SSDC_OK;
}

int ssdc_decode_collect(ssdc_ctx* ctx, double* out_rows, int64_t capacity_rows,
                        int32_t* out_counts, int32_t* out_anchor_idx, int64_t* total_rows) {
    if (!ctx) { set_error("ctx is NULL"); return SSDC_ERR_ARG; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    const int n = (int)ctx->devs.size();
    std::vector<int64_t> totals(n, 0);
    int64_t total = 0;
    for (int i = 0; i < n; ++i) {
        SSDC_TRY(decode_finish_dev(ctx, &ctx->devs[i], &totals[i]));
        total += totals[i];
    }
    if (total_rows) *total_rows = total;
    // counts are always delivered
    for (int i = 0; i < n; ++i) {
        DevCtx& d = ctx->devs[i];
        if (d.job.B == 0 || !out_counts) continue;
        SSDC_CUDA(cudaSetDevice(d.device));
        SSDC_CUDA(cudaMemcpyAsync(out_counts + d.job.b0, d.out_count.p, (size_t)d.job.B * sizeof(int), cudaMemcpyDeviceToHost, d.stream));
    }
    if (capacity_rows < total || (total > 0 && !out_rows)) {
        for (DevCtx& d : ctx->devs) { cudaSetDevice(d.device); cudaStreamSynchronize(d.stream); }
        set_error("decode: output buffer holds %lld rows, %lld needed", (long long)capacity_rows, (long long)total);
        return SSDC_ERR_CAPACITY;
    }
    int64_t row0 = 0;
    for (int i = 0; i < n; ++i) {
        DevCtx& d = ctx->devs[i];
        if (d.job.B == 0) continue;
        SSDC_TRY(decode_emit_all_dev(ctx, &d, totals[i]));
        SSDC_CUDA(cudaSetDevice(d.device));
        if (totals[i] > 0) {
            SSDC_CUDA(cudaMemcpyAsync(out_rows + row0 * 6, d.out_rows.p, (size_t)totals[i] * 6 * sizeof(double), cudaMemcpyDeviceToHost, d.stream));
            if (out_anchor_idx)
                SSDC_CUDA(cudaMemcpyAsync(out_anchor_idx + row0, d.out_anchor.p, (size_t)totals[i] * sizeof(int), cudaMemcpyDeviceToHost, d.stream));
        }
        row0 += totals[i];
    }
    for (DevCtx& d : ctx->devs) {
        SSDC_CUDA(cudaSetDevice(d.device));
        SSDC_CUDA(cudaStreamSynchronize(d.stream));
    }
    return SSDC_OK;
}

int ssdc_decode_results_dev(ssdc_ctx* ctx, int dev_slot, const double** rows, const int32_t** anchors,
                            const int32_t** counts, int64_t* b0, int64_t* n_images, int32_t* top_k) {
    if (!ctx || dev_slot < 0 || dev_slot >= (int)ctx->devs.size()) { set_error("ssdc_decode_results_dev: bad argument"); return SSDC_ERR_ARG; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevCtx& d = ctx->devs[dev_slot];
    if (!d.job.valid || !d.job.padded) {
        set_error("ssdc_decode_results_dev: no device-resident padded result (needs a finite top_k, float32 input, per-class or layer mode)");
        return SSDC_ERR_STATE;
    }
    if (rows) *rows = d.pad_rows.as<double>();
    if (anchors) *anchors = d.pad_anchor.as<int32_t>();
    if (counts) *counts = d.out_count.as<int32_t>();
    if (b0) *b0 = d.job.b0;
    if (n_images) *n_images = d.job.B;
    if (top_k) *top_k = d.job.p.top_k;
    return SSDC_OK;
}

int ssdc_decode_stats(ssdc_ctx* ctx, int64_t* out3) {
    if (!ctx || !out3) { set_error("ssdc_decode_stats: bad argument"); return SSDC_ERR_ARG; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    out3[0] = out3[1] = out3[2] = 0;
    for (DevCtx& d : ctx->devs) {
        if (!d.job.valid || !d.job.padded) continue;
        out3[0] += d.job.stat_keys; out3[1] += d.job.stat_floored; out3[2] += d.job.stat_fallback;
    }
    return SSDC_OK;
}

int ssdc_decode(ssdc_ctx* ctx, const void* y_pred, int dtype, int64_t B, int64_t A, int C,
                const ssdc_decode_params* p, double* out_rows, int64_t capacity_rows,
                int32_t* out_counts, int32_t* out_anchor_idx, int64_t* total_rows) {
    SSDC_TRY(ssdc_decode_submit(ctx, y_pred, dtype, 0, B, A, C, p));
    return ssdc_decode_collect(ctx, out_rows, capacity_rows, out_counts, out_anchor_idx, total_rows);
}

}  // extern "C"
