// ctx.cuh - context, per-device state and scratch management of libssdcodec.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <mutex>
#include <vector>
#include <atomic>
#include <utility>

#include "../../include/ssdcodec.h"

namespace ssdc {

void set_error(const char* fmt, ...);

#define SSDC_CUDA(call)                                                                    \
    do {                                                                                   \
        cudaError_t _e = (call);                                                           \
        if (_e != cudaSuccess) {                                                           \
            ssdc::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call,             \
                            cudaGetErrorString(_e));                                       \
            return SSDC_ERR_CUDA;                                                          \
        }                                                                                  \
    } while (0)

#define SSDC_TRY(call)                                                                     \
    do {                                                                                   \
        int _r = (call);                                                                   \
        if (_r != SSDC_OK) return _r;                                                      \
    } while (0)

// A device buffer that only ever grows.
struct Buf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return SSDC_OK;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + (bytes >> 3) + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            e = cudaMalloc(&p, bytes);   // retry without head-room
            want = bytes;
        }
        if (e != cudaSuccess) {
            set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
            p = nullptr; cap = 0;
            return SSDC_ERR_CUDA;
        }
        cap = want;
        return SSDC_OK;
    }
    // grows to exactly `bytes` (no head-room): brings a sibling buffer to the capacity another one already has
    int reserve(size_t bytes) {
        if (bytes <= cap) return SSDC_OK;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {
            set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
            p = nullptr; cap = 0;
            return SSDC_ERR_CUDA;
        }
        cap = bytes;
        return SSDC_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return SSDC_OK;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        cudaError_t e = cudaMallocHost(&p, bytes);
        if (e != cudaSuccess) { set_error("cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e)); p = nullptr; return SSDC_ERR_CUDA; }
        cap = bytes;
        return SSDC_OK;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// State of one submitted decode on one device (results stay on the device
// until collected).
struct DecodeJob {
    bool valid = false;
    int64_t b0 = 0, B = 0, A = 0;   // image range of this device's shard
    int C = 0, NS = 0, dtype = 0;
    ssdc_decode_params p;
    bool emitted = false;           // rows already written on the device
    bool padded = false;            // image-sweep path: (B, top_k, 6) rows + counts are on the device (pad_rows / pad_anchor / out_count)
    bool scan_pending = false;      // image-sweep path: padded rows are on the device, the packed row offsets are not computed yet
    int64_t out_capacity = 0;       // rows the device out buffer can hold
    int iou_f32 = 0;
    int64_t stat_keys = 0, stat_floored = 0, stat_fallback = 0;     // image sweep statistics, valid after the collect
};

struct DevCtx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;              // side stream for work that overlaps the main stream (encoder template)
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // decode scratch
    Buf y_in, ints, keys, boxes, aux_class, sort_scratch, merge_scratch, out_rows, out_anchor, out_count, row_offset;
    Buf hist;                        // image-sweep path: per-image 256-bin score histograms left by D1 (zero between decodes)
    bool hist_clean = false;
    // Pipelining of consecutive image-sweep decodes: the sweep of decode i runs on its own (lower priority) stream beside
    // D1 of decode i + 1.  The scratch a decode owns from its D1 to its sweep (ints, keys, hist, pad_rows, pad_anchor,
    // out_count, hist_clean above) exists twice; the members above are the bank of the LATEST decode, `shadow` holds the
    // other one, and every image-sweep submit swaps them.
    struct Bank { Buf ints, keys, hist, pad_rows, pad_anchor, out_count; bool hist_clean = false; } shadow;
    int bank = 0;                                            // index of the current bank (events below)
    cudaStream_t stream_nms = nullptr;
    cudaEvent_t ev_d1[2] = {nullptr, nullptr};               // D1 of the bank's decode finished (recorded on `stream`)
    cudaEvent_t ev_sweep[2] = {nullptr, nullptr};            // sweep of the bank's decode finished (recorded on `stream_nms`)
    bool sweep_pending[2] = {false, false};
    void swap_banks() {
        std::swap(ints, shadow.ints); std::swap(keys, shadow.keys); std::swap(hist, shadow.hist);
        std::swap(pad_rows, shadow.pad_rows); std::swap(pad_anchor, shadow.pad_anchor); std::swap(out_count, shadow.out_count);
        std::swap(hist_clean, shadow.hist_clean);
        bank ^= 1;
    }
    // makes `stream` wait for the sweeps still in flight on `stream_nms` (bank b, or both for b < 0)
    int wait_sweeps(int b = -1) {
        for (int k = 0; k < 2; ++k) {
            if ((b >= 0 && k != b) || !sweep_pending[k]) continue;
            cudaError_t e = cudaStreamWaitEvent(stream, ev_sweep[k], 0);
            if (e != cudaSuccess) { set_error("cudaStreamWaitEvent(sweep) failed: %s", cudaGetErrorString(e)); return SSDC_ERR_CUDA; }
            sweep_pending[k] = false;
        }
        return SSDC_OK;
    }
    Buf pad_rows, pad_anchor;        // image-sweep path: (B, top_k, 6) float64 rows + anchor ids, as the sweep leaves them
    // small pinned staging areas for asynchronous H2D copies of per-call host data: a ring guarded by events, so a
    // call that only enqueues work never overwrites bytes an earlier call's copy has not read yet
    static constexpr int H_RING = 8;
    PinnedBuf h_ring[H_RING];
    cudaEvent_t h_ev[H_RING] = {};
    bool h_busy[H_RING] = {};
    int h_next = 0;
    // Stages `bytes` of per-call host data: returns a pinned area the caller fills and then copies from with
    // cudaMemcpyAsync on `stream`, followed by staged_done().
    int stage_acquire(size_t bytes, void** out, int* slot_out) {
        const int s = h_next;
        h_next = (h_next + 1) % H_RING;
        if (h_busy[s]) {
            cudaError_t e = cudaEventSynchronize(h_ev[s]);
            if (e != cudaSuccess) { set_error("cudaEventSynchronize(staging) failed: %s", cudaGetErrorString(e)); return SSDC_ERR_CUDA; }
            h_busy[s] = false;
        }
        if (!h_ring[s].p) {
            // first use: every slot of the ring is allocated now (a page-locked allocation costs milliseconds; it must not
            // land in the fourth call of a steady loop)
            const size_t want = bytes > ((size_t)256 << 10) ? bytes + (bytes >> 2) : ((size_t)256 << 10);
            for (int k = 0; k < H_RING; ++k)
                if (!h_ring[k].p) { int r = h_ring[k].ensure(want); if (r != SSDC_OK) return r; }
        }
        int r = h_ring[s].ensure(bytes);
        if (r != SSDC_OK) return r;
        *out = h_ring[s].p; *slot_out = s;
        return SSDC_OK;
    }
    int stage_done(int s, cudaStream_t st) {
        if (!h_ev[s]) {
            cudaError_t e = cudaEventCreateWithFlags(&h_ev[s], cudaEventDisableTiming);
            if (e != cudaSuccess) { set_error("cudaEventCreate(staging) failed: %s", cudaGetErrorString(e)); return SSDC_ERR_CUDA; }
        }
        cudaError_t e = cudaEventRecord(h_ev[s], st);
        if (e != cudaSuccess) { set_error("cudaEventRecord(staging) failed: %s", cudaGetErrorString(e)); return SSDC_ERR_CUDA; }
        h_busy[s] = true;
        return SSDC_OK;
    }
    // host input of a decode: chunked copies on stream2, pageable sources staged through pinned buffers
    static constexpr int FEED_RING = 3;
    PinnedBuf feed_buf[FEED_RING];
    cudaEvent_t feed_ev[FEED_RING] = {nullptr, nullptr, nullptr};      // DMA out of the staging buffer finished
    bool feed_busy[FEED_RING] = {false, false, false};
    cudaEvent_t chunk_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // chunk landed on the device
    DecodeJob job;
    // encode scratch
    Buf gt, gt_off, partial, matches, enc_out, enc_out2, enc_idx, enc_flags;
    size_t cand_clean = 0;           // leading bytes of `matches` known to hold -1 (the sparse path's decision array between calls)
    // Pipelining of consecutive device-output encodes (ssdc_encode with on_device = 1 only enqueues work): call i runs on
    // lane i % ENC_LANES - its own stream pair, events and scratch - so the latency chain of one call (upload, seed, pair,
    // greedy rounds, patch) runs beside the chains of its neighbours and beside their template streams.  A lane starts
    // behind everything enqueued on `stream` before the call; `stream` (and with it every other entry point of the
    // library) waits for the lanes still in flight through wait_encodes().
    static constexpr int ENC_LANES = 6;
    struct EncLane {
        cudaStream_t st = nullptr, ts = nullptr;
        cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_done = nullptr;
        Buf gt, partial, matches;
        size_t cand_clean = 0;
        bool pending = false;
        const char* out_lo[3] = {nullptr, nullptr, nullptr};      // output ranges of the call in flight (WAW guard between lanes)
        const char* out_hi[3] = {nullptr, nullptr, nullptr};
    } enc_lane[ENC_LANES];
    cudaEvent_t ev_order = nullptr;  // recorded on `stream` at the start of a lane call
    int enc_next = 0;
    int lanes_init() {
        if (ev_order) return SSDC_OK;
        for (int k = 0; k < ENC_LANES; ++k) {
            EncLane& L = enc_lane[k];
            if (cudaStreamCreateWithFlags(&L.st, cudaStreamNonBlocking) != cudaSuccess ||
                cudaStreamCreateWithFlags(&L.ts, cudaStreamNonBlocking) != cudaSuccess ||
                cudaEventCreateWithFlags(&L.ev_fork, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&L.ev_join, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&L.ev_done, cudaEventDisableTiming) != cudaSuccess) {
                set_error("encode lanes: stream / event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
                return SSDC_ERR_CUDA;
            }
        }
        if (cudaEventCreateWithFlags(&ev_order, cudaEventDisableTiming) != cudaSuccess) {
            set_error("encode lanes: event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
            return SSDC_ERR_CUDA;
        }
        return SSDC_OK;
    }
    // makes `stream` wait for the encodes still in flight on the lanes
    int wait_encodes() {
        for (int k = 0; k < ENC_LANES; ++k) {
            EncLane& L = enc_lane[k];
            if (!L.pending) continue;
            cudaError_t e = cudaStreamWaitEvent(stream, L.ev_done, 0);
            if (e != cudaSuccess) { set_error("cudaStreamWaitEvent(encode lane) failed: %s", cudaGetErrorString(e)); return SSDC_ERR_CUDA; }
            L.pending = false;
        }
        return SSDC_OK;
    }
    void lanes_release() {
        for (int k = 0; k < ENC_LANES; ++k) {
            EncLane& L = enc_lane[k];
            if (L.st) cudaStreamSynchronize(L.st);
            if (L.ts) cudaStreamSynchronize(L.ts);
            L.gt.release(); L.partial.release(); L.matches.release();
            if (L.ev_fork) cudaEventDestroy(L.ev_fork);
            if (L.ev_join) cudaEventDestroy(L.ev_join);
            if (L.ev_done) cudaEventDestroy(L.ev_done);
            if (L.ts) cudaStreamDestroy(L.ts);
            if (L.st) cudaStreamDestroy(L.st);
            L = EncLane();
        }
        if (ev_order) { cudaEventDestroy(ev_order); ev_order = nullptr; }
    }
    // thin ops scratch
    Buf t0buf, t1buf, t2buf, t3buf;
};

}  // namespace ssdc

struct ssdc_ctx {
    std::vector<ssdc::DevCtx> devs;
    std::mutex mu;
    std::atomic<int64_t> launches{0};
    int64_t opt[SSDC_OPT_COUNT] = {0};
    bool profile = false;
    double prof_ms[SSDC_K_COUNT];
    int64_t prof_n[SSDC_K_COUNT];
    struct ProfEv { cudaEvent_t a, b; int family; int dev; };
    std::vector<ProfEv> prof_pending;
    std::vector<cudaEvent_t> ev_pool;
};

namespace ssdc {

// RAII bracket around a kernel launch: counts it and, when profiling is on,
// records CUDA events on the launching stream.
struct LaunchScope {
    ssdc_ctx* ctx; DevCtx* d; int family; cudaEvent_t a = nullptr, b = nullptr;
    LaunchScope(ssdc_ctx* c, DevCtx* dev, int fam) : ctx(c), d(dev), family(fam) {
        ctx->launches.fetch_add(1, std::memory_order_relaxed);
        if (ctx->profile) {
            a = get_ev(); b = get_ev();
            cudaEventRecord(a, d->stream);
        }
    }
    ~LaunchScope() {
        if (ctx->profile) {
            cudaEventRecord(b, d->stream);
            ctx->prof_pending.push_back({a, b, family, (int)(d - ctx->devs.data())});
        }
    }
    cudaEvent_t get_ev() {
        if (!ctx->ev_pool.empty()) { cudaEvent_t e = ctx->ev_pool.back(); ctx->ev_pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
};

int check_launch(const char* what);
// Raises a kernel's dynamic shared-memory limit on `device` when it is below `bytes` (remembered per device and kernel:
// the hot paths do not pay a driver call per launch for it).
int ensure_dyn_smem(int device, const void* fn, size_t bytes);

// implemented in decode.cu / encode.cu / thin.cu
int decode_submit_dev(ssdc_ctx* ctx, DevCtx* d, const void* y_pred, int dtype, int on_device,
                      int64_t b0, int64_t B, int64_t A, int C, const ssdc_decode_params* p);
int decode_finish_dev(ssdc_ctx* ctx, DevCtx* d, int64_t* total_rows);
int decode_emit_all_dev(ssdc_ctx* ctx, DevCtx* d, int64_t total_rows);

}  // namespace ssdc
