// evk_api.cu — the C-ABI of include/evk.h: handle, arena, loaders, downsample dispatch, k-means
// loop, streaming windows, profiling.  Host orchestration only; kernels are in the other .cu files.
// Replaces the OpenCL host code of ACCEL/store.cpp:55-139,219-326,370-568 and
// KM/assign_to_centers2.c:105-568.  There is no CPU path: every entry point needs the device.
#include <math.h>
#include <stdarg.h>
#include <ctype.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>

#include "evk_internal.cuh"

// stage timestamps; inside a stream capture the record becomes an event-record NODE of the graph
// (cudaEventRecordExternal), so the stage times stay measurable when the fused step is replayed
void evk_prof_rec(evk_handle* h, int i) {
    if (!h->profiling) return;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(h->stream, &cs);
    cudaEventRecordWithFlags(h->ev[i], h->stream,
                             cs == cudaStreamCaptureStatusActive ? cudaEventRecordExternal
                                                                 : cudaEventRecordDefault);
}

int evk_fail(evk_handle* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    return code;
}

namespace {

size_t next_pow2(size_t v) {
    size_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

void prof_rec(evk_handle* h, int i) { evk_prof_rec(h, i); }
float prof_ms(evk_handle* h, int a, int b) {
    float ms = 0.f;
    if (h->profiling && cudaEventElapsedTime(&ms, h->ev[a], h->ev[b]) != cudaSuccess) {
        cudaGetLastError();
        ms = 0.f;
    }
    return ms;
}

int check_handle(evk_handle* h) { return h ? EVK_OK : EVK_ERR_INVALID; }

void invalidate_results(evk_handle* h) {
    h->have_voxels = false;
    h->perm_valid = false;
    h->reps_valid = false;
    h->voxels_foreign = false;
    h->t_range_valid = false;
    h->n_unique = h->n_repeated = 0;
    h->n_labels = 0;
    h->pix_valid = false;
}

int km_validate(evk_handle* h, const evk_km_params* p) {
    if (!p) return evk_fail(h, EVK_ERR_INVALID, "km params are NULL");
    if (p->K < 1 || p->K > EVK_MAX_K) return evk_fail(h, EVK_ERR_INVALID, "K must be in [1,%d]", EVK_MAX_K);
    if (p->D < 2 || p->D > EVK_MAX_D) return evk_fail(h, EVK_ERR_INVALID, "D must be 2, 3 or 4");
    if (p->iters < 1) return evk_fail(h, EVK_ERR_INVALID, "iters must be >= 1");
    return EVK_OK;
}

// pixel images of the current frame size (lazy; frames beyond 64 Mpixel use the per-point scan)
bool ensure_images(evk_handle* h, int width, int height) {
    const size_t need = (size_t)width * (size_t)height;
    if (need > (64ull << 20)) return false;
    if (h->image_pixels < need) {
        // cached graphs (fused step, Lloyd loop, sharded step) hold the old pointers: their keys
        // carry image_gen, so the next call of each re-captures.  A graph may still be in flight.
        DeviceGuard g(h->device);
        if (h->stream) cudaStreamSynchronize(h->stream);
        if (h->side) cudaStreamSynchronize(h->side);
        h->image_gen++;
        if (h->d_label_map) cudaFree(h->d_label_map);
        if (h->d_pixcnt) cudaFree(h->d_pixcnt);
        h->d_label_map = nullptr;
        h->d_pixcnt = nullptr;
        h->image_pixels = 0;
        h->pix_valid = false;
        if (cudaMalloc((void**)&h->d_label_map, need) != cudaSuccess ||
            cudaMalloc((void**)&h->d_pixcnt, need * sizeof(uint32_t)) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        h->image_pixels = need;
    }
    return true;
}

KmLaunch km_launch_params(const evk_handle* h, const evk_km_params* p) {
    KmLaunch kl;
    kl.K = p->K;
    kl.D = p->D;
    kl.best2 = (p->max_dist > 0.f && !isinf(p->max_dist)) ? p->max_dist * p->max_dist : INFINITY;
    kl.t_scale = p->t_scale;
    kl.p_scale = p->p_scale;
    kl.t0 = h->have_ds ? h->ds.t0_us : 0;
    kl.write_labels = 1;
    return kl;
}

}  // namespace

int evk_km_validate(evk_handle* h, const evk_km_params* p) { return km_validate(h, p); }
KmLaunch evk_km_launch_params(const evk_handle* h, const evk_km_params* p) {
    return km_launch_params(h, p);
}
bool evk_ensure_images(evk_handle* h, int width, int height) {
    return ensure_images(h, width, height);
}
void evk_invalidate_results(evk_handle* h) { invalidate_results(h); }

int evk_make_key_params(evk_handle* h, const evk_ds_params* p, KeyParams* kp) {
    if (!p) return evk_fail(h, EVK_ERR_INVALID, "ds params are NULL");
    if (p->width < 1 || p->height < 1 || p->width > 65536 || p->height > 65536)
        return evk_fail(h, EVK_ERR_INVALID, "width/height must be in [1,65536]");
    memset(kp, 0, sizeof *kp);
    kp->keyfn = p->keyfn;
    kp->width = p->width;
    kp->height = p->height;
    if (p->keyfn == EVK_KEY_REF_HASH8192) {
        kp->use_p = 0;
        kp->mx = kp->my = 0;
        kp->sx = kp->sy = 0;
        kp->NX = kp->NY = kp->P = 1;
        kp->vt = 0;
        kp->vt_shift = -1;
        kp->cells = 8192;
        return EVK_OK;
    }
    if (p->keyfn != EVK_KEY_VOXEL) return evk_fail(h, EVK_ERR_INVALID, "unknown keyfn %d", p->keyfn);
    if (p->vx < 1 || p->vy < 1 || p->vx > 65536 || p->vy > 65536)
        return evk_fail(h, EVK_ERR_INVALID, "vx/vy must be in [1,65536]");
    kp->use_p = p->use_polarity ? 1 : 0;
    kp->P = kp->use_p ? 2 : 1;
    auto magic = [](int v, uint32_t* m, int32_t* s) {
        *s = -1;
        *m = 0;
        if ((v & (v - 1)) == 0) {
            int k = 0;
            while ((1 << k) < v) k++;
            *s = k;
        } else {
            *m = (uint32_t)(((1ull << 32) + (uint64_t)v - 1) / (uint64_t)v);
        }
    };
    magic(p->vx, &kp->mx, &kp->sx);
    magic(p->vy, &kp->my, &kp->sy);
    kp->NX = (uint32_t)((p->width + p->vx - 1) / p->vx);
    kp->NY = (uint32_t)((p->height + p->vy - 1) / p->vy);
    kp->cells = (uint64_t)kp->NX * kp->NY * kp->P;
    kp->t0 = p->t0_us;
    kp->vt = p->vt_us > 0 ? p->vt_us : 0;
    kp->vt_shift = -1;
    if (kp->vt > 0) {
        uint64_t v = (uint64_t)kp->vt;
        if ((v & (v - 1)) == 0) {
            int s = 0;
            while ((1ull << s) < v) s++;
            kp->vt_shift = s;
        } else {
            kp->vt_limit = 0xFFFFFFFFFFFFFFFFull / v;      // dt <= limit: magic quotient exact
            kp->vt_magic = 0xFFFFFFFFFFFFFFFFull / v + 1;  // ceil(2^64 / v) (v not a power of 2)
        }
    }
    return EVK_OK;
}

extern "C" {

// debug aid (not part of the ABI): the runtime's pending error, without clearing it
EVK_API const char* evk_debug_peek_error(void) { return cudaGetErrorString(cudaPeekAtLastError()); }

const char* evk_version(void) { return "evk-b200 0.1 (sm_100a)"; }

const char* evk_last_error(const evk_handle* h) { return h ? h->err.c_str() : "null handle"; }

int evk_create(evk_handle** out, int device, size_t max_events) {
    if (!out) return EVK_ERR_INVALID;
    *out = nullptr;
    if (max_events == 0 || max_events >= 0xFF000000ull) return EVK_ERR_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return EVK_ERR_CUDA;  // no fallback
    if (device < 0 || device >= ndev) return EVK_ERR_INVALID;
    evk_handle* h = new (std::nothrow) evk_handle();
    if (!h) return EVK_ERR_NOMEM;
    h->device = device;
    h->max_events = max_events;
    auto bail = [&](int code) {
        if (getenv("EVK_DEBUG"))
            fprintf(stderr, "evk_create: %s\n", cudaGetErrorString(cudaGetLastError()));
        evk_destroy(h);
        return code;
    };
    if (cudaSetDevice(device) != cudaSuccess) return bail(EVK_ERR_CUDA);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(EVK_ERR_CUDA);
    h->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess)
        return bail(EVK_ERR_CUDA);
    const size_t m = max_events;
    h->table_cap = next_pow2(2 * m + 2);
    h->max_bins = 1u << 22;
    h->cand_cap = 1u << 16;
    h->flush_bytes = 256ull << 20;
#define ALLOC(ptr, bytes)                                                          \
    if (cudaMalloc((void**)&(ptr), (bytes)) != cudaSuccess) return bail(EVK_ERR_NOMEM)
    ALLOC(h->d_events, m * sizeof(evk_event));
    ALLOC(h->d_tkeys, h->table_cap * sizeof(uint64_t));
    ALLOC(h->d_tfirst, h->table_cap * sizeof(uint32_t));
    // the slab kernel hands out output slots in CTA-private chunks: room for the unfilled tails
    h->out_cap = m + (size_t)EVK_SLAB_CHUNK *
                         (3 * (size_t)h->sm_count * evk_slab_ctas_per_sm() + 4);
    ALLOC(h->d_keys, h->out_cap * sizeof(uint64_t));
    ALLOC(h->d_first, h->out_cap * sizeof(uint32_t));
    ALLOC(h->d_xy, h->out_cap * sizeof(uint32_t));
    ALLOC(h->d_slab_scratch, evk_slab_scratch_bytes(h->sm_count));
    ALLOC(h->d_labels, h->out_cap * sizeof(int32_t));  // the fused step labels un-compacted slots
    ALLOC(h->d_bin_start, (h->max_bins + 2) * sizeof(uint32_t));
    ALLOC(h->d_cnt, sizeof(DsCounters));
    ALLOC(h->d_cent, EVK_MAX_K * EVK_MAX_D * sizeof(float));
    ALLOC(h->d_acc, EVK_MAX_K * 5 * sizeof(unsigned long long));
    ALLOC(h->d_counts, EVK_MAX_K * sizeof(unsigned long long));
    ALLOC(h->d_shift, sizeof(float));
    ALLOC(h->d_prune_lists, EVK_PRUNE_TILES * 16);
    ALLOC(h->d_quads, EVK_MAX_QUADS);
    ALLOC(h->d_n_points, sizeof(unsigned long long));
    ALLOC(h->d_sticky, sizeof(unsigned long long));
    ALLOC(h->d_t0, sizeof(long long));
    ALLOC(h->d_cand, 2 * h->cand_cap * sizeof(uint32_t));
    ALLOC(h->d_flush, h->flush_bytes);
#undef ALLOC
    if (cudaMallocHost((void**)&h->h_cnt, sizeof(DsCounters)) != cudaSuccess) return bail(EVK_ERR_NOMEM);
    if (cudaMallocHost((void**)&h->h_shift, sizeof(float)) != cudaSuccess) return bail(EVK_ERR_NOMEM);
    for (auto& e : h->ev)
        if (cudaEventCreate(&e) != cudaSuccess) return bail(EVK_ERR_CUDA);
    for (auto& e : h->ev_timer)
        if (cudaEventCreate(&e) != cudaSuccess) return bail(EVK_ERR_CUDA);
    if (evk_launch_table_clear(h->d_tkeys, h->d_tfirst, h->table_cap, h->stream) != cudaSuccess)
        return bail(EVK_ERR_CUDA);
    cudaMemsetAsync(h->d_acc, 0, EVK_MAX_K * 5 * sizeof(unsigned long long), h->stream);
    cudaMemsetAsync(h->d_cnt, 0, sizeof(DsCounters), h->stream);
    cudaMemsetAsync(h->d_sticky, 0, sizeof(unsigned long long), h->stream);
    cudaMemsetAsync(h->d_slab_scratch, 0, evk_slab_scratch_bytes(h->sm_count), h->stream);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return bail(EVK_ERR_CUDA);
    *out = h;
    return EVK_OK;
}

int evk_comm_destroy(evk_handle* h);

int evk_destroy(evk_handle* h) {
    if (!h) return EVK_OK;
    DeviceGuard g(h->device);
    const bool dbg = getenv("EVK_DEBUG") != nullptr;
    auto chk = [&](cudaError_t e, const char* what) {
        if (dbg && e != cudaSuccess) fprintf(stderr, "evk_destroy: %s: %s\n", what, cudaGetErrorString(e));
    };
    if (h->stream) chk(cudaStreamSynchronize(h->stream), "sync");
    evk_comm_destroy(h);
    evk_aec_destroy(h);
    evk_ts_destroy(h);
    evk_filter_corners_destroy(h);
    evk_optics_destroy(h);
    evk_dbscan_destroy(h);
    evk_partition_free(h);
    // graphs first: they reference the buffers, events and streams released below
    if (h->fused_exec) chk(cudaGraphExecDestroy(h->fused_exec), "fused graph");
    if (h->loop_exec) chk(cudaGraphExecDestroy(h->loop_exec), "loop graph");
    void* ptrs[] = {h->d_events, h->d_tkeys,  h->d_tfirst, h->d_keys,   h->d_first,  h->d_xy,
                    h->d_reps,   h->d_labels, h->d_perm,   h->d_sort_tmp, h->d_sort_a, h->d_sort_b,
                    h->d_sort_c, h->d_sk_in,  h->d_sk_out, h->d_si_in,  h->d_si_out, h->d_sv_tmp,
                    h->d_bin_start, h->d_slab_scratch, h->d_cnt, h->d_cent,   h->d_acc,    h->d_counts, h->d_shift,
                    h->d_cand,   h->d_flush, h->d_prune_lists, h->d_label_map, h->d_pixcnt, h->d_quads,
                    h->d_win_stage, h->d_n_points, h->d_raw, h->d_raw_blk, h->d_sticky, h->d_t0, h->d_prune3_lists};
    for (void* p : ptrs)
        if (p) chk(cudaFree(p), "free");
    if (h->h_cnt) chk(cudaFreeHost(h->h_cnt), "free host");
    if (h->h_shift) chk(cudaFreeHost(h->h_shift), "free host");
    for (auto& e : h->ev)
        if (e) chk(cudaEventDestroy(e), "event");
    for (auto& e : h->ev_timer)
        if (e) chk(cudaEventDestroy(e), "event");
    if (h->ev_fork) chk(cudaEventDestroy(h->ev_fork), "event");
    if (h->ev_join) chk(cudaEventDestroy(h->ev_join), "event");
    if (h->side) chk(cudaStreamDestroy(h->side), "side stream");
    if (h->stream) chk(cudaStreamDestroy(h->stream), "stream");
    chk(cudaGetLastError(), "left-over error state");  // never hand a stale error to the next call
    delete h;
    return EVK_OK;
}

// ---- load ---------------------------------------------------------------------------------
static int append_host(evk_handle* h, const evk_event* begin, const evk_event* end, bool replace) {
    EVK_TRY(check_handle(h));
    if ((!begin && end != begin) || end < begin) return evk_fail(h, EVK_ERR_INVALID, "bad event range");
    // a queued step publishes first indices into the events it ran on: collect it before they change
    EVK_TRY(evk_collect_pending(h));
    DeviceGuard g(h->device);
    const size_t n = (size_t)(end - begin);
    const size_t at = replace ? 0 : h->n_events;
    if (at + n > h->max_events)
        return evk_fail(h, EVK_ERR_CAPACITY, "%zu events exceed the handle capacity %zu", at + n,
                        h->max_events);
    if (n)
        EVK_CUDA(h, cudaMemcpyAsync(h->d_events + at, begin, n * sizeof(evk_event),
                                    cudaMemcpyHostToDevice, h->stream));
    h->n_events = at + n;
    invalidate_results(h);
    return EVK_OK;
}

int evk_load_events(evk_handle* h, const evk_event* begin, const evk_event* end) {
    return append_host(h, begin, end, true);
}
int evk_append_events(evk_handle* h, const evk_event* begin, const evk_event* end) {
    return append_host(h, begin, end, false);
}

int evk_load_events_soa(evk_handle* h, const uint16_t* x, const uint16_t* y, const int64_t* t,
                        const uint8_t* p, size_t n) {
    EVK_TRY(check_handle(h));
    if (n && (!x || !y)) return evk_fail(h, EVK_ERR_INVALID, "x / y are NULL");
    if (n > h->max_events) return evk_fail(h, EVK_ERR_CAPACITY, "too many events");
    EVK_TRY(evk_collect_pending(h));
    DeviceGuard g(h->device);
    invalidate_results(h);
    h->n_events = n;
    if (!n) return EVK_OK;
    // stage the columns in the (idle) label / key buffers, then pack on the device
    uint16_t* dx = reinterpret_cast<uint16_t*>(h->d_labels);
    uint16_t* dy = dx + n;
    int64_t* dt = reinterpret_cast<int64_t*>(h->d_keys);
    uint8_t* dp = reinterpret_cast<uint8_t*>(h->d_first);
    EVK_CUDA(h, cudaMemcpyAsync(dx, x, n * 2, cudaMemcpyHostToDevice, h->stream));
    EVK_CUDA(h, cudaMemcpyAsync(dy, y, n * 2, cudaMemcpyHostToDevice, h->stream));
    if (t) EVK_CUDA(h, cudaMemcpyAsync(dt, t, n * 8, cudaMemcpyHostToDevice, h->stream));
    if (p) EVK_CUDA(h, cudaMemcpyAsync(dp, p, n, cudaMemcpyHostToDevice, h->stream));
    EVK_CUDA(h, evk_launch_soa_pack(dx, dy, t ? dt : nullptr, p ? dp : nullptr, n, h->d_events,
                                    h->stream));
    return EVK_OK;
}

int evk_load_coords_i32(evk_handle* h, const int32_t* xy, size_t n_pairs) {
    EVK_TRY(check_handle(h));
    if (n_pairs && !xy) return evk_fail(h, EVK_ERR_INVALID, "xy is NULL");
    if (n_pairs > h->max_events) return evk_fail(h, EVK_ERR_CAPACITY, "too many coordinates");
    EVK_TRY(evk_collect_pending(h));
    DeviceGuard g(h->device);
    invalidate_results(h);
    h->n_events = n_pairs;
    if (!n_pairs) return EVK_OK;
    int32_t* d = reinterpret_cast<int32_t*>(h->d_keys);
    EVK_CUDA(h, cudaMemcpyAsync(d, xy, n_pairs * 8, cudaMemcpyHostToDevice, h->stream));
    EVK_CUDA(h, evk_launch_coords_pack(d, n_pairs, h->d_events, h->stream));
    return EVK_OK;
}

int evk_load_csv(evk_handle* h, const char* path) {
    EVK_TRY(check_handle(h));
    if (!path) return evk_fail(h, EVK_ERR_INVALID, "path is NULL");
    FILE* f = fopen(path, "r");
    if (!f) return evk_fail(h, EVK_ERR_IO, "cannot open %s", path);
    std::vector<evk_event> ev;
    char line[256];
    while (fgets(line, sizeof line, f)) {
        long x, y, p;
        long long t;
        if (sscanf(line, "%ld,%ld,%lld,%ld", &x, &y, &t, &p) != 4) continue;
        evk_event e;
        e.x = (uint16_t)x;
        e.y = (uint16_t)y;
        e.p = (int16_t)p;
        e._pad = 0;
        e.t = (int64_t)t;
        ev.push_back(e);
    }
    fclose(f);
    int st = evk_load_events(h, ev.data(), ev.data() + ev.size());
    if (st == EVK_OK) cudaStreamSynchronize(h->stream);  // ev goes out of scope
    return st;
}

// RAW sensor words: uploaded as they are, decoded on the device.  fmt 2: EVT 2.0, 32-bit words,
// 4 B per event (evk_evt2.cu); fmt 3: EVT 3.0, 16-bit words (evk_evt3.cu).
static int load_raw_words(evk_handle* h, int fmt, const void* words, size_t n_words,
                          size_t* n_events) {
    EVK_TRY(check_handle(h));
    if (n_words && !words) return evk_fail(h, EVK_ERR_INVALID, "words is NULL");
    EVK_TRY(evk_collect_pending(h));
    DeviceGuard g(h->device);
    invalidate_results(h);
    h->n_events = 0;
    if (n_events) *n_events = 0;
    if (!n_words) return EVK_OK;
    const size_t bytes = n_words * (fmt == 2 ? 4 : 2);
    const size_t n_u32 = (bytes + 3) / 4;
    if (h->raw_cap_words < n_u32) {
        if (h->d_raw) cudaFree(h->d_raw);
        h->d_raw = nullptr;
        h->raw_cap_words = 0;
        if (cudaMalloc((void**)&h->d_raw, (n_u32 + 4) * sizeof(uint32_t)) != cudaSuccess) {
            cudaGetLastError();
            return evk_fail(h, EVK_ERR_NOMEM, "RAW staging of %zu words", n_words);
        }
        h->raw_cap_words = n_u32;
    }
    // block summaries: 2 u32 per block (EVT 2.0) or 7 (EVT 3.0); capacity counted in pairs
    const size_t nb = fmt == 2 ? evk_evt2_blocks(n_words) : evk_evt3_blocks(n_words);
    const size_t pairs = fmt == 2 ? nb : (7 * nb + 1) / 2;
    if (h->raw_cap_blocks < pairs) {
        if (h->d_raw_blk) cudaFree(h->d_raw_blk);
        h->d_raw_blk = nullptr;
        h->raw_cap_blocks = 0;
        if (cudaMalloc((void**)&h->d_raw_blk, 2 * pairs * sizeof(uint32_t)) != cudaSuccess) {
            cudaGetLastError();
            return evk_fail(h, EVK_ERR_NOMEM, "RAW block summaries");
        }
        h->raw_cap_blocks = pairs;
    }
    EVK_CUDA(h, cudaMemcpyAsync(h->d_raw, words, bytes, cudaMemcpyHostToDevice, h->stream));
    if (fmt == 2)
        EVK_CUDA(h, evk_launch_evt2_decode(h->d_raw, n_words, h->d_raw_blk, h->d_n_points,
                                           h->d_events, h->max_events, h->stream));
    else
        EVK_CUDA(h, evk_launch_evt3_decode(reinterpret_cast<const uint16_t*>(h->d_raw), n_words,
                                           h->d_raw_blk, h->d_n_points, h->d_events,
                                           h->max_events, h->stream));
    EVK_CUDA(h, cudaMemcpyAsync(&h->h_cnt->scratch[5], h->d_n_points, sizeof(unsigned long long),
                                cudaMemcpyDeviceToHost, h->stream));
    EVK_CUDA(h, cudaStreamSynchronize(h->stream));
    const size_t n = (size_t)h->h_cnt->scratch[5];
    if (n > h->max_events)
        return evk_fail(h, EVK_ERR_CAPACITY, "%zu CD events exceed the handle capacity %zu", n,
                        h->max_events);
    h->n_events = n;
    if (n_events) *n_events = n;
    return EVK_OK;
}
int evk_load_evt2(evk_handle* h, const uint32_t* words, size_t n_words, size_t* n_events) {
    return load_raw_words(h, 2, words, n_words, n_events);
}
int evk_load_evt3(evk_handle* h, const uint16_t* words, size_t n_words, size_t* n_events) {
    return load_raw_words(h, 3, words, n_words, n_events);
}

// A Metavision RAW recording: ASCII header lines starting with '%', then the binary payload.
int evk_load_raw(evk_handle* h, const char* path, size_t* n_events) {
    EVK_TRY(check_handle(h));
    if (!path) return evk_fail(h, EVK_ERR_INVALID, "path is NULL");
    FILE* f = fopen(path, "rb");
    if (!f) return evk_fail(h, EVK_ERR_IO, "cannot open %s", path);
    std::vector<unsigned char> buf;
    {
        fseek(f, 0, SEEK_END);
        const long sz = ftell(f);
        fseek(f, 0, SEEK_SET);
        if (sz < 0) {
            fclose(f);
            return evk_fail(h, EVK_ERR_IO, "cannot size %s", path);
        }
        buf.resize((size_t)sz);
        if (sz && fread(buf.data(), 1, (size_t)sz, f) != (size_t)sz) {
            fclose(f);
            return evk_fail(h, EVK_ERR_IO, "short read on %s", path);
        }
        fclose(f);
    }
    size_t pos = 0;
    bool evt2 = false, evt3 = false, other = false;
    while (pos < buf.size() && buf[pos] == '%') {  // header lines
        size_t eol = pos;
        while (eol < buf.size() && buf[eol] != '\n') eol++;
        std::string line(buf.begin() + pos, buf.begin() + eol);
        for (auto& c : line) c = (char)tolower((unsigned char)c);
        if (line.find("evt 2.0") != std::string::npos || line.find("evt2;") != std::string::npos ||
            line.find("format evt2") != std::string::npos)
            evt2 = true;
        else if (line.find("evt 3.0") != std::string::npos || line.find("evt3;") != std::string::npos ||
                 line.find("format evt3") != std::string::npos)
            evt3 = true;
        else if (line.find("% evt ") == 0 || line.find("% format ") == 0)
            other = true;
        pos = eol < buf.size() ? eol + 1 : eol;
    }
    if (other && !evt2 && !evt3)
        return evk_fail(h, EVK_ERR_IO, "%s: only the EVT 2.0 and EVT 3.0 payload formats are decoded",
                        path);
    if (evt3) {
        const size_t n16 = (buf.size() - pos) / 2;
        std::vector<uint16_t> w16(n16);
        if (n16) memcpy(w16.data(), buf.data() + pos, n16 * 2);
        return evk_load_evt3(h, w16.data(), n16, n_events);
    }
    const size_t n_words = (buf.size() - pos) / 4;
    std::vector<uint32_t> words(n_words);
    if (n_words) memcpy(words.data(), buf.data() + pos, n_words * 4);  // (payload may be unaligned)
    return evk_load_evt2(h, words.data(), n_words, n_events);
}

int evk_synth(evk_handle* h, const evk_synth_params* sp) {
    EVK_TRY(check_handle(h));
    if (!sp || sp->rate_eps == 0 || sp->width < 1 || sp->height < 1 || sp->n_blobs < 1)
        return evk_fail(h, EVK_ERR_INVALID, "bad synth params");
    if (sp->n_events > h->max_events) return evk_fail(h, EVK_ERR_CAPACITY, "too many events");
    EVK_TRY(evk_collect_pending(h));
    DeviceGuard g(h->device);
    invalidate_results(h);
    h->n_events = (size_t)sp->n_events;
    if (sp->n_events) EVK_CUDA(h, evk_launch_synth(*sp, h->d_events, h->stream));
    return EVK_OK;
}

int evk_num_events(const evk_handle* h, size_t* n) {
    if (!h || !n) return EVK_ERR_INVALID;
    *n = h->n_events;
    return EVK_OK;
}

int evk_get_events(evk_handle* h, evk_event* out, size_t first, size_t count) {
    EVK_TRY(check_handle(h));
    if (first + count > h->n_events) return evk_fail(h, EVK_ERR_INVALID, "range beyond stream");
    if (!count) return EVK_OK;
    if (!out) return evk_fail(h, EVK_ERR_INVALID, "out is NULL");
    DeviceGuard g(h->device);
    EVK_CUDA(h, cudaMemcpyAsync(out, h->d_events + first, count * sizeof(evk_event),
                                cudaMemcpyDeviceToHost, h->stream));
    EVK_CUDA(h, cudaStreamSynchronize(h->stream));
    return EVK_OK;
}

// ---- downsample -----------------------------------------------------------------------------
int evk_downsample_local(evk_handle* h, const evk_ds_params* p) {
    KeyParams kp;
    EVK_TRY(evk_make_key_params(h, p, &kp));
    DeviceGuard g(h->device);
    invalidate_results(h);
    h->ds = *p;
    h->kp = kp;
    h->have_ds = true;
    h->times.ds_total_ms = h->times.ds_main_ms = h->times.ds_compact_ms = 0.f;
    int launches = 0;
    EVK_CUDA(h, cudaMemsetAsync(h->d_cnt, 0, sizeof(DsCounters), h->stream));
    prof_rec(h, 0);
    int algo = p->algo;
    if (algo == EVK_ALGO_AUTO) algo = evk_slab_supported(h, kp) ? EVK_ALGO_SLAB : EVK_ALGO_TABLE;
    if (algo == EVK_ALGO_SLAB) {
        bool ok = false;
        if (evk_slab_supported(h, kp)) EVK_TRY(evk_downsample_slab(h, kp, p->count_repeated, &ok, &launches));
        if (!ok) {  // not partitioned by time bin (or unsupported shape)
            algo = EVK_ALGO_PARTITION;
            EVK_CUDA(h, cudaMemsetAsync(h->d_sticky, 0, sizeof(unsigned long long), h->stream));
        } else {
            prof_rec(h, 1);
            prof_rec(h, 2);
        }
    }
    if (algo == EVK_ALGO_PARTITION) {  // bring the stream into bin order first, then the slab kernel
        bool ok = false;
        EVK_CUDA(h, cudaMemsetAsync(h->d_cnt, 0, sizeof(DsCounters), h->stream));
        prof_rec(h, 0);
        EVK_TRY(evk_downsample_partitioned(h, kp, p->count_repeated, &ok, &launches));
        if (!ok) {  // key space or bin count the partition does not take: the general table
            algo = EVK_ALGO_TABLE;
            EVK_CUDA(h, cudaMemsetAsync(h->d_sticky, 0, sizeof(unsigned long long), h->stream));
            EVK_CUDA(h, cudaMemsetAsync(h->d_cnt, 0, sizeof(DsCounters), h->stream));
            prof_rec(h, 0);
        } else {
            prof_rec(h, 1);
            prof_rec(h, 2);
        }
    }
    if (algo == EVK_ALGO_TABLE) {
        EVK_CUDA(h, evk_launch_table_insert(kp, h->d_events, h->n_events, h->d_tkeys, h->d_tfirst,
                                            h->table_cap, p->count_repeated, h->sm_count,
                                            h->stream));
        prof_rec(h, 1);
        EVK_CUDA(h, evk_launch_table_compact(h->d_events, h->d_tkeys, h->d_tfirst, h->table_cap,
                                             h->d_keys, h->d_first, h->d_xy,
                                             (uint32_t)h->shard_first, h->d_cnt, h->sm_count,
                                             h->stream));
        prof_rec(h, 2);
        launches += 2;
    } else if (algo == EVK_ALGO_SORT) {
        EVK_TRY(evk_downsample_sort(h, kp, &launches));
        prof_rec(h, 1);
        prof_rec(h, 2);
    } else if (algo != EVK_ALGO_SLAB && algo != EVK_ALGO_PARTITION) {
        return evk_fail(h, EVK_ERR_INVALID, "unknown algo %d", p->algo);
    }
    EVK_CUDA(h, cudaMemcpyAsync(h->h_cnt, h->d_cnt, sizeof(DsCounters), cudaMemcpyDeviceToHost,
                                h->stream));
    EVK_CUDA(h, cudaStreamSynchronize(h->stream));
    h->n_unique = (size_t)h->h_cnt->n_unique;
    h->n_repeated = p->count_repeated ? (size_t)h->h_cnt->n_repeated : 0;
    h->have_voxels = true;
    // the slab kernels leave the first time bin and the bin count behind: every representative's
    // timestamp lies in that range (used by the space-time pruning of the D = 3 / 4 k-means)
    if ((algo == EVK_ALGO_SLAB || algo == EVK_ALGO_PARTITION) && kp.vt > 0 && h->n_unique) {
        h->t_range_lo = kp.t0 + (long long)h->h_cnt->scratch[2] * kp.vt;
        h->t_range_hi = h->t_range_lo + (long long)h->h_cnt->scratch[0] * kp.vt - 1;
        h->t_range_valid = true;
    }
    h->times.ds_algo_used = algo;
    h->times.ds_launches = launches;
    if (h->profiling) {
        h->times.ds_total_ms = prof_ms(h, 0, 2);
        h->times.ds_main_ms = algo == EVK_ALGO_SLAB ? prof_ms(h, 5, 6) : prof_ms(h, 0, 1);
        h->times.ds_compact_ms = prof_ms(h, 1, 2);
    }
    return EVK_OK;
}

// A queued fused step (evk_downsample_kmeans[_sharded]_submit) is collected before anything reads or
// replaces its results: the getters and the separate calls never see half a step.
int evk_collect_pending(evk_handle* h) {
    if (!h || !h->step_pending) return EVK_OK;
    if (h->step_sharded) return evk_downsample_kmeans_sharded_wait(h, nullptr, nullptr, nullptr);
    return evk_downsample_kmeans_wait(h, nullptr, nullptr, nullptr);
}

int evk_downsample(evk_handle* h, const evk_ds_params* p, size_t* n_unique, size_t* n_repeated) {
    EVK_TRY(check_handle(h));
    EVK_TRY(evk_collect_pending(h));
    h->shard_first = h->comm ? h->shard_first : 0;
    EVK_TRY(evk_downsample_local(h, p));
    if (n_unique) *n_unique = h->n_unique;
    if (n_repeated) *n_repeated = h->n_repeated;
    return EVK_OK;
}

// the counters of the current voxel shard (the blocking reads of unique_count / repeated_count,
// ACCEL/store.cpp:418-430): what the last downsample, fused step or completed window produced
int evk_num_voxels(const evk_handle* h, size_t* n_unique, size_t* n_repeated) {
    if (!h) return EVK_ERR_INVALID;
    if (h->step_pending == 1) {  // (const handle: cannot collect) a queued step has not been waited for
        if (n_unique) *n_unique = 0;
        if (n_repeated) *n_repeated = 0;
        return EVK_ERR_STATE;
    }
    if (n_unique) *n_unique = h->have_voxels ? h->n_unique : 0;
    if (n_repeated) *n_repeated = h->have_voxels ? h->n_repeated : 0;
    return h->have_voxels ? EVK_OK : EVK_ERR_STATE;
}

int evk_get_voxels(evk_handle* h, uint64_t* keys, evk_event* reps, uint32_t* first_idx, size_t cap) {
    EVK_TRY(check_handle(h));
    EVK_TRY(evk_collect_pending(h));
    if (!h->have_voxels) return evk_fail(h, EVK_ERR_STATE, "evk_downsample has not run");
    const size_t n = h->n_unique;
    if (cap < n) return evk_fail(h, EVK_ERR_CAPACITY, "cap %zu < %zu voxels", cap, n);
    if (!n) return EVK_OK;
    if (reps && h->voxels_foreign && !h->reps_valid)
        return evk_fail(h, EVK_ERR_STATE, "the representatives of hash-owned voxels were not "
                        "exchanged by the fused step: use evk_downsample_sharded");
    DeviceGuard g(h->device);
    EVK_TRY(evk_ensure_perm(h));
    // gather into the (now idle) sort-variant / scratch buffers, one column at a time
    if (!h->d_sort_c) EVK_CUDA(h, cudaMalloc(&h->d_sort_c, h->max_events * sizeof(evk_event)));
    void* scratch = h->d_sort_c;
    if (keys) {
        EVK_CUDA(h, evk_launch_gather_voxels(h, (uint64_t*)scratch, nullptr, nullptr, n));
        EVK_CUDA(h, cudaMemcpyAsync(keys, scratch, n * 8, cudaMemcpyDeviceToHost, h->stream));
    }
    if (first_idx) {
        EVK_CUDA(h, evk_launch_gather_voxels(h, nullptr, nullptr, (uint32_t*)scratch, n));
        EVK_CUDA(h, cudaMemcpyAsync(first_idx, scratch, n * 4, cudaMemcpyDeviceToHost, h->stream));
    }
    if (reps) {
        EVK_CUDA(h, evk_launch_gather_voxels(h, nullptr, (evk_event*)scratch, nullptr, n));
        EVK_CUDA(h, cudaMemcpyAsync(reps, scratch, n * 16, cudaMemcpyDeviceToHost, h->stream));
    }
    EVK_CUDA(h, cudaStreamSynchronize(h->stream));
    return EVK_OK;
}

// ---- k-means --------------------------------------------------------------------------------
int evk_set_centroids(evk_handle* h, const float* c, int K, int D) {
    EVK_TRY(check_handle(h));
    if (!c || K < 1 || K > EVK_MAX_K || D < 2 || D > EVK_MAX_D)
        return evk_fail(h, EVK_ERR_INVALID, "bad centroids (K=%d, D=%d)", K, D);
    DeviceGuard g(h->device);
    EVK_CUDA(h, cudaMemcpyAsync(h->d_cent, c, (size_t)K * D * sizeof(float), cudaMemcpyHostToDevice,
                                h->stream));
    EVK_CUDA(h, cudaStreamSynchronize(h->stream));  // c is borrowed
    h->K = K;
    h->D = D;
    h->have_centroids = true;
    return EVK_OK;
}

int evk_init_centroids_first_k(evk_handle* h, const evk_km_params* p) {
    EVK_TRY(check_handle(h));
    EVK_TRY(km_validate(h, p));
    if (!h->have_voxels) return evk_fail(h, EVK_ERR_STATE, "evk_downsample has not run");
    if (h->n_unique < (size_t)p->K)
        return evk_fail(h, EVK_ERR_INVALID, "only %zu voxels for K=%d", h->n_unique, p->K);
    DeviceGuard g(h->device);
    KmLaunch kl = km_launch_params(h, p);
    unsigned long long* d_count = &h->d_cnt->scratch[0];
    // fast path: walk the head of the stream until K distinct keys have been met
    if (!h->voxels_foreign && h->n_events) {
        const size_t n_scan = h->n_events < (1u << 20) ? h->n_events : (1u << 20);
        EVK_CUDA(h, evk_launch_init_first_k_walk(h->kp, kl, h->d_events, n_scan, h->d_cent, d_count,
                                                 h->stream));
        EVK_CUDA(h, cudaMemcpyAsync(&h->h_cnt->scratch[0], d_count, sizeof(unsigned long long),
                                    cudaMemcpyDeviceToHost, h->stream));
        EVK_CUDA(h, cudaStreamSynchronize(h->stream));
        if (h->h_cnt->scratch[0] == (unsigned long long)p->K) {
            h->K = p->K;
            h->D = p->D;
            h->have_centroids = true;
            return EVK_OK;
        }
    }
    // general path: voxels with first index below `bound` are the only candidates for the K lowest ones;
    // first indices are global in sharded mode, so the bound starts at the shard offset
    uint64_t span = 2048 + 32ull * p->K;
    for (int attempt = 0; attempt < 8; attempt++) {
        uint64_t bound64 = h->shard_first + span;
        uint32_t bound = bound64 > 0xFFFFFFFEull ? 0xFFFFFFFEu : (uint32_t)bound64;
        EVK_CUDA(h, cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), h->stream));
        EVK_CUDA(h, evk_launch_collect_below(h->d_first, h->n_unique, bound, h->d_cand,
                                             (uint32_t)h->cand_cap, d_count, h->stream));
        EVK_CUDA(h, cudaMemcpyAsync(&h->h_cnt->scratch[0], d_count, sizeof(unsigned long long),
                                    cudaMemcpyDeviceToHost, h->stream));
        EVK_CUDA(h, cudaStreamSynchronize(h->stream));
        const unsigned long long c = h->h_cnt->scratch[0];
        if (c >= (unsigned long long)p->K && c <= h->cand_cap) {
            const evk_event* ev0 = h->d_events - h->shard_first;  // indexable by global index
            EVK_CUDA(h, evk_launch_init_from_cand(kl, h->d_cand, (uint32_t)c, h->d_xy, ev0,
                                                  h->reps_valid ? h->d_reps : nullptr, h->d_cent,
                                                  h->stream));
            h->K = p->K;
            h->D = p->D;
            h->have_centroids = true;
            return EVK_OK;
        }
        if (c > h->cand_cap) break;
        span *= 16;
    }
    return evk_fail(h, EVK_ERR_CAPACITY, "centroid initialisation found no compact candidate set");
}

int evk_kmeans_run(evk_handle* h, const evk_km_params* p, int* iters_done,
                   int (*reduce)(evk_handle*, int K, int D)) {
    EVK_TRY(km_validate(h, p));
    if (!h->have_centroids || h->K != p->K || h->D != p->D)
        return evk_fail(h, EVK_ERR_STATE, "centroids for K=%d, D=%d have not been set", p->K, p->D);
    if (!p->on_events && !h->have_voxels)
        return evk_fail(h, EVK_ERR_STATE, "evk_downsample has not run");
    DeviceGuard g(h->device);
    KmLaunch kl = km_launch_params(h, p);
    const size_t n = p->on_events ? h->n_events : h->n_unique;
    const uint32_t* xy = p->on_events ? nullptr : h->d_xy;
    const evk_event* ev = p->on_events ? h->d_events : h->d_events - h->shard_first;
    if (!p->on_events && p->D > 2 && h->voxels_foreign)
        return evk_fail(h, EVK_ERR_INVALID, "D > 2 is not supported on a sharded voxel table");
    h->times.km_total_ms = h->times.km_assign_ms = 0.f;
    int it = 0, launches = 0;
    prof_rec(h, 3);
    // D == 2 on voxels: integer pixels, so labels come from a per-pixel label map (exact).  With
    // three or more iterations (or a tolerance stop) the iterations run on the W x H histogram of
    // the representatives and the voxel list is touched twice in total (histogram, final labels);
    // else every iteration is one two-level pass over the voxel list.
    // (EVK_KEY_REF_HASH8192 gates inclusively, x <= width: its representatives may lie one pixel
    // outside the W x H images, so that key function takes the per-point scan)
    const bool in_frame = xy && h->have_ds && h->kp.keyfn == EVK_KEY_VOXEL;
    const bool image = in_frame && p->D == 2 && p->K <= 254 &&
                       ensure_images(h, h->ds.width, h->ds.height);
    const bool on_hist = image && (p->iters >= 3 || p->tol >= 0.f);
    if (on_hist && p->tol < 0.f && !reduce && n) {
        // fixed number of iterations on the pixel histogram: the whole loop (histogram, iters x
        // (candidate lists, image pass, finalise), final labels) is one CUDA graph; the point count
        // is read on the device so the graph survives from call to call
        evk_handle::LoopKey key;
        memset(&key, 0, sizeof key);
        key.width = h->ds.width;
        key.height = h->ds.height;
        key.K = p->K;
        key.iters = p->iters;
        key.need_hist = h->pix_valid ? 0 : 1;
        key.profiling = h->profiling ? 1 : 0;
        key.max_dist = p->max_dist;
        key.cap = h->out_cap;
        key.image_gen = h->image_gen;
        EVK_CUDA(h, evk_launch_set_u64(h->d_n_points, (unsigned long long)n, h->stream));
        if (!h->loop_exec || memcmp(&key, &h->loop_key, sizeof key) != 0) {
            if (h->loop_exec) {
                // (driver 580 / CUDA 12.9: destroying this executable graph intermittently
                // reports cudaErrorInvalidValue although it was instantiated and replayed
                // correctly; the error state is cleared so that it cannot leak into later calls)
                cudaGraphExecDestroy(h->loop_exec);
                cudaGetLastError();
            }
            h->loop_exec = nullptr;
            cudaGraph_t graph = nullptr;
            EVK_CUDA(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
            cudaError_t ce = cudaSuccess;
            if (key.need_hist)
                ce = evk_launch_pix_hist(xy, h->out_cap, h->ds.width, h->ds.height, h->d_pixcnt,
                                         h->sm_count, h->stream, h->d_n_points);
            for (int i = 0; i < p->iters && ce == cudaSuccess; i++) {
                ce = evk_launch_km_image(kl, h->ds.width, h->ds.height, h->d_prune_lists, h->d_cent,
                                         h->d_pixcnt, h->d_label_map,
                                         i == p->iters - 1 ? h->d_quads : nullptr, h->d_acc,
                                         h->stream);
                if (ce == cudaSuccess)
                    ce = evk_launch_km_finalise(kl, h->d_cent, h->d_acc, h->d_counts, h->d_shift,
                                                h->stream);
            }
            if (ce == cudaSuccess)
                ce = evk_launch_km_assign_tiles(kl, h->ds.width, h->ds.height, h->d_quads,
                                                h->d_label_map, xy, h->out_cap, h->d_n_points,
                                                false, h->d_acc, h->d_labels, h->sm_count,
                                                h->stream);
            const cudaError_t ce2 = cudaStreamEndCapture(h->stream, &graph);
            if (ce != cudaSuccess || ce2 != cudaSuccess) {
                if (graph) cudaGraphDestroy(graph);
                cudaGetLastError();
                return evk_fail(h, EVK_ERR_CUDA, "k-means loop capture: %s",
                                cudaGetErrorString(ce != cudaSuccess ? ce : ce2));
            }
            ce = cudaGraphInstantiate(&h->loop_exec, graph, 0);
            cudaGraphDestroy(graph);
            EVK_CUDA(h, ce);
            h->loop_key = key;
        }
        EVK_CUDA(h, cudaGraphLaunch(h->loop_exec, h->stream));
        h->pix_valid = true;
        prof_rec(h, 4);
        h->n_labels = n;
        h->labels_on_events = false;
        h->km_last = *p;
        h->times.km_iters = p->iters;
        h->times.km_launches = 3 * p->iters + 2 + key.need_hist;
        if (h->profiling) {
            EVK_CUDA(h, cudaStreamSynchronize(h->stream));
            h->times.km_total_ms = prof_ms(h, 3, 4);
            h->times.km_assign_ms = h->times.km_total_ms;
        }
        if (iters_done) *iters_done = p->iters;
        return EVK_OK;
    }
    if (on_hist && !h->pix_valid) {
        EVK_CUDA(h, evk_launch_pix_hist(xy, n, h->ds.width, h->ds.height, h->d_pixcnt, h->sm_count,
                                        h->stream));
        h->pix_valid = true;
        launches++;
    }
    while (it < p->iters) {
        kl.write_labels = (p->tol >= 0.f || it == p->iters - 1) ? 1 : 0;
        cudaError_t ce = cudaErrorNotSupported;
        if (image) {
            const bool last = it == p->iters - 1;  // quads: only where a per-voxel pass follows
            ce = evk_launch_km_image(kl, h->ds.width, h->ds.height, h->d_prune_lists, h->d_cent,
                                     on_hist ? h->d_pixcnt : nullptr, h->d_label_map,
                                     (!on_hist || last || p->tol >= 0.f) ? h->d_quads : nullptr,
                                     h->d_acc, h->stream);
            if (ce == cudaSuccess && !on_hist)
                ce = evk_launch_km_assign_tiles(kl, h->ds.width, h->ds.height, h->d_quads,
                                                h->d_label_map, xy, n, nullptr, true, h->d_acc,
                                                h->d_labels, h->sm_count, h->stream);
        } else {
            if (in_frame && p->D > 2 && h->t_range_valid && !h->voxels_foreign) {
                // D = 3 / 4: candidate lists per (time slab, pixel tile)
                if (!h->d_prune3_lists &&
                    cudaMalloc(&h->d_prune3_lists, (size_t)EVK_PRUNE3_LISTS * 16) != cudaSuccess) {
                    cudaGetLastError();
                    h->d_prune3_lists = nullptr;
                }
                if (h->d_prune3_lists)
                    ce = evk_launch_km_assign_pruned3(kl, h->ds.width, h->ds.height,
                                                      h->d_prune3_lists, h->t_range_lo,
                                                      h->t_range_hi, xy, ev, h->d_first, n,
                                                      h->d_cent, h->d_acc, h->d_labels,
                                                      h->sm_count, h->stream);
            }
            if (ce == cudaErrorNotSupported && in_frame)  // D = 2: exact candidate pruning per tile
                ce = evk_launch_km_assign_pruned(kl, h->ds.width, h->ds.height, h->d_prune_lists,
                                                 xy, n, h->d_cent, h->d_acc, h->d_labels,
                                                 h->sm_count, h->stream);
            if (ce == cudaErrorNotSupported)
                ce = evk_launch_km_assign(kl, xy, ev, h->d_first, n, h->d_cent, h->d_acc,
                                          h->d_labels, h->sm_count, h->stream);
        }
        EVK_CUDA(h, ce);
        if (reduce) EVK_TRY(reduce(h, p->K, p->D));
        EVK_CUDA(h, evk_launch_km_finalise(kl, h->d_cent, h->d_acc, h->d_counts, h->d_shift,
                                           h->stream));
        launches += image ? (on_hist ? 3 : 5) : (n ? 2 : 1);
        it++;
        if (p->tol >= 0.f) {
            EVK_CUDA(h, cudaMemcpyAsync(h->h_shift, h->d_shift, sizeof(float),
                                        cudaMemcpyDeviceToHost, h->stream));
            EVK_CUDA(h, cudaStreamSynchronize(h->stream));
            if (*h->h_shift <= p->tol) break;
        }
    }
    if (on_hist) {  // the label map holds the assignment of the last iteration
        EVK_CUDA(h, evk_launch_km_assign_tiles(kl, h->ds.width, h->ds.height, h->d_quads,
                                               h->d_label_map, xy, n, nullptr, false, h->d_acc,
                                               h->d_labels, h->sm_count, h->stream));
        launches++;
    }
    prof_rec(h, 4);
    h->n_labels = n;
    h->labels_on_events = p->on_events != 0;
    h->km_last = *p;
    h->times.km_iters = it;
    h->times.km_launches = launches;
    if (h->profiling) {
        EVK_CUDA(h, cudaStreamSynchronize(h->stream));
        h->times.km_total_ms = prof_ms(h, 3, 4);
        h->times.km_assign_ms = h->times.km_total_ms;  // finalise is a single tiny block
    }
    if (iters_done) *iters_done = it;
    return EVK_OK;
}

int evk_kmeans(evk_handle* h, const evk_km_params* p, int* iters_done) {
    EVK_TRY(check_handle(h));
    EVK_TRY(evk_collect_pending(h));
    return evk_kmeans_run(h, p, iters_done, nullptr);
}

// ---- fused step -----------------------------------------------------------------------------
// One submission, one host synchronisation: a replayed CUDA graph of nine kernels -- bins -> slab
// -> fix-up on the main stream, first-K walk -> candidate lists -> label map -> quads beside them on
// the side stream, then assign + accumulate over the new voxels -> finalise.  (Assigning inside the
// slab kernel with consumer warps was measured slower)  Shapes the fused pass
// does not take (unordered stream, D > 2, K > 254, raw-event clustering, key space too large for
// shared memory) run the three separate calls: same results either way.
static int step_unfused(evk_handle* h, const evk_ds_params* ds, const evk_km_params* km,
                        int init_first_k, int* iters_done) {
    EVK_TRY(evk_downsample_local(h, ds));
    if (init_first_k) EVK_TRY(evk_init_centroids_first_k(h, km));
    return evk_kmeans_run(h, km, iters_done, nullptr);
}

// everything the fused step submits, on h->stream and h->side (captured into a graph by the caller)
static int enqueue_fused(evk_handle* h, const KeyParams& kp, const evk_ds_params* ds,
                         const evk_km_params* km, int init_first_k, int* launches) {
    KmLaunch kl = km_launch_params(h, km);
    EVK_CUDA(h, cudaMemsetAsync(h->d_cnt, 0, sizeof(DsCounters), h->stream));
    prof_rec(h, 0);
    EVK_CUDA(h, cudaEventRecord(h->ev_fork, h->stream));
    bool ok = false;
    EVK_TRY(evk_downsample_slab(h, kp, ds->count_repeated, &ok, launches, false, nullptr, h->d_t0));
    // side stream: everything that depends on the centroids only (first-K walk over the head
    // of the stream, candidate lists, label map, quads) runs beside the downsample
    EVK_CUDA(h, cudaStreamWaitEvent(h->side, h->ev_fork, 0));
    if (init_first_k) {
        const size_t n_scan = h->n_events < (1u << 20) ? h->n_events : (1u << 20);
        EVK_CUDA(h, evk_launch_init_first_k_walk(kp, kl, h->d_events, n_scan, h->d_cent,
                                                 &h->d_cnt->scratch[4], h->side, h->d_t0));
        (*launches)++;
    } else {  // warm start: keep a copy in case the stream check sends us to the general path
        EVK_CUDA(h, cudaMemcpyAsync(h->d_cent + EVK_MAX_K * 2, h->d_cent,
                                    (size_t)km->K * 2 * sizeof(float), cudaMemcpyDeviceToDevice,
                                    h->side));
    }
    EVK_CUDA(h, evk_launch_km_image(kl, ds->width, ds->height, h->d_prune_lists, h->d_cent,
                                    nullptr, h->d_label_map, h->d_quads, h->d_acc, h->side));
    EVK_CUDA(h, cudaEventRecord(h->ev_join, h->side));
    EVK_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    prof_rec(h, 2);
    prof_rec(h, 3);
    // one pass over the new voxels, whose count never leaves the device
    EVK_CUDA(h, evk_launch_km_assign_tiles(kl, ds->width, ds->height, h->d_quads, h->d_label_map,
                                           h->d_xy, h->n_events, &h->d_cnt->n_unique, true,
                                           h->d_acc, h->d_labels, h->sm_count, h->stream));
    // (finalising in the last CTA of the assign kernel was measured slower than this launch)
    EVK_CUDA(h, evk_launch_km_finalise(kl, h->d_cent, h->d_acc, h->d_counts, h->d_shift,
                                       h->stream));
    prof_rec(h, 4);
    *launches += 5;
    EVK_CUDA(h, cudaMemcpyAsync(h->h_cnt, h->d_cnt, sizeof(DsCounters), cudaMemcpyDeviceToHost,
                                h->stream));
    return EVK_OK;
}

// Submission half of the fused step.  A fusable shape goes out as one graph launch and the call
// returns at once (h->step_pending = 1: results are collected by evk_downsample_kmeans_wait);
// any other shape runs the three calls here, synchronously, and wait only reports their results.
// Several submissions may be queued behind each other (a pipeline of slices on one stream): every
// replay works on the events resident at ITS turn in stream order, wait reports the last one.
int evk_downsample_kmeans_submit(evk_handle* h, const evk_ds_params* ds, const evk_km_params* km,
                                 int init_first_k) {
    EVK_TRY(check_handle(h));
    EVK_TRY(km_validate(h, km));
    KeyParams kp;
    EVK_TRY(evk_make_key_params(h, ds, &kp));
    DeviceGuard g(h->device);  // before anything that may allocate (ensure_images)
    h->shard_first = h->comm ? h->shard_first : 0;
    // (a queued step leaves centroids behind: a warm start may follow it without a wait)
    if (!init_first_k && h->step_pending != 1 &&
        (!h->have_centroids || h->K != km->K || h->D != km->D))
        return evk_fail(h, EVK_ERR_STATE, "centroids for K=%d, D=%d have not been set", km->K, km->D);
    const bool fusable = km->D == 2 && !km->on_events && km->K <= 254 && h->n_events &&
                         km->iters == 1 && km->tol < 0.f &&
                         (ds->algo == EVK_ALGO_AUTO || ds->algo == EVK_ALGO_SLAB) &&
                         kp.keyfn == EVK_KEY_VOXEL && evk_slab_supported(h, kp) &&
                         ensure_images(h, ds->width, ds->height);
    if (!fusable) {
        if (h->step_pending == 1)
            EVK_TRY(evk_downsample_kmeans_wait(h, nullptr, nullptr, nullptr));
        h->step_pending = 0;
        h->step_iters = 0;
        EVK_TRY(step_unfused(h, ds, km, init_first_k, &h->step_iters));
        h->step_pending = 2;  // finished: wait has nothing to collect from the device
        return EVK_OK;
    }
    h->step_ds = *ds;
    h->step_km = *km;
    h->step_init = init_first_k;
    h->step_iters = 0;
    invalidate_results(h);
    h->ds = *ds;
    h->kp = kp;
    h->have_ds = true;
    // The whole pass is one CUDA graph, re-instantiated only when the shape of the call changes
    // (event count, parameters, profiling): a replay costs one launch instead of a dozen.
    FusedKey key;
    memset(&key, 0, sizeof key);
    key.n = h->n_events;
    key.ds = *ds;
    key.ds.t0_us = 0;  // the time origin is a device-side value (d_t0): windows replay one graph
    key.km = *km;
    key.init = init_first_k ? 1 : 0;
    key.profiling = h->profiling ? 1 : 0;
    key.shard_first = h->shard_first;
    key.image_gen = h->image_gen;
    int launches = 0;
    EVK_CUDA(h, evk_launch_set_u64(reinterpret_cast<unsigned long long*>(h->d_t0),
                                   (unsigned long long)ds->t0_us, h->stream));
    if (h->fused_exec && memcmp(&key, &h->fused_key, sizeof key) == 0) {
        launches = h->fused_launches;
    } else {
        if (h->step_pending == 1) {  // the graph in flight is about to be replaced: let it finish
            EVK_CUDA(h, cudaStreamSynchronize(h->stream));
        }
        if (h->fused_exec) cudaGraphExecDestroy(h->fused_exec);
        h->fused_exec = nullptr;
        cudaGraph_t graph = nullptr;
        EVK_CUDA(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        int st_enq = enqueue_fused(h, kp, ds, km, init_first_k, &launches);
        cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
        if (st_enq != EVK_OK) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            return st_enq;
        }
        EVK_CUDA(h, ce);
        ce = cudaGraphInstantiate(&h->fused_exec, graph, 0);
        cudaGraphDestroy(graph);
        EVK_CUDA(h, ce);
        h->fused_key = key;
        h->fused_launches = launches;
    }
    EVK_CUDA(h, cudaGraphLaunch(h->fused_exec, h->stream));
    h->steps_queued = h->step_pending == 1 ? h->steps_queued + 1 : 1;
    h->step_pending = 1;
    h->step_sharded = false;
    return EVK_OK;
}

// Collection half: the step's single host synchronisation.  A stream the slab kernel rejected
// (not time-ordered, events before t0, ...) is rerun here on the general path -- same results.
int evk_downsample_kmeans_wait(evk_handle* h, size_t* n_unique, size_t* n_repeated,
                               int* iters_done) {
    EVK_TRY(check_handle(h));
    if (!h->step_pending)
        return evk_fail(h, EVK_ERR_STATE, "evk_downsample_kmeans_wait: no step has been submitted");
    const int pending = h->step_pending;
    h->step_pending = 0;
    if (pending == 1) {
        DeviceGuard g(h->device);
        const evk_ds_params* ds = &h->step_ds;
        const evk_km_params* km = &h->step_km;
        const int init_first_k = h->step_init;
        const int launches = h->fused_launches;
        const int queued = h->steps_queued;
        h->steps_queued = 0;
        if (queued > 1)  // earlier queued steps: did the fast path reject any of them?
            EVK_CUDA(h, cudaMemcpyAsync(&h->h_cnt->scratch[5], h->d_sticky, sizeof(unsigned long long),
                                        cudaMemcpyDeviceToHost, h->stream));
        EVK_CUDA(h, cudaStreamSynchronize(h->stream));
        if (queued > 1) {
            // (the last step's own rejection is handled below by the rerun; it counts once here)
            const unsigned long long rejected = h->h_cnt->scratch[5];
            const unsigned long long last = h->h_cnt->slab_violation ? 1ull : 0ull;
            if (rejected) EVK_CUDA(h, cudaMemsetAsync(h->d_sticky, 0, sizeof(unsigned long long), h->stream));
            if (rejected > last)
                return evk_fail(h, EVK_ERR_STATE,
                                "%llu of %d queued steps were rejected by the time-slab path (slices "
                                "must be time-ordered to be queued): resubmit them one at a time",
                                rejected - last, queued);
        }
        const bool ok = h->h_cnt->slab_violation == 0 && h->h_cnt->overflow == 0 &&
                        (!init_first_k || h->h_cnt->scratch[4] == (unsigned long long)km->K);
        if (ok) {
            h->n_unique = (size_t)h->h_cnt->n_unique;
            h->n_repeated = ds->count_repeated ? (size_t)h->h_cnt->n_repeated : 0;
            h->have_voxels = true;
            if (ds->vt_us > 0 && h->n_unique) {
                h->t_range_lo = ds->t0_us + (long long)h->h_cnt->scratch[2] * ds->vt_us;
                h->t_range_hi = h->t_range_lo + (long long)h->h_cnt->scratch[0] * ds->vt_us - 1;
                h->t_range_valid = true;
            }
            h->K = km->K;
            h->D = km->D;
            h->have_centroids = true;
            h->n_labels = h->n_unique;
            h->labels_on_events = false;
            h->km_last = *km;
            h->times.ds_algo_used = EVK_ALGO_SLAB;
            h->times.ds_launches = launches - 5;
            h->times.km_launches = 5;
            h->times.km_iters = 1;
            h->times.ds_compact_ms = 0.f;
            if (h->profiling) {
                h->times.ds_total_ms = prof_ms(h, 0, 2);
                h->times.ds_main_ms = prof_ms(h, 5, 6);
                h->times.km_total_ms = h->times.km_assign_ms = prof_ms(h, 3, 4);
            }
            h->step_iters = 1;
        } else {
            EVK_CUDA(h, cudaMemsetAsync(h->d_sticky, 0, sizeof(unsigned long long), h->stream));
            if (!init_first_k) {  // finalise has overwritten the caller's centroids
                if (!h->have_centroids || h->K != km->K || h->D != km->D)
                    return evk_fail(h, EVK_ERR_STATE,
                                    "centroids for K=%d, D=%d have not been set", km->K, km->D);
                EVK_CUDA(h, cudaMemcpyAsync(h->d_cent, h->d_cent + EVK_MAX_K * 2,
                                            (size_t)km->K * 2 * sizeof(float),
                                            cudaMemcpyDeviceToDevice, h->stream));
            }
            EVK_TRY(step_unfused(h, ds, km, init_first_k, &h->step_iters));
        }
    }
    if (n_unique) *n_unique = h->n_unique;
    if (n_repeated) *n_repeated = h->n_repeated;
    if (iters_done) *iters_done = h->step_iters;
    return EVK_OK;
}

int evk_downsample_kmeans(evk_handle* h, const evk_ds_params* ds, const evk_km_params* km,
                          int init_first_k, size_t* n_unique, size_t* n_repeated,
                          int* iters_done) {
    EVK_TRY(check_handle(h));
    if (h->step_pending) EVK_TRY(evk_downsample_kmeans_wait(h, nullptr, nullptr, nullptr));
    EVK_TRY(evk_downsample_kmeans_submit(h, ds, km, init_first_k));
    return evk_downsample_kmeans_wait(h, n_unique, n_repeated, iters_done);
}

int evk_get_labels(evk_handle* h, int32_t* labels, size_t cap) {
    EVK_TRY(check_handle(h));
    EVK_TRY(evk_collect_pending(h));
    if (!h->km_last.K) return evk_fail(h, EVK_ERR_STATE, "evk_kmeans has not run");
    const size_t n = h->n_labels;
    if (cap < n) return evk_fail(h, EVK_ERR_CAPACITY, "cap %zu < %zu labels", cap, n);
    if (!n) return EVK_OK;
    if (!labels) return evk_fail(h, EVK_ERR_INVALID, "labels is NULL");
    DeviceGuard g(h->device);
    if (h->labels_on_events) {
        EVK_CUDA(h, cudaMemcpyAsync(labels, h->d_labels, n * 4, cudaMemcpyDeviceToHost, h->stream));
    } else {
        EVK_TRY(evk_ensure_perm(h));
        int32_t* scratch = reinterpret_cast<int32_t*>(h->d_sort_b);
        EVK_CUDA(h, evk_launch_gather_labels(h->d_labels, h->d_perm, scratch, n, h->stream));
        EVK_CUDA(h, cudaMemcpyAsync(labels, scratch, n * 4, cudaMemcpyDeviceToHost, h->stream));
    }
    EVK_CUDA(h, cudaStreamSynchronize(h->stream));
    return EVK_OK;
}

int evk_get_centroids(evk_handle* h, float* c, uint64_t* counts) {
    EVK_TRY(check_handle(h));
    EVK_TRY(evk_collect_pending(h));
    if (!h->have_centroids) return evk_fail(h, EVK_ERR_STATE, "no centroids");
    DeviceGuard g(h->device);
    if (c)
        EVK_CUDA(h, cudaMemcpyAsync(c, h->d_cent, (size_t)h->K * h->D * sizeof(float),
                                    cudaMemcpyDeviceToHost, h->stream));
    if (counts)
        EVK_CUDA(h, cudaMemcpyAsync(counts, h->d_counts, (size_t)h->K * sizeof(uint64_t),
                                    cudaMemcpyDeviceToHost, h->stream));
    EVK_CUDA(h, cudaStreamSynchronize(h->stream));
    return EVK_OK;
}

// ---- streaming windows ----------------------------------------------------------------------
int evk_window_config(evk_handle* h, const evk_ds_params* ds, const evk_km_params* km,
                      int64_t window_us) {
    EVK_TRY(check_handle(h));
    KeyParams kp;
    EVK_TRY(evk_make_key_params(h, ds, &kp));
    EVK_TRY(km_validate(h, km));
    if (window_us <= 0) return evk_fail(h, EVK_ERR_INVALID, "window_us must be > 0");
    if (km->on_events) return evk_fail(h, EVK_ERR_INVALID, "windows cluster voxels, not raw events");
    h->win_ds = *ds;
    h->win_km = *km;
    h->win_us = window_us;
    h->win_events = 0;
    h->win_cfg = true;
    h->win_started = false;
    h->win_pending = 0;
    h->win_count = 0;
    h->have_centroids = false;
    return EVK_OK;
}

// Condition::make_n_events(nevents) of the reslicer (SAMP/store.cpp:336): a window is complete after
// exactly n_events events; its time bins start at its first event.
int evk_window_config_events(evk_handle* h, const evk_ds_params* ds, const evk_km_params* km,
                             size_t n_events) {
    EVK_TRY(check_handle(h));
    if (n_events == 0) return evk_fail(h, EVK_ERR_INVALID, "n_events must be > 0");
    if (n_events > h->max_events)
        return evk_fail(h, EVK_ERR_CAPACITY, "window of %zu events exceeds the handle capacity %zu",
                        n_events, h->max_events);
    EVK_TRY(evk_window_config(h, ds, km, 1));
    h->win_events = n_events;
    return EVK_OK;
}

static int window_run(evk_handle* h) {
    // one slice: the body of on_new_slice (ACCEL/store.cpp:370-568) minus consumer and drawing,
    // as ONE fused submission (downsample + warm-started k-means) with one host synchronisation.
    // The slice's events are already on the device (staged as they arrived); they replace the
    // previous slice's only now, so that slice's results stay readable until this point.
    {
        DeviceGuard g(h->device);
        EVK_CUDA(h, cudaMemcpyAsync(h->d_events, h->d_win_stage, h->win_pending * sizeof(evk_event),
                                    cudaMemcpyDeviceToDevice, h->stream));
        h->n_events = h->win_pending;
        invalidate_results(h);
    }
    evk_ds_params ds = h->win_ds;
    ds.t0_us = h->win_start;  // time bins restart with every window
    const bool seeded = h->have_centroids && h->K == h->win_km.K && h->D == h->win_km.D;
    int st = evk_downsample_kmeans(h, &ds, &h->win_km, seeded ? 0 : 1, nullptr, nullptr, nullptr);
    // fewer voxels than clusters in the first windows: downsampled only, seeding waits
    if (st == EVK_ERR_INVALID && !seeded && h->have_voxels && h->n_unique < (size_t)h->win_km.K)
        st = EVK_OK;
    EVK_TRY(st);
    h->win_count++;
    h->win_pending = 0;
    return EVK_OK;
}

// Events of a push are time-ordered (the SDK delivers them so, ACCEL/store.cpp:614-615): the end of
// the current window inside the range is found by binary search and whole sub-ranges go to the
// device with one copy each -- no per-event host work, no host-side staging buffer.
int evk_window_push(evk_handle* h, const evk_event* begin, const evk_event* end, int* windows_done) {
    EVK_TRY(check_handle(h));
    if (!h->win_cfg) return evk_fail(h, EVK_ERR_STATE, "evk_window_config has not been called");
    if ((!begin && end != begin) || end < begin) return evk_fail(h, EVK_ERR_INVALID, "bad event range");
    EVK_TRY(evk_collect_pending(h));
    int done = 0;
    const evk_event* p = begin;
    while (h->win_events && p != end) {  // count-based windows
        if (h->win_pending == 0) h->win_start = p->t;
        const size_t m = std::min((size_t)(end - p), h->win_events - h->win_pending);
        DeviceGuard g(h->device);
        if (!h->d_win_stage)
            EVK_CUDA(h, cudaMalloc((void**)&h->d_win_stage, h->max_events * sizeof(evk_event)));
        EVK_CUDA(h, cudaMemcpyAsync(h->d_win_stage + h->win_pending, p, m * sizeof(evk_event),
                                    cudaMemcpyHostToDevice, h->stream));
        h->win_pending += m;
        p += m;
        if (h->win_pending == h->win_events) {
            EVK_TRY(window_run(h));
            done++;
        }
    }
    while (p != end) {
        if (!h->win_started) {
            h->win_started = true;
            h->win_start = p->t - (p->t % h->win_us);
        }
        const int64_t win_end = h->win_start + h->win_us;
        const evk_event* q =
            std::partition_point(p, end, [win_end](const evk_event& e) { return e.t < win_end; });
        if (q != p) {
            const size_t m = (size_t)(q - p);
            if (h->win_pending + m > h->max_events)
                return evk_fail(h, EVK_ERR_CAPACITY, "window exceeds the handle capacity");
            DeviceGuard g(h->device);
            if (!h->d_win_stage)
                EVK_CUDA(h, cudaMalloc((void**)&h->d_win_stage, h->max_events * sizeof(evk_event)));
            EVK_CUDA(h, cudaMemcpyAsync(h->d_win_stage + h->win_pending, p, m * sizeof(evk_event),
                                        cudaMemcpyHostToDevice, h->stream));
            h->win_pending += m;
        }
        if (q != end) {  // an event of a later window has arrived: this one is complete
            if (h->win_pending) {
                EVK_TRY(window_run(h));
                done++;
            }
            h->win_start += h->win_us;
            while (q->t >= h->win_start + h->win_us) h->win_start += h->win_us;  // empty windows
        }
        p = q;
    }
    // the ranges are borrowed: their copies must have left host memory before we return
    {
        DeviceGuard g(h->device);
        EVK_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    if (windows_done) *windows_done = done;
    return EVK_OK;
}

int evk_window_flush(evk_handle* h, int* windows_done) {
    EVK_TRY(check_handle(h));
    if (!h->win_cfg) return evk_fail(h, EVK_ERR_STATE, "evk_window_config has not been called");
    int done = 0;
    if (h->win_pending) {
        EVK_TRY(window_run(h));
        done = 1;
        if (!h->win_events) h->win_start += h->win_us;
    }
    if (windows_done) *windows_done = done;
    return EVK_OK;
}

// ---- profiling / measurement ----------------------------------------------------------------
int evk_set_profiling(evk_handle* h, int enabled) {
    EVK_TRY(check_handle(h));
    h->profiling = enabled != 0;
    return EVK_OK;
}
int evk_get_stage_times(const evk_handle* h, evk_stage_times* out) {
    if (!h || !out) return EVK_ERR_INVALID;
    *out = h->times;
    return EVK_OK;
}
int evk_timer_start(evk_handle* h) {
    EVK_TRY(check_handle(h));
    DeviceGuard g(h->device);
    EVK_CUDA(h, cudaEventRecord(h->ev_timer[0], h->stream));
    return EVK_OK;
}
int evk_timer_stop(evk_handle* h, float* elapsed_ms) {
    EVK_TRY(check_handle(h));
    DeviceGuard g(h->device);
    EVK_CUDA(h, cudaEventRecord(h->ev_timer[1], h->stream));
    EVK_CUDA(h, cudaEventSynchronize(h->ev_timer[1]));
    float ms = 0.f;
    EVK_CUDA(h, cudaEventElapsedTime(&ms, h->ev_timer[0], h->ev_timer[1]));
    if (elapsed_ms) *elapsed_ms = ms;
    return EVK_OK;
}
int evk_sync(evk_handle* h) {
    EVK_TRY(check_handle(h));
    DeviceGuard g(h->device);
    EVK_CUDA(h, cudaStreamSynchronize(h->stream));
    return EVK_OK;
}
int evk_flush_l2(evk_handle* h) {
    EVK_TRY(check_handle(h));
    DeviceGuard g(h->device);
    EVK_CUDA(h, evk_launch_fill_u8(h->d_flush, 0, h->flush_bytes, h->stream));
    return EVK_OK;
}

}  // extern "C"
