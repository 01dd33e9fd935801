// evk_internal.cuh — shared declarations of the B200 (sm_100a) implementation behind include/evk.h.
//
// Data layout in HBM (all owned by the handle, allocated once in evk_create):
//   d_events   evk_event[max_events]   16-B packed AoS records, read with one 128-bit load
//   d_tkeys    u64[table_cap]          open-addressing table keys (EMPTY = ~0; bit 63 = "hit twice")
//   d_tfirst   u32[table_cap]          lowest stream index per slot (atomicMin)
//   d_keys     u64[max_events]         voxel keys, emission order            \  SoA voxel shard:
//   d_first    u32[max_events]         lowest stream index of each voxel      | 16 B per voxel
//   d_xy       u32[max_events]         representative's x | y << 16          /
//   d_labels   i32[max_events]         k-means label of each voxel (or event)
//   d_cent     f32[K*D], d_acc u64[K*(D+1)]  centroids and exact integer partial sums
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <vector>

#include "../../include/evk.h"

#define EVK_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull
#define EVK_REP_FLAG 0x8000000000000000ull
#define EVK_EMPTY_IDX 0xFFFFFFFFu
#define EVK_MAX_K 1024
#define EVK_MAX_D 4
#define EVK_SLAB_CHUNK 8192u  // output slots per CTA-private chunk of the slab kernel

// ---- key arithmetic -------------------------------------------------------------------------
struct KeyParams {
    int32_t keyfn, width, height, use_p;
    uint32_t mx, my;  // ceil(2^32 / v) for v >= 2: __umulhi(x, m) == x / v exactly for x < 65536
    int32_t sx, sy;   // >= 0: v is a power of two, quotient is a shift
    uint32_t NX, NY, P;
    int32_t vt_shift;  // >= 0: vt is a power of two
    int64_t t0, vt;    // vt <= 0: one time bin
    uint64_t vt_magic, vt_limit;
    uint64_t cells;    // NX * NY * P  (keys per time bin)
};

__host__ __device__ __forceinline__ uint64_t evk_mix64(uint64_t z) {
    z = (z ^ (z >> 33)) * 0xFF51AFD7ED558CCDull;
    z = (z ^ (z >> 33)) * 0xC4CEB9FE1A85EC53ull;
    return z ^ (z >> 33);
}

#ifdef __CUDACC__
__device__ __forceinline__ uint4 ld_event(const evk_event* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
// fields of the 16-B record held in a uint4: x = .x & 0xffff, y = .x >> 16, p = (int16).y, t = .z|.w<<32
__device__ __forceinline__ uint32_t ev_x(const uint4& e) { return e.x & 0xFFFFu; }
__device__ __forceinline__ uint32_t ev_y(const uint4& e) { return e.x >> 16; }
__device__ __forceinline__ uint32_t ev_pbit(const uint4& e) {
    return ((int32_t)(int16_t)(e.y & 0xFFFFu)) > 0 ? 1u : 0u;
}
__device__ __forceinline__ int64_t ev_t(const uint4& e) {
    return (int64_t)((uint64_t)e.z | ((uint64_t)e.w << 32));
}

__device__ __forceinline__ uint64_t evk_tbin(const KeyParams& kp, int64_t t) {
    if (kp.vt <= 0) return 0;
    uint64_t dt = (uint64_t)(t - kp.t0);
    if (kp.vt_shift >= 0) return dt >> kp.vt_shift;
    if (dt <= kp.vt_limit) return __umul64hi(dt, kp.vt_magic);
    return dt / (uint64_t)kp.vt;
}

// spatial cell index inside one time bin: (ybin * NX + xbin) * P + pbit
__device__ __forceinline__ uint32_t evk_cell(const KeyParams& kp, const uint4& e) {
    uint32_t xb = kp.sx >= 0 ? ev_x(e) >> kp.sx : __umulhi(ev_x(e), kp.mx);
    uint32_t yb = kp.sy >= 0 ? ev_y(e) >> kp.sy : __umulhi(ev_y(e), kp.my);
    uint32_t c = yb * kp.NX + xb;
    return kp.use_p ? c * 2u + ev_pbit(e) : c;
}

// full key; returns false when the event is gated out
__device__ __forceinline__ bool evk_key(const KeyParams& kp, const uint4& e, uint64_t& key) {
    const uint32_t x = ev_x(e), y = ev_y(e);
    if (kp.keyfn == EVK_KEY_REF_HASH8192) {
        if (x > (uint32_t)kp.width || y > (uint32_t)kp.height) return false;
        key = (uint64_t)((x * 1619u + y * 31u) & 8191u);
        return true;
    }
    const int64_t t = ev_t(e);
    if (x >= (uint32_t)kp.width || y >= (uint32_t)kp.height || t < kp.t0) return false;
    key = evk_tbin(kp, t) * kp.cells + evk_cell(kp, e);
    return true;
}
#endif

// ---- handle ---------------------------------------------------------------------------------
struct DsCounters {  // device-side counters, mirrored into pinned host memory after each call
    unsigned long long n_unique;
    unsigned long long n_repeated;
    unsigned long long n_valid;
    unsigned int slab_violation;  // slab kernel found an event outside its bin's index range
    unsigned int overflow;
    unsigned long long scratch[6];
    // fused sharded step: [0] = any reason to abandon the pass (+ 2^32: a peer-memory wait timed
    // out), [1] = voxels of all ranks, [2] = repeated cells of all ranks (k_p2p_step_tail)
    unsigned long long step_tail[3];
};

struct CommState;
struct AecHost;  // evk_aec.cu
struct TsHost;   // evk_corner.cu
struct DbHost;   // evk_dbscan.cu
struct NmsHost;  // evk_corner.cu
struct OpticsHost;  // evk_optics.cu (box non-maximum suppression of corner lists)

struct FusedKey {  // what the captured fused-step graph depends on
    size_t n;
    evk_ds_params ds;
    evk_km_params km;
    int init, profiling;
    uint64_t shard_first;
    uint64_t image_gen;  // bumped when the pixel images are reallocated (their pointers are baked in)
};

struct evk_handle {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    size_t max_events = 0, n_events = 0;
    evk_event* d_events = nullptr;
    // table
    size_t table_cap = 0;
    uint64_t* d_tkeys = nullptr;
    uint32_t* d_tfirst = nullptr;
    // voxel shard (emission order)
    uint64_t* d_keys = nullptr;
    uint32_t* d_first = nullptr;
    uint32_t* d_xy = nullptr;
    evk_event* d_reps = nullptr;  // only materialised in sharded mode (voxels received from peers)
    int32_t* d_labels = nullptr;
    size_t n_unique = 0, n_repeated = 0;
    bool have_voxels = false;
    bool reps_valid = false;
    bool voxels_foreign = false;  // the shard holds voxels of other ranks' events (hash ownership)
    // canonical order (lazy): d_perm[i] = emission position of the i-th voxel by first index
    uint32_t* d_perm = nullptr;
    bool perm_valid = false;
    void* d_sort_tmp = nullptr;
    size_t sort_tmp_bytes = 0;
    uint32_t *d_sort_a = nullptr, *d_sort_b = nullptr, *d_sort_c = nullptr;  // perm scratch
    // sort-variant scratch (lazy)
    uint64_t *d_sk_in = nullptr, *d_sk_out = nullptr;
    uint32_t *d_si_in = nullptr, *d_si_out = nullptr;
    void* d_sv_tmp = nullptr;
    size_t sv_tmp_bytes = 0;
    // partition path (lazy): events in time-bin order, their original indices, strip x bin offsets
    evk_event* d_part_events = nullptr;
    uint32_t* d_part_orig = nullptr;
    uint32_t* d_part_hist = nullptr;
    size_t part_hist_cap = 0;
    evk_event* d_part_events2 = nullptr;  // two-level partition of scattered streams: level-1 copy
    uint32_t* d_part_orig2 = nullptr;
    // slab scratch
    uint32_t* d_bin_start = nullptr;  // [max_bins + 1]
    void* d_slab_scratch = nullptr;   // fix-up plan + per-CTA chunk list
    size_t out_cap = 0;               // capacity of d_keys / d_first / d_xy (max_events + slack)
    size_t max_bins = 0;
    // counters
    DsCounters* d_cnt = nullptr;
    DsCounters* h_cnt = nullptr;  // pinned
    // k-means
    evk_ds_params ds{};
    KeyParams kp{};
    bool have_ds = false;
    int K = 0, D = 0;
    bool have_centroids = false;
    float* d_cent = nullptr;                 // [EVK_MAX_K * EVK_MAX_D]
    unsigned long long* d_acc = nullptr;     // [EVK_MAX_K * (EVK_MAX_D + 1)]
    unsigned long long* d_counts = nullptr;  // [EVK_MAX_K] counts of the last iteration
    void* d_prune_lists = nullptr;           // [EVK_PRUNE_TILES] uint4 candidate lists
    void* d_prune3_lists = nullptr;          // [EVK_PRUNE3_LISTS] uint4, lazy (D == 3 / 4)
    // microsecond range of the current voxel shard's time bins (slab / partition paths)
    bool t_range_valid = false;
    long long t_range_lo = 0, t_range_hi = 0;
    // pixel-image k-means (lazy, D == 2): label of every pixel, voxel representatives per pixel
    uint8_t* d_label_map = nullptr;          // [height * width]
    uint32_t* d_pixcnt = nullptr;            // [height * width]
    uint8_t* d_quads = nullptr;              // [EVK_MAX_QUADS] label of uniformly labelled squares
    cudaStream_t side = nullptr;             // centroid-only kernels run beside the downsample
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // fixed-count Lloyd loops on the pixel histogram as one graph (evk_kmeans, iters >= 3, tol < 0)
    cudaGraphExec_t loop_exec = nullptr;
    struct LoopKey {
        int width, height, K, iters, need_hist, profiling;
        float max_dist;
        size_t cap;
        uint64_t image_gen;
    } loop_key{};
    unsigned long long* d_n_points = nullptr;  // point count read by the loop graph's kernels
    cudaGraphExec_t fused_exec = nullptr;    // the fused step as one graph (evk_downsample_kmeans)
    FusedKey fused_key{};
    int fused_launches = 0;
    // evk_downsample_kmeans_submit / _wait: 0 = nothing submitted, 1 = graph in flight, 2 = the
    // step ran synchronously on the general path (wait only reports)
    int step_pending = 0;
    // steps queued since the last wait, and a device counter of queued steps the fast path rejected
    // (never cleared by the per-step memset): wait reports them instead of silently skipping a slice
    int steps_queued = 0;
    unsigned long long* d_sticky = nullptr;
    long long* d_t0 = nullptr;  // time origin of the queued fused step (read by its kernels)
    evk_ds_params step_ds{};
    evk_km_params step_km{};
    int step_init = 0, step_iters = 0;
    bool step_sharded = false;  // the pending step was queued by the sharded submit
    size_t image_pixels = 0;                 // capacity of both
    uint64_t image_gen = 0;                  // generation of the image allocations (graph keys)
    bool pix_valid = false;                  // d_pixcnt matches the current voxel shard
    float* d_shift = nullptr;                // [1]
    float* h_shift = nullptr;                // pinned
    size_t n_labels = 0;
    bool labels_on_events = false;
    evk_km_params km_last{};
    // init-centroid scratch
    uint32_t* d_cand = nullptr;  // pairs (first, pos)
    size_t cand_cap = 0;
    // profiling
    bool profiling = false;
    evk_stage_times times{};
    cudaEvent_t ev[8] = {};
    cudaEvent_t ev_timer[2] = {};
    // L2 flush buffer
    void* d_flush = nullptr;
    size_t flush_bytes = 0;
    // streaming windows
    bool win_cfg = false;
    evk_ds_params win_ds{};
    evk_km_params win_km{};
    int64_t win_us = 0, win_start = 0;
    size_t win_events = 0;  // > 0: windows of exactly this many events (evk_window_config_events)
    bool win_started = false;
    size_t win_pending = 0;  // events of the open window already on the device (d_win_stage)
    evk_event* d_win_stage = nullptr;  // [max_events], lazy
    // RAW ingest staging (lazy, grow-only): EVT 2.0 words and the decoder's block summaries
    uint32_t* d_raw = nullptr;
    size_t raw_cap_words = 0;
    uint32_t* d_raw_blk = nullptr;
    size_t raw_cap_blocks = 0;
    size_t win_count = 0;
    // multi-GPU
    CommState* comm = nullptr;
    // asynchronous event clustering consumer (evk_aec.cu), created by evk_aec_create
    AecHost* aec = nullptr;
    // time surface + corner test (evk_corner.cu), created by evk_ts_create
    TsHost* ts = nullptr;
    NmsHost* nms = nullptr;
    OpticsHost* optics = nullptr;
    // DBSCAN buffers and last results (evk_dbscan.cu), created on first use
    DbHost* db = nullptr;
    uint64_t shard_first = 0;
    std::string err;
};

// makes the handle's device current for the scope of an entry point and restores the caller's
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        int cur = -1;
        cudaGetDevice(&cur);
        if (prev >= 0 && cur != prev) cudaSetDevice(prev);
    }
};

// ---- error handling -------------------------------------------------------------------------
int evk_fail(evk_handle* h, int code, const char* fmt, ...);
void evk_prof_rec(evk_handle* h, int i);  // cudaEventRecord(h->ev[i]) when profiling is on
#define EVK_CUDA(h, expr)                                                                    \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess)                                                               \
            return evk_fail((h), EVK_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #expr,    \
                            cudaGetErrorString(_e));                                         \
    } while (0)
#define EVK_TRY(expr)                 \
    do {                              \
        int _s = (expr);              \
        if (_s != EVK_OK) return _s;  \
    } while (0)

// ---- internal entry points shared between translation units ----------------------------------
extern "C" int evk_downsample_local(evk_handle* h, const evk_ds_params* p);
extern "C" int evk_kmeans_run(evk_handle* h, const evk_km_params* p, int* iters_done,
                              int (*reduce)(evk_handle*, int K, int D));

// host helpers of evk_api.cu used by the sharded step (evk_comm.cu)
struct KmLaunch;
int evk_km_validate(evk_handle* h, const evk_km_params* p);
extern "C" int evk_collect_pending(evk_handle* h);  // wait for a queued fused step, if any
bool evk_ensure_images(evk_handle* h, int width, int height);
void evk_invalidate_results(evk_handle* h);

// ---- kernel launchers (implemented in the .cu files) ----------------------------------------
int evk_make_key_params(evk_handle* h, const evk_ds_params* p, KeyParams* kp);
// synth
cudaError_t evk_launch_synth(const evk_synth_params& sp, evk_event* out, cudaStream_t s);
cudaError_t evk_launch_soa_pack(const uint16_t* x, const uint16_t* y, const int64_t* t,
                                const uint8_t* p, size_t n, evk_event* out, cudaStream_t s);
cudaError_t evk_launch_coords_pack(const int32_t* xy, size_t n, evk_event* out, cudaStream_t s);
// downsample: table
cudaError_t evk_launch_table_clear(uint64_t* tkeys, uint32_t* tfirst, size_t cap, cudaStream_t s);
cudaError_t evk_launch_table_insert(const KeyParams& kp, const evk_event* ev, size_t n,
                                    uint64_t* tkeys, uint32_t* tfirst, size_t cap,
                                    int count_repeated, int sm_count, cudaStream_t s);
cudaError_t evk_launch_table_compact(const evk_event* ev, uint64_t* tkeys, uint32_t* tfirst,
                                     size_t cap, uint64_t* keys, uint32_t* first, uint32_t* xy,
                                     uint32_t first_offset, DsCounters* cnt, int sm_count,
                                     cudaStream_t s);
// downsample: sort + unique
int evk_downsample_sort(evk_handle* h, const KeyParams& kp, int* launches);
// downsample: time-slab kernel.  sync = false: everything is only enqueued; the caller copies the
// counters, synchronises and reads slab_violation / overflow itself.
// range (device, sharded runs): [0] give-up flag, [3] events to skip, [4] events to keep behind n.
// t0_dev (device): when not null the time origin is read from there instead of kp.t0, so that a
// captured graph can be replayed for slices with different origins (streaming windows).
int evk_downsample_slab(evk_handle* h, const KeyParams& kp, int count_repeated, bool* ok,
                        int* launches, bool sync = true,
                        const unsigned long long* range = nullptr,
                        const long long* t0_dev = nullptr);
bool evk_slab_supported(const evk_handle* h, const KeyParams& kp);
// unordered streams: stable partition by time bin + the slab kernel (evk_partition.cu)
int evk_downsample_partitioned(evk_handle* h, const KeyParams& kp, int count_repeated, bool* ok,
                               int* launches);
void evk_partition_free(evk_handle* h);
size_t evk_slab_scratch_bytes(int sm_count);
int evk_slab_ctas_per_sm();
// canonical order
int evk_ensure_perm(evk_handle* h);
cudaError_t evk_launch_gather_voxels(const evk_handle* h, uint64_t* keys, evk_event* reps,
                                     uint32_t* first, size_t n);
cudaError_t evk_launch_gather_labels(const int32_t* labels, const uint32_t* perm, int32_t* out,
                                     size_t n, cudaStream_t s);
// k-means
// accumulator slots per cluster in global memory (d_acc)
enum { ACC_CNT = 0, ACC_X = 1, ACC_Y = 2, ACC_T = 3, ACC_P = 4, ACC_STRIDE = 5 };
struct KmLaunch {
    int K, D;
    float best2;  // gate on the squared distance (+inf: none)
    float t_scale, p_scale;
    int64_t t0;
    int write_labels;
};
KmLaunch evk_km_launch_params(const evk_handle* h, const evk_km_params* p);
#ifdef __CUDACC__
// centroid = exact sum / count rounded once to fp32; empty clusters keep their centroid;
// shift = max_k |delta c_k|_inf; accumulators are zeroed for the next iteration.  Threads
// [0, K) of one CTA (k_km_finalise; the tail kernel of the fused sharded step).
__device__ __forceinline__ void evk_km_finalise_body(const KmLaunch& kl, float* cent,
                                                     unsigned long long* acc,
                                                     unsigned long long* counts, float* shift) {
    __shared__ unsigned int s_shift;
    if (threadIdx.x == 0) s_shift = 0;
    __syncthreads();
    const int k = threadIdx.x;
    if (k < kl.K) {
        unsigned long long c = acc[k * ACC_STRIDE + ACC_CNT];
        counts[k] = c;
        float mx = 0.f;
        if (c) {
            for (int d = 0; d < kl.D; d++) {
                double s;
                if (d == 0) s = (double)acc[k * ACC_STRIDE + ACC_X];
                else if (d == 1) s = (double)acc[k * ACC_STRIDE + ACC_Y];
                else if (d == 2) s = (double)(long long)acc[k * ACC_STRIDE + ACC_T];
                else s = (double)acc[k * ACC_STRIDE + ACC_P];
                double m = s / (double)c;
                if (d == 2) m *= (double)kl.t_scale;
                if (d == 3) m *= (double)kl.p_scale;
                float nc = (float)m;
                float dl = fabsf(nc - cent[k * kl.D + d]);
                mx = fmaxf(mx, dl);
                cent[k * kl.D + d] = nc;
            }
        }
        atomicMax(&s_shift, __float_as_uint(mx));  // non-negative floats order like uints
#pragma unroll
        for (int j = 0; j < ACC_STRIDE; j++) acc[k * ACC_STRIDE + j] = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) *shift = __uint_as_float(s_shift);
}
#endif
cudaError_t evk_launch_km_assign(const KmLaunch& kl, const uint32_t* xy, const evk_event* ev,
                                 const uint32_t* first, size_t n, const float* cent,
                                 unsigned long long* acc, int32_t* labels, int sm_count,
                                 cudaStream_t s);
#define EVK_PRUNE_TILES 4096
// exact candidate pruning (evk_kmeans.cu): the frame is cut into (1 << shift)-pixel square tiles,
// tx * ty <= EVK_PRUNE_TILES; lists[tile] = up to 16 ascending centroid indices (0xFF = end,
// first byte 0xFE = scan all K)
struct PruneGrid {
    int32_t width, height, shift, tx, ty;
};
PruneGrid evk_make_prune_grid(int width, int height);
cudaError_t evk_launch_km_candidates(const KmLaunch& kl, const PruneGrid& pg, const float* cent,
                                     void* lists, cudaStream_t s);
cudaError_t evk_launch_km_assign_pruned(const KmLaunch& kl, int width, int height, void* lists,
                                        const uint32_t* xy, size_t n, const float* cent,
                                        unsigned long long* acc, int32_t* labels, int sm_count,
                                        cudaStream_t s);
// D == 3 / 4 with pruning in space and time: lists3 = EVK_PRUNE3_LISTS uint4 candidate lists
#define EVK_PRUNE3_LISTS (64 * EVK_PRUNE_TILES)
cudaError_t evk_launch_km_assign_pruned3(const KmLaunch& kl, int width, int height, void* lists3,
                                         long long t_lo, long long t_hi, const uint32_t* xy,
                                         const evk_event* ev, const uint32_t* first, size_t n,
                                         const float* cent, unsigned long long* acc,
                                         int32_t* labels, int sm_count, cudaStream_t s);
// pixel-image k-means (D == 2, K <= 254; evk_kmeans.cu)
cudaError_t evk_launch_pix_hist(const uint32_t* xy, size_t n, int width, int height,
                                uint32_t* pixcnt, int sm_count, cudaStream_t s,
                                const unsigned long long* n_dev = nullptr);
cudaError_t evk_launch_set_u64(unsigned long long* dst, unsigned long long v, cudaStream_t s);
#define EVK_MAX_QUADS 65536  // 4 x 4-px quads for a 1280 x 720 sensor (57 600)
struct QuadGrid {  // squares of (1 << shift) pixels, tx * ty <= EVK_MAX_QUADS
    int32_t width, shift, tx, ty;
};
QuadGrid evk_make_quad_grid(int width, int height);
cudaError_t evk_launch_km_image(const KmLaunch& kl, int width, int height, void* lists,
                                const float* cent, const uint32_t* pixcnt, uint8_t* map,
                                uint8_t* quads, unsigned long long* acc, cudaStream_t s);
cudaError_t evk_launch_km_assign_tiles(const KmLaunch& kl, int width, int height,
                                       const uint8_t* quads, const uint8_t* map,
                                       const uint32_t* xy, size_t n,
                                       const unsigned long long* n_dev, bool accumulate,
                                       unsigned long long* acc, int32_t* labels, int sm_count,
                                       cudaStream_t s);
cudaError_t evk_launch_km_finalise(const KmLaunch& kl, float* cent, unsigned long long* acc,
                                   unsigned long long* counts, float* shift, cudaStream_t s);
cudaError_t evk_launch_collect_below(const uint32_t* first, size_t n, uint32_t bound,
                                     uint32_t* cand, uint32_t cand_cap, unsigned long long* count,
                                     cudaStream_t s);
cudaError_t evk_launch_init_from_cand(const KmLaunch& kl, const uint32_t* cand, uint32_t n_cand,
                                      const uint32_t* xy, const evk_event* ev,
                                      const evk_event* reps, float* cent, cudaStream_t s);
cudaError_t evk_launch_init_first_k_walk(const KeyParams& kp, const KmLaunch& kl,
                                         const evk_event* ev, size_t n_scan, float* cent,
                                         unsigned long long* found, cudaStream_t s,
                                         const long long* t0_dev = nullptr);
cudaError_t evk_launch_fill_u8(void* p, int v, size_t bytes, cudaStream_t s);
// RAW EVT 2.0 decode (evk_evt2.cu): words on the device -> packed events, CD count in *total
cudaError_t evk_launch_evt2_decode(const uint32_t* words, size_t n_words, uint32_t* blk,
                                   unsigned long long* total, evk_event* out, size_t cap,
                                   cudaStream_t s);
size_t evk_evt2_blocks(size_t n_words);
// RAW EVT 3.0 decode (evk_evt3.cu): 16-bit words; blk = scratch of 7 * evk_evt3_blocks() u32
cudaError_t evk_launch_evt3_decode(const uint16_t* words, size_t n_words, uint32_t* blk,
                                   unsigned long long* total, evk_event* out, size_t cap,
                                   cudaStream_t s);
size_t evk_evt3_blocks(size_t n_words);
