// evk_corner.cu — time surface + per-event corner test on the device.
//
// Reference: the event callback of the corner tracker, event-cam-tracking/
// event-cam-fast-corner-tracker/metavision_time_surface_periodic_group_track.cpp (FCT below):
//   :786       Metavision::MostRecentTimestampBuffer time_surface(height, width, 1), zero-initialised
//   :888-923   every event of the callback range stamps its pixel: surface[y][x] = t (stream order,
//              the last one wins)
//   :931-1057  then every event of the range is tested against the UPDATED surface (the update
//              inside that loop is commented out, :935): a streak of 3..6 newest pixels on the
//              16-pixel circle of radius 3 (:44) AND a streak of 4..8 on the 20-pixel circle of
//              radius 4 (:45) -- the Arc*-style test of the event-based FAST family
//   :948-955   an event closer than 4 pixels to the border BREAKS the loop (the rest of the range is
//              not tested); the evident intent is `continue`.  Both are offered (literal_break).
// Because the surface is frozen while the range is tested, the tests are independent: one thread per
// event, 36 gathers from an L2-resident int64 image (7.4 MB for Gen4) -- a gather / issue-bound
// kernel, not an HBM-bound one.  Algorithmic bytes per event: 16 (read the event) + 8 (stamp) +
// 36 x 8 (circle reads, L2) + 1 (flag).
//   k_ts_stamp_idx   atomicMax of (stream index + 1) per pixel: which event stamps last; also the
//                    index of the first border event (atomicMin)
//   k_ts_stamp_t     the winner writes its timestamp and clears the index image (self-cleaning: no
//                    memset per call)
//   k_corner_detect  the two streak tests per event, flag + per-block corner count
//   k_corner_prefix / k_corner_scatter   ordered compaction of the corner events' stream indices
#include "evk_internal.cuh"

namespace {

constexpr int kT = 256;

__constant__ int c_circle3[16][2] = {{0, 3},  {1, 3},  {2, 2},  {3, 1},   {3, 0},   {3, -1},
                                     {2, -2}, {1, -3}, {0, -3}, {-1, -3}, {-2, -2}, {-3, -1},
                                     {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};
__constant__ int c_circle4[20][2] = {{0, 4},   {1, 4},  {2, 3},  {3, 2},  {4, 1},  {4, 0},  {4, -1},
                                     {3, -2},  {2, -3}, {1, -4}, {0, -4}, {-1, -4}, {-2, -3}, {-3, -2},
                                     {-4, -1}, {-4, 0}, {-4, 1}, {-3, 2}, {-2, 3}, {-1, 4}};

struct TsScalars {
    unsigned long long first_border;  // stream index of the first event within 4 px of the border
    unsigned long long n_corners;
};

__device__ __forceinline__ bool is_border(uint32_t x, uint32_t y, int W, int H) {
    return x < 4u || x >= (uint32_t)(W - 4) || y < 4u || y >= (uint32_t)(H - 4);
}

__global__ void __launch_bounds__(kT)
    k_ts_stamp_idx(const evk_event* __restrict__ ev, size_t n, int W, int H, uint32_t* last_idx,
                   TsScalars* sc) {
    const size_t i = (size_t)blockIdx.x * kT + threadIdx.x;
    if (i >= n) return;
    const uint4 e = ld_event(ev + i);
    const uint32_t x = ev_x(e), y = ev_y(e);
    if (x < (uint32_t)W && y < (uint32_t)H) atomicMax(&last_idx[(size_t)y * W + x], (uint32_t)i + 1u);
    if (is_border(x, y, W, H)) atomicMin(&sc->first_border, (unsigned long long)i);
}

__global__ void __launch_bounds__(kT)
    k_ts_stamp_t(const evk_event* __restrict__ ev, size_t n, int W, int H, uint32_t* last_idx,
                 long long* surf) {
    const size_t i = (size_t)blockIdx.x * kT + threadIdx.x;
    if (i >= n) return;
    const uint4 e = ld_event(ev + i);
    const uint32_t x = ev_x(e), y = ev_y(e);
    if (x >= (uint32_t)W || y >= (uint32_t)H) return;
    const size_t p = (size_t)y * W + x;
    if (last_idx[p] == (uint32_t)i + 1u) {
        surf[p] = ev_t(e);
        last_idx[p] = 0u;
    }
}

// the streak test exactly as the reference writes it (FCT:958-1003): every start, every size
template <int N, int SMIN, int SMAX>
__device__ __forceinline__ bool streak_literal(const long long (&t)[N]) {
    for (int i = 0; i < N; i++) {
        const int im1 = i == 0 ? N - 1 : i - 1;
        if (t[i] < t[im1]) continue;  // (the same for every streak size)
        for (int sz = SMIN; sz <= SMAX; sz++) {
            int a = i + sz - 1, b = i + sz;
            if (a >= N) a -= N;
            if (b >= N) b -= N;
            if (t[a] < t[b]) continue;
            long long min_t = t[i];
            int k = i;
            for (int j = 1; j < sz; j++) {
                k = k + 1 == N ? 0 : k + 1;
                min_t = t[k] < min_t ? t[k] : min_t;
            }
            bool ok = true;
            for (int j = sz; j < N; j++) {
                k = k + 1 == N ? 0 : k + 1;
                if (t[k] >= min_t) {
                    ok = false;
                    break;
                }
            }
            if (ok) return true;
        }
    }
    return false;
}

#ifndef EVK_CORNER_FAST
#define EVK_CORNER_FAST 1  // 1: the arc-growth form below (3.1x faster); 0: the literal loops
#endif

// The streak test of FCT:958-1003 (circle of N pixels, streak sizes SMIN..SMAX) on the N surface
// values t[] read around the event.  As written it tries every start and every size; what it decides
// is: "is there an arc A of SMIN..SMAX consecutive pixels with min(A) > max(all other pixels)?" (the
// two neighbour comparisons at :962 and :966 follow from that, the arc being shorter than the
// circle).  Such an arc holds the newest pixel, and growing an arc from the newest pixel by always
// taking the newer of its two neighbours reaches it (pixels inside are strictly newer than pixels
// outside), so ONE arc per size has to be examined: ~N + 2 SMAX + (N - SMAX) reads instead of the
// nested loops.  Equal to the literal loops on 200 000 random circles with ties (checked on the CPU)
// and, through the parity tests, to the oracle's literal restatement.
template <int N, int SMIN, int SMAX>
__device__ __forceinline__ bool streak(const long long (&t)[N]) {
    int m = 0;
#pragma unroll
    for (int k = 1; k < N; k++)
        if (t[k] > t[m]) m = k;
    int lo = m, hi = m;
    long long cur = t[m];
    long long mins[SMAX + 1], added[SMAX + 1];
    mins[1] = cur;
#pragma unroll
    for (int size = 2; size <= SMAX; size++) {
        const int l = lo == 0 ? N - 1 : lo - 1, r = hi == N - 1 ? 0 : hi + 1;
        const long long L = t[l], R = t[r];
        long long a;
        if (L >= R) {
            lo = l;
            a = L;
        } else {
            hi = r;
            a = R;
        }
        cur = a < cur ? a : cur;
        mins[size] = cur;
        added[size] = a;
    }
    long long rest = t[hi == N - 1 ? 0 : hi + 1];  // newest pixel outside the largest arc
    for (int k = hi == N - 1 ? 0 : hi + 1; k != lo; k = k == N - 1 ? 0 : k + 1)
        rest = t[k] > rest ? t[k] : rest;
#pragma unroll
    for (int size = SMAX; size >= SMIN; size--) {
        if (mins[size] > rest) return true;
        rest = added[size] > rest ? added[size] : rest;
    }
    return false;
}

__global__ void __launch_bounds__(kT)
    k_corner_detect(const evk_event* __restrict__ ev, size_t n, int W, int H,
                    const long long* __restrict__ surf, const TsScalars* sc, int literal_break,
                    uint8_t* flags, uint32_t* blk_cnt) {
    const size_t i = (size_t)blockIdx.x * kT + threadIdx.x;
    const size_t limit = literal_break ? (size_t)min(sc->first_border, (unsigned long long)n) : n;
    bool corner = false;
    if (i < limit) {
        const uint4 e = ld_event(ev + i);
        const int x = (int)ev_x(e), y = (int)ev_y(e);
        if (!is_border((uint32_t)x, (uint32_t)y, W, H)) {
            long long t3[16];
#pragma unroll
            for (int k = 0; k < 16; k++)
                t3[k] = surf[(size_t)(y + c_circle3[k][0]) * W + (x + c_circle3[k][1])];
            if (EVK_CORNER_FAST ? streak<16, 3, 6>(t3) : streak_literal<16, 3, 6>(t3)) {
                long long t4[20];
#pragma unroll
                for (int k = 0; k < 20; k++)
                    t4[k] = surf[(size_t)(y + c_circle4[k][0]) * W + (x + c_circle4[k][1])];
                corner = EVK_CORNER_FAST ? streak<20, 4, 8>(t4) : streak_literal<20, 4, 8>(t4);
            }
        }
    }
    if (i < n) flags[i] = corner ? 1 : 0;
    const int c = __syncthreads_count(corner);
    if (threadIdx.x == 0) blk_cnt[blockIdx.x] = (uint32_t)c;
}

// one CTA: exclusive scan of the per-block corner counts, in place; total -> sc->n_corners
__global__ void __launch_bounds__(1024)
    k_corner_prefix(uint32_t* blk, uint32_t nb, TsScalars* sc) {
    __shared__ unsigned long long s_sum[1024];
    const uint32_t per = (nb + 1023) / 1024;
    const uint32_t b0 = min(nb, threadIdx.x * per), b1 = min(nb, b0 + per);
    unsigned long long sum = 0;
    for (uint32_t b = b0; b < b1; b++) sum += blk[b];
    s_sum[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        unsigned long long a = 0;
        if ((int)threadIdx.x >= o) a = s_sum[threadIdx.x - o];
        __syncthreads();
        s_sum[threadIdx.x] += a;
        __syncthreads();
    }
    unsigned long long off = threadIdx.x ? s_sum[threadIdx.x - 1] : 0ull;
    for (uint32_t b = b0; b < b1; b++) {
        const uint32_t c = blk[b];
        blk[b] = (uint32_t)off;
        off += c;
    }
    if (threadIdx.x == 1023) sc->n_corners = s_sum[1023];
}

__global__ void __launch_bounds__(kT)
    k_corner_scatter(const uint8_t* __restrict__ flags, size_t n, const uint32_t* __restrict__ blk_off,
                     uint32_t* out) {
    __shared__ uint32_t s_w[kT / 32];
    const size_t i = (size_t)blockIdx.x * kT + threadIdx.x;
    const bool c = i < n && flags[i];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t bal = __ballot_sync(0xffffffffu, c);
    if (lane == 0) s_w[wid] = __popc(bal);
    __syncthreads();
    uint32_t base = blk_off[blockIdx.x];
    for (int k = 0; k < wid; k++) base += s_w[k];
    if (c) out[base + __popc(bal & ((1u << lane) - 1u))] = (uint32_t)i;
}

}  // namespace

struct TsHost {
    int W = 0, H = 0;
    long long* d_surf = nullptr;     // [H][W] most recent timestamp per pixel
    uint32_t* d_last = nullptr;      // [H][W] scratch: index + 1 of the last event at the pixel
    uint8_t* d_flags = nullptr;      // [max_events]
    uint32_t* d_blk = nullptr;       // [max_events / kT + 1]
    uint32_t* d_corners = nullptr;   // [max_events] stream indices of the corner events
    TsScalars* d_sc = nullptr;
    size_t n_corners = 0;
    bool have = false;
};

extern "C" {

int evk_ts_destroy(evk_handle* h) {
    if (!h || !h->ts) return EVK_OK;
    TsHost* t = h->ts;
    if (h->stream) cudaStreamSynchronize(h->stream);
    void* ptrs[] = {t->d_surf, t->d_last, t->d_flags, t->d_blk, t->d_corners, t->d_sc};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    delete t;
    h->ts = nullptr;
    return EVK_OK;
}

int evk_ts_create(evk_handle* h, int width, int height) {
    if (!h) return EVK_ERR_INVALID;
    if (width < 9 || height < 9 || width > 65535 || height > 65535)
        return evk_fail(h, EVK_ERR_INVALID, "evk_ts_create: %d x %d", width, height);
    evk_ts_destroy(h);
    cudaSetDevice(h->device);
    TsHost* t = new TsHost;
    t->W = width;
    t->H = height;
    const size_t px = (size_t)width * height, ne = h->max_events ? h->max_events : 1;
    cudaError_t ce = cudaMalloc((void**)&t->d_surf, px * sizeof(long long));
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&t->d_last, px * sizeof(uint32_t));
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&t->d_flags, ne);
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&t->d_blk, (ne / kT + 2) * sizeof(uint32_t));
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&t->d_corners, ne * sizeof(uint32_t));
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&t->d_sc, sizeof(TsScalars));
    if (ce == cudaSuccess) ce = cudaMemsetAsync(t->d_surf, 0, px * sizeof(long long), h->stream);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(t->d_last, 0, px * sizeof(uint32_t), h->stream);
    h->ts = t;
    if (ce != cudaSuccess) {
        cudaGetLastError();
        evk_ts_destroy(h);
        return evk_fail(h, EVK_ERR_NOMEM, "evk_ts_create: %s", cudaGetErrorString(ce));
    }
    return EVK_OK;
}

int evk_ts_corners(evk_handle* h, int literal_break, size_t* n_corners) {
    if (!h) return EVK_ERR_INVALID;
    if (!h->ts) return evk_fail(h, EVK_ERR_STATE, "evk_ts_create has not been called");
    TsHost* t = h->ts;
    cudaSetDevice(h->device);
    const size_t n = h->n_events;
    t->n_corners = 0;
    t->have = true;
    if (n_corners) *n_corners = 0;
    if (n == 0) return EVK_OK;
    const unsigned nb = (unsigned)((n + kT - 1) / kT);
    const TsScalars init = {~0ull, 0ull};
    EVK_CUDA(h, cudaMemcpyAsync(t->d_sc, &init, sizeof init, cudaMemcpyHostToDevice, h->stream));
    k_ts_stamp_idx<<<nb, kT, 0, h->stream>>>(h->d_events, n, t->W, t->H, t->d_last, t->d_sc);
    k_ts_stamp_t<<<nb, kT, 0, h->stream>>>(h->d_events, n, t->W, t->H, t->d_last, t->d_surf);
    k_corner_detect<<<nb, kT, 0, h->stream>>>(h->d_events, n, t->W, t->H, t->d_surf, t->d_sc,
                                              literal_break, t->d_flags, t->d_blk);
    k_corner_prefix<<<1, 1024, 0, h->stream>>>(t->d_blk, nb, t->d_sc);
    k_corner_scatter<<<nb, kT, 0, h->stream>>>(t->d_flags, n, t->d_blk, t->d_corners);
    EVK_CUDA(h, cudaGetLastError());
    TsScalars sc;
    EVK_CUDA(h, cudaMemcpyAsync(&sc, t->d_sc, sizeof sc, cudaMemcpyDeviceToHost, h->stream));
    EVK_CUDA(h, cudaStreamSynchronize(h->stream));
    t->n_corners = (size_t)sc.n_corners;
    if (n_corners) *n_corners = t->n_corners;
    return EVK_OK;
}

int evk_ts_get_corners(evk_handle* h, uint32_t* event_index, size_t cap) {
    if (!h) return EVK_ERR_INVALID;
    if (!h->ts || !h->ts->have) return evk_fail(h, EVK_ERR_STATE, "evk_ts_corners has not run");
    TsHost* t = h->ts;
    if (t->n_corners == 0) return EVK_OK;
    if (!event_index || cap < t->n_corners)
        return evk_fail(h, EVK_ERR_CAPACITY, "%zu corners, room for %zu", t->n_corners, cap);
    cudaSetDevice(h->device);
    EVK_CUDA(h, cudaMemcpyAsync(event_index, t->d_corners, t->n_corners * sizeof(uint32_t),
                                cudaMemcpyDeviceToHost, h->stream));
    EVK_CUDA(h, cudaStreamSynchronize(h->stream));
    return EVK_OK;
}

int evk_ts_get_surface(evk_handle* h, int64_t* out, size_t cap_pixels) {
    if (!h) return EVK_ERR_INVALID;
    if (!h->ts) return evk_fail(h, EVK_ERR_STATE, "evk_ts_create has not been called");
    TsHost* t = h->ts;
    const size_t px = (size_t)t->W * t->H;
    if (!out || cap_pixels < px) return evk_fail(h, EVK_ERR_CAPACITY, "surface of %zu pixels", px);
    cudaSetDevice(h->device);
    EVK_CUDA(h, cudaMemcpyAsync(out, t->d_surf, px * sizeof(int64_t), cudaMemcpyDeviceToHost,
                                h->stream));
    EVK_CUDA(h, cudaStreamSynchronize(h->stream));
    return EVK_OK;
}

}  // extern "C"

// ---- corner post-processing: box non-maximum suppression -----------------------------------------
// CornerFilter::filterCorners of the reference's corner tracker (FCT:81-151, call site FCT:832 with
// box size 15): corners in input order; a corner is kept when its box [x - h, x + h] x [y - h, y + h]
// (h = box / 2, clipped to the image) touches no box of a corner kept before it; kept corners are
// labelled 0, 1, 2, ... in order.  The reference walks the list with a mask image.  Here the same
// greedy rule runs on the pairwise relation "boxes intersect":
//   k_nms_boxes   the clipped box of every corner
//   k_nms_mask    upper triangle of the n x n conflict matrix as 64-bit words (one thread per word)
//   k_nms_scan    one CTA walks the corners 64 at a time: the 64 x 64 diagonal block is resolved
//                 sequentially from shared memory by one thread (the only serial part), then every
//                 thread ORs the rows of the corners just kept into its word of the "suppressed" set
// Exact for any order and density; n <= 32768 corners per call (the matrix is n^2 / 8 bytes).
namespace {

constexpr int kNmsMaxWords = 512;  // 64-bit words of the suppressed set = threads of the scan CTA

__global__ void __launch_bounds__(256)
    k_nms_xy_from_events(const evk_event* __restrict__ ev, const uint32_t* __restrict__ idx,
                         uint32_t n, int32_t* xy) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t w = *reinterpret_cast<const uint32_t*>(ev + idx[i]);
    xy[2 * i] = (int32_t)(w & 0xFFFFu);
    xy[2 * i + 1] = (int32_t)(w >> 16);
}

__global__ void __launch_bounds__(256)
    k_nms_boxes(const int32_t* __restrict__ xy, uint32_t n, int W, int H, int half, int4* box) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int x = xy[2 * i], y = xy[2 * i + 1];
    // (64-bit: coordinates come from the caller)
    const long long sx = max(0ll, (long long)x - half), ex = min((long long)W - 1, (long long)x + half);
    const long long sy = max(0ll, (long long)y - half), ey = min((long long)H - 1, (long long)y + half);
    const bool empty = sx > ex || sy > ey;  // centre outside the image: touches and marks nothing
    box[i] = empty ? make_int4(1, 0, 1, 0) : make_int4((int)sx, (int)ex, (int)sy, (int)ey);
}

// mask[i][w] bit b: corner j = 64 w + b comes after corner i and their boxes intersect
__global__ void __launch_bounds__(256)
    k_nms_mask(const int4* __restrict__ box, uint32_t n, uint32_t n_w, unsigned long long* mask) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t i = blockIdx.y;
    if (w >= n_w || w < (i >> 6)) return;
    const int4 a = box[i];
    unsigned long long bits = 0;
    const uint32_t j0 = w << 6;
#pragma unroll 4
    for (uint32_t b = 0; b < 64; b++) {
        const uint32_t j = j0 + b;
        if (j <= i || j >= n) continue;
        const int4 c = box[j];
        const bool hit = max(a.x, c.x) <= min(a.y, c.y) && max(a.z, c.z) <= min(a.w, c.w);
        bits |= (unsigned long long)hit << b;
    }
    mask[(size_t)i * n_w + w] = bits;
}

__global__ void __launch_bounds__(kNmsMaxWords)
    k_nms_scan(const unsigned long long* __restrict__ mask, uint32_t n, uint32_t n_w,
               uint32_t* kept, unsigned long long* n_kept) {
    __shared__ unsigned long long s_diag[64];
    __shared__ unsigned long long s_rem, s_keep;
    __shared__ uint32_t s_count;
    const uint32_t t = threadIdx.x;
    unsigned long long rem = 0;  // suppressed corners of word t
    if (t == 0) s_count = 0;
    for (uint32_t W0 = 0; W0 < n_w; W0++) {
        const uint32_t base = W0 << 6;
        if (t == W0) s_rem = rem;
        if (t < 64) s_diag[t] = base + t < n ? mask[(size_t)(base + t) * n_w + W0] : 0ull;
        __syncthreads();
        if (t == 0) {  // the 64 corners of this word, in order
            unsigned long long r = s_rem, keep = 0;
            uint32_t c = s_count;
            const uint32_t m = n - base < 64 ? n - base : 64;
            for (uint32_t b = 0; b < m; b++) {
                if (!((r >> b) & 1ull)) {
                    keep |= 1ull << b;
                    r |= s_diag[b];
                    kept[c++] = base + b;
                }
            }
            s_keep = keep;
            s_count = c;
        }
        __syncthreads();
        if (t > W0 && t < n_w) {  // the kept corners suppress later ones
            unsigned long long k = s_keep;
            while (k) {
                const int b = __ffsll((long long)k) - 1;
                k &= k - 1;
                rem |= mask[(size_t)(base + b) * n_w + t];
            }
        }
        __syncthreads();
    }
    if (t == 0) *n_kept = s_count;
}

__global__ void __launch_bounds__(256)
    k_nms_emit(const int32_t* __restrict__ xy, const uint32_t* __restrict__ kept, uint32_t m,
               evk_corner* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    evk_corner c;
    c.x = xy[2 * kept[i]];
    c.y = xy[2 * kept[i] + 1];
    c.label = (int32_t)i;
    out[i] = c;
}

}  // namespace

struct NmsHost {
    size_t cap = 0;  // corners
    int32_t* d_xy = nullptr;
    int4* d_box = nullptr;
    unsigned long long* d_mask = nullptr;
    size_t mask_words = 0;
    uint32_t* d_kept = nullptr;
    evk_corner* d_out = nullptr;
    unsigned long long* d_n = nullptr;
    size_t n_kept = 0;
    bool have = false;
};

static int nms_reserve(evk_handle* h, size_t n) {
    if (!h->nms) h->nms = new NmsHost;
    NmsHost* s = h->nms;
    const size_t n_w = (n + 63) / 64;
    if (s->cap < n) {
        void* ptrs[] = {s->d_xy, s->d_box, s->d_kept, s->d_out};
        for (void* p : ptrs)
            if (p) cudaFree(p);
        s->cap = 0;
        const size_t c = n < 1024 ? 1024 : n;
        EVK_CUDA(h, cudaMalloc((void**)&s->d_xy, c * 2 * sizeof(int32_t)));
        EVK_CUDA(h, cudaMalloc((void**)&s->d_box, c * sizeof(int4)));
        EVK_CUDA(h, cudaMalloc((void**)&s->d_kept, c * sizeof(uint32_t)));
        EVK_CUDA(h, cudaMalloc((void**)&s->d_out, c * sizeof(evk_corner)));
        s->cap = c;
    }
    if (s->mask_words < n * n_w) {
        if (s->d_mask) cudaFree(s->d_mask);
        s->d_mask = nullptr;
        s->mask_words = 0;
        EVK_CUDA(h, cudaMalloc((void**)&s->d_mask, n * n_w * sizeof(unsigned long long)));
        s->mask_words = n * n_w;
    }
    if (!s->d_n) EVK_CUDA(h, cudaMalloc((void**)&s->d_n, sizeof(unsigned long long)));
    return EVK_OK;
}

// the corners are in s->d_xy[0 .. n)
static int nms_run(evk_handle* h, size_t n, int width, int height, int box_size, size_t* n_kept) {
    NmsHost* s = h->nms;
    s->n_kept = 0;
    s->have = true;
    if (n_kept) *n_kept = 0;
    if (n == 0) return EVK_OK;
    const uint32_t n32 = (uint32_t)n, n_w = (uint32_t)((n + 63) / 64);
    k_nms_boxes<<<(n32 + 255) / 256, 256, 0, h->stream>>>(s->d_xy, n32, width, height, box_size / 2,
                                                         s->d_box);
    k_nms_mask<<<dim3((n_w + 255) / 256, n32), 256, 0, h->stream>>>(s->d_box, n32, n_w, s->d_mask);
    k_nms_scan<<<1, kNmsMaxWords, 0, h->stream>>>(s->d_mask, n32, n_w, s->d_kept, s->d_n);
    EVK_CUDA(h, cudaGetLastError());
    unsigned long long m = 0;
    EVK_CUDA(h, cudaMemcpyAsync(&m, s->d_n, sizeof m, cudaMemcpyDeviceToHost, h->stream));
    EVK_CUDA(h, cudaStreamSynchronize(h->stream));
    if (m) {
        k_nms_emit<<<((uint32_t)m + 255) / 256, 256, 0, h->stream>>>(s->d_xy, s->d_kept, (uint32_t)m,
                                                                    s->d_out);
        EVK_CUDA(h, cudaGetLastError());
    }
    s->n_kept = (size_t)m;
    if (n_kept) *n_kept = s->n_kept;
    return EVK_OK;
}

extern "C" {

int evk_filter_corners_destroy(evk_handle* h) {
    if (!h || !h->nms) return EVK_OK;
    NmsHost* s = h->nms;
    if (h->stream) cudaStreamSynchronize(h->stream);
    void* ptrs[] = {s->d_xy, s->d_box, s->d_mask, s->d_kept, s->d_out, s->d_n};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    delete s;
    h->nms = nullptr;
    return EVK_OK;
}

int evk_filter_corners(evk_handle* h, const int32_t* xy, size_t n, int width, int height,
                       int box_size, size_t* n_kept) {
    if (!h) return EVK_ERR_INVALID;
    if (n && !xy) return evk_fail(h, EVK_ERR_INVALID, "xy is NULL");
    if (width < 1 || height < 1 || box_size < 0)
        return evk_fail(h, EVK_ERR_INVALID, "evk_filter_corners: %d x %d, box %d", width, height, box_size);
    if (n > (size_t)kNmsMaxWords * 64)
        return evk_fail(h, EVK_ERR_CAPACITY, "at most %d corners per call", kNmsMaxWords * 64);
    DeviceGuard g(h->device);
    EVK_TRY(nms_reserve(h, n ? n : 1));
    if (n)
        EVK_CUDA(h, cudaMemcpyAsync(h->nms->d_xy, xy, n * 2 * sizeof(int32_t), cudaMemcpyHostToDevice,
                                    h->stream));
    return nms_run(h, n, width, height, box_size, n_kept);
}

int evk_ts_filter_corners(evk_handle* h, int box_size, size_t* n_kept) {
    if (!h) return EVK_ERR_INVALID;
    if (!h->ts || !h->ts->have) return evk_fail(h, EVK_ERR_STATE, "evk_ts_corners has not run");
    if (box_size < 0) return evk_fail(h, EVK_ERR_INVALID, "box size %d", box_size);
    TsHost* t = h->ts;
    const size_t n = t->n_corners;
    if (n > (size_t)kNmsMaxWords * 64)
        return evk_fail(h, EVK_ERR_CAPACITY, "%zu corners: at most %d per call", n, kNmsMaxWords * 64);
    DeviceGuard g(h->device);
    EVK_TRY(nms_reserve(h, n ? n : 1));
    if (n) {
        k_nms_xy_from_events<<<((uint32_t)n + 255) / 256, 256, 0, h->stream>>>(
            h->d_events, t->d_corners, (uint32_t)n, h->nms->d_xy);
        EVK_CUDA(h, cudaGetLastError());
    }
    return nms_run(h, n, t->W, t->H, box_size, n_kept);
}

int evk_get_filtered_corners(evk_handle* h, evk_corner* out, size_t cap) {
    if (!h) return EVK_ERR_INVALID;
    if (!h->nms || !h->nms->have) return evk_fail(h, EVK_ERR_STATE, "no corner list has been filtered");
    NmsHost* s = h->nms;
    if (s->n_kept == 0) return EVK_OK;
    if (!out || cap < s->n_kept)
        return evk_fail(h, EVK_ERR_CAPACITY, "%zu corners, room for %zu", s->n_kept, cap);
    DeviceGuard g(h->device);
    EVK_CUDA(h, cudaMemcpyAsync(out, s->d_out, s->n_kept * sizeof(evk_corner), cudaMemcpyDeviceToHost,
                                h->stream));
    EVK_CUDA(h, cudaStreamSynchronize(h->stream));
    return EVK_OK;
}

}  // extern "C"

