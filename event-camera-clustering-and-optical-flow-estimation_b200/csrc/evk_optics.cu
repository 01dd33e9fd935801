// evk_optics.cu — OPTICS reachability ordering of a (coarse) cloud: SURVEY 8f rank 4, second half.
//
// The reference orders event coordinates with OPTICS (event-cam-clustering/optics-clustering/
// include/optics/optics.hpp:413-590, app: test/cluster_event_data.cpp:333-338, min_pts 2, epsilon 10,
// threshold 10 on 6573 integer points).  Its cost is the neighbourhood work -- an epsilon query per
// point and the core distance (the neighbour of rank min_pts - 1) -- which is independent per point
// and runs here on the device: one thread per point streams the cloud through shared memory in tiles
// (k_optics_count: neighbour count + the min_pts smallest squared distances in registers;
// k_optics_fill: the neighbour lists in ascending index order, CSR).  The ordering itself is a
// priority-queue walk in which every step depends on the previous pop (a std::set in the reference):
// it runs on the host over the device-built lists, with the reference's total order (reachability,
// then index) -- the result does not depend on the order inside a neighbour list.
// Integer coordinates, exact squared distances (64-bit), distances = IEEE sqrt in double: the same
// numbers the reference computes.  Clouds of up to 65536 points (the all-pairs pass is O(n^2)).
#include <math.h>

#include <set>
#include <utility>

#include "evk_internal.cuh"

namespace {

constexpr int kT = 256;
constexpr int kMaxMinPts = 32;
constexpr size_t kMaxPoints = 65536;

template <int D>
__global__ void __launch_bounds__(kT)
    k_optics_count(const int32_t* __restrict__ pts, uint32_t n, double r2, int min_pts,
                   uint32_t* count, long long* core_sq) {
    __shared__ int32_t s_p[kT * 3];
    const uint32_t i = blockIdx.x * kT + threadIdx.x;
    int32_t a[3] = {0, 0, 0};
    if (i < n)
        for (int d = 0; d < D; d++) a[d] = pts[(size_t)i * D + d];
    long long best[kMaxMinPts];  // the min_pts smallest squared distances, ascending
    for (int k = 0; k < kMaxMinPts; k++) best[k] = 0x7FFFFFFFFFFFFFFFll;
    uint32_t c = 0;
    for (uint32_t j0 = 0; j0 < n; j0 += kT) {
        const uint32_t j = j0 + threadIdx.x;
        for (int d = 0; d < D; d++) s_p[threadIdx.x * 3 + d] = j < n ? pts[(size_t)j * D + d] : 0;
        __syncthreads();
        const uint32_t m = n - j0 < (uint32_t)kT ? n - j0 : (uint32_t)kT;
        if (i < n)
            for (uint32_t q = 0; q < m; q++) {
                long long s = 0;
#pragma unroll
                for (int d = 0; d < D; d++) {
                    const long long v = (long long)a[d] - s_p[q * 3 + d];
                    s += v * v;
                }
                if ((double)s <= r2) {
                    c++;
                    if (s < best[min_pts - 1]) {  // insertion into the short sorted list
                        int k = min_pts - 1;
                        while (k > 0 && best[k - 1] > s) {
                            best[k] = best[k - 1];
                            k--;
                        }
                        best[k] = s;
                    }
                }
            }
        __syncthreads();
    }
    if (i < n) {
        count[i] = c;
        core_sq[i] = c >= (uint32_t)min_pts ? best[min_pts - 1] : -1;
    }
}

// exclusive scan of count[0..n) into off[0..n] (one CTA; n <= 65536)
__global__ void __launch_bounds__(1024) k_optics_scan(const uint32_t* count, uint32_t n,
                                                      unsigned long long* off) {
    __shared__ unsigned long long s_w[32];
    __shared__ unsigned long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t b0 = 0; b0 <= n; b0 += 1024) {
        const uint32_t b = b0 + threadIdx.x;
        const unsigned long long v = b < n ? count[b] : 0ull;
        unsigned long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_w[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            unsigned long long w = s_w[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            s_w[lane] = w;
        }
        __syncthreads();
        const unsigned long long carry = s_carry;
        if (b <= n) off[b] = carry + (wid ? s_w[wid - 1] : 0ull) + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + s_w[31];
        __syncthreads();
    }
}

template <int D>
__global__ void __launch_bounds__(kT)
    k_optics_fill(const int32_t* __restrict__ pts, uint32_t n, double r2,
                  const unsigned long long* __restrict__ off, uint32_t* nb) {
    __shared__ int32_t s_p[kT * 3];
    const uint32_t i = blockIdx.x * kT + threadIdx.x;
    int32_t a[3] = {0, 0, 0};
    if (i < n)
        for (int d = 0; d < D; d++) a[d] = pts[(size_t)i * D + d];
    unsigned long long w = i < n ? off[i] : 0ull;
    for (uint32_t j0 = 0; j0 < n; j0 += kT) {
        const uint32_t j = j0 + threadIdx.x;
        for (int d = 0; d < D; d++) s_p[threadIdx.x * 3 + d] = j < n ? pts[(size_t)j * D + d] : 0;
        __syncthreads();
        const uint32_t m = n - j0 < (uint32_t)kT ? n - j0 : (uint32_t)kT;
        if (i < n)
            for (uint32_t q = 0; q < m; q++) {
                long long s = 0;
#pragma unroll
                for (int d = 0; d < D; d++) {
                    const long long v = (long long)a[d] - s_p[q * 3 + d];
                    s += v * v;
                }
                if ((double)s <= r2) nb[w++] = j0 + q;
            }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kT)
    k_optics_voxel_points(const uint32_t* __restrict__ xy, const uint32_t* __restrict__ perm,
                          uint32_t n, int32_t* pts) {
    const uint32_t i = blockIdx.x * kT + threadIdx.x;
    if (i >= n) return;
    const uint32_t w = xy[perm[i]];  // canonical order
    pts[2 * i] = (int32_t)(w & 0xFFFFu);
    pts[2 * i + 1] = (int32_t)(w >> 16);
}

}  // namespace

struct OpticsHost {
    size_t n = 0;
    std::vector<uint32_t> order;
    std::vector<double> reach;  // by ordering position; -1 = none
    bool have = false;
};

// d_pts: n x D int32 on the device
static int optics_run(evk_handle* h, const int32_t* d_pts, std::vector<int32_t>& pts, size_t n, int D,
                      int min_pts, double eps) {
    if (!h->optics) h->optics = new OpticsHost;
    OpticsHost* o = h->optics;
    o->n = n;
    o->order.assign(n, 0);
    o->reach.assign(n, -1.0);
    o->have = true;
    if (n == 0) return EVK_OK;
    const uint32_t n32 = (uint32_t)n, grid = (n32 + kT - 1) / kT;
    const double r2 = eps * eps;
    uint32_t* d_count = nullptr;
    long long* d_core = nullptr;
    unsigned long long* d_off = nullptr;
    uint32_t* d_nb = nullptr;
    auto cleanup = [&]() {
        void* p[] = {d_count, d_core, d_off, d_nb};
        for (void* q : p)
            if (q) cudaFree(q);
    };
    cudaError_t ce = cudaMalloc((void**)&d_count, n * sizeof(uint32_t));
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&d_core, n * sizeof(long long));
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&d_off, (n + 1) * sizeof(unsigned long long));
    if (ce != cudaSuccess) {
        cleanup();
        cudaGetLastError();
        return evk_fail(h, EVK_ERR_NOMEM, "OPTICS scratch");
    }
    if (D == 2) k_optics_count<2><<<grid, kT, 0, h->stream>>>(d_pts, n32, r2, min_pts, d_count, d_core);
    else k_optics_count<3><<<grid, kT, 0, h->stream>>>(d_pts, n32, r2, min_pts, d_count, d_core);
    k_optics_scan<<<1, 1024, 0, h->stream>>>(d_count, n32, d_off);
    std::vector<unsigned long long> off(n + 1);
    std::vector<long long> core_sq(n);
    ce = cudaMemcpyAsync(off.data(), d_off, (n + 1) * sizeof(unsigned long long),
                         cudaMemcpyDeviceToHost, h->stream);
    if (ce == cudaSuccess)
        ce = cudaMemcpyAsync(core_sq.data(), d_core, n * sizeof(long long), cudaMemcpyDeviceToHost,
                             h->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream);
    const unsigned long long E = ce == cudaSuccess ? off[n] : 0;
    if (ce == cudaSuccess && E > (1ull << 31)) {
        cleanup();
        return evk_fail(h, EVK_ERR_CAPACITY, "OPTICS: %llu neighbour pairs (epsilon too large for this cloud)", E);
    }
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&d_nb, (E ? E : 1) * sizeof(uint32_t));
    std::vector<uint32_t> nb((size_t)E);
    if (ce == cudaSuccess) {
        if (D == 2) k_optics_fill<2><<<grid, kT, 0, h->stream>>>(d_pts, n32, r2, d_off, d_nb);
        else k_optics_fill<3><<<grid, kT, 0, h->stream>>>(d_pts, n32, r2, d_off, d_nb);
        ce = cudaGetLastError();
    }
    if (ce == cudaSuccess && E)
        ce = cudaMemcpyAsync(nb.data(), d_nb, (size_t)E * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                             h->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream);
    cleanup();
    if (ce != cudaSuccess) {
        cudaGetLastError();
        return evk_fail(h, EVK_ERR_CUDA, "OPTICS neighbourhoods: %s", cudaGetErrorString(ce));
    }
    // the walk (optics.hpp:522-560): seeds ordered by (reachability, index)
    std::vector<double> reach(n, -1.0);
    std::vector<char> done(n, 0);
    std::set<std::pair<double, uint32_t>> seeds;
    auto dist = [&](size_t a, size_t b) {
        long long s = 0;
        for (int d = 0; d < D; d++) {
            const long long v = (long long)pts[a * D + d] - pts[b * D + d];
            s += v * v;
        }
        return sqrt((double)s);
    };
    auto update = [&](size_t p) {
        if (core_sq[p] < 0) return;
        const double core = sqrt((double)core_sq[p]);
        for (unsigned long long q = off[p]; q < off[p + 1]; q++) {
            const uint32_t v = nb[(size_t)q];
            if (done[v]) continue;
            const double dd = dist(p, v), nr = core > dd ? core : dd;
            if (reach[v] < 0.0) {
                reach[v] = nr;
                seeds.insert({nr, v});
            } else if (nr < reach[v]) {
                seeds.erase({reach[v], v});
                reach[v] = nr;
                seeds.insert({nr, v});
            }
        }
    };
    size_t emitted = 0;
    for (size_t start = 0; start < n; start++) {
        if (done[start]) continue;
        done[start] = 1;
        o->order[emitted++] = (uint32_t)start;
        update(start);
        while (!seeds.empty()) {
            const uint32_t s = seeds.begin()->second;
            seeds.erase(seeds.begin());
            done[s] = 1;
            o->order[emitted++] = s;
            update(s);
        }
    }
    for (size_t k = 0; k < n; k++) o->reach[k] = reach[o->order[k]];
    return EVK_OK;
}

extern "C" {

int evk_optics_destroy(evk_handle* h) {
    if (!h || !h->optics) return EVK_OK;
    delete h->optics;
    h->optics = nullptr;
    return EVK_OK;
}

int evk_optics_points(evk_handle* h, const int32_t* pts, size_t n, int D, int min_pts, double eps) {
    if (!h) return EVK_ERR_INVALID;
    if ((n && !pts) || (D != 2 && D != 3) || min_pts < 1 || min_pts > kMaxMinPts || !(eps > 0))
        return evk_fail(h, EVK_ERR_INVALID, "evk_optics_points: D %d, min_pts %d, eps %g", D, min_pts, eps);
    if (n > kMaxPoints)
        return evk_fail(h, EVK_ERR_CAPACITY, "OPTICS clouds hold at most %zu points", kMaxPoints);
    DeviceGuard g(h->device);
    std::vector<int32_t> host(pts, pts + n * D);
    int32_t* d_pts = nullptr;
    if (n) {
        EVK_CUDA(h, cudaMalloc((void**)&d_pts, n * D * sizeof(int32_t)));
        const cudaError_t ce = cudaMemcpyAsync(d_pts, pts, n * D * sizeof(int32_t),
                                               cudaMemcpyHostToDevice, h->stream);
        if (ce != cudaSuccess) {
            cudaFree(d_pts);
            return evk_fail(h, EVK_ERR_CUDA, "%s", cudaGetErrorString(ce));
        }
    }
    const int st = optics_run(h, d_pts, host, n, D, min_pts, eps);
    if (d_pts) cudaFree(d_pts);
    return st;
}

int evk_optics_voxels(evk_handle* h, int min_pts, double eps) {
    if (!h) return EVK_ERR_INVALID;
    EVK_TRY(evk_collect_pending(h));
    if (!h->have_voxels) return evk_fail(h, EVK_ERR_STATE, "evk_downsample has not run");
    if (min_pts < 1 || min_pts > kMaxMinPts || !(eps > 0))
        return evk_fail(h, EVK_ERR_INVALID, "evk_optics_voxels: min_pts %d, eps %g", min_pts, eps);
    const size_t n = h->n_unique;
    if (n > kMaxPoints)
        return evk_fail(h, EVK_ERR_CAPACITY, "OPTICS clouds hold at most %zu points (%zu voxels: "
                        "downsample more coarsely)", kMaxPoints, n);
    DeviceGuard g(h->device);
    std::vector<int32_t> host(n * 2);
    int32_t* d_pts = nullptr;
    if (n) {
        EVK_TRY(evk_ensure_perm(h));
        EVK_CUDA(h, cudaMalloc((void**)&d_pts, n * 2 * sizeof(int32_t)));
        k_optics_voxel_points<<<((uint32_t)n + kT - 1) / kT, kT, 0, h->stream>>>(h->d_xy, h->d_perm,
                                                                                  (uint32_t)n, d_pts);
        cudaError_t ce = cudaMemcpyAsync(host.data(), d_pts, n * 2 * sizeof(int32_t),
                                         cudaMemcpyDeviceToHost, h->stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream);
        if (ce != cudaSuccess) {
            cudaFree(d_pts);
            return evk_fail(h, EVK_ERR_CUDA, "%s", cudaGetErrorString(ce));
        }
    }
    const int st = optics_run(h, d_pts, host, n, 2, min_pts, eps);
    if (d_pts) cudaFree(d_pts);
    return st;
}

int evk_optics_get(evk_handle* h, uint32_t* order, double* reach, size_t cap, size_t* n) {
    if (!h) return EVK_ERR_INVALID;
    if (!h->optics || !h->optics->have) return evk_fail(h, EVK_ERR_STATE, "no OPTICS ordering");
    OpticsHost* o = h->optics;
    if (n) *n = o->n;
    if (!order && !reach) return EVK_OK;
    if (cap < o->n) return evk_fail(h, EVK_ERR_CAPACITY, "%zu points, room for %zu", o->n, cap);
    if (order && o->n) memcpy(order, o->order.data(), o->n * sizeof(uint32_t));
    if (reach && o->n) memcpy(reach, o->reach.data(), o->n * sizeof(double));
    return EVK_OK;
}

int evk_optics_clusters(evk_handle* h, double threshold, uint32_t* cluster, size_t cap,
                        size_t* n_clusters) {
    if (!h) return EVK_ERR_INVALID;
    if (!h->optics || !h->optics->have) return evk_fail(h, EVK_ERR_STATE, "no OPTICS ordering");
    OpticsHost* o = h->optics;
    if (cluster && cap < o->n) return evk_fail(h, EVK_ERR_CAPACITY, "%zu points, room for %zu", o->n, cap);
    size_t nc = 0;
    for (size_t k = 0; k < o->n; k++) {
        if (o->reach[k] < 0.0 || o->reach[k] >= threshold || nc == 0) nc++;
        if (cluster) cluster[k] = (uint32_t)(nc - 1);
    }
    if (n_clusters) *n_clusters = nc;
    return EVK_OK;
}

}  // extern "C"
