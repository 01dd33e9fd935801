// evk_dbscan.cu — density clustering of the downsampled cloud on the device.
//
// Reference: event-cam-clustering/point-cloud-clustering/DBSCAN_simple.h (PCC below; the app,
// pcl_cluster.cpp:112-123, instantiates the kd-tree variant DBSCAN_kdtree.h, which replaces only the
// radius search).  The reference walks the points in index order and grows one cluster at a time
// through a seed queue (:27-95).  What that sequential walk computes has an ORDER-FREE statement,
// which is what runs here (the test suite checks this statement, in numpy, against the reference's
// own code compiled in place):
//   core(i)      #{j : |p_j - p_i|^2 <= eps^2} >= minPts, i itself counted (:36-39), double
//                arithmetic on the float coordinates (:117-140)
//   cluster      a connected component of core points under the eps relation; its SEED is its
//                lowest-index core point, and clusters are discovered in seed order
//   border point a non-core point with a core neighbour: member of the cluster with the LOWEST seed
//                among its core neighbours' clusters (the first one to reach it) -- and ALSO of every
//                other cluster whose seed POINT is its neighbour, because the seed's neighbours are
//                queued whatever their state (:45-50); those second memberships are the `extra` pairs
//   kept         clusters with min <= members <= max (:73), largest first (:89; ties by seed)
// Kernels: points are bucketed into a grid of eps-sized cells by one radix sort of 64-bit cell keys
// (CUB, library code, as on the sort cross-check path); a neighbourhood is then 3 (2-D) or 9 (3-D)
// contiguous runs of the sorted array, found by binary search.
//   k_db_keys / k_db_gather   cell key per point; points gathered into sorted order
//   k_db_count                neighbours within eps -> core flags
//   k_db_union                lock-free union-find over core-core edges, smaller index wins the root
//                             (so the root of a component IS its seed)
//   k_db_label                roots flattened; border points take the lowest neighbouring root and
//                             emit their extra pairs; member counts per seed
// Neighbour scans are gathers over an L2-resident point array: gather / atomic-bound, no roofline
// claim is made for this row.
#include <algorithm>
#include <utility>
#include <cub/device/device_radix_sort.cuh>

#include "evk_internal.cuh"

namespace {

constexpr int kT = 256;
constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr long long kBias = 1ll << 20;  // cell coordinates are biased into 21 bits

struct DbArgs {
    const float4* spts;   // points in cell order (x, y, z, -)
    const uint64_t* skey; // their cell keys, ascending
    const uint32_t* sidx; // their original indices
    uint32_t n;
    int rows;             // 3 (z = const) or 9
    double r2, inv_cell;
};

__device__ __forceinline__ uint64_t cell_key(long long cx, long long cy, long long cz) {
    return ((uint64_t)(cz + kBias) << 42) | ((uint64_t)(cy + kBias) << 21) | (uint64_t)(cx + kBias);
}
__device__ __forceinline__ uint32_t lower_bound(const uint64_t* a, uint32_t n, uint64_t v) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(kT)
    k_db_keys(const float4* __restrict__ pts, uint32_t n, double inv_cell, uint64_t* key,
              uint32_t* idx, unsigned int* bad) {
    const uint32_t i = blockIdx.x * kT + threadIdx.x;
    if (i >= n) return;
    const float4 p = pts[i];
    const long long cx = (long long)floor((double)p.x * inv_cell);
    const long long cy = (long long)floor((double)p.y * inv_cell);
    const long long cz = (long long)floor((double)p.z * inv_cell);
    if (cx <= -kBias + 1 || cx >= kBias - 2 || cy <= -kBias + 1 || cy >= kBias - 2 ||
        cz <= -kBias + 1 || cz >= kBias - 2 || !(p.x == p.x) || !(p.y == p.y) || !(p.z == p.z))
        atomicOr(bad, 1u);
    key[i] = cell_key(cx, cy, cz);
    idx[i] = i;
}
__global__ void __launch_bounds__(kT)
    k_db_gather(const float4* __restrict__ pts, const uint32_t* __restrict__ sidx, uint32_t n,
                float4* spts) {
    const uint32_t k = blockIdx.x * kT + threadIdx.x;
    if (k < n) spts[k] = pts[sidx[k]];
}

// calls f(m) for every sorted position m whose point lies within eps of sorted position k
template <typename F>
__device__ __forceinline__ void for_neighbours(const DbArgs& a, uint32_t k, F f) {
    const float4 p = a.spts[k];
    const long long cx = (long long)floor((double)p.x * a.inv_cell);
    const long long cy = (long long)floor((double)p.y * a.inv_cell);
    const long long cz = (long long)floor((double)p.z * a.inv_cell);
    for (int r = 0; r < a.rows; r++) {
        const long long dy = r % 3 - 1, dz = a.rows == 9 ? r / 3 - 1 : 0;
        const uint32_t lo = lower_bound(a.skey, a.n, cell_key(cx - 1, cy + dy, cz + dz));
        const uint64_t last = cell_key(cx + 1, cy + dy, cz + dz);
        for (uint32_t m = lo; m < a.n && a.skey[m] <= last; m++) {
            const float4 q = a.spts[m];
            const double dx = (double)q.x - (double)p.x, dyy = (double)q.y - (double)p.y,
                         dzz = (double)q.z - (double)p.z;
            // (PCC:131: x*x + y*y + z*z, left to right, no contraction)
            const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dyy, dyy)),
                                        __dmul_rn(dzz, dzz));
            if (d2 <= a.r2) f(m);
        }
    }
}

__global__ void __launch_bounds__(kT) k_db_count(DbArgs a, int min_pts, uint8_t* score) {
    const uint32_t k = blockIdx.x * kT + threadIdx.x;
    if (k >= a.n) return;
    int cnt = 0;
    for_neighbours(a, k, [&](uint32_t) { cnt++; });
    score[k] = cnt >= min_pts ? 1 : 0;
}

__device__ __forceinline__ uint32_t uf_find(uint32_t* parent, uint32_t x) {
    for (;;) {
        const uint32_t p = parent[x];
        if (p == x) return x;
        const uint32_t g = parent[p];
        if (g != p) parent[x] = g;  // path halving (a benign race: any ancestor is a valid parent)
        x = p;
    }
}
__device__ __forceinline__ void uf_union(uint32_t* parent, uint32_t a, uint32_t b) {
    for (;;) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        const uint32_t hi = a > b ? a : b, lo = a > b ? b : a;
        if (atomicCAS(&parent[hi], hi, lo) == hi) return;  // the smaller index stays the root
    }
}

__global__ void __launch_bounds__(kT)
    k_db_union(DbArgs a, const uint8_t* __restrict__ score, uint32_t* parent) {
    const uint32_t k = blockIdx.x * kT + threadIdx.x;
    if (k >= a.n || !score[k]) return;
    const uint32_t i = a.sidx[k];
    for_neighbours(a, k, [&](uint32_t m) {
        if (score[m]) {
            const uint32_t j = a.sidx[m];
            if (j < i) uf_union(parent, i, j);
        }
    });
}

// label[i] = seed of the cluster that holds point i first (kNone: noise); size[seed] += 1 per member
__global__ void __launch_bounds__(kT)
    k_db_label(DbArgs a, const uint8_t* __restrict__ score, uint32_t* parent, uint32_t* label,
               uint32_t* size, uint32_t* extra, uint32_t cap_extra, unsigned int* n_extra) {
    const uint32_t k = blockIdx.x * kT + threadIdx.x;
    if (k >= a.n) return;
    const uint32_t i = a.sidx[k];
    if (score[k]) {
        const uint32_t r = uf_find(parent, i);
        label[i] = r;
        atomicAdd(&size[r], 1u);
        return;
    }
    uint32_t prim = kNone;
    for_neighbours(a, k, [&](uint32_t m) {
        if (score[m]) prim = min(prim, uf_find(parent, a.sidx[m]));
    });
    label[i] = prim;
    if (prim == kNone) return;
    atomicAdd(&size[prim], 1u);
    for_neighbours(a, k, [&](uint32_t m) {  // seed points of OTHER clusters next to this border point
        if (!score[m]) return;
        const uint32_t j = a.sidx[m];
        if (j != prim && parent[j] == j) {
            const uint32_t o = atomicAdd(n_extra, 1u);
            if (o < cap_extra) {
                extra[2 * o] = i;
                extra[2 * o + 1] = j;
            }
            atomicAdd(&size[j], 1u);
        }
    });
}

__global__ void __launch_bounds__(kT) k_db_iota(uint32_t* p, uint32_t n) {
    const uint32_t i = blockIdx.x * kT + threadIdx.x;
    if (i < n) p[i] = i;
}

// voxel representatives in canonical order -> points (x, y, z): z = 0 (D = 2) or (t - t0) * t_scale
__global__ void __launch_bounds__(kT)
    k_db_voxel_points(const uint32_t* __restrict__ xy, const uint32_t* __restrict__ perm,
                      const uint32_t* __restrict__ first, const evk_event* __restrict__ ev,
                      uint32_t first_offset, uint32_t n, int D, double t_scale, long long t0,
                      float4* pts) {
    const uint32_t i = blockIdx.x * kT + threadIdx.x;
    if (i >= n) return;
    const uint32_t e = perm[i], v = xy[e];
    float z = 0.f;
    if (D == 3) {
        const uint4 r = ld_event(ev + (first[e] - first_offset));
        z = (float)((double)(ev_t(r) - t0) * t_scale);
    }
    pts[i] = make_float4((float)(v & 0xFFFFu), (float)(v >> 16), z, 0.f);
}

}  // namespace

struct DbHost {
    size_t cap = 0;  // points the device buffers hold
    float4 *d_pts = nullptr, *d_spts = nullptr;
    uint64_t *d_key = nullptr, *d_skey = nullptr;
    uint32_t *d_idx = nullptr, *d_sidx = nullptr, *d_parent = nullptr, *d_label = nullptr;
    uint32_t *d_size = nullptr, *d_extra = nullptr;
    uint8_t* d_score = nullptr;
    unsigned int* d_flags = nullptr;  // [0] bad coordinates, [1] extra pairs
    void* d_tmp = nullptr;
    size_t tmp_bytes = 0;
    // results of the last run (host)
    std::vector<int32_t> labels;
    std::vector<uint32_t> sizes, seeds, extra;  // extra: pairs (point, cluster rank)
    bool have = false;
};

static void db_free(DbHost* d) {
    void* ptrs[] = {d->d_pts,  d->d_spts,  d->d_key,  d->d_skey,  d->d_idx,   d->d_sidx, d->d_parent,
                    d->d_label, d->d_size, d->d_extra, d->d_score, d->d_flags, d->d_tmp};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    d->d_pts = d->d_spts = nullptr;
    d->d_key = d->d_skey = nullptr;
    d->d_idx = d->d_sidx = d->d_parent = d->d_label = d->d_size = d->d_extra = nullptr;
    d->d_score = nullptr;
    d->d_flags = nullptr;
    d->d_tmp = nullptr;
    d->cap = d->tmp_bytes = 0;
}

static int db_reserve(evk_handle* h, size_t n) {
    if (!h->db) h->db = new DbHost;
    DbHost* d = h->db;
    if (n <= d->cap) return EVK_OK;
    db_free(d);
    const size_t c = n < 4096 ? 4096 : n;
    cudaError_t e = cudaMalloc((void**)&d->d_pts, c * sizeof(float4));
    if (e == cudaSuccess) e = cudaMalloc((void**)&d->d_spts, c * sizeof(float4));
    if (e == cudaSuccess) e = cudaMalloc((void**)&d->d_key, c * 8);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d->d_skey, c * 8);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d->d_idx, c * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d->d_sidx, c * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d->d_parent, c * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d->d_label, c * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d->d_size, c * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d->d_extra, c * 8);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d->d_score, c);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d->d_flags, 2 * sizeof(unsigned int));
    size_t bytes = 0;
    if (e == cudaSuccess)
        e = cub::DeviceRadixSort::SortPairs(nullptr, bytes, d->d_key, d->d_skey, d->d_idx, d->d_sidx,
                                            (int64_t)c, 0, 63, h->stream);
    if (e == cudaSuccess) e = cudaMalloc(&d->d_tmp, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        db_free(d);
        return evk_fail(h, EVK_ERR_NOMEM, "dbscan buffers for %zu points: %s", n, cudaGetErrorString(e));
    }
    d->tmp_bytes = bytes;
    d->cap = c;
    return EVK_OK;
}

// the points are in d->d_pts[0..n)
static int db_run(evk_handle* h, size_t n, const evk_dbscan_params* p, bool flat, size_t* n_clusters,
                  size_t* n_extra) {
    DbHost* d = h->db;
    d->have = false;
    d->labels.assign(n, -1);
    d->sizes.clear();
    d->seeds.clear();
    d->extra.clear();
    if (n_clusters) *n_clusters = 0;
    if (n_extra) *n_extra = 0;
    if (n == 0) {
        d->have = true;
        return EVK_OK;
    }
    const unsigned nb = (unsigned)((n + kT - 1) / kT);
    const double inv_cell = 1.0 / (p->eps * 1.000001);  // cells a hair wider than eps: a neighbour
                                                        // is never more than one cell away
    EVK_CUDA(h, cudaMemsetAsync(d->d_flags, 0, 2 * sizeof(unsigned int), h->stream));
    EVK_CUDA(h, cudaMemsetAsync(d->d_size, 0, n * 4, h->stream));
    k_db_keys<<<nb, kT, 0, h->stream>>>(d->d_pts, (uint32_t)n, inv_cell, d->d_key, d->d_idx, d->d_flags);
    size_t bytes = d->tmp_bytes;
    EVK_CUDA(h, cub::DeviceRadixSort::SortPairs(d->d_tmp, bytes, d->d_key, d->d_skey, d->d_idx,
                                                d->d_sidx, (int64_t)n, 0, 63, h->stream));
    k_db_gather<<<nb, kT, 0, h->stream>>>(d->d_pts, d->d_sidx, (uint32_t)n, d->d_spts);
    DbArgs a;
    a.spts = d->d_spts;
    a.skey = d->d_skey;
    a.sidx = d->d_sidx;
    a.n = (uint32_t)n;
    a.rows = flat ? 3 : 9;
    a.r2 = p->eps * p->eps;
    a.inv_cell = inv_cell;
    k_db_count<<<nb, kT, 0, h->stream>>>(a, p->min_pts, d->d_score);
    k_db_iota<<<nb, kT, 0, h->stream>>>(d->d_parent, (uint32_t)n);
    k_db_union<<<nb, kT, 0, h->stream>>>(a, d->d_score, d->d_parent);
    k_db_label<<<nb, kT, 0, h->stream>>>(a, d->d_score, d->d_parent, d->d_label, d->d_size, d->d_extra,
                                         (uint32_t)n, d->d_flags + 1);
    EVK_CUDA(h, cudaGetLastError());
    // the cluster table is small bookkeeping: built on the host from the seeds' member counts
    std::vector<uint32_t> label(n), size(n);
    unsigned int flags[2];
    EVK_CUDA(h, cudaMemcpyAsync(label.data(), d->d_label, n * 4, cudaMemcpyDeviceToHost, h->stream));
    EVK_CUDA(h, cudaMemcpyAsync(size.data(), d->d_size, n * 4, cudaMemcpyDeviceToHost, h->stream));
    EVK_CUDA(h, cudaMemcpyAsync(flags, d->d_flags, sizeof flags, cudaMemcpyDeviceToHost, h->stream));
    EVK_CUDA(h, cudaStreamSynchronize(h->stream));
    if (flags[0]) return evk_fail(h, EVK_ERR_INVALID, "dbscan: a coordinate is NaN or beyond 2^20 cells");
    if (flags[1] > n) return evk_fail(h, EVK_ERR_CAPACITY, "dbscan: %u second memberships", flags[1]);
    std::vector<uint32_t> ex(2 * (size_t)flags[1]);
    if (flags[1])
        EVK_CUDA(h, cudaMemcpy(ex.data(), d->d_extra, ex.size() * 4, cudaMemcpyDeviceToHost));
    struct Cl {
        uint32_t size, seed;
    };
    std::vector<Cl> cl;
    for (size_t i = 0; i < n; i++)
        if (size[i] && (long long)size[i] >= p->min_cluster && (long long)size[i] <= p->max_cluster)
            cl.push_back({size[i], (uint32_t)i});
    std::sort(cl.begin(), cl.end(), [](const Cl& x, const Cl& y) {
        return x.size != y.size ? x.size > y.size : x.seed < y.seed;
    });
    std::vector<int32_t> rank(n, -1);  // seed -> position in the output order
    for (size_t c = 0; c < cl.size(); c++) {
        rank[cl[c].seed] = (int32_t)c;
        d->sizes.push_back(cl[c].size);
        d->seeds.push_back(cl[c].seed);
    }
    for (size_t i = 0; i < n; i++) d->labels[i] = label[i] == kNone ? -1 : rank[label[i]];
    // second memberships in kept clusters, by (point, seed).  A point whose first cluster was dropped
    // by the size filter is labelled with its next kept one (lowest seed) instead of staying -1.
    std::vector<std::pair<uint32_t, uint32_t>> pairs;
    for (size_t q = 0; q + 1 < ex.size(); q += 2)
        if (rank[ex[q + 1]] >= 0) pairs.emplace_back(ex[q], ex[q + 1]);
    std::sort(pairs.begin(), pairs.end());
    for (const auto& pr : pairs) {
        if (d->labels[pr.first] < 0) {
            d->labels[pr.first] = rank[pr.second];
        } else {
            d->extra.push_back(pr.first);
            d->extra.push_back((uint32_t)rank[pr.second]);
        }
    }
    d->have = true;
    if (n_clusters) *n_clusters = cl.size();
    if (n_extra) *n_extra = d->extra.size() / 2;
    return EVK_OK;
}

static int db_check(evk_handle* h, const evk_dbscan_params* p) {
    if (!h) return EVK_ERR_INVALID;
    if (!p || !(p->eps > 0) || p->min_pts < 1 || p->min_cluster < 0 || p->max_cluster < p->min_cluster)
        return evk_fail(h, EVK_ERR_INVALID, "bad dbscan parameters");
    return EVK_OK;
}

extern "C" {

int evk_dbscan_destroy(evk_handle* h) {
    if (!h || !h->db) return EVK_OK;
    if (h->stream) cudaStreamSynchronize(h->stream);
    db_free(h->db);
    delete h->db;
    h->db = nullptr;
    return EVK_OK;
}

int evk_dbscan_points(evk_handle* h, const float* xyz, size_t n, const evk_dbscan_params* p,
                      size_t* n_clusters, size_t* n_extra) {
    EVK_TRY(db_check(h, p));
    if (n && !xyz) return evk_fail(h, EVK_ERR_INVALID, "evk_dbscan_points: null points");
    if (n >= 0xFFFFFFF0ull) return evk_fail(h, EVK_ERR_CAPACITY, "dbscan: too many points");
    cudaSetDevice(h->device);
    EVK_TRY(db_reserve(h, n));
    bool flat = true;
    if (n) {
        std::vector<float4> tmp(n);
        for (size_t i = 0; i < n; i++) {
            tmp[i] = make_float4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0.f);
            flat = flat && xyz[3 * i + 2] == xyz[2];
        }
        EVK_CUDA(h, cudaMemcpy(h->db->d_pts, tmp.data(), n * sizeof(float4), cudaMemcpyHostToDevice));
    }
    return db_run(h, n, p, flat, n_clusters, n_extra);
}

int evk_dbscan_voxels(evk_handle* h, const evk_dbscan_params* p, size_t* n_clusters,
                      size_t* n_extra) {
    EVK_TRY(db_check(h, p));
    EVK_TRY(evk_collect_pending(h));
    if (!h->have_voxels) return evk_fail(h, EVK_ERR_STATE, "evk_dbscan_voxels: no voxel shard");
    if (p->D != 2 && p->D != 3) return evk_fail(h, EVK_ERR_INVALID, "dbscan: D must be 2 or 3");
    if (p->D == 3 && (h->comm || h->reps_valid || h->voxels_foreign))
        return evk_fail(h, EVK_ERR_INVALID, "dbscan: D = 3 needs the shard's own events (not sharded)");
    cudaSetDevice(h->device);
    const size_t n = h->n_unique;
    EVK_TRY(db_reserve(h, n));
    if (n) {
        EVK_TRY(evk_ensure_perm(h));
        k_db_voxel_points<<<(unsigned)((n + kT - 1) / kT), kT, 0, h->stream>>>(
            h->d_xy, h->d_perm, h->d_first, h->d_events, (uint32_t)h->shard_first, (uint32_t)n, p->D,
            p->t_scale, (long long)p->t0_us, h->db->d_pts);
        EVK_CUDA(h, cudaGetLastError());
    }
    return db_run(h, n, p, p->D == 2, n_clusters, n_extra);
}

int evk_dbscan_get(evk_handle* h, int32_t* labels, size_t cap_points, uint32_t* sizes,
                   uint32_t* seeds, size_t cap_clusters, uint32_t* extra_pairs, size_t cap_extra) {
    if (!h) return EVK_ERR_INVALID;
    if (!h->db || !h->db->have) return evk_fail(h, EVK_ERR_STATE, "evk_dbscan_* has not run");
    DbHost* d = h->db;
    if ((labels && cap_points < d->labels.size()) ||
        ((sizes || seeds) && cap_clusters < d->sizes.size()) ||
        (extra_pairs && cap_extra < d->extra.size() / 2))
        return evk_fail(h, EVK_ERR_CAPACITY, "evk_dbscan_get: %zu points, %zu clusters, %zu extra pairs",
                        d->labels.size(), d->sizes.size(), d->extra.size() / 2);
    if (labels && !d->labels.empty()) memcpy(labels, d->labels.data(), d->labels.size() * 4);
    if (sizes && !d->sizes.empty()) memcpy(sizes, d->sizes.data(), d->sizes.size() * 4);
    if (seeds && !d->seeds.empty()) memcpy(seeds, d->seeds.data(), d->seeds.size() * 4);
    if (extra_pairs && !d->extra.empty()) memcpy(extra_pairs, d->extra.data(), d->extra.size() * 4);
    return EVK_OK;
}

}  // extern "C"
