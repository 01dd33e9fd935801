// evk_aec.cu — the asynchronous event clustering consumer of the downsampler's output on the device
// (SURVEY.md 8f rank 1).
//
// Reference (paths relative to event-cam-clustering-accel/event-cam-clustering-downsampling-accel/):
// every unique coordinate the slice callback keeps is pushed through AEClustering::update
// (metavision_sdk_get_started5_opencl_store.cpp:435-445; AEClustering.cpp:48-123): the event joins
// the first cluster whose moving average (MyCluster.cpp:200-202) lies within a Manhattan radius
// (or, for clusters above minN events, within the radius of kappa randomly sampled stored events,
// MyCluster.cpp:72-103), clusters it bridges are merged (AEClustering.cpp:148-211), events older
// than the szBuffer-th last one are forgotten (MyCluster.cpp:54-65, AEClustering.cpp:137-146), and
// after the slice every cluster with >= minN events reports its centroid and the displacement from
// its previous report -- the optical-flow arrow (store.cpp:461-521).
//
// The algorithm is sequential BY DEFINITION: event i sees the clusters exactly as event i-1 left
// them, and results must equal the reference's state for state (same cluster order, ids, moving
// averages to the last bit, same stored events).  So this is a latency kernel, not a bandwidth
// kernel: ONE warp walks the events in order and the parallelism is ACROSS CLUSTERS and across
// stored events inside each step --
//   * cluster table (n, ring head, storage slot, id, mu, time of the oldest stored event) lives in
//     shared memory for the whole call, one lane per cluster: forgetting, the distance test and
//     the three-way classification (empty / near / needs sampling) are one pass of ballots;
//   * the time buffer (last szBuffer event times) is a shared-memory ring;
//   * stored events live in HBM/L2 as one fixed-capacity ring per cluster slot (SoA), touched only
//     by the append (fire and forget), by a pop (the new oldest time) and by sampling / merging;
//   * a merge is a parallel rank computation: the reference's "repeatedly take the list whose
//     head is oldest" equals a stable merge on the running maximum of each list's times, so every
//     stored event finds its output position with binary searches instead of a sequential walk;
//   * std::rand() (MyCluster.cpp:88) is the glibc TYPE_3 additive generator, kept in shared memory
//     with 31 values generated per step (the recurrence unrolled eleven times), consumed in the
//     reference's draw order (clusters in list order, kappa each).
// All arithmetic is IEEE double with one rounding per operation (__dmul_rn / __dadd_rn: no FMA
// contraction), as the reference's x86-64 build evaluates it.
//
// Nothing leaves the device between the downsample and the report: evk_aec_update_voxels feeds the
// representatives of the current voxel shard (canonical order) straight into the update kernel.
#include <float.h>

#include "evk_internal.cuh"

namespace {

constexpr int kMaxC = 1024;      // simultaneous clusters (one lane each, 32 per pass)
constexpr int kMaxKappa = 256;   // draws per sampled cluster
constexpr int kDrawCap = 2048;   // draws generated per batch of sampled clusters
constexpr int kMaxIds = 16384;   // centroid_prev[16384][2], store.cpp:188
constexpr unsigned kFull = 0xffffffffu;

enum { AEC_OK = 0, AEC_ERR_CLUSTERS = 1, AEC_ERR_POINTS = 2, AEC_ERR_IDS = 3 };

struct AecDev {  // device-resident header; the arrays follow it in the same allocation
    // parameters
    int sz_buffer, kappa, min_n, max_clusters, cap;
    double radius, alpha;
    // scalar state of AEClustering
    double t0;
    int event_id, next_id, last, nc;
    int tb_head, tb_n, n_free;
    int error;
    long long error_event, events_done;
    long long rcount;  // glibc random() TYPE_3: index of the next value of r[i] = r[i-31] + r[i-3]
    // arrays
    double* tbuf;                                // [sz_buffer + 1] ring
    int *c_slot, *c_id, *c_n, *c_head;           // [kMaxC] by position in the cluster list
    double *c_mux, *c_muy;                       // [kMaxC]
    int* free_slots;                             // [max_clusters] stack
    double *p_x, *p_y, *p_t;                     // [max_clusters * cap] rings by slot
    int* p_id;
    unsigned char* p_pol;
    double *m_x, *m_y, *m_t, *m_kp;              // merge scratch [cap]
    int* m_id;
    unsigned char* m_pol;
    double* prev;                                // [kMaxIds][2]
    uint32_t* rhist;                             // [128] the last 128 values of r[], ring by i & 127
    int n_report;                                // records written by the last report
};

struct AecSmem {  // carved out of dynamic shared memory
    double *mux, *muy, *ft, *tb;
    int *n, *head, *slot, *id, *asg, *rem, *draw;
    uint32_t *hist, *hit;  // generator history ring [128]; sampling hits by 32 positions [32]
};
__device__ __forceinline__ AecSmem carve(unsigned char* raw, int ring) {
    AecSmem s;
    double* d = reinterpret_cast<double*>(raw);
    s.mux = d;
    s.muy = d + kMaxC;
    s.ft = d + 2 * kMaxC;
    s.tb = d + 3 * kMaxC;
    int* i = reinterpret_cast<int*>(s.tb + ring);
    s.n = i;
    s.head = i + kMaxC;
    s.slot = i + 2 * kMaxC;
    s.id = i + 3 * kMaxC;
    s.asg = i + 4 * kMaxC;
    s.rem = i + 5 * kMaxC;
    s.draw = i + 6 * kMaxC;
    s.hist = reinterpret_cast<uint32_t*>(s.draw + kDrawCap);
    s.hit = s.hist + 128;
    return s;
}
size_t aec_smem_bytes(int ring) {
    return sizeof(double) * (3 * kMaxC + (size_t)ring) + sizeof(int) * (6 * kMaxC + kDrawCap + 128 + 32);
}

__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(kFull, v, o));
    return v;
}
__device__ __forceinline__ double manhattan(double x, double y, double mx, double my) {
    return __dadd_rn(fabs(__dsub_rn(x, mx)), fabs(__dsub_rn(y, my)));
}

// exclusive prefix sum of v over the lanes; *total = sum over the warp
__device__ __forceinline__ int warp_excl_scan(int v, int lane, int* total) {
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int u = __shfl_up_sync(kFull, inc, d);
        if (lane >= d) inc += u;
    }
    *total = __shfl_sync(kFull, inc, 31);
    return inc - v;
}

// Remove the clusters whose bits are set in `gone` (lane c holds the mask of positions
// 32c .. 32c+31) from the list, keeping the order of the others (std::deque::erase,
// AEClustering.cpp:116-121,208-210); their storage slots go back to the free stack.
// Warp-synchronous; returns the new cluster count.
__device__ __forceinline__ int aec_erase(const AecSmem& s, const AecDev* S, unsigned gone, int nc,
                                         int lane, int* n_free) {
    int k;
    const int before = warp_excl_scan(__popc(gone), lane, &k);  // removed positions in earlier chunks
    const unsigned lt = (1u << lane) - 1u;
    const unsigned chunks = __ballot_sync(kFull, gone != 0);
    const int first_ch = __ffs(chunks) - 1;  // positions before the first removal stay put
    for (int ch = first_ch, base = first_ch << 5; base < nc; ch++, base += 32) {
        const int i = base + lane;
        const unsigned gm = __shfl_sync(kFull, gone, ch);
        const int below = __shfl_sync(kFull, before, ch) + __popc(gm & lt);
        const bool g = (gm >> lane) & 1u;
        double mx = 0, my = 0, ft = 0;
        int cn = 0, hd = 0, sl = 0, id = 0;
        if (i < nc) {
            mx = s.mux[i];
            my = s.muy[i];
            ft = s.ft[i];
            cn = s.n[i];
            hd = s.head[i];
            sl = s.slot[i];
            id = s.id[i];
        }
        if (g) S->free_slots[*n_free + __popc(gm & lt)] = sl;
        *n_free += __popc(gm);
        __syncwarp();
        if (i < nc && !g && below) {
            const int j = i - below;
            s.mux[j] = mx;
            s.muy[j] = my;
            s.ft[j] = ft;
            s.n[j] = cn;
            s.head[j] = hd;
            s.slot[j] = sl;
            s.id[j] = id;
        }
        __syncwarp();
    }
    return nc - k;
}

// positions of the set bits of `mask` (lane c: positions 32c..) in ascending order -> lst[]
__device__ __forceinline__ void aec_list(unsigned mask, int lane, int* lst) {
    int tot;
    int o = warp_excl_scan(__popc(mask), lane, &tot);
    while (mask) {
        lst[o++] = (lane << 5) + __ffs(mask) - 1;
        mask &= mask - 1;
    }
    __syncwarp();
}

// AEClustering::merge_clusters_ (AEClustering.cpp:148-211) for the clusters at positions
// s.asg[0..m).  Returns false when the merged cluster does not fit a storage ring.
__device__ __forceinline__ bool aec_merge(const AecSmem& s, const AecDev* S, int m, int lane) {
    const int cap = S->cap;
    int aux_n = 0;
    for (int ii = 0; ii < m; ii++) aux_n += s.n[s.asg[ii]];
    if (aux_n > cap) return false;
    double aux0 = 0.0, aux1 = 0.0;  // :170-173, coefficient-wise, list order
    for (int ii = 0; ii < m; ii++) {
        const int c = s.asg[ii];
        const double w = (double)s.n[c] / (double)aux_n;
        aux0 = __dadd_rn(aux0, __dmul_rn(w, s.mux[c]));
        aux1 = __dadd_rn(aux1, __dmul_rn(w, s.muy[c]));
    }
    // offsets of the lists in the scratch arrays
    if (lane == 0) {
        int off = 0;
        for (int ii = 0; ii < m; ii++) {
            s.rem[ii] = off;
            off += s.n[s.asg[ii]];
        }
    }
    __syncwarp();
    // 1. copy every list to scratch, with the running maximum of its times as merge key
    for (int ii = 0; ii < m; ii++) {
        const int c = s.asg[ii], cn = s.n[c], hd = s.head[c], off = s.rem[ii];
        const size_t o = (size_t)s.slot[c] * cap;
        double carry = -DBL_MAX;
        for (int q0 = 0; q0 < cn; q0 += 32) {
            const int q = q0 + lane;
            double v = -DBL_MAX;
            if (q < cn) {
                int idx = hd + q;
                if (idx >= cap) idx -= cap;
                v = S->p_t[o + idx];
                S->m_x[off + q] = S->p_x[o + idx];
                S->m_y[off + q] = S->p_y[o + idx];
                S->m_t[off + q] = v;
                S->m_id[off + q] = S->p_id[o + idx];
                S->m_pol[off + q] = S->p_pol[o + idx];
            }
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double u = __shfl_up_sync(kFull, v, d);
                if (lane >= d) v = fmax(v, u);
            }
            v = fmax(v, carry);
            if (q < cn) S->m_kp[off + q] = v;
            carry = __shfl_sync(kFull, v, 31);
        }
    }
    __syncwarp();
    // 2. every stored event computes its position in the merged list: its own position in its list
    //    plus, for every other list, the events that leave before it (ties: the lower list first)
    const int c0 = s.asg[0];
    const size_t o0 = (size_t)s.slot[c0] * cap;
    for (int g0 = 0; g0 < aux_n; g0 += 32) {
        const int g = g0 + lane;
        if (g < aux_n) {
            int k = 0;
            while (k + 1 < m && s.rem[k + 1] <= g) k++;
            const double key = S->m_kp[g];
            int rank = g - s.rem[k];
            for (int j = 0; j < m; j++) {
                if (j == k) continue;
                const int off = s.rem[j], nj = s.n[s.asg[j]];
                int lo = 0, hi = nj;
                if (j < k) {  // events of an earlier list with key <= mine
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (S->m_kp[off + mid] <= key) lo = mid + 1;
                        else hi = mid;
                    }
                } else {  // events of a later list with key < mine
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (S->m_kp[off + mid] < key) lo = mid + 1;
                        else hi = mid;
                    }
                }
                rank += lo;
            }
            const double tt = S->m_t[g];
            S->p_x[o0 + rank] = S->m_x[g];
            S->p_y[o0 + rank] = S->m_y[g];
            S->p_t[o0 + rank] = tt;
            S->p_id[o0 + rank] = S->m_id[g];
            S->p_pol[o0 + rank] = S->m_pol[g];
            if (rank == 0) s.ft[c0] = tt;
        }
    }
    __syncwarp();
    if (lane == 0) {
        s.head[c0] = 0;
        s.n[c0] = aux_n;
        s.mux[c0] = aux0;
        s.muy[c0] = aux1;
    }
    __syncwarp();
    return true;
}

// glibc random_r, TYPE_3: r[i] = r[i-31] + r[i-3].  Substituting the second term eleven times gives
// r[i] = r[i-33] + sum_{j=0..10} r[i-31-3j]: every index is at least 31 back, so 31 consecutive
// outputs are independent of each other -- one per lane from a 128-entry history ring.
// Writes `count` draws (rand() values) to out[]; *rcount = index of the next value to generate.
__device__ __forceinline__ void aec_rand_fill(uint32_t* hist, long long* rcount, int count, int* out,
                                              int lane) {
    long long K = *rcount;
    for (int done = 0; done < count; done += 31, K += 31) {
        uint32_t v = 0;
        const bool act = lane < 31 && done + lane < count;
        if (act) {
            const long long i = K + lane;
            v = hist[(i - 33) & 127];
#pragma unroll
            for (int j = 0; j <= 10; j++) v += hist[(i - 31 - 3 * j) & 127];
        }
        __syncwarp();
        if (act) {
            hist[(K + lane) & 127] = v;
            out[done + lane] = (int)(v >> 1);
        }
        __syncwarp();
    }
    *rcount += count;
}

// n calls of AEClustering::update, in order.  ev = n x {t, x, y, p} doubles on the device.
__global__ void __launch_bounds__(32) k_aec_update(AecDev* S_global, const double* __restrict__ ev,
                                                   long long n) {
    extern __shared__ __align__(16) unsigned char aec_raw[];
    const int lane = threadIdx.x;
    // the header (parameters, array pointers) is read ONCE into registers: through the pointer every
    // use would be a dependent global load in front of the access it addresses
    AecDev* const Sg = S_global;
    const AecDev hdr = *Sg;
    const AecDev* const S = &hdr;
    const int ring = S->sz_buffer + 1;
    const AecSmem s = carve(aec_raw, ring);
    if (S->error) return;
    // ---- state in ----
    int nc = S->nc, n_free = S->n_free, tb_head = S->tb_head, tb_n = S->tb_n;
    int event_id = S->event_id, next_id = S->next_id, last = S->last;
    double t0 = S->t0;
    long long rcount = S->rcount;
    const int sz = S->sz_buffer, kappa = S->kappa, min_n = S->min_n, cap = S->cap;
    const int max_c = S->max_clusters;
    const double radius = S->radius, alpha = S->alpha, om_alpha = 1 - S->alpha;
    for (int i = lane; i < nc; i += 32) {
        s.mux[i] = S->c_mux[i];
        s.muy[i] = S->c_muy[i];
        s.n[i] = S->c_n[i];
        s.head[i] = S->c_head[i];
        s.slot[i] = S->c_slot[i];
        s.id[i] = S->c_id[i];
        s.ft[i] = S->c_n[i] > 0 ? S->p_t[(size_t)S->c_slot[i] * cap + S->c_head[i]] : 0.0;
    }
    for (int i = lane; i < ring; i += 32) s.tb[i] = S->tbuf[i];
    for (int i = lane; i < 128; i += 32) s.hist[i] = S->rhist[i];
    s.hit[lane] = 0;
    __syncwarp();
    const unsigned lt = (1u << lane) - 1u;
    int err = AEC_OK;
    long long e = 0;
    // the next event is fetched while the current one is processed (a dependent L2 round trip per
    // event otherwise)
    const double2* ev2 = reinterpret_cast<const double2*>(ev);
    double2 nx_a = make_double2(0, 0), nx_b = make_double2(0, 0);
    if (n > 0) {
        nx_a = ev2[0];
        nx_b = ev2[1];
    }
    for (; e < n; e++) {
        const double et = nx_a.x, x = nx_a.y, y = nx_b.x, pv = nx_b.y;
        if (e + 1 < n) {
            nx_a = ev2[2 * (e + 1)];
            nx_b = ev2[2 * (e + 1) + 1];
        }
        if (t0 < 0) t0 = et;  // AEClustering.cpp:49-51
        const double t = __dsub_rn(et, t0);
        // updateBuffer_, AEClustering.cpp:137-146
        {
            int w = tb_head + tb_n;
            if (w >= ring) w -= ring;
            if (lane == 0) s.tb[w] = t;
            tb_n++;
            if (tb_n > sz) {
                tb_head = tb_head + 1 == ring ? 0 : tb_head + 1;
                tb_n--;
            }
        }
        __syncwarp();
        const double tmin = s.tb[tb_head];
        // ---- proximity pass over the cluster list, AEClustering.cpp:69-94: lane c keeps the three
        //      masks of positions 32c .. 32c+31 (empty / near the moving average / to be sampled)
        unsigned my_rem = 0, my_near = 0, my_samp = 0;
        for (int ch = 0, base = 0; base < nc; ch++, base += 32) {
            const int i = base + lane;
            bool rem = false, near = false, samp = false;
            if (i < nc) {
                int cn = s.n[i];
                if (cn > 0 && s.ft[i] < tmin) {  // MyCluster::forget
                    int hd = s.head[i];
                    const size_t o = (size_t)s.slot[i] * cap;
                    double ft = 0.0;
                    do {
                        hd = hd + 1 == cap ? 0 : hd + 1;
                        cn--;
                        if (cn > 0) ft = S->p_t[o + hd];
                    } while (cn > 0 && ft < tmin);
                    s.head[i] = hd;
                    s.n[i] = cn;
                    if (cn > 0) s.ft[i] = ft;
                }
                rem = cn == 0;
                near = !rem && manhattan(x, y, s.mux[i], s.muy[i]) <= radius;
                samp = !rem && !near && cn > min_n;
            }
            const unsigned b_rem = __ballot_sync(kFull, rem);
            const unsigned b_near = __ballot_sync(kFull, near);
            const unsigned b_samp = __ballot_sync(kFull, samp);
            if (lane == ch) {
                my_rem = b_rem;
                my_near = b_near;
                my_samp = b_samp;
            }
        }
        // ---- manhattanDistanceWithSampling (MyCluster.cpp:72-103) for the flagged clusters, in
        //      list order: one lane per cluster, the draws of a batch generated up front
        if (kappa > 0 && __any_sync(kFull, my_samp != 0)) {
            __syncwarp();
            const int F = __reduce_add_sync(kFull, __popc(my_samp));
            aec_list(my_samp, lane, s.asg);  // s.asg[f] = position of the f-th flagged cluster
            for (int f0 = 0; f0 < F;) {
                // batch: as many clusters as fit the draw buffer (those with kappa > n draw none)
                int f1 = f0, nd = 0;
                while (f1 < F) {
                    const int d = kappa > s.n[s.asg[f1]] ? 0 : kappa;
                    if (nd + d > kDrawCap) break;
                    s.rem[f1] = nd;  // draw offset of cluster f1 inside the batch
                    nd += d;
                    f1++;
                }
                __syncwarp();
                aec_rand_fill(s.hist, &rcount, nd, s.draw, lane);
                for (int fb = f0; fb < f1; fb += 32) {
                    const int f = fb + lane;
                    if (f < f1) {
                        const int j = s.asg[f], cn = s.n[j], hd = s.head[j];
                        const size_t o = (size_t)s.slot[j] * cap;
                        double ma = DBL_MAX;
                        if (kappa > cn) {
                            for (int q = 0; q < cn; q++) {
                                int idx = hd + q;
                                if (idx >= cap) idx -= cap;
                                ma = fmin(ma, manhattan(x, y, S->p_x[o + idx], S->p_y[o + idx]));
                            }
                        } else {
                            const int* dr = s.draw + s.rem[f];
#pragma unroll 5
                            for (int q = 0; q < kappa; q++) {
                                int idx = hd + dr[q] % cn;
                                if (idx >= cap) idx -= cap;
                                ma = fmin(ma, manhattan(x, y, S->p_x[o + idx], S->p_y[o + idx]));
                            }
                        }
                        if (ma <= radius) atomicOr(&s.hit[j >> 5], 1u << (j & 31));
                    }
                }
                __syncwarp();
                f0 = f1;
            }
            my_near |= s.hit[lane];
            s.hit[lane] = 0;
            __syncwarp();
        }
        const int na = __reduce_add_sync(kFull, __popc(my_near));
        const int nr = __reduce_add_sync(kFull, __popc(my_rem));
        // ---- no proximity -> new cluster; else join the first one, AEClustering.cpp:97-110 ----
        if (na == 0) {
            if (nc >= max_c || nc >= kMaxC || n_free == 0) {
                err = AEC_ERR_CLUSTERS;
                break;
            }
            const int slot = S->free_slots[n_free - 1];
            n_free--;
            if (lane == 0) {
                const size_t o = (size_t)slot * cap;
                S->p_x[o] = x;
                S->p_y[o] = y;
                S->p_t[o] = t;
                S->p_id[o] = event_id;
                S->p_pol[o] = pv != 0.0;
                s.slot[nc] = slot;
                s.id[nc] = next_id;
                s.n[nc] = 1;
                s.head[nc] = 0;
                s.mux[nc] = x;
                s.muy[nc] = y;
                s.ft[nc] = t;
            }
            next_id++;
            event_id++;
            last = nc;
            nc++;
            __syncwarp();
        } else {
            const unsigned chs = __ballot_sync(kFull, my_near != 0);
            const int ch0 = __ffs(chs) - 1;
            const unsigned m0 = __shfl_sync(kFull, my_near, ch0);
            const int a0 = (ch0 << 5) + __ffs(m0) - 1;
            const int cn = s.n[a0];
            if (cn >= cap) {
                err = AEC_ERR_POINTS;
                break;
            }
            if (lane == 0) {  // MyCluster::add + updateMu_ (cn > 0 here)
                int idx = s.head[a0] + cn;
                if (idx >= cap) idx -= cap;
                const size_t o = (size_t)s.slot[a0] * cap + idx;
                S->p_x[o] = x;
                S->p_y[o] = y;
                S->p_t[o] = t;
                S->p_id[o] = event_id;
                S->p_pol[o] = pv != 0.0;
                s.mux[a0] = __dadd_rn(__dmul_rn(om_alpha, s.mux[a0]), __dmul_rn(alpha, x));
                s.muy[a0] = __dadd_rn(__dmul_rn(om_alpha, s.muy[a0]), __dmul_rn(alpha, y));
                s.n[a0] = cn + 1;
            }
            event_id++;
            last = a0;
            __syncwarp();
            if (na >= 2) {  // proximity to more than one cluster -> merge, then return (:103-107)
                aec_list(my_near, lane, s.asg);
                if (!aec_merge(s, S, na, lane)) {
                    err = AEC_ERR_POINTS;
                    break;
                }
                if (lane == ch0) my_near &= ~(1u << (a0 & 31));  // every assigned cluster but the first
                nc = aec_erase(s, S, my_near, nc, lane, &n_free);
                continue;
            }
        }
        if (nr) {  // AEClustering.cpp:113-121
            // lastUpdatedCluster_ moves down by the removed positions below it
            const int lch = last >> 5;
            const int below = lane < lch ? __popc(my_rem)
                              : (lane == lch ? __popc(my_rem & ((1u << (last & 31)) - 1u)) : 0);
            last -= __reduce_add_sync(kFull, below);
            nc = aec_erase(s, S, my_rem, nc, lane, &n_free);
        }
    }
    __syncwarp();
    // ---- state out ----
    for (int i = lane; i < nc; i += 32) {
        Sg->c_mux[i] = s.mux[i];
        Sg->c_muy[i] = s.muy[i];
        Sg->c_n[i] = s.n[i];
        Sg->c_head[i] = s.head[i];
        Sg->c_slot[i] = s.slot[i];
        Sg->c_id[i] = s.id[i];
    }
    for (int i = lane; i < ring; i += 32) Sg->tbuf[i] = s.tb[i];
    for (int i = lane; i < 128; i += 32) S->rhist[i] = s.hist[i];
    if (lane == 0) {
        Sg->nc = nc;
        Sg->n_free = n_free;
        Sg->tb_head = tb_head;
        Sg->tb_n = tb_n;
        Sg->event_id = event_id;
        Sg->next_id = next_id;
        Sg->last = last;
        Sg->t0 = t0;
        Sg->rcount = rcount;
        Sg->events_done = hdr.events_done + e;
        if (err) {
            Sg->error = err;
            Sg->error_event = hdr.events_done + e;
        }
    }
}

// getClusterCentroid (MyCluster.cpp:171-186): sequential sums in stored order, one division
__device__ __forceinline__ void aec_centroid(const AecDev* S, int c, double* cx, double* cy) {
    const int cn = S->c_n[c], hd = S->c_head[c], cap = S->cap;
    const size_t o = (size_t)S->c_slot[c] * cap;
    double xa = 0, ya = 0;
    for (int q = 0; q < cn; q++) {
        int idx = hd + q;
        if (idx >= cap) idx -= cap;
        xa = __dadd_rn(xa, S->p_x[o + idx]);
        ya = __dadd_rn(ya, S->p_y[o + idx]);
    }
    *cx = xa / (double)cn;
    *cy = ya / (double)cn;
}

__global__ void k_aec_export(const AecDev* S, evk_aec_cluster* out, int cap_out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= S->nc || c >= cap_out) return;
    evk_aec_cluster r;
    r.id = S->c_id[c];
    r.n = S->c_n[c];
    r.mu[0] = S->c_mux[c];
    r.mu[1] = S->c_muy[c];
    aec_centroid(S, c, &r.centroid[0], &r.centroid[1]);
    out[c] = r;
}

__global__ void k_aec_points(const AecDev* S, int c, int32_t* ids, double* xy, double* t,
                             uint8_t* pol, int cap_out) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= S->nc || q >= S->c_n[c] || q >= cap_out) return;
    int idx = S->c_head[c] + q;
    if (idx >= S->cap) idx -= S->cap;
    const size_t o = (size_t)S->c_slot[c] * S->cap + idx;
    ids[q] = S->p_id[o];
    xy[2 * q] = S->p_x[o];
    xy[2 * q + 1] = S->p_y[o];
    t[q] = S->p_t[o];
    pol[q] = S->p_pol[o];
}

// The per-slice report, store.cpp:461-521: one warp, one lane per cluster, records in list order.
__global__ void __launch_bounds__(32) k_aec_report(AecDev* S, evk_aec_flow* out, int cap_out) {
    const int lane = threadIdx.x;
    const int nc = S->nc, min_n = S->min_n;
    int n_out = 0;
    for (int base = 0; base < nc; base += 32) {
        const int c = base + lane;
        bool take = false;
        evk_aec_flow r;
        if (c < nc && S->c_n[c] >= min_n) {
            take = true;
            r.id = S->c_id[c];
            r.n = S->c_n[c];
            if (r.id < 0 || r.id >= kMaxIds) {
                S->error = AEC_ERR_IDS;
                take = false;
            }
        }
        if (take) {
            aec_centroid(S, c, &r.centroid[0], &r.centroid[1]);
            double* prev = S->prev + 2 * r.id;
            r.prev[0] = prev[0];
            r.prev[1] = prev[1];
            r.has_arrow = (prev[0] > 0 && prev[1] > 0) ? 1 : 0;
            r._pad = 0;
            r.arrow_end[0] = __dadd_rn(prev[0], __dsub_rn(r.centroid[0], prev[0]));
            r.arrow_end[1] = __dadd_rn(prev[1], __dsub_rn(r.centroid[1], prev[1]));
            prev[0] = r.centroid[0];
            prev[1] = r.centroid[1];
        }
        const unsigned m = __ballot_sync(kFull, take);
        if (take) {
            const int p = n_out + __popc(m & ((1u << lane) - 1u));
            if (p < cap_out) out[p] = r;
        }
        n_out += __popc(m);
    }
    if (lane == 0) S->n_report = n_out;
}

// hand-off from the voxel shard (store.cpp:435-445): event k = representative of the voxel at
// canonical position start + k * step, pseudo-time t for all, polarity 0
__global__ void k_aec_gather(const uint32_t* __restrict__ xy, const uint32_t* __restrict__ perm,
                             size_t start, size_t step, size_t count, double t, double* out) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const uint32_t v = xy[perm[start + k * step]];
    out[4 * k] = t;
    out[4 * k + 1] = (double)(v & 0xFFFFu);
    out[4 * k + 2] = (double)(v >> 16);
    out[4 * k + 3] = 0.0;
}

__global__ void k_aec_init(AecDev* S, unsigned seed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < S->max_clusters) S->free_slots[i] = S->max_clusters - 1 - i;  // slot 0 on top
    for (int q = i; q < 2 * kMaxIds; q += gridDim.x * blockDim.x) S->prev[q] = 0.0;
    for (int q = i; q <= S->sz_buffer; q += gridDim.x * blockDim.x) S->tbuf[q] = 0.0;
    if (i == 0) {  // glibc srandom_r, TYPE_3: r[0..30] seeded, r[31..33] = r[0..2], 310 values discarded
        int32_t word = seed ? (int32_t)seed : 1;
        uint32_t* h = S->rhist;  // ring indexed by i & 127
        h[0] = (uint32_t)word;
        for (int k = 1; k < 31; k++) {
            const long long hi = word / 127773, lo = word % 127773;
            long long w = 16807 * lo - 2836 * hi;
            if (w < 0) w += 2147483647;
            word = (int32_t)w;
            h[k] = (uint32_t)word;
        }
        for (int k = 31; k < 34; k++) h[k] = h[k - 31];
        for (int k = 34; k < 344; k++) h[k & 127] = h[(k - 31) & 127] + h[(k - 3) & 127];
        S->rcount = 344;  // rand() number j is r[344 + j] >> 1
    }
}

}  // namespace

struct AecHost {
    AecDev* d = nullptr;       // header + arrays, one allocation
    AecDev hdr{};              // host copy of the header as created (pointers, parameters)
    size_t smem = 0;
    double* d_ev = nullptr;    // staging for update events
    size_t ev_cap = 0;         // events
    void* d_out = nullptr;     // staging for exports
    size_t out_bytes = 0;
};

static int aec_check(evk_handle* h) {
    if (!h) return EVK_ERR_INVALID;
    if (!h->aec) return evk_fail(h, EVK_ERR_STATE, "evk_aec_create has not been called");
    return EVK_OK;
}
static int aec_reserve_out(evk_handle* h, size_t bytes) {
    AecHost* a = h->aec;
    if (bytes <= a->out_bytes) return EVK_OK;
    if (a->d_out) cudaFree(a->d_out);
    a->d_out = nullptr;
    a->out_bytes = 0;
    if (cudaMalloc(&a->d_out, bytes) != cudaSuccess) {
        cudaGetLastError();
        return evk_fail(h, EVK_ERR_NOMEM, "aec export staging (%zu bytes)", bytes);
    }
    a->out_bytes = bytes;
    return EVK_OK;
}
static int aec_reserve_ev(evk_handle* h, size_t n) {
    AecHost* a = h->aec;
    if (n <= a->ev_cap) return EVK_OK;
    if (a->d_ev) cudaFree(a->d_ev);
    a->d_ev = nullptr;
    a->ev_cap = 0;
    const size_t cap = n < 4096 ? 4096 : n + n / 2;
    if (cudaMalloc(&a->d_ev, cap * 4 * sizeof(double)) != cudaSuccess) {
        cudaGetLastError();
        return evk_fail(h, EVK_ERR_NOMEM, "aec event staging (%zu events)", cap);
    }
    a->ev_cap = cap;
    return EVK_OK;
}
// run the update kernel over n staged events and report a capacity failure, if any
static int aec_run(evk_handle* h, size_t n) {
    AecHost* a = h->aec;
    if (n) {
        k_aec_update<<<1, 32, a->smem, h->stream>>>(a->d, a->d_ev, (long long)n);
        EVK_CUDA(h, cudaGetLastError());
    }
    struct {
        int error;
        long long error_event, events_done;
    } st;
    AecDev probe;
    EVK_CUDA(h, cudaMemcpyAsync(&probe, a->d, sizeof probe, cudaMemcpyDeviceToHost, h->stream));
    EVK_CUDA(h, cudaStreamSynchronize(h->stream));
    st.error = probe.error;
    st.error_event = probe.error_event;
    if (st.error == AEC_ERR_CLUSTERS)
        return evk_fail(h, EVK_ERR_CAPACITY, "aec: more than %d simultaneous clusters at event %lld",
                        a->hdr.max_clusters, st.error_event);
    if (st.error == AEC_ERR_POINTS)
        return evk_fail(h, EVK_ERR_CAPACITY, "aec: a cluster holds more than %d events at event %lld",
                        a->hdr.cap, st.error_event);
    if (st.error)
        return evk_fail(h, EVK_ERR_CAPACITY, "aec: cluster id beyond the %d-entry report table",
                        kMaxIds);
    return EVK_OK;
}

extern "C" {

int evk_aec_destroy(evk_handle* h) {
    if (!h || !h->aec) return EVK_OK;
    AecHost* a = h->aec;
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (a->d) cudaFree(a->d);
    if (a->d_ev) cudaFree(a->d_ev);
    if (a->d_out) cudaFree(a->d_out);
    delete a;
    h->aec = nullptr;
    return EVK_OK;
}

int evk_aec_create(evk_handle* h, const evk_aec_params* p) {
    if (!h) return EVK_ERR_INVALID;
    if (!p) return evk_fail(h, EVK_ERR_INVALID, "evk_aec_create: null parameters");
    evk_aec_destroy(h);
    AecDev hd;
    memset(&hd, 0, sizeof hd);
    // AEClustering::AEClustering(), AEClustering.cpp:7-18
    hd.min_n = 10;
    hd.sz_buffer = 800;
    hd.radius = 40;
    hd.alpha = 0.5;
    hd.kappa = 0;
    if (p->use_init) {  // AEClustering::init, :20-26
        hd.sz_buffer = p->sz_buffer;
        hd.radius = p->radius;
        hd.alpha = p->alpha;
        hd.min_n = p->min_n;
        hd.kappa = p->kappa;
    }
    hd.t0 = -1;
    hd.last = -1;
    hd.max_clusters = p->max_clusters > 0 ? p->max_clusters : kMaxC;
    hd.cap = p->max_points > 0 ? p->max_points : 4096;
    if (hd.sz_buffer < 1 || hd.sz_buffer > 16384)
        return evk_fail(h, EVK_ERR_INVALID, "aec: sz_buffer %d outside [1, 16384]", hd.sz_buffer);
    if (hd.kappa < 0 || hd.kappa > kMaxKappa)
        return evk_fail(h, EVK_ERR_INVALID, "aec: kappa %d outside [0, %d]", hd.kappa, kMaxKappa);
    if (hd.max_clusters > kMaxC)
        return evk_fail(h, EVK_ERR_INVALID, "aec: max_clusters %d > %d", hd.max_clusters, kMaxC);
    if (!(hd.radius >= 0) || !(hd.alpha == hd.alpha))
        return evk_fail(h, EVK_ERR_INVALID, "aec: radius / alpha");
    cudaSetDevice(h->device);
    // one allocation: header, then the arrays (8-byte ones first)
    const size_t np = (size_t)hd.max_clusters * hd.cap;
    size_t off = (sizeof(AecDev) + 15) & ~(size_t)15;
    auto take = [&](size_t bytes) {
        const size_t o = off;
        off += (bytes + 15) & ~(size_t)15;
        return o;
    };
    const size_t o_tbuf = take(8 * (size_t)(hd.sz_buffer + 1));
    const size_t o_mux = take(8 * kMaxC), o_muy = take(8 * kMaxC);
    const size_t o_px = take(8 * np), o_py = take(8 * np), o_pt = take(8 * np);
    const size_t o_mx = take(8 * (size_t)hd.cap), o_my = take(8 * (size_t)hd.cap);
    const size_t o_mt = take(8 * (size_t)hd.cap), o_mkp = take(8 * (size_t)hd.cap);
    const size_t o_prev = take(8 * 2 * (size_t)kMaxIds);
    const size_t o_slot = take(4 * kMaxC), o_id = take(4 * kMaxC), o_n = take(4 * kMaxC);
    const size_t o_head = take(4 * kMaxC), o_free = take(4 * (size_t)hd.max_clusters);
    const size_t o_pid = take(4 * np), o_mid = take(4 * (size_t)hd.cap);
    const size_t o_ppol = take(np), o_mpol = take((size_t)hd.cap);
    const size_t o_rh = take(4 * 128);
    unsigned char* base = nullptr;
    if (cudaMalloc(&base, off) != cudaSuccess) {
        cudaGetLastError();
        return evk_fail(h, EVK_ERR_NOMEM, "aec state (%zu bytes)", off);
    }
    hd.tbuf = (double*)(base + o_tbuf);
    hd.c_mux = (double*)(base + o_mux);
    hd.c_muy = (double*)(base + o_muy);
    hd.p_x = (double*)(base + o_px);
    hd.p_y = (double*)(base + o_py);
    hd.p_t = (double*)(base + o_pt);
    hd.m_x = (double*)(base + o_mx);
    hd.m_y = (double*)(base + o_my);
    hd.m_t = (double*)(base + o_mt);
    hd.m_kp = (double*)(base + o_mkp);
    hd.prev = (double*)(base + o_prev);
    hd.c_slot = (int*)(base + o_slot);
    hd.c_id = (int*)(base + o_id);
    hd.c_n = (int*)(base + o_n);
    hd.c_head = (int*)(base + o_head);
    hd.free_slots = (int*)(base + o_free);
    hd.p_id = (int*)(base + o_pid);
    hd.m_id = (int*)(base + o_mid);
    hd.p_pol = base + o_ppol;
    hd.m_pol = base + o_mpol;
    hd.rhist = (uint32_t*)(base + o_rh);
    hd.n_free = hd.max_clusters;
    AecHost* a = new AecHost;
    a->d = (AecDev*)base;
    a->hdr = hd;
    a->smem = aec_smem_bytes(hd.sz_buffer + 1);
    h->aec = a;
    cudaError_t ce = cudaMemcpyAsync(base, &hd, sizeof hd, cudaMemcpyHostToDevice, h->stream);
    if (ce == cudaSuccess)
        ce = cudaFuncSetAttribute(k_aec_update, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)a->smem);
    if (ce == cudaSuccess) {
        k_aec_init<<<64, 256, 0, h->stream>>>(a->d, p->rand_seed);
        ce = cudaGetLastError();
    }
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream);
    if (ce != cudaSuccess) {
        evk_aec_destroy(h);
        return evk_fail(h, EVK_ERR_CUDA, "evk_aec_create: %s", cudaGetErrorString(ce));
    }
    return EVK_OK;
}

int evk_aec_update(evk_handle* h, const double* e, size_t n) {
    EVK_TRY(aec_check(h));
    if (n && !e) return evk_fail(h, EVK_ERR_INVALID, "evk_aec_update: null events");
    cudaSetDevice(h->device);
    EVK_TRY(aec_reserve_ev(h, n));
    if (n)
        EVK_CUDA(h, cudaMemcpyAsync(h->aec->d_ev, e, n * 4 * sizeof(double), cudaMemcpyHostToDevice,
                                    h->stream));
    return aec_run(h, n);
}

int evk_aec_update_voxels(evk_handle* h, double t, size_t start, size_t step, size_t count) {
    EVK_TRY(aec_check(h));
    EVK_TRY(evk_collect_pending(h));
    if (!h->have_voxels) return evk_fail(h, EVK_ERR_STATE, "evk_aec_update_voxels: no voxel shard");
    if (step == 0) return evk_fail(h, EVK_ERR_INVALID, "evk_aec_update_voxels: step = 0");
    if (count && (start >= h->n_unique || (count - 1) > (h->n_unique - 1 - start) / step))
        return evk_fail(h, EVK_ERR_INVALID,
                        "evk_aec_update_voxels: %zu voxels from %zu by %zu exceed the %zu of the shard",
                        count, start, step, h->n_unique);
    cudaSetDevice(h->device);
    if (count == 0) return aec_run(h, 0);
    EVK_TRY(evk_ensure_perm(h));
    EVK_TRY(aec_reserve_ev(h, count));
    k_aec_gather<<<(unsigned)((count + 255) / 256), 256, 0, h->stream>>>(
        h->d_xy, h->d_perm, start, step, count, t, h->aec->d_ev);
    EVK_CUDA(h, cudaGetLastError());
    return aec_run(h, count);
}

int evk_aec_get_clusters(evk_handle* h, evk_aec_cluster* out, size_t cap, size_t* n,
                         int* last_updated) {
    EVK_TRY(aec_check(h));
    cudaSetDevice(h->device);
    AecHost* a = h->aec;
    AecDev probe;
    EVK_CUDA(h, cudaMemcpyAsync(&probe, a->d, sizeof probe, cudaMemcpyDeviceToHost, h->stream));
    EVK_CUDA(h, cudaStreamSynchronize(h->stream));
    if (n) *n = (size_t)probe.nc;
    if (last_updated) *last_updated = probe.last;
    if (!out || probe.nc == 0) return EVK_OK;
    if (cap < (size_t)probe.nc)
        return evk_fail(h, EVK_ERR_CAPACITY, "evk_aec_get_clusters: %d clusters, room for %zu",
                        probe.nc, cap);
    EVK_TRY(aec_reserve_out(h, (size_t)probe.nc * sizeof(evk_aec_cluster)));
    k_aec_export<<<(probe.nc + 63) / 64, 64, 0, h->stream>>>(a->d, (evk_aec_cluster*)a->d_out,
                                                            probe.nc);
    EVK_CUDA(h, cudaGetLastError());
    EVK_CUDA(h, cudaMemcpyAsync(out, a->d_out, (size_t)probe.nc * sizeof(evk_aec_cluster),
                                cudaMemcpyDeviceToHost, h->stream));
    EVK_CUDA(h, cudaStreamSynchronize(h->stream));
    return EVK_OK;
}

int evk_aec_get_points(evk_handle* h, size_t cluster, int32_t* ids, double* xy, double* t,
                       uint8_t* pol, size_t cap, size_t* n) {
    EVK_TRY(aec_check(h));
    cudaSetDevice(h->device);
    AecHost* a = h->aec;
    AecDev probe;
    EVK_CUDA(h, cudaMemcpyAsync(&probe, a->d, sizeof probe, cudaMemcpyDeviceToHost, h->stream));
    EVK_CUDA(h, cudaStreamSynchronize(h->stream));
    if (cluster >= (size_t)probe.nc)
        return evk_fail(h, EVK_ERR_INVALID, "evk_aec_get_points: cluster %zu of %d", cluster, probe.nc);
    int cn = 0;
    EVK_CUDA(h, cudaMemcpy(&cn, a->hdr.c_n + cluster, sizeof cn, cudaMemcpyDeviceToHost));
    if (n) *n = (size_t)cn;
    if (cn == 0 || (!ids && !xy && !t && !pol)) return EVK_OK;
    if (cap < (size_t)cn)
        return evk_fail(h, EVK_ERR_CAPACITY, "evk_aec_get_points: %d events, room for %zu", cn, cap);
    const size_t c8 = ((size_t)cn + 1) & ~(size_t)1;
    EVK_TRY(aec_reserve_out(h, c8 * (4 + 16 + 8 + 1) + 64));
    unsigned char* b = (unsigned char*)a->d_out;
    double* d_xy = (double*)b;
    double* d_t = d_xy + 2 * c8;
    int32_t* d_id = (int32_t*)(d_t + c8);
    uint8_t* d_pol = (uint8_t*)(d_id + c8);
    k_aec_points<<<(cn + 127) / 128, 128, 0, h->stream>>>(a->d, (int)cluster, d_id, d_xy, d_t, d_pol,
                                                         cn);
    EVK_CUDA(h, cudaGetLastError());
    if (ids) EVK_CUDA(h, cudaMemcpyAsync(ids, d_id, 4 * (size_t)cn, cudaMemcpyDeviceToHost, h->stream));
    if (xy) EVK_CUDA(h, cudaMemcpyAsync(xy, d_xy, 16 * (size_t)cn, cudaMemcpyDeviceToHost, h->stream));
    if (t) EVK_CUDA(h, cudaMemcpyAsync(t, d_t, 8 * (size_t)cn, cudaMemcpyDeviceToHost, h->stream));
    if (pol) EVK_CUDA(h, cudaMemcpyAsync(pol, d_pol, (size_t)cn, cudaMemcpyDeviceToHost, h->stream));
    EVK_CUDA(h, cudaStreamSynchronize(h->stream));
    return EVK_OK;
}

int evk_aec_report(evk_handle* h, evk_aec_flow* out, size_t cap, size_t* n) {
    EVK_TRY(aec_check(h));
    cudaSetDevice(h->device);
    AecHost* a = h->aec;
    EVK_TRY(aec_reserve_out(h, (size_t)kMaxC * sizeof(evk_aec_flow)));
    k_aec_report<<<1, 32, 0, h->stream>>>(a->d, (evk_aec_flow*)a->d_out, kMaxC);
    EVK_CUDA(h, cudaGetLastError());
    EVK_TRY(aec_run(h, 0));  // synchronises, reports an id beyond the table
    AecDev probe;
    EVK_CUDA(h, cudaMemcpy(&probe, a->d, sizeof probe, cudaMemcpyDeviceToHost));
    if (n) *n = (size_t)probe.n_report;
    if (!out || probe.n_report == 0) return EVK_OK;
    if (cap < (size_t)probe.n_report)
        return evk_fail(h, EVK_ERR_CAPACITY, "evk_aec_report: %d records, room for %zu",
                        probe.n_report, cap);
    EVK_CUDA(h, cudaMemcpy(out, a->d_out, (size_t)probe.n_report * sizeof(evk_aec_flow),
                           cudaMemcpyDeviceToHost));
    return EVK_OK;
}

}  // extern "C"
