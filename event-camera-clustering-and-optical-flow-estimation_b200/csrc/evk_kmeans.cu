// evk_kmeans.cu — fused k-means assign + accumulate and centroid finalise for sm_100a.
//
// Replaces the reference's three launches + four blocking reads per iteration
// (KM/assign_to_centers.cl:1-140, KM/assign_to_centers2.c:218-512):
//   assign_to_centers   -> argmin in registers, centroids broadcast from shared memory
//   assign_data_cluster -> gone: no scatter into 4096-float slabs (D12, D13)
//   reduction_scalar    -> block-private INTEGER partial sums in shared memory (coordinates are
//                          integers, so sums are exact and independent of summation order),
//                          merged with K*(D+1) 64-bit global atomics per 32 Ki points
//   host centroid update-> k_km_finalise on the device (exact sum / count, rounded once to fp32)
// Arithmetic contract: d2 = fmaf(dy,dy, dx*dx) (+ one fmaf per extra dimension),
// dx = c.x - p.x, strict '<' so the lowest k wins ties, gate d2 < max_dist^2, label -1 if none.
#include "evk_internal.cuh"

namespace {

constexpr int kBlock = 256;
constexpr int kPPT = 4;                        // points per thread per pass
constexpr int kChunk = kBlock * kPPT * 32;     // 32768 points between flushes: sums fit u32

// SRC: 0 = packed xy[i] (voxels, D == 2), 1 = raw events ev[i], 2 = gather ev[first[i]]
template <int D, int SRC>
__global__ void __launch_bounds__(kBlock)
    k_km_assign(KmLaunch kl, const uint32_t* __restrict__ xy, const evk_event* __restrict__ ev,
                const uint32_t* __restrict__ first, size_t n, const float* __restrict__ cent,
                unsigned long long* __restrict__ acc, int32_t* __restrict__ labels) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = kl.K;
    float* s_c = reinterpret_cast<float*>(smem_raw);                         // [K * D]
    unsigned long long* s_t = reinterpret_cast<unsigned long long*>(s_c + ((K * D + 3) & ~3));
    uint32_t* s_cnt = reinterpret_cast<uint32_t*>(s_t + K);                  // [K]
    uint32_t* s_x = s_cnt + K;
    uint32_t* s_y = s_x + K;
    uint32_t* s_p = s_y + K;

    for (int i = threadIdx.x; i < K * D; i += kBlock) s_c[i] = cent[i];
    for (int i = threadIdx.x; i < K; i += kBlock) {
        s_t[i] = 0;
        s_cnt[i] = 0;
        s_x[i] = 0;
        s_y[i] = 0;
        s_p[i] = 0;
    }
    __syncthreads();

    const size_t n_chunks = (n + kChunk - 1) / kChunk;
    for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const size_t cbase = c * (size_t)kChunk;
        for (int r = 0; r < kChunk / (kBlock * kPPT); r++) {
            const size_t base = cbase + (size_t)r * (kBlock * kPPT);
            if (base >= n) break;
            float px[kPPT], py[kPPT], pt[kPPT], pp[kPPT];
            uint32_t ix[kPPT], iy[kPPT], ip[kPPT];
            long long it[kPPT];
            bool ok[kPPT];
#pragma unroll
            for (int q = 0; q < kPPT; q++) {
                const size_t i = base + (size_t)q * kBlock + threadIdx.x;
                ok[q] = i < n;
                uint32_t w = 0;
                ip[q] = 0;
                it[q] = 0;
                if (ok[q]) {
                    if (SRC == 0) {
                        w = xy[i];
                    } else {
                        const evk_event* src = SRC == 2 ? ev + first[i] : ev + i;
                        if (D == 2) {
                            w = *reinterpret_cast<const uint32_t*>(src);
                        } else {
                            uint4 e = ld_event(src);
                            w = e.x;
                            ip[q] = ev_pbit(e);
                            it[q] = ev_t(e) - kl.t0;
                        }
                    }
                }
                ix[q] = w & 0xFFFFu;
                iy[q] = w >> 16;
                px[q] = (float)ix[q];
                py[q] = (float)iy[q];
                pt[q] = D > 2 ? __fmul_rn((float)it[q], kl.t_scale) : 0.f;
                pp[q] = D > 3 ? __fmul_rn(ip[q] ? 1.0f : 0.0f, kl.p_scale) : 0.f;
            }
            float best[kPPT];
            int lab[kPPT];
#pragma unroll
            for (int q = 0; q < kPPT; q++) {
                best[q] = kl.best2;
                lab[q] = -1;
            }
#pragma unroll 4
            for (int k = 0; k < K; k++) {
                float cx, cy, ct = 0.f, cp = 0.f;
                if (D == 2) {
                    float2 c2 = reinterpret_cast<const float2*>(s_c)[k];
                    cx = c2.x;
                    cy = c2.y;
                } else if (D == 4) {
                    float4 c4 = reinterpret_cast<const float4*>(s_c)[k];
                    cx = c4.x;
                    cy = c4.y;
                    ct = c4.z;
                    cp = c4.w;
                } else {
                    cx = s_c[k * 3];
                    cy = s_c[k * 3 + 1];
                    ct = s_c[k * 3 + 2];
                }
#pragma unroll
                for (int q = 0; q < kPPT; q++) {
                    float dx = __fsub_rn(cx, px[q]);
                    float dy = __fsub_rn(cy, py[q]);
                    float d2 = __fmaf_rn(dy, dy, __fmul_rn(dx, dx));
                    if (D > 2) {
                        float dt = __fsub_rn(ct, pt[q]);
                        d2 = __fmaf_rn(dt, dt, d2);
                    }
                    if (D > 3) {
                        float dp = __fsub_rn(cp, pp[q]);
                        d2 = __fmaf_rn(dp, dp, d2);
                    }
                    if (d2 < best[q]) {
                        best[q] = d2;
                        lab[q] = k;
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < kPPT; q++) {
                if (!ok[q]) continue;
                const size_t i = base + (size_t)q * kBlock + threadIdx.x;
                if (kl.write_labels) labels[i] = lab[q];
                if (lab[q] >= 0) {
                    atomicAdd(&s_cnt[lab[q]], 1u);
                    atomicAdd(&s_x[lab[q]], ix[q]);
                    atomicAdd(&s_y[lab[q]], iy[q]);
                    if (D > 2) atomicAdd(&s_t[lab[q]], (unsigned long long)it[q]);
                    if (D > 3) atomicAdd(&s_p[lab[q]], ip[q]);
                }
            }
        }
        // flush the block-private sums (u32 cannot overflow within one chunk: 32768 * 65535)
        __syncthreads();
        for (int k = threadIdx.x; k < K; k += kBlock) {
            if (s_cnt[k]) {
                atomicAdd(&acc[k * ACC_STRIDE + ACC_CNT], (unsigned long long)s_cnt[k]);
                atomicAdd(&acc[k * ACC_STRIDE + ACC_X], (unsigned long long)s_x[k]);
                atomicAdd(&acc[k * ACC_STRIDE + ACC_Y], (unsigned long long)s_y[k]);
                if (D > 2) atomicAdd(&acc[k * ACC_STRIDE + ACC_T], s_t[k]);
                if (D > 3) atomicAdd(&acc[k * ACC_STRIDE + ACC_P], (unsigned long long)s_p[k]);
                s_cnt[k] = 0;
                s_x[k] = 0;
                s_y[k] = 0;
                s_t[k] = 0;
                s_p[k] = 0;
            }
        }
        __syncthreads();
    }
}

// centroid update on the device (evk_km_finalise_body, evk_internal.cuh)
__global__ void __launch_bounds__(EVK_MAX_K)
    k_km_finalise(KmLaunch kl, float* cent, unsigned long long* acc, unsigned long long* counts,
                  float* shift) {
    evk_km_finalise_body(kl, cent, acc, counts, shift);
}

// candidates for "first K voxels in canonical order": voxels whose first index is below `bound`
__global__ void __launch_bounds__(kBlock)
    k_collect_below(const uint32_t* __restrict__ first, size_t n, uint32_t bound, uint32_t* cand,
                    uint32_t cand_cap, unsigned long long* count) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t f = first[i];
        if (f < bound) {
            unsigned long long o = atomicAdd(count, 1ull);
            if (o < cand_cap) {
                cand[2 * o] = f;
                cand[2 * o + 1] = (uint32_t)i;
            }
        }
    }
}

// rank the candidates by first index (all distinct) and write the K lowest as centroids
__global__ void __launch_bounds__(1024)
    k_init_from_cand(KmLaunch kl, const uint32_t* __restrict__ cand, uint32_t n_cand,
                     const uint32_t* __restrict__ xy, const evk_event* __restrict__ ev,
                     const evk_event* __restrict__ reps, uint64_t first_offset, float* cent) {
    for (uint32_t i = threadIdx.x; i < n_cand; i += blockDim.x) {
        const uint32_t f = cand[2 * i];
        uint32_t rank = 0;
        for (uint32_t j = 0; j < n_cand; j++) rank += cand[2 * j] < f;
        if (rank < (uint32_t)kl.K) {
            const uint32_t pos = cand[2 * i + 1];
            const uint32_t w = xy[pos];
            float* c = cent + rank * kl.D;
            c[0] = (float)(w & 0xFFFFu);
            c[1] = (float)(w >> 16);
            if (kl.D > 2) {
                const evk_event e = reps ? reps[pos] : ev[(uint64_t)f - first_offset];
                c[2] = __fmul_rn((float)(e.t - kl.t0), kl.t_scale);
                if (kl.D > 3) c[3] = __fmul_rn(e.p > 0 ? 1.0f : 0.0f, kl.p_scale);
            }
        }
    }
}

// ---- exact candidate pruning (D == 2, voxels) ------------------------------------------------
// The frame is cut into G x G pixel tiles.  For a tile T and centroid k let dmin(k) / dmax(k) be
// the smallest / largest distance from c_k to any pixel of T.  With ub = min_k dmax(k), the
// nearest centroid of EVERY pixel of T has dmin(k) <= ub, so only those k (ascending, which keeps
// the lowest-k tie rule) need the contract arithmetic.  The test carries a relative + absolute
// slack far above fp32 rounding, so the labels are bit-identical to the full scan; tiles with
// more than 16 candidates are marked and scanned in full.
constexpr int kListLen = 16;

__global__ void __launch_bounds__(128)
    k_km_candidates(KmLaunch kl, PruneGrid pg, const float* __restrict__ cent, uint4* lists) {
    extern __shared__ float2 s_cc[];
    for (int i = threadIdx.x; i < kl.K; i += blockDim.x)
        s_cc[i] = reinterpret_cast<const float2*>(cent)[i];
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= pg.tx * pg.ty) return;
    const int ty = t / pg.tx, tx = t - ty * pg.tx;
    const float x0 = (float)(tx << pg.shift), y0 = (float)(ty << pg.shift);
    const float x1 = (float)min(pg.width - 1, ((tx + 1) << pg.shift) - 1);
    const float y1 = (float)min(pg.height - 1, ((ty + 1) << pg.shift) - 1);
    float ub = INFINITY;
    for (int k = 0; k < kl.K; k++) {
        const float2 c = s_cc[k];
        const float ax = fmaxf(fabsf(c.x - x0), fabsf(c.x - x1));
        const float ay = fmaxf(fabsf(c.y - y0), fabsf(c.y - y1));
        ub = fminf(ub, ax * ax + ay * ay);
    }
    const float lim = ub * 1.0001f + 0.01f;
    unsigned char out[kListLen];
#pragma unroll
    for (int i = 0; i < kListLen; i++) out[i] = 0xFF;
    int cnt = 0;
    for (int k = 0; k < kl.K; k++) {
        const float2 c = s_cc[k];
        const float ix = fmaxf(0.f, fmaxf(x0 - c.x, c.x - x1));
        const float iy = fmaxf(0.f, fmaxf(y0 - c.y, c.y - y1));
        if (ix * ix + iy * iy <= lim) {
#pragma unroll
            for (int i = 0; i < kListLen; i++)
                if (i == cnt) out[i] = (unsigned char)k;
            cnt++;
        }
    }
    if (cnt > kListLen) out[0] = 0xFE;  // too many candidates: full scan for this tile
    uint4 w;
    w.x = out[0] | (out[1] << 8) | (out[2] << 16) | ((uint32_t)out[3] << 24);
    w.y = out[4] | (out[5] << 8) | (out[6] << 16) | ((uint32_t)out[7] << 24);
    w.z = out[8] | (out[9] << 8) | (out[10] << 16) | ((uint32_t)out[11] << 24);
    w.w = out[12] | (out[13] << 8) | (out[14] << 16) | ((uint32_t)out[15] << 24);
    lists[t] = w;
}

constexpr int kRep = 8;  // accumulator copies: lane l adds into copy l % 8 (fewer same-address hits)
__global__ void __launch_bounds__(kBlock)
    k_km_assign_pruned(KmLaunch kl, PruneGrid pg, const uint4* __restrict__ lists,
                       const uint32_t* __restrict__ xy, size_t n, const float* __restrict__ cent,
                       unsigned long long* __restrict__ acc, int32_t* __restrict__ labels) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = kl.K;
    float2* s_c = reinterpret_cast<float2*>(smem_raw);            // [K]
    uint32_t* s_acc = reinterpret_cast<uint32_t*>(s_c + K);       // [kRep][3][K]
    for (int i = threadIdx.x; i < K; i += kBlock) s_c[i] = reinterpret_cast<const float2*>(cent)[i];
    for (int i = threadIdx.x; i < kRep * 3 * K; i += kBlock) s_acc[i] = 0;
    __syncthreads();
    uint32_t* my_acc = s_acc + (threadIdx.x & (kRep - 1)) * 3 * K;

    const size_t n_chunks = (n + kChunk - 1) / kChunk;
    for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const size_t cbase = c * (size_t)kChunk;
        for (int r = 0; r < kChunk / (kBlock * kPPT); r++) {
            const size_t base = cbase + (size_t)r * (kBlock * kPPT);
            if (base >= n) break;
            uint32_t w[kPPT];
            uint4 lst[kPPT];
#pragma unroll
            for (int q = 0; q < kPPT; q++) {
                const size_t i = base + (size_t)q * kBlock + threadIdx.x;
                w[q] = i < n ? xy[i] : 0u;
            }
#pragma unroll
            for (int q = 0; q < kPPT; q++) {
                const uint32_t t = ((w[q] >> 16) >> pg.shift) * pg.tx + ((w[q] & 0xFFFFu) >> pg.shift);
                lst[q] = __ldg(lists + t);
            }
#pragma unroll
            for (int q = 0; q < kPPT; q++) {
                const size_t i = base + (size_t)q * kBlock + threadIdx.x;
                const bool ok = i < n;
                const uint32_t ix = w[q] & 0xFFFFu, iy = w[q] >> 16;
                const float px = (float)ix, py = (float)iy;
                float best = kl.best2;
                int lab = -1;
                const bool full = (lst[q].x & 0xFFu) == 0xFEu;
                if (__any_sync(0xffffffffu, full)) {
                    if (full) {
                        for (int k = 0; k < K; k++) {
                            const float2 cc = s_c[k];
                            const float dx = __fsub_rn(cc.x, px), dy = __fsub_rn(cc.y, py);
                            const float d2 = __fmaf_rn(dy, dy, __fmul_rn(dx, dx));
                            if (d2 < best) {
                                best = d2;
                                lab = k;
                            }
                        }
                    }
                }
                if (!full) {
                    const uint32_t lw[4] = {lst[q].x, lst[q].y, lst[q].z, lst[q].w};
#pragma unroll
                    for (int s = 0; s < kListLen; s++) {
                        const uint32_t k = (lw[s >> 2] >> (8 * (s & 3))) & 0xFFu;
                        if (k == 0xFFu) break;
                        const float2 cc = s_c[k];
                        const float dx = __fsub_rn(cc.x, px), dy = __fsub_rn(cc.y, py);
                        const float d2 = __fmaf_rn(dy, dy, __fmul_rn(dx, dx));
                        if (d2 < best) {
                            best = d2;
                            lab = (int)k;
                        }
                    }
                }
                if (ok) {
                    if (kl.write_labels) labels[i] = lab;
                    if (lab >= 0) {
                        atomicAdd(&my_acc[lab], 1u);
                        atomicAdd(&my_acc[K + lab], ix);
                        atomicAdd(&my_acc[2 * K + lab], iy);
                    }
                }
            }
        }
        __syncthreads();
        for (int k = threadIdx.x; k < K; k += kBlock) {
            unsigned long long sc = 0, sx = 0, sy = 0;
#pragma unroll
            for (int rr = 0; rr < kRep; rr++) {
                uint32_t* a = s_acc + rr * 3 * K;
                sc += a[k];
                sx += a[K + k];
                sy += a[2 * K + k];
                a[k] = 0;
                a[K + k] = 0;
                a[2 * K + k] = 0;
            }
            if (sc) {
                atomicAdd(&acc[k * ACC_STRIDE + ACC_CNT], sc);
                atomicAdd(&acc[k * ACC_STRIDE + ACC_X], sx);
                atomicAdd(&acc[k * ACC_STRIDE + ACC_Y], sy);
            }
        }
        __syncthreads();
    }
}

// ---- exact candidate pruning in space AND time (D == 3 / 4, voxels) ---------------------------
// The per-point scan of K centroids is issue-bound (K x 9 instructions per point).  Voxel
// representatives lie inside the frame and inside the downsample's time-bin range, so the same
// exact pruning as above works on boxes [16 x 16 px tile] x [time slab of 2^sh us] (x [0, p_scale]
// for D == 4): for a box B and centroid k, dmin(k) / dmax(k) are the smallest / largest distance
// from c_k to any point of B; only centroids with dmin(k) <= min_k dmax(k) (+ slack far above
// fp32 rounding) can be the nearest of a point of B.  lists[slab][tile] = up to 16 ascending
// centroid indices (keeps the lowest-k tie rule), 0xFE = scan all K.  A point looks its list up
// with its tile and the slab of its representative's timestamp, then runs the CONTRACT arithmetic
// on the listed centroids only: same operands, operations and order, bit-identical labels.
struct Prune3 {
    int32_t shift_us;   // slab width = 2^shift_us microseconds
    int32_t n_slabs;
    long long t_lo;     // first microsecond of slab 0
};

template <int D>
__global__ void __launch_bounds__(128)
    k_km_candidates3(KmLaunch kl, PruneGrid pg, Prune3 p3, const float* __restrict__ cent,
                     uint4* lists) {
    extern __shared__ float s_cc3[];  // [K * D]
    for (int i = threadIdx.x; i < kl.K * D; i += blockDim.x) s_cc3[i] = cent[i];
    __syncthreads();
    const int tiles = pg.tx * pg.ty;
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= tiles * p3.n_slabs) return;
    const int slab = id / tiles, t = id - slab * tiles;
    const int ty = t / pg.tx, tx = t - ty * pg.tx;
    float lo[4], hi[4];
    lo[0] = (float)(tx << pg.shift);
    lo[1] = (float)(ty << pg.shift);
    hi[0] = (float)min(pg.width - 1, ((tx + 1) << pg.shift) - 1);
    hi[1] = (float)min(pg.height - 1, ((ty + 1) << pg.shift) - 1);
    // the point coordinate is fmul_rn((float)(t - t0), t_scale): monotone in t
    const long long ta = p3.t_lo + ((long long)slab << p3.shift_us);
    const long long tb = ta + (1ll << p3.shift_us) - 1;
    lo[2] = __fmul_rn((float)(ta - kl.t0), kl.t_scale);
    hi[2] = __fmul_rn((float)(tb - kl.t0), kl.t_scale);
    if (lo[2] > hi[2]) {  // (negative t_scale)
        const float sw = lo[2];
        lo[2] = hi[2];
        hi[2] = sw;
    }
    lo[3] = fminf(0.f, kl.p_scale);
    hi[3] = fmaxf(0.f, kl.p_scale);
    // Reference centroid r = the one with the smallest worst-case distance at the slab's mid time.
    // Centroid k can be dropped when it is farther than r from EVERY point of the box:
    //   d_k - d_r = [s_k - s_r] + g(t),  s = squared distance in the spatial (+ polarity) dimensions,
    //   g(t) = (ct_k - t)^2 - (ct_r - t)^2 is LINEAR in t, so its minimum over the slab is at an end.
    // (A box bound that takes the worst t for k and for r independently loses 2 |dt| x slab width,
    // which is more than the spatial separation of the centroids once |dt| is a few hundred.)
    const float tm = 0.5f * (lo[2] + hi[2]);
    float ub = INFINITY;
    int kr = 0;
    for (int k = 0; k < kl.K; k++) {
        float m = 0.f;
#pragma unroll
        for (int d = 0; d < D; d++) {
            if (d == 2) continue;
            const float c = s_cc3[k * D + d];
            const float a = fmaxf(fabsf(c - lo[d]), fabsf(c - hi[d]));
            m += a * a;
        }
        const float dt = s_cc3[k * D + 2] - tm;
        m += dt * dt;
        if (m < ub) {
            ub = m;
            kr = k;
        }
    }
    float smax_r = 0.f;  // largest spatial (+ polarity) squared distance from r to the box
#pragma unroll
    for (int d = 0; d < D; d++) {
        if (d == 2) continue;
        const float c = s_cc3[kr * D + d];
        const float a = fmaxf(fabsf(c - lo[d]), fabsf(c - hi[d]));
        smax_r += a * a;
    }
    const float ctr = s_cc3[kr * D + 2];
    const float dra = ctr - lo[2], drb = ctr - hi[2];
    // slack: far above the rounding of this test and of the fp32 contract arithmetic it guards
    const float scale = smax_r + fmaxf(dra * dra, drb * drb);
    const float tol = scale * 1.0e-4f + 0.5f;
    unsigned char out[kListLen];
#pragma unroll
    for (int i = 0; i < kListLen; i++) out[i] = 0xFF;
    int cnt = 0;
    for (int k = 0; k < kl.K; k++) {
        float smin = 0.f;
#pragma unroll
        for (int d = 0; d < D; d++) {
            if (d == 2) continue;
            const float c = s_cc3[k * D + d];
            const float a = fmaxf(0.f, fmaxf(lo[d] - c, c - hi[d]));
            smin += a * a;
        }
        const float ctk = s_cc3[k * D + 2];
        const float ka = ctk - lo[2], kb = ctk - hi[2];
        const float g = fminf(ka * ka - dra * dra, kb * kb - drb * drb);
        const float ktol = tol + (smin + fmaxf(ka * ka, kb * kb)) * 1.0e-4f;
        if (smin - smax_r + g <= ktol) {
#pragma unroll
            for (int i = 0; i < kListLen; i++)
                if (i == cnt) out[i] = (unsigned char)k;
            cnt++;
        }
    }
    if (cnt > kListLen || !(ub < INFINITY)) out[0] = 0xFE;  // too many candidates: full scan
    uint4 w;
    w.x = out[0] | (out[1] << 8) | (out[2] << 16) | ((uint32_t)out[3] << 24);
    w.y = out[4] | (out[5] << 8) | (out[6] << 16) | ((uint32_t)out[7] << 24);
    w.z = out[8] | (out[9] << 8) | (out[10] << 16) | ((uint32_t)out[11] << 24);
    w.w = out[12] | (out[13] << 8) | (out[14] << 16) | ((uint32_t)out[15] << 24);
    lists[id] = w;
}

template <int D>
__device__ __forceinline__ float km_d2(const float* c, float px, float py, float pt, float pp) {
    const float dx = __fsub_rn(c[0], px), dy = __fsub_rn(c[1], py);
    float d2 = __fmaf_rn(dy, dy, __fmul_rn(dx, dx));
    const float dt = __fsub_rn(c[2], pt);
    d2 = __fmaf_rn(dt, dt, d2);
    if (D > 3) {
        const float dp = __fsub_rn(c[3], pp);
        d2 = __fmaf_rn(dp, dp, d2);
    }
    return d2;
}

template <int D>
__global__ void __launch_bounds__(kBlock)
    k_km_assign_pruned3(KmLaunch kl, PruneGrid pg, Prune3 p3, const uint4* __restrict__ lists,
                        const uint32_t* __restrict__ xy, const evk_event* __restrict__ ev,
                        const uint32_t* __restrict__ first, size_t n,
                        const float* __restrict__ cent, unsigned long long* __restrict__ acc,
                        int32_t* __restrict__ labels, int n_rep) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = kl.K;
    float* s_c = reinterpret_cast<float*>(smem_raw);                                   // [K * D]
    // sum of (t - t0): 16-bit limbs in u32 accumulators (a 64-bit shared atomicAdd is a CAS spin
    // loop); bits 32 and up, zero for streams shorter than 71 minutes, keep the 64-bit path
    unsigned long long* s_t = reinterpret_cast<unsigned long long*>(s_c + ((K * D + 3) & ~3));  // [n_rep][K]
    uint32_t* s_acc = reinterpret_cast<uint32_t*>(s_t + n_rep * K);                    // [n_rep][6][K]
    for (int i = threadIdx.x; i < K * D; i += kBlock) s_c[i] = cent[i];
    for (int i = threadIdx.x; i < n_rep * K; i += kBlock) s_t[i] = 0;
    for (int i = threadIdx.x; i < n_rep * 6 * K; i += kBlock) s_acc[i] = 0;
    __syncthreads();
    uint32_t* my_acc = s_acc + (threadIdx.x & (n_rep - 1)) * 6 * K;
    unsigned long long* my_t = s_t + (threadIdx.x & (n_rep - 1)) * K;
    const int tiles = pg.tx * pg.ty;

    const size_t n_chunks = (n + kChunk - 1) / kChunk;
    for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const size_t cbase = c * (size_t)kChunk;
        for (int r = 0; r < kChunk / (kBlock * kPPT); r++) {
            const size_t base = cbase + (size_t)r * (kBlock * kPPT);
            if (base >= n) break;
            uint32_t w[kPPT], ip[kPPT];
            long long it[kPPT];
            uint4 lst[kPPT];
#pragma unroll
            for (int q = 0; q < kPPT; q++) {  // all gathers of the pass in flight together
                const size_t i = base + (size_t)q * kBlock + threadIdx.x;
                w[q] = 0;
                it[q] = 0;
                ip[q] = 0;
                if (i < n) {
                    w[q] = xy[i];
                    const evk_event* src = ev + first[i];
                    if (D > 3) {
                        const uint4 e = ld_event(src);
                        ip[q] = ev_pbit(e);
                        it[q] = ev_t(e);
                    } else {
                        it[q] = __ldg(reinterpret_cast<const long long*>(
                            reinterpret_cast<const char*>(src) + 8));
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < kPPT; q++) {
                const uint32_t t = ((w[q] >> 16) >> pg.shift) * pg.tx + ((w[q] & 0xFFFFu) >> pg.shift);
                long long sl = (it[q] - p3.t_lo) >> p3.shift_us;
                // (a representative outside the announced time range: scan everything)
                const bool in = sl >= 0 && sl < p3.n_slabs;
                lst[q] = in ? __ldg(lists + (size_t)sl * tiles + t) : make_uint4(0xFEu, 0, 0, 0);
                it[q] -= kl.t0;
            }
#pragma unroll
            for (int q = 0; q < kPPT; q++) {
                const size_t i = base + (size_t)q * kBlock + threadIdx.x;
                const bool ok = i < n;
                const uint32_t ix = w[q] & 0xFFFFu, iy = w[q] >> 16;
                const float px = (float)ix, py = (float)iy;
                const float pt = __fmul_rn((float)it[q], kl.t_scale);
                const float pp = D > 3 ? __fmul_rn(ip[q] ? 1.0f : 0.0f, kl.p_scale) : 0.f;
                float best = kl.best2;
                int lab = -1;
                const bool full = (lst[q].x & 0xFFu) == 0xFEu;
                if (__any_sync(0xffffffffu, full)) {
                    if (full) {
                        for (int k = 0; k < K; k++) {
                            const float d2 = km_d2<D>(s_c + k * D, px, py, pt, pp);
                            if (d2 < best) {
                                best = d2;
                                lab = k;
                            }
                        }
                    }
                }
                if (!full) {
                    const uint32_t lw[4] = {lst[q].x, lst[q].y, lst[q].z, lst[q].w};
#pragma unroll
                    for (int s = 0; s < kListLen; s++) {
                        const uint32_t k = (lw[s >> 2] >> (8 * (s & 3))) & 0xFFu;
                        if (k == 0xFFu) break;
                        const float d2 = km_d2<D>(s_c + k * D, px, py, pt, pp);
                        if (d2 < best) {
                            best = d2;
                            lab = (int)k;
                        }
                    }
                }
                if (ok) {
                    if (kl.write_labels) labels[i] = lab;
                    if (lab >= 0) {
                        atomicAdd(&my_acc[lab], 1u);
                        atomicAdd(&my_acc[K + lab], ix);
                        atomicAdd(&my_acc[2 * K + lab], iy);
                        const unsigned long long ut = (unsigned long long)it[q];
                        atomicAdd(&my_acc[4 * K + lab], (uint32_t)(ut & 0xFFFFu));
                        atomicAdd(&my_acc[5 * K + lab], (uint32_t)((ut >> 16) & 0xFFFFu));
                        if (ut >> 32) atomicAdd(&my_t[lab], ut >> 32);
                        if (D > 3) atomicAdd(&my_acc[3 * K + lab], ip[q]);
                    }
                }
            }
        }
        __syncthreads();
        for (int k = threadIdx.x; k < K; k += kBlock) {
            unsigned long long sc = 0, sx = 0, sy = 0, st = 0, sp = 0;
#pragma unroll
            for (int rr = 0; rr < n_rep; rr++) {
                uint32_t* a = s_acc + rr * 6 * K;
                sc += a[k];
                sx += a[K + k];
                sy += a[2 * K + k];
                sp += a[3 * K + k];
                st += (unsigned long long)a[4 * K + k] + ((unsigned long long)a[5 * K + k] << 16) +
                      (s_t[rr * K + k] << 32);
#pragma unroll
                for (int f = 0; f < 6; f++) a[f * K + k] = 0;
                s_t[rr * K + k] = 0;
            }
            if (sc) {
                atomicAdd(&acc[k * ACC_STRIDE + ACC_CNT], sc);
                atomicAdd(&acc[k * ACC_STRIDE + ACC_X], sx);
                atomicAdd(&acc[k * ACC_STRIDE + ACC_Y], sy);
                atomicAdd(&acc[k * ACC_STRIDE + ACC_T], st);
                if (D > 3) atomicAdd(&acc[k * ACC_STRIDE + ACC_P], sp);
            }
        }
        __syncthreads();
    }
}

// ---- pixel-image k-means (D == 2, voxels, K <= 254) -------------------------------------------
// Voxel representatives have integer pixel coordinates, so with D == 2 everything an iteration
// needs from the voxel list is the histogram pixcnt[y][x] = number of representatives at that pixel
// (built once per downsample: by the slab kernel itself in the fused step, else by k_pix_hist).
//   assign      : label[y][x] = argmin over the tile's candidate list, the contract arithmetic
//                 evaluated once per PIXEL (same operands, operations and order as the per-point
//                 scan, so labels are bit-identical); 0xFF = unassigned (gate)
//   accumulate  : sum_k += pixcnt * (1, x, y) -- exact integers, same sums as adding point by point
// A Lloyd iteration therefore costs O(W * H) (0.9 M pixels for Gen4, L2-resident) instead of
// O(U) (55 M voxels); the per-voxel labels are one 1-byte gather at the end.
__global__ void __launch_bounds__(256)
    k_pix_hist(const uint32_t* __restrict__ xy, size_t n, const unsigned long long* n_dev,
               uint32_t width, uint32_t* pixcnt) {
    if (n_dev) n = (size_t)*n_dev;  // point count kept on the device (graph replay)
    const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
    for (size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i0 < n; i0 += stride) {
        uint32_t w[4];
        int m = 4;
        if (i0 + 4 <= n) {
            const uint4 v = __ldcs(reinterpret_cast<const uint4*>(xy + i0));
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        } else {
            m = (int)(n - i0);
            for (int q = 0; q < m; q++) w[q] = xy[i0 + q];
        }
        for (int q = 0; q < m; q++) atomicAdd(pixcnt + (size_t)(w[q] >> 16) * width + (w[q] & 0xFFFFu), 1u);
    }
}

// one thread per pixel: label into `map`, pixcnt-weighted sums into acc (warp-aggregated: a warp's
// 32 neighbouring pixels nearly always share one label)
__global__ void __launch_bounds__(256)
    k_km_image(KmLaunch kl, PruneGrid pg, const uint4* __restrict__ lists,
               const float* __restrict__ cent, const uint32_t* __restrict__ pixcnt,
               uint8_t* __restrict__ map, unsigned long long* __restrict__ acc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s_cc = reinterpret_cast<float2*>(smem_raw);                                // [K]
    unsigned long long* s_acc = reinterpret_cast<unsigned long long*>(s_cc + kl.K);    // [K][3]
    for (int i = threadIdx.x; i < kl.K; i += blockDim.x)
        s_cc[i] = reinterpret_cast<const float2*>(cent)[i];
    for (int i = threadIdx.x; i < 3 * kl.K; i += blockDim.x) s_acc[i] = 0;
    __syncthreads();
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const bool in = x < pg.width;
    uint32_t lab = 0xFFu, c = 0;
    if (in) {
        const uint4 l4 = __ldg(lists + (y >> pg.shift) * pg.tx + (x >> pg.shift));
        c = pixcnt ? pixcnt[(size_t)y * pg.width + x] : 0u;
        const float px = (float)x, py = (float)y;
        float best = kl.best2;
        if ((l4.x & 0xFFu) == 0xFEu) {  // more than 16 candidates: scan all K
            for (int k = 0; k < kl.K; k++) {
                const float2 cc = s_cc[k];
                const float dx = __fsub_rn(cc.x, px), dy = __fsub_rn(cc.y, py);
                const float d2 = __fmaf_rn(dy, dy, __fmul_rn(dx, dx));
                if (d2 < best) {
                    best = d2;
                    lab = (uint32_t)k;
                }
            }
        } else {
            const uint32_t lw[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
            for (int s = 0; s < kListLen; s++) {
                const uint32_t k = (lw[s >> 2] >> (8 * (s & 3))) & 0xFFu;
                if (k == 0xFFu) break;
                const float2 cc = s_cc[k];
                const float dx = __fsub_rn(cc.x, px), dy = __fsub_rn(cc.y, py);
                const float d2 = __fmaf_rn(dy, dy, __fmul_rn(dx, dx));
                if (d2 < best) {
                    best = d2;
                    lab = k;
                }
            }
        }
        map[(size_t)y * pg.width + x] = (uint8_t)lab;
    }
    // warp-aggregated accumulation: one leader per distinct label in the warp
    const int lane = threadIdx.x & 31;
    uint32_t todo = __ballot_sync(0xffffffffu, in && c != 0 && lab != 0xFFu);
    const uint64_t vx = (uint64_t)c * (uint32_t)x, vy = (uint64_t)c * (uint32_t)y;  // < 2^48
    while (todo) {
        const int leader = __ffs(todo) - 1;
        const uint32_t l0 = __shfl_sync(0xffffffffu, lab, leader);
        const uint32_t grp = __ballot_sync(0xffffffffu, lab == l0) & todo;
        const bool mine = (grp >> lane) & 1u;
        // 24-bit limbs: a warp's sum of 32 limbs fits 32 bits
        const uint32_t c0 = __reduce_add_sync(0xffffffffu, mine ? (c & 0xFFFFFFu) : 0u);
        const uint32_t c1 = __reduce_add_sync(0xffffffffu, mine ? (c >> 24) : 0u);
        const uint32_t x0 = __reduce_add_sync(0xffffffffu, mine ? (uint32_t)(vx & 0xFFFFFFu) : 0u);
        const uint32_t x1 = __reduce_add_sync(0xffffffffu, mine ? (uint32_t)(vx >> 24) : 0u);
        const uint32_t y0 = __reduce_add_sync(0xffffffffu, mine ? (uint32_t)(vy & 0xFFFFFFu) : 0u);
        const uint32_t y1 = __reduce_add_sync(0xffffffffu, mine ? (uint32_t)(vy >> 24) : 0u);
        if (lane == leader) {
            atomicAdd(&s_acc[l0 * 3 + 0], (unsigned long long)c0 + ((unsigned long long)c1 << 24));
            atomicAdd(&s_acc[l0 * 3 + 1], (unsigned long long)x0 + ((unsigned long long)x1 << 24));
            atomicAdd(&s_acc[l0 * 3 + 2], (unsigned long long)y0 + ((unsigned long long)y1 << 24));
        }
        todo &= ~grp;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < kl.K; k += blockDim.x) {
        if (s_acc[k * 3]) {
            atomicAdd(&acc[k * ACC_STRIDE + ACC_CNT], s_acc[k * 3 + 0]);
            atomicAdd(&acc[k * ACC_STRIDE + ACC_X], s_acc[k * 3 + 1]);
            atomicAdd(&acc[k * ACC_STRIDE + ACC_Y], s_acc[k * 3 + 2]);
        }
    }
}

// quads[t] = the label shared by every pixel of the (1 << shift)-pixel square t, or 0xFE (mixed)
__global__ void __launch_bounds__(128)
    k_km_quads(const uint8_t* __restrict__ map, int width, int height, int shift, int qx, int qy,
               uint8_t* __restrict__ quads) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= qx * qy) return;
    const int ty = t / qx, tx = t - ty * qx;
    const int x0 = tx << shift, y0 = ty << shift;
    const int x1 = min(width, x0 + (1 << shift)), y1 = min(height, y0 + (1 << shift));
    const uint8_t first = map[(size_t)y0 * width + x0];
    bool uniform = true;
    for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++) uniform &= map[(size_t)y * width + x] == first;
    quads[t] = uniform ? first : (uint8_t)0xFE;
}

// Per-voxel pass, two levels: the label map is piecewise constant (Voronoi cells), so most small
// squares ("quads", 8 x 8 px for Gen4) carry ONE label -- a byte from a <= 16 KB table in shared
// memory; only points in mixed quads (0xFE) gather their label from the pixel map (L2; a random
// L2 access per point costs an SM about a cycle per lane, so it pays to keep them rare).
// ACC: also accumulate the exact sums (single-iteration path); else labels only (gather at the end
// of the image iterations).  Accumulators: one private column per lane (bank == lane: no conflicts)
// shared by the CTA's warps through 32-bit shared atomics, flushed with 64-bit global atomics.
constexpr int kTilesBlock = 512;        // two CTAs per SM share the SM's shared memory: a table of
                                        // 4 x 4-px quads for Gen4 (57.6 KB) + the accumulators
constexpr int kUnit = kTilesBlock * 8;  // points per work unit
constexpr int kFlushUnits = 32;

__device__ __forceinline__ void tiles_flush(uint32_t* s_acc, int K, int copies,
                                            unsigned long long* acc) {
    __syncthreads();
    for (int j = threadIdx.x; j < 3 * K; j += kTilesBlock) {
        unsigned long long sum = 0;
        for (int cc = 0; cc < copies; cc++) {
            const int idx = j * copies + ((cc + threadIdx.x) & (copies - 1));  // skewed: no conflicts
            sum += s_acc[idx];
            s_acc[idx] = 0;
        }
        const int which = j / K, k = j - which * K;  // 0: count, 1: sum x, 2: sum y
        if (sum)
            atomicAdd(&acc[k * ACC_STRIDE + (which == 0 ? ACC_CNT : which == 1 ? ACC_X : ACC_Y)], sum);
    }
    __syncthreads();
}

template <bool ACC>
__global__ void __launch_bounds__(kTilesBlock, 2)
    k_km_assign_tiles(KmLaunch kl, QuadGrid pg, uint32_t tl_bytes, const uint8_t* __restrict__ quads,
                      const uint8_t* __restrict__ map, const uint32_t* __restrict__ xy, size_t n,
                      const unsigned long long* __restrict__ n_dev, int copies,
                      unsigned long long* __restrict__ acc, int32_t* __restrict__ labels) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint8_t* s_tl = smem_raw;                                            // [tl_bytes]
    uint32_t* s_acc = reinterpret_cast<uint32_t*>(smem_raw + tl_bytes);  // [3][K][copies]
    const int K = kl.K;
    if (n_dev) n = (size_t)*n_dev;  // voxel count still on the device (fused step)
    for (int t = threadIdx.x; t < (pg.tx * pg.ty + 3) / 4; t += kTilesBlock)  // table is padded to 4
        reinterpret_cast<uint32_t*>(s_tl)[t] = __ldg(reinterpret_cast<const uint32_t*>(quads) + t);
    if (ACC)
        for (int i = threadIdx.x; i < 3 * K * copies; i += kTilesBlock) s_acc[i] = 0;
    __syncthreads();
    const uint32_t col = threadIdx.x & (copies - 1);
    // work unit: kUnit points = two 16-B loads per thread, units dealt round-robin to the CTAs
    const size_t n_units = (n + kUnit - 1) / kUnit;
    int since_flush = 0;
    // the next unit's coordinates are requested before the current unit is processed
    auto fetch = [&](size_t u, uint32_t (&w)[8]) {
#pragma unroll
        for (int hlf = 0; hlf < 2; hlf++) {
            const size_t i0 = u * (size_t)kUnit + (size_t)hlf * (kUnit / 2) + (size_t)threadIdx.x * 4;
            if (i0 + 4 <= n) {
                const uint4 v = __ldcs(reinterpret_cast<const uint4*>(xy + i0));
                w[4 * hlf + 0] = v.x; w[4 * hlf + 1] = v.y; w[4 * hlf + 2] = v.z; w[4 * hlf + 3] = v.w;
            } else {
#pragma unroll
                for (int q = 0; q < 4; q++) w[4 * hlf + q] = i0 + q < n ? xy[i0 + q] : 0u;
            }
        }
    };
    uint32_t wn[8];
    if (blockIdx.x < n_units) fetch(blockIdx.x, wn);
    for (size_t u = blockIdx.x; u < n_units; u += gridDim.x) {
        const size_t ubase = u * (size_t)kUnit;
        uint32_t w[8], lab[8];
        size_t i0[2];
        bool full[2];
#pragma unroll
        for (int q = 0; q < 8; q++) w[q] = wn[q];
        if (u + gridDim.x < n_units) fetch(u + gridDim.x, wn);
#pragma unroll
        for (int hlf = 0; hlf < 2; hlf++) {
            i0[hlf] = ubase + (size_t)hlf * (kUnit / 2) + (size_t)threadIdx.x * 4;
            full[hlf] = i0[hlf] + 4 <= n;
        }
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const uint32_t px = w[q] & 0xFFFFu, py = w[q] >> 16;
            lab[q] = s_tl[(py >> pg.shift) * pg.tx + (px >> pg.shift)];  // quad
        }
#pragma unroll
        for (int q = 0; q < 8; q++)
            if (lab[q] == 0xFEu)
                lab[q] = __ldg(map + (size_t)(w[q] >> 16) * pg.width + (w[q] & 0xFFFFu));
        if (!ACC || kl.write_labels) {
#pragma unroll
            for (int hlf = 0; hlf < 2; hlf++) {
                int4 o;
                o.x = lab[4 * hlf + 0] == 0xFFu ? -1 : (int)lab[4 * hlf + 0];
                o.y = lab[4 * hlf + 1] == 0xFFu ? -1 : (int)lab[4 * hlf + 1];
                o.z = lab[4 * hlf + 2] == 0xFFu ? -1 : (int)lab[4 * hlf + 2];
                o.w = lab[4 * hlf + 3] == 0xFFu ? -1 : (int)lab[4 * hlf + 3];
                if (full[hlf]) {
                    __stcs(reinterpret_cast<int4*>(labels + i0[hlf]), o);
                } else {
                    const int ov[4] = {o.x, o.y, o.z, o.w};
                    for (int q = 0; q < 4; q++)
                        if (i0[hlf] + q < n) labels[i0[hlf] + q] = ov[q];
                }
            }
        }
        if (ACC) {
#pragma unroll
            for (int q = 0; q < 8; q++) {
                if (i0[q >> 2] + (q & 3) < n && lab[q] != 0xFFu) {
                    uint32_t* a = s_acc + lab[q] * copies + col;
                    atomicAdd(a, 1u);
                    atomicAdd(a + K * copies, w[q] & 0xFFFFu);
                    atomicAdd(a + 2 * K * copies, w[q] >> 16);
                }
            }
            // per unit a column receives 8 points from each of the kTilesBlock / 32 = 16 warps,
            // times 32 / copies lanes per column (copies >= 8 for K <= 254): at most 512 points,
            // and 512 * kFlushUnits * 65535 < 2^32
            if (++since_flush == kFlushUnits) {
                since_flush = 0;
                tiles_flush(s_acc, K, copies, acc);
            }
        }
    }
    if (ACC) tiles_flush(s_acc, K, copies, acc);
}

// "first K voxel representatives in canonical order" == the first K distinct keys met when the
// stream is walked in order.  One warp walks the head of the stream: lane l holds event base+l,
// the 32 events are committed one after the other, the list of keys found so far is compared
// 32 entries at a time.  A few hundred events suffice for realistic streams; `found` tells the
// host when the walk ran out (it then falls back to the rank-by-first-index path).
__global__ void __launch_bounds__(32)
    k_init_first_k_walk(KeyParams kp, KmLaunch kl, const evk_event* __restrict__ ev, size_t n_scan,
                        float* cent, unsigned long long* found_out, const long long* t0_dev) {
    __shared__ uint64_t s_keys[EVK_MAX_K];
    if (t0_dev) kp.t0 = *t0_dev;  // (replayed graphs: the time origin is a device-side value)
    const int lane = threadIdx.x;
    int found = 0;
    for (size_t base = 0; base < n_scan && found < kl.K; base += 32) {
        const size_t i = base + lane;
        uint4 e = make_uint4(0, 0, 0, 0);
        uint64_t key = 0;
        bool valid = false;
        if (i < n_scan) {
            e = ld_event(ev + i);
            valid = evk_key(kp, e, key);
        }
        const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
        for (int l = 0; l < 32 && found < kl.K; l++) {
            if (!((vmask >> l) & 1u)) continue;
            const uint64_t kl_key = __shfl_sync(0xffffffffu, key, l);
            bool match = false;
            for (int q = lane; q < found; q += 32) match |= s_keys[q] == kl_key;
            if (!__any_sync(0xffffffffu, match)) {
                if (lane == l) {
                    s_keys[found] = key;
                    float* c = cent + found * kl.D;
                    c[0] = (float)ev_x(e);
                    c[1] = (float)ev_y(e);
                    if (kl.D > 2) c[2] = __fmul_rn((float)(ev_t(e) - kl.t0), kl.t_scale);
                    if (kl.D > 3) c[3] = __fmul_rn(ev_pbit(e) ? 1.0f : 0.0f, kl.p_scale);
                }
                found++;
                __syncwarp();
            }
        }
    }
    if (lane == 0) *found_out = (unsigned long long)found;
}

size_t km_smem_bytes(int K, int D) {
    return (size_t)((K * D + 3) & ~3) * sizeof(float) + (size_t)K * sizeof(unsigned long long) +
           (size_t)K * 4 * sizeof(uint32_t);
}

template <int D, int SRC>
cudaError_t launch_assign(const KmLaunch& kl, const uint32_t* xy, const evk_event* ev,
                          const uint32_t* first, size_t n, const float* cent,
                          unsigned long long* acc, int32_t* labels, int sm_count,
                          cudaStream_t s) {
    size_t chunks = (n + kChunk - 1) / kChunk;
    size_t cap = (size_t)sm_count * 8;
    int grid = (int)(chunks < cap ? chunks : cap);
    k_km_assign<D, SRC><<<grid, kBlock, km_smem_bytes(kl.K, D), s>>>(kl, xy, ev, first, n, cent,
                                                                      acc, labels);
    return cudaGetLastError();
}

}  // namespace

// xy != nullptr: D == 2 reads the packed voxel coordinates; D > 2 gathers ev[first[i]].
// xy == nullptr: points are the raw events ev[i].
cudaError_t evk_launch_km_assign(const KmLaunch& kl, const uint32_t* xy, const evk_event* ev,
                                 const uint32_t* first, size_t n, const float* cent,
                                 unsigned long long* acc, int32_t* labels, int sm_count,
                                 cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const bool on_events = xy == nullptr;
    switch (kl.D) {
        case 2:
            return on_events
                       ? launch_assign<2, 1>(kl, xy, ev, first, n, cent, acc, labels, sm_count, s)
                       : launch_assign<2, 0>(kl, xy, ev, first, n, cent, acc, labels, sm_count, s);
        case 3:
            return on_events
                       ? launch_assign<3, 1>(kl, xy, ev, first, n, cent, acc, labels, sm_count, s)
                       : launch_assign<3, 2>(kl, xy, ev, first, n, cent, acc, labels, sm_count, s);
        case 4:
            return on_events
                       ? launch_assign<4, 1>(kl, xy, ev, first, n, cent, acc, labels, sm_count, s)
                       : launch_assign<4, 2>(kl, xy, ev, first, n, cent, acc, labels, sm_count, s);
    }
    return cudaErrorInvalidValue;
}

// D == 2 on the voxel shard with 16 < K <= 254: candidate lists per tile, then the pruned scan.
// lists must hold EVK_PRUNE_TILES uint4.  Returns cudaErrorNotSupported when the shape does not fit.
cudaError_t evk_launch_km_assign_pruned(const KmLaunch& kl, int width, int height, void* lists,
                                        const uint32_t* xy, size_t n, const float* cent,
                                        unsigned long long* acc, int32_t* labels, int sm_count,
                                        cudaStream_t s) {
    if (kl.D != 2 || kl.K <= 16 || kl.K > 254) return cudaErrorNotSupported;
    const PruneGrid pg = evk_make_prune_grid(width, height);
    if (n == 0) return cudaSuccess;
    evk_launch_km_candidates(kl, pg, cent, lists, s);
    size_t chunks = (n + kChunk - 1) / kChunk;
    size_t cap = (size_t)sm_count * 8;
    int grid = (int)(chunks < cap ? chunks : cap);
    size_t smem = (size_t)kl.K * sizeof(float2) + (size_t)kRep * 3 * kl.K * sizeof(uint32_t);
    k_km_assign_pruned<<<grid, kBlock, smem, s>>>(kl, pg, reinterpret_cast<const uint4*>(lists),
                                                  xy, n, cent, acc, labels);
    return cudaGetLastError();
}

// D == 3 / 4 on voxels whose representatives' timestamps lie in [t_lo, t_hi] (the downsample's
// time-bin range): lists3 holds n_slabs x tiles candidate lists (EVK_PRUNE3_LISTS at most)
cudaError_t evk_launch_km_assign_pruned3(const KmLaunch& kl, int width, int height, void* lists3,
                                         long long t_lo, long long t_hi, const uint32_t* xy,
                                         const evk_event* ev, const uint32_t* first, size_t n,
                                         const float* cent, unsigned long long* acc,
                                         int32_t* labels, int sm_count, cudaStream_t s) {
    if ((kl.D != 3 && kl.D != 4) || kl.K <= 8 || kl.K > 254 || t_hi < t_lo)
        return cudaErrorNotSupported;
    if (n == 0) return cudaSuccess;
    const PruneGrid pg = evk_make_prune_grid(width, height);
    const int tiles = pg.tx * pg.ty;
    Prune3 p3;
    const int want = EVK_PRUNE3_LISTS / tiles < 64 ? EVK_PRUNE3_LISTS / tiles : 64;
    if (want < 1) return cudaErrorNotSupported;
    p3.t_lo = t_lo;
    p3.shift_us = 0;
    const unsigned long long span = (unsigned long long)(t_hi - t_lo) + 1;
    while (((span + (1ull << p3.shift_us) - 1) >> p3.shift_us) > (unsigned long long)want) p3.shift_us++;
    p3.n_slabs = (int)((span + (1ull << p3.shift_us) - 1) >> p3.shift_us);
    const int n_lists = tiles * p3.n_slabs;
    const size_t csmem = (size_t)kl.K * kl.D * sizeof(float);
    if (kl.D == 3)
        k_km_candidates3<3><<<(n_lists + 127) / 128, 128, csmem, s>>>(
            kl, pg, p3, cent, reinterpret_cast<uint4*>(lists3));
    else
        k_km_candidates3<4><<<(n_lists + 127) / 128, 128, csmem, s>>>(
            kl, pg, p3, cent, reinterpret_cast<uint4*>(lists3));
    size_t chunks = (n + kChunk - 1) / kChunk;
    size_t cap = (size_t)sm_count * 8;
    int grid = (int)(chunks < cap ? chunks : cap);
    const int n_rep = kl.K > 128 ? 4 : 8;  // accumulator copies (fewer same-address atomics)
    const size_t smem = (size_t)((kl.K * kl.D + 3) & ~3) * sizeof(float) +
                        (size_t)n_rep * kl.K * sizeof(unsigned long long) +
                        (size_t)n_rep * 6 * kl.K * sizeof(uint32_t);
    if (smem > 48 * 1024) return cudaErrorNotSupported;
    if (kl.D == 3)
        k_km_assign_pruned3<3><<<grid, kBlock, smem, s>>>(
            kl, pg, p3, reinterpret_cast<const uint4*>(lists3), xy, ev, first, n, cent, acc, labels,
            n_rep);
    else
        k_km_assign_pruned3<4><<<grid, kBlock, smem, s>>>(
            kl, pg, p3, reinterpret_cast<const uint4*>(lists3), xy, ev, first, n, cent, acc, labels,
            n_rep);
    return cudaGetLastError();
}

PruneGrid evk_make_prune_grid(int width, int height) {
    PruneGrid pg;
    pg.width = width;
    pg.height = height;
    pg.shift = 2;
    for (;;) {
        pg.tx = (width + (1 << pg.shift) - 1) >> pg.shift;
        pg.ty = (height + (1 << pg.shift) - 1) >> pg.shift;
        if ((long long)pg.tx * pg.ty <= EVK_PRUNE_TILES) break;
        pg.shift++;
    }
    return pg;
}

cudaError_t evk_launch_km_candidates(const KmLaunch& kl, const PruneGrid& pg, const float* cent,
                                     void* lists, cudaStream_t s) {
    const int tiles = pg.tx * pg.ty;
    k_km_candidates<<<(tiles + 127) / 128, 128, kl.K * sizeof(float2), s>>>(
        kl, pg, cent, reinterpret_cast<uint4*>(lists));
    return cudaGetLastError();
}

cudaError_t evk_launch_pix_hist(const uint32_t* xy, size_t n, int width, int height,
                                uint32_t* pixcnt, int sm_count, cudaStream_t s,
                                const unsigned long long* n_dev) {
    cudaError_t e = cudaMemsetAsync(pixcnt, 0, (size_t)width * height * sizeof(uint32_t), s);
    if (e != cudaSuccess || n == 0) return e;
    const size_t need = (n + 1023) / 1024;
    const size_t cap = (size_t)sm_count * 16;
    k_pix_hist<<<(int)(need < cap ? need : cap), 256, 0, s>>>(xy, n, n_dev, (uint32_t)width, pixcnt);
    return cudaGetLastError();
}

// one Lloyd assign + accumulate over the pixel image: candidate lists -> labels + weighted sums
QuadGrid evk_make_quad_grid(int width, int height) {
    QuadGrid qg;
    qg.width = width;
    qg.shift = 1;
    for (;;) {
        qg.tx = (width + (1 << qg.shift) - 1) >> qg.shift;
        qg.ty = (height + (1 << qg.shift) - 1) >> qg.shift;
        if ((long long)qg.tx * qg.ty <= EVK_MAX_QUADS) break;
        qg.shift++;
    }
    return qg;
}

// quads != nullptr: also the table of uniformly labelled squares for evk_launch_km_assign_tiles
cudaError_t evk_launch_km_image(const KmLaunch& kl, int width, int height, void* lists,
                                const float* cent, const uint32_t* pixcnt, uint8_t* map,
                                uint8_t* quads, unsigned long long* acc, cudaStream_t s) {
    if (kl.D != 2 || kl.K > 254 || !map) return cudaErrorNotSupported;
    const PruneGrid pg = evk_make_prune_grid(width, height);
    evk_launch_km_candidates(kl, pg, cent, lists, s);
    dim3 grid((width + 255) / 256, height);
    const size_t smem = (size_t)kl.K * (sizeof(float2) + 3 * sizeof(unsigned long long));
    k_km_image<<<grid, 256, smem, s>>>(kl, pg, reinterpret_cast<const uint4*>(lists), cent, pixcnt,
                                       map, acc);
    if (quads) {
        const QuadGrid qg = evk_make_quad_grid(width, height);
        k_km_quads<<<(qg.tx * qg.ty + 127) / 128, 128, 0, s>>>(map, width, height, qg.shift, qg.tx,
                                                              qg.ty, quads);
    }
    return cudaGetLastError();
}

// Per-voxel pass over the voxel shard with the two-level label lookup.  accumulate: also the exact
// sums (one Lloyd assign + accumulate); the candidate lists and the pixel map must already have
// been built from the current centroids (evk_launch_km_image with or without a histogram).
// n_dev != nullptr: the voxel count is read on the device, n is only an upper bound for the grid.
cudaError_t evk_launch_km_assign_tiles(const KmLaunch& kl, int width, int height,
                                       const uint8_t* quads, const uint8_t* map,
                                       const uint32_t* xy, size_t n,
                                       const unsigned long long* n_dev, bool accumulate,
                                       unsigned long long* acc, int32_t* labels, int sm_count,
                                       cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const QuadGrid pg = evk_make_quad_grid(width, height);
    int copies = 32;
    while (copies > 1 && (size_t)3 * kl.K * copies * 4 > 48 * 1024) copies >>= 1;
    const uint32_t tl_bytes = (uint32_t)((pg.tx * pg.ty + 15) & ~15);
    const size_t smem = tl_bytes + (accumulate ? (size_t)3 * kl.K * copies * 4 : 0);
    // (a function attribute is per device: set on every call, whichever device the handle lives on)
    if (accumulate)
        cudaFuncSetAttribute(k_km_assign_tiles<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             EVK_MAX_QUADS + 48 * 1024);
    else
        cudaFuncSetAttribute(k_km_assign_tiles<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             EVK_MAX_QUADS);
    // one wave: as many CTAs as are resident at once, units dealt round-robin
    int per_sm = 1;
    if (accumulate)
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_km_assign_tiles<true>, kTilesBlock, smem);
    else
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_km_assign_tiles<false>, kTilesBlock, smem);
    if (per_sm < 1) per_sm = 1;
    const size_t units = (n + kUnit - 1) / kUnit;
    const size_t cap = (size_t)sm_count * per_sm;
    const int grid = (int)(units < cap ? units : cap);
    if (accumulate) {
        k_km_assign_tiles<true><<<grid, kTilesBlock, smem, s>>>(kl, pg, tl_bytes, quads, map, xy, n,
                                                                n_dev, copies, acc, labels);
    } else {
        k_km_assign_tiles<false><<<grid, kTilesBlock, smem, s>>>(kl, pg, tl_bytes, quads, map, xy, n,
                                                                 n_dev, copies, acc, labels);
    }
    return cudaGetLastError();
}

__global__ void k_set_u64(unsigned long long* dst, unsigned long long v) { *dst = v; }
cudaError_t evk_launch_set_u64(unsigned long long* dst, unsigned long long v, cudaStream_t s) {
    k_set_u64<<<1, 1, 0, s>>>(dst, v);
    return cudaGetLastError();
}

cudaError_t evk_launch_km_finalise(const KmLaunch& kl, float* cent, unsigned long long* acc,
                                   unsigned long long* counts, float* shift, cudaStream_t s) {
    int threads = ((kl.K + 31) / 32) * 32;
    k_km_finalise<<<1, threads, 0, s>>>(kl, cent, acc, counts, shift);
    return cudaGetLastError();
}

cudaError_t evk_launch_collect_below(const uint32_t* first, size_t n, uint32_t bound,
                                     uint32_t* cand, uint32_t cand_cap, unsigned long long* count,
                                     cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    size_t need = (n + kBlock - 1) / kBlock;
    int grid = (int)(need < 148 * 8 ? need : 148 * 8);
    k_collect_below<<<grid, kBlock, 0, s>>>(first, n, bound, cand, cand_cap, count);
    return cudaGetLastError();
}

cudaError_t evk_launch_init_from_cand(const KmLaunch& kl, const uint32_t* cand, uint32_t n_cand,
                                      const uint32_t* xy, const evk_event* ev,
                                      const evk_event* reps, float* cent, cudaStream_t s) {
    // first_offset is folded into ev by the caller (ev points at global index 0 of the shard)
    k_init_from_cand<<<1, 1024, 0, s>>>(kl, cand, n_cand, xy, ev, reps, 0, cent);
    return cudaGetLastError();
}

cudaError_t evk_launch_init_first_k_walk(const KeyParams& kp, const KmLaunch& kl,
                                         const evk_event* ev, size_t n_scan, float* cent,
                                         unsigned long long* found, cudaStream_t s,
                                         const long long* t0_dev) {
    k_init_first_k_walk<<<1, 32, 0, s>>>(kp, kl, ev, n_scan, cent, found, t0_dev);
    return cudaGetLastError();
}
