// evk_slab.cu — time-slab downsample: the sm_100a fast path for time-ordered event streams.
//
// Event streams are time-ordered (the SDK delivers them so: ACCEL/store.cpp:614-615), and the time
// bin is the most significant part of the voxel key, so all events of one time bin are contiguous.
// One CTA owns one time bin at a time and keeps the WHOLE key space of that bin on chip:
//   s_seen   1 bit per spatial cell (NX*NY*P bits; 57.6 KB for Gen4 2x2 px + polarity)
//   s_rep    1 bit per cell "hit at least twice" (the kernel's repeated_count semantics,
//            ACCEL/build/coordinate_processor.cl:73-75)
//   s_hash   a 4096-slot tile table of packed (cell << 11 | index-in-tile) words: a 32-bit
//            atomicCAS claims, a 32-bit atomicMin keeps the LOWEST stream index
// Global memory sees each event read once (128-bit loads) and each voxel written once (16 B SoA);
// no table lives in HBM.  This is the "perfect-hash" specialisation of the mandated
// open-addressing table (evk_downsample.cu), which remains the general path: the slab kernel
// verifies on the fly that [bin_start[b], bin_start[b+1]) holds only events of bin b and that the
// ranges partition the stream; any violation makes the host fall back to the table.
#include "evk_internal.cuh"

namespace {

constexpr int kThreads = 1024;
constexpr int kLogTile = 11;
constexpr int kTile = 1 << kLogTile;       // events per tile
constexpr int kPer = kTile / kThreads;     // events per thread per tile
constexpr int kLogHash = 12;
constexpr int kHash = 1 << kLogHash;       // tile table slots (load <= 0.5)
constexpr int kSlotsPer = kHash / kThreads;
constexpr uint32_t kEmpty = 0xFFFFFFFFu;

struct SlabArgs {
    KeyParams kp;
    const evk_event* ev;
    size_t n;
    uint32_t* bin_start;
    uint64_t* keys;
    uint32_t* first;
    uint32_t* xy;
    DsCounters* cnt;
    uint32_t words;  // bitmap words per bin
    uint32_t max_bins;
    uint32_t min_bins;
    int count_repeated;
};

// scratch[0] = n_bins, scratch[1] = work counter, scratch[2] = tb0
__global__ void __launch_bounds__(256) k_slab_bins(SlabArgs a) {
    DsCounters* cnt = a.cnt;
    const KeyParams& kp = a.kp;
    const uint4 e0 = ld_event(a.ev);
    const uint4 e1 = ld_event(a.ev + (a.n - 1));
    const int64_t t_first = ev_t(e0), t_last = ev_t(e1);
    bool bad = t_first < kp.t0 || t_last < t_first;
    uint64_t tb0 = 0, nb = 0;
    if (!bad) {
        tb0 = evk_tbin(kp, t_first);
        nb = evk_tbin(kp, t_last) - tb0 + 1;
        if (nb > a.max_bins || nb < a.min_bins) bad = true;
    }
    if (bad) {
        if (blockIdx.x == 0 && threadIdx.x == 0) cnt->slab_violation = 2;
        return;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        cnt->scratch[0] = nb;
        cnt->scratch[2] = tb0;
    }
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b <= nb; b += stride) {
        size_t lo = 0, hi = a.n;  // lower_bound: first i with tbin(ev[i]) >= tb0 + b
        if (b == 0) hi = 0;
        else if (b == nb) lo = a.n;
        while (lo < hi) {
            size_t mid = (lo + hi) >> 1;
            const int64_t t = *reinterpret_cast<const int64_t*>(
                reinterpret_cast<const char*>(a.ev + mid) + 8);
            bool ge = t >= kp.t0 && evk_tbin(kp, t) >= tb0 + b;
            if (ge) hi = mid;
            else lo = mid + 1;
        }
        a.bin_start[b] = (uint32_t)lo;
    }
}

__device__ __forceinline__ uint32_t hash_slot(uint32_t cell) {
    return (cell * 0x9E3779B1u) >> (32 - kLogHash);
}

template <bool COUNT_REP>
__global__ void __launch_bounds__(kThreads, 1) k_slab_main(SlabArgs a) {
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t* s_seen = smem;
    uint32_t* s_rep = s_seen + a.words;                          // only touched when COUNT_REP
    uint32_t* s_hash = s_rep + (COUNT_REP ? a.words : 0);        // [kHash]
    uint32_t* s_xy = s_hash + kHash;                             // [kTile]
    __shared__ int s_warp[kThreads / 32 + 1];
    __shared__ unsigned long long s_base;
    __shared__ uint32_t s_bin;

    DsCounters* cnt = a.cnt;
    if (cnt->slab_violation) return;
    const KeyParams& kp = a.kp;
    const uint32_t nb = (uint32_t)cnt->scratch[0];
    const uint64_t tb0 = cnt->scratch[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int i = tid; i < kHash; i += kThreads) s_hash[i] = kEmpty;

    for (;;) {
        __syncthreads();  // previous bin fully retired (also orders the s_hash init)
        if (tid == 0) s_bin = (uint32_t)atomicAdd(&cnt->scratch[1], 1ull);
        __syncthreads();
        const uint32_t b = s_bin;
        if (b >= nb) break;
        const uint32_t lo = a.bin_start[b], hi = a.bin_start[b + 1];
        if (hi < lo) {  // ranges do not partition the stream: not time-ordered
            if (tid == 0) atomicOr(&cnt->slab_violation, 1u);
            continue;
        }
        if (hi == lo) continue;
        const uint64_t tb = tb0 + b;
        for (uint32_t i = tid; i < a.words; i += kThreads) {
            s_seen[i] = 0;
            if (COUNT_REP) s_rep[i] = 0;
        }
        // prefetch the first tile
        uint4 e[kPer];
#pragma unroll
        for (int j = 0; j < kPer; j++) {
            uint32_t i = lo + j * kThreads + tid;
            if (i < hi) e[j] = ld_event(a.ev + i);
        }
        __syncthreads();

        for (uint32_t base = lo; base < hi; base += kTile) {
            // ---- phase A: classify against the bin bitmap, insert candidates in the tile table
#pragma unroll
            for (int j = 0; j < kPer; j++) {
                const uint32_t li = j * kThreads + tid;
                const uint32_t i = base + li;
                if (i >= hi) continue;
                const uint4 ev = e[j];
                s_xy[li] = ev.x;
                const int64_t t = ev_t(ev);
                if (ev_x(ev) >= (uint32_t)kp.width || ev_y(ev) >= (uint32_t)kp.height) continue;
                if (t < kp.t0 || evk_tbin(kp, t) != tb) {
                    atomicOr(&cnt->slab_violation, 1u);
                    continue;
                }
                const uint32_t cell = evk_cell(kp, ev);
                const uint32_t w = cell >> 5, bit = 1u << (cell & 31);
                if (s_seen[w] & bit) {  // seen in an earlier tile: duplicate
                    if (COUNT_REP && !(s_rep[w] & bit)) atomicOr(&s_rep[w], bit);
                    continue;
                }
                const uint32_t v = (cell << kLogTile) | li;
                uint32_t s = hash_slot(cell);
                for (;;) {
                    uint32_t old = atomicCAS(&s_hash[s], kEmpty, v);
                    if (old == kEmpty) break;
                    if ((old >> kLogTile) == cell) {
                        atomicMin(&s_hash[s], v);
                        if (COUNT_REP && !(s_rep[w] & bit)) atomicOr(&s_rep[w], bit);
                        break;
                    }
                    s = (s + 1) & (kHash - 1);
                }
            }
            // prefetch the next tile while the table is drained
#pragma unroll
            for (int j = 0; j < kPer; j++) {
                uint32_t i = base + kTile + j * kThreads + tid;
                if (i < hi) e[j] = ld_event(a.ev + i);
            }
            __syncthreads();
            // ---- phase B: every occupied slot is one new voxel with its lowest index
            uint32_t v[kSlotsPer];
            int mine = 0;
#pragma unroll
            for (int j = 0; j < kSlotsPer; j++) {
                const int s = j * kThreads + tid;
                v[j] = s_hash[s];
                if (v[j] != kEmpty) {
                    mine++;
                    s_hash[s] = kEmpty;
                    const uint32_t cell = v[j] >> kLogTile;
                    atomicOr(&s_seen[cell >> 5], 1u << (cell & 31));
                }
            }
            // block exclusive scan of `mine`
            int inc = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int nn = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += nn;
            }
            if (lane == 31) s_warp[warp] = inc;
            __syncthreads();
            if (warp == 0) {
                int wv = s_warp[lane];
                int winc = wv;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    int nn = __shfl_up_sync(0xffffffffu, winc, o);
                    if (lane >= o) winc += nn;
                }
                s_warp[lane] = winc - wv;
                if (lane == 31) {
                    s_warp[32] = winc;
                    if (winc) s_base = atomicAdd(&cnt->n_unique, (unsigned long long)winc);
                }
            }
            __syncthreads();
            if (s_warp[32]) {
                size_t o = (size_t)s_base + s_warp[warp] + inc - mine;
#pragma unroll
                for (int j = 0; j < kSlotsPer; j++) {
                    if (v[j] != kEmpty) {
                        const uint32_t cell = v[j] >> kLogTile, li = v[j] & (kTile - 1);
                        a.keys[o] = tb * kp.cells + cell;
                        a.first[o] = base + li;
                        a.xy[o] = s_xy[li];
                        o++;
                    }
                }
            }
            __syncthreads();  // s_xy / s_hash / s_warp reusable
        }
        if (COUNT_REP) {
            int r = 0;
            for (uint32_t i = tid; i < a.words; i += kThreads) r += __popc(s_rep[i]);
            r = __reduce_add_sync(0xffffffffu, r);
            if (lane == 0 && r) atomicAdd(&cnt->n_repeated, (unsigned long long)r);
        }
    }
}

size_t slab_smem_bytes(uint32_t words, bool count_rep) {
    return ((size_t)words * (count_rep ? 2 : 1) + kHash + kTile) * sizeof(uint32_t);
}
constexpr size_t kSmemLimit = 200 * 1024;

}  // namespace

bool evk_slab_supported(const evk_handle* h, const KeyParams& kp) {
    if (kp.keyfn != EVK_KEY_VOXEL || kp.vt <= 0 || h->n_events == 0) return false;
    if (kp.cells >= (1ull << (32 - kLogTile))) return false;  // packed (cell, index) word
    const uint32_t words = (uint32_t)((kp.cells + 31) / 32);
    return slab_smem_bytes(words, true) <= kSmemLimit;
}

int evk_downsample_slab(evk_handle* h, const KeyParams& kp, int count_repeated, bool* ok,
                        int* launches) {
    *ok = false;
    SlabArgs a;
    a.kp = kp;
    a.ev = h->d_events;
    a.n = h->n_events;
    a.bin_start = h->d_bin_start;
    a.keys = h->d_keys;
    a.first = h->d_first;
    a.xy = h->d_xy;
    a.cnt = h->d_cnt;
    a.words = (uint32_t)((kp.cells + 31) / 32);
    a.max_bins = (uint32_t)h->max_bins;
    // a single CTA walks a bin sequentially: with few bins and many events the table is faster
    a.min_bins = h->n_events > (1u << 22) ? 32 : 1;
    a.count_repeated = count_repeated;
    const size_t smem = slab_smem_bytes(a.words, count_repeated != 0);
    if (count_repeated)
        EVK_CUDA(h, cudaFuncSetAttribute(k_slab_main<true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kSmemLimit));
    else
        EVK_CUDA(h, cudaFuncSetAttribute(k_slab_main<false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kSmemLimit));
    k_slab_bins<<<h->sm_count, 256, 0, h->stream>>>(a);
    EVK_CUDA(h, cudaGetLastError());
    if (h->profiling) cudaEventRecord(h->ev[5], h->stream);
    if (count_repeated)
        k_slab_main<true><<<h->sm_count, kThreads, smem, h->stream>>>(a);
    else
        k_slab_main<false><<<h->sm_count, kThreads, smem, h->stream>>>(a);
    EVK_CUDA(h, cudaGetLastError());
    if (h->profiling) cudaEventRecord(h->ev[6], h->stream);
    *launches += 2;
    // the verification flag decides whether the result stands
    EVK_CUDA(h, cudaMemcpyAsync(h->h_cnt, h->d_cnt, sizeof(DsCounters), cudaMemcpyDeviceToHost,
                                h->stream));
    EVK_CUDA(h, cudaStreamSynchronize(h->stream));
    *ok = h->h_cnt->slab_violation == 0;
    return EVK_OK;
}
