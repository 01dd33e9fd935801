// evk_partition.cu — unordered streams: stable partition by time bin, then the time-slab kernel.
//
// The slab kernel (evk_slab.cu) needs every time bin's events contiguous.  A stream that is not
// time-ordered (shuffled packets, a replay stitched from several files, events before t0) used to
// drop to the global open-addressing table (evk_downsample.cu: DRAM-random, 16 ms for 100 M
// events).  Here the stream is first brought into bin order by ONE stable counting pass on the
// time bin -- stable, so that inside a bin the original stream order survives and "lowest
// position wins" in the slab kernel still means "lowest stream index wins"  -- and the
// slab kernel runs on the partitioned copy; the voxels' first indices are translated back
// through the index map the scatter wrote.  Gated events (outside the frame, before t0) are
// dropped by the partition.
//
//   k_part_minmax   min / max time bin of the gated-in events              (16 N bytes read)
//   k_part_hist     one warp per strip of <= 32 K events: private shared-memory histogram over
//                   the bins, written out as a row hist[strip][bin]         (16 N read)
//   k_part_colsum / _binscan / _offsets
//                   offsets[strip][bin] = start of the bin + events of the bin in earlier strips
//   k_part_scatter  one warp per strip walks its events in rows of 32, in order: the rank of a lane
//                   among the lanes of its row with the same bin (one ballot per distinct bin of the
//                   row) + the strip's running cursor of that bin = its slot
//                   (16 N read, 16 N + 4 N written)
//   k_part_translate  first[v] = orig_index[first[v]]                       (8 U)
// The key space per strip is bounded by shared memory (u32 cursor per bin and warp); streams
// with more bins than fit take the table path as before.
#include "evk_internal.cuh"

namespace {

constexpr int kBlock = 256;
constexpr uint32_t kStrip = 16384;  // events per strip (one warp walks a strip in order)
constexpr int kRows = 8;            // rows of 32 events a warp keeps in flight (4 KB per warp)

struct PartArgs {
    KeyParams kp;
    const evk_event* ev;
    size_t n;
    uint64_t tb_min;   // first time bin
    uint32_t nb;       // number of digits of this pass (bins, or groups of `div` bins)
    uint32_t div;      // digit = (time bin - tb_min) / div
    uint32_t n_strips;
    unsigned long long* spread;  // += number of non-empty digits of every strip (histogram pass)
};

__device__ __forceinline__ bool part_bin(const KeyParams& kp, const uint4& e, uint64_t& tb) {
    const uint32_t x = ev_x(e), y = ev_y(e);
    const int64_t t = ev_t(e);
    if (x >= (uint32_t)kp.width || y >= (uint32_t)kp.height || t < kp.t0) return false;
    tb = evk_tbin(kp, t);
    return true;
}
__device__ __forceinline__ bool part_digit(const PartArgs& a, const uint4& e, uint32_t& d) {
    uint64_t tb;
    if (!part_bin(a.kp, e, tb)) return false;
    d = (uint32_t)(tb - a.tb_min);
    if (a.div > 1) d /= a.div;
    return true;
}

// stats[0] = min bin, stats[1] = max bin, stats[2] = gated-in events
__global__ void __launch_bounds__(kBlock)
    k_part_minmax(KeyParams kp, const evk_event* __restrict__ ev, size_t n,
                  unsigned long long* stats) {
    unsigned long long lo = ~0ull, hi = 0, cnt = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint4 e = ld_event(ev + i);
        uint64_t tb;
        if (part_bin(kp, e, tb)) {
            lo = tb < lo ? tb : lo;
            hi = tb > hi ? tb : hi;
            cnt++;
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const unsigned long long l2 = __shfl_xor_sync(0xffffffffu, lo, o);
        const unsigned long long h2 = __shfl_xor_sync(0xffffffffu, hi, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        lo = l2 < lo ? l2 : lo;
        hi = h2 > hi ? h2 : hi;
    }
    if ((threadIdx.x & 31) == 0 && cnt) {
        atomicMin(&stats[0], lo);
        atomicMax(&stats[1], hi);
        atomicAdd(&stats[2], cnt);
    }
}

// one warp per strip; dynamic shared memory: [warps per CTA][nb] u32
__global__ void __launch_bounds__(1024) k_part_hist(PartArgs a, uint32_t* hist) {
    extern __shared__ uint32_t s_cnt[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    uint32_t* mine = s_cnt + (size_t)warp * a.nb;
    for (uint32_t s0 = blockIdx.x * wpc; s0 < a.n_strips; s0 += gridDim.x * wpc) {
        const uint32_t s = s0 + warp;
        for (uint32_t b = lane; b < a.nb; b += 32) mine[b] = 0;
        __syncwarp();
        if (s < a.n_strips) {
            const size_t lo = (size_t)s * kStrip;
            const size_t hi = lo + kStrip < a.n ? lo + kStrip : a.n;
            for (size_t i0 = lo + lane; i0 < hi; i0 += 32 * kRows) {  // kRows loads in flight
                uint4 e[kRows];
#pragma unroll
                for (int r = 0; r < kRows; r++)
                    if (i0 + 32 * r < hi) e[r] = ld_event(a.ev + i0 + 32 * r);
#pragma unroll
                for (int r = 0; r < kRows; r++) {
                    uint32_t d;
                    if (i0 + 32 * r < hi && part_digit(a, e[r], d)) atomicAdd(&mine[d], 1u);
                }
            }
            __syncwarp();
            uint32_t* row = hist + (size_t)s * a.nb;
            uint32_t nz = 0;
            for (uint32_t b = lane; b < a.nb; b += 32) {
                const uint32_t c = mine[b];
                row[b] = c;
                nz += c != 0;
            }
            nz = __reduce_add_sync(0xffffffffu, nz);
            if (lane == 0 && a.spread) atomicAdd(a.spread, (unsigned long long)nz);
        }
        __syncwarp();
    }
}

// The strip x digit matrix is scanned column-wise in chunks of kChunkStrips strips so that the three
// small kernels below have (chunks x digits) parallelism instead of one thread per digit.
constexpr uint32_t kChunkStrips = 64;

// partial[c][b] = sum of hist[s][b] over the strips of chunk c
__global__ void __launch_bounds__(kBlock)
    k_part_colsum(const uint32_t* __restrict__ hist, uint32_t nb, uint32_t n_strips,
                  uint32_t* partial) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const uint32_t s0 = blockIdx.y * kChunkStrips;
    const uint32_t s1 = s0 + kChunkStrips < n_strips ? s0 + kChunkStrips : n_strips;
    uint32_t t = 0;
#pragma unroll 8
    for (uint32_t s = s0; s < s1; s++) t += hist[(size_t)s * nb + b];
    partial[(size_t)blockIdx.y * nb + b] = t;
}

// one CTA: partial[c][b] -> exclusive prefix over the chunks of digit b; binstart = exclusive scan of
// the digit totals (binstart[nb] = number of partitioned events)
__global__ void __launch_bounds__(1024)
    k_part_binscan(uint32_t* partial, uint32_t n_chunks, uint32_t* binstart, uint32_t nb) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t b0 = 0; b0 <= nb; b0 += 1024) {
        const uint32_t b = b0 + threadIdx.x;
        uint32_t v = 0;
        if (b < nb)
            for (uint32_t c = 0; c < n_chunks; c++) {
                const size_t at = (size_t)c * nb + b;
                const uint32_t x = partial[at];
                partial[at] = v;
                v += x;
            }
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        const uint32_t excl = carry + (wid ? s_warp[wid - 1] : 0u) + inc - v;
        if (b <= nb) binstart[b] = excl;  // (b == nb: the grand total, v = 0)
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + s_warp[31];
        __syncthreads();
    }
}

// hist[s][b] -> first slot of strip s in digit b
__global__ void __launch_bounds__(kBlock)
    k_part_offsets(uint32_t* hist, uint32_t nb, uint32_t n_strips,
                   const uint32_t* __restrict__ binstart, const uint32_t* __restrict__ partial) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const uint32_t s0 = blockIdx.y * kChunkStrips;
    const uint32_t s1 = s0 + kChunkStrips < n_strips ? s0 + kChunkStrips : n_strips;
    uint32_t run = binstart[b] + partial[(size_t)blockIdx.y * nb + b];
    for (uint32_t s = s0; s < s1; s++) {
        const size_t at = (size_t)s * nb + b;
        const uint32_t c = hist[at];
        hist[at] = run;
        run += c;
    }
}

__global__ void __launch_bounds__(1024)
    k_part_scatter(PartArgs a, const uint32_t* __restrict__ offsets, evk_event* out,
                   uint32_t* orig, const uint32_t* __restrict__ orig_in) {
    extern __shared__ uint32_t s_cnt[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    uint32_t* cur = s_cnt + (size_t)warp * a.nb;
    const uint32_t lt = (1u << lane) - 1u;
    for (uint32_t s0 = blockIdx.x * wpc; s0 < a.n_strips; s0 += gridDim.x * wpc) {
        const uint32_t s = s0 + warp;
        if (s < a.n_strips) {
            const uint32_t* row = offsets + (size_t)s * a.nb;
            for (uint32_t b = lane; b < a.nb; b += 32) cur[b] = row[b];
            __syncwarp();
            const size_t lo = (size_t)s * kStrip;
            const size_t hi = lo + kStrip < a.n ? lo + kStrip : a.n;
            for (size_t g0 = lo; g0 < hi; g0 += 32 * kRows) {  // kRows loads in flight, rows in order
                uint4 ev[kRows];
#pragma unroll
                for (int r = 0; r < kRows; r++) {
                    const size_t i = g0 + 32 * r + lane;
                    ev[r] = i < hi ? ld_event(a.ev + i) : make_uint4(0xFFFFFFFFu, 0, 0, 0);
                }
#pragma unroll
                for (int r = 0; r < kRows; r++) {
                    const size_t i = g0 + 32 * r + lane;
                    const uint4 e = ev[r];
                    uint32_t dg = 0;
                    const bool ok = i < hi && part_digit(a, e, dg);
                    const uint32_t bin = ok ? dg : 0xFFFFFFFFu;
                    const uint32_t okm = __ballot_sync(0xffffffffu, ok);
                    if (!okm) continue;
                    // lanes of the row that share my digit: one ballot per distinct digit of the row
                    // (a time-ordered row is uniform: one round; match.any costs more than 32 rounds)
                    uint32_t peers = 0, rest = okm;
                    while (rest) {
                        const uint32_t b0 = __shfl_sync(0xffffffffu, bin, __ffs(rest) - 1);
                        const uint32_t m = __ballot_sync(0xffffffffu, bin == b0);
                        if (bin == b0) peers = m;
                        rest &= ~m;
                    }
                    if (ok) {
                        const uint32_t base = cur[bin];  // (read by every peer before the update)
                        const uint32_t pos = base + __popc(peers & lt);
                        reinterpret_cast<uint4*>(out)[pos] = e;
                        orig[pos] = orig_in ? orig_in[i] : (uint32_t)i;
                    }
                    __syncwarp();
                    if (ok && (peers & lt) == 0) cur[bin] += __popc(peers);  // lowest lane of the group
                    __syncwarp();
                }
            }
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(kBlock)
    k_part_translate(uint32_t* first, const unsigned long long* n_dev,
                     const uint32_t* __restrict__ orig, uint32_t first_offset) {
    const size_t n = (size_t)*n_dev;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        first[i] = orig[first[i] - first_offset] + first_offset;
}

}  // namespace

// one stable counting pass: events `in` (n of them) -> `out` ordered by digit; orig_out[pos] = the
// original index of the event now at pos (through orig_in when this is the second level).
// Leaves binstart[0 .. nd] (exclusive prefix of the digit totals) behind the strip x digit matrix.
static int partition_pass(evk_handle* h, PartArgs a, const uint32_t* orig_in, evk_event* out,
                          uint32_t* orig_out, bool hist_done, int* launches) {
    constexpr size_t kSmem = 200 * 1024;
    int wpc = (int)(kSmem / ((size_t)a.nb * 4));
    wpc = wpc > 32 ? 32 : wpc;
    const size_t smem = (size_t)wpc * a.nb * 4;
    uint32_t* hist = h->d_part_hist;
    const uint32_t n_chunks = (a.n_strips + kChunkStrips - 1) / kChunkStrips;
    uint32_t* binstart = hist + (size_t)a.n_strips * a.nb;  // [nb + 1]
    uint32_t* partial = binstart + a.nb + 2;                // [n_chunks][nb]
    int grid = (int)((a.n_strips + wpc - 1) / wpc);
    const int ctas_per_sm = smem <= 100 * 1024 ? 2 : 1;
    if (grid > h->sm_count * ctas_per_sm) grid = h->sm_count * ctas_per_sm;
    if (!hist_done) {
        k_part_hist<<<grid, wpc * 32, smem, h->stream>>>(a, hist);
        (*launches)++;
    }
    if (!out) return cudaGetLastError() == cudaSuccess ? EVK_OK : evk_fail(h, EVK_ERR_CUDA, "partition histogram");
    const dim3 g2((a.nb + kBlock - 1) / kBlock, n_chunks);
    k_part_colsum<<<g2, kBlock, 0, h->stream>>>(hist, a.nb, a.n_strips, partial);
    k_part_binscan<<<1, 1024, 0, h->stream>>>(partial, n_chunks, binstart, a.nb);
    k_part_offsets<<<g2, kBlock, 0, h->stream>>>(hist, a.nb, a.n_strips, binstart, partial);
    k_part_scatter<<<grid, wpc * 32, smem, h->stream>>>(a, hist, out, orig_out, orig_in);
    EVK_CUDA(h, cudaGetLastError());
    *launches += 4;
    return EVK_OK;
}

// Partition h->d_events by time bin into h->d_part_events / d_part_orig, run the slab kernel on the
// copy and translate the first indices back.  *ok = false: shape not supported (too many bins for
// shared memory, too few bins, slab kernel not applicable) -> the caller takes the table path.
// A stream whose strips each touch many bins (fully shuffled) is partitioned in two levels (groups
// of ~sqrt(bins) bins first): a one-level scatter into thousands of open segments per strip turns
// every 16-B record into a DRAM read-modify-write.
int evk_downsample_partitioned(evk_handle* h, const KeyParams& kp, int count_repeated, bool* ok,
                               int* launches) {
    *ok = false;
    const size_t n = h->n_events;
    if (n == 0 || n >= 0xFFFFFFFFull || !evk_slab_supported(h, kp)) return EVK_OK;
    unsigned long long* d_stats = h->d_cnt->scratch;  // [0..3] (the slab kernels use them later)
    const unsigned long long init[4] = {~0ull, 0ull, 0ull, 0ull};
    EVK_CUDA(h, cudaMemcpyAsync(d_stats, init, sizeof init, cudaMemcpyHostToDevice, h->stream));
    k_part_minmax<<<h->sm_count * 8, kBlock, 0, h->stream>>>(kp, h->d_events, n, d_stats);
    EVK_CUDA(h, cudaGetLastError());
    EVK_CUDA(h, cudaMemcpyAsync(h->h_cnt->scratch, d_stats, 3 * sizeof(unsigned long long),
                                cudaMemcpyDeviceToHost, h->stream));
    EVK_CUDA(h, cudaStreamSynchronize(h->stream));
    (*launches)++;
    const uint64_t tb_min = h->h_cnt->scratch[0], tb_max = h->h_cnt->scratch[1];
    const size_t n_valid = (size_t)h->h_cnt->scratch[2];
    if (n_valid == 0) return EVK_OK;  // nothing gated in: the table path reports the empty result
    const uint64_t nb64 = tb_max - tb_min + 1;
    // one u32 cursor per bin and warp in shared memory; at least 4 warps per CTA
    constexpr size_t kSmem = 200 * 1024;
    if (nb64 > kSmem / (4 * 4)) return EVK_OK;
    const uint32_t nb = (uint32_t)nb64;
    const uint32_t n_strips = (uint32_t)((n + kStrip - 1) / kStrip);
    // buffers (lazy, grow-only): partitioned events, original indices, strip x bin offsets
    if (!h->d_part_events) {
        EVK_CUDA(h, cudaMalloc((void**)&h->d_part_events, h->max_events * sizeof(evk_event)));
        EVK_CUDA(h, cudaMalloc((void**)&h->d_part_orig, h->max_events * sizeof(uint32_t)));
    }
    const size_t need = (size_t)n_strips * nb + nb + 2 +
                        (size_t)((n_strips + kChunkStrips - 1) / kChunkStrips) * nb;
    if (h->part_hist_cap < need) {
        if (h->d_part_hist) cudaFree(h->d_part_hist);
        h->d_part_hist = nullptr;
        h->part_hist_cap = 0;
        if (cudaMalloc((void**)&h->d_part_hist, need * sizeof(uint32_t)) != cudaSuccess) {
            cudaGetLastError();
            return EVK_OK;  // too many strips x bins: table path
        }
        h->part_hist_cap = need;
    }
    EVK_CUDA(h, cudaFuncSetAttribute(k_part_hist, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)kSmem));
    EVK_CUDA(h, cudaFuncSetAttribute(k_part_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)kSmem));
    PartArgs a;
    a.kp = kp;
    a.ev = h->d_events;
    a.n = n;
    a.tb_min = tb_min;
    a.nb = nb;
    a.div = 1;
    a.n_strips = n_strips;
    a.spread = &d_stats[3];
    // histogram over the bins first: it also tells how many bins a strip touches on average
    EVK_TRY(partition_pass(h, a, nullptr, nullptr, nullptr, false, launches));
    EVK_CUDA(h, cudaMemcpyAsync(&h->h_cnt->scratch[3], &d_stats[3], sizeof(unsigned long long),
                                cudaMemcpyDeviceToHost, h->stream));
    EVK_CUDA(h, cudaStreamSynchronize(h->stream));
    const double spread = (double)h->h_cnt->scratch[3] / (double)n_strips;
    a.spread = nullptr;
    if (spread <= 128.0 || nb < 256) {
        EVK_TRY(partition_pass(h, a, nullptr, h->d_part_events, h->d_part_orig, true, launches));
    } else {
        // level 1: groups of `div` bins; level 2: the bins, on the grouped copy
        if (!h->d_part_events2) {
            EVK_CUDA(h, cudaMalloc((void**)&h->d_part_events2, h->max_events * sizeof(evk_event)));
            EVK_CUDA(h, cudaMalloc((void**)&h->d_part_orig2, h->max_events * sizeof(uint32_t)));
        }
        uint32_t div = 1;
        while ((uint64_t)div * div < nb) div++;
        PartArgs a1 = a;
        a1.div = div;
        a1.nb = (nb + div - 1) / div;
        EVK_TRY(partition_pass(h, a1, nullptr, h->d_part_events2, h->d_part_orig2, false, launches));
        PartArgs a2 = a;
        a2.ev = h->d_part_events2;
        a2.n = n_valid;
        a2.n_strips = (uint32_t)((n_valid + kStrip - 1) / kStrip);
        EVK_TRY(partition_pass(h, a2, h->d_part_orig2, h->d_part_events, h->d_part_orig, false,
                               launches));
    }
    // the slab kernel on the partitioned copy (every event of it is gated in and in bin order)
    evk_event* ev_saved = h->d_events;
    h->d_events = h->d_part_events;
    h->n_events = n_valid;
    EVK_CUDA(h, cudaMemsetAsync(h->d_cnt, 0, sizeof(DsCounters), h->stream));
    bool slab_ok = false;
    const int st = evk_downsample_slab(h, kp, count_repeated, &slab_ok, launches, true);
    h->d_events = ev_saved;
    h->n_events = n;
    EVK_TRY(st);
    if (!slab_ok) return EVK_OK;
    k_part_translate<<<h->sm_count * 8, kBlock, 0, h->stream>>>(
        h->d_first, &h->d_cnt->n_unique, h->d_part_orig, (uint32_t)h->shard_first);
    EVK_CUDA(h, cudaGetLastError());
    (*launches)++;
    *ok = true;
    return EVK_OK;
}

void evk_partition_free(evk_handle* h) {
    if (h->d_part_events) cudaFree(h->d_part_events);
    if (h->d_part_orig) cudaFree(h->d_part_orig);
    if (h->d_part_hist) cudaFree(h->d_part_hist);
    if (h->d_part_events2) cudaFree(h->d_part_events2);
    if (h->d_part_orig2) cudaFree(h->d_part_orig2);
    h->d_part_events2 = nullptr;
    h->d_part_orig2 = nullptr;
    h->d_part_events = nullptr;
    h->d_part_orig = nullptr;
    h->d_part_hist = nullptr;
    h->part_hist_cap = 0;
}
