// evk_downsample.cu — voxel-hash downsample kernels (general paths) for sm_100a.
//
// Replaces the reference's single-work-group process_coordinates kernel
// (ACCEL/build/coordinate_processor.cl:16-89: 8192 lossy buckets in __local memory, zeroed by one
// work-item, no collision resolution, first-arrival representative) by
//   TABLE : a global open-addressing table on 64-bit voxel keys — atomicCAS claims a slot,
//           atomicMin keeps the lowest stream index (deterministic representative),
//           bit 63 of the stored key records "hit at least twice" (the kernel's repeated_count,
//           :73-75); a block-scan stream compaction emits the SoA voxel shard and resets the
//           slots it visits so the table is clean for the next call;
//   SORT  : radix sort of (key, index) pairs + head detection — the cross-check variant.
// The time-slab fast path lives in evk_slab.cu.
#include <cub/device/device_radix_sort.cuh>

#include "evk_internal.cuh"

namespace {

constexpr int kBlock = 256;

__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

// exclusive scan of one int per thread across a kBlock-thread block; returns the block total
__device__ __forceinline__ int block_excl_scan(int v, int& total, int* s_warp /*[kBlock/32+1]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < kBlock / 32 ? s_warp[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < kBlock / 32; o <<= 1) {
            int n = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += n;
        }
        if (lane < kBlock / 32) s_warp[lane] = winc - w;
        if (lane == kBlock / 32 - 1) s_warp[kBlock / 32] = winc;
    }
    __syncthreads();
    total = s_warp[kBlock / 32];
    int r = s_warp[warp] + inc - v;
    __syncthreads();
    return r;
}

// ---------------------------------------------------------------------------- table ----------
__global__ void __launch_bounds__(kBlock) k_table_clear(uint64_t* tkeys, uint32_t* tfirst,
                                                        size_t cap) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += stride) {
        tkeys[i] = EVK_EMPTY_KEY;
        tfirst[i] = EVK_EMPTY_IDX;
    }
}

__device__ __forceinline__ void table_put(uint64_t key, uint32_t idx, uint64_t* tkeys,
                                          uint32_t* tfirst, uint64_t mask, int count_repeated) {
    uint64_t slot = evk_mix64(key) & mask;
    for (;;) {
        uint64_t cur = ld_relaxed_u64(tkeys + slot);
        if (cur == EVK_EMPTY_KEY) {
            cur = atomicCAS((unsigned long long*)(tkeys + slot), EVK_EMPTY_KEY, key);
            if (cur == EVK_EMPTY_KEY) {
                atomicMin(tfirst + slot, idx);
                return;
            }
        }
        if ((cur & ~EVK_REP_FLAG) == key) {
            atomicMin(tfirst + slot, idx);
            if (count_repeated && !(cur & EVK_REP_FLAG))
                atomicOr((unsigned long long*)(tkeys + slot), EVK_REP_FLAG);
            return;
        }
        slot = (slot + 1) & mask;
    }
}

constexpr int kInsUnroll = 4;
__global__ void __launch_bounds__(kBlock)
    k_table_insert(KeyParams kp, const evk_event* __restrict__ ev, size_t n, uint64_t* tkeys,
                   uint32_t* tfirst, uint64_t mask, int count_repeated) {
    const size_t tile = (size_t)kBlock * kInsUnroll;
    for (size_t base = (size_t)blockIdx.x * tile; base < n; base += (size_t)gridDim.x * tile) {
        uint4 e[kInsUnroll];
#pragma unroll
        for (int j = 0; j < kInsUnroll; j++) {
            size_t i = base + (size_t)j * kBlock + threadIdx.x;
            if (i < n) e[j] = ld_event(ev + i);
        }
#pragma unroll
        for (int j = 0; j < kInsUnroll; j++) {
            size_t i = base + (size_t)j * kBlock + threadIdx.x;
            uint64_t key;
            if (i < n && evk_key(kp, e[j], key))
                table_put(key, (uint32_t)i, tkeys, tfirst, mask, count_repeated);
        }
    }
}

constexpr int kCmpPer = 8;
__global__ void __launch_bounds__(kBlock)
    k_table_compact(const evk_event* __restrict__ ev, uint64_t* tkeys, uint32_t* tfirst,
                    size_t cap, uint64_t* __restrict__ keys, uint32_t* __restrict__ first,
                    uint32_t* __restrict__ xy, uint32_t first_offset, DsCounters* cnt) {
    __shared__ int s_warp[kBlock / 32 + 1];
    __shared__ unsigned long long s_base;
    const size_t tile = (size_t)kBlock * kCmpPer;
    for (size_t base = (size_t)blockIdx.x * tile; base < cap; base += (size_t)gridDim.x * tile) {
        uint64_t k[kCmpPer];
        int mine = 0, rep = 0;
#pragma unroll
        for (int j = 0; j < kCmpPer; j++) {
            size_t s = base + (size_t)j * kBlock + threadIdx.x;
            k[j] = s < cap ? tkeys[s] : EVK_EMPTY_KEY;
            if (k[j] != EVK_EMPTY_KEY) {
                mine++;
                rep += (k[j] & EVK_REP_FLAG) ? 1 : 0;
            }
        }
        int total;
        int off = block_excl_scan(mine, total, s_warp);
        // piggy-back the repeated count on a second scan-free reduction
        rep = __reduce_add_sync(0xffffffffu, rep);
        if (total == 0) continue;  // uniform across the block
        if ((threadIdx.x & 31) == 0 && rep) atomicAdd(&cnt->n_repeated, (unsigned long long)rep);
        if (threadIdx.x == 0) s_base = atomicAdd(&cnt->n_unique, (unsigned long long)total);
        __syncthreads();
        size_t o = (size_t)s_base + off;
#pragma unroll
        for (int j = 0; j < kCmpPer; j++) {
            if (k[j] != EVK_EMPTY_KEY) {
                size_t s = base + (size_t)j * kBlock + threadIdx.x;
                uint32_t f = tfirst[s];
                keys[o] = k[j] & ~EVK_REP_FLAG;
                first[o] = f + first_offset;
                xy[o] = *reinterpret_cast<const uint32_t*>(ev + f);  // x | y << 16
                o++;
                tkeys[s] = EVK_EMPTY_KEY;  // leave the table clean for the next call
                tfirst[s] = EVK_EMPTY_IDX;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------- sort ------------
__global__ void __launch_bounds__(kBlock)
    k_make_pairs(KeyParams kp, const evk_event* __restrict__ ev, size_t n, uint64_t* keys,
                 uint32_t* idx) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint4 e = ld_event(ev + i);
        uint64_t key;
        keys[i] = evk_key(kp, e, key) ? key : EVK_EMPTY_KEY;  // gated events sort last
        idx[i] = (uint32_t)i;
    }
}

// sorted (key, idx) pairs -> one record per run of equal keys; the sort is stable and idx was
// ascending, so the head of a run carries the lowest stream index
__global__ void __launch_bounds__(kBlock)
    k_unique_heads(const evk_event* __restrict__ ev, const uint64_t* __restrict__ sk,
                   const uint32_t* __restrict__ si, size_t n, uint64_t* __restrict__ keys,
                   uint32_t* __restrict__ first, uint32_t* __restrict__ xy, uint32_t first_offset,
                   DsCounters* cnt) {
    __shared__ int s_warp[kBlock / 32 + 1];
    __shared__ unsigned long long s_base;
    for (size_t base = (size_t)blockIdx.x * kBlock; base < n; base += (size_t)gridDim.x * kBlock) {
        size_t j = base + threadIdx.x;
        uint64_t k = j < n ? sk[j] : EVK_EMPTY_KEY;
        bool head = k != EVK_EMPTY_KEY && (j == 0 || sk[j - 1] != k);
        int rep = head && j + 1 < n && sk[j + 1] == k;
        int total;
        int off = block_excl_scan(head ? 1 : 0, total, s_warp);
        rep = __reduce_add_sync(0xffffffffu, rep);
        if (total == 0) continue;
        if ((threadIdx.x & 31) == 0 && rep) atomicAdd(&cnt->n_repeated, (unsigned long long)rep);
        if (threadIdx.x == 0) s_base = atomicAdd(&cnt->n_unique, (unsigned long long)total);
        __syncthreads();
        if (head) {
            size_t o = (size_t)s_base + off;
            uint32_t f = si[j];
            keys[o] = k;
            first[o] = f + first_offset;
            xy[o] = *reinterpret_cast<const uint32_t*>(ev + f);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kBlock) k_iota(uint32_t* a, size_t n) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        a[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(kBlock)
    k_gather_voxels(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ first,
                    const evk_event* __restrict__ ev, const evk_event* __restrict__ reps,
                    uint64_t first_offset, const uint32_t* __restrict__ perm, size_t n,
                    uint64_t* okeys, evk_event* oreps, uint32_t* ofirst) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t p = perm[i];
        uint32_t f = first[p];
        if (okeys) okeys[i] = keys[p];
        if (ofirst) ofirst[i] = f;
        if (oreps) oreps[i] = reps ? reps[p] : ev[(uint64_t)f - first_offset];
    }
}

__global__ void __launch_bounds__(kBlock)
    k_gather_labels(const int32_t* __restrict__ labels, const uint32_t* __restrict__ perm,
                    int32_t* out, size_t n) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = labels[perm[i]];
}

__global__ void __launch_bounds__(kBlock) k_fill_u8(uint4* p, uint32_t v, size_t n16) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    uint4 w = make_uint4(v, v, v, v);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) p[i] = w;
}

inline int grid_for(size_t items, int per_block, int sm_count, int waves = 8) {
    size_t need = (items + per_block - 1) / per_block;
    size_t cap = (size_t)sm_count * waves;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

}  // namespace

cudaError_t evk_launch_fill_u8(void* p, int v, size_t bytes, cudaStream_t s) {
    uint32_t b = (uint32_t)(v & 0xFF);
    b |= b << 8;
    b |= b << 16;
    size_t n16 = bytes / 16;
    if (n16) k_fill_u8<<<148 * 8, kBlock, 0, s>>>((uint4*)p, b, n16);
    if (bytes % 16) return cudaMemsetAsync((char*)p + n16 * 16, v, bytes % 16, s);
    return cudaGetLastError();
}

cudaError_t evk_launch_table_clear(uint64_t* tkeys, uint32_t* tfirst, size_t cap,
                                   cudaStream_t s) {
    k_table_clear<<<148 * 8, kBlock, 0, s>>>(tkeys, tfirst, cap);
    return cudaGetLastError();
}

cudaError_t evk_launch_table_insert(const KeyParams& kp, const evk_event* ev, size_t n,
                                    uint64_t* tkeys, uint32_t* tfirst, size_t cap,
                                    int count_repeated, int sm_count, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    int grid = grid_for(n, kBlock * kInsUnroll, sm_count);
    k_table_insert<<<grid, kBlock, 0, s>>>(kp, ev, n, tkeys, tfirst, (uint64_t)cap - 1,
                                           count_repeated);
    return cudaGetLastError();
}

cudaError_t evk_launch_table_compact(const evk_event* ev, uint64_t* tkeys, uint32_t* tfirst,
                                     size_t cap, uint64_t* keys, uint32_t* first, uint32_t* xy,
                                     uint32_t first_offset, DsCounters* cnt, int sm_count,
                                     cudaStream_t s) {
    int grid = grid_for(cap, kBlock * kCmpPer, sm_count);
    k_table_compact<<<grid, kBlock, 0, s>>>(ev, tkeys, tfirst, cap, keys, first, xy, first_offset,
                                            cnt);
    return cudaGetLastError();
}

int evk_downsample_sort(evk_handle* h, const KeyParams& kp, int* launches) {
    const size_t n = h->n_events;
    if (n == 0) return EVK_OK;
    if (!h->d_sk_in) {
        const size_t m = h->max_events;
        EVK_CUDA(h, cudaMalloc(&h->d_sk_in, m * sizeof(uint64_t)));
        EVK_CUDA(h, cudaMalloc(&h->d_sk_out, m * sizeof(uint64_t)));
        EVK_CUDA(h, cudaMalloc(&h->d_si_in, m * sizeof(uint32_t)));
        EVK_CUDA(h, cudaMalloc(&h->d_si_out, m * sizeof(uint32_t)));
        size_t bytes = 0;
        EVK_CUDA(h, cub::DeviceRadixSort::SortPairs(nullptr, bytes, h->d_sk_in, h->d_sk_out,
                                                    h->d_si_in, h->d_si_out, (int64_t)m, 0, 64,
                                                    h->stream));
        EVK_CUDA(h, cudaMalloc(&h->d_sv_tmp, bytes));
        h->sv_tmp_bytes = bytes;
    }
    int grid = grid_for(n, kBlock, h->sm_count);
    k_make_pairs<<<grid, kBlock, 0, h->stream>>>(kp, h->d_events, n, h->d_sk_in, h->d_si_in);
    EVK_CUDA(h, cudaGetLastError());
    size_t bytes = h->sv_tmp_bytes;
    EVK_CUDA(h, cub::DeviceRadixSort::SortPairs(h->d_sv_tmp, bytes, h->d_sk_in, h->d_sk_out,
                                                h->d_si_in, h->d_si_out, (int64_t)n, 0, 64,
                                                h->stream));
    k_unique_heads<<<grid, kBlock, 0, h->stream>>>(h->d_events, h->d_sk_out, h->d_si_out, n,
                                                   h->d_keys, h->d_first, h->d_xy,
                                                   (uint32_t)h->shard_first, h->d_cnt);
    EVK_CUDA(h, cudaGetLastError());
    *launches += 2 + 8;  // key build, head compaction, ~8 radix passes (library kernels)
    return EVK_OK;
}

// ---- canonical order ------------------------------------------------------------------------
// perm[i] = emission position of the voxel with the i-th lowest first stream index.  First indices
// are distinct event indices inside the handle's shard, so the rank of a voxel is a population
// count: one bit per event ("is a representative"), an exclusive prefix of the word popcounts, and
// rank(f) = prefix[f / 32] + popc(word & below(f)).  Three light passes over the voxel list and a
// 1-bit-per-event map instead of a radix sort of the first indices (4 passes over 8-byte pairs).
namespace {
constexpr int kScanWords = 16;  // flag words per thread in the prefix passes

__global__ void __launch_bounds__(kBlock)
    k_perm_flag(const uint32_t* __restrict__ first, size_t n, uint32_t base, uint32_t* flags) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t f = first[i] - base;
        atomicOr(&flags[f >> 5], 1u << (f & 31));
    }
}

// pass 1: popcount of each block of kBlock * kScanWords flag words
__global__ void __launch_bounds__(kBlock)
    k_perm_blocksum(const uint32_t* __restrict__ flags, size_t words, uint32_t* blk) {
    __shared__ int s_warp[kBlock / 32 + 1];
    const size_t w0 = ((size_t)blockIdx.x * kBlock + threadIdx.x) * kScanWords;
    int c = 0;
#pragma unroll
    for (int j = 0; j < kScanWords; j++)
        if (w0 + j < words) c += __popc(flags[w0 + j]);
    int total;
    block_excl_scan(c, total, s_warp);
    if (threadIdx.x == 0) blk[blockIdx.x] = (uint32_t)total;
}

// pass 2: exclusive scan of the block sums, one CTA, carried across chunks of kBlock
__global__ void __launch_bounds__(kBlock) k_perm_blockscan(uint32_t* blk, size_t n_blocks) {
    __shared__ int s_warp[kBlock / 32 + 1];
    uint32_t carry = 0;
    for (size_t b0 = 0; b0 < n_blocks; b0 += kBlock) {
        const size_t b = b0 + threadIdx.x;
        const int v = b < n_blocks ? (int)blk[b] : 0;
        int total;
        const int off = block_excl_scan(v, total, s_warp);
        if (b < n_blocks) blk[b] = carry + (uint32_t)off;
        carry += (uint32_t)total;
    }
}

// pass 3: exclusive prefix per flag word
__global__ void __launch_bounds__(kBlock)
    k_perm_wordprefix(const uint32_t* __restrict__ flags, size_t words,
                      const uint32_t* __restrict__ blk, uint32_t* prefix) {
    __shared__ int s_warp[kBlock / 32 + 1];
    const size_t w0 = ((size_t)blockIdx.x * kBlock + threadIdx.x) * kScanWords;
    uint32_t v[kScanWords];
    int c = 0;
#pragma unroll
    for (int j = 0; j < kScanWords; j++) {
        v[j] = w0 + j < words ? flags[w0 + j] : 0u;
        c += __popc(v[j]);
    }
    int total;
    uint32_t run = blk[blockIdx.x] + (uint32_t)block_excl_scan(c, total, s_warp);
#pragma unroll
    for (int j = 0; j < kScanWords; j++) {
        if (w0 + j < words) prefix[w0 + j] = run;
        run += __popc(v[j]);
    }
}

__global__ void __launch_bounds__(kBlock)
    k_perm_scatter(const uint32_t* __restrict__ first, size_t n, uint32_t base,
                   const uint32_t* __restrict__ flags, const uint32_t* __restrict__ prefix,
                   uint32_t* perm) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t f = first[i] - base;
        const uint32_t w = f >> 5;
        const uint32_t rank = prefix[w] + __popc(flags[w] & ((1u << (f & 31)) - 1u));
        perm[rank] = (uint32_t)i;
    }
}
}  // namespace

int evk_ensure_perm(evk_handle* h) {
    if (h->perm_valid) return EVK_OK;
    const size_t n = h->n_unique;
    const size_t m = h->max_events;
    if (!h->d_perm) {
        EVK_CUDA(h, cudaMalloc(&h->d_perm, m * sizeof(uint32_t)));
        EVK_CUDA(h, cudaMalloc(&h->d_sort_a, (m + 64) * sizeof(uint32_t)));
        EVK_CUDA(h, cudaMalloc(&h->d_sort_b, (m + 64) * sizeof(uint32_t)));
    }
    if (n && !h->voxels_foreign) {
        // local voxels: first indices lie in [shard_first, shard_first + max_events)
        const uint32_t base = (uint32_t)h->shard_first;
        const size_t words = (m + 31) / 32 + 1;
        const size_t per_block = (size_t)kBlock * kScanWords;
        const size_t n_blocks = (words + per_block - 1) / per_block;
        uint32_t* flags = h->d_sort_a;            // [words]
        uint32_t* prefix = h->d_sort_b;           // [words]
        uint32_t* blk = h->d_sort_a + words + 4;  // [n_blocks]  (words + n_blocks << m)
        if (words + 4 + n_blocks > m + 64) return evk_fail(h, EVK_ERR_CAPACITY, "perm scratch");
        EVK_CUDA(h, cudaMemsetAsync(flags, 0, words * sizeof(uint32_t), h->stream));
        const int grid = grid_for(n, kBlock, h->sm_count);
        k_perm_flag<<<grid, kBlock, 0, h->stream>>>(h->d_first, n, base, flags);
        k_perm_blocksum<<<(unsigned)n_blocks, kBlock, 0, h->stream>>>(flags, words, blk);
        k_perm_blockscan<<<1, kBlock, 0, h->stream>>>(blk, n_blocks);
        k_perm_wordprefix<<<(unsigned)n_blocks, kBlock, 0, h->stream>>>(flags, words, blk, prefix);
        k_perm_scatter<<<grid, kBlock, 0, h->stream>>>(h->d_first, n, base, flags, prefix, h->d_perm);
        EVK_CUDA(h, cudaGetLastError());
    } else if (n) {
        // voxels received from peers (hash ownership): first indices span the whole global stream;
        // radix sort of (first index, position) pairs -- library code, off the hot path
        if (!h->d_sort_tmp) {
            size_t bytes = 0;
            EVK_CUDA(h, cub::DeviceRadixSort::SortPairs(nullptr, bytes, h->d_first, h->d_sort_b,
                                                        h->d_sort_a, h->d_perm, (int64_t)m, 0, 32,
                                                        h->stream));
            EVK_CUDA(h, cudaMalloc(&h->d_sort_tmp, bytes));
            h->sort_tmp_bytes = bytes;
        }
        k_iota<<<grid_for(n, kBlock, h->sm_count), kBlock, 0, h->stream>>>(h->d_sort_a, n);
        EVK_CUDA(h, cudaGetLastError());
        size_t bytes = h->sort_tmp_bytes;
        EVK_CUDA(h, cub::DeviceRadixSort::SortPairs(h->d_sort_tmp, bytes, h->d_first, h->d_sort_b,
                                                    h->d_sort_a, h->d_perm, (int64_t)n, 0, 32,
                                                    h->stream));
    }
    h->perm_valid = true;
    return EVK_OK;
}

cudaError_t evk_launch_gather_voxels(const evk_handle* h, uint64_t* keys, evk_event* reps,
                                     uint32_t* first, size_t n) {
    if (n == 0) return cudaSuccess;
    k_gather_voxels<<<grid_for(n, kBlock, h->sm_count), kBlock, 0, h->stream>>>(
        h->d_keys, h->d_first, h->d_events, h->reps_valid ? h->d_reps : nullptr, h->shard_first,
        h->d_perm, n, keys, reps, first);
    return cudaGetLastError();
}

cudaError_t evk_launch_gather_labels(const int32_t* labels, const uint32_t* perm, int32_t* out,
                                     size_t n, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    k_gather_labels<<<grid_for(n, kBlock, 148), kBlock, 0, s>>>(labels, perm, out, n);
    return cudaGetLastError();
}
