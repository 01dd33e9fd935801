// evk_evt2.cu — RAW EVT 2.0 ingest: the sensor's own 4-byte words go over PCIe and are decoded to
// the 16-byte packed records on the device.
//
// The reference reads recordings with Metavision::Camera::from_file(argv[1]) (ACCEL/store.cpp:336,
// `traffic_data.raw` in ACCEL/Readme.md:21): the SDK decodes the RAW payload on the CPU and hands
// EventCD ranges to the callback (:614-615).  Here the payload is uploaded as it is (4 B per event
// instead of 16) and decoded by three small kernels, so the end-to-end path moves a quarter of the
// bytes.  Format (Prophesee "EVT 2.0", word layout as in include/evk.h): type in bits 31..28;
// CD_OFF 0x0 / CD_ON 0x1: [27:22] t bits 5..0, [21:11] x, [10:0] y; EVT_TIME_HIGH 0x8: [27:0] t
// bits 33..6; every other type carries no CD event.
//
// Decode = an order-preserving stream compaction (CD words only) + a "last EVT_TIME_HIGH seen"
// scan.  k_evt2_scan: per 2048-word block, CD count and last time-high.  k_evt2_prefix: one CTA
// scans the block summaries (exclusive CD offsets, carried time-high).  k_evt2_decode: per block
// the same local scans again, then every thread walks its 8 consecutive words and writes its CD
// events into a shared-memory tile, which the block then copies out with fully coalesced 16-B
// stores.  HBM traffic: 8 B read + 16 B written per event.
#include "evk_internal.cuh"

namespace {

constexpr int kT = 256;              // threads per block
constexpr int kWpt = 8;              // consecutive words per thread (two 16-B loads)
constexpr int kWpb = kT * kWpt;      // words per block
constexpr uint32_t kNoTh = 0xFFFFFFFFu;

__device__ __forceinline__ void load16(const uint32_t* __restrict__ words, size_t n_words,
                                       size_t at, uint32_t (&w)[kWpt]) {
    if (at + kWpt <= n_words) {
#pragma unroll
        for (int q = 0; q < kWpt / 4; q++) {
            const uint4 v = __ldcs(reinterpret_cast<const uint4*>(words + at) + q);
            w[4 * q + 0] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int q = 0; q < kWpt; q++) w[q] = at + q < n_words ? words[at + q] : 0xE0000000u;  // OTHERS
    }
}

// per thread: number of CD words and the last time-high among its words (kNoTh: none)
__device__ __forceinline__ void summarise(const uint32_t (&w)[kWpt], uint32_t& cd, uint32_t& th) {
    cd = 0;
    th = kNoTh;
#pragma unroll
    for (int q = 0; q < kWpt; q++) {
        const uint32_t type = w[q] >> 28;
        cd += type <= 1u;
        if (type == 8u) th = w[q] & 0x0FFFFFFFu;
    }
}

// block-wide exclusive scan of (cd, th): cd adds up, th = the right-most one that exists.
// Returns the exclusive prefix for this thread; *tot_* = block totals.
__device__ __forceinline__ void block_scan(uint32_t cd, uint32_t th, uint32_t& cd_ex,
                                           uint32_t& th_ex, uint32_t& tot_cd, uint32_t& tot_th) {
    __shared__ uint32_t s_cd[kT / 32], s_th[kT / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t icd = cd, ith = th;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t c = __shfl_up_sync(0xffffffffu, icd, o);
        const uint32_t t = __shfl_up_sync(0xffffffffu, ith, o);
        if (lane >= o) {
            icd += c;
            if (ith == kNoTh) ith = t;
        }
    }
    if (lane == 31) {
        s_cd[wid] = icd;
        s_th[wid] = ith;
    }
    __syncthreads();
    uint32_t base_cd = 0, base_th = kNoTh;
    tot_cd = 0;
    tot_th = kNoTh;
#pragma unroll
    for (int k = 0; k < kT / 32; k++) {
        if (k < wid) {
            base_cd += s_cd[k];
            if (s_th[k] != kNoTh) base_th = s_th[k];
        }
        tot_cd += s_cd[k];
        if (s_th[k] != kNoTh) tot_th = s_th[k];
    }
    // exclusive within the warp
    const uint32_t pcd = __shfl_up_sync(0xffffffffu, icd, 1);
    const uint32_t pth = __shfl_up_sync(0xffffffffu, ith, 1);
    cd_ex = base_cd + (lane ? pcd : 0u);
    th_ex = lane && pth != kNoTh ? pth : base_th;
    __syncthreads();
}

__global__ void __launch_bounds__(kT)
    k_evt2_scan(const uint32_t* __restrict__ words, size_t n_words, uint32_t* blk_cd,
                uint32_t* blk_th) {
    uint32_t w[kWpt], cd, th, cd_ex, th_ex, tot_cd, tot_th;
    load16(words, n_words, (size_t)blockIdx.x * kWpb + (size_t)threadIdx.x * kWpt, w);
    summarise(w, cd, th);
    block_scan(cd, th, cd_ex, th_ex, tot_cd, tot_th);
    if (threadIdx.x == 0) {
        blk_cd[blockIdx.x] = tot_cd;
        blk_th[blockIdx.x] = tot_th;
    }
}

// one CTA: exclusive scan of the block summaries, in place (blk_cd -> offsets, blk_th -> carry-in)
__global__ void __launch_bounds__(1024)
    k_evt2_prefix(uint32_t* blk_cd, uint32_t* blk_th, uint32_t n_blocks,
                  unsigned long long* total) {
    __shared__ unsigned long long s_sum[1024];
    __shared__ uint32_t s_last[1024];
    const uint32_t per = (n_blocks + 1023) / 1024;
    const uint32_t b0 = threadIdx.x * per, b1 = min(n_blocks, b0 + per);
    unsigned long long sum = 0;
    uint32_t last = kNoTh;
    for (uint32_t b = b0; b < b1; b++) {
        sum += blk_cd[b];
        if (blk_th[b] != kNoTh) last = blk_th[b];
    }
    s_sum[threadIdx.x] = sum;
    s_last[threadIdx.x] = last;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan
        unsigned long long a = 0;
        uint32_t l = kNoTh;
        if ((int)threadIdx.x >= o) {
            a = s_sum[threadIdx.x - o];
            l = s_last[threadIdx.x - o];
        }
        __syncthreads();
        if ((int)threadIdx.x >= o) {
            s_sum[threadIdx.x] += a;
            if (s_last[threadIdx.x] == kNoTh) s_last[threadIdx.x] = l;
        }
        __syncthreads();
    }
    unsigned long long off = threadIdx.x ? s_sum[threadIdx.x - 1] : 0ull;
    uint32_t carry = threadIdx.x ? s_last[threadIdx.x - 1] : kNoTh;
    for (uint32_t b = b0; b < b1; b++) {
        const uint32_t c = blk_cd[b], t = blk_th[b];
        blk_cd[b] = (uint32_t)off;  // < 2^32: bounded by the handle capacity (checked by the host)
        blk_th[b] = carry == kNoTh ? 0u : carry;  // before the first EVT_TIME_HIGH: 0
        off += c;
        if (t != kNoTh) carry = t;
    }
    if (threadIdx.x == 1023) *total = s_sum[1023];
}

__global__ void __launch_bounds__(kT)
    k_evt2_decode(const uint32_t* __restrict__ words, size_t n_words,
                  const uint32_t* __restrict__ blk_off, const uint32_t* __restrict__ blk_carry,
                  evk_event* __restrict__ out, size_t cap) {
    uint32_t w[kWpt], cd, th, cd_ex, th_ex, tot_cd, tot_th;
    load16(words, n_words, (size_t)blockIdx.x * kWpb + (size_t)threadIdx.x * kWpt, w);
    summarise(w, cd, th);
    block_scan(cd, th, cd_ex, th_ex, tot_cd, tot_th);
    // stage the block's events in shared memory in stream order, then copy out coalesced
    __shared__ uint4 s_ev[kWpb];
    uint32_t o = cd_ex;
    uint64_t time_high = th_ex != kNoTh ? th_ex : blk_carry[blockIdx.x];
#pragma unroll
    for (int q = 0; q < kWpt; q++) {
        const uint32_t type = w[q] >> 28;
        if (type == 8u) time_high = w[q] & 0x0FFFFFFFu;
        if (type <= 1u) {
            const uint64_t t = (time_high << 6) | ((w[q] >> 22) & 0x3Fu);
            s_ev[o++] = make_uint4(((w[q] >> 11) & 0x7FFu) | ((w[q] & 0x7FFu) << 16), type,
                                   (uint32_t)t, (uint32_t)(t >> 32));
        }
    }
    __syncthreads();
    const size_t base = blk_off[blockIdx.x];
    uint4* dst = reinterpret_cast<uint4*>(out);
    for (uint32_t i = threadIdx.x; i < tot_cd; i += kT)
        if (base + i < cap) __stcs(dst + base + i, s_ev[i]);
}

}  // namespace

// words: device pointer.  blk: scratch of 2 * n_blocks u32.  total: device scalar (CD events).
cudaError_t evk_launch_evt2_decode(const uint32_t* words, size_t n_words, uint32_t* blk,
                                   unsigned long long* total, evk_event* out, size_t cap,
                                   cudaStream_t s) {
    const uint32_t n_blocks = (uint32_t)((n_words + kWpb - 1) / kWpb);
    k_evt2_scan<<<n_blocks, kT, 0, s>>>(words, n_words, blk, blk + n_blocks);
    k_evt2_prefix<<<1, 1024, 0, s>>>(blk, blk + n_blocks, n_blocks, total);
    k_evt2_decode<<<n_blocks, kT, 0, s>>>(words, n_words, blk, blk + n_blocks, out, cap);
    return cudaGetLastError();
}
size_t evk_evt2_blocks(size_t n_words) { return (n_words + kWpb - 1) / kWpb; }
