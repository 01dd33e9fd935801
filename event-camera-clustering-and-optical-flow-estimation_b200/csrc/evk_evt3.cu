// evk_evt3.cu — RAW EVT 3.0 ingest: 16-bit sensor words over PCIe, decoded on the device.
//
// The reference replays recordings through Metavision::Camera::from_file (ACCEL/store.cpp:336); the
// SDK decodes the payload on the CPU.  EVT 3.0 is the format current Prophesee sensors (Gen4 / IMX636)
// write: a stateful stream of 16-bit words, type in bits 15..12 --
//   0x0 EVT_ADDR_Y    [10:0] y                      -> sets the row of the following events
//   0x2 EVT_ADDR_X    [11] polarity, [10:0] x       -> ONE event (x, row, polarity, time)
//   0x3 VECT_BASE_X   [11] polarity, [10:0] x       -> base column + polarity of the vectors below
//   0x4 VECT_12       [11:0] mask                   -> an event at base + i per set bit; base += 12
//   0x5 VECT_8        [7:0] mask                    -> likewise, 8 columns; base += 8
//   0x6 EVT_TIME_LOW  [11:0] time bits 11..0
//   0x8 EVT_TIME_HIGH [11:0] time bits 23..12 (a value lower than the previous one = a 2^24 us wrap)
//   others (EXT_TRIGGER 0xA, OTHERS 0xE, CONTINUED 0x7 / 0xF): no CD event
// so a word's meaning depends on the last row, the last two time words, the wraps so far and the
// running vector base.  Every one of those is an associative "last writer / running sum" state, so
// the decode is a scan: the effect of a run of words on the decoder is a small record (E3 below)
// with an associative combine.  Three kernels, as for EVT 2.0 (evk_evt2.cu):
//   k_evt3_scan    per 4096-word block: every thread folds its 16 words, block scan -> block effect
//   k_evt3_prefix  one CTA: exclusive scan of the block effects (state at the start of each block)
//   k_evt3_decode  per block: the same local scan, then every thread replays its 16 words from its
//                  exact entry state and writes its events at its exact offset (stream order)
// HBM traffic: 2 x 2 B per word read + 16 B per event written (~1 word per event in sparse scenes,
// down to 1/12 in dense ones).
#include "evk_internal.cuh"

namespace {

constexpr int kT = 256;          // threads per block
constexpr int kWpt = 16;         // consecutive 16-bit words per thread (two 16-B loads)
constexpr int kWpb = kT * kWpt;  // words per block
constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr uint32_t kSet = 0x80000000u;

struct E3 {              // the effect of a run of words on the decoder state
    uint32_t n;          // CD events the run emits
    uint32_t th_first;   // first / last EVT_TIME_HIGH payload in the run (kNone: none)
    uint32_t th_last;
    uint32_t wraps;      // EVT_TIME_HIGH words of the run lower than the one before them IN the run
    uint32_t tl;         // last EVT_TIME_LOW payload (kNone: none)
    uint32_t y;          // last EVT_ADDR_Y payload (kNone: none)
    uint32_t base;       // bits 14..0 column, bit 15 polarity.  kSet | (polarity, base column after the
                         // run) when the run holds a VECT_BASE_X, else the columns the run's vector
                         // words advance an earlier base by
};
__device__ __forceinline__ E3 e3_identity() { return E3{0, kNone, kNone, 0, kNone, kNone, 0}; }
// the effect of a followed by b
__device__ __forceinline__ E3 e3_combine(const E3& a, const E3& b) {
    E3 r;
    r.n = a.n + b.n;
    if (b.th_first == kNone) {
        r.th_first = a.th_first;
        r.th_last = a.th_last;
        r.wraps = a.wraps;
    } else if (a.th_last == kNone) {
        r.th_first = b.th_first;
        r.th_last = b.th_last;
        r.wraps = b.wraps;
    } else {
        r.th_first = a.th_first;
        r.th_last = b.th_last;
        r.wraps = a.wraps + b.wraps + (b.th_first < a.th_last ? 1u : 0u);
    }
    r.tl = b.tl != kNone ? b.tl : a.tl;
    r.y = b.y != kNone ? b.y : a.y;
    if (b.base & kSet) r.base = b.base;
    else if (a.base & kSet) r.base = (a.base & (kSet | 0x8000u)) | (((a.base & 0x7FFFu) + b.base) & 0x7FFFu);
    else r.base = (a.base + b.base) & 0x7FFFu;
    return r;
}
__device__ __forceinline__ E3 e3_shfl_up(const E3& v, int o) {
    E3 r;
    r.n = __shfl_up_sync(0xffffffffu, v.n, o);
    r.th_first = __shfl_up_sync(0xffffffffu, v.th_first, o);
    r.th_last = __shfl_up_sync(0xffffffffu, v.th_last, o);
    r.wraps = __shfl_up_sync(0xffffffffu, v.wraps, o);
    r.tl = __shfl_up_sync(0xffffffffu, v.tl, o);
    r.y = __shfl_up_sync(0xffffffffu, v.y, o);
    r.base = __shfl_up_sync(0xffffffffu, v.base, o);
    return r;
}
// fold one word into an effect
__device__ __forceinline__ void e3_word(E3& e, uint32_t w) {
    const uint32_t type = w >> 12, v = w & 0xFFFu;
    switch (type) {
        case 0x0: e.y = v & 0x7FFu; break;
        case 0x2: e.n += 1; break;
        case 0x3: e.base = kSet | ((v & 0x800u) << 4) | (v & 0x7FFu); break;
        case 0x4:
            e.n += __popc(v);
            e.base = (e.base & (kSet | 0x8000u)) | (((e.base & 0x7FFFu) + 12u) & 0x7FFFu);
            break;
        case 0x5:
            e.n += __popc(v & 0xFFu);
            e.base = (e.base & (kSet | 0x8000u)) | (((e.base & 0x7FFFu) + 8u) & 0x7FFFu);
            break;
        case 0x6: e.tl = v; break;
        case 0x8:
            if (e.th_last != kNone && v < e.th_last) e.wraps++;
            if (e.th_first == kNone) e.th_first = v;
            e.th_last = v;
            break;
        default: break;
    }
}

__device__ __forceinline__ void load_words(const uint16_t* __restrict__ words, size_t n_words,
                                           size_t at, uint32_t (&w)[kWpt]) {
    if (at + kWpt <= n_words) {
#pragma unroll
        for (int q = 0; q < kWpt / 8; q++) {
            const uint4 v = __ldcs(reinterpret_cast<const uint4*>(words + at) + q);
            const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                w[8 * q + 2 * k] = u[k] & 0xFFFFu;
                w[8 * q + 2 * k + 1] = u[k] >> 16;
            }
        }
    } else {
#pragma unroll
        for (int q = 0; q < kWpt; q++) w[q] = at + q < n_words ? words[at + q] : 0xE000u;  // OTHERS
    }
}

// block-wide exclusive scan of the thread effects; *total = the block's effect
__device__ __forceinline__ E3 block_scan(const E3& mine, E3* total) {
    __shared__ E3 s_w[kT / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    E3 inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const E3 u = e3_shfl_up(inc, o);
        if (lane >= o) inc = e3_combine(u, inc);
    }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    E3 base = e3_identity(), tot = e3_identity();
#pragma unroll
    for (int k = 0; k < kT / 32; k++) {
        if (k < wid) base = e3_combine(base, s_w[k]);
        tot = e3_combine(tot, s_w[k]);
    }
    *total = tot;
    const E3 prev = e3_shfl_up(inc, 1);
    const E3 ex = lane ? e3_combine(base, prev) : base;
    __syncthreads();
    return ex;
}

__device__ __forceinline__ void e3_store(uint32_t* blk, uint32_t nb, uint32_t b, const E3& e) {
    blk[b] = e.n;
    blk[nb + b] = e.th_first;
    blk[2 * nb + b] = e.th_last;
    blk[3 * nb + b] = e.wraps;
    blk[4 * nb + b] = e.tl;
    blk[5 * nb + b] = e.y;
    blk[6 * nb + b] = e.base;
}
__device__ __forceinline__ E3 e3_load(const uint32_t* blk, uint32_t nb, uint32_t b) {
    return E3{blk[b],          blk[nb + b],     blk[2 * nb + b], blk[3 * nb + b],
              blk[4 * nb + b], blk[5 * nb + b], blk[6 * nb + b]};
}

__global__ void __launch_bounds__(kT)
    k_evt3_scan(const uint16_t* __restrict__ words, size_t n_words, uint32_t* blk, uint32_t nb) {
    uint32_t w[kWpt];
    load_words(words, n_words, (size_t)blockIdx.x * kWpb + (size_t)threadIdx.x * kWpt, w);
    E3 e = e3_identity();
#pragma unroll
    for (int q = 0; q < kWpt; q++) e3_word(e, w[q]);
    E3 tot;
    block_scan(e, &tot);
    if (threadIdx.x == 0) e3_store(blk, nb, blockIdx.x, tot);
}

// one CTA: blk[b] <- the combined effect of blocks 0 .. b-1 (in place); total CD events -> *total
__global__ void __launch_bounds__(1024)
    k_evt3_prefix(uint32_t* blk, uint32_t nb, unsigned long long* total) {
    __shared__ E3 s_e[1024];
    __shared__ unsigned long long s_n[1024];  // event counts in 64 bits (the E3 count wraps at 2^32)
    const uint32_t per = (nb + 1023) / 1024;
    const uint32_t b0 = threadIdx.x * per, b1 = min(nb, b0 + per);
    E3 acc = e3_identity();
    unsigned long long cnt = 0;
    for (uint32_t b = b0; b < b1; b++) {
        const E3 e = e3_load(blk, nb, b);
        acc = e3_combine(acc, e);
        cnt += e.n;
    }
    s_e[threadIdx.x] = acc;
    s_n[threadIdx.x] = cnt;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan
        E3 a = e3_identity();
        unsigned long long c = 0;
        if ((int)threadIdx.x >= o) {
            a = s_e[threadIdx.x - o];
            c = s_n[threadIdx.x - o];
        }
        __syncthreads();
        if ((int)threadIdx.x >= o) {
            s_e[threadIdx.x] = e3_combine(a, s_e[threadIdx.x]);
            s_n[threadIdx.x] += c;
        }
        __syncthreads();
    }
    E3 carry = threadIdx.x ? s_e[threadIdx.x - 1] : e3_identity();
    for (uint32_t b = b0; b < b1; b++) {
        const E3 e = e3_load(blk, nb, b);
        e3_store(blk, nb, b, carry);
        carry = e3_combine(carry, e);
    }
    if (threadIdx.x == 1023) *total = s_n[1023];
}

__global__ void __launch_bounds__(kT)
    k_evt3_decode(const uint16_t* __restrict__ words, size_t n_words,
                  const uint32_t* __restrict__ blk, uint32_t nb, evk_event* __restrict__ out,
                  size_t cap) {
    uint32_t w[kWpt];
    load_words(words, n_words, (size_t)blockIdx.x * kWpb + (size_t)threadIdx.x * kWpt, w);
    E3 e = e3_identity();
#pragma unroll
    for (int q = 0; q < kWpt; q++) e3_word(e, w[q]);
    E3 tot;
    const E3 ex = block_scan(e, &tot);
    // the decoder state in front of this thread's first word
    const E3 in = e3_combine(e3_load(blk, nb, blockIdx.x), ex);
    size_t o = in.n;  // (< 2^32: bounded by the handle capacity, checked by the host)
    uint32_t th = in.th_last != kNone ? in.th_last : 0u, wraps = in.wraps;
    uint32_t tl = in.tl != kNone ? in.tl : 0u, y = in.y != kNone ? in.y : 0u;
    uint32_t base = in.base & 0xFFFFu;
    bool have_th = in.th_last != kNone;
    uint4* dst = reinterpret_cast<uint4*>(out);
#pragma unroll
    for (int q = 0; q < kWpt; q++) {
        const uint32_t type = w[q] >> 12, v = w[q] & 0xFFFu;
        if (type == 0x0) y = v & 0x7FFu;
        else if (type == 0x6) tl = v;
        else if (type == 0x8) {
            if (have_th && v < th) wraps++;
            th = v;
            have_th = true;
        } else if (type == 0x3) base = ((v & 0x800u) << 4) | (v & 0x7FFu);
        else if (type == 0x2 || type == 0x4 || type == 0x5) {
            const uint64_t t = ((uint64_t)wraps << 24) | ((uint64_t)th << 12) | tl;
            if (type == 0x2) {
                if (o < cap)
                    dst[o] = make_uint4((v & 0x7FFu) | (y << 16), (v >> 11) & 1u, (uint32_t)t,
                                        (uint32_t)(t >> 32));
                o++;
            } else {
                uint32_t m = type == 0x4 ? v : (v & 0xFFu);
                const uint32_t bx = base & 0x7FFFu, pol = base >> 15;  // polarity: bit 11 of VECT_BASE_X
                while (m) {
                    const uint32_t i = __ffs(m) - 1;
                    m &= m - 1;
                    if (o < cap)
                        dst[o] = make_uint4(((bx + i) & 0x7FFFu) | (y << 16), pol, (uint32_t)t,
                                            (uint32_t)(t >> 32));
                    o++;
                }
                base = (base & 0x8000u) | ((bx + (type == 0x4 ? 12u : 8u)) & 0x7FFFu);
            }
        }
    }
}

}  // namespace

// words: device pointer.  blk: scratch of 7 * n_blocks u32.  total: device scalar (CD events).
cudaError_t evk_launch_evt3_decode(const uint16_t* words, size_t n_words, uint32_t* blk,
                                   unsigned long long* total, evk_event* out, size_t cap,
                                   cudaStream_t s) {
    const uint32_t nb = (uint32_t)((n_words + kWpb - 1) / kWpb);
    k_evt3_scan<<<nb, kT, 0, s>>>(words, n_words, blk, nb);
    k_evt3_prefix<<<1, 1024, 0, s>>>(blk, nb, total);
    k_evt3_decode<<<nb, kT, 0, s>>>(words, n_words, blk, nb, out, cap);
    return cudaGetLastError();
}
size_t evk_evt3_blocks(size_t n_words) { return (n_words + kWpb - 1) / kWpb; }
