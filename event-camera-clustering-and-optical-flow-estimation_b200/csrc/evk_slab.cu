// evk_slab.cu — time-slab downsample: the sm_100a fast path for time-ordered event streams.
//
// Event streams are time-ordered (the SDK delivers them so: ACCEL/store.cpp:614-615), and the time
// bin is the most significant part of the voxel key, so all events of one time bin are contiguous.
// One CTA owns one time bin at a time and keeps the WHOLE key space of that bin on chip:
//   s_map    2 bits per spatial cell: "seen" and "hit at least twice" (the kernel's
//            repeated_count semantics, ACCEL/build/coordinate_processor.cl:73-75), 16 cells per
//            word so one load answers both (115 KB for Gen4 2x2 px + polarity)
//   s_ev     a 2-stage ring of 3072-event tiles (48 KB each) filled by TMA bulk copies
//            (cp.async.bulk + mbarrier complete_tx) issued by an elected thread across bin
//            boundaries (bins are dealt round-robin, so the order is known).  The classify pass
//            drains a tile into registers, so its slot is refilled one barrier later (tile + 2):
//            the prefetch distance of a three-stage ring at two thirds of the memory, which is
//            what lets a tile hold three events per thread (round 2: -3.8 %, profiles/r02)
//   s_late   two small tables (one per tile parity) for the rare events that lose a claim race
// Per tile:  classify  one bitmap load per event: seen -> duplicate of an earlier tile (its cell's
//                      "hit twice" bit is set by a reduction every lane issues, most with a zero)
//            -- barrier: every read of the tile precedes every claim of the tile --
//            claim     each unseen event does ONE returning atomicOr on the bitmap: bit was clear
//                      -> it owns the cell ("claimant"); bit was set -> a peer of the SAME tile got
//                      there first ("late peer"): its packed (cell, index) word goes into s_late,
//                      where a 32-bit atomicMin keeps the lowest index of the cell, and the cell's
//                      "hit twice" bit is set
//            -- barrier --
//            resolve   a claimant whose cell is now "hit twice" probes s_late: a late peer with a
//                      lower index replaces it as representative (lowest stream index wins, SURVEY
//                      8a; its coordinates are re-read from global memory: rare, L2-resident); then
//                      warp ballot + one shared atomic per warp claim output slots and the 16-B
//                      record goes to HBM from registers through predicated stores
// Two block barriers per tile and no global atomic: output slots come from a CTA-private chunk of
// the output arrays (claimed from a global counter one chunk ahead); a small fix-up pass moves
// the tail of the last chunks into the holes so the voxel shard is dense.  HBM sees each event
// read once and each voxel written once (16 B SoA); no table lives in HBM.
//
// This is the "perfect-hash" specialisation of the mandated open-addressing table
// (evk_downsample.cu), which remains the general path: the kernel verifies on the fly that
// [bin_start[b], bin_start[b+1]) holds only events of bin b and that the ranges partition the
// stream; any violation makes the host fall back to the table.
#include <type_traits>

#include "evk_internal.cuh"

namespace {

#ifndef EVK_SLAB_THREADS
#define EVK_SLAB_THREADS 1024
#endif
#ifndef EVK_SLAB_PER
#define EVK_SLAB_PER 3
#endif
#ifndef EVK_SLAB_STAGES
#define EVK_SLAB_STAGES 2
#endif
#ifndef EVK_SLAB_CHECKS
#define EVK_SLAB_CHECKS 0  // 1: every index of the hot kernel is bounds-checked on the device (a failed
#endif                     //    check sets DsCounters::overflow bit 1); compute-sanitizer is closed
                           //    on this pool, so the parity / stress tests are run against this build
// (round 2's A/B switches -- every-lane hit-twice reduction, short-row skip, predicate for the
// out-of-bin check, ... -- are folded into the code: profiles/r02/slab_ab_runs.md has the runs)
constexpr int kThreads = EVK_SLAB_THREADS;  // CTA size (the hardware maximum by default)
constexpr int kCtasPerSm = kThreads <= 512 ? 2 : 1;
constexpr int kPer = EVK_SLAB_PER;          // events per thread per tile
constexpr int kTile = kThreads * kPer;      // events per tile (stage buffers hold this many)
constexpr int kLogTile = kTile > 2048 ? 12 : kTile > 1024 ? 11 : 10;  // index-in-tile field
constexpr uint32_t kIdxMask = (1u << kLogTile) - 1u;
constexpr int kLogHash = kTile > 2048 ? 11 : kLogTile;
constexpr int kHash = 1 << kLogHash;    // late-peer table slots per parity (a late cell has at
                                        // least two events in the tile: at most kTile / 2 entries)
constexpr int kStages = EVK_SLAB_STAGES;
static_assert(kTile / 2 < kHash, "late-peer table must never fill up");
static_assert(kTile <= (1 << kLogTile), "index field too narrow");
constexpr uint32_t kEmpty = 0xFFFFFFFFu;
constexpr uint32_t kChunk = EVK_SLAB_CHUNK;  // output slots per CTA-private chunk (>= 2 tiles)
constexpr uint32_t kNoChunk = 0xFFFFFFFFu;
static_assert(kChunk >= 2 * kTile, "a fresh chunk must absorb a whole tile");
static_assert(kThreads * kPer == kTile, "tile = events per thread x CTA size");
static_assert(kThreads >= kHash / 4, "one 16-B store per thread clears a late-peer table");

// what a CTA leaves behind for the fix-up pass: the two output chunks it may have left partly
// filled (the current one and the one claimed ahead, never written to)
struct ChunkTail {
    uint32_t cur_base, cur_filled;    // current chunk: first slot, slots written
    uint32_t next_base, next_filled;  // chunk claimed ahead (kNoChunk-based if none); always 0 filled
};

#if EVK_SLAB_CHECKS
#define SLAB_CHECK(cond)                                   \
    do {                                                   \
        if (!(cond)) atomicOr(&cnt->overflow, 2u);         \
    } while (0)
#else
#define SLAB_CHECK(cond) ((void)0)
#endif

struct SlabArgs {
    KeyParams kp;
    const evk_event* ev;
    size_t n;
    uint32_t* bin_start;
    uint64_t* keys;
    uint32_t* first;
    uint32_t* xy;
    DsCounters* cnt;
    ChunkTail* chunk_list;  // [grid]
    uint32_t first_offset;  // added to every emitted first index (global index of event 0)
    uint32_t words;         // bitmap words per bin
    uint32_t max_bins;
    uint32_t min_bins;
    uint32_t out_cap;       // slots of keys / first / xy (checked build)
    // sharded runs: [0] != 0 -> give up; [3] = leading events that belong to the previous rank's
    // last bin; [4] = events received behind my own that belong to my last bin.  Read on the device
    // so that the halo exchange and the downsample need no host round trip in between.
    const unsigned long long* range;
    const long long* t0_dev;  // not null: the time origin (overrides kp.t0; replayed graphs)
};

// the event range the kernels work on
struct SlabRange {
    const evk_event* ev;
    size_t n;
    uint32_t first_offset;
    bool bad;
};
__device__ __forceinline__ SlabRange slab_range(const SlabArgs& a) {
    SlabRange r;
    r.ev = a.ev;
    r.n = a.n;
    r.first_offset = a.first_offset;
    r.bad = false;
    if (a.range) {
        const size_t skip = (size_t)a.range[3], keep = (size_t)a.range[4];
        r.bad = a.range[0] != 0 || skip > a.n;
        if (!r.bad) {
            r.ev = a.ev + skip;
            r.n = a.n - skip + keep;
            r.first_offset = a.first_offset + (uint32_t)skip;
        }
    }
    return r;
}

// ---- PTX helpers: mbarrier + 1-D TMA bulk copy ------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "EVK_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra EVK_DONE_%=;\n\t"
        "bra EVK_WAIT_%=;\n\t"
        "EVK_DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// scratch[0] = n_bins, scratch[1] = bin work counter, scratch[2] = tb0, scratch[3] = chunk counter
// bin_start[b] = first index whose time bin is >= tb0 + b.  One warp per bin, 32-way search: six
// rounds of 32 parallel probes instead of 27 dependent loads.
__global__ void __launch_bounds__(256) k_slab_bins(SlabArgs a) {
    DsCounters* cnt = a.cnt;
    KeyParams kp = a.kp;
    if (a.t0_dev) kp.t0 = *a.t0_dev;
    const SlabRange rg = slab_range(a);
    if (rg.bad || rg.n == 0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) cnt->slab_violation = 2;
        return;
    }
    const evk_event* ev = rg.ev;
    const size_t n = rg.n;
    const uint4 e0 = ld_event(ev);
    const uint4 e1 = ld_event(ev + (n - 1));
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
    const int lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t b = warp; b <= nb; b += n_warps) {
        size_t lo = 0, hi = n;  // invariant: pred(lo - 1) false, pred(hi) true (pred(n) = true)
        if (b == 0) hi = 0;
        else if (b == nb) lo = n;
        while (lo < hi) {
            const size_t span = hi - lo;
            // 32 probes spread over [lo, hi): lo + span * (l + 1) / 33 (span < 2^33: no overflow)
            const size_t p = lo + span * (size_t)(lane + 1) / 33u;
            const int64_t t = *reinterpret_cast<const int64_t*>(
                reinterpret_cast<const char*>(ev + p) + 8);
            const bool ge = t >= kp.t0 && evk_tbin(kp, t) >= tb0 + b;
            const uint32_t m = __ballot_sync(0xffffffffu, ge);
            // first lane whose probe satisfies the predicate bounds hi; the lane before it bounds lo
            const int f = m ? __ffs(m) - 1 : 32;
            const size_t p_f = __shfl_sync(0xffffffffu, p, f < 32 ? f : 0);
            const size_t p_b = __shfl_sync(0xffffffffu, p, f > 0 ? f - 1 : 0);
            if (f < 32) hi = p_f;
            if (f > 0) lo = p_b + 1;
        }
        if (lane == 0) a.bin_start[b] = (uint32_t)lo;
    }
}

__device__ __forceinline__ uint32_t hash_slot(uint32_t cell) {
    return (cell * 0x9E3779B1u) >> (32 - kLogHash);
}

// lane 0 of the last producer warp issues the bulk copies; thread 0 keeps the output-chunk
// bookkeeping

template <int NT>
__device__ __forceinline__ void prod_sync() {
    asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
}

// the CTA's tile stream: bins blockIdx.x, blockIdx.x + gridDim.x, ...; TILE-event tiles in each
struct TileIt {
    uint32_t b, hi, base;  // tile [base, min(base + TILE, hi)) of bin b; b >= nb: end of stream
};
// first tile of the first non-empty bin at or after it.b
__device__ __forceinline__ void it_enter(TileIt& it, const uint32_t* bin_start, uint32_t nb) {
    while (it.b < nb) {
        const uint32_t lo = bin_start[it.b], hi = bin_start[it.b + 1];
        if (hi > lo) {  // (hi < lo: the stream is not time-ordered; the consumer side reports it)
            it.base = lo;
            it.hi = hi;
            return;
        }
        it.b += gridDim.x;
    }
}

template <bool COUNT_REP, bool POW2, bool USE_P>
__global__ void __launch_bounds__(kThreads, kCtasPerSm) k_slab_main(SlabArgs a) {
    constexpr int NT = kThreads;
    constexpr int TILE = NT * kPer;  // events per tile
    constexpr int kTmaThread = NT - 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint4* s_ev = reinterpret_cast<uint4*>(smem_raw);                        // [kStages][kTile]
    uint32_t* s_late = reinterpret_cast<uint32_t*>(s_ev + kStages * kTile);  // [2][kHash]
    // bin bitmap.  COUNT_REP: 16 cells per word, bit c = "seen", bit 16 + c = "hit at least twice"
    // (one load answers both questions); else 32 cells per word, "seen" only.
    uint32_t* s_map = s_late + 2 * kHash;  // [words]
    __shared__ __align__(8) uint64_t s_bar[kStages];
    __shared__ uint32_t s_cursor[2];  // voxels emitted by the current tile (by tile parity)
    // output chunk state + the early-stop flag (some CTA has found the stream unordered), one
    // 16-byte group: the output pass reads it with one load
    __shared__ __align__(16) uint32_t s_chunk[4];
    uint32_t& s_chunk_pos = s_chunk[0];
    uint32_t& s_chunk_end = s_chunk[1];
    uint32_t& s_next_base = s_chunk[2];
    uint32_t& s_stop = s_chunk[3];
    // Per-bin and per-launch constants the passes read back ONCE PER TILE through shared memory.
    // Held in registers across the tile loop they do not survive: ptxas rematerialises the 64-bit
    // products (six instructions in front of every key store, ncu source page of round 2) and
    // re-reads the column bases from the constant bank in front of every record.
    __shared__ int64_t s_t_lo;  // first microsecond of the current bin
    // [0..2] global addresses of the key / first-index / xy columns, [3] first key of the current bin
    __shared__ __align__(16) uint64_t s_out[4];
    uint64_t& s_key_base = s_out[3];

    DsCounters* cnt = a.cnt;
    if (cnt->slab_violation) return;
    const KeyParams& kp = a.kp;
    const int64_t t0 = a.t0_dev ? *a.t0_dev : a.kp.t0;  // (replayed graphs: device-side origin)
    const uint32_t nb = (uint32_t)cnt->scratch[0];
    const uint64_t tb0 = cnt->scratch[2];
    const int tid = threadIdx.x;
    const SlabRange rg = slab_range(a);
    const evk_event* const evs = rg.ev;
    const uint32_t first_offset = rg.first_offset;
    uint32_t lane_lt;  // lanes below mine (the output pass ranks a lane among the emitting lanes)
    asm volatile("mov.u32 %0, %%lanemask_lt;" : "=r"(lane_lt));

    for (int i = tid; i < 2 * kHash; i += NT) s_late[i] = kEmpty;
    for (uint32_t i = tid; i < a.words; i += NT) s_map[i] = 0;
    if (tid == 0) {
        for (int s = 0; s < kStages; s++) mbar_init(&s_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_cursor[0] = s_cursor[1] = 0;
        s_out[0] = (uint64_t)__cvta_generic_to_global(a.keys);
        s_out[1] = (uint64_t)__cvta_generic_to_global(a.first);
        s_out[2] = (uint64_t)__cvta_generic_to_global(a.xy);
        // two chunks up front: the current one and the one after it
        const uint32_t c0 = (uint32_t)atomicAdd(&cnt->scratch[3], 2ull);
        s_chunk_pos = c0 * kChunk;
        s_chunk_end = s_chunk_pos + kChunk;
        s_next_base = (c0 + 1) * kChunk;
    }
    __syncthreads();
    uint32_t pend_chunk = kNoChunk;  // thread 0: chunk index requested but not yet published
    uint32_t book_par = 2;           // thread 0: parity of the tile whose output is not yet booked
    uint32_t tile_seq = 0;           // tiles consumed by this CTA (stage = seq % kStages)
    uint32_t viol = 0;

    // thread 0, once every warp has claimed its output slots of a tile: advance the chunk state
    auto book = [&]() {
        if (book_par > 1) return;
        if (pend_chunk != kNoChunk) {  // requested one tile ago: has arrived by now
            s_next_base = pend_chunk * kChunk;
            pend_chunk = kNoChunk;
        }
        const uint32_t c = s_cursor[book_par];  // voxels that tile emitted
        const uint32_t room = s_chunk_end - s_chunk_pos;
        if (c >= room) {  // spilled into the next chunk: make it current, request another
            const uint32_t nb0 = s_next_base;
            s_chunk_pos = nb0 + (c - room);
            s_chunk_end = nb0 + kChunk;
            s_next_base = kNoChunk;
            pend_chunk = (uint32_t)atomicAdd(&cnt->scratch[3], 1ull);
        } else {
            s_chunk_pos += c;
        }
        s_cursor[book_par] = 0;
        book_par = 2;
    };

    // the TMA thread runs two tiles ahead of the consumers through the same tile stream
    TileIt tma;
    tma.b = blockIdx.x;
    tma.hi = tma.base = 0;
    uint32_t tma_seq = 0;
    auto fetch = [&]() {  // issue the next tile of the stream, if any
        if (tma.b >= nb) return;
        const uint32_t cntev = min((uint32_t)TILE, tma.hi - tma.base);
        uint64_t* bar = &s_bar[tma_seq % kStages];
        mbar_expect_tx(bar, cntev * 16u);
        tma_load_1d(s_ev + (tma_seq % kStages) * kTile, evs + tma.base, cntev * 16u, bar);
        tma_seq++;
        tma.base += TILE;
        if (tma.base >= tma.hi) {
            tma.b += gridDim.x;
            it_enter(tma, a.bin_start, nb);
        }
    };
    if (tid == kTmaThread) {
        it_enter(tma, a.bin_start, nb);
#pragma unroll
        for (int s = 0; s < kStages; s++) fetch();
    }

    // y < height  <=>  (y << 16 | x) <= ((height - 1) << 16 | 0xFFFF): no need to extract y
    const uint32_t y_lim = ((uint32_t)(kp.height - 1) << 16) | 0xFFFFu;
    const uint32_t ysh = 16u + (uint32_t)(kp.sy >= 0 ? kp.sy : 0);
    for (uint32_t b = blockIdx.x; b < nb; b += gridDim.x) {
        const uint32_t lo = a.bin_start[b], hi = a.bin_start[b + 1];
        if (hi < lo) viol = 1;  // ranges do not partition the stream: not time-ordered
        if (hi <= lo) continue;
        if (tid == 0) s_stop = *reinterpret_cast<volatile unsigned int*>(&cnt->slab_violation);
        prod_sync<NT>();  // every thread has left the previous bin (bitmap, cursor, key base)
        if (s_stop) break;  // (an early stop inside the tile loop also ends here)
        if (tid == 0) {
            // (only now: slower warps were still writing the previous bin's records, which read
            // s_key_base, until the barrier above)
            const uint64_t tb = tb0 + b;
            s_key_base = tb * kp.cells;
            s_t_lo = t0 + (int64_t)(tb * (uint64_t)kp.vt);
            book();
        }
        {  // count the previous bin's repeated cells and clear the bitmap in one pass
            uint32_t r = 0;
            for (uint32_t i = tid; i < a.words; i += NT) {
                if (COUNT_REP) r += __popc(s_map[i] >> 16);
                s_map[i] = 0;
            }
            if (COUNT_REP) {
                r = __reduce_add_sync(0xffffffffu, r);
                if ((tid & 31) == 0 && r) atomicAdd(&cnt->n_repeated, (unsigned long long)r);
            }
        }
        prod_sync<NT>();

        for (uint32_t base = lo; base < hi; base += TILE, tile_seq++) {
            const uint32_t stage = tile_seq % kStages, par = tile_seq & 1;
            uint4* tile = s_ev + stage * kTile;
            uint32_t* late_tbl = s_late + par * kHash;
            mbar_wait(&s_bar[stage], (tile_seq / kStages) & 1);
            const int64_t t_lo = *reinterpret_cast<volatile int64_t*>(&s_t_lo);
            // ---- classify: cv = candidate word (cell << kLogTile | index in tile) or kEmpty;
            // cw / cb = bitmap word index and bit of the event's cell (reused by the claim pass)
            uint32_t cv[kPer], cxy[kPer], cw[kPer], cb[kPer];
            bool tile_viol = false;
            auto classify_row = [&](int j) {
                const uint32_t li = j * NT + tid;
                const uint4 ev = tile[li];
                const uint32_t x = ev.x & 0xFFFFu, y = ev.x >> 16;
                const bool gate = (x < (uint32_t)kp.width) & (ev.x <= y_lim);
                const bool inbin = (uint64_t)(ev_t(ev) - t_lo) < (uint64_t)kp.vt;
                const bool ok = gate & inbin;
                // a gated-in event that is not of this bin: the stream is not what the bins say
                tile_viol |= gate != ok;
                uint32_t cell;
                if (POW2) cell = (ev.x >> ysh) * kp.NX + (x >> kp.sx);
                else cell = (kp.sy >= 0 ? y >> kp.sy : __umulhi(y, kp.my)) * kp.NX +
                            (kp.sx >= 0 ? x >> kp.sx : __umulhi(x, kp.mx));
                // polarity bit = (int16)p > 0: the sign test on the halfword moved to bit 31
                if (USE_P) cell = cell * 2u + ((int32_t)(ev.y << 16) > 0 ? 1u : 0u);
                uint32_t w = COUNT_REP ? cell >> 4 : cell >> 5;
                const uint32_t sbit = 1u << (cell & (COUNT_REP ? 15u : 31u));
                SLAB_CHECK(!ok || (w < a.words && li < (uint32_t)kTile));
                w = ok ? w : 0u;  // (every lane issues the reduction below: a word any lane may touch)
                const uint32_t wv = ok ? s_map[w] : 0xFFFFFFFFu;  // gated events: nothing to do
                if (COUNT_REP) {  // duplicate of an earlier tile's voxel: mark it "hit twice"
                    // every lane issues the reduction, most with a zero (gated events read all
                    // ones, so theirs is zero too): ptxas turns a predicated shared atomic into a
                    // branch around it (BSSY / BRA / BSYNC) although some lane of the warp nearly
                    // always takes it
                    atomicOr(&s_map[w], ((wv & sbit) << 16) & ~wv);
                }
                cv[j] = (wv & sbit) ? kEmpty : ((cell << kLogTile) | li);
                cxy[j] = ev.x;
                cw[j] = w;
                cb[j] = sbit;
            };
            if (hi - base < (uint32_t)TILE) {  // (uniform over the CTA)
                // the last tile of a bin is short: its unused slots still hold an older tile.
                // Overwrite them with an event outside the frame (x = y = 0xFFFF), so that the
                // per-event test "index < end of bin" -- two instructions per event of every tile
                // -- is not needed
                const uint32_t live = hi - base;
#pragma unroll
                for (int j = 0; j < kPer; j++)
                    if ((uint32_t)(j * NT + tid) >= live)
                        tile[j * NT + tid] = make_uint4(0xFFFFFFFFu, 0u, 0u, 0u);
                // (the slot is next written by a bulk copy: order the two proxies)
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                // (its own copy of the classify pass: the rows beyond the end of the bin -- on
                // average a tile's worth of rows per bin -- are skipped)
#pragma unroll
                for (int j = 0; j < kPer; j++) {
                    if ((uint32_t)(j * NT) < live) {
                        classify_row(j);
                    } else {
                        cv[j] = kEmpty;
                        cxy[j] = cw[j] = cb[j] = 0;
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < kPer; j++) classify_row(j);
            }
            if (tile_viol) viol |= 1u;
            // a rejected stream is rerun on a general path: publish the violation at once and let
            // every CTA stop at its next tile (uniformly: thread 0 reads, the barrier broadcasts)
            const bool poll = (tile_seq & 15u) == 0;  // (every 16th tile: polling every 4th cost 2 %)
            if (poll) {
                if (viol == 1) {
                    atomicOr(&cnt->slab_violation, 1u);
                    viol = 3;
                }
                if (tid == 0)
                    s_stop = *reinterpret_cast<volatile unsigned int*>(&cnt->slab_violation);
            }
            prod_sync<NT>();  // B0: every bitmap read of this tile precedes every claim below
            if (poll && s_stop) break;
            // this tile's ring slot has been drained into registers: refill it with the tile after
            // the next one
            if (tid == kTmaThread) fetch();
            if (tid == 0) book();
            // ---- claim: one returning atomic per unseen event
            bool late[kPer];
#pragma unroll
            for (int j = 0; j < kPer; j++) {
                // (issued by every lane with a zero for lanes without a candidate, this returning
                // atomic costs more than the branch ptxas builds around the predicated one: A/B)
                uint32_t old = 0;  // (stays 0 for lanes without a candidate)
                asm volatile(
                    "{ .reg .pred q; setp.ne.u32 q, %3, 0xFFFFFFFF; "
                    "@q atom.shared.or.b32 %0, [%1], %2; }"
                    : "+r"(old)
                    : "r"(smem_u32(&s_map[cw[j]])), "r"(cb[j]), "r"(cv[j])
                    : "memory");
                late[j] = (old & cb[j]) != 0;
            }
#pragma unroll
            for (int j = 0; j < kPer; j++) {  // a peer of this tile claimed the cell first (rare)
                if (!late[j]) continue;
                const uint32_t cell = cv[j] >> kLogTile;
                if (COUNT_REP) atomicOr(&s_map[cw[j]], cb[j] << 16);  // the cell is hit twice
                uint32_t s = hash_slot(cell);
#if EVK_SLAB_CHECKS
                uint32_t probes = 0;
#endif
                for (;;) {
                    SLAB_CHECK(s < (uint32_t)kHash && ++probes <= (uint32_t)kHash);
                    const uint32_t old = atomicCAS(&late_tbl[s], kEmpty, cv[j]);
                    if (old == kEmpty) break;
                    if ((old >> kLogTile) == cell) {  // keep the lowest index of the cell
                        atomicMin(&late_tbl[s], cv[j]);
                        break;
                    }
                    s = (s + 1) & (kHash - 1);
                }
                cv[j] = kEmpty;
            }
            prod_sync<NT>();  // S1: every claim of the tile is in the bitmap, s_late is complete
            // ---- resolve: the claimant is the new voxel unless a late peer has a lower index.
            // COUNT_REP: a late peer has set the cell's "hit twice" bit (so may a warp already
            // classifying the next tile: then the probe finds nothing) -- one load of a word whose
            // address and bit are in registers answers "is there a late peer" without hashing the
            // cell; else the first probe of the table (nearly always an empty slot) does
            uint32_t bal[kPer], wtot = 0, w0[kPer];
#pragma unroll
            for (int j = 0; j < kPer; j++) {
                if (COUNT_REP) w0[j] = cv[j] != kEmpty ? s_map[cw[j]] & (cb[j] << 16) : 0u;
                else w0[j] = cv[j] != kEmpty ? ~late_tbl[hash_slot(cv[j] >> kLogTile)] : 0u;
            }
#pragma unroll
            for (int j = 0; j < kPer; j++) {
                if (w0[j]) {
                    const uint32_t cell = cv[j] >> kLogTile;
                    uint32_t s = hash_slot(cell);
                    for (;;) {
                        const uint32_t w = late_tbl[s];
                        if (w == kEmpty) break;
                        if ((w >> kLogTile) == cell) {
                            if (w < cv[j]) {  // lower index: that event is the representative
                                cv[j] = w;
                                // (its ring slot may already hold a later tile: rare, L2-resident)
                                SLAB_CHECK((size_t)base + (w & kIdxMask) < rg.n);
                                cxy[j] = __ldg(reinterpret_cast<const uint32_t*>(
                                    evs + (base + (w & kIdxMask))));
                            }
                            break;
                        }
                        s = (s + 1) & (kHash - 1);
                    }
                }
                bal[j] = __ballot_sync(0xffffffffu, cv[j] != kEmpty);
                wtot += __popc(bal[j]);
            }
            // the other parity's table was last read one tile ago: clean it for the next tile
            if (tid < kHash / 4)
                reinterpret_cast<uint4*>(s_late + (par ^ 1) * kHash)[tid] =
                    make_uint4(kEmpty, kEmpty, kEmpty, kEmpty);
            // one shared atomic per warp claims the warp's output slots.  elect.sync names the one
            // lane: behind a lane test ptxas cannot prove the atomic is single-lane and wraps it in
            // its own same-address aggregation (VOTE / FLO / POPC / SHFL, a dozen instructions)
            uint32_t wbase = 0, leader;
            asm volatile(
                "{ .reg .pred q; elect.sync %1|q, 0xffffffff; @q atom.shared.add.u32 %0, [%2], %3; }"
                : "+r"(wbase), "=r"(leader)
                : "r"(smem_u32(&s_cursor[par])), "r"(wtot)
                : "memory");
            wbase = __shfl_sync(0xffffffffu, wbase, leader);
            {
                // chunk state, column bases and the bin's first key: three 16-byte shared loads
                uint32_t pos0, end0, nxt0, stop_unused;
                asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(pos0), "=r"(end0), "=r"(nxt0), "=r"(stop_unused)
                             : "r"(smem_u32(s_chunk)));
                (void)stop_unused;
                uint64_t keys_g, first_g, xy_g, key_base;
                asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];"
                             : "=l"(keys_g), "=l"(first_g)
                             : "r"(smem_u32(&s_out[0])));
                asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];"
                             : "=l"(xy_g), "=l"(key_base)
                             : "r"(smem_u32(&s_out[2])));
                const uint32_t room = end0 - pos0;  // slots left in the current chunk
                // the warp's slots [wbase, wbase + wtot) nearly always lie on one side of the chunk
                // boundary: then slot = rank + one warp-uniform offset, no per-record select
                const bool straddle = (wbase < room) & (wbase + wtot > room);
                auto emit = [&](auto per_lane) {
                    const uint32_t off = wbase < room ? pos0 : nxt0 - room;
                    uint32_t wb = wbase;
#pragma unroll
                    for (int j = 0; j < kPer; j++) {
                        const uint32_t o = wb + __popc(bal[j] & lane_lt);
                        uint32_t p;
                        if (decltype(per_lane)::value) p = o < room ? pos0 + o : nxt0 + (o - room);
                        else p = o + off;
                        SLAB_CHECK(cv[j] == kEmpty || (p < a.out_cap && (o < room || nxt0 != kNoChunk)));
                        asm volatile(
                            "{ .reg .pred q; setp.ne.u32 q, %6, 0xFFFFFFFF;\n\t"
                            "@q st.global.u64 [%0], %1;\n\t"
                            "@q st.global.u32 [%2], %3;\n\t"
                            "@q st.global.u32 [%4], %5; }" ::"l"(keys_g + (uint64_t)p * 8u),
                            "l"(key_base + (cv[j] >> kLogTile)), "l"(first_g + (uint64_t)p * 4u),
                            "r"(base + (cv[j] & kIdxMask) + first_offset),
                            "l"(xy_g + (uint64_t)p * 4u), "r"(cxy[j]), "r"(cv[j])
                            : "memory");
                        wb += __popc(bal[j]);
                    }
                };
                if (straddle) emit(std::true_type{});
                else emit(std::false_type{});
            }
            if (tid == 0) book_par = par;
        }
    }
    // (early stop: bulk copies still in flight must land before the CTA gives up its shared memory)
    if (tid == kTmaThread)
        for (uint32_t q = tile_seq; q < tma_seq; q++) mbar_wait(&s_bar[q % kStages], (q / kStages) & 1);
    prod_sync<NT>();
    if (tid == 0) book();
    if (COUNT_REP) {  // the last bin's repeated cells
        uint32_t r = 0;
        for (uint32_t i = tid; i < a.words; i += NT) r += __popc(s_map[i] >> 16);
        r = __reduce_add_sync(0xffffffffu, r);
        if ((tid & 31) == 0 && r) atomicAdd(&cnt->n_repeated, (unsigned long long)r);
    }
    if (viol) atomicOr(&cnt->slab_violation, 1u);
    if (tid == 0) {  // publish the two chunks this CTA leaves partly filled
        if (pend_chunk != kNoChunk) s_next_base = pend_chunk * kChunk;
        ChunkTail ct;
        ct.cur_base = s_chunk_end - kChunk;
        ct.cur_filled = kChunk - (s_chunk_end - s_chunk_pos);
        ct.next_base = s_next_base;
        ct.next_filled = 0;
        a.chunk_list[blockIdx.x] = ct;
    }
}

// ---- fix-up: make the voxel shard dense -------------------------------------------------------
// Chunks not listed are full.  U = total filled.  Every live slot at a position >= U is moved into
// a hole (unfilled slot of a listed chunk) below U; the counts match by construction.  One launch:
// every CTA rebuilds the (tiny) plan in shared memory from the chunk list, then moves its share.
constexpr int kMaxList = 1024;
constexpr int kFixThreads = 1024;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_warp,
                                                         uint32_t* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
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
        s_warp[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const uint32_t base = wid ? s_warp[wid - 1] : 0u;
    *total = s_warp[31];
    __syncthreads();
    return base + inc - v;
}

__device__ __forceinline__ uint32_t plan_locate(const uint32_t* start, const uint32_t* prefix,
                                                uint32_t n, uint32_t j) {
    uint32_t lo = 0, hi = n;  // last segment with prefix <= j (skips empty segments)
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (prefix[mid] <= j) lo = mid;
        else hi = mid;
    }
    return start[lo] + (j - prefix[lo]);
}

__global__ void __launch_bounds__(kFixThreads)
    k_slab_fix(const ChunkTail* chunk_list, uint32_t n_ctas, DsCounters* cnt, uint64_t* keys,
               uint32_t* first, uint32_t* xy, unsigned long long* sticky) {
    __shared__ uint32_t s_base[kMaxList], s_fill[kMaxList];
    __shared__ uint32_t s_src_start[kMaxList + 1], s_src_prefix[kMaxList + 1];
    __shared__ uint32_t s_dst_start[kMaxList], s_dst_prefix[kMaxList];
    __shared__ uint32_t s_warp[32];
    if (cnt->slab_violation) {  // the step is rejected: remembered across queued steps
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(sticky, 1ull);
        return;
    }
    const uint32_t M = (uint32_t)cnt->scratch[3];  // chunks handed out
    const uint32_t i = threadIdx.x;
    const uint32_t n_list = 2 * n_ctas;  // two listed chunks per CTA
    uint32_t hole = 0;
    if (i < n_list) {
        const ChunkTail ct = chunk_list[i >> 1];
        s_base[i] = (i & 1) ? ct.next_base : ct.cur_base;
        s_fill[i] = (i & 1) ? ct.next_filled : ct.cur_filled;
        hole = kChunk - s_fill[i];
    }
    uint32_t holes;
    block_exclusive_scan(hole, s_warp, &holes);
    const uint64_t U = (uint64_t)M * kChunk - holes;
    const uint32_t m0 = (uint32_t)(U / kChunk);  // first chunk that may hold slots >= U
    const uint32_t n_src = min(M - m0, (uint32_t)kMaxList);  // <= n_list + 1 by construction
    // destination segments: the part of each listed chunk's hole that lies below U
    uint32_t dlen = 0;
    if (i < n_list) {
        const uint64_t h0 = (uint64_t)s_base[i] + s_fill[i], h1 = (uint64_t)s_base[i] + kChunk;
        const uint64_t e = h1 < U ? h1 : U;
        s_dst_start[i] = (uint32_t)h0;
        dlen = h0 < e ? (uint32_t)(e - h0) : 0;
    }
    uint32_t dtot;
    const uint32_t dpre = block_exclusive_scan(dlen, s_warp, &dtot);
    if (i < n_list) s_dst_prefix[i] = dpre;
    // source segments: live slots at positions >= U, chunk by chunk (a chunk that is not listed is
    // full; the listed ones scatter their fill into the table)
    if (i < n_src) s_src_prefix[i] = kChunk;
    __syncthreads();
    if (i < n_list && s_base[i] % kChunk == 0) {
        const uint32_t m = s_base[i] / kChunk;
        if (m >= m0 && m - m0 < n_src) s_src_prefix[m - m0] = s_fill[i];
    }
    __syncthreads();
    uint32_t slen = 0;
    if (i < n_src) {
        const uint32_t m = m0 + i;
        const uint32_t fill = s_src_prefix[i];
        const uint64_t l0 = (uint64_t)m * kChunk, l1 = l0 + fill;
        const uint64_t s = l0 > U ? l0 : U;
        s_src_start[i] = (uint32_t)s;
        slen = s < l1 ? (uint32_t)(l1 - s) : 0;
    }
    uint32_t stot;
    const uint32_t spre = block_exclusive_scan(slen, s_warp, &stot);
    if (i < n_src) s_src_prefix[i] = spre;
    __syncthreads();
    if (blockIdx.x == 0 && i == 0) {
        cnt->n_unique = U;
        if (stot != dtot || M - m0 > (uint32_t)kMaxList) cnt->overflow = 1;  // cannot happen
    }
    const uint32_t total = stot < dtot ? stot : dtot;
    if (n_src == 0 || n_list == 0) return;
    // four independent moves per thread and pass: the searches and the loads overlap
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t j0 = blockIdx.x * blockDim.x + i; j0 < total; j0 += 4 * stride) {
        uint32_t sp[4], dp[4], f[4], x[4];
        uint64_t k[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t j = j0 + q * stride;
            sp[q] = j < total ? plan_locate(s_src_start, s_src_prefix, n_src, j) : 0xFFFFFFFFu;
            dp[q] = j < total ? plan_locate(s_dst_start, s_dst_prefix, n_list, j) : 0u;
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (sp[q] == 0xFFFFFFFFu) continue;
            k[q] = keys[sp[q]];
            f[q] = first[sp[q]];
            x[q] = xy[sp[q]];
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (sp[q] == 0xFFFFFFFFu) continue;
            keys[dp[q]] = k[q];
            first[dp[q]] = f[q];
            xy[dp[q]] = x[q];
        }
    }
}

uint32_t map_words(uint64_t cells, bool count_rep) {
    return (uint32_t)(count_rep ? (cells + 15) / 16 : (cells + 31) / 32);
}
size_t slab_smem_bytes(uint64_t cells, bool count_rep) {
    return (size_t)kStages * kTile * 16 + (size_t)2 * kHash * 4 +
           (size_t)((map_words(cells, count_rep) + 3) & ~3u) * 4;
}
constexpr size_t kSmemLimit = 232448 - 1024;  // 227 KB opt-in maximum minus the static part

}  // namespace

size_t evk_slab_scratch_bytes(int sm_count) {
    return (size_t)sm_count * kCtasPerSm * sizeof(ChunkTail);
}
int evk_slab_ctas_per_sm() { return kCtasPerSm; }

bool evk_slab_supported(const evk_handle* h, const KeyParams& kp) {
    if (kp.keyfn != EVK_KEY_VOXEL || kp.vt <= 0 || h->n_events == 0) return false;
    if (kp.cells >= (1ull << (32 - kLogTile))) return false;  // packed (cell, index) word
    if (2 * kCtasPerSm * h->sm_count > kMaxList) return false;
    // (short tiles are padded with the event x = y = 0xFFFF, which must be outside the frame)
    if (kp.width > 65535 && kp.height > 65535) return false;
    return slab_smem_bytes(kp.cells, true) <= kSmemLimit;
}

int evk_downsample_slab(evk_handle* h, const KeyParams& kp, int count_repeated, bool* ok,
                        int* launches, bool sync, const unsigned long long* range,
                        const long long* t0_dev) {
    *ok = false;
    const int grid = h->sm_count * kCtasPerSm;
    SlabArgs a;
    a.kp = kp;
    a.ev = h->d_events;
    a.n = h->n_events;
    a.bin_start = h->d_bin_start;
    a.keys = h->d_keys;
    a.first = h->d_first;
    a.xy = h->d_xy;
    a.cnt = h->d_cnt;
    a.chunk_list = reinterpret_cast<ChunkTail*>(h->d_slab_scratch);
    a.first_offset = (uint32_t)h->shard_first;
    a.words = map_words(kp.cells, count_repeated != 0);
    a.max_bins = (uint32_t)h->max_bins;
    // a single CTA walks a bin sequentially: with few bins and many events the table is faster
    a.min_bins = h->n_events > (1u << 22) ? 32 : 1;
    a.out_cap = (uint32_t)(h->out_cap > 0xFFFFFFFFull ? 0xFFFFFFFFull : h->out_cap);
    a.range = range;
    a.t0_dev = t0_dev;
    const size_t smem = slab_smem_bytes(kp.cells, count_repeated != 0);
    const bool pow2 = kp.sx >= 0 && kp.sy >= 0;
    static void (*const kerns[8])(SlabArgs) = {
        k_slab_main<false, false, false>, k_slab_main<false, false, true>,
        k_slab_main<false, true, false>,  k_slab_main<false, true, true>,
        k_slab_main<true, false, false>,  k_slab_main<true, false, true>,
        k_slab_main<true, true, false>,   k_slab_main<true, true, true>};
    void (*kern)(SlabArgs) = kerns[(count_repeated ? 4 : 0) + (pow2 ? 2 : 0) + (kp.use_p ? 1 : 0)];
    EVK_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)kSmemLimit));
    k_slab_bins<<<2 * grid, 256, 0, h->stream>>>(a);
    EVK_CUDA(h, cudaGetLastError());
    evk_prof_rec(h, 5);
    kern<<<grid, kThreads, smem, h->stream>>>(a);
    EVK_CUDA(h, cudaGetLastError());
    evk_prof_rec(h, 6);
    k_slab_fix<<<grid, kFixThreads, 0, h->stream>>>(a.chunk_list, grid, h->d_cnt, h->d_keys,
                                                   h->d_first, h->d_xy, h->d_sticky);
    EVK_CUDA(h, cudaGetLastError());
    *launches += 3;
    if (!sync) return EVK_OK;
    // the verification flag decides whether the result stands
    EVK_CUDA(h, cudaMemcpyAsync(h->h_cnt, h->d_cnt, sizeof(DsCounters), cudaMemcpyDeviceToHost,
                                h->stream));
    EVK_CUDA(h, cudaStreamSynchronize(h->stream));
    *ok = h->h_cnt->slab_violation == 0 && h->h_cnt->overflow == 0;
    return EVK_OK;
}
