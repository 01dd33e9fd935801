// evk_comm.cu — multi-GPU path: one process per GPU, NCCL over NVLink 5 / NVSwitch.
//
// The reference is single-device (one clCreateContext / one in-order queue per program:
// ACCEL/store.cpp:237,277; KM/assign_to_centers2.c:171,207); this file is the build's addition
// (SURVEY.md 8e).  Events are sharded by contiguous index range.  Two exchange schemes:
//
//  EVK_OWNER_TIME_RANGE (default; time-ordered streams).  Ownership of a voxel key is decided by
//    its time bin: rank r owns every bin that starts inside its index shard.  Only the bin that
//    straddles a shard boundary has events on two GPUs, so rank r+1 sends the head of its shard
//    (a fixed block of `halo` events, one ncclSend/ncclRecv pair per neighbour) and rank r keeps
//    the received events that belong to its last bin; rank r+1 skips them.  After that every bin
//    is wholly on one GPU, the unchanged single-GPU kernels run, and no voxel has to move.  The
//    slab kernel's partition check doubles as the check that the stream really is time-ordered;
//    ranks agree on the outcome with one allreduce and otherwise fall through to
//  EVK_OWNER_MIX64 (general).  owner = mix64(key) % G.  Local downsample, bucket the voxels by
//    owner, exchange counts (allgather), NCCL all-to-all of (key, first index, xy, representative)
//    with grouped send/recv, then the owner merges what it received in its hash-owned table
//    (lowest global first index wins).
//
// K-means: each rank accumulates exact integer partial sums over its voxel shard; one
// ncclAllReduce of K x 5 u64 per iteration; every rank finalises identical centroids.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): inside a PyTorch process this resolves to the
// copy torch already loaded, so both share one NCCL.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

#include "evk_internal.cuh"

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi g_nccl;

bool load_nccl() {
    if (g_nccl.ok) return true;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return false;
    g_nccl.lib = lib;
#define SYM(field, name)                                                    \
    *reinterpret_cast<void**>(&g_nccl.field) = dlsym(lib, name);            \
    if (!g_nccl.field) return false
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllReduce, "ncclAllReduce");
    SYM(AllGather, "ncclAllGather");
    SYM(Broadcast, "ncclBroadcast");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    g_nccl.ok = true;
    return true;
}

constexpr int kMaxWorld = 64;
constexpr int kBlock = 256;

}  // namespace

// peer-memory mailbox of one rank (see "peer-memory exchange" below)
constexpr int kMailWords = 2048;  // >= 254 * 5 + 3
struct Mailbox {
    unsigned long long sums[2][kMaxWorld][kMailWords];  // [parity][from rank][word]
    unsigned long long sum_flag[2][kMaxWorld];          // seq of the partial above
    unsigned long long ready_flag;                      // next rank: "my events of step seq are loaded"
    unsigned long long keep_count;                      // next rank: how many of its leading events are mine
                                                        // (>= halo: its first bin does not end in the block)
    unsigned long long cent_flag;                       // rank 0: "centroids of step seq are below"
    unsigned long long seq;                             // my step counter
    unsigned long long err;                             // a wait timed out
    float cent[EVK_MAX_K * 2];
    // hash-owned exchange (k_mix_*): what every rank tells every other rank, by step parity
    unsigned long long mix_info[2][kMaxWorld][4];       // [src]: first bin, last bin, all-suspect, seq
    unsigned long long mix_cnt[2][kMaxWorld][kMaxWorld][2];  // [src][dst]: records (plain, suspect)
    unsigned long long mix_cnt_flag[2][kMaxWorld];      // seq of row src
    unsigned long long mix_done_flag[2][kMaxWorld];     // seq: src has finished writing into me
};

struct MixPeer {  // one rank's buffers as seen from this process
    uint64_t* keys;
    uint32_t* first;
    uint32_t* xy;
    evk_event* reps;
    uint64_t* sk;  // suspect staging
    uint32_t* sf;
    uint32_t* sx;
    evk_event* sr;
};

struct CommState {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    uint32_t halo = 1u << 18;              // events sent to the previous rank (4 MB)
    unsigned long long* d_stats = nullptr;  // [8 + kMaxWorld * (kMaxWorld + 2)]
    unsigned long long* h_stats = nullptr;  // pinned mirror
    // general-mode staging (lazy): voxels bucketed by owner / received from peers
    uint64_t *sk = nullptr, *rk = nullptr;
    uint32_t *sf = nullptr, *rf = nullptr, *sx = nullptr, *rx = nullptr;
    evk_event *sr = nullptr, *rr = nullptr;
    size_t stage_cap = 0;
    int last_mode = -1;
    // peer-memory exchange (NVLink P2P through CUDA IPC): mailboxes and event buffers of all ranks
    bool p2p = false;
    Mailbox* mail = nullptr;                        // mine (device)
    Mailbox* peer_mail[64] = {};                    // [r]: rank r's mailbox mapped into this process
    const evk_event* peer_events[64] = {};          // [r]: rank r's event buffer
    Mailbox** d_peer_mail = nullptr;                // the table above, on the device
    // hash-owned exchange over peer memory: the local downsample writes into (lk, lf, lx), the
    // bucket scatter writes every record straight into its owner's voxel shard (d_keys / d_first /
    // d_xy / d_reps, mapped below) or, for keys another rank may also hold, into the owner's
    // staging (rk / rf / rx / rr) for the merge
    bool mix_p2p = false;
    uint64_t* lk = nullptr;
    uint32_t *lf = nullptr, *lx = nullptr;
    struct MixPeer* d_mix_peers = nullptr;      // [world], device
    unsigned long long* d_mix = nullptr;        // device scratch (see MX_*)
    unsigned long long* h_mix = nullptr;        // pinned mirror of the tail
    void* ipc_opened[64][8] = {};
    cudaGraphExec_t step_exec = nullptr;  // the fused sharded step as one graph
    int step_owner = 0;                   // evk_downsample_kmeans_sharded_submit: owner mode,
    bool step_fusable = false;            // and whether the submission went out as the graph
    FusedKey step_key{};
    int step_launches = 0;
};

#define EVK_NCCL(h, expr)                                                                  \
    do {                                                                                   \
        ncclResult_t _r = (expr);                                                          \
        if (_r != ncclSuccess)                                                             \
            return evk_fail((h), EVK_ERR_COMM, "%s:%d %s: %s", __FILE__, __LINE__, #expr,  \
                            g_nccl.GetErrorString(_r));                                    \
    } while (0)

namespace {

// stats layout
enum { ST_FLAG = 0, ST_U = 1, ST_R = 2, ST_SKIP = 3, ST_KEEP = 4, ST_HIST = 8 };

// time-range halo: how many of my leading events belong to my first bin (they go to the previous
// rank) and how many of the received events belong to the sender's first bin (I keep them).
// One warp; run lengths by 32-way search (assumes time order; the slab kernel verifies the
// resulting partition).
__global__ void __launch_bounds__(32)
    k_halo_range(KeyParams kp, const evk_event* ev, uint32_t n_own, uint32_t halo, int rank,
                 int world, unsigned long long* stats) {
    const int lane = threadIdx.x;
    unsigned long long flag = 0, skip = 0, keep = 0;
    auto bin_of = [&](uint32_t i, bool& ok) -> uint64_t {
        const int64_t t = ev[i].t;
        ok = t >= kp.t0;
        return ok ? evk_tbin(kp, t) : 0;
    };
    auto run_len = [&](uint32_t start, uint32_t len) -> uint32_t {  // events sharing ev[start]'s bin
        bool ok;
        const uint64_t b0 = bin_of(start, ok);
        if (!ok) {
            flag = 1;
            return 0;
        }
        uint32_t lo = 0, hi = len;  // first offset whose bin differs
        while (lo < hi) {
            const uint32_t span = hi - lo;
            const uint32_t p = lo + (uint32_t)((uint64_t)span * (uint32_t)(lane + 1) / 33u);
            bool ok2;
            const uint64_t b = bin_of(start + p, ok2);
            const bool differs = !(ok2 && b == b0);
            const uint32_t m = __ballot_sync(0xffffffffu, differs);
            const int f = m ? __ffs(m) - 1 : 32;
            const uint32_t p_f = __shfl_sync(0xffffffffu, p, f < 32 ? f : 0);
            const uint32_t p_b = __shfl_sync(0xffffffffu, p, f > 0 ? f - 1 : 0);
            if (f < 32) hi = p_f;
            if (f > 0) lo = p_b + 1;
        }
        return lo;
    };
    if (n_own < halo) flag = 1;  // my block for the previous rank is not all real events
    if (!flag && rank > 0) {
        skip = run_len(0, halo);
        if (skip >= halo || skip >= n_own) flag = 1;  // first bin does not end inside the block
    }
    if (!flag && rank < world - 1) {
        keep = run_len(n_own, halo);
        if (keep >= halo) flag = 1;
    }
    if (lane == 0) {
        stats[ST_FLAG] = flag;
        stats[ST_SKIP] = skip;
        stats[ST_KEEP] = keep;
    }
}

// ---- peer-memory exchange ---------------------------------------------------------------------
// The fused sharded step moves three small things between GPUs: a 4 MB boundary block to the
// previous rank, K x 2 centroids from rank 0, and 2.6 KB of partial sums all-to-all.  Through NCCL
// each costs a kernel launch and tens of microseconds of latency; here the ranks read and write
// each other's memory directly over NVLink (buffers mapped with CUDA IPC) and synchronise with
// sequence-number flags, so each exchange is part of an ordinary kernel of the step:
//   k_p2p_tick     seq++; tell the previous rank "my events are loaded"
//   k_p2p_push_cent rank 0 (beside its downsample): centroids into every mailbox
//   k_p2p_pull     wait for the next rank's flag, PULL its boundary block behind my own events
//   k_p2p_cent     wait for rank 0's centroids, copy them in
//   k_p2p_allreduce push my partial sums into every rank's mailbox, wait for everybody's, add up
// Every wait is bounded (about nine seconds of spinning): a timeout raises Mailbox::err, the pass is
// abandoned by all ranks through the give-up flag, and the communicator falls back to NCCL.

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// spin until *flag >= want; false on timeout
__device__ __forceinline__ bool wait_flag(const unsigned long long* flag, unsigned long long want) {
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < want) {
        if (clock64() - t0 > (1ll << 34)) return false;  // about nine seconds: a peer is gone
        __nanosleep(100);
    }
    return true;
}

__global__ void k_p2p_tick(Mailbox* mine, Mailbox* const* peers, int rank) {
    const unsigned long long seq = mine->seq + 1;
    mine->seq = seq;
    if (rank > 0) st_release_sys(&peers[rank - 1]->ready_flag, seq);
}

// The fused step's tick: one warp.  seq++; the events at the head of my shard that still belong to the
// previous rank's last time bin are counted HERE (a 32-way search over my own, local events) and the
// count travels with the flag, so the receiver pulls exactly those events and needs no search of its
// own (k_halo_range's two searches would otherwise sit between the pull and the downsample).
__global__ void __launch_bounds__(1024)
    k_p2p_tick_range(KeyParams kp, const evk_event* ev, uint32_t n_own, uint32_t halo, int rank,
                     int world, Mailbox* mine, Mailbox* const* peers, unsigned long long* stats) {
    // One CTA of 1024 threads: a 1024-way search finds the end of the first bin's run inside the
    // boundary block in two rounds of (independent) global loads -- the search sits in front of the
    // flag the previous rank's pull is waiting for, and of this rank's own downsample.
    __shared__ unsigned int s_first;  // lowest thread whose probe lies behind the run
    const unsigned int tid = threadIdx.x, NT = blockDim.x;
    const unsigned long long seq = mine->seq + 1;
    __syncthreads();
    if (tid == 0) mine->seq = seq;
    unsigned long long flag = n_own < halo ? 1 : 0, skip = 0;
    if (!flag && rank > 0) {  // (uniform over the CTA)
        const int64_t t0v = ev[0].t;
        if (t0v < kp.t0) {
            flag = 1;
        } else {
            const uint64_t b0 = evk_tbin(kp, t0v);
            uint32_t lo = 0, hi = halo;  // the first offset whose bin differs lies in [lo, hi]
            while (lo < hi) {
                const uint32_t span = hi - lo;
                const bool dense = span <= NT;  // every offset of the range has its own thread
                const uint32_t p = dense ? lo + tid
                                         : lo + (uint32_t)((uint64_t)span * (tid + 1) / (NT + 1));
                if (tid == 0) s_first = NT;
                __syncthreads();
                if (!dense || tid < span) {
                    const int64_t t = ev[p].t;
                    if (!(t >= kp.t0 && evk_tbin(kp, t) == b0)) atomicMin(&s_first, tid);
                }
                __syncthreads();
                const uint32_t f = s_first;
                __syncthreads();  // (s_first is reset at the top of the next round)
                if (dense) {
                    lo = hi = f < NT ? lo + f : hi;
                } else {
                    // probe of thread q: lo + span * (q + 1) / (NT + 1), increasing in q
                    const uint32_t p_f = lo + (uint32_t)((uint64_t)span * (f + 1) / (NT + 1));
                    const uint32_t p_b = lo + (uint32_t)((uint64_t)span * f / (NT + 1));  // thread f - 1
                    const uint32_t lo_old = lo;
                    if (f < NT) hi = p_f;
                    if (f > 0) lo = p_b + 1;
                    (void)lo_old;
                }
            }
            skip = lo;
            if (skip >= halo || skip >= n_own) flag = 1;  // first bin does not end inside the block
        }
    }
    if (tid == 0) {
        stats[ST_FLAG] = flag;
        stats[ST_SKIP] = flag ? 0 : skip;
        stats[ST_KEEP] = 0;
        if (rank > 0) {
            peers[rank - 1]->keep_count = flag ? (unsigned long long)halo : skip;
            st_release_sys(&peers[rank - 1]->ready_flag, seq);
        }
    }
}

// the events of the next rank that belong to my last bin -> behind my own: exactly keep_count of them
__global__ void __launch_bounds__(256)
    k_p2p_pull_exact(Mailbox* mine, const evk_event* next_events, evk_event* dst, uint32_t halo,
                     unsigned long long* stats) {
    __shared__ unsigned long long s_keep;
    if (threadIdx.x == 0) {
        unsigned long long keep = halo;  // (time-out: give up)
        if (wait_flag(&mine->ready_flag, mine->seq))
            keep = *reinterpret_cast<volatile unsigned long long*>(&mine->keep_count);
        else
            mine->err = 1;
        s_keep = keep;
        if (blockIdx.x == 0) {
            if (keep >= halo) stats[ST_FLAG] = 1;
            else stats[ST_KEEP] = keep;
        }
    }
    __syncthreads();
    const unsigned long long keep = s_keep;
    if (keep >= halo) return;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < keep)
        reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(next_events)[i];
}

// rank 0, beside its downsample: the centroids the walk kernel has just written go to every rank
__global__ void __launch_bounds__(256)
    k_p2p_push_cent(Mailbox* mine, Mailbox* const* peers, int world, const float* cent, int K) {
    const unsigned long long seq = mine->seq;
    for (int r = 1; r < world; r++)
        for (int i = threadIdx.x; i < 2 * K; i += blockDim.x) peers[r]->cent[i] = cent[i];
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x > 0 && (int)threadIdx.x < world)
        st_release_sys(&peers[threadIdx.x]->cent_flag, seq);
}

// the next rank's first `halo` events -> behind my own (a pull: the reader knows where they go)
__global__ void __launch_bounds__(256)
    k_p2p_pull(Mailbox* mine, const evk_event* next_events, evk_event* dst, uint32_t halo) {
    __shared__ int s_ok;
    if (threadIdx.x == 0) {
        s_ok = wait_flag(&mine->ready_flag, mine->seq) ? 1 : 0;
        if (!s_ok) mine->err = 1;
    }
    __syncthreads();
    if (!s_ok) return;
    const uint4* src = reinterpret_cast<const uint4*>(next_events);
    uint4* d = reinterpret_cast<uint4*>(dst);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < halo; i += gridDim.x * blockDim.x)
        d[i] = src[i];
}

__global__ void __launch_bounds__(256) k_p2p_cent(Mailbox* mine, float* cent, int K) {
    __shared__ int s_ok;
    if (threadIdx.x == 0) {
        s_ok = wait_flag(&mine->cent_flag, mine->seq) ? 1 : 0;
        if (!s_ok) mine->err = 1;
    }
    __syncthreads();
    if (!s_ok) return;
    for (int i = threadIdx.x; i < 2 * K; i += blockDim.x) cent[i] = __ldcg(&mine->cent[i]);
}

// one-shot allreduce (sum) of acc[0..n): every rank pushes its partial into every mailbox
__global__ void __launch_bounds__(1024)
    k_p2p_allreduce(Mailbox* mine, Mailbox* const* peers, int rank, int world,
                    unsigned long long* acc, int n) {
    __shared__ int s_ok;
    const unsigned long long seq = mine->seq;
    const int par = (int)(seq & 1);
    if (threadIdx.x == 0) s_ok = 1;
    for (int r = 0; r < world; r++)
        for (int i = threadIdx.x; i < n; i += blockDim.x) peers[r]->sums[par][rank][i] = acc[i];
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < world) st_release_sys(&peers[threadIdx.x]->sum_flag[par][rank], seq);
    if ((int)threadIdx.x < world && !wait_flag(&mine->sum_flag[par][threadIdx.x], seq)) {
        s_ok = 0;
        mine->err = 1;
    }
    __syncthreads();
    if (!s_ok) {
        if (threadIdx.x == 0) acc[n - 3] = 1ull << 32;  // give-up flag of the step (k_pack_step_stats)
        return;
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        unsigned long long sum = 0;
        for (int r = 0; r < world; r++) sum += __ldcg(&mine->sums[par][r][i]);  // written by peers
        acc[i] = sum;
    }
}

// fused sharded step: [K * 5 + 0] = any reason to abandon the pass, [+1] = voxels, [+2] = repeated
// (a peer-memory wait that timed out on this rank adds 2^32: the host then drops to NCCL everywhere)
__global__ void k_pack_step_stats(const DsCounters* cnt, const unsigned long long* range,
                                  unsigned long long found_want, int check_found,
                                  const Mailbox* mail, unsigned long long* tail) {
    unsigned long long flag = cnt->slab_violation | cnt->overflow | range[ST_FLAG];
    if (check_found && cnt->scratch[4] != found_want) flag |= 1;
    tail[0] = (flag ? 1 : 0) + ((mail && mail->err) ? (1ull << 32) : 0ull);
    tail[1] = cnt->n_unique;
    tail[2] = cnt->n_repeated;
}

// The fused time-range step's tail as ONE kernel (three launches and a copy node in round 1):
// pack the step's stats behind the partial sums, the one-shot allreduce over peer memory, the
// allreduced stats into the counters block (it travels to the host with the step's one D2H copy),
// the centroid update.
__global__ void __launch_bounds__(1024)
    k_p2p_step_tail(Mailbox* mine, Mailbox* const* peers, int rank, int world,
                    unsigned long long* acc, int n, DsCounters* cnt, const unsigned long long* range,
                    unsigned long long found_want, int check_found, KmLaunch kl, float* cent,
                    unsigned long long* counts, float* shift) {
    __shared__ int s_ok;
    const unsigned long long seq = mine->seq;
    const int par = (int)(seq & 1);
    if (threadIdx.x == 0) {
        s_ok = 1;
        unsigned long long flag = cnt->slab_violation | cnt->overflow | range[ST_FLAG];
        if (check_found && cnt->scratch[4] != found_want) flag |= 1;
        acc[n - 3] = (flag ? 1 : 0) + (mine->err ? (1ull << 32) : 0ull);
        acc[n - 2] = cnt->n_unique;
        acc[n - 1] = cnt->n_repeated;
    }
    __syncthreads();
    for (int r = 0; r < world; r++)
        for (int i = threadIdx.x; i < n; i += blockDim.x) peers[r]->sums[par][rank][i] = acc[i];
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < world) st_release_sys(&peers[threadIdx.x]->sum_flag[par][rank], seq);
    if ((int)threadIdx.x < world && !wait_flag(&mine->sum_flag[par][threadIdx.x], seq)) {
        s_ok = 0;
        mine->err = 1;
    }
    __syncthreads();
    if (s_ok) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            unsigned long long sum = 0;
            for (int r = 0; r < world; r++) sum += __ldcg(&mine->sums[par][r][i]);  // written by peers
            acc[i] = sum;
        }
    } else if (threadIdx.x == 0) {
        acc[n - 3] = 1ull << 32;  // give up: the host drops to NCCL everywhere
    }
    __syncthreads();
    if (threadIdx.x < 3) cnt->step_tail[threadIdx.x] = acc[n - 3 + threadIdx.x];
    evk_km_finalise_body(kl, cent, acc, counts, shift);
}

// owner of a key: the top 32 bits of mix64(key ^ golden) range-reduced to [0, G) by a multiply
// (a 64-bit modulo costs ~40 instructions per record; the rule is restated in sharding.py)
// the same for the hash-owned step: the exchange's error word instead of the halo range flag
__global__ void k_pack_mix_stats(const DsCounters* cnt, const unsigned long long* mx_tail,
                                 unsigned long long found_want, int check_found,
                                 const Mailbox* mail, unsigned long long* tail) {
    unsigned long long flag = mx_tail[2];
    if (check_found && cnt->scratch[4] != found_want) flag |= 1;
    tail[0] = (flag ? 1 : 0) + ((mail && mail->err) ? (1ull << 32) : 0ull);
    tail[1] = cnt->n_unique;
    tail[2] = 0;
}

__device__ __forceinline__ int owner_of(uint64_t key, int world) {
    const uint32_t hi = (uint32_t)(evk_mix64(key ^ 0x9E3779B97F4A7C15ull) >> 32);
    return (int)__umulhi(hi, (uint32_t)world);
}

__global__ void __launch_bounds__(kBlock)
    k_owner_hist(const uint64_t* __restrict__ keys, size_t n, int world, unsigned long long* hist) {
    __shared__ unsigned int s_h[kMaxWorld];
    for (int i = threadIdx.x; i < world; i += kBlock) s_h[i] = 0;
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        atomicAdd(&s_h[owner_of(keys[i], world)], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < world; i += kBlock)
        if (s_h[i]) atomicAdd(&hist[i], (unsigned long long)s_h[i]);
}

__global__ void __launch_bounds__(kBlock)
    k_owner_scatter(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ first,
                    const uint32_t* __restrict__ xy, const evk_event* __restrict__ ev0, size_t n,
                    int world, const unsigned long long* __restrict__ send_off,
                    unsigned long long* cursor, uint64_t* sk, uint32_t* sf, uint32_t* sx,
                    evk_event* sr) {
    // block-aggregated bucket cursors: one global atomic per (block tile, owner)
    __shared__ unsigned int s_cnt[kMaxWorld];
    __shared__ unsigned long long s_base[kMaxWorld];
    const size_t n_pad = (n + kBlock - 1) / kBlock * kBlock;
    for (size_t i0 = (size_t)blockIdx.x * kBlock; i0 < n_pad; i0 += (size_t)gridDim.x * kBlock) {
        const size_t i = i0 + threadIdx.x;
        for (int o = threadIdx.x; o < world; o += kBlock) s_cnt[o] = 0;
        __syncthreads();
        uint64_t k = 0;
        int o = 0;
        unsigned int local = 0;
        if (i < n) {
            k = keys[i];
            o = owner_of(k, world);
            local = atomicAdd(&s_cnt[o], 1u);
        }
        __syncthreads();
        for (int q = threadIdx.x; q < world; q += kBlock)
            s_base[q] = s_cnt[q] ? atomicAdd(&cursor[q], (unsigned long long)s_cnt[q]) : 0ull;
        __syncthreads();
        if (i < n) {
            const size_t p = (size_t)send_off[o] + (size_t)s_base[o] + local;
            const uint32_t f = first[i];
            sk[p] = k;
            sf[p] = f;
            sx[p] = xy[i];
            sr[p] = ev0[f];  // ev0 is indexable by global index
        }
        __syncthreads();
    }
}

__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

// owner-side merge, pass 1: key -> lowest global first index
__global__ void __launch_bounds__(kBlock)
    k_merge_insert(const uint64_t* __restrict__ rk, const uint32_t* __restrict__ rf, size_t m,
                   uint64_t* tkeys, uint32_t* tfirst, uint64_t mask) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        const uint64_t key = rk[i];
        uint64_t slot = evk_mix64(key) & mask;
        for (;;) {
            uint64_t cur = ld_relaxed_u64(tkeys + slot);
            if (cur == EVK_EMPTY_KEY)
                cur = atomicCAS((unsigned long long*)(tkeys + slot), EVK_EMPTY_KEY, key);
            if (cur == EVK_EMPTY_KEY || cur == key) {
                atomicMin(tfirst + slot, rf[i]);
                break;
            }
            slot = (slot + 1) & mask;
        }
    }
}

// pass 2: the record whose first index survived in the table is the voxel; compact the winners
__global__ void __launch_bounds__(kBlock)
    k_merge_emit(const uint64_t* __restrict__ rk, const uint32_t* __restrict__ rf,
                 const uint32_t* __restrict__ rx, const evk_event* __restrict__ rr, size_t m,
                 const uint64_t* __restrict__ tkeys, const uint32_t* __restrict__ tfirst,
                 uint64_t mask, uint64_t* keys, uint32_t* first, uint32_t* xy, evk_event* reps,
                 uint32_t* slot_out, DsCounters* cnt) {
    const int lane = threadIdx.x & 31;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t m_pad = (m + 31) & ~(size_t)31;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < m_pad; i += stride) {
        bool win = false;
        uint64_t key = 0;
        uint32_t f = 0;
        if (i < m) {
            key = rk[i];
            f = rf[i];
            uint64_t slot = evk_mix64(key) & mask;
            while (tkeys[slot] != key) slot = (slot + 1) & mask;
            win = tfirst[slot] == f;
            slot_out[i] = (uint32_t)slot;  // pass 3 resets exactly this slot
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, win);
        unsigned long long base = 0;
        if (lane == 0 && bal) base = atomicAdd(&cnt->n_unique, (unsigned long long)__popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (win) {
            const size_t o = (size_t)base + __popc(bal & ((1u << lane) - 1u));
            keys[o] = key;
            first[o] = f;
            xy[o] = rx[i];
            reps[o] = rr[i];
        }
    }
}

// pass 3: leave the table clean
__global__ void __launch_bounds__(kBlock)
    k_merge_reset(const uint32_t* __restrict__ slots, size_t m, uint64_t* tkeys, uint32_t* tfirst) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        const uint32_t slot = slots[i];  // records of one key write the same values: benign
        tkeys[slot] = EVK_EMPTY_KEY;
        tfirst[slot] = EVK_EMPTY_IDX;
    }
}

// ---- hash-owned exchange over peer memory ------------------------------------------------------
// owner = mix64(key) % G (the north star's design).  Every rank downsamples its own index shard,
// then writes each voxel record STRAIGHT into its owner's voxel shard over NVLink: no staging copy,
// no NCCL call, no host round trip between the local downsample and the last kernel.
//   k_mix_publish_info  my first / last time bin (or "unordered") into every mailbox
//   k_mix_hist          records per owner, split into plain and SUSPECT ones: a key can only exist
//                       on two ranks when its time bin straddles a shard boundary (index-sharded,
//                       time-ordered stream), i.e. when its bin is some rank's first or last bin;
//                       a rank whose shard is not time-ordered marks everything suspect
//   k_mix_publish_counts / k_mix_offsets   the G x G count matrix travels through the mailboxes;
//                       every rank derives where its records start in every owner's arrays
//   k_mix_scatter       block-aggregated bucket cursors; plain records go into the owner's
//                       d_keys / d_first / d_xy (/ d_reps), suspects into the owner's staging
//   k_mix_done          "my writes into you are complete" to everyone, wait for everyone's
//   k_merge_*           the owner resolves the (few) suspects in its hash-owned table: lowest
//                       global first index wins; winners are appended behind the plain records
enum { MX_HIST = 0, MX_OFF_PLAIN = 2 * kMaxWorld, MX_OFF_SUSP = 3 * kMaxWorld,
       MX_CUR = 4 * kMaxWorld, MX_TAIL = 6 * kMaxWorld, MX_WORDS = 6 * kMaxWorld + 8 };
// tail: [0] plain records I own, [1] suspects I own, [2] error (capacity / time-out)

__global__ void __launch_bounds__(64)
    k_mix_publish_info(Mailbox* mine, Mailbox* const* peers, int rank, int world,
                       unsigned long long tb_first, unsigned long long tb_last,
                       unsigned long long all_suspect) {
    const unsigned long long seq = mine->seq;
    const int par = (int)(seq & 1);
    const int r = threadIdx.x;
    if (r < world) {
        unsigned long long* info = peers[r]->mix_info[par][rank];
        info[0] = tb_first;
        info[1] = tb_last;
        info[2] = all_suspect;
        __threadfence_system();
        st_release_sys(&info[3], seq);
    }
}

// exact key / cells without a 64-bit division: q = hi64(key * ceil(2^64 / cells)) is the quotient
// or the quotient + 1
struct MixDiv {
    uint64_t cells, magic;    // magic = floor(2^64 / cells) + 1 (0 when cells == 1)
    uint64_t lo_key, hi_key;  // keys in [lo_key, hi_key) lie strictly inside this rank's bin range
};
__device__ __forceinline__ uint64_t mix_tbin(uint64_t key, const MixDiv& d) {
    if (d.cells <= 1) return key;
    uint64_t q = __umul64hi(key, d.magic);
    if (q * d.cells > key) q--;
    return q;
}
__device__ __forceinline__ bool mix_suspect(uint64_t key, const MixDiv& d, const uint64_t* s_bins,
                                            int n_bins, bool all) {
    if (all) return true;
    // a time-ordered shard can share a bin with another rank only in its own first or last bin:
    // keys strictly between them need no division and no search
    if (key >= d.lo_key && key < d.hi_key) return false;
    const uint64_t tb = mix_tbin(key, d);
    bool s = false;
    for (int i = 0; i < n_bins; i++) s |= s_bins[i] == tb;
    return s;
}

// shared by the histogram and the scatter: wait for every rank's info, list the suspect bins
__device__ __forceinline__ bool mix_load_info(Mailbox* mine, int world, uint64_t* s_bins,
                                              int* s_flags) {
    const unsigned long long seq = mine->seq;
    const int par = (int)(seq & 1);
    if (threadIdx.x == 0) {
        s_flags[0] = 0;  // all suspect
        s_flags[1] = 1;  // ok
    }
    __syncthreads();
    if ((int)threadIdx.x < world) {
        unsigned long long* info = mine->mix_info[par][threadIdx.x];
        if (!wait_flag(&info[3], seq)) {
            s_flags[1] = 0;
            mine->err = 1;
        } else {
            s_bins[2 * threadIdx.x] = __ldcg(&info[0]);
            s_bins[2 * threadIdx.x + 1] = __ldcg(&info[1]);
            if (__ldcg(&info[2])) s_flags[0] = 1;
        }
    }
    __syncthreads();
    return s_flags[1] != 0;
}

// bucket of every record (2 * owner + suspect) -> bkt[], and the bucket sizes
__global__ void __launch_bounds__(kBlock)
    k_mix_hist(const uint64_t* __restrict__ keys, size_t n, int world, MixDiv dv, Mailbox* mine,
               unsigned long long* mx, uint8_t* bkt) {
    __shared__ uint64_t s_bins[2 * kMaxWorld];
    __shared__ int s_flags[2];
    __shared__ unsigned int s_h[2 * kMaxWorld];
    if (!mix_load_info(mine, world, s_bins, s_flags)) {
        if (threadIdx.x == 0) mx[MX_TAIL + 2] = 1;
        return;
    }
    for (int i = threadIdx.x; i < 2 * world; i += kBlock) s_h[i] = 0;
    __syncthreads();
    const bool all = s_flags[0] != 0;
    const int lane = threadIdx.x & 31;
    constexpr int kU = 4;  // keys in flight per thread (one load per iteration is latency-bound)
    const size_t tile = (size_t)kBlock * kU;
    const size_t n_tiles = (n + tile - 1) / tile;
    for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        uint64_t k[kU];
#pragma unroll
        for (int u = 0; u < kU; u++) {
            const size_t i = t * tile + (size_t)u * kBlock + threadIdx.x;
            k[u] = i < n ? keys[i] : 0;
        }
#pragma unroll
        for (int u = 0; u < kU; u++) {
            const size_t i = t * tile + (size_t)u * kBlock + threadIdx.x;
            int b = -1;
            if (i < n) {
                b = 2 * owner_of(k[u], world) + (mix_suspect(k[u], dv, s_bins, 2 * world, all) ? 1 : 0);
                bkt[i] = (uint8_t)b;
            }
            // one shared atomic per (warp, bucket present)
            uint32_t rest = __ballot_sync(0xffffffffu, b >= 0);
            while (rest) {
                const int b0 = __shfl_sync(0xffffffffu, b, __ffs(rest) - 1);
                const uint32_t m = __ballot_sync(0xffffffffu, b == b0);
                if (lane == __ffs(m) - 1) atomicAdd(&s_h[b0], (unsigned int)__popc(m));
                rest &= ~m;
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * world; i += kBlock)
        if (s_h[i]) atomicAdd(&mx[MX_HIST + i], (unsigned long long)s_h[i]);
}

// my row of the count matrix into every mailbox
__global__ void __launch_bounds__(kBlock)
    k_mix_publish_counts(Mailbox* mine, Mailbox* const* peers, int rank, int world,
                         const unsigned long long* mx) {
    const unsigned long long seq = mine->seq;
    const int par = (int)(seq & 1);
    for (int i = threadIdx.x; i < world * 2 * world; i += kBlock) {
        const int r = i / (2 * world), j = i % (2 * world);
        peers[r]->mix_cnt[par][rank][j >> 1][j & 1] = mx[MX_HIST + j];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < world) st_release_sys(&peers[threadIdx.x]->mix_cnt_flag[par][rank], seq);
}

// where my records start in every owner's arrays; how many records I own
__global__ void __launch_bounds__(64)
    k_mix_offsets(Mailbox* mine, int rank, int world, unsigned long long cap_plain,
                  unsigned long long cap_susp, unsigned long long* mx, DsCounters* cnt) {
    __shared__ int s_ok;
    const unsigned long long seq = mine->seq;
    const int par = (int)(seq & 1);
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    if ((int)threadIdx.x < world && !wait_flag(&mine->mix_cnt_flag[par][threadIdx.x], seq)) {
        s_ok = 0;
        mine->err = 1;
    }
    __syncthreads();
    if (!s_ok) {
        if (threadIdx.x == 0) mx[MX_TAIL + 2] = 1;
        return;
    }
    const int d = threadIdx.x;  // one owner per thread
    if (d < world) {
        unsigned long long op = 0, os = 0, tp = 0, ts = 0;
        for (int s = 0; s < world; s++) {
            const unsigned long long cp = __ldcg(&mine->mix_cnt[par][s][d][0]);
            const unsigned long long cs = __ldcg(&mine->mix_cnt[par][s][d][1]);
            if (s < rank) {
                op += cp;
                os += cs;
            }
            tp += cp;
            ts += cs;
        }
        mx[MX_OFF_PLAIN + d] = op;
        mx[MX_OFF_SUSP + d] = os;
        if (d == rank) {
            mx[MX_TAIL + 0] = tp;
            mx[MX_TAIL + 1] = ts;
            cnt->n_unique = tp;  // the merge appends its winners behind the plain records
        }
        // every rank checks every owner's capacity: all of them see the same matrix and agree
        if (tp + ts > cap_plain || ts > cap_susp) mx[MX_TAIL + 2] = 1;
    }
}

// Bucket scatter into the owners' memory.  A block sorts a tile of kScTile records by bucket in
// shared memory, claims one range per bucket with one global atomic each, and then writes the
// tile out with consecutive threads on consecutive slots: every bucket's run is one contiguous,
// coalesced burst over NVLink instead of 8 / 4 / 4-byte stores scattered across the owner's arrays.
constexpr int kScPer = 2, kScTile = kBlock * kScPer;
constexpr int kScCtas = 8;  // resident CTAs per SM (registers capped accordingly)
__global__ void __launch_bounds__(kBlock, kScCtas)
    k_mix_scatter(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ first,
                  const uint32_t* __restrict__ xy, const evk_event* __restrict__ ev0,
                  const uint8_t* __restrict__ bkt, size_t n, int world,
                  const MixPeer* __restrict__ peers, unsigned long long* mx, int want_reps,
                  int local_only) {
    __shared__ unsigned int s_cnt[2 * kMaxWorld], s_start[2 * kMaxWorld + 1];
    __shared__ unsigned long long s_dst[2 * kMaxWorld];  // first slot of this tile's run per bucket
    __shared__ uint64_t s_k[kScTile];
    __shared__ uint32_t s_f[kScTile], s_x[kScTile];
    __shared__ uint8_t s_b[kScTile];
    if (mx[MX_TAIL + 2]) return;  // capacity error or time-out: nothing is written anywhere
    const int nb = 2 * world;
    const size_t n_tiles = (n + kScTile - 1) / kScTile;
    for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (int o = threadIdx.x; o < nb; o += kBlock) s_cnt[o] = 0;
        __syncthreads();
        uint64_t k[kScPer];
        uint32_t f[kScPer], x[kScPer], loc[kScPer];
        int b[kScPer];
        const int lane = threadIdx.x & 31;
        const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
        for (int j = 0; j < kScPer; j++) {
            const size_t i = t * kScTile + (size_t)j * kBlock + threadIdx.x;
            b[j] = -1;
            loc[j] = 0;
            if (i < n) {
                k[j] = keys[i];
                f[j] = first[i];
                x[j] = xy[i];
                b[j] = bkt[i];
            }
            // slot inside the tile's bucket: one shared atomic per (warp, bucket present)
            uint32_t rest = __ballot_sync(0xffffffffu, b[j] >= 0);
            while (rest) {
                const int b0 = __shfl_sync(0xffffffffu, b[j], __ffs(rest) - 1);
                const uint32_t m = __ballot_sync(0xffffffffu, b[j] == b0);
                unsigned int base = 0;
                if (lane == __ffs(m) - 1) base = atomicAdd(&s_cnt[b0], (unsigned int)__popc(m));
                base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
                if (b[j] == b0) loc[j] = base + __popc(m & lt);
                rest &= ~m;
            }
        }
        __syncthreads();
        if (threadIdx.x < 32) {  // exclusive scan of the bucket sizes (<= 128 buckets: 4 per lane)
            unsigned int v[4], sum = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int o = threadIdx.x * 4 + q;
                v[q] = o < nb ? s_cnt[o] : 0u;
                sum += v[q];
            }
            unsigned int inc = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned int u = __shfl_up_sync(0xffffffffu, inc, d);
                if ((int)threadIdx.x >= d) inc += u;
            }
            unsigned int run = inc - sum;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int o = threadIdx.x * 4 + q;
                if (o < nb) {
                    s_start[o] = run;
                    // the tile's run of bucket o in its owner's arrays
                    const unsigned long long off = mx[((o & 1) ? MX_OFF_SUSP : MX_OFF_PLAIN) + (o >> 1)];
                    s_dst[o] = off + (v[q] ? atomicAdd(&mx[MX_CUR + o], (unsigned long long)v[q]) : 0ull);
                }
                run += v[q];
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kScPer; j++) {
            if (b[j] < 0) continue;
            const unsigned int slot = s_start[b[j]] + loc[j];
            s_k[slot] = k[j];
            s_f[slot] = f[j];
            s_x[slot] = x[j];
            s_b[slot] = (uint8_t)b[j];
        }
        __syncthreads();
        const size_t left = n - t * kScTile;
        const unsigned int in_tile = left < (size_t)kScTile ? (unsigned int)left : (unsigned int)kScTile;
#pragma unroll
        for (int j = 0; j < kScPer; j++) {
            const unsigned int slot = j * kBlock + threadIdx.x;
            if (slot >= in_tile) continue;
            const int bb = s_b[slot];
            const MixPeer& pr = peers[local_only >= 0 ? local_only : (bb >> 1)];
            const size_t p = (size_t)s_dst[bb] + (slot - s_start[bb]);
            const uint32_t ff = s_f[slot];
            if (bb & 1) {
                pr.sk[p] = s_k[slot];
                pr.sf[p] = ff;
                pr.sx[p] = s_x[slot];
                if (want_reps) pr.sr[p] = ev0[ff];  // ev0 is indexable by global index
            } else {
                pr.keys[p] = s_k[slot];
                pr.first[p] = ff;
                pr.xy[p] = s_x[slot];
                if (want_reps) pr.reps[p] = ev0[ff];
            }
        }
        __syncthreads();
    }
    // (no fence here: a system-scope fence per thread costs more than the scatter itself; the
    // kernel boundary orders these writes before k_mix_done, which fences once and raises the flags)
}

__global__ void __launch_bounds__(64)
    k_mix_done(Mailbox* mine, Mailbox* const* peers, int rank, int world, unsigned long long* mx) {
    __shared__ int s_ok;
    const unsigned long long seq = mine->seq;
    const int par = (int)(seq & 1);
    if (threadIdx.x == 0) s_ok = 1;
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < world) {
        st_release_sys(&peers[threadIdx.x]->mix_done_flag[par][rank], seq);
        if (!wait_flag(&mine->mix_done_flag[par][threadIdx.x], seq)) {
            s_ok = 0;
            mine->err = 1;
        }
    }
    __syncthreads();
    if (!s_ok && threadIdx.x == 0) mx[MX_TAIL + 2] = 1;
}

// the suspect merge with its record count read on the device (mx[MX_TAIL + 1])
__global__ void __launch_bounds__(kBlock)
    k_mix_merge_insert(const uint64_t* __restrict__ rk, const uint32_t* __restrict__ rf,
                       const unsigned long long* mx, uint64_t* tkeys, uint32_t* tfirst,
                       uint64_t mask) {
    const size_t m = mx[MX_TAIL + 2] ? 0 : (size_t)mx[MX_TAIL + 1];
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        const uint64_t key = rk[i];
        uint64_t slot = evk_mix64(key) & mask;
        for (;;) {
            uint64_t cur = ld_relaxed_u64(tkeys + slot);
            if (cur == EVK_EMPTY_KEY)
                cur = atomicCAS((unsigned long long*)(tkeys + slot), EVK_EMPTY_KEY, key);
            if (cur == EVK_EMPTY_KEY || cur == key) {
                atomicMin(tfirst + slot, rf[i]);
                break;
            }
            slot = (slot + 1) & mask;
        }
    }
}
__global__ void __launch_bounds__(kBlock)
    k_mix_merge_emit(const uint64_t* __restrict__ rk, const uint32_t* __restrict__ rf,
                     const uint32_t* __restrict__ rx, const evk_event* __restrict__ rr,
                     const unsigned long long* mx, const uint64_t* __restrict__ tkeys,
                     const uint32_t* __restrict__ tfirst, uint64_t mask, uint64_t* keys,
                     uint32_t* first, uint32_t* xy, evk_event* reps, uint32_t* slot_out,
                     DsCounters* cnt, int want_reps) {
    const size_t m = mx[MX_TAIL + 2] ? 0 : (size_t)mx[MX_TAIL + 1];
    const int lane = threadIdx.x & 31;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t m_pad = (m + 31) & ~(size_t)31;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < m_pad; i += stride) {
        bool win = false;
        uint64_t key = 0;
        uint32_t f = 0;
        if (i < m) {
            key = rk[i];
            f = rf[i];
            uint64_t slot = evk_mix64(key) & mask;
            while (tkeys[slot] != key) slot = (slot + 1) & mask;
            win = tfirst[slot] == f;
            slot_out[i] = (uint32_t)slot;
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, win);
        unsigned long long base = 0;
        if (lane == 0 && bal) base = atomicAdd(&cnt->n_unique, (unsigned long long)__popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (win) {
            const size_t o = (size_t)base + __popc(bal & ((1u << lane) - 1u));
            keys[o] = key;
            first[o] = f;
            xy[o] = rx[i];
            if (want_reps) reps[o] = rr[i];
        }
    }
}
__global__ void __launch_bounds__(kBlock)
    k_mix_merge_reset(const uint32_t* __restrict__ slots, const unsigned long long* mx,
                      uint64_t* tkeys, uint32_t* tfirst) {
    const size_t m = mx[MX_TAIL + 2] ? 0 : (size_t)mx[MX_TAIL + 1];
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        const uint32_t slot = slots[i];
        tkeys[slot] = EVK_EMPTY_KEY;
        tfirst[slot] = EVK_EMPTY_IDX;
    }
}

// flag = sum over ranks of (1 if the basic set-up failed) + (2 if the hash-owned set-up failed)
bool ok_all(unsigned long long sum, int world, int what) {
    // decode: each rank contributes 0..3; without carries between the two bits this needs them
    // separated, so ranks add 1 and 2 * (world + 1) instead (see the callers)
    (void)world;
    return what == 1 ? (sum & 0xFFFFull) == 0 : (sum >> 16) == 0;
}

int grid_for(size_t n, int sm) {
    size_t need = (n + kBlock - 1) / kBlock;
    size_t cap = (size_t)sm * 8;
    return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

int stage_alloc(evk_handle* h, CommState* c) {
    if (c->sk) return EVK_OK;
    const size_t m = h->max_events;
    EVK_CUDA(h, cudaMalloc(&c->sk, m * 8));
    EVK_CUDA(h, cudaMalloc(&c->rk, m * 8));
    EVK_CUDA(h, cudaMalloc(&c->sf, m * 4));
    EVK_CUDA(h, cudaMalloc(&c->rf, m * 4));
    EVK_CUDA(h, cudaMalloc(&c->sx, m * 4));
    EVK_CUDA(h, cudaMalloc(&c->rx, m * 4));
    EVK_CUDA(h, cudaMalloc(&c->sr, m * 16));
    EVK_CUDA(h, cudaMalloc(&c->rr, m * 16));
    if (!h->d_reps) EVK_CUDA(h, cudaMalloc(&h->d_reps, h->out_cap * 16));
    c->stage_cap = m;
    return EVK_OK;
}

// allreduce (sum) of the first `n` stats words, result mirrored to the host
int stats_allreduce(evk_handle* h, CommState* c, int n) {
    EVK_NCCL(h, g_nccl.AllReduce(c->d_stats, c->d_stats, n, ncclUint64, ncclSum, c->comm,
                                 h->stream));
    EVK_CUDA(h, cudaMemcpyAsync(c->h_stats, c->d_stats, n * 8, cudaMemcpyDeviceToHost, h->stream));
    EVK_CUDA(h, cudaStreamSynchronize(h->stream));
    return EVK_OK;
}

int sharded_time_range(evk_handle* h, const evk_ds_params* p, bool* ok) {
    CommState* c = h->comm;
    *ok = false;
    KeyParams kp;
    EVK_TRY(evk_make_key_params(h, p, &kp));
    const size_t n_own = h->n_events;
    const uint32_t halo = c->halo;
    if (n_own + halo > h->max_events)
        return evk_fail(h, EVK_ERR_CAPACITY,
                        "sharded downsample needs max_events >= n_events + %u (halo)", halo);
    unsigned long long local_flag = evk_slab_supported(h, kp) ? 0 : 1;
    // 1. boundary block: my head goes to the previous rank, the next rank's head arrives behind
    //    my own events (contiguous in the global index space)
    EVK_NCCL(h, g_nccl.GroupStart());
    if (c->rank > 0)
        EVK_NCCL(h, g_nccl.Send(h->d_events, (size_t)halo * 16, ncclUint8, c->rank - 1, c->comm,
                                h->stream));
    if (c->rank < c->world - 1)
        EVK_NCCL(h, g_nccl.Recv(h->d_events + n_own, (size_t)halo * 16, ncclUint8, c->rank + 1,
                                c->comm, h->stream));
    EVK_NCCL(h, g_nccl.GroupEnd());
    // 2. who keeps what
    k_halo_range<<<1, 32, 0, h->stream>>>(kp, h->d_events, (uint32_t)n_own, halo, c->rank,
                                          c->world, c->d_stats);
    EVK_CUDA(h, cudaGetLastError());
    EVK_CUDA(h, cudaMemcpyAsync(c->h_stats, c->d_stats, 8 * 8, cudaMemcpyDeviceToHost, h->stream));
    EVK_CUDA(h, cudaStreamSynchronize(h->stream));
    local_flag |= c->h_stats[ST_FLAG];
    const size_t skip = local_flag ? 0 : (size_t)c->h_stats[ST_SKIP];
    const size_t keep = local_flag ? 0 : (size_t)c->h_stats[ST_KEEP];
    // 3. the single-GPU path on my bins: events [skip, n_own + keep)
    evk_event* ev_saved = h->d_events;
    const uint64_t first_saved = h->shard_first;
    int st = EVK_OK;
    if (!local_flag) {
        h->d_events = ev_saved + skip;
        h->n_events = n_own - skip + keep;
        h->shard_first = first_saved + skip;
        evk_ds_params q = *p;
        q.algo = EVK_ALGO_SLAB;
        st = evk_downsample_local(h, &q);
        if (st == EVK_OK && h->times.ds_algo_used != EVK_ALGO_SLAB) local_flag = 1;  // unordered
        h->d_events = ev_saved;
        h->n_events = n_own;
        h->shard_first = first_saved;
    }
    if (st != EVK_OK) local_flag = 1;
    // 4. all ranks agree
    c->h_stats[ST_FLAG] = local_flag;
    c->h_stats[ST_U] = local_flag ? 0 : h->n_unique;
    c->h_stats[ST_R] = local_flag ? 0 : h->n_repeated;
    EVK_CUDA(h, cudaMemcpyAsync(c->d_stats, c->h_stats, 3 * 8, cudaMemcpyHostToDevice, h->stream));
    EVK_TRY(stats_allreduce(h, c, 3));
    if (st != EVK_OK && c->h_stats[ST_FLAG] == 0) return st;
    *ok = c->h_stats[ST_FLAG] == 0;
    return EVK_OK;
}

int sharded_mix64(evk_handle* h, const evk_ds_params* p) {
    CommState* c = h->comm;
    const int G = c->world;
    if (h->table_cap > (1ull << 32))
        return evk_fail(h, EVK_ERR_CAPACITY, "hash-owned exchange supports max_events < 2^31");
    EVK_TRY(stage_alloc(h, c));
    evk_ds_params q = *p;
    if (q.algo == EVK_ALGO_SLAB) q.algo = EVK_ALGO_AUTO;
    EVK_TRY(evk_downsample_local(h, &q));  // local voxels, global first indices
    const size_t U = h->n_unique;
    unsigned long long* d_hist = c->d_stats + ST_HIST;          // [G] mine
    unsigned long long* d_all = d_hist + kMaxWorld;             // [G][G] everyone's
    unsigned long long* d_off = d_all + kMaxWorld * kMaxWorld;  // [G] send offsets
    unsigned long long* d_cur = d_off + kMaxWorld;              // [G] cursors
    EVK_CUDA(h, cudaMemsetAsync(d_hist, 0, kMaxWorld * 8, h->stream));
    EVK_CUDA(h, cudaMemsetAsync(d_cur, 0, kMaxWorld * 8, h->stream));
    if (U) k_owner_hist<<<grid_for(U, h->sm_count), kBlock, 0, h->stream>>>(h->d_keys, U, G, d_hist);
    EVK_CUDA(h, cudaGetLastError());
    EVK_NCCL(h, g_nccl.AllGather(d_hist, d_all, G, ncclUint64, c->comm, h->stream));
    unsigned long long* hs = c->h_stats + ST_HIST;
    EVK_CUDA(h, cudaMemcpyAsync(hs, d_all, (size_t)G * G * 8, cudaMemcpyDeviceToHost, h->stream));
    EVK_CUDA(h, cudaStreamSynchronize(h->stream));
    // counts[src][dst] = hs[src * G + dst]
    std::vector<unsigned long long> soff(G + 1, 0), roff(G + 1, 0);
    for (int d = 0; d < G; d++) soff[d + 1] = soff[d] + hs[c->rank * G + d];
    for (int s = 0; s < G; s++) roff[s + 1] = roff[s] + hs[s * G + c->rank];
    const size_t M = (size_t)roff[G];
    if (M > c->stage_cap || M > h->max_events)
        return evk_fail(h, EVK_ERR_CAPACITY, "rank %d would own %zu voxels (capacity %zu)", c->rank,
                        M, h->max_events);
    EVK_CUDA(h, cudaMemcpyAsync(d_off, soff.data(), G * 8, cudaMemcpyHostToDevice, h->stream));
    const evk_event* ev0 = h->d_events - h->shard_first;
    if (U)
        k_owner_scatter<<<grid_for(U, h->sm_count), kBlock, 0, h->stream>>>(
            h->d_keys, h->d_first, h->d_xy, ev0, U, G, d_off, d_cur, c->sk, c->sf, c->sx, c->sr);
    EVK_CUDA(h, cudaGetLastError());
    // all-to-all of the voxel records (one grouped exchange; own bucket is a device copy)
    EVK_NCCL(h, g_nccl.GroupStart());
    for (int peer = 0; peer < G; peer++) {
        const size_t sn = (size_t)(soff[peer + 1] - soff[peer]), so = (size_t)soff[peer];
        const size_t rn = (size_t)(roff[peer + 1] - roff[peer]), ro = (size_t)roff[peer];
        if (peer == c->rank) continue;
        if (sn) {
            EVK_NCCL(h, g_nccl.Send(c->sk + so, sn * 8, ncclUint8, peer, c->comm, h->stream));
            EVK_NCCL(h, g_nccl.Send(c->sf + so, sn * 4, ncclUint8, peer, c->comm, h->stream));
            EVK_NCCL(h, g_nccl.Send(c->sx + so, sn * 4, ncclUint8, peer, c->comm, h->stream));
            EVK_NCCL(h, g_nccl.Send(c->sr + so, sn * 16, ncclUint8, peer, c->comm, h->stream));
        }
        if (rn) {
            EVK_NCCL(h, g_nccl.Recv(c->rk + ro, rn * 8, ncclUint8, peer, c->comm, h->stream));
            EVK_NCCL(h, g_nccl.Recv(c->rf + ro, rn * 4, ncclUint8, peer, c->comm, h->stream));
            EVK_NCCL(h, g_nccl.Recv(c->rx + ro, rn * 4, ncclUint8, peer, c->comm, h->stream));
            EVK_NCCL(h, g_nccl.Recv(c->rr + ro, rn * 16, ncclUint8, peer, c->comm, h->stream));
        }
    }
    EVK_NCCL(h, g_nccl.GroupEnd());
    {
        const size_t sn = (size_t)(soff[c->rank + 1] - soff[c->rank]), so = (size_t)soff[c->rank];
        const size_t ro = (size_t)roff[c->rank];
        if (sn) {
            EVK_CUDA(h, cudaMemcpyAsync(c->rk + ro, c->sk + so, sn * 8, cudaMemcpyDeviceToDevice, h->stream));
            EVK_CUDA(h, cudaMemcpyAsync(c->rf + ro, c->sf + so, sn * 4, cudaMemcpyDeviceToDevice, h->stream));
            EVK_CUDA(h, cudaMemcpyAsync(c->rx + ro, c->sx + so, sn * 4, cudaMemcpyDeviceToDevice, h->stream));
            EVK_CUDA(h, cudaMemcpyAsync(c->rr + ro, c->sr + so, sn * 16, cudaMemcpyDeviceToDevice, h->stream));
        }
    }
    // owner-side merge in the hash-owned table
    EVK_CUDA(h, cudaMemsetAsync(h->d_cnt, 0, sizeof(DsCounters), h->stream));
    if (M) {
        const uint64_t mask = (uint64_t)h->table_cap - 1;
        const int grid = grid_for(M, h->sm_count);
        k_merge_insert<<<grid, kBlock, 0, h->stream>>>(c->rk, c->rf, M, h->d_tkeys, h->d_tfirst, mask);
        k_merge_emit<<<grid, kBlock, 0, h->stream>>>(c->rk, c->rf, c->rx, c->rr, M, h->d_tkeys,
                                                     h->d_tfirst, mask, h->d_keys, h->d_first,
                                                     h->d_xy, h->d_reps, c->sx, h->d_cnt);
        k_merge_reset<<<grid, kBlock, 0, h->stream>>>(c->sx, M, h->d_tkeys, h->d_tfirst);
        EVK_CUDA(h, cudaGetLastError());
    }
    EVK_CUDA(h, cudaMemcpyAsync(h->h_cnt, h->d_cnt, sizeof(DsCounters), cudaMemcpyDeviceToHost, h->stream));
    EVK_CUDA(h, cudaStreamSynchronize(h->stream));
    h->n_unique = (size_t)h->h_cnt->n_unique;
    h->n_repeated = 0;  // per-key hit counts are not exchanged in this mode
    h->perm_valid = false;
    h->reps_valid = true;
    h->voxels_foreign = true;
    h->have_voxels = true;
    c->h_stats[ST_FLAG] = 0;
    c->h_stats[ST_U] = h->n_unique;
    c->h_stats[ST_R] = 0;
    EVK_CUDA(h, cudaMemcpyAsync(c->d_stats, c->h_stats, 3 * 8, cudaMemcpyHostToDevice, h->stream));
    EVK_TRY(stats_allreduce(h, c, 3));
    return EVK_OK;
}

// Hash-owned downsample over peer memory (see the k_mix_* kernels).  One host synchronisation after
// the local downsample (its algorithm -- slab, partition, table -- is chosen on the host) and one at
// the end.  want_reps: also deliver the representative events (evk_get_voxels needs them; the fused
// step does not).
int sharded_mix64_p2p(evk_handle* h, const evk_ds_params* p, bool want_reps) {
    CommState* c = h->comm;
    const int G = c->world;
    if (h->table_cap > (1ull << 32))
        return evk_fail(h, EVK_ERR_CAPACITY, "hash-owned exchange supports max_events < 2^31");
    // 1. local downsample into the local voxel buffers (the shard arrays are written by peers)
    evk_ds_params q = *p;
    if (q.algo == EVK_ALGO_SLAB) q.algo = EVK_ALGO_AUTO;
    std::swap(h->d_keys, c->lk);
    std::swap(h->d_first, c->lf);
    std::swap(h->d_xy, c->lx);
    const int st = evk_downsample_local(h, &q);
    std::swap(h->d_keys, c->lk);
    std::swap(h->d_first, c->lf);
    std::swap(h->d_xy, c->lx);
    // (an error on one rank only would leave the others waiting: it is reported after the exchange)
    const size_t U = st == EVK_OK ? h->n_unique : 0;
    const bool ordered = st == EVK_OK && h->times.ds_algo_used == EVK_ALGO_SLAB;
    // the slab kernels leave the first bin and the bin count behind (DsCounters::scratch[2], [0])
    const unsigned long long tb_first = ordered ? h->h_cnt->scratch[2] : 0;
    const unsigned long long tb_last = ordered ? tb_first + h->h_cnt->scratch[0] - 1 : 0;
    KeyParams kp;
    EVK_TRY(evk_make_key_params(h, p, &kp));
    MixDiv dv;
    dv.cells = kp.cells ? kp.cells : 1;
    dv.magic = dv.cells > 1 ? ~0ull / dv.cells + 1 : 0;
    dv.lo_key = dv.hi_key = 0;
    if (ordered && kp.vt > 0 && kp.keyfn == EVK_KEY_VOXEL) {
        dv.lo_key = (tb_first + 1) * dv.cells;  // first key behind my first bin
        dv.hi_key = tb_last * dv.cells;         // first key of my last bin
        if (dv.hi_key < dv.lo_key) dv.hi_key = dv.lo_key;
    }
    uint8_t* bkt = reinterpret_cast<uint8_t*>(c->sx);  // (idle send staging: one byte per record)
    const bool all_suspect = !ordered || kp.keyfn != EVK_KEY_VOXEL || kp.vt <= 0;
    unsigned long long* mx = c->d_mix;
    static const bool trace = getenv("EVK_MIX_TRACE") != nullptr;
    static cudaEvent_t tev[8] = {};
    auto mark = [&](int i) {
        if (!trace) return;
        if (!tev[i]) cudaEventCreate(&tev[i]);
        cudaEventRecord(tev[i], h->stream);
    };
    EVK_CUDA(h, cudaMemsetAsync(mx, 0, MX_WORDS * sizeof(unsigned long long), h->stream));
    EVK_CUDA(h, cudaMemsetAsync(h->d_cnt, 0, sizeof(DsCounters), h->stream));
    evk_prof_rec(h, 0);
    mark(0);
    k_p2p_tick<<<1, 1, 0, h->stream>>>(c->mail, c->d_peer_mail, 0);  // seq++ (no neighbour flag)
    k_mix_publish_info<<<1, 64, 0, h->stream>>>(c->mail, c->d_peer_mail, c->rank, G, tb_first,
                                                tb_last, all_suspect ? 1ull : 0ull);
    const int grid = grid_for(U ? U : 1, h->sm_count);
    k_mix_hist<<<grid, kBlock, 0, h->stream>>>(c->lk, U, G, dv, c->mail, mx, bkt);
    mark(1);
    k_mix_publish_counts<<<1, kBlock, 0, h->stream>>>(c->mail, c->d_peer_mail, c->rank, G, mx);
    k_mix_offsets<<<1, 64, 0, h->stream>>>(c->mail, c->rank, G, (unsigned long long)h->out_cap,
                                           (unsigned long long)c->stage_cap, mx, h->d_cnt);
    const evk_event* ev0 = h->d_events - h->shard_first;
    mark(2);
    static bool carve_set = false;
    if (!carve_set) {  // (latency-bound: ncu showed 3 resident CTAs at 72 registers)
        cudaFuncSetAttribute(k_mix_scatter, cudaFuncAttributePreferredSharedMemoryCarveout, 60);
        carve_set = true;
    }
    k_mix_scatter<<<h->sm_count * kScCtas, kBlock, 0, h->stream>>>(c->lk, c->lf, c->lx, ev0, bkt, U, G,
                                                             c->d_mix_peers, mx, want_reps ? 1 : 0,
                                                             getenv("EVK_MIX_TIMING_LOCAL_ONLY") ? c->rank : -1);
    mark(3);
    k_mix_done<<<1, 64, 0, h->stream>>>(c->mail, c->d_peer_mail, c->rank, G, mx);
    mark(4);
    // 2. the owner resolves its suspects: lowest global first index wins, winners are appended
    {
        const uint64_t mask = (uint64_t)h->table_cap - 1;
        const int mg = h->sm_count * 8;
        k_mix_merge_insert<<<mg, kBlock, 0, h->stream>>>(c->rk, c->rf, mx, h->d_tkeys, h->d_tfirst,
                                                         mask);
        k_mix_merge_emit<<<mg, kBlock, 0, h->stream>>>(c->rk, c->rf, c->rx, c->rr, mx, h->d_tkeys,
                                                       h->d_tfirst, mask, h->d_keys, h->d_first,
                                                       h->d_xy, h->d_reps, c->sf, h->d_cnt,
                                                       want_reps ? 1 : 0);
        k_mix_merge_reset<<<mg, kBlock, 0, h->stream>>>(c->sf, mx, h->d_tkeys, h->d_tfirst);
    }
    EVK_CUDA(h, cudaGetLastError());
    mark(5);
    if (trace) {
        cudaEventSynchronize(tev[5]);
        float t[5];
        for (int i = 0; i < 5; i++) cudaEventElapsedTime(&t[i], tev[i], tev[i + 1]);
        fprintf(stderr, "[evk mix64 rank %d] U=%zu hist %.3f counts+offsets %.3f scatter %.3f "
                "done-wait %.3f merge %.3f ms\n", c->rank, U, t[0], t[1], t[2], t[3], t[4]);
    }
    evk_prof_rec(h, 1);
    evk_prof_rec(h, 2);
    EVK_CUDA(h, cudaMemcpyAsync(c->h_mix, mx + MX_TAIL, 8 * sizeof(unsigned long long),
                                cudaMemcpyDeviceToHost, h->stream));
    EVK_CUDA(h, cudaMemcpyAsync(h->h_cnt, h->d_cnt, sizeof(DsCounters), cudaMemcpyDeviceToHost,
                                h->stream));
    return st;
}

// collection half of the exchange: counters to the host, state of the handle
int sharded_mix64_p2p_finish(evk_handle* h, int st_local, bool want_reps) {
    CommState* c = h->comm;
    EVK_CUDA(h, cudaStreamSynchronize(h->stream));
    if (c->h_mix[2])
        return evk_fail(h, EVK_ERR_CAPACITY, "hash-owned exchange: a rank would own more voxels than "
                        "its capacity, or a peer did not answer (rank %d owns %llu + %llu)", c->rank,
                        c->h_mix[0], c->h_mix[1]);
    EVK_TRY(st_local);
    h->n_unique = (size_t)h->h_cnt->n_unique;
    h->n_repeated = 0;  // per-key hit counts are not exchanged in this mode
    h->perm_valid = false;
    h->reps_valid = want_reps;
    h->voxels_foreign = true;
    h->have_voxels = true;
    h->pix_valid = false;
    h->times.ds_launches += 10;
    return EVK_OK;
}

// Map every rank's mailbox and event buffer into this process (CUDA IPC; the 64-byte handles travel
// through one NCCL allgather).  Any failure leaves c->p2p false: the NCCL path is used instead.
void p2p_setup(evk_handle* h, CommState* c) {
    struct Handles {
        cudaIpcMemHandle_t mail, events;
    };
    static_assert(sizeof(Handles) * kMaxWorld <= 8 * (ST_HIST + (size_t)kMaxWorld * (kMaxWorld + 3)),
                  "the stats scratch doubles as the handle exchange buffer");
    Handles mine;
    memset(&mine, 0, sizeof mine);
    Handles* d_all = reinterpret_cast<Handles*>(c->d_stats);
    std::vector<Handles> all(c->world);
    bool ok = cudaMalloc((void**)&c->mail, sizeof(Mailbox)) == cudaSuccess &&
              cudaMemset(c->mail, 0, sizeof(Mailbox)) == cudaSuccess &&
              cudaIpcGetMemHandle(&mine.mail, c->mail) == cudaSuccess &&
              cudaIpcGetMemHandle(&mine.events, h->d_events) == cudaSuccess;
    // both collectives below are reached by every rank whatever happened above: a rank whose
    // set-up failed publishes zero handles and votes "no" in the allreduce, and all fall back
    bool coll = cudaMemcpy(d_all + c->rank, &mine, sizeof mine, cudaMemcpyHostToDevice) == cudaSuccess;
    coll = g_nccl.AllGather(d_all + c->rank, d_all, sizeof(Handles), ncclUint8, c->comm,
                            h->stream) == ncclSuccess && coll;
    coll = cudaStreamSynchronize(h->stream) == cudaSuccess && coll;
    coll = cudaMemcpy(all.data(), d_all, sizeof(Handles) * c->world, cudaMemcpyDeviceToHost) ==
               cudaSuccess && coll;
    ok = ok && coll;
    for (int r = 0; ok && r < c->world; r++) {
        if (r == c->rank) {
            c->peer_mail[r] = c->mail;
            c->peer_events[r] = h->d_events;
            continue;
        }
        void *pm = nullptr, *pe = nullptr;
        ok = cudaIpcOpenMemHandle(&pm, all[r].mail, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess &&
             cudaIpcOpenMemHandle(&pe, all[r].events, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
        c->peer_mail[r] = static_cast<Mailbox*>(pm);
        c->peer_events[r] = static_cast<const evk_event*>(pe);
    }
    if (ok)
        ok = cudaMalloc((void**)&c->d_peer_mail, sizeof(Mailbox*) * kMaxWorld) == cudaSuccess &&
             cudaMemcpy(c->d_peer_mail, c->peer_mail, sizeof(Mailbox*) * kMaxWorld,
                        cudaMemcpyHostToDevice) == cudaSuccess;
    // hash-owned exchange: every rank's voxel shard and staging arrays, mapped the same way
    bool mix = ok && getenv("EVK_NO_MIX_P2P") == nullptr;
    struct MixHandles {
        cudaIpcMemHandle_t a[8];
    };
    MixHandles mh;
    memset(&mh, 0, sizeof mh);
    MixHandles* d_mh = nullptr;
    std::vector<MixHandles> allm(c->world);
    if (mix) {
        mix = stage_alloc(h, c) == EVK_OK &&
              cudaMalloc((void**)&c->lk, h->out_cap * 8) == cudaSuccess &&
              cudaMalloc((void**)&c->lf, h->out_cap * 4) == cudaSuccess &&
              cudaMalloc((void**)&c->lx, h->out_cap * 4) == cudaSuccess &&
              cudaMalloc((void**)&c->d_mix, MX_WORDS * 8) == cudaSuccess &&
              cudaMallocHost((void**)&c->h_mix, 8 * 8) == cudaSuccess &&
              cudaMalloc((void**)&c->d_mix_peers, sizeof(MixPeer) * kMaxWorld) == cudaSuccess;
        void* mineptr[8] = {h->d_keys, h->d_first, h->d_xy, h->d_reps, c->rk, c->rf, c->rx, c->rr};
        for (int i = 0; mix && i < 8; i++)
            mix = cudaIpcGetMemHandle(&mh.a[i], mineptr[i]) == cudaSuccess;
    }
    // (collectives below are reached by every rank whatever happened above)
    bool coll2 = cudaMalloc((void**)&d_mh, sizeof(MixHandles) * c->world) == cudaSuccess;
    coll2 = coll2 && cudaMemcpy(d_mh + c->rank, &mh, sizeof mh, cudaMemcpyHostToDevice) == cudaSuccess;
    coll2 = g_nccl.AllGather(d_mh ? d_mh + c->rank : nullptr, d_mh, sizeof(MixHandles), ncclUint8,
                             c->comm, h->stream) == ncclSuccess && coll2;
    coll2 = cudaStreamSynchronize(h->stream) == cudaSuccess && coll2;
    coll2 = coll2 && cudaMemcpy(allm.data(), d_mh, sizeof(MixHandles) * c->world,
                                cudaMemcpyDeviceToHost) == cudaSuccess;
    if (d_mh) cudaFree(d_mh);
    mix = mix && coll2;
    std::vector<MixPeer> mp(kMaxWorld);
    for (int r = 0; mix && r < c->world; r++) {
        void* ptr[8] = {h->d_keys, h->d_first, h->d_xy, h->d_reps, c->rk, c->rf, c->rx, c->rr};
        if (r != c->rank)
            for (int i = 0; mix && i < 8; i++) {
                mix = cudaIpcOpenMemHandle(&ptr[i], allm[r].a[i], cudaIpcMemLazyEnablePeerAccess) ==
                      cudaSuccess;
                c->ipc_opened[r][i] = mix ? ptr[i] : nullptr;
            }
        mp[r].keys = static_cast<uint64_t*>(ptr[0]);
        mp[r].first = static_cast<uint32_t*>(ptr[1]);
        mp[r].xy = static_cast<uint32_t*>(ptr[2]);
        mp[r].reps = static_cast<evk_event*>(ptr[3]);
        mp[r].sk = static_cast<uint64_t*>(ptr[4]);
        mp[r].sf = static_cast<uint32_t*>(ptr[5]);
        mp[r].sx = static_cast<uint32_t*>(ptr[6]);
        mp[r].sr = static_cast<evk_event*>(ptr[7]);
    }
    if (mix)
        mix = cudaMemcpy(c->d_mix_peers, mp.data(), sizeof(MixPeer) * kMaxWorld,
                         cudaMemcpyHostToDevice) == cudaSuccess;
    // all ranks agree: P2P only if it works everywhere
    unsigned long long flag = (ok ? 0 : 1) + (mix ? 0 : (1ull << 16));
    if (cudaMemcpy(c->d_stats, &flag, 8, cudaMemcpyHostToDevice) == cudaSuccess &&
        g_nccl.AllReduce(c->d_stats, c->d_stats, 1, ncclUint64, ncclSum, c->comm, h->stream) ==
            ncclSuccess &&
        cudaStreamSynchronize(h->stream) == cudaSuccess &&
        cudaMemcpy(&flag, c->d_stats, 8, cudaMemcpyDeviceToHost) == cudaSuccess) {
        // (sum over ranks: bit 0 of any rank -> odd contributions; test the two conditions apart)
        c->p2p = ok_all(flag, c->world, 1);
        c->mix_p2p = c->p2p && ok_all(flag, c->world, 2);
    }
    cudaGetLastError();
}

int allreduce_acc(evk_handle* h, int K, int D) {
    (void)D;
    CommState* c = h->comm;
    EVK_NCCL(h, g_nccl.AllReduce(h->d_acc, h->d_acc, (size_t)K * 5, ncclUint64, ncclSum, c->comm,
                                 h->stream));
    return EVK_OK;
}

}  // namespace

extern "C" {

int evk_comm_unique_id(uint8_t* id128) {
    if (!id128 || !load_nccl()) return EVK_ERR_COMM;
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return EVK_ERR_COMM;
    memcpy(id128, &id, 128);
    return EVK_OK;
}

int evk_comm_init(evk_handle* h, int rank, int world, const uint8_t* id128) {
    if (!h) return EVK_ERR_INVALID;
    if (!id128 || world < 1 || world > kMaxWorld || rank < 0 || rank >= world)
        return evk_fail(h, EVK_ERR_INVALID, "bad communicator shape rank=%d world=%d", rank, world);
    if (!load_nccl()) return evk_fail(h, EVK_ERR_COMM, "libnccl.so.2 not found: %s", dlerror());
    if (h->comm) evk_comm_destroy(h);
    DeviceGuard dev_guard(h->device);
    CommState* c = new CommState();
    c->rank = rank;
    c->world = world;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
        delete c;
        return evk_fail(h, EVK_ERR_COMM, "ncclCommInitRank: %s", g_nccl.GetErrorString(r));
    }
    const size_t words = ST_HIST + (size_t)kMaxWorld * (kMaxWorld + 3);
    if (cudaMalloc(&c->d_stats, words * 8) != cudaSuccess ||
        cudaMallocHost(&c->h_stats, words * 8) != cudaSuccess) {
        delete c;
        return evk_fail(h, EVK_ERR_NOMEM, "communicator scratch");
    }
    if ((size_t)c->halo * 4 > h->max_events) c->halo = (uint32_t)(h->max_events / 4);
    h->comm = c;
    // (EVK_P2P_SINGLE: the peer-memory paths with one rank, for profiling their kernels under ncu)
    if ((world > 1 || getenv("EVK_P2P_SINGLE")) && !getenv("EVK_NO_P2P"))
        p2p_setup(h, c);  // best effort: NCCL stays the fallback
    return EVK_OK;
}

int evk_comm_destroy(evk_handle* h) {
    if (!h || !h->comm) return EVK_OK;
    CommState* c = h->comm;
    DeviceGuard dev_guard(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (c->step_exec) cudaGraphExecDestroy(c->step_exec);
    for (int r = 0; r < c->world; r++) {
        if (r == c->rank) continue;
        if (c->peer_mail[r]) cudaIpcCloseMemHandle(c->peer_mail[r]);
        if (c->peer_events[r]) cudaIpcCloseMemHandle(const_cast<evk_event*>(c->peer_events[r]));
    }
    for (int r = 0; r < c->world; r++)
        for (int i = 0; i < 8; i++)
            if (c->ipc_opened[r][i]) cudaIpcCloseMemHandle(c->ipc_opened[r][i]);
    if (c->d_peer_mail) cudaFree(c->d_peer_mail);
    if (c->mail) cudaFree(c->mail);
    void* mixptrs[] = {c->lk, c->lf, c->lx, c->d_mix, c->d_mix_peers};
    for (void* q : mixptrs)
        if (q) cudaFree(q);
    if (c->h_mix) cudaFreeHost(c->h_mix);
    cudaGetLastError();
    if (c->comm && g_nccl.ok) g_nccl.CommDestroy(c->comm);
    void* ptrs[] = {c->d_stats, c->sk, c->rk, c->sf, c->rf, c->sx, c->rx, c->sr, c->rr};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    if (c->h_stats) cudaFreeHost(c->h_stats);
    delete c;
    h->comm = nullptr;
    return EVK_OK;
}

int evk_set_shard(evk_handle* h, uint64_t first_global_index) {
    if (!h) return EVK_ERR_INVALID;
    if (first_global_index + h->max_events >= 0xFF000000ull)
        return evk_fail(h, EVK_ERR_INVALID, "global indices must stay below 2^32");
    h->shard_first = first_global_index;
    return EVK_OK;
}

int evk_downsample_sharded(evk_handle* h, const evk_ds_params* p, int owner_mode,
                           size_t* n_unique_local, size_t* n_unique_global) {
    if (!h) return EVK_ERR_INVALID;
    if (!h->comm) return evk_fail(h, EVK_ERR_COMM, "evk_comm_init has not been called");
    if (!p) return evk_fail(h, EVK_ERR_INVALID, "ds params are NULL");
    CommState* c = h->comm;
    DeviceGuard dev_guard(h->device);
    bool done = false;
    if (owner_mode == EVK_OWNER_TIME_RANGE) {
        EVK_TRY(sharded_time_range(h, p, &done));
        if (done) c->last_mode = EVK_OWNER_TIME_RANGE;
    } else if (owner_mode != EVK_OWNER_MIX64) {
        return evk_fail(h, EVK_ERR_INVALID, "unknown owner mode %d", owner_mode);
    }
    if (!done) {
        if (c->mix_p2p) {
            const int st_local = sharded_mix64_p2p(h, p, true);
            EVK_TRY(sharded_mix64_p2p_finish(h, st_local, true));
            // the global voxel count for the caller (and for the fallback of the fused step)
            c->h_stats[ST_FLAG] = 0;
            c->h_stats[ST_U] = h->n_unique;
            c->h_stats[ST_R] = 0;
            EVK_CUDA(h, cudaMemcpyAsync(c->d_stats, c->h_stats, 3 * 8, cudaMemcpyHostToDevice,
                                        h->stream));
            EVK_TRY(stats_allreduce(h, c, 3));
        } else {
            EVK_TRY(sharded_mix64(h, p));
        }
        c->last_mode = EVK_OWNER_MIX64;
    }
    if (n_unique_local) *n_unique_local = h->n_unique;
    if (n_unique_global) *n_unique_global = (size_t)c->h_stats[ST_U];
    return EVK_OK;
}

int evk_kmeans_sharded(evk_handle* h, const evk_km_params* p, int* iters_done) {
    if (!h) return EVK_ERR_INVALID;
    if (!h->comm) return evk_fail(h, EVK_ERR_COMM, "evk_comm_init has not been called");
    if (p && p->on_events) return evk_fail(h, EVK_ERR_INVALID, "sharded k-means clusters voxels");
    return evk_kmeans_run(h, p, iters_done, allreduce_acc);
}

// everything the fused sharded step submits (captured into a graph by the caller)
static int enqueue_fused_sharded(evk_handle* h, const KeyParams& kp, const evk_ds_params* ds,
                                 const evk_km_params* km, int init_first_k, int* launches) {
    CommState* c = h->comm;
    const size_t n_own = h->n_events;
    const uint32_t halo = c->halo;
    const KmLaunch kl = evk_km_launch_params(h, km);
    unsigned long long* tail = h->d_acc + (size_t)km->K * 5;
    EVK_CUDA(h, cudaMemsetAsync(h->d_cnt, 0, sizeof(DsCounters), h->stream));
    evk_prof_rec(h, 0);
    // NCCL path: rank 0 walks the head of the global stream for the initial centroids first, then
    // ONE group carries the boundary blocks and the centroid broadcast (collectives cannot start
    // beside the downsample: its CTAs leave no shared memory on any SM).  Peer-memory path: the
    // walk and the centroid exchange run on the side stream, beside the downsample.
    if (init_first_k && c->rank == 0 && !c->p2p) {
        const size_t n_scan = n_own < (1u << 20) ? n_own : (1u << 20);
        if (n_scan)
            EVK_CUDA(h, evk_launch_init_first_k_walk(kp, kl, h->d_events, n_scan, h->d_cent,
                                                     &h->d_cnt->scratch[4], h->stream));
    }
    if (!init_first_k)  // warm start: keep a copy in case the pass is abandoned
        EVK_CUDA(h, cudaMemcpyAsync(h->d_cent + EVK_MAX_K * 2, h->d_cent,
                                    (size_t)km->K * 2 * sizeof(float), cudaMemcpyDeviceToDevice,
                                    h->stream));
    if (c->p2p) {  // peer memory over NVLink: flags + direct loads / stores (see k_p2p_*)
        // who keeps what is decided by the SENDER of a boundary block (a search over its own events)
        // and travels with the flag; the receiver pulls exactly its share
        k_p2p_tick_range<<<1, 1024, 0, h->stream>>>(kp, h->d_events, (uint32_t)n_own, halo, c->rank,
                                                  c->world, c->mail, c->d_peer_mail, c->d_stats);
        EVK_CUDA(h, cudaEventRecord(h->ev_fork, h->stream));  // the side stream reads the new seq
        if (c->rank < c->world - 1)
            // (one 16-B load per thread: the whole share is in flight over NVLink at once)
            k_p2p_pull_exact<<<(halo + 255) / 256, 256, 0, h->stream>>>(
                c->mail, c->peer_events[c->rank + 1], h->d_events + n_own, halo, c->d_stats);
        EVK_CUDA(h, cudaGetLastError());
    } else {
        EVK_NCCL(h, g_nccl.GroupStart());
        if (c->rank > 0)
            EVK_NCCL(h, g_nccl.Send(h->d_events, (size_t)halo * 16, ncclUint8, c->rank - 1,
                                    c->comm, h->stream));
        if (c->rank < c->world - 1)
            EVK_NCCL(h, g_nccl.Recv(h->d_events + n_own, (size_t)halo * 16, ncclUint8,
                                    c->rank + 1, c->comm, h->stream));
        if (init_first_k)
            EVK_NCCL(h, g_nccl.Broadcast(h->d_cent, h->d_cent, (size_t)km->K * 2, ncclFloat, 0,
                                         c->comm, h->stream));
        EVK_NCCL(h, g_nccl.GroupEnd());
        EVK_CUDA(h, cudaEventRecord(h->ev_fork, h->stream));
        // who keeps what (on the device), then the downsample on that range
        k_halo_range<<<1, 32, 0, h->stream>>>(kp, h->d_events, (uint32_t)n_own, halo, c->rank,
                                              c->world, c->d_stats);
        EVK_CUDA(h, cudaGetLastError());
    }
    bool ok = false;
    EVK_TRY(evk_downsample_slab(h, kp, ds->count_repeated, &ok, launches, false, c->d_stats));
    // side stream, beside the downsample: candidate lists, label map, quads
    EVK_CUDA(h, cudaStreamWaitEvent(h->side, h->ev_fork, 0));
    if (c->p2p && init_first_k) {
        if (c->rank == 0) {
            const size_t n_scan = n_own < (1u << 20) ? n_own : (1u << 20);
            if (n_scan)
                EVK_CUDA(h, evk_launch_init_first_k_walk(kp, kl, h->d_events, n_scan, h->d_cent,
                                                         &h->d_cnt->scratch[4], h->side));
            k_p2p_push_cent<<<1, 256, 0, h->side>>>(c->mail, c->d_peer_mail, c->world, h->d_cent,
                                                    km->K);
        } else {
            k_p2p_cent<<<1, 256, 0, h->side>>>(c->mail, h->d_cent, km->K);
        }
        EVK_CUDA(h, cudaGetLastError());
    }
    EVK_CUDA(h, evk_launch_km_image(kl, ds->width, ds->height, h->d_prune_lists, h->d_cent,
                                    nullptr, h->d_label_map, h->d_quads, h->d_acc, h->side));
    EVK_CUDA(h, cudaEventRecord(h->ev_join, h->side));
    EVK_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    evk_prof_rec(h, 2);
    evk_prof_rec(h, 3);
    EVK_CUDA(h, evk_launch_km_assign_tiles(kl, ds->width, ds->height, h->d_quads, h->d_label_map,
                                           h->d_xy, n_own + halo, &h->d_cnt->n_unique, true,
                                           h->d_acc, h->d_labels, h->sm_count, h->stream));
    if (c->p2p) {
        k_p2p_step_tail<<<1, 1024, 0, h->stream>>>(
            c->mail, c->d_peer_mail, c->rank, c->world, h->d_acc, km->K * 5 + 3, h->d_cnt, c->d_stats,
            (unsigned long long)km->K, init_first_k && c->rank == 0, kl, h->d_cent, h->d_counts,
            h->d_shift);
        EVK_CUDA(h, cudaGetLastError());
    } else {
        k_pack_step_stats<<<1, 1, 0, h->stream>>>(h->d_cnt, c->d_stats, (unsigned long long)km->K,
                                                  init_first_k && c->rank == 0, nullptr, tail);
        EVK_CUDA(h, cudaGetLastError());
        EVK_NCCL(h, g_nccl.AllReduce(h->d_acc, h->d_acc, (size_t)km->K * 5 + 3, ncclUint64,
                                     ncclSum, c->comm, h->stream));
        EVK_CUDA(h, cudaMemcpyAsync(h->d_cnt->step_tail, tail, 3 * 8, cudaMemcpyDeviceToDevice,
                                    h->stream));
        EVK_CUDA(h, evk_launch_km_finalise(kl, h->d_cent, h->d_acc, h->d_counts, h->d_shift,
                                           h->stream));
    }
    evk_prof_rec(h, 4);
    EVK_CUDA(h, cudaMemcpyAsync(h->h_cnt, h->d_cnt, sizeof(DsCounters), cudaMemcpyDeviceToHost,
                                h->stream));
    return EVK_OK;
}

// Fused sharded step (time-range ownership): halo exchange, downsample, centroid broadcast, one
// assign + accumulate pass and ONE allreduce (partial sums + voxel counters + give-up flag), all
// stream-ordered with a single host synchronisation at the end.  The event range a rank works on
// (after giving its first partial bin to the previous rank and keeping the next rank's) stays on
// the device.  Anything the fused pass does not take -- other ownership mode, D > 2, several
// iterations, an unordered stream on any rank -- runs the separate sharded calls.
// Submission half of the fused sharded step: a fusable shape goes out as one graph launch per rank
// and returns (h->step_pending = 1); steps may be queued behind each other -- every flag of the
// peer-memory exchange is a device-side sequence number and the mailboxes are double-buffered by
// step parity, and no rank can run more than one step ahead of its slowest peer (it needs that
// peer's partial sums to finish a step).  Other shapes run the three calls synchronously here.
int evk_downsample_kmeans_sharded_submit(evk_handle* h, const evk_ds_params* ds,
                                         const evk_km_params* km, int init_first_k,
                                         int owner_mode) {
    if (!h) return EVK_ERR_INVALID;
    if (!h->comm) return evk_fail(h, EVK_ERR_COMM, "evk_comm_init has not been called");
    EVK_TRY(evk_km_validate(h, km));
    KeyParams kp;
    EVK_TRY(evk_make_key_params(h, ds, &kp));
    CommState* c = h->comm;
    DeviceGuard dev_guard(h->device);
    if (!init_first_k && h->step_pending != 1 &&
        (!h->have_centroids || h->K != km->K || h->D != km->D))
        return evk_fail(h, EVK_ERR_STATE, "centroids for K=%d, D=%d have not been set", km->K, km->D);
    const size_t n_own = h->n_events;
    const uint32_t halo = c->halo;
    // rank-independent conditions only: every rank must take the same branch
    bool fusable = owner_mode == EVK_OWNER_TIME_RANGE && km->D == 2 && !km->on_events &&
                   km->K <= 254 && km->iters == 1 && km->tol < 0.f &&
                   (ds->algo == EVK_ALGO_AUTO || ds->algo == EVK_ALGO_SLAB) &&
                   ds->keyfn == EVK_KEY_VOXEL && kp.vt > 0 && (uint64_t)ds->width * ds->height <= (64ull << 20);
    if (fusable) {
        evk_handle probe_shape = evk_handle();  // slab shape check without the rank's event count
        probe_shape.n_events = 1;
        probe_shape.sm_count = h->sm_count;
        fusable = evk_slab_supported(&probe_shape, kp);
    }
    h->step_ds = *ds;
    h->step_km = *km;
    h->step_init = init_first_k;
    h->step_iters = 0;
    c->step_owner = owner_mode;
    c->step_fusable = fusable;
    // hash ownership over peer memory: local downsample (host-synchronous: its algorithm is chosen
    // on the host), then exchange + merge + the k-means pass + the allreduce as one stream of
    // kernels with a single synchronisation in _wait.  Representatives are not exchanged here.
    const bool mix_fused = owner_mode == EVK_OWNER_MIX64 && c->mix_p2p && km->D == 2 &&
                           !km->on_events && km->K <= 254 && km->iters == 1 && km->tol < 0.f &&
                           ds->keyfn == EVK_KEY_VOXEL &&
                           (uint64_t)ds->width * ds->height <= (64ull << 20);
    if (mix_fused) {
        if (h->step_pending)
            EVK_TRY(evk_downsample_kmeans_sharded_wait(h, nullptr, nullptr, nullptr));
        if (!evk_ensure_images(h, ds->width, ds->height))
            return evk_fail(h, EVK_ERR_NOMEM, "pixel images");
        c->step_fusable = false;
        const int st_local = sharded_mix64_p2p(h, ds, false);
        EVK_TRY(st_local);  // (a local failure after the exchange was queued: peers see U = 0)
        h->ds = *ds;
        h->kp = kp;
        h->have_ds = true;
        const KmLaunch kl = evk_km_launch_params(h, km);
        unsigned long long* tail = h->d_acc + (size_t)km->K * 5;
        EVK_CUDA(h, cudaEventRecord(h->ev_fork, h->stream));
        EVK_CUDA(h, cudaStreamWaitEvent(h->side, h->ev_fork, 0));
        if (init_first_k) {
            if (c->rank == 0) {
                const size_t n_scan = n_own < (1u << 20) ? n_own : (1u << 20);
                if (n_scan)
                    EVK_CUDA(h, evk_launch_init_first_k_walk(kp, kl, h->d_events, n_scan, h->d_cent,
                                                             &h->d_cnt->scratch[4], h->side));
                k_p2p_push_cent<<<1, 256, 0, h->side>>>(c->mail, c->d_peer_mail, c->world, h->d_cent,
                                                        km->K);
            } else {
                k_p2p_cent<<<1, 256, 0, h->side>>>(c->mail, h->d_cent, km->K);
            }
            EVK_CUDA(h, cudaGetLastError());
        }
        EVK_CUDA(h, evk_launch_km_image(kl, ds->width, ds->height, h->d_prune_lists, h->d_cent,
                                        nullptr, h->d_label_map, h->d_quads, h->d_acc, h->side));
        EVK_CUDA(h, cudaEventRecord(h->ev_join, h->side));
        EVK_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_join, 0));
        evk_prof_rec(h, 3);
        EVK_CUDA(h, evk_launch_km_assign_tiles(kl, ds->width, ds->height, h->d_quads, h->d_label_map,
                                               h->d_xy, h->out_cap, &h->d_cnt->n_unique, true,
                                               h->d_acc, h->d_labels, h->sm_count, h->stream));
        k_pack_mix_stats<<<1, 1, 0, h->stream>>>(h->d_cnt, c->d_mix + MX_TAIL,
                                                 (unsigned long long)km->K,
                                                 init_first_k && c->rank == 0, c->mail, tail);
        k_p2p_allreduce<<<1, 1024, 0, h->stream>>>(c->mail, c->d_peer_mail, c->rank, c->world,
                                                   h->d_acc, km->K * 5 + 3);
        EVK_CUDA(h, cudaGetLastError());
        EVK_CUDA(h, cudaMemcpyAsync(c->h_stats, tail, 3 * 8, cudaMemcpyDeviceToHost, h->stream));
        EVK_CUDA(h, evk_launch_km_finalise(kl, h->d_cent, h->d_acc, h->d_counts, h->d_shift,
                                           h->stream));
        evk_prof_rec(h, 4);
        EVK_CUDA(h, cudaMemcpyAsync(h->h_cnt, h->d_cnt, sizeof(DsCounters), cudaMemcpyDeviceToHost,
                                    h->stream));
        h->step_pending = 3;  // hash-owned step in flight
        h->step_sharded = true;
        return EVK_OK;
    }
    if (!fusable) {
        if (h->step_pending == 1 || h->step_pending == 3)
            EVK_TRY(evk_downsample_kmeans_sharded_wait(h, nullptr, nullptr, nullptr));
        h->step_pending = 0;
        EVK_TRY(evk_downsample_sharded(h, ds, owner_mode, nullptr, nullptr));
        if (init_first_k) EVK_TRY(evk_init_centroids_first_k_sharded(h, km));
        EVK_TRY(evk_kmeans_sharded(h, km, &h->step_iters));
        h->step_pending = 2;
        return EVK_OK;
    }
    if (n_own + halo > h->max_events)
        return evk_fail(h, EVK_ERR_CAPACITY,
                        "sharded downsample needs max_events >= n_events + %u (halo)", halo);
    if (!evk_ensure_images(h, ds->width, ds->height))
        return evk_fail(h, EVK_ERR_NOMEM, "pixel images");
    evk_invalidate_results(h);
    h->ds = *ds;
    h->kp = kp;
    h->have_ds = true;
    // one CUDA graph (NCCL operations included), re-captured only when the call's shape changes
    FusedKey key;
    memset(&key, 0, sizeof key);
    key.n = n_own;
    key.ds = *ds;
    key.km = *km;
    key.init = init_first_k ? 1 : 0;
    key.profiling = h->profiling ? 1 : 0;
    key.shard_first = h->shard_first;
    key.image_gen = h->image_gen;
    int launches = 0;
    if (c->step_exec && memcmp(&key, &c->step_key, sizeof key) == 0) {
        launches = c->step_launches;
    } else {
        if (h->step_pending == 1)  // the graph in flight is about to be replaced
            EVK_TRY(evk_downsample_kmeans_sharded_wait(h, nullptr, nullptr, nullptr));
        if (c->step_exec) cudaGraphExecDestroy(c->step_exec);
        c->step_exec = nullptr;
        cudaGraph_t graph = nullptr;
        EVK_CUDA(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        const int st_enq = enqueue_fused_sharded(h, kp, ds, km, init_first_k, &launches);
        cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
        if (st_enq != EVK_OK) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            return st_enq;
        }
        EVK_CUDA(h, ce);
        ce = cudaGraphInstantiate(&c->step_exec, graph, 0);
        cudaGraphDestroy(graph);
        EVK_CUDA(h, ce);
        c->step_key = key;
        c->step_launches = launches;
    }
    EVK_CUDA(h, cudaGraphLaunch(c->step_exec, h->stream));
    h->step_pending = 1;
    h->step_sharded = true;
    return EVK_OK;
}

// Collection half: the step's one host synchronisation; a step the fast path gave up on (unordered
// stream somewhere, peer-memory time-out) is rerun on the general path -- every rank sees the same
// flags, so every rank takes the same branch.
int evk_downsample_kmeans_sharded_wait(evk_handle* h, size_t* n_unique_local,
                                       size_t* n_unique_global, int* iters_done) {
    if (!h) return EVK_ERR_INVALID;
    if (!h->comm) return evk_fail(h, EVK_ERR_COMM, "evk_comm_init has not been called");
    if (!h->step_pending)
        return evk_fail(h, EVK_ERR_STATE, "evk_downsample_kmeans_sharded_wait: nothing submitted");
    CommState* c = h->comm;
    const int pending = h->step_pending;
    h->step_pending = 0;
    if (pending == 3) {  // hash-owned step: exchange + k-means pass were queued by submit
        DeviceGuard dev_guard(h->device);
        const evk_km_params* km = &h->step_km;
        EVK_TRY(sharded_mix64_p2p_finish(h, EVK_OK, false));
        if (c->h_stats[0])
            return evk_fail(h, (c->h_stats[0] >> 32) ? EVK_ERR_COMM : EVK_ERR_CAPACITY,
                            "hash-owned step abandoned (a peer timed out, a rank ran out of room, "
                            "or rank 0 holds fewer than K distinct voxels in its first 2^20 events)");
        c->h_stats[ST_U] = c->h_stats[1];
        h->K = km->K;
        h->D = km->D;
        h->have_centroids = true;
        h->n_labels = h->n_unique;
        h->labels_on_events = false;
        h->km_last = *km;
        c->last_mode = EVK_OWNER_MIX64;
        h->times.km_launches = 6;
        h->times.km_iters = 1;
        if (h->profiling) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, h->ev[3], h->ev[4]) == cudaSuccess)
                h->times.km_total_ms = h->times.km_assign_ms = ms;
            cudaGetLastError();
        }
        h->step_iters = 1;
    }
    if (pending == 1) {
        DeviceGuard dev_guard(h->device);
        const evk_ds_params* ds = &h->step_ds;
        const evk_km_params* km = &h->step_km;
        const int init_first_k = h->step_init;
        const int launches = c->step_launches;
        EVK_CUDA(h, cudaStreamSynchronize(h->stream));
        for (int i = 0; i < 3; i++) c->h_stats[i] = h->h_cnt->step_tail[i];  // (with the counters)
        if (c->h_stats[0] >> 32) {  // a peer-memory wait timed out somewhere: NCCL from now on
            c->p2p = false;
            cudaGraphExecDestroy(c->step_exec);
            c->step_exec = nullptr;
            cudaGetLastError();
        }
        if (c->h_stats[0] == 0) {
            h->n_unique = (size_t)h->h_cnt->n_unique;
            h->n_repeated = ds->count_repeated ? (size_t)h->h_cnt->n_repeated : 0;
            c->h_stats[ST_U] = c->h_stats[1];
            h->have_voxels = true;
            h->K = km->K;
            h->D = km->D;
            h->have_centroids = true;
            h->n_labels = h->n_unique;
            h->labels_on_events = false;
            h->km_last = *km;
            c->last_mode = EVK_OWNER_TIME_RANGE;
            h->times.ds_algo_used = EVK_ALGO_SLAB;
            h->times.ds_launches = launches + 2;
            h->times.km_launches = c->p2p ? 5 : 6;  // candidates, image, quads, assign, tail (NCCL: + pack, finalise)
            h->times.km_iters = 1;
            if (h->profiling) {
                float ms = 0.f;
                if (cudaEventElapsedTime(&ms, h->ev[0], h->ev[2]) == cudaSuccess) h->times.ds_total_ms = ms;
                if (cudaEventElapsedTime(&ms, h->ev[5], h->ev[6]) == cudaSuccess) h->times.ds_main_ms = ms;
                if (cudaEventElapsedTime(&ms, h->ev[3], h->ev[4]) == cudaSuccess)
                    h->times.km_total_ms = h->times.km_assign_ms = ms;
                cudaGetLastError();
            }
            h->step_iters = 1;
        } else {
            EVK_CUDA(h, cudaMemsetAsync(h->d_sticky, 0, sizeof(unsigned long long), h->stream));
            if (!init_first_k) {  // finalise has overwritten the caller's centroids
                if (!h->have_centroids || h->K != km->K || h->D != km->D)
                    return evk_fail(h, EVK_ERR_STATE,
                                    "centroids for K=%d, D=%d have not been set", km->K, km->D);
                EVK_CUDA(h, cudaMemcpyAsync(h->d_cent, h->d_cent + EVK_MAX_K * 2,
                                            (size_t)km->K * 2 * sizeof(float),
                                            cudaMemcpyDeviceToDevice, h->stream));
            }
            EVK_TRY(evk_downsample_sharded(h, ds, EVK_OWNER_MIX64, nullptr, nullptr));
            if (init_first_k) EVK_TRY(evk_init_centroids_first_k_sharded(h, km));
            EVK_TRY(evk_kmeans_sharded(h, km, &h->step_iters));
        }
    }
    if (n_unique_local) *n_unique_local = h->n_unique;
    if (n_unique_global) *n_unique_global = (size_t)c->h_stats[ST_U];
    if (iters_done) *iters_done = h->step_iters;
    return EVK_OK;
}

int evk_downsample_kmeans_sharded(evk_handle* h, const evk_ds_params* ds, const evk_km_params* km,
                                  int init_first_k, int owner_mode, size_t* n_unique_local,
                                  size_t* n_unique_global, int* iters_done) {
    if (!h) return EVK_ERR_INVALID;
    if (h->step_pending && h->comm)
        EVK_TRY(evk_downsample_kmeans_sharded_wait(h, nullptr, nullptr, nullptr));
    EVK_TRY(evk_downsample_kmeans_sharded_submit(h, ds, km, init_first_k, owner_mode));
    return evk_downsample_kmeans_sharded_wait(h, n_unique_local, n_unique_global, iters_done);
}

int evk_init_centroids_first_k_sharded(evk_handle* h, const evk_km_params* p) {
    if (!h) return EVK_ERR_INVALID;
    if (!h->comm) return evk_fail(h, EVK_ERR_COMM, "evk_comm_init has not been called");
    if (!p || p->K < 1 || p->K > EVK_MAX_K || p->D < 2 || p->D > EVK_MAX_D)
        return evk_fail(h, EVK_ERR_INVALID, "bad k-means params");
    if (!h->have_ds) return evk_fail(h, EVK_ERR_STATE, "evk_downsample_sharded has not run");
    CommState* c = h->comm;
    DeviceGuard dev_guard(h->device);
    // the K globally lowest first indices are the first K distinct keys of rank 0's shard
    unsigned long long* d_found = c->d_stats + 5;
    EVK_CUDA(h, cudaMemsetAsync(d_found, 0, 8, h->stream));
    if (c->rank == 0) {
        KmLaunch kl;
        kl.K = p->K;
        kl.D = p->D;
        kl.best2 = 0.f;
        kl.t_scale = p->t_scale;
        kl.p_scale = p->p_scale;
        kl.t0 = h->ds.t0_us;
        kl.write_labels = 0;
        const size_t n_scan = h->n_events < (1u << 20) ? h->n_events : (1u << 20);
        if (n_scan)
            EVK_CUDA(h, evk_launch_init_first_k_walk(h->kp, kl, h->d_events, n_scan, h->d_cent,
                                                     d_found, h->stream));
    }
    EVK_NCCL(h, g_nccl.Broadcast(h->d_cent, h->d_cent, (size_t)p->K * p->D, ncclFloat, 0, c->comm,
                                 h->stream));
    EVK_NCCL(h, g_nccl.Broadcast(d_found, d_found, 1, ncclUint64, 0, c->comm, h->stream));
    EVK_CUDA(h, cudaMemcpyAsync(c->h_stats + 5, d_found, 8, cudaMemcpyDeviceToHost, h->stream));
    EVK_CUDA(h, cudaStreamSynchronize(h->stream));
    if (c->h_stats[5] != (unsigned long long)p->K)
        return evk_fail(h, EVK_ERR_CAPACITY, "rank 0 holds only %llu distinct voxels in its first "
                        "2^20 events (K=%d)", c->h_stats[5], p->K);
    h->K = p->K;
    h->D = p->D;
    h->have_centroids = true;
    return EVK_OK;
}

}  // extern "C"
