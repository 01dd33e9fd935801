// microbench.cu — B200 measurements that drive the kernel design (results in profiles/*.md):
// shared-memory atomic throughput (CAS / MIN / OR / ADD, 32 and 64 bit), global RED/ATOM on an
// L2-resident table, match.any, and the read-only streaming bandwidth of 16-byte records.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu && ./microbench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x)                                                                       \
    do {                                                                            \
        cudaError_t e = (x);                                                        \
        if (e != cudaSuccess) {                                                     \
            printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__);      \
            return 1;                                                               \
        }                                                                           \
    } while (0)

__device__ __forceinline__ uint32_t xs(uint32_t& s) {
    s ^= s << 13;
    s ^= s >> 17;
    s ^= s << 5;
    return s;
}

constexpr int kIters = 2048;
constexpr int kSlots = 4096;

// MODE 0 add32, 1 or32, 2 min32, 3 cas32, 4 add64, 5 min64(cas), 6 plain LDS+STS, 7 match.any
template <int MODE>
__global__ void __launch_bounds__(1024) k_smem(uint32_t* out, int spread) {
    __shared__ unsigned long long tab64[kSlots];
    uint32_t* tab = reinterpret_cast<uint32_t*>(tab64);
    for (int i = threadIdx.x; i < kSlots; i += blockDim.x) tab64[i] = 0xFFFFFFFFFFFFFFFFull;
    __syncthreads();
    uint32_t s = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 1u;
    uint32_t acc = 0;
    for (int it = 0; it < kIters; it++) {
        uint32_t r = xs(s);
        uint32_t a = spread ? (r & (kSlots - 1)) : ((r & 31) | ((threadIdx.x >> 5) << 5));
        if (MODE == 0) atomicAdd(&tab[a], r);
        if (MODE == 1) atomicOr(&tab[a], 1u << (r >> 27));
        if (MODE == 2) atomicMin(&tab[a], r);
        if (MODE == 3) acc += atomicCAS(&tab[a], 0xFFFFFFFFu, r);
        if (MODE == 4) atomicAdd(&tab64[a], (unsigned long long)r);
        if (MODE == 5) atomicMin(&tab64[a], (unsigned long long)r);
        if (MODE == 6) {
            acc += tab[a];
            tab[a ^ 1] = r;
        }
        if (MODE == 7) acc += __match_any_sync(0xffffffffu, r & 255);
    }
    __syncthreads();
    if (acc == 0x12345678u) out[0] = acc + tab[threadIdx.x];
    if (threadIdx.x == 0) out[blockIdx.x + 1] = tab[5];
}

// global atomics on an L2-resident table: MODE 0 RED.min (no return), 1 ATOM.min (return used),
// 2 RED.or, 3 ATOM.cas
template <int MODE>
__global__ void __launch_bounds__(256) k_gatom(uint32_t* tab, uint32_t mask, uint32_t* out) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 7u;
    uint32_t acc = 0;
    for (int it = 0; it < 256; it++) {
        uint32_t r = xs(s);
        uint32_t a = r & mask;
        if (MODE == 0) atomicMin(&tab[a], r);
        if (MODE == 1) acc += atomicMin(&tab[a], r);
        if (MODE == 2) atomicOr(&tab[a], r);
        if (MODE == 3) acc += atomicCAS(&tab[a], 0xFFFFFFFFu, r);
    }
    if (acc == 0x12345678u) out[0] = acc;
}

__global__ void __launch_bounds__(256) k_read16(const uint4* __restrict__ p, size_t n, uint32_t* out) {
    uint32_t acc = 0;
    size_t stride = (size_t)gridDim.x * blockDim.x * 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x * 4 + threadIdx.x; i < n; i += stride) {
        uint4 v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            size_t k = i + (size_t)j * blockDim.x;
            if (k < n)
                asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(v[j].x), "=r"(v[j].y), "=r"(v[j].z), "=r"(v[j].w)
                             : "l"(p + k));
            else
                v[j] = make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) acc += v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

template <typename F>
float time_ms(F f, int reps = 5) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    f();
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(a);
        f();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, sms, clk_khz);
    uint32_t* out;
    CK(cudaMalloc(&out, 4096 * 4));
    const char* names[] = {"add32", "or32", "min32", "cas32", "add64", "min64", "lds_sts", "match_any"};
    for (int spread = 1; spread >= 0; spread--) {
        for (int m = 0; m < 8; m++) {
            float ms = 0;
            auto run = [&]() {
                switch (m) {
                    case 0: k_smem<0><<<sms, 1024>>>(out, spread); break;
                    case 1: k_smem<1><<<sms, 1024>>>(out, spread); break;
                    case 2: k_smem<2><<<sms, 1024>>>(out, spread); break;
                    case 3: k_smem<3><<<sms, 1024>>>(out, spread); break;
                    case 4: k_smem<4><<<sms, 1024>>>(out, spread); break;
                    case 5: k_smem<5><<<sms, 1024>>>(out, spread); break;
                    case 6: k_smem<6><<<sms, 1024>>>(out, spread); break;
                    case 7: k_smem<7><<<sms, 1024>>>(out, spread); break;
                }
            };
            ms = time_ms(run);
            double ops = (double)sms * 1024 * kIters;
            printf("{\"bench\": \"smem_%s\", \"spread\": %d, \"ms\": %.4f, \"Gops_chip\": %.1f, "
                   "\"ns_per_warp_instr_per_sm\": %.2f}\n",
                   names[m], spread, ms, ops / ms * 1e-6, ms * 1e6 / (1024.0 / 32 * kIters));
        }
    }
    // global atomics, table sizes 256 KB .. 64 MB (L2-resident) and 1 GB (DRAM)
    size_t sizes[] = {1u << 16, 1u << 19, 1u << 22, 1u << 24, 1u << 28};
    const char* gnames[] = {"red_min", "atom_min", "red_or", "atom_cas"};
    for (size_t words : sizes) {
        uint32_t* tab;
        CK(cudaMalloc(&tab, words * 4));
        CK(cudaMemset(tab, 0xFF, words * 4));
        for (int m = 0; m < 4; m++) {
            int grid = sms * 16;
            auto run = [&]() {
                switch (m) {
                    case 0: k_gatom<0><<<grid, 256>>>(tab, (uint32_t)words - 1, out); break;
                    case 1: k_gatom<1><<<grid, 256>>>(tab, (uint32_t)words - 1, out); break;
                    case 2: k_gatom<2><<<grid, 256>>>(tab, (uint32_t)words - 1, out); break;
                    case 3: k_gatom<3><<<grid, 256>>>(tab, (uint32_t)words - 1, out); break;
                }
            };
            float ms = time_ms(run);
            double ops = (double)grid * 256 * 256;
            printf("{\"bench\": \"global_%s\", \"table_MB\": %.2f, \"ms\": %.4f, \"Gops\": %.1f}\n",
                   gnames[m], words * 4 / 1048576.0, ms, ops / ms * 1e-6);
        }
        cudaFree(tab);
    }
    {
        size_t n = 100000000;
        uint4* buf;
        CK(cudaMalloc(&buf, n * 16));
        CK(cudaMemset(buf, 1, n * 16));
        for (int waves : {4, 8, 16, 32}) {
            int grid = sms * waves;
            float ms = time_ms([&]() { k_read16<<<grid, 256>>>(buf, n, out); });
            printf("{\"bench\": \"read16_stream\", \"grid\": %d, \"ms\": %.4f, \"GBps\": %.1f}\n", grid,
                   ms, n * 16.0 / ms * 1e-6);
        }
        cudaFree(buf);
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
