// evk_synth.cu — event ingest kernels: on-device synthetic stream (include/evk_synth.h) and the
// packers that turn column (SoA) or reference-style interleaved int32 input into the 16-byte packed
// record.  Replaces the host packing loop of ACCEL/store.cpp:587-599 (EventCD -> int data[16384]).
#include "../../include/evk_synth.h"
#include "evk_internal.cuh"

namespace {
constexpr int kBlock = 256;

__global__ void __launch_bounds__(kBlock) k_synth(evk_synth_params sp, evk_event* out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < sp.n_events; j += stride) {
        evk_event e = evk_synth_event(&sp, sp.first_index + j);
        reinterpret_cast<uint4*>(out)[j] = *reinterpret_cast<const uint4*>(&e);
    }
}

__global__ void __launch_bounds__(kBlock)
    k_soa_pack(const uint16_t* __restrict__ x, const uint16_t* __restrict__ y,
               const int64_t* __restrict__ t, const uint8_t* __restrict__ p, size_t n,
               evk_event* out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t tt = t ? (uint64_t)t[i] : 0ull;
        uint4 w;
        w.x = (uint32_t)x[i] | ((uint32_t)y[i] << 16);
        w.y = p ? (uint32_t)p[i] : 0u;
        w.z = (uint32_t)tt;
        w.w = (uint32_t)(tt >> 32);
        reinterpret_cast<uint4*>(out)[i] = w;
    }
}

// interleaved int32 pairs (the reference kernel's input_coords).  Values outside [0,65535] cannot
// be a sensor coordinate: they are stored as x = y = 65535, t = INT64_MIN, which every gate rejects
// (REF gate: x <= width <= 65536 would admit 65535 only for width >= 65535 — such frames do not
// exist; VOXEL gate: t >= t0 fails).
__global__ void __launch_bounds__(kBlock)
    k_coords_pack(const int32_t* __restrict__ xy, size_t n, evk_event* out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int2 c = reinterpret_cast<const int2*>(xy)[i];
        uint4 w;
        if (c.x < 0 || c.x > 65535 || c.y < 0 || c.y > 65535) {
            w.x = 0xFFFFFFFFu;
            w.y = 0;
            w.z = 0;
            w.w = 0x80000000u;
        } else {
            w.x = (uint32_t)c.x | ((uint32_t)c.y << 16);
            w.y = 0;
            w.z = 0;
            w.w = 0;
        }
        reinterpret_cast<uint4*>(out)[i] = w;
    }
}

int grid_for(size_t n) {
    size_t need = (n + kBlock - 1) / kBlock;
    return (int)(need < 148 * 16 ? (need ? need : 1) : 148 * 16);
}
}  // namespace

cudaError_t evk_launch_synth(const evk_synth_params& sp, evk_event* out, cudaStream_t s) {
    k_synth<<<grid_for(sp.n_events), kBlock, 0, s>>>(sp, out);
    return cudaGetLastError();
}
cudaError_t evk_launch_soa_pack(const uint16_t* x, const uint16_t* y, const int64_t* t,
                                const uint8_t* p, size_t n, evk_event* out, cudaStream_t s) {
    k_soa_pack<<<grid_for(n), kBlock, 0, s>>>(x, y, t, p, n, out);
    return cudaGetLastError();
}
cudaError_t evk_launch_coords_pack(const int32_t* xy, size_t n, evk_event* out, cudaStream_t s) {
    k_coords_pack<<<grid_for(n), kBlock, 0, s>>>(xy, n, out);
    return cudaGetLastError();
}
