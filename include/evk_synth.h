/* evk_synth.h — definition of the synthetic event stream, shared by the CUDA
 * generator kernel, the C oracle and the tests.  It defines INPUT DATA only, not the algorithm
 * under test.  Integer arithmetic throughout (no libm), so gcc and nvcc produce identical
 * events from (seed, index): every GPU shard and the CPU baseline see the same stream.
 *
 *   t_i      = floor(i * 1e6 / rate)                    (non-decreasing microseconds)
 *   u0,u1,u2 = splitmix64 chain seeded with seed ^ i
 *   noise_q16/65536 of the events: uniform background over the frame
 *   the rest: one of n_blobs moving Gaussian blobs (Irwin-Hall(8) normal, sigma_q8/256 px),
 *             blob centre c_m(t) = c_m0 + v_m * t, reflected at the borders
 *   p        = one random bit;  coordinates clamped to [0,W) x [0,H)
 */
#ifndef EVK_SYNTH_H_
#define EVK_SYNTH_H_

#include <stdint.h>
#include "evk.h"

#if defined(__CUDACC__)
#define EVK_HD __host__ __device__ __forceinline__
#else
#define EVK_HD static inline
#endif

EVK_HD uint64_t evk_sm64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* sum of the 8 bytes of u minus 1020: zero-mean, variance 8*(256^2-1)/12 = 43690 (std 209.02) */
EVK_HD int32_t evk_ih8(uint64_t u) {
    uint64_t s = (u & 0x00FF00FF00FF00FFull) + ((u >> 8) & 0x00FF00FF00FF00FFull);
    s = (s & 0x0000FFFF0000FFFFull) + ((s >> 16) & 0x0000FFFF0000FFFFull);
    s = (s & 0xFFFFFFFFull) + (s >> 32);
    return (int32_t)s - 1020;
}

/* reflect a q8 coordinate into [0, extent_q8) */
EVK_HD int64_t evk_reflect_q8(int64_t pos, int64_t extent_q8) {
    int64_t period = 2 * extent_q8;
    int64_t r = pos % period;
    if (r < 0) r += period;
    if (r >= extent_q8) r = period - 1 - r;
    return r;
}

EVK_HD evk_event evk_synth_event(const evk_synth_params* sp, uint64_t i) {
    evk_event e;
    const uint64_t u0 = evk_sm64(sp->seed ^ i);
    const uint64_t u1 = evk_sm64(u0);
    const uint64_t u2 = evk_sm64(u1);
    const int64_t t = (int64_t)((i * 1000000ull) / sp->rate_eps);
    int32_t x, y;
    if ((int32_t)(u0 & 0xFFFFu) < sp->noise_q16) {
        x = (int32_t)(((u1 & 0xFFFFFFFFull) * (uint64_t)sp->width) >> 32);
        y = (int32_t)(((u1 >> 32) * (uint64_t)sp->height) >> 32);
    } else {
        const uint32_t m = (uint32_t)((((u0 >> 16) & 0xFFFFu) * (uint64_t)sp->n_blobs) >> 16);
        const uint64_t b0 = evk_sm64(sp->seed ^ 0xB10B00000000ull ^ ((uint64_t)m << 8));
        const uint64_t b1 = evk_sm64(b0);
        const int64_t wq = (int64_t)sp->width * 256, hq = (int64_t)sp->height * 256;
        const int64_t c0x = (int64_t)(((b0 & 0xFFFFFFFFull) * (uint64_t)wq) >> 32);
        const int64_t c0y = (int64_t)(((b0 >> 32) * (uint64_t)hq) >> 32);
        const int64_t vspan = 2 * (int64_t)sp->vmax_pps * 256 + 1; /* q8 px per second */
        const int64_t vx = (int64_t)(((b1 & 0xFFFFFFFFull) * (uint64_t)vspan) >> 32) -
                           (int64_t)sp->vmax_pps * 256;
        const int64_t vy =
            (int64_t)(((b1 >> 32) * (uint64_t)vspan) >> 32) - (int64_t)sp->vmax_pps * 256;
        const int64_t cx = evk_reflect_q8(c0x + (vx * t) / 1000000, wq);
        const int64_t cy = evk_reflect_q8(c0y + (vy * t) / 1000000, hq);
        const int64_t px = cx + ((int64_t)evk_ih8(u1) * sp->sigma_q8) / 209;
        const int64_t py = cy + ((int64_t)evk_ih8(u2) * sp->sigma_q8) / 209;
        x = (int32_t)(px >> 8);
        y = (int32_t)(py >> 8);
    }
    if (x < 0) x = 0;
    if (x >= sp->width) x = sp->width - 1;
    if (y < 0) y = 0;
    if (y >= sp->height) y = sp->height - 1;
    e.x = (uint16_t)x;
    e.y = (uint16_t)y;
    e.p = (int16_t)((u0 >> 32) & 1u);
    e._pad = 0;
    e.t = t;
    return e;
}

#endif /* EVK_SYNTH_H_ */
