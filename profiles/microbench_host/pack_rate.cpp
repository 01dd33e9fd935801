// Host-side microbenchmark: how fast can T threads pack 16-byte EventCD records into 8-byte records
// (x | y << 16, (t - t_base) << 1 | p)?  Decides whether packing in front of the PCIe copy pays
// (the copy runs at ~55 GB/s of 16-byte records).  g++ -O3 -march=native -pthread
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
struct Ev { uint16_t x, y; int16_t p; uint16_t pad; int64_t t; };
static uint32_t pack(const Ev* in, uint64_t* out, size_t n, int64_t tb) {
    uint64_t bad = 0;
    for (size_t i = 0; i < n; i++) {
        uint32_t w0;
        memcpy(&w0, &in[i], 4);
        const uint64_t dt = (uint64_t)(in[i].t - tb);
        bad |= (dt >> 31) | (uint64_t)(uint16_t)in[i].p >> 1 | in[i].pad;
        out[i] = (uint64_t)w0 | ((dt << 1 | (uint64_t)(in[i].p & 1)) << 32);
    }
    return bad != 0;
}
int main(int argc, char** argv) {
    const size_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 100000000ull;
    Ev* in = (Ev*)aligned_alloc(4096, n * sizeof(Ev));
    uint64_t* out = (uint64_t*)aligned_alloc(4096, n * 8);
    const unsigned hw = std::thread::hardware_concurrency();
    printf("hardware_concurrency %u\n", hw);
    {   // first touch + fill in parallel
        std::vector<std::thread> th;
        for (unsigned t = 0; t < hw; t++)
            th.emplace_back([&, t] {
                for (size_t i = n * t / hw; i < n * (t + 1) / hw; i++) {
                    in[i] = Ev{(uint16_t)(i * 7 % 1280), (uint16_t)(i * 13 % 720), (int16_t)(i & 1), 0, (int64_t)(i / 100)};
                    out[i] = 0;
                }
            });
        for (auto& x : th) x.join();
    }
    for (unsigned T : {1u, 2u, 4u, 8u, 16u, 32u, 64u}) {
        if (T > hw) break;
        double best = 1e9;
        for (int rep = 0; rep < 3; rep++) {
            auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> th;
            std::vector<uint32_t> bad(T);
            for (unsigned t = 0; t < T; t++)
                th.emplace_back([&, t] {
                    const size_t a = n * t / T, b = n * (t + 1) / T;
                    bad[t] = pack(in + a, out + a, b - a, in[a].t);
                });
            for (auto& x : th) x.join();
            double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (s < best) best = s;
        }
        printf("threads %2u: %.2f ms, %.1f GB/s of records read, %.1f Mev/s\n", T, best * 1e3, n * 16 / best / 1e9, n / best / 1e6);
    }
    // plain memcpy for reference
    {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        const unsigned T = hw < 16 ? hw : 16;
        for (unsigned t = 0; t < T; t++)
            th.emplace_back([&, t] { const size_t a = n * t / T, b = n * (t + 1) / T; memcpy(out + a / 2, in + a / 2, (b - a) * 8); });
        for (auto& x : th) x.join();
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("memcpy %u threads of %zu MB: %.2f ms, %.1f GB/s\n", T, n * 8 >> 20, s * 1e3, n * 8 / s / 1e9);
    }
    return 0;
}
