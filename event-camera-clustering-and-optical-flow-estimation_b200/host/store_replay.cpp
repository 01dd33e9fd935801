// store_replay.cpp — the reference's streaming sample without the SDK: replays a CSV dump
// ("x,y,t,p" rows) or a synthetic stream in 50 ms slices through evk::Pipeline and prints one line
// per slice (events, unique voxels, repeated, centroids of the K clusters) and, with a third argument
// "aec", the report of the reference's own consumer (AEClustering: centroid + flow arrow per
// cluster, ACCEL/store.cpp:461-521) computed on the device.
//
// Mirrors main() of ACCEL/metavision_sdk_get_started5_opencl_store.cpp:179-643:
//   Camera::from_file(argv[1])              -> CSV path in argv[1] (or "synth:<n_events>")
//   cam.cd().add_callback(aggregate_events) -> pipe.add_events(chunk)   (uneven chunks, :614-615)
//   reslicer on_new_slice (50 ms)           -> pipe.on_new_slice(...)   (:370)
// Build: g++ -std=c++17 store_replay.cpp -L.. -levk -o store_replay   (see build.py)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/evk_synth.h"
#include "evk.hpp"

static std::vector<evk_event> read_csv(const char* path) {
    std::vector<evk_event> ev;
    FILE* f = std::fopen(path, "r");
    if (!f) return ev;
    char line[256];
    while (std::fgets(line, sizeof line, f)) {
        long x, y, p;
        long long t;
        if (std::sscanf(line, "%ld,%ld,%lld,%ld", &x, &y, &t, &p) != 4) continue;
        evk_event e{};
        e.x = (uint16_t)x;
        e.y = (uint16_t)y;
        e.p = (int16_t)p;
        e.t = (int64_t)t;
        ev.push_back(e);
    }
    std::fclose(f);
    return ev;
}

int main(int argc, char** argv) {
    const char* src = argc > 1 ? argv[1] : "synth:2000000";
    const int K = argc > 2 ? std::atoi(argv[2]) : 8;
    std::vector<evk_event> ev;
    int W = 1280, H = 720;
    if (std::strncmp(src, "synth:", 6) == 0) {
        evk_synth_params sp{};
        sp.seed = 0xE7CA0005;
        sp.n_events = std::strtoull(src + 6, nullptr, 10);
        sp.rate_eps = 10000000;
        sp.width = W;
        sp.height = H;
        sp.n_blobs = K;
        sp.sigma_q8 = 1536;
        sp.noise_q16 = 16384;
        sp.vmax_pps = 200;
        ev.resize(sp.n_events);
        for (uint64_t i = 0; i < sp.n_events; i++) ev[i] = evk_synth_event(&sp, i);
    } else {
        ev = read_csv(src);
    }
    if (ev.empty()) {
        std::fprintf(stderr, "no events from %s\n", src);
        return 1;
    }
    evk_ds_params ds{};
    ds.width = W;
    ds.height = H;
    ds.vx = ds.vy = 4;
    ds.vt_us = 1000;
    ds.use_polarity = 1;
    ds.keyfn = EVK_KEY_VOXEL;
    ds.algo = EVK_ALGO_AUTO;
    ds.count_repeated = 1;
    evk_km_params km{};
    km.K = K;
    km.D = 2;
    km.max_dist = 0.f;
    km.iters = 2;
    km.tol = -1.f;
    km.t_scale = 1e-3f;
    km.p_scale = 1.f;
    try {
        evk::Pipeline pipe(ev.size(), ds, km, 50000);
        int n_slices = 0;
        if (argc > 3 && std::strcmp(argv[3], "aec") == 0) {
            evk_aec_params ap{};  // use_init = 0: the default-constructed consumer of the reference
            ap.rand_seed = 1;
            ap.max_points = 1 << 15;
            pipe.enable_aec(ap, /*literal_stride=*/true);
        }
        pipe.on_new_slice([&](const evk::Slice& s) {
            std::printf("slice %d t=%lld us events=%zu unique=%zu repeated=%zu", n_slices++,
                        (long long)s.t_begin_us, s.n_events, s.n_unique, s.n_repeated);
            for (size_t k = 0; k < s.counts.size() && k < 4; k++)
                std::printf(" c%zu=(%.1f,%.1f)x%llu", k, s.centroids[2 * k], s.centroids[2 * k + 1],
                            (unsigned long long)s.counts[k]);
            std::printf("\n");
            for (const evk_aec_flow& f : s.flow)
                if (f.has_arrow)
                    std::printf("  cluster %d n=%d (%.1f,%.1f) -> (%.1f,%.1f)\n", f.id, f.n,
                                f.prev[0], f.prev[1], f.arrow_end[0], f.arrow_end[1]);
        });
        // the SDK delivers events in uneven buffers; do the same
        size_t at = 0, chunk = 1000;
        while (at < ev.size()) {
            size_t n = std::min(chunk, ev.size() - at);
            pipe.add_events(ev.data() + at, ev.data() + at + n);
            at += n;
            chunk = chunk * 3 / 2 + 17;
            if (chunk > 300000) chunk = 1000;
        }
        pipe.flush();
        std::printf("%d slices\n", n_slices);
    } catch (const evk::Error& e) {
        std::fprintf(stderr, "evk error %d: %s\n", e.status, e.what());
        return 2;
    }
    return 0;
}
