// evk.hpp — C++17 RAII view of the C-ABI (include/evk.h) with the reference's call shape.
//
// The reference wires everything inside main(): an SDK callback packs events
// (ACCEL/store.cpp:570-611), a reslicer fires on_new_slice every 50 ms (:349-352,370), which
// uploads, launches process_coordinates, waits, reads back (:389-430) and hands the unique
// coordinates to the consumer (:435-445).  `evk::Pipeline` keeps that order:
//   add_events(begin,end)  <->  cam.cd().add_callback / reslicer.process_events  (:614-615)
//   on_new_slice(fn)       <->  reslicer.set_on_new_slice_callback               (:370)
// with the slice body running on the GPU through libevk.so.  Errors are exceptions carrying the
// evk_status (the reference calls exit(1)).
#pragma once

#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/evk.h"

namespace evk {

struct Error : std::runtime_error {
    int status;
    Error(int st, const std::string& what) : std::runtime_error(what), status(st) {}
};

class Handle {
  public:
    explicit Handle(std::size_t max_events, int device = 0) {
        int st = evk_create(&h_, device, max_events);
        if (st != EVK_OK) throw Error(st, "evk_create failed (no CUDA device? there is no CPU path)");
    }
    ~Handle() { evk_destroy(h_); }
    Handle(const Handle&) = delete;
    Handle& operator=(const Handle&) = delete;
    evk_handle* get() const { return h_; }

    void check(int st) const {
        if (st != EVK_OK) throw Error(st, evk_last_error(h_));
    }
    // load events
    void load_events(const evk_event* b, const evk_event* e) { check(evk_load_events(h_, b, e)); }
    void append_events(const evk_event* b, const evk_event* e) { check(evk_append_events(h_, b, e)); }
    void load_csv(const std::string& path) { check(evk_load_csv(h_, path.c_str())); }
    void synth(const evk_synth_params& sp) { check(evk_synth(h_, &sp)); }
    // downsample
    struct Counts {
        std::size_t unique = 0, repeated = 0;
    };
    Counts downsample(const evk_ds_params& p) {
        Counts c;
        check(evk_downsample(h_, &p, &c.unique, &c.repeated));
        n_unique_ = c.unique;
        return c;
    }
    void voxels(std::vector<uint64_t>* keys, std::vector<evk_event>* reps,
                std::vector<uint32_t>* first) {
        if (keys) keys->resize(n_unique_);
        if (reps) reps->resize(n_unique_);
        if (first) first->resize(n_unique_);
        check(evk_get_voxels(h_, keys ? keys->data() : nullptr, reps ? reps->data() : nullptr,
                             first ? first->data() : nullptr, n_unique_));
    }
    // cluster
    void set_centroids(const std::vector<float>& c, int K, int D) {
        check(evk_set_centroids(h_, c.data(), K, D));
    }
    void init_centroids_first_k(const evk_km_params& p) { check(evk_init_centroids_first_k(h_, &p)); }
    int kmeans(const evk_km_params& p) {
        int it = 0;
        check(evk_kmeans(h_, &p, &it));
        km_ = p;
        return it;
    }
    // fused step: downsample + (first-K init | warm start) + k-means in one submission
    Counts downsample_kmeans(const evk_ds_params& ds, const evk_km_params& km, bool init_first_k,
                             int* iters = nullptr) {
        Counts c;
        int it = 0;
        check(evk_downsample_kmeans(h_, &ds, &km, init_first_k ? 1 : 0, &c.unique, &c.repeated, &it));
        n_unique_ = c.unique;
        km_ = km;
        if (iters) *iters = it;
        return c;
    }
    // the same step in two halves: queue it, do host work, collect (evk_downsample_kmeans_submit/_wait)
    void submit_downsample_kmeans(const evk_ds_params& ds, const evk_km_params& km,
                                  bool init_first_k) {
        check(evk_downsample_kmeans_submit(h_, &ds, &km, init_first_k ? 1 : 0));
        km_ = km;
    }
    Counts wait_downsample_kmeans(int* iters = nullptr) {
        Counts c;
        int it = 0;
        check(evk_downsample_kmeans_wait(h_, &c.unique, &c.repeated, &it));
        n_unique_ = c.unique;
        if (iters) *iters = it;
        return c;
    }
    // labels and centroids
    std::vector<int32_t> labels() {
        std::size_t n = n_unique_;
        if (km_.on_events) check(evk_num_events(h_, &n));
        std::vector<int32_t> l(n);
        check(evk_get_labels(h_, l.data(), n));
        return l;
    }
    void centroids(std::vector<float>* c, std::vector<uint64_t>* counts) {
        c->resize(static_cast<std::size_t>(km_.K) * km_.D);
        counts->resize(km_.K);
        check(evk_get_centroids(h_, c->data(), counts->data()));
    }
    std::size_t n_unique() const { return n_unique_; }

    // ---- consumer: the reference's `AEClustering *eclustering` (ACCEL/store.cpp:42) on the device
    void aec_create(const evk_aec_params& p) { check(evk_aec_create(h_, &p)); }
    // eclustering->update(ev_data) for n events {t, x, y, p} (AEClustering.h:37)
    void aec_update(const double* e, std::size_t n) { check(evk_aec_update(h_, e, n)); }
    // the hand-off loop of the slice callback (ACCEL/store.cpp:435-445) without leaving the device
    void aec_update_voxels(double t, std::size_t start, std::size_t step, std::size_t count) {
        check(evk_aec_update_voxels(h_, t, start, step, count));
    }
    std::vector<evk_aec_cluster> aec_clusters(int* last_updated = nullptr) {
        std::size_t n = 0;
        check(evk_aec_get_clusters(h_, nullptr, 0, &n, last_updated));
        std::vector<evk_aec_cluster> c(n);
        if (n) check(evk_aec_get_clusters(h_, c.data(), n, &n, last_updated));
        return c;
    }
    // centroid + flow arrow per cluster with >= minN events (ACCEL/store.cpp:461-521)
    std::vector<evk_aec_flow> aec_report() {
        std::vector<evk_aec_flow> f(1024);
        std::size_t n = 0;
        check(evk_aec_report(h_, f.data(), f.size(), &n));
        f.resize(n);
        return f;
    }

  private:
    evk_handle* h_ = nullptr;
    std::size_t n_unique_ = 0;
    evk_km_params km_{};
};

// One slice result, what the reference's on_new_slice body has in hand at ACCEL/store.cpp:430-460
struct Slice {
    int64_t t_begin_us = 0;
    std::size_t n_events = 0, n_unique = 0, n_repeated = 0;
    std::vector<float> centroids;   // K x D
    std::vector<uint64_t> counts;   // K
    std::vector<evk_aec_flow> flow; // enable_aec(): one centroid + arrow per reported cluster
};

// Replays the reference's streaming structure: events arrive in arbitrary chunks, a slice is
// closed every `slice_us` of event time, the slice callback sees the clustered result.
class Pipeline {
  public:
    Pipeline(std::size_t max_events_per_slice, const evk_ds_params& ds, const evk_km_params& km,
             int64_t slice_us = 50000, int device = 0)
        : h_(max_events_per_slice, device), ds_(ds), km_(km), slice_us_(slice_us) {}

    void on_new_slice(std::function<void(const Slice&)> fn) { cb_ = std::move(fn); }

    // Attach the reference's consumer (ACCEL/store.cpp:42,435-521): after every slice the unique
    // voxels are handed to the asynchronous event clustering on the device and Slice::flow holds
    // its report.  literal_stride reproduces the hand-off as written (every 2nd pair, a quarter of
    // the slice); otherwise every voxel is handed over.
    // max_per_slice bounds the hand-off (the reference's kernel emits at most 8192 pairs per
    // launch, so its loop makes at most 2048 updates per slice).
    void enable_aec(const evk_aec_params& p, bool literal_stride = false,
                    std::size_t max_per_slice = 2048) {
        h_.aec_create(p);
        aec_ = true;
        aec_literal_ = literal_stride;
        aec_max_ = max_per_slice;
    }

    // the event callback: a borrowed range, valid only during the call
    void add_events(const evk_event* begin, const evk_event* end) {
        for (const evk_event* e = begin; e != end; ++e) {
            if (!started_) {
                started_ = true;
                t0_ = e->t - (e->t % slice_us_);
            }
            while (e->t >= t0_ + slice_us_) close_slice();
            buf_.push_back(*e);
        }
    }
    void flush() {
        if (!buf_.empty()) close_slice();
    }
    Handle& handle() { return h_; }

  private:
    void close_slice() {
        if (!buf_.empty()) {
            Slice s;
            s.t_begin_us = t0_;
            s.n_events = buf_.size();
            h_.load_events(buf_.data(), buf_.data() + buf_.size());
            evk_ds_params ds = ds_;
            ds.t0_us = t0_;
            // one fused submission per slice; later slices start warm.  A first slice with fewer
            // voxels than clusters is only downsampled (seeding waits for the next one).
            Handle::Counts c;
            try {
                c = h_.downsample_kmeans(ds, km_, !warm_);
                warm_ = true;
                h_.centroids(&s.centroids, &s.counts);
            } catch (const Error& e) {
                if (warm_ || e.status != EVK_ERR_INVALID) throw;
                c = h_.downsample(ds);
            }
            s.n_unique = c.unique;
            s.n_repeated = c.repeated;
            if (aec_) {  // pseudo-time = cumulative unique count / 1000.0 (ACCEL/store.cpp:440)
                unique_total_ += c.unique;
                const double t = static_cast<double>(unique_total_) / 1000.0;
                std::size_t cnt = aec_literal_ ? (c.unique + 3) / 4 : c.unique;
                if (cnt > aec_max_) cnt = aec_max_;
                h_.aec_update_voxels(t, 0, aec_literal_ ? 2 : 1, cnt);
                s.flow = h_.aec_report();
            }
            if (cb_) cb_(s);
            buf_.clear();
        }
        t0_ += slice_us_;
    }
    Handle h_;
    evk_ds_params ds_;
    evk_km_params km_;
    int64_t slice_us_, t0_ = 0;
    bool started_ = false, warm_ = false, aec_ = false, aec_literal_ = false;
    std::size_t unique_total_ = 0, aec_max_ = 2048;
    std::vector<evk_event> buf_;
    std::function<void(const Slice&)> cb_;
};

}  // namespace evk
