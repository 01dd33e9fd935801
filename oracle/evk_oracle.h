/* evk_oracle.h — CPU ORACLE for the downsample + k-means hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may build, load or call it.  The product
 * library (libevk.so) never links or calls anything in oracle/ and has no CPU fallback.
 *
 * It restates, sequentially and in plain C, the algorithm of the reference's OpenCL kernels
 * and host loops (paths relative to the reference root):
 *   ACCEL = event-cam-clustering-accel/event-cam-clustering-downsampling-accel
 *   KM    = event-cam-clustering-accel/event-cam-k-means-clustering
 *   FCT   = event-cam-tracking/event-cam-fast-corner-tracker
 *
 * PARITY PINNING.  The reference holds no test, golden vector or expected value for this path
 * (SURVEY.md 4, 8c), and its host programs cannot be built here (OpenCL headers + ICD, Metavision
 * SDK, Eigen, OpenCV C++ are absent).  Its two OpenCL kernel files, however, compile AS C where
 * they lie through oracle/cl_shim.h (Makefile target `ref` -> oracle/_ref/libref.so), and the
 * oracle is pinned against that -- the reference's own code run here:
 * tests/test_reference_kernels.py compares, on random launches and on fixtures F1-F3 of SURVEY.md
 * 8c (the k-means data of KM/assign_to_centers2.c:121-131, the all-zero warm-up launch of
 * ACCEL/store.cpp:209-215,317-326, the reference's event dump
 * optics-clustering/test/event_raw_data8.csv), the kernels with the literal functions below AND
 * with the contract functions; the vectors it produced are committed as
 * tests/golden/ref_kernel_golden.json (tests/golden/make_ref_golden.py).  The generalisation the
 * reference does not have (64-bit voxel keys, time bins, polarity) is pinned by the known answers
 * SURVEY.md 8c derives and by an independent numpy restatement (tests/golden/make_golden.py).
 */
#ifndef EVK_ORACLE_H_
#define EVK_ORACLE_H_

#include <stddef.h>
#include <stdint.h>
#include "../include/evk.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- literal restatements (the reference as written, executed sequentially) ---------------- */

/* ACCEL/build/coordinate_processor.cl:16-89 for work-items i = 0..total_coords-1 in order.
 * Counters are ADDED to (atomic_add at :84-87; the host never resets them).  Returns the number of
 * unique pairs written by this call. */
int orc_ref_process_coordinates(const int* input_coords, int total_coords, int* unique_coords,
                                int* repeated_count, int* unique_count);

/* FCT/metavision_time_surface_periodic.cpp:56-120 (findCoordinate / analyzeCoordinates):
 * exact unique (x,y) in first-seen order with counts.  n_ints = number of ints in data. */
int orc_ref_analyze_coordinates(const int* data, int n_ints, int* xs, int* ys, int* counts);

/* One trip round KERNEL_RESTART of KM/assign_to_centers2.c:184-548 exactly as written
 * (quirks D11-D15 of SURVEY.md appendix A): assign_to_centers with sqrt distances and 2k / 255
 * labels (assign_to_centers.cl:1-34), scatter into 4096-float slabs in gid order (:36-119),
 * 1024-wide fp32 tree sums (:121-140), centroid formula with the stride-2 indexing
 * (assign_to_centers2.c:507-512) and the selective overwrite with integer abs() (:525-531).
 * data: 4096 floats; centroids: 16 floats (updated in place as the reference does);
 * output: 32768 floats carried across calls (the reference re-uploads it, :195);
 * returns error_max. */
float orc_ref_kmeans_trip(const float* data, float* centroids, float* output, int* assign,
                          int* cluster_index, float* scalar_sum, float* new_centroids);

/* ---- contract semantics (SURVEY.md 8a "Semantics contract") -------------------------------- */

/* Voxel / reference-hash key of one event; returns 0 when the event is gated out. */
int orc_event_key(const evk_event* e, const evk_ds_params* p, uint64_t* key);

/* Sequential downsample: for i = 0..n-1 keep the first event of every distinct key.
 * Output in canonical order (ascending first stream index).  Returns U.
 * keys / first_idx must hold n entries; n_repeated (= #keys hit >= 2 times) may be NULL. */
size_t orc_downsample(const evk_event* ev, size_t n, const evk_ds_params* p, uint64_t* keys,
                      uint32_t* first_idx, size_t* n_repeated);

/* Same result as orc_downsample computed with `threads` OpenMP threads (CPU baseline with all
 * host cores).  canonical = 0 skips the final sort by first index (set semantics only). */
size_t orc_downsample_mt(const evk_event* ev, size_t n, const evk_ds_params* p, int threads,
                         int canonical, uint64_t* keys, uint32_t* first_idx, size_t* n_repeated);

/* Point features of the representatives: x, y, (t - t0) * t_scale, pbit * p_scale. */
void orc_points(const evk_event* ev, const uint32_t* first_idx, size_t U, int D, int64_t t0_us,
                float t_scale, float p_scale, float* pts);

/* KM/assign_to_centers.cl:1-34 under the contract: fp32 squared distance built with explicit
 * single-rounding fmaf, strict '<' (lowest k wins ties), gate max_dist (<= 0 / inf: none),
 * labels k or -1.  use_sqrt = 1 compares length() like the reference (:17-21). */
void orc_kmeans_assign(const float* pts, size_t P, int D, const float* cent, int K,
                       float max_dist, int use_sqrt, int32_t* labels);

/* Centroid update: exact sums, centroid = (float)(sum / count); empty cluster keeps its centroid.
 * sums (K*D doubles) and counts (K) may be NULL. Returns max_k |delta c_k|_inf. */
float orc_kmeans_update(const float* pts, size_t P, int D, const int32_t* labels, int K,
                        float* cent, uint64_t* counts, double* sums);

/* Lloyd loop: iters maximum, stop when shift <= tol (tol < 0: never). Returns iterations run. */
int orc_kmeans(const float* pts, size_t P, int D, float* cent, int K, float max_dist, int iters,
               float tol, int32_t* labels, uint64_t* counts);
int orc_kmeans_mt(const float* pts, size_t P, int D, float* cent, int K, float max_dist,
                  int iters, float tol, int threads, int32_t* labels, uint64_t* counts);

/* ---- inputs -------------------------------------------------------------------------------- */
void orc_synth(const evk_synth_params* sp, evk_event* out);
void orc_synth_mt(const evk_synth_params* sp, evk_event* out, int threads);
/* rows "x,y,t,p"; returns the number of events read (<= cap) or -1 */
long orc_load_csv(const char* path, evk_event* out, size_t cap);
int orc_max_threads(void);

/* ---- RAW EVT 2.0 (third-party format: Prophesee Metavision "EVT 2.0", the payload of the
 * `traffic_data.raw`-style recordings the reference opens with Metavision::Camera::from_file,
 * ACCEL/store.cpp:336; the SDK itself is absent from /root/reference, version unpinned, so the
 * published word layout is restated here) -------------------------------------------------------
 * 32-bit little-endian words, type in bits 31..28:
 *   0x0 CD_OFF / 0x1 CD_ON : [27:22] timestamp bits 5..0, [21:11] x, [10:0] y
 *   0x8 EVT_TIME_HIGH      : [27:0]  timestamp bits 33..6   (t_us = time_high << 6 | low 6 bits)
 *   0xA EXT_TRIGGER, 0xE OTHERS, 0xF CONTINUED : carry no CD event (skipped)
 * decode: sequential scan keeping the last EVT_TIME_HIGH (0 before the first one); returns the
 * number of CD events written (<= cap).  encode: one EVT_TIME_HIGH whenever bits 33..6 change
 * (and at the start); events need x, y < 2048 and 0 <= t < 2^34; returns words written (<= cap)
 * or (size_t)-1 when an event does not fit the format. */
size_t orc_evt2_decode(const uint32_t* words, size_t n_words, evk_event* out, size_t cap);
size_t orc_evt2_encode(const evk_event* ev, size_t n, uint32_t* words, size_t cap);
/* RAW EVT 3.0 (16-bit words; see evk_oracle.c): decode returns the number of CD events in the stream
 * (which may exceed cap), encode the number of words or (size_t)-1 */
/* time surface + corner test of the reference's corner tracker (see evk_oracle.c) */
size_t orc_ts_corners(const evk_event* ev, size_t n, int W, int H, int64_t* surface,
                      int literal_break, uint8_t* flags);
size_t orc_evt3_decode(const uint16_t* words, size_t n_words, evk_event* out, size_t cap);
size_t orc_evt3_encode(const evk_event* ev, size_t n, uint16_t* words, size_t cap);

#ifdef __cplusplus
}
#endif
/* box non-maximum suppression of a corner list (the reference's CornerFilter::filterCorners) */
size_t orc_filter_corners(const int32_t* xy, size_t n, int width, int height, int box_size,
                          uint32_t* kept);

#endif
