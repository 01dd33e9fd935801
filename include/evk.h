/* evk.h — C-ABI boundary of the B200-native event-cloud downsample + k-means path.
 *
 * This header is the drop-in boundary for ONE hot path of the reference
 * (LogicTronixInc/Event-Camera-Clustering-and-Optical-Flow-Estimation):
 *
 *     load events -> voxel-hash downsample -> k-means assign + centroid update
 *                 -> return labels and centroids
 *
 * and, widening into the path's callers on either side: RAW EVT 2.0 / 3.0
 * ingest (evk_load_evt2 / _evt3 / _raw), the reference's consumer of the downsampled coordinates
 * (evk_aec_*: asynchronous event clustering + per-slice flow arrows), the corner tracker's time
 * surface + corner test (evk_ts_*) and DBSCAN on the downsampled cloud (evk_dbscan_*).
 *
 * The reference has no library / plugin API for this path: every call is inlined in main()
 * of three sample programs.  Each entry point below therefore cites the reference call site it
 * replaces (paths relative to the reference root; abbreviations:
 *   ACCEL = event-cam-clustering-accel/event-cam-clustering-downsampling-accel
 *   SAMP  = event-cam-pre-processing-opencl/event-cam-sampling
 *   KM    = event-cam-clustering-accel/event-cam-k-means-clustering
 *   store.cpp = metavision_sdk_get_started5_opencl_store.cpp ).
 *
 * Conventions
 *  - plain C, plain pointers and sizes; no C++ / torch / CUDA types in any signature;
 *  - every function returns EVK_OK (0) or a negative evk_status; nothing calls exit() and no
 *    exception crosses the boundary (the reference does perror()+exit(1), ACCEL/store.cpp:64-68);
 *  - the caller owns every host pointer; the handle owns all device memory (arena sized in
 *    evk_create; the hot calls never allocate — the reference leaks a cl_mem per slice,
 *    ACCEL/store.cpp:389);
 *  - one handle is used from one thread at a time (the reference does pack -> launch -> wait ->
 *    consume on the single SDK decoding thread, ACCEL/store.cpp:370-615);
 *  - calls return when their results are complete, except the *_submit halves of the fused step,
 *    which only queue it on the handle's stream; the matching *_wait (or any getter) collects it;
 *  - there is NO CPU fallback: without a CUDA device evk_create fails with EVK_ERR_CUDA.
 */
#ifndef EVK_H_
#define EVK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define EVK_API
#else
#define EVK_API __attribute__((visibility("default")))
#endif

typedef enum evk_status {
    EVK_OK = 0,
    EVK_ERR_INVALID = -1,  /* bad argument / parameter combination                      */
    EVK_ERR_CUDA = -2,     /* CUDA runtime error (see evk_last_error)                    */
    EVK_ERR_NOMEM = -3,    /* device or host allocation failed                           */
    EVK_ERR_STATE = -4,    /* call order violated (e.g. kmeans before downsample)        */
    EVK_ERR_CAPACITY = -5, /* more events / voxels than the handle was created for       */
    EVK_ERR_IO = -6,       /* file could not be read / parsed                            */
    EVK_ERR_COMM = -7      /* NCCL error or communicator not initialised                 */
} evk_status;

/* 16-byte packed event == layout of Metavision::EventCD {uint16 x; uint16 y; int16 p; int64 t}
 * as consumed at ACCEL/store.cpp:575-590 (ev->x, ev->y, ev->t, ev->p).  Kernels read this record
 * directly with one 128-bit load; the reference's int[2] repack loop (ACCEL/store.cpp:587-599) is
 * gone. */
typedef struct evk_event {
    uint16_t x, y;
    int16_t p;      /* polarity: > 0 is "on" */
    uint16_t _pad;
    int64_t t;      /* microseconds */
} evk_event;

enum { EVK_KEY_VOXEL = 0, EVK_KEY_REF_HASH8192 = 1 };
/* downsample algorithm selector */
enum {
    EVK_ALGO_AUTO = 0,  /* time-slab kernel when the stream allows it, else partition, else table */
    EVK_ALGO_TABLE = 1, /* global open-addressing table, 64-bit keys, atomicCAS insert       */
    EVK_ALGO_SORT = 2,  /* radix sort + unique (cross-check variant)                         */
    EVK_ALGO_SLAB = 3,  /* per-time-bin shared-memory kernel; a stream that is not partitioned by
                           time bin fails over to PARTITION, then to TABLE                  */
    EVK_ALGO_PARTITION = 4 /* reported in ds_algo_used (also selectable): stable partition of
                           the stream by time bin on the device, then the time-slab kernel  */
};

/* Downsample parameters.  Replaces the compile-time constants of the reference kernel
 * (sensor bounds, bucket count: ACCEL/build/coordinate_processor.cl:7-12,29-30,56) and the unused
 * width/height kernel arguments (:24-25).
 *   keyfn = EVK_KEY_VOXEL:  key = ((tbin*NY + ybin)*NX + xbin) * (use_polarity?2:1) + pbit
 *           xbin = x / vx, ybin = y / vy, tbin = (t - t0_us) / vt_us (vt_us <= 0: one bin),
 *           NX = ceil(width / vx), NY = ceil(height / vy); gate x < width, y < height, t >= t0_us.
 *   keyfn = EVK_KEY_REF_HASH8192: key = (1619*x + 31*y) mod 8192, gate x <= width, y <= height
 *           (inclusive as coordinate_processor.cl:56); vx, vy, vt_us, t0_us, use_polarity ignored. */
typedef struct evk_ds_params {
    int32_t width, height;
    int32_t vx, vy;
    int64_t vt_us;
    int64_t t0_us;
    int32_t use_polarity;
    int32_t keyfn;
    int32_t algo;
    int32_t count_repeated; /* 0: n_repeated is not computed (returned as 0) */
} evk_ds_params;

/* K-means parameters.  Replaces K=8 / D=2 / threshold 50 hard-wired in
 * KM/assign_to_centers.cl:11,14 and the error_max > 10 loop of KM/assign_to_centers2.c:545.
 * Point features, fixed order: x, y, (t - t0_us) * t_scale, pbit * p_scale  (D = 2, 3 or 4). */
typedef struct evk_km_params {
    int32_t K, D;
    float max_dist; /* <= 0 or +inf: no gate; reference: 50.0f */
    int32_t iters;  /* maximum Lloyd iterations (>= 1) */
    float tol;      /* stop when max_k |delta c_k|_inf <= tol; < 0: always run `iters` */
    float t_scale, p_scale;
    int32_t on_events; /* 0: cluster the downsampled voxel representatives; 1: the raw events */
} evk_km_params;

/* Synthetic stream.  Counter-based: event i depends on (seed, i) only, so any
 * index shard can be generated independently, bit-identically on CPU and GPU (evk_synth.h). */
typedef struct evk_synth_params {
    uint64_t seed;
    uint64_t first_index; /* global index of the first generated event */
    uint64_t n_events;    /* number of events to generate            */
    uint64_t rate_eps;    /* events per second: t_i = i * 1e6 / rate */
    int32_t width, height;
    int32_t n_blobs;      /* moving Gaussian blobs                   */
    int32_t sigma_q8;     /* blob sigma in 1/256 px (6 px = 1536)    */
    int32_t noise_q16;    /* P(background event) * 65536 (25 % = 16384) */
    int32_t vmax_pps;     /* max |blob velocity| per axis, px/s      */
} evk_synth_params;

/* Per-stage device times of the last evk_downsample / evk_kmeans call, in milliseconds, measured
 * with CUDA events on the handle's stream when profiling is enabled.  Mirrors the reference's
 * CL_QUEUE_PROFILING_ENABLE + clGetEventProfilingInfo use (ACCEL/store.cpp:277-278,450-454). */
typedef struct evk_stage_times {
    float ds_total_ms;    /* all downsample kernels                          */
    float ds_main_ms;     /* dominant downsample kernel (insert or slab)     */
    float ds_compact_ms;  /* table compaction (TABLE) / 0                    */
    float km_total_ms;    /* all k-means kernels of the last evk_kmeans call */
    float km_assign_ms;   /* sum over iterations of the fused assign+accumulate kernel */
    int32_t ds_algo_used; /* EVK_ALGO_* actually run                         */
    int32_t km_iters;
    int32_t ds_launches;  /* kernels launched by the last evk_downsample     */
    int32_t km_launches;  /* kernels launched by the last evk_kmeans         */
} evk_stage_times;

typedef struct evk_handle evk_handle;

/* ---- lifetime ------------------------------------------------------------------------------ */
/* Replaces create_device/clCreateContext/build_program/clCreateBuffer x5 (ACCEL/store.cpp:55-139,
 * 237-268; KM/assign_to_centers2.c:151-197).  Kernels are precompiled for sm_100a. */
EVK_API int evk_create(evk_handle** out, int device, size_t max_events);
EVK_API int evk_destroy(evk_handle* h);
EVK_API const char* evk_last_error(const evk_handle* h); /* never NULL */
EVK_API const char* evk_version(void);

/* ---- load events --------------------------------------------------------------------------- */
/* Replaces the event callback + packing loop aggregate_events_fct(begin,end)
 * (ACCEL/store.cpp:570-611, SAMP/store.cpp:419-460) and the per-slice COPY_HOST_PTR upload
 * (ACCEL/store.cpp:389).  Borrowed host range, copied H2D; the stream held by the handle is
 * replaced. evk_append_events adds to it (the callback may fire several times per slice). */
EVK_API int evk_load_events(evk_handle* h, const evk_event* begin, const evk_event* end);
EVK_API int evk_append_events(evk_handle* h, const evk_event* begin, const evk_event* end);
EVK_API int evk_load_events_soa(evk_handle* h, const uint16_t* x, const uint16_t* y,
                                const int64_t* t, const uint8_t* p, size_t n);
/* Reference-literal input: interleaved int32 (x0,y0,x1,y1,...) exactly as the kernel argument
 * input_coords of process_coordinates (coordinate_processor.cl:17-18); t = 0, p = 0.
 * Coordinates outside [0,65535] are carried as invalid events (always gated out). */
EVK_API int evk_load_coords_i32(evk_handle* h, const int32_t* xy, size_t n_pairs);
/* rows "x,y,t,p" as ACCEL-era dumps (optics-clustering/test/event_raw_data8.csv) */
EVK_API int evk_load_csv(evk_handle* h, const char* path);
/* RAW EVT 2.0 words (Prophesee; the payload of the recordings the reference opens with
 * Metavision::Camera::from_file(argv[1]), ACCEL/store.cpp:336, e.g. `traffic_data.raw`,
 * ACCEL/Readme.md:21).  Replaces the SDK's CPU decoder in front of the event callback (:614-615):
 * the 4-byte words are uploaded as they are and decoded to the 16-byte records on the device.
 * Word layout: type = bits 31..28; 0x0 CD_OFF / 0x1 CD_ON: [27:22] t bits 5..0, [21:11] x,
 * [10:0] y; 0x8 EVT_TIME_HIGH: [27:0] t bits 33..6; other types carry no CD event.
 * n_events (may be NULL) receives the number of CD events decoded. */
EVK_API int evk_load_evt2(evk_handle* h, const uint32_t* words, size_t n_words, size_t* n_events);
/* RAW EVT 3.0 words (16 bit; the format current Prophesee sensors record): type = bits 15..12;
 * 0x0 EVT_ADDR_Y [10:0] y; 0x2 EVT_ADDR_X [11] p, [10:0] x (one event); 0x3 VECT_BASE_X [11] p,
 * [10:0] x; 0x4 VECT_12 / 0x5 VECT_8: one event at base + i per set mask bit, then base += 12 / 8;
 * 0x6 EVT_TIME_LOW t bits 11..0; 0x8 EVT_TIME_HIGH t bits 23..12 (a lower value than the previous
 * one counts a 2^24 us wrap); other types carry no CD event.  Decoded on the device as a scan over
 * the decoder state (evk_evt3.cu). */
EVK_API int evk_load_evt3(evk_handle* h, const uint16_t* words, size_t n_words, size_t* n_events);
/* a RAW file: '%'-prefixed ASCII header lines ("% evt 2.0" / "% evt 3.0" / "% format EVT3;..."),
 * then the payload in that format */
EVK_API int evk_load_raw(evk_handle* h, const char* path, size_t* n_events);
/* generate on device (benchmarks; no host copy) */
EVK_API int evk_synth(evk_handle* h, const evk_synth_params* sp);
EVK_API int evk_num_events(const evk_handle* h, size_t* n);
EVK_API int evk_get_events(evk_handle* h, evk_event* out, size_t first, size_t count);

/* ---- downsample ---------------------------------------------------------------------------- */
/* Replaces clSetKernelArg + clEnqueueNDRangeKernel(process_coordinates) + clFinish + the two
 * blocking reads of unique_count (ACCEL/store.cpp:397-430; kernel coordinate_processor.cl:16-89).
 * Counts are per call (the reference's counters are cumulative and never reset, :267-268,557).
 * n_unique / n_repeated may be NULL. */
EVK_API int evk_downsample(evk_handle* h, const evk_ds_params* p, size_t* n_unique,
                           size_t* n_repeated);
/* Replaces clEnqueueReadBuffer(unique_buffer) (ACCEL/store.cpp:412-413).  Canonical order =
 * ascending first stream index (what a sequential run of the reference kernel produces).
 * Any of keys / reps / first_idx may be NULL.  cap = capacity of the arrays in records. */
EVK_API int evk_get_voxels(evk_handle* h, uint64_t* keys, evk_event* reps, uint32_t* first_idx,
                           size_t cap);
/* The counters of the current voxel shard -- what the last evk_downsample, fused step or completed
 * window produced (the reads of unique_count / repeated_count, ACCEL/store.cpp:418-430).
 * EVK_ERR_STATE (and zeros) when there is none. */
EVK_API int evk_num_voxels(const evk_handle* h, size_t* n_unique, size_t* n_repeated);

/* ---- cluster ------------------------------------------------------------------------------- */
/* Replaces the host-initialised float centroids[16] (KM/assign_to_centers2.c:131). K*D floats. */
EVK_API int evk_set_centroids(evk_handle* h, const float* c, int K, int D);
/* Deterministic initialisation: the first K voxel representatives in canonical order. */
EVK_API int evk_init_centroids_first_k(evk_handle* h, const evk_km_params* p);
/* Replaces one or more trips round KERNEL_RESTART: assign_to_centers -> assign_data_cluster ->
 * reduction_scalar -> host centroid update (KM/assign_to_centers2.c:184-548; kernels
 * KM/assign_to_centers.cl:1-140) by a fused assign+accumulate kernel and a finalise kernel. */
EVK_API int evk_kmeans(evk_handle* h, const evk_km_params* p, int* iters_done);

/* Fused step: the same results as evk_downsample, then (init_first_k != 0) evk_init_centroids_first_k
 * or (== 0) the centroids already held by the handle (warm start), then evk_kmeans -- submitted as
 * one pass with a single host synchronisation.  On time-ordered streams with D = 2, K <= 254 and one
 * iteration the pass is ONE replayed CUDA graph of nine kernels -- time-slab downsample (bins, slab,
 * fix-up) with the centroid-only work (first-K walk, candidate lists, label map, quads) beside it
 * on a second stream, then one assign + accumulate pass over the new voxels and the finalise; the
 * voxel count never leaves the device in between.  (Assigning inside the downsample kernel was
 * measured slower: DESIGN.md section 7.)  Other shapes run the three calls one after the other
 * with the same results.  Replaces the per-slice sequence
 * launch -> clFinish -> read -> consumer of ACCEL/store.cpp:397-445 when the consumer is k-means. */
EVK_API int evk_downsample_kmeans(evk_handle* h, const evk_ds_params* ds, const evk_km_params* km,
                                  int init_first_k, size_t* n_unique, size_t* n_repeated,
                                  int* iters_done);
/* The same step in two halves, for callers that keep the device busy across slices (the reference
 * blocks its producer thread in clFinish for every slice, ACCEL/store.cpp:397-411): submit returns
 * as soon as the pass is queued on the handle's stream, wait is the step's one synchronisation and
 * reports its counts.  Steps may be queued behind each other (each one works on the events resident
 * at its turn in stream order); wait then reports the last one.  evk_downsample_kmeans ==
 * submit + wait.  Shapes the fused pass does not take run synchronously inside submit.
 * A loader called while steps are queued collects them first (their results refer back to the
 * events they ran on), so results must be read before the next load.  Queued steps need
 * time-ordered slices: if the time-slab path rejects an EARLIER queued step, wait returns
 * EVK_ERR_STATE instead of silently skipping that slice (one step at a time falls back to the
 * general path). */
EVK_API int evk_downsample_kmeans_submit(evk_handle* h, const evk_ds_params* ds,
                                         const evk_km_params* km, int init_first_k);
EVK_API int evk_downsample_kmeans_wait(evk_handle* h, size_t* n_unique, size_t* n_repeated,
                                       int* iters_done);

/* ---- return labels and centroids ----------------------------------------------------------- */
/* Replaces clEnqueueReadBuffer(assign_buffer) (KM/assign_to_centers2.c:259-265).  labels[i] in
 * [0,K) or -1 (unassigned; the reference writes 2k or 255, assign_to_centers.cl:12,22,26), in the
 * same canonical order as evk_get_voxels (or stream order when on_events = 1). */
EVK_API int evk_get_labels(evk_handle* h, int32_t* labels, size_t cap);
/* Replaces new_centroids / cluster_index of KM/assign_to_centers2.c:500-512.  counts may be NULL */
EVK_API int evk_get_centroids(evk_handle* h, float* c, uint64_t* counts);

/* ---- streaming windows --------------------------------------------------------------------- */
/* Replaces the 50 ms reslicer + on_new_slice callback (ACCEL/store.cpp:329,349-352,370-568):
 * events are appended; whenever window_us of event time is complete the window is downsampled
 * and clustered with centroids warm-started from the previous window.  *windows_done receives
 * the number of windows completed by this call (may be NULL). */
EVK_API int evk_window_config(evk_handle* h, const evk_ds_params* ds, const evk_km_params* km,
                              int64_t window_us);
/* The reslicer's other condition, make_n_events(nevents) (SAMP/store.cpp:336, ACCEL/store.cpp:350):
 * a window is complete after exactly n_events events; its time bins start at its first event. */
EVK_API int evk_window_config_events(evk_handle* h, const evk_ds_params* ds,
                                     const evk_km_params* km, size_t n_events);
EVK_API int evk_window_push(evk_handle* h, const evk_event* begin, const evk_event* end,
                            int* windows_done);
EVK_API int evk_window_flush(evk_handle* h, int* windows_done);

/* ---- consumer: asynchronous event clustering ----------------------------- */
/* The reference hands every unique coordinate a slice keeps to AEClustering::update and draws one
 * centroid + flow arrow per cluster afterwards (ACCEL/store.cpp:435-445,461-521; ACCEL/
 * AEClustering.{h,cpp}, ACCEL/MyCluster.{h,cpp}).  These calls keep that consumer on the device:
 * the same sequential algorithm, state for state (cluster order, ids, moving averages to the last
 * bit, stored events), so the voxels never cross PCIe.  One consumer per handle. */
typedef struct {
    int32_t use_init;  /* 0: `new AEClustering` with init() never called, what the reference app does
                        *    (store.cpp:42; AEClustering.cpp:7-18: szBuffer 800, radius 40, alpha 0.5,
                        *    minN 10, kappa 0 -> the sampling test never fires);
                        * 1: AEClustering::init(sz_buffer, radius, kappa, alpha, min_n), :20-26 */
    int32_t sz_buffer; /* events remembered: older ones are forgotten (AEClustering.cpp:137-146)   */
    double radius;     /* Manhattan radius around a cluster's moving average                      */
    int32_t kappa;     /* stored events sampled per cluster above min_n (MyCluster.cpp:72-103)    */
    int32_t min_n;     /* clusters with fewer events are neither sampled nor reported             */
    double alpha;      /* moving average: mu = (1 - alpha) mu + alpha pix (MyCluster.cpp:200-202) */
    uint32_t rand_seed;/* std::srand value; 1 = a process that never calls srand (the reference)  */
    int32_t max_clusters; /* capacity: simultaneous clusters, <= 1024; 0 = 1024                   */
    int32_t max_points;   /* capacity: stored events per cluster; 0 = 4096                        */
    int32_t _pad;
} evk_aec_params;
typedef struct {
    int32_t id, n;       /* getClusterId(), getN()                                                */
    double mu[2];        /* getMu(): the moving average                                           */
    double centroid[2];  /* getClusterCentroid(): mean of the stored events (NaN when n = 0)      */
} evk_aec_cluster;
typedef struct {         /* one cluster of the per-slice report, ACCEL/store.cpp:466-521          */
    int32_t id, n;
    double centroid[2];  /* cen                                                                   */
    double prev[2];      /* centroid_prev[id] before this report (0,0 = none yet)                 */
    int32_t has_arrow;   /* prev x and y both > 0 (:498): the flow arrow prev -> arrow_end is drawn */
    int32_t _pad;
    double arrow_end[2]; /* prev + (cen - prev), :499-500                                         */
} evk_aec_flow;
/* new AEClustering (+ init); replaces the consumer the handle already has, if any */
EVK_API int evk_aec_create(evk_handle* h, const evk_aec_params* p);
EVK_API int evk_aec_destroy(evk_handle* h);
/* n calls of AEClustering::update (AEClustering.h:37) in order; e = n x {t, x, y, p} doubles in host
 * memory, the std::deque<double> of the reference.  EVK_ERR_CAPACITY when max_clusters or
 * max_points is exceeded (the consumer is then unusable until re-created). */
EVK_API int evk_aec_update(evk_handle* h, const double* e, size_t n);
/* The hand-off loop (ACCEL/store.cpp:435-445) without leaving the device: `count` updates with the
 * representatives of the current voxel shard at canonical positions start, start + step, ...;
 * every event carries pseudo-time t (the reference uses uniqueCount / 1000.0) and polarity 0.
 * As written the reference steps the FLAT x,y array by 4 up to the PAIR count (SURVEY appendix A,
 * D4): that is step = 2, count = ceil(n_unique / 4); step = 1, count = n_unique feeds them all. */
EVK_API int evk_aec_update_voxels(evk_handle* h, double t, size_t start, size_t step, size_t count);
/* eclustering->clusters in list order + getLastUpdatedClusterIdx(); out may be NULL (count only) */
EVK_API int evk_aec_get_clusters(evk_handle* h, evk_aec_cluster* out, size_t cap, size_t* n,
                                 int* last_updated);
/* the stored events of one cluster, oldest first: getDatId / getDat / getDatT / getDatPol */
EVK_API int evk_aec_get_points(evk_handle* h, size_t cluster, int32_t* ids, double* xy, double* t,
                               uint8_t* pol, size_t cap, size_t* n);
/* The per-slice report (ACCEL/store.cpp:461-521): one record per cluster with n >= min_n, list
 * order; centroid_prev[id] then becomes the centroid (16384 ids as in the reference, :188). Call it
 * once per slice: it advances centroid_prev whether or not `out` has room. */
EVK_API int evk_aec_report(evk_handle* h, evk_aec_flow* out, size_t cap, size_t* n);

/* ---- time surface + corner test ----------------------------------------- */
/* The event callback of the reference's corner tracker (event-cam-tracking/
 * event-cam-fast-corner-tracker/metavision_time_surface_periodic_group_track.cpp, FCT): a
 * Metavision::MostRecentTimestampBuffer (:786) stamped by every event (:888-923), then every event
 * of the callback range tested for an Arc*-style corner on the circles of radius 3 and 4 against
 * the updated surface (:931-1057).  One time surface per handle, zero-initialised. */
EVK_API int evk_ts_create(evk_handle* h, int width, int height);
EVK_API int evk_ts_destroy(evk_handle* h);
/* One callback range = the events resident in the handle (evk_load_events ...): stamp, then test.
 * literal_break != 0 reproduces FCT:948-955 as written (the first event within 4 px of the border
 * ends the range's tests); 0 skips that event only.  Timestamps must stay below 2^53 (the reference
 * compares them as doubles). */
EVK_API int evk_ts_corners(evk_handle* h, int literal_break, size_t* n_corners);
/* stream indices (ascending) of the corner events of the last evk_ts_corners = the events pushed to
 * `corners` at FCT:1050 */
EVK_API int evk_ts_get_corners(evk_handle* h, uint32_t* event_index, size_t cap);
/* Corner post-processing: CornerFilter::filterCorners of the same file (FCT:81-151; call site
 * FCT:832 with box_size 15 -- its `threshold` argument and response sort are commented out there).
 * Corners in list order; a corner is kept when its box [x - box_size / 2, x + box_size / 2]^2, clipped
 * to the image, touches no box of a corner kept before it; kept corners are labelled 0, 1, 2, ...
 * At most 32768 corners per call. */
typedef struct evk_corner {
    int32_t x, y, label;
} evk_corner;
/* on the corner list of the last evk_ts_corners (device-resident) */
EVK_API int evk_ts_filter_corners(evk_handle* h, int box_size, size_t* n_kept);
/* on a host list of n (x, y) pairs, as the reference's std::vector<Corner> */
EVK_API int evk_filter_corners(evk_handle* h, const int32_t* xy, size_t n, int width, int height,
                               int box_size, size_t* n_kept);
EVK_API int evk_get_filtered_corners(evk_handle* h, evk_corner* out, size_t cap);
EVK_API int evk_filter_corners_destroy(evk_handle* h);
/* the surface, row-major [height][width] */
EVK_API int evk_ts_get_surface(evk_handle* h, int64_t* out, size_t cap_pixels);

/* ---- density clustering of the downsampled cloud ------------------------- */
/* DBSCAN as the reference runs it on voxel-grid-downsampled clouds (event-cam-clustering/
 * point-cloud-clustering/DBSCAN_simple.h:27-140, parameters as pcl_cluster.cpp:112-120): the same
 * clusters, member for member, as its sequential seed-queue walk -- computed as core flags,
 * connected components of core points (seed = lowest-index core point) and border assignment. */
typedef struct {
    double eps;           /* setClusterTolerance: neighbours are points with distance^2 <= eps^2      */
    int32_t min_pts;      /* setCorePointMinPts: neighbours (the point itself included) of a core point */
    int32_t min_cluster;  /* setMinClusterSize / setMaxClusterSize: clusters outside are dropped      */
    int32_t max_cluster;
    int32_t D;            /* evk_dbscan_voxels only: 2 = (x, y); 3 = (x, y, (t - t0_us) * t_scale)    */
    double t_scale;
    int64_t t0_us;
} evk_dbscan_params;
/* a cloud of n points, xyz = n x 3 floats in host memory (pcl::PointCloud<pcl::PointXYZ>) */
EVK_API int evk_dbscan_points(evk_handle* h, const float* xyz, size_t n, const evk_dbscan_params* p,
                              size_t* n_clusters, size_t* n_extra);
/* the current voxel shard: point i = the representative of the i-th voxel in canonical order */
EVK_API int evk_dbscan_voxels(evk_handle* h, const evk_dbscan_params* p, size_t* n_clusters,
                              size_t* n_extra);
/* Results of the last run.  Clusters are ordered largest first (DBSCAN_simple.h:89; ties: lowest
 * seed first).  labels[i] = position in that order of the first kept cluster that holds point i
 * (lowest seed), -1 = noise or member of dropped clusters only; sizes / seeds per cluster (seed = its lowest-index core point); extra_pairs = (point,
 * cluster) pairs of the second memberships the reference creates when a border point held by an
 * earlier cluster neighbours a later cluster's seed point (:45-50).  Any pointer may be NULL. */
EVK_API int evk_dbscan_get(evk_handle* h, int32_t* labels, size_t cap_points, uint32_t* sizes,
                           uint32_t* seeds, size_t cap_clusters, uint32_t* extra_pairs,
                           size_t cap_extra);
EVK_API int evk_dbscan_destroy(evk_handle* h);

/* OPTICS reachability ordering as the reference computes it on event coordinates (event-cam-
 * clustering/optics-clustering/include/optics/optics.hpp:413-590 compute_reachability_dists; app:
 * test/cluster_event_data.cpp:333-338 with min_pts 2, epsilon 10, threshold 10): neighbours are the
 * points with squared distance <= epsilon^2 (the point itself included), the core distance is the
 * distance to the neighbour of rank min_pts - 1, seeds are popped in (reachability, index) order.
 * Integer coordinates, D = 2 or 3; at most 65536 points, min_pts <= 32.  The neighbourhood work runs
 * on the device, the priority-queue walk on the host. */
EVK_API int evk_optics_points(evk_handle* h, const int32_t* pts, size_t n, int D, int min_pts,
                              double eps);
/* the current voxel shard: point i = (x, y) of the representative of the i-th voxel, canonical order */
EVK_API int evk_optics_voxels(evk_handle* h, int min_pts, double eps);
/* order[k] = index of the k-th point of the ordering, reach[k] = its reachability (-1: none);
 * either may be NULL; *n = number of points */
EVK_API int evk_optics_get(evk_handle* h, uint32_t* order, double* reach, size_t cap, size_t* n);
/* get_cluster_indices(reach_dists, threshold) (optics.hpp:674-690): cluster[k] = id of the cluster of
 * the k-th point of the ordering (a point without reachability or with one >= threshold opens one) */
EVK_API int evk_optics_clusters(evk_handle* h, double threshold, uint32_t* cluster, size_t cap,
                                size_t* n_clusters);
EVK_API int evk_optics_destroy(evk_handle* h);

/* ---- profiling / measurement --------------------------------------------------------------- */
EVK_API int evk_set_profiling(evk_handle* h, int enabled);
EVK_API int evk_get_stage_times(const evk_handle* h, evk_stage_times* out);
/* CUDA-event stopwatch on the handle's stream (bench.py times the HBM-resident step with it) */
EVK_API int evk_timer_start(evk_handle* h);
EVK_API int evk_timer_stop(evk_handle* h, float* elapsed_ms);
EVK_API int evk_sync(evk_handle* h);
/* write `bytes` of device memory (L2 flush between timed iterations) */
EVK_API int evk_flush_l2(evk_handle* h);

/* ---- multi-GPU (one process per GPU, NCCL over NVLink / NVSwitch) --------------------------- */
enum { EVK_OWNER_TIME_RANGE = 0, EVK_OWNER_MIX64 = 1 };
EVK_API int evk_comm_unique_id(uint8_t* id128);                 /* rank 0: ncclGetUniqueId */
EVK_API int evk_comm_init(evk_handle* h, int rank, int world, const uint8_t* id128);
EVK_API int evk_comm_destroy(evk_handle* h);
/* index offset of this rank's shard in the global stream (first_idx becomes global) */
EVK_API int evk_set_shard(evk_handle* h, uint64_t first_global_index);
/* local downsample -> all-to-all of (key, first_idx) by ownership -> owner-side merge.
 * n_unique_local = voxels owned by this rank; n_unique_global = allreduced total. */
EVK_API int evk_downsample_sharded(evk_handle* h, const evk_ds_params* p, int owner_mode,
                                   size_t* n_unique_local, size_t* n_unique_global);
/* k-means on the local voxel shard with an allreduce of the K*(D+1) exact partial sums */
EVK_API int evk_kmeans_sharded(evk_handle* h, const evk_km_params* p, int* iters_done);
/* Fused sharded step: evk_downsample_sharded + (init_first_k != 0) evk_init_centroids_first_k_sharded
 * + evk_kmeans_sharded as one stream-ordered pass with a single host synchronisation: boundary
 * block exchange, downsample on a device-side event range, centroid broadcast beside it, one
 * assign + accumulate pass, ONE allreduce of the K x 5 partial sums together with the voxel
 * counters.  Shapes it does not take run the three calls; same results either way. */
EVK_API int evk_downsample_kmeans_sharded(evk_handle* h, const evk_ds_params* ds,
                                          const evk_km_params* km, int init_first_k,
                                          int owner_mode, size_t* n_unique_local,
                                          size_t* n_unique_global, int* iters_done);
/* The same step in two halves (see evk_downsample_kmeans_submit / _wait): every rank queues its
 * pass, steps may be queued behind each other, wait is the one synchronisation.  Ranks must make
 * the same sequence of calls. */
EVK_API int evk_downsample_kmeans_sharded_submit(evk_handle* h, const evk_ds_params* ds,
                                                 const evk_km_params* km, int init_first_k,
                                                 int owner_mode);
EVK_API int evk_downsample_kmeans_sharded_wait(evk_handle* h, size_t* n_unique_local,
                                               size_t* n_unique_global, int* iters_done);
/* broadcast-free deterministic init: the K globally lowest first indices */
EVK_API int evk_init_centroids_first_k_sharded(evk_handle* h, const evk_km_params* p);

#ifdef __cplusplus
}
#endif
#endif /* EVK_H_ */
