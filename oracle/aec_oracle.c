/* aec_oracle.c -- CPU restatement of the reference's asynchronous event clustering consumer
 * (SURVEY.md 8f rank 1).  TEST INFRASTRUCTURE ONLY: nothing here is linked into, or called by,
 * the product library (libevk.so); only tests/, __graft_entry__.smoke() and bench.py's CPU leg
 * use it, as the checker.
 *
 * Parity is PINNED: the reference's own AEClustering.cpp / MyCluster.cpp compile where they lie
 * (oracle/Makefile target ref_aec, oracle/eigen_shim) and tests/test_aec.py checks this
 * restatement against that library state for state (cluster order, ids, n, mu, stored events).
 *
 * Reference paths, relative to /root/reference/event-cam-clustering-accel/
 * event-cam-clustering-downsampling-accel/ :
 *   AEClustering.cpp:7-26    constructor defaults / init()           -> orc_aec_new
 *   AEClustering.cpp:48-123  update()                                -> aec_update_one
 *   AEClustering.cpp:137-146 updateBuffer_()                         -> (inside aec_update_one)
 *   AEClustering.cpp:148-211 merge_clusters_()                       -> aec_merge
 *   MyCluster.cpp:27-52      add(), :54-65 forget(), :67-70 manhattanDistance(),
 *   MyCluster.cpp:72-103     manhattanDistanceWithSampling() (std::rand, glibc TYPE_3 generator
 *                            restated in orc_glibc_rand), :171-186 getClusterCentroid(),
 *   MyCluster.cpp:200-202    updateMu_()
 *   metavision_sdk_get_started5_opencl_store.cpp:435-445  hand-off loop   -> orc_aec_handoff
 *   metavision_sdk_get_started5_opencl_store.cpp:461-521  per-slice centroid / arrow report
 *                                                                     -> orc_aec_report
 * abs(): the reference calls unqualified abs() on doubles (MyCluster.cpp:61,75,92).  With Eigen's
 * own headers the C++ <stdlib.h> wrapper is in scope, so that is std::abs(double) -- see
 * oracle/eigen_shim/eigen3/Eigen/Core.  fabs() here.
 * All arithmetic is double, one rounding per operation (-ffp-contract=off). */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int id;
    double x, y, t;
    int pol;
} aec_pt;

typedef struct {
    aec_pt* d; /* the four parallel deques datId_/dat_/datT_/datPol_ as one array, front at d[0] */
    int n, cap;
    double mu[2];
    int cluster_id;
} aec_cluster;

typedef struct orc_aec {
    int min_n, sz_buffer, kappa;
    double radius, alpha, t0, t_min;
    int event_id, next_cluster_id, last_updated;
    double* tbuf; /* tBuffer_ */
    int tb_n;
    aec_cluster* c;
    int nc, cc_cap;
    /* glibc random() TYPE_3 state (std::rand, MyCluster.cpp:88) */
    int32_t r[31];
    int rf, rb;
    long n_merges, max_merged, n_draws; /* test-coverage counters */
} orc_aec;

/* glibc srandom_r / random_r, TYPE_3 (x^31 + x^3 + 1): what std::rand() is in the reference build */
static void glibc_srand(orc_aec* a, unsigned seed) {
    int32_t word = seed ? (int32_t)seed : 1;
    a->r[0] = word;
    for (int i = 1; i < 31; i++) {
        const long hi = word / 127773, lo = word % 127773;
        long w = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        word = (int32_t)w;
        a->r[i] = word;
    }
    a->rf = 3;
    a->rb = 0;
    for (int i = 0; i < 310; i++) {
        a->r[a->rf] = (int32_t)((uint32_t)a->r[a->rf] + (uint32_t)a->r[a->rb]);
        a->rf = (a->rf + 1) % 31;
        a->rb = (a->rb + 1) % 31;
    }
}
static int glibc_rand(orc_aec* a) {
    const uint32_t v = (uint32_t)a->r[a->rf] + (uint32_t)a->r[a->rb];
    a->r[a->rf] = (int32_t)v;
    a->rf = (a->rf + 1) % 31;
    a->rb = (a->rb + 1) % 31;
    return (int)(v >> 1);
}
/* exposed so a test can compare the restated generator with this libc's rand() */
void orc_glibc_rand_seq(unsigned seed, int* out, int n) {
    orc_aec a;
    glibc_srand(&a, seed);
    for (int i = 0; i < n; i++) out[i] = glibc_rand(&a);
}

orc_aec* orc_aec_new(int use_init, int sz_buffer, double radius, int kappa, double alpha,
                     int min_n, unsigned rand_seed) {
    orc_aec* a = (orc_aec*)calloc(1, sizeof *a);
    a->min_n = 10; /* AEClustering.cpp:7-18 */
    a->sz_buffer = 800;
    a->t_min = 0;
    a->radius = 40;
    a->alpha = 0.5;
    a->t0 = -1;
    a->kappa = 0;
    a->last_updated = -1;
    if (use_init) { /* AEClustering.cpp:20-26 */
        a->sz_buffer = sz_buffer;
        a->radius = radius;
        a->alpha = alpha;
        a->min_n = min_n;
        a->kappa = kappa;
    }
    a->tbuf = (double*)malloc(sizeof(double) * (size_t)(a->sz_buffer + 2));
    glibc_srand(a, rand_seed);
    return a;
}
void orc_aec_free(orc_aec* a) {
    if (!a) return;
    for (int i = 0; i < a->nc; i++) free(a->c[i].d);
    free(a->c);
    free(a->tbuf);
    free(a);
}

static void cl_push(aec_cluster* k, aec_pt p) {
    if (k->n == k->cap) {
        k->cap = k->cap ? 2 * k->cap : 16;
        k->d = (aec_pt*)realloc(k->d, sizeof(aec_pt) * (size_t)k->cap);
    }
    k->d[k->n++] = p;
}
/* MyCluster::add, MyCluster.cpp:27-52 */
static void cl_add(orc_aec* a, aec_cluster* k, const double* e) {
    aec_pt p;
    p.id = a->event_id;
    p.x = e[1];
    p.y = e[2];
    p.t = e[0] - a->t0;
    p.pol = e[3] != 0.0;
    const int n_before = k->n;
    cl_push(k, p);
    if (n_before == 0) {
        k->mu[0] = p.x;
        k->mu[1] = p.y;
    } else { /* updateMu_, MyCluster.cpp:200-202: (1-alpha)*mu + alpha*pix, coefficient-wise */
        const double om = 1 - a->alpha;
        const double a0 = om * k->mu[0], a1 = om * k->mu[1];
        const double b0 = a->alpha * p.x, b1 = a->alpha * p.y;
        k->mu[0] = a0 + b0;
        k->mu[1] = a1 + b1;
    }
    a->event_id++;
}
/* MyCluster::forget, MyCluster.cpp:54-65 */
static void cl_forget(aec_cluster* k, double t) {
    int pop = 0;
    while (pop < k->n && k->d[pop].t < t) pop++;
    if (pop) {
        memmove(k->d, k->d + pop, sizeof(aec_pt) * (size_t)(k->n - pop));
        k->n -= pop;
    }
}
static double manhattan(double x, double y, double mx, double my) {
    return fabs(x - mx) + fabs(y - my);
}
/* MyCluster::manhattanDistanceWithSampling, MyCluster.cpp:72-103 */
static double cl_sampling(orc_aec* a, const aec_cluster* k, double x, double y) {
    double ma = DBL_MAX;
    if (a->kappa > k->n) {
        for (int i = 0; i < k->n; i++) {
            const double foo = manhattan(x, y, k->d[i].x, k->d[i].y);
            if (foo < ma) ma = foo;
        }
    } else {
        for (int ii = 0; ii < a->kappa; ii++) {
            const int idx = glibc_rand(a) % k->n;
            a->n_draws++;
            const double foo = manhattan(x, y, k->d[idx].x, k->d[idx].y);
            if (foo < ma) ma = foo;
        }
    }
    return ma;
}
static void cl_erase(orc_aec* a, int pos) {
    free(a->c[pos].d);
    memmove(a->c + pos, a->c + pos + 1, sizeof(aec_cluster) * (size_t)(a->nc - pos - 1));
    a->nc--;
}
/* AEClustering::merge_clusters_, AEClustering.cpp:148-211 */
static void aec_merge(orc_aec* a, const int* assigned, int m) {
    int aux_n = 0;
    for (int ii = 0; ii < m; ii++) aux_n += a->c[assigned[ii]].n;
    a->n_merges++;
    if (aux_n > a->max_merged) a->max_merged = aux_n;
    double aux_mu[2] = {0.0, 0.0};
    for (int ii = 0; ii < m; ii++) {
        const aec_cluster* k = &a->c[assigned[ii]];
        const double w = (double)k->n / (double)aux_n;
        const double p0 = w * k->mu[0], p1 = w * k->mu[1];
        aux_mu[0] += p0;
        aux_mu[1] += p1;
    }
    aec_pt* out = (aec_pt*)malloc(sizeof(aec_pt) * (size_t)(aux_n ? aux_n : 1));
    int* count = (int*)calloc((size_t)m, sizeof(int));
    int no = 0, idx = 1;
    while (idx >= 0) { /* repeatedly take the list whose head is oldest; ties: lowest list */
        idx = -1;
        double tt = DBL_MAX;
        for (int jj = 0; jj < m; jj++) {
            const aec_cluster* k = &a->c[assigned[jj]];
            if (count[jj] < k->n && k->d[count[jj]].t < tt) {
                idx = jj;
                tt = k->d[count[jj]].t;
            }
        }
        if (idx >= 0) out[no++] = a->c[assigned[idx]].d[count[idx]++];
    }
    aec_cluster* k0 = &a->c[assigned[0]];
    free(k0->d);
    k0->d = out;
    k0->n = no;
    k0->cap = aux_n ? aux_n : 1;
    k0->mu[0] = aux_mu[0];
    k0->mu[1] = aux_mu[1];
    free(count);
    for (int ii = m - 1; ii > 0; ii--) cl_erase(a, assigned[ii]);
}

/* AEClustering::update, AEClustering.cpp:48-123 */
static void aec_update_one(orc_aec* a, const double* e) {
    if (a->t0 < 0) a->t0 = e[0];
    const double x = e[1], y = e[2];
    const double t = e[0] - a->t0;
    /* updateBuffer_, :137-146 */
    a->tbuf[a->tb_n++] = t;
    if (a->tb_n > a->sz_buffer) {
        memmove(a->tbuf, a->tbuf + 1, sizeof(double) * (size_t)(a->tb_n - 1));
        a->tb_n--;
    }
    a->t_min = a->tbuf[0];
    int* assigned = (int*)malloc(sizeof(int) * (size_t)(a->nc + 1));
    int* removed = (int*)malloc(sizeof(int) * (size_t)(a->nc + 1));
    int na = 0, nr = 0;
    for (int ii = 0; ii < a->nc; ii++) {
        aec_cluster* k = &a->c[ii];
        cl_forget(k, a->t_min);
        if (k->n == 0) removed[nr++] = ii;
        else if (manhattan(x, y, k->mu[0], k->mu[1]) <= a->radius) assigned[na++] = ii;
        else if (k->n > a->min_n) {
            if (cl_sampling(a, k, x, y) <= a->radius) assigned[na++] = ii;
        }
    }
    if (na == 0) {
        if (a->nc == a->cc_cap) {
            a->cc_cap = a->cc_cap ? 2 * a->cc_cap : 16;
            a->c = (aec_cluster*)realloc(a->c, sizeof(aec_cluster) * (size_t)a->cc_cap);
        }
        aec_cluster* k = &a->c[a->nc++];
        memset(k, 0, sizeof *k);
        cl_add(a, k, e);
        k->cluster_id = a->next_cluster_id++;
        a->last_updated = a->nc - 1;
    } else {
        a->last_updated = assigned[0];
        cl_add(a, &a->c[assigned[0]], e);
        if (na >= 2) {
            aec_merge(a, assigned, na);
            free(assigned);
            free(removed);
            return; /* :106-107: the empty clusters found above stay until a later update */
        }
    }
    for (int ii = nr - 1; ii >= 0; ii--) {
        if (a->last_updated > removed[ii]) a->last_updated--;
        cl_erase(a, removed[ii]);
    }
    free(assigned);
    free(removed);
}

void orc_aec_update(orc_aec* a, const double* e, long n) {
    for (long i = 0; i < n; i++) aec_update_one(a, e + 4 * i);
}
int orc_aec_n_clusters(const orc_aec* a) { return a->nc; }
void orc_aec_coverage(const orc_aec* a, long* out3) {
    out3[0] = a->n_merges;
    out3[1] = a->max_merged;
    out3[2] = a->n_draws;
}
int orc_aec_last_updated(const orc_aec* a) { return a->last_updated; }
/* getClusterCentroid, MyCluster.cpp:171-186: sequential sums in deque order, then one division */
static void cl_centroid(const aec_cluster* k, double* cen) {
    double xa = 0, ya = 0;
    for (int i = 0; i < k->n; i++) {
        xa = xa + k->d[i].x;
        ya = ya + k->d[i].y;
    }
    cen[0] = xa / (double)k->n;
    cen[1] = ya / (double)k->n;
}
void orc_aec_get_clusters(const orc_aec* a, int* ids, int* ns, double* mu, double* cen) {
    for (int c = 0; c < a->nc; c++) {
        ids[c] = a->c[c].cluster_id;
        ns[c] = a->c[c].n;
        mu[2 * c] = a->c[c].mu[0];
        mu[2 * c + 1] = a->c[c].mu[1];
        cl_centroid(&a->c[c], cen + 2 * c);
    }
}
int orc_aec_get_points(const orc_aec* a, int c, int* ids, double* xy, double* t, int* pol,
                       int cap) {
    const aec_cluster* k = &a->c[c];
    for (int i = 0; i < k->n && i < cap; i++) {
        ids[i] = k->d[i].id;
        xy[2 * i] = k->d[i].x;
        xy[2 * i + 1] = k->d[i].y;
        t[i] = k->d[i].t;
        pol[i] = k->d[i].pol;
    }
    return k->n;
}

/* The hand-off loop of the slice callback, store.cpp:435-445, as written: the FLAT coordinate
 * array is stepped by 4 and bounded by the PAIR count, every event carries
 * the same pseudo-time uniqueCount / 1000.0 and polarity 0.  Writes n x {t, x, y, p}; returns n. */
long orc_aec_handoff(const int* unique_coords, int unique_count_diff, int unique_count,
                     double* e_out) {
    long n = 0;
    for (int i = 0; i < unique_count_diff; i += 4) {
        e_out[4 * n] = unique_count / 1000.0;
        e_out[4 * n + 1] = unique_coords[i];
        e_out[4 * n + 2] = unique_coords[i + 1];
        e_out[4 * n + 3] = 0;
        n++;
    }
    return n;
}

/* Per-slice report, store.cpp:461-521: for every cluster holding at least minN events, its data
 * centroid, the centroid it had at the previous report (centroid_prev[id], zero-initialised,
 * :188-193) and -- when both previous coordinates are positive -- the flow arrow
 * prev -> prev + (centroid - prev).  centroid_prev[id] then becomes the centroid.
 * rec: n_rec x {id, n, cen_x, cen_y, prev_x, prev_y, has_arrow, end_x, end_y} as doubles.
 * Returns the number of records, or -1 when an id does not fit centroid_prev (the reference
 * indexes double[16384][2] unchecked). */
int orc_aec_report(const orc_aec* a, double* centroid_prev, int max_ids, double* rec, int cap) {
    int n = 0;
    for (int c = 0; c < a->nc; c++) {
        const aec_cluster* k = &a->c[c];
        if (k->n < a->min_n) continue;
        if (k->cluster_id < 0 || k->cluster_id >= max_ids) return -1;
        double cen[2];
        cl_centroid(k, cen);
        double* prev = centroid_prev + 2 * k->cluster_id;
        if (n < cap) {
            double* r = rec + 9 * n;
            r[0] = k->cluster_id;
            r[1] = k->n;
            r[2] = cen[0];
            r[3] = cen[1];
            r[4] = prev[0];
            r[5] = prev[1];
            const double dx = cen[0] - prev[0], dy = cen[1] - prev[1];
            r[6] = (prev[0] > 0 && prev[1] > 0) ? 1.0 : 0.0;
            r[7] = prev[0] + dx;
            r[8] = prev[1] + dy;
        }
        prev[0] = cen[0];
        prev[1] = cen[1];
        n++;
    }
    return n;
}
