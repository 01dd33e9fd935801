/* optics_oracle.c -- CPU restatement of the reference's OPTICS (TEST INFRASTRUCTURE ONLY).
 *
 * Follows event-cam-clustering/optics-clustering/include/optics/optics.hpp:
 *   :413-590  compute_reachability_dists: neighbours of every point first (the KDTREE method is
 *             compiled in, :410: kdTree.hpp:222 keeps j when square_distance(p_j, p_i) <= radius^2,
 *             the point itself included), then the walk: the lowest unprocessed index starts a
 *             run; a point is emitted, its core distance is computed (:285-298: none with fewer than
 *             min_pts neighbours, else the distance to the neighbour of rank min_pts - 1 by squared
 *             distance), update() (:315-340) gives every unprocessed neighbour the reachability
 *             max(core distance, distance) if that is new or smaller, and the seed with the smallest
 *             (reachability, index) (:66-68, a std::set) is popped next
 *   :674-690  get_cluster_indices(reach, threshold): a point with no reachability (< 0) or one
 *             >= threshold opens a new cluster, every other point joins the current one
 * The app clusters integer (x, y) event coordinates with min_pts 2, epsilon 10, threshold 10
 * (test/cluster_event_data.cpp:333-338).
 * PARITY UNPINNED by the reference: optics.hpp needs boost.geometry, FunctionalPlus ("fplus") and
 * "geometry" (geom::), none of them vendored or present here, so it cannot be compiled; the
 * restatement is pinned by hand-derived known answers and by an independent O(n^2) Python form
 * (tests/test_optics.py).  Distances are IEEE doubles: sqrt of the exact integer squared distance. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double r; uint32_t i; } seed_t;
static int seed_less(seed_t a, seed_t b) { return a.r == b.r ? a.i < b.i : a.r < b.r; }

/* binary heap with lazy deletion: an entry is stale when reach[i] no longer equals its key or the
 * point has been processed; pops come out in the std::set's order */
typedef struct { seed_t* a; size_t n, cap; } heap_t;
static void heap_push(heap_t* h, seed_t s) {
    if (h->n == h->cap) { h->cap = h->cap ? 2 * h->cap : 64; h->a = (seed_t*)realloc(h->a, h->cap * sizeof(seed_t)); }
    size_t k = h->n++;
    while (k && seed_less(s, h->a[(k - 1) / 2])) { h->a[k] = h->a[(k - 1) / 2]; k = (k - 1) / 2; }
    h->a[k] = s;
}
static seed_t heap_pop(heap_t* h) {
    seed_t top = h->a[0], last = h->a[--h->n];
    size_t k = 0;
    for (;;) {
        size_t c = 2 * k + 1;
        if (c >= h->n) break;
        if (c + 1 < h->n && seed_less(h->a[c + 1], h->a[c])) c++;
        if (!seed_less(h->a[c], last)) break;
        h->a[k] = h->a[c];
        k = c;
    }
    if (h->n) h->a[k] = last;
    return top;
}

static int cmp_i64(const void* a, const void* b) {
    const int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
    return x < y ? -1 : x > y;
}

/* pts: n x D int32 (D = 2 or 3).  order[k] = index of the k-th point of the ordering, reach[k] its
 * reachability (-1: none).  Returns 0, or -1 on bad arguments. */
int orc_optics(const int32_t* pts, size_t n, int D, int min_pts, double eps, uint32_t* order,
               double* reach_out) {
    if (D < 2 || D > 3 || min_pts < 1 || !(eps > 0)) return -1;
    const double r2 = eps * eps;
    /* neighbour lists (CSR) */
    size_t* off = (size_t*)calloc(n + 1, sizeof(size_t));
    for (int pass = 0; pass < 2; pass++) {
        static uint32_t* nb;
        if (pass == 1) { nb = (uint32_t*)malloc((off[n] ? off[n] : 1) * sizeof(uint32_t)); }
        size_t run = 0;
        for (size_t i = 0; i < n; i++) {
            size_t c = 0;
            for (size_t j = 0; j < n; j++) {
                int64_t s = 0;
                for (int d = 0; d < D; d++) { const int64_t q = (int64_t)pts[i * D + d] - pts[j * D + d]; s += q * q; }
                if ((double)s <= r2) { if (pass == 1) nb[run + c] = (uint32_t)j; c++; }
            }
            if (pass == 0) off[i + 1] = off[i] + c; else run += c;
        }
        if (pass == 1) {
            double* reach = (double*)malloc((n ? n : 1) * sizeof(double));
            uint8_t* done = (uint8_t*)calloc(n ? n : 1, 1);
            int64_t* tmp = (int64_t*)malloc((n ? n : 1) * sizeof(int64_t));
            for (size_t i = 0; i < n; i++) reach[i] = -1.0;
            heap_t h = {0, 0, 0};
            size_t emitted = 0;
            for (size_t start = 0; start < n; start++) {
                if (done[start]) continue;
                size_t p = start;
                for (;;) {
                    done[p] = 1;
                    order[emitted++] = (uint32_t)p;
                    const size_t m = off[p + 1] - off[p];
                    if (m >= (size_t)min_pts) {
                        for (size_t q = 0; q < m; q++) {
                            const size_t j = nb[off[p] + q];
                            int64_t s = 0;
                            for (int d = 0; d < D; d++) { const int64_t v = (int64_t)pts[p * D + d] - pts[j * D + d]; s += v * v; }
                            tmp[q] = s;
                        }
                        qsort(tmp, m, sizeof(int64_t), cmp_i64);
                        const double core = sqrt((double)tmp[min_pts - 1]);
                        for (size_t q = 0; q < m; q++) {
                            const size_t o = nb[off[p] + q];
                            if (done[o]) continue;
                            int64_t s = 0;
                            for (int d = 0; d < D; d++) { const int64_t v = (int64_t)pts[p * D + d] - pts[o * D + d]; s += v * v; }
                            const double dist = sqrt((double)s), nr = core > dist ? core : dist;
                            if (reach[o] < 0.0 || nr < reach[o]) {
                                reach[o] = nr;
                                const seed_t sd = {nr, (uint32_t)o};
                                heap_push(&h, sd);
                            }
                        }
                    }
                    /* next seed of this run: the smallest live (reachability, index) */
                    int found = 0;
                    while (h.n) {
                        const seed_t s = heap_pop(&h);
                        if (!done[s.i] && reach[s.i] == s.r) { p = s.i; found = 1; break; }
                    }
                    if (!found) break;
                }
            }
            for (size_t k = 0; k < n; k++) reach_out[k] = reach[order[k]];
            free(reach); free(done); free(tmp); free(h.a); free(nb);
        }
    }
    free(off);
    return 0;
}

/* get_cluster_indices(reach_dists, threshold): cluster[k] = id of the cluster the k-th point of the
 * ordering belongs to.  Returns the number of clusters. */
size_t orc_optics_clusters(const double* reach, size_t n, double threshold, uint32_t* cluster) {
    size_t nc = 0;
    for (size_t k = 0; k < n; k++) {
        if (reach[k] < 0.0 || reach[k] >= threshold || nc == 0) nc++;
        cluster[k] = (uint32_t)(nc - 1);
    }
    return nc;
}
