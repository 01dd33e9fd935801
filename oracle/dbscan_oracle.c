/* dbscan_oracle.c -- CPU restatement of the reference's DBSCAN.
 * TEST INFRASTRUCTURE ONLY: never linked into or called by the product library.
 * Follows event-cam-clustering/point-cloud-clustering/DBSCAN_simple.h:
 *   :27-95   extract(): points in index order; a point with fewer than minPts neighbours (itself
 *            included, :36-39) is marked noise and skipped; otherwise it seeds a cluster: ALL its
 *            neighbours join the seed queue whatever their state (:45-50), then the queue is walked
 *            (:52-72): noise or processed points are not expanded, core points add their
 *            still-unprocessed neighbours; clusters outside [min, max] points are dropped (:73);
 *            the clusters are returned largest first (:89)
 *   :117-140 radiusSearch(): brute force, double arithmetic on the float coordinates,
 *            distance^2 <= radius^2
 * Parity is PINNED: the header compiles where it lies against oracle/pcl_shim (Makefile target
 * ref_dbscan) and tests/test_dbscan.py compares this restatement with it cluster for cluster.
 * Output: labels[i] = index of i's cluster in the returned order or -1; members of the clusters
 * (sorted) back to back; a point can be a member of a second cluster (see :45-50: the seed's
 * neighbours are taken even when an earlier cluster already holds them) -- labels[] then keeps the
 * first cluster that took it, members[] lists it in both. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int nb_search(const float* p, int n, int i, double r2, int* out) {
    int m = 0;
    out[m++] = i;
    for (int j = 0; j < n; j++) {
        if (j == i) continue;
        const double dx = (double)p[3 * j] - (double)p[3 * i];
        const double dy = (double)p[3 * j + 1] - (double)p[3 * i + 1];
        const double dz = (double)p[3 * j + 2] - (double)p[3 * i + 2];
        if (dx * dx + dy * dy + dz * dz <= r2) out[m++] = j;
    }
    return m;
}
static int cmp_int(const void* a, const void* b) {
    const int x = *(const int*)a, y = *(const int*)b;
    return (x > y) - (x < y);
}
typedef struct {
    int size, start; /* start: offset into members */
    int seed;
} orc_cl;
static int cmp_cl(const void* a, const void* b) { /* largest first; ties: lower seed first */
    const orc_cl *x = (const orc_cl*)a, *y = (const orc_cl*)b;
    if (x->size != y->size) return x->size > y->size ? -1 : 1;
    return (x->seed > y->seed) - (x->seed < y->seed);
}

/* returns the number of clusters; sizes[c], seeds[c] (lowest-index core point = the point that
 * seeded it), members (sorted per cluster, clusters back to back, *n_members in total) */
int orc_dbscan(const float* xyz, int n, double eps, int min_pts, int min_cluster, int max_cluster,
               int* labels, int* sizes, int* seeds, int cap_clusters, int* members,
               long cap_members, long* n_members) {
    enum { UN = 0, PROCESSING = 1, PROCESSED = 2 };
    const double r2 = eps * eps;
    int* nn = (int*)malloc(sizeof(int) * (size_t)(n + 1));
    int* queue = (int*)malloc(sizeof(int) * (size_t)(2 * n + 2));
    unsigned char* types = (unsigned char*)calloc((size_t)n + 1, 1);
    unsigned char* noise = (unsigned char*)calloc((size_t)n + 1, 1);
    orc_cl* cl = (orc_cl*)malloc(sizeof(orc_cl) * (size_t)(n + 1));
    int* all = (int*)malloc(sizeof(int) * (size_t)(2 * n + 2));
    int nc = 0;
    long total = 0;
    for (int i = 0; i < n; i++) {
        if (types[i] == PROCESSED) continue;
        int m = nb_search(xyz, n, i, r2, nn);
        if (m < min_pts) {
            noise[i] = 1;
            continue;
        }
        int q = 0;
        queue[q++] = i;
        types[i] = PROCESSED;
        for (int j = 0; j < m; j++)
            if (nn[j] != i) {
                queue[q++] = nn[j];
                types[nn[j]] = PROCESSING;
            }
        for (int s = 1; s < q; s++) {
            const int c = queue[s];
            if (noise[c] || types[c] == PROCESSED) {
                types[c] = PROCESSED;
                continue;
            }
            m = nb_search(xyz, n, c, r2, nn);
            if (m >= min_pts)
                for (int j = 0; j < m; j++)
                    if (types[nn[j]] == UN) {
                        queue[q++] = nn[j];
                        types[nn[j]] = PROCESSING;
                    }
            types[c] = PROCESSED;
        }
        if (q >= min_cluster && q <= max_cluster) {
            qsort(queue, (size_t)q, sizeof(int), cmp_int);
            int u = 0;
            for (int k = 0; k < q; k++)
                if (k == 0 || queue[k] != queue[k - 1]) queue[u++] = queue[k];
            if (total + u > 2L * n + 2) break; /* cannot happen */
            memcpy(all + total, queue, sizeof(int) * (size_t)u);
            cl[nc].size = u;
            cl[nc].start = (int)total;
            cl[nc].seed = i;
            nc++;
            total += u;
        }
    }
    qsort(cl, (size_t)nc, sizeof(orc_cl), cmp_cl);
    for (int i = 0; i < n; i++) labels[i] = -1;
    long o = 0;
    for (int c = 0; c < nc; c++) {
        if (c < cap_clusters) {
            sizes[c] = cl[c].size;
            seeds[c] = cl[c].seed;
        }
        for (int k = 0; k < cl[c].size; k++) {
            const int v = all[cl[c].start + k];
            if (members && o < cap_members) members[o] = v;
            o++;
        }
    }
    /* labels: the cluster that took the point first = the one with the lowest seed among its clusters */
    {
        int* best_seed = (int*)malloc(sizeof(int) * (size_t)(n + 1));
        for (int i = 0; i < n; i++) best_seed[i] = 0x7fffffff;
        for (int c = 0; c < nc; c++)
            for (int k = 0; k < cl[c].size; k++) {
                const int v = all[cl[c].start + k];
                if (cl[c].seed < best_seed[v]) {
                    best_seed[v] = cl[c].seed;
                    labels[v] = c;
                }
            }
        free(best_seed);
    }
    if (n_members) *n_members = o;
    free(nn);
    free(queue);
    free(types);
    free(noise);
    free(cl);
    free(all);
    return nc;
}
