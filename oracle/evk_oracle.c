/* evk_oracle.c — CPU ORACLE (test infrastructure only; see evk_oracle.h for the rules).
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fopenmp -shared -fPIC (oracle/Makefile).
 * -ffp-contract=off matters: the contract arithmetic is "dx*dx rounded, then ONE fmaf", and gcc
 * must not fuse or split anything on its own.
 */
#include "evk_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/evk_synth.h"

/* ============================================================================================
 * 1. Literal restatements
 * ========================================================================================== */

/* coordinate_processor.cl:3-14 — hash_coordinate(); the dead locals width/height are omitted */
static int ref_hash_coordinate(int x, int y) { return (x * 1619 + y * 31) % 8192; }

int orc_ref_process_coordinates(const int* input_coords, int total_coords, int* unique_coords,
                                int* repeated_count, int* unique_count) {
    /* :29-44 — __local int coordinate_map[8192] zeroed, local counters zeroed */
    static int coordinate_map[8192];
    int local_repeated_count = 0, local_unique_count = 0;
    memset(coordinate_map, 0, sizeof coordinate_map);
    /* :50-78 — the grid-stride loop, run for one work-item after the other (i ascending) */
    for (int i = 0; i < total_coords; i++) {
        int x = input_coords[i * 2];     /* :52 */
        int y = input_coords[i * 2 + 1]; /* :53 */
        if (x >= 0 && x <= 1280 && y >= 0 && y <= 720) { /* :56, inclusive as written */
            int hash = ref_hash_coordinate(x, y);         /* :58 */
            int prev_value = coordinate_map[hash]++;      /* :62 atomic_inc returns old */
            if (prev_value == 0) {                        /* :65 first occurrence */
                int index = local_unique_count++;         /* :66 */
                unique_coords[index * 2] = x;             /* :68 */
                unique_coords[index * 2 + 1] = y;         /* :69 */
            } else if (prev_value == 1) {                 /* :73 second occurrence */
                local_repeated_count++;                   /* :74 */
            }
        }
    }
    *repeated_count += local_repeated_count; /* :85 atomic_add to a never-reset counter */
    *unique_count += local_unique_count;     /* :86 */
    return local_unique_count;
}

int orc_ref_analyze_coordinates(const int* data, int n_ints, int* xs, int* ys, int* counts) {
    /* FCT/metavision_time_surface_periodic.cpp:72-98 with findCoordinate (:57-67) inlined */
    int uniqueCount = 0;
    for (int i = 0; i + 1 < n_ints; i += 2) {
        int x = data[i], y = data[i + 1];
        int existingIndex = -1;
        for (int j = 0; j < uniqueCount; j++)
            if (xs[j] == x && ys[j] == y) {
                existingIndex = j;
                break;
            }
        if (existingIndex != -1) {
            counts[existingIndex]++;
        } else {
            xs[uniqueCount] = x;
            ys[uniqueCount] = y;
            counts[uniqueCount] = 1;
            uniqueCount++;
        }
    }
    return uniqueCount;
}

float orc_ref_kmeans_trip(const float* data, float* centroids, float* output, int* assign,
                          int* cluster_index, float* scalar_sum, float* new_centroids) {
    enum { ARRAY_SIZE = 4096, P = ARRAY_SIZE / 2, CLUSTER_NUM = 8 };
    /* assign_to_centers2.c:186-188 */
    for (int i = 0; i < 8; i++) cluster_index[i] = 0;

    /* K1 assign_to_centers, assign_to_centers.cl:1-34, global size 2048 */
    for (int g = 0; g < P; g++) {
        unsigned gx = (unsigned)g * 2;
        float data_x = data[gx], data_y = data[gx + 1];
        float threshold_dd = 50.0f;
        unsigned char indMin = (unsigned char)-1; /* :12 */
        for (int i = 0; i < 16; i += 2) {
            float localDx = centroids[i] - data_x;
            float localDy = centroids[i + 1] - data_y;
            /* length((float3)(dx,dy,0)) :17-18 */
            float localD = sqrtf(localDx * localDx + localDy * localDy + 0.0f);
            if (localD < threshold_dd) {
                indMin = (unsigned char)i;
                threshold_dd = localD;
            }
            assign[gx / 2] = indMin; /* :26 */
        }
    }
    /* K2 assign_data_cluster, assign_to_centers.cl:36-119, work-items in gid order */
    for (int g = 0; g < P; g++) {
        unsigned gid = (unsigned)g * 2;
        unsigned assign_cluster = (unsigned)assign[gid / 2] / 2; /* :43 */
        unsigned address_offset = assign_cluster * 4096;          /* :45 */
        unsigned address_offset_y = address_offset + 2048;        /* :46 */
        if (assign_cluster < 8) {                                 /* the if-chain :48-111 */
            int index = cluster_index[assign_cluster]++;          /* atomic_fetch_add */
            output[address_offset + index] = data[gid];
            output[address_offset_y + index] = data[gid + 1];
        }
    }
    /* K3 reduction_scalar, assign_to_centers.cl:121-140: 32 groups of 1024, fp32 tree */
    for (int grp = 0; grp < 32; grp++) {
        float partial[1024];
        for (int l = 0; l < 1024; l++) partial[l] = output[grp * 1024 + l];
        for (int i = 512; i > 0; i >>= 1)
            for (int l = 0; l < i; l++) partial[l] += partial[l + i];
        scalar_sum[grp] = partial[0];
    }
    /* host: assign_to_centers2.c:500-512 (stride-2 indexing, as written) */
    unsigned y_offset = 2;
    for (int j = 0; j < CLUSTER_NUM * 2; j += 2) {
        new_centroids[j] = (scalar_sum[j] + scalar_sum[j + 1]) / cluster_index[j / 2];
        new_centroids[j + 1] =
            (scalar_sum[j + y_offset] + scalar_sum[j + 1 + y_offset]) / cluster_index[j / 2];
    }
    /* :514-532 */
    float error[16];
    float error_max = 0.0f;
    for (int j = 0; j < CLUSTER_NUM * 2; j += 2) {
        error[j] = new_centroids[j] - centroids[j];
        error[j + 1] = new_centroids[j + 1] - centroids[j + 1];
    }
    for (int j = 0; j < CLUSTER_NUM * 2; j++) {
        if (abs((int)error[j]) > error_max) { /* C int abs() on a float, :526 */
            error_max = (float)abs((int)error[j]);
            centroids[j] = new_centroids[j];
        }
    }
    return error_max;
}

/* ============================================================================================
 * 2. Contract semantics
 * ========================================================================================== */

int orc_event_key(const evk_event* e, const evk_ds_params* p, uint64_t* key) {
    if (p->keyfn == EVK_KEY_REF_HASH8192) {
        /* coordinate_processor.cl:56 gate (inclusive) and :12 hash */
        int x = e->x, y = e->y;
        if (!(x >= 0 && x <= p->width && y >= 0 && y <= p->height)) return 0;
        *key = (uint64_t)((x * 1619 + y * 31) % 8192);
        return 1;
    }
    /* VOXEL(vx,vy,vt,use_p): SURVEY.md 8a */
    if ((int)e->x >= p->width || (int)e->y >= p->height) return 0;
    if (e->t < p->t0_us) return 0;
    uint64_t NX = (uint64_t)((p->width + p->vx - 1) / p->vx);
    uint64_t NY = (uint64_t)((p->height + p->vy - 1) / p->vy);
    uint64_t xbin = (uint64_t)(e->x / p->vx), ybin = (uint64_t)(e->y / p->vy);
    uint64_t tbin = p->vt_us > 0 ? (uint64_t)((e->t - p->t0_us) / p->vt_us) : 0;
    uint64_t k = (tbin * NY + ybin) * NX + xbin;
    if (p->use_polarity) k = k * 2 + (e->p > 0 ? 1u : 0u);
    *key = k;
    return 1;
}

static inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 33)) * 0xFF51AFD7ED558CCDull;
    z = (z ^ (z >> 33)) * 0xC4CEB9FE1A85EC53ull;
    return z ^ (z >> 33);
}

typedef struct {
    uint64_t* keys;  /* EMPTY = ~0 */
    uint32_t* first;
    uint8_t* rep;
    size_t mask;
} orc_map;

static int map_init(orc_map* m, size_t n) {
    size_t cap = 16;
    while (cap < 2 * n + 2) cap <<= 1;
    m->keys = (uint64_t*)malloc(cap * sizeof(uint64_t));
    m->first = (uint32_t*)malloc(cap * sizeof(uint32_t));
    m->rep = (uint8_t*)calloc(cap, 1);
    if (!m->keys || !m->first || !m->rep) return -1;
    memset(m->keys, 0xFF, cap * sizeof(uint64_t));
    m->mask = cap - 1;
    return 0;
}
static void map_free(orc_map* m) {
    free(m->keys);
    free(m->first);
    free(m->rep);
}
/* returns slot; *fresh = 1 when the key was inserted by this call */
static inline size_t map_put(orc_map* m, uint64_t key, int* fresh) {
    size_t s = (size_t)mix64(key) & m->mask;
    for (;;) {
        if (m->keys[s] == key) {
            *fresh = 0;
            return s;
        }
        if (m->keys[s] == ~0ull) {
            m->keys[s] = key;
            *fresh = 1;
            return s;
        }
        s = (s + 1) & m->mask;
    }
}

size_t orc_downsample(const evk_event* ev, size_t n, const evk_ds_params* p, uint64_t* keys,
                      uint32_t* first_idx, size_t* n_repeated) {
    orc_map m;
    if (map_init(&m, n)) return 0;
    size_t U = 0, R = 0;
    /* the reference kernel's loop body (coordinate_processor.cl:50-78) for i ascending: the
     * first event that finds its bucket empty is emitted, in arrival order */
    for (size_t i = 0; i < n; i++) {
        uint64_t k;
        if (!orc_event_key(&ev[i], p, &k)) continue;
        int fresh;
        size_t s = map_put(&m, k, &fresh);
        if (fresh) {
            m.first[s] = (uint32_t)i;
            keys[U] = k;
            first_idx[U] = (uint32_t)i;
            U++;
        } else if (!m.rep[s]) { /* prev_value == 1 branch, :73-75 */
            m.rep[s] = 1;
            R++;
        }
    }
    map_free(&m);
    if (n_repeated) *n_repeated = R;
    return U;
}

typedef struct {
    uint64_t key;
    uint32_t first;
    uint32_t rep;
} orc_rec;

static int cmp_first(const void* a, const void* b) {
    uint32_t x = ((const orc_rec*)a)->first, y = ((const orc_rec*)b)->first;
    return x < y ? -1 : (x > y ? 1 : 0);
}

size_t orc_downsample_mt(const evk_event* ev, size_t n, const evk_ds_params* p, int threads,
                         int canonical, uint64_t* keys, uint32_t* first_idx,
                         size_t* n_repeated) {
    if (threads < 1) threads = 1;
    const int T = threads;
    orc_rec** lists = (orc_rec**)calloc((size_t)T, sizeof(orc_rec*));
    size_t* lens = (size_t*)calloc((size_t)T, sizeof(size_t));
    orc_rec** outs = (orc_rec**)calloc((size_t)T, sizeof(orc_rec*));
    size_t* olens = (size_t*)calloc((size_t)T, sizeof(size_t));
    /* phase 1: every thread de-duplicates its contiguous index range */
#pragma omp parallel num_threads(T)
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num();
#else
        int tid = 0;
#endif
        size_t lo = n * (size_t)tid / (size_t)T, hi = n * (size_t)(tid + 1) / (size_t)T;
        orc_map m;
        map_init(&m, hi - lo);
        orc_rec* L = (orc_rec*)malloc((hi - lo + 1) * sizeof(orc_rec));
        size_t u = 0;
        for (size_t i = lo; i < hi; i++) {
            uint64_t k;
            if (!orc_event_key(&ev[i], p, &k)) continue;
            int fresh;
            size_t s = map_put(&m, k, &fresh);
            if (fresh) {
                m.first[s] = (uint32_t)u; /* position in L */
                L[u].key = k;
                L[u].first = (uint32_t)i;
                L[u].rep = 0;
                u++;
            } else {
                L[m.first[s]].rep = 1;
            }
        }
        map_free(&m);
        lists[tid] = L;
        lens[tid] = u;
    }
    size_t total = 0;
    for (int t = 0; t < T; t++) total += lens[t];
    /* phase 2: thread q merges the keys it owns (mix64(key) % T == q) from all ranges, in range
     * order, so the lowest stream index wins */
#pragma omp parallel num_threads(T)
    {
#ifdef _OPENMP
        int q = omp_get_thread_num();
#else
        int q = 0;
#endif
        size_t mine = 0;
        for (int t = 0; t < T; t++)
            for (size_t j = 0; j < lens[t]; j++)
                if ((int)(mix64(lists[t][j].key ^ 0x5bd1e995u) % (uint64_t)T) == q) mine++;
        orc_map m;
        map_init(&m, mine);
        orc_rec* O = (orc_rec*)malloc((mine + 1) * sizeof(orc_rec));
        size_t u = 0;
        for (int t = 0; t < T; t++)
            for (size_t j = 0; j < lens[t]; j++) {
                const orc_rec* r = &lists[t][j];
                if ((int)(mix64(r->key ^ 0x5bd1e995u) % (uint64_t)T) != q) continue;
                int fresh;
                size_t s = map_put(&m, r->key, &fresh);
                if (fresh) {
                    m.first[s] = (uint32_t)u;
                    O[u] = *r;
                    u++;
                } else {
                    O[m.first[s]].rep = 1;
                }
            }
        map_free(&m);
        outs[q] = O;
        olens[q] = u;
    }
    (void)total;
    size_t U = 0, R = 0;
    for (int q = 0; q < T; q++) U += olens[q];
    orc_rec* all = (orc_rec*)malloc((U + 1) * sizeof(orc_rec));
    size_t w = 0;
    for (int q = 0; q < T; q++) {
        memcpy(all + w, outs[q], olens[q] * sizeof(orc_rec));
        w += olens[q];
    }
    if (canonical) qsort(all, U, sizeof(orc_rec), cmp_first);
    for (size_t i = 0; i < U; i++) {
        keys[i] = all[i].key;
        first_idx[i] = all[i].first;
        R += all[i].rep;
    }
    if (n_repeated) *n_repeated = R;
    for (int t = 0; t < T; t++) {
        free(lists[t]);
        free(outs[t]);
    }
    free(all);
    free(lists);
    free(lens);
    free(outs);
    free(olens);
    return U;
}

void orc_points(const evk_event* ev, const uint32_t* first_idx, size_t U, int D, int64_t t0_us,
                float t_scale, float p_scale, float* pts) {
    for (size_t i = 0; i < U; i++) {
        const evk_event* e = &ev[first_idx ? first_idx[i] : i];
        float* q = pts + i * (size_t)D;
        q[0] = (float)e->x;
        q[1] = (float)e->y;
        if (D > 2) q[2] = (float)(e->t - t0_us) * t_scale;
        if (D > 3) q[3] = (e->p > 0 ? 1.0f : 0.0f) * p_scale;
    }
}

static inline int32_t assign_one(const float* q, int D, const float* cent, int K, float best,
                                 int use_sqrt) {
    int32_t lab = -1;
    for (int k = 0; k < K; k++) {
        const float* c = cent + (size_t)k * (size_t)D;
        float dx = c[0] - q[0]; /* operand order of assign_to_centers.cl:15-16 */
        float dy = c[1] - q[1];
        float d = fmaf(dy, dy, dx * dx);
        for (int j = 2; j < D; j++) {
            float dj = c[j] - q[j];
            d = fmaf(dj, dj, d);
        }
        if (use_sqrt) d = sqrtf(d); /* length(), :17-18 */
        if (d < best) {             /* strict '<' :21 => lowest k wins ties */
            lab = k;
            best = d;
        }
    }
    return lab;
}

static float gate_value(float max_dist, int use_sqrt) {
    if (!(max_dist > 0.0f) || isinf(max_dist)) return INFINITY;
    return use_sqrt ? max_dist : max_dist * max_dist;
}

void orc_kmeans_assign(const float* pts, size_t P, int D, const float* cent, int K,
                       float max_dist, int use_sqrt, int32_t* labels) {
    float best = gate_value(max_dist, use_sqrt);
    for (size_t i = 0; i < P; i++)
        labels[i] = assign_one(pts + i * (size_t)D, D, cent, K, best, use_sqrt);
}

static float finalise(int K, int D, const double* sums, const uint64_t* counts, float* cent) {
    float shift = 0.0f;
    for (int k = 0; k < K; k++) {
        if (counts[k] == 0) continue; /* empty cluster keeps its centroid (contract, D15) */
        for (int j = 0; j < D; j++) {
            float nc = (float)(sums[(size_t)k * D + j] / (double)counts[k]);
            float d = fabsf(nc - cent[(size_t)k * D + j]);
            if (d > shift) shift = d;
            cent[(size_t)k * D + j] = nc;
        }
    }
    return shift;
}

float orc_kmeans_update(const float* pts, size_t P, int D, const int32_t* labels, int K,
                        float* cent, uint64_t* counts, double* sums) {
    double* s = (double*)calloc((size_t)K * D, sizeof(double));
    uint64_t* c = (uint64_t*)calloc((size_t)K, sizeof(uint64_t));
    for (size_t i = 0; i < P; i++) {
        int32_t l = labels[i];
        if (l < 0) continue; /* unassigned points contribute nothing (a8 skip) */
        c[l]++;
        for (int j = 0; j < D; j++) s[(size_t)l * D + j] += (double)pts[i * (size_t)D + j];
    }
    float shift = finalise(K, D, s, c, cent);
    if (counts) memcpy(counts, c, (size_t)K * sizeof(uint64_t));
    if (sums) memcpy(sums, s, (size_t)K * D * sizeof(double));
    free(s);
    free(c);
    return shift;
}

int orc_kmeans(const float* pts, size_t P, int D, float* cent, int K, float max_dist, int iters,
               float tol, int32_t* labels, uint64_t* counts) {
    int it = 0;
    while (it < iters) {
        orc_kmeans_assign(pts, P, D, cent, K, max_dist, 0, labels);
        float shift = orc_kmeans_update(pts, P, D, labels, K, cent, counts, NULL);
        it++;
        if (tol >= 0.0f && shift <= tol) break;
    }
    return it;
}

int orc_kmeans_mt(const float* pts, size_t P, int D, float* cent, int K, float max_dist,
                  int iters, float tol, int threads, int32_t* labels, uint64_t* counts) {
    if (threads < 1) threads = 1;
    const int T = threads;
    float best = gate_value(max_dist, 0);
    double* S = (double*)malloc((size_t)T * K * D * sizeof(double));
    uint64_t* C = (uint64_t*)malloc((size_t)T * K * sizeof(uint64_t));
    double* s = (double*)malloc((size_t)K * D * sizeof(double));
    uint64_t* c = (uint64_t*)malloc((size_t)K * sizeof(uint64_t));
    int it = 0;
    while (it < iters) {
        memset(S, 0, (size_t)T * K * D * sizeof(double));
        memset(C, 0, (size_t)T * K * sizeof(uint64_t));
#pragma omp parallel num_threads(T)
        {
#ifdef _OPENMP
            int tid = omp_get_thread_num();
#else
            int tid = 0;
#endif
            size_t lo = P * (size_t)tid / (size_t)T, hi = P * (size_t)(tid + 1) / (size_t)T;
            double* ms = S + (size_t)tid * K * D;
            uint64_t* mc = C + (size_t)tid * K;
            for (size_t i = lo; i < hi; i++) {
                int32_t l = assign_one(pts + i * (size_t)D, D, cent, K, best, 0);
                labels[i] = l;
                if (l < 0) continue;
                mc[l]++;
                for (int j = 0; j < D; j++) ms[(size_t)l * D + j] += (double)pts[i * (size_t)D + j];
            }
        }
        memset(s, 0, (size_t)K * D * sizeof(double));
        memset(c, 0, (size_t)K * sizeof(uint64_t));
        for (int t = 0; t < T; t++) {
            for (int k = 0; k < K * D; k++) s[k] += S[(size_t)t * K * D + k];
            for (int k = 0; k < K; k++) c[k] += C[(size_t)t * K + k];
        }
        float shift = finalise(K, D, s, c, cent);
        it++;
        if (tol >= 0.0f && shift <= tol) break;
    }
    if (counts) memcpy(counts, c, (size_t)K * sizeof(uint64_t));
    free(S);
    free(C);
    free(s);
    free(c);
    return it;
}

/* ============================================================================================
 * 3. Inputs
 * ========================================================================================== */

void orc_synth(const evk_synth_params* sp, evk_event* out) {
    for (uint64_t j = 0; j < sp->n_events; j++) out[j] = evk_synth_event(sp, sp->first_index + j);
}

void orc_synth_mt(const evk_synth_params* sp, evk_event* out, int threads) {
    if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int64_t j = 0; j < (int64_t)sp->n_events; j++)
        out[j] = evk_synth_event(sp, sp->first_index + (uint64_t)j);
}

long orc_load_csv(const char* path, evk_event* out, size_t cap) {
    /* same row shape as optics-clustering/test/cluster_event_data.cpp:21-55 reads: x,y,t,p */
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    char line[256];
    size_t n = 0;
    while (n < cap && fgets(line, sizeof line, f)) {
        long x, y, p;
        long long t;
        if (sscanf(line, "%ld,%ld,%lld,%ld", &x, &y, &t, &p) != 4) continue;
        out[n].x = (uint16_t)x;
        out[n].y = (uint16_t)y;
        out[n].p = (int16_t)p;
        out[n]._pad = 0;
        out[n].t = (int64_t)t;
        n++;
    }
    fclose(f);
    return (long)n;
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- RAW EVT 2.0 (see evk_oracle.h) ----------------------------------------------------------- */
size_t orc_evt2_decode(const uint32_t* words, size_t n_words, evk_event* out, size_t cap) {
    uint64_t time_high = 0;
    size_t n = 0;
    for (size_t i = 0; i < n_words; i++) {
        const uint32_t w = words[i];
        const uint32_t type = w >> 28;
        if (type == 0x8u) {
            time_high = w & 0x0FFFFFFFu;
        } else if (type <= 0x1u) {
            if (n >= cap) break;
            evk_event e;
            e.x = (uint16_t)((w >> 11) & 0x7FFu);
            e.y = (uint16_t)(w & 0x7FFu);
            e.p = (int16_t)type;
            e._pad = 0;
            e.t = (int64_t)((time_high << 6) | ((w >> 22) & 0x3Fu));
            out[n++] = e;
        }
    }
    return n;
}

size_t orc_evt2_encode(const evk_event* ev, size_t n, uint32_t* words, size_t cap) {
    size_t m = 0;
    uint64_t time_high = ~0ull;
    for (size_t i = 0; i < n; i++) {
        const evk_event* e = &ev[i];
        if (e->x >= 2048 || e->y >= 2048 || e->t < 0 || e->t >= (1ll << 34)) return (size_t)-1;
        const uint64_t th = (uint64_t)e->t >> 6;
        if (th != time_high) {
            if (m >= cap) return m;
            words[m++] = 0x80000000u | (uint32_t)(th & 0x0FFFFFFFu);
            time_high = th;
        }
        if (m >= cap) return m;
        words[m++] = ((e->p > 0 ? 1u : 0u) << 28) | (((uint32_t)e->t & 0x3Fu) << 22) |
                     ((uint32_t)e->x << 11) | (uint32_t)e->y;
    }
    return m;
}

/* ---- RAW EVT 3.0 (third-party format: Prophesee "EVT 3.0"; the Metavision SDK that decodes it
 * for the reference -- Camera::from_file, ACCEL/store.cpp:336 -- is absent from /root/reference
 * and its version is unpinned, so the published word layout is restated here; PARITY UNPINNED by
 * the reference: pinned only by known-answer words and encode/decode round trips).
 * 16-bit words, type = bits 15..12: 0x0 EVT_ADDR_Y [10:0] y; 0x2 EVT_ADDR_X [11] p [10:0] x (one
 * event); 0x3 VECT_BASE_X [11] p [10:0] x; 0x4 VECT_12 [11:0] mask, 0x5 VECT_8 [7:0] mask (an event
 * at base + i per set bit, then base += 12 / 8); 0x6 EVT_TIME_LOW t[11:0]; 0x8 EVT_TIME_HIGH
 * t[23:12], a value lower than the previous EVT_TIME_HIGH counts one 2^24 us wrap; every other type
 * carries no CD event.  State before the first word of each kind: 0. */
size_t orc_evt3_decode(const uint16_t* words, size_t n_words, evk_event* out, size_t cap) {
    uint32_t th = 0, tl = 0, y = 0, base = 0, pol = 0;
    uint64_t wraps = 0;
    int have_th = 0;
    size_t n = 0;
    for (size_t i = 0; i < n_words; i++) {
        const uint32_t w = words[i], type = w >> 12, v = w & 0xFFFu;
        if (type == 0x0) y = v & 0x7FFu;
        else if (type == 0x6) tl = v;
        else if (type == 0x8) {
            if (have_th && v < th) wraps++;
            th = v;
            have_th = 1;
        } else if (type == 0x3) {
            base = v & 0x7FFu;
            pol = (v >> 11) & 1u;
        } else if (type == 0x2 || type == 0x4 || type == 0x5) {
            const int64_t t = (int64_t)((wraps << 24) | ((uint64_t)th << 12) | tl);
            if (type == 0x2) {
                if (n < cap) {
                    evk_event e = {(uint16_t)(v & 0x7FFu), (uint16_t)y, (int16_t)((v >> 11) & 1u), 0, t};
                    out[n] = e;
                }
                n++;
            } else {
                const uint32_t bits = type == 0x4 ? 12u : 8u;
                for (uint32_t b = 0; b < bits; b++)
                    if (v & (1u << b)) {
                        if (n < cap) {
                            evk_event e = {(uint16_t)((base + b) & 0x7FFFu), (uint16_t)y, (int16_t)pol, 0, t};
                            out[n] = e;
                        }
                        n++;
                    }
                base = (base + bits) & 0x7FFFu;
            }
        }
    }
    return n; /* may exceed cap: the number of CD events in the stream */
}

/* A writer that uses every word kind: time words only when they change, a row word when the row
 * changes, runs of same-time same-row same-polarity events with increasing columns as vectors
 * (VECT_8 when the run spans < 8 columns, else VECT_12; a vector that continues exactly where the
 * previous one stopped reuses the running base without a new VECT_BASE_X), single events as
 * EVT_ADDR_X.  Events must be time-ordered with gaps < 2^24 us and start below 2^24 us (the wrap
 * count is implicit in the format).  Returns the number of words, or (size_t)-1. */
size_t orc_evt3_encode(const evk_event* ev, size_t n, uint16_t* words, size_t cap) {
    size_t m = 0;
    int64_t cur_hi = -1, cur_tl = -1, cur_y = -1, prev_t = 0;
    int64_t vec_next = -1; /* running vector base, valid while t, y, p stay the same */
    int vec_pol = 0;
#define PUT(w)                        \
    do {                              \
        if (m >= cap) return m;       \
        words[m++] = (uint16_t)(w);   \
    } while (0)
    for (size_t i = 0; i < n;) {
        const evk_event* e = &ev[i];
        const int pol = e->p > 0;
        if (e->x >= 2048 || e->y >= 2048 || e->t < 0) return (size_t)-1;
        if (i == 0 ? e->t >= (1ll << 24) : (e->t < prev_t || e->t - prev_t >= (1ll << 24) - 4096))
            return (size_t)-1;
        prev_t = e->t;
        if ((e->t >> 12) != cur_hi) {
            cur_hi = e->t >> 12;
            PUT(0x8000u | (uint32_t)(cur_hi & 0xFFF));
            vec_next = -1;
        }
        if ((e->t & 0xFFF) != cur_tl) {
            cur_tl = e->t & 0xFFF;
            PUT(0x6000u | (uint32_t)cur_tl);
            vec_next = -1;
        }
        if (e->y != cur_y) {
            cur_y = e->y;
            PUT(0x0000u | (uint32_t)cur_y);
            vec_next = -1;
        }
        /* the run of events that fit one vector word starting at column x0 */
        int64_t x0 = e->x;
        const int cont = vec_next >= 0 && vec_pol == pol && e->x >= vec_next && e->x - vec_next < 12;
        if (cont) x0 = vec_next;
        size_t j = i + 1;
        while (j < n && ev[j].t == e->t && ev[j].y == e->y && (ev[j].p > 0) == pol &&
               ev[j].x > ev[j - 1].x && ev[j].x - x0 < 12)
            j++;
        if (j - i >= 2 || cont) {
            uint32_t mask = 0;
            for (size_t k = i; k < j; k++) mask |= 1u << (ev[k].x - x0);
            if (!cont) PUT(0x3000u | ((uint32_t)pol << 11) | (uint32_t)x0);
            if (mask < 256u && !cont) {
                PUT(0x5000u | mask);
                vec_next = x0 + 8;
            } else {
                PUT(0x4000u | mask);
                vec_next = x0 + 12;
            }
            vec_pol = pol;
            i = j;
        } else {
            PUT(0x2000u | ((uint32_t)pol << 11) | (uint32_t)e->x);
            i++; /* (an EVT_ADDR_X word leaves the vector base alone) */
        }
    }
#undef PUT
    return m;
}

/* ---- time surface + corner test --------------------------------------------
 * Restates the event callback of the reference's corner tracker, event-cam-tracking/
 * event-cam-fast-corner-tracker/metavision_time_surface_periodic_group_track.cpp:883-1063 (the code
 * sits inside a lambda of main() that needs the Metavision SDK, so it cannot be compiled here:
 * PARITY UNPINNED by the reference, pinned by hand-derived known answers only).
 *   :888-923   every event of the callback range stamps its pixel of the MostRecentTimestampBuffer
 *              (:786, zero-initialised): surface[y][x] = t, in stream order (the last one wins)
 *   :931-1057  then every event of the range is tested against the UPDATED surface: a streak of
 *              3..6 newest pixels on the 16-pixel circle of radius 3 (:44) whose neighbours on both
 *              sides are older-or-equal and whose oldest member is strictly newer than every other
 *              pixel of the circle, and likewise a streak of 4..8 on the 20-pixel circle of radius
 *              4 (:45); circle entry [k][0] is added to y, [k][1] to x (:962)
 *   :948-955   events closer than 4 pixels to the border: the loop BREAKS there, i.e. the rest of
 *              the range is not tested (literal_break != 0); the evident intent is to skip that
 *              event only (literal_break == 0)
 * flags[i] = 1 when event i is a corner.  Returns the number of corners. */
static const int orc_circle3[16][2] = {{0, 3},  {1, 3},   {2, 2},   {3, 1},  {3, 0},  {3, -1},
                                       {2, -2}, {1, -3},  {0, -3},  {-1, -3}, {-2, -2}, {-3, -1},
                                       {-3, 0}, {-3, 1},  {-2, 2},  {-1, 3}};
static const int orc_circle4[20][2] = {{0, 4},   {1, 4},   {2, 3},   {3, 2},  {4, 1},  {4, 0},  {4, -1},
                                       {3, -2},  {2, -3},  {1, -4},  {0, -4}, {-1, -4}, {-2, -3}, {-3, -2},
                                       {-4, -1}, {-4, 0},  {-4, 1},  {-3, 2}, {-2, 3}, {-1, 4}};
static int orc_streak(const int64_t* s, int W, int x, int y, const int (*c)[2], int n, int smin,
                      int smax) {
#define TS(k) s[(size_t)(y + c[(k)][0]) * W + (x + c[(k)][1])]
    for (int i = 0; i < n; i++) {
        for (int sz = smin; sz <= smax; sz++) {
            if (TS(i) < TS((i - 1 + n) % n)) continue;
            if (TS((i + sz - 1) % n) < TS((i + sz) % n)) continue;
            double min_t = (double)TS(i);
            for (int j = 1; j < sz; j++) {
                const double tj = (double)TS((i + j) % n);
                if (tj < min_t) min_t = tj;
            }
            int did_break = 0;
            for (int j = sz; j < n; j++) {
                const double tj = (double)TS((i + j) % n);
                if (tj >= min_t) {
                    did_break = 1;
                    break;
                }
            }
            if (!did_break) return 1;
        }
    }
#undef TS
    return 0;
}
size_t orc_ts_corners(const evk_event* ev, size_t n, int W, int H, int64_t* surface,
                      int literal_break, uint8_t* flags) {
    for (size_t i = 0; i < n; i++)
        if (ev[i].x < W && ev[i].y < H) surface[(size_t)ev[i].y * W + ev[i].x] = ev[i].t;
    size_t nc = 0;
    memset(flags, 0, n);
    for (size_t i = 0; i < n; i++) {
        const int x = ev[i].x, y = ev[i].y;
        if (x < 4 || x >= W - 4 || y < 4 || y >= H - 4) {
            if (literal_break) break;
            continue;
        }
        if (orc_streak(surface, W, x, y, orc_circle3, 16, 3, 6) &&
            orc_streak(surface, W, x, y, orc_circle4, 20, 4, 8)) {
            flags[i] = 1;
            nc++;
        }
    }
    return nc;
}

/* ---- corner post-processing: box non-maximum suppression ---------------------
 * Restates CornerFilter::filterCorners, event-cam-tracking/event-cam-fast-corner-tracker/
 * metavision_time_surface_periodic_group_track.cpp:81-151 (call site :832, box size 15; the
 * `threshold` argument and the response sort are commented out there, :93-98,112).  Corners are
 * taken in input order; a corner is kept when no pixel of its box [x - h, x + h] x [y - h, y + h]
 * (h = box_size / 2, clipped to the image, :118-121) has been marked by a corner kept before it
 * (:124-136); a kept corner gets label = its rank among the kept ones (:144) and marks its box
 * (:148-151).  Unlike the reference, a corner whose clipped box is empty (centre outside the image)
 * reads nothing here -- the reference would read outside its mask.  Pinned against the reference's
 * own class compiled where it lies (oracle/_ref/libref_fct.so, tests/test_corner_filter.py).
 * kept[i] = index of the i-th kept corner.  Returns the number kept. */
size_t orc_filter_corners(const int32_t* xy, size_t n, int width, int height, int box_size,
                          uint32_t* kept) {
    if (n == 0 || width < 1 || height < 1) return 0;
    uint8_t* mask = (uint8_t*)calloc((size_t)width * height, 1);
    const int half = box_size / 2;
    size_t nk = 0;
    for (size_t i = 0; i < n; i++) {
        const int x = xy[2 * i], y = xy[2 * i + 1];
        const int sx = x - half > 0 ? x - half : 0, ex = x + half < width - 1 ? x + half : width - 1;
        const int sy = y - half > 0 ? y - half : 0, ey = y + half < height - 1 ? y + half : height - 1;
        int is_max = 1;
        for (int yy = sy; yy <= ey && is_max; yy++)
            for (int xx = sx; xx <= ex; xx++)
                if (mask[(size_t)yy * width + xx]) {
                    is_max = 0;
                    break;
                }
        if (is_max) {
            kept[nk++] = (uint32_t)i;
            for (int yy = sy; yy <= ey; yy++)
                for (int xx = sx; xx <= ex; xx++) mask[(size_t)yy * width + xx] = 255;
        }
    }
    free(mask);
    return nk;
}
