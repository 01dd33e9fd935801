// ref_aec_driver.cpp -- C entry points around the REFERENCE's asynchronous event clustering
// consumer (TEST INFRASTRUCTURE ONLY).  AEClustering.cpp and MyCluster.cpp are the reference's
// files, compiled where they lie under /root/reference against oracle/eigen_shim (Makefile target
// _ref/libref_aec.so); nothing of them is copied into this repository.  Call shapes follow the
// reference host code: `AEClustering *eclustering(new AEClustering)` with NO init() call
// (ACCEL/metavision_sdk_get_started5_opencl_store.cpp:42) -> use_init = 0; init() is the class's
// own public way to set parameters (AEClustering.h:33) -> use_init = 1.
#include <cstdlib>
#include <deque>
#include <iostream>

#include "AEClustering.h"

extern "C" {

void* ref_aec_new(int use_init, int sz_buffer, double radius, int kappa, double alpha, int min_n,
                  unsigned rand_seed) {
    std::cout.setstate(std::ios_base::failbit);  // the classes narrate every step on stdout
    std::srand(rand_seed);                       // MyCluster.cpp:88 draws from std::rand()
    AEClustering* a = new AEClustering;
    if (use_init) a->init(sz_buffer, radius, kappa, alpha, min_n);
    return a;
}
void ref_aec_free(void* p) { delete static_cast<AEClustering*>(p); }

// n calls of AEClustering::update (AEClustering.h:37); e = n x {t, x, y, p}
void ref_aec_update(void* p, const double* e, long n) {
    AEClustering* a = static_cast<AEClustering*>(p);
    for (long i = 0; i < n; i++) {
        std::deque<double> ev{e[4 * i], e[4 * i + 1], e[4 * i + 2], e[4 * i + 3]};
        a->update(ev);
    }
}
int ref_aec_n_clusters(void* p) { return (int)static_cast<AEClustering*>(p)->clusters.size(); }
int ref_aec_last_updated(void* p) {
    return static_cast<AEClustering*>(p)->getLastUpdatedClusterIdx();
}
int ref_aec_min_n(void* p) { return static_cast<AEClustering*>(p)->getMinN(); }
// per cluster, in deque order: id, n, mu[2], getClusterCentroid()[2] (NaN when the cluster is empty)
void ref_aec_get_clusters(void* p, int* ids, int* ns, double* mu, double* cen) {
    AEClustering* a = static_cast<AEClustering*>(p);
    for (std::size_t c = 0; c < a->clusters.size(); c++) {
        MyCluster& k = a->clusters[c];
        ids[c] = k.getClusterId();
        ns[c] = k.getN();
        Eigen::VectorXd m(k.getMu());
        mu[2 * c] = m[0];
        mu[2 * c + 1] = m[1];
        Eigen::VectorXd g(k.getClusterCentroid());
        cen[2 * c] = g[0];
        cen[2 * c + 1] = g[1];
    }
}
// the stored events of cluster c, front to back: event id, x, y, t (relative), polarity
int ref_aec_get_points(void* p, int c, int* ids, double* xy, double* t, int* pol, int cap) {
    MyCluster& k = static_cast<AEClustering*>(p)->clusters[(std::size_t)c];
    std::deque<int> id = k.getDatId();
    std::deque<Eigen::VectorXd> d = k.getDat();
    std::deque<double> tt = k.getDatT();
    std::deque<bool> pp = k.getDatPol();
    const int n = (int)d.size();
    for (int i = 0; i < n && i < cap; i++) {
        ids[i] = id[(std::size_t)i];
        xy[2 * i] = d[(std::size_t)i][0];
        xy[2 * i + 1] = d[(std::size_t)i][1];
        t[i] = tt[(std::size_t)i];
        pol[i] = pp[(std::size_t)i] ? 1 : 0;
    }
    return n;
}
}
