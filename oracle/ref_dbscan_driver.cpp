// ref_dbscan_driver.cpp -- C entry point around the REFERENCE's DBSCANSimpleCluster (TEST
// INFRASTRUCTURE ONLY).  DBSCAN_simple.h is the reference's file, compiled where it lies under
// /root/reference against oracle/pcl_shim (Makefile target _ref/libref_dbscan.so); nothing of it is
// copied into this repository.  Parameters as the reference app sets them
// (event-cam-clustering/point-cloud-clustering/pcl_cluster.cpp:112-120: setCorePointMinPts,
// setClusterTolerance, setMinClusterSize, setMaxClusterSize, extract).  The app instantiates the
// kd-tree variant (DBSCAN_kdtree.h), which overrides only the radius search; the simple variant's
// brute-force search returns the same neighbour sets.
#include <pcl/point_types.h>

#include "DBSCAN_simple.h"

extern "C" {
// xyz: n x 3 floats.  Writes the clusters in the order extract() returns them: sizes[c], then the
// (sorted) member indices of every cluster back to back in members[].  Returns the number of
// clusters; *n_members the total member count (the call can be made with members = NULL first).
int ref_dbscan_simple(const float* xyz, int n, double eps, int min_pts, int min_cluster,
                      int max_cluster, int* sizes, int cap_clusters, int* members, long cap_members,
                      long* n_members) {
    pcl::PointCloud<pcl::PointXYZ>::Ptr cloud(new pcl::PointCloud<pcl::PointXYZ>);
    cloud->points.resize((size_t)n);
    for (int i = 0; i < n; i++) cloud->points[(size_t)i] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    DBSCANSimpleCluster<pcl::PointXYZ> ec;
    ec.setCorePointMinPts(min_pts);
    ec.setClusterTolerance(eps);
    ec.setMinClusterSize(min_cluster);
    ec.setMaxClusterSize(max_cluster);
    ec.setInputCloud(cloud);
    std::vector<pcl::PointIndices> out;
    ec.extract(out);
    long total = 0;
    for (size_t c = 0; c < out.size(); c++) {
        if ((int)c < cap_clusters && sizes) sizes[c] = (int)out[c].indices.size();
        for (int v : out[c].indices) {
            if (members && total < cap_members) members[total] = v;
            total++;
        }
    }
    if (n_members) *n_members = total;
    return (int)out.size();
}
}
