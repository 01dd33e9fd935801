// Minimal stand-in for the PCL types the reference's DBSCAN_simple.h uses (TEST INFRASTRUCTURE ONLY;
// PCL is absent from this image).  It exists so that
//   event-cam-clustering/point-cloud-clustering/DBSCAN_simple.h
// compiles unmodified, where it lies, into oracle/_ref/libref_dbscan.so (oracle/Makefile, target
// ref_dbscan).  DBSCANSimpleCluster does its own brute-force radius search, so only the containers
// are needed: PointXYZ, PointCloud<T>::points / header / Ptr, PointIndices, search::KdTree<T>::Ptr
// (a pointer type that is never dereferenced by the simple variant).  The standard headers the
// real <pcl/point_types.h> drags in and DBSCAN_simple.h relies on are included here.
#ifndef EVK_PCL_SHIM_POINT_TYPES
#define EVK_PCL_SHIM_POINT_TYPES
#include <algorithm>
#include <cmath>
#include <limits>
#include <memory>
#include <vector>

namespace pcl {
struct PCLHeader {
    unsigned seq = 0;
};
struct PointXYZ {
    float x, y, z;
};
struct PointIndices {
    PCLHeader header;
    std::vector<int> indices;
};
template <typename PointT>
struct PointCloud {
    typedef std::shared_ptr<PointCloud<PointT>> Ptr;
    PCLHeader header;
    std::vector<PointT> points;
};
namespace search {
template <typename PointT>
struct KdTree {
    typedef std::shared_ptr<KdTree<PointT>> Ptr;
};
}  // namespace search
}  // namespace pcl
#endif
