/* ref_driver.c — C entry points around the reference's OpenCL kernels compiled as C through
 * cl_shim.h (TEST INFRASTRUCTURE ONLY).  The kernels themselves are the reference's files, compiled
 * where they lie under /root/reference (Makefile target _ref/libref.so); nothing of them is copied
 * into this repository.  Launch shapes follow the reference host code:
 *   process_coordinates : ACCEL/metavision_sdk_get_started5_opencl_store.cpp:301-314, here with one
 *                         work-item (global size 1) = the sequential run of the kernel
 *   assign_to_centers   : KM/assign_to_centers2.c:226-240, one work-item per point
 *   assign_data_cluster : KM/assign_to_centers2.c:309-330, one work-item per point */
#include <stdatomic.h>
#include <stddef.h>
#include <string.h>

_Thread_local size_t ref_gid, ref_gsize = 1, ref_lid, ref_lsize = 1, ref_group;

void process_coordinates(const int* input_coords, int* repeated_coords, int* unique_coords,
                         int* repeated_count, int* unique_count, const int total_coords,
                         const int width, const int height);
void assign_to_centers(float* data, float* centers, int* assignments);
void assign_data_cluster(float* data, unsigned int* assign, atomic_int* cluster_index,
                         float* output);

/* counters are cumulative in the reference (never reset by the kernel): the caller passes them in */
void ref_process_coordinates(const int* coords, int total_coords, int* unique_coords,
                             int* unique_count, int* repeated_count) {
    ref_gid = ref_lid = ref_group = 0;
    ref_gsize = ref_lsize = 1;
    process_coordinates(coords, NULL, unique_coords, repeated_count, unique_count, total_coords,
                        1280, 720);
}

void ref_assign_to_centers(float* data, float* centers, int* assignments, int n_points) {
    ref_gsize = (size_t)n_points;
    ref_lsize = 1;
    for (int i = 0; i < n_points; i++) {
        ref_gid = (size_t)i;
        ref_lid = 0;
        ref_group = (size_t)i;
        assign_to_centers(data, centers, assignments);
    }
}

/* output: 8 slabs of 4096 floats (2048 x then 2048 y), cluster_index: 8 counters */
void ref_assign_data_cluster(float* data, int* assignments, int* cluster_index, float* output,
                             int n_points) {
    atomic_int idx[8];
    for (int k = 0; k < 8; k++) atomic_init(&idx[k], cluster_index[k]);
    ref_gsize = (size_t)n_points;
    for (int i = 0; i < n_points; i++) {
        ref_gid = (size_t)i;
        assign_data_cluster(data, (unsigned int*)assignments, idx, output);
    }
    for (int k = 0; k < 8; k++) cluster_index[k] = atomic_load(&idx[k]);
}
