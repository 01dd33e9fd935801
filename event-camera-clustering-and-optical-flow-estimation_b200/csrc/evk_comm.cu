// evk_comm.cu — multi-GPU (one process per GPU) entry points.  Placeholder until the sharded
// path lands: every call reports EVK_ERR_COMM.
#include "evk_internal.cuh"

extern "C" {
int evk_comm_unique_id(uint8_t*) { return EVK_ERR_COMM; }
int evk_comm_init(evk_handle* h, int, int, const uint8_t*) {
    return evk_fail(h, EVK_ERR_COMM, "sharded path not built");
}
int evk_comm_destroy(evk_handle*) { return EVK_OK; }
int evk_set_shard(evk_handle* h, uint64_t first_global_index) {
    if (!h) return EVK_ERR_INVALID;
    h->shard_first = first_global_index;
    return EVK_OK;
}
int evk_downsample_sharded(evk_handle* h, const evk_ds_params*, int, size_t*, size_t*) {
    return evk_fail(h, EVK_ERR_COMM, "sharded path not built");
}
int evk_kmeans_sharded(evk_handle* h, const evk_km_params*, int*) {
    return evk_fail(h, EVK_ERR_COMM, "sharded path not built");
}
int evk_init_centroids_first_k_sharded(evk_handle* h, const evk_km_params*) {
    return evk_fail(h, EVK_ERR_COMM, "sharded path not built");
}
}
