// TEMPORARY: replaced by nbx_sort.cu / nbx_bvh.cu / nbx_octree.cu
#include "nbx_internal.cuh"
namespace nbx {
int sorter_create(nbx_engine*, uint32_t) { return fail(NBX_ERR_INVALID, "not implemented"); }
void sorter_destroy(nbx_engine*) {}
int sort_pairs(nbx_engine*, const uint64_t*, uint32_t, int, uint32_t*, uint64_t*) { return fail(NBX_ERR_INVALID, "not implemented"); }
int bvh_create(nbx_engine*) { return fail(NBX_ERR_INVALID, "not implemented"); }
void bvh_destroy(nbx_engine*) {}
int bvh_bounding_box(nbx_engine*) { return fail(NBX_ERR_INVALID, "not implemented"); }
int bvh_hilbert_sort(nbx_engine*) { return fail(NBX_ERR_INVALID, "not implemented"); }
int bvh_build_tree(nbx_engine*) { return fail(NBX_ERR_INVALID, "not implemented"); }
int bvh_compute_force(nbx_engine*) { return fail(NBX_ERR_INVALID, "not implemented"); }
int bvh_get_bbox(nbx_engine*, void*, void*) { return fail(NBX_ERR_INVALID, "not implemented"); }
int bvh_get_keys(nbx_engine*, uint64_t*, uint32_t*) { return fail(NBX_ERR_INVALID, "not implemented"); }
int bvh_get_nodes(nbx_engine*, uint64_t*, void*, void*, void*) { return fail(NBX_ERR_INVALID, "not implemented"); }
int octree_create(nbx_engine*) { return fail(NBX_ERR_INVALID, "not implemented"); }
void octree_destroy(nbx_engine*) {}
int octree_build(nbx_engine*) { return fail(NBX_ERR_INVALID, "not implemented"); }
int octree_compute_force(nbx_engine*) { return fail(NBX_ERR_INVALID, "not implemented"); }
int octree_get_root(nbx_engine*, void*, void*, uint64_t*) { return fail(NBX_ERR_INVALID, "not implemented"); }
int octree_get_canonical(nbx_engine*, uint64_t*, uint32_t*, uint64_t*, uint32_t*, void*) { return fail(NBX_ERR_INVALID, "not implemented"); }
}
