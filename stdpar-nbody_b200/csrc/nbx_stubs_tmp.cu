// TEMPORARY: replaced by nbx_sort.cu / nbx_bvh.cu / nbx_octree.cu
#include "nbx_internal.cuh"
namespace nbx {
int octree_create(nbx_engine*) { return fail(NBX_ERR_INVALID, "not implemented"); }
void octree_destroy(nbx_engine*) {}
int octree_build(nbx_engine*) { return fail(NBX_ERR_INVALID, "not implemented"); }
int octree_compute_force(nbx_engine*) { return fail(NBX_ERR_INVALID, "not implemented"); }
int octree_get_root(nbx_engine*, void*, void*, uint64_t*) { return fail(NBX_ERR_INVALID, "not implemented"); }
int octree_get_canonical(nbx_engine*, uint64_t*, uint32_t*, uint64_t*, uint32_t*, void*) { return fail(NBX_ERR_INVALID, "not implemented"); }
}
