// tile_pipe.cuh -- persistent-tile plumbing shared by the FFT kernels: cp.async (LDGSTS) helpers
// with zero fill, the 4-slot tile-descriptor ring, L2 prefetch and programmatic-dependent-launch
// hooks.
#pragma once
#include "spev_internal.cuh"

namespace spev {

struct BatchView {
    int n_ftiles, n_ctiles;
    const spev_tile* ftiles;
    const spev_tile* ctiles;
};

static BatchView view_of(const spev_batch* b) {
    BatchView v;
    v.n_ftiles = b->n_ftiles; v.n_ctiles = b->n_ctiles;
    v.ftiles = b->ftiles; v.ctiles = b->ctiles;
    return v;
}

constexpr float kTiny = 1.17549435e-38f;   // np.finfo(np.float32).tiny
constexpr int kRing = 4;

// ---------------------------------------------------------------------------------------
// cp.async helpers (LDGSTS): zero-fill through the src-size operand
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem, int src_bytes) {
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// Warp-cooperative L2 prefetch of [p, p+bytes): moves DRAM->L2 traffic off the critical path
// (issued at tile start, consumed after the FFT phase), no registers or shared memory needed.
__device__ __forceinline__ void warp_prefetch_l2(const void* p, int bytes, int lane) {
    const uintptr_t base = reinterpret_cast<uintptr_t>(p) & ~static_cast<uintptr_t>(127);
    const int n = static_cast<int>((reinterpret_cast<uintptr_t>(p) + bytes - base + 127) >> 7);
    for (int i = lane; i < n; i += 32)
        asm volatile("prefetch.global.L2 [%0];\n" ::"l"(base + (static_cast<uintptr_t>(i) << 7)));
}

__device__ __forceinline__ void fetch_desc(spev_tile* slot, const spev_tile* g) {
    static_assert(sizeof(spev_tile) == 48, "spev_tile must be 48 bytes");
    if (threadIdx.x == 0) {
        cp_async16(reinterpret_cast<char*>(slot), reinterpret_cast<const char*>(g), 16);
        cp_async16(reinterpret_cast<char*>(slot) + 16, reinterpret_cast<const char*>(g) + 16, 16);
        cp_async16(reinterpret_cast<char*>(slot) + 32, reinterpret_cast<const char*>(g) + 32, 16);
    }
}

// Programmatic dependent launch: a kernel launched with the stream-serialization attribute may
// start while its predecessor is still draining; everything before pdl_wait() must only touch
// data the predecessor does not write (constant tables, tile descriptors).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }

// Persistent-loop bookkeeping shared by the kernels: prologue of the descriptor ring.
__device__ __forceinline__ int ring_prologue(spev_tile* s_ring, const spev_tile* tiles, int n_tiles) {
    const int first = blockIdx.x, stride = gridDim.x;
    const int my_n = first < n_tiles ? (n_tiles - first + stride - 1) / stride : 0;
    for (int j = 0; j < 3 && j < my_n; ++j) fetch_desc(s_ring + j, tiles + first + j * stride);
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();
    return my_n;
}

}  // namespace spev
