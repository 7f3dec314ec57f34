// Any-hit pack ordering: triangles sorted by descending area.
//
// `ray_intersect_any_triangle` (reference differt/src/differt/geometry/_utils.py:1353-1537) is an OR
// over triangles, so the order in which they are tested cannot change the result — only how soon a
// blocked ray can stop.  The probability that a random segment crosses a triangle is proportional
// to its area, so the all-pairs engine keeps the first tiles of an area-sorted pack resident in shared
// memory and tests every new ray against them first (intersect_core.cuh, "head tiles").
// Triangle indices are lost, so this ordering is never used for first-hit / visibility queries.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace drt {

// key = bits of |e1 x e2|^2 (non-negative float: its bit pattern is order-preserving);
// never-hit records (NaN origin: padding and masked-out triangles) get key 0 and sort last
__global__ void area_keys_kernel(int64_t n, const Tri48 *__restrict__ pack, uint32_t *__restrict__ keys,
                                 uint32_t *__restrict__ idx) {
    const int64_t j = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (j >= n) return;
    const float4 a = pack[j].a, b = pack[j].b, c = pack[j].c;
    const Tri t = unpack(a, b, c);
    const float3 nrm = cross3(t.e1, t.e2);
    float area2 = dot3(nrm, nrm);
    if (!(area2 > 0.0f) || a.x != a.x) area2 = 0.0f;  // NaN / never-hit / degenerate
    if (isinf(area2)) area2 = 3.0e38f;
    keys[j] = __float_as_uint(area2);
    idx[j] = static_cast<uint32_t>(j);
}

__global__ void iota_kernel(int64_t n, uint32_t *__restrict__ idx) {
    const int64_t j = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (j < n) idx[j] = static_cast<uint32_t>(j);
}

__global__ void gather_pack_kernel(int64_t n, const Tri48 *__restrict__ in, const uint32_t *__restrict__ idx,
                                   Tri48 *__restrict__ out) {
    const int64_t j = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (j >= n) return;
    const Tri48 t = in[idx[j]];
    out[j] = t;
}

inline size_t align256s(size_t x) { return (x + 255) & ~size_t(255); }

struct SortLayout {
    size_t keys_in, keys_out, idx_in, idx_out, cub, cub_bytes, total;
};

inline SortLayout sort_layout(int64_t n) {
    SortLayout l{};
    const size_t arr = align256s(size_t(n) * sizeof(uint32_t));
    l.keys_in = 0;
    l.keys_out = arr;
    l.idx_in = 2 * arr;
    l.idx_out = sort_indices_offset(n);  // = 3 * arr: read by the culled first-hit (walk.cuh)
    l.cub = 4 * arr;
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, static_cast<const uint32_t *>(nullptr),
                                              static_cast<uint32_t *>(nullptr),
                                              static_cast<const uint32_t *>(nullptr),
                                              static_cast<uint32_t *>(nullptr), static_cast<int>(n));
    l.cub_bytes = align256s(bytes);
    l.total = l.cub + l.cub_bytes;
    return l;
}

}  // namespace drt

using namespace drt;

extern "C" {

size_t drt_mesh_pack_sort_workspace_bytes(int64_t num_triangles) {
    if (num_triangles < 0 || num_triangles > (int64_t(1) << 30)) return 0;
    return sort_layout(int64_t(drt_mesh_pack_bytes(num_triangles) / sizeof(Tri48))).total;
}

int drt_mesh_pack_sort_by_area(drt_stream_t stream, int64_t num_triangles, const void *pack_in,
                               void *workspace, size_t workspace_bytes, void *pack_out) {
    if (num_triangles < 0 || num_triangles > (int64_t(1) << 30)) return DRT_ERR_BAD_EXTENT;
    if (!pack_in || !pack_out || !workspace) return DRT_ERR_NULL_POINTER;
    if (pack_in == pack_out) return DRT_ERR_UNSUPPORTED;
    const int64_t n = int64_t(drt_mesh_pack_bytes(num_triangles) / sizeof(Tri48));
    const SortLayout l = sort_layout(n);
    if (workspace_bytes < l.total) return DRT_ERR_WORKSPACE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    uint32_t *keys_in = reinterpret_cast<uint32_t *>(ws + l.keys_in);
    uint32_t *keys_out = reinterpret_cast<uint32_t *>(ws + l.keys_out);
    uint32_t *idx_in = reinterpret_cast<uint32_t *>(ws + l.idx_in);
    uint32_t *idx_out = reinterpret_cast<uint32_t *>(ws + l.idx_out);
    const unsigned blocks = unsigned((n + 255) / 256);
    area_keys_kernel<<<blocks, 256, 0, s>>>(n, static_cast<const Tri48 *>(pack_in), keys_in, idx_in);
    size_t cub_bytes = l.cub_bytes;
    if (cub::DeviceRadixSort::SortPairsDescending(ws + l.cub, cub_bytes, keys_in, keys_out, idx_in,
                                                  idx_out, static_cast<int>(n), 0, 32, s) != cudaSuccess)
        return DRT_ERR_CUDA;
    gather_pack_kernel<<<blocks, 256, 0, s>>>(n, static_cast<const Tri48 *>(pack_in), idx_out,
                                              static_cast<Tri48 *>(pack_out));
    return cudaGetLastError() == cudaSuccess ? DRT_OK : DRT_ERR_CUDA;
}

}  // extern "C"

// internal: the same for an explicit number of records (a sub-range of a pack)
namespace drt {
int drt_sort_records_by_keys(drt_stream_t stream, int64_t n, const void *pack_in, const uint32_t *keys,
                             void *workspace, size_t workspace_bytes, void *pack_out);
}

extern "C" {

int drt_mesh_pack_sort_by_keys(drt_stream_t stream, int64_t num_triangles, const void *pack_in,
                               const uint32_t *keys, void *workspace, size_t workspace_bytes,
                               void *pack_out) {
    if (num_triangles < 0 || num_triangles > (int64_t(1) << 30)) return DRT_ERR_BAD_EXTENT;
    return drt_sort_records_by_keys(stream, int64_t(drt_mesh_pack_bytes(num_triangles) / sizeof(Tri48)), pack_in,
                                    keys, workspace, workspace_bytes, pack_out);
}

}  // extern "C"

int drt::drt_sort_records_by_keys(drt_stream_t stream, int64_t n, const void *pack_in, const uint32_t *keys,
                                  void *workspace, size_t workspace_bytes, void *pack_out) {
    if (n <= 0) return DRT_ERR_BAD_EXTENT;
    if (!pack_in || !pack_out || !workspace || !keys) return DRT_ERR_NULL_POINTER;
    if (pack_in == pack_out) return DRT_ERR_UNSUPPORTED;
    const SortLayout l = sort_layout(n);
    if (workspace_bytes < l.total) return DRT_ERR_WORKSPACE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    uint32_t *keys_out = reinterpret_cast<uint32_t *>(ws + l.keys_out);
    uint32_t *idx_in = reinterpret_cast<uint32_t *>(ws + l.idx_in);
    uint32_t *idx_out = reinterpret_cast<uint32_t *>(ws + l.idx_out);
    const unsigned blocks = unsigned((n + 255) / 256);
    iota_kernel<<<blocks, 256, 0, s>>>(n, idx_in);
    size_t cub_bytes = l.cub_bytes;
    if (cub::DeviceRadixSort::SortPairsDescending(ws + l.cub, cub_bytes, keys, keys_out, idx_in, idx_out,
                                                  static_cast<int>(n), 0, 32, s) != cudaSuccess)
        return DRT_ERR_CUDA;
    gather_pack_kernel<<<blocks, 256, 0, s>>>(n, static_cast<const Tri48 *>(pack_in), idx_out,
                                              static_cast<Tri48 *>(pack_out));
    return cudaGetLastError() == cudaSuccess ? DRT_OK : DRT_ERR_CUDA;
}
