// N1: path candidates of the visibility-pruned graph, decoded on the device from the linear index.
//
// Reference: HybridPathTracer.generate_path_candidates (differt/src/differt/geometry/_solvers.py:
// 993-1058) builds DiGraph.from_complete_graph(n), inserts a `from` node connected to the primitives
// visible from the transmitters and a `to` node reachable from those visible from the receivers
// (differt-core/src/geometry/graph.rs:636-691), clears the outgoing edges of masked-out primitives
// (filter_by_mask, fast mode, graph.rs:879-915) and enumerates all paths from → c_1 … c_k → to with a
// DFS whose children are visited in ascending order (graph.rs:1063-1108).  The resulting list is, in
// lexicographic order, every tuple with
//     c_1 ∈ A∩M,  c_i ∈ M,  c_i ≠ c_{i-1},  c_k ∈ B∩M
// (A = visible from TX, B = visible from RX, M = active).  Instead of a single-threaded host DFS and
// a host → device copy, the number of completions g_j(c) of every node at every position is computed
// once (k tiny kernels + prefix sums), after which candidate i is unranked independently by one
// thread: k binary searches over the prefix sums.
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace drt {

inline size_t align256c(size_t x) { return (x + 255) & ~size_t(255); }

struct DiGraphLayout {
    size_t cum, g, cub, cub_bytes, total;  // cum: [order][n+1] int64, g: [n+1] int64
};

inline DiGraphLayout digraph_layout(int64_t n, int order) {
    DiGraphLayout l{};
    const size_t row = align256c(size_t(n + 1) * sizeof(int64_t));
    l.cum = 0;
    l.g = size_t(order > 0 ? order : 1) * row;
    l.cub = l.g + row;
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, static_cast<const int64_t *>(nullptr),
                                  static_cast<int64_t *>(nullptr), static_cast<int>(n + 1));
    l.cub_bytes = align256c(bytes);
    l.total = l.cub + l.cub_bytes;
    return l;
}

// g_j(c) = elig_j(c) * (last ? 1 : S_{j+1} - g_{j+1}(c)),  g_j(n) = 0
__global__ void digraph_level_kernel(int64_t n, int j, int order, const uint8_t *__restrict__ from_mask,
                                     const uint8_t *__restrict__ to_mask,
                                     const uint8_t *__restrict__ active,
                                     const int64_t *__restrict__ cum_next, int64_t *__restrict__ g) {
    const int64_t c = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (c > n) return;
    if (c == n) {
        g[c] = 0;
        return;
    }
    bool ok = active == nullptr || active[c] != 0;
    if (j == 0 && from_mask != nullptr) ok = ok && from_mask[c] != 0;
    if (j == order - 1 && to_mask != nullptr) ok = ok && to_mask[c] != 0;
    int64_t v = 0;
    if (ok) v = (j == order - 1) ? 1 : cum_next[n] - (cum_next[c + 1] - cum_next[c]);
    g[c] = v;
}

__global__ void digraph_total_kernel(const int64_t *cum0, int64_t n, int64_t *total) { *total = cum0[n]; }

__global__ void digraph_decode_kernel(int64_t n, int order, size_t row_elems,
                                      const int64_t *__restrict__ cum, int64_t start, int64_t count,
                                      int mult, int32_t *__restrict__ out) {
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= count) return;
    int64_t r = start + i;
    int64_t prev = -1;
    for (int j = 0; j < order; ++j) {
        const int64_t *cj = cum + size_t(j) * row_elems;
        const int64_t gp = prev >= 0 ? cj[prev + 1] - cj[prev] : 0;
        // smallest c with F(c) = cum[c+1] - (prev <= c ? gp : 0) > r
        int64_t lo = 0, hi = n - 1;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            const int64_t F = cj[mid + 1] - ((prev >= 0 && prev <= mid) ? gp : 0);
            if (F > r) hi = mid; else lo = mid + 1;
        }
        const int64_t c = lo;
        r -= cj[c] - ((prev >= 0 && prev < c) ? gp : 0);
        out[i * order + j] = int32_t(c * mult);
        prev = c;
    }
}

}  // namespace drt

using namespace drt;

extern "C" {

size_t drt_digraph_candidates_workspace_bytes(int64_t num_nodes, int32_t order) {
    if (num_nodes < 0 || order < 0 || num_nodes > (int64_t(1) << 30)) return 0;
    return digraph_layout(num_nodes, order).total;
}

int drt_digraph_candidates_prepare(drt_stream_t stream, int64_t num_nodes, int32_t order,
                                   const uint8_t *from_mask, const uint8_t *to_mask,
                                   const uint8_t *active_mask, void *workspace,
                                   size_t workspace_bytes, int64_t *total_out) {
    if (num_nodes < 0 || order < 0 || num_nodes > (int64_t(1) << 30)) return DRT_ERR_BAD_EXTENT;
    if (order > 2 * DRT_MAX_ORDER) return DRT_ERR_UNSUPPORTED;
    if (total_out == nullptr) return DRT_ERR_NULL_POINTER;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (order == 0 || num_nodes == 0) {
        // order 0: `from` has no edge to `to` (direct_path=False, _solvers.py:1034-1037), so the DFS
        // yields nothing (graph.rs:1076-1092); an empty graph has no candidates either
        return cudaMemsetAsync(total_out, 0, sizeof(int64_t), s) == cudaSuccess ? DRT_OK : DRT_ERR_CUDA;
    }
    if (workspace == nullptr) return DRT_ERR_NULL_POINTER;
    const DiGraphLayout l = digraph_layout(num_nodes, order);
    if (workspace_bytes < l.total) return DRT_ERR_WORKSPACE;
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    const size_t row = align256c(size_t(num_nodes + 1) * sizeof(int64_t));
    int64_t *g = reinterpret_cast<int64_t *>(ws + l.g);
    const unsigned blocks = unsigned((num_nodes + 1 + 255) / 256);
    for (int j = order - 1; j >= 0; --j) {
        int64_t *cum_j = reinterpret_cast<int64_t *>(ws + l.cum + size_t(j) * row);
        const int64_t *cum_next =
            j + 1 < order ? reinterpret_cast<const int64_t *>(ws + l.cum + size_t(j + 1) * row) : nullptr;
        digraph_level_kernel<<<blocks, 256, 0, s>>>(num_nodes, j, order, from_mask, to_mask,
                                                    active_mask, cum_next, g);
        size_t cub_bytes = l.cub_bytes;
        if (cub::DeviceScan::ExclusiveSum(ws + l.cub, cub_bytes, g, cum_j, static_cast<int>(num_nodes + 1),
                                          s) != cudaSuccess)
            return DRT_ERR_CUDA;
    }
    digraph_total_kernel<<<1, 1, 0, s>>>(reinterpret_cast<const int64_t *>(ws + l.cum), num_nodes, total_out);
    return cudaGetLastError() == cudaSuccess ? DRT_OK : DRT_ERR_CUDA;
}

int drt_digraph_candidates(drt_stream_t stream, int64_t num_nodes, int32_t order, const void *workspace,
                           int64_t start, int64_t count, int32_t stride_multiplier, int32_t *out) {
    if (num_nodes < 0 || order < 0 || start < 0 || count < 0) return DRT_ERR_BAD_EXTENT;
    if (order > 2 * DRT_MAX_ORDER) return DRT_ERR_UNSUPPORTED;
    if (count == 0 || order == 0) return DRT_OK;
    if (num_nodes == 0) return DRT_ERR_BAD_EXTENT;
    if (workspace == nullptr || out == nullptr) return DRT_ERR_NULL_POINTER;
    const DiGraphLayout l = digraph_layout(num_nodes, order);
    const size_t row_elems = align256c(size_t(num_nodes + 1) * sizeof(int64_t)) / sizeof(int64_t);
    digraph_decode_kernel<<<unsigned((count + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        num_nodes, order, row_elems,
        reinterpret_cast<const int64_t *>(static_cast<const unsigned char *>(workspace) + l.cum), start,
        count, stride_multiplier > 0 ? stride_multiplier : 1, out);
    return cudaGetLastError() == cudaSuccess ? DRT_OK : DRT_ERR_CUDA;
}

}  // extern "C"
