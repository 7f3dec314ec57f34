// Mesh packing + K1 (element-wise Möller–Trumbore), K2 (any-hit), K3 (first-hit), K4 (visibility).
// Reference: differt/src/differt/geometry/_utils.py:1157-1960 and the Warp launchers
// differt/src/differt/geometry/_mesh.py:142-223, 347-401.
#include "intersect_core.cuh"
#include "walk.cuh"

namespace drt {

// ------------------------------------------------------------------------------------------------
// packing
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ void store_tri(Tri48 *out, float3 v0, float3 v1, float3 v2) {
    const float3 e1 = sub3(v1, v0), e2 = sub3(v2, v0);
    const float3 n = unit_normal(v0, v1, v2);
    out->a = make_float4(v0.x, v0.y, v0.z, e1.x);
    out->b = make_float4(e1.y, e1.z, e2.x, e2.y);
    out->c = make_float4(e2.z, n.x, n.y, n.z);
}

__device__ __forceinline__ void store_never_hit(Tri48 *out) {
    // NaN origin: every comparison of the intersection test is false
    const float q = CUDART_NAN_F;
    out->a = make_float4(__uint_as_float(kNeverHitBits), q, q, 0.f);
    out->b = make_float4(0.f, 0.f, 0.f, 0.f);
    out->c = make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void pack_indexed_kernel(int64_t V, int64_t T, int64_t T_pad, const float *__restrict__ verts,
                                    const int32_t *__restrict__ tris,
                                    const uint8_t *__restrict__ mask, Tri48 *__restrict__ out) {
    const int64_t j = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (j >= T_pad) return;
    if (j >= T || (mask != nullptr && mask[j] == 0)) {
        store_never_hit(out + j);
        return;
    }
    int64_t i0 = tris[3 * j], i1 = tris[3 * j + 1], i2 = tris[3 * j + 2];
    i0 = min(max(i0, int64_t(0)), V - 1);
    i1 = min(max(i1, int64_t(0)), V - 1);
    i2 = min(max(i2, int64_t(0)), V - 1);
    store_tri(out + j, ld3(verts + 3 * i0), ld3(verts + 3 * i1), ld3(verts + 3 * i2));
}

__global__ void pack_triangle_vertices_kernel(int64_t T, int64_t T_pad, const float *__restrict__ tv,
                                              const uint8_t *__restrict__ mask,
                                              Tri48 *__restrict__ out) {
    const int64_t j = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (j >= T_pad) return;
    if (j >= T || (mask != nullptr && mask[j] == 0)) {
        store_never_hit(out + j);
        return;
    }
    store_tri(out + j, ld3(tv + 9 * j), ld3(tv + 9 * j + 3), ld3(tv + 9 * j + 6));
}

inline int64_t padded_triangles(int64_t T) {
    const int64_t tiles = (T + kTile - 1) / kTile;
    return (tiles > 0 ? tiles : 1) * kTile;
}

// ------------------------------------------------------------------------------------------------
// K1
// ------------------------------------------------------------------------------------------------

__global__ void mt_elementwise_kernel(int64_t n, Batch4 bt, const float *__restrict__ o,
                                      const float *__restrict__ d, const float *__restrict__ tri,
                                      float eps, float *__restrict__ t_out,
                                      uint8_t *__restrict__ hit_out) {
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += stride) {
        int64_t oo, od, ot, unused;
        bt.offsets(i, oo, od, ot, unused);
        const float3 v0 = ld3(tri + ot), v1 = ld3(tri + ot + 3), v2 = ld3(tri + ot + 6);
        Tri tr;
        tr.v0 = v0;
        tr.e1 = sub3(v1, v0);
        tr.e2 = sub3(v2, v0);
        float t;
        const bool hit = mt_exact(ld3(o + oo), ld3(d + od), tr, eps, t);
        t_out[i] = t;
        hit_out[i] = hit ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------------
// sources / sinks of the all-pairs engine
// ------------------------------------------------------------------------------------------------

template <int RPW>
struct FlatRays {
    const float *o, *d;
    int64_t R;
    __device__ __forceinline__ uint32_t load(int64_t unit, float3 (&oo)[RPW], float3 (&dd)[RPW]) const {
        uint32_t valid = 0;
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int64_t ray = unit * RPW + r;
            if (ray < R) {
                oo[r] = ld3(o + 3 * ray);
                dd[r] = ld3(d + 3 * ray);
                valid |= 1u << r;
            } else {
                oo[r] = make_float3(0.f, 0.f, 0.f);
                dd[r] = make_float3(0.f, 0.f, 0.f);
            }
        }
        return valid;
    }
};

// rays of K4: origin = vertices[ray / n_rays], direction = dirs[ray]
template <int RPW>
struct VertexRays {
    const float *vertices, *d;
    int64_t R, n_rays;
    __device__ __forceinline__ uint32_t load(int64_t unit, float3 (&oo)[RPW], float3 (&dd)[RPW]) const {
        uint32_t valid = 0;
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int64_t ray = unit * RPW + r;
            if (ray < R) {
                oo[r] = ld3(vertices + 3 * (ray / n_rays));
                dd[r] = ld3(d + 3 * ray);
                valid |= 1u << r;
            } else {
                oo[r] = make_float3(0.f, 0.f, 0.f);
                dd[r] = make_float3(0.f, 0.f, 0.f);
            }
        }
        return valid;
    }
};

template <int RPW>
struct AnySink {
    uint8_t *out;
    __device__ __forceinline__ void any(int64_t unit, uint32_t hit, uint32_t valid) const {
#pragma unroll
        for (int r = 0; r < RPW; ++r)
            if (valid & (1u << r)) out[unit * RPW + r] = (hit >> r) & 1u;
    }
    __device__ __forceinline__ void first(int64_t, int, int32_t, float) const {}
};

template <int RPW>
struct FirstSink {
    int32_t *idx;
    float *t;
    __device__ __forceinline__ void any(int64_t, uint32_t, uint32_t) const {}
    __device__ __forceinline__ void first(int64_t unit, int r, int32_t i, float tt) const {
        const bool fin = isfinite(tt);  // _utils.py:1957-1959
        idx[unit * RPW + r] = fin ? i : -1;
        t[unit * RPW + r] = fin ? tt : CUDART_INF_F;
    }
};

// K4: visible[batch, face] = 1 (concurrent identical stores are benign, cf. _mesh.py:365)
template <int RPW>
struct VisibleSink {
    uint8_t *out;
    int64_t n_rays, T;
    __device__ __forceinline__ void any(int64_t, uint32_t, uint32_t) const {}
    __device__ __forceinline__ void first(int64_t unit, int r, int32_t i, float tt) const {
        if (i >= 0 && isfinite(tt)) out[((unit * RPW + r) / n_rays) * T + i] = 1;
    }
};

__global__ void fill_first_miss_kernel(int64_t R, int32_t *idx, float *t) {
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i < R) {
        idx[i] = -1;
        t[i] = CUDART_INF_F;
    }
}

// Rays per warp: 4 amortises each triangle load over four tests and gives four independent chains,
// but a small batch then yields too few warps to fill 148 SMs — fewer rays per warp below ~2^17 rays
// (the reference's own benchmark harness launches 10 000).
inline int rays_per_warp(int64_t R) { return R >= (int64_t(1) << 17) ? 4 : (R >= (int64_t(1) << 15) ? 2 : 1); }

#define DRT_DISPATCH_RPW(R, ...)                                \
    switch (rays_per_warp(R)) {                                  \
        case 4: { constexpr int RPW = 4; __VA_ARGS__ } break;    \
        case 2: { constexpr int RPW = 2; __VA_ARGS__ } break;    \
        default: { constexpr int RPW = 1; __VA_ARGS__ } break;   \
    }

}  // namespace drt

using namespace drt;

#define DRT_CHECK_CUDA(expr)                   \
    do {                                       \
        cudaError_t e__ = (expr);              \
        if (e__ != cudaSuccess) return DRT_ERR_CUDA; \
    } while (0)

extern "C" {

int drt_abi_version(void) { return DRT_ABI_VERSION; }

const char *drt_error_string(int code) {
    switch (code) {
        case DRT_OK: return "ok";
        case DRT_ERR_NULL_POINTER: return "a required pointer is NULL";
        case DRT_ERR_BAD_EXTENT: return "negative or out-of-range extent";
        case DRT_ERR_UNSUPPORTED: return "unsupported configuration (order or batch rank too large)";
        case DRT_ERR_WORKSPACE: return "workspace too small";
        case DRT_ERR_CUDA: return "CUDA runtime error";
        default: return "unknown error";
    }
}

size_t drt_mesh_pack_bytes(int64_t num_triangles) {
    if (num_triangles < 0) return 0;
    return size_t(padded_triangles(num_triangles)) * sizeof(Tri48);
}

int drt_mesh_pack(drt_stream_t stream, int64_t V, int64_t T, const float *vertices,
                  const int32_t *triangles, const uint8_t *mask, void *pack_out) {
    if (V < 0 || T < 0 || T > (int64_t(1) << 30)) return DRT_ERR_BAD_EXTENT;
    if (pack_out == nullptr || (T > 0 && (vertices == nullptr || triangles == nullptr)))
        return DRT_ERR_NULL_POINTER;
    if (T > 0 && V == 0) return DRT_ERR_BAD_EXTENT;
    const int64_t T_pad = padded_triangles(T);
    const int threads = 256;
    pack_indexed_kernel<<<unsigned((T_pad + threads - 1) / threads), threads, 0,
                          static_cast<cudaStream_t>(stream)>>>(
        V, T, T_pad, vertices, triangles, mask, static_cast<Tri48 *>(pack_out));
    DRT_CHECK_CUDA(cudaGetLastError());
    return DRT_OK;
}

int drt_mesh_pack_triangle_vertices(drt_stream_t stream, int64_t T, const float *tv,
                                    const uint8_t *mask, void *pack_out) {
    if (T < 0 || T > (int64_t(1) << 30)) return DRT_ERR_BAD_EXTENT;
    if (pack_out == nullptr || (T > 0 && tv == nullptr)) return DRT_ERR_NULL_POINTER;
    const int64_t T_pad = padded_triangles(T);
    const int threads = 256;
    pack_triangle_vertices_kernel<<<unsigned((T_pad + threads - 1) / threads), threads, 0,
                                    static_cast<cudaStream_t>(stream)>>>(
        T, T_pad, tv, mask, static_cast<Tri48 *>(pack_out));
    DRT_CHECK_CUDA(cudaGetLastError());
    return DRT_OK;
}

int drt_ray_intersect_triangle(drt_stream_t stream, int32_t ndim, const int64_t *shape,
                               const float *o, const int64_t *os, const float *d, const int64_t *ds,
                               const float *tri, const int64_t *ts, float epsilon, float *t_out,
                               uint8_t *hit_out) {
    if (ndim < 0 || ndim > DRT_MAX_BATCH_DIMS) return DRT_ERR_UNSUPPORTED;
    if (ndim > 0 && (shape == nullptr || os == nullptr || ds == nullptr || ts == nullptr))
        return DRT_ERR_NULL_POINTER;
    Batch4 bt;
    int64_t n = 1;
    for (int i = 0; i < 4; ++i) {
        const int src = i - (4 - ndim);
        bt.shape[i] = src >= 0 ? shape[src] : 1;
        bt.s0[i] = src >= 0 ? os[src] : 0;
        bt.s1[i] = src >= 0 ? ds[src] : 0;
        bt.s2[i] = src >= 0 ? ts[src] : 0;
        bt.s3[i] = 0;
        if (bt.shape[i] < 0) return DRT_ERR_BAD_EXTENT;
        n *= bt.shape[i];
    }
    if (n == 0) return DRT_OK;
    if (o == nullptr || d == nullptr || tri == nullptr || t_out == nullptr || hit_out == nullptr)
        return DRT_ERR_NULL_POINTER;
    const int threads = 256;
    const int64_t blocks = (n + threads - 1) / threads;
    const unsigned grid = unsigned(blocks < int64_t(drt::device_sm_count()) * 16 ? blocks : int64_t(drt::device_sm_count()) * 16);
    mt_elementwise_kernel<<<grid, threads, 0, static_cast<cudaStream_t>(stream)>>>(
        n, bt, o, d, tri, epsilon, t_out, hit_out);
    DRT_CHECK_CUDA(cudaGetLastError());
    return DRT_OK;
}

int drt_ray_intersect_any_triangle(drt_stream_t stream, int64_t R, const float *o, const float *d,
                                   const void *pack, int64_t T, float epsilon, float hit_tol,
                                   uint8_t *out, int64_t *tests_done) {
    if (R < 0 || T < 0) return DRT_ERR_BAD_EXTENT;
    if (R == 0) return DRT_OK;
    if (out == nullptr) return DRT_ERR_NULL_POINTER;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (T == 0) {  // _utils.py:1441-1450
        DRT_CHECK_CUDA(cudaMemsetAsync(out, 0, size_t(R), s));
        return DRT_OK;
    }
    if (o == nullptr || d == nullptr || pack == nullptr) return DRT_ERR_NULL_POINTER;
    CoreParams p{};
    p.pack = static_cast<const Tri48 *>(pack);
    p.num_tiles = int(padded_triangles(T) / kTile);
    p.num_units_dev = nullptr;
    p.eps = epsilon;
    p.thr = 1.0f - hit_tol;
    p.batch_size = 0;
    p.num_triangles = T;
    p.tests_done = tests_done;
    DRT_DISPATCH_RPW(R, {
        p.num_units = (R + RPW - 1) / RPW;
        FlatRays<RPW> src{o, d, R};
        AnySink<RPW> sink{out};
        DRT_CHECK_CUDA((launch_intersect<RPW, MODE_ANY>(s, p, src, sink, p.num_units)));
    })
    return DRT_OK;
}

size_t drt_any_hit_workspace_bytes(int64_t T) {
    if (T < 0) return 0;
    const int64_t records = padded_triangles(T);
    return ((cull_layout(records).total + 255) & ~size_t(255)) + drt_mesh_pack_sort_workspace_bytes(T) + 256;
}

int drt_ray_intersect_any_triangle_culled(drt_stream_t stream, int64_t R, const float *o, const float *d,
                                          const void *pack, int64_t T, float epsilon, float hit_tol,
                                          void *workspace, size_t workspace_bytes, uint8_t *out,
                                          int64_t *tests_done) {
    if (R < 0 || T < 0) return DRT_ERR_BAD_EXTENT;
    const int64_t records = padded_triangles(T);
    const float thr = 1.0f - hit_tol;
    // small meshes, or parameters outside the range the cull's proof covers: the plain all-pairs engine
    if (R == 0 || T == 0 || workspace == nullptr || records <= int64_t(kCullHead) * kTile ||
        !(epsilon >= 1.17549435e-38f) || !(thr > 0.0f && thr <= 1.0f))
        return drt_ray_intersect_any_triangle(stream, R, o, d, pack, T, epsilon, hit_tol, out, tests_done);
    if (out == nullptr || o == nullptr || d == nullptr || pack == nullptr) return DRT_ERR_NULL_POINTER;
    if (workspace_bytes < drt_any_hit_workspace_bytes(T)) return DRT_ERR_WORKSPACE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    const CullLayout l = cull_layout(records);
    const size_t sort_off = (l.total + 255) & ~size_t(255);
    const size_t sort_bytes = drt_mesh_pack_sort_workspace_bytes(T);
    unsigned long long *cursor = reinterpret_cast<unsigned long long *>(ws + sort_off + sort_bytes);
    int rc = cull_build(s, records, static_cast<const Tri48 *>(pack), ws, l, ws + sort_off, sort_bytes);
    if (rc != DRT_OK) return rc;
    DRT_CHECK_CUDA(cudaMemsetAsync(cursor, 0, 8, s));
    DRT_CHECK_CUDA(cudaMemsetAsync(out, 0, size_t(R), s));
    const int64_t wblocks = (R + kWalkWarps - 1) / kWalkWarps;
    const int64_t wres = int64_t(device_sm_count()) * DRT_WALK_CTAS;
    path_walk_kernel<1, true><<<unsigned(wblocks < wres ? wblocks : wres), kWalkWarps * 32, 0, s>>>(
        reinterpret_cast<const Tri48 *>(ws + l.pack), reinterpret_cast<const CullNode *>(ws + l.walk), l.levels,
        static_cast<const Tri48 *>(pack), R, nullptr, o, d, nullptr, epsilon, thr, out, cursor, tests_done);
    DRT_CHECK_CUDA(cudaGetLastError());
    return DRT_OK;
}

int drt_first_triangle_hit_by_ray(drt_stream_t stream, int64_t R, const float *o, const float *d,
                                  const void *pack, int64_t T, float epsilon, int64_t batch_size,
                                  int32_t *out_index, float *out_t, int64_t *tests_done) {
    if (R < 0 || T < 0) return DRT_ERR_BAD_EXTENT;
    if (R == 0) return DRT_OK;
    if (out_index == nullptr || out_t == nullptr) return DRT_ERR_NULL_POINTER;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (T == 0) {  // _utils.py:1848-1857
        fill_first_miss_kernel<<<unsigned((R + 255) / 256), 256, 0, s>>>(R, out_index, out_t);
        DRT_CHECK_CUDA(cudaGetLastError());
        return DRT_OK;
    }
    if (o == nullptr || d == nullptr || pack == nullptr) return DRT_ERR_NULL_POINTER;
    CoreParams p{};
    p.pack = static_cast<const Tri48 *>(pack);
    p.num_tiles = int(padded_triangles(T) / kTile);
    p.eps = epsilon;
    p.thr = 0.f;
    p.batch_size = batch_size;
    p.num_triangles = T;
    p.tests_done = tests_done;
    DRT_DISPATCH_RPW(R, {
        p.num_units = (R + RPW - 1) / RPW;
        FlatRays<RPW> src{o, d, R};
        FirstSink<RPW> sink{out_index, out_t};
        DRT_CHECK_CUDA((launch_intersect<RPW, MODE_FIRST>(s, p, src, sink, p.num_units)));
    })
    return DRT_OK;
}

int drt_first_triangle_hit_by_ray_culled(drt_stream_t stream, int64_t R, const float *o, const float *d,
                                         const void *pack, int64_t T, float epsilon, int64_t batch_size,
                                         void *workspace, size_t workspace_bytes, int32_t *out_index,
                                         float *out_t, int64_t *tests_done) {
    if (R < 0 || T < 0) return DRT_ERR_BAD_EXTENT;
    const int64_t records = padded_triangles(T);
    // small meshes, or an epsilon outside the range the cull's proof covers: the plain all-pairs engine
    if (R == 0 || T == 0 || workspace == nullptr || records <= int64_t(kCullHead) * kTile ||
        !(epsilon >= 1.17549435e-38f))
        return drt_first_triangle_hit_by_ray(stream, R, o, d, pack, T, epsilon, batch_size, out_index, out_t,
                                             tests_done);
    if (out_index == nullptr || out_t == nullptr || o == nullptr || d == nullptr || pack == nullptr)
        return DRT_ERR_NULL_POINTER;
    if (workspace_bytes < drt_any_hit_workspace_bytes(T)) return DRT_ERR_WORKSPACE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    const CullLayout l = cull_layout(records);
    const size_t sort_off = (l.total + 255) & ~size_t(255);
    const size_t sort_bytes = drt_mesh_pack_sort_workspace_bytes(T);
    unsigned long long *cursor = reinterpret_cast<unsigned long long *>(ws + sort_off + sort_bytes);
    int rc = cull_build(s, records, static_cast<const Tri48 *>(pack), ws, l, ws + sort_off, sort_bytes);
    if (rc != DRT_OK) return rc;
    DRT_CHECK_CUDA(cudaMemsetAsync(cursor, 0, 8, s));
    const int64_t wblocks = (R + kWalkWarps - 1) / kWalkWarps;
    const int64_t wres = int64_t(device_sm_count()) * DRT_WALK_CTAS;
    ray_first_walk_kernel<<<unsigned(wblocks < wres ? wblocks : wres), kWalkWarps * 32, 0, s>>>(
        reinterpret_cast<const Tri48 *>(ws + l.pack),
        reinterpret_cast<const uint32_t *>(ws + sort_off + sort_indices_offset(records)),
        reinterpret_cast<const CullNode *>(ws + l.walk), l.levels, R, o, d, epsilon, batch_size, T, out_index, out_t,
        cursor, tests_done);
    DRT_CHECK_CUDA(cudaGetLastError());
    return DRT_OK;
}

int drt_triangles_visible_from_vertex(drt_stream_t stream, int64_t B, int64_t n_rays,
                                      const float *vertices, const float *dirs, const void *pack,
                                      int64_t T, float epsilon, uint8_t *out, int64_t *tests_done) {
    if (B < 0 || n_rays < 0 || T < 0) return DRT_ERR_BAD_EXTENT;
    if (B == 0 || T == 0) return DRT_OK;
    if (out == nullptr) return DRT_ERR_NULL_POINTER;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DRT_CHECK_CUDA(cudaMemsetAsync(out, 0, size_t(B) * size_t(T), s));  // _mesh.py:386
    const int64_t R = B * n_rays;
    if (R == 0) return DRT_OK;
    if (vertices == nullptr || dirs == nullptr || pack == nullptr) return DRT_ERR_NULL_POINTER;
    CoreParams p{};
    p.pack = static_cast<const Tri48 *>(pack);
    p.num_tiles = int(padded_triangles(T) / kTile);
    p.eps = epsilon;
    p.batch_size = 0;  // the reference calls first-hit with batch_size=None here (_utils.py:1717)
    p.num_triangles = T;
    p.tests_done = tests_done;
    DRT_DISPATCH_RPW(R, {
        p.num_units = (R + RPW - 1) / RPW;
        VertexRays<RPW> src{vertices, dirs, R, n_rays};
        VisibleSink<RPW> sink{out, n_rays, T};
        DRT_CHECK_CUDA((launch_intersect<RPW, MODE_FIRST>(s, p, src, sink, p.num_units)));
    })
    return DRT_OK;
}

}  // extern "C"
