// K5 image_method and K5b its reverse mode over a broadcast batch.
// Reference: differt/src/differt/geometry/_solver_image_method.py:206-363 (jnp.vectorize over the
// batch of two lax.scans); gradients there come from JAX autodiff of the scans.
// One thread per batch element: the k mirrors, images and path points live in registers
// (k <= DRT_MAX_ORDER) — the kernel is a pure HBM stream (36k bytes per element).
#include "image_core.cuh"

namespace drt {

constexpr int kMaxGenericOrder = 64;

template <int K>
__global__ void __launch_bounds__(256)
image_method_kernel(int64_t n, Batch4 bt, const float *__restrict__ from, const float *__restrict__ to,
                    const float *__restrict__ mv, const float *__restrict__ mn,
                    float *__restrict__ out) {
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += stride) {
        int64_t of, ot, ov, on;
        bt.offsets(i, of, ot, ov, on);
        float3 full[K + 2], a[K], b[K];
        full[0] = ld3(from + of);
        full[K + 1] = ld3(to + ot);
#pragma unroll
        for (int j = 0; j < K; ++j) {
            a[j] = ld3(mv + ov + 3 * j);
            b[j] = ld3(mn + on + 3 * j);
        }
        image_method_path<K>(full, a, b);
#pragma unroll
        for (int j = 0; j < K; ++j) st3(out + (i * K + j) * 3, full[j + 1]);
    }
}

// runtime-order fallback (8 < k <= 64): same arithmetic, arrays in local memory
__global__ void __launch_bounds__(128)
image_method_generic_kernel(int64_t n, int K, Batch4 bt, const float *__restrict__ from,
                            const float *__restrict__ to, const float *__restrict__ mv,
                            const float *__restrict__ mn, float *__restrict__ out) {
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += stride) {
        int64_t of, ot, ov, on;
        bt.offsets(i, of, ot, ov, on);
        float3 img[kMaxGenericOrder];
        float3 prev = ld3(from + of);
        for (int j = 0; j < K; ++j) {
            prev = mirror_image(prev, ld3(mv + ov + 3 * j), ld3(mn + on + 3 * j));
            img[j] = prev;
        }
        prev = ld3(to + ot);
        for (int j = K - 1; j >= 0; --j) {
            prev = back_step(prev, img[j], ld3(mv + ov + 3 * j), ld3(mn + on + 3 * j), nullptr);
            st3(out + (i * K + j) * 3, prev);
        }
    }
}

template <int K>
__global__ void __launch_bounds__(256)
image_method_vjp_kernel(int64_t n, Batch4 bt, const float *__restrict__ from,
                        const float *__restrict__ to, const float *__restrict__ mv,
                        const float *__restrict__ mn, const float *__restrict__ g_paths,
                        float *__restrict__ g_from, float *__restrict__ g_to,
                        float *__restrict__ g_mv, float *__restrict__ g_mn) {
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += stride) {
        int64_t of, ot, ov, on;
        bt.offsets(i, of, ot, ov, on);
        float3 a[K], b[K], g[K], ga[K], gb[K];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            a[j] = ld3(mv + ov + 3 * j);
            b[j] = ld3(mn + on + 3 * j);
            g[j] = ld3(g_paths + (i * K + j) * 3);
            ga[j] = make_float3(0.f, 0.f, 0.f);
            gb[j] = make_float3(0.f, 0.f, 0.f);
        }
        float3 gf = make_float3(0.f, 0.f, 0.f), gt = make_float3(0.f, 0.f, 0.f);
        image_method_reverse<K>(ld3(from + of), ld3(to + ot), a, b, g, gf, gt, ga, gb);
        st3(g_from + 3 * i, gf);
        st3(g_to + 3 * i, gt);
#pragma unroll
        for (int j = 0; j < K; ++j) {
            st3(g_mv + (i * K + j) * 3, ga[j]);
            st3(g_mn + (i * K + j) * 3, gb[j]);
        }
    }
}


// a5 / a6 / a8 as stand-alone element-wise kernels (the reference exports them as public functions)
__global__ void __launch_bounds__(256)
image_of_vertex_kernel(int64_t n, Batch4 bt, const float *__restrict__ p, const float *__restrict__ m,
                       const float *__restrict__ nrm, float *__restrict__ out) {
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += stride) {
        int64_t op, om, on, unused;
        bt.offsets(i, op, om, on, unused);
        st3(out + 3 * i, mirror_image(ld3(p + op), ld3(m + om), ld3(nrm + on)));
    }
}

__global__ void __launch_bounds__(256)
ray_plane_kernel(int64_t n, Batch4 bt, const float *__restrict__ o, const float *__restrict__ d,
                 const float *__restrict__ pv, const float *__restrict__ pn, float *__restrict__ out) {
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += stride) {
        int64_t oo, od, ov, on;
        bt.offsets(i, oo, od, ov, on);
        // _solver_image_method.py:110-135 (no inf guard here: that lives in the scan body)
        const float3 ro = ld3(o + oo), u = ld3(d + od), nn = ld3(pn + on);
        const float3 w = sub3(ld3(pv + ov), ro);
        float un = dot3(u, nn);
        const float vn = dot3(w, nn);
        const bool par = (un == 0.0f);
        un = par ? 1.0f : un;
        const float t = __fdiv_rn(vn, un);
        float3 r = make_float3(ro.x + u.x * t, ro.y + u.y * t, ro.z + u.z * t);
        if (par && vn != 0.0f) r = make_float3(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F);
        st3(out + 3 * i, r);
    }
}

// vertices [n, k+2, 3] (strided batch), mirrors [n, k, 3] → out [n, k] u8
__global__ void __launch_bounds__(256)
same_side_kernel(int64_t n, int K, Batch4 bt, const float *__restrict__ v, const float *__restrict__ mv,
                 const float *__restrict__ mn, uint8_t *__restrict__ out) {
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n * K; i += stride) {
        const int64_t b = i / K;
        const int j = int(i % K);
        int64_t ov, om, on, unused;
        bt.offsets(b, ov, om, on, unused);
        const float3 m = ld3(mv + om + 3 * j), nn = ld3(mn + on + 3 * j);
        const float dp = dot3(sub3(ld3(v + ov + 3 * j), m), nn);
        const float dn = dot3(sub3(ld3(v + ov + 3 * (j + 2)), m), nn);
        const float sp = float(dp > 0.0f) - float(dp < 0.0f), sn = float(dn > 0.0f) - float(dn < 0.0f);
        out[i] = (sp == sn && dp == dp && dn == dn) ? 1 : 0;
    }
}

static int fill_batch(int32_t ndim, const int64_t *shape, const int64_t *s0, const int64_t *s1,
                      const int64_t *s2, const int64_t *s3, Batch4 &bt, int64_t &n) {
    if (ndim < 0 || ndim > DRT_MAX_BATCH_DIMS) return DRT_ERR_UNSUPPORTED;
    if (ndim > 0 && (!shape || !s0 || !s1 || !s2 || !s3)) return DRT_ERR_NULL_POINTER;
    n = 1;
    for (int i = 0; i < 4; ++i) {
        const int src = i - (4 - ndim);
        bt.shape[i] = src >= 0 ? shape[src] : 1;
        bt.s0[i] = src >= 0 ? s0[src] : 0;
        bt.s1[i] = src >= 0 ? s1[src] : 0;
        bt.s2[i] = src >= 0 ? s2[src] : 0;
        bt.s3[i] = src >= 0 ? s3[src] : 0;
        if (bt.shape[i] < 0) return DRT_ERR_BAD_EXTENT;
        n *= bt.shape[i];
    }
    return DRT_OK;
}

static unsigned grid_for(int64_t n, int threads) {
    const int64_t blocks = (n + threads - 1) / threads;
    return unsigned(blocks < int64_t(drt::device_sm_count()) * 16 ? blocks : int64_t(drt::device_sm_count()) * 16);
}

}  // namespace drt

using namespace drt;

extern "C" {

int drt_image_method(drt_stream_t stream, int32_t ndim, const int64_t *shape, int32_t order,
                     const float *from, const int64_t *fs, const float *to, const int64_t *ts,
                     const float *mv, const int64_t *vs, const float *mn, const int64_t *ns,
                     float *out) {
    if (order < 0) return DRT_ERR_BAD_EXTENT;
    if (order > kMaxGenericOrder) return DRT_ERR_UNSUPPORTED;
    Batch4 bt;
    int64_t n;
    const int rc = fill_batch(ndim, shape, fs, ts, vs, ns, bt, n);
    if (rc != DRT_OK) return rc;
    if (n == 0 || order == 0) return DRT_OK;  // _solver_image_method.py:349-358
    if (!from || !to || !mv || !mn || !out) return DRT_ERR_NULL_POINTER;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
#define DRT_IM_CASE(K)                                                                          \
    case K:                                                                                     \
        image_method_kernel<K><<<grid_for(n, 256), 256, 0, s>>>(n, bt, from, to, mv, mn, out);  \
        break;
    switch (order) {
        DRT_IM_CASE(1) DRT_IM_CASE(2) DRT_IM_CASE(3) DRT_IM_CASE(4)
        DRT_IM_CASE(5) DRT_IM_CASE(6) DRT_IM_CASE(7) DRT_IM_CASE(8)
        default:
            image_method_generic_kernel<<<grid_for(n, 128), 128, 0, s>>>(n, order, bt, from, to, mv,
                                                                         mn, out);
    }
#undef DRT_IM_CASE
    return cudaGetLastError() == cudaSuccess ? DRT_OK : DRT_ERR_CUDA;
}

int drt_image_method_vjp(drt_stream_t stream, int32_t ndim, const int64_t *shape, int32_t order,
                         const float *from, const int64_t *fs, const float *to, const int64_t *ts,
                         const float *mv, const int64_t *vs, const float *mn, const int64_t *ns,
                         const float *g_paths, float *g_from, float *g_to, float *g_mv,
                         float *g_mn) {
    if (order < 0) return DRT_ERR_BAD_EXTENT;
    if (order > DRT_MAX_ORDER) return DRT_ERR_UNSUPPORTED;
    Batch4 bt;
    int64_t n;
    const int rc = fill_batch(ndim, shape, fs, ts, vs, ns, bt, n);
    if (rc != DRT_OK) return rc;
    if (n == 0) return DRT_OK;
    if (!g_from || !g_to) return DRT_ERR_NULL_POINTER;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (order == 0) {
        if (cudaMemsetAsync(g_from, 0, size_t(n) * 12, s) != cudaSuccess) return DRT_ERR_CUDA;
        if (cudaMemsetAsync(g_to, 0, size_t(n) * 12, s) != cudaSuccess) return DRT_ERR_CUDA;
        return DRT_OK;
    }
    if (!from || !to || !mv || !mn || !g_paths || !g_mv || !g_mn) return DRT_ERR_NULL_POINTER;
#define DRT_IM_CASE(K)                                                                          \
    case K:                                                                                     \
        image_method_vjp_kernel<K><<<grid_for(n, 256), 256, 0, s>>>(n, bt, from, to, mv, mn,    \
                                                                    g_paths, g_from, g_to, g_mv, \
                                                                    g_mn);                       \
        break;
    switch (order) {
        DRT_IM_CASE(1) DRT_IM_CASE(2) DRT_IM_CASE(3) DRT_IM_CASE(4)
        DRT_IM_CASE(5) DRT_IM_CASE(6) DRT_IM_CASE(7) DRT_IM_CASE(8)
    }
#undef DRT_IM_CASE
    return cudaGetLastError() == cudaSuccess ? DRT_OK : DRT_ERR_CUDA;
}

int drt_image_of_vertex_with_respect_to_mirror(drt_stream_t stream, int32_t ndim, const int64_t *shape,
                                               const float *vertex, const int64_t *ps,
                                               const float *mirror_vertex, const int64_t *ms,
                                               const float *mirror_normal, const int64_t *ns,
                                               float *out) {
    Batch4 bt;
    int64_t n;
    const int rc = fill_batch(ndim, shape, ps, ms, ns, ns, bt, n);
    if (rc != DRT_OK) return rc;
    if (n == 0) return DRT_OK;
    if (!vertex || !mirror_vertex || !mirror_normal || !out) return DRT_ERR_NULL_POINTER;
    image_of_vertex_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        n, bt, vertex, mirror_vertex, mirror_normal, out);
    return cudaGetLastError() == cudaSuccess ? DRT_OK : DRT_ERR_CUDA;
}

int drt_intersection_of_ray_with_plane(drt_stream_t stream, int32_t ndim, const int64_t *shape,
                                       const float *o, const int64_t *os, const float *d,
                                       const int64_t *ds, const float *pv, const int64_t *vs,
                                       const float *pn, const int64_t *ns, float *out) {
    Batch4 bt;
    int64_t n;
    const int rc = fill_batch(ndim, shape, os, ds, vs, ns, bt, n);
    if (rc != DRT_OK) return rc;
    if (n == 0) return DRT_OK;
    if (!o || !d || !pv || !pn || !out) return DRT_ERR_NULL_POINTER;
    ray_plane_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(n, bt, o, d, pv,
                                                                                     pn, out);
    return cudaGetLastError() == cudaSuccess ? DRT_OK : DRT_ERR_CUDA;
}

int drt_consecutive_vertices_are_on_same_side_of_mirror(drt_stream_t stream, int32_t ndim,
                                                        const int64_t *shape, int32_t order,
                                                        const float *vertices, const int64_t *vs,
                                                        const float *mv, const int64_t *ms,
                                                        const float *mn, const int64_t *ns,
                                                        uint8_t *out) {
    if (order < 0) return DRT_ERR_BAD_EXTENT;
    Batch4 bt;
    int64_t n;
    const int rc = fill_batch(ndim, shape, vs, ms, ns, ns, bt, n);
    if (rc != DRT_OK) return rc;
    if (n == 0 || order == 0) return DRT_OK;
    if (!vertices || !mv || !mn || !out) return DRT_ERR_NULL_POINTER;
    same_side_kernel<<<grid_for(n * order, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        n, order, bt, vertices, mv, mn, out);
    return cudaGetLastError() == cudaSuccess ? DRT_OK : DRT_ERR_CUDA;
}

}  // extern "C"
