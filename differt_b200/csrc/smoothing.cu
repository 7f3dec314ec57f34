// N4 (forward): the smoothed variants of the intersection primitives.
// Reference: differt/src/differt/utils.py:70-89 (smoothing_function = sigmoid(x * alpha)),
// differt/src/differt/geometry/_utils.py:1279-1318 (ray_intersect_triangle), :1465-1476
// (ray_intersect_any_triangle), _solver_image_method.py:448-454 (same side of mirrors).
// Comparisons become sigmoids, AND becomes min, the OR over triangles becomes a sum clipped at 1.
// Outputs are floats in [0, 1]; with the transcendental involved parity is to tolerance (1e-5), not
// bit-exact.  The smoothed trace (_solvers.py:599-713) and the gradients are not built.
#include "common.cuh"

namespace drt {

__device__ __forceinline__ float smooth(float x, float alpha) {  // jax.nn.sigmoid(x * alpha)
    return __fdiv_rn(1.0f, 1.0f + expf(-(x * alpha)));
}

// jnp.min / jnp.minimum propagate NaN; fminf would drop it
__device__ __forceinline__ float nanmin(float a, float b) { return (a != a || b != b) ? CUDART_NAN_F : fminf(a, b); }

// _utils.py:1263-1322 with smoothing_factor; returns the smoothed hit, writes t
__device__ __forceinline__ float mt_smooth(const float3 o, const float3 d, const Tri &tr, const float eps,
                                           const float alpha, float &t) {
    const float3 h = cross3(d, tr.e2);
    float a = dot3(h, tr.e1);
    a = (a == 0.0f) ? CUDART_INF_F : a;
    float hit = smooth(fabsf(a) - eps, alpha);
    const float f = __frcp_rn(a);
    const float3 s = sub3(o, tr.v0);
    const float u = f * dot3(s, h);
    hit = nanmin(nanmin(hit, smooth(u - 0.0f, alpha)), nanmin(smooth(1.0f - u, alpha), 1.0f));
    const float3 q = cross3(s, tr.e1);
    const float v = f * dot3(q, d);
    hit = nanmin(nanmin(hit, smooth(v - 0.0f, alpha)), nanmin(smooth(1.0f - (u + v), alpha), 1.0f));
    t = f * dot3(q, tr.e2);
    return nanmin(hit, smooth(t - eps, alpha));
}

__global__ void mt_smooth_elementwise_kernel(int64_t n, Batch4 bt, const float *__restrict__ o,
                                             const float *__restrict__ d, const float *__restrict__ tri,
                                             float eps, float alpha, float *__restrict__ t_out,
                                             float *__restrict__ hit_out) {
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
        hit_out[i] = mt_smooth(ld3(o + oo), ld3(d + od), tr, eps, alpha, t);
        t_out[i] = t;
    }
}

// one warp per ray: sum over the (active) triangles of min(hit, sigmoid((thr - t) alpha)), clipped at 1
__global__ void __launch_bounds__(256)
any_smooth_kernel(int64_t R, int64_t T, const float *__restrict__ o, const float *__restrict__ d,
                  const Tri48 *__restrict__ pack, float eps, float thr, float alpha,
                  float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t ray = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
    if (ray >= R) return;
    const float3 oo = ld3(o + 3 * ray), dd = ld3(d + 3 * ray);
    float acc = 0.0f;
    for (int64_t j = lane; j < T; j += 32) {
        const float4 a = pack[j].a, b = pack[j].b, c = pack[j].c;
        if (a.x != a.x && a.y != a.y && a.z != a.z && a.w == 0.0f) continue;  // never-hit record: inactive
        float t;
        const float hit = mt_smooth(oo, dd, unpack(a, b, c), eps, alpha, t);
        acc += nanmin(hit, smooth(thr - t, alpha));
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(kFull, acc, off);
    if (lane == 0) out[ray] = nanmin(acc, 1.0f);  // (left + right).clip(max=1), _utils.py:1474-1476
}

__global__ void __launch_bounds__(256)
same_side_smooth_kernel(int64_t n, int K, Batch4 bt, const float *__restrict__ v, const float *__restrict__ mv,
                        const float *__restrict__ mn, float alpha, float *__restrict__ out) {
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n * K; i += stride) {
        const int64_t b = i / K;
        const int j = int(i % K);
        int64_t ov, om, on, unused;
        bt.offsets(b, ov, om, on, unused);
        const float3 m = ld3(mv + om + 3 * j), nn = ld3(mn + on + 3 * j);
        const float dp = dot3(sub3(ld3(v + ov + 3 * j), m), nn);
        const float dn = dot3(sub3(ld3(v + ov + 3 * (j + 2)), m), nn);
        // jnp.sign: -1, 0, +1, NaN for NaN
        const float sp = dp != dp ? dp : float(dp > 0.0f) - float(dp < 0.0f);
        const float sn = dn != dn ? dn : float(dn > 0.0f) - float(dn < 0.0f);
        out[i] = smooth(sp * sn, alpha);
    }
}

static int fill_batch3(int32_t ndim, const int64_t *shape, const int64_t *s0, const int64_t *s1,
                       const int64_t *s2, Batch4 &bt, int64_t &n) {
    if (ndim < 0 || ndim > DRT_MAX_BATCH_DIMS) return DRT_ERR_UNSUPPORTED;
    if (ndim > 0 && (!shape || !s0 || !s1 || !s2)) return DRT_ERR_NULL_POINTER;
    n = 1;
    for (int i = 0; i < 4; ++i) {
        const int src = i - (4 - ndim);
        bt.shape[i] = src >= 0 ? shape[src] : 1;
        bt.s0[i] = src >= 0 ? s0[src] : 0;
        bt.s1[i] = src >= 0 ? s1[src] : 0;
        bt.s2[i] = src >= 0 ? s2[src] : 0;
        bt.s3[i] = 0;
        if (bt.shape[i] < 0) return DRT_ERR_BAD_EXTENT;
        n *= bt.shape[i];
    }
    return DRT_OK;
}

}  // namespace drt

using namespace drt;

extern "C" {

int drt_ray_intersect_triangle_smooth(drt_stream_t stream, int32_t ndim, const int64_t *shape,
                                      const float *o, const int64_t *os, const float *d, const int64_t *ds,
                                      const float *tri, const int64_t *ts, float epsilon,
                                      float smoothing_factor, float *t_out, float *hit_out) {
    Batch4 bt;
    int64_t n;
    const int rc = fill_batch3(ndim, shape, os, ds, ts, bt, n);
    if (rc != DRT_OK) return rc;
    if (n == 0) return DRT_OK;
    if (!o || !d || !tri || !t_out || !hit_out) return DRT_ERR_NULL_POINTER;
    const int64_t blocks = (n + 255) / 256;
    mt_smooth_elementwise_kernel<<<unsigned(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0,
                                   static_cast<cudaStream_t>(stream)>>>(n, bt, o, d, tri, epsilon,
                                                                        smoothing_factor, t_out, hit_out);
    return cudaGetLastError() == cudaSuccess ? DRT_OK : DRT_ERR_CUDA;
}

int drt_ray_intersect_any_triangle_smooth(drt_stream_t stream, int64_t R, const float *o, const float *d,
                                          const void *pack, int64_t T, float epsilon, float hit_tol,
                                          float smoothing_factor, float *out) {
    if (R < 0 || T < 0) return DRT_ERR_BAD_EXTENT;
    if (R == 0) return DRT_OK;
    if (!out) return DRT_ERR_NULL_POINTER;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (T == 0) return cudaMemsetAsync(out, 0, size_t(R) * sizeof(float), s) == cudaSuccess ? DRT_OK : DRT_ERR_CUDA;
    if (!o || !d || !pack) return DRT_ERR_NULL_POINTER;
    any_smooth_kernel<<<unsigned((R * 32 + 255) / 256), 256, 0, s>>>(R, T, o, d, static_cast<const Tri48 *>(pack),
                                                                     epsilon, 1.0f - hit_tol, smoothing_factor, out);
    return cudaGetLastError() == cudaSuccess ? DRT_OK : DRT_ERR_CUDA;
}

int drt_consecutive_vertices_are_on_same_side_of_mirror_smooth(
    drt_stream_t stream, int32_t ndim, const int64_t *shape, int32_t order, const float *vertices,
    const int64_t *vs, const float *mv, const int64_t *ms, const float *mn, const int64_t *ns,
    float smoothing_factor, float *out) {
    if (order < 0) return DRT_ERR_BAD_EXTENT;
    Batch4 bt;
    int64_t n;
    const int rc = fill_batch3(ndim, shape, vs, ms, ns, bt, n);
    if (rc != DRT_OK) return rc;
    if (n == 0 || order == 0) return DRT_OK;
    if (!vertices || !mv || !mn || !out) return DRT_ERR_NULL_POINTER;
    const int64_t blocks = (n * order + 255) / 256;
    same_side_smooth_kernel<<<unsigned(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0,
                              static_cast<cudaStream_t>(stream)>>>(n, order, bt, vertices, mv, mn,
                                                                   smoothing_factor, out);
    return cudaGetLastError() == cudaSuccess ? DRT_OK : DRT_ERR_CUDA;
}

}  // extern "C"
