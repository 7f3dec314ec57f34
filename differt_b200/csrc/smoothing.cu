// N4 (forward): the smoothed variants of the intersection primitives.
// Reference: differt/src/differt/utils.py:70-89 (smoothing_function = sigmoid(x * alpha)),
// differt/src/differt/geometry/_utils.py:1279-1318 (ray_intersect_triangle), :1465-1476
// (ray_intersect_any_triangle), _solver_image_method.py:448-454 (same side of mirrors).
// Comparisons become sigmoids, AND becomes min, the OR over triangles becomes a sum clipped at 1.
// Outputs are floats in [0, 1]; with the transcendental involved parity is to tolerance (1e-5), not
// bit-exact.  The smoothed trace (_solvers.py:599-713) is the last entry point of this file; the
// gradients of the relaxed outputs are not built.
#include "common.cuh"
#include "image_core.cuh"

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

// jnp.max propagates NaN as well
__device__ __forceinline__ float nanmax(float a, float b) { return (a != a || b != b) ? CUDART_NAN_F : fmaxf(a, b); }

struct SmoothTraceArgs {
    const Tri48 *pack;        // geometry of every triangle (mask NOT applied): mirrors + inside test
    const Tri48 *pack_active; // mask applied: blockage
    const uint8_t *tri_mask;  // nullable
    const float *tx, *rx;
    const int32_t *cand;
    int64_t T, ntx, nrx, C, P;
    float eps, thr, min_len, alpha;
    float *out_vertices;
    int32_t *out_objects;
    float *out_mask;
    uint8_t *flags;           // [P] bit 0: path is finite, bit 1: every mirror of the candidate is active
};

// Steps 2 - 3.2, 3.4, 3.5 of the relaxed _trace_path_candidates (_solvers.py:576-660, 684-703), one
// thread per path: image method (same arithmetic as the hard trace), then min over the interactions
// of the smoothed inside test (max over the two triangles of a quad) and of the smoothed same-side
// test, max over the segments of sigmoid((min_len - |d|^2) alpha), and the hard finiteness test.
// Writes the dense vertices / objects and min(inside, same, 1 - too_small, finite) into out_mask; the
// blockage kernel below folds 1 - blocked in.
template <int K, bool QUADS>
__global__ void __launch_bounds__(128) trace_smooth_stage_kernel(const SmoothTraceArgs a) {
    constexpr int KK = K > 0 ? K : 1;
    constexpr int NT = QUADS ? 2 : 1;
    const int64_t p = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (p >= a.P) return;
    const int64_t c = p % a.C, pair = p / a.C;
    const int64_t irx = pair % a.nrx, itx = pair / a.nrx;

    float3 mv[KK], mn[KK];
    Tri tri[KK][NT];
    int32_t ci[KK];
    bool active = true;
#pragma unroll
    for (int i = 0; i < K; ++i) {
        int32_t t = a.cand[c * K + i];
        ci[i] = t;
        t = min(max(t, 0), int32_t(a.T - (QUADS ? 2 : 1)));
#pragma unroll
        for (int q = 0; q < NT; ++q) {
            const float4 ta = a.pack[t + q].a, tb = a.pack[t + q].b, tc = a.pack[t + q].c;
            tri[i][q] = unpack(ta, tb, tc);
            if (q == 0) {
                mv[i] = make_float3(ta.x, ta.y, ta.z);
                mn[i] = make_float3(tc.y, tc.z, tc.w);
            }
            if (a.tri_mask != nullptr) active = active && a.tri_mask[t + q] != 0;
        }
    }
    float3 full[K + 2];
    full[0] = ld3(a.tx + 3 * itx);
    full[K + 1] = ld3(a.rx + 3 * irx);
    image_method_path<K>(full, mv, mn);

    float inside = 1.0f, same = 1.0f, small = 0.0f;
    bool finite = true;
#pragma unroll
    for (int i = 0; i <= K; ++i) {
        const float3 o = full[i];
        const float3 d = sub3(full[i + 1], full[i]);
        small = nanmax(small, smooth(a.min_len - dot3(d, d), a.alpha));
        if (i < K) {
            float tt;
            float hit = mt_smooth(o, d, tri[i][0], a.eps, a.alpha, tt);
            if (QUADS) hit = nanmax(nanmax(hit, mt_smooth(o, d, tri[i][NT - 1], a.eps, a.alpha, tt)), 0.0f);
            inside = nanmin(inside, hit);
            const float dp = dot3(sub3(full[i], mv[i]), mn[i]);
            const float dn = dot3(sub3(full[i + 2], mv[i]), mn[i]);
            const float sp = dp != dp ? dp : float(dp > 0.0f) - float(dp < 0.0f);
            const float sn = dn != dn ? dn : float(dn > 0.0f) - float(dn < 0.0f);
            same = nanmin(same, smooth(sp * sn, a.alpha));
        }
    }
#pragma unroll
    for (int i = 0; i < K + 2; ++i) finite = finite && finite3(full[i]);

    float *ov = a.out_vertices + p * (K + 2) * 3;
#pragma unroll
    for (int i = 0; i < K + 2; ++i) st3(ov + 3 * i, finite ? full[i] : make_float3(0.f, 0.f, 0.f));
    int32_t *oo = a.out_objects + p * (K + 2);
    oo[0] = int32_t(itx);
#pragma unroll
    for (int i = 0; i < K; ++i) oo[i + 1] = ci[i];
    oo[K + 1] = int32_t(irx);
    a.out_mask[p] = nanmin(nanmin(inside, same), nanmin(1.0f - small, finite ? 1.0f : 0.0f));
    a.flags[p] = uint8_t((finite ? 1 : 0) | (active ? 2 : 0));
}

// Step 3.3 (_solvers.py:662-672) + the final min and active_rays product (:705-713), one warp per
// path: every segment of the path against every active triangle, sum per segment of
// min(hit, sigmoid((1 - hit_tol - t) alpha)) clipped at 1, max over the segments.  Non-finite paths
// keep what the stage kernel wrote (0 or NaN, "invalid" either way): their rays were never finite.
template <int NSEG>
__global__ void __launch_bounds__(256) trace_smooth_blocked_kernel(const SmoothTraceArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t p = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
    if (p >= a.P) return;
    const uint8_t fl = a.flags[p];
    float m = a.out_mask[p];
    if (fl & 1) {
        const float *v = a.out_vertices + p * (NSEG + 1) * 3;
        float3 o[NSEG], d[NSEG];
        float acc[NSEG];
#pragma unroll
        for (int s = 0; s < NSEG; ++s) {
            o[s] = ld3(v + 3 * s);
            d[s] = sub3(ld3(v + 3 * s + 3), o[s]);
            acc[s] = 0.0f;
        }
        for (int64_t j = lane; j < a.T; j += 32) {
            const float4 ra = a.pack_active[j].a, rb = a.pack_active[j].b, rc = a.pack_active[j].c;
            if (ra.x != ra.x && ra.y != ra.y && ra.z != ra.z && ra.w == 0.0f) continue;  // inactive
            const Tri tr = unpack(ra, rb, rc);
#pragma unroll
            for (int s = 0; s < NSEG; ++s) {
                float t;
                const float hit = mt_smooth(o[s], d[s], tr, a.eps, a.alpha, t);
                acc[s] += nanmin(hit, smooth(a.thr - t, a.alpha));
            }
        }
        float blocked = 0.0f;
#pragma unroll
        for (int s = 0; s < NSEG; ++s) {
            float x = acc[s];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(kFull, x, off);
            blocked = nanmax(blocked, nanmin(x, 1.0f));
        }
        m = nanmin(m, 1.0f - blocked);
    }
    if (lane == 0) a.out_mask[p] = (a.tri_mask != nullptr) ? m * ((fl & 2) ? 1.0f : 0.0f) : m;
}

template <int K>
static int launch_trace_smooth(const SmoothTraceArgs &a, bool quads, cudaStream_t s) {
    const unsigned blocks = unsigned((a.P + 127) / 128);
    if (quads)
        trace_smooth_stage_kernel<K, true><<<blocks, 128, 0, s>>>(a);
    else
        trace_smooth_stage_kernel<K, false><<<blocks, 128, 0, s>>>(a);
    if (cudaGetLastError() != cudaSuccess) return DRT_ERR_CUDA;
    // an empty mesh blocks nothing, but the active_rays product still applies (trivially: K == 0)
    trace_smooth_blocked_kernel<K + 1><<<unsigned((a.P * 32 + 255) / 256), 256, 0, s>>>(a);
    return cudaGetLastError() == cudaSuccess ? DRT_OK : DRT_ERR_CUDA;
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

size_t drt_trace_smooth_workspace_bytes(int64_t T, int64_t ntx, int64_t nrx, int64_t C) {
    if (T < 0 || ntx < 0 || nrx < 0 || C < 0) return 0;
    const size_t flags = (size_t(ntx) * size_t(nrx) * size_t(C) + 255) & ~size_t(255);
    return 2 * drt_mesh_pack_bytes(T) + flags + 256;
}

int drt_trace_path_candidates_smooth(drt_stream_t stream, int64_t V, int64_t T, const float *vertices,
                                     const int32_t *triangles, const uint8_t *triangle_mask,
                                     int32_t assume_quads, int64_t ntx, const float *tx, int64_t nrx,
                                     const float *rx, int64_t C, int32_t order, const int32_t *cand,
                                     float epsilon, float hit_tol, float min_len, float smoothing_factor,
                                     void *workspace, size_t workspace_bytes, float *out_vertices,
                                     int32_t *out_objects, float *out_mask) {
    if (V < 0 || T < 0 || ntx < 0 || nrx < 0 || C < 0 || order < 0) return DRT_ERR_BAD_EXTENT;
    if (order > DRT_MAX_ORDER) return DRT_ERR_UNSUPPORTED;
    const int64_t P = ntx * nrx * C;
    if (P >= (int64_t(1) << 32)) return DRT_ERR_BAD_EXTENT;
    if (P == 0) return DRT_OK;
    if (!tx || !rx || !out_vertices || !out_objects || !out_mask || (order > 0 && !cand)) return DRT_ERR_NULL_POINTER;
    if (order > 0 && T < (assume_quads ? 2 : 1)) return DRT_ERR_BAD_EXTENT;  // candidates index triangles
    if (T > 0 && (!vertices || !triangles)) return DRT_ERR_NULL_POINTER;
    if (!workspace || workspace_bytes < drt_trace_smooth_workspace_bytes(T, ntx, nrx, C)) return DRT_ERR_WORKSPACE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    char *ws = static_cast<char *>(workspace);
    const size_t pb = drt_mesh_pack_bytes(T);
    SmoothTraceArgs a{};
    a.pack = reinterpret_cast<const Tri48 *>(ws);
    a.pack_active = a.pack;
    a.flags = reinterpret_cast<uint8_t *>(ws + 2 * pb);
    if (T > 0) {
        int rc = drt_mesh_pack(stream, V, T, vertices, triangles, nullptr, ws);
        if (rc != DRT_OK) return rc;
        if (triangle_mask) {
            rc = drt_mesh_pack(stream, V, T, vertices, triangles, triangle_mask, ws + pb);
            if (rc != DRT_OK) return rc;
            a.pack_active = reinterpret_cast<const Tri48 *>(ws + pb);
        }
    }
    a.tri_mask = triangle_mask;
    a.tx = tx; a.rx = rx; a.cand = cand;
    a.T = T; a.ntx = ntx; a.nrx = nrx; a.C = C; a.P = P;
    a.eps = epsilon; a.thr = 1.0f - hit_tol; a.min_len = min_len; a.alpha = smoothing_factor;
    a.out_vertices = out_vertices; a.out_objects = out_objects; a.out_mask = out_mask;
    const bool q = assume_quads != 0;
    switch (order) {
        case 0: return launch_trace_smooth<0>(a, q, s);
        case 1: return launch_trace_smooth<1>(a, q, s);
        case 2: return launch_trace_smooth<2>(a, q, s);
        case 3: return launch_trace_smooth<3>(a, q, s);
        case 4: return launch_trace_smooth<4>(a, q, s);
        case 5: return launch_trace_smooth<5>(a, q, s);
        case 6: return launch_trace_smooth<6>(a, q, s);
        case 7: return launch_trace_smooth<7>(a, q, s);
        case 8: return launch_trace_smooth<8>(a, q, s);
        default: return DRT_ERR_UNSUPPORTED;
    }
}

}  // extern "C"
