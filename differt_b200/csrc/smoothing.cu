// N4 (forward): the smoothed variants of the intersection primitives.
// Reference: differt/src/differt/utils.py:70-89 (smoothing_function = sigmoid(x * alpha)),
// differt/src/differt/geometry/_utils.py:1279-1318 (ray_intersect_triangle), :1465-1476
// (ray_intersect_any_triangle), _solver_image_method.py:448-454 (same side of mirrors).
// Comparisons become sigmoids, AND becomes min, the OR over triangles becomes a sum clipped at 1.
// Outputs are floats in [0, 1]; with the transcendental involved parity is to tolerance (1e-5), not
// bit-exact.  The relaxed trace (_solvers.py:599-713) and its reverse mode are the last two entry
// points of this file.
#include "common.cuh"
#include "image_core.cuh"

namespace drt {

__device__ __forceinline__ float smooth(float x, float alpha) {  // jax.nn.sigmoid(x * alpha)
    // 1 / y correctly rounded: the same value as __fdiv_rn(1, y), without the generic division path
    return __frcp_rn(1.0f + expf(-(x * alpha)));
}

// jnp.min / jnp.minimum propagate NaN; fminf would drop it.  One FMNMX.NAN instead of two compares,
// a select and a min.
__device__ __forceinline__ float nanmin(float a, float b) {
    float r;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

// The sigmoid is monotone, so the reference's min over sigmoids (_utils.py:1279-1318: AND → min) is the
// sigmoid of the min of their arguments (max for a negative slope): ONE exponential per relaxed test
// instead of seven.  NaN arguments propagate like in jnp.min; the constant 1 of the reference's stacks
// never wins (sigmoid <= 1).  `m` is min_i x_i for alpha >= 0, max_i x_i otherwise.
struct RelaxedArgs {
    float sgn;  // +1 (alpha >= 0) or -1
    float m;    // running extremum of sgn * x_i
    __device__ __forceinline__ explicit RelaxedArgs(float alpha, float x0) : sgn(alpha < 0.0f ? -1.0f : 1.0f), m(sgn * x0) {}
    __device__ __forceinline__ void add(float x) { m = nanmin(m, sgn * x); }
    __device__ __forceinline__ float value(float alpha) const { return smooth(sgn * m, alpha); }
};

// _utils.py:1263-1322 with smoothing_factor; returns the smoothed hit, writes t.  `extra` (blockage):
// also folds sigmoid((thr - t) alpha) into the min (_utils.py:1465-1473).
__device__ __forceinline__ float mt_smooth(const float3 o, const float3 d, const Tri &tr, const float eps,
                                           const float alpha, float &t, const bool extra = false,
                                           const float thr = 0.0f) {
    const float3 h = cross3(d, tr.e2);
    float a = dot3(h, tr.e1);
    a = (a == 0.0f) ? CUDART_INF_F : a;
    RelaxedArgs r(alpha, fabsf(a) - eps);
    const float f = __frcp_rn(a);
    const float3 s = sub3(o, tr.v0);
    const float u = f * dot3(s, h);
    r.add(u - 0.0f);
    r.add(1.0f - u);
    const float3 q = cross3(s, tr.e1);
    const float v = f * dot3(q, d);
    r.add(v - 0.0f);
    r.add(1.0f - (u + v));
    t = f * dot3(q, tr.e2);
    r.add(t - eps);
    if (extra) r.add(thr - t);
    return r.value(alpha);
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
        if (is_never_hit(a)) continue;  // never-hit record: inactive
        float t;
        acc += mt_smooth(oo, dd, unpack(a, b, c), eps, alpha, t, true, thr);
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
__device__ __forceinline__ float nanmax(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

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
            if (is_never_hit(ra)) continue;  // inactive
            const Tri tr = unpack(ra, rb, rc);
#pragma unroll
            for (int s = 0; s < NSEG; ++s) {
                float t;
                acc[s] += mt_smooth(o[s], d[s], tr, a.eps, a.alpha, t, true, a.thr);
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

// ------------------------------------------------------------------------------------------------
// reverse mode of the relaxed trace (what jax.grad gives on _solvers.py:576-713)
// ------------------------------------------------------------------------------------------------
// min / max route the cotangent to their (first) extremal argument — JAX shares it between exact
// ties, which only happen between saturated terms whose sigmoid derivative is 0 anyway; the
// same-side term is a function of signs (zero gradient); the clip of the blockage sum passes the
// cotangent while the sum is below 1.  Non-finite paths and NaN confidences get a zero gradient.

struct MtGrad {  // cotangents of one relaxed Möller–Trumbore evaluation
    float3 o, d, v0, e1, e2;
};

__device__ __forceinline__ void atomic_add3s(float *p, float3 v) {
    if (v.x != 0.f) atomicAdd(p, v.x);
    if (v.y != 0.f) atomicAdd(p + 1, v.y);
    if (v.z != 0.f) atomicAdd(p + 2, v.z);
}

// Seeds (ga, gu, gv, gt) → cotangents of (o, d, v0, e1, e2).  Recomputes the forward
// (h = d x e2, a = h.e1, f = 1/a, s = o - v0, u = f s.h, q = s x e1, v = f q.d, t = f q.e2).
// `select` != 0: add g times the derivative of the relaxed hit (1) or of min(hit, sigmoid((thr - t) alpha))
// (2), routed to the first minimal term.  Returns false when nothing flows (a == 0, NaN, saturated).
__device__ __forceinline__ bool mt_smooth_adjoint(const float3 o, const float3 d, const Tri &tr, const float eps,
                                                  const float alpha, const float thr, const int select,
                                                  const float g, float gt, MtGrad &out) {
    const float3 h = cross3(d, tr.e2);
    const float a = dot3(h, tr.e1);
    if (a == 0.0f) return false;  // where(a == 0, inf, a): f = 0 and every output is constant
    const float f = __frcp_rn(a);
    const float3 s = sub3(o, tr.v0);
    const float su = dot3(s, h);
    const float u = f * su;
    const float3 q = cross3(s, tr.e1);
    const float qv = dot3(q, d);
    const float v = f * qv;
    const float qt = dot3(q, tr.e2);
    const float t = f * qt;
    float ga = 0.f, gu = 0.f, gv = 0.f;
    if (select != 0) {
        // first extremal argument (the forward's RelaxedArgs), then one sigmoid derivative
        const float sg = alpha < 0.0f ? -1.0f : 1.0f;
        float best = sg * (fabsf(a) - eps);
        int which = 0;
        bool nan = best != best;
#define DRT_CONSIDER(val, id)            \
    {                                    \
        const float x_ = sg * (val);     \
        nan = nan || (x_ != x_);         \
        if (x_ < best) {                 \
            best = x_;                   \
            which = (id);                \
        }                                \
    }
        DRT_CONSIDER(u - 0.0f, 1)
        DRT_CONSIDER(1.0f - u, 2)
        DRT_CONSIDER(v - 0.0f, 3)
        DRT_CONSIDER(1.0f - (u + v), 4)
        DRT_CONSIDER(t - eps, 5)
        if (select == 2) DRT_CONSIDER(thr - t, 6)
#undef DRT_CONSIDER
        const float sv = smooth(sg * best, alpha);
        float ds = g * alpha * sv * (1.0f - sv);  // d sigmoid(x alpha) / dx, times the cotangent
        if (nan || ds != ds) ds = 0.0f;
        switch (which) {
            case 0: ga = a < 0.0f ? -ds : ds; break;
            case 1: gu = ds; break;
            case 2: gu = -ds; break;
            case 3: gv = ds; break;
            case 4: gu = -ds; gv = -ds; break;
            case 5: gt += ds; break;
            default: gt -= ds; break;
        }
    }
    if (ga == 0.0f && gu == 0.0f && gv == 0.0f && gt == 0.0f) return false;
    const float gf = gu * su + gv * qv + gt * qt;
    const float gsu = gu * f, gqv = gv * f, gqt = gt * f;
    ga -= gf * f * f;                                            // f = 1 / a
    const float3 gq = add3(scale3(d, gqv), scale3(tr.e2, gqt));  // qv = q.d, qt = q.e2
    const float3 gh = add3(scale3(s, gsu), scale3(tr.e1, ga));   // su = s.h, a = h.e1
    const float3 gs = add3(scale3(h, gsu), cross3(tr.e1, gq));   // q = s x e1
    out.o = gs;                                                  // s = o - v0
    out.v0 = make_float3(-gs.x, -gs.y, -gs.z);
    out.d = add3(scale3(q, gqv), cross3(tr.e2, gh));             // h = d x e2
    out.e1 = add3(scale3(h, ga), cross3(gq, s));
    out.e2 = add3(scale3(q, gqt), cross3(gh, d));
    return true;
}

// g: cotangent of hit (blockage == false) or of min(hit, sigmoid((thr - t) alpha)) (blockage == true)
__device__ __forceinline__ bool mt_smooth_grad(const float3 o, const float3 d, const Tri &tr, const float eps,
                                               const float alpha, const float thr, const bool blockage,
                                               const float g, MtGrad &out) {
    return mt_smooth_adjoint(o, d, tr, eps, alpha, thr, blockage ? 2 : 1, g, 0.0f, out);
}

// reverse mode of the relaxed primitives called on their own ------------------------------------

// elementwise: cotangents of (t, hit) → (o, d, triangle vertices) per element; the host sums over
// broadcast axes (autograd of expand)
__global__ void mt_smooth_vjp_kernel(int64_t n, const float *__restrict__ o, const float *__restrict__ d,
                                     const float *__restrict__ tri, float eps, float alpha,
                                     const float *__restrict__ g_t, const float *__restrict__ g_hit,
                                     float *__restrict__ g_o, float *__restrict__ g_d, float *__restrict__ g_tri) {
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    const float3 v0 = ld3(tri + 9 * i), v1 = ld3(tri + 9 * i + 3), v2 = ld3(tri + 9 * i + 6);
    Tri tr;
    tr.v0 = v0;
    tr.e1 = sub3(v1, v0);
    tr.e2 = sub3(v2, v0);
    MtGrad mg;
    const float3 z = make_float3(0.f, 0.f, 0.f);
    mg.o = mg.d = mg.v0 = mg.e1 = mg.e2 = z;
    const float gh = g_hit ? g_hit[i] : 0.0f, gt = g_t ? g_t[i] : 0.0f;
    if (!mt_smooth_adjoint(ld3(o + 3 * i), ld3(d + 3 * i), tr, eps, alpha, 0.0f, gh != 0.0f ? 1 : 0, gh, gt, mg))
        mg.o = mg.d = mg.v0 = mg.e1 = mg.e2 = z;
    st3(g_o + 3 * i, mg.o);
    st3(g_d + 3 * i, mg.d);
    st3(g_tri + 9 * i, sub3(sub3(mg.v0, mg.e1), mg.e2));
    st3(g_tri + 9 * i + 3, mg.e1);
    st3(g_tri + 9 * i + 6, mg.e2);
}

// one warp per ray: recompute the clipped sum exactly like any_smooth_kernel; below the clip the
// cotangent goes through every active triangle's term
__global__ void __launch_bounds__(256)
any_smooth_vjp_kernel(int64_t R, int64_t T, const float *__restrict__ o, const float *__restrict__ d,
                      const Tri48 *__restrict__ pack, float eps, float thr, float alpha,
                      const float *__restrict__ g_out, float *__restrict__ g_o, float *__restrict__ g_d,
                      float *g_tri) {
    const int lane = threadIdx.x & 31;
    const int64_t ray = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
    if (ray >= R) return;
    const float3 oo = ld3(o + 3 * ray), dd = ld3(d + 3 * ray);
    float acc = 0.0f;
    for (int64_t j = lane; j < T; j += 32) {
        const float4 a = pack[j].a, b = pack[j].b, c = pack[j].c;
        if (is_never_hit(a)) continue;
        float t;
        acc += mt_smooth(oo, dd, unpack(a, b, c), eps, alpha, t, true, thr);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(kFull, acc, off);
    float3 go = make_float3(0.f, 0.f, 0.f), gd = make_float3(0.f, 0.f, 0.f);
    const float g = g_out[ray];
    if (acc < 1.0f && g != 0.0f && g == g) {
        for (int64_t j = lane; j < T; j += 32) {
            const float4 a = pack[j].a, b = pack[j].b, c = pack[j].c;
            if (is_never_hit(a)) continue;
            MtGrad mg;
            if (mt_smooth_adjoint(oo, dd, unpack(a, b, c), eps, alpha, thr, 2, g, 0.0f, mg)) {
                go = add3(go, mg.o);
                gd = add3(gd, mg.d);
                atomic_add3s(g_tri + 9 * j, sub3(sub3(mg.v0, mg.e1), mg.e2));
                atomic_add3s(g_tri + 9 * j + 3, mg.e1);
                atomic_add3s(g_tri + 9 * j + 6, mg.e2);
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        go.x += __shfl_xor_sync(kFull, go.x, off);
        go.y += __shfl_xor_sync(kFull, go.y, off);
        go.z += __shfl_xor_sync(kFull, go.z, off);
        gd.x += __shfl_xor_sync(kFull, gd.x, off);
        gd.y += __shfl_xor_sync(kFull, gd.y, off);
        gd.z += __shfl_xor_sync(kFull, gd.z, off);
    }
    if (lane == 0) {
        st3(g_o + 3 * ray, go);
        st3(g_d + 3 * ray, gd);
    }
}

struct SmoothVjpArgs {
    SmoothTraceArgs f;          // the forward's arguments (out_* unused except out_mask = saved confidences)
    const int32_t *triangles;   // [T, 3]
    int64_t V;
    const float *g_mask;        // [P]
    const float *g_out_vertices;  // nullable [P, k+2, 3]: cotangent of the dense path vertices
    float *g_full;              // [P, k+2, 3] workspace: total cotangent of the path vertices
    float *g_verts_direct;      // [V, 3] workspace: cotangent reaching the mesh through triangle geometry
};

// scatter of (v0, e1, e2) cotangents of triangle `tri` onto its three mesh vertices
__device__ __forceinline__ void scatter_triangle_grad(const SmoothVjpArgs &a, int64_t tri, const MtGrad &m) {
    const int32_t *ix = a.triangles + 3 * tri;
    const int64_t i0 = min(max(int64_t(ix[0]), int64_t(0)), a.V - 1);
    const int64_t i1 = min(max(int64_t(ix[1]), int64_t(0)), a.V - 1);
    const int64_t i2 = min(max(int64_t(ix[2]), int64_t(0)), a.V - 1);
    atomic_add3s(a.g_verts_direct + 3 * i0, sub3(sub3(m.v0, m.e1), m.e2));  // e1 = v1 - v0, e2 = v2 - v0
    atomic_add3s(a.g_verts_direct + 3 * i1, m.e1);
    atomic_add3s(a.g_verts_direct + 3 * i2, m.e2);
}

// one thread per path: recompute the forward's stage terms, find which one the confidence came from,
// write the total cotangent of the path vertices (downstream + stage) and flag the paths whose
// confidence came from the blockage term for the warp-per-path kernel below.
template <int K, bool QUADS>
__global__ void __launch_bounds__(128) trace_smooth_vjp_stage_kernel(const SmoothVjpArgs a) {
    constexpr int KK = K > 0 ? K : 1;
    constexpr int NT = QUADS ? 2 : 1;
    const SmoothTraceArgs &fa = a.f;
    const int64_t p = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (p >= fa.P) return;
    const int64_t c = p % fa.C, pair = p / fa.C;
    const int64_t irx = pair % fa.nrx, itx = pair / fa.nrx;

    float3 mv[KK], mn[KK];
    Tri tri[KK][NT];
    int32_t ti[KK];
    bool active = true;
#pragma unroll
    for (int i = 0; i < K; ++i) {
        int32_t t = fa.cand[c * K + i];
        t = min(max(t, 0), int32_t(fa.T - (QUADS ? 2 : 1)));
        ti[i] = t;
#pragma unroll
        for (int q = 0; q < NT; ++q) {
            const float4 ta = fa.pack[t + q].a, tb = fa.pack[t + q].b, tc = fa.pack[t + q].c;
            tri[i][q] = unpack(ta, tb, tc);
            if (q == 0) {
                mv[i] = make_float3(ta.x, ta.y, ta.z);
                mn[i] = make_float3(tc.y, tc.z, tc.w);
            }
            if (fa.tri_mask != nullptr) active = active && fa.tri_mask[t + q] != 0;
        }
    }
    float3 full[K + 2];
    full[0] = ld3(fa.tx + 3 * itx);
    full[K + 1] = ld3(fa.rx + 3 * irx);
    image_method_path<K>(full, mv, mn);

    float inside = 1.0f, same = 1.0f, small = 0.0f;
    int inside_i = -1, inside_q = 0, small_s = -1;
    bool finite = true;
#pragma unroll
    for (int i = 0; i <= K; ++i) {
        const float3 o = full[i];
        const float3 d = sub3(full[i + 1], full[i]);
        const float sm = smooth(fa.min_len - dot3(d, d), fa.alpha);
        if (sm > small) {
            small = sm;
            small_s = i;
        }
        if (sm != sm) small = sm;
        if (i < K) {
            float tt;
            float hit = mt_smooth(o, d, tri[i][0], fa.eps, fa.alpha, tt);
            int q = 0;
            if (QUADS) {
                const float h1 = mt_smooth(o, d, tri[i][NT - 1], fa.eps, fa.alpha, tt);
                if (h1 > hit) q = 1;
                hit = nanmax(nanmax(hit, h1), 0.0f);
            }
            if (hit < inside) {
                inside = hit;
                inside_i = i;
                inside_q = q;
            }
            if (hit != hit) inside = hit;
            const float dp = dot3(sub3(full[i], mv[i]), mn[i]);
            const float dn = dot3(sub3(full[i + 2], mv[i]), mn[i]);
            const float sp = dp != dp ? dp : float(dp > 0.0f) - float(dp < 0.0f);
            const float sn = dn != dn ? dn : float(dn > 0.0f) - float(dn < 0.0f);
            same = nanmin(same, smooth(sp * sn, fa.alpha));
        }
    }
#pragma unroll
    for (int i = 0; i < K + 2; ++i) finite = finite && finite3(full[i]);

    float3 gfull[K + 2];
#pragma unroll
    for (int i = 0; i < K + 2; ++i)
        gfull[i] = a.g_out_vertices ? ld3(a.g_out_vertices + (p * (K + 2) + i) * 3) : make_float3(0.f, 0.f, 0.f);

    const float g = a.g_mask[p];
    const float m_stage = nanmin(nanmin(inside, same), 1.0f - small);
    const float conf = fa.out_mask[p];  // the forward's min(stage, 1 - blocked) * active
    uint8_t to_blockage = 0;
    if (finite && active && g != 0.0f && g == g && m_stage == m_stage && conf == conf) {
        if (conf < m_stage) {
            to_blockage = 1;  // 1 - blocked is the minimum
        } else if (inside <= same && inside <= 1.0f - small) {
            if (inside_i >= 0) {
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    if (i != inside_i) continue;
                    MtGrad mg;
                    const float3 o = full[i], d = sub3(full[i + 1], full[i]);
                    if (mt_smooth_grad(o, d, tri[i][(QUADS && inside_q) ? NT - 1 : 0], fa.eps, fa.alpha, 0.0f, false,
                                       g, mg)) {
                        gfull[i] = add3(gfull[i], sub3(mg.o, mg.d));  // d = full[i+1] - full[i]
                        gfull[i + 1] = add3(gfull[i + 1], mg.d);
                        scatter_triangle_grad(a, ti[i] + ((QUADS && inside_q) ? 1 : 0), mg);
                    }
                }
            }
        } else if (!(same <= 1.0f - small)) {
            // 1 - too_small is the minimum: d(1 - sigmoid((min_len - d.d) alpha)) = 2 alpha s (1 - s) d
#pragma unroll
            for (int i = 0; i <= K; ++i) {
                if (i != small_s) continue;
                const float3 d = sub3(full[i + 1], full[i]);
                const float3 gd = scale3(d, 2.0f * g * fa.alpha * small * (1.0f - small));
                gfull[i] = sub3(gfull[i], gd);
                gfull[i + 1] = add3(gfull[i + 1], gd);
            }
        }
    }
    a.f.flags[p] = to_blockage;
    float *gf = a.g_full + p * (K + 2) * 3;
#pragma unroll
    for (int i = 0; i < K + 2; ++i) st3(gf + 3 * i, gfull[i]);
}

// one warp per flagged path: recompute the per-segment blockage sums, pick the segment the max came
// from, and if its sum is below the clip send -g through every active triangle's term.
template <int NSEG>
__global__ void __launch_bounds__(256) trace_smooth_vjp_blocked_kernel(const SmoothVjpArgs a) {
    const SmoothTraceArgs &fa = a.f;
    const int lane = threadIdx.x & 31;
    const int64_t p = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
    if (p >= fa.P || fa.flags[p] == 0) return;  // warp-uniform
    const float *v = fa.out_vertices + p * (NSEG + 1) * 3;
    float3 o[NSEG], d[NSEG];
    float acc[NSEG];
#pragma unroll
    for (int s = 0; s < NSEG; ++s) {
        o[s] = ld3(v + 3 * s);
        d[s] = sub3(ld3(v + 3 * s + 3), o[s]);
        acc[s] = 0.0f;
    }
    for (int64_t j = lane; j < fa.T; j += 32) {
        const float4 ra = fa.pack_active[j].a, rb = fa.pack_active[j].b, rc = fa.pack_active[j].c;
        if (is_never_hit(ra)) continue;
        const Tri tr = unpack(ra, rb, rc);
#pragma unroll
        for (int s = 0; s < NSEG; ++s) {
            float t;
            acc[s] += mt_smooth(o[s], d[s], tr, fa.eps, fa.alpha, t, true, fa.thr);
        }
    }
    int best_s = 0;
    float best = -1.0f;
#pragma unroll
    for (int s = 0; s < NSEG; ++s) {
        float x = acc[s];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(kFull, x, off);
        if (x > best) {
            best = x;
            best_s = s;
        }
    }
    if (!(best < 1.0f)) return;  // clipped (or NaN): the cotangent stops here
    const float g = -a.g_mask[p];  // confidence = 1 - blocked
    const float3 os = ld3(v + 3 * best_s);
    const float3 ds = sub3(ld3(v + 3 * best_s + 3), os);
    float3 go = make_float3(0.f, 0.f, 0.f), gd = make_float3(0.f, 0.f, 0.f);
    for (int64_t j = lane; j < fa.T; j += 32) {
        const float4 ra = fa.pack_active[j].a, rb = fa.pack_active[j].b, rc = fa.pack_active[j].c;
        if (is_never_hit(ra)) continue;
        MtGrad mg;
        if (mt_smooth_grad(os, ds, unpack(ra, rb, rc), fa.eps, fa.alpha, fa.thr, true, g, mg)) {
            go = add3(go, mg.o);
            gd = add3(gd, mg.d);
            scatter_triangle_grad(a, j, mg);
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        go.x += __shfl_xor_sync(kFull, go.x, off);
        go.y += __shfl_xor_sync(kFull, go.y, off);
        go.z += __shfl_xor_sync(kFull, go.z, off);
        gd.x += __shfl_xor_sync(kFull, gd.x, off);
        gd.y += __shfl_xor_sync(kFull, gd.y, off);
        gd.z += __shfl_xor_sync(kFull, gd.z, off);
    }
    if (lane == 0) {  // this warp owns the path: plain read-modify-write
        float *gf = a.g_full + (p * (NSEG + 1) + best_s) * 3;
        const float3 g0 = add3(ld3(gf), sub3(go, gd)), g1 = add3(ld3(gf + 3), gd);
        st3(gf, g0);
        st3(gf + 3, g1);
    }
}

__global__ void add_inplace_kernel(int64_t n, float *__restrict__ dst, const float *__restrict__ src) {
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i < n) dst[i] += src[i];
}

template <int K>
static int launch_trace_smooth_vjp(const SmoothVjpArgs &a, bool quads, cudaStream_t s) {
    const unsigned blocks = unsigned((a.f.P + 127) / 128);
    if (quads)
        trace_smooth_vjp_stage_kernel<K, true><<<blocks, 128, 0, s>>>(a);
    else
        trace_smooth_vjp_stage_kernel<K, false><<<blocks, 128, 0, s>>>(a);
    if (cudaGetLastError() != cudaSuccess) return DRT_ERR_CUDA;
    if (a.f.T > 0) trace_smooth_vjp_blocked_kernel<K + 1><<<unsigned((a.f.P * 32 + 255) / 256), 256, 0, s>>>(a);
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
    mt_smooth_elementwise_kernel<<<unsigned(blocks < int64_t(drt::device_sm_count()) * 16 ? blocks : int64_t(drt::device_sm_count()) * 16), 256, 0,
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
    same_side_smooth_kernel<<<unsigned(blocks < int64_t(drt::device_sm_count()) * 16 ? blocks : int64_t(drt::device_sm_count()) * 16), 256, 0,
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

size_t drt_trace_smooth_vjp_workspace_bytes(int64_t V, int64_t T, int64_t ntx, int64_t nrx, int64_t C,
                                            int32_t order) {
    if (V < 0 || T < 0 || ntx < 0 || nrx < 0 || C < 0 || order < 0) return 0;
    const size_t P = size_t(ntx) * size_t(nrx) * size_t(C);
    const size_t a256 = 255;
    return 2 * drt_mesh_pack_bytes(T) + ((P + a256) & ~a256) + ((P * size_t(order + 2) * 12 + a256) & ~a256) +
           ((size_t(V) * 12 + a256) & ~a256) + 256;
}

int drt_trace_path_candidates_smooth_vjp(drt_stream_t stream, int64_t V, int64_t T, const float *vertices,
                                         const int32_t *triangles, const uint8_t *triangle_mask,
                                         int32_t assume_quads, int64_t ntx, const float *tx, int64_t nrx,
                                         const float *rx, int64_t C, int32_t order, const int32_t *cand,
                                         float epsilon, float hit_tol, float min_len, float smoothing_factor,
                                         const float *out_vertices, const float *out_mask, const float *g_out_vertices,
                                         const float *g_out_mask, void *workspace, size_t workspace_bytes,
                                         float *g_tx, float *g_rx, float *g_vertices) {
    if (V < 0 || T < 0 || ntx < 0 || nrx < 0 || C < 0 || order < 0) return DRT_ERR_BAD_EXTENT;
    if (order > DRT_MAX_ORDER) return DRT_ERR_UNSUPPORTED;
    const int64_t P = ntx * nrx * C;
    if (P >= (int64_t(1) << 32)) return DRT_ERR_BAD_EXTENT;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (P == 0 || g_out_mask == nullptr)  // nothing but (possibly) the path-vertex cotangent
        return drt_trace_path_candidates_vjp(stream, V, T, vertices, triangles, ntx, tx, nrx, rx, C, order, cand,
                                             g_out_vertices, g_tx, g_rx, g_vertices);
    if (!tx || !rx || !out_vertices || !out_mask || (order > 0 && !cand)) return DRT_ERR_NULL_POINTER;
    if (order > 0 && T < (assume_quads ? 2 : 1)) return DRT_ERR_BAD_EXTENT;
    if (T > 0 && (!vertices || !triangles)) return DRT_ERR_NULL_POINTER;
    if (!workspace || workspace_bytes < drt_trace_smooth_vjp_workspace_bytes(V, T, ntx, nrx, C, order))
        return DRT_ERR_WORKSPACE;
    char *ws = static_cast<char *>(workspace);
    const size_t pb = drt_mesh_pack_bytes(T), a256 = 255;
    const size_t off_flags = 2 * pb;
    const size_t off_gfull = off_flags + ((size_t(P) + a256) & ~a256);
    const size_t off_gv = off_gfull + ((size_t(P) * size_t(order + 2) * 12 + a256) & ~a256);
    SmoothVjpArgs a{};
    SmoothTraceArgs &f = a.f;
    f.pack = reinterpret_cast<const Tri48 *>(ws);
    f.pack_active = f.pack;
    f.flags = reinterpret_cast<uint8_t *>(ws + off_flags);
    if (T > 0) {
        int rc = drt_mesh_pack(stream, V, T, vertices, triangles, nullptr, ws);
        if (rc != DRT_OK) return rc;
        if (triangle_mask) {
            rc = drt_mesh_pack(stream, V, T, vertices, triangles, triangle_mask, ws + pb);
            if (rc != DRT_OK) return rc;
            f.pack_active = reinterpret_cast<const Tri48 *>(ws + pb);
        }
    }
    f.tri_mask = triangle_mask;
    f.tx = tx; f.rx = rx; f.cand = cand;
    f.T = T; f.ntx = ntx; f.nrx = nrx; f.C = C; f.P = P;
    f.eps = epsilon; f.thr = 1.0f - hit_tol; f.min_len = min_len; f.alpha = smoothing_factor;
    f.out_vertices = const_cast<float *>(out_vertices);
    f.out_mask = const_cast<float *>(out_mask);
    a.triangles = triangles;
    a.V = V;
    a.g_mask = g_out_mask;
    a.g_out_vertices = g_out_vertices;
    a.g_full = reinterpret_cast<float *>(ws + off_gfull);
    a.g_verts_direct = reinterpret_cast<float *>(ws + off_gv);
    if (V > 0 && cudaMemsetAsync(a.g_verts_direct, 0, size_t(V) * 12, s) != cudaSuccess) return DRT_ERR_CUDA;
    const bool q = assume_quads != 0;
    int rc = DRT_ERR_UNSUPPORTED;
    switch (order) {
        case 0: rc = launch_trace_smooth_vjp<0>(a, q, s); break;
        case 1: rc = launch_trace_smooth_vjp<1>(a, q, s); break;
        case 2: rc = launch_trace_smooth_vjp<2>(a, q, s); break;
        case 3: rc = launch_trace_smooth_vjp<3>(a, q, s); break;
        case 4: rc = launch_trace_smooth_vjp<4>(a, q, s); break;
        case 5: rc = launch_trace_smooth_vjp<5>(a, q, s); break;
        case 6: rc = launch_trace_smooth_vjp<6>(a, q, s); break;
        case 7: rc = launch_trace_smooth_vjp<7>(a, q, s); break;
        case 8: rc = launch_trace_smooth_vjp<8>(a, q, s); break;
        default: break;
    }
    if (rc != DRT_OK) return rc;
    // through the image method, the mirror gathers and the normals (K6b), then the direct part
    rc = drt_trace_path_candidates_vjp(stream, V, T, vertices, triangles, ntx, tx, nrx, rx, C, order, cand,
                                       a.g_full, g_tx, g_rx, g_vertices);
    if (rc != DRT_OK) return rc;
    if (V > 0) {
        add_inplace_kernel<<<unsigned((V * 3 + 255) / 256), 256, 0, s>>>(V * 3, g_vertices, a.g_verts_direct);
        if (cudaGetLastError() != cudaSuccess) return DRT_ERR_CUDA;
    }
    return DRT_OK;
}

int drt_ray_intersect_triangle_smooth_vjp(drt_stream_t stream, int64_t n, const float *o, const float *d,
                                          const float *tri, float epsilon, float smoothing_factor,
                                          const float *g_t, const float *g_hit, float *g_o, float *g_d,
                                          float *g_tri) {
    if (n < 0) return DRT_ERR_BAD_EXTENT;
    if (n == 0) return DRT_OK;
    if (!o || !d || !tri || !g_o || !g_d || !g_tri) return DRT_ERR_NULL_POINTER;
    mt_smooth_vjp_kernel<<<unsigned((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        n, o, d, tri, epsilon, smoothing_factor, g_t, g_hit, g_o, g_d, g_tri);
    return cudaGetLastError() == cudaSuccess ? DRT_OK : DRT_ERR_CUDA;
}

int drt_ray_intersect_any_triangle_smooth_vjp(drt_stream_t stream, int64_t R, const float *o, const float *d,
                                              const void *pack, int64_t T, float epsilon, float hit_tol,
                                              float smoothing_factor, const float *g_out, float *g_o,
                                              float *g_d, float *g_tri) {
    if (R < 0 || T < 0) return DRT_ERR_BAD_EXTENT;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (T > 0 && !g_tri) return DRT_ERR_NULL_POINTER;
    if (T > 0 && cudaMemsetAsync(g_tri, 0, size_t(T) * 36, s) != cudaSuccess) return DRT_ERR_CUDA;
    if (R == 0) return DRT_OK;
    if (!g_o || !g_d) return DRT_ERR_NULL_POINTER;
    if (T == 0) {
        if (cudaMemsetAsync(g_o, 0, size_t(R) * 12, s) != cudaSuccess) return DRT_ERR_CUDA;
        return cudaMemsetAsync(g_d, 0, size_t(R) * 12, s) == cudaSuccess ? DRT_OK : DRT_ERR_CUDA;
    }
    if (!o || !d || !pack || !g_out) return DRT_ERR_NULL_POINTER;
    any_smooth_vjp_kernel<<<unsigned((R * 32 + 255) / 256), 256, 0, s>>>(
        R, T, o, d, static_cast<const Tri48 *>(pack), epsilon, 1.0f - hit_tol, smoothing_factor, g_out, g_o, g_d,
        g_tri);
    return cudaGetLastError() == cudaSuccess ? DRT_OK : DRT_ERR_CUDA;
}

}  // extern "C"
