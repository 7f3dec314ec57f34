// Image-method arithmetic shared by K5 (image_method), K5b (its VJP) and K6 (fused trace).
// Reference: differt/src/differt/geometry/_solver_image_method.py:68-79, 110-135, 138-203.
#pragma once

#include "common.cuh"

namespace drt {

// p - 2 ((p - m)·n) n      (_solver_image_method.py:73-79)
__device__ __forceinline__ float3 mirror_image(float3 p, float3 m, float3 n) {
    const float c = 2.0f * dot3(sub3(p, m), n);
    return make_float3(p.x - c * n.x, p.y - c * n.y, p.z - c * n.z);
}

struct BackStep {  // what the reverse sweep needs from one backward step
    float3 dir, w;
    float un, vn, t;
    bool parallel;
    unsigned no_prev;  // bit c set: component c of the previous point was ±inf
};

// one backward step (_solver_image_method.py:152-182 + 110-135): intersection of the ray
// prev → image with the mirror plane, with the reference's inf guard.
__device__ __forceinline__ float3 back_step(float3 prev, float3 image, float3 m, float3 n,
                                            BackStep *rec) {
    const bool ix = isinf(prev.x), iy = isinf(prev.y), iz = isinf(prev.z);
    const float3 p0 = make_float3(ix ? 0.0f : prev.x, iy ? 0.0f : prev.y, iz ? 0.0f : prev.z);
    const float3 u = sub3(image, p0);
    const float3 w = sub3(m, p0);
    float un = dot3(u, n);
    const float vn = dot3(w, n);
    const bool par = (un == 0.0f);
    un = par ? 1.0f : un;
    const float t = __fdiv_rn(vn, un);
    float3 r = make_float3(p0.x + u.x * t, p0.y + u.y * t, p0.z + u.z * t);
    if (par && vn != 0.0f) r = make_float3(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F);
    if (ix) r.x = CUDART_INF_F;
    if (iy) r.y = CUDART_INF_F;
    if (iz) r.z = CUDART_INF_F;
    if (rec != nullptr) {
        rec->dir = u;
        rec->w = w;
        rec->un = un;
        rec->vn = vn;
        rec->t = t;
        rec->parallel = par;
        rec->no_prev = (ix ? 1u : 0u) | (iy ? 2u : 0u) | (iz ? 4u : 0u);
    }
    return r;
}

// full[0] = from, full[K+1] = to on entry; fills full[1..K].
template <int K>
__device__ __forceinline__ void image_method_path(float3 (&full)[K + 2], const float3 (&mv)[K > 0 ? K : 1],
                                                  const float3 (&mn)[K > 0 ? K : 1]) {
    float3 img[K > 0 ? K : 1];
    float3 prev = full[0];
#pragma unroll
    for (int i = 0; i < K; ++i) {
        prev = mirror_image(prev, mv[i], mn[i]);
        img[i] = prev;
    }
    prev = full[K + 1];
#pragma unroll
    for (int i = K - 1; i >= 0; --i) {
        prev = back_step(prev, img[i], mv[i], mn[i], nullptr);
        full[i + 1] = prev;
    }
}

// Reverse sweep.  g_p[i] is the cotangent of path point i+1 (i = 0..K-1).
// Accumulates into g_from, g_to, g_mv[i], g_mn[i] (all must be initialised by the caller).
template <int K>
__device__ __forceinline__ void image_method_reverse(float3 from, float3 to,
                                                     const float3 (&mv)[K > 0 ? K : 1],
                                                     const float3 (&mn)[K > 0 ? K : 1],
                                                     const float3 (&g_p)[K > 0 ? K : 1], float3 &g_from,
                                                     float3 &g_to, float3 (&g_mv)[K > 0 ? K : 1],
                                                     float3 (&g_mn)[K > 0 ? K : 1]) {
    float3 img[K + 1];
    BackStep rec[K > 0 ? K : 1];
    img[0] = from;
#pragma unroll
    for (int i = 0; i < K; ++i) img[i + 1] = mirror_image(img[i], mv[i], mn[i]);
    float3 prev = to;
#pragma unroll
    for (int i = K - 1; i >= 0; --i) prev = back_step(prev, img[i + 1], mv[i], mn[i], &rec[i]);

    float3 g_img[K + 1];
#pragma unroll
    for (int i = 0; i <= K; ++i) g_img[i] = make_float3(0.f, 0.f, 0.f);
    float3 carry = make_float3(0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < K; ++i) {
        const BackStep &r = rec[i];
        float3 g = add3(g_p[i], carry);
        const bool dead = r.parallel && r.vn != 0.0f;
        if (dead || (r.no_prev & 1u)) g.x = 0.f;
        if (dead || (r.no_prev & 2u)) g.y = 0.f;
        if (dead || (r.no_prev & 4u)) g.z = 0.f;
        const float g_t = dot3(g, r.dir);
        float3 g_dir = scale3(g, r.t);
        const float g_vn = __fdiv_rn(g_t, r.un);
        const float g_un = r.parallel ? 0.0f : -__fdiv_rn(g_t * r.t, r.un);
        g_dir = add3(g_dir, scale3(mn[i], g_un));
        g_mn[i] = add3(g_mn[i], add3(scale3(r.dir, g_un), scale3(r.w, g_vn)));
        const float3 g_w = scale3(mn[i], g_vn);
        g_mv[i] = add3(g_mv[i], g_w);
        g_img[i + 1] = add3(g_img[i + 1], g_dir);
        float3 gp = sub3(sub3(g, g_dir), g_w);
        if (r.no_prev & 1u) gp.x = 0.f;
        if (r.no_prev & 2u) gp.y = 0.f;
        if (r.no_prev & 4u) gp.z = 0.f;
        carry = gp;
    }
    g_to = add3(g_to, carry);
#pragma unroll
    for (int i = K - 1; i >= 0; --i) {
        const float3 gI = g_img[i + 1];
        const float3 inc = sub3(img[i], mv[i]);
        const float c = dot3(inc, mn[i]);
        const float g_c = -2.0f * dot3(gI, mn[i]);
        g_img[i] = add3(g_img[i], add3(gI, scale3(mn[i], g_c)));
        g_mv[i] = sub3(g_mv[i], scale3(mn[i], g_c));
        g_mn[i] = add3(g_mn[i], sub3(scale3(inc, g_c), scale3(gI, 2.0f * c)));
    }
    g_from = add3(g_from, g_img[0]);
}

}  // namespace drt
