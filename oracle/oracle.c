/* CPU oracle (C, OpenMP) for the DiffeRT geometric hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Same arithmetic as oracle/differt_oracle.py (which is pinned to the reference's golden vectors):
 * float32 everywhere, no fused multiply-add (build with -ffp-contract=off, never -ffast-math),
 * IEEE division and sqrt, three-term sums left to right.  It exists so that parity tests can run at
 * sizes NumPy cannot reach in seconds and so that bench.py has an all-cores CPU baseline.
 * tests/test_oracle_c.py checks it bit-for-bit against the NumPy oracle.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this library.
 *
 * Citations are relative to /root/reference/differt/src/differt/geometry/.
 *
 * Build: gcc -O3 -fopenmp -ffp-contract=off -fno-fast-math -shared -fPIC oracle.c -o liboracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

typedef struct { float x, y, z; } v3;

static inline v3 v3sub(v3 a, v3 b) { v3 r = {a.x - b.x, a.y - b.y, a.z - b.z}; return r; }
static inline float v3dot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline v3 v3cross(v3 a, v3 b) {
    v3 r = {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
    return r;
}
static inline v3 ld3(const float *p) { v3 r = {p[0], p[1], p[2]}; return r; }

/* _utils.py:1263-1322 — Möller–Trumbore with e1/e2 precomputed (same subtraction, same bits). */
static inline int mt_test(v3 o, v3 d, v3 v0, v3 e1, v3 e2, float eps, float *t_out) {
    v3 h = v3cross(d, e2);
    float a = v3dot(h, e1);
    if (a == 0.0f) a = INFINITY;
    int hit = fabsf(a) > eps;
    float f = 1.0f / a;
    v3 s = v3sub(o, v0);
    float u = f * v3dot(s, h);
    hit &= (u >= 0.0f) & (u <= 1.0f);
    v3 q = v3cross(s, e1);
    float v = f * v3dot(q, d);
    hit &= (v >= 0.0f) & (u + v <= 1.0f);
    float t = f * v3dot(q, e2);
    hit &= t > eps;
    *t_out = t;
    return hit;
}

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* a1: elementwise pairs. tri is [n,3,3]. */
ORC_API void orc_ray_intersect_triangle(int64_t n, const float *o, const float *d, const float *tri,
                                        float eps, float *t_out, uint8_t *hit_out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        v3 v0 = ld3(tri + 9 * i), v1 = ld3(tri + 9 * i + 3), v2 = ld3(tri + 9 * i + 6);
        float t;
        int hit = mt_test(ld3(o + 3 * i), ld3(d + 3 * i), v0, v3sub(v1, v0), v3sub(v2, v0), eps, &t);
        t_out[i] = t;
        hit_out[i] = (uint8_t)hit;
    }
}

typedef struct {
    int64_t T;
    float *v0x, *v0y, *v0z, *e1x, *e1y, *e1z, *e2x, *e2y, *e2z;
} soa_t;

static soa_t soa_build(int64_t T, const float *tri) {
    soa_t s;
    s.T = T;
    float *buf = (float *)malloc(sizeof(float) * 9 * (size_t)(T > 0 ? T : 1));
    s.v0x = buf; s.v0y = buf + T; s.v0z = buf + 2 * T;
    s.e1x = buf + 3 * T; s.e1y = buf + 4 * T; s.e1z = buf + 5 * T;
    s.e2x = buf + 6 * T; s.e2y = buf + 7 * T; s.e2z = buf + 8 * T;
    for (int64_t j = 0; j < T; ++j) {
        v3 v0 = ld3(tri + 9 * j), v1 = ld3(tri + 9 * j + 3), v2 = ld3(tri + 9 * j + 6);
        v3 e1 = v3sub(v1, v0), e2 = v3sub(v2, v0);
        s.v0x[j] = v0.x; s.v0y[j] = v0.y; s.v0z[j] = v0.z;
        s.e1x[j] = e1.x; s.e1y[j] = e1.y; s.e1z[j] = e1.z;
        s.e2x[j] = e2.x; s.e2y[j] = e2.y; s.e2z[j] = e2.z;
    }
    return s;
}
static void soa_free(soa_t *s) { free(s->v0x); }

/* One ray against triangles [j0, j1) of the SoA mesh: OR of (t < thr) & hit & active.
 * Dense (no early exit inside the span) so that the loop vectorises, like the reference's
 * batched map/reduce (_utils.py:1454-1469). */
__attribute__((target_clones("avx2", "default")))
static int any_hit_span(const soa_t *m, const uint8_t *active, int64_t j0, int64_t j1, v3 o, v3 d,
                        float eps, float thr) {
    int any = 0;
#pragma omp simd reduction(| : any)
    for (int64_t j = j0; j < j1; ++j) {
        float e2x = m->e2x[j], e2y = m->e2y[j], e2z = m->e2z[j];
        float e1x = m->e1x[j], e1y = m->e1y[j], e1z = m->e1z[j];
        float hx = d.y * e2z - d.z * e2y, hy = d.z * e2x - d.x * e2z, hz = d.x * e2y - d.y * e2x;
        float a = (hx * e1x + hy * e1y) + hz * e1z;
        a = (a == 0.0f) ? INFINITY : a;
        int hit = fabsf(a) > eps;
        float f = 1.0f / a;
        float sx = o.x - m->v0x[j], sy = o.y - m->v0y[j], sz = o.z - m->v0z[j];
        float u = f * ((sx * hx + sy * hy) + sz * hz);
        hit &= (u >= 0.0f) & (u <= 1.0f);
        float qx = sy * e1z - sz * e1y, qy = sz * e1x - sx * e1z, qz = sx * e1y - sy * e1x;
        float v = f * ((qx * d.x + qy * d.y) + qz * d.z);
        hit &= (v >= 0.0f) & (u + v <= 1.0f);
        float t = f * ((qx * e2x + qy * e2y) + qz * e2z);
        hit &= (t > eps) & (t < thr);
        if (active) hit &= active[j] != 0;
        any |= hit;
    }
    return any;
}

/* a2: _utils.py:1414-1537.  early_exit != 0 stops a ray after the first 512-triangle span that hits
 * (result identical; it is what a tuned CPU implementation would do). */
ORC_API void orc_ray_intersect_any_triangle(int64_t R, int64_t T, const float *o, const float *d,
                                            const float *tri, const uint8_t *active, float eps,
                                            float hit_tol, int early_exit, uint8_t *out) {
    if (T <= 0) { memset(out, 0, (size_t)R); return; }
    soa_t m = soa_build(T, tri);
    float thr = 1.0f - hit_tol;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t r = 0; r < R; ++r) {
        v3 oo = ld3(o + 3 * r), dd = ld3(d + 3 * r);
        int any = 0;
        for (int64_t j0 = 0; j0 < T; j0 += 512) {
            int64_t j1 = j0 + 512 < T ? j0 + 512 : T;
            any |= any_hit_span(&m, active, j0, j1, oo, dd, eps, thr);
            if (any && early_exit) break;
        }
        out[r] = (uint8_t)any;
    }
    soa_free(&m);
}

/* a3: _utils.py:1821-1960 with the reference's tie rule: first minimum inside a batch (:1886),
 * carry kept only if carry_t < new_t across batches (:1865-1868). */
ORC_API void orc_first_triangle_hit_by_ray(int64_t R, int64_t T, const float *o, const float *d,
                                           const float *tri, const uint8_t *active, float eps,
                                           int64_t batch_size, int32_t *idx_out, float *t_out) {
    if (T <= 0) {
        for (int64_t r = 0; r < R; ++r) { idx_out[r] = -1; t_out[r] = INFINITY; }
        return;
    }
    soa_t m = soa_build(T, tri);
    int64_t bs = batch_size <= 0 ? T : (batch_size < T ? batch_size : T);
    if (bs < 1) bs = 1;
    int64_t nb = T / bs, rem = T % bs;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t r = 0; r < R; ++r) {
        v3 oo = ld3(o + 3 * r), dd = ld3(d + 3 * r);
        int32_t ci = -1;
        float ct = INFINITY;
        for (int64_t b = 0; b < nb + (rem > 0); ++b) {
            int64_t j0 = b < nb ? b * bs : T - rem;
            int64_t j1 = b < nb ? j0 + bs : T;
            int32_t bi = 0;
            float bt = INFINITY;
            for (int64_t j = j0; j < j1; ++j) {
                v3 v0 = {m.v0x[j], m.v0y[j], m.v0z[j]};
                v3 e1 = {m.e1x[j], m.e1y[j], m.e1z[j]};
                v3 e2 = {m.e2x[j], m.e2y[j], m.e2z[j]};
                float t;
                int hit = mt_test(oo, dd, v0, e1, e2, eps, &t);
                if (active) hit &= active[j] != 0;
                t = hit ? t : INFINITY;
                if (t < bt) { bt = t; bi = (int32_t)(j - j0); }
            }
            bi = isinf(bt) ? -1 : bi;
            bi += (int32_t)j0;
            if (!(ct < bt)) { ct = bt; ci = bi; }
        }
        int fin = isfinite(ct);
        idx_out[r] = fin ? ci : -1;
        t_out[r] = fin ? ct : INFINITY;
    }
    soa_free(&m);
}

/* a4 with directions given: first hit per ray (batch_size=None), scatter True. out is [B,T]. */
ORC_API void orc_triangles_visible_from_vertex(int64_t B, int64_t n_rays, int64_t T,
                                               const float *vertex, const float *dirs,
                                               const float *tri, const uint8_t *active, float eps,
                                               uint8_t *out) {
    memset(out, 0, (size_t)(B * T));
    if (T <= 0 || n_rays <= 0) return;
    int64_t R = B * n_rays;
    float *o = (float *)malloc(sizeof(float) * 3 * (size_t)R);
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)R);
    float *t = (float *)malloc(sizeof(float) * (size_t)R);
    for (int64_t r = 0; r < R; ++r) memcpy(o + 3 * r, vertex + 3 * (r / n_rays), 3 * sizeof(float));
    orc_first_triangle_hit_by_ray(R, T, o, dirs, tri, active, eps, 0, idx, t);
    for (int64_t r = 0; r < R; ++r)
        if (idx[r] >= 0) out[(r / n_rays) * T + idx[r]] = 1;
    free(o); free(idx); free(t);
}

/* _solver_image_method.py:138-203 for one path; k <= 16. Writes paths[k]. */
static void image_method_one(v3 from, v3 to, int k, const v3 *mv, const v3 *mn, v3 *paths) {
    v3 img[16];
    v3 prev = from;
    for (int i = 0; i < k; ++i) {
        v3 inc = v3sub(prev, mv[i]);
        float c = 2.0f * v3dot(inc, mn[i]);
        v3 r = {prev.x - c * mn[i].x, prev.y - c * mn[i].y, prev.z - c * mn[i].z};
        img[i] = r;
        prev = r;
    }
    prev = to;
    for (int i = k - 1; i >= 0; --i) {
        int ix = isinf(prev.x), iy = isinf(prev.y), iz = isinf(prev.z);
        v3 p0 = {ix ? 0.0f : prev.x, iy ? 0.0f : prev.y, iz ? 0.0f : prev.z};
        v3 u = v3sub(img[i], p0);
        v3 w = v3sub(mv[i], p0);
        float un = v3dot(u, mn[i]);
        float vn = v3dot(w, mn[i]);
        int par = un == 0.0f;
        if (par) un = 1.0f;
        float t = vn / un;
        v3 r = {p0.x + u.x * t, p0.y + u.y * t, p0.z + u.z * t};
        if (par && vn != 0.0f) { r.x = INFINITY; r.y = INFINITY; r.z = INFINITY; }
        if (ix) r.x = INFINITY;
        if (iy) r.y = INFINITY;
        if (iz) r.z = INFINITY;
        paths[i] = r;
        prev = r;
    }
}

/* a7 flat: from/to [N,3], mv/mn [N,k,3] → out [N,k,3]. */
ORC_API void orc_image_method(int64_t N, int k, const float *from, const float *to, const float *mv,
                              const float *mn, float *out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        v3 a[16], b[16], p[16];
        for (int j = 0; j < k; ++j) { a[j] = ld3(mv + (i * k + j) * 3); b[j] = ld3(mn + (i * k + j) * 3); }
        image_method_one(ld3(from + 3 * i), ld3(to + 3 * i), k, a, b, p);
        for (int j = 0; j < k; ++j) { out[(i * k + j) * 3] = p[j].x; out[(i * k + j) * 3 + 1] = p[j].y; out[(i * k + j) * 3 + 2] = p[j].z; }
    }
}

static inline float sgnf(float x) { return (x > 0.0f) - (x < 0.0f); } /* NaN → 0, like comparisons */

/* a9: _solvers.py:514-770, non-smoothing branch, blockage = pure-JAX any-hit over the whole mesh
 * for every segment of every candidate (dense, as the reference evaluates it).
 * early_exit: 0 = dense like the reference's fori_loop; 1 = a ray stops at the first 512-triangle span that
 * hits; 2 = in addition a candidate stops at its first blocked segment.  Same outputs in all three.
 * tri_mask may be NULL.  stage_out (may be NULL) receives 5 flag bytes per path:
 * inside, same_side, blocked, too_small, finite.  tests_done (may be NULL) receives the number of
 * ray-triangle tests evaluated. */
ORC_API void orc_trace_path_candidates(int64_t V, int64_t T, const float *verts, const int32_t *tris,
                                       const uint8_t *tri_mask, int assume_quads, int64_t ntx,
                                       const float *tx, int64_t nrx, const float *rx, int64_t C,
                                       int k, const int32_t *cand, float eps, float hit_tol,
                                       float min_len, int early_exit, float *out_vertices,
                                       int32_t *out_objects, uint8_t *out_mask, uint8_t *stage_out,
                                       int64_t *tests_done) {
    (void)V;
    float *tri = (float *)malloc(sizeof(float) * 9 * (size_t)(T > 0 ? T : 1));
    float *nrm = (float *)malloc(sizeof(float) * 3 * (size_t)(T > 0 ? T : 1));
    for (int64_t j = 0; j < T; ++j) {
        for (int c = 0; c < 3; ++c) memcpy(tri + 9 * j + 3 * c, verts + 3 * (int64_t)tris[3 * j + c], 12);
        v3 v0 = ld3(tri + 9 * j), v1 = ld3(tri + 9 * j + 3), v2 = ld3(tri + 9 * j + 6);
        v3 n = v3cross(v3sub(v1, v0), v3sub(v2, v1)); /* _mesh.py:953-956 */
        float len = sqrtf(v3dot(n, n));
        if (len == 0.0f) len = 1.0f;               /* _utils.py:66-72 */
        nrm[3 * j] = n.x / len; nrm[3 * j + 1] = n.y / len; nrm[3 * j + 2] = n.z / len;
    }
    soa_t m = soa_build(T, tri);
    float thr = 1.0f - hit_tol;
    int q = assume_quads ? 2 : 1;
    int64_t P = ntx * nrx * C;
    int64_t ntests = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : ntests)
    for (int64_t p = 0; p < P; ++p) {
        int64_t c = p % C, irx = (p / C) % nrx, itx = p / (C * nrx);
        v3 mv[16], mn[16], full[18];
        int active = 1;
        for (int i = 0; i < k; ++i) {
            int32_t ti = cand[c * k + i];
            mv[i] = ld3(tri + 9 * (int64_t)ti);
            mn[i] = ld3(nrm + 3 * (int64_t)ti);
            if (tri_mask) for (int s = 0; s < q; ++s) active &= tri_mask[ti + s] != 0;
        }
        full[0] = ld3(tx + 3 * itx);
        full[k + 1] = ld3(rx + 3 * irx);
        image_method_one(full[0], full[k + 1], k, mv, mn, full + 1);
        int inside = 1, same = 1, blocked = 0, small = 0, finite = 1;
        for (int i = 0; i <= k; ++i) {
            v3 o = full[i], d = v3sub(full[i + 1], full[i]);
            if (i < k) {
                int any = 0;
                for (int s = 0; s < q; ++s) {
                    int64_t ti = cand[c * k + i] + s;
                    v3 v0 = ld3(tri + 9 * ti), v1 = ld3(tri + 9 * ti + 3), v2 = ld3(tri + 9 * ti + 6);
                    float t;
                    any |= mt_test(o, d, v0, v3sub(v1, v0), v3sub(v2, v0), eps, &t);
                }
                inside &= any;
                float dp = v3dot(v3sub(full[i], mv[i]), mn[i]);
                float dn = v3dot(v3sub(full[i + 2], mv[i]), mn[i]);
                same &= (sgnf(dp) == sgnf(dn)) & !isnan(dp) & !isnan(dn); /* sign(NaN)=NaN != NaN */
            }
            small |= v3dot(d, d) < min_len;
            if (T > 0) {
                int any = 0;
                for (int64_t j0 = 0; j0 < T; j0 += 512) {
                    int64_t j1 = j0 + 512 < T ? j0 + 512 : T;
                    any |= any_hit_span(&m, tri_mask, j0, j1, o, d, eps, thr);
                    ntests += j1 - j0;
                    if (any && early_exit) break;
                }
                blocked |= any;
                /* early_exit >= 2: a blocked candidate is invalid whatever its other segments do, so a
                 * tuned CPU implementation stops here (mask identical; the per-stage flags would not be,
                 * so this level is only honoured when they are not requested) */
                if (blocked && early_exit >= 2 && !stage_out) break;
            }
        }
        for (int i = 0; i < k + 2; ++i) finite &= isfinite(full[i].x) && isfinite(full[i].y) && isfinite(full[i].z);
        float *ov = out_vertices + p * (k + 2) * 3;
        for (int i = 0; i < k + 2; ++i) {
            ov[3 * i] = finite ? full[i].x : 0.0f;
            ov[3 * i + 1] = finite ? full[i].y : 0.0f;
            ov[3 * i + 2] = finite ? full[i].z : 0.0f;
        }
        int32_t *oo = out_objects + p * (k + 2);
        oo[0] = (int32_t)itx;
        for (int i = 0; i < k; ++i) oo[i + 1] = cand[c * k + i];
        oo[k + 1] = (int32_t)irx;
        out_mask[p] = (uint8_t)(inside & same & !blocked & !small & finite & active);
        if (stage_out) {
            uint8_t *so = stage_out + 5 * p;
            so[0] = (uint8_t)inside; so[1] = (uint8_t)same; so[2] = (uint8_t)blocked;
            so[3] = (uint8_t)small; so[4] = (uint8_t)finite;
        }
    }
    if (tests_done) *tests_done = ntests;
    soa_free(&m);
    free(tri);
    free(nrm);
}
