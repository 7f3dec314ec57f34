// Exact conservative culling for the path-blockage test (any-hit of the k+1 segments of a path
// candidate against every active triangle; reference differt/src/differt/geometry/_solvers.py:655-680
// → _utils.py:1414-1537).
//
// The blockage result is an OR over (segment, triangle) pairs of the reference's fp32 Möller–Trumbore
// decision.  A pair may be skipped only if that decision is PROVEN to be "no hit" — including the
// noise-dominated cases (a segment nearly parallel to a triangle's plane) in which the fp32 test
// reports hits far away from the triangle, which is why a plain bounding-volume test is not enough.
//
// ---- The bound --------------------------------------------------------------------------------------
// Notation: u = 2^-24; data o, d, v0, e1, e2 (fp32, exact as stored); s = o - v0 (real);
// h = d x e2, a = h.e1, Nu = s.h, q = s x e1, Nv = q.d, Nt = q.e2 (real); hats = what
// _utils.py:1263-1322 computes in fp32 without FMA.  Adjugate identity (holds for ANY a, also a = 0):
//        a s = Nu e1 + Nv e2 - Nt d.
// Standard forward error bounds of the cross/dot products give, with W = |d||e1||e2|,
//        |â - a| <= 8u W,  |N̂u - Nu| <= 9u |s||d||e2|,  |N̂v - Nv| <= 9u |s||e1||d|,  |N̂t - Nt| <= 9u |s||e1||e2|,
// and û = (N̂u/â)(1+δ), |δ| <= 2.01u (correctly rounded reciprocal, one multiplication), same for v̂, t̂.
// A reported hit has |â| > eps > 0, û, v̂ in [0,1], û + v̂ <= 1 + u, t̂ in (eps, thr) ⊂ (0, 1) (we only
// cull when FLT_MIN <= eps and 0 < thr <= 1), hence |N̂u| <= |â|(1 + 2.01u) etc.  Put P1 = o + t̂ d (a point
// OF the segment) and P2 = v0 + û e1 + v̂ e2 (within u(|e1|+|e2|) of the triangle).  Then
//        â (P2 - P1) = (a - â) s + [(N̂u(1+δu) - Nu) e1 + (N̂v(1+δv) - Nv) e2 - (N̂t(1+δt) - Nt) d]
//   ⇒   |P2 - P1| <= 35u W|s| / |â| + 2.02u (|e1| + |e2| + |d|).
// With rho = |a| / W = |d^.n^| sin(theta)  (n^ = unit normal, theta = angle between e1 and e2) and
// |â| >= |a| - 8u W:
//   (*)  reported hit  ⇒  dist(segment, triangle) <= 35u |s| / (rho - 8u) + 3.1u (|e1| + |e2| + |d|)   (rho > 8u).
// Overflow to inf/NaN anywhere makes the reference's comparisons false (no hit), so (*) cannot be
// violated by overflow; underflow needs products below 2^-126, excluded by the guards below.
//
// ---- The test ---------------------------------------------------------------------------------------
// A NODE (a group of 8 triangles, or 8 nodes of the level below) stores its bounding box (centre, half extent), up to three
// unit axes such that every triangle normal is within angle alpha of ±one axis (sin/cos alpha), the
// smallest sin(theta), E >= |e1|+|e2| and R >= every |coordinate|.  For a segment (o, d):
//     g  = sin_theta_min (min_k |d^.c_k| cos_alpha - sin_alpha) - 2e-5      <= rho - 8u for every triangle
//     S  = sqrt(3) max_k(|o_k - ctr_k| + half_k) + 1e-7 (R_seg + R_node)    >= |o - v0| for every triangle
//     m  = 6e-6 S / g + 1e-6 E + 4e-6 (R_seg + R_node)                      >= 2.8 x the bound (*) + rounding of
//                                                                              the box arithmetic itself
// and the node is culled for that segment iff g > 0 and the segment misses the box inflated by m (slab
// test on t in [0,1] with 1e-5 slack in t; reciprocals of |d_k| < 2^-100 are replaced by ±FLT_MAX, which
// only enlarges the slab interval as long as m >= 2^-100).  Degenerate triangles (sin(theta) < 2^-10, an
// edge shorter than 2^-20, non-finite data) put sin_theta_min = 0: such a node is never culled; neither is
// a segment shorter than 2^-30 nor data with R_seg + R_node outside [1e-20, 1e9] — with those guards
// W >= 2^-70 and the absolute errors of products that underflow (<= 2^-150 each; <= 2^-147 (R + 1) on a^, N^u, N^v, N^t)
// add at most 5e-23 (R_seg + R_node)^2 / g to (*), which the 1e-7 (R_seg + R_node) term inside S covers
// (6e-13 (R_seg + R_node) / g in m).  The segment guard sits at 2^-30 and not higher because rounding residues ARE
// segments: a path that reflects twice on one plane has two vertices one ulp apart (|d| ~ 1e-7 at city scale), and
// an un-cullable segment walks the whole hierarchy.  Segments with d = 0 or
// non-finite vertices can never hit (a = 0 → inf → t = 0, or NaN comparisons) and are dropped up front.
// Everything the cull lets through is evaluated by the same mt_any_fast / mt_exact as everywhere else.
//
// ---- Exactly axis-aligned triangles -----------------------------------------------------------------
// A segment lying IN the plane of a triangle has rho = 0: (*) says nothing and the node is descended into —
// and a path that reflects twice in a row on one plane (the two triangles of a wall, the floor slabs and the
// ground of a city model: 3 % of the bench's candidates) has such a segment with respect to EVERY triangle of
// that plane, so it would walk the whole hierarchy.  For a triangle whose edges both have an exactly zero
// j-th component (normal = +-e_j) and a segment with d_j == 0 exactly, the reference's own arithmetic is
// exact where it matters: h = d x e2 has h_i = d_j e2_m - d_m e2_j = 0 for both i != j (products with an
// exact zero are exact zeros), so a = h.e1 = 0 + 0 + h_j * 0 is exactly 0 (NaN if h_j overflowed), a is replaced by
// inf, f = 0, t = 0 and `t > eps` fails: NO hit, whatever the distance.  A node whose triangles are ALL of
// that kind (flag: sign bit of c1.w; its axes are stored as exact unit vectors) drops, for such a segment, the
// axes along which d is exactly zero from the min over |d^.c_k|: those triangles cannot be hit, the others are
// covered by (*) as before.  make_seg_cull keeps d^_j == 0 <=> d_j == 0 so that the node test sees it.
#pragma once

#include "common.cuh"

namespace drt {

constexpr int kCullGroup = 8;  // triangles per leaf node (group)

struct __align__(16) CullNode {  // 80 bytes
    float4 ctr;   // xyz = box centre,      w = cos(alpha)   (lower bound)
    float4 half;  // xyz = box half extent, w = sin(alpha)   (upper bound); half.x < 0 → empty node
    float4 c0;    // xyz = axis 0,          w = min sin(theta) (lower bound; 0 → never culled)
    float4 c1;    // xyz = axis 1,          |w| = R: max |coordinate| of the node's triangles; sign bit set: every
                  //                        triangle is exactly axis-aligned and the axes are exact unit vectors
    float4 c2;    // xyz = axis 2,          w = E: max (|e1| + |e2|)
};
static_assert(sizeof(CullNode) == 80, "CullNode must be 80 bytes");

struct SegCull {  // per-segment constants of the node test (warp-uniform)
    float3 o, dhat, inv;
    float rseg;
};

// MUFU.RCP alone: relative error <= 2^-22, for normal inputs (every use below is guarded to |x| >= 2^-100).  The
// node test has slack for it: 2e-5 absolute in g (|d^.c| <= 1) and 1e-5 relative in the slab parameters.
__device__ __forceinline__ float rcp_fast(const float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ SegCull make_seg_cull(const float3 o, const float3 d) {
    SegCull s;
    s.o = o;
    const float len = sqrtf(dot3(d, d));
    const float rl = len >= 9.313225746e-10f ? rcp_fast(len) : 0.0f;  // |d| < 2^-30: d^ = 0 → g < 0 → never culled
    s.dhat = make_float3(d.x * rl, d.y * rl, d.z * rl);
    // d^_j == 0 <=> d_j == 0 (header, axis-aligned triangles): a non-zero component that underflowed (or rl = 0)
    // becomes a tiny non-zero value, which only makes g smaller
    if (d.x != 0.0f && s.dhat.x == 0.0f) s.dhat.x = 1e-37f;
    if (d.y != 0.0f && s.dhat.y == 0.0f) s.dhat.y = 1e-37f;
    if (d.z != 0.0f && s.dhat.z == 0.0f) s.dhat.z = 1e-37f;
    const float kTiny = 7.888609052e-31f;  // 2^-100
    s.inv.x = fabsf(d.x) >= kTiny ? rcp_fast(d.x) : copysignf(3.402823466e38f, d.x);
    s.inv.y = fabsf(d.y) >= kTiny ? rcp_fast(d.y) : copysignf(3.402823466e38f, d.y);
    s.inv.z = fabsf(d.z) >= kTiny ? rcp_fast(d.z) : copysignf(3.402823466e38f, d.z);
    const float3 e = add3(o, d);
    s.rseg = fmaxf(fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(e.x))),
                   fmaxf(fabsf(e.y), fabsf(e.z)));
    return s;
}

// The same for a RAY (first-hit queries: t > eps only, no upper bound).  (*) gains the term
// 2.02u t̂ |d|, and t̂ |d| = |P1 - o| <= |P2 - o| + |P2 - P1| <= S + m, i.e. 2.02u (S + m): inside the
// slack of the 6e-6 S / g term (100u S / g against the 37u S / g needed).  R_seg only has to bound the
// coordinates of o (the slab arithmetic never forms o + d).
__device__ __forceinline__ SegCull make_ray_cull(const float3 o, const float3 d) {
    SegCull s = make_seg_cull(o, d);
    s.rseg = fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fabsf(o.z));
    return s;
}

// true iff the node is PROVEN to contain no triangle the reference's test would report as hit by s at
// a parameter t <= tmax (1 for a segment; the best distance so far, or +inf, for a first-hit ray)
__device__ __forceinline__ bool node_culled(const SegCull &s, const CullNode &n, const float tmax_seg = 1.0f) {
    // (an empty node — only never-hit records: half.x < 0, sin_theta_min = 0 — is reported as culled where g fails;
    // no early test on half.x: it would put a second dependent L1 round trip in front of every node test)
    float p0 = fabsf(__fmaf_rn(s.dhat.x, n.c0.x, __fmaf_rn(s.dhat.y, n.c0.y, s.dhat.z * n.c0.z)));
    float p1 = fabsf(__fmaf_rn(s.dhat.x, n.c1.x, __fmaf_rn(s.dhat.y, n.c1.y, s.dhat.z * n.c1.z)));
    float p2 = fabsf(__fmaf_rn(s.dhat.x, n.c2.x, __fmaf_rn(s.dhat.y, n.c2.y, s.dhat.z * n.c2.z)));
    if (__float_as_int(n.c1.w) < 0) {
        // every axis is an exact +-e_j: p_k = |d^_j| exactly, and p_k == 0 <=> d_j == 0 <=> the triangles of that axis
        // cannot be hit (header): the axis does not bind (any value >= 1 >= |d^.c| stands for "no constraint")
        p0 = p0 == 0.0f ? 2.0f : p0;
        p1 = p1 == 0.0f ? 2.0f : p1;
        p2 = p2 == 0.0f ? 2.0f : p2;
    }
    const float pmin = fminf(fminf(p0, p1), p2);
    const float g = __fmaf_rn(n.c0.w, __fmaf_rn(pmin, n.ctr.w, -n.half.w), -2e-5f);
    if (!(g > 0.0f)) return n.half.x < 0.0f;  // grazing, degenerate or NaN: cannot be proven (empty: c0.w = 0 → here)
    const float sx = fabsf(s.o.x - n.ctr.x) + n.half.x;
    const float sy = fabsf(s.o.y - n.ctr.y) + n.half.y;
    const float sz = fabsf(s.o.z - n.ctr.z) + n.half.z;
    const float rsum = s.rseg + fabsf(n.c1.w);
    if (!(rsum >= 1e-20f && rsum <= 1e9f)) return false;  // keep clear of underflow / overflow (header)
    const float S = __fmaf_rn(1.7320509f, fmaxf(fmaxf(sx, sy), sz), 1e-7f * rsum);
    const float m = __fmaf_rn(6e-6f * S, __fdividef(1.0f, g) * 1.0001f, __fmaf_rn(1e-6f, n.c2.w, 4e-6f * rsum));
    const float hx = n.half.x + m, hy = n.half.y + m, hz = n.half.z + m;
    const float ax = ((n.ctr.x - hx) - s.o.x) * s.inv.x, bx = ((n.ctr.x + hx) - s.o.x) * s.inv.x;
    const float ay = ((n.ctr.y - hy) - s.o.y) * s.inv.y, by = ((n.ctr.y + hy) - s.o.y) * s.inv.y;
    const float az = ((n.ctr.z - hz) - s.o.z) * s.inv.z, bz = ((n.ctr.z + hz) - s.o.z) * s.inv.z;
    const float tmin = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));
    const float tmax = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tmax_seg));
    // 1e-5 relative slack in t (the computed parameters carry a relative error of a few u)
    return tmin > __fmaf_rn(tmax, 1e-5f, tmax) + 1e-30f;
}

// The 8-ary hierarchy of the per-thread traversal (path_walk_kernel): level 0 is the top (<= 8 nodes,
// children of a virtual root), the last level holds the groups; node i of level l has the children
// 8i .. 8i+7 of level l+1.  Offsets are in nodes from the start of the `walk` array.
constexpr int kWalkFan = 8;
constexpr int kWalkMaxLevels = 8;
struct WalkLevels {
    int num_levels;
    int offset[kWalkMaxLevels];
    int size[kWalkMaxLevels];
};

size_t cull_workspace_bytes(int64_t records);
// Builds the spatially ordered pack and its node levels from a packed mesh of `records` records
// (a multiple of kTile).  Layout inside `ws`: see CullLayout.
struct CullLayout {
    size_t pack, groups, walk, bounds, keys, total;
    int64_t num_groups;
    WalkLevels levels;  // groups are ALSO the last level of `walk` (stored once, inside `walk`)
};
CullLayout cull_layout(int64_t records);
int cull_build(cudaStream_t s, int64_t records, const Tri48 *pack_in, unsigned char *ws, const CullLayout &l,
               void *sort_ws, size_t sort_bytes);

}  // namespace drt
