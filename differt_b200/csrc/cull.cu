// Builds the structure behind the exact conservative cull of the blockage test (cull.cuh): the packed
// triangles in Morton order of their centroids, one CullNode per group of 8 consecutive triangles and
// the levels of an 8-ary hierarchy above them.  Stateless like everything else in the library: rebuilt per call from the
// packed mesh (three tiny kernels + one CUB radix sort; ~30 µs for 10 000 triangles).
#include "cull.cuh"

namespace drt {

int drt_sort_records_by_keys(drt_stream_t stream, int64_t n, const void *pack_in, const uint32_t *keys,
                             void *workspace, size_t workspace_bytes, void *pack_out);  // pack_sort.cu

namespace {

__device__ __forceinline__ uint32_t spread10(uint32_t v) {  // 10 bits → every third bit
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__device__ __forceinline__ bool record_live(const Tri &t) {  // false for never-hit records (NaN origin)
    return t.v0.x == t.v0.x;
}

// Geometry of one triangle as the node builders need it.  `good` = every quantity finite, both edges
// >= 2^-20 and sin(theta) >= 2^-10; otherwise the triangle makes its node un-cullable.
struct TriInfo {
    float3 lo, hi, nrm;
    float r, e, sin_theta, beta;
    bool good;
    int aligned;  // j if both edges have an exactly zero j-th component (normal = +-e_j, cull.cuh), else -1
};

__device__ __forceinline__ TriInfo tri_info(const Tri &t) {
    TriInfo i;
    const float3 v1 = add3(t.v0, t.e1), v2 = add3(t.v0, t.e2);
    i.lo = make_float3(fminf(t.v0.x, fminf(v1.x, v2.x)), fminf(t.v0.y, fminf(v1.y, v2.y)),
                       fminf(t.v0.z, fminf(v1.z, v2.z)));
    i.hi = make_float3(fmaxf(t.v0.x, fmaxf(v1.x, v2.x)), fmaxf(t.v0.y, fmaxf(v1.y, v2.y)),
                       fmaxf(t.v0.z, fmaxf(v1.z, v2.z)));
    i.r = fmaxf(fmaxf(fmaxf(fabsf(i.lo.x), fabsf(i.lo.y)), fmaxf(fabsf(i.lo.z), fabsf(i.hi.x))),
                fmaxf(fabsf(i.hi.y), fabsf(i.hi.z)));
    const float l1 = sqrtf(dot3(t.e1, t.e1)), l2 = sqrtf(dot3(t.e2, t.e2));
    i.e = (l1 + l2) * 1.0001f;
    const float3 n = cross3(t.e1, t.e2);
    const float ln = sqrtf(dot3(n, n));
    const float st = ln / (l1 * l2);  // sin(theta), relative error of a few u plus 2.9u absolute (header)
    i.sin_theta = st * 0.9999f - 1e-6f;
    i.good = isfinite(i.r) && isfinite(l1) && isfinite(l2) && isfinite(ln) && l1 >= 9.5367431640625e-7f &&
             l2 >= 9.5367431640625e-7f && l1 * l2 >= 1e-30f && st >= 9.765625e-4f && st <= 1.001f;
    const float rn = i.good ? 1.0f / ln : 0.0f;
    i.nrm = make_float3(n.x * rn, n.y * rn, n.z * rn);
    // angle between the computed and the true normal: |N_c - N| <= 2.9u |e1||e2| → asin(2.9u / sin theta)
    i.beta = i.good ? 2.4e-7f / st : 0.0f;
    i.aligned = -1;
    if (i.good) {
        const bool zx = t.e1.x == 0.0f && t.e2.x == 0.0f, zy = t.e1.y == 0.0f && t.e2.y == 0.0f,
                   zz = t.e1.z == 0.0f && t.e2.z == 0.0f;
        if (int(zx) + int(zy) + int(zz) == 1) {
            i.aligned = zx ? 0 : (zy ? 1 : 2);
            i.nrm = make_float3(zx ? 1.f : 0.f, zy ? 1.f : 0.f, zz ? 1.f : 0.f);  // the exact axis (sign is irrelevant)
        }
    }
    return i;
}

// Running description of the normals seen so far: up to three axes, every normal within asin(sa) of
// ±one of them.  A new axis is opened while fewer than three exist and the normal is more than
// asin(0.05) away from all of them; otherwise the cone grows.
struct AxisSet {
    float3 c[3];
    int n = 0;
    float sa = 0.0f;
    __device__ __forceinline__ void add(const float3 v, const float extra) {
        float best = 2.0f;
        for (int k = 0; k < n; ++k) {
            const float3 x = cross3(v, c[k]);
            best = fminf(best, sqrtf(dot3(x, x)));
        }
        if (n == 0 || (best > 0.05f && n < 3)) {
            c[n++] = v;
            best = 0.0f;
        }
        sa = fmaxf(sa, best + extra);
    }
    __device__ __forceinline__ void finish(CullNode &node, const float sin_theta_min, const bool cullable) const {
        const float3 a0 = n > 0 ? c[0] : make_float3(1.f, 0.f, 0.f);
        const float3 a1 = n > 1 ? c[1] : a0, a2 = n > 2 ? c[2] : a0;
        // sa bounds the sine of the deviation; beyond ~0.7 the cone is useless anyway
        const float s = fminf(sa * 1.001f + 2e-6f, 1.0f);
        node.ctr.w = sqrtf(fmaxf(1.0f - s * s, 0.0f)) * 0.9999f;
        node.half.w = s;
        node.c0 = make_float4(a0.x, a0.y, a0.z, cullable && n > 0 ? fmaxf(sin_theta_min, 0.0f) : 0.0f);
        node.c1.x = a1.x, node.c1.y = a1.y, node.c1.z = a1.z;
        node.c2.x = a2.x, node.c2.y = a2.y, node.c2.z = a2.z;
    }
};

// scene bounds over the live records: single block → bounds[6]
__global__ void __launch_bounds__(1024) cull_bounds_kernel(int64_t n, const Tri48 *__restrict__ pack,
                                                           float *__restrict__ bounds) {
    __shared__ float sm[6][32];
    float lo[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, hi[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
        const Tri t = unpack(pack[j].a, pack[j].b, pack[j].c);
        if (!record_live(t)) continue;
        const float3 c = make_float3(t.v0.x + (t.e1.x + t.e2.x) * (1.0f / 3.0f), t.v0.y + (t.e1.y + t.e2.y) * (1.0f / 3.0f),
                                     t.v0.z + (t.e1.z + t.e2.z) * (1.0f / 3.0f));
        if (!(isfinite(c.x) && isfinite(c.y) && isfinite(c.z))) continue;
        lo[0] = fminf(lo[0], c.x), lo[1] = fminf(lo[1], c.y), lo[2] = fminf(lo[2], c.z);
        hi[0] = fmaxf(hi[0], c.x), hi[1] = fmaxf(hi[1], c.y), hi[2] = fmaxf(hi[2], c.z);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        for (int off = 16; off > 0; off >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(kFull, lo[k], off));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(kFull, hi[k], off));
        }
        if (lane == 0) sm[k][w] = lo[k], sm[3 + k][w] = hi[k];
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float a = lane < int(blockDim.x >> 5) ? sm[k][lane] : CUDART_INF_F;
            float b = lane < int(blockDim.x >> 5) ? sm[3 + k][lane] : -CUDART_INF_F;
            for (int off = 16; off > 0; off >>= 1) {
                a = fminf(a, __shfl_xor_sync(kFull, a, off));
                b = fmaxf(b, __shfl_xor_sync(kFull, b, off));
            }
            if (lane == 0) bounds[k] = a, bounds[3 + k] = b;
        }
    }
}

// Sort key (descending sort): live, well-shaped triangles by Morton code of the centroid; degenerate
// ones after them (key 1: they end up in the same few, un-cullable nodes); never-hit records last.
__global__ void cull_keys_kernel(int64_t n, const Tri48 *__restrict__ pack, const float *__restrict__ bounds,
                                 uint32_t *__restrict__ keys) {
    const int64_t j = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (j >= n) return;
    const Tri t = unpack(pack[j].a, pack[j].b, pack[j].c);
    uint32_t key = 0;
    if (record_live(t)) {
        key = 1;
        if (tri_info(t).good) {
            const float c[3] = {t.v0.x + (t.e1.x + t.e2.x) * (1.0f / 3.0f), t.v0.y + (t.e1.y + t.e2.y) * (1.0f / 3.0f),
                                t.v0.z + (t.e1.z + t.e2.z) * (1.0f / 3.0f)};
            uint32_t q[3];
            // ONE scale for the three axes (the largest extent): a city is 20 x wider than tall, and cells
            // that are cubes in space keep the triangles of a building together
            const float ext = fmaxf(fmaxf(bounds[3] - bounds[0], bounds[4] - bounds[1]), bounds[5] - bounds[2]);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float x = ext > 0.0f ? (c[k] - bounds[k]) / ext : 0.0f;
                x = fminf(fmaxf(x * 1024.0f, 0.0f), 1023.0f);  // NaN → 0
                q[k] = uint32_t(x);
            }
            key = 2u + ((spread10(q[0]) << 2) | (spread10(q[1]) << 1) | spread10(q[2]));
        }
    }
    keys[j] = key;
}

// one thread per group of kCullGroup consecutive records of the ordered pack
__global__ void cull_group_nodes_kernel(int64_t num_groups, const Tri48 *__restrict__ pack,
                                        CullNode *__restrict__ nodes) {
    const int64_t gidx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (gidx >= num_groups) return;
    float3 lo = make_float3(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F);
    float3 hi = make_float3(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
    float r = 0.0f, e = 0.0f, st = 1.0f;
    bool any = false, cullable = true, aligned = true;
    AxisSet axes;
    for (int k = 0; k < kCullGroup; ++k) {
        const Tri48 rec = pack[gidx * kCullGroup + k];
        const Tri t = unpack(rec.a, rec.b, rec.c);
        if (!record_live(t)) continue;
        any = true;
        const TriInfo i = tri_info(t);
        if (!i.good) {
            cullable = false;
            continue;
        }
        lo = make_float3(fminf(lo.x, i.lo.x), fminf(lo.y, i.lo.y), fminf(lo.z, i.lo.z));
        hi = make_float3(fmaxf(hi.x, i.hi.x), fmaxf(hi.y, i.hi.y), fmaxf(hi.z, i.hi.z));
        r = fmaxf(r, i.r), e = fmaxf(e, i.e), st = fminf(st, i.sin_theta);
        aligned = aligned && i.aligned >= 0;
        axes.add(i.nrm, i.beta);
    }
    CullNode node;
    if (!any) {  // only never-hit records: always culled
        node.ctr = make_float4(0.f, 0.f, 0.f, 1.f);
        node.half = make_float4(-1.f, -1.f, -1.f, 0.f);
        node.c0 = make_float4(1.f, 0.f, 0.f, 0.f);
        node.c1 = make_float4(1.f, 0.f, 0.f, 0.f);
        node.c2 = make_float4(1.f, 0.f, 0.f, 0.f);
    } else {
        if (!cullable || !(lo.x <= hi.x)) {  // a degenerate triangle: the node is never culled
            lo = hi = make_float3(0.f, 0.f, 0.f);
            cullable = false;
        }
        node.ctr = make_float4(0.5f * lo.x + 0.5f * hi.x, 0.5f * lo.y + 0.5f * hi.y, 0.5f * lo.z + 0.5f * hi.z, 0.f);
        // half extent measured from the rounded centre, rounded up
        node.half = make_float4(fmaxf(hi.x - node.ctr.x, node.ctr.x - lo.x) * 1.000001f,
                                fmaxf(hi.y - node.ctr.y, node.ctr.y - lo.y) * 1.000001f,
                                fmaxf(hi.z - node.ctr.z, node.ctr.z - lo.z) * 1.000001f, 0.f);
        axes.finish(node, st, cullable);
        node.c1.w = (aligned && cullable) ? -r : r;  // r > 0 for a cullable node of finite triangles unless all sit at the origin
        node.c2.w = e;
    }
    nodes[gidx] = node;
}

// one thread per parent of `fan` consecutive nodes of the level below: union of the boxes, axes
// re-clustered from the children's axes (a normal within asin(sa_g) of a group axis that is within asin(x) of a tile axis is
// within asin(x) + asin(sa_g) of it, and sin(p + q) <= sin p + sin q)
__global__ void cull_tile_nodes_kernel(int64_t num_tiles, int64_t num_groups, const CullNode *__restrict__ groups,
                                       CullNode *__restrict__ tiles, const int fan) {
    const int64_t tidx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (tidx >= num_tiles) return;
    float3 lo = make_float3(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F);
    float3 hi = make_float3(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
    float r = 0.0f, e = 0.0f, st = 1.0f;
    bool any = false, cullable = true, aligned = true;
    AxisSet axes;
    for (int k = 0; k < fan; ++k) {
        const int64_t gidx = tidx * fan + k;
        if (gidx >= num_groups) break;
        const CullNode g = groups[gidx];
        if (g.half.x < 0.0f) continue;  // empty group
        any = true;
        if (!(g.c0.w > 0.0f)) {
            cullable = false;
            continue;
        }
        lo = make_float3(fminf(lo.x, g.ctr.x - g.half.x), fminf(lo.y, g.ctr.y - g.half.y), fminf(lo.z, g.ctr.z - g.half.z));
        hi = make_float3(fmaxf(hi.x, g.ctr.x + g.half.x), fmaxf(hi.y, g.ctr.y + g.half.y), fmaxf(hi.z, g.ctr.z + g.half.z));
        r = fmaxf(r, fabsf(g.c1.w)), e = fmaxf(e, g.c2.w), st = fminf(st, g.c0.w);
        aligned = aligned && __float_as_int(g.c1.w) < 0;
        axes.add(make_float3(g.c0.x, g.c0.y, g.c0.z), g.half.w);
        axes.add(make_float3(g.c1.x, g.c1.y, g.c1.z), g.half.w);
        axes.add(make_float3(g.c2.x, g.c2.y, g.c2.z), g.half.w);
    }
    CullNode node;
    if (!any) {
        node.ctr = make_float4(0.f, 0.f, 0.f, 1.f);
        node.half = make_float4(-1.f, -1.f, -1.f, 0.f);
        node.c0 = make_float4(1.f, 0.f, 0.f, 0.f);
        node.c1 = make_float4(1.f, 0.f, 0.f, 0.f);
        node.c2 = make_float4(1.f, 0.f, 0.f, 0.f);
    } else {
        if (!cullable || !(lo.x <= hi.x)) {
            lo = hi = make_float3(0.f, 0.f, 0.f);
            cullable = false;
        }
        node.ctr = make_float4(0.5f * lo.x + 0.5f * hi.x, 0.5f * lo.y + 0.5f * hi.y, 0.5f * lo.z + 0.5f * hi.z, 0.f);
        node.half = make_float4(fmaxf(hi.x - node.ctr.x, node.ctr.x - lo.x) * 1.000001f + 1e-6f * r,
                                fmaxf(hi.y - node.ctr.y, node.ctr.y - lo.y) * 1.000001f + 1e-6f * r,
                                fmaxf(hi.z - node.ctr.z, node.ctr.z - lo.z) * 1.000001f + 1e-6f * r, 0.f);
        axes.finish(node, st, cullable);
        node.c1.w = (aligned && cullable) ? -(r * 1.000001f) : r * 1.000001f;
        node.c2.w = e;
    }
    tiles[tidx] = node;
}

inline size_t a256(size_t x) { return (x + 255) & ~size_t(255); }

}  // namespace

CullLayout cull_layout(int64_t records) {
    CullLayout l{};
    const int64_t n = records > 0 ? records : 1;
    l.num_groups = (n + kCullGroup - 1) / kCullGroup;
    size_t off = 0;
    // levels of the 8-ary hierarchy, bottom-up sizes, stored top-first
    int sizes[kWalkMaxLevels], nb = 0;
    int64_t cur = l.num_groups;
    sizes[nb++] = int(cur);
    while (cur > kWalkFan && nb < kWalkMaxLevels) {
        cur = (cur + kWalkFan - 1) / kWalkFan;
        sizes[nb++] = int(cur);
    }
    l.levels.num_levels = nb;
    int64_t walk_nodes = 0;
    for (int i = 0; i < nb; ++i) {
        l.levels.size[i] = sizes[nb - 1 - i];
        l.levels.offset[i] = int(walk_nodes);
        walk_nodes += sizes[nb - 1 - i];
    }
    for (int i = nb; i < kWalkMaxLevels; ++i) l.levels.size[i] = l.levels.offset[i] = 0;
    l.pack = off, off += a256(size_t(n) * sizeof(Tri48));
    l.walk = off, off += a256(size_t(walk_nodes) * sizeof(CullNode));
    l.groups = l.walk + size_t(l.levels.offset[nb - 1]) * sizeof(CullNode);  // the last level
    l.bounds = off, off += 256;
    l.keys = off, off += a256(size_t(n) * sizeof(uint32_t));
    l.total = off;
    return l;
}

size_t cull_workspace_bytes(int64_t records) { return cull_layout(records).total; }

int cull_build(cudaStream_t s, int64_t records, const Tri48 *pack_in, unsigned char *ws, const CullLayout &l,
               void *sort_ws, size_t sort_bytes) {
    if (records <= 0 || records % kCullGroup != 0) return DRT_ERR_BAD_EXTENT;
    float *bounds = reinterpret_cast<float *>(ws + l.bounds);
    uint32_t *keys = reinterpret_cast<uint32_t *>(ws + l.keys);
    Tri48 *pack = reinterpret_cast<Tri48 *>(ws + l.pack);
    CullNode *groups = reinterpret_cast<CullNode *>(ws + l.groups);
    cull_bounds_kernel<<<1, 1024, 0, s>>>(records, pack_in, bounds);
    cull_keys_kernel<<<unsigned((records + 255) / 256), 256, 0, s>>>(records, pack_in, bounds, keys);
    if (cudaGetLastError() != cudaSuccess) return DRT_ERR_CUDA;
    const int rc = drt_sort_records_by_keys(s, records, pack_in, keys, sort_ws, sort_bytes, pack);
    if (rc != DRT_OK) return rc;
    cull_group_nodes_kernel<<<unsigned((l.num_groups + 127) / 128), 128, 0, s>>>(l.num_groups, pack, groups);
    CullNode *walk = reinterpret_cast<CullNode *>(ws + l.walk);
    for (int lev = l.levels.num_levels - 2; lev >= 0; --lev)  // bottom-up: level lev from level lev + 1
        cull_tile_nodes_kernel<<<unsigned((l.levels.size[lev] + 63) / 64), 64, 0, s>>>(
            l.levels.size[lev], l.levels.size[lev + 1], walk + l.levels.offset[lev + 1], walk + l.levels.offset[lev],
            kWalkFan);
    return cudaGetLastError() == cudaSuccess ? DRT_OK : DRT_ERR_CUDA;
}

}  // namespace drt
