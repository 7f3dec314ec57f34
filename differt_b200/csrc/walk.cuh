// The culled traversal kernel (see cull.cuh for the proof behind node_culled): shared by the path-blockage
// stage of the trace (trace.cu) and the flat any-hit query (intersect.cu).
#pragma once

#include "cull.cuh"

namespace drt {

// ------------------------------------------------------------------------------------------------
// Blockage, culled traversal (cull.cuh): ONE WARP per candidate walks the 8-ary hierarchy over the
// Morton-ordered pack, one segment at a time — the LAST segment first: it ends at the receiver, at
// street level, and is the likeliest to be blocked — with all 32 lanes busy on that segment:
//   * node step: up to four pending nodes are popped from a per-warp stack in shared memory and
//     lane (j, c) tests child c of node j with node_culled (4 x 8 children, 80-byte nodes read as
//     contiguous 640-byte blocks); the surviving children are pushed back — groups on a second stack;
//   * group step: up to four pending groups are popped and lane (j, k) evaluates the exact
//     Möller–Trumbore test of triangle k of group j.  The first hit retires the candidate.
// Control flow is warp-uniform (ballots + popcounts, no divergence), the stacks hold at most a few
// dozen entries (depth-first order), and a candidate only ever touches the nodes and triangles around
// its own segments.
//
// Measured alternatives (profiles/README.md): a thread per candidate — the classic BVH traversal — runs
// at 29 % SIMT efficiency with the L1 data pipe saturated by 32 lanes fetching 32 different nodes
// (131 ms per bench step); all segments of a candidate walking together, lane = (segment, child),
// visits every node ANY segment touches before the blocking one is found (153 ms).
// ------------------------------------------------------------------------------------------------

#ifndef DRT_WALK_WARPS
#define DRT_WALK_WARPS 8
#endif
#ifndef DRT_WALK_CTAS
#define DRT_WALK_CTAS 4
#endif
#ifndef DRT_WALK_CHUNK
#define DRT_WALK_CHUNK 4
#endif
#ifndef DRT_WALK_HEAD_ROWS
#define DRT_WALK_HEAD_ROWS 1
#endif
#ifndef DRT_WALK_MIN_TILES
#define DRT_WALK_MIN_TILES 1
#endif
constexpr int kCullHead = DRT_WALK_MIN_TILES;  // meshes with more tiles than this take the culled traversal
constexpr int kWalkWarps = DRT_WALK_WARPS;
constexpr int kWalkHeadRows = DRT_WALK_HEAD_ROWS;  // rows of 32 largest triangles tested before the walk
constexpr int kWalkStack = 256;      // node stack: a step pops <= 4 and pushes <= 32, depth-first: <= 7 x 28 + 32
constexpr int kWalkGroupStack = 64;  // group stack: drained 4 at a time as soon as it holds 4: <= 3 + 32

// RAYS = false: unit = path candidate, `vertices` [units, NSEG + 1, 3]; a hit writes mask[path] = 0.
// RAYS = true (NSEG = 1): unit = ray, `vertices` = origins [units, 3], `directions` [units, 3] (used as
// given: o + d - o would not reproduce d); a hit writes mask[ray] = 1 (the caller zero-fills it).
template <int NSEG, bool RAYS>
__global__ void __launch_bounds__(kWalkWarps * 32, DRT_WALK_CTAS)
path_walk_kernel(const Tri48 *__restrict__ pack, const CullNode *__restrict__ nodes, const WalkLevels lv,
                 const Tri48 *__restrict__ head_row, const int64_t num_units_host,
                 const int64_t *__restrict__ num_units_dev, const float *__restrict__ vertices,
                 const float *__restrict__ directions, const uint32_t *__restrict__ list, const float eps,
                 const float thr, uint8_t *__restrict__ mask, unsigned long long *cursor, int64_t *tests_done) {
    static_assert(!RAYS || NSEG == 1, "flat rays are single segments");
    static_assert(kWalkFan == 8 && kCullGroup == 8, "lane = 8 x entry + child / triangle");
    // entry = (L << 28) | index: node `index` of level L - 1 (L = 0: the virtual root); its children are
    // the nodes 8 index .. 8 index + 7 of level L
    __shared__ uint32_t node_stack_all[kWalkWarps][kWalkStack];
    __shared__ uint32_t group_stack_all[kWalkWarps][kWalkGroupStack];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *node_stack = node_stack_all[warp], *group_stack = group_stack_all[warp];
    __shared__ int level_offset[kWalkMaxLevels], level_size[kWalkMaxLevels];
    if (threadIdx.x < kWalkMaxLevels) {
        level_offset[threadIdx.x] = lv.offset[threadIdx.x];
        level_size[threadIdx.x] = lv.size[threadIdx.x];
    }
    __syncthreads();
    const int64_t num_units = num_units_dev ? *num_units_dev : num_units_host;
    constexpr int NV = NSEG + 1;
    static_assert(3 * NV <= 32, "one float per lane prefetch");
    auto path_of = [&](int64_t u) -> int64_t { return list != nullptr ? int64_t(list[u]) : u; };
    auto fetch = [&](int64_t path) -> float {
        if (RAYS) return lane < 3 ? vertices[path * 3 + lane] : (lane < 6 ? directions[path * 3 + lane - 3] : 0.0f);
        return lane < 3 * NV ? vertices[path * (3 * NV) + lane] : 0.0f;
    };
    auto seg_dir = [&](const float3 next, const float3 o) -> float3 {
        return RAYS ? next : sub3(next, o);  // jnp.diff (_solvers.py:593) for paths
    };
    const int leaf = lv.num_levels - 1;
    int start_level = 0;
    while (start_level < leaf && lv.size[start_level + 1] <= 32) ++start_level;
    const int sub = lane & 7, slot = lane >> 3;  // child / triangle, and which of the (up to) 4 popped entries
    // head test: every segment of a candidate against the kHeadPer = 32 / NSEG largest triangles (front of the
    // area-sorted pack) in ONE Möller–Trumbore evaluation, lane = (segment, triangle) — a random segment is blocked
    // by a triangle with a probability proportional to its area: the two ground triangles alone block 32 % of the
    // bench batch, the 8 largest 32.7 %, the 32 largest 35.7 % (a full row of 32 per segment cost four evaluations
    // for those last 3 %)
    constexpr int kHeadPer = 32 / NSEG;
    const int head_seg = lane / kHeadPer;  // >= NSEG: idle lane
    const bool head_lane = kWalkHeadRows > 0 && head_seg < NSEG;
    const int head_src = 3 * (head_lane ? head_seg : 0);
    Tri head_tri;
    {
        const Tri48 &hr = head_row[lane % kHeadPer];
        head_tri = unpack(hr.a, hr.b, hr.c);
    }

    // Work distribution: chunks of kChunk consecutive candidates from a global cursor (the cost of a
    // candidate varies by two orders of magnitude; a static stride would leave most warps idle at the
    // end).  Small chunks: a warp that draws a run of unblocked candidates is busy for ~30 us each, and
    // nothing may wait for it at the tail of a small batch (guided self-scheduling — reading the cursor to
    // size the chunk — was measured slower: 28.5 vs 27.8 ms on the bench batch).
    constexpr int kChunk = DRT_WALK_CHUNK;
    int64_t tests = 0;
    while (true) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(cursor, (unsigned long long)kChunk);
        const int64_t unit0 = int64_t(__shfl_sync(kFull, base, 0));
        if (unit0 >= num_units) break;
        const int64_t unit1 = unit0 + kChunk < num_units ? unit0 + kChunk : num_units;
        float pf = fetch(path_of(unit0));
    for (int64_t unit = unit0; unit < unit1; ++unit) {
        const float pv = pf;  // this candidate's vertices, one float per lane
        const int64_t path = path_of(unit);
        if (unit + 1 < unit1) pf = fetch(path_of(unit + 1));

        bool blocked = false;
        if (kWalkHeadRows > 0) {
            const float3 ho = make_float3(__shfl_sync(kFull, pv, head_src), __shfl_sync(kFull, pv, head_src + 1),
                                          __shfl_sync(kFull, pv, head_src + 2));
            const float3 hn = make_float3(__shfl_sync(kFull, pv, head_src + 3), __shfl_sync(kFull, pv, head_src + 4),
                                          __shfl_sync(kFull, pv, head_src + 5));
            const float3 hd = seg_dir(hn, ho);
            bool weird = false;
            bool hit = head_lane && mt_any_fast(ho, hd, head_tri, eps, thr, weird);
            if (head_lane && weird) {  // |a| outside the fast reciprocal's range: the general test decides
                float t;
                hit = mt_exact(ho, hd, head_tri, eps, t) && t < thr;
            }
            tests += NSEG * kHeadPer;
            blocked = __any_sync(kFull, hit);
        }
        for (int sgm = NSEG - 1; sgm >= 0 && !blocked; --sgm) {
            const float3 o = make_float3(__shfl_sync(kFull, pv, 3 * sgm), __shfl_sync(kFull, pv, 3 * sgm + 1),
                                         __shfl_sync(kFull, pv, 3 * sgm + 2));
            const float3 next = make_float3(__shfl_sync(kFull, pv, 3 * sgm + 3), __shfl_sync(kFull, pv, 3 * sgm + 4),
                                            __shfl_sync(kFull, pv, 3 * sgm + 5));
            const float3 d = seg_dir(next, o);
            // d = 0 → a = 0 → no hit; a non-finite origin or direction → NaN/inf comparisons → no hit
            if ((d.x == 0.0f && d.y == 0.0f && d.z == 0.0f) || !finite3(o) || !finite3(d)) continue;
            const SegCull sc = make_seg_cull(o, d);

            // first step: lane l tests node l of the deepest level that has at most 32 nodes (one step
            // instead of walking the 3-node and 20-node levels of a 10 000-triangle mesh one after the other)
            int nn = 0, ng = 0;  // stack heights (warp-uniform)
            {
                const bool keep = lane < level_size[start_level] &&
                                  !node_culled(sc, nodes[level_offset[start_level] + lane]);
                const unsigned b0 = __ballot_sync(kFull, keep);
                const int pos = __popc(b0 & ((1u << lane) - 1u));
                __syncwarp();
                if (keep) {
                    if (start_level == leaf) group_stack[pos] = uint32_t(lane);
                    else node_stack[pos] = (uint32_t(start_level + 1) << 28) | uint32_t(lane);
                }
                if (start_level == leaf) ng = __popc(b0);
                else nn = __popc(b0);
                __syncwarp();
            }
            while (nn > 0 || ng > 0) {
                if (ng >= 4 || nn == 0) {
                    // ---- group step: up to 4 groups x 8 triangles
                    const int take = ng < 4 ? ng : 4;
                    ng -= take;
                    bool hit = false;
                    if (slot < take) {
                        const uint32_t g = group_stack[ng + slot];
                        const Tri48 *rec = pack + size_t(g) * kCullGroup + sub;
                        const float4 ta = rec->a, tb = rec->b, tc = rec->c;
                        const Tri tr = unpack(ta, tb, tc);
                        bool weird = false;
                        hit = mt_any_fast(o, d, tr, eps, thr, weird);
                        if (weird) {  // |a| outside the fast reciprocal's range: the general test decides
                            float t;
                            hit = mt_exact(o, d, tr, eps, t) && t < thr;
                        }
                    }
                    tests += take * kCullGroup;
                    if (__any_sync(kFull, hit)) {
                        blocked = true;
                        break;
                    }
                } else {
                    // ---- node step: up to 4 nodes x 8 children
                    const int take = nn < 4 ? nn : 4;
                    nn -= take;
                    bool keep = false;
                    uint32_t child = 0;
                    int L = 0;
                    if (slot < take) {
                        const uint32_t e = node_stack[nn + slot];
                        L = int(e >> 28);
                        child = (e & 0x0fffffffu) * kWalkFan + sub;
                        if (int(child) < level_size[L]) keep = !node_culled(sc, nodes[level_offset[L] + child]);
                    }
                    __syncwarp();  // every lane has read its entry before the stacks are overwritten
                    const bool to_group = keep && L == leaf;
                    const bool to_node = keep && L != leaf;
                    const unsigned bg = __ballot_sync(kFull, to_group), bn = __ballot_sync(kFull, to_node);
                    const unsigned below = (1u << lane) - 1u;
                    if (to_group) group_stack[ng + __popc(bg & below)] = child;
                    if (to_node) node_stack[nn + __popc(bn & below)] = (uint32_t(L + 1) << 28) | child;
                    ng += __popc(bg);
                    nn += __popc(bn);
                    __syncwarp();
                }
            }
        }
        if (blocked && lane == 0) mask[path] = RAYS ? 1 : 0;
        __syncwarp();
    }
    }
    if (tests_done != nullptr && lane == 0 && tests)
        atomicAdd(reinterpret_cast<unsigned long long *>(tests_done), (unsigned long long)tests);
}


// ------------------------------------------------------------------------------------------------
// First hit (nearest triangle) with the same culled traversal: one warp per ray.  A node is skipped only
// if node_culled proves that no triangle in it is reported as hit at a distance t <= the best found so
// far (ties included: the reference's tie rule picks among EQUAL distances by index, tie_key), so the
// winner and its distance are those of the all-pairs reduction bit for bit.  The bound shrinks as hits
// are found; children are not ordered front to back (any order is exact; the Morton order of the
// hierarchy is spatially coherent, which is most of the benefit).
// `orig` maps a position in the Morton-ordered pack to the triangle's index in the caller's mesh.
// ------------------------------------------------------------------------------------------------

// tie key of the reference's first-hit reduction (_utils.py:1865-1868, 1886): smaller wins.
__device__ __forceinline__ uint32_t first_hit_tie_key(int64_t j, int64_t bs, int64_t T) {
    if (bs <= 0 || bs >= T) return static_cast<uint32_t>(j);
    const int64_t nb = (T + bs - 1) / bs;  // batches incl. the remainder batch
    const int64_t b = j / bs;
    return static_cast<uint32_t>((nb - 1 - b) * bs + (j - b * bs));
}

static __global__ void __launch_bounds__(kWalkWarps * 32, DRT_WALK_CTAS)
ray_first_walk_kernel(const Tri48 *__restrict__ pack, const uint32_t *__restrict__ orig,
                      const CullNode *__restrict__ nodes, const WalkLevels lv, const int64_t num_rays,
                      const float *__restrict__ origins, const float *__restrict__ directions, const float eps,
                      const int64_t batch_size, const int64_t num_triangles, int32_t *__restrict__ out_index,
                      float *__restrict__ out_t, unsigned long long *cursor, int64_t *tests_done) {
    __shared__ uint32_t node_stack_all[kWalkWarps][kWalkStack];
    __shared__ uint32_t group_stack_all[kWalkWarps][kWalkGroupStack];
    __shared__ int level_offset[kWalkMaxLevels], level_size[kWalkMaxLevels];
    if (threadIdx.x < kWalkMaxLevels) {
        level_offset[threadIdx.x] = lv.offset[threadIdx.x];
        level_size[threadIdx.x] = lv.size[threadIdx.x];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *node_stack = node_stack_all[warp], *group_stack = group_stack_all[warp];
    const int leaf = lv.num_levels - 1;
    int start_level = 0;
    while (start_level < leaf && lv.size[start_level + 1] <= 32) ++start_level;
    const int sub = lane & 7, slot = lane >> 3;
    const bool fast_ok = eps >= 1.17549435e-38f;
    constexpr int kChunk = DRT_WALK_CHUNK;
    int64_t tests = 0;
    while (true) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(cursor, (unsigned long long)kChunk);
        const int64_t unit0 = int64_t(__shfl_sync(kFull, base, 0));
        if (unit0 >= num_rays) break;
        const int64_t unit1 = unit0 + kChunk < num_rays ? unit0 + kChunk : num_rays;
        for (int64_t ray = unit0; ray < unit1; ++ray) {
            const float pv = lane < 3 ? origins[ray * 3 + lane] : (lane < 6 ? directions[ray * 3 + lane - 3] : 0.0f);
            const float3 o = make_float3(__shfl_sync(kFull, pv, 0), __shfl_sync(kFull, pv, 1), __shfl_sync(kFull, pv, 2));
            const float3 d = make_float3(__shfl_sync(kFull, pv, 3), __shfl_sync(kFull, pv, 4), __shfl_sync(kFull, pv, 5));
            // warp-uniform running minimum
            float best_t = CUDART_INF_F;
            uint32_t best_key = 0xffffffffu;
            int32_t best_idx = -1;
            const bool dead = (d.x == 0.0f && d.y == 0.0f && d.z == 0.0f) || !finite3(o) || !finite3(d);
            if (!dead) {
                const SegCull sc = make_ray_cull(o, d);
                int nn = 0, ng = 0;
                {   // first step: the deepest level with at most 32 nodes, one node per lane
                    const bool keep = lane < level_size[start_level] &&
                                      !node_culled(sc, nodes[level_offset[start_level] + lane], best_t);
                    const unsigned b0 = __ballot_sync(kFull, keep);
                    const int pos = __popc(b0 & ((1u << lane) - 1u));
                    __syncwarp();
                    if (keep) {
                        if (start_level == leaf) group_stack[pos] = uint32_t(lane);
                        else node_stack[pos] = (uint32_t(start_level + 1) << 28) | uint32_t(lane);
                    }
                    if (start_level == leaf) ng = __popc(b0);
                    else nn = __popc(b0);
                    __syncwarp();
                }
                while (nn > 0 || ng > 0) {
                    if (ng >= 4 || nn == 0) {
                        const int take = ng < 4 ? ng : 4;
                        ng -= take;
                        float t = CUDART_INF_F;
                        uint32_t key = 0xffffffffu;
                        int32_t idx = -1;
                        if (slot < take) {
                            const uint32_t pos = group_stack[ng + slot] * kCullGroup + sub;
                            const float4 ta = pack[pos].a, tb = pack[pos].b, tc = pack[pos].c;
                            const Tri tr = unpack(ta, tb, tc);
                            bool weird = !fast_ok, hit = false;
                            float tt = 0.0f;
                            if (fast_ok) hit = mt_first_fast(o, d, tr, eps, tt, weird);
                            if (weird) hit = mt_exact(o, d, tr, eps, tt);
                            if (hit && tt <= best_t) {
                                t = tt + 0.0f;  // -0 and +0 are the same distance: canonicalise
                                idx = int32_t(orig[pos]);
                                key = first_hit_tie_key(idx, batch_size, num_triangles);
                            }
                        }
                        tests += take * kCullGroup;
                        // warp minimum of (t, key): ordered bits of t, then the tie key
                        const uint32_t tb_ = idx >= 0 ? float_order_bits(t) : 0xffffffffu;
                        const uint32_t tmin = __reduce_min_sync(kFull, tb_);
                        if (tmin != 0xffffffffu) {
                            const uint32_t k2 = (tb_ == tmin) ? key : 0xffffffffu;
                            const uint32_t kmin = __reduce_min_sync(kFull, k2);
                            const unsigned owner = __ballot_sync(kFull, tb_ == tmin && k2 == kmin);
                            const int src = __ffs(owner) - 1;
                            const float wt = __shfl_sync(kFull, t, src);
                            const int32_t widx = __shfl_sync(kFull, idx, src);
                            if (wt < best_t || (wt == best_t && kmin < best_key)) {
                                best_t = wt;
                                best_key = kmin;
                                best_idx = widx;
                            }
                        }
                    } else {
                        const int take = nn < 4 ? nn : 4;
                        nn -= take;
                        bool keep = false;
                        uint32_t child = 0;
                        int L = 0;
                        if (slot < take) {
                            const uint32_t e = node_stack[nn + slot];
                            L = int(e >> 28);
                            child = (e & 0x0fffffffu) * kWalkFan + sub;
                            if (int(child) < level_size[L])
                                keep = !node_culled(sc, nodes[level_offset[L] + child], best_t);
                        }
                        __syncwarp();
                        const bool to_group = keep && L == leaf;
                        const bool to_node = keep && L != leaf;
                        const unsigned bg = __ballot_sync(kFull, to_group), bn = __ballot_sync(kFull, to_node);
                        const unsigned below = (1u << lane) - 1u;
                        if (to_group) group_stack[ng + __popc(bg & below)] = child;
                        if (to_node) node_stack[nn + __popc(bn & below)] = (uint32_t(L + 1) << 28) | child;
                        ng += __popc(bg);
                        nn += __popc(bn);
                        __syncwarp();
                    }
                }
            }
            if (lane == 0) {  // a "hit" at a non-finite distance is a miss (_utils.py:1957-1959)
                const bool fin = isfinite(best_t);
                out_index[ray] = fin ? best_idx : -1;
                out_t[ray] = fin ? best_t : CUDART_INF_F;
            }
            __syncwarp();
        }
    }
    if (tests_done != nullptr && lane == 0 && tests)
        atomicAdd(reinterpret_cast<unsigned long long *>(tests_done), (unsigned long long)tests);
}

}  // namespace drt
