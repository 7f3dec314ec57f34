// K6 fused trace + validate, its reverse mode, TracedPaths.masked() compaction and the on-device
// path-candidate decode.
// Reference: differt/src/differt/geometry/_solvers.py:499-770 (`_trace_path_candidates`),
// differt/src/differt/geometry/_paths.py:299-328 (`masked`),
// differt-core/src/geometry/graph.rs:286-491 (complete-graph candidates).
//
// Stage A (one thread per (tx, rx, candidate)): gather mirrors from the packed mesh → image method →
//   inside-triangle / same-side / min-length / finite tests in registers → writes the dense
//   TracedPaths fields and a work list of the candidates that still need the blockage test.
// Stage B (all-pairs engine, one warp per candidate, its k+1 segments register-blocked): any-hit of
//   every segment against the whole mesh; a hit on any segment retires the candidate.
#include "walk.cuh"
#include "image_core.cuh"
#include "intersect_core.cuh"

namespace drt {

__device__ __forceinline__ float sgn(float x) { return float(x > 0.0f) - float(x < 0.0f); }

struct TraceArgs {
    const Tri48 *pack;        // geometry of every triangle (mask NOT applied)
    const uint8_t *tri_mask;  // nullable
    const float *tx, *rx;
    const int32_t *cand;
    int64_t T, ntx, nrx, C, P;
    float eps, min_len;
    float *out_vertices;
    int32_t *out_objects;
    uint8_t *out_mask;
    uint32_t *list;           // nullable (dense blockage): indices of candidates to test
    int64_t *list_count;
    // compact mode (drt_trace_valid_path_candidates): no dense outputs; candidates that pass the cheap
    // tests are appended to out_vertices / out_objects / out_mask / compact_index [capacity]
    int64_t capacity;
    int64_t *compact_index;
};

// Candidate-major decomposition: thread = (candidate c, transmitter, chunk of receivers).  The
// candidate's mirrors (first vertex + unit normal) and the triangles of its inside test are gathered
// from the packed mesh ONCE and kept in registers for every receiver of the chunk — the gather is the
// uncoalesced part of this stage (ncu: L1TEX-bound when done per path).  Lanes of a warp hold 32
// consecutive candidates of the same receiver, i.e. 32 consecutive paths: their dense outputs are
// staged in shared memory and leave as contiguous 16-byte stores.
template <int K, bool QUADS, bool COMPACT>
__global__ void __launch_bounds__(128)
trace_stage_a_kernel(const TraceArgs a, const int64_t rx_per_chunk) {
    constexpr int KK = K > 0 ? K : 1;
    constexpr int NT = QUADS ? 2 : 1;  // triangles per primitive in the inside test
    constexpr int NV3 = (K + 2) * 3;
    __shared__ __align__(16) float stage_v_all[4][32 * NV3];
    __shared__ __align__(16) int32_t stage_o_all[4][32 * (K + 2)];
    const int lane = threadIdx.x & 31;
    float *stage_v = stage_v_all[threadIdx.x >> 5];
    int32_t *stage_o = stage_o_all[threadIdx.x >> 5];

    const int64_t c = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    const int64_t c0 = c - lane;  // first candidate of the warp
    if (c0 >= a.C) return;        // warp-uniform
    const int64_t itx = blockIdx.z;
    const int64_t rx0 = int64_t(blockIdx.y) * rx_per_chunk;
    const int64_t rx1 = rx0 + rx_per_chunk < a.nrx ? rx0 + rx_per_chunk : a.nrx;
    const bool have = c < a.C;
    const unsigned have_mask = __ballot_sync(kFull, have);
    const int n_warp = int((a.C - c0) < 32 ? (a.C - c0) : 32);

    float3 mv[KK], mn[KK];
    Tri tri[KK][NT];
    int32_t ci[KK];
    bool active = true;
    if (have) {
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
    }
    const float3 from = ld3(a.tx + 3 * itx);

    for (int64_t irx = rx0; irx < rx1; ++irx) {
        const int64_t p0 = (itx * a.nrx + irx) * a.C + c0;  // first path of the warp
        const int64_t p = p0 + lane;
        bool prevalid = false;
        if (have) {
            float3 full[K + 2];
            full[0] = from;
            full[K + 1] = ld3(a.rx + 3 * irx);
            image_method_path<K>(full, mv, mn);

            bool inside = true, same = true, small = false, finite = true;
#pragma unroll
            for (int i = 0; i <= K; ++i) {
                const float3 o = full[i];
                const float3 d = sub3(full[i + 1], full[i]);
                small = small || (dot3(d, d) < a.min_len);
                if (i < K) {
                    float tt;
                    bool hit = mt_exact(o, d, tri[i][0], a.eps, tt);
                    if (QUADS) hit = hit || mt_exact(o, d, tri[i][NT - 1], a.eps, tt);
                    inside = inside && hit;
                    const float dp = dot3(sub3(full[i], mv[i]), mn[i]);
                    const float dn = dot3(sub3(full[i + 2], mv[i]), mn[i]);
                    same = same && (sgn(dp) == sgn(dn)) && (dp == dp) && (dn == dn);
                }
            }
#pragma unroll
            for (int i = 0; i < K + 2; ++i) finite = finite && finite3(full[i]);

            prevalid = inside && same && !small && finite && active;
            if (!COMPACT) {
                float *sv = stage_v + lane * NV3;
#pragma unroll
                for (int i = 0; i < K + 2; ++i)
                    st3(sv + 3 * i, finite ? full[i] : make_float3(0.f, 0.f, 0.f));
                int32_t *so = stage_o + lane * (K + 2);
                so[0] = int32_t(itx);
#pragma unroll
                for (int i = 0; i < K; ++i) so[i + 1] = ci[i];
                so[K + 1] = int32_t(irx);
                a.out_mask[p] = prevalid ? 1 : 0;
            } else {
                // the few candidates that pass go straight to their compact slot
                const unsigned m = __ballot_sync(have_mask, prevalid);
                if (prevalid) {
                    const int leader = __ffs(m) - 1;
                    int64_t start = 0;
                    if (lane == leader)
                        start = (int64_t)atomicAdd(reinterpret_cast<unsigned long long *>(a.list_count),
                                                   (unsigned long long)__popc(m));
                    start = __shfl_sync(m, start, leader);
                    const int64_t slot = start + __popc(m & ((1u << lane) - 1u));
                    if (slot < a.capacity) {
                        float *ov = a.out_vertices + slot * NV3;
#pragma unroll
                        for (int i = 0; i < K + 2; ++i) st3(ov + 3 * i, full[i]);
                        int32_t *oo = a.out_objects + slot * (K + 2);
                        oo[0] = int32_t(itx);
#pragma unroll
                        for (int i = 0; i < K; ++i) oo[i + 1] = ci[i];
                        oo[K + 1] = int32_t(irx);
                        a.out_mask[slot] = 1;
                        a.compact_index[slot] = p;
                    }
                }
            }
        }
        if (COMPACT) continue;
        __syncwarp();
        {
            float *gv = a.out_vertices + p0 * NV3;
            int32_t *go = a.out_objects + p0 * (K + 2);
            const bool wide = n_warp == 32 &&
                              ((reinterpret_cast<uintptr_t>(gv) | reinterpret_cast<uintptr_t>(go)) & 15) == 0;
            if (wide) {
                float4 *gv4 = reinterpret_cast<float4 *>(gv);
                const float4 *sv4 = reinterpret_cast<const float4 *>(stage_v);
#pragma unroll
                for (int i = lane; i < 8 * NV3; i += 32) gv4[i] = sv4[i];
                int4 *go4 = reinterpret_cast<int4 *>(go);
                const int4 *so4 = reinterpret_cast<const int4 *>(stage_o);
#pragma unroll
                for (int i = lane; i < 8 * (K + 2); i += 32) go4[i] = so4[i];
            } else {
                for (int i = lane; i < n_warp * NV3; i += 32) gv[i] = stage_v[i];
                for (int i = lane; i < n_warp * (K + 2); i += 32) go[i] = stage_o[i];
            }
        }
        __syncwarp();
        if (a.list != nullptr) {  // warp-aggregated append
            const unsigned m = __ballot_sync(kFull, prevalid);
            if (m) {
                int64_t start = 0;
                if (lane == 0)
                    start = (int64_t)atomicAdd(reinterpret_cast<unsigned long long *>(a.list_count),
                                               (unsigned long long)__popc(m));
                start = __shfl_sync(kFull, start, 0);
                if (prevalid) a.list[start + __popc(m & ((1u << lane) - 1u))] = uint32_t(p);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Blockage pass: a list of candidates against `num_head_tiles` consecutive tiles of the ordered pack,
// which stay resident in shared memory.  No ring, no CTA barrier after the initial load: warps run
// free, one candidate at a time, with the next candidate's vertices prefetched (one float per lane)
// while the current one is tested.  A hit retires the candidate (mask = 0); survivors are appended to
// `out_list` (32 at a time per warp) for the next pass of the cascade (tiles [8,16), [16,24), ...).
// ------------------------------------------------------------------------------------------------

#ifndef DRT_GREEDY_TILES
#define DRT_GREEDY_TILES 4
#endif
int drt_sort_records_by_keys(drt_stream_t stream, int64_t n, const void *pack_in, const uint32_t *keys,
                             void *workspace, size_t workspace_bytes, void *pack_out);  // pack_sort.cu
#ifndef DRT_PATH_HEAD_TILES
#define DRT_PATH_HEAD_TILES 8
#endif
#ifndef DRT_PATH_HEAD_WARPS
#define DRT_PATH_HEAD_WARPS 24
#endif
#ifndef DRT_PATH_HEAD_CTAS
#define DRT_PATH_HEAD_CTAS 1
#endif
constexpr int kPathHead = DRT_PATH_HEAD_TILES;
constexpr int kPathHeadWarps = DRT_PATH_HEAD_WARPS;
constexpr size_t kPathHeadSmem = size_t(kPathHead) * kTile * sizeof(Tri48) + 16;

template <int NSEG>
__global__ void __launch_bounds__(kPathHeadWarps * 32, DRT_PATH_HEAD_CTAS)
path_head_kernel(const Tri48 *__restrict__ pack, const int num_head_tiles, const int64_t num_units_host,
                 const int64_t *__restrict__ num_units_dev, const float *__restrict__ vertices,
                 const uint32_t *__restrict__ list, const float eps, const float thr,
                 uint8_t *__restrict__ mask, uint32_t *__restrict__ out_list, int64_t *out_count,
                 int64_t *tests_done) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Tri48 *head = reinterpret_cast<Tri48 *>(smem_raw);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + size_t(kPathHead) * kTile * sizeof(Tri48));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t num_units = num_units_dev ? *num_units_dev : num_units_host;
    const int64_t total_warps = int64_t(gridDim.x) * kPathHeadWarps;
    if (int64_t(blockIdx.x) * kPathHeadWarps >= num_units) return;
    constexpr uint32_t kTileBytes = kTile * sizeof(Tri48);
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(bar, uint32_t(num_head_tiles) * kTileBytes);
        for (int h = 0; h < num_head_tiles; ++h)
            bulk_g2s(head + size_t(h) * kTile, pack + size_t(h) * kTile, kTileBytes, bar);
    }

    constexpr int NV = NSEG + 1;  // vertices per path, 3 * NV floats <= 32 lanes
    static_assert(3 * NV <= 32, "one float per lane prefetch");
    auto path_of = [&](int64_t u) -> int64_t { return list != nullptr ? int64_t(list[u]) : u; };
    auto fetch = [&](int64_t path) -> float {
        return lane < 3 * NV ? vertices[path * (3 * NV) + lane] : 0.0f;
    };

    int64_t unit = int64_t(blockIdx.x) * kPathHeadWarps + warp;
    // software pipeline: vertices of `unit` in pf, path id of the unit after it in path_next
    int64_t path_cur = unit < num_units ? path_of(unit) : 0;
    float pf = unit < num_units ? fetch(path_cur) : 0.0f;
    int64_t path_next = unit + total_warps < num_units ? path_of(unit + total_warps) : 0;

    mbar_wait(bar, 0u);

    const bool fast_ok = eps >= 1.17549435e-38f;
    int64_t tests = 0;
    uint32_t keep = 0;  // lane i holds the i-th buffered survivor
    int nkeep = 0;
    for (; unit < num_units; unit += total_warps) {
        float3 o[NSEG], d[NSEG];
        {
            float3 prev = make_float3(__shfl_sync(kFull, pf, 0), __shfl_sync(kFull, pf, 1),
                                      __shfl_sync(kFull, pf, 2));
#pragma unroll
            for (int sgm = 0; sgm < NSEG; ++sgm) {
                const float3 next = make_float3(__shfl_sync(kFull, pf, 3 * sgm + 3),
                                                __shfl_sync(kFull, pf, 3 * sgm + 4),
                                                __shfl_sync(kFull, pf, 3 * sgm + 5));
                o[sgm] = prev;
                d[sgm] = sub3(next, prev);  // jnp.diff (_solvers.py:593)
                prev = next;
            }
        }
        const int64_t path = path_cur;
        // prefetch the next candidate
        path_cur = path_next;
        if (unit + total_warps < num_units) pf = fetch(path_cur);
        path_next = unit + 2 * total_warps < num_units ? path_of(unit + 2 * total_warps) : 0;

        // a segment with d == 0 gives h = d x e2 = 0 and a = 0 for every triangle: it can never hit.
        // Paths whose vertices were zeroed because they are not finite (_solvers.py:696-699) consist of
        // such segments only: they are unblocked by construction and need no test at all.
        // (Only for eps >= 0: with a negative epsilon the reference's a == 0 → t = 0 passes `t > eps`.)
        bool degenerate = eps >= 0.0f;
#pragma unroll
        for (int sgm = 0; sgm < NSEG; ++sgm)
            degenerate = degenerate && d[sgm].x == 0.0f && d[sgm].y == 0.0f && d[sgm].z == 0.0f;
        if (degenerate) continue;

        bool blocked = false;
        for (int h = 0; h < num_head_tiles && !blocked; ++h) {
            int rows;
            const uint32_t hits = scan_tile_any<NSEG, true>(head + size_t(h) * kTile, lane, o, d,
                                                            (1u << NSEG) - 1u, eps, thr, fast_ok, rows);
            tests += int64_t(NSEG) * 32 * rows;
            blocked = __any_sync(kFull, hits != 0);
        }
        if (blocked) {
            if (lane == 0) mask[path] = 0;
        } else {
            if (lane == nkeep) keep = uint32_t(path);
            if (++nkeep == 32) {
                int64_t base = 0;
                if (lane == 0)
                    base = (int64_t)atomicAdd(reinterpret_cast<unsigned long long *>(out_count), 32ull);
                base = __shfl_sync(kFull, base, 0);
                out_list[base + lane] = keep;
                nkeep = 0;
            }
        }
    }
    if (nkeep > 0) {
        int64_t base = 0;
        if (lane == 0)
            base = (int64_t)atomicAdd(reinterpret_cast<unsigned long long *>(out_count),
                                      (unsigned long long)nkeep);
        base = __shfl_sync(kFull, base, 0);
        if (lane < nkeep) out_list[base + lane] = keep;
    }
    if (tests_done != nullptr && lane == 0 && tests)
        atomicAdd(reinterpret_cast<unsigned long long *>(tests_done), (unsigned long long)tests);
}

// ------------------------------------------------------------------------------------------------
// Blockage, ordering pass: which triangles block THIS batch?  A sample of the candidates (every
// `stride`-th path) is tested against every triangle with no early exit, one thread per triangle, and
// the number of sampled candidates each triangle blocks is counted.  The pack is then re-sorted by
// that count (stable: ties keep the area order), so the head pass meets the triangles that block
// most of this transmitter / receiver configuration first.  Any-hit results do not depend on the
// order, so this is purely a work-reduction heuristic — which is also why the sample can use the
// fast test without its exactness fallback.
// ------------------------------------------------------------------------------------------------

__global__ void set_i64_kernel(int64_t *p, int64_t v) { *p = v; }

__global__ void sample_list_kernel(int64_t n, int64_t stride, uint32_t *__restrict__ list) {
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i < n) list[i] = uint32_t(i * stride);
}

constexpr int kCountBatch = 64;  // sampled candidates staged in shared memory at a time

template <int NSEG>
__global__ void __launch_bounds__(256)
hit_count_kernel(const Tri48 *__restrict__ pack, const int64_t num_records,
                 const float *__restrict__ vertices, const int64_t stride,
                 const uint32_t *__restrict__ sample_list /* nullable: sample s = path s * stride */,
                 const int64_t num_samples_host, const int64_t *__restrict__ num_samples_dev,
                 const int64_t samples_per_chunk, const float eps, const float thr,
                 uint32_t *__restrict__ counts) {
    constexpr int NV3 = (NSEG + 1) * 3;
    __shared__ float sv[kCountBatch][NV3];
    const int64_t j = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    Tri tr;
    {
        const int64_t jj = j < num_records ? j : num_records - 1;
        tr = unpack(pack[jj].a, pack[jj].b, pack[jj].c);
    }
    const int64_t num_samples = num_samples_dev != nullptr ? *num_samples_dev : num_samples_host;
    const int64_t s0 = int64_t(blockIdx.y) * samples_per_chunk;
    const int64_t s1 = s0 + samples_per_chunk < num_samples ? s0 + samples_per_chunk : num_samples;
    uint32_t c = 0;
    for (int64_t base = s0; base < s1; base += kCountBatch) {
        const int nb = int(s1 - base < kCountBatch ? s1 - base : kCountBatch);
        __syncthreads();
        for (int i = threadIdx.x; i < nb * NV3; i += blockDim.x) {
            const int u = i / NV3, k = i - u * NV3;
            const int64_t path = sample_list != nullptr ? int64_t(sample_list[base + u]) : (base + u) * stride;
            sv[u][k] = vertices[path * NV3 + k];
        }
        __syncthreads();
        for (int u = 0; u < nb; ++u) {
            bool any = false, weird = false;
            float3 prev = make_float3(sv[u][0], sv[u][1], sv[u][2]);
#pragma unroll
            for (int sgm = 0; sgm < NSEG; ++sgm) {
                const float3 next = make_float3(sv[u][3 * sgm + 3], sv[u][3 * sgm + 4], sv[u][3 * sgm + 5]);
                any = mt_any_fast(prev, sub3(next, prev), tr, eps, thr, weird) || any;
                prev = next;
            }
            c += any ? 1u : 0u;
        }
    }
    if (c != 0 && j < num_records) atomicAdd(&counts[j], c);
}

// generic order: flat (slot, segment) rays, RPW per warp, no path-level early exit
template <int RPW>
struct SegRays {
    const float *vertices;
    const uint32_t *list;
    const int64_t *count_dev;  // number of slots (device) or null
    int64_t count_host;
    int nseg;
    __device__ __forceinline__ uint32_t load(int64_t unit, float3 (&o)[RPW], float3 (&d)[RPW]) const {
        const int64_t n = (count_dev != nullptr ? *count_dev : count_host) * nseg;
        uint32_t valid = 0;
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int64_t ray = unit * RPW + r;
            o[r] = d[r] = make_float3(0.f, 0.f, 0.f);
            if (ray < n) {
                const int64_t slot = ray / nseg;
                const int s = int(ray % nseg);
                const int64_t p = list != nullptr ? int64_t(list[slot]) : slot;
                const float *v = vertices + (p * (nseg + 1) + s) * 3;
                o[r] = ld3(v);
                d[r] = sub3(ld3(v + 3), o[r]);
                valid |= 1u << r;
            }
        }
        return valid;
    }
};

template <int RPW>
struct SegSink {
    uint8_t *mask;
    const uint32_t *list;
    int nseg;
    __device__ __forceinline__ void any(int64_t unit, uint32_t hit, uint32_t valid) const {
#pragma unroll
        for (int r = 0; r < RPW; ++r)
            if ((hit & valid) & (1u << r)) {
                const int64_t slot = (unit * RPW + r) / nseg;
                mask[list != nullptr ? int64_t(list[slot]) : slot] = 0;
            }
    }
    __device__ __forceinline__ void first(int64_t, int, int32_t, float) const {}
};

__global__ void clamp_count_kernel(const int64_t *count, int64_t capacity, int64_t *clamped) {
    *clamped = *count < capacity ? *count : capacity;
}

__global__ void seg_units_kernel(const int64_t *count, int nseg, int rpw, int64_t *units) {
    *units = (*count * nseg + rpw - 1) / rpw;
}

// ------------------------------------------------------------------------------------------------
// reverse mode
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ void atomic_add3(float *p, float3 v) {
    if (v.x != 0.f) atomicAdd(p, v.x);
    if (v.y != 0.f) atomicAdd(p + 1, v.y);
    if (v.z != 0.f) atomicAdd(p + 2, v.z);
}

// sum over the lanes that share `key` with lane 0's... simple form: if the whole warp shares the key
// do one shuffle reduction and a single atomic, else fall back to per-lane atomics.
__device__ __forceinline__ void warp_accumulate3(float *base, int64_t key, float3 v, bool live) {
    const int64_t k0 = __shfl_sync(kFull, key, 0);
    const bool uniform = __all_sync(kFull, !live || key == k0);
    if (uniform) {
        if (!live) v = make_float3(0.f, 0.f, 0.f);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            v.x += __shfl_xor_sync(kFull, v.x, off);
            v.y += __shfl_xor_sync(kFull, v.y, off);
            v.z += __shfl_xor_sync(kFull, v.z, off);
        }
        const unsigned any_live = __ballot_sync(kFull, live);
        if (any_live && (threadIdx.x & 31) == (__ffs(any_live) - 1)) {
            const int64_t kk = key;
            atomic_add3(base + 3 * kk, v);
        }
    } else if (live) {
        atomic_add3(base + 3 * key, v);
    }
}

// Candidate-major decomposition: thread = (candidate c, transmitter, chunk of receivers).  The mirror
// data of candidate c (gather of the three vertices of each triangle, unit normals) is built once and
// reused for every receiver of the chunk; the cotangents that flow into the mirrors are accumulated
// in registers over the chunk, so the scatter into g_vertices costs 9k atomics per THREAD instead of
// per path, and the chain through `normalize((v1-v0) x (v2-v1))` — linear in the accumulated
// cotangent — is applied once.  Lanes of a warp hold consecutive candidates of the same receiver, so
// the cotangent reads are coalesced and g_rx needs one shuffle reduction + one atomic per warp.
template <int K>
__global__ void __launch_bounds__(128, (K <= 3 ? 4 : (K <= 5 ? 3 : 2)))
trace_vjp_kernel(int64_t V, int64_t T, const float *__restrict__ verts,
                 const int32_t *__restrict__ tris, int64_t ntx, const float *__restrict__ tx,
                 int64_t nrx, const float *__restrict__ rx, int64_t C,
                 const int32_t *__restrict__ cand, const float *__restrict__ g_out,
                 int64_t rx_per_chunk, float *g_tx, float *g_rx, float *g_verts) {
    constexpr int KK = K > 0 ? K : 1;
    constexpr int NV3 = (K + 2) * 3;  // floats of cotangent per path
    // The cotangents of a warp's 32 consecutive paths are one contiguous block of 32 * NV3 floats.  It is
    // copied into shared memory with 16-byte cp.async one receiver AHEAD (double buffer, no registers
    // held), and each lane then reads its own NV3 floats from there (stride NV3 = 3 (K + 2) words:
    // conflict-free for odd orders, 4-way at worst for even ones) — instead of NV3 scalar loads per lane
    // at a 4 * NV3-byte stride, issued only once the previous receiver is done.
    constexpr int NF4 = 8 * NV3;  // float4 per block
    __shared__ __align__(16) float stage_all[4][2][32 * NV3];
    const int lane = threadIdx.x & 31;
    float(*stage)[32 * NV3] = stage_all[threadIdx.x >> 5];
    const int64_t c = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    const int64_t itx = blockIdx.z;
    const int64_t rx0 = int64_t(blockIdx.y) * rx_per_chunk;
    const int64_t rx1 = rx0 + rx_per_chunk < nrx ? rx0 + rx_per_chunk : nrx;
    const bool have = c < C;

    // persistent per-thread state is kept small (the kernel is latency bound: registers buy warps):
    // first vertex + unit normal of every mirror and the accumulated cotangents; the triangle's vertex
    // numbers are re-read and the other two vertices re-loaded in the epilogue
    float3 v0[KK], mn[KK];
    if (have) {
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const int64_t t = min(max(int64_t(cand[c * K + i]), int64_t(0)), T - 1);
            const int64_t i0 = min(max(int64_t(tris[3 * t]), int64_t(0)), V - 1);
            const int64_t i1 = min(max(int64_t(tris[3 * t + 1]), int64_t(0)), V - 1);
            const int64_t i2 = min(max(int64_t(tris[3 * t + 2]), int64_t(0)), V - 1);
            v0[i] = ld3(verts + 3 * i0);
            mn[i] = unit_normal(v0[i], ld3(verts + 3 * i1), ld3(verts + 3 * i2));
        }
    }
    const float3 from = ld3(tx + 3 * itx);
    float3 acc_from = make_float3(0.f, 0.f, 0.f);
    float3 acc_mv[KK], acc_mn[KK];
#pragma unroll
    for (int i = 0; i < KK; ++i) acc_mv[i] = acc_mn[i] = make_float3(0.f, 0.f, 0.f);

    const int64_t c0 = c - lane;      // first candidate of the warp
    const bool whole = c0 + 32 <= C;  // a full warp of candidates
    auto fetch_ahead = [&](int64_t irx, float *dst) {
        const float *src = g_out + ((itx * nrx + irx) * C + c0) * NV3;
        if (whole && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
            for (int i = lane; i < NF4; i += 32)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + 4 * i)), "l"(src + 4 * i)
                             : "memory");
        } else if (c0 < C) {  // ragged or unaligned block: plain loads, same layout
            const int n = int(C - c0 < 32 ? C - c0 : 32) * NV3;
            for (int i = lane; i < n; i += 32) dst[i] = src[i];
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (rx0 < rx1) fetch_ahead(rx0, stage[0]);
    int buf = 0;
    for (int64_t irx = rx0; irx < rx1; ++irx, buf ^= 1) {
        float3 g_to = make_float3(0.f, 0.f, 0.f);
        if (irx + 1 < rx1) {
            fetch_ahead(irx + 1, stage[buf ^ 1]);  // in flight while this receiver is processed
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp();
        if (have) {
            const float *g = stage[buf] + lane * NV3;
            float3 gp[KK];
            const float3 g0 = ld3(g), g1 = ld3(g + 3 * (K + 1));
            bool nz = (g0.x != 0.f) || (g0.y != 0.f) || (g0.z != 0.f) || (g1.x != 0.f) ||
                      (g1.y != 0.f) || (g1.z != 0.f);
#pragma unroll
            for (int i = 0; i < K; ++i) {
                gp[i] = ld3(g + 3 * (i + 1));
                nz = nz || (gp[i].x != 0.f) || (gp[i].y != 0.f) || (gp[i].z != 0.f);
            }
            if (nz) {
                float3 full[K + 2];
                full[0] = from;
                full[K + 1] = ld3(rx + 3 * irx);
                image_method_path<K>(full, v0, mn);
                bool finite = true;
#pragma unroll
                for (int i = 0; i < K + 2; ++i) finite = finite && finite3(full[i]);
                if (finite) {  // where(is_finite, full_paths, 0): no gradient otherwise
                    acc_from = add3(acc_from, g0);
                    g_to = g1;
                    if (K > 0)
                        image_method_reverse<K>(full[0], full[K + 1], v0, mn, gp, acc_from, g_to, acc_mv,
                                                acc_mn);
                }
            }
        }
        __syncwarp();  // every lane is done with stage[buf] before the next iteration refills it
        // every lane of the warp works on the same receiver
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            g_to.x += __shfl_xor_sync(kFull, g_to.x, off);
            g_to.y += __shfl_xor_sync(kFull, g_to.y, off);
            g_to.z += __shfl_xor_sync(kFull, g_to.z, off);
        }
        if (lane == 0) atomic_add3(g_rx + 3 * irx, g_to);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        acc_from.x += __shfl_xor_sync(kFull, acc_from.x, off);
        acc_from.y += __shfl_xor_sync(kFull, acc_from.y, off);
        acc_from.z += __shfl_xor_sync(kFull, acc_from.z, off);
    }
    if (lane == 0) atomic_add3(g_tx + 3 * itx, acc_from);
    if (have) {
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const int64_t t = min(max(int64_t(cand[c * K + i]), int64_t(0)), T - 1);
            const int64_t i0 = min(max(int64_t(tris[3 * t]), int64_t(0)), V - 1);
            const int64_t i1 = min(max(int64_t(tris[3 * t + 1]), int64_t(0)), V - 1);
            const int64_t i2 = min(max(int64_t(tris[3 * t + 2]), int64_t(0)), V - 1);
            const float3 v1 = ld3(verts + 3 * i1), v2 = ld3(verts + 3 * i2);
            // n = N / len, N = A × B, A = v1 - v0, B = v2 - v1
            const float3 A = sub3(v1, v0[i]), B = sub3(v2, v1);
            const float3 N = cross3(A, B);
            const float len = __fsqrt_rn(dot3(N, N));
            float3 gN;
            if (len == 0.0f) {
                gN = acc_mn[i];
            } else {
                const float proj = dot3(mn[i], acc_mn[i]);
                gN = scale3(sub3(acc_mn[i], scale3(mn[i], proj)), __fdiv_rn(1.0f, len));
            }
            const float3 gA = cross3(B, gN), gB = cross3(gN, A);
            atomic_add3(g_verts + 3 * i0, sub3(acc_mv[i], gA));
            atomic_add3(g_verts + 3 * i1, sub3(gA, gB));
            atomic_add3(g_verts + 3 * i2, gB);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// compaction of valid paths (stable, row-major order)
// ------------------------------------------------------------------------------------------------

constexpr int kCompactThreads = 256;
constexpr int kCompactItems = 16;
constexpr int kCompactChunk = kCompactThreads * kCompactItems;  // 4096 paths per block

__device__ __forceinline__ int block_exclusive_scan(int v, int *total) {
    __shared__ int warp_sums[kCompactThreads / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int y = __shfl_up_sync(kFull, x, off);
        if (lane >= off) x += y;
    }
    if (lane == 31) warp_sums[w] = x;
    __syncthreads();
    if (w == 0) {
        int s = lane < kCompactThreads / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int y = __shfl_up_sync(kFull, s, off);
            if (lane >= off) s += y;
        }
        if (lane < kCompactThreads / 32) warp_sums[lane] = s;
    }
    __syncthreads();
    const int before = w > 0 ? warp_sums[w - 1] : 0;
    *total = warp_sums[kCompactThreads / 32 - 1];
    __syncthreads();
    return before + x - v;
}

// bit i = mask[start + i] != 0 for the thread's kCompactItems (= 16) consecutive paths: one 16-byte load
// when the span is whole and aligned (the mask base comes from an allocator: 256-byte aligned; start is a
// multiple of 16), byte loads at the ragged end
static_assert(kCompactItems == 16, "one uint4 of mask bytes per thread");
__device__ __forceinline__ unsigned load_mask_bits(const uint8_t *__restrict__ mask, const int64_t start,
                                                   const int64_t P) {
    unsigned bits = 0;
    if (start + kCompactItems <= P && (reinterpret_cast<uintptr_t>(mask + start) & 15) == 0) {
        const uint4 w = *reinterpret_cast<const uint4 *>(mask + start);
        const unsigned words[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if ((words[q] >> (8 * b)) & 0xffu) bits |= 1u << (4 * q + b);
    } else {
#pragma unroll
        for (int i = 0; i < kCompactItems; ++i)
            if (start + i < P && mask[start + i] != 0) bits |= 1u << i;
    }
    return bits;
}

__global__ void __launch_bounds__(kCompactThreads)
compact_count_kernel(int64_t P, const uint8_t *__restrict__ mask, int32_t *__restrict__ block_counts) {
    const int64_t start = int64_t(blockIdx.x) * kCompactChunk + threadIdx.x * kCompactItems;
    const int n = __popc(load_mask_bits(mask, start, P));
    int total;
    block_exclusive_scan(n, &total);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024)
compact_scan_kernel(int64_t num_blocks, const int32_t *__restrict__ block_counts,
                    int64_t *__restrict__ block_offsets, int64_t *out_count) {
    __shared__ int64_t warp_sums[32];
    __shared__ int64_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int64_t base = 0; base < num_blocks; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const int64_t v = i < num_blocks ? block_counts[i] : 0;
        int64_t x = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int64_t y = __shfl_up_sync(kFull, x, off);
            if (lane >= off) x += y;
        }
        if (lane == 31) warp_sums[w] = x;
        __syncthreads();
        if (w == 0) {
            int64_t s = warp_sums[lane];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int64_t y = __shfl_up_sync(kFull, s, off);
                if (lane >= off) s += y;
            }
            warp_sums[lane] = s;
        }
        __syncthreads();
        const int64_t before = (w > 0 ? warp_sums[w - 1] : 0) + carry_s;
        if (i < num_blocks) block_offsets[i] = before + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = before + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) *out_count = carry_s;
}

__global__ void __launch_bounds__(kCompactThreads)
compact_scatter_kernel(int64_t P, int nvert, const float *__restrict__ vertices,
                       const int32_t *__restrict__ objects, const uint8_t *__restrict__ mask,
                       const int64_t *__restrict__ block_offsets, int64_t capacity,
                       int64_t *__restrict__ out_index, float *__restrict__ out_vertices,
                       int32_t *__restrict__ out_objects) {
    const int64_t start = int64_t(blockIdx.x) * kCompactChunk + threadIdx.x * kCompactItems;
    const unsigned bits = load_mask_bits(mask, start, P);
    const int n = __popc(bits);
    int total;
    int64_t pos = block_offsets[blockIdx.x] + block_exclusive_scan(n, &total);
    for (int i = 0; i < kCompactItems; ++i) {
        if (bits & (1u << i)) {
            if (pos < capacity) {
                const int64_t p = start + i;
                if (out_index != nullptr) out_index[pos] = p;
                if (out_vertices != nullptr)
                    for (int q = 0; q < nvert * 3; ++q)
                        out_vertices[pos * nvert * 3 + q] = vertices[p * nvert * 3 + q];
                if (out_objects != nullptr)
                    for (int q = 0; q < nvert; ++q)
                        out_objects[pos * nvert + q] = objects[p * nvert + q];
            }
            ++pos;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// N1: complete-graph candidates from the linear index
// ------------------------------------------------------------------------------------------------

__global__ void complete_graph_candidates_kernel(int64_t n, int order, int64_t start, int64_t count,
                                                 int mult, int32_t *__restrict__ out) {
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= count) return;
    const int64_t base = n > 1 ? n - 1 : 1;
    int64_t rem = start + i;
    int64_t digits[DRT_MAX_ORDER * 2];
    for (int j = order - 1; j >= 1; --j) {
        digits[j] = rem % base;
        rem /= base;
    }
    digits[0] = rem;
    int64_t prev = digits[0];
    out[i * order] = int32_t(prev * mult);
    for (int j = 1; j < order; ++j) {
        const int64_t node = digits[j] + (digits[j] >= prev ? 1 : 0);
        out[i * order + j] = int32_t(node * mult);
        prev = node;
    }
}

// Workspace of the trace entry points.  The first `prefix` bytes depend on the mesh only (geometry pack,
// masked pack, area-sorted pack, the culled hierarchy, the sort scratch): drt_trace_prepare fills them
// once and a call flagged DRT_TRACE_PREPARED finds them already in place.
struct TraceWorkspace {
    size_t pack_geom, pack_active, pack_sorted, cull, sort_ws, sort_bytes, prefix;
    size_t pack_sorted2, pack_sorted3, hit_counts, list, list2, list3, counters, total;
    CullLayout cull_layout;
};

inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

inline TraceWorkspace trace_workspace_layout(int64_t T, int64_t P) {
    TraceWorkspace w;
    const size_t pack = drt_mesh_pack_bytes(T);
    size_t off = 0;
    w.pack_geom = off;
    off += align256(pack);
    w.pack_active = off;
    off += align256(pack);
    w.pack_sorted = off;  // blockage pack: active triangles in descending area order (pack_sort.cu)
    off += align256(pack);
    w.cull = off;  // spatially ordered pack + its node levels (cull.cuh)
    w.cull_layout = cull_layout(int64_t(pack / sizeof(Tri48)));
    off += align256(w.cull_layout.total);
    w.sort_ws = off;
    w.sort_bytes = drt_mesh_pack_sort_workspace_bytes(T);
    off += align256(w.sort_bytes);
    w.prefix = off;
    w.pack_sorted2 = off;  // ordering pass: the pack re-sorted by the hit counts of a sample of this batch,
    off += align256(pack);
    w.pack_sorted3 = off;  // ... and its ping-pong partner during the greedy rounds
    off += align256(pack);
    w.hit_counts = off;
    off += align256(pack / sizeof(Tri48) * sizeof(uint32_t));
    w.counters = off;
    off += 256;
    w.list = off;  // candidates that reach the blockage test (stage A → head pass)
    off += align256(size_t(P > 0 ? P : 1) * sizeof(uint32_t));
    w.list2 = off;  // candidates that survive a blockage pass (ping-pong with list3)
    off += align256(size_t(P > 0 ? P : 1) * sizeof(uint32_t));
    w.list3 = off;
    off += align256(size_t(P > 0 ? P : 1) * sizeof(uint32_t));
    w.total = off;
    return w;
}

// Mesh-only part of the trace: packs, area order, culled hierarchy → the prefix of `ws`.
inline int trace_prepare_mesh(drt_stream_t stream, int64_t V, int64_t T, const float *vertices,
                              const int32_t *triangles, const uint8_t *triangle_mask, unsigned char *ws,
                              const TraceWorkspace &w) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    Tri48 *pack_geom = reinterpret_cast<Tri48 *>(ws + w.pack_geom);
    Tri48 *pack_active = reinterpret_cast<Tri48 *>(ws + w.pack_active);
    int rc = drt_mesh_pack(stream, V, T, vertices, triangles, nullptr, pack_geom);
    if (rc != DRT_OK) return rc;
    rc = drt_mesh_pack(stream, V, T, vertices, triangles, triangle_mask, pack_active);  // null mask: a second copy
    if (rc != DRT_OK) return rc;
    if (T > 0) {
        Tri48 *pack_sorted = reinterpret_cast<Tri48 *>(ws + w.pack_sorted);
        rc = drt_mesh_pack_sort_by_area(stream, T, pack_active, ws + w.sort_ws, w.sort_bytes, pack_sorted);
        if (rc != DRT_OK) return rc;
        const int64_t records = int64_t(drt_mesh_pack_bytes(T) / sizeof(Tri48));
        if (records > int64_t(kCullHead) * kTile) {
            rc = cull_build(s, records, pack_sorted, ws + w.cull, w.cull_layout, ws + w.sort_ws, w.sort_bytes);
            if (rc != DRT_OK) return rc;
        }
    }
    return DRT_OK;
}

// per-thread profile ring (DRT_TRACE_PROFILE): event pairs around the blockage kernel
struct ProfileRing {
    cudaEvent_t start[DRT_PROFILE_SLOTS], stop[DRT_PROFILE_SLOTS];
    int count = 0;
    bool created = false;
    bool ensure() {
        if (created) return true;
        for (int i = 0; i < DRT_PROFILE_SLOTS; ++i)
            if (cudaEventCreate(&start[i]) != cudaSuccess || cudaEventCreate(&stop[i]) != cudaSuccess)
                return false;
        created = true;
        return true;
    }
};
static thread_local ProfileRing g_profile;

template <int K>
int trace_launch(cudaStream_t s, const TraceArgs &a, bool quads, bool dense, bool profile,
                 const Tri48 *pack_active, float hit_tol, int64_t *tests_done,
                 int64_t *units_scratch, uint32_t *list2, uint32_t *list3, int64_t *list2_count,
                 uint32_t *hit_counts,
                 Tri48 *pack_sorted2, Tri48 *pack_sorted3, void *sort_ws, size_t sort_bytes,
                 const unsigned char *cull_ws, const CullLayout &cull) {
    // candidates along x, receiver chunks along y (enough of them to fill the GPU), transmitters along z
    const int threads = 128;
    const int64_t cblocks = (a.C + threads - 1) / threads;
    if (a.ntx > 65535) return DRT_ERR_UNSUPPORTED;
    int64_t chunks = (int64_t(device_sm_count()) * 32 + cblocks * a.ntx - 1) / (cblocks * a.ntx);
    chunks = chunks < 1 ? 1 : (chunks > a.nrx ? a.nrx : chunks);
    if (chunks > 65535) chunks = 65535;
    const int64_t rx_per_chunk = (a.nrx + chunks - 1) / chunks;
    const dim3 grid(unsigned(cblocks), unsigned((a.nrx + rx_per_chunk - 1) / rx_per_chunk), unsigned(a.ntx));
    const bool compact = a.compact_index != nullptr;
    if (compact) {
        if (quads)
            trace_stage_a_kernel<K, true, true><<<grid, threads, 0, s>>>(a, rx_per_chunk);
        else
            trace_stage_a_kernel<K, false, true><<<grid, threads, 0, s>>>(a, rx_per_chunk);
    } else {
        if (quads)
            trace_stage_a_kernel<K, true, false><<<grid, threads, 0, s>>>(a, rx_per_chunk);
        else
            trace_stage_a_kernel<K, false, false><<<grid, threads, 0, s>>>(a, rx_per_chunk);
    }
    if (cudaGetLastError() != cudaSuccess) return DRT_ERR_CUDA;
    if (a.T == 0) return DRT_OK;  // empty mesh: nothing can block (_mesh.py:3053-3057)

    CoreParams p{};
    p.pack = pack_active;
    p.num_tiles = int(drt_mesh_pack_bytes(a.T) / sizeof(Tri48) / kTile);
    p.eps = a.eps;
    p.thr = 1.0f - hit_tol;
    p.num_triangles = a.T;
    p.tests_done = tests_done;
    const uint32_t *list = dense ? nullptr : a.list;
    constexpr int NSEG = K + 1;
    cudaError_t e;
    int slot = -1;
    if (profile && g_profile.ensure() && g_profile.count < DRT_PROFILE_SLOTS) {
        slot = g_profile.count++;
        cudaEventRecord(g_profile.start[slot], s);
    }
    if constexpr (NSEG <= 6) {
        // Meshes with more tiles than a resident pass holds: per-thread traversal of the culled
        // hierarchy for every candidate (cull.cuh; its proof needs FLT_MIN <= eps and 0 < thr <= 1,
        // other parameters keep the row-per-warp passes below).
        if (cull_ws != nullptr && p.num_tiles > kCullHead && p.eps >= 1.17549435e-38f && p.thr > 0.0f &&
            p.thr <= 1.0f) {
            const int64_t bound = compact ? a.capacity : a.P;
            if (compact) clamp_count_kernel<<<1, 1, 0, s>>>(a.list_count, a.capacity, units_scratch);
            const int64_t wblocks = (bound + kWalkWarps - 1) / kWalkWarps;
            const int64_t wres = int64_t(device_sm_count()) * DRT_WALK_CTAS;
            path_walk_kernel<NSEG, false><<<unsigned(wblocks < wres ? wblocks : wres), kWalkWarps * 32, 0, s>>>(
                reinterpret_cast<const Tri48 *>(cull_ws + cull.pack),
                reinterpret_cast<const CullNode *>(cull_ws + cull.walk), cull.levels, pack_active, bound,
                compact ? units_scratch : (dense ? nullptr : a.list_count), a.out_vertices,
                nullptr, compact ? nullptr : list, a.eps, p.thr, a.out_mask,
                reinterpret_cast<unsigned long long *>(list2_count + 4), tests_done);
            e = cudaGetLastError();
            if (e == cudaSuccess && tests_done != nullptr) set_i64_kernel<<<1, 1, 0, s>>>(tests_done + 3, 256);
            if (slot >= 0) cudaEventRecord(g_profile.stop[slot], s);
            return e == cudaSuccess ? DRT_OK : DRT_ERR_CUDA;
        }
        // ordering pass (dense batches only: a pruned work list is too short to pay for it)
        constexpr int64_t kSamples = 32768;
        if (dense && !compact && a.P >= 8 * kSamples && p.eps >= 1.17549435e-38f) {
            int64_t stride = (a.P / kSamples) | 1;
            auto gcd = [](int64_t x, int64_t y) { while (y) { const int64_t r = x % y; x = y; y = r; } return x; };
            while (gcd(stride, a.C) != 1) stride += 2;  // walk across candidates AND receivers
            const int64_t num_samples = (a.P + stride - 1) / stride;
            const int64_t records = int64_t(p.num_tiles) * kTile;
            if (cudaMemsetAsync(hit_counts, 0, size_t(records) * sizeof(uint32_t), s) != cudaSuccess)
                return DRT_ERR_CUDA;
            const int chunks = 32;
            const int64_t spc = (num_samples + chunks - 1) / chunks;
            const dim3 cgrid(unsigned((records + 255) / 256), chunks);
            hit_count_kernel<NSEG><<<cgrid, 256, 0, s>>>(pack_active, records, a.out_vertices, stride, nullptr,
                                                         num_samples, nullptr, spc, a.eps, p.thr, hit_counts);
            if (cudaGetLastError() != cudaSuccess) return DRT_ERR_CUDA;
            int rc = drt_mesh_pack_sort_by_keys(s, a.T, pack_active, hit_counts, sort_ws, sort_bytes, pack_sorted2);
            if (rc != DRT_OK) return rc;
            Tri48 *cur = pack_sorted2;
            Tri48 *other = pack_sorted3;  // (the area-sorted pack stays intact: it may belong to a prepared prefix)
#if DRT_GREEDY_TILES > 0
            // Greedy refinement (set-cover heuristic): raw hit counts are redundant — the triangles
            // that block the most samples tend to block the SAME samples.  Tile by tile: keep the
            // samples the tiles chosen so far do not block (one resident pass of the cascade kernel on
            // the sample list), recount the remaining triangles on those samples only, re-sort the
            // remaining records.  Sample lists ping-pong between list2 and list3 (free at this point).
            {
                auto hk0 = path_head_kernel<NSEG>;
                static unsigned long long configured0 = 0;
                if (ensure_dynamic_smem(hk0, kPathHeadSmem, configured0) != cudaSuccess) return DRT_ERR_CUDA;
                const int sms0 = device_sm_count();
                uint32_t *sl_in = list2, *sl_out = list3;
                int64_t *cnt_in = list2_count, *cnt_out = list2_count + 1;
                sample_list_kernel<<<unsigned((num_samples + 255) / 256), 256, 0, s>>>(num_samples, stride, sl_in);
                set_i64_kernel<<<1, 1, 0, s>>>(cnt_in, num_samples);  // (a kernel, not a pageable H2D copy: stays
                                                                      //  capturable in a CUDA graph)
                const int rounds = p.num_tiles - 1 < DRT_GREEDY_TILES ? p.num_tiles - 1 : DRT_GREEDY_TILES;
                for (int r = 0; r < rounds; ++r) {
                    if (cudaMemsetAsync(cnt_out, 0, sizeof(int64_t), s) != cudaSuccess) return DRT_ERR_CUDA;
                    // samples that tile r does not block
                    const int64_t sb = (num_samples + kPathHeadWarps - 1) / kPathHeadWarps;
                    hk0<<<unsigned(sb < sms0 ? sb : sms0), kPathHeadWarps * 32, kPathHeadSmem, s>>>(
                        cur + size_t(r) * kTile, 1, num_samples, cnt_in, a.out_vertices, sl_in, a.eps, p.thr,
                        a.out_mask, sl_out, cnt_out, nullptr);
                    // recount the records after tile r on the surviving samples, re-sort them
                    const int64_t rest = records - int64_t(r + 1) * kTile;
                    if (cudaMemsetAsync(hit_counts, 0, size_t(rest) * sizeof(uint32_t), s) != cudaSuccess)
                        return DRT_ERR_CUDA;
                    const dim3 rgrid(unsigned((rest + 255) / 256), chunks);
                    hit_count_kernel<NSEG><<<rgrid, 256, 0, s>>>(cur + size_t(r + 1) * kTile, rest, a.out_vertices,
                                                                 stride, sl_out, num_samples, cnt_out, spc, a.eps,
                                                                 p.thr, hit_counts);
                    if (cudaGetLastError() != cudaSuccess) return DRT_ERR_CUDA;
                    rc = drt_sort_records_by_keys(s, rest, cur + size_t(r + 1) * kTile, hit_counts, sort_ws,
                                                  sort_bytes, other + size_t(r + 1) * kTile);
                    if (rc != DRT_OK) return rc;
                    if (cudaMemcpyAsync(other, cur, size_t(r + 1) * kTile * sizeof(Tri48), cudaMemcpyDeviceToDevice,
                                        s) != cudaSuccess)
                        return DRT_ERR_CUDA;
                    Tri48 *tp = cur; cur = other; other = tp;
                    uint32_t *tl = sl_in; sl_in = sl_out; sl_out = tl;
                    int64_t *tc = cnt_in; cnt_in = cnt_out; cnt_out = tc;
                }
                // the cascade below starts from clean survivor counters
                if (cudaMemsetAsync(list2_count, 0, 2 * sizeof(int64_t), s) != cudaSuccess) return DRT_ERR_CUDA;
            }
#endif
            pack_active = cur;
            p.pack = pack_active;
            // stats[3]: 1 + number of greedy rounds — proof for the tests that this branch ran
            if (tests_done != nullptr)
                set_i64_kernel<<<1, 1, 0, s>>>(tests_done + 3,
                                               1 + (p.num_tiles - 1 < DRT_GREEDY_TILES ? p.num_tiles - 1
                                                                                        : DRT_GREEDY_TILES));
        }
        // head pass: every candidate against the likeliest blockers, resident, barrier free
        const int NT = p.num_tiles;
        auto hk = path_head_kernel<NSEG>;
        static unsigned long long configured = 0;  // per-device bits (common.cuh)
        if (ensure_dynamic_smem(hk, kPathHeadSmem, configured) != cudaSuccess) return DRT_ERR_CUDA;
        const int sms = device_sm_count();
        // compact mode: the units are the compact slots 0 .. min(count, capacity) - 1
        const int64_t bound = compact ? a.capacity : a.P;
        if (compact) clamp_count_kernel<<<1, 1, 0, s>>>(a.list_count, a.capacity, units_scratch);
        const int64_t hblocks = (bound + kPathHeadWarps - 1) / kPathHeadWarps;
        const int64_t hres = int64_t(sms) * DRT_PATH_HEAD_CTAS;
        // cascade of resident passes: tiles [0,8) for everyone, [8,16) for the survivors, ... — every
        // pass barrier-free, survivor lists ping-pong between list2 and list3
        const uint32_t *in_list = compact ? nullptr : list;
        const int64_t *in_count = compact ? units_scratch : (dense ? nullptr : a.list_count);
        uint32_t *out_list = list2;
        int64_t *out_count = list2_count;  // counters[2]; counters[3] is list3's
        e = cudaSuccess;
        const int head_step = kPathHead;
        for (int t0 = 0; t0 < NT && e == cudaSuccess; t0 += head_step) {
            const int nh = NT - t0 < head_step ? NT - t0 : head_step;
            hk<<<unsigned(hblocks < hres ? hblocks : hres), kPathHeadWarps * 32, kPathHeadSmem, s>>>(
                pack_active + size_t(t0) * kTile, nh, bound, in_count, a.out_vertices, in_list, a.eps, p.thr,
                a.out_mask, out_list, out_count, tests_done);
            e = cudaGetLastError();
            if (t0 == 0 && e == cudaSuccess)  // stats[2]: survivors of the first pass
                e = cudaMemcpyAsync(list2_count + 3, list2_count, sizeof(int64_t), cudaMemcpyDeviceToDevice, s);
            in_list = out_list;
            in_count = out_count;
            const bool to3 = out_list == list2;
            out_list = to3 ? list3 : list2;
            out_count = to3 ? list2_count + 1 : list2_count;
            if (e == cudaSuccess && t0 + head_step < NT) e = cudaMemsetAsync(out_count, 0, sizeof(int64_t), s);
        }
    } else {
        if (compact) return DRT_ERR_UNSUPPORTED;  // compact mode is built for orders <= 5
        constexpr int RPW = 3;
        p.num_units = (a.P * NSEG + RPW - 1) / RPW;
        p.num_units_dev = nullptr;
        if (!dense) {
            seg_units_kernel<<<1, 1, 0, s>>>(a.list_count, NSEG, RPW, units_scratch);
            p.num_units_dev = units_scratch;
        }
        SegRays<RPW> src{a.out_vertices, list, dense ? nullptr : a.list_count, a.P, NSEG};
        SegSink<RPW> sink{a.out_mask, list, NSEG};
        e = launch_intersect<RPW, MODE_ANY>(s, p, src, sink, p.num_units);
    }
    if (slot >= 0) cudaEventRecord(g_profile.stop[slot], s);
    return e == cudaSuccess ? DRT_OK : DRT_ERR_CUDA;
}

}  // namespace drt

using namespace drt;

extern "C" {

size_t drt_trace_workspace_bytes(int64_t T, int64_t ntx, int64_t nrx, int64_t C) {
    if (T < 0 || ntx < 0 || nrx < 0 || C < 0) return 0;
    return trace_workspace_layout(T, ntx * nrx * C).total;
}

int drt_trace_path_candidates(drt_stream_t stream, int64_t V, int64_t T, const float *vertices,
                              const int32_t *triangles, const uint8_t *triangle_mask,
                              int32_t assume_quads, int64_t ntx, const float *tx, int64_t nrx,
                              const float *rx, int64_t C, int32_t order, const int32_t *cand,
                              float epsilon, float hit_tol, float min_len, uint32_t flags,
                              void *workspace, size_t workspace_bytes, float *out_vertices,
                              int32_t *out_objects, uint8_t *out_mask, int64_t *stats) {
    if (V < 0 || T < 0 || ntx < 0 || nrx < 0 || C < 0 || order < 0) return DRT_ERR_BAD_EXTENT;
    if (order > DRT_MAX_ORDER) return DRT_ERR_UNSUPPORTED;
    const int64_t P = ntx * nrx * C;
    if (P >= (int64_t(1) << 32)) return DRT_ERR_BAD_EXTENT;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (stats != nullptr && cudaMemsetAsync(stats, 0, 4 * sizeof(int64_t), s) != cudaSuccess)
        return DRT_ERR_CUDA;
    if (P == 0) return DRT_OK;
    if (!tx || !rx || !out_vertices || !out_objects || !out_mask || !workspace)
        return DRT_ERR_NULL_POINTER;
    if (order > 0 && cand == nullptr) return DRT_ERR_NULL_POINTER;
    if (order > 0 && T < (assume_quads ? 2 : 1)) return DRT_ERR_BAD_EXTENT;  // candidates index triangles
    const TraceWorkspace w = trace_workspace_layout(T, P);
    if (workspace_bytes < w.total) return DRT_ERR_WORKSPACE;
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    Tri48 *pack_geom = reinterpret_cast<Tri48 *>(ws + w.pack_geom);
    int rc = DRT_OK;
    if ((flags & DRT_TRACE_PREPARED) == 0) {
        rc = trace_prepare_mesh(stream, V, T, vertices, triangles, triangle_mask, ws, w);
        if (rc != DRT_OK) return rc;
    }
    const Tri48 *pack_active = reinterpret_cast<const Tri48 *>(ws + (T > 0 ? w.pack_sorted : w.pack_active));
    const unsigned char *cull_ws = nullptr;
    if (order + 1 <= 6 && int64_t(drt_mesh_pack_bytes(T) / sizeof(Tri48)) > int64_t(kCullHead) * kTile)
        cull_ws = ws + w.cull;
    int64_t *counters = reinterpret_cast<int64_t *>(ws + w.counters);
    if (cudaMemsetAsync(counters, 0, 256, s) != cudaSuccess) return DRT_ERR_CUDA;
    const bool dense = (flags & DRT_TRACE_DENSE_BLOCKAGE) != 0;
    const bool profile = (flags & DRT_TRACE_PROFILE) != 0;

    TraceArgs a{};
    a.pack = pack_geom;
    a.tri_mask = triangle_mask;
    a.tx = tx;
    a.rx = rx;
    a.cand = cand;
    a.T = T;
    a.ntx = ntx;
    a.nrx = nrx;
    a.C = C;
    a.P = P;
    a.eps = epsilon;
    a.min_len = min_len;
    a.out_vertices = out_vertices;
    a.out_objects = out_objects;
    a.out_mask = out_mask;
    a.list = reinterpret_cast<uint32_t *>(ws + w.list);
    a.list_count = counters;
    int64_t *tests_done = stats;  // stats[0]
    int64_t *units_scratch = counters + 1;
    uint32_t *list2 = reinterpret_cast<uint32_t *>(ws + w.list2);
    uint32_t *list3 = reinterpret_cast<uint32_t *>(ws + w.list3);
    const bool quads = assume_quads != 0;
#define DRT_TRACE_CASE(K)                                                                         \
    case K:                                                                                       \
        rc = trace_launch<K>(s, a, quads, dense, profile, pack_active, hit_tol, tests_done,        \
                             units_scratch, list2, list3, counters + 2,                           \
                             reinterpret_cast<uint32_t *>(ws + w.hit_counts),                      \
                             reinterpret_cast<Tri48 *>(ws + w.pack_sorted2),                      \
                             reinterpret_cast<Tri48 *>(ws + w.pack_sorted3), ws + w.sort_ws,      \
                             w.sort_bytes, cull_ws, w.cull_layout);                               \
        break;
    switch (order) {
        DRT_TRACE_CASE(0) DRT_TRACE_CASE(1) DRT_TRACE_CASE(2) DRT_TRACE_CASE(3) DRT_TRACE_CASE(4)
        DRT_TRACE_CASE(5) DRT_TRACE_CASE(6) DRT_TRACE_CASE(7) DRT_TRACE_CASE(8)
    }
#undef DRT_TRACE_CASE
    if (rc != DRT_OK) return rc;
    if (stats != nullptr) {
        if (cudaMemcpyAsync(stats + 1, counters, sizeof(int64_t), cudaMemcpyDeviceToDevice, s) != cudaSuccess ||
            cudaMemcpyAsync(stats + 2, counters + 5, sizeof(int64_t), cudaMemcpyDeviceToDevice, s) != cudaSuccess)
            return DRT_ERR_CUDA;
    }
    return DRT_OK;
}

size_t drt_trace_prepared_bytes(int64_t T) {
    if (T < 0) return 0;
    return trace_workspace_layout(T, 0).prefix;
}

int drt_trace_prepare(drt_stream_t stream, int64_t V, int64_t T, const float *vertices, const int32_t *triangles,
                      const uint8_t *triangle_mask, void *prepared, size_t prepared_bytes) {
    if (V < 0 || T < 0) return DRT_ERR_BAD_EXTENT;
    if (!prepared) return DRT_ERR_NULL_POINTER;
    const TraceWorkspace w = trace_workspace_layout(T, 0);
    if (prepared_bytes < w.prefix) return DRT_ERR_WORKSPACE;
    return trace_prepare_mesh(stream, V, T, vertices, triangles, triangle_mask, static_cast<unsigned char *>(prepared), w);
}

size_t drt_trace_valid_workspace_bytes(int64_t T, int64_t capacity) {
    if (T < 0 || capacity < 0) return 0;
    return trace_workspace_layout(T, capacity).total;
}

int drt_trace_valid_path_candidates(drt_stream_t stream, int64_t V, int64_t T, const float *vertices,
                                    const int32_t *triangles, const uint8_t *triangle_mask,
                                    int32_t assume_quads, int64_t ntx, const float *tx, int64_t nrx,
                                    const float *rx, int64_t C, int32_t order, const int32_t *cand,
                                    float epsilon, float hit_tol, float min_len, uint32_t flags,
                                    int64_t capacity, void *workspace, size_t workspace_bytes, int64_t *out_count,
                                    int64_t *out_index, float *out_vertices, int32_t *out_objects,
                                    uint8_t *out_valid) {
    if (V < 0 || T < 0 || ntx < 0 || nrx < 0 || C < 0 || order < 0 || capacity <= 0) return DRT_ERR_BAD_EXTENT;
    if (order > 5) return DRT_ERR_UNSUPPORTED;
    if (capacity >= (int64_t(1) << 32)) return DRT_ERR_BAD_EXTENT;
    if (!out_count) return DRT_ERR_NULL_POINTER;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (cudaMemsetAsync(out_count, 0, sizeof(int64_t), s) != cudaSuccess) return DRT_ERR_CUDA;
    const int64_t P = ntx * nrx * C;
    if (P == 0) return DRT_OK;
    if (!tx || !rx || !out_vertices || !out_objects || !out_valid || !out_index || !workspace)
        return DRT_ERR_NULL_POINTER;
    if (order > 0 && cand == nullptr) return DRT_ERR_NULL_POINTER;
    if (order > 0 && T < (assume_quads ? 2 : 1)) return DRT_ERR_BAD_EXTENT;
    const TraceWorkspace w = trace_workspace_layout(T, capacity);
    if (workspace_bytes < w.total) return DRT_ERR_WORKSPACE;
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    Tri48 *pack_geom = reinterpret_cast<Tri48 *>(ws + w.pack_geom);
    int rc = DRT_OK;
    if ((flags & DRT_TRACE_PREPARED) == 0) {
        rc = trace_prepare_mesh(stream, V, T, vertices, triangles, triangle_mask, ws, w);
        if (rc != DRT_OK) return rc;
    }
    const Tri48 *pack_active = reinterpret_cast<const Tri48 *>(ws + (T > 0 ? w.pack_sorted : w.pack_active));
    const unsigned char *cull_ws = nullptr;
    if (int64_t(drt_mesh_pack_bytes(T) / sizeof(Tri48)) > int64_t(kCullHead) * kTile)
        cull_ws = ws + w.cull;
    int64_t *counters = reinterpret_cast<int64_t *>(ws + w.counters);
    if (cudaMemsetAsync(counters, 0, 256, s) != cudaSuccess) return DRT_ERR_CUDA;
    TraceArgs a{};
    a.pack = pack_geom;
    a.tri_mask = triangle_mask;
    a.tx = tx;
    a.rx = rx;
    a.cand = cand;
    a.T = T;
    a.ntx = ntx;
    a.nrx = nrx;
    a.C = C;
    a.P = P;
    a.eps = epsilon;
    a.min_len = min_len;
    a.out_vertices = out_vertices;
    a.out_objects = out_objects;
    a.out_mask = out_valid;
    a.list = nullptr;
    a.list_count = out_count;
    a.capacity = capacity;
    a.compact_index = out_index;
    const bool quads = assume_quads != 0;
#define DRT_TRACE_CASE(K)                                                                               \
    case K:                                                                                             \
        rc = trace_launch<K>(s, a, quads, false, false, pack_active, hit_tol, nullptr, counters + 1,    \
                             reinterpret_cast<uint32_t *>(ws + w.list2),                                 \
                             reinterpret_cast<uint32_t *>(ws + w.list3), counters + 2,                   \
                             reinterpret_cast<uint32_t *>(ws + w.hit_counts),                            \
                             reinterpret_cast<Tri48 *>(ws + w.pack_sorted2),                            \
                             reinterpret_cast<Tri48 *>(ws + w.pack_sorted3), ws + w.sort_ws,            \
                             w.sort_bytes, cull_ws, w.cull_layout);                                     \
        break;
    switch (order) {
        DRT_TRACE_CASE(0) DRT_TRACE_CASE(1) DRT_TRACE_CASE(2) DRT_TRACE_CASE(3) DRT_TRACE_CASE(4)
        DRT_TRACE_CASE(5)
    }
#undef DRT_TRACE_CASE
    return rc;
}

int drt_trace_path_candidates_vjp(drt_stream_t stream, int64_t V, int64_t T, const float *vertices,
                                  const int32_t *triangles, int64_t ntx, const float *tx,
                                  int64_t nrx, const float *rx, int64_t C, int32_t order,
                                  const int32_t *cand, const float *g_out, float *g_tx, float *g_rx,
                                  float *g_vertices) {
    if (V < 0 || T < 0 || ntx < 0 || nrx < 0 || C < 0 || order < 0) return DRT_ERR_BAD_EXTENT;
    if (order > DRT_MAX_ORDER) return DRT_ERR_UNSUPPORTED;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (ntx > 0 && g_tx == nullptr) return DRT_ERR_NULL_POINTER;
    if (nrx > 0 && g_rx == nullptr) return DRT_ERR_NULL_POINTER;
    if (V > 0 && g_vertices == nullptr) return DRT_ERR_NULL_POINTER;
    if (ntx > 0 && cudaMemsetAsync(g_tx, 0, size_t(ntx) * 12, s) != cudaSuccess) return DRT_ERR_CUDA;
    if (nrx > 0 && cudaMemsetAsync(g_rx, 0, size_t(nrx) * 12, s) != cudaSuccess) return DRT_ERR_CUDA;
    if (V > 0 && cudaMemsetAsync(g_vertices, 0, size_t(V) * 12, s) != cudaSuccess) return DRT_ERR_CUDA;
    const int64_t P = ntx * nrx * C;
    if (P == 0) return DRT_OK;
    if (!tx || !rx || !g_out) return DRT_ERR_NULL_POINTER;
    if (order > 0 && (!cand || !vertices || !triangles || T == 0 || V == 0)) return DRT_ERR_NULL_POINTER;
    // candidates along x, receiver chunks along y (enough of them to fill the GPU), transmitters along z
    const int threads = 128;
    const int64_t cblocks = (C + threads - 1) / threads;
    if (ntx > 65535) return DRT_ERR_UNSUPPORTED;
    int64_t chunks = (int64_t(device_sm_count()) * 16 + cblocks * ntx - 1) / (cblocks * ntx);
    chunks = chunks < 1 ? 1 : (chunks > nrx ? nrx : chunks);
    if (chunks > 65535) chunks = 65535;
    const int64_t rx_per_chunk = (nrx + chunks - 1) / chunks;
    const dim3 grid(unsigned(cblocks), unsigned((nrx + rx_per_chunk - 1) / rx_per_chunk), unsigned(ntx));
#define DRT_VJP_CASE(K)                                                                          \
    case K:                                                                                      \
        trace_vjp_kernel<K><<<grid, threads, 0, s>>>(V, T, vertices, triangles, ntx, tx, nrx, rx, \
                                                     C, cand, g_out, rx_per_chunk, g_tx, g_rx,    \
                                                     g_vertices);                                 \
        break;
    switch (order) {
        DRT_VJP_CASE(0) DRT_VJP_CASE(1) DRT_VJP_CASE(2) DRT_VJP_CASE(3) DRT_VJP_CASE(4)
        DRT_VJP_CASE(5) DRT_VJP_CASE(6) DRT_VJP_CASE(7) DRT_VJP_CASE(8)
    }
#undef DRT_VJP_CASE
    return cudaGetLastError() == cudaSuccess ? DRT_OK : DRT_ERR_CUDA;
}

int drt_profile_reset(void) {
    g_profile.count = 0;
    return g_profile.ensure() ? DRT_OK : DRT_ERR_CUDA;
}

int drt_profile_count(void) { return g_profile.count; }

int drt_profile_elapsed_ms(int32_t slot, float *ms_host) {
    if (ms_host == nullptr) return DRT_ERR_NULL_POINTER;
    if (slot < 0 || slot >= g_profile.count) return DRT_ERR_BAD_EXTENT;
    if (cudaEventSynchronize(g_profile.stop[slot]) != cudaSuccess) return DRT_ERR_CUDA;
    return cudaEventElapsedTime(ms_host, g_profile.start[slot], g_profile.stop[slot]) == cudaSuccess
               ? DRT_OK
               : DRT_ERR_CUDA;
}

size_t drt_compact_workspace_bytes(int64_t P) {
    if (P < 0) return 0;
    const int64_t nb = (P + kCompactChunk - 1) / kCompactChunk;
    return align256(size_t(nb > 0 ? nb : 1) * sizeof(int32_t)) +
           align256(size_t(nb > 0 ? nb : 1) * sizeof(int64_t));
}

int drt_compact_valid_paths(drt_stream_t stream, int64_t P, int32_t order, const float *vertices,
                            const int32_t *objects, const uint8_t *mask, int64_t capacity,
                            void *workspace, size_t workspace_bytes, int64_t *out_count,
                            int64_t *out_index, float *out_vertices, int32_t *out_objects) {
    if (P < 0 || order < 0 || capacity < 0) return DRT_ERR_BAD_EXTENT;
    if (out_count == nullptr) return DRT_ERR_NULL_POINTER;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (P == 0) {
        return cudaMemsetAsync(out_count, 0, sizeof(int64_t), s) == cudaSuccess ? DRT_OK : DRT_ERR_CUDA;
    }
    if (!mask || !workspace) return DRT_ERR_NULL_POINTER;
    if ((out_vertices && !vertices) || (out_objects && !objects)) return DRT_ERR_NULL_POINTER;
    if (workspace_bytes < drt_compact_workspace_bytes(P)) return DRT_ERR_WORKSPACE;
    const int64_t nb = (P + kCompactChunk - 1) / kCompactChunk;
    int32_t *counts = static_cast<int32_t *>(workspace);
    int64_t *offsets = reinterpret_cast<int64_t *>(static_cast<unsigned char *>(workspace) +
                                                   align256(size_t(nb) * sizeof(int32_t)));
    compact_count_kernel<<<unsigned(nb), kCompactThreads, 0, s>>>(P, mask, counts);
    compact_scan_kernel<<<1, 1024, 0, s>>>(nb, counts, offsets, out_count);
    compact_scatter_kernel<<<unsigned(nb), kCompactThreads, 0, s>>>(
        P, order + 2, vertices, objects, mask, offsets, capacity, out_index, out_vertices, out_objects);
    return cudaGetLastError() == cudaSuccess ? DRT_OK : DRT_ERR_CUDA;
}

int drt_complete_graph_candidates(drt_stream_t stream, int64_t num_nodes, int32_t order,
                                  int64_t start, int64_t count, int32_t stride_multiplier,
                                  int32_t *out) {
    if (num_nodes < 0 || order < 0 || start < 0 || count < 0) return DRT_ERR_BAD_EXTENT;
    if (order > 2 * DRT_MAX_ORDER) return DRT_ERR_UNSUPPORTED;
    if (count == 0 || order == 0) return DRT_OK;
    if (out == nullptr) return DRT_ERR_NULL_POINTER;
    complete_graph_candidates_kernel<<<unsigned((count + 255) / 256), 256, 0,
                                       static_cast<cudaStream_t>(stream)>>>(
        num_nodes, order, start, count, stride_multiplier > 0 ? stride_multiplier : 1, out);
    return cudaGetLastError() == cudaSuccess ? DRT_OK : DRT_ERR_CUDA;
}

}  // extern "C"
