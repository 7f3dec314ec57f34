// All-pairs ray × triangle engine shared by K2 (any-hit), K3 (first-hit), K4 (visibility) and the
// blockage stage of K6.
//
// Mapping (BASELINE.json north_star): one warp per work unit — a unit is RPW rays (flat kernels) or
// the k+1 segments of one path candidate (K6) — the 32 lanes stride over the triangles of a
// shared-memory tile of 512 packed triangles (24 KB).
//
// Tiles come from two places:
//   * HEAD tiles: the first DRT_HEAD_TILES tiles of the pack are loaded once per CTA (one TMA bulk
//     copy each) and stay resident.  Every new unit is tested against them first, in order.  With an
//     area-sorted any-hit pack (pack_sort.cu) these hold the largest triangles — the likeliest
//     blockers — so most blocked units are decided here and never touch the ring.
//   * RING tiles: the remaining tiles are streamed cyclically from the L2-resident pack by the TMA
//     engine (cp.async.bulk → mbarrier complete_tx) through DRT_STAGES slots, the copy of the next
//     tile overlapping the arithmetic on the current one.  any-hit / first-hit are order independent,
//     so a unit joins the ring at whatever tile is in flight and leaves after a full cycle.
//
// CTAs are persistent (grid = SMs × CTAs/SM) and walk the ring in lockstep — one barrier per step,
// which also returns the consumed slot to the TMA producer — but every WARP owns its own sequence of
// units (unit = global warp id + n · total warps): per step it processes exactly one tile (head or
// ring) for its current unit and moves to its next unit the moment the current one is decided.
//
// Each lane keeps one triangle in registers and tests it against the unit's RPW rays (register
// blocking: 48 B of shared memory traffic amortised over RPW tests, RPW independent dependency
// chains for ILP).  Reductions are warp votes/shuffles only: __any_sync / __reduce_or_sync for
// any-hit, __reduce_min_sync on (ordered t bits, tie key) for the nearest hit.
#pragma once

#include "common.cuh"

namespace drt {

#ifndef DRT_STAGES
#define DRT_STAGES 2
#endif
#ifndef DRT_WARPS
#define DRT_WARPS 16
#endif
#ifndef DRT_HEAD_TILES
#define DRT_HEAD_TILES 2
#endif
#ifndef DRT_UNROLL
#define DRT_UNROLL 2
#endif
#ifndef DRT_CTAS_PER_SM
#define DRT_CTAS_PER_SM 2
#endif
constexpr int kStages = DRT_STAGES;
constexpr int kWarps = DRT_WARPS;
constexpr int kCtasPerSm = DRT_CTAS_PER_SM;
constexpr int kUnroll = DRT_UNROLL;
// first-hit units carry three running minima per ray: 12 warps x 2 CTAs leaves them 80 registers
#ifndef DRT_WARPS_FIRST
#define DRT_WARPS_FIRST 12
#endif
template <int MODE>
struct WarpsFor {
    static constexpr int value = MODE == 1 /*MODE_FIRST*/ ? DRT_WARPS_FIRST : kWarps;
};
constexpr int kHead = DRT_HEAD_TILES;
constexpr size_t kHeadBytes = size_t(kHead) * kTile * sizeof(Tri48);
constexpr size_t kRingBytes = size_t(kStages) * kTile * sizeof(Tri48);
constexpr size_t kSmemBytes = kHeadBytes + kRingBytes + (kStages + 1) * sizeof(uint64_t);

enum : int { MODE_ANY = 0, MODE_FIRST = 1 };

struct CoreParams {
    const Tri48 *pack;
    int num_tiles;       // padded triangle count / kTile
    int64_t num_units;   // work units (warps' worth of rays)
    const int64_t *num_units_dev;  // if non-null, read the unit count from device memory
    float eps;
    float thr;           // 1 - hit_tol (MODE_ANY)
    int64_t batch_size;  // tie rule (MODE_FIRST); <= 0 → single batch
    int64_t num_triangles;
    int64_t *tests_done;  // nullable
};

// tie key of the reference's first-hit reduction (_utils.py:1865-1868, 1886): smaller wins.
__device__ __forceinline__ uint32_t tie_key(int64_t j, int64_t bs, int64_t T) {
    if (bs <= 0 || bs >= T) return static_cast<uint32_t>(j);
    const int64_t nb = (T + bs - 1) / bs;  // batches incl. the remainder batch
    const int64_t b = j / bs;
    return static_cast<uint32_t>((nb - 1 - b) * bs + (j - b * bs));
}

// One lane's share of a tile for an any-hit unit: returns the mask of rays that hit one of this
// lane's triangles; `rows` = 32-triangle rows evaluated.  PATH units stop at the first row in which
// any lane found a hit (the unit is decided).  FAST selects mt_any_fast; the caller falls back to
// the exact variant for the tile when `weird` comes back set on any lane.
template <int RPW, bool PATH, bool FAST>
__device__ __forceinline__ uint32_t scan_tile_any(const Tri48 *__restrict__ tile, const int lane,
                                                  const float3 (&o)[RPW], const float3 (&d)[RPW],
                                                  const uint32_t active, const float eps,
                                                  const float thr, int &rows, bool &weird) {
    uint32_t lane_hits = 0;
    rows = 0;
#pragma unroll kUnroll
    for (int j = lane; j < kTile; j += 32) {
        const float4 a = tile[j].a, b = tile[j].b, c = tile[j].c;
        const Tri tr = unpack(a, b, c);
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            // PATH units keep all their rays until the unit is decided: no per-ray branch
            if (PATH || (active & (1u << r))) {
                bool hit;
                if (FAST) {
                    hit = mt_any_fast(o[r], d[r], tr, eps, thr, weird);
                } else {
                    float t;
                    hit = mt_exact(o[r], d[r], tr, eps, t) && (t < thr);
                }
                lane_hits |= hit ? (1u << r) : 0u;
            }
        }
        ++rows;
        if (PATH && __any_sync(kFull, lane_hits != 0)) break;
    }
    return lane_hits;
}

template <int RPW, bool PATH>
__device__ __forceinline__ uint32_t scan_tile_any(const Tri48 *__restrict__ tile, const int lane,
                                                  const float3 (&o)[RPW], const float3 (&d)[RPW],
                                                  const uint32_t active, const float eps,
                                                  const float thr, const bool fast_ok, int &rows) {
    if (fast_ok) {
        bool weird = false;
        const uint32_t hits = scan_tile_any<RPW, PATH, true>(tile, lane, o, d, active, eps, thr, rows, weird);
        if (!__any_sync(kFull, weird)) return hits;
    }
    bool unused = false;
    return scan_tile_any<RPW, PATH, false>(tile, lane, o, d, active, eps, thr, rows, unused);
}

// Src:  __device__ uint32_t load(int64_t unit, float3 (&o)[RPW], float3 (&d)[RPW])  → active mask
// Sink: __device__ void any(int64_t unit, uint32_t hit_mask, uint32_t valid_mask)           (ANY)
//       __device__ void first(int64_t unit, int r, int32_t idx, float t)                    (FIRST)
//
// Scheduling: the CTA walks the tile ring in lockstep (one barrier per tile, which also hands the
// consumed slot back to the TMA producer), but every WARP owns its own sequence of work units
// (unit = global warp id + n * total warps) and moves on to its next unit the moment the current one
// is decided — after an early exit or after it has seen all NT tiles, wherever in the cycle that
// happens.  No warp idles while another one of its CTA still works on a long unit.
template <int RPW, int MODE, class Src, class Sink>
__global__ void __launch_bounds__(WarpsFor<MODE>::value * 32, kCtasPerSm)
intersect_kernel(const CoreParams p, const Src src, const Sink sink) {
    constexpr int kWarpsK = WarpsFor<MODE>::value;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Tri48 *head = reinterpret_cast<Tri48 *>(smem_raw);
    Tri48 *ring = reinterpret_cast<Tri48 *>(smem_raw + kHeadBytes);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + kHeadBytes + kRingBytes);  // [kStages] ring, [kStages] head

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t num_units = p.num_units_dev ? *p.num_units_dev : p.num_units;
    const int64_t total_warps = int64_t(gridDim.x) * kWarpsK;
    if (int64_t(blockIdx.x) * kWarpsK >= num_units) return;

    const int NT = p.num_tiles;
    const int NH = NT < kHead ? NT : kHead;  // resident head tiles
    const int NR = NT - NH;                  // tiles streamed through the ring
    constexpr uint32_t kTileBytes = kTile * sizeof(Tri48);

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s <= kStages; ++s) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    uint32_t it = 0;      // ring steps consumed (CTA uniform)
    uint32_t issued = 0;  // ring tile loads issued (CTA uniform)
    auto issue = [&]() {
        if (tid == 0) {
            const uint32_t stage = issued % kStages;
            const uint32_t tile = NH + issued % static_cast<uint32_t>(NR);
            mbar_arrive_expect_tx(&bars[stage], kTileBytes);
            bulk_g2s(ring + size_t(stage) * kTile, p.pack + size_t(tile) * kTile, kTileBytes,
                     &bars[stage]);
        }
        ++issued;
    };
    if (tid == 0 && NH > 0) {
        mbar_arrive_expect_tx(&bars[kStages], uint32_t(NH) * kTileBytes);
        for (int h = 0; h < NH; ++h)
            bulk_g2s(head + size_t(h) * kTile, p.pack + size_t(h) * kTile, kTileBytes, &bars[kStages]);
    }
    if (NR > 0) {
#pragma unroll
        for (int s = 0; s < kStages - 1; ++s) issue();
    }
    if (NH > 0) mbar_wait(&bars[kStages], 0u);

    // per-warp state of the unit in flight
    int64_t unit = int64_t(blockIdx.x) * kWarpsK + warp;
    bool live = unit < num_units;
    float3 o[RPW], d[RPW];
    uint32_t valid = 0, active = 0, hit_any = 0;
    int pos = 0;  // tiles of the pack this unit has been tested against (head tiles first)
    float best_t[RPW];
    uint32_t best_key[RPW];
    int32_t best_idx[RPW];
    auto begin_unit = [&]() {
        valid = src.load(unit, o, d);
        active = valid;
        hit_any = 0;
        pos = 0;
        if (MODE == MODE_FIRST) {
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                best_t[r] = CUDART_INF_F;
                best_key[r] = 0xffffffffu;
                best_idx[r] = -1;
            }
        }
    };
    if (live) begin_unit();

    const bool fast_ok = p.eps >= 1.17549435e-38f;  // FLT_MIN: precondition of mt_any_fast
    int64_t tests = 0;
    while (true) {
        uint32_t stage = 0, ring_tile = 0;
        if (NR > 0) {
            issue();
            stage = it % kStages;
            mbar_wait(&bars[stage], (it / kStages) & 1u);
            ring_tile = NH + it % static_cast<uint32_t>(NR);
            ++it;
        }

        if (live) {
            const bool in_head = pos < NH;
            const Tri48 *tile = in_head ? head + size_t(pos) * kTile : ring + size_t(stage) * kTile;
            const uint32_t tile_index = in_head ? uint32_t(pos) : ring_tile;
            uint32_t lane_hits = 0;
            int rows = 0;
            if (MODE == MODE_ANY) {
                lane_hits = scan_tile_any<RPW, false>(tile, lane, o, d, active, p.eps, p.thr, fast_ok, rows);
            } else {
                // first-hit: fast reciprocal first; pairs whose |a| is out of its range are skipped and
                // flagged, and the tile is then redone with the fully general test (idempotent)
                bool weird = !fast_ok;
                if (fast_ok) {
#pragma unroll kUnroll
                    for (int j = lane; j < kTile; j += 32) {
                        const float4 a = tile[j].a, b = tile[j].b, c = tile[j].c;
                        const Tri tr = unpack(a, b, c);
#pragma unroll
                        for (int r = 0; r < RPW; ++r) {
                            float t;
                            const bool hit = mt_first_fast(o[r], d[r], tr, p.eps, t, weird);
                            if (hit && t <= best_t[r] && (active & (1u << r))) {
                                const int64_t gj = int64_t(tile_index) * kTile + j;
                                const uint32_t key = tie_key(gj, p.batch_size, p.num_triangles);
                                if (t < best_t[r] || key < best_key[r]) {
                                    best_t[r] = t;
                                    best_key[r] = key;
                                    best_idx[r] = static_cast<int32_t>(gj);
                                }
                            }
                        }
                    }
                    weird = __any_sync(kFull, weird);
                }
                if (weird) {
#pragma unroll kUnroll
                    for (int j = lane; j < kTile; j += 32) {
                        const float4 a = tile[j].a, b = tile[j].b, c = tile[j].c;
                        const Tri tr = unpack(a, b, c);
#pragma unroll
                        for (int r = 0; r < RPW; ++r) {
                            if (active & (1u << r)) {
                                float t;
                                const bool hit = mt_exact(o[r], d[r], tr, p.eps, t);
                                if (hit && t <= best_t[r]) {
                                    const int64_t gj = int64_t(tile_index) * kTile + j;
                                    const uint32_t key = tie_key(gj, p.batch_size, p.num_triangles);
                                    if (t < best_t[r] || key < best_key[r]) {
                                        best_t[r] = t;
                                        best_key[r] = key;
                                        best_idx[r] = static_cast<int32_t>(gj);
                                    }
                                }
                            }
                        }
                    }
                }
                rows = kTile / 32;
            }
            tests += int64_t(__popc(active)) * 32 * rows;
            ++pos;
            if (MODE == MODE_ANY) {
                const uint32_t m = __reduce_or_sync(kFull, lane_hits) & active;
                hit_any |= m;
                active &= ~m;
            }
            if (pos == NT || active == 0) {  // unit decided: emit, move on to this warp's next unit
                if (MODE == MODE_ANY) {
                    if (lane == 0) sink.any(unit, hit_any, valid);
                } else {
#pragma unroll
                    for (int r = 0; r < RPW; ++r) {
                        // -0 and +0 are the same distance (they tie and the index decides): + 0.0f
                        // canonicalises the sign before the bit-pattern comparison
                        const uint32_t tb = float_order_bits(best_t[r] + 0.0f);
                        const uint32_t tmin = __reduce_min_sync(kFull, tb);
                        const uint32_t key = (tb == tmin) ? best_key[r] : 0xffffffffu;
                        const uint32_t kmin = __reduce_min_sync(kFull, key);
                        const uint32_t owner = __ballot_sync(kFull, tb == tmin && key == kmin);
                        const int src_lane = __ffs(owner) - 1;
                        const int32_t idx = __shfl_sync(kFull, best_idx[r], src_lane);
                        const float t = __shfl_sync(kFull, best_t[r], src_lane);
                        if (lane == 0 && (valid & (1u << r))) sink.first(unit, r, idx, t);
                    }
                }
                unit += total_warps;
                live = unit < num_units;
                if (live) begin_unit();
            }
        }
        // every warp is done with this ring stage (the producer refills it next step); stop when no
        // warp of the CTA has a unit left
        if (!__syncthreads_or(live)) break;
    }

    // never exit with bulk copies still in flight
    while (it < issued) {
        mbar_wait(&bars[it % kStages], (it / kStages) & 1u);
        ++it;
    }
    if (p.tests_done != nullptr) {
        // one atomic per warp (lanes hold identical counts)
        if (lane == 0 && tests) atomicAdd(reinterpret_cast<unsigned long long *>(p.tests_done),
                                          static_cast<unsigned long long>(tests));
    }
}

template <int RPW, int MODE, class Src, class Sink>
inline cudaError_t launch_intersect(cudaStream_t stream, const CoreParams &p, const Src &src,
                                    const Sink &sink, int64_t max_units) {
    auto kern = intersect_kernel<RPW, MODE, Src, Sink>;
    static unsigned long long configured = 0;  // per-device bits (common.cuh)
    if (cudaError_t e = ensure_dynamic_smem(kern, kSmemBytes, configured); e != cudaSuccess) return e;
    const int sms = device_sm_count();
    constexpr int kWarpsK = WarpsFor<MODE>::value;
    const int64_t blocks = (max_units + kWarpsK - 1) / kWarpsK;
    if (blocks <= 0) return cudaSuccess;
    const int64_t resident_ctas = int64_t(sms) * kCtasPerSm;
    const int grid = static_cast<int>(blocks < resident_ctas ? blocks : resident_ctas);
    kern<<<grid, kWarpsK * 32, kSmemBytes, stream>>>(p, src, sink);
    return cudaGetLastError();
}

}  // namespace drt
