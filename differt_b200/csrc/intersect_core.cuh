// All-pairs ray × triangle engine shared by K2 (any-hit), K3 (first-hit), K4 (visibility) and the
// blockage stage of K6.
//
// Mapping (BASELINE.json north_star): one warp per work unit — a unit is RPW rays (flat kernels) or
// the k+1 segments of one path candidate (K6) — the 32 lanes stride over the triangles of a
// shared-memory tile; tiles of 512 packed triangles (24 KB) are streamed from the L2-resident packed
// mesh by the TMA engine (cp.async.bulk → mbarrier complete_tx) through a 4-deep ring, so the copy of
// tile i+3 overlaps the arithmetic on tile i.  When the whole mesh fits in the ring it is loaded
// once per CTA and stays resident.  CTAs are persistent (grid = SMs × 2) and keep walking the ring
// cyclically across work blocks: any-hit / first-hit are order independent, so a block may start at
// whatever tile is next in flight — an early exit never drains or restarts the pipeline.
//
// Each lane keeps one triangle in registers and tests it against the unit's RPW rays (register
// blocking: 48 B of shared memory traffic amortised over RPW tests, RPW independent dependency
// chains for ILP).  Reductions are warp shuffles only: __reduce_or_sync for any-hit,
// __reduce_min_sync on (ordered t bits, tie key) for the nearest hit.
#pragma once

#include "common.cuh"

namespace drt {

#ifndef DRT_STAGES
#define DRT_STAGES 2
#endif
#ifndef DRT_WARPS
#define DRT_WARPS 16
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
constexpr int kThreads = kWarps * 32;
constexpr size_t kRingBytes = size_t(kStages) * kTile * sizeof(Tri48);
constexpr size_t kSmemBytes = kRingBytes + kStages * sizeof(uint64_t);

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

// Src:  __device__ uint32_t load(int64_t unit, float3 (&o)[RPW], float3 (&d)[RPW])  → active mask
// Sink: __device__ void any(int64_t unit, uint32_t hit_mask, uint32_t valid_mask)           (ANY)
//       __device__ void first(int64_t unit, int r, int32_t idx, float t)                    (FIRST)
// PATH = true: the unit is finished as soon as any of its rays hits (K6 blockage).
//
// Scheduling: the CTA walks the tile ring in lockstep (one barrier per tile, which also hands the
// consumed slot back to the TMA producer), but every WARP owns its own sequence of work units
// (unit = global warp id + n * total warps) and moves on to its next unit the moment the current one
// is decided — after an early exit or after it has seen all NT tiles, wherever in the cycle that
// happens.  No warp idles while another one of its CTA still works on a long unit.
template <int RPW, int MODE, bool PATH, class Src, class Sink>
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
intersect_kernel(const CoreParams p, const Src src, const Sink sink) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Tri48 *ring = reinterpret_cast<Tri48 *>(smem_raw);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + kRingBytes);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t num_units = p.num_units_dev ? *p.num_units_dev : p.num_units;
    const int64_t total_warps = int64_t(gridDim.x) * kWarps;
    if (int64_t(blockIdx.x) * kWarps >= num_units) return;

    const int NT = p.num_tiles;
    const bool resident = NT <= kStages;
    constexpr uint32_t kTileBytes = kTile * sizeof(Tri48);

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    uint32_t it = 0;      // ring iterations consumed (CTA uniform)
    uint32_t issued = 0;  // tile loads issued (CTA uniform)
    auto issue = [&]() {
        if (resident && issued >= static_cast<uint32_t>(NT)) return;
        if (tid == 0) {
            const uint32_t stage = issued % kStages;
            const uint32_t tile = issued % static_cast<uint32_t>(NT);
            mbar_arrive_expect_tx(&bars[stage], kTileBytes);
            bulk_g2s(ring + size_t(stage) * kTile, p.pack + size_t(tile) * kTile, kTileBytes,
                     &bars[stage]);
        }
        ++issued;
    };
#pragma unroll
    for (int s = 0; s < kStages - 1; ++s) issue();

    // per-warp state of the unit in flight
    int64_t unit = int64_t(blockIdx.x) * kWarps + warp;
    bool live = unit < num_units;
    float3 o[RPW], d[RPW];
    uint32_t valid = 0, active = 0, hit_any = 0;
    int seen = 0;
    float best_t[RPW];
    uint32_t best_key[RPW];
    int32_t best_idx[RPW];
    auto begin_unit = [&]() {
        valid = src.load(unit, o, d);
        active = valid;
        hit_any = 0;
        seen = 0;
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

    int64_t tests = 0;
    while (true) {
        issue();
        const uint32_t stage = resident ? it % static_cast<uint32_t>(NT) : it % kStages;
        if (!resident || it < static_cast<uint32_t>(NT)) mbar_wait(&bars[stage], (it / kStages) & 1u);
        const uint32_t tile_index = it % static_cast<uint32_t>(NT);
        ++it;

        if (live) {
            const Tri48 *tile = ring + size_t(stage) * kTile;
            uint32_t lane_hits = 0;
#pragma unroll kUnroll
            for (int j = lane; j < kTile; j += 32) {
                const float4 a = tile[j].a, b = tile[j].b, c = tile[j].c;
                const Tri tr = unpack(a, b, c);
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    // PATH units keep all their rays until the unit is decided: no per-ray branch
                    if (PATH || (active & (1u << r))) {
                        float t;
                        const bool hit = mt_exact(o[r], d[r], tr, p.eps, t);
                        if (MODE == MODE_ANY) {
                            lane_hits |= (hit && (t < p.thr)) ? (1u << r) : 0u;
                        } else {
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
            tests += int64_t(__popc(active)) * kTile;
            ++seen;
            if (MODE == MODE_ANY) {
                const uint32_t m = __reduce_or_sync(kFull, lane_hits) & active;
                hit_any |= m;
                active = (PATH && m) ? 0u : (active & ~m);
            }
            if (seen == NT || active == 0) {  // unit decided: emit, move on to this warp's next unit
                if (MODE == MODE_ANY) {
                    if (lane == 0) sink.any(unit, hit_any, valid);
                } else {
#pragma unroll
                    for (int r = 0; r < RPW; ++r) {
                        const uint32_t tb = float_order_bits(best_t[r]);
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
        // every warp is done with this stage (the producer refills it next iteration); stop when no
        // warp of the CTA has a unit left
        if (!__syncthreads_or(live)) break;
    }

    // never exit with bulk copies still in flight
    while (it < issued) {
        mbar_wait(&bars[it % kStages], resident ? 0u : ((it / kStages) & 1u));
        ++it;
    }
    if (p.tests_done != nullptr) {
        // one atomic per warp (lanes hold identical counts)
        if (lane == 0 && tests) atomicAdd(reinterpret_cast<unsigned long long *>(p.tests_done),
                                          static_cast<unsigned long long>(tests));
    }
}

template <int RPW, int MODE, bool PATH, class Src, class Sink>
inline cudaError_t launch_intersect(cudaStream_t stream, const CoreParams &p, const Src &src,
                                    const Sink &sink, int64_t max_units) {
    auto kern = intersect_kernel<RPW, MODE, PATH, Src, Sink>;
    static bool configured = false;  // benign race: idempotent attribute set
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(kSmemBytes));
        if (e != cudaSuccess) return e;
        configured = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t blocks = (max_units + kWarps - 1) / kWarps;
    if (blocks <= 0) return cudaSuccess;
    const int64_t resident_ctas = int64_t(sms) * kCtasPerSm;
    const int grid = static_cast<int>(blocks < resident_ctas ? blocks : resident_ctas);
    kern<<<grid, kThreads, kSmemBytes, stream>>>(p, src, sink);
    return cudaGetLastError();
}

}  // namespace drt
