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

constexpr int kStages = 4;
constexpr int kWarps = 8;
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
template <int RPW, int MODE, bool PATH, class Src, class Sink>
__global__ void __launch_bounds__(kThreads, 2)
intersect_kernel(const CoreParams p, const Src src, const Sink sink) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Tri48 *ring = reinterpret_cast<Tri48 *>(smem_raw);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + kRingBytes);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t num_units = p.num_units_dev ? *p.num_units_dev : p.num_units;
    const int64_t num_blocks = (num_units + kWarps - 1) / kWarps;
    if (static_cast<int64_t>(blockIdx.x) >= num_blocks) return;

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

    int64_t tests = 0;
    for (int64_t blk = blockIdx.x; blk < num_blocks; blk += gridDim.x) {
        const int64_t unit = blk * kWarps + warp;
        float3 o[RPW], d[RPW];
        uint32_t valid = 0;
        if (unit < num_units) valid = src.load(unit, o, d);
        uint32_t active = valid;  // warp-uniform mask of rays still being tested
        uint32_t hit_any = 0;     // MODE_ANY: warp-uniform mask of rays that hit
        float best_t[RPW];
        uint32_t best_key[RPW];
        int32_t best_idx[RPW];
        if (MODE == MODE_FIRST) {
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                best_t[r] = CUDART_INF_F;
                best_key[r] = 0xffffffffu;
                best_idx[r] = -1;
            }
        }

        for (int tl = 0; tl < NT; ++tl) {
            issue();
            const uint32_t stage = resident ? it % static_cast<uint32_t>(NT) : it % kStages;
            if (!resident || it < static_cast<uint32_t>(NT)) mbar_wait(&bars[stage], (it / kStages) & 1u);
            const uint32_t tile_index = it % static_cast<uint32_t>(NT);
            ++it;

            if (active) {
                const Tri48 *tile = ring + size_t(stage) * kTile;
                uint32_t lane_hits = 0;
#pragma unroll 2
                for (int j = lane; j < kTile; j += 32) {
                    const float4 a = tile[j].a, b = tile[j].b, c = tile[j].c;
                    const Tri tr = unpack(a, b, c);
#pragma unroll
                    for (int r = 0; r < RPW; ++r) {
                        if (active & (1u << r)) {
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
                if (MODE == MODE_ANY) {
                    const uint32_t m = __reduce_or_sync(kFull, lane_hits);
                    hit_any |= m;
                    active = (PATH && m) ? 0u : (active & ~m);
                }
            }
            // all warps are done with this stage; the producer may refill it next iteration
            if (__syncthreads_and(active == 0)) break;
        }

        if (unit < num_units) {
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
        }
    }

    // never exit with bulk copies still in flight
    if (!resident) {
        while (it < issued) {
            mbar_wait(&bars[it % kStages], (it / kStages) & 1u);
            ++it;
        }
    } else {
        while (it < issued) {
            mbar_wait(&bars[it % kStages], 0u);
            ++it;
        }
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
    const int grid = static_cast<int>(blocks < int64_t(sms) * 2 ? blocks : int64_t(sms) * 2);
    kern<<<grid, kThreads, kSmemBytes, stream>>>(p, src, sink);
    return cudaGetLastError();
}

}  // namespace drt
