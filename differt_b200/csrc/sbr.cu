// N3: shooting-and-bouncing rays.  One bounce step of SBRPathLauncher.launch_paths
// (reference: differt/src/differt/geometry/_solvers.py:279-356 bounce_rays / filter_rays, scan body
// :407-444) and one step of the multipath-lifetime-map kernel (differt/src/differt/geometry/_scene.py:
// 81-171).  The nearest hit of every ray is found by K3 (first_triangle_hit_by_ray, the all-pairs
// engine) between the steps; these kernels are the element-wise remainder, fused per bounce.
#include "common.cuh"

namespace drt {

// filter_rays (:320-356) for every receiver, then bounce_rays (:279-318), one thread per (tx, ray).
// masks: [num_tx, num_rx, num_rays] for THIS bounce.  origins / directions / valid updated in place.
__global__ void __launch_bounds__(256)
sbr_bounce_kernel(int64_t num_tx, int64_t num_rays, int64_t num_rx, int64_t T,
                  const Tri48 *__restrict__ pack, float *__restrict__ origins,
                  float *__restrict__ directions, uint8_t *__restrict__ valid,
                  const int32_t *__restrict__ faces, const float *__restrict__ t_hit,
                  const float *__restrict__ rx, float max_dist, uint8_t *__restrict__ masks,
                  float *__restrict__ vertices_out) {
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= num_tx * num_rays) return;
    const int64_t itx = i / num_rays, iray = i - itx * num_rays;
    float3 o = ld3(origins + 3 * i), d = ld3(directions + 3 * i);
    const float th = t_hit[i];
    bool ok = valid[i] != 0;

    // receivers in the vicinity of the segment [o, o + th d)
    uint8_t *m = masks + itx * num_rx * num_rays + iray;
    for (int64_t r = 0; r < num_rx; ++r) {
        const float3 v = sub3(ld3(rx + 3 * r), o);
        const float3 c = cross3(d, v);
        const float dist2 = (c.x * c.x + c.y * c.y) + c.z * c.z;
        const float t_rx = dot3(d, v);
        const bool near = (t_rx > 0.0f) && (t_rx < th) && ok && (dist2 < max_dist);
        m[r * num_rays] = near ? 1 : 0;
    }

    // bounce
    const bool inside = isfinite(th);
    ok = ok && inside;
    const float t = inside ? th : 0.0f;
    o = make_float3(o.x + t * d.x, o.y + t * d.y, o.z + t * d.z);
    int64_t f = faces[i];
    if (f < 0) f += T;  // jnp.take wraps negative indices (a miss reports face -1): last triangle
    f = f < 0 ? 0 : (f >= T ? T - 1 : f);
    const float4 tc = pack[f].c;
    const float3 n = make_float3(tc.y, tc.z, tc.w);
    const float k = 2.0f * dot3(d, n);
    d = make_float3(d.x - k * n.x, d.y - k * n.y, d.z - k * n.z);
    st3(origins + 3 * i, o);
    st3(directions + 3 * i, d);
    valid[i] = ok ? 1 : 0;
    if (vertices_out != nullptr) st3(vertices_out + 3 * i, o);
}

// ---- multipath lifetime map (reference _scene.py:62-171) -------------------------------------------

__device__ __forceinline__ uint32_t combine_hashes(uint32_t h1, uint32_t h2) {  // _scene.py:67-69
    return h1 ^ (h2 + 0x9E3779B9u + (h1 << 6) + (h1 >> 2));
}
__device__ __forceinline__ uint32_t hash_int(uint32_t x) {  // _scene.py:74-78
    x = ((x >> 16) ^ x) * 0x045D9F3Bu;
    x = ((x >> 16) ^ x) * 0x045D9F3Bu;
    return (x >> 16) ^ x;
}

// One iteration t of the loop of _compute_tx_mlm_kernel (_scene.py:108-171) given the first hit
// (face, distance from the query origin) of every live ray.  query origins / directions / hashes /
// alive flags are updated in place; output [num_tx, dim_x, dim_y] receives the atomic ORs.
__global__ void __launch_bounds__(256)
mlm_step_kernel(int64_t num_tx, int64_t num_rays, int64_t T, const Tri48 *__restrict__ pack,
                float *__restrict__ origins, float *__restrict__ directions,
                uint32_t *__restrict__ hashes, uint8_t *__restrict__ alive,
                const int32_t *__restrict__ faces, const float *__restrict__ t_first, int iteration,
                int min_order, int assume_quads, float receiver_height, float min_x, float max_x,
                float min_y, float max_y, int dim_x, int dim_y, float epsilon,
                uint32_t *__restrict__ output) {
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= num_tx * num_rays || alive[i] == 0) return;
    const int64_t itx = i / num_rays;
    const float3 qo = ld3(origins + 3 * i);  // already offset by epsilon * d for iteration > 0
    float3 d = ld3(directions + 3 * i);
    const int32_t face = faces[i];
    const bool hit = face >= 0;
    const float res_t = t_first[i];
    float t_hit = CUDART_INF_F;
    if (hit) t_hit = iteration > 0 ? res_t + epsilon : res_t;
    const float dx = (max_x - min_x) / float(dim_x), dy = (max_y - min_y) / float(dim_y);
    const uint32_t h = hashes[i];
    if (fabsf(d.z) > 1e-6f) {
        const float u = __fdiv_rn(receiver_height - qo.z, d.z);
        const float3 P = make_float3(qo.x + d.x * u, qo.y + d.y * u, qo.z + d.z * u);
        if (u > 0.0f && u < t_hit && iteration >= min_order && P.x >= min_x && P.x <= max_x &&
            P.y >= min_y && P.y <= max_y) {
            int ix = int(floorf(__fdiv_rn(P.x - min_x, dx)));
            int iy = int(floorf(__fdiv_rn(P.y - min_y, dy)));
            ix = ix < 0 ? 0 : (ix > dim_x - 1 ? dim_x - 1 : ix);
            iy = iy < 0 ? 0 : (iy > dim_y - 1 ? dim_y - 1 : iy);
            atomicOr(output + (itx * dim_x + ix) * dim_y + iy, h);
        }
    }
    if (!hit) {
        alive[i] = 0;  // the ray leaves the scene
        return;
    }
    float3 o = make_float3(qo.x + d.x * res_t, qo.y + d.y * res_t, qo.z + d.z * res_t);
    const int64_t f = face >= T ? T - 1 : face;
    const float4 tc = pack[f].c;
    const float3 n = make_float3(tc.y, tc.z, tc.w);
    const float k = 2.0f * dot3(d, n);
    d = make_float3(d.x - k * n.x, d.y - k * n.y, d.z - k * n.z);
    hashes[i] = combine_hashes(h, hash_int(uint32_t(assume_quads ? face / 2 : face)));
    // next query origin: current origin + direction * epsilon (_scene.py:110-112)
    o = make_float3(o.x + d.x * epsilon, o.y + d.y * epsilon, o.z + d.z * epsilon);
    st3(origins + 3 * i, o);
    st3(directions + 3 * i, d);
}

}  // namespace drt

using namespace drt;

extern "C" {

int drt_sbr_bounce(drt_stream_t stream, int64_t num_tx, int64_t num_rays, int64_t num_rx,
                   int64_t num_triangles, const void *pack, float *origins, float *directions,
                   uint8_t *valid, const int32_t *faces, const float *t_hit, const float *rx,
                   float max_dist, uint8_t *masks, float *vertices_out) {
    if (num_tx < 0 || num_rays < 0 || num_rx < 0 || num_triangles < 0) return DRT_ERR_BAD_EXTENT;
    const int64_t n = num_tx * num_rays;
    if (n == 0) return DRT_OK;
    if (num_triangles == 0) return DRT_ERR_BAD_EXTENT;  // the caller handles empty meshes
    if (!pack || !origins || !directions || !valid || !faces || !t_hit) return DRT_ERR_NULL_POINTER;
    if (num_rx > 0 && (!rx || !masks)) return DRT_ERR_NULL_POINTER;
    sbr_bounce_kernel<<<unsigned((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        num_tx, num_rays, num_rx, num_triangles, static_cast<const Tri48 *>(pack), origins, directions,
        valid, faces, t_hit, rx, max_dist, masks, vertices_out);
    return cudaGetLastError() == cudaSuccess ? DRT_OK : DRT_ERR_CUDA;
}

int drt_mlm_step(drt_stream_t stream, int64_t num_tx, int64_t num_rays, int64_t num_triangles,
                 const void *pack, float *origins, float *directions, uint32_t *hashes,
                 uint8_t *alive, const int32_t *faces, const float *t_first, int32_t iteration,
                 int32_t min_order, int32_t assume_quads, float receiver_height, float min_x,
                 float max_x, float min_y, float max_y, int32_t dim_x, int32_t dim_y, float epsilon,
                 uint32_t *output) {
    if (num_tx < 0 || num_rays < 0 || num_triangles < 0 || dim_x <= 0 || dim_y <= 0) return DRT_ERR_BAD_EXTENT;
    const int64_t n = num_tx * num_rays;
    if (n == 0) return DRT_OK;
    if (!origins || !directions || !hashes || !alive || !faces || !t_first || !output) return DRT_ERR_NULL_POINTER;
    if (num_triangles > 0 && !pack) return DRT_ERR_NULL_POINTER;
    mlm_step_kernel<<<unsigned((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        num_tx, num_rays, num_triangles, static_cast<const Tri48 *>(pack), origins, directions, hashes,
        alive, faces, t_first, iteration, min_order, assume_quads, receiver_height, min_x, max_x, min_y,
        max_y, dim_x, dim_y, epsilon, output);
    return cudaGetLastError() == cudaSuccess ? DRT_OK : DRT_ERR_CUDA;
}

}  // extern "C"
