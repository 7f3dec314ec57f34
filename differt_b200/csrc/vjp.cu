// K3b: reverse mode of the first-hit distance with the winning faces held fixed.
// Reference: differt/src/differt/geometry/_mesh.py:226-255 (`_differentiable_distance`) pulled back
// by jax.vjp in `_first_triangle_hit_by_ray_helper_bwd` (_mesh.py:308-338).
//   t = (q·e2) / (h·e1),  h = d × e2,  q = (o - v0) × e1;  rays with face == -1 or a == 0 give zero.
// One thread per ray; the three vertex cotangents are scatter-added with float atomics (several rays
// share a vertex), origins/directions are plain coalesced stores.
#include "common.cuh"

namespace drt {

__global__ void __launch_bounds__(256)
first_hit_vjp_kernel(int64_t R, int64_t V, int64_t T, const float *__restrict__ verts,
                     const int32_t *__restrict__ tris, const float *__restrict__ o,
                     const float *__restrict__ d, const int32_t *__restrict__ faces,
                     const float *__restrict__ g_t, float *g_verts, float *__restrict__ g_o,
                     float *__restrict__ g_d) {
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t r = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; r < R; r += stride) {
        float3 go = make_float3(0.f, 0.f, 0.f), gd = make_float3(0.f, 0.f, 0.f);
        const int32_t face = faces[r];
        const float g = g_t[r];
        if (face >= 0 && face < T && g != 0.0f) {
            int64_t vi[3];
#pragma unroll
            for (int q = 0; q < 3; ++q)
                vi[q] = min(max(int64_t(tris[3 * int64_t(face) + q]), int64_t(0)), V - 1);
            const float3 v0 = ld3(verts + 3 * vi[0]), v1 = ld3(verts + 3 * vi[1]),
                         v2 = ld3(verts + 3 * vi[2]);
            const float3 oo = ld3(o + 3 * r), dd = ld3(d + 3 * r);
            const float3 e1 = sub3(v1, v0), e2 = sub3(v2, v0);
            const float3 h = cross3(dd, e2);
            const float a = dot3(h, e1);
            if (a != 0.0f) {
                const float f = __frcp_rn(a);
                const float3 s = sub3(oo, v0);
                const float3 q = cross3(s, e1);
                const float qe2 = dot3(q, e2);
                const float g_f = g * qe2;
                const float g_qe2 = g * f;
                const float3 g_q = scale3(e2, g_qe2);
                float3 g_e2 = scale3(q, g_qe2);
                const float g_a = -(g_f * f) * f;
                const float3 g_h = scale3(e1, g_a);
                float3 g_e1 = scale3(h, g_a);
                gd = cross3(e2, g_h);
                g_e2 = add3(g_e2, cross3(g_h, dd));
                const float3 g_s = cross3(e1, g_q);
                g_e1 = add3(g_e1, cross3(g_q, s));
                go = g_s;
                const float3 g_v0 = sub3(sub3(make_float3(-g_s.x, -g_s.y, -g_s.z), g_e1), g_e2);
                atomicAdd(g_verts + 3 * vi[0], g_v0.x);
                atomicAdd(g_verts + 3 * vi[0] + 1, g_v0.y);
                atomicAdd(g_verts + 3 * vi[0] + 2, g_v0.z);
                atomicAdd(g_verts + 3 * vi[1], g_e1.x);
                atomicAdd(g_verts + 3 * vi[1] + 1, g_e1.y);
                atomicAdd(g_verts + 3 * vi[1] + 2, g_e1.z);
                atomicAdd(g_verts + 3 * vi[2], g_e2.x);
                atomicAdd(g_verts + 3 * vi[2] + 1, g_e2.y);
                atomicAdd(g_verts + 3 * vi[2] + 2, g_e2.z);
            }
        }
        st3(g_o + 3 * r, go);
        st3(g_d + 3 * r, gd);
    }
}

}  // namespace drt

extern "C" int drt_first_triangle_hit_by_ray_vjp(drt_stream_t stream, int64_t R, int64_t V, int64_t T,
                                                 const float *vertices, const int32_t *triangles,
                                                 const float *o, const float *d,
                                                 const int32_t *faces, const float *g_t,
                                                 float *g_vertices, float *g_o, float *g_d) {
    if (R < 0 || V < 0 || T < 0) return DRT_ERR_BAD_EXTENT;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (V > 0) {
        if (g_vertices == nullptr) return DRT_ERR_NULL_POINTER;
        if (cudaMemsetAsync(g_vertices, 0, size_t(V) * 12, s) != cudaSuccess) return DRT_ERR_CUDA;
    }
    if (R == 0) return DRT_OK;
    if (!g_o || !g_d) return DRT_ERR_NULL_POINTER;
    if (T == 0 || V == 0) {
        if (cudaMemsetAsync(g_o, 0, size_t(R) * 12, s) != cudaSuccess) return DRT_ERR_CUDA;
        if (cudaMemsetAsync(g_d, 0, size_t(R) * 12, s) != cudaSuccess) return DRT_ERR_CUDA;
        return DRT_OK;
    }
    if (!vertices || !triangles || !o || !d || !faces || !g_t) return DRT_ERR_NULL_POINTER;
    const int threads = 256;
    const int64_t blocks = (R + threads - 1) / threads;
    const unsigned grid = unsigned(blocks < int64_t(drt::device_sm_count()) * 16 ? blocks : int64_t(drt::device_sm_count()) * 16);
    drt::first_hit_vjp_kernel<<<grid, threads, 0, s>>>(R, V, T, vertices, triangles, o, d, faces, g_t,
                                                       g_vertices, g_o, g_d);
    return cudaGetLastError() == cudaSuccess ? DRT_OK : DRT_ERR_CUDA;
}
