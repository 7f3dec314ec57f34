// First EM consumer of the traced paths (second half): Fresnel coefficients and the
// per-path field chain, fused with the accumulation per (transmitter, receiver) pair.
// Reference: differt/src/differt/em/_fresnel.py:183-213 (`fresnel_coefficients`), em/_utils.py:243-262
// (`sp_directions`), :289-302 (`sp_rotation_matrix`), geometry/_utils.py:66-72, 100-108 (`normalize`,
// `perpendicular_vector`), and the composition of those pieces in plugins/deepmimo.py:348-405, 516-665
// (spherical basis, slab coefficients, J = R_out diag(r_s, r_p) R_in per interaction, projection on the
// receive polarisation, spreading 1/s, phase exp(-j 2 pi f s / c), lambda / 4 pi).
//
// One thread per path, everything in registers; the complex coefficient of a path is added to its
// (tx, rx) pair's field with two float atomics.  Parity with the oracle (oracle/em_oracle.py) is to a
// tolerance, not bit-exact: the chain goes through complex square roots, acos / atan2 and sin / cos of
// phases of 10^4 rad, whose last bits are libm-specific in the reference as well.
#include "common.cuh"

namespace drt {

struct cplx {
    float re, im;
};
__device__ __forceinline__ cplx cmk(float re, float im) { return cplx{re, im}; }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return cmk(a.re + b.re, a.im + b.im); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return cmk(a.re - b.re, a.im - b.im); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return cmk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
__device__ __forceinline__ cplx cscale(cplx a, float s) { return cmk(a.re * s, a.im * s); }
// utils.py:60-67 safe_divide: 0 where the denominator is exactly 0 (Smith's algorithm otherwise)
__device__ __forceinline__ cplx cdiv_safe(cplx a, cplx b) {
    if (b.re == 0.0f && b.im == 0.0f) return cmk(0.0f, 0.0f);
    if (fabsf(b.re) >= fabsf(b.im)) {
        const float r = __fdiv_rn(b.im, b.re), den = b.re + b.im * r;
        return cmk(__fdiv_rn(a.re + a.im * r, den), __fdiv_rn(a.im - a.re * r, den));
    }
    const float r = __fdiv_rn(b.re, b.im), den = b.re * r + b.im;
    return cmk(__fdiv_rn(a.re * r + a.im, den), __fdiv_rn(a.im * r - a.re, den));
}
// principal square root
__device__ __forceinline__ cplx csqrt_(cplx z) {
    if (z.re == 0.0f && z.im == 0.0f) return cmk(0.0f, z.im);
    const float r = hypotf(z.re, z.im);
    if (z.re >= 0.0f) {
        const float t = __fsqrt_rn(0.5f * (r + z.re));
        return cmk(t, __fdiv_rn(z.im, 2.0f * t));
    }
    const float t = __fsqrt_rn(0.5f * (r - z.re));
    return cmk(__fdiv_rn(fabsf(z.im), 2.0f * t), copysignf(t, z.im));
}

// em/_fresnel.py:183-213
__device__ __forceinline__ void fresnel(cplx n_r, float cos_theta_i, cplx &r_s, cplx &r_p, cplx &t_s, cplx &t_p) {
    const float c = fabsf(cos_theta_i);
    const cplx n2 = cmul(n_r, n_r);
    const float c2 = c * c;
    const cplx n2c = cscale(n2, c);
    const cplx nct = csqrt_(cmk((n2.re + c2) - 1.0f, n2.im));
    const cplx cc = cmk(c, 0.0f);
    const float two_c = 2.0f * c;
    r_s = cdiv_safe(csub(cc, nct), cadd(cc, nct));
    t_s = cdiv_safe(cmk(two_c, 0.0f), cadd(cc, nct));
    r_p = cdiv_safe(csub(n2c, nct), cadd(n2c, nct));
    t_p = cdiv_safe(cscale(n_r, two_c), cadd(n2c, nct));
}

__global__ void __launch_bounds__(256)
em_fresnel_kernel(int64_t n, const float *__restrict__ n_r, int64_t n_r_stride, const float *__restrict__ cos_theta,
                  int64_t cos_stride, float *__restrict__ r_s, float *__restrict__ r_p, float *__restrict__ t_s,
                  float *__restrict__ t_p) {
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    cplx a, b, c, d;
    fresnel(cmk(n_r[2 * i * n_r_stride], n_r[2 * i * n_r_stride + 1]), cos_theta[i * cos_stride], a, b, c, d);
    if (r_s) r_s[2 * i] = a.re, r_s[2 * i + 1] = a.im;
    if (r_p) r_p[2 * i] = b.re, r_p[2 * i + 1] = b.im;
    if (t_s) t_s[2 * i] = c.re, t_s[2 * i + 1] = c.im;
    if (t_p) t_p[2 * i] = d.re, t_p[2 * i + 1] = d.im;
}

// geometry/_utils.py:66-72: x / |x| (divide by 1 when |x| = 0) and |x|
__device__ __forceinline__ float3 normalize3(float3 v, float &len) {
    len = __fsqrt_rn((v.x * v.x + v.y * v.y) + v.z * v.z);
    const float d = len == 0.0f ? 1.0f : len;
    return make_float3(__fdiv_rn(v.x, d), __fdiv_rn(v.y, d), __fdiv_rn(v.z, d));
}
// geometry/_utils.py:100-108
__device__ __forceinline__ float3 perpendicular3(float3 u) {
    const float3 v = fabsf(u.x) > fabsf(u.y) ? make_float3(-u.y, u.x, 0.0f) : make_float3(0.0f, -u.z, u.y);
    float l;
    return normalize3(cross3(u, v), l);
}
// plugins/deepmimo.py:348-363: theta = acos(clip(k.z)), phi = atan2(k.y, k.x) and their sines / cosines,
// evaluated algebraically (cos theta = z, sin theta = sqrt((1 - z)(1 + z)), cos phi = x / rho, sin phi = y / rho:
// each within 2 ulp of the exact value, like the libm chain it replaces, at a fifth of the instructions)
__device__ __forceinline__ void spherical_basis(float3 k, float3 &theta_hat, float3 &phi_hat) {
    const float ct = fminf(fmaxf(k.z, -1.0f), 1.0f);
    const float st = __fsqrt_rn((1.0f - ct) * (1.0f + ct));
    const float rho = __fsqrt_rn(k.x * k.x + k.y * k.y);
    float cp, sp;
    if (rho > 0.0f) {
        cp = __fdiv_rn(k.x, rho), sp = __fdiv_rn(k.y, rho);
    } else {  // atan2(+-0, +0) = +-0, atan2(+-0, -0) = +-pi
        cp = signbit(k.x) ? -1.0f : 1.0f, sp = 0.0f;
    }
    theta_hat = make_float3(ct * cp, ct * sp, -st);
    phi_hat = make_float3(-sp, cp, 0.0f);
}

// Sum of x over the lanes of `peers` (the lanes of this warp holding the same key), for any peer pattern, in
// log2(32) shuffle rounds; every lane of the warp must call it.  The lowest lane of each group ends with the sum.
__device__ __forceinline__ void reduce_peers3(unsigned peers, int lane, float &x, float &y, float &z) {
    int rel = __popc(peers & ((1u << lane) - 1u));   // rank among the peers
    peers &= (0xfffffffeu << lane);                  // peers above this lane
    while (__any_sync(0xffffffffu, peers != 0u)) {
        const int next = __ffs(peers);               // 1 + lane of the next peer (0: none)
        const int src = next ? next - 1 : lane;
        const float tx = __shfl_sync(0xffffffffu, x, src), ty = __shfl_sync(0xffffffffu, y, src),
                    tz = __shfl_sync(0xffffffffu, z, src);
        if (next) x += tx, y += ty, z += tz;
        const bool done = rel & 1;                   // odd ranks have been absorbed by their left neighbour
        peers &= __ballot_sync(0xffffffffu, !done);
        rel >>= 1;
    }
}

// em/_utils.py:243-262 on flat [n, 3] operands (the host broadcasts)
__global__ void __launch_bounds__(256)
em_sp_directions_kernel(int64_t n, const float *__restrict__ k_i, const float *__restrict__ k_r,
                        const float *__restrict__ normals, float *__restrict__ e_i_s, float *__restrict__ e_i_p,
                        float *__restrict__ e_r_p) {
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    const float3 ki = ld3(k_i + 3 * i), kr = ld3(k_r + 3 * i), nrm = ld3(normals + 3 * i);
    float nl, l2;
    float3 s = normalize3(cross3(ki, nrm), nl);
    if (nl == 0.0f) s = perpendicular3(ki);
    const float3 pi = normalize3(cross3(s, ki), l2);
    const float3 pr = normalize3(cross3(s, kr), l2);
    e_i_s[3 * i] = s.x, e_i_s[3 * i + 1] = s.y, e_i_s[3 * i + 2] = s.z;
    e_i_p[3 * i] = pi.x, e_i_p[3 * i + 1] = pi.y, e_i_p[3 * i + 2] = pi.z;
    e_r_p[3 * i] = pr.x, e_r_p[3 * i + 1] = pr.y, e_r_p[3 * i + 2] = pr.z;
}

struct EmArgs {
    int64_t n;
    const float *vertices;     // [n, K + 2, 3]
    const int32_t *objects;    // [n, K + 2]
    int64_t T;
    const Tri48 *pack;         // unit normals in c.yzw (Mesh.normals, _mesh.py:950-956)
    const float *n_r;          // [T, 2] complex relative refractive index per triangle
    const float *thickness;    // [T] or null (half spaces)
    float wavelength;          // c / f, rounded once from double like the reference's Python scalars
    float omega;               // -2 pi f
    float scale;               // wavelength / 4 pi
    int tx_pol, rx_pol;        // 0 = V, 1 = H
    float *out_a;              // [n, 2] or null
    float *out_length;         // [n] or null
    const int64_t *pair_index; // [n] or null
    int64_t num_pairs;
    float *field;              // [num_pairs, 2] or null (atomic accumulate)
    float *power;              // [num_pairs] or null (atomic accumulate of |a|^2)
};

template <int K>
__global__ void __launch_bounds__(128) em_path_kernel(const EmArgs a) {
    const int64_t p0 = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    const bool live = p0 < a.n;
    const int64_t p = live ? p0 : a.n - 1;  // idle lanes of the last warp recompute the last path: they take part in the shuffles
    const float *v = a.vertices + p * (K + 2) * 3;
    float3 prev = ld3(v);
    float s_tot = 0.0f;
    // field in the (theta_hat, phi_hat) basis of the current segment
    cplx e0 = a.tx_pol == 0 ? cmk(1.0f, 0.0f) : cmk(0.0f, 0.0f);
    cplx e1 = a.tx_pol == 0 ? cmk(0.0f, 0.0f) : cmk(1.0f, 0.0f);
    float3 k_in, th_in, ph_in;
    {
        const float3 next = ld3(v + 3);
        float len;
        k_in = normalize3(sub3(next, prev), len);
        s_tot = len;
        spherical_basis(k_in, th_in, ph_in);
        prev = next;
    }
    const float wavelength = a.wavelength;
#pragma unroll
    for (int i = 0; i < K; ++i) {
        const float3 next = ld3(v + 3 * (i + 2));
        float len;
        const float3 k_out = normalize3(sub3(next, prev), len);
        s_tot += len;
        prev = next;
        float3 th_out, ph_out;
        spherical_basis(k_out, th_out, ph_out);
        int32_t t = a.objects[p * (K + 2) + i + 1];
        t = min(max(t, 0), int32_t(a.T - 1));
        const float4 pc = a.pack[t].c;
        const float3 nrm = make_float3(pc.y, pc.z, pc.w);
        // em/_utils.py:243-262
        float nl;
        float3 e_i_s = normalize3(cross3(k_in, nrm), nl);
        if (nl == 0.0f) e_i_s = perpendicular3(k_in);
        float l2;
        const float3 e_i_p = normalize3(cross3(e_i_s, k_in), l2);
        const float3 e_r_p = normalize3(cross3(e_i_s, k_out), l2);
        const float cos_i = -((nrm.x * k_in.x + nrm.y * k_in.y) + nrm.z * k_in.z);
        // plugins/deepmimo.py:390-405
        const cplx nr = cmk(a.n_r[2 * t], a.n_r[2 * t + 1]);
        cplx r_s, r_p, t_s, t_p;
        fresnel(nr, cos_i, r_s, r_p, t_s, t_p);
        const float thick = a.thickness != nullptr ? a.thickness[t] : -1.0f;
        if (thick >= 0.0f) {
            const cplx eta = cmul(nr, nr);
            const cplx aa = csqrt_(cmk(eta.re - (1.0f - cos_i * cos_i), eta.im));
            const float qs = __fdiv_rn(6.2831853071795864f * thick, wavelength);
            const cplx q = cscale(aa, qs);
            // exp(-2j q) = exp(2 q.im) (cos(2 q.re) - j sin(2 q.re))
            float sn, cs;
            sincosf(2.0f * q.re, &sn, &cs);
            const float mag = expf(2.0f * q.im);
            const cplx ex = cmk(mag * cs, -mag * sn);
            const cplx one = cmk(1.0f, 0.0f);
            r_s = cdiv_safe(cmul(r_s, csub(one, ex)), csub(one, cmul(cmul(r_s, r_s), ex)));
            r_p = cdiv_safe(cmul(r_p, csub(one, ex)), csub(one, cmul(cmul(r_p, r_p), ex)));
        }
        // in_rot: (theta_in, phi_in) → (e_i_s, e_i_p); d = diag(r_s, r_p); out_rot: (e_r_s, e_r_p) → (theta_out, phi_out)
        const float i11 = dot3(e_i_s, th_in), i12 = dot3(e_i_s, ph_in), i21 = dot3(e_i_p, th_in), i22 = dot3(e_i_p, ph_in);
        const cplx fs = cmul(r_s, cadd(cscale(e0, i11), cscale(e1, i12)));
        const cplx fp = cmul(r_p, cadd(cscale(e0, i21), cscale(e1, i22)));
        const float o11 = dot3(th_out, e_i_s), o12 = dot3(th_out, e_r_p), o21 = dot3(ph_out, e_i_s), o22 = dot3(ph_out, e_r_p);
        e0 = cadd(cscale(fs, o11), cscale(fp, o12));
        e1 = cadd(cscale(fs, o21), cscale(fp, o22));
        k_in = k_out, th_in = th_out, ph_in = ph_out;
    }
    // projection on the receive polarisation (plugins/deepmimo.py:641-658)
    float3 th_neg, ph_neg;
    spherical_basis(make_float3(-k_in.x, -k_in.y, -k_in.z), th_neg, ph_neg);
    const float a_coeff = dot3(th_in, th_neg);
    cplx ar = a.rx_pol == 0 ? cscale(e0, a_coeff) : cscale(e1, -a_coeff);
    // spreading and phase (:660-665), lambda / 4 pi (:694-696)
    const float spread = s_tot == 0.0f ? 0.0f : __fdiv_rn(1.0f, s_tot);
    const float phase = __fdiv_rn(a.omega * s_tot, 299792458.0f);
    float sn, cs;
    sincosf(phase, &sn, &cs);
    ar = cmul(ar, cmk(spread * cs, spread * sn));
    ar = cscale(ar, a.scale);
    if (live) {
        if (a.out_a) reinterpret_cast<float2 *>(a.out_a)[p] = make_float2(ar.re, ar.im);
        if (a.out_length) a.out_length[p] = s_tot;
    }
    if (a.pair_index != nullptr) {
        // paths of one (tx, rx) pair are neighbours in the compacted order: one atomic per pair and warp, not per path
        int64_t q = live ? a.pair_index[p] : -1;
        if (q >= a.num_pairs) q = -1;
        const int lane = threadIdx.x & 31;
        const unsigned peers = __match_any_sync(0xffffffffu, q);
        float fr = ar.re, fi = ar.im, pw = ar.re * ar.re + ar.im * ar.im;
        reduce_peers3(peers, lane, fr, fi, pw);
        if (q >= 0 && lane == __ffs(peers) - 1) {
            if (a.field) {
                atomicAdd(a.field + 2 * q, fr);
                atomicAdd(a.field + 2 * q + 1, fi);
            }
            if (a.power) atomicAdd(a.power + q, pw);
        }
    }
}

}  // namespace drt

using namespace drt;

extern "C" {

int drt_em_fresnel_coefficients(drt_stream_t stream, int64_t n, const float *n_r, int64_t n_r_stride,
                                const float *cos_theta_i, int64_t cos_stride, float *r_s, float *r_p,
                                float *t_s, float *t_p) {
    if (n < 0 || n_r_stride < 0 || cos_stride < 0) return DRT_ERR_BAD_EXTENT;
    if (n == 0) return DRT_OK;
    if (!n_r || !cos_theta_i) return DRT_ERR_NULL_POINTER;
    em_fresnel_kernel<<<unsigned((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        n, n_r, n_r_stride, cos_theta_i, cos_stride, r_s, r_p, t_s, t_p);
    return cudaGetLastError() == cudaSuccess ? DRT_OK : DRT_ERR_CUDA;
}

int drt_em_sp_directions(drt_stream_t stream, int64_t n, const float *k_i, const float *k_r, const float *normals,
                         float *e_i_s, float *e_i_p, float *e_r_p) {
    if (n < 0) return DRT_ERR_BAD_EXTENT;
    if (n == 0) return DRT_OK;
    if (!k_i || !k_r || !normals || !e_i_s || !e_i_p || !e_r_p) return DRT_ERR_NULL_POINTER;
    em_sp_directions_kernel<<<unsigned((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        n, k_i, k_r, normals, e_i_s, e_i_p, e_r_p);
    return cudaGetLastError() == cudaSuccess ? DRT_OK : DRT_ERR_CUDA;
}

int drt_em_path_coefficients(drt_stream_t stream, int64_t num_paths, int32_t order, const float *vertices,
                             const int32_t *objects, int64_t num_triangles, const void *pack,
                             const float *n_r, const float *thickness, double frequency, int32_t tx_polarization,
                             int32_t rx_polarization, float *out_a, float *out_length, const int64_t *pair_index,
                             int64_t num_pairs, float *field, float *power) {
    if (num_paths < 0 || order < 0 || num_triangles < 0 || num_pairs < 0) return DRT_ERR_BAD_EXTENT;
    if (order > DRT_MAX_ORDER) return DRT_ERR_UNSUPPORTED;
    if (!(frequency > 0.0)) return DRT_ERR_BAD_EXTENT;
    if (tx_polarization < 0 || tx_polarization > 1 || rx_polarization < 0 || rx_polarization > 1)
        return DRT_ERR_UNSUPPORTED;
    if (num_paths == 0) return DRT_OK;
    if (!vertices) return DRT_ERR_NULL_POINTER;
    if (order > 0 && (!objects || !pack || !n_r || num_triangles == 0)) return DRT_ERR_NULL_POINTER;
    if ((field || power) && !pair_index) return DRT_ERR_NULL_POINTER;
    const double kPi = 3.14159265358979323846, f = frequency, wl = 299792458.0 / f;
    EmArgs a{num_paths, vertices, objects, num_triangles, static_cast<const Tri48 *>(pack), n_r, thickness, float(wl), float(-2.0 * kPi * f), float(wl / (4.0 * kPi)),
             tx_polarization, rx_polarization, out_a, out_length, pair_index, num_pairs, field, power};
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const unsigned grid = unsigned((num_paths + 127) / 128);
#define DRT_EM_CASE(K) \
    case K:            \
        em_path_kernel<K><<<grid, 128, 0, s>>>(a); \
        break;
    switch (order) {
        DRT_EM_CASE(0) DRT_EM_CASE(1) DRT_EM_CASE(2) DRT_EM_CASE(3) DRT_EM_CASE(4) DRT_EM_CASE(5) DRT_EM_CASE(6)
        DRT_EM_CASE(7) DRT_EM_CASE(8)
    }
#undef DRT_EM_CASE
    return cudaGetLastError() == cudaSuccess ? DRT_OK : DRT_ERR_CUDA;
}

}  // extern "C"
