// XLA typed-FFI handlers over the differt_b200 C ABI — the reference-side shim a DiffeRT maintainer
// would compile next to jaxlib (`g++ -shared -fPIC -I$(python -c "import jax; print(jax.ffi.include_dir())")
// -I include integration/xla_ffi.cc -L differt_b200 -ldiffert_b200 -o libdiffert_b200_xla.so`).
//
// NOT BUILT IN THIS IMAGE: jaxlib (and therefore xla/ffi/api/ffi.h) is not installed, so this file is kept out of
// differt_b200/csrc (build.py compiles *.cu only) and has never run.  What IS checked here, on every CPU test run
// (tests/test_abi.py::test_xla_ffi_shim_type_checks): `g++ -fsyntax-only` against include/differt_b200.h and a
// declaration-only stand-in for the XLA header (integration/stub/xla/ffi/api/ffi.h) — every drt_* call matches the
// C ABI, every handler's parameter list matches the binding it is registered with.  It replaces
// the three Warp launchers the reference registers through wp.jax_callable:
//   _ray_intersect_any_triangle_anyhit_func   differt/src/differt/geometry/_mesh.py:160-181
//   _first_triangle_hit_by_ray_func           differt/src/differt/geometry/_mesh.py:202-223
//   _triangles_visible_from_vertex_func       differt/src/differt/geometry/_mesh.py:369-401
// and adds the fused trace step (_solvers.py:499-770) as a fourth target.
#include <cuda_runtime_api.h>

#include "differt_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error Check(int rc) {
    if (rc == DRT_OK) return ffi::Error::Success();
    return ffi::Error(rc == DRT_ERR_WORKSPACE ? ffi::ErrorCode::kResourceExhausted
                                              : ffi::ErrorCode::kInvalidArgument,
                      drt_error_string(rc));
}

// vertices [V,3] f32, triangles [T,3] i32, mask [T] u8 (T may be 0-sized → no mask),
// origins / directions [R,3] f32 → hit [R] pred
static ffi::Error AnyHitImpl(cudaStream_t stream, ffi::ScratchAllocator scratch,
                             ffi::Buffer<ffi::F32> vertices, ffi::Buffer<ffi::S32> triangles,
                             ffi::Buffer<ffi::U8> mask, ffi::Buffer<ffi::F32> origins,
                             ffi::Buffer<ffi::F32> directions, float epsilon, float hit_tol,
                             ffi::ResultBuffer<ffi::PRED> hit) {
    const int64_t V = vertices.dimensions()[0], T = triangles.dimensions()[0];
    const int64_t R = origins.dimensions()[0];
    auto pack = scratch.Allocate(drt_mesh_pack_bytes(T));
    if (!pack.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "pack scratch");
    const uint8_t *m = mask.element_count() ? mask.typed_data() : nullptr;
    if (int rc = drt_mesh_pack(stream, V, T, vertices.typed_data(), triangles.typed_data(), m, *pack))
        return Check(rc);
    return Check(drt_ray_intersect_any_triangle(stream, R, origins.typed_data(),
                                                directions.typed_data(), *pack, T, epsilon, hit_tol,
                                                reinterpret_cast<uint8_t *>(hit->typed_data()),
                                                nullptr));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(DrtRayIntersectAnyTriangle, AnyHitImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Attr<float>("epsilon")
                                  .Attr<float>("hit_tol")
                                  .Ret<ffi::Buffer<ffi::PRED>>());

// → face [R] i32, t [R] f32 ; (-1, +inf) on a miss
static ffi::Error FirstHitImpl(cudaStream_t stream, ffi::ScratchAllocator scratch,
                               ffi::Buffer<ffi::F32> vertices, ffi::Buffer<ffi::S32> triangles,
                               ffi::Buffer<ffi::U8> mask, ffi::Buffer<ffi::F32> origins,
                               ffi::Buffer<ffi::F32> directions, float epsilon, int64_t batch_size,
                               ffi::ResultBuffer<ffi::S32> face, ffi::ResultBuffer<ffi::F32> t) {
    const int64_t V = vertices.dimensions()[0], T = triangles.dimensions()[0];
    auto pack = scratch.Allocate(drt_mesh_pack_bytes(T));
    if (!pack.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "pack scratch");
    const uint8_t *m = mask.element_count() ? mask.typed_data() : nullptr;
    if (int rc = drt_mesh_pack(stream, V, T, vertices.typed_data(), triangles.typed_data(), m, *pack))
        return Check(rc);
    return Check(drt_first_triangle_hit_by_ray(stream, origins.dimensions()[0], origins.typed_data(),
                                               directions.typed_data(), *pack, T, epsilon, batch_size,
                                               face->typed_data(), t->typed_data(), nullptr));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(DrtFirstTriangleHitByRay, FirstHitImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Attr<float>("epsilon")
                                  .Attr<int64_t>("batch_size")
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());

// backward of the hit distance (custom_vjp bwd, _mesh.py:308-338)
static ffi::Error FirstHitVjpImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> vertices,
                                  ffi::Buffer<ffi::S32> triangles, ffi::Buffer<ffi::F32> origins,
                                  ffi::Buffer<ffi::F32> directions, ffi::Buffer<ffi::S32> faces,
                                  ffi::Buffer<ffi::F32> g_t, ffi::ResultBuffer<ffi::F32> g_vertices,
                                  ffi::ResultBuffer<ffi::F32> g_origins,
                                  ffi::ResultBuffer<ffi::F32> g_directions) {
    return Check(drt_first_triangle_hit_by_ray_vjp(
        stream, origins.dimensions()[0], vertices.dimensions()[0], triangles.dimensions()[0],
        vertices.typed_data(), triangles.typed_data(), origins.typed_data(), directions.typed_data(),
        faces.typed_data(), g_t.typed_data(), g_vertices->typed_data(), g_origins->typed_data(),
        g_directions->typed_data()));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(DrtFirstTriangleHitByRayVjp, FirstHitVjpImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());

// viewing vertices [B,3], ray directions [B,n,3] → visible [B,T] pred
static ffi::Error VisibleImpl(cudaStream_t stream, ffi::ScratchAllocator scratch,
                              ffi::Buffer<ffi::F32> vertices, ffi::Buffer<ffi::S32> triangles,
                              ffi::Buffer<ffi::U8> mask, ffi::Buffer<ffi::F32> viewers,
                              ffi::Buffer<ffi::F32> directions, float epsilon,
                              ffi::ResultBuffer<ffi::PRED> visible) {
    const int64_t V = vertices.dimensions()[0], T = triangles.dimensions()[0];
    auto pack = scratch.Allocate(drt_mesh_pack_bytes(T));
    if (!pack.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "pack scratch");
    const uint8_t *m = mask.element_count() ? mask.typed_data() : nullptr;
    if (int rc = drt_mesh_pack(stream, V, T, vertices.typed_data(), triangles.typed_data(), m, *pack))
        return Check(rc);
    return Check(drt_triangles_visible_from_vertex(
        stream, viewers.dimensions()[0], directions.dimensions()[1], viewers.typed_data(),
        directions.typed_data(), *pack, T, epsilon, reinterpret_cast<uint8_t *>(visible->typed_data()),
        nullptr));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(DrtTrianglesVisibleFromVertex, VisibleImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Attr<float>("epsilon")
                                  .Ret<ffi::Buffer<ffi::PRED>>());

// fused _trace_path_candidates: → vertices [Ntx,Nrx,C,k+2,3], objects [Ntx,Nrx,C,k+2], mask [Ntx,Nrx,C]
static ffi::Error TraceImpl(cudaStream_t stream, ffi::ScratchAllocator scratch,
                            ffi::Buffer<ffi::F32> vertices, ffi::Buffer<ffi::S32> triangles,
                            ffi::Buffer<ffi::U8> mask, ffi::Buffer<ffi::F32> tx,
                            ffi::Buffer<ffi::F32> rx, ffi::Buffer<ffi::S32> candidates,
                            bool assume_quads, float epsilon, float hit_tol, float min_len,
                            ffi::ResultBuffer<ffi::F32> out_vertices,
                            ffi::ResultBuffer<ffi::S32> out_objects,
                            ffi::ResultBuffer<ffi::PRED> out_mask) {
    const int64_t V = vertices.dimensions()[0], T = triangles.dimensions()[0];
    const int64_t ntx = tx.dimensions()[0], nrx = rx.dimensions()[0];
    const int64_t C = candidates.dimensions()[0], k = candidates.dimensions()[1];
    const size_t ws_bytes = drt_trace_workspace_bytes(T, ntx, nrx, C);
    auto ws = scratch.Allocate(ws_bytes);
    if (!ws.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "trace workspace");
    const uint8_t *m = mask.element_count() ? mask.typed_data() : nullptr;
    return Check(drt_trace_path_candidates(
        stream, V, T, vertices.typed_data(), triangles.typed_data(), m, assume_quads ? 1 : 0, ntx,
        tx.typed_data(), nrx, rx.typed_data(), C, static_cast<int32_t>(k), candidates.typed_data(),
        epsilon, hit_tol, min_len, /*flags=*/0u, *ws, ws_bytes, out_vertices->typed_data(),
        out_objects->typed_data(), reinterpret_cast<uint8_t *>(out_mask->typed_data()), nullptr));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(DrtTracePathCandidates, TraceImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Attr<bool>("assume_quads")
                                  .Attr<float>("epsilon")
                                  .Attr<float>("hit_tol")
                                  .Attr<float>("min_len")
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::PRED>>());

// compact trace for exhaustive searches: static capacity (an attribute), → count [1] s64,
// index [capacity] s64, vertices [capacity,k+2,3], objects [capacity,k+2], valid [capacity] pred.
// The caller compares count with capacity and re-traces with a larger one on overflow.
static ffi::Error TraceValidImpl(cudaStream_t stream, ffi::ScratchAllocator scratch,
                                 ffi::Buffer<ffi::F32> vertices, ffi::Buffer<ffi::S32> triangles,
                                 ffi::Buffer<ffi::U8> mask, ffi::Buffer<ffi::F32> tx,
                                 ffi::Buffer<ffi::F32> rx, ffi::Buffer<ffi::S32> candidates,
                                 bool assume_quads, float epsilon, float hit_tol, float min_len,
                                 int64_t capacity, ffi::ResultBuffer<ffi::S64> out_count,
                                 ffi::ResultBuffer<ffi::S64> out_index,
                                 ffi::ResultBuffer<ffi::F32> out_vertices,
                                 ffi::ResultBuffer<ffi::S32> out_objects,
                                 ffi::ResultBuffer<ffi::PRED> out_valid) {
    const int64_t V = vertices.dimensions()[0], T = triangles.dimensions()[0];
    const size_t ws_bytes = drt_trace_valid_workspace_bytes(T, capacity);
    auto ws = scratch.Allocate(ws_bytes);
    if (!ws.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "trace workspace");
    const uint8_t *m = mask.element_count() ? mask.typed_data() : nullptr;
    return Check(drt_trace_valid_path_candidates(
        stream, V, T, vertices.typed_data(), triangles.typed_data(), m, assume_quads ? 1 : 0,
        tx.dimensions()[0], tx.typed_data(), rx.dimensions()[0], rx.typed_data(), candidates.dimensions()[0],
        static_cast<int32_t>(candidates.dimensions()[1]), candidates.typed_data(), epsilon, hit_tol, min_len,
        /*flags=*/0u, capacity, *ws, ws_bytes, out_count->typed_data(), out_index->typed_data(), out_vertices->typed_data(),
        out_objects->typed_data(), reinterpret_cast<uint8_t *>(out_valid->typed_data())));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(DrtTraceValidPathCandidates, TraceValidImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Attr<bool>("assume_quads")
                                  .Attr<float>("epsilon")
                                  .Attr<float>("hit_tol")
                                  .Attr<float>("min_len")
                                  .Attr<int64_t>("capacity")
                                  .Ret<ffi::Buffer<ffi::S64>>()
                                  .Ret<ffi::Buffer<ffi::S64>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::PRED>>());

// candidates start .. start+count-1 of the complete graph, decoded on the device: replaces the host
// DFS + host→device copy of ExhaustivePathTracer.generate_path_candidates (_solvers.py:803-848)
static ffi::Error CompleteGraphCandidatesImpl(cudaStream_t stream, int64_t num_nodes, int64_t start,
                                              int64_t stride_multiplier,
                                              ffi::ResultBuffer<ffi::S32> out) {
    return Check(drt_complete_graph_candidates(stream, num_nodes, static_cast<int32_t>(out->dimensions()[1]),
                                               start, out->dimensions()[0],
                                               static_cast<int32_t>(stride_multiplier), out->typed_data()));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(DrtCompleteGraphCandidates, CompleteGraphCandidatesImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("num_nodes")
                                  .Attr<int64_t>("start")
                                  .Attr<int64_t>("stride_multiplier")
                                  .Ret<ffi::Buffer<ffi::S32>>());

// relaxed (smoothing_factor) trace, forward: float confidences (_solvers.py:599-713)
static ffi::Error TraceSmoothImpl(cudaStream_t stream, ffi::ScratchAllocator scratch,
                                  ffi::Buffer<ffi::F32> vertices, ffi::Buffer<ffi::S32> triangles,
                                  ffi::Buffer<ffi::U8> mask, ffi::Buffer<ffi::F32> tx, ffi::Buffer<ffi::F32> rx,
                                  ffi::Buffer<ffi::S32> candidates, bool assume_quads, float epsilon,
                                  float hit_tol, float min_len, float smoothing_factor,
                                  ffi::ResultBuffer<ffi::F32> out_vertices, ffi::ResultBuffer<ffi::S32> out_objects,
                                  ffi::ResultBuffer<ffi::F32> out_mask) {
    const int64_t V = vertices.dimensions()[0], T = triangles.dimensions()[0];
    const int64_t ntx = tx.dimensions()[0], nrx = rx.dimensions()[0], C = candidates.dimensions()[0];
    const size_t ws_bytes = drt_trace_smooth_workspace_bytes(T, ntx, nrx, C);
    auto ws = scratch.Allocate(ws_bytes);
    if (!ws.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "trace workspace");
    const uint8_t *m = mask.element_count() ? mask.typed_data() : nullptr;
    return Check(drt_trace_path_candidates_smooth(
        stream, V, T, vertices.typed_data(), triangles.typed_data(), m, assume_quads ? 1 : 0, ntx, tx.typed_data(),
        nrx, rx.typed_data(), C, static_cast<int32_t>(candidates.dimensions()[1]), candidates.typed_data(), epsilon,
        hit_tol, min_len, smoothing_factor, *ws, ws_bytes, out_vertices->typed_data(), out_objects->typed_data(),
        out_mask->typed_data()));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(DrtTracePathCandidatesSmooth, TraceSmoothImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Attr<bool>("assume_quads")
                                  .Attr<float>("epsilon")
                                  .Attr<float>("hit_tol")
                                  .Attr<float>("min_len")
                                  .Attr<float>("smoothing_factor")
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());

// ... and its reverse mode: residuals (out_vertices, out_mask), cotangents of both → vertices, tx, rx
static ffi::Error TraceSmoothVjpImpl(cudaStream_t stream, ffi::ScratchAllocator scratch,
                                     ffi::Buffer<ffi::F32> vertices, ffi::Buffer<ffi::S32> triangles,
                                     ffi::Buffer<ffi::U8> mask, ffi::Buffer<ffi::F32> tx, ffi::Buffer<ffi::F32> rx,
                                     ffi::Buffer<ffi::S32> candidates, ffi::Buffer<ffi::F32> out_vertices,
                                     ffi::Buffer<ffi::F32> out_mask, ffi::Buffer<ffi::F32> g_out_vertices,
                                     ffi::Buffer<ffi::F32> g_out_mask, bool assume_quads, float epsilon,
                                     float hit_tol, float min_len, float smoothing_factor,
                                     ffi::ResultBuffer<ffi::F32> g_vertices, ffi::ResultBuffer<ffi::F32> g_tx,
                                     ffi::ResultBuffer<ffi::F32> g_rx) {
    const int64_t V = vertices.dimensions()[0], T = triangles.dimensions()[0];
    const int64_t ntx = tx.dimensions()[0], nrx = rx.dimensions()[0], C = candidates.dimensions()[0];
    const int32_t k = static_cast<int32_t>(candidates.dimensions()[1]);
    const size_t ws_bytes = drt_trace_smooth_vjp_workspace_bytes(V, T, ntx, nrx, C, k);
    auto ws = scratch.Allocate(ws_bytes);
    if (!ws.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "trace VJP workspace");
    const uint8_t *m = mask.element_count() ? mask.typed_data() : nullptr;
    return Check(drt_trace_path_candidates_smooth_vjp(
        stream, V, T, vertices.typed_data(), triangles.typed_data(), m, assume_quads ? 1 : 0, ntx, tx.typed_data(),
        nrx, rx.typed_data(), C, k, candidates.typed_data(), epsilon, hit_tol, min_len, smoothing_factor,
        out_vertices.typed_data(), out_mask.typed_data(), g_out_vertices.typed_data(), g_out_mask.typed_data(), *ws,
        ws_bytes, g_tx->typed_data(), g_rx->typed_data(), g_vertices->typed_data()));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(DrtTracePathCandidatesSmoothVjp, TraceSmoothVjpImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Attr<bool>("assume_quads")
                                  .Attr<float>("epsilon")
                                  .Attr<float>("hit_tol")
                                  .Attr<float>("min_len")
                                  .Attr<float>("smoothing_factor")
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());

// Fresnel coefficients (differt/src/differt/em/_fresnel.py:46-213): n_r [n] c64, cos_theta_i [n] f32 (the caller
// broadcasts) → r_s, r_p, t_s, t_p [n] c64
static ffi::Error FresnelImpl(cudaStream_t stream, ffi::Buffer<ffi::C64> n_r, ffi::Buffer<ffi::F32> cos_theta_i,
                              ffi::ResultBuffer<ffi::C64> r_s, ffi::ResultBuffer<ffi::C64> r_p,
                              ffi::ResultBuffer<ffi::C64> t_s, ffi::ResultBuffer<ffi::C64> t_p) {
    const int64_t n = static_cast<int64_t>(r_s->element_count());
    return Check(drt_em_fresnel_coefficients(
        stream, n, reinterpret_cast<const float *>(n_r.typed_data()), n_r.element_count() == 1 ? 0 : 1,
        cos_theta_i.typed_data(), cos_theta_i.element_count() == 1 ? 0 : 1, reinterpret_cast<float *>(r_s->typed_data()),
        reinterpret_cast<float *>(r_p->typed_data()), reinterpret_cast<float *>(t_s->typed_data()),
        reinterpret_cast<float *>(t_p->typed_data())));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(DrtEmFresnelCoefficients, FresnelImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::C64>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::C64>>()
                                  .Ret<ffi::Buffer<ffi::C64>>()
                                  .Ret<ffi::Buffer<ffi::C64>>()
                                  .Ret<ffi::Buffer<ffi::C64>>());

// per-path field chain of deepmimo.export (differt/src/differt/plugins/deepmimo.py:516-665, 694-696) on compacted
// paths: vertices [n,k+2,3], objects [n,k+2], mesh vertices / triangles (normals come from the pack), n_r [T] c64,
// thickness [T] f32 (negative: half space), pair_index [n] s64 → a [n] c64, length [n] f32, field [pairs] c64,
// power [pairs] f32 (the two accumulators are zero-filled here: XLA hands over uninitialised result buffers)
static ffi::Error PathCoefficientsImpl(cudaStream_t stream, ffi::ScratchAllocator scratch,
                                       ffi::Buffer<ffi::F32> mesh_vertices, ffi::Buffer<ffi::S32> triangles,
                                       ffi::Buffer<ffi::F32> vertices, ffi::Buffer<ffi::S32> objects,
                                       ffi::Buffer<ffi::C64> n_r, ffi::Buffer<ffi::F32> thickness,
                                       ffi::Buffer<ffi::S64> pair_index, double frequency, int64_t tx_polarization,
                                       int64_t rx_polarization, ffi::ResultBuffer<ffi::C64> a,
                                       ffi::ResultBuffer<ffi::F32> length, ffi::ResultBuffer<ffi::C64> field,
                                       ffi::ResultBuffer<ffi::F32> power) {
    const int64_t V = mesh_vertices.dimensions()[0], T = triangles.dimensions()[0];
    const int64_t n = vertices.dimensions()[0];
    const int32_t order = static_cast<int32_t>(vertices.dimensions()[1]) - 2;
    auto pack = scratch.Allocate(drt_mesh_pack_bytes(T));
    if (!pack.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "pack scratch");
    if (int rc = drt_mesh_pack(stream, V, T, mesh_vertices.typed_data(), triangles.typed_data(), nullptr, *pack))
        return Check(rc);
    const int64_t pairs = static_cast<int64_t>(power->element_count());
    if (cudaMemsetAsync(field->typed_data(), 0, field->size_bytes(), stream) != cudaSuccess ||
        cudaMemsetAsync(power->typed_data(), 0, power->size_bytes(), stream) != cudaSuccess)
        return ffi::Error(ffi::ErrorCode::kInternal, "cudaMemsetAsync");
    return Check(drt_em_path_coefficients(
        stream, n, order, vertices.typed_data(), objects.typed_data(), T, *pack,
        reinterpret_cast<const float *>(n_r.typed_data()), thickness.element_count() ? thickness.typed_data() : nullptr,
        frequency, static_cast<int32_t>(tx_polarization), static_cast<int32_t>(rx_polarization),
        reinterpret_cast<float *>(a->typed_data()), length->typed_data(), pair_index.typed_data(), pairs,
        reinterpret_cast<float *>(field->typed_data()), power->typed_data()));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(DrtEmPathCoefficients, PathCoefficientsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::C64>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S64>>()
                                  .Attr<double>("frequency")
                                  .Attr<int64_t>("tx_polarization")
                                  .Attr<int64_t>("rx_polarization")
                                  .Ret<ffi::Buffer<ffi::C64>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::C64>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());
