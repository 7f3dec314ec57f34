/* differt_b200 — C ABI of the B200-native DiffeRT geometric hot path.
 *
 * This is the drop-in boundary: flat device buffers + a CUDA stream, exactly the shape of the
 * launchers the reference registers with XLA through `wp.jax_callable`
 * (reference: differt/src/differt/geometry/_mesh.py:160-181, 202-223, 369-401, 3082-3092) and of the
 * pure-JAX functions those launchers accelerate (differt/src/differt/geometry/_utils.py:1157-1960,
 * _solver_image_method.py:11-454, _solvers.py:499-770).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - all floats are IEEE binary32, all indices int32, all masks/booleans one byte (0/1);
 *   - calls are asynchronous on `stream` (a `cudaStream_t`), never synchronise, never allocate,
 *     are re-entrant and keep no global mutable state;
 *   - return value: DRT_OK (0) or a negative DRT_ERR_* code (`drt_error_string`); nothing throws;
 *   - degenerate extents (0 rays, 0 triangles, 0 candidates) are legal and produce the constant
 *     outputs the reference produces (`False`, `(-1, +inf)`, empty);
 *   - arithmetic follows the reference's pure-JAX operation order without fused multiply-add, so
 *     boolean outputs are bit-identical to the CPU algorithm ("Parity contract").
 *
 * Semantics delta (documented in DESIGN.md): the entry points implement the *pure-JAX* definitions
 * (`t > epsilon`, `t < 1 - hit_tol` on the un-normalised ray).  The reference's Warp launchers
 * instead shorten the normalised ray by `hit_tol` at both ends (_mesh.py:3065-3070) and offset the
 * first-hit origin by 1e-5 (_mesh.py:195-199); the reference's own tests pin both forms to the same
 * results (differt/tests/geometry/test_mesh.py:1984-2090).
 */
#ifndef DIFFERT_B200_H
#define DIFFERT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRT_ABI_VERSION 2

#define DRT_OK 0
#define DRT_ERR_NULL_POINTER (-1)     /* a required pointer is NULL */
#define DRT_ERR_BAD_EXTENT (-2)       /* a negative extent, or one beyond the int32 index range */
#define DRT_ERR_UNSUPPORTED (-3)      /* order > DRT_MAX_ORDER, ndim > DRT_MAX_BATCH_DIMS ... */
#define DRT_ERR_WORKSPACE (-4)        /* workspace too small (see *_workspace_bytes) */
#define DRT_ERR_CUDA (-5)             /* a CUDA runtime call failed (sticky error on the context) */

#define DRT_MAX_ORDER 8               /* mirrors per path candidate the fused kernels are built for */
#define DRT_MAX_BATCH_DIMS 4          /* broadcast batch dims of the element-wise entry points */

/* trace flags */
#define DRT_TRACE_DENSE_BLOCKAGE 1u   /* test every candidate's segments against the mesh, like the
                                         reference does; default (0) skips candidates that already
                                         failed a cheaper test — same outputs, less work */

#define DRT_TRACE_PREPARED 4u         /* the first drt_trace_prepared_bytes(T) bytes of `workspace` already hold
                                         what drt_trace_prepare wrote for THIS mesh and mask (the caller
                                         copied them there): skip the mesh-only work of the call */
#define DRT_TRACE_PROFILE 2u          /* record a CUDA-event pair around the blockage (all-pairs)
                                         kernel of this call into the calling thread's profile
                                         ring (drt_profile_*); measurement hook for bench.py */

typedef void *drt_stream_t;           /* cudaStream_t */

#if defined(__GNUC__)
#define DRT_API __attribute__((visibility("default")))
#else
#define DRT_API
#endif

DRT_API int drt_abi_version(void);
DRT_API const char *drt_error_string(int code);

/* ---------------------------------------------------------------------------------------------
 * Packed mesh.  48 bytes per triangle: v0, e1 = v1-v0, e2 = v2-v0, unit normal; padded to whole
 * tiles of DRT_TILE_TRIANGLES with never-hit triangles.  Inactive triangles (mask[j] == 0) are
 * stored as never-hit triangles when `mask` is given.  Replaces the per-call `wp.Mesh(...)` BVH
 * build + `_WARP_MESHES_CACHE` (_mesh.py:55, 170-174): packing is stateless and takes microseconds.
 * ------------------------------------------------------------------------------------------- */
#define DRT_TILE_TRIANGLES 512
DRT_API size_t drt_mesh_pack_bytes(int64_t num_triangles);

/* from Mesh.vertices [V,3] + Mesh.triangles [T,3] (reference layout: _mesh.py:612-700) */
DRT_API int drt_mesh_pack(drt_stream_t stream, int64_t num_vertices, int64_t num_triangles,
                  const float *vertices, const int32_t *triangles, const uint8_t *mask /*nullable*/,
                  void *pack_out);
/* from triangle_vertices [T,3,3] (the form the pure-JAX functions take: _utils.py:1353-1364) */
DRT_API int drt_mesh_pack_triangle_vertices(drt_stream_t stream, int64_t num_triangles,
                                    const float *triangle_vertices,
                                    const uint8_t *mask /*nullable*/, void *pack_out);

/* Any-hit ordering of a pack: the same records in descending triangle-area order (never-hit
 * records last).  An any-hit query is an OR over triangles, so the order cannot change its result —
 * only how early a blocked ray stops: the all-pairs engine keeps the first tiles of the pack resident
 * in shared memory and tests every new ray against them first.  Triangle indices are lost: use the
 * result with drt_ray_intersect_any_triangle only.  pack_out must not alias pack_in. */
DRT_API size_t drt_mesh_pack_sort_workspace_bytes(int64_t num_triangles);
DRT_API int drt_mesh_pack_sort_by_area(drt_stream_t stream, int64_t num_triangles, const void *pack_in,
                               void *workspace, size_t workspace_bytes, void *pack_out);
/* Same, by caller-supplied keys (device uint32, one per record of the padded pack =
 * drt_mesh_pack_bytes / 48 entries), descending and STABLE: records with equal keys keep their
 * order.  The fused trace uses it with per-triangle hit counts of a sample of its own rays. */
DRT_API int drt_mesh_pack_sort_by_keys(drt_stream_t stream, int64_t num_triangles, const void *pack_in,
                               const uint32_t *keys, void *workspace, size_t workspace_bytes,
                               void *pack_out);

/* ---------------------------------------------------------------------------------------------
 * K1  ray_intersect_triangle — element-wise Möller–Trumbore over a broadcast batch
 *     (reference: _utils.py:1157-1322).  Operand element (i0..i3) lives at
 *     base + sum_d i_d * stride_d (strides in floats; 0 broadcasts).  Outputs are contiguous.
 * ------------------------------------------------------------------------------------------- */
DRT_API int drt_ray_intersect_triangle(drt_stream_t stream, int32_t ndim, const int64_t *shape_host,
                               const float *ray_origins, const int64_t *o_strides_host,
                               const float *ray_directions, const int64_t *d_strides_host,
                               const float *triangle_vertices, const int64_t *tri_strides_host,
                               float epsilon, float *t_out, uint8_t *hit_out);

/* ---------------------------------------------------------------------------------------------
 * K2  ray_intersect_any_triangle (reference: _utils.py:1353-1537; launcher _mesh.py:160-181).
 *     out[r] = any_j [ hit(r,j; epsilon) and t(r,j) < 1 - hit_tol ].
 *     `tests_done` (nullable, device int64) is incremented by the ray–triangle tests evaluated.
 * ------------------------------------------------------------------------------------------- */
DRT_API int drt_ray_intersect_any_triangle(drt_stream_t stream, int64_t num_rays, const float *ray_origins,
                                   const float *ray_directions, const void *pack,
                                   int64_t num_triangles, float epsilon, float hit_tol,
                                   uint8_t *out, int64_t *tests_done /*nullable*/);
/* K2 with the exact conservative cull of csrc/cull.cuh in front of the same Möller–Trumbore test: a warp
 * per ray walks an 8-ary hierarchy over the Morton-ordered pack and only evaluates the triangles the
 * cull cannot PROVE to be misses — identical results, O(log T) instead of O(T) per ray.  Falls back to
 * the all-pairs engine for meshes of <= 512 triangles and for epsilon < FLT_MIN or hit_tol outside
 * [0, 1) (outside the proof).  `pack` as above (its first 32 records are tested first: pass an
 * area-sorted pack when you have one).  workspace: drt_any_hit_workspace_bytes(num_triangles). */
DRT_API size_t drt_any_hit_workspace_bytes(int64_t num_triangles);
DRT_API int drt_ray_intersect_any_triangle_culled(drt_stream_t stream, int64_t num_rays, const float *ray_origins,
                                          const float *ray_directions, const void *pack,
                                          int64_t num_triangles, float epsilon, float hit_tol,
                                          void *workspace, size_t workspace_bytes, uint8_t *out,
                                          int64_t *tests_done /*nullable*/);

/* ---------------------------------------------------------------------------------------------
 * K3  first_triangle_hit_by_ray (reference: _utils.py:1775-1960; launcher _mesh.py:202-223).
 *     miss = (-1, +inf).  `batch_size` reproduces the reference's tie rule on exactly equal
 *     distances (lowest index inside a batch, latest batch across batches); <= 0 means one batch.
 * K3b VJP of the hit distance w.r.t. vertices / origins / directions with the winning faces fixed
 *     (reference: _mesh.py:226-255, 308-338).  g_vertices [V,3] is zero-filled by the callee.
 * ------------------------------------------------------------------------------------------- */
DRT_API int drt_first_triangle_hit_by_ray(drt_stream_t stream, int64_t num_rays, const float *ray_origins,
                                  const float *ray_directions, const void *pack,
                                  int64_t num_triangles, float epsilon, int64_t batch_size,
                                  int32_t *out_index, float *out_t,
                                  int64_t *tests_done /*nullable*/);
/* K3 behind the same exact cull (one warp per ray; a node is skipped only if it provably holds no hit
 * at a distance <= the best so far, ties included): index and distance identical to the all-pairs
 * reduction.  Falls back to it for meshes of <= 512 triangles and for epsilon < FLT_MIN.  `pack` must
 * be in the mesh's own triangle order (indices are reported).  workspace: drt_any_hit_workspace_bytes. */
DRT_API int drt_first_triangle_hit_by_ray_culled(drt_stream_t stream, int64_t num_rays, const float *ray_origins,
                                         const float *ray_directions, const void *pack,
                                         int64_t num_triangles, float epsilon, int64_t batch_size,
                                         void *workspace, size_t workspace_bytes, int32_t *out_index,
                                         float *out_t, int64_t *tests_done /*nullable*/);
DRT_API int drt_first_triangle_hit_by_ray_vjp(drt_stream_t stream, int64_t num_rays, int64_t num_vertices,
                                      int64_t num_triangles, const float *vertices,
                                      const int32_t *triangles, const float *ray_origins,
                                      const float *ray_directions, const int32_t *faces,
                                      const float *g_t, float *g_vertices, float *g_origins,
                                      float *g_directions);

/* ---------------------------------------------------------------------------------------------
 * K4  triangles_visible_from_vertex with the ray directions supplied by the caller, as the
 *     reference's own launcher takes them (_mesh.py:369-401; pure JAX: _utils.py:1702-1772).
 *     vertices [B,3], ray_directions [B,num_rays,3] → out [B,T]; the callee zero-fills `out`.
 * ------------------------------------------------------------------------------------------- */
DRT_API int drt_triangles_visible_from_vertex(drt_stream_t stream, int64_t num_vertices_batch,
                                      int64_t num_rays, const float *vertices,
                                      const float *ray_directions, const void *pack,
                                      int64_t num_triangles, float epsilon, uint8_t *out,
                                      int64_t *tests_done /*nullable*/);

/* ---------------------------------------------------------------------------------------------
 * K5  image_method over a broadcast batch (reference: _solver_image_method.py:206-363).
 *     from/to: vec3 per batch element; mirrors: [k,3] per batch element (inner layout contiguous);
 *     out [N,k,3] contiguous.
 * K5b reverse mode; cotangent g_paths [N,k,3] contiguous; gradients are written per batch element
 *     ([N,3],[N,3],[N,k,3],[N,k,3], contiguous) — the caller reduces over broadcast axes.
 * ------------------------------------------------------------------------------------------- */
DRT_API int drt_image_method(drt_stream_t stream, int32_t ndim, const int64_t *shape_host, int32_t order,
                     const float *from_vertices, const int64_t *from_strides_host,
                     const float *to_vertices, const int64_t *to_strides_host,
                     const float *mirror_vertices, const int64_t *mv_strides_host,
                     const float *mirror_normals, const int64_t *mn_strides_host, float *out_paths);
DRT_API int drt_image_method_vjp(drt_stream_t stream, int32_t ndim, const int64_t *shape_host,
                         int32_t order, const float *from_vertices,
                         const int64_t *from_strides_host, const float *to_vertices,
                         const int64_t *to_strides_host, const float *mirror_vertices,
                         const int64_t *mv_strides_host, const float *mirror_normals,
                         const int64_t *mn_strides_host, const float *g_paths, float *g_from,
                         float *g_to, float *g_mirror_vertices, float *g_mirror_normals);

/* a5 / a6 / a8 as stand-alone element-wise entry points over a broadcast batch
 * (reference: _solver_image_method.py:11-79, 82-135, 386-454).  Outputs contiguous. */
DRT_API int drt_image_of_vertex_with_respect_to_mirror(
    drt_stream_t stream, int32_t ndim, const int64_t *shape_host, const float *vertex,
    const int64_t *vertex_strides_host, const float *mirror_vertex, const int64_t *mv_strides_host,
    const float *mirror_normal, const int64_t *mn_strides_host, float *out);
DRT_API int drt_intersection_of_ray_with_plane(
    drt_stream_t stream, int32_t ndim, const int64_t *shape_host, const float *ray_origin,
    const int64_t *o_strides_host, const float *ray_direction, const int64_t *d_strides_host,
    const float *plane_vertex, const int64_t *pv_strides_host, const float *plane_normal,
    const int64_t *pn_strides_host, float *out);
/* vertices: [k+2,3] per batch element, mirrors: [k,3] per batch element → out [N,k] u8 */
DRT_API int drt_consecutive_vertices_are_on_same_side_of_mirror(
    drt_stream_t stream, int32_t ndim, const int64_t *shape_host, int32_t order,
    const float *vertices, const int64_t *v_strides_host, const float *mirror_vertices,
    const int64_t *mv_strides_host, const float *mirror_normals, const int64_t *mn_strides_host,
    uint8_t *out);

/* ---------------------------------------------------------------------------------------------
 * K6  fused trace + validate (reference: _solvers.py:499-770, non-smoothing branch).
 *     Inputs in the reference's own layouts: Mesh.vertices [V,3], Mesh.triangles [T,3],
 *     Mesh.mask [T] (nullable), tx [Ntx,3], rx [Nrx,3], path_candidates [C,k] int32.
 *     Outputs = the fields of TracedPaths (_paths.py:77-116), dense and contiguous:
 *       vertices [Ntx,Nrx,C,k+2,3] f32, objects [Ntx,Nrx,C,k+2] i32, mask [Ntx,Nrx,C] u8.
 *     `stats` (nullable, device int64[4]): [0] ray–triangle tests evaluated, [1] candidates that
 *     passed the cheap tests (the only ones blockage-tested unless DRT_TRACE_DENSE_BLOCKAGE),
 *     [2] candidates still unblocked after the first resident pass (8 tiles), [3] bit field: low byte =
 *     1 + greedy rounds of the batch-specific ordering pass (0 if it did not run), bit 8 = the exactly
 *     culled blockage pass (csrc/cull.cuh) ran.
 *     The callee zero-fills it.
 * K6b reverse mode of `vertices` w.r.t. tx, rx and Mesh.vertices (mask carries no cotangent,
 *     reference: _mesh.py:3087-3094).  g_* outputs are zero-filled by the callee.
 * ------------------------------------------------------------------------------------------- */
DRT_API size_t drt_trace_workspace_bytes(int64_t num_triangles, int64_t num_tx, int64_t num_rx,
                                 int64_t num_candidates);
/* The mesh-only part of K6 / K6c (packed mesh, area order, the culled hierarchy of csrc/cull.cuh),
 * hoisted for callers that trace many batches against one mesh (chunked searches, per-rank shards,
 * training steps between two mesh updates): drt_trace_prepare writes it into caller-owned memory;
 * a later call copies those bytes to the start of its workspace and passes DRT_TRACE_PREPARED.
 * The library still keeps no state — the caller owns the bytes and knows when its mesh changed. */
DRT_API size_t drt_trace_prepared_bytes(int64_t num_triangles);
DRT_API int drt_trace_prepare(drt_stream_t stream, int64_t num_vertices, int64_t num_triangles,
                      const float *vertices, const int32_t *triangles,
                      const uint8_t *triangle_mask /*nullable*/, void *prepared, size_t prepared_bytes);
DRT_API int drt_trace_path_candidates(drt_stream_t stream, int64_t num_vertices, int64_t num_triangles,
                              const float *vertices, const int32_t *triangles,
                              const uint8_t *triangle_mask /*nullable*/, int32_t assume_quads,
                              int64_t num_tx, const float *tx, int64_t num_rx, const float *rx,
                              int64_t num_candidates, int32_t order, const int32_t *path_candidates,
                              float epsilon, float hit_tol, float min_len, uint32_t flags,
                              void *workspace, size_t workspace_bytes, float *out_vertices,
                              int32_t *out_objects, uint8_t *out_mask,
                              int64_t *stats /*nullable*/);
/* K6c compact form of K6 for exhaustive searches: no dense outputs.  Every candidate is traced and
 *     validated like in drt_trace_path_candidates; the ones that pass the cheap tests are appended (in
 *     no particular order) to out_index (flat (tx, rx, candidate) index) / out_vertices / out_objects
 *     [capacity], then blockage-tested: out_valid[slot] = 1 iff the path is valid.  *out_count
 *     (device) = candidates that passed the cheap tests; if it exceeds `capacity` the excess was
 *     dropped and the caller retries with a larger capacity.  Orders 0..5.
 *     workspace: drt_trace_valid_workspace_bytes(num_triangles, capacity). */
DRT_API size_t drt_trace_valid_workspace_bytes(int64_t num_triangles, int64_t capacity);
DRT_API int drt_trace_valid_path_candidates(drt_stream_t stream, int64_t num_vertices, int64_t num_triangles,
                                    const float *vertices, const int32_t *triangles,
                                    const uint8_t *triangle_mask /*nullable*/, int32_t assume_quads,
                                    int64_t num_tx, const float *tx, int64_t num_rx, const float *rx,
                                    int64_t num_candidates, int32_t order,
                                    const int32_t *path_candidates, float epsilon, float hit_tol,
                                    float min_len, uint32_t flags, int64_t capacity, void *workspace,
                                    size_t workspace_bytes, int64_t *out_count, int64_t *out_index,
                                    float *out_vertices, int32_t *out_objects, uint8_t *out_valid);
DRT_API int drt_trace_path_candidates_vjp(drt_stream_t stream, int64_t num_vertices, int64_t num_triangles,
                                  const float *vertices, const int32_t *triangles,
                                  int64_t num_tx, const float *tx, int64_t num_rx, const float *rx,
                                  int64_t num_candidates, int32_t order,
                                  const int32_t *path_candidates, const float *g_out_vertices,
                                  float *g_tx, float *g_rx, float *g_vertices);

/* ---------------------------------------------------------------------------------------------
 * First EM consumer of the traced paths.
 * drt_em_fresnel_coefficients: reference em/_fresnel.py:46-213.  Complex values are interleaved
 *     (re, im) float pairs; `n_r` / `cos_theta_i` are read at index i * stride (stride 0 broadcasts a
 *     scalar); any of the four outputs may be null.
 * drt_em_path_coefficients: one complex coefficient per path — the field chain of the reference's
 *     consumers (em/_utils.py:243-302 `sp_directions` / `sp_rotation_matrix`, the composition in
 *     plugins/deepmimo.py:348-405, 516-665: per interaction J = R_out diag(r_s, r_p) R_in with the slab
 *     formula when thickness >= 0, projection on the receive polarisation, 1 / length, phase
 *     exp(-j 2 pi f length / c), lambda / 4 pi) — for paths given as compacted TracedPaths fields
 *     (vertices [n, order+2, 3], objects [n, order+2]); normals come from the mesh pack (drt_mesh_pack),
 *     `n_r` [T] complex and `thickness` [T] (nullable: half spaces) are per triangle; polarisations
 *     0 = V, 1 = H.  Optionally fused with the accumulation per (tx, rx) pair: field[pair_index[i]] +=
 *     a_i (coherent), power[pair_index[i]] += |a_i|^2; the caller zero-fills `field` / `power`.
 * ------------------------------------------------------------------------------------------- */
DRT_API int drt_em_fresnel_coefficients(drt_stream_t stream, int64_t n, const float *n_r, int64_t n_r_stride,
                                const float *cos_theta_i, int64_t cos_theta_stride, float *r_s /*nullable*/,
                                float *r_p /*nullable*/, float *t_s /*nullable*/, float *t_p /*nullable*/);
/* em/_utils.py:84-262 `sp_directions` on flat, contiguous [n, 3] operands; e_r_s equals e_i_s. */
DRT_API int drt_em_sp_directions(drt_stream_t stream, int64_t n, const float *k_i, const float *k_r,
                         const float *normals, float *e_i_s, float *e_i_p, float *e_r_p);
DRT_API int drt_em_path_coefficients(drt_stream_t stream, int64_t num_paths, int32_t order, const float *vertices,
                             const int32_t *objects, int64_t num_triangles, const void *pack,
                             const float *n_r, const float *thickness /*nullable*/, double frequency,
                             int32_t tx_polarization, int32_t rx_polarization, float *out_a /*nullable*/,
                             float *out_length /*nullable*/, const int64_t *pair_index /*nullable*/,
                             int64_t num_pairs, float *field /*nullable*/, float *power /*nullable*/);

/* Profile ring (per host thread, 64 slots).  A call made with DRT_TRACE_PROFILE records one event
 * pair on its stream, immediately before and after the blockage kernel, into the next free slot.
 * drt_profile_elapsed_ms waits for that slot's stop event and returns the device time between the
 * two events.  Nothing else in the library keeps state. */
#define DRT_PROFILE_SLOTS 64
DRT_API int drt_profile_reset(void);
DRT_API int drt_profile_count(void);
DRT_API int drt_profile_elapsed_ms(int32_t slot, float *ms_host);

/* ---------------------------------------------------------------------------------------------
 * TracedPaths.masked() (reference: _paths.py:299-328): stable compaction of the valid paths in
 * row-major (tx, rx, candidate) order.  out_count: device int64 = number of valid paths (may
 * exceed `capacity`; only the first `capacity` are stored).  out_index: flat path index of each
 * survivor.  workspace: drt_compact_workspace_bytes(num_paths).
 * ------------------------------------------------------------------------------------------- */
DRT_API size_t drt_compact_workspace_bytes(int64_t num_paths);
DRT_API int drt_compact_valid_paths(drt_stream_t stream, int64_t num_paths, int32_t order,
                            const float *vertices, const int32_t *objects, const uint8_t *mask,
                            int64_t capacity, void *workspace, size_t workspace_bytes,
                            int64_t *out_count, int64_t *out_index, float *out_vertices,
                            int32_t *out_objects);

/* ---------------------------------------------------------------------------------------------
 * N1  path candidates of the complete graph, decoded on the device from the linear index
 *     (reference: differt-core/src/geometry/graph.rs:286-491, count formula :356-362).
 *     out [count, order] int32 = candidates start .. start+count-1 in the reference's order.
 * ------------------------------------------------------------------------------------------- */
DRT_API int drt_complete_graph_candidates(drt_stream_t stream, int64_t num_nodes, int32_t order,
                                  int64_t start, int64_t count, int32_t stride_multiplier,
                                  int32_t *out);

/* ---------------------------------------------------------------------------------------------
 * N1b path candidates of the visibility-pruned graph of HybridPathTracer
 *     (reference: differt/src/differt/geometry/_solvers.py:993-1058;
 *      differt-core/src/geometry/graph.rs:636-691 insert_from_and_to_nodes, :879-915 filter_by_mask,
 *      :1063-1108 DFS order): every tuple with c_1 in from_mask, c_k in to_mask, all c_i in
 *     active_mask, c_i != c_{i-1}, in lexicographic order.  NULL masks mean "every node".
 *     `prepare` fills the workspace with the per-position completion counts and writes the total
 *     number of candidates to total_out (device int64); `drt_digraph_candidates` then decodes
 *     candidates start .. start+count-1 from the prepared workspace.  The caller must keep
 *     num_nodes^order below 2^63.
 * ------------------------------------------------------------------------------------------- */
DRT_API size_t drt_digraph_candidates_workspace_bytes(int64_t num_nodes, int32_t order);
DRT_API int drt_digraph_candidates_prepare(drt_stream_t stream, int64_t num_nodes, int32_t order,
                                   const uint8_t *from_mask /*nullable*/,
                                   const uint8_t *to_mask /*nullable*/,
                                   const uint8_t *active_mask /*nullable*/, void *workspace,
                                   size_t workspace_bytes, int64_t *total_out);
DRT_API int drt_digraph_candidates(drt_stream_t stream, int64_t num_nodes, int32_t order,
                           const void *workspace, int64_t start, int64_t count,
                           int32_t stride_multiplier, int32_t *out);

/* ---------------------------------------------------------------------------------------------
 * N4  smoothed variants of the primitives and of the trace, forward and reverse mode: comparisons → sigmoid(x * smoothing_factor)
 *     (reference: differt/src/differt/utils.py:70-89), AND → min, OR over triangles → sum clipped at 1
 *     (_utils.py:1279-1318, 1465-1476; _solver_image_method.py:448-454).  Float outputs in [0,1];
 *     parity to 1e-5 (transcendental), not bit-exact.  Masked-out triangles of a pack contribute 0.
 * ------------------------------------------------------------------------------------------- */
DRT_API int drt_ray_intersect_triangle_smooth(drt_stream_t stream, int32_t ndim, const int64_t *shape_host,
                                      const float *ray_origins, const int64_t *o_strides_host,
                                      const float *ray_directions, const int64_t *d_strides_host,
                                      const float *triangle_vertices, const int64_t *tri_strides_host,
                                      float epsilon, float smoothing_factor, float *t_out,
                                      float *hit_out);
DRT_API int drt_ray_intersect_any_triangle_smooth(drt_stream_t stream, int64_t num_rays,
                                          const float *ray_origins, const float *ray_directions,
                                          const void *pack, int64_t num_triangles, float epsilon,
                                          float hit_tol, float smoothing_factor, float *out);
DRT_API int drt_consecutive_vertices_are_on_same_side_of_mirror_smooth(
    drt_stream_t stream, int32_t ndim, const int64_t *shape_host, int32_t order,
    const float *vertices, const int64_t *v_strides_host, const float *mirror_vertices,
    const int64_t *mv_strides_host, const float *mirror_normals, const int64_t *mn_strides_host,
    float smoothing_factor, float *out);

/* Reverse mode of the two relaxed primitives called on their own (flat, already broadcast inputs).
 * Elementwise: cotangents g_t / g_hit [n] (each nullable) → g_o, g_d [n,3], g_triangle_vertices
 * [n,3,3] (overwritten; the caller sums over its broadcast axes).  Any-hit: cotangent g_out [R] →
 * g_o, g_d [R,3] (overwritten) and g_triangle_vertices [T,3,3] (zeroed, then accumulated with float
 * atomics); rays whose sum reached the clip at 1 pass nothing.  The same-side relaxation depends on
 * signs only: its gradient is identically zero and has no entry point. */
DRT_API int drt_ray_intersect_triangle_smooth_vjp(drt_stream_t stream, int64_t n, const float *ray_origins,
                                          const float *ray_directions, const float *triangle_vertices,
                                          float epsilon, float smoothing_factor, const float *g_t,
                                          const float *g_hit, float *g_origins, float *g_directions,
                                          float *g_triangle_vertices);
DRT_API int drt_ray_intersect_any_triangle_smooth_vjp(drt_stream_t stream, int64_t num_rays,
                                              const float *ray_origins, const float *ray_directions,
                                              const void *pack, int64_t num_triangles, float epsilon,
                                              float hit_tol, float smoothing_factor, const float *g_out,
                                              float *g_origins, float *g_directions,
                                              float *g_triangle_vertices);

/* The relaxed trace + validate step: _trace_path_candidates with smoothing_factor (reference:
 * _solvers.py:576-719, the `smoothing_factor is not None` branches).  Same inputs and dense
 * out_vertices / out_objects as drt_trace_path_candidates (K6; vertices bit-identical to it), but
 * out_mask is float32 [Ntx, Nrx, C]: min(inside, same_side, 1 - blocked, 1 - too_small, finite),
 * times 1/0 for candidates that use a masked-out triangle.  NaN propagates as in jnp.min.
 * Cost: every path segment against every active triangle (the relaxed any-hit is a sum: no
 * early exit).  workspace: device scratch of drt_trace_smooth_workspace_bytes(...) bytes. */
DRT_API size_t drt_trace_smooth_workspace_bytes(int64_t num_triangles, int64_t num_tx, int64_t num_rx,
                                        int64_t num_candidates);
DRT_API int drt_trace_path_candidates_smooth(
    drt_stream_t stream, int64_t num_vertices, int64_t num_triangles, const float *vertices,
    const int32_t *triangles, const uint8_t *triangle_mask, int32_t assume_quads, int64_t num_tx,
    const float *tx_vertices, int64_t num_rx, const float *rx_vertices, int64_t num_candidates,
    int32_t order, const int32_t *path_candidates, float epsilon, float hit_tol, float min_len,
    float smoothing_factor, void *workspace, size_t workspace_bytes, float *out_vertices,
    int32_t *out_objects, float *out_mask);

/* Reverse mode of the relaxed trace: what jax.grad gives on _solvers.py:576-713.  Inputs: the
 * forward's arguments, its saved out_vertices / out_mask, the cotangents g_out_vertices [P, k+2, 3]
 * (nullable) and g_out_mask [P] (nullable: then this is drt_trace_path_candidates_vjp).  Outputs
 * (overwritten): g_tx [Ntx,3], g_rx [Nrx,3], g_vertices [V,3].  The cotangent of a confidence goes to
 * the term its min came from: the relaxed inside test of one interaction (path vertices + that
 * triangle's vertices), the too-small segment, or — when the blockage sum of the worst segment is
 * below its clip — every active triangle of the mesh (float atomics: not bit-reproducible); the
 * same-side term depends on signs only.  min / max send the cotangent to their FIRST extremal
 * argument where JAX shares it between exact ties; non-finite paths get none. */
DRT_API size_t drt_trace_smooth_vjp_workspace_bytes(int64_t num_vertices, int64_t num_triangles,
                                            int64_t num_tx, int64_t num_rx, int64_t num_candidates,
                                            int32_t order);
DRT_API int drt_trace_path_candidates_smooth_vjp(
    drt_stream_t stream, int64_t num_vertices, int64_t num_triangles, const float *vertices,
    const int32_t *triangles, const uint8_t *triangle_mask, int32_t assume_quads, int64_t num_tx,
    const float *tx_vertices, int64_t num_rx, const float *rx_vertices, int64_t num_candidates,
    int32_t order, const int32_t *path_candidates, float epsilon, float hit_tol, float min_len,
    float smoothing_factor, const float *out_vertices, const float *out_mask,
    const float *g_out_vertices, const float *g_out_mask, void *workspace, size_t workspace_bytes,
    float *g_tx, float *g_rx, float *g_vertices);

/* ---------------------------------------------------------------------------------------------
 * N2  OPT-IN bounding-volume hierarchy (linear BVH over Morton-sorted triangles) for the queries the
 *     reference answers with Warp's BVH (wp.mesh_query_ray[_anyhit], _mesh.py:142-223, 347-401).
 *     Every triangle that is reached is tested with the same arithmetic as the brute-force kernels
 *     (bit-identical t, same tie rule); node boxes are padded by relative_pad x scene size.  Results
 *     equal the brute-force ones except for rays grazing a triangle's plane, where the reference's
 *     fp32 test can accept hits no bounding volume contains — hence opt-in.
 *     The BVH blob is caller-owned and stateless like the pack: build, query, discard.
 * ------------------------------------------------------------------------------------------- */
DRT_API size_t drt_bvh_bytes(int64_t num_triangles);
DRT_API size_t drt_bvh_workspace_bytes(int64_t num_triangles);
DRT_API int drt_bvh_build(drt_stream_t stream, int64_t num_triangles, const void *pack, float relative_pad,
                  void *workspace, size_t workspace_bytes, void *bvh_out);
DRT_API int drt_bvh_ray_intersect_any_triangle(drt_stream_t stream, int64_t num_rays,
                                       const float *ray_origins, const float *ray_directions,
                                       const void *bvh, int64_t num_triangles, float epsilon,
                                       float hit_tol, uint8_t *out);
DRT_API int drt_bvh_first_triangle_hit_by_ray(drt_stream_t stream, int64_t num_rays,
                                      const float *ray_origins, const float *ray_directions,
                                      const void *bvh, int64_t num_triangles, float epsilon,
                                      int64_t batch_size, int32_t *out_index, float *out_t);
/* out [B,T] u8 (zero-filled by the callee): out[b, first_hit_index[b, r]] = 1 for every hit */
DRT_API int drt_scatter_visible(drt_stream_t stream, int64_t num_vertices_batch, int64_t num_rays,
                        int64_t num_triangles, const int32_t *first_hit_index, uint8_t *out);

/* ---------------------------------------------------------------------------------------------
 * N3  shooting-and-bouncing rays.  The nearest hit of every ray comes from
 *     drt_first_triangle_hit_by_ray between the steps; these entry points are the element-wise
 *     remainder of one bounce.
 *  drt_sbr_bounce: SBRPathLauncher.launch_paths scan body (reference: _solvers.py:279-356, 407-444):
 *     masks [num_tx, num_rx, num_rays] u8 of THIS bounce = receiver within sqrt(max_dist) of the ray
 *     before its hit (filter_rays), then origins / directions / valid [num_tx, num_rays] are bounced
 *     in place (bounce_rays); vertices_out (nullable) [num_tx, num_rays, 3] receives the new origins.
 *  drt_mlm_step: one iteration of the multipath-lifetime-map kernel (reference: _scene.py:81-171):
 *     receiver-plane crossing test + atomic OR of the path hash into output [num_tx, dim_x, dim_y],
 *     then reflection, hash update (hash constants _scene.py:60-78) and the epsilon offset of the
 *     next query origin.  Rays that miss are retired (alive = 0).
 * ------------------------------------------------------------------------------------------- */
DRT_API int drt_sbr_bounce(drt_stream_t stream, int64_t num_tx, int64_t num_rays, int64_t num_rx,
                   int64_t num_triangles, const void *pack, float *origins, float *directions,
                   uint8_t *valid, const int32_t *faces, const float *t_hit, const float *rx,
                   float max_dist, uint8_t *masks, float *vertices_out /*nullable*/);
DRT_API int drt_mlm_step(drt_stream_t stream, int64_t num_tx, int64_t num_rays, int64_t num_triangles,
                 const void *pack, float *origins, float *directions, uint32_t *hashes,
                 uint8_t *alive, const int32_t *faces, const float *t_first, int32_t iteration,
                 int32_t min_order, int32_t assume_quads, float receiver_height, float min_x,
                 float max_x, float min_y, float max_y, int32_t dim_x, int32_t dim_y, float epsilon,
                 uint32_t *output);

#ifdef __cplusplus
}
#endif
#endif /* DIFFERT_B200_H */
