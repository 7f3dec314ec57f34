/* A pure-C99 consumer of the differt_b200 C ABI: no Python, no PyTorch, no C++.
 *
 *   gcc -std=c99 -Iinclude -I/usr/local/cuda/include integration/c_abi_example.c \
 *       -Ldiffert_b200 -ldiffert_b200 -L/usr/local/cuda/lib64 -lcudart -o c_abi_example
 *   LD_LIBRARY_PATH=differt_b200 ./c_abi_example
 *
 * It builds the unit box of the reference's visibility tests (Mesh.box(with_top=True), 12 triangles,
 * differt/src/differt/geometry/_mesh.py:2172-2208), asks which of six axis rays from the centre are
 * blocked (all of them), finds the nearest hits (t = 0.5 on every axis), then traces every order-1
 * path candidate between two points inside the box and prints the valid ones.  tests/test_abi.py
 * compiles this file on the CPU box (the header must be valid C) and the GPU suite runs it. */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "differt_b200.h"

#define CHECK_CUDA(x)                                                        \
    do {                                                                     \
        cudaError_t e_ = (x);                                                \
        if (e_ != cudaSuccess) {                                             \
            fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(e_));     \
            return 2;                                                        \
        }                                                                    \
    } while (0)
#define CHECK_DRT(x)                                                         \
    do {                                                                     \
        int rc_ = (x);                                                       \
        if (rc_ != DRT_OK) {                                                 \
            fprintf(stderr, "differt_b200: %s\n", drt_error_string(rc_));    \
            return 3;                                                        \
        }                                                                    \
    } while (0)

static void *to_device(const void *host, size_t bytes) {
    void *d = NULL;
    if (cudaMalloc(&d, bytes) != cudaSuccess) return NULL;
    if (cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return NULL;
    return d;
}

int main(void) {
    /* unit box centred at the origin, the reference's vertex and triangle order */
    const float h = 0.5f;
    const float vertices[8][3] = {{+h, +h, +h}, {+h, +h, -h}, {-h, +h, -h}, {-h, +h, +h},
                                  {-h, -h, -h}, {-h, -h, +h}, {+h, -h, -h}, {+h, -h, +h}};
    const int32_t triangles[12][3] = {{0, 1, 2}, {0, 2, 3}, {3, 2, 4}, {3, 4, 5}, {5, 4, 6}, {5, 6, 7},
                                      {7, 6, 1}, {7, 1, 0}, {1, 4, 2}, {1, 6, 4}, {0, 3, 5}, {0, 5, 7}};
    const int64_t V = 8, T = 12, R = 6;
    const float origins[6][3] = {{0}};
    const float directions[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    const float eps = 10.0f * 1.1920929e-7f, hit_tol = 100.0f * 1.1920929e-7f;

    if (drt_abi_version() != DRT_ABI_VERSION) return 1;
    cudaStream_t stream;
    CHECK_CUDA(cudaStreamCreate(&stream));
    float *d_v = (float *)to_device(vertices, sizeof vertices);
    int32_t *d_t = (int32_t *)to_device(triangles, sizeof triangles);
    float *d_o = (float *)to_device(origins, sizeof origins);
    float *d_d = (float *)to_device(directions, sizeof directions);
    void *d_pack = NULL;
    uint8_t *d_hit = NULL;
    int32_t *d_idx = NULL;
    float *d_dist = NULL;
    CHECK_CUDA(cudaMalloc(&d_pack, drt_mesh_pack_bytes(T)));
    CHECK_CUDA(cudaMalloc((void **)&d_hit, R));
    CHECK_CUDA(cudaMalloc((void **)&d_idx, R * sizeof(int32_t)));
    CHECK_CUDA(cudaMalloc((void **)&d_dist, R * sizeof(float)));
    if (!d_v || !d_t || !d_o || !d_d) return 2;

    CHECK_DRT(drt_mesh_pack(stream, V, T, d_v, d_t, NULL, d_pack));
    CHECK_DRT(drt_ray_intersect_any_triangle(stream, R, d_o, d_d, d_pack, T, eps, hit_tol, d_hit, NULL));
    CHECK_DRT(drt_first_triangle_hit_by_ray(stream, R, d_o, d_d, d_pack, T, eps, 512, d_idx, d_dist, NULL));
    uint8_t hit[6];
    int32_t idx[6];
    float dist[6];
    CHECK_CUDA(cudaMemcpyAsync(hit, d_hit, sizeof hit, cudaMemcpyDeviceToHost, stream));
    CHECK_CUDA(cudaMemcpyAsync(idx, d_idx, sizeof idx, cudaMemcpyDeviceToHost, stream));
    CHECK_CUDA(cudaMemcpyAsync(dist, d_dist, sizeof dist, cudaMemcpyDeviceToHost, stream));
    CHECK_CUDA(cudaStreamSynchronize(stream));
    int ok = 1;
    for (int r = 0; r < 6; ++r) {
        printf("ray %d: blocked=%d first_hit=%d t=%.3f\n", r, hit[r], idx[r], dist[r]);
        ok = ok && hit[r] == 1 && idx[r] >= 0 && dist[r] == 0.5f;
    }

    /* fused trace: every order-1 candidate between two points inside the box */
    const float tx[1][3] = {{0.2f, 0.1f, 0.0f}}, rx[1][3] = {{-0.3f, 0.2f, 0.1f}};
    int32_t cand[12][1];
    for (int c = 0; c < 12; ++c) cand[c][0] = c;
    float *d_tx = (float *)to_device(tx, sizeof tx), *d_rx = (float *)to_device(rx, sizeof rx);
    int32_t *d_cand = (int32_t *)to_device(cand, sizeof cand);
    const size_t ws_bytes = drt_trace_workspace_bytes(T, 1, 1, 12);
    void *d_ws = NULL;
    float *d_pv = NULL;
    int32_t *d_po = NULL;
    uint8_t *d_pm = NULL;
    CHECK_CUDA(cudaMalloc(&d_ws, ws_bytes));
    CHECK_CUDA(cudaMalloc((void **)&d_pv, 12 * 3 * 3 * sizeof(float)));
    CHECK_CUDA(cudaMalloc((void **)&d_po, 12 * 3 * sizeof(int32_t)));
    CHECK_CUDA(cudaMalloc((void **)&d_pm, 12));
    CHECK_DRT(drt_trace_path_candidates(stream, V, T, d_v, d_t, NULL, 0, 1, d_tx, 1, d_rx, 12, 1, d_cand, eps,
                                        hit_tol, eps, 0u, d_ws, ws_bytes, d_pv, d_po, d_pm, NULL));
    uint8_t mask[12];
    float pv[12][3][3];
    CHECK_CUDA(cudaMemcpyAsync(mask, d_pm, sizeof mask, cudaMemcpyDeviceToHost, stream));
    CHECK_CUDA(cudaMemcpyAsync(pv, d_pv, sizeof pv, cudaMemcpyDeviceToHost, stream));
    CHECK_CUDA(cudaStreamSynchronize(stream));
    int valid = 0;
    for (int c = 0; c < 12; ++c)
        if (mask[c]) {
            ++valid;
            printf("valid order-1 path via triangle %d: reflection point (%.4f, %.4f, %.4f)\n", c, pv[c][1][0],
                   pv[c][1][1], pv[c][1][2]);
        }
    /* inside a closed box every one of the 6 faces reflects exactly one path */
    ok = ok && valid == 6;
    printf("%s: %d valid order-1 paths\n", ok ? "OK" : "MISMATCH", valid);
    return ok ? 0 : 4;
}
