// infomat.cu -- calculate_information_matrix_from_pcd (system/modules/utils.py:60-104, the
// pytorch3d branch): the 6x6 G^T G of the point-to-point ICP Jacobian over the correspondences
// "nearest target point within `radius` of every transformed source point".  The reference runs
// it on every odometry / scan-to-map / loop edge (odometry.py:115, mapping.py:159,
// loop_closure.py:247) as knn_points(K=1) + boolean-mask indexing + three batched outer products.
//
// Here: transform + pack (one launch), cell grid over the target, 1-NN capped at the radius over
// the 27-cell neighbourhood (knn_grid_kernel<PAD>, same (d2, index) order as knn_points), and one
// reduction kernel.  G^T G only needs ten sums over the matched target points t = (x, y, z):
//   n, Sx, Sy, Sz, Sxx, Syy, Szz, Sxy, Sxz, Syz   (rows of G per point: (0, z,-y,1,0,0), (-z,0,x,0,1,0), (y,-x,0,0,0,1))
// accumulated in fp64, so there is no (n,6,6) intermediate and no host sync.
#include "common.cuh"

namespace dpm {

// src (3,N1) channel-first -> q4 = R.p + T ; dst (3,N2) channel-first -> p4
__global__ void __launch_bounds__(256)
im_pack_kernel(const float *__restrict__ src, int N1, const float *__restrict__ dst, int N2, const float *__restrict__ SE3,
               float4 *__restrict__ q4, float4 *__restrict__ p4) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < N1) {
        const float x = src[i], y = src[(size_t)N1 + i], z = src[(size_t)2 * N1 + i];
        q4[i] = make_float4(fmaf(SE3[2], z, fmaf(SE3[1], y, SE3[0] * x)) + SE3[3],
                            fmaf(SE3[6], z, fmaf(SE3[5], y, SE3[4] * x)) + SE3[7],
                            fmaf(SE3[10], z, fmaf(SE3[9], y, SE3[8] * x)) + SE3[11], 0.f);
    }
    if (i < N2) p4[i] = make_float4(dst[i], dst[(size_t)N2 + i], dst[(size_t)2 * N2 + i], 0.f);
}

__global__ void __launch_bounds__(256)
im_accum_kernel(const float4 *__restrict__ p4, const int32_t *__restrict__ nn, const float *__restrict__ d2, float cap,
                int N1, double *__restrict__ acc) {
    __shared__ double red[8][10];
    double a[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) a[k] = 0.0;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < N1; i += gridDim.x * 256) {
        const int j = nn[i];
        if (j < 0 || (d2 && !(d2[i] < cap))) continue;
        const float4 t = p4[j];
        const double x = t.x, y = t.y, z = t.z;
        a[0] += 1.0; a[1] += x; a[2] += y; a[3] += z;
        a[4] += x * x; a[5] += y * y; a[6] += z * z;
        a[7] += x * y; a[8] += x * z; a[9] += y * z;
    }
#pragma unroll
    for (int k = 0; k < 10; ++k) a[k] = warp_sum_d(a[k]);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < 10; ++k) red[warp][k] = a[k];
    __syncthreads();
    if (threadIdx.x < 10) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
        atomicAdd(&acc[threadIdx.x], t);
    }
}

__global__ void im_final_kernel(const double *__restrict__ acc, float *__restrict__ info, int32_t *__restrict__ ncorr) {
    if (threadIdx.x != 0) return;
    const double n = acc[0], sx = acc[1], sy = acc[2], sz = acc[3], xx = acc[4], yy = acc[5], zz = acc[6], xy = acc[7],
                 xz = acc[8], yz = acc[9];
    double G[6][6] = {{zz + yy, -xy, -xz, 0.0, -sz, sy},
                      {-xy, zz + xx, -yz, sz, 0.0, -sx},
                      {-xz, -yz, yy + xx, -sy, sx, 0.0},
                      {0.0, sz, -sy, n, 0.0, 0.0},
                      {-sz, 0.0, sx, 0.0, n, 0.0},
                      {sy, -sx, 0.0, 0.0, 0.0, n}};
    for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 6; ++c) info[6 * r + c] = (float)G[r][c];
    if (ncorr) *ncorr = (int32_t)n;
}

struct ImWs {
    float4 *q4, *p4;
    int32_t *nn;
    float *d2;
    int *len;
    double *acc;
    GridWs grid;
    bool use_grid;
};

static bool im_carve(Arena &a, int N1, int N2, ImWs *w) {
    w->q4 = a.get<float4>((size_t)N1);
    w->p4 = a.get<float4>((size_t)N2);
    w->nn = a.get<int32_t>((size_t)N1);
    w->d2 = a.get<float>((size_t)N1);
    w->len = a.get<int>(1);
    w->acc = a.get<double>(16);
    w->use_grid = N2 >= GRID_MIN_N && N2 <= GRID_MAX_N;
    if (w->use_grid) grid_ws_carve(a, 1, N2, &w->grid);
    return a.ok();
}

}  // namespace dpm

using namespace dpm;

extern "C" size_t dpm_information_matrix_workspace_bytes(int N1, int N2) {
    if (N1 <= 0 || N2 <= 0) return 0;
    Arena a(nullptr, 0);
    ImWs w;
    im_carve(a, N1, N2, &w);
    return a.off + 256;
}

extern "C" int dpm_information_matrix_f32(const float *src, int N1, const float *dst, int N2, const float *SE3,
                                          float radius, float *info, int32_t *n_corr, void *ws, size_t ws_bytes,
                                          dpm_stream_t stream) {
    if (!src || !dst || !SE3 || !info || !ws) return fail(DPM_ERR_ARG, "information_matrix: null pointer");
    if (N1 <= 0 || N2 <= 0) return fail(DPM_ERR_SHAPE, "information_matrix: bad shape N1=%d N2=%d", N1, N2);
    if (!(radius >= 0.f)) return fail(DPM_ERR_ARG, "information_matrix: radius must be >= 0");
    cudaStream_t st = (cudaStream_t)stream;
    Arena a(ws, ws_bytes);
    ImWs w;
    if (!im_carve(a, N1, N2, &w)) return fail(DPM_ERR_WORKSPACE, "information_matrix: workspace too small");
    prof_mark(st);
    const int nmax = N1 > N2 ? N1 : N2;
    im_pack_kernel<<<(nmax + 255) / 256, 256, 0, st>>>(src, N1, dst, N2, SE3, w.q4, w.p4);
    DPM_CHECK_LAUNCH("im_pack", st);
    DPM_CHECK_CUDA(cudaMemsetAsync(w.acc, 0, sizeof(double) * 16, st));
    const float r2 = radius * radius;  // `dists <= radius ** 2` in fp32 (utils.py:84)
    if (w.use_grid) {
        DPM_TRY(lengths_to_i32_launch(nullptr, 1, N2, w.len, st));
        DPM_TRY(grid_build_launch(w.p4, 1, N2, w.len, radius * 1.001f, w.grid, st));
        DPM_TRY(knn_grid_launch(w.grid, w.q4, w.p4, 1, N1, N2, nullptr, 1, r2, nullptr, w.nn, st, /*pad=*/true));
        im_accum_kernel<<<148, 256, 0, st>>>(w.p4, w.nn, nullptr, 0.f, N1, w.acc);
    } else {
        DPM_TRY(knn_launch(w.q4, w.p4, 1, N1, N2, nullptr, nullptr, 1, 0.f, KNN_MODE_KNN, nullptr, w.nn, w.d2, st));
        im_accum_kernel<<<148, 256, 0, st>>>(w.p4, w.nn, w.d2, __builtin_nextafterf(r2, __builtin_inff()), N1, w.acc);
    }
    DPM_CHECK_LAUNCH("im_accum", st);
    im_final_kernel<<<1, 32, 0, st>>>(w.acc, info, n_corr);
    DPM_CHECK_LAUNCH("im_final", st);
    return DPM_OK;
}
