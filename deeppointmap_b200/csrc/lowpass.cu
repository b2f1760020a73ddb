// lowpass.cu -- LowPassFilter (dataloader/transforms.py:256-297), the last data-dependent step of the shipped
// transform chain, as one native call with no host sync:
//   normals      = per-point PCA normal of the neighbours within normals_radius          (open3d estimate_normals)
//   idx          = knn_points(xyz, xyz, K = normals_num + 1)[..., 1:]                     (the point itself dropped)
//   similarity   = |normals[idx] . normals|                                               (N, normals_num)
//   sim          = sum of the `flux` largest similarities per point
//   keep         = sim > sim.mean() - filter_std * sim.std()                              (unbiased std)
// survivors keep their order; the division of CoordinatesNormalization can ride along in the emit pass.  The
// reference's `max_remain` re-ranking (off in every shipped YAML) is done by the Python front-end from `sim`.
// Points whose statistic lies within rounding distance of the threshold, or whose neighbourhood has two nearly equal
// smallest eigenvalues (the normal is then ill-defined in ANY implementation), may fall on the other side than with
// open3d's fp64 kd-tree / eigen-solver; tests/test_gpu_frontend.py bounds that fraction.
#include "common.cuh"

namespace dpm {

constexpr int LP_CH = 2048;  // points per block in the compaction passes (256 threads x 8)
constexpr int LP_MAXFLUX = 8;

__global__ void __launch_bounds__(256)
lp_sim_kernel(const float *__restrict__ normals, const int32_t *__restrict__ idx, int N, int K, int flux,
              float *__restrict__ sim, double *__restrict__ acc) {
    __shared__ double red[8][2];
    const int i = blockIdx.x * 256 + threadIdx.x;
    double s = 0.0, ss = 0.0;
    if (i < N) {
        const float nx = normals[(size_t)i * 3], ny = normals[(size_t)i * 3 + 1], nz = normals[(size_t)i * 3 + 2];
        float top[LP_MAXFLUX];
#pragma unroll
        for (int f = 0; f < LP_MAXFLUX; ++f) top[f] = -1.f;  // similarities are >= 0
        for (int k = 1; k < K; ++k) {
            const int j = idx[(size_t)i * K + k];
            float v = 0.f;
            if (j >= 0) v = fabsf(normals[(size_t)j * 3] * nx + normals[(size_t)j * 3 + 1] * ny + normals[(size_t)j * 3 + 2] * nz);
#pragma unroll
            for (int f = 0; f < LP_MAXFLUX; ++f) {  // insertion into the descending list
                if (f < flux && v > top[f]) { const float t = top[f]; top[f] = v; v = t; }
            }
        }
        float sum = 0.f;  // torch.topk(..., sorted) then .sum(1): largest first
#pragma unroll
        for (int f = 0; f < LP_MAXFLUX; ++f)
            if (f < flux && top[f] >= 0.f) sum = __fadd_rn(sum, top[f]);
        sim[i] = sum;
        s = (double)sum;
        ss = (double)sum * (double)sum;
    }
    s = warp_sum_d(s);
    ss = warp_sum_d(ss);
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = s; red[threadIdx.x >> 5][1] = ss; }
    __syncthreads();
    if (threadIdx.x < 2) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
        atomicAdd(&acc[threadIdx.x], t);
    }
}

__global__ void lp_thr_kernel(const double *__restrict__ acc, int N, float filter_std, float *__restrict__ thr) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double mean = acc[0] / (double)N;
    double var = N > 1 ? (acc[1] - (double)N * mean * mean) / (double)(N - 1) : __longlong_as_double(0x7ff8000000000000LL);
    if (var < 0.0) var = 0.0;
    const float mf = (float)mean, sf = (float)sqrt(var);
    *thr = __fsub_rn(mf, __fmul_rn(filter_std, sf));  // sim.mean() - filter_std * sim.std() on fp32 tensors
}

__global__ void __launch_bounds__(256)
lp_count_kernel(const float *__restrict__ stat, const float *__restrict__ thr, int N, int *__restrict__ bcount) {
    __shared__ int wsum[8];
    const float t = *thr;
    int c = 0;
    for (int k = 0; k < 8; ++k) {
        const int i = blockIdx.x * LP_CH + threadIdx.x * 8 + k;
        if (i < N && stat[i] > t) ++c;
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < 8; ++w) s += wsum[w];
        bcount[blockIdx.x] = s;
    }
}

// exclusive scan of the block counts by one block (same scheme as the outlier filter's)
__global__ void __launch_bounds__(1024) lp_scan_kernel(int *__restrict__ bcount, int nblk, int32_t *__restrict__ count) {
    __shared__ int wsum[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (nblk + 1023) / 1024;
    const int b0 = tid * per, b1 = min(nblk, b0 + per);
    int s = 0;
    for (int b = b0; b < b1; ++b) s += bcount[b];
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int v = wsum[lane], iv = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, iv, o);
            if (lane >= o) iv += t;
        }
        wsum[lane] = iv - v;
    }
    __syncthreads();
    int run = wsum[warp] + incl - s;
    for (int b = b0; b < b1; ++b) {
        const int n = bcount[b];
        bcount[b] = run;
        run += n;
    }
    if (tid == 1023) *count = run;
}

__global__ void __launch_bounds__(256)
lp_emit_kernel(const float *__restrict__ rows, int stride, const float *__restrict__ stat, const float *__restrict__ thr, int N,
               const int *__restrict__ boffset, float *__restrict__ out, uint8_t *__restrict__ mask, float div) {
    __shared__ int wsum[8];
    const float t = *thr;
    unsigned keep = 0u;
    int c = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int i = blockIdx.x * LP_CH + threadIdx.x * 8 + k;
        const bool kp = i < N && stat[i] > t;
        if (kp) { keep |= 1u << k; ++c; }
        if (mask && i < N) mask[i] = kp ? 1 : 0;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int before = boffset[blockIdx.x] + incl - c;
    for (int w = 0; w < warp; ++w) before += wsum[w];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (keep & (1u << k)) {
            const int i = blockIdx.x * LP_CH + threadIdx.x * 8 + k;
            float *o = out + (size_t)before * 3;
            o[0] = __fdiv_rn(rows[(size_t)i * stride], div);
            o[1] = __fdiv_rn(rows[(size_t)i * stride + 1], div);
            o[2] = __fdiv_rn(rows[(size_t)i * stride + 2], div);
            ++before;
        }
    }
}

struct LpWs {
    float4 *p4;
    float *normals, *sim, *thr;
    int32_t *idx;
    double *acc;
    int *bcount, *len;
    GridWs grid;
    int nblk;
    bool use_grid;
};

static bool lp_carve(Arena &a, int N, int K, LpWs *w) {
    w->p4 = a.get<float4>((size_t)N);
    w->normals = a.get<float>((size_t)N * 3);
    w->idx = a.get<int32_t>((size_t)N * K);
    w->sim = a.get<float>((size_t)N);
    w->thr = a.get<float>(4);
    w->acc = a.get<double>(4);
    w->nblk = (N + LP_CH - 1) / LP_CH;
    w->bcount = a.get<int>((size_t)w->nblk);
    w->len = a.get<int>(1);
    w->use_grid = N >= GRID_MIN_N && N <= GRID_MAX_N;
    if (w->use_grid) grid_ws_carve(a, 1, N, &w->grid);
    return a.ok();
}

}  // namespace dpm

using namespace dpm;

extern "C" size_t dpm_low_pass_filter_workspace_bytes(int N, int normals_num) {
    if (N <= 0 || normals_num < 1 || normals_num > 31) return 0;
    Arena a(nullptr, 0);
    LpWs w;
    lp_carve(a, N, normals_num + 1, &w);
    return a.off + 256;
}

extern "C" int dpm_low_pass_filter_f32(const float *rows, int N, int stride, float normals_radius, int normals_num,
                                       float filter_std, int flux, float out_divisor, float *out_rows, uint8_t *mask,
                                       float *sim_out, int32_t *count, void *ws, size_t ws_bytes, dpm_stream_t stream) {
    if (!rows || !out_rows || !count || !ws) return fail(DPM_ERR_ARG, "low_pass_filter: null pointer");
    if (N <= 0 || stride < 3) return fail(DPM_ERR_SHAPE, "low_pass_filter: bad shape N=%d stride=%d", N, stride);
    if (!(out_divisor > 0.f) || !(normals_radius > 0.f)) return fail(DPM_ERR_ARG, "low_pass_filter: radius / divisor must be > 0");
    if (normals_num < 1 || normals_num > 31) return fail(DPM_ERR_UNSUPPORTED, "low_pass_filter: normals_num=%d not in 1..31", normals_num);
    if (flux < 1 || flux > LP_MAXFLUX || flux > normals_num) return fail(DPM_ERR_UNSUPPORTED, "low_pass_filter: flux=%d not in 1..%d", flux, LP_MAXFLUX);
    if (N > GRID_MAX_N) return fail(DPM_ERR_UNSUPPORTED, "low_pass_filter: N=%d exceeds the limit %d", N, GRID_MAX_N);
    const int K = normals_num + 1;
    cudaStream_t st = (cudaStream_t)stream;
    Arena a(ws, ws_bytes);
    LpWs w;
    if (!lp_carve(a, N, K, &w)) return fail(DPM_ERR_WORKSPACE, "low_pass_filter: workspace too small");
    prof_mark(st);
    DPM_TRY(pack_xyz4_launch(rows, 1, N, stride, w.p4, st));
    DPM_CHECK_CUDA(cudaMemsetAsync(w.acc, 0, sizeof(double) * 4, st));
    if (w.use_grid) {
        DPM_TRY(lengths_to_i32_launch(nullptr, 1, N, w.len, st));
        DPM_TRY(grid_build_launch(w.p4, 1, N, w.len, normals_radius * 1.001f, w.grid, st));
        DPM_TRY(radius_normals_launch(&w.grid, w.p4, N, normals_radius, w.normals, st));
        DPM_TRY(knn_ring_launch(w.grid, w.p4, 1, N, nullptr, K, nullptr, w.idx, nullptr, st));
    } else {
        DPM_TRY(radius_normals_launch(nullptr, w.p4, N, normals_radius, w.normals, st));
        DPM_TRY(knn_launch(w.p4, w.p4, 1, N, N, nullptr, nullptr, K, 0.f, KNN_MODE_KNN, nullptr, w.idx, nullptr, st));
    }
    lp_sim_kernel<<<(N + 255) / 256, 256, 0, st>>>(w.normals, w.idx, N, K, flux, w.sim, w.acc);
    DPM_CHECK_LAUNCH("lp_sim", st);
    lp_thr_kernel<<<1, 32, 0, st>>>(w.acc, N, filter_std, w.thr);
    DPM_CHECK_LAUNCH("lp_thr", st);
    lp_count_kernel<<<w.nblk, 256, 0, st>>>(w.sim, w.thr, N, w.bcount);
    DPM_CHECK_LAUNCH("lp_count", st);
    lp_scan_kernel<<<1, 1024, 0, st>>>(w.bcount, w.nblk, count);
    DPM_CHECK_LAUNCH("lp_scan", st);
    lp_emit_kernel<<<w.nblk, 256, 0, st>>>(rows, stride, w.sim, w.thr, N, w.bcount, out_rows, mask, out_divisor);
    DPM_CHECK_LAUNCH("lp_emit", st);
    if (sim_out) DPM_CHECK_CUDA(cudaMemcpyAsync(sim_out, w.sim, sizeof(float) * (size_t)N, cudaMemcpyDeviceToDevice, st));
    return DPM_OK;
}
