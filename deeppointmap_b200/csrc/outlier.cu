// outlier.cu -- OutlierFilter, the reference's CUDA branch (dataloader/transforms.py:230-246), as one native
// call with no host sync: self-kNN (K = nb_neighbors + 1, first column = the point itself, dropped),
// per-point mean of the neighbour DISTANCES (sqrt of the squared distances), global mean / unbiased std of that
// statistic, keep `stat <= mean + std_ratio * std`, survivors in their original order.
// kNN = knn_ring_kernel / knn_kernel (bit-exact pytorch3d contract); the statistic is summed left to right in
// fp32 like a row mean, the global moments in fp64 (rounded to fp32 before the threshold is formed, as the
// reference's fp32 tensors are) -- a point whose statistic sits within an ulp-scale distance of the threshold
// can therefore fall on the other side than in a particular torch build; everything else is index work.
#include "common.cuh"

namespace dpm {

constexpr int OC_CH = 2048;  // points per block in the compaction passes (256 threads x 8)

__global__ void __launch_bounds__(256)
oc_stat_kernel(const float *__restrict__ d2, int N, int K, float *__restrict__ stat, double *__restrict__ acc) {
    __shared__ double red[8][2];
    const int i = blockIdx.x * 256 + threadIdx.x;
    double s = 0.0, ss = 0.0;
    if (i < N) {
        float sum = 0.f;
        for (int k = 1; k < K; ++k) sum = __fadd_rn(sum, sqrtf(d2[(size_t)i * K + k]));  // torch.sqrt(dists[:, 1:]).mean(1)
        const float m = __fdiv_rn(sum, (float)(K - 1));
        stat[i] = m;
        s = (double)m;
        ss = (double)m * (double)m;
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

__global__ void oc_thr_kernel(const double *__restrict__ acc, int N, float ratio, float *__restrict__ thr) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double mean = acc[0] / (double)N;
    double var = N > 1 ? (acc[1] - (double)N * mean * mean) / (double)(N - 1) : __longlong_as_double(0x7ff8000000000000LL);
    if (var < 0.0) var = 0.0;
    const float mf = (float)mean, sf = (float)sqrt(var);   // N == 1: std is NaN, nothing passes `<=` (as torch)
    *thr = __fadd_rn(mf, __fmul_rn(ratio, sf));          // mean + std_ratio * std on fp32 tensors
}

__global__ void __launch_bounds__(256)
oc_count_kernel(const float *__restrict__ stat, const float *__restrict__ thr, int N, int *__restrict__ bcount) {
    __shared__ int wsum[8];
    const float t = *thr;
    int c = 0;
    for (int k = 0; k < 8; ++k) {
        const int i = blockIdx.x * OC_CH + threadIdx.x * 8 + k;
        if (i < N && stat[i] <= t) ++c;
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

__global__ void __launch_bounds__(1024) oc_scan_kernel(int *__restrict__ bcount, int nblk, int32_t *__restrict__ count) {
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
oc_emit_kernel(const float *__restrict__ rows, int stride, const float *__restrict__ stat, const float *__restrict__ thr, int N,
               const int *__restrict__ boffset, float *__restrict__ out, uint8_t *__restrict__ mask, float div) {
    __shared__ int wsum[8];
    const float t = *thr;
    unsigned keep = 0u;
    int c = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int i = blockIdx.x * OC_CH + threadIdx.x * 8 + k;
        const bool kp = i < N && stat[i] <= t;
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
            const int i = blockIdx.x * OC_CH + threadIdx.x * 8 + k;
            float *o = out + (size_t)before * 3;
            o[0] = __fdiv_rn(rows[(size_t)i * stride], div);  // CoordinatesNormalization rides along (IEEE division)
            o[1] = __fdiv_rn(rows[(size_t)i * stride + 1], div);
            o[2] = __fdiv_rn(rows[(size_t)i * stride + 2], div);
            ++before;
        }
    }
}

struct OcWs {
    float4 *p4;
    float *d2, *stat, *thr;
    double *acc;
    int *bcount, *len;
    GridWs grid;
    int nblk;
    bool use_grid;
};

static bool oc_carve(Arena &a, int N, int K, OcWs *w) {
    w->p4 = a.get<float4>((size_t)N);
    w->d2 = a.get<float>((size_t)N * K);
    w->stat = a.get<float>((size_t)N);
    w->thr = a.get<float>(4);
    w->acc = a.get<double>(4);
    w->nblk = (N + OC_CH - 1) / OC_CH;
    w->bcount = a.get<int>((size_t)w->nblk);
    w->len = a.get<int>(1);
    w->use_grid = N >= GRID_MIN_N && N <= GRID_MAX_N;
    if (w->use_grid) grid_ws_carve(a, 1, N, &w->grid);
    return a.ok();
}

}  // namespace dpm

using namespace dpm;

extern "C" size_t dpm_outlier_filter_workspace_bytes(int N, int nb_neighbors) {
    if (N <= 0 || nb_neighbors < 1 || nb_neighbors > 31) return 0;
    Arena a(nullptr, 0);
    OcWs w;
    oc_carve(a, N, nb_neighbors + 1, &w);
    return a.off + 256;
}

extern "C" int dpm_outlier_filter_f32(const float *rows, int N, int stride, int nb_neighbors, float std_ratio,
                                      float out_divisor, float *out_rows, uint8_t *mask, int32_t *count, void *ws,
                                      size_t ws_bytes, dpm_stream_t stream) {
    if (!rows || !out_rows || !count || !ws) return fail(DPM_ERR_ARG, "outlier_filter: null pointer");
    if (N <= 0 || stride < 3) return fail(DPM_ERR_SHAPE, "outlier_filter: bad shape N=%d stride=%d", N, stride);
    if (!(out_divisor > 0.f)) return fail(DPM_ERR_ARG, "outlier_filter: out_divisor must be > 0 (1 = none)");
    if (nb_neighbors < 1 || nb_neighbors > 31) return fail(DPM_ERR_UNSUPPORTED, "outlier_filter: nb_neighbors=%d not in 1..31", nb_neighbors);
    const int K = nb_neighbors + 1;
    cudaStream_t st = (cudaStream_t)stream;
    Arena a(ws, ws_bytes);
    OcWs w;
    if (!oc_carve(a, N, K, &w)) return fail(DPM_ERR_WORKSPACE, "outlier_filter: workspace too small");
    prof_mark(st);
    DPM_TRY(pack_xyz4_launch(rows, 1, N, stride, w.p4, st));
    DPM_CHECK_CUDA(cudaMemsetAsync(w.acc, 0, sizeof(double) * 4, st));
    if (w.use_grid) {
        DPM_TRY(lengths_to_i32_launch(nullptr, 1, N, w.len, st));
        DPM_TRY(grid_build_launch(w.p4, 1, N, w.len, 0.f, w.grid, st));
        DPM_TRY(knn_ring_launch(w.grid, w.p4, 1, N, nullptr, K, nullptr, nullptr, w.d2, st));
    } else {
        DPM_TRY(knn_launch(w.p4, w.p4, 1, N, N, nullptr, nullptr, K, 0.f, KNN_MODE_KNN, nullptr, nullptr, w.d2, st));
    }
    oc_stat_kernel<<<(N + 255) / 256, 256, 0, st>>>(w.d2, N, K, w.stat, w.acc);
    DPM_CHECK_LAUNCH("oc_stat", st);
    oc_thr_kernel<<<1, 32, 0, st>>>(w.acc, N, std_ratio, w.thr);
    DPM_CHECK_LAUNCH("oc_thr", st);
    oc_count_kernel<<<w.nblk, 256, 0, st>>>(w.stat, w.thr, N, w.bcount);
    DPM_CHECK_LAUNCH("oc_count", st);
    oc_scan_kernel<<<1, 1024, 0, st>>>(w.bcount, w.nblk, count);
    DPM_CHECK_LAUNCH("oc_scan", st);
    oc_emit_kernel<<<w.nblk, 256, 0, st>>>(rows, stride, w.stat, w.thr, N, w.bcount, out_rows, mask, out_divisor);
    DPM_CHECK_LAUNCH("oc_emit", st);
    return DPM_OK;
}
