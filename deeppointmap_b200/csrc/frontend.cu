// frontend.cu -- the per-frame preprocessing chain of the inference YAML that needs no neighbourhood search
// (SURVEY.md section 8f rank 4), raw KITTI .bin rows in, encoder input out, without a CPU round trip:
//   BinReader                 dataloader/heads/bin.py:16-17        float32 (N,4) rows, NaN rows dropped
//   VoxelSample(v, 'first')   dataloader/transforms.py:331-356     one point per voxel: the lowest original
//                                                                  index (np.unique return_index), output in
//                                                                  ascending voxel-id order
//   DistanceSample(lo, hi)    dataloader/transforms.py:387-397     lo <= |p| <= hi
//   CoordinatesNormalization  dataloader/transforms.py:400-407     p / ratio
// HBM-bound integer work: a dense "first index per voxel" table (atomicMin), then an ordered compaction of
// the table (two passes: per-chunk counts, scan, emit).  Everything data dependent (grid extent, number of
// survivors) stays on the device; `count` = -1 flags a voxel grid larger than the caller's table.
#include <limits.h>

#include "common.cuh"

namespace dpm {

struct FeParams {
    float mnx, mny, mnz, pad;
    int X, Y, Z, valid;
    long long nvox;
};

constexpr int FE_CH = 4096;  // table entries per block in the compaction passes (256 threads x 16)

__device__ __forceinline__ bool fe_row_ok(float x, float y, float z) { return !(isnan(x) || isnan(y) || isnan(z)); }

__global__ void __launch_bounds__(1024)
fe_minmax_kernel(const float *__restrict__ raw, int N, int stride, float voxel, long long max_voxels, FeParams *__restrict__ P,
                 int32_t *__restrict__ count) {
    __shared__ float red[6][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float INF = __int_as_float(0x7f800000);
    float lo[3] = {INF, INF, INF}, hi[3] = {-INF, -INF, -INF};
    for (int i = tid; i < N; i += 1024) {
        const float x = raw[(size_t)i * stride], y = raw[(size_t)i * stride + 1], z = raw[(size_t)i * stride + 2];
        if (!fe_row_ok(x, y, z)) continue;
        lo[0] = fminf(lo[0], x); hi[0] = fmaxf(hi[0], x);
        lo[1] = fminf(lo[1], y); hi[1] = fmaxf(hi[1], y);
        lo[2] = fminf(lo[2], z); hi[2] = fmaxf(hi[2], z);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if (lane == 0) { red[a][warp] = lo[a]; red[3 + a][warp] = hi[a]; }
    }
    __syncthreads();
    if (tid == 0) {
        for (int a = 0; a < 3; ++a)
            for (int w = 1; w < 32; ++w) {
                red[a][0] = fminf(red[a][0], red[a][w]);
                red[3 + a][0] = fmaxf(red[3 + a][0], red[3 + a][w]);
            }
        FeParams p;
        p.mnx = red[0][0]; p.mny = red[1][0]; p.mnz = red[2][0]; p.pad = 0.f;
        p.valid = red[0][0] <= red[3][0];  // at least one non-NaN row
        p.X = p.Y = p.Z = 0;
        p.nvox = 0;
        if (p.valid) {
            // ((xyz_max - xyz_min) / voxel_size).astype(np.int32) + 1, fp32 (transforms.py:339)
            const float ex = __fdiv_rn(__fsub_rn(red[3][0], red[0][0]), voxel), ey = __fdiv_rn(__fsub_rn(red[4][0], red[1][0]), voxel),
                        ez = __fdiv_rn(__fsub_rn(red[5][0], red[2][0]), voxel);
            if (ex < 2.0e9f && ey < 2.0e9f && ez < 2.0e9f) {
                p.X = (int)ex + 1; p.Y = (int)ey + 1; p.Z = (int)ez + 1;
                const double nv = (double)p.X * (double)p.Y * (double)p.Z;
                p.nvox = nv <= (double)max_voxels ? (long long)nv : -1;
            } else {
                p.nvox = -1;
            }
        }
        *P = p;
        *count = p.nvox < 0 ? -1 : 0;
    }
}

__global__ void __launch_bounds__(256) fe_fill_kernel(const FeParams *__restrict__ P, int32_t *__restrict__ table) {
    const long long nvox = P->nvox;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < nvox; i += (long long)gridDim.x * 256) table[i] = INT_MAX;
}

__global__ void __launch_bounds__(256)
fe_scatter_kernel(const float *__restrict__ raw, int N, int stride, float voxel, const FeParams *__restrict__ P,
                  int32_t *__restrict__ table) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    const FeParams p = *P;
    if (p.nvox <= 0) return;
    const float x = raw[(size_t)i * stride], y = raw[(size_t)i * stride + 1], z = raw[(size_t)i * stride + 2];
    if (!fe_row_ok(x, y, z)) return;
    // voxel_xyz = ((xyz - xyz_min) / voxel_size).astype(np.int32); id = vx + vy * X + vz * X * Y (transforms.py:341-343)
    const int vx = (int)__fdiv_rn(__fsub_rn(x, p.mnx), voxel), vy = (int)__fdiv_rn(__fsub_rn(y, p.mny), voxel),
              vz = (int)__fdiv_rn(__fsub_rn(z, p.mnz), voxel);
    const long long id = (long long)vx + (long long)vy * p.X + (long long)vz * p.X * p.Y;
    if (id >= 0 && id < p.nvox) atomicMin(&table[id], i);
}

// does table entry e survive?  (its point, when it does)
__device__ __forceinline__ bool fe_keep(const float *__restrict__ raw, int stride, int32_t idx, float lo, float hi, float &x,
                                        float &y, float &z) {
    if (idx == INT_MAX) return false;
    x = raw[(size_t)idx * stride]; y = raw[(size_t)idx * stride + 1]; z = raw[(size_t)idx * stride + 2];
    const float d = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));  // torch.norm(p=2, dim=1)
    return lo <= d && d <= hi;
}

__global__ void __launch_bounds__(256)
fe_count_kernel(const float *__restrict__ raw, int stride, const FeParams *__restrict__ P, const int32_t *__restrict__ table,
                float lo, float hi, int *__restrict__ bcount) {
    __shared__ int wsum[8];
    const long long nvox = P->nvox;
    const long long e0 = (long long)blockIdx.x * FE_CH;
    if (e0 >= nvox) return;  // block-uniform
    int c = 0;
    for (int k = 0; k < 16; ++k) {
        const long long e = e0 + (long long)threadIdx.x * 16 + k;
        float x, y, z;
        if (e < nvox && fe_keep(raw, stride, table[e], lo, hi, x, y, z)) ++c;
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += wsum[w];
        bcount[blockIdx.x] = t;
    }
}

// exclusive scan of the per-chunk counts (one block), total -> count
__global__ void __launch_bounds__(1024)
fe_scan_kernel(const FeParams *__restrict__ P, int *__restrict__ bcount, int32_t *__restrict__ count) {
    __shared__ int wsum[32];
    const long long nvox = P->nvox;
    if (nvox < 0) return;  // count already holds -1
    const int nblk = (int)((nvox + FE_CH - 1) / FE_CH);
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
    if (tid == 1023) *count = run;  // the last thread's running sum is the total
}

__global__ void __launch_bounds__(256)
fe_emit_kernel(const float *__restrict__ raw, int stride, const FeParams *__restrict__ P, const int32_t *__restrict__ table,
               float lo, float hi, float ratio, const int *__restrict__ boffset, float *__restrict__ out) {
    __shared__ int wsum[8];
    const long long nvox = P->nvox;
    const long long e0 = (long long)blockIdx.x * FE_CH;
    if (e0 >= nvox) return;
    float px[16], py[16], pz[16];
    unsigned keep = 0u;
    int c = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const long long e = e0 + (long long)threadIdx.x * 16 + k;
        if (e < nvox && fe_keep(raw, stride, table[e], lo, hi, px[k], py[k], pz[k])) { keep |= 1u << k; ++c; }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int before = boffset[blockIdx.x] + incl - c;
    for (int w = 0; w < warp; ++w) before += wsum[w];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        if (keep & (1u << k)) {
            float *o = out + (size_t)before * 3;
            o[0] = __fdiv_rn(px[k], ratio); o[1] = __fdiv_rn(py[k], ratio); o[2] = __fdiv_rn(pz[k], ratio);
            ++before;
        }
    }
}

struct FeWs {
    FeParams *params;
    int32_t *table;
    int *bcount;
    int nblk;
};

static bool fe_carve(Arena &a, long long max_voxels, FeWs *w) {
    w->params = a.get<FeParams>(1);
    w->table = a.get<int32_t>((size_t)max_voxels);
    w->nblk = (int)((max_voxels + FE_CH - 1) / FE_CH);
    w->bcount = a.get<int>((size_t)w->nblk);
    return a.ok();
}

}  // namespace dpm

using namespace dpm;

extern "C" size_t dpm_frontend_workspace_bytes(long long max_voxels) {
    if (max_voxels <= 0 || max_voxels > (1LL << 31) - 1) return 0;
    Arena a(nullptr, 0);
    FeWs w;
    fe_carve(a, max_voxels, &w);
    return a.off + 256;
}

extern "C" int dpm_frontend_f32(const float *raw, int N, int stride, float voxel_size, float min_dis, float max_dis,
                                float ratio, long long max_voxels, float *out_rows, int32_t *count, void *ws,
                                size_t ws_bytes, dpm_stream_t stream) {
    if (!raw || !out_rows || !count || !ws) return fail(DPM_ERR_ARG, "frontend: null pointer");
    if (N <= 0 || stride < 3) return fail(DPM_ERR_SHAPE, "frontend: bad shape N=%d stride=%d", N, stride);
    if (!(voxel_size > 0.f) || !(ratio > 0.f)) return fail(DPM_ERR_ARG, "frontend: voxel_size and ratio must be > 0");
    if (max_voxels <= 0 || max_voxels > (1LL << 31) - 1) return fail(DPM_ERR_ARG, "frontend: max_voxels out of range");
    cudaStream_t st = (cudaStream_t)stream;
    Arena a(ws, ws_bytes);
    FeWs w;
    if (!fe_carve(a, max_voxels, &w)) return fail(DPM_ERR_WORKSPACE, "frontend: workspace too small");
    prof_mark(st);
    fe_minmax_kernel<<<1, 1024, 0, st>>>(raw, N, stride, voxel_size, max_voxels, w.params, count);
    DPM_CHECK_LAUNCH("fe_minmax", st);
    const int sms = device_sm_count();
    fe_fill_kernel<<<sms * 8, 256, 0, st>>>(w.params, w.table);
    DPM_CHECK_LAUNCH("fe_fill", st);
    fe_scatter_kernel<<<(N + 255) / 256, 256, 0, st>>>(raw, N, stride, voxel_size, w.params, w.table);
    DPM_CHECK_LAUNCH("fe_scatter", st);
    fe_count_kernel<<<w.nblk, 256, 0, st>>>(raw, stride, w.params, w.table, min_dis, max_dis, w.bcount);
    DPM_CHECK_LAUNCH("fe_count", st);
    fe_scan_kernel<<<1, 1024, 0, st>>>(w.params, w.bcount, count);
    DPM_CHECK_LAUNCH("fe_scan", st);
    fe_emit_kernel<<<w.nblk, 256, 0, st>>>(raw, stride, w.params, w.table, min_dis, max_dis, ratio, w.bcount, out_rows);
    DPM_CHECK_LAUNCH("fe_emit", st);
    return DPM_OK;
}
