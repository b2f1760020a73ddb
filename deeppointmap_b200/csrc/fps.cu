// fps.cu -- farthest point sampling, bit-exact with the reference's Sampler.fps
// (network/encoder/utils.py:209-270) / pytorch3d sample_farthest_points.
//
// Design (B200): one thread-block CLUSTER per cloud.  The whole cloud lives in registers
// (P points per thread: x, y, z, running min-distance) for the entire K-1 step chain, loaded
// once with coalesced float4 reads; nothing is re-read from HBM.  Per step:
//   1. every thread updates its P min-distances against the last pick (exact fp32,
//      (dx*dx+dy*dy)+dz*dz, no FMA) and keeps the local max;
//   2. warp arg-max with two REDUX ops (max of the value bits, then min index among ties
//      = "first maximum");
//   3. each warp publishes ONE 24-byte record (value bits | ~index, xyz) into the shared
//      memory of EVERY CTA of the cluster (DSMEM stores), double-buffered by step parity;
//   4. one cluster barrier; every warp reduces the CS*16 records locally and knows the pick
//      and its coordinates -- no second barrier, no global-memory round trip.
#include <stdlib.h>

#include "common.cuh"

namespace dpm {

constexpr int FPS_T = 512;
constexpr int FPS_NW = FPS_T / 32;

__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned map_to_rank(const void *local_smem, unsigned rank) {
    unsigned l = (unsigned)__cvta_generic_to_shared(local_smem), r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(l), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_u64(unsigned addr, unsigned long long v) {
    asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_f4(unsigned addr, float4 v) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

template <int P, int CS>
__global__ void __launch_bounds__(FPS_T, 1)
fps_kernel(const float4 *__restrict__ xyz4, int N, const int *__restrict__ len32, int K,
           int64_t *__restrict__ idx64, int32_t *__restrict__ idx32, float4 *__restrict__ new_xyz4,
           uint8_t *__restrict__ new_pad, int *__restrict__ new_len32) {
    extern __shared__ float4 spts[];  // P * FPS_T: this CTA's points, for the winner's xyz
    __shared__ unsigned long long skey[2][CS * FPS_NW];
    __shared__ float4 sxyz[2][CS * FPS_NW];

    const int b = blockIdx.y;
    const unsigned rank = (CS > 1) ? cluster_ctarank() : 0u;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int len = len32 ? min(len32[b], N) : N;
    const float4 *pts = xyz4 + (size_t)b * N;
    const int kn = min(len, K);
    const bool writer = (rank == 0 && tid == 0);

    float x[P], y[P], z[P], m[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
        const int i = j * (CS * FPS_T) + (int)rank * FPS_T + tid;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        // points past `len` never win: min-distance 0 and a higher index than any valid point
        m[j] = 0.f;
        if (i < len) {
            p = pts[i];
            m[j] = __int_as_float(0x7f800000);
        }
        x[j] = p.x; y[j] = p.y; z[j] = p.z;
        spts[j * FPS_T + tid] = p;
    }

    float sx = 0.f, sy = 0.f, sz = 0.f;
    if (len > 0) {
        const float4 p0 = pts[0];
        sx = p0.x; sy = p0.y; sz = p0.z;
    }
    const size_t ob = (size_t)b * K;
    if (writer && kn > 0) {
        if (idx64) idx64[ob] = 0;
        if (idx32) idx32[ob] = 0;
        if (new_xyz4) new_xyz4[ob] = make_float4(sx, sy, sz, 0.f);
        if (new_pad) new_pad[ob] = 0;
    }
    if (CS > 1) cluster_barrier(); else __syncthreads();  // spts visible; all CTAs alive before DSMEM traffic

    int par = 0;
    for (int k = 1; k < kn; ++k) {
        float bm = 0.f;
#pragma unroll
        for (int j = 0; j < P; ++j) {
            const float d = d2_exact(sx, sy, sz, x[j], y[j], z[j]);
            m[j] = fminf(m[j], d);
            bm = fmaxf(bm, m[j]);
        }
        const unsigned bits = __float_as_uint(bm);  // bm >= 0: bit pattern is order preserving
        const unsigned wmax = __reduce_max_sync(0xffffffffu, bits);
        unsigned cand = 0xffffffffu;
        if (bits == wmax) {
            int jj = 0;
#pragma unroll
            for (int j = P - 1; j >= 0; --j)
                if (m[j] == bm) jj = j;  // lowest j = lowest index among this thread's ties
            cand = (unsigned)(jj * (CS * FPS_T) + (int)rank * FPS_T + tid);
        }
        const unsigned wmin = __reduce_min_sync(0xffffffffu, cand);
        const int src = __ffs(__ballot_sync(0xffffffffu, cand == wmin)) - 1;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane == src) {
            const int jj = ((int)wmin - (int)rank * FPS_T - tid) / (CS * FPS_T);
            p = spts[jj * FPS_T + tid];
        }
        p.x = __shfl_sync(0xffffffffu, p.x, src);
        p.y = __shfl_sync(0xffffffffu, p.y, src);
        p.z = __shfl_sync(0xffffffffu, p.z, src);
        const unsigned long long key = ((unsigned long long)wmax << 32) | (unsigned long long)(0xffffffffu - wmin);
        const int slot = (int)rank * FPS_NW + warp;
        if (CS > 1) {
            if (lane < CS) {
                st_cluster_u64(map_to_rank(&skey[par][slot], (unsigned)lane), key);
                st_cluster_f4(map_to_rank(&sxyz[par][slot], (unsigned)lane), p);
            }
            cluster_barrier();
        } else {
            if (lane == 0) {
                skey[par][slot] = key;
                sxyz[par][slot] = p;
            }
            __syncthreads();
        }
        // every warp reduces the CS*NW records on its own
        unsigned long long bk = 0ull;
        int bs = 0;
#pragma unroll
        for (int r = lane; r < CS * FPS_NW; r += 32) {
            const unsigned long long kk = skey[par][r];
            if (kk > bk) { bk = kk; bs = r; }
        }
        const unsigned hi = (unsigned)(bk >> 32), lo = (unsigned)bk;
        const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
        const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
        const int wl = __ffs(__ballot_sync(0xffffffffu, hi == mh && lo == ml)) - 1;
        bs = __shfl_sync(0xffffffffu, bs, wl);
        const float4 w = sxyz[par][bs];
        sx = w.x; sy = w.y; sz = w.z;
        if (writer) {
            const unsigned sel = 0xffffffffu - ml;
            if (idx64) idx64[ob + k] = (int64_t)sel;
            if (idx32) idx32[ob + k] = (int32_t)sel;
            if (new_xyz4) new_xyz4[ob + k] = make_float4(sx, sy, sz, 0.f);
            if (new_pad) new_pad[ob + k] = 0;
        }
        par ^= 1;
    }
    if (rank == 0) {
        for (int k = kn + tid; k < K; k += FPS_T) {  // K > len: idx -1, zero rows, padded
            if (idx64) idx64[ob + k] = -1;
            if (idx32) idx32[ob + k] = -1;
            if (new_xyz4) new_xyz4[ob + k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (new_pad) new_pad[ob + k] = 1;
        }
        if (tid == 0 && new_len32) new_len32[b] = kn;
    }
    if (CS > 1) cluster_barrier();  // no CTA may exit while peers can still address its smem
}

template <int P, int CS>
static int fps_launch_t(const float4 *xyz4, int B, int N, const int *len32, int K, int64_t *idx64,
                        int32_t *idx32, float4 *new_xyz4, uint8_t *new_pad, int *new_len32, cudaStream_t st) {
    auto kern = fps_kernel<P, CS>;
    const size_t smem = (size_t)P * FPS_T * sizeof(float4);
    static thread_local unsigned long long configured = 0ull;  // one bit per device: function attributes are per context
    const unsigned long long devbit = 1ull << (current_device() & 63);
    if (!(configured & devbit)) {
        DPM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (CS > 8) DPM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        configured |= devbit;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS, B, 1);
    cfg.blockDim = dim3(FPS_T, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    DPM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, xyz4, N, len32, K, idx64, idx32, new_xyz4, new_pad, new_len32));
    count_launch("fps", st);
    return DPM_OK;
}

// choose (points per thread, cluster size): the smallest cluster that holds the cloud in
// registers, widened (fewer points per thread = shorter steps) while the grid would
// otherwise leave SMs idle.
static void fps_pick(int N, int B, int *P, int *CS) {
    int cs = 1;
    while (cs < 16 && (long long)cs * FPS_T * 16 < N) cs *= 2;
    int p = 1;
    while (p < 16 && (long long)cs * FPS_T * p < N) p *= 2;
    const int sms = device_sm_count();
    // Widening used to pay when a batch left SMs idle; batches that reach this kernel are throughput work now (<= 15 clouds
    // take the cluster kernels of fps_cluster.cu), and with several streams in flight SM time is what counts: clusters of
    // 4 for the 4096-point level cost 7 % of the 32-frame step (6890 against 7480 frames/s).  DPM_FPS_REG_MAXCS=8 restores it.
    static const int maxcs = getenv("DPM_FPS_REG_MAXCS") ? atoi(getenv("DPM_FPS_REG_MAXCS")) : 1;
    while (cs < maxcs && p > 2 && (long long)B * cs * 2 <= sms) {
        cs *= 2;
        p /= 2;
    }
    *P = p;
    *CS = cs;
}

// Clouds of <= 8192 points stay in registers (512 threads x <= 16 points, 80 registers at 8 points: half an SM's threads
// and 60 % of its registers, against a whole SM for the pruned kernel, whose bucket bookkeeping only pays from ~10 000
// points on): level 1 of the encoder (4096 -> 1024) 7130 -> 7240 frames/s in the driver's 20-step run, 7480 -> 7535 at
// 200 steps.  DPM_FPS_GRID_MIN_N=2048 restores the old split.
int fps_grid_min_n() {
    static const int v = getenv("DPM_FPS_GRID_MIN_N") ? atoi(getenv("DPM_FPS_GRID_MIN_N")) : FPS_BRUTE_CLUSTER_MAX_N + 1;
    return v < GRID_MIN_N ? GRID_MIN_N : v;
}

int fps_launch(const float4 *xyz4, int B, int N, const int *len32, int K, int64_t *idx64, int32_t *idx32,
               float4 *new_xyz4, uint8_t *new_pad, int *new_len32, cudaStream_t st) {
    if (B <= 0 || N <= 0 || K <= 0) return fail(DPM_ERR_SHAPE, "fps: bad shape B=%d N=%d K=%d", B, N, K);
    if ((long long)N > 16LL * FPS_T * 16)
        return fail(DPM_ERR_UNSUPPORTED, "fps: N=%d exceeds the register-resident limit %d", N, 16 * FPS_T * 16);
    if (N <= FPS_BRUTE_CLUSTER_MAX_N && fps_cluster_mode_small(B))
        return fps_brute_cluster_launch(xyz4, B, N, len32, K, idx64, idx32, new_xyz4, new_pad, new_len32, st);
    int P, CS;
    fps_pick(N, B, &P, &CS);
    prof_note(N, K);
#define DPM_FPS_CASE(p, cs)                                                                             \
    if (P == p && CS == cs)                                                                             \
        return fps_launch_t<p, cs>(xyz4, B, N, len32, K, idx64, idx32, new_xyz4, new_pad, new_len32, st);
#define DPM_FPS_ROW(p) DPM_FPS_CASE(p, 1) DPM_FPS_CASE(p, 2) DPM_FPS_CASE(p, 4) DPM_FPS_CASE(p, 8) DPM_FPS_CASE(p, 16)
    DPM_FPS_ROW(1) DPM_FPS_ROW(2) DPM_FPS_ROW(4) DPM_FPS_ROW(8) DPM_FPS_ROW(16)
#undef DPM_FPS_ROW
#undef DPM_FPS_CASE
    return fail(DPM_ERR_UNSUPPORTED, "fps: no kernel for P=%d CS=%d", P, CS);
}

// ---- small helpers shared with knn.cu ---------------------------------------------------
__global__ void pack_xyz4_kernel(const float *__restrict__ src, long long rows, int D, float4 *__restrict__ dst) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows) {
        const float *p = src + i * D;
        dst[i] = make_float4(p[0], p[1], p[2], 0.f);
    }
}
int pack_xyz4_launch(const float *src, int B, int N, int D, float4 *dst, cudaStream_t st) {
    long long rows = (long long)B * N;
    if (rows == 0) return DPM_OK;
    pack_xyz4_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(src, rows, D, dst);
    DPM_CHECK_LAUNCH("pack_xyz4", st);
    return DPM_OK;
}
__global__ void lengths_to_i32_kernel(const int64_t *len64, int B, int N, int *len32) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) {
        long long v = len64 ? len64[i] : N;
        len32[i] = (int)(v < 0 ? 0 : (v > N ? N : v));
    }
}
int lengths_to_i32_launch(const int64_t *len64, int B, int N, int *len32, cudaStream_t st) {
    lengths_to_i32_kernel<<<(B + 127) / 128, 128, 0, st>>>(len64, B, N, len32);
    DPM_CHECK_LAUNCH("lengths_to_i32", st);
    return DPM_OK;
}

__global__ void gather_rows_kernel(const float *__restrict__ points, const int64_t *__restrict__ idx, int B, int N,
                                   int K, int D, float *__restrict__ out) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * K * D;
    if (t >= total) return;
    int d = (int)(t % D);
    long long r = t / D;
    int b = (int)(r / K);
    int64_t i = idx[r];
    out[t] = i < 0 ? 0.f : points[((size_t)b * N + i) * D + d];
}

}  // namespace dpm

using namespace dpm;

extern "C" size_t dpm_fps_workspace_bytes(int B, int N, int D, int K) {
    Arena a(nullptr, 0);
    a.get<float4>((size_t)B * N);
    a.get<int>(B);
    return a.off + grid_ws_bytes(B, N) + 256;
}

extern "C" int dpm_fps_f32(const float *points, int B, int N, int D, const int64_t *lengths, int K,
                           int64_t *idx_out, float *sampled_out, void *ws, size_t ws_bytes, dpm_stream_t stream) {
    if (!points || !idx_out) return fail(DPM_ERR_ARG, "fps: null pointer");
    if (B <= 0 || N <= 0 || D < 3 || K <= 0) return fail(DPM_ERR_SHAPE, "fps: bad shape B=%d N=%d D=%d K=%d", B, N, D, K);
    if (!ws || ws_bytes < dpm_fps_workspace_bytes(B, N, D, K))
        return fail(DPM_ERR_WORKSPACE, "fps: workspace too small (%zu < %zu)", ws_bytes, dpm_fps_workspace_bytes(B, N, D, K));
    cudaStream_t st = (cudaStream_t)stream;
    Arena a(ws, ws_bytes);
    float4 *xyz4 = a.get<float4>((size_t)B * N);
    int *len32 = a.get<int>(B);
    DPM_TRY(pack_xyz4_launch(points, B, N, D, xyz4, st));
    DPM_TRY(lengths_to_i32_launch(lengths, B, N, len32, st));
    const bool brute = N <= FPS_BRUTE_CLUSTER_MAX_N && fps_cluster_mode_small(B);
    if (N >= fps_grid_min_n() && N <= GRID_MAX_N && !brute) {
        GridWs g;
        if (!grid_ws_carve(a, B, N, &g)) return fail(DPM_ERR_WORKSPACE, "fps: workspace too small");
        DPM_TRY(grid_build_launch(xyz4, B, N, len32, 0.f, g, st));
        DPM_TRY(fps_grid_launch(g, xyz4, B, N, K, idx_out, nullptr, nullptr, nullptr, nullptr, st));
    } else {
        DPM_TRY(fps_launch(xyz4, B, N, len32, K, idx_out, nullptr, nullptr, nullptr, nullptr, st));
    }
    if (sampled_out) {
        long long total = (long long)B * K * D;
        gather_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(points, idx_out, B, N, K, D, sampled_out);
        DPM_CHECK_LAUNCH("gather_rows", st);
    }
    return DPM_OK;
}
