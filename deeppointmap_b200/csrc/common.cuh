// common.cuh -- shared helpers for libdpm_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/dpm_b200.h"

namespace dpm {

// ---- error plumbing (thread-local, see dpm_last_error) --------------------------------
char *err_buf();
int fail(int code, const char *fmt, ...);
// counts the launch; when a profile is open (dpm_prof_begin) also records a CUDA event on `st`
// so consecutive events bracket each kernel (everything of one call is on one stream).
void count_launch(const char *tag, cudaStream_t st);
void prof_note(long long a, long long b);
bool prof_active();
void prof_mark(cudaStream_t st);  // call boundary: time since the previous record is not kernel time  // detail columns of the NEXT launch record

#define DPM_CHECK_CUDA(expr)                                                                   \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return dpm::fail(DPM_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                             __FILE__, __LINE__);                                              \
    } while (0)

#define DPM_CHECK_LAUNCH(tag, st)                                                              \
    do {                                                                                       \
        dpm::count_launch(tag, st);                                                            \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess)                                                                 \
            return dpm::fail(DPM_ERR_CUDA, "kernel launch failed: %s (%s:%d)",                 \
                             cudaGetErrorString(_e), __FILE__, __LINE__);                      \
    } while (0)

#define DPM_TRY(expr)                \
    do {                             \
        int _rc = (expr);            \
        if (_rc != DPM_OK) return _rc; \
    } while (0)

// ---- bump allocator over the caller's workspace ---------------------------------------
struct Arena {
    char *base;
    size_t cap, off;
    bool dry;  // dry run: only measure
    Arena(void *p, size_t n) : base((char *)p), cap(n), off(0), dry(p == nullptr) {}
    template <typename T>
    T *get(size_t count) {
        size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
        size_t o = off;
        off += bytes;
        if (dry) return nullptr;
        if (off > cap) return nullptr;
        return (T *)(base + o);
    }
    bool ok() const { return dry || off <= cap; }
};

int device_sm_count();
int current_device();

// ---- device helpers ---------------------------------------------------------------------
#ifdef __CUDACC__
// Exact reference arithmetic: (dx*dx + dy*dy) + dz*dz, fp32, round-to-nearest, no FMA
// contraction (network/encoder/utils.py:255-256).
__device__ __forceinline__ float d2_exact(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// in-place bitonic sort (descending) of n = 2^m 64-bit keys in shared memory by all threads of the block
__device__ __forceinline__ void bitonic_sort_desc(unsigned long long *s, int n, int tid, int nthreads) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < (n >> 1); i += nthreads) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const unsigned long long a = s[lo], b = s[hi];
                if (desc ? (a < b) : (a > b)) { s[lo] = b; s[hi] = a; }
            }
            __syncthreads();
        }
    }
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif

// ---- per-cloud uniform cell grid (grid.cu) ------------------------------------------------
constexpr int GRID_MAXCELL = 32768;   // cells per cloud (the cell size grows until the grid fits)
constexpr int GRID_MAX_N = 1048576;          // 1024 buckets x 32 lanes x 32 points per lane
constexpr int GRID_CLUSTER_MAX_N = 262144;   // the cluster / tuned one-SM FPS kernels: <= 8 points per lane
constexpr int GRID_MIN_N = 2048;      // below this the brute-force / register kernels win
struct GridDesc {
    float ox, oy, oz, inv_h;  // cell of p = floor((p - o) * inv_h), x fastest
    int gx, gy, gz, ncell;
    float h;
    int nvalid;               // points in the grid (= lengths[b])
    int pad0, pad1;
};
struct GridWs {
    float4 *sorted;   // (B, npad) cell-sorted points, w = original index bits; tail = sentinels
    int *cell_start;  // (B, GRID_MAXCELL + 1)
    int *cursor;      // (B, GRID_MAXCELL + 1) scatter cursors
    int *cellid;      // (B, N)
    float *mind;      // (B, npad) FPS running min-distances, sorted order
    GridDesc *desc;   // (B)
    int npad;
};
inline int grid_npad(int N) { return (N + 255) & ~255; }
inline int grid_ppl(int N) { return N <= 32768 ? 1 : (N <= 65536 ? 2 : (N <= 131072 ? 4 : (N <= 262144 ? 8 : (N <= 524288 ? 16 : 32)))); }
size_t grid_ws_bytes(int B, int N);
bool grid_ws_carve(Arena &a, int B, int N, GridWs *g);
// hmin: lower bound of the cell size (1.001 x the largest query radius served by this grid; 0 = FPS only)
int grid_build_launch(const float4 *xyz4, int B, int N, const int *len32, float hmin, const GridWs &g, cudaStream_t st);
int fps_grid_launch(const GridWs &g, const float4 *xyz4, int B, int N, int K, int64_t *idx64, int32_t *idx32,
                    float4 *new_xyz4, uint8_t *new_pad, int *new_len32, cudaStream_t st);
// latency variants (fps_cluster.cu): one cloud per 8-CTA cluster, chosen by fps_cluster_mode(B)
constexpr int FPS_BRUTE_CLUSTER_MAX_N = 8192;  // 32 warps x 32 lanes x 8 points in registers
bool fps_cluster_mode(int B);
bool fps_packed_mode();  // dpm_set_fps_mode(3): two clouds per SM in the one-SM grid kernel
bool fps_cluster_mode_small(int B);  // for clouds of <= FPS_BRUTE_CLUSTER_MAX_N points
// smallest cloud that takes the pruned (grid) FPS; smaller ones stay in registers (fps.cu)
int fps_grid_min_n();
int fps_grid_cluster_launch(const GridWs &g, const float4 *xyz4, int B, int N, int K, int64_t *idx64, int32_t *idx32,
                            float4 *new_xyz4, uint8_t *new_pad, int *new_len32, cudaStream_t st);
int fps_grid_onesm_launch(const GridWs &g, const float4 *xyz4, int B, int N, int K, int64_t *idx64, int32_t *idx32,
                          float4 *new_xyz4, uint8_t *new_pad, int *new_len32, cudaStream_t st);
int fps_brute_cluster_launch(const float4 *xyz4, int B, int N, const int *len32, int K, int64_t *idx64, int32_t *idx32,
                             float4 *new_xyz4, uint8_t *new_pad, int *new_len32, cudaStream_t st);
int knn_grid_launch(const GridWs &g, const float4 *q4, const float4 *p4, int B, int S, int N, const int *qlen32,
                    int K, float r2, int64_t *idx64, int32_t *idx32, cudaStream_t st, bool pad = false);

// exact uncapped kNN (knn_points contract) over the grid: 3x3x3 block, then shells until provably complete
int knn_ring_launch(const GridWs &g, const float4 *q4, int B, int S, const int *qlen32, int K, int64_t *idx64,
                    int32_t *idx32, float *d2out, cudaStream_t st);

// per-point surface normal from the neighbours within `radius` (grid cell size must be >= radius; g == nullptr: brute force)
int radius_normals_launch(const GridWs *g, const float4 *q4, int S, float radius, float *normals, cudaStream_t st);

// ---- internal launchers shared between translation units ------------------------------
// xyz4 buffers are float4 (x,y,z,0) rows.
int fps_launch(const float4 *xyz4, int B, int N, const int *len32, int K, int64_t *idx64,
               int32_t *idx32, float4 *new_xyz4, uint8_t *new_pad, int *new_len32, cudaStream_t st);
int knn_launch(const float4 *q4, const float4 *p4, int B, int S, int N, const int *qlen32,
               const int *plen32, int K, float r2, int mode, int64_t *idx64, int32_t *idx32,
               float *d2out, cudaStream_t st);
enum { KNN_MODE_KNN = 0, KNN_MODE_HYBRID = 1 };
int pack_xyz4_launch(const float *src, int B, int N, int D, float4 *dst, cudaStream_t st);
int lengths_to_i32_launch(const int64_t *len64, int B, int N, int *len32, cudaStream_t st);
int linear_launch(const float *X, int ldx, const float *W, int ldw, const float *bias, const float *res, int ldres,
                  float *Y, int ldy, int M, int N, int K, int act, cudaStream_t st);
// tcgen05 3xTF32 path (gemm_tc.cu); linear_launch routes to it when eligible
bool linear_tc_eligible(const float *X, int ldx, long long sX, const float *W, int ldw, long long sW, int M, int N, int K);
int linear_tc_launch(const float *X, int ldx, long long sX, const float *W, int ldw, long long sW, const float *bias,
                     const float *res, int ldres, float *Y, int ldy, long long sY, int M, int N, int K, int nbatch,
                     int act, cudaStream_t st);
// per-call pre-split (hi/lo tf32) copies of the weights for the tensor-core path (gemm_tc.cu)
size_t split_floats(int rows, int cols);  // floats of the hi + lo copies of a rows x cols weight
void split_begin();
void split_add(Arena &a, const float *W, int rows, int cols, int ld);
int split_run(cudaStream_t st);
const float *split_lookup(const float *W, int rows, int cols, int ld);
int linear_batched_launch(const float *X, int ldx, long long sX, const float *W, int ldw, long long sW,
                          const float *bias, const float *res, int ldres, float *Y, int ldy, long long sY, int M,
                          int N, int K, int nbatch, int act, cudaStream_t st);
bool linear_ln_tc_launch(const float *X, int ldx, const float *W, int ldw, const float *bias, const float *res,
                         int ldres, const float *gamma, const float *beta, const float *post, int ldpost, float *Y,
                         int ldy, int M, int N, int K, int act, cudaStream_t st, int *rc, float *tmp = nullptr);
// Y = act(LayerNorm_N(X W^T + bias + res) * gamma + beta + post): fused when eligible, else linear + layernorm via
// the scratch buffer `tmp` (M x N, may be Y itself when Y does not alias res / post)
int linear_ln_launch(const float *X, int ldx, const float *W, int ldw, const float *bias, const float *res, int ldres,
                     const float *gamma, const float *beta, const float *post, int ldpost, float *tmp, float *Y, int ldy,
                     int M, int N, int K, int act, cudaStream_t st, int tmp_copies = 1);
int layernorm_launch(const float *X, int ldx, const float *gamma, const float *beta, const float *post, int ldpost,
                     float *Y, int ldy, int M, int C, int act, cudaStream_t st);
int group_launch(const float *Z, const float4 *xyz4, const float4 *ctr4, const int32_t *gidx, const float *Wxyz,
                 int ldw, const float *gamma, const float *beta, float radius, float *out, int B, int N, int S,
                 int K, int Cout, cudaStream_t st);
bool group_from_xyz_supported(int Cout);
// stage-0 SA when the gathered features are the stem conv of the coordinates: no per-point matrix at all
int group_from_xyz_launch(const float *Wsa, int ldw, const float *bias, const float *W0, const float *b0, int width,
                          float4 *comp, const float4 *xyz4, const float4 *ctr4, const int32_t *gidx, const float *gamma,
                          const float *beta, float radius, float *out, int B, int N, int S, int K, int Cout,
                          cudaStream_t st);
// attention_tc5.cu: tcgen05 flash attention for long key ranges (pairs layout of decoder.cu)
int attention_tc5_launch(const float *q, int ldq, const float *k, int ldk, const float *v, int ldv, float *out, int ldo,
                         int P, int M, int N, int mode, int heads, const uint8_t *kmask, cudaStream_t st);
// pairing.cu: dual softmax + global top-k of the similarity matrices, three launches
constexpr int PAIR_MAXK = 4096;
size_t pairing_ws_bytes(int P);
int pairing_launch(float *S, int P, int M, int N, float tau, int k, float2 *rs, float2 *cs, void *ws, int32_t *si,
                   int32_t *di, float *conf, cudaStream_t st);
int fp_interp_launch(const float4 *xyz1, const float4 *xyz2, const float *fea1, const float *fea2,
                     const uint8_t *pad2, float *out, int B, int N, int S, int C1, int C2, cudaStream_t st);

}  // namespace dpm
