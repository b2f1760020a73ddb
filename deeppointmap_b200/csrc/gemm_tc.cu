// gemm_tc.cu -- Y = act(X W^T + bias + res) on the 5th-generation tensor cores (tcgen05),
// fp32 in / fp32 out with ERROR-COMPENSATED TF32 ("3xTF32"): every operand is split on the fly
// into hi = tf32(x) and lo = tf32(x - hi) and the product is accumulated as
// hi*hi + hi*lo + lo*hi in the fp32 TMEM accumulator, which keeps the result within ~1e-6 of
// the fp32 reference (plain TF32 / BF16 would break the 1e-4 parity bar of the decoder).
//
// Replaces the 1x1 Conv1d / Conv2d / Linear layers of build_mlp
// (network/encoder/utils.py:358-389), nn.MultiheadAttention's in/out projections and the
// decoder heads (network/decoder/heads.py) for M >= 128 rows.
//
// One CTA = one 128 x BN output tile.
//   warps 0-7 : two producer groups of 4 warps that take alternate 32-wide K blocks (so the global
//               load latency of one block overlaps the split + store of the other) -- coalesced
//               16-byte global loads of the X (128 x 32) and W (BN x 32) fp32 tiles, hi/lo split in
//               registers, st.shared into the canonical K-major SWIZZLE_128B layout the UMMA
//               descriptors describe; then the epilogue -- tcgen05.ld of the accumulator (one output
//               row per thread, each group half of the columns), bias / residual / ReLU, 16-byte stores.
//   warp 8    : allocates TMEM; one elected lane issues 12 tcgen05.mma.kind::tf32 (128 x BN x 8) per
//               K block and releases the stage with tcgen05.commit -> mbarrier.
// The split is why the operands are staged by the producer warps instead of by TMA: TMA cannot
// transform data in flight.
#include <stdlib.h>

#include "common.cuh"
#include "tc5.cuh"

namespace dpm {

namespace tc {

constexpr int BM = 128;
constexpr int BK = 32;            // fp32 elements per K block = one 128-byte swizzle row
constexpr int THREADS = 288;      // 8 producer/epilogue warps + 1 MMA warp
constexpr int A_TILE = BM * 128;  // bytes of one 128 x 32 fp32 tile

template <int ROWS>
__device__ __forceinline__ void produce_tile(const float *__restrict__ G, int ld, int rows_total, int r0, int K, int k0,
                                             unsigned char *hi, unsigned char *lo, int tid) {
    constexpr int ITER = ROWS * 8 / 128;
    float4 v[ITER];
#pragma unroll
    for (int i = 0; i < ITER; ++i) {
        const int idx = tid + i * 128, r = idx >> 3, c = idx & 7;
        const int gr = r0 + r, gk = k0 + c * 4;
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gr < rows_total && gk < K) v[i] = __ldg(reinterpret_cast<const float4 *>(G + (size_t)gr * ld + gk));
    }
#pragma unroll
    for (int i = 0; i < ITER; ++i) {
        const int idx = tid + i * 128, r = idx >> 3, c = idx & 7;
        float4 h, l;  // Veltkamp split on the FP32 pipe (cvt.rna.tf32 is a quarter-rate conversion), see tc5.cuh
        split_tf32(v[i].x, h.x, l.x); split_tf32(v[i].y, h.y, l.y); split_tf32(v[i].z, h.z, l.z); split_tf32(v[i].w, h.w, l.w);
        const unsigned off = swz(r, c);
        *reinterpret_cast<float4 *>(hi + off) = h;
        *reinterpret_cast<float4 *>(lo + off) = l;
    }
}

// ---- cluster plumbing for the pre-split weight tiles --------------------------------------------------
__device__ __forceinline__ unsigned cluster_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cluster_size() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> this CTA's shared memory, completing `bytes` on the CTA's own mbarrier
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// ... and into the same offsets of EVERY CTA of the cluster in `mask` (each one's mbarrier gets the bytes)
__device__ __forceinline__ void bulk_g2s_multicast(unsigned dst, const void *src, unsigned bytes, unsigned bar,
                                                   unsigned short mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar), "h"(mask)
        : "memory");
}
// commit -> one arrival on the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void mma_commit_multicast(unsigned bar, unsigned short mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}

__device__ __forceinline__ void cp_async16(unsigned dst, const void *src, unsigned src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// W tiles of PRE-SPLIT weights (split_weights_kernel, once per call): hi and lo are stored in global memory
// K-block-major and already swizzled, [K block][row n][32 floats, 16-byte chunk c at (c ^ n) & 7], so the BN x 32 tile
// of K block kb is ONE contiguous run of BN x 128 bytes that lands in the canonical SWIZZLE_128B layout by a plain 1-D
// bulk copy (TMA engine, no registers, no ALU).  CTAs with consecutive row tiles form a CLUSTER along M: each one
// fetches 1 / cluster of the tile's rows and MULTICASTS them into every CTA of the cluster, so a weight tile crosses
// L2 -> SM once per cluster instead of once per CTA (the 64 MB of redundant weight reads of a 16 384 x 256 x 256
// layer were what bounded its main loop: 12 us against 6.4 us of MMAs).
#ifdef DPM_TC_PROFILE
// developer build: wall-clock (globaltimer, ns) of the phases of every CTA of the last launch
__device__ unsigned long long tc_prof[2048][6];
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define TC_STAMP(i)                                                                                       \
    do {                                                                                                  \
        const int _c = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);                    \
        if (_c < 2048) tc_prof[_c][i] = gtimer();                                                         \
    } while (0)
#else
#define TC_STAMP(i) do { } while (0)
#endif

// the two warps that share a TMEM lane quarter (immediate barrier ids: a register id makes ptxas reserve all 16)
__device__ __forceinline__ void pair_barrier(int quarter) {
    if (quarter == 0) asm volatile("bar.sync 1, 64;" ::: "memory");
    else if (quarter == 1) asm volatile("bar.sync 2, 64;" ::: "memory");
    else if (quarter == 2) asm volatile("bar.sync 3, 64;" ::: "memory");
    else asm volatile("bar.sync 4, 64;" ::: "memory");
}

// LayerNorm arithmetic shared by the fused epilogue and ln_fused_order_kernel (the few-rows path): written with explicit
// round-to-nearest intrinsics so that both kernels contract nothing differently -- the same row gives the same bits
// whichever path its batch size selects.
__device__ __forceinline__ float ln_sq4(const float4 o, float mean) {
    const float dx = __fsub_rn(o.x, mean), dy = __fsub_rn(o.y, mean), dz = __fsub_rn(o.z, mean), dw = __fsub_rn(o.w, mean);
    return __fadd_rn(__fmaf_rn(dx, dx, __fmul_rn(dy, dy)), __fmaf_rn(dz, dz, __fmul_rn(dw, dw)));
}
__device__ __forceinline__ float ln_sum4(const float4 o) { return __fadd_rn(__fadd_rn(o.x, o.y), __fadd_rn(o.z, o.w)); }
__device__ __forceinline__ float ln_out(float o, float mean, float rstd, float g, float b, float p) {
    return __fadd_rn(__fmaf_rn(__fmul_rn(__fsub_rn(o, mean), rstd), g, b), p);
}
__device__ __forceinline__ float ln_mean(float s0, float s1, int N) { return __fdiv_rn(__fadd_rn(s0, s1), (float)N); }
__device__ __forceinline__ float ln_rstd(float q0, float q1, int N) {
    return __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fdiv_rn(__fadd_rn(q0, q1), (float)N), 1e-5f)));
}

struct LnArgs {  // fused LayerNorm epilogue: Y = act(LN_N(acc + bias + res) * gamma + beta + post)
    const float *gamma, *beta, *post;
    int ldpost;
};

template <int BN, int STAGES, bool PRESPLIT, bool LN, int MINB = 1>
__global__ void __launch_bounds__(THREADS, MINB)
linear_tc_kernel(const float *__restrict__ X, int ldx, const float *__restrict__ W, int ldw,
                 const float *__restrict__ bias, const float *res, int ldres, float *Y,
                 int ldy, int M, int N, int K, int act, long long sX, long long sW, long long sY, long long wlo_off,
                 LnArgs ln) {
    X += (size_t)blockIdx.z * sX;
    W += (size_t)blockIdx.z * sW;
    Y += (size_t)blockIdx.z * sY;
    if (res) res += (size_t)blockIdx.z * sY;
    constexpr int B_TILE = BN * 128;
    constexpr int STAGE_BYTES = 2 * A_TILE + 2 * B_TILE;
    // 1024-byte aligned by declaration (SWIZZLE_128B atoms); used directly so that the compiler keeps the shared
    // address space (an integer round trip to align it by hand turned every staging store into a generic ST)
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[2 * STAGES + 1];
    __shared__ unsigned tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int nkb = (K + BK - 1) / BK;
    const unsigned full0 = s32(&bars[0]), empty0 = s32(&bars[STAGES]), accum = s32(&bars[2 * STAGES]);

    const unsigned csize = PRESPLIT ? cluster_size() : 1u, crank = PRESPLIT ? cluster_rank() : 0u;
    const unsigned short cmask = (unsigned short)((1u << csize) - 1u);
    if (tid == 0) TC_STAMP(0);
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, PRESPLIT ? 129 : 128);  // X producers (+ the expect_tx arrival of the W copies)
            mbar_init(empty0 + 8 * s, csize);                // the MMA warp of every CTA that receives this stage's W
        }
        mbar_init(accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // two accumulators: [0, BN) takes hi*hi, [BN, 2BN) the two cross terms.  The tensor core's fp32
    // accumulation truncates, and that bias grows with the number of additions into a LARGE accumulator;
    // the cross terms are 2^-11 of the main term, so parking them in their own accumulator cuts the
    // additions into the main one by 3x (they are summed with IEEE adds in the epilogue).
    constexpr unsigned TCOLS = 2 * BN < 32 ? 32 : 2 * BN;
    if (warp == 8) tmem_alloc(s32(&tmem_base_s), TCOLS);
    tc_fence_before();
    __syncthreads();
    if (csize > 1) cluster_sync_all();  // nobody multicasts into a CTA whose mbarriers are not initialised yet
    tc_fence_after();
    const unsigned tmem_d = tmem_base_s;
    if (tid == 0) TC_STAMP(1);

    if (warp < 8) {
        // ================= producers: group g takes the K blocks kb = g, g + 2, ... =================
        const int grp = warp >> 2, ptid = tid & 127;
        for (int kb = grp; kb < nkb; kb += 2) {
            const int s = kb % STAGES;
            mbar_wait(empty0 + 8 * s, (unsigned)(((kb / STAGES) & 1) ^ 1));
            unsigned char *st = smem + (size_t)s * STAGE_BYTES;
            if (PRESPLIT) {
                if (ptid == 0) {  // every CTA of the cluster has released this stage (empty barrier): send my rows
                    const int rows = min(BN, N - n0);
                    mbar_expect_tx(full0 + 8 * s, 2u * (unsigned)rows * 128u);
                    const int per = (((rows + (int)csize - 1) / (int)csize) + 7) & ~7;
                    const int r0 = (int)crank * per, nr = min(per, rows - r0);
                    if (nr > 0) {
                        const float *src = W + ((size_t)kb * N + n0 + r0) * BK;
                        const unsigned dh = s32(st + 2 * A_TILE) + (unsigned)r0 * 128u, dl = dh + B_TILE;
                        if (csize > 1) {
                            bulk_g2s_multicast(dh, src, (unsigned)nr * 128u, full0 + 8 * s, cmask);
                            bulk_g2s_multicast(dl, src + wlo_off, (unsigned)nr * 128u, full0 + 8 * s, cmask);
                        } else {
                            bulk_g2s(dh, src, (unsigned)nr * 128u, full0 + 8 * s);
                            bulk_g2s(dl, src + wlo_off, (unsigned)nr * 128u, full0 + 8 * s);
                        }
                    }
                }
            }
            produce_tile<BM>(X, ldx, M, m0, K, kb * BK, st, st + A_TILE, ptid);
            if (!PRESPLIT) produce_tile<BN>(W, ldw, N, n0, K, kb * BK, st + 2 * A_TILE, st + 2 * A_TILE + B_TILE, ptid);
            fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
            mbar_arrive(full0 + 8 * s);
        }
        // ================= epilogue =================
        // tcgen05.ld hands thread t of a warp row (lane quarter base + t); writing rows from there would
        // touch 32 different 128-byte lines per instruction (and so would the residual read).  The 32x32
        // chunk goes through shared memory instead (the pipeline stages are idle by now) and leaves /
        // meets the residual as whole 128-byte row segments, 4 rows per warp instruction.
        mbar_wait(accum, 0u);
        tc_fence_after();
        if (tid == 0) TC_STAMP(3);
        const int quarter = warp & 3;  // the TMEM lane quarter this warp may read
        constexpr int CHALF = BN >= 64 ? BN / 2 : BN;  // columns per producer group in the epilogue
        constexpr int TLD = 36;                        // staging row stride (floats): 16-byte aligned, conflict-free
        float *tb = reinterpret_cast<float *>(smem) + warp * (32 * TLD);
        const bool vec = ((ldy & 3) == 0) && ((((uintptr_t)Y) & 15) == 0) && (!bias || ((((uintptr_t)bias) & 15) == 0)) &&
                         (!res || (((ldres & 3) == 0) && ((((uintptr_t)res) & 15) == 0)));
        const int rsub = lane >> 3, c4 = (lane & 7) * 4;  // this lane's (row within a group of 4, column) when storing
        if (LN) {
            // The whole row lives in this CTA (N <= BN, one column tile).  The 128 x N tile of acc + bias + res is
            // parked in shared memory (row stride BN + 4), the two warps that share a TMEM lane quarter (one per
            // column half) exchange their partial row sums through shared memory, and the normalised rows leave
            // as whole 128-byte segments.  Two-pass variance, like the stand-alone LayerNorm kernel.
            constexpr int TS = BN + 4;
            float *tile = reinterpret_cast<float *>(smem);
            float *stat = tile + 128 * TS;  // [2 passes][2 halves][128 rows]
            const int cbeg = grp * CHALF, cend = (BN >= 64 ? (grp + 1) * CHALF : (grp == 0 ? BN : 0));
            float sum[8], mean[8], rstd[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) sum[i] = 0.f;
#pragma unroll 1
            for (int cb = cbeg; cb < cend; cb += 32) {
                if (cb >= N) break;  // warp-uniform
                float v[32], vc[32];
                tmem_ld32(tmem_d + ((unsigned)(quarter * 32) << 16) + (unsigned)cb, v);
                tmem_ld32(tmem_d + ((unsigned)(quarter * 32) << 16) + (unsigned)(BN + cb), vc);
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    *reinterpret_cast<float4 *>(tile + (quarter * 32 + lane) * TS + cb + 4 * q) =
                        make_float4(v[4 * q] + vc[4 * q], v[4 * q + 1] + vc[4 * q + 1], v[4 * q + 2] + vc[4 * q + 2],
                                    v[4 * q + 3] + vc[4 * q + 3]);
                __syncwarp();
                const int col = cb + c4;
                const bool cok = col < N;  // N % 4 == 0
                float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
                if (bias && cok) bb = __ldg(reinterpret_cast<const float4 *>(bias + col));
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int lr = quarter * 32 + i * 4 + rsub, row = m0 + lr;
                    float4 o = *reinterpret_cast<const float4 *>(tile + lr * TS + col);
                    float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (res && cok && row < M) r4 = *reinterpret_cast<const float4 *>(res + (size_t)row * ldres + col);
                    o.x += bb.x + r4.x; o.y += bb.y + r4.y; o.z += bb.z + r4.z; o.w += bb.w + r4.w;
                    if (!cok) o = make_float4(0.f, 0.f, 0.f, 0.f);
                    *reinterpret_cast<float4 *>(tile + lr * TS + col) = o;
                    sum[i] = __fadd_rn(sum[i], ln_sum4(o));
                }
            }
            // ---- mean ----
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                sum[i] += __shfl_xor_sync(0xffffffffu, sum[i], 1);
                sum[i] += __shfl_xor_sync(0xffffffffu, sum[i], 2);
                sum[i] += __shfl_xor_sync(0xffffffffu, sum[i], 4);
                if ((lane & 7) == 0) stat[grp * 128 + quarter * 32 + i * 4 + rsub] = sum[i];
            }
            pair_barrier(quarter);  // the two warps of this lane quarter
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int lr = quarter * 32 + i * 4 + rsub;
                mean[i] = ln_mean(stat[lr], stat[128 + lr], N);
                sum[i] = 0.f;
            }
            // ---- variance ----
#pragma unroll 1
            for (int cb = cbeg; cb < cend; cb += 32) {
                if (cb >= N) break;
                const int col = cb + c4;
                if (col < N) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 o = *reinterpret_cast<const float4 *>(tile + (quarter * 32 + i * 4 + rsub) * TS + col);
                        sum[i] = __fadd_rn(sum[i], ln_sq4(o, mean[i]));
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                sum[i] += __shfl_xor_sync(0xffffffffu, sum[i], 1);
                sum[i] += __shfl_xor_sync(0xffffffffu, sum[i], 2);
                sum[i] += __shfl_xor_sync(0xffffffffu, sum[i], 4);
                if ((lane & 7) == 0) stat[256 + grp * 128 + quarter * 32 + i * 4 + rsub] = sum[i];
            }
            pair_barrier(quarter);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int lr = quarter * 32 + i * 4 + rsub;
                rstd[i] = ln_rstd(stat[256 + lr], stat[384 + lr], N);
            }
            // ---- normalise, affine, post-add, activation, store ----
#pragma unroll 1
            for (int cb = cbeg; cb < cend; cb += 32) {
                if (cb >= N) break;
                const int col = cb + c4;
                if (col >= N) continue;
                const float4 g4 = __ldg(reinterpret_cast<const float4 *>(ln.gamma + col));
                const float4 b4 = __ldg(reinterpret_cast<const float4 *>(ln.beta + col));
                float4 p4[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = m0 + quarter * 32 + i * 4 + rsub;
                    p4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ln.post && row < M) p4[i] = *reinterpret_cast<const float4 *>(ln.post + (size_t)row * ln.ldpost + col);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int lr = quarter * 32 + i * 4 + rsub, row = m0 + lr;
                    float4 o = *reinterpret_cast<const float4 *>(tile + lr * TS + col);
                    o.x = ln_out(o.x, mean[i], rstd[i], g4.x, b4.x, p4[i].x);
                    o.y = ln_out(o.y, mean[i], rstd[i], g4.y, b4.y, p4[i].y);
                    o.z = ln_out(o.z, mean[i], rstd[i], g4.z, b4.z, p4[i].z);
                    o.w = ln_out(o.w, mean[i], rstd[i], g4.w, b4.w, p4[i].w);
                    if (act == DPM_ACT_RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    if (row < M) *reinterpret_cast<float4 *>(Y + (size_t)row * ldy + col) = o;
                }
            }
        } else
#pragma unroll 1
        for (int cb = grp * CHALF; cb < (BN >= 64 ? (grp + 1) * CHALF : (grp == 0 ? BN : 0)); cb += 32) {
            if (n0 + cb >= N) break;  // warp-uniform
            float v[32], vc[32];
            tmem_ld32(tmem_d + ((unsigned)(quarter * 32) << 16) + (unsigned)cb, v);
            tmem_ld32(tmem_d + ((unsigned)(quarter * 32) << 16) + (unsigned)(BN + cb), vc);
#pragma unroll
            for (int q = 0; q < 8; ++q)
                *reinterpret_cast<float4 *>(tb + lane * TLD + 4 * q) =
                    make_float4(v[4 * q] + vc[4 * q], v[4 * q + 1] + vc[4 * q + 1], v[4 * q + 2] + vc[4 * q + 2],
                                v[4 * q + 3] + vc[4 * q + 3]);
            __syncwarp();
            const int col = n0 + cb + c4;
            if (vec && n0 + cb + 32 <= N) {
                float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
                if (bias) bb = __ldg(reinterpret_cast<const float4 *>(bias + col));
                float4 r4[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = m0 + quarter * 32 + i * 4 + rsub;
                    r4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (res && row < M) r4[i] = *reinterpret_cast<const float4 *>(res + (size_t)row * ldres + col);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = m0 + quarter * 32 + i * 4 + rsub;
                    float4 o = *reinterpret_cast<const float4 *>(tb + (i * 4 + rsub) * TLD + c4);
                    o.x += bb.x + r4[i].x; o.y += bb.y + r4[i].y; o.z += bb.z + r4[i].z; o.w += bb.w + r4[i].w;
                    if (act == DPM_ACT_RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    if (row < M) *reinterpret_cast<float4 *>(Y + (size_t)row * ldy + col) = o;
                }
            } else {
#pragma unroll 4
                for (int r = 0; r < 32; ++r) {
                    const int row = m0 + quarter * 32 + r, cc = n0 + cb + lane;
                    if (row < M && cc < N) {
                        float o = tb[r * TLD + lane];
                        float add = bias ? bias[cc] : 0.f;  // acc + (bias + res): the association of the vector path
                        if (res) add += res[(size_t)row * ldres + cc];
                        o += add;
                        if (act == DPM_ACT_RELU) o = fmaxf(o, 0.f);
                        Y[(size_t)row * ldy + cc] = o;
                    }
                }
            }
            __syncwarp();
        }
    } else if (lane == 0) {
        // ================= MMA issuer: one thread =================
        constexpr unsigned IDESC = instr_desc(BN < 16 ? 16 : BN);
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % STAGES;
            mbar_wait(full0 + 8 * s, (unsigned)((kb / STAGES) & 1));
            tc_fence_after();
            if (kb == 0) TC_STAMP(2);
            const unsigned sa = s32(smem + (size_t)s * STAGE_BYTES);
            const unsigned long long ah = smem_desc(sa), al = smem_desc(sa + A_TILE);
            const unsigned long long bh = smem_desc(sa + 2 * A_TILE), bl = smem_desc(sa + 2 * A_TILE + B_TILE);
#pragma unroll
            for (int j = 0; j < BK / 8; ++j) {
                const unsigned long long ko = (unsigned long long)(j * 2);  // 32 bytes >> 4 along K inside the swizzle row
                mma_tf32(tmem_d, ah + ko, bh + ko, IDESC, (kb | j) != 0 ? 1u : 0u);
                mma_tf32(tmem_d + BN, ah + ko, bl + ko, IDESC, (kb | j) != 0 ? 1u : 0u);
                mma_tf32(tmem_d + BN, al + ko, bh + ko, IDESC, 1u);
            }
            // stage free once these MMAs have read it -- for every CTA that multicasts into it
            if (csize > 1) mma_commit_multicast(empty0 + 8 * s, cmask);
            else mma_commit(empty0 + 8 * s);
        }
        mma_commit(accum);
    }
    if (tid == 0) TC_STAMP(4);
    tc_fence_before();
    __syncthreads();
    if (csize > 1) cluster_sync_all();  // no CTA leaves while a peer's commits / copies can still target it
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem_d, TCOLS);
    }
    if (tid == 0) TC_STAMP(5);
}

template <int BN, int STAGES, bool PRESPLIT, bool LN, int MINB>
static int launch_s(const float *X, int ldx, long long sX, const float *W, int ldw, long long sW, const float *bias,
                    const float *res, int ldres, float *Y, int ldy, long long sY, int M, int N, int K, int nbatch,
                    int act, long long wlo_off, cudaStream_t st, LnArgs ln);

// SHALLOW contractions (K <= 64 / 128: the early encoder stages, 1024 row tiles per launch): the default stage count
// would hold 160-200 KB of shared memory for a K loop of 1-4 blocks and pin one CTA per SM, whose 2.8 us fill and
// 5-6 us epilogue then run with the tensor core idle.  With as many stages as K blocks (<= 2) two CTAs share an SM
// (<= 113 KB, <= 256 TMEM columns each) and one's epilogue overlaps the other's loads.
template <int BN, int STAGES, bool PRESPLIT, bool LN = false>
static int launch_t(const float *X, int ldx, long long sX, const float *W, int ldw, long long sW, const float *bias,
                    const float *res, int ldres, float *Y, int ldy, long long sY, int M, int N, int K, int nbatch,
                    int act, long long wlo_off, cudaStream_t st, LnArgs ln = LnArgs{nullptr, nullptr, nullptr, 0}) {
    static const bool shallow_off = getenv("DPM_TC_NO_SHALLOW") != nullptr;
    const int nkb = (K + BK - 1) / BK, mt = (M + BM - 1) / BM;
    if constexpr (PRESPLIT && BN <= 128) {
        if (!shallow_off && mt > 148) {
            if constexpr (BN <= 64) {
                if (nkb <= 8 && nkb >= 2)
                    return launch_s<BN, 2, PRESPLIT, LN, 2>(X, ldx, sX, W, ldw, sW, bias, res, ldres, Y, ldy, sY, M, N, K, nbatch, act,
                                                            wlo_off, st, ln);
            }
            if (nkb == 1)
                return launch_s<BN, 1, PRESPLIT, LN, 2>(X, ldx, sX, W, ldw, sW, bias, res, ldres, Y, ldy, sY, M, N, K, nbatch, act,
                                                        wlo_off, st, ln);
        }
    }
    return launch_s<BN, STAGES, PRESPLIT, LN, 1>(X, ldx, sX, W, ldw, sW, bias, res, ldres, Y, ldy, sY, M, N, K, nbatch, act, wlo_off,
                                                 st, ln);
}

template <int BN, int STAGES, bool PRESPLIT, bool LN, int MINB>
static int launch_s(const float *X, int ldx, long long sX, const float *W, int ldw, long long sW, const float *bias,
                    const float *res, int ldres, float *Y, int ldy, long long sY, int M, int N, int K, int nbatch,
                    int act, long long wlo_off, cudaStream_t st, LnArgs ln) {
    auto kern = linear_tc_kernel<BN, STAGES, PRESPLIT, LN, MINB>;
    // the pipeline stages double as the epilogue's staging area: LayerNorm parks the whole 128 x (BN + 4) tile + statistics
    constexpr size_t stage_bytes = (size_t)STAGES * (2 * A_TILE + 2 * BN * 128);
    constexpr size_t epi_bytes = LN ? (size_t)(128 * (BN + 4) + 4 * 128) * 4 : (size_t)8 * 32 * 36 * 4;
    const size_t smem = stage_bytes > epi_bytes ? stage_bytes : epi_bytes;
    static thread_local unsigned long long configured = 0ull;  // one bit per device: function attributes are per context
    const unsigned long long devbit = 1ull << (current_device() & 63);
    if (!(configured & devbit)) {
        DPM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured |= devbit;
    }
    const int mt = (M + BM - 1) / BM;
    // pre-split weights: clusters of row tiles share every weight tile by multicast (the grid is padded to whole
    // clusters; a CTA past the last row tile loads zeros, stores nothing and keeps the cluster's pipeline in step)
    static const int cl_env = getenv("DPM_TC_CLUSTER") ? atoi(getenv("DPM_TC_CLUSTER")) : 0;
    int cl = 1;
    if (PRESPLIT) {
        // measured (round 2, 32-frame steps on 5 streams): cluster 1 -> 6613 frames/s, 2 -> 6474, 4 -> 6304.  The weight
        // re-reads were NOT what bounds the main loop (it runs at 76 % of the tf32 MMA peak either way); clusters only
        // add co-scheduling constraints and two cluster barriers.  So the default is one CTA per "cluster" (plain bulk
        // copies); DPM_TC_CLUSTER=2|4 keeps the multicast path measurable.
        cl = cl_env > 0 ? cl_env : 1;
        if (cl != 1 && cl != 2 && cl != 4 && cl != 8) cl = 1;
        if (cl > mt) cl = 1;
    }
    dim3 grid((mt + cl - 1) / cl * cl, (N + BN - 1) / BN, nbatch);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = cl > 1 ? 1 : 0;
    DPM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, X, ldx, W, ldw, bias, res, ldres, Y, ldy, M, N, K, act, sX, sW, sY, wlo_off, ln));
    count_launch(LN ? "linear_ln_tc" : "linear_tc", st);
    return DPM_OK;
}

// LayerNorm of rows that a NARROW-tile GEMM left in global memory (acc + bias + res already added), in exactly the
// summation order of the fused epilogue of a 128 x BN tile: 8 lanes per row, 4 columns per lane and 32-column chunk,
// the two column halves of the tile summed separately (they are two warps there) and then added.  Same bits as the
// fused kernel for the same row.
__global__ void __launch_bounds__(256)
ln_fused_order_kernel(const float *T, int ldt, const float *__restrict__ gamma, const float *__restrict__ beta,
                      const float *post, int ldpost, float *Y, int ldy, int M, int N, int BN, int act) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = (blockIdx.x * 8 + warp) * 4 + (lane >> 3), c4 = (lane & 7) * 4;
    const bool rok = row < M;
    const int CHALF = BN >= 64 ? BN / 2 : BN;
    const float *t = T + (size_t)(rok ? row : 0) * ldt;
    // the lane's 4 columns of every 32-column chunk (N <= 256: at most 8 chunks), all loads in flight together
    float4 o[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int col = c * 32 + c4;
        o[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rok && col < N) o[c] = *reinterpret_cast<const float4 *>(t + col);
    }
    float s[2] = {0.f, 0.f}, q[2] = {0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int cb = c * 32;
        if (cb < N && cb < BN) {                       // the fused loop: cb in [half begin, half end), break at cb >= N
            const int h = (BN >= 64 && cb >= CHALF) ? 1 : 0;
            s[h] = __fadd_rn(s[h], ln_sum4(o[c]));
        }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        s[h] += __shfl_xor_sync(0xffffffffu, s[h], 1);
        s[h] += __shfl_xor_sync(0xffffffffu, s[h], 2);
        s[h] += __shfl_xor_sync(0xffffffffu, s[h], 4);
    }
    const float mean = ln_mean(s[0], s[1], N);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int cb = c * 32;
        if (cb < N && cb < BN && cb + c4 < N) {
            const int h = (BN >= 64 && cb >= CHALF) ? 1 : 0;
            q[h] = __fadd_rn(q[h], ln_sq4(o[c], mean));
        }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        q[h] += __shfl_xor_sync(0xffffffffu, q[h], 1);
        q[h] += __shfl_xor_sync(0xffffffffu, q[h], 2);
        q[h] += __shfl_xor_sync(0xffffffffu, q[h], 4);
    }
    const float rstd = ln_rstd(q[0], q[1], N);
    if (!rok) return;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int col = c * 32 + c4;
        if (col >= N) continue;
        const float4 g4 = __ldg(reinterpret_cast<const float4 *>(gamma + col));
        const float4 b4 = __ldg(reinterpret_cast<const float4 *>(beta + col));
        float4 p4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (post) p4 = *reinterpret_cast<const float4 *>(post + (size_t)row * ldpost + col);
        float4 r;
        r.x = ln_out(o[c].x, mean, rstd, g4.x, b4.x, p4.x);
        r.y = ln_out(o[c].y, mean, rstd, g4.y, b4.y, p4.y);
        r.z = ln_out(o[c].z, mean, rstd, g4.z, b4.z, p4.z);
        r.w = ln_out(o[c].w, mean, rstd, g4.w, b4.w, p4.w);
        if (act == DPM_ACT_RELU) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
        *reinterpret_cast<float4 *>(Y + (size_t)row * ldy + col) = r;
    }
}

}  // namespace tc

// FEW ROWS (one frame per call: 512 decoder tokens, 16 ... 1024 points of the deeper encoder stages): a 128 x 256 tile
// leaves 1-8 CTAs on the chip, each walking all K blocks with 12 wide MMAs per block and, when LayerNorm rides in the
// epilogue, parking the whole 128 x 256 tile in shared memory -- 20-26 us per layer, pure latency.  Narrow column tiles
// spread the same rows over 4x the CTAs with a quarter of the MMA time each; LayerNorm then runs as its own (small) launch.
// DPM_TC_SMALL_M: rows up to which the narrow tiles are taken (0 = never); DPM_TC_SMALL_BN: their width (32 / 64 / 128).
static int small_m_rows() {
    static const int v = getenv("DPM_TC_SMALL_M") ? atoi(getenv("DPM_TC_SMALL_M")) : 1024;
    return v;
}
static int small_m_bn() {
    static const int v = getenv("DPM_TC_SMALL_BN") ? atoi(getenv("DPM_TC_SMALL_BN")) : 64;
    return (v == 32 || v == 64 || v == 128) ? v : 64;
}

int linear_tc_launch(const float *X, int ldx, long long sX, const float *W, int ldw, long long sW, const float *bias,
                     const float *res, int ldres, float *Y, int ldy, long long sY, int M, int N, int K, int nbatch,
                     int act, cudaStream_t st);

// Y = act(LayerNorm_N(X W^T + bias + res) * gamma + beta + post) in ONE launch when the row fits one column
// tile (N <= 256) and the weights were pre-split for this call; false = not eligible (caller runs the two
// kernels).  Y may alias res and / or post (a CTA reads its own rows before it writes them).
bool linear_ln_tc_launch(const float *X, int ldx, const float *W, int ldw, const float *bias, const float *res,
                         int ldres, const float *gamma, const float *beta, const float *post, int ldpost, float *Y,
                         int ldy, int M, int N, int K, int act, cudaStream_t st, int *rc, float *tmp) {
    static const bool off = getenv("DPM_NO_LN_FUSION") != nullptr;
    if (off || N > 256 || (N & 3) || M < 1 || (K & 3) || (ldx & 3) || (((uintptr_t)X) & 15)) return false;
    if ((ldy & 3) || (((uintptr_t)Y) & 15) || (bias && (((uintptr_t)bias) & 15)) || (((uintptr_t)gamma) & 15) ||
        (((uintptr_t)beta) & 15))
        return false;
    if (res && ((ldres & 3) || (((uintptr_t)res) & 15))) return false;
    if (post && ((ldpost & 3) || (((uintptr_t)post) & 15))) return false;
    const float *Ws = split_lookup(W, N, K, ldw);
    if (!Ws || getenv("DPM_NO_TC")) return false;
    if (M <= small_m_rows() && N > small_m_bn() && tmp && tmp != X && tmp != res && (((uintptr_t)tmp) & 15) == 0) {
        // few rows: narrow column tiles (4x the CTAs, a quarter of the MMA time each) leave acc + bias + res in `tmp`;
        // the LayerNorm is a second, small launch that sums in the fused epilogue's order -- same bits either way
        *rc = linear_tc_launch(X, ldx, 0, W, ldw, 0, bias, res, ldres, tmp, N, 0, M, N, K, 1, DPM_ACT_NONE, st);
        if (*rc != DPM_OK) return true;
        const int BN = N <= 128 ? 128 : 256;
        tc::ln_fused_order_kernel<<<(M + 31) / 32, 256, 0, st>>>(tmp, N, gamma, beta, post, ldpost, Y, ldy, M, N, BN, act);
        count_launch("layernorm", st);
        *rc = cudaGetLastError() == cudaSuccess ? DPM_OK : fail(DPM_ERR_CUDA, "layernorm launch failed");
        return true;
    }
    prof_note((long long)M, (long long)N * K);
    const tc::LnArgs ln{gamma, beta, post, ldpost};
    const long long lo = (long long)N * ((K + 31) & ~31);
#define DPM_TC_ARGS X, ldx, 0, Ws, K, 0, bias, res, ldres, Y, ldy, 0, M, N, K, 1, act, lo, st, ln
    if (N <= 32) *rc = tc::launch_t<32, 4, true, true>(DPM_TC_ARGS);
    else if (N <= 64) *rc = tc::launch_t<64, 4, true, true>(DPM_TC_ARGS);
    else if (N <= 128) *rc = tc::launch_t<128, 3, true, true>(DPM_TC_ARGS);
    else *rc = tc::launch_t<256, 2, true, true>(DPM_TC_ARGS);
#undef DPM_TC_ARGS
    return true;
}

bool linear_tc_eligible(const float *X, int ldx, long long sX, const float *W, int ldw, long long sW, int M, int N, int K) {
    static const bool off = getenv("DPM_NO_TC") != nullptr;
    if (off) return false;
    if (M < 1 || N < 8 || K < 4) return false;
    if ((K & 3) || (ldx & 3) || (sX & 3) || (((uintptr_t)X) & 15)) return false;
    if (split_lookup(W, N, K, ldw)) return true;  // compact pre-split copy: the original layout does not matter
    if ((ldw & 3) || (sW & 3) || (((uintptr_t)W) & 15)) return false;
    return true;
}

int linear_tc_launch(const float *X, int ldx, long long sX, const float *W, int ldw, long long sW, const float *bias,
                     const float *res, int ldres, float *Y, int ldy, long long sY, int M, int N, int K, int nbatch,
                     int act, cudaStream_t st) {
    prof_note((long long)M * nbatch, (long long)N * K);
    // pre-split weights registered for this call (split_weights_*): hi at the returned pointer, lo right after
    const float *Ws = (nbatch == 1) ? split_lookup(W, N, K, ldw) : nullptr;
    if (Ws) {
        const long long lo = (long long)N * ((K + 31) & ~31);
#define DPM_TC_ARGS X, ldx, sX, Ws, K, 0, bias, res, ldres, Y, ldy, sY, M, N, K, nbatch, act, lo, st
        if (M <= small_m_rows() && N > small_m_bn()) {
            const int bn = small_m_bn();
            if (bn == 32) return tc::launch_t<32, 4, true>(DPM_TC_ARGS);
            if (bn == 64) return tc::launch_t<64, 4, true>(DPM_TC_ARGS);
            return tc::launch_t<128, 3, true>(DPM_TC_ARGS);
        }
        if (N <= 32) return tc::launch_t<32, 4, true>(DPM_TC_ARGS);
        if (N <= 64) return tc::launch_t<64, 4, true>(DPM_TC_ARGS);
        if (N <= 128) return tc::launch_t<128, 3, true>(DPM_TC_ARGS);
        return tc::launch_t<256, 2, true>(DPM_TC_ARGS);
#undef DPM_TC_ARGS
    }
#define DPM_TC_ARGS X, ldx, sX, W, ldw, sW, bias, res, ldres, Y, ldy, sY, M, N, K, nbatch, act, 0, st
    if (N <= 32) return tc::launch_t<32, 4, false>(DPM_TC_ARGS);
    if (N <= 64) return tc::launch_t<64, 4, false>(DPM_TC_ARGS);
    if (N <= 128) return tc::launch_t<128, 3, false>(DPM_TC_ARGS);
    return tc::launch_t<256, 2, false>(DPM_TC_ARGS);
#undef DPM_TC_ARGS
}

// ---------------------------------------------------------------------------------------
// per-call weight split: every 2-D weight the call will push through the tensor-core path is split ONCE
// into compact hi / lo tf32 matrices in the caller's workspace (one launch for all of them), instead of
// once per CTA per K block inside the GEMM.
// ---------------------------------------------------------------------------------------
constexpr int SPLIT_MAX = 56;
struct SplitJob {
    const float *src;
    float *dst;
    int rows, cols, ld, pad;
};
struct SplitTable {
    SplitJob job[SPLIT_MAX];
};
static thread_local SplitTable g_split;
static thread_local int g_nsplit = 0;

// dst: hi then lo, each (K blocks of 32) x rows x 32 floats, chunk c of row n stored at chunk (c ^ n) & 7; K padded
// to a multiple of 32 with zeros
__global__ void __launch_bounds__(256) split_weights_kernel(const __grid_constant__ SplitTable t) {
    const SplitJob j = t.job[blockIdx.y];
    const int kpad = j.pad, n = j.rows * kpad;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const int r = i / kpad, c = i - r * kpad;
        const float v = c < j.cols ? j.src[(size_t)r * j.ld + c] : 0.f;
        const float h = tc::tf32_rna(v);
        const int kb = c >> 5, ch = (c >> 2) & 7, e = c & 3;
        const size_t o = ((size_t)kb * j.rows + r) * 32 + (((ch ^ r) & 7) << 2) + e;
        j.dst[o] = h;
        j.dst[(size_t)n + o] = tc::tf32_rna(v - h);
    }
}

size_t split_floats(int rows, int cols) { return (size_t)2 * rows * ((cols + 31) & ~31); }

void split_begin() { g_nsplit = 0; }

void split_add(Arena &a, const float *W, int rows, int cols, int ld) {
    if (rows < 8 || cols < 4 || (cols & 3)) return;  // never taken by the tensor-core path
    float *dst = a.get<float>(split_floats(rows, cols));
    if (a.dry || !dst || !W || g_nsplit >= SPLIT_MAX) return;
    SplitJob &j = g_split.job[g_nsplit++];
    j.src = W; j.dst = dst; j.rows = rows; j.cols = cols; j.ld = ld; j.pad = (cols + 31) & ~31;
}

// ---- optional reuse of the split copies across calls ---------------------------------------------------
// By default every call re-splits its weights (nothing derived from the weights outlives a call).  A caller that KNOWS
// its weights and its workspace are unchanged since its previous call can say so with dpm_set_weights_epoch(e != 0): the
// split launch is then skipped when the same epoch last wrote the same jobs (same sources, shapes, destinations).  The
// epoch must change whenever a weight tensor or the workspace buffer does (deeppointmap_b200 derives it from the
// parameters' version counters and the scratch buffer's identity); 0 switches the reuse off again.
static thread_local unsigned long long g_epoch = 0ull;
struct SplitSeen {
    unsigned long long epoch, hash;
    const float *dst0;
};
static thread_local SplitSeen g_seen[16];
static thread_local int g_seen_next = 0;

static unsigned long long split_hash() {
    unsigned long long h = 1469598103934665603ull;
    auto mix = [&](unsigned long long v) { h = (h ^ v) * 1099511628211ull; };
    for (int i = 0; i < g_nsplit; ++i) {
        const SplitJob &j = g_split.job[i];
        mix((unsigned long long)(uintptr_t)j.src); mix((unsigned long long)(uintptr_t)j.dst);
        mix(((unsigned long long)j.rows << 32) | (unsigned)j.cols); mix((unsigned long long)j.ld);
    }
    return h;
}

int split_run(cudaStream_t st) {
    if (g_nsplit == 0) return DPM_OK;
    if (g_epoch != 0ull) {
        const unsigned long long h = split_hash();
        const float *dst0 = g_split.job[0].dst;
        for (int i = 0; i < 16; ++i)
            if (g_seen[i].dst0 == dst0) {
                if (g_seen[i].epoch == g_epoch && g_seen[i].hash == h) return DPM_OK;  // still valid: no launch
                g_seen[i].dst0 = nullptr;
            }
        g_seen[g_seen_next] = SplitSeen{g_epoch, h, dst0};
        g_seen_next = (g_seen_next + 1) & 15;
    }
    int maxn = 0;
    for (int i = 0; i < g_nsplit; ++i) maxn = g_split.job[i].rows * g_split.job[i].pad > maxn ? g_split.job[i].rows * g_split.job[i].pad : maxn;
    int gx = (maxn + 1023) / 1024;  // ~4 elements per thread for the largest matrix
    gx = gx < 1 ? 1 : (gx > 1024 ? 1024 : gx);
    split_weights_kernel<<<dim3(gx, g_nsplit, 1), 256, 0, st>>>(g_split);
    DPM_CHECK_LAUNCH("split_weights", st);
    return DPM_OK;
}

const float *split_lookup(const float *W, int rows, int cols, int ld) {
    for (int i = 0; i < g_nsplit; ++i) {
        const SplitJob &j = g_split.job[i];
        if (j.src == W && j.rows == rows && j.cols == cols && j.ld == ld) return j.dst;
    }
    return nullptr;
}

}  // namespace dpm

#ifdef DPM_TC_PROFILE
extern "C" int dpm_debug_tc_profile(unsigned long long *out, int nctas) {
    if (cudaMemcpyFromSymbol(out, dpm::tc::tc_prof, sizeof(unsigned long long) * 6 * (nctas > 2048 ? 2048 : nctas)) != cudaSuccess) return -1;
    return 0;
}
#endif

extern "C" void dpm_set_weights_epoch(unsigned long long epoch) { dpm::g_epoch = epoch; }
