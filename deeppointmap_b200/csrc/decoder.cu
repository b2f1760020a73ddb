// decoder.cu -- Decoder.registration_forward / loop_detection_forward
// (network/decoder/decoder.py:91-265, descriptor_attention.py, heads.py) as native calls.
//
// Token layout: one row per descriptor.  For pair p the M src rows are followed by the N dst
// rows (R = P*(M+N) rows in total), so every shared-weight layer (projection, in/out
// projections, MLP, LayerNorms, heads) is ONE launch over all rows of all pairs, and
// self/cross attention are one launch over 2P (query-range, key-range) problems.
// Data-dependent shapes of the reference (offset filter, sigma clipping) are restated with
// counts + masks on the device: there is no host sync anywhere in here.
#include <stdlib.h>

#include "common.cuh"

namespace dpm {

// ---------------------------------------------------------------------------------------
// channel-first descriptors -> row-major features + float4 xyz
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dec_unpack_kernel(const float *__restrict__ src, const float *__restrict__ dst, int M, int N, int Cf,
                  float *__restrict__ fea, float4 *__restrict__ xyz, const uint8_t *__restrict__ src_pad,
                  const uint8_t *__restrict__ dst_pad, uint8_t *__restrict__ rowmask) {
    __shared__ float tile[32][33];
    const int z = blockIdx.z, p = z >> 1, side = z & 1;
    const int L = side ? N : M;
    const float *in = (side ? dst : src) + (size_t)p * (Cf + 3) * L;
    const size_t row0 = (size_t)p * (M + N) + (side ? M : 0);
    const int l0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    if (l0 >= L) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r, l = l0 + tx;
        tile[r][tx] = (c < Cf && l < L) ? in[(size_t)c * L + l] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int l = l0 + r, c = c0 + tx;
        if (l < L && c < Cf) fea[(row0 + l) * Cf + c] = tile[tx][r];
    }
    if (blockIdx.y == 0 && ty == 0) {
        const int l = l0 + tx;
        if (l < L) {
            xyz[row0 + l] = make_float4(in[(size_t)Cf * L + l], in[(size_t)(Cf + 1) * L + l], in[(size_t)(Cf + 2) * L + l], 0.f);
            if (rowmask) {
                const uint8_t *pm = side ? dst_pad : src_pad;
                rowmask[row0 + l] = pm ? (pm[(size_t)p * L + l] != 0) : 0;
            }
        }
    }
}

// PositionEmbeddingCoordsSine.forward (descriptor_attention.py:66-83)
__global__ void __launch_bounds__(256)
posenc_kernel(const float *__restrict__ xyz, int ldx, const float *__restrict__ dim_t, int npf, float *__restrict__ emb,
              int R, int C) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)R * C) return;
    const int r = (int)(t / C), c = (int)(t % C);
    float v = 0.f;
    if (c < 3 * npf) {
        const int a = c / npf, i = c % npf;
        const float x = __fmul_rn(xyz[(size_t)r * ldx + a], 3.14159274101257324f);  // coor * fp32(pi)
        const float pd = __fdiv_rn(x, dim_t[i]);
        v = (i & 1) ? cosf(pd) : sinf(pd);
    }
    emb[t] = v;
}

int posenc_launch(const float *xyz, int ldx, const float *dim_t, int npf, float *emb, int R, int C, cudaStream_t st) {
    if (R <= 0 || C <= 0 || npf <= 0) return fail(DPM_ERR_SHAPE, "posenc: bad shape");
    const long long total = (long long)R * C;
    posenc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(xyz, ldx, dim_t, npf, emb, R, C);
    DPM_CHECK_LAUNCH("posenc", st);
    return DPM_OK;
}

// ---------------------------------------------------------------------------------------
// multi-head attention core (head_dim 32) on the tensor cores: mma.sync m16n8k8 TF32 with the 3xTF32 error
// compensation (x = hi + lo, products hi*hi + hi*lo + lo*hi), fp32 accumulation, flash-style
// online softmax in registers.  CTA = 128 queries (8 warps x 16 rows) x one head; keys / values
// are staged 64 at a time in shared memory already split into hi / lo (row stride 36 floats:
// conflict-free fragment loads).  S = Q K^T accumulates in the C fragments; those registers ARE the
// A fragments of P V once the 8 keys of a k-step are taken in the order (0,2,4,6,1,3,5,7), which
// only changes which V rows are loaded into the B fragment.
// ---------------------------------------------------------------------------------------
constexpr int AT_BQ = 128, AT_BK = 64, AT_LD = 36, AT_T = AT_BQ * 2;  // 16 query rows per warp; 64 / 256 queries per CTA measured 10 % slower

__device__ __forceinline__ unsigned tf32_bits(float x) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32_16n8k8(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// LONG (key ranges of more than 512 rows, i.e. map tiles): the tensor core's fp32 accumulate TRUNCATES, a bias
// of ~2^-24 per k-step that is invisible over the 32 k-steps of a 256-key set but reaches 2e-5 over the 512
// k-steps of a 4096-key map tile (measured against fp64).  There every 64-key tile accumulates P V into fresh
// registers and is folded into the running output with one IEEE fma per element.
template <bool LONG>
__global__ void __launch_bounds__(AT_T, LONG ? 1 : 512 / AT_T)
attention_tc_kernel(const float *__restrict__ Q, int ldq, const float *__restrict__ Kp, int ldk,
                    const float *__restrict__ Vp, int ldv, float *__restrict__ O, int ldo, const int *__restrict__ prob,
                    int M, int N, int mode, const uint8_t *__restrict__ kmask) {
    __shared__ __align__(16) unsigned Kh[AT_BK * AT_LD], Kl[AT_BK * AT_LD], Vh[AT_BK * AT_LD], Vl[AT_BK * AT_LD];
    __shared__ uint8_t kdrop[AT_BK];  // key_padding_mask of the staged keys (descriptor_attention.py:33-42): 1 = ignored
    const int z = blockIdx.z, head = blockIdx.y;
    int q0, Lq, k0, Lk;
    if (prob) {
        q0 = prob[4 * z]; Lq = prob[4 * z + 1]; k0 = prob[4 * z + 2]; Lk = prob[4 * z + 3];
    } else {
        const int p = z >> 1, side = z & 1, base = p * (M + N);
        q0 = base + (side ? M : 0);
        Lq = side ? N : M;
        const int kvside = mode ? !side : side;  // mode 0: self, 1: cross
        k0 = base + (kvside ? M : 0);
        Lk = kvside ? N : M;
    }
    if (blockIdx.x * AT_BQ >= Lq) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int r0 = blockIdx.x * AT_BQ + warp * 16;  // this warp's 16 query rows
    const float NEG = -__int_as_float(0x7f800000);

    // Q fragments (scaled by 1/sqrt(d) first, as nn.MultiheadAttention does), split once
    unsigned qh[4][4], ql[4][4];
    {
        const float scale = 0.17677669529663687f;
        const int ra = min(r0 + g, Lq - 1), rb = min(r0 + g + 8, Lq - 1);
        const float *qa = Q + (size_t)(q0 + ra) * ldq + head * 32, *qb = Q + (size_t)(q0 + rb) * ldq + head * 32;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const float v[4] = {qa[8 * s + t] * scale, qb[8 * s + t] * scale, qa[8 * s + t + 4] * scale,
                                qb[8 * s + t + 4] * scale};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                qh[s][i] = tf32_bits(v[i]);
                ql[s][i] = tf32_bits(v[i] - __uint_as_float(qh[s][i]));
            }
        }
    }
    float oh[4][4], ox[4][4];  // O accumulators: hi*hi terms / cross terms (kept apart: the MMA's fp32 add truncates)
#pragma unroll
    for (int d = 0; d < 4; ++d)
#pragma unroll
        for (int i = 0; i < 4; ++i) { oh[d][i] = 0.f; ox[d][i] = 0.f; }
    float m0 = NEG, m1 = NEG, l0 = 0.f, l1 = 0.f;  // rows g and g+8; l is this thread's partial row sum

    for (int c0 = 0; c0 < Lk; c0 += AT_BK) {
        __syncthreads();
        for (int e = tid; e < AT_BK * 8; e += AT_T) {
            const int key = e >> 3, part = e & 7;
            float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
            if (c0 + key < Lk) {
                const size_t row = (size_t)(k0 + c0 + key);
                kk = *reinterpret_cast<const float4 *>(Kp + row * ldk + head * 32 + 4 * part);
                vv = *reinterpret_cast<const float4 *>(Vp + row * ldv + head * 32 + 4 * part);
            }
            uint4 h, l;
            h.x = tf32_bits(kk.x); l.x = tf32_bits(kk.x - __uint_as_float(h.x));
            h.y = tf32_bits(kk.y); l.y = tf32_bits(kk.y - __uint_as_float(h.y));
            h.z = tf32_bits(kk.z); l.z = tf32_bits(kk.z - __uint_as_float(h.z));
            h.w = tf32_bits(kk.w); l.w = tf32_bits(kk.w - __uint_as_float(h.w));
            *reinterpret_cast<uint4 *>(&Kh[key * AT_LD + 4 * part]) = h;
            *reinterpret_cast<uint4 *>(&Kl[key * AT_LD + 4 * part]) = l;
            h.x = tf32_bits(vv.x); l.x = tf32_bits(vv.x - __uint_as_float(h.x));
            h.y = tf32_bits(vv.y); l.y = tf32_bits(vv.y - __uint_as_float(h.y));
            h.z = tf32_bits(vv.z); l.z = tf32_bits(vv.z - __uint_as_float(h.z));
            h.w = tf32_bits(vv.w); l.w = tf32_bits(vv.w - __uint_as_float(h.w));
            *reinterpret_cast<uint4 *>(&Vh[key * AT_LD + 4 * part]) = h;
            *reinterpret_cast<uint4 *>(&Vl[key * AT_LD + 4 * part]) = l;
            if (kmask && part == 0) kdrop[key] = (c0 + key < Lk) ? kmask[k0 + c0 + key] : 1;
        }
        __syncthreads();

        // ---- S = Q K^T for 16 rows x 64 keys: sc[j] = keys 8j..8j+7 ----
        float sc[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const int o = (8 * j + g) * AT_LD + 8 * s + t;
                const unsigned b0h = Kh[o], b1h = Kh[o + 4], b0l = Kl[o], b1l = Kl[o + 4];
                mma_tf32_16n8k8(sc[j], ql[s], b0h, b1h);
                mma_tf32_16n8k8(sc[j], qh[s], b0l, b1l);
                mma_tf32_16n8k8(sc[j], qh[s], b0h, b1h);
            }
        }
        if (kmask) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (kdrop[8 * j + 2 * t]) { sc[j][0] = NEG; sc[j][2] = NEG; }
                if (kdrop[8 * j + 2 * t + 1]) { sc[j][1] = NEG; sc[j][3] = NEG; }
            }
        }
        const int nvalid = Lk - c0;
        if (nvalid < AT_BK) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (8 * j + 2 * t >= nvalid) { sc[j][0] = NEG; sc[j][2] = NEG; }
                if (8 * j + 2 * t + 1 >= nvalid) { sc[j][1] = NEG; sc[j][3] = NEG; }
            }
        }
        // ---- online softmax ----
        float mx0 = NEG, mx1 = NEG;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            mx0 = fmaxf(mx0, fmaxf(sc[j][0], sc[j][1]));
            mx1 = fmaxf(mx1, fmaxf(sc[j][2], sc[j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        if (kmask) {  // every key so far masked: keep the exponents finite (the tile then contributes exp(-inf) = 0)
            mn0 = mn0 == NEG ? 0.f : mn0;
            mn1 = mn1 == NEG ? 0.f : mn1;
        }
        const float corr0 = expf(m0 - mn0), corr1 = expf(m1 - mn1);
        m0 = mn0; m1 = mn1;
        l0 *= corr0; l1 *= corr1;
        float th[LONG ? 4 : 1][4], tx[LONG ? 4 : 1][4];  // LONG: this tile's P V (hi*hi / cross terms)
        if (LONG) {
#pragma unroll
            for (int d = 0; d < (LONG ? 4 : 1); ++d)
#pragma unroll
                for (int i = 0; i < 4; ++i) { th[d][i] = 0.f; tx[d][i] = 0.f; }
        } else {
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                oh[d][0] *= corr0; oh[d][1] *= corr0; oh[d][2] *= corr1; oh[d][3] *= corr1;
                ox[d][0] *= corr0; ox[d][1] *= corr0; ox[d][2] *= corr1; ox[d][3] *= corr1;
            }
        }
        // ---- O += P V, 8 keys per k-step; k index t <-> key 2t, k index t+4 <-> key 2t+1 ----
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float p0 = expf(sc[j][0] - mn0), p1 = expf(sc[j][1] - mn0);
            const float p2 = expf(sc[j][2] - mn1), p3 = expf(sc[j][3] - mn1);
            l0 += p0 + p1;
            l1 += p2 + p3;
            unsigned ph[4], pl[4];
            ph[0] = tf32_bits(p0); pl[0] = tf32_bits(p0 - __uint_as_float(ph[0]));  // (row g,   key 2t)
            ph[1] = tf32_bits(p2); pl[1] = tf32_bits(p2 - __uint_as_float(ph[1]));  // (row g+8, key 2t)
            ph[2] = tf32_bits(p1); pl[2] = tf32_bits(p1 - __uint_as_float(ph[2]));  // (row g,   key 2t+1)
            ph[3] = tf32_bits(p3); pl[3] = tf32_bits(p3 - __uint_as_float(ph[3]));  // (row g+8, key 2t+1)
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                const int o = (8 * j + 2 * t) * AT_LD + 8 * d + g;
                const unsigned b0h = Vh[o], b1h = Vh[o + AT_LD], b0l = Vl[o], b1l = Vl[o + AT_LD];
                if (LONG) {
                    mma_tf32_16n8k8(th[LONG ? d : 0], ph, b0h, b1h);
                    mma_tf32_16n8k8(tx[LONG ? d : 0], ph, b0l, b1l);
                    mma_tf32_16n8k8(tx[LONG ? d : 0], pl, b0h, b1h);
                } else {
                    mma_tf32_16n8k8(oh[d], ph, b0h, b1h);
                    mma_tf32_16n8k8(ox[d], ph, b0l, b1l);
                    mma_tf32_16n8k8(ox[d], pl, b0h, b1h);
                }
            }
        }
        if (LONG) {
#pragma unroll
            for (int d = 0; d < (LONG ? 4 : 1); ++d) {
                oh[d][0] = fmaf(oh[d][0], corr0, th[d][0] + tx[d][0]);
                oh[d][1] = fmaf(oh[d][1], corr0, th[d][1] + tx[d][1]);
                oh[d][2] = fmaf(oh[d][2], corr1, th[d][2] + tx[d][2]);
                oh[d][3] = fmaf(oh[d][3], corr1, th[d][3] + tx[d][3]);
            }
        }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const int ra = r0 + g, rb = r0 + g + 8;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        if (ra < Lq)
            *reinterpret_cast<float2 *>(O + (size_t)(q0 + ra) * ldo + head * 32 + 8 * d + 2 * t) =
                make_float2((oh[d][0] + ox[d][0]) / l0, (oh[d][1] + ox[d][1]) / l0);
        if (rb < Lq)
            *reinterpret_cast<float2 *>(O + (size_t)(q0 + rb) * ldo + head * 32 + 8 * d + 2 * t) =
                make_float2((oh[d][2] + ox[d][2]) / l1, (oh[d][3] + ox[d][3]) / l1);
    }
}

int attention_launch(const float *q, int ldq, const float *k, int ldk, const float *v, int ldv, float *out, int ldo,
                     const int *prob, int nprob, int maxLq, int M, int N, int mode, int heads, cudaStream_t st,
                     const uint8_t *kmask = nullptr) {
    if (nprob <= 0 || heads <= 0 || maxLq <= 0) return fail(DPM_ERR_SHAPE, "attention: bad shape");
    if ((ldq | ldk | ldv | ldo) & 3) return fail(DPM_ERR_UNSUPPORTED, "attention: leading dimensions must be multiples of 4");
    // the decoder's pairs: flash attention on tcgen05 / TMEM (attention_tc5.cu) at every size -- 2.2x this kernel on map
    // tiles, 1.3x on the 256-descriptor odometry pairs (0.61 -> 0.47 ms per 32-pair step).  DPM_ATT_IMPL=1 keeps mma.sync.
    static const int impl_env = getenv("DPM_ATT_IMPL") ? atoi(getenv("DPM_ATT_IMPL")) : 0;
    if (!prob && impl_env != 1)
        return attention_tc5_launch(q, ldq, k, ldk, v, ldv, out, ldo, nprob / 2, M, N, mode, heads, kmask, st);
    dim3 grid((maxLq + AT_BQ - 1) / AT_BQ, heads, nprob);
    // the longest key range of the launch decides (prob == nullptr: the two sides of the pairs; else the caller's table,
    // which this host code cannot read: the accurate variant then)
    const bool lng = prob ? true : (M > 512 || N > 512);
    if (lng) attention_tc_kernel<true><<<grid, AT_T, 0, st>>>(q, ldq, k, ldk, v, ldv, out, ldo, prob, M, N, mode, kmask);
    else attention_tc_kernel<false><<<grid, AT_T, 0, st>>>(q, ldq, k, ldk, v, ldv, out, ldo, prob, M, N, mode, kmask);
    DPM_CHECK_LAUNCH("attention", st);
    return DPM_OK;
}

// ---------------------------------------------------------------------------------------
// pairing: L2 normalise, dual softmax, global top-k   (decoder.py:180-192)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) l2norm_rows_kernel(float *__restrict__ X, int R, int C) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= R) return;
    float *x = X + (size_t)row * C;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) q = fmaf(x[c], x[c], q);
    const float nrm = fmaxf(sqrtf(warp_sum(q)), 1e-12f);  // F.normalize eps
    for (int c = lane; c < C; c += 32) x[c] = x[c] / nrm;
}

// rows of the offset head input: r < k: [F_src[i_r], F_dst[j_r]] ; k + r: [F_dst[j_r], F_src[i_r]]
__global__ void __launch_bounds__(256)
pair_gather_kernel(const float *__restrict__ F, int C, int M, int N, int k, const int32_t *__restrict__ si,
                   const int32_t *__restrict__ di, float *__restrict__ X) {
    const int p = blockIdx.y, r = blockIdx.x;  // r in [0, 2k)
    const int rr = r < k ? r : r - k;
    const size_t base = (size_t)p * (M + N);
    const float *fs = F + (base + si[(size_t)p * k + rr]) * C;
    const float *fd = F + (base + M + di[(size_t)p * k + rr]) * C;
    float *x = X + ((size_t)p * 2 * k + r) * 2 * C;
    const float *a = r < k ? fs : fd, *b = r < k ? fd : fs;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        x[c] = a[c];
        x[C + c] = b[c];
    }
}

// _get_corres_sets (decoder.py:202-225): 2k candidate correspondences, offset filter, ordered
// compaction.  One CTA per pair.  out: src/dst (P,3,2k), w (P,2k), count (P).
__global__ void __launch_bounds__(256)
corres_kernel(const float4 *__restrict__ xyz, const float *__restrict__ off, const int32_t *__restrict__ si,
              const int32_t *__restrict__ di, const float *__restrict__ conf, int M, int N, int k, float lim,
              float *__restrict__ csrc, float *__restrict__ cdst, float *__restrict__ cw, int32_t *__restrict__ count) {
    __shared__ int s_base, s_wcnt[8];
    const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t base = (size_t)p * (M + N);
    const int K2 = 2 * k;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int r0 = 0; r0 < K2; r0 += 256) {
        const int r = r0 + tid;
        bool keep = false;
        float3 s = make_float3(0.f, 0.f, 0.f), d = make_float3(0.f, 0.f, 0.f);
        float w = 0.f;
        if (r < K2) {
            const int rr = r < k ? r : r - k;
            const float4 xs = xyz[base + si[(size_t)p * k + rr]];
            const float4 xd = xyz[base + M + di[(size_t)p * k + rr]];
            const float *o = off + ((size_t)p * K2 + r) * 3;
            const float ox = o[0], oy = o[1], oz = o[2];
            keep = (ox * ox + oy * oy + oz * oz) <= lim;
            if (r < k) { s = make_float3(xs.x + ox, xs.y + oy, xs.z + oz); d = make_float3(xd.x, xd.y, xd.z); }
            else       { s = make_float3(xs.x, xs.y, xs.z); d = make_float3(xd.x + ox, xd.y + oy, xd.z + oz); }
            w = conf[(size_t)p * k + rr];
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_wcnt[warp] = __popc(bal);
        __syncthreads();
        int before = s_base;
        for (int ww = 0; ww < warp; ++ww) before += s_wcnt[ww];
        const int pos = before + __popc(bal & ((1u << lane) - 1u));
        if (keep) {
            float *ps = csrc + (size_t)p * 3 * K2, *pd = cdst + (size_t)p * 3 * K2;
            ps[pos] = s.x; ps[K2 + pos] = s.y; ps[2 * K2 + pos] = s.z;
            pd[pos] = d.x; pd[K2 + pos] = d.y; pd[2 * K2 + pos] = d.z;
            cw[(size_t)p * K2 + pos] = w;
        }
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int ww = 0; ww < 8; ++ww) t += s_wcnt[ww];
            s_base += t;
        }
        __syncthreads();
    }
    if (tid == 0) count[p] = s_base;
}

// ---------------------------------------------------------------------------------------
// weighted Kabsch + 3-sigma loop (decoder.py:227-265).  One CTA per problem.
// ---------------------------------------------------------------------------------------
constexpr int KAB_T = 256;
constexpr int KAB_MAX = 4096;

template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *red /* [8][NV] */, int tid) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum_d(v[i]);
    __syncthreads();
    if ((tid & 31) == 0)
#pragma unroll
        for (int i = 0; i < NV; ++i) red[(tid >> 5) * NV + i] = v[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double t = 0.0;
        for (int w = 0; w < KAB_T / 32; ++w) t += red[w * NV + i];
        v[i] = t;
    }
}

// one-sided Jacobi SVD of a 3x3 (fp64): A = U diag(s) V^T; returns R = V U^T
__device__ void svd3_rot(const double S[9], double R[9]) {
    double A[3][3], V[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { A[i][j] = S[3 * i + j]; V[i][j] = i == j ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 30; ++sweep) {
        double offmax = 0.0;
        for (int pp = 0; pp < 2; ++pp)
            for (int q = pp + 1; q < 3; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < 3; ++i) { alpha += A[i][pp] * A[i][pp]; beta += A[i][q] * A[i][q]; gamma += A[i][pp] * A[i][q]; }
                if (gamma == 0.0) continue;
                const double rel = fabs(gamma) / sqrt(alpha * beta + 1e-300);
                offmax = fmax(offmax, rel);
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int i = 0; i < 3; ++i) {
                    const double ap = A[i][pp], aq = A[i][q];
                    A[i][pp] = c * ap - s * aq; A[i][q] = s * ap + c * aq;
                    const double vp = V[i][pp], vq = V[i][q];
                    V[i][pp] = c * vp - s * vq; V[i][q] = s * vp + c * vq;
                }
            }
        if (offmax < 1e-15) break;
    }
    double U[3][3], sig[3];
    for (int j = 0; j < 3; ++j) {
        sig[j] = sqrt(A[0][j] * A[0][j] + A[1][j] * A[1][j] + A[2][j] * A[2][j]);
    }
    const double smax = fmax(sig[0], fmax(sig[1], sig[2]));
    int bad = -1, nbad = 0;
    for (int j = 0; j < 3; ++j) {
        if (sig[j] > smax * 1e-14 && sig[j] > 0.0) {
            for (int i = 0; i < 3; ++i) U[i][j] = A[i][j] / sig[j];
        } else { bad = j; ++nbad; }
    }
    if (nbad == 1) {  // rank 2: complete U with the cross product of the other two columns
        const int a = (bad + 1) % 3, b = (bad + 2) % 3;
        U[0][bad] = U[1][a] * U[2][b] - U[2][a] * U[1][b];
        U[1][bad] = U[2][a] * U[0][b] - U[0][a] * U[2][b];
        U[2][bad] = U[0][a] * U[1][b] - U[1][a] * U[0][b];
    } else if (nbad > 1) {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) U[i][j] = V[i][j];  // degenerate: R = I
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[3 * i + j] = V[i][0] * U[j][0] + V[i][1] * U[j][1] + V[i][2] * U[j][2];
}

__global__ void __launch_bounds__(KAB_T)
kabsch_kernel(const float *__restrict__ srcp, const float *__restrict__ dstp, const float *__restrict__ wp,
              const int32_t *__restrict__ count, int ldk, float *__restrict__ result, uint8_t *__restrict__ inl_out,
              float *__restrict__ conf_out) {
    __shared__ unsigned long long skey[KAB_MAX];
    float *serr = reinterpret_cast<float *>(skey);  // the sort keys are dead once the top-64 are marked
    __shared__ uint8_t sinl[KAB_MAX];
    __shared__ double red[8 * 16];
    __shared__ float sRT[12];
    __shared__ int s_flag[2], s_wcnt[8], s_base;
    const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = min(count[p], min(ldk, KAB_MAX));
    const float *sx = srcp + (size_t)p * 3 * ldk, *sy = sx + ldk, *sz = sy + ldk;
    const float *dx = dstp + (size_t)p * 3 * ldk, *dy = dx + ldk, *dz = dy + ldk;
    const float *w = wp + (size_t)p * ldk;
    float *res = result + (size_t)p * DPM_REG_STRIDE;

    // initial inliers: w > 0.5, plus the 64 largest weights (ties: lowest index)
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    if (np2 < 2) np2 = 2;
    for (int i = tid; i < np2; i += KAB_T)
        skey[i] = i < n ? (((unsigned long long)__float_as_uint(fmaxf(w[i], 0.f)) << 32) | (unsigned)(0xffffffffu - (unsigned)i)) : 0ull;
    for (int i = tid; i < n; i += KAB_T) sinl[i] = w[i] > 0.5f;
    __syncthreads();
    bitonic_sort_desc(skey, np2, tid, KAB_T);
    if (tid < 64 && tid < n) sinl[0xffffffffu - (unsigned)skey[tid]] = 1;
    __syncthreads();

    int iters = 0;
    float rmse = 0.f;
    while (true) {
        double a[16];
        for (int i = 0; i < 16; ++i) a[i] = 0.0;
        for (int i = tid; i < n; i += KAB_T)
            if (sinl[i]) {
                const double wi = w[i];
                a[0] += wi;
                a[1] += wi * sx[i]; a[2] += wi * sy[i]; a[3] += wi * sz[i];
                a[4] += wi * dx[i]; a[5] += wi * dy[i]; a[6] += wi * dz[i];
            }
        block_sum<7>(reinterpret_cast<double(&)[7]>(a), red, tid);
        const float wsum = (float)a[0];
        const float csx = (float)a[1] / wsum, csy = (float)a[2] / wsum, csz = (float)a[3] / wsum;
        const float cdx = (float)a[4] / wsum, cdy = (float)a[5] / wsum, cdz = (float)a[6] / wsum;
        double h[9];
        for (int i = 0; i < 9; ++i) h[i] = 0.0;
        for (int i = tid; i < n; i += KAB_T)
            if (sinl[i]) {
                const float ax = sx[i] - csx, ay = sy[i] - csy, az = sz[i] - csz;
                const float bx = dx[i] - cdx, by = dy[i] - cdy, bz = dz[i] - cdz;
                const double wi = w[i];
                h[0] += wi * ax * bx; h[1] += wi * ax * by; h[2] += wi * ax * bz;
                h[3] += wi * ay * bx; h[4] += wi * ay * by; h[5] += wi * ay * bz;
                h[6] += wi * az * bx; h[7] += wi * az * by; h[8] += wi * az * bz;
            }
        block_sum<9>(h, red, tid);
        if (tid == 0) {
            double Sm[9], R[9];
            for (int i = 0; i < 9; ++i) Sm[i] = (double)(float)h[i];  // the reference forms S in fp32, then .double()
            svd3_rot(Sm, R);
            const double cs[3] = {csx, csy, csz}, cd[3] = {cdx, cdy, cdz};
            for (int i = 0; i < 3; ++i) {
                const double t = cd[i] - (R[3 * i] * cs[0] + R[3 * i + 1] * cs[1] + R[3 * i + 2] * cs[2]);
                sRT[9 + i] = (float)t;
            }
            for (int i = 0; i < 9; ++i) sRT[i] = (float)R[i];
        }
        __syncthreads();
        float R[9], T[3];
        for (int i = 0; i < 9; ++i) R[i] = sRT[i];
        for (int i = 0; i < 3; ++i) T[i] = sRT[9 + i];
        double e[2] = {0.0, 0.0};
        int cnt = 0;
        for (int i = tid; i < n; i += KAB_T) {
            const float ex = R[0] * sx[i] + R[1] * sy[i] + R[2] * sz[i] + T[0] - dx[i];
            const float ey = R[3] * sx[i] + R[4] * sy[i] + R[5] * sz[i] + T[1] - dy[i];
            const float ez = R[6] * sx[i] + R[7] * sy[i] + R[8] * sz[i] + T[2] - dz[i];
            const float er = sqrtf(ex * ex + ey * ey + ez * ez);
            serr[i] = er;
            if (sinl[i]) { e[0] += er; ++cnt; }
        }
        e[1] = (double)cnt;
        block_sum<2>(e, red, tid);
        const double ninl = e[1];
        const float mean = (float)(e[0] / ninl);
        double v[1] = {0.0};
        for (int i = tid; i < n; i += KAB_T)
            if (sinl[i]) { const double dd = (double)serr[i] - (double)mean; v[0] += dd * dd; }
        block_sum<1>(v, red, tid);
        const float stdv = (float)sqrt(v[0] / (ninl - 1.0));  // unbiased; NaN when ninl <= 1, as torch.std
        const float thr = mean + 3.0f * stdv;
        if (tid < 2) s_flag[tid] = 0;
        __syncthreads();
        int changed = 0, newcnt = 0;
        for (int i = tid; i < n; i += KAB_T) {
            const uint8_t ni = serr[i] <= thr ? 1 : 0;  // false for NaN thr
            changed |= (ni != sinl[i]);
            newcnt += ni;
            sinl[i] = ni;
        }
        if (changed) atomicOr(&s_flag[0], 1);
        if (newcnt) atomicAdd(&s_flag[1], newcnt);
        __syncthreads();
        ++iters;
        const bool stop = iters >= 3 || s_flag[0] == 0 || s_flag[1] < 30;
        const int nfinal = s_flag[1];
        __syncthreads();
        if (stop) {
            double q[1] = {0.0};
            for (int i = tid; i < n; i += KAB_T)
                if (sinl[i]) q[0] += (double)serr[i] * (double)serr[i];
            block_sum<1>(q, red, tid);
            rmse = (float)sqrt(q[0] / (double)nfinal);
            if (tid == 0) {
                for (int i = 0; i < 9; ++i) res[DPM_REG_R + i] = R[i];
                for (int i = 0; i < 3; ++i) res[DPM_REG_T + i] = T[i];
                res[DPM_REG_RMSE] = rmse;
                res[DPM_REG_NCORR] = (float)n;
                res[DPM_REG_NINLIER] = (float)nfinal;
                res[DPM_REG_ITERS] = (float)iters;
            }
            break;
        }
    }
    // inlier mask + ordered compaction of the inlier confidences
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int r0 = 0; r0 < ldk; r0 += KAB_T) {
        const int i = r0 + tid;
        const bool in = i < n && sinl[i];
        if (inl_out && i < ldk) inl_out[(size_t)p * ldk + i] = in ? 1 : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        if (lane == 0) s_wcnt[warp] = __popc(bal);
        __syncthreads();
        int before = s_base;
        for (int ww = 0; ww < warp; ++ww) before += s_wcnt[ww];
        if (in && conf_out) conf_out[(size_t)p * ldk + before + __popc(bal & ((1u << lane) - 1u))] = w[i];
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int ww = 0; ww < 8; ++ww) t += s_wcnt[ww];
            s_base += t;
        }
        __syncthreads();
    }
}

int kabsch_launch(const float *src, const float *dst, const float *w, const int32_t *count, int P, int ldk,
                  float *result, uint8_t *inlier, float *conf_out, cudaStream_t st) {
    if (P <= 0 || ldk <= 0) return fail(DPM_ERR_SHAPE, "kabsch: bad shape");
    if (ldk > KAB_MAX) return fail(DPM_ERR_UNSUPPORTED, "kabsch: %d correspondences exceed the limit %d", ldk, KAB_MAX);
    kabsch_kernel<<<P, KAB_T, 0, st>>>(src, dst, w, count, ldk, result, inlier, conf_out);
    DPM_CHECK_LAUNCH("kabsch", st);
    return DPM_OK;
}

// mean over the tokens of each (pair, side) -> (P, 2C) = [mean src ; mean dst]   (heads.py:64-67)
__global__ void __launch_bounds__(256)
token_mean_kernel(const float *__restrict__ X, int C, int M, int N, float *__restrict__ out) {
    const int p = blockIdx.x, side = blockIdx.y;
    const int L = side ? N : M;
    const float *x = X + ((size_t)p * (M + N) + (side ? M : 0)) * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int l = 0; l < L; ++l) s += x[(size_t)l * C + c];
        out[(size_t)p * 2 * C + side * C + c] = s / (float)L;
    }
}

__global__ void sigmoid_kernel(const float *__restrict__ x, float *__restrict__ y, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = 1.0f / (1.0f + expf(-x[i]));
}

// ---------------------------------------------------------------------------------------
// pipelines
// ---------------------------------------------------------------------------------------
static int dec_num_weights(const dpm_decoder_desc *d) { return 2 + d->attention_layers * 18 + 4 + 10 + 8 + 4; }

struct DecW {
    const float *proj_w, *proj_b;
    struct Layer {
        const float *sa_in_w, *sa_in_b, *sa_out_w, *sa_out_b, *ca_in_w, *ca_in_b, *ca_out_w, *ca_out_b;
        const float *m0_w, *m0_b, *m2_w, *m2_b, *n1_w, *n1_b, *n2_w, *n2_b, *n3_w, *n3_b;
    } layer[16];
    const float *sim0_w, *sim0_b, *sim2_w, *sim2_b;
    const float *off0_w, *off0_b, *off2_w, *off2_b, *off4_w, *off4_b, *offd_w, *offd_b, *offh_w, *offh_b;
    const float *lp0_w, *lp0_b, *lp2_w, *lp2_b, *lq0_w, *lq0_b, *lq2_w, *lq2_b;
    const float *dim_t;  // extra (last) entry: positional-embedding frequencies
};

static void dec_bind(const dpm_decoder_desc *d, const float *const *w, DecW &o) {
    int i = 0;
    o.proj_w = w[i++]; o.proj_b = w[i++];
    for (int l = 0; l < d->attention_layers; ++l) {
        DecW::Layer &L = o.layer[l];
        L.sa_in_w = w[i++]; L.sa_in_b = w[i++]; L.sa_out_w = w[i++]; L.sa_out_b = w[i++];
        L.ca_in_w = w[i++]; L.ca_in_b = w[i++]; L.ca_out_w = w[i++]; L.ca_out_b = w[i++];
        L.m0_w = w[i++]; L.m0_b = w[i++]; L.m2_w = w[i++]; L.m2_b = w[i++];
        L.n1_w = w[i++]; L.n1_b = w[i++]; L.n2_w = w[i++]; L.n2_b = w[i++]; L.n3_w = w[i++]; L.n3_b = w[i++];
    }
    o.sim0_w = w[i++]; o.sim0_b = w[i++]; o.sim2_w = w[i++]; o.sim2_b = w[i++];
    o.off0_w = w[i++]; o.off0_b = w[i++]; o.off2_w = w[i++]; o.off2_b = w[i++]; o.off4_w = w[i++]; o.off4_b = w[i++];
    o.offd_w = w[i++]; o.offd_b = w[i++]; o.offh_w = w[i++]; o.offh_b = w[i++];
    o.lp0_w = w[i++]; o.lp0_b = w[i++]; o.lp2_w = w[i++]; o.lp2_b = w[i++];
    o.lq0_w = w[i++]; o.lq0_b = w[i++]; o.lq2_w = w[i++]; o.lq2_b = w[i++];
    i += 4;  // coarse_pairing_head: training only (decoder.py:46-49)
    o.dim_t = w[i++];
}

// pre-split (hi/lo tf32) copies of the weights this call pushes through the tensor-core GEMM
static int dec_split(const dpm_decoder_desc *d, const DecW *w, bool registration, Arena &a, cudaStream_t st) {
    const int C = d->model_channel, Cf = d->in_channel;
    split_begin();
    split_add(a, w ? w->proj_w : nullptr, C, Cf, Cf);
    for (int l = 0; l < d->attention_layers; ++l) {
        const DecW::Layer *L = w ? &w->layer[l] : nullptr;
        split_add(a, L ? L->sa_in_w : nullptr, 3 * C, C, C);
        split_add(a, L ? L->sa_out_w : nullptr, C, C, C);
        split_add(a, L ? L->ca_in_w : nullptr, 3 * C, C, C);
        split_add(a, L ? L->ca_out_w : nullptr, C, C, C);
        split_add(a, L ? L->m0_w : nullptr, C, C, C);
        split_add(a, L ? L->m2_w : nullptr, C, C, C);
    }
    if (registration) {
        split_add(a, w ? w->sim0_w : nullptr, C, C, C);
        split_add(a, w ? w->sim2_w : nullptr, C, C, C);
        split_add(a, w ? w->off0_w : nullptr, C, 2 * C, 2 * C);
        split_add(a, w ? w->off2_w : nullptr, C / 2, C, C);
        split_add(a, w ? w->offd_w : nullptr, C / 4, 2 * C, 2 * C);
        split_add(a, w ? w->off4_w : nullptr, C / 4, C / 2, C / 2);
    } else {
        split_add(a, w ? w->lp0_w : nullptr, C, C, C);
        split_add(a, w ? w->lp2_w : nullptr, C, C, C);
        split_add(a, w ? w->lq0_w : nullptr, 2 * C, 2 * C, 2 * C);
    }
    if (!a.ok()) return fail(DPM_ERR_WORKSPACE, "decoder: workspace too small");
    if (!a.dry) DPM_TRY(split_run(st));
    return DPM_OK;
}

// _descriptor_attention_forward (decoder.py:145-162): -> F (R, C) correlated features, xyz (R)
static int attention_stack(const dpm_decoder_desc *d, const DecW &w, const float *src, const float *dst,
                           const uint8_t *src_pad, const uint8_t *dst_pad, int P, int M, int N, Arena &a, float **F_out,
                           float4 **xyz_out, cudaStream_t st) {
    const bool dry = a.dry;
    const int C = d->model_channel, Cf = d->in_channel, H = d->heads;
    const int R = P * (M + N);
    if (C != H * 32) return fail(DPM_ERR_UNSUPPORTED, "decoder: head_dim must be 32 (model_channel=%d heads=%d)", C, H);
    if (d->attention_layers < 0 || d->attention_layers > 16) return fail(DPM_ERR_SHAPE, "decoder: attention_layers");
    float *fea = a.get<float>((size_t)R * Cf);
    float4 *xyz = a.get<float4>((size_t)R);
    float *pos = a.get<float>((size_t)R * C);
    float *x = a.get<float>((size_t)R * C);
    float *y = a.get<float>((size_t)R * C);
    float *qkv = a.get<float>((size_t)R * 3 * C);
    float *att = a.get<float>((size_t)R * C);
    uint8_t *rowmask = a.get<uint8_t>((size_t)R);  // key_padding_mask per token row (only read when a mask was given)
    if (!a.ok()) return fail(DPM_ERR_WORKSPACE, "decoder: workspace too small");
    const uint8_t *km = (src_pad || dst_pad) ? rowmask : nullptr;
    *F_out = x;
    *xyz_out = xyz;
    if (dry) return DPM_OK;
    const int npf = C / 3 / 2 * 2;
    const int maxL = M > N ? M : N;
    dim3 g((maxL + 31) / 32, (Cf + 31) / 32, 2 * P);
    dec_unpack_kernel<<<g, 256, 0, st>>>(src, dst, M, N, Cf, fea, xyz, src_pad, dst_pad, km ? rowmask : nullptr);
    DPM_CHECK_LAUNCH("dec_unpack", st);
    DPM_TRY(posenc_launch(reinterpret_cast<const float *>(xyz), 4, w.dim_t, npf, pos, R, C, st));
    // x = projection(fea) + pos   (the "+ pos" of the first layer, descriptor_attention.py:31)
    DPM_TRY(linear_launch(fea, Cf, w.proj_w, Cf, w.proj_b, pos, C, x, C, R, C, Cf, DPM_ACT_NONE, st));
    for (int l = 0; l < d->attention_layers; ++l) {
        const DecW::Layer &L = w.layer[l];
        const bool last = l == d->attention_layers - 1;
        // self attention (shared weights for src and dst), add & norm1, then "+ pos" again
        DPM_TRY(linear_launch(x, C, L.sa_in_w, C, L.sa_in_b, nullptr, 0, qkv, 3 * C, R, 3 * C, C, DPM_ACT_NONE, st));
        DPM_TRY(attention_launch(qkv, 3 * C, qkv + C, 3 * C, qkv + 2 * C, 3 * C, att, C, nullptr, 2 * P, maxL, M, N, 0, H, st, km));
        DPM_TRY(linear_ln_launch(att, C, L.sa_out_w, C, L.sa_out_b, x, C, L.n1_w, L.n1_b, pos, C, y, x, C, R, C, C, DPM_ACT_NONE, st));
        // cross attention both ways, add & norm2
        DPM_TRY(linear_launch(x, C, L.ca_in_w, C, L.ca_in_b, nullptr, 0, qkv, 3 * C, R, 3 * C, C, DPM_ACT_NONE, st));
        DPM_TRY(attention_launch(qkv, 3 * C, qkv + C, 3 * C, qkv + 2 * C, 3 * C, att, C, nullptr, 2 * P, maxL, M, N, 1, H, st, km));
        DPM_TRY(linear_ln_launch(att, C, L.ca_out_w, C, L.ca_out_b, x, C, L.n2_w, L.n2_b, nullptr, 0, y, x, C, R, C, C, DPM_ACT_NONE, st));
        // mlp, add & norm3 (+ pos for the next layer)
        DPM_TRY(linear_launch(x, C, L.m0_w, C, L.m0_b, nullptr, 0, att, C, R, C, C, DPM_ACT_RELU, st));
        DPM_TRY(linear_ln_launch(att, C, L.m2_w, C, L.m2_b, x, C, L.n3_w, L.n3_b, last ? nullptr : pos, C, y, x, C, R, C, C,
                                 DPM_ACT_NONE, st));
    }
    return DPM_OK;
}

static int registration_run(const dpm_decoder_desc *d, const float *const *weights, const float *src,
                            const float *dst, const uint8_t *src_pad, const uint8_t *dst_pad, int P, int M, int N, int k,
                            float *result, float *conf_out, Arena &a, cudaStream_t st) {
    const bool dry = a.dry;
    if (P <= 0 || M <= 0 || N <= 0) return fail(DPM_ERR_SHAPE, "registration: bad shape P=%d M=%d N=%d", P, M, N);
    if (k <= 0 || (long long)k > (long long)M * N) return fail(DPM_ERR_SHAPE, "registration: k=%d pairs out of range", k);
    if (k > PAIR_MAXK || 2 * k > KAB_MAX) return fail(DPM_ERR_UNSUPPORTED, "registration: k=%d exceeds the limit %d", k, KAB_MAX / 2);
    DecW w;
    if (!dry) dec_bind(d, weights, w);
    const int C = d->model_channel, R = P * (M + N), K2 = 2 * k;
    float *F = nullptr;
    float4 *xyz = nullptr;
    if (d->attention_layers < 0 || d->attention_layers > 16) return fail(DPM_ERR_SHAPE, "decoder: attention_layers");
    DPM_TRY(dec_split(d, dry ? nullptr : &w, true, a, st));
    DPM_TRY(attention_stack(d, w, src, dst, src_pad, dst_pad, P, M, N, a, &F, &xyz, st));
    float *h = a.get<float>((size_t)R * C);
    float *sim = a.get<float>((size_t)R * C);
    float *S = a.get<float>((size_t)P * M * N);
    float2 *rs = a.get<float2>((size_t)P * M);
    float2 *cs = a.get<float2>((size_t)P * N);
    void *pws = a.get<unsigned char>(pairing_ws_bytes(P));
    int32_t *si = a.get<int32_t>((size_t)P * k);
    int32_t *di = a.get<int32_t>((size_t)P * k);
    float *conf = a.get<float>((size_t)P * k);
    float *X = a.get<float>((size_t)P * K2 * 2 * C);
    float *o1 = a.get<float>((size_t)P * K2 * C);
    float *o2 = a.get<float>((size_t)P * K2 * (C / 2));
    float *oi = a.get<float>((size_t)P * K2 * (C / 4));
    float *o3 = a.get<float>((size_t)P * K2 * (C / 4));
    float *off = a.get<float>((size_t)P * K2 * 3);
    float *csrc = a.get<float>((size_t)P * 3 * K2);
    float *cdst = a.get<float>((size_t)P * 3 * K2);
    float *cw = a.get<float>((size_t)P * K2);
    int32_t *cnt = a.get<int32_t>((size_t)P);
    if (!a.ok()) return fail(DPM_ERR_WORKSPACE, "registration: workspace too small");
    if (dry) return DPM_OK;

    // similarity head + L2 normalise (decoder.py:181-185)
    DPM_TRY(linear_launch(F, C, w.sim0_w, C, w.sim0_b, nullptr, 0, h, C, R, C, C, DPM_ACT_RELU, st));
    DPM_TRY(linear_launch(h, C, w.sim2_w, C, w.sim2_b, nullptr, 0, sim, C, R, C, C, DPM_ACT_NONE, st));
    l2norm_rows_kernel<<<(R + 7) / 8, 256, 0, st>>>(sim, R, C);
    DPM_CHECK_LAUNCH("l2norm_rows", st);
    // S_p = A_src . A_dst^T
    DPM_TRY(linear_batched_launch(sim, C, (long long)(M + N) * C, sim + (size_t)M * C, C, (long long)(M + N) * C, nullptr,
                                  nullptr, 0, S, N, (long long)M * N, M, N, C, P, DPM_ACT_NONE, st));
    DPM_TRY(pairing_launch(S, P, M, N, d->tau, k, rs, cs, pws, si, di, conf, st));
    // offset head on [f_s;f_d] and [f_d;f_s]  (heads.py:22-42)
    pair_gather_kernel<<<dim3(K2, P, 1), 256, 0, st>>>(F, C, M, N, k, si, di, X);
    DPM_CHECK_LAUNCH("pair_gather", st);
    const int RO = P * K2;
    DPM_TRY(linear_launch(X, 2 * C, w.off0_w, 2 * C, w.off0_b, nullptr, 0, o1, C, RO, C, 2 * C, DPM_ACT_RELU, st));
    DPM_TRY(linear_launch(o1, C, w.off2_w, C, w.off2_b, nullptr, 0, o2, C / 2, RO, C / 2, C, DPM_ACT_RELU, st));
    DPM_TRY(linear_launch(X, 2 * C, w.offd_w, 2 * C, w.offd_b, nullptr, 0, oi, C / 4, RO, C / 4, 2 * C, DPM_ACT_NONE, st));
    DPM_TRY(linear_launch(o2, C / 2, w.off4_w, C / 2, w.off4_b, oi, C / 4, o3, C / 4, RO, C / 4, C / 2, DPM_ACT_RELU, st));
    DPM_TRY(linear_launch(o3, C / 4, w.offh_w, C / 4, w.offh_b, nullptr, 0, off, 3, RO, 3, C / 4, DPM_ACT_NONE, st));
    corres_kernel<<<P, 256, 0, st>>>(xyz, off, si, di, conf, M, N, k, d->eps_offset * d->eps_offset, csrc, cdst, cw, cnt);
    DPM_CHECK_LAUNCH("corres", st);
    DPM_TRY(kabsch_launch(csrc, cdst, cw, cnt, P, K2, result, nullptr, conf_out, st));
    return DPM_OK;
}

static int loop_run(const dpm_decoder_desc *d, const float *const *weights, const float *src, const float *dst,
                    const uint8_t *src_pad, const uint8_t *dst_pad, int P, int M, int N, float *prob, Arena &a,
                    cudaStream_t st) {
    const bool dry = a.dry;
    if (P <= 0 || M <= 0 || N <= 0) return fail(DPM_ERR_SHAPE, "loop_detection: bad shape");
    DecW w;
    if (!dry) dec_bind(d, weights, w);
    const int C = d->model_channel, R = P * (M + N);
    float *F = nullptr;
    float4 *xyz = nullptr;
    if (d->attention_layers < 0 || d->attention_layers > 16) return fail(DPM_ERR_SHAPE, "decoder: attention_layers");
    DPM_TRY(dec_split(d, dry ? nullptr : &w, false, a, st));
    DPM_TRY(attention_stack(d, w, src, dst, src_pad, dst_pad, P, M, N, a, &F, &xyz, st));
    float *h = a.get<float>((size_t)R * C);
    float *g = a.get<float>((size_t)R * C);
    float *mean = a.get<float>((size_t)P * 2 * C);
    float *p1 = a.get<float>((size_t)P * 2 * C);
    float *logit = a.get<float>((size_t)P);
    if (!a.ok()) return fail(DPM_ERR_WORKSPACE, "loop_detection: workspace too small");
    if (dry) return DPM_OK;
    DPM_TRY(linear_launch(F, C, w.lp0_w, C, w.lp0_b, nullptr, 0, h, C, R, C, C, DPM_ACT_RELU, st));
    DPM_TRY(linear_launch(h, C, w.lp2_w, C, w.lp2_b, nullptr, 0, g, C, R, C, C, DPM_ACT_NONE, st));
    token_mean_kernel<<<dim3(P, 2, 1), 256, 0, st>>>(g, C, M, N, mean);
    DPM_CHECK_LAUNCH("token_mean", st);
    DPM_TRY(linear_launch(mean, 2 * C, w.lq0_w, 2 * C, w.lq0_b, nullptr, 0, p1, 2 * C, P, 2 * C, 2 * C, DPM_ACT_RELU, st));
    DPM_TRY(linear_launch(p1, 2 * C, w.lq2_w, 2 * C, w.lq2_b, nullptr, 0, logit, 1, P, 1, 2 * C, DPM_ACT_NONE, st));
    sigmoid_kernel<<<(P + 127) / 128, 128, 0, st>>>(logit, prob, P);
    DPM_CHECK_LAUNCH("sigmoid", st);
    return DPM_OK;
}

}  // namespace dpm

using namespace dpm;

/* weights = the decoder state_dict tensors in state_dict order PLUS one trailing entry:
 * dim_t (model_channel/3/2*2 floats), the positional-embedding frequencies. */
extern "C" int dpm_decoder_num_weights(const dpm_decoder_desc *desc) { return desc ? dec_num_weights(desc) + 1 : 0; }

extern "C" size_t dpm_registration_workspace_bytes(const dpm_decoder_desc *desc, int P, int M, int N, int k) {
    if (!desc) return 0;
    Arena a(nullptr, 0);
    if (registration_run(desc, nullptr, nullptr, nullptr, nullptr, nullptr, P, M, N, k, nullptr, nullptr, a, nullptr) != DPM_OK) return 0;
    // loop detection shares the attention stack; its extra buffers are smaller than registration's
    return a.off + 256;
}

extern "C" int dpm_registration_forward(const dpm_decoder_desc *desc, const float *const *weights, int n_weights,
                                        const float *src, const float *dst, const uint8_t *src_pad,
                                        const uint8_t *dst_pad, int P, int M, int N, int k, float *result,
                                        float *conf_out, void *ws, size_t ws_bytes, dpm_stream_t stream) {
    if (!desc || !weights || !src || !dst || !result || !conf_out || !ws) return fail(DPM_ERR_ARG, "registration: null pointer");
    if (n_weights != dec_num_weights(desc) + 1)
        return fail(DPM_ERR_SHAPE, "registration: got %d weight tensors, expected %d", n_weights, dec_num_weights(desc) + 1);
    prof_mark((cudaStream_t)stream);
    Arena a(ws, ws_bytes);
    return registration_run(desc, weights, src, dst, src_pad, dst_pad, P, M, N, k, result, conf_out, a, (cudaStream_t)stream);
}

extern "C" size_t dpm_loop_detection_workspace_bytes(const dpm_decoder_desc *desc, int P, int M, int N) {
    if (!desc) return 0;
    Arena a(nullptr, 0);
    if (loop_run(desc, nullptr, nullptr, nullptr, nullptr, nullptr, P, M, N, nullptr, a, nullptr) != DPM_OK) return 0;
    return a.off + 256;
}

extern "C" int dpm_loop_detection_forward(const dpm_decoder_desc *desc, const float *const *weights, int n_weights,
                                          const float *src, const float *dst, const uint8_t *src_pad,
                                          const uint8_t *dst_pad, int P, int M, int N, float *prob, void *ws,
                                          size_t ws_bytes, dpm_stream_t stream) {
    if (!desc || !weights || !src || !dst || !prob || !ws) return fail(DPM_ERR_ARG, "loop_detection: null pointer");
    if (n_weights != dec_num_weights(desc) + 1)
        return fail(DPM_ERR_SHAPE, "loop_detection: got %d weight tensors, expected %d", n_weights, dec_num_weights(desc) + 1);
    prof_mark((cudaStream_t)stream);
    Arena a(ws, ws_bytes);
    return loop_run(desc, weights, src, dst, src_pad, dst_pad, P, M, N, prob, a, (cudaStream_t)stream);
}

extern "C" int dpm_posenc_f32(const float *xyz, int ldx, const float *dim_t, int npf, float *emb, int R, int C,
                              dpm_stream_t stream) {
    if (!xyz || !dim_t || !emb) return fail(DPM_ERR_ARG, "posenc: null pointer");
    return posenc_launch(xyz, ldx, dim_t, npf, emb, R, C, (cudaStream_t)stream);
}

extern "C" int dpm_attention_f32(const float *q, int ldq, const float *k, int ldk, const float *v, int ldv, float *out,
                                 int ldo, const int *prob, int nprob, int max_lq, int heads, dpm_stream_t stream) {
    if (!q || !k || !v || !out || !prob) return fail(DPM_ERR_ARG, "attention: null pointer");
    return attention_launch(q, ldq, k, ldk, v, ldv, out, ldo, prob, nprob, max_lq, 0, 0, 0, heads, (cudaStream_t)stream);
}

/* the attention core in the decoder's own layout: P pairs, rows p*(M+N).. = M src then N dst tokens; mode 0 self,
 * 1 cross; kmask optional (P*(M+N) bytes); impl 0 = as the decoder chooses, 1 = mma.sync kernel, 2 = tcgen05 kernel */
extern "C" int dpm_attention_pairs_f32(const float *q, int ldq, const float *k, int ldk, const float *v, int ldv,
                                       float *out, int ldo, int P, int M, int N, int mode, int heads,
                                       const uint8_t *kmask, int impl, dpm_stream_t stream) {
    if (!q || !k || !v || !out) return fail(DPM_ERR_ARG, "attention: null pointer");
    if (P <= 0 || M <= 0 || N <= 0) return fail(DPM_ERR_SHAPE, "attention: bad shape");
    if (impl == 2) return attention_tc5_launch(q, ldq, k, ldk, v, ldv, out, ldo, P, M, N, mode, heads, kmask, (cudaStream_t)stream);
    if (impl == 1) {
        dim3 grid(((M > N ? M : N) + AT_BQ - 1) / AT_BQ, heads, 2 * P);
        attention_tc_kernel<true><<<grid, AT_T, 0, (cudaStream_t)stream>>>(q, ldq, k, ldk, v, ldv, out, ldo, nullptr, M, N, mode, kmask);
        DPM_CHECK_LAUNCH("attention", (cudaStream_t)stream);
        return DPM_OK;
    }
    return attention_launch(q, ldq, k, ldk, v, ldv, out, ldo, nullptr, 2 * P, M > N ? M : N, M, N, mode, heads, (cudaStream_t)stream, kmask);
}

extern "C" int dpm_kabsch_f32(const float *src, const float *dst, const float *w, const int32_t *count, int P, int ldk,
                              float *result, uint8_t *inlier, float *conf_out, dpm_stream_t stream) {
    if (!src || !dst || !w || !count || !result) return fail(DPM_ERR_ARG, "kabsch: null pointer");
    return kabsch_launch(src, dst, w, count, P, ldk, result, inlier, conf_out, (cudaStream_t)stream);
}
