// attention_tc5.cu -- the multi-head attention core (head_dim 32) for LONG key ranges (map tiles: up to 4096
// descriptors, system/modules/pose_graph.py:513) on the 5th-generation tensor cores: flash attention with S = Q K^T
// and O_tile = P V as tcgen05.mma kind::tf32 with fp32 accumulators in TMEM, 3xTF32 error compensation
// (hi*hi | hi*lo + lo*hi in separate accumulators) like gemm_tc.cu, online softmax in registers.
// Replaces nn.MultiheadAttention's scaled-dot-product core in DescriptorAttentionLayer
// (network/decoder/descriptor_attention.py:24-51) when a side of the pair has more than 512 descriptors; the
// register-resident mma.sync kernel of decoder.cu stays for the 256-descriptor odometry pairs, where it wins.
//
// CTA = 128 queries x 1 head, key tiles of 64:
//   warps 0-7  softmax: thread <-> (query row = TMEM lane, one half of the tile's 64 keys); tcgen05.ld of the
//              scores, running max / sum, P = exp(S - max) split into hi / lo and stored as the A operand of P V in
//              the canonical K-major SWIZZLE_128B layout; folds the PREVIOUS tile's O_tile (read back from TMEM) into
//              the running output with one IEEE fma per element.  Every tile accumulates into a FRESH TMEM
//              accumulator: the tensor core's fp32 accumulate truncates, and a 4096-key row would otherwise carry a
//              2e-5 bias (measured; see decoder.cu).
//   warps 8-11 loaders: K tile -> [K_hi ; K_lo] (128 rows x 32, the B operand of S: one MMA yields Q_hi K_hi^T and
//              Q_hi K_lo^T side by side), V tile -> TRANSPOSED [V_hi^T ; V_lo^T] (64 rows x 64 keys, the B operand
//              of P V); a ring of two K stages and one of two V stages.
//   warp 12    one lane issues the MMAs: S(i+2) is in flight while the softmax warps work on tile i and P V (i)
//              runs as soon as P(i) is in shared memory; completion goes through tcgen05.commit -> mbarriers.
// Measured and rejected: 16 softmax warps (4 per TMEM lane quarter, 16 keys each) instead of 8 -- 0.359 against 0.335 ms
// at 4096 x 4096: the max / exp phase does not shrink (1 315 cycles per tile), because all softmax warps wait for the same
// S tile and then hit the special-function unit together (8 448 ex2 per tile at 16 per clock = 530 cycles) behind a wider
// barrier.  What would help is two query tiles per CTA in ping-pong (one group's exp under the other's MMAs).
// Also measured and rejected: the 8 softmax warps as two ping-pong groups over even / odd key tiles (thread = whole row of 64
// keys, own running max / sum / output per group, merged at the end, four O buffers): bit-for-bit fine, 0.350 ms -- the
// groups then wait ~1 400 cycles per tile for S(i+2), which cannot be issued before P V (i) because it overwrites P(i)
// in tensor memory, and the 512 TMEM columns are all in use (S / P 2 x 128, O 4 x 64).
// P never touches shared memory: the softmax warps write it back into TENSOR MEMORY over the scores it came from
// (tcgen05.st) and P V takes its A operand from there -- a first version that staged P in shared memory (64 KB per
// tile written, 64 KB read by the MMAs) was bound by the shared-memory pipe: 3 400 cycles per tile, none of them
// waiting for the tensor core.  S / P are double-buffered and O_tile(i-1) is read back AFTER P(i) has been handed over.
// Per tile: 8 MMAs for S (N = 128 / 64) and 16 for P V (N = 64 / 32), K = 8 each.
#include "common.cuh"
#include "tc5.cuh"

namespace dpm {
namespace at5 {

using namespace tc;

constexpr int BQ = 128, BKEY = 64, NSTAGE = 2;
constexpr int THREADS = 13 * 32;
constexpr int Q_BYTES = 2 * BQ * 128;            // Q_hi, Q_lo
constexpr int K_BYTES = 2 * BKEY * 128;          // [K_hi ; K_lo]
constexpr int V_BYTES = 2 * 2 * 32 * 128;        // 2 chunks of [V_hi^T ; V_lo^T]
constexpr int OFF_K = Q_BYTES, OFF_V = OFF_K + NSTAGE * K_BYTES;
constexpr int OFF_MISC = OFF_V + NSTAGE * V_BYTES;   // mbarriers, TMEM base, row-max exchange, key masks
constexpr int SMEM_BYTES = OFF_MISC + 128 + 2 * 2 * BQ * 4 + 4 * BKEY + 1024;  // + room to align the base to 1024 bytes
constexpr unsigned TM_S = 0, TM_O = 256;         // TMEM columns: S buffers 2 x (64 + 64), O buffers 2 x (32 + 32)
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

#ifdef DPM_AT5_PROFILE
__device__ unsigned long long at5_prof[16];
#define AT5_TICK(i) do { const long long _t = clock64(); pacc[i] += _t - tprev; tprev = _t; } while (0)
#else
#define AT5_TICK(i) do { } while (0)
#endif

// 32 lanes x 32 consecutive fp32 columns, thread t of the warp writes row (lane base + t)
__device__ __forceinline__ void tmem_st32(unsigned taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]),
          "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]),
          "f"(v[18]), "f"(v[19]), "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]),
          "f"(v[27]), "f"(v[28]), "f"(v[29]), "f"(v[30]), "f"(v[31])
        : "memory");
}
// 2^x on the special-function unit, one instruction (x <= 0 here; -inf -> 0; relative error 2^-22)
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (128 rows x 8 tf32) is read from tensor memory, lane = row
__device__ __forceinline__ void mma_tf32_ts(unsigned d_tmem, unsigned a_tmem, unsigned long long bdesc, unsigned idesc,
                                            unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void tmem_ld16(unsigned taddr, float (&v)[16]) {
    unsigned r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__global__ void __launch_bounds__(THREADS, 1)
attention_tc5_kernel(const float *__restrict__ Q, int ldq, const float *__restrict__ Kp, int ldk,
                     const float *__restrict__ Vp, int ldv, float *__restrict__ O, int ldo, int M, int N, int mode,
                     const uint8_t *__restrict__ kmask) {
    extern __shared__ __align__(1024) unsigned char smem[];  // used directly: the compiler keeps the shared address space
    unsigned char *sQ = smem, *sK = smem + OFF_K, *sV = smem + OFF_V;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + OFF_MISC);       // 14 mbarriers
    unsigned &tmem_base_s = *reinterpret_cast<unsigned *>(smem + OFF_MISC + 120);
    float (*xm)[2][BQ] = reinterpret_cast<float (*)[2][BQ]>(smem + OFF_MISC + 128);           // [tile parity][key half][row]
    uint8_t (*kdrop)[BKEY] = reinterpret_cast<uint8_t (*)[BKEY]>(smem + OFF_MISC + 128 + 2 * 2 * BQ * 4);  // [tile & 3][key]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int z = blockIdx.z, head = blockIdx.y;
    const int p = z >> 1, side = z & 1, base = p * (M + N);
    const int q0 = base + (side ? M : 0), Lq = side ? N : M;
    const int kvside = mode ? !side : side;  // mode 0: self, 1: cross
    const int k0 = base + (kvside ? M : 0), Lk = kvside ? N : M;
    if (blockIdx.x * BQ >= Lq) return;       // block-uniform, before any barrier / allocation
    const int nt = (Lk + BKEY - 1) / BKEY;

    const unsigned k_full = s32(&bars[0]), k_empty = s32(&bars[2]), v_full = s32(&bars[4]), v_empty = s32(&bars[6]),
                   s_full = s32(&bars[8]), o_full = s32(&bars[10]), p_full = s32(&bars[12]);
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(k_full + 8 * b, 128);
            mbar_init(v_full + 8 * b, 128);
            mbar_init(k_empty + 8 * b, 1);
            mbar_init(v_empty + 8 * b, 1);
            mbar_init(s_full + 8 * b, 1);
            mbar_init(o_full + 8 * b, 1);
            mbar_init(p_full + 8 * b, 256);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 12) tmem_alloc(s32(&tmem_base_s), 512);
    // Q tile: scaled by 1/sqrt(d) (as nn.MultiheadAttention does), split hi / lo, canonical layout
    for (int e = tid; e < BQ * 8; e += THREADS) {
        const int r = e >> 3, c = e & 7;
        const int row = min(blockIdx.x * BQ + r, Lq - 1);
        float4 v = *reinterpret_cast<const float4 *>(Q + (size_t)(q0 + row) * ldq + head * 32 + 4 * c);
        const float sc = 0.17677669529663687f * 1.4426950408889634f;  // 1/sqrt(d) (nn.MultiheadAttention) x log2(e): scores in log2 units
        v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
        float4 h, l;
        split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
        const unsigned off = swz(r, c);
        *reinterpret_cast<float4 *>(sQ + off) = h;
        *reinterpret_cast<float4 *>(sQ + BQ * 128 + off) = l;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = tmem_base_s;

    if (warp < 8) {
        // ======================= softmax warps =======================
        const int wg = warp >> 2, quarter = warp & 3, row = quarter * 32 + lane, col0 = 32 * wg;
        const unsigned tl = tmem + ((unsigned)(quarter * 32) << 16);
        const float NEG = -__int_as_float(0x7f800000);
        float o_run[16];
#pragma unroll
        for (int d = 0; d < 16; ++d) o_run[d] = 0.f;
        float m_run = NEG, l_run = 0.f, corr_prev = 0.f;
#ifdef DPM_AT5_PROFILE
        long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#endif
        for (int i = 0; i < nt; ++i) {
            const int b = i & 1;
            AT5_TICK(0);
            mbar_wait(s_full + 8 * b, (unsigned)((i >> 1) & 1));
            tc_fence_after();
            AT5_TICK(1);
            float s[32];
            {
                unsigned rh[32], rx[32];  // both loads in flight, one wait
                tmem_ld32_nowait(tl + TM_S + 128u * b + (unsigned)col0, rh);
                tmem_ld32_nowait(tl + TM_S + 128u * b + 64u + (unsigned)col0, rx);
                tmem_wait_ld();
#pragma unroll
                for (int c = 0; c < 32; ++c) s[c] = __uint_as_float(rh[c]) + __uint_as_float(rx[c]);
            }
            AT5_TICK(2);
            const int nvalid = Lk - i * BKEY - col0;  // keys of my half that exist
            if (nvalid < 32) {
#pragma unroll
                for (int c = 0; c < 32; ++c)
                    if (c >= nvalid) s[c] = NEG;
            }
            if (kmask) {
                const uint8_t *kd = kdrop[i & 3] + col0;
#pragma unroll
                for (int c = 0; c < 32; ++c)
                    if (kd[c]) s[c] = NEG;
            }
            float mx4[4] = {NEG, NEG, NEG, NEG};  // four independent chains instead of one of 32 dependent fmax
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                mx4[0] = fmaxf(mx4[0], s[c]); mx4[1] = fmaxf(mx4[1], s[c + 1]);
                mx4[2] = fmaxf(mx4[2], s[c + 2]); mx4[3] = fmaxf(mx4[3], s[c + 3]);
            }
            const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
            xm[b][wg][row] = mx;
            asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");  // the two warps that share these rows
            // (slot b is rewritten two tiles later, i.e. after another barrier that the partner only reaches once it has read it)
            float m_new = fmaxf(m_run, fmaxf(mx, xm[b][wg ^ 1][row]));
            if (m_new == NEG) m_new = 0.f;  // every key so far masked: keep the exponents finite
            const float corr = ex2(m_run - m_new);
            float ls4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                s[c] = ex2(s[c] - m_new);         ls4[0] += s[c];
                s[c + 1] = ex2(s[c + 1] - m_new); ls4[1] += s[c + 1];
                s[c + 2] = ex2(s[c + 2] - m_new); ls4[2] += s[c + 2];
                s[c + 3] = ex2(s[c + 3] - m_new); ls4[3] += s[c + 3];
            }
            l_run = fmaf(l_run, corr, (ls4[0] + ls4[1]) + (ls4[2] + ls4[3]));
            m_run = m_new;
            AT5_TICK(3);
            // P (my 32 keys) -> the A operand of P V, IN TENSOR MEMORY and in place of the scores it was computed from:
            // the hi parts over the S_hi columns, the lo parts over the cross-term columns (lane = query row, column =
            // key: exactly the layout a TMEM A operand has).  S(i+2) is issued after P V (i), so it cannot overwrite them early.
            {
                float lo[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    float h;
                    split_tf32(s[c], h, lo[c]);
                    s[c] = h;
                }
                tmem_st32(tl + TM_S + 128u * b + (unsigned)col0, s);
                tmem_st32(tl + TM_S + 128u * b + 64u + (unsigned)col0, lo);
                tmem_wait_st();
            }
            tc_fence_before();
            mbar_arrive(p_full + 8 * b);
            AT5_TICK(4);
            // off the tensor core's critical path: fold the previous tile's P V into the running output
            if (i > 0) {
                const int pb = (i - 1) & 1;
                mbar_wait(o_full + 8 * pb, (unsigned)(((i - 1) >> 1) & 1));
                tc_fence_after();
                AT5_TICK(5);
                float ot[16], ox[16];
                tmem_ld16(tl + TM_O + 64u * pb + 16u * wg, ot);
                tmem_ld16(tl + TM_O + 64u * pb + 32u + 16u * wg, ox);
                tc_fence_before();
#pragma unroll
                for (int d = 0; d < 16; ++d) o_run[d] = fmaf(o_run[d], corr_prev, ot[d] + ox[d]);
            }
            corr_prev = corr;
            AT5_TICK(6);
        }
#ifdef DPM_AT5_PROFILE
        if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
            for (int e = 0; e < 8; ++e) at5_prof[e] = (unsigned long long)pacc[e];
            at5_prof[8] = (unsigned long long)nt;
        }
#endif
        {   // last tile
            const int pb = (nt - 1) & 1;
            mbar_wait(o_full + 8 * pb, (unsigned)(((nt - 1) >> 1) & 1));
            tc_fence_after();
            float ot[16], ox[16];
            tmem_ld16(tl + TM_O + 64u * pb + 16u * wg, ot);
            tmem_ld16(tl + TM_O + 64u * pb + 32u + 16u * wg, ox);
#pragma unroll
            for (int d = 0; d < 16; ++d) o_run[d] = fmaf(o_run[d], corr_prev, ot[d] + ox[d]);
        }
        // both halves hold partial row sums under the same running max
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");  // the partner is past its last read of xm
        xm[0][wg][row] = l_run;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
        const float l = l_run + xm[0][wg ^ 1][row];
        const int qrow = blockIdx.x * BQ + row;
        if (qrow < Lq) {
            float *o = O + (size_t)(q0 + qrow) * ldo + head * 32 + 16 * wg;
#pragma unroll
            for (int d = 0; d < 16; d += 4)
                *reinterpret_cast<float4 *>(o + d) = make_float4(o_run[d] / l, o_run[d + 1] / l, o_run[d + 2] / l, o_run[d + 3] / l);
        }
    } else if (warp < 12) {
        // ======================= loaders =======================
        const int ltid = tid - 256;
        for (int i = 0; i < nt; ++i) {
            const int s = i & 1;
            const unsigned par = (unsigned)(((i >> 1) - 1) & 1);
            float4 kk[4], vv[4];
#pragma unroll
            for (int it = 0; it < 4; ++it) {  // global loads first: they do not need the stage
                const int idx = ltid + it * 128, key = idx >> 3, c = idx & 7;
                const int gk = i * BKEY + key;
                kk[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                vv[it] = kk[it];
                if (gk < Lk) {
                    const size_t r = (size_t)(k0 + gk);
                    kk[it] = *reinterpret_cast<const float4 *>(Kp + r * ldk + head * 32 + 4 * c);
                    vv[it] = *reinterpret_cast<const float4 *>(Vp + r * ldv + head * 32 + 4 * c);
                }
            }
            if (i >= 2) mbar_wait(k_empty + 8 * s, par);   // S(i-2) has read this K stage
            unsigned char *kst = sK + (size_t)s * K_BYTES;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int idx = ltid + it * 128, key = idx >> 3, c = idx & 7;
                float4 h, l;
                split_tf32(kk[it].x, h.x, l.x); split_tf32(kk[it].y, h.y, l.y);
                split_tf32(kk[it].z, h.z, l.z); split_tf32(kk[it].w, h.w, l.w);
                const unsigned off = swz(key, c);
                *reinterpret_cast<float4 *>(kst + off) = h;                 // rows 0..63   = K_hi
                *reinterpret_cast<float4 *>(kst + BKEY * 128 + off) = l;    // rows 64..127 = K_lo
                // the softmax warps read the mask of tile i after S(i), i.e. after k_full(i); slot (i & 3) was last
                // read for tile i - 4, whose P the tensor core consumed before S(i - 2) -- hence before this stage was freed
                if (kmask && c == 0) {
                    const int gk = i * BKEY + key;
                    kdrop[i & 3][key] = gk < Lk ? kmask[k0 + gk] : 1;
                }
            }
            fence_proxy_async();
            mbar_arrive(k_full + 8 * s);
            if (i >= 2) mbar_wait(v_empty + 8 * s, par);   // P V (i-2) has read this V stage (and kdrop[s] is free)
            unsigned char *vst = sV + (size_t)s * V_BYTES;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int idx = ltid + it * 128, key = idx >> 3, c = idx & 7;
                // V transposed: element (dim n, key) of chunk key / 32; rows 0..31 = V_hi^T, 32..63 = V_lo^T
                unsigned char *vb = vst + (key >> 5) * (64 * 128);
                const int kq = key & 31;
                const float ve[4] = {vv[it].x, vv[it].y, vv[it].z, vv[it].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int n = 4 * c + e;
                    float vh, vl;
                    split_tf32(ve[e], vh, vl);
                    const unsigned o2 = (unsigned)(((n >> 3) << 10) + ((n & 7) << 7) + ((((kq >> 2) ^ n) & 7) << 4) + ((kq & 3) << 2));
                    *reinterpret_cast<float *>(vb + o2) = vh;
                    *reinterpret_cast<float *>(vb + 32 * 128 + o2) = vl;
                }
            }
            fence_proxy_async();
            mbar_arrive(v_full + 8 * s);
        }
    } else if (lane == 0) {
        // ======================= MMA issuer =======================
        constexpr unsigned ID_S2 = instr_desc(128), ID_S1 = instr_desc(64), ID_O2 = instr_desc(64), ID_O1 = instr_desc(32);
        const unsigned long long qh = smem_desc(s32(sQ)), ql = smem_desc(s32(sQ + BQ * 128));
        auto issue_s = [&](int i) {
            const int b = i & 1;
            mbar_wait(k_full + 8 * b, (unsigned)((i >> 1) & 1));
            tc_fence_after();
            const unsigned long long kd = smem_desc(s32(sK + (size_t)b * K_BYTES));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned long long ko = (unsigned long long)(2 * j);  // 32 bytes >> 4 inside the swizzle row
                mma_tf32(tmem + TM_S + 128u * b, qh + ko, kd + ko, ID_S2, j ? 1u : 0u);       // Q_hi [K_hi;K_lo]^T
                mma_tf32(tmem + TM_S + 128u * b + 64u, ql + ko, kd + ko, ID_S1, 1u);          // + Q_lo K_hi^T
            }
            mma_commit(s_full + 8 * b);
            mma_commit(k_empty + 8 * b);
        };
        if (nt > 0) issue_s(0);
        if (nt > 1) issue_s(1);
        for (int i = 0; i < nt; ++i) {
            const int b = i & 1;
            const unsigned par = (unsigned)((i >> 1) & 1);
            mbar_wait(v_full + 8 * b, par);
            mbar_wait(p_full + 8 * b, par);
            tc_fence_after();
            const unsigned vbase = s32(sV + (size_t)b * V_BYTES);
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int c = jj >> 2, j = jj & 3;
                const unsigned long long vd = smem_desc(vbase + c * (64 * 128)) + (unsigned long long)(2 * j);
                const unsigned pa = tmem + TM_S + 128u * b + 8u * jj;   // keys 8 jj .. 8 jj + 7 of the tile
                mma_tf32_ts(tmem + TM_O + 64u * b, pa, vd, ID_O2, jj ? 1u : 0u);              // P_hi [V_hi;V_lo]
                mma_tf32_ts(tmem + TM_O + 64u * b + 32u, pa + 64u, vd, ID_O1, 1u);            // + P_lo V_hi
            }
            mma_commit(o_full + 8 * b);
            mma_commit(v_empty + 8 * b);
            if (i + 2 < nt) issue_s(i + 2);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 12) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

}  // namespace at5

// self (mode 0) / cross (mode 1) attention of P (src, dst) pairs laid out as in decoder.cu: rows p*(M+N) .. are the
// M src tokens followed by the N dst tokens.  kmask: optional key-padding mask per token row.
int attention_tc5_launch(const float *q, int ldq, const float *k, int ldk, const float *v, int ldv, float *out, int ldo,
                         int P, int M, int N, int mode, int heads, const uint8_t *kmask, cudaStream_t st) {
    if (P <= 0 || M <= 0 || N <= 0 || heads <= 0) return fail(DPM_ERR_SHAPE, "attention: bad shape");
    if ((ldq | ldk | ldv | ldo) & 3) return fail(DPM_ERR_UNSUPPORTED, "attention: leading dimensions must be multiples of 4");
    static thread_local unsigned long long configured = 0ull;
    const unsigned long long devbit = 1ull << (current_device() & 63);
    if (!(configured & devbit)) {
        DPM_CHECK_CUDA(cudaFuncSetAttribute(at5::attention_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, at5::SMEM_BYTES));
        configured |= devbit;
    }
    const int maxL = M > N ? M : N;
    dim3 grid((maxL + at5::BQ - 1) / at5::BQ, heads, 2 * P);
    at5::attention_tc5_kernel<<<grid, at5::THREADS, at5::SMEM_BYTES, st>>>(q, ldq, k, ldk, v, ldv, out, ldo, M, N, mode, kmask);
    DPM_CHECK_LAUNCH("attention_tc5", st);
    return DPM_OK;
}

}  // namespace dpm

#ifdef DPM_AT5_PROFILE
extern "C" int dpm_debug_at5_profile(unsigned long long *out16) {
    return cudaMemcpyFromSymbol(out16, dpm::at5::at5_prof, sizeof(unsigned long long) * 16) == cudaSuccess ? 0 : -1;
}
#endif
