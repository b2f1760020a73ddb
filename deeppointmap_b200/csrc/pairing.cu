// pairing.cu -- the pairing stage of Decoder.registration_forward (network/decoder/decoder.py:185-192):
//     P = softmax(S / tau, dim=2) * softmax(S / tau, dim=1);  conf, flat = topk(P.flatten(), k)
// over the cosine-similarity matrices S (pairs, M, N) in THREE launches whatever M x N is (256 x 256 odometry pairs,
// 4096 x 256 scan-to-map, 4096 x 4096 map-to-map), every one of them spread over the whole chip:
//
//   1. pair_stats_kernel         row (max, sum exp) and column (max, sum exp) of S / tau in one launch (row blocks and
//                                column blocks side by side); also clears the histogram / counters of step 2 and 3.
//   2. pair_softmax_hist_kernel  P in place over S + a 2048-bin histogram of the top bits of every P value; the LAST
//                                block of a pair (threadfence + counter) turns the histogram into the bin that holds
//                                the k-th largest value.
//   3. pair_select_kernel        every block appends its elements at or above that bin to the pair's candidate list
//                                (typically k + a few hundred entries); the LAST block of a pair radix-selects the k
//                                largest 64-bit keys (value bits, ~flat index) among them, sorts them and emits
//                                (src index, dst index, confidence) in torch.topk's order (value descending; equal
//                                values: lower flat index first).
//
// Deterministic: integer atomics only decide WHERE a candidate is parked, never which candidates there are, and the
// final order is a total order on unique keys.  If a pathological matrix puts more than PAIR_CAND elements into the
// threshold bin the last block falls back to selecting over the whole matrix (slow, exact).
#include "common.cuh"

namespace dpm {

constexpr int PAIR_BINS = 2048;    // bits >> 19 of a fp32 in [0, 1]: 8 exponent + 4 mantissa bits (<= 2032)
constexpr int PAIR_SHIFT = 19;
constexpr int PAIR_CAND = 16384;   // candidate slots per pair
constexpr int PAIR_T2 = 256, PAIR_CHUNK2 = 4096;     // step 2: threads, elements per block
constexpr int PAIR_T3 = 1024, PAIR_CHUNK3 = 16384;   // step 3
constexpr int PAIR_CTL = 8;        // ints per pair: [0] blocks done in step 2, [1] threshold bin, [2] candidates, [3] blocks done in step 3

__global__ void __launch_bounds__(256)
pair_stats_kernel(const float *__restrict__ S, int P, int M, int N, float tau, float2 *__restrict__ rs,
                  float2 *__restrict__ cs, int *__restrict__ hist, int *__restrict__ ctl, int nrowblk, int colblk) {
    __shared__ float smx[8][32], ssum[8][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = blockIdx.x * 256 + tid; i < P * PAIR_BINS; i += gridDim.x * 256) hist[i] = 0;
    for (int i = blockIdx.x * 256 + tid; i < P * PAIR_CTL; i += gridDim.x * 256) ctl[i] = 0;
    const float NEG = -__int_as_float(0x7f800000);
    if ((int)blockIdx.x < nrowblk) {  // ---- rows: one warp per row of (P*M, N) ----
        const int row = blockIdx.x * 8 + warp;
        if (row >= P * M) return;
        const float *s = S + (size_t)row * N;
        float mx = NEG;
        for (int j = lane; j < N; j += 32) mx = fmaxf(mx, __fdiv_rn(s[j], tau));
        mx = warp_max(mx);
        float sum = 0.f;
        for (int j = lane; j < N; j += 32) sum += expf(__fdiv_rn(s[j], tau) - mx);
        sum = warp_sum(sum);
        if (lane == 0) rs[row] = make_float2(mx, sum);
        return;
    }
    // ---- columns: 32 columns x 8 row slices per block ----
    const int b = blockIdx.x - nrowblk, p = b / colblk, j = (b % colblk) * 32 + lane;
    const float *s = S + (size_t)p * M * N;
    float mx = NEG, sum = 0.f;
    if (j < N)
        for (int i = warp; i < M; i += 8) mx = fmaxf(mx, __fdiv_rn(s[(size_t)i * N + j], tau));
    smx[warp][lane] = mx;
    __syncthreads();
    float gm = smx[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) gm = fmaxf(gm, smx[w][lane]);
    if (j < N)
        for (int i = warp; i < M; i += 8) sum += expf(__fdiv_rn(s[(size_t)i * N + j], tau) - gm);
    ssum[warp][lane] = sum;
    __syncthreads();
    if (warp == 0 && j < N) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += ssum[w][lane];
        cs[(size_t)p * N + j] = make_float2(gm, t);
    }
}

// true in every thread of the LAST block of pair `p` to get here (all earlier blocks' global writes are visible)
__device__ __forceinline__ bool pair_last_block(int *counter, int nblocks) {
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(counter, 1) == nblocks - 1;
    __syncthreads();
    if (s_last) __threadfence();
    return s_last != 0;
}

__global__ void __launch_bounds__(PAIR_T2)
pair_softmax_hist_kernel(float *__restrict__ S, int M, int N, float tau, const float2 *__restrict__ rs,
                         const float2 *__restrict__ cs, int k, int *__restrict__ hist, int *__restrict__ ctl) {
    __shared__ int h[PAIR_BINS];
    __shared__ int s_wsum[PAIR_T2 / 32];
    const int p = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    const long long MN = (long long)M * N;
    for (int i = tid; i < PAIR_BINS; i += PAIR_T2) h[i] = 0;
    __syncthreads();
    float *Sp = S + (size_t)p * MN;
    const long long base = (long long)blockIdx.x * PAIR_CHUNK2;
#pragma unroll 4
    for (int it = 0; it < PAIR_CHUNK2 / PAIR_T2; ++it) {
        const long long i = base + it * PAIR_T2 + tid;
        const bool ok = i < MN;
        unsigned bin = PAIR_BINS;
        if (ok) {
            const int row = (int)(i / N), j = (int)(i - (long long)row * N);
            const float x = __fdiv_rn(Sp[i], tau);
            const float2 r = rs[(size_t)p * M + row], c = cs[(size_t)p * N + j];
            const float v = (expf(x - r.x) / r.y) * (expf(x - c.x) / c.y);
            Sp[i] = v;
            bin = __float_as_uint(v) >> PAIR_SHIFT;
            bin = bin < PAIR_BINS ? bin : PAIR_BINS - 1;  // never (v <= 1), but a NaN must not index outside
        }
        const unsigned peers = __match_any_sync(0xffffffffu, bin);
        if (ok && lane == __ffs(peers) - 1) atomicAdd(&h[bin], __popc(peers));
    }
    __syncthreads();
    int *hp = hist + (size_t)p * PAIR_BINS;
    for (int i = tid; i < PAIR_BINS; i += PAIR_T2)
        if (h[i]) atomicAdd(&hp[i], h[i]);
    if (!pair_last_block(&ctl[p * PAIR_CTL + 0], gridDim.x)) return;
    // ---- last block of the pair: the bin T with count(bins > T) < k <= count(bins >= T) ----
    // thread t owns the 8 bins [B - 8(t+1), B - 8t): descending order over t
    constexpr int PER = PAIR_BINS / PAIR_T2;
    int mine[PER], seg = 0;
#pragma unroll
    for (int e = 0; e < PER; ++e) {
        mine[e] = __ldcg(&hp[PAIR_BINS - 1 - (tid * PER + e)]);
        seg += mine[e];
    }
    int incl = seg;  // inclusive scan over the block, thread 0 first
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) s_wsum[tid >> 5] = incl;
    __syncthreads();
    int before = 0;
    for (int w = 0; w < (tid >> 5); ++w) before += s_wsum[w];
    incl += before;
    const int excl = incl - seg;
    if (excl < k && incl >= k) {  // exactly one thread
        int cum = excl, e = 0;
        for (; e < PER - 1; ++e) {
            if (cum + mine[e] >= k) break;
            cum += mine[e];
        }
        ctl[p * PAIR_CTL + 1] = PAIR_BINS - 1 - (tid * PER + e);
    }
}

// the k largest of n unique 64-bit keys -> sel[0..k) sorted descending.  FULL: key i = (matrix[i], ~i) for i < n;
// else the candidate list.  8 radix passes of 8 bits from the top; all PAIR_T3 threads of one block.
template <bool FULL>
__device__ void pair_select(const unsigned long long *__restrict__ cand, const unsigned *__restrict__ mat, long long n,
                            int k, int kp2, unsigned long long *sel, int *hist256, int *s_misc) {
    const int tid = threadIdx.x, lane = tid & 31;
    auto key = [&](long long i) -> unsigned long long {
        if (FULL) return ((unsigned long long)__ldcg(&mat[i]) << 32) | (unsigned)(0xffffffffu - (unsigned)i);
        return __ldcg(&cand[i]);
    };
    unsigned long long prefix = 0ull;
    int remaining = k;
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = 56 - 8 * pass;
        const unsigned long long himask = pass == 0 ? 0ull : (~0ull << (shift + 8));
        if (tid < 256) hist256[tid] = 0;
        __syncthreads();
        for (long long i0 = 0; i0 < n; i0 += PAIR_T3) {
            const long long i = i0 + tid;
            unsigned digit = 256u;
            if (i < n) {
                const unsigned long long q = key(i);
                if ((q & himask) == prefix) digit = (unsigned)(q >> shift) & 255u;
            }
            const unsigned peers = __match_any_sync(0xffffffffu, digit);
            if (digit < 256u && lane == __ffs(peers) - 1) atomicAdd(&hist256[digit], __popc(peers));
        }
        __syncthreads();
        if (tid == 0) {
            int cum = 0, d = 255;
            for (; d > 0; --d) {
                if (cum + hist256[d] >= remaining) break;
                cum += hist256[d];
            }
            s_misc[0] = d;
            s_misc[1] = remaining - cum;
        }
        __syncthreads();
        prefix |= (unsigned long long)s_misc[0] << shift;
        remaining = s_misc[1];
        __syncthreads();
    }
    // prefix = the k-th largest key (keys are unique): everything >= it is a winner
    if (tid == 0) s_misc[2] = 0;
    for (int i = tid; i < kp2; i += PAIR_T3) sel[i] = 0ull;
    __syncthreads();
    for (long long i0 = 0; i0 < n; i0 += PAIR_T3) {
        const long long i = i0 + tid;
        if (i < n) {
            const unsigned long long q = key(i);
            if (q >= prefix) {
                const int pos = atomicAdd(&s_misc[2], 1);
                if (pos < kp2) sel[pos] = q;
            }
        }
    }
    __syncthreads();
    bitonic_sort_desc(sel, kp2, tid, PAIR_T3);
}

__global__ void __launch_bounds__(PAIR_T3)
pair_select_kernel(const float *__restrict__ Pm, int M, int N, int k, int kp2, int *__restrict__ ctl,
                   unsigned long long *__restrict__ cand, int32_t *__restrict__ si, int32_t *__restrict__ di,
                   float *__restrict__ conf) {
    extern __shared__ __align__(16) unsigned char pair_smem[];
    unsigned long long *sel = reinterpret_cast<unsigned long long *>(pair_smem);  // kp2
    __shared__ int hist256[256];
    __shared__ int s_misc[4];
    const int p = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    const long long MN = (long long)M * N;
    const unsigned *v = reinterpret_cast<const unsigned *>(Pm + (size_t)p * MN);
    unsigned long long *cp = cand + (size_t)p * PAIR_CAND;
    int *c = ctl + p * PAIR_CTL;
    const unsigned tbits = (unsigned)c[1] << PAIR_SHIFT;
    const long long base = (long long)blockIdx.x * PAIR_CHUNK3;
#pragma unroll 4
    for (int it = 0; it < PAIR_CHUNK3 / PAIR_T3; ++it) {
        const long long i = base + it * PAIR_T3 + tid;
        const unsigned u = i < MN ? v[i] : 0u;
        const bool take = i < MN && u >= tbits;
        const unsigned bal = __ballot_sync(0xffffffffu, take);
        if (bal) {
            int pos = 0;
            if (lane == __ffs(bal) - 1) pos = atomicAdd(&c[2], __popc(bal));
            pos = __shfl_sync(0xffffffffu, pos, __ffs(bal) - 1) + __popc(bal & ((1u << lane) - 1u));
            if (take && pos < PAIR_CAND) cp[pos] = ((unsigned long long)u << 32) | (unsigned)(0xffffffffu - (unsigned)i);
        }
    }
    if (!pair_last_block(&c[3], gridDim.x)) return;
    const int ncand = __ldcg(&c[2]);
    if (ncand <= PAIR_CAND) pair_select<false>(cp, v, ncand, k, kp2, sel, hist256, s_misc);
    else pair_select<true>(cp, v, MN, k, kp2, sel, hist256, s_misc);
    for (int r = tid; r < k; r += PAIR_T3) {
        const unsigned long long e = sel[r];
        const unsigned flat = 0xffffffffu - (unsigned)e;
        si[(size_t)p * k + r] = (int)(flat / (unsigned)N);
        di[(size_t)p * k + r] = (int)(flat % (unsigned)N);
        conf[(size_t)p * k + r] = __uint_as_float((unsigned)(e >> 32));
    }
}

size_t pairing_ws_bytes(int P) {
    return (size_t)P * (PAIR_BINS + PAIR_CTL) * sizeof(int) + (size_t)P * PAIR_CAND * sizeof(unsigned long long) + 512;
}

// S (P, M, N) cosine similarities -> in place P = dual softmax; (si, di, conf) (P, k) = its top-k.
// rs (P*M), cs (P*N), ws: pairing_ws_bytes(P) bytes.
int pairing_launch(float *S, int P, int M, int N, float tau, int k, float2 *rs, float2 *cs, void *ws, int32_t *si,
                   int32_t *di, float *conf, cudaStream_t st) {
    if (k > PAIR_MAXK) return fail(DPM_ERR_UNSUPPORTED, "pairing: k=%d exceeds the limit %d", k, PAIR_MAXK);
    if ((long long)M * N >= (1ll << 31)) return fail(DPM_ERR_UNSUPPORTED, "pairing: M*N=%lld too large", (long long)M * N);
    unsigned long long *cand = reinterpret_cast<unsigned long long *>(ws);
    int *hist = reinterpret_cast<int *>(cand + (size_t)P * PAIR_CAND);
    int *ctl = hist + (size_t)P * PAIR_BINS;
    const int nrowblk = (P * M + 7) / 8, colblk = (N + 31) / 32;
    pair_stats_kernel<<<nrowblk + colblk * P, 256, 0, st>>>(S, P, M, N, tau, rs, cs, hist, ctl, nrowblk, colblk);
    DPM_CHECK_LAUNCH("pair_stats", st);
    const long long MN = (long long)M * N;
    pair_softmax_hist_kernel<<<dim3((unsigned)((MN + PAIR_CHUNK2 - 1) / PAIR_CHUNK2), P, 1), PAIR_T2, 0, st>>>(
        S, M, N, tau, rs, cs, k, hist, ctl);
    DPM_CHECK_LAUNCH("pair_softmax_hist", st);
    int kp2 = 2;
    while (kp2 < k) kp2 <<= 1;
    pair_select_kernel<<<dim3((unsigned)((MN + PAIR_CHUNK3 - 1) / PAIR_CHUNK3), P, 1), PAIR_T3, (size_t)kp2 * 8, st>>>(
        S, M, N, k, kp2, ctl, cand, si, di, conf);
    DPM_CHECK_LAUNCH("pair_select", st);
    return DPM_OK;
}

}  // namespace dpm
