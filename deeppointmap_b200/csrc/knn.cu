// knn.cu -- brute-force K nearest neighbours (+ radius mask = the reference's "hybrid"
// query, + ball query), bit-exact under the total order (d2, index) with
// d2 = (dx*dx+dy*dy)+dz*dz in fp32 without FMA.
//
// Replaces pytorch3d knn_points / ball_query as called from Querier.*_t3d,
// network/encoder/utils.py:91-123.
//
// Design (B200): warp-cooperative.  A CTA of 8 warps owns 8*QW queries of one cloud and
// streams the cloud's points through shared memory in 2048-point float4 tiles, staged by
// the TMA engine (cp.async.bulk 1-D copies completing on mbarriers, double buffered) so all
// 8 warps reuse each tile.  Every lane tests one point per step against the warp's QW
// queries; the running K-best list of a query is distributed over the warp (lane l holds
// the l-th best), so "is this point closer than the current K-th" is one register compare,
// and the rare insertion is a ballot + shuffle-up.  For the hybrid query the list is
// additionally capped at the radius, which removes the warm-up insertions.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace dpm {

constexpr int KNN_T = 256;
constexpr int KNN_WARPS = KNN_T / 32;
constexpr int KNN_TILE = 2048;  // points per stage (32 KB)

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// insert every candidate flagged in `cmask` (lane order = index order) into the
// warp-distributed sorted list (ld, li); thr = current K-th distance.
__device__ __forceinline__ void knn_insert(unsigned cmask, float d, int gi, float &ld, int &li, float &thr,
                                           const int K, const unsigned kmask, const int lane) {
    while (cmask) {
        const int src = __ffs(cmask) - 1;
        cmask &= cmask - 1;
        const float dc = __shfl_sync(0xffffffffu, d, src);
        const int ic = __shfl_sync(0xffffffffu, gi, src);
        if (dc < thr) {  // warp-uniform; ties keep the earlier (lower) index
            const int pos = __popc(__ballot_sync(0xffffffffu, ld <= dc) & kmask);
            const float ud = __shfl_up_sync(0xffffffffu, ld, 1);
            const int ui = __shfl_up_sync(0xffffffffu, li, 1);
            if (lane > pos) { ld = ud; li = ui; }
            if (lane == pos) { ld = dc; li = ic; }
            thr = fminf(thr, __shfl_sync(0xffffffffu, ld, K - 1));
        }
    }
}

template <int QW>
__global__ void __launch_bounds__(KNN_T)
knn_kernel(const float4 *__restrict__ q4, const float4 *__restrict__ p4, int S, int N,
           const int *__restrict__ qlen32, const int *__restrict__ plen32, int K, float cap, int mode,
           int64_t *__restrict__ idx64, int32_t *__restrict__ idx32, float *__restrict__ d2out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *stile = reinterpret_cast<float4 *>(smem_raw);
    __shared__ __align__(8) unsigned long long full[2];

    const int b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int len = plen32 ? min(plen32[b], N) : N;
    const int qlen = qlen32 ? min(qlen32[b], S) : S;
    const float4 *pts = p4 + (size_t)b * N;
    const float INF = __int_as_float(0x7f800000);
    const unsigned kmask = K >= 32 ? 0xffffffffu : ((1u << K) - 1u);
    const int ntiles = (len + KNN_TILE - 1) / KNN_TILE;

    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        for (int t = 0; t < 2 && t < ntiles; ++t) {
            const unsigned bytes = (unsigned)min(KNN_TILE, len - t * KNN_TILE) * 16u;
            mbar_expect_tx(&full[t], bytes);
            bulk_g2s(stile + t * KNN_TILE, pts + (size_t)t * KNN_TILE, bytes, &full[t]);
        }
    }

    const int s0 = (blockIdx.x * KNN_WARPS + warp) * QW;
    float qx[QW], qy[QW], qz[QW], ld[QW], thr[QW];
    int li[QW];
#pragma unroll
    for (int q = 0; q < QW; ++q) {
        const int s = s0 + q;
        float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
        thr[q] = -1.f;  // inactive query: nothing is ever closer
        if (s < qlen) {
            c = q4[(size_t)b * S + s];
            thr[q] = cap;
        }
        qx[q] = c.x; qy[q] = c.y; qz[q] = c.z;
        ld[q] = INF;
        li[q] = 0;
    }

    for (int t = 0; t < ntiles; ++t) {
        const int stage = t & 1;
        mbar_wait(&full[stage], (unsigned)((t >> 1) & 1));
        const float4 *tile = stile + stage * KNN_TILE;
        const int base = t * KNN_TILE;
        const int cnt = min(KNN_TILE, len - base);
#pragma unroll 2
        for (int off = 0; off < cnt; off += 32) {
            const int pi = off + lane;
            float4 p = tile[pi];
            if (pi >= cnt) p.x = INF;  // -> d2 = +inf, never a candidate
            float d[QW];
            bool any = false;
#pragma unroll
            for (int q = 0; q < QW; ++q) {
                d[q] = d2_exact(qx[q], qy[q], qz[q], p.x, p.y, p.z);
                any |= d[q] < thr[q];
            }
            if (__any_sync(0xffffffffu, any)) {
#pragma unroll
                for (int q = 0; q < QW; ++q) {
                    const unsigned cm = __ballot_sync(0xffffffffu, d[q] < thr[q]);
                    if (cm) knn_insert(cm, d[q], base + pi, ld[q], li[q], thr[q], K, kmask, lane);
                }
            }
        }
        __syncthreads();  // every warp is done with this stage before it is refilled
        if (tid == 0 && t + 2 < ntiles) {
            const unsigned bytes = (unsigned)min(KNN_TILE, len - (t + 2) * KNN_TILE) * 16u;
            mbar_expect_tx(&full[stage], bytes);
            bulk_g2s(stile + stage * KNN_TILE, pts + (size_t)(t + 2) * KNN_TILE, bytes, &full[stage]);
        }
    }

#pragma unroll
    for (int q = 0; q < QW; ++q) {
        const int s = s0 + q;
        if (s >= S) continue;  // warp-uniform
        const size_t o = ((size_t)b * S + s) * K;
        int oi = 0;
        float od = 0.f;
        if (s < qlen) {
            const int count = __popc(__ballot_sync(0xffffffffu, ld[q] < INF) & kmask);
            const int kvalid = min(len, K);
            if (mode == KNN_MODE_HYBRID) {
                int first = __shfl_sync(0xffffffffu, li[q], 0);
                if (count == 0 && len > 0) {
                    // nothing inside the radius: slot 0 of the uncapped kNN is the nearest point
                    unsigned long long best = ~0ull;
                    for (int i = lane; i < len; i += 32) {
                        const float4 p = pts[i];
                        const float dd = d2_exact(qx[q], qy[q], qz[q], p.x, p.y, p.z);
                        const unsigned long long key = ((unsigned long long)__float_as_uint(dd) << 32) | (unsigned)i;
                        best = key < best ? key : best;
                    }
#pragma unroll
                    for (int sft = 16; sft > 0; sft >>= 1) {
                        const unsigned long long o2 = __shfl_xor_sync(0xffffffffu, best, sft);
                        best = o2 < best ? o2 : best;
                    }
                    first = (int)(unsigned)best;
                }
                oi = lane < count ? li[q] : (lane < kvalid ? first : 0);
            } else {
                oi = lane < count ? li[q] : 0;
                od = lane < count ? ld[q] : 0.f;
            }
        }
        if (lane < K) {
            if (idx64) idx64[o + lane] = (int64_t)oi;
            if (idx32) idx32[o + lane] = oi;
            if (d2out) d2out[o + lane] = od;
        }
    }
}

// K > 32 (pytorch3d knn_points has no limit on K; nothing in the reference asks for more than 32): warp per query,
// ceil(K / 32) passes over the cloud.  Pass p keeps the 32 smallest keys (d2, index) that come AFTER the last key of
// pass p - 1 in the same total order, in the same lane-distributed list as knn_kernel, so the concatenation of the
// passes is the ascending top K.  Points are read straight from L2; O(N K / 32) per query -- a completeness path.
__global__ void __launch_bounds__(KNN_T)
knn_big_kernel(const float4 *__restrict__ q4, const float4 *__restrict__ p4, int S, int N,
               const int *__restrict__ qlen32, const int *__restrict__ plen32, int K, float cap, int mode,
               int64_t *__restrict__ idx64, int32_t *__restrict__ idx32, float *__restrict__ d2out) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int s = blockIdx.x * KNN_WARPS + (threadIdx.x >> 5);
    if (s >= S) return;  // warp-uniform
    const int len = plen32 ? min(plen32[b], N) : N;
    const int qlen = qlen32 ? min(qlen32[b], S) : S;
    const size_t o = ((size_t)b * S + s) * K;
    const float INF = __int_as_float(0x7f800000);
    const float4 *pts = p4 + (size_t)b * N;
    int done = 0;  // slots written so far
    int first = 0;
    if (s < qlen) {
        const float4 c = q4[(size_t)b * S + s];
        float pd = -1.f;  // last key of the previous pass: every d2 is >= 0, so the first pass takes everything
        int pi = -1;
        while (done < K) {
            const int kk = min(32, K - done);
            const unsigned kmask = kk >= 32 ? 0xffffffffu : ((1u << kk) - 1u);
            float ld = INF, thr = cap;
            int li = 0;
            for (int off = 0; off < len; off += 32) {
                const int i = off + lane;
                float d = INF;
                if (i < len) {
                    const float4 p = pts[i];
                    d = d2_exact(c.x, c.y, c.z, p.x, p.y, p.z);
                }
                const bool after = d > pd || (d == pd && i > pi);
                const unsigned cm = __ballot_sync(0xffffffffu, after && d < thr);
                if (cm) knn_insert(cm, d, i, ld, li, thr, kk, kmask, lane);
            }
            const int count = __popc(__ballot_sync(0xffffffffu, ld < INF) & kmask);
            if (done == 0) first = __shfl_sync(0xffffffffu, li, 0);
            if (lane < count) {
                if (idx64) idx64[o + done + lane] = (int64_t)li;
                if (idx32) idx32[o + done + lane] = li;
                if (d2out) d2out[o + done + lane] = ld;
            }
            done += count;
            if (count < kk) break;  // the cloud (or the radius) is exhausted
            pd = __shfl_sync(0xffffffffu, ld, kk - 1);
            pi = __shfl_sync(0xffffffffu, li, kk - 1);
        }
        if (mode == KNN_MODE_HYBRID && done == 0 && len > 0) {
            // nothing inside the radius: slot 0 of the uncapped kNN is the nearest point
            unsigned long long best = ~0ull;
            for (int i = lane; i < len; i += 32) {
                const float4 p = pts[i];
                const float dd = d2_exact(c.x, c.y, c.z, p.x, p.y, p.z);
                const unsigned long long key = ((unsigned long long)__float_as_uint(dd) << 32) | (unsigned)i;
                best = key < best ? key : best;
            }
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) {
                const unsigned long long o2 = __shfl_xor_sync(0xffffffffu, best, sft);
                best = o2 < best ? o2 : best;
            }
            first = (int)(unsigned)best;
        }
    }
    // padding: hybrid repeats slot 0 over the slots a cloud of `len` points could fill, everything else is 0
    const int kvalid = (mode == KNN_MODE_HYBRID && s < qlen) ? min(len, K) : 0;
    for (int k = done + lane; k < K; k += 32) {
        const int oi = k < kvalid ? first : 0;
        if (idx64) idx64[o + k] = (int64_t)oi;
        if (idx32) idx32[o + k] = oi;
        if (d2out) d2out[o + k] = 0.f;
    }
}

template <int QW>
static int knn_launch_t(const float4 *q4, const float4 *p4, int B, int S, int N, const int *qlen32,
                        const int *plen32, int K, float cap, int mode, int64_t *idx64, int32_t *idx32,
                        float *d2out, cudaStream_t st) {
    auto kern = knn_kernel<QW>;
    const size_t smem = 2 * (size_t)KNN_TILE * sizeof(float4);
    static thread_local unsigned long long configured = 0ull;  // one bit per device: function attributes are per context
    const unsigned long long devbit = 1ull << (current_device() & 63);
    if (!(configured & devbit)) {
        DPM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured |= devbit;
    }
    dim3 grid((S + KNN_WARPS * QW - 1) / (KNN_WARPS * QW), B, 1);
    kern<<<grid, KNN_T, smem, st>>>(q4, p4, S, N, qlen32, plen32, K, cap, mode, idx64, idx32, d2out);
    DPM_CHECK_LAUNCH("knn", st);
    return DPM_OK;
}

int knn_launch(const float4 *q4, const float4 *p4, int B, int S, int N, const int *qlen32, const int *plen32,
               int K, float r2, int mode, int64_t *idx64, int32_t *idx32, float *d2out, cudaStream_t st) {
    if (B <= 0 || S <= 0 || N <= 0) return fail(DPM_ERR_SHAPE, "knn: bad shape B=%d S=%d N=%d", B, S, N);
    if (K <= 0) return fail(DPM_ERR_SHAPE, "knn: K=%d", K);
    float cap = __builtin_inff();
    if (mode == KNN_MODE_HYBRID) {
        // keep d2 <= r2  <=>  d2 < nextafter(r2, +inf)
        cap = r2 >= 0.f ? __builtin_nextafterf(r2, __builtin_inff()) : 0.f;
    }
    if (K > 32) {
        prof_note(S, N);
        dim3 grid((S + KNN_WARPS - 1) / KNN_WARPS, B, 1);
        knn_big_kernel<<<grid, KNN_T, 0, st>>>(q4, p4, S, N, qlen32, plen32, K, cap, mode, idx64, idx32, d2out);
        DPM_CHECK_LAUNCH("knn", st);
        return DPM_OK;
    }
    const long long sms = device_sm_count();
    const long long warps = (long long)B * ((S + KNN_WARPS - 1) / KNN_WARPS);  // CTAs at QW=1
    prof_note(S, N);
    if (warps >= 8 * sms)
        return knn_launch_t<4>(q4, p4, B, S, N, qlen32, plen32, K, cap, mode, idx64, idx32, d2out, st);
    if (warps >= 3 * sms)
        return knn_launch_t<2>(q4, p4, B, S, N, qlen32, plen32, K, cap, mode, idx64, idx32, d2out, st);
    return knn_launch_t<1>(q4, p4, B, S, N, qlen32, plen32, K, cap, mode, idx64, idx32, d2out, st);
}

// ball query: first K points in index order with d2 < r2 (pytorch3d contract); warp per query.
__global__ void __launch_bounds__(256)
ball_query_kernel(const float4 *__restrict__ q4, const float4 *__restrict__ p4, int S, int N,
                  const int *__restrict__ qlen32, const int *__restrict__ plen32, int K, float r2,
                  int64_t *__restrict__ idx64, float *__restrict__ d2out) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int s = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (s >= S) return;
    const int len = plen32 ? min(plen32[b], N) : N;
    const int qlen = qlen32 ? min(qlen32[b], S) : S;
    const size_t o = ((size_t)b * S + s) * K;
    int count = 0;
    if (s < qlen) {
        const float4 c = q4[(size_t)b * S + s];
        const float4 *pts = p4 + (size_t)b * N;
        for (int base = 0; base < len && count < K; base += 32) {
            const int i = base + lane;
            float d = __int_as_float(0x7f800000);
            if (i < len) {
                const float4 p = pts[i];
                d = d2_exact(c.x, c.y, c.z, p.x, p.y, p.z);
            }
            const unsigned m = __ballot_sync(0xffffffffu, d < r2);
            const int rank = count + __popc(m & ((1u << lane) - 1u));
            if ((m >> lane & 1u) && rank < K) {
                idx64[o + rank] = i;
                if (d2out) d2out[o + rank] = d;
            }
            count += __popc(m);
        }
        count = min(count, K);
    }
    for (int k = count + lane; k < K; k += 32) {
        idx64[o + k] = -1;
        if (d2out) d2out[o + k] = 0.f;
    }
}

}  // namespace dpm

using namespace dpm;

extern "C" size_t dpm_knn_workspace_bytes(int B, int S, int N, int K) {
    (void)K;
    Arena a(nullptr, 0);
    a.get<float4>((size_t)B * S);
    a.get<float4>((size_t)B * N);
    a.get<int>(B);
    a.get<int>(B);
    return a.off + grid_ws_bytes(B, N) + 256;
}

static int knn_common(const float *p1, int D1, const float *p2, int D2, int B, int S, int N,
                      const int64_t *lengths1, const int64_t *lengths2, int K, float r2, int which,
                      int64_t *idx_out, float *d2_out, void *ws, size_t ws_bytes, cudaStream_t st) {
    if (!p1 || !p2 || !idx_out) return fail(DPM_ERR_ARG, "knn: null pointer");
    if (B <= 0 || S <= 0 || N <= 0 || D1 < 3 || D2 < 3 || K <= 0)
        return fail(DPM_ERR_SHAPE, "knn: bad shape B=%d S=%d N=%d D1=%d D2=%d K=%d", B, S, N, D1, D2, K);
    if (!ws || ws_bytes < dpm_knn_workspace_bytes(B, S, N, K))
        return fail(DPM_ERR_WORKSPACE, "knn: workspace too small (%zu < %zu)", ws_bytes, dpm_knn_workspace_bytes(B, S, N, K));
    Arena a(ws, ws_bytes);
    float4 *q4 = a.get<float4>((size_t)B * S);
    float4 *p4 = a.get<float4>((size_t)B * N);
    int *l1 = a.get<int>(B);
    int *l2 = a.get<int>(B);
    DPM_TRY(pack_xyz4_launch(p1, B, S, D1, q4, st));
    DPM_TRY(pack_xyz4_launch(p2, B, N, D2, p4, st));
    if (lengths1) DPM_TRY(lengths_to_i32_launch(lengths1, B, S, l1, st));
    if (lengths2) DPM_TRY(lengths_to_i32_launch(lengths2, B, N, l2, st));
    if (which == 2) {
        dim3 grid((S + 7) / 8, B, 1);
        ball_query_kernel<<<grid, 256, 0, st>>>(q4, p4, S, N, lengths1 ? l1 : nullptr, lengths2 ? l2 : nullptr, K, r2,
                                                idx_out, d2_out);
        DPM_CHECK_LAUNCH("ball_query", st);
        return DPM_OK;
    }
    if (which == 1 && N >= GRID_MIN_N && N <= GRID_MAX_N && r2 >= 0.f && K <= 32) {
        // radius-capped query: only the 3x3x3 cell neighbourhood of a query can hold its answers
        GridWs g;
        if (!grid_ws_carve(a, B, N, &g)) return fail(DPM_ERR_WORKSPACE, "knn: workspace too small");
        DPM_TRY(grid_build_launch(p4, B, N, lengths2 ? l2 : nullptr, 1.001f * sqrtf(r2), g, st));
        return knn_grid_launch(g, q4, p4, B, S, N, lengths1 ? l1 : nullptr, K, r2, idx_out, nullptr, st);
    }
    static const bool brute = getenv("DPM_KNN_BRUTE") != nullptr;  // developer A/B switch
    if (which == 0 && N >= GRID_MIN_N && N <= GRID_MAX_N && K <= 32 && !brute) {
        // plain kNN on a large cloud: cell grid + shells until the K-th distance is provably final
        GridWs g;
        if (!grid_ws_carve(a, B, N, &g)) return fail(DPM_ERR_WORKSPACE, "knn: workspace too small");
        DPM_TRY(grid_build_launch(p4, B, N, lengths2 ? l2 : nullptr, 0.f, g, st));
        return knn_ring_launch(g, q4, B, S, lengths1 ? l1 : nullptr, K, idx_out, nullptr, d2_out, st);
    }
    return knn_launch(q4, p4, B, S, N, lengths1 ? l1 : nullptr, lengths2 ? l2 : nullptr, K, r2,
                      which == 1 ? KNN_MODE_HYBRID : KNN_MODE_KNN, idx_out, nullptr, d2_out, st);
}

extern "C" int dpm_knn_f32(const float *p1, int D1, const float *p2, int D2, int B, int S, int N,
                           const int64_t *lengths1, const int64_t *lengths2, int K, int64_t *idx_out,
                           float *d2_out, void *ws, size_t ws_bytes, dpm_stream_t stream) {
    return knn_common(p1, D1, p2, D2, B, S, N, lengths1, lengths2, K, 0.f, 0, idx_out, d2_out, ws, ws_bytes,
                      (cudaStream_t)stream);
}

extern "C" int dpm_knn_radius_f32(const float *p1, int D1, const float *p2, int D2, int B, int S, int N,
                                  const int64_t *lengths2, int K, float radius2, int64_t *idx_out, void *ws,
                                  size_t ws_bytes, dpm_stream_t stream) {
    return knn_common(p1, D1, p2, D2, B, S, N, nullptr, lengths2, K, radius2, 1, idx_out, nullptr, ws, ws_bytes,
                      (cudaStream_t)stream);
}

extern "C" int dpm_ball_query_f32(const float *p1, int D1, const float *p2, int D2, int B, int S, int N,
                                  const int64_t *lengths1, const int64_t *lengths2, int K, float radius2,
                                  int64_t *idx_out, float *d2_out, void *ws, size_t ws_bytes, dpm_stream_t stream) {
    return knn_common(p1, D1, p2, D2, B, S, N, lengths1, lengths2, K, radius2, 2, idx_out, d2_out, ws, ws_bytes,
                      (cudaStream_t)stream);
}
