// grid.cu -- a per-cloud uniform cell grid and the two index kernels that use it:
//
//   * fps_grid_kernel : EXACT farthest point sampling with bucket pruning.  The cloud is
//     cell-sorted and cut into buckets of 32*PPL consecutive points; every bucket keeps its
//     bounding box, its largest min-distance and the point that attains it.  A new sample only
//     touches the buckets whose box is closer than that largest min-distance -- for all other
//     buckets min(m_i, d_i) = m_i for every point, so skipping them changes nothing.  The
//     arithmetic on the touched points is the reference's ((dx*dx+dy*dy)+dz*dz, fp32, no FMA,
//     first maximum = lowest ORIGINAL index), so the picks are bit-identical to
//     Sampler.fps (network/encoder/utils.py:209-270) while the work per pick drops from N
//     points to a few hundred.  One CTA per cloud (the points stay in L2), so a batch of
//     clouds fills the chip instead of 8 SMs per cloud.
//   * knn_grid_kernel : the "hybrid" query (kNN capped at the radius,
//     Querier.hybrid_query_t3d, utils.py:112-123) over the 3x3x3 cell neighbourhood of the
//     query instead of the whole cloud; same total order (d2, index) => same rows.
//
// The within-cell order of the sorted array depends on atomic arrival order; every result is
// made independent of it by explicit (value, original index) comparisons.
#include <stdlib.h>

#include "common.cuh"

namespace dpm {

// ---------------------------------------------------------------------------------------
// grid build: bbox -> cell size -> count -> scan -> scatter
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int cell_coord(float p, float o, float inv_h, int g) {
    const float t = (p - o) * inv_h;
    int c = (int)fminf(fmaxf(t, 0.f), (float)(g - 1));
    return c;
}
// unclamped cell coordinate of a query (may lie outside the grid)
__device__ __forceinline__ int cell_coord_free(float p, float o, float inv_h, int g) {
    const float t = floorf((p - o) * inv_h);
    return (int)fminf(fmaxf(t, -2.f), (float)(g + 1));
}

template <int T>
__global__ void __launch_bounds__(T)
grid_bbox_kernel(const float4 *__restrict__ xyz4, int N, const int *__restrict__ len32, float hmin, float ppc,
                 float ppc_flat, GridDesc *__restrict__ desc) {
    __shared__ float red[6][32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int len = len32 ? min(len32[b], N) : N;
    const float4 *pts = xyz4 + (size_t)b * N;
    const float INF = __int_as_float(0x7f800000);
    float lo[3] = {INF, INF, INF}, hi[3] = {-INF, -INF, -INF};
    for (int i = tid; i < len; i += T) {
        const float4 p = pts[i];
        lo[0] = fminf(lo[0], p.x); hi[0] = fmaxf(hi[0], p.x);
        lo[1] = fminf(lo[1], p.y); hi[1] = fmaxf(hi[1], p.y);
        lo[2] = fminf(lo[2], p.z); hi[2] = fmaxf(hi[2], p.z);
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
            for (int w = 1; w < T / 32; ++w) {
                red[a][0] = fminf(red[a][0], red[a][w]);
                red[3 + a][0] = fmaxf(red[3 + a][0], red[3 + a][w]);
            }
        GridDesc d;
        if (len <= 0) {
            d.ox = d.oy = d.oz = 0.f; d.h = 1.f; d.inv_h = 1.f; d.gx = d.gy = d.gz = 1; d.ncell = 1; d.nvalid = 0;
        } else {
            float ex = red[3][0] - red[0][0], ey = red[4][0] - red[1][0], ez = red[5][0] - red[2][0];
            if (!(ex >= 0.f && ex < 1e30f)) ex = 0.f;  // NaN / inf coordinates: degenerate grid, still correct
            if (!(ey >= 0.f && ey < 1e30f)) ey = 0.f;
            if (!(ez >= 0.f && ez < 1e30f)) ez = 0.f;
            const float emax = fmaxf(ex, fmaxf(ey, ez));
            const float fl = fmaxf(0.02f * emax, 1e-20f);
            // ~32 points per occupied cell keeps a bucket within one or two cells
            // target points per cell: 32 keeps a 64-point bucket within one or two cells of a cloud that FILLS its box
            // (uniform cube: 5.5 ms per 65 536-point FPS against 6.1 with 20); a LiDAR scan is a thin slab (z extent a
            // few per cent of x / y) whose points sit on surfaces, and there 20 per cell prunes better (4.77 against 5.13 ms;
            // with the encoder's radius bound on the cell size: 7350 -> 7520 frames/s in the 20-step run)
            const float emin = fminf(ex, fminf(ey, ez));
            const float target = emin < 0.25f * emax ? ppc_flat : ppc;
            float h = cbrtf(fmaxf(ex, fl) * fmaxf(ey, fl) * fmaxf(ez, fl) * target / (float)len);
            h = fmaxf(h, hmin);
            if (!(h > 0.f) || !(h < 1e30f)) h = 1.f;
            int gx, gy, gz;
            for (int it = 0; it < 200; ++it) {
                gx = (int)fminf(ex / h, 1e6f) + 1;
                gy = (int)fminf(ey / h, 1e6f) + 1;
                gz = (int)fminf(ez / h, 1e6f) + 1;
                if ((long long)gx * gy * gz <= GRID_MAXCELL) break;
                h *= 1.26f;
            }
            if ((long long)gx * gy * gz > GRID_MAXCELL) { gx = gy = gz = 1; h = fmaxf(emax, 1.f) * 2.f; }
            d.ox = red[0][0]; d.oy = red[1][0]; d.oz = red[2][0];
            d.h = h; d.inv_h = 1.0f / h; d.gx = gx; d.gy = gy; d.gz = gz; d.ncell = gx * gy * gz; d.nvalid = len;
        }
        d.pad0 = d.pad1 = 0;
        desc[b] = d;
    }
}

__global__ void __launch_bounds__(256)
grid_count_kernel(const float4 *__restrict__ xyz4, int N, const GridDesc *__restrict__ desc,
                  int *__restrict__ counts, int *__restrict__ cellid) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * 256 + threadIdx.x;
    const GridDesc d = desc[b];
    if (i >= d.nvalid) return;
    const float4 p = xyz4[(size_t)b * N + i];
    const int cx = cell_coord(p.x, d.ox, d.inv_h, d.gx), cy = cell_coord(p.y, d.oy, d.inv_h, d.gy),
              cz = cell_coord(p.z, d.oz, d.inv_h, d.gz);
    const int c = (cz * d.gy + cy) * d.gx + cx;
    cellid[(size_t)b * N + i] = c;
    atomicAdd(&counts[(size_t)b * (GRID_MAXCELL + 1) + c], 1);
}

// exclusive scan of the per-cell counts (in place) + a copy as the scatter cursor
template <int T>
__global__ void __launch_bounds__(T)
grid_scan_kernel(const GridDesc *__restrict__ desc, int *__restrict__ cell_start, int *__restrict__ cursor) {
    __shared__ int wsum[32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ncell = desc[b].ncell;
    int *cs = cell_start + (size_t)b * (GRID_MAXCELL + 1);
    int *cu = cursor + (size_t)b * (GRID_MAXCELL + 1);
    const int per = (ncell + T - 1) / T;
    const int c0 = tid * per, c1 = min(ncell, c0 + per);
    int s = 0;
    for (int c = c0; c < c1; ++c) s += cs[c];
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int v = lane < T / 32 ? wsum[lane] : 0, iv = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, iv, o);
            if (lane >= o) iv += t;
        }
        wsum[lane] = iv - v;
    }
    __syncthreads();
    int run = wsum[warp] + incl - s;
    for (int c = c0; c < c1; ++c) {
        const int n = cs[c];
        cs[c] = run;
        cu[c] = run;
        run += n;
    }
    if (tid == T - 1) cs[ncell] = run;  // the last thread's running sum is the total (its range may be empty)
}

__global__ void __launch_bounds__(256)
grid_scatter_kernel(const float4 *__restrict__ xyz4, int N, int npad, const GridDesc *__restrict__ desc,
                    const int *__restrict__ cellid, int *__restrict__ cursor, float4 *__restrict__ sorted) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int nvalid = desc[b].nvalid;
    if (i >= npad) return;
    float4 *dst = sorted + (size_t)b * npad;
    if (i < nvalid) {
        const float4 p = xyz4[(size_t)b * N + i];
        const int c = cellid[(size_t)b * N + i];
        const int pos = atomicAdd(&cursor[(size_t)b * (GRID_MAXCELL + 1) + c], 1);
        dst[pos] = make_float4(p.x, p.y, p.z, __int_as_float(i));
    } else {
        // tail [nvalid, npad): sentinels that never win (min-distance 0, highest index)
        dst[i] = make_float4(0.f, 0.f, 0.f, __int_as_float(0x7fffffff));
    }
}

size_t grid_ws_bytes(int B, int N) {
    Arena a(nullptr, 0);
    const int npad = grid_npad(N);
    a.get<float4>((size_t)B * npad);
    a.get<int>((size_t)B * (GRID_MAXCELL + 1));
    a.get<int>((size_t)B * (GRID_MAXCELL + 1));
    a.get<int>((size_t)B * N);
    a.get<float>((size_t)B * npad);
    a.get<GridDesc>((size_t)B);
    return a.off;
}

bool grid_ws_carve(Arena &a, int B, int N, GridWs *g) {
    g->npad = grid_npad(N);
    g->sorted = a.get<float4>((size_t)B * g->npad);
    g->cell_start = a.get<int>((size_t)B * (GRID_MAXCELL + 1));
    g->cursor = a.get<int>((size_t)B * (GRID_MAXCELL + 1));
    g->cellid = a.get<int>((size_t)B * N);
    g->mind = a.get<float>((size_t)B * g->npad);
    g->desc = a.get<GridDesc>((size_t)B);
    return a.ok();
}

int grid_build_launch(const float4 *xyz4, int B, int N, const int *len32, float hmin, const GridWs &g, cudaStream_t st) {
    if (B <= 0 || N <= 0) return fail(DPM_ERR_SHAPE, "grid: bad shape B=%d N=%d", B, N);
    prof_note(N, 0);
    // team width of the two one-CTA-per-cloud kernels (DPM_GRID_T, A/B switch): a 1024-thread CTA needs half an SM's
    // thread slots at once, which it rarely finds while other streams' kernels are resident
    static const int team = getenv("DPM_GRID_T") ? atoi(getenv("DPM_GRID_T")) : 1024;
    // target points per occupied cell (the cell size never drops below hmin): DPM_GRID_PPC, A/B switch
    static const float ppc = getenv("DPM_GRID_PPC") ? (float)atof(getenv("DPM_GRID_PPC")) : 32.f;
    static const float ppc_flat = getenv("DPM_GRID_PPC_FLAT") ? (float)atof(getenv("DPM_GRID_PPC_FLAT")) : 20.f;
    if (team == 256) grid_bbox_kernel<256><<<B, 256, 0, st>>>(xyz4, N, len32, hmin, ppc, ppc_flat, g.desc);
    else grid_bbox_kernel<1024><<<B, 1024, 0, st>>>(xyz4, N, len32, hmin, ppc, ppc_flat, g.desc);
    DPM_CHECK_LAUNCH("grid_bbox", st);
    DPM_CHECK_CUDA(cudaMemsetAsync(g.cell_start, 0, sizeof(int) * (size_t)B * (GRID_MAXCELL + 1), st));
    dim3 grid((N + 255) / 256, B, 1);
    grid_count_kernel<<<grid, 256, 0, st>>>(xyz4, N, g.desc, g.cell_start, g.cellid);
    DPM_CHECK_LAUNCH("grid_count", st);
    if (team == 256) grid_scan_kernel<256><<<B, 256, 0, st>>>(g.desc, g.cell_start, g.cursor);
    else grid_scan_kernel<1024><<<B, 1024, 0, st>>>(g.desc, g.cell_start, g.cursor);
    DPM_CHECK_LAUNCH("grid_scan", st);
    dim3 grid2((g.npad + 255) / 256, B, 1);
    grid_scatter_kernel<<<grid2, 256, 0, st>>>(xyz4, N, g.npad, g.desc, g.cellid, g.cursor, g.sorted);
    DPM_CHECK_LAUNCH("grid_scatter", st);
    return DPM_OK;
}

// ---------------------------------------------------------------------------------------
// pruned exact FPS: one CTA of T threads per cloud.  The <= 1024 buckets are dealt round-robin
// over the W = T/32 warps (bucket s*T + lane*W + warp is slot s of that lane), so the buckets a
// new sample touches -- spatial neighbours, i.e. consecutive ids -- spread over all warps.  An
// owner keeps its buckets' boxes, largest min-distances and the points attaining them in
// registers.  Per pick: every owner tests its 1024/T boxes; a warp then updates its touched
// buckets, up to D per slot per round with all loads of a round in flight together (one L2 round
// trip per round); warp arg-max -> shared memory -> one barrier -> every warp reduces the W
// records.  Measured on B200 (N = 65536, K = 4096): T = 1024 / 512 / 256 / 128 -> 4.98 / 5.3 /
// 7.1 / 11.3 ms -- a pick is bound by the dependent-instruction latency of ONE warp's chain
// (test, load, update, reduce), so one bucket per thread wins; 2048 buckets of 32 points (two per
// thread) lose 17 %, and a shared-memory work queue that balances the touched buckets over the
// warps costs two more barriers per pick, which is what the balance gains.
// clock64() phase timing of one warp (cycles per pick, 2 260 uninstrumented): box tests 270, a round of
// touched buckets 1 130 (L2 round trip ~800 + update), warp arg-max 400, block arg-max 260; the other
// warps wait at the barrier for the ones that had a round.  Replacing the second REDUX of an arg-max by a
// ballot + shuffle fast path for unique maxima was slower (divergent branch around warp collectives).
// Keeping a 4096-point level entirely in shared memory (no L2 round trip at all) does not change its
// 0.75 us per pick: what bounds a pick is the ~100-instruction dependent chain of tests and three levels of
// arg-max (64 points -> bucket, 32 buckets -> warp, 32 warps -> block), not where the points live.
// Letting untouched warps republish their previous record instead of recomputing it: 4.89 -> 5.36 ms.
// ---------------------------------------------------------------------------------------
// TEAMS > 1: a CTA of TEAMS * T threads holds TEAMS independent teams, each with its own cloud, its own records and its
// own named barrier -- two 512-thread teams share one SM (the block scheduler would spread two 512-thread CTAs over two
// SMs and block half of each).
template <int T, int TEAMS>
__device__ __forceinline__ void fps_team_sync(int team) {
    if (TEAMS == 1) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(T) : "memory");
}

template <int PPL, int T, int BPT, int D, int TEAMS = 1>
__global__ void __launch_bounds__(T * TEAMS, (T * TEAMS >= 1024 || PPL >= 16) ? 1 : 2)  // narrower CTAs keep <= 64 registers (the rest of the SM stays free), huge clouds get 128
fps_grid_kernel(const float4 *__restrict__ sorted, float *__restrict__ mind, int npad,
                const GridDesc *__restrict__ desc, const float4 *__restrict__ xyz4, int N, int K,
                int64_t *__restrict__ idx64, int32_t *__restrict__ idx32, float4 *__restrict__ new_xyz4,
                uint8_t *__restrict__ new_pad, int *__restrict__ new_len32, int nclouds) {
    constexpr int BS = 32 * PPL, W = T / 32;
    __shared__ unsigned long long skey_all[TEAMS][2][W];
    __shared__ float4 sxyz_all[TEAMS][2][W];
    const int team = TEAMS == 1 ? 0 : (int)threadIdx.x / T;
    const int b = blockIdx.x * TEAMS + team;
    if (b >= nclouds) return;  // a whole team: its named barrier is never used
    unsigned long long (*skey)[W] = skey_all[team];
    float4 (*sxyz)[W] = sxyz_all[team];
    const int tid = TEAMS == 1 ? (int)threadIdx.x : (int)threadIdx.x % T, lane = tid & 31, warp = tid >> 5;
    const int len = desc[b].nvalid;
    const int kn = min(len, K);
    const int nb = (len + BS - 1) / BS;  // <= T * BPT
    const float4 *P = sorted + (size_t)b * npad;
    float *M = mind + (size_t)b * npad;
    const size_t ob = (size_t)b * K;
    const float INF = __int_as_float(0x7f800000);

    // ---- prologue: bucket boxes, min-distances = +inf (sentinels 0) -------------------------
    float lox[BPT], loy[BPT], loz[BPT], hix[BPT], hiy[BPT], hiz[BPT], ax[BPT], ay[BPT], az[BPT];
    unsigned maxbits[BPT], argidx[BPT];
    bool owns[BPT];
#pragma unroll
    for (int s = 0; s < BPT; ++s) {
        lox[s] = loy[s] = loz[s] = hix[s] = hiy[s] = hiz[s] = ax[s] = ay[s] = az[s] = 0.f;
        maxbits[s] = 0u;
        argidx[s] = 0xffffffffu;
        owns[s] = s * T + lane * W + warp < nb;
        for (int L = 0; L < 32; ++L) {
            const int bk = s * T + L * W + warp;
            if (bk >= nb) break;  // warp-uniform
            float l0 = INF, l1 = INF, l2 = INF, h0 = -INF, h1 = -INF, h2 = -INF;
            unsigned mi = 0xffffffffu;
#pragma unroll
            for (int j = 0; j < PPL; ++j) {
                const int i = bk * BS + j * 32 + lane;
                const float4 p = P[i];
                const unsigned id = __float_as_uint(p.w);
                const bool valid = id != 0x7fffffffu;
                M[i] = valid ? INF : 0.f;
                if (valid) {
                    l0 = fminf(l0, p.x); h0 = fmaxf(h0, p.x);
                    l1 = fminf(l1, p.y); h1 = fmaxf(h1, p.y);
                    l2 = fminf(l2, p.z); h2 = fmaxf(h2, p.z);
                    mi = min(mi, id);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                l0 = fminf(l0, __shfl_xor_sync(0xffffffffu, l0, o)); h0 = fmaxf(h0, __shfl_xor_sync(0xffffffffu, h0, o));
                l1 = fminf(l1, __shfl_xor_sync(0xffffffffu, l1, o)); h1 = fmaxf(h1, __shfl_xor_sync(0xffffffffu, h1, o));
                l2 = fminf(l2, __shfl_xor_sync(0xffffffffu, l2, o)); h2 = fmaxf(h2, __shfl_xor_sync(0xffffffffu, h2, o));
            }
            mi = __reduce_min_sync(0xffffffffu, mi);
            if (lane == L) {
                lox[s] = l0; loy[s] = l1; loz[s] = l2; hix[s] = h0; hiy[s] = h1; hiz[s] = h2;
                maxbits[s] = 0x7f800000u;  // +inf: every bucket is touched by the first sample
                argidx[s] = mi;
            }
        }
    }

    float sx = 0.f, sy = 0.f, sz = 0.f;
    if (len > 0) {
        const float4 p0 = xyz4[(size_t)b * N];
        sx = p0.x; sy = p0.y; sz = p0.z;
    }
    if (tid == 0 && kn > 0) {
        if (idx64) idx64[ob] = 0;
        if (idx32) idx32[ob] = 0;
        if (new_xyz4) new_xyz4[ob] = make_float4(sx, sy, sz, 0.f);
        if (new_pad) new_pad[ob] = 0;
    }
    fps_team_sync<T, TEAMS>(team);  // M initialised before any bucket is processed (a bucket is only ever touched by its own warp,
                      // but keep the prologue and the loop cleanly separated)

    int par = 0;
    for (int k = 1; k < kn; ++k) {
        // ---- which of my warp's buckets can change?  (d2 >= lb for every point of the box) -------
        unsigned mask[BPT];
        unsigned anym = 0u;
#pragma unroll
        for (int s = 0; s < BPT; ++s) {
            bool act = false;
            if (owns[s]) {
                const float dx = fmaxf(fmaxf(lox[s] - sx, sx - hix[s]), 0.f);
                const float dy = fmaxf(fmaxf(loy[s] - sy, sy - hiy[s]), 0.f);
                const float dz = fmaxf(fmaxf(loz[s] - sz, sz - hiz[s]), 0.f);
                const float lb = (dx * dx + dy * dy + dz * dz) * 0.99999f;  // conservative lower bound of every d2
                act = lb < __uint_as_float(maxbits[s]);
            }
            mask[s] = __ballot_sync(0xffffffffu, act);
            anym |= mask[s];
        }
        while (anym) {
            // one round: the first D touched buckets of every slot; all their loads are issued before the first
            // one is consumed
            int Ls[BPT * D];
            float4 p[BPT * D][PPL];
            float m[BPT * D][PPL];
#pragma unroll
            for (int s = 0; s < BPT; ++s) {
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    const int q = s * D + d;
                    Ls[q] = -1;
                    if (mask[s]) {
                        Ls[q] = __ffs(mask[s]) - 1;
                        mask[s] &= mask[s] - 1;
                        const int base = (s * T + Ls[q] * W + warp) * BS + lane;
#pragma unroll
                        for (int j = 0; j < PPL; ++j) {
                            p[q][j] = P[base + j * 32];
                            m[q][j] = M[base + j * 32];
                        }
                    }
                }
            }
            anym = 0u;
#pragma unroll
            for (int s = 0; s < BPT; ++s) {
                anym |= mask[s];
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    const int q = s * D + d;
                    if (Ls[q] < 0) continue;  // warp-uniform
                    const int base = (s * T + Ls[q] * W + warp) * BS + lane;
                    float bestv = -1.f, bx = 0.f, by = 0.f, bz = 0.f;
                    unsigned besti = 0xffffffffu;
#pragma unroll
                    for (int j = 0; j < PPL; ++j) {
                        const float dd = d2_exact(sx, sy, sz, p[q][j].x, p[q][j].y, p[q][j].z);
                        const float nm = fminf(m[q][j], dd);
                        if (nm < m[q][j]) M[base + j * 32] = nm;
                        const unsigned id = __float_as_uint(p[q][j].w);
                        if (nm > bestv || (nm == bestv && id < besti)) {
                            bestv = nm; besti = id; bx = p[q][j].x; by = p[q][j].y; bz = p[q][j].z;
                        }
                    }
                    const unsigned bits = __float_as_uint(bestv);  // bestv >= 0: the bit pattern is order preserving
                    const unsigned wmax = __reduce_max_sync(0xffffffffu, bits);
                    const unsigned wmin = __reduce_min_sync(0xffffffffu, bits == wmax ? besti : 0xffffffffu);
                    const int src = __ffs(__ballot_sync(0xffffffffu, bits == wmax && besti == wmin)) - 1;
                    const float wx = __shfl_sync(0xffffffffu, bx, src);
                    const float wy = __shfl_sync(0xffffffffu, by, src);
                    const float wz = __shfl_sync(0xffffffffu, bz, src);
                    if (lane == Ls[q]) { maxbits[s] = wmax; argidx[s] = wmin; ax[s] = wx; ay[s] = wy; az[s] = wz; }
                }
            }
        }
        // ---- arg-max over all buckets: (value desc, original index asc) -------------------------
        unsigned vb = 0u, vi = 0xffffffffu;
        float vx = 0.f, vy = 0.f, vz = 0.f;
#pragma unroll
        for (int s = 0; s < BPT; ++s) {
            if (owns[s] && (maxbits[s] > vb || (maxbits[s] == vb && argidx[s] < vi))) {
                vb = maxbits[s]; vi = argidx[s]; vx = ax[s]; vy = ay[s]; vz = az[s];
            }
        }
        const unsigned wmax = __reduce_max_sync(0xffffffffu, vb);
        const unsigned wmin = __reduce_min_sync(0xffffffffu, vb == wmax ? vi : 0xffffffffu);
        const int src = __ffs(__ballot_sync(0xffffffffu, vb == wmax && vi == wmin)) - 1;
        if (lane == max(src, 0)) {
            skey[par][warp] = wmin == 0xffffffffu ? 0ull : (((unsigned long long)wmax << 32) | (unsigned long long)(0xffffffffu - wmin));
            sxyz[par][warp] = make_float4(vx, vy, vz, 0.f);
        }
        fps_team_sync<T, TEAMS>(team);
        const unsigned long long kk = lane < W ? skey[par][lane] : 0ull;
        const unsigned hi = (unsigned)(kk >> 32), lo = (unsigned)kk;
        const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
        const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
        const int slot = __ffs(__ballot_sync(0xffffffffu, lane < W && hi == mh && lo == ml)) - 1;
        const float4 w = sxyz[par][max(slot, 0)];
        sx = w.x; sy = w.y; sz = w.z;
        if (tid == 0) {
            const unsigned sel = 0xffffffffu - ml;
            if (idx64) idx64[ob + k] = (int64_t)sel;
            if (idx32) idx32[ob + k] = (int32_t)sel;
            if (new_xyz4) new_xyz4[ob + k] = make_float4(sx, sy, sz, 0.f);
            if (new_pad) new_pad[ob + k] = 0;
        }
        par ^= 1;
    }
    for (int k = kn + tid; k < K; k += T) {  // K > len: idx -1, zero rows, padded
        if (idx64) idx64[ob + k] = -1;
        if (idx32) idx32[ob + k] = -1;
        if (new_xyz4) new_xyz4[ob + k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (new_pad) new_pad[ob + k] = 1;
    }
    if (tid == 0 && new_len32) new_len32[b] = kn;
}

static inline bool p_ok_for_pair(int ppl) { return ppl <= 2; }

int fps_grid_launch(const GridWs &g, const float4 *xyz4, int B, int N, int K, int64_t *idx64, int32_t *idx32,
                    float4 *new_xyz4, uint8_t *new_pad, int *new_len32, cudaStream_t st) {
    if (B <= 0 || N <= 0 || K <= 0) return fail(DPM_ERR_SHAPE, "fps: bad shape B=%d N=%d K=%d", B, N, K);
    if (N > GRID_MAX_N) return fail(DPM_ERR_UNSUPPORTED, "fps: N=%d exceeds the limit %d", N, GRID_MAX_N);
    if (fps_cluster_mode(B) && N <= GRID_CLUSTER_MAX_N)  // few clouds: 8 SMs per cloud instead of one (fps_cluster.cu)
        return fps_grid_cluster_launch(g, xyz4, B, N, K, idx64, idx32, new_xyz4, new_pad, new_len32, st);
    // Measured and rejected (round 2): the cluster kernel's algorithm (broadcast arg-max, two picks per round, mbarrier
    // record exchange) in a ONE-CTA geometry of 32 warps, fps_grid_onesm_launch: bit-exact, but 6.4 ms against this
    // kernel's 4.8 ms per 65 536-point cloud -- with 8 warps per scheduler the exchange through st.async + mbarrier
    // costs more than one __syncthreads, and the bucket maxima sit on the critical path when the tiles are not in
    // shared memory.  DPM_FPS_ONESM=1 selects it for A/B runs.
    static const bool onesm_new = getenv("DPM_FPS_ONESM") && atoi(getenv("DPM_FPS_ONESM")) == 1;
    if (onesm_new && N <= GRID_CLUSTER_MAX_N) return fps_grid_onesm_launch(g, xyz4, B, N, K, idx64, idx32, new_xyz4, new_pad, new_len32, st);
    const int ppl = grid_ppl(N);
    prof_note(N, K);
#define DPM_FG_ARGS g.sorted, g.mind, g.npad, g.desc, xyz4, N, K, idx64, idx32, new_xyz4, new_pad, new_len32, B
    // team width: 1024 threads own one bucket each; 512 / 256 threads own two / four and leave the rest of the SM's
    // threads and registers to the other streams' kernels (DPM_FPS_T, A/B switch)
    static const int team = getenv("DPM_FPS_T") ? atoi(getenv("DPM_FPS_T")) : 1024;
    // two 512-thread teams (two clouds) per CTA: dpm_set_fps_mode(3); clouds of <= 65 536 points (no register spills)
    const bool pair = fps_packed_mode() && p_ok_for_pair(ppl);
#define DPM_FG_CASE(p)                                                \
    case p:                                                           \
        if (pair) fps_grid_kernel<p, 512, 2, 1, 2><<<(B + 1) / 2, 1024, 0, st>>>(DPM_FG_ARGS); \
        else if (team == 512) fps_grid_kernel<p, 512, 2, 1><<<B, 512, 0, st>>>(DPM_FG_ARGS); \
        else if (team == 256) fps_grid_kernel<p, 256, 4, 1><<<B, 256, 0, st>>>(DPM_FG_ARGS); \
        else fps_grid_kernel<p, 1024, 1, (p <= 2 ? 2 : 1)><<<B, 1024, 0, st>>>(DPM_FG_ARGS); \
        break;
    switch (ppl) {
        DPM_FG_CASE(1) DPM_FG_CASE(2) DPM_FG_CASE(4) DPM_FG_CASE(8)
        // clouds of 262 145 .. 1 048 576 points (beyond any LiDAR frame; pytorch3d has no limit): the same kernel with
        // 16 / 32 points per lane, two buckets per thread so that a thread may hold 128 registers
        case 16: fps_grid_kernel<16, 512, 2, 1><<<B, 512, 0, st>>>(DPM_FG_ARGS); break;
        case 32: fps_grid_kernel<32, 512, 2, 1><<<B, 512, 0, st>>>(DPM_FG_ARGS); break;
        default:
            return fail(DPM_ERR_UNSUPPORTED, "fps: no kernel for %d points per lane", ppl);
    }
#undef DPM_FG_CASE
    DPM_CHECK_LAUNCH("fps", st);
    return DPM_OK;
}

// ---------------------------------------------------------------------------------------
// hybrid query over the cell neighbourhood.  Warp per query; the running K-best list is
// distributed over the lanes (lane l = l-th best), ordered by (d2, original index).
// ---------------------------------------------------------------------------------------
constexpr int KG_T = 256;

// PAD = false: the reference's hybrid query (slots past the in-radius count repeat slot 0; a query with
// nothing in the radius falls back to its overall nearest point).  PAD = true: slots past the count are
// -1 and there is no fallback -- "the K nearest within the radius" (information matrix, pre-filters).
template <bool PAD>
__global__ void __launch_bounds__(KG_T)
knn_grid_kernel(const float4 *__restrict__ q4, const float4 *__restrict__ sorted, int npad,
                const int *__restrict__ cell_start, const GridDesc *__restrict__ desc,
                const float4 *__restrict__ p4, int S, int N, const int *__restrict__ qlen32, int K, float cap,
                int64_t *__restrict__ idx64, int32_t *__restrict__ idx32) {
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int s = blockIdx.x * (KG_T / 32) + (threadIdx.x >> 5);
    if (s >= S) return;  // warp-uniform
    const GridDesc d = desc[b];
    const int len = d.nvalid;
    const int qlen = qlen32 ? min(qlen32[b], S) : S;
    const size_t o = ((size_t)b * S + s) * K;
    const unsigned kmask = K >= 32 ? 0xffffffffu : ((1u << K) - 1u);
    const float INF = __int_as_float(0x7f800000);
    if (s >= qlen) {
        if (lane < K) {
            if (idx64) idx64[o + lane] = 0;
            if (idx32) idx32[o + lane] = 0;
        }
        return;
    }
    const float4 c = q4[(size_t)b * S + s];
    const float4 *P = sorted + (size_t)b * npad;
    const int *cs = cell_start + (size_t)b * (GRID_MAXCELL + 1);
    const int cx = cell_coord_free(c.x, d.ox, d.inv_h, d.gx), cy = cell_coord_free(c.y, d.oy, d.inv_h, d.gy),
              cz = cell_coord_free(c.z, d.oz, d.inv_h, d.gz);
    // lane r < 9 owns the x-row (cy + r%3 - 1, cz + r/3 - 1): one contiguous range of the sorted array
    // Cells that cannot hold an answer are not scanned: a lower bound of the distance from the query to a neighbouring
    // cell's slab (its near face, pulled in by a margin that covers the rounding of the cell assignment, ~1e-6 of the
    // extent) already exceeds the cap.  With cells ~1.2-1.5 radii wide a third of the 3 x 3 x 3 block's candidates goes
    // (stage 0: 665 -> 448 per query); what is dropped could never have passed `d2 < cap`, so the result is unchanged.
    int rs = 0, re = 0;
    if (lane < 9) {
        const int dyi = (lane % 3) - 1, dzi = (lane / 3) - 1;
        const int yy = cy + dyi, zz = cz + dzi;
        int x0 = max(cx - 1, 0), x1 = min(cx + 1, d.gx - 1);
        if (yy >= 0 && yy < d.gy && zz >= 0 && zz < d.gz && x0 <= x1) {
            const float margin = 1e-5f * d.h * (float)max(d.gx, max(d.gy, d.gz));
            const float capx = cap * 1.00001f;  // +inf stays +inf: nothing is culled without a radius
            float fy = dyi == 0 ? 0.f : (dyi < 0 ? c.y - (d.oy + (float)cy * d.h) : (d.oy + (float)(cy + 1) * d.h) - c.y);
            float fz = dzi == 0 ? 0.f : (dzi < 0 ? c.z - (d.oz + (float)cz * d.h) : (d.oz + (float)(cz + 1) * d.h) - c.z);
            fy = fmaxf(fy - margin, 0.f);
            fz = fmaxf(fz - margin, 0.f);
            const float base = fy * fy + fz * fz;
            if (!(base > capx)) {
                if (x0 < cx) {  // the cell below the query's own along x
                    const float fx = fmaxf(c.x - (d.ox + (float)cx * d.h) - margin, 0.f);
                    if (fx * fx + base > capx) x0 = cx;
                }
                if (x1 > cx) {  // the cell above
                    const float fx = fmaxf((d.ox + (float)(cx + 1) * d.h) - c.x - margin, 0.f);
                    if (fx * fx + base > capx) x1 = cx;
                }
                if (x0 <= x1) {  // (x0 = cx > x1 happens for a query outside the grid whose only cell was culled)
                    const int row = (zz * d.gy + yy) * d.gx;
                    rs = cs[row + x0];
                    re = cs[row + x1 + 1];
                }
            }
        }
    }
    float ld = INF, thrd = cap;  // candidates must satisfy (d, i) < (thrd, thri)
    int li = 0x7fffffff, thri = -1;
    // Candidates that beat the current bound are appended to a per-warp staging buffer; every 32 of them
    // are sorted across the lanes (bitonic) and merged into the running list in one go.  The bound is
    // only refreshed at a flush, so a few candidates pass that a per-candidate insertion would have
    // rejected -- the merge drops them, the final list is the same (d2, index)-ordered top K.
    __shared__ float sbd[KG_T / 32][64];
    __shared__ int sbi[KG_T / 32][64];
    float *bd = sbd[threadIdx.x >> 5];
    int *bi = sbi[threadIdx.x >> 5];
    int fill = 0;
    const unsigned lt = (1u << lane) - 1u;
    auto flush = [&](float cd, int ci) {
        // sort the 32 staged candidates ascending by (d, i)
#pragma unroll
        for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {
                const float od = __shfl_xor_sync(0xffffffffu, cd, j);
                const int oi = __shfl_xor_sync(0xffffffffu, ci, j);
                const bool keep_min = (((lane & k) == 0) == ((lane & j) == 0));
                const bool less = od < cd || (od == cd && oi < ci);
                if (keep_min == less) { cd = od; ci = oi; }
            }
        }
        // the 32 smallest of (list U candidates): element-wise min of the list and the reversed candidates
        // is bitonic; one bitonic merge sorts it
        const float rd = __shfl_sync(0xffffffffu, cd, 31 - lane);
        const int ri = __shfl_sync(0xffffffffu, ci, 31 - lane);
        if (rd < ld || (rd == ld && ri < li)) { ld = rd; li = ri; }
#pragma unroll
        for (int j = 16; j > 0; j >>= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, ld, j);
            const int oi = __shfl_xor_sync(0xffffffffu, li, j);
            const bool keep_min = (lane & j) == 0;
            const bool less = od < ld || (od == ld && oi < li);
            if (keep_min == less) { ld = od; li = oi; }
        }
        if (lane >= K) { ld = INF; li = 0x7fffffff; }
        const float kd = __shfl_sync(0xffffffffu, ld, K - 1);
        const int ki = __shfl_sync(0xffffffffu, li, K - 1);
        if (kd < INF) { thrd = kd; thri = ki; }
    };
    // (visiting the query's own x-row first so that the bound tightens sooner was measured: no change -- at the encoder's
    // radii most queries see fewer than K points inside the radius, so every in-radius candidate passes in any order)
    for (int r = 0; r < 9; ++r) {
        const int st = __shfl_sync(0xffffffffu, rs, r), en = __shfl_sync(0xffffffffu, re, r);
        for (int i0 = st; i0 < en; i0 += 32) {
            const int i = i0 + lane;
            float dd = INF;
            int gi = 0x7fffffff;
            if (i < en) {
                const float4 p = P[i];
                dd = d2_exact(c.x, c.y, c.z, p.x, p.y, p.z);
                gi = __float_as_int(p.w);
            }
            const bool pass = dd < thrd || (dd == thrd && gi < thri);
            const unsigned cm = __ballot_sync(0xffffffffu, pass);
            if (cm) {
                if (pass) {
                    const int slot = fill + __popc(cm & lt);
                    bd[slot] = dd;
                    bi[slot] = gi;
                }
                fill += __popc(cm);
                if (fill >= 32) {
                    __syncwarp();
                    const float cd = bd[lane];
                    const int ci = bi[lane];
                    const float td = bd[32 + lane];
                    const int ti = bi[32 + lane];
                    __syncwarp();
                    fill -= 32;
                    if (lane < fill) { bd[lane] = td; bi[lane] = ti; }
                    flush(cd, ci);
                }
            }
        }
    }
    if (fill > 0) {
        __syncwarp();
        flush(lane < fill ? bd[lane] : INF, lane < fill ? bi[lane] : 0x7fffffff);
    }
    const int count = __popc(__ballot_sync(0xffffffffu, ld < INF) & kmask);
    const int kvalid = min(len, K);
    int first = __shfl_sync(0xffffffffu, li, 0);
    if (!PAD && count == 0 && len > 0) {
        // nothing inside the radius: slot 0 of the uncapped kNN is the nearest point overall
        const float4 *pts = p4 + (size_t)b * N;
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
    const int oi = lane < count ? li : (PAD ? -1 : (lane < kvalid ? first : 0));
    if (lane < K) {
        if (idx64) idx64[o + lane] = (int64_t)oi;
        if (idx32) idx32[o + lane] = oi;
    }
}

// ---------------------------------------------------------------------------------------
// plain (uncapped) kNN over the grid: the 3x3x3 block first, then cubic shells of cells until the K-th
// distance found is provably final -- every point not yet scanned lies beyond one of the block's faces, so
// it is at least as far as the nearest face that still has cells behind it.  Same (d2, index) order and
// the same buffered 32-at-a-time merge as knn_grid_kernel; pytorch3d knn_points contract (ascending,
// idx / d2 zero-padded when the cloud has fewer than K points).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(KG_T)
knn_ring_kernel(const float4 *__restrict__ q4, const float4 *__restrict__ sorted, int npad,
                const int *__restrict__ cell_start, const GridDesc *__restrict__ desc, int S,
                const int *__restrict__ qlen32, int K, int64_t *__restrict__ idx64, int32_t *__restrict__ idx32,
                float *__restrict__ d2out) {
    __shared__ float sbd[KG_T / 32][64];
    __shared__ int sbi[KG_T / 32][64];
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int s = blockIdx.x * (KG_T / 32) + (threadIdx.x >> 5);
    if (s >= S) return;  // warp-uniform
    const GridDesc d = desc[b];
    const int qlen = qlen32 ? min(qlen32[b], S) : S;
    const size_t o = ((size_t)b * S + s) * K;
    const float INF = __int_as_float(0x7f800000);
    float ld = INF;
    int li = 0x7fffffff;
    if (s < qlen && d.nvalid > 0) {
        const float4 c = q4[(size_t)b * S + s];
        const float4 *P = sorted + (size_t)b * npad;
        const int *cs = cell_start + (size_t)b * (GRID_MAXCELL + 1);
        const int cx = cell_coord_free(c.x, d.ox, d.inv_h, d.gx), cy = cell_coord_free(c.y, d.oy, d.inv_h, d.gy),
                  cz = cell_coord_free(c.z, d.oz, d.inv_h, d.gz);
        float thrd = INF;
        int thri = 0x7fffffff;
        float *bd = sbd[threadIdx.x >> 5];
        int *bi = sbi[threadIdx.x >> 5];
        int fill = 0;
        const unsigned lt = (1u << lane) - 1u;
        auto flush = [&](float cd, int ci) {
#pragma unroll
            for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
                for (int j = k >> 1; j > 0; j >>= 1) {
                    const float od = __shfl_xor_sync(0xffffffffu, cd, j);
                    const int oi = __shfl_xor_sync(0xffffffffu, ci, j);
                    const bool keep_min = (((lane & k) == 0) == ((lane & j) == 0));
                    const bool less = od < cd || (od == cd && oi < ci);
                    if (keep_min == less) { cd = od; ci = oi; }
                }
            }
            const float rd = __shfl_sync(0xffffffffu, cd, 31 - lane);
            const int ri = __shfl_sync(0xffffffffu, ci, 31 - lane);
            if (rd < ld || (rd == ld && ri < li)) { ld = rd; li = ri; }
#pragma unroll
            for (int j = 16; j > 0; j >>= 1) {
                const float od = __shfl_xor_sync(0xffffffffu, ld, j);
                const int oi = __shfl_xor_sync(0xffffffffu, li, j);
                const bool keep_min = (lane & j) == 0;
                const bool less = od < ld || (od == ld && oi < li);
                if (keep_min == less) { ld = od; li = oi; }
            }
            if (lane >= K) { ld = INF; li = 0x7fffffff; }
            const float kd = __shfl_sync(0xffffffffu, ld, K - 1);
            const int ki = __shfl_sync(0xffffffffu, li, K - 1);
            if (kd < INF) { thrd = kd; thri = ki; }
        };
        auto scan = [&](int st, int en) {  // warp-uniform range of the sorted array
            for (int i0 = st; i0 < en; i0 += 32) {
                const int i = i0 + lane;
                float dd = INF;
                int gi = 0x7fffffff;
                if (i < en) {
                    const float4 p = P[i];
                    dd = d2_exact(c.x, c.y, c.z, p.x, p.y, p.z);
                    gi = __float_as_int(p.w);
                }
                const bool pass = dd < thrd || (dd == thrd && gi < thri);
                const unsigned cm = __ballot_sync(0xffffffffu, pass);
                if (cm) {
                    if (pass) {
                        const int slot = fill + __popc(cm & lt);
                        bd[slot] = dd;
                        bi[slot] = gi;
                    }
                    fill += __popc(cm);
                    if (fill >= 32) {
                        __syncwarp();
                        const float cd = bd[lane];
                        const int ci = bi[lane];
                        const float td = bd[32 + lane];
                        const int ti = bi[32 + lane];
                        __syncwarp();
                        fill -= 32;
                        if (lane < fill) { bd[lane] = td; bi[lane] = ti; }
                        flush(cd, ci);
                    }
                }
            }
        };
        const int gmax = max(d.gx, max(d.gy, d.gz));
        for (int R = 1;; ++R) {
            // ---- cells of shell R (R = 1: the whole 3x3x3 block): rows (dy, dz), one or two x-segments each ----
            const int side = 2 * R + 1;
            for (int t0 = 0; t0 < side * side; t0 += 32) {
                const int t = t0 + lane;
                int s0 = 0, e0 = 0, s1 = 0, e1 = 0;
                if (t < side * side) {
                    const int dy = t % side - R, dz = t / side - R;
                    const int yy = cy + dy, zz = cz + dz;
                    if (yy >= 0 && yy < d.gy && zz >= 0 && zz < d.gz) {
                        const int row = (zz * d.gy + yy) * d.gx;
                        if (R == 1 || abs(dy) == R || abs(dz) == R) {
                            const int x0 = max(cx - R, 0), x1 = min(cx + R, d.gx - 1);
                            if (x0 <= x1) { s0 = cs[row + x0]; e0 = cs[row + x1 + 1]; }
                        } else {
                            const int xa = cx - R, xb = cx + R;
                            if (xa >= 0 && xa < d.gx) { s0 = cs[row + xa]; e0 = cs[row + xa + 1]; }
                            if (xb >= 0 && xb < d.gx) { s1 = cs[row + xb]; e1 = cs[row + xb + 1]; }
                        }
                    }
                }
                unsigned nz = __ballot_sync(0xffffffffu, e0 > s0 || e1 > s1);
                while (nz) {
                    const int src = __ffs(nz) - 1;
                    nz &= nz - 1;
                    const int a0 = __shfl_sync(0xffffffffu, s0, src), b0 = __shfl_sync(0xffffffffu, e0, src);
                    const int a1 = __shfl_sync(0xffffffffu, s1, src), b1 = __shfl_sync(0xffffffffu, e1, src);
                    scan(a0, b0);
                    scan(a1, b1);
                }
            }
            if (fill > 0) {
                __syncwarp();
                flush(lane < fill ? bd[lane] : INF, lane < fill ? bi[lane] : 0x7fffffff);
                fill = 0;
            }
            // ---- done?  every unscanned point is beyond a face of the block that still has cells behind it ----
            const bool cover = cx - R <= 0 && cx + R >= d.gx - 1 && cy - R <= 0 && cy + R >= d.gy - 1 && cz - R <= 0 &&
                               cz + R >= d.gz - 1;
            if (cover || R > gmax + 3) break;
            float gap = INF;
            if (cx - R > 0) gap = fminf(gap, c.x - (d.ox + (float)(cx - R) * d.h));
            if (cx + R < d.gx - 1) gap = fminf(gap, (d.ox + (float)(cx + R + 1) * d.h) - c.x);
            if (cy - R > 0) gap = fminf(gap, c.y - (d.oy + (float)(cy - R) * d.h));
            if (cy + R < d.gy - 1) gap = fminf(gap, (d.oy + (float)(cy + R + 1) * d.h) - c.y);
            if (cz - R > 0) gap = fminf(gap, c.z - (d.oz + (float)(cz - R) * d.h));
            if (cz + R < d.gz - 1) gap = fminf(gap, (d.oz + (float)(cz + R + 1) * d.h) - c.z);
            // the cell of a point is floor((p - o) * inv_h) in fp32: leave a margin for that rounding
            gap -= 1e-4f * (d.h + fabsf(c.x - d.ox) + fabsf(c.y - d.oy) + fabsf(c.z - d.oz));
            const float kd = __shfl_sync(0xffffffffu, ld, min(K, d.nvalid) - 1);
            if (gap > 0.f && kd <= gap * gap) break;
        }
    }
    const unsigned kmask = K >= 32 ? 0xffffffffu : ((1u << K) - 1u);
    const int count = __popc(__ballot_sync(0xffffffffu, ld < INF) & kmask);
    if (lane < K) {
        const int oi = lane < count ? li : 0;
        if (idx64) idx64[o + lane] = (int64_t)oi;
        if (idx32) idx32[o + lane] = oi;
        if (d2out) d2out[o + lane] = lane < count ? ld : 0.f;
    }
}

int knn_ring_launch(const GridWs &g, const float4 *q4, int B, int S, const int *qlen32, int K, int64_t *idx64,
                    int32_t *idx32, float *d2out, cudaStream_t st) {
    if (B <= 0 || S <= 0) return fail(DPM_ERR_SHAPE, "knn: bad shape B=%d S=%d", B, S);
    if (K <= 0 || K > 32) return fail(DPM_ERR_UNSUPPORTED, "knn: K=%d not in 1..32", K);
    prof_note(S, 0);
    dim3 grid((S + KG_T / 32 - 1) / (KG_T / 32), B, 1);
    knn_ring_kernel<<<grid, KG_T, 0, st>>>(q4, g.sorted, g.npad, g.cell_start, g.desc, S, qlen32, K, idx64, idx32, d2out);
    DPM_CHECK_LAUNCH("knn", st);
    return DPM_OK;
}

// ---------------------------------------------------------------------------------------
// surface normals from the points within a radius (open3d estimate_normals with KDTreeSearchParamRadius, as
// LowPassFilter uses it, dataloader/transforms.py:269-272): covariance of the neighbours (the point itself
// included), eigenvector of its smallest eigenvalue; fewer than 3 neighbours -> (0, 0, 1).  One warp per point over
// the 3 x 3 x 3 cell block (cell size >= radius); moments are summed relative to the query in fp32 (|d| <= radius),
// reduced and diagonalised in fp64 (cyclic Jacobi).  The sign of a normal is arbitrary, as in open3d without an
// orientation step; its only consumer takes |n_i . n_j|.
// ---------------------------------------------------------------------------------------
__device__ void smallest_eigvec3(double a00, double a01, double a02, double a11, double a12, double a22, float *n) {
    double A[3][3] = {{a00, a01, a02}, {a01, a11, a12}, {a02, a12, a22}};
    double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 12; ++sweep) {
        const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
        if (off < 1e-30) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (fabs(A[p][q]) < 1e-300) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < 3; ++k) {
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - sn * akq;
                    A[k][q] = sn * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - sn * aqk;
                    A[q][k] = sn * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - sn * vkq;
                    V[k][q] = sn * vkp + c * vkq;
                }
            }
    }
    int m = 0;
    if (A[1][1] < A[m][m]) m = 1;
    if (A[2][2] < A[m][m]) m = 2;
    const double l = sqrt(V[0][m] * V[0][m] + V[1][m] * V[1][m] + V[2][m] * V[2][m]);
    n[0] = (float)(V[0][m] / l); n[1] = (float)(V[1][m] / l); n[2] = (float)(V[2][m] / l);
}

template <bool GRID>
__global__ void __launch_bounds__(KG_T)
radius_normals_kernel(const float4 *__restrict__ q4, const float4 *__restrict__ sorted, const int *__restrict__ cell_start,
                      const GridDesc *__restrict__ desc, int S, int N, float r2, float *__restrict__ normals) {
    const int lane = threadIdx.x & 31;
    const int s = blockIdx.x * (KG_T / 32) + (threadIdx.x >> 5);
    if (s >= S) return;  // warp-uniform
    const float4 c = q4[s];
    float m[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int cnt = 0;
    auto take = [&](const float4 p) {
        const float dx = p.x - c.x, dy = p.y - c.y, dz = p.z - c.z;
        if (dx * dx + dy * dy + dz * dz <= r2) {
            m[0] += dx; m[1] += dy; m[2] += dz;
            m[3] += dx * dx; m[4] += dx * dy; m[5] += dx * dz; m[6] += dy * dy; m[7] += dy * dz; m[8] += dz * dz;
            ++cnt;
        }
    };
    if (GRID) {
        const GridDesc d = desc[0];
        const int cx = cell_coord_free(c.x, d.ox, d.inv_h, d.gx), cy = cell_coord_free(c.y, d.oy, d.inv_h, d.gy),
                  cz = cell_coord_free(c.z, d.oz, d.inv_h, d.gz);
        int rs = 0, re = 0;
        if (lane < 9) {
            const int yy = cy + (lane % 3) - 1, zz = cz + (lane / 3) - 1;
            const int x0 = max(cx - 1, 0), x1 = min(cx + 1, d.gx - 1);
            if (yy >= 0 && yy < d.gy && zz >= 0 && zz < d.gz && x0 <= x1) {
                const int row = (zz * d.gy + yy) * d.gx;
                rs = cell_start[row + x0];
                re = cell_start[row + x1 + 1];
            }
        }
        for (int r = 0; r < 9; ++r) {
            const int b0 = __shfl_sync(0xffffffffu, rs, r), b1 = __shfl_sync(0xffffffffu, re, r);
            for (int i = b0 + lane; i < b1; i += 32) take(sorted[i]);
        }
    } else {
        for (int i = lane; i < N; i += 32) take(q4[i]);
    }
    double acc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = warp_sum_d((double)m[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) {
        float n[3] = {0.f, 0.f, 1.f};
        if (cnt >= 3) {
            const double inv = 1.0 / (double)cnt;
            const double mx = acc[0] * inv, my = acc[1] * inv, mz = acc[2] * inv;
            smallest_eigvec3(acc[3] * inv - mx * mx, acc[4] * inv - mx * my, acc[5] * inv - mx * mz, acc[6] * inv - my * my,
                             acc[7] * inv - my * mz, acc[8] * inv - mz * mz, n);
        }
        normals[(size_t)s * 3] = n[0]; normals[(size_t)s * 3 + 1] = n[1]; normals[(size_t)s * 3 + 2] = n[2];
    }
}

// normals (S,3) of the cloud q4 (S points, one cloud); g == nullptr: no grid, every point is tested (small clouds)
int radius_normals_launch(const GridWs *g, const float4 *q4, int S, float radius, float *normals, cudaStream_t st) {
    if (S <= 0 || !(radius > 0.f)) return fail(DPM_ERR_SHAPE, "normals: bad arguments");
    const float r2 = radius * radius;
    dim3 grid((S + KG_T / 32 - 1) / (KG_T / 32), 1, 1);
    if (g) radius_normals_kernel<true><<<grid, KG_T, 0, st>>>(q4, g->sorted, g->cell_start, g->desc, S, S, r2, normals);
    else radius_normals_kernel<false><<<grid, KG_T, 0, st>>>(q4, nullptr, nullptr, nullptr, S, S, r2, normals);
    DPM_CHECK_LAUNCH("normals", st);
    return DPM_OK;
}

int knn_grid_launch(const GridWs &g, const float4 *q4, const float4 *p4, int B, int S, int N, const int *qlen32,
                    int K, float r2, int64_t *idx64, int32_t *idx32, cudaStream_t st, bool pad) {
    if (B <= 0 || S <= 0 || N <= 0) return fail(DPM_ERR_SHAPE, "knn: bad shape B=%d S=%d N=%d", B, S, N);
    if (K <= 0 || K > 32) return fail(DPM_ERR_UNSUPPORTED, "knn: K=%d not in 1..32", K);
    // keep d2 <= r2  <=>  d2 < nextafter(r2, +inf)
    const float cap = r2 >= 0.f ? __builtin_nextafterf(r2, __builtin_inff()) : 0.f;
    prof_note(S, N);
    dim3 grid((S + KG_T / 32 - 1) / (KG_T / 32), B, 1);
    if (pad) knn_grid_kernel<true><<<grid, KG_T, 0, st>>>(q4, g.sorted, g.npad, g.cell_start, g.desc, p4, S, N, qlen32, K, cap, idx64, idx32);
    else knn_grid_kernel<false><<<grid, KG_T, 0, st>>>(q4, g.sorted, g.npad, g.cell_start, g.desc, p4, S, N, qlen32, K, cap, idx64, idx32);
    DPM_CHECK_LAUNCH("knn", st);
    return DPM_OK;
}

}  // namespace dpm
