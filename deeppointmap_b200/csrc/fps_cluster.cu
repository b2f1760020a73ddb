// fps_cluster.cu -- the LATENCY variants of farthest point sampling: one cloud spread over a cluster of
// 8 CTAs x 4 warps (8 SMs), bit-exact with Sampler.fps (network/encoder/utils.py:209-270) like every other
// FPS kernel here.
//
// Why: the one-CTA-per-cloud kernel of grid.cu keeps 32 warps on ONE SM, so the ~100-instruction chain of a
// pick is issued 8 warps deep per scheduler (issue-active 50 %, 1.2 us per pick, 32 of 148 SMs busy for a
// batch of 32).  That is the right shape when other streams fill the rest of the chip; for a single frame
// (pipeline/infer.py runs batch 1) it leaves 147 SMs idle behind a 4095-pick chain.  Here every warp has a
// scheduler to itself, the cloud is resident in the cluster's shared memory (65 536 points x 20 B = 1.25 MB
// over 8 x 160 KB), and the per-pick arg-max crosses the cluster without a cluster barrier:
//
//   * every warp reduces its own candidate (two REDUX + ballot), then 16 lanes push the 32-byte record
//     (value bits, ~index, xyz) into the record table of EVERY CTA with st.async -- the store itself
//     signals the destination's mbarrier (complete_tx), so there is no separate arrive and no
//     barrier.cluster per pick;
//   * every warp waits on its own CTA's mbarrier (32 records x 32 B), loads one record per lane and
//     reduces them -- all 32 warps reach the same winner without another exchange;
//   * records and mbarriers are double-buffered by pick parity; a warp can only publish pick k+2 after
//     it has seen every warp's record of pick k+1, which those warps sent after reading pick k -- so two
//     buffers are enough and the mbarrier of pick k is re-armed (expect_tx) right after its wait.
//
//   fps_grid_cluster_kernel  : the bucket-pruned FPS of grid.cu (same buckets, same arithmetic), buckets
//                              dealt over the 32 warps of the cluster, tiles in shared memory (N <= 65 536)
//                              or left in L2 (larger clouds).
//   fps_brute_cluster_kernel : clouds of <= 8192 points: every point in registers (<= 8 per thread), all
//                              min-distances updated per pick -- cheaper than a box test at that size.
#include <stdlib.h>

#include <atomic>

#include "common.cuh"

namespace dpm {

// Geometry of a cloud's team: CS CTAs x WPC warps = 32 warps (one record per lane of the reducing warp).
//   Geo<8, 4>  : the latency mapping described above (8 SMs per cloud, one warp per scheduler)
//   Geo<1, 32> : the same algorithm inside ONE CTA of 32 warps (one SM per cloud, st.async to itself + mbarrier instead of a
//                __syncthreads per pick, tiles in L2).  Bit-exact, measured 6.4 ms per 65 536-point cloud against 4.8 ms for
//                grid.cu's fps_grid_kernel, which therefore stays the throughput mapping; kept behind DPM_FPS_ONESM=1.
template <int CS_, int WPC_>
struct Geo {
    static constexpr int CS = CS_, WPC = WPC_, T = WPC_ * 32, NW = CS_ * WPC_;
    static_assert(NW == 32, "the record reduce takes one record per lane");
};
constexpr int FC_NW = 32;                 // warps per cloud = records per pick
constexpr unsigned FC_TX = FC_NW * 32u;   // bytes that complete one pick's mbarrier phase

template <int WPC>
struct __align__(16) FcSharedT {
    uint4 rec[2][FC_NW][2];   // [parity][warp of the team]: {value bits, ~index, x, y}, {z, second value of the warp, -, -}
    uint4 stage[2][WPC][2];   // [parity][warp]: this warp's record, written by its winning lane, pushed out by 2 CS lanes
    uint4 win[2][WPC][3];     // [parity][warp]: the round's result as this warp reduced it: winner {~index, x, y, z},
                              //   {second value of the winner's warp}, runner-up {~index, x, y, z}
    uint4 tmp[2][WPC];        // [toggle][warp]: broadcast slot of the bucket-level arg-max
    unsigned long long bar[2];
};

// developer build (-DDPM_FC_PROFILE): cycles per phase of a pick, summed over the warps of the cluster
#ifdef DPM_FC_PROFILE
__device__ unsigned long long fc_prof[16];
#define FC_TICK(i)                                               \
    do {                                                         \
        const long long _t = clock64();                          \
        pacc[i] += _t - tprev;                                   \
        tprev = _t;                                              \
    } while (0)
#else
#define FC_TICK(i) do { } while (0)
#endif

__device__ __forceinline__ unsigned fc_s32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned fc_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void fc_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned fc_mapa(unsigned local, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
__device__ __forceinline__ void fc_mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// one arrival + `bytes` expected transaction bytes: the phase completes when they have all landed
__device__ __forceinline__ void fc_mbar_arm(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fc_mbar_wait(unsigned bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
// 16-byte store into a peer CTA's shared memory that completes 16 transaction bytes on the peer's mbarrier
__device__ __forceinline__ void fc_st_async16(unsigned raddr, uint4 v, unsigned rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr),
                 "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(rbar)
                 : "memory");
}

template <typename SH>
__device__ __forceinline__ void fc_setup(SH &sh, int tid) {
    if (tid == 0) {
        fc_mbar_init(fc_s32(&sh.bar[0]), 1);
        fc_mbar_init(fc_s32(&sh.bar[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fc_mbar_arm(fc_s32(&sh.bar[0]), FC_TX);
        fc_mbar_arm(fc_s32(&sh.bar[1]), FC_TX);
    }
}

// ---- the arg-max machinery -------------------------------------------------------------------------------
// Candidates are ordered by (value bits, tie key), both "larger wins": value = fp32 min-distance (>= 0, so its bit
// pattern is order preserving), tie key = ~original index (first maximum = lowest index; 0 = "no candidate").
// A warp arg-max is ONE REDUX over the values; the lane that holds the maximum drops its payload into a
// shared-memory slot that the others read back (STS + LDS, ~40 cycles).  Measured alternatives on B200
// (tools/microbench.cu): REDUX ~21 cycles and not pipelined (five REDUX.OR broadcasts: 127), ballot + ffs ~65,
// shfl ~30 -- the classic REDUX, REDUX, ballot, ffs, 3 x shfl chain is ~180.  When several lanes hold the maximum
// (rare: distinct points almost never share a min-distance) they race on the slot and a second REDUX over the tie
// keys picks the one that rewrites it.
//
// TWO PICKS PER ROUND.  The exchange across the cluster (~400 cycles of st.async flight + mbarrier wake-up) is the
// largest item of a pick, so a round tries to settle two: next to its best candidate every lane / bucket / warp
// tracks the VALUE of its second best point, and a warp's record carries the second best value of the warp.  After
// the exchange every warp knows the winner G1, the best G2 of the other 31 warps, and the second value v2 of G1's
// warp.  If value(G2) > v2, G2 is the strict global runner-up; if moreover d2(G1, G2) >= value(G2), inserting G1
// does not lower G2's min-distance -- and since min-distances only ever shrink, G2 is exactly the next farthest
// point (same value, same tie-break) and is emitted in the same round.  Otherwise the round yields one pick, as
// before.  Every warp evaluates the same test on the same 32 records, so the cluster stays in step.  On KITTI-shaped
// clouds ~9 of 10 rounds yield two picks.
struct FcCand {
    unsigned b, n, b2;  // best: value bits, tie key; second best: value bits
    float x, y, z;      // coordinates of the best
};
__device__ __forceinline__ void fc_merge(FcCand &c, unsigned bits, unsigned nid, float x, float y, float z) {
    if (bits > c.b || (bits == c.b && nid > c.n)) {
        c.b2 = c.b; c.b = bits; c.n = nid; c.x = x; c.y = y; c.z = z;
    } else {
        c.b2 = max(c.b2, bits);
    }
}

// (max, second value, arg) of one bucket from the lanes' candidates over its points; only `owner` keeps the result.
// `slot` must not be the slot of the previous call (toggle).
__device__ __forceinline__ void fc_bucket(uint4 *slot, const FcCand &c, bool owner, unsigned &maxbits, unsigned &max2bits,
                                          unsigned &argidx, float &ax, float &ay, float &az) {
    const unsigned vmax = __reduce_max_sync(0xffffffffu, c.b);
    const bool eq = c.b == vmax;
    if (eq) *slot = make_uint4(c.n, __float_as_uint(c.x), __float_as_uint(c.y), __float_as_uint(c.z));
    const unsigned bal = __ballot_sync(0xffffffffu, eq);
    unsigned v2 = __reduce_max_sync(0xffffffffu, eq ? c.b2 : c.b);
    __syncwarp();
    if (bal & (bal - 1u)) {  // warp-uniform
        v2 = vmax;
        const unsigned tmax = __reduce_max_sync(0xffffffffu, eq ? c.n : 0u);
        if (eq && c.n == tmax) *slot = make_uint4(c.n, __float_as_uint(c.x), __float_as_uint(c.y), __float_as_uint(c.z));
        __syncwarp();
    }
    if (owner) {
        const uint4 w = *slot;
        maxbits = vmax; max2bits = v2; argidx = 0xffffffffu - w.x;
        ax = __uint_as_float(w.y); ay = __uint_as_float(w.z); az = __uint_as_float(w.w);
    }
}

// This warp's record of round r: the winning lane writes it, 16 lanes push it into the record table of every CTA of
// the cluster (ra / rb: this lane's remote record / mbarrier address for the parity of r).
template <int CS, typename SH>
__device__ __forceinline__ void fc_publish(SH &sh, int r, int warp, int lane, unsigned ra, unsigned rb, const FcCand &c) {
    uint4 *st = sh.stage[r & 1][warp];
    const unsigned vmax = __reduce_max_sync(0xffffffffu, c.b);
    const bool eq = c.b == vmax;
    if (eq) {
        st[0] = make_uint4(c.n ? vmax : 0u, c.n, __float_as_uint(c.x), __float_as_uint(c.y));
        st[1] = make_uint4(__float_as_uint(c.z), 0u, 0u, 0u);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, eq);
    unsigned v2 = __reduce_max_sync(0xffffffffu, eq ? c.b2 : c.b);  // second value of the warp when one lane holds the best
    __syncwarp();
    if (bal & (bal - 1u)) {  // warp-uniform
        v2 = vmax;
        const unsigned tmax = __reduce_max_sync(0xffffffffu, eq ? c.n : 0u);
        if (eq && c.n == tmax) {
            st[0] = make_uint4(c.n ? vmax : 0u, c.n, __float_as_uint(c.x), __float_as_uint(c.y));
            st[1] = make_uint4(__float_as_uint(c.z), 0u, 0u, 0u);
        }
        __syncwarp();
    }
    if (lane < 2 * CS) {
        uint4 q = st[lane & 1];
        if (lane & 1) q.y = v2;
        fc_st_async16(ra, q, rb);
    }
}

struct FcPick {
    unsigned sel1, sel2;
    float x1, y1, z1, x2, y2, z2;
    bool two;
};
// wait for the 32 records of round r and reduce them: the same result in every warp of the cluster
template <typename SH>
__device__ __forceinline__ void fc_collect(SH &sh, int r, int warp, int lane, bool armer, bool allow_two, FcPick &o) {
    const int par = r & 1;
    const unsigned parity = (unsigned)(((r - 1) >> 1) & 1);
    const unsigned bar = fc_s32(&sh.bar[par]);
    fc_mbar_wait(bar, parity);
    const uint4 a = sh.rec[par][lane][0];
    const uint2 zb = *reinterpret_cast<const uint2 *>(&sh.rec[par][lane][1]);
    if (armer) fc_mbar_arm(bar, FC_TX);  // round r + 2 (nobody can publish it before this CTA has published r + 1)
    uint4 *w = sh.win[par][warp];
    const unsigned m1 = __reduce_max_sync(0xffffffffu, a.x);
    const bool e1 = a.x == m1;
    const unsigned m2 = __reduce_max_sync(0xffffffffu, e1 ? 0u : a.x);
    const bool e2 = !e1 && a.x == m2;
    if (e1) {
        w[0] = make_uint4(a.y, a.z, a.w, zb.x);
        w[1] = make_uint4(zb.y, 0u, 0u, 0u);
    }
    if (e2) w[2] = make_uint4(a.y, a.z, a.w, zb.x);
    const unsigned bal1 = __ballot_sync(0xffffffffu, e1), bal2 = __ballot_sync(0xffffffffu, e2);
    __syncwarp();
    const bool multi = (bal1 & (bal1 - 1u)) != 0u;
    if (multi) {  // warp-uniform: several warps offer the same value, the lowest index wins; one pick this round
        const unsigned tmax = __reduce_max_sync(0xffffffffu, e1 ? a.y : 0u);
        if (e1 && a.y == tmax) w[0] = make_uint4(a.y, a.z, a.w, zb.x);
        __syncwarp();
    }
    const uint4 g1 = w[0], g2 = w[2];
    const unsigned v2w = w[1].x;
    o.sel1 = 0xffffffffu - g1.x;
    o.x1 = __uint_as_float(g1.y); o.y1 = __uint_as_float(g1.z); o.z1 = __uint_as_float(g1.w);
    bool two = allow_two && !multi && bal2 != 0u && (bal2 & (bal2 - 1u)) == 0u && m2 > 0u && m2 > v2w;
    o.sel2 = 0xffffffffu - g2.x;
    o.x2 = __uint_as_float(g2.y); o.y2 = __uint_as_float(g2.z); o.z2 = __uint_as_float(g2.w);
    if (two) two = !(d2_exact(o.x1, o.y1, o.z1, o.x2, o.y2, o.z2) < __uint_as_float(m2));
    o.two = two;
}
// this lane's remote addresses (lanes < 16: destination CTA lane / 2, record half lane & 1) for both parities
template <int CS, typename SH>
__device__ __forceinline__ void fc_remote(SH &sh, int gw, int lane, unsigned (&ra)[2], unsigned (&rb)[2]) {
    const unsigned dest = ((unsigned)lane >> 1) & (CS - 1), half = (unsigned)lane & 1u;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        ra[p] = fc_mapa(fc_s32(&sh.rec[p][gw][half]), dest);
        rb[p] = fc_mapa(fc_s32(&sh.bar[p]), dest);
    }
}

// ---------------------------------------------------------------------------------------------------------
// bucket-pruned FPS over the cell-sorted cloud (see fps_grid_kernel in grid.cu for the pruning argument):
// bucket id = lane * 32 + (warp of the cluster), one bucket of 32 * PPL points per thread.
// RES: the warp's 32 buckets (points + min-distances) live in this CTA's shared memory.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fc_emit(size_t o, unsigned sel, float x, float y, float z, int lane, int64_t *idx64,
                                        int32_t *idx32, float4 *new_xyz4, uint8_t *new_pad) {  // one store per lane
    if (lane == 0 && idx64) idx64[o] = (int64_t)sel;
    if (lane == 1 && idx32) idx32[o] = (int32_t)sel;
    if (lane == 2 && new_xyz4) new_xyz4[o] = make_float4(x, y, z, 0.f);
    if (lane == 3 && new_pad) new_pad[o] = 0;
}

template <int PPL, bool RES, typename G>
__global__ void __launch_bounds__(G::T, 1)
fps_grid_cluster_kernel(const float4 *__restrict__ sorted, float *__restrict__ mind, int npad,
                        const GridDesc *__restrict__ desc, const float4 *__restrict__ xyz4, int N, int K,
                        int64_t *__restrict__ idx64, int32_t *__restrict__ idx32, float4 *__restrict__ new_xyz4,
                        uint8_t *__restrict__ new_pad, int *__restrict__ new_len32) {
    constexpr int BS = 32 * PPL, D = PPL <= 2 ? 2 : 1;
    constexpr int FC_CS = G::CS, FC_WPC = G::WPC, FC_T = G::T;
    using FcShared = FcSharedT<G::WPC>;
    extern __shared__ __align__(16) unsigned char fc_smem[];
    FcShared &sh = *reinterpret_cast<FcShared *>(fc_smem);
    float4 *spts = reinterpret_cast<float4 *>(fc_smem + sizeof(FcShared));  // [FC_WPC * 32 * BS]
    float *smin = reinterpret_cast<float *>(spts + (RES ? FC_WPC * 32 * BS : 0));

    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned rank = FC_CS > 1 ? fc_rank() : 0u;
    const int gw = (int)rank * FC_WPC + warp;
    const int len = desc[b].nvalid;
    const int kn = min(len, K);
    const int nb = (len + BS - 1) / BS;  // <= 1024
    const float4 *P = sorted + (size_t)b * npad;
    float *M = mind + (size_t)b * npad;
    const size_t ob = (size_t)b * K;
    const float INF = __int_as_float(0x7f800000);
    fc_setup(sh, tid);
    unsigned ra[2], rb[2];
    fc_remote<FC_CS>(sh, gw, lane, ra, rb);
    int tog = 0;

    // ---- prologue: tiles -> shared memory, bucket boxes, min-distances = +inf (sentinels 0) ----
    float lox = 0.f, loy = 0.f, loz = 0.f, hix = 0.f, hiy = 0.f, hiz = 0.f, ax = 0.f, ay = 0.f, az = 0.f;
    unsigned maxbits = 0u, max2bits = 0u, argidx = 0xffffffffu;
    const bool owns = lane * FC_NW + gw < nb;
    for (int L = 0; L < 32; ++L) {
        const int bk = L * FC_NW + gw;
        if (bk >= nb) break;  // warp-uniform
        float l0 = INF, l1 = INF, l2 = INF, h0 = -INF, h1 = -INF, h2 = -INF;
        unsigned mi = 0xffffffffu;
#pragma unroll
        for (int j = 0; j < PPL; ++j) {
            const int i = bk * BS + j * 32 + lane;
            const float4 p = P[i];
            const unsigned id = __float_as_uint(p.w);
            const bool valid = id != 0x7fffffffu;
            if (RES) {
                spts[(warp * 32 + L) * BS + j * 32 + lane] = p;
                smin[(warp * 32 + L) * BS + j * 32 + lane] = valid ? INF : 0.f;
            } else {
                M[i] = valid ? INF : 0.f;
            }
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
            lox = l0; loy = l1; loz = l2; hix = h0; hiy = h1; hiz = h2;
            maxbits = max2bits = 0x7f800000u;  // +inf: every bucket is touched by the first sample
            argidx = mi;
        }
    }

    // samples picked but not yet applied to the min-distances: s1, and s2 when the last round settled two picks
    float s1x = 0.f, s1y = 0.f, s1z = 0.f, s2x = 0.f, s2y = 0.f, s2z = 0.f;
    bool two = false;
    if (len > 0) {
        const float4 p0 = xyz4[(size_t)b * N];
        s1x = p0.x; s1y = p0.y; s1z = p0.z;
    }
    if (gw == 0 && kn > 0) fc_emit(ob, 0u, s1x, s1y, s1z, lane, idx64, idx32, new_xyz4, new_pad);
    if (FC_CS > 1) fc_cluster_sync();  // every CTA's mbarriers are initialised and armed before the first record arrives
    else __syncthreads();

#ifdef DPM_FC_PROFILE
    long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
    int rounds = 0;
#endif
    int r = 1;
    for (int k = 1; k < kn; ++r) {
        // ---- phase A, the critical path: box tests, update of the touched tiles, this warp's candidate ----
        FC_TICK(7);
        bool act = false;
        if (owns) {  // can my bucket change?  (d2 >= lb for every point of the box)
            const float mb = __uint_as_float(maxbits);
            float dx = fmaxf(fmaxf(lox - s1x, s1x - hix), 0.f);
            float dy = fmaxf(fmaxf(loy - s1y, s1y - hiy), 0.f);
            float dz = fmaxf(fmaxf(loz - s1z, s1z - hiz), 0.f);
            act = (dx * dx + dy * dy + dz * dz) * 0.99999f < mb;  // conservative lower bound of every d2
            if (two) {
                dx = fmaxf(fmaxf(lox - s2x, s2x - hix), 0.f);
                dy = fmaxf(fmaxf(loy - s2y, s2y - hiy), 0.f);
                dz = fmaxf(fmaxf(loz - s2z, s2z - hiz), 0.f);
                act = act || (dx * dx + dy * dy + dz * dz) * 0.99999f < mb;
            }
        }
        const unsigned mask0 = __ballot_sync(0xffffffffu, act);
        // lane-level candidate: the cached (maximum, second value) of my bucket while it is untouched, merged with my
        // points of every touched tile of the warp
        FcCand c;
        const bool cached = owns && !act && argidx != 0xffffffffu;
        c.b = cached ? maxbits : 0u; c.n = cached ? 0xffffffffu - argidx : 0u; c.b2 = cached ? max2bits : 0u;
        c.x = ax; c.y = ay; c.z = az;
        unsigned mask = mask0;
        FC_TICK(0);
        while (mask) {
            int Ls[D];
            float4 p[D][PPL];
            float m[D][PPL];
#pragma unroll
            for (int d = 0; d < D; ++d) {
                Ls[d] = -1;
                if (mask) {
                    Ls[d] = 31 - __clz(mask);
                    mask &= ~(1u << Ls[d]);
#pragma unroll
                    for (int j = 0; j < PPL; ++j) {
                        if (RES) {
                            const int o = (warp * 32 + Ls[d]) * BS + j * 32 + lane;
                            p[d][j] = spts[o];
                            m[d][j] = smin[o];
                        } else {
                            const int o = (Ls[d] * FC_NW + gw) * BS + j * 32 + lane;
                            p[d][j] = P[o];
                            m[d][j] = M[o];
                        }
                    }
                }
            }
#pragma unroll
            for (int d = 0; d < D; ++d) {
                if (Ls[d] < 0) continue;  // warp-uniform
                FcCand t;
                t.b = t.n = t.b2 = 0u; t.x = t.y = t.z = 0.f;
#pragma unroll
                for (int j = 0; j < PPL; ++j) {
                    float nm = fminf(m[d][j], d2_exact(s1x, s1y, s1z, p[d][j].x, p[d][j].y, p[d][j].z));
                    if (two) nm = fminf(nm, d2_exact(s2x, s2y, s2z, p[d][j].x, p[d][j].y, p[d][j].z));
                    if (nm < m[d][j]) {
                        if (RES) smin[(warp * 32 + Ls[d]) * BS + j * 32 + lane] = nm;
                        else M[(Ls[d] * FC_NW + gw) * BS + j * 32 + lane] = nm;
                    }
                    // nm >= 0: the bit pattern is order preserving
                    fc_merge(t, __float_as_uint(nm), 0xffffffffu - __float_as_uint(p[d][j].w), p[d][j].x, p[d][j].y, p[d][j].z);
                }
                fc_merge(c, t.b, t.n, t.x, t.y, t.z);
                c.b2 = max(c.b2, t.b2);
                if (!RES) {  // tiles in L2: the bucket's new maximum now, while its points are in registers
                    fc_bucket(&sh.tmp[tog][warp], t, lane == Ls[d], maxbits, max2bits, argidx, ax, ay, az);
                    tog ^= 1;
                }
            }
        }
        FC_TICK(1);
        fc_publish<FC_CS>(sh, r, warp, lane, (r & 1) ? ra[1] : ra[0], (r & 1) ? rb[1] : rb[0], c);
        FC_TICK(3);
        // ---- phase B, in the shadow of the exchange: the touched buckets' new maxima, for their owners ----
        if (RES) {
            mask = mask0;
            while (mask) {
                const int L = 31 - __clz(mask);
                mask &= ~(1u << L);
                FcCand t;
                t.b = t.n = t.b2 = 0u; t.x = t.y = t.z = 0.f;
#pragma unroll
                for (int j = 0; j < PPL; ++j) {
                    const int o = (warp * 32 + L) * BS + j * 32 + lane;
                    const float4 q = spts[o];
                    fc_merge(t, __float_as_uint(smin[o]), 0xffffffffu - __float_as_uint(q.w), q.x, q.y, q.z);
                }
                fc_bucket(&sh.tmp[tog][warp], t, lane == L, maxbits, max2bits, argidx, ax, ay, az);
                tog ^= 1;
            }
        }
        // ---- the round's result: one pick, or two ----
        FC_TICK(4);
        FcPick o;
        fc_collect(sh, r, warp, lane, tid == 0, k + 1 < kn, o);
        FC_TICK(5);
        if (gw == 0) {
            fc_emit(ob + k, o.sel1, o.x1, o.y1, o.z1, lane, idx64, idx32, new_xyz4, new_pad);
            if (o.two) fc_emit(ob + k + 1, o.sel2, o.x2, o.y2, o.z2, lane, idx64, idx32, new_xyz4, new_pad);
        }
        s1x = o.x1; s1y = o.y1; s1z = o.z1;
        s2x = o.x2; s2y = o.y2; s2z = o.z2;
        two = o.two;
        k += two ? 2 : 1;
#ifdef DPM_FC_PROFILE
        ++rounds;
#endif
    }
#ifdef DPM_FC_PROFILE
    if (lane == 0 && b == 0)
        for (int i = 0; i < 8; ++i) atomicAdd(&fc_prof[i], (unsigned long long)pacc[i]);
    if (tid == 0 && rank == 0 && b == 0) {
        atomicAdd(&fc_prof[8], (unsigned long long)(kn - 1));
        atomicAdd(&fc_prof[9], (unsigned long long)rounds);
    }
#endif
    if (rank == 0) {
        for (int k = kn + tid; k < K; k += FC_T) {  // K > len: idx -1, zero rows, padded
            if (idx64) idx64[ob + k] = -1;
            if (idx32) idx32[ob + k] = -1;
            if (new_xyz4) new_xyz4[ob + k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (new_pad) new_pad[ob + k] = 1;
        }
        if (tid == 0 && new_len32) new_len32[b] = kn;
    }
    if (FC_CS > 1) fc_cluster_sync();  // no CTA may exit while a peer can still store into its shared memory
}

// ---------------------------------------------------------------------------------------------------------
// small clouds: P points per thread in registers, point i = j * 1024 + (warp of the cluster) * 32 + lane
// ---------------------------------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(Geo<8, 4>::T, 1)
fps_brute_cluster_kernel(const float4 *__restrict__ xyz4, int N, const int *__restrict__ len32, int K,
                         int64_t *__restrict__ idx64, int32_t *__restrict__ idx32, float4 *__restrict__ new_xyz4,
                         uint8_t *__restrict__ new_pad, int *__restrict__ new_len32) {
    constexpr int FC_CS = 8, FC_WPC = 4, FC_T = 128;
    __shared__ FcSharedT<4> sh;
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned rank = fc_rank();
    const int gw = (int)rank * FC_WPC + warp;
    const int len = len32 ? min(len32[b], N) : N;
    const int kn = min(len, K);
    const float4 *pts = xyz4 + (size_t)b * N;
    const size_t ob = (size_t)b * K;
    fc_setup(sh, tid);
    unsigned ra[2], rb[2];
    fc_remote<FC_CS>(sh, gw, lane, ra, rb);

    float x[P], y[P], z[P], m[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
        const int i = j * (FC_NW * 32) + gw * 32 + lane;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        m[j] = 0.f;  // points past `len` never win: min-distance 0 and a higher index than any valid point
        if (i < len) {
            p = pts[i];
            m[j] = __int_as_float(0x7f800000);
        }
        x[j] = p.x; y[j] = p.y; z[j] = p.z;
    }
    float s1x = 0.f, s1y = 0.f, s1z = 0.f, s2x = 0.f, s2y = 0.f, s2z = 0.f;
    bool two = false;
    if (len > 0) {
        const float4 p0 = pts[0];
        s1x = p0.x; s1y = p0.y; s1z = p0.z;
    }
    if (gw == 0 && kn > 0) fc_emit(ob, 0u, s1x, s1y, s1z, lane, idx64, idx32, new_xyz4, new_pad);
    fc_cluster_sync();

    int r = 1;
    for (int k = 1; k < kn; ++r) {
        FcCand c;
        c.b = c.n = c.b2 = 0u; c.x = c.y = c.z = 0.f;
#pragma unroll
        for (int j = P - 1; j >= 0; --j) {  // descending j: among this thread's ties the lowest index is merged last and wins
            m[j] = fminf(m[j], d2_exact(s1x, s1y, s1z, x[j], y[j], z[j]));
            if (two) m[j] = fminf(m[j], d2_exact(s2x, s2y, s2z, x[j], y[j], z[j]));
            // m >= 0: the bit pattern is order preserving; tie key = ~index (never 0: index < 2^31)
            fc_merge(c, __float_as_uint(m[j]), 0xffffffffu - (unsigned)(j * (FC_NW * 32) + gw * 32 + lane), x[j], y[j], z[j]);
        }
        fc_publish<FC_CS>(sh, r, warp, lane, (r & 1) ? ra[1] : ra[0], (r & 1) ? rb[1] : rb[0], c);
        FcPick o;
        fc_collect(sh, r, warp, lane, tid == 0, k + 1 < kn, o);
        if (gw == 0) {
            fc_emit(ob + k, o.sel1, o.x1, o.y1, o.z1, lane, idx64, idx32, new_xyz4, new_pad);
            if (o.two) fc_emit(ob + k + 1, o.sel2, o.x2, o.y2, o.z2, lane, idx64, idx32, new_xyz4, new_pad);
        }
        s1x = o.x1; s1y = o.y1; s1z = o.z1;
        s2x = o.x2; s2y = o.y2; s2z = o.z2;
        two = o.two;
        k += two ? 2 : 1;
    }
    if (rank == 0) {
        for (int k = kn + tid; k < K; k += FC_T) {
            if (idx64) idx64[ob + k] = -1;
            if (idx32) idx32[ob + k] = -1;
            if (new_xyz4) new_xyz4[ob + k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (new_pad) new_pad[ob + k] = 1;
        }
        if (tid == 0 && new_len32) new_len32[b] = kn;
    }
    fc_cluster_sync();
}

// ---------------------------------------------------------------------------------------------------------
// policy + launchers
// ---------------------------------------------------------------------------------------------------------
static std::atomic<int> g_fps_mode{0};  // 0 auto, 1 one CTA per cloud, 2 cluster per cloud, 3 two clouds per CTA (packed)

using GeoCluster = Geo<8, 4>;   // latency mapping
using GeoOneSm = Geo<1, 32>;    // throughput mapping
constexpr int FC_CS = GeoCluster::CS, FC_T = GeoCluster::T;

template <typename G>
static size_t fc_grid_smem(int ppl, bool res) {
    return sizeof(FcSharedT<G::WPC>) + (res ? (size_t)G::WPC * 32 * 32 * ppl * (sizeof(float4) + sizeof(float)) : 0);
}

// clouds whose clusters are co-resident: what the hardware can place of the largest cluster kernel (8 CTAs x 166 KB
// of shared memory, one GPC per cluster), asked once per device
int fps_cluster_capacity() {
    static thread_local int cached[64];
    const int dev = current_device() & 63;
    if (cached[dev] == 0) {
        int n = 0;
        auto kern = fps_grid_cluster_kernel<2, true, GeoCluster>;
        const size_t smem = fc_grid_smem<GeoCluster>(2, true);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(FC_CS, 64, 1);
        cfg.blockDim = dim3(FC_T, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = FC_CS;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) {
            cudaGetLastError();
            n = device_sm_count() / (2 * FC_CS);  // conservative
        }
        cached[dev] = n < 1 ? 1 : n;
    }
    return cached[dev];
}

bool fps_cluster_mode(int B) {
    static const char *env = getenv("DPM_FPS_MODE");  // developer A/B switch: 1 / 2 as dpm_set_fps_mode
    int mode = g_fps_mode.load(std::memory_order_relaxed);
    if (mode == 0 && env) mode = atoi(env);
    if (mode == 1 || mode == 3) return false;
    if (mode == 2) return true;
    // few clouds: 8 SMs each, all clusters resident at once.  Larger batches are throughput work: one SM per cloud
    // and the rest of the chip for the other kernels / streams (a second wave of clusters would double the latency).
    return B <= fps_cluster_capacity();
}

// clouds of <= FPS_BRUTE_CLUSTER_MAX_N points (every level below the first): the register-resident cluster kernel needs
// no shared-memory tiles, so 16 of its CTAs fit on an SM and a whole batch of clusters is co-resident
bool fps_cluster_mode_small(int B) {
    static const char *env = getenv("DPM_FPS_MODE");
    static const char *env_small = getenv("DPM_FPS_SMALL_MAXB");
    int mode = g_fps_mode.load(std::memory_order_relaxed);
    if (mode == 0 && env) mode = atoi(env);
    if (mode == 1 || mode == 3) return false;
    if (mode == 2) return true;
    // measured at 32 clouds per step: 256 light cluster CTAs instead of 32 one-SM CTAs halve the level-1 FPS (0.78 -> 0.40 ms)
    // but crowd the other streams' kernels: 6778 -> 6450 frames/s.  So the default is the same bound as the big clouds'.
    const int maxb = env_small ? atoi(env_small) : 0;
    return B <= (maxb > fps_cluster_capacity() ? maxb : fps_cluster_capacity());
}

template <typename G, typename Kern, typename... Args>
static int fc_launch(Kern kern, size_t smem, int B, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(G::CS, B, 1);
    cfg.blockDim = dim3(G::T, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = G::CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;   // also for one CTA per cloud: the records travel by st.async to shared::cluster addresses
    DPM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, args...));
    count_launch("fps", st);
    return DPM_OK;
}

template <int PPL, bool RES, typename G>
static int fps_grid_cluster_t(const GridWs &g, const float4 *xyz4, int B, int N, int K, int64_t *idx64, int32_t *idx32,
                              float4 *new_xyz4, uint8_t *new_pad, int *new_len32, cudaStream_t st) {
    auto kern = fps_grid_cluster_kernel<PPL, RES, G>;
    const size_t smem = fc_grid_smem<G>(PPL, RES);
    static thread_local unsigned long long configured = 0ull;  // one bit per device: function attributes are per context
    const unsigned long long devbit = 1ull << (current_device() & 63);
    if (!(configured & devbit)) {
        DPM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured |= devbit;
    }
    return fc_launch<G>(kern, smem, B, st, (const float4 *)g.sorted, g.mind, g.npad, (const GridDesc *)g.desc, xyz4, N, K, idx64,
                        idx32, new_xyz4, new_pad, new_len32);
}

int fps_grid_cluster_launch(const GridWs &g, const float4 *xyz4, int B, int N, int K, int64_t *idx64, int32_t *idx32,
                            float4 *new_xyz4, uint8_t *new_pad, int *new_len32, cudaStream_t st) {
    if (N > GRID_CLUSTER_MAX_N) return fail(DPM_ERR_UNSUPPORTED, "fps: N=%d exceeds the limit %d", N, GRID_CLUSTER_MAX_N);
    prof_note(N, K);
    switch (grid_ppl(N)) {
        case 1: return fps_grid_cluster_t<1, true, GeoCluster>(g, xyz4, B, N, K, idx64, idx32, new_xyz4, new_pad, new_len32, st);
        case 2: return fps_grid_cluster_t<2, true, GeoCluster>(g, xyz4, B, N, K, idx64, idx32, new_xyz4, new_pad, new_len32, st);
        case 4: return fps_grid_cluster_t<4, false, GeoCluster>(g, xyz4, B, N, K, idx64, idx32, new_xyz4, new_pad, new_len32, st);
        case 8: return fps_grid_cluster_t<8, false, GeoCluster>(g, xyz4, B, N, K, idx64, idx32, new_xyz4, new_pad, new_len32, st);
    }
    return fail(DPM_ERR_UNSUPPORTED, "fps: no cluster kernel for N=%d", N);
}

// the same algorithm with one CTA (one SM) per cloud, tiles in L2: the throughput mapping
int fps_grid_onesm_launch(const GridWs &g, const float4 *xyz4, int B, int N, int K, int64_t *idx64, int32_t *idx32,
                          float4 *new_xyz4, uint8_t *new_pad, int *new_len32, cudaStream_t st) {
    if (N > GRID_CLUSTER_MAX_N) return fail(DPM_ERR_UNSUPPORTED, "fps: N=%d exceeds the limit %d", N, GRID_CLUSTER_MAX_N);
    prof_note(N, K);
    switch (grid_ppl(N)) {
        case 1: return fps_grid_cluster_t<1, false, GeoOneSm>(g, xyz4, B, N, K, idx64, idx32, new_xyz4, new_pad, new_len32, st);
        case 2: return fps_grid_cluster_t<2, false, GeoOneSm>(g, xyz4, B, N, K, idx64, idx32, new_xyz4, new_pad, new_len32, st);
        case 4: return fps_grid_cluster_t<4, false, GeoOneSm>(g, xyz4, B, N, K, idx64, idx32, new_xyz4, new_pad, new_len32, st);
        case 8: return fps_grid_cluster_t<8, false, GeoOneSm>(g, xyz4, B, N, K, idx64, idx32, new_xyz4, new_pad, new_len32, st);
    }
    return fail(DPM_ERR_UNSUPPORTED, "fps: no kernel for N=%d", N);
}

int fps_brute_cluster_launch(const float4 *xyz4, int B, int N, const int *len32, int K, int64_t *idx64, int32_t *idx32,
                             float4 *new_xyz4, uint8_t *new_pad, int *new_len32, cudaStream_t st) {
    prof_note(N, K);
#define DPM_FB_CASE(p)                                                                                              \
    if (N <= p * FC_NW * 32)                                                                                        \
        return fc_launch<GeoCluster>(fps_brute_cluster_kernel<p>, 0, B, st, xyz4, N, len32, K, idx64, idx32, new_xyz4, new_pad, \
                                     new_len32);
    DPM_FB_CASE(1) DPM_FB_CASE(2) DPM_FB_CASE(4) DPM_FB_CASE(8)
#undef DPM_FB_CASE
    return fail(DPM_ERR_UNSUPPORTED, "fps: N=%d exceeds the register-resident cluster limit %d", N, FPS_BRUTE_CLUSTER_MAX_N);
}

}  // namespace dpm

#ifdef DPM_FC_PROFILE
extern "C" int dpm_debug_fc_profile(unsigned long long *out16, int reset) {
    if (cudaMemcpyFromSymbol(out16, dpm::fc_prof, sizeof(unsigned long long) * 16) != cudaSuccess) return -1;
    if (reset) {
        unsigned long long z[16] = {0};
        cudaMemcpyToSymbol(dpm::fc_prof, z, sizeof(z));
    }
    return 0;
}
#endif
extern "C" int dpm_fps_cluster_capacity(void) { return dpm::fps_cluster_capacity(); }
namespace dpm {
// mode 3 ("packed"): the one-SM kernel with TWO clouds per SM (two 512-thread teams in one CTA): 22 % less SM time for the
// FPS of a batch at 1.55x its latency -- pays when the caller keeps >= 8 streams of batches in flight (bench.py), loses
// on fewer (one stream: 9.0 instead of 5.8 ms per 32 clouds), hence opt-in
bool fps_packed_mode() {
    static const bool env = getenv("DPM_FPS_PAIR") && atoi(getenv("DPM_FPS_PAIR")) == 1;
    return env || g_fps_mode.load(std::memory_order_relaxed) == 3;
}
}  // namespace dpm
extern "C" void dpm_set_fps_mode(int mode) { dpm::g_fps_mode.store(mode < 0 || mode > 3 ? 0 : mode, std::memory_order_relaxed); }
