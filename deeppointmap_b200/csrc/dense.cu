// dense.cu -- row-major fp32 building blocks: linear (+bias/residual/ReLU), LayerNorm,
// the fused group kernel (gather + geometry term + LayerNorm + ReLU + max over K) and the
// feature-propagation interpolation.  fp32 SIMT: the 1e-4 parity tolerance rules out plain
// TF32/BF16 tensor-core math, and after hoisting the 1x1 conv in front of the gather the
// whole encoder is ~0.65 GFLOP/frame.
#include "common.cuh"

namespace dpm {

// ---------------------------------------------------------------------------------------
// Y = act(X W^T + bias + res)      X (M,K) ldx, W (N,K) ldw, Y (M,N) ldy
// 64x64x16 tiles, 256 threads, 4x4 outputs per thread.
// ---------------------------------------------------------------------------------------
constexpr int GM = 64, GN = 64, GK = 16;

template <bool VEC>
__device__ __forceinline__ void load_tile(const float *__restrict__ A, int lda, int rows, int K, int r0, int k0,
                                          float (*sm)[GM + 4], int tid) {
    const int r = tid >> 2, kq = (tid & 3) * 4;
    const int gr = r0 + r;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (gr < rows) {
        const float *p = A + (size_t)gr * lda + k0 + kq;
        if (VEC && k0 + kq + 3 < K) {
            const float4 t = *reinterpret_cast<const float4 *>(p);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (k0 + kq + i < K) v[i] = p[i];
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) sm[kq + i][r] = v[i];
}

template <bool VX, bool VW>
__global__ void __launch_bounds__(256)
linear_kernel(const float *__restrict__ X, int ldx, const float *__restrict__ W, int ldw,
              const float *__restrict__ bias, const float *__restrict__ res, int ldres, float *__restrict__ Y,
              int ldy, int M, int N, int K, int act, long long sX, long long sW, long long sY) {
    X += (size_t)blockIdx.z * sX;
    W += (size_t)blockIdx.z * sW;
    Y += (size_t)blockIdx.z * sY;
    if (res) res += (size_t)blockIdx.z * sY;
    __shared__ float As[GK][GM + 4];
    __shared__ float Bs[GK][GN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * GM, n0 = blockIdx.y * GN;  // M tiles on x: no 65535 limit
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += GK) {
        load_tile<VX>(X, ldx, M, K, m0, k0, As, tid);
        load_tile<VW>(W, ldw, N, K, n0, k0, Bs, tid);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GK; ++k) {
            const float4 a = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (bias) v += bias[n];
            if (res) v += res[(size_t)m * ldres + n];
            if (act == DPM_ACT_RELU) v = fmaxf(v, 0.f);
            Y[(size_t)m * ldy + n] = v;
        }
    }
}

int linear_batched_launch(const float *X, int ldx, long long sX, const float *W, int ldw, long long sW,
                          const float *bias, const float *res, int ldres, float *Y, int ldy, long long sY, int M,
                          int N, int K, int nbatch, int act, cudaStream_t st) {
    if (M <= 0 || N <= 0 || K <= 0 || nbatch <= 0) return fail(DPM_ERR_SHAPE, "linear: bad shape M=%d N=%d K=%d", M, N, K);
    // the tensor-core path is taken whenever the operand layout allows it, whatever M is (a short tile is
    // zero-filled): the K-summation order of an output element never depends on how many rows share the
    // launch, so a unit's result is bit-identical whatever else is batched with it.
    if (linear_tc_eligible(X, ldx, sX, W, ldw, sW, M, N, K) && (nbatch == 1 || (sY & 3) == 0))
        return linear_tc_launch(X, ldx, sX, W, ldw, sW, bias, res, ldres, Y, ldy, sY, M, N, K, nbatch, act, st);
    dim3 grid((M + GM - 1) / GM, (N + GN - 1) / GN, nbatch);
    prof_note((long long)M * nbatch, (long long)N * K);
    const bool vx = (ldx % 4 == 0) && (((uintptr_t)X & 15) == 0) && (sX % 4 == 0);
    const bool vw = (ldw % 4 == 0) && (((uintptr_t)W & 15) == 0) && (sW % 4 == 0);
    if (vx && vw) linear_kernel<true, true><<<grid, 256, 0, st>>>(X, ldx, W, ldw, bias, res, ldres, Y, ldy, M, N, K, act, sX, sW, sY);
    else if (vx) linear_kernel<true, false><<<grid, 256, 0, st>>>(X, ldx, W, ldw, bias, res, ldres, Y, ldy, M, N, K, act, sX, sW, sY);
    else if (vw) linear_kernel<false, true><<<grid, 256, 0, st>>>(X, ldx, W, ldw, bias, res, ldres, Y, ldy, M, N, K, act, sX, sW, sY);
    else linear_kernel<false, false><<<grid, 256, 0, st>>>(X, ldx, W, ldw, bias, res, ldres, Y, ldy, M, N, K, act, sX, sW, sY);
    DPM_CHECK_LAUNCH("linear", st);
    return DPM_OK;
}

int linear_launch(const float *X, int ldx, const float *W, int ldw, const float *bias, const float *res, int ldres,
                  float *Y, int ldy, int M, int N, int K, int act, cudaStream_t st) {
    return linear_batched_launch(X, ldx, 0, W, ldw, 0, bias, res, ldres, Y, ldy, 0, M, N, K, 1, act, st);
}

// ---------------------------------------------------------------------------------------
// LayerNorm over the row (eps 1e-5), optional post-add and ReLU.  Warp per row.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
layernorm_kernel(const float *X, int ldx, const float *__restrict__ gamma, const float *__restrict__ beta,
                 const float *post, int ldpost, float *Y, int ldy, int M, int C, int act) {  // X may alias Y
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= M) return;
    const float *x = X + (size_t)row * ldx;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += x[c];
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) {
        const float d = x[c] - mean;
        q = fmaf(d, d, q);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)C + 1e-5f);
    for (int c = lane; c < C; c += 32) {
        float v = (x[c] - mean) * rstd * gamma[c] + beta[c];
        if (post) v += post[(size_t)row * ldpost + c];
        if (act == DPM_ACT_RELU) v = fmaxf(v, 0.f);
        Y[(size_t)row * ldy + c] = v;
    }
}

int layernorm_launch(const float *X, int ldx, const float *gamma, const float *beta, const float *post, int ldpost,
                     float *Y, int ldy, int M, int C, int act, cudaStream_t st) {
    if (M <= 0 || C <= 0) return fail(DPM_ERR_SHAPE, "layernorm: bad shape M=%d C=%d", M, C);
    layernorm_kernel<<<(M + 7) / 8, 256, 0, st>>>(X, ldx, gamma, beta, post, ldpost, Y, ldy, M, C, act);
    DPM_CHECK_LAUNCH("layernorm", st);
    return DPM_OK;
}

// ---------------------------------------------------------------------------------------
// group kernel: warp per centre; lanes over output channels (CPL per lane); loop over K.
//   y_k[c] = Zfea[idx_k][c] + ((xyz[idx_k]-ctr)/r) . Wxyz[c]      (bias already inside Zfea)
//   out[c] = max_k relu(LN_c(y_k) * gamma[c] + beta[c])
// ---------------------------------------------------------------------------------------
template <int CPL>
__global__ void __launch_bounds__(256)
group_kernel(const float *__restrict__ Z, const float4 *__restrict__ xyz4, const float4 *__restrict__ ctr4,
             const int32_t *__restrict__ gidx, const float *__restrict__ Wxyz, int ldw,
             const float *__restrict__ gamma, const float *__restrict__ beta, float radius,
             float *__restrict__ out, int N, int S, int K) {
    constexpr int C = CPL * 32;
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int s = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (s >= S) return;
    const float4 c4 = ctr4[(size_t)b * S + s];
    const int myidx = lane < K ? gidx[((size_t)b * S + s) * K + lane] : 0;
    float wx[CPL], wy[CPL], wz[CPL], g[CPL], be[CPL], best[CPL];
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
        const int c = lane + 32 * i;
        wx[i] = Wxyz[(size_t)c * ldw];
        wy[i] = Wxyz[(size_t)c * ldw + 1];
        wz[i] = Wxyz[(size_t)c * ldw + 2];
        g[i] = gamma[c];
        be[i] = beta[c];
        best[i] = 0.f;  // max over K of relu(.) == relu(max): start from 0
    }
    const float *Zb = Z + (size_t)b * N * C;
    const float4 *pb = xyz4 + (size_t)b * N;
    for (int k = 0; k < K; ++k) {
        const int j = __shfl_sync(0xffffffffu, myidx, k);
        const float4 p = pb[j];
        const float dx = __fdiv_rn(__fsub_rn(p.x, c4.x), radius);
        const float dy = __fdiv_rn(__fsub_rn(p.y, c4.y), radius);
        const float dz = __fdiv_rn(__fsub_rn(p.z, c4.z), radius);
        const float *zr = Zb + (size_t)j * C;
        float y[CPL];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            y[i] = zr[lane + 32 * i];
            y[i] = fmaf(dx, wx[i], y[i]);
            y[i] = fmaf(dy, wy[i], y[i]);
            y[i] = fmaf(dz, wz[i], y[i]);
            sum += y[i];
        }
        const float mean = warp_sum(sum) * (1.0f / (float)C);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            y[i] -= mean;
            q = fmaf(y[i], y[i], q);
        }
        const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / (float)C) + 1e-5f);
#pragma unroll
        for (int i = 0; i < CPL; ++i) best[i] = fmaxf(best[i], fmaf(y[i] * rstd, g[i], be[i]));
    }
    float *o = out + ((size_t)b * S + s) * C;
#pragma unroll
    for (int i = 0; i < CPL; ++i) o[lane + 32 * i] = best[i];
}

// ---------------------------------------------------------------------------------------
// group kernel, lane = neighbour (Cout <= 128): every lane pulls ITS neighbour's whole row into
// registers (independent 16-byte loads, no per-neighbour shuffles), does the LayerNorm locally,
// and the max over the K lanes is a butterfly transpose-reduce (C-1 shuffles in total) that
// leaves C/32 output channels per lane.  Wxyz / gamma / beta are staged in shared memory and
// read as warp-wide broadcasts.  (A row-coalesced variant -- C/4 lanes per neighbour row, 32/(C/4) rows per
// step, LayerNorm reduced across those lanes -- was measured at 1.28 ms/step against 0.86 for this one: the
// kernel is bound by the shuffle / dependent-math chain per neighbour, not by how the rows arrive.)
// ---------------------------------------------------------------------------------------
//
// FROMXYZ: the gathered features are themselves an affine map of the neighbour's coordinates
// (stage 0: the stem Conv1d 3->width, encoder.py:53, feeds the SA conv with nothing in between), so
// Z is not a per-point matrix but the composed table comp[C][4] = (Wfea.W0 | Wfea.b0 + b) made by
// compose_stem_kernel, and the "gather" is 3 more FMAs per channel on the neighbour's own xyz.
template <int C, bool FROMXYZ>
__global__ void __launch_bounds__(128)
group_lane_kernel(const float *__restrict__ Z, const float4 *__restrict__ xyz4, const float4 *__restrict__ ctr4,
                  const int32_t *__restrict__ gidx, const float *__restrict__ Wxyz, int ldw,
                  const float *__restrict__ gamma, const float *__restrict__ beta, float radius,
                  float *__restrict__ out, int N, int S, int K) {
    __shared__ __align__(16) float sw[FROMXYZ ? 9 : 5][C];  // wx, wy, wz, gamma, beta [, ax, ay, az, a0]
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    for (int c = tid; c < C; c += 128) {
        sw[0][c] = Wxyz[(size_t)c * ldw];
        sw[1][c] = Wxyz[(size_t)c * ldw + 1];
        sw[2][c] = Wxyz[(size_t)c * ldw + 2];
        sw[3][c] = gamma[c];
        sw[4][c] = beta[c];
        if (FROMXYZ) {
            const float4 a = reinterpret_cast<const float4 *>(Z)[c];
            sw[5][c] = a.x; sw[6][c] = a.y; sw[7][c] = a.z; sw[8][c] = a.w;
        }
    }
    __syncthreads();
    const int s = blockIdx.x * 4 + (tid >> 5);
    if (s >= S) return;
    const float4 c4 = ctr4[(size_t)b * S + s];
    float v[C];
    const bool act = lane < K;
    if (act) {
        const int j = gidx[((size_t)b * S + s) * K + lane];
        const float4 p = xyz4[(size_t)b * N + j];
        const float dx = __fdiv_rn(__fsub_rn(p.x, c4.x), radius);
        const float dy = __fdiv_rn(__fsub_rn(p.y, c4.y), radius);
        const float dz = __fdiv_rn(__fsub_rn(p.z, c4.z), radius);
        const float4 *zr = reinterpret_cast<const float4 *>(Z + ((size_t)b * N + j) * C);
        float sum = 0.f;
#pragma unroll
        for (int q = 0; q < C / 4; ++q) {
            float4 z;
            if (FROMXYZ) {
                const float4 ax = *reinterpret_cast<const float4 *>(&sw[5][4 * q]);
                const float4 ay = *reinterpret_cast<const float4 *>(&sw[6][4 * q]);
                const float4 az = *reinterpret_cast<const float4 *>(&sw[7][4 * q]);
                const float4 a0 = *reinterpret_cast<const float4 *>(&sw[8][4 * q]);
                z.x = fmaf(p.z, az.x, fmaf(p.y, ay.x, fmaf(p.x, ax.x, a0.x)));
                z.y = fmaf(p.z, az.y, fmaf(p.y, ay.y, fmaf(p.x, ax.y, a0.y)));
                z.z = fmaf(p.z, az.z, fmaf(p.y, ay.z, fmaf(p.x, ax.z, a0.z)));
                z.w = fmaf(p.z, az.w, fmaf(p.y, ay.w, fmaf(p.x, ax.w, a0.w)));
            } else {
                z = zr[q];
            }
            const float4 wx = *reinterpret_cast<const float4 *>(&sw[0][4 * q]);
            const float4 wy = *reinterpret_cast<const float4 *>(&sw[1][4 * q]);
            const float4 wz = *reinterpret_cast<const float4 *>(&sw[2][4 * q]);
            v[4 * q + 0] = fmaf(dz, wz.x, fmaf(dy, wy.x, fmaf(dx, wx.x, z.x)));
            v[4 * q + 1] = fmaf(dz, wz.y, fmaf(dy, wy.y, fmaf(dx, wx.y, z.y)));
            v[4 * q + 2] = fmaf(dz, wz.z, fmaf(dy, wy.z, fmaf(dx, wx.z, z.z)));
            v[4 * q + 3] = fmaf(dz, wz.w, fmaf(dy, wy.w, fmaf(dx, wx.w, z.w)));
            sum += (v[4 * q] + v[4 * q + 1]) + (v[4 * q + 2] + v[4 * q + 3]);
        }
        const float mean = sum * (1.0f / (float)C);
        float q2 = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            v[c] -= mean;
            q2 = fmaf(v[c], v[c], q2);
        }
        const float rstd = 1.0f / sqrtf(q2 * (1.0f / (float)C) + 1e-5f);
#pragma unroll
        for (int q = 0; q < C / 4; ++q) {
            const float4 g = *reinterpret_cast<const float4 *>(&sw[3][4 * q]);
            const float4 be = *reinterpret_cast<const float4 *>(&sw[4][4 * q]);
            v[4 * q + 0] = fmaxf(fmaf(v[4 * q + 0] * rstd, g.x, be.x), 0.f);
            v[4 * q + 1] = fmaxf(fmaf(v[4 * q + 1] * rstd, g.y, be.y), 0.f);
            v[4 * q + 2] = fmaxf(fmaf(v[4 * q + 2] * rstd, g.z, be.z), 0.f);
            v[4 * q + 3] = fmaxf(fmaf(v[4 * q + 3] * rstd, g.w, be.w), 0.f);
        }
    } else {
#pragma unroll
        for (int c = 0; c < C; ++c) v[c] = 0.f;  // identity of max over relu(.)
    }
    // butterfly: after the step with xor-distance d a lane keeps the half of its channels selected by
    // its bit d, so lane l ends with channels [l*C/32, (l+1)*C/32)
    int h = C / 2;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const bool up = (lane & d) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
            const float send = up ? v[i] : v[i + h];
            const float keep = up ? v[i + h] : v[i];
            v[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, d));
        }
        h >>= 1;
    }
    float *o = out + ((size_t)b * S + s) * C + lane * (C / 32);
#pragma unroll
    for (int i = 0; i < C / 32; ++i) o[i] = v[i];
}

// comp[c] = (sum_k Wfea[c][k] W0[k][0..2], sum_k Wfea[c][k] b0[k] + b[c]): the stem folded into the
// per-point half of the stage-0 SA conv (fp64 accumulation, rounded once)
__global__ void compose_stem_kernel(const float *__restrict__ Wfea, int ldw, const float *__restrict__ bias,
                                    const float *__restrict__ W0, const float *__restrict__ b0, int width, int Cout,
                                    float4 *__restrict__ comp) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= Cout) return;
    double a[4] = {0.0, 0.0, 0.0, bias ? (double)bias[c] : 0.0};
    for (int k = 0; k < width; ++k) {
        const double w = Wfea[(size_t)c * ldw + k];
        a[0] += w * W0[k * 3 + 0];
        a[1] += w * W0[k * 3 + 1];
        a[2] += w * W0[k * 3 + 2];
        if (b0) a[3] += w * b0[k];
    }
    comp[c] = make_float4((float)a[0], (float)a[1], (float)a[2], (float)a[3]);
}

bool group_from_xyz_supported(int Cout) { return Cout == 32 || Cout == 64 || Cout == 128; }

int group_from_xyz_launch(const float *Wsa, int ldw, const float *bias, const float *W0, const float *b0, int width,
                          float4 *comp, const float4 *xyz4, const float4 *ctr4, const int32_t *gidx, const float *gamma,
                          const float *beta, float radius, float *out, int B, int N, int S, int K, int Cout,
                          cudaStream_t st) {
    if (B <= 0 || N <= 0 || S <= 0) return fail(DPM_ERR_SHAPE, "group: bad shape");
    if (K <= 0 || K > 32) return fail(DPM_ERR_UNSUPPORTED, "group: K=%d not in 1..32", K);
    if (!group_from_xyz_supported(Cout)) return fail(DPM_ERR_UNSUPPORTED, "group_from_xyz: Cout=%d", Cout);
    compose_stem_kernel<<<(Cout + 127) / 128, 128, 0, st>>>(Wsa, ldw, bias, W0, b0, width, Cout, comp);
    DPM_CHECK_LAUNCH("compose_stem", st);
    prof_note(S, Cout);
    const float *Z = reinterpret_cast<const float *>(comp);
    const float *Wxyz = Wsa + width;
    dim3 g4((S + 3) / 4, B, 1);
    if (Cout == 32) group_lane_kernel<32, true><<<g4, 128, 0, st>>>(Z, xyz4, ctr4, gidx, Wxyz, ldw, gamma, beta, radius, out, N, S, K);
    else if (Cout == 64) group_lane_kernel<64, true><<<g4, 128, 0, st>>>(Z, xyz4, ctr4, gidx, Wxyz, ldw, gamma, beta, radius, out, N, S, K);
    else group_lane_kernel<128, true><<<g4, 128, 0, st>>>(Z, xyz4, ctr4, gidx, Wxyz, ldw, gamma, beta, radius, out, N, S, K);
    DPM_CHECK_LAUNCH("group", st);
    return DPM_OK;
}

// sum of `ks` partial products (+ bias + res) -> LayerNorm -> (+ post) -> activation; warp per row
__global__ void __launch_bounds__(256)
splitk_ln_kernel(const float *__restrict__ part, int ks, long long pstride, const float *__restrict__ bias,
                 const float *res, int ldres, const float *__restrict__ gamma, const float *__restrict__ beta,
                 const float *post, int ldpost, float *Y, int ldy, int M, int C, int act) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= M) return;
    auto val = [&](int c) {
        float v = part[(size_t)row * C + c];
        for (int z = 1; z < ks; ++z) v += part[(size_t)z * pstride + (size_t)row * C + c];  // fixed order: deterministic
        if (bias) v += bias[c];
        if (res) v += res[(size_t)row * ldres + c];
        return v;
    };
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += val(c);
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) {
        const float d = val(c) - mean;
        q = fmaf(d, d, q);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)C + 1e-5f);
    for (int c = lane; c < C; c += 32) {
        float v = (val(c) - mean) * rstd * gamma[c] + beta[c];
        if (post) v += post[(size_t)row * ldpost + c];
        if (act == DPM_ACT_RELU) v = fmaxf(v, 0.f);
        Y[(size_t)row * ldy + c] = v;
    }
}

// tmp: scratch of tmp_copies x M x N floats.  With tmp_copies >= 2 and a long contraction (K >= 1024) that cannot use the
// fused epilogue, the product is computed as tmp_copies SPLIT-K partial GEMMs in one launch (K chunks as the batch
// dimension of the tcgen05 kernel) and summed with IEEE adds inside the LayerNorm kernel: the few row tiles of the deepest
// stage spread over 4x as many CTAs, and the tensor core's truncating accumulate runs over 64 k-steps instead of 256
// (2e-4 -> < 1e-4 of a channel's rms at the 512-channel stage, tools/probe_stage_error.py).
int linear_ln_launch(const float *X, int ldx, const float *W, int ldw, const float *bias, const float *res, int ldres,
                     const float *gamma, const float *beta, const float *post, int ldpost, float *tmp, float *Y, int ldy,
                     int M, int N, int K, int act, cudaStream_t st, int tmp_copies) {
    int rc = DPM_OK;
    if (linear_ln_tc_launch(X, ldx, W, ldw, bias, res, ldres, gamma, beta, post, ldpost, Y, ldy, M, N, K, act, st, &rc, tmp))
        return rc;
    int ks = tmp_copies;
    while (ks > 1 && (K % (ks * 32) != 0)) --ks;
    if (ks >= 2 && K >= 1024 && ldx == K && ldw == K && tmp != Y &&
        linear_tc_eligible(X, ldx, K / ks, W, ldw, K / ks, M, N, K / ks)) {
        const int Kc = K / ks;
        DPM_TRY(linear_tc_launch(X, ldx, Kc, W, ldw, Kc, nullptr, nullptr, 0, tmp, N, (long long)M * N, M, N, Kc, ks,
                                 DPM_ACT_NONE, st));
        splitk_ln_kernel<<<(M + 7) / 8, 256, 0, st>>>(tmp, ks, (long long)M * N, bias, res, ldres, gamma, beta, post, ldpost, Y,
                                                      ldy, M, N, act);
        DPM_CHECK_LAUNCH("layernorm", st);
        return DPM_OK;
    }
    DPM_TRY(linear_launch(X, ldx, W, ldw, bias, res, ldres, tmp, N, M, N, K, DPM_ACT_NONE, st));
    return layernorm_launch(tmp, N, gamma, beta, post, ldpost, Y, ldy, M, N, act, st);
}

int group_launch(const float *Z, const float4 *xyz4, const float4 *ctr4, const int32_t *gidx, const float *Wxyz,
                 int ldw, const float *gamma, const float *beta, float radius, float *out, int B, int N, int S,
                 int K, int Cout, cudaStream_t st) {
    if (B <= 0 || N <= 0 || S <= 0) return fail(DPM_ERR_SHAPE, "group: bad shape");
    if (K <= 0 || K > 32) return fail(DPM_ERR_UNSUPPORTED, "group: K=%d not in 1..32", K);
    prof_note(S, Cout);
    if ((Cout == 32 || Cout == 64 || Cout == 128) && (((uintptr_t)Z & 15) == 0)) {
        dim3 g4((S + 3) / 4, B, 1);
        if (Cout == 32) group_lane_kernel<32, false><<<g4, 128, 0, st>>>(Z, xyz4, ctr4, gidx, Wxyz, ldw, gamma, beta, radius, out, N, S, K);
        else if (Cout == 64) group_lane_kernel<64, false><<<g4, 128, 0, st>>>(Z, xyz4, ctr4, gidx, Wxyz, ldw, gamma, beta, radius, out, N, S, K);
        else group_lane_kernel<128, false><<<g4, 128, 0, st>>>(Z, xyz4, ctr4, gidx, Wxyz, ldw, gamma, beta, radius, out, N, S, K);
        DPM_CHECK_LAUNCH("group", st);
        return DPM_OK;
    }
    dim3 grid((S + 7) / 8, B, 1);
#define DPM_GROUP_CASE(cpl)                                                                                     \
    case cpl * 32:                                                                                              \
        group_kernel<cpl><<<grid, 256, 0, st>>>(Z, xyz4, ctr4, gidx, Wxyz, ldw, gamma, beta, radius, out, N, S, K); \
        break;
    switch (Cout) {
        DPM_GROUP_CASE(1) DPM_GROUP_CASE(2) DPM_GROUP_CASE(4) DPM_GROUP_CASE(8) DPM_GROUP_CASE(16)
        default:
            return fail(DPM_ERR_UNSUPPORTED, "group: Cout=%d not in {32,64,128,256,512}", Cout);
    }
#undef DPM_GROUP_CASE
    DPM_CHECK_LAUNCH("group", st);
    return DPM_OK;
}

// ---------------------------------------------------------------------------------------
// feature propagation: 3-NN (direct differences) inverse-squared-distance interpolation +
// concat([fea1, interp]).  Warp per target point.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fp_interp_kernel(const float4 *__restrict__ xyz1, const float4 *__restrict__ xyz2, const float *__restrict__ fea1,
                 const float *__restrict__ fea2, const uint8_t *__restrict__ pad2, float *__restrict__ out, int N,
                 int S, int C1, int C2) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (n >= N) return;
    const float4 c = xyz1[(size_t)b * N + n];
    const float4 *p2 = xyz2 + (size_t)b * S;
    float *o = out + ((size_t)b * N + n) * (C1 + C2);
    const float *f1 = fea1 + ((size_t)b * N + n) * C1;
    for (int ch = lane; ch < C1; ch += 32) o[ch] = f1[ch];
    const float *f2 = fea2 + (size_t)b * S * C2;
    if (S == 1) {
        for (int ch = lane; ch < C2; ch += 32) o[C1 + ch] = f2[ch];
        return;
    }
    // lane-local best three (ascending by (d2, idx)), then three rounds of warp arg-min
    unsigned long long k0 = ~0ull, k1 = ~0ull, k2 = ~0ull;
    for (int i = lane; i < S; i += 32) {
        const float4 p = p2[i];
        float d = d2_exact(c.x, c.y, c.z, p.x, p.y, p.z);
        if (pad2 && pad2[(size_t)b * S + i]) d = 1e30f;  // reference pushes padded points far away
        const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)i;
        if (key < k0) { k2 = k1; k1 = k0; k0 = key; }
        else if (key < k1) { k2 = k1; k1 = key; }
        else if (key < k2) { k2 = key; }
    }
    float dsel[3];
    int isel[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        unsigned long long m = k0;
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) {
            const unsigned long long o2 = __shfl_xor_sync(0xffffffffu, m, sft);
            m = o2 < m ? o2 : m;
        }
        dsel[r] = __uint_as_float((unsigned)(m >> 32));
        isel[r] = (int)(unsigned)m;
        if (k0 == m && m != ~0ull) { k0 = k1; k1 = k2; k2 = ~0ull; }
    }
    const int cnt = S < 3 ? S : 3;
    float w[3] = {0.f, 0.f, 0.f}, wsum = 0.f;
    for (int r = 0; r < cnt; ++r) {
        w[r] = 1.0f / fmaxf(dsel[r], 1e-8f);
        wsum += w[r];
    }
    for (int r = 0; r < cnt; ++r) w[r] = w[r] / wsum;
    for (int ch = lane; ch < C2; ch += 32) {
        float acc = 0.f;
        for (int r = 0; r < cnt; ++r) acc += f2[(size_t)isel[r] * C2 + ch] * w[r];
        o[C1 + ch] = acc;
    }
}

int fp_interp_launch(const float4 *xyz1, const float4 *xyz2, const float *fea1, const float *fea2,
                     const uint8_t *pad2, float *out, int B, int N, int S, int C1, int C2, cudaStream_t st) {
    if (B <= 0 || N <= 0 || S <= 0 || C1 < 0 || C2 <= 0) return fail(DPM_ERR_SHAPE, "fp_interp: bad shape");
    dim3 grid((N + 7) / 8, B, 1);
    fp_interp_kernel<<<grid, 256, 0, st>>>(xyz1, xyz2, fea1, fea2, pad2, out, N, S, C1, C2);
    DPM_CHECK_LAUNCH("fp_interp", st);
    return DPM_OK;
}

}  // namespace dpm

using namespace dpm;

extern "C" int dpm_linear_f32(const float *X, int ldx, const float *W, int ldw, const float *bias, const float *res,
                              int ldres, float *Y, int ldy, int M, int N, int K, int act, dpm_stream_t stream) {
    if (!X || !W || !Y) return fail(DPM_ERR_ARG, "linear: null pointer");
    split_begin();  // no pre-split weights outside the encoder / decoder calls
    return linear_launch(X, ldx, W, ldw, bias, res, ldres, Y, ldy, M, N, K, act, (cudaStream_t)stream);
}

extern "C" size_t dpm_linear_workspace_bytes(int N, int K) {
    Arena a(nullptr, 0);
    a.get<float>(split_floats(N > 0 ? N : 0, K > 0 ? K : 0));
    return a.off + 256;
}

extern "C" int dpm_linear_ws_f32(const float *X, int ldx, const float *W, int ldw, const float *bias, const float *res,
                                 int ldres, float *Y, int ldy, int M, int N, int K, int act, void *ws, size_t ws_bytes,
                                 dpm_stream_t stream) {
    if (!X || !W || !Y || !ws) return fail(DPM_ERR_ARG, "linear: null pointer");
    if (ws_bytes < dpm_linear_workspace_bytes(N, K)) return fail(DPM_ERR_WORKSPACE, "linear: workspace too small");
    Arena a(ws, ws_bytes);
    split_begin();
    split_add(a, W, N, K, ldw);
    DPM_TRY(split_run((cudaStream_t)stream));
    const int rc = linear_launch(X, ldx, W, ldw, bias, res, ldres, Y, ldy, M, N, K, act, (cudaStream_t)stream);
    split_begin();
    return rc;
}

extern "C" size_t dpm_linear_ln_workspace_bytes(int M, int N, int K) {
    Arena a(nullptr, 0);
    a.get<float>(split_floats(N > 0 ? N : 0, K > 0 ? K : 0));
    a.get<float>((size_t)(M > 0 ? M : 0) * (N > 0 ? N : 0));
    return a.off + 256;
}

extern "C" int dpm_linear_ln_ws_f32(const float *X, int ldx, const float *W, int ldw, const float *bias,
                                    const float *res, int ldres, const float *gamma, const float *beta,
                                    const float *post, int ldpost, float *Y, int ldy, int M, int N, int K, int act,
                                    void *ws, size_t ws_bytes, dpm_stream_t stream) {
    if (!X || !W || !Y || !gamma || !beta || !ws) return fail(DPM_ERR_ARG, "linear_ln: null pointer");
    if (M <= 0 || N <= 0 || K <= 0) return fail(DPM_ERR_SHAPE, "linear_ln: bad shape M=%d N=%d K=%d", M, N, K);
    if (ws_bytes < dpm_linear_ln_workspace_bytes(M, N, K)) return fail(DPM_ERR_WORKSPACE, "linear_ln: workspace too small");
    Arena a(ws, ws_bytes);
    split_begin();
    split_add(a, W, N, K, ldw);
    float *tmp = a.get<float>((size_t)M * N);
    DPM_TRY(split_run((cudaStream_t)stream));
    const int rc = linear_ln_launch(X, ldx, W, ldw, bias, res, ldres, gamma, beta, post, ldpost, tmp, Y, ldy, M, N, K, act,
                                    (cudaStream_t)stream);
    split_begin();
    return rc;
}

extern "C" int dpm_layernorm_f32(const float *X, int ldx, const float *gamma, const float *beta, const float *post,
                                 int ldpost, float *Y, int ldy, int M, int C, int act, dpm_stream_t stream) {
    if (!X || !gamma || !beta || !Y) return fail(DPM_ERR_ARG, "layernorm: null pointer");
    return layernorm_launch(X, ldx, gamma, beta, post, ldpost, Y, ldy, M, C, act, (cudaStream_t)stream);
}

extern "C" int dpm_group_ln_relu_max_f32(const float *Zfea, const float *xyz4, const float *ctr4, const int32_t *gidx,
                                         const float *Wxyz, int ldw, const float *gamma, const float *beta,
                                         float radius, float *out, int B, int N, int S, int K, int Cout,
                                         dpm_stream_t stream) {
    if (!Zfea || !xyz4 || !ctr4 || !gidx || !Wxyz || !gamma || !beta || !out) return fail(DPM_ERR_ARG, "group: null pointer");
    return group_launch(Zfea, (const float4 *)xyz4, (const float4 *)ctr4, gidx, Wxyz, ldw, gamma, beta, radius, out, B,
                        N, S, K, Cout, (cudaStream_t)stream);
}

extern "C" int dpm_fp_interp_f32(const float *xyz1_4, const float *xyz2_4, const float *fea1, const float *fea2,
                                 const uint8_t *pad2, float *out, int B, int N, int S, int C1, int C2,
                                 dpm_stream_t stream) {
    if (!xyz1_4 || !xyz2_4 || !fea1 || !fea2 || !out) return fail(DPM_ERR_ARG, "fp_interp: null pointer");
    return fp_interp_launch((const float4 *)xyz1_4, (const float4 *)xyz2_4, fea1, fea2, pad2, out, B, N, S, C1, C2,
                            (cudaStream_t)stream);
}
