// microbench.cu -- latencies of the primitives the FPS pick chain is made of (B200, sm_100a):
// warp collectives, shared memory, and three ways of crossing an 8-CTA cluster.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu && tools/microbench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ unsigned s32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// ---- dependent chains of warp collectives (one warp) ----
__global__ void chain_kernel(int iters, long long *out, unsigned *sink) {
    const int lane = threadIdx.x;
    unsigned v = lane * 2654435761u + 17u;
    long long t0, t1;
    // REDUX max
    t0 = clock64();
    for (int i = 0; i < iters; ++i) v = __reduce_max_sync(0xffffffffu, v ^ (unsigned)lane) + (unsigned)i;
    t1 = clock64();
    if (lane == 0) out[0] = t1 - t0;
    // shfl
    t0 = clock64();
    for (int i = 0; i < iters; ++i) v = __shfl_sync(0xffffffffu, v, (v + i) & 31) + 1u;
    t1 = clock64();
    if (lane == 0) out[1] = t1 - t0;
    // ballot + ffs
    t0 = clock64();
    for (int i = 0; i < iters; ++i) v = __ffs(__ballot_sync(0xffffffffu, ((v + lane) & 3u) == 0u) | 0x80000000u) + v;
    t1 = clock64();
    if (lane == 0) out[2] = t1 - t0;
    // shfl_xor butterfly max over a 64-bit key (5 steps)
    unsigned long long key = ((unsigned long long)v << 32) | lane;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, o);
            key = ok > key ? ok : key;
        }
        key += (unsigned long long)lane << 33;
    }
    t1 = clock64();
    if (lane == 0) out[3] = t1 - t0;
    // shared-memory load chain
    __shared__ unsigned sm[1024];
    for (int i = lane; i < 1024; i += 32) sm[i] = (i * 37 + 11) & 1023;
    __syncwarp();
    unsigned a = lane;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) a = sm[a];
    t1 = clock64();
    if (lane == 0) out[4] = t1 - t0;
    // match_any
    t0 = clock64();
    for (int i = 0; i < iters; ++i) v = __match_any_sync(0xffffffffu, v & 7u) + v;
    t1 = clock64();
    if (lane == 0) out[5] = t1 - t0;
    // fp32 dependent add chain
    float f = (float)lane;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) f = __fadd_rn(f, 1.5f);
    t1 = clock64();
    if (lane == 0) out[6] = t1 - t0;
    sink[lane] = v + a + (unsigned)key + (unsigned)f;
}

// ---- cluster exchange ----
struct __align__(16) Sh {
    uint4 rec[2][32][2];
    unsigned long long bar[2];
    uint4 crec[2][4][2];  // CTA-level staging for the hierarchical variant
};
__device__ __forceinline__ unsigned rank_() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void csync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ unsigned mapa(unsigned l, unsigned r) { unsigned o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(l), "r"(r)); return o; }
__device__ __forceinline__ void arm(unsigned bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void waitp(unsigned bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void st_async16(unsigned ra, uint4 v, unsigned rb) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(ra), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(rb) : "memory");
}
__device__ __forceinline__ void st_remote16(unsigned ra, uint4 v) {
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ra), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_vol16(const uint4 *p) {
    uint4 v;
    asm volatile("ld.volatile.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(s32(p)) : "memory");
    return v;
}

// variant 0: st.async + mbarrier, 32 records direct (what fps_cluster.cu does)
// variant 1: plain remote stores carrying a pick tag, readers poll their record with volatile loads
// variant 2: cluster barrier per pick (what fps.cu does)
// variant 3: one 16-byte st.async per record (value, index, packed position): 16 B records
template <int V>
__global__ void __launch_bounds__(128, 1) exch_kernel(int iters, long long *out, unsigned *sink) {
    __shared__ Sh sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned rank = rank_();
    const int gw = rank * 4 + warp;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&sh.bar[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&sh.bar[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const unsigned tx = V == 3 ? 32 * 16 : 32 * 32;
        arm(s32(&sh.bar[0]), tx);
        arm(s32(&sh.bar[1]), tx);
    }
    for (int i = tid; i < 2 * 32 * 2; i += 128) (&sh.rec[0][0][0])[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    csync();
    unsigned v = gw * 977u + 13u;
    float x = (float)gw;
    const long long t0 = clock64();
    for (int k = 1; k <= iters; ++k) {
        const int par = k & 1;
        const unsigned parity = (unsigned)(((k - 1) >> 1) & 1);
        // warp-uniform candidate (as after the warp arg-max)
        const unsigned vb = (v * 2654435761u) >> 3, vi = (unsigned)gw;
        uint4 a;
        unsigned zb = 0;
        if (V == 0 || V == 3) {
            if (V == 0) {
                if (lane < 16) {
                    const unsigned dest = lane >> 1, half = lane & 1;
                    const uint4 r = half ? make_uint4(__float_as_uint(x), 0, 0, 0) : make_uint4(vb, 0xffffffffu - vi, __float_as_uint(x), __float_as_uint(x));
                    st_async16(mapa(s32(&sh.rec[par][gw][half]), dest), r, mapa(s32(&sh.bar[par]), dest));
                }
            } else {
                if (lane < 8) st_async16(mapa(s32(&sh.rec[par][gw][0]), lane), make_uint4(vb, 0xffffffffu - vi, __float_as_uint(x), 0), mapa(s32(&sh.bar[par]), lane));
            }
            const unsigned bar = s32(&sh.bar[par]);
            waitp(bar, parity);
            a = sh.rec[par][lane][0];
            if (V == 0) zb = sh.rec[par][lane][1].x;
            if (tid == 0) arm(bar, V == 3 ? 32 * 16 : 32 * 32);
        } else if (V == 1) {
            if (lane < 16) {
                const unsigned dest = lane >> 1, half = lane & 1;
                const uint4 r = half ? make_uint4(__float_as_uint(x), __float_as_uint(x), (unsigned)k, (unsigned)k) : make_uint4(vb, 0xffffffffu - vi, __float_as_uint(x), (unsigned)k);
                st_remote16(mapa(s32(&sh.rec[par][gw][half]), dest), r);
            }
            uint4 b;
            do {
                a = ld_vol16(&sh.rec[par][lane][0]);
                b = ld_vol16(&sh.rec[par][lane][1]);
            } while (!__all_sync(0xffffffffu, a.w == (unsigned)k && b.w == (unsigned)k));
            zb = b.x;
        } else {
            if (lane < 16) {
                const unsigned dest = lane >> 1, half = lane & 1;
                const uint4 r = half ? make_uint4(__float_as_uint(x), 0, 0, 0) : make_uint4(vb, 0xffffffffu - vi, __float_as_uint(x), __float_as_uint(x));
                st_remote16(mapa(s32(&sh.rec[par][gw][half]), dest), r);
            }
            csync();
            a = sh.rec[par][lane][0];
            zb = sh.rec[par][lane][1].x;
        }
        const unsigned mh = __reduce_max_sync(0xffffffffu, a.x);
        const unsigned bal = __ballot_sync(0xffffffffu, a.x == mh);
        int src = __ffs(bal) - 1;
        unsigned ml = __shfl_sync(0xffffffffu, a.y, src);
        if (bal & (bal - 1)) {
            ml = __reduce_max_sync(0xffffffffu, a.x == mh ? a.y : 0u);
            src = __ffs(__ballot_sync(0xffffffffu, a.x == mh && a.y == ml)) - 1;
        }
        x = __uint_as_float(__shfl_sync(0xffffffffu, a.z, src)) + __uint_as_float(__shfl_sync(0xffffffffu, zb, src)) * 0.f + 1.f;
        v = mh + ml + (unsigned)k;
    }
    const long long t1 = clock64();
    if (rank == 0 && tid == 0) out[0] = t1 - t0;
    sink[blockIdx.x * 128 + tid] = v + (unsigned)x;
    csync();
}

template <int V>
static int run_exch(const char *name, int iters, long long *out, unsigned *sink) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(8, 1, 1);
    cfg.blockDim = dim3(128, 1, 1);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 8; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaLaunchKernelEx(&cfg, exch_kernel<V>, iters, out, sink));
        CK(cudaDeviceSynchronize());
    }
    long long h;
    CK(cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost));
    printf("exchange %-34s %8.1f cycles / pick\n", name, (double)h / iters);
    return 0;
}

int main() {
    long long *out;
    unsigned *sink;
    CK(cudaMalloc(&out, 64 * 8));
    CK(cudaMalloc(&sink, 4096 * 4));
    const int iters = 4096;
    for (int rep = 0; rep < 2; ++rep) {
        chain_kernel<<<1, 32>>>(iters, out, sink);
        CK(cudaDeviceSynchronize());
    }
    long long h[8];
    CK(cudaMemcpy(h, out, 7 * 8, cudaMemcpyDeviceToHost));
    const char *names[7] = {"REDUX.max + add", "shfl + add", "ballot + ffs + add", "64-bit butterfly max (5 shfl.b64)", "LDS chain", "match_any + add", "FADD"};
    for (int i = 0; i < 7; ++i) printf("chain %-36s %7.1f cycles / op\n", names[i], (double)h[i] / iters);
    if (run_exch<0>("st.async + mbarrier, 32 B records", iters, out, sink)) return 1;
    if (run_exch<3>("st.async + mbarrier, 16 B records", iters, out, sink)) return 1;
    if (run_exch<1>("remote st + tag polling", iters, out, sink)) return 1;
    if (run_exch<2>("remote st + barrier.cluster", iters, out, sink)) return 1;
    int dev = 0, clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    printf("sm clock (attr) %d kHz\n", clk);
    return 0;
}
