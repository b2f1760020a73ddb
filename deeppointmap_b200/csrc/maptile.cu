// maptile.cu -- the scan-to-map input stage on the device:
// PoseGraph.__global_mapping + the centring of global_map_query_graph
// (system/modules/pose_graph.py:373-409, 499-511): for the m <= 16 key-frames around a scan, move every
// key-frame's 256 descriptor coordinates into the world with its pose, concatenate the (131, 256)
// descriptor sets along the point axis and re-centre the coordinates on the querying scan's pose:
//     tile[:, i*S:(i+1)*S] = [ fea_i ;  Rc^T ( (R_i xyz_i + t_i) - tc ) ]
// The reference does this with per-scan .to(device) / matmul / concat / .cpu(); here the descriptor sets stay
// in a device-resident store (n, Cd, S) and the tile (Cd, m*S) is written by one launch, ready to be the
// `dst` of dpm_registration_forward (M = 256, N = m*256).
#include "common.cuh"

namespace dpm {

__global__ void __launch_bounds__(256)
map_tile_kernel(const float *__restrict__ store, int n_store, int Cd, int S, const int32_t *__restrict__ ids,
                const float *__restrict__ poses, const float *__restrict__ center, int m, float *__restrict__ tile) {
    const int i = blockIdx.y;  // key-frame slot
    const int id = ids[i];
    const float *src = store + (size_t)id * Cd * S;
    const int Cf = Cd - 3;
    const size_t ld = (size_t)m * S;
    const bool ok = id >= 0 && id < n_store;
    // features: plain copy of Cf rows
    for (int e = blockIdx.x * 256 + threadIdx.x; e < Cf * S; e += gridDim.x * 256) {
        const int c = e / S, s = e - c * S;
        tile[(size_t)c * ld + (size_t)i * S + s] = ok ? src[(size_t)c * S + s] : 0.f;
    }
    // coordinates: world = R_i p + t_i (pose_graph.py:391), then Rc^T (world - tc) (pose_graph.py:507)
    const float *T = poses + (size_t)i * 16;
    for (int s = blockIdx.x * 256 + threadIdx.x; s < S; s += gridDim.x * 256) {
        float x = 0.f, y = 0.f, z = 0.f;
        if (ok) {
            const float px = src[(size_t)Cf * S + s], py = src[(size_t)(Cf + 1) * S + s], pz = src[(size_t)(Cf + 2) * S + s];
            x = fmaf(T[2], pz, fmaf(T[1], py, T[0] * px)) + T[3];
            y = fmaf(T[6], pz, fmaf(T[5], py, T[4] * px)) + T[7];
            z = fmaf(T[10], pz, fmaf(T[9], py, T[8] * px)) + T[11];
            if (center) {
                const float dx = x - center[3], dy = y - center[7], dz = z - center[11];
                x = fmaf(center[8], dz, fmaf(center[4], dy, center[0] * dx));  // rows of Rc^T = columns of Rc
                y = fmaf(center[9], dz, fmaf(center[5], dy, center[1] * dx));
                z = fmaf(center[10], dz, fmaf(center[6], dy, center[2] * dx));
            }
        }
        tile[(size_t)Cf * ld + (size_t)i * S + s] = x;
        tile[(size_t)(Cf + 1) * ld + (size_t)i * S + s] = y;
        tile[(size_t)(Cf + 2) * ld + (size_t)i * S + s] = z;
    }
}

}  // namespace dpm

using namespace dpm;

extern "C" int dpm_map_tile_f32(const float *store, int n_store, int Cd, int S, const int32_t *ids, const float *poses,
                                const float *center, int m, float *tile, dpm_stream_t stream) {
    if (!store || !ids || !poses || !tile) return fail(DPM_ERR_ARG, "map_tile: null pointer");
    if (n_store <= 0 || Cd < 4 || S <= 0 || m <= 0) return fail(DPM_ERR_SHAPE, "map_tile: bad shape n=%d Cd=%d S=%d m=%d", n_store, Cd, S, m);
    cudaStream_t st = (cudaStream_t)stream;
    prof_mark(st);
    dim3 grid(((Cd - 3) * S + 1023) / 1024 > 0 ? ((Cd - 3) * S + 1023) / 1024 : 1, m, 1);
    map_tile_kernel<<<grid, 256, 0, st>>>(store, n_store, Cd, S, ids, poses, center, m, tile);
    DPM_CHECK_LAUNCH("map_tile", st);
    return DPM_OK;
}
