/*
 * dpm_b200.h -- C ABI of libdpm_b200.so: the B200 (sm_100a) implementation of
 * DeepPointMap's per-frame hot path (point-cloud encoder + registration decoder).
 *
 * Conventions (all entry points):
 *   - plain C: raw DEVICE pointers + sizes, no torch types.  `stream` is a cudaStream_t
 *     passed as void*; all work is enqueued on it, nothing synchronises internally
 *     unless the function is documented as "_host" (then buffers are HOST pointers and
 *     the call returns when the result is in host memory).
 *   - return 0 on success, a negative DPM_ERR_* otherwise; dpm_last_error() gives the
 *     thread-local message.  The callee never allocates or frees caller memory: scratch
 *     comes from the caller's `ws` buffer (size from the matching *_workspace_bytes).
 *   - tensors are dense row-major fp32 unless stated; indices int64 (reference dtype);
 *     masks are 1 byte per element (torch.bool).
 *   - re-entrant from several host threads as long as each call has its own workspace.
 *
 * Each declaration cites the reference interface it replaces, relative to
 * ZhangXiaze/DeepPointMap (/root/reference) -- or pytorch3d 0.7.4 ("[t3d]"), the
 * un-vendored dependency whose ops the reference calls at those lines.
 */
#ifndef DPM_B200_H
#define DPM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPM_OK 0
#define DPM_ERR_SHAPE (-1)       /* bad / inconsistent sizes                          */
#define DPM_ERR_UNSUPPORTED (-2) /* valid request this build does not implement        */
#define DPM_ERR_WORKSPACE (-3)   /* ws_bytes smaller than *_workspace_bytes()          */
#define DPM_ERR_CUDA (-4)        /* a CUDA runtime call failed                         */
#define DPM_ERR_ARG (-5)         /* null pointer / bad flag                            */

#define DPM_MAX_STAGES 8
#define DPM_MAX_BLOCKS 4 /* radius_list entries per stage (1 SA + up to 3 InvResMLP) */

typedef void *dpm_stream_t; /* cudaStream_t */

int dpm_version(void);
const char *dpm_last_error(void);
/* number of kernel launches enqueued by this host thread since the last reset
 * (bench.py's "gpu_launches") */
long long dpm_launch_count(void);
void dpm_launch_count_reset(void);
/* per-launch device-time profile of this host thread (bench.py's kernel breakdown and the
 * roofline.achieved figure).  dpm_prof_begin records a start event on `stream`; while open every
 * launch of this thread records an event after itself.  dpm_prof_end synchronises on them,
 * closes the profile and writes one line per launch -- "<kernel> <a> <b> <ms>\n", (a, b) =
 * kernel-specific sizes, e.g. (N, K) for fps and (S, N) for knn -- into buf; returns the number
 * of launches (>= 0) or a negative DPM_ERR_*. */
int dpm_prof_begin(dpm_stream_t stream);
int dpm_prof_end(char *buf, size_t buf_bytes);

/* ------------------------------------------------------------------------------------
 * index ops -- seam #3 (pytorch3d.ops) and seam #2 (Sampler / Querier)
 * ---------------------------------------------------------------------------------- */

/* [t3d] sample_farthest_points(points (B,N,D), lengths, K, random_start_point=False)
 * called from network/encoder/utils.py:278,282 (Sampler.fps_t3d); same result as the
 * reference's own Sampler.fps, utils.py:209-270.  Start index 0; d2 = (dx*dx+dy*dy)+dz*dz
 * in fp32 without FMA; argmax = first maximum; idx = -1 once k >= lengths[b].
 * idx_out (B,K) int64.  sampled_out (B,K,D) optional (masked_gather, utils.py:298-343:
 * rows with idx -1 are 0).  lengths (B) int64 on device, NULL = all N. */
int dpm_fps_f32(const float *points, int B, int N, int D, const int64_t *lengths, int K,
                int64_t *idx_out, float *sampled_out, void *ws, size_t ws_bytes, dpm_stream_t stream);
size_t dpm_fps_workspace_bytes(int B, int N, int D, int K);
/* How a cloud is mapped onto the chip by every FPS in this library (same picks either way): 0 = auto (a
 * cluster of 8 SMs per cloud while all clouds of the call fit the chip at once, B <= dpm_fps_cluster_capacity() -- the latency
 * shape of pipeline/infer.py's batch of 1; one SM per cloud for larger batches, which leaves the other SMs
 * to the kernels of concurrent streams), 1 = always one SM per cloud, 2 = always a cluster, 3 = "packed": TWO clouds
 * per SM in the one-SM kernel (clouds of <= 65 536 points) -- 22 % less SM time per batch at 1.55x the latency of its
 * FPS, for callers that keep >= 8 streams of batches in flight.  Process-wide. */
void dpm_set_fps_mode(int mode);
/* clouds per call up to which mode 0 takes the cluster mapping: the number of 8-CTA clusters of the largest FPS
 * kernel the current device can hold at once (cudaOccupancyMaxActiveClusters) */
int dpm_fps_cluster_capacity(void);

/* [t3d] knn_points(p1 (B,S,D1), p2 (B,N,D2), lengths1, lengths2, K) -> dists (B,S,K)
 * squared, ascending by (d2, index); idx (B,S,K) int64.  Only xyz (first 3 columns) is
 * used (utils.py:94,115 pass [..., :3]).  Slots k >= lengths2[b] and rows
 * s >= lengths1[b] are 0.  K <= 32.  d2_out may be NULL. */
int dpm_knn_f32(const float *p1, int D1, const float *p2, int D2, int B, int S, int N,
                const int64_t *lengths1, const int64_t *lengths2, int K, int64_t *idx_out,
                float *d2_out, void *ws, size_t ws_bytes, dpm_stream_t stream);

/* Querier.hybrid_query_t3d, network/encoder/utils.py:112-123: knn_points then every
 * slot with d2 > radius2 takes slot 0's index.  radius2 = fp32(radius**2). */
int dpm_knn_radius_f32(const float *p1, int D1, const float *p2, int D2, int B, int S, int N,
                       const int64_t *lengths2, int K, float radius2, int64_t *idx_out, void *ws,
                       size_t ws_bytes, dpm_stream_t stream);

/* [t3d] ball_query(p1, p2, lengths1, lengths2, K, radius) as used at utils.py:100-110:
 * the first K points (index order) with d2 < radius2; idx -1 / d2 0 padded. */
int dpm_ball_query_f32(const float *p1, int D1, const float *p2, int D2, int B, int S, int N,
                       const int64_t *lengths1, const int64_t *lengths2, int K, float radius2,
                       int64_t *idx_out, float *d2_out, void *ws, size_t ws_bytes,
                       dpm_stream_t stream);
size_t dpm_knn_workspace_bytes(int B, int S, int N, int K);

/* ------------------------------------------------------------------------------------
 * dense building blocks (row-major activations: one point / token per row)
 * ---------------------------------------------------------------------------------- */

#define DPM_ACT_NONE 0
#define DPM_ACT_RELU 1

/* Y (M,N) = act( X (M,K) . W (N,K)^T + bias (N) + res (M,N) ), leading dimensions in
 * elements.  Replaces the 1x1 Conv1d/Conv2d/Linear calls built by build_mlp
 * (network/encoder/utils.py:358-389) and the decoder's projections.  bias/res may be NULL. */
int dpm_linear_f32(const float *X, int ldx, const float *W, int ldw, const float *bias,
                   const float *res, int ldres, float *Y, int ldy, int M, int N, int K, int act,
                   dpm_stream_t stream);

/* Y (M,C) = act( LayerNorm_C(X) * gamma + beta + post ), eps 1e-5: LayerNorm1d/2d,
 * network/encoder/utils.py:392-413, and nn.LayerNorm in descriptor_attention.py:21-23.
 * post (M,C) optional (residual identity of InvResMLP, pointnext.py:136; positional
 * embedding of the next attention block, descriptor_attention.py:31,39). */
/* Same, with a caller workspace of dpm_linear_workspace_bytes(N, K) bytes: the weight matrix is split
 * once into hi / lo TF32 copies there (one extra launch) and the tensor-core GEMM reads those instead of
 * splitting W in every CTA -- what the encoder / decoder calls do internally for all their layers. */
size_t dpm_linear_workspace_bytes(int N, int K);
int dpm_linear_ws_f32(const float *X, int ldx, const float *W, int ldw, const float *bias,
                      const float *res, int ldres, float *Y, int ldy, int M, int N, int K, int act,
                      void *workspace, size_t ws_bytes, dpm_stream_t stream);

/* One build_mlp block in one launch (network/encoder/utils.py:358-389: conv -> LayerNorm -> ReLU), also the
 * decoder's "add & norm" (descriptor_attention.py:33-50):
 *   Y = act( LayerNorm_N( X W^T + bias + res ) * gamma + beta + post )
 * The LayerNorm rides in the GEMM epilogue when the row fits one column tile (N <= 256, N % 4 == 0, 16-byte
 * aligned operands); otherwise the two kernels run back to back through the workspace.  Y may alias res / post.
 * workspace: dpm_linear_ln_workspace_bytes(M, N, K). */
size_t dpm_linear_ln_workspace_bytes(int M, int N, int K);
int dpm_linear_ln_ws_f32(const float *X, int ldx, const float *W, int ldw, const float *bias,
                         const float *res, int ldres, const float *gamma, const float *beta,
                         const float *post, int ldpost, float *Y, int ldy, int M, int N, int K, int act,
                         void *workspace, size_t ws_bytes, dpm_stream_t stream);

int dpm_layernorm_f32(const float *X, int ldx, const float *gamma, const float *beta,
                      const float *post, int ldpost, float *Y, int ldy, int M, int C, int act,
                      dpm_stream_t stream);

/* Fused gather + [fea, (xyz-centre)/r] + 1x1 conv + LayerNorm + ReLU + max over K:
 * SetAbstraction.forward pointnext.py:52-61 / LocalAggregation.forward :97-107.
 * The 1x1 conv is split algebraically: Zfea (B,N,Cout) = fea . Wfea^T + bias is computed
 * per POINT beforehand (dpm_linear_f32), the geometric part Wxyz (Cout x 3, row stride
 * ldw) is applied per gathered neighbour here.  xyz4 (B,N,4), ctr4 (B,S,4) are float4
 * (x,y,z,0); gidx (B,S,K) int32; out (B,S,Cout). */
int dpm_group_ln_relu_max_f32(const float *Zfea, const float *xyz4, const float *ctr4,
                              const int32_t *gidx, const float *Wxyz, int ldw, const float *gamma,
                              const float *beta, float radius, float *out, int B, int N, int S,
                              int K, int Cout, dpm_stream_t stream);

/* FeaturePropagation.forward pointnext.py:188-211: 3-NN inverse-squared-distance
 * interpolation of fea2 (B,S,C2) onto xyz1 (B,N,.) and concat -> out (B,N,C1+C2). */
int dpm_fp_interp_f32(const float *xyz1_4, const float *xyz2_4, const float *fea1, const float *fea2,
                      const uint8_t *pad2, float *out, int B, int N, int S, int C1, int C2,
                      dpm_stream_t stream);

/* ------------------------------------------------------------------------------------
 * whole-path entry points -- seam #1 (network.encoder.Encoder / network.decoder.Decoder)
 * ---------------------------------------------------------------------------------- */

typedef struct dpm_encoder_desc {
    int n_stages;                           /* len(encoder.npoint)                       */
    int in_channel, width, expansion;       /* encoder.in_channel / width / expansion    */
    int out_channel, upsample_layers;       /* encoder.out_channel / upsample_layers     */
    int npoint[DPM_MAX_STAGES];             /* encoder.npoint                            */
    int n_blocks[DPM_MAX_STAGES];           /* len(radius_list[i])                       */
    double radius[DPM_MAX_STAGES][DPM_MAX_BLOCKS]; /* encoder.radius_list (Python doubles) */
    int nsample[DPM_MAX_STAGES][DPM_MAX_BLOCKS];
} dpm_encoder_desc;

/* Encoder.forward, network/encoder/encoder.py:51-69 (+ the descriptor glue of
 * ExtractionThread.process, system/modules/odometry.py:46-49 when desc_out != NULL).
 *   points (B,C,N) channel-first, C >= 3; padding (B,N) bool or NULL (= all valid)
 *   weights: device pointers of the encoder state_dict tensors, in state_dict order
 *            (point_mlp0.weight, point_mlp0.bias, downsampler.0.sa.mlp.0.weight, ...)
 *   out_coor (B,3,S), out_fea (B,out_channel,S), out_pad (B,S) bool, S = npoint of the
 *   level the FPN ends on; desc_out (B,out_channel+3,S) = [fea ; coor*coor_scale] or NULL.
 *   trace_fps / trace_knn: optional device buffers receiving every stage's FPS indices
 *   (int64, concatenated (B,npoint[i])) and every query's group indices (int32,
 *   concatenated (B,S,K)) for parity tests; NULL in production. */
int dpm_encoder_forward(const dpm_encoder_desc *desc, const float *const *weights, int n_weights,
                        const float *points, int C, const uint8_t *padding, int B, int N,
                        float *out_coor, float *out_fea, uint8_t *out_pad, float *desc_out,
                        float coor_scale, int64_t *trace_fps, int32_t *trace_knn, void *ws,
                        size_t ws_bytes, dpm_stream_t stream);
size_t dpm_encoder_workspace_bytes(const dpm_encoder_desc *desc, int B, int N);
int dpm_encoder_num_weights(const dpm_encoder_desc *desc);
int dpm_encoder_out_points(const dpm_encoder_desc *desc);

/* Weight-copy reuse across calls (thread-local, default 0 = off): every encoder / decoder / linear_ws call writes
 * hi / lo tf32 copies of its GEMM weights into the head of the caller's workspace.  With a NON-ZERO epoch the calling
 * thread promises that neither the weight tensors nor that workspace changed since its previous call under the same
 * epoch; the copy launch is then skipped.  Change the epoch (or set 0) whenever a weight or the workspace buffer does. */
void dpm_set_weights_epoch(unsigned long long epoch);

typedef struct dpm_decoder_desc {
    int in_channel, model_channel, attention_layers, heads; /* decoder.* ; heads = 8   */
    float tau, eps_offset;                                   /* loss.tau / eps_offset   */
} dpm_decoder_desc;

/* result record of one registration (device or host memory, 64 floats) */
#define DPM_REG_R 0        /* 9 floats, row-major R                                    */
#define DPM_REG_T 9        /* 3 floats                                                 */
#define DPM_REG_RMSE 12    /* inlier rmse                                              */
#define DPM_REG_NCORR 13   /* K' = correspondences kept by the offset filter (as float) */
#define DPM_REG_NINLIER 14 /* K'' = inliers after the SVD loop                          */
#define DPM_REG_ITERS 15
#define DPM_REG_STRIDE 16

/* Decoder.registration_forward, network/decoder/decoder.py:91-127, for P independent
 * (src, dst) pairs (the reference asserts P == 1 per call; P > 1 is the batched form).
 *   src (P,Cd,M), dst (P,Cd,N) channel-first unified descriptors, Cd = in_channel+3,
 *   xyz rows in metres.  k = num_pairs sampled per pair (decoder.py:170-178, computed by
 *   the caller).  weights: decoder state_dict tensors in state_dict order.
 *   result (P,DPM_REG_STRIDE) floats; conf_out (P,2k) = pairing confidence of the kept
 *   correspondences in reference order with the inlier ones FIRST COMPACTED:
 *   conf_out[p][0..K''-1] is what the reference returns. */
/* src_pad (P,M) / dst_pad (P,N): the reference's key-padding masks (1 = padded descriptor, ignored as an attention
 * KEY, descriptor_attention.py:33-42; like the reference the padded rows still take part in everything else), or
 * NULL (shipped inference always passes None). */
int dpm_registration_forward(const dpm_decoder_desc *desc, const float *const *weights,
                             int n_weights, const float *src, const float *dst,
                             const uint8_t *src_pad, const uint8_t *dst_pad, int P, int M,
                             int N, int k, float *result, float *conf_out, void *ws,
                             size_t ws_bytes, dpm_stream_t stream);
size_t dpm_registration_workspace_bytes(const dpm_decoder_desc *desc, int P, int M, int N, int k);

/* Decoder.loop_detection_forward, decoder.py:129-143 + OverlapHead heads.py:45-69:
 * src, dst (P,Cd,L) -> prob (P). */
int dpm_loop_detection_forward(const dpm_decoder_desc *desc, const float *const *weights,
                               int n_weights, const float *src, const float *dst,
                               const uint8_t *src_pad, const uint8_t *dst_pad, int P, int M,
                               int N, float *prob, void *ws, size_t ws_bytes, dpm_stream_t stream);
size_t dpm_loop_detection_workspace_bytes(const dpm_decoder_desc *desc, int P, int M, int N);
/* number of entries of `weights` for the decoder calls: the 82 state_dict tensors in
 * state_dict order (projection, descriptor_attention.{l}.{self_attn,cross_attn}.{in_proj_weight,
 * in_proj_bias,out_proj.weight,out_proj.bias}, .mlp.{0,2}, .norm{1,2,3}, similarity_head,
 * offset_head.{mlp.0,mlp.2,mlp.4,downsample,head}, loop_head.{mlp.0,mlp.2,projection.0,
 * projection.2}, coarse_pairing_head) PLUS one trailing entry: dim_t, the
 * model_channel/3/2*2 positional-embedding frequencies temperature**(2*(i//2)/npf)
 * (descriptor_attention.py:70-71), computed once by the host. */
int dpm_decoder_num_weights(const dpm_decoder_desc *desc);

/* decoder pieces exposed for parity tests */
/* PositionEmbeddingCoordsSine.forward descriptor_attention.py:66-83: xyz (R,3) metres ->
 * emb (R,C).  dim_t (npf) = temperature**(2*(i//2)/npf), precomputed by the caller. */
int dpm_posenc_f32(const float *xyz, int ldx, const float *dim_t, int npf, float *emb, int R, int C,
                   dpm_stream_t stream);
/* multi-head attention core: out (Lq, H*32) = softmax(q k^T / sqrt(32)) v per head, for
 * `nprob` (q-range, kv-range) problems over row-major qkv buffers (nn.MultiheadAttention
 * core, descriptor_attention.py:33-42).  head_dim must be 32. */
int dpm_attention_f32(const float *q, int ldq, const float *k, int ldk, const float *v, int ldv,
                      float *out, int ldo, const int *prob /* device, nprob x 4: q0,Lq,k0,Lk */,
                      int nprob, int max_lq, int heads, dpm_stream_t stream);
/* The attention core of DescriptorAttentionLayer (network/decoder/descriptor_attention.py:33-42, head_dim 32) in the
 * decoder's token layout: P pairs, rows p*(M+N) .. p*(M+N)+M-1 are the src tokens, the next N the dst tokens; q / k / v
 * (rows x heads*32, leading dimensions ldq / ldk / ldv) already projected.  mode 0: self attention of each side,
 * 1: cross attention (src queries over dst keys and vice versa).  kmask: key-padding mask per token row or NULL.
 * impl 0: as the decoder chooses (mma.sync for <= 512 keys, tcgen05 flash attention above), 1 / 2 force one. */
int dpm_attention_pairs_f32(const float *q, int ldq, const float *k, int ldk, const float *v, int ldv, float *out,
                            int ldo, int P, int M, int N, int mode, int heads, const uint8_t *kmask, int impl,
                            dpm_stream_t stream);

/* weighted Kabsch + 3-sigma loop, decoder.py:227-265, for P problems of Kc <= ldk
 * correspondences each: src/dst (P,3,ldk), w (P,ldk), count (P) int32 -> result
 * (P,DPM_REG_STRIDE), inlier mask (P,ldk) bytes (optional), conf_out (P,ldk) = the inlier
 * weights compacted in order (optional).  ldk <= 4096. */
int dpm_kabsch_f32(const float *src, const float *dst, const float *w, const int32_t *count, int P,
                   int ldk, float *result, uint8_t *inlier, float *conf_out, dpm_stream_t stream);

/* calculate_information_matrix_from_pcd, system/modules/utils.py:60-104 (pytorch3d branch): src (3,N1),
 * dst (3,N2) channel-first fp32, SE3 16 floats row-major (device memory); every transformed source point
 * R.p+T is matched to its nearest dst point when d2 <= radius^2; info (36 floats, device) = sum over the
 * matched dst points of G^T G, G = d(R(w) t + v)/d(w, v).  n_corr (device int32, optional) = matches.
 * No host sync; SURVEY.md section 8f rank 2. */
size_t dpm_information_matrix_workspace_bytes(int N1, int N2);
int dpm_information_matrix_f32(const float *src, int N1, const float *dst, int N2, const float *SE3,
                               float radius, float *info, int32_t *n_corr, void *workspace,
                               size_t ws_bytes, dpm_stream_t stream);

/* Raw frame -> encoder input on the device: BinReader's NaN-row drop
 * (dataloader/heads/bin.py:16-17), VoxelSample(voxel_size, 'first') (dataloader/transforms.py:331-356),
 * DistanceSample(min_dis, max_dis) (:387-397) and CoordinatesNormalization(ratio) (:400-407).
 * raw: N rows of `stride` floats (x, y, z first; 4 for a KITTI .bin).  out_rows: room for N x 3 floats; the
 * first *count rows are the surviving points / ratio in ascending voxel-id order (the reference's order).
 * count (device int32) = -1 when the voxel grid of this frame has more than max_voxels cells (nothing is
 * written then; crop the frame or raise max_voxels).  No host sync. */
size_t dpm_frontend_workspace_bytes(long long max_voxels);
int dpm_frontend_f32(const float *raw, int N, int stride, float voxel_size, float min_dis, float max_dis,
                     float ratio, long long max_voxels, float *out_rows, int32_t *count, void *workspace,
                     size_t ws_bytes, dpm_stream_t stream);

/* Scan-to-map input stage: PoseGraph.__global_mapping + the centring of
 * global_map_query_graph, system/modules/pose_graph.py:373-409, 499-511.  store (n_store, Cd, S): device-resident
 * descriptor sets [fea ; xyz]; ids (m) int32 slots of the key-frames; poses (m,16) their SE3_pred, row-major;
 * center (16) the centring SE3 or NULL (identity).  tile (Cd, m*S):
 *   tile[:, i*S:(i+1)*S] = [ fea_i ; Rc^T ((R_i xyz_i + t_i) - tc) ]   -- the `dst` of dpm_registration_forward. */
int dpm_map_tile_f32(const float *store, int n_store, int Cd, int S, const int32_t *ids, const float *poses,
                     const float *center, int m, float *tile, dpm_stream_t stream);

/* OutlierFilter, the reference's CUDA branch (dataloader/transforms.py:230-246): rows (N x stride floats, xyz
 * first) -> the rows whose mean distance to their nb_neighbors nearest neighbours is <= mean + std_ratio * std
 * of that statistic over the cloud, in their original order, divided by out_divisor (1 = as they are; the
 * CoordinatesNormalization ratio when the filter is the last step before it).  out_rows: room for N x 3 floats;
 * mask (N bytes, optional) the keep flags; count (device int32) the number of survivors.  nb_neighbors <= 31.
 * No host sync. */
size_t dpm_outlier_filter_workspace_bytes(int N, int nb_neighbors);
int dpm_outlier_filter_f32(const float *rows, int N, int stride, int nb_neighbors, float std_ratio,
                           float out_divisor, float *out_rows, uint8_t *mask, int32_t *count, void *workspace,
                           size_t ws_bytes, dpm_stream_t stream);

/* LowPassFilter (dataloader/transforms.py:256-297): rows (N x stride floats, xyz first, metres) -> the rows whose
 * normal agrees with its neighbours': sim_i = sum of the `flux` largest |n_i . n_j| over the normals_num nearest
 * neighbours j (n = PCA normal of the points within normals_radius, open3d's estimate_normals), kept when
 * sim_i > mean(sim) - filter_std * std(sim); original order, divided by out_divisor.  out_rows: room for N x 3
 * floats; mask (N bytes) and sim_out (N floats) optional; count (device int32) the number of survivors.
 * normals_num <= 31, flux <= 8.  The reference's max_remain re-ranking is left to the caller (sim_out).  No host sync. */
size_t dpm_low_pass_filter_workspace_bytes(int N, int normals_num);
int dpm_low_pass_filter_f32(const float *rows, int N, int stride, float normals_radius, int normals_num,
                            float filter_std, int flux, float out_divisor, float *out_rows, uint8_t *mask,
                            float *sim_out, int32_t *count, void *workspace, size_t ws_bytes, dpm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DPM_B200_H */
