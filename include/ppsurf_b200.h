/* ppsurf_b200 -- C ABI of the B200-native PPSurf occupancy hot path.
 *
 * The reference (cg-tuwien/ppsurf) is pure Python and has no FFI of its own, so every entry point
 * below replaces a Python call site of the reference; the call site is cited as  file:line  relative to the
 * reference root.  The Python side of this repo (ppsurf_b200/) binds these symbols with ctypes; INTEGRATION.md
 * shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless the name ends in _host
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); every call is asynchronous on it
 *     unless documented otherwise, never allocates device memory, and works only inside caller-owned buffers
 *   - return value: 0 on success, negative pps_status on error; pps_last_error() returns a message for the
 *     calling thread
 *   - layouts are POINT-MAJOR (row = one point / query, channels contiguous): pts [N,3] f32, latents [N,C] f32,
 *     indices int32.  The reference's [B,C,N] tensors are transposed once at the Python boundary.
 */
#ifndef PPSURF_B200_H
#define PPSURF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum pps_status {
    PPS_OK = 0,
    PPS_ERR_INVALID = -1,   /* bad argument (null pointer, size out of range, k > supported ...) */
    PPS_ERR_WORKSPACE = -2, /* caller-provided workspace too small                                 */
    PPS_ERR_CUDA = -3,      /* a CUDA runtime call or kernel launch failed                         */
    PPS_ERR_NO_DEVICE = -4  /* no sm_100 device: there is NO CPU fallback                          */
} pps_status;

const char* pps_last_error(void);
/* library version and the SM architecture the kernels were compiled for (100 for sm_100a) */
int pps_version(void);
int pps_compiled_arch(void);
/* 64-bit digest of the CUDA sources and this header the library was built from (ppsurf_b200/build.py computes it, the
 * Python binding recomputes it from the sources next to the library): a stale prebuilt binary cannot pass for the
 * current sources */
unsigned long long pps_source_hash(void);
/* 0 if the current device can run the kernels, PPS_ERR_NO_DEVICE otherwise */
int pps_check_device(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
unsigned long long pps_launch_count(void);
/* Measurement hooks: while enabled, the decoder brackets its dominant kernels (the fc2/fc3/fc_query GEMMs of the
 * global branch) with CUDA events on the launching stream; pps_profile_read() synchronises, returns the summed
 * device time and the number of brackets since the last read, and resets. */
void pps_profile_enable(int on);
int pps_profile_read(double* total_ms, long long* brackets);

/* ---------------------------------------------------------------------------------------------------------------
 * a6  exact k nearest neighbours
 *     replaces  knn()                   source/poco_utils.py:257-273
 *               make_kdtree/query_kdtree source/base/proximity.py:40-81  (pykdtree, CPU, rebuilt per call)
 * Index = points sorted by 3-D Morton code + start offsets of the finest cells (an implicit octree).
 * Distances are float32 ((dx*dx + dy*dy) + dz*dz, no FMA) like pykdtree's float path; results ascend by
 * (dist2, index); k must be <= n (the reference clamps k to n before the query, poco_utils.py:259-260).
 * ------------------------------------------------------------------------------------------------------------- */
size_t pps_knn_index_bytes(int64_t n);
int pps_knn_build(const float* pts, int64_t n, void* index, size_t index_bytes, void* stream);
/* idx_out [q,k] int32 (original point numbering), dist2_out [q,k] f32 or NULL */
/* tuning knob: consecutive queries one warp handles in the seeded search (33 <= k <= 256); returns the previous value */
int pps_debug_knn_run(int run);
/* tuning knob: finest octree cells per point of indices built from now on (query an index under the setting it was built with) */
int pps_debug_knn_cells(int factor);
/* tuning knob: an inner octree node of at most this many points per unpruned child is scanned as one contiguous range instead of
 * being traversed (results do not depend on it); returns the previous value */
int pps_debug_knn_scan_child(int points);
/* tuning knob: points a run of consecutive queries may scan before its remaining queries are deferred to a second pass of one warp per
 * query (results do not depend on it); returns the previous value, points < 1 only reads */
int pps_debug_knn_scan_cap(int points);
int pps_knn_query(const void* index, int64_t n, const float* queries, int64_t q, int k, int32_t* idx_out,
                  float* dist2_out, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * a7  local patches in patch space
 *     replaces  _get_pts_local_ps                    source/poco_utils.py:67-72
 *               PPSurfDataset.normalize_patches      source/ppsurf_data_loader.py:91-123
 * idx [q,k_stride] are the (ascending) neighbours from pps_knn_query; the first p of each row form the patch;
 * radius = sqrt(dist2[q,p-1]).  out [q,p,3] f32.
 * ------------------------------------------------------------------------------------------------------------- */
int pps_patch_normalize(const float* pts, const float* queries, const int32_t* idx, const float* dist2,
                        int64_t q, int p, int k_stride, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * generic fused pointwise layer  Y = act( X[rows or gathered rows] . W^T + bias + residual )
 *     replaces every 1x1 Conv1d/Conv2d/Linear (+ folded eval BatchNorm + ReLU) of
 *     ResidualBlock / FKAConvNetwork decoder / STN / PointNetfeat / MLP
 *     source/base/nn.py:162-190,305-373,415-417,438-450,530-548
 * X [m or n_src, k] f32, W [n,k] f32 (row = output channel), Y [m,n].  gather (nullable) picks row gather[i] of X
 * for output row i (the 1-NN `interpolate`, nn.py:684-697); residual (nullable, may alias Y) is added before the
 * activation; act: 0 none, 1 ReLU.
 * ------------------------------------------------------------------------------------------------------------- */
int pps_linear(const float* x, const float* w, const float* bias, const float* residual, const int32_t* gather,
               float* y, int64_t m, int n, int k, int ldx, int ldy, int act, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * a8-a11  occupancy decoder (global attention-interpolation branch + local PointNet branch + MLP + softmax)
 *     replaces  PPSurfNetwork.from_latent             source/ppsurf_model.py:82-117
 *               InterpAttentionKHeadsNet.forward      source/poco_model.py:381-419
 *               PointNetfeat.forward / STN / AttentionPoco   source/base/nn.py:305-373,162-190,84-96
 *               MLP.forward                           source/base/nn.py:415-417
 *               _predict_from_latent                  source/poco_utils.py:74-82
 * Weights are packed on the host by ppsurf_b200.packing (eval BatchNorm folded, algebraic merges listed in
 * DESIGN.md); all matrices are [out,in] row-major f32.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct pps_decoder_weights {
    int32_t latent;        /* C = 256 */
    int32_t heads;         /* 64 (fc_query outputs) */
    int32_t k;             /* 64 neighbours of the global branch */
    int32_t num_pts_local; /* P = 50 / 200 */
    /* global branch */
    const float* w1_lat; /* [C,C]   fc1 weight, latent columns                     */
    const float* w1_xyz; /* [C,3]   fc1 weight, xyz columns                        */
    const float* b1;     /* [C]                                                    */
    const float* w2;     /* [C,C] */
    const float* b2;
    const float* w3;     /* [C,C] */
    const float* b3;
    const float* wq;     /* [heads,C] */
    const float* bq;
    const float* wv8;    /* [C,C]   fc8 . fc_value                                 */
    const float* bv8;    /* [C]     fc8 . b_value + b8                             */
    /* local branch (BatchNorm folded) */
    const float* pn0a_w; /* [64,3]  */
    const float* pn0a_b;
    const float* pn0b_w; /* [64,64] */
    const float* pn0b_b;
    const float* stn1_w; /* [64,64] */
    const float* stn1_b;
    const float* stn2_w; /* [128,64] */
    const float* stn2_b;
    const float* stn3_w; /* [S,128]  S = pointnet_latent_size */
    const float* stn3_b;
    const float* stnf1_w; /* [S/2,S] */
    const float* stnf1_b;
    const float* stnf2_w; /* [S/4,S/2] */
    const float* stnf2_b;
    const float* stnf3_w; /* [4096,S/4] */
    const float* stnf3_b; /* [4096]  fc3 bias + identity                            */
    const float* pn1_w;   /* [64,64] */
    const float* pn1_b;
    const float* pn2_w;   /* [128,64] */
    const float* pn2_b;
    const float* pnq_w;   /* [128]   att.fc_query . bn3 . conv3                     */
    float pnq_b;
    int32_t stn_size;     /* S */
    const float* pnv_w;   /* [C,128] att.fc_value . bn3 . conv3                     */
    const float* pnv_b;   /* [C]                                                    */
    /* MLP (BatchNorm folded) */
    const float* m0_w; /* [C,C] */
    const float* m0_b;
    const float* m1_w; /* [C,C] */
    const float* m1_b;
    const float* m2_w; /* [2,C] */
    const float* m2_b; /* [2]   */
    /* tensor-core pack of fc2 / fc3 / fc_query for path 1 (pps_decoder_tc_pack_bytes() bytes, nullable): per layer 16
     * k16 stages, each [W_hi k8-block 0 | W_hi k8-block 1 | W_lo block 0 | W_lo block 1], a block = N rows x 8 fp16 */
    const void* tc_wpack;
    /* same stage format for the local branch (P <= 256): [conv0b | stn.conv1 | stn.conv2 | stn.conv3 rows 0-127 | rows
     * 128-255] and [conv1 | conv2]; pps_decoder_tc_pn_stn_bytes() / pps_decoder_tc_pn_feat_bytes() bytes, nullable */
    const void* tc_pn_stn;
    const void* tc_pn_feat;
    /* per-query chains: [stn.fc1 | stn.fc2 | stn.fc3 in 16 blocks of 256 rows] and [(W8 Wv) | (Wv_att A3) | mlp.0 | mlp.1];
     * tc_bias_feat [C] = bv8 + pnv_b (bias of the summed branches); nullable */
    const void* tc_stn_fc;
    const void* tc_mlp;
    const float* tc_bias_feat;
} pps_decoder_weights;

size_t pps_decoder_tc_pack_bytes(void);
/* debug aid: while `counters` (128 x int64, device memory) is set, the projection runs as its instrumented instance: CTA 0 writes
 * per-phase cycle counters [0..9], every CTA pair its total [32 + pair] (tools/tc_phase_profile.py); NULL switches it off */
void pps_debug_tc_profile(long long* counters);
/* retired tuning knob (no-op, kept for ABI stability) */
void pps_debug_tc_cluster(int cs);
/* Split-fp16 terms of the tensor-core global branch (pass ablation, profiles/r02_pass_ablation.md): bit 3*layer + t with layer
 * 0 = fc2, 1 = fc3, 2 = fc_query and t = 0: x_hi*w_hi (always on), 1: x_lo*w_hi, 2: x_hi*w_lo.  Default 0x0FF (fc_query without its
 * weight-lo term: 1e-6 on the logits, 7 % of the kernel); 0x1FF = three terms everywhere.  mask < 0 only
 * reads.  Returns the previous mask. */
int pps_decoder_tc_terms(int mask);
/* debug: number of CTA pairs (2-CTA clusters) of the projection kernel the device holds at once */
int pps_debug_tc_max_clusters(void);
size_t pps_decoder_tc_pn_stn_bytes(void);
size_t pps_decoder_tc_pn_feat_bytes(void);
size_t pps_decoder_tc_stn_fc_bytes(void);
size_t pps_decoder_tc_mlp_bytes(void);

/* per-point table  U[n,:] = W1_lat . latent[n] - W1_xyz . pts[n] + b1   (fc1 hoisted out of the (query,neighbour)
 * loop: fc1([latent_j, q - p_j]) = U_j + W1_xyz . q).  table [n,C] f32. */
int pps_decoder_point_table(const pps_decoder_weights* w, const float* pts, const float* latents, int64_t n,
                            float* table, void* stream);

/* workspace bytes for decoding chunks of at most `chunk` queries */
size_t pps_decoder_workspace_bytes(const pps_decoder_weights* w, int64_t chunk);

/* Decode q queries: kNN (k = max(w->k, P)) -> both branches -> MLP.
 *   logits_out [q,2] f32 or NULL;  occ_out [q] f32 (softmax(l)[0] - softmax(l)[1]) or NULL.
 *   idx_out [q,kmax] int32 or NULL (the neighbour ids, = reference proj_ids for the first w->k columns).
 *   path: 0 = fp32 SIMT kernels, 1 = tcgen05 split-fp16 tensor-core kernels (both branches; the local branch needs
 *   P <= 256 -- a patch then spans ceil(P/64) half-tiles -- and otherwise stays on the fp32 kernels). */
int pps_decoder_decode(const pps_decoder_weights* w, const void* knn_index, const float* pts, const float* table,
                       int64_t n, const float* queries, int64_t q, int64_t chunk, void* workspace,
                       size_t workspace_bytes, float* logits_out, float* occ_out, int32_t* idx_out, int path,
                       void* stream);

/* Same from HOST memory (the reference hands `pts_query` over as a CPU tensor, source/poco_utils.py:220 and
 * source/poco_model.py:392, and reads the result back with .cpu(), poco_utils.py:81): queries_host [q,3] and
 * occ_host [q] should be pinned; uploads/downloads are chunked on `copy_stream` and overlap the kernels.
 * Synchronous: returns when occ_host is complete.  staging: device buffer of q*16 bytes. */
int pps_decoder_decode_host(const pps_decoder_weights* w, const void* knn_index, const float* pts,
                            const float* table, int64_t n, const float* queries_host, int64_t q, int64_t chunk,
                            void* workspace, size_t workspace_bytes, void* staging, size_t staging_bytes,
                            float* occ_host, int path, void* stream, void* copy_stream);

/* stand-alone branches for the parity tests (same kernels as pps_decoder_decode) */
int pps_decoder_projection(const pps_decoder_weights* w, const float* pts, const float* table, const float* queries,
                           const int32_t* idx, int k_stride, int64_t q, void* workspace, size_t workspace_bytes,
                           float* feat_out /* [q,C] */, int path, void* stream);
int pps_decoder_pointnet(const pps_decoder_weights* w, const float* patches /* [q,P,3] */, int64_t q,
                         void* workspace, size_t workspace_bytes, float* feat_out /* [q,C] */, int path, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * a11  dense marching-cubes query grid
 *     replaces  the coordinate expression of _create_volume   source/poco_utils.py:212-213
 * out [(r^3),3] f32, C order, coordinate = idx * step + bmin_pad with separate multiply and add roundings.
 * Only vertices [first, first+count) of the flattened grid are produced (multi-GPU slabs).
 * ------------------------------------------------------------------------------------------------------------- */
int pps_grid_queries(int r, float step, float bmin_pad, int64_t first, int64_t count, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * a11 / f1  region-growing bookkeeping of the occupancy volume, on the device
 *     replaces  _create_volume's masks   source/poco_utils.py:178-254
 *               (_dilate_binary 181-196, np.argwhere 210-213, volume[mask] = z 229, sign-change frontier 232-244)
 * Voxels are C-order linear indices into the r^3 volume (r = resolution + 2*padding); every list comes out in ascending
 * index order (= np.argwhere order).  volume f32 [r^3] (NaN = not decoded), to_see u8 [r^3].
 *   pps_region_init      volume = NaN, to_see = 1
 *   pps_region_pending   voxels within `dilation` of a seed (clipped box) that are not decoded yet -> out_ids, *out_count
 *                        (device scalars; a voxel decoded in an earlier sweep keeps its value: the decode is deterministic)
 *   pps_region_queries   ids -> coordinates idx * step + bmin_pad (separate fp32 multiply and add)
 *   pps_region_scatter   volume[ids] = values
 *   pps_region_frontier  to_see[seeds] = 0; the new seeds are the to_see voxels with value >= 0 within `dilation` of a seed
 *                        with value <= 0, or with value <= 0 within `dilation` of a seed with value >= 0
 *   pps_region_finish    the `padding` outer layers = out_value   (poco_utils.py:246-251)
 * out_ids must hold r^3 entries; workspace: pps_region_workspace_bytes(r).
 * ------------------------------------------------------------------------------------------------------------- */
size_t pps_region_workspace_bytes(int r);
int pps_region_init(int r, float* volume, uint8_t* to_see, void* stream);
int pps_region_pending(const int32_t* seeds, int64_t n_seeds, int r, int dilation, const float* volume, void* workspace,
                       size_t workspace_bytes, int32_t* out_ids, long long* out_count, void* stream);
int pps_region_queries(const int32_t* ids, int64_t n, int r, float step, float bmin_pad, float* out, void* stream);
int pps_region_scatter(const int32_t* ids, const float* values, int64_t n, float* volume, void* stream);
int pps_region_frontier(const int32_t* seeds, int64_t n_seeds, int r, int dilation, const float* volume, uint8_t* to_see,
                        void* workspace, size_t workspace_bytes, int32_t* out_ids, long long* out_count, void* stream);
int pps_region_finish(float* volume, int r, int padding, float out_value, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * a2  quantised support sampling and the encoder's index tensors
 *     replaces  sampling_quantized   source/poco_data_loader.py:59-134   (torch_geometric voxel_grid + host loop)
 *               get_fkaconv_ids      source/poco_data_loader.py:137-209  (4 samplings at ratio 1/4, 13 kNN tensors)
 * pps_sample_quantized picks n_support of the n points: one representative per voxel of a randomly rotated grid, voxel
 * edge ||bbox||_2 / sqrt(n_support), halved until enough points are found, random trim of the last round.  rotations
 * [n_rot,9] are row-major 3x3 matrices in DEVICE memory (the caller draws the random angles), one per round (n_rot >= 6
 * recommended).  No host synchronisation.  sel_out [n_support] int32.  Random by construction in the reference too:
 * parity is distributional.
 * pps_encoder_ids runs the whole get_fkaconv_ids for a batch of clouds pts [b,n0,3]: level sizes n_{l+1} =
 * max(1, n_l / 4); rotations [b,4,n_rot,9]; outputs (caller-allocated, device): support[l-1] [b,n_l,3];
 * ids16[p] [b,n_c,min(16,n_a)] for (a,c) = (0,0),(0,1),(1,1),(1,2),(2,2),(2,3),(3,3),(3,4),(4,4);
 * ids1[p] [b,n_c,1] for (a,c) = (4,3),(3,2),(2,1),(1,0).
 * ------------------------------------------------------------------------------------------------------------- */
size_t pps_sample_workspace_bytes(int64_t n);
int pps_sample_quantized(const float* pts, int64_t n, int64_t n_support, const float* rotations, int n_rot, uint32_t seed,
                         void* workspace, size_t workspace_bytes, int32_t* sel_out, void* stream);
typedef struct pps_encoder_ids_out {
    float* support[4];
    int32_t* ids16[9];
    int32_t* ids1[4];
} pps_encoder_ids_out;
size_t pps_encoder_ids_workspace_bytes(int64_t n0);
int pps_encoder_ids(const float* pts, int64_t b, int64_t n0, const float* rotations, int n_rot, uint32_t seed, void* workspace,
                    size_t workspace_bytes, const pps_encoder_ids_out* out, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * a3  FKAConv layer (point-major activations)
 *     replaces  FKAConvLayer.forward   source/base/nn.py:592-652   (eval mode; InstanceNorm statistics are
 *     per sample over (Ns,16), so the layer runs as stats1 -> stats2 -> fused gather/contract)
 * x [b,n_in,cin], pts [b,n_in,3], support [b,n_s,3], ids [b,n_s,kn] int32 with 1 <= kn <= 16 (the reference asks for
 * 16 neighbours and clamps to n_in; kn == 1 skips both InstanceNorms like nn.py:627-636), out [b,n_s,cout].
 * cv_w is [cout, 16*cin] with column = m*cin + c  (repacked from the reference [cout,cin,1,16]); the eval
 * BatchNorm that always follows the layer (nn.py:441,519) is folded into cv_w / out_bias, out_relu applies its ReLU.
 * act: 0 ReLU (POCO), 1 SiLU (PPSurf).
 * Two statistics launches (InstanceNorm 1 and 2 are global per sample), then ONE fused kernel: index gather ->
 * kernel-weight MLP -> neighbourhood weighted sum -> tcgen05 contraction -> bias / ReLU (csrc/fka_tc.cu; needs kn == 16,
 * cin % 4 == 0, n_s >= 16, tc_pack).  Other shapes run the unfused fp32 kernels (third MLP launch, gather * weights kernel,
 * fp32 contraction) through the same entry point.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct pps_fkaconv_weights {
    int32_t cin, cout, act;
    float alpha, beta, norm_radius;
    const float* fc1;   /* [16,3]  */
    const float* fc2;   /* [16,32] */
    const float* fc3;   /* [16,32] */
    const float* in1_w; /* [16] InstanceNorm affine */
    const float* in1_b;
    const float* in2_w;
    const float* in2_b;
    const float* cv_w;     /* [cout,16*cin] */
    const float* out_bias; /* [cout], nullable */
    int32_t out_relu;
    /* operand pack of the fused tensor-core kernel (csrc/fka_tc.cu), nullable: cv_w as fp16 hi/lo k16 stages in the
     * shared-memory operand layout, K order k = ((c/2)*4 + m/4)*8 + (m%4)*2 + c%2, per slice of min(cout,256) rows;
     * 4*16*cin*cout bytes.  NULL selects the unfused fp32 kernels. */
    const void* tc_pack;
    /* HOST pointer (the only one in this struct), nullable: fc1 | fc2 | fc3 contiguous, 48 + 512 + 512 floats.  The fused kernels
     * take the kernel-weight MLP in their parameter block (constant-bank operands); NULL selects the unfused kernels. */
    const float* mlp_host;
    /* tc_pack holds cv_w / tc_out_scale (a power of two chosen so that the fp16 lo parts stay normal); the fused kernel multiplies
     * its accumulator by tc_out_scale */
    float tc_out_scale;
} pps_fkaconv_weights;

size_t pps_fkaconv_workspace_bytes(int64_t b, int64_t n_s, int cin); /* upper bound for any layer of that input width */
/* exact requirement of one layer call (the fused kernel only needs the statistics: 512 B per sample) */
size_t pps_fkaconv_workspace_bytes_for(const pps_fkaconv_weights* w, int kn, int64_t b, int64_t n_s);
/* debug / parity: 0 forces the unfused fp32 kernels even when the fused kernel supports the shape (default 1) */
void pps_debug_fka_fused(int on);
int pps_fkaconv_forward(const pps_fkaconv_weights* w, const float* x, const float* pts, const float* support,
                        const int32_t* ids, int kn, int64_t b, int64_t n_in, int64_t n_s, void* workspace,
                        size_t workspace_bytes, float* out, void* stream);

/* gather-max over the 16 neighbours (shortcut branch of a strided ResidualBlock)
 *     replaces  max_pool   source/base/nn.py:677-680
 * x [b,n_in,c], ids [b,n_s,16] -> out [b,n_s,c] */
int pps_gather_max(const float* x, const int32_t* ids, int64_t b, int64_t n_in, int64_t n_s, int c, int kn,
                   float* out, void* stream);
/* max over all points of a sample: x [b,n,c] -> out [b,c]   (nn.py:531) */
int pps_global_max(const float* x, int64_t b, int64_t n, int c, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * a1  latent accumulation of one encoder pass
 *     replaces  latent[ids] += partial; counts[ids] += 1   source/poco_model.py:228-229
 *     (torch advanced-index semantics: a duplicated id receives ONE of the duplicates' values, once)
 * ids [n] int32 must already be de-duplicated by the caller when exact reference semantics are needed.
 * ------------------------------------------------------------------------------------------------------------- */
int pps_latent_accumulate(const float* partial, const int32_t* ids, int64_t n, int c, float* latent, float* counts,
                          void* stream);
/* row-selecting variant: latent[ids[i]] += partial[rows[i]] (partial [*,c] point-major, c % 4 == 0); the caller passes each
 * destination once per call (de-duplicated pass), calls of different passes are ordered on the stream */
int pps_latent_accumulate_rows(const float* partial, const int32_t* rows, const int32_t* ids, int64_t n, int c, float* latent,
                               float* counts, void* stream);
int pps_latent_finalize(float* latent, const float* counts, int64_t n, int c, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * f3  marching cubes + bisection refinement on the device
 *     replaces  skimage.measure.marching_cubes + the refinement loop   source/poco_utils.py:96,111-168
 * volume [r,r,r] f32 (NaN = never decoded; cells with a NaN corner emit nothing).  The case table (ppsurf_b200/mc_tables.py,
 * [256,width] int8, -1 padded, bit i of the case = corner i below the level) and the three cell-edge tables are supplied by the
 * caller.  One vertex per crossed GRID edge (no duplicate vertices); vert_edge = 3 * linear index of the edge's lower grid vertex
 * + axis; verts are in volume-index coordinates like skimage's.
 *   pps_mc_count : counts_out (device int64[2]) = {vertices, triangles}; leaves the scans in `workspace`
 *   pps_mc_emit  : same arguments and workspace -> verts_out [nv,3] f32, vert_edge_out [nv] i32, faces_out [nt,3] i32
 *   pps_refine_init / pps_refine_update : bracketing state of the bisection (va, vb [nv,3]; pa, pb [nv]; v [nv,3]; active [nv] =
 *     vertex strictly inside its edge with both end values decoded) and one sweep given the occupancy `pred` at v
 * ------------------------------------------------------------------------------------------------------------- */
int pps_mc_set_edges(const int8_t* edge_corner_host, const int8_t* edge_axis_host, const int8_t* edge_origin_host);
size_t pps_mc_workspace_bytes(int r);
int pps_mc_count(const float* volume, int r, float level, const int8_t* tri_table, int width, void* workspace, size_t workspace_bytes,
                 int64_t* counts_out, void* stream);
int pps_mc_emit(const float* volume, int r, float level, const int8_t* tri_table, int width, const void* workspace, float* verts_out,
                int32_t* vert_edge_out, int32_t* faces_out, void* stream);
int pps_refine_init(const float* volume, int r, const float* verts, const int32_t* vert_edge, int64_t nv, float step, float bmin_pad,
                    float* va, float* vb, float* pa, float* pb, float* v, unsigned char* active, void* stream);
int pps_refine_update(const float* pred, int64_t n, float* va, float* vb, float* pa, float* pb, float* v, void* stream);

/* ===============================================================================================================
 * config 5  --  the training step (`pps.py fit`): forward in TRAIN mode and backward (SURVEY.md §8b "*_bwd")
 *     replaces  PocoModel.training_step -> network.forward -> cross_entropy -> Lightning backward (torch autograd)
 *               source/poco_model.py:75-88,108-125, source/ppsurf_model.py:70-117
 * Train mode differs from the predict path in ways that forbid the folded / fused predict kernels: BatchNorm uses BATCH
 * statistics (and updates its running statistics), FKAConvLayer updates norm_radius (nn.py:608-613), dropout is active,
 * and every activation has to be kept for the backward pass.  The training path is therefore a set of PRIMITIVES with
 * explicit forward / backward entry points; ppsurf_b200/autograd.py wraps each pair in a torch.autograd.Function and
 * ppsurf_b200/training.py composes the reference's network from them.  All tensors row-major f32 [rows, channels]; a
 * "group" is a run of consecutive rows.  No entry point allocates; scratch is passed in.
 * =============================================================================================================== */

/* Strided, batched GEMM:  C_b[m,n] (+)= sum_k A_b[m,k] . B_b[k,n] (+ bias[n])
 *   A_b[m,k] = a[b*sa_b + m*sa_m + k*sa_k],  B_b[k,n] = b[b*sb_b + k*sb_k + n*sb_n],  C_b[m,n] = c[b*sc_b + m*ldc + n]
 * Serves Y = X.W^T (Linear / Conv1d / Conv2d 1x1: nn.py, poco_model.py:400-411), dX = dY.W, dW = dY^T.X (k = rows; split over
 * the grid and reduced with atomics) and the per-query feature transform torch.bmm(trans2, x) (nn.py:347) with its two
 * gradients.  accumulate != 0 adds into C.  precision 0 = fp32 SIMT, 1 = bf16 operands on tcgen05 with fp32 accumulation
 * (BASELINE config 5 "bf16"; batch == 1 problems with m >= 64, n >= 16, k >= 32, others fall to fp32). */
int pps_gemm(const float* a, int64_t sa_b, int64_t sa_m, int64_t sa_k, const float* b, int64_t sb_b, int64_t sb_k, int64_t sb_n,
             float* c, int64_t sc_b, int64_t ldc, int64_t batch, int64_t m, int n, int64_t k, const float* bias, int accumulate,
             int precision, void* stream);
/* out[c] (+)= sum_r x[r*ld + c]   (bias gradients) */
int pps_colsum(const float* x, int64_t rows, int c, int64_t ld, float* out, int accumulate, void* stream);

/* Normalisation over the rows of each group, per channel, affine, with the following activation fused (act 0 none, 1 ReLU,
 * 2 SiLU):  y = act((x - mean_g) / sqrt(var_g + eps) * gamma + beta),  x [groups, rows, c].
 *   groups = 1        BatchNorm1d in train mode (nn.py:162-190,305-373,376-417,438-450,508-548); mean / var [c] feed
 *                     pps_bn_running_update (momentum 0.1, unbiased variance, like torch)
 *   groups = samples  InstanceNorm2d(affine=True) of FKAConvLayer (nn.py:586-587,630,638), rows = n_s * 16
 * mean / var [groups, c] are outputs of the forward and inputs of the backward (biased variance).  Backward: dx, and
 * dgamma / dbeta [c] summed over the groups.  workspace: pps_norm_workspace_bytes(groups, c). */
size_t pps_norm_workspace_bytes(int64_t groups, int c);
int pps_norm_fwd(const float* x, int64_t groups, int64_t rows, int c, const float* gamma, const float* beta, float eps, int act, float* y,
                 float* mean, float* var, void* workspace, size_t workspace_bytes, void* stream);
int pps_norm_bwd(const float* x, const float* dy, int64_t groups, int64_t rows, int c, const float* gamma, const float* beta,
                 const float* mean, const float* var, float eps, int act, float* dx, float* dgamma, float* dbeta, void* workspace,
                 size_t workspace_bytes, void* stream);
int pps_bn_running_update(const float* mean, const float* var, int64_t count, float momentum, int c, float* running_mean, float* running_var,
                          void* stream);

/* element-wise pairs: activation (0 none, 1 ReLU, 2 SiLU; backward takes the PRE-activation x), dropout with a stored keep-mask
 * (MLP, p = 0.3, nn.py:399), y = x * w[row] (the distance weights of nn.py:631,639,644), concat of per-row and per-group
 * channels torch.cat([mat, mp.expand], 1) (nn.py:634-635,642-643) */
int pps_act_fwd(const float* x, int64_t n, int act, float* y, void* stream);
int pps_act_bwd(const float* x, const float* dy, int64_t n, int act, float* dx, void* stream);
/* the mask is a hash of (element, seed + *draw_counter): draw_counter (nullable) is a DEVICE counter the caller advances between draws, so a
 * captured CUDA graph that is replayed draws a new mask every step */
int pps_dropout_fwd(const float* x, int64_t n, float p, uint32_t seed, const uint32_t* draw_counter, float* y, uint8_t* mask,
                    void* stream);
int pps_dropout_bwd(const float* dy, const uint8_t* mask, int64_t n, float p, float* dx, void* stream);
int pps_rowscale_fwd(const float* x, const float* w, int64_t rows, int c, float* y, void* stream);
int pps_rowscale_bwd(const float* x, const float* w, const float* dy, int64_t rows, int c, float* dx, float* dw, void* stream);
int pps_concat_bcast_fwd(const float* x, const float* v, int64_t groups, int s, int c, float* out, void* stream);
int pps_concat_bcast_bwd(const float* dout, int64_t groups, int s, int c, float* dx, float* dv, void* stream);

/* batch_gather (nn.py:655-674) as a row gather and its gradient (atomic scatter-add into a caller-zeroed dx) */
int pps_gather_rows(const float* x, const int32_t* idx, int64_t m, int c, float* y, void* stream);
int pps_scatter_add_rows(const float* dy, const int32_t* idx, int64_t m, int c, float* dx, void* stream);
/* y[g,c] = max_s x[g,s,c] * w[g,s] (w nullable), arg = winning s: the STN max over the patch (nn.py:170-172), the weighted
 * neighbourhood maxima of FKAConvLayer (nn.py:631-633,639-641), the global max of the U-Net (nn.py:535).  Backward fills
 * dx [groups,s,c] and, when dw is given, dw [groups,s] (gradient of the distance weights). */
int pps_seg_max_fwd(const float* x, const float* w, int64_t groups, int s, int c, float* y, int32_t* arg, void* stream);
int pps_seg_max_bwd(const float* dy, const int32_t* arg, const float* x, const float* w, int64_t groups, int s, int c, float* dx, float* dw,
                    void* stream);
/* max_pool(x, ids) (nn.py:677-680) with the winning source row kept for the backward (atomic adds into a caller-zeroed dx) */
int pps_gather_max_fwd(const float* x, const int32_t* ids, int64_t b, int64_t n_in, int64_t n_s, int c, int kn, float* y, int32_t* arg,
                       void* stream);
int pps_gather_max_bwd(const float* dy, const int32_t* arg, int64_t rows, int c, float* dx, void* stream);

/* attention pooling: prob = softmax over the s rows of a group per head, a = mean over the h heads, out = sum_s a[s] v[s,:]
 *     InterpAttentionKHeadsNet (poco_model.py:413-416; h = 64, s = k = 64) and AttentionPoco (nn.py:84-96; h = 1, s = P) */
int pps_attn_pool_fwd(const float* scores, const float* v, int64_t groups, int s, int h, int c, float* prob, float* a, float* out,
                      void* stream);
int pps_attn_pool_bwd(const float* dout, const float* prob, const float* a, const float* v, int64_t groups, int s, int h, int c,
                      float* dscores, float* dv, void* stream);

/* FKAConvLayer.forward geometry (nn.py:598-624): offs [R,3] = (pts[ids] - support) / norm_radius, dist [R], sig [R] =
 * sigmoid(-alpha d + beta), dw [R] = normalised distance weights (R = b * n_s * kn rows).  alpha / beta / norm_radius are DEVICE
 * scalars; with update_radius the norm_radius buffer is first moved towards the mean neighbourhood radius (train mode,
 * nn.py:608-613).  scratch: one double.  pps_fka_weights_bwd: d loss / d dw -> {d alpha, d beta} (two doubles). */
int pps_fka_geometry_fwd(const float* pts, const float* support, const int32_t* ids, int64_t b, int64_t n_in, int64_t n_s, int kn,
                         const float* alpha, const float* beta, float* norm_radius, float momentum, int update_radius, float* offs,
                         float* dist, float* sig, float* dw, double* scratch, void* stream);
int pps_fka_weights_bwd(const float* ddw, const float* sig, const float* dist, int64_t points, int kn, double* dalpha_dbeta, void* stream);
/* feat[p, c*16 + m] = sum_j x[ids[p,j], c] * mat[p,j,m]  (nn.py:647-649; the column order of cv.weight.view(cout, cin*16)) and its
 * gradients dmat [R,16] and dx [b*n_in, cin] (zero-filled inside, atomic adds) */
int pps_fka_feat_fwd(const float* x, const int32_t* ids, const float* mat, int64_t b, int64_t n_in, int64_t n_s, int kn, int cin, float* feat,
                     void* stream);
int pps_fka_feat_bwd(const float* dfeat, const float* x, const int32_t* ids, const float* mat, int64_t b, int64_t n_in, int64_t n_s, int kn,
                     int cin, float* dx, float* dmat, void* stream);

/* cross entropy of compute_loss (poco_model.py:75-88): per-row losses and their sum (one double); backward with a per-row scale */
int pps_ce_fwd(const float* logits, const int64_t* target, int64_t m, int c, float* loss_rows, double* loss_sum, void* stream);
int pps_ce_bwd(const float* logits, const int64_t* target, const float* scale, int64_t m, int c, float* dlogits, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PPSURF_B200_H */
