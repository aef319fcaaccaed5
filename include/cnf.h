/* cnf.h -- C ABI of libcnf_sm100.so: B200 (sm_100a) kernels for the category-level neural-field
 * training hot path of Taekbum/category-nerf-reconstruction-official.
 *
 * The reference has no FFI layer (it is pure Python/PyTorch); the boundary a
 * maintainer binds is therefore this header, one entry point per reference function on the path.
 * Each declaration cites the reference interface it replaces (paths relative to the reference repo).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller (PyTorch allocates everything, including
 *    workspaces); the library allocates nothing persistent and keeps no per-device global state;
 *  - `stream` is the cudaStream_t to launch on (torch.cuda.current_stream().cuda_stream);
 *    all entry points are asynchronous w.r.t. the host (no synchronisation inside);
 *  - tensors are dense row-major fp32 unless stated; a leading `n_cat` (K) dimension is the
 *    category batch the reference obtains with functorch.vmap (train.py:154-155);
 *  - return value: 0 = OK, <0 = argument/shape/arch error (see cnf_last_error_string),
 *    >0 = cudaError_t.  The library never calls exit() and has NO CPU fallback.
 */
#ifndef CNF_H_
#define CNF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* cnf_stream_t;

#define CNF_OK 0
#define CNF_ERR_ARG (-1)
#define CNF_ERR_ARCH (-2)
#define CNF_ERR_WORKSPACE (-3)
#define CNF_ERR_UNSUPPORTED (-4)

#define CNF_EMB_DIM 129   /* 3 + 6*21, src/trainer.py:20-21 */
#define CNF_EMB1 87       /* geometry branch input  */
#define CNF_EMB2 42       /* colour branch input    */
#define CNF_N_DIRS 21
#define CNF_N_FREQ 6

int cnf_version(void);
const char* cnf_last_error_string(void);
/* 0 if `device` is an sm_100 part, CNF_ERR_ARCH otherwise (no fallback exists). */
int cnf_device_check(int device);

/* ------------------------------------------------------------------------------------------------
 * (1) Sampler -- sceneCategory.sample_3d_points, stratified_bins, normal_bins_sampling
 *     (src/scene_cateogries.py:51-96, 453-546).  Random draws are INPUTS so results are bit-exact.
 *     rgbs u8 [K,R,4] (last column = label 0/1/2), depth [K,R], origins/dirs [K,R,3].
 *     u_inv [K,R,S], u_val [K,R,N], n_obj [K,R,M] (already N(0,eps/3)), u_oth [K,R,M]: per category,
 *     row k is consumed by the k-th ray (in ray order) of the respective case, as the reference's
 *     boolean-mask assignment does.  lin_S/lin_N/lin_M = torch.linspace(0,1,n+1) tables (not k/n).
 *     Outputs: z [K,R,S], pcs [K,R,S,3] (may be NULL), valid u8 [K,R], labels u8 [K,R],
 *     stats [K,8] int32 = {nI, nV, nO, nT, N_depth&obj, N_obj, N_sem, float_bits(max depth)}.
 *     ws: >= cnf_sample_workspace_bytes(K,R) bytes.
 *     max_depth_override [K] (may be NULL): use these upper bounds for invalid-depth rays instead of the batch
 *     maximum (scene_cateogries.py:486) -- needed when one category's batch is split across GPUs and the
 *     maximum has been all-reduced by the caller.
 * ---------------------------------------------------------------------------------------------- */
size_t cnf_sample_workspace_bytes(int n_cat, int n_rays);
int cnf_sample_fwd(int n_cat, int n_rays, int n_cam2surf, int n_surf,
                   const uint8_t* rgbs, const float* depth, const float* origins, const float* dirs,
                   const float* u_inv, const float* u_val, const float* n_obj, const float* u_oth,
                   const float* lin_S, const float* lin_N, const float* lin_M,
                   float min_bound, float eps, float stop_eps, int this_obj, const float* max_depth_override,
                   float* z, float* pcs, uint8_t* valid, uint8_t* labels, int32_t* stats,
                   void* ws, size_t ws_bytes, cnf_stream_t stream);

/* Count pass of the sampler alone: stats [K,8] as cnf_sample_fwd writes them (case counts, loss-mask counts
 * N_depth&obj / N_obj / N_sem = the denominators of src/render_rays.py:86, bits of the batch maximum depth,
 * src/scene_cateogries.py:486).  When one category's rays are split across GPUs every rank all-gathers its rows and
 * cnf_reduce_stats turns the gathered rows [world,K,8] into the GLOBAL counts int32 [K,4] (sum) and max depth [K] (max),
 * which the sampler (max_depth_override) and the fused step (counts) then use.  ws as for cnf_sample_fwd. */
int cnf_sample_stats(int n_cat, int n_rays, const uint8_t* rgbs, const float* depth, float min_bound, int this_obj,
                     int32_t* stats, void* ws, size_t ws_bytes, cnf_stream_t stream);
int cnf_reduce_stats(int world, int n_cat, const int32_t* gathered, int32_t* counts, float* max_depth, cnf_stream_t stream);

/* stratified_bins(min, max, n_bins, n_rays) module-level function (src/scene_cateogries.py:51-81;
 * also called by src/category_registration.py:146).  min_d/max_d: [n_rays] or broadcast scalar
 * (stride 0 when *_is_scalar).  u [n_rays,n_bins], lin [n_bins+1] -> z [n_rays,n_bins]. */
int cnf_stratified_bins(int n_rays, int n_bins, const float* min_d, int min_is_scalar,
                        const float* max_d, int max_is_scalar, const float* u, const float* lin,
                        float* z, cnf_stream_t stream);

/* origin_dirs_O / origin_dirs_W (src/scene_cateogries.py:24-47; called per step by get_training_samples :428-433):
 * T [n_rays,4,4] (per-ray camera pose T_CO or T_WC, 16-byte aligned), dirs_c [n_rays,3] ->
 * invert = 1: origins = inv(T)[:3,3], dirs_out = inv(T)[:3,:3] dirs_c (general 4x4 inverse);  invert = 0: the same with T itself. */
int cnf_origin_dirs(int n_rays, int invert, const float* T, const float* dirs_c, float* origins, float* dirs_out,
                    cnf_stream_t stream);

/* Device-resident ray pools -- the per-step slicing of sceneCategory.get_training_samples
 * (src/scene_cateogries.py:352-451) for all categories in one launch.  The pools of n_cat categories are
 * concatenated structure-of-arrays: rgbs u8 [P,4], depth [P], dirs_c [P,3] (camera frame), inst i64 [P]
 * (batch_indices_all, may be NULL), poses [F,4,4] (16-byte aligned) with pose_index i32 [P] (NULL = one pose per
 * pool row, t_co_batch_all :424-431; else the per-frame table t_wc_batch_dict[frame_batch] :373-379).
 * perm i32 [P]: pool row held at each (shuffled) position; seg i64 [n_cat] first position of each category;
 * start i64 [n_cat] = i_batch of each category.  Ray r of category k is pool row perm[seg[k]+start[k]+r];
 * outputs [n_cat,n_rays,...]: rgbs u8 x4, depth, origins, dirs (as cnf_origin_dirs), inst (may be NULL). */
int cnf_pool_gather(int n_cat, int n_rays, int invert, const int64_t* seg, const int64_t* start, const int32_t* perm,
                    const uint8_t* rgbs, const float* depth, const float* dirs_c, const float* poses,
                    const int32_t* pose_index, const int64_t* inst, uint8_t* out_rgbs, float* out_depth,
                    float* out_origins, float* out_dirs, int64_t* out_inst, cnf_stream_t stream);
/* Epoch reshuffle `pool = pool[rand_idx]` (src/scene_cateogries.py:439-449) as a composition of permutations:
 * perm_out[i] = perm_in[idx[i]], i < n (out of place; idx = randperm(n) of one category's segment). */
int cnf_pool_compose(int64_t n, const int32_t* perm_in, const int64_t* idx, int32_t* perm_out, cnf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (2) Embedding -- UniDirsEmbed.forward (src/embedding.py:82-92) and its backward.
 *     x [K,P,3], B [K,21,3], scale [K] -> emb [K,P,129] = cat(x/s, sin(pi * 2^f * (x/s) B^T)).
 *     bwd: g_emb [K,P,129] -> g_B [K,21,3] (ACCUMULATED into), g_x [K,P,3] (may be NULL).
 * ---------------------------------------------------------------------------------------------- */
int cnf_embed_fwd(int n_cat, int n_pts, const float* x, const float* B, const float* scale,
                  float* emb, cnf_stream_t stream);
int cnf_embed_bwd(int n_cat, int n_pts, const float* x, const float* B, const float* scale,
                  const float* g_emb, float* g_B, float* g_x, cnf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (3) MLPs, fp32 CUDA-core path (module-level drop-in, exact-precision reference on device).
 *     CodeNeRF.forward (src/model.py:56-84): params = 28 tensors in named_parameters() order, each
 *     stacked [K,...] and packed back-to-back in ONE flat buffer per category (layout from
 *     cnf_codenerf_param_count / offsets in SURVEY appendix D order).
 *     emb [K,P,129]; latent rows: shape_lat/tex_lat [K,NR,L] with point p using row p / pts_per_row
 *     (the reference passes [R,1,L] per-ray rows -> NR=R, pts_per_row=S; eval_points passes one
 *     code -> NR=1, pts_per_row=P).  Outputs sigma [K,P] (already x10), rgb [K,P,3].
 *     ws keeps the activations needed by the backward (cnf_codenerf_workspace_floats per category).
 * ---------------------------------------------------------------------------------------------- */
size_t cnf_codenerf_param_count(int W, int L);
size_t cnf_codenerf_workspace_floats(int n_pts, int n_rows, int W);
int cnf_codenerf_fwd(int n_cat, int n_pts, int n_rows, int pts_per_row, int W, int L,
                     const float* params, const float* emb, const float* shape_lat, const float* tex_lat,
                     float* sigma, float* rgb, float* ws, size_t ws_floats, cnf_stream_t stream);
/* g_params is ACCUMULATED into (caller zeroes); g_emb/g_shape_lat/g_tex_lat are overwritten (may be NULL). */
int cnf_codenerf_bwd(int n_cat, int n_pts, int n_rows, int pts_per_row, int W, int L,
                     const float* params, const float* emb, const float* shape_lat, const float* tex_lat,
                     const float* rgb, const float* g_sigma, const float* g_rgb,
                     float* ws, size_t ws_floats,
                     float* g_params, float* g_emb, float* g_shape_lat, float* g_tex_lat,
                     cnf_stream_t stream);

/*     OccupancyMap.forward (src/model.py:124-154), hidden size H (128 for the background model). */
size_t cnf_occmap_param_count(int H);
size_t cnf_occmap_workspace_floats(int n_pts, int H);
int cnf_occmap_fwd(int n_cat, int n_pts, int H, const float* params, const float* emb,
                   float* alpha, float* rgb, float* ws, size_t ws_floats, cnf_stream_t stream);
int cnf_occmap_bwd(int n_cat, int n_pts, int H, const float* params, const float* emb, const float* rgb,
                   const float* g_alpha, const float* g_rgb, float* ws, size_t ws_floats,
                   float* g_params, float* g_emb, cnf_stream_t stream);

/*     OccupancyMap.forward with hidden size 128 (the reference's background model, train.py:172-180) on tcgen05 tensor cores:
 *     bf16 operands, fp32 accumulation; same flat parameter layout as cnf_occmap_* with H = 128; ONE model (n_cat = 1).
 *     fwd: emb [P,129] -> alpha [P] (already x10), rgb [P,3]; keeps bf16 activation images in ws for the backward.
 *     bwd (after fwd on the same ws, same params): rgb (fwd output), g_alpha [P], g_rgb [P,3] ->
 *          g_params [cnf_occmap_param_count(128)] OVERWRITTEN (not accumulated), g_emb [P,129] (may be NULL).
 *     ws >= cnf_occmap_tc_workspace_bytes(P) (1568 B/point of activation images + 0.4 MiB). */
size_t cnf_occmap_tc_workspace_bytes(int n_pts);
int cnf_occmap_tc_fwd(int n_pts, const float* params, const float* emb, float* alpha, float* rgb,
                      void* ws, size_t ws_bytes, cnf_stream_t stream);
/*     The same with the embedding (UniDirsEmbed, src/embedding.py:82-92) computed inside the kernels: points [P,3], pe_B [21,3],
 *     pe_scale [1] instead of emb; no [P,129] embedding or embedding gradient exists.  bwd_points: g_B [21,3] OVERWRITTEN
 *     (the B_layer gradient); same workspace contract (fwd_points first, same ws and params). */
int cnf_occmap_tc_fwd_points(int n_pts, const float* params, const float* pe_B, const float* pe_scale, const float* points,
                             float* alpha, float* rgb, void* ws, size_t ws_bytes, int flags, cnf_stream_t stream);
int cnf_occmap_tc_bwd_points(int n_pts, const float* params, const float* points, const float* rgb, const float* g_alpha,
                             const float* g_rgb, void* ws, size_t ws_bytes, float* g_params, float* g_B, int flags,
                             cnf_stream_t stream);
/*     flags bit0: the workspace was prepared for this step by cnf_occmap_tc_prepare (operand images + constants from params / pe_B /
 *     pe_scale, zeroed gradient buffer): the parameter-only half of a step, which may run on another stream next to the sampler. */
int cnf_occmap_tc_prepare(int n_pts, const float* params, const float* pe_B, const float* pe_scale, void* ws, size_t ws_bytes,
                          cnf_stream_t stream);
/*     Forward-only evaluation at arbitrary points (Trainer.eval_points of the background model, src/trainer.py:125-151):
 *     occ [P] = sigmoid(alpha) (occupancy_activation), rgb [P,3]; no activation images are kept. */
size_t cnf_occmap_tc_eval_workspace_bytes(void);
int cnf_occmap_tc_eval_points(int n_pts, const float* params, const float* pe_B, const float* pe_scale, const float* points,
                              float* occ, float* rgb, void* ws, size_t ws_bytes, cnf_stream_t stream);
/*     Forward-only rendering of rays through the hidden-128 OccupancyMap (dense views of the background; any S): the counterpart
 *     of cnf_render_rays.  origins/dirs [R,3], z [R,S] -> out_rays [R,6] = {depth, variance, r, g, b, opacity}, out_w [R,S]
 *     termination probabilities (may be NULL). */
size_t cnf_occmap_tc_render_workspace_bytes(int n_rays, int S);
int cnf_occmap_tc_render_rays(int n_rays, int S, const float* params, const float* pe_B, const float* pe_scale,
                              const float* origins, const float* dirs, const float* z, float* out_rays, float* out_w,
                              void* ws, size_t ws_bytes, cnf_stream_t stream);
int cnf_occmap_tc_bwd(int n_pts, const float* params, const float* rgb, const float* g_alpha, const float* g_rgb,
                      void* ws, size_t ws_bytes, float* g_params, float* g_emb, cnf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (4) Compositing and losses -- occupancy_activation / occupancy_to_termination / render /
 *     render_loss / reduce_batch_loss (src/render_rays.py:3-95) and step_batch_loss (src/loss.py:18-74).
 * ---------------------------------------------------------------------------------------------- */
/* termination weights only: occ [n_rays,S] -> w [n_rays,S] (src/render_rays.py:25-33). */
int cnf_termination_fwd(int n_rays, int S, const float* occ, float* w, cnf_stream_t stream);
/* its backward (the reference function is differentiable): g_w [n_rays,S] -> g_occ [n_rays,S]. */
int cnf_termination_bwd(int n_rays, int S, const float* occ, const float* g_w, float* g_occ, cnf_stream_t stream);
/* Uncertainty fields of a rendered ray bundle (src/category_registration.py:152-162): termination weights w [n_rays,S] ->
 * out [n_rays,3] = {entropy = sum -w log(w + 1e-10), mass = sum w, heuristic = mass * exp(-0.5 entropy)}. */
int cnf_ray_uncertainty(int n_rays, int S, const float* w, float* out, cnf_stream_t stream);
/* mask counts for the loss denominators: labels u8 [K,R], mask_depth u8 [K,R] -> counts int32 [K,4]
 * = {N_depth&obj, N_obj, N_sem, 0}. */
int cnf_mask_counts(int n_cat, int n_rays, const uint8_t* labels, const uint8_t* mask_depth,
                    int32_t* counts, cnf_stream_t stream);
/* forward: alpha [K,R,S], color [K,R,S,3], z [K,R,S], gt_depth [K,R], gt_color [K,R,3] (0..1),
 * labels/mask_depth u8 [K,R], counts from cnf_mask_counts.
 * out_rays [K,R,6] = {D, var, C0, C1, C2, O};  out_loss [4 + 3K] =
 * {loss, any_empty_depth, any_empty_obj, any_empty_sem, depth[K], color[K], opacity[K]}.
 * ws: >= cnf_composite_workspace_bytes(K,R). */
size_t cnf_composite_workspace_bytes(int n_cat, int n_rays);
int cnf_composite_loss_fwd(int n_cat, int n_rays, int S, const float* alpha, const float* color,
                           const float* z, const float* gt_depth, const float* gt_color,
                           const uint8_t* labels, const uint8_t* mask_depth, const int32_t* counts,
                           float color_scaling, float opacity_scaling,
                           float* out_rays, float* out_loss, void* ws, size_t ws_bytes, cnf_stream_t stream);
/* backward: g_loss = device pointer to the upstream scalar gradient -> g_alpha [K,R,S], g_color [K,R,S,3]. */
int cnf_composite_loss_bwd(int n_cat, int n_rays, int S, const float* alpha, const float* color,
                           const float* z, const float* gt_depth, const float* gt_color,
                           const uint8_t* labels, const uint8_t* mask_depth, const int32_t* counts,
                           float color_scaling, float opacity_scaling, const float* g_loss,
                           float* g_alpha, float* g_color, cnf_stream_t stream);
/* forward and backward in ONE pass for S <= 32 (upstream gradient of the loss = 1): the backward only needs per-ray quantities and
 * the input-only mask counts, so loss terms and gradients leave the same kernel.  Same arguments as the two calls above. */
int cnf_composite_loss_fwd_bwd(int n_cat, int n_rays, int S, const float* alpha, const float* color,
                               const float* z, const float* gt_depth, const float* gt_color,
                               const uint8_t* labels, const uint8_t* mask_depth, const int32_t* counts,
                               float color_scaling, float opacity_scaling, float* out_rays, float* out_loss,
                               float* g_alpha, float* g_color, void* ws, size_t ws_bytes, cnf_stream_t stream);
/* step_batch_loss_reg (src/loss.py:5-15): codes [n_obj,L] -> out[0] = sum_i ||codes_i||_2 (0 if n_obj==1);
 * g_codes (may be NULL) += weight * codes_i/||codes_i||. */
int cnf_code_reg(int n_obj, int L, const float* codes, float weight, float* out, float* g_codes,
                 cnf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (5) Fused training step on tcgen05 tensor cores (bf16 operands, fp32 accumulation in TMEM):
 *     embed -> CodeNeRF(W=32) -> composite -> losses, forward AND backward in one launch, for K
 *     categories (train.py:136-182 minus the optimizer).  Samples z come from cnf_sample_fwd.
 *     See DESIGN.md for the tile pipeline.  Code tables are indexed in-kernel (never gathered).
 *
 *     params      [K, cnf_codenerf_param_count(32,L)]   fp32 master weights (appendix-D order)
 *     pe_B [K,21,3], pe_scale [K], shape_codes/tex_codes [K,n_obj_max,L], n_obj int32 [K]
 *     origins/dirs [K,R,3], z [K,R,S], gt_depth [K,R], gt_rgb u8 [K,R,4] (rgb + label),
 *     mask_depth u8 [K,R], inst int32 [K,R], counts int32 [K,4] (cnf_mask_counts / sampler stats+4)
 *     outputs (all OVERWRITTEN): g_params [K,n_param], g_B [K,21,3], g_shape/g_tex [K,n_obj_max,L],
 *     out_loss [4+3K] as in cnf_composite_loss_fwd plus reg terms appended: [.., reg_shape[K], reg_tex[K]].
 *     limits: S <= 32, 1 <= n_obj_max <= 4096 (CNF_ERR_UNSUPPORTED otherwise).
 *     flags: bit0 = also add the code regulariser (weight reg_scaling) to loss and code grads;
 *            bit1 = the workspace was prepared by cnf_train_step_prepare (same arguments) -- skip that half;
 *            bit2 = this call is a secondary part of a ray-split step whose outputs are summed over parts / ranks (with the GLOBAL
 *                   counts): the non-additive slots of out_loss (empty-mask flags [1..3], regulariser values) are written as 0;
 *            bits 8-15 = row stride of `counts` in int32 elements (0 means 4), so cnf_sample_fwd's stats[K,8] + 4 can be passed as is.
 * ---------------------------------------------------------------------------------------------- */
size_t cnf_train_step_workspace_bytes(int n_cat, int n_rays, int S, int L, int n_obj_max);
/* Optional first half of a step: everything that does not depend on the sampled depths.  `what` selects
 *   1 = the OPERAND half: fp16 weight images, constants and per-instance latent biases -- functions of the parameters only, so a
 *       training loop refreshes them once per optimizer step (like the low-precision weight copy of mixed-precision training);
 *   2 = the INPUT half: instance sort of this step's rays (inst) and the accumulator clear;   3 = both.
 * Call it on ANY stream (e.g. a side stream overlapping the sampler), make the stream of cnf_train_step wait for it, then pass
 * flags | 2 to cnf_train_step with the same arguments / workspace (both halves must have run since the last parameter change /
 * for this step's rays). */
int cnf_train_step_prepare(int n_cat, int n_rays, int S, int L, int n_obj_max,
                           const float* params, const float* pe_B, const float* pe_scale,
                           const float* shape_codes, const float* tex_codes, const int32_t* n_obj,
                           const int32_t* inst, int what, void* ws, size_t ws_bytes, cnf_stream_t stream);
int cnf_train_step(int n_cat, int n_rays, int S, int L, int n_obj_max,
                   const float* params, const float* pe_B, const float* pe_scale,
                   const float* shape_codes, const float* tex_codes, const int32_t* n_obj,
                   const float* origins, const float* dirs, const float* z,
                   const float* gt_depth, const uint8_t* gt_rgb, const uint8_t* mask_depth,
                   const int32_t* inst, const int32_t* counts,
                   float color_scaling, float opacity_scaling, float reg_scaling, int flags,
                   float* g_params, float* g_B, float* g_shape, float* g_tex, float* out_loss,
                   void* ws, size_t ws_bytes, cnf_stream_t stream);
/* The whole step in one call WITH the sampler inside the fused kernel (north_star: "a fused sample -> embed -> MLP -> composite ->
 * loss" step), for up to cnf_train_step_sampled_max_rays() rays per category (8192): no sampler launch and no z round trip through
 * HBM.  The prepare launch (instance sort of the rays) also ranks every ray inside its sampling case in ray order -- the row of the
 * random tensors the reference's boolean-mask assignment gives it (src/scene_cateogries.py:484-539) -- and writes the stats rows;
 * each role of the fused kernel then derives a point's depth in its tile prologue with the arithmetic of cnf_sample_fwd (bit-exact;
 * the sort of the M normal draws by warp shuffles).  Inputs as cnf_sample_fwd (rgbs, depth, origins, dirs, the four random tensors,
 * linspace tables, min_bound / eps / stop_eps / this_obj, max_depth_override) + inst; gt_depth = depth, gt_rgb = rgbs, the depth mask
 * is depth > min_bound (train.py:125-131).  Outputs as cnf_train_step + stats int32 [K,8] (as cnf_sample_fwd) and, if z_out != NULL,
 * the depths z [K,R,S] the kernel used (verification).  flags: bit0 / bit2 as cnf_train_step; bit3 (8) = the operand half of the
 * workspace is up to date (cnf_train_step_prepare(what = 1) has run since the last parameter change). */
int cnf_train_step_sampled_max_rays(void);
int cnf_train_step_sampled(int n_cat, int n_rays, int N, int M, int L, int n_obj_max,
                           const float* params, const float* pe_B, const float* pe_scale,
                           const float* shape_codes, const float* tex_codes, const int32_t* n_obj,
                           const uint8_t* rgbs, const float* depth, const float* origins, const float* dirs, const int32_t* inst,
                           const float* u_inv, const float* u_val, const float* n_draws, const float* u_oth,
                           const float* lin_S, const float* lin_N, const float* lin_M,
                           float min_bound, float eps, float stop_eps, int this_obj, const float* max_depth_override,
                           float color_scaling, float opacity_scaling, float reg_scaling, int flags,
                           float* g_params, float* g_B, float* g_shape, float* g_tex, float* out_loss, int32_t* stats,
                           float* z_out, void* ws, size_t ws_bytes, cnf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (5b) Forward-only tensor-core evaluation of CodeNeRF (W=32) at arbitrary points with ONE instance's codes:
 *      Trainer.eval_points (src/trainer.py:125-151) -- meshing grids, mesh-vertex colours, dense render sweeps.
 *      params [cnf_codenerf_param_count(32,L)], pe_B [21,3], pe_scale [1], shape_code / tex_code [L], points [P,3] ->
 *      occ [P] = sigmoid(10 * sigma) (occupancy_activation), rgb [P,3].  bf16 operands, fp32 accumulation.
 * ---------------------------------------------------------------------------------------------- */
size_t cnf_eval_points_workspace_bytes(void);
int cnf_eval_points(int n_points, int L, const float* params, const float* pe_B, const float* pe_scale,
                    const float* shape_code, const float* tex_code, const float* points, float* occ, float* rgb,
                    void* ws, size_t ws_bytes, cnf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (5c) Forward-only rendering of rays with ONE instance's codes (BASELINE config 5: dense 512x512 views per instance;
 *      the 96-sample uncertainty-field renders of src/category_registration.py:146-152): points origin + dir * z
 *      (src/scene_cateogries.py:543-544) -> UniDirsEmbed -> CodeNeRF -> occupancy_activation -> occupancy_to_termination
 *      -> render (src/render_rays.py:3-50, src/loss.py:44-48).  Any S >= 1.
 *      origins/dirs [R,3], z [R,S] -> out_rays [R,6] = {depth, variance, r, g, b, opacity};
 *      out_w [R,S] = termination probabilities (may be NULL).
 *      The tensor-core point kernel runs in chunks of 4 Mi points whose {occ, rgb} stay L2-resident for the per-ray
 *      compositing kernel; ws >= cnf_render_rays_workspace_bytes(R, S) (<= 64 MiB + 32 KiB).
 * ---------------------------------------------------------------------------------------------- */
size_t cnf_render_rays_workspace_bytes(int n_rays, int S);
int cnf_render_rays(int n_rays, int S, int L, const float* params, const float* pe_B, const float* pe_scale,
                    const float* shape_code, const float* tex_code, const float* origins, const float* dirs,
                    const float* z, float* out_rays, float* out_w, void* ws, size_t ws_bytes, cnf_stream_t stream);

/* Measurement hook: when both handles are non-NULL (cudaEvent_t created by the caller, e.g.
 * torch.cuda.Event(enable_timing=True).cuda_event), the NEXT cnf_train_step call on this host thread records
 * them on its stream immediately before / after the fused tcgen05 kernel (and only that kernel), then clears
 * the hook.  Lets bench.py time the dominant kernel live without a profiler. */
int cnf_set_kernel_events(void* ev_start, void* ev_stop);

/* Fused AdamW over a flat fp32 buffer (train.py:183): decoupled weight decay as
 * torch.optim.AdamW.  step is 1-based. */
int cnf_adamw_step(size_t n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                   float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                   cnf_stream_t stream);

/* Same update for up to CNF_ADAMW_MAX_GROUPS parameter groups (train.py:54-64: MLP + PE with lr / weight_decay, code tables
 * with code_lr / code_weight_decay) in ONE launch, with the step count kept on the device: state = int32[2] {steps done, 0},
 * zero-initialised by the caller; the kernel uses t = state[0] + 1 and increments it, so the call can be captured into a CUDA
 * graph and replayed (cnf_adamw_step takes t from the host and would freeze it). `groups` is a host array. */
#define CNF_ADAMW_MAX_GROUPS 8
typedef struct {
  float* param; const float* grad; float* exp_avg; float* exp_avg_sq;
  size_t n; float lr; float weight_decay;
} cnf_adamw_group;
int cnf_adamw_groups(int n_groups, const cnf_adamw_group* groups, float beta1, float beta2, float eps,
                     int32_t* state, cnf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CNF_H_ */
