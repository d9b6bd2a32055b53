/*
 * xsup_b200 — C ABI of the B200-native integral + multi-hypothesis reprojection-loss path.
 *
 * The reference (Charrrrrlie/X-as-Supervision) is pure Python and has no operator/plugin
 * interface of its own; the seam is the parameter-free code after `self.net(x)`.  Each entry
 * point below names the reference lines it replaces.  Plain pointers and sizes only: no torch
 * types, no allocation, no hidden synchronisation.  See INTEGRATION.md for the reference-side
 * binding (ctypes stub inside `KPDetector3DMulti.forward` / `Counter3DModel.forward`).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller, 16-byte aligned, contiguous;
 *   - `stream` is a `cudaStream_t` passed as `void*`; all work is enqueued on it, nothing
 *     synchronises, every call is CUDA-graph capturable;
 *   - return value 0 = success; negative = rejected before launch (XSUP_E_*), positive = a
 *     `cudaError_t` from the launch.  `xsup_last_error()` returns a thread-local message;
 *   - there is NO CPU fallback: without a CUDA device the compute calls fail with a
 *     cudaError, they never compute on the host.
 */
#ifndef XSUP_B200_H_
#define XSUP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XSUP_ABI_VERSION 11

enum { XSUP_F32 = 0, XSUP_BF16 = 1 };
enum { XSUP_HEAD_MULTI = 0, XSUP_HEAD_SINGLE = 1 };
enum { XSUP_REDUCE_BATCH = 0, XSUP_REDUCE_SAMPLE = 1, XSUP_REDUCE_JOINT = 2 };
enum {
    XSUP_OK = 0,
    XSUP_E_SHAPE = -1,     /* unsupported / inconsistent shape (D != W, NS even, NH > D-2, ...) */
    XSUP_E_ALIGN = -2,     /* pointer not 16-byte aligned */
    XSUP_E_NULL = -3,      /* required pointer is NULL */
    XSUP_E_DTYPE = -4,
    XSUP_E_DEVICE = -5     /* not an sm_100 device / no device */
};

/* One heat-map volume batch: logits are `[B, K*D, H, W]` contiguous (w fastest), i.e. each
 * (b,k) "unit" is one contiguous D*H*W block — keypoint_detector_integral_multi.py:69-74. */
typedef struct {
    int32_t B, K, D, H, W;
    int32_t NH;        /* num_hypo      (keypoint_detector_integral_multi.py:17); 1 in single mode */
    int32_t NS;        /* neighbor_size (…:18), odd; ignored in single mode */
    int32_t dtype;     /* XSUP_F32 | XSUP_BF16 : element type of logits and of g_logits */
    int32_t head;      /* XSUP_HEAD_MULTI (…_multi.py:66-88) | XSUP_HEAD_SINGLE (keypoint_detector_integral.py:45-65) */
} xsup_shape_t;

/* Per-sample camera tensors exactly as the reference's data loader hands them over
 * (human_utils/dataloader/dataloader.py:170-182; unpacked at modules/util.py:129-134). */
typedef struct {
    const float* trans_image;  /* [B,2,3] crop affine          */
    const float* pelvis;       /* [B,3]   root joint, camera mm */
    const float* k_mat;        /* [B,3,3] intrinsics            */
    const float* trans_world;  /* [B,3]   extrinsic translation */
    const float* rot_world;    /* [B,3,3] extrinsic rotation    */
} xsup_cam_t;

/* Loss graph of one camera: modules/model.py:71-79 (world lift per hypothesis),
 * :105-114 (symmetry, min over hypotheses), :158-162 (pseudo-GT MSE, min over hypotheses). */
typedef struct {
    int32_t B, K, NH;
    int32_t img_h, img_w;     /* params['{mode}_img'].shape[-2:]  (util.py:130,137-138) */
    float rect_width;         /* RECT_WIDTH (util.py:128), 2000 mm */
    float w_mse;              /* smpl_pseudo_img_loss.weight (model.py:164) */
    float w_bone, w_kp, w_kp2d; /* symmetry_loss.weight.{bone,kp,kp_2d} (model.py:108-112) */
    int32_t use_sym;          /* 0: symmetry term absent from the loss config */
    int32_t reduction;        /* XSUP_REDUCE_* */
    int32_t batch_total;      /* denominator batch size: B, or the global batch when sharded */
} xsup_loss_cfg_t;

#define XSUP_LOSS_TERMS 4     /* mse, bone, kp, kp2d */

int xsup_abi_version(void);
const char* xsup_last_error(void);
/* kernels launched by this library since load (all threads); evidence for bench.py's gpu_launches */
uint64_t xsup_launch_count(void);

/* floats per (b,k) unit of the saved-for-backward block / of the backward coefficient block */
size_t xsup_stats_stride(const xsup_shape_t* s);
size_t xsup_coef_stride(const xsup_shape_t* s);
/* total floats the caller must allocate for `stats` / `coef_ws`: B*K*stride plus a few words the kernels
 * use as their work-claim counter (SMs of a B200 see different HBM bandwidth, so work is claimed, not dealt) */
size_t xsup_stats_floats(const xsup_shape_t* s);
size_t xsup_coef_floats(const xsup_shape_t* s);

/* Replaces keypoint_detector_integral_multi.py:69-88 (softmax over D*H*W, the three marginals,
 * x/y expectations, find_peak + topk, windowed depth expectation, normalisation, assembly).
 *   logits          [B,K*D,H,W]   (s->dtype)
 *   kps             [B,NH,K,3]    fp32
 *   depth_prob_map  [K,D]         fp32, sample 0 (…:48)
 *   peak_idx        [B,K,NH]      int64 (…:24-34); NULL allowed
 *   stats           [xsup_stats_floats]      fp32, consumed by xsup_integral_bwd            */
int xsup_integral_fwd(const void* logits, float* kps, float* depth_prob_map, int64_t* peak_idx,
                      float* stats, const xsup_shape_t* s, void* stream);

/* Replaces KPDetector3DMulti.find_peak (keypoint_detector_integral_multi.py:24-34) on a depth
 * marginal `pz [rows, D]` fp32 -> `idx [rows, NH]` int64 (ties and filler slots: lowest bin first). */
int xsup_find_peak(const float* pz, int64_t* idx, int32_t rows, int32_t D, int32_t NH, void* stream);

/* The autograd of the above in one pass over the volume.
 *   g_kps    [B,NH,K,3] fp32  (d loss / d kps)
 *   g_logits [B,K*D,H,W]      (s->dtype); may alias `logits` for an in-place gradient
 *   coef_ws  [xsup_coef_floats] fp32 scratch                                               */
int xsup_integral_bwd(const void* logits, const float* stats, const float* g_kps, void* g_logits,
                      float* coef_ws, const xsup_shape_t* s, void* stream);

/* Replaces modules/util.py:128-152 (convert_patch_to_world = :61-82 then :85-95) on `[B,J,3]`.
 * flags: bit0 is_norm, bit1 mono (:145-150), bit2 patch.  `_bwd` is its vector-Jacobian product
 * (it takes the forward input `kps` again: the Jacobian depends on the projected point). */
int xsup_patch_to_world_fwd(const float* kps, const xsup_cam_t* cam, float* world, int32_t B, int32_t J,
                            int32_t img_h, int32_t img_w, float rect_width, int32_t flags, void* stream);
int xsup_patch_to_world_bwd(const float* kps, const float* g_world, const xsup_cam_t* cam, float* g_kps, int32_t B, int32_t J,
                            int32_t img_h, int32_t img_w, float rect_width, int32_t flags, void* stream);
/* Replaces modules/util.py:155-168 (convert_world_to_patch = :116-125 then :98-113): the forward
 * perspective projection, inverse of the above. */
int xsup_world_to_patch_fwd(const float* world, const xsup_cam_t* cam, float* kps, int32_t B, int32_t J,
                            int32_t img_h, int32_t img_w, float rect_width, int32_t flags, void* stream);

/* The same geometry stage by stage, with the reference's free parameters (modules/util.py:61-125): every one of
 * convert_patch_to_image (:61-82), convert_image_to_world (:85-95), convert_image_to_patch (:98-113),
 * convert_world_to_image (:116-125) and the two composites (:128-152, :155-168) is one call, selected by `flags`:
 *   XSUP_GEOM_PATCH_STAGE   the patch <-> image stage (crop affine `trans_image`, px <-> mm depth via `depth_scale`, `pelvis`)
 *   XSUP_GEOM_CAMERA_STAGE  the image <-> world stage (pinhole fx, fy, cx, cy; extrinsics rot_world / trans_world)
 *   XSUP_GEOM_NORM          is_norm (patch coordinates in [-1,1] / [0,1])        XSUP_GEOM_MONO  util.py:145-150
 * fx, fy, cx, cy: element b at [b*intr_stride] (1 for the reference's [B,1] tensors; for a k_mat [B,3,3] pass
 * k_mat+0, k_mat+4, k_mat+2, k_mat+5 with stride 9).  Tensors of a stage that is not selected may be NULL.
 * `xsup_geom_patch_to_world` runs patch -> image -> world (the selected stages), `xsup_geom_world_to_patch` runs
 * world -> image -> patch; `in`, `out` are [B,J,3].  The `_vjp` forms take the forward input again plus g_out = d loss / d out
 * and write g_in = d loss / d in. */
enum { XSUP_GEOM_NORM = 1, XSUP_GEOM_MONO = 2, XSUP_GEOM_PATCH_STAGE = 4, XSUP_GEOM_CAMERA_STAGE = 8 };
typedef struct {
    int32_t B, J;
    int32_t img_d, img_h, img_w;   /* image_depth, image_height, image_width (util.py:61) */
    float depth_scale;             /* mm per depth pixel: RECT_WIDTH / image width in the composites (util.py:138) */
    int32_t flags;
    int32_t intr_stride;
    const float* trans_image;      /* [B,2,3] */
    const float* pelvis;           /* [B,3]   */
    const float* fx; const float* fy; const float* cx; const float* cy;
    const float* trans_world;      /* [B,3]   */
    const float* rot_world;        /* [B,3,3] */
} xsup_geom_t;
int xsup_geom_patch_to_world(const float* in, float* out, const xsup_geom_t* g, void* stream);
int xsup_geom_patch_to_world_vjp(const float* in, const float* g_out, float* g_in, const xsup_geom_t* g, void* stream);
int xsup_geom_world_to_patch(const float* in, float* out, const xsup_geom_t* g, void* stream);
int xsup_geom_world_to_patch_vjp(const float* in, const float* g_out, float* g_in, const xsup_geom_t* g, void* stream);

/* Replaces the per-hypothesis Python loops of modules/model.py:71-79,105-114,158-162 and the loss
 * primitives modules/base_losses/loss_func.py:18-52 for one camera.
 *   kps      [B,NH,K,3]  from xsup_integral_fwd        target [B,K,3] pseudo joints
 *   world    [B,NH,K,3]  out: world mm
 *   sample_terms [B,XSUP_LOSS_TERMS,NH] out: per-sample un-normalised sums (16-byte aligned: rows are summed as float4s)
 *   partial  [XSUP_LOSS_TERMS,NH] out: their fixed-order sum over this rank's batch
 *            (the only thing that crosses ranks: one all-reduce(SUM) in global scope)          */
int xsup_reproj_loss_fwd(const float* kps, const float* target, const xsup_cam_t* cam, float* world,
                         float* sample_terms, float* partial, const xsup_loss_cfg_t* cfg, void* stream);

/* The one exchange step of the path when the batch is sharded over GPUs and the winner is chosen on the
 * GLOBAL batch: all-reduce(SUM) of `partial[n]` (n <= XSUP_XCHG_SLOT-1 floats: [XSUP_LOSS_TERMS, NH] for every NH the head accepts) done by ONE kernel over NVLink peer memory.
 * Every rank owns a zero-initialised mailbox of xsup_xchg_floats(world) floats that all peers have mapped
 * (e.g. torch symmetric memory); `peer_bufs` is a DEVICE array of the `world` mailbox addresses as seen from
 * this process (own included).  The kernel stores its partial sums into slot [step&1][rank] of every mailbox
 * (P2P stores + st.release.sys flag), acquire-spins on the `world` flags of its own mailbox, and sums the slots
 * in rank order, so all ranks end with bit-identical sums.  `step` is the call sequence number (1, 2, ...),
 * identical on all ranks.  All ranks must make the call; a peer that never arrives turns the result into NaN
 * after ~10 s instead of hanging the GPU.  Replaces nothing in the reference (it never reduces across ranks,
 * model.py:114,162 run on the rank-local batch); it exists so that N GPUs reproduce the single-process result. */
typedef struct {
    void* const* peer_bufs;
    int32_t rank, world;
    uint32_t step;
    uint32_t* seq;   /* DEVICE counter (zero-initialised, private to this rank) or NULL.  When given, the kernel itself
                      * takes the sequence number as ++(*seq) and `step` is ignored: the call can then be captured in
                      * a CUDA graph and replayed (a by-value step would repeat).  All ranks must make the same calls. */
    uint32_t* err;   /* DEVICE word (zero-initialised) or NULL: set to 1, and never cleared, when a wait for a peer timed
                      * out.  The sums of that call are NaN and the mailboxes are out of step from then on: the host
                      * must check the word (it costs a read-back, so not on every step) and rebuild the exchange. */
} xsup_xchg_t;
#define XSUP_XCHG_SLOT 1024   /* floats per (parity, source rank) slot: n data floats + the flag in the last word */
size_t xsup_xchg_floats(int32_t world);
int xsup_partial_allreduce(float* partial, int32_t n, const xsup_xchg_t* x, void* stream);

/* min / argmin over hypotheses (model.py:114,162; loss_func.py:59; eval.py:138-145).
 *   loss [2] out: (pseudo term * w_mse, symmetry term)
 *   sel  out int64: batch -> [2] (slot or -1); sample -> [2,B]; joint -> [B,K]                 */
int xsup_reproj_select(const float* kps, const float* target, const float* sample_terms, const float* partial,
                       float* loss, int64_t* sel, const xsup_loss_cfg_t* cfg, void* stream);

/* d (g_loss[0]*pseudo + g_loss[1]*symmetry) / d kps -> g_kps [B,NH,K,3]; only the selected slots
 * receive gradient (torch.min semantics; exact ties resolve to the lowest slot).              */
int xsup_reproj_loss_bwd(const float* kps, const float* target, const xsup_cam_t* cam, const int64_t* sel,
                         const float* g_loss, float* g_kps, const xsup_loss_cfg_t* cfg, void* stream);

/* The three calls above (and the exchange between them) as ONE launch: per-(sample, hypothesis) world lift + loss
 * terms, then the last CTA to finish forms the fixed-order batch sums, all-reduces them over NVLink peer memory when
 * `xchg` is given (reduction 'batch'; for 'sample' / 'joint' it is the reported loss that is summed over ranks) and
 * selects the slots.  Bit-identical to xsup_reproj_loss_fwd -> [xsup_partial_allreduce] -> xsup_reproj_select.
 *   xchg    NULL (rank-local selection, what the reference does under DDP) or the mailbox description
 *   ticket  one zero-initialised DEVICE word; the kernel leaves it zero again.  xsup_integral_fwd zeroes the
 *           XSUP_SCHED_WORDS ints that follow the statistics (stats + B*K*xsup_stats_stride): word 0 is its own
 *           work-claim counter, word 1 is free for this ticket. */
#define XSUP_SCHED_WORDS 16
int xsup_reproj_fused_fwd(const float* kps, const float* target, const xsup_cam_t* cam, float* world, float* sample_terms,
                          float* partial, float* loss, int64_t* sel, const xsup_loss_cfg_t* cfg, const xsup_xchg_t* xchg,
                          uint32_t* ticket, void* stream);

/* Backward of the fused per-camera op up to the coefficient blocks of the streaming head backward, as ONE launch:
 * xsup_reproj_loss_bwd, the sum with upstream gradients on kps / kps_world, and xsup_integral_coef, without writing
 * d loss / d kps to HBM.  Follow it with xsup_integral_bwd_apply.
 *   g_lp, g_ls   DEVICE scalars d L / d loss_pseudo, d L / d loss_sym (NULL = 0)
 *   g_kps_in     [B,NH,K,3] upstream gradient on kps (e.g. from the skeleton rasteriser, model.py:91) or NULL
 *   g_world      [B,NH,K,3] upstream gradient on kps_world (e.g. the generator loss with use_aug, model.py:138) or NULL
 *   stats        from xsup_integral_fwd;  coef_ws [xsup_coef_floats] out;  g_kps_out [B,NH,K,3] out or NULL */
int xsup_reproj_fused_bwd(const float* kps, const float* target, const xsup_cam_t* cam, const int64_t* sel, const float* g_lp,
                          const float* g_ls, const float* g_kps_in, const float* g_world, const float* stats, float* coef_ws,
                          float* g_kps_out, const xsup_loss_cfg_t* cfg, const xsup_shape_t* s, void* stream);
/* The streaming half of xsup_integral_bwd: coefficient blocks (from xsup_reproj_fused_bwd or xsup_integral_coef) ->
 * g_logits in one read + one write of the volume. */
int xsup_integral_bwd_apply(const void* logits, const float* coef_ws, void* g_logits, const xsup_shape_t* s, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Skeleton rasteriser + mask-reconstruction loss.
 *
 * Lines run from joint child[l] (start) to joint parent[l] (end), as built by cal_links
 * (modules/model.py:8-22).  `kps` holds the 2-D patch coordinates in [-1,1]: element (b, j) is at
 * kps[b*kp_batch_stride + j*kp_joint_stride + {0,1}], so the reference's strided view
 * kps_ori[cam][:, 0, :, :2] of the [B,NH,K,3] head output (model.py:91) is passed without a copy
 * (kp_batch_stride = NH*K*3, kp_joint_stride = 3). */
#define XSUP_MAX_LINES 32
typedef struct {
    int32_t B, K;              /* samples, joints */
    int32_t S;                 /* image_size: the heat-maps are S x S (x['{cam}_img'].shape[-1]); S % 4 == 0 */
    int32_t L;                 /* number of lines, 1..XSUP_MAX_LINES; L >= 21 halves the width of lines 11,12,14,15 (util.py:50-53) */
    int32_t kp_batch_stride, kp_joint_stride;
    float body_width;          /* cfg body_width * 1e-3 (model.py:31-32) */
    int32_t parent[XSUP_MAX_LINES];
    int32_t child[XSUP_MAX_LINES];
} xsup_skel_t;

/* compute_mask_reconstruction_loss (modules/base_losses/loss_func.py:4-16) over n = numel(mask) elements:
 *   XSUP_MASK_MSE        weight=None, use_clip=False : mean((m-gt)^2)
 *   XSUP_MASK_CLIP_MEAN  weight=None, use_clip=True  : the reference returns the TENSOR mean((m-gt)^2) * (m>0.1),
 *                        which the trainer reduces with .mean() (train.py:182); this mode is that mean
 *   XSUP_MASK_WEIGHTED   weight given               : mean((m-gt)^2 * [m>0.1 if use_clip] * weight)          */
enum { XSUP_MASK_MSE = 0, XSUP_MASK_CLIP_MEAN = 1, XSUP_MASK_WEIGHTED = 2 };
typedef struct {
    int64_t n;
    int32_t mode;
    int32_t use_clip;          /* XSUP_MASK_WEIGHTED only; CLIP_MEAN implies the filter, MSE ignores it */
} xsup_mask_loss_t;
#define XSUP_MASK_SUMS 4      /* loss_sums: sum (m-gt)^2, sum filter, sum (m-gt)^2*filter*weight, loss */

/* floats of scratch the calls below need (per-CTA partial sums; nothing is kept between calls) */
size_t xsup_skel_ws_floats(const xsup_skel_t* s);
size_t xsup_draw_lines_ws_floats(const xsup_skel_t* s);
size_t xsup_mask_loss_ws_floats(int64_t n);

/* Replaces draw_lines (modules/util.py:21-59): heat [B,L,S,S] fp32 = exp(-c_l * dist^2(pixel, segment l) / body_width).
 * `_bwd`: g_kps [B,K,2] (contiguous) = vector-Jacobian product with g_heat [B,L,S,S]. */
int xsup_draw_lines_fwd(const float* kps, const xsup_skel_t* s, float* heat, void* stream);
int xsup_draw_lines_bwd(const float* kps, const xsup_skel_t* s, const float* heat, const float* g_heat, float* g_kps,
                        float* ws, void* stream);

/* Replaces draw_lines + torch.max(heatmaps, dim=1, keepdim=True)[0] (modules/model.py:91-96) without
 * materialising the L heat-maps, optionally fused with compute_mask_reconstruction_loss of the result
 * against `gt` / `weight` [B,1,S,S] (model.py:185-188).
 *   recon     [B,1,S,S] fp32 out
 *   line_idx  [B,S,S]   uint8 out: winning line per pixel (lowest index on exact ties), consumed by _bwd
 *   loss      NULL -> no loss (gt, weight, loss_sums ignored); else loss_sums[XSUP_MASK_SUMS] out
 * `_bwd`: g_kps [B,K,2] = d( <g_recon, recon> + g_loss * loss ) / d kps; g_recon may be NULL (no other consumer
 * of recon), loss may be NULL (then gt, weight, loss_sums, g_loss are ignored); g_loss is a DEVICE scalar. */
int xsup_skeleton_mask_fwd(const float* kps, const xsup_skel_t* s, float* recon, uint8_t* line_idx, const float* gt,
                           const float* weight, const xsup_mask_loss_t* loss, float* loss_sums, float* ws, void* stream);
int xsup_skeleton_mask_bwd(const float* kps, const xsup_skel_t* s, const float* recon, const uint8_t* line_idx,
                           const float* g_recon, const float* gt, const float* weight, const xsup_mask_loss_t* loss,
                           const float* loss_sums, const float* g_loss, float* g_kps, float* ws, void* stream);

/* compute_mask_reconstruction_loss on an arbitrary mask tensor (the physique network's output, model.py:176).
 * filter_out (nullable): the (m > 0.1) map as floats, for callers that need the reference's tensor-valued result.
 * `_bwd`: g_mask[n] = g_loss * d loss / d mask. */
int xsup_mask_loss_fwd(const float* mask, const float* gt, const float* weight, float* filter_out,
                       const xsup_mask_loss_t* cfg, float* loss_sums, float* ws, void* stream);
int xsup_mask_loss_bwd(const float* mask, const float* gt, const float* weight, const xsup_mask_loss_t* cfg,
                       const float* loss_sums, const float* g_loss, float* g_mask, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Eval-side selection + triangulation, discriminator-side glue. */

/* Replaces eval.py:122-148 + eval_utils.py:7-41 for one camera: normalise the pixel-space ground truth
 * `joints_px [B,K,3]` (x,y -> [-1,1], z -> [0,1] by img_size-1), undo left/right swaps per hypothesis
 * (`switch_points`, per joint, decided on the L1 error in x,y; perm[k] = the joint k swaps with), and per joint keep
 * the hypothesis with the smallest squared error (best != 0, eval mode 'best') or hypothesis 0 ('confident').
 *   kps [B,NH,K,3] -> kp3d [B,K,3], kp2d [B,K,2] (selected on the 2-D error), is_trans [B,K] uint8 (of the LAST
 *   hypothesis, as the reference's loop leaves it), err2d [B] (`per_act_mse`), best_idx / best_2d_idx [B,K] int64,
 *   gt_norm [B,K,3].  is_trans, err2d, best_idx, best_2d_idx, gt_norm may be NULL.  K <= 32. */
typedef struct {
    int32_t B, NH, K;
    float img_size;            /* Eval.img_size (eval.py:72), 256; 0 = `joints_px` is already normalised (switch_points on its own, eval_utils.py:7) */
    int32_t best;              /* 1: mode 'best', 0: 'confident' */
    int32_t perm[32];
} xsup_eval_t;
int xsup_eval_select(const float* kps, const float* joints_px, const xsup_eval_t* cfg, float* kp3d, float* kp2d,
                     uint8_t* is_trans, float* err2d, int64_t* best_idx, int64_t* best_2d_idx, float* gt_norm, void* stream);

/* Replaces triangulation / batch_triangulate (modules/util.py:171-230): V <= 8 cameras, each with its patch
 * keypoints kps[v] [B,K,3] and camera tensors; world [B,K,3] out = the DLT solution (right singular vector of
 * the smallest singular value of the [2V,4] system, rows weighted by the metric depth as in the reference). */
#define XSUP_MAX_VIEWS 8
typedef struct {
    int32_t V, B, K;
    int32_t img_h, img_w, is_norm;
    float rect_width;
    const float* kps[XSUP_MAX_VIEWS];
    xsup_cam_t cam[XSUP_MAX_VIEWS];
} xsup_tri_t;
int xsup_triangulate(const xsup_tri_t* t, float* world, void* stream);

/* modules/model.py:123-124, literally `(w - w[:, [0], :]) / 1000` then the first `dim` (DISC_SUP_DIMENSION) of every
 * coordinate triple: world [N, M, R] -> out [N, M, R/3*dim], item 0 along the second axis subtracted.  For the stacked
 * [B, NH, K, 3] tensor the reference passes, N = B, M = NH, R = 3K: every hypothesis relative to HYPOTHESIS 0 (what
 * that expression means on a 4-D tensor); for [B, K, 3], N = B, M = K, R = 3: relative to the root joint.  `_bwd` is the
 * vector-Jacobian product (g_world [N, M, R]). */
int xsup_root_centre_fwd(const float* world, float* out, int32_t N, int32_t M, int32_t R, int32_t dim, void* stream);
int xsup_root_centre_bwd(const float* g_out, float* g_world, int32_t N, int32_t M, int32_t R, int32_t dim, void* stream);

/* One term of compute_disc_loss (modules/base_losses/loss_func.py:54-76) on logits [B,NH,C]:
 * loss = mean over (b,c) of min over h of (x - target)^2; sel [B,C] int64 = the argmin (first minimum).  NH = 1
 * is the 2-D-input case.  `_bwd`: g_logits [B,NH,C] = g_loss * d loss / d logits (only the selected slot). */
int xsup_disc_min_loss_fwd(const float* logits, int32_t B, int32_t NH, int32_t C, float target, float* loss, int64_t* sel,
                           void* stream);
int xsup_disc_min_loss_bwd(const float* logits, const int64_t* sel, const float* g_loss, int32_t B, int32_t NH, int32_t C,
                           float target, float* g_logits, void* stream);

/* ------------------------------------------------------------------------------------------------
 * The head's final 1x1 convolution fused into the integral head, forward only.
 * Replaces `Conv2d(C, K*D, 1)` (modules/integral_base_modules/deconv_head.py:33-35, the last layer of `self.net`)
 * followed by keypoint_detector_integral_multi.py:69-88 on the eval path (eval.py:120): the logits never touch HBM.
 *   x_nhwc  [B, H, W, C]  bf16, channels-last (torch.channels_last storage of a [B,C,H,W] tensor)
 *   weight  [K*D, C]      bf16 (the conv weight [K*D, C, 1, 1])
 *   bias    [K*D]         fp32 or NULL
 *   kps, depth_prob_map, peak_idx, stats: as xsup_integral_fwd (s->dtype is ignored: the operands are bf16, the
 *   accumulation and all statistics are fp32); logits_out: NULL, or [B, K*D, H, W] fp32 to also materialise the
 *   logits (validation).  Constraints: 128 % D == 0, (H*W) % 128 == 0, W % 32 == 0, C % 64 == 0, C <= 256. */
/* [B, C, H*W] fp32 (NCHW-contiguous activations of a plain backbone) -> [B, H*W, C] bf16, the operand layout of
 * xsup_conv_head_fwd, in one pass (transpose + round-to-nearest-even).  C % 64 == 0, (H*W) % 64 == 0. */
int xsup_pack_nhwc_bf16(const float* x_nchw, void* x_nhwc_bf16, int32_t B, int32_t C, int32_t HW, void* stream);
int xsup_conv_head_fwd(const void* x_nhwc, const void* weight, const float* bias, float* kps, float* depth_prob_map,
                       int64_t* peak_idx, float* stats, float* logits_out, const xsup_shape_t* s, int32_t C, void* stream);
/* The same with fp32 operands and tf32 tensor-core arithmetic (`tcgen05.mma.kind::tf32`, fp32 accumulation): the precision of
 * the reference's own `Conv2d` on this GPU (PyTorch runs cuDNN convolutions in TF32 by default).  x_nhwc_f32 [B, H, W, C] fp32
 * channels-last, weight_f32 [K*D, C] fp32; everything else as xsup_conv_head_fwd; same shape constraints.  About twice the
 * tensor-core time of the bf16 call (half-rate MMAs, 64-pixel tiles); forward only - the backward call rounds to bf16. */
int xsup_conv_head_fwd_tf32(const void* x_nhwc_f32, const void* weight_f32, const float* bias, float* kps, float* depth_prob_map,
                            int64_t* peak_idx, float* stats, float* logits_out, const xsup_shape_t* s, int32_t C, void* stream);

/* Backward of the conv-fused head, logits-free and on the tensor cores (csrc/conv_head_bwd.cu):
 *   (1) xsup_integral_coef turns g_kps + the saved statistics into the per-unit coefficient blocks (the first half of
 *       xsup_integral_bwd);
 *   (2) xsup_conv_head_bwd recomputes the logit tiles with tcgen05.mma, forms d loss / d logits in the epilogue, keeps it in
 *       shared memory as a bf16 MMA operand and contracts it at once:
 *         dx     [B, H, W, C]  channels-last, bf16 (dx_f32 == 0) or fp32 (dx_f32 == 1)   = G^T W     (NULL: skipped)
 *         dw     [K*D, C]      fp32                                                     = sum_b G X (NULL: skipped, with dbias)
 *         dbias  [K*D]         fp32 or NULL                                             = sum_{b,p} G
 *       dw / dbias are zeroed by the call and accumulated with fp32 atomics (the summation order over samples is not fixed).
 *       rowcoef_ws: xsup_conv_bwd_ws_floats(s) floats of scratch.  Same shape constraints as xsup_conv_head_fwd.
 *   xsup_conv_head_bwd_g (validation / diagnostics) writes d loss / d logits itself as bf16 `g [B, K*D, H*W]` plus
 *   `gbias_part [B, 4, K*D]` fp32 (sum over the first two axes = d loss / d bias). */
int xsup_integral_coef(const float* stats, const float* g_kps, float* coef_ws, const xsup_shape_t* s, void* stream);
size_t xsup_conv_bwd_ws_floats(const xsup_shape_t* s);
int xsup_conv_head_bwd(const void* x_nhwc, const void* weight, const float* bias, const float* coef_ws, float* rowcoef_ws, void* dx,
                       int32_t dx_f32, float* dw, float* dbias, const xsup_shape_t* s, int32_t C, void* stream);
int xsup_conv_head_bwd_g(const void* x_nhwc, const void* weight, const float* bias, const float* coef_ws, void* g_out,
                         float* gbias_part, const xsup_shape_t* s, int32_t C, void* stream);

/* The three pose loss terms on their own (modules/base_losses/loss_func.py:18-52), for callers that use them outside
 * the fused per-camera op: x [B,K,C] fp32 -> loss (device scalar), and the vector-Jacobian product g_x [B,K,C].
 *   XSUP_TERM_MSE   compute_supervision(keypoint, keypoint_gt, feature_shape, mode): gt [B,K,C]; feature_shape (HOST pointer
 *                   to 3 floats, or NULL) rescales x as :39-45; flag 0 = mode 'mean' (divide by B*K*C), 1 = 'sum' (by B, :51)
 *   XSUP_TERM_BONE  compute_bone_sym_loss(keypoints): needs K >= 17; mean over B*4; flag ignored
 *   XSUP_TERM_KP    compute_kp_sym_loss(keypoints, is_3D = flag): needs K >= 15; mean over B*2*C
 * sample_ws: B floats of scratch. */
enum { XSUP_TERM_MSE = 0, XSUP_TERM_BONE = 1, XSUP_TERM_KP = 2 };
int xsup_pose_term_fwd(const float* x, const float* gt, const float* feature_shape, int32_t term, int32_t flag, int32_t B,
                       int32_t K, int32_t C, float* sample_ws, float* loss, void* stream);
int xsup_pose_term_bwd(const float* x, const float* gt, const float* feature_shape, int32_t term, int32_t flag, int32_t B,
                       int32_t K, int32_t C, const float* g_loss, float* g_x, void* stream);
/* compute_supervision(..., mode='none') (loss_func.py:46-47, nn.MSELoss(reduction='none')): out [B,K,C] = (x' - gt)^2 with
 * x' the optionally feature_shape-rescaled x.  g_out == NULL: forward; else out = g_out * d out / d x (the VJP). */
int xsup_pose_sqerr(const float* x, const float* gt, const float* feature_shape, int32_t B, int32_t K, int32_t C,
                    const float* g_out, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XSUP_B200_H_ */
