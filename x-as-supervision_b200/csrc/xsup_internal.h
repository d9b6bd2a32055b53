// Host/device declarations shared by the xsup_b200 translation units (not part of the C ABI).
#pragma once
#include "xsup_common.cuh"

namespace xsup {

constexpr int kMaxFinalisers = 3;    // 20 warps = 5 per scheduler partition: the register cap stays at 96 per thread
constexpr int kFwdThreads = (kConsumerWarps + 1 + kMaxFinalisers) * 32;   // 16 consumers + producer + up to 3 finalisers (FwdParams::nfin)
constexpr int kBwdThreads = (kConsumerWarps + 1) * 32;   // 16 consumers + producer
constexpr size_t kSmemBudget = 227 * 1024;               // per-CTA opt-in maximum on sm_100
constexpr int kMaxStages = 16;
constexpr int kSchedWords = XSUP_SCHED_WORDS;                         // ints reserved after stats / coef for the work-claim counter

struct FwdParams {
    const void* logits;
    float* kps;
    float* dmap;
    int64_t* peak_idx;
    float* stats;
    int n_units, K, NH, NS, head, stats_stride;
    int nst;
    int nfin;          // finaliser warps = partial buffers (2..kMaxFinalisers), set by launch_integral_fwd
    int* counter;      // work-claim counter (zeroed by the launcher), in the caller's stats buffer
    Tiling t;
};

struct BwdParams {
    const void* logits;
    const float* coef;
    void* g_logits;
    int n_units, coef_stride;
    int nst, slot_bytes;
    int chunk;         // ring stages per claim
    int* counter;      // work-claim counter (zeroed by integral_coef_kernel), in the caller's coef buffer
    Tiling t;
};

struct CoefParams {
    const float* stats;
    const float* g_kps;
    float* coef;
    int n_units, K, D, H, W, NH, NS, head, stats_stride, coef_stride;
    int* counter;
};

cudaError_t launch_integral_fwd(FwdParams p, bool fast, int dtype, int num_sms, cudaStream_t st);
cudaError_t launch_find_peak(const float* pz, int64_t* idx, int rows, int D, int NH, cudaStream_t st);
cudaError_t launch_integral_coef(const CoefParams& p, cudaStream_t st);
cudaError_t launch_integral_bwd(BwdParams p, bool fast, int dtype, int num_sms, cudaStream_t st);

cudaError_t launch_geom(int dir, const float* in, const float* g_out, float* out, const xsup_geom_t& g, cudaStream_t st);

cudaError_t launch_reproj_loss_fwd(const float* kps, const float* target, const xsup_cam_t& cam, float* world,
                                   float* sample_terms, float* partial, const xsup_loss_cfg_t& c, cudaStream_t st);
cudaError_t launch_partial_allreduce(float* partial, int n, const xsup_xchg_t& x, cudaStream_t st);
cudaError_t launch_reproj_select(const float* kps, const float* target, const float* sample_terms, const float* partial,
                                 float* loss, int64_t* sel, const xsup_loss_cfg_t& c, cudaStream_t st);
cudaError_t launch_reproj_loss_bwd(const float* kps, const float* target, const xsup_cam_t& cam, const int64_t* sel,
                                   const float* g_loss, float* g_kps, const xsup_loss_cfg_t& c, cudaStream_t st);
cudaError_t launch_reproj_fused_fwd(const float* kps, const float* target, const xsup_cam_t& cam, float* world, float* sample_terms,
                                    float* partial, float* loss, int64_t* sel, const xsup_loss_cfg_t& c, const xsup_xchg_t& x,
                                    unsigned int* ticket, cudaStream_t st);
cudaError_t launch_reproj_fused_bwd(const float* kps, const float* target, const xsup_cam_t& cam, const int64_t* sel, const float* g_lp,
                                    const float* g_ls, const float* g_kps_in, const float* g_world, float* g_kps_out,
                                    const xsup_loss_cfg_t& c, const CoefParams& p, cudaStream_t st);

// skeleton rasteriser + mask loss (skeleton_mask.cu)
struct SkelParams {
    const float* kps;
    int kbs, kjs;              // batch / joint strides of `kps` in floats (x at +0, y at +1)
    int B, K, S, L;
    float bw;                  // body_width (model.py:31-32)
    int tiles_x, tiles, NC;    // 16x8 pixel tiles per row / per sample, CTAs per sample
    unsigned tiles_x_magic;    // floor(2^32 / tiles_x)
    int parent[XSUP_MAX_LINES], child[XSUP_MAX_LINES];
};
int skel_chunks(int S);
int draw_lines_chunks(int S);
int mask_loss_ctas(long long n);
cudaError_t launch_skeleton_mask_fwd(SkelParams p, float* recon, uint8_t* line_idx, const float* gt, const float* weight,
                                     const xsup_mask_loss_t* loss, float* loss_sums, float* ws, cudaStream_t st);
cudaError_t launch_skeleton_mask_bwd(SkelParams p, const float* recon, const uint8_t* line_idx, const float* g_recon,
                                     const float* gt, const float* weight, const xsup_mask_loss_t* loss, const float* loss_sums,
                                     const float* g_loss, float* g_kps, float* ws, cudaStream_t st);
cudaError_t launch_draw_lines_fwd(SkelParams p, float* heat, cudaStream_t st);
cudaError_t launch_draw_lines_bwd(SkelParams p, const float* heat, const float* g_heat, float* g_kps, float* ws, cudaStream_t st);
cudaError_t launch_mask_loss_fwd(const float* mask, const float* gt, const float* weight, float* filter_out,
                                 const xsup_mask_loss_t& c, float* loss_sums, float* ws, cudaStream_t st);
cudaError_t launch_mask_loss_bwd(const float* mask, const float* gt, const float* weight, const xsup_mask_loss_t& c,
                                 const float* loss_sums, const float* g_loss, float* g_mask, int num_sms, cudaStream_t st);

// eval selection, triangulation, discriminator glue (eval_disc.cu)
struct EvalParams {
    const float* kps;
    const float* joints_px;
    float img_size;
    int B, NH, K, best;
    float* kp3d;
    float* kp2d;
    uint8_t* is_trans;
    float* err2d;
    int64_t* best_idx;
    int64_t* best_2d_idx;
    float* gt_norm;
    int perm[32];
};
typedef xsup_tri_t TriParams;
struct PoseTermParams {
    const float* x;
    const float* gt;
    int B, K, C, term, use_fs, is_3d;
    float fs[3];
};
cudaError_t launch_pose_term_fwd(const PoseTermParams& p, double denom, float* sample_sums, float* loss, cudaStream_t st);
cudaError_t launch_pose_term_bwd(const PoseTermParams& p, double denom, const float* g_loss, float* g_x, cudaStream_t st);
cudaError_t launch_pose_sqerr(const PoseTermParams& p, const float* g_out, float* out, cudaStream_t st);
cudaError_t launch_eval_select(const EvalParams& p, cudaStream_t st);
cudaError_t launch_triangulate(const TriParams& p, float* world, cudaStream_t st);
cudaError_t launch_root_centre_fwd(const float* world, float* out, int N, int M, int R, int dim, cudaStream_t st);
cudaError_t launch_root_centre_bwd(const float* g_out, float* g_world, int N, int M, int R, int dim, cudaStream_t st);
cudaError_t launch_disc_min_loss_fwd(const float* logits, int B, int NH, int C, float target, float* loss, int64_t* sel, cudaStream_t st);
cudaError_t launch_disc_min_loss_bwd(const float* logits, const int64_t* sel, const float* g_loss, int B, int NH, int C, float target,
                                     float* g_logits, cudaStream_t st);

// conv-fused forward (conv_head_fwd.cu)
cudaError_t launch_conv_head_fwd(const void* x_nhwc, const void* w, const float* bias, float* logits_out, FwdParams f, int B, int C,
                                 int num_sms, cudaStream_t st);

cudaError_t launch_conv_head_fwd_tf32(const void* x_nhwc_f32, const void* w_f32, const float* bias, float* logits_out, FwdParams f, int B, int C,
                                      int num_sms, cudaStream_t st);
cudaError_t launch_conv_head_bwd_g(const void* x_nhwc, const void* w, const float* bias, const float* coef, int coef_stride, void* g_out,
                                   float* gbias_part, FwdParams f, int B, int C, int num_sms, cudaStream_t st);
cudaError_t launch_pack_nhwc_bf16(const float* x, void* y, int B, int C, int HW, cudaStream_t st);
// conv-fused backward (conv_head_bwd.cu)
int conv_bwd_rows_pad(int K, int D);
cudaError_t launch_conv_head_bwd(const void* x_nhwc, const void* w, const float* bias, const float* coef, int coef_stride, float* rowcoef_ws,
                                 void* dx, int dx_f32, float* dw, float* dbias, int B, int K, int D, int H, int W, int C, int num_sms,
                                 cudaStream_t st);

void count_launches(int n);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel instantiation, device), not on every launch
template <typename Kern>
inline cudaError_t ensure_max_smem(Kern kern, unsigned long long& done_mask) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (__atomic_load_n(&done_mask, __ATOMIC_ACQUIRE) & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget);
    if (e == cudaSuccess) __atomic_fetch_or(&done_mask, bit, __ATOMIC_RELEASE);
    return e;
}

}  // namespace xsup
