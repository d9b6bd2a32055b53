// extern "C" entry points of libxsup_b200.so (declared in include/xsup_b200.h).
// Host-side validation happens here; nothing in this file computes on the CPU.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "xsup_internal.h"

namespace xsup {

static thread_local char t_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void count_launches(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
    return code;
}

static int cuda_fail(cudaError_t e, const char* what) {
    snprintf(t_err, sizeof(t_err), "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return (int)e;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// returns true when the TMA-ring kernels can cut this volume; false -> generic kernels
static bool make_tiling(const xsup_shape_t& s, Tiling& t) {
    t = Tiling{};
    t.D = s.D; t.H = s.H; t.W = s.W;
    t.esize = s.dtype == XSUP_F32 ? 4 : 2;
    t.unit_bytes = (long long)s.D * s.H * s.W * t.esize;
    const int vec = 16 / t.esize;
    if (s.W % vec) return false;
    t.lpr = s.W / vec;
    if (!is_pow2(t.lpr) || t.lpr > 32) return false;
    t.lpr_log2 = 0;
    while ((1 << t.lpr_log2) < t.lpr) ++t.lpr_log2;
    const long long slice_bytes = (long long)s.H * s.W * t.esize;
    if (slice_bytes % 512) return false;
    const long long per = slice_bytes / 512;
    t.U = kMaxU;
    while (per % t.U) t.U >>= 1;
    t.task_bytes = 512 * t.U;
    t.parts = (int)(slice_bytes / t.task_bytes);
    t.parts_log2 = -1;
    if (is_pow2(t.parts)) { t.parts_log2 = 0; while ((1 << t.parts_log2) < t.parts) ++t.parts_log2; }
    t.rows_per_task = 32 * t.U / t.lpr;
    const long long tu = (long long)s.D * t.parts;
    if (tu > 4096) return false;
    t.tasks_per_unit = (int)tu;
    // small tasks (2 KB for 32^3 bf16) are batched per warp so that a ring stage stays ~16 KB - the producer / consumer handshake and
    // the consumers' per-iteration bookkeeping are per stage -, as long as a unit still spans at least one stage per warp group
    t.rounds = 1;
    while (t.rounds < 4 && kTasksPerStage * t.task_bytes * t.rounds * 2 <= 16384 && t.tasks_per_unit >= kTasksPerStage * kGroups * t.rounds * 2)
        t.rounds *= 2;
    const int per_stage = kTasksPerStage * t.rounds;
    t.stages_per_unit = (t.tasks_per_unit + per_stage - 1) / per_stage;
    t.stage_bytes = per_stage * t.task_bytes;
    return true;
}

static int check_shape(const xsup_shape_t* s) {
    if (!s) return fail(XSUP_E_NULL, "shape is NULL");
    if (s->dtype != XSUP_F32 && s->dtype != XSUP_BF16) return fail(XSUP_E_DTYPE, "dtype must be XSUP_F32 or XSUP_BF16 (got %d)", s->dtype);
    if (s->head != XSUP_HEAD_MULTI && s->head != XSUP_HEAD_SINGLE) return fail(XSUP_E_SHAPE, "unknown head mode %d", s->head);
    if (s->B < 0 || s->K <= 0 || s->D <= 0 || s->H <= 0 || s->W <= 0) return fail(XSUP_E_SHAPE, "non-positive dimension");
    if (s->D != s->W)
        return fail(XSUP_E_SHAPE, "depth_dim (%d) must equal width (%d): the reference multiplies the depth marginal by arange(W)", s->D, s->W);
    if (s->D > kMaxD) return fail(XSUP_E_SHAPE, "depth_dim %d > %d", s->D, kMaxD);
    if ((long long)s->B * s->K > 0x7fffffffLL / 64) return fail(XSUP_E_SHAPE, "too many (b,k) units");
    if (s->head == XSUP_HEAD_MULTI) {
        if (s->NH < 1 || s->NH > s->D - 2) return fail(XSUP_E_SHAPE, "num_hypo %d must be in [1, D-2=%d] (topk over the interior bins)", s->NH, s->D - 2);
        if (s->NS < 1 || !(s->NS & 1)) return fail(XSUP_E_SHAPE, "neighbor_size %d must be odd and positive", s->NS);
    } else if (s->NH != 1) {
        return fail(XSUP_E_SHAPE, "single-hypothesis head needs NH == 1");
    }
    return XSUP_OK;
}

static int device_info(int& num_sms) {
    static thread_local int cached_dev = -1, cached_sms = 0, cached_major = 0;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    if (dev != cached_dev) {
        cudaDeviceProp prop;
        e = cudaGetDeviceProperties(&prop, dev);
        if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties");
        cached_dev = dev; cached_sms = prop.multiProcessorCount; cached_major = prop.major;
    }
    if (cached_major != 10) return fail(XSUP_E_DEVICE, "xsup_b200 kernels are built for sm_100a only (device is sm_%d*)", cached_major * 10);
    num_sms = cached_sms;
    return XSUP_OK;
}

static size_t stats_stride(const xsup_shape_t& s) { return (size_t)((4 + s.D + 3 * s.NH + 3) / 4 * 4); }
static size_t coef_stride(const xsup_shape_t& s) { return (size_t)((8 + s.D + 3) / 4 * 4); }

}  // namespace xsup

using namespace xsup;

extern "C" {

int xsup_abi_version(void) { return XSUP_ABI_VERSION; }
const char* xsup_last_error(void) { return t_err; }
uint64_t xsup_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

size_t xsup_stats_stride(const xsup_shape_t* s) { return s ? stats_stride(*s) : 0; }
size_t xsup_coef_stride(const xsup_shape_t* s) { return s ? coef_stride(*s) : 0; }
size_t xsup_stats_floats(const xsup_shape_t* s) { return s ? (size_t)s->B * s->K * stats_stride(*s) + kSchedWords : 0; }
size_t xsup_coef_floats(const xsup_shape_t* s) { return s ? (size_t)s->B * s->K * coef_stride(*s) + kSchedWords : 0; }

int xsup_integral_fwd(const void* logits, float* kps, float* depth_prob_map, int64_t* peak_idx, float* stats,
                      const xsup_shape_t* s, void* stream) {
    if (int rc = check_shape(s)) return rc;
    if (s->B == 0) return XSUP_OK;                      // empty batch: nothing to do (pointers may be NULL)
    if (!logits || !kps || !depth_prob_map || !stats) return fail(XSUP_E_NULL, "xsup_integral_fwd: NULL pointer");
    if (!aligned16(logits) || !aligned16(stats)) return fail(XSUP_E_ALIGN, "xsup_integral_fwd: logits/stats must be 16-byte aligned");
    int sms = 0;
    if (int rc = device_info(sms)) return rc;
    FwdParams p{};
    p.logits = logits; p.kps = kps; p.dmap = depth_prob_map; p.peak_idx = peak_idx; p.stats = stats;
    p.n_units = s->B * s->K; p.K = s->K; p.NH = s->NH; p.NS = s->NS; p.head = s->head;
    p.stats_stride = (int)stats_stride(*s);
    p.counter = reinterpret_cast<int*>(stats + (size_t)p.n_units * p.stats_stride);
    const bool fast = make_tiling(*s, p.t);
    cudaError_t e = cudaSuccess;
    // all XSUP_SCHED_WORDS: word 0 is the work-claim counter, word 1 the ticket of xsup_reproj_fused_fwd
    e = cudaMemsetAsync(p.counter, 0, kSchedWords * sizeof(int), (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "xsup_integral_fwd counter reset");
    e = launch_integral_fwd(p, fast, s->dtype, sms, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "xsup_integral_fwd launch");
    count_launches(1);
    return XSUP_OK;
}

int xsup_integral_bwd(const void* logits, const float* stats, const float* g_kps, void* g_logits, float* coef_ws,
                      const xsup_shape_t* s, void* stream) {
    if (int rc = check_shape(s)) return rc;
    if (s->B == 0) return XSUP_OK;
    if (!logits || !stats || !g_kps || !g_logits || !coef_ws) return fail(XSUP_E_NULL, "xsup_integral_bwd: NULL pointer");
    if (!aligned16(logits) || !aligned16(g_logits) || !aligned16(coef_ws))
        return fail(XSUP_E_ALIGN, "xsup_integral_bwd: logits/g_logits/coef_ws must be 16-byte aligned");
    int sms = 0;
    if (int rc = device_info(sms)) return rc;
    CoefParams c{};
    c.stats = stats; c.g_kps = g_kps; c.coef = coef_ws;
    c.n_units = s->B * s->K; c.K = s->K; c.D = s->D; c.H = s->H; c.W = s->W; c.NH = s->NH; c.NS = s->NS; c.head = s->head;
    c.stats_stride = (int)stats_stride(*s); c.coef_stride = (int)coef_stride(*s);
    c.counter = reinterpret_cast<int*>(coef_ws + (size_t)c.n_units * c.coef_stride);
    cudaError_t e = launch_integral_coef(c, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "xsup_integral_bwd coefficient launch");
    BwdParams p{};
    p.logits = logits; p.coef = coef_ws; p.g_logits = g_logits;
    p.n_units = c.n_units; p.coef_stride = c.coef_stride; p.counter = c.counter;
    const bool fast = make_tiling(*s, p.t);
    e = launch_integral_bwd(p, fast, s->dtype, sms, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "xsup_integral_bwd launch");
    count_launches(2);
    return XSUP_OK;
}

int xsup_integral_bwd_apply(const void* logits, const float* coef_ws, void* g_logits, const xsup_shape_t* s, void* stream) {
    if (int rc = check_shape(s)) return rc;
    if (s->B == 0) return XSUP_OK;
    if (!logits || !g_logits || !coef_ws) return fail(XSUP_E_NULL, "xsup_integral_bwd_apply: NULL pointer");
    if (!aligned16(logits) || !aligned16(g_logits) || !aligned16(coef_ws))
        return fail(XSUP_E_ALIGN, "xsup_integral_bwd_apply: logits/g_logits/coef_ws must be 16-byte aligned");
    int sms = 0;
    if (int rc = device_info(sms)) return rc;
    BwdParams p{};
    p.logits = logits; p.coef = coef_ws; p.g_logits = g_logits;
    p.n_units = s->B * s->K; p.coef_stride = (int)coef_stride(*s);
    p.counter = reinterpret_cast<int*>(const_cast<float*>(coef_ws) + (size_t)p.n_units * p.coef_stride);   // zeroed by the coefficient kernel
    const bool fast = make_tiling(*s, p.t);
    cudaError_t e = launch_integral_bwd(p, fast, s->dtype, sms, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "xsup_integral_bwd_apply launch");
    count_launches(1);
    return XSUP_OK;
}

int xsup_find_peak(const float* pz, int64_t* idx, int32_t rows, int32_t D, int32_t NH, void* stream) {
    if (!pz || !idx) return fail(XSUP_E_NULL, "xsup_find_peak: NULL pointer");
    if (rows < 0 || D < 3 || D > kMaxD || NH < 1 || NH > D - 2) return fail(XSUP_E_SHAPE, "xsup_find_peak: need 3 <= D <= %d and 1 <= NH <= D-2", kMaxD);
    if (rows == 0) return XSUP_OK;
    cudaError_t e = launch_find_peak(pz, idx, rows, D, NH, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "xsup_find_peak launch");
    count_launches(1);
    return XSUP_OK;
}

static int check_cam(const xsup_cam_t* cam, const char* who) {
    if (!cam) return fail(XSUP_E_NULL, "%s: cam is NULL", who);
    if (!cam->trans_image || !cam->pelvis || !cam->k_mat || !cam->trans_world || !cam->rot_world)
        return fail(XSUP_E_NULL, "%s: a camera tensor is NULL", who);
    return XSUP_OK;
}

// ---- geometry: one validated description, four entry points + the three round-1 wrappers
static int check_geom(const xsup_geom_t* g, const char* who) {
    if (!g) return fail(XSUP_E_NULL, "%s: geometry description is NULL", who);
    if (g->B < 0 || g->J <= 0) return fail(XSUP_E_SHAPE, "%s: bad sizes (B=%d J=%d)", who, g->B, g->J);
    if (g->flags & ~(XSUP_GEOM_NORM | XSUP_GEOM_MONO | XSUP_GEOM_PATCH_STAGE | XSUP_GEOM_CAMERA_STAGE)) return fail(XSUP_E_SHAPE, "%s: unknown flag bits 0x%x", who, g->flags);
    if (g->flags & XSUP_GEOM_PATCH_STAGE) {
        if (!g->trans_image || !g->pelvis) return fail(XSUP_E_NULL, "%s: the patch stage needs trans_image and pelvis", who);
        if (!(g->depth_scale > 0.0f)) return fail(XSUP_E_SHAPE, "%s: depth_scale must be positive", who);
        if ((g->flags & XSUP_GEOM_NORM) && (g->img_d <= 1 || g->img_h <= 1 || g->img_w <= 1)) return fail(XSUP_E_SHAPE, "%s: image extents must exceed 1", who);
    }
    if ((g->flags & XSUP_GEOM_CAMERA_STAGE) && !(g->flags & XSUP_GEOM_MONO)) {
        if (!g->fx || !g->fy || !g->cx || !g->cy || !g->trans_world || !g->rot_world) return fail(XSUP_E_NULL, "%s: the camera stage needs fx, fy, cx, cy, trans_world, rot_world", who);
        if (g->intr_stride < 1) return fail(XSUP_E_SHAPE, "%s: intr_stride must be >= 1", who);
    }
    return XSUP_OK;
}

static int run_geom(int dir, const float* in, const float* g_out, float* out, bool vjp, const xsup_geom_t* g, void* stream, const char* who) {
    if (int rc = check_geom(g, who)) return rc;
    if (g->B == 0) return XSUP_OK;
    if (!in || !out || (vjp && !g_out)) return fail(XSUP_E_NULL, "%s: NULL pointer", who);
    cudaError_t e = launch_geom(dir, in, vjp ? g_out : nullptr, out, *g, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, who);
    count_launches(1);
    return XSUP_OK;
}

int xsup_geom_patch_to_world(const float* in, float* out, const xsup_geom_t* g, void* stream) {
    return run_geom(0, in, nullptr, out, false, g, stream, "xsup_geom_patch_to_world");
}
int xsup_geom_patch_to_world_vjp(const float* in, const float* g_out, float* g_in, const xsup_geom_t* g, void* stream) {
    return run_geom(0, in, g_out, g_in, true, g, stream, "xsup_geom_patch_to_world_vjp");
}
int xsup_geom_world_to_patch(const float* in, float* out, const xsup_geom_t* g, void* stream) {
    return run_geom(1, in, nullptr, out, false, g, stream, "xsup_geom_world_to_patch");
}
int xsup_geom_world_to_patch_vjp(const float* in, const float* g_out, float* g_in, const xsup_geom_t* g, void* stream) {
    return run_geom(1, in, g_out, g_in, true, g, stream, "xsup_geom_world_to_patch_vjp");
}

// the composites with the data loader's tensors (k_mat instead of fx..cy; depth extent = image width, util.py:137-138)
static int geom_from_cam(xsup_geom_t& g, const xsup_cam_t* cam, int B, int J, int img_h, int img_w, float rect_width, int flags, const char* who) {
    if (int rc = check_cam(cam, who)) return rc;
    if (B < 0 || J <= 0 || img_h <= 1 || img_w <= 1) return fail(XSUP_E_SHAPE, "%s: bad sizes", who);
    g = xsup_geom_t{};
    g.B = B; g.J = J; g.img_d = img_w; g.img_h = img_h; g.img_w = img_w;
    g.depth_scale = 1.0f / (float)img_w * rect_width;
    g.flags = (flags & (XSUP_GEOM_NORM | XSUP_GEOM_MONO | XSUP_GEOM_PATCH_STAGE)) | XSUP_GEOM_CAMERA_STAGE;
    g.intr_stride = 9;
    g.trans_image = cam->trans_image; g.pelvis = cam->pelvis;
    g.fx = cam->k_mat; g.fy = cam->k_mat + 4; g.cx = cam->k_mat + 2; g.cy = cam->k_mat + 5;
    g.trans_world = cam->trans_world; g.rot_world = cam->rot_world;
    return XSUP_OK;
}

int xsup_patch_to_world_fwd(const float* kps, const xsup_cam_t* cam, float* world, int32_t B, int32_t J, int32_t img_h,
                            int32_t img_w, float rect_width, int32_t flags, void* stream) {
    xsup_geom_t g;
    if (int rc = geom_from_cam(g, cam, B, J, img_h, img_w, rect_width, flags, "xsup_patch_to_world_fwd")) return rc;
    return run_geom(0, kps, nullptr, world, false, &g, stream, "xsup_patch_to_world_fwd");
}

int xsup_patch_to_world_bwd(const float* kps, const float* g_world, const xsup_cam_t* cam, float* g_kps, int32_t B, int32_t J,
                            int32_t img_h, int32_t img_w, float rect_width, int32_t flags, void* stream) {
    xsup_geom_t g;
    if (int rc = geom_from_cam(g, cam, B, J, img_h, img_w, rect_width, flags, "xsup_patch_to_world_bwd")) return rc;
    return run_geom(0, kps, g_world, g_kps, true, &g, stream, "xsup_patch_to_world_bwd");
}

int xsup_world_to_patch_fwd(const float* world, const xsup_cam_t* cam, float* kps, int32_t B, int32_t J, int32_t img_h,
                            int32_t img_w, float rect_width, int32_t flags, void* stream) {
    xsup_geom_t g;
    if (int rc = geom_from_cam(g, cam, B, J, img_h, img_w, rect_width, flags | XSUP_GEOM_PATCH_STAGE, "xsup_world_to_patch_fwd")) return rc;
    return run_geom(1, world, nullptr, kps, false, &g, stream, "xsup_world_to_patch_fwd");
}

static int check_cfg(const xsup_loss_cfg_t* c, const char* who) {
    if (!c) return fail(XSUP_E_NULL, "%s: cfg is NULL", who);
    if (c->B < 0 || c->K <= 0 || c->NH <= 0) return fail(XSUP_E_SHAPE, "%s: bad sizes", who);
    if (c->K > 32) return fail(XSUP_E_SHAPE, "%s: num_kp %d > 32 (one joint per lane)", who, c->K);
    if (c->use_sym && c->K < 17) return fail(XSUP_E_SHAPE, "%s: symmetry terms index joints up to 16 (loss_func.py:20), num_kp is %d", who, c->K);
    if (c->use_sym && c->reduction == XSUP_REDUCE_JOINT) return fail(XSUP_E_SHAPE, "%s: symmetry terms are undefined per joint", who);
    if (c->reduction < XSUP_REDUCE_BATCH || c->reduction > XSUP_REDUCE_JOINT) return fail(XSUP_E_SHAPE, "%s: unknown reduction %d", who, c->reduction);
    if (c->batch_total < c->B || c->img_h <= 1 || c->img_w <= 1) return fail(XSUP_E_SHAPE, "%s: batch_total < B or bad image size", who);
    return XSUP_OK;
}

int xsup_reproj_loss_fwd(const float* kps, const float* target, const xsup_cam_t* cam, float* world, float* sample_terms,
                         float* partial, const xsup_loss_cfg_t* cfg, void* stream) {
    if (int rc = check_cfg(cfg, "xsup_reproj_loss_fwd")) return rc;
    if (int rc = check_cam(cam, "xsup_reproj_loss_fwd")) return rc;
    if (!kps || !target || !world || !sample_terms || !partial) return fail(XSUP_E_NULL, "xsup_reproj_loss_fwd: NULL pointer");
    if (cfg->B == 0) return fail(XSUP_E_SHAPE, "xsup_reproj_loss_fwd: empty batch");
    if (!aligned16(sample_terms)) return fail(XSUP_E_ALIGN, "xsup_reproj_loss_fwd: sample_terms must be 16-byte aligned");
    cudaError_t e = launch_reproj_loss_fwd(kps, target, *cam, world, sample_terms, partial, *cfg, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "xsup_reproj_loss_fwd launch");
    count_launches(2);
    return XSUP_OK;
}

static int check_xchg(const xsup_xchg_t* x, int n, const char* who);
size_t xsup_xchg_floats(int32_t world) { return world > 0 ? (size_t)2 * world * XSUP_XCHG_SLOT : 0; }

int xsup_partial_allreduce(float* partial, int32_t n, const xsup_xchg_t* x, void* stream) {
    if (!partial || !x) return fail(XSUP_E_NULL, "xsup_partial_allreduce: NULL pointer");
    if (int rc = check_xchg(x, n, "xsup_partial_allreduce")) return rc;
    cudaError_t e = launch_partial_allreduce(partial, n, *x, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "xsup_partial_allreduce launch");
    count_launches(1);
    return XSUP_OK;
}

int xsup_reproj_select(const float* kps, const float* target, const float* sample_terms, const float* partial, float* loss,
                       int64_t* sel, const xsup_loss_cfg_t* cfg, void* stream) {
    if (int rc = check_cfg(cfg, "xsup_reproj_select")) return rc;
    if (!kps || !target || !sample_terms || !partial || !loss || !sel) return fail(XSUP_E_NULL, "xsup_reproj_select: NULL pointer");
    cudaError_t e = launch_reproj_select(kps, target, sample_terms, partial, loss, sel, *cfg, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "xsup_reproj_select launch");
    count_launches(1);
    return XSUP_OK;
}

int xsup_reproj_loss_bwd(const float* kps, const float* target, const xsup_cam_t* cam, const int64_t* sel, const float* g_loss,
                         float* g_kps, const xsup_loss_cfg_t* cfg, void* stream) {
    if (int rc = check_cfg(cfg, "xsup_reproj_loss_bwd")) return rc;
    if (int rc = check_cam(cam, "xsup_reproj_loss_bwd")) return rc;
    if (!kps || !target || !sel || !g_loss || !g_kps) return fail(XSUP_E_NULL, "xsup_reproj_loss_bwd: NULL pointer");
    if (cfg->B == 0) return XSUP_OK;
    cudaError_t e = launch_reproj_loss_bwd(kps, target, *cam, sel, g_loss, g_kps, *cfg, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "xsup_reproj_loss_bwd launch");
    count_launches(1);
    return XSUP_OK;
}

static int check_xchg(const xsup_xchg_t* x, int n, const char* who) {
    if (!x->peer_bufs) return fail(XSUP_E_NULL, "%s: peer_bufs is NULL", who);
    if (n < 1 || n > XSUP_XCHG_SLOT - 1 || x->world < 1 || x->world > 64 || x->rank < 0 || x->rank >= x->world || (x->step == 0 && !x->seq))
        return fail(XSUP_E_SHAPE, "%s: need 1 <= n <= %d, 1 <= world <= 64, 0 <= rank < world, step >= 1 (or a device sequence counter)", who, XSUP_XCHG_SLOT - 1);
    return XSUP_OK;
}

int xsup_reproj_fused_fwd(const float* kps, const float* target, const xsup_cam_t* cam, float* world, float* sample_terms,
                          float* partial, float* loss, int64_t* sel, const xsup_loss_cfg_t* cfg, const xsup_xchg_t* xchg,
                          uint32_t* ticket, void* stream) {
    if (int rc = check_cfg(cfg, "xsup_reproj_fused_fwd")) return rc;
    if (int rc = check_cam(cam, "xsup_reproj_fused_fwd")) return rc;
    if (!kps || !target || !world || !sample_terms || !partial || !loss || !sel || !ticket) return fail(XSUP_E_NULL, "xsup_reproj_fused_fwd: NULL pointer");
    if (!aligned16(sample_terms)) return fail(XSUP_E_ALIGN, "xsup_reproj_fused_fwd: sample_terms must be 16-byte aligned");
    if (cfg->B == 0) return fail(XSUP_E_SHAPE, "xsup_reproj_fused_fwd: empty batch");
    xsup_xchg_t x{};
    if (xchg) {
        if (int rc = check_xchg(xchg, XSUP_LOSS_TERMS * cfg->NH, "xsup_reproj_fused_fwd")) return rc;
        x = *xchg;
    }
    cudaError_t e = launch_reproj_fused_fwd(kps, target, *cam, world, sample_terms, partial, loss, sel, *cfg, x, ticket, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "xsup_reproj_fused_fwd launch");
    count_launches(1);
    return XSUP_OK;
}

int xsup_reproj_fused_bwd(const float* kps, const float* target, const xsup_cam_t* cam, const int64_t* sel, const float* g_lp,
                          const float* g_ls, const float* g_kps_in, const float* g_world, const float* stats, float* coef_ws,
                          float* g_kps_out, const xsup_loss_cfg_t* cfg, const xsup_shape_t* s, void* stream) {
    if (int rc = check_cfg(cfg, "xsup_reproj_fused_bwd")) return rc;
    if (int rc = check_cam(cam, "xsup_reproj_fused_bwd")) return rc;
    if (int rc = check_shape(s)) return rc;
    if (s->head != XSUP_HEAD_MULTI || s->B != cfg->B || s->K != cfg->K || s->NH != cfg->NH)
        return fail(XSUP_E_SHAPE, "xsup_reproj_fused_bwd: shape and loss cfg disagree (B, K, NH) or the head is not the multi-hypothesis one");
    if (!kps || !target || !sel || !stats || !coef_ws) return fail(XSUP_E_NULL, "xsup_reproj_fused_bwd: NULL pointer");
    if (cfg->B == 0) return XSUP_OK;
    CoefParams c{};
    c.stats = stats; c.g_kps = nullptr; c.coef = coef_ws;
    c.n_units = s->B * s->K; c.K = s->K; c.D = s->D; c.H = s->H; c.W = s->W; c.NH = s->NH; c.NS = s->NS; c.head = s->head;
    c.stats_stride = (int)stats_stride(*s); c.coef_stride = (int)coef_stride(*s);
    c.counter = reinterpret_cast<int*>(coef_ws + (size_t)c.n_units * c.coef_stride);
    cudaError_t e = launch_reproj_fused_bwd(kps, target, *cam, sel, g_lp, g_ls, g_kps_in, g_world, g_kps_out, *cfg, c, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "xsup_reproj_fused_bwd launch");
    count_launches(1);
    return XSUP_OK;
}

// ------------------------------------------------------------------------------------------------ skeleton rasteriser + mask loss
static int skel_params(SkelParams& p, const float* kps, const xsup_skel_t* s, const char* who) {
    if (!s) return fail(XSUP_E_NULL, "%s: skeleton description is NULL", who);
    if (s->B < 0 || s->K <= 0 || s->S < 4) return fail(XSUP_E_SHAPE, "%s: bad sizes (B=%d K=%d S=%d)", who, s->B, s->K, s->S);
    if (s->S % 4 || s->S > 8192) return fail(XSUP_E_SHAPE, "%s: image_size %d must be a multiple of 4 (128-bit rows) and <= 8192", who, s->S);
    if (s->L < 1 || s->L > XSUP_MAX_LINES) return fail(XSUP_E_SHAPE, "%s: %d lines, need 1..%d", who, s->L, XSUP_MAX_LINES);
    if (!(s->body_width > 0.0f)) return fail(XSUP_E_SHAPE, "%s: body_width must be positive", who);
    if (s->kp_joint_stride < 2 || s->kp_batch_stride < 0) return fail(XSUP_E_SHAPE, "%s: bad keypoint strides", who);
    for (int l = 0; l < s->L; ++l)
        if (s->parent[l] < 0 || s->parent[l] >= s->K || s->child[l] < 0 || s->child[l] >= s->K)
            return fail(XSUP_E_SHAPE, "%s: line %d joins joints (%d,%d) outside [0,%d)", who, l, s->parent[l], s->child[l], s->K);
    if (s->B > 0 && !kps) return fail(XSUP_E_NULL, "%s: kps is NULL", who);
    p = SkelParams{};
    p.kps = kps; p.kbs = s->kp_batch_stride; p.kjs = s->kp_joint_stride;
    p.B = s->B; p.K = s->K; p.S = s->S; p.L = s->L; p.bw = s->body_width;
    for (int l = 0; l < s->L; ++l) { p.parent[l] = s->parent[l]; p.child[l] = s->child[l]; }
    return XSUP_OK;
}

static int check_mask_cfg(const xsup_mask_loss_t* c, const char* who) {
    if (!c) return fail(XSUP_E_NULL, "%s: loss cfg is NULL", who);
    if (c->n <= 0) return fail(XSUP_E_SHAPE, "%s: empty mask", who);
    if (c->mode < XSUP_MASK_MSE || c->mode > XSUP_MASK_WEIGHTED) return fail(XSUP_E_SHAPE, "%s: unknown loss mode %d", who, c->mode);
    return XSUP_OK;
}

size_t xsup_skel_ws_floats(const xsup_skel_t* s) {
    if (!s || s->S < 4 || s->B < 0) return 0;
    return (size_t)s->B * skel_chunks(s->S) * (XSUP_MAX_LINES * 4 + 4);
}
size_t xsup_draw_lines_ws_floats(const xsup_skel_t* s) {
    if (!s || s->S < 4 || s->B < 0) return 0;
    return (size_t)s->B * draw_lines_chunks(s->S) * XSUP_MAX_LINES * 4;
}
size_t xsup_mask_loss_ws_floats(int64_t n) { return n > 0 ? (size_t)mask_loss_ctas(n) * 4 : 0; }

int xsup_draw_lines_fwd(const float* kps, const xsup_skel_t* s, float* heat, void* stream) {
    SkelParams p;
    if (int rc = skel_params(p, kps, s, "xsup_draw_lines_fwd")) return rc;
    if (p.B == 0) return XSUP_OK;
    if (!heat) return fail(XSUP_E_NULL, "xsup_draw_lines_fwd: NULL pointer");
    if (!aligned16(heat)) return fail(XSUP_E_ALIGN, "xsup_draw_lines_fwd: heat must be 16-byte aligned");
    if (p.B > 65535) return fail(XSUP_E_SHAPE, "xsup_draw_lines_fwd: B > 65535");
    cudaError_t e = launch_draw_lines_fwd(p, heat, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "xsup_draw_lines_fwd launch");
    count_launches(1);
    return XSUP_OK;
}

int xsup_draw_lines_bwd(const float* kps, const xsup_skel_t* s, const float* heat, const float* g_heat, float* g_kps, float* ws,
                        void* stream) {
    SkelParams p;
    if (int rc = skel_params(p, kps, s, "xsup_draw_lines_bwd")) return rc;
    if (p.B == 0) return XSUP_OK;
    if (!heat || !g_heat || !g_kps || !ws) return fail(XSUP_E_NULL, "xsup_draw_lines_bwd: NULL pointer");
    if (!aligned16(heat) || !aligned16(g_heat) || !aligned16(ws)) return fail(XSUP_E_ALIGN, "xsup_draw_lines_bwd: heat/g_heat/ws must be 16-byte aligned");
    if (p.B > 65535) return fail(XSUP_E_SHAPE, "xsup_draw_lines_bwd: B > 65535");
    cudaError_t e = launch_draw_lines_bwd(p, heat, g_heat, g_kps, ws, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "xsup_draw_lines_bwd launch");
    count_launches(2);
    return XSUP_OK;
}

int xsup_skeleton_mask_fwd(const float* kps, const xsup_skel_t* s, float* recon, uint8_t* line_idx, const float* gt,
                           const float* weight, const xsup_mask_loss_t* loss, float* loss_sums, float* ws, void* stream) {
    SkelParams p;
    if (int rc = skel_params(p, kps, s, "xsup_skeleton_mask_fwd")) return rc;
    if (loss) {
        if (int rc = check_mask_cfg(loss, "xsup_skeleton_mask_fwd")) return rc;
        if (loss->n != (int64_t)p.B * p.S * p.S) return fail(XSUP_E_SHAPE, "xsup_skeleton_mask_fwd: loss n must be B*S*S");
        if (!gt || !loss_sums || !ws) return fail(XSUP_E_NULL, "xsup_skeleton_mask_fwd: gt/loss_sums/ws is NULL");
        if (loss->mode == XSUP_MASK_WEIGHTED && !weight) return fail(XSUP_E_NULL, "xsup_skeleton_mask_fwd: weighted mode without a weight map");
        if (!aligned16(gt) || !aligned16(weight) || !aligned16(ws)) return fail(XSUP_E_ALIGN, "xsup_skeleton_mask_fwd: gt/weight/ws must be 16-byte aligned");
    }
    if (p.B == 0) return XSUP_OK;
    if (!recon || !line_idx) return fail(XSUP_E_NULL, "xsup_skeleton_mask_fwd: NULL pointer");
    if (!aligned16(recon) || (reinterpret_cast<uintptr_t>(line_idx) & 3u)) return fail(XSUP_E_ALIGN, "xsup_skeleton_mask_fwd: recon (16 B) / line_idx (4 B) misaligned");
    if (p.B > 65535) return fail(XSUP_E_SHAPE, "xsup_skeleton_mask_fwd: B > 65535");
    cudaError_t e = launch_skeleton_mask_fwd(p, recon, line_idx, gt, loss && loss->mode == XSUP_MASK_WEIGHTED ? weight : nullptr, loss,
                                             loss_sums, ws, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "xsup_skeleton_mask_fwd launch");
    count_launches(loss ? 2 : 1);
    return XSUP_OK;
}

int xsup_skeleton_mask_bwd(const float* kps, const xsup_skel_t* s, const float* recon, const uint8_t* line_idx, const float* g_recon,
                           const float* gt, const float* weight, const xsup_mask_loss_t* loss, const float* loss_sums,
                           const float* g_loss, float* g_kps, float* ws, void* stream) {
    SkelParams p;
    if (int rc = skel_params(p, kps, s, "xsup_skeleton_mask_bwd")) return rc;
    if (loss) {
        if (int rc = check_mask_cfg(loss, "xsup_skeleton_mask_bwd")) return rc;
        if (loss->n != (int64_t)p.B * p.S * p.S) return fail(XSUP_E_SHAPE, "xsup_skeleton_mask_bwd: loss n must be B*S*S");
        if (!gt || !loss_sums || !g_loss) return fail(XSUP_E_NULL, "xsup_skeleton_mask_bwd: gt/loss_sums/g_loss is NULL");
        if (loss->mode == XSUP_MASK_WEIGHTED && !weight) return fail(XSUP_E_NULL, "xsup_skeleton_mask_bwd: weighted mode without a weight map");
        if (!aligned16(gt) || !aligned16(weight)) return fail(XSUP_E_ALIGN, "xsup_skeleton_mask_bwd: gt/weight must be 16-byte aligned");
    }
    if (p.B == 0) return XSUP_OK;
    if (!recon || !line_idx || !g_kps || !ws) return fail(XSUP_E_NULL, "xsup_skeleton_mask_bwd: NULL pointer");
    if (!aligned16(recon) || !aligned16(g_recon) || !aligned16(ws) || (reinterpret_cast<uintptr_t>(line_idx) & 3u))
        return fail(XSUP_E_ALIGN, "xsup_skeleton_mask_bwd: recon/g_recon/ws (16 B) / line_idx (4 B) misaligned");
    if (p.B > 65535) return fail(XSUP_E_SHAPE, "xsup_skeleton_mask_bwd: B > 65535");
    cudaError_t e = launch_skeleton_mask_bwd(p, recon, line_idx, g_recon, gt, loss && loss->mode == XSUP_MASK_WEIGHTED ? weight : nullptr,
                                             loss, loss_sums, g_loss, g_kps, ws, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "xsup_skeleton_mask_bwd launch");
    count_launches(2);
    return XSUP_OK;
}

int xsup_mask_loss_fwd(const float* mask, const float* gt, const float* weight, float* filter_out, const xsup_mask_loss_t* cfg,
                       float* loss_sums, float* ws, void* stream) {
    if (int rc = check_mask_cfg(cfg, "xsup_mask_loss_fwd")) return rc;
    if (!mask || !gt || !loss_sums || !ws) return fail(XSUP_E_NULL, "xsup_mask_loss_fwd: NULL pointer");
    if (cfg->mode == XSUP_MASK_WEIGHTED && !weight) return fail(XSUP_E_NULL, "xsup_mask_loss_fwd: weighted mode without a weight map");
    if (!aligned16(mask) || !aligned16(gt) || !aligned16(weight) || !aligned16(filter_out))
        return fail(XSUP_E_ALIGN, "xsup_mask_loss_fwd: tensors must be 16-byte aligned");
    cudaError_t e = launch_mask_loss_fwd(mask, gt, cfg->mode == XSUP_MASK_WEIGHTED ? weight : nullptr, filter_out, *cfg, loss_sums, ws,
                                         (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "xsup_mask_loss_fwd launch");
    count_launches(2);
    return XSUP_OK;
}

int xsup_mask_loss_bwd(const float* mask, const float* gt, const float* weight, const xsup_mask_loss_t* cfg, const float* loss_sums,
                       const float* g_loss, float* g_mask, void* stream) {
    if (int rc = check_mask_cfg(cfg, "xsup_mask_loss_bwd")) return rc;
    if (!mask || !gt || !loss_sums || !g_loss || !g_mask) return fail(XSUP_E_NULL, "xsup_mask_loss_bwd: NULL pointer");
    if (cfg->mode == XSUP_MASK_WEIGHTED && !weight) return fail(XSUP_E_NULL, "xsup_mask_loss_bwd: weighted mode without a weight map");
    int sms = 0;
    if (int rc = device_info(sms)) return rc;
    cudaError_t e = launch_mask_loss_bwd(mask, gt, cfg->mode == XSUP_MASK_WEIGHTED ? weight : nullptr, *cfg, loss_sums, g_loss, g_mask, sms,
                                         (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "xsup_mask_loss_bwd launch");
    count_launches(1);
    return XSUP_OK;
}

// ------------------------------------------------------------------------------------------------ eval selection, triangulation, discriminator glue
int xsup_eval_select(const float* kps, const float* joints_px, const xsup_eval_t* cfg, float* kp3d, float* kp2d, uint8_t* is_trans,
                     float* err2d, int64_t* best_idx, int64_t* best_2d_idx, float* gt_norm, void* stream) {
    if (!cfg) return fail(XSUP_E_NULL, "xsup_eval_select: cfg is NULL");
    if (cfg->B < 0 || cfg->NH < 1 || cfg->K < 1 || cfg->K > 32) return fail(XSUP_E_SHAPE, "xsup_eval_select: need B >= 0, NH >= 1, 1 <= K <= 32 (one joint per lane)");
    if (!(cfg->img_size > 1.0f) && cfg->img_size != 0.0f) return fail(XSUP_E_SHAPE, "xsup_eval_select: img_size must exceed 1 (or be 0: ground truth already normalised)");
    for (int k = 0; k < cfg->K; ++k)
        if (cfg->perm[k] < 0 || cfg->perm[k] >= cfg->K) return fail(XSUP_E_SHAPE, "xsup_eval_select: perm[%d] = %d outside [0,%d)", k, cfg->perm[k], cfg->K);
    if (cfg->B == 0) return XSUP_OK;
    if (!kps || !joints_px || !kp3d || !kp2d) return fail(XSUP_E_NULL, "xsup_eval_select: NULL pointer");
    EvalParams p{};
    p.kps = kps; p.joints_px = joints_px; p.img_size = cfg->img_size; p.B = cfg->B; p.NH = cfg->NH; p.K = cfg->K; p.best = cfg->best;
    p.kp3d = kp3d; p.kp2d = kp2d; p.is_trans = is_trans; p.err2d = err2d; p.best_idx = best_idx; p.best_2d_idx = best_2d_idx; p.gt_norm = gt_norm;
    for (int k = 0; k < 32; ++k) p.perm[k] = k < cfg->K ? cfg->perm[k] : k;
    cudaError_t e = launch_eval_select(p, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "xsup_eval_select launch");
    count_launches(1);
    return XSUP_OK;
}

int xsup_triangulate(const xsup_tri_t* t, float* world, void* stream) {
    if (!t) return fail(XSUP_E_NULL, "xsup_triangulate: description is NULL");
    if (t->V < 2 || t->V > XSUP_MAX_VIEWS) return fail(XSUP_E_SHAPE, "xsup_triangulate: %d views, need 2..%d", t->V, XSUP_MAX_VIEWS);
    if (t->B < 0 || t->K < 1 || t->img_h <= 1 || t->img_w <= 1) return fail(XSUP_E_SHAPE, "xsup_triangulate: bad sizes");
    if (t->B == 0) return XSUP_OK;
    if (!world) return fail(XSUP_E_NULL, "xsup_triangulate: world is NULL");
    for (int v = 0; v < t->V; ++v) {
        if (!t->kps[v]) return fail(XSUP_E_NULL, "xsup_triangulate: kps[%d] is NULL", v);
        if (int rc = check_cam(&t->cam[v], "xsup_triangulate")) return rc;
    }
    cudaError_t e = launch_triangulate(*t, world, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "xsup_triangulate launch");
    count_launches(1);
    return XSUP_OK;
}

static int check_root(const void* a, const void* b, int N, int M, int R, int dim, const char* who) {
    if (N < 0 || M < 1 || R < 3 || R % 3 || dim < 1 || dim > 3) return fail(XSUP_E_SHAPE, "%s: need N >= 0, M >= 1, R a positive multiple of 3, 1 <= dim <= 3", who);
    if (N > 0 && (!a || !b)) return fail(XSUP_E_NULL, "%s: NULL pointer", who);
    return XSUP_OK;
}
int xsup_root_centre_fwd(const float* world, float* out, int32_t N, int32_t M, int32_t R, int32_t dim, void* stream) {
    if (int rc = check_root(world, out, N, M, R, dim, "xsup_root_centre_fwd")) return rc;
    if (N == 0) return XSUP_OK;
    cudaError_t e = launch_root_centre_fwd(world, out, N, M, R, dim, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "xsup_root_centre_fwd launch");
    count_launches(1);
    return XSUP_OK;
}
int xsup_root_centre_bwd(const float* g_out, float* g_world, int32_t N, int32_t M, int32_t R, int32_t dim, void* stream) {
    if (int rc = check_root(g_out, g_world, N, M, R, dim, "xsup_root_centre_bwd")) return rc;
    if (N == 0) return XSUP_OK;
    cudaError_t e = launch_root_centre_bwd(g_out, g_world, N, M, R, dim, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "xsup_root_centre_bwd launch");
    count_launches(1);
    return XSUP_OK;
}

int xsup_disc_min_loss_fwd(const float* logits, int32_t B, int32_t NH, int32_t C, float target, float* loss, int64_t* sel, void* stream) {
    if (B < 1 || NH < 1 || C < 1) return fail(XSUP_E_SHAPE, "xsup_disc_min_loss_fwd: need B, NH, C >= 1 (mean of an empty batch is undefined)");
    if (!logits || !loss || !sel) return fail(XSUP_E_NULL, "xsup_disc_min_loss_fwd: NULL pointer");
    cudaError_t e = launch_disc_min_loss_fwd(logits, B, NH, C, target, loss, sel, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "xsup_disc_min_loss_fwd launch");
    count_launches(1);
    return XSUP_OK;
}
int xsup_disc_min_loss_bwd(const float* logits, const int64_t* sel, const float* g_loss, int32_t B, int32_t NH, int32_t C, float target,
                           float* g_logits, void* stream) {
    if (B < 1 || NH < 1 || C < 1) return fail(XSUP_E_SHAPE, "xsup_disc_min_loss_bwd: need B, NH, C >= 1");
    if (!logits || !sel || !g_loss || !g_logits) return fail(XSUP_E_NULL, "xsup_disc_min_loss_bwd: NULL pointer");
    cudaError_t e = launch_disc_min_loss_bwd(logits, sel, g_loss, B, NH, C, target, g_logits, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "xsup_disc_min_loss_bwd launch");
    count_launches(1);
    return XSUP_OK;
}

// ------------------------------------------------------------------------------------------------ conv-fused forward
static int check_conv_shape(const xsup_shape_t* s, int C, const char* who);
int xsup_pack_nhwc_bf16(const float* x_nchw, void* x_nhwc_bf16, int32_t B, int32_t C, int32_t HW, void* stream) {
    if (B < 0 || C < 64 || C % 64 || HW < 64 || HW % 64) return fail(XSUP_E_SHAPE, "xsup_pack_nhwc_bf16: need C and H*W multiples of 64");
    if (B > 65535) return fail(XSUP_E_SHAPE, "xsup_pack_nhwc_bf16: B > 65535");
    if (B == 0) return XSUP_OK;
    if (!x_nchw || !x_nhwc_bf16) return fail(XSUP_E_NULL, "xsup_pack_nhwc_bf16: NULL pointer");
    if (!aligned16(x_nchw) || !aligned16(x_nhwc_bf16)) return fail(XSUP_E_ALIGN, "xsup_pack_nhwc_bf16: tensors must be 16-byte aligned");
    cudaError_t e = launch_pack_nhwc_bf16(x_nchw, x_nhwc_bf16, B, C, HW, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "xsup_pack_nhwc_bf16 launch");
    count_launches(1);
    return XSUP_OK;
}

int xsup_conv_head_fwd(const void* x_nhwc, const void* weight, const float* bias, float* kps, float* depth_prob_map, int64_t* peak_idx,
                       float* stats, float* logits_out, const xsup_shape_t* s, int32_t C, void* stream) {
    if (int rc = check_conv_shape(s, C, "xsup_conv_head_fwd")) return rc;
    if (s->B == 0) return XSUP_OK;
    if (!x_nhwc || !weight || !kps || !depth_prob_map || !stats) return fail(XSUP_E_NULL, "xsup_conv_head_fwd: NULL pointer");
    if (!aligned16(x_nhwc) || !aligned16(weight) || !aligned16(logits_out)) return fail(XSUP_E_ALIGN, "xsup_conv_head_fwd: x/weight/logits_out must be 16-byte aligned");
    int sms = 0;
    if (int rc = device_info(sms)) return rc;
    FwdParams f{};
    f.kps = kps; f.dmap = depth_prob_map; f.peak_idx = peak_idx; f.stats = stats;
    f.n_units = s->B * s->K; f.K = s->K; f.NH = s->NH; f.NS = s->NS; f.head = s->head;
    f.stats_stride = (int)stats_stride(*s);
    f.t.D = s->D; f.t.H = s->H; f.t.W = s->W;
    // the scheduling words after the statistics (word 1 = ticket of xsup_reproj_fused_fwd) start at zero, as after xsup_integral_fwd
    cudaError_t e = cudaMemsetAsync(stats + (size_t)s->B * s->K * f.stats_stride, 0, kSchedWords * sizeof(int), (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "xsup_conv_head_fwd scheduling-word reset");
    e = launch_conv_head_fwd(x_nhwc, weight, bias, logits_out, f, s->B, C, sms, (cudaStream_t)stream);
    if (e == cudaErrorNotSupported) return fail(XSUP_E_DEVICE, "xsup_conv_head_fwd: cuTensorMapEncodeTiled unavailable or rejected the tensors");
    if (e != cudaSuccess) return cuda_fail(e, "xsup_conv_head_fwd launch");
    count_launches(1);
    return XSUP_OK;
}

int xsup_conv_head_fwd_tf32(const void* x_nhwc_f32, const void* weight_f32, const float* bias, float* kps, float* depth_prob_map,
                            int64_t* peak_idx, float* stats, float* logits_out, const xsup_shape_t* s, int32_t C, void* stream) {
    if (int rc = check_conv_shape(s, C, "xsup_conv_head_fwd_tf32")) return rc;
    if (s->B == 0) return XSUP_OK;
    if (!x_nhwc_f32 || !weight_f32 || !kps || !depth_prob_map || !stats) return fail(XSUP_E_NULL, "xsup_conv_head_fwd_tf32: NULL pointer");
    if (!aligned16(x_nhwc_f32) || !aligned16(weight_f32) || !aligned16(logits_out))
        return fail(XSUP_E_ALIGN, "xsup_conv_head_fwd_tf32: x/weight/logits_out must be 16-byte aligned");
    int sms = 0;
    if (int rc = device_info(sms)) return rc;
    FwdParams f{};
    f.kps = kps; f.dmap = depth_prob_map; f.peak_idx = peak_idx; f.stats = stats;
    f.n_units = s->B * s->K; f.K = s->K; f.NH = s->NH; f.NS = s->NS; f.head = s->head;
    f.stats_stride = (int)stats_stride(*s);
    f.t.D = s->D; f.t.H = s->H; f.t.W = s->W;
    cudaError_t e = cudaMemsetAsync(stats + (size_t)s->B * s->K * f.stats_stride, 0, kSchedWords * sizeof(int), (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "xsup_conv_head_fwd_tf32 scheduling-word reset");
    e = launch_conv_head_fwd_tf32(x_nhwc_f32, weight_f32, bias, logits_out, f, s->B, C, sms, (cudaStream_t)stream);
    if (e == cudaErrorNotSupported) return fail(XSUP_E_DEVICE, "xsup_conv_head_fwd_tf32: cuTensorMapEncodeTiled unavailable or rejected the tensors");
    if (e != cudaSuccess) return cuda_fail(e, "xsup_conv_head_fwd_tf32 launch");
    count_launches(1);
    return XSUP_OK;
}

static int check_conv_shape(const xsup_shape_t* s, int C, const char* who) {
    if (int rc = check_shape(s)) return rc;
    if (128 % s->D) return fail(XSUP_E_SHAPE, "%s: depth_dim %d must divide 128 (output rows per CTA)", who, s->D);
    if ((s->H * s->W) % 128 || s->W % 32) return fail(XSUP_E_SHAPE, "%s: H*W must be a multiple of 128 and W of 32", who);
    if (C < 64 || C % 64 || C > 256) return fail(XSUP_E_SHAPE, "%s: channels %d must be 64, 128, 192 or 256", who, C);
    if ((long long)s->B * s->H * s->W > 0x7fffffffLL) return fail(XSUP_E_SHAPE, "%s: too many pixels", who);
    return XSUP_OK;
}

int xsup_integral_coef(const float* stats, const float* g_kps, float* coef_ws, const xsup_shape_t* s, void* stream) {
    if (int rc = check_shape(s)) return rc;
    if (s->B == 0) return XSUP_OK;
    if (!stats || !g_kps || !coef_ws) return fail(XSUP_E_NULL, "xsup_integral_coef: NULL pointer");
    CoefParams c{};
    c.stats = stats; c.g_kps = g_kps; c.coef = coef_ws;
    c.n_units = s->B * s->K; c.K = s->K; c.D = s->D; c.H = s->H; c.W = s->W; c.NH = s->NH; c.NS = s->NS; c.head = s->head;
    c.stats_stride = (int)stats_stride(*s); c.coef_stride = (int)coef_stride(*s);
    c.counter = reinterpret_cast<int*>(coef_ws + (size_t)c.n_units * c.coef_stride);
    cudaError_t e = launch_integral_coef(c, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "xsup_integral_coef launch");
    count_launches(1);
    return XSUP_OK;
}

size_t xsup_conv_bwd_ws_floats(const xsup_shape_t* s) { return s ? (size_t)s->B * conv_bwd_rows_pad(s->K, s->D) * 4 : 0; }

int xsup_conv_head_bwd(const void* x_nhwc, const void* weight, const float* bias, const float* coef_ws, float* rowcoef_ws, void* dx,
                       int32_t dx_f32, float* dw, float* dbias, const xsup_shape_t* s, int32_t C, void* stream) {
    if (int rc = check_conv_shape(s, C, "xsup_conv_head_bwd")) return rc;
    if (s->B == 0) {
        if (dw) {
            cudaError_t e = cudaMemsetAsync(dw, 0, (size_t)s->K * s->D * C * sizeof(float), (cudaStream_t)stream);
            if (e == cudaSuccess && dbias) e = cudaMemsetAsync(dbias, 0, (size_t)s->K * s->D * sizeof(float), (cudaStream_t)stream);
            if (e != cudaSuccess) return cuda_fail(e, "xsup_conv_head_bwd memset");
        }
        return XSUP_OK;
    }
    if (!x_nhwc || !weight || !coef_ws || !rowcoef_ws) return fail(XSUP_E_NULL, "xsup_conv_head_bwd: NULL pointer");
    if (dbias && !dw) return fail(XSUP_E_NULL, "xsup_conv_head_bwd: dbias is produced by the dw launch, pass dw as well");
    if (!aligned16(x_nhwc) || !aligned16(weight) || !aligned16(rowcoef_ws) || !aligned16(dx) || !aligned16(dw))
        return fail(XSUP_E_ALIGN, "xsup_conv_head_bwd: x/weight/rowcoef_ws/dx/dw must be 16-byte aligned");
    int sms = 0;
    if (int rc = device_info(sms)) return rc;
    cudaError_t e = launch_conv_head_bwd(x_nhwc, weight, bias, coef_ws, (int)coef_stride(*s), rowcoef_ws, dx, dx_f32 ? 1 : 0, dw, dbias, s->B,
                                         s->K, s->D, s->H, s->W, C, sms, (cudaStream_t)stream);
    if (e == cudaErrorNotSupported) return fail(XSUP_E_DEVICE, "xsup_conv_head_bwd: cuTensorMapEncodeTiled unavailable or rejected the tensors");
    if (e != cudaSuccess) return cuda_fail(e, "xsup_conv_head_bwd launch");
    count_launches(1 + (dw ? 1 : 0) + (dx ? 1 : 0));
    return XSUP_OK;
}

int xsup_conv_head_bwd_g(const void* x_nhwc, const void* weight, const float* bias, const float* coef_ws, void* g_out, float* gbias_part,
                         const xsup_shape_t* s, int32_t C, void* stream) {
    if (int rc = check_conv_shape(s, C, "xsup_conv_head_bwd_g")) return rc;
    if (s->B == 0) return XSUP_OK;
    if (!x_nhwc || !weight || !coef_ws || !g_out) return fail(XSUP_E_NULL, "xsup_conv_head_bwd_g: NULL pointer");
    if (!aligned16(x_nhwc) || !aligned16(weight) || !aligned16(g_out)) return fail(XSUP_E_ALIGN, "xsup_conv_head_bwd_g: x/weight/g_out must be 16-byte aligned");
    int sms = 0;
    if (int rc = device_info(sms)) return rc;
    FwdParams f{};
    f.n_units = s->B * s->K; f.K = s->K; f.NH = s->NH; f.NS = s->NS; f.head = s->head;
    f.t.D = s->D; f.t.H = s->H; f.t.W = s->W;
    cudaError_t e = launch_conv_head_bwd_g(x_nhwc, weight, bias, coef_ws, (int)coef_stride(*s), g_out, gbias_part, f, s->B, C, sms,
                                           (cudaStream_t)stream);
    if (e == cudaErrorNotSupported) return fail(XSUP_E_DEVICE, "xsup_conv_head_bwd_g: cuTensorMapEncodeTiled unavailable or rejected the tensors");
    if (e != cudaSuccess) return cuda_fail(e, "xsup_conv_head_bwd_g launch");
    count_launches(1);
    return XSUP_OK;
}

// ------------------------------------------------------------------------------------------------ stand-alone pose loss terms
static int pose_term_params(PoseTermParams& p, double& denom, const float* x, const float* gt, const float* fs, int term, int sum_mode,
                            int B, int K, int C, const char* who) {
    if (term < XSUP_TERM_MSE || term > XSUP_TERM_KP) return fail(XSUP_E_SHAPE, "%s: unknown term %d", who, term);
    if (B < 1 || K < 1 || C < 2 || C > 3) return fail(XSUP_E_SHAPE, "%s: need B >= 1, K >= 1, C in {2,3}", who);
    if (term == XSUP_TERM_BONE && K < 17) return fail(XSUP_E_SHAPE, "%s: the bone term indexes joints up to 16 (loss_func.py:20), K is %d", who, K);
    if (term == XSUP_TERM_KP && K < 15) return fail(XSUP_E_SHAPE, "%s: the keypoint term indexes joints up to 14 (loss_func.py:28), K is %d", who, K);
    if (!x || (term == XSUP_TERM_MSE && !gt)) return fail(XSUP_E_NULL, "%s: NULL pointer", who);
    p = PoseTermParams{};
    p.x = x; p.gt = gt; p.B = B; p.K = K; p.C = C; p.term = term; p.use_fs = (term == XSUP_TERM_MSE && fs) ? 1 : 0;
    if (p.use_fs) { p.fs[0] = fs[0]; p.fs[1] = fs[1]; p.fs[2] = fs[2]; }
    p.is_3d = (term == XSUP_TERM_KP && sum_mode) ? 1 : 0;
    if (term == XSUP_TERM_MSE) denom = sum_mode ? (double)B : (double)B * K * C;
    else if (term == XSUP_TERM_BONE) denom = (double)B * 4.0;
    else denom = (double)B * 2.0 * C;
    return XSUP_OK;
}

int xsup_pose_term_fwd(const float* x, const float* gt, const float* feature_shape, int32_t term, int32_t sum_mode, int32_t B, int32_t K,
                       int32_t C, float* sample_ws, float* loss, void* stream) {
    PoseTermParams p;
    double denom = 1.0;
    if (int rc = pose_term_params(p, denom, x, gt, feature_shape, term, sum_mode, B, K, C, "xsup_pose_term_fwd")) return rc;
    if (!sample_ws || !loss) return fail(XSUP_E_NULL, "xsup_pose_term_fwd: NULL pointer");
    cudaError_t e = launch_pose_term_fwd(p, denom, sample_ws, loss, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "xsup_pose_term_fwd launch");
    count_launches(2);
    return XSUP_OK;
}

int xsup_pose_term_bwd(const float* x, const float* gt, const float* feature_shape, int32_t term, int32_t sum_mode, int32_t B, int32_t K,
                       int32_t C, const float* g_loss, float* g_x, void* stream) {
    PoseTermParams p;
    double denom = 1.0;
    if (int rc = pose_term_params(p, denom, x, gt, feature_shape, term, sum_mode, B, K, C, "xsup_pose_term_bwd")) return rc;
    if (!g_loss || !g_x) return fail(XSUP_E_NULL, "xsup_pose_term_bwd: NULL pointer");
    cudaError_t e = launch_pose_term_bwd(p, denom, g_loss, g_x, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "xsup_pose_term_bwd launch");
    count_launches(1);
    return XSUP_OK;
}

int xsup_pose_sqerr(const float* x, const float* gt, const float* feature_shape, int32_t B, int32_t K, int32_t C, const float* g_out,
                    float* out, void* stream) {
    PoseTermParams p;
    double denom = 1.0;
    if (B == 0) return XSUP_OK;
    if (int rc = pose_term_params(p, denom, x, gt, feature_shape, XSUP_TERM_MSE, 0, B, K, C, "xsup_pose_sqerr")) return rc;
    if (!out) return fail(XSUP_E_NULL, "xsup_pose_sqerr: NULL pointer");
    cudaError_t e = launch_pose_sqerr(p, g_out, out, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "xsup_pose_sqerr launch");
    count_launches(1);
    return XSUP_OK;
}

}  // extern "C"
