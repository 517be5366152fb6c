// c_api.cu -- extern "C" boundary (include/gg_raster.h).  Argument validation, workspace
// carving, launch sequencing, error reporting.  No device-memory allocation.  Process-wide state: the launch counter and
// the optional per-kernel timing events (diagnostic, off by default), environment switches read once, and the forward's
// fork sets (a side stream + two events per (device, caller stream), see ForkSet); the compute path is re-entrant.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <cstring>
#include <utility>
#include "common.cuh"

namespace {

thread_local char g_err[512] = "";
// profiling facilities (off by default).  Process-wide because autograd runs backward on a
// worker thread: launch counts and per-kernel events must be visible from the caller's thread.
std::atomic<int64_t> g_launches{0};

enum KernelSlot { K_PROJECT = 0, K_SCAN, K_SHCOLOR, K_EMIT, K_SORTPACK, K_BLENDFWD, K_BLENDBWD, K_PREBWD, K_MESHFWD, K_MESHBWD, K_LOSSFWD, K_LOSSBWD, K_RAYCAST, K_COUNT };
const char* const kKernelNames[K_COUNT] = {"project", "tile_scan", "sh_color", "emit", "sort_pack",
                                           "blend_fwd", "blend_bwd", "preprocess_bwd", "mesh_bind_fwd",
                                           "mesh_bind_bwd", "photometric_fwd", "photometric_bwd", "cast_rays"};
std::atomic<int> g_timing{0};
std::mutex g_timing_mu;
// per-kernel timing events belong to ONE device (the one that was current when timing was enabled): launches on any
// other device are simply not timed (recording a foreign device's event is an invalid-resource-handle error)
cudaEvent_t g_ev[K_COUNT][2];
bool g_ev_made = false;
int g_ev_device = -1;
bool g_ev_used[K_COUNT] = {false};

struct ScopedKernelTimer {
    int slot;
    cudaStream_t s;
    bool on;
    ScopedKernelTimer(int slot_, cudaStream_t s_) : slot(slot_), s(s_), on(g_timing.load() != 0) {
        if (on) {
            int dev = -1;
            cudaGetDevice(&dev);
            std::lock_guard<std::mutex> lk(g_timing_mu);
            on = g_ev_made && dev == g_ev_device;
            if (on) cudaEventRecord(g_ev[slot][0], s);
        }
    }
    ~ScopedKernelTimer() {
        if (on) {
            std::lock_guard<std::mutex> lk(g_timing_mu);
            cudaEventRecord(g_ev[slot][1], s);
            g_ev_used[slot] = true;
        }
    }
};

int fail(int code, const char* fmt, const char* what = "") {
    snprintf(g_err, sizeof(g_err), fmt, what);
    return code;
}
int cuda_fail(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
    return (int)e;
}
#define GG_CUDA(call)                                         \
    do {                                                      \
        cudaError_t _e = (call);                              \
        if (_e != cudaSuccess) return cuda_fail(_e, #call);   \
    } while (0)

// ---- forward fork: SH -> RGB on a side stream, concurrently with the tile scan and the instance emission ----------
// sh_color is HBM-bound and needs only the projection's radii; tile_scan is ONE CTA and emit is atomics-bound, so the
// three overlap almost for free.  One fork set per (device, caller stream): gg_forward_project records `projected`
// right behind project_kernel, gg_forward_color launches on `side` behind it and records `colored`, gg_forward_render
// joins before the first reader of the colours.  Event record / wait pairs are capturable, so the fork survives inside
// a CUDA graph.  Side streams come from a small per-device pool created on first (eager) use -- nothing is created
// while a capture is in progress except events.
struct ForkSet {
    cudaStream_t side = nullptr;
    cudaEvent_t projected = nullptr, colored = nullptr;
    bool armed = false;        // `projected` was recorded by the latest gg_forward_project on this stream
    bool pending = false;      // sh_color is in flight on `side`: the next render on this stream must join
};
constexpr int FORK_POOL = 4, FORK_MAX_DEV = 64;
std::mutex g_fork_mu;
std::map<std::pair<int, cudaStream_t>, ForkSet> g_forks;
cudaStream_t g_fork_pool[FORK_MAX_DEV][FORK_POOL];
bool g_fork_pool_made[FORK_MAX_DEV] = {false};
int g_fork_next[FORK_MAX_DEV] = {0};

// nullptr when forking is off (GG_FWD_FORK=0, per-kernel timing on, debug mode) or the pool cannot be set up here
ForkSet* fork_set(int device, cudaStream_t s, bool debug) {
    static const bool off = []() { const char* e = getenv("GG_FWD_FORK"); return e && !strcmp(e, "0"); }();
    if (off || debug || g_timing.load() != 0 || device < 0 || device >= FORK_MAX_DEV) return nullptr;
    std::lock_guard<std::mutex> lk(g_fork_mu);
    if (!g_fork_pool_made[device]) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(s, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return nullptr;
        for (int i = 0; i < FORK_POOL; i++)
            if (cudaStreamCreateWithFlags(&g_fork_pool[device][i], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        g_fork_pool_made[device] = true;
    }
    auto key = std::make_pair(device, s);
    auto it = g_forks.find(key);
    if (it == g_forks.end()) {
        ForkSet f;
        f.side = g_fork_pool[device][g_fork_next[device]++ % FORK_POOL];
        if (cudaEventCreateWithFlags(&f.projected, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&f.colored, cudaEventDisableTiming) != cudaSuccess)
            return nullptr;
        it = g_forks.emplace(key, f).first;
    }
    return &it->second;
}

// environment switches are read ONCE per process (function-local statics at the call sites)
inline bool env_is(const char* name, const char* value) {
    const char* e = getenv(name);
    return e && !strcmp(e, value);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int check_view(const gg_view* v) {
    if (!v) return fail(GG_E_BADARG, "view is NULL");
    if (v->num_gaussians < 0 || v->image_width < 0 || v->image_height < 0) return fail(GG_E_BADARG, "negative size in view");
    if (v->sh_degree < 0 || v->sh_degree > 3) return fail(GG_E_BADARG, "sh_degree must be in 0..3");
    if (v->image_width > 65535 * GG_TILE || v->image_height > 65535 * GG_TILE) return fail(GG_E_BADARG, "image too large");
    return 0;
}

int check_inputs(const gg_view* v, const gg_inputs* in) {
    if (!in) return fail(GG_E_BADARG, "inputs is NULL");
    if (!in->bg || !in->viewmatrix || !in->projmatrix || !in->campos) return fail(GG_E_BADARG, "bg/viewmatrix/projmatrix/campos must be non-NULL");
    if (v->num_gaussians == 0) return 0;
    if (!in->means3D || !in->opacities) return fail(GG_E_BADARG, "means3D/opacities must be non-NULL");
    if ((in->shs == nullptr) == (in->colors_precomp == nullptr)) return fail(GG_E_BADARG, "provide exactly one of shs / colors_precomp");
    const bool sr = in->scales && in->rotations;
    if (sr == (in->cov3D_precomp != nullptr) || (!sr && (in->scales || in->rotations)))
        return fail(GG_E_BADARG, "provide exactly one of (scales, rotations) / cov3D_precomp");
    if (in->shs && v->sh_coeffs < (v->sh_degree + 1) * (v->sh_degree + 1)) return fail(GG_E_BADARG, "sh_coeffs < (sh_degree+1)^2");
    if (in->shs && v->sh_coeffs > 16) return fail(GG_E_BADARG, "sh_coeffs > 16 (degree <= 3) not supported");
    if (in->rotations && !aligned16(in->rotations)) return fail(GG_E_ALIGN, "rotations not 16-byte aligned");
    if (in->shs && v->sh_coeffs == 16 && !aligned16(in->shs)) return fail(GG_E_ALIGN, "shs not 16-byte aligned");
    return 0;
}

int after_launch(const gg_view* v, cudaStream_t s, const char* where) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, where);
    if (v && v->debug) {
        e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) return cuda_fail(e, where);
    }
    return 0;
}
#define GG_AFTER(where)                                 \
    do {                                                \
        int _rc = after_launch(view, s, where);         \
        if (_rc) return _rc;                            \
    } while (0)

}  // namespace

using namespace gg;

extern "C" {

int gg_abi_version(void) { return GG_ABI_VERSION; }
const char* gg_version(void) { return "gg_raster 0.1.0 (sm_100a)"; }
const char* gg_last_error(void) { return g_err; }
int64_t gg_launch_count(int reset) {
    const int64_t v = g_launches.load();
    if (reset) g_launches.store(0);
    return v;
}

int gg_kernel_timing(int enable) {
    std::lock_guard<std::mutex> lk(g_timing_mu);
    int dev = -1;
    GG_CUDA(cudaGetDevice(&dev));
    if (enable && g_ev_made && dev != g_ev_device) {          // timing moves to the caller's current device
        for (int k = 0; k < K_COUNT; k++)
            for (int j = 0; j < 2; j++) cudaEventDestroy(g_ev[k][j]);
        g_ev_made = false;
    }
    if (enable && !g_ev_made) {
        for (int k = 0; k < K_COUNT; k++)
            for (int j = 0; j < 2; j++) GG_CUDA(cudaEventCreate(&g_ev[k][j]));
        g_ev_made = true;
        g_ev_device = dev;
    }
    for (int k = 0; k < K_COUNT; k++) g_ev_used[k] = false;
    g_timing.store(enable ? 1 : 0);
    return 0;
}

int gg_kernel_count(void) { return K_COUNT; }
const char* gg_kernel_name(int slot) { return (slot >= 0 && slot < K_COUNT) ? kKernelNames[slot] : ""; }

int gg_kernel_times(float* ms_out) {
    if (!ms_out) return fail(GG_E_BADARG, "ms_out is NULL");
    std::lock_guard<std::mutex> lk(g_timing_mu);
    for (int k = 0; k < K_COUNT; k++) {
        ms_out[k] = -1.f;
        if (!g_ev_made || !g_ev_used[k]) continue;
        GG_CUDA(cudaEventSynchronize(g_ev[k][1]));
        float ms = 0.f;
        GG_CUDA(cudaEventElapsedTime(&ms, g_ev[k][0], g_ev[k][1]));
        ms_out[k] = ms;
    }
    return 0;
}

int gg_forward_workspace_bytes(const gg_view* view, size_t* geom_bytes, size_t* tile_bytes, size_t* image_bytes) {
    if (int rc = check_view(view)) return rc;
    const int64_t gx = (view->image_width + TILE - 1) / TILE, gy = (view->image_height + TILE - 1) / TILE;
    if (geom_bytes) *geom_bytes = geom_layout(nullptr, view->num_gaussians > 0 ? view->num_gaussians : 1, nullptr);
    if (tile_bytes) *tile_bytes = tile_layout(nullptr, gx * gy, nullptr);
    if (image_bytes) *image_bytes = image_layout(nullptr, (int64_t)view->image_width * view->image_height, nullptr);
    return 0;
}

int gg_instance_workspace_bytes(int64_t num_rendered, size_t* key_bytes, size_t* record_bytes) {
    if (num_rendered < 0 || num_rendered > 0xfffffff0ll) return fail(GG_E_BADARG, "num_rendered out of range");
    // two key arrays: tile-bucketed keys + the depth-bucketed copy of the lazy forward path
    if (key_bytes) *key_bytes = 2 * align_up((size_t)(num_rendered > 0 ? num_rendered : 1) * 8);
    if (record_bytes) *record_bytes = record_layout(nullptr, num_rendered, nullptr);
    return 0;
}

int gg_backward_workspace_bytes(const gg_view* view, size_t* accum_bytes) {
    if (int rc = check_view(view)) return rc;
    if (accum_bytes) *accum_bytes = accum_layout(nullptr, view->num_gaussians, nullptr);
    return 0;
}

int gg_forward_project(const gg_view* view, const gg_inputs* in, void* geom_ws, void* tile_ws, int32_t* radii,
                       uint32_t* num_rendered_host, int device, void* stream) {
    if (int rc = check_view(view)) return rc;
    if (int rc = check_inputs(view, in)) return rc;
    if (!geom_ws || !tile_ws || (!radii && view->num_gaussians > 0)) return fail(GG_E_BADARG, "workspace / radii is NULL");
    GG_CUDA(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream;
    const int gx = (view->image_width + TILE - 1) / TILE, gy = (view->image_height + TILE - 1) / TILE;
    const int T = gx * gy;
    GeomWS g;
    TileWS t;
    geom_layout(geom_ws, view->num_gaussians > 0 ? view->num_gaussians : 1, &g);
    tile_layout(tile_ws, T, &t);
    // count | fill | misc are adjacent: one memset covers all three
    GG_CUDA(cudaMemsetAsync(t.count, 0, (size_t)((char*)t.offset - (char*)t.count), s));
    { ScopedKernelTimer kt(K_PROJECT, s); g_launches += launch_project(*view, *in, g, t, radii, s); }
    GG_AFTER("project_kernel");
    if (ForkSet* f = fork_set(device, s, view->debug != 0)) {      // fork point for the colour kernel (see ForkSet)
        GG_CUDA(cudaEventRecord(f->projected, s));
        f->armed = true;
    }
    { ScopedKernelTimer kt(K_SCAN, s); g_launches += launch_tile_scan(T, t, s); }
    GG_AFTER("tile_scan_kernel");
    if (num_rendered_host) GG_CUDA(cudaMemcpyAsync(num_rendered_host, t.misc, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    return 0;
}

int gg_forward_color(const gg_view* view, const gg_inputs* in, void* geom_ws, const int32_t* radii, int device,
                     void* stream) {
    if (int rc = check_view(view)) return rc;
    if (int rc = check_inputs(view, in)) return rc;
    if (view->num_gaussians == 0) return 0;
    if (!geom_ws || !radii) return fail(GG_E_BADARG, "workspace / radii is NULL");
    GG_CUDA(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream;
    GeomWS g;
    geom_layout(geom_ws, view->num_gaussians, &g);
    ForkSet* f = fork_set(device, s, view->debug != 0);
    if (f && f->armed) {            // side stream, behind the projection only; joined by the next render on `s`
        f->armed = false;
        GG_CUDA(cudaStreamWaitEvent(f->side, f->projected, 0));
        // a resident-size grid (default 4 CTAs per SM on 148 SMs) that strides over the slabs: see sh_color16_kernel
        static const int fork_blocks = []() { const char* e = getenv("GG_SH_FORK_BLOCKS"); return e ? atoi(e) : 592; }();
        g_launches += launch_sh_color(*view, *in, g, radii, f->side, fork_blocks);
        GG_CUDA(cudaGetLastError());
        GG_CUDA(cudaEventRecord(f->colored, f->side));
        f->pending = true;
        return 0;
    }
    { ScopedKernelTimer kt(K_SHCOLOR, s); g_launches += launch_sh_color(*view, *in, g, radii, s); }
    GG_AFTER("sh_color_kernel");
    return 0;
}

static int forward_render_impl(const gg_view* view, const gg_inputs* in, const void* geom_ws, void* tile_ws, void* key_ws,
                               void* record_ws, int64_t instance_capacity, int64_t max_tile_instances, void* image_ws,
                               const int32_t* radii, float* out_color, float* out_depth, float* out_alpha, bool late_color,
                               void* color_gate_event, uint32_t* color_gate_words, int device, void* stream) {
    if (int rc = check_view(view)) return rc;
    if (int rc = check_inputs(view, in)) return rc;
    if (!geom_ws || !tile_ws || !key_ws || !record_ws || !image_ws || !out_color || !out_depth || !out_alpha)
        return fail(GG_E_BADARG, "NULL workspace or output");
    if (instance_capacity < 0 || instance_capacity > 0xfffffff0ll) return fail(GG_E_BADARG, "instance_capacity out of range");
    GG_CUDA(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream;
    const int gx = (view->image_width + TILE - 1) / TILE, gy = (view->image_height + TILE - 1) / TILE;
    const int T = gx * gy;
    const size_t P = (size_t)view->image_width * view->image_height;
    if (view->num_gaussians == 0) {   // upstream: all-zero outputs when there is nothing to draw
        GG_CUDA(cudaMemsetAsync(out_color, 0, 3 * P * sizeof(float), s));
        GG_CUDA(cudaMemsetAsync(out_depth, 0, P * sizeof(float), s));
        GG_CUDA(cudaMemsetAsync(out_alpha, 0, P * sizeof(float), s));
        return 0;
    }
    GeomWS g;
    TileWS t;
    RecordWS r;
    ImageWS img;
    geom_layout(const_cast<void*>(geom_ws), view->num_gaussians, &g);
    tile_layout(tile_ws, T, &t);
    record_layout(record_ws, instance_capacity, &r);
    image_layout(image_ws, (int64_t)P, &img);
    const uint32_t cap = (uint32_t)instance_capacity;
    // emit cursors start at zero (also makes a re-run after GG_E_OVERFLOW-style capacity growth self-contained)
    GG_CUDA(cudaMemsetAsync(t.fill, 0, (size_t)T * sizeof(uint32_t), s));
    { ScopedKernelTimer kt(K_EMIT, s); g_launches += launch_emit(*view, g, t, radii, (uint64_t*)key_ws, cap, s); }
    GG_AFTER("emit_kernel");
    {   // join a colour kernel forked by gg_forward_color: sort_pack / the lazy kernel read geom.rgb
        std::unique_lock<std::mutex> lk(g_fork_mu);
        auto it = g_forks.find(std::make_pair(device, s));
        if (it != g_forks.end() && it->second.pending) {
            it->second.pending = false;
            cudaEvent_t ev = it->second.colored;
            lk.unlock();
            GG_CUDA(cudaStreamWaitEvent(s, ev, 0));
        }
    }
    // Forward path: "tma" = per-tile full sort + pack, then the bulk-TMA streamed blend;
    //               "lazy" = fused bucket-sort + pack + blend that stops at tile saturation (dense scenes).
    // auto: lazy when some tile holds more instances than the default shared-memory sort handles.
    bool lazy = max_tile_instances > 4096;
    if (const char* e = getenv("GG_FWD_PATH")) {   // test/diagnostic override (tests flip it between calls)
        if (!strcmp(e, "lazy")) lazy = true;
        else if (!strcmp(e, "tma")) lazy = false;
    }
    auto colours = [&]() -> int {                  // late-colour mode: SH -> RGB behind the gate, after the sort
        if (color_gate_event) GG_CUDA(cudaStreamWaitEvent(s, (cudaEvent_t)color_gate_event, 0));
        if (color_gate_words) g_launches += launch_gate_wait(color_gate_words, s);      // device-side gate (graph replays)
        { ScopedKernelTimer kt(K_SHCOLOR, s); g_launches += launch_sh_color(*view, *in, g, radii, s); }
        return after_launch(view, s, "sh_color_kernel");
    };
    if (lazy) {
        if (late_color) { if (int rc = colours()) return rc; }     // the fused kernel packs colours itself
        uint64_t* keys2 = (uint64_t*)((char*)key_ws + align_up((size_t)(instance_capacity > 0 ? instance_capacity : 1) * 8));
        ScopedKernelTimer kt(K_BLENDFWD, s);
        g_launches += launch_blend_fwd_lazy(*view, *in, g, t, (uint64_t*)key_ws, keys2, r, img, cap, out_color,
                                            out_depth, out_alpha, s);
    } else {
        {
            ScopedKernelTimer kt(K_SORTPACK, s);
            g_launches += launch_sort_pack(*view, g, t, (uint64_t*)key_ws, r, cap,
                                           (uint32_t)(max_tile_instances < 0 ? 0 : max_tile_instances), !late_color, s);
        }
        GG_AFTER("sort_pack_kernel");
        if (late_color) {
            if (int rc = colours()) return rc;
            g_launches += launch_color_fill(*view, g, t, r, cap, s);
            GG_AFTER("color_fill_kernel");
        }
        // "v2" (default): decoupled warps, branch-free body (blend_fwd2.cu); "v1": round 1's kernel, kept for A/B runs
        static const bool fwd_v1 = env_is("GG_FWD_KERNEL", "v1");
        ScopedKernelTimer kt(K_BLENDFWD, s);
        g_launches += fwd_v1 ? launch_blend_fwd(*view, *in, t, r, img, cap, out_color, out_depth, out_alpha, s)
                             : launch_blend_fwd2(*view, *in, t, r, img, cap, out_color, out_depth, out_alpha, s);
    }
    GG_AFTER("blend_fwd_kernel");
    return 0;
}

int gg_forward_render(const gg_view* view, const gg_inputs* in, const void* geom_ws, void* tile_ws, void* key_ws,
                      void* record_ws, int64_t instance_capacity, int64_t max_tile_instances, void* image_ws,
                      const int32_t* radii, float* out_color, float* out_depth, float* out_alpha, int device,
                      void* stream) {
    return forward_render_impl(view, in, geom_ws, tile_ws, key_ws, record_ws, instance_capacity, max_tile_instances,
                               image_ws, radii, out_color, out_depth, out_alpha, false, nullptr, nullptr, device, stream);
}

int gg_forward_render_late_color(const gg_view* view, const gg_inputs* in, void* geom_ws, void* tile_ws, void* key_ws,
                                 void* record_ws, int64_t instance_capacity, int64_t max_tile_instances, void* image_ws,
                                 const int32_t* radii, float* out_color, float* out_depth, float* out_alpha,
                                 void* color_gate_event, uint32_t* color_gate_words, int device, void* stream) {
    return forward_render_impl(view, in, geom_ws, tile_ws, key_ws, record_ws, instance_capacity, max_tile_instances,
                               image_ws, radii, out_color, out_depth, out_alpha, true, color_gate_event, color_gate_words,
                               device, stream);
}

int gg_gate_signal(uint32_t* gate_words, int device, void* stream) {
    if (!gate_words) return fail(GG_E_BADARG, "gate_words is NULL");
    GG_CUDA(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream;
    const gg_view* view = nullptr;
    g_launches += launch_gate_signal(gate_words, s);
    GG_AFTER("gate_signal_kernel");
    return 0;
}

int gg_forward_overflow_check(const gg_view* view, const void* tile_ws, int64_t instance_capacity, uint32_t* flag2,
                              int device, void* stream) {
    if (int rc = check_view(view)) return rc;
    if (!tile_ws || !flag2) return fail(GG_E_BADARG, "NULL argument");
    if (instance_capacity < 0 || instance_capacity > 0xfffffff0ll) return fail(GG_E_BADARG, "instance_capacity out of range");
    GG_CUDA(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream;
    const int gx = (view->image_width + TILE - 1) / TILE, gy = (view->image_height + TILE - 1) / TILE;
    TileWS t;
    tile_layout(const_cast<void*>(tile_ws), gx * gy, &t);
    g_launches += launch_overflow_flag(t, (uint32_t)instance_capacity, flag2, s);
    GG_AFTER("overflow_flag_kernel");
    return 0;
}

int gg_backward(const gg_view* view, const gg_inputs* in, const void* tile_ws, const void* record_ws,
                int64_t instance_capacity, const void* image_ws, const int32_t* radii, void* accum_ws,
                const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha, float* dL_dmeans3D,
                float* dL_dmeans2D, float* dL_dshs, float* dL_dcolors_precomp, float* dL_dopacities,
                float* dL_dscales, float* dL_drotations, float* dL_dcov3D, int device, void* stream) {
    if (int rc = check_view(view)) return rc;
    if (int rc = check_inputs(view, in)) return rc;
    if (view->num_gaussians == 0) return 0;
    if (!tile_ws || !record_ws || !image_ws || !radii || !accum_ws) return fail(GG_E_BADARG, "NULL workspace");
    if (instance_capacity < 0 || instance_capacity > 0xfffffff0ll) return fail(GG_E_BADARG, "instance_capacity out of range");
    if (dL_drotations && !aligned16(dL_drotations)) return fail(GG_E_ALIGN, "dL_drotations not 16-byte aligned");
    if (dL_dshs && view->sh_coeffs == 16 && !aligned16(dL_dshs)) return fail(GG_E_ALIGN, "dL_dshs not 16-byte aligned");
    GG_CUDA(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream;
    const int gx = (view->image_width + TILE - 1) / TILE, gy = (view->image_height + TILE - 1) / TILE;
    const int T = gx * gy;
    TileWS t;
    RecordWS r;
    ImageWS img;
    AccumWS acc;
    tile_layout(const_cast<void*>(tile_ws), T, &t);
    record_layout(const_cast<void*>(record_ws), instance_capacity, &r);
    image_layout(const_cast<void*>(image_ws), (int64_t)view->image_width * view->image_height, &img);
    accum_layout(accum_ws, view->num_gaussians, &acc);
    GG_CUDA(cudaMemsetAsync(accum_ws, 0, accum_layout(nullptr, view->num_gaussians, nullptr), s));
    {
        // "v2" (default): pixel-parallel evaluation + splat-parallel moment reduction (blend_bwd2.cu);
        // "v1": round 1's shuffle reduce-scatter kernel (blend_bwd.cu), kept for A/B measurements.
        static const bool bwd_v1 = env_is("GG_BWD_PATH", "v1");
        static const int bwd2_minb = env_is("GG_BWD2_MINB", "3") ? 3 : 4;
        ScopedKernelTimer kt(K_BLENDBWD, s);
        g_launches += bwd_v1 ? launch_blend_bwd(*view, *in, t, r, img, dL_dcolor, dL_ddepth, dL_dalpha, acc, s)
                             : launch_blend_bwd2(*view, *in, t, r, img, dL_dcolor, dL_ddepth, dL_dalpha, acc, bwd2_minb, s);
    }
    GG_AFTER("blend_bwd_kernel");
    {
        ScopedKernelTimer kt(K_PREBWD, s);
        g_launches += launch_preprocess_bwd(*view, *in, radii, acc, dL_dmeans3D, dL_dmeans2D, dL_dshs,
                                            dL_dcolors_precomp, dL_dopacities, dL_dscales, dL_drotations, dL_dcov3D, s);
    }
    GG_AFTER("preprocess_bwd_kernel");
    return 0;
}

int gg_mark_visible(int32_t num_gaussians, const float* means3D, const float* viewmatrix, const float* projmatrix,
                    uint8_t* visible, int device, void* stream) {
    (void)projmatrix;
    if (num_gaussians < 0) return fail(GG_E_BADARG, "negative num_gaussians");
    if (num_gaussians == 0) return 0;
    if (!means3D || !viewmatrix || !visible) return fail(GG_E_BADARG, "NULL argument");
    GG_CUDA(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream;
    const gg_view* view = nullptr;
    g_launches += launch_mark_visible(num_gaussians, means3D, viewmatrix, visible, s);
    GG_AFTER("mark_visible_kernel");
    return 0;
}

int gg_debug_read_geom(const gg_view* view, const void* geom_ws, float* xy, float* depth, float* conic_opacity,
                       float* rgb, uint32_t* rect, int device, void* stream) {
    if (int rc = check_view(view)) return rc;
    if (!geom_ws) return fail(GG_E_BADARG, "geom_ws is NULL");
    const int64_t N = view->num_gaussians;
    if (N == 0) return 0;
    GG_CUDA(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream;
    GeomWS g;
    geom_layout(const_cast<void*>(geom_ws), N, &g);
    const cudaMemcpyKind k = cudaMemcpyDefault;
    if (xy) GG_CUDA(cudaMemcpyAsync(xy, g.xy, N * 8, k, s));
    if (depth) GG_CUDA(cudaMemcpyAsync(depth, g.depth, N * 4, k, s));
    if (conic_opacity) GG_CUDA(cudaMemcpyAsync(conic_opacity, g.conic_o, N * 16, k, s));
    if (rgb) GG_CUDA(cudaMemcpyAsync(rgb, g.rgb, N * 12, k, s));
    if (rect) {   // unpack on the host side is not possible for device destinations: copy packed pairs
        GG_CUDA(cudaMemcpyAsync(rect, g.rect, N * 8, k, s));
    }
    return 0;
}

int gg_debug_lazy_phase_counters(uint64_t* counters2) {
    set_lazy_phase_counters(reinterpret_cast<unsigned long long*>(counters2));
    return 0;
}

int gg_debug_read_binning(const gg_view* view, const void* tile_ws, const void* record_ws, int64_t instance_capacity,
                          uint32_t* tile_offsets, uint32_t* sorted_ids, int device, void* stream) {
    if (int rc = check_view(view)) return rc;
    if (!tile_ws || !record_ws) return fail(GG_E_BADARG, "NULL workspace");
    if (instance_capacity < 0 || instance_capacity > 0xfffffff0ll) return fail(GG_E_BADARG, "instance_capacity out of range");
    GG_CUDA(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream;
    const int gx = (view->image_width + TILE - 1) / TILE, gy = (view->image_height + TILE - 1) / TILE;
    const int T = gx * gy;
    TileWS t;
    RecordWS r;
    tile_layout(const_cast<void*>(tile_ws), T, &t);
    record_layout(const_cast<void*>(record_ws), instance_capacity, &r);
    if (tile_offsets) GG_CUDA(cudaMemcpyAsync(tile_offsets, t.offset, (size_t)(T + 1) * 4, cudaMemcpyDefault, s));
    if (sorted_ids && instance_capacity > 0)   // ids sit in the .w lane of plane p2: strided gather
        GG_CUDA(cudaMemcpy2DAsync(sorted_ids, 4, reinterpret_cast<const char*>(r.p2) + 12, 16, 4,
                                  (size_t)instance_capacity, cudaMemcpyDefault, s));
    return 0;
}

// ---- fused mesh-binding transform ------------------------------------------
int gg_mesh_bind_workspace_bytes(int32_t num_faces, size_t* frame_bytes) {
    if (num_faces < 0) return fail(GG_E_BADARG, "negative num_faces");
    if (frame_bytes) *frame_bytes = align_up((size_t)(num_faces > 0 ? num_faces : 1) * 17 * sizeof(float));
    return 0;
}

int gg_mesh_bind_forward_ex(int32_t num_vertices, int32_t num_faces, int32_t num_gaussians, const float* verts,
                            const int32_t* faces, const int32_t* binding, const float* local_xyz,
                            const float* local_log_scaling, const float* local_rotation, const float* barycentric,
                            const float* face_scaling_remembered, void* frame_ws, float* out_xyz, float* out_scaling,
                            float* out_rotation, int device, void* stream) {
    (void)num_vertices;
    if (num_faces < 0 || num_gaussians < 0) return fail(GG_E_BADARG, "negative size");
    if (num_gaussians > 0 && (!verts || !faces || !binding || !local_xyz || !local_log_scaling || !local_rotation ||
                              !frame_ws || !out_xyz || !out_scaling || !out_rotation))
        return fail(GG_E_BADARG, "NULL argument");
    if (!aligned16(local_rotation) || !aligned16(out_rotation)) return fail(GG_E_ALIGN, "rotation tensors not 16-byte aligned");
    GG_CUDA(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream;
    const gg_view* view = nullptr;
    {
        ScopedKernelTimer kt(K_MESHFWD, s);
        g_launches += launch_mesh_bind_forward(num_faces, num_gaussians, verts, faces, binding, local_xyz, local_log_scaling,
                                               local_rotation, barycentric, face_scaling_remembered, (float*)frame_ws,
                                               out_xyz, out_scaling, out_rotation, s);
    }
    GG_AFTER("mesh_bind_forward");
    return 0;
}

int gg_mesh_bind_forward(int32_t num_vertices, int32_t num_faces, int32_t num_gaussians, const float* verts,
                         const int32_t* faces, const int32_t* binding, const float* local_xyz,
                         const float* local_log_scaling, const float* local_rotation, void* frame_ws, float* out_xyz,
                         float* out_scaling, float* out_rotation, int device, void* stream) {
    return gg_mesh_bind_forward_ex(num_vertices, num_faces, num_gaussians, verts, faces, binding, local_xyz,
                                   local_log_scaling, local_rotation, nullptr, nullptr, frame_ws, out_xyz, out_scaling,
                                   out_rotation, device, stream);
}

int gg_mesh_bind_backward_ex(int32_t num_vertices, int32_t num_faces, int32_t num_gaussians, const float* verts,
                             const int32_t* faces, const int32_t* binding, const float* local_xyz,
                             const float* local_log_scaling, const float* local_rotation, const float* barycentric,
                             const float* face_scaling_remembered, const void* frame_ws, void* frame_grad_ws,
                             const float* dL_dxyz, const float* dL_dscaling, const float* dL_drotation, float* dL_dverts,
                             float* dL_dlocal_xyz, float* dL_dlocal_log_scaling, float* dL_dlocal_rotation, int device,
                             void* stream) {
    if (num_vertices < 0 || num_faces < 0 || num_gaussians < 0) return fail(GG_E_BADARG, "negative size");
    if (num_gaussians == 0) {
        if (dL_dverts && num_vertices > 0) {
            GG_CUDA(cudaSetDevice(device));
            GG_CUDA(cudaMemsetAsync(dL_dverts, 0, (size_t)num_vertices * 3 * sizeof(float), (cudaStream_t)stream));
        }
        return 0;
    }
    if (!verts || !faces || !binding || !local_xyz || !local_log_scaling || !local_rotation || !frame_ws)
        return fail(GG_E_BADARG, "NULL argument");
    if (dL_dverts && !frame_grad_ws) return fail(GG_E_BADARG, "frame_grad_ws is required for dL_dverts");
    if (!aligned16(local_rotation) || (dL_drotation && !aligned16(dL_drotation)) ||
        (dL_dlocal_rotation && !aligned16(dL_dlocal_rotation)))
        return fail(GG_E_ALIGN, "rotation tensors not 16-byte aligned");
    GG_CUDA(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream;
    const gg_view* view = nullptr;
    if (dL_dverts) {
        GG_CUDA(cudaMemsetAsync(frame_grad_ws, 0, (size_t)num_faces * 17 * sizeof(float), s));
        GG_CUDA(cudaMemsetAsync(dL_dverts, 0, (size_t)num_vertices * 3 * sizeof(float), s));
    }
    {
        ScopedKernelTimer kt(K_MESHBWD, s);
        g_launches += launch_mesh_bind_backward(num_faces, num_gaussians, verts, faces, binding, local_xyz, local_log_scaling,
                                                local_rotation, barycentric, face_scaling_remembered, (const float*)frame_ws,
                                                dL_dxyz, dL_dscaling, dL_drotation, (float*)frame_grad_ws, dL_dverts,
                                                dL_dlocal_xyz, dL_dlocal_log_scaling, dL_dlocal_rotation, s);
    }
    GG_AFTER("mesh_bind_backward");
    return 0;
}

int gg_mesh_bind_backward(int32_t num_vertices, int32_t num_faces, int32_t num_gaussians, const float* verts,
                          const int32_t* faces, const int32_t* binding, const float* local_xyz,
                          const float* local_log_scaling, const float* local_rotation, const void* frame_ws,
                          void* frame_grad_ws, const float* dL_dxyz, const float* dL_dscaling, const float* dL_drotation,
                          float* dL_dverts, float* dL_dlocal_xyz, float* dL_dlocal_log_scaling, float* dL_dlocal_rotation,
                          int device, void* stream) {
    return gg_mesh_bind_backward_ex(num_vertices, num_faces, num_gaussians, verts, faces, binding, local_xyz,
                                    local_log_scaling, local_rotation, nullptr, nullptr, frame_ws, frame_grad_ws, dL_dxyz,
                                    dL_dscaling, dL_drotation, dL_dverts, dL_dlocal_xyz, dL_dlocal_log_scaling,
                                    dL_dlocal_rotation, device, stream);
}

// ---- on-device visibility ray cast -------------------------------------------
int gg_cast_rays_workspace_bytes(int32_t num_vertices, int32_t num_faces, size_t* ws_bytes, int64_t* list_capacity) {
    if (num_vertices < 0 || num_faces < 0) return fail(GG_E_BADARG, "negative size");
    const int64_t cap = 32ll * num_faces + 65536;      // cell-list entries; more than this -> brute-force kernel
    if (list_capacity) *list_capacity = cap;
    if (ws_bytes) *ws_bytes = vis_workspace_bytes(num_vertices, cap);
    return 0;
}

int gg_cast_rays_from_point(int32_t num_vertices, int32_t num_faces, int32_t num_rays, const float* verts,
                            const int32_t* faces, const float* targets, const float* origin, const float* look_at,
                            void* ws, int64_t list_capacity, int32_t force_bruteforce, int32_t* primitive_ids,
                            float* t_hit, int device, void* stream) {
    if (num_vertices < 0 || num_faces < 0 || num_rays < 0) return fail(GG_E_BADARG, "negative size");
    if (num_rays == 0) return 0;
    if (!targets || !origin || !look_at || !ws || !primitive_ids) return fail(GG_E_BADARG, "NULL argument");
    if (num_faces > 0 && (!verts || !faces)) return fail(GG_E_BADARG, "NULL mesh");
    if (list_capacity < 0 || list_capacity > 0xfffffff0ll) return fail(GG_E_BADARG, "list_capacity out of range");
    GG_CUDA(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream;
    const gg_view* view = nullptr;
    {
        ScopedKernelTimer kt(K_RAYCAST, s);
        g_launches += launch_cast_rays(num_vertices, num_faces, num_rays, verts, faces, targets, origin, look_at, ws,
                                       list_capacity, force_bruteforce, primitive_ids, t_hit, s);
    }
    GG_AFTER("cast_rays");
    return 0;
}

// ---- in-switch gradient average ------------------------------------------------------
int gg_nvls_allreduce_f32(void* multicast_base, const void* signal_pads_dev, int32_t rank, int32_t world_size,
                          int64_t elem_offset, int64_t elem_count, float scale, int32_t pad_slot0, int32_t num_blocks,
                          int device, void* stream) {
    if (!multicast_base || !signal_pads_dev) return fail(GG_E_BADARG, "NULL multicast pointer / signal pads");
    if (world_size < 1 || rank < 0 || rank >= world_size) return fail(GG_E_BADARG, "bad rank / world_size");
    if (world_size > 32) return fail(GG_E_BADARG, "world_size > 32 not supported");
    if (elem_offset < 0 || elem_count < 0 || (elem_offset & 3) || (elem_count & 3))
        return fail(GG_E_ALIGN, "elem_offset and elem_count must be multiples of 4 floats");
    if (num_blocks < 1 || num_blocks > 148 || pad_slot0 < 0) return fail(GG_E_BADARG, "bad num_blocks / pad_slot0");
    float* mc = reinterpret_cast<float*>(multicast_base) + elem_offset;
    if (!aligned16(mc)) return fail(GG_E_ALIGN, "multicast pointer not 16-byte aligned");
    GG_CUDA(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream;
    const gg_view* view = nullptr;
    g_launches += launch_nvls_allreduce(mc, reinterpret_cast<uint32_t* const*>(signal_pads_dev), rank, world_size, pad_slot0,
                                        elem_count / 4, scale, num_blocks, s);
    GG_AFTER("nvls_allreduce_kernel");
    return 0;
}

// ---- fused photometric loss --------------------------------------------------
int gg_photometric_workspace_bytes(int32_t width, int32_t height, size_t* map_bytes) {
    if (width < 0 || height < 0) return fail(GG_E_BADARG, "negative size");
    // two double accumulators + three partial-derivative maps [3,H,W] (the maps are untouched when with_ssim = 0)
    if (map_bytes) *map_bytes = 3 * align_up((size_t)3 * width * height * sizeof(float)) + 1024;
    return 0;
}

static void photometric_carve(void* ws, int W, int H, float** m1, float** m2, float** m3, double** sums) {
    const size_t plane = align_up((size_t)3 * W * H * sizeof(float));
    char* b = (char*)ws;
    *sums = (double*)b;                      // 2 x 64 double slots
    *m1 = (float*)(b + 1024);
    *m2 = (float*)(b + 1024 + plane);
    *m3 = (float*)(b + 1024 + 2 * plane);
}

int gg_photometric_forward(int32_t width, int32_t height, const float* image, const float* gt, const float* mask,
                           void* map_ws, int32_t with_ssim, int device, void* stream) {
    if (width <= 0 || height <= 0 || !image || !gt || !map_ws) return fail(GG_E_BADARG, "bad size or NULL argument");
    GG_CUDA(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream;
    const gg_view* view = nullptr;
    float *m1, *m2, *m3;
    double* sums;
    photometric_carve(map_ws, width, height, &m1, &m2, &m3, &sums);
    GG_CUDA(cudaMemsetAsync(sums, 0, 1024, s));
    {
        ScopedKernelTimer kt(K_LOSSFWD, s);
        if (!with_ssim) m1 = m2 = m3 = nullptr;      // L1 only: no partial-derivative maps
        g_launches += launch_photometric_fwd(width, height, image, gt, mask, m1, m2, m3, sums, s);
    }
    GG_AFTER("photometric_fwd_kernel");
    return 0;
}

// L1-only loss (lambda_dssim = 0) against an 8-bit ground truth [3,H,W], value / 255 (frames as stored / shipped over PCIe).
// dL_dimage == NULL: forward (accumulates into map_ws like gg_photometric_forward with with_ssim = 0);
// dL_dimage != NULL: backward (like gg_photometric_backward with coeff_ssim = 0).
int gg_photometric_l1_u8(int32_t width, int32_t height, const float* image, const uint8_t* gt_u8, const float* mask,
                         void* map_ws, float coeff_l1, const float* upstream_scalar, float* dL_dimage, int device,
                         void* stream) {
    if (width <= 0 || height <= 0 || !image || !gt_u8 || !map_ws) return fail(GG_E_BADARG, "bad size or NULL argument");
    const size_t plane = (size_t)width * height;
    if (plane % 4 != 0 || !aligned16(image) || (mask && !aligned16(mask)) || (dL_dimage && !aligned16(dL_dimage)) ||
        (reinterpret_cast<uintptr_t>(gt_u8) & 3u))
        return fail(GG_E_ALIGN, "gg_photometric_l1_u8 needs W*H % 4 == 0 and 16-byte aligned float tensors");
    GG_CUDA(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream;
    const gg_view* view = nullptr;
    if (!dL_dimage) GG_CUDA(cudaMemsetAsync(map_ws, 0, 1024, s));
    {
        ScopedKernelTimer kt(dL_dimage ? K_LOSSBWD : K_LOSSFWD, s);
        g_launches += launch_photometric_l1_u8(width, height, image, gt_u8, mask, (double*)map_ws, coeff_l1, upstream_scalar,
                                               dL_dimage, s);
    }
    GG_AFTER("photometric_l1_u8");
    return 0;
}

int gg_photometric_reduce(int32_t width, int32_t height, const void* map_ws, float lambda_dssim, float* out3, int device,
                          void* stream) {
    if (width <= 0 || height <= 0 || !map_ws || !out3) return fail(GG_E_BADARG, "bad size or NULL argument");
    GG_CUDA(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream;
    const gg_view* view = nullptr;
    g_launches += launch_photometric_finalize((const double*)map_ws, 1.0 / (3.0 * (double)width * (double)height),
                                              lambda_dssim, out3, s);
    GG_AFTER("photometric_finalize_kernel");
    return 0;
}

int gg_photometric_backward(int32_t width, int32_t height, const float* image, const float* gt, const float* mask,
                            const void* map_ws, float coeff_l1, float coeff_ssim, const float* upstream_scalar,
                            float* dL_dimage, int device, void* stream) {
    if (width <= 0 || height <= 0 || !image || !gt || !map_ws || !dL_dimage) return fail(GG_E_BADARG, "bad size or NULL argument");
    GG_CUDA(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream;
    const gg_view* view = nullptr;
    float *m1, *m2, *m3;
    double* sums;
    photometric_carve(const_cast<void*>(map_ws), width, height, &m1, &m2, &m3, &sums);
    if (coeff_ssim == 0.f) m1 = m2 = m3 = nullptr;       // forward was L1 only
    {
        ScopedKernelTimer kt(K_LOSSBWD, s);
        g_launches += launch_photometric_bwd(width, height, image, gt, mask, m1, m2, m3, coeff_l1, coeff_ssim, upstream_scalar, dL_dimage, s);
    }
    GG_AFTER("photometric_bwd_kernel");
    return 0;
}

}  // extern "C"
