// api.cu — the extern "C" vxrt_cuda_* layer (include/vxrt_cuda.h).  Never throws, never aborts;
// every entry point validates its arguments and returns a vxrt_status.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include <vector>
#include <unordered_map>

#include "ctx.h"

static thread_local char g_err[512] = "";

int vxrt_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
int vxrt_check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return VXRT_OK;
    return vxrt_fail(VXRT_E_CUDA, "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
}

// Every entry point runs with the context's device current and restores the caller's device on return, so contexts on
// different GPUs can live in one process, be called from any thread, and survive a cudaSetDevice / torch.cuda.set_device
// of the caller in between.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
// lane 1 of the pass-level concurrency (ctx.h): `stream` waits for everything queued there
static int vxrt_join_lane1(vxrt_ctx* c, bool keep_fork) {
    if (c->lane1_pending) {
        VX_CUDA(cudaStreamWaitEvent(c->stream, c->lane1_tail, 0));
        c->lane1_pending = false; c->lane1_reads_gi = false;
    }
    if (c->lane2_pending) {
        VX_CUDA(cudaStreamWaitEvent(c->stream, c->lane2_tail, 0));
        c->lane2_pending = false;
    }
    c->lane1_pre_refl_valid = false;
    if (!keep_fork) c->gi_fork_valid = false;
    return VXRT_OK;
}
// Every entry point joins lane 1 first, except the ray passes that know about the lanes (REQUIRE_CTX_LANES).  Calls that only read what
// passes have produced (read-backs, copies, statistics) keep the frame's fork point (REQUIRE_CTX_READER): a lane-1 pass issued after them
// still waits only for what preceded the GI.  Anything else may produce inputs of a later lane-1 pass, so it drops the fork point and
// that pass waits for everything queued on lane 0.
#define REQUIRE_CTX_LANES(c)                                                    \
    if (!(c)) return vxrt_fail(VXRT_E_INVALID, "%s: ctx is NULL", __func__);    \
    DeviceGuard _device_guard((c)->device)
#define REQUIRE_CTX(c)                                                          \
    REQUIRE_CTX_LANES(c);                                                       \
    if ((c)->lane1_pending || (c)->lane2_pending || (c)->gi_fork_valid) { if (int _jrc = vxrt_join_lane1(c, false)) return _jrc; }
#define REQUIRE_CTX_READER(c)                                                   \
    REQUIRE_CTX_LANES(c);                                                       \
    if ((c)->lane1_pending || (c)->lane2_pending) { if (int _jrc = vxrt_join_lane1(c, true)) return _jrc; }

static bool lanes_on(const vxrt_ctx* c) { return c->pass_overlap && !c->probe_on && !(c->trace_caps | c->trace_spill); }
static int ensure_lanes(vxrt_ctx* c) {
    if (c->lane1) return VXRT_OK;
    for (cudaEvent_t* e : {&c->gi_fork, &c->gi_done, &c->lane0_mark, &c->lane1_tail, &c->lane1_pre_refl, &c->lane2_tail})
        VX_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    VX_CUDA(cudaStreamCreateWithFlags(&c->lane2, cudaStreamNonBlocking));
    VX_CUDA(cudaStreamCreateWithFlags(&c->lane1, cudaStreamNonBlocking));   // last: c->lane1 != nullptr means all of the above exist
    return VXRT_OK;
}
// a lane-1 pass (sun-shadow trace, direct term): waits for what lane 0 had queued before the frame's diffuse_trace (or for everything, if there
// was none), runs with c->stream swapped, leaves its tail event behind
struct Lane1Pass {
    vxrt_ctx* c; cudaStream_t saved; void* wf; size_t wf_cap; bool on, arena, to2 = false;
    explicit Lane1Pass(vxrt_ctx* c_, bool own_arena = false) : c(c_), saved(c_->stream), wf(c_->d_wf), wf_cap(c_->wf_cap), on(lanes_on(c_)), arena(own_arena) {}
    // reads_gi: the pass reads the GI attachments; gi_first: from its first kernel on (otherwise from the point where the launcher waits for
    // c->refl_gi_event).  gi_free: the pass reads neither the GI nor the reflections (the direct term), so it may overtake a pending
    // reflection pass on lane 2 (ctx.h).
    int begin(bool reads_gi = false, bool gi_first = false, bool gi_free = false) {
        if (!on) return VXRT_OK;
        if (int rc = ensure_lanes(c)) return rc;
        if (gi_free && c->lane2_direct && c->gi_fork_valid && c->lane1_pending && c->lane1_reads_gi && c->lane1_pre_refl_valid) {
            VX_CUDA(cudaStreamWaitEvent(c->lane2, c->gi_fork, 0));
            VX_CUDA(cudaStreamWaitEvent(c->lane2, c->lane1_pre_refl, 0));
            to2 = true;
            c->stream = c->lane2;
            return VXRT_OK;
        }
        if (c->lane2_pending) VX_CUDA(cudaStreamWaitEvent(c->lane1, c->lane2_tail, 0));   // lane 2 may still read what this pass rewrites
        if (c->gi_fork_valid) {
            VX_CUDA(cudaStreamWaitEvent(c->lane1, c->gi_fork, 0));
            if (reads_gi) {
                VX_CUDA(cudaEventRecord(c->lane1_pre_refl, c->lane1));   // everything lane 1 holds before this pass (and before its wait for the GI)
                c->lane1_pre_refl_valid = true;
                if (gi_first) VX_CUDA(cudaStreamWaitEvent(c->lane1, c->gi_done, 0));
                else c->refl_gi_event = c->gi_done;
            }
        } else {   // no diffuse_trace since the last join: everything lane 0 has queued (a GI issued before an intervening call included)
            VX_CUDA(cudaEventRecord(c->lane0_mark, c->stream));
            VX_CUDA(cudaStreamWaitEvent(c->lane1, c->lane0_mark, 0));
        }
        if (reads_gi) c->lane1_reads_gi = true;
        c->stream = c->lane1;
        if (arena) { c->d_wf = c->d_wf1; c->wf_cap = c->wf1_cap; }
        return VXRT_OK;
    }
    int end(int rc) {
        if (on && to2 && c->stream == c->lane2) {
            c->stream = saved;
            const cudaError_t e = cudaEventRecord(c->lane2_tail, c->lane2);
            c->lane2_pending = true;
            if (rc == VXRT_OK && e != cudaSuccess) return vxrt_check_cuda(e, "cudaEventRecord(lane2_tail)");
            return rc;
        }
        if (!on || c->stream != c->lane1) return rc;
        if (arena) { c->d_wf1 = c->d_wf; c->wf1_cap = c->wf_cap; c->d_wf = wf; c->wf_cap = wf_cap; }
        c->refl_gi_event = nullptr;
        c->stream = saved;
        const cudaError_t e = cudaEventRecord(c->lane1_tail, c->lane1);
        c->lane1_pending = true;
        if (rc == VXRT_OK && e != cudaSuccess) return vxrt_check_cuda(e, "cudaEventRecord(lane1_tail)");
        return rc;
    }
};
#define REQUIRE_PTR(p) \
    if (!(p)) return vxrt_fail(VXRT_E_INVALID, "%s: %s is NULL", __func__, #p)

static int sync_copy_streams(vxrt_ctx* c) {
    if (c->copy_stream) VX_CUDA(cudaStreamSynchronize(c->copy_stream));
    for (cudaStream_t cs : c->copy_stream_lane) if (cs) VX_CUDA(cudaStreamSynchronize(cs));
    return VXRT_OK;
}
// Orders a copy of attachment `id` behind the pass that produced it and returns the copy stream to use.  With copy_lanes an attachment that a
// pending lane-1 / lane-2 pass wrote waits for that lane alone and goes to that lane's own copy stream; an attachment of lane 0 waits for lane 0
// alone.  Without it the lanes are joined into `stream` first (keeping the frame's fork point), as every reader does.
static int begin_copy(vxrt_ctx* c, int id, cudaStream_t* cs_out) {
    if (!c->copy_stream) VX_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    if (!c->att_ready[id]) {
        VX_CUDA(cudaEventCreateWithFlags(&c->att_ready[id], cudaEventDisableTiming));
        VX_CUDA(cudaEventCreateWithFlags(&c->att_read_done[id], cudaEventDisableTiming));
    }
    cudaStream_t prod = c->stream, cs = c->copy_stream;
    int lane = 0;
    if (c->copy_lanes) {
        if (c->att_lane[id] == 1 && c->lane1_pending) { prod = c->lane1; lane = 1; }
        else if (c->att_lane[id] == 2 && c->lane2_pending) { prod = c->lane2; lane = 2; }
    } else if (c->lane1_pending || c->lane2_pending) {
        if (int jrc = vxrt_join_lane1(c, true)) return jrc;
    }
    if (lane) {
        cudaStream_t& ls = c->copy_stream_lane[lane - 1];
        if (!ls) VX_CUDA(cudaStreamCreateWithFlags(&ls, cudaStreamNonBlocking));
        cs = ls;
    }
    if (c->att_read_pending[id]) VX_CUDA(cudaStreamWaitEvent(cs, c->att_read_done[id], 0));   // an earlier copy of it may be on the other copy stream
    VX_CUDA(cudaEventRecord(c->att_ready[id], prod));               // everything issued there so far has produced the attachment
    VX_CUDA(cudaStreamWaitEvent(cs, c->att_ready[id], 0));
    *cs_out = cs;
    return VXRT_OK;
}
static int end_copy(vxrt_ctx* c, int id, cudaStream_t cs) {
    VX_CUDA(cudaEventRecord(c->att_read_done[id], cs));
    c->att_read_pending[id] = true;
    return VXRT_OK;
}
int vxrt_ensure_attachment(vxrt_ctx* c, int id, int w, int h, int bpp) {
    if (id < 0 || id >= VXRT_ATT_COUNT) return vxrt_fail(VXRT_E_INVALID, "bad attachment id %d", id);
    if (w <= 0 || h <= 0 || w > 16384 || h > 16384) return vxrt_fail(VXRT_E_INVALID, "bad attachment size %dx%d", w, h);
    Attachment& a = c->att[id];
    size_t need = (size_t)w * h * bpp;
    if (c->att_read_pending[id]) {  // the pass about to write this attachment must not overtake a read-back in flight
        VX_CUDA(cudaStreamWaitEvent(c->stream, c->att_read_done[id], 0));
        c->att_read_pending[id] = false;
    }
    if (need > a.capacity && a.external)
        return vxrt_fail(VXRT_E_INVALID, "attachment %d: bound storage holds %zu bytes, the pass needs %zu", id, a.capacity, need);
    c->att_lane[id] = (c->lane1 && c->stream == c->lane1) ? 1 : (c->lane2 && c->stream == c->lane2) ? 2 : 0;
    if (need > a.capacity) {
        if (int src = sync_copy_streams(c)) return src;
        if (a.ptr) VX_CUDA(cudaFree(a.ptr));
        a.ptr = nullptr;
        a.capacity = 0;
        VX_CUDA(cudaMalloc(&a.ptr, need));
        a.capacity = need;
    }
    if (a.width != w || a.height != h || a.bpp != bpp) {
        // a tile-sharded pass only writes its rows: keep the rest defined
        VX_CUDA(cudaMemsetAsync(a.ptr, 0, need, c->stream));
    }
    a.width = w; a.height = h; a.bpp = bpp;
    return VXRT_OK;
}

// The grids (37.7 MB) are what every traversal iteration gathers from, while the wavefront passes stream ~0.5 GB of
// path state through the 126 MB L2 per frame.  An access-policy window marks the grids as persisting and everything
// else a kernel of this stream touches as streaming, so the state traffic does not evict them.
// Measured (config 4 / config 5, same box): primary 0.099 -> 0.096 ms, but GI 1.088 -> 1.150 ms and 7.34 -> 7.88 ms:
// the 37.7 MB set-aside costs the path-state streams more L2 than the grids gain.  Off by default.
int vxrt_apply_l2_policy(vxrt_ctx* c) {
    if (!c->d_df || c->l2_persist_max == 0) return VXRT_OK;
    cudaStreamAttrValue v = {};
    if (c->l2_persist) {
        size_t want = 2 * c->nvox;
        if (want > c->l2_persist_max) want = c->l2_persist_max;
        VX_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
        v.accessPolicyWindow.base_ptr = c->d_df;
        v.accessPolicyWindow.num_bytes = 2 * c->nvox;
        v.accessPolicyWindow.hitRatio = (float)((double)want / (double)(2 * c->nvox));
        v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    } else {
        v.accessPolicyWindow.num_bytes = 0;  // disables the window
        VX_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0));  // and gives the set-aside back
    }
    VX_CUDA(cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &v));
    return VXRT_OK;
}

int vxrt_ensure_trace_cont(vxrt_ctx* c, size_t rays, TraceCont* out) {
    if (rays > c->trace_cont_cap) {
        VX_CUDA(cudaStreamSynchronize(c->stream));
        if (c->d_trace_cont) VX_CUDA(cudaFree(c->d_trace_cont));
        c->d_trace_cont = nullptr; c->trace_cont_cap = 0;
        VX_CUDA(cudaMalloc(&c->d_trace_cont, 2 * rays * (sizeof(float4) + sizeof(unsigned)) + 256));
        c->trace_cont_cap = rays;
    }
    uint8_t* p = (uint8_t*)c->d_trace_cont;
    out->count = reinterpret_cast<int*>(p);
    out->q[0] = reinterpret_cast<float4*>(p + 256);
    out->q[1] = out->q[0] + c->trace_cont_cap;
    out->meta[0] = reinterpret_cast<unsigned*>(out->q[1] + c->trace_cont_cap);
    out->meta[1] = out->meta[0] + c->trace_cont_cap;
    return VXRT_OK;
}

cudaEvent_t vxrt_probe_event(vxrt_ctx* c) {
    if (!c->probe_on) return nullptr;
    if (c->probe_used == c->probe_ev.size()) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
        c->probe_ev.push_back(e);
    }
    return c->probe_ev[c->probe_used++];
}

extern "C" {

const char* vxrt_cuda_last_error(void) { return g_err; }

int vxrt_cuda_create(vxrt_ctx** out, int device, const int32_t* dims) {
    REQUIRE_PTR(out);
    *out = nullptr;
    int nx = VXRT_WORLD_SIZE_X, ny = VXRT_WORLD_SIZE_Y, nz = VXRT_WORLD_SIZE_Z;
    if (dims) { nx = dims[0]; ny = dims[1]; nz = dims[2]; }
    if (nx < 16 || ny < 16 || nz < 16 || nx > 1024 || ny > 1024 || nz > 1024 || (nx % 16) != 0 ||
        (size_t)nx * ny > 65536)
        return vxrt_fail(VXRT_E_INVALID, "unsupported grid dims %dx%dx%d", nx, ny, nz);
    int ndev = 0;
    VX_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return vxrt_fail(VXRT_E_INVALID, "device %d out of range (%d devices)", device, ndev);
    DeviceGuard _device_guard(device);  // the caller's current device is restored on return
    VX_CUDA(cudaSetDevice(device));
    vxrt_ctx* c = new (std::nothrow) vxrt_ctx();
    if (!c) return vxrt_fail(VXRT_E_NOMEM, "out of host memory");
    c->device = device;
    if (const char* e = getenv("VXRT_WAVEFRONT")) c->wavefront = atoi(e) != 0;
    c->nx = nx; c->ny = ny; c->nz = nz;
    c->nvox = (size_t)nx * ny * nz;
    cudaDeviceProp prop;
    int rc = vxrt_check_cuda(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
    if (rc == VXRT_OK && prop.major < 10) rc = vxrt_fail(VXRT_E_UNSUPPORTED, "sm_%d%d device; this library is built for sm_100a only", prop.major, prop.minor);
    if (rc == VXRT_OK) { c->sm_count = prop.multiProcessorCount; rc = vxrt_check_cuda(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking), "cudaStreamCreate"); }
    // one allocation for both grids (distance field | block ids) so that a single L2 access-policy window covers them
    if (rc == VXRT_OK) { c->stream = c->own_stream; rc = vxrt_check_cuda(cudaMalloc(&c->d_df, 2 * c->nvox), "cudaMalloc(grids)"); }
    if (rc == VXRT_OK) {
        c->d_blocks = c->d_df + c->nvox;
        c->l2_persist_max = (size_t)prop.persistingL2CacheMaxSize;
        if (const char* e = getenv("VXRT_L2_PERSIST")) c->l2_persist = atoi(e) != 0;
        rc = vxrt_apply_l2_policy(c);
    }
    if (rc == VXRT_OK) rc = vxrt_check_cuda(cudaMalloc(&c->d_block_data, 6 * 128 * sizeof(int32_t)), "cudaMalloc(block_data)");
    if (rc == VXRT_OK) rc = vxrt_check_cuda(cudaMalloc(&c->d_stats, 2 * sizeof(TraceStatsDev)), "cudaMalloc(stats)");
    if (rc == VXRT_OK) rc = vxrt_check_cuda(cudaMemset(c->d_stats, 0, 2 * sizeof(TraceStatsDev)), "cudaMemset(stats)");
    if (rc == VXRT_OK) rc = vxrt_check_cuda(cudaMemset(c->d_block_data, 0xff, 6 * 128 * sizeof(int32_t)), "cudaMemset(block_data)");
    if (rc != VXRT_OK) { vxrt_cuda_destroy(c); return rc; }
    *out = c;
    return VXRT_OK;
}

int vxrt_cuda_destroy(vxrt_ctx* c) {
    if (!c) return VXRT_OK;
    DeviceGuard _device_guard(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->d_df); cudaFree(c->d_block_data); cudaFree(c->d_blue_noise);
    cudaFree(c->d_blue_tex); cudaFree(c->d_edit_buf); cudaFree(c->d_stats); cudaFree(c->d_sky); cudaFree(c->d_slab_z0); cudaFree(c->d_wf); cudaFree(c->d_ray_buf);
    cudaFree(c->d_lpv); cudaFree(c->d_lpv_work); cudaFree(c->d_trace_cont);
    if (c->h_lpv_flag) cudaFreeHost(c->h_lpv_flag);
    cudaFree(c->d_lpv_avg);
    for (int k = 0; k < 4; ++k) { cudaFree(c->d_tex_data[k]); cudaFree(c->d_tex_decode[k]); }
    for (int i = 0; i < VXRT_ATT_COUNT; ++i)
        if (!c->att[i].external) cudaFree(c->att[i].ptr);
    for (cudaEvent_t e : c->probe_ev) cudaEventDestroy(e);
    for (cudaStream_t cs : {c->copy_stream, c->copy_stream_lane[0], c->copy_stream_lane[1]}) if (cs) { cudaStreamSynchronize(cs); cudaStreamDestroy(cs); }
    if (c->aux_stream) { cudaStreamSynchronize(c->aux_stream); cudaStreamDestroy(c->aux_stream); }
    for (cudaStream_t ls : {c->lane1, c->lane2}) if (ls) { cudaStreamSynchronize(ls); cudaStreamDestroy(ls); }
    for (cudaEvent_t e : {c->gi_fork, c->gi_done, c->lane0_mark, c->lane1_tail, c->lane1_pre_refl, c->lane2_tail}) if (e) cudaEventDestroy(e);
    cudaFree(c->d_wf1);
    for (int l = 0; l < 2; ++l) {
        for (auto& b : c->band[l]) {
            if (b.stream) { cudaStreamSynchronize(b.stream); cudaStreamDestroy(b.stream); }
            if (b.aux) { cudaStreamSynchronize(b.aux); cudaStreamDestroy(b.aux); }
            for (cudaEvent_t e : {b.done, b.aux_fork, b.aux_join}) if (e) cudaEventDestroy(e);
            cudaFree(b.wf);
        }
        if (c->band_fork[l]) cudaEventDestroy(c->band_fork[l]);
    }
    if (c->aux_fork) cudaEventDestroy(c->aux_fork);
    if (c->aux_join) cudaEventDestroy(c->aux_join);
    if (c->copies_joined) cudaEventDestroy(c->copies_joined);
    for (int i = 0; i < VXRT_ATT_COUNT; ++i) {
        if (c->att_ready[i]) cudaEventDestroy(c->att_ready[i]);
        if (c->att_read_done[i]) cudaEventDestroy(c->att_read_done[i]);
    }
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
    return VXRT_OK;
}

int vxrt_cuda_set_stream(vxrt_ctx* c, void* s) {
    REQUIRE_CTX(c);
    cudaStream_t next = s ? (cudaStream_t)s : c->own_stream;
    if (next == c->stream) return VXRT_OK;
    VX_CUDA(cudaStreamSynchronize(c->stream));
    c->stream = next;
    return vxrt_apply_l2_policy(c);
}
int vxrt_cuda_synchronize(vxrt_ctx* c) {
    REQUIRE_CTX_READER(c);
    VX_CUDA(cudaStreamSynchronize(c->stream));
    return VXRT_OK;
}
int vxrt_cuda_set_option(vxrt_ctx* c, const char* name, int32_t value) {
    REQUIRE_CTX(c); REQUIRE_PTR(name);
    if (!strcmp(name, "wavefront")) { c->wavefront = value != 0; return VXRT_OK; }
    if (!strcmp(name, "probe")) { c->probe_on = value != 0; return VXRT_OK; }
    if (!strcmp(name, "l2_persist")) { c->l2_persist = value != 0; return vxrt_apply_l2_policy(c); }
    if (!strcmp(name, "lpv_coop")) { c->lpv_coop = value != 0; return VXRT_OK; }
    if (!strcmp(name, "filter_snap")) {   // value / 65536: 256 = 1 / 256 (the weight resolution of the hardware texture unit), 0 = bit-faithful
        if (value < 0 || value > 16384) return vxrt_fail(VXRT_E_INVALID, "filter_snap %d outside [0, 16384] (units of 1 / 65536)", value);
        c->filter_snap = (float)value / 65536.0f;
        return VXRT_OK;
    }
    if (!strcmp(name, "gi_fuse_final")) { c->gi_fuse_final = value != 0; return VXRT_OK; }
    if (!strcmp(name, "gi_overlap")) { c->gi_overlap = value != 0; return VXRT_OK; }
    if (!strcmp(name, "pass_overlap")) { c->pass_overlap = value != 0; return VXRT_OK; }
    if (!strcmp(name, "lane2_direct")) { c->lane2_direct = value != 0; return VXRT_OK; }
    if (!strcmp(name, "lane1_gbuffer")) { c->lane1_gbuffer = value != 0; return VXRT_OK; }
    if (!strcmp(name, "refl_defer_gi")) { c->refl_defer_gi = value != 0; return VXRT_OK; }
    if (!strcmp(name, "copy_lanes")) { c->copy_lanes = value != 0; return VXRT_OK; }
    if (!strcmp(name, "wf_bands")) { if (value < 1 || value > 4) return vxrt_fail(VXRT_E_INVALID, "wf_bands: 1..4"); c->wf_bands = value; return VXRT_OK; }
    if (!strcmp(name, "trace_caps")) { c->trace_caps = value & 0xffffff; return VXRT_OK; }
    if (!strcmp(name, "trace_spill")) { c->trace_spill = value & 0xffffff; return VXRT_OK; }
    if (!strcmp(name, "df_stage")) { c->df_stage = value; return VXRT_OK; }
    if (!strcmp(name, "df_dbg")) { c->df_dbg = value; return VXRT_OK; }
    if (!strcmp(name, "df_zver")) { c->df_zver = value; return VXRT_OK; }
    if (!strcmp(name, "df_xyver")) { c->df_xyver = value; return VXRT_OK; }
    if (!strcmp(name, "df_sx")) { c->df_sx = value < 0 ? 0 : (value > 8 ? 8 : value); return VXRT_OK; }
    if (!strcmp(name, "df_sy")) { c->df_sy = value < 0 ? 0 : (value > 8 ? 8 : value); return VXRT_OK; }
    return vxrt_fail(VXRT_E_INVALID, "unknown option '%s'", name);
}
int64_t vxrt_cuda_launch_count(vxrt_ctx* c) { return c ? c->launches : -1; }

int vxrt_cuda_upload_world(vxrt_ctx* c, const uint8_t* blocks) {
    REQUIRE_CTX(c); REQUIRE_PTR(blocks);
    VX_CUDA(cudaMemcpyAsync(c->d_blocks, blocks, c->nvox, cudaMemcpyHostToDevice, c->stream));
    VX_CUDA(cudaStreamSynchronize(c->stream));  // host buffer is only borrowed for the call
    c->world_uploaded = true;
    c->df_valid = false;
    return VXRT_OK;
}
int vxrt_cuda_download_world(vxrt_ctx* c, uint8_t* out) {
    REQUIRE_CTX(c); REQUIRE_PTR(out);
    if (!c->world_uploaded) return vxrt_fail(VXRT_E_STATE, "download_world before upload_world");
    VX_CUDA(cudaMemcpyAsync(out, c->d_blocks, c->nvox, cudaMemcpyDeviceToHost, c->stream));
    VX_CUDA(cudaStreamSynchronize(c->stream));
    return VXRT_OK;
}

int vxrt_cuda_edit_blocks(vxrt_ctx* c, const int32_t* e, int32_t n) {
    REQUIRE_CTX(c);
    if (n < 0) return vxrt_fail(VXRT_E_INVALID, "edit_blocks: n < 0");
    if (n == 0) return VXRT_OK;
    REQUIRE_PTR(e);
    if (!c->world_uploaded) return vxrt_fail(VXRT_E_STATE, "edit_blocks before upload_world");
    for (int i = 0; i < n; ++i) {
        int x = e[4 * i], y = e[4 * i + 1], z = e[4 * i + 2], id = e[4 * i + 3];
        if (x < 0 || y < 0 || z < 0 || x >= c->nx || y >= c->ny || z >= c->nz || id < 0 || id > 255)
            return vxrt_fail(VXRT_E_INVALID, "edit %d out of range: (%d,%d,%d) id %d", i, x, y, z, id);
    }
    // the reference applies edits one by one: a later edit of the same voxel wins.  Keep only the
    // last edit per voxel so the parallel scatter is deterministic.
    std::vector<int32_t> last;
    const int32_t* src = e;
    int m = n;
    {
        std::unordered_map<int64_t, int> seen;
        seen.reserve((size_t)n * 2);
        bool dup = false;
        for (int i = 0; i < n; ++i) {
            int64_t key = (int64_t)e[4 * i] + (int64_t)c->nx * (e[4 * i + 1] + (int64_t)c->ny * e[4 * i + 2]);
            auto it = seen.find(key);
            if (it != seen.end()) { dup = true; it->second = i; } else seen.emplace(key, i);
        }
        if (dup) {
            last.reserve(seen.size() * 4);
            for (int i = 0; i < n; ++i) {
                int64_t key = (int64_t)e[4 * i] + (int64_t)c->nx * (e[4 * i + 1] + (int64_t)c->ny * e[4 * i + 2]);
                if (seen[key] == i) last.insert(last.end(), e + 4 * i, e + 4 * i + 4);
            }
            src = last.data();
            m = (int)(last.size() / 4);
        }
    }
    size_t bytes = (size_t)m * 4 * sizeof(int32_t);
    if (bytes > c->edit_cap) {
        if (c->d_edit_buf) VX_CUDA(cudaFree(c->d_edit_buf));
        c->d_edit_buf = nullptr; c->edit_cap = 0;
        VX_CUDA(cudaMalloc(&c->d_edit_buf, bytes));
        c->edit_cap = bytes;
    }
    VX_CUDA(cudaMemcpyAsync(c->d_edit_buf, src, bytes, cudaMemcpyHostToDevice, c->stream));
    int rc = vxrt_launch_edit_blocks(c, c->d_edit_buf, m);
    if (rc) return rc;
    VX_CUDA(cudaStreamSynchronize(c->stream));  // src may be a temporary / borrowed buffer
    c->df_valid = false;
    return VXRT_OK;
}

int vxrt_cuda_generate_distance_field(vxrt_ctx* c) {
    REQUIRE_CTX(c);
    if (!c->world_uploaded) return vxrt_fail(VXRT_E_STATE, "generate_distance_field before upload_world");
    int rc = vxrt_launch_distance_field(c);
    if (rc) return rc;
    c->df_valid = true;
    return VXRT_OK;
}
int vxrt_cuda_download_distance_field(vxrt_ctx* c, uint8_t* out) {
    REQUIRE_CTX(c); REQUIRE_PTR(out);
    if (!c->df_valid) return vxrt_fail(VXRT_E_STATE, "distance field has not been generated");
    VX_CUDA(cudaMemcpyAsync(out, c->d_df, c->nvox, cudaMemcpyDeviceToHost, c->stream));
    VX_CUDA(cudaStreamSynchronize(c->stream));
    return VXRT_OK;
}
int vxrt_cuda_upload_distance_field(vxrt_ctx* c, const uint8_t* df) {
    REQUIRE_CTX(c); REQUIRE_PTR(df);
    VX_CUDA(cudaMemcpyAsync(c->d_df, df, c->nvox, cudaMemcpyHostToDevice, c->stream));
    VX_CUDA(cudaStreamSynchronize(c->stream));
    c->df_valid = true;
    return VXRT_OK;
}

static int check_slabs(vxrt_ctx* c, const char* fn, int32_t slab, int32_t nslabs, const int32_t* z0) {
    if (!z0) return vxrt_fail(VXRT_E_INVALID, "%s: slab_z0 is NULL", fn);
    if (nslabs < 1 || nslabs > 64 || slab < 0 || slab >= nslabs) return vxrt_fail(VXRT_E_INVALID, "%s: bad slab %d of %d", fn, slab, nslabs);
    if (z0[0] != 0 || z0[nslabs] != c->nz) return vxrt_fail(VXRT_E_INVALID, "%s: slabs must cover [0, %d)", fn, c->nz);
    for (int i = 0; i < nslabs; ++i)
        if (z0[i + 1] <= z0[i]) return vxrt_fail(VXRT_E_INVALID, "%s: slab %d is empty", fn, i);
    return VXRT_OK;
}
int vxrt_cuda_df_slab_phase_a(vxrt_ctx* c, int32_t slab, int32_t nslabs, const int32_t* z0) {
    REQUIRE_CTX(c);
    if (!c->world_uploaded) return vxrt_fail(VXRT_E_STATE, "df_slab_phase_a before upload_world");
    int rc = check_slabs(c, __func__, slab, nslabs, z0);
    if (rc) return rc;
    c->df_valid = false;
    return vxrt_launch_df_slab_phase_a(c, z0[slab], z0[slab + 1]);
}
int vxrt_cuda_df_slab_phase_b(vxrt_ctx* c, int32_t slab, int32_t nslabs, const int32_t* z0, const void* first_planes, const void* last_planes) {
    REQUIRE_CTX(c); REQUIRE_PTR(first_planes); REQUIRE_PTR(last_planes);
    int rc = check_slabs(c, __func__, slab, nslabs, z0);
    if (rc) return rc;
    if (!c->d_slab_z0) VX_CUDA(cudaMalloc(&c->d_slab_z0, 65 * sizeof(int32_t)));
    VX_CUDA(cudaMemcpyAsync(c->d_slab_z0, z0, (size_t)(nslabs + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    rc = vxrt_launch_df_slab_phase_b(c, slab, nslabs, c->d_slab_z0, z0[slab], z0[slab + 1], first_planes, last_planes);
    if (rc) return rc;
    VX_CUDA(cudaStreamSynchronize(c->stream));  // slab_z0 is borrowed
    return VXRT_OK;
}
int vxrt_cuda_df_plane_device(vxrt_ctx* c, int32_t z, void** p) {
    REQUIRE_CTX(c); REQUIRE_PTR(p);
    if (z < 0 || z >= c->nz) return vxrt_fail(VXRT_E_INVALID, "df_plane_device: plane %d out of range", z);
    *p = c->d_df + (size_t)z * c->nx * c->ny;
    return VXRT_OK;
}
int vxrt_cuda_grid_device(vxrt_ctx* c, void** blocks, void** df) {
    REQUIRE_CTX(c);
    if (blocks) *blocks = c->d_blocks;
    if (df) *df = c->d_df;
    return VXRT_OK;
}
int vxrt_cuda_df_commit(vxrt_ctx* c) {
    REQUIRE_CTX(c);
    if (!c->world_uploaded) return vxrt_fail(VXRT_E_STATE, "df_commit before upload_world");
    c->df_valid = true;
    return VXRT_OK;
}

int vxrt_cuda_set_block_data(vxrt_ctx* c, const int32_t* table) {
    REQUIRE_CTX(c); REQUIRE_PTR(table);
    VX_CUDA(cudaMemcpyAsync(c->d_block_data, table, 6 * 128 * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    VX_CUDA(cudaStreamSynchronize(c->stream));
    return VXRT_OK;
}
int vxrt_cuda_set_blue_noise(vxrt_ctx* c, const int32_t* data, int32_t count) {
    REQUIRE_CTX(c); REQUIRE_PTR(data);
    if (count != 256 * 256 + 2 * 128 * 128 * 8) return vxrt_fail(VXRT_E_INVALID, "blue noise table must hold 327680 ints, got %d", count);
    if (!c->d_blue_noise) VX_CUDA(cudaMalloc(&c->d_blue_noise, (size_t)count * sizeof(int32_t)));
    VX_CUDA(cudaMemcpyAsync(c->d_blue_noise, data, (size_t)count * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    VX_CUDA(cudaStreamSynchronize(c->stream));
    c->blue_noise_count = count;
    return VXRT_OK;
}
int vxrt_cuda_set_blue_noise_texture(vxrt_ctx* c, const uint8_t* rgba, int32_t w, int32_t h) {
    REQUIRE_CTX(c); REQUIRE_PTR(rgba);
    if (w <= 0 || h <= 0 || w > 4096 || h > 4096) return vxrt_fail(VXRT_E_INVALID, "bad blue-noise texture size %dx%d", w, h);
    if (c->d_blue_tex) { VX_CUDA(cudaFree(c->d_blue_tex)); c->d_blue_tex = nullptr; }
    VX_CUDA(cudaMalloc(&c->d_blue_tex, (size_t)w * h * 4));
    VX_CUDA(cudaMemcpyAsync(c->d_blue_tex, rgba, (size_t)w * h * 4, cudaMemcpyHostToDevice, c->stream));
    VX_CUDA(cudaStreamSynchronize(c->stream));
    c->blue_w = w; c->blue_h = h;
    return VXRT_OK;
}

int vxrt_cuda_read_attachment(vxrt_ctx* c, int32_t id, void* dst, size_t bytes) {
    REQUIRE_CTX_READER(c); REQUIRE_PTR(dst);
    if (id < 0 || id >= VXRT_ATT_COUNT) return vxrt_fail(VXRT_E_INVALID, "bad attachment id %d", id);
    const Attachment& a = c->att[id];
    if (!a.ptr) return vxrt_fail(VXRT_E_STATE, "attachment %d has not been rendered", id);
    size_t have = (size_t)a.width * a.height * a.bpp;
    if (bytes != have) return vxrt_fail(VXRT_E_INVALID, "attachment %d holds %zu bytes, caller asked for %zu", id, have, bytes);
    VX_CUDA(cudaMemcpyAsync(dst, a.ptr, have, cudaMemcpyDeviceToHost, c->stream));
    VX_CUDA(cudaStreamSynchronize(c->stream));
    return VXRT_OK;
}
int vxrt_cuda_write_attachment(vxrt_ctx* c, int32_t id, int32_t width, int32_t height, int32_t bytes_per_pixel, const void* src) {
    REQUIRE_CTX(c); REQUIRE_PTR(src);
    if (id < 0 || id >= VXRT_ATT_COUNT) return vxrt_fail(VXRT_E_INVALID, "bad attachment id %d", id);
    if (bytes_per_pixel < 1 || bytes_per_pixel > 16) return vxrt_fail(VXRT_E_INVALID, "write_attachment: %d bytes per pixel", bytes_per_pixel);
    int rc = vxrt_ensure_attachment(c, id, width, height, bytes_per_pixel);
    if (rc) return rc;
    VX_CUDA(cudaMemcpyAsync(c->att[id].ptr, src, (size_t)width * height * bytes_per_pixel, cudaMemcpyHostToDevice, c->stream));
    VX_CUDA(cudaStreamSynchronize(c->stream));  // src is only borrowed for the call
    return VXRT_OK;
}
int vxrt_cuda_read_attachment_async(vxrt_ctx* c, int32_t id, void* dst, size_t bytes) {
    REQUIRE_CTX_LANES(c); REQUIRE_PTR(dst);
    if (id < 0 || id >= VXRT_ATT_COUNT) return vxrt_fail(VXRT_E_INVALID, "bad attachment id %d", id);
    const Attachment& a = c->att[id];
    if (!a.ptr) return vxrt_fail(VXRT_E_STATE, "attachment %d has not been rendered", id);
    size_t have = (size_t)a.width * a.height * a.bpp;
    if (bytes != have) return vxrt_fail(VXRT_E_INVALID, "attachment %d holds %zu bytes, caller asked for %zu", id, have, bytes);
    cudaStream_t cs;
    if (int rc = begin_copy(c, id, &cs)) return rc;
    VX_CUDA(cudaMemcpyAsync(dst, a.ptr, have, cudaMemcpyDeviceToHost, cs));
    return end_copy(c, id, cs);
}
int vxrt_cuda_copy_attachment_rows_async(vxrt_ctx* c, int32_t id, int32_t row0, int32_t rows, void* dst) {
    REQUIRE_CTX_LANES(c); REQUIRE_PTR(dst);
    if (id < 0 || id >= VXRT_ATT_COUNT) return vxrt_fail(VXRT_E_INVALID, "bad attachment id %d", id);
    const Attachment& a = c->att[id];
    if (!a.ptr) return vxrt_fail(VXRT_E_STATE, "attachment %d has not been rendered", id);
    if (rows == 0) { row0 = 0; rows = a.height; }
    if (row0 < 0 || rows < 0 || row0 + rows > a.height) return vxrt_fail(VXRT_E_INVALID, "copy_attachment_rows: rows [%d,+%d) of %d", row0, rows, a.height);
    const size_t row_bytes = (size_t)a.width * a.bpp;
    cudaStream_t cs;
    if (int rc = begin_copy(c, id, &cs)) return rc;
    VX_CUDA(cudaMemcpyAsync(dst, (const uint8_t*)a.ptr + (size_t)row0 * row_bytes, (size_t)rows * row_bytes, cudaMemcpyDefault, cs));
    return end_copy(c, id, cs);
}
int vxrt_cuda_copy_attachment_rect_async(vxrt_ctx* c, int32_t id, int32_t row0, int32_t rows, int32_t col0, int32_t cols, void* dst_image) {
    REQUIRE_CTX_LANES(c); REQUIRE_PTR(dst_image);
    if (id < 0 || id >= VXRT_ATT_COUNT) return vxrt_fail(VXRT_E_INVALID, "bad attachment id %d", id);
    const Attachment& a = c->att[id];
    if (!a.ptr) return vxrt_fail(VXRT_E_STATE, "attachment %d has not been rendered", id);
    if (rows == 0) { row0 = 0; rows = a.height; }
    if (cols == 0) { col0 = 0; cols = a.width; }
    if (row0 < 0 || rows < 0 || row0 + rows > a.height || col0 < 0 || cols < 0 || col0 + cols > a.width)
        return vxrt_fail(VXRT_E_INVALID, "copy_attachment_rect: rows [%d,+%d) x columns [%d,+%d) of %dx%d", row0, rows, col0, cols, a.width, a.height);
    const size_t pitch = (size_t)a.width * a.bpp, off = (size_t)row0 * pitch + (size_t)col0 * a.bpp;
    cudaStream_t cs;
    if (int rc = begin_copy(c, id, &cs)) return rc;
    if (rows > 0 && cols == a.width)   // whole rows are one contiguous block: a linear copy instead of `rows` strided ones
        VX_CUDA(cudaMemcpyAsync((uint8_t*)dst_image + off, (const uint8_t*)a.ptr + off, (size_t)rows * pitch, cudaMemcpyDefault, cs));
    else if (rows > 0 && cols > 0)
        VX_CUDA(cudaMemcpy2DAsync((uint8_t*)dst_image + off, pitch, (const uint8_t*)a.ptr + off, pitch, (size_t)cols * a.bpp, (size_t)rows, cudaMemcpyDefault, cs));
    return end_copy(c, id, cs);
}
int vxrt_cuda_shared_alloc(vxrt_ctx* c, size_t bytes, void** dev_ptr, uint8_t handle[64]) {
    REQUIRE_CTX(c); REQUIRE_PTR(dev_ptr); REQUIRE_PTR(handle);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the ABI carries the IPC handle as 64 bytes");
    if (bytes == 0) return vxrt_fail(VXRT_E_INVALID, "shared_alloc: 0 bytes");
    void* p = nullptr;
    VX_CUDA(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return vxrt_check_cuda(e, "cudaIpcGetMemHandle"); }
    memcpy(handle, &h, 64);
    *dev_ptr = p;
    return VXRT_OK;
}
int vxrt_cuda_shared_free(vxrt_ctx* c, void* dev_ptr) {
    REQUIRE_CTX(c);
    if (int src = sync_copy_streams(c)) return src;
    VX_CUDA(cudaFree(dev_ptr));
    return VXRT_OK;
}
int vxrt_cuda_shared_open(vxrt_ctx* c, const uint8_t handle[64], void** dev_ptr) {
    REQUIRE_CTX(c); REQUIRE_PTR(dev_ptr); REQUIRE_PTR(handle);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    VX_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return VXRT_OK;
}
int vxrt_cuda_shared_close(vxrt_ctx* c, void* dev_ptr) {
    REQUIRE_CTX(c);
    if (int src = sync_copy_streams(c)) return src;
    VX_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return VXRT_OK;
}
int vxrt_cuda_join_reads(vxrt_ctx* c) {
    REQUIRE_CTX_READER(c);
    if (!c->copy_stream) return VXRT_OK;
    if (!c->copies_joined) VX_CUDA(cudaEventCreateWithFlags(&c->copies_joined, cudaEventDisableTiming));
    for (cudaStream_t cs : {c->copy_stream, c->copy_stream_lane[0], c->copy_stream_lane[1]}) {   // a wait captures the event's state at the call, so one event serves both
        if (!cs) continue;
        VX_CUDA(cudaEventRecord(c->copies_joined, cs));
        VX_CUDA(cudaStreamWaitEvent(c->stream, c->copies_joined, 0));
    }
    return VXRT_OK;
}
int vxrt_cuda_wait_reads(vxrt_ctx* c) {
    REQUIRE_CTX_READER(c);
    return sync_copy_streams(c);
}
int vxrt_cuda_bind_attachment(vxrt_ctx* c, int32_t id, void* dev_ptr, size_t capacity) {
    REQUIRE_CTX(c);
    if (id < 0 || id >= VXRT_ATT_COUNT) return vxrt_fail(VXRT_E_INVALID, "bad attachment id %d", id);
    if (dev_ptr && capacity == 0) return vxrt_fail(VXRT_E_INVALID, "bind_attachment: capacity is 0");  // before anything is released
    if (int src = sync_copy_streams(c)) return src;
    c->att_read_pending[id] = false;
    Attachment& a = c->att[id];
    if (a.ptr && !a.external) {
        VX_CUDA(cudaStreamSynchronize(c->stream));  // passes in flight may still use the old storage
        void* old = a.ptr;
        a.ptr = nullptr; a.capacity = 0; a.width = a.height = a.bpp = 0;
        VX_CUDA(cudaFree(old));
    }
    if (dev_ptr) {
        a.ptr = dev_ptr; a.capacity = capacity; a.external = true;
    } else {
        a.ptr = nullptr; a.capacity = 0; a.external = false; a.width = a.height = a.bpp = 0;
    }
    return VXRT_OK;
}
int vxrt_cuda_attachment_device(vxrt_ctx* c, int32_t id, void** p, int32_t* w, int32_t* h, int32_t* bpp) {
    REQUIRE_CTX_READER(c);
    if (id < 0 || id >= VXRT_ATT_COUNT) return vxrt_fail(VXRT_E_INVALID, "bad attachment id %d", id);
    const Attachment& a = c->att[id];
    if (!a.ptr) return vxrt_fail(VXRT_E_STATE, "attachment %d has not been rendered", id);
    if (p) *p = a.ptr;
    if (w) *w = a.width;
    if (h) *h = a.height;
    if (bpp) *bpp = a.bpp;
    return VXRT_OK;
}

static int check_frame(const char* fn, int w, int h, const vxrt_tile& t, bool rows_only = false) {
    if (w <= 0 || h <= 0 || w > 16384 || h > 16384) return vxrt_fail(VXRT_E_INVALID, "%s: bad dimensions %dx%d", fn, w, h);
    if (t.rows < 0 || t.row0 < 0 || (t.rows > 0 && t.row0 >= h)) return vxrt_fail(VXRT_E_INVALID, "%s: bad tile rows [%d,+%d) of %d", fn, t.row0, t.rows, h);
    if (t.cols < 0 || t.col0 < 0 || (t.cols > 0 && t.col0 >= w)) return vxrt_fail(VXRT_E_INVALID, "%s: bad tile columns [%d,+%d) of %d", fn, t.col0, t.cols, w);
    if (rows_only && t.cols != 0) return vxrt_fail(VXRT_E_INVALID, "%s: the screen-space filters take row bands only (tile.cols must be 0)", fn);
    return VXRT_OK;
}

int vxrt_cuda_initial_trace(vxrt_ctx* c, const vxrt_primary_params* p) {
    REQUIRE_CTX(c); REQUIRE_PTR(p);
    if (!c->df_valid) return vxrt_fail(VXRT_E_STATE, "initial_trace needs a world and a distance field");
    int rc = check_frame(__func__, p->width, p->height, p->tile);
    if (rc) return rc;
    if (p->alpha_test && !c->tex_set[VXRT_TEX_ALBEDO])
        return vxrt_fail(VXRT_E_STATE, "alpha-tested traversal needs the albedo texture array (vxrt_cuda_set_texture_array) and the block table");
    if (p->alpha_test && !(p->fov > 0.0f && p->fov < 180.0f)) return vxrt_fail(VXRT_E_INVALID, "alpha test: fov must be in (0, 180) degrees");
    if (p->render_distance < 0) return vxrt_fail(VXRT_E_INVALID, "render_distance < 0");
    return vxrt_launch_initial_trace(c, *p);
}

int vxrt_cuda_trace_rays(vxrt_ctx* c, const float* origins, const float* directions, int32_t n, int32_t max_iterations, vxrt_ray_hit* hits) {
    REQUIRE_CTX(c);
    if (n < 0 || max_iterations < 0) return vxrt_fail(VXRT_E_INVALID, "trace_rays: n < 0 or max_iterations < 0");
    if (n == 0) return VXRT_OK;
    REQUIRE_PTR(origins); REQUIRE_PTR(directions); REQUIRE_PTR(hits);
    if (!c->df_valid) return vxrt_fail(VXRT_E_STATE, "trace_rays needs a world and a distance field");
    const size_t vec_bytes = ((size_t)n * 3 * sizeof(float) + 255) / 256 * 256, need = 2 * vec_bytes + (size_t)n * sizeof(vxrt_ray_hit);
    if (need > c->ray_cap) {
        if (c->d_ray_buf) VX_CUDA(cudaFree(c->d_ray_buf));
        c->d_ray_buf = nullptr; c->ray_cap = 0;
        VX_CUDA(cudaMalloc(&c->d_ray_buf, need));
        c->ray_cap = need;
    }
    char* base = (char*)c->d_ray_buf;
    float* d_o = (float*)base;
    float* d_d = (float*)(base + vec_bytes);
    vxrt_ray_hit* d_h = (vxrt_ray_hit*)(base + 2 * vec_bytes);
    VX_CUDA(cudaMemcpyAsync(d_o, origins, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    VX_CUDA(cudaMemcpyAsync(d_d, directions, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    int rc = vxrt_launch_trace_rays(c, d_o, d_d, n, max_iterations, d_h);
    if (rc) return rc;
    VX_CUDA(cudaMemcpyAsync(hits, d_h, (size_t)n * sizeof(vxrt_ray_hit), cudaMemcpyDeviceToHost, c->stream));
    VX_CUDA(cudaStreamSynchronize(c->stream));  // host buffers are only borrowed for the call
    return VXRT_OK;
}

int vxrt_cuda_raycast_detect(vxrt_ctx* c, const float* positions, const float* directions, int32_t n, int32_t* out) {
    REQUIRE_CTX(c);
    if (n < 0) return vxrt_fail(VXRT_E_INVALID, "raycast_detect: n < 0");
    if (n == 0) return VXRT_OK;
    REQUIRE_PTR(positions); REQUIRE_PTR(directions); REQUIRE_PTR(out);
    if (!c->world_uploaded) return vxrt_fail(VXRT_E_STATE, "raycast_detect before upload_world");
    const size_t vec_bytes = ((size_t)n * 3 * sizeof(float) + 255) / 256 * 256, need = 2 * vec_bytes + (size_t)n * 8 * sizeof(int32_t);
    if (need > c->ray_cap) {
        if (c->d_ray_buf) VX_CUDA(cudaFree(c->d_ray_buf));
        c->d_ray_buf = nullptr; c->ray_cap = 0;
        VX_CUDA(cudaMalloc(&c->d_ray_buf, need));
        c->ray_cap = need;
    }
    char* base = (char*)c->d_ray_buf;
    float* d_o = (float*)base;
    float* d_d = (float*)(base + vec_bytes);
    int32_t* d_r = (int32_t*)(base + 2 * vec_bytes);
    VX_CUDA(cudaMemcpyAsync(d_o, positions, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    VX_CUDA(cudaMemcpyAsync(d_d, directions, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    int rc = vxrt_launch_raycast_detect(c, d_o, d_d, n, d_r);
    if (rc) return rc;
    VX_CUDA(cudaMemcpyAsync(out, d_r, (size_t)n * 8 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    VX_CUDA(cudaStreamSynchronize(c->stream));
    return VXRT_OK;
}

int vxrt_cuda_shadow_trace(vxrt_ctx* c, const vxrt_shadow_params* p) {
    REQUIRE_CTX_LANES(c); REQUIRE_PTR(p);
    if (!c->df_valid) return vxrt_fail(VXRT_E_STATE, "shadow_trace needs a world and a distance field");
    int rc = check_frame(__func__, p->width, p->height, p->tile);
    if (rc) return rc;
    if (p->alpha_test && !c->tex_set[VXRT_TEX_ALBEDO])
        return vxrt_fail(VXRT_E_STATE, "alpha-tested traversal needs the albedo texture array (vxrt_cuda_set_texture_array) and the block table");
    if (p->alpha_test && !(p->fov > 0.0f && p->fov < 180.0f)) return vxrt_fail(VXRT_E_INVALID, "alpha test: fov must be in (0, 180) degrees");
    if (!c->att[VXRT_ATT_INITIAL_T].ptr || !c->att[VXRT_ATT_INITIAL_NORMAL].ptr)
        return vxrt_fail(VXRT_E_STATE, "shadow_trace consumes the primary G-buffer: run initial_trace first");
    if (p->soft_shadows && !c->d_blue_tex) return vxrt_fail(VXRT_E_STATE, "soft shadows need set_blue_noise_texture");
    if (p->max_iterations < 0) return vxrt_fail(VXRT_E_INVALID, "max_iterations < 0");
    Lane1Pass lane(c);
    if ((rc = lane.begin())) return rc;
    return lane.end(vxrt_launch_shadow_trace(c, *p));
}

int vxrt_cuda_set_texture_array(vxrt_ctx* c, int32_t kind, int32_t layers, int32_t w, int32_t h, const uint8_t* rgba8) {
    REQUIRE_CTX(c); REQUIRE_PTR(rgba8);
    if (kind < 0 || kind > 3) return vxrt_fail(VXRT_E_INVALID, "set_texture_array: bad kind %d", kind);
    if (layers < 1 || layers > 255) return vxrt_fail(VXRT_E_INVALID, "set_texture_array: %d layers (1..255, TextureArray.cpp:17-23)", layers);
    if (w < 1 || h < 1 || w > 2048 || h > 2048 || (w & (w - 1)) || (h & (h - 1)) || w != h)
        return vxrt_fail(VXRT_E_INVALID, "set_texture_array: size %dx%d must be a square power of two, at most 2048", w, h);
    return vxrt_set_texture_array(c, kind, layers, w, h, rgba8);
}
int vxrt_cuda_set_skymap(vxrt_ctx* c, int32_t res, const float* rgb_faces) {
    REQUIRE_CTX(c); REQUIRE_PTR(rgb_faces);
    if (res < 1 || res > 2048) return vxrt_fail(VXRT_E_INVALID, "set_skymap: bad resolution %d", res);
    return vxrt_set_skymap(c, res, rgb_faces);
}

static int require_att(vxrt_ctx* c, const char* fn, int id, const char* producer) {
    if (!c->att[id].ptr) return vxrt_fail(VXRT_E_STATE, "%s consumes attachment %d: run %s first", fn, id, producer);
    return VXRT_OK;
}
static int require_textures(vxrt_ctx* c, const char* fn, bool need_normal) {
    for (int k = 0; k < 4; ++k) {
        if (k == VXRT_TEX_NORMAL && !need_normal) continue;
        if (!c->tex_set[k]) return vxrt_fail(VXRT_E_STATE, "%s needs texture array %d (vxrt_cuda_set_texture_array)", fn, k);
    }
    return VXRT_OK;
}

int vxrt_cuda_generate_gbuffer(vxrt_ctx* c, const vxrt_gbuffer_params* p) {
    // With the lanes on, lane 1 (ctx.h lane1_gbuffer): the material G-buffer is read by the reflection pass, the direct term and the filters,
    // never by the GI, so the frame's diffuse_trace - issued next on lane 0 - need not wait for it.  Lane 1 waits for what lane 0 has
    // queued so far (the primary pass); its later passes follow in stream order, lane 2 waits for lane 1's state before the reflection pass.
    REQUIRE_CTX_LANES(c); REQUIRE_PTR(p);
    const bool on_lane1 = c->lane1_gbuffer && lanes_on(c);
    if (!on_lane1 && (c->lane1_pending || c->lane2_pending || c->gi_fork_valid)) { if (int jrc = vxrt_join_lane1(c, false)) return jrc; }
    int rc = check_frame(__func__, p->width, p->height, p->tile);
    if (rc) return rc;
    if ((rc = require_att(c, __func__, VXRT_ATT_INITIAL_INVT, "vxrt_cuda_initial_trace"))) return rc;
    if ((rc = require_textures(c, __func__, true))) return rc;
    if (!on_lane1) return vxrt_launch_generate_gbuffer(c, *p);
    Lane1Pass lane(c);
    if ((rc = lane.begin())) return rc;
    return lane.end(vxrt_launch_generate_gbuffer(c, *p));
}
int vxrt_cuda_shade_direct(vxrt_ctx* c, const vxrt_direct_params* p) {
    REQUIRE_CTX_LANES(c); REQUIRE_PTR(p);
    int rc = check_frame(__func__, p->width, p->height, p->tile);
    if (rc) return rc;
    if ((rc = require_att(c, __func__, VXRT_ATT_INITIAL_INVT, "vxrt_cuda_initial_trace"))) return rc;
    if ((rc = require_att(c, __func__, VXRT_ATT_GBUF_ALBEDO, "vxrt_cuda_generate_gbuffer"))) return rc;
    if ((rc = require_att(c, __func__, c->shadow_source, c->shadow_source == VXRT_ATT_SHADOW ? "vxrt_cuda_shadow_trace" : "the shadow denoiser"))) return rc;
    Lane1Pass lane(c);
    if ((rc = lane.begin(false, false, true))) return rc;
    return lane.end(vxrt_launch_shade_direct(c, *p));
}
int vxrt_cuda_diffuse_trace(vxrt_ctx* c, const vxrt_gi_params* p) {
    REQUIRE_CTX_LANES(c); REQUIRE_PTR(p);
    // lane 0.  The GI reads the primary G-buffer and writes the GI attachments and its own arena: it waits for lane 1 only when a reflection pass
    // queued there (before this call) reads those attachments
    if (c->lane1_pending && (c->lane1_reads_gi || !lanes_on(c))) { if (int jrc = vxrt_join_lane1(c, false)) return jrc; }
    if (!c->df_valid) return vxrt_fail(VXRT_E_STATE, "diffuse_trace needs a world and a distance field");
    int rc = check_frame(__func__, p->width, p->height, p->tile);
    if (rc) return rc;
    if ((rc = require_att(c, __func__, VXRT_ATT_INITIAL_T, "vxrt_cuda_initial_trace"))) return rc;
    if ((rc = require_textures(c, __func__, false))) return rc;
    if (!c->d_blue_noise) return vxrt_fail(VXRT_E_STATE, "diffuse_trace needs vxrt_cuda_set_blue_noise");
    if (!c->sky.data) return vxrt_fail(VXRT_E_STATE, "diffuse_trace needs vxrt_cuda_set_skymap");
    if (!p->use_blue_noise) return vxrt_fail(VXRT_E_UNSUPPORTED, "the fract(sin()) hash RNG (u_UseBlueNoise = false) is not portable and not implemented");
    if (p->spp < 1 || p->trace_length < 0 || p->shadow_trace_length < 0) return vxrt_fail(VXRT_E_INVALID, "diffuse_trace: bad spp / trace length");
    if (lanes_on(c)) {
        if ((rc = ensure_lanes(c))) return rc;
        if (!c->gi_fork_valid) { VX_CUDA(cudaEventRecord(c->gi_fork, c->stream)); c->gi_fork_valid = true; }   // what lane-1 passes of this frame wait for
        rc = vxrt_launch_diffuse_trace(c, *p);
        VX_CUDA(cudaEventRecord(c->gi_done, c->stream));
        return rc;
    }
    return vxrt_launch_diffuse_trace(c, *p);
}
int vxrt_cuda_svgf_temporal(vxrt_ctx* c, const vxrt_svgf_temporal_params* p) {
    REQUIRE_CTX(c); REQUIRE_PTR(p);
    int rc = check_frame(__func__, p->width, p->height, p->tile, true);
    if (rc) return rc;
    return vxrt_launch_svgf_temporal(c, *p);
}
int vxrt_cuda_svgf_prespatial(vxrt_ctx* c, const vxrt_svgf_prespatial_params* p) {
    REQUIRE_CTX(c); REQUIRE_PTR(p);
    int rc = check_frame(__func__, p->width, p->height, p->tile, true);
    if (rc) return rc;
    return vxrt_launch_svgf_prespatial(c, *p);
}
int vxrt_cuda_svgf_variance(vxrt_ctx* c, const vxrt_svgf_variance_params* p) {
    REQUIRE_CTX(c); REQUIRE_PTR(p);
    int rc = check_frame(__func__, p->width, p->height, p->tile, true);
    if (rc) return rc;
    return vxrt_launch_svgf_variance(c, *p);
}
int vxrt_cuda_svgf_spatial(vxrt_ctx* c, const vxrt_svgf_spatial_params* p) {
    REQUIRE_CTX(c); REQUIRE_PTR(p);
    int rc = check_frame(__func__, p->width, p->height, p->tile, true);
    if (rc) return rc;
    return vxrt_launch_svgf_spatial(c, *p);
}
int vxrt_cuda_svgf_end_frame(vxrt_ctx* c) {
    REQUIRE_CTX(c);
    return vxrt_launch_svgf_end_frame(c);
}
int vxrt_cuda_end_frame(vxrt_ctx* c) {
    REQUIRE_CTX(c);
    return vxrt_launch_svgf_end_frame(c);
}
int vxrt_cuda_select_shadow(vxrt_ctx* c, int32_t id) {
    REQUIRE_CTX(c);
    if (id != VXRT_ATT_SHADOW && id != VXRT_ATT_SHADOW_TEMPORAL_A && id != VXRT_ATT_SHADOW_TEMPORAL_B && id != VXRT_ATT_SHADOW_FILTERED)
        return vxrt_fail(VXRT_E_INVALID, "select_shadow: attachment %d is not a shadow image", id);
    c->shadow_source = id;
    return VXRT_OK;
}
int vxrt_cuda_specular_temporal(vxrt_ctx* c, const vxrt_specular_temporal_params* p) {
    REQUIRE_CTX(c); REQUIRE_PTR(p);
    int rc = check_frame(__func__, p->width, p->height, p->tile, true);
    if (rc) return rc;
    return vxrt_launch_specular_temporal(c, *p);
}
int vxrt_cuda_reflection_denoise(vxrt_ctx* c, const vxrt_reflection_denoise_params* p) {
    REQUIRE_CTX(c); REQUIRE_PTR(p);
    int rc = check_frame(__func__, p->width, p->height, p->tile, true);
    if (rc) return rc;
    return vxrt_launch_reflection_denoise(c, *p);
}
int vxrt_cuda_shadow_temporal(vxrt_ctx* c, const vxrt_shadow_temporal_params* p) {
    REQUIRE_CTX(c); REQUIRE_PTR(p);
    int rc = check_frame(__func__, p->width, p->height, p->tile, true);
    if (rc) return rc;
    return vxrt_launch_shadow_temporal(c, *p);
}
int vxrt_cuda_shadow_filter(vxrt_ctx* c, const vxrt_shadow_filter_params* p) {
    REQUIRE_CTX(c); REQUIRE_PTR(p);
    int rc = check_frame(__func__, p->width, p->height, p->tile, true);
    if (rc) return rc;
    return vxrt_launch_shadow_filter(c, *p);
}
int vxrt_cuda_reflection_trace(vxrt_ctx* c, const vxrt_reflection_params* p) {
    // lane 1 (behind the sun shadow it reads).  The GI's SH attachments are its ambient base (ReflectionTraceFrag.glsl main(): SHToIrridiance of
    // u_DiffuseSHy), read by the shading kernels; ray generation and the first trace run beside the GI unless DeriveFromDiffuseSH reads them at once
    REQUIRE_CTX_LANES(c); REQUIRE_PTR(p);
    if (!c->df_valid) return vxrt_fail(VXRT_E_STATE, "reflection_trace needs a world and a distance field");
    int rc = check_frame(__func__, p->width, p->height, p->tile);
    if (rc) return rc;
    if ((rc = require_att(c, __func__, VXRT_ATT_INITIAL_T, "vxrt_cuda_initial_trace"))) return rc;
    if ((rc = require_att(c, __func__, VXRT_ATT_GBUF_NORMAL, "vxrt_cuda_generate_gbuffer"))) return rc;
    if ((rc = require_att(c, __func__, VXRT_ATT_GI_SH, "vxrt_cuda_diffuse_trace"))) return rc;
    if ((rc = require_att(c, __func__, c->shadow_source, c->shadow_source == VXRT_ATT_SHADOW ? "vxrt_cuda_shadow_trace" : "the shadow denoiser"))) return rc;
    if ((rc = require_textures(c, __func__, true))) return rc;
    if (!c->d_blue_noise) return vxrt_fail(VXRT_E_STATE, "reflection_trace needs vxrt_cuda_set_blue_noise");
    if (!c->sky.data) return vxrt_fail(VXRT_E_STATE, "reflection_trace needs vxrt_cuda_set_skymap");
    if (!p->use_blue_noise) return vxrt_fail(VXRT_E_UNSUPPORTED, "the fract(sin()) hash RNG is not implemented");
    if (p->spp < 1 || p->trace_length < 0 || p->shadow_trace_length < 0) return vxrt_fail(VXRT_E_INVALID, "reflection_trace: bad spp / trace length");
    Lane1Pass lane(c, true);
    if ((rc = lane.begin(true, p->derive_from_diffuse_sh != 0))) return rc;
    return lane.end(vxrt_launch_reflection_trace(c, *p));
}

int vxrt_cuda_join_passes(vxrt_ctx* c) {
    REQUIRE_CTX(c);   // the join itself
    return VXRT_OK;
}
int vxrt_cuda_stats_enable(vxrt_ctx* c, int32_t on) {
    REQUIRE_CTX(c);
    c->stats_on = on != 0;
    return VXRT_OK;
}
int vxrt_cuda_stats_read(vxrt_ctx* c, vxrt_trace_stats* out, int32_t reset) {
    REQUIRE_CTX_READER(c); REQUIRE_PTR(out);
    TraceStatsDev h[2];
    VX_CUDA(cudaMemcpyAsync(h, c->d_stats, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    VX_CUDA(cudaStreamSynchronize(c->stream));
    out->rays = h[0].rays + h[1].rays; out->iterations = h[0].iterations + h[1].iterations;
    out->dda_steps = h[0].dda_steps + h[1].dda_steps; out->hits = h[0].hits + h[1].hits;
    if (reset) {
        c->probe_acc.rays += h[1].rays; c->probe_acc.iterations += h[1].iterations;
        c->probe_acc.dda_steps += h[1].dda_steps; c->probe_acc.hits += h[1].hits;
        VX_CUDA(cudaMemsetAsync(c->d_stats, 0, 2 * sizeof(TraceStatsDev), c->stream));
    }
    return VXRT_OK;
}
// ---- world producers (world.cu) ----
static int ensure_staging(vxrt_ctx* c, size_t need) {
    if (need > c->ray_cap) {
        if (c->d_ray_buf) VX_CUDA(cudaFree(c->d_ray_buf));
        c->d_ray_buf = nullptr; c->ray_cap = 0;
        VX_CUDA(cudaMalloc(&c->d_ray_buf, need));
        c->ray_cap = need;
    }
    return VXRT_OK;
}

int vxrt_cuda_generate_world(vxrt_ctx* c, const vxrt_worldgen_params* p) {
    REQUIRE_CTX(c); REQUIRE_PTR(p);
    const int ids[4] = {p->grass_id, p->dirt_id, p->stone_id, p->sand_id};
    for (int k = 0; k < 4; ++k)
        if (ids[k] < 0 || ids[k] > 255) return vxrt_fail(VXRT_E_INVALID, "generate_world: block id %d out of range", ids[k]);
    int rc = vxrt_launch_generate_world(c, *p);
    if (rc) return rc;
    c->world_uploaded = true;
    c->df_valid = false;
    return VXRT_OK;
}

int vxrt_cuda_import_sections(vxrt_ctx* c, const uint8_t* block_ids, const uint8_t* data_nibbles, const uint8_t* has_data,
                              const int32_t* section_origins, int32_t n, const int32_t import_origin[3], const uint8_t lut[256],
                              int32_t clear_first) {
    REQUIRE_CTX(c);
    if (n < 0) return vxrt_fail(VXRT_E_INVALID, "import_sections: n < 0");
    REQUIRE_PTR(import_origin); REQUIRE_PTR(lut);
    if (!clear_first && !c->world_uploaded) return vxrt_fail(VXRT_E_STATE, "import_sections without clear_first needs a world");
    if (clear_first) VX_CUDA(cudaMemsetAsync(c->d_blocks, 0, c->nvox, c->stream));
    if (n > 0) {
        REQUIRE_PTR(block_ids); REQUIRE_PTR(data_nibbles); REQUIRE_PTR(has_data); REQUIRE_PTR(section_origins);
        const size_t b_ids = (size_t)n * 4096, b_nib = (size_t)n * 2048, b_org = ((size_t)n * 3 * sizeof(int32_t) + 255) / 256 * 256;
        int rc = ensure_staging(c, b_ids + b_nib + b_org + (size_t)n);
        if (rc) return rc;
        uint8_t* base = (uint8_t*)c->d_ray_buf;
        uint8_t* d_ids = base; uint8_t* d_nib = base + b_ids; int32_t* d_org = (int32_t*)(base + b_ids + b_nib); uint8_t* d_has = base + b_ids + b_nib + b_org;
        VX_CUDA(cudaMemcpyAsync(d_ids, block_ids, b_ids, cudaMemcpyHostToDevice, c->stream));
        VX_CUDA(cudaMemcpyAsync(d_nib, data_nibbles, b_nib, cudaMemcpyHostToDevice, c->stream));
        VX_CUDA(cudaMemcpyAsync(d_org, section_origins, (size_t)n * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        VX_CUDA(cudaMemcpyAsync(d_has, has_data, (size_t)n, cudaMemcpyHostToDevice, c->stream));
        rc = vxrt_launch_import_sections(c, d_ids, d_nib, d_has, d_org, n, import_origin, lut);
        if (rc) return rc;
        VX_CUDA(cudaStreamSynchronize(c->stream));  // host buffers are only borrowed for the call
    }
    c->world_uploaded = true;
    c->df_valid = false;
    return VXRT_OK;
}

int vxrt_cuda_collect_lights(vxrt_ctx* c, int32_t* xyz_out, int32_t capacity, int32_t* count) {
    REQUIRE_CTX(c); REQUIRE_PTR(count);
    if (capacity < 0) return vxrt_fail(VXRT_E_INVALID, "collect_lights: capacity < 0");
    if (capacity > 0) REQUIRE_PTR(xyz_out);
    if (!c->world_uploaded) return vxrt_fail(VXRT_E_STATE, "collect_lights before a world exists");
    const int chunks = vxrt_lights_chunks(c);
    const size_t b_counts = ((size_t)(chunks + 1) * sizeof(unsigned) + 255) / 256 * 256;
    int rc = ensure_staging(c, b_counts + (size_t)capacity * 3 * sizeof(int32_t));
    if (rc) return rc;
    unsigned* d_counts = (unsigned*)c->d_ray_buf;
    int32_t* d_out = (int32_t*)((uint8_t*)c->d_ray_buf + b_counts);
    rc = vxrt_launch_collect_lights(c, d_counts, d_out, capacity);
    if (rc) return rc;
    unsigned total = 0;
    VX_CUDA(cudaMemcpyAsync(&total, d_counts + chunks, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    VX_CUDA(cudaStreamSynchronize(c->stream));
    *count = (int32_t)total;
    const size_t wr = (size_t)(total < (unsigned)capacity ? total : (unsigned)capacity);
    if (wr) {
        VX_CUDA(cudaMemcpyAsync(xyz_out, d_out, wr * 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        VX_CUDA(cudaStreamSynchronize(c->stream));
    }
    return VXRT_OK;
}

int vxrt_cuda_lpv_repropagate(vxrt_ctx* c, const int32_t* lights_xyz, int32_t n_lights, int32_t distance_limit) {
    REQUIRE_CTX(c);
    if (distance_limit < 0 || distance_limit > 255) return vxrt_fail(VXRT_E_INVALID, "lpv_repropagate: distance_limit out of range");
    if (n_lights < 0 || (size_t)n_lights > c->nvox) return vxrt_fail(VXRT_E_INVALID, "lpv_repropagate: n_lights out of range");
    if (!c->world_uploaded) return vxrt_fail(VXRT_E_STATE, "lpv_repropagate before a world exists");
    const int chunks = vxrt_lights_chunks(c);
    const size_t b_counts = ((size_t)(chunks + 2) * sizeof(unsigned) + 255) / 256 * 256;
    int rc;
    if (lights_xyz) {
        rc = ensure_staging(c, b_counts + (size_t)(n_lights > 0 ? n_lights : 1) * 3 * sizeof(int32_t));
        if (rc) return rc;
        unsigned* d_count = (unsigned*)c->d_ray_buf;
        int32_t* d_lights = (int32_t*)((uint8_t*)c->d_ray_buf + b_counts);
        const unsigned n = (unsigned)n_lights;
        VX_CUDA(cudaMemcpyAsync(d_count, &n, sizeof(unsigned), cudaMemcpyHostToDevice, c->stream));
        if (n_lights) VX_CUDA(cudaMemcpyAsync(d_lights, lights_xyz, (size_t)n_lights * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        rc = c->lpv_coop ? vxrt_launch_lpv_repropagate_coop(c, d_lights, n_lights, distance_limit) : VXRT_E_UNSUPPORTED;
        if (rc == VXRT_E_UNSUPPORTED) rc = vxrt_launch_lpv_repropagate(c, d_lights, d_count, n_lights, distance_limit);
        if (rc) return rc;
        VX_CUDA(cudaStreamSynchronize(c->stream));   // host buffers are only borrowed for the call
    } else {
        // the light list of LoadWorld, scanned on the device and consumed there: nothing comes back to the host
        if (c->lpv_coop) {
            rc = vxrt_launch_lpv_repropagate_coop(c, nullptr, 0, distance_limit);
            if (rc != VXRT_E_UNSUPPORTED) {
                if (rc == VXRT_OK) c->lpv_valid = true;
                return rc;
            }
        }
        const int capacity = 1 << 20;
        rc = ensure_staging(c, b_counts + (size_t)capacity * 3 * sizeof(int32_t));
        if (rc) return rc;
        unsigned* d_counts = (unsigned*)c->d_ray_buf;
        int32_t* d_lights = (int32_t*)((uint8_t*)c->d_ray_buf + b_counts);
        rc = vxrt_launch_collect_lights(c, d_counts, d_lights, capacity);
        if (rc) return rc;
        rc = vxrt_launch_lpv_repropagate(c, d_lights, d_counts + chunks, capacity, distance_limit);
        if (rc) return rc;
    }
    c->lpv_valid = true;
    return VXRT_OK;
}

int vxrt_cuda_lpv_edit(vxrt_ctx* c, int32_t op, int32_t x, int32_t y, int32_t z, int32_t block, int32_t distance_limit) {
    REQUIRE_CTX(c);
    if (op != 0 && op != 1) return vxrt_fail(VXRT_E_INVALID, "lpv_edit: op must be 0 (break) or 1 (place)");
    if (distance_limit < 0 || distance_limit > 255) return vxrt_fail(VXRT_E_INVALID, "lpv_edit: distance_limit out of range");
    if (block < 0 || block > 255) return vxrt_fail(VXRT_E_INVALID, "lpv_edit: block id out of range");
    // World::Raycast returns before touching anything when the voxel is on or outside the faces of the grid (World.cpp:267-271)
    if (x <= 0 || y <= 0 || z <= 0 || x >= c->nx || y >= c->ny || z >= c->nz) return vxrt_fail(VXRT_E_INVALID, "lpv_edit: position not strictly inside the grid");
    if (!c->world_uploaded) return vxrt_fail(VXRT_E_STATE, "lpv_edit before a world exists");
    int overflowed = 0;
    int rc = vxrt_launch_lpv_edit(c, op, x, y, z, block, distance_limit, &overflowed);
    if (rc) return rc;
    if (overflowed) return vxrt_fail(VXRT_E_NOMEM, "lpv_edit: queue capacity exceeded");
    c->lpv_valid = true;
    return VXRT_OK;
}

int vxrt_cuda_lpv_average_colors(vxrt_ctx* c, float* rgba_out) {
    REQUIRE_CTX(c);
    if (!c->tex_set[VXRT_TEX_ALBEDO]) return vxrt_fail(VXRT_E_STATE, "lpv_average_colors: the albedo texture array has not been set");
    int rc = vxrt_launch_lpv_average_colors(c);
    if (rc) return rc;
    if (rgba_out) {
        VX_CUDA(cudaMemcpyAsync(rgba_out, c->d_lpv_avg, 128 * 4 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        VX_CUDA(cudaStreamSynchronize(c->stream));
    }
    return VXRT_OK;
}

int vxrt_cuda_lpv_set_average_colors(vxrt_ctx* c, const float* rgba) {
    REQUIRE_CTX(c); REQUIRE_PTR(rgba);
    if (!c->d_lpv_avg) VX_CUDA(cudaMalloc(&c->d_lpv_avg, 128 * 4 * sizeof(float)));
    VX_CUDA(cudaMemcpyAsync(c->d_lpv_avg, rgba, 128 * 4 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    VX_CUDA(cudaStreamSynchronize(c->stream));
    return VXRT_OK;
}

int vxrt_cuda_lpv_sample(vxrt_ctx* c, const float* points, int32_t n, const float dither[3], float* rgb_out) {
    REQUIRE_CTX(c); REQUIRE_PTR(dither);
    if (n < 0) return vxrt_fail(VXRT_E_INVALID, "lpv_sample: n < 0");
    if (n == 0) return VXRT_OK;
    REQUIRE_PTR(points); REQUIRE_PTR(rgb_out);
    if (!c->lpv_valid) return vxrt_fail(VXRT_E_STATE, "lpv_sample: no light propagation volume yet (lpv_repropagate / lpv_upload)");
    if (!c->d_lpv_avg) return vxrt_fail(VXRT_E_STATE, "lpv_sample: no average block colours yet (lpv_average_colors)");
    const size_t bytes = ((size_t)n * 3 * sizeof(float) + 255) / 256 * 256;
    int rc = ensure_staging(c, 2 * bytes);
    if (rc) return rc;
    float* d_points = (float*)c->d_ray_buf;
    float* d_out = (float*)((uint8_t*)c->d_ray_buf + bytes);
    VX_CUDA(cudaMemcpyAsync(d_points, points, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    rc = vxrt_launch_lpv_sample(c, d_points, n, dither, d_out);
    if (rc) return rc;
    VX_CUDA(cudaMemcpyAsync(rgb_out, d_out, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    VX_CUDA(cudaStreamSynchronize(c->stream));
    return VXRT_OK;
}

int vxrt_cuda_lpv_download(vxrt_ctx* c, uint8_t* level, uint8_t* block_type) {
    REQUIRE_CTX(c);
    int rc = vxrt_lpv_ensure(c);
    if (rc) return rc;
    if (level) VX_CUDA(cudaMemcpyAsync(level, c->d_lpv, c->nvox, cudaMemcpyDeviceToHost, c->stream));
    if (block_type) VX_CUDA(cudaMemcpyAsync(block_type, c->d_lpv + c->nvox, c->nvox, cudaMemcpyDeviceToHost, c->stream));
    VX_CUDA(cudaStreamSynchronize(c->stream));
    return VXRT_OK;
}

int vxrt_cuda_lpv_upload(vxrt_ctx* c, const uint8_t* level, const uint8_t* block_type) {
    REQUIRE_CTX(c); REQUIRE_PTR(level); REQUIRE_PTR(block_type);
    int rc = vxrt_lpv_ensure(c);
    if (rc) return rc;
    VX_CUDA(cudaMemcpyAsync(c->d_lpv, level, c->nvox, cudaMemcpyHostToDevice, c->stream));
    VX_CUDA(cudaMemcpyAsync(c->d_lpv + c->nvox, block_type, c->nvox, cudaMemcpyHostToDevice, c->stream));
    VX_CUDA(cudaStreamSynchronize(c->stream));
    c->lpv_valid = true;
    return VXRT_OK;
}

int vxrt_cuda_gather_peak(vxrt_ctx* c, int32_t rounds, double* sectors_per_second) {
    REQUIRE_CTX(c); REQUIRE_PTR(sectors_per_second);
    if (rounds < 1 || rounds > (1 << 16)) return vxrt_fail(VXRT_E_INVALID, "gather_peak: rounds out of range");
    return vxrt_launch_gather_peak(c, rounds, sectors_per_second);
}
int vxrt_cuda_probe_read(vxrt_ctx* c, double* total_ms, int64_t* launches, vxrt_trace_stats* stats, int32_t reset) {
    REQUIRE_CTX_READER(c);
    VX_CUDA(cudaStreamSynchronize(c->stream));
    double ms = 0.0;
    for (size_t i = 0; i + 1 < c->probe_used; i += 2) {
        float t = 0.0f;
        VX_CUDA(cudaEventElapsedTime(&t, c->probe_ev[i], c->probe_ev[i + 1]));
        ms += t;
    }
    if (total_ms) *total_ms = ms;
    if (launches) *launches = (int64_t)(c->probe_used / 2);
    if (stats) {
        TraceStatsDev h;
        VX_CUDA(cudaMemcpy(&h, c->d_stats + 1, sizeof(h), cudaMemcpyDeviceToHost));
        stats->rays = h.rays + c->probe_acc.rays; stats->iterations = h.iterations + c->probe_acc.iterations;
        stats->dda_steps = h.dda_steps + c->probe_acc.dda_steps; stats->hits = h.hits + c->probe_acc.hits;
    }
    if (reset) {
        c->probe_used = 0;
        c->probe_acc = TraceStatsDev{0, 0, 0, 0};
        VX_CUDA(cudaMemsetAsync(c->d_stats + 1, 0, sizeof(TraceStatsDev), c->stream));
    }
    return VXRT_OK;
}

}  // extern "C"
