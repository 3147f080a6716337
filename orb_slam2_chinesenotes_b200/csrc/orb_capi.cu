// orb_capi.cu -- context management and the extern "C" boundary of the extractor
// (include/orb_b200.h).  Host code only: owns device memory, streams, the per-shape plan and
// the chunked, pipelined batch loop; all arithmetic lives in the kernels.
//
// Batch loop: a batch is cut into chunks of ctx->chunk frames.  With host buffers the chunks
// rotate through NSLOT buffer sets so that chunk i+1's host->device copy, chunk i's kernels and
// chunk i-1's device->host copy run concurrently on three streams (copy engines are full
// duplex).  Chunks also alternate between NSLOT compute streams (forked from / joined to the
// context's stream by events), so that the latency-bound sparse kernels of one chunk (quadtree,
// descriptors) share the SMs with the issue-bound dense kernels of the next; while stage timing
// is enabled (orbx_profile) everything runs on the one stream so the CUDA events bracket kernels.
#include <cstdarg>
#include <cstdio>
#include <chrono>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only; a no-op unless a profiler is attached

#include "orb_device.cuh"
#include "orb_launch.h"
#include "orb_stereo.h"

namespace {

#ifndef ORB_NCOMP
#define ORB_NCOMP 2      // compute streams the chunks of the piped path alternate between (<= ORB_NSLOT).  Re-measured with the final
                         // kernels (1024 KITTI frames from pinned host memory, alternating runs, tools/ncomp_exp.sh): 2 streams 88.4-88.5 k
                         // frames/s per synchronous call and 99.1-99.2 k submitted; 3 streams 84.6-87.1 k / 91.6-98.3 k; 1 stream 82.0 k /
                         // 91.4 k -- two chunks side by side hide each other's launch gaps, a third only takes cache and SM space
#endif
#ifndef ORB_NSLOT
#define ORB_NSLOT 3
#endif
const int NSLOT = ORB_NSLOT;

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= bytes) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        cudaError_t e = cudaMalloc(&p, n);
        if (e != cudaSuccess) return e;
        bytes = n;
        // fresh device memory holds whatever its last owner left: entries of an output behind n_out are unspecified by
        // contract, but they should not differ from run to run (a staged output is copied out whole)
        // (the context's streams are non-blocking, i.e. NOT ordered against the default stream the memset runs on: wait for it
        // here, once per allocation, or a kernel queued next could be overwritten by it)
        e = cudaMemset(p, 0, n);
        return e == cudaSuccess ? cudaStreamSynchronize(0) : e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};

struct PinBuf {     // pinned host memory
    void* p = nullptr;
    size_t bytes = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= bytes) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; bytes = 0;
        cudaError_t e = cudaHostAlloc(&p, n, cudaHostAllocDefault);
        if (e == cudaSuccess) bytes = n;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; bytes = 0; }
};

// One set of per-chunk device buffers.
struct Slot {
    DevBuf img_stage, pyr, blur, cand, node_of, counts, lkp, out_kps, out_desc, out_n;
    DevBuf st_sad, st_sorted, st_rows, out_ur, out_depth, out_ns;   // stereo: scratch and staged mvuRight / mvDepth / counts
    int frames = 0;                       // frames the work buffers are sized for
    OrbFastMaps fmaps;                    // FAST's tensor maps for the batch layout last seen in this slot
    cudaEvent_t h2d_done = nullptr, compute_done = nullptr, d2h_done = nullptr;
    cudaStream_t aux = nullptr;            // the blur runs here, beside FAST + quadtree (both only need the pyramid)
    cudaEvent_t pyr_done = nullptr, blur_done = nullptr;
    // a handful of frames (latency): FAST + quadtree of level l run on their own stream as soon as level l exists
    cudaStream_t lvl[ORB_MAX_LEVELS] = { nullptr };
    cudaEvent_t lvl_ready[ORB_MAX_LEVELS] = { nullptr }, lvl_done[ORB_MAX_LEVELS] = { nullptr };
    bool used = false;
    void release()
    {
        DevBuf* b[] = { &img_stage, &pyr, &blur, &cand, &node_of, &counts, &lkp, &out_kps, &out_desc, &out_n,
                        &st_sad, &st_sorted, &st_rows, &out_ur, &out_depth, &out_ns };
        for (DevBuf* x : b) x->release();
        frames = 0;
    }
};

// What orbx_extract_stereo_batch adds to a batch call: frames are L0,R0,L1,R1,...
struct StereoReq {
    float bf, fx;
    float* u_right; float* depth; int* n_stereo;
};

// The single-call path (one frame or one stereo pair from host memory): everything one call does -- the upload, every
// kernel, the download of the results and of the padded pyramid -- is ONE instantiated CUDA graph over fixed buffers.
struct FramePath {
    cudaGraphExec_t exec = nullptr;
    // what the graph was captured for
    int w = 0, h = 0, pitch = 0, frames = 0, cap = 0, nlevels = 0;
    bool stereo = false, pyr = false;
    float bf = 0.f, fx = 0.f;
    cudaStream_t stream = nullptr;
    PinBuf h_in, h_kps, h_desc, h_n, h_ur, h_dep, h_ns, h_pad;
    DevBuf d_in, d_kps, d_desc, d_n, d_ur, d_dep, d_ns, d_pad;
    OrbBorderJob job;
    double last_us[3] = { 0, 0, 0 };       // host time of the last call: staging copy, graph launch, wait for the results
    void release()
    {
        if (exec) cudaGraphExecDestroy(exec);
        exec = nullptr;
        PinBuf* hb[] = { &h_in, &h_kps, &h_desc, &h_n, &h_ur, &h_dep, &h_ns, &h_pad };
        for (PinBuf* x : hb) x->release();
        DevBuf* db[] = { &d_in, &d_kps, &d_desc, &d_n, &d_ur, &d_dep, &d_ns, &d_pad };
        for (DevBuf* x : db) x->release();
    }
};

struct StageTimer {
    cudaEvent_t ev[2];
    int stage;
};

} // namespace

struct orbx_ctx {
    OrbParams params;
    int device = 0;
    cudaStream_t own_stream = nullptr;     // default compute stream
    cudaStream_t stream = nullptr;         // compute stream in use (own or caller's)
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    cudaStream_t xstream[NSLOT - 1] = { nullptr };   // extra compute streams (slots 1..)
    cudaEvent_t fork_ev = nullptr, join_ev[NSLOT - 1] = { nullptr };
    // Frames per internal pass.  With host buffers the chunks are what the copy / compute pipeline overlaps: smaller
    // chunks cost launch tails (a 64-frame launch list takes 0.71 ms, an eighth of a 512-frame one 0.57 ms), larger ones a
    // longer pipeline fill and drain, so the size is chosen per call from the frame size and the batch (run_batch has the
    // rule and the measurements) unless orbx_set_chunk fixed it.  Slots and compute streams (ORB_NSLOT / ORB_NCOMP = 3/2,
    // 3/1, 4/2, 6/3, 6/2, 8/4) all measured 80-86 k frames/s on KITTI before the last kernel changes: the pipeline is not short of
    // buffers; with the final kernels two compute streams beat three and one (see ORB_NCOMP above).  With everything
    // device resident there is nothing to overlap and bigger launches are simply more efficient (1241x376: 6.66 ms per 512
    // frames in chunks of 64, 6.03 ms as one chunk), so the chunk only bounds the work buffers (about 2.6 MB per frame).
    int chunk = 0, chunk_resident = 512;   // chunk 0: chosen per call (run_batch)
    std::string err;

    bool have_plan = false;
    OrbPlan plan;
    DevBuf taps, border_tmp;
    Slot slot[NSLOT + 1];                  // the last one belongs to the single-call path
    FramePath fp;

    // what is resident from the last call (for orbx_pyramid_level / stage taps)
    int last_slot = -1, last_first = 0, last_count = 0;
    const uint8_t* last_img0 = nullptr; size_t last_img0_stride = 0; int last_img0_pitch = 0;

    bool profile = false;
    std::vector<StageTimer> pending;
    std::vector<cudaEvent_t> free_events;
    float stage_ms[ORBX_STAGE_COUNT] = { 0 };
    int stage_launches[ORBX_STAGE_COUNT] = { 0 };
};

namespace {

int fail(orbx_ctx* c, int code, const char* fmt, ...)
{
    if (c) {
        char buf[512];
        va_list ap; va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        c->err = buf;
    }
    return code;
}

#define CU(c, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return fail((c), ORBX_E_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); } while (0)

bool is_device_ptr(const void* p)
{
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int sync_all(orbx_ctx* c)
{
    CU(c, cudaStreamSynchronize(c->h2d_stream));
    CU(c, cudaStreamSynchronize(c->stream));
    for (cudaStream_t x : c->xstream) CU(c, cudaStreamSynchronize(x));
    for (Slot& s : c->slot) {
        CU(c, cudaStreamSynchronize(s.aux));
        for (cudaStream_t x : s.lvl) if (x) CU(c, cudaStreamSynchronize(x));
    }
    CU(c, cudaStreamSynchronize(c->d2h_stream));
    return ORBX_OK;
}

int ensure_plan(orbx_ctx* c, int w, int h)
{
    if (c->have_plan && c->plan.w == w && c->plan.h == h) return ORBX_OK;
    std::vector<OrbTap> taps;
    OrbPlan plan;
    if (orb_plan_build(&c->params, w, h, &plan, &taps)) return fail(c, ORBX_E_SHAPE, "unsupported image shape %dx%d", w, h);
    int rc = sync_all(c);
    if (rc) return rc;
    CU(c, c->taps.reserve(sizeof(OrbTap) * (taps.size() + 1)));
    if (!taps.empty()) CU(c, cudaMemcpy(c->taps.p, taps.data(), sizeof(OrbTap) * taps.size(), cudaMemcpyHostToDevice));
    c->plan = plan;
    c->have_plan = true;
    for (Slot& s : c->slot) s.frames = 0;   // work buffers are re-sized lazily (reserve only grows)
    c->last_slot = -1; c->last_count = 0;
    if (c->fp.exec) { cudaGraphExecDestroy(c->fp.exec); c->fp.exec = nullptr; }   // captured for the old shape
    return ORBX_OK;
}

int ensure_slot(orbx_ctx* c, Slot& s, int frames, int cap, bool stage_in, size_t in_frame_bytes, bool stage_out, bool stereo)
{
    const OrbPlan& P = c->plan;
    if (stereo) {
        const size_t pairs = (size_t)(frames + 1) / 2;
        if (s.st_sad.bytes < pairs * cap * 4 || s.st_rows.bytes < pairs * (P.h + 2) * 4 || (stage_out && s.out_ur.bytes < pairs * cap * 4)) {
            int rc = sync_all(c);
            if (rc) return rc;
            CU(c, s.st_sad.reserve(pairs * cap * 4));
            CU(c, s.st_sorted.reserve(pairs * cap * 16));
            CU(c, s.st_rows.reserve(pairs * (P.h + 2) * 4));
            if (stage_out) {
                CU(c, s.out_ur.reserve(pairs * cap * 4));
                CU(c, s.out_depth.reserve(pairs * cap * 4));
                CU(c, s.out_ns.reserve(pairs * 4));
            }
        }
    }
    if (frames > s.frames) {
        int rc = sync_all(c);
        if (rc) return rc;
        CU(c, s.pyr.reserve((size_t)frames * P.pyr_bytes + 256));
        CU(c, s.blur.reserve((size_t)frames * P.blur_bytes + 256));
        CU(c, s.cand.reserve((size_t)frames * P.cand_per_frame * 4 + 256));
        CU(c, s.node_of.reserve((size_t)frames * P.cand_per_frame * 2 + 256));
        CU(c, s.counts.reserve((size_t)frames * ORB_MAX_LEVELS * 4 * 2));
        CU(c, s.lkp.reserve((size_t)frames * P.kp_per_frame * 4 + 256));
        s.frames = frames;
    }
    if (stage_in && s.img_stage.bytes < (size_t)frames * in_frame_bytes + 256) {
        int rc = sync_all(c);
        if (rc) return rc;
        CU(c, s.img_stage.reserve((size_t)frames * in_frame_bytes + 256));
    }
    if (stage_out && (s.out_kps.bytes < (size_t)frames * cap * sizeof(orbx_kp) || s.out_n.bytes < (size_t)frames * 4)) {
        int rc = sync_all(c);
        if (rc) return rc;
        CU(c, s.out_kps.reserve((size_t)frames * cap * sizeof(orbx_kp)));
        CU(c, s.out_desc.reserve((size_t)frames * cap * 32));
        CU(c, s.out_n.reserve((size_t)frames * 4));
    }
    return ORBX_OK;
}

cudaEvent_t get_event(orbx_ctx* c)
{
    if (!c->free_events.empty()) { cudaEvent_t e = c->free_events.back(); c->free_events.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

// One stage of a chunk: an NVTX range around its launches (shows up in Nsight Systems / ncu --nvtx; free when no tool
// is attached) and, while stage timing is on, a CUDA event pair around them.
const char* const kStageName[ORBX_STAGE_COUNT] = { "orb:pyramid", "orb:fast", "orb:blur", "orb:quadtree", "orb:describe", "orb:stereo" };
struct StageScope {
    orbx_ctx* c; StageTimer t; bool on; cudaStream_t st;
    StageScope(orbx_ctx* ctx, int stage, cudaStream_t stream) : c(ctx), on(ctx->profile), st(stream)
    {
        nvtxRangePushA(stage >= 0 && stage < ORBX_STAGE_COUNT ? kStageName[stage] : "orb");
        if (on) { t.stage = stage; t.ev[0] = get_event(c); t.ev[1] = get_event(c); cudaEventRecord(t.ev[0], st); }
    }
    ~StageScope() { if (on) { cudaEventRecord(t.ev[1], st); c->pending.push_back(t); } nvtxRangePop(); }
};

void collect_timers(orbx_ctx* c)
{
    for (StageTimer& t : c->pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, t.ev[0], t.ev[1]) == cudaSuccess) { c->stage_ms[t.stage] += ms; c->stage_launches[t.stage] += 1; }
        else cudaGetLastError();
        c->free_events.push_back(t.ev[0]); c->free_events.push_back(t.ev[1]);
    }
    c->pending.clear();
}

int ensure_level_streams(orbx_ctx* c, Slot& s, int nlevels)
{
    for (int l = 0; l < nlevels; ++l)
        if (!s.lvl[l]) {
            CU(c, cudaStreamCreateWithFlags(&s.lvl[l], cudaStreamNonBlocking));
            CU(c, cudaEventCreateWithFlags(&s.lvl_ready[l], cudaEventDisableTiming));
            CU(c, cudaEventCreateWithFlags(&s.lvl_done[l], cudaEventDisableTiming));
        }
    return ORBX_OK;
}

// Enqueue the whole extractor for `frames` frames whose level-0 images are device resident.
int enqueue_chunk(orbx_ctx* c, Slot& s, cudaStream_t st, const uint8_t* d_img, size_t frame_stride, int pitch, int frames,
                  orbx_kp* d_kps, uint8_t* d_desc, int* d_n, int cap,
                  const StereoReq* sr = nullptr, float* d_ur = nullptr, float* d_depth = nullptr, int* d_ns = nullptr)
{
    const OrbPlan& P = c->plan;
    OrbBatch io;
    io.img0 = d_img; io.img0_stride = frame_stride; io.img0_pitch = pitch;
    io.pyr = (uint8_t*)s.pyr.p; io.blur = (uint8_t*)s.blur.p;
    io.cand = (uint32_t*)s.cand.p; io.node_of = (uint16_t*)s.node_of.p;
    io.cand_count = (int*)s.counts.p;
    io.lkp_count = io.cand_count + (size_t)s.frames * ORB_MAX_LEVELS;
    io.lkp = (uint32_t*)s.lkp.p;
    io.kps = d_kps; io.desc = d_desc; io.n_out = d_n; io.cap = cap;
    io.taps = (const OrbTap*)c->taps.p;
    CU(c, cudaMemsetAsync(io.cand_count, 0, (size_t)frames * ORB_MAX_LEVELS * 4, st));
    // A handful of frames (latency): the quadtree of one level is the longest single piece of the call and needs only
    // that level, so FAST + quadtree of level l start on their own stream as soon as level l exists (level 0: at once),
    // beside the rest of the pyramid, the blur and the other levels.
    const bool split = !c->profile && frames <= 2 && P.nlevels > 1;
    if (split) {
        int rc = ensure_level_streams(c, s, P.nlevels);
        if (rc) return rc;
        fast_maps_prepare(P, io, frames, &s.fmaps);      // once, not per level
        for (int l = 0; l < P.nlevels; ++l) {
            if (l > 0) CU(c, orb_launch_pyramid_level(P, io, frames, l, st));
            CU(c, cudaEventRecord(s.lvl_ready[l], st));
            CU(c, cudaStreamWaitEvent(s.lvl[l], s.lvl_ready[l], 0));
            CU(c, orb_launch_fast(P, io, frames, s.lvl[l], &s.fmaps, l, l + 1));
            CU(c, orb_launch_octree(P, io, frames, s.lvl[l], l, l + 1));
            CU(c, cudaEventRecord(s.lvl_done[l], s.lvl[l]));
        }
    } else {
        StageScope t(c, ORBX_STAGE_PYRAMID, st);
        CU(c, orb_launch_pyramid(P, io, frames, st));
    }
    if (c->profile) {
        { StageScope t(c, ORBX_STAGE_FAST, st); CU(c, orb_launch_fast(P, io, frames, st, &s.fmaps)); }
        { StageScope t(c, ORBX_STAGE_BLUR, st); CU(c, orb_launch_blur(P, io, frames, st)); }
        { StageScope t(c, ORBX_STAGE_OCTREE, st); CU(c, orb_launch_octree(P, io, frames, st)); }
    } else {
        CU(c, cudaEventRecord(s.pyr_done, st));
        CU(c, cudaStreamWaitEvent(s.aux, s.pyr_done, 0));
        { StageScope t(c, ORBX_STAGE_BLUR, s.aux); CU(c, orb_launch_blur(P, io, frames, s.aux)); }
        CU(c, cudaEventRecord(s.blur_done, s.aux));
        if (!split) {
            { StageScope t(c, ORBX_STAGE_FAST, st); CU(c, orb_launch_fast(P, io, frames, st, &s.fmaps)); }
            { StageScope t(c, ORBX_STAGE_OCTREE, st); CU(c, orb_launch_octree(P, io, frames, st)); }
        }
        CU(c, cudaStreamWaitEvent(st, s.blur_done, 0));
        if (split) for (int l = 0; l < P.nlevels; ++l) CU(c, cudaStreamWaitEvent(st, s.lvl_done[l], 0));
    }
    { StageScope t(c, ORBX_STAGE_DESCRIBE, st); CU(c, orb_launch_describe(P, io, frames, st)); }
    if (sr) {
        // ComputeStereoMatches on the pairs (2p, 2p+1) of this chunk, straight from the buffers written above
        OrbStereoView V;
        memset(&V, 0, sizeof(V));
        V.nlevels = P.nlevels;
        for (int l = 0; l < P.nlevels; ++l) {
            const OrbLevel& L = P.lv[l];
            if (l == 0) { V.l[l] = d_img; V.r[l] = d_img + frame_stride; V.lstride[l] = V.rstride[l] = 2 * frame_stride; V.lpitch[l] = V.rpitch[l] = pitch; }
            else {
                V.l[l] = io.pyr + L.img_off; V.r[l] = io.pyr + P.pyr_bytes + L.img_off;
                V.lstride[l] = V.rstride[l] = 2 * (size_t)P.pyr_bytes; V.lpitch[l] = V.rpitch[l] = L.pitch;
            }
            V.w[l] = L.w; V.h[l] = L.h;
            V.scale[l] = c->params.scale[l]; V.inv_scale[l] = c->params.inv_scale[l];
        }
        V.kl = d_kps; V.kr = d_kps + cap;
        V.dl = (const uint32_t*)d_desc; V.dr = (const uint32_t*)(d_desc + (size_t)cap * 32);
        V.kstride = 2 * (size_t)cap;
        V.nl = d_n; V.nr = d_n + 1; V.nstride = 2;
        V.cap = cap;
        V.bf = sr->bf; V.mb = sr->bf / sr->fx;                                  // src/Frame.cc:121
        V.u_right = d_ur; V.depth = d_depth; V.ostride = (size_t)cap; V.n_stereo = d_ns;
        V.sad = (int*)s.st_sad.p; V.rec = (uint4*)s.st_sorted.p; V.row_start = (int*)s.st_rows.p;
        StageScope t(c, ORBX_STAGE_STEREO, st);
        CU(c, orb_launch_stereo(V, frames / 2, cap, st));
    }
    return ORBX_OK;
}

int run_batch(orbx_ctx* c, const uint8_t* imgs, size_t frame_stride, int batch, int w, int h, size_t pitch,
              orbx_kp* kps, uint8_t* desc, int cap, int* n_out, bool async_only, const StereoReq* sr = nullptr)
{
    if (!c) return ORBX_E_ARG;
    if (sr && (batch % 2 || cap > 65535 || !sr->u_right || !sr->depth || !sr->n_stereo || !(sr->fx > 0.f)))
        return fail(c, ORBX_E_ARG, "bad stereo argument");
    if (!imgs || batch <= 0 || w <= 0 || h <= 0) return fail(c, ORBX_E_EMPTY, "empty image");
    if (pitch < (size_t)w || cap <= 0 || !kps || !desc || !n_out) return fail(c, ORBX_E_ARG, "bad argument");
    if (batch > 1 && frame_stride < pitch * (size_t)(h - 1) + (size_t)w) return fail(c, ORBX_E_ARG, "frame_stride too small");
    CU(c, cudaSetDevice(c->device));
    int rc = ensure_plan(c, w, h);
    if (rc) return rc;
    const bool in_dev = is_device_ptr(imgs);
    const bool kps_dev = is_device_ptr(kps), desc_dev = is_device_ptr(desc), n_dev = is_device_ptr(n_out);
    const bool ur_dev = sr && is_device_ptr(sr->u_right), dep_dev = sr && is_device_ptr(sr->depth), ns_dev = sr && is_device_ptr(sr->n_stereo);
    // (the _async entry points take host buffers too: the call then only enqueues the copies and kernels of its chunks, and
    // several calls in flight overlap -- the next call's uploads run beside this call's last kernels and downloads)
    const bool stage_out = !(kps_dev && desc_dev && n_dev) || (sr && !(ur_dev && dep_dev && ns_dev));
    const bool piped = !in_dev || stage_out;           // any host buffer: rotate slots and overlap the copies
    // piped chunks: at most 128 MB of frames (1920x1080: 64 frames) and at least four chunks per call when the batch allows,
    // 16..128 frames; measured end to end: 1024 KITTI frames 64 / 96 / 128 / 160 = 82.6-84.4 / 85.3 / 86.1-86.4 / 83.0 k frames/s;
    // 256 EuRoC frames 32 / 48 / 64 / 96 / 128 = 91.9 / 95.1 / 95.7 / 86.1 / 82.6 k; 1024 HD frames 32 / 64 / 128 / 256 = 22.3 /
    // 24.0-24.3 / 22.7 / 20.5 k
    int auto_chunk = (int)(((size_t)128 << 20) / (pitch * (size_t)h));
    if (auto_chunk > batch / 4) auto_chunk = batch / 4;
    auto_chunk = auto_chunk < 16 ? 16 : auto_chunk > 128 ? 128 : auto_chunk;
    auto_chunk &= ~1;
    const int want = piped ? (c->chunk > 0 ? c->chunk : auto_chunk) : c->chunk_resident;
    int chunk = batch < want ? batch : want;
    if (sr && chunk % 2) chunk = chunk > 1 ? chunk - 1 : 2;      // a pair never straddles two chunks
    // Chunk schedule.  On the piped path the first chunk's upload and the last chunk's kernels + download have nothing to
    // overlap with, so the batch starts and ends with smaller chunks (chunk/4, chunk/2, chunk ... chunk, chunk/2, chunk/4)
    // when ORB_TAPER=1 asks for it (measured: 80.9 k against 81.5 k frames/s end to end, so off by default).
    std::vector<int> sched;
    {
        static const bool taper_on = []{ const char* e = getenv("ORB_TAPER"); return e && e[0] == '1'; }();
        int left = batch;
        const int q = (chunk / 4) & ~1, hlf = (chunk / 2) & ~1;
        const bool taper = taper_on && piped && q >= 2 && batch >= 6 * chunk;
        int tail[2] = { 0, 0 };
        if (taper) { sched.push_back(q); sched.push_back(hlf); left -= q + hlf; tail[0] = hlf; tail[1] = q; left -= hlf + q; }
        while (left > 0) { const int nfc = left < chunk ? left : chunk; sched.push_back(nfc); left -= nfc; }
        if (taper) { sched.push_back(tail[0]); sched.push_back(tail[1]); }
    }
    const int nchunks = (int)sched.size();
    const int nslot = nchunks < NSLOT ? nchunks : NSLOT;
    // Chunks alternate between compute streams only on the piped path (small chunks: one chunk's latency-bound sparse
    // kernels share the SMs with the next chunk's dense ones).  Device-resident chunks are large; running two of
    // them side by side measured slower than back to back (6.84 vs 6.40 ms per 512 frames in chunks of 256; again with the
    // round-2 kernels, 1024 frames: 512-frame chunks 105.0 k frames/s back to back vs 88.8 k on two streams, 256: 100.1 vs 90.6 k,
    // 128: 95.3 vs 88.1 k -- tools/multi_stream_exp.sh).
    static const bool force_multi = []{ const char* e = getenv("ORB_FORCE_MULTI"); return e && e[0] == '1'; }();   // experiment knob
    const bool multi = nslot > 1 && !c->profile && (piped || force_multi);
    const size_t in_frame_bytes = pitch * (size_t)h;
    const size_t frame_copy_bytes = pitch * (size_t)(h - 1) + (size_t)w;   // never read past the last row's pixels
    for (int s = 0; s < nslot; ++s) {
        rc = ensure_slot(c, c->slot[s], chunk, cap, !in_dev, in_frame_bytes, stage_out, sr != nullptr);
        if (rc) return rc;
    }
    const int ncomp = nslot < ORB_NCOMP ? nslot : ORB_NCOMP;   // slot s always runs on stream s % ncomp when ncomp divides nslot; else its events order it
    if (multi) {                                       // fork: the extra streams start after what is already queued
        CU(c, cudaEventRecord(c->fork_ev, c->stream));
        for (int k = 1; k < ncomp; ++k) CU(c, cudaStreamWaitEvent(c->xstream[k - 1], c->fork_ev, 0));
    }
    for (int i = 0, f0 = 0; i < nchunks; f0 += sched[i], ++i) {
        const int nf = sched[i];
        Slot& s = c->slot[i % nslot];
        cudaStream_t st = multi && i % ncomp ? c->xstream[i % ncomp - 1] : c->stream;
        const uint8_t* d_img; size_t d_stride;
        if (in_dev) { d_img = imgs + (size_t)f0 * frame_stride; d_stride = frame_stride; }
        else {
            // host frames -> staging on the copy stream, keeping the caller's row pitch; the
            // slot's previous kernels must have finished reading the staging buffer
            d_img = (const uint8_t*)s.img_stage.p; d_stride = in_frame_bytes;
            if (s.used) CU(c, cudaStreamWaitEvent(c->h2d_stream, s.compute_done, 0));
            if (frame_stride == in_frame_bytes || nf == 1)
                CU(c, cudaMemcpyAsync(s.img_stage.p, imgs + (size_t)f0 * frame_stride, (size_t)(nf - 1) * in_frame_bytes + frame_copy_bytes,
                                      cudaMemcpyHostToDevice, c->h2d_stream));
            else
                CU(c, cudaMemcpy2DAsync(s.img_stage.p, in_frame_bytes, imgs + (size_t)f0 * frame_stride, frame_stride,
                                        frame_copy_bytes, nf, cudaMemcpyHostToDevice, c->h2d_stream));
            CU(c, cudaEventRecord(s.h2d_done, c->h2d_stream));
            CU(c, cudaStreamWaitEvent(st, s.h2d_done, 0));
        }
        orbx_kp* dk = stage_out ? (orbx_kp*)s.out_kps.p : kps + (size_t)f0 * cap;
        uint8_t* dd = stage_out ? (uint8_t*)s.out_desc.p : desc + (size_t)f0 * cap * 32;
        int* dn = stage_out ? (int*)s.out_n.p : n_out + f0;
        if (stage_out && s.used) CU(c, cudaStreamWaitEvent(st, s.d2h_done, 0));   // output staging still being drained
        if (nslot > 1 && s.used && (!piped || multi)) CU(c, cudaStreamWaitEvent(st, s.compute_done, 0));   // the slot's work buffers may have served another stream
        float* dur = nullptr; float* ddep = nullptr; int* dns = nullptr;
        if (sr) {
            dur = stage_out ? (float*)s.out_ur.p : sr->u_right + (size_t)(f0 / 2) * cap;
            ddep = stage_out ? (float*)s.out_depth.p : sr->depth + (size_t)(f0 / 2) * cap;
            dns = stage_out ? (int*)s.out_ns.p : sr->n_stereo + f0 / 2;
        }
        rc = enqueue_chunk(c, s, st, d_img, d_stride, (int)pitch, nf, dk, dd, dn, cap, sr, dur, ddep, dns);
        if (rc) return rc;
        CU(c, cudaEventRecord(s.compute_done, st));
        if (stage_out) {
            const cudaMemcpyKind kk = kps_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
            const cudaMemcpyKind kd = desc_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
            const cudaMemcpyKind kn = n_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
            CU(c, cudaStreamWaitEvent(c->d2h_stream, s.compute_done, 0));
            CU(c, cudaMemcpyAsync(kps + (size_t)f0 * cap, dk, (size_t)nf * cap * sizeof(orbx_kp), kk, c->d2h_stream));
            CU(c, cudaMemcpyAsync(desc + (size_t)f0 * cap * 32, dd, (size_t)nf * cap * 32, kd, c->d2h_stream));
            CU(c, cudaMemcpyAsync(n_out + f0, dn, (size_t)nf * 4, kn, c->d2h_stream));
            if (sr) {
                const size_t np = (size_t)nf / 2;
                CU(c, cudaMemcpyAsync(sr->u_right + (size_t)(f0 / 2) * cap, dur, np * cap * 4, ur_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->d2h_stream));
                CU(c, cudaMemcpyAsync(sr->depth + (size_t)(f0 / 2) * cap, ddep, np * cap * 4, dep_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->d2h_stream));
                CU(c, cudaMemcpyAsync(sr->n_stereo + f0 / 2, dns, np * 4, ns_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->d2h_stream));
            }
            CU(c, cudaEventRecord(s.d2h_done, c->d2h_stream));
        }
        s.used = true;
        c->last_slot = i % nslot; c->last_first = f0; c->last_count = nf;
        c->last_img0 = d_img; c->last_img0_stride = d_stride; c->last_img0_pitch = (int)pitch;
    }
    if (multi) {                                       // join: later work on the context's stream sees every chunk
        for (int k = 1; k < ncomp; ++k) {
            CU(c, cudaEventRecord(c->join_ev[k - 1], c->xstream[k - 1]));
            CU(c, cudaStreamWaitEvent(c->stream, c->join_ev[k - 1], 0));
        }
    }
    if (async_only) return ORBX_OK;                    // the slots' events keep ordering the next call against this one
    rc = sync_all(c);
    if (rc) return rc;
    for (Slot& s : c->slot) s.used = false;
    collect_timers(c);
    if (!n_dev) for (int f = 0; f < batch; ++f) if (n_out[f] > cap) return fail(c, ORBX_E_CAPACITY, "frame %d: %d keypoints > capacity %d", f, n_out[f], cap);
    return ORBX_OK;
}

// ---- single-call path ------------------------------------------------------------------------------------------
// Capture the graph of one call: frames (1, or 2 = a stereo pair) of w x h with row pitch `pitch`.
int frame_path_build(orbx_ctx* c, int w, int h, int pitch, int frames, const StereoReq* sr, bool pyr)
{
    FramePath& F = c->fp;
    const OrbPlan& P = c->plan;
    if (F.exec) { cudaGraphExecDestroy(F.exec); F.exec = nullptr; }
    int cap = 0;
    for (int l = 0; l < P.nlevels; ++l) cap += P.lv[l].kp_cap;           // no frame can return more (orbx_max_keypoints)
    Slot& s = c->slot[NSLOT];
    int rc = ensure_slot(c, s, frames, cap, false, 0, false, sr != nullptr);
    if (rc) return rc;
    rc = sync_all(c);
    if (rc) return rc;
    const size_t in_bytes = (size_t)frames * pitch * h;
    CU(c, F.h_in.reserve(in_bytes)); CU(c, F.d_in.reserve(in_bytes + 256));
    CU(c, F.h_kps.reserve((size_t)frames * cap * sizeof(orbx_kp))); CU(c, F.d_kps.reserve((size_t)frames * cap * sizeof(orbx_kp)));
    CU(c, F.h_desc.reserve((size_t)frames * cap * 32)); CU(c, F.d_desc.reserve((size_t)frames * cap * 32));
    CU(c, F.h_n.reserve(64)); CU(c, F.d_n.reserve(64));
    if (sr) {
        CU(c, F.h_ur.reserve((size_t)cap * 4)); CU(c, F.d_ur.reserve((size_t)cap * 4));
        CU(c, F.h_dep.reserve((size_t)cap * 4)); CU(c, F.d_dep.reserve((size_t)cap * 4));
        CU(c, F.h_ns.reserve(64)); CU(c, F.d_ns.reserve(64));
    }
    size_t pad = 0;
    for (int l = 0; l < P.nlevels; ++l) {
        F.job.off[l] = (uint32_t)pad;
        F.job.pitch[l] = (P.lv[l].w + 2 * ORB_EDGE + 15) & ~15;
        pad += ((size_t)F.job.pitch[l] * (P.lv[l].h + 2 * ORB_EDGE) + 255) & ~(size_t)255;
    }
    F.job.frame_bytes = pad;
    if (pyr) { CU(c, F.h_pad.reserve(pad * frames)); CU(c, F.d_pad.reserve(pad * frames)); }

    rc = ensure_level_streams(c, s, P.nlevels);          // nothing may be created while the stream is capturing
    if (rc) return rc;
    cudaStream_t st = c->stream;
    CU(c, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    cudaGraph_t graph = nullptr;
    int erc = ORBX_OK;
    do {
        if (cudaMemcpyAsync(F.d_in.p, F.h_in.p, in_bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) { erc = ORBX_E_CUDA; break; }
        StereoReq dsr;
        if (sr) { dsr = *sr; dsr.u_right = (float*)F.d_ur.p; dsr.depth = (float*)F.d_dep.p; dsr.n_stereo = (int*)F.d_ns.p; }
        erc = enqueue_chunk(c, s, st, (const uint8_t*)F.d_in.p, (size_t)pitch * h, pitch, frames, (orbx_kp*)F.d_kps.p, (uint8_t*)F.d_desc.p,
                            (int*)F.d_n.p, cap, sr ? &dsr : nullptr, sr ? (float*)F.d_ur.p : nullptr, sr ? (float*)F.d_dep.p : nullptr, sr ? (int*)F.d_ns.p : nullptr);
        if (erc) break;
        bool ok = true;
        if (pyr) {
            // the padded levels leave on the copy stream as soon as the pyramid exists, beside FAST / quadtree / descriptors
            OrbBatch io;
            memset(&io, 0, sizeof(io));
            io.img0 = (const uint8_t*)F.d_in.p; io.img0_stride = (size_t)pitch * h; io.img0_pitch = pitch; io.pyr = (uint8_t*)s.pyr.p;
            ok = ok && cudaStreamWaitEvent(c->d2h_stream, s.pyr_done, 0) == cudaSuccess;
            ok = ok && orb_launch_border_levels(P, io, F.job, (uint8_t*)F.d_pad.p, frames, c->d2h_stream) == cudaSuccess;
            ok = ok && cudaMemcpyAsync(F.h_pad.p, F.d_pad.p, pad * frames, cudaMemcpyDeviceToHost, c->d2h_stream) == cudaSuccess;
            ok = ok && cudaEventRecord(s.d2h_done, c->d2h_stream) == cudaSuccess;
        }
        ok = ok && cudaMemcpyAsync(F.h_n.p, F.d_n.p, (size_t)frames * 4, cudaMemcpyDeviceToHost, st) == cudaSuccess;
        ok = ok && cudaMemcpyAsync(F.h_kps.p, F.d_kps.p, (size_t)frames * cap * sizeof(orbx_kp), cudaMemcpyDeviceToHost, st) == cudaSuccess;
        ok = ok && cudaMemcpyAsync(F.h_desc.p, F.d_desc.p, (size_t)frames * cap * 32, cudaMemcpyDeviceToHost, st) == cudaSuccess;
        if (sr) {
            ok = ok && cudaMemcpyAsync(F.h_ur.p, F.d_ur.p, (size_t)cap * 4, cudaMemcpyDeviceToHost, st) == cudaSuccess;
            ok = ok && cudaMemcpyAsync(F.h_dep.p, F.d_dep.p, (size_t)cap * 4, cudaMemcpyDeviceToHost, st) == cudaSuccess;
            ok = ok && cudaMemcpyAsync(F.h_ns.p, F.d_ns.p, 4, cudaMemcpyDeviceToHost, st) == cudaSuccess;
        }
        if (pyr) ok = ok && cudaStreamWaitEvent(st, s.d2h_done, 0) == cudaSuccess;
        if (!ok) erc = ORBX_E_CUDA;
    } while (0);
    const cudaError_t ec = cudaStreamEndCapture(st, &graph);
    if (erc || ec != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        return fail(c, ORBX_E_CUDA, "single-call graph capture failed (%s)", cudaGetErrorString(ec));
    }
    const cudaError_t ei = cudaGraphInstantiate(&F.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ei != cudaSuccess) { F.exec = nullptr; return fail(c, ORBX_E_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ei)); }
    F.w = w; F.h = h; F.pitch = pitch; F.frames = frames; F.cap = cap; F.nlevels = P.nlevels; F.stereo = sr != nullptr; F.pyr = pyr;
    F.bf = sr ? sr->bf : 0.f; F.fx = sr ? sr->fx : 0.f; F.stream = st;
    return ORBX_OK;
}

// One frame (or one pair) from HOST memory through the graph; the results are left in the path's pinned buffers.
int frame_path_run(orbx_ctx* c, const uint8_t* imgs, size_t frame_stride, int frames, int w, int h, size_t pitch, const StereoReq* sr, bool pyr)
{
    CU(c, cudaSetDevice(c->device));
    int rc = ensure_plan(c, w, h);
    if (rc) return rc;
    FramePath& F = c->fp;
    const bool fit = F.exec && F.w == w && F.h == h && F.pitch == (int)pitch && F.frames == frames && F.stereo == (sr != nullptr) &&
                     (F.pyr || !pyr) && F.stream == c->stream && (!sr || (F.bf == sr->bf && F.fx == sr->fx));
    if (!fit) {
        rc = frame_path_build(c, w, h, (int)pitch, frames, sr, pyr || F.pyr);
        if (rc) return rc;
    }
    const size_t fbytes = pitch * (size_t)h;
    // The frames always go through the path's own pinned buffer (13 us for a 1241x376 frame).  Re-pointing the graph's
    // upload node at a caller buffer that is itself pinned was tried: cudaGraphExecMemcpyNodeSetParams1D rejects a
    // source registered by someone else (torch's pinned allocator) with cudaErrorInvalidValue.
    const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    for (int f = 0; f < frames; ++f) memcpy((uint8_t*)F.h_in.p + (size_t)f * fbytes, imgs + (size_t)f * frame_stride, pitch * (size_t)(h - 1) + (size_t)w);
    const std::chrono::steady_clock::time_point t1 = std::chrono::steady_clock::now();
    CU(c, cudaGraphLaunch(F.exec, c->stream));
    const std::chrono::steady_clock::time_point t2 = std::chrono::steady_clock::now();
    CU(c, cudaStreamSynchronize(c->stream));
    const std::chrono::steady_clock::time_point t3 = std::chrono::steady_clock::now();
    F.last_us[0] = std::chrono::duration<double, std::micro>(t1 - t0).count();
    F.last_us[1] = std::chrono::duration<double, std::micro>(t2 - t1).count();
    F.last_us[2] = std::chrono::duration<double, std::micro>(t3 - t2).count();
    c->last_slot = NSLOT; c->last_first = 0; c->last_count = frames;
    c->last_img0 = (const uint8_t*)F.d_in.p; c->last_img0_stride = fbytes; c->last_img0_pitch = (int)pitch;
    return ORBX_OK;
}

// Copy the first min(n, cap) records of frame f from the path's pinned buffers to the caller's (host) arrays.
int frame_path_copy_out(orbx_ctx* c, int f, orbx_kp* kps, uint8_t* desc, int cap, int* n_out)
{
    const FramePath& F = c->fp;
    const int n = ((const int*)F.h_n.p)[f];
    *n_out = n;
    const int m = n < cap ? n : cap;
    memcpy(kps, (const orbx_kp*)F.h_kps.p + (size_t)f * F.cap, (size_t)m * sizeof(orbx_kp));
    memcpy(desc, (const uint8_t*)F.h_desc.p + (size_t)f * F.cap * 32, (size_t)m * 32);
    return n > cap ? fail(c, ORBX_E_CAPACITY, "frame %d: %d keypoints > capacity %d", f, n, cap) : ORBX_OK;
}

bool all_host(const void* a, const void* b, const void* d, const void* e)
{
    return !is_device_ptr(a) && !is_device_ptr(b) && !is_device_ptr(d) && !is_device_ptr(e);
}

int resident_frame(orbx_ctx* c, int frame, int level)
{
    if (!c || !c->have_plan || c->last_count == 0 || c->last_slot < 0) return -1;
    if (level < 0 || level >= c->plan.nlevels) return -1;
    const int rel = frame - c->last_first;
    if (rel < 0 || rel >= c->last_count) return -1;
    return rel;
}

int copy_out_2d(orbx_ctx* c, uint8_t* dst, size_t dst_pitch, const uint8_t* d_src, size_t src_pitch, int w, int h)
{
    const cudaMemcpyKind k = is_device_ptr(dst) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    CU(c, cudaMemcpy2DAsync(dst, dst_pitch, d_src, src_pitch, (size_t)w, (size_t)h, k, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return ORBX_OK;
}

} // namespace

int orb_ctx_levels(orbx_ctx* c, int frame, const uint8_t** ptr, int* pitch, int* w, int* h,
                   float* scale, float* inv_scale, int* nlevels, int* device)
{
    const int rel = resident_frame(c, frame, 0);
    if (rel < 0) return 1;
    if (cudaSetDevice(c->device) != cudaSuccess || sync_all(c)) return 1;
    const Slot& s = c->slot[c->last_slot];
    for (int l = 0; l < c->plan.nlevels; ++l) {
        const OrbLevel& L = c->plan.lv[l];
        if (ptr) ptr[l] = l == 0 ? c->last_img0 + (size_t)rel * c->last_img0_stride
                                 : (const uint8_t*)s.pyr.p + (size_t)rel * c->plan.pyr_bytes + L.img_off;
        if (pitch) pitch[l] = l == 0 ? c->last_img0_pitch : L.pitch;
        if (w) w[l] = L.w;
        if (h) h[l] = L.h;
        if (scale) scale[l] = c->params.scale[l];
        if (inv_scale) inv_scale[l] = c->params.inv_scale[l];
    }
    if (nlevels) *nlevels = c->plan.nlevels;
    if (device) *device = c->device;
    return 0;
}

extern "C" {

int orbx_create(orbx_ctx** out, int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int device)
{
    if (!out) return ORBX_E_ARG;
    *out = nullptr;
    OrbParams p;
    if (orb_params_init(&p, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)) return ORBX_E_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); return ORBX_E_CUDA; }
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return ORBX_E_CUDA; }
    orbx_ctx* c = new (std::nothrow) orbx_ctx();
    if (!c) return ORBX_E_ARG;
    c->params = p; c->device = device;
    bool ok = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&c->fork_ev, cudaEventDisableTiming) == cudaSuccess;
    for (int k = 0; k < NSLOT - 1; ++k)
        ok = ok && cudaStreamCreateWithFlags(&c->xstream[k], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&c->join_ev[k], cudaEventDisableTiming) == cudaSuccess;
    for (Slot& s : c->slot)
        ok = ok && cudaEventCreateWithFlags(&s.h2d_done, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&s.compute_done, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&s.d2h_done, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&s.pyr_done, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&s.blur_done, cudaEventDisableTiming) == cudaSuccess &&

             cudaStreamCreateWithFlags(&s.aux, cudaStreamNonBlocking) == cudaSuccess;
    if (!ok) { cudaGetLastError(); delete c; return ORBX_E_CUDA; }
    c->stream = c->own_stream;
    *out = c;
    return ORBX_OK;
}

void orbx_destroy(orbx_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    sync_all(c);
    collect_timers(c);
    for (cudaEvent_t e : c->free_events) cudaEventDestroy(e);
    for (Slot& s : c->slot) {
        s.release();
        if (s.h2d_done) cudaEventDestroy(s.h2d_done);
        if (s.compute_done) cudaEventDestroy(s.compute_done);
        if (s.d2h_done) cudaEventDestroy(s.d2h_done);
        if (s.pyr_done) cudaEventDestroy(s.pyr_done);
        if (s.blur_done) cudaEventDestroy(s.blur_done);
        for (int l = 0; l < ORB_MAX_LEVELS; ++l) {
            if (s.lvl_ready[l]) cudaEventDestroy(s.lvl_ready[l]);
            if (s.lvl_done[l]) cudaEventDestroy(s.lvl_done[l]);
            if (s.lvl[l]) cudaStreamDestroy(s.lvl[l]);
        }
        if (s.aux) cudaStreamDestroy(s.aux);
    }
    c->fp.release();
    c->taps.release(); c->border_tmp.release();
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->h2d_stream) cudaStreamDestroy(c->h2d_stream);
    if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
    for (cudaStream_t x : c->xstream) if (x) cudaStreamDestroy(x);
    for (cudaEvent_t e : c->join_ev) if (e) cudaEventDestroy(e);
    if (c->fork_ev) cudaEventDestroy(c->fork_ev);
    delete c;
}

int orbx_levels(const orbx_ctx* c) { return c ? c->params.nlevels : 0; }

int orbx_tables(const orbx_ctx* c, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int* per_level)
{
    if (!c) return ORBX_E_ARG;
    for (int i = 0; i < c->params.nlevels; ++i) {
        if (scale) scale[i] = c->params.scale[i];
        if (inv_scale) inv_scale[i] = c->params.inv_scale[i];
        if (sigma2) sigma2[i] = c->params.sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = c->params.inv_sigma2[i];
        if (per_level) per_level[i] = c->params.per_level[i];
    }
    return ORBX_OK;
}

int orbx_shape_supported(const orbx_ctx* c, int w, int h)
{
    if (!c) return 0;
    OrbPlan plan;
    return orb_plan_build(&c->params, w, h, &plan, nullptr) == 0;
}

int orbx_extract(orbx_ctx* c, const uint8_t* img, int w, int h, size_t pitch, orbx_kp* kps, uint8_t* desc, int capacity, int* n_out)
{
    // one host image: the single-call graph (stage timing needs the launch-by-launch path)
    if (c && !c->profile && img && w > 0 && h > 0 && pitch >= (size_t)w && capacity > 0 && kps && desc && n_out && all_host(img, kps, desc, n_out)) {
        const int rc = frame_path_run(c, img, pitch * (size_t)h, 1, w, h, pitch, nullptr, false);
        return rc ? rc : frame_path_copy_out(c, 0, kps, desc, capacity, n_out);
    }
    return run_batch(c, img, pitch * (size_t)(h > 0 ? h : 0), 1, w, h, pitch, kps, desc, capacity, n_out, false);
}

int orbx_debug_last_call_us(const orbx_ctx* c, double* us3)
{
    if (!c || !us3) return ORBX_E_ARG;
    for (int i = 0; i < 3; ++i) us3[i] = c->fp.last_us[i];
    return ORBX_OK;
}

int orbx_extract_frame(orbx_ctx* c, const uint8_t* image, int w, int h, size_t pitch, int with_pyramid, orbx_frame_out* out)
{
    if (!c || !out) return ORBX_E_ARG;
    if (!image || w <= 0 || h <= 0) return fail(c, ORBX_E_EMPTY, "empty image");
    if (pitch < (size_t)w || is_device_ptr(image)) return fail(c, ORBX_E_ARG, "bad argument (a host image is expected)");
    const int rc = frame_path_run(c, image, pitch * (size_t)h, 1, w, h, pitch, nullptr, with_pyramid != 0);
    if (rc) return rc;
    const FramePath& F = c->fp;
    memset(out, 0, sizeof(*out));
    out->n = ((const int*)F.h_n.p)[0];
    out->kps = (const orbx_kp*)F.h_kps.p;
    out->desc = (const uint8_t*)F.h_desc.p;
    if (with_pyramid) {
        out->nlevels = c->plan.nlevels;
        for (int l = 0; l < c->plan.nlevels; ++l) {
            out->level_w[l] = c->plan.lv[l].w; out->level_h[l] = c->plan.lv[l].h; out->level_pitch[l] = (size_t)F.job.pitch[l];
            out->level[l] = (const uint8_t*)F.h_pad.p + F.job.off[l] + (size_t)ORB_EDGE * F.job.pitch[l] + ORB_EDGE;
        }
    }
    return ORBX_OK;
}

int orbx_extract_batch(orbx_ctx* c, const uint8_t* imgs, size_t frame_stride, int batch, int w, int h, size_t pitch,
                       orbx_kp* kps, uint8_t* desc, int cap_per_frame, int* n_out)
{
    return run_batch(c, imgs, frame_stride, batch, w, h, pitch, kps, desc, cap_per_frame, n_out, false);
}

int orbx_extract_batch_async(orbx_ctx* c, const uint8_t* d_imgs, size_t frame_stride, int batch, int w, int h, size_t pitch,
                             orbx_kp* d_kps, uint8_t* d_desc, int cap_per_frame, int* d_n_out)
{
    return run_batch(c, d_imgs, frame_stride, batch, w, h, pitch, d_kps, d_desc, cap_per_frame, d_n_out, true);
}

int orbx_extract_stereo_batch(orbx_ctx* c, const uint8_t* imgs, size_t frame_stride, int pairs, int w, int h, size_t pitch,
                              orbx_kp* kps, uint8_t* desc, int cap_per_frame, int* n_out, float bf, float fx,
                              float* u_right, float* depth, int* n_stereo)
{
    const StereoReq sr = { bf, fx, u_right, depth, n_stereo };
    // one host pair: the single-call graph
    if (c && !c->profile && pairs == 1 && imgs && w > 0 && h > 0 && pitch >= (size_t)w && cap_per_frame > 0 && cap_per_frame <= 65535 && kps && desc && n_out &&
        u_right && depth && n_stereo && fx > 0.f && frame_stride >= pitch * (size_t)(h - 1) + (size_t)w && all_host(imgs, kps, desc, n_out) && all_host(u_right, depth, n_stereo, n_stereo)) {
        int rc = frame_path_run(c, imgs, frame_stride, 2, w, h, pitch, &sr, false);
        if (rc) return rc;
        rc = frame_path_copy_out(c, 0, kps, desc, cap_per_frame, n_out);
        const int rc1 = frame_path_copy_out(c, 1, kps + cap_per_frame, desc + (size_t)cap_per_frame * 32, cap_per_frame, n_out + 1);
        const FramePath& F = c->fp;
        const int m = n_out[0] < cap_per_frame ? n_out[0] : cap_per_frame;
        for (int i = 0; i < cap_per_frame; ++i) { u_right[i] = -1.0f; depth[i] = -1.0f; }
        memcpy(u_right, F.h_ur.p, (size_t)m * 4);
        memcpy(depth, F.h_dep.p, (size_t)m * 4);
        n_stereo[0] = ((const int*)F.h_ns.p)[0];
        return rc ? rc : rc1;
    }
    return run_batch(c, imgs, frame_stride, 2 * pairs, w, h, pitch, kps, desc, cap_per_frame, n_out, false, &sr);
}

int orbx_extract_stereo_batch_async(orbx_ctx* c, const uint8_t* d_imgs, size_t frame_stride, int pairs, int w, int h, size_t pitch,
                                    orbx_kp* d_kps, uint8_t* d_desc, int cap_per_frame, int* d_n_out, float bf, float fx,
                                    float* d_u_right, float* d_depth, int* d_n_stereo)
{
    const StereoReq sr = { bf, fx, d_u_right, d_depth, d_n_stereo };
    return run_batch(c, d_imgs, frame_stride, 2 * pairs, w, h, pitch, d_kps, d_desc, cap_per_frame, d_n_out, true, &sr);
}

int orbx_sync(orbx_ctx* c)
{
    if (!c) return ORBX_E_ARG;
    CU(c, cudaSetDevice(c->device));
    int rc = sync_all(c);
    if (rc) return rc;
    collect_timers(c);
    return ORBX_OK;
}

int orbx_pyramid_level(orbx_ctx* c, int frame, int level, int with_border, uint8_t* dst, size_t dst_pitch, int* w, int* h)
{
    if (!c) return ORBX_E_ARG;
    const int rel = resident_frame(c, frame, level);
    if (rel < 0) return fail(c, ORBX_E_ARG, "frame %d level %d is not resident", frame, level);
    const OrbLevel& L = c->plan.lv[level];
    const Slot& s = c->slot[c->last_slot];
    const int b = with_border ? ORB_EDGE : 0;
    if (w) *w = L.w + 2 * b;
    if (h) *h = L.h + 2 * b;
    if (!dst) return ORBX_OK;
    CU(c, cudaSetDevice(c->device));
    const uint8_t* src; int spitch;
    if (level == 0) { src = c->last_img0 + (size_t)rel * c->last_img0_stride; spitch = c->last_img0_pitch; }
    else { src = (const uint8_t*)s.pyr.p + (size_t)rel * c->plan.pyr_bytes + L.img_off; spitch = L.pitch; }
    if (!with_border) return copy_out_2d(c, dst, dst_pitch, src, (size_t)spitch, L.w, L.h);
    const int bw = L.w + 2 * b, bh = L.h + 2 * b;
    CU(c, c->border_tmp.reserve((size_t)bw * bh));
    CU(c, orb_launch_border(src, L.w, L.h, spitch, (uint8_t*)c->border_tmp.p, bw, b, c->stream));
    return copy_out_2d(c, dst, dst_pitch, (const uint8_t*)c->border_tmp.p, (size_t)bw, bw, bh);
}

int orbx_set_stream(orbx_ctx* c, void* cuda_stream)
{
    if (!c) return ORBX_E_ARG;
    CU(c, cudaSetDevice(c->device));
    int rc = sync_all(c);
    if (rc) return rc;
    collect_timers(c);
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return ORBX_OK;
}

void* orbx_stream(const orbx_ctx* c) { return c ? (void*)c->stream : nullptr; }

int orbx_set_chunk(orbx_ctx* c, int frames_per_chunk)
{
    if (!c || frames_per_chunk < 1) return ORBX_E_ARG;
    c->chunk = c->chunk_resident = frames_per_chunk;
    return ORBX_OK;
}

const char* orbx_last_error(const orbx_ctx* c) { return c ? c->err.c_str() : "null context"; }

int orbx_debug_blurred(orbx_ctx* c, int frame, int level, uint8_t* dst, size_t dst_pitch)
{
    if (!c || !dst) return ORBX_E_ARG;
    const int rel = resident_frame(c, frame, level);
    if (rel < 0) return fail(c, ORBX_E_ARG, "frame %d level %d is not resident", frame, level);
    const OrbLevel& L = c->plan.lv[level];
    CU(c, cudaSetDevice(c->device));
    return copy_out_2d(c, dst, dst_pitch, (const uint8_t*)c->slot[c->last_slot].blur.p + (size_t)rel * c->plan.blur_bytes + L.blur_off,
                       (size_t)L.pitch, L.w, L.h);
}

static int debug_packed(orbx_ctx* c, int frame, int level, bool after_octree, int* xys, int cap, int* n)
{
    if (!c || !n) return ORBX_E_ARG;
    const int rel = resident_frame(c, frame, level);
    if (rel < 0) return fail(c, ORBX_E_ARG, "frame %d level %d is not resident", frame, level);
    const OrbLevel& L = c->plan.lv[level];
    const Slot& s = c->slot[c->last_slot];
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    const int* counts = (const int*)s.counts.p + (after_octree ? (size_t)s.frames * ORB_MAX_LEVELS : 0);
    int cnt = 0;
    CU(c, cudaMemcpy(&cnt, counts + (size_t)rel * ORB_MAX_LEVELS + level, 4, cudaMemcpyDeviceToHost));
    *n = cnt;
    if (!xys || cnt == 0) return ORBX_OK;
    const int lim = after_octree ? L.kp_cap : L.cand_cap;
    const int m = cnt < lim ? cnt : lim;
    std::vector<uint32_t> tmp((size_t)m);
    const uint32_t* src = after_octree ? (const uint32_t*)s.lkp.p + (size_t)rel * c->plan.kp_per_frame + L.kp_off
                                       : (const uint32_t*)s.cand.p + (size_t)rel * c->plan.cand_per_frame + L.cand_off;
    CU(c, cudaMemcpy(tmp.data(), src, (size_t)m * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < m && i < cap; ++i) { xys[3 * i] = ORB_PX(tmp[i]); xys[3 * i + 1] = ORB_PY(tmp[i]); xys[3 * i + 2] = ORB_PS(tmp[i]); }
    return ORBX_OK;
}

int orbx_debug_candidates(orbx_ctx* c, int frame, int level, int* xys, int cap, int* n) { return debug_packed(c, frame, level, false, xys, cap, n); }
int orbx_debug_level_keypoints(orbx_ctx* c, int frame, int level, int* xys, int cap, int* n) { return debug_packed(c, frame, level, true, xys, cap, n); }

int orbx_profile(orbx_ctx* c, int enable)
{
    if (!c) return ORBX_E_ARG;
    c->profile = enable != 0;
    return ORBX_OK;
}

int orbx_stage_ms(orbx_ctx* c, float* ms, int* launches, int reset)
{
    if (!c) return ORBX_E_ARG;
    CU(c, cudaSetDevice(c->device));
    int rc = sync_all(c);
    if (rc) return rc;
    collect_timers(c);
    for (int i = 0; i < ORBX_STAGE_COUNT; ++i) {
        if (ms) ms[i] = c->stage_ms[i];
        if (launches) launches[i] = c->stage_launches[i];
        if (reset) { c->stage_ms[i] = 0.f; c->stage_launches[i] = 0; }
    }
    return ORBX_OK;
}

// Host-only view of the plan for tests (no GPU needed): level sizes, processed cells, quotas, roots.
int orbx_plan_describe(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int w, int h,
                       int* level_w, int* level_h, int* cells, int* quota, int* n_ini, int* cand_cap)
{
    OrbParams p;
    if (orb_params_init(&p, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)) return ORBX_E_ARG;
    OrbPlan plan;
    if (orb_plan_build(&p, w, h, &plan, nullptr)) return ORBX_E_SHAPE;
    for (int l = 0; l < nlevels; ++l) {
        if (level_w) level_w[l] = plan.lv[l].w;
        if (level_h) level_h[l] = plan.lv[l].h;
        if (cells) cells[l] = plan.lv[l].ncx * plan.lv[l].ncy;
        if (quota) quota[l] = plan.lv[l].quota;
        if (n_ini) n_ini[l] = plan.lv[l].nIni;
        if (cand_cap) cand_cap[l] = plan.lv[l].cand_cap;
    }
    return ORBX_OK;
}

// Upper bound on the keypoints operator() can return for one image of this shape: a level's DistributeOctTree list
// never exceeds max(N + 2, 4 * nIni) nodes (src/ORBextractor.cc:617-766) and every node yields one keypoint.
int orbx_launches_per_chunk(orbx_ctx* c, int stereo)
{
    if (!c) return -ORBX_E_ARG;
    if (c->plan.nlevels <= 0 || c->plan.w <= 0) return -ORBX_E_EMPTY;
    return orb_pyramid_launch_count(c->plan) + 4 + (stereo ? 3 : 0);
}

int orbx_max_keypoints(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int w, int h)
{
    OrbParams p;
    if (orb_params_init(&p, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)) return -ORBX_E_ARG;
    OrbPlan plan;
    if (orb_plan_build(&p, w, h, &plan, nullptr)) return -ORBX_E_SHAPE;
    int n = 0;
    for (int l = 0; l < nlevels; ++l) n += plan.lv[l].quota + 2 > 4 * plan.lv[l].nIni ? plan.lv[l].quota + 2 : 4 * plan.lv[l].nIni;
    return n;
}

} // extern "C"
