// Whole-frame driver: the B200 replacement for Frame::Frame(stereo+lines) (reference src/Frame.cc:136-221).
// The reference runs ExtractORB(L|R) and ExtractLine(L|R) on four std::threads (src/Frame.cc:164-171).  Here the CALLING
// thread is the only host thread of a batch rig: the ORB extraction of every image of the call and the stereo point
// matcher are one asynchronous chain on the rig's first stream (no host step inside: quadtree and trig run on the device),
// the line extraction of all images is one batched chain on the second stream whose three host steps (libm trig of the
// O(#regions) rectangles, KeyLine construction, top-N) run on the calling thread while the first stream works.
// A single-frame rig (latency) keeps ONE helper thread so that the two eyes' LSD chains run side by side.
#include "common.cuh"
#include "orb.h"
#include "line.h"
#include "match.h"
#include <thread>
#include <mutex>
#include <condition_variable>
#include <functional>
#include <chrono>

namespace olf {

class Worker {
public:
    Worker() : th_([this] { run(); }) {}
    ~Worker() { { std::lock_guard<std::mutex> l(m_); stop_ = true; } cv_.notify_all(); th_.join(); }
    void submit(std::function<int()> f) { { std::lock_guard<std::mutex> l(m_); job_ = std::move(f); has_ = true; done_ = false; } cv_.notify_all(); }
    int wait() { std::unique_lock<std::mutex> l(m_); cv_done_.wait(l, [this] { return done_; }); return rc_; }
private:
    void run() {
        for (;;) {
            std::function<int()> f;
            { std::unique_lock<std::mutex> l(m_); cv_.wait(l, [this] { return has_ || stop_; }); if (stop_) return; f = std::move(job_); has_ = false; }
            const int rc = f();
            { std::lock_guard<std::mutex> l(m_); rc_ = rc; done_ = true; }
            cv_done_.notify_all();
        }
    }
    std::mutex m_; std::condition_variable cv_, cv_done_;
    std::function<int()> job_; bool has_ = false, stop_ = false, done_ = true; int rc_ = 0;
    std::thread th_;
};

struct FrontendImpl {
    olf_frontend_params P;
    int device;
    int max_frames = 1;                  // frames per olf_frontend_process_batch call
    // slot 2*f + eye.  TWO streams per rig: every ORB extractor (and the stereo point matcher) rides on the first ORB extractor's
    // stream; the line extractors of a batch rig are driven through the first one's stream by ONE batched chain of launches
    std::vector<OrbImpl*> orb;
    std::vector<LineImpl*> line;
    StereoWs sws[OLF_MAX_BATCH_FRAMES];
    SyncEvent ev_orb;
    MatchCtx* mctx = nullptr;            // scratch of the stereo line matcher (owned by the rig, not by the calling thread)
    Worker* worker = nullptr;            // single-frame rig only: the right eye's line extraction
    olf_frame_offsets off;
    std::string err;
    int timing[8] = {0};                 // host view of the last call, microseconds: ORB enqueue, lines, ORB wait, collect, stereo lines, total
};

static uint64_t a64(uint64_t v) { return (v + 63) / 64 * 64; }
int frame_layout(int cap_p, int cap_l, olf_frame_offsets* o) {
    if (!o || cap_p < 0 || cap_l < 0) return OLF_ERR_ARG;
    uint64_t p = a64(sizeof(olf_frame_header));
    o->kps_l = p; p = a64(p + (uint64_t)cap_p * sizeof(olf_keypoint));
    o->desc_l = p; p = a64(p + (uint64_t)cap_p * 32);
    o->kps_r = p; p = a64(p + (uint64_t)cap_p * sizeof(olf_keypoint));
    o->desc_r = p; p = a64(p + (uint64_t)cap_p * 32);
    o->u_right = p; p = a64(p + (uint64_t)cap_p * 4);
    o->depth = p; p = a64(p + (uint64_t)cap_p * 4);
    o->kls_l = p; p = a64(p + (uint64_t)cap_l * sizeof(olf_keyline));
    o->ldesc_l = p; p = a64(p + (uint64_t)cap_l * 32);
    o->kls_r = p; p = a64(p + (uint64_t)cap_l * sizeof(olf_keyline));
    o->ldesc_r = p; p = a64(p + (uint64_t)cap_l * 32);
    o->lmatch = p; p = a64(p + (uint64_t)cap_l * 4);
    o->ldisp = p; p = a64(p + (uint64_t)cap_l * 8);
    o->lle = p; p = a64(p + (uint64_t)cap_l * 24);
    o->total = p;
    return OLF_OK;
}

void frontend_destroy(FrontendImpl* h);
FrontendImpl* frontend_create(const olf_frontend_params* p, int device, int max_frames) {
    if (!p || p->cap_points < p->nfeatures || p->cap_lines < 0 || max_frames < 1 || max_frames > OLF_MAX_BATCH_FRAMES) {
        set_last_error("olf_frontend_create: bad arguments (1..8 frames per batch)"); return nullptr;
    }
    FrontendImpl* h = new FrontendImpl();
    h->P = *p; h->device = device; h->max_frames = max_frames;
    frame_layout(p->cap_points, p->cap_lines, &h->off);
    const bool batch = max_frames > 1;
    // A device has 32 hardware work queues and streams beyond that alias (a long LSD chain then blocks an unrelated rig), so
    // streams are what limits the number of rigs in flight.  A batch rig therefore uses ONE stream by default: the short ORB +
    // stereo chain of the call runs first, the batched line chain behind it (OLF_RIG_STREAMS=2: ORB on a stream of its own).
    const bool one_stream = batch && p->has_lines && !(getenv("OLF_RIG_STREAMS") && atoi(getenv("OLF_RIG_STREAMS")) == 2);
    bool ok = true;
    for (int k = 0; k < 2 * max_frames && ok; ++k) {
        // a single-frame rig (latency matters) keeps one stream per eye and extracts the two eyes' lines side by side; a batch
        // rig (throughput matters) drives every line extractor through ONE stream and one batched chain of launches
        if (p->has_lines) { h->line.push_back(line_create(&p->line, device, (h->line.empty() || !batch) ? nullptr : line_stream(h->line[0]), batch)); ok = h->line.back() != nullptr; }
        if (ok) {
            h->orb.push_back(orb_create(p->nfeatures, p->scale_factor, p->nlevels, p->ini_th_fast, p->min_th_fast, device,
                                        one_stream ? line_stream(h->line[0]) : (h->orb.empty() ? nullptr : orb_stream(h->orb[0]))));
            ok = h->orb.back() != nullptr;
        }
    }
    ok = ok && h->ev_orb.create(batch) == cudaSuccess;
    if (ok) h->mctx = match_ctx_create(device);
    if (!ok) {
        const std::string e = olf_last_error();
        frontend_destroy(h); set_last_error(e); return nullptr;
    }
    if (!batch && p->has_lines) h->worker = new Worker();
    return h;
}
void frontend_destroy(FrontendImpl* h) {
    if (!h) return;
    delete h->worker;
    cudaSetDevice(h->device);
    if (!h->orb.empty()) cudaStreamSynchronize(orb_stream(h->orb[0]));
    for (int f = 0; f < OLF_MAX_BATCH_FRAMES; ++f) stereo_ws_release(&h->sws[f]);
    h->ev_orb.destroy();
    if (h->mctx) match_ctx_destroy(h->mctx);
    for (size_t k = h->orb.size(); k-- > 0;) orb_destroy(h->orb[k]);     // the borrowers before the owner of the stream
    for (size_t k = h->line.size(); k-- > 0;) line_destroy(h->line[k]);
    delete h;
}

// Frame::Frame(stereo+lines) (src/Frame.cc:136-221) for `nframes` independent stereo frames at once
int frontend_process_batch(FrontendImpl* h, const uint8_t* const* img_l, const uint8_t* const* img_r, int nframes, int w, int hgt, int stride,
                           int on_device, void* const* results) {
    if (!h || !img_l || !img_r || !results || nframes < 1 || nframes > h->max_frames || w <= 0 || hgt <= 0 || stride < w) {
        set_last_error("olf_frontend_process: bad arguments"); return OLF_ERR_ARG;
    }
    for (int f = 0; f < nframes; ++f) if (!img_l[f] || !img_r[f] || !results[f]) { set_last_error("olf_frontend_process: bad arguments"); return OLF_ERR_ARG; }
    const olf_frame_offsets& o = h->off;
    const int nimg = 2 * nframes, capP = h->P.cap_points;
    std::vector<int> n(nimg, 0), m(nimg, 0);
    std::vector<const uint8_t*> img(nimg);
    std::vector<olf_keypoint*> kps(nimg); std::vector<uint8_t*> desc(nimg), ldesc(nimg); std::vector<olf_keyline*> kls(nimg);
    for (int f = 0; f < nframes; ++f) {
        uint8_t* base = (uint8_t*)results[f];
        olf_frame_header* hd = (olf_frame_header*)base;
        memset(hd, 0, sizeof(*hd));
        hd->cap_points = capP; hd->cap_lines = h->P.cap_lines;
        img[2 * f] = img_l[f]; img[2 * f + 1] = img_r[f];
        kps[2 * f] = (olf_keypoint*)(base + o.kps_l); kps[2 * f + 1] = (olf_keypoint*)(base + o.kps_r);
        desc[2 * f] = base + o.desc_l; desc[2 * f + 1] = base + o.desc_r;
        kls[2 * f] = (olf_keyline*)(base + o.kls_l); kls[2 * f + 1] = (olf_keyline*)(base + o.kls_r);
        ldesc[2 * f] = base + o.ldesc_l; ldesc[2 * f + 1] = base + o.ldesc_r;
    }
    auto now = [] { return std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const long long t0 = now();
    // ExtractORB(0|1) of every frame + ComputeStereoMatches: one asynchronous chain on the first stream
    cudaStream_t so = orb_stream(h->orb[0]);
    int rc = orb_enqueue(h->orb.data(), nimg, img.data(), w, hgt, stride, on_device != 0, capP, so);
    for (int f = 0; f < nframes && !rc; ++f)
        rc = stereo_points_enqueue(&h->sws[f], h->orb[2 * f], h->orb[2 * f + 1], h->P.cam.bf, h->P.cam.fx, capP, so);
    if (!rc && h->ev_orb.record(so) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "cudaEventRecord", __FILE__, __LINE__);
    if (rc) { cudaStreamSynchronize(so); return rc; }
    const long long t1 = now();
    // ExtractLine(0|1) of every frame on the second stream while the first one works
    int rc_l = OLF_OK;
    if (h->P.has_lines) {
        if (h->worker) {           // one frame: the two eyes side by side on their own streams
            h->worker->submit([=, &m, &kls, &ldesc, &img]() {
                const int r = line_extract(h->line[1], img[1], w, hgt, stride, on_device != 0, kls[1], ldesc[1], h->P.cap_lines, &m[1]);
                if (r) h->err = olf_last_error();
                return r;
            });
            rc_l = line_extract(h->line[0], img[0], w, hgt, stride, on_device != 0, kls[0], ldesc[0], h->P.cap_lines, &m[0]);
            const int r1 = h->worker->wait();
            if (!rc_l && r1) { rc_l = r1; set_last_error(h->err); }
        } else
            rc_l = line_extract_batch(h->line.data(), nimg, img.data(), w, hgt, stride, on_device != 0, kls.data(), ldesc.data(), h->P.cap_lines, m.data());
    }
    const std::string err_l = rc_l ? olf_last_error() : "";
    const long long t2 = now();
    if (h->ev_orb.wait() != cudaSuccess) return cuda_fail(cudaGetLastError(), "ORB chain", __FILE__, __LINE__);
    const long long t3 = now();
    for (int k = 0; k < nimg && !rc; ++k) rc = orb_collect(h->orb[k], kps[k], desc[k], capP, &n[k]);
    const long long t4 = now();
    if (!rc && rc_l) { rc = rc_l; set_last_error(err_l); }
    match_use_stream(h->P.has_lines ? line_stream(h->line[0]) : so);
    match_use_ctx(h->mctx);
    // stereo association of the lines: the frames of the call are enqueued together and waited for once
    const olf_keyline* b_kl[OLF_MAX_BATCH_FRAMES]; const uint8_t* b_dl[OLF_MAX_BATCH_FRAMES]; const olf_keyline* b_kr[OLF_MAX_BATCH_FRAMES]; const uint8_t* b_dr[OLF_MAX_BATCH_FRAMES];
    int b_n1[OLF_MAX_BATCH_FRAMES], b_n2[OLF_MAX_BATCH_FRAMES], nb = 0;
    int* b_m[OLF_MAX_BATCH_FRAMES]; float* b_d[OLF_MAX_BATCH_FRAMES]; double* b_le[OLF_MAX_BATCH_FRAMES];
    for (int f = 0; f < nframes; ++f) {
        uint8_t* base = (uint8_t*)results[f];
        olf_frame_header* hd = (olf_frame_header*)base;
        const int kl = 2 * f, kr = 2 * f + 1;
        hd->n_l = n[kl]; hd->n_r = n[kr]; hd->m_l = m[kl]; hd->m_r = m[kr];
        if (!rc && n[kl] > 0) {     // if(mvKeys.empty()) return;  (src/Frame.cc:176)
            memcpy(base + o.u_right, h->sws[f].out.p, (size_t)n[kl] * 4);
            memcpy(base + o.depth, h->sws[f].out.p + h->sws[f].cap, (size_t)n[kl] * 4);
        }
        if (!rc && n[kl] > 0 && h->P.has_lines) {
            b_kl[nb] = kls[kl]; b_dl[nb] = ldesc[kl]; b_n1[nb] = m[kl]; b_kr[nb] = kls[kr]; b_dr[nb] = ldesc[kr]; b_n2[nb] = m[kr];
            b_m[nb] = (int*)(base + o.lmatch); b_d[nb] = (float*)(base + o.ldisp); b_le[nb] = (double*)(base + o.lle); ++nb;
        } else if (h->P.has_lines) {  // ComputeStereoMatches_Lines not reached: every line unmatched (mvDisparity_l = (-1,-1), mvle_l = 0)
            int* lm = (int*)(base + o.lmatch); float* ld = (float*)(base + o.ldisp); double* le = (double*)(base + o.lle);
            for (int i = 0; i < m[kl]; ++i) { lm[i] = -1; ld[2 * i] = ld[2 * i + 1] = -1.f; le[3 * i] = le[3 * i + 1] = le[3 * i + 2] = 0; }
        }
    }
    if (!rc && nb > 0) rc = stereo_lines_batch(nb, b_kl, b_dl, b_n1, b_kr, b_dr, b_n2, w, hgt, &h->P.line_match, b_m, b_d, b_le, h->device);
    for (int f = 0; f < nframes; ++f) ((olf_frame_header*)results[f])->status = rc;
    match_use_stream(nullptr); match_use_ctx(nullptr);
    const long long t5 = now();
    h->timing[0] = (int)(t1 - t0); h->timing[1] = (int)(t2 - t1); h->timing[2] = (int)(t3 - t2); h->timing[3] = (int)(t4 - t3); h->timing[4] = (int)(t5 - t4); h->timing[5] = (int)(t5 - t0);
    return rc;
}
int frontend_process(FrontendImpl* h, const uint8_t* img_l, const uint8_t* img_r, int w, int hgt, int stride, int on_device, void* result) {
    if (!img_l || !img_r || !result) { set_last_error("olf_frontend_process: bad arguments"); return OLF_ERR_ARG; }
    return frontend_process_batch(h, &img_l, &img_r, 1, w, hgt, stride, on_device, &result);
}

}  // namespace olf
using namespace olf;
extern "C" {
int olf_frame_layout(int cap_points, int cap_lines, olf_frame_offsets* out) { return frame_layout(cap_points, cap_lines, out); }
olf_frontend* olf_frontend_create(const olf_frontend_params* p, int device) { return (olf_frontend*)frontend_create(p, device, 1); }
olf_frontend* olf_frontend_create_batch(const olf_frontend_params* p, int device, int max_frames) { return (olf_frontend*)frontend_create(p, device, max_frames); }
int olf_frontend_process_batch(olf_frontend* h, const uint8_t* const* img_l, const uint8_t* const* img_r, int nframes, int width, int height, int stride,
                               int on_device, void* const* results) {
    return frontend_process_batch((FrontendImpl*)h, img_l, img_r, nframes, width, height, stride, on_device, results);
}
void olf_frontend_destroy(olf_frontend* h) { frontend_destroy((FrontendImpl*)h); }
int olf_frontend_process(olf_frontend* h, const uint8_t* img_l, const uint8_t* img_r, int width, int height, int stride, int on_device, void* result) {
    return frontend_process((FrontendImpl*)h, img_l, img_r, width, height, stride, on_device, result);
}
int olf_frontend_last_timing(const olf_frontend* h, int* out8) { if (!h || !out8) return OLF_ERR_ARG; for (int i = 0; i < 8; ++i) out8[i] = ((const FrontendImpl*)h)->timing[i]; return OLF_OK; }
olf_orb* olf_frontend_orb(olf_frontend* h, int eye) { return h && eye >= 0 && eye < (int)((FrontendImpl*)h)->orb.size() ? (olf_orb*)((FrontendImpl*)h)->orb[eye] : nullptr; }
olf_line* olf_frontend_line(olf_frontend* h, int eye) { return h && eye >= 0 && eye < (int)((FrontendImpl*)h)->line.size() ? (olf_line*)((FrontendImpl*)h)->line[eye] : nullptr; }
}
