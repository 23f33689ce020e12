// Whole-frame driver: the B200 replacement for Frame::Frame(stereo+lines) (reference src/Frame.cc:136-221).
// Four persistent host threads play the role of the reference's four std::threads (src/Frame.cc:164-171): each owns one
// extractor (own CUDA stream), so the left/right ORB and line pipelines overlap on the device.
#include "common.cuh"
#include "orb.h"
#include "line.h"
#include "match.h"
#include <thread>
#include <mutex>
#include <condition_variable>
#include <functional>
#include <atomic>

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
    OrbImpl* orb[2] = {nullptr, nullptr};
    LineImpl* line[2] = {nullptr, nullptr};
    Worker* workers[4] = {nullptr, nullptr, nullptr, nullptr};
    olf_frame_offsets off;
    std::string err[4];
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

FrontendImpl* frontend_create(const olf_frontend_params* p, int device) {
    if (!p || p->cap_points < p->nfeatures || p->cap_lines < 0) { set_last_error("olf_frontend_create: bad arguments"); return nullptr; }
    FrontendImpl* h = new FrontendImpl();
    h->P = *p; h->device = device;
    frame_layout(p->cap_points, p->cap_lines, &h->off);
    const bool share = getenv("OLF_RIG_STREAMS") == nullptr || atoi(getenv("OLF_RIG_STREAMS")) <= 2;
    for (int e = 0; e < 2; ++e) {
        if (p->has_lines) h->line[e] = line_create(&p->line, device);
        // each eye's ORB work rides on that eye's line stream (2 streams per rig); without lines the two eyes share one
        cudaStream_t ext = (p->has_lines && share) ? line_stream(h->line[e]) : nullptr;
        if (!(p->has_lines && !h->line[e]))
            h->orb[e] = orb_create(p->nfeatures, p->scale_factor, p->nlevels, p->ini_th_fast, p->min_th_fast, device, ext);
        if (!h->orb[e] || (p->has_lines && !h->line[e])) {
            for (int k = 0; k < 2; ++k) { orb_destroy(h->orb[k]); line_destroy(h->line[k]); }
            delete h; return nullptr;
        }
    }
    for (int i = 0; i < 4; ++i) h->workers[i] = new Worker();
    return h;
}
void frontend_destroy(FrontendImpl* h) {
    if (!h) return;
    for (int i = 0; i < 4; ++i) delete h->workers[i];
    for (int e = 0; e < 2; ++e) orb_destroy(h->orb[e]);    // before the line extractors whose streams they may borrow
    for (int e = 0; e < 2; ++e) line_destroy(h->line[e]);
    delete h;
}

int frontend_process(FrontendImpl* h, const uint8_t* img_l, const uint8_t* img_r, int w, int hgt, int stride, int on_device, void* result) {
    if (!h || !img_l || !img_r || !result || w <= 0 || hgt <= 0 || stride < w) { set_last_error("olf_frontend_process: bad arguments"); return OLF_ERR_ARG; }
    uint8_t* base = (uint8_t*)result;
    olf_frame_header* hd = (olf_frame_header*)base;
    memset(hd, 0, sizeof(*hd));
    hd->cap_points = h->P.cap_points; hd->cap_lines = h->P.cap_lines;
    const olf_frame_offsets& o = h->off;
    olf_keypoint* kps[2] = {(olf_keypoint*)(base + o.kps_l), (olf_keypoint*)(base + o.kps_r)};
    uint8_t* desc[2] = {base + o.desc_l, base + o.desc_r};
    olf_keyline* kls[2] = {(olf_keyline*)(base + o.kls_l), (olf_keyline*)(base + o.kls_r)};
    uint8_t* ldesc[2] = {base + o.ldesc_l, base + o.ldesc_r};
    const uint8_t* img[2] = {img_l, img_r};
    int n[2] = {0, 0}, m[2] = {0, 0};
    // ExtractORB(0|1), ExtractLine(0|1) on four threads (src/Frame.cc:164-171).  Per eye the two extractors share a stream:
    // the line thread enqueues its long LSD chain only after the ORB thread has enqueued (and marked) its pre-quadtree
    // kernels, so the ORB host work (quadtree) overlaps the LSD phases instead of queueing behind them.
    struct Gate { std::mutex m; std::condition_variable cv; bool open = false; } gate[2];
    std::function<void()> open_gate[2];
    for (int e = 0; e < 2; ++e) {
        Gate* g = &gate[e];
        open_gate[e] = [g]() { { std::lock_guard<std::mutex> l(g->m); g->open = true; } g->cv.notify_all(); };
        const std::function<void()>* hook = &open_gate[e];
        h->workers[e]->submit([=, &n, &h]() {
            const int rc = orb_extract(h->orb[e], img[e], w, hgt, stride, on_device != 0, kps[e], desc[e], h->P.cap_points, &n[e], hook);
            (*hook)();                                     // idempotent: covers the early-return paths
            if (rc) h->err[e] = olf_last_error();
            return rc;
        });
        if (h->P.has_lines)
            h->workers[2 + e]->submit([=, &m, &h]() {
                { std::unique_lock<std::mutex> l(g->m); g->cv.wait(l, [g] { return g->open; }); }
                const int rc = line_extract(h->line[e], img[e], w, hgt, stride, on_device != 0, kls[e], ldesc[e], h->P.cap_lines, &m[e]);
                if (rc) h->err[2 + e] = olf_last_error();
                return rc;
            });
    }
    int rc = OLF_OK;
    for (int i = 0; i < 4; ++i) {
        if (i >= 2 && !h->P.has_lines) break;
        const int r = h->workers[i]->wait();
        if (r && !rc) { rc = r; set_last_error(h->err[i]); }
    }
    hd->n_l = n[0]; hd->n_r = n[1]; hd->m_l = m[0]; hd->m_r = m[1];
    if (!rc && n[0] > 0)       // if(mvKeys.empty()) return;  (src/Frame.cc:176)
        rc = stereo_points(h->orb[0], h->orb[1], kps[0], desc[0], n[0], kps[1], desc[1], n[1], h->P.cam.bf, h->P.cam.fx,
                           (float*)(base + o.u_right), (float*)(base + o.depth));
    if (!rc && n[0] > 0 && h->P.has_lines)
        rc = stereo_lines(kls[0], ldesc[0], m[0], kls[1], ldesc[1], m[1], w, hgt, &h->P.line_match,
                          (int*)(base + o.lmatch), (float*)(base + o.ldisp), (double*)(base + o.lle), h->device);
    hd->status = rc;
    return rc;
}

}  // namespace olf
using namespace olf;
extern "C" {
int olf_frame_layout(int cap_points, int cap_lines, olf_frame_offsets* out) { return frame_layout(cap_points, cap_lines, out); }
olf_frontend* olf_frontend_create(const olf_frontend_params* p, int device) { return (olf_frontend*)frontend_create(p, device); }
void olf_frontend_destroy(olf_frontend* h) { frontend_destroy((FrontendImpl*)h); }
int olf_frontend_process(olf_frontend* h, const uint8_t* img_l, const uint8_t* img_r, int width, int height, int stride, int on_device, void* result) {
    return frontend_process((FrontendImpl*)h, img_l, img_r, width, height, stride, on_device, result);
}
olf_orb* olf_frontend_orb(olf_frontend* h, int eye) { return h && eye >= 0 && eye < 2 ? (olf_orb*)((FrontendImpl*)h)->orb[eye] : nullptr; }
olf_line* olf_frontend_line(olf_frontend* h, int eye) { return h && eye >= 0 && eye < 2 ? (olf_line*)((FrontendImpl*)h)->line[eye] : nullptr; }
}
