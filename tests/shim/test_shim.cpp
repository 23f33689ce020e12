// Exercises the reference-shaped C++ class surfaces (orb_line_slam_b200/shim) end to end on a raw 8-bit image file:
//   test_shim <w> <h> <left.raw> <right.raw>   -> prints counts and FNV hashes that the pytest compares with the C-ABI path.
#include "../../orb_line_slam_b200/shim/ORBextractor.h"
#include "../../orb_line_slam_b200/shim/LineExtractor.h"
#include "../../orb_line_slam_b200/shim/LineMatcher.h"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace ORB_SLAM2;
static unsigned long long fnv(const void* p, size_t n, unsigned long long h = 1469598103934665603ull) {
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
static cv::Mat load(const char* path, int w, int h) {
    cv::Mat m(h, w, CV_8UC1);
    FILE* f = fopen(path, "rb");
    if (!f || fread(m.data, 1, (size_t)w * h, f) != (size_t)w * h) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
    fclose(f);
    return m;
}
int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: test_shim w h left.raw right.raw\n"); return 2; }
    const int w = atoi(argv[1]), h = atoi(argv[2]);
    try {
        cv::Mat L = load(argv[3], w, h), R = load(argv[4], w, h), mask;
        ORBextractor eL(1000, 1.2f, 8, 20, 7), eR(1000, 1.2f, 8, 20, 7);
        Lineextractor lL(200, 0.025, 0, 1.2, 0.6, 2.0, 22.5, 1.0, 0.6, 1024), lR(200, 0.025, 0, 1.2, 0.6, 2.0, 22.5, 1.0, 0.6, 1024);
        std::vector<cv::KeyPoint> kL, kR; cv::Mat dL, dR;
        eL(L, mask, kL, dL); eR(R, mask, kR, dR);
        std::vector<cv::line_descriptor::KeyLine> klL, klR; cv::Mat ldL, ldR;
        lL(L, mask, klL, ldL); lR(R, mask, klR, ldR);
        std::vector<float> uR, depth;
        ComputeStereoMatches(&eL, &eR, kL, dL, kR, dR, 47.90639384423901f, 435.2046959714599f, uR, depth);
        std::vector<std::pair<float, float>> disp; std::vector<double> le;
        ComputeStereoMatches_Lines(klL, ldL, klR, ldR, w, h, disp, le);
        std::vector<int> m12;
        const int nm = match(ldL, ldR, 0.9f, m12);
        unsigned long long hd = 1469598103934665603ull;
        for (int i = 0; i < dL.rows; ++i) hd = fnv(dL.ptr(i), 32, hd);
        unsigned long long hl = 1469598103934665603ull;
        for (int i = 0; i < ldL.rows; ++i) hl = fnv(ldL.ptr(i), 32, hl);
        eL.SyncImagePyramid();
        printf("nL %zu nR %zu mL %zu mR %zu desc %llu ldesc %llu uright %llu disp %llu match %d %llu pyr7 %dx%d levels %d\n", kL.size(), kR.size(), klL.size(), klR.size(),
               hd, hl, fnv(uR.data(), uR.size() * 4), fnv(disp.data(), disp.size() * 8), nm, fnv(m12.data(), m12.size() * 4),
               eL.mvImagePyramid[7].cols, eL.mvImagePyramid[7].rows, eL.GetLevels());
    } catch (const std::exception& e) { printf("EXCEPTION %s\n", e.what()); return 1; }
    return 0;
}
