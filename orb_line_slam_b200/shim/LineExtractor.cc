#include "LineExtractor.h"
#include <stdexcept>
#include <string>
#include <iostream>

namespace ORB_SLAM2 {
int Lineextractor::device = 0;
bool Lineextractor::has_lines = true;
static_assert(sizeof(cv::line_descriptor::KeyLine) == sizeof(olf_keyline), "KeyLine layout must match olf_keyline");

Lineextractor::Lineextractor(int _lsd_nfeatures, double _llength_th, int _lsd_refine, double _lsd_scale, double _lsd_sigma_scale,
                             double _lsd_quant, double _lsd_ang_th, double _lsd_log_eps, double _lsd_density_th, int _lsd_n_bins, bool _bFLD)
    : lsd_nfeatures(_lsd_nfeatures), min_line_length(_llength_th), bFLD(_bFLD) {
    olf_line_params p;
    p.lsd_nfeatures = _lsd_nfeatures; p.min_line_length = _llength_th; p.lsd_refine = _lsd_refine; p.lsd_scale = _lsd_scale;
    p.lsd_sigma_scale = _lsd_sigma_scale; p.lsd_quant = _lsd_quant; p.lsd_ang_th = _lsd_ang_th; p.lsd_log_eps = _lsd_log_eps;
    p.lsd_density_th = _lsd_density_th; p.lsd_n_bins = _lsd_n_bins;
    h_ = olf_line_create(&p, device);
    if (!h_) throw std::runtime_error(std::string("[Lineextractor] ") + olf_last_error());
}
Lineextractor::~Lineextractor() { olf_line_destroy(h_); }

void Lineextractor::operator()(const cv::Mat& img, const cv::Mat& /*mask*/, std::vector<cv::line_descriptor::KeyLine>& keylines, cv::Mat& descriptors_line) {
    keylines.clear();
    if (!has_lines || bFLD) return;                              // src/LineExtractor.cc:37-41, 68
    if (img.type() != CV_8UC1) throw std::runtime_error("Error, depth image!= 0");   // LSDDetector_custom.cpp:236-237
    const int cap = lsd_nfeatures > 0 ? lsd_nfeatures : 16384;
    keylines.resize(cap);
    std::vector<uint8_t> desc((size_t)cap * 32);
    int n = 0;
    const int rc = olf_line_extract(h_, img.ptr(), img.cols, img.rows, (int)img.step, (olf_keyline*)keylines.data(), desc.data(), cap, &n);
    if (rc != OLF_OK) { keylines.clear(); throw std::runtime_error(std::string("[Lineextractor] ") + olf_last_error()); }
    keylines.resize(n);
    if (n == 0) { std::cout << "Error: keypoint list is empty" << std::endl; return; }   // binary_descriptor_custom.cpp:556-560
    descriptors_line.create(n, 32, CV_8UC1);
    for (int i = 0; i < n; ++i) memcpy(descriptors_line.ptr(i), desc.data() + (size_t)i * 32, 32);
}
}  // namespace ORB_SLAM2
