// Drop-in for include/LineExtractor.h:40-72 of the reference (class ORB_SLAM2::Lineextractor).
#pragma once
#include <vector>
#include "cv_min.h"
#include "../../include/olf_abi.h"

namespace ORB_SLAM2 {
class Lineextractor {
public:
    Lineextractor(int _lsd_nfeatures, double _llength_th, int _lsd_refine, double _lsd_scale, double _lsd_sigma_scale,
                  double _lsd_quant, double _lsd_ang_th, double _lsd_log_eps, double _lsd_density_th, int _lsd_n_bins, bool _bFLD = false);
    ~Lineextractor();
    Lineextractor(const Lineextractor&) = delete;
    Lineextractor& operator=(const Lineextractor&) = delete;
    void operator()(const cv::Mat& image, const cv::Mat& mask, std::vector<cv::line_descriptor::KeyLine>& keylines, cv::Mat& descriptors_line);
    olf_line* handle() { return h_; }
    static int device;
    static bool has_lines;       // stands for Config::hasLines() (src/LineExtractor.cc:37); the integrator wires it to Config
protected:
    int lsd_nfeatures; double min_line_length; bool bFLD;
    olf_line* h_;
};
}  // namespace ORB_SLAM2
