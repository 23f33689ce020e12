// internal C++ interface of the ORB half (see orb.cu)
#pragma once
#include <cstdint>
#include "../../include/olf_abi.h"
namespace olf {
struct OrbImpl;
struct OrbDeviceView {          // device-resident pyramid of the last extract (mvImagePyramid), for the stereo matcher
    const uint8_t* pyr; int nlevels; int device;
    int w[OLF_MAX_LEVELS], h[OLF_MAX_LEVELS], pitch[OLF_MAX_LEVELS]; unsigned off[OLF_MAX_LEVELS];
    float scale[OLF_MAX_LEVELS], inv_scale[OLF_MAX_LEVELS];
};
OrbImpl* orb_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th, int device);
void orb_destroy(OrbImpl* h);
int orb_extract(OrbImpl* h, const uint8_t* img, int w, int hgt, int stride, bool on_device, olf_keypoint* kps, uint8_t* desc, int cap, int* n);
int orb_level_size(const OrbImpl* h, int level, int* w, int* hh);
int orb_get_level(OrbImpl* h, int level, uint8_t* dst, int dst_stride);
int orb_last_candidates(OrbImpl* h, int* out, int cap, int* n);
const OrbDeviceView orb_device_view(const OrbImpl* h);
void orb_scale_tables(const OrbImpl* h, const float** s, const float** is, const float** s2, const float** is2, const int** fpl, int* nlevels);
}
