// internal C++ interface of the ORB half (see orb.cu)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <functional>
#include "../../include/olf_abi.h"
namespace olf {
struct OrbImpl;
struct OrbDeviceView {          // device-resident pyramid of the last extract (mvImagePyramid), for the stereo matcher
    const uint8_t* pyr; int nlevels; int device;
    int w[OLF_MAX_LEVELS], h[OLF_MAX_LEVELS], pitch[OLF_MAX_LEVELS]; unsigned off[OLF_MAX_LEVELS];
    float scale[OLF_MAX_LEVELS], inv_scale[OLF_MAX_LEVELS];
};
// ext_stream: run on a stream owned by someone else (a rig puts each eye's ORB work on that eye's line-extractor stream:
// 2 streams per rig keep many rigs within the 32 hardware work queues); the owner must outlive this extractor
OrbImpl* orb_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th, int device, cudaStream_t ext_stream = nullptr);
void orb_destroy(OrbImpl* h);
cudaStream_t orb_stream(const OrbImpl* h);
// on_phase1_enqueued: called once the GPU phase before the quadtree is enqueued and marked (a rig uses it to let the line
// extractor of the same eye enqueue its long LSD chain BEHIND the ORB kernels on their shared stream)
int orb_extract(OrbImpl* h, const uint8_t* img, int w, int hgt, int stride, bool on_device, olf_keypoint* kps, uint8_t* desc, int cap, int* n,
                const std::function<void()>* on_phase1_enqueued = nullptr);
int orb_level_size(const OrbImpl* h, int level, int* w, int* hh);
int orb_get_level(OrbImpl* h, int level, uint8_t* dst, int dst_stride);
int orb_last_candidates(OrbImpl* h, int* out, int cap, int* n);
const OrbDeviceView orb_device_view(const OrbImpl* h);
void orb_scale_tables(const OrbImpl* h, const float** s, const float** is, const float** s2, const float** is2, const int** fpl, int* nlevels);
}
