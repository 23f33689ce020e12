// internal C++ interface of the ORB half (see orb.cu)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/olf_abi.h"
namespace olf {
struct OrbImpl;
struct OrbDeviceView {          // device-resident pyramid of the last extract (mvImagePyramid), for the stereo matcher
    const uint8_t* pyr; int nlevels; int device;
    int w[OLF_MAX_LEVELS], h[OLF_MAX_LEVELS], pitch[OLF_MAX_LEVELS]; unsigned off[OLF_MAX_LEVELS];
    float scale[OLF_MAX_LEVELS], inv_scale[OLF_MAX_LEVELS];
    // device-resident result of the last extract: keypoints, descriptors, n[0] = count (valid once the stream has passed them)
    const olf_keypoint* kps; const uint8_t* desc; const int* n; int cap;
};
// ext_stream: run on a stream owned by someone else (a rig puts each eye's ORB work on that eye's line-extractor stream:
// 2 streams per rig keep many rigs within the 32 hardware work queues); the owner must outlive this extractor
OrbImpl* orb_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th, int device, cudaStream_t ext_stream = nullptr);
void orb_destroy(OrbImpl* h);
cudaStream_t orb_stream(const OrbImpl* h);
int orb_extract(OrbImpl* h, const uint8_t* img, int w, int hgt, int stride, bool on_device, olf_keypoint* kps, uint8_t* desc, int cap, int* n);
// the two halves of orb_extract for a batch of images (one handle each): everything is enqueued on `s` and nothing waits;
// once `s` has been waited for, orb_collect copies one image's result out of its pinned staging
int orb_enqueue(OrbImpl* const* hs, int n, const uint8_t* const* imgs, int w, int hgt, int stride, bool on_device, int cap, cudaStream_t s);
int orb_collect(OrbImpl* h, olf_keypoint* kps, uint8_t* desc, int cap, int* n);
// parity hook: {cos, sin} of the device's glibc-exact sinf/cosf for the floats with bit patterns first + i * stride
int orb_trig_sweep(unsigned first, unsigned stride, unsigned count, float* cos_sin_out, int device);
int orb_level_size(const OrbImpl* h, int level, int* w, int* hh);
int orb_get_level(OrbImpl* h, int level, uint8_t* dst, int dst_stride);
int orb_last_candidates(OrbImpl* h, int* out, int cap, int* n);
const OrbDeviceView orb_device_view(const OrbImpl* h);
void orb_scale_tables(const OrbImpl* h, const float** s, const float** is, const float** s2, const float** is2, const int** fpl, int* nlevels);
}
