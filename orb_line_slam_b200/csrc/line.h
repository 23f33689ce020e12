// internal C++ interface of the line half (see line.cu)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/olf_abi.h"
namespace olf {
struct LineImpl;
// blocking_sync: waits sleep on the driver's interrupt instead of polling (batch rigs)
LineImpl* line_create(const olf_line_params* p, int device, cudaStream_t ext_stream = nullptr, bool blocking_sync = false);
void line_destroy(LineImpl* h);
int line_lsd_detect(LineImpl* h, const uint8_t* img, int w, int hgt, int stride, bool on_device, float* segs, int cap, int* n);
int line_lbd_compute(LineImpl* h, const uint8_t* img, int w, int hgt, int stride, const olf_keyline* kls, int n, uint8_t* desc);
int line_extract(LineImpl* h, const uint8_t* img, int w, int hgt, int stride, bool on_device, olf_keyline* kls, uint8_t* desc, int cap, int* n);
// a batch of images (one per handle, at most 8) through ONE chain of launches on the first handle's stream
int line_extract_batch(LineImpl* const* hs, int nimg, const uint8_t* const* imgs, int w, int hgt, int stride, bool on_device,
                       olf_keyline* const* kls, uint8_t* const* desc, int cap, int* n);
int line_trace(LineImpl* h, int* out, int max_rounds);
cudaStream_t line_stream(const LineImpl* h);
void line_last_stats(const LineImpl* h, int* out8);   // [0] rounds, [1] waves, [2] accepted regions
}
