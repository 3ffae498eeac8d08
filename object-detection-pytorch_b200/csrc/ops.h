// Internal (non-ABI) interfaces between the translation units of libssdbox.so.
#pragma once
#include "common.h"
#include "sizes.h"
#include "ssdbox_dev.cuh"

namespace ssdbox {

constexpr unsigned long long kBestInit = 0x00000000FFFFFFFFull;  // iou 0.0, prior 0 (~0u)

struct MatchArgs {
  const float* gt;
  const int32_t* gt_offsets;
  int gmax;
  const float* priors;
  long long prior_stride;  // floats between images (0 = shared priors)
  const float* anchors_xyxy;
  int B, P;
  float threshold;
  int binarize;
  RefineArgs rf;           // RefineDet fused: anchors decoded from arm_loc on the fly (anchors_xyxy / prior_stride unused)
};

// fills `best` with kBestInit and zeroes up to three uint32 ranges (one launch)
int launch_init(unsigned long long* best, size_t nbest, uint32_t* z0, size_t n0, uint32_t* z1, size_t n1,
                uint32_t* z2, size_t n2, cudaStream_t st);

// batched box_utils.match without the encode: lab/tidx (+ optional overlap) for every prior.
// `w.gt_best` / `w.done` must have been initialised by launch_init on the same stream.
int launch_match(const MatchArgs& a, const MatchWs& w, int16_t* lab, int16_t* tidx, float* overlap,
                 cudaStream_t st);

// conf_t / loc_t / match_idx materialisation from lab/tidx
int launch_materialize(const MatchArgs& a, float var0, float var1, const int16_t* lab, const int16_t* tidx,
                       float* loc_t, int64_t* conf_t, int32_t* match_idx, cudaStream_t st);

}  // namespace ssdbox
