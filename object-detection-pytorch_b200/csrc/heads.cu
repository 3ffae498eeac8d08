// Head-output layout: the producer side of loc / conf.
// lib/models/ssd_v3.py:114-121 (same in rfb_net.py:213-220): every multibox head output
// [B, A_k*K, H_k, W_k] (NCHW) is permuted to NHWC, made contiguous, flattened and the layers are
// concatenated along dim 1, giving [B, P*K] = the [B,P,4] / [B,P,C] tensors of the box path.
// torch does that with one copy per layer (permute().contiguous()) plus one more for the cat; here
// ONE launch transposes every layer straight into its slice of the final tensor through
// 64 (channel) x 64 (h*w) shared-memory tiles: coalesced 256-byte reads along H*W, coalesced 256-byte
// writes along the channel axis, 16 independent loads in flight per thread; the channel tiles of one
// h*w range are handed to consecutive CTAs so that their partial 128-byte lines meet in L2.
// (Measured alternatives: 32 x 128 tiles 3 % slower; whole-channel tiles written back as one
// contiguous run 60 % slower.)
#include "ops.h"
#include "ssdbox_dev.cuh"

namespace ssdbox {

constexpr int kHeadTileC = 64;        // channels per tile
constexpr int kHeadTileS = 64;        // h*w positions per tile
constexpr int kHeadThreads = 256;     // 32 x 8

struct HeadsPlan {
  int num_layers, B;
  int channels[SSDBOX_MAX_HEADS];
  int hw[SSDBOX_MAX_HEADS];
  const float* src[SSDBOX_MAX_HEADS];
  long long out_off[SSDBOX_MAX_HEADS];      // first float of layer k inside one image's row
  long long tile_start[SSDBOX_MAX_HEADS + 1];   // first tile of layer k (tiles of all images of a layer are consecutive)
  int tiles_ch[SSDBOX_MAX_HEADS], tiles_hw[SSDBOX_MAX_HEADS];
  long long row_len;                        // floats per image in `out`
};

__global__ void __launch_bounds__(kHeadThreads) heads_to_rows_kernel(HeadsPlan p, float* __restrict__ out) {
  __shared__ float tile[kHeadTileC][kHeadTileS + 1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (long long t = blockIdx.x; t < p.tile_start[p.num_layers]; t += gridDim.x) {
    int k = 0;
    while (t >= p.tile_start[k + 1]) ++k;
    long long local = t - p.tile_start[k];
    const int per_img = p.tiles_ch[k] * p.tiles_hw[k];
    const int b = (int)(local / per_img);
    const int r = (int)(local - (long long)b * per_img);
    const int th = r / p.tiles_ch[k], tc = r - th * p.tiles_ch[k];     // channel tiles of one h*w range are consecutive:
                                                                       // their partial 128-byte lines meet in L2
    const int CH = p.channels[k], HW = p.hw[k];
    const int ch0 = tc * kHeadTileC, hw0 = th * kHeadTileS;
    const float* src = p.src[k] + (size_t)b * CH * HW;
    float* dst = out + (size_t)b * p.row_len + p.out_off[k];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kHeadTileC; i += kHeadThreads / 32) {
      const int ch = ch0 + ty + i;
#pragma unroll
      for (int j = 0; j < kHeadTileS; j += 32) {
        const int hw = hw0 + tx + j;
        if (ch < CH && hw < HW) tile[ty + i][tx + j] = __ldcs(src + (size_t)ch * HW + hw);
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int i = 0; i < kHeadTileS; i += kHeadThreads / 32) {
      const int hw = hw0 + ty + i;
#pragma unroll
      for (int j = 0; j < kHeadTileC; j += 32) {
        const int ch = ch0 + tx + j;
        if (ch < CH && hw < HW) dst[(size_t)hw * CH + ch] = tile[tx + j][ty + i];
      }
    }
  }
}

}  // namespace ssdbox

using namespace ssdbox;

extern "C" int ssdbox_heads_to_rows(const ssdbox_heads_cfg* cfg, float* out, ssdbox_stream_t stream) {
  SSDBOX_REQUIRE(cfg, SSDBOX_EINVAL, "heads: null cfg");
  SSDBOX_REQUIRE(cfg->num_layers >= 0 && cfg->num_layers <= SSDBOX_MAX_HEADS && cfg->B >= 0, SSDBOX_EINVAL,
                 "heads: bad layer count / batch");
  HeadsPlan p{};
  p.num_layers = cfg->num_layers;
  p.B = cfg->B;
  long long off = 0, tiles = 0;
  for (int k = 0; k < cfg->num_layers; ++k) {
    SSDBOX_REQUIRE(cfg->channels[k] >= 1 && cfg->hw[k] >= 1, SSDBOX_ESHAPE, "heads: layer %d has an empty shape", k);
    SSDBOX_REQUIRE(cfg->src[k] || cfg->B == 0, SSDBOX_EINVAL, "heads: layer %d has a null pointer", k);
    p.channels[k] = cfg->channels[k];
    p.hw[k] = cfg->hw[k];
    p.src[k] = cfg->src[k];
    p.out_off[k] = off;
    off += (long long)cfg->channels[k] * cfg->hw[k];
    p.tiles_ch[k] = (cfg->channels[k] + kHeadTileC - 1) / kHeadTileC;
    p.tiles_hw[k] = (cfg->hw[k] + kHeadTileS - 1) / kHeadTileS;
    p.tile_start[k] = tiles;
    tiles += (long long)cfg->B * p.tiles_ch[k] * p.tiles_hw[k];
  }
  p.tile_start[cfg->num_layers] = tiles;
  p.row_len = off;
  if (tiles == 0) return SSDBOX_OK;
  SSDBOX_REQUIRE(out, SSDBOX_EINVAL, "heads: null output");
  SSDBOX_REQUIRE(tiles < (1ll << 40), SSDBOX_ESHAPE, "heads: too many tiles");
  DevInfo dev;
  int rc = get_dev_info(&dev);
  if (rc) return rc;
  long long grid = tiles < (long long)dev.sm_count * 16 ? tiles : (long long)dev.sm_count * 16;
  heads_to_rows_kernel<<<(int)grid, kHeadThreads, 0, static_cast<cudaStream_t>(stream)>>>(p, out);
  SSDBOX_LAUNCH_OK("heads_to_rows_kernel");
  return SSDBOX_OK;
}
