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
#include <cuda.h>      // CUtensorMap types only: the encoder is fetched through cudaGetDriverEntryPoint (no -lcuda)

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

// ---- the large layers: 2-D tiled TMA loads into a swizzled ring (H*W % 4 == 0) ---------------------------------
// A work item is ALL channels x 32 consecutive h*w positions of one image: its output is one contiguous run of
// 32 * CH floats.  One producer thread requests the item as one or two {32 positions x <= 256 channels} boxes of the
// layer's {HW, CH, B} tensor map (SWIZZLE_128B: the 16-byte chunk c of channel row r lands at chunk c ^ (r & 7));
// consumer warp q reads chunk q of 32 channel rows with one conflict-free LDS.128 per lane (lanes = consecutive
// channels, the swizzle spreads them over the banks) and writes four coalesced 128-byte pieces of four output rows.
// No register staging of the loads, kHeadTmaStages items in flight per SM, no partial lines inside a run.
constexpr int kHeadTmaPos = 32;          // positions per item = one 128-byte swizzle row
constexpr int kHeadTmaWarps = 8;         // consumer warps = 16-byte chunks per row
constexpr int kHeadTmaLayers = 8;
constexpr int kHeadTmaMaxStages = 6;
constexpr int kHeadTmaMaxRows = 576;     // channel rows of one item (73.7 KB): three stages fit

struct HeadsTmaPlan {
  CUtensorMap map[kHeadTmaLayers];
  int num_layers, stages;
  uint32_t stage_bytes;
  int channels[kHeadTmaLayers], hw[kHeadTmaLayers], tiles_hw[kHeadTmaLayers];
  int box_rows[kHeadTmaLayers], boxes[kHeadTmaLayers];
  long long out_off[kHeadTmaLayers];
  long long item_start[kHeadTmaLayers + 1];
  long long row_len;
};

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar,
                                            uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%2, %3, %4}], [%5], %6;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

// the producer thread of a CTA: requests the CTA's items (blockIdx.x, + gridDim.x, ...) into the ring, in order
__device__ __forceinline__ void heads_tma_produce(const HeadsTmaPlan& p, unsigned char* ring, uint64_t* full, uint64_t* empty) {
  const uint64_t pol = l2_evict_first_policy();
  const long long total = p.item_start[p.num_layers];
  const int stages = p.stages;
  int st = 0;
  uint32_t ph = 0;
  long long n = 0;
  for (long long item = blockIdx.x; item < total; item += gridDim.x, ++n) {
    if (n >= stages) mbar_wait(&empty[st], ph ^ 1u);
    int k = 0;
    while (item >= p.item_start[k + 1]) ++k;
    const long long local = item - p.item_start[k];
    const int b = (int)(local / p.tiles_hw[k]);
    const int th = (int)(local - (long long)b * p.tiles_hw[k]);
    const uint32_t box_bytes = (uint32_t)p.box_rows[k] * 128u;
    mbar_arrive_expect_tx(&full[st], box_bytes * (uint32_t)p.boxes[k]);
    unsigned char* dst = ring + (size_t)st * p.stage_bytes;
    for (int i = 0; i < p.boxes[k]; ++i)
      tma_load_3d(dst + (size_t)i * box_bytes, &p.map[k], th * kHeadTmaPos, i * p.box_rows[k], b, &full[st], pol);
    if (++st == stages) { st = 0; ph ^= 1u; }
  }
}

// MODE (experiment builds only): 0 = the kernel, 1 = loads without stores, 2 = stores without loads
template <int MODE, bool ALIGN>
__global__ void __launch_bounds__((kHeadTmaWarps + 1) * 32, 1)
heads_tma_kernel(const __grid_constant__ HeadsTmaPlan p, float* __restrict__ out) {
  extern __shared__ unsigned char smem_heads[];
  __shared__ __align__(8) uint64_t full[kHeadTmaMaxStages], empty[kHeadTmaMaxStages];
  unsigned char* ring = smem_heads + ((1024u - (smem_u32(smem_heads) & 1023u)) & 1023u);     // SWIZZLE_128B: 1 KB aligned
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stages = p.stages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kHeadTmaWarps);
    }
    fence_mbar_init();
  }
  __syncthreads();
  const long long total = p.item_start[p.num_layers];
  int st = 0;
  uint32_t ph = 0;
  if (warp == kHeadTmaWarps) {
    if (lane != 0 || MODE == 2) return;
    heads_tma_produce(p, ring, full, empty);
    return;
  }
  for (long long item = blockIdx.x; item < total; item += gridDim.x) {
    int k = 0;
    while (item >= p.item_start[k + 1]) ++k;
    const long long local = item - p.item_start[k];
    const int b = (int)(local / p.tiles_hw[k]);
    const int th = (int)(local - (long long)b * p.tiles_hw[k]);
    const int CH = p.channels[k];
    const int pos0 = th * kHeadTmaPos + warp * 4;                  // this warp's four positions
    const int npos = p.hw[k] - pos0;                               // how many of them exist
    float* dst = out + (size_t)b * p.row_len + p.out_off[k] + (size_t)pos0 * CH;
    const unsigned char* tile = ring + (size_t)st * p.stage_bytes;
    if (MODE != 2) mbar_wait(&full[st], ph);
    float keep = 0.f;
    if (ALIGN && MODE != 1) {
      // every store instruction writes ONE whole 128-byte line of the output: row j of this warp starts sh[j] floats
      // into a line, so lane l stores channel 32 * g - sh[j] + l, which it gets from lane (l - sh[j]) & 31 of
      // channel group g (or g - 1, for the lanes that wrap) with one shuffle
      int sh[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) sh[j] = (int)((reinterpret_cast<uintptr_t>(dst + (size_t)j * CH) >> 2) & 31u);
      float4 prev = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int c0 = 0; c0 < CH + 31; c0 += 128) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int ch = c0 + u * 32 + lane;
          v[u] = ch < CH && MODE != 2 ? *reinterpret_cast<const float4*>(tile + (size_t)ch * 128 + ((warp ^ (ch & 7)) << 4))
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float cur[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
          const float old[4] = {prev.x, prev.y, prev.z, prev.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float x = lane < 32 - sh[j] ? cur[j] : old[j];
            const float y = __shfl_sync(SSDBOX_FULL_MASK, x, (lane - sh[j]) & 31);
            const int ch = c0 + u * 32 - sh[j] + lane;
            if (ch >= 0 && ch < CH && j < npos) __stcs(dst + (size_t)j * CH + ch, y);
          }
          prev = v[u];
        }
      }
    } else {
    for (int c0 = 0; c0 < CH; c0 += 128) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int ch = c0 + u * 32 + lane;
        v[u] = ch < CH && MODE != 2 ? *reinterpret_cast<const float4*>(tile + (size_t)ch * 128 + ((warp ^ (ch & 7)) << 4))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (MODE == 1) {
#pragma unroll
        for (int u = 0; u < 4; ++u) keep += v[u].x + v[u].y + v[u].z + v[u].w;
        continue;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int ch = c0 + u * 32 + lane;
        if (ch < CH) {
          if (npos > 0) __stcs(dst + ch, v[u].x);
          if (npos > 1) __stcs(dst + CH + ch, v[u].y);
          if (npos > 2) __stcs(dst + 2 * CH + ch, v[u].z);
          if (npos > 3) __stcs(dst + 3 * CH + ch, v[u].w);
        }
      }
    }
    }
    if (MODE == 1 && keep == 123.456f) dst[0] = keep;
    __syncwarp();
    if (lane == 0 && MODE != 2) mbar_arrive(&empty[st]);
    if (++st == stages) { st = 0; ph ^= 1u; }
  }
}

typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeFn tensor_map_encoder() {
  static TensorMapEncodeFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<TensorMapEncodeFn>(f);
  }();
  return fn;
}

}  // namespace ssdbox

using namespace ssdbox;

extern "C" int ssdbox_heads_to_rows(const ssdbox_heads_cfg* cfg, float* out, ssdbox_stream_t stream) {
  SSDBOX_REQUIRE(cfg, SSDBOX_EINVAL, "heads: null cfg");
  SSDBOX_REQUIRE(cfg->num_layers >= 0 && cfg->num_layers <= SSDBOX_MAX_HEADS && cfg->B >= 0, SSDBOX_EINVAL,
                 "heads: bad layer count / batch");
  HeadsPlan p{};          // layers that go through the generic tile kernel
  HeadsTmaPlan q{};       // layers that go through the TMA ring
  DevInfo dev;
  int rc = get_dev_info(&dev);
  if (rc) return rc;
  TensorMapEncodeFn encode = cfg->B > 0 ? tensor_map_encoder() : nullptr;
  long long off = 0, tiles = 0, items = 0;
  uint32_t stage_bytes = 0;
  for (int k = 0; k < cfg->num_layers; ++k) {
    SSDBOX_REQUIRE(cfg->channels[k] >= 1 && cfg->hw[k] >= 1, SSDBOX_ESHAPE, "heads: layer %d has an empty shape", k);
    SSDBOX_REQUIRE(cfg->src[k] || cfg->B == 0, SSDBOX_EINVAL, "heads: layer %d has a null pointer", k);
    const int CH = cfg->channels[k], HW = cfg->hw[k];
    bool tma = encode && q.num_layers < kHeadTmaLayers && HW % 4 == 0 && HW >= 2 * kHeadTmaPos && CH <= kHeadTmaMaxRows &&
               CH >= 32 && (reinterpret_cast<uintptr_t>(cfg->src[k]) & 15u) == 0 && (long long)CH * HW * 4 < (1ll << 40);
    if (tma) {
      const int j = q.num_layers;
      const int boxes = (CH + 255) / 256;
      const int box_rows = (((CH + boxes - 1) / boxes) + 7) & ~7;
      cuuint64_t gdim[3] = {(cuuint64_t)HW, (cuuint64_t)CH, (cuuint64_t)cfg->B};
      cuuint64_t gstride[2] = {(cuuint64_t)HW * 4, (cuuint64_t)CH * HW * 4};
      cuuint32_t box[3] = {(cuuint32_t)kHeadTmaPos, (cuuint32_t)box_rows, 1};
      cuuint32_t estride[3] = {1, 1, 1};
      CUresult r = encode(&q.map[j], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(cfg->src[k]), gdim, gstride, box,
                          estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        tma = false;                     // the generic kernel takes every layout
      } else {
        q.channels[j] = CH;
        q.hw[j] = HW;
        q.tiles_hw[j] = (HW + kHeadTmaPos - 1) / kHeadTmaPos;
        q.box_rows[j] = box_rows;
        q.boxes[j] = boxes;
        q.out_off[j] = off;
        q.item_start[j] = items;
        items += (long long)cfg->B * q.tiles_hw[j];
        const uint32_t bytes = (uint32_t)boxes * box_rows * 128u;
        stage_bytes = bytes > stage_bytes ? bytes : stage_bytes;
        ++q.num_layers;
      }
    }
    if (!tma) {
      const int j = p.num_layers++;
      p.channels[j] = CH;
      p.hw[j] = HW;
      p.src[j] = cfg->src[k];
      p.out_off[j] = off;
      p.tiles_ch[j] = (CH + kHeadTileC - 1) / kHeadTileC;
      p.tiles_hw[j] = (HW + kHeadTileS - 1) / kHeadTileS;
      p.tile_start[j] = tiles;
      tiles += (long long)cfg->B * p.tiles_ch[j] * p.tiles_hw[j];
    }
    off += (long long)CH * HW;
  }
  p.B = cfg->B;
  p.tile_start[p.num_layers] = tiles;
  p.row_len = off;
  q.item_start[q.num_layers] = items;
  q.row_len = off;
  if (tiles == 0 && items == 0) return SSDBOX_OK;
  SSDBOX_REQUIRE(out, SSDBOX_EINVAL, "heads: null output");
  SSDBOX_REQUIRE(tiles < (1ll << 40), SSDBOX_ESHAPE, "heads: too many tiles");
  if (items > 0) {
    const int budget = dev.max_smem_optin - 1024 - 256;
    int stages = budget / (int)stage_bytes;
    stages = stages > kHeadTmaMaxStages ? kHeadTmaMaxStages : stages;
    SSDBOX_REQUIRE(stages >= 2, SSDBOX_ESHAPE, "heads: %u-byte stages do not fit", stage_bytes);
    q.stages = stages;
    q.stage_bytes = stage_bytes;
    const size_t smem = (size_t)stages * stage_bytes + 1024;
    const long long grid = items < dev.sm_count ? items : dev.sm_count;
    void (*kern)(HeadsTmaPlan, float*) = heads_tma_kernel<0, true>;
#ifdef SSDBOX_EXPERIMENTS
    const bool al = !getenv("SSDBOX_HEADS_ALIGN") || atoi(getenv("SSDBOX_HEADS_ALIGN")) != 0;
    if (!al) kern = heads_tma_kernel<0, false>;
    if (const char* e = getenv("SSDBOX_HEADS_MODE"))
      kern = atoi(e) == 1 ? heads_tma_kernel<1, true> : (atoi(e) == 2 ? (al ? heads_tma_kernel<2, true> : heads_tma_kernel<2, false>) : kern);
    if (const char* e = getenv("SSDBOX_HEADS_STAGES")) q.stages = atoi(e) >= 2 && atoi(e) < stages ? atoi(e) : stages;
#endif
    SSDBOX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(int)grid, (kHeadTmaWarps + 1) * 32, smem, static_cast<cudaStream_t>(stream)>>>(q, out);
    SSDBOX_LAUNCH_OK("heads_tma_kernel");
  }
  if (tiles > 0) {
    long long grid = tiles < (long long)dev.sm_count * 16 ? tiles : (long long)dev.sm_count * 16;
    heads_to_rows_kernel<<<(int)grid, kHeadThreads, 0, static_cast<cudaStream_t>(stream)>>>(p, out);
    SSDBOX_LAUNCH_OK("heads_to_rows_kernel");
  }
  return SSDBOX_OK;
}
