// Prior generation and the element-wise box algebra of lib/layers/box_utils.py
// (point_form, jaccard, encode, decode, log_sum_exp) plus the RefineDet ARM filter.
#include "ops.h"
#include "ssdbox_dev.cuh"

namespace ssdbox {

// ------------------------------------------------------------------------------------------------
// PriorBoxSSD.forward, prior_box.py:92-111 + _create_prior :122-143.  One thread per feature-map
// cell; fp64 arithmetic in the reference's operation order, one rounding to fp32, optional clamp.
// ------------------------------------------------------------------------------------------------
struct PriorPlan {
  long long cell_start[SSDBOX_MAX_LAYERS + 1];   // first cell of each layer
  long long prior_start[SSDBOX_MAX_LAYERS + 1];  // first prior of each layer
  int per_cell[SSDBOX_MAX_LAYERS];
};

static int per_cell_count(const ssdbox_prior_cfg* c, int k) {
  int per_min = 1 + (c->has_max ? 1 : 0) + c->num_ratio[k] * (c->flip ? 2 : 1);
  return c->num_min[k] * per_min;
}

static int make_plan(const ssdbox_prior_cfg* c, PriorPlan* pl) {
  if (!c) return fail(SSDBOX_EINVAL, "priorbox: null cfg");
  if (c->num_layers < 0 || c->num_layers > SSDBOX_MAX_LAYERS)
    return fail(SSDBOX_ESHAPE, "priorbox: num_layers %d out of range", c->num_layers);
  if (!(c->image_h > 0) || !(c->image_w > 0)) return fail(SSDBOX_EINVAL, "priorbox: image size must be > 0");
  pl->cell_start[0] = 0;
  pl->prior_start[0] = 0;
  for (int k = 0; k < c->num_layers; ++k) {
    if (c->num_min[k] < 1 || c->num_min[k] > SSDBOX_MAX_MIN_SIZES || c->num_ratio[k] < 0 ||
        c->num_ratio[k] > SSDBOX_MAX_RATIOS || c->feat_h[k] < 0 || c->feat_w[k] < 0 || !(c->step[k] > 0))
      return fail(SSDBOX_ESHAPE, "priorbox: layer %d has an unsupported size list", k);
    pl->per_cell[k] = per_cell_count(c, k);
    long long cells = (long long)c->feat_h[k] * c->feat_w[k];
    pl->cell_start[k + 1] = pl->cell_start[k] + cells;
    pl->prior_start[k + 1] = pl->prior_start[k] + cells * pl->per_cell[k];
  }
  return SSDBOX_OK;
}

__device__ __forceinline__ void put_prior(float* out, long long idx, double cx, double cy, double w, double h,
                                          int clip) {
  float4 v = make_float4(__double2float_rn(cx), __double2float_rn(cy), __double2float_rn(w),
                         __double2float_rn(h));
  if (clip) {
    v.x = fminf(fmaxf(v.x, 0.f), 1.f);
    v.y = fminf(fmaxf(v.y, 0.f), 1.f);
    v.z = fminf(fmaxf(v.z, 0.f), 1.f);
    v.w = fminf(fmaxf(v.w, 0.f), 1.f);
  }
  *reinterpret_cast<float4*>(out + idx * 4) = v;
}

__global__ void priorbox_kernel(ssdbox_prior_cfg c, PriorPlan pl, float* __restrict__ out) {
  long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= pl.cell_start[c.num_layers]) return;
  int k = 0;
  while (cell >= pl.cell_start[k + 1]) ++k;
  long long local = cell - pl.cell_start[k];
  int i = (int)(local / c.feat_w[k]);
  int j = (int)(local - (long long)i * c.feat_w[k]);
  double sx = __ddiv_rn(c.image_w, c.step[k]);       // prior_box.py:99-102
  double sy = __ddiv_rn(c.image_h, c.step[k]);
  double cx = __ddiv_rn(__dadd_rn((double)j, 0.5), sx);
  double cy = __ddiv_rn(__dadd_rn((double)i, 0.5), sy);
  long long o = pl.prior_start[k] + local * pl.per_cell[k];
  for (int m = 0; m < c.num_min[k]; ++m) {
    double ms = c.min_size[k][m];
    double sh = __ddiv_rn(ms, c.image_h);           // s_i
    double sw = __ddiv_rn(ms, c.image_w);           // s_j
    put_prior(out, o++, cx, cy, sw, sh, c.clip);
    if (c.has_max) {                                  // :133-137
      double wp = __dsqrt_rn(__dmul_rn(sw, __ddiv_rn(c.max_size[k], c.image_w)));
      double hp = __dsqrt_rn(__dmul_rn(sh, __ddiv_rn(c.max_size[k], c.image_h)));
      put_prior(out, o++, cx, cy, wp, hp, c.clip);
    }
    for (int r = 0; r < c.num_ratio[k]; ++r) {        // :139-142
      double q = __dsqrt_rn(c.ratio[k][r]);
      put_prior(out, o++, cx, cy, __dmul_rn(sw, q), __ddiv_rn(sh, q), c.clip);
      if (c.flip) put_prior(out, o++, cx, cy, __ddiv_rn(sw, q), __dmul_rn(sh, q), c.clip);
    }
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void point_form_kernel(const float* __restrict__ in, long long n, float* __restrict__ out, int to_center) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float4 v = *reinterpret_cast<const float4*>(in + i * 4);
    float4 r;
    if (to_center) {
      Box b;
      b.x1 = v.x; b.y1 = v.y; b.x2 = v.z; b.y2 = v.w;
      r = center_form(b);
    } else {
      Box b = point_form(v);
      r = make_float4(b.x1, b.y1, b.x2, b.y2);
    }
    *reinterpret_cast<float4*>(out + i * 4) = r;
  }
}

__global__ void jaccard_kernel(const float* __restrict__ a, int G, const float* __restrict__ bx, int P,
                               float* __restrict__ out) {
  int g = blockIdx.y;
  const float* r = a + (size_t)g * 4;
  Box t;
  t.x1 = r[0]; t.y1 = r[1]; t.x2 = r[2]; t.y2 = r[3];
  float ta = box_area(t);
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
    float4 v = *reinterpret_cast<const float4*>(bx + (size_t)p * 4);
    Box q;
    q.x1 = v.x; q.y1 = v.y; q.x2 = v.z; q.y2 = v.w;
    out[(size_t)g * P + p] = iou_jaccard_full(t, ta, q, box_area(q));
  }
}

__global__ void encode_kernel(const float* __restrict__ m, const float* __restrict__ pr, long long n, float var0,
                              float var1, float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float4 v = *reinterpret_cast<const float4*>(m + i * 4);
    Box b;
    b.x1 = v.x; b.y1 = v.y; b.x2 = v.z; b.y2 = v.w;
    *reinterpret_cast<float4*>(out + i * 4) = encode_box(b, *reinterpret_cast<const float4*>(pr + i * 4), var0, var1);
  }
}

__global__ void decode_kernel(const float* __restrict__ loc, const float* __restrict__ pr, long long n,
                              long long prior_rows, float var0, float var1, float* __restrict__ out,
                              float* __restrict__ out_center) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float4 l = *reinterpret_cast<const float4*>(loc + i * 4);
    float4 p = *reinterpret_cast<const float4*>(pr + (i % prior_rows) * 4);
    Box b = decode_box(l, p, var0, var1);
    if (out) *reinterpret_cast<float4*>(out + i * 4) = make_float4(b.x1, b.y1, b.x2, b.y2);
    if (out_center) *reinterpret_cast<float4*>(out_center + i * 4) = center_form(b);
  }
}

// log_sum_exp, box_utils.py:272-273: ONE max over the whole tensor, then log(sum(exp(x - max))) + max
__global__ void global_max_kernel(const float* __restrict__ x, long long n, uint32_t* __restrict__ gmax_ord) {
  float m = -INFINITY;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, x[i]);
  uint32_t o = __reduce_max_sync(SSDBOX_FULL_MASK, f2ord(m));
  if ((threadIdx.x & 31) == 0) atomicMax(gmax_ord, o);
}

__global__ void ord_to_double_kernel(const uint32_t* __restrict__ gmax_ord, double* __restrict__ out) {
  const uint32_t o = *gmax_ord;
  out[0] = o == 0u ? -(double)INFINITY : (double)ord2f(o);      // key 0 = nothing seen
}

__global__ void lse_rows_kernel(const float* __restrict__ x, long long rows, int C, const uint32_t* __restrict__ gmax_ord,
                                float* __restrict__ out) {
  float gm = ord2f(*gmax_ord);
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    const float* row = x + r * C;
    float s = 0.f;
    for (int c = 0; c < C; ++c) s = __fadd_rn(s, expf(__fsub_rn(row[c], gm)));
    out[r] = __fadd_rn(logf(s), gm);
  }
}

__global__ void arm_filter_kernel(const float* __restrict__ arm_conf, long long n, float theta, uint8_t* __restrict__ keep) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float2 v = *reinterpret_cast<const float2*>(arm_conf + i * 2);
    float obj = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(__fsub_rn(v.x, v.y))));
    keep[i] = obj > theta ? 1 : 0;
  }
}

static inline int grid_for(long long n, int threads = 256, int cap = 148 * 8) {
  long long b = (n + threads - 1) / threads;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace ssdbox

using namespace ssdbox;

extern "C" int64_t ssdbox_priorbox_count(const ssdbox_prior_cfg* cfg) {
  PriorPlan pl;
  int rc = make_plan(cfg, &pl);
  if (rc) return rc;
  return pl.prior_start[cfg->num_layers];
}

extern "C" int ssdbox_priorbox(const ssdbox_prior_cfg* cfg, float* out, int64_t out_rows, ssdbox_stream_t stream) {
  PriorPlan pl;
  int rc = make_plan(cfg, &pl);
  if (rc) return rc;
  long long total = pl.prior_start[cfg->num_layers];
  SSDBOX_REQUIRE(out_rows == total, SSDBOX_ESHAPE, "priorbox: out has %lld rows, configuration yields %lld",
                 (long long)out_rows, total);
  if (total == 0) return SSDBOX_OK;
  SSDBOX_REQUIRE(out && aligned16(out), SSDBOX_EALIGN, "priorbox: out must be a 16-byte aligned device pointer");
  long long cells = pl.cell_start[cfg->num_layers];
  priorbox_kernel<<<(unsigned)((cells + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(*cfg, pl, out);
  SSDBOX_LAUNCH_OK("priorbox_kernel");
  return SSDBOX_OK;
}

extern "C" int ssdbox_point_form(const float* boxes, int64_t n, float* out, ssdbox_stream_t stream) {
  SSDBOX_REQUIRE(n >= 0, SSDBOX_EINVAL, "point_form: negative n");
  if (n == 0) return SSDBOX_OK;
  SSDBOX_REQUIRE(boxes && out, SSDBOX_EINVAL, "point_form: null pointer");
  SSDBOX_REQUIRE(aligned16(boxes) && aligned16(out), SSDBOX_EALIGN, "point_form: 16-byte alignment required");
  point_form_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(boxes, n, out, 0);
  SSDBOX_LAUNCH_OK("point_form_kernel");
  return SSDBOX_OK;
}

extern "C" int ssdbox_center_form(const float* boxes, int64_t n, float* out, ssdbox_stream_t stream) {
  SSDBOX_REQUIRE(n >= 0, SSDBOX_EINVAL, "center_form: negative n");
  if (n == 0) return SSDBOX_OK;
  SSDBOX_REQUIRE(boxes && out, SSDBOX_EINVAL, "center_form: null pointer");
  SSDBOX_REQUIRE(aligned16(boxes) && aligned16(out), SSDBOX_EALIGN, "center_form: 16-byte alignment required");
  point_form_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(boxes, n, out, 1);
  SSDBOX_LAUNCH_OK("point_form_kernel");
  return SSDBOX_OK;
}

extern "C" int ssdbox_jaccard(const float* a, int32_t G, const float* b, int32_t P, float* out, ssdbox_stream_t stream) {
  SSDBOX_REQUIRE(G >= 0 && P >= 0, SSDBOX_EINVAL, "jaccard: negative size");
  SSDBOX_REQUIRE(G <= 65535, SSDBOX_ESHAPE, "jaccard: G %d > 65535", G);
  if (G == 0 || P == 0) return SSDBOX_OK;
  SSDBOX_REQUIRE(a && b && out, SSDBOX_EINVAL, "jaccard: null pointer");
  SSDBOX_REQUIRE(aligned16(b), SSDBOX_EALIGN, "jaccard: box_b must be 16-byte aligned");
  dim3 grid(grid_for(P, 256, 148), G);
  jaccard_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, G, b, P, out);
  SSDBOX_LAUNCH_OK("jaccard_kernel");
  return SSDBOX_OK;
}

extern "C" int ssdbox_encode(const float* matched, const float* priors, int64_t n, float var0, float var1, float* out,
                             ssdbox_stream_t stream) {
  SSDBOX_REQUIRE(n >= 0, SSDBOX_EINVAL, "encode: negative n");
  if (n == 0) return SSDBOX_OK;
  SSDBOX_REQUIRE(matched && priors && out, SSDBOX_EINVAL, "encode: null pointer");
  SSDBOX_REQUIRE(aligned16(matched) && aligned16(priors) && aligned16(out), SSDBOX_EALIGN,
                 "encode: 16-byte alignment required");
  encode_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(matched, priors, n, var0, var1, out);
  SSDBOX_LAUNCH_OK("encode_kernel");
  return SSDBOX_OK;
}

extern "C" int ssdbox_decode(const float* loc, const float* priors, int64_t n, int64_t prior_rows, float var0,
                             float var1, float* out, float* out_center, ssdbox_stream_t stream) {
  SSDBOX_REQUIRE(n >= 0 && prior_rows >= 0, SSDBOX_EINVAL, "decode: negative size");
  if (n == 0) return SSDBOX_OK;
  SSDBOX_REQUIRE(loc && priors && (out || out_center) && prior_rows > 0, SSDBOX_EINVAL, "decode: null pointer");
  SSDBOX_REQUIRE(aligned16(loc) && aligned16(priors) && (!out || aligned16(out)) && (!out_center || aligned16(out_center)),
                 SSDBOX_EALIGN, "decode: 16-byte alignment required");
  decode_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(loc, priors, n, prior_rows, var0, var1, out,
                                                                           out_center);
  SSDBOX_LAUNCH_OK("decode_kernel");
  return SSDBOX_OK;
}

extern "C" int ssdbox_log_sum_exp(const float* x, int64_t rows, int32_t C, float* out, void* ws, size_t ws_bytes,
                                  ssdbox_stream_t stream) {
  SSDBOX_REQUIRE(rows >= 0 && C >= 1, SSDBOX_EINVAL, "log_sum_exp: bad shape");
  if (rows == 0) return SSDBOX_OK;
  SSDBOX_REQUIRE(x && out && ws, SSDBOX_EINVAL, "log_sum_exp: null pointer");
  SSDBOX_REQUIRE(ws_bytes >= 256, SSDBOX_EWORKSPACE, "log_sum_exp: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint32_t* g = static_cast<uint32_t*>(ws);
  int rc = launch_init(nullptr, 0, g, 1, nullptr, 0, nullptr, 0, st);
  if (rc) return rc;
  global_max_kernel<<<grid_for(rows * C), 256, 0, st>>>(x, rows * (long long)C, g);
  SSDBOX_LAUNCH_OK("global_max_kernel");
  lse_rows_kernel<<<grid_for(rows), 256, 0, st>>>(x, rows, C, g, out);
  SSDBOX_LAUNCH_OK("lse_rows_kernel");
  return SSDBOX_OK;
}

extern "C" int ssdbox_global_max(const float* x, int64_t n, double* out, void* ws, size_t ws_bytes, ssdbox_stream_t stream) {
  SSDBOX_REQUIRE(n >= 0, SSDBOX_EINVAL, "global_max: negative n");
  SSDBOX_REQUIRE(out && ws && (x || n == 0), SSDBOX_EINVAL, "global_max: null pointer");
  SSDBOX_REQUIRE(ws_bytes >= 256, SSDBOX_EWORKSPACE, "global_max: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint32_t* g = static_cast<uint32_t*>(ws);
  int rc = launch_init(nullptr, 0, g, 1, nullptr, 0, nullptr, 0, st);      // ordered key 0 = below every float
  if (rc) return rc;
  if (n > 0) {
    global_max_kernel<<<grid_for(n), 256, 0, st>>>(x, (long long)n, g);
    SSDBOX_LAUNCH_OK("global_max_kernel");
  }
  ord_to_double_kernel<<<1, 1, 0, st>>>(g, out);
  SSDBOX_LAUNCH_OK("ord_to_double_kernel");
  return SSDBOX_OK;
}

extern "C" int ssdbox_arm_filter(const float* arm_conf, int64_t n, float theta, uint8_t* keep, ssdbox_stream_t stream) {
  SSDBOX_REQUIRE(n >= 0, SSDBOX_EINVAL, "arm_filter: negative n");
  if (n == 0) return SSDBOX_OK;
  SSDBOX_REQUIRE(arm_conf && keep, SSDBOX_EINVAL, "arm_filter: null pointer");
  SSDBOX_REQUIRE((reinterpret_cast<uintptr_t>(arm_conf) & 7u) == 0, SSDBOX_EALIGN, "arm_filter: 8-byte alignment required");
  arm_filter_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(arm_conf, n, theta, keep);
  SSDBOX_LAUNCH_OK("arm_filter_kernel");
  return SSDBOX_OK;
}
