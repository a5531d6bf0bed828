// normalize.cu -- K1 l2norm_rows and K2 mm_fuse_normalize: vectorised, coalesced HBM passes.
//
// Replaces eval_mm_protocol.py:46-48 (`l2n` = F.normalize(x, dim=-1)) and :328-365
// (`extract_query_feat`: per-modality l2n -> weighted sum -> l2n).  One warp per output row;
// a row of d fp32 is read with 128-bit streaming loads (lane l owns elements 4l..4l+3 of every
// 128-element block), reduced with a butterfly, and written as fp32 and/or fp16 (the fp16 copy is
// the tensor-core operand of the similarity GEMM).  Roofline: HBM.  Algorithmic bytes per row:
//   K1: d*4 read + d*4 (f32) + d*2 (f16) written;  K2: k*d*4 read + d*4 + d*2 written.
#include "common.cuh"

namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kMaxVec = 8;  // supports d <= 8*128 = 1024 in registers

template <int NV>
__device__ __forceinline__ void load_row(const float* __restrict__ row, int d, int lane, float4 (&v)[NV]) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane * 4 + i * 128;
    v[i] = (c < d) ? ldg_stream(reinterpret_cast<const float4*>(row + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

template <int NV>
__device__ __forceinline__ float row_sumsq(const float4 (&v)[NV]) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    s = fmaf(v[i].x, v[i].x, s); s = fmaf(v[i].y, v[i].y, s);
    s = fmaf(v[i].z, v[i].z, s); s = fmaf(v[i].w, v[i].w, s);
  }
  return warp_sum(s);
}

template <int NV>
__device__ __forceinline__ void store_row(const float4 (&v)[NV], float* out_f32, __half* out_f16,
                                          int64_t row, int d, int lane) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane * 4 + i * 128;
    if (c < d) {
      if (out_f32) *reinterpret_cast<float4*>(out_f32 + row * d + c) = v[i];
      if (out_f16) {
        __half2 lo = __floats2half2_rn(v[i].x, v[i].y);
        __half2 hi = __floats2half2_rn(v[i].z, v[i].w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&lo);
        pk.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(out_f16 + row * d + c) = pk;
      }
    }
  }
}

// x / max(||x||, eps) as ATen's normalize forms it (denominator, then an IEEE division per element).  The
// division is the Markstein sequence on the correctly rounded reciprocal r = RN(1/denom), computed once per
// row:  q0 = RN(x r),  e = x - denom q0 (exact in the FMA),  q = RN(q0 + e r)  -- correctly rounded like
// div.rn for |x| <= denom (always true here: no overflow, no special cases), at 3 instructions per element
// instead of the ~9 of the generic division; K2 (five normalisations per query) was issue-bound on it.
template <int NV>
__device__ __forceinline__ void normalize_inplace(float4 (&v)[NV], float eps) {
  const float denom = fmaxf(sqrtf(row_sumsq<NV>(v)), eps);
  const float r = __frcp_rn(denom);
  auto div = [&](float x) -> float {
    const float q0 = __fmul_rn(x, r);
    const float e = __fmaf_rn(-denom, q0, x);
    return __fmaf_rn(e, r, q0);
  };
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i].x = div(v[i].x); v[i].y = div(v[i].y);
    v[i].z = div(v[i].z); v[i].w = div(v[i].w);
  }
}

template <int NV>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
l2norm_rows_kernel(const float* __restrict__ x, float* __restrict__ out_f32, __half* __restrict__ out_f16,
                   int64_t rows, int d, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  const int64_t stride = (int64_t)gridDim.x * kWarpsPerCta;
  for (int64_t row = warp0; row < rows; row += stride) {
    float4 v[NV];
    load_row<NV>(x + row * d, d, lane, v);
    normalize_inplace<NV>(v, eps);
    store_row<NV>(v, out_f32, out_f16, row, d, lane);
  }
}

template <int NV>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
mm_fuse_normalize_kernel(const float* __restrict__ feats, const int32_t* __restrict__ mod_id,
                         const float* __restrict__ w, int n_mod, float* __restrict__ out_f32,
                         __half* __restrict__ out_f16, int64_t Q, int k, int d) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  const int64_t stride = (int64_t)gridDim.x * kWarpsPerCta;
  const float eps = 1e-12f;  // F.normalize default (eval_mm_protocol.py:48)
  for (int64_t q = warp0; q < Q; q += stride) {
    float4 acc[NV];
    bool first = true;
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < k; ++j) {
      const int m = mod_id[q * k + j];
      if (m < 0) continue;                        // empty slot (ragged MM-k batches)
      float4 v[NV];
      load_row<NV>(feats + (q * k + j) * (int64_t)d, d, lane, v);
      normalize_inplace<NV>(v, eps);               // :353
      if (k == 1) {                                // fuse_features_if_any returns feats[0] (:209-210)
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] = v[i];
        continue;
      }
      const float wm = (m < n_mod) ? w[m] : 1.0f;  // weight_cfg.get(m, 1.0) (:362)
      // (stack * w[:, None]).sum(0): rounded product, then left-to-right adds (:364-365)
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float px = __fmul_rn(v[i].x, wm), py = __fmul_rn(v[i].y, wm);
        const float pz = __fmul_rn(v[i].z, wm), pw = __fmul_rn(v[i].w, wm);
        if (first) { acc[i] = make_float4(px, py, pz, pw); }
        else {
          acc[i].x = __fadd_rn(acc[i].x, px); acc[i].y = __fadd_rn(acc[i].y, py);
          acc[i].z = __fadd_rn(acc[i].z, pz); acc[i].w = __fadd_rn(acc[i].w, pw);
        }
      }
      first = false;
    }
    normalize_inplace<NV>(acc, eps);               // :359 / :365
    store_row<NV>(acc, out_f32, out_f16, q, d, lane);
  }
}

inline int grid_for_rows(int64_t rows, int waves = 4) {
  int64_t ctas = (rows + kWarpsPerCta - 1) / kWarpsPerCta;
  const int64_t cap = 148LL * 8 * waves;  // `waves` waves of 8 resident CTAs per SM, grid-stride beyond
  if (ctas > cap) ctas = cap;
  if (ctas < 1) ctas = 1;
  return (int)ctas;
}

}  // namespace

extern "C" int reid_l2norm_rows(const float* x, float* out_f32, void* out_f16, int64_t rows, int d,
                                float eps, void* stream) {
  if (!x || (!out_f32 && !out_f16) || rows < 0 || d <= 0 || d % 4 != 0 || d > kMaxVec * 128) return REID_E_INVALID;
  if (rows == 0) return REID_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int nv = (d + 127) / 128;
  const int grid = grid_for_rows(rows);
#define LAUNCH(NV) l2norm_rows_kernel<NV><<<grid, kWarpsPerCta * 32, 0, st>>>(x, out_f32, (__half*)out_f16, rows, d, eps)
  switch (nv) {
    case 1: LAUNCH(1); break; case 2: LAUNCH(2); break; case 3: LAUNCH(3); break; case 4: LAUNCH(4); break;
    case 5: LAUNCH(5); break; case 6: LAUNCH(6); break; case 7: LAUNCH(7); break; default: LAUNCH(8); break;
  }
#undef LAUNCH
  REID_CHECK_LAUNCH();
  return REID_OK;
}

extern "C" int reid_mm_fuse_normalize(const float* feats, const int32_t* mod_id, const float* w, int n_mod,
                                      float* out_f32, void* out_f16, int64_t Q, int k, int d, void* stream) {
  if (!feats || !mod_id || !w || (!out_f32 && !out_f16) || Q < 0 || k <= 0 || d <= 0 || d % 4 != 0 ||
      d > kMaxVec * 128)
    return REID_E_INVALID;
  if (Q == 0) return REID_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int nv = (d + 127) / 128;
  // a query is 8-11 KB of traffic: one query per warp up to 300k queries (an uneven 2-vs-3 grid-stride split of
  // the last wave costs more than the extra CTA launches; measured 62 % -> 67 % of the HBM peak at Q = 100k).
  // Prefetching all k rows of a query, or software-pipelining them, costs occupancy and was slower (48-50 %).
  const int grid = grid_for_rows(Q, 32);
#define LAUNCH(NV) mm_fuse_normalize_kernel<NV><<<grid, kWarpsPerCta * 32, 0, st>>>(feats, mod_id, w, n_mod, out_f32, (__half*)out_f16, Q, k, d)
  switch (nv) {
    case 1: LAUNCH(1); break; case 2: LAUNCH(2); break; case 3: LAUNCH(3); break; case 4: LAUNCH(4); break;
    case 5: LAUNCH(5); break; case 6: LAUNCH(6); break; case 7: LAUNCH(7); break; default: LAUNCH(8); break;
  }
#undef LAUNCH
  REID_CHECK_LAUNCH();
  return REID_OK;
}
