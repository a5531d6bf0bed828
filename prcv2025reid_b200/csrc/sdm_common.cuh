// sdm_common.cuh -- definitions shared by the two SDM-loss code paths:
//   sdm.cu     fp32 CUDA-core path (any shape / dtype, exact fp32 arithmetic; the small-batch case C2)
//   sdm_tc.cu  bf16 tcgen05 path   (64 <= N,M <= 512, d % 64 == 0; the large-batch case C5)
// Both replace models/sdm_loss.py:13-149 and its autograd backward.
#pragma once
#include "common.cuh"

namespace sdm {

struct Batch {
  reid_sdm_pair p[REID_SDM_MAX_PAIRS];
  int n_pairs;
};

// status bits (include/reid_b200.h): 1 = the reference returned its non-differentiable zero,
// 2 = non-finite feature (:79-81), 4 = non-finite S (:89-91), 8 = no positives (:105-106), 16 = bad result (:145-147)

// element type of the features: DT = REID_DTYPE_F32 (0), REID_DTYPE_BF16 (1) or REID_DTYPE_F16 (2).  The template
// parameter keeps its historic name BF16: non-zero = a 16-bit type whose rounding the reference's normalisation sees.
template <int BF16>
__device__ __forceinline__ float round_dt(float v) {
  if (BF16 == REID_DTYPE_BF16) return __bfloat162float(__float2bfloat16_rn(v));
  if (BF16 == REID_DTYPE_F16) return __half2float(__float2half_rn(v));
  return v;
}
template <int BF16>
__device__ __forceinline__ float ld_elem(const void* base, size_t i) {
  if (BF16 == REID_DTYPE_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[i]);
  if (BF16 == REID_DTYPE_F16) return __half2float(reinterpret_cast<const __half*>(base)[i]);
  return reinterpret_cast<const float*>(base)[i];
}
// normalised element exactly as the reference forms it: fp32: x / den ; bf16 / fp16: round(x / den)
template <int BF16>
__device__ __forceinline__ float norm_elem(const void* base, size_t i, float den) {
  return round_dt<BF16>(__fdiv_rn(ld_elem<BF16>(base, i), den));
}
template <int BF16>
__device__ __forceinline__ void st_out(void* base, size_t i, float v) {
  if (BF16 == REID_DTYPE_BF16) reinterpret_cast<__nv_bfloat16*>(base)[i] = __float2bfloat16_rn(v);
  else if (BF16 == REID_DTYPE_F16) reinterpret_cast<__half*>(base)[i] = __float2half_rn(v);
  else reinterpret_cast<float*>(base)[i] = v;
}

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Positive mask of a pair as the CUDA-core kernels (sdm.cu) read it: the dense y (y[i][j] > 0), or the LABEL FORM
// (y == NULL, include/reid_b200.h; models/model.py:570-605): y[i][j] = row_label[i] == col_label[j] between the rows
// that take part; a row whose valid byte is 0 is absent from the problem (the reference indexes it away before the loss).
struct PosMask {
  const float* y; const int64_t* rl; const int64_t* cl; const uint8_t* rv; const uint8_t* cv; int M;
  __device__ __forceinline__ explicit PosMask(const reid_sdm_pair& P)
      : y(P.y), rl(P.row_label), cl(P.col_label), rv(P.y ? nullptr : P.row_valid), cv(P.y ? nullptr : P.col_valid), M(P.M) {}
  __device__ __forceinline__ bool row_in(int i) const { return !rv || rv[i] != 0; }
  __device__ __forceinline__ bool col_in(int j) const { return !cv || cv[j] != 0; }
  // (only asked for rows / columns that take part)
  __device__ __forceinline__ bool pos(int i, int j) const { return y ? y[(size_t)i * M + j] > 0.f : rl[i] == cl[j]; }
};

// ---- tcgen05 path: layout of the per-pair `saved` buffer (offsets in floats unless noted) ----
// [den_q N][den_g M][lse_r N][lse_c M][cnt_r N][cnt_c M][ce_r N][ce_c M][hdr 128][S N*M][St M*N] then, 128-byte
// aligned, the positive masks of y as bit rows (ybits [N][16] / ybitsT [M][16] 32-bit words) and four 16-bit operand images (Qn / Gn bf16; the transposed QnT / GnT fp16 unless -DREID_SDM_DS_F16=0) in the UMMA K-major 128B-swizzle layout (8-row x 64-element atoms of 1024 bytes,
// [k block][row group]) so that an operand tile is ONE contiguous cp.async.bulk:
//   Qn  [Np rows][d]   Gn  [Mp rows][d]   QnT [d rows][Np]   GnT [d rows][Mp]      (Np, Mp = N, M rounded up to 128)
// hdr: [0] nR  [1] nC  [2] status bits (int)  [3] loss  [4] CTA completion counter (int)
//      [8 .. 8+64)  non-finite-feature flag per 16-row slab of qry / gal (int, 32 each)
//      [72 .. 80)   non-finite-S flag per forward CTA (int)
//      [128 .. 160) validity bits of the qry / gal rows (16 words each)
struct TcLayout {
  int N, M, d, Np, Mp;
  size_t den_q, den_g, lse_r, lse_c, cnt_r, cnt_c, ce_r, ce_c, hdr, S, St;
  size_t ybits, ybitsT;            // byte offsets of the bit masks (64 bytes per row)
  size_t qn, gn, qnt, gnt;         // byte offsets of the images
  size_t total_bytes;
};
constexpr int TC_HDR_FLOATS = 192;    // (words 80..111: optional phase time stamps of REID_SDM_TIMING builds)
constexpr int TC_HDR_VALID_Q = 128;   // 16 words: validity bits of the qry rows (label form; all ones otherwise)
constexpr int TC_HDR_VALID_G = 144;   // 16 words: validity bits of the gal rows
__host__ __device__ inline TcLayout tc_layout(int N, int M, int d) {
  TcLayout L;
  L.N = N; L.M = M; L.d = d; L.Np = round_up(N, 128); L.Mp = round_up(M, 128);
  size_t o = 0;
  L.den_q = o; o += N; L.den_g = o; o += M;
  L.lse_r = o; o += N; L.lse_c = o; o += M;
  L.cnt_r = o; o += N; L.cnt_c = o; o += M;
  L.ce_r = o; o += N; L.ce_c = o; o += M;
  L.hdr = o; o += TC_HDR_FLOATS;
  o = (o + 3) / 4 * 4;                                      // S rows are written with 16-byte stores
  L.S = o; o += (size_t)N * M;
  L.St = o; o += (size_t)N * M;
  size_t b = (o * 4 + 127) / 128 * 128;
  L.ybits = b; b += (size_t)N * 64;
  L.ybitsT = b; b += (size_t)M * 64;
  b = (b + 127) / 128 * 128;
  L.qn = b; b += (size_t)L.Np * d * 2;
  L.gn = b; b += (size_t)L.Mp * d * 2;
  L.qnt = b; b += (size_t)d * L.Np * 2;
  L.gnt = b; b += (size_t)d * L.Mp * 2;
  L.total_bytes = b + 128;                                   // slack: the base is aligned up to 128 bytes
  return L;
}
__host__ __device__ inline float* tc_base(float* saved) {
  return reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(saved) + 127) & ~(uintptr_t)127);
}
// byte offset of element (r, k) inside a K-major swizzled image with `rows_p` rows (multiple of 8)
__host__ __device__ inline size_t sw_offset(int r, int k, int rows_p) {
  const int kb = k >> 6, ch = (k & 63) >> 3, rr = r & 7;
  return (size_t)kb * ((size_t)rows_p * 128) + (size_t)(r >> 3) * 1024 + (size_t)rr * 128 + (size_t)((ch ^ rr) << 4) +
         (size_t)(k & 7) * 2;
}

bool tc_eligible(const reid_sdm_pair* pairs, int n_pairs, int dtype, int d);
int tc_forward(const reid_sdm_pair* pairs, int n_pairs, int d, float tau, float eps, cudaStream_t st);
int tc_backward(const reid_sdm_pair* pairs, int n_pairs, int d, float tau, float eps, cudaStream_t st);

}  // namespace sdm
