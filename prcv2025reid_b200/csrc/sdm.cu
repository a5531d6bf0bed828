// sdm.cu -- SDM cross-modal alignment loss, forward and backward, batched over modality pairs.
//
// Replaces models/sdm_loss.py:13-149 (`sdm_loss_stable`) and its autograd backward.
// One launch per direction covers every (modality -> vis) pair of a training step: the reference
// issues ~25 ATen ops and >= 10 host syncs per pair; here all guards
// (non-finite features / similarities, rows without positives, negative result) are evaluated on
// the device and reported in `status`, the loss is then the reference's zero.
//
//   S      = q^ g^T / tau_eff, clamp(+-20)                     (:28, :31-32, :86, :94)
//   L      = 0.5 * [ mean_{i in R} (lse_i - mean_{pos} S_ij) + mean_{j in C} (lse_j - mean_{pos} S_ij) ]
//   dL/dS  = 0.5 * [ 1_R (softmax_row - y/cnt_row)/|R| + 1_C (softmax_col - y/cnt_col)/|C| ]
//   dq^ = dS g^ / tau, dg^ = dS^T q^ / tau, then the normalisation Jacobian.
//
// All arithmetic is fp32 on CUDA cores (bf16 inputs are first rounded exactly as the reference's
// bf16 F.normalize rounds them); a grid of CTAs per pair cooperates through grid-wide barriers
// (cooperative launch), a single CTA per pair is used when the pair is one tile (C2: N=M=8).
// HBM roofline per pair, fwd+bwd: 3*(N+M)*d*s + 2*N*M*4 bytes.
#include "sdm_common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace {

constexpr int TB = 256;      // threads per CTA
constexpr int TM = 64;       // tile rows
constexpr int TN = 64;       // tile cols
constexpr int KC = 16;       // k chunk

using SdmBatch = sdm::Batch;
using sdm::ld_elem;
using sdm::norm_elem;
using sdm::st_out;
using sdm::round_dt;
using sdm::PosMask;

struct Saved {
  float *den_q, *den_g, *lse_r, *lse_c, *cnt_r, *cnt_c, *ce_r, *ce_c, *hdr, *S, *dqn, *dgn, *qn, *gn;
};
__host__ __device__ inline size_t saved_floats(int N, int M, int d) {
  return (size_t)4 * N + (size_t)4 * M + 8 + (size_t)N * M + (size_t)2 * (N + M) * d;
}
__device__ inline Saved carve(float* base, int N, int M, int d) {
  Saved s;
  s.den_q = base; s.den_g = s.den_q + N;
  s.lse_r = s.den_g + M; s.lse_c = s.lse_r + N;
  s.cnt_r = s.lse_c + M; s.cnt_c = s.cnt_r + N;
  s.ce_r = s.cnt_c + M; s.ce_c = s.ce_r + N;
  s.hdr = s.ce_c + M;              // [0]=nR [1]=nC [2]=flags(as int bits) [3]=loss
  s.S = s.hdr + 8;
  s.dqn = s.S + (size_t)N * M;
  s.dgn = s.dqn + (size_t)N * d;
  s.qn = s.dgn + (size_t)M * d;          // normalised features exactly as the reference forms them (fp32 copy)
  s.gn = s.qn + (size_t)N * d;
  return s;
}

__device__ __forceinline__ void sync_all(bool multi) {
  if (multi) cg::this_grid().sync(); else __syncthreads();
}

// 64x64 output tile, 256 threads, thread (ty,tx) owns rows ty*4.., cols tx*4..
// la(r,k) / lb(k,c) are tile-local element functors returning 0 outside the matrix.
template <bool A_KCONTIG, bool B_KCONTIG, class LA, class LB>
__device__ __forceinline__ void tile_gemm(int K, LA la, LB lb, float (&acc)[4][4], float (*As)[TM + 4], float (*Bs)[TN + 4]) {
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += KC) {
    if (A_KCONTIG) { const int r = t >> 2, kq = (t & 3) * 4;
#pragma unroll
      for (int u = 0; u < 4; ++u) As[kq + u][r] = (k0 + kq + u < K) ? la(r, k0 + kq + u) : 0.f;
    } else { const int kk = t >> 4, r4 = (t & 15) * 4;
#pragma unroll
      for (int u = 0; u < 4; ++u) As[kk][r4 + u] = (k0 + kk < K) ? la(r4 + u, k0 + kk) : 0.f;
    }
    if (B_KCONTIG) { const int c = t >> 2, kq = (t & 3) * 4;
#pragma unroll
      for (int u = 0; u < 4; ++u) Bs[kq + u][c] = (k0 + kq + u < K) ? lb(k0 + kq + u, c) : 0.f;
    } else { const int kk = t >> 4, c4 = (t & 15) * 4;
#pragma unroll
      for (int u = 0; u < 4; ++u) Bs[kk][c4 + u] = (k0 + kk < K) ? lb(k0 + kk, c4 + u) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
}

// row denominators max(||x||, eps) in the reference's dtype path + non-finite detection
template <int BF16>
__device__ void phase_norms(const void* x, int rows, int d, float eps, float* den, float* xn, int* flags, int cta, int nctas,
                            const uint8_t* valid) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = cta * (TB / 32) + warp; r < rows; r += nctas * (TB / 32)) {
    if (valid && !valid[r]) {              // label form: the row takes no part -- zero image, unit denominator, never inspected
      for (int c = lane; c < d; c += 32) xn[(size_t)r * d + c] = 0.f;
      if (lane == 0) den[r] = 1.f;
      continue;
    }
    float ss = 0.f;
    for (int c = lane; c < d; c += 32) { const float v = ld_elem<BF16>(x, (size_t)r * d + c); ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    float nrm = sqrtf(ss), e = eps;
    if (BF16) { nrm = round_dt<BF16>(nrm); e = round_dt<BF16>(eps); }
    const float dn = fmaxf(nrm, e);
    bool bad = false;
    for (int c = lane; c < d; c += 32) {
      const float v = norm_elem<BF16>(x, (size_t)r * d + c, dn);
      xn[(size_t)r * d + c] = v;
      bad |= !isfinite(v);
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(flags, 2);     // sdm_loss.py:79-81
    if (lane == 0) den[r] = dn;
  }
}

template <int BF16>
__global__ void __launch_bounds__(TB)
sdm_fwd_kernel(SdmBatch batch, int d, float tau_eff, float eps) {
  __shared__ __align__(16) float As[KC][TM + 4];
  __shared__ __align__(16) float Bs[KC][TN + 4];
  __shared__ double red[TB / 32][4];
  const bool multi = gridDim.x > 1;
  const reid_sdm_pair& P = batch.p[blockIdx.y];
  const int N = P.N, M = P.M, cta = blockIdx.x, nctas = gridDim.x;
  Saved sv = carve(P.saved, N, M, d);
  const PosMask pm(P);
  int* flags = reinterpret_cast<int*>(sv.hdr + 2);
  if (cta == 0 && threadIdx.x == 0) *flags = 0;
  sync_all(multi);
  // ---- phase 0: denominators (:31-32) ----
  phase_norms<BF16>(P.qry, N, d, eps, sv.den_q, sv.qn, flags, cta, nctas, pm.rv);
  phase_norms<BF16>(P.gal, M, d, eps, sv.den_g, sv.gn, flags, cta, nctas, pm.cv);
  sync_all(multi);
  // ---- phase 1: S = q^ g^T / tau, clamp (:86, :94) ----
  const int tiles_m = (N + TM - 1) / TM, tiles_n = (M + TN - 1) / TN;
  for (int tile = cta; tile < tiles_m * tiles_n; tile += nctas) {
    const int i0 = (tile / tiles_n) * TM, j0 = (tile % tiles_n) * TN;
    float acc[4][4];
    auto la = [&](int r, int k) { const int i = i0 + r; return i < N ? sv.qn[(size_t)i * d + k] : 0.f; };
    auto lb = [&](int k, int c) { const int j = j0 + c; return j < M ? sv.gn[(size_t)j * d + k] : 0.f; };
    tile_gemm<true, true>(d, la, lb, acc, As, Bs);
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    bool bad = false;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int i = i0 + ty * 4 + a, j = j0 + tx * 4 + b;
        if (i < N && j < M) {
          const float s = __fdiv_rn(acc[a][b], tau_eff);
          bad |= !isfinite(s);
          sv.S[(size_t)i * M + j] = fminf(fmaxf(s, -20.f), 20.f);
        }
      }
    if (bad) atomicOr(flags, 4);                                            // :89-91
  }
  sync_all(multi);
  // ---- phase 2: per-row and per-column log-sum-exp and cross-entropy (:34-57) ----
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // (label form: only the rows / columns that take part enter the sums; an absent row keeps cnt = 0 and is never counted)
    for (int i = cta * (TB / 32) + warp; i < N; i += nctas * (TB / 32)) {
      const bool rin = pm.row_in(i);
      float mx = -INFINITY;
      for (int j = lane; j < M; j += 32) if (rin && pm.col_in(j)) mx = fmaxf(mx, sv.S[(size_t)i * M + j]);
      mx = warp_max(mx);
      float se = 0.f, ps = 0.f, pc = 0.f;
      for (int j = lane; j < M; j += 32) {
        if (!rin || !pm.col_in(j)) continue;
        const float s = sv.S[(size_t)i * M + j];
        se += expf(s - mx);
        if (pm.pos(i, j)) { ps += s; pc += 1.f; }
      }
      se = warp_sum(se); ps = warp_sum(ps); pc = warp_sum(pc);
      if (lane == 0) {
        const float lse = se > 0.f ? mx + logf(se) : 0.f;      // (se >= 1 whenever an element took part)
        sv.lse_r[i] = lse; sv.cnt_r[i] = pc;
        sv.ce_r[i] = pc > 0.f ? (lse - ps / pc) : 0.f;     // -(q * log_p).sum, q uniform over positives
      }
    }
    // columns: warp handles one column; lanes stride over rows (S is L2 resident, N*M*4 bytes)
    for (int j = cta * (TB / 32) + warp; j < M; j += nctas * (TB / 32)) {
      const bool cin = pm.col_in(j);
      float mx = -INFINITY;
      for (int i = lane; i < N; i += 32) if (cin && pm.row_in(i)) mx = fmaxf(mx, sv.S[(size_t)i * M + j]);
      mx = warp_max(mx);
      float se = 0.f, ps = 0.f, pc = 0.f;
      for (int i = lane; i < N; i += 32) {
        if (!cin || !pm.row_in(i)) continue;
        const float s = sv.S[(size_t)i * M + j];
        se += expf(s - mx);
        if (pm.pos(i, j)) { ps += s; pc += 1.f; }
      }
      se = warp_sum(se); ps = warp_sum(ps); pc = warp_sum(pc);
      if (lane == 0) {
        const float lse = se > 0.f ? mx + logf(se) : 0.f;
        sv.lse_c[j] = lse; sv.cnt_c[j] = pc;
        sv.ce_c[j] = pc > 0.f ? (lse - ps / pc) : 0.f;
      }
    }
  }
  sync_all(multi);
  // ---- phase 3: means over valid rows / columns, guards (:60-68, :101-106, :121-123, :142-147) ----
  if (cta == 0) {
    double sr = 0.0, nr = 0.0, sc = 0.0, nc = 0.0;
    for (int i = threadIdx.x; i < N; i += TB)
      if (sv.cnt_r[i] > 0.f && isfinite(sv.ce_r[i])) { sr += sv.ce_r[i]; nr += 1.0; }
    for (int j = threadIdx.x; j < M; j += TB)
      if (sv.cnt_c[j] > 0.f && isfinite(sv.ce_c[j])) { sc += sv.ce_c[j]; nc += 1.0; }
    double v[4] = {sr, nr, sc, nc};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
      if (lane == 0) red[warp][k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double t[4] = {0, 0, 0, 0};
      for (int w = 0; w < TB / 32; ++w)
        for (int k = 0; k < 4; ++k) t[k] += red[w][k];
      int anypos = 0;
      for (int i = 0; i < N && !anypos; ++i) anypos = sv.cnt_r[i] > 0.f;
      int st = *flags;
      if (!anypos) st |= 8;
      const float lr = t[1] > 0 ? (float)(t[0] / t[1]) : 0.f;
      const float lc = t[3] > 0 ? (float)(t[2] / t[3]) : 0.f;
      float loss = 0.5f * (lr + lc);
      if (!(st & (2 | 4 | 8)) && (isnan(loss) || isinf(loss) || loss < 0.f)) st |= 16;
      if (st & (2 | 4 | 8 | 16)) { st |= 1; loss = 0.f; }
      sv.hdr[0] = (float)t[1]; sv.hdr[1] = (float)t[3]; sv.hdr[3] = loss;
      *flags = st;
      *P.loss = loss;
      *P.status = st;
    }
  }
}

template <int BF16>
__global__ void __launch_bounds__(TB)
sdm_bwd_kernel(SdmBatch batch, int d, float tau_eff, float eps) {
  __shared__ __align__(16) float As[KC][TM + 4];
  __shared__ __align__(16) float Bs[KC][TN + 4];
  const bool multi = gridDim.x > 1;
  const reid_sdm_pair& P = batch.p[blockIdx.y];
  const int N = P.N, M = P.M, cta = blockIdx.x, nctas = gridDim.x;
  Saved sv = carve(P.saved, N, M, d);
  const PosMask pm(P);
  const int st = *reinterpret_cast<const int*>(sv.hdr + 2);
  const float nR = sv.hdr[0], nC = sv.hdr[1];
  const float gscale = (st & 1) ? 0.f : (*P.grad_out) * 0.5f / tau_eff;
  const float wr = nR > 0.f ? gscale / nR : 0.f, wc = nC > 0.f ? gscale / nC : 0.f;
  // dL/dS element (chain through clamp: zero where saturated)
  auto dS = [&](int i, int j) -> float {
    if (i >= N || j >= M) return 0.f;
    if (!pm.row_in(i) || !pm.col_in(j)) return 0.f;          // label form: absent rows / columns
    const float s = sv.S[(size_t)i * M + j];
    if (s >= 20.f || s <= -20.f) return 0.f;
    const float pos = pm.pos(i, j) ? 1.f : 0.f;
    float g = 0.f;
    const float cr = sv.cnt_r[i], cc = sv.cnt_c[j];
    if (cr > 0.f && isfinite(sv.ce_r[i])) g += wr * (expf(s - sv.lse_r[i]) - pos / cr);
    if (cc > 0.f && isfinite(sv.ce_c[j])) g += wc * (expf(s - sv.lse_c[j]) - pos / cc);
    return g;
  };
  const int tiles_d = (d + TN - 1) / TN;
  const int tq = ((N + TM - 1) / TM) * tiles_d, tg = ((M + TM - 1) / TM) * tiles_d;
  const bool dead = (st & 1) != 0;   // reference returned its non-differentiable zero: gradients are zero
  for (int tile = cta; tile < tq + tg && !dead; tile += nctas) {
    float acc[4][4];
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    if (tile < tq) {        // dq^[i][c] = sum_j dS[i][j] g^[j][c]
      const int i0 = (tile / tiles_d) * TM, c0 = (tile % tiles_d) * TN;
      auto la = [&](int r, int k) { return dS(i0 + r, k); };
      auto lb = [&](int k, int c) { return (c0 + c < d) ? sv.gn[(size_t)k * d + c0 + c] : 0.f; };
      tile_gemm<true, false>(M, la, lb, acc, As, Bs);
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int i = i0 + ty * 4 + a, c = c0 + tx * 4 + b;
          if (i < N && c < d) sv.dqn[(size_t)i * d + c] = acc[a][b];
        }
    } else {                // dg^[j][c] = sum_i dS[i][j] q^[i][c]
      const int t2 = tile - tq;
      const int j0 = (t2 / tiles_d) * TM, c0 = (t2 % tiles_d) * TN;
      auto la = [&](int r, int k) { return dS(k, j0 + r); };
      auto lb = [&](int k, int c) { return (c0 + c < d) ? sv.qn[(size_t)k * d + c0 + c] : 0.f; };
      tile_gemm<false, false>(N, la, lb, acc, As, Bs);
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int j = j0 + ty * 4 + a, c = c0 + tx * 4 + b;
          if (j < M && c < d) sv.dgn[(size_t)j * d + c] = acc[a][b];
        }
    }
  }
  sync_all(multi);
  // normalisation Jacobian: dx = (dxn - x^ (dxn . x^)) / den   (den = ||x|| unless clamped by eps)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = cta * (TB / 32) + warp; r < N + M; r += nctas * (TB / 32)) {
    const bool isq = r < N;
    const int row = isq ? r : r - N;
    void* out = isq ? P.dqry : P.dgal;
    const float den = isq ? sv.den_q[row] : sv.den_g[row];
    const float* dxn = (isq ? sv.dqn : sv.dgn) + (size_t)row * d;
    float e = eps;
    if (BF16) e = round_dt<BF16>(eps);
    const bool clamped = !(den > e);   // norm <= eps: denominator is the constant eps
    if (dead) {
      for (int c = lane; c < d; c += 32) st_out<BF16>(out, (size_t)row * d + c, 0.f);
      continue;
    }
    float dot = 0.f;
    const float* xnr = (isq ? sv.qn : sv.gn) + (size_t)row * d;
    for (int c = lane; c < d; c += 32) dot = fmaf(dxn[c], xnr[c], dot);
    dot = warp_sum(dot);
    if (clamped) dot = 0.f;
    for (int c = lane; c < d; c += 32) {
      st_out<BF16>(out, (size_t)row * d + c, (dxn[c] - xnr[c] * dot) / den);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Small-batch path (BASELINE C2: P x K = 4 x 2 -> N = M = 8; any N, M <= 32, d % 128 == 0, d <= 512): one CTA per
// pair, everything in shared memory, no grid barrier.  The step is pure latency at this size (a few KB).
// Uses the same `saved` layout as the general path (den, lse, cnt, ce, hdr, S).
// ------------------------------------------------------------------------------------------------
constexpr int SMALL_MAX = 32;

// one row (d <= 512, d % 128 == 0) into registers: lane owns elements k*128 + lane*4 .. +3; all loads issued up front
template <int BF16>
__device__ __forceinline__ void small_load_row(const void* x, size_t row, int d, int lane, float (&v)[16]) {
  const int nchunk = d >> 7;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k < nchunk) {
      const size_t o = row * d + k * 128 + lane * 4;
      if (BF16 == REID_DTYPE_BF16) {
        const uint2 t = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(x) + o);
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.x));
        const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.y));
        v[4 * k] = a.x; v[4 * k + 1] = a.y; v[4 * k + 2] = b.x; v[4 * k + 3] = b.y;
      } else if (BF16 == REID_DTYPE_F16) {
        const uint2 t = *reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(x) + o);
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
        v[4 * k] = a.x; v[4 * k + 1] = a.y; v[4 * k + 2] = b.x; v[4 * k + 3] = b.y;
      } else {
        const float4 t = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + o);
        v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
      }
    } else {
      v[4 * k] = v[4 * k + 1] = v[4 * k + 2] = v[4 * k + 3] = 0.f;
    }
  }
}
// normalised row into shared memory (reference dtype path); returns whether every element is finite
template <int BF16>
__device__ __forceinline__ bool small_store_norm(const float (&v)[16], float dn, float* dst, int d, int lane) {
  const int nchunk = d >> 7;
  bool fin = true;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (k < nchunk) {
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        o[e] = __fdiv_rn(v[4 * k + e], dn);
        o[e] = round_dt<BF16>(o[e]);
        fin = fin && isfinite(o[e]);
      }
      *reinterpret_cast<float4*>(dst + k * 128 + lane * 4) = make_float4(o[0], o[1], o[2], o[3]);
    }
  return fin;
}

__device__ __forceinline__ void small_zero_row(float* dst, int d, int lane) {
  const int nchunk = d >> 7;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (k < nchunk) *reinterpret_cast<float4*>(dst + k * 128 + lane * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
}

template <int BF16>
__global__ void __launch_bounds__(TB)
sdm_small_fwd_kernel(SdmBatch batch, int d, float tau_eff, float eps) {
  extern __shared__ __align__(16) float small_smem[];
  const reid_sdm_pair& P = batch.p[blockIdx.x];
  const int N = P.N, M = P.M;
  Saved sv = carve(P.saved, N, M, d);
  float* xs = small_smem;                              // [N + M][d] normalised rows
  float* Ss = xs + (size_t)(N + M) * d;                // [N][M + 1]
  float* st_r = Ss + N * (M + 1);                      // cnt / ce per row, then per column
  __shared__ int s_flags;
  const PosMask pm(P);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_flags = 0;
  __syncthreads();
  // ---- rows: denominators in the reference's dtype path (:31-32), normalised copies, non-finite check (:79-81)
  for (int r = warp; r < N + M; r += TB / 32) {
    const bool isq = r < N;
    const int row = isq ? r : r - N;
    const void* x = isq ? P.qry : P.gal;
    if (!(isq ? pm.row_in(row) : pm.col_in(row))) {     // label form: the row takes no part (zero image, never inspected)
      small_zero_row(xs + (size_t)r * d, d, lane);
      if (lane == 0) (isq ? sv.den_q : sv.den_g)[row] = 1.f;
      continue;
    }
    float v[16];
    small_load_row<BF16>(x, (size_t)row, d, lane, v);
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) ss = fmaf(v[c], v[c], ss);
    ss = warp_sum(ss);
    float nrm = sqrtf(ss), e = eps;
    if (BF16) { nrm = round_dt<BF16>(nrm); e = round_dt<BF16>(eps); }
    const float dn = fmaxf(nrm, e);
    const bool bad = !small_store_norm<BF16>(v, dn, xs + (size_t)r * d, d, lane);
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(&s_flags, 2);
    if (lane == 0) (isq ? sv.den_q : sv.den_g)[row] = dn;
  }
  __syncthreads();
  // ---- S = q^ g^T / tau, clamp (:86, :94): one warp per element
  for (int idx = warp; idx < N * M; idx += TB / 32) {
    const int i = idx / M, j = idx % M;
    const float dot = warp_dot(xs + (size_t)i * d, xs + (size_t)(N + j) * d, d, lane);
    if (lane == 0) {
      const float s = __fdiv_rn(dot, tau_eff);
      if (!isfinite(s)) atomicOr(&s_flags, 4);                               // :89-91
      const float sc = fminf(fmaxf(s, -20.f), 20.f);
      Ss[i * (M + 1) + j] = sc;
      sv.S[(size_t)i * M + j] = sc;
    }
  }
  __syncthreads();
  // ---- per-row / per-column log-sum-exp and cross-entropy (:34-57): one warp per row or column, one lane per element
  for (int r = warp; r < N + M; r += TB / 32) {
    const bool isrow = r < N;
    const int a = isrow ? r : r - N, len = isrow ? M : N;
    // (label form: only the rows / columns that take part enter the sums; an absent row keeps cnt = 0 and is never counted)
    const bool in = lane < len && (isrow ? (pm.row_in(a) && pm.col_in(lane)) : (pm.col_in(a) && pm.row_in(lane)));
    const float s = in ? (isrow ? Ss[a * (M + 1) + lane] : Ss[lane * (M + 1) + a]) : -INFINITY;
    const bool yv = in && (isrow ? pm.pos(a, lane) : pm.pos(lane, a));
    const float mx = warp_max(s);
    const float se = warp_sum(in ? expf(s - mx) : 0.f);
    const float ps = warp_sum(yv ? s : 0.f), pc = warp_sum(yv ? 1.f : 0.f);
    if (lane == 0) {
      const float lse = se > 0.f ? mx + logf(se) : 0.f;        // (se >= 1 whenever an element took part)
      const float ce = pc > 0.f ? (lse - ps / pc) : 0.f;
      (isrow ? sv.lse_r : sv.lse_c)[a] = lse;
      (isrow ? sv.cnt_r : sv.cnt_c)[a] = pc;
      (isrow ? sv.ce_r : sv.ce_c)[a] = ce;
      st_r[2 * r] = pc; st_r[2 * r + 1] = ce;
    }
  }
  __syncthreads();
  // ---- means over valid rows / columns and the guards (:60-68, :101-106, :121-123, :142-147)
  if (threadIdx.x == 0) {
    double t[4] = {0, 0, 0, 0};
    int anypos = 0;
    for (int r = 0; r < N; ++r) {
      const float pc = st_r[2 * r], ce = st_r[2 * r + 1];
      if (pc > 0.f) { anypos = 1; if (isfinite(ce)) { t[0] += ce; t[1] += 1.0; } }
    }
    for (int r = N; r < N + M; ++r) {
      const float pc = st_r[2 * r], ce = st_r[2 * r + 1];
      if (pc > 0.f && isfinite(ce)) { t[2] += ce; t[3] += 1.0; }
    }
    int st = s_flags;
    if (!anypos) st |= 8;
    const float lr = t[1] > 0 ? (float)(t[0] / t[1]) : 0.f;
    const float lc = t[3] > 0 ? (float)(t[2] / t[3]) : 0.f;
    float loss = 0.5f * (lr + lc);
    if (!(st & (2 | 4 | 8)) && (isnan(loss) || isinf(loss) || loss < 0.f)) st |= 16;
    if (st & (2 | 4 | 8 | 16)) { st |= 1; loss = 0.f; }
    sv.hdr[0] = (float)t[1]; sv.hdr[1] = (float)t[3]; sv.hdr[3] = loss;
    *reinterpret_cast<int*>(sv.hdr + 2) = st;
    *P.loss = loss;
    *P.status = st;
  }
}

template <int BF16>
__global__ void __launch_bounds__(TB)
sdm_small_bwd_kernel(SdmBatch batch, int d, float tau_eff, float eps) {
  extern __shared__ __align__(16) float small_smem[];
  const reid_sdm_pair& P = batch.p[blockIdx.x];
  const int N = P.N, M = P.M;
  Saved sv = carve(P.saved, N, M, d);
  float* xs = small_smem;                              // [N + M][d] normalised rows
  float* dSs = xs + (size_t)(N + M) * d;               // [N][M + 1]
  const PosMask pm(P);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int st = *reinterpret_cast<const int*>(sv.hdr + 2);
  if (st & 1) {                                        // the reference returned its non-differentiable zero
    for (int idx = threadIdx.x; idx < (N + M) * d; idx += TB) {
      const bool isq = idx < N * d;
      st_out<BF16>(isq ? P.dqry : P.dgal, isq ? idx : idx - N * d, 0.f);
    }
    return;
  }
  const float nR = sv.hdr[0], nC = sv.hdr[1];
  const float gscale = (*P.grad_out) * 0.5f / tau_eff;
  const float wr = nR > 0.f ? gscale / nR : 0.f, wc = nC > 0.f ? gscale / nC : 0.f;
  for (int idx = threadIdx.x; idx < N * M; idx += TB) {
    const int i = idx / M, j = idx % M;
    const float s = sv.S[idx];
    float g = 0.f;
    if (s < 20.f && s > -20.f && pm.row_in(i) && pm.col_in(j)) {
      const float pos = pm.pos(i, j) ? 1.f : 0.f;
      const float cr = sv.cnt_r[i], cc = sv.cnt_c[j];
      if (cr > 0.f && isfinite(sv.ce_r[i])) g += wr * (expf(s - sv.lse_r[i]) - pos / cr);
      if (cc > 0.f && isfinite(sv.ce_c[j])) g += wc * (expf(s - sv.lse_c[j]) - pos / cc);
    }
    dSs[i * (M + 1) + j] = g;
  }
  for (int r = warp; r < N + M; r += TB / 32) {
    const bool isq = r < N;
    const int row = isq ? r : r - N;
    const void* x = isq ? P.qry : P.gal;
    const float dn = (isq ? sv.den_q : sv.den_g)[row];
    if (!(isq ? pm.row_in(row) : pm.col_in(row))) { small_zero_row(xs + (size_t)r * d, d, lane); continue; }
    float v[16];
    small_load_row<BF16>(x, (size_t)row, d, lane, v);
    small_store_norm<BF16>(v, dn, xs + (size_t)r * d, d, lane);
  }
  __syncthreads();
  // one warp per output row: dx^ = sum_k dS * (other side's x^), then the normalisation Jacobian
  const int nchunk = d >> 7;                           // 128 columns per chunk, 4 per lane (d <= 512)
  float e = eps;
  if (BF16) e = round_dt<BF16>(eps);
  for (int r = warp; r < N + M; r += TB / 32) {
    const bool isq = r < N;
    const int row = isq ? r : r - N, len = isq ? M : N;
    const float* other = xs + (size_t)(isq ? N : 0) * d;
    float4 acc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k2 = 0; k2 < len; ++k2) {
      const float w = isq ? dSs[row * (M + 1) + k2] : dSs[k2 * (M + 1) + row];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k < nchunk) {
          const float4 v = *reinterpret_cast<const float4*>(other + (size_t)k2 * d + k * 128 + lane * 4);
          acc[k].x = fmaf(w, v.x, acc[k].x); acc[k].y = fmaf(w, v.y, acc[k].y);
          acc[k].z = fmaf(w, v.z, acc[k].z); acc[k].w = fmaf(w, v.w, acc[k].w);
        }
    }
    const float den = (isq ? sv.den_q : sv.den_g)[row];
    float dot = 0.f;
    float4 xh[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < nchunk) {
        xh[k] = *reinterpret_cast<const float4*>(xs + (size_t)r * d + k * 128 + lane * 4);
        dot = fmaf(acc[k].x, xh[k].x, dot); dot = fmaf(acc[k].y, xh[k].y, dot);
        dot = fmaf(acc[k].z, xh[k].z, dot); dot = fmaf(acc[k].w, xh[k].w, dot);
      }
    dot = warp_sum(dot);
    if (!(den > e)) dot = 0.f;                         // norm <= eps: the denominator is the constant eps
    void* out = isq ? P.dqry : P.dgal;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < nchunk) {
        const size_t o = (size_t)row * d + k * 128 + lane * 4;
        st_out<BF16>(out, o, (acc[k].x - xh[k].x * dot) / den);
        st_out<BF16>(out, o + 1, (acc[k].y - xh[k].y * dot) / den);
        st_out<BF16>(out, o + 2, (acc[k].z - xh[k].z * dot) / den);
        st_out<BF16>(out, o + 3, (acc[k].w - xh[k].w * dot) / den);
      }
  }
}

// Forward AND backward of a small pair in ONE launch (reid_sdm_step): the normalised rows, S, y and the row / column
// statistics stay in shared memory between the two halves, so the step reads every input once and takes one launch
// instead of two dependent ones (the step is pure latency at C2).  The arithmetic is the two kernels' above, operation
// for operation (tests compare the results bit for bit); `saved` still receives the statistics a later
// reid_sdm_bwd call would need.  grad_out[p] (the weight of loss p in the objective) and y are fetched up front, so
// their latency hides behind the feature rows.  16 warps: at C2 (N = M = 8) every phase is ONE pass -- a row per warp,
// four independent dots per warp in flight, a row / column statistic per warp, an output row per warp.
constexpr int STB = 512;
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// LBL: some pair of the batch is in the label form (y == NULL).  The dense instantiation carries none of the validity
// tests: at C2 the kernel is a pure latency chain and its time is what the benchmark reports.
template <int BF16, bool D512, bool LBL>              // D512: the reference's feature width as a compile-time constant
__global__ void __launch_bounds__(STB, 1)
sdm_small_step_kernel(SdmBatch batch, int d_arg, float tau_eff, float eps) {
  extern __shared__ __align__(16) float small_smem[];
  constexpr int NW = STB / 32;
  const int d = D512 ? 512 : d_arg;                    // (halves the code of this run-once kernel: no chunk predicates)
  const reid_sdm_pair& P = batch.p[blockIdx.x];
  const int N = P.N, M = P.M;
  Saved sv = carve(P.saved, N, M, d);
  float* xs = small_smem;                              // [N + M][d] normalised rows
  float* Ss = xs + (size_t)(N + M) * d;                // [N][M + 1]  S, then dL/dS
  float* st_r = Ss + N * (M + 1);                      // [N + M][4]  cnt, ce, lse, den per row, then per column
  float* Ys = st_r + 4 * (N + M);                      // [N][M]      y
  __shared__ int s_flags, s_status;
  __shared__ float s_nR, s_nC;
  __shared__ unsigned char s_in[2 * SMALL_MAX];         // label form: does row r of [qry; gal] take part (all ones otherwise)
  const PosMask pm(P);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float gout = *P.grad_out;
  float yreg[2];                                        // N * M <= 1024 = 2 per thread
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int i = threadIdx.x + u * STB;
    yreg[u] = 0.f;
    if (i < N * M) {
      if (!LBL || pm.y) yreg[u] = P.y[i];
      else yreg[u] = (pm.row_in(i / M) && pm.col_in(i % M) && pm.rl[i / M] == pm.cl[i % M]) ? 1.f : 0.f;
    }
  }
  if (threadIdx.x == 0) s_flags = 0;
  if (LBL && threadIdx.x < N + M) s_in[threadIdx.x] = threadIdx.x < N ? pm.row_in(threadIdx.x) : pm.col_in(threadIdx.x - N);
  __syncthreads();
  for (int r = warp; r < N + M; r += NW) {
    const bool isq = r < N;
    const int row = isq ? r : r - N;
    const void* x = isq ? P.qry : P.gal;
    if (LBL && !s_in[r]) {                              // label form: the row takes no part (zero image, never inspected)
      small_zero_row(xs + (size_t)r * d, d, lane);
      if (lane == 0) { (isq ? sv.den_q : sv.den_g)[row] = 1.f; st_r[4 * r + 3] = 1.f; }
      continue;
    }
    float v[16];
    small_load_row<BF16>(x, (size_t)row, d, lane, v);
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) ss = fmaf(v[c], v[c], ss);
    ss = warp_sum(ss);
    float nrm = sqrtf(ss), e = eps;
    if (BF16) { nrm = round_dt<BF16>(nrm); e = round_dt<BF16>(eps); }
    const float dn = fmaxf(nrm, e);
    const bool bad = !small_store_norm<BF16>(v, dn, xs + (size_t)r * d, d, lane);
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(&s_flags, 2);
    if (lane == 0) { (isq ? sv.den_q : sv.den_g)[row] = dn; st_r[4 * r + 3] = dn; }
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) { const int i = threadIdx.x + u * STB; if (i < N * M) Ys[i] = yreg[u]; }
  __syncthreads();
  // S: four independent dots per warp and pass (indices past the end are clamped and discarded)
  const int nm = N * M;
  for (int idx0 = warp; idx0 < nm; idx0 += 4 * NW) {
    float dots[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = min(idx0 + u * NW, nm - 1);
      dots[u] = warp_dot(xs + (size_t)(idx / M) * d, xs + (size_t)(N + idx % M) * d, d, lane);
    }
    if (lane < 4 && idx0 + lane * NW < nm) {
      const int idx = idx0 + lane * NW;
      const float dot = lane == 0 ? dots[0] : lane == 1 ? dots[1] : lane == 2 ? dots[2] : dots[3];
      const float sc0 = __fdiv_rn(dot, tau_eff);
      if (!isfinite(sc0)) atomicOr(&s_flags, 4);
      const float sc = fminf(fmaxf(sc0, -20.f), 20.f);
      Ss[(idx / M) * (M + 1) + idx % M] = sc;
      sv.S[idx] = sc;
    }
  }
  __syncthreads();
  for (int r = warp; r < N + M; r += NW) {
    const bool isrow = r < N;
    const int a = isrow ? r : r - N, len = isrow ? M : N;
    const bool in = lane < len && (!LBL || (s_in[r] && s_in[isrow ? N + lane : lane]));
    const float sv_ = in ? (isrow ? Ss[a * (M + 1) + lane] : Ss[lane * (M + 1) + a]) : -INFINITY;
    const float yv = in ? (isrow ? Ys[a * M + lane] : Ys[lane * M + a]) : 0.f;
    const float mx = warp_max(sv_);
    const float se = warp_sum(in ? expf(sv_ - mx) : 0.f);
    const float ps = warp_sum(yv > 0.f ? sv_ : 0.f), pc = warp_sum(yv > 0.f ? 1.f : 0.f);
    if (lane == 0) {
      const float lse = (!LBL || se > 0.f) ? mx + logf(se) : 0.f;
      const float ce = pc > 0.f ? (lse - ps / pc) : 0.f;
      (isrow ? sv.lse_r : sv.lse_c)[a] = lse;
      (isrow ? sv.cnt_r : sv.cnt_c)[a] = pc;
      (isrow ? sv.ce_r : sv.ce_c)[a] = ce;
      st_r[4 * r] = pc; st_r[4 * r + 1] = ce; st_r[4 * r + 2] = lse;
    }
  }
  __syncthreads();
  // means over valid rows / columns and the guards: one warp, a row and a column per lane.  The float64 sums of <= 32
  // fp32 terms are exact, so their order does not matter (same value as the serial loop of sdm_small_fwd_kernel).
  if (warp == 0) {
    double t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    int anypos = 0;
    if (lane < N) {
      const float pc = st_r[4 * lane], ce = st_r[4 * lane + 1];
      if (pc > 0.f) { anypos = 1; if (isfinite(ce)) { t0 = ce; t1 = 1.0; } }
    }
    if (lane < M) {
      const float pc = st_r[4 * (N + lane)], ce = st_r[4 * (N + lane) + 1];
      if (pc > 0.f && isfinite(ce)) { t2 = ce; t3 = 1.0; }
    }
    t0 = warp_sum_f64(t0); t1 = warp_sum_f64(t1); t2 = warp_sum_f64(t2); t3 = warp_sum_f64(t3);
    anypos = __any_sync(0xffffffffu, anypos);
    if (lane == 0) {
      int st = s_flags;
      if (!anypos) st |= 8;
      const float lr = t1 > 0 ? (float)(t0 / t1) : 0.f;
      const float lc = t3 > 0 ? (float)(t2 / t3) : 0.f;
      float loss = 0.5f * (lr + lc);
      if (!(st & (2 | 4 | 8)) && (isnan(loss) || isinf(loss) || loss < 0.f)) st |= 16;
      if (st & (2 | 4 | 8 | 16)) { st |= 1; loss = 0.f; }
      sv.hdr[0] = (float)t1; sv.hdr[1] = (float)t3; sv.hdr[3] = loss;
      *reinterpret_cast<int*>(sv.hdr + 2) = st;
      *P.loss = loss;
      *P.status = st;
      s_status = st; s_nR = (float)t1; s_nC = (float)t3;
    }
  }
  __syncthreads();
  // ---------------------------------------------------------------- backward half (sdm_small_bwd_kernel)
  if (s_status & 1) {
    for (int idx = threadIdx.x; idx < (N + M) * d; idx += STB) {
      const bool isq = idx < N * d;
      st_out<BF16>(isq ? P.dqry : P.dgal, isq ? idx : idx - N * d, 0.f);
    }
    return;
  }
  const float nR = s_nR, nC = s_nC;
  const float gscale = gout * 0.5f / tau_eff;
  const float wr = nR > 0.f ? gscale / nR : 0.f, wc = nC > 0.f ? gscale / nC : 0.f;
  for (int idx = threadIdx.x; idx < nm; idx += STB) {
    const int i = idx / M, j = idx % M;
    const float s = Ss[i * (M + 1) + j];
    float g = 0.f;
    if (s < 20.f && s > -20.f && (!LBL || (s_in[i] && s_in[N + j]))) {
      const float pos = Ys[idx] > 0.f ? 1.f : 0.f;
      const float cr = st_r[4 * i], cc = st_r[4 * (N + j)];
      if (cr > 0.f && isfinite(st_r[4 * i + 1])) g += wr * (expf(s - st_r[4 * i + 2]) - pos / cr);
      if (cc > 0.f && isfinite(st_r[4 * (N + j) + 1])) g += wc * (expf(s - st_r[4 * (N + j) + 2]) - pos / cc);
    }
    Ss[i * (M + 1) + j] = g;                              // (each element is read and rewritten by one thread)
  }
  __syncthreads();
  const int nchunk = d >> 7;
  float e = eps;
  if (BF16) e = round_dt<BF16>(eps);
  for (int r = warp; r < N + M; r += NW) {
    const bool isq = r < N;
    const int row = isq ? r : r - N, len = isq ? M : N;
    const float* other = xs + (size_t)(isq ? N : 0) * d;
    float4 acc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k2 = 0; k2 < len; ++k2) {
      const float w = isq ? Ss[row * (M + 1) + k2] : Ss[k2 * (M + 1) + row];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k < nchunk) {
          const float4 v = *reinterpret_cast<const float4*>(other + (size_t)k2 * d + k * 128 + lane * 4);
          acc[k].x = fmaf(w, v.x, acc[k].x); acc[k].y = fmaf(w, v.y, acc[k].y);
          acc[k].z = fmaf(w, v.z, acc[k].z); acc[k].w = fmaf(w, v.w, acc[k].w);
        }
    }
    const float den = st_r[4 * r + 3];
    float dot = 0.f;
    float4 xh[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < nchunk) {
        xh[k] = *reinterpret_cast<const float4*>(xs + (size_t)r * d + k * 128 + lane * 4);
        dot = fmaf(acc[k].x, xh[k].x, dot); dot = fmaf(acc[k].y, xh[k].y, dot);
        dot = fmaf(acc[k].z, xh[k].z, dot); dot = fmaf(acc[k].w, xh[k].w, dot);
      }
    dot = warp_sum(dot);
    if (!(den > e)) dot = 0.f;
    void* out = isq ? P.dqry : P.dgal;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < nchunk) {
        const size_t o = (size_t)row * d + k * 128 + lane * 4;
        st_out<BF16>(out, o, (acc[k].x - xh[k].x * dot) / den);
        st_out<BF16>(out, o + 1, (acc[k].y - xh[k].y * dot) / den);
        st_out<BF16>(out, o + 2, (acc[k].z - xh[k].z * dot) / den);
        st_out<BF16>(out, o + 3, (acc[k].w - xh[k].w * dot) / den);
      }
  }
}

bool small_eligible(const reid_sdm_pair* pairs, int n_pairs, int d) {
  if (!pairs || n_pairs <= 0 || n_pairs > REID_SDM_MAX_PAIRS || d % 128 != 0 || d > 512) return false;
  for (int i = 0; i < n_pairs; ++i) {
    if (pairs[i].N <= 0 || pairs[i].M <= 0 || pairs[i].N > SMALL_MAX || pairs[i].M > SMALL_MAX) return false;
    if ((reinterpret_cast<uintptr_t>(pairs[i].qry) | reinterpret_cast<uintptr_t>(pairs[i].gal)) & 15u) return false;   // vector row loads
  }
  return true;
}

template <class K>
int launch_small(K kernel, const reid_sdm_pair* pairs, int n_pairs, int d, float tau, float eps, bool bwd, cudaStream_t st,
                 int threads = TB) {
  SdmBatch b;
  b.n_pairs = n_pairs;
  size_t smem = 0;
  for (int i = 0; i < n_pairs; ++i) {
    const reid_sdm_pair& p = pairs[i];
    if (!p.y && !(p.row_label && p.col_label)) return REID_E_INVALID;          // dense y, or the label form
    if (!p.qry || !p.gal || !p.loss || !p.status || !p.saved) return REID_E_INVALID;
    if (bwd && (!p.grad_out || !p.dqry || !p.dgal)) return REID_E_INVALID;
    b.p[i] = p;
    const size_t need = ((size_t)(p.N + p.M) * d + (size_t)p.N * (p.M + 1) + 4 * (size_t)(p.N + p.M) + (size_t)p.N * p.M) * sizeof(float);
    if (need > smem) smem = need;
  }
  const float tau_eff = fmaxf(0.15f, fminf(0.5f, tau));                    // sdm_loss.py:28
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return REID_E_CUDA;
  kernel<<<n_pairs, threads, smem, st>>>(b, d, tau_eff, eps);
  REID_CHECK_LAUNCH();
  return REID_OK;
}

template <class K>
int launch_sdm(K kernel, const reid_sdm_pair* pairs, int n_pairs, int d, float tau, float eps, bool bwd, cudaStream_t st) {
  if (!pairs || n_pairs <= 0 || n_pairs > REID_SDM_MAX_PAIRS || d <= 0) return REID_E_INVALID;
  SdmBatch b;
  b.n_pairs = n_pairs;
  int max_tiles = 1;
  for (int i = 0; i < n_pairs; ++i) {
    const reid_sdm_pair& p = pairs[i];
    if (!p.y && !(p.row_label && p.col_label)) return REID_E_INVALID;          // dense y, or the label form
    if (!p.qry || !p.gal || !p.loss || !p.status || !p.saved || p.N <= 0 || p.M <= 0) return REID_E_INVALID;
    if (bwd && (!p.grad_out || !p.dqry || !p.dgal)) return REID_E_INVALID;
    b.p[i] = p;
    const int tm = (p.N + TM - 1) / TM, tn = (p.M + TN - 1) / TN, td = (d + TN - 1) / TN;
    const int tiles = bwd ? (tm + tn) * td : tm * tn;
    const int rows = (p.N + p.M + 7) / 8;
    const int want = bwd ? tiles : (tiles > 1 ? (tiles > rows ? tiles : rows) : 1);
    if (want > max_tiles) max_tiles = want;
  }
  const float tau_eff = fmaxf(0.15f, fminf(0.5f, tau));                    // sdm_loss.py:28
  int per_sm = 0, sms = 0, dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return REID_E_CUDA;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return REID_E_CUDA;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TB, 0) != cudaSuccess) return REID_E_CUDA;
  int ctas = max_tiles;
  const int cap = (per_sm * sms) / n_pairs;
  if (ctas > cap) ctas = cap;
  if (ctas < 1) ctas = 1;
  dim3 grid(ctas, n_pairs);
  if (ctas == 1) {
    kernel<<<grid, TB, 0, st>>>(b, d, tau_eff, eps);
  } else {
    void* args[] = {&b, &d, (void*)&tau_eff, &eps};
    if (cudaLaunchCooperativeKernel((const void*)kernel, grid, dim3(TB), args, 0, st) != cudaSuccess) return REID_E_CUDA;
  }
  REID_CHECK_LAUNCH();
  return REID_OK;
}

}  // namespace

// large enough for either code path (the tcgen05 path keeps bf16 operand images instead of fp32 copies)
extern "C" size_t reid_sdm_saved_floats(int N, int M, int d) {
  const size_t a = saved_floats(N, M, d);
  const size_t b = (sdm::tc_layout(N, M, d).total_bytes + 3) / 4;
  return a > b ? a : b;
}

extern "C" int reid_sdm_uses_tensor_cores(const reid_sdm_pair* pairs, int n_pairs, int dtype, int d) {
  return sdm::tc_eligible(pairs, n_pairs, dtype, d) ? 1 : 0;
}

extern "C" int reid_sdm_fwd(const reid_sdm_pair* pairs, int n_pairs, int dtype, int d, float tau, float eps, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (small_eligible(pairs, n_pairs, d)) {
    if (dtype == REID_DTYPE_F32) return launch_small(sdm_small_fwd_kernel<REID_DTYPE_F32>, pairs, n_pairs, d, tau, eps, false, st);
    if (dtype == REID_DTYPE_BF16) return launch_small(sdm_small_fwd_kernel<REID_DTYPE_BF16>, pairs, n_pairs, d, tau, eps, false, st);
    if (dtype == REID_DTYPE_F16) return launch_small(sdm_small_fwd_kernel<REID_DTYPE_F16>, pairs, n_pairs, d, tau, eps, false, st);
  }
  if (dtype == REID_DTYPE_F32) return launch_sdm(sdm_fwd_kernel<REID_DTYPE_F32>, pairs, n_pairs, d, tau, eps, false, st);
  if (sdm::tc_eligible(pairs, n_pairs, dtype, d)) return sdm::tc_forward(pairs, n_pairs, d, tau, eps, st);
  if (dtype == REID_DTYPE_BF16) return launch_sdm(sdm_fwd_kernel<REID_DTYPE_BF16>, pairs, n_pairs, d, tau, eps, false, st);
  if (dtype == REID_DTYPE_F16) return launch_sdm(sdm_fwd_kernel<REID_DTYPE_F16>, pairs, n_pairs, d, tau, eps, false, st);
  return REID_E_UNSUPPORTED;
}

extern "C" int reid_sdm_bwd(const reid_sdm_pair* pairs, int n_pairs, int dtype, int d, float tau, float eps, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (small_eligible(pairs, n_pairs, d)) {
    if (dtype == REID_DTYPE_F32) return launch_small(sdm_small_bwd_kernel<REID_DTYPE_F32>, pairs, n_pairs, d, tau, eps, true, st);
    if (dtype == REID_DTYPE_BF16) return launch_small(sdm_small_bwd_kernel<REID_DTYPE_BF16>, pairs, n_pairs, d, tau, eps, true, st);
    if (dtype == REID_DTYPE_F16) return launch_small(sdm_small_bwd_kernel<REID_DTYPE_F16>, pairs, n_pairs, d, tau, eps, true, st);
  }
  if (dtype == REID_DTYPE_F32) return launch_sdm(sdm_bwd_kernel<REID_DTYPE_F32>, pairs, n_pairs, d, tau, eps, true, st);
  if (sdm::tc_eligible(pairs, n_pairs, dtype, d)) return sdm::tc_backward(pairs, n_pairs, d, tau, eps, st);
  if (dtype == REID_DTYPE_BF16) return launch_sdm(sdm_bwd_kernel<REID_DTYPE_BF16>, pairs, n_pairs, d, tau, eps, true, st);
  if (dtype == REID_DTYPE_F16) return launch_sdm(sdm_bwd_kernel<REID_DTYPE_F16>, pairs, n_pairs, d, tau, eps, true, st);
  return REID_E_UNSUPPORTED;
}

template <int DT>
static int launch_small_step(const reid_sdm_pair* pairs, int n_pairs, int d, float tau, float eps, cudaStream_t st) {
  bool lbl = false;
  for (int i = 0; i < n_pairs; ++i) lbl = lbl || !pairs[i].y;
  if (d == 512)
    return lbl ? launch_small(sdm_small_step_kernel<DT, true, true>, pairs, n_pairs, d, tau, eps, true, st, STB)
               : launch_small(sdm_small_step_kernel<DT, true, false>, pairs, n_pairs, d, tau, eps, true, st, STB);
  return lbl ? launch_small(sdm_small_step_kernel<DT, false, true>, pairs, n_pairs, d, tau, eps, true, st, STB)
             : launch_small(sdm_small_step_kernel<DT, false, false>, pairs, n_pairs, d, tau, eps, true, st, STB);
}

// Forward + backward of a step in as few launches as the path allows: one for small pairs, otherwise the two calls above.
// Every pair carries its backward slots (grad_out = the weight of loss p in the objective, dqry, dgal).
extern "C" int reid_sdm_step(const reid_sdm_pair* pairs, int n_pairs, int dtype, int d, float tau, float eps, void* stream) {
  if (small_eligible(pairs, n_pairs, d)) {
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == REID_DTYPE_F32) return launch_small_step<REID_DTYPE_F32>(pairs, n_pairs, d, tau, eps, st);
    if (dtype == REID_DTYPE_BF16) return launch_small_step<REID_DTYPE_BF16>(pairs, n_pairs, d, tau, eps, st);
    if (dtype == REID_DTYPE_F16) return launch_small_step<REID_DTYPE_F16>(pairs, n_pairs, d, tau, eps, st);
  }
  const int rc = reid_sdm_fwd(pairs, n_pairs, dtype, d, tau, eps, stream);
  return rc != REID_OK ? rc : reid_sdm_bwd(pairs, n_pairs, dtype, d, tau, eps, stream);
}

// kernel launches reid_sdm_step issues for this batch (1 small / 3 tcgen05: pack + forward + backward / 2 general)
extern "C" int reid_sdm_step_launches(const reid_sdm_pair* pairs, int n_pairs, int dtype, int d) {
  if (small_eligible(pairs, n_pairs, d) && (dtype == REID_DTYPE_F32 || dtype == REID_DTYPE_BF16 || dtype == REID_DTYPE_F16)) return 1;
  return (dtype == REID_DTYPE_BF16 && sdm::tc_eligible(pairs, n_pairs, dtype, d)) ? 3 : 2;
}
