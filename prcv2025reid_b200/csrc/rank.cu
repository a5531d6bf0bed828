// rank.cu -- exact fp32 retrieval (SIMT), candidate re-scoring, shard merge, metrics.
//
// Replaces eval_mm_protocol.py:401-469 (sims, mask, argsort, CMC, AP) without an argsort:
//   rank_j = 1 + #{valid non-positive g : s_g > s_pos_j} + j   (positives sorted descending)
//   AP     = (1/P) sum_j (j+1) / rank_j          (== the reference's walk down the full ranking)
//   hit@k  = rank_0 <= k
// reid_retrieve_exact is the all-fp32 CUDA-core form (fallback for flagged queries, tiny
// problems, in-library cross-check of the tcgen05 path); reid_rescore_topk turns per-chunk
// candidate buffers into an exactly ordered top list; reid_metrics_reduce produces the dict.
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// exact path: CTA = 8 warps, QT queries staged in smem, gallery chunk streamed by warps.
// ------------------------------------------------------------------------------------------
constexpr int XQT = 8;        // queries per CTA
constexpr int XWARPS = 8;

__global__ void __launch_bounds__(XWARPS * 32)
retrieve_exact_kernel(const float* __restrict__ q_f32, const float* __restrict__ g_f32,
                      const int32_t* __restrict__ q_code, const int32_t* __restrict__ g_code,
                      const int32_t* __restrict__ excl, int E, const float* __restrict__ pos_thr,
                      const int32_t* __restrict__ n_pos, const int32_t* __restrict__ q_sel, int64_t n_sel,
                      int64_t G_local, int64_t g_offset, int d, int Pmax, int n_chunks, int cand_cap,
                      int32_t* __restrict__ pos_above, float* __restrict__ cand_score,
                      int32_t* __restrict__ cand_idx, int32_t* __restrict__ cand_count) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sq = reinterpret_cast<float*>(smem_raw);                  // [XQT][d]
  float* sthr = sq + XQT * d;                                      // [XQT][Pmax]
  int32_t* scnt = reinterpret_cast<int32_t*>(sthr + XQT * Pmax);   // [XQT][Pmax]
  int32_t* sexcl = scnt + XQT * Pmax;                              // [XQT][max(E,1)]
  float* ms = reinterpret_cast<float*>(sexcl + XQT * (E > 0 ? E : 1));   // [XQT][XWARPS*32] merge staging: scores
  int* mi = reinterpret_cast<int*>(ms + XQT * XWARPS * 32);              // [XQT][XWARPS*32] merge staging: rows
  __shared__ int s_qidx[XQT], s_code[XQT], s_npos[XQT];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunk = blockIdx.y;
  const int64_t qt0 = (int64_t)blockIdx.x * XQT;
  const int Ee = E > 0 ? E : 1;

  if (threadIdx.x < XQT) {
    const int64_t sel = qt0 + threadIdx.x;
    int qi = -1;
    if (sel < n_sel) qi = q_sel ? q_sel[sel] : (int)sel;
    s_qidx[threadIdx.x] = qi;
    s_code[threadIdx.x] = qi >= 0 ? q_code[qi] : -2;
    s_npos[threadIdx.x] = qi >= 0 ? min(n_pos[qi], Pmax) : 0;
  }
  __syncthreads();
  for (int t = 0; t < XQT; ++t) {
    const int qi = s_qidx[t];
    for (int c = threadIdx.x; c < d; c += blockDim.x) sq[t * d + c] = qi >= 0 ? q_f32[(int64_t)qi * d + c] : 0.f;
    for (int p = threadIdx.x; p < Pmax; p += blockDim.x) {
      sthr[t * Pmax + p] = qi >= 0 ? pos_thr[(int64_t)qi * Pmax + p] : INFINITY;
      scnt[t * Pmax + p] = 0;
    }
    for (int e = threadIdx.x; e < Ee; e += blockDim.x)
      sexcl[t * Ee + e] = (qi >= 0 && E > 0) ? excl[(int64_t)qi * E + e] : -1;
  }
  __syncthreads();

  // per-warp running top list: lane i holds entry i of each query's list
  // (all of s, lmin, nfill are warp-uniform; only lv / li differ per lane)
  float lv[XQT]; int li[XQT]; float lmin[XQT]; int nfill[XQT];
#pragma unroll
  for (int t = 0; t < XQT; ++t) { lv[t] = REID_NEG_INF; li[t] = -1; lmin[t] = REID_NEG_INF; nfill[t] = 0; }

  const int64_t rows_per_chunk = (G_local + n_chunks - 1) / n_chunks;
  const int64_t c0 = chunk * rows_per_chunk;
  const int64_t c1 = reid_min64(G_local, c0 + rows_per_chunk);
  // gridDim.z CTAs share one (query tile, chunk): each takes a contiguous slice of the chunk
  const int64_t rows_per_split = (c1 - c0 + gridDim.z - 1) / gridDim.z;
  const int64_t r0 = c0 + blockIdx.z * rows_per_split;
  const int64_t r1 = reid_min64(c1, r0 + rows_per_split);
  for (int64_t r = r0 + warp; r < r1; r += XWARPS) {
    const float* grow = g_f32 + r * (int64_t)d;
    const int gcode = g_code[r];
    const int32_t gidx = (int32_t)(g_offset + r);
#pragma unroll
    for (int t = 0; t < XQT; ++t) {
      if (s_qidx[t] < 0) continue;
      const float s = warp_dot(sq + t * d, grow, d, lane);
      bool masked = false;
      for (int e = 0; e < Ee; ++e) masked |= (sexcl[t * Ee + e] == gidx);
      if (masked) continue;                                     // sims_masked = -1e9 (:421-422)
      const int np = s_npos[t];
      if (gcode != s_code[t] && np > 0 && s > sthr[t * Pmax + np - 1]) {
        for (int p = lane; p < np; p += 32)
          if (s > sthr[t * Pmax + p]) atomicAdd(&scnt[t * Pmax + p], 1);
      }
      if (nfill[t] < 32) {                                      // list not full yet: fill lane by lane
        if (lane == nfill[t]) { lv[t] = s; li[t] = (int)r; }
        if (++nfill[t] == 32) lmin[t] = warp_min(lv[t]);
      } else if (s > lmin[t]) {                                 // replace the current minimum entry
        const unsigned holders = __ballot_sync(0xffffffffu, lv[t] == lmin[t]);
        if (lane == __ffs(holders) - 1) { lv[t] = s; li[t] = (int)r; }
        lmin[t] = warp_min(lv[t]);
      }
    }
  }
  __syncthreads();
  // flush: counts and per-warp lists
  for (int t = 0; t < XQT; ++t) {
    const int qi = s_qidx[t];
    if (qi < 0) continue;
    for (int p = threadIdx.x; p < s_npos[t]; p += blockDim.x) {
      const int c = scnt[t * Pmax + p];
      if (c) atomicAdd(&pos_above[(int64_t)qi * Pmax + p], c);
    }
  }
  // merge the XWARPS per-warp lists of every query into the CTA's best 32 (staged in shared memory)
  // so that a (query, chunk) buffer receives at most 32 entries per CTA
  __syncthreads();
#pragma unroll
  for (int t = 0; t < XQT; ++t) {
    ms[t * XWARPS * 32 + warp * 32 + lane] = (li[t] >= 0) ? lv[t] : REID_NEG_INF;
    mi[t * XWARPS * 32 + warp * 32 + lane] = li[t];
  }
  __syncthreads();
  if (warp < XQT && s_qidx[warp] >= 0) {
    const int t = warp, qi = s_qidx[t];
    float v[XWARPS]; int ix[XWARPS];
#pragma unroll
    for (int k = 0; k < XWARPS; ++k) { v[k] = ms[t * XWARPS * 32 + k * 32 + lane]; ix[k] = mi[t * XWARPS * 32 + k * 32 + lane]; }
    float outv = REID_NEG_INF; int outi = -1;
    for (int round = 0; round < 32; ++round) {      // 32 rounds of warp-wide arg-max (score desc, row asc)
      float bv = REID_NEG_INF; int bi = 0x7fffffff; int bk = -1;
#pragma unroll
      for (int k = 0; k < XWARPS; ++k)
        if (ix[k] >= 0 && ranks_before(v[k], ix[k], bv, bi)) { bv = v[k]; bi = ix[k]; bk = k; }
      float wv = bv; int wi = bi;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, wv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
        if (ranks_before(ov, oi, wv, wi)) { wv = ov; wi = oi; }
      }
      if (wi == 0x7fffffff) break;                   // fewer than 32 rows in this slice
      if (bk >= 0 && bi == wi) {                     // the owning lane retires its entry
#pragma unroll
        for (int k = 0; k < XWARPS; ++k) if (k == bk) ix[k] = -1;
      }
      if (lane == round) { outv = wv; outi = wi; }
    }
    const unsigned have = __ballot_sync(0xffffffffu, outi >= 0);
    if (have) {
      int base = 0;
      if (lane == 0) base = atomicAdd(&cand_count[(int64_t)qi * n_chunks + chunk], __popc(have));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (outi >= 0) {
        const int slot = base + __popc(have & ((1u << lane) - 1));
        if (slot < cand_cap) {
          const int64_t o = ((int64_t)qi * n_chunks + chunk) * cand_cap + slot;
          cand_score[o] = outv; cand_idx[o] = outi;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// candidate selection + re-scoring
//   cand_select:  one CTA (4 warps) per query: gather the query's candidate slots, sort by approximate score, keep the
//                 REID_RTOP best -> sel_score / sel_idx (sorted descending) and sel_cut = the kx-th best approximate score.
//   (multi-shard: the host all-reduces sel_cut with MAX -> bound: the best shard's kx-th best score)
//   rescore_topk: one CTA per query: the selected rows at or above `bound` are re-scored in fp32, sorted, written as the
//                 shard's exact top list; exact local counts for the positives above bound + eps.
//   topk_check:   after the exchange: is the merged top-k / CMC decidable within eps?
// Completeness: a local row that is NOT re-scored has an approximate score <= bound (below it inside the selected list, or
// <= the REID_KLIST-th best <= sel_cut <= bound outside it), hence an exact score <= bound + eps.
// ------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 128;

__device__ void block_bitonic_desc(float* key, int32_t* val, int n2) {
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const float a = key[i], b = key[ixj];
          const int ia = val[i], ib = val[ixj];
          const bool first_block = ((i & k) == 0);
          // descending overall: in a "first" block the better element goes to the lower index
          const bool swap = first_block ? ranks_before(b, ib, a, ia) : ranks_before(a, ia, b, ib);
          if (swap) { key[i] = b; key[ixj] = a; val[i] = ib; val[ixj] = ia; }
        }
      }
      __syncthreads();
    }
  }
}

// one query by a whole CTA (any number of candidates up to n2): the fallback of cand_select_warp_kernel
__device__ void cand_select_cta(int qi, const float* __restrict__ cand_score, const int32_t* __restrict__ cand_idx,
                                const int32_t* __restrict__ cand_count, const float* __restrict__ cand_thr, int n_chunks,
                                int cand_cap, int kx, int n2, float* __restrict__ sel_score, int32_t* __restrict__ sel_idx,
                                int32_t* __restrict__ sel_n, float* __restrict__ sel_cut, int32_t* __restrict__ sel_flag) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* key = reinterpret_cast<float*>(smem_raw);        // [n2]
  int32_t* val = reinterpret_cast<int32_t*>(key + n2);    // [n2]
  __shared__ int s_total, s_overflow;
  __syncthreads();                                        // (the previous query of this CTA is done with the shared state)
  if (threadIdx.x == 0) { s_total = 0; s_overflow = 0; }
  __syncthreads();
  // gather the chunk buffers (compact, order irrelevant: sorted next).  When the producer supplied a
  // per-query threshold with >= KLIST candidates at or above it, only those can reach the top list.
  // The chunk counts are fetched first and the candidates are read over the flattened index space, four
  // independent loads per thread in flight (one memory round trip per 512 candidates, not one per chunk slice).
  const float keep = cand_thr ? cand_thr[qi] : REID_NEG_INF;
  __shared__ int s_pre[65];                                   // prefix sums of the clamped chunk counts (n_chunks <= 64)
  const int nck = n_chunks < 64 ? n_chunks : 64;
  if (threadIdx.x < nck) {
    int cnt = cand_count[(int64_t)qi * n_chunks + threadIdx.x];
    if (cnt > cand_cap) { cnt = cand_cap; s_overflow = 1; }
    s_pre[threadIdx.x + 1] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    s_pre[0] = 0;
    for (int c = 0; c < nck; ++c) s_pre[c + 1] += s_pre[c];
  }
  __syncthreads();
  const int total_in = s_pre[nck];
  const int64_t qbase = (int64_t)qi * n_chunks * cand_cap;
  for (int b = threadIdx.x; b < total_in; b += 4 * RS_THREADS) {
    float v[4]; int64_t off[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = b + u * RS_THREADS;
      v[u] = REID_NEG_INF; off[u] = 0;
      if (e < total_in) {
        int c = 0;
        while (c + 1 < nck && e >= s_pre[c + 1]) ++c;
        off[u] = qbase + (int64_t)c * cand_cap + (e - s_pre[c]);
        v[u] = cand_score[off[u]];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (b + u * RS_THREADS < total_in && v[u] >= keep) {
        const int slot = atomicAdd(&s_total, 1);
        if (slot < n2) { key[slot] = v[u]; val[slot] = cand_idx[off[u]]; }
      }
    }
  }
  for (int c = 64; c < n_chunks; ++c) {                        // (more than 64 chunks: never produced by the host side)
    int cnt = cand_count[(int64_t)qi * n_chunks + c];
    if (cnt > cand_cap) { cnt = cand_cap; if (threadIdx.x == 0) s_overflow = 1; }
    const int64_t o = qbase + (int64_t)c * cand_cap;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      const float v = cand_score[o + i];
      if (v >= keep) {
        const int slot = atomicAdd(&s_total, 1);
        if (slot < n2) { key[slot] = v; val[slot] = cand_idx[o + i]; }
      }
    }
  }
  __syncthreads();
  int total = s_total;
  if (total > n2) { total = n2; if (threadIdx.x == 0) s_overflow = 1; }   // staging full: re-run exactly
  int ns = 32;                       // sort only the occupied power-of-two prefix
  while (ns < total) ns <<= 1;
  for (int i = total + threadIdx.x; i < ns; i += blockDim.x) { key[i] = REID_NEG_INF; val[i] = 0x7fffffff; }
  __syncthreads();
  block_bitonic_desc(key, val, ns);
  const int R = min(total, REID_RTOP);
  if (threadIdx.x < REID_RTOP) {
    sel_score[(int64_t)qi * REID_RTOP + threadIdx.x] = threadIdx.x < R ? key[threadIdx.x] : REID_NEG_INF;
    sel_idx[(int64_t)qi * REID_RTOP + threadIdx.x] = threadIdx.x < R ? val[threadIdx.x] : -1;
  }
  if (threadIdx.x == 0) {
    sel_n[qi] = R;
    // completeness cut-off: every local row whose approximate score exceeds it is among the selected rows
    // (fewer than kx candidates: the shard has no other rows to offer, -inf)
    sel_cut[qi] = (total >= kx) ? key[kx - 1] : REID_NEG_INF;
    sel_flag[qi] = s_overflow ? 1 : 0;                        // bit0: a candidate buffer overflowed
  }
}

// Fallback pass: the queries cand_select_warp_kernel marked (sel_n == -1: more than CSW_MAX candidates at or above the
// threshold), one CTA per query; the CTAs of unmarked queries leave at once.
__global__ void __launch_bounds__(RS_THREADS)
cand_select_kernel(const float* __restrict__ cand_score, const int32_t* __restrict__ cand_idx,
                   const int32_t* __restrict__ cand_count, const float* __restrict__ cand_thr, int64_t Q, int n_chunks,
                   int cand_cap, int kx, int n2, float* __restrict__ sel_score, int32_t* __restrict__ sel_idx,
                   int32_t* __restrict__ sel_n, float* __restrict__ sel_cut, int32_t* __restrict__ sel_flag) {
  const int qi = blockIdx.x;
  if (sel_n[qi] != -1) return;                               // (uniform over the CTA)
  cand_select_cta(qi, cand_score, cand_idx, cand_count, cand_thr, n_chunks, cand_cap, kx, n2, sel_score, sel_idx, sel_n, sel_cut,
                  sel_flag);
}

// The common case -- at most CSW_MAX (128) candidates at or above the query's threshold -- by ONE WARP per query, in
// registers: the kept candidates are compacted into a per-warp staging list (ballot + popc), every 32 of them are sorted with
// a shuffle bitonic network (score desc, index asc: the total order of ranks_before) and folded into the running best 32
// (best(a[i], b[31 - i]) of two sorted lists is a bitonic sequence that holds the 32 best of their union; five merge stages
// sort it).  Same outputs as the CTA form: the REID_RTOP best in order, their number, the kx-th best as the cut-off, the
// overflow bit.  Queries with more candidates are marked (sel_n = -1) for cand_select_kernel.
constexpr int CSW_MAX = 128;
constexpr int CSW_WARPS = 8;

__device__ __forceinline__ void warp_bitonic_merge_desc(float& s, int& gi, int lane, int jmax) {
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) {
    if (j > jmax) continue;
    const float so = __shfl_xor_sync(0xffffffffu, s, j);
    const int io = __shfl_xor_sync(0xffffffffu, gi, j);
    const bool other_better = ranks_before(so, io, s, gi);
    const bool take = ((lane & j) == 0) ? other_better : !other_better;   // the lower lane keeps the better element
    if (take) { s = so; gi = io; }
  }
}
__device__ __forceinline__ void warp_bitonic_sort_desc(float& s, int& gi, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const float so = __shfl_xor_sync(0xffffffffu, s, j);
      const int io = __shfl_xor_sync(0xffffffffu, gi, j);
      const bool lower = (lane & j) == 0;
      const bool first_block = (lane & k) == 0 || k == 32;                // (the last level sorts the whole warp descending)
      const bool other_better = ranks_before(so, io, s, gi);
      const bool take = (lower == first_block) ? other_better : !other_better;
      if (take) { s = so; gi = io; }
    }
  }
}

__global__ void __launch_bounds__(CSW_WARPS * 32)
cand_select_warp_kernel(const float* __restrict__ cand_score, const int32_t* __restrict__ cand_idx,
                        const int32_t* __restrict__ cand_count, const float* __restrict__ cand_thr, int64_t Q, int n_chunks,
                        int cand_cap, int kx, float* __restrict__ sel_score, int32_t* __restrict__ sel_idx,
                        int32_t* __restrict__ sel_n, float* __restrict__ sel_cut, int32_t* __restrict__ sel_flag) {
  __shared__ float s_key[CSW_WARPS][CSW_MAX];
  __shared__ int32_t s_val[CSW_WARPS][CSW_MAX];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t q = (int64_t)blockIdx.x * CSW_WARPS + warp;
  if (q >= Q) return;
  const float keep = cand_thr ? cand_thr[q] : REID_NEG_INF;
  const unsigned lt = (1u << lane) - 1u;
  int kept = 0;
  bool overflow = false;
  for (int c = 0; c < n_chunks; ++c) {
    int cnt = cand_count[q * n_chunks + c];
    if (cnt > cand_cap) { cnt = cand_cap; overflow = true; }
    const int64_t o = (q * n_chunks + c) * (int64_t)cand_cap;
    for (int b = 0; b < cnt; b += 64) {                       // two independent loads per lane in flight
      const int i0 = b + lane, i1 = b + 32 + lane;
      const float v0 = i0 < cnt ? cand_score[o + i0] : REID_NEG_INF, v1 = i1 < cnt ? cand_score[o + i1] : REID_NEG_INF;
      const bool k0 = i0 < cnt && v0 >= keep, k1 = i1 < cnt && v1 >= keep;
      const unsigned m0 = __ballot_sync(0xffffffffu, k0), m1 = __ballot_sync(0xffffffffu, k1);
      if (k0) { const int pos = kept + __popc(m0 & lt); if (pos < CSW_MAX) { s_key[warp][pos] = v0; s_val[warp][pos] = cand_idx[o + i0]; } }
      kept += __popc(m0);
      if (k1) { const int pos = kept + __popc(m1 & lt); if (pos < CSW_MAX) { s_key[warp][pos] = v1; s_val[warp][pos] = cand_idx[o + i1]; } }
      kept += __popc(m1);
    }
  }
  if (kept > CSW_MAX) {                                       // (warp-uniform) too many for this form: the CTA pass takes the query
    if (lane == 0) sel_n[q] = -1;
    return;
  }
  __syncwarp();
  float bs = REID_NEG_INF; int bi = 0x7fffffff;               // running best 32, sorted descending over the lanes
  for (int c0 = 0; c0 < kept; c0 += 32) {
    float s = c0 + lane < kept ? s_key[warp][c0 + lane] : REID_NEG_INF;
    int gi = c0 + lane < kept ? s_val[warp][c0 + lane] : 0x7fffffff;
    warp_bitonic_sort_desc(s, gi, lane);
    if (c0 == 0) { bs = s; bi = gi; continue; }
    const float so = __shfl_sync(0xffffffffu, s, 31 - lane);
    const int io = __shfl_sync(0xffffffffu, gi, 31 - lane);
    if (ranks_before(so, io, bs, bi)) { bs = so; bi = io; }
    warp_bitonic_merge_desc(bs, bi, lane, 16);
  }
  const int R = min(kept, REID_RTOP);
  sel_score[q * REID_RTOP + lane] = lane < R ? bs : REID_NEG_INF;
  sel_idx[q * REID_RTOP + lane] = lane < R ? bi : -1;
  const float cut = __shfl_sync(0xffffffffu, bs, kx - 1);
  if (lane == 0) {
    sel_n[q] = R;
    sel_cut[q] = kept >= kx ? cut : REID_NEG_INF;
    sel_flag[q] = overflow ? 1 : 0;
  }
}

__global__ void __launch_bounds__(RS_THREADS)
rescore_topk_kernel(const float* __restrict__ q_f32, const float* __restrict__ g_f32,
                    const int32_t* __restrict__ q_code, const int32_t* __restrict__ g_code,
                    const float* __restrict__ pos_thr, const int32_t* __restrict__ n_pos,
                    const float* __restrict__ sel_score, const int32_t* __restrict__ sel_idx,
                    const int32_t* __restrict__ sel_n, const float* __restrict__ bound_in,
                    int64_t g_offset, int d, int Pmax, float eps, int32_t* __restrict__ pos_above,
                    float* __restrict__ top_score, int32_t* __restrict__ top_idx, int32_t* __restrict__ lb0) {
  __shared__ float ex_s[REID_RTOP];
  __shared__ int32_t ex_i[REID_RTOP];
  __shared__ int s_neg[REID_RTOP];                             // re-scored row r is NOT a positive of the query
  const int qi = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float b = bound_in[qi];
  // rows to re-score: the prefix of the (descending) selected list at or above the bound
  const int n_sel = sel_n[qi];
  const float my = lane < n_sel ? sel_score[(int64_t)qi * REID_RTOP + lane] : REID_NEG_INF;
  const int R = __popc(__ballot_sync(0xffffffffu, lane < n_sel && my >= b));
  // exact fp32 re-score (same dot routine as reid_pos_scores)
  constexpr int RW = RS_THREADS / 32;
  const float* qrow = q_f32 + (int64_t)qi * d;
  for (int r = warp; r < REID_RTOP; r += 2 * RW) {               // two rows per warp at a time: both gathers in flight
    const int r2 = r + RW;
    float s0 = REID_NEG_INF, s1 = REID_NEG_INF; int g0 = 0x7fffffff, g1 = 0x7fffffff;
    if (r < R) g0 = sel_idx[(int64_t)qi * REID_RTOP + r];
    if (r2 < R) g1 = sel_idx[(int64_t)qi * REID_RTOP + r2];
    if (d == 512 && r2 < R) {
      warp_dot2_512(qrow, g_f32 + (int64_t)g0 * d, g_f32 + (int64_t)g1 * d, lane, s0, s1);
    } else {
      if (r < R) s0 = warp_dot(qrow, g_f32 + (int64_t)g0 * d, d, lane);
      if (r2 < R) s1 = warp_dot(qrow, g_f32 + (int64_t)g1 * d, d, lane);
    }
    if (lane == 0) { ex_s[r] = s0; ex_i[r] = g0; if (r2 < REID_RTOP) { ex_s[r2] = s1; ex_i[r2] = g1; } }
  }
  __syncthreads();
  if (warp == 0) {
    // warp bitonic over 32 entries (score desc, idx asc)
    float s = ex_s[lane]; int gi = ex_i[lane];
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
      for (int j = k >> 1; j > 0; j >>= 1) {
        const float so = __shfl_xor_sync(0xffffffffu, s, j);
        const int io = __shfl_xor_sync(0xffffffffu, gi, j);
        const bool lower = (lane & j) == 0;
        const bool first_block = (lane & k) == 0;
        const bool other_better = ranks_before(so, io, s, gi);
        // lower lane of a descending block keeps the better element
        const bool take = (lower == first_block) ? other_better : !other_better;
        if (take) { s = so; gi = io; }
      }
    }
    ex_s[lane] = s; ex_i[lane] = gi;
    top_score[(int64_t)qi * REID_RTOP + lane] = s;
    top_idx[(int64_t)qi * REID_RTOP + lane] = (lane < R) ? (int32_t)(g_offset + gi) : -1;
  }
  __syncthreads();
  // exact counts for the positives above the completeness bound
  const int np = min(n_pos[qi], Pmax);
  const int qcode = q_code[qi];
  const float bound = b + eps;     // no local row that was not re-scored can score above this
  if (threadIdx.x < REID_RTOP) s_neg[threadIdx.x] = (threadIdx.x < R && g_code[ex_i[threadIdx.x]] != qcode) ? 1 : 0;
  __syncthreads();
  for (int j = threadIdx.x; j < np; j += blockDim.x) {
    const float t = pos_thr[(int64_t)qi * Pmax + j];
    int lb = 0;
    for (int r = 0; r < R; ++r) lb += (s_neg[r] && ex_s[r] > t) ? 1 : 0;
    int32_t* dst = pos_above + (int64_t)qi * Pmax + j;
    if (t > bound || b == REID_NEG_INF) *dst = lb;         // exact local count
    else if (*dst < lb) *dst = lb;                         // lb is a rigorous lower bound
    if (j == 0) lb0[qi] = lb;
  }
  if (np == 0 && threadIdx.x == 0) lb0[qi] = 0;
}

// After the exchange (or directly, one shard): can the top-k list and CMC@10 be decided within eps?
//   top: the merged exact top list [Q, list_len]; bound: the (gallery-wide) completeness cut-off; lb0: re-scored rows above
//   the best positive, summed over the shards; flag |= 2 (top-k) / 4 (CMC); flag bit0 (overflow) is kept.
__global__ void __launch_bounds__(256)
topk_check_kernel(const float* __restrict__ top, int list_len, int topk, const float* __restrict__ bound_in, float eps,
                  const float* __restrict__ pos_thr, const int32_t* __restrict__ n_pos, int Pmax,
                  const int32_t* __restrict__ lb0, int64_t Q, int32_t* __restrict__ flag) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < Q; q += (int64_t)gridDim.x * blockDim.x) {
    const float b = bound_in[q];
    int f = flag[q] ? 1 : 0;
    if (b > REID_NEG_INF) {
      const float bound = b + eps;
      if (!(top[q * list_len + topk - 1] >= bound)) f |= 2;                       // k-th best not above the bound (or missing)
      if (n_pos[q] > 0 && !(pos_thr[q * Pmax] > bound) && lb0[q] < 10) f |= 4;    // CMC@10 undecidable within eps
    }
    flag[q] = f;
  }
}

// merge per-shard top lists: one warp-multiple CTA per query, bitonic over n_lists*RTOP entries
__global__ void __launch_bounds__(128)
merge_topk_kernel(const float* __restrict__ scores, const int32_t* __restrict__ idx, int n_lists, int64_t Q,
                  int list_len, int topk, int n2, float* __restrict__ out_score, int32_t* __restrict__ out_idx) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* key = reinterpret_cast<float*>(smem_raw);
  int32_t* val = reinterpret_cast<int32_t*>(key + n2);
  const int64_t q = blockIdx.x;
  const int n = n_lists * list_len;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    float s = REID_NEG_INF; int gi = 0x7fffffff;
    if (i < n) {
      const int l = i / list_len, r = i % list_len;
      const int64_t o = ((int64_t)l * Q + q) * list_len + r;
      gi = idx[o]; s = scores[o];
      if (gi < 0) { gi = 0x7fffffff; s = REID_NEG_INF; }
    }
    key[i] = s; val[i] = gi;
  }
  __syncthreads();
  block_bitonic_desc(key, val, n2);
  for (int i = threadIdx.x; i < topk; i += blockDim.x) {
    const bool ok = (i < n2) && (val[i] != 0x7fffffff);
    out_score[q * topk + i] = ok ? key[i] : REID_NEG_INF;
    out_idx[q * topk + i] = ok ? val[i] : -1;
  }
}

// small unions (n_lists * list_len <= 96: eight shards x top-10): one WARP per query, three entries per lane, topk rounds of a
// warp-wide arg-best (score desc, index asc); eight queries per CTA.  Same result as merge_topk_kernel.
__global__ void __launch_bounds__(256)
merge_topk_warp_kernel(const float* __restrict__ scores, const int32_t* __restrict__ idx, int n_lists, int64_t Q,
                       int list_len, int topk, float* __restrict__ out_score, int32_t* __restrict__ out_idx) {
  const int lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (q >= Q) return;
  const int n = n_lists * list_len;
  float v[3]; int g[3];
#pragma unroll
  for (int u = 0; u < 3; ++u) {
    const int i = lane + 32 * u;
    v[u] = REID_NEG_INF; g[u] = 0x7fffffff;
    if (i < n) {
      const int l = i / list_len, r = i % list_len;
      const int64_t o = ((int64_t)l * Q + q) * list_len + r;
      const int gi = idx[o];
      if (gi >= 0) { g[u] = gi; v[u] = scores[o]; }
    }
  }
  float os = REID_NEG_INF; int oi = -1;
  for (int round = 0; round < topk; ++round) {
    float bv = REID_NEG_INF; int bi = 0x7fffffff;
#pragma unroll
    for (int u = 0; u < 3; ++u)
      if (g[u] != 0x7fffffff && ranks_before(v[u], g[u], bv, bi)) { bv = v[u]; bi = g[u]; }
    float wv = bv; int wi = bi;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, wv, o);
      const int oi2 = __shfl_xor_sync(0xffffffffu, wi, o);
      if (ranks_before(ov, oi2, wv, wi)) { wv = ov; wi = oi2; }
    }
    if (wi == 0x7fffffff) break;                       // the union is exhausted
#pragma unroll
    for (int u = 0; u < 3; ++u)
      if (g[u] == wi && v[u] == wv) g[u] = 0x7fffffff;  // (a gallery row appears in exactly one shard's list)
    if (lane == round) { os = wv; oi = wi; }
  }
  if (lane < topk) { out_score[q * topk + lane] = os; out_idx[q * topk + lane] = oi; }
}

// ------------------------------------------------------------------------------------------
// metrics
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
metrics_kernel(const int32_t* __restrict__ pos_above, const int32_t* __restrict__ n_pos, int64_t Q, int Pmax,
               double* __restrict__ acc /*[5]*/, double* __restrict__ ap_per_query) {
  double ap_sum = 0.0, h1 = 0.0, h5 = 0.0, h10 = 0.0, nv = 0.0;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < Q; q += (int64_t)gridDim.x * blockDim.x) {
    const int np = min(n_pos[q], Pmax);
    double ap = 0.0;
    if (np > 0) {                                            // else: skipped (:430-432)
      const int32_t* row = pos_above + q * (int64_t)Pmax;
      for (int j = 0; j < np; ++j) ap += (double)(j + 1) / (double)(row[j] + j + 1);   // hit / rank_idx (:451)
      ap /= (double)np;                                      // :455
      const int first = row[0] + 1;
      h1 += first <= 1; h5 += first <= 5; h10 += first <= 10;   // :436-438
      ap_sum += ap; nv += 1.0;
    }
    if (ap_per_query) ap_per_query[q] = (np > 0) ? ap : -1.0;
  }
  __shared__ double red[5][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double v[5] = {ap_sum, h1, h5, h10, nv};
#pragma unroll
  for (int k = 0; k < 5; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if (lane == 0) red[k][warp] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
    // The order in which the blocks arrive is not fixed.  The hit counts are integers (exact in a double whatever the order);
    // the AP sum of a block is added as a 64-bit FIXED-POINT number (2^-40 units: integer addition is associative), so the
    // result is bit-identical from run to run and from rank to rank (every rank of a sharded run reduces the same counts).
    if (threadIdx.x == 0) atomicAdd(reinterpret_cast<unsigned long long*>(&acc[0]), (unsigned long long)__double2ll_rn(s * 1099511627776.0));
    else atomicAdd(&acc[threadIdx.x], s);
  }
}

__global__ void metrics_finalize_kernel(double* acc) {
  const double n = acc[4];
  acc[0] = (double)(*reinterpret_cast<unsigned long long*>(&acc[0])) * (1.0 / 1099511627776.0);   // fixed point -> double
  for (int k = 0; k < 4; ++k) acc[k] = n > 0.0 ? acc[k] / n : 0.0;   // np.mean / empty -> 0.0 (:458-461)
}

}  // namespace

extern "C" int reid_retrieve_exact(const float* q_f32, const float* g_f32, const int32_t* q_code,
                                   const int32_t* g_code, const int32_t* excl, int E, const float* pos_thr,
                                   const int32_t* n_pos, const int32_t* q_sel, int64_t n_sel, int64_t Q,
                                   int64_t G_local, int64_t g_offset, int d, int Pmax, int n_chunks, int cand_cap,
                                   int32_t* pos_above, float* cand_score, int32_t* cand_idx, int32_t* cand_count,
                                   void* stream) {
  if (!q_f32 || !g_f32 || !q_code || !g_code || !pos_thr || !n_pos || !pos_above || !cand_score || !cand_idx ||
      !cand_count || d <= 0 || d % 4 != 0 || Pmax <= 0 || n_chunks <= 0 || (E > 0 && !excl))
    return REID_E_INVALID;
  if (cand_cap < 32) return REID_E_INVALID;
  if (!q_sel) n_sel = Q;
  if (n_sel <= 0 || G_local <= 0) return REID_OK;
  const size_t smem = (size_t)XQT * d * 4 + (size_t)XQT * Pmax * 8 + (size_t)XQT * (E > 0 ? E : 1) * 4 +
                      (size_t)XQT * XWARPS * 32 * 8;
  if (smem > 200 * 1024) return REID_E_UNSUPPORTED;
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(retrieve_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return REID_E_CUDA;
  // enough CTAs to fill the GPU even for a handful of flagged queries: split every chunk further;
  // each CTA appends at most 32 candidates per (query, chunk)
  const int64_t qtiles = (n_sel + XQT - 1) / XQT;
  int64_t splits = (148 * 4 + qtiles * n_chunks - 1) / (qtiles * n_chunks);
  const int64_t rows_chunk = (G_local + n_chunks - 1) / n_chunks;
  if (splits > rows_chunk / 512) splits = rows_chunk / 512;
  if (splits > cand_cap / 32) splits = cand_cap / 32;
  if (splits < 1) splits = 1;
  dim3 grid((unsigned)qtiles, (unsigned)n_chunks, (unsigned)splits);
  retrieve_exact_kernel<<<grid, XWARPS * 32, smem, (cudaStream_t)stream>>>(
      q_f32, g_f32, q_code, g_code, excl, E, pos_thr, n_pos, q_sel, n_sel, G_local, g_offset, d, Pmax, n_chunks,
      cand_cap, pos_above, cand_score, cand_idx, cand_count);
  REID_CHECK_LAUNCH();
  return REID_OK;
}

extern "C" int reid_cand_select(const float* cand_score, const int32_t* cand_idx, const int32_t* cand_count,
                                const float* cand_thr, int64_t Q, int n_chunks, int cand_cap, int kx, float* sel_score,
                                int32_t* sel_idx, int32_t* sel_n, float* sel_cut, int32_t* sel_flag, void* stream) {
  if (!cand_score || !cand_idx || !cand_count || !sel_score || !sel_idx || !sel_n || !sel_cut || !sel_flag || n_chunks <= 0 ||
      cand_cap <= 0 || kx <= 0 || kx > REID_KLIST)
    return REID_E_INVALID;
  if (Q <= 0) return REID_OK;
  int n2 = 32;                                   // staging entries: enough for every candidate, at most 4096
  while (n2 < n_chunks * cand_cap && n2 < 4096) n2 <<= 1;
  const size_t smem = (size_t)n2 * 8;
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(cand_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return REID_E_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  // one warp per query for up to 128 kept candidates; the CTA pass only does the queries the warp pass marked
  cand_select_warp_kernel<<<(unsigned)((Q + CSW_WARPS - 1) / CSW_WARPS), CSW_WARPS * 32, 0, st>>>(
      cand_score, cand_idx, cand_count, cand_thr, Q, n_chunks, cand_cap, kx, sel_score, sel_idx, sel_n, sel_cut, sel_flag);
  REID_CHECK_LAUNCH();
  cand_select_kernel<<<(unsigned)Q, RS_THREADS, smem, st>>>(cand_score, cand_idx, cand_count, cand_thr, Q, n_chunks, cand_cap, kx, n2,
                                                     sel_score, sel_idx, sel_n, sel_cut, sel_flag);
  REID_CHECK_LAUNCH();
  return REID_OK;
}

extern "C" int reid_rescore_topk(const float* q_f32, const float* g_f32, const int32_t* q_code, const int32_t* g_code,
                                 const float* pos_thr, const int32_t* n_pos, const float* sel_score, const int32_t* sel_idx,
                                 const int32_t* sel_n, const float* bound, int64_t Q, int64_t G_local, int64_t g_offset,
                                 int d, int Pmax, float eps, int32_t* pos_above, float* top_score, int32_t* top_idx,
                                 int32_t* lb0, void* stream) {
  if (!q_f32 || !g_f32 || !q_code || !g_code || !pos_thr || !n_pos || !sel_score || !sel_idx || !sel_n || !bound ||
      !pos_above || !top_score || !top_idx || !lb0 || d <= 0 || d % 4 != 0 || Pmax <= 0 || G_local <= 0)
    return REID_E_INVALID;
  if (Q <= 0) return REID_OK;
  rescore_topk_kernel<<<(unsigned)Q, RS_THREADS, 0, (cudaStream_t)stream>>>(q_f32, g_f32, q_code, g_code, pos_thr, n_pos, sel_score,
                                                                          sel_idx, sel_n, bound, g_offset, d, Pmax, eps, pos_above,
                                                                          top_score, top_idx, lb0);
  REID_CHECK_LAUNCH();
  return REID_OK;
}

extern "C" int reid_topk_check(const float* top_score, int list_len, int topk, const float* bound, float eps,
                               const float* pos_thr, const int32_t* n_pos, int Pmax, const int32_t* lb0, int64_t Q,
                               int32_t* flag, void* stream) {
  if (!top_score || !bound || !pos_thr || !n_pos || !lb0 || !flag || topk <= 0 || topk > list_len || Pmax <= 0) return REID_E_INVALID;
  if (Q <= 0) return REID_OK;
  topk_check_kernel<<<(int)reid_min64((Q + 255) / 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(top_score, list_len, topk, bound, eps,
                                                                                               pos_thr, n_pos, Pmax, lb0, Q, flag);
  REID_CHECK_LAUNCH();
  return REID_OK;
}

extern "C" int reid_merge_topk(const float* scores, const int32_t* idx, int n_lists, int64_t Q, int list_len, int topk,
                               float* out_score, int32_t* out_idx, void* stream) {
  if (!scores || !idx || !out_score || !out_idx || n_lists <= 0 || topk <= 0 || list_len <= 0 || n_lists * list_len > 8192)
    return REID_E_INVALID;
  if (Q <= 0) return REID_OK;
  if (n_lists * list_len <= 96 && topk <= 32) {
    merge_topk_warp_kernel<<<(unsigned)((Q + 7) / 8), 256, 0, (cudaStream_t)stream>>>(scores, idx, n_lists, Q, list_len, topk,
                                                                                    out_score, out_idx);
    REID_CHECK_LAUNCH();
    return REID_OK;
  }
  int n2 = 32;
  while (n2 < n_lists * list_len) n2 <<= 1;
  merge_topk_kernel<<<(unsigned)Q, 128, (size_t)n2 * 8, (cudaStream_t)stream>>>(scores, idx, n_lists, Q, list_len, topk, n2,
                                                                                 out_score, out_idx);
  REID_CHECK_LAUNCH();
  return REID_OK;
}

extern "C" int reid_metrics_reduce(const int32_t* pos_above, const int32_t* n_pos, int64_t Q, int Pmax,
                                   double* out, double* ap_per_query, void* stream) {
  if (!pos_above || !n_pos || !out || Q < 0 || Pmax <= 0) return REID_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(out, 0, 5 * sizeof(double), st) != cudaSuccess) return REID_E_CUDA;
  if (Q > 0) {
    const int grid = (int)reid_min64((Q + 255) / 256, 148 * 4);
    metrics_kernel<<<grid, 256, 0, st>>>(pos_above, n_pos, Q, Pmax, out, ap_per_query);
  }
  metrics_finalize_kernel<<<1, 1, 0, st>>>(out);
  REID_CHECK_LAUNCH();
  return REID_OK;
}


// ---------------------------------------------------------------------------------------------
// Train-time evaluator reductions (train.py:101-138): given each query's exact top-k gallery list,
//   compute_map : AP@k = mean over the matched positions r (1-based) of (#matches among the first r) / r   (:116-124)
//   compute_cmc : hit@k = any match inside the top-k                                                          (:133-136)
// One warp per query; ap[q] = -1 when the top-k holds no match (the reference skips such queries, :120).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
topk_label_metrics_kernel(const int32_t* __restrict__ top_idx, const int64_t* __restrict__ q_label,
                          const int64_t* __restrict__ g_label, int64_t Q, int list_len, int k,
                          float* __restrict__ ap, int32_t* __restrict__ hit) {
  const int lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= Q) return;
  const int64_t ql = q_label[q];
  int before = 0;        // matches in earlier 32-position groups
  float sum = 0.f;
  for (int r0 = 0; r0 < k; r0 += 32) {
    const int r = r0 + lane;
    bool m = false;
    if (r < k) { const int32_t gi = top_idx[q * list_len + r]; m = gi >= 0 && g_label[gi] == ql; }
    const unsigned b = __ballot_sync(0xffffffffu, m);
    if (m) sum += (float)(before + __popc(b & ((2u << lane) - 1u))) / (float)(r + 1);     // fp32 like the reference
    before += __popc(b);
  }
  sum = warp_sum(sum);
  if (lane == 0) { ap[q] = before > 0 ? sum / (float)before : -1.f; hit[q] = before > 0 ? 1 : 0; }
}

extern "C" int reid_topk_label_metrics(const int32_t* top_idx, const int64_t* q_label, const int64_t* g_label,
                                       int64_t Q, int list_len, int k, float* ap, int32_t* hit, void* stream) {
  if (!top_idx || !q_label || !g_label || !ap || !hit || Q <= 0 || k <= 0 || k > list_len) return REID_E_INVALID;
  topk_label_metrics_kernel<<<(unsigned)((Q + 7) / 8), 256, 0, (cudaStream_t)stream>>>(top_idx, q_label, g_label, Q, list_len, k, ap, hit);
  REID_CHECK_LAUNCH();
  return REID_OK;
}
