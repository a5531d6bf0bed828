// pos_index.cu -- gallery identity index and exact positive scores.
//
// Replaces the per-query `is_pos = (g_pids == pid) & mask` scan of eval_mm_protocol.py:427-428:
// the gallery is sorted by pid once (CUB radix sort = library plumbing, one-time per gallery);
// a query's positives are then the contiguous run order[code .. code+count).  Their scores are
// computed exactly in fp32 (reid_pos_scores) and sorted descending (reid_pos_sort); these are the
// thresholds the fused GEMM epilogue counts against (rank_j = 1 + #above_j + j).
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>

namespace {

__global__ void iota_kernel(int32_t* v, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    v[i] = (int32_t)i;
}

// longest run of equal keys in a sorted array: a run starts where key[i] != key[i-1]; its length
// is found by scanning forward from the run head (runs are short: images per identity)
__global__ void max_run_kernel(const int64_t* __restrict__ sorted, int64_t n, int32_t* max_run) {
  int best = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (i == 0 || sorted[i] != sorted[i - 1]) {
      int64_t j = i + 1;
      while (j < n && sorted[j] == sorted[i]) ++j;
      best = max(best, (int)(j - i));
    }
  }
  best = __reduce_max_sync(0xffffffffu, best);
  if ((threadIdx.x & 31) == 0 && best > 0) atomicMax(max_run, best);
}

__global__ void pid_lookup_kernel(const int64_t* __restrict__ sorted, int64_t G, const int64_t* __restrict__ pids,
                                  int64_t n, int32_t* __restrict__ code, int32_t* __restrict__ count) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t key = pids[i];
    int64_t lo = 0, hi = G;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (sorted[mid] < key) lo = mid + 1; else hi = mid; }
    const int64_t lb = lo;
    hi = G;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (sorted[mid] <= key) lo = mid + 1; else hi = mid; }
    const bool found = (lb < G) && (sorted[lb] == key);
    code[i] = found ? (int32_t)lb : -1;
    if (count) count[i] = found ? (int32_t)(lo - lb) : 0;
  }
}

// one CTA (4 warps) per query; warp w scores slots w, w+4, ...
__global__ void __launch_bounds__(128)
pos_scores_kernel(const float* __restrict__ q_f32, const float* __restrict__ g_f32, const int32_t* __restrict__ order,
                  const int32_t* __restrict__ q_code, const int32_t* __restrict__ q_count,
                  const int32_t* __restrict__ excl, int E, int64_t G_local, int64_t g_offset, int d, int Pmax,
                  float* __restrict__ pos_score) {
  const int64_t q = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int code = q_code[q];
  const int cnt = (code >= 0) ? min(q_count[q], Pmax) : 0;
  const float* qrow = q_f32 + q * (int64_t)d;
  // two slots per warp at a time: both row gathers in flight (results bit-identical to warp_dot)
  for (int s = warp; s < Pmax; s += 8) {
    const int s2 = s + 4;
    int64_t g0 = -1, g1 = -1;
    if (s < cnt) g0 = order[code + s];
    if (s2 < cnt) g1 = order[code + s2];
    bool ok0 = (g0 >= g_offset) && (g0 < g_offset + G_local);
    bool ok1 = (g1 >= g_offset) && (g1 < g_offset + G_local);
    for (int e = 0; e < E; ++e) {
      const int32_t x = excl[q * E + e];
      ok0 = ok0 && (x != (int32_t)g0);
      ok1 = ok1 && (x != (int32_t)g1);
    }
    float v0 = REID_NEG_INF, v1 = REID_NEG_INF;
    if (d == 512 && ok0 && ok1) {
      warp_dot2_512(qrow, g_f32 + (g0 - g_offset) * (int64_t)d, g_f32 + (g1 - g_offset) * (int64_t)d, lane, v0, v1);
    } else {
      if (ok0) v0 = warp_dot(qrow, g_f32 + (g0 - g_offset) * (int64_t)d, d, lane);
      if (ok1) v1 = warp_dot(qrow, g_f32 + (g1 - g_offset) * (int64_t)d, d, lane);
    }
    if (lane == 0) {
      pos_score[q * Pmax + s] = v0;
      if (s2 < Pmax) pos_score[q * Pmax + s2] = v1;
    }
  }
}

// one CTA per query: bitonic sort (descending) of Pmax (<= 2048) scores in shared memory
__global__ void __launch_bounds__(256)
pos_sort_kernel(float* __restrict__ pos_score, int32_t* __restrict__ n_pos, int Pmax, int P2) {
  extern __shared__ float sm[];
  const int64_t q = blockIdx.x;
  float* row = pos_score + q * (int64_t)Pmax;
  int local_cnt = 0;
  for (int i = threadIdx.x; i < P2; i += blockDim.x) {
    const float v = (i < Pmax) ? row[i] : REID_NEG_INF;
    sm[i] = v;
    local_cnt += (v > REID_NEG_INF) ? 1 : 0;
  }
  __syncthreads();
  for (int k = 2; k <= P2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const float a = sm[i], b = sm[ixj];
          const bool desc = ((i & k) == 0);
          if (desc ? (a < b) : (a > b)) { sm[i] = b; sm[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < Pmax; i += blockDim.x) row[i] = sm[i];
  // count of finite entries
  __shared__ int total;
  if (threadIdx.x == 0) total = 0;
  __syncthreads();
  local_cnt = __reduce_add_sync(0xffffffffu, local_cnt);
  if ((threadIdx.x & 31) == 0 && local_cnt) atomicAdd(&total, local_cnt);
  __syncthreads();
  if (threadIdx.x == 0) n_pos[q] = total;
}

// Pmax <= 64 (the usual case: images per identity): one WARP per query, two values per lane, the bitonic network in
// registers (shuffles); 8 queries per CTA.  Same result as pos_sort_kernel.
__global__ void __launch_bounds__(256)
pos_sort_warp_kernel(float* __restrict__ pos_score, int32_t* __restrict__ n_pos, int64_t Q, int Pmax) {
  const int lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (q >= Q) return;
  float* row = pos_score + q * (int64_t)Pmax;
  float v0 = lane < Pmax ? row[lane] : REID_NEG_INF;
  float v1 = lane + 32 < Pmax ? row[lane + 32] : REID_NEG_INF;
  const int cnt = __popc(__ballot_sync(0xffffffffu, v0 > REID_NEG_INF)) + __popc(__ballot_sync(0xffffffffu, v1 > REID_NEG_INF));
#pragma unroll
  for (int k = 2; k <= 64; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j == 32) {                                   // (k == 64) partner = the lane's other value; block 0: descending
        const float a = fmaxf(v0, v1), b = fminf(v0, v1);
        v0 = a; v1 = b;
      } else {
        const float o0 = __shfl_xor_sync(0xffffffffu, v0, j), o1 = __shfl_xor_sync(0xffffffffu, v1, j);
        const bool lower = (lane & j) == 0;
        const bool d0 = (lane & k) == 0, d1 = ((lane + 32) & k) == 0;      // descending sub-block?
        v0 = (lower == d0) ? fmaxf(v0, o0) : fminf(v0, o0);
        v1 = (lower == d1) ? fmaxf(v1, o1) : fminf(v1, o1);
      }
    }
  }
  if (lane < Pmax) row[lane] = v0;
  if (lane + 32 < Pmax) row[lane + 32] = v1;
  if (lane == 0) n_pos[q] = cnt;
}

}  // namespace

static size_t pid_index_ws_bytes(int64_t G) {
  size_t temp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, temp, (const int64_t*)nullptr, (int64_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)G);
  const size_t iota = ((size_t)G * sizeof(int32_t) + 255) & ~(size_t)255;
  return iota + temp + 256;
}

extern "C" size_t reid_pid_index_workspace_bytes(int64_t G) { return pid_index_ws_bytes(G); }

extern "C" int reid_pid_index_build(const int64_t* g_pid, int64_t G, int64_t* sorted_pid, int32_t* order,
                                    int32_t* max_run, void* workspace, size_t workspace_bytes, void* stream) {
  if (!g_pid || !sorted_pid || !order || !max_run || !workspace || G <= 0 || G > 0x7fffffffLL) return REID_E_INVALID;
  if (workspace_bytes < pid_index_ws_bytes(G)) return REID_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  int32_t* iota = (int32_t*)workspace;
  const size_t iota_bytes = ((size_t)G * sizeof(int32_t) + 255) & ~(size_t)255;
  void* temp = (char*)workspace + iota_bytes;
  size_t temp_bytes = workspace_bytes - iota_bytes;
  iota_kernel<<<(int)reid_min64((G + 255) / 256, 148 * 8), 256, 0, st>>>(iota, G);
  if (cub::DeviceRadixSort::SortPairs(temp, temp_bytes, g_pid, sorted_pid, iota, order, (int)G, 0, 64, st) != cudaSuccess)
    return REID_E_CUDA;
  if (cudaMemsetAsync(max_run, 0, sizeof(int32_t), st) != cudaSuccess) return REID_E_CUDA;
  max_run_kernel<<<(int)reid_min64((G + 255) / 256, 148 * 8), 256, 0, st>>>(sorted_pid, G, max_run);
  REID_CHECK_LAUNCH();
  return REID_OK;
}

extern "C" int reid_pid_lookup(const int64_t* sorted_pid, int64_t G, const int64_t* pids, int64_t n,
                               int32_t* code, int32_t* count, void* stream) {
  if (!sorted_pid || !pids || !code || G <= 0 || n < 0) return REID_E_INVALID;
  if (n == 0) return REID_OK;
  pid_lookup_kernel<<<(int)reid_min64((n + 255) / 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(
      sorted_pid, G, pids, n, code, count);
  REID_CHECK_LAUNCH();
  return REID_OK;
}

extern "C" int reid_pos_scores(const float* q_f32, const float* g_f32, const int32_t* order, const int32_t* q_code,
                               const int32_t* q_count, const int32_t* excl, int E, int64_t Q, int64_t G_local,
                               int64_t g_offset, int d, int Pmax, float* pos_score, void* stream) {
  if (!q_f32 || !g_f32 || !order || !q_code || !q_count || !pos_score || Q < 0 || d <= 0 || d % 4 != 0 || Pmax <= 0 ||
      (E > 0 && !excl))
    return REID_E_INVALID;
  if (Q == 0) return REID_OK;
  pos_scores_kernel<<<(unsigned)Q, 128, 0, (cudaStream_t)stream>>>(q_f32, g_f32, order, q_code, q_count, excl, E,
                                                                  G_local, g_offset, d, Pmax, pos_score);
  REID_CHECK_LAUNCH();
  return REID_OK;
}

extern "C" int reid_pos_sort(float* pos_score, int32_t* n_pos, int64_t Q, int Pmax, void* stream) {
  if (!pos_score || !n_pos || Q < 0 || Pmax <= 0 || Pmax > 2048) return REID_E_INVALID;
  if (Q == 0) return REID_OK;
  if (Pmax <= 64) {
    pos_sort_warp_kernel<<<(unsigned)((Q + 7) / 8), 256, 0, (cudaStream_t)stream>>>(pos_score, n_pos, Q, Pmax);
    REID_CHECK_LAUNCH();
    return REID_OK;
  }
  int P2 = 1;
  while (P2 < Pmax) P2 <<= 1;
  pos_sort_kernel<<<(unsigned)Q, 256, P2 * sizeof(float), (cudaStream_t)stream>>>(pos_score, n_pos, Pmax, P2);
  REID_CHECK_LAUNCH();
  return REID_OK;
}
