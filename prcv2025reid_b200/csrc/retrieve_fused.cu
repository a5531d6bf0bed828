// retrieve_fused.cu -- K3+K4: similarity GEMM with the ranking statistics computed in its epilogue.
//
// Replaces the per-query loop body of rank_and_metrics, eval_mm_protocol.py:401-455 (cosine_sim,
// same-image mask, argsort, CMC, AP walk) for one gallery shard, without materialising S or a sort.
//
// Two layouts of the same kernel (template PAIR; the kernel comment further down has the MMA shapes):
//   PAIR = true (default)  "R layout": a cta_group::2 pair of a 2-CTA cluster per work item = (256 queries, gallery
//     chunk); each CTA keeps its 128 queries resident in shared memory (A operand, 128 KB) and streams its half of
//     every 256-row gallery tile through a 3-stage TMA ring; the leader issues M256 x N256 MMAs into one of two
//     256-column TMEM accumulators.  TMEM lane = query, column = gallery row.
//   PAIR = false "T layout": one CTA per work item = (128 queries, gallery chunk), M128 (gallery rows) x N128 (queries)
//     MMAs into one of four 128-column accumulators; TMEM lane = gallery row, column = query.
// Epilogue (16 warps, four per TMEM lane quadrant, tcgen05.ld 32x32b.x16): every score is compared with
// min(candidate threshold, lowest exactly-counted positive threshold) of its query (R: per-lane registers, T:
// warp-uniform shared values); ONE redux.or per 16-column step tells whether the warp has a hit at all.  Hits are
// compacted round by round (n-th hit of every lane: register select tree + ballot / popc) into a per-warp queue in
// shared memory and drained 32 at a time, lane-parallel, mostly after the accumulator has been handed back:
//       (a) counting: a binary search over the query's positive thresholds (sorted descending, in
//           shared memory) gives the bucket b = #thresholds >= score; hist[q][b]++ (packed 16-bit
//           counters in shared memory, spilled to a global histogram every 128 tiles); the count of
//           rows ranked above positive j is the prefix sum over buckets <= j.
//           Thresholds whose rank inside the chunk is estimated (calibration pre-pass over a strided
//           2048-row sample) to exceed max(32*SAMPLE_W / total_chunks, 8 calibration hits) rows are "deep":
//           they are counted on a fixed 1/SAMPLE_W stratified row sample with weight SAMPLE_W (>= 32 sampled
//           rows above such a threshold gallery-wide: <= 18% unbiased error on a rank > 1000, which moves a
//           query's AP by ~1e-5 and mAP by ~1e-6); all shallower thresholds are counted exactly on every row;
//       (b) candidates: rows above the query's running threshold are appended (lane-parallel) to its
//           candidate buffer in global memory (slots pre-filled with -inf); at tile boundaries the owning warp
//           raises the threshold to the 32nd largest of the last 64 appended scores, so >= 32 appended rows
//           always lie above it.  The threshold reached by one gallery chunk is published (atomicMax) and
//           warm-starts the later chunks of the same query.
//     The epilogue has no CTA barrier inside an item: every shared structure is atomic-safe.
// Roofline: tensor cores, 2*Q*G*d flop; algorithmic HBM bytes are only operands + outputs.
#include "common.cuh"
#include "tc_common.cuh"
#include <stdlib.h>
#include <stddef.h>

namespace {

// debug instrumentation (REID_FUSED_DEBUG bit 13 = 8192): cycle counters summed over CTAs
__device__ unsigned long long g_dbg[8];

constexpr int NQ = 128;        // queries per block  (MMA N, TMEM columns)
constexpr int TMG = 128;       // gallery rows per tile (MMA M, TMEM lanes)
constexpr int BK = 64;         // K chunk: one 128-byte swizzle atom of fp16
constexpr int KL = REID_KLIST; // candidates are complete down to the KL-th best score of a chunk
constexpr int A_STAGE = TMG * BK * 2;   // 16 KB
constexpr int B_CHUNK = NQ * BK * 2;    // 16 KB
constexpr int MAX_STAGES = 8;
constexpr int EPI_WARP0 = 4;   // warps 4.. are the epilogue: EPI_WARPS/4 warps per TMEM lane quadrant
#ifndef REID_EPI_WARPS
#define REID_EPI_WARPS 16
#endif
constexpr int EPI_WARPS = REID_EPI_WARPS;
constexpr int UPD_PER_WARP = NQ / EPI_WARPS;           // queries whose candidate threshold a warp owns
static_assert(EPI_WARPS == 8 || EPI_WARPS == 16, "epilogue warps");
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int THREADS = EPI_WARP0 * 32 + EPI_THREADS;
constexpr int NBUF = 4;                  // TMEM accumulators: the epilogue may lag the MMA by up to 3 tiles
constexpr uint32_t TMEM_COLS = NBUF * NQ; // 4 x 128 fp32 columns = all of TMEM
constexpr int QCAP = 96;                 // per-warp hit queue entries (8 bytes each: score + packed meta)
constexpr int FLUSH_TILES = 128;         // 16-bit counters: <= 240 weighted increments per tile
#ifndef REID_SAMPLE_W
#define REID_SAMPLE_W 32
#endif
constexpr int SAMPLE_W = REID_SAMPLE_W;  // deep thresholds: rows with (row % SAMPLE_W) == 5, weight SAMPLE_W
static_assert(SAMPLE_W >= 16 && SAMPLE_W % 16 == 0, "sampled rows are column 5 of a 16-column step");
#ifndef REID_CALIB_ROWS
#define REID_CALIB_ROWS 2048
#endif
constexpr int CALIB_ROWS = REID_CALIB_ROWS;   // strided gallery sample of the calibration pre-pass
// queue meta word: bits 0-6 query column, 7 SAMPLED, 10-31 local gallery row (< 2^22)
constexpr uint32_t M_SAMPLED = 1u << 7;
constexpr int M_ROW_SHIFT = 10;
constexpr int64_t MAX_ROWS = 1ll << 22;
static_assert(KL == 32, "threshold update takes the 32nd largest of a 64-entry window");

struct Params {
  const int32_t* q_code; const int32_t* g_code; const int32_t* excl; int E;
  const float* pos_thr; const int32_t* n_pos;
  int64_t Q, G_local, g_offset;
  int Pmax, pcap, kchunks, stages, n_chunks, n_qblocks, cand_cap;
  int debug;       // REID_FUSED_DEBUG env: bit0 = no counting, bit1 = no hits at all (GEMM + scan only),
                   // bit2 = epilogue only hands the accumulator back (mainloop only),
                   // bit6 (64) = count every threshold exactly (no deep sampling)
  int64_t rows_per_chunk;
  int32_t* hist;   // [Q, Pmax] global bucket histogram (workspace)
  uint32_t* thr_share;   // [Q] best known candidate threshold per query (ordered key), shared by all chunks
  int32_t* n_exact;      // [Q] number of positive thresholds counted exactly (the rest on the row sample)
  int calib;             // 1 = calibration pre-pass: all thresholds exact, no candidates
  int64_t row_stride;    // gallery row stride of this pass (1, or the sample stride of the pre-pass)
  float* cand_score; int32_t* cand_idx; int32_t* cand_count;
};

// per-CTA shared state of the epilogue (one query block)
struct EpiState {
  float s_min[NQ];      // fast-path test of ordinary rows: min(candidate threshold, lowest EXACT positive threshold)
  float s_minS[NQ];     // fast-path test of sampled rows: also covers the deep thresholds
  float s_thrtop[NQ];   // candidate threshold: >= 32 appended rows lie above it (or -inf)
  float s_threx[NQ];    // lowest exactly-counted positive threshold (+inf when none)
  float s_thrlow[NQ];   // lowest positive threshold (+inf when the query has no positive)
  int s_qcode[NQ];
  int s_npos[NQ];
  int s_nexact[NQ];
  int s_candcnt[NQ];    // candidate slots allocated so far (may exceed cand_cap: overflow is flagged by the re-scorer)
  int s_nextupd[NQ];    // append count at which the candidate threshold is next refreshed
  int s_hasexcl[NQ];
};

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory"); }

// order-preserving integer key of a float (a single redux.sync.min replaces a shuffle reduction)
__device__ __forceinline__ uint32_t key32(float x) { const uint32_t b = __float_as_uint(x); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
__device__ __forceinline__ float unkey32(uint32_t k) {
  if (k <= 0x007FFFFFu) return -INFINITY;              // includes zero-initialised memory
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

// 32-bit shared-window addresses of the per-CTA epilogue state.  The out-of-line drain / refresh
// routines access shared memory through explicit ld/st/atom.shared PTX: one copy of the code (the
// kernel must stay inside the 32 KB instruction cache) without degrading to generic loads.
struct EpiAddr {
  uint32_t es;     // EpiState
  uint32_t thr;    // float [NQ][pcap] positive thresholds, sorted descending
  uint32_t hist;   // u32   [NQ][pcap/2] packed 16-bit bucket counters
  uint32_t qs;     // float [EPI_WARPS][QCAP] hit queue: score
  uint32_t qm;     // u32   [EPI_WARPS][QCAP] hit queue: packed meta (query column, flags, local gallery row)
};
#define ES_OFF(field) ((uint32_t)offsetof(EpiState, field))

__device__ __forceinline__ float lds_f32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ int lds_s32(uint32_t a) { int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_s32(uint32_t a, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ int atoms_add(uint32_t a, int v) { int o; asm volatile("atom.shared.add.s32 %0, [%1], %2;" : "=r"(o) : "r"(a), "r"(v) : "memory"); return o; }
__device__ __forceinline__ void reds_add(uint32_t a, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t atoms_exch(uint32_t a, uint32_t v) { uint32_t o; asm volatile("atom.shared.exch.b32 %0, [%1], %2;" : "=r"(o) : "r"(a), "r"(v) : "memory"); return o; }

// Drain queue entries [first, first + n) (n <= 32) of epilogue warp `ew`, one entry per lane.
__device__ __noinline__ void epi_drain32(EpiAddr A, const Params* pp, int ew, int first, int n, int lane, int64_t q0, int chunk) {
  const Params& p = *pp;
  __syncwarp();
  bool have = lane < n;
  const uint32_t qoff = (uint32_t)(ew * QCAP + first + lane) * 4u;
  const float s = have ? lds_f32(A.qs + qoff) : 0.f;
  const uint32_t meta = have ? lds_u32(A.qm + qoff) : 0u;
  const int row = (int)(meta >> M_ROW_SHIFT);
  have = have && row < p.G_local;                              // rows past the shard end (TMA zero fill) are no rows
  const uint32_t ql4 = (meta & 127u) * 4u;
  const int ql = meta & 127;
  bool ok = have;
  if (have && lds_s32(A.es + ES_OFF(s_hasexcl) + ql4)) {        // same-image mask (eval_mm_protocol.py:408-418)
    const int32_t gidx = (int32_t)(p.g_offset + (int64_t)row * p.row_stride);
#pragma unroll 1
    for (int x = 0; x < p.E; ++x) ok = ok && (p.excl[(q0 + ql) * p.E + x] != gidx);
  }
  // (a) bucket among the query's positive thresholds: b = #{j : t_j >= s}.  Rows that are positives of
  //     the query are skipped (positives are ordered exactly among themselves: rank_j = 1 + above_j + j).
  //     Ordinary rows only see the exactly-counted thresholds [0, n_exact); sampled rows see all of them
  //     and stand for SAMPLE_W rows in the deep buckets.
  if (ok) {
    const int ne = lds_s32(A.es + ES_OFF(s_nexact) + ql4);
    int hi = ((meta & M_SAMPLED) ? lds_s32(A.es + ES_OFF(s_npos) + ql4) : ne) - 1;
    const uint32_t t = A.thr + (uint32_t)(ql * p.pcap) * 4u;
    if (hi >= 0 && s > lds_f32(t + hi * 4) &&                   // invariant: t[hi] < s
        p.g_code[(int64_t)row * p.row_stride] != lds_s32(A.es + ES_OFF(s_qcode) + ql4)) {
      int lo = 0;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (lds_f32(t + mid * 4) < s) hi = mid; else lo = mid + 1;
      }
      const uint32_t w = (lo < ne) ? 1u : (uint32_t)SAMPLE_W;
      reds_add(A.hist + (uint32_t)((ql * p.pcap + lo) >> 1) * 4u, w << ((lo & 1) * 16));   // pcap is even
    }
  }
  // (b) candidates: lane-parallel append; the threshold is refreshed at tile boundaries
  if (ok && s > lds_f32(A.es + ES_OFF(s_thrtop) + ql4)) {
    const int slot = atoms_add(A.es + ES_OFF(s_candcnt) + ql4, 1);
    if (slot < p.cand_cap) {
      const int64_t o = ((q0 + ql) * p.n_chunks + chunk) * (int64_t)p.cand_cap + slot;
      p.cand_score[o] = s;
      p.cand_idx[o] = row;
    }
  }
  __syncwarp();
}

// Candidate-threshold refresh for query qq (whole warp): 32nd largest of the last 64 appended scores.
__device__ __noinline__ void epi_refresh_thr(EpiAddr A, const Params* pp, int qq, int cnt, int lane, int64_t q0, int chunk) {
  const Params& p = *pp;
  const float* win = p.cand_score + ((q0 + qq) * p.n_chunks + chunk) * (int64_t)p.cand_cap + (cnt - 64);
  const float a0 = __ldcg(win + lane), a1 = __ldcg(win + 32 + lane);
  int r0 = 0, r1 = 0;                                          // number of window values greater than a0 / a1
#pragma unroll 4
  for (int k = 0; k < 32; ++k) {
    const float b0 = __shfl_sync(0xffffffffu, a0, k), b1 = __shfl_sync(0xffffffffu, a1, k);
    r0 += (b0 > a0) + (b1 > a0);
    r1 += (b0 > a1) + (b1 > a1);
  }
  // the values with fewer than 32 greater ones are >= 32 scores; their minimum bounds the 32nd largest
  const float cand = fminf(r0 < 32 ? a0 : INFINITY, r1 < 32 ? a1 : INFINITY);
  const float nthr = unkey32(__reduce_min_sync(0xffffffffu, key32(cand)));
  if (lane == 0) {
    const uint32_t q4 = (uint32_t)qq * 4u;
    sts_s32(A.es + ES_OFF(s_nextupd) + q4, cnt + 16);
    if (nthr > lds_f32(A.es + ES_OFF(s_thrtop) + q4)) {
      sts_f32(A.es + ES_OFF(s_thrtop) + q4, nthr);
      const float ma = fminf(nthr, lds_f32(A.es + ES_OFF(s_threx) + q4));
      sts_f32(A.es + ES_OFF(s_min) + q4, ma);
      sts_f32(A.es + ES_OFF(s_minS) + q4, fminf(ma, lds_f32(A.es + ES_OFF(s_thrlow) + q4)));
    }
  }
  __syncwarp();
}

// spill the packed 16-bit counters of this CTA into the global histogram
__device__ __noinline__ void epi_flush_hist(EpiAddr A, const Params* pp, int et, int64_t q0) {
  const Params& p = *pp;
  const int words = NQ * p.pcap / 2;
#pragma unroll 1
  for (int i = et; i < words; i += EPI_THREADS) {
    if (lds_u32(A.hist + i * 4) == 0) continue;
    const uint32_t w = atoms_exch(A.hist + i * 4, 0u);         // other warps keep adding concurrently
    if (w) {
      const int ql = (2 * i) / p.pcap, b = (2 * i) % p.pcap;
      const int64_t q = q0 + ql;
      if (q < p.Q) {
        if ((w & 0xFFFFu) && b < p.Pmax) atomicAdd(&p.hist[q * p.Pmax + b], (int)(w & 0xFFFFu));
        if ((w >> 16) && b + 1 < p.Pmax) atomicAdd(&p.hist[q * p.Pmax + b + 1], (int)(w >> 16));
      }
    }
  }
}

// calibration: thresholds whose estimated rank inside a chunk exceeds `limit` rows are counted on the
// row sample only.  hist holds bucket counts over the n_sample calibration rows; it is zeroed for the
// main pass.
__global__ void calib_split_kernel(int32_t* __restrict__ hist, const int32_t* __restrict__ n_pos, int64_t Q, int Pmax,
                                   float scale, float limit, int32_t* __restrict__ n_exact) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < Q; q += (int64_t)gridDim.x * blockDim.x) {
    const int np = min(n_pos[q], Pmax);
    int acc = 0, ne = 0;
    bool open = true;
    for (int j = 0; j < Pmax; ++j) {
      acc += hist[q * Pmax + j];
      hist[q * Pmax + j] = 0;
      if (j < np && open) {
        if ((float)acc * scale <= limit) ne = j + 1; else open = false;
      }
    }
    n_exact[q] = ne;
  }
}
__global__ void fill_n_exact_kernel(const int32_t* __restrict__ n_pos, int64_t Q, int Pmax, int32_t* __restrict__ n_exact) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < Q; q += (int64_t)gridDim.x * blockDim.x)
    n_exact[q] = min(n_pos[q], Pmax);
}

// pos_above[q, j] += sum_{b <= j} hist[q, b]
__global__ void hist_to_above_kernel(const int32_t* __restrict__ hist, const int32_t* __restrict__ n_pos, int64_t Q,
                                     int Pmax, int32_t* __restrict__ pos_above, const uint32_t* __restrict__ thr_share,
                                     float* __restrict__ cand_thr) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < Q; q += (int64_t)gridDim.x * blockDim.x) {
    if (cand_thr) cand_thr[q] = unkey32(thr_share[q]);
    const int np = min(n_pos[q], Pmax);
    int acc = 0;
    for (int j = 0; j < np; ++j) {
      acc += hist[q * Pmax + j];
      pos_above[q * Pmax + j] += acc;
    }
  }
}

// PAIR = false ("T" layout, one CTA): A = gallery tile (M = 128 rows -> TMEM lanes, streamed), B = 128 resident
//   queries (N = 128 -> TMEM columns).  An M128 x N128 SS-mode MMA is bound by the A-operand shared-memory
//   read (~146 cycles per instruction, 43% of the tensor peak: measured with gallery loads disabled).
// PAIR = true ("R" layout, cta_group::2 pair of a 2-CTA cluster): A = 256 resident queries (128 per CTA -> the
//   128 TMEM lanes of that CTA), B = 256 streamed gallery rows (each CTA loads half of every tile), one
//   M256 x N256 MMA issued by the leader: ~175 cycles for 4x the math (73% of the tensor peak = the cuBLAS burst
//   rate), half the L2->SM operand traffic per flop.  Lane = query, column = gallery row: the per-query
//   thresholds sit in registers, the sampled rows are compile-time columns.  Hit queue / drain are shared.
template <bool PAIR>
__global__ void __launch_bounds__(THREADS, 1)
retrieve_fused_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmQ,
                      const __grid_constant__ Params prm) {
  constexpr int TROWS = PAIR ? 256 : TMG;              // gallery rows per tile step
  constexpr int DCOLS = PAIR ? 256 : NQ;               // accumulator columns per TMEM buffer
  constexpr int NB = PAIR ? 2 : NBUF;                  // TMEM accumulators
  constexpr int QSTEP = PAIR ? 2 * NQ : NQ;            // queries per work item
  const uint32_t crank = PAIR ? tc::cluster_ctarank() : 0u;
  const bool leader = crank == 0;
  const int unit = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;      // scheduling unit: CTA or CTA pair
  const int n_units = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  const Params& p = prm;
  uint8_t* sB = smem;                                         // [kchunks][B_CHUNK] resident query block
  uint8_t* sA = sB + p.kchunks * B_CHUNK;                      // [stages][A_STAGE]  gallery ring
  float* s_thr = reinterpret_cast<float*>(sA + p.stages * A_STAGE);              // [NQ][pcap]
  uint32_t* s_hist32 = reinterpret_cast<uint32_t*>(s_thr + NQ * p.pcap);         // [NQ][pcap/2]
  float* s_qs = reinterpret_cast<float*>(s_hist32 + NQ * p.pcap / 2);            // [EPI_WARPS][QCAP]
  uint32_t* s_qm = reinterpret_cast<uint32_t*>(s_qs + EPI_WARPS * QCAP);         // [EPI_WARPS][QCAP]
  EpiState* es = reinterpret_cast<EpiState*>(s_qm + EPI_WARPS * QCAP);
  __shared__ __align__(8) uint64_t full[MAX_STAGES], empty[MAX_STAGES], bfull, bempty, tfull[NBUF], tempty[NBUF];   // (pair mode uses 2 of the NBUF)
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    tc::mbar_init(&bfull, 1); tc::mbar_init(&bempty, 1);
    // in a pair the leader's tempty collects the epilogue warps of BOTH CTAs
    for (int b = 0; b < NBUF; ++b) { tc::mbar_init(&tfull[b], 1); tc::mbar_init(&tempty[b], PAIR ? 2 * EPI_WARPS : EPI_WARPS); }
    tc::fence_barrier_init();
    tc::prefetch_tensormap(&tmG); tc::prefetch_tensormap(&tmQ);
  }
  if (warp == 2) { if (PAIR) tc::tmem_alloc_pair(&tmem_base_s, TMEM_COLS); else tc::tmem_alloc(&tmem_base_s, TMEM_COLS); }
  tc::fence_before_sync();
  if (PAIR) tc::cluster_sync_all(); else __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  const int n_items = p.n_qblocks * p.n_chunks;     // n_qblocks counts blocks of QSTEP queries

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------------ TMA producer
    // sB holds the RESIDENT query block of this CTA (128 queries; the MMA's B operand in T layout, A in R);
    // sA is the gallery ring (128 rows per stage per CTA; in pair mode the two CTAs load the two halves of a tile)
    uint32_t it = 0, ph = 0;
    int st = 0;
    const uint32_t bfull_l = tc::leader_addr(&bfull);
    for (int item = unit; item < n_items; item += n_units, ++it) {
      const int chunk = item / p.n_qblocks, qb = item % p.n_qblocks;
      const int64_t row0 = chunk * p.rows_per_chunk;
      const int64_t row1 = reid_min64(p.G_local, row0 + p.rows_per_chunk);
      const int ntiles = (int)((row1 - row0 + TROWS - 1) / TROWS);
      const int qrow = qb * QSTEP + (int)crank * NQ;
      tc::mbar_wait(&bempty, (it & 1) ^ 1);            // previous item's MMAs have finished with the queries
      if (leader) tc::mbar_arrive_expect_tx(&bfull, (uint32_t)(p.kchunks * B_CHUNK) * (PAIR ? 2u : 1u));
#pragma unroll 1
      for (int kc = 0; kc < p.kchunks; ++kc) {
        if (PAIR) tc::tma_load_2d_pair(sB + kc * B_CHUNK, &tmQ, bfull_l, kc * BK, qrow);
        else tc::tma_load_2d(sB + kc * B_CHUNK, &tmQ, &bfull, kc * BK, qrow);
      }
#pragma unroll 1
      for (int t = 0; t < ntiles; ++t) {
        const int grow = (int)(row0 + (int64_t)t * TROWS) + (int)crank * TMG;
#pragma unroll 1
        for (int kc = 0; kc < p.kchunks && !(p.debug & 128); ++kc) {     // debug 128: no gallery loads at all
          tc::mbar_wait(&empty[st], ph ^ 1);
          if (leader) tc::mbar_arrive_expect_tx(&full[st], A_STAGE * (PAIR ? 2u : 1u));
          if (PAIR) tc::tma_load_2d_pair(sA + st * A_STAGE, &tmG, tc::leader_addr(&full[st]), kc * BK, grow);
          else tc::tma_load_2d(sA + st * A_STAGE, &tmG, &full[st], kc * BK, grow);
          if (++st == p.stages) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    // ------------------------------------------------------------------ MMA issuer (pair: leader CTA only)
    // T: D[gallery 128 x queries 128] = A(gallery stage) . B(queries)^T ;  R: D[queries 256 x gallery 256] = A(queries) . B(gallery)^T
    constexpr uint32_t idesc_n = PAIR ? tc::make_idesc_f16(256, 256, 0) : tc::make_idesc_f16(TMG, NQ, 0);
    // debug 512 / 1024 (timing experiments in T layout, results invalid): issue N=256 / N=192 instructions
    const uint32_t idesc = (!PAIR && (p.debug & 512)) ? tc::make_idesc_f16(TMG, 256, 0)
                         : (!PAIR && (p.debug & 1024)) ? tc::make_idesc_f16(TMG, 192, 0) : idesc_n;
    uint32_t it = 0, ph = 0, tilecount = 0;
    int st = 0;
    for (int item = unit; item < n_items; item += n_units, ++it) {
      const int chunk = item / p.n_qblocks;
      const int64_t row0 = chunk * p.rows_per_chunk;
      const int64_t row1 = reid_min64(p.G_local, row0 + p.rows_per_chunk);
      const int ntiles = (int)((row1 - row0 + TROWS - 1) / TROWS);
      tc::mbar_wait(&bfull, it & 1);
      tc::fence_after_sync();
#pragma unroll 1
      for (int t = 0; t < ntiles; ++t, ++tilecount) {
        const uint32_t buf = tilecount % NB, bph = (tilecount / NB) & 1;
        const long long tw0 = (p.debug & 8192) ? clock64() : 0;
        tc::mbar_wait(&tempty[buf], bph ^ 1);          // epilogue has drained this accumulator
        tc::fence_after_sync();
        if (p.debug & 8192) { atomicAdd(&g_dbg[2], (unsigned long long)(clock64() - tw0)); atomicAdd(&g_dbg[5], 1ull); }
#pragma unroll 1
        for (int kc = 0; kc < p.kchunks; ++kc) {
          const long long tf0 = (p.debug & 8192) ? clock64() : 0;
          if (!(p.debug & 128)) tc::mbar_wait(&full[st], ph);
          tc::fence_after_sync();
          if (p.debug & 8192) atomicAdd(&g_dbg[3], (unsigned long long)(clock64() - tf0));
          const uint64_t gd = tc::make_smem_desc_sw128(tc::smem_u32(sA + st * A_STAGE));    // gallery stage
          const uint64_t qd = tc::make_smem_desc_sw128(tc::smem_u32(sB + kc * B_CHUNK));    // resident queries
#pragma unroll
          for (int k = 0; k < ((p.debug & 256) ? 0 : BK / 16); ++k) {   // debug 256: loads only, no MMA
            if (PAIR) {
              tc::mma_f16_ss_pair(tmem_base + buf * DCOLS, tc::advance_desc_k(qd, k), tc::advance_desc_k(gd, k), idesc, (kc | k) != 0);
            } else {
              const uint32_t dcol = (p.debug & (512 | 1024)) ? (buf & 1) * 256 : buf * DCOLS;   // wide-N experiment stays inside TMEM
              tc::mma_f16_ss(tmem_base + dcol, tc::advance_desc_k(gd, k), tc::advance_desc_k(qd, k), idesc, (kc | k) != 0);
            }
          }
          if (!(p.debug & 128)) { if (PAIR) tc::mma_commit_pair(&empty[st]); else tc::mma_commit(&empty[st]); }   // frees the gallery stage
          if (++st == p.stages) { st = 0; ph ^= 1; }
        }
        if (PAIR) tc::mma_commit_pair(&tfull[buf]); else tc::mma_commit(&tfull[buf]);   // accumulator complete -> epilogue
      }
      if (PAIR) tc::mma_commit_pair(&bempty); else tc::mma_commit(&bempty);             // query block no longer read
    }
  } else if (warp >= EPI_WARP0) {
    // ------------------------------------------------------------------ epilogue (EPI_WARPS warps)
    const int quad = warp & 3;                          // TMEM lane quadrant of this warp
    const int ew = warp - EPI_WARP0;                    // 0..EPI_WARPS-1
    const int part = ew >> 2;                           // which slice of the accumulator columns
    const int et = threadIdx.x - EPI_WARP0 * 32;        // 0..EPI_THREADS-1
    const uint32_t tmem_q = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t lt_mask = (1u << lane) - 1u;
    // T layout: lane = gallery row; the rows of this lane belong to the 1/SAMPLE_W stratified sample or not
    const bool smp = ((quad * 32 + lane) & (SAMPLE_W - 1)) == 5;
    const uint32_t lane_bits = smp ? M_SAMPLED : 0u;
    float* my_qs = s_qs + ew * QCAP;
    uint32_t* my_qm = s_qm + ew * QCAP;
    EpiAddr A;
    A.es = tc::smem_u32(es); A.thr = tc::smem_u32(s_thr); A.hist = tc::smem_u32(s_hist32);
    A.qs = tc::smem_u32(s_qs); A.qm = tc::smem_u32(s_qm);
#pragma unroll 1
    for (int i = et; i < NQ * p.pcap / 2; i += EPI_THREADS) s_hist32[i] = 0;
    epi_bar();
    uint32_t tilecount = 0;
    for (int item = unit; item < n_items; item += n_units) {
      const int chunk = item / p.n_qblocks, qb = item % p.n_qblocks;
      const int64_t row0 = chunk * p.rows_per_chunk;
      const int64_t row1 = reid_min64(p.G_local, row0 + p.rows_per_chunk);
      const int ntiles = (int)((row1 - row0 + TROWS - 1) / TROWS);
      const int64_t q0 = (int64_t)qb * QSTEP + (int64_t)crank * NQ;   // first query of THIS CTA's block
      // ---- item setup: per-query state
      if (et < NQ) {
        const int64_t q = q0 + et;
        const bool live = q < p.Q;
        const int np = (live && !(p.debug & 1)) ? min(p.n_pos[q], p.Pmax) : 0;
        const int ne = p.calib ? np : (live ? min(p.n_exact[q], np) : 0);
        es->s_qcode[et] = live ? p.q_code[q] : -2;
        es->s_npos[et] = np;
        es->s_nexact[et] = ne;
        const float tex = ne > 0 ? p.pos_thr[q * p.Pmax + ne - 1] : INFINITY;
        const float tl = np > 0 ? p.pos_thr[q * p.Pmax + np - 1] : INFINITY;
        es->s_threx[et] = tex;
        es->s_thrlow[et] = tl;
        // candidates: warm start from the best threshold any earlier chunk of this query has published;
        // padded queries and the calibration pre-pass never append
        const float tt = (live && !p.calib && !(p.debug & 2)) ? unkey32(__ldcg(&p.thr_share[q])) : INFINITY;
        es->s_thrtop[et] = tt;
        es->s_min[et] = fminf(tt, tex);
        es->s_minS[et] = fminf(fminf(tt, tex), tl);
        es->s_candcnt[et] = 0;
        es->s_nextupd[et] = 64;
        int he = 0;
        if (live) {
#pragma unroll 1
          for (int e = 0; e < p.E; ++e) he |= (p.excl[q * p.E + e] >= 0);
        }
        es->s_hasexcl[et] = he;
      }
      if (!p.calib) {                                          // -inf pre-fill of this item's candidate score slots
        const float4 ninf = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll 1
        for (int ql = ew; ql < NQ; ql += EPI_WARPS) {
          if (q0 + ql >= p.Q) break;
          float4* dst = reinterpret_cast<float4*>(p.cand_score + ((q0 + ql) * p.n_chunks + chunk) * (int64_t)p.cand_cap);
#pragma unroll 1
          for (int j = lane; j < p.cand_cap / 4; j += 32) dst[j] = ninf;
        }
      }
#pragma unroll 1
      for (int ql = ew; ql < NQ; ql += EPI_WARPS) {          // one warp per query row of the threshold table
        const int64_t qq = q0 + ql;
#pragma unroll 1
        for (int j = lane; j < p.pcap; j += 32)
          s_thr[ql * p.pcap + j] = (qq < p.Q && j < p.Pmax) ? p.pos_thr[qq * p.Pmax + j] : -INFINITY;
      }
      epi_bar();
      int qn = 0;                                          // queued hits of this warp (persist across tiles)
      // ---- tiles
      for (int t = 0; t < ntiles; ++t, ++tilecount) {
        const uint32_t buf = tilecount % NB, bph = (tilecount / NB) & 1;
        const int tile_row0 = (int)(row0 + (int64_t)t * TROWS);
        // T: this lane's gallery row / R: this lane's query
        const int grow_local = tile_row0 + quad * 32 + lane;
        const bool valid = PAIR ? true : (grow_local < row1);
        const int myq = quad * 32 + lane;                               // R: query column id of this lane
        float minA = 0.f, minS = 0.f;
        if (PAIR) { minA = es->s_min[myq]; minS = es->s_minS[myq]; }   // refreshed thresholds, once per tile
        // fp16 images of the thresholds for the packed fast-path test (rounded DOWN after subtracting the rounding bound)
        const __half lowA = __float2half_rd(minA - 5e-4f), lowS = __float2half_rd(minS - 5e-4f);
        const __half2 hAA = __halves2half2(lowA, lowA), hAS = __halves2half2(lowA, lowS);
        const long long te0 = (p.debug & 8192) ? clock64() : 0;
        tc::mbar_wait(&tfull[buf], bph);
        tc::fence_after_sync();
        const long long te1 = (p.debug & 8192) ? clock64() : 0;
        constexpr int STEPS = (PAIR ? DCOLS : NQ) * 4 / EPI_WARPS / 16;
#pragma unroll 1
        for (int step = 0; step < ((p.debug & 4) ? 0 : STEPS); ++step) {
          const int c0 = part * (STEPS * 16) + step * 16;
          const bool step_sampled = ((c0 + 5) & (SAMPLE_W - 1)) == 5;     // uniform
          uint32_t r[16];
          tc::tmem_ld_x16(tmem_q + buf * DCOLS + c0, r);
          float mm[16];
          if (!PAIR) {
            const float* msrc = smp ? es->s_minS : es->s_min;      // lane-constant choice of threshold set
#pragma unroll
            for (int i4 = 0; i4 < 16; i4 += 4) {
              const float4 m4 = *reinterpret_cast<const float4*>(&msrc[c0 + i4]);
              mm[i4] = m4.x; mm[i4 + 1] = m4.y; mm[i4 + 2] = m4.z; mm[i4 + 3] = m4.w;
            }
          }
          tc::tmem_wait_ld();
          // fast path: a per-lane bit mask of the columns that MAY hit, ONE warp-wide OR per step.
          // R layout: the 16 scores are packed to half2 (one F2FP per pair) and compared pairwise against the
          // thresholds lowered by the fp16 rounding bound (5e-4 >= half an ulp of any |score| <= 1, rounded down), so
          // the mask is a superset of the exact hits (a few % more); the slow path re-tests exactly in fp32.
          // Mask bit p / 16 + p = column 2p / 2p + 1.
          unsigned colmask = 0, bits = 0;
          if (!(p.debug & 32)) {
            if (PAIR) {
#pragma unroll
              for (int pr = 0; pr < 8; ++pr) {
                const __half2 h = __floats2half2_rn(__uint_as_float(r[2 * pr]), __uint_as_float(r[2 * pr + 1]));
                // only column 5 of a 16-column step can be a sampled row (rows = 5 mod SAMPLE_W)
                const __half2 thr = (pr == 2 && step_sampled) ? hAS : hAA;
                bits |= __hgt2_mask(h, thr) & ((1u << pr) | (1u << (16 + pr)));
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const bool hit = valid && __uint_as_float(r[i]) > mm[i];
                bits |= hit ? (1u << (i >> 1) << ((i & 1) * 16)) : 0u;
              }
            }
            colmask = __reduce_or_sync(0xffffffffu, bits);
          }
          // slow path (single copy of the code: the kernel must fit the instruction cache).  Round n takes
          // the n-th flagged column of EVERY lane at once: per-lane select tree for the score, exact fp32 test, one
          // ballot + prefix popc to compact the round into the dense warp queue.  Rounds per step = max flagged
          // columns per lane (usually 1), not the number of hit columns.
          if (p.debug & 4096) colmask = 0;               // debug: fast path only (hits ignored)
          if (colmask) {
            unsigned b = bits;
#pragma unroll 1
            while (true) {
              const bool has = b != 0;
              if (!__any_sync(0xffffffffu, has)) break;
              const int bp = has ? (__ffs(b) - 1) : 0;
              const int i = ((bp & 15) << 1) | (bp >> 4);
              b &= b - 1;                                // (0 stays 0)
              // the lane's score of column i: 4-level select tree over the 16 registers (no memory)
              uint32_t t8[8], t4[4], t2[2];
#pragma unroll
              for (int u = 0; u < 8; ++u) t8[u] = (i & 1) ? r[2 * u + 1] : r[2 * u];
#pragma unroll
              for (int u = 0; u < 4; ++u) t4[u] = (i & 2) ? t8[2 * u + 1] : t8[2 * u];
#pragma unroll
              for (int u = 0; u < 2; ++u) t2[u] = (i & 4) ? t4[2 * u + 1] : t4[2 * u];
              const float sc = __uint_as_float((i & 8) ? t2[1] : t2[0]);
              const bool smp_col = PAIR && i == 5 && step_sampled;
              const bool hit = has && (!PAIR || sc > (smp_col ? minS : minA));     // exact test (T layout: already exact)
              const unsigned c = __ballot_sync(0xffffffffu, hit);
              if (hit) {
                const int pos = qn + __popc(c & lt_mask);
                my_qs[pos] = sc;
                my_qm[pos] = PAIR ? ((uint32_t)myq | (smp_col ? M_SAMPLED : 0u) |
                                     ((uint32_t)(tile_row0 + c0 + i) << M_ROW_SHIFT))
                                  : ((uint32_t)(c0 + i) | lane_bits | ((uint32_t)grow_local << M_ROW_SHIFT));
              }
              qn += __popc(c);
              if (qn > QCAP - 32) {                      // queue nearly full: drain the newest 32 now (order is irrelevant)
                qn -= 32;
                epi_drain32(A, &p, ew, qn, 32, lane, q0, chunk);
              }
            }
          }
        }
        tc::fence_before_sync();
        if (lane == 0) { if (PAIR && !leader) tc::mbar_arrive_remote(&tempty[buf], 0); else tc::mbar_arrive(&tempty[buf]); }
        if ((p.debug & 8192) && lane == 0) {
          const long long te2 = clock64();
          if (ew == 0) { atomicAdd(&g_dbg[0], (unsigned long long)(te2 - te1)); atomicAdd(&g_dbg[1], (unsigned long long)(te1 - te0)); }
        }
        // the accumulator is handed back; full batches are drained now, overlapping the next tile's MMA
        while (qn >= 32) { qn -= 32; epi_drain32(A, &p, ew, qn, 32, lane, q0, chunk); }
        // ---- tile boundary: this warp owns the candidate thresholds of a fixed group of queries
        //      (UPD_PER_WARP each).  The score slots of an item are pre-filled with -inf, so a window that
        //      contains a slot whose store is still in flight only yields a more conservative bound:
        //      no CTA barrier and no completion protocol are needed.
        if (!p.calib) {
          const int uq = ew * UPD_PER_WARP + (lane & (UPD_PER_WARP - 1));
          const int cnt = min(*(volatile int*)&es->s_candcnt[uq], p.cand_cap);
          unsigned um = __ballot_sync(0xffffffffu, lane < UPD_PER_WARP && cnt >= 64 && cnt >= es->s_nextupd[uq]);
          while (um) {
            const int src = __ffs(um) - 1;
            um &= um - 1;
            epi_refresh_thr(A, &p, ew * UPD_PER_WARP + src, __shfl_sync(0xffffffffu, cnt, src), lane, q0, chunk);
          }
        }
        if (((t + 1) % FLUSH_TILES) == 0) epi_flush_hist(A, &p, et, q0);
      }
      // ---- item end: drain the queue tail, spill the histogram, publish candidate state
      if (qn) epi_drain32(A, &p, ew, 0, qn, lane, q0, chunk);
      epi_bar();
      epi_flush_hist(A, &p, et, q0);
      if (et < NQ && q0 + et < p.Q && !p.calib) {
        p.cand_count[(q0 + et) * p.n_chunks + chunk] = es->s_candcnt[et];
        atomicMax(&p.thr_share[q0 + et], key32(es->s_thrtop[et]));
      }
      epi_bar();
    }
  }
  tc::fence_before_sync();
  if (PAIR) tc::cluster_sync_all(); else __syncthreads();   // the peer may still be signalled / read until here
  if (warp == 2) { if (PAIR) tc::tmem_dealloc_pair(tmem_base, TMEM_COLS); else tc::tmem_dealloc(tmem_base, TMEM_COLS); }
}

size_t fused_smem_bytes(int kchunks, int stages, int pcap, bool /*pair*/) {
  return (size_t)kchunks * B_CHUNK + (size_t)stages * A_STAGE + (size_t)NQ * pcap * 4 /*thr*/ +
         (size_t)NQ * pcap * 2 /*hist*/ + (size_t)EPI_WARPS * QCAP * 8 /*queues*/ +
         sizeof(EpiState) + 1024;
}

}  // namespace

// workspace = the global bucket histogram [Q, 64] int32 (Pmax <= 64)
// workspace = histogram [Q, Pmax<=64] + candidate threshold [Q] + exact-threshold count [Q]
extern "C" int reid_debug_counters(unsigned long long* out, int reset) {
  if (out && cudaMemcpyFromSymbol(out, g_dbg, sizeof(g_dbg)) != cudaSuccess) return REID_E_CUDA;
  if (reset) { unsigned long long z[8] = {0}; if (cudaMemcpyToSymbol(g_dbg, z, sizeof(z)) != cudaSuccess) return REID_E_CUDA; }
  return REID_OK;
}

extern "C" size_t reid_retrieve_fused_workspace_bytes(int64_t Q, int64_t, int) { return (size_t)Q * 66 * sizeof(int32_t); }

extern "C" int reid_retrieve_fused(const void* q_f16, const void* g_f16, const int32_t* q_code, const int32_t* g_code,
                                   const int32_t* excl, int E, const float* pos_thr, const int32_t* n_pos, int64_t Q,
                                   int64_t G_local, int64_t g_offset, int d, int Pmax, int n_chunks, int total_chunks, int cand_cap,
                                   int32_t* pos_above, float* cand_score, int32_t* cand_idx, int32_t* cand_count,
                                   float* cand_thr, void* workspace, size_t workspace_bytes, void* stream) {
  if (!q_f16 || !g_f16 || !q_code || !g_code || !pos_thr || !n_pos || !pos_above || !cand_score || !cand_idx ||
      !cand_count || Q <= 0 || G_local <= 0 || n_chunks <= 0 || cand_cap < 64 || cand_cap % 4 != 0 || (E > 0 && !excl) || E < 0)
    return REID_E_INVALID;
  if (d % BK != 0 || d > 512 || Pmax <= 0 || Pmax > 64 || G_local > MAX_ROWS) return REID_E_UNSUPPORTED;
  // cta_group::2 pair variant (R layout, M256 x N256): REID_FUSED_PAIR=0 selects the single-CTA T layout
  const char* pair_env = getenv("REID_FUSED_PAIR");
  const bool pair = pair_env ? atoi(pair_env) != 0 : true;
  const int n_lchunks = n_chunks;
  const int trows = pair ? 256 : TMG;
  const int qstep = pair ? 2 * NQ : NQ;
  Params p;
  p.q_code = q_code; p.g_code = g_code; p.excl = excl; p.E = E; p.pos_thr = pos_thr; p.n_pos = n_pos;
  p.Q = Q; p.G_local = G_local; p.g_offset = g_offset;
  if (!workspace || workspace_bytes < (size_t)Q * (Pmax + 2) * sizeof(int32_t)) return REID_E_WORKSPACE;
  p.Pmax = Pmax; p.pcap = (Pmax + 3) / 4 * 4; p.kchunks = d / BK;
  p.n_chunks = n_chunks; p.n_qblocks = (int)((Q + qstep - 1) / qstep); p.cand_cap = cand_cap;
  const int64_t rpc = (G_local + n_lchunks - 1) / n_lchunks;
  p.rows_per_chunk = (rpc + trows - 1) / trows * trows;
  p.hist = (int32_t*)workspace; p.thr_share = (uint32_t*)workspace + (size_t)Q * Pmax;
  p.n_exact = (int32_t*)workspace + (size_t)Q * (Pmax + 1); p.calib = 0; p.row_stride = 1; p.cand_score = cand_score;
  { const char* dbg = getenv("REID_FUSED_DEBUG"); p.debug = dbg ? atoi(dbg) : 0; }
  p.cand_idx = cand_idx; p.cand_count = cand_count;
  int stages = MAX_STAGES;
  const size_t smem_max = 227 * 1024;
  while (stages > 2 && fused_smem_bytes(p.kchunks, stages, p.pcap, pair) > smem_max) --stages;
  if (fused_smem_bytes(p.kchunks, stages, p.pcap, pair) > smem_max) return REID_E_UNSUPPORTED;
  p.stages = stages;
  const size_t smem = fused_smem_bytes(p.kchunks, stages, p.pcap, pair);
  CUtensorMap tmG, tmQ;
  if (!tc_host::make_map_f16(&tmG, g_f16, G_local, d, TMG) ||
      !tc_host::make_map_f16(&tmQ, q_f16, Q, d, NQ))
    return REID_E_CUDA;
  auto kern1 = retrieve_fused_kernel<false>;
  auto kern2 = retrieve_fused_kernel<true>;
  if (cudaFuncSetAttribute(pair ? kern2 : kern1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return REID_E_CUDA;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return REID_E_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  // one persistent CTA per SM (pair: one CTA pair per two SMs, launched as clusters of 2)
  auto launch = [&](const CUtensorMap& mg, const Params& pp) -> bool {
    const int n_items = pp.n_qblocks * pp.n_chunks;
    if (!pair) {
      const int grid = n_items < sms ? n_items : sms;
      kern1<<<grid, THREADS, smem, st>>>(mg, tmQ, pp);
      return cudaGetLastError() == cudaSuccess;
    }
    const int units = n_items < sms / 2 ? n_items : sms / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * units); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern2, mg, tmQ, pp) == cudaSuccess;
  };
  if (cudaMemsetAsync(workspace, 0, (size_t)Q * (Pmax + 2) * sizeof(int32_t), st) != cudaSuccess) return REID_E_CUDA;
  const int aux_grid = (int)reid_min64((Q + 255) / 256, 148 * 8);
  const bool sample_deep = (G_local >= 16 * CALIB_ROWS) && !(p.debug & 64);
  if (sample_deep) {
    // calibration pre-pass: the same kernel over a strided sample of CALIB_ROWS gallery rows, every
    // threshold counted exactly, no candidates -> per-query bucket histogram of the sample
    Params c = p;
    c.calib = 1;
    c.row_stride = G_local / CALIB_ROWS;
    c.G_local = CALIB_ROWS;
    c.n_chunks = 1;
    c.rows_per_chunk = CALIB_ROWS;
    CUtensorMap tmS;
    {
      tc_host::EncodeTiledFn enc = tc_host::get_encode();
      if (!enc) return REID_E_CUDA;
      cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)CALIB_ROWS};
      cuuint64_t strides[1] = {(cuuint64_t)d * 2 * (cuuint64_t)c.row_stride};
      cuuint32_t box[2] = {64, (cuuint32_t)TMG};
      cuuint32_t estr[2] = {1, 1};
      if (enc(&tmS, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(g_f16), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return REID_E_CUDA;
    }
    if (!launch(tmS, c)) return REID_E_CUDA;
    // deep = the row sample is expected to hold >= 32 rows above the threshold OVER ALL CHUNKS OF ALL RANKS
    // (17.7% rank error at the boundary); never classify on fewer than 8 calibration hits
    const float scale = (float)p.rows_per_chunk / (float)CALIB_ROWS;
    // (budget = 32 SAMPLE_W rows GALLERY-WIDE: a per-chunk budget would count 8 x 1024 rows exactly when a 100k-row
    //  gallery is cut into 8 chunks -- 8 % of all rows as hits)
    const int tc_all = total_chunks > n_chunks ? total_chunks : n_chunks;
    const float limit = fmaxf(32.f * (float)SAMPLE_W / (float)tc_all, 8.f * scale);
    calib_split_kernel<<<aux_grid, 256, 0, st>>>(p.hist, n_pos, Q, Pmax, scale, limit, p.n_exact);
  } else {
    fill_n_exact_kernel<<<aux_grid, 256, 0, st>>>(n_pos, Q, Pmax, p.n_exact);
  }
  REID_CHECK_LAUNCH();
  if (!launch(tmG, p)) return REID_E_CUDA;
  hist_to_above_kernel<<<aux_grid, 256, 0, st>>>(p.hist, n_pos, Q, Pmax, pos_above, p.thr_share, cand_thr);
  REID_CHECK_LAUNCH();
  return REID_OK;
}
