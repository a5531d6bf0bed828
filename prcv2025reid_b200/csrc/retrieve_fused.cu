// retrieve_fused.cu -- K3+K4: similarity GEMM with the ranking statistics computed in its epilogue.
//
// Replaces the per-query loop body of rank_and_metrics, eval_mm_protocol.py:401-455 (cosine_sim,
// same-image mask, argsort, CMC, AP walk) for one gallery shard, without materialising S or a sort.
//
// Layout: a cta_group::2 pair of a 2-CTA cluster per work item = (256 queries, gallery chunk); each CTA keeps its
// 128 queries resident in shared memory (A operand, 128 KB) and streams its half of every 256-row gallery tile
// through a TMA ring; the leader issues M256 x N256 MMAs into one of two 256-column TMEM accumulators.
// TMEM lane = query, column = gallery row: the per-query thresholds sit in registers, the rows of the row samples
// are compile-time columns.
// Epilogue (16 warps, four per TMEM lane quadrant, tcgen05.ld 32x32b.x16): every score is compared with
// min(candidate threshold, lowest exactly-counted positive threshold) of its query; ONE redux.or per 16-column
// step tells whether the warp has a hit at all.  Hits are compacted round by round (n-th hit of every lane:
// register select tree + ballot / popc) into a per-warp queue in shared memory and drained 32 at a time,
// lane-parallel, after the accumulator has been handed back:
//       (a) counting: a binary search over the query's positive thresholds (sorted descending, in
//           shared memory) gives the bucket b = #thresholds >= score; hist[q][b]++ (packed 16-bit
//           counters in shared memory, spilled to a global histogram every 128 tiles); the count of
//           rows ranked above positive j is the weighted prefix sum over buckets <= j.
//           Counting classes (calibration pre-pass over a strided CALIB_ROWS-row sample estimates every
//           threshold's rank inside a chunk): thresholds [0, n_exact) are counted on EVERY row; thresholds
//           [n_exact, n_l1) -- estimated in-chunk rank above max(32 * W1 / total_chunks, 8 calibration hits) --
//           on the level-1 row sample (local row % W1 == 5, weight W1 = 32); thresholds [n_l1, n_pos) -- rank above
//           32 * W2 / total_chunks -- on the level-2 sample (local row % W2 == 5, weight W2 = 1024).  A sampled
//           count rests on >= 32 sampled rows gallery-wide (<= 18 % standard error on a rank > 1000 resp. > 32768,
//           which moves a query's AP by ~1e-5 and mAP by ~1e-6); REID_FUSED_EXACT_COUNTS counts everything exactly.
//       (b) candidates: rows above the query's running threshold are appended (lane-parallel) to its
//           candidate buffer in global memory (slots pre-filled with -inf); at tile boundaries the owning warp
//           raises the threshold to the 32nd largest of the last 64 appended scores, so >= 32 appended rows
//           always lie above it.  The threshold reached by one gallery chunk is published (atomicMax) and
//           warm-starts the later chunks of the same query.
//     The epilogue has no CTA barrier inside an item: every shared structure is atomic-safe.
// Roofline: tensor cores, 2*Q*G*d flop; algorithmic HBM bytes are only operands + outputs.
//
// Build-time switches (all off in the shipped library; nothing in this file reads the environment):
//   -DREID_DEBUG=bits       timing experiments (bits below; results may be INVALID); 8192 adds reid_debug_counters
//   -DREID_PF_DIST=n        L2 prefetch of the gallery tile n tiles ahead, one prefetching pair per tile
//   -DREID_SCAN_FIRST=0/1   an accumulator that is already waiting is scanned before queued hits are drained
//   -DREID_SLOW_X1=0/1      hit columns re-read from TMEM one by one (1) or selected from the step's registers (0)
//   -DREID_HIST_GLOBAL=0/1  bucket counters as red.global on the workspace histogram (1) or packed 16-bit counters in shared memory (0)
//   -DREID_MAX_STAGES=n     cap on the gallery ring depth
//   -DREID_QCAP=n           per-warp hit queue entries
#include "common.cuh"
#include "tc_common.cuh"
#include <stddef.h>

namespace {

#ifndef REID_DEBUG
#define REID_DEBUG 0
#endif
#if REID_DEBUG
__device__ unsigned long long g_dbg[16];
#endif
#define FDBG(bits) ((REID_DEBUG) & (bits))      // compile-time: an experiment build has exactly the code it measures
// debug bits: 1 = no counting, 2 = no candidates, 4 = epilogue only hands the accumulator back (mainloop only),
// 32 = no fast-path test, 128 = no gallery loads, 256 = loads only (no MMA), 4096 = scan only (hits ignored),
// 8192 = cycle counters

constexpr int NQ = 128;        // queries per CTA (TMEM lanes); a pair works on 2 * NQ
constexpr int TMG = 128;       // gallery rows per CTA and tile (TMA box rows)
constexpr int TROWS = 2 * TMG; // gallery rows per tile (MMA N, accumulator columns)
constexpr int BK = 64;         // K chunk: one 128-byte swizzle atom of fp16
constexpr int KL = REID_KLIST; // candidates are complete down to the KL-th best score of a chunk
constexpr int A_STAGE = TMG * BK * 2;   // 16 KB: this CTA's half of a gallery tile, one K chunk
constexpr int B_CHUNK = NQ * BK * 2;    // 16 KB: one K chunk of the resident queries
constexpr int MAX_STAGES = 8;
constexpr int EPI_WARP0 = 4;   // warps 4.. are the epilogue: EPI_WARPS/4 warps per TMEM lane quadrant
constexpr int EPI_WARPS = 16;
constexpr int UPD_PER_WARP = NQ / EPI_WARPS;           // queries whose candidate threshold a warp owns
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int THREADS = EPI_WARP0 * 32 + EPI_THREADS;
constexpr int NB = 2;                    // TMEM accumulators of TROWS fp32 columns
constexpr uint32_t TMEM_COLS = NB * TROWS; // = 512: all of TMEM
#ifndef REID_QCAP
#define REID_QCAP 96
#endif
constexpr int QCAP = REID_QCAP;          // per-warp hit queue entries (8 bytes each: score + packed meta)
constexpr int FLUSH_TILES = 128;         // 16-bit counters: <= 256 increments per bucket and tile
constexpr int W1 = 32;                   // level-1 row sample: rows with (row % W1) == 5, weight W1
constexpr int W2 = 1024;                 // level-2 row sample: rows with (row % W2) == 5, weight W2
#ifndef REID_CALIB_ROWS
#define REID_CALIB_ROWS 8192
#endif
constexpr int CALIB_ROWS = REID_CALIB_ROWS;   // strided gallery sample of the calibration pre-pass: at most this many rows
constexpr int CALIB_MIN = 2048;               // ... at least this many (shards below 16 * CALIB_MIN rows are never sampled)
#ifndef REID_PF_DIST
#define REID_PF_DIST 0
#endif
constexpr int PF_DIST = REID_PF_DIST;
#ifndef REID_SCAN_FIRST
#define REID_SCAN_FIRST 0
#endif
#ifndef REID_SLOW_X1
#define REID_SLOW_X1 0
#endif
#ifndef REID_HIST_GLOBAL
#define REID_HIST_GLOBAL 0
#endif
#ifndef REID_MAX_STAGES
#define REID_MAX_STAGES 8
#endif
// queue meta word: bits 0-6 query column, 7 level-1 sampled row, 8 level-2 sampled row, 10-31 local gallery row (< 2^22)
constexpr uint32_t M_S1 = 1u << 7;
constexpr uint32_t M_S2 = 1u << 8;
constexpr int M_ROW_SHIFT = 10;
constexpr int64_t MAX_ROWS = 1ll << 22;
static_assert(KL == 32, "threshold update takes the 32nd largest of a 64-entry window");
static_assert(W2 % TROWS == 0 && TROWS % W1 == 0, "sampled rows are fixed columns of a tile");

struct Params {
  const int32_t* q_code; const int32_t* g_code; const int32_t* excl; int E;
  const float* pos_thr; const int32_t* n_pos;
  int64_t Q, G_local, g_offset;
  int Pmax, pos_stride, pcap, kchunks, stages, n_chunks, n_qblocks, n_full, cand_cap;
  int debug;             // FDBG bits (always 0 unless built with -DREID_DEBUG)
  int no_cand;           // REID_FUSED_NO_CANDIDATES: counting only
  int klist;             // candidates are complete down to the klist-th best score of a slot: 32, or 16 (REID_FUSED_KLIST16)
  int64_t rows_per_chunk;
  int32_t* hist;         // [Q, Pmax] global bucket histogram (workspace), raw hit counts
  uint32_t* thr_share;   // [Q] best known candidate threshold per query (ordered key), shared by all chunks
  int32_t* n_exact;      // [Q] thresholds [0, n_exact) are counted on every row
  int32_t* n_l1;         // [Q] thresholds [n_exact, n_l1) on the level-1 row sample, [n_l1, n_pos) on the level-2 sample
  int calib;             // 1 = calibration pre-pass: all thresholds exact, no candidates
  int calib_cap;         // calibration: sample rows above a threshold that put its estimated rank beyond the level-2 limit
  int calib_floor;       // ... a threshold is never retired on fewer rows than this (>= klist)
  float calib_rate;      // calib_cap / (rows of the sample)
  int64_t row_stride;    // gallery row stride of this pass (1, or the sample stride of the pre-pass)
  float* cand_score; int32_t* cand_idx; int32_t* cand_count;
};

// per-CTA shared state of the epilogue (one query block)
struct EpiState {
  float s_min[NQ];      // fast-path test of ordinary rows: min(candidate threshold, lowest EXACT positive threshold)
  float s_minS1[NQ];    // ... of level-1 sampled rows: also covers thresholds [n_exact, n_l1)
  float s_minS2[NQ];    // ... of level-2 sampled rows: covers every threshold
  float s_thrtop[NQ];   // candidate threshold: >= 32 appended rows lie above it (or -inf)
  float s_threx[NQ];    // lowest exactly-counted positive threshold (+inf when none)
  float s_thrl1[NQ];    // lowest threshold counted on the level-1 sample or exactly (+inf when none)
  float s_thrlow[NQ];   // lowest positive threshold (+inf when the query has no positive)
  int s_qcode[NQ];
  uint32_t s_cnts[NQ];  // n_pos | n_exact << 8 | n_l1 << 16 | has_excl << 24
  int s_candcnt[NQ];    // candidate slots allocated so far (may exceed cand_cap: overflow is flagged by the re-scorer)
  int s_nextupd[NQ];    // append count at which the candidate threshold is next refreshed
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

// non-blocking probe of an mbarrier phase
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}\n" : "=r"(ok) : "r"(tc::smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Work items.  Query blocks [0, n_full) -- a whole number of waves of the persistent pairs -- are processed against the
// WHOLE shard (one candidate slot per query, one cold start of the candidate threshold); the remaining blocks are cut
// into n_chunks gallery chunks each so that the last wave is full as well (items are handed out round-robin).
struct Item { int qb, slot; int64_t row0, row1; };
__device__ __forceinline__ Item get_item(const Params& p, int item) {
  Item it;
  if (item < p.n_full) { it.qb = item; it.slot = 0; it.row0 = 0; it.row1 = p.G_local; return it; }
  const int r = item - p.n_full;
  it.qb = p.n_full + r / p.n_chunks; it.slot = r % p.n_chunks;
  it.row0 = it.slot * p.rows_per_chunk;
  it.row1 = reid_min64(p.G_local, it.row0 + p.rows_per_chunk);
  return it;
}
__device__ __forceinline__ int item_tiles(const Item& it) { return it.row1 > it.row0 ? (int)((it.row1 - it.row0 + TROWS - 1) / TROWS) : 0; }

// L2 prefetch of one box of a tiled tensor map
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1) : "memory");
}

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
  const uint32_t cnts = have ? lds_u32(A.es + ES_OFF(s_cnts) + ql4) : 0u;
  bool ok = have;
  if (have && (cnts >> 24)) {                                  // same-image mask (eval_mm_protocol.py:408-418)
    const int32_t gidx = (int32_t)(p.g_offset + (int64_t)row * p.row_stride);
#pragma unroll 1
    for (int x = 0; x < p.E; ++x) ok = ok && (p.excl[(q0 + ql) * p.E + x] != gidx);
  }
  // (a) bucket among the query's positive thresholds: b = #{j : t_j >= s}.  Rows that are positives of
  //     the query are skipped (positives are ordered exactly among themselves: rank_j = 1 + above_j + j).
  //     Ordinary rows only see the exactly-counted thresholds [0, n_exact); level-1 sampled rows see
  //     [0, n_l1), level-2 sampled rows all of them.  The histogram holds RAW hit counts: the weight of a
  //     bucket (1, W1 or W2) follows from its index and is applied by hist_to_above_kernel.
  if (ok) {
    const int np = cnts & 255, ne = (cnts >> 8) & 255, n1 = (cnts >> 16) & 255;
    int hi = ((meta & M_S2) ? np : (meta & M_S1) ? n1 : ne) - 1;
    const uint32_t t = A.thr + (uint32_t)(ql * p.pcap) * 4u;
    if (hi >= 0 && s > lds_f32(t + hi * 4) &&                   // invariant: t[hi] < s
        p.g_code[(int64_t)row * p.row_stride] != lds_s32(A.es + ES_OFF(s_qcode) + ql4)) {
      int lo = 0;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (lds_f32(t + mid * 4) < s) hi = mid; else lo = mid + 1;
      }
#if REID_HIST_GLOBAL
      atomicAdd(&p.hist[(q0 + ql) * p.Pmax + lo], 1);
#else
      reds_add(A.hist + (uint32_t)((ql * p.pcap + lo) >> 1) * 4u, 1u << ((lo & 1) * 16));   // pcap is even
#endif
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

// Candidate-threshold refresh for query qq (whole warp): the klist-th largest of the last 2 * klist appended scores
// (klist = 32: a 64-entry window, two values per lane; klist = 16: one value per lane).
__device__ __noinline__ void epi_refresh_thr(EpiAddr A, const Params* pp, int qq, int cnt, int lane, int64_t q0, int chunk) {
  const Params& p = *pp;
  const int kl = p.klist;
  const float* win = p.cand_score + ((q0 + qq) * p.n_chunks + chunk) * (int64_t)p.cand_cap + (cnt - 2 * kl);
  const float a0 = __ldcg(win + lane), a1 = kl == 32 ? __ldcg(win + 32 + lane) : -INFINITY;
  int r0 = 0, r1 = 0;                                          // number of window values greater than a0 / a1
#pragma unroll 4
  for (int k = 0; k < 32; ++k) {
    const float b0 = __shfl_sync(0xffffffffu, a0, k), b1 = __shfl_sync(0xffffffffu, a1, k);
    r0 += (b0 > a0) + (b1 > a0);
    r1 += (b0 > a1) + (b1 > a1);
  }
  // the values with fewer than klist greater ones are >= klist scores; their minimum bounds the klist-th largest
  const float cand = fminf(r0 < kl ? a0 : INFINITY, (kl == 32 && r1 < kl) ? a1 : INFINITY);
  const float nthr = unkey32(__reduce_min_sync(0xffffffffu, key32(cand)));
  if (lane == 0) {
    const uint32_t q4 = (uint32_t)qq * 4u;
    sts_s32(A.es + ES_OFF(s_nextupd) + q4, cnt + kl / 2);
    if (nthr > lds_f32(A.es + ES_OFF(s_thrtop) + q4)) {
      sts_f32(A.es + ES_OFF(s_thrtop) + q4, nthr);
      sts_f32(A.es + ES_OFF(s_min) + q4, fminf(nthr, lds_f32(A.es + ES_OFF(s_threx) + q4)));
      sts_f32(A.es + ES_OFF(s_minS1) + q4, fminf(nthr, lds_f32(A.es + ES_OFF(s_thrl1) + q4)));
      sts_f32(A.es + ES_OFF(s_minS2) + q4, fminf(nthr, lds_f32(A.es + ES_OFF(s_thrlow) + q4)));
    }
  }
  __syncwarp();
}

// spill the packed 16-bit counters of this CTA into the global histogram
__device__ __noinline__ void epi_flush_hist(EpiAddr A, const Params* pp, int et, int64_t q0) {
  const Params& p = *pp;
  if (REID_HIST_GLOBAL) return;
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

// calibration: hist holds bucket counts over the CALIB_ROWS calibration rows (zeroed here for the main pass).
// Threshold j is counted exactly while the estimated in-chunk rank of thresholds 0..j stays <= limit1, on the
// level-1 sample while it stays <= limit2, else on the level-2 sample.
// One warp per query (Pmax <= 64: two buckets per lane, warp-wide inclusive scan), coalesced over the histogram rows.
__global__ void __launch_bounds__(256)
calib_split_kernel(int32_t* __restrict__ hist, const int32_t* __restrict__ n_pos, int64_t Q, int Pmax,
                   float scale, float limit1, float limit2, int32_t* __restrict__ n_exact,
                   int32_t* __restrict__ n_l1, const float* __restrict__ pos_thr, int pos_stride, int klist,
                   uint32_t* __restrict__ thr_share) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t q = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); q < Q; q += warps) {
    const int np = min(n_pos[q], Pmax);
    int32_t* h = hist + q * Pmax;
    const int j0 = lane, j1 = lane + 32;
    const int h0 = j0 < Pmax ? h[j0] : 0, h1 = j1 < Pmax ? h[j1] : 0;
    if (j0 < Pmax) h[j0] = 0;
    if (j1 < Pmax) h[j1] = 0;
    int a0 = h0, a1 = h1;                                    // inclusive prefix sums over buckets 0..31 and 32..63
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u0 = __shfl_up_sync(0xffffffffu, a0, o), u1 = __shfl_up_sync(0xffffffffu, a1, o);
      if (lane >= o) { a0 += u0; a1 += u1; }
    }
    a1 += __shfl_sync(0xffffffffu, a0, 31);
    // first bucket (in order) whose estimate passes a limit / reaches klist rows: 64-bit ballots, lowest set bit
    auto first = [&](bool c0, bool c1) -> int {
      const unsigned b0 = __ballot_sync(0xffffffffu, c0 && j0 < np), b1 = __ballot_sync(0xffffffffu, c1 && j1 < np);
      return b0 ? __ffs(b0) - 1 : (b1 ? 32 + __ffs(b1) - 1 : -1);
    };
    const int f1 = first((float)a0 * scale > limit1, (float)a1 * scale > limit1);
    const int f2 = first((float)a0 * scale > limit2, (float)a1 * scale > limit2);
    const int seed = first(a0 >= klist, a1 >= klist);
    if (lane == 0) {
      const int ne = f1 < 0 ? np : f1, n1 = f2 < 0 ? np : f2;
      n_exact[q] = ne;
      n_l1[q] = max(n1, ne);
      // seed of the candidate threshold: the sample rows are shard rows, so the shallowest positive threshold with >= klist
      // sample rows (valid, non-positive) strictly above it has >= klist shard rows above it -- the main pass appends exactly
      // the rows above it and skips the cold start (-inf: every row of the first tiles)
      if (thr_share && seed >= 0) thr_share[q] = key32(pos_thr[q * (int64_t)pos_stride + seed]);
    }
  }
}
__global__ void fill_n_exact_kernel(const int32_t* __restrict__ n_pos, int64_t Q, int Pmax, int32_t* __restrict__ n_exact,
                                    int32_t* __restrict__ n_l1) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < Q; q += (int64_t)gridDim.x * blockDim.x)
    n_exact[q] = n_l1[q] = min(n_pos[q], Pmax);
}

// pos_above[q, j] += sum_{b <= j} weight(b) * hist[q, b],  weight = 1 / W1 / W2 by the bucket's counting class
__global__ void hist_to_above_kernel(const int32_t* __restrict__ hist, const int32_t* __restrict__ n_pos,
                                     const int32_t* __restrict__ n_exact, const int32_t* __restrict__ n_l1, int64_t Q,
                                     int Pmax, int pos_stride, int32_t* __restrict__ pos_above,
                                     const uint32_t* __restrict__ thr_share, float* __restrict__ cand_thr) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < Q; q += (int64_t)gridDim.x * blockDim.x) {
    if (cand_thr) cand_thr[q] = unkey32(thr_share[q]);
    const int np = min(n_pos[q], Pmax), ne = n_exact[q], n1 = n_l1[q];
    int acc = 0;
    for (int j = 0; j < np; ++j) {
      acc += hist[q * Pmax + j] * (j < ne ? 1 : (j < n1 ? W1 : W2));
      pos_above[q * (int64_t)pos_stride + j] += acc;
    }
  }
}

// A = 256 resident queries (128 per CTA -> the 128 TMEM lanes of that CTA), B = 256 streamed gallery rows (each CTA
// loads half of every tile), one M256 x N256 x K16 MMA per issue by the leader CTA.
__global__ void __launch_bounds__(THREADS, 1)
retrieve_fused_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmQ,
                      const __grid_constant__ Params prm) {
  const uint32_t crank = tc::cluster_ctarank();
  const bool leader = crank == 0;
  const int unit = (int)(blockIdx.x >> 1);              // scheduling unit: CTA pair
  const int n_units = (int)(gridDim.x >> 1);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  const Params& p = prm;
  uint8_t* sB = smem;                                         // [kchunks][B_CHUNK] resident query block
  uint8_t* sA = sB + p.kchunks * B_CHUNK;                      // [stages][A_STAGE]  gallery ring
  float* s_thr = reinterpret_cast<float*>(sA + p.stages * A_STAGE);              // [NQ][pcap]
  uint32_t* s_hist32 = reinterpret_cast<uint32_t*>(s_thr + NQ * p.pcap);         // [NQ][pcap/2]
  float* s_qs = reinterpret_cast<float*>(s_hist32 + (REID_HIST_GLOBAL ? 0 : NQ * p.pcap / 2));   // [EPI_WARPS][QCAP]
  uint32_t* s_qm = reinterpret_cast<uint32_t*>(s_qs + EPI_WARPS * QCAP);         // [EPI_WARPS][QCAP]
  EpiState* es = reinterpret_cast<EpiState*>(s_qm + EPI_WARPS * QCAP);
  __shared__ __align__(8) uint64_t full[MAX_STAGES], empty[MAX_STAGES], bfull, bempty, tfull[NB], tempty[NB];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    tc::mbar_init(&bfull, 1); tc::mbar_init(&bempty, 1);
    // the leader's tempty collects the epilogue warps of BOTH CTAs
    for (int b = 0; b < NB; ++b) { tc::mbar_init(&tfull[b], 1); tc::mbar_init(&tempty[b], 2 * EPI_WARPS); }
    tc::fence_barrier_init();
    tc::prefetch_tensormap(&tmG); tc::prefetch_tensormap(&tmQ);
  }
  if (warp == 2) tc::tmem_alloc_pair(&tmem_base_s, TMEM_COLS);
  tc::fence_before_sync();
  tc::cluster_sync_all();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  const int n_items = p.n_full + (p.n_qblocks - p.n_full) * p.n_chunks;     // n_qblocks counts blocks of 2 * NQ queries

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------------ TMA producer
    // sB holds the RESIDENT query block of this CTA (128 queries, the MMA's A operand); sA is the gallery ring
    // (128 rows per stage per CTA: the two CTAs load the two halves of a tile)
    uint32_t it = 0, ph = 0;
    int st = 0;
    const uint32_t bfull_l = tc::leader_addr(&bfull);
    for (int item = unit; item < n_items; item += n_units, ++it) {
      const Item im = get_item(p, item);
      const int64_t row0 = im.row0;
      const int ntiles = item_tiles(im);
      const int qrow = im.qb * 2 * NQ + (int)crank * NQ;
      tc::mbar_wait(&bempty, (it & 1) ^ 1);            // previous item's MMAs have finished with the queries
      if (leader) tc::mbar_arrive_expect_tx(&bfull, (uint32_t)(p.kchunks * B_CHUNK) * 2u);
#pragma unroll 1
      for (int kc = 0; kc < p.kchunks; ++kc) tc::tma_load_2d_pair(sB + kc * B_CHUNK, &tmQ, bfull_l, kc * BK, qrow);
#pragma unroll 1
      for (int t = 0; t < ntiles; ++t) {
        const int grow = (int)(row0 + (int64_t)t * TROWS) + (int)crank * TMG;
        if (PF_DIST > 0) {
          // far L2 prefetch: the pairs of a wave stream the same chunk in loose lock-step; tile t + PF_DIST is
          // requested ONCE, by the pair it maps to, long before the pack of pairs reaches it
          const int tp = t + PF_DIST;
          if (tp < ntiles && (tp % n_units) == unit) {
#pragma unroll 1
            for (int kc = 0; kc < p.kchunks; ++kc) tma_prefetch_2d(&tmG, kc * BK, grow + PF_DIST * TROWS);
          }
        }
#pragma unroll 1
        for (int kc = 0; kc < p.kchunks && !FDBG(128); ++kc) {
          tc::mbar_wait(&empty[st], ph ^ 1);
          if (leader) tc::mbar_arrive_expect_tx(&full[st], A_STAGE * 2u);
          tc::tma_load_2d_pair(sA + st * A_STAGE, &tmG, tc::leader_addr(&full[st]), kc * BK, grow);
          if (++st == p.stages) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    // D[queries 256 x gallery 256] = A(queries) . B(gallery)^T
    constexpr uint32_t idesc = tc::make_idesc_f16(2 * NQ, TROWS, 0);
    uint32_t it = 0, ph = 0, tilecount = 0;
    int st = 0;
    for (int item = unit; item < n_items; item += n_units, ++it) {
      const int ntiles = item_tiles(get_item(p, item));
      tc::mbar_wait(&bfull, it & 1);
      tc::fence_after_sync();
#pragma unroll 1
      for (int t = 0; t < ntiles; ++t, ++tilecount) {
        const uint32_t buf = tilecount % NB, bph = (tilecount / NB) & 1;
#if REID_DEBUG & 8192
        const long long tw0 = FDBG(8192) ? clock64() : 0;
#endif
        tc::mbar_wait(&tempty[buf], bph ^ 1);          // epilogue has drained this accumulator
        tc::fence_after_sync();
#if REID_DEBUG & 8192
        if (FDBG(8192)) { atomicAdd(&g_dbg[2], (unsigned long long)(clock64() - tw0)); atomicAdd(&g_dbg[5], 1ull); }
#endif
#pragma unroll 1
        for (int kc = 0; kc < p.kchunks; ++kc) {
#if REID_DEBUG & 8192
          const long long tf0 = FDBG(8192) ? clock64() : 0;
#endif
          if (!FDBG(128)) tc::mbar_wait(&full[st], ph);
          tc::fence_after_sync();
#if REID_DEBUG & 8192
          if (FDBG(8192)) atomicAdd(&g_dbg[3], (unsigned long long)(clock64() - tf0));
#endif
          const uint64_t gd = tc::make_smem_desc_sw128(tc::smem_u32(sA + st * A_STAGE));    // gallery stage
          const uint64_t qd = tc::make_smem_desc_sw128(tc::smem_u32(sB + kc * B_CHUNK));    // resident queries
#pragma unroll
          for (int k = 0; k < (FDBG(256) ? 0 : BK / 16); ++k)
            tc::mma_f16_ss_pair(tmem_base + buf * TROWS, tc::advance_desc_k(qd, k), tc::advance_desc_k(gd, k), idesc, (kc | k) != 0);
          if (!FDBG(128)) tc::mma_commit_pair(&empty[st]);   // frees the gallery stage
          if (++st == p.stages) { st = 0; ph ^= 1; }
        }
        tc::mma_commit_pair(&tfull[buf]);   // accumulator complete -> epilogue of both CTAs
      }
      tc::mma_commit_pair(&bempty);         // query block no longer read
    }
  } else if (warp >= EPI_WARP0) {
    // ------------------------------------------------------------------ epilogue (EPI_WARPS warps)
    const int quad = warp & 3;                          // TMEM lane quadrant of this warp
    const int ew = warp - EPI_WARP0;                    // 0..EPI_WARPS-1
    const int part = ew >> 2;                           // which slice of the accumulator columns
    const int et = threadIdx.x - EPI_WARP0 * 32;        // 0..EPI_THREADS-1
    const uint32_t tmem_q = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int myq = quad * 32 + lane;                   // query (TMEM lane) of this thread
    float* my_qs = s_qs + ew * QCAP;
    uint32_t* my_qm = s_qm + ew * QCAP;
    EpiAddr A;
    A.es = tc::smem_u32(es); A.thr = tc::smem_u32(s_thr); A.hist = tc::smem_u32(s_hist32);
    A.qs = tc::smem_u32(s_qs); A.qm = tc::smem_u32(s_qm);
#pragma unroll 1
    for (int i = et; i < (REID_HIST_GLOBAL ? 0 : NQ * p.pcap / 2); i += EPI_THREADS) s_hist32[i] = 0;
    epi_bar();
    uint32_t tilecount = 0;
    for (int item = unit; item < n_items; item += n_units) {
      const Item im = get_item(p, item);
      const int chunk = im.slot;                           // candidate slot of this item
      const int64_t row0 = im.row0;
      const int ntiles = item_tiles(im);
      const int64_t q0 = (int64_t)im.qb * 2 * NQ + (int64_t)crank * NQ;   // first query of THIS CTA's block
      const bool appends = !p.calib && !p.no_cand;
      // ---- item setup: per-query state
      if (et < NQ) {
        const int64_t q = q0 + et;
        const bool live = q < p.Q;
        const int np = (live && !FDBG(1)) ? min(p.n_pos[q], p.Pmax) : 0;
        const int ne = p.calib ? np : (live ? min(p.n_exact[q], np) : 0);
        const int n1 = p.calib ? np : (live ? max(ne, min(p.n_l1[q], np)) : 0);
        es->s_qcode[et] = live ? p.q_code[q] : -2;
        const float* tq = p.pos_thr + q * (int64_t)p.pos_stride;
        const float tex = ne > 0 ? tq[ne - 1] : INFINITY;
        const float t1 = n1 > 0 ? tq[n1 - 1] : INFINITY;
        const float tl = np > 0 ? tq[np - 1] : INFINITY;
        es->s_threx[et] = tex;
        es->s_thrl1[et] = t1;
        es->s_thrlow[et] = tl;
        // candidates: warm start from the best threshold any earlier chunk of this query has published;
        // padded queries, the calibration pre-pass and counting-only passes never append
        const float tt = (live && appends && !FDBG(2)) ? unkey32(__ldcg(&p.thr_share[q])) : INFINITY;
        es->s_thrtop[et] = tt;
        es->s_min[et] = fminf(tt, tex);
        es->s_minS1[et] = fminf(tt, t1);
        es->s_minS2[et] = fminf(tt, tl);
        es->s_candcnt[et] = 0;
        es->s_nextupd[et] = 2 * p.klist;
        uint32_t he = 0;
        if (live) {
#pragma unroll 1
          for (int e = 0; e < p.E; ++e) he |= (p.excl[q * p.E + e] >= 0) ? 1u : 0u;
        }
        es->s_cnts[et] = (uint32_t)np | ((uint32_t)ne << 8) | ((uint32_t)n1 << 16) | (he << 24);
      }
      if (appends) {                                           // -inf pre-fill of this item's candidate score slots
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
          s_thr[ql * p.pcap + j] = (qq < p.Q && j < p.Pmax) ? p.pos_thr[qq * (int64_t)p.pos_stride + j] : -INFINITY;
      }
      epi_bar();
      int qn = 0;                                          // queued hits of this warp (persist across tiles)
      // ---- tiles
      for (int t = 0; t < ntiles; ++t, ++tilecount) {
        const uint32_t buf = tilecount % NB, bph = (tilecount / NB) & 1;
        const int tile_row0 = (int)(row0 + (int64_t)t * TROWS);
        const bool tile_l2 = (tile_row0 & (W2 - 1)) == 0;              // column 5 of this tile is a level-2 sampled row
        const float minA = es->s_min[myq], minS1 = es->s_minS1[myq], minS2 = es->s_minS2[myq];   // refreshed thresholds, once per tile
        // fp16 images of the thresholds for the packed fast-path test (rounded DOWN after subtracting the rounding bound)
        const __half lowA = __float2half_rd(minA - 5e-4f);
        const __half2 hAA = __halves2half2(lowA, lowA);
        const __half2 hAS1 = __halves2half2(lowA, __float2half_rd(minS1 - 5e-4f));
        const __half2 hAS2 = __halves2half2(lowA, __float2half_rd(minS2 - 5e-4f));
#if REID_DEBUG & 8192
        const long long te0 = FDBG(8192) ? clock64() : 0;
#endif
        tc::mbar_wait(&tfull[buf], bph);
        tc::fence_after_sync();
#if REID_DEBUG & 8192
        const long long te1 = FDBG(8192) ? clock64() : 0;
#endif
        constexpr int STEPS = TROWS * 4 / EPI_WARPS / 16;       // 16-column steps of this warp's slice (64 columns)
#pragma unroll 1
        for (int step = 0; step < (FDBG(4) ? 0 : STEPS); ++step) {
          const int c0 = part * (STEPS * 16) + step * 16;
          // only column 5 of every second step can be a sampled row (rows = 5 mod W1); uniform
          const bool step_s1 = ((c0 + 5) & (W1 - 1)) == 5;
          const bool step_s2 = tile_l2 && c0 == 0;
          const float minS = step_s2 ? minS2 : minS1;
          const __half2 hAS = step_s2 ? hAS2 : hAS1;
          uint32_t r[16];
          tc::tmem_ld_x16(tmem_q + buf * TROWS + c0, r);
          tc::tmem_wait_ld();
          // fast path: a per-lane bit mask of the columns that MAY hit, ONE warp-wide OR per step.  The 16 scores
          // are packed to half2 (one F2FP per pair) and compared pairwise against the thresholds lowered by the fp16
          // rounding bound (5e-4 >= half an ulp of any |score| <= 1, rounded down), so the mask is a superset of the
          // exact hits (a few % more); the slow path re-tests exactly in fp32.  Mask bit p / 16 + p = column 2p / 2p + 1.
          unsigned colmask = 0, bits = 0;
          if (!FDBG(32)) {
#pragma unroll
            for (int pr = 0; pr < 8; ++pr) {
              const __half2 h = __floats2half2_rn(__uint_as_float(r[2 * pr]), __uint_as_float(r[2 * pr + 1]));
              const __half2 thr = (pr == 2 && step_s1) ? hAS : hAA;
              bits |= __hgt2_mask(h, thr) & ((1u << pr) | (1u << (16 + pr)));
            }
            colmask = __reduce_or_sync(0xffffffffu, bits);
          }
          // slow path (single copy of the code: the kernel must fit the instruction cache).  Round n takes
          // the n-th flagged column of EVERY lane at once: per-lane select tree for the score, exact fp32 test, one
          // ballot + prefix popc to compact the round into the dense warp queue.  Rounds per step = max flagged
          // columns per lane (usually 1), not the number of hit columns.
          if (FDBG(4096)) colmask = 0;
#if REID_SLOW_X1 == 2
          // slow path, column by column (warp-uniform loop over the flagged columns, usually one or two per step): the
          // column index is uniform, so the lane's score comes out of the 16 registers through a jump table (one MOV)
          // instead of a per-lane select tree; exact fp32 test, one ballot + prefix popc into the dense warp queue.
          while (colmask) {
            const int bp = __ffs(colmask) - 1;
            colmask &= colmask - 1;
            const int i = ((bp & 15) << 1) | (bp >> 4);
            uint32_t rv;
            switch (i) {
              case 0: rv = r[0]; break;   case 1: rv = r[1]; break;   case 2: rv = r[2]; break;   case 3: rv = r[3]; break;
              case 4: rv = r[4]; break;   case 5: rv = r[5]; break;   case 6: rv = r[6]; break;   case 7: rv = r[7]; break;
              case 8: rv = r[8]; break;   case 9: rv = r[9]; break;   case 10: rv = r[10]; break; case 11: rv = r[11]; break;
              case 12: rv = r[12]; break; case 13: rv = r[13]; break; case 14: rv = r[14]; break; default: rv = r[15]; break;
            }
            const float sc = __uint_as_float(rv);
            const bool smp_col = i == 5 && step_s1;
            const bool hit = sc > (smp_col ? minS : minA);             // exact test
            const unsigned c = __ballot_sync(0xffffffffu, hit);
            if (hit) {
              const int pos = qn + __popc(c & lt_mask);
              my_qs[pos] = sc;
              my_qm[pos] = (uint32_t)myq | (smp_col ? (step_s2 ? (M_S1 | M_S2) : M_S1) : 0u) |
                           ((uint32_t)(tile_row0 + c0 + i) << M_ROW_SHIFT);
            }
            qn += __popc(c);
            if (qn > QCAP - 32) {                        // queue nearly full: drain the newest 32 now (order is irrelevant)
              qn -= 32;
              epi_drain32(A, &p, ew, qn, 32, lane, q0, chunk);
            }
          }
#elif REID_SLOW_X1
          // slow path, column by column (warp-uniform loop over the flagged columns, usually one or two per step): the
          // column is read again from TMEM (one 32x32b.x1 load: every lane gets ITS score of that column -- no register
          // select tree), tested exactly in fp32, and the hits of the 32 lanes are compacted into the dense warp queue with
          // one ballot + prefix popc.
          while (colmask) {
            const int bp = __ffs(colmask) - 1;
            colmask &= colmask - 1;
            const int i = ((bp & 15) << 1) | (bp >> 4);
            const float sc = __uint_as_float(tc::tmem_ld_x1(tmem_q + buf * TROWS + c0 + i));
            tc::tmem_wait_ld();
            const bool smp_col = i == 5 && step_s1;
            const bool hit = sc > (smp_col ? minS : minA);             // exact test
            const unsigned c = __ballot_sync(0xffffffffu, hit);
            if (hit) {
              const int pos = qn + __popc(c & lt_mask);
              my_qs[pos] = sc;
              my_qm[pos] = (uint32_t)myq | (smp_col ? (step_s2 ? (M_S1 | M_S2) : M_S1) : 0u) |
                           ((uint32_t)(tile_row0 + c0 + i) << M_ROW_SHIFT);
            }
            qn += __popc(c);
            if (qn > QCAP - 32) {                        // queue nearly full: drain the newest 32 now (order is irrelevant)
              qn -= 32;
              epi_drain32(A, &p, ew, qn, 32, lane, q0, chunk);
            }
          }
#else
          // slow path (single copy of the code: the kernel must fit the instruction cache).  Round n takes
          // the n-th flagged column of EVERY lane at once: per-lane select tree for the score, exact fp32 test, one
          // ballot + prefix popc to compact the round into the dense warp queue.  Rounds per step = max flagged
          // columns per lane (usually 1), not the number of hit columns.
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
              const bool smp_col = i == 5 && step_s1;
              const bool hit = has && sc > (smp_col ? minS : minA);     // exact test
              const unsigned c = __ballot_sync(0xffffffffu, hit);
              if (hit) {
                const int pos = qn + __popc(c & lt_mask);
                my_qs[pos] = sc;
                my_qm[pos] = (uint32_t)myq | (smp_col ? (step_s2 ? (M_S1 | M_S2) : M_S1) : 0u) |
                             ((uint32_t)(tile_row0 + c0 + i) << M_ROW_SHIFT);
              }
              qn += __popc(c);
              if (qn > QCAP - 32) {                      // queue nearly full: drain the newest 32 now (order is irrelevant)
                qn -= 32;
                epi_drain32(A, &p, ew, qn, 32, lane, q0, chunk);
              }
            }
          }
#endif
        }
        tc::fence_before_sync();
        if (lane == 0) { if (!leader) tc::mbar_arrive_remote(&tempty[buf], 0); else tc::mbar_arrive(&tempty[buf]); }
#if REID_DEBUG & 8192
        const long long te2 = FDBG(8192) ? clock64() : 0;
        if (FDBG(8192) && lane == 0) {
          atomicAdd(&g_dbg[0], (unsigned long long)(te2 - te1)); atomicAdd(&g_dbg[1], (unsigned long long)(te1 - te0));
          atomicAdd(&g_dbg[9], 1ull);
        }
#endif
        // the accumulator is handed back; full batches are drained now, overlapping the next tile's MMA
        while (qn >= 32) {
          if (REID_SCAN_FIRST && qn <= QCAP - 64 && t + 1 < ntiles &&
              mbar_test(&tfull[(tilecount + 1) % NB], ((tilecount + 1) / NB) & 1)) break;   // the next accumulator is already waiting
          qn -= 32;
          epi_drain32(A, &p, ew, qn, 32, lane, q0, chunk);
#if REID_DEBUG & 8192
          if (FDBG(8192) && lane == 0) atomicAdd(&g_dbg[6], 1ull);
#endif
        }
#if REID_DEBUG & 8192
        const long long te3 = FDBG(8192) ? clock64() : 0;
#endif
        // ---- tile boundary: this warp owns the candidate thresholds of a fixed group of queries
        //      (UPD_PER_WARP each).  The score slots of an item are pre-filled with -inf, so a window that
        //      contains a slot whose store is still in flight only yields a more conservative bound:
        //      no CTA barrier and no completion protocol are needed.
        if (appends) {
          const int uq = ew * UPD_PER_WARP + (lane & (UPD_PER_WARP - 1));
          const int cnt = min(*(volatile int*)&es->s_candcnt[uq], p.cand_cap);
          unsigned um = __ballot_sync(0xffffffffu, lane < UPD_PER_WARP && cnt >= 2 * p.klist && cnt >= es->s_nextupd[uq]);
          while (um) {
            const int src = __ffs(um) - 1;
            um &= um - 1;
            epi_refresh_thr(A, &p, ew * UPD_PER_WARP + src, __shfl_sync(0xffffffffu, cnt, src), lane, q0, chunk);
#if REID_DEBUG & 8192
            if (FDBG(8192) && lane == 0) atomicAdd(&g_dbg[8], 1ull);
#endif
          }
        }
        if (p.calib) {
          // (deterministic classes: every hit of this tile is counted, by all warps, before the owners look at the counts, and
          //  every warp sees the retired state before it scans the next tile -- two barriers per tile, calibration only)
          while (qn > 0) { const int n = qn < 32 ? qn : 32; qn -= n; epi_drain32(A, &p, ew, qn, n, lane, q0, chunk); }
          epi_bar();
          if (lane < UPD_PER_WARP) {
            // calibration: a threshold whose ESTIMATED in-shard rank (sample rows above it so far, scaled by the part of the
            // sample seen) has passed the level-2 limit is retired on the spot -- never on fewer than calib_floor rows, so a
            // retired threshold also has >= klist shard rows above it (the seed of the candidate threshold).  Retiring stops
            // the rows between the deep thresholds from being hits at all (most sample rows score above a query's DEEPEST
            // positives).  The bucket of the shallowest retired threshold receives calib_cap extra counts: calib_split then
            // sees every retired threshold beyond the level-2 limit.
            const int uq = ew * UPD_PER_WARP + lane;
            const uint32_t cn = es->s_cnts[uq];
            const int np = cn & 255;
            const int seen = min((t + 1) * TROWS, (int)p.G_local);            // (the sample has <= CALIB_ROWS rows)
            const int cap = max(p.calib_floor, (int)(p.calib_rate * (float)seen));
            int acc = 0, na = np;
#pragma unroll 1
            for (int j = 0; j < np; ++j) {
              const uint32_t w = s_hist32[(uq * p.pcap + j) >> 1];
              acc += (j & 1) ? (int)(w >> 16) : (int)(w & 0xFFFFu);
              if (acc >= cap) { na = j; break; }
            }
            if (na < np) {
              es->s_cnts[uq] = (cn & 0xFF000000u) | (uint32_t)na | ((uint32_t)na << 8) | ((uint32_t)na << 16);
              const float tl = na > 0 ? s_thr[uq * p.pcap + na - 1] : INFINITY;
              es->s_threx[uq] = tl; es->s_thrl1[uq] = tl; es->s_thrlow[uq] = tl;
              es->s_min[uq] = tl; es->s_minS1[uq] = tl; es->s_minS2[uq] = tl;
              reds_add(A.hist + (uint32_t)((uq * p.pcap + na) >> 1) * 4u, (uint32_t)p.calib_cap << ((na & 1) * 16));
            }
          }
          epi_bar();
        }
        if (((t + 1) % FLUSH_TILES) == 0) epi_flush_hist(A, &p, et, q0);
#if REID_DEBUG & 8192
        if (FDBG(8192) && lane == 0) {
          const long long te4 = clock64();
          atomicAdd(&g_dbg[4], (unsigned long long)(te3 - te2)); atomicAdd(&g_dbg[7], (unsigned long long)(te4 - te3));
          atomicMax(&g_dbg[10], (unsigned long long)(te4 - te1));
        }
#endif
      }
      // ---- item end: drain the queue, spill the histogram, publish candidate state
      while (qn > 0) { const int n = qn < 32 ? qn : 32; qn -= n; epi_drain32(A, &p, ew, qn, n, lane, q0, chunk); }
      epi_bar();
      epi_flush_hist(A, &p, et, q0);
      if (et < NQ && q0 + et < p.Q && appends) {
        p.cand_count[(q0 + et) * p.n_chunks + chunk] = es->s_candcnt[et];
        atomicMax(&p.thr_share[q0 + et], key32(es->s_thrtop[et]));
      }
      epi_bar();
    }
  }
  tc::fence_before_sync();
  tc::cluster_sync_all();   // the peer may still be signalled / read until here
  if (warp == 2) tc::tmem_dealloc_pair(tmem_base, TMEM_COLS);
}

size_t fused_smem_bytes(int kchunks, int stages, int pcap) {
  return (size_t)kchunks * B_CHUNK + (size_t)stages * A_STAGE + (size_t)NQ * pcap * 4 /*thr*/ +
         (REID_HIST_GLOBAL ? 0 : (size_t)NQ * pcap * 2) /*hist*/ + (size_t)EPI_WARPS * QCAP * 8 /*queues*/ +
         sizeof(EpiState) + 1024;
}

}  // namespace

#if REID_DEBUG
// timing experiments only (never part of the shipped library)
extern "C" int reid_debug_counters(unsigned long long* out, int reset) {
  if (out && cudaMemcpyFromSymbol(out, g_dbg, sizeof(g_dbg)) != cudaSuccess) return REID_E_CUDA;
  if (reset) { unsigned long long z[16] = {0}; if (cudaMemcpyToSymbol(g_dbg, z, sizeof(z)) != cudaSuccess) return REID_E_CUDA; }
  return REID_OK;
}
#endif

// workspace = histogram [Q, Pmax <= 64] + candidate threshold [Q] + exact / level-1 threshold counts [Q] each
extern "C" size_t reid_retrieve_fused_workspace_bytes(int64_t Q, int64_t, int) { return (size_t)Q * 67 * sizeof(int32_t); }

extern "C" int reid_retrieve_fused(const void* q_f16, const void* g_f16, const int32_t* q_code, const int32_t* g_code,
                                   const int32_t* excl, int E, const float* pos_thr, const int32_t* n_pos, int64_t Q,
                                   int64_t G_local, int64_t g_offset, int d, int Pmax, int pos_stride, int n_chunks,
                                   int n_shards, int cand_cap, int flags, int32_t* pos_above, float* cand_score,
                                   int32_t* cand_idx, int32_t* cand_count, float* cand_thr, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  const bool no_cand = (flags & REID_FUSED_NO_CANDIDATES) != 0;
  if (!q_f16 || !g_f16 || !q_code || !g_code || !pos_thr || !n_pos || !pos_above || Q <= 0 || G_local <= 0 ||
      n_chunks <= 0 || (E > 0 && !excl) || E < 0 ||
      (flags & ~(REID_FUSED_EXACT_COUNTS | REID_FUSED_NO_CANDIDATES | REID_FUSED_KLIST16)))
    return REID_E_INVALID;
  if (!no_cand && (!cand_score || !cand_idx || !cand_count || cand_cap < 64 || cand_cap % 4 != 0)) return REID_E_INVALID;
  if (pos_stride == 0) pos_stride = Pmax;
  if (pos_stride < Pmax) return REID_E_INVALID;
  if (d % BK != 0 || d > 512 || Pmax <= 0 || Pmax > 64 || G_local > MAX_ROWS) return REID_E_UNSUPPORTED;
  Params p;
  p.q_code = q_code; p.g_code = g_code; p.excl = excl; p.E = E; p.pos_thr = pos_thr; p.n_pos = n_pos;
  p.Q = Q; p.G_local = G_local; p.g_offset = g_offset;
  if (!workspace || workspace_bytes < (size_t)Q * (Pmax + 3) * sizeof(int32_t)) return REID_E_WORKSPACE;
  p.Pmax = Pmax; p.pos_stride = pos_stride; p.pcap = (Pmax + 3) / 4 * 4; p.kchunks = d / BK;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return REID_E_CUDA;
  const int max_units = sms / 2;                       // persistent CTA pairs
  p.n_chunks = n_chunks; p.n_qblocks = (int)((Q + 2 * NQ - 1) / (2 * NQ)); p.cand_cap = cand_cap;
  // whole waves of query blocks run against the whole shard; the blocks of the last, partial wave are cut into n_chunks
  p.n_full = n_chunks > 1 ? (p.n_qblocks / max_units) * max_units : p.n_qblocks;
  const int64_t rpc = (G_local + n_chunks - 1) / n_chunks;
  p.rows_per_chunk = (rpc + W2 - 1) / W2 * W2;         // chunk starts are multiples of W2: (local row % W) is a tile column
  p.hist = (int32_t*)workspace; p.thr_share = (uint32_t*)workspace + (size_t)Q * Pmax;
  p.n_exact = (int32_t*)workspace + (size_t)Q * (Pmax + 1); p.n_l1 = (int32_t*)workspace + (size_t)Q * (Pmax + 2);
  p.calib = 0; p.calib_cap = 0; p.calib_floor = 0; p.calib_rate = 0.f; p.row_stride = 1; p.no_cand = no_cand ? 1 : 0;
  p.klist = (flags & REID_FUSED_KLIST16) ? 16 : KL;
  p.cand_score = cand_score; p.cand_idx = cand_idx; p.cand_count = cand_count;
  p.debug = REID_DEBUG;
  int stages = REID_MAX_STAGES < MAX_STAGES ? REID_MAX_STAGES : MAX_STAGES;
  const size_t smem_max = 227 * 1024 - 512;            // (static shared: barriers)
  while (stages > 2 && fused_smem_bytes(p.kchunks, stages, p.pcap) > smem_max) --stages;
  if (fused_smem_bytes(p.kchunks, stages, p.pcap) > smem_max) return REID_E_UNSUPPORTED;
  p.stages = stages;
  const size_t smem = fused_smem_bytes(p.kchunks, stages, p.pcap);
  CUtensorMap tmG, tmQ;
  if (!tc_host::make_map_f16(&tmG, g_f16, G_local, d, TMG) ||
      !tc_host::make_map_f16(&tmQ, q_f16, Q, d, NQ))
    return REID_E_CUDA;
  if (cudaFuncSetAttribute(retrieve_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return REID_E_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  // one persistent CTA pair per two SMs, launched as clusters of 2
  auto launch = [&](const CUtensorMap& mg, const Params& pp) -> bool {
    const int n_items = pp.n_full + (pp.n_qblocks - pp.n_full) * pp.n_chunks;
    const int units = n_items < max_units ? n_items : max_units;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * units); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, retrieve_fused_kernel, mg, tmQ, pp) == cudaSuccess;
  };
  if (cudaMemsetAsync(workspace, 0, (size_t)Q * (Pmax + 3) * sizeof(int32_t), st) != cudaSuccess) return REID_E_CUDA;
  const int aux_grid = (int)reid_min64((Q + 255) / 256, 148 * 8);
  bool sample_deep = (G_local >= 16 * CALIB_MIN) && !(flags & REID_FUSED_EXACT_COUNTS);
  // calibration sample: 1/32 of the shard in whole tiles, CALIB_MIN .. CALIB_ROWS rows
  int64_t calib_rows = G_local / 32 / TROWS * TROWS;
  if (calib_rows > CALIB_ROWS) calib_rows = CALIB_ROWS;
  if (calib_rows < CALIB_MIN) calib_rows = CALIB_MIN;
  if (FDBG(64)) sample_deep = false;
  if (sample_deep) {
    // calibration pre-pass: the same kernel over a strided sample of CALIB_ROWS gallery rows, every
    // threshold counted exactly, no candidates -> per-query bucket histogram of the sample
    Params c = p;
    c.calib = 1;
    c.row_stride = G_local / calib_rows;
    c.G_local = calib_rows;
    c.n_chunks = 1;
    c.n_full = c.n_qblocks;
    c.rows_per_chunk = calib_rows;
    CUtensorMap tmS;
    {
      tc_host::EncodeTiledFn enc = tc_host::get_encode();
      if (!enc) return REID_E_CUDA;
      cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)calib_rows};
      cuuint64_t strides[1] = {(cuuint64_t)d * 2 * (cuuint64_t)c.row_stride};
      cuuint32_t box[2] = {64, (cuuint32_t)TMG};
      cuuint32_t estr[2] = {1, 1};
      if (enc(&tmS, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(g_f16), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return REID_E_CUDA;
    }
    // a sampled count must rest on >= 32 sampled rows above the threshold OVER ALL SHARDS (17.7 % standard error at
    // the boundary): the budgets are gallery-wide, a threshold's class is the same in every chunk of a shard;
    // never classify on fewer than 8 calibration hits.  scale: calibration hits -> estimated rank inside the shard
    const float scale = (float)G_local / (float)calib_rows;
    const int shards = n_shards > 1 ? n_shards : 1;
    const float limit1 = fmaxf(32.f * (float)W1 / (float)shards, 8.f * scale);
    const float limit2 = fmaxf(32.f * (float)W2 / (float)shards, limit1);
    c.calib_cap = (int)(limit2 / scale) + 8;
    c.calib_floor = 32;
    c.calib_rate = (float)c.calib_cap / (float)calib_rows;
    if (!launch(tmS, c)) return REID_E_CUDA;
    calib_split_kernel<<<(int)reid_min64((Q + 7) / 8, 148 * 16), 256, 0, st>>>(p.hist, n_pos, Q, Pmax, scale, limit1, limit2, p.n_exact, p.n_l1, pos_thr,
                                                 pos_stride, p.klist, no_cand ? nullptr : p.thr_share);
  } else {
    fill_n_exact_kernel<<<aux_grid, 256, 0, st>>>(n_pos, Q, Pmax, p.n_exact, p.n_l1);
  }
  REID_CHECK_LAUNCH();
  if (!launch(tmG, p)) return REID_E_CUDA;
  hist_to_above_kernel<<<aux_grid, 256, 0, st>>>(p.hist, n_pos, p.n_exact, p.n_l1, Q, Pmax, pos_stride, pos_above,
                                                 p.thr_share, no_cand ? nullptr : cand_thr);
  REID_CHECK_LAUNCH();
  return REID_OK;
}
