// retrieve_fused.cu -- K3+K4: similarity GEMM with the ranking statistics computed in its epilogue.
//
// Replaces the per-query loop body of rank_and_metrics, eval_mm_protocol.py:401-455 (cosine_sim,
// same-image mask, argsort, CMC, AP walk) for one gallery shard, without materialising S or a sort.
//
// Work item = (block of NQ=128 queries, gallery chunk).  A persistent CTA per SM loops over items:
//   * the query block (B operand, 128 x d fp16 = up to 128 KB) is TMA-loaded once per item and stays
//     resident in shared memory; gallery tiles (A operand, 128 rows) stream through a TMA ring;
//   * one thread issues tcgen05.mma (M=128 gallery rows -> TMEM lanes, N=128 queries -> TMEM columns,
//     fp16 x fp16 -> fp32) into one of two 128-column TMEM accumulators;
//   * four epilogue warps read the accumulator with tcgen05.ld: lane l of warp w owns gallery row
//     32w+l, a column is a query, so every per-query quantity is WARP-UNIFORM and a whole column
//     of 32 scores is tested with one compare + ballot against min(top-list threshold, lowest
//     positive threshold).  Hits are compacted into a per-warp queue in shared memory and drained
//     lane-parallel:
//       (a) counting: a binary search over the query's positive thresholds (sorted descending, in
//           shared memory) gives the bucket b = #thresholds >= score; hist[q][b]++ (packed 16-bit
//           counters in shared memory, spilled to a global histogram every 256 tiles); the count of
//           rows ranked above positive j is the prefix sum over buckets <= j;
//       (b) top list: a 32-entry running list per query (one entry per lane, bf16 rounded down so
//           its minimum is a valid lower bound) gives the threshold above which rows are appended
//           to the query's candidate buffer in global memory.
//     The four warps walk the four 32-column groups of a tile in rotated order with a named barrier
//     between phases, so a query's shared state is owned by exactly one warp at a time.
// Roofline: tensor cores, 2*Q*G*d flop; algorithmic HBM bytes are only operands + outputs.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int NQ = 128;        // queries per block  (MMA N, TMEM columns)
constexpr int TMG = 128;       // gallery rows per tile (MMA M, TMEM lanes)
constexpr int BK = 64;         // K chunk: one 128-byte swizzle atom of fp16
constexpr int KL = REID_KLIST; // running top-list length (== warp size)
constexpr int A_STAGE = TMG * BK * 2;   // 16 KB
constexpr int B_CHUNK = NQ * BK * 2;    // 16 KB
constexpr int MAX_STAGES = 4;
constexpr int THREADS = 256;
constexpr int EPI_WARP0 = 4;   // warps 4..7 are the epilogue
constexpr uint32_t TMEM_COLS = 256;      // two 128-column fp32 accumulators
constexpr int QCAP = 128;                // per-warp hit queue entries
constexpr int FLUSH_TILES = 256;         // 16-bit counters: <= 128 increments per tile
constexpr uint32_t M_NOTPOS = 1u << 16;
static_assert(KL == 32, "one list entry per lane");

struct Params {
  const int32_t* q_code; const int32_t* g_code; const int32_t* excl; int E;
  const float* pos_thr; const int32_t* n_pos;
  int64_t Q, G_local, g_offset;
  int Pmax, pcap, kchunks, stages, n_chunks, n_qblocks, cand_cap;
  int64_t rows_per_chunk;
  int32_t* hist;   // [Q, Pmax] global bucket histogram (workspace)
  float* cand_score; int32_t* cand_idx; int32_t* cand_count;
};

// per-CTA shared state of the epilogue (one query block)
struct EpiState {
  float s_min[NQ];      // min(top-list threshold, lowest positive threshold): the fast-path test
  float s_thrtop[NQ];   // current top-list threshold (min of the list), -inf until the list is full
  float s_thrlow[NQ];   // lowest positive threshold (+inf when the query has no positive)
  int s_qcode[NQ];
  int s_npos[NQ];
  int s_candcnt[NQ];
  int s_hasexcl[NQ];
};

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// views of the per-CTA epilogue state in dynamic shared memory
struct EpiShared {
  EpiState* es;
  const float* thr;        // [NQ][pcap] positive thresholds, sorted descending
  uint32_t* hist32;        // [NQ][pcap/2] packed 16-bit bucket counters
  uint16_t* list16;        // [NQ][KL] running top list, bf16 bit patterns (rounded down)
  float* q_s;              // [4][QCAP] hit queue: score
  uint32_t* q_m;           // [4][QCAP] hit queue: ql | row_in_tile << 8 | NOTPOS
};

__device__ __forceinline__ uint16_t bf16_round_down(float x) {
  uint32_t b = __float_as_uint(x);
  uint32_t t = b >> 16;
  if ((b & 0x80000000u) && (b & 0xFFFFu)) t += 1;    // negative: truncation rounds up, step one down
  return (uint16_t)t;
}
__device__ __forceinline__ float bf16_bits_to_float(uint16_t h) { return __uint_as_float((uint32_t)h << 16); }

// Drain `n` queued hits of this warp (all from the tile starting at local gallery row tile_row0).
__device__ __noinline__ void epi_drain(const EpiShared* shp, const Params* pp, int ew, int n, int lane, int tile_row0,
                                       int64_t q0, int chunk) {
  const EpiShared& sh = *shp;
  const Params& p = *pp;
  EpiState* es = sh.es;
  __syncwarp();
  for (int base = 0; base < n; base += 32) {
    const int e = base + lane;
    const bool have = e < n;
    const float s = have ? sh.q_s[ew * QCAP + e] : 0.f;
    const uint32_t meta = have ? sh.q_m[ew * QCAP + e] : 0u;
    const int ql = meta & 127;
    const int row = (meta >> 8) & 127;
    bool ok = have;
    if (have && es->s_hasexcl[ql]) {                 // same-image mask (eval_mm_protocol.py:408-418)
      const int32_t gidx = (int32_t)(p.g_offset + tile_row0 + row);
      for (int x = 0; x < p.E; ++x) ok = ok && (p.excl[(q0 + ql) * p.E + x] != gidx);
    }
    // (a) bucket among the query's positive thresholds: b = #{j : t_j >= s}; rows that are positives of
    //     the query are skipped (positives are ordered exactly among themselves: rank_j = 1 + above_j + j)
    if (ok && (meta & M_NOTPOS) && s > es->s_thrlow[ql]) {
      const float* t = sh.thr + ql * p.pcap;
      int lo = 0, hi = es->s_npos[ql] - 1;           // invariant: t[hi] < s
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (t[mid] < s) hi = mid; else lo = mid + 1;
      }
      atomicAdd(&sh.hist32[(ql * p.pcap + lo) >> 1], (lo & 1) ? 0x10000u : 1u);    // pcap is even
    }
    // (b) running top list -> candidate buffer (serial per accepted row; rare after warm-up)
    unsigned tm = __ballot_sync(0xffffffffu, ok && s > es->s_thrtop[ql]);
    while (tm) {
      const int src = __ffs(tm) - 1;
      tm &= tm - 1;
      const float v = __shfl_sync(0xffffffffu, s, src);
      const uint32_t m2 = __shfl_sync(0xffffffffu, meta, src);
      const int qq = m2 & 127;
      const float thr = es->s_thrtop[qq];             // may have moved within this batch
      if (v > thr) {
        const float lv = bf16_bits_to_float(sh.list16[qq * KL + lane]);
        const float lmin = warp_min(lv);                // == thr once the list is full, -inf before
        const unsigned holders = __ballot_sync(0xffffffffu, lv == lmin);
        const int victim = __ffs(holders) - 1;
        const uint16_t nv16 = bf16_round_down(v);
        const float nlv = (lane == victim) ? bf16_bits_to_float(nv16) : lv;
        const float nthr = warp_min(nlv);
        if (lane == victim) sh.list16[qq * KL + lane] = nv16;
        if (lane == 0) {
          const int cc = es->s_candcnt[qq];
          if (cc < p.cand_cap) {
            const int64_t o = ((q0 + qq) * p.n_chunks + chunk) * (int64_t)p.cand_cap + cc;
            p.cand_score[o] = v;
            p.cand_idx[o] = tile_row0 + (int)((m2 >> 8) & 127);
          }
          es->s_candcnt[qq] = cc + 1;
          es->s_thrtop[qq] = nthr;
          es->s_min[qq] = fminf(nthr, es->s_thrlow[qq]);
        }
        __syncwarp();
      }
    }
  }
  __syncwarp();
}

// spill the packed 16-bit counters of this CTA into the global histogram
__device__ __forceinline__ void epi_flush_hist(const EpiShared& sh, const Params& p, int et, int64_t q0) {
  const int words = NQ * p.pcap / 2;
  for (int i = et; i < words; i += 128) {
    const uint32_t w = sh.hist32[i];
    if (w) {
      const int ql = (2 * i) / p.pcap, b = (2 * i) % p.pcap;
      const int64_t q = q0 + ql;
      if (q < p.Q) {
        if ((w & 0xFFFFu) && b < p.Pmax) atomicAdd(&p.hist[q * p.Pmax + b], (int)(w & 0xFFFFu));
        if ((w >> 16) && b + 1 < p.Pmax) atomicAdd(&p.hist[q * p.Pmax + b + 1], (int)(w >> 16));
      }
      sh.hist32[i] = 0;
    }
  }
}

// pos_above[q, j] += sum_{b <= j} hist[q, b]
__global__ void hist_to_above_kernel(const int32_t* __restrict__ hist, const int32_t* __restrict__ n_pos, int64_t Q,
                                     int Pmax, int32_t* __restrict__ pos_above) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < Q; q += (int64_t)gridDim.x * blockDim.x) {
    const int np = min(n_pos[q], Pmax);
    int acc = 0;
    for (int j = 0; j < np; ++j) {
      acc += hist[q * Pmax + j];
      pos_above[q * Pmax + j] += acc;
    }
  }
}

__global__ void __launch_bounds__(THREADS, 1)
retrieve_fused_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmQ,
                      const __grid_constant__ Params prm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  const Params& p = prm;
  uint8_t* sB = smem;                                         // [kchunks][B_CHUNK] resident query block
  uint8_t* sA = sB + p.kchunks * B_CHUNK;                      // [stages][A_STAGE]  gallery ring
  float* s_thr = reinterpret_cast<float*>(sA + p.stages * A_STAGE);              // [NQ][pcap]
  uint32_t* s_hist32 = reinterpret_cast<uint32_t*>(s_thr + NQ * p.pcap);         // [NQ][pcap/2]
  uint16_t* s_list16 = reinterpret_cast<uint16_t*>(s_hist32 + NQ * p.pcap / 2);  // [NQ][KL]
  float* s_qs = reinterpret_cast<float*>(s_list16 + NQ * KL);                    // [4][QCAP]
  uint32_t* s_qm = reinterpret_cast<uint32_t*>(s_qs + 4 * QCAP);                 // [4][QCAP]
  EpiState* es = reinterpret_cast<EpiState*>(s_qm + 4 * QCAP);
  __shared__ EpiShared sh_views;
  __shared__ __align__(8) uint64_t full[MAX_STAGES], empty[MAX_STAGES], bfull, bempty, tfull[2], tempty[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    tc::mbar_init(&bfull, 1); tc::mbar_init(&bempty, 1);
    for (int b = 0; b < 2; ++b) { tc::mbar_init(&tfull[b], 1); tc::mbar_init(&tempty[b], 4); }
    tc::fence_barrier_init();
    tc::prefetch_tensormap(&tmG); tc::prefetch_tensormap(&tmQ);
  }
  if (warp == 2) tc::tmem_alloc(&tmem_base_s, TMEM_COLS);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  const int n_items = p.n_qblocks * p.n_chunks;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------------ TMA producer
    uint32_t it = 0, ring = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int chunk = item / p.n_qblocks, qb = item % p.n_qblocks;
      const int64_t row0 = chunk * p.rows_per_chunk;
      const int64_t row1 = reid_min64(p.G_local, row0 + p.rows_per_chunk);
      const int ntiles = (int)((row1 - row0 + TMG - 1) / TMG);
      tc::mbar_wait(&bempty, (it & 1) ^ 1);            // previous item's MMAs have finished with B
      tc::mbar_arrive_expect_tx(&bfull, (uint32_t)(p.kchunks * B_CHUNK));
      for (int kc = 0; kc < p.kchunks; ++kc) tc::tma_load_2d(sB + kc * B_CHUNK, &tmQ, &bfull, kc * BK, qb * NQ);
      for (int t = 0; t < ntiles; ++t) {
        for (int kc = 0; kc < p.kchunks; ++kc, ++ring) {
          const int st = ring % p.stages; const uint32_t ph = (ring / p.stages) & 1;
          tc::mbar_wait(&empty[st], ph ^ 1);
          tc::mbar_arrive_expect_tx(&full[st], A_STAGE);
          tc::tma_load_2d(sA + st * A_STAGE, &tmG, &full[st], kc * BK, (int)(row0 + (int64_t)t * TMG));
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = tc::make_idesc_f16(TMG, NQ, 0);
    uint32_t it = 0, ring = 0, tilecount = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int chunk = item / p.n_qblocks;
      const int64_t row0 = chunk * p.rows_per_chunk;
      const int64_t row1 = reid_min64(p.G_local, row0 + p.rows_per_chunk);
      const int ntiles = (int)((row1 - row0 + TMG - 1) / TMG);
      tc::mbar_wait(&bfull, it & 1);
      tc::fence_after_sync();
      for (int t = 0; t < ntiles; ++t, ++tilecount) {
        const uint32_t buf = tilecount & 1, bph = (tilecount >> 1) & 1;
        tc::mbar_wait(&tempty[buf], bph ^ 1);          // epilogue has drained this accumulator
        tc::fence_after_sync();
        for (int kc = 0; kc < p.kchunks; ++kc, ++ring) {
          const int st = ring % p.stages; const uint32_t ph = (ring / p.stages) & 1;
          tc::mbar_wait(&full[st], ph);
          tc::fence_after_sync();
          const uint64_t ad = tc::make_smem_desc_sw128(tc::smem_u32(sA + st * A_STAGE));
          const uint64_t bd = tc::make_smem_desc_sw128(tc::smem_u32(sB + kc * B_CHUNK));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            tc::mma_f16_ss(tmem_base + buf * NQ, tc::advance_desc_k(ad, k), tc::advance_desc_k(bd, k), idesc,
                           (kc | k) != 0);
          tc::mma_commit(&empty[st]);                   // frees the gallery stage when the MMAs retire
        }
        tc::mma_commit(&tfull[buf]);                    // accumulator complete -> epilogue
      }
      tc::mma_commit(&bempty);                          // query block no longer read
    }
  } else if (warp >= EPI_WARP0) {
    // ------------------------------------------------------------------ epilogue (128 threads)
    const int quad = warp & 3;                          // TMEM lane quadrant of this warp
    const int et = threadIdx.x - EPI_WARP0 * 32;        // 0..127
    const uint32_t tmem_q = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t row_bits = (uint32_t)(quad * 32 + lane) << 8;
    float* my_qs = s_qs + quad * QCAP;
    uint32_t* my_qm = s_qm + quad * QCAP;
    if (et == 0) {
      sh_views.es = es; sh_views.thr = s_thr; sh_views.hist32 = s_hist32; sh_views.list16 = s_list16;
      sh_views.q_s = s_qs; sh_views.q_m = s_qm;
    }
    for (int i = et; i < NQ * p.pcap / 2; i += 128) s_hist32[i] = 0;
    epi_bar();
    const EpiShared& sh = sh_views;
    uint32_t tilecount = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int chunk = item / p.n_qblocks, qb = item % p.n_qblocks;
      const int64_t row0 = chunk * p.rows_per_chunk;
      const int64_t row1 = reid_min64(p.G_local, row0 + p.rows_per_chunk);
      const int ntiles = (int)((row1 - row0 + TMG - 1) / TMG);
      const int64_t q0 = (int64_t)qb * NQ;
      // ---- item setup: per-query state
      {
        const int64_t q = q0 + et;
        const bool live = q < p.Q;
        const int np = live ? min(p.n_pos[q], p.Pmax) : 0;
        es->s_qcode[et] = live ? p.q_code[q] : -2;
        es->s_npos[et] = np;
        es->s_thrlow[et] = np > 0 ? p.pos_thr[q * p.Pmax + np - 1] : INFINITY;
        es->s_thrtop[et] = live ? -INFINITY : INFINITY;      // padded queries never hit
        es->s_min[et] = live ? -INFINITY : INFINITY;
        es->s_candcnt[et] = 0;
        int he = 0;
        if (live) for (int e = 0; e < p.E; ++e) he |= (p.excl[q * p.E + e] >= 0);
        es->s_hasexcl[et] = he;
        for (int i = et; i < NQ * KL; i += 128) s_list16[i] = 0xFF80;   // bf16 -inf
        for (int i = et; i < NQ * p.pcap; i += 128) {
          const int ql = i / p.pcap, j = i % p.pcap;
          const int64_t qq = q0 + ql;
          s_thr[i] = (qq < p.Q && j < p.Pmax) ? p.pos_thr[qq * p.Pmax + j] : -INFINITY;
        }
      }
      epi_bar();
      // ---- tiles
      for (int t = 0; t < ntiles; ++t, ++tilecount) {
        const uint32_t buf = tilecount & 1, bph = (tilecount >> 1) & 1;
        const int tile_row0 = (int)(row0 + (int64_t)t * TMG);
        const int grow_local = tile_row0 + quad * 32 + lane;
        const bool valid = grow_local < row1;
        const int gcode = valid ? p.g_code[grow_local] : -3;
        tc::mbar_wait(&tfull[buf], bph);
        tc::fence_after_sync();
#pragma unroll 1
        for (int ph = 0; ph < 4; ++ph) {
          const int cg = (quad + ph) & 3;                 // rotated column group: exclusive per warp
          uint32_t r[32];
          tc::tmem_ld_x32(tmem_q + buf * NQ + cg * 32, r);
          tc::tmem_wait_ld();
          int qn = 0;                                      // queued hits of this warp
#pragma unroll
          for (int i4 = 0; i4 < 32; i4 += 4) {
            const float4 m4 = *reinterpret_cast<const float4*>(&es->s_min[cg * 32 + i4]);
            const float mm[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float s = __uint_as_float(r[i4 + u]);
              const bool hit = valid && s > mm[u];
              const unsigned c = __ballot_sync(0xffffffffu, hit);
              if (c) {                                     // warp-uniform
                const int ql = cg * 32 + i4 + u;
                if (hit) {
                  const int pos = qn + __popc(c & lt_mask);
                  my_qs[pos] = s;
                  my_qm[pos] = (uint32_t)ql | row_bits | ((gcode != es->s_qcode[ql]) ? M_NOTPOS : 0u);
                }
                qn += __popc(c);
                if (qn > QCAP - 32) { epi_drain(&sh, &p, quad, qn, lane, tile_row0, q0, chunk); qn = 0; }
              }
            }
          }
          if (qn) epi_drain(&sh, &p, quad, qn, lane, tile_row0, q0, chunk);
          epi_bar();
        }
        tc::fence_before_sync();
        if (lane == 0) tc::mbar_arrive(&tempty[buf]);
        if (((t + 1) % FLUSH_TILES) == 0) { epi_flush_hist(sh, p, et, q0); epi_bar(); }
      }
      // ---- item flush
      epi_flush_hist(sh, p, et, q0);
      if (q0 + et < p.Q) p.cand_count[(q0 + et) * p.n_chunks + chunk] = es->s_candcnt[et];
      epi_bar();
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

size_t fused_smem_bytes(int kchunks, int stages, int pcap) {
  return (size_t)kchunks * B_CHUNK + (size_t)stages * A_STAGE + (size_t)NQ * pcap * 4 /*thr*/ +
         (size_t)NQ * pcap * 2 /*hist*/ + (size_t)NQ * KL * 2 /*list*/ + (size_t)4 * QCAP * 8 /*queues*/ +
         sizeof(EpiState) + 1024;
}

}  // namespace

// workspace = the global bucket histogram [Q, 64] int32 (Pmax <= 64)
extern "C" size_t reid_retrieve_fused_workspace_bytes(int64_t Q, int64_t, int) { return (size_t)Q * 64 * sizeof(int32_t); }

extern "C" int reid_retrieve_fused(const void* q_f16, const void* g_f16, const int32_t* q_code, const int32_t* g_code,
                                   const int32_t* excl, int E, const float* pos_thr, const int32_t* n_pos, int64_t Q,
                                   int64_t G_local, int64_t g_offset, int d, int Pmax, int n_chunks, int cand_cap,
                                   int32_t* pos_above, float* cand_score, int32_t* cand_idx, int32_t* cand_count,
                                   void* workspace, size_t workspace_bytes, void* stream) {
  if (!q_f16 || !g_f16 || !q_code || !g_code || !pos_thr || !n_pos || !pos_above || !cand_score || !cand_idx ||
      !cand_count || Q <= 0 || G_local <= 0 || n_chunks <= 0 || cand_cap <= 0 || (E > 0 && !excl) || E < 0)
    return REID_E_INVALID;
  if (d % BK != 0 || d > 512 || Pmax <= 0 || Pmax > 64 || G_local > 0x7fffff00LL) return REID_E_UNSUPPORTED;
  Params p;
  p.q_code = q_code; p.g_code = g_code; p.excl = excl; p.E = E; p.pos_thr = pos_thr; p.n_pos = n_pos;
  p.Q = Q; p.G_local = G_local; p.g_offset = g_offset;
  if (!workspace || workspace_bytes < (size_t)Q * Pmax * sizeof(int32_t)) return REID_E_WORKSPACE;
  p.Pmax = Pmax; p.pcap = (Pmax + 3) / 4 * 4; p.kchunks = d / BK;
  p.n_chunks = n_chunks; p.n_qblocks = (int)((Q + NQ - 1) / NQ); p.cand_cap = cand_cap;
  const int64_t rpc = (G_local + n_chunks - 1) / n_chunks;
  p.rows_per_chunk = (rpc + TMG - 1) / TMG * TMG;
  p.hist = (int32_t*)workspace; p.cand_score = cand_score; p.cand_idx = cand_idx; p.cand_count = cand_count;
  int stages = MAX_STAGES;
  const size_t smem_max = 227 * 1024;
  while (stages > 2 && fused_smem_bytes(p.kchunks, stages, p.pcap) > smem_max) --stages;
  if (fused_smem_bytes(p.kchunks, stages, p.pcap) > smem_max) return REID_E_UNSUPPORTED;
  p.stages = stages;
  const size_t smem = fused_smem_bytes(p.kchunks, stages, p.pcap);
  CUtensorMap tmG, tmQ;
  if (!tc_host::make_map_f16(&tmG, g_f16, G_local, d, TMG) || !tc_host::make_map_f16(&tmQ, q_f16, Q, d, NQ))
    return REID_E_CUDA;
  if (cudaFuncSetAttribute(retrieve_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return REID_E_CUDA;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return REID_E_CUDA;
  const int n_items = p.n_qblocks * n_chunks;
  const int grid = n_items < sms ? n_items : sms;
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(workspace, 0, (size_t)Q * Pmax * sizeof(int32_t), st) != cudaSuccess) return REID_E_CUDA;
  retrieve_fused_kernel<<<grid, THREADS, smem, st>>>(tmG, tmQ, p);
  REID_CHECK_LAUNCH();
  hist_to_above_kernel<<<(int)reid_min64((Q + 255) / 256, 148 * 8), 256, 0, st>>>(p.hist, n_pos, Q, Pmax, pos_above);
  REID_CHECK_LAUNCH();
  return REID_OK;
}
