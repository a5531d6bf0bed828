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
//     positive threshold).  Only columns with a hit run the slow path:
//       (a) counting: lane p holds positive threshold p of the query (read from the spare TMEM
//           columns 256.., written there once per item) and counts the hit scores above it;
//           counters live in shared memory -> pos_above[q, p] at the end of the item;
//       (b) top list: a 32-entry running list per query (one entry per lane) gives the threshold
//           above which rows are appended to the query's candidate buffer in global memory.
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
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t THR_COL0 = 256;      // thresholds p in [0,32) at column 256+q, p in [32,64) at 384+q
static_assert(KL == 32, "one list entry per lane");

struct Params {
  const int32_t* q_code; const int32_t* g_code; const int32_t* excl; int E;
  const float* pos_thr; const int32_t* n_pos;
  int64_t Q, G_local, g_offset;
  int Pmax, pcap, kchunks, stages, n_chunks, n_qblocks, cand_cap;
  int64_t rows_per_chunk;
  int32_t* pos_above; float* cand_score; int32_t* cand_idx; int32_t* cand_count;
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

// slow path of one (warp, query column): called warp-uniformly.
__device__ __noinline__ void epi_slow(EpiState* es, int* s_cnt, float* s_list, const Params* pp, uint32_t tmem_thr,
                                      int ql, int64_t qglob, int chunk, float s, bool valid, int gcode,
                                      int grow_local, int lane) {
  const Params& p = *pp;
  bool ok = valid;
  if (es->s_hasexcl[ql]) {                        // same-image mask (eval_mm_protocol.py:408-418)
    const int32_t gidx = (int32_t)(p.g_offset + grow_local);
    for (int e = 0; e < p.E; ++e) ok = ok && (p.excl[qglob * p.E + e] != gidx);
  }
  // (a) rows ranked above the query's positives (non-positives only; positives are ordered exactly
  //     among themselves on the host side of the formula: rank_j = 1 + above_j + j)
  const int npos = es->s_npos[ql];
  const unsigned cm = __ballot_sync(0xffffffffu, ok && gcode != es->s_qcode[ql] && s > es->s_thrlow[ql]);
  if (cm) {
    for (int pb = 0; pb < npos; pb += 32) {
      const float t = __uint_as_float(tc::tmem_ld_x1(tmem_thr + (uint32_t)((pb >> 5) * 128 + ql)));
      tc::tmem_wait_ld();
      int cnt = 0;
      unsigned m = cm;
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        cnt += (__shfl_sync(0xffffffffu, s, src) > t) ? 1 : 0;
      }
      if (cnt) s_cnt[ql * p.pcap + pb + lane] += cnt;   // t = +inf beyond npos, so no stray writes
    }
  }
  // (b) running top list -> candidate buffer
  float thr = es->s_thrtop[ql];
  unsigned tm = __ballot_sync(0xffffffffu, ok && s > thr);
  if (tm) {
    float lv = s_list[ql * KL + lane];
    int cc = es->s_candcnt[ql];
    const int64_t cbase = (qglob * p.n_chunks + chunk) * (int64_t)p.cand_cap;
    while (tm) {
      const int src = __ffs(tm) - 1;
      tm &= tm - 1;
      const float v = __shfl_sync(0xffffffffu, s, src);
      const int gi = __shfl_sync(0xffffffffu, grow_local, src);
      if (v > thr) {                                 // thr == min(list) is warp-uniform
        if (lane == 0 && cc < p.cand_cap) { p.cand_score[cbase + cc] = v; p.cand_idx[cbase + cc] = gi; }
        ++cc;
        const unsigned holders = __ballot_sync(0xffffffffu, lv == thr);
        if (lane == __ffs(holders) - 1) lv = v;
        thr = warp_min(lv);
      }
    }
    s_list[ql * KL + lane] = lv;
    if (lane == 0) {
      es->s_thrtop[ql] = thr;
      es->s_candcnt[ql] = cc;
      es->s_min[ql] = fminf(thr, es->s_thrlow[ql]);
    }
  }
  __syncwarp();
}

__global__ void __launch_bounds__(THREADS, 1)
retrieve_fused_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmQ,
                      const __grid_constant__ Params prm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  const Params& p = prm;
  uint8_t* sB = smem;                                         // [kchunks][B_CHUNK] resident query block
  uint8_t* sA = sB + p.kchunks * B_CHUNK;                      // [stages][A_STAGE]  gallery ring
  int* s_cnt = reinterpret_cast<int*>(sA + p.stages * A_STAGE);   // [NQ][pcap]
  float* s_list = reinterpret_cast<float*>(s_cnt + NQ * p.pcap);  // [NQ][KL]
  EpiState* es = reinterpret_cast<EpiState*>(s_list + NQ * KL);
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
        const float tl = np > 0 ? p.pos_thr[q * p.Pmax + np - 1] : INFINITY;
        es->s_thrlow[et] = tl;
        es->s_thrtop[et] = live ? -INFINITY : INFINITY;      // padded queries never hit
        es->s_min[et] = live ? -INFINITY : INFINITY;
        es->s_candcnt[et] = 0;
        int he = 0;
        if (live) for (int e = 0; e < p.E; ++e) he |= (p.excl[q * p.E + e] >= 0);
        es->s_hasexcl[et] = he;
        for (int i = et; i < NQ * p.pcap; i += 128) s_cnt[i] = 0;
        for (int i = et; i < NQ * KL; i += 128) s_list[i] = -INFINITY;
        // positive thresholds -> spare TMEM columns of this warp's quadrant: lane = threshold index
        for (int ql = 0; ql < NQ; ++ql) {
          const int64_t qq = q0 + ql;
          const int npq = (qq < p.Q) ? min(p.n_pos[qq], p.Pmax) : 0;
          const float t0 = (lane < npq) ? p.pos_thr[qq * p.Pmax + lane] : INFINITY;
          tc::tmem_st_x1(tmem_q + THR_COL0 + ql, __float_as_uint(t0));
          if (p.Pmax > 32) {
            const float t1 = (32 + lane < npq) ? p.pos_thr[qq * p.Pmax + 32 + lane] : INFINITY;
            tc::tmem_st_x1(tmem_q + THR_COL0 + 128 + ql, __float_as_uint(t1));
          }
        }
        tc::tmem_wait_st();
      }
      epi_bar();
      // ---- tiles
      for (int t = 0; t < ntiles; ++t, ++tilecount) {
        const uint32_t buf = tilecount & 1, bph = (tilecount >> 1) & 1;
        const int grow_local = (int)(row0 + (int64_t)t * TMG) + quad * 32 + lane;
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
#pragma unroll
          for (int i4 = 0; i4 < 32; i4 += 4) {
            const float4 m4 = *reinterpret_cast<const float4*>(&es->s_min[cg * 32 + i4]);
            const float mm[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float s = __uint_as_float(r[i4 + u]);
              if (__any_sync(0xffffffffu, valid && s > mm[u])) {
                const int ql = cg * 32 + i4 + u;
                epi_slow(es, s_cnt, s_list, &p, tmem_q + THR_COL0, ql, q0 + ql, chunk, s, valid, gcode, grow_local, lane);
              }
            }
          }
          epi_bar();
        }
        tc::fence_before_sync();
        if (lane == 0) tc::mbar_arrive(&tempty[buf]);
      }
      // ---- item flush
      for (int i = et; i < NQ * p.pcap; i += 128) {
        const int ql = i / p.pcap, pi = i % p.pcap;
        const int c = s_cnt[i];
        if (c && pi < es->s_npos[ql]) atomicAdd(&p.pos_above[(q0 + ql) * p.Pmax + pi], c);
      }
      if (q0 + et < p.Q) p.cand_count[(q0 + et) * p.n_chunks + chunk] = es->s_candcnt[et];
      epi_bar();
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

size_t fused_smem_bytes(int kchunks, int stages, int pcap) {
  return (size_t)kchunks * B_CHUNK + (size_t)stages * A_STAGE + (size_t)NQ * pcap * 4 + (size_t)NQ * KL * 4 +
         sizeof(EpiState) + 1024;
}

}  // namespace

extern "C" size_t reid_retrieve_fused_workspace_bytes(int64_t, int64_t, int) { return 0; }

extern "C" int reid_retrieve_fused(const void* q_f16, const void* g_f16, const int32_t* q_code, const int32_t* g_code,
                                   const int32_t* excl, int E, const float* pos_thr, const int32_t* n_pos, int64_t Q,
                                   int64_t G_local, int64_t g_offset, int d, int Pmax, int n_chunks, int cand_cap,
                                   int32_t* pos_above, float* cand_score, int32_t* cand_idx, int32_t* cand_count,
                                   void* workspace, size_t workspace_bytes, void* stream) {
  (void)workspace; (void)workspace_bytes;
  if (!q_f16 || !g_f16 || !q_code || !g_code || !pos_thr || !n_pos || !pos_above || !cand_score || !cand_idx ||
      !cand_count || Q <= 0 || G_local <= 0 || n_chunks <= 0 || cand_cap <= 0 || (E > 0 && !excl) || E < 0)
    return REID_E_INVALID;
  if (d % BK != 0 || d > 512 || Pmax <= 0 || Pmax > 64 || G_local > 0x7fffff00LL) return REID_E_UNSUPPORTED;
  Params p;
  p.q_code = q_code; p.g_code = g_code; p.excl = excl; p.E = E; p.pos_thr = pos_thr; p.n_pos = n_pos;
  p.Q = Q; p.G_local = G_local; p.g_offset = g_offset;
  p.Pmax = Pmax; p.pcap = (Pmax + 3) / 4 * 4; p.kchunks = d / BK;
  p.n_chunks = n_chunks; p.n_qblocks = (int)((Q + NQ - 1) / NQ); p.cand_cap = cand_cap;
  const int64_t rpc = (G_local + n_chunks - 1) / n_chunks;
  p.rows_per_chunk = (rpc + TMG - 1) / TMG * TMG;
  p.pos_above = pos_above; p.cand_score = cand_score; p.cand_idx = cand_idx; p.cand_count = cand_count;
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
  retrieve_fused_kernel<<<grid, THREADS, smem, (cudaStream_t)stream>>>(tmG, tmQ, p);
  REID_CHECK_LAUNCH();
  return REID_OK;
}
