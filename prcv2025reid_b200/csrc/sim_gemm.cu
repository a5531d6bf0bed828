// sim_gemm.cu -- K3: cosine-similarity product S = Q^ G^T as a TMA-fed tcgen05 GEMM (unfused).
//
// Replaces eval_mm_protocol.py:50-53 `cosine_sim` (a @ b.T).  fp16 operands (unit-norm rows never
// leave the fp16 normal range; 8x finer than bf16 at the same tensor rate), fp32 accumulation in
// TMEM.  Layout: the GALLERY tile is the MMA M operand (128 rows -> 128 TMEM lanes), the QUERY
// tile the N operand (128 columns), so in the epilogue lane l of a warp owns one gallery row and a
// TMEM column is one query -- for a fixed query the 32 lanes of a warp write 32 consecutive floats
// of S[q, :], a coalesced 128-byte store.  The fused retrieval kernel (retrieve_fused.cu) uses
// the same operand roles; this kernel materialises S and is the drop-in for `cosine_sim` and
// the debug / verification mode of the fused path.
// Roofline: tensor (2*Q*G*d flop) but the fp32 S store makes it HBM-bound for d = 512
// (256 flop per output byte vs a ridge of ~212 flop/B): algorithmic bytes = 4*Q*G written (+ operands).  It is NOT the
// benchmarked path -- the ranking step never materialises S (retrieve_fused.cu) -- but the drop-in for `cosine_sim`
// and the verification kernel of the fused path.  One tile per CTA, 2 ring stages (65 KB) so that THREE CTAs share an SM
// (3 x 128 TMEM columns): the store epilogue of one tile overlaps the loads / MMAs of the next ones.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int BM = 128;      // gallery rows per CTA tile (MMA M)
constexpr int BN = 128;      // queries per CTA tile (MMA N)
constexpr int BK = 64;       // one 128-byte swizzle atom of fp16
constexpr int STAGES = 2;
constexpr int A_BYTES = BM * BK * 2;
constexpr int B_BYTES = BN * BK * 2;

__global__ void __launch_bounds__(128, 3)
sim_gemm_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmQ,
                float* __restrict__ S, int64_t Q, int64_t G, int64_t ldS, int kchunks) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment in the shared window
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                         // [STAGES][A_BYTES]
  uint8_t* sB = smem + STAGES * A_BYTES;      // [STAGES][B_BYTES]
  __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES], tfull;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    tc::mbar_init(&tfull, 1);
    tc::fence_barrier_init();
    tc::prefetch_tensormap(&tmG); tc::prefetch_tensormap(&tmQ);
  }
  if (warp == 2) tc::tmem_alloc(&tmem_base_s, BN);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  const int g0 = blockIdx.x * BM, q0 = blockIdx.y * BN;

  if (warp == 0 && lane == 0) {                 // ---- TMA producer
    for (int kc = 0; kc < kchunks; ++kc) {
      const int st = kc % STAGES; const uint32_t ph = (kc / STAGES) & 1;
      tc::mbar_wait(&empty[st], ph ^ 1);
      tc::mbar_arrive_expect_tx(&full[st], A_BYTES + B_BYTES);
      tc::tma_load_2d(sA + st * A_BYTES, &tmG, &full[st], kc * BK, g0);
      tc::tma_load_2d(sB + st * B_BYTES, &tmQ, &full[st], kc * BK, q0);
    }
  } else if (warp == 1 && lane == 0) {          // ---- MMA issuer
    constexpr uint32_t idesc = tc::make_idesc_f16(BM, BN, 0);
    for (int kc = 0; kc < kchunks; ++kc) {
      const int st = kc % STAGES; const uint32_t ph = (kc / STAGES) & 1;
      tc::mbar_wait(&full[st], ph);
      tc::fence_after_sync();
      const uint64_t ad = tc::make_smem_desc_sw128(tc::smem_u32(sA + st * A_BYTES));
      const uint64_t bd = tc::make_smem_desc_sw128(tc::smem_u32(sB + st * B_BYTES));
#pragma unroll
      for (int k = 0; k < BK / 16; ++k)
        tc::mma_f16_ss(tmem_base, tc::advance_desc_k(ad, k), tc::advance_desc_k(bd, k), idesc, (kc | k) != 0);
      tc::mma_commit(&empty[st]);
    }
    tc::mma_commit(&tfull);
  }
  __syncwarp();
  // ---- epilogue: all four warps, warp w owns TMEM lanes 32w..32w+31 (gallery rows)
  tc::mbar_wait(&tfull, 0);
  tc::fence_after_sync();
  const int64_t g = (int64_t)g0 + warp * 32 + lane;
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += 32) {
    uint32_t r[32];
    tc::tmem_ld_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, r);
    tc::tmem_wait_ld();
    if (g < G) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int64_t q = (int64_t)q0 + c0 + i;
        if (q < Q) S[q * ldS + g] = __uint_as_float(r[i]);
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem_base, BN);
}

}  // namespace

extern "C" int reid_sim_gemm(const void* q_f16, const void* g_f16, float* S, int64_t Q, int64_t G, int d, int64_t ldS,
                             void* stream) {
  if (!q_f16 || !g_f16 || !S || Q <= 0 || G <= 0 || d <= 0 || d % BK != 0 || ldS < G) return REID_E_INVALID;
  CUtensorMap tmG, tmQ;
  if (!tc_host::make_map_f16(&tmG, g_f16, G, d, BM) || !tc_host::make_map_f16(&tmQ, q_f16, Q, d, BN)) return REID_E_CUDA;
  const int smem = STAGES * (A_BYTES + B_BYTES) + 1024;
  if (cudaFuncSetAttribute(sim_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
    return REID_E_CUDA;
  dim3 grid((unsigned)((G + BM - 1) / BM), (unsigned)((Q + BN - 1) / BN));
  sim_gemm_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(tmG, tmQ, S, Q, G, ldS, d / BK);
  REID_CHECK_LAUNCH();
  return REID_OK;
}
