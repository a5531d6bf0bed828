// sdm_tc.cu -- SDM loss forward / backward on the 5th-generation tensor cores (bf16 inputs, large batch).
//
// Replaces models/sdm_loss.py:13-149 (`sdm_loss_stable`) and its autograd backward for the large-batch
// configuration (BASELINE C5: P x K = 64 x 8 -> N = M = 512, d = 512, bf16, up to 16 modality pairs per
// launch).  With bf16 inputs the reference normalises in bf16 (sdm_loss.py:31-32) and only then converts
// to fp32 (:75), so the operands of S = q^ g^T (:86) are exactly representable in bf16: a kind::f16
// tcgen05.mma with bf16 operands and fp32 accumulation computes the reference's S up to summation order.
//
// Three launches per step, no host synchronisation, no cooperative grid barrier:
//   tc_prep_kernel  normalise rows (reference bf16 rounding), write each operand twice -- row-major-K image
//                   and transposed image -- directly in the UMMA K-major 128B-swizzle layout, so that every
//                   operand tile is ONE contiguous cp.async.bulk (no tensor maps, any number of pairs); a third
//                   slice of its grid turns a dense y into the positive-mask bit rows both kernels read.
//   tc_fwd_kernel   CTA = (pair, side, 128-row block).  side 0: rows = qry, columns = gal; side 1 = the transposed
//                   problem.  The whole row block x all columns (<= 512) accumulates in TMEM (128 lanes x 512
//                   fp32 columns = all of it), so the epilogue thread that owns a TMEM lane owns a full row of S:
//                   S/tau, clamp (:94), log-sum-exp and the positive statistics of _one_side_ce (:34-57) need no
//                   cross-thread reduction.  The last CTA of a pair (completion counter) reduces the per-row
//                   cross-entropies to the loss and evaluates the guards (:60-68, :79-81, :89-91, :105-106, :142-147).
//   tc_bwd_kernel   CTA = (pair, side, 128-row block).  side 0: dq^ = dS g^ (K = gal rows), side 1: dg^ = dS^T q^.
//                   Sixteen producer warps form dL/dS tiles from S, the row / column LSE and the mask bits as one
//                   scaled fp16 plane (see REID_SDM_DS_F16) and store them in the swizzled operand layout; the MMA warp
//                   multiplies them with the transposed operand image streamed by bulk copies; the accumulator
//                   (128 rows x d columns) again gives every epilogue thread a full output row, so the
//                   normalisation Jacobian (I - x^ x^T)/den is applied in registers and the gradient is written
//                   once, in bf16.
// Rooflines: algorithmic HBM bytes per pair fwd+bwd = 3*(N+M)*d*2 + 2*N*M*4; tensor work 3 * 2*N*M*d flop
// (+ the lo pass of the backward).  See DESIGN.md section 6.
#include "sdm_common.cuh"
#include "tc_common.cuh"

// -DREID_SDM_PAIR=1 builds the cta_group::2 pair variant of the forward / backward kernels (an experiment that is kept
// parity-tested but is slower on C5: see tc_forward).  Nothing in this file reads the environment.
#ifndef REID_SDM_PAIR
#define REID_SDM_PAIR 0
#endif

// -DREID_SDM_DS_F16=0 keeps dL/dS as two bf16 planes (hi + lo, 16 mantissa bits, two MMAs per K step) in the backward; the
// default is ONE fp16 plane of dL/dS for grad_out = 1, scaled by 2^12 (|value| <= 8192; 11 mantissa bits, far inside the
// bf16 rounding of the result).  kind::f16 wants A and B in the SAME 16-bit type (a mixed fp16 x bf16 descriptor is an
// illegal instruction on sm_100a), so the prep kernel writes the TRANSPOSED operand images -- read by the backward only --
// in fp16: the normalised bf16 values (8 mantissa bits, |x| <= 1) are exact in fp16 down to 2^-14 and lose at most 2^-25
// absolutely below that.
#ifndef REID_SDM_DS_F16
#define REID_SDM_DS_F16 1
#endif
// -DREID_SDM_PDL=1 launches the forward / backward with programmatic dependent launch (their prologues then overlap the
// tail of the kernel before them).  Parity-tested, but OFF by default: measured on C5 as a CUDA graph the step takes 69.3-69.8 us
// with it against 68.2 us without (profiles/r02t_sdm_ab.txt) -- the early CTAs compete with the running kernel for issue slots
// and the launch gaps inside a graph are already short.
#ifndef REID_SDM_PDL
#define REID_SDM_PDL 0
#endif

namespace sdm {
namespace {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float bf16r(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return v;
}
// exp(x) for |x| <= ~40 (the clamp :94 bounds every argument): one FMUL + MUFU.EX2, relative error ~2^-22
__device__ __forceinline__ float fast_exp(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
  return y;
}
// REID_SDM_TIMING build only (scripts/sdm_phase_times.py): phase time stamps (ns) of the forward CTA (row block 0,
// side 0) of every pair in hdr words 80..95
#ifdef REID_SDM_TIMING
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define SDM_STAMP(cond, slot) do { if (cond) reinterpret_cast<unsigned long long*>(hdr_i + 80)[slot] = gtime(); } while (0)
#else
#define SDM_STAMP(cond, slot) do { } while (0)
#endif
__device__ __forceinline__ void named_bar(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// =============================================================================================== prep
#ifndef REID_PREP_ROWS
#define REID_PREP_ROWS 16
#endif
constexpr int PREP_ROWS = REID_PREP_ROWS;   // rows per CTA (8-row K chunks of the transposed image)

constexpr int PREP_THREADS = 256;

// dense form: y [N][M] fp32 -> positive masks as bit rows, ybits [N][16 words] (bit c of row i: y[i][c] > 0) and
// ybitsT [M][16 words] (bit r of row j: y[r][j] > 0), with coalesced 128-byte loads and one ballot / one bit per load.
// Runs as the third y-slice of the prep grid, next to the normalisation CTAs (the forward used to form the masks
// itself, one row per thread: 12 us on its critical path at N = M = 512).
__device__ void prep_mask_bits(const reid_sdm_pair& P, int d) {
  const int N = P.N, M = P.M;
  const TcLayout L = tc_layout(N, M, d);
  uint8_t* bytes = reinterpret_cast<uint8_t*>(tc_base(P.saved));
  uint32_t* ybits = reinterpret_cast<uint32_t*>(bytes + L.ybits);
  uint32_t* ybitsT = reinterpret_cast<uint32_t*>(bytes + L.ybitsT);
  const int lane = threadIdx.x & 31;
  const int wid = blockIdx.x * (PREP_THREADS / 32) + (threadIdx.x >> 5);
  const int nw = gridDim.x * (PREP_THREADS / 32);
  const int wm = (M + 31) >> 5, wn = (N + 31) >> 5;                  // words per bit row (<= 16)
  for (int r = wid; r < N; r += nw) {                                // ybits: one warp per row of y
    const float* yr = P.y + (size_t)r * M;
    float v[16];
#pragma unroll
    for (int w = 0; w < 16; ++w) v[w] = (w < wm && 32 * w + lane < M) ? yr[32 * w + lane] : 0.f;
    uint32_t mine = 0;
#pragma unroll
    for (int w = 0; w < 16; ++w) {
      const uint32_t b = __ballot_sync(0xffffffffu, v[w] > 0.f);
      if (lane == w) mine = b;
    }
    if (lane < 16) ybits[(size_t)r * 16 + lane] = mine;
  }
  for (int task = wid; task < wm * wn; task += nw) {                 // ybitsT: one warp per (32 columns, 32 rows) block
    const int cg = task / wn, w = task % wn;
    const int col = 32 * cg + lane;
    uint32_t word = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {                                     // 16 loads in flight at a time
      float v[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int r = 32 * w + 16 * h + e;
        v[e] = (r < N && col < M) ? P.y[(size_t)r * M + col] : 0.f;
      }
#pragma unroll
      for (int e = 0; e < 16; ++e) word |= (v[e] > 0.f ? 1u : 0u) << (16 * h + e);
    }
    if (col < M) ybitsT[(size_t)col * 16 + w] = word;
  }
  // (words beyond wn / wm of a bit row are never read: the forward masks with the validity words, zero there)
}

__global__ void __launch_bounds__(PREP_THREADS, 5)
tc_prep_kernel(const __grid_constant__ Batch batch, int d, float eps) {
  tc::grid_launch_dependents();                      // the forward's CTAs may take their SMs as soon as every prep CTA runs
  const reid_sdm_pair& P = batch.p[blockIdx.z];
  if (blockIdx.y == 2) {                             // mask slice of the grid (dense form only)
    if (P.y) prep_mask_bits(P, d);
    return;
  }
  const int which = blockIdx.y;                      // 0 = qry, 1 = gal
  const int R = which ? P.M : P.N;
  const int Rp = round_up(R, 128);
  const int r0 = blockIdx.x * PREP_ROWS;
  if (r0 >= Rp) return;
  const TcLayout L = tc_layout(P.N, P.M, d);
  float* base = tc_base(P.saved);
  uint8_t* bytes = reinterpret_cast<uint8_t*>(base);
  const bf16* x = reinterpret_cast<const bf16*>(which ? P.gal : P.qry);
  float* den = base + (which ? L.den_g : L.den_q);
  uint8_t* img = bytes + (which ? L.gn : L.qn);
  uint8_t* imgT = bytes + (which ? L.gnt : L.qnt);
  int* hdr_i = reinterpret_cast<int*>(base + L.hdr);
  extern __shared__ __align__(16) uint8_t prep_smem[];
  bf16* tile = reinterpret_cast<bf16*>(prep_smem);   // [PREP_ROWS][d + 8]
  __shared__ int s_bad;
  const int ld = d + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) s_bad = 0;
  if (blockIdx.x == 0 && which == 0 && threadIdx.x == 0) hdr_i[4] = 0;     // forward completion counter
  __syncthreads();
  // label form: rows whose valid byte is 0 take no part (models/model.py:570-602): their normalised image is zero, their
  // denominator 1, they raise no non-finite flag; validity bit rows for the forward / backward go to the header
  const uint8_t* valid = P.y ? nullptr : (which ? P.col_valid : P.row_valid);
  if (blockIdx.x == 0 && threadIdx.x < 16) {
    uint32_t w = 0;
    for (int e = 0; e < 32; ++e) {
      const int r = threadIdx.x * 32 + e;
      if (r < R && (!valid || valid[r])) w |= 1u << e;
    }
    hdr_i[(which ? TC_HDR_VALID_G : TC_HDR_VALID_Q) + threadIdx.x] = (int)w;
  }
  const float e_b = bf16r(eps);
  const int nchunk = d >> 3;                         // 16-byte chunks per row (<= 64)
  bool bad = false;
  constexpr int RPW = PREP_ROWS / (PREP_THREADS / 32);   // rows per warp: all their loads are issued before the first use
  uint4 raw[RPW][2];
#pragma unroll
  for (int a = 0; a < RPW; ++a) {
    const int r = r0 + warp + a * (PREP_THREADS / 32);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int c = lane + 32 * u;
      raw[a][u] = make_uint4(0, 0, 0, 0);
      if (r < R && c < nchunk) raw[a][u] = *reinterpret_cast<const uint4*>(x + (size_t)r * d + c * 8);
    }
  }
#pragma unroll
  for (int a = 0; a < RPW; ++a) {
    const int rr = warp + a * (PREP_THREADS / 32);
    const int r = r0 + rr;
    const bool live = r < R && (!valid || valid[r]);
    float ss = 0.f;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float f[8];
      unpack8(raw[a][u], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) ss = fmaf(f[e], f[e], ss);
    }
    ss = warp_sum(ss);
    const float dn = live ? fmaxf(bf16r(sqrtf(ss)), e_b) : 1.f;   // F.normalize in bf16: max(||x||, eps), :31-32
    if (lane == 0 && r < R) den[r] = dn;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int c = lane + 32 * u;
      if (c >= nchunk) continue;
      float f[8];
      unpack8(raw[a][u], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        f[e] = live ? bf16r(__fdiv_rn(f[e], dn)) : 0.f;
        bad |= !isfinite(f[e]);
      }
      const uint4 o = pack8(f);
      const int kb = c >> 3, ch = c & 7;
      *reinterpret_cast<uint4*>(img + (size_t)kb * ((size_t)Rp * 128) + (size_t)(r >> 3) * 1024 + (r & 7) * 128 +
                                ((ch ^ (r & 7)) << 4)) = o;
      *reinterpret_cast<uint4*>(tile + rr * ld + c * 8) = o;
    }
  }
  if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(&s_bad, 1);
  __syncthreads();
  if (threadIdx.x == 0) hdr_i[8 + which * 32 + blockIdx.x] = s_bad;          // sdm_loss.py:79-81
  // transposed image: rows = feature index c, K = row index r (this CTA: 4 chunks of 8 rows)
  for (int c = threadIdx.x; c < d; c += PREP_THREADS) {
#pragma unroll
    for (int q = 0; q < PREP_ROWS / 8; ++q) {
      uint4 o;
      bf16* ob = reinterpret_cast<bf16*>(&o);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (REID_SDM_DS_F16) reinterpret_cast<__half*>(ob)[e] = __float2half_rn(__bfloat162float(tile[(q * 8 + e) * ld + c]));
        else ob[e] = tile[(q * 8 + e) * ld + c];
      }
      const int k0 = r0 + q * 8;
      const int kb = k0 >> 6, ch = (k0 & 63) >> 3;
      *reinterpret_cast<uint4*>(imgT + (size_t)kb * ((size_t)d * 128) + (size_t)(c >> 3) * 1024 + (c & 7) * 128 +
                                ((ch ^ (c & 7)) << 4)) = o;
    }
  }
}

// =============================================================================================== forward
constexpr int FWD_EPI_WARPS = 16;     // four warps per TMEM lane quadrant, a quarter of the columns each
constexpr int FWD_EPI_THREADS = FWD_EPI_WARPS * 32;
constexpr int FWD_THREADS = 64 + FWD_EPI_THREADS;   // warp 0: bulk-copy producer, warp 1: MMA + TMEM, warps 2-17: epilogue
constexpr int A_TILE = 128 * 128;        // 128 rows x 64 bf16 = 16 KB
constexpr int B_TILE_MAX = 512 * 128;    // up to 512 rows x 64 bf16 = 64 KB
constexpr int FWD_STAGE = A_TILE + B_TILE_MAX;
constexpr int FWD_STAGES = 2;
constexpr size_t FWD_SMEM = (size_t)FWD_STAGES * FWD_STAGE + 1024;

__device__ __forceinline__ uint32_t tmem_cols_for(int c) { return c <= 32 ? 32u : c <= 64 ? 64u : c <= 128 ? 128u : c <= 256 ? 256u : 512u; }

// PAIR = true: the two CTAs of a cluster take row blocks 2p and 2p+1 of one (pair, side) and issue ONE
// cta_group::2 MMA per step (M = 256: 128 accumulator lanes in each CTA).  Each CTA loads its own A tile and only
// HALF of the column operand (the leader's MMA reads both halves), so the operand stream per CTA drops from 80 KB to
// 48 KB per K block and the ring is 4 deep instead of 2.  Tiles come through per-pair tensor maps over the operand
// images (cp.async.bulk.tensor ... cta_group::2 signals the LEADER's barrier from both CTAs).
struct PairMaps {
  CUtensorMap q[REID_SDM_MAX_PAIRS];      // image Qn viewed as [Np * d/64][64] bf16 (one row = one swizzled 128-byte row)
  CUtensorMap g[REID_SDM_MAX_PAIRS];
  CUtensorMap qt[REID_SDM_MAX_PAIRS];     // transposed images QnT / GnT viewed as [d * Np/64][64] (backward)
  CUtensorMap gt[REID_SDM_MAX_PAIRS];
};
constexpr int FWDP_STAGE = A_TILE + 2 * A_TILE;      // own A tile + this CTA's half of up to two 256-column chunks
constexpr int FWDP_STAGES = 4;
constexpr size_t FWDP_SMEM = (size_t)FWDP_STAGES * FWDP_STAGE + 1024;

template <bool PAIR>
__global__ void __launch_bounds__(FWD_THREADS, 1)
tc_fwd_kernel(const __grid_constant__ Batch batch, const __grid_constant__ PairMaps maps, int d, float tau_eff) {
  constexpr int STAGES = PAIR ? FWDP_STAGES : FWD_STAGES;
  constexpr int STAGE = PAIR ? FWDP_STAGE : FWD_STAGE;
  const reid_sdm_pair& P = batch.p[blockIdx.z];
  const int side = blockIdx.y;
  const int N = P.N, M = P.M;
  const int R = side ? M : N, C = side ? N : M;              // rows / columns of this side's view of S
  const int Rp = round_up(R, 128), Cp = round_up(C, 128);
  const int rb = blockIdx.x;
  const uint32_t crank = PAIR ? (uint32_t)(rb & 1) : 0u;     // cluster rank (cluster = 2 consecutive CTAs in x)
  const bool leader = crank == 0;
  if ((PAIR ? (rb & ~1) : rb) * 128 >= R) return;            // (uniform over the pair)
  const bool active = rb * 128 < R;                          // pair mode: the peer of the last odd block only feeds operands
  const TcLayout L = tc_layout(N, M, d);
  float* base = tc_base(P.saved);
  const uint8_t* bytes = reinterpret_cast<const uint8_t*>(base);
  const uint8_t* imgA = bytes + (side ? L.gn : L.qn);
  const uint8_t* imgB = bytes + (side ? L.qn : L.gn);
  int* hdr_i = reinterpret_cast<int*>(base + L.hdr);
  extern __shared__ uint8_t fwd_smem_raw[];
  uint8_t* smem = fwd_smem_raw + ((1024u - (tc::smem_u32(fwd_smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES], accfull;
  __shared__ uint32_t tmem_base_s;
  __shared__ int s_bad, s_last, s_st;
  __shared__ float s_part[FWD_EPI_WARPS / 4 - 1][3][128];
  __shared__ double s_red[FWD_EPI_WARPS][5];
  __shared__ int64_t s_collab[512];                          // label form: the labels of this side's columns
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = d >> 6;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    tc::mbar_init(&accfull, 1);
    tc::fence_barrier_init();
    s_bad = 0; s_last = 0; s_st = 0;
    if (PAIR) { tc::prefetch_tensormap(&maps.q[blockIdx.z]); tc::prefetch_tensormap(&maps.g[blockIdx.z]); }
  }
  const uint32_t ncols = tmem_cols_for(Cp);
  if (warp == 1) { if (PAIR) tc::tmem_alloc_pair(&tmem_base_s, ncols); else tc::tmem_alloc(&tmem_base_s, ncols); }
  tc::fence_before_sync();
  if (PAIR) tc::cluster_sync_all(); else __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  const bool stamp = rb == 0 && side == 0;
  (void)stamp;
  // programmatic dependent launch: everything above ran while the prep kernel was still finishing; the backward's CTAs
  // may start their own prologue on the SMs this grid leaves free
  tc::grid_launch_dependents();
  tc::grid_dependency_wait();
  SDM_STAMP(stamp && threadIdx.x == 0, 0);                   // prologue done

  if (warp == 0 && lane == 0) {
    if (PAIR) {
      const CUtensorMap* mapA = side ? &maps.g[blockIdx.z] : &maps.q[blockIdx.z];
      const CUtensorMap* mapB = side ? &maps.q[blockIdx.z] : &maps.g[blockIdx.z];
      const int rbA = active ? rb : 0;                       // an idle peer still streams (unused) rows: no special cases
      uint32_t tx = 2u * A_TILE;                             // bytes of BOTH CTAs per stage
      for (int n0 = 0; n0 < Cp; n0 += 256) tx += 2u * (uint32_t)(((Cp - n0 < 256 ? Cp - n0 : 256) / 2) * 128);
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES;
        tc::mbar_wait(&empty[s], ((kb / STAGES) & 1) ^ 1);
        if (leader) tc::mbar_arrive_expect_tx(&full[s], tx);
        const uint32_t lbar = tc::leader_addr(&full[s]);
        uint8_t* st = smem + s * STAGE;
        const int rowA = kb * Rp + rbA * 128;
        tc::tma_load_2d_pair(st, mapA, lbar, 0, rowA);
        tc::tma_load_2d_pair(st + A_TILE / 2, mapA, lbar, 0, rowA + 64);
        int c = 0;
        for (int n0 = 0; n0 < Cp; n0 += 256, ++c) {
          const int hr = (Cp - n0 < 256 ? Cp - n0 : 256) / 2;     // this CTA's rows of the chunk: 64 or 128
          const int rowB = kb * Cp + n0 + (int)crank * hr;
          tc::tma_load_2d_pair(st + A_TILE + c * A_TILE, mapB, lbar, 0, rowB);
          if (hr > 64) tc::tma_load_2d_pair(st + A_TILE + c * A_TILE + A_TILE / 2, mapB, lbar, 0, rowB + 64);
        }
      }
    } else {
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES;
        tc::mbar_wait(&empty[s], ((kb / STAGES) & 1) ^ 1);
        tc::mbar_arrive_expect_tx(&full[s], (uint32_t)(A_TILE + Cp * 128));
        tc::bulk_load(smem + s * STAGE, imgA + (size_t)kb * ((size_t)Rp * 128) + (size_t)rb * A_TILE, A_TILE, &full[s]);
        tc::bulk_load(smem + s * STAGE + A_TILE, imgB + (size_t)kb * ((size_t)Cp * 128), (uint32_t)(Cp * 128), &full[s]);
      }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    for (int kb = 0; kb < KB; ++kb) {
      const int s = kb % STAGES;
      tc::mbar_wait(&full[s], (kb / STAGES) & 1);
      tc::fence_after_sync();
      const uint64_t ad = tc::make_smem_desc_sw128(tc::smem_u32(smem + s * STAGE));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        int c = 0;
        for (int n0 = 0; n0 < Cp; n0 += 256, ++c) {
          const int n = Cp - n0 < 256 ? Cp - n0 : 256;
          if (PAIR) {
            const uint64_t bd = tc::make_smem_desc_sw128(tc::smem_u32(smem + s * STAGE + A_TILE + c * A_TILE));
            tc::mma_f16_ss_pair(tmem_base + n0, tc::advance_desc_k(ad, k), tc::advance_desc_k(bd, k),
                                tc::make_idesc_f16(256, n, 1), (kb | k) != 0);
          } else {
            const uint64_t bd = tc::make_smem_desc_sw128(tc::smem_u32(smem + s * STAGE + A_TILE + n0 * 128));
            tc::mma_f16_ss(tmem_base + n0, tc::advance_desc_k(ad, k), tc::advance_desc_k(bd, k),
                           tc::make_idesc_f16(128, n, 1), (kb | k) != 0);
          }
        }
      }
      if (PAIR) tc::mma_commit_pair(&empty[s]); else tc::mma_commit(&empty[s]);
    }
    if (PAIR) tc::mma_commit_pair(&accfull); else tc::mma_commit(&accfull);
  } else if (warp >= 2 && active) {
    // ---------------------------------------------------------------- epilogue: four threads per row of S (a quarter of the columns each)
    const int quad = warp & 3;
    constexpr int PARTS = FWD_EPI_WARPS / 4;
    const int part = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;                       // 0..FWD_EPI_THREADS-1
    const int row = quad * 32 + lane;
    const int i = rb * 128 + row;
    const int* vrow = hdr_i + (side ? TC_HDR_VALID_G : TC_HDR_VALID_Q);      // validity bits of this side's rows / columns
    const int* vcol = hdr_i + (side ? TC_HDR_VALID_Q : TC_HDR_VALID_G);
    const bool live = i < R && ((__ldcg(vrow + (i >> 5)) >> (i & 31)) & 1);
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int cstep = round_up((C + PARTS - 1) / PARTS, 32);
    const int c_begin = part * cstep < C ? part * cstep : C, c_end = (part + 1) * cstep < C ? (part + 1) * cstep : C;
    if (!P.y) {                                            // label form: the column labels of this side, once per CTA
      const int64_t* cl = side ? P.row_label : P.col_label;
      for (int c = et; c < C; c += FWD_EPI_THREADS) s_collab[c] = cl[c];
      named_bar(1, FWD_EPI_THREADS);
    }
    // positive mask of this thread's columns (<= 128 = 4 words), formed from y WHILE the MMAs run and kept in
    // registers; also stored as bit rows (ybits [N][16] / ybitsT [M][16]) for the backward
    uint32_t mb[4] = {0u, 0u, 0u, 0u}, vb[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int w = 0; w < 4; ++w)
      if (c_begin + 32 * w < c_end) vb[w] = (uint32_t)__ldcg(vcol + ((c_begin + 32 * w) >> 5));
    if (live) {
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const int cw = c_begin + 32 * w;
        if (cw < c_end) {
          if (!P.y) {
            // y[i][j] = (label_i == label_j) (models/model.py:605), restricted to valid columns below
            const int64_t rl = (side ? P.col_label : P.row_label)[i];
#pragma unroll 8
            for (int e = 0; e < 32; ++e) mb[w] |= (cw + e < c_end && s_collab[cw + e] == rl ? 1u : 0u) << e;
          } else {
            // dense form: the bit rows were formed by the prep launch (prep_mask_bits)
            mb[w] = __ldcg(reinterpret_cast<const uint32_t*>(bytes + (side ? L.ybitsT : L.ybits)) + (size_t)i * 16 + (cw >> 5));
          }
          mb[w] &= vb[w];
          reinterpret_cast<uint32_t*>(const_cast<uint8_t*>(bytes) + (side ? L.ybitsT : L.ybits))[(size_t)i * 16 + (cw >> 5)] = mb[w];
        }
      }
    } else if (i < R) {                                    // a row that takes no part: no positives anywhere
#pragma unroll
      for (int w = 0; w < 4; ++w)
        if (c_begin + 32 * w < c_end)
          reinterpret_cast<uint32_t*>(const_cast<uint8_t*>(bytes) + (side ? L.ybitsT : L.ybits))[(size_t)i * 16 + ((c_begin + 32 * w) >> 5)] = 0u;
    }
    const float inv_tau = 1.f / tau_eff;
    float se = 0.f, ps = 0.f, pc = 0.f;
    bool bad = false;
    SDM_STAMP(stamp && et == 0, 1);                          // masks formed
    tc::mbar_wait(&accfull, 0);
    tc::fence_after_sync();
    SDM_STAMP(stamp && et == 0, 2);                          // accumulator complete
    // S (side 0) / St (side 1) leave through a per-warp staging tile in the (now idle) operand ring: a thread owns a
    // row, so direct stores would touch 32 rows with 16 bytes each per instruction; staged, every store instruction
    // writes 256 contiguous bytes of two rows (XOR-swizzled 16-byte chunks keep both directions conflict-free).
    uint8_t* tile = smem + (warp - 2) * 8192;                  // [32 rows][64 columns] fp32
    float* Sbase = base + (side ? L.St : L.S);
#pragma unroll 1
    for (int cg = c_begin; cg < c_end; cg += 64) {
#pragma unroll 1
      for (int it = 0; it < 4; ++it) {
        const int c0 = cg + 16 * it;
        if (c0 >= c_end) break;                                // (warp-uniform)
        uint32_t r[16];
        __syncwarp();                                          // tcgen05.ld is .sync.aligned: reconverge first
        const int wsel = (c0 - c_begin) >> 5;
        const uint32_t bw = (wsel == 0 ? mb[0] : wsel == 1 ? mb[1] : wsel == 2 ? mb[2] : mb[3]) >> ((c0 - c_begin) & 31);
        const uint32_t vw = (wsel == 0 ? vb[0] : wsel == 1 ? vb[1] : wsel == 2 ? vb[2] : vb[3]) >> ((c0 - c_begin) & 31);
        tc::tmem_ld_x16(taddr + c0, r);
        tc::tmem_wait_ld();
        const int nv = c_end - c0 < 16 ? c_end - c0 : 16;       // 16, or 8 in the ragged tail (C is a multiple of 8)
        float sv[16];
        float se2 = 0.f;                                        // two accumulation chains
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          float s = __uint_as_float(r[e]) * inv_tau;            // :86 (x 1/tau: <= 1 ulp from the reference's division)
          const bool in = live && e < nv && ((vw >> e) & 1u);    // (columns that take no part are not in the softmax)
          bad |= in && !(fabsf(s) <= 3.0e38f);                  // :89-91 (NaN or Inf)
          s = fminf(fmaxf(s, -20.f), 20.f);                     // :94 (and :46)
          sv[e] = s;
          if (in) {
            if (e & 1) se2 += fast_exp(s); else se += fast_exp(s);   // |s| <= 20: no overflow without max subtraction
            if ((bw >> e) & 1u) { ps += s; pc += 1.f; }
          }
        }
        se += se2;
#pragma unroll
        for (int e4 = 0; e4 < 4; ++e4)
          *reinterpret_cast<float4*>(tile + lane * 256 + (((it * 4 + e4) ^ (lane & 15)) << 4)) =
              make_float4(sv[4 * e4], sv[4 * e4 + 1], sv[4 * e4 + 2], sv[4 * e4 + 3]);
      }
      __syncwarp();
#pragma unroll 4
      for (int rr = 0; rr < 32; rr += 2) {
        const int rw = rr + (lane >> 4), lc = lane & 15;
        const int col = cg + lc * 4;
        const int gi = rb * 128 + quad * 32 + rw;
        if (col < c_end && gi < R)
          *reinterpret_cast<float4*>(Sbase + (size_t)gi * C + col) =
              *reinterpret_cast<const float4*>(tile + rw * 256 + ((lc ^ (rw & 15)) << 4));
      }
      __syncwarp();
    }
    SDM_STAMP(stamp && et == 0, 3);                          // column loop done
    if (part > 0) { s_part[part - 1][0][row] = se; s_part[part - 1][1][row] = ps; s_part[part - 1][2][row] = pc; }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(&s_bad, 1);
    named_bar(1, FWD_EPI_THREADS);
    if (part == 0 && i < R) {
#pragma unroll
      for (int q = 0; q < PARTS - 1; ++q) { se += s_part[q][0][row]; ps += s_part[q][1][row]; pc += s_part[q][2][row]; }
      const float lse = se > 0.f ? logf(se) : 0.f;            // (se == 0: a row that takes no part, or no valid column)
      float* lse_a = base + (side ? L.lse_c : L.lse_r);
      float* cnt_a = base + (side ? L.cnt_c : L.cnt_r);
      float* ce_a = base + (side ? L.ce_c : L.ce_r);
      lse_a[i] = lse; cnt_a[i] = pc;
      ce_a[i] = pc > 0.f ? (lse - ps / pc) : 0.f;               // -(q * log_p).sum, q uniform over positives (:49-57)
    }
    named_bar(1, FWD_EPI_THREADS);
    SDM_STAMP(stamp && et == 0, 4);                          // row statistics stored
    if (et == 0) {
      hdr_i[72 + side * 4 + rb] = s_bad;
      __threadfence();
      const int total = (N + 127) / 128 + (M + 127) / 128;
      s_last = (atomicAdd(&hdr_i[4], 1) == total - 1);
    }
    named_bar(1, FWD_EPI_THREADS);
    if (s_last) {
      // ---- the last CTA of the pair: means over valid rows / columns and the guards
      __threadfence();
      if (et < 64) {                                              // non-finite feature flags of the 16-row slabs (:79-81)
        const int lim = ((et >> 5) ? L.Mp : L.Np) / PREP_ROWS;
        if ((et & 31) < lim && __ldcg(&hdr_i[8 + et])) atomicOr(&s_st, 2);
      } else if (et < 72) {                                       // non-finite S flags of the forward CTAs (:89-91)
        const int lim = ((((et - 64) >> 2) ? M : N) + 127) / 128;
        if (((et - 64) & 3) < lim && __ldcg(&hdr_i[72 + et - 64])) atomicOr(&s_st, 4);
      }
      double v[5] = {0, 0, 0, 0, 0};                              // sum ce_r, #valid rows, sum ce_c, #valid cols, #rows with a positive
      for (int a = et; a < N; a += FWD_EPI_THREADS) {
        const float c = __ldcg(base + L.cnt_r + a), ce = __ldcg(base + L.ce_r + a);
        if (c > 0.f) { v[4] += 1.0; if (isfinite(ce)) { v[0] += ce; v[1] += 1.0; } }
      }
      for (int a = et; a < M; a += FWD_EPI_THREADS) {
        const float c = __ldcg(base + L.cnt_c + a), ce = __ldcg(base + L.ce_c + a);
        if (c > 0.f && isfinite(ce)) { v[2] += ce; v[3] += 1.0; }
      }
#pragma unroll
      for (int k = 0; k < 5; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if (lane == 0) s_red[warp - 2][k] = v[k];
      }
      named_bar(1, FWD_EPI_THREADS);
      if (et == 0) {
        double t[5] = {0, 0, 0, 0, 0};
        for (int w = 0; w < FWD_EPI_WARPS; ++w)
          for (int k = 0; k < 5; ++k) t[k] += s_red[w][k];
        int st = s_st;
        if (t[4] == 0.0) st |= 8;                                                  // :105-106
        const float lr = t[1] > 0 ? (float)(t[0] / t[1]) : 0.f;
        const float lc = t[3] > 0 ? (float)(t[2] / t[3]) : 0.f;
        float loss = 0.5f * (lr + lc);                                             // :123
        if (!(st & (2 | 4 | 8)) && (isnan(loss) || isinf(loss) || loss < 0.f)) st |= 16;   // :145-147
        if (st & (2 | 4 | 8 | 16)) { st |= 1; loss = 0.f; }
        float* hdr = base + L.hdr;
        hdr[0] = (float)t[1]; hdr[1] = (float)t[3]; hdr[3] = loss;
        hdr_i[2] = st;
        *P.loss = loss;
        *P.status = st;
      }
    }
  }
  tc::fence_before_sync();
  if (PAIR) tc::cluster_sync_all(); else __syncthreads();    // (pair: the peer may still be signalled / read until here)
  if (warp == 1) { if (PAIR) tc::tmem_dealloc_pair(tmem_base, ncols); else tc::tmem_dealloc(tmem_base, ncols); }
}

// =============================================================================================== backward
constexpr int BWD_PROD_WARPS = 16;      // (8 left the dS loop latency-bound: two warps per scheduler)
constexpr int BWD_PROD_THREADS = BWD_PROD_WARPS * 32;
constexpr int BWD_THREADS = 64 + BWD_PROD_THREADS;   // warp 0: bulk-copy producer, warp 1: MMA + TMEM, warps 2-17: dS producers, then epilogue
constexpr int BWD_ROWS_PASS = BWD_PROD_THREADS / 8;  // tile rows covered by one pass of the producers (8 threads per row)
constexpr int BWD_U = 128 / BWD_ROWS_PASS;           // passes per 128-row tile
constexpr int BWD_PARTS = BWD_PROD_WARPS / 4;        // epilogue: threads per output row (a 1/BWD_PARTS slice of the columns each)
constexpr int BWD_STAGE = 2 * A_TILE + B_TILE_MAX;   // dS hi, dS lo, transposed operand chunk (d rows x 64)
constexpr int BWD_STAGES = 2;
constexpr int BWDP_STAGE = 2 * A_TILE + 2 * A_TILE;  // pair mode: dS hi, dS lo, this CTA's half of the two 256-row chunks
constexpr int BWDP_STAGES = 3;
constexpr int BWD_KMAX = 512;
constexpr size_t BWD_SMEM = (size_t)BWD_STAGES * BWD_STAGE + (4 * BWD_KMAX + 4 * 128) * sizeof(float) + 1024;
constexpr size_t BWDP_SMEM = (size_t)BWDP_STAGES * BWDP_STAGE + (4 * BWD_KMAX + 4 * 128) * sizeof(float) + 1024;

// PAIR = true: row blocks 2p / 2p+1 of one (pair, side) as a cta_group::2 pair (see tc_fwd_kernel): every CTA forms
// its own dS tiles, loads half of the transposed operand chunk, the leader issues M = 256 MMAs; ring depth 3.
template <bool PAIR>
__global__ void __launch_bounds__(BWD_THREADS, 1)
tc_bwd_kernel(const __grid_constant__ Batch batch, const __grid_constant__ PairMaps maps, int d, float tau_eff, float eps) {
  constexpr int STAGES = PAIR ? BWDP_STAGES : BWD_STAGES;
  constexpr int STAGE = PAIR ? BWDP_STAGE : BWD_STAGE;
  const reid_sdm_pair& P = batch.p[blockIdx.z];
  const int side = blockIdx.y;                               // 0: dqry rows, K = gal rows; 1: dgal rows, K = qry rows
  const int N = P.N, M = P.M;
  const int R = side ? M : N, K = side ? N : M;
  const int rb = blockIdx.x;
  const uint32_t crank = PAIR ? (uint32_t)(rb & 1) : 0u;
  const bool leader = crank == 0;
  if ((PAIR ? (rb & ~1) : rb) * 128 >= R) return;            // (uniform over the pair)
  const bool active = rb * 128 < R;                          // pair mode: the peer of the last odd block only feeds operands
  const int row0 = rb * 128;
  const TcLayout L = tc_layout(N, M, d);
  float* base = tc_base(P.saved);
  const uint8_t* bytes = reinterpret_cast<const uint8_t*>(base);
  const uint8_t* imgB = bytes + (side ? L.qnt : L.gnt);      // [d rows][K] transposed operand image
  const int* hdr_i = reinterpret_cast<const int*>(base + L.hdr);
  const float* hdr = base + L.hdr;
  bf16* out = reinterpret_cast<bf16*>(side ? P.dgal : P.dqry);
  const int rows_here = !active ? 0 : (R - row0 < 128 ? R - row0 : 128);
  extern __shared__ uint8_t bwd_smem_raw[];
  uint8_t* smem = bwd_smem_raw + ((1024u - (tc::smem_u32(bwd_smem_raw) & 1023u)) & 1023u);
  float* KL = reinterpret_cast<float*>(smem + STAGES * STAGE);           // per K index: lse, weight, 1/count, takes part (1 / 0)
  float* KW = KL + BWD_KMAX;
  float* KIC = KW + BWD_KMAX;
  float* KV = KIC + BWD_KMAX;
  float* RL = KV + BWD_KMAX;                                             // per tile row
  float* RW = RL + 128;
  float* RIC = RW + 128;
  float* RV = RIC + 128;
  __shared__ __align__(8) uint64_t bfull[STAGES], afull[STAGES], empty[STAGES], accfull, xfull;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = (K + 63) >> 6;
  if (threadIdx.x == 0) {
    // pair mode: the leader's afull collects the producer warps of BOTH CTAs
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&bfull[s], 1); tc::mbar_init(&afull[s], PAIR ? 2 * BWD_PROD_WARPS : BWD_PROD_WARPS); tc::mbar_init(&empty[s], 1); }
    tc::mbar_init(&accfull, 1); tc::mbar_init(&xfull, 1);
    tc::fence_barrier_init();
    if (PAIR) tc::prefetch_tensormap(side ? &maps.qt[blockIdx.z] : &maps.gt[blockIdx.z]);
  }
  const uint32_t ncols = tmem_cols_for(d);
  if (warp == 1) { if (PAIR) tc::tmem_alloc_pair(&tmem_base_s, ncols); else tc::tmem_alloc(&tmem_base_s, ncols); }
  tc::fence_before_sync();
  if (PAIR) tc::cluster_sync_all(); else __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
#ifdef REID_SDM_TIMING
  const bool stamp = rb == 0 && side == 0;
#define BSTAMP(cond, slot) do { if (cond) reinterpret_cast<unsigned long long*>(const_cast<int*>(hdr_i) + 80)[8 + (slot)] = gtime(); } while (0)
#else
#define BSTAMP(cond, slot) do { } while (0)
#endif
  // programmatic dependent launch: the prologue above overlapped the forward's tail; nothing the forward writes (the
  // status word included) is read before this point
  tc::grid_dependency_wait();
  const int st = __ldcg(&hdr_i[2]);
  BSTAMP(stamp && threadIdx.x == 64, 0);                     // prologue done

  if (st & 1) {      // the reference returned its non-differentiable zero: gradients are exact zeros (uniform over the pair)
    uint4* o = reinterpret_cast<uint4*>(out + (size_t)row0 * d);
    for (int a = threadIdx.x; a < rows_here * d / 8; a += BWD_THREADS) o[a] = make_uint4(0, 0, 0, 0);
  } else if (warp == 0 && lane == 0) {
    if (PAIR) {
      const CUtensorMap* mapB = side ? &maps.qt[blockIdx.z] : &maps.gt[blockIdx.z];
      uint32_t tx = 0;                                        // bytes of BOTH CTAs per stage
      for (int n0 = 0; n0 < d; n0 += 256) tx += 2u * (uint32_t)(((d - n0 < 256 ? d - n0 : 256) / 2) * 128);
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES;
        tc::mbar_wait(&empty[s], ((kb / STAGES) & 1) ^ 1);
        if (leader) tc::mbar_arrive_expect_tx(&bfull[s], tx);
        const uint32_t lbar = tc::leader_addr(&bfull[s]);
        uint8_t* sb = smem + s * STAGE + 2 * A_TILE;
        int c = 0;
        for (int n0 = 0; n0 < d; n0 += 256, ++c) {
          const int hr = (d - n0 < 256 ? d - n0 : 256) / 2;       // this CTA's rows of the chunk: 32 .. 128 (d % 64 == 0)
          const int rowB = kb * d + n0 + (int)crank * hr;
          for (int r = 0; r < hr; r += 32) tc::tma_load_2d_pair(sb + c * A_TILE + r * 128, mapB, lbar, 0, rowB + r);
        }
      }
    } else {
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES;
        tc::mbar_wait(&empty[s], ((kb / STAGES) & 1) ^ 1);
        tc::mbar_arrive_expect_tx(&bfull[s], (uint32_t)(d * 128));
        tc::bulk_load(smem + s * STAGE + 2 * A_TILE, imgB + (size_t)kb * ((size_t)d * 128), (uint32_t)(d * 128), &bfull[s]);
      }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    for (int kb = 0; kb < KB; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      tc::mbar_wait(&bfull[s], ph);
      tc::mbar_wait(&afull[s], ph);
      tc::fence_after_sync();
      const uint64_t ah = tc::make_smem_desc_sw128(tc::smem_u32(smem + s * STAGE));
      const uint64_t al = tc::make_smem_desc_sw128(tc::smem_u32(smem + s * STAGE + A_TILE));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        int c = 0;
        for (int n0 = 0; n0 < d; n0 += 256, ++c) {
          const int n = d - n0 < 256 ? d - n0 : 256;
          if (PAIR) {
            const uint64_t bd = tc::make_smem_desc_sw128(tc::smem_u32(smem + s * STAGE + 2 * A_TILE + c * A_TILE));
            const uint32_t idesc = tc::make_idesc_f16(256, n, REID_SDM_DS_F16 ? 0 : 1);
            tc::mma_f16_ss_pair(tmem_base + n0, tc::advance_desc_k(ah, k), tc::advance_desc_k(bd, k), idesc, (kb | k) != 0);
            if (!REID_SDM_DS_F16) tc::mma_f16_ss_pair(tmem_base + n0, tc::advance_desc_k(al, k), tc::advance_desc_k(bd, k), idesc, 1);
          } else {
            const uint64_t bd = tc::make_smem_desc_sw128(tc::smem_u32(smem + s * STAGE + 2 * A_TILE + n0 * 128));
            const uint32_t idesc = tc::make_idesc_f16(128, n, REID_SDM_DS_F16 ? 0 : 1);
            tc::mma_f16_ss(tmem_base + n0, tc::advance_desc_k(ah, k), tc::advance_desc_k(bd, k), idesc, (kb | k) != 0);
            if (!REID_SDM_DS_F16) tc::mma_f16_ss(tmem_base + n0, tc::advance_desc_k(al, k), tc::advance_desc_k(bd, k), idesc, 1);
          }
        }
      }
      if (PAIR) tc::mma_commit_pair(&empty[s]); else tc::mma_commit(&empty[s]);
    }
    if (PAIR) tc::mma_commit_pair(&accfull); else tc::mma_commit(&accfull);
  } else if (warp >= 2) {
    // ---------------------------------------------------------------- dS producers
    const int t = threadIdx.x - 64;                          // 0..255
    const float nR = hdr[0], nC = hdr[1];
    const float gscale = (*P.grad_out) * 0.5f / tau_eff;
    // fp16 plane: dL/dS is formed for grad_out = 1 and scaled by 2^12 (the operand then never leaves the fp16 range whatever
    // the upstream gradient is); the epilogue multiplies the rows by gscale / 2^12
    const float wscale = REID_SDM_DS_F16 ? 4096.f : gscale;
    const float out_scale = REID_SDM_DS_F16 ? gscale * (1.f / 4096.f) : 1.f;
    const float wr = nR > 0.f ? wscale / nR : 0.f, wc = nC > 0.f ? wscale / nC : 0.f;
    {
      const float* lse_k = base + (side ? L.lse_r : L.lse_c);
      const float* cnt_k = base + (side ? L.cnt_r : L.cnt_c);
      const float* ce_k = base + (side ? L.ce_r : L.ce_c);
      const float wk = side ? wr : wc;
      const int* vk = hdr_i + (side ? TC_HDR_VALID_Q : TC_HDR_VALID_G);       // rows of the OTHER operand index K
      const int* vr = hdr_i + (side ? TC_HDR_VALID_G : TC_HDR_VALID_Q);
      for (int k = t; k < KB * 64; k += BWD_PROD_THREADS) {
        const bool in = k < K;
        const float c = in ? cnt_k[k] : 0.f;
        const bool valid = in && c > 0.f && isfinite(ce_k[k]);
        KL[k] = in ? lse_k[k] : 0.f;
        KW[k] = valid ? wk : 0.f;
        KIC[k] = c > 0.f ? 1.f / c : 0.f;
        KV[k] = (in && ((vk[k >> 5] >> (k & 31)) & 1)) ? 1.f : 0.f;
      }
      const float* lse_r = base + (side ? L.lse_c : L.lse_r);
      const float* cnt_r = base + (side ? L.cnt_c : L.cnt_r);
      const float* ce_r = base + (side ? L.ce_c : L.ce_r);
      const float wrow = side ? wc : wr;
      if (t < 128) {
        const int gi = row0 + t;
        const bool in = gi < R;
        const float c = in ? cnt_r[gi] : 0.f;
        const bool valid = in && c > 0.f && isfinite(ce_r[gi]);
        RL[t] = in ? lse_r[gi] : 0.f;
        RW[t] = valid ? wrow : 0.f;
        RIC[t] = c > 0.f ? 1.f / c : 0.f;
        RV[t] = (in && ((vr[gi >> 5] >> (gi & 31)) & 1)) ? 1.f : 0.f;
      }
    }
    named_bar(1, BWD_PROD_THREADS);
    BSTAMP(stamp && t == 0, 1);                           // statistics staged
    // this side's view of S (side 0: S [N][M], side 1: St [M][N]) and of the positive mask: rows = tile rows, K contiguous
    const float* Sv = base + (side ? L.St : L.S);
    const uint32_t* bitsv = reinterpret_cast<const uint32_t*>(bytes + (side ? L.ybitsT : L.ybits));
    const int ch = t & 7;                                    // 8 lanes cover 64 consecutive K indices of one row
    // global loads of one K block: 4 rows x 32 bytes of S + 4 mask words per thread
    auto load_block = [&](int kb, float4 (&sa)[BWD_U], float4 (&sb)[BWD_U], uint32_t (&bw)[BWD_U]) {
      const int k0 = kb * 64 + ch * 8;
#pragma unroll
      for (int u = 0; u < BWD_U; ++u) {
        const int gr = row0 + (t >> 3) + BWD_ROWS_PASS * u;
        sa[u] = sb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        bw[u] = 0;
        if (gr < R && k0 < K) {
          sa[u] = *reinterpret_cast<const float4*>(Sv + (size_t)gr * K + k0);
          sb[u] = *reinterpret_cast<const float4*>(Sv + (size_t)gr * K + k0 + 4);
          bw[u] = bitsv[(size_t)gr * 16 + (k0 >> 5)] >> (k0 & 31);
        }
      }
    };
    float4 sa[BWD_U], sb[BWD_U];
    uint32_t bw[BWD_U];
    load_block(0, sa, sb, bw);
#pragma unroll 1
    for (int kb = 0; kb < KB; ++kb) {
      const int s = kb % STAGES;
      const int k0 = kb * 64 + ch * 8;
      float4 na[BWD_U], nb[BWD_U];
      uint32_t nw[BWD_U];
      if (kb + 1 < KB) load_block(kb + 1, na, nb, nw);       // next block's loads fly while this one is formed
      float kl[8], kw[8], kic[8];
      uint32_t kvm = 0;                                        // K indices of this thread that take part
#pragma unroll
      for (int e = 0; e < 8; ++e) { kl[e] = KL[k0 + e]; kw[e] = KW[k0 + e]; kic[e] = KIC[k0 + e]; kvm |= (KV[k0 + e] != 0.f ? 1u : 0u) << e; }
      tc::mbar_wait(&empty[s], ((kb / STAGES) & 1) ^ 1);
      uint8_t* Ahi = smem + s * STAGE;
      uint8_t* Alo = Ahi + A_TILE;
#pragma unroll
      for (int u = 0; u < BWD_U; ++u) {
        const int row = (t >> 3) + BWD_ROWS_PASS * u;
        const bool in = row0 + row < R && k0 < K && RV[row] != 0.f;   // (rows that take no part: dS = 0)
        const float rl = RL[row], rw = RW[row], ric = RIC[row];
        const float sv[8] = {sa[u].x, sa[u].y, sa[u].z, sa[u].w, sb[u].x, sb[u].y, sb[u].z, sb[u].w};
        float hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          // dL/dS (chain through the clamp: zero where saturated); row statistics r*, K-index statistics k*
          const float pos = ((bw[u] >> e) & 1u) ? 1.f : 0.f;
          float g = rw * (fast_exp(sv[e] - rl) - pos * ric) + kw[e] * (fast_exp(sv[e] - kl[e]) - pos * kic[e]);
          if (!in || !((kvm >> e) & 1u) || sv[e] >= 20.f || sv[e] <= -20.f) g = 0.f;
          hi[e] = REID_SDM_DS_F16 ? g : bf16r(g);
          lo[e] = g - hi[e];
        }
        const uint32_t off = (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u + (uint32_t)((ch ^ (row & 7)) << 4);
        if (REID_SDM_DS_F16) {
          uint4 o;
          __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
          for (int e2 = 0; e2 < 4; ++e2) oh[e2] = __floats2half2_rn(hi[2 * e2], hi[2 * e2 + 1]);
          *reinterpret_cast<uint4*>(Ahi + off) = o;
        } else {
          *reinterpret_cast<uint4*>(Ahi + off) = pack8(hi);
          *reinterpret_cast<uint4*>(Alo + off) = pack8(lo);
        }
      }
      tc::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) { if (PAIR && !leader) tc::mbar_arrive_remote_release(&afull[s], 0); else tc::mbar_arrive(&afull[s]); }
#pragma unroll
      for (int u = 0; u < BWD_U; ++u) { sa[u] = na[u]; sb[u] = nb[u]; bw[u] = nw[u]; }
    }
    BSTAMP(stamp && t == 0, 2);                           // all dS tiles formed
    // ---------------------------------------------------------------- epilogue: BWD_PARTS threads per output row (a slice of the columns each)
    if (active) {
      const int quad = warp & 3, half = (warp - 2) >> 2;      // (half: index of this thread's column slice)
      const int row = quad * 32 + lane, gi = row0 + row;
      const bool live = gi < R;
      const int gl = live ? gi : row0;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
      const float den = (base + (side ? L.den_g : L.den_q))[gl];
      const float rden = out_scale / den;
      const bool clamped = !(den > bf16r(eps));          // norm <= eps: the denominator is the constant eps
      const int dh = d / BWD_PARTS;                      // (d % 64 == 0: a multiple of 16)
      const int c_begin = half * dh, c_end = c_begin + dh;
      float* pdot = RL;                                  // the row statistics RL / RW / RIC / RV are dead now: reuse as [BWD_PARTS][128] partial dots
      tc::mbar_wait(&accfull, 0);                        // every MMA has retired: the stage buffers are free
      tc::fence_after_sync();
      BSTAMP(stamp && t == 0, 3);                     // accumulator complete
      // x^ of this row block (the values the forward multiplied): 128 rows x d of the K-major operand image,
      // one contiguous 16 KB bulk copy per 64-feature block, into the (now idle) stage buffers
      if (t == 0) {
        const uint8_t* ximg = bytes + (side ? L.gn : L.qn);
        const size_t kstride = (size_t)(side ? L.Mp : L.Np) * 128;
        tc::mbar_arrive_expect_tx(&xfull, (uint32_t)((d >> 6) * A_TILE));
        for (int kb = 0; kb < (d >> 6); ++kb)
          tc::bulk_load(smem + kb * A_TILE, ximg + (size_t)kb * kstride + (size_t)rb * A_TILE, A_TILE, &xfull);
      }
      named_bar(1, BWD_PROD_THREADS);                    // every producer is past its last read of RL / RW
      tc::mbar_wait(&xfull, 0);
      BSTAMP(stamp && t == 0, 4);                     // x^ tile in shared memory
      const uint8_t* xrow = smem + (row >> 3) * 1024 + (row & 7) * 128;
      auto xchunk = [&](int c) -> uint4 {                // 8 consecutive features starting at c (multiple of 8)
        return *reinterpret_cast<const uint4*>(xrow + (c >> 6) * A_TILE + ((((c & 63) >> 3) ^ (row & 7)) << 4));
      };
      float dot = 0.f, dot2 = 0.f;
#pragma unroll 1
      for (int c0 = c_begin; c0 < c_end; c0 += 16) {
        uint32_t r[16];
        __syncwarp();
        tc::tmem_ld_x16(taddr + c0, r);
        float fa[8], fb[8];
        unpack8(xchunk(c0), fa); unpack8(xchunk(c0 + 8), fb);
        tc::tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          dot = fmaf(__uint_as_float(r[e]), fa[e], dot);
          dot2 = fmaf(__uint_as_float(r[8 + e]), fb[e], dot2);
        }
      }
      BSTAMP(stamp && t == 0, 5);                     // pass 1 (row dots) done
      pdot[half * 128 + row] = dot + dot2;
      named_bar(1, BWD_PROD_THREADS);
      dot = 0.f;
#pragma unroll
      for (int q = 0; q < BWD_PARTS; ++q) dot += pdot[q * 128 + row];
      if (clamped) dot = 0.f;
      // the gradient rows leave through a per-warp staging tile (free ring space above the x^ tile): a thread owns a
      // row, so direct stores would write 32 bytes to each of 32 rows per step; staged, every store instruction writes
      // 128 contiguous bytes of four rows (d/2 is a multiple of 32: groups of 64 columns, the last may hold 32)
      uint8_t* otile = smem + (size_t)(d >> 6) * A_TILE + (warp - 2) * 4096;   // [32 rows][64 columns] bf16
      const int wrow0 = row0 + quad * 32;
#pragma unroll 1
      for (int cg = c_begin; cg < c_end; cg += 64) {
#pragma unroll 1
        for (int it = 0; it < 4; ++it) {
          const int c0 = cg + 16 * it;
          if (c0 >= c_end) break;                            // (warp-uniform)
          uint32_t r[16];
          __syncwarp();
          tc::tmem_ld_x16(taddr + c0, r);
          float fa[8], fb[8], oa[8], ob[8];
          unpack8(xchunk(c0), fa); unpack8(xchunk(c0 + 8), fb);
          tc::tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            oa[e] = (__uint_as_float(r[e]) - fa[e] * dot) * rden;
            ob[e] = (__uint_as_float(r[8 + e]) - fb[e] * dot) * rden;
          }
          *reinterpret_cast<uint4*>(otile + lane * 128 + (((2 * it) ^ (lane & 7)) << 4)) = pack8(oa);
          *reinterpret_cast<uint4*>(otile + lane * 128 + (((2 * it + 1) ^ (lane & 7)) << 4)) = pack8(ob);
        }
        __syncwarp();
#pragma unroll 4
        for (int rr = 0; rr < 32; rr += 4) {
          const int rw = rr + (lane >> 3), lc = lane & 7;
          const int col = cg + lc * 8;
          if (col < c_end && wrow0 + rw < R)
            *reinterpret_cast<uint4*>(out + (size_t)(wrow0 + rw) * d + col) =
                *reinterpret_cast<const uint4*>(otile + rw * 128 + ((lc ^ (rw & 7)) << 4));
        }
        __syncwarp();
      }
      BSTAMP(stamp && t == 0, 6);                     // pass 2 (gradient rows written) done
    }
  }
  tc::fence_before_sync();
  if (PAIR) tc::cluster_sync_all(); else __syncthreads();    // (pair: the peer may still be signalled / read until here)
  if (warp == 1) { if (PAIR) tc::tmem_dealloc_pair(tmem_base, ncols); else tc::tmem_dealloc(tmem_base, ncols); }
}

// per-pair tensor maps over the operand images of the saved buffers (pair kernels)
bool fill_maps(PairMaps& maps, const reid_sdm_pair* pairs, int n_pairs, int d, bool transposed) {
  for (int i = 0; i < n_pairs; ++i) {
    const TcLayout L = tc_layout(pairs[i].N, pairs[i].M, d);
    const uint8_t* bytes = reinterpret_cast<const uint8_t*>(tc_base(pairs[i].saved));
    if (!transposed) {
      if (!tc_host::make_map_image(&maps.q[i], bytes + L.qn, (int64_t)L.Np * (d / 64), 64) ||
          !tc_host::make_map_image(&maps.g[i], bytes + L.gn, (int64_t)L.Mp * (d / 64), 64))
        return false;
    } else {
      if (!tc_host::make_map_image(&maps.qt[i], bytes + L.qnt, (int64_t)d * (L.Np / 64), 32) ||
          !tc_host::make_map_image(&maps.gt[i], bytes + L.gnt, (int64_t)d * (L.Mp / 64), 32))
        return false;
    }
  }
  return true;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int fill_batch(Batch& b, const reid_sdm_pair* pairs, int n_pairs, bool bwd, int* max_blocks) {
  if (!pairs || n_pairs <= 0 || n_pairs > REID_SDM_MAX_PAIRS) return REID_E_INVALID;
  b.n_pairs = n_pairs;
  int mb = 1;
  for (int i = 0; i < n_pairs; ++i) {
    const reid_sdm_pair& p = pairs[i];
    if (!p.qry || !p.gal || !p.loss || !p.status || !p.saved) return REID_E_INVALID;
    if (!p.y && (!p.row_label || !p.col_label)) return REID_E_INVALID;         // dense y, or the label form
    if (bwd && (!p.grad_out || !p.dqry || !p.dgal || !aligned16(p.dqry) || !aligned16(p.dgal))) return REID_E_INVALID;
    b.p[i] = p;
    const int blocks = ((p.N > p.M ? p.N : p.M) + 127) / 128;
    if (blocks > mb) mb = blocks;
  }
  *max_blocks = mb;
  return REID_OK;
}

}  // namespace

// the tcgen05 path serves bf16 batches whose every pair is one TMEM-resident problem
bool tc_eligible(const reid_sdm_pair* pairs, int n_pairs, int dtype, int d) {
  if (dtype != REID_DTYPE_BF16 || !pairs || n_pairs <= 0 || n_pairs > REID_SDM_MAX_PAIRS) return false;
  if (d % 64 != 0 || d < 64 || d > 512) return false;
  for (int i = 0; i < n_pairs; ++i) {
    const reid_sdm_pair& p = pairs[i];
    if (p.N < 64 || p.M < 64 || p.N > 512 || p.M > 512 || (p.N & 7) || (p.M & 7)) return false;
    if (!aligned16(p.qry) || !aligned16(p.gal) || (p.y && !aligned16(p.y))) return false;
    if (!p.y && (!p.row_label || !p.col_label)) return false;
  }
  return true;
}

int tc_forward(const reid_sdm_pair* pairs, int n_pairs, int d, float tau, float eps, cudaStream_t st) {
  Batch b;
  int mb = 1;
  const int rc = fill_batch(b, pairs, n_pairs, false, &mb);
  if (rc != REID_OK) return rc;
  const float tau_eff = fmaxf(0.15f, fminf(0.5f, tau));                    // sdm_loss.py:28
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(tc_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM) != cudaSuccess) return REID_E_CUDA;
    if (cudaFuncSetAttribute(tc_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWDP_SMEM) != cudaSuccess) return REID_E_CUDA;
    attr_done = true;
  }
  const size_t prep_smem = (size_t)PREP_ROWS * (d + 8) * 2;
  tc_prep_kernel<<<dim3(mb * 128 / PREP_ROWS, 3, n_pairs), PREP_THREADS, prep_smem, st>>>(b, d, eps);
  REID_CHECK_LAUNCH();
  // -DREID_SDM_PAIR=1 selects the cta_group::2 pair kernels.  They are parity-tested but NOT the default: measured on
  // C5 (10 pairs) the step takes 90 us against 78 us with the single-CTA kernels -- the six to eight 4-8 KB tensor-map
  // loads per stage and the cluster launch cost more than the halved operand stream and the deeper ring save.
  const bool pair = REID_SDM_PAIR != 0;
  PairMaps maps;                                             // (only the entries of this launch are read by the kernel)
  if (pair) {
    if (!fill_maps(maps, pairs, n_pairs, d, false)) return REID_E_CUDA;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((mb + 1) / 2 * 2, 2, n_pairs); cfg.blockDim = dim3(FWD_THREADS); cfg.dynamicSmemBytes = FWDP_SMEM; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = REID_SDM_PDL ? 2 : 1;
    if (cudaLaunchKernelEx(&cfg, tc_fwd_kernel<true>, b, maps, d, tau_eff) != cudaSuccess) return REID_E_CUDA;
  } else {
    // programmatic dependent launch: the forward's CTAs start (barriers, TMEM) while the prep kernel drains
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(mb, 2, n_pairs); cfg.blockDim = dim3(FWD_THREADS); cfg.dynamicSmemBytes = FWD_SMEM; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = REID_SDM_PDL ? 1 : 0;
    if (cudaLaunchKernelEx(&cfg, tc_fwd_kernel<false>, b, maps, d, tau_eff) != cudaSuccess) return REID_E_CUDA;
  }
  REID_CHECK_LAUNCH();
  return REID_OK;
}

int tc_backward(const reid_sdm_pair* pairs, int n_pairs, int d, float tau, float eps, cudaStream_t st) {
  Batch b;
  int mb = 1;
  const int rc = fill_batch(b, pairs, n_pairs, true, &mb);
  if (rc != REID_OK) return rc;
  const float tau_eff = fmaxf(0.15f, fminf(0.5f, tau));
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(tc_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM) != cudaSuccess) return REID_E_CUDA;
    if (cudaFuncSetAttribute(tc_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWDP_SMEM) != cudaSuccess) return REID_E_CUDA;
    attr_done = true;
  }
  const bool pair = REID_SDM_PAIR != 0;                     // (see tc_forward)
  PairMaps maps;
  if (pair) {
    if (!fill_maps(maps, pairs, n_pairs, d, true)) return REID_E_CUDA;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((mb + 1) / 2 * 2, 2, n_pairs); cfg.blockDim = dim3(BWD_THREADS); cfg.dynamicSmemBytes = BWDP_SMEM; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = REID_SDM_PDL ? 2 : 1;
    if (cudaLaunchKernelEx(&cfg, tc_bwd_kernel<true>, b, maps, d, tau_eff, eps) != cudaSuccess) return REID_E_CUDA;
  } else {
    // programmatic dependent launch: whatever kernel precedes this one on the stream (the forward, in a fused step),
    // the prologue overlaps its tail; the kernel waits (griddepcontrol.wait) before it reads any input
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(mb, 2, n_pairs); cfg.blockDim = dim3(BWD_THREADS); cfg.dynamicSmemBytes = BWD_SMEM; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = REID_SDM_PDL ? 1 : 0;
    if (cudaLaunchKernelEx(&cfg, tc_bwd_kernel<false>, b, maps, d, tau_eff, eps) != cudaSuccess) return REID_E_CUDA;
  }
  REID_CHECK_LAUNCH();
  return REID_OK;
}

}  // namespace sdm
