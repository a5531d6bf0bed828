/* reid_b200.h -- C ABI of libreid_b200.so: B200 (sm_100a) kernels for the PRCV2025REID hot path.
 *
 * The reference (LingmaFuture/PRCV2025REID) has no FFI layer; its boundary for this path is a
 * set of Python functions.  Each entry point below names the reference
 * code it replaces (file:line under the reference tree).  Conventions:
 *   - C linkage, POD arguments only, caller-owned DEVICE memory on the current CUDA device;
 *   - every launch is ordered on the passed stream (a cudaStream_t as void*), no host sync,
 *     no hidden allocation -- scratch comes from the caller (`workspace`), sized by
 *     reid_workspace_bytes();
 *   - return 0 = OK, negative = REID_E_* (see reid_strerror); never throws.
 *   - gallery / query feature rows are dense row-major, row stride = d elements, d % 8 == 0.
 */
#ifndef REID_B200_H
#define REID_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define REID_OK 0
#define REID_E_INVALID -1    /* bad argument (null pointer, unsupported d / k / topk ...) */
#define REID_E_CUDA -2       /* a CUDA runtime / driver call failed (cudaGetLastError kept) */
#define REID_E_WORKSPACE -3  /* workspace too small */
#define REID_E_UNSUPPORTED -4

#define REID_DTYPE_F32 0
#define REID_DTYPE_BF16 1
#define REID_DTYPE_F16 2

/* fixed sizes of the fused retrieval path */
#define REID_KLIST 32        /* per-(query, gallery-chunk) running top list; candidates are complete down to it */
#define REID_RTOP 32         /* candidates re-scored in fp32 per query and shard */

const char* reid_strerror(int code);
int reid_abi_version(void);

/* ---- K1: eval_mm_protocol.py:46-48 `l2n` (F.normalize(x, dim=-1)): out = x / max(||x||2, eps).
 * out_f32 and/or out_f16 may be NULL.  out_f16 is the tensor-core operand copy. */
int reid_l2norm_rows(const float* x, float* out_f32, void* out_f16, int64_t rows, int d, float eps,
                     void* stream);

/* ---- K2: eval_mm_protocol.py:328-365 `extract_query_feat` on pre-extracted features:
 * out[q] = l2n( sum_j w[mod_id[q,j]] * l2n(feats[q,j,:]) ), slots with mod_id < 0 are skipped;
 * k == 1 (single modality) reproduces l2n(l2n(f)) without the weight (:209-210, :359). */
int reid_mm_fuse_normalize(const float* feats, const int32_t* mod_id, const float* w, int n_mod,
                           float* out_f32, void* out_f16, int64_t Q, int k, int d, void* stream);

/* ---- K3: eval_mm_protocol.py:50-53 `cosine_sim` (a @ b.T) as a TMA-fed tcgen05 GEMM on fp16
 * operands with fp32 accumulation in TMEM; S is [Q, ldS] fp32 row-major, ldS >= G. */
int reid_sim_gemm(const void* q_f16, const void* g_f16, float* S, int64_t Q, int64_t G, int d,
                  int64_t ldS, void* stream);

/* ---- gallery identity index (replaces the per-query `(g_pids == pid)` scans of
 * eval_mm_protocol.py:390,427).  sorted_pid/order: gallery rows sorted by pid (stable).
 * max_run (device int32[1]) = largest number of gallery rows sharing one pid (= Pmax bound). */
int reid_pid_index_build(const int64_t* g_pid, int64_t G, int64_t* sorted_pid, int32_t* order,
                         int32_t* max_run, void* workspace, size_t workspace_bytes, void* stream);
/* code[i] = first position of pids[i] in sorted_pid, or -1 when absent; count[i] = run length. */
int reid_pid_lookup(const int64_t* sorted_pid, int64_t G, const int64_t* pids, int64_t n,
                    int32_t* code, int32_t* count, void* stream);

/* ---- exact fp32 scores of every query's positives (eval_mm_protocol.py:427 `is_pos`),
 * slot j of query q is gallery row order[q_code[q] + j].  Rows outside the local shard
 * [g_offset, g_offset + G_local), masked rows (excl) and slots >= q_count[q] get -inf.
 * excl: [Q, E] global gallery indices masked for the query (same-image rule :408-418), -1 pad. */
int reid_pos_scores(const float* q_f32, const float* g_f32, const int32_t* order,
                    const int32_t* q_code, const int32_t* q_count, const int32_t* excl, int E,
                    int64_t Q, int64_t G_local, int64_t g_offset, int d, int Pmax,
                    float* pos_score, void* stream);
/* sort each row of pos_score descending in place; n_pos[q] = number of finite entries. */
int reid_pos_sort(float* pos_score, int32_t* n_pos, int64_t Q, int Pmax, void* stream);

/* ---- K3+K4 fused: eval_mm_protocol.py:401-455 for one gallery shard without materialising
 * S or an argsort.  tcgen05 GEMM (fp16 in, fp32 TMEM accumulators) whose epilogue
 *   (a) counts, for every positive threshold pos_thr[q,j] (sorted desc), the local gallery rows
 *       that are neither positives of q nor masked and score above it  -> pos_above[q,j] (+=)
 *   (b) appends every row that beats the running REID_KLIST-th best of its (query, chunk) to
 *       cand_score/cand_idx[q, chunk, :cand_cap] (local row index), count in cand_count[q,chunk].
 * g_code / q_code: pid codes from reid_pid_lookup.  n_chunks = candidate slots per query (buffers are sized with it):
 * work items are (block of 256 queries, gallery rows); the query blocks that fill whole waves of the persistent CTA
 * pairs run against the whole shard (slot 0 only), the blocks of the last, partial wave are cut into n_chunks gallery
 * chunks so that this wave is full too.  pos_above and cand_count must be zeroed.
 * Pmax <= 64 thresholds per query and call; pos_stride = row stride (elements) of pos_thr and pos_above (0 = Pmax),
 * so that a query with more positives is processed in windows of 64 thresholds (pointers advanced by the caller).
 * cand_thr [Q] (optional): per-query score with >= REID_KLIST candidates at or above it (-inf if fewer);
 * every candidate the re-scorer can need lies at or above it.
 * n_shards = ranks the WHOLE gallery of a query is partitioned over (0 / 1 = this shard is the whole gallery).
 * Counting classes (shards of >= 32768 rows; a calibration pre-pass over a strided 1/32 of the shard, 2048 .. 8192
 * rows, estimates every threshold's rank inside the shard): thresholds ranked above max(1024 / n_shards, 8 calibration hits) rows are
 * counted on a 1/32 row sample, those above 32768 / n_shards rows on a 1/1024 row sample, so that a sampled count
 * rests on >= 32 sampled rows gallery-wide; all shallower thresholds on every row.  The pre-pass also seeds the running
 * candidate threshold of (b): a positive threshold with >= REID_KLIST (16) sample rows above it has that many shard rows above it.
 * Results are bit-identical from run to run (the classes do not depend on the order in which warps finish).
 * flags: REID_FUSED_EXACT_COUNTS = count every threshold on every row (no sampling; slow for deep positives);
 *        REID_FUSED_NO_CANDIDATES = counting only (cand_* may be NULL);
 *        REID_FUSED_KLIST16 = candidates complete down to the 16th (not the REID_KLIST-th) best score of a slot: what the
 *        re-scorer needs when the completeness bound is exchanged over several shards (reid_cand_select kx = 16). */
#define REID_FUSED_EXACT_COUNTS 1
#define REID_FUSED_NO_CANDIDATES 2
#define REID_FUSED_KLIST16 4
int reid_retrieve_fused(const void* q_f16, const void* g_f16, const int32_t* q_code,
                        const int32_t* g_code, const int32_t* excl, int E, const float* pos_thr,
                        const int32_t* n_pos, int64_t Q, int64_t G_local, int64_t g_offset, int d,
                        int Pmax, int pos_stride, int n_chunks, int n_shards, int cand_cap, int flags,
                        int32_t* pos_above, float* cand_score, int32_t* cand_idx, int32_t* cand_count,
                        float* cand_thr, void* workspace, size_t workspace_bytes, void* stream);

/* ---- exact fp32 SIMT form of the same step (all scores in fp32, CUDA cores).  Used for the
 * queries the fused path flags, for tiny problems, and as the in-library cross-check.
 * q_sel: optional list of n_sel query indices to process (NULL = all Q). */
int reid_retrieve_exact(const float* q_f32, const float* g_f32, const int32_t* q_code,
                        const int32_t* g_code, const int32_t* excl, int E, const float* pos_thr,
                        const int32_t* n_pos, const int32_t* q_sel, int64_t n_sel, int64_t Q,
                        int64_t G_local, int64_t g_offset, int d, int Pmax, int n_chunks,
                        int cand_cap, int32_t* pos_above, float* cand_score, int32_t* cand_idx,
                        int32_t* cand_count, void* stream);

/* ---- candidate selection, re-scoring and the decidability check (replace `argsort` :423 for the head of the ranking).
 * reid_cand_select: gathers a query's candidate slots (those at or above cand_thr when given), sorts them by approximate
 *   score and keeps the REID_RTOP best -> sel_score / sel_idx [Q, REID_RTOP] (descending; local row index, -1 pad),
 *   sel_n [Q]; sel_cut [Q] = the kx-th best approximate score (kx <= REID_KLIST; -inf when the shard offers fewer than
 *   kx candidates); sel_flag [Q] = 1 when a candidate buffer overflowed.
 * With several gallery shards the host all-reduces sel_cut with MAX (`bound`: the best shard's kx-th best score), so
 *   that every shard re-scores only the few rows that can still reach the gallery-wide head; one shard: bound = sel_cut.
 * reid_rescore_topk: re-scores the selected rows at or above `bound` in fp32 (same dot routine as reid_pos_scores),
 *   sorts them (score desc, index asc) -> top_score / top_idx [Q, REID_RTOP] (global gallery index, -1 pad).  A local
 *   row that is not re-scored has an exact score <= bound + eps (eps = bound on |approx - exact|, 0 for the exact
 *   path): positives whose threshold lies above bound + eps get their exact local count written to pos_above, the
 *   others keep max(count so far, re-scored rows above them).  lb0 [Q] = re-scored non-positive rows above the best positive.
 * reid_topk_check (after the exchange; top_score = the merged list [Q, list_len], lb0 summed over the shards): flag |= 2
 *   when the k-th best exact score is not above bound + eps, |= 4 when CMC@10 is undecidable (best positive not above
 *   bound + eps and fewer than 10 re-scored rows above it); flag != 0 -> re-run the query through reid_retrieve_exact. */
int reid_cand_select(const float* cand_score, const int32_t* cand_idx, const int32_t* cand_count,
                     const float* cand_thr /* optional, from reid_retrieve_fused */, int64_t Q, int n_chunks,
                     int cand_cap, int kx, float* sel_score, int32_t* sel_idx, int32_t* sel_n, float* sel_cut,
                     int32_t* sel_flag, void* stream);
int reid_rescore_topk(const float* q_f32, const float* g_f32, const int32_t* q_code, const int32_t* g_code,
                      const float* pos_thr, const int32_t* n_pos, const float* sel_score, const int32_t* sel_idx,
                      const int32_t* sel_n, const float* bound, int64_t Q, int64_t G_local, int64_t g_offset, int d,
                      int Pmax, float eps, int32_t* pos_above, float* top_score, int32_t* top_idx, int32_t* lb0,
                      void* stream);
int reid_topk_check(const float* top_score, int list_len, int topk, const float* bound, float eps,
                    const float* pos_thr, const int32_t* n_pos, int Pmax, const int32_t* lb0, int64_t Q,
                    int32_t* flag, void* stream);

/* merge n_lists per-shard top lists [n_lists, Q, list_len] (each sorted; the global top-k is contained in the union of
 * the shards' top-k, so list_len = topk suffices) into out[Q, topk] (score desc, idx asc) */
int reid_merge_topk(const float* scores, const int32_t* idx, int n_lists, int64_t Q, int list_len, int topk,
                    float* out_score, int32_t* out_idx, void* stream);

/* ---- metrics: eval_mm_protocol.py:435-469.  rank_j = 1 + pos_above[q,j] + j;
 * AP = (1/P) sum_j (j+1)/rank_j in float64; hit@k = rank_0 <= k; queries with n_pos == 0 are
 * skipped (:430-432).  out[5] = {mAP, R@1, R@5, R@10, num_valid}; ap_per_query optional [Q]. */
int reid_metrics_reduce(const int32_t* pos_above, const int32_t* n_pos, int64_t Q, int Pmax,
                        double* out, double* ap_per_query, void* stream);

/* ---- train-time evaluator reductions, train.py:101-138 (`compute_map` AP@k :116-124, `compute_cmc` :133-136):
 * top_idx [Q, list_len] exact ranking (global gallery indices, -1 pad), labels int64.  ap[q] = AP over the
 * matches inside the first k positions (fp32 like the reference), -1 when there is none; hit[q] = any match. */
int reid_topk_label_metrics(const int32_t* top_idx, const int64_t* q_label, const int64_t* g_label,
                            int64_t Q, int list_len, int k, float* ap, int32_t* hit, void* stream);

/* ---- SDM loss: models/sdm_loss.py:13-149 `sdm_loss_stable`, batched over modality pairs.
 * One launch computes every pair; pair p uses qry[p] [N_p, d], gal[p] [M_p, d] (dtype F32, BF16 or
 * F16), y[p] [N_p, M_p] float {0,1}.  loss[p] (fp32), status[p] (bit0: returned the reference's
 * non-differentiable zero; bit1: non-finite feature; bit2: non-finite S; bit3: no positives).
 * saved[p]: scratch of reid_sdm_saved_floats(N,M,d) floats kept for the backward (S, row / column
 * statistics and the normalised operands). */
typedef struct {
  const void* qry; const void* gal; const float* y;
  int32_t N; int32_t M;
  float* loss; int32_t* status; float* saved;
  const float* grad_out;   /* backward only: upstream scalar gradient */
  void* dqry; void* dgal;  /* backward only: gradients in the input dtype */
  /* LABEL FORM (y == NULL), the SDM section of compute_loss, models/model.py:586-622, without materialising y:
   *   y[i][j] = row_valid[i] && col_valid[j] && row_label[i] == col_label[j]   (:605)
   * and rows of qry / gal whose valid byte is 0 (the feature masks, :570-602) are left out of the loss altogether --
   * out of the softmax denominators, the means and the guards -- and receive exact-zero gradients.  row_valid /
   * col_valid may be NULL (all valid).  Served by every code path (tcgen05, small-batch, general CUDA-core kernels);
   * a pair with neither y nor both label arrays is REID_E_INVALID. */
  const int64_t* row_label; const int64_t* col_label;
  const uint8_t* row_valid; const uint8_t* col_valid;
} reid_sdm_pair;
#define REID_SDM_MAX_PAIRS 16
size_t reid_sdm_saved_floats(int N, int M, int d);
/* 1 when this batch runs on the tcgen05 path (csrc/sdm_tc.cu: bf16, 64 <= N,M <= 512 multiples of 8,
 * d % 64 == 0, d <= 512, 16-byte aligned tensors), 0 when it runs on the fp32 CUDA-core path (csrc/sdm.cu). */
int reid_sdm_uses_tensor_cores(const reid_sdm_pair* pairs, int n_pairs, int dtype, int d);
int reid_sdm_fwd(const reid_sdm_pair* pairs, int n_pairs, int dtype, int d, float tau, float eps,
                 void* stream);
int reid_sdm_bwd(const reid_sdm_pair* pairs, int n_pairs, int dtype, int d, float tau, float eps,
                 void* stream);
/* One training step = forward + backward (the autograd pass train.py:969-972 triggers) with the upstream gradient
 * grad_out[p] of every pair known up front (the weight of loss p in the objective): one launch for small pairs
 * (N, M <= 32: rows, S and statistics never leave shared memory), otherwise reid_sdm_fwd followed by reid_sdm_bwd. */
int reid_sdm_step(const reid_sdm_pair* pairs, int n_pairs, int dtype, int d, float tau, float eps,
                  void* stream);
/* kernel launches reid_sdm_step issues for this batch: 1 (small pairs), 3 (tcgen05: pack + forward + backward), 2 (general) */
int reid_sdm_step_launches(const reid_sdm_pair* pairs, int n_pairs, int dtype, int d);

/* scratch sizes.  which: 0 = reid_pid_index_build(G), 1 = reid_retrieve_fused */
size_t reid_workspace_bytes(int which, int64_t Q, int64_t G, int d);

/* device properties the host side sizes grids with */
int reid_device_sm_count(void);

#ifdef __cplusplus
}
#endif
#endif
