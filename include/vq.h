/*
 * vq.h — C ABI of libvq_b200: the B200-native match-scoring path of Video Query.
 *
 * The reference (PARC-projects/video-query-algorithms) is pure Python and has no FFI; its seam
 * for this path is method dispatch on Ticket / TargetClip / Hyperparameter driven by
 * compute_matches (reference src/models/compute_matches.py:8).  Each entry point below replaces
 * the arithmetic of one reference method; the Python host layer (video_query_algorithms_b200/)
 * binds them with ctypes and keeps the reference's method names and semantics.
 *
 * Conventions
 *   - every function returns 0 on success, < 0 on error; vq_last_error() gives the text
 *     (thread-local).  No C++ exception crosses this boundary.
 *   - all buffers are caller-allocated and caller-owned; HOST pointers unless the name says _dev.
 *   - a store is an opaque handle that owns one HBM-resident shard: clip-major rows of
 *     n_streams * stream_len fp32 (stream_len = n_splits * dim), 16-byte aligned.
 *   - rows are identified by GLOBAL row number = first_global_row + local row; lists come back
 *     in ascending row order (= database order, which the reference's seeded sampling needs,
 *     ticket.py:326-333).
 *   - calls on one store must be serialised by the caller (the broker runs one job at a time,
 *     broker.py:90-92) but may come from any thread: the library sets the device per call.
 *   - `stream` arguments are a cudaStream_t passed as void*; NULL = the store's own stream.
 */
#ifndef VQ_B200_H
#define VQ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VQ_MAX_STREAMS 4
#define VQ_MAX_TOPK 1024
#define VQ_ABI_VERSION 2

typedef struct vq_store vq_store;

/* ---------------------------------------------------------------- errors / environment */
const char *vq_last_error(void);
int vq_abi_version(void);
int vq_device_count(int *count_out);

/* ---------------------------------------------------------------- feature store (A2)
 * Replaces Ticket._get_candidate_features (ticket.py:358-382): instead of pulling every feature
 * row over HTTP per job, rows live in HBM for the life of the broker process.               */
int vq_store_create(vq_store **out, int device, int64_t n_rows, int n_streams, int n_splits,
                    int dim, int64_t first_global_row);
int vq_store_destroy(vq_store *s);
/* Incremental growth (load_db.py adds clips to a search set, reference load_db.py:10-28): append n_new rows after the
 * last row of the shard; the buffers grow geometrically (vq_store_reserve pre-sizes them).  Per-row split weights set
 * with vq_store_set_split_weights are dropped by an append and must be set again for the grown shard.              */
int vq_store_reserve(vq_store *s, int64_t capacity);
int vq_store_append(vq_store *s, int64_t n_new, const float *rows /* [n_new][n_streams][n_splits][dim] host */);
int vq_store_describe(const vq_store *s, int64_t *n_rows, int *n_streams, int *n_splits, int *dim,
                      int64_t *first_global_row, int *device);
/* rows: [n_rows][n_streams][n_splits][dim] fp32, local row range.  Pinned memory is copied
 * asynchronously on the store's stream; the call returns after the copy has completed.       */
int vq_store_upload(vq_store *s, int64_t first_row, int64_t n_rows, const float *rows);
int vq_store_download(vq_store *s, int64_t first_row, int64_t n_rows, float *rows_out);
/* Pipelined ingest (north_star item 1: "uploaded from pinned memory"): the host fills one pinned chunk while the
 * previous one is in flight.  vq_store_upload_async only enqueues ONE copy on the store's stream (rows must stay
 * untouched until vq_store_sync returns); vq_pinned_alloc / vq_pinned_free hand out page-locked staging buffers.       */
int vq_store_upload_async(vq_store *s, int64_t first_row, int64_t n_rows, const float *rows_pinned);
int vq_store_sync(vq_store *s);
int vq_pinned_alloc(void **out, int64_t bytes);
int vq_pinned_free(void *p);
/* 1 / (number of splits the clip actually has) per (row, stream); NULL restores "all present".
 * Missing (row, stream, split) slots must be uploaded as zeros (ticket.py:155-157).          */
int vq_store_set_split_weights(vq_store *s, const float *inv_counts /* [n_rows][n_streams] */);
/* VQSYN-1 counter-based generator (oracle/synth.py is the CPU twin), fills the whole shard.  */
int vq_store_fill_synthetic(vq_store *s, uint64_t seed, const float *stream_means);
/* device base pointer of the shard (for zero-copy consumers such as bench.py)               */
int vq_store_device_ptr(vq_store *s, void **rows_dev);

/* ---------------------------------------------------------------- single-query scan (A3-A7)
 * compute_similarities (ticket.py:120-163) + compute_scores (:165-180) + the two candidate
 * scans of select_clips_to_review (:325-327) + ranking (:266), one pass over the shard.     */
typedef struct vq_scan_params {
    double weights[VQ_MAX_STREAMS]; /* w_s in stream order (ticket.py:176)                   */
    double threshold;               /* match set  = { score >= threshold }                   */
    double lower_limit;             /* near set   = { lower_limit <= score < threshold }     */
    double eps;                     /* tie band   = |score - threshold| < eps or |score - lower_limit| < eps */
    int32_t topk;                   /* 0..VQ_MAX_TOPK best rows (score desc, row asc)         */
    int32_t want_sims;              /* also keep per-stream similarities [n_rows][n_streams] */
} vq_scan_params;

typedef struct vq_scan_counts {
    int64_t n_match, n_near, n_tie;
    int32_t n_topk;
    float scan_ms;                  /* device time of the fused scan kernel (K1), CUDA events */
} vq_scan_counts;

/* target: [n_streams][n_splits][dim] fp32 HOST.  Synchronous: copies the target in, runs the
 * scan and its selection kernels, and publishes counts, top-k and the three ordered lists into
 * a pinned host mirror owned by the library (one stream synchronisation).  vq_fetch_* copy
 * from that mirror into caller buffers; vq_scan_host_list hands out read-only views of it.   */
int vq_scan(vq_store *s, const float *target, const vq_scan_params *p, vq_scan_counts *out);
/* The selection round's variant (ticket.py:311-356 samples a few dozen clips out of lists that can hold millions):
 * same scan, but only the counts, the top-k, the tie band and the BEST near miss (highest score, first in database
 * order, ticket.py:335-340; row -1 when there is none) cross PCIe; the caller draws list positions itself and fetches
 * exactly those with vq_gather_list (positions index the ordered match / near-miss / tie lists of this scan).        */
int vq_scan_select(vq_store *s, const float *target, const vq_scan_params *p, vq_scan_counts *out,
                   int64_t *near_best_pos, int64_t *near_best_row, float *near_best_score);
int vq_gather_list(vq_store *s, int32_t which, int64_t n_idx, const int64_t *positions, int64_t *rows_out,
                   float *scores_out);
/* The broker's arrangement — ONE process, one job at a time (reference src/broker.py:62-92) — with the search set sharded
 * over the GPUs of the box: the same scan on every shard from one call and one thread.  The target copy, K1, K2 and the
 * publish kernel are enqueued on every shard's own stream first, then each stream is waited for once; the shards'
 * ranked top-k lists are merged into the search set's (score descending, global row ascending).  lists = 1 behaves like
 * vq_scan on every shard (whole lists in the host mirrors), lists = 0 like vq_scan_select.  counts_out [n_shards];
 * near_best_out [n_shards][3] = position in the shard's near-miss list, global row (-1: none), fp32 score bits;
 * list positions are per shard: the shards' lists in shard order are the search set's lists in database order.
 * With lists = 1 the match and near-miss lists of ALL shards land back to back in one host mirror (every device writes its
 * segment at the offset the counts give): vq_scan_multi_host_list hands out views of the search set's lists, no host copy;
 * the tie band, the top-k and the counts stay per shard (vq_scan_host_list which = 2, 3).                             */
int vq_scan_multi(vq_store *const *shards, int32_t n_shards, const float *target, const vq_scan_params *p,
                  int32_t lists, vq_scan_counts *counts_out, int64_t *near_best_out, int32_t topk_cap,
                  int64_t *topk_rows_out, float *topk_scores_out, int32_t *n_topk_out);
/* Entries of the search set's lists after a vq_scan_multi: which[e] names the list (0 matches, 1 near misses, 2 ties),
 * positions[e] the position in it; gather kernels on all owning shards first, then one wait per shard.                */
int vq_gather_list_multi(vq_store *const *shards, int32_t n_shards, int64_t n_idx, const int32_t *which,
                         const int64_t *positions, int64_t *rows_out, float *scores_out);
int vq_scan_multi_host_list(vq_store *first_shard, int32_t which, const int64_t **rows, const float **scores, int64_t *n);
/* Same work, enqueued only: target already on the device, nothing copied back, no sync.
 * Used for device-side timing and for multi-GPU merges that read the results in place.      */
int vq_scan_enqueue(vq_store *s, const float *target_dev, const vq_scan_params *p, void *stream);
int vq_scan_wait(vq_store *s, void *stream, vq_scan_counts *out);

int vq_fetch_matches(vq_store *s, int64_t cap, int64_t *rows_out, float *scores_out);
int vq_fetch_near(vq_store *s, int64_t cap, int64_t *rows_out, float *scores_out);
int vq_fetch_ties(vq_store *s, int64_t cap, int64_t *rows_out, float *scores_out);
int vq_fetch_topk(vq_store *s, int32_t cap, int64_t *rows_out, float *scores_out);
/* Read-only view of the host mirror filled by the last vq_scan on this store: which = 0 matches,
 * 1 near misses, 2 tie band, 3 top-k; rows are GLOBAL, lists in database order (top-k ranked).
 * Valid until the next scan on this store; the caller must not free or write it.             */
int vq_scan_host_list(vq_store *s, int32_t which, const int64_t **rows, const float **scores, int64_t *n);
/* Full ranking of a list of the last scan (which = 0 matches, 1 near misses), sorted on the device: score descending,
 * database order among equal scores (the report order of ticket.py:266).                                         */
int vq_fetch_ranked(vq_store *s, int32_t which, int64_t cap, int64_t *rows_out, float *scores_out);
/* The same list sorted by (score descending, caller's tie-break ascending): the report order of ticket.py:266 is a stable
 * sort of the clips in SELECTION order (for a finalize round a seeded permutation of the lists, ticket.py:333,341), so
 * equal scores keep that order, not the database's.  tiebreak [n] = each list entry's rank in selection order (distinct
 * values < 2^32 - 16); tiebreak_out [n] = the same values in report order, scores_out their scores.                   */
int vq_rank_list(vq_store *s, int32_t which, int64_t n, const uint32_t *tiebreak, uint32_t *tiebreak_out,
                 float *scores_out);
int vq_fetch_scores(vq_store *s, int64_t first_row, int64_t n_rows, float *scores_out);
/* scores of the last scan at n arbitrary LOCAL rows in one round trip (the forced clips of ticket.py:346-356)          */
int vq_fetch_scores_at(vq_store *s, int64_t n, const int64_t *local_rows, float *scores_out);
int vq_fetch_sims(vq_store *s, int64_t first_row, int64_t n_rows, float *sims_out);

/* device views of the last scan's results (valid until the next scan on this store)         */
typedef struct vq_scan_device_view {
    const float *scores_dev;        /* [n_rows]                                               */
    const float *sims_dev;          /* [n_rows][n_streams] or NULL                            */
    const int64_t *counts_dev;      /* [4] = n_match, n_near, n_tie, n_topk                   */
    const float *topk_scores_dev;   /* [VQ_MAX_TOPK], -inf padded                             */
    const int64_t *topk_rows_dev;   /* [VQ_MAX_TOPK] GLOBAL rows, -1 padded                   */
    const uint32_t *match_rows_dev; /* [n_match] LOCAL rows ascending                         */
    const uint32_t *near_rows_dev;  /* [n_near]                                               */
    const float *match_scores_dev;  /* [n_match] scores of the listed rows, same order        */
    const float *near_scores_dev;   /* [n_near]                                               */
    const uint32_t *tie_rows_dev;   /* [n_tie]                                                */
    const float *tie_scores_dev;    /* [n_tie]                                                */
} vq_scan_device_view;
int vq_scan_view(vq_store *s, vq_scan_device_view *out);

/* Multi-GPU merge (the one collective of the path).  After a scan, each rank's payload
 * [4 + 2*topk] int64 = counts | top-k global rows | top-k score bits sits at *payload_dev; ranks
 * exchange it with one NCCL allgather and every rank merges the gathered buffer on its device:
 * merged_dev gets the same layout with counts summed and the global top-k.                    */
int vq_scan_payload(vq_store *s, const int64_t **payload_dev, int32_t *n_int64);
int vq_merge_payloads_enqueue(int device, const int64_t *gathered_dev, int32_t n_lists, int32_t topk,
                              int64_t *merged_dev, void *stream);

/* Peer-memory variant of the same exchange (one process per GPU, same node): every rank stores its payload
 * into all peers' inboxes over NVLink (CUDA IPC mappings), then waits for the others and merges — one
 * kernel behind the selection kernels, no NCCL call.  Setup: create, exchange the 64-byte local handles
 * between ranks by any means (torch.distributed all_gather, MPI, files), connect.                    */
typedef struct vq_exchange vq_exchange;
int vq_exchange_create(vq_exchange **out, int device, int world, int rank);
int vq_exchange_local_handle(vq_exchange *x, void *handle_out /* 64 bytes */);
int vq_exchange_connect(vq_exchange *x, const void *all_handles /* [world][64] */);
/* The same for exchanges that live in ONE process (all[r] = rank r's exchange; one per GPU with peer access between the
 * devices, or several on one device): the inboxes are addressed directly, no IPC mapping.                      */
int vq_exchange_connect_local(vq_exchange *const *all, int32_t world);
int vq_exchange_destroy(vq_exchange *x);
int vq_scan_exchange_enqueue(vq_store *s, vq_exchange *x, void *stream);
/* A stream of queries: push this step's payload, merge the PREVIOUS step's (no rank waits for the slowest rank of the
 * current step); vq_exchange_flush_enqueue merges the last pushed step.  vq_exchange_merged then holds step i-1 after
 * the i-th lagged call and the last step after the flush.                                                          */
int vq_scan_exchange_enqueue_lagged(vq_store *s, vq_exchange *x, void *stream);
int vq_exchange_flush_enqueue(vq_exchange *x, void *stream);
int vq_exchange_merged(vq_exchange *x, const int64_t **merged_dev /* [4 + 2*topk] */);
/* The exchange kernel runs on a stream of its own behind the scan stream (the next scan does not queue behind it, nor
 * behind the peer it may be waiting for); vq_exchange_flush_enqueue also makes `stream` wait for it, so call it before
 * reading the merged buffer on `stream`.  A peer that never delivers ends the kernel after VQ_EXCHANGE_TIMEOUT_S
 * (default 10 s) with negative counts in the merged buffer: vq_exchange_check (after a synchronisation) reports it.
 * vq_exchange_kernel_times: each exchange kernel's own time since the last call (ring of 256), in ms, taken by the kernel
 * on the global timer, optionally split into its three phases.                                                       */
int vq_exchange_check(vq_exchange *x);
int vq_exchange_kernel_times(vq_exchange *x, int32_t cap, float *ms_out, float *parts_out /* [cap][3]: push, wait, merge; or NULL */,
                             int32_t *n_out);

/* Host mailbox: all-gather of small records (up to slot_bytes each) between the rank processes of ONE box through a
 * POSIX shared-memory segment — the host-side twin of the exchange above, for what the host needs from its peers per
 * query: counts, top-k, tie band, best near miss, sampled list entries (a few KB; a collective library would stage
 * them through the devices with two stream synchronisations).  `name` ("/vq-<unique per job>") is created by rank 0
 * and opened by the others — the caller puts a barrier between the two — and can be unlinked as soon as all ranks are
 * attached.  vq_hostx_allgather is a collective: every rank makes the same calls with the same nbytes; all_out is
 * [world][nbytes]; it fails (instead of hanging) when a peer has not arrived within timeout_s (<= 0: 60 s).           */
typedef struct vq_hostx vq_hostx;
int vq_hostx_create(vq_hostx **out, const char *name, int world, int rank, int64_t slot_bytes);
int vq_hostx_unlink(vq_hostx *x);
int vq_hostx_allgather(vq_hostx *x, const void *mine, int64_t nbytes, void *all_out, double timeout_s);
int vq_hostx_destroy(vq_hostx *x);

/* Per-query summary record of the rank-level path (one process per GPU): what a rank tells its peers after its local
 * scan — first global row, counts, ranked top-k, best near miss, tie band (up to tie_cap entries) — as
 * int64 [5 + 2k + 3 + 2 tie_cap], and the merge of the gathered records of all ranks (host code; on the per-query path). */
int vq_summary_pack(int64_t first_row, const int64_t *counts3, int32_t k, int32_t n_topk, const int64_t *topk_rows,
                    const float *topk_scores, const int64_t *near_best, int32_t tie_cap, int32_t n_ties,
                    const int64_t *tie_rows, const float *tie_scores, int64_t *rec_out);
int vq_summary_merge(const int64_t *gathered, int32_t world, int32_t k, int32_t tie_cap, int64_t *first_rows_out,
                     int64_t *counts_out, int64_t *topk_rows_out, float *topk_scores_out, int32_t *n_topk_out,
                     int64_t *best_out, int64_t *tie_rows_out, float *tie_scores_out, int64_t *n_ties_out);

/* per-launch device times of K1 recorded since the last call (ring of 1024), in ms           */
int vq_scan_kernel_times(vq_store *s, int32_t cap, float *ms_out, int32_t *n_out);
/* the same ring split by phase: K1 (scan) and K2a-c (selection) per launch; either output may be NULL                */
int vq_scan_phase_times(vq_store *s, int32_t cap, float *scan_ms_out, float *select_ms_out, int32_t *n_out);
/* merge per-shard top-k lists (score desc, global row asc) — the host/rank-0 side of the
 * NCCL allgather merge; lists: [n_lists][k] with -inf / -1 padding.                          */
int vq_merge_topk(int32_t n_lists, int32_t k, const float *scores, const int64_t *rows,
                  float *scores_out, int64_t *rows_out, int32_t *n_out);
/* the same for Q queries at once (per-shard / per-rank results of vq_scan_batch): lists
 * [n_lists][n_queries][k] -> [n_queries][k] (padded with -inf / -1), n_out [n_queries].      */
int vq_merge_topk_batch(int32_t n_lists, int32_t n_queries, int32_t k, const float *scores,
                        const int64_t *rows, float *scores_out, int64_t *rows_out, int32_t *n_out);

/* ---------------------------------------------------------------- labelled subset, fp64 (A6, A8)
 * Similarities of a short list of rows against an fp64 target, fp64 accumulation
 * (the inputs of Hyperparameter.optimize_weights, hyperparameter.py:57-65).                  */
int vq_labelled_sims(vq_store *s, const double *target /* [S][P][dim] */, const int64_t *rows,
                     int64_t n, double *sims_out /* [n][S] */);
/* losses[r][iw][ith] for R replicate index sets over L labelled clips (hyperparameter.py:56-65);
 * replicate r uses labelled indices rep_index[rep_offset[r] .. rep_offset[r+1]).             */
int vq_loss_grid(int device, const double *sims /* [L][2] */, const uint8_t *labels, int64_t L,
                 const double *weight_grid, int32_t n_w, const double *threshold_grid,
                 int32_t n_th, double ballast, const int32_t *rep_offset, const int32_t *rep_index,
                 int32_t R, double *losses_out);

/* ---------------------------------------------------------------- target bootstrap, fp64 (A10)
 * New target from user-labelled rows: TargetClip._bootstrap_valid_matches (target_clip.py:161-199)
 * when n_invalid == 0, else _bootstrap_valid_plus_invalid (:201-261).  One (stream, split)
 * problem per slot, all solved in one call.  Resampling stays with the caller (Python RNG).  */
int vq_bootstrap_target(vq_store *s, const int64_t *valid_rows, int32_t n_valid,
                        const int64_t *invalid_rows, int32_t n_invalid, double mu,
                        const uint8_t *slot_mask /* [S][P] nonzero = solve this slot; NULL = all */,
                        double *target_out /* [S][P][dim]; slots not solved are zero */);
/* A singular system in a solved slot (a labelled clip listed twice, dependent rows: numpy.linalg.inv raises LinAlgError in
 * the reference, target_clip.py:194,248) or a non-finite result is an error, not a NaN target.                        */

/* ---------------------------------------------------------------- batched queries (tcgen05)
 * Q targets scored against the shard in one pass; per query: counts and top-k.               */
int vq_scan_batch(vq_store *s, const float *targets /* [Q][S][P][dim] */, int32_t n_queries,
                  const vq_scan_params *p, int64_t *counts_out /* [Q][2] match, near */,
                  int64_t *topk_rows_out /* [Q][topk] */, float *topk_scores_out /* [Q][topk] */,
                  float *kernel_ms_out);

/* The same plus every query's tie band (north_star: ties within COMPUTE_EPS of a boundary are reported): rows whose
 * fp32 score is within p->eps of the threshold or of the near-miss limit, compared in double like the single-query scan.
 * tie_counts_out [Q] (exact); tie_rows_out / tie_scores_out [Q][tie_cap] in database order, -1 / -inf padded, may be
 * NULL; at most 4096 entries per query are kept.  p->eps <= 0: counts are zero and the faster kernel runs.            */
int vq_scan_batch_ties(vq_store *s, const float *targets, int32_t n_queries, const vq_scan_params *p,
                       int64_t *counts_out, int64_t *topk_rows_out, float *topk_scores_out,
                       int64_t *tie_counts_out, int32_t tie_cap, int64_t *tie_rows_out, float *tie_scores_out,
                       float *kernel_ms_out);

/* Test hook: the full [Q][n_rows] fp32 score matrix of the batched path (small shards only). */
int vq_scan_batch_scores(vq_store *s, const float *targets, int32_t n_queries, const vq_scan_params *p,
                         float *scores_out);

/* ---------------------------------------------------------------- seeded sampling (host code)
 * `random.sample(range(n), k)` on CPython's Mersenne Twister state (ticket.py:333,341 draw the review set with it; the
 * finalize round permutes EVERY match, one interpreter loop iteration per clip): same picks, same state afterwards.
 * mt_state [624] and mt_pos are random.getstate()[1][:624] and [624], updated in place.                              */
int vq_mt_sample_range(uint32_t *mt_state, int32_t *mt_pos, int64_t n, int64_t k, int64_t *picks_out);

/* ---------------------------------------------------------------- feature-CSV ingest (host code)
 * The files the TSN extractor writes and load_db.py loads (src/api/api_load_records.py:41-58:
 * csv.reader + float() per cell): header line of 5 `key =value` fields, then `clip_no, f0 .. f{dim-1}`
 * per row.  vq_csv_shape returns the number of data rows, the feature length and the header line;
 * vq_csv_read parses the rows with n_threads threads (0 = all cores) into caller-owned arrays —
 * correctly rounded doubles, identical to Python's float().  Malformed rows are an error.       */
int vq_csv_shape(const char *path, int64_t *n_rows_out, int32_t *dim_out, char *header_out, int32_t header_cap);
int vq_csv_read(const char *path, int32_t n_threads, int64_t cap_rows, int32_t dim,
                int64_t *clip_numbers_out /* [n] */, double *features_out /* [n][dim] */, int64_t *n_rows_out);

#ifdef __cplusplus
}
#endif
#endif /* VQ_B200_H */
