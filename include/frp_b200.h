/*
 * frp_b200.h — C ABI of the B200-native recognition hot path (align + 1:N match).
 *
 * The reference (achiever04/face-recognition-platform) is a pure-Python service
 * with no FFI; the seam this library sits behind is the Python singleton
 * `face_service` (backend/app/services/face_service.py:769) and the gallery dict
 * `ENCODINGS` (backend/app/state.py:78).  Each entry point below names the
 * reference code it replaces.  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - every function returns 0 (FRP_OK) or a negative FRP_E_* code; the message
 *     of the last error of the calling thread is frp_last_error().
 *   - "_device" variants take device pointers and a cudaStream_t (passed as
 *     void*; NULL = the legacy default stream) and are asynchronous.
 *   - host variants copy host<->device inside the call and return synchronised.
 *   - there is NO CPU fallback: without a usable sm_100 device every compute
 *     entry point fails with FRP_E_CUDA / FRP_E_NODEVICE.
 *   - scores: metric L2 -> Euclidean distance (smaller is better, match iff
 *     d <= tau, face_service.py:410-411); metric COS -> cosine similarity
 *     (larger is better, match iff s >= tau; insightface compute_sim).
 *   - ties are broken by gallery row (= dict insertion order; the reference's
 *     sort is stable over dict order, face_service.py:431).
 */
#ifndef FRP_B200_H
#define FRP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FRP_ABI_VERSION 2

/* error codes */
#define FRP_OK            0
#define FRP_E_INVALID    -1   /* bad argument */
#define FRP_E_CUDA       -2   /* CUDA runtime error (see frp_last_error) */
#define FRP_E_NODEVICE   -3   /* no sm_100 device visible */
#define FRP_E_NOMEM      -4
#define FRP_E_IO         -5
#define FRP_E_UNSUPPORTED -6
#define FRP_E_RANGE      -7   /* row index out of range / dead row */

/* metrics */
#define FRP_METRIC_L2   0   /* face_recognition.face_distance, face_service.py:410,597 */
#define FRP_METRIC_COS  1   /* insightface ArcFaceONNX.compute_sim (third-party) */

/* frp_match flags */
#define FRP_MATCH_THRESHOLD   0x0  /* threshold then top-k (routes/face.py:692-693) */
#define FRP_MATCH_NO_THRESHOLD 0x1 /* pure top-k (face_service.py:590-612 find_k_nearest) */
#define FRP_MATCH_FORCE_SCAN  0x2  /* use the streaming scan kernel for every batch size */
#define FRP_MATCH_FORCE_GEMM  0x4  /* use the tcgen05 kernel for every batch size */
#define FRP_MATCH_FORCE_EXACT 0x8  /* scan the fp32 master instead of the fp16 copy */
#define FRP_MATCH_STRICT      0x10 /* strict comparison (d < tau): store_face duplicate check, face_service.py:360 */
#define FRP_MATCH_WIDE        0x20 /* frp_match_device: widest candidate lists (64 per CTA); second-level fallback */
#define FRP_MATCH_NO_SAMPLE   0x40 /* pure top-k: skip the sampling pass that lifts the prefilter cut (A/B measurements) */

/* per-query status written by frp_match */
#define FRP_STATUS_OK        0  /* result certified exact w.r.t. the fp32 master */
#define FRP_STATUS_FALLBACK  1  /* fp16 prefilter could not be certified; result came from the fp32 scan */
#define FRP_STATUS_UNCERTIFIED 2 /* still uncertified after both fallbacks (> 64 exact ties at the cut):
                                    the caller must use frp_scores for this query */

typedef struct frp_gallery frp_gallery_t;   /* opaque; replaces state.ENCODINGS (state.py:78) */

/* ---- library / device -------------------------------------------------- */
int         frp_abi_version(void);
const char* frp_last_error(void);
/* number of visible CUDA devices; 0 and FRP_E_NODEVICE when there is none */
int         frp_device_count(int* count);
/* name, SM count, global memory bytes and compute capability of `device` */
int         frp_device_info(int device, char* name, size_t name_len, int* sm_count,
                            uint64_t* mem_bytes, int* cc_major, int* cc_minor);

/* ---- gallery store: replaces ENCODINGS (state.py:78), the per-call
 *      np.array([...]) rebuild (face_service.py:409) and store/delete
 *      (face_service.py:374, :521-523) ---------------------------------------- */
/* dim: embedding length (128 / 512 / ...); capacity: initial row capacity
 * (grows on demand).  HBM layout: fp16 scan copy [cap][dimp], fp32 master
 * [cap][dimp], |g|^2 f32 [cap], alive u8 [cap]; COS rows are L2-normalised on
 * ingest.  dimp = dim rounded up to a power of two >= 64. */
int frp_gallery_create(int device, int dim, int metric, uint64_t capacity, frp_gallery_t** out);
int frp_gallery_destroy(frp_gallery_t* g);
/* append n rows (host, row-major [n][dim]); *first_row receives the row index
 * of the first appended row.  Row index == insertion sequence == tie order. */
int frp_gallery_append_f32(frp_gallery_t* g, const float* rows, uint64_t n, uint64_t* first_row);
int frp_gallery_append_f64(frp_gallery_t* g, const double* rows, uint64_t n, uint64_t* first_row);
/* same, rows already in device memory (fp32, row-major [n][dim]) */
int frp_gallery_append_device(frp_gallery_t* g, const float* rows_dev, uint64_t n,
                              uint64_t* first_row, void* stream);
/* overwrite a live row in place (dict assignment to an existing key keeps its
 * position, face_service.py:374) */
int frp_gallery_update_f32(frp_gallery_t* g, uint64_t row, const float* vec);
int frp_gallery_update_f64(frp_gallery_t* g, uint64_t row, const double* vec);
/* tombstone a row (del ENCODINGS[t], face_service.py:521-523) */
int frp_gallery_remove(frp_gallery_t* g, uint64_t row);
/* rows = appended rows incl. tombstones; live = rows - tombstones */
int frp_gallery_count(const frp_gallery_t* g, uint64_t* rows, uint64_t* live);
int frp_gallery_info(const frp_gallery_t* g, int* device, int* dim, int* dimp, int* metric,
                     uint64_t* capacity);
/* new gallery holding rows idx[0..n) of g in that order (duplicates allowed):
 * compaction after deletes, and the target_names subset of face_service.py:403 */
int frp_gallery_gather(const frp_gallery_t* g, const uint64_t* idx, uint64_t n, frp_gallery_t** out);
/* read back the fp32 master row (COS: normalised) */
int frp_gallery_get_row(const frp_gallery_t* g, uint64_t row, float* out);
/* flat binary snapshot: header + fp32 master rows + alive flags.  save writes
 * `path + ".tmp"` and renames it over `path`; load validates the header against the file size
 * and rebuilds the fp16 copy / norms / error bound from the stored master VERBATIM (no
 * re-normalisation), so a loaded gallery answers bit-identically to the saved one. */
int frp_gallery_save(const frp_gallery_t* g, const char* path);
int frp_gallery_load(int device, const char* path, frp_gallery_t** out);
/* global row index of local row 0 (row-sharded galleries); added to out_idx */
int frp_gallery_set_base(frp_gallery_t* g, uint64_t base_row);
/* Mutable row-sharded galleries ("appends to the least-full shard with global seq
 * kept per row"): give rows [first_row, first_row + n) explicit global sequence numbers.  Once
 * a gallery carries sequences, out_idx reports seq[row] instead of base_row + row, and the
 * cross-shard merge breaks ties by it.  Sequences must increase with the local row (append in
 * global insertion order), so the local tie order equals the global one.  Not persisted by
 * frp_gallery_save (the sharded dict keeps them with its names). */
int frp_gallery_set_seq(frp_gallery_t* g, uint64_t first_row, uint64_t n, const uint64_t* seq);

/* ---- 1:N match: replaces face_distance + `<= tolerance` + sort + [:top_k]
 *      (face_service.py:410-411,431-432; routes/face.py:692-693) and the
 *      argpartition top-k of find_k_nearest (face_service.py:601-603) ------- */
/* queries [B][dim] fp32 host.  Outputs (host): out_idx [B][k] (global row, -1
 * pad), out_score [B][k] (distance or similarity, NaN pad), out_count [B]
 * (results per query <= k), out_status [B] (FRP_STATUS_*; may be NULL). */
int frp_match(frp_gallery_t* g, const float* queries, int B, double tau, int k, int flags,
              int64_t* out_idx, float* out_score, int32_t* out_count, int32_t* out_status);
/* same with device buffers; asynchronous on `stream`.  Queries are uncertified
 * here: out_status must be checked by the caller (FRP_STATUS_FALLBACK rows must
 * be re-run through frp_match_device with FRP_MATCH_FORCE_EXACT). */
int frp_match_device(frp_gallery_t* g, const float* queries_dev, int B, double tau, int k,
                     int flags, int64_t* out_idx_dev, float* out_score_dev,
                     int32_t* out_count_dev, int32_t* out_status_dev, void* stream);
/* all N scores of every query from the fp32 master in float64 (the O(N) dict
 * list compare_faces returns, face_service.py:414-429).  out [B][out_stride] (row b of the
 * result starts at out + b * out_stride); the gallery's row count at the time of the call is
 * returned in *rows_out (may be NULL) and must not exceed out_stride (FRP_E_RANGE otherwise,
 * nothing written): the count is read and used under one lock, so a concurrent store_face
 * can never make the call write past the caller's buffer.  NaN marks deleted rows. */
int frp_scores(frp_gallery_t* g, const float* queries, int B, double* out, uint64_t out_stride,
               uint64_t* rows_out);
int frp_scores_device(frp_gallery_t* g, const float* queries_dev, int B, double* out_dev,
                      uint64_t out_stride, uint64_t* rows_out, void* stream);
/* Row-sharded galleries (one rank per GPU): frp_match_record_device runs the local match and
 * writes ONE packed record straight into the caller's all-gather send buffer
 *     idx[B][k] i64 | score[B][k] f64 | count[B] i32 | status[B] i32   (padded to 16 B)
 * of frp_shard_record_bytes(B,k) bytes (scores travel in float64 so the cross-shard merge
 * ranks exactly like the single-gallery re-rank); ncclAllGather concatenates the R records;
 * frp_merge_packed_device merges them into the global top-k [B][k] (better score first, ties
 * by global row = gallery base_row + local row) and ORs the per-shard status words into
 * *out_any_status (may be NULL).  out_score64 (may be NULL) receives the float64 scores. */
size_t frp_shard_record_bytes(int B, int k);
int frp_match_record_device(frp_gallery_t* g, const float* queries_dev, int B, double tau, int k,
                            int flags, void* record_dev, void* stream);
int frp_merge_packed_device(int metric, int R, int B, int k, const void* records_dev,
                            int64_t* out_idx, float* out_score, double* out_score64,
                            int32_t* out_count, int32_t* out_any_status, void* stream);

/* Fused exchange + merge over peer memory (NVLink/NVSwitch), replacing the all-gather:
 * every rank owns an exchange buffer of frp_exchange_buffer_bytes(R,B,k) bytes that is
 * mapped into all ranks (CUDA IPC / symmetric memory; zero-initialised once).
 * peers_dev is a DEVICE array of the R mapped base pointers (index = rank).  One kernel
 * pushes this rank's packed record into slot [epoch&1][rank] of every peer, publishes a
 * system-scope flag, waits for the R flags of its own buffer and merges.  `epoch` must
 * increase by one per call, identically on all ranks; counter_dev is a zero-initialised
 * device word private to this rank.  Bit 0x100 of *out_any_status = a peer timed out. */
size_t frp_exchange_buffer_bytes(int R, int B, int k);
int frp_push_merge_device(int metric, int R, int rank, int B, int k, const void* send_record_dev,
                          void* const* peers_dev, uint32_t epoch, unsigned int* counter_dev,
                          int64_t* out_idx, float* out_score, int32_t* out_count,
                          int32_t* out_any_status, void* stream);

/* Number of SMs the tcgen05 prefilter leaves out of its grid (default 0).  Its CTAs occupy a whole
 * SM's shared memory, so work on another stream that should overlap it — the all-gather + merge
 * of the previous step in the pipelined sharded matcher — only runs concurrently if a few SMs
 * stay free.  Process-wide; ShardedMatcher sets 2 for its pipelined mode on more than one rank. */
int frp_set_reserved_sms(int n);

/* ---- gallery against itself: cluster_faces (face_service.py:552-585) -------------------
 * Greedy seed clustering on the device gallery: rows in insertion order; an unassigned live
 * row seeds a cluster and absorbs every unassigned live row within tau (inclusive `<=` / `>=`,
 * strict with FRP_MATCH_STRICT).  The gallery is matched against itself in slabs of up to 256
 * seed candidates through the same prefilter + exact re-rank as frp_match (tcgen05 kernel);
 * only the thresholded neighbour lists come back to the host, which runs the greedy pass.
 * labels [rows] receives the cluster id of every row (ids number the seeds in order; -1 for
 * deleted rows), *n_clusters the number of clusters.  stats (may be NULL) receives
 * {slabs, rows matched on the device, rows that needed the dense fallback, kernel launches}. */
int frp_cluster(frp_gallery_t* g, double tau, int flags, int32_t* labels, uint64_t labels_len,
                int32_t* n_clusters, uint64_t* stats4);

/* ---- match -> tracking hand-off: cooldown table on the device ---------
 * tracking_service.record_detection drops a (person, camera) detection that arrives within
 * `cooldown` seconds of the last RECORDED one (tracking_service.py:118-134, :179).  The table
 * is an open-addressing hash in HBM keyed by (global gallery row, camera id); the filter runs
 * on the match results while they are still on the device, so cooled-down hits are never
 * copied back or turned into Python objects. */
typedef struct frp_cooldown frp_cooldown_t;
/* capacity = number of (row, camera) slots (rounded up to a power of two, >= 1024) */
int frp_cooldown_create(int device, uint64_t capacity, frp_cooldown_t** out);
int frp_cooldown_destroy(frp_cooldown_t* c);
int frp_cooldown_clear(frp_cooldown_t* c, void* stream);
/* Inputs (device): idx [B][k] / score [B][k] / count [B] as written by frp_match_device;
 * camera [B] i32 = camera id of each query; limit [B] f32 (may be NULL) = per-query second
 * filter `distance <= limit` (`similarity >= limit` for metric COS), routes/camera.py:250.
 * A hit (b, j) is RECORDED when it passes the limit and its (row, camera) key was last
 * recorded more than `cooldown` seconds before `now` (or never); within one call the first
 * hit in (b, j) order wins a key, like the reference's sequential loop.  Outputs (device):
 * hits [B*k] records {int32 b, int32 j, int64 idx, float score, int32 camera} compacted in
 * (b, j) order, *n_hits.  A full table drops nothing: the hit is recorded and bit 0 of
 * *overflow is set. */
typedef struct frp_hit { int32_t query; int32_t rank; int64_t row; float score; int32_t camera; } frp_hit_t;
int frp_cooldown_filter_device(frp_cooldown_t* c, int metric, const int64_t* idx_dev,
                               const float* score_dev, const int32_t* count_dev,
                               const int32_t* camera_dev, const float* limit_dev, int B, int k,
                               double now, double cooldown, frp_hit_t* hits_dev, int32_t* n_hits_dev,
                               int32_t* overflow_dev, void* stream);
/* host-buffer convenience: match + filter in one call, only the recorded hits come back.
 * hits [B*k]; returns the number of hits in *n_hits. */
int frp_match_hits(frp_gallery_t* g, frp_cooldown_t* c, const float* queries, int B, double tau, int k,
                   int flags, const int32_t* camera, const float* limit, double now, double cooldown,
                   frp_hit_t* hits, int32_t* n_hits);

/* ---- alignment: insightface face_align.norm_crop + blobFromImages
 *      (third-party; only call site backend/app/utils/deepfake_utils.py:138) -- */
/* 2x3 forward similarity matrices from 5 landmarks (skimage _umeyama onto
 * arcface_dst, float64).  lmk [n][5][2] fp32 host -> M [n][6] f64 host. */
int frp_estimate_norm(const float* lmk, int n, double* M_out);
/* The estimate as insightface 0.7.3 actually evaluates it: float32 landmarks and template make
 * skimage's _umeyama run its means, covariance and variance in float32 (numpy keeps the dtype)
 * and only the composition in float64.  Same closed form on those float32 intermediates;
 * agrees with skimage-on-float32 to ~1e-6 relative (LAPACK's float32 SVD differs in the last
 * bits).  Pass the result as `M` to frp_align to reproduce insightface's crops within 1 LSB. */
int frp_estimate_norm_f32(const float* lmk, int n, double* M_out);
/* frames: F BGR u8 images [F][H][stride] (stride >= 3*W bytes, host);
 * frame_idx [n]; lmk [n][5][2] fp32 (or NULL when M [n][6] f64 is given).
 * out_blob [n][3][112][112] fp16 (RGB planes, (v-127.5)/127.5);
 * out_aligned [n][112][112][3] u8 BGR (optional, may be NULL). */
int frp_align(int device, const uint8_t* frames, int F, int H, int W, size_t stride,
              const int32_t* frame_idx, const float* lmk, const double* M, int n,
              uint16_t* out_blob, uint8_t* out_aligned);
int frp_align_device(int device, const uint8_t* frames_dev, int F, int H, int W, size_t stride,
                     const int32_t* frame_idx_dev, const float* lmk_dev, const double* M_dev,
                     int n, uint16_t* out_blob_dev, uint8_t* out_aligned_dev, void* stream);

/* ---- instrumentation ----------------------------------------------------- */
/* kernels launched by this library since load (all threads) */
uint64_t frp_kernel_launches(void);
/* per-kernel timing (CUDA events on the caller's stream) of the matches issued while
 * enabled; frp_profile_last waits for the last match of the calling thread and
 * returns the device time of its three stages in milliseconds */
int frp_profile_enable(int on);
int frp_profile_last(float* prep_ms, float* prefilter_ms, float* rerank_ms);
/* same for the alignment kernel of frp_align_device */
int frp_profile_align_enable(int on);
int frp_profile_align_last(float* kernel_ms);
/* name of the kernel family the last frp_match* call of this thread used:
 * "small_exact" (<= 4 096 rows: one fused exact launch), "scan_f16", "scan_f32", "gemm_tcgen05" */
const char* frp_last_match_path(void);

#ifdef __cplusplus
}
#endif
#endif /* FRP_B200_H */
