/*
 * flechasdb_b200.h -- C ABI of libflechasdb_b200.so, the B200 (sm_100a) engine
 * behind flechasdb's IVF-PQ build and query path.
 *
 * The reference (codemonger-io/flechasdb) is pure Rust with no FFI; this header
 * is the boundary its hot-path bodies bind to (INTEGRATION.md shows the Rust
 * `extern "C"` block and build.rs).  Every entry point names the reference code
 * it replaces (paths relative to the reference repo root).
 *
 * Conventions
 *  - plain C: pointers + sizes, no exceptions cross the boundary;
 *  - every function returns FDB_OK (0) or a negative FDB_ERR_*; the message is
 *    kept per thread and read with fdb_last_error();
 *  - the caller owns all host buffers; the library owns device memory behind
 *    opaque handles; one handle is used from one thread at a time;
 *  - there is NO CPU fallback: without a CUDA device fdb_ctx_create fails.
 *  - "nb" = number of independent k-means problems solved side by side on the
 *    strided sub-vector views of one vector set (nb = 1 for the coarse
 *    quantiser, nb = D for the PQ codebooks of all divisions at once).  All
 *    per-problem arrays are laid out [nb][...].
 */
#ifndef FLECHASDB_B200_H
#define FLECHASDB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FDB_OK 0
#define FDB_ERR_INVALID_ARGS (-1)    /* Error::InvalidArgs      src/error.rs:7   */
#define FDB_ERR_INVALID_DATA (-2)    /* Error::InvalidData      src/error.rs:9   */
#define FDB_ERR_INVALID_CONTEXT (-3) /* Error::InvalidContext   src/error.rs:11  */
#define FDB_ERR_EMPTY_CLUSTER (-4)   /* panic: assert_ne!(count, 0) src/kmeans.rs:259 */
#define FDB_ERR_WEIGHTS (-5)         /* panic: WeightedIndex .unwrap() src/kmeans.rs:199,207,216 */
#define FDB_ERR_NAN (-6)             /* panic: min_index.unwrap() src/kmeans.rs:304; partial_cmp().unwrap() src/db/stored.rs:385,426 */
#define FDB_ERR_CUDA (-7)            /* CUDA runtime failure (no reference analogue) */
#define FDB_ERR_UNSUPPORTED (-8)     /* shape outside what the device layout supports (e.g. C > 256 with u8 codes) */
#define FDB_ERR_NCCL (-9)            /* NCCL failure or libnccl missing (multi-GPU entry points only) */

#define FDB_KMEANS_MAX_ROUNDS 100    /* const R: usize = 100;   src/kmeans.rs:114 */
#define FDB_KMEANS_EPSILON 1e-6f     /* f32 default_epsilon     src/kmeans.rs:24-28 */

#define FDB_QUERY_STORED 0           /* stored::Database::query  (NBestByKey selection) */
#define FDB_QUERY_BUILD 1            /* build::Database::query   (stable sort + truncate) */

typedef struct fdb_ctx fdb_ctx;     /* one CUDA device + stream                                */
typedef struct fdb_vs fdb_vs;       /* BlockVectorSet<f32> resident in HBM   src/vector.rs:28-100 */
typedef struct fdb_km fdb_km;       /* nb k-means problems (Codebook<f32> each) src/kmeans.rs:62-68 */
typedef struct fdb_index fdb_index; /* queryable IVF-PQ index: stored::Database src/db/stored.rs:41-57 */
typedef struct fdb_comm fdb_comm;   /* one rank of an NCCL communicator, bound to a context (no reference analogue) */

/* ---- library / context ------------------------------------------------------ */
const char *fdb_last_error(void);
int fdb_version(void);
int fdb_device_count(void);
int fdb_ctx_create(int device, fdb_ctx **out);
void fdb_ctx_destroy(fdb_ctx *ctx);
int fdb_ctx_sync(fdb_ctx *ctx);
/* CUDA events on the library's stream, for callers that time device work */
int fdb_ctx_timer_start(fdb_ctx *ctx);
int fdb_ctx_timer_stop(fdb_ctx *ctx, float *milliseconds);
/* number of kernels this context has launched so far */
uint64_t fdb_ctx_launch_count(const fdb_ctx *ctx);

/* ---- vector sets: BlockVectorSet::chunk + VectorSet  (src/vector.rs:40-100) ---- */
/* copies n*dim floats host -> device.  Fails like chunk() on a zero dim. */
int fdb_vs_upload(fdb_ctx *ctx, const float *rows, size_t n, size_t dim, fdb_vs **out);
/* borrows caller-owned device memory (row-major n*dim, 16-byte aligned) */
int fdb_vs_from_device(fdb_ctx *ctx, float *device_rows, size_t n, size_t dim, fdb_vs **out);
/* synthetic uniform [0,1) rows generated on the device (bench / tests):
 * element e of the set = (splitmix64(seed, start+e) >> 40) * 2^-24 */
int fdb_vs_generate(fdb_ctx *ctx, size_t n, size_t dim, uint64_t seed, uint64_t start, fdb_vs **out);
int fdb_vs_download(fdb_vs *vs, float *rows);
int fdb_vs_download_rows(fdb_vs *vs, size_t first_row, size_t nrows, float *rows);
size_t fdb_vs_len(const fdb_vs *vs);         /* VectorSet::len          */
size_t fdb_vs_vector_size(const fdb_vs *vs); /* VectorSet::vector_size  */
float *fdb_vs_device_ptr(fdb_vs *vs);
void fdb_vs_destroy(fdb_vs *vs);
/* v_j -= centroid[indices[j]] in place: Partitioning::partition_with_events,
 * src/partitions.rs:128-138.  km must be an nb=1 problem over the full rows. */
int fdb_vs_subtract_assigned(fdb_vs *vs, const fdb_km *km);

/* ---- k-means: src/kmeans.rs ---------------------------------------------------- */
/* Problem b clusters the SubVectorSet (src/vector.rs:103-149) made of columns
 * [col_off + b*dim, col_off + (b+1)*dim) of vs.  divide_vector_set
 * (src/vector.rs:154-174) is fdb_kmeans_begin(vs, 0, N/D, D, C).
 * Err(InvalidArgs) when n < k, like cluster_with_events (src/kmeans.rs:116-120). */
int fdb_kmeans_begin(fdb_vs *vs, size_t col_off, size_t dim, size_t nb, size_t k, fdb_km **out);
void fdb_kmeans_destroy(fdb_km *km);

/* k-means++ seeding, initialize_centroids (src/kmeans.rs:142-229), one call per
 * reference statement group so the host keeps the RNG and the loop:
 *   seed_first : ci = rng.gen_range(0..n); first centre; initial D^2 weights (:172-199)
 *   seed_total : WeightedIndex::total_weight (src/distribution.rs:45,76)
 *   seed_pick  : WeightedIndex::sample for the draw u in [0,1) (src/distribution.rs:104-121;
 *                the sample value is u*total as rand 0.8.5 UniformFloat<f32> computes it)
 *   seed_add   : round i for the chosen vector ci (:203-220)
 * exact != 0 keeps the reference's sequential f32 running total and cumulative
 * scan bit for bit (slow, single-threaded per problem); exact == 0 uses a
 * deterministic parallel scan (same distribution, not bit-comparable at
 * cumulative-sum boundaries).  ci / u01 / totals are arrays of nb. */
int fdb_kmeans_seed_first(fdb_km *km, const uint32_t *ci);
int fdb_kmeans_seed_total(fdb_km *km, float *totals);
int fdb_kmeans_seed_pick(fdb_km *km, const float *u01, int exact, uint32_t *ci_out);
int fdb_kmeans_seed_add(fdb_km *km, size_t i, const uint32_t *ci, int exact);
/* the whole seeding loop on the device without host round trips:
 * first[nb], u01[nb][k-1] -> picked[nb][k] (may be NULL) */
int fdb_kmeans_seed_run(fdb_km *km, const uint32_t *first, const float *u01, int exact,
                        uint32_t *picked);
/* the same loop with the picks injected ("k-means++ seeds fixed"): chosen[nb][k] */
int fdb_kmeans_seed_chosen(fdb_km *km, const uint32_t *chosen);
/* Rows sharded over several processes (multi-GPU build).  The host combines the shards:
 *   seed_round_ext : round i with the chosen vectors given by value, centres[nb][dim]
 *                    (host), local_ci[b] = the vector's index in THIS shard or 0xFFFFFFFF;
 *                    afterwards seed_total returns this shard's totals
 *   seed_pick_value: WeightedIndex::sample for an absolute sample value inside this shard's
 *                    cumulative weights (negative = the draw falls into another shard) */
int fdb_kmeans_seed_round_ext(fdb_km *km, size_t i, const float *centres, const uint32_t *local_ci);
int fdb_kmeans_seed_pick_value(fdb_km *km, const float *sample_values, uint32_t *ci_out);
/* The same sharded k-means++ without host round trips: every call only enqueues kernels on the
 * context's stream (fdb_ctx_stream), the caller runs the all-gathers (NCCL) on that stream in between.
 *   begin  -> device addresses of: totals [nb], pick [nb], centre_send [nb][dim], u01 [nb] (the caller
 *             copies this round's draws there), picked_global [k][nb]
 *   round 0: first(local index of the first centre or 0xFFFFFFFF) -> all-gather pick, centre_send -> round(0, ..)
 *   round i: total -> all-gather totals [world][nb] -> pick(all_totals) -> all-gather pick [world][nb] and
 *            centre_send [world][nb][dim] -> round(i, all_picks, all_centres)
 *   finish -> the picked global indices [nb][k] (the only host synchronisation) */
void *fdb_ctx_stream(fdb_ctx *ctx);
int fdb_kmeans_seed_sharded_begin(fdb_km *km, float **d_totals, uint32_t **d_pick, float **d_centre_send,
                                  float **d_u01, uint32_t **d_picked_global);
int fdb_kmeans_seed_sharded_first(fdb_km *km, const uint32_t *local_first);
int fdb_kmeans_seed_sharded_total(fdb_km *km);
int fdb_kmeans_seed_sharded_pick(fdb_km *km, const float *d_all_totals, int world, int rank);
int fdb_kmeans_seed_sharded_round(fdb_km *km, size_t i, const uint32_t *d_all_picks, const float *d_all_centres,
                                  int world, int rank, size_t n_global);
int fdb_kmeans_seed_sharded_finish(fdb_km *km, uint32_t *picked_global);
/* test hook / resume: set centroids [nb][k][dim] and (optionally) indices [nb][n] */
int fdb_kmeans_set_state(fdb_km *km, const float *centroids, const uint32_t *indices);

/* update_centroids (src/kmeans.rs:232-276): members are added in ascending vector
 * index, scaled by 1/count; gradients[nb] = max||old-new|| / max||new||.
 * FDB_ERR_EMPTY_CLUSTER mirrors the reference's assert.  active (nb bytes, may be
 * NULL = all) selects the problems to step. */
int fdb_kmeans_update(fdb_km *km, const uint8_t *active, float *gradients);
/* reassign_centroids (src/kmeans.rs:279-306): argmin_j dot(v-c_j, v-c_j) in the
 * reference's summation order, lowest j wins ties. */
int fdb_kmeans_reassign(fdb_km *km, const uint8_t *active);
/* the loop of cluster_with_events (src/kmeans.rs:125-137) for all nb problems:
 * gradients[nb][max_rounds] receives the FinishedCentroidUpdate values,
 * rounds[nb] the number of updates, reassigns[nb] the number of reassignments. */
int fdb_kmeans_run(fdb_km *km, size_t max_rounds, float epsilon, float *gradients,
                   uint32_t *rounds, uint32_t *reassigns);
/* how the last reassignment ran: on the tensor pipe (GEMM filter + exact re-check of the
 * rows with several candidates) or on the exact fp32 kernel; the row counts are only
 * collected when the environment variable FDB_TC_STATS is set. */
int fdb_kmeans_last_assign_info(fdb_km *km, uint32_t *used_tensor_cores, uint32_t *rechecked_rows,
                                uint32_t *overflow_rows);
/* Codebook { centroids, indices }: centroids [nb][k][dim], indices [nb][n] */
int fdb_kmeans_get(fdb_km *km, float *centroids, uint32_t *indices);
int fdb_kmeans_get_weights(fdb_km *km, float *weights); /* [nb][n] D^2 weights */

/* multi-GPU build (rows sharded across processes, centroids replicated): the
 * per-rank half of update_centroids.  update_partial leaves [nb][k][dim] sums
 * followed by [nb][k] counts (as floats) in a device buffer the caller
 * all-reduces (NCCL sum) in place; update_finish divides and computes the gradient. */
int fdb_kmeans_update_partial(fdb_km *km, float **device_buf, size_t *nfloats);
int fdb_kmeans_update_finish(fdb_km *km, float *gradients);
/* The same loop without host round trips (everything is enqueued on fdb_ctx_stream): begin; per round
 * partial_async -> all-reduce of the buffer on that stream -> finish_async (divide, gradient, device-side
 * convergence flags, reassignment of the active problems); poll reads the flags (the caller chooses how
 * often: converged problems are frozen on the device, extra rounds do nothing); end returns what
 * fdb_kmeans_run returns, gradients [nb][FDB_KMEANS_MAX_ROUNDS]. */
int fdb_kmeans_sharded_loop_begin(fdb_km *km);
int fdb_kmeans_sharded_partial_async(fdb_km *km, float **device_buf, size_t *nfloats);
int fdb_kmeans_sharded_finish_async(fdb_km *km, float epsilon);
int fdb_kmeans_sharded_poll(fdb_km *km, uint8_t *active /*[nb]*/);
int fdb_kmeans_sharded_loop_end(fdb_km *km, float *gradients, uint32_t *rounds, uint32_t *reassigns);

/* ---- index + query: src/db/build.rs:446-482, src/db/stored.rs:331-442,549-597 ---- */
/* Host-side constructor (what load_database/load_partition feed):
 *   coarse    [P][N]       partition centroids
 *   codebooks [D][C][N/D]  PQ code vectors
 *   offsets   [P+1]        partition p holds vectors offsets[p]..offsets[p+1]
 *   codes     [M][D] u8    partition-major, ascending vector index inside a partition
 *                          (Partition::new order, src/db/build.rs:459-473) */
int fdb_index_create(fdb_ctx *ctx, size_t N, size_t P, size_t D, size_t C, const float *coarse,
                     const float *codebooks, const uint64_t *offsets, const uint8_t *codes,
                     fdb_index **out);
/* stored::Database loads its partitions lazily (get_partition, src/db/stored.rs:269-293): create_lazy uploads the
 * partition centroids and the codebooks only (what load_database reads, src/db/stored.rs:659-798); a partition's
 * code list ([n][D] u8, ascending vector index) is uploaded by set_partition when the host first needs it -- a
 * partition that has not arrived is an empty list to the kernels.  missing_partitions runs the probe selection of a
 * batch and returns the distinct probed partitions that have not been loaded (ascending; out may be NULL, up to
 * cap are written, *n_missing is their number): the host loads them, then queries. */
int fdb_index_create_lazy(fdb_ctx *ctx, size_t N, size_t P, size_t D, size_t C, const float *coarse,
                          const float *codebooks, fdb_index **out);
int fdb_index_set_partition(fdb_index *ix, size_t p, const uint8_t *codes, size_t n);
int fdb_index_partition_loaded(const fdb_index *ix, size_t p);
int fdb_index_missing_partitions(fdb_index *ix, const float *queries, size_t nq, size_t nprobe, int mode,
                                 uint32_t *out, size_t cap, size_t *n_missing);
/* Device-side constructor straight from a finished build (no host round trip):
 * coarse = nb=1 problem with k=P, pq = nb=D problem with k=C over the residues.
 * order_out (may be NULL) receives [M] the global vector index stored at each
 * partition-major position (the vector_ids order of Partition::new). */
int fdb_index_from_build(fdb_ctx *ctx, const fdb_km *coarse, const fdb_km *pq, fdb_index **out);
int fdb_index_get_layout(fdb_index *ix, uint64_t *offsets /*P+1*/, uint32_t *order /*M or NULL*/,
                         uint8_t *codes /*M*D or NULL*/);
size_t fdb_index_num_vectors(const fdb_index *ix);
void fdb_index_destroy(fdb_index *ix);
/* Database::query for a batch of nq queries (host buffers).  Outputs are [nq][k]
 * in ascending distance; out_count[q] <= k entries are valid.  mode selects the
 * selection semantics (FDB_QUERY_STORED / FDB_QUERY_BUILD).
 * Err(InvalidArgs) when nprobe > P (src/db/stored.rs:403-409). */
int fdb_index_query(fdb_index *ix, const float *queries, size_t nq, size_t k, size_t nprobe,
                    int mode, uint32_t *out_partition, uint32_t *out_vector_index,
                    float *out_sqdist, uint32_t *out_count);
/* the same with queries already resident in HBM and results left there
 * (device pointers; used to time the kernels without host copies) */
int fdb_index_query_device(fdb_index *ix, const float *d_queries, size_t nq, size_t k,
                           size_t nprobe, int mode, uint32_t *d_partition,
                           uint32_t *d_vector_index, float *d_sqdist, uint32_t *d_count);
/* step-wise pieces for parity tests: probe order and one ADC table */
int fdb_index_probe(fdb_index *ix, const float *queries, size_t nq, size_t nprobe, int mode,
                    uint32_t *out_partition /*[nq][nprobe]*/, float *out_sqdist /*[nq][nprobe]*/);
int fdb_index_table(fdb_index *ix, const float *query, uint32_t partition, float *table /*[D][C]*/);
/* Code lists sharded over several GPUs (one process each; no reference analogue, the result is what
 * build::Database::query src/db/build.rs:307-340 returns for the whole database): every rank holds the
 * coarse centroids and codebooks and the lists of the partitions it owns (the others are empty), answers
 * the batch with fdb_index_query_device(..., FDB_QUERY_BUILD, ...), the per-rank lists are all-gathered
 * ([world][nq][k] / [world][nq]) and merged by the canonical key (distance, probe rank, vector index).
 * All pointers are device pointers; both calls are enqueued on fdb_ctx_stream and do not synchronise. */
int fdb_index_probe_device(fdb_index *ix, const float *d_queries, size_t nq, size_t nprobe, int mode,
                           uint32_t *d_partition /*[nq][nprobe]*/);
/* the probe lists the last fdb_index_query_device call used, when it selected them exactly (reference order);
 * FDB_ERR_INVALID_CONTEXT when it did not (probe filter: a set in no particular order) -> fdb_index_probe_device */
int fdb_index_last_probes_device(fdb_index *ix, size_t nq, size_t nprobe, uint32_t *d_partition /*[nq][nprobe]*/);
int fdb_merge_topk_device(fdb_ctx *ctx, int world, size_t nq, size_t k, size_t nprobe, const uint32_t *d_partition,
                          const uint32_t *d_vector_index, const float *d_sqdist, const uint32_t *d_count,
                          const uint32_t *d_probes, uint32_t *d_out_partition, uint32_t *d_out_vector_index,
                          float *d_out_sqdist, uint32_t *d_out_count, uint32_t *d_tie_flag);
/* d_probes may be NULL when the probe order is not at hand (fdb_index_last_probes_device failed): the
 * partition id then stands in for the probe rank and d_tie_flag[q] (device, [nq], zeroed by the caller) is
 * set for the queries in which two candidates of different partitions with exactly equal f32 distances sit
 * at or above the k-th place -- the only case in which the order matters (a few per 10 000 queries on
 * 100M vectors); the caller merges those queries again with fdb_index_probe_device's lists. */
/* per-phase timing is off by default (it adds events and one probe read-back per call) */
int fdb_index_set_timing(fdb_index *ix, int enabled);
/* per-phase device milliseconds of the last fdb_index_query* call (timing enabled):
 * [0] coarse distances, [1] probe selection, [2] localise, [3] ADC tables,
 * [4] code scan + per-partition n-best, [5] merge; and the algorithmic scan bytes */
int fdb_index_last_timing(fdb_index *ix, float ms[6], uint64_t *scan_bytes);
/* how the last fdb_index_query* call was answered: [0] queries decided by the ADC filter
 * path (approximate tables + exact re-check of the candidates inside the error band),
 * [1] queries answered by the exact pipeline (ties, NaN, shapes the filter does not take),
 * [2] candidates the filter path evaluated exactly, [3] code vectors it scanned.
 * Results are identical on both paths; FDB_QUERY_EXACT=1 in the environment forces [1]. */
int fdb_index_last_stats(fdb_index *ix, uint64_t out[4]);
/* which code-scan kernel the last fdb_index_query* call ran (the choice depends on the shape):
 * 0 none (exact pipeline only), 1 fscan_kernel (query-major), 2 pscan_kernel (partition-major, f32 tables),
 * 3 pscan16_kernel (partition-major, 16-bit tables), 4 vscan_kernel (vector-lane, packed 16-bit tables).
 * kernel_ms (may be NULL; timing enabled): device milliseconds of that kernel's launches alone, CUDA events on the
 * launching stream right around them -- the code-scan PHASE of fdb_index_last_timing also holds the helpers around it
 * (grouping of the pairs by partition, table quantisation, per-query merge of the item lists) */
int fdb_index_last_scan_kernel(fdb_index *ix, int *kind, float *kernel_ms);
/* test hook: the filter path's view of the nq queries of the last fdb_index_query_device call
 * (one slice): E[q] = its bound on |approximate - true| distance, the up to 32 smallest approximate
 * distances per query (ascending) with their positions in the concatenation of the probed lists,
 * and the probe lists it scanned (the reference's probe set, in no particular order). */
int fdb_index_debug_band(fdb_index *ix, size_t nq, size_t nprobe, float *E /*[nq]*/,
                         float *cand_approx /*[nq][32]*/, uint32_t *cand_flat /*[nq][32]*/,
                         uint32_t *cand_cnt /*[nq]*/, uint32_t *probes /*[nq][nprobe]*/);

/* ---- multi-GPU: one rank (thread or process) per GPU, NCCL owned by the library --------------------
 * No reference analogue (the reference is single-threaded); this is how the build
 * (DatabaseBuilder::build, src/db/build.rs:73-129) and the query (Database::query,
 * src/db/stored.rs:315-389, src/db/build.rs:294-340) shard over the GPUs of one box.
 * Rank 0 makes an id with fdb_comm_unique_id and hands it to the other ranks (any way it likes);
 * every rank then calls fdb_comm_create with its own context.  All collectives run on the
 * context's stream.  libnccl.so.2 is opened at run time (FDB_NCCL_LIB overrides the name);
 * world == 1 needs no NCCL at all. */
#define FDB_COMM_ID_BYTES 128
int fdb_comm_unique_id(uint8_t *id /*[FDB_COMM_ID_BYTES]*/);
int fdb_comm_create(fdb_ctx *ctx, int world, int rank, const uint8_t *id, fdb_comm **out);
void fdb_comm_destroy(fdb_comm *comm);
int fdb_comm_world(const fdb_comm *comm);
int fdb_comm_rank(const fdb_comm *comm);
uint64_t fdb_comm_collective_count(const fdb_comm *comm); /* collectives enqueued so far */
/* building blocks (device pointers, enqueued on the context's stream): in-place sum, all-gather of bytes */
int fdb_comm_allreduce_device(fdb_comm *comm, float *d_buf, size_t n);
int fdb_comm_allgather_device(fdb_comm *comm, const void *d_send, void *d_recv, size_t bytes_per_rank);
/* max over the ranks of n <= 64 host doubles (device timings); synchronises: also the barrier */
int fdb_comm_max_f64(fdb_comm *comm, double *values, size_t n);
/* cluster_with_events (src/kmeans.rs:104-139) over row shards: rank r holds the rows
 * [n_global*r/world, n_global*(r+1)/world) of the vector set behind km; centroids are replicated.
 *   seed_run_sharded: initialize_centroids (:142-229); first_global[nb] = gen_range(0..n) per problem,
 *     u01[nb][k-1] the draws (identical on every rank); ONE packed all-gather per round (shard total,
 *     local pick, picked row); picked_global[nb][k] (may be NULL) = the chosen global row indices.
 *     Two-stage sampling (shard by total, row inside the shard by weight): the reference's distribution.
 *   run_sharded: the loop of :125-137; ONE all-reduce of [nb*k*dim sums || nb*k counts] per round;
 *     outputs as fdb_kmeans_run.  Assignments are bit-exact per row given the centroids; centroids
 *     differ from one GPU only by the order in which the shards' partial sums are added. */
int fdb_kmeans_seed_run_sharded(fdb_km *km, fdb_comm *comm, size_t n_global, const uint32_t *first_global,
                                const float *u01, uint32_t *picked_global);
int fdb_kmeans_run_sharded(fdb_km *km, fdb_comm *comm, size_t max_rounds, float epsilon, float *gradients,
                           uint32_t *rounds, uint32_t *reassigns);
/* Database::query with the code lists sharded: every rank holds the coarse centroids, the codebooks and
 * the lists of the partitions it owns (the other partitions are empty on it) and is given the SAME batch
 * (device pointer); the per-rank candidates travel in ONE packed all-gather and are merged on every rank.
 * mode FDB_QUERY_BUILD: canonical order (distance, probe rank, vector index) == build::Database::query on
 * the whole database.  mode FDB_QUERY_STORED: the same result whenever no two of the k+1 best candidates
 * have equal distances; the queries where they do (NBestByKey push history decides, src/nbest.rs:52-64)
 * are answered again from the per-partition slot lists of the owning ranks.  Outputs: device pointers,
 * [nq][k] / [nq], identical on every rank.  Synchronises once at the end. */
int fdb_index_query_sharded(fdb_index *ix, fdb_comm *comm, const float *d_queries, size_t nq, size_t k,
                            size_t nprobe, int mode, uint32_t *d_partition, uint32_t *d_vector_index,
                            float *d_sqdist, uint32_t *d_count);

/* how many queries of the last fdb_index_query_sharded call were merged a second time (distance ties) */
int fdb_index_last_sharded_ties(fdb_index *ix, uint32_t *ties);

/* Host buffers.  The query / upload / download calls take any host pointer.  Page-locked memory (registered
 * here, or allocated pinned by the caller) is copied by asynchronous DMA, so fdb_index_query overlaps the copy of
 * a batch with answering it; pageable memory (a plain Vec<f32>) is staged slice by slice -- by a few host threads
 * into a page-locked ring of the index, then by DMA -- while the GPU answers the previous slice.  Registering costs about as much as one copy: worth it for buffers that are
 * reused.  No reference analogue (the reference never leaves host memory). */
int fdb_host_register(fdb_ctx *ctx, void *p, size_t bytes);
int fdb_host_unregister(fdb_ctx *ctx, void *p);

/* raw device buffers for benches that keep inputs resident */
int fdb_device_alloc(fdb_ctx *ctx, size_t bytes, void **out);
int fdb_device_free(fdb_ctx *ctx, void *p);
int fdb_device_upload(fdb_ctx *ctx, void *d_dst, const void *src, size_t bytes);     /* synchronous */
int fdb_device_download(fdb_ctx *ctx, void *dst, const void *d_src, size_t bytes);   /* synchronous */
int fdb_device_fill_uniform(fdb_ctx *ctx, float *d, size_t count, uint64_t seed, uint64_t start);
int fdb_device_flush_l2(fdb_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
