/*
 * flechas_oracle.h -- CPU restatement of flechasdb's IVF-PQ build/query arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under flechasdb_b200/ may include, link or
 * call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / the CPU arm.
 *
 * Every function cites the reference lines (paths relative to /root/reference)
 * whose order of floating-point operations it follows.  Compile with
 * -ffp-contract=off and without -ffast-math (see oracle/Makefile): the reference
 * is Rust, which never contracts a*b+c into an FMA and never reassociates.
 *
 * Pinning: primitives are pinned by the reference's own unit-test vectors
 * (src/linalg.rs:366-869, src/distribution.rs:125-374, src/vector.rs:177-267;
 * transcribed in tests/test_oracle_kats.py).  The reference has NO tests for
 * kmeans / partitions / nbest / db::build / db::stored and cannot be compiled
 * here (no rustc/cargo): for those functions PARITY IS UNPINNED by the
 * reference and rests on this line-by-line restatement.
 */
#ifndef FLECHAS_ORACLE_H
#define FLECHAS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* error codes (mirror the reference's Err / panic classes) */
#define FO_OK 0
#define FO_ERR_INVALID_ARGS (-1)  /* Error::InvalidArgs                        */
#define FO_ERR_PANIC_EMPTY_CLUSTER (-4) /* assert_ne!(count,0) src/kmeans.rs:259 */
#define FO_ERR_PANIC_WEIGHTS (-5) /* WeightedIndex unwrap() src/kmeans.rs:199,207,216 */
#define FO_ERR_PANIC_NAN (-6)     /* min_index.unwrap() / partial_cmp().unwrap()  */

/* A VectorSet (src/vector.rs:11-25): BlockVectorSet when off=0,dim=stride;
 * SubVectorSet (src/vector.rs:103-149) otherwise. */
typedef struct {
    float *base;   /* first row of the underlying block vector set */
    size_t n;      /* number of vectors                             */
    size_t stride; /* floats between consecutive rows               */
    size_t off;    /* column offset of the sub-vector               */
    size_t dim;    /* vector_size()                                 */
} fo_view;

/* ---- src/linalg.rs ---------------------------------------------------- */
float fo_dot(const float *x, const float *y, size_t n);        /* :12-40  */
float fo_dot_naive(const float *x, const float *y, size_t n);  /* :43-53  */
float fo_norm2(const float *x, size_t n);                      /* :61-75  */
float fo_sum(const float *x, size_t n);                        /* :208-235 */
int fo_min(const float *x, size_t n, float *out);              /* :252-283, returns 0 if None */
int fo_max_abs(const float *x, size_t n, float *out);          /* :306-345, returns 0 if None */
void fo_add_in(float *l, const float *r, size_t n);            /* :149-155 */
void fo_subtract(const float *l, const float *r, float *o, size_t n); /* :158-165 */
void fo_subtract_in(float *l, const float *r, size_t n);       /* :168-174 */
void fo_scale_in(float *x, float a, size_t n);                 /* :188-193 */
/* squared distance exactly as the hot loops compute it: subtract then dot(d,d) */
float fo_sqdist(const float *v, const float *c, size_t n, float *buf);

/* ---- src/vector.rs ----------------------------------------------------- */
int fo_chunk_check(size_t data_len, size_t vector_size);       /* :40-57  */
int fo_divide(const fo_view *vs, size_t d, fo_view *out);      /* :154-174 */

/* ---- src/distribution.rs ----------------------------------------------- */
typedef struct {
    float *weights;
    size_t n;
    float total;
    float scale; /* rand 0.8.5 UniformFloat<f32>::new(0,total).scale */
} fo_wi;
int fo_wi_new(fo_wi *wi, const float *weights, size_t n);      /* :35-54  */
void fo_wi_free(fo_wi *wi);
int fo_wi_update(fo_wi *wi, const size_t *idx, const float *w, size_t m); /* :63-91 */
float fo_wi_get_weight(const fo_wi *wi, size_t i);             /* :94-96  */
/* the cumulative scan of sample() for an already drawn sample value :108-120 */
size_t fo_wi_pick(const fo_wi *wi, float sample);
/* rand 0.8.5 UniformFloat<f32>::sample with the 23-bit draw u=(bits>>9)*2^-23 */
float fo_wi_sample_value(const fo_wi *wi, float u01);

/* ---- src/kmeans.rs ------------------------------------------------------ */
/* k-means++ (:142-229).  RNG injection (the reference's thread_rng is unseeded):
 *   first      = the index gen_range(0..n) returned (:172)
 *   chosen     = if non-NULL, k-1 indices to use instead of sampling (entries 1..k)
 *   u01        = otherwise k-1 uniform draws in [0,1) fed to the sampler (:202)
 * Outputs: centroids k*dim, indices n, optional weights n (final D^2 weights),
 *          optional picked k (the indices that became centres). */
int fo_kmeans_init(const fo_view *vs, size_t k, size_t first,
                   const uint32_t *chosen, const float *u01,
                   float *centroids, uint32_t *indices,
                   float *weights_out, uint32_t *picked_out);
/* update_centroids (:232-276): returns gradient through *gradient */
int fo_kmeans_update(const fo_view *vs, size_t k, float *centroids,
                     const uint32_t *indices, float *gradient);
/* reassign_centroids (:279-306); nthreads>1 splits rows (results identical) */
int fo_kmeans_reassign(const fo_view *vs, size_t k, const float *centroids,
                       uint32_t *indices, int nthreads);
/* cluster_with_events (:104-139) after initialisation; records the event stream:
 * gradients[r] for every FinishedCentroidUpdate(r), *rounds = number of updates,
 * *reassigns = number of reassignments.  max_rounds=100 in the reference (:114). */
int fo_kmeans_lloyd(const fo_view *vs, size_t k, float *centroids,
                    uint32_t *indices, size_t max_rounds, float epsilon,
                    float *gradients, size_t *rounds, size_t *reassigns,
                    int nthreads);

/* ---- src/partitions.rs:128-138 ------------------------------------------ */
void fo_residues(fo_view *vs, size_t p, const float *centroids,
                 const uint32_t *indices);

/* ---- src/db/build.rs:446-482 -------------------------------------------- */
/* Partition-major code layout: offsets P+1, order M (global index of the
 * vector at each partition-major position), codes M*D u32 partition-major. */
void fo_extract_partitions(size_t M, size_t P, size_t D,
                           const uint32_t *part_idx, const uint32_t *codes_div_major,
                           uint64_t *offsets, uint32_t *order, uint32_t *codes_pm);

/* ---- src/nbest.rs -------------------------------------------------------- */
typedef struct { float key; uint32_t a; uint32_t b; } fo_item;
typedef struct { fo_item *items; size_t n; size_t len; } fo_nbest;
void fo_nbest_push(fo_nbest *nb, fo_item cand);                 /* :52-64 */

/* ---- query --------------------------------------------------------------- */
typedef struct {
    size_t N, P, D, C;
    const float *coarse;       /* P*N            */
    const float *codebooks;    /* D*C*(N/D)      */
    const uint64_t *offsets;   /* P+1            */
    const uint32_t *codes_pm;  /* M*D, partition-major, ascending global id */
} fo_index;
/* stored::Database::query (src/db/stored.rs:331-442,549-597): NBestByKey selection.
 * mode 0 = stored (NBestByKey), mode 1 = build (src/db/build.rs:307-382,521-565:
 * full stable sorts).  Outputs k entries per query (count in out_count). */
int fo_query(const fo_index *ix, const float *q, size_t nq, size_t k, size_t nprobe,
             int mode, uint32_t *out_part, uint32_t *out_vidx, float *out_dist,
             uint32_t *out_count, int nthreads);
/* pieces, exposed for step-wise parity */
int fo_query_probe(const fo_index *ix, const float *q, size_t nprobe, int mode,
                   uint32_t *probe_part, float *probe_dist);
void fo_query_table(const fo_index *ix, const float *q, uint32_t part, float *table);

/* ---- synthetic data -------------------------------- */
void fo_fill_uniform(float *out, size_t count, uint64_t seed, uint64_t start);
uint64_t fo_splitmix64(uint64_t seed, uint64_t i);

#ifdef __cplusplus
}
#endif
#endif
