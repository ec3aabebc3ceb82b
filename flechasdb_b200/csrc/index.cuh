// Device state of a queryable IVF-PQ index (stored::Database, src/db/stored.rs:41-57).
#pragma once
#include "kmeans.cuh"

#include <vector>

namespace fdb { struct FilterState; }
#ifndef FDB_FILTER_SLOTS
#define FDB_FILTER_SLOTS 8   // scratch slots (upper bound) = compute streams that slices of a host batch rotate through
#endif
// how many of them a host batch uses (FDB_QUERY_STREAMS in the environment, read once)
int fdb_filter_slots_in_use();

struct fdb_index {
    fdb_ctx *ctx = nullptr;
    size_t N = 0, P = 0, D = 0, C = 0, s = 0, M = 0;
    fdb::DevBuf<float> coarse, codebooks;
    fdb::DevBuf<uint8_t> codes;
    fdb::DevBuf<uint32_t> part_off;    // [P+1] vectors before partition p
    fdb::DevBuf<uint64_t> part_cstart; // [P] byte offset of the partition's code list
    fdb::DevBuf<uint32_t> order;       // [M] global vector index at each partition-major position
    std::vector<uint32_t> h_off;
    std::vector<uint64_t> h_cstart;
    // lazily loaded partitions (fdb_index_create_lazy / fdb_index_set_partition): the code lists arrive one by one
    // and are appended to `codes`; the offsets are rebuilt before the next query
    bool lazy = false, layout_dirty = false;
    std::vector<uint8_t> loaded;       // [P]
    std::vector<uint32_t> sizes;       // [P] vectors per partition (0 while not loaded)
    size_t codes_used = 0;             // bytes of `codes` in use
    // scratch
    fdb::DevBuf<float> q_dev, dist, loc, tables, part_d, out_d, probe_d;
    fdb::DevBuf<uint32_t> probes, part_v, part_cnt, out_p, out_v, out_c;
    std::vector<cudaEvent_t> events;
    cudaStream_t copy_stream = nullptr;   // host->device copies of fdb_index_query, overlapped with the kernels
    std::vector<cudaEvent_t> copy_events;
    float phase_ms[6] = {0, 0, 0, 0, 0, 0};
    uint64_t scan_bytes = 0;
    size_t last_npairs = 0;
    bool timing = false;
    size_t chunk_pairs = 8192;
    fdb::DevBuf<float> fb_q, fb_d;        // queries the filter path handed to the exact pipeline
    fdb::DevBuf<uint32_t> fb_p, fb_v, fb_c, fb_probes;
    bool last_filter = false;
    int last_scan_kind = 0;   // fdb_index_last_scan_kernel
    // stored-semantic probes from sparse rows: the queries with tied distances, their rows and full distance rows
    fdb::DevBuf<uint32_t> tied_list;
    fdb::DevBuf<float> tied_q, tied_dist;
    unsigned last_probe_ties = 0;
    float *h_qstage = nullptr;     // page-locked ring (two slots) pageable query batches are staged through
    size_t h_qstage_floats = 0;
    uint32_t *h_stage = nullptr;   // page-locked staging for the handed-back rows of a host batch
    size_t h_stage_words = 0;
    // timing mode: events around the code-scan kernel itself (its launches of the last call, summed)
    std::vector<cudaEvent_t> kev;
    size_t kev_used = 0;
    float scan_kernel_ms = 0.f;
    int kev_mark() {
        if (!timing) return FDB_OK;
        if (kev_used == kev.size()) {
            cudaEvent_t e;
            FDB_CUDA(cudaEventCreate(&e));
            kev.push_back(e);
        }
        FDB_CUDA(cudaEventRecord(kev[kev_used++], ctx->stream));
        return FDB_OK;
    }
    fdb::FilterState *filter = nullptr;   // ADC filter path (adc_filter.cu), null when not usable
    bool last_probes_exact = false;       // `probes` holds the last device batch's lists in the reference's order
    size_t last_probes_nq = 0, last_probes_nprobe = 0;
    fdb::DevBuf<uint32_t> sh_flags, sh_list;   // sharded query: tie flags (+ a counter), list of the flagged queries
    uint32_t last_sharded_ties = 0;
    uint64_t last_stats[4] = {0, 0, 0, 0};  // queries on the filter path, exact fallbacks, exact candidates, scanned vectors
    ~fdb_index();                         // adc_filter.cu (FilterState is complete there)
};

namespace fdb {

// CUDA-event log of the query phases (0 coarse, 1 probe, 2 localise / pair constants,
// 3 ADC tables, 4 code scan, 5 selection / merge); a phase ends where the next mark starts
struct EventLog {
    fdb_index *ix;
    size_t used = 0;
    std::vector<std::pair<int, size_t>> marks;  // (phase, index of the start event)
    int mark(int phase) {
        if (!ix->timing) return FDB_OK;
        if (used == ix->events.size()) {
            cudaEvent_t e;
            FDB_CUDA(cudaEventCreate(&e));
            ix->events.push_back(e);
        }
        FDB_CUDA(cudaEventRecord(ix->events[used], ix->ctx->stream));
        marks.emplace_back(phase, used);
        used++;
        return FDB_OK;
    }
    int finish() {
        for (int i = 0; i < 6; ++i) ix->phase_ms[i] = 0.f;
        ix->scan_kernel_ms = 0.f;
        if (!ix->timing) return FDB_OK;
        for (size_t i = 0; i + 1 < ix->kev_used; i += 2) {
            float ms = 0.f;
            FDB_CUDA(cudaEventElapsedTime(&ms, ix->kev[i], ix->kev[i + 1]));
            ix->scan_kernel_ms += ms;
        }
        for (size_t i = 0; i + 1 < marks.size(); ++i) {
            if (marks[i].first < 0) continue;
            float ms = 0.f;
            FDB_CUDA(cudaEventElapsedTime(&ms, ix->events[marks[i].second], ix->events[marks[i + 1].second]));
            ix->phase_ms[marks[i].first] += ms;
        }
        return FDB_OK;
    }
};

// ---- ADC filter path (adc_filter.cu) ---------------------------------------------------
// filter_prepare: per-index tables and bounds, once the coarse centroids and codebooks are
// resident.  filter_query: decides every query it can from approximate tables + an exact
// re-check of the few candidates inside the error band; the others are appended to
// *d_fb_list (count in *h_nfb) for the exact pipeline.
int filter_prepare(fdb_index *ix);
void filter_free(fdb_index *ix);
bool filter_eligible(const fdb_index *ix, size_t nq, size_t k, size_t nprobe);
// filter_probe: probe lists from approximate coarse scores (tensor pipe) + exact re-check;
// *done = false when the shape is not taken (the caller then runs the exact probe kernels).
// *h_nhard = queries among the *h_nfb handed back whose probe list is not the reference's.
int filter_probe(fdb_index *ix, const float *d_q, size_t nq, size_t nprobe, EventLog *log, bool *done);
// nprobe beyond the probe filter, build semantic: dense rows of coarse distances for the exact selection kernel,
// exact for the partitions the tensor-pipe scores cannot rule out and +inf elsewhere
int filter_probe_dense(fdb_index *ix, const float *d_q, size_t nq, size_t nprobe, float *d_dist, bool *done);
// A batch is answered slice by slice (filter_probe + filter_query per slice, nothing waits on the
// host); the queries a slice could not decide are appended, with their probe lists, to a batch
// list that filter_batch_end hands to the exact pipeline.  *h_nhard = those among the *h_nfb
// whose probe list is not the reference's.
int filter_batch_begin(fdb_index *ix, size_t nq_total, size_t nprobe);
// two scratch slots: consecutive slices of a host batch run on two streams when
// filter_can_overlap (fork after batch_begin, join before batch_end)
bool filter_can_overlap(const fdb_index *ix, size_t nprobe);
int filter_use_slot(fdb_index *ix, int slot, bool own_stream, cudaStream_t *stream);
int filter_fork(fdb_index *ix);
int filter_slot_wait_fork(fdb_index *ix, int slot);
int filter_slot_done(fdb_index *ix, int slot);
int filter_join(fdb_index *ix);
int filter_debug_band(fdb_index *ix, size_t nq, size_t nprobe, float *E, float *cand_approx, uint32_t *cand_flat,
                      uint32_t *cand_cnt, uint32_t *probes);
int filter_query(fdb_index *ix, const float *d_q, size_t q_base, size_t nq, size_t k, size_t nprobe, uint32_t *d_p,
                 uint32_t *d_v, float *d_d, uint32_t *d_c, EventLog *log);
int filter_batch_end(fdb_index *ix, size_t nq_total, const uint32_t **d_fb_q, const uint32_t **d_fb_probes,
                     unsigned *h_nfb, unsigned *h_nhard);

}  // namespace fdb
