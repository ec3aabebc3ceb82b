// C++17 mirror of the reference's host-side API for the IVF-PQ path, on top of the C ABI
// (include/flechasdb_b200.h).  Header only.  The reference is Rust; Rust is not available in
// this build environment, so this is the compiled host side (rust_shim/ holds the Rust bodies
// a maintainer drops into the reference crate; they make the same calls in the same order).
//
// Mirrors, with the reference's names and semantics:
//   flechasdb::vector::BlockVectorSet                      src/vector.rs:28-100
//   flechasdb::kmeans::{Codebook, ClusterEvent, cluster_with_events}   src/kmeans.rs:62-139
//   flechasdb::partitions::Partitioning                    src/partitions.rs:96-144
//   flechasdb::db::build::{DatabaseBuilder, Database, BuildEvent, QueryEvent, QueryResult}
//                                                          src/db/build.rs:23-587
//   flechasdb::error::Error                                src/error.rs:5-18
#pragma once

#include <array>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/flechasdb_b200.h"

namespace flechasdb {

// ---- error.rs ------------------------------------------------------------------------------
struct Error : std::runtime_error {
    enum Kind { InvalidArgs, InvalidData, InvalidContext, Panic, Cuda, Unsupported } kind;
    Error(Kind k, const std::string &m) : std::runtime_error(m), kind(k) {}
};

inline void check(int rc) {
    if (rc == FDB_OK) return;
    const std::string msg = fdb_last_error();
    switch (rc) {
        case FDB_ERR_INVALID_ARGS: throw Error(Error::InvalidArgs, msg);
        case FDB_ERR_INVALID_DATA: throw Error(Error::InvalidData, msg);
        case FDB_ERR_INVALID_CONTEXT: throw Error(Error::InvalidContext, msg);
        case FDB_ERR_EMPTY_CLUSTER:   // assert_ne!(count, 0)            -> panic in the reference
        case FDB_ERR_WEIGHTS:         // WeightedIndex ... .unwrap()      -> panic
        case FDB_ERR_NAN:             // unwrap() on None / partial_cmp   -> panic
            throw Error(Error::Panic, msg);
        case FDB_ERR_UNSUPPORTED: throw Error(Error::Unsupported, msg);
        default: throw Error(Error::Cuda, msg);
    }
}

struct Context {
    fdb_ctx *h = nullptr;
    explicit Context(int device = 0) { check(fdb_ctx_create(device, &h)); }
    ~Context() { fdb_ctx_destroy(h); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
};

// ---- vector.rs -----------------------------------------------------------------------------
// BlockVectorSet<f32>: row-major rows; `chunk` fails when the data is not a multiple of the size.
class BlockVectorSet {
  public:
    static BlockVectorSet chunk(std::shared_ptr<Context> ctx, const std::vector<float> &data, size_t vector_size) {
        if (vector_size == 0) throw Error(Error::InvalidArgs, "vector size must be non-zero");
        if (!data.empty() && data.size() % vector_size != 0)
            throw Error(Error::InvalidArgs, "data size (" + std::to_string(data.size()) +
                                                ") is not a multiple of vector size (" +
                                                std::to_string(vector_size) + ")");
        BlockVectorSet vs;
        vs.ctx_ = ctx;
        check(fdb_vs_upload(ctx->h, data.data(), data.size() / vector_size, vector_size, &vs.h_));
        return vs;
    }
    BlockVectorSet(BlockVectorSet &&o) noexcept : ctx_(std::move(o.ctx_)), h_(o.h_) { o.h_ = nullptr; }
    BlockVectorSet &operator=(BlockVectorSet &&o) noexcept {
        std::swap(ctx_, o.ctx_);
        std::swap(h_, o.h_);
        return *this;
    }
    ~BlockVectorSet() { fdb_vs_destroy(h_); }
    size_t len() const { return fdb_vs_len(h_); }
    size_t vector_size() const { return fdb_vs_vector_size(h_); }
    std::vector<float> get(size_t i) const {
        std::vector<float> v(vector_size());
        check(fdb_vs_download_rows(h_, i, 1, v.data()));
        return v;
    }
    fdb_vs *handle() const { return h_; }
    const std::shared_ptr<Context> &context() const { return ctx_; }

  private:
    BlockVectorSet() = default;
    std::shared_ptr<Context> ctx_;
    fdb_vs *h_ = nullptr;
};

// ---- kmeans.rs -----------------------------------------------------------------------------
struct ClusterEvent {
    enum Kind {
        StartingCentroidInitialization, FinishedCentroidInitialization, StartingCentroidUpdate,
        FinishedCentroidUpdate, StartingCentroidReassignment, FinishedCentroidReassignment
    } kind;
    size_t round = 0;
    float gradient = 0.0f;  // FinishedCentroidUpdate only
};

struct Codebook {
    std::vector<float> centroids;  // [k][vector_size]
    size_t vector_size = 0;
    std::vector<uint32_t> indices; // [n]
};

// The draws the reference takes from rand::thread_rng() (src/kmeans.rs:148,172,202).
struct SeedSource {
    std::mt19937_64 rng;
    explicit SeedSource(uint64_t seed = std::random_device{}()) : rng(seed) {}
    uint32_t first(size_t n) { return (uint32_t)(rng() % n); }          // gen_range(0..n)
    float draw() { return (float)(rng() >> 41) * 1.1920929e-07f; }      // (u32 >> 9) * 2^-23
};

using ClusterEventHandler = std::function<void(const ClusterEvent &)>;

// replays the event sequence of cluster_with_events for one finished problem
inline void replay_cluster_events(const ClusterEventHandler &ev, const float *grads, size_t rounds, size_t reassigns) {
    ev({ClusterEvent::StartingCentroidInitialization});
    ev({ClusterEvent::FinishedCentroidInitialization});
    for (size_t r = 0; r < rounds; ++r) {
        ev({ClusterEvent::StartingCentroidUpdate, r});
        ev({ClusterEvent::FinishedCentroidUpdate, r, grads[r]});
        if (r < reassigns) {
            ev({ClusterEvent::StartingCentroidReassignment, r});
            ev({ClusterEvent::FinishedCentroidReassignment, r});
        }
    }
}

// nb side-by-side problems on strided sub-vector views of `vs` (nb = 1: cluster_with_events)
struct KMeansRun {
    fdb_km *km = nullptr;
    std::vector<float> gradients;          // [nb][FDB_KMEANS_MAX_ROUNDS]
    std::vector<uint32_t> rounds, reassigns;
    ~KMeansRun() { fdb_kmeans_destroy(km); }
};

inline std::unique_ptr<KMeansRun> cluster_device(const BlockVectorSet &vs, size_t col_off, size_t dim, size_t nb,
                                                 size_t k, SeedSource &seeds) {
    auto run = std::make_unique<KMeansRun>();
    check(fdb_kmeans_begin(vs.handle(), col_off, dim, nb, k, &run->km));  // Err(InvalidArgs) if n < k
    std::vector<uint32_t> first(nb);
    std::vector<float> u(nb * (k > 0 ? k - 1 : 0));
    for (size_t b = 0; b < nb; ++b) {
        first[b] = seeds.first(vs.len());
        for (size_t i = 0; i + 1 < k; ++i) u[b * (k - 1) + i] = seeds.draw();
    }
    check(fdb_kmeans_seed_run(run->km, first.data(), u.data(), /*exact=*/0, nullptr));
    run->gradients.resize(nb * FDB_KMEANS_MAX_ROUNDS);
    run->rounds.resize(nb);
    run->reassigns.resize(nb);
    check(fdb_kmeans_run(run->km, FDB_KMEANS_MAX_ROUNDS, FDB_KMEANS_EPSILON, run->gradients.data(),
                         run->rounds.data(), run->reassigns.data()));
    return run;
}

// The same loop driven step by step from the host (fdb_kmeans_update / fdb_kmeans_reassign per round), so that
// every ClusterEvent fires WHEN its phase starts / ends, like the reference's (src/kmeans.rs:121-137): callers that
// time the phases with Instant::now() between events (src/main.rs:53-94) see real durations.  One host
// synchronisation per phase; cluster_device above runs the whole loop on the device and replays the events.
inline std::unique_ptr<KMeansRun> cluster_device_live(const BlockVectorSet &vs, size_t col_off, size_t dim, size_t k,
                                                      SeedSource &seeds, const ClusterEventHandler &ev) {
    auto run = std::make_unique<KMeansRun>();
    check(fdb_kmeans_begin(vs.handle(), col_off, dim, 1, k, &run->km));
    ev({ClusterEvent::StartingCentroidInitialization});
    const uint32_t first = seeds.first(vs.len());
    std::vector<float> u(k > 0 ? k - 1 : 0);
    for (float &x : u) x = seeds.draw();
    check(fdb_kmeans_seed_run(run->km, &first, u.data(), /*exact=*/0, nullptr));
    ev({ClusterEvent::FinishedCentroidInitialization});
    run->gradients.assign(FDB_KMEANS_MAX_ROUNDS, 0.0f);
    run->rounds.assign(1, 0);
    run->reassigns.assign(1, 0);
    for (size_t r = 0; r < FDB_KMEANS_MAX_ROUNDS; ++r) {
        ev({ClusterEvent::StartingCentroidUpdate, r});
        float g = 0.0f;
        check(fdb_kmeans_update(run->km, nullptr, &g));
        run->gradients[r] = g;
        run->rounds[0] = (uint32_t)(r + 1);
        ev({ClusterEvent::FinishedCentroidUpdate, r, g});
        if (g < FDB_KMEANS_EPSILON) break;
        ev({ClusterEvent::StartingCentroidReassignment, r});
        check(fdb_kmeans_reassign(run->km, nullptr));
        run->reassigns[0] = (uint32_t)(r + 1);
        ev({ClusterEvent::FinishedCentroidReassignment, r});
    }
    return run;
}

inline Codebook cluster_with_events(const BlockVectorSet &vs, size_t k, SeedSource &seeds,
                                    const ClusterEventHandler &ev = [](const ClusterEvent &) {}) {
    if (k == 0) throw Error(Error::InvalidArgs, "k must be non-zero");
    auto run = cluster_device(vs, 0, vs.vector_size(), 1, k, seeds);
    replay_cluster_events(ev, run->gradients.data(), run->rounds[0], run->reassigns[0]);
    Codebook cb;
    cb.vector_size = vs.vector_size();
    cb.centroids.resize(k * cb.vector_size);
    cb.indices.resize(vs.len());
    check(fdb_kmeans_get(run->km, cb.centroids.data(), cb.indices.data()));
    return cb;
}

// ---- db/build.rs -----------------------------------------------------------------------------
struct BuildEvent {
    enum Kind {
        StartingIdAssignment, FinishedIdAssignment, StartingPartitioning, FinishedPartitioning,
        StartingSubvectorDivision, FinishedSubvectorDivision, StartingQuantization, FinishedQuantization,
        Cluster
    } kind;
    size_t division = 0;
    ClusterEvent cluster{};
};

struct QueryEvent {
    enum Kind {
        StartingPartitionSelection, FinishedPartitionSelection, StartingPartitionQuery,
        FinishedPartitionQuery, StartingResultSelection, FinishedResultSelection,
        StartingQueryInitialization, FinishedQueryInitialization   // stored::Database only (src/db/stored.rs:342-360)
    } kind;
    size_t partition_index = 0;
};

using Uuid = std::array<uint8_t, 16>;

struct QueryResult {
    size_t partition_index;
    Uuid vector_id;
    size_t vector_index;
    float squared_distance;
};

// AttributeValue { String, Uint64 } (src/db/mod.rs), Attributes = name -> value, AttributeTable = vector id -> Attributes.
// Attributes live on the host only.
struct AttributeValue {
    bool is_string = false;
    std::string string_value;
    uint64_t uint64_value = 0;
    AttributeValue() = default;
    AttributeValue(const std::string &s) : is_string(true), string_value(s) {}
    AttributeValue(const char *s) : is_string(true), string_value(s) {}
    AttributeValue(uint64_t v) : uint64_value(v) {}
    bool operator==(const AttributeValue &o) const {
        return is_string == o.is_string && string_value == o.string_value && uint64_value == o.uint64_value;
    }
};
using Attributes = std::map<std::string, AttributeValue>;
using AttributeTable = std::map<Uuid, Attributes>;

class Database {
  public:
    size_t num_vectors() const { return vector_ids_.size(); }
    size_t vector_size() const { return vector_size_; }
    size_t num_partitions() const { return num_partitions_; }
    size_t num_divisions() const { return num_divisions_; }
    size_t subvector_size() const { return vector_size_ / num_divisions_; }
    size_t num_clusters() const { return num_clusters_; }
    const std::vector<Uuid> &vector_ids() const { return vector_ids_; }

    // build::Database::query_with_events (src/db/build.rs:307-340)
    std::vector<QueryResult> query(const std::vector<float> &v, size_t k, size_t nprobe,
                                   const std::function<void(const QueryEvent &)> &ev = [](const QueryEvent &) {},
                                   int mode = FDB_QUERY_BUILD) const {
        if (k == 0 || nprobe == 0) throw Error(Error::InvalidArgs, "NonZeroUsize");
        ev({QueryEvent::StartingPartitionSelection});
        std::vector<uint32_t> probes(nprobe);
        check(fdb_index_probe(index_, v.data(), 1, nprobe, mode, probes.data(), nullptr));
        ev({QueryEvent::FinishedPartitionSelection});
        std::vector<uint32_t> part(k), vidx(k);
        std::vector<float> dist(k);
        uint32_t count = 0;
        check(fdb_index_query(index_, v.data(), 1, k, nprobe, mode, part.data(), vidx.data(), dist.data(), &count));
        for (uint32_t p : probes) {
            ev({QueryEvent::StartingPartitionQuery, p});
            ev({QueryEvent::FinishedPartitionQuery, p});
        }
        ev({QueryEvent::StartingResultSelection});
        std::vector<QueryResult> out;
        for (uint32_t i = 0; i < count; ++i)
            out.push_back({part[i], vector_ids_[order_[offsets_[part[i]] + vidx[i]]], vidx[i], dist[i]});
        ev({QueryEvent::FinishedResultSelection});
        return out;
    }
    // batched form: outputs [nq][k]
    void query_batch(const float *queries, size_t nq, size_t k, size_t nprobe, int mode, uint32_t *part,
                     uint32_t *vidx, float *dist, uint32_t *count) const {
        check(fdb_index_query(index_, queries, nq, k, nprobe, mode, part, vidx, dist, count));
    }
    // Database::get_attribute (src/db/build.rs:228-245): null when the vector has no such attribute; InvalidArgs when
    // no attribute was ever set for this id (the reference looks the id up in its attribute table only)
    const AttributeValue *get_attribute(const Uuid &id, const std::string &key) const {
        auto it = attribute_table_.find(id);
        if (it == attribute_table_.end()) throw Error(Error::InvalidArgs, "no such vector ID");
        auto a = it->second.find(key);
        return a == it->second.end() ? nullptr : &a->second;
    }
    // Database::set_attribute_at (src/db/build.rs:252-285): replaces an existing value; InvalidArgs when i is out of bounds
    void set_attribute_at(size_t i, const std::string &key, const AttributeValue &value) {
        if (i >= vector_ids_.size()) throw Error(Error::InvalidArgs, "vector index out of bounds: " + std::to_string(i));
        attribute_table_[vector_ids_[i]][key] = value;
    }
    const AttributeTable &attribute_table() const { return attribute_table_; }
    fdb_index *index() const { return index_; }
    // partition centroids [P][N] and codebooks [D][C][N/D] (what serialize_database writes)
    void quantisers(float *coarse, float *codebooks) const {
        check(fdb_kmeans_get(coarse_->km, coarse, nullptr));
        check(fdb_kmeans_get(pq_->km, codebooks, nullptr));
    }
    ~Database() {
        fdb_index_destroy(index_);
        pq_.reset();
        coarse_.reset();
    }

  private:
    friend class DatabaseBuilder;
    Database() = default;
    size_t vector_size_ = 0, num_partitions_ = 0, num_divisions_ = 0, num_clusters_ = 0;
    std::vector<Uuid> vector_ids_;
    AttributeTable attribute_table_;
    std::vector<uint64_t> offsets_;
    std::vector<uint32_t> order_;
    std::unique_ptr<KMeansRun> coarse_, pq_;
    std::unique_ptr<BlockVectorSet> residues_;
    fdb_index *index_ = nullptr;
};

class DatabaseBuilder {
  public:
    explicit DatabaseBuilder(BlockVectorSet vs) : vs_(std::move(vs)) {}  // consumes vs (src/db/build.rs:44-52)
    DatabaseBuilder &with_partitions(size_t p) { num_partitions_ = nonzero(p); return *this; }
    DatabaseBuilder &with_divisions(size_t d) { num_divisions_ = nonzero(d); return *this; }
    DatabaseBuilder &with_clusters(size_t c) { num_clusters_ = nonzero(c); return *this; }
    DatabaseBuilder &with_seed(uint64_t s) { seeds_ = SeedSource(s); return *this; }

    std::unique_ptr<Database> build() { return build_with_events([](const BuildEvent &) {}); }

    // src/db/build.rs:78-129
    std::unique_ptr<Database> build_with_events(const std::function<void(const BuildEvent &)> &event) {
        const size_t M = vs_.len(), N = vs_.vector_size();
        std::unique_ptr<Database> db(new Database);
        event({BuildEvent::StartingIdAssignment});
        db->vector_ids_.resize(M);
        for (auto &id : db->vector_ids_) {  // Uuid::new_v4()
            uint64_t a = seeds_.rng(), b = seeds_.rng();
            for (int i = 0; i < 8; ++i) { id[i] = (uint8_t)(a >> (8 * i)); id[8 + i] = (uint8_t)(b >> (8 * i)); }
            id[6] = (id[6] & 0x0F) | 0x40;
            id[8] = (id[8] & 0x3F) | 0x80;
        }
        event({BuildEvent::FinishedIdAssignment});
        auto cluster_ev = [&](const ClusterEvent &e) { event({BuildEvent::Cluster, 0, e}); };
        event({BuildEvent::StartingPartitioning});
        db->coarse_ = cluster_device(vs_, 0, N, 1, num_partitions_, seeds_);      // partition_with_events
        replay_cluster_events(cluster_ev, db->coarse_->gradients.data(), db->coarse_->rounds[0],
                              db->coarse_->reassigns[0]);
        check(fdb_vs_subtract_assigned(vs_.handle(), db->coarse_->km));           // residues in place
        event({BuildEvent::FinishedPartitioning});
        event({BuildEvent::StartingSubvectorDivision});
        if (N % num_divisions_ != 0)                                              // divide_vector_set
            throw Error(Error::InvalidArgs, "vector size (" + std::to_string(N) + ") is not divisible by " +
                                                std::to_string(num_divisions_));
        event({BuildEvent::FinishedSubvectorDivision});
        db->pq_ = cluster_device(vs_, 0, N / num_divisions_, num_divisions_, num_clusters_, seeds_);
        for (size_t di = 0; di < num_divisions_; ++di) {
            event({BuildEvent::StartingQuantization, di});
            replay_cluster_events(cluster_ev, db->pq_->gradients.data() + di * FDB_KMEANS_MAX_ROUNDS,
                                  db->pq_->rounds[di], db->pq_->reassigns[di]);
            event({BuildEvent::FinishedQuantization, di});
        }
        check(fdb_index_from_build(vs_.context()->h, db->coarse_->km, db->pq_->km, &db->index_));
        db->offsets_.resize(num_partitions_ + 1);
        db->order_.resize(M);
        check(fdb_index_get_layout(db->index_, db->offsets_.data(), db->order_.data(), nullptr));
        db->vector_size_ = N;
        db->num_partitions_ = num_partitions_;
        db->num_divisions_ = num_divisions_;
        db->num_clusters_ = num_clusters_;
        db->residues_.reset(new BlockVectorSet(std::move(vs_)));
        return db;
    }

  private:
    static size_t nonzero(size_t v) {
        if (v == 0) throw Error(Error::InvalidArgs, "NonZeroUsize");
        return v;
    }
    BlockVectorSet vs_;
    size_t num_partitions_ = 10, num_divisions_ = 8, num_clusters_ = 16;  // src/db/build.rs:48-50
    SeedSource seeds_;
};

}  // namespace flechasdb
