// build -> serialize_database -> load_database (lazy) -> query: the reference's save-and-reload round trip
// (examples/build-random writes the database, examples/query-sync loads and queries it), against the C++ host
// mirror.  Exits 0 and prints "STORED_OK ..." when the lazily loaded stored database answers exactly like the
// in-memory one (stored semantic), partitions are loaded only when probed, and live ClusterEvents arrive in the
// reference's order.   usage: example_stored <dir> [M N D P C]
#include <cstdio>
#include <cstdlib>

#include "flechasdb_stored.hpp"

using namespace flechasdb;

int main(int argc, char **argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: example_stored <dir> [M N D P C]\n");
        return 2;
    }
    const std::string base = argv[1];
    const size_t M = argc > 2 ? atol(argv[2]) : 6000, N = argc > 3 ? atol(argv[3]) : 64;
    const size_t D = argc > 4 ? atol(argv[4]) : 4, P = argc > 5 ? atol(argv[5]) : 16, C = argc > 6 ? atol(argv[6]) : 32;
    const size_t K = 7, NPROBE = 3, NQ = 24;
    try {
        auto ctx = std::make_shared<Context>(0);
        std::vector<float> data(M * N);
        std::mt19937 rng(3);
        for (auto &x : data) x = (float)(rng() >> 8) * 5.9604645e-08f;
        // live events: a k-means driven step by step fires every ClusterEvent when its phase happens
        {
            auto vs = BlockVectorSet::chunk(ctx, data, N);
            SeedSource seeds(11);
            std::vector<int> kinds;
            auto run = cluster_device_live(vs, 0, N, 8, seeds, [&](const ClusterEvent &e) { kinds.push_back((int)e.kind); });
            bool ok = kinds.size() >= 4 && kinds[0] == ClusterEvent::StartingCentroidInitialization &&
                      kinds[1] == ClusterEvent::FinishedCentroidInitialization && kinds[2] == ClusterEvent::StartingCentroidUpdate &&
                      kinds[3] == ClusterEvent::FinishedCentroidUpdate && kinds.back() == ClusterEvent::FinishedCentroidUpdate + 0 * 0;
            // the loop ends after an update (convergence) or after the 100th reassignment
            ok = ok && (kinds.back() == ClusterEvent::FinishedCentroidUpdate || kinds.back() == ClusterEvent::FinishedCentroidReassignment);
            if (!ok) {
                fprintf(stderr, "live ClusterEvent order is wrong\n");
                return 1;
            }
        }
        auto db = DatabaseBuilder(BlockVectorSet::chunk(ctx, data, N)).with_partitions(P).with_divisions(D).with_clusters(C).with_seed(7).build();
        // attributes (src/db/build.rs:252-285): set on the built database, written as one log per partition
        for (size_t i = 0; i < M; i += 3) {
            db->set_attribute_at(i, "index", AttributeValue((uint64_t)i));
            if (i % 2 == 0) db->set_attribute_at(i, "label", AttributeValue("vector " + std::to_string(i)));
        }
        db->set_attribute_at(0, "label", AttributeValue("replaced"));
        const std::string header = stored::serialize_database(*db, base);
        auto sdb = stored::Database::load_database(ctx, base, header + ".binpb");
        if (sdb->loaded_partitions() != 0) {
            fprintf(stderr, "load_database must not load partitions\n");
            return 1;
        }
        size_t checked = 0;
        for (size_t qi = 0; qi < NQ; ++qi) {
            std::vector<float> qv(N);
            for (auto &x : qv) x = (float)(rng() >> 8) * 5.9604645e-08f;
            const auto want = db->query(qv, K, NPROBE, [](const QueryEvent &) {}, FDB_QUERY_STORED);
            const auto got = sdb->query(qv, K, NPROBE);
            if (got.size() != want.size()) {
                fprintf(stderr, "query %zu: %zu results, expected %zu\n", qi, got.size(), want.size());
                return 1;
            }
            for (size_t i = 0; i < got.size(); ++i)
                if (got[i].partition_index != want[i].partition_index || got[i].vector_index != want[i].vector_index ||
                    got[i].squared_distance != want[i].squared_distance || got[i].vector_id != want[i].vector_id) {
                    fprintf(stderr, "query %zu result %zu differs\n", qi, i);
                    return 1;
                }
            // QueryResult::get_attribute (src/db/stored.rs:621-634): what the built database holds for that vector id
            for (const auto &r : got) {
                size_t gi = 0;
                while (gi < M && db->vector_ids()[gi] != r.vector_id) ++gi;
                const AttributeValue *a = sdb->get_attribute(r, "index"), *l = sdb->get_attribute(r, "label");
                const bool has = gi % 3 == 0;
                const std::string label = gi == 0 ? "replaced" : "vector " + std::to_string(gi);
                if (gi == M || (a != nullptr) != has || (a && (a->is_string || a->uint64_value != gi)) ||
                    (l != nullptr) != (has && gi % 2 == 0) || (l && (!l->is_string || l->string_value != label))) {
                    fprintf(stderr, "query %zu: attributes of vector %zu differ\n", qi, gi);
                    return 1;
                }
            }
            if (qi == 0 && sdb->loaded_partitions() > NPROBE) {
                fprintf(stderr, "the first query loaded %zu partitions, it probes %zu\n", sdb->loaded_partitions(), NPROBE);
                return 1;
            }
            checked += got.size();
        }
        // batched form: loads what the batch probes, then one device batch
        std::vector<float> batch(NQ * N);
        for (auto &x : batch) x = (float)(rng() >> 8) * 5.9604645e-08f;
        std::vector<uint32_t> part(NQ * K), vidx(NQ * K), cnt(NQ), part2(NQ * K), vidx2(NQ * K), cnt2(NQ);
        std::vector<float> dist(NQ * K), dist2(NQ * K);
        sdb->query_batch(batch.data(), NQ, K, NPROBE, part.data(), vidx.data(), dist.data(), cnt.data());
        db->query_batch(batch.data(), NQ, K, NPROBE, FDB_QUERY_STORED, part2.data(), vidx2.data(), dist2.data(), cnt2.data());
        if (part != part2 || vidx != vidx2 || dist != dist2 || cnt != cnt2) {
            fprintf(stderr, "batched stored query differs from the in-memory database\n");
            return 1;
        }
        // Database::get_attribute loads every log; an unknown id is InvalidArgs
        const AttributeValue *v0 = sdb->get_attribute(db->vector_ids()[0], "label");
        bool threw = false;
        try {
            sdb->get_attribute(Uuid{}, "label");
        } catch (const Error &e) {
            threw = e.kind == Error::InvalidArgs;
        }
        if (!v0 || v0->string_value != "replaced" || !threw || sdb->loaded_partitions() != P) {
            fprintf(stderr, "get_attribute over the whole table is wrong\n");
            return 1;
        }
        printf("STORED_OK header=%s results=%zu loaded_partitions=%zu/%zu\n", header.c_str(), checked, sdb->loaded_partitions(), P);
    } catch (const Error &e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
