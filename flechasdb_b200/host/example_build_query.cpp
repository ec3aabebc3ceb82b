// examples/build-random + examples/query-sync of the reference (M N D P C K NPROBE from argv),
// written against the C++ host mirror.  Prints the same lines the reference's examples print.
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "flechasdb.hpp"

using namespace flechasdb;
using clk = std::chrono::steady_clock;
static double since(clk::time_point t) { return std::chrono::duration<double>(clk::now() - t).count(); }

int main(int argc, char **argv) {
    const size_t M = argc > 1 ? atol(argv[1]) : 100000, N = argc > 2 ? atol(argv[2]) : 1536;
    const size_t D = argc > 3 ? atol(argv[3]) : 12, P = argc > 4 ? atol(argv[4]) : 100;
    const size_t C = argc > 5 ? atol(argv[5]) : 256, K = 10, NPROBE = 5;
    try {
        auto ctx = std::make_shared<Context>(0);
        auto t = clk::now();
        std::vector<float> data(M * N);
        std::mt19937 rng(1);
        for (auto &x : data) x = (float)(rng() >> 8) * 5.9604645e-08f;  // rng.fill(&mut [f32])
        auto vs = BlockVectorSet::chunk(ctx, data, N);
        printf("prepared data in %f s\n", since(t));
        t = clk::now();
        size_t updates = 0;
        auto db = DatabaseBuilder(std::move(vs)).with_partitions(P).with_divisions(D).with_clusters(C).with_seed(7)
                      .build_with_events([&](const BuildEvent &e) {
                          if (e.kind == BuildEvent::Cluster && e.cluster.kind == ClusterEvent::FinishedCentroidUpdate) updates++;
                      });
        printf("built database in %f s (%zu centroid updates)\n", since(t), updates);
        std::vector<float> qv(N);
        for (auto &x : qv) x = (float)(rng() >> 8) * 5.9604645e-08f;
        for (int r = 0; r < 2; ++r) {
            t = clk::now();
            auto results = db->query(qv, K, NPROBE);
            printf("[%d] queried k-NN in %f s\n", r, since(t));
            for (size_t i = 0; i < results.size(); ++i)
                printf("\t%zu: partition=%zu, approx. distance^2=%f, vector_index=%zu\n", i,
                       results[i].partition_index, results[i].squared_distance, results[i].vector_index);
        }
    } catch (const Error &e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
