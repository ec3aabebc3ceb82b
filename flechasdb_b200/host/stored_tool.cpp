// The stored layout without a GPU (tests/test_stored_layout.py cross-checks it against the Python mirror):
//   stored_tool write <dir> <seed> <N> <P> <D> <C> <M> [attrs]   writes a database of pseudo-random arrays (and attributes), prints its header id
//   stored_tool read  <dir> <header.binpb>               reads it back (header, centroids, codebooks, every partition)
//                                                         and prints shapes + an FNV-1a checksum of every array
#include <cstdio>
#include <cstdlib>

#include "flechasdb_stored.hpp"

using namespace flechasdb;

static uint64_t lcg(uint64_t &s) {
    s = s * 6364136223846793005ULL + 1442695040888963407ULL;
    return s >> 33;
}
static uint64_t fnv(uint64_t h, const void *p, size_t n) {
    const uint8_t *b = (const uint8_t *)p;
    for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 1099511628211ULL;
    return h;
}

int main(int argc, char **argv) {
    try {
        if (argc >= 9 && std::string(argv[1]) == "write") {
            uint64_t s = strtoull(argv[3], nullptr, 10);
            const size_t N = atol(argv[4]), P = atol(argv[5]), D = atol(argv[6]), C = atol(argv[7]), M = atol(argv[8]);
            std::vector<float> coarse(P * N), cbs(D * C * (N / D));
            for (auto &x : coarse) x = (float)(lcg(s) & 0xffffff) * 5.9604645e-08f;
            for (auto &x : cbs) x = (float)(lcg(s) & 0xffffff) * 5.9604645e-08f - 0.5f;
            std::vector<uint64_t> off(P + 1, 0);
            for (size_t p = 0; p < P; ++p) off[p + 1] = off[p] + (p == 1 ? 0 : M / P + (p % 3));   // partition 1 is empty
            const size_t total = off[P];
            std::vector<uint8_t> codes(total * D);
            for (auto &c : codes) c = (uint8_t)(lcg(s) % C);
            std::vector<Uuid> ids(total);
            for (auto &id : ids)
                for (auto &b : id) b = (uint8_t)lcg(s);
            // optional 9th argument "attrs": vector i (partition-major) gets idx = i when i % 7 == 0 and name = "v<i>" when i % 5 == 0
            AttributeTable table;
            if (argc >= 10 && std::string(argv[9]) == "attrs")
                for (size_t i = 0; i < total; ++i) {
                    if (i % 7 == 0) table[ids[i]]["idx"] = AttributeValue((uint64_t)i);
                    if (i % 5 == 0) table[ids[i]]["name"] = AttributeValue("v" + std::to_string(i));
                }
            const std::string h = stored::serialize_arrays(argv[2], N, P, D, C, coarse.data(), cbs.data(), off.data(), codes.data(), ids.data(),
                                                           table.empty() ? nullptr : &table);
            printf("%s\n", h.c_str());
            return 0;
        }
        if (argc >= 4 && std::string(argv[1]) == "read") {
            const stored::Header h = stored::read_header(argv[2], argv[3]);
            uint64_t hc = fnv(1469598103934665603ULL, h.coarse.data(), h.coarse.size() * 4);
            uint64_t hb = fnv(1469598103934665603ULL, h.codebooks.data(), h.codebooks.size() * 4);
            uint64_t hk = 1469598103934665603ULL, hi = 1469598103934665603ULL;
            size_t total = 0;
            for (size_t p = 0; p < h.P; ++p) {
                const stored::PartitionData pd = stored::read_partition(argv[2], h.partition_ids[p], p, h.N, h.D, h.C);
                hk = fnv(hk, pd.codes.data(), pd.codes.size());
                for (const Uuid &u : pd.ids) hi = fnv(hi, u.data(), 16);
                total += pd.ids.size();
            }
            // the attributes logs in file order: id bytes, name, value ("s:<string>" / "u:<number>")
            uint64_t ha = 1469598103934665603ULL;
            size_t entries = 0;
            for (size_t p = 0; p < h.P && p < h.attributes_log_ids.size(); ++p)
                for (const auto &e : stored::read_attributes_log(argv[2], h.attributes_log_ids[p], h.partition_ids[p], p, h.attribute_names)) {
                    const std::string v = e.value.is_string ? "s:" + e.value.string_value : "u:" + std::to_string(e.value.uint64_value);
                    ha = fnv(fnv(fnv(ha, e.id.data(), 16), e.name.data(), e.name.size()), v.data(), v.size());
                    ++entries;
                }
            std::string names;
            for (const auto &n : h.attribute_names) names += (names.empty() ? "" : ",") + n;
            printf("N=%zu P=%zu D=%zu C=%zu M=%zu coarse=%016llx codebooks=%016llx codes=%016llx ids=%016llx attr_entries=%zu attr_names=%s attrs=%016llx\n",
                   h.N, h.P, h.D, h.C, total, (unsigned long long)hc, (unsigned long long)hb, (unsigned long long)hk, (unsigned long long)hi,
                   entries, names.c_str(), (unsigned long long)ha);
            return 0;
        }
    } catch (const Error &e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    fprintf(stderr, "usage: stored_tool write <dir> <seed> N P D C M | read <dir> <header.binpb>\n");
    return 2;
}
