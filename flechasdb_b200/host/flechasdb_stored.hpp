// The reference's on-disk layout, natively: writer (serialize_database, src/db/build/proto.rs:25-267), the
// content-addressed files of src/io.rs:170-300, and stored::Database with lazily loaded partitions
// (load_database / load_partition_centroids / load_codebook / get_partition / query,
// src/db/stored.rs:41-57,269-293,315-389,659-880) on top of the C ABI (fdb_index_create_lazy /
// fdb_index_set_partition / fdb_index_missing_partitions).  Header only; link with -lz.
//
//   <base>/<h>.binpb                 Database header                    (zlib)
//   <base>/partitions/<h>.binpb      one Partition per partition        (zlib)
//   <base>/partitions/<h>.binpb      the partition centroids VectorSet  (plain)
//   <base>/codebooks/<h>.binpb       one VectorSet per division         (plain)
//   <base>/attributes/<h>.binpb      one AttributesLog per partition    (zlib)
//
// <h> = URL-safe base64 (no padding) of the SHA-256 of the bytes ON DISK (the hasher sits under the zlib encoder,
// src/io.rs:97-106,231-235).  Messages follow src/protos/database.proto; rust-protobuf 3.2.0 writes repeated scalars
// UNPACKED (one tag per element), which the writer reproduces; the reader accepts packed and unpacked.
#pragma once

#include <zlib.h>

#include <cstdio>
#include <cstring>
#include <fstream>
#include <algorithm>
#include <map>
#include <set>
#include <sstream>
#include <sys/stat.h>

#include "flechasdb.hpp"

namespace flechasdb {
namespace stored {

// ---- SHA-256 (FIPS 180-4) ----------------------------------------------------------------------------------
inline std::array<uint8_t, 32> sha256(const std::string &data) {
    static const uint32_t K[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
        0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
        0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
        0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
        0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
        0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
        0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    std::string m = data;
    const uint64_t bits = (uint64_t)data.size() * 8;
    m.push_back((char)0x80);
    while (m.size() % 64 != 56) m.push_back((char)0);
    for (int i = 7; i >= 0; --i) m.push_back((char)(bits >> (8 * i)));
    auto rotr = [](uint32_t x, int n) { return (x >> n) | (x << (32 - n)); };
    for (size_t off = 0; off < m.size(); off += 64) {
        uint32_t w[64];
        for (int i = 0; i < 16; ++i)
            w[i] = ((uint32_t)(uint8_t)m[off + 4 * i] << 24) | ((uint32_t)(uint8_t)m[off + 4 * i + 1] << 16) |
                   ((uint32_t)(uint8_t)m[off + 4 * i + 2] << 8) | (uint32_t)(uint8_t)m[off + 4 * i + 3];
        for (int i = 16; i < 64; ++i) {
            const uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
            const uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 64; ++i) {
            const uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25), ch = (e & f) ^ (~e & g);
            const uint32_t t1 = hh + S1 + ch + K[i] + w[i];
            const uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22), mj = (a & b) ^ (a & c) ^ (b & c);
            const uint32_t t2 = S0 + mj;
            hh = g, g = f, f = e, e = d + t1, d = c, c = b, b = a, a = t1 + t2;
        }
        h[0] += a, h[1] += b, h[2] += c, h[3] += d, h[4] += e, h[5] += f, h[6] += g, h[7] += hh;
    }
    std::array<uint8_t, 32> out;
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 4; ++j) out[4 * i + j] = (uint8_t)(h[i] >> (24 - 8 * j));
    return out;
}

// URL-safe base64 without padding (base64::engine::general_purpose::URL_SAFE_NO_PAD, src/io.rs:243-255)
inline std::string base64url(const uint8_t *p, size_t n) {
    static const char *A = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789-_";
    std::string out;
    for (size_t i = 0; i < n; i += 3) {
        const uint32_t v = ((uint32_t)p[i] << 16) | ((i + 1 < n ? (uint32_t)p[i + 1] : 0u) << 8) | (i + 2 < n ? (uint32_t)p[i + 2] : 0u);
        out.push_back(A[(v >> 18) & 63]);
        out.push_back(A[(v >> 12) & 63]);
        if (i + 1 < n) out.push_back(A[(v >> 6) & 63]);
        if (i + 2 < n) out.push_back(A[v & 63]);
    }
    return out;
}

inline std::string zlib_compress(const std::string &in) {
    uLongf cap = compressBound((uLong)in.size());
    std::string out(cap, '\0');
    if (compress2((Bytef *)&out[0], &cap, (const Bytef *)in.data(), (uLong)in.size(), 6) != Z_OK)
        throw Error(Error::InvalidContext, "zlib compress failed");
    out.resize(cap);
    return out;
}
inline std::string zlib_decompress(const std::string &in) {
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit(&zs) != Z_OK) throw Error(Error::InvalidContext, "zlib init failed");
    zs.next_in = (Bytef *)in.data();
    zs.avail_in = (uInt)in.size();
    std::string out;
    char buf[1 << 16];
    int rc;
    do {
        zs.next_out = (Bytef *)buf;
        zs.avail_out = sizeof(buf);
        rc = inflate(&zs, Z_NO_FLUSH);
        if (rc != Z_OK && rc != Z_STREAM_END) {
            inflateEnd(&zs);
            throw Error(Error::InvalidData, "zlib stream is corrupt");
        }
        out.append(buf, sizeof(buf) - zs.avail_out);
    } while (rc != Z_STREAM_END);
    inflateEnd(&zs);
    return out;
}

// ---- protobuf wire format -----------------------------------------------------------------------------------
struct Writer {
    std::string b;
    void varint(uint64_t v) {
        while (v >= 0x80) {
            b.push_back((char)((v & 0x7F) | 0x80));
            v >>= 7;
        }
        b.push_back((char)v);
    }
    void tag(uint32_t field, uint32_t wire) { varint(((uint64_t)field << 3) | wire); }
    void uint32_field(uint32_t field, uint32_t v) {   // proto3: a zero is not written
        if (v) tag(field, 0), varint(v);
    }
    void string_field(uint32_t field, const std::string &s, bool always) {
        if (s.empty() && !always) return;
        tag(field, 2), varint(s.size()), b += s;
    }
    void message_field(uint32_t field, const std::string &payload) { tag(field, 2), varint(payload.size()), b += payload; }
    // `for v in data { os.write_float(field, v) }`: one tag per element
    void floats_unpacked(uint32_t field, const float *a, size_t n) {
        b.reserve(b.size() + 5 * n);
        for (size_t i = 0; i < n; ++i) {
            tag(field, 5);
            char raw[4];
            memcpy(raw, a + i, 4);
            b.append(raw, 4);
        }
    }
    void uint32s_unpacked(uint32_t field, const uint8_t *a, size_t n) {
        b.reserve(b.size() + 3 * n);
        for (size_t i = 0; i < n; ++i) tag(field, 0), varint(a[i]);
    }
    void fixed64_field(uint32_t field, uint64_t v) {
        if (!v) return;
        tag(field, 1);
        char raw[8];
        memcpy(raw, &v, 8);
        b.append(raw, 8);
    }
};

struct Field {
    uint32_t wire;
    uint64_t value;       // wire 0 / 1 / 5
    std::string bytes;    // wire 2
};
inline std::multimap<uint32_t, Field> parse(const std::string &buf) {
    std::multimap<uint32_t, Field> out;
    size_t i = 0;
    auto varint = [&]() {
        uint64_t v = 0;
        int shift = 0;
        while (true) {
            if (i >= buf.size()) throw Error(Error::InvalidData, "truncated protobuf message");
            const uint8_t c = (uint8_t)buf[i++];
            v |= (uint64_t)(c & 0x7F) << shift;
            shift += 7;
            if (!(c & 0x80)) return v;
        }
    };
    while (i < buf.size()) {
        const uint64_t key = varint();
        Field f;
        f.wire = (uint32_t)(key & 7);
        f.value = 0;
        if (f.wire == 0) f.value = varint();
        else if (f.wire == 1 || f.wire == 5) {
            const size_t n = f.wire == 1 ? 8 : 4;
            if (i + n > buf.size()) throw Error(Error::InvalidData, "truncated protobuf message");
            memcpy(&f.value, buf.data() + i, n);
            i += n;
        } else if (f.wire == 2) {
            const uint64_t n = varint();
            if (i + n > buf.size()) throw Error(Error::InvalidData, "truncated protobuf message");
            f.bytes.assign(buf, i, n);
            i += n;
        } else throw Error(Error::InvalidData, "unsupported wire type");
        out.emplace((uint32_t)(key >> 3), std::move(f));
    }
    return out;
}
// repeated float / uint32, unpacked or packed
inline std::vector<float> floats_of(const std::multimap<uint32_t, Field> &m, uint32_t field) {
    std::vector<float> out;
    auto r = m.equal_range(field);
    for (auto it = r.first; it != r.second; ++it) {
        if (it->second.wire == 5) {
            float v;
            memcpy(&v, &it->second.value, 4);
            out.push_back(v);
        } else if (it->second.wire == 2) {
            const size_t n = it->second.bytes.size() / 4, o = out.size();
            out.resize(o + n);
            memcpy(out.data() + o, it->second.bytes.data(), 4 * n);
        }
    }
    return out;
}
inline std::vector<uint32_t> uint32s_of(const std::multimap<uint32_t, Field> &m, uint32_t field) {
    std::vector<uint32_t> out;
    auto r = m.equal_range(field);
    for (auto it = r.first; it != r.second; ++it) {
        if (it->second.wire == 0) out.push_back((uint32_t)it->second.value);
        else if (it->second.wire == 2) {
            const std::string &b = it->second.bytes;
            size_t i = 0;
            while (i < b.size()) {
                uint64_t v = 0;
                int shift = 0;
                uint8_t c;
                do {
                    c = (uint8_t)b[i++];
                    v |= (uint64_t)(c & 0x7F) << shift;
                    shift += 7;
                } while ((c & 0x80) && i < b.size());
                out.push_back((uint32_t)v);
            }
        }
    }
    return out;
}
inline uint32_t uint32_of(const std::multimap<uint32_t, Field> &m, uint32_t field) {
    auto it = m.find(field);
    return it == m.end() ? 0u : (uint32_t)it->second.value;
}
inline std::vector<std::string> strings_of(const std::multimap<uint32_t, Field> &m, uint32_t field) {
    std::vector<std::string> out;
    auto r = m.equal_range(field);
    for (auto it = r.first; it != r.second; ++it) out.push_back(it->second.bytes);
    return out;
}

// ---- content-addressed files (src/io.rs:170-300) --------------------------------------------------------------
inline void make_dirs(const std::string &path) {
    for (size_t i = 1; i <= path.size(); ++i)
        if (i == path.size() || path[i] == '/') mkdir(path.substr(0, i).c_str(), 0777);
}
inline std::string persist(const std::string &base, const std::string &sub, const std::string &payload, bool compressed) {
    const std::string data = compressed ? zlib_compress(payload) : payload;
    const auto digest = sha256(data);
    const std::string h = base64url(digest.data(), digest.size());
    const std::string dir = sub.empty() ? base : base + "/" + sub;
    make_dirs(dir);
    std::ofstream f(dir + "/" + h + ".binpb", std::ios::binary);
    if (!f) throw Error(Error::InvalidContext, "cannot write " + dir + "/" + h + ".binpb");
    f.write(data.data(), (std::streamsize)data.size());
    return h;
}
inline std::string open_file(const std::string &base, const std::string &rel, bool compressed, bool verify = true) {
    const std::string path = base + "/" + rel;
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Error(Error::InvalidContext, "cannot read " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string data = ss.str();
    if (verify) {   // HashedFileIn::verify (src/io.rs:287-299)
        const auto digest = sha256(data);
        const std::string h = base64url(digest.data(), digest.size());
        std::string stem = rel.substr(rel.find_last_of('/') == std::string::npos ? 0 : rel.find_last_of('/') + 1);
        stem = stem.substr(0, stem.find_last_of('.'));
        if (h != stem) throw Error(Error::InvalidData, "VerificationFailure: expected hash " + stem + ", got " + h);
    }
    return compressed ? zlib_decompress(data) : data;
}

inline std::string vector_set_message(const float *data, size_t n, size_t vector_size) {
    Writer w;
    w.uint32_field(1, (uint32_t)vector_size);
    w.floats_unpacked(10, data, n * vector_size);
    return w.b;
}

// serialize_database (src/db/build/proto.rs:25-267) from plain arrays: coarse [P][N], codebooks [D][C][N/D],
// offsets [P+1], codes [M][D] partition-major, ids [M] partition-major.  Returns the header id.
inline std::string uuid_message(const Uuid &id) {   // Uuid { fixed64 upper = 1; fixed64 lower = 2 } (src/protos/mod.rs:21-27)
    uint64_t upper = 0, lower = 0;
    for (int i = 0; i < 8; ++i) upper = (upper << 8) | id[i], lower = (lower << 8) | id[8 + i];
    Writer u;
    u.fixed64_field(1, upper);
    u.fixed64_field(2, lower);
    return u.b;
}
// AttributeValue (src/db/proto.rs:15-24): a oneof member is written even when it holds the default value
inline std::string attribute_value_message(const AttributeValue &v) {
    Writer w;
    if (v.is_string) w.tag(1, 2), w.varint(v.string_value.size()), w.b += v.string_value;
    else w.tag(2, 0), w.varint(v.uint64_value);
    return w.b;
}
// AttributesLog of one partition (src/db/build/proto.rs:174-200): one OperationSetAttribute per (vector of the
// partition in ascending vector index, attribute of that vector); name_index = position in the sorted names
inline std::string attributes_log_message(const std::string &partition_id, const Uuid *ids, size_t n,
                                          const AttributeTable *table, const std::vector<std::string> &names) {
    Writer w;
    w.string_field(1, partition_id, false);
    for (size_t v = 0; v < n && table; ++v) {
        auto it = table->find(ids[v]);
        if (it == table->end()) continue;
        for (const auto &kv : it->second) {
            const auto pos = std::lower_bound(names.begin(), names.end(), kv.first);
            if (pos == names.end() || *pos != kv.first) throw Error(Error::InvalidContext, "attribute name must be encoded: " + kv.first);
            Writer e;
            e.message_field(1, uuid_message(ids[v]));
            e.uint32_field(2, (uint32_t)(pos - names.begin()));
            e.message_field(3, attribute_value_message(kv.second));
            w.message_field(10, e.b);
        }
    }
    return w.b;
}

inline std::string serialize_arrays(const std::string &base, size_t N, size_t P, size_t D, size_t C, const float *coarse,
                                    const float *codebooks, const uint64_t *offsets, const uint8_t *codes, const Uuid *ids,
                                    const AttributeTable *attributes = nullptr) {
    std::vector<std::string> partition_ids, codebook_ids, log_ids;
    for (size_t p = 0; p < P; ++p) {
        const size_t lo = offsets[p], hi = offsets[p + 1];
        Writer enc, w;
        enc.uint32_field(1, (uint32_t)D);
        enc.uint32s_unpacked(10, codes + lo * D, (hi - lo) * D);
        w.uint32_field(1, (uint32_t)N);
        w.uint32_field(2, (uint32_t)D);
        w.floats_unpacked(10, coarse + p * N, N);
        w.message_field(11, enc.b);
        for (size_t v = lo; v < hi; ++v) w.message_field(12, uuid_message(ids[v]));
        partition_ids.push_back(persist(base, "partitions", w.b, true));
    }
    const std::string centroids_id = persist(base, "partitions", vector_set_message(coarse, P, N), false);
    const size_t s = N / D;
    for (size_t d = 0; d < D; ++d) codebook_ids.push_back(persist(base, "codebooks", vector_set_message(codebooks + d * C * s, C, s), false));
    // get_sorted_attribute_names (src/db/build/proto.rs:149-158): a BTreeSet<String>, i.e. sorted by bytes
    std::vector<std::string> names;
    if (attributes) {
        std::set<std::string> sorted;
        for (const auto &kv : *attributes)
            for (const auto &a : kv.second) sorted.insert(a.first);
        names.assign(sorted.begin(), sorted.end());
    }
    for (size_t p = 0; p < P; ++p)
        log_ids.push_back(persist(base, "attributes",
                                  attributes_log_message(partition_ids[p], ids + offsets[p], offsets[p + 1] - offsets[p], attributes, names), true));
    Writer h;
    h.uint32_field(1, (uint32_t)N), h.uint32_field(2, (uint32_t)P), h.uint32_field(3, (uint32_t)D), h.uint32_field(4, (uint32_t)C);
    for (const auto &x : partition_ids) h.string_field(10, x, true);
    h.string_field(11, centroids_id, false);
    for (const auto &x : codebook_ids) h.string_field(12, x, true);
    for (const auto &x : log_ids) h.string_field(13, x, true);
    for (const auto &x : names) h.string_field(14, x, true);
    return persist(base, "", h.b, true);
}

// serialize_database for a database built on the GPU
inline std::string serialize_database(const flechasdb::Database &db, const std::string &base) {
    const size_t N = db.vector_size(), P = db.num_partitions(), D = db.num_divisions(), C = db.num_clusters(), M = db.num_vectors();
    std::vector<float> coarse(P * N), cbs(D * C * (N / D));
    db.quantisers(coarse.data(), cbs.data());
    std::vector<uint64_t> off(P + 1);
    std::vector<uint32_t> order(M);
    std::vector<uint8_t> codes(M * D);
    check(fdb_index_get_layout(db.index(), off.data(), order.data(), codes.data()));
    std::vector<Uuid> ids(M);
    for (size_t i = 0; i < M; ++i) ids[i] = db.vector_ids()[order[i]];
    return serialize_arrays(base, N, P, D, C, coarse.data(), cbs.data(), off.data(), codes.data(), ids.data(), &db.attribute_table());
}

// what load_database reads (src/db/stored.rs:659-798): the header, the partition centroids, the codebooks
struct Header {
    size_t N = 0, P = 0, D = 0, C = 0;
    std::vector<std::string> partition_ids, attributes_log_ids, attribute_names;
    std::vector<float> coarse, codebooks;   // [P][N], [D][C][N/D]
};
inline Header read_header(const std::string &base, const std::string &path) {
    Header h;
    const auto hdr = parse(open_file(base, path, true));
    const size_t N = uint32_of(hdr, 1), P = uint32_of(hdr, 2), D = uint32_of(hdr, 3), C = uint32_of(hdr, 4);
    if (N == 0) throw Error(Error::InvalidData, "vector_size is zero");     // src/db/stored.rs:671-706
    if (D == 0) throw Error(Error::InvalidData, "num_divisions is zero");
    if (P == 0) throw Error(Error::InvalidData, "num_partitions is zero");
    if (C == 0) throw Error(Error::InvalidData, "num_codes is zero");
    if (N % D) throw Error(Error::InvalidData, "vector_size " + std::to_string(N) + " is not multiple of num_divisions " + std::to_string(D));
    h.partition_ids = strings_of(hdr, 10);
    const auto codebook_ids = strings_of(hdr, 12);
    const auto centroid_ids = strings_of(hdr, 11);
    if (h.partition_ids.size() != P)
        throw Error(Error::InvalidData, "num_partitions " + std::to_string(P) + " and partition_ids.len() " + std::to_string(h.partition_ids.size()) + " do not match");
    if (codebook_ids.size() != D)
        throw Error(Error::InvalidData, "num_divisions " + std::to_string(D) + " and codebook_ids.len() " + std::to_string(codebook_ids.size()) + " do not match");
    if (centroid_ids.empty()) throw Error(Error::InvalidData, "partition_centroids_id is missing");
    h.N = N, h.P = P, h.D = D, h.C = C;
    h.attributes_log_ids = strings_of(hdr, 13);
    h.attribute_names = strings_of(hdr, 14);
    // load_partition_centroids never calls verify() (src/db/stored.rs:729-755)
    const auto cen = parse(open_file(base, "partitions/" + centroid_ids[0] + ".binpb", false, false));
    if (uint32_of(cen, 1) != N) throw Error(Error::InvalidData, "partition centroids vector size mismatch");
    h.coarse = floats_of(cen, 10);
    if (h.coarse.size() != P * N) throw Error(Error::InvalidData, "partition centroids data length mismatch");
    const size_t s = N / D;
    for (size_t d = 0; d < D; ++d) {   // load_codebook (src/db/stored.rs:757-798)
        const auto cb = parse(open_file(base, "codebooks/" + codebook_ids[d] + ".binpb", false));
        const std::vector<float> data = floats_of(cb, 10);
        if (uint32_of(cb, 1) != s || data.size() != C * s) throw Error(Error::InvalidData, "codebook " + std::to_string(d) + " shape mismatch");
        h.codebooks.insert(h.codebooks.end(), data.begin(), data.end());
    }
    return h;
}
// load_partition (src/db/stored.rs:800-880): the codes [n][D] (u8) and vector ids [n] of one partition
struct PartitionData {
    std::vector<uint8_t> codes;
    std::vector<Uuid> ids;
};
inline PartitionData read_partition(const std::string &base, const std::string &id, size_t p, size_t N, size_t D, size_t C) {
    PartitionData out;
    const auto f = parse(open_file(base, "partitions/" + id + ".binpb", true));
    if (uint32_of(f, 1) != N || uint32_of(f, 2) != D) throw Error(Error::InvalidData, "partition " + std::to_string(p) + " shape mismatch");
    std::vector<uint32_t> data;
    auto enc = f.find(11);
    if (enc != f.end()) data = uint32s_of(parse(enc->second.bytes), 10);
    if (data.size() % D) throw Error(Error::InvalidData, "encoded vectors of partition " + std::to_string(p) + " are not a multiple of " + std::to_string(D));
    const size_t n = data.size() / D;
    out.codes.resize(data.size());
    for (size_t i = 0; i < data.size(); ++i) {
        // a code indexes table[di * num_codes + code] (src/db/stored.rs:585): the reference panics past the table
        if (data[i] >= C || data[i] > 255) throw Error(Error::InvalidData, "partition " + std::to_string(p) + " holds the code " + std::to_string(data[i]));
        out.codes[i] = (uint8_t)data[i];
    }
    auto r = f.equal_range(12);
    for (auto it = r.first; it != r.second; ++it) {
        const auto u = parse(it->second.bytes);
        uint64_t upper = 0, lower = 0;
        auto a = u.find(1), b = u.find(2);
        if (a != u.end()) upper = a->second.value;
        if (b != u.end()) lower = b->second.value;
        Uuid uid;
        for (int i = 0; i < 8; ++i) uid[i] = (uint8_t)(upper >> (56 - 8 * i)), uid[8 + i] = (uint8_t)(lower >> (56 - 8 * i));
        out.ids.push_back(uid);
    }
    if (out.ids.size() != n) throw Error(Error::InvalidData, "partition " + std::to_string(p) + ": vector ids do not match the vectors");
    return out;
}

// one AttributesLog, validated like load_attributes_log (src/db/stored.rs:185-249): (vector id, name, value) per entry
struct AttributeEntry {
    Uuid id;
    std::string name;
    AttributeValue value;
};
inline Uuid uuid_of(const std::string &bytes) {
    const auto u = parse(bytes);
    uint64_t upper = 0, lower = 0;
    auto a = u.find(1), b = u.find(2);
    if (a != u.end()) upper = a->second.value;
    if (b != u.end()) lower = b->second.value;
    Uuid uid;
    for (int i = 0; i < 8; ++i) uid[i] = (uint8_t)(upper >> (56 - 8 * i)), uid[8 + i] = (uint8_t)(lower >> (56 - 8 * i));
    return uid;
}
inline std::vector<AttributeEntry> read_attributes_log(const std::string &base, const std::string &log_id, const std::string &partition_id,
                                                       size_t p, const std::vector<std::string> &names) {
    const auto f = parse(open_file(base, "attributes/" + log_id + ".binpb", true));
    auto pid = f.find(1);
    const std::string got = pid == f.end() ? std::string() : pid->second.bytes;
    if (got != partition_id) throw Error(Error::InvalidData, "inconsistent partition IDs: " + got + " vs " + partition_id);
    std::vector<AttributeEntry> out;
    auto r = f.equal_range(10);
    size_t i = 0;
    for (auto it = r.first; it != r.second; ++it, ++i) {
        const auto e = parse(it->second.bytes);
        const uint32_t ni = uint32_of(e, 2);
        if (ni >= names.size()) throw Error(Error::InvalidData, "attribute name index out of bounds: " + std::to_string(ni));
        auto id = e.find(1), val = e.find(3);
        const std::string where = "attributes log[" + std::to_string(p) + ", " + std::to_string(i) + "]: ";
        if (id == e.end()) throw Error(Error::InvalidData, where + "missing vector ID");
        if (val == e.end()) throw Error(Error::InvalidData, where + "missing value");
        const auto v = parse(val->second.bytes);
        AttributeEntry ent;
        ent.id = uuid_of(id->second.bytes);
        ent.name = names[ni];
        auto sv = v.find(1), uv = v.find(2);
        if (sv != v.end()) ent.value = AttributeValue(sv->second.bytes);
        else if (uv != v.end()) ent.value = AttributeValue((uint64_t)uv->second.value);
        else throw Error(Error::InvalidData, where + "missing value");
        out.push_back(std::move(ent));
    }
    return out;
}

// stored::Database<f32, LocalFileSystem> (src/db/stored.rs:41-57): header, partition centroids and codebooks are
// read by load_database; a partition (codes + vector ids) is read and uploaded when a query first probes it
// (get_partition, src/db/stored.rs:269-293).
class Database {
  public:
    static std::unique_ptr<Database> load_database(std::shared_ptr<Context> ctx, const std::string &base, const std::string &path) {
        std::unique_ptr<Database> db(new Database);
        db->ctx_ = ctx;
        db->base_ = base;
        const Header h = read_header(base, path);
        db->N_ = h.N, db->P_ = h.P, db->D_ = h.D, db->C_ = h.C;
        db->partition_ids_ = h.partition_ids;
        db->attributes_log_ids_ = h.attributes_log_ids;
        db->attribute_names_ = h.attribute_names;
        db->log_loaded_.assign(h.P, false);
        check(fdb_index_create_lazy(ctx->h, h.N, h.P, h.D, h.C, h.coarse.data(), h.codebooks.data(), &db->index_));
        db->ids_.resize(h.P);
        return db;
    }

    // get_partition (src/db/stored.rs:269-293): read, validate and upload partition p once
    void load_partition(size_t p) {
        if (fdb_index_partition_loaded(index_, p)) return;
        PartitionData pd = read_partition(base_, partition_ids_[p], p, N_, D_, C_);
        check(fdb_index_set_partition(index_, p, pd.codes.data(), pd.ids.size()));
        ids_[p] = std::move(pd.ids);
    }

    // stored::Database::query_with_events (src/db/stored.rs:331-389); event order is the reference's
    std::vector<QueryResult> query(const std::vector<float> &v, size_t k, size_t nprobe,
                                   const std::function<void(const QueryEvent &)> &ev = [](const QueryEvent &) {}) {
        if (k == 0 || nprobe == 0) throw Error(Error::InvalidArgs, "NonZeroUsize");
        if (v.size() != N_) throw Error(Error::InvalidArgs, "query vector size mismatch");
        // (the partition centroids and the codebooks are resident since load_database: nothing to initialise lazily)
        ev({QueryEvent::StartingQueryInitialization});
        ev({QueryEvent::FinishedQueryInitialization});
        ev({QueryEvent::StartingPartitionSelection});
        std::vector<uint32_t> probes(nprobe);
        check(fdb_index_probe(index_, v.data(), 1, nprobe, FDB_QUERY_STORED, probes.data(), nullptr));
        ev({QueryEvent::FinishedPartitionSelection});
        for (uint32_t p : probes) {
            ev({QueryEvent::StartingPartitionQuery, p});
            load_partition(p);     // lazy: the partition's file is read on its first probe
            ev({QueryEvent::FinishedPartitionQuery, p});
        }
        std::vector<uint32_t> part(k), vidx(k);
        std::vector<float> dist(k);
        uint32_t count = 0;
        check(fdb_index_query(index_, v.data(), 1, k, nprobe, FDB_QUERY_STORED, part.data(), vidx.data(), dist.data(), &count));
        ev({QueryEvent::StartingResultSelection});
        std::vector<QueryResult> out;
        for (uint32_t i = 0; i < count; ++i) out.push_back({part[i], ids_[part[i]][vidx[i]], vidx[i], dist[i]});
        ev({QueryEvent::FinishedResultSelection});
        return out;
    }

    // batched form (no reference analogue): loads whatever the batch probes, then one device batch
    void query_batch(const float *queries, size_t nq, size_t k, size_t nprobe, uint32_t *part, uint32_t *vidx, float *dist,
                     uint32_t *count) {
        size_t missing = 0;
        std::vector<uint32_t> need(P_);
        check(fdb_index_missing_partitions(index_, queries, nq, nprobe, FDB_QUERY_STORED, need.data(), need.size(), &missing));
        for (size_t i = 0; i < missing; ++i) load_partition(need[i]);
        check(fdb_index_query(index_, queries, nq, k, nprobe, FDB_QUERY_STORED, part, vidx, dist, count));
    }
    // Database::get_attribute (src/db/stored.rs:118-131): loads every attributes log on the first call; null when the
    // vector exists but has no such attribute, InvalidArgs when no vector has this id
    const AttributeValue *get_attribute(const Uuid &id, const std::string &key) {
        if (!table_loaded_) {
            for (size_t p = 0; p < P_; ++p) load_attributes_log(p);
            table_loaded_ = true;
        }
        return get_attribute_internal(id, key);
    }
    // QueryResult::get_attribute (src/db/stored.rs:621-634): only the log of the result's partition is loaded
    const AttributeValue *get_attribute(const QueryResult &r, const std::string &key) {
        load_attributes_log(r.partition_index);
        return get_attribute_internal(r.vector_id, key);
    }
    // load_attributes_log (src/db/stored.rs:185-260): also loads the partition; its vectors get empty attribute maps
    void load_attributes_log(size_t p) {
        if (p >= P_) throw Error(Error::InvalidArgs, "partition index out of bounds");
        if (log_loaded_[p]) return;
        load_partition(p);
        if (p >= attributes_log_ids_.size()) throw Error(Error::InvalidData, "no attributes log for partition " + std::to_string(p));
        for (auto &e : read_attributes_log(base_, attributes_log_ids_[p], partition_ids_[p], p, attribute_names_))
            attribute_table_[e.id][e.name] = e.value;      // the last set operation wins
        for (const Uuid &id : ids_[p]) attribute_table_[id];
        log_loaded_[p] = true;
    }
    size_t loaded_partitions() const {
        size_t n = 0;
        for (size_t p = 0; p < P_; ++p) n += fdb_index_partition_loaded(index_, p);
        return n;
    }
    size_t vector_size() const { return N_; }
    size_t num_partitions() const { return P_; }
    size_t num_divisions() const { return D_; }
    size_t num_codes() const { return C_; }
    ~Database() { fdb_index_destroy(index_); }

  private:
    Database() = default;
    const AttributeValue *get_attribute_internal(const Uuid &id, const std::string &key) const {
        auto it = attribute_table_.find(id);
        if (it == attribute_table_.end()) throw Error(Error::InvalidArgs, "no such vector ID");
        auto a = it->second.find(key);
        return a == it->second.end() ? nullptr : &a->second;
    }
    std::vector<std::string> attributes_log_ids_, attribute_names_;
    std::vector<bool> log_loaded_;
    bool table_loaded_ = false;
    AttributeTable attribute_table_;
    std::shared_ptr<Context> ctx_;
    std::string base_;
    size_t N_ = 0, P_ = 0, D_ = 0, C_ = 0;
    std::vector<std::string> partition_ids_;
    std::vector<std::vector<Uuid>> ids_;
    fdb_index *index_ = nullptr;
};

}  // namespace stored
}  // namespace flechasdb
