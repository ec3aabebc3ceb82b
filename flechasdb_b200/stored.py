"""The reference's on-disk layout: writer and reader (host-side I/O, no arithmetic).

Mirrors serialize_database (src/db/build/proto.rs:25-267), the content-addressed files of
src/io.rs:170-300 and load_database / load_partition_centroids / load_codebook /
load_partition of src/db/stored.rs:659-880:

    <base>/<h>.binpb                 Database header                    (zlib)
    <base>/partitions/<h>.binpb      one Partition per partition        (zlib)
    <base>/partitions/<h>.binpb      the partition centroids VectorSet  (plain)
    <base>/codebooks/<h>.binpb       one VectorSet per division         (plain)
    <base>/attributes/<h>.binpb      one AttributesLog per partition    (zlib)

<h> = URL-safe base64 (no padding) of the SHA-256 of the bytes ON DISK (the hasher sits under
the zlib encoder, src/io.rs:97-106,231-235).  Messages follow src/protos/database.proto;
rust-protobuf 3.2.0 writes repeated scalars UNPACKED (one tag per element), which this writer
reproduces; the reader accepts packed and unpacked.  Compressed bytes (and therefore file
names) depend on the zlib implementation; the files verify against their own names.
"""
import base64
import hashlib
import os
import struct
import uuid
import zlib

import numpy as np

EXT = "binpb"


# ---- protobuf wire format ----------------------------------------------------------------------
def _varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _tag(field, wire):
    return _varint((field << 3) | wire)


def _uint32_field(field, v):
    return b"" if v == 0 else _tag(field, 0) + _varint(int(v))   # proto3: zero is not written


def _string_field(field, s, always=False):
    b = s.encode()
    return b"" if (not b and not always) else _tag(field, 2) + _varint(len(b)) + b


def _message_field(field, payload):
    return _tag(field, 2) + _varint(len(payload)) + payload


def _floats_unpacked(field, a):
    """`for v in data { os.write_float(field, v) }`: tag + 4 little-endian bytes per element"""
    a = np.ascontiguousarray(a, "<f4").reshape(-1)
    tag = _tag(field, 5)
    assert len(tag) == 1
    out = np.empty((a.size, 5), np.uint8)
    out[:, 0] = tag[0]
    out[:, 1:] = a.view(np.uint8).reshape(-1, 4)
    return out.tobytes()


def _uint32s_unpacked(field, a):
    """`for v in data { os.write_uint32(field, v) }` for values < 2^14 (PQ codes)"""
    a = np.ascontiguousarray(a, np.uint32).reshape(-1)
    assert a.size == 0 or int(a.max()) < (1 << 14)
    tag = _tag(field, 0)
    assert len(tag) == 1
    two = a >= 128
    length = 2 + two.astype(np.int64)
    pos = np.concatenate([[0], np.cumsum(length)])
    out = np.empty(int(pos[-1]), np.uint8)
    start = pos[:-1]
    out[start] = tag[0]
    out[start + 1] = np.where(two, (a & 0x7F) | 0x80, a).astype(np.uint8)
    out[start[two] + 2] = (a[two] >> 7).astype(np.uint8)
    return out.tobytes()


def _uuid_message(u16):
    """Uuid { fixed64 upper = 1; fixed64 lower = 2 } from 16 big-endian bytes
    (Uuid::as_u64_pair, src/protos/mod.rs:21-27); zero halves are not written."""
    upper, lower = struct.unpack(">QQ", bytes(u16))
    out = b""
    if upper:
        out += _tag(1, 1) + struct.pack("<Q", upper)
    if lower:
        out += _tag(2, 1) + struct.pack("<Q", lower)
    return out


def parse(buf):
    """generic wire parser -> {field: [values]} (varint int, 64-bit bytes, length-delimited
    bytes, 32-bit bytes)"""
    fields = {}
    i, n = 0, len(buf)
    while i < n:
        key = 0
        shift = 0
        while True:
            b = buf[i]
            i += 1
            key |= (b & 0x7F) << shift
            shift += 7
            if not b & 0x80:
                break
        field, wire = key >> 3, key & 7
        if wire == 0:
            v = 0
            shift = 0
            while True:
                b = buf[i]
                i += 1
                v |= (b & 0x7F) << shift
                shift += 7
                if not b & 0x80:
                    break
        elif wire == 1:
            v = buf[i:i + 8]
            i += 8
        elif wire == 5:
            v = buf[i:i + 4]
            i += 4
        elif wire == 2:
            ln = 0
            shift = 0
            while True:
                b = buf[i]
                i += 1
                ln |= (b & 0x7F) << shift
                shift += 7
                if not b & 0x80:
                    break
            v = buf[i:i + ln]
            i += ln
        else:
            raise ValueError("unsupported wire type %d" % wire)
        fields.setdefault(field, []).append((wire, v))
    return fields


def _parse_floats(buf, field):
    """repeated float, unpacked fast path (tag + 4 bytes per element) or generic"""
    tag = _tag(field, 5)[0]
    b = np.frombuffer(buf, np.uint8)
    # find where the run of unpacked elements starts: skip leading other fields generically
    f = parse_prefix(buf, stop_field=field)
    rest = b[f:]
    if rest.size and rest.size % 5 == 0 and (rest.reshape(-1, 5)[:, 0] == tag).all():
        return np.ascontiguousarray(rest.reshape(-1, 5)[:, 1:]).view("<f4").reshape(-1).copy()
    out = []
    for wire, v in parse(buf).get(field, []):
        if wire == 5:
            out.append(np.frombuffer(v, "<f4"))
        else:  # packed
            out.append(np.frombuffer(v, "<f4"))
    return np.concatenate(out) if out else np.zeros(0, np.float32)


def parse_prefix(buf, stop_field):
    """offset of the first occurrence of `stop_field` when every earlier field is a varint"""
    i, n = 0, len(buf)
    while i < n:
        j = i
        key = 0
        shift = 0
        while True:
            b = buf[j]
            j += 1
            key |= (b & 0x7F) << shift
            shift += 7
            if not b & 0x80:
                break
        if key >> 3 == stop_field or key & 7 != 0:
            return i
        while buf[j] & 0x80:
            j += 1
        i = j + 1
    return n


def _parse_uint32s(buf, field):
    out = []
    for wire, v in parse(buf).get(field, []):
        if wire == 0:
            out.append(v)
        else:  # packed varints
            i = 0
            while i < len(v):
                x = 0
                shift = 0
                while True:
                    b = v[i]
                    i += 1
                    x |= (b & 0x7F) << shift
                    shift += 7
                    if not b & 0x80:
                        break
                out.append(x)
    return np.array(out, np.uint32)


# ---- content-addressed files (src/io.rs) -----------------------------------------------------------
class Error(Exception):
    def __init__(self, kind, msg):
        super().__init__(msg)
        self.kind = kind


def _persist(base, sub, payload, compressed):
    data = zlib.compress(payload, 6) if compressed else payload
    h = base64.urlsafe_b64encode(hashlib.sha256(data).digest()).rstrip(b"=").decode()
    d = os.path.join(base, sub) if sub else base
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, h + "." + EXT), "wb") as f:
        f.write(data)
    return h


def _open(base, rel, compressed, verify=True):
    path = os.path.join(base, rel)
    data = open(path, "rb").read()
    if verify:
        h = base64.urlsafe_b64encode(hashlib.sha256(data).digest()).rstrip(b"=").decode()
        stem = os.path.splitext(os.path.basename(path))[0]
        if h != stem:
            raise Error("VerificationFailure", "Expected hash %r, but got %s" % (stem, h))
    return zlib.decompress(data) if compressed else data


# ---- messages ---------------------------------------------------------------------------------------
def vector_set_message(data2d):
    data2d = np.ascontiguousarray(data2d, np.float32)
    return _uint32_field(1, data2d.shape[1]) + _floats_unpacked(10, data2d)


def partition_message(centroid, codes, ids16):
    codes = np.ascontiguousarray(codes, np.uint32)
    D = codes.shape[1]
    enc = _uint32_field(1, D) + _uint32s_unpacked(10, codes)
    out = _uint32_field(1, len(centroid)) + _uint32_field(2, D) + _floats_unpacked(10, centroid)
    out += _message_field(11, enc)
    out += b"".join(_message_field(12, _uuid_message(u)) for u in ids16)
    return out


def database_message(N, P, D, C, partition_ids, centroids_id, codebook_ids, attributes_log_ids,
                     attribute_names=()):
    out = _uint32_field(1, N) + _uint32_field(2, P) + _uint32_field(3, D) + _uint32_field(4, C)
    out += b"".join(_string_field(10, s, True) for s in partition_ids)
    out += _string_field(11, centroids_id)
    out += b"".join(_string_field(12, s, True) for s in codebook_ids)
    out += b"".join(_string_field(13, s, True) for s in attributes_log_ids)
    out += b"".join(_string_field(14, s, True) for s in attribute_names)
    return out


def attribute_value_message(value):
    """AttributeValue (src/db/proto.rs:15-24): oneof { string string_value = 1; uint64 uint64_value = 2; }.  A oneof
    member is written even when it holds the default value."""
    if isinstance(value, str):
        b = value.encode()
        return _tag(1, 2) + _varint(len(b)) + b
    return _tag(2, 0) + _varint(int(value))


def attributes_log_message(partition_id, ids16, attribute_table, attribute_names):
    """AttributesLog of one partition (src/db/build/proto.rs:174-200): one OperationSetAttribute per (vector of the
    partition in ascending vector index, attribute of that vector); name_index = position in the sorted names."""
    out = _string_field(1, partition_id)
    index = {n: i for i, n in enumerate(attribute_names)}
    for u in ids16:
        attrs = attribute_table.get(bytes(u)) if attribute_table else None
        if not attrs:
            continue
        for name, value in attrs.items():
            if name not in index:
                raise Error("InvalidContext", "attribute name must be encoded: %s" % name)
            entry = _message_field(1, _uuid_message(u)) + _uint32_field(2, index[name]) + \
                _message_field(3, attribute_value_message(value))
            out += _message_field(10, entry)
    return out


def serialize_arrays(base, coarse, codebooks, offsets, codes_pm, ids16, attribute_table=None):
    """serialize_database from plain arrays: coarse [P][N], codebooks [D][C][s], offsets [P+1],
    codes_pm [M][D] partition-major, ids16 [M][16] uint8 partition-major; attribute_table: {16 id bytes: {name:
    str | int}} (src/db/build.rs:174-175).  Returns the header id."""
    coarse = np.ascontiguousarray(coarse, np.float32)
    codebooks = np.ascontiguousarray(codebooks, np.float32)
    P, N = coarse.shape
    D, C, _ = codebooks.shape
    offsets = np.asarray(offsets, np.int64)
    partition_ids = []
    for p in range(P):
        lo, hi = int(offsets[p]), int(offsets[p + 1])
        msg = partition_message(coarse[p], np.asarray(codes_pm[lo:hi]).reshape(hi - lo, D), ids16[lo:hi])
        partition_ids.append(_persist(base, "partitions", msg, True))
    centroids_id = _persist(base, "partitions", vector_set_message(coarse), False)
    codebook_ids = [_persist(base, "codebooks", vector_set_message(codebooks[d]), False) for d in range(D)]
    # get_sorted_attribute_names (src/db/build/proto.rs:149-158): a BTreeSet, i.e. sorted by bytes
    names = sorted({n for a in (attribute_table or {}).values() for n in a}, key=lambda n: n.encode())
    log_ids = []
    for p, pid in enumerate(partition_ids):
        lo, hi = int(offsets[p]), int(offsets[p + 1])
        log_ids.append(_persist(base, "attributes", attributes_log_message(pid, ids16[lo:hi], attribute_table, names), True))
    return _persist(base, "", database_message(N, P, D, C, partition_ids, centroids_id, codebook_ids, log_ids, names), True)


def serialize_database(db, base):
    """db: flechasdb_b200.db.Database (built on the GPU).  Returns the header id; the database
    is then loadable with load_database(base, "<id>.binpb") here or by the reference."""
    coarse, _ = db.ckm.get()
    cbs, _ = db.pkm.get()
    off, order, codes = db.index.layout(order=True)
    return serialize_arrays(base, coarse[0], cbs, off, codes, db._id_bytes[order], db._attribute_table)


class StoredArrays:
    """what stored::Database holds after all lazy loads"""

    def __init__(self, N, P, D, C, coarse, codebooks, offsets, codes_pm, ids16, attribute_names=(), attribute_table=None):
        self.vector_size, self.num_partitions, self.num_divisions, self.num_codes = N, P, D, C
        self.coarse, self.codebooks, self.offsets, self.codes_pm, self.ids16 = coarse, codebooks, offsets, codes_pm, ids16
        self.attribute_names = list(attribute_names)
        self.attribute_table = attribute_table if attribute_table is not None else {}

    def get_attribute(self, vector_id, key):
        """stored::Database::get_attribute (src/db/stored.rs:118-131) once everything is loaded"""
        idb = vector_id.bytes if isinstance(vector_id, uuid.UUID) else bytes(vector_id)
        if idb not in self.attribute_table:
            raise Error("InvalidArgs", "no such vector ID: %s" % uuid.UUID(bytes=idb))
        return self.attribute_table[idb].get(key)

    def vector_id(self, partition_index, vector_index):
        return uuid.UUID(bytes=bytes(self.ids16[int(self.offsets[partition_index]) + vector_index]))


def load_database(base, path):
    """load_database + every lazy loader, with the reference's validation (src/db/stored.rs:659-880)"""
    hdr = parse(_open(base, path, True))

    def u(field):
        v = hdr.get(field)
        return int(v[0][1]) if v else 0

    def strs(field):
        return [v.decode() for _, v in hdr.get(field, [])]

    N, P, D, C = u(1), u(2), u(3), u(4)
    for name, v in (("vector_size", N), ("num_divisions", D), ("num_partitions", P), ("num_codes", C)):
        if v == 0:
            raise Error("InvalidData", "%s is zero" % name)
    if N % D:
        raise Error("InvalidData", "vector_size %d is not multiple of num_divisions %d" % (N, D))
    partition_ids, codebook_ids = strs(10), strs(12)
    if len(partition_ids) != P:
        raise Error("InvalidData", "num_partitions %d and partition_ids.len() %d do not match" % (P, len(partition_ids)))
    if len(codebook_ids) != D:
        raise Error("InvalidData", "num_divisions %d and codebook_ids.len() %d do not match" % (D, len(codebook_ids)))
    # load_partition_centroids never calls verify() (src/db/stored.rs:729-755)
    cen = _open(base, "partitions/%s.%s" % (strs(11)[0], EXT), False, verify=False)
    coarse = _parse_floats(cen, 10)
    cvs = int(parse(cen[:parse_prefix(cen, 10)]).get(1, [(0, 0)])[0][1])
    if cvs != N:
        raise Error("InvalidData", "partition centroids vector size mismatch: expected %d, got %d" % (N, cvs))
    if coarse.size != P * N:
        raise Error("InvalidData", "partition centroids data length mismatch: expected %d, got %d" % (P, coarse.size // N))
    s = N // D
    cbs = np.zeros((D, C, s), np.float32)
    for d in range(D):
        raw = _open(base, "codebooks/%s.%s" % (codebook_ids[d], EXT), False)
        data = _parse_floats(raw, 10)
        if data.size != C * s:
            raise Error("InvalidData", "codebook %d has %d elements, expected %d" % (d, data.size, C * s))
        cbs[d] = data.reshape(C, s)
    offsets = [0]
    codes, ids = [], []
    for p in range(P):
        f = parse(_open(base, "partitions/%s.%s" % (partition_ids[p], EXT), True))
        pv, pd = int(f.get(1, [(0, 0)])[0][1]), int(f.get(2, [(0, 0)])[0][1])
        if pv != N or pd != D:
            raise Error("InvalidData", "partition %d shape mismatch" % p)
        enc = f.get(11)
        data = _parse_uint32s(enc[0][1], 10) if enc else np.zeros(0, np.uint32)
        if data.size % D:
            raise Error("InvalidData", "encoded vectors of partition %d are not a multiple of %d" % (p, D))
        n_p = data.size // D
        # a code indexes table[di * num_codes + code] (src/db/stored.rs:585): the reference panics past the table
        if data.size and int(data.max()) >= C:
            raise Error("InvalidData", "partition %d holds the code %d, num_codes is %d" % (p, int(data.max()), C))
        pid = np.zeros((n_p, 16), np.uint8)
        msgs = f.get(12, [])
        if len(msgs) != n_p:
            raise Error("InvalidData", "partition %d: %d vector ids for %d vectors" % (p, len(msgs), n_p))
        for i, (_, m) in enumerate(msgs):
            g = parse(m)
            upper = struct.unpack("<Q", g[1][0][1])[0] if 1 in g else 0
            lower = struct.unpack("<Q", g[2][0][1])[0] if 2 in g else 0
            pid[i] = np.frombuffer(struct.pack(">QQ", upper, lower), np.uint8)
        codes.append(data.reshape(n_p, D))
        ids.append(pid)
        offsets.append(offsets[-1] + n_p)
    codes_pm = np.concatenate(codes) if codes else np.zeros((0, D), np.uint32)
    ids16 = np.concatenate(ids) if ids else np.zeros((0, 16), np.uint8)
    # load_attribute_table (src/db/stored.rs:173-178): every partition's log; vectors without attributes get empty maps
    log_ids, names = strs(13), strs(14)
    table = {}
    for p in range(min(P, len(log_ids))):
        for idb, name, value in load_attributes_log(base, log_ids[p], partition_ids[p], p, names):
            table.setdefault(idb, {})[name] = value
    for row in ids16:
        table.setdefault(bytes(row), {})
    return StoredArrays(N, P, D, C, coarse.reshape(P, N), cbs, np.array(offsets, np.uint64), codes_pm, ids16, names, table)


def load_header(base, path):
    """load_database proper (src/db/stored.rs:659-727): the header, the partition centroids and the codebooks; the
    partitions stay on disk.  Returns (N, P, D, C, coarse [P][N], codebooks [D][C][s], partition_ids)."""
    hdr = parse(_open(base, path, True))

    def u(field):
        v = hdr.get(field)
        return int(v[0][1]) if v else 0

    def strs(field):
        return [v.decode() for _, v in hdr.get(field, [])]

    N, P, D, C = u(1), u(2), u(3), u(4)
    for name, v in (("vector_size", N), ("num_divisions", D), ("num_partitions", P), ("num_codes", C)):
        if v == 0:
            raise Error("InvalidData", "%s is zero" % name)
    if N % D:
        raise Error("InvalidData", "vector_size %d is not multiple of num_divisions %d" % (N, D))
    partition_ids, codebook_ids = strs(10), strs(12)
    if len(partition_ids) != P:
        raise Error("InvalidData", "num_partitions %d and partition_ids.len() %d do not match" % (P, len(partition_ids)))
    if len(codebook_ids) != D:
        raise Error("InvalidData", "num_divisions %d and codebook_ids.len() %d do not match" % (D, len(codebook_ids)))
    cen = _open(base, "partitions/%s.%s" % (strs(11)[0], EXT), False, verify=False)
    coarse = _parse_floats(cen, 10)
    if coarse.size != P * N:
        raise Error("InvalidData", "partition centroids data length mismatch: expected %d, got %d" % (P, coarse.size // N))
    s = N // D
    cbs = np.zeros((D, C, s), np.float32)
    for d in range(D):
        data = _parse_floats(_open(base, "codebooks/%s.%s" % (codebook_ids[d], EXT), False), 10)
        if data.size != C * s:
            raise Error("InvalidData", "codebook %d has %d elements, expected %d" % (d, data.size, C * s))
        cbs[d] = data.reshape(C, s)
    return N, P, D, C, coarse.reshape(P, N), cbs, partition_ids, strs(13), strs(14)


def parse_attribute_value(buf):
    """AttributeValue -> str | int (src/db/proto.rs:26-37); InvalidData when neither member is present"""
    f = parse(buf)
    if 1 in f:
        return f[1][-1][1].decode()
    if 2 in f:
        return int(f[2][-1][1])
    raise Error("InvalidData", "missing attribute value")


def load_attributes_log(base, log_id, partition_id, partition_index, attribute_names):
    """the entries of one AttributesLog as (16 id bytes, name, value), validated like load_attributes_log
    (src/db/stored.rs:185-249)"""
    f = parse(_open(base, "attributes/%s.%s" % (log_id, EXT), True))
    got = f[1][0][1].decode() if 1 in f else ""
    if got != partition_id:
        raise Error("InvalidData", "inconsistent partition IDs: %s vs %s" % (got, partition_id))
    out = []
    for i, (_, m) in enumerate(f.get(10, [])):
        e = parse(m)
        ni = int(e[2][0][1]) if 2 in e else 0
        if ni >= len(attribute_names):
            raise Error("InvalidData", "attribute name index out of bounds: %d" % ni)
        if 1 not in e:
            raise Error("InvalidData", "attributes log[%d, %d]: missing vector ID" % (partition_index, i))
        if 3 not in e:
            raise Error("InvalidData", "attributes log[%d, %d]: missing value" % (partition_index, i))
        g = parse(e[1][0][1])
        upper = struct.unpack("<Q", g[1][0][1])[0] if 1 in g else 0
        lower = struct.unpack("<Q", g[2][0][1])[0] if 2 in g else 0
        out.append((struct.pack(">QQ", upper, lower), attribute_names[ni], parse_attribute_value(e[3][0][1])))
    return out


def load_partition(base, partition_id, p, N, D, C):
    """load_partition (src/db/stored.rs:800-880): codes [n][D] and vector ids [n][16] of one partition"""
    f = parse(_open(base, "partitions/%s.%s" % (partition_id, EXT), True))
    pv, pd = int(f.get(1, [(0, 0)])[0][1]), int(f.get(2, [(0, 0)])[0][1])
    if pv != N or pd != D:
        raise Error("InvalidData", "partition %d shape mismatch" % p)
    enc = f.get(11)
    data = _parse_uint32s(enc[0][1], 10) if enc else np.zeros(0, np.uint32)
    if data.size % D:
        raise Error("InvalidData", "encoded vectors of partition %d are not a multiple of %d" % (p, D))
    if data.size and int(data.max()) >= C:
        raise Error("InvalidData", "partition %d holds the code %d, num_codes is %d" % (p, int(data.max()), C))
    n_p = data.size // D
    msgs = f.get(12, [])
    if len(msgs) != n_p:
        raise Error("InvalidData", "partition %d: %d vector ids for %d vectors" % (p, len(msgs), n_p))
    pid = np.zeros((n_p, 16), np.uint8)
    for i, (_, m) in enumerate(msgs):
        g = parse(m)
        upper = struct.unpack("<Q", g[1][0][1])[0] if 1 in g else 0
        lower = struct.unpack("<Q", g[2][0][1])[0] if 2 in g else 0
        pid[i] = np.frombuffer(struct.pack(">QQ", upper, lower), np.uint8)
    return data.reshape(n_p, D), pid


class StoredDatabase:
    """stored::Database<f32, LocalFileSystem> on the GPU (src/db/stored.rs:41-57,315-389).  Like the reference it
    loads lazily: load_database reads the header, the partition centroids and the codebooks; a partition's file is
    read (and its codes uploaded, fdb_index_set_partition) when a query first probes it (get_partition,
    src/db/stored.rs:269-293)."""

    def __init__(self, ctx, base, N, P, D, C, coarse, codebooks, partition_ids, attributes_log_ids=(),
                 attribute_names=()):
        from .engine import Index
        if C > 256:
            raise Error("InvalidData", "num_codes > 256 is not supported by the u8 device layout")
        self.base, self.partition_ids = base, partition_ids
        self.vector_size, self.num_partitions, self.num_divisions, self.num_codes = N, P, D, C
        self.index = Index.create_lazy(ctx, coarse, codebooks)
        self.ids = [None] * P
        self.partition_loads = 0
        # attributes stay on the host (src/db/stored.rs:53-56): one log per partition, loaded on first use
        self.attributes_log_ids, self.attribute_names = list(attributes_log_ids), list(attribute_names)
        self.attributes_log_load_flags = [False] * P
        self.attribute_table = None

    @classmethod
    def load_database(cls, ctx, base, path):
        return cls(ctx, base, *load_header(base, path))

    def _get_partition(self, p):
        if self.ids[p] is None:
            codes, ids = load_partition(self.base, self.partition_ids[p], p, self.vector_size, self.num_divisions,
                                        self.num_codes)
            self.index.set_partition(p, codes.astype(np.uint8))
            self.ids[p] = ids
            self.partition_loads += 1

    # ---- attributes (src/db/stored.rs:108-260) ----------------------------------------------------------------
    def get_attribute(self, vector_id, key):
        """Database::get_attribute: loads every attributes log on the first call; None when the vector exists but
        has no such attribute; InvalidArgs when no vector has this id"""
        if self.attribute_table is None:
            for p in range(self.num_partitions):
                self.load_attributes_log(p)
        return self._get_attribute_internal(vector_id, key)

    def get_attribute_in_partition(self, partition_index, vector_id, key):
        self.load_attributes_log(partition_index)
        return self._get_attribute_internal(vector_id, key)

    def _get_attribute_internal(self, vector_id, key):
        idb = vector_id.bytes if isinstance(vector_id, uuid.UUID) else bytes(vector_id)
        attrs = (self.attribute_table or {}).get(idb)
        if attrs is None:
            raise Error("InvalidArgs", "no such vector ID: %s" % uuid.UUID(bytes=idb))
        return attrs.get(key)

    def load_attributes_log(self, p):
        """load_attributes_log (src/db/stored.rs:185-260): also loads the partition, whose vector ids get empty
        attribute maps so that get_attribute does not fail for a vector without attributes"""
        if self.attributes_log_load_flags[p]:
            return
        self._get_partition(p)
        if p >= len(self.attributes_log_ids):
            raise Error("InvalidData", "no attributes log for partition %d" % p)
        entries = load_attributes_log(self.base, self.attributes_log_ids[p], self.partition_ids[p], p,
                                      self.attribute_names)
        if self.attribute_table is None:
            self.attribute_table = {}
        for idb, name, value in entries:
            self.attribute_table.setdefault(idb, {})[name] = value      # the last set operation wins
        for row in self.ids[p]:
            self.attribute_table.setdefault(bytes(row), {})
        self.attributes_log_load_flags[p] = True

    def query(self, v, k, nprobe, event=lambda e: None):
        """stored::Database::query_with_events (src/db/stored.rs:331-389), QueryEvent order included"""
        from . import _capi as capi
        from .db import QueryResult
        v = np.asarray(v, np.float32).reshape(1, -1)
        event(("StartingQueryInitialization",))     # centroids and codebooks are resident since load_database
        event(("FinishedQueryInitialization",))
        event(("StartingPartitionSelection",))
        probes, _ = self.index.probe(v, nprobe, capi.QUERY_STORED)
        event(("FinishedPartitionSelection",))
        for p in probes[0]:
            event(("StartingPartitionQuery", int(p)))
            self._get_partition(int(p))
            event(("FinishedPartitionQuery", int(p)))
        p, vi, d, c = self.index.query(v, k, nprobe, capi.QUERY_STORED)
        event(("StartingResultSelection",))
        out = [QueryResult(int(p[0, i]), uuid.UUID(bytes=bytes(self.ids[int(p[0, i])][int(vi[0, i])])), int(vi[0, i]),
                           float(d[0, i]), self) for i in range(int(c[0]))]
        event(("FinishedResultSelection",))
        return out

    def query_batch(self, queries, k, nprobe):
        """batched form: loads whatever the batch probes, then one device batch"""
        from . import _capi as capi
        for p in self.index.missing_partitions(queries, nprobe, capi.QUERY_STORED):
            self._get_partition(int(p))
        return self.index.query(queries, k, nprobe, capi.QUERY_STORED)

    def close(self):
        self.index.close()
