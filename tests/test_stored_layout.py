"""The reference's on-disk layout (src/db/build/proto.rs, src/io.rs, src/protos/database.proto):
wire-format known answers, round trips, integrity checks.  No GPU."""
import base64
import hashlib
import os
import struct
import zlib

import numpy as np
import pytest

from flechasdb_b200 import stored


def test_vector_set_wire_format_matches_rust_protobuf_unpacked_encoding():
    # VectorSet { vector_size: 2, data: [0,1,2,3,4,5] } (src/vector/proto.rs:61-70): field 1 varint,
    # then one (tag 0x55, 4 LE bytes) group per float -- rust-protobuf 3.2.0 `os.write_float(10, *v)`
    msg = stored.vector_set_message(np.arange(6, dtype=np.float32).reshape(3, 2))
    want = b"\x08\x02" + b"".join(b"\x55" + struct.pack("<f", v) for v in range(6))
    assert msg == want
    assert (stored._parse_floats(msg, 10) == np.arange(6, dtype=np.float32)).all()
    # a packed encoding of the same message parses to the same values
    packed = b"\x08\x02" + b"\x52" + bytes([24]) + np.arange(6, dtype="<f4").tobytes()
    assert (stored._parse_floats(packed, 10) == np.arange(6, dtype=np.float32)).all()


def test_encoded_vector_set_and_uuid_wire_format():
    enc = stored._uint32_field(1, 3) + stored._uint32s_unpacked(10, np.array([1, 2, 3, 200, 0, 255], np.uint32))
    assert enc == b"\x08\x03" + b"\x50\x01\x50\x02\x50\x03" + b"\x50\xc8\x01" + b"\x50\x00" + b"\x50\xff\x01"
    assert stored._parse_uint32s(enc, 10).tolist() == [1, 2, 3, 200, 0, 255]
    # Uuid::from_u64_pair(0xa1a2a3a4b1b2c1c2, 0xd1d2d3d4d5d6d7d8) (src/protos/mod.rs:95-103)
    u = struct.pack(">QQ", 0xa1a2a3a4b1b2c1c2, 0xd1d2d3d4d5d6d7d8)
    m = stored._uuid_message(u)
    assert m == b"\x09" + struct.pack("<Q", 0xa1a2a3a4b1b2c1c2) + b"\x11" + struct.pack("<Q", 0xd1d2d3d4d5d6d7d8)
    g = stored.parse(m)
    assert struct.unpack("<Q", g[1][0][1])[0] == 0xa1a2a3a4b1b2c1c2


def _arrays(seed=0, M=500, N=16, P=5, D=4, C=300):
    rng = np.random.default_rng(seed)
    coarse = rng.random((P, N), dtype=np.float32)
    cbs = rng.random((D, C, N // D), dtype=np.float32)
    sizes = rng.multinomial(M, np.ones(P) / P)
    sizes[2] = 0  # an empty partition
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    codes = rng.integers(0, C, (int(off[-1]), D)).astype(np.uint32)
    ids = rng.integers(0, 256, (int(off[-1]), 16)).astype(np.uint8)
    return coarse, cbs, off, codes, ids


def test_round_trip_and_content_addressing(tmp_path):
    coarse, cbs, off, codes, ids = _arrays()
    base = str(tmp_path / "db")
    h = stored.serialize_arrays(base, coarse, cbs, off, codes, ids)
    # file layout and names: URL-safe base64 (no pad) of the SHA-256 of the bytes on disk
    assert set(os.listdir(base)) == {"attributes", "codebooks", "partitions", h + ".binpb"}
    assert len(os.listdir(os.path.join(base, "codebooks"))) == cbs.shape[0]
    for sub in ("", "partitions", "codebooks", "attributes"):
        d = os.path.join(base, sub)
        for f in os.listdir(d):
            p = os.path.join(d, f)
            if os.path.isfile(p):
                digest = base64.urlsafe_b64encode(hashlib.sha256(open(p, "rb").read()).digest()).rstrip(b"=").decode()
                assert f == digest + ".binpb"
    # the header and the partitions are zlib streams, centroids and codebooks are plain
    assert zlib.decompress(open(os.path.join(base, h + ".binpb"), "rb").read())[:2] == b"\x08\x10"
    got = stored.load_database(base, h + ".binpb")
    assert (got.vector_size, got.num_partitions, got.num_divisions, got.num_codes) == (16, 5, 4, 300)
    assert (got.coarse == coarse).all() and (got.codebooks == cbs).all()
    assert (got.offsets == off).all() and (got.codes_pm == codes).all() and (got.ids16 == ids).all()


def test_integrity_and_validation_errors(tmp_path):
    coarse, cbs, off, codes, ids = _arrays(1, C=16)
    base = str(tmp_path / "db")
    h = stored.serialize_arrays(base, coarse, cbs, off, codes, ids)
    # flip one byte of a codebook file: verify() fails (src/io.rs:287-299)
    cdir = os.path.join(base, "codebooks")
    victim = os.path.join(cdir, sorted(os.listdir(cdir))[0])
    raw = bytearray(open(victim, "rb").read())
    raw[-1] ^= 1
    open(victim, "wb").write(bytes(raw))
    with pytest.raises(stored.Error) as e:
        stored.load_database(base, h + ".binpb")
    assert e.value.kind == "VerificationFailure"
    # header validation (src/db/stored.rs:671-706)
    bad = stored.database_message(16, 5, 3, 16, ["a"] * 5, "c", ["b"] * 3, [])
    hb = stored._persist(base, "", bad, True)
    with pytest.raises(stored.Error) as e:
        stored.load_database(base, hb + ".binpb")
    assert e.value.kind == "InvalidData" and "not multiple" in str(e.value)
    bad = stored.database_message(16, 5, 4, 16, ["a"] * 4, "c", ["b"] * 4, [])
    hb = stored._persist(base, "", bad, True)
    with pytest.raises(stored.Error) as e:
        stored.load_database(base, hb + ".binpb")
    assert "partition_ids.len()" in str(e.value)


# ---- the native (C++) writer / reader of flechasdb_b200/host/flechasdb_stored.hpp against this Python mirror -------
def _fnv(arr):
    h = 1469598103934665603
    for b in np.ascontiguousarray(arr).view(np.uint8).reshape(-1).tolist():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


def _tool():
    import subprocess
    from flechasdb_b200.host import build as hb
    hb.build()
    return hb.EXE_TOOL, subprocess


def test_cpp_writer_is_read_by_the_python_reader_and_back(tmp_path):
    """stored_tool (C++ serialize_arrays / read_header / read_partition, no GPU needed) and stored.py read each
    other's files: same protobuf wire bytes, zlib streams, SHA-256 names."""
    tool, subprocess = _tool()
    base = str(tmp_path / "cpp")
    N, P, D, C, M = 16, 5, 4, 20, 200
    out = subprocess.run([tool, "write", base, "5", str(N), str(P), str(D), str(C), str(M)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    h = out.stdout.strip()
    got = stored.load_database(base, h + ".binpb")          # verifies every content-addressed name on the way
    assert (got.vector_size, got.num_partitions, got.num_divisions, got.num_codes) == (N, P, D, C)
    assert got.offsets[2] == got.offsets[1]                  # the tool leaves partition 1 empty
    rd = subprocess.run([tool, "read", base, h + ".binpb"], capture_output=True, text=True)
    assert rd.returncode == 0, rd.stderr
    line = rd.stdout.strip()
    assert "N=%d P=%d D=%d C=%d M=%d" % (N, P, D, C, int(got.offsets[-1])) in line
    assert "coarse=" + _fnv(got.coarse) in line and "codebooks=" + _fnv(got.codebooks) in line
    assert "codes=" + _fnv(got.codes_pm.astype(np.uint8)) in line and "ids=" + _fnv(got.ids16) in line
    # the other direction: Python writes, the C++ reader parses the same arrays
    coarse, cbs, off, codes, ids = _arrays(3, C=16)
    base2 = str(tmp_path / "py")
    h2 = stored.serialize_arrays(base2, coarse, cbs, off, codes, ids)
    rd = subprocess.run([tool, "read", base2, h2 + ".binpb"], capture_output=True, text=True)
    assert rd.returncode == 0, rd.stderr
    line = rd.stdout.strip()
    assert "coarse=" + _fnv(coarse) in line and "codebooks=" + _fnv(cbs) in line
    assert "codes=" + _fnv(codes.astype(np.uint8)) in line and "ids=" + _fnv(ids) in line
    # a flipped byte fails the native verify() too
    cdir = os.path.join(base2, "codebooks")
    victim = os.path.join(cdir, sorted(os.listdir(cdir))[0])
    raw = bytearray(open(victim, "rb").read())
    raw[-1] ^= 1
    open(victim, "wb").write(bytes(raw))
    rd = subprocess.run([tool, "read", base2, h2 + ".binpb"], capture_output=True, text=True)
    assert rd.returncode == 1 and "VerificationFailure" in rd.stderr


def test_a_code_beyond_num_codes_is_invalid_data(tmp_path):
    """a stored partition whose codes do not fit num_codes must not reach the device tables"""
    coarse, cbs, off, codes, ids = _arrays(2, C=16)
    codes = codes.copy()
    codes[3, 1] = 16                                          # == num_codes: one past the table
    base = str(tmp_path / "bad")
    h = stored.serialize_arrays(base, coarse, cbs, off, codes, ids)
    with pytest.raises(stored.Error) as e:
        stored.load_database(base, h + ".binpb")
    assert e.value.kind == "InvalidData" and "num_codes" in str(e.value)
    tool, subprocess = _tool()
    rd = subprocess.run([tool, "read", base, h + ".binpb"], capture_output=True, text=True)
    assert rd.returncode == 1 and "holds the code" in rd.stderr


def test_attributes_log_wire_bytes_and_round_trip(tmp_path):
    """AttributeValue / OperationSetAttribute / AttributesLog (src/protos/database.proto:92-123) byte for byte against
    hand-derived answers, sorted attribute names in the header (src/db/build/proto.rs:149-158), and
    get_attribute's semantics after load (src/db/stored.rs:118-260)."""
    # oneof members are written even when they hold the default value
    assert stored.attribute_value_message("ab") == bytes([0x0A, 2]) + b"ab"
    assert stored.attribute_value_message("") == bytes([0x0A, 0])
    assert stored.attribute_value_message(300) == bytes([0x10, 0xAC, 0x02])
    assert stored.attribute_value_message(0) == bytes([0x10, 0x00])
    assert stored.parse_attribute_value(stored.attribute_value_message(0x123456789ABCDEF0)) == 0x123456789ABCDEF0   # src/db/proto.rs:63-76
    assert stored.parse_attribute_value(stored.attribute_value_message("string")) == "string"
    uid = bytes(range(1, 17))
    log = stored.attributes_log_message("pid", [uid], {uid: {"b": 7}}, ["a", "b"])
    uuid_msg = bytes([0x09]) + bytes(range(8, 0, -1)) + bytes([0x11]) + bytes(range(16, 8, -1))
    entry = bytes([0x0A, len(uuid_msg)]) + uuid_msg + bytes([0x10, 1]) + bytes([0x1A, 2, 0x10, 7])
    assert log == bytes([0x0A, 3]) + b"pid" + bytes([0x52, len(entry)]) + entry

    coarse, cbs, off, codes, ids = _arrays(4, C=16)
    table = {bytes(ids[0]): {"title": "first", "year": 2023}, bytes(ids[5]): {"year": 0, "empty": ""},
             bytes(ids[int(off[-1]) - 1]): {"title": "last"}}
    base = str(tmp_path / "attrs")
    h = stored.serialize_arrays(base, coarse, cbs, off, codes, ids, table)
    got = stored.load_database(base, h + ".binpb")
    assert got.attribute_names == ["empty", "title", "year"]
    assert got.get_attribute(bytes(ids[0]), "title") == "first" and got.get_attribute(bytes(ids[0]), "year") == 2023
    assert got.get_attribute(bytes(ids[5]), "year") == 0 and got.get_attribute(bytes(ids[5]), "empty") == ""
    assert got.get_attribute(bytes(ids[5]), "title") is None
    assert got.get_attribute(bytes(ids[1]), "title") is None          # a vector without attributes exists
    with pytest.raises(stored.Error) as e:
        got.get_attribute(bytes(16), "title")
    assert e.value.kind == "InvalidArgs" and "no such vector ID" in str(e.value)
    # a log that names another partition is InvalidData (src/db/stored.rs:196-202)
    hdr = stored.parse(stored._open(base, h + ".binpb", True))
    pids = [v.decode() for _, v in hdr[10]]
    lids = [v.decode() for _, v in hdr[13]]
    with pytest.raises(stored.Error) as e:
        stored.load_attributes_log(base, lids[0], pids[1], 1, got.attribute_names)
    assert e.value.kind == "InvalidData" and "inconsistent partition IDs" in str(e.value)
    with pytest.raises(stored.Error) as e:                             # name index past the header's list
        stored.load_attributes_log(base, lids[0], pids[0], 0, [])
    assert e.value.kind == "InvalidData" and "out of bounds" in str(e.value)


def test_attributes_cross_the_language_border(tmp_path):
    """C++ writes attributes logs that the Python reader understands and vice versa (same wire bytes, sorted names,
    entries in ascending vector index per partition)."""
    tool, subprocess = _tool()
    base = str(tmp_path / "cpp")
    N, P, D, C, M = 16, 4, 4, 20, 120
    out = subprocess.run([tool, "write", base, "9", str(N), str(P), str(D), str(C), str(M), "attrs"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    got = stored.load_database(base, out.stdout.strip() + ".binpb")
    assert got.attribute_names == ["idx", "name"]
    total = int(got.offsets[-1])
    for i in range(total):
        idb = bytes(got.ids16[i])
        assert got.get_attribute(idb, "idx") == (i if i % 7 == 0 else None)
        assert got.get_attribute(idb, "name") == ("v%d" % i if i % 5 == 0 else None)
    # Python writes, C++ reads: the tool hashes (id, name, value) of every entry in file order
    coarse, cbs, off, codes, ids = _arrays(3, C=16)
    table = {}
    for i in range(0, len(ids), 3):
        table[bytes(ids[i])] = {"zeta": "z%d" % i, "alpha": i * 1000003}
    base2 = str(tmp_path / "py")
    h2 = stored.serialize_arrays(base2, coarse, cbs, off, codes, ids, table)
    rd = subprocess.run([tool, "read", base2, h2 + ".binpb"], capture_output=True, text=True)
    assert rd.returncode == 0, rd.stderr
    want, n = 1469598103934665603, 0
    for i in range(len(ids)):
        for name, value in table.get(bytes(ids[i]), {}).items():
            v = ("s:" + value) if isinstance(value, str) else ("u:%d" % value)
            for b in bytes(ids[i]) + name.encode() + v.encode():
                want = ((want ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
            n += 1
    line = rd.stdout.strip()
    assert "attr_entries=%d attr_names=alpha,zeta attrs=%016x" % (n, want) in line, line
