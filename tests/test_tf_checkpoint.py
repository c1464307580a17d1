"""TensorFlow V2 checkpoint reader/writer: format round trips, checksums, the
QuerySAT variable mapping and the ``model_path`` entry of ``load_weights`` (reference
``satuniformity/DiffusionSampler.py:215-227``)."""
import os
import struct

import numpy as np
import pytest

from diffusionsat_b200 import tf_checkpoint as tfc
from diffusionsat_b200.weights import init_weights, load_weights, save_weights


def test_crc32c_known_answers():
    # RFC 3720 appendix B.4 vectors
    assert tfc.crc32c(b"123456789") == 0xE3069283
    assert tfc.crc32c(bytes(32)) == 0x8A9136AA
    assert tfc.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
    assert tfc.crc32c(bytes(range(32))) == 0x46DD794E
    # incremental == one shot, and the LevelDB mask is a bijection with a known form
    assert tfc.crc32c(b"6789", tfc.crc32c(b"12345")) == 0xE3069283
    crc = 0xE3069283
    assert tfc.mask_crc(crc) == ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def test_varint_and_proto_roundtrip():
    for v in (0, 1, 127, 128, 300, 2 ** 32 - 1, 2 ** 63 - 1):
        enc = tfc._write_varint(v)
        assert tfc._read_varint(enc, 0) == (v, len(enc))
    msg = tfc._pb_varint(1, 150) + tfc._pb_bytes(2, b"testing") + tfc._pb_fixed32(6, 0xDEADBEEF)
    assert tfc.parse_proto(msg) == [(1, 0, 150), (2, 2, b"testing"), (6, 5, 0xDEADBEEF)]
    assert tfc._pb_varint(1, 150) == bytes([0x08, 0x96, 0x01])          # the protobuf documentation's example


def test_snappy_decompress():
    # literal "abcd" + copy(offset 4, length 8) -> "abcdabcdabcd"; then a 1-byte-offset copy overlapping itself
    stream = bytes([12]) + bytes([3 << 2]) + b"abcd" + bytes([((8 - 4) << 2) | 1, 4])
    assert tfc.snappy_decompress(stream) == b"abcdabcdabcd"
    stream = bytes([6]) + bytes([0 << 2]) + b"x" + bytes([((5 - 4) << 2) | 1, 1])
    assert tfc.snappy_decompress(stream) == b"xxxxxx"
    with pytest.raises(ValueError):
        tfc.snappy_decompress(bytes([5]) + bytes([0 << 2]) + b"x")


def test_table_roundtrip_many_blocks(tmp_path):
    writer = tfc._TableWriter(block_size=256)
    items = [(("key/%05d/suffix" % i).encode(), os.urandom(i % 40)) for i in range(500)]
    for k, v in items:
        writer.add(k, v)
    path = tmp_path / "t.index"
    path.write_bytes(writer.finish())
    table = tfc.read_table(str(path))
    assert list(table.items()) == items
    assert len(writer.index) > 10                                       # really several data blocks
    # corruption is detected
    raw = bytearray(path.read_bytes())
    raw[100] ^= 0x40
    path.write_bytes(bytes(raw))
    with pytest.raises(ValueError):
        tfc.read_table(str(path))
    raw[100] ^= 0x40
    raw[-1] ^= 0x01
    path.write_bytes(bytes(raw))
    with pytest.raises(ValueError):
        tfc.read_table(str(path))
    with pytest.raises(ValueError):
        writer2 = tfc._TableWriter()
        writer2.add(b"b", b"")
        writer2.add(b"a", b"")


def test_checkpoint_roundtrip_dtypes(tmp_path):
    rng = np.random.default_rng(0)
    tensors = {"a/kernel": rng.standard_normal((5, 7)).astype(np.float32), "a/bias": np.zeros(7, np.float32),
               "step": np.asarray(42, dtype=np.int64), "b/c/d": rng.integers(-5, 5, size=(2, 3, 4)).astype(np.int32),
               "e": rng.standard_normal(3)}
    prefix = str(tmp_path / "ckpt-7")
    tfc.write_checkpoint(prefix, tensors)
    assert tfc.latest_checkpoint(str(tmp_path)) == prefix
    reader = tfc.CheckpointReader(prefix)
    assert set(reader.keys()) == {k + tfc.VALUE_SUFFIX for k in tensors} | {tfc.OBJECT_GRAPH_KEY}
    for k, v in tensors.items():
        got = reader.tensor(k + tfc.VALUE_SUFFIX)
        assert got.dtype == v.dtype and got.shape == v.shape
        np.testing.assert_array_equal(got, v)
    nodes = tfc.parse_object_graph(reader.tensor(tfc.OBJECT_GRAPH_KEY))
    assert tfc._resolve(nodes, ["b", "c", "d"]) == "b/c/d" + tfc.VALUE_SUFFIX
    assert tfc._resolve(nodes, ["b", "missing"]) is None
    # a flipped data byte fails the tensor checksum
    data_path = prefix + ".data-00000-of-00001"
    raw = bytearray(open(data_path, "rb").read())
    raw[reader.entry("a/kernel" + tfc.VALUE_SUFFIX).offset + 3] ^= 0x10
    open(data_path, "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        tfc.CheckpointReader(prefix).tensor("a/kernel" + tfc.VALUE_SUFFIX)
    tfc.CheckpointReader(prefix).tensor("a/bias" + tfc.VALUE_SUFFIX)              # the others still verify
    tfc.CheckpointReader(prefix, verify=False).tensor("a/kernel" + tfc.VALUE_SUFFIX)


def test_bfloat16_and_header_parsing(tmp_path):
    # hand-built bundle with a DT_BFLOAT16 tensor and an explicit header
    values = np.array([1.0, -2.5, 0.15625], dtype=np.float32)
    blob = (values.view(np.uint32) >> 16).astype("<u2").tobytes()
    table = tfc._TableWriter()
    table.add(b"", tfc._pb_varint(1, 1) + tfc._pb_varint(2, 0))
    entry = tfc._pb_varint(1, tfc.DT_BFLOAT16) + tfc._pb_bytes(2, tfc._pb_bytes(2, tfc._pb_varint(1, 3)))
    entry += tfc._pb_varint(4, 0) + tfc._pb_varint(5, len(blob)) + tfc._pb_fixed32(6, tfc.mask_crc(tfc.crc32c(blob)))
    table.add(b"x", entry)
    prefix = str(tmp_path / "bf")
    open(prefix + ".index", "wb").write(table.finish())
    open(prefix + ".data-00000-of-00001", "wb").write(blob)
    np.testing.assert_array_equal(tfc.CheckpointReader(prefix).tensor("x"), values)


def test_querysat_weights_through_tf_checkpoint(tmp_path):
    for f, q in ((128, 128), (64, 32)):
        w = init_weights(f, q, seed=3, bias_scale=0.1)
        d = tmp_path / ("model_%d_%d" % (f, q))
        prefix = tfc.save_querysat_checkpoint(str(d), w, step=12)
        assert os.path.basename(prefix) == "ckpt-12"
        keys = tfc.CheckpointReader(prefix, verify=False).keys()
        # the names the reference's object graph produces (model/query_sat.py:117-122, model/mlp.py:24,39)
        assert "model/lit_mlp/dense_layers/2/kernel/.ATTRIBUTES/VARIABLE_VALUE" in keys
        assert "model/clause_mlp/dense_layers/0/bias/.ATTRIBUTES/VARIABLE_VALUE" in keys
        assert "step/.ATTRIBUTES/VARIABLE_VALUE" in keys
        for path in (str(d), prefix, prefix + ".index"):
            got = load_weights(path) if path == str(d) else tfc.load_querysat_weights(path)
            assert (got.feature_maps, got.query_maps) == (f, q)
            assert list(got.layers) == list(w.layers)
            for name in w.layers:
                np.testing.assert_array_equal(got.layers[name][0], w.layers[name][0])
                np.testing.assert_array_equal(got.layers[name][1], w.layers[name][1])


def test_object_graph_fallback_when_keys_are_renamed(tmp_path):
    # same object graph, but the flat keys use different names (e.g. layer_with_weights-N aliases):
    # the loader must follow the graph's checkpoint_key attributes
    w = init_weights(64, 64, seed=5)
    tensors, alias = {}, {}
    for i, (name, (kernel, bias)) in enumerate(w.layers.items()):
        mlp, li = name.rsplit("/", 1)
        base = "model/%s/dense_layers/%s" % (tfc.MLP_ATTRIBUTES[mlp], li)
        alias[base + "/kernel"] = "model/layer_with_weights-%d/kernel" % i
        alias[base + "/bias"] = "model/layer_with_weights-%d/bias" % i
        tensors[alias[base + "/kernel"]] = kernel
        tensors[alias[base + "/bias"]] = bias
    prefix = str(tmp_path / "ckpt-1")
    tfc.write_checkpoint(prefix, tensors)
    # overwrite the object graph with one whose attribute paths are the reference's but whose keys are the aliases
    nodes = [{"children": {}, "attributes": {}}]
    for path, target in alias.items():
        node = 0
        for part in path.split("/"):
            if part not in nodes[node]["children"]:
                nodes.append({"children": {}, "attributes": {}})
                nodes[node]["children"][part] = len(nodes) - 1
            node = nodes[node]["children"][part]
        nodes[node]["attributes"]["VARIABLE_VALUE"] = target + tfc.VALUE_SUFFIX
    blob = b""
    for n in nodes:
        body = b"".join(tfc._pb_bytes(1, tfc._pb_varint(1, i) + tfc._pb_bytes(2, k.encode())) for k, i in n["children"].items())
        body += b"".join(tfc._pb_bytes(2, tfc._pb_bytes(1, k.encode()) + tfc._pb_bytes(3, v.encode())) for k, v in n["attributes"].items())
        blob += tfc._pb_bytes(1, body)
    reader = tfc.CheckpointReader(prefix)
    items = {k: (reader.entry(k).dtype, reader.entry(k).shape, reader.raw(k)) for k in reader.keys() if k != tfc.OBJECT_GRAPH_KEY}
    graph_blob = tfc._write_varint(len(blob)) + struct.pack("<I", 0) + blob
    data, table = bytearray(), tfc._TableWriter()
    table.add(b"", tfc._pb_varint(1, 1))
    items[tfc.OBJECT_GRAPH_KEY] = (tfc.DT_STRING, (), graph_blob)
    for key in sorted(items, key=lambda k: k.encode()):
        dtype, shape, raw = items[key]
        shape_pb = b"".join(tfc._pb_bytes(2, tfc._pb_varint(1, d)) for d in shape)
        table.add(key.encode(), tfc._pb_varint(1, dtype) + tfc._pb_bytes(2, shape_pb) + tfc._pb_varint(4, len(data)) +
                  tfc._pb_varint(5, len(raw)) + tfc._pb_fixed32(6, tfc.mask_crc(tfc.crc32c(raw))))
        data += raw
    open(prefix + ".index", "wb").write(table.finish())
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    got = tfc.load_querysat_weights(prefix)
    for name in w.layers:
        np.testing.assert_array_equal(got.layers[name][0], w.layers[name][0])


def test_missing_and_npz_paths(tmp_path):
    with pytest.raises(FileNotFoundError):
        load_weights(str(tmp_path / "nothing"))
    w = init_weights(64, 64, seed=1)
    save_weights(str(tmp_path / "w.npz"), w)
    got = load_weights(str(tmp_path / "w.npz"))
    np.testing.assert_array_equal(got.layers["update_gate/1"][0], w.layers["update_gate/1"][0])
    empty = tmp_path / "empty"
    empty.mkdir()
    with pytest.raises(FileNotFoundError):
        load_weights(str(empty))
