"""TensorFlow V2 ("tensor bundle") checkpoints without TensorFlow: reader, writer and the QuerySAT mapping.

The reference restores its model with ``tf.train.Checkpoint(step=..., model=model)`` +
``tf.train.CheckpointManager(ckpt, model_path).latest_checkpoint``
(``satuniformity/DiffusionSampler.py:215-227``), i.e. ``model_path`` is a directory holding

* ``checkpoint``                      text proto, ``model_checkpoint_path: "ckpt-N"``
* ``ckpt-N.index``                    an SSTable (LevelDB table format) ``tensor key -> BundleEntryProto``
* ``ckpt-N.data-00000-of-00001``      the raw little-endian tensor bytes

TensorFlow is not installable offline, so the published formats are restated here (tensor_bundle.proto,
the LevelDB table format, trackable_object_graph.proto).  Keys of an object-graph checkpoint are attribute
paths: the twelve Dense layers of the reference model are
``model/<mlp attribute>/dense_layers/<i>/{kernel,bias}/.ATTRIBUTES/VARIABLE_VALUE`` with the attributes of
``model/query_sat.py:117-122`` (``update_gate``, ``variables_output``, ``variables_query``, ``clause_mlp``,
``lit_mlp``) and the layer list of ``model/mlp.py:24,39`` (``dense_layers``).  If the flat keys are not found the
serialized object graph (key ``_CHECKPOINTABLE_OBJECT_GRAPH``) is walked instead.

Not validated against a file written by TensorFlow itself in this container (none is available); the reader is
exercised against :func:`write_checkpoint`, which follows the same published layout, and it checks the block
checksums and the table magic, so a layout mismatch fails loudly instead of loading garbage.
"""

from __future__ import annotations

import os
import re
import struct
from collections import OrderedDict

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
OBJECT_GRAPH_KEY = "_CHECKPOINTABLE_OBJECT_GRAPH"
VALUE_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"

# DataType enum of tensorflow/core/framework/types.proto
_DTYPES = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 4: np.dtype("u1"), 5: np.dtype("<i2"),
           6: np.dtype("i1"), 9: np.dtype("<i8"), 10: np.dtype("?"), 17: np.dtype("<u2"), 19: np.dtype("<f2"),
           22: np.dtype("<u4"), 23: np.dtype("<u8")}
DT_STRING, DT_BFLOAT16 = 7, 14
_DTYPE_CODES = {np.dtype("float32"): 1, np.dtype("float64"): 2, np.dtype("int32"): 3, np.dtype("int64"): 9}

# reference attribute name of each MLP (model/query_sat.py:117-122) keyed by this package's MLP names
MLP_ATTRIBUTES = OrderedDict([("variables_query", "variables_query"), ("lit_query", "lit_mlp"),
                              ("clause_update", "clause_mlp"), ("update_gate", "update_gate"),
                              ("variables_output", "variables_output")])


# ------------------------------------------------------------------------------------------ crc32c
def _make_crc_table():
    table = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        table.append(c)
    return table


_CRC_TABLE = _make_crc_table()


def crc32c(data: bytes, crc: int = 0) -> int:
    """CRC-32C (Castagnoli), the checksum of LevelDB tables and tensor bundles."""
    c = crc ^ 0xFFFFFFFF
    table = _CRC_TABLE
    for b in data:
        c = table[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(crc: int) -> int:
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------ varints / protobuf
def _read_varint(buf, pos):
    result, shift = 0, 0
    while True:
        if pos >= len(buf):
            raise ValueError("truncated varint")
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 70:
            raise ValueError("varint too long")


def _write_varint(value: int) -> bytes:
    if value < 0:
        value += 1 << 64
    out = bytearray()
    while True:
        b = value & 0x7F
        value >>= 7
        if value:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def parse_proto(buf):
    """Generic protobuf wire parse: ``[(field, wire_type, value)]``; length-delimited values stay bytes."""
    fields, pos = [], 0
    buf = bytes(buf)
    while pos < len(buf):
        tag, pos = _read_varint(buf, pos)
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            value, pos = _read_varint(buf, pos)
        elif wire == 1:
            value = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wire == 2:
            n, pos = _read_varint(buf, pos)
            value = buf[pos:pos + n]
            if len(value) != n:
                raise ValueError("truncated length-delimited field")
            pos += n
        elif wire == 5:
            value = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wire)
        fields.append((field, wire, value))
    return fields


def _pb_varint(field, value):
    return _write_varint(field << 3) + _write_varint(value)


def _pb_bytes(field, value: bytes):
    return _write_varint((field << 3) | 2) + _write_varint(len(value)) + value


def _pb_fixed32(field, value):
    return _write_varint((field << 3) | 5) + struct.pack("<I", value)


# ------------------------------------------------------------------------------------------ snappy (blocks may be compressed)
def snappy_decompress(data: bytes) -> bytes:
    n, pos = _read_varint(data, 0)
    out = bytearray()
    while pos < len(data):
        tag = data[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
            length = tag >> 2
            if length >= 60:
                extra = length - 59
                length = int.from_bytes(data[pos:pos + extra], "little")
                pos += extra
            length += 1
            out += data[pos:pos + length]
            pos += length
            continue
        if kind == 1:
            length = ((tag >> 2) & 7) + 4
            offset = ((tag >> 5) << 8) | data[pos]
            pos += 1
        elif kind == 2:
            length = (tag >> 2) + 1
            offset = int.from_bytes(data[pos:pos + 2], "little")
            pos += 2
        else:
            length = (tag >> 2) + 1
            offset = int.from_bytes(data[pos:pos + 4], "little")
            pos += 4
        if offset == 0 or offset > len(out):
            raise ValueError("corrupt snappy stream")
        for _ in range(length):                      # copies may overlap their own output
            out.append(out[-offset])
    if len(out) != n:
        raise ValueError("snappy length mismatch")
    return bytes(out)


# ------------------------------------------------------------------------------------------ SSTable
def _read_block(buf, offset, size, verify=True):
    contents = buf[offset:offset + size]
    trailer = buf[offset + size:offset + size + 5]
    if len(contents) != size or len(trailer) != 5:
        raise ValueError("table block out of range")
    if verify:
        expect = struct.unpack("<I", trailer[1:])[0]
        if mask_crc(crc32c(contents + trailer[:1])) != expect:
            raise ValueError("table block checksum mismatch")
    if trailer[0] == 1:
        contents = snappy_decompress(contents)
    elif trailer[0] != 0:
        raise ValueError("unknown block compression %d" % trailer[0])
    return contents


def _block_entries(block):
    if len(block) < 4:
        raise ValueError("table block too small")
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    if end < 0:
        raise ValueError("corrupt restart array")
    pos, key = 0, b""
    while pos < end:
        shared, pos = _read_varint(block, pos)
        non_shared, pos = _read_varint(block, pos)
        value_len, pos = _read_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + value_len]
        pos += value_len


def read_table(path, verify=True):
    """All ``(key, value)`` pairs of an SSTable file, in key order."""
    with open(path, "rb") as f:
        buf = f.read()
    if len(buf) < 48:
        raise ValueError("%s: too small for an SSTable" % path)
    footer = buf[-48:]
    if struct.unpack("<Q", footer[40:])[0] != TABLE_MAGIC:
        raise ValueError("%s: not an SSTable (bad magic)" % path)
    pos = 0
    _, pos = _read_varint(footer, pos)               # metaindex handle (unused)
    _, pos = _read_varint(footer, pos)
    index_off, pos = _read_varint(footer, pos)
    index_size, pos = _read_varint(footer, pos)
    out = OrderedDict()
    for _, handle in _block_entries(_read_block(buf, index_off, index_size, verify)):
        off, p = _read_varint(handle, 0)
        size, p = _read_varint(handle, p)
        for key, value in _block_entries(_read_block(buf, off, size, verify)):
            out[key] = value
    return out


class _TableWriter:
    """Minimal SSTable builder: prefix-compressed data blocks, one index block, footer."""

    def __init__(self, block_size=4096, restart_interval=16):
        self.buf = bytearray()
        self.block_size, self.restart_interval = block_size, restart_interval
        self.index = []                               # (last key of block, offset, size)
        self._reset()

    def _reset(self):
        self.block, self.restarts, self.count, self.last = bytearray(), [0], 0, b""

    def add(self, key: bytes, value: bytes):
        if self.count and key <= self.last:
            raise ValueError("keys must be added in strictly increasing order")
        shared = 0
        if self.count % self.restart_interval == 0:
            if self.count:
                self.restarts.append(len(self.block))
        else:
            limit = min(len(key), len(self.last))
            while shared < limit and key[shared] == self.last[shared]:
                shared += 1
        self.block += _write_varint(shared) + _write_varint(len(key) - shared) + _write_varint(len(value))
        self.block += key[shared:] + value
        self.last, self.count = key, self.count + 1
        if len(self.block) >= self.block_size:
            self._flush()

    def _emit(self, contents: bytes):
        offset = len(self.buf)
        self.buf += contents + b"\x00" + struct.pack("<I", mask_crc(crc32c(contents + b"\x00")))
        return offset, len(contents)

    def _flush(self):
        if not self.count:
            return
        contents = bytes(self.block) + b"".join(struct.pack("<I", r) for r in self.restarts)
        contents += struct.pack("<I", len(self.restarts))
        off, size = self._emit(contents)
        self.index.append((self.last, off, size))
        self._reset()

    def finish(self) -> bytes:
        self._flush()
        meta_off, meta_size = self._emit(struct.pack("<I", 0) + struct.pack("<I", 1))      # empty metaindex block
        block, restarts = bytearray(), []
        for key, off, size in self.index:             # index block: restart interval 1, no prefix sharing
            restarts.append(len(block))
            handle = _write_varint(off) + _write_varint(size)
            block += _write_varint(0) + _write_varint(len(key)) + _write_varint(len(handle)) + key + handle
        if not restarts:
            restarts = [0]
        contents = bytes(block) + b"".join(struct.pack("<I", r) for r in restarts) + struct.pack("<I", len(restarts))
        idx_off, idx_size = self._emit(contents)
        footer = _write_varint(meta_off) + _write_varint(meta_size) + _write_varint(idx_off) + _write_varint(idx_size)
        footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
        self.buf += footer
        return bytes(self.buf)


# ------------------------------------------------------------------------------------------ tensor bundle
class BundleEntry:
    __slots__ = ("dtype", "shape", "shard_id", "offset", "size", "crc32c", "sliced")

    def __init__(self):
        self.dtype, self.shape, self.shard_id, self.offset, self.size, self.crc32c, self.sliced = 0, (), 0, 0, 0, None, False


def _parse_entry(value) -> BundleEntry:
    e = BundleEntry()
    for field, _, v in parse_proto(value):
        if field == 1:
            e.dtype = v
        elif field == 2:                              # TensorShapeProto: repeated Dim dim = 2 { int64 size = 1 }
            dims = []
            for f2, _, dim in parse_proto(v):
                if f2 == 2:
                    size = 0
                    for f3, _, x in parse_proto(dim):
                        if f3 == 1:
                            size = x
                    dims.append(size)
            e.shape = tuple(dims)
        elif field == 3:
            e.shard_id = v
        elif field == 4:
            e.offset = v
        elif field == 5:
            e.size = v
        elif field == 6:
            e.crc32c = v
        elif field == 7:
            e.sliced = True
    return e


class CheckpointReader:
    """``CheckpointReader(prefix)``: ``keys()``, ``entry(key)``, ``tensor(key)`` of one V2 checkpoint."""

    def __init__(self, prefix: str, verify: bool = True):
        self.prefix = prefix
        self.verify = verify
        table = read_table(prefix + ".index", verify)
        self.num_shards = 1
        self.entries = OrderedDict()
        for key, value in table.items():
            if key == b"":                            # BundleHeaderProto {num_shards = 1, endianness = 2, version = 3}
                for field, _, v in parse_proto(value):
                    if field == 1:
                        self.num_shards = v
                    elif field == 2 and v != 0:
                        raise ValueError("big-endian checkpoints are not supported")
                continue
            self.entries[key.decode("utf-8")] = _parse_entry(value)
        self._shards = {}

    def keys(self):
        return list(self.entries)

    def entry(self, key) -> BundleEntry:
        return self.entries[key]

    def _shard(self, shard_id):
        if shard_id not in self._shards:
            path = "%s.data-%05d-of-%05d" % (self.prefix, shard_id, self.num_shards)
            with open(path, "rb") as f:
                self._shards[shard_id] = f.read()
        return self._shards[shard_id]

    def raw(self, key) -> bytes:
        e = self.entries[key]
        if e.sliced:
            raise ValueError("%s: partitioned (sliced) variables are not supported" % key)
        data = self._shard(e.shard_id)[e.offset:e.offset + e.size]
        if len(data) != e.size:
            raise ValueError("%s: data shard is truncated" % key)
        if self.verify and e.crc32c is not None and e.dtype != DT_STRING and mask_crc(crc32c(data)) != e.crc32c:
            raise ValueError("%s: tensor checksum mismatch" % key)
        return data

    def tensor(self, key):
        e = self.entries[key]
        data = self.raw(key)
        if e.dtype == DT_STRING:                      # varint lengths, 4-byte checksum of the lengths, then the bytes
            count = int(np.prod(e.shape)) if e.shape else 1
            pos, lengths = 0, []
            for _ in range(count):
                n, pos = _read_varint(data, pos)
                lengths.append(n)
            pos += 4
            out = []
            for n in lengths:
                out.append(bytes(data[pos:pos + n]))
                pos += n
            return out[0] if not e.shape else np.array(out, dtype=object).reshape(e.shape)
        if e.dtype == DT_BFLOAT16:
            bits = np.frombuffer(data, dtype="<u2").astype(np.uint32) << 16
            return bits.view(np.float32).reshape(e.shape)
        if e.dtype not in _DTYPES:
            raise ValueError("%s: unsupported dtype enum %d" % (key, e.dtype))
        return np.frombuffer(data, dtype=_DTYPES[e.dtype]).reshape(e.shape).copy()


def latest_checkpoint(directory: str):
    """Prefix named by ``<directory>/checkpoint`` (``tf.train.latest_checkpoint``), else the newest ``*.index``."""
    state = os.path.join(directory, "checkpoint")
    if os.path.isfile(state):
        with open(state, "r", encoding="utf-8") as f:
            m = re.search(r'^\s*model_checkpoint_path:\s*"(.*)"\s*$', f.read(), re.M)
        if m:
            prefix = m.group(1)
            if not os.path.isabs(prefix):
                prefix = os.path.join(directory, prefix)
            if os.path.isfile(prefix + ".index"):
                return prefix
    cands = sorted((os.path.getmtime(os.path.join(directory, f)), os.path.join(directory, f[:-len(".index")]))
                   for f in os.listdir(directory) if f.endswith(".index"))
    return cands[-1][1] if cands else None


# ------------------------------------------------------------------------------------------ object graph
def parse_object_graph(blob: bytes):
    """``[{"children": {local_name: node_id}, "attributes": {name: checkpoint_key}}]`` of a TrackableObjectGraph."""
    nodes = []
    for field, _, node in parse_proto(blob):
        if field != 1:
            continue
        children, attributes = OrderedDict(), OrderedDict()
        for f2, _, v in parse_proto(node):
            if f2 == 1:                               # ObjectReference {node_id = 1, local_name = 2}
                node_id, name = 0, ""
                for f3, _, x in parse_proto(v):
                    if f3 == 1:
                        node_id = x
                    elif f3 == 2:
                        name = x.decode("utf-8")
                children[name] = node_id
            elif f2 == 2:                             # SerializedTensor {name = 1, full_name = 2, checkpoint_key = 3}
                name, key = "", ""
                for f3, _, x in parse_proto(v):
                    if f3 == 1:
                        name = x.decode("utf-8")
                    elif f3 == 3:
                        key = x.decode("utf-8")
                attributes[name] = key
        nodes.append({"children": children, "attributes": attributes})
    return nodes


def _resolve(nodes, path):
    node = 0
    for name in path:
        node = nodes[node]["children"].get(name)
        if node is None:
            return None
    return nodes[node]["attributes"].get("VARIABLE_VALUE")


def _build_object_graph(keys):
    """Object graph whose attribute paths are the given ``a/b/c`` variable paths (writer side)."""
    nodes = [{"children": OrderedDict(), "attributes": OrderedDict()}]
    for path in keys:
        node = 0
        for name in path.split("/"):
            nxt = nodes[node]["children"].get(name)
            if nxt is None:
                nodes.append({"children": OrderedDict(), "attributes": OrderedDict()})
                nxt = len(nodes) - 1
                nodes[node]["children"][name] = nxt
            node = nxt
        nodes[node]["attributes"]["VARIABLE_VALUE"] = path + VALUE_SUFFIX
    blob = b""
    for n in nodes:
        body = b""
        for name, node_id in n["children"].items():
            body += _pb_bytes(1, _pb_varint(1, node_id) + _pb_bytes(2, name.encode("utf-8")))
        for name, key in n["attributes"].items():
            body += _pb_bytes(2, _pb_bytes(1, name.encode("utf-8")) + _pb_bytes(3, key.encode("utf-8")))
        blob += _pb_bytes(1, body)
    return blob


# ------------------------------------------------------------------------------------------ writer
def write_checkpoint(prefix: str, tensors, update_state: bool = True):
    """Write ``{variable path: ndarray}`` as a V2 checkpoint ``prefix.index`` + ``prefix.data-00000-of-00001``
    (variable paths get the ``/.ATTRIBUTES/VARIABLE_VALUE`` suffix and a matching object graph)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    items = {}
    for path, arr in tensors.items():
        arr = np.asarray(arr)                          # tobytes() is C order; 0-d arrays keep their shape
        if arr.dtype not in _DTYPE_CODES:
            raise ValueError("%s: dtype %s not supported by the writer" % (path, arr.dtype))
        items[path + VALUE_SUFFIX] = (_DTYPE_CODES[arr.dtype], arr.shape, arr.astype(arr.dtype.newbyteorder("<")).tobytes())
    graph = _build_object_graph(list(tensors))
    items[OBJECT_GRAPH_KEY] = (DT_STRING, (), _write_varint(len(graph)) +
                               struct.pack("<I", mask_crc(crc32c(_write_varint(len(graph))))) + graph)
    data = bytearray()
    table = _TableWriter()
    table.add(b"", _pb_varint(1, 1) + _pb_bytes(3, _pb_varint(1, 1)))              # header: one shard, little endian
    for key in sorted(items, key=lambda k: k.encode("utf-8")):
        dtype, shape, blob = items[key]
        shape_pb = b"".join(_pb_bytes(2, _pb_varint(1, d)) for d in shape)
        entry = _pb_varint(1, dtype) + _pb_bytes(2, shape_pb) + _pb_varint(4, len(data)) + _pb_varint(5, len(blob))
        entry += _pb_fixed32(6, mask_crc(crc32c(blob)))
        table.add(key.encode("utf-8"), entry)
        data += blob
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))
    with open(prefix + ".index", "wb") as f:
        f.write(table.finish())
    if update_state:
        name = os.path.basename(prefix)
        with open(os.path.join(os.path.dirname(os.path.abspath(prefix)), "checkpoint"), "w", encoding="utf-8") as f:
            f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (name, name))


# ------------------------------------------------------------------------------------------ QuerySAT mapping
def _layer_paths(mlp_attr, i):
    base = ["model", mlp_attr, "dense_layers", str(i)]
    return base + ["kernel"], base + ["bias"]


def load_querysat_weights(path: str, verify: bool = True):
    """``QuerySATWeights`` from a reference checkpoint directory or prefix (``DiffusionSampler.py:215-227``)."""
    from .weights import QuerySATWeights, mlp_layer_dims

    prefix = path
    if os.path.isdir(path):
        prefix = latest_checkpoint(path)
        if prefix is None:
            raise FileNotFoundError("no TensorFlow checkpoint under %s" % path)
    if prefix.endswith(".index"):
        prefix = prefix[:-len(".index")]
    if not os.path.isfile(prefix + ".index"):
        raise FileNotFoundError(prefix + ".index")
    reader = CheckpointReader(prefix, verify)
    nodes = None

    def find(parts):
        nonlocal nodes
        key = "/".join(parts) + VALUE_SUFFIX
        if key in reader.entries:
            return key
        if nodes is None:
            nodes = parse_object_graph(reader.tensor(OBJECT_GRAPH_KEY)) if OBJECT_GRAPH_KEY in reader.entries else []
        key = _resolve(nodes, parts) if nodes else None
        if key is None or key not in reader.entries:
            raise KeyError("checkpoint %s has no variable %s" % (prefix, "/".join(parts)))
        return key

    # widths from the shapes: variables_output/0 is [F, F], variables_query's last layer is [int(1.2 Q), Q]
    f = reader.entry(find(_layer_paths("variables_output", 0)[0])).shape[0]
    q = reader.entry(find(_layer_paths("variables_query", 1)[0])).shape[1]
    layers = OrderedDict()
    for mlp, dims in mlp_layer_dims(f, q).items():
        for i, (n_in, n_out) in enumerate(dims):
            kp, bp = _layer_paths(MLP_ATTRIBUTES[mlp], i)
            kernel = np.asarray(reader.tensor(find(kp)), dtype=np.float32)
            bias = np.asarray(reader.tensor(find(bp)), dtype=np.float32)
            if kernel.shape != (n_in, n_out) or bias.shape != (n_out,):
                raise ValueError("%s/%d: checkpoint shapes %r %r, expected (%d, %d)" %
                                 (mlp, i, kernel.shape, bias.shape, n_in, n_out))
            layers["%s/%d" % (mlp, i)] = (kernel, bias)
    return QuerySATWeights(layers, f, q)


def save_querysat_checkpoint(directory: str, weights, step: int = 0):
    """Write ``weights`` under the reference's variable names (``<directory>/ckpt-<step>``)."""
    tensors = OrderedDict()
    for name, (kernel, bias) in weights.layers.items():
        mlp, i = name.rsplit("/", 1)
        base = "model/%s/dense_layers/%s" % (MLP_ATTRIBUTES[mlp], i)
        tensors[base + "/kernel"] = np.asarray(kernel, dtype=np.float32)
        tensors[base + "/bias"] = np.asarray(bias, dtype=np.float32)
    tensors["step"] = np.asarray(step, dtype=np.int64)
    prefix = os.path.join(directory, "ckpt-%d" % step)
    write_checkpoint(prefix, tensors)
    return prefix


def is_tf_checkpoint(path: str) -> bool:
    if os.path.isdir(path):
        return os.path.isfile(os.path.join(path, "checkpoint")) or any(f.endswith(".index") for f in os.listdir(path))
    return os.path.isfile(path + ".index") or (path.endswith(".index") and os.path.isfile(path))
