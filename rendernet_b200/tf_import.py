"""Readers for the weight containers the reference ships its networks in, without TensorFlow:

* frozen GraphDef `.pb` — what `demo/RenderNet_converter.py:11-18` writes with `convert_variables_to_constants` and
  `RenderNet_demo.py:23-30` loads: every variable is a `Const` node carrying a TensorProto under attr "value";
* TF-1 checkpoint V2 (`<prefix>.index` + `<prefix>.data-00000-of-0000N`) — what `tf.train.Saver().save`
  (`RenderNet_Shader.py:171-200`) writes and `RenderNet_converter.py:7-8` restores;
* the npz directory of `tools/model_util.py:26-39` lives in `model_util.load_weights`.

Only the protobuf *wire format* is decoded here (varints, length-delimited fields); field numbers are those of
tensorflow/core/framework/{graph,node_def,attr_value,tensor,tensor_shape}.proto and
tensorflow/core/protobuf/tensor_bundle.proto.  The `.pb` reader is tested against GraphDefs produced by a real
protobuf encoder over the TF schema (tensorboard's compiled protos); the checkpoint reader is tested against a writer
that follows the published table format, plain and Snappy-compressed blocks (tests/test_weight_import.py) — no TF-written checkpoint exists in this
environment to pin it against, so treat it as unpinned until one is tried.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, Iterator, Tuple

import numpy as np

# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64,
           10: np.bool_, 19: np.float16, 22: np.uint32, 23: np.uint64}
_VARIABLE_SUFFIXES = ("weights", "biases", "alpha")


class TFImportError(ValueError):
    pass


# ------------------------------------------------------------------------------------------ protobuf wire format
def _varint(buf, pos: int) -> Tuple[int, int]:
    out = shift = 0
    while True:
        if pos >= len(buf):
            raise TFImportError("truncated varint")
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7
        if shift > 63:
            raise TFImportError("varint too long")


def _fields(buf) -> Iterator[Tuple[int, int, object]]:
    """Yield (field_number, wire_type, value) of one message; value is int (varint / fixed) or a memoryview."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = bytes(buf[pos:pos + 8]); pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            if pos + ln > n:
                raise TFImportError("length-delimited field runs past the end of its message")
            v = buf[pos:pos + ln]; pos += ln
        elif wt == 5:
            v = bytes(buf[pos:pos + 4]); pos += 4
        else:
            raise TFImportError(f"unsupported wire type {wt}")
        yield fno, wt, v


def _signed64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _shape(buf) -> Tuple[int, ...]:
    """TensorShapeProto: dim = 2 { size = 1 }."""
    dims = []
    for fno, wt, v in _fields(buf):
        if fno == 2 and wt == 2:
            size = 0
            for f2, w2, v2 in _fields(v):
                if f2 == 1 and w2 == 0:
                    size = _signed64(v2)
            dims.append(size)
    return tuple(dims)


def _repeated(values, wt, v, fmt, width):
    if wt == 2:      # packed
        values.extend(np.frombuffer(bytes(v), dtype=fmt).tolist())
    else:
        values.append(np.frombuffer(v if isinstance(v, bytes) else int(v).to_bytes(width, "little"), dtype=fmt)[0])


def _tensor(buf) -> np.ndarray:
    """TensorProto: dtype = 1, tensor_shape = 2, tensor_content = 4, float_val = 5, double_val = 6, int_val = 7,
    int64_val = 10, bool_val = 11, half_val = 13."""
    dtype_enum, shape, content = 0, (), None
    vals = []
    for fno, wt, v in _fields(buf):
        if fno == 1 and wt == 0:
            dtype_enum = v
        elif fno == 2 and wt == 2:
            shape = _shape(v)
        elif fno == 4 and wt == 2:
            content = bytes(v)
        elif fno == 5:
            _repeated(vals, wt, v, "<f4", 4)
        elif fno == 6:
            _repeated(vals, wt, v, "<f8", 8)
        elif fno in (7, 10, 11, 13):
            if wt == 2:
                p = 0
                while p < len(v):
                    x, p = _varint(v, p)
                    vals.append(_signed64(x))
            else:
                vals.append(_signed64(v))
    if dtype_enum not in _DTYPES:
        raise TFImportError(f"unsupported tensor dtype enum {dtype_enum}")
    dt = np.dtype(_DTYPES[dtype_enum])
    count = int(np.prod(shape)) if shape else 1
    if content is not None:
        if len(content) != count * dt.itemsize:
            raise TFImportError(f"tensor_content holds {len(content)} bytes, shape {shape} needs {count * dt.itemsize}")
        return np.frombuffer(content, dtype=dt.newbyteorder("<")).astype(dt).reshape(shape)
    if dtype_enum == 19:      # half_val carries the raw 16-bit patterns
        arr = np.asarray(vals, np.uint16).view(np.float16)
    else:
        arr = np.asarray(vals, dtype=dt)
    if arr.size == count:
        return arr.reshape(shape)
    if arr.size == 1:         # TF's "splat" encoding of a constant-filled tensor
        return np.full(shape, arr.reshape(())[()], dtype=dt)
    if arr.size == 0:
        return np.zeros(shape, dtype=dt)
    # fewer values than elements: the last one repeats (tensor_util.MakeNdarray)
    out = np.empty(count, dtype=dt)
    out[:arr.size] = arr
    out[arr.size:] = arr[-1]
    return out.reshape(shape)


def _const_node(buf):
    """NodeDef: name = 1, op = 2, attr = 5 (map entry: key = 1, value = 2 AttrValue{tensor = 8})."""
    name = op = None
    tensor = None
    for fno, wt, v in _fields(buf):
        if fno == 1 and wt == 2:
            name = bytes(v).decode("utf-8")
        elif fno == 2 and wt == 2:
            op = bytes(v).decode("utf-8")
            if op != "Const":
                return name, op, None
        elif fno == 5 and wt == 2:
            key, val = None, None
            for f2, w2, v2 in _fields(v):
                if f2 == 1 and w2 == 2:
                    key = bytes(v2)
                elif f2 == 2 and w2 == 2:
                    val = v2
            if key == b"value" and val is not None:
                for f3, w3, v3 in _fields(val):
                    if f3 == 8 and w3 == 2:
                        tensor = v3
    return name, op, tensor


def read_frozen_graph(path: str, variables_only: bool = True) -> Dict[str, np.ndarray]:
    """{node name: value} of the `Const` nodes of a frozen GraphDef.  `variables_only` keeps the nodes whose last
    path component is one the reference's layers create (`weights`, `biases`, `alpha`: tools/layer_util.py:35,142,158)
    — a frozen graph also holds hundreds of shape/stride/meshgrid constants."""
    with open(path, "rb") as f:
        buf = memoryview(f.read())
    out: Dict[str, np.ndarray] = {}
    seen_node = False
    for fno, wt, v in _fields(buf):
        if fno != 1 or wt != 2:       # GraphDef.node = 1
            continue
        seen_node = True
        name, op, tensor = _const_node(v)
        if op != "Const" or tensor is None or name is None:
            continue
        if variables_only and name.rsplit("/", 1)[-1] not in _VARIABLE_SUFFIXES:
            continue
        out[name] = _tensor(tensor)
    if not seen_node:
        raise TFImportError(f"{path}: no GraphDef nodes found")
    return out


def graph_has_node(path: str, node_name: str) -> bool:
    with open(path, "rb") as f:
        buf = memoryview(f.read())
    for fno, wt, v in _fields(buf):
        if fno == 1 and wt == 2:
            for f2, w2, v2 in _fields(v):
                if f2 == 1 and w2 == 2:
                    if bytes(v2).decode("utf-8") == node_name:
                        return True
                    break
    return False


# ------------------------------------------------------------------------------------------ checkpoint V2
_TABLE_MAGIC = 0xdb4775248b80fb57


def snappy_decompress(src) -> bytes:
    """Raw Snappy (the block format LevelDB tables may use for their blocks): varint uncompressed length, then literal /
    copy elements (tag & 3: 0 literal, 1 copy with 11-bit offset, 2 copy with 16-bit offset, 3 copy with 32-bit offset)."""
    src = bytes(src)
    n, pos = _varint(src, 0)
    out = bytearray()
    while pos < len(src):
        tag = src[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(src[pos:pos + nb], "little")
                pos += nb
            ln += 1
            if pos + ln > len(src):
                raise TFImportError("snappy: literal runs past the end of the block")
            out += src[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | src[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = int.from_bytes(src[pos:pos + 2], "little")
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(src[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise TFImportError("snappy: copy offset outside the data produced so far")
        for _ in range(ln):                       # byte-wise: copies may overlap their own output
            out.append(out[-off])
    if len(out) != n:
        raise TFImportError(f"snappy: produced {len(out)} bytes, header says {n}")
    return bytes(out)


def _table_block(buf, offset: int, size: int) -> Iterator[Tuple[bytes, bytes]]:
    """One block of the index table (the LevelDB table format TF's tensor bundle uses): prefix-compressed
    entries `shared | non_shared | value_len | key_delta | value`, then the restart array and its length; a
    1-byte compression tag and a 4-byte crc follow the block."""
    if offset + size + 5 > len(buf):
        raise TFImportError("index block runs past the end of the file")
    ctype = buf[offset + size]
    if ctype == 1:                                # kSnappyCompression
        blk = memoryview(snappy_decompress(buf[offset:offset + size]))
        size = len(blk)
    elif ctype == 0:
        blk = buf[offset:offset + size]
    else:
        raise TFImportError(f"index block compression type {ctype} is not supported")
    n_restarts = struct.unpack_from("<I", blk, size - 4)[0]
    end = size - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _varint(blk, pos)
        non_shared, pos = _varint(blk, pos)
        vlen, pos = _varint(blk, pos)
        key = key[:shared] + bytes(blk[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(blk[pos:pos + vlen])
        pos += vlen


def _bundle_entry(buf):
    """BundleEntryProto: dtype = 1, shape = 2, shard_id = 3, offset = 4, size = 5, crc32c = 6 (fixed32), slices = 7."""
    e = {"dtype": 0, "shape": (), "shard": 0, "offset": 0, "size": 0, "sliced": False}
    for fno, wt, v in _fields(memoryview(buf)):
        if fno == 1 and wt == 0:
            e["dtype"] = v
        elif fno == 2 and wt == 2:
            e["shape"] = _shape(v)
        elif fno == 3 and wt == 0:
            e["shard"] = v
        elif fno == 4 and wt == 0:
            e["offset"] = v
        elif fno == 5 and wt == 0:
            e["size"] = v
        elif fno == 7:
            e["sliced"] = True
    return e


def read_checkpoint(prefix: str, variables_only: bool = True) -> Dict[str, np.ndarray]:
    """{variable name: value} of a TF-1 V2 checkpoint (`prefix.index`, `prefix.data-XXXXX-of-YYYYY`)."""
    index_path = prefix + ".index"
    with open(index_path, "rb") as f:
        buf = memoryview(f.read())
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != _TABLE_MAGIC:
        raise TFImportError(f"{index_path}: not a tensor-bundle index (bad table magic)")
    footer = buf[len(buf) - 48:]
    pos = 0
    _, pos = _varint(footer, pos)           # metaindex handle
    _, pos = _varint(footer, pos)
    idx_off, pos = _varint(footer, pos)     # index handle
    idx_size, pos = _varint(footer, pos)
    entries = {}
    num_shards = 1
    for _, handle in _table_block(buf, idx_off, idx_size):
        hv = memoryview(handle)
        off, p = _varint(hv, 0)
        size, p = _varint(hv, p)
        for key, value in _table_block(buf, off, size):
            if key == b"":
                for fno, wt, v in _fields(memoryview(value)):   # BundleHeaderProto.num_shards = 1
                    if fno == 1 and wt == 0:
                        num_shards = v
                continue
            entries[key.decode("utf-8")] = _bundle_entry(value)
    shards = {}
    out: Dict[str, np.ndarray] = {}
    for name, e in entries.items():
        if variables_only and name.rsplit("/", 1)[-1] not in _VARIABLE_SUFFIXES:
            continue
        if e["sliced"]:
            raise TFImportError(f"{name}: partitioned variables are not supported")
        if e["dtype"] not in _DTYPES:
            raise TFImportError(f"{name}: unsupported dtype enum {e['dtype']}")
        dt = np.dtype(_DTYPES[e["dtype"]])
        count = int(np.prod(e["shape"])) if e["shape"] else 1
        if e["size"] != count * dt.itemsize:
            raise TFImportError(f"{name}: entry size {e['size']} != shape {e['shape']} x {dt.itemsize}")
        if e["shard"] not in shards:
            shards[e["shard"]] = np.memmap(f"{prefix}.data-{e['shard']:05d}-of-{num_shards:05d}", dtype=np.uint8,
                                           mode="r")
        raw = shards[e["shard"]][e["offset"]:e["offset"] + e["size"]]
        if raw.size != e["size"]:
            raise TFImportError(f"{name}: data shard too short")
        out[name] = np.frombuffer(raw.tobytes(), dtype=dt.newbyteorder("<")).astype(dt).reshape(e["shape"])
    return out


# ------------------------------------------------------------------------------------------ dispatch
def load_variables(path: str) -> Dict[str, np.ndarray]:
    """Weights from any container the reference uses: npz-dir (model_util.py:26-39), `.npz`, frozen `.pb`
    (RenderNet_demo.py:23-30) or a checkpoint prefix (RenderNet_converter.py:7-8)."""
    if os.path.isdir(path):
        from .model_util import load_weights
        w = load_weights(path)
        if not w:
            raise TFImportError(f"{path}: no *.txt.npz files")
        return w
    if path.endswith(".npz"):
        with np.load(path) as z:
            return {k: z[k] for k in z.files}
    if path.endswith(".pb"):
        return read_frozen_graph(path)
    for suffix in (".index", ".meta"):
        if path.endswith(suffix):
            path = path[:-len(suffix)]
    if os.path.exists(path + ".index"):
        return read_checkpoint(path)
    raise FileNotFoundError(f"{path}: not an npz directory, .npz, frozen .pb or checkpoint prefix")
