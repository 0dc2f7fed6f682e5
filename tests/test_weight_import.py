"""Weight containers of the reference: frozen GraphDef .pb (demo/RenderNet_converter.py:11-18,
RenderNet_demo.py:23-30), TF-1 checkpoint V2 (RenderNet_converter.py:7-8) and the npz directory
(tools/model_util.py:26-39), read without TensorFlow by rendernet_b200/tf_import.py.

The .pb fixtures are produced by a real protobuf encoder over TensorFlow's own schema (tensorboard ships the compiled
graph/tensor protos); the checkpoint fixtures by the small table writer below, which follows the published
LevelDB-table / tensor-bundle layout independently of the reader's code path."""
import os
import struct

import numpy as np
import pytest

from rendernet_b200 import tf_import
from rendernet_b200.RenderNet_demo import load_graph


def _variables(seed=0, grey=False):
    rng = np.random.default_rng(seed)
    return {
        "encoder/e_conv1/e_conv1/weights": rng.standard_normal((5, 5, 5, 1, 8)).astype(np.float32),
        "encoder/e_conv1/e_conv1/biases": np.full((8,), 0.001, np.float32),
        "encoder/e_conv1/alpha": rng.uniform(0, 0.3, (8,)).astype(np.float32),
        "encoder/res2_3/con1_3X3/weights": rng.standard_normal((3, 3, 16, 16)).astype(np.float32),
        "encoder/e_conv11/weights": rng.standard_normal((4, 4, 1 if grey else 3, 16)).astype(np.float32),
        "encoder/e_conv11/biases": np.zeros((1 if grey else 3,), np.float32),
    }


# ---------------------------------------------------------------------------------------------- frozen graph
def _graph_def(variables):
    gp = pytest.importorskip("tensorboard.compat.proto.graph_pb2")
    from tensorboard.compat.proto import tensor_pb2, tensor_shape_pb2, types_pb2
    g = gp.GraphDef()

    def const(name, arr, how="content"):
        n = g.node.add()
        n.name, n.op = name, "Const"
        n.attr["dtype"].type = types_pb2.DT_FLOAT if arr.dtype == np.float32 else types_pb2.DT_INT32
        t = tensor_pb2.TensorProto(dtype=n.attr["dtype"].type,
                                   tensor_shape=tensor_shape_pb2.TensorShapeProto(
                                       dim=[tensor_shape_pb2.TensorShapeProto.Dim(size=int(s)) for s in arr.shape]))
        if how == "content":
            t.tensor_content = arr.tobytes()
        elif how == "vals":
            (t.float_val if arr.dtype == np.float32 else t.int_val).extend(arr.ravel().tolist())
        elif how == "splat":
            t.float_val.append(float(arr.ravel()[0]))
        n.attr["value"].tensor.CopyFrom(t)
        return n

    ph = g.node.add()
    ph.name, ph.op = "real_model_in", "Placeholder"
    for i, (k, v) in enumerate(variables.items()):
        how = "splat" if k.endswith("e_conv1/e_conv1/biases") else ("vals" if i % 2 else "content")
        const(k, v, how)
        rd = g.node.add()                      # convert_variables_to_constants leaves the Identity "read" nodes
        rd.name, rd.op = k + "/read", "Identity"
        rd.input.append(k)
    const("encoder/e_conv1/e_conv1/strides", np.array([1, 2, 2, 2, 1], np.int32), "vals")
    const("encoder/Reshape/shape", np.array([-1, 64, 64, 1024], np.int32), "vals")
    out = g.node.add()
    out.name, out.op = "encoder/output", "Sigmoid"
    return g


def test_frozen_graph_reader_roundtrip(tmp_path):
    variables = _variables()
    g = _graph_def(variables)
    path = tmp_path / "3d2d_render.pb"
    path.write_bytes(g.SerializeToString())
    got = tf_import.read_frozen_graph(str(path))
    assert set(got) == set(variables)                      # shape/stride constants filtered out
    for k, v in variables.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape
        np.testing.assert_array_equal(got[k], v)
    everything = tf_import.read_frozen_graph(str(path), variables_only=False)
    np.testing.assert_array_equal(everything["encoder/Reshape/shape"], [-1, 64, 64, 1024])   # negative varints
    assert tf_import.graph_has_node(str(path), "encoder/output")
    assert not tf_import.graph_has_node(str(path), "decoder/output")
    graph = load_graph(str(path))
    assert graph.is_greyscale is False and set(graph.weights) == set(variables)


def test_frozen_graph_greyscale_detected_and_garbage_rejected(tmp_path):
    path = tmp_path / "grey.pb"
    path.write_bytes(_graph_def(_variables(grey=True)).SerializeToString())
    assert load_graph(str(path)).is_greyscale is True
    bad = tmp_path / "bad.pb"
    bad.write_bytes(b"\x0a\xff\xff\xff\xff\x0f not a graph")
    with pytest.raises(tf_import.TFImportError):
        tf_import.read_frozen_graph(str(bad))
    empty = tmp_path / "empty.pb"
    empty.write_bytes(b"")
    with pytest.raises(tf_import.TFImportError):
        tf_import.read_frozen_graph(str(empty))
    unrelated = tmp_path / "other.pb"
    unrelated.write_bytes(_graph_def({"encoder/x/weights": np.ones((2, 2), np.float32)}).SerializeToString())
    with pytest.raises(ValueError, match="no RenderNet variables"):
        load_graph(str(unrelated))


# ---------------------------------------------------------------------------------------------- checkpoint V2
def _vi(n):
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def _pb_field(fno, wt, payload):
    return _vi((fno << 3) | wt) + payload


def _block(items, restart_interval=16):
    """LevelDB table block with prefix compression."""
    out, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        out += _vi(shared) + _vi(len(k) - shared) + _vi(len(v)) + k[shared:] + v
        prev = k
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _snappy(raw: bytes) -> bytes:
    """A valid (if naive) Snappy stream: runs of a repeated byte become a literal + overlapping copy, the rest literals."""
    out = bytearray(_vi(len(raw)))
    i = 0
    while i < len(raw):
        j = i
        while j < len(raw) and raw[j] == raw[i]:
            j += 1
        run = j - i
        if run >= 8:
            out += bytes([0 << 2 | 0, raw[i]])                    # 1-byte literal
            left = run - 1
            while left > 0:
                ln = min(left, 64)
                if ln < 4 and ln != left:
                    ln = left
                out += bytes([((ln - 1) << 2) | 2, 1, 0])          # copy, 2-byte offset = 1 (overlapping)
                left -= ln
            i = j
            continue
        k = i
        while k < len(raw) and k - i < 60:
            m = k
            while m < len(raw) and raw[m] == raw[k]:
                m += 1
            if m - k >= 8:
                break
            k = m if m - k < 8 else k
            if k == i:
                k += 1
        k = max(k, i + 1)
        k = min(k, i + 60)
        out += bytes([((k - i - 1) << 2) | 0]) + raw[i:k]
        i = k
    return bytes(out)


def _write_checkpoint(prefix, variables, entries_per_block=4, snappy=False):
    names = sorted(variables)
    data = bytearray()
    items = [(b"", _pb_field(1, 0, _vi(1)) + _pb_field(3, 2, _vi(2) + _pb_field(1, 0, _vi(1))))]  # header: 1 shard
    for name in names:
        arr = variables[name]
        shape = b"".join(_pb_field(2, 2, _vi(len(d)) + d) for d in (_pb_field(1, 0, _vi(s)) for s in arr.shape))
        dtype = {np.dtype(np.float32): 1, np.dtype(np.int32): 3}[arr.dtype]
        entry = _pb_field(1, 0, _vi(dtype)) + _pb_field(2, 2, _vi(len(shape)) + shape)
        if len(data):
            entry += _pb_field(4, 0, _vi(len(data)))
        entry += _pb_field(5, 0, _vi(arr.nbytes)) + _pb_field(6, 5, struct.pack("<I", 0xdeadbeef))
        items.append((name.encode(), entry))
        data += arr.tobytes()
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(data)
    table, index_items = bytearray(), []
    for i in range(0, len(items), entries_per_block):
        chunk = items[i:i + entries_per_block]
        blk = _block(chunk, restart_interval=3)
        if snappy:
            blk = _snappy(blk)
        index_items.append((chunk[-1][0] + b"~", _vi(len(table)) + _vi(len(blk))))
        table += blk + (b"\x01" if snappy else b"\x00") + struct.pack("<I", 0)
    meta = _block([])
    meta_handle = _vi(len(table)) + _vi(len(meta))
    table += meta + b"\x00" + struct.pack("<I", 0)
    idx = _block(index_items, restart_interval=1)
    idx_handle = _vi(len(table)) + _vi(len(idx))
    table += idx + b"\x00" + struct.pack("<I", 0)
    footer = meta_handle + idx_handle
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", 0xdb4775248b80fb57)
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(table) + footer)


def test_checkpoint_reader_roundtrip(tmp_path):
    variables = _variables(3)
    variables["global_step"] = np.array(7, np.int32).reshape(())
    variables["encoder/e_conv1/e_conv1/weights/Adam"] = np.ones((5, 5, 5, 1, 8), np.float32)   # optimizer slots
    prefix = str(tmp_path / "3d2d_renderer")
    _write_checkpoint(prefix, variables)
    got = tf_import.read_checkpoint(prefix)
    want = {k: v for k, v in variables.items() if k.rsplit("/", 1)[-1] in ("weights", "biases", "alpha")}
    assert set(got) == set(want)
    for k, v in want.items():
        np.testing.assert_array_equal(got[k], v)
    allv = tf_import.read_checkpoint(prefix, variables_only=False)
    assert int(allv["global_step"]) == 7 and "encoder/e_conv1/e_conv1/weights/Adam" in allv
    for spelled in (prefix, prefix + ".index", prefix + ".meta"):
        assert set(tf_import.load_variables(spelled)) == set(want)
    with open(prefix + ".index", "r+b") as f:
        f.seek(-1, os.SEEK_END)
        f.write(b"\x00")
    with pytest.raises(tf_import.TFImportError, match="magic"):
        tf_import.read_checkpoint(prefix)


def test_checkpoint_reader_snappy_blocks(tmp_path):
    """LevelDB tables may Snappy-compress their blocks; the reader carries its own decompressor."""
    raw = bytes(range(50)) + b"\x00" * 300 + b"abcabcabc" + b"\x07" * 9 + bytes(range(200, 256))
    assert tf_import.snappy_decompress(_snappy(raw)) == raw
    assert tf_import.snappy_decompress(bytes([11, (5 - 1) << 2]) + b"hello" + bytes([((6 - 4) << 2) | 1 | (0 << 5), 5])) \
        == b"hellohelloh"                                       # copy with 11-bit offset, overlapping its own output
    with pytest.raises(tf_import.TFImportError):
        tf_import.snappy_decompress(bytes([4, (2 - 1) << 2]) + b"ab" + bytes([((4 - 4) << 2) | 1, 9]))   # offset too far
    variables = _variables(8)
    prefix = str(tmp_path / "ckpt")
    _write_checkpoint(prefix, variables, snappy=True)
    got = tf_import.read_checkpoint(prefix)
    assert set(got) == set(variables)
    for k, v in variables.items():
        np.testing.assert_array_equal(got[k], v)


def test_npz_dir_and_npz_file(tmp_path):
    variables = _variables(5)
    d = tmp_path / "weights"
    d.mkdir()
    for k, v in variables.items():
        np.savez(str(d / (k[len("encoder/"):].replace("/", "_") + ".txt.npz")), v)   # model_util.py:32-38 spelling
    got = tf_import.load_variables(str(d))
    assert set(got) == {k[len("encoder/"):].replace("/", "_") for k in variables}
    np.testing.assert_array_equal(got["e_conv1_alpha"], variables["encoder/e_conv1/alpha"])
    assert load_graph(str(d)).is_greyscale is False
    f = tmp_path / "w.npz"
    np.savez(str(f), **variables)
    assert set(tf_import.load_variables(str(f))) == set(variables)
    with pytest.raises(FileNotFoundError):
        tf_import.load_variables(str(tmp_path / "missing"))
    (tmp_path / "emptydir").mkdir()
    with pytest.raises(tf_import.TFImportError):
        tf_import.load_variables(str(tmp_path / "emptydir"))
