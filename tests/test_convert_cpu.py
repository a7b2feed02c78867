"""Converters (first step): ONNX initializers parsed from the protobuf wire format, FAISS flat /
.npy retrieval matrices, written as .rvcw.  The ONNX / FAISS files are built here byte by byte from the public
format definitions (onnx.proto3 field numbers, faiss/impl/index_write.cpp) - no real checkpoint exists offline."""
import os
import struct

import numpy as np
import pytest

from rvc_b200 import convert


def _varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _ld(fno, payload):            # length-delimited field
    return _varint((fno << 3) | 2) + _varint(len(payload)) + payload


def _vi(fno, v):                  # varint field
    return _varint(fno << 3) + _varint(v)


def _tensor(name, arr, how):
    dims = b"".join(_vi(1, d) for d in arr.shape) if how != "packed_dims" else _ld(1, b"".join(_varint(d) for d in arr.shape))
    if arr.dtype == np.float32 and how == "float_data":
        body = _ld(4, arr.astype("<f4").tobytes())
        dt = 1
    elif arr.dtype == np.int64:
        body = _ld(7, b"".join(_varint(int(v) & ((1 << 64) - 1)) for v in arr.ravel()))
        dt = 7
    elif arr.dtype == np.float16:
        body = _ld(9, arr.astype("<f2").tobytes())
        dt = 10
    else:
        body = _ld(9, arr.astype("<f4").tobytes())
        dt = 1
    return dims + _vi(2, dt) + _ld(8, name.encode()) + body


def _onnx(tensors):
    graph = _ld(1, b"") + b"".join(_ld(5, t) for t in tensors)          # a (empty) node, then the initializers
    return _vi(1, 8) + _ld(2, b"pytorch") + _ld(7, graph) + _ld(8, _ld(2, b""))   # ir_version, producer, graph, opset


def test_onnx_initializers_to_rvcw(tmp_path):
    from oracle.weights import read_rvcw   # independent reader of the container (test infrastructure)
    rng = np.random.default_rng(0)
    w_conv = rng.standard_normal((4, 3, 5)).astype(np.float32)
    w_emb = rng.standard_normal((6, 8)).astype(np.float32)
    w_half = rng.standard_normal((2, 2)).astype(np.float16)
    shape_c = np.array([1, -1, 192], np.int64)
    folded = rng.standard_normal((8, 8)).astype(np.float32)
    path = tmp_path / "m.onnx"
    path.write_bytes(_onnx([
        _tensor("enc_p.emb_phone.weight", w_emb, "float_data"),
        _tensor("dec.conv_pre.weight", w_conv, "raw"),
        _tensor("flow.half", w_half, "raw"),
        _tensor("shape_const", shape_c, "raw"),
        _tensor("onnx::MatMul_1234", folded, "packed_dims"),
    ]))
    init = convert.read_onnx_initializers(str(path))
    assert set(init) == {"enc_p.emb_phone.weight", "dec.conv_pre.weight", "flow.half", "shape_const", "onnx::MatMul_1234"}
    np.testing.assert_array_equal(init["dec.conv_pre.weight"], w_conv)
    np.testing.assert_array_equal(init["enc_p.emb_phone.weight"], w_emb)
    np.testing.assert_array_equal(init["shape_const"], shape_c)
    np.testing.assert_array_equal(init["onnx::MatMul_1234"], folded)

    out = tmp_path / "m.rvcw"
    written, unresolved = convert.onnx_to_rvcw(str(path), str(out))
    assert unresolved == ["onnx::MatMul_1234"] and "dec.conv_pre.weight" in written
    back = read_rvcw(str(out))
    np.testing.assert_array_equal(back["dec.conv_pre.weight"], w_conv)
    np.testing.assert_array_equal(back["flow.half"], w_half.astype(np.float32))
    np.testing.assert_array_equal(back["shape_const"], shape_c.astype(np.int32))
    # the caller names the constant-folded Linear weight
    written, unresolved = convert.onnx_to_rvcw(str(path), str(out), rename={"onnx::MatMul_1234": "enc_p.proj.weight"})
    assert unresolved == [] and "enc_p.proj.weight" in written
    np.testing.assert_array_equal(read_rvcw(str(out))["enc_p.proj.weight"], folded)


def test_onnx_parser_rejects_garbage(tmp_path):
    p = tmp_path / "bad.onnx"
    p.write_bytes(_ld(7, _ld(5, _vi(1, 4) + _vi(2, 1) + _ld(8, b"w") + _ld(9, b"\0" * 8))))   # 4 elements declared, 2 present
    with pytest.raises(ValueError):
        convert.read_onnx_initializers(str(p))
    p.write_bytes(b"\x3a\xff\xff\xff\xff\x0f")    # graph field longer than the file
    with pytest.raises(ValueError):
        convert.read_onnx_initializers(str(p))


def test_index_conversion(tmp_path):
    from oracle.weights import read_rvcw
    rng = np.random.default_rng(1)
    rows = rng.standard_normal((37, 8)).astype(np.float32)
    np.save(tmp_path / "total_fea.npy", rows)
    assert convert.index_to_rvcw(str(tmp_path / "total_fea.npy"), str(tmp_path / "a.rvcw")) == (37, 8)
    np.testing.assert_array_equal(read_rvcw(str(tmp_path / "a.rvcw"))["big_npy"], rows)
    # FAISS IndexFlatL2 file: fourcc, d, ntotal, 2 dummies, is_trained, metric, vector<float>
    blob = b"IxF2" + struct.pack("<iqqqBi", 8, 37, 1 << 20, 1 << 20, 1, 1) + struct.pack("<Q", rows.size) + rows.tobytes()
    (tmp_path / "flat.index").write_bytes(blob)
    assert convert.index_to_rvcw(str(tmp_path / "flat.index"), str(tmp_path / "b.rvcw")) == (37, 8)
    np.testing.assert_array_equal(read_rvcw(str(tmp_path / "b.rvcw"))["big_npy"], rows)
    (tmp_path / "bad_ivf.index").write_bytes(b"IwFl" + blob[4:])          # IVF fourcc on a flat body: must be refused
    with pytest.raises((ValueError, struct.error)):
        convert.index_to_rvcw(str(tmp_path / "bad_ivf.index"), str(tmp_path / "c.rvcw"))


def _ivf_flat_file(rows, nlist, sparse, seed=0):
    """An IndexIVFFlat file written to FAISS's published layout (faiss/impl/index_write.cpp): what upstream RVC's
    `added_IVF{nlist}_Flat_nprobe_1_*.index` holds.  Vectors are dealt to lists at random; ids keep the original order."""
    n, d = rows.shape
    rng = np.random.default_rng(seed)
    assign = rng.integers(0, nlist, n)
    if sparse:
        assign[assign == 1] = 0                                            # an empty list -> FAISS picks the sparse size table
    hdr = struct.pack("<iqqqBi", d, n, 1 << 20, 1 << 20, 1, 1)
    cent = rng.standard_normal((nlist, d)).astype(np.float32)
    quant = b"IxF2" + struct.pack("<iqqqBi", d, nlist, 1 << 20, 1 << 20, 1, 1) + struct.pack("<Q", cent.size) + cent.tobytes()
    direct_map = struct.pack("<B", 0) + struct.pack("<Q", 0)               # DirectMap::NoMap, empty array
    sizes = np.bincount(assign, minlength=nlist).astype(np.uint64)
    if sparse:
        nz = np.nonzero(sizes)[0]
        table = b"sprs" + struct.pack("<Q", 2 * len(nz)) + np.stack([nz.astype(np.uint64), sizes[nz]], 1).tobytes()
    else:
        table = b"full" + struct.pack("<Q", nlist) + sizes.tobytes()
    lists = b""
    for l in range(nlist):
        ids = np.nonzero(assign == l)[0].astype(np.int64)
        ids = ids[rng.permutation(len(ids))]
        if len(ids):
            lists += rows[ids].tobytes() + ids.tobytes()
    return b"IwFl" + hdr + struct.pack("<QQ", nlist, 1) + quant + direct_map + b"ilar" + struct.pack("<QQ", nlist, 4 * d) + table + lists


@pytest.mark.parametrize("sparse", [False, True])
def test_faiss_ivf_flat_index_is_decoded_in_id_order(tmp_path, sparse):
    from oracle.weights import read_rvcw
    rows = np.random.default_rng(4).standard_normal((203, 16)).astype(np.float32)
    path = tmp_path / "added_IVF8_Flat_nprobe_1.index"
    path.write_bytes(_ivf_flat_file(rows, 8, sparse))
    np.testing.assert_array_equal(convert.read_faiss_ivf_flat(str(path)), rows)
    assert convert.index_to_rvcw(str(path), str(tmp_path / "ivf.rvcw")) == (203, 16)
    np.testing.assert_array_equal(read_rvcw(str(tmp_path / "ivf.rvcw"))["big_npy"], rows)
    blob = bytearray(path.read_bytes())
    blob[-8:] = struct.pack("<q", 10 ** 6)                                  # an id outside [0, ntotal)
    (tmp_path / "corrupt.index").write_bytes(bytes(blob))
    with pytest.raises(ValueError):
        convert.read_faiss_ivf_flat(str(tmp_path / "corrupt.index"))


def test_converted_model_runs_through_the_engine_packer(tmp_path):
    """End to end on the CPU: the synthesizer's tensors, re-encoded here as ONNX initializers (raw_data, state_dict
    names), converted with onnx_to_rvcw, are read by the engine's own container reader + weight packer (via the
    test-only plan interpreter) and give the same window, bit for bit, as the original .rvcw."""
    import tempfile
    import planexec
    from oracle import pipeline, weights
    from oracle.weights import read_rvcw
    data = weights.make_data_dir(os.path.join(tempfile.gettempdir(), "rvc_b200_data_seed7"), seed=7, index_rows=40000)
    orig = read_rvcw(data["model"])
    onnx_path = tmp_path / "voice.onnx"
    enc = []
    for name, arr in orig.items():
        enc.append(_tensor(name, arr.astype(np.int64) if arr.dtype == np.int32 else arr, "raw"))
    onnx_path.write_bytes(_onnx(enc))
    out = tmp_path / "voice.rvcw"
    written, unresolved = convert.onnx_to_rvcw(str(onnx_path), str(out))
    assert unresolved == [] and set(written) == set(orig)
    g = pipeline.BASELINE_GEOM
    x = pipeline.synthetic_pcm(g["n16k"], seed=4)
    audio = []
    for model in (data["model"], str(out)):
        pe = planexec.PlanExec(data["data"])
        pe.load(0, data["contentvec"]); pe.load(1, data["f0"]); pe.load(2, model)
        pe.set_params(seed=0, noise_mode=1, index_k=8)
        pe.run(planexec.PLAN_INFER, x, g["sf16k"], 12, g["skip_head"], g["return_length"])
        audio.append(pe.get("audio"))
        pe.close()
    np.testing.assert_array_equal(audio[0], audio[1])
    assert float(np.sqrt(np.mean(audio[0] ** 2))) > 0.01
