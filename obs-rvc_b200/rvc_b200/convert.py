"""Offline converters into the engine's `.rvcw` container (first step).

The reference loads three opaque `.onnx` graphs (`rvc/src/models.rs:48-76`; names at the call sites
`rvc/src/rvc.rs:92-93,196-207`, `rvc/src/f0/rmvpe.rs:235-236`) and, in the plugin settings, a FAISS `.index`
(`obs-rvc/src/lib.rs:337`).  This engine keeps weights in its own arenas, so real artefacts have to be
converted once:

* `read_onnx_initializers` - the initializer tensors of an ONNX file, parsed straight from the protobuf wire
  format (no `onnx` package needed): ModelProto.graph (field 7) -> GraphProto.initializer (field 5) ->
  TensorProto {dims 1, data_type 2, float_data 4, int32_data 5, int64_data 7, name 8, raw_data 9}.
* `onnx_to_rvcw` - writes them as an `.rvcw` file under the upstream state_dict names the packer
  (`csrc/model.cpp`) looks up.  Exporters keep parameter names for conv / embedding / norm weights; Linear
  weights that `torch.onnx` constant-folds into anonymous `onnx::MatMul_*` tensors are listed as unresolved (the
  caller supplies a `rename` map for them) - there is no real checkpoint in this environment to test a graph
  tracer against, so none is pretended.
* `index_to_rvcw` - the retrieval matrix (`big_npy`) from a `.npy` file or a FAISS `IndexFlat` file
  (fourcc `IxF2` / `IxFI`: d, ntotal, two dummies, is_trained, metric, then the float vector) or a FAISS `IndexIVFFlat`
  file (`IwFl`, what upstream RVC ships: vectors recovered from the inverted lists in id order).  PQ / other IVF codecs are
  not decoded; upstream RVC saves `total_fea.npy` next to them.

Only the container layout is shared with `oracle/weights.py` (test infrastructure); nothing here imports it.
"""
import os
import struct

import numpy as np

MAGIC = b"RVCW0001"
_DT = {np.dtype("float32"): 0, np.dtype("int32"): 1}

# ONNX TensorProto.DataType -> numpy
_ONNX_DT = {1: np.float32, 6: np.int32, 7: np.int64, 10: np.float16, 11: np.float64}


def write_rvcw(path: str, tensors: dict) -> None:
    """name -> float32 / int32 array; layout: magic, (n, table bytes, data offset, data bytes), table, 64-byte
    aligned tensors (`csrc/model.cpp` RvcwFile::load)."""
    table = bytearray()
    off = 0
    entries = []
    for name, arr in tensors.items():
        arr = np.ascontiguousarray(arr)
        if arr.dtype not in _DT:
            raise TypeError(f"{name}: unsupported dtype {arr.dtype}")
        nb = name.encode()
        table += struct.pack("<H", len(nb)) + nb
        table += struct.pack("<BB", _DT[arr.dtype], arr.ndim)
        table += struct.pack(f"<{arr.ndim}I", *arr.shape)
        table += struct.pack("<QQ", off, arr.nbytes)
        entries.append((off, arr))
        off = (off + arr.nbytes + 63) & ~63
    data_offset = (32 + len(table) + 63) & ~63
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<IIQQ", len(tensors), len(table), data_offset, off))
        f.write(table)
        f.write(b"\0" * (data_offset - 32 - len(table)))
        pos = 0
        for o, arr in entries:
            f.write(b"\0" * (o - pos))
            f.write(arr.tobytes())
            pos = o + arr.nbytes
        f.write(b"\0" * (off - pos))


# ---------------------------------------------------------------------------------------------- protobuf wire


def _varint(buf, p):
    v = 0
    shift = 0
    while True:
        b = buf[p]
        p += 1
        v |= (b & 0x7F) << shift
        if not b & 0x80:
            return v, p
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def _fields(buf):
    """Yields (field number, wire type, value) of one message; value is an int (varint / fixed) or a memoryview."""
    p, n = 0, len(buf)
    while p < n:
        key, p = _varint(buf, p)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, p = _varint(buf, p)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, p)[0]
            p += 8
        elif wt == 2:
            ln, p = _varint(buf, p)
            if p + ln > n:
                raise ValueError("truncated length-delimited field")
            v = buf[p:p + ln]
            p += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, p)[0]
            p += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield fno, wt, v


def _packed_varints(buf):
    out, p = [], 0
    while p < len(buf):
        v, p = _varint(buf, p)
        out.append(v)
    return out


def _tensor(buf):
    dims, dtype, name, raw = [], 1, "", None
    floats, i32, i64 = [], [], []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            dims += _packed_varints(v) if wt == 2 else [v]
        elif fno == 2:
            dtype = v
        elif fno == 4:
            floats.append(np.frombuffer(v, "<f4") if wt == 2 else np.array([struct.unpack("<f", struct.pack("<I", v))[0]], "<f4"))
        elif fno == 5:
            i32 += _packed_varints(v) if wt == 2 else [v]
        elif fno == 7:
            i64 += _packed_varints(v) if wt == 2 else [v]
        elif fno == 8:
            name = bytes(v).decode()
        elif fno == 9:
            raw = bytes(v)
        elif fno == 13 and wt == 2:
            raise ValueError(f"tensor {name!r} keeps its data in an external file (not supported)")
    if dtype not in _ONNX_DT:
        raise ValueError(f"tensor {name!r}: unsupported ONNX data type {dtype}")
    np_dt = np.dtype(_ONNX_DT[dtype])
    if raw is not None:
        arr = np.frombuffer(raw, np_dt.newbyteorder("<"))
    elif floats:
        arr = np.concatenate(floats)
    elif i64:
        arr = np.array([v - (1 << 64) if v >= (1 << 63) else v for v in i64], np.int64)
    elif i32:
        arr = np.array([v - (1 << 64) if v >= (1 << 63) else v for v in i32], np.int64).astype(np.int32)
    else:
        arr = np.zeros(0, np_dt)
    shape = tuple(int(d) for d in dims)
    if int(np.prod(shape, dtype=np.int64)) != arr.size:
        raise ValueError(f"tensor {name!r}: {arr.size} elements for shape {shape}")
    return name, arr.reshape(shape)


def read_onnx_initializers(path: str) -> dict:
    """{initializer name: array} of an ONNX model file."""
    with open(path, "rb") as f:
        buf = memoryview(f.read())
    out = {}
    for fno, wt, v in _fields(buf):
        if fno == 7 and wt == 2:                      # ModelProto.graph
            for gno, gwt, gv in _fields(v):
                if gno == 5 and gwt == 2:             # GraphProto.initializer
                    name, arr = _tensor(gv)
                    out[name] = arr
    return out


def onnx_to_rvcw(onnx_path: str, out_path: str, rename: dict = None, strip_prefix: str = ""):
    """Writes the float / int initializers of `onnx_path` to `out_path`.  Returns (written names, unresolved names):
    tensors whose name starts with `onnx::` (constant-folded by the exporter) are unresolved unless `rename`
    maps them to a state_dict name."""
    rename = rename or {}
    tensors, unresolved = {}, []
    for name, arr in read_onnx_initializers(onnx_path).items():
        name = rename.get(name, name)
        if name.startswith("onnx::") or not name:
            unresolved.append(name)
            continue
        if strip_prefix and name.startswith(strip_prefix):
            name = name[len(strip_prefix):]
        if arr.dtype in (np.float16, np.float64):
            arr = arr.astype(np.float32)
        elif arr.dtype == np.int64:
            if arr.size and (arr.max() > 2**31 - 1 or arr.min() < -2**31):
                unresolved.append(name)
                continue
            arr = arr.astype(np.int32)
        tensors[name] = np.ascontiguousarray(arr)
    write_rvcw(out_path, tensors)
    return sorted(tensors), sorted(unresolved)


# ---------------------------------------------------------------------------------------------- retrieval index


def read_faiss_flat(path: str) -> np.ndarray:
    """Vectors of a FAISS IndexFlatL2 / IndexFlatIP file (faiss/impl/index_write.cpp: fourcc, index header, codes)."""
    with open(path, "rb") as f:
        buf = f.read()
    fourcc = buf[:4]
    if fourcc not in (b"IxF2", b"IxFI", b"IxFl"):
        raise ValueError(f"{path}: fourcc {fourcc!r} is not a flat index (IVF / PQ indices are not decoded; use total_fea.npy)")
    d, ntotal = struct.unpack_from("<iq", buf, 4)
    p = 4 + 4 + 8 + 8 + 8 + 1 + 4          # d, ntotal, 2 dummies, is_trained, metric_type
    (n_floats,) = struct.unpack_from("<Q", buf, p)
    p += 8
    if n_floats != d * ntotal or p + 4 * n_floats > len(buf):
        raise ValueError(f"{path}: inconsistent flat index header (d={d}, ntotal={ntotal}, floats={n_floats})")
    return np.frombuffer(buf, "<f4", count=n_floats, offset=p).reshape(ntotal, d).copy()


def _index_header(buf, p):
    """faiss/impl/index_write.cpp write_index_header: d, ntotal, two dummies, is_trained, metric_type (+ metric_arg)."""
    d, ntotal = struct.unpack_from("<iq", buf, p)
    p += 4 + 8 + 8 + 8
    p += 1                                   # is_trained
    (metric,) = struct.unpack_from("<i", buf, p)
    p += 4
    if metric > 1:
        p += 4                               # metric_arg
    return d, ntotal, p


def read_faiss_ivf_flat(path: str) -> np.ndarray:
    """Vectors of a FAISS IndexIVFFlat file - what upstream RVC ships as `added_IVF{n}_Flat_nprobe_1_*.index` - in their
    ORIGINAL row order (the ids of the inverted lists), i.e. the `big_npy` the exact search needs.  Layout
    (faiss/impl/index_write.cpp): fourcc `IwFl`, index header, nlist, nprobe, the coarse quantizer as a nested index
    (IndexFlat: header + centroid vector), the direct map (type byte + id array [+ hashtable pairs]), then the
    ArrayInvertedLists: fourcc `ilar`, nlist, code_size, list-size table (`full`: one size per list, `sprs`: (list, size)
    pairs), and per non-empty list its codes (size x code_size bytes) followed by its ids (size x int64).
    No real .index file exists offline: the reader is tested on files written to the same published layout."""
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:4] != b"IwFl":
        raise ValueError(f"{path}: fourcc {buf[:4]!r} is not an IndexIVFFlat file")
    d, ntotal, p = _index_header(buf, 4)
    nlist, _nprobe = struct.unpack_from("<QQ", buf, p)
    p += 16
    q4 = buf[p:p + 4]
    if q4 not in (b"IxF2", b"IxFI", b"IxFl"):
        raise ValueError(f"{path}: coarse quantizer {q4!r} is not a flat index")
    qd, _qn, p = _index_header(buf, p + 4)
    (qfloats,) = struct.unpack_from("<Q", buf, p)
    p += 8 + 4 * qfloats                     # centroids: not needed for the exact search
    dm_type = buf[p]
    p += 1
    (dm_n,) = struct.unpack_from("<Q", buf, p)
    p += 8 + 8 * dm_n
    if dm_type == 2:                         # DirectMap::Hashtable: vector of (id, lo) pairs
        (hn,) = struct.unpack_from("<Q", buf, p)
        p += 8 + 16 * hn
    if buf[p:p + 4] != b"ilar":
        raise ValueError(f"{path}: inverted lists {buf[p:p + 4]!r} are not ArrayInvertedLists")
    il_nlist, code_size = struct.unpack_from("<QQ", buf, p + 4)
    p += 4 + 16
    if il_nlist != nlist or code_size != 4 * d or qd != d:
        raise ValueError(f"{path}: inconsistent IVF header (nlist {nlist}/{il_nlist}, code_size {code_size}, d {d}/{qd})")
    kind = buf[p:p + 4]
    p += 4
    sizes = np.zeros(nlist, np.int64)
    (n,) = struct.unpack_from("<Q", buf, p)
    p += 8
    if kind == b"full":
        sizes[:] = np.frombuffer(buf, "<u8", count=n, offset=p)
        p += 8 * n
    elif kind == b"sprs":
        pairs = np.frombuffer(buf, "<u8", count=n, offset=p).reshape(-1, 2)
        sizes[pairs[:, 0].astype(np.int64)] = pairs[:, 1].astype(np.int64)
        p += 8 * n
    else:
        raise ValueError(f"{path}: unknown list-size table {kind!r}")
    if int(sizes.sum()) != ntotal:
        raise ValueError(f"{path}: list sizes sum to {int(sizes.sum())}, ntotal is {ntotal}")
    rows = np.zeros((ntotal, d), np.float32)
    seen = np.zeros(ntotal, bool)
    for sz in sizes:
        sz = int(sz)
        if sz == 0:
            continue
        codes = np.frombuffer(buf, "<f4", count=sz * d, offset=p).reshape(sz, d)
        p += sz * code_size
        ids = np.frombuffer(buf, "<i8", count=sz, offset=p)
        p += 8 * sz
        if ids.min() < 0 or ids.max() >= ntotal:
            raise ValueError(f"{path}: vector id out of range")
        rows[ids] = codes
        seen[ids] = True
    if not seen.all():
        raise ValueError(f"{path}: {int((~seen).sum())} vector ids missing from the inverted lists")
    return rows


def index_to_rvcw(src_path: str, out_path: str) -> tuple:
    """`big_npy` for `rvc_load_index` from total_fea.npy / big_npy.npy, a FAISS flat index or a FAISS IVF-Flat index.
    Returns its shape."""
    if src_path.endswith(".npy"):
        rows = np.load(src_path)
    else:
        with open(src_path, "rb") as f:
            fourcc = f.read(4)
        rows = read_faiss_ivf_flat(src_path) if fourcc == b"IwFl" else read_faiss_flat(src_path)
    rows = np.ascontiguousarray(rows, np.float32)
    if rows.ndim != 2 or rows.shape[1] % 4 != 0:
        raise ValueError(f"index rows must be [N, C] with C a multiple of 4, got {rows.shape}")
    write_rvcw(out_path, {"big_npy": rows})
    return rows.shape
