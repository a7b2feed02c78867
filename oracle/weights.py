"""RVCW weight container I/O and seeded synthetic weights (TEST INFRASTRUCTURE).

The reference loads opaque .onnx graphs (rvc/src/models.rs:48-76) and ships no weights
(SURVEY.md section 0.2); the engine loads the same tensors from an `.rvcw` container whose
tensor names follow the upstream state_dict names the .onnx files were exported from.  The
container is trivial on purpose (an ONNX-initialiser importer is "next" row #3):

    0   char  magic[8] = "RVCW0001"
    8   u32   n_tensors
    12  u32   table_bytes
    16  u64   data_offset  (absolute, 64-byte aligned)
    24  u64   data_bytes
    32  table: { u16 name_len; char name[]; u8 dtype(0=f32,1=i32); u8 ndim; u32 dims[ndim];
                 u64 offset (relative to data_offset, 64-aligned); u64 nbytes } * n_tensors

Synthetic weights are drawn per tensor from numpy PCG64 seeded by (seed, crc32(name)) so that a
file can be regenerated bit-identically anywhere (the GPU box regenerates them; nothing large is
committed).  Gains are chosen so activations stay O(1) through the un-normalised stacks.
"""
from __future__ import annotations

import os
import struct
import zlib

import numpy as np

MAGIC = b"RVCW0001"
_DT = {np.dtype(np.float32): 0, np.dtype(np.int32): 1}
_DT_INV = {0: np.float32, 1: np.int32}


def write_rvcw(path: str, tensors: dict) -> None:
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
            if o > pos:
                f.write(b"\0" * (o - pos))
            f.write(arr.tobytes())
            pos = o + arr.nbytes
        if off > pos:
            f.write(b"\0" * (off - pos))


def read_rvcw(path: str) -> dict:
    with open(path, "rb") as f:
        head = f.read(32)
        if head[:8] != MAGIC:
            raise ValueError("not an RVCW file")
        n, tb, data_off, _ = struct.unpack("<IIQQ", head[8:])
        table = f.read(tb)
        out = {}
        p = 0
        metas = []
        for _ in range(n):
            (nl,) = struct.unpack_from("<H", table, p); p += 2
            name = table[p:p + nl].decode(); p += nl
            dt, nd = struct.unpack_from("<BB", table, p); p += 2
            dims = struct.unpack_from(f"<{nd}I", table, p); p += 4 * nd
            off, nb = struct.unpack_from("<QQ", table, p); p += 16
            metas.append((name, dt, dims, off, nb))
        for name, dt, dims, off, nb in metas:
            f.seek(data_off + off)
            out[name] = np.frombuffer(f.read(nb), dtype=_DT_INV[dt]).reshape(dims).copy()
    return out


# ----------------------------------------------------------------------------------------------
# synthetic weights
# ----------------------------------------------------------------------------------------------


class _Gen:
    def __init__(self, seed: int):
        self.seed = seed
        self.t = {}

    def rng(self, name):
        return np.random.Generator(np.random.PCG64([self.seed, zlib.crc32(name.encode())]))

    def normal(self, name, shape, std, mean=0.0):
        a = self.rng(name).standard_normal(shape, dtype=np.float32) * np.float32(std)
        if mean:
            a += np.float32(mean)
        self.t[name] = a.astype(np.float32)
        return self.t[name]

    def uniform(self, name, shape, lo, hi):
        a = self.rng(name).random(shape, dtype=np.float32) * np.float32(hi - lo) + np.float32(lo)
        self.t[name] = a.astype(np.float32)
        return self.t[name]

    def norm_affine(self, prefix, c):
        self.normal(prefix + ".weight", (c,), 0.1, 1.0)
        self.normal(prefix + ".bias", (c,), 0.1)

    def meta(self, name, *vals):
        self.t[name] = np.asarray(vals, dtype=np.int32)


HUBERT_KERNELS = (10, 3, 3, 3, 3, 2, 2)
HUBERT_STRIDES = (5, 2, 2, 2, 2, 2, 2)


def synth_contentvec(seed: int = 1234, n_layers: int = 12, final_proj: bool = False) -> dict:
    """HuBERT-base / ContentVec; names follow fairseq's state_dict.
    Weight-norm on pos_conv is stored resolved (`encoder.pos_conv.0.weight`)."""
    g = _Gen(seed)
    cin = 1
    for i, k in enumerate(HUBERT_KERNELS):
        g.normal(f"feature_extractor.conv_layers.{i}.0.weight", (512, cin, k),
                 1.5 * np.sqrt(1.0 / (cin * k)))
        cin = 512
    g.norm_affine("feature_extractor.conv_layers.0.2", 512)
    g.norm_affine("layer_norm", 512)
    g.normal("post_extract_proj.weight", (768, 512), np.sqrt(1.0 / 512))
    g.normal("post_extract_proj.bias", (768,), 0.02)
    g.normal("encoder.pos_conv.0.weight", (768, 48, 128), np.sqrt(1.0 / (48 * 128)))
    g.normal("encoder.pos_conv.0.bias", (768,), 0.02)
    g.norm_affine("encoder.layer_norm", 768)
    for i in range(n_layers):
        p = f"encoder.layers.{i}."
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            g.normal(p + f"self_attn.{nm}.weight", (768, 768), np.sqrt(1.0 / 768))
            g.normal(p + f"self_attn.{nm}.bias", (768,), 0.02)
        g.norm_affine(p + "self_attn_layer_norm", 768)
        g.normal(p + "fc1.weight", (3072, 768), np.sqrt(1.0 / 768))
        g.normal(p + "fc1.bias", (3072,), 0.02)
        g.normal(p + "fc2.weight", (768, 3072), np.sqrt(1.0 / 3072))
        g.normal(p + "fc2.bias", (768,), 0.02)
        g.norm_affine(p + "final_layer_norm", 768)
    if final_proj:
        g.normal("final_proj.weight", (256, 768), np.sqrt(1.0 / 768))
        g.normal("final_proj.bias", (256,), 0.02)
    g.meta("meta.n_layers", n_layers)
    g.meta("meta.out_dim", 256 if final_proj else 768)
    return g.t


def _rmvpe_convblockres(g: _Gen, p: str, cin: int, cout: int):
    g.normal(p + "conv.0.weight", (cout, cin, 3, 3), np.sqrt(2.0 / (9 * cin)))
    _bn(g, p + "conv.1", cout, 1.0)
    g.normal(p + "conv.3.weight", (cout, cout, 3, 3), np.sqrt(2.0 / (9 * cout)))
    _bn(g, p + "conv.4", cout, 0.25)
    if cin != cout:
        g.normal(p + "shortcut.weight", (cout, cin, 1, 1), np.sqrt(1.0 / cin))
        g.normal(p + "shortcut.bias", (cout,), 0.02)


def _bn(g: _Gen, p: str, c: int, gain: float):
    g.uniform(p + ".weight", (c,), 0.8 * gain, 1.2 * gain)
    g.normal(p + ".bias", (c,), 0.1)
    g.normal(p + ".running_mean", (c,), 0.1)
    g.uniform(p + ".running_var", (c,), 0.8, 1.2)


def synth_rmvpe(seed: int = 2345) -> dict:
    """RMVPE E2E(4, 1, (2,2)); names follow the upstream state_dict."""
    g = _Gen(seed)
    # input BN on the log-mel (values ~ [-11.5, 3])
    g.t["unet.encoder.bn.weight"] = np.asarray([1.0], np.float32)
    g.t["unet.encoder.bn.bias"] = np.asarray([0.0], np.float32)
    g.t["unet.encoder.bn.running_mean"] = np.asarray([-5.0], np.float32)
    g.t["unet.encoder.bn.running_var"] = np.asarray([9.0], np.float32)
    cin, cout = 1, 16
    for i in range(5):
        for j in range(4):
            _rmvpe_convblockres(g, f"unet.encoder.layers.{i}.conv.{j}.", cin if j == 0 else cout,
                                cout)
        cin, cout = cout, cout * 2
    # intermediate: 256 -> 512, then 512 -> 512
    cin, cout = 256, 512
    for i in range(4):
        for j in range(4):
            _rmvpe_convblockres(g, f"unet.intermediate.layers.{i}.conv.{j}.",
                                cin if (i == 0 and j == 0) else cout, cout)
    cin = 512
    for i in range(5):
        cout = cin // 2
        p = f"unet.decoder.layers.{i}."
        g.normal(p + "conv1.0.weight", (cin, cout, 3, 3), np.sqrt(2.0 / (9 * cin / 4)))
        _bn(g, p + "conv1.1", cout, 1.0)
        for j in range(4):
            _rmvpe_convblockres(g, p + f"conv2.{j}.", cout * 2 if j == 0 else cout, cout)
        cin = cout
    g.normal("cnn.weight", (3, 16, 3, 3), 0.25 * np.sqrt(1.0 / (9 * 16)))
    g.normal("cnn.bias", (3,), 0.02)
    k = 1.0 / np.sqrt(256.0)
    for sfx in ("", "_reverse"):
        g.uniform(f"fc.0.gru.weight_ih_l0{sfx}", (768, 384), -k, k)
        g.uniform(f"fc.0.gru.weight_hh_l0{sfx}", (768, 256), -k, k)
        g.uniform(f"fc.0.gru.bias_ih_l0{sfx}", (768,), -k, k)
        g.uniform(f"fc.0.gru.bias_hh_l0{sfx}", (768,), -k, k)
    g.normal("fc.1.weight", (360, 512), 0.09)
    g.normal("fc.1.bias", (360,), 0.3, -5.2)
    return g.t


def synth_voice(seed: int = 3456, sr: int = 40000, phone_dim: int = 768) -> dict:
    """SynthesizerTrnMs768NSFsid, 40k config; upstream state_dict names,
    weight-norm resolved.  Speaker id 0 is baked in like the reference's exported graph."""
    assert sr in (32000, 40000, 48000), "generator configs: 32k / 40k / 48k (oracle/nets.py GEN_CONFIGS)"
    g = _Gen(seed)
    H = 192
    g.normal("enc_p.emb_phone.weight", (H, phone_dim), np.sqrt(1.0 / phone_dim) / np.sqrt(H) * 2.0)
    g.normal("enc_p.emb_phone.bias", (H,), 0.01)
    g.normal("enc_p.emb_pitch.weight", (256, H), 1.0 / np.sqrt(H))
    for i in range(6):
        p = f"enc_p.encoder.attn_layers.{i}."
        for nm in ("conv_q", "conv_k", "conv_v", "conv_o"):
            g.normal(p + nm + ".weight", (H, H, 1), np.sqrt(1.0 / H))
            g.normal(p + nm + ".bias", (H,), 0.02)
        g.normal(p + "emb_rel_k", (1, 21, 96), 96 ** -0.5)
        g.normal(p + "emb_rel_v", (1, 21, 96), 96 ** -0.5)
        g.normal(f"enc_p.encoder.norm_layers_1.{i}.gamma", (H,), 0.1, 1.0)
        g.normal(f"enc_p.encoder.norm_layers_1.{i}.beta", (H,), 0.1)
        p = f"enc_p.encoder.ffn_layers.{i}."
        g.normal(p + "conv_1.weight", (768, H, 3), np.sqrt(2.0 / (3 * H)))
        g.normal(p + "conv_1.bias", (768,), 0.02)
        g.normal(p + "conv_2.weight", (H, 768, 3), np.sqrt(1.0 / (3 * 768)))
        g.normal(p + "conv_2.bias", (H,), 0.02)
        g.normal(f"enc_p.encoder.norm_layers_2.{i}.gamma", (H,), 0.1, 1.0)
        g.normal(f"enc_p.encoder.norm_layers_2.{i}.beta", (H,), 0.1)
    g.normal("enc_p.proj.weight", (2 * H, H, 1), np.sqrt(1.0 / H) * 0.5)
    g.normal("enc_p.proj.bias", (2 * H,), 0.02)
    for f in range(4):
        p = f"flow.flows.{2 * f}."
        g.normal(p + "pre.weight", (H, 96, 1), np.sqrt(1.0 / 96))
        g.normal(p + "pre.bias", (H,), 0.02)
        g.normal(p + "enc.cond_layer.weight", (2 * H * 3, 256, 1), np.sqrt(1.0 / 256) * 0.5)
        g.normal(p + "enc.cond_layer.bias", (2 * H * 3,), 0.02)
        for i in range(3):
            g.normal(p + f"enc.in_layers.{i}.weight", (2 * H, H, 5), np.sqrt(1.0 / (5 * H)))
            g.normal(p + f"enc.in_layers.{i}.bias", (2 * H,), 0.02)
            rs = 2 * H if i < 2 else H
            g.normal(p + f"enc.res_skip_layers.{i}.weight", (rs, H, 1), np.sqrt(1.0 / H))
            g.normal(p + f"enc.res_skip_layers.{i}.bias", (rs,), 0.02)
        g.normal(p + "post.weight", (96, H, 1), np.sqrt(1.0 / H) * 0.3)
        g.normal(p + "post.bias", (96,), 0.02)
    g.normal("emb_g.weight", (109, 256), 1.0)
    # decoder (GeneratorNSF)
    g.t["dec.m_source.l_linear.weight"] = np.asarray([[0.9]], np.float32)
    g.t["dec.m_source.l_linear.bias"] = np.asarray([0.01], np.float32)
    g.normal("dec.conv_pre.weight", (512, H, 7), np.sqrt(1.0 / (7 * H)))
    g.normal("dec.conv_pre.bias", (512,), 0.02)
    g.normal("dec.cond.weight", (512, 256, 1), np.sqrt(1.0 / 256) * 0.3)
    g.normal("dec.cond.bias", (512,), 0.02)
    from .nets import GEN_CONFIGS
    rates, kernels = GEN_CONFIGS[sr]
    for i in range(4):
        cin, cout = 512 >> i, 512 >> (i + 1)
        k, u = kernels[i], rates[i]
        # each output sample sees k/u taps of cin channels
        g.normal(f"dec.ups.{i}.weight", (cin, cout, k), np.sqrt(1.6 / (cin * k / u)))
        g.normal(f"dec.ups.{i}.bias", (cout,), 0.02)
        if i + 1 < 4:
            sf = int(np.prod(rates[i + 1:]))
            g.normal(f"dec.noise_convs.{i}.weight", (cout, 1, sf * 2), 2.0 / np.sqrt(sf * 2.0))
        else:
            g.normal(f"dec.noise_convs.{i}.weight", (cout, 1, 1), 2.0)
        g.normal(f"dec.noise_convs.{i}.bias", (cout,), 0.02)
        for j, rk in enumerate((3, 7, 11)):
            p = f"dec.resblocks.{i * 3 + j}."
            for d in range(3):
                g.normal(p + f"convs1.{d}.weight", (cout, cout, rk), np.sqrt(2.0 / (cout * rk)))
                g.normal(p + f"convs1.{d}.bias", (cout,), 0.02)
                g.normal(p + f"convs2.{d}.weight", (cout, cout, rk),
                         0.35 * np.sqrt(2.0 / (cout * rk)))
                g.normal(p + f"convs2.{d}.bias", (cout,), 0.02)
    g.normal("dec.conv_post.weight", (1, 32, 7), 0.35 * np.sqrt(1.0 / (32 * 7)))
    g.meta("meta.sr", sr)
    g.meta("meta.sid", 0)
    g.meta("meta.phone_dim", phone_dim)
    return g.t


def synth_index(seed: int, n: int, c: int, std: float = 0.34) -> np.ndarray:
    """i.i.d. N(0, 0.34^2) rows (per-element rms of feats.npy)."""
    rng = np.random.Generator(np.random.PCG64([seed, n, c]))
    return (rng.standard_normal((n, c), dtype=np.float32) * np.float32(std)).astype(np.float32)


def make_data_dir(root: str, seed: int = 7, v1: bool = False, index_rows: int = 0, sr: int = 40000) -> dict:
    """Writes the reference's on-disk layout (rvc.rs:48,57,66; models.rs:58-61,71-73) with
    `.rvcw` instead of `.onnx`:  <root>/contentvec/vec-{C}-layer-{L}.rvcw, <root>/f0/rmvpe.rvcw,
    <root>/voice.rvcw [, <root>/voice.index.rvcw].  Returns the paths."""
    c, l = (256, 9) if v1 else (768, 12)
    paths = {
        "data": root,
        "contentvec": os.path.join(root, "contentvec", f"vec-{c}-layer-{l}.rvcw"),
        "f0": os.path.join(root, "f0", "rmvpe.rvcw"),
        "model": os.path.join(root, "voice.rvcw"),
    }
    if not os.path.exists(paths["contentvec"]):
        write_rvcw(paths["contentvec"], synth_contentvec(seed + 1, l, v1))
    if not os.path.exists(paths["f0"]):
        write_rvcw(paths["f0"], synth_rmvpe(seed + 2))
    if not os.path.exists(paths["model"]):
        write_rvcw(paths["model"], synth_voice(seed + 3, sr, c))
    if index_rows:
        paths["index"] = os.path.join(root, f"voice.{index_rows}x{c}.index.rvcw")
        if not os.path.exists(paths["index"]):
            write_rvcw(paths["index"], {"big_npy": synth_index(seed + 4, index_rows, c)})
    return paths
