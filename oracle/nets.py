"""torch-CPU fp32 restatement of the three networks behind the reference's opaque .onnx graphs
(TEST INFRASTRUCTURE; PARITY UNPINNED - see oracle/__init__.py).

Call sites in the reference: ContentVec `rvc/src/rvc.rs:92` ("source" -> "embed"), RMVPE
`rvc/src/f0/rmvpe.rs:235` ("input" -> "output"), synthesizer `rvc/src/rvc.rs:195`
("phone","pitch","pitchf" -> "audio").  Bodies: SURVEY.md Appendix C ([UPSTREAM] public RVC
WebUI / fairseq definitions).  All functions are pure: `w` is a dict name -> torch tensor
(`oracle.weights`), `rec` (optional dict) receives named intermediates in the engine's
channels-last layouts for stage-by-stage parity.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from .weights import HUBERT_KERNELS, HUBERT_STRIDES


def to_torch(w: dict) -> dict:
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in w.items()}


def _rec(rec, name, t):
    if rec is not None:
        rec[name] = t.detach().contiguous().numpy().copy()


# ----------------------------------------------------------------------------------------------
# ContentVec / HuBERT-base
# ----------------------------------------------------------------------------------------------


@torch.no_grad()
def hubert_forward(w: dict, pcm: torch.Tensor, rec=None) -> torch.Tensor:
    """pcm (N,) -> (T, C) with T = (N-400)/320+1.  V2: layer-12 hidden (768); V1: layer-9
    hidden -> final_proj (256) (rvc-common/src/enums.rs:9-23)."""
    n_layers = int(w["meta.n_layers"][0])
    x = pcm.reshape(1, 1, -1).float()
    for i, (k, s) in enumerate(zip(HUBERT_KERNELS, HUBERT_STRIDES)):
        x = F.conv1d(x, w[f"feature_extractor.conv_layers.{i}.0.weight"], stride=s)
        if i == 0:
            x = F.group_norm(x, 512, w["feature_extractor.conv_layers.0.2.weight"],
                             w["feature_extractor.conv_layers.0.2.bias"], eps=1e-5)
        x = F.gelu(x)
        _rec(rec, f"cv.conv{i}", x[0].t())
    x = x.transpose(1, 2)                                            # (1,T,512)
    x = F.layer_norm(x, (512,), w["layer_norm.weight"], w["layer_norm.bias"], 1e-5)
    x = F.linear(x, w["post_extract_proj.weight"], w["post_extract_proj.bias"])
    _rec(rec, "cv.proj", x[0])
    pc = F.conv1d(x.transpose(1, 2), w["encoder.pos_conv.0.weight"], w["encoder.pos_conv.0.bias"],
                  padding=64, groups=16)
    pc = F.gelu(pc[:, :, :-1])                                       # SamePad(128) drops the last
    x = x + pc.transpose(1, 2)
    x = F.layer_norm(x, (768,), w["encoder.layer_norm.weight"], w["encoder.layer_norm.bias"], 1e-5)
    _rec(rec, "cv.enc_in", x[0])
    T = x.shape[1]
    for i in range(n_layers):
        p = f"encoder.layers.{i}."
        q = F.linear(x, w[p + "self_attn.q_proj.weight"], w[p + "self_attn.q_proj.bias"]) * 0.125
        k = F.linear(x, w[p + "self_attn.k_proj.weight"], w[p + "self_attn.k_proj.bias"])
        v = F.linear(x, w[p + "self_attn.v_proj.weight"], w[p + "self_attn.v_proj.bias"])
        q, k, v = (t.reshape(T, 12, 64).transpose(0, 1) for t in (q, k, v))   # (12,T,64)
        a = torch.softmax(q @ k.transpose(1, 2), dim=-1) @ v                   # (12,T,64)
        a = a.transpose(0, 1).reshape(1, T, 768)
        a = F.linear(a, w[p + "self_attn.out_proj.weight"], w[p + "self_attn.out_proj.bias"])
        x = F.layer_norm(x + a, (768,), w[p + "self_attn_layer_norm.weight"],
                         w[p + "self_attn_layer_norm.bias"], 1e-5)
        h = F.gelu(F.linear(x, w[p + "fc1.weight"], w[p + "fc1.bias"]))
        h = F.linear(h, w[p + "fc2.weight"], w[p + "fc2.bias"])
        x = F.layer_norm(x + h, (768,), w[p + "final_layer_norm.weight"],
                         w[p + "final_layer_norm.bias"], 1e-5)
        _rec(rec, f"cv.layer{i}", x[0])
    if "final_proj.weight" in w:
        x = F.linear(x, w["final_proj.weight"], w["final_proj.bias"])
    _rec(rec, "cv.out", x[0])
    return x[0]


# ----------------------------------------------------------------------------------------------
# RMVPE  E2E(n_blocks=4, n_gru=1, kernel (2,2))
# ----------------------------------------------------------------------------------------------


def _bn2d(w, p, x, eps=1e-5):
    return F.batch_norm(x, w[p + ".running_mean"], w[p + ".running_var"], w[p + ".weight"],
                        w[p + ".bias"], False, 0.0, eps)


def _conv_block_res(w, p, x):
    y = F.conv2d(x, w[p + "conv.0.weight"], padding=1)
    y = F.relu(_bn2d(w, p + "conv.1", y))
    y = F.conv2d(y, w[p + "conv.3.weight"], padding=1)
    y = F.relu(_bn2d(w, p + "conv.4", y))
    if p + "shortcut.weight" in w:
        return y + F.conv2d(x, w[p + "shortcut.weight"], w[p + "shortcut.bias"])
    return y + x


def gru_direction(x, w_ih, w_hh, b_ih, b_hh, reverse: bool):
    """Single-layer GRU, PyTorch gate order [r, z, n]; x (T, I) -> (T, H)."""
    T = x.shape[0]
    H = w_hh.shape[1]
    gi = x @ w_ih.t() + b_ih
    h = torch.zeros(H, dtype=x.dtype)
    out = torch.zeros(T, H, dtype=x.dtype)
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        gh = w_hh @ h + b_hh
        r = torch.sigmoid(gi[t, :H] + gh[:H])
        z = torch.sigmoid(gi[t, H:2 * H] + gh[H:2 * H])
        n = torch.tanh(gi[t, 2 * H:] + r * gh[2 * H:])
        h = (1.0 - z) * n + z * h
        out[t] = h
    return out


@torch.no_grad()
def rmvpe_forward(w: dict, mel: torch.Tensor, rec=None) -> torch.Tensor:
    """mel (128, T) f32 log-mel, T a multiple of 32 -> salience (T, 360) in (0,1)."""
    x = mel.t().reshape(1, 1, mel.shape[1], 128).float()              # (1,1,T,128)
    x = _bn2d(w, "unet.encoder.bn", x)
    skips = []
    for i in range(5):
        for j in range(4):
            x = _conv_block_res(w, f"unet.encoder.layers.{i}.conv.{j}.", x)
        _rec(rec, f"rm.enc{i}", x[0].permute(1, 2, 0))
        skips.append(x)
        x = F.avg_pool2d(x, (2, 2))
    for i in range(4):
        for j in range(4):
            x = _conv_block_res(w, f"unet.intermediate.layers.{i}.conv.{j}.", x)
    _rec(rec, "rm.inter", x[0].permute(1, 2, 0))
    for i in range(5):
        p = f"unet.decoder.layers.{i}."
        x = F.conv_transpose2d(x, w[p + "conv1.0.weight"], stride=(2, 2), padding=(1, 1),
                               output_padding=(1, 1))
        x = F.relu(_bn2d(w, p + "conv1.1", x))
        x = torch.cat((x, skips[-1 - i]), dim=1)
        for j in range(4):
            x = _conv_block_res(w, p + f"conv2.{j}.", x)
        _rec(rec, f"rm.dec{i}", x[0].permute(1, 2, 0))
    x = F.conv2d(x, w["cnn.weight"], w["cnn.bias"], padding=1)       # (1,3,T,128)
    x = x.transpose(1, 2).flatten(-2)[0]                              # (T, 384), index c*128+f
    _rec(rec, "rm.cnn", x)
    g = "fc.0.gru."
    hf = gru_direction(x, w[g + "weight_ih_l0"], w[g + "weight_hh_l0"], w[g + "bias_ih_l0"],
                       w[g + "bias_hh_l0"], False)
    hb = gru_direction(x, w[g + "weight_ih_l0_reverse"], w[g + "weight_hh_l0_reverse"],
                       w[g + "bias_ih_l0_reverse"], w[g + "bias_hh_l0_reverse"], True)
    h = torch.cat((hf, hb), dim=1)                                    # (T,512)
    _rec(rec, "rm.gru", h)
    out = torch.sigmoid(F.linear(h, w["fc.1.weight"], w["fc.1.bias"]))
    _rec(rec, "rm.salience", out)
    return out


# ----------------------------------------------------------------------------------------------
# SynthesizerTrnMs768NSFsid (40k)
# ----------------------------------------------------------------------------------------------

H = 192
WINDOW = 10


def _ln_c(x, gamma, beta, eps=1e-5):
    """attentions.LayerNorm: over channels of (B,C,T)."""
    return F.layer_norm(x.transpose(1, -1), (x.shape[1],), gamma, beta, eps).transpose(1, -1)


def _rel_attention(w, p, x):
    """MultiHeadAttention(192,192,n_heads=2,window_size=10,heads_share=True) on (1,C,T):
    scores[i,j] += q_i . Ek[j-i+W], out[i] += sum_j p[i,j] Ev[j-i+W] for |j-i| <= W."""
    T = x.shape[2]
    q = F.conv1d(x, w[p + "conv_q.weight"], w[p + "conv_q.bias"])
    k = F.conv1d(x, w[p + "conv_k.weight"], w[p + "conv_k.bias"])
    v = F.conv1d(x, w[p + "conv_v.weight"], w[p + "conv_v.bias"])
    q, k, v = (t.reshape(2, 96, T).transpose(1, 2) for t in (q, k, v))          # (2,T,96)
    qs = q / math.sqrt(96.0)
    scores = qs @ k.transpose(1, 2)                                               # (2,T,T)
    ek, ev = w[p + "emb_rel_k"][0], w[p + "emb_rel_v"][0]                         # (21,96)
    ii = torch.arange(T)
    rel = ii[None, :] - ii[:, None] + WINDOW                                      # j-i+W
    valid = (rel >= 0) & (rel <= 2 * WINDOW)
    relc = rel.clamp(0, 2 * WINDOW)
    rl = torch.einsum("hid,ijd->hij", qs, ek[relc])                               # (2,T,T)
    scores = scores + torch.where(valid[None], rl, torch.zeros_like(rl))
    pa = torch.softmax(scores, dim=-1)
    out = pa @ v
    pv = torch.where(valid[None], pa, torch.zeros_like(pa))
    out = out + torch.einsum("hij,ijd->hid", pv, ev[relc])
    out = out.transpose(1, 2).reshape(1, H, T)
    return F.conv1d(out, w[p + "conv_o.weight"], w[p + "conv_o.bias"])


def text_encoder(w, phone, pitch, rec=None):
    """enc_p: phone (T,768) f32, pitch (T,) int -> m, logs each (1,192,T)."""
    x = F.linear(phone[None], w["enc_p.emb_phone.weight"], w["enc_p.emb_phone.bias"])
    x = x + w["enc_p.emb_pitch.weight"][pitch.long()][None]
    x = x * math.sqrt(H)
    x = F.leaky_relu(x, 0.1).transpose(1, 2)                                      # (1,192,T)
    _rec(rec, "sy.emb", x[0].t())
    for i in range(6):
        y = _rel_attention(w, f"enc_p.encoder.attn_layers.{i}.", x)
        x = _ln_c(x + y, w[f"enc_p.encoder.norm_layers_1.{i}.gamma"],
                  w[f"enc_p.encoder.norm_layers_1.{i}.beta"])
        p = f"enc_p.encoder.ffn_layers.{i}."
        y = F.relu(F.conv1d(x, w[p + "conv_1.weight"], w[p + "conv_1.bias"], padding=1))
        y = F.conv1d(y, w[p + "conv_2.weight"], w[p + "conv_2.bias"], padding=1)
        x = _ln_c(x + y, w[f"enc_p.encoder.norm_layers_2.{i}.gamma"],
                  w[f"enc_p.encoder.norm_layers_2.{i}.beta"])
        _rec(rec, f"sy.enc{i}", x[0].t())
    stats = F.conv1d(x, w["enc_p.proj.weight"], w["enc_p.proj.bias"])
    return stats[:, :H], stats[:, H:]


def _wn(w, p, x, g):
    """modules.WN(192, 5, 1, 3, gin=256)."""
    out = torch.zeros_like(x)
    gc = F.conv1d(g, w[p + "cond_layer.weight"], w[p + "cond_layer.bias"])        # (1,1152,1)
    for i in range(3):
        x_in = F.conv1d(x, w[p + f"in_layers.{i}.weight"], w[p + f"in_layers.{i}.bias"], padding=2)
        a = x_in + gc[:, i * 2 * H:(i + 1) * 2 * H]
        acts = torch.tanh(a[:, :H]) * torch.sigmoid(a[:, H:])
        rs = F.conv1d(acts, w[p + f"res_skip_layers.{i}.weight"],
                      w[p + f"res_skip_layers.{i}.bias"])
        if i < 2:
            x = x + rs[:, :H]
            out = out + rs[:, H:]
        else:
            out = out + rs
    return out


def flow_reverse(w, z, g, rec=None):
    """ResidualCouplingBlock(192,192,5,1,3, n_flows=4, gin=256).forward(reverse=True):
    for flow in reversed([C0,Flip,C1,Flip,C2,Flip,C3,Flip])."""
    for f in (3, 2, 1, 0):
        z = torch.flip(z, [1])
        p = f"flow.flows.{2 * f}."
        x0, x1 = z[:, :96], z[:, 96:]
        h = F.conv1d(x0, w[p + "pre.weight"], w[p + "pre.bias"])
        h = _wn(w, p + "enc.", h, g)
        m = F.conv1d(h, w[p + "post.weight"], w[p + "post.bias"])
        z = torch.cat([x0, x1 - m], dim=1)
        _rec(rec, f"sy.flow{f}", z[0].t())
    return z


def sine_gen(f0: torch.Tensor, upp: int, sr: int, noise: torch.Tensor, sine_amp=0.1,
             noise_std=0.003, voiced_threshold=0.0) -> torch.Tensor:
    """SineGen(harmonic_num=0).forward(f0, upp): f0 (T,) -> (T*upp,)."""
    f0 = f0.reshape(1, -1, 1).float()
    rad = (f0 / sr) % 1
    tmp = torch.cumsum(rad, 1) * upp
    tmp = F.interpolate(tmp.transpose(2, 1), scale_factor=float(upp), mode="linear",
                        align_corners=True).transpose(2, 1)
    rad_up = F.interpolate(rad.transpose(2, 1), scale_factor=float(upp),
                           mode="nearest").transpose(2, 1)
    tmp = tmp % 1
    idx = (tmp[:, 1:, :] - tmp[:, :-1, :]) < 0
    shift = torch.zeros_like(rad_up)
    shift[:, 1:, :] = idx * -1.0
    sine = torch.sin(torch.cumsum(rad_up + shift, dim=1) * 2 * np.pi) * sine_amp
    uv = (f0 > voiced_threshold).float()
    uv = F.interpolate(uv.transpose(2, 1), scale_factor=float(upp), mode="nearest").transpose(2, 1)
    noise_amp = uv * noise_std + (1 - uv) * sine_amp / 3
    sine = sine * uv + noise_amp * noise.reshape(1, -1, 1)
    return sine[0, :, 0]


# upstream RVC v2 generator configs (configs/v2/{32k,48k}.json, configs/v1/40k.json): upsample_rates / upsample_kernel_sizes
GEN_CONFIGS = {32000: ((10, 8, 2, 2), (20, 16, 4, 4)), 40000: ((10, 10, 2, 2), (16, 16, 4, 4)), 48000: ((12, 10, 2, 2), (24, 20, 4, 4))}
RATES, UP_KERNELS = GEN_CONFIGS[40000]
RES_KERNELS = (3, 7, 11)
RES_DILATIONS = (1, 3, 5)


def generator_nsf(w, z, f0, g, sr, noise_sine, rec=None):
    """GeneratorNSF.forward: z (1,192,T), f0 (T,) -> audio (T * sr / 100,)."""
    RATES, UP_KERNELS = GEN_CONFIGS[int(sr)]
    upp = int(np.prod(RATES))
    sine = sine_gen(f0, upp, sr, noise_sine)
    _rec(rec, "sy.sine", sine)
    har = torch.tanh(sine * w["dec.m_source.l_linear.weight"][0, 0] +
                     w["dec.m_source.l_linear.bias"][0]).reshape(1, 1, -1)
    x = F.conv1d(z, w["dec.conv_pre.weight"], w["dec.conv_pre.bias"], padding=3)
    x = x + F.conv1d(g, w["dec.cond.weight"], w["dec.cond.bias"])
    _rec(rec, "sy.conv_pre", x[0].t())
    for i in range(4):
        k, u = UP_KERNELS[i], RATES[i]
        x = F.leaky_relu(x, 0.1)
        x = F.conv_transpose1d(x, w[f"dec.ups.{i}.weight"], w[f"dec.ups.{i}.bias"], stride=u,
                               padding=(k - u) // 2)
        if i + 1 < 4:
            sf = int(np.prod(RATES[i + 1:]))
            xs = F.conv1d(har, w[f"dec.noise_convs.{i}.weight"], w[f"dec.noise_convs.{i}.bias"],
                          stride=sf, padding=sf // 2)
        else:
            xs = F.conv1d(har, w[f"dec.noise_convs.{i}.weight"], w[f"dec.noise_convs.{i}.bias"])
        x = x + xs
        _rec(rec, f"sy.up{i}", x[0].t())
        acc = None
        for j, rk in enumerate(RES_KERNELS):
            p = f"dec.resblocks.{i * 3 + j}."
            y = x
            for d, dil in enumerate(RES_DILATIONS):
                t = F.leaky_relu(y, 0.1)
                t = F.conv1d(t, w[p + f"convs1.{d}.weight"], w[p + f"convs1.{d}.bias"],
                             dilation=dil, padding=(rk * dil - dil) // 2)
                t = F.leaky_relu(t, 0.1)
                t = F.conv1d(t, w[p + f"convs2.{d}.weight"], w[p + f"convs2.{d}.bias"],
                             padding=(rk - 1) // 2)
                y = t + y
            acc = y if acc is None else acc + y
        x = acc / 3
        _rec(rec, f"sy.stage{i}", x[0].t())
    x = F.leaky_relu(x)                                   # default slope 0.01 (upstream)
    x = torch.tanh(F.conv1d(x, w["dec.conv_post.weight"], None, padding=3))
    return x[0, 0]


@torch.no_grad()
def synth_forward(w: dict, phone: torch.Tensor, pitch: torch.Tensor, pitchf: torch.Tensor,
                  noise_z: torch.Tensor, noise_sine: torch.Tensor, rec=None) -> torch.Tensor:
    """phone (T,768) f32, pitch (T,) i32 coarse, pitchf (T,) f32 Hz, noise_z (T,192),
    noise_sine (T*sr/100,) -> audio (T*sr/100,) f32."""
    sr = int(w["meta.sr"][0])
    sid = int(w["meta.sid"][0])
    g = w["emb_g.weight"][sid].reshape(1, 256, 1)
    m, logs = text_encoder(w, phone.float(), pitch, rec)
    _rec(rec, "sy.m", m[0].t())
    _rec(rec, "sy.logs", logs[0].t())
    z_p = m + torch.exp(logs) * noise_z.t()[None] * 0.66666
    _rec(rec, "sy.z_p", z_p[0].t())
    z = flow_reverse(w, z_p, g, rec)
    audio = generator_nsf(w, z, pitchf.float(), g, sr, noise_sine, rec)
    _rec(rec, "sy.audio", audio)
    return audio
