"""ORACLE (test infrastructure, NOT product code): CPU restatement of MAGVITv2.decode_code.

Restates MMaDA-Parallel-M/models/modeling_magvitv2.py: LFQuantizer.get_codebook_entry :208-221, VQGANDecoder.forward
:365-399, MAGVITv2.decode_code :429-433, and the blocks of models/common_modules.py: nonlinearity :12-14,
Normalize :17-20, Upsample :27-40, AttnBlock :168-211, ResnetBlock :298-357 - as one functional forward over a state
dict with the reference's parameter names. fp32 throughout, like the reference (the VQ model is loaded without dtype).
Pinned against the real modules by oracle/make_golden_magvit.py.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, List

import torch
import torch.nn.functional as F

from .sampling import lfq_codebook_entry


def decoder_config(ch=128, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=(4, 4, 3, 4, 3), z_channels=13, out_ch=3):
    """Defaults of VQGANDecoder.__init__ (modeling_magvitv2.py:278-288)."""
    return SimpleNamespace(ch=ch, ch_mult=tuple(ch_mult), num_res_blocks=tuple(num_res_blocks), z_channels=z_channels,
                           out_ch=out_ch)


def param_shapes(cfg) -> Dict[str, tuple]:
    """Names/shapes of the decoder parameters exactly as the reference module registers them (prefix 'decoder.')."""
    sh: Dict[str, tuple] = {}

    def conv(name, cout, cin, k):
        sh[name + ".weight"] = (cout, cin, k, k)
        sh[name + ".bias"] = (cout,)

    def norm(name, c):
        sh[name + ".weight"] = (c,)
        sh[name + ".bias"] = (c,)

    def resblock(name, cin, cout):
        norm(name + ".norm1", cin)
        conv(name + ".conv1", cout, cin, 3)
        norm(name + ".norm2", cout)
        conv(name + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(name + ".nin_shortcut", cout, cin, 1)

    nres = len(cfg.ch_mult)
    block_in = cfg.ch * cfg.ch_mult[nres - 1]
    conv("decoder.conv_in", block_in, cfg.z_channels, 3)
    resblock("decoder.mid.block_1", block_in, block_in)
    norm("decoder.mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        conv("decoder.mid.attn_1." + n, block_in, block_in, 1)
    resblock("decoder.mid.block_2", block_in, block_in)
    for i_level in reversed(range(nres)):
        block_out = cfg.ch * cfg.ch_mult[i_level]
        for i_block in range(cfg.num_res_blocks[i_level]):
            resblock(f"decoder.up.{i_level}.block.{i_block}", block_in, block_out)
            block_in = block_out
        if i_level != 0:
            conv(f"decoder.up.{i_level}.upsample.conv", block_in, block_in, 3)
    norm("decoder.norm_out", block_in)
    conv("decoder.conv_out", cfg.out_ch, block_in, 3)
    conv("decoder.post_quant_conv", cfg.z_channels, cfg.z_channels, 1)
    return sh


def make_weights(cfg, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Deterministic synthetic decoder weights (fp32), fan-in scaled so activations stay O(1) through the stack."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in param_shapes(cfg).items():
        if name.endswith(".weight") and len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            sd[name] = torch.randn(shape, generator=g) * (1.0 / fan_in) ** 0.5
        elif len(shape) == 1 and ".norm" in name and name.endswith(".weight"):
            sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            sd[name] = 0.05 * torch.randn(shape, generator=g)
    return sd


def nonlinearity(x):
    return x * torch.sigmoid(x)                                                   # common_modules.py:12-14


def _norm(x, w, name):
    return F.group_norm(x, 32, w[name + ".weight"], w[name + ".bias"], eps=1e-6)  # Normalize, :17-20


def _conv(x, w, name, padding):
    return F.conv2d(x, w[name + ".weight"], w[name + ".bias"], stride=1, padding=padding)


def _resblock(x, w, name):
    """ResnetBlock.forward with temb None, dropout 0 (common_modules.py:335-357)."""
    h = _conv(nonlinearity(_norm(x, w, name + ".norm1")), w, name + ".conv1", 1)
    h = _conv(nonlinearity(_norm(h, w, name + ".norm2")), w, name + ".conv2", 1)
    if name + ".nin_shortcut.weight" in w:
        x = _conv(x, w, name + ".nin_shortcut", 0)
    return x + h


def _attn(x, w, name):
    """AttnBlock.forward (common_modules.py:186-211)."""
    h_ = _norm(x, w, name + ".norm")
    q, k, v = (_conv(h_, w, name + "." + n, 0) for n in ("q", "k", "v"))
    b, c, h, wd = q.shape
    q = q.reshape(b, c, h * wd).permute(0, 2, 1)
    k = k.reshape(b, c, h * wd)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, h * wd)
    h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, h, wd)
    return x + _conv(h_, w, name + ".proj_out", 0)


@torch.no_grad()
def decoder_forward(z: torch.Tensor, w: Dict[str, torch.Tensor], cfg) -> torch.Tensor:
    """VQGANDecoder.forward (modeling_magvitv2.py:365-399): z [B, 13, h, w] -> [B, 3, 16h', 16w'] for the default config."""
    nres = len(cfg.ch_mult)
    z = _conv(z, w, "decoder.post_quant_conv", 0)
    h = _conv(z, w, "decoder.conv_in", 1)
    h = _resblock(h, w, "decoder.mid.block_1")
    h = _attn(h, w, "decoder.mid.attn_1")
    h = _resblock(h, w, "decoder.mid.block_2")
    for i_level in reversed(range(nres)):
        for i_block in range(cfg.num_res_blocks[i_level]):
            h = _resblock(h, w, f"decoder.up.{i_level}.block.{i_block}")
        if i_level != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")               # Upsample, common_modules.py:36-40
            h = _conv(h, w, f"decoder.up.{i_level}.upsample.conv", 1)
    h = nonlinearity(_norm(h, w, "decoder.norm_out"))
    return _conv(h, w, "decoder.conv_out", 1)


@torch.no_grad()
def decode_code(indices: torch.Tensor, w: Dict[str, torch.Tensor], cfg, shape=None) -> torch.Tensor:
    """MAGVITv2.decode_code (modeling_magvitv2.py:429-433)."""
    b, n = indices.shape
    if shape is None:
        hh = ww = int(n ** 0.5)
    else:
        hh, ww = shape
    z_q = lfq_codebook_entry(indices, cfg.z_channels).view(b, cfg.z_channels, hh, ww)
    return decoder_forward(z_q, w, cfg)


# ---------------------------------------------------------------------------------------------------------------
# encoder side: MAGVITv2.get_code (modeling_magvitv2.py:423-427) = VQGANEncoder.forward (:143-169) + LFQ sign bits
# ---------------------------------------------------------------------------------------------------------------
def encoder_config(ch=128, ch_mult=(1, 2, 2, 4, 4), num_res_blocks=(4, 3, 4, 3, 4), z_channels=13, in_ch=3):
    """Defaults of VQGANEncoder.__init__ (modeling_magvitv2.py:62-73)."""
    return SimpleNamespace(ch=ch, ch_mult=tuple(ch_mult), num_res_blocks=tuple(num_res_blocks), z_channels=z_channels, in_ch=in_ch)


def encoder_param_shapes(cfg) -> Dict[str, tuple]:
    sh: Dict[str, tuple] = {}

    def conv(name, cout, cin, k):
        sh[name + ".weight"] = (cout, cin, k, k)
        sh[name + ".bias"] = (cout,)

    def norm(name, c):
        sh[name + ".weight"] = (c,)
        sh[name + ".bias"] = (c,)

    def resblock(name, cin, cout):
        norm(name + ".norm1", cin)
        conv(name + ".conv1", cout, cin, 3)
        norm(name + ".norm2", cout)
        conv(name + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(name + ".nin_shortcut", cout, cin, 1)

    conv("encoder.conv_in", cfg.ch, cfg.in_ch, 3)
    in_ch_mult = (1,) + tuple(cfg.ch_mult)
    nres = len(cfg.ch_mult)
    block_in = cfg.ch
    for i_level in range(nres):
        block_in = cfg.ch * in_ch_mult[i_level]
        block_out = cfg.ch * cfg.ch_mult[i_level]
        for i_block in range(cfg.num_res_blocks[i_level]):
            resblock(f"encoder.down.{i_level}.block.{i_block}", block_in, block_out)
            block_in = block_out
        if i_level != nres - 1:
            conv(f"encoder.down.{i_level}.downsample.conv", block_in, block_in, 3)
    resblock("encoder.mid.block_1", block_in, block_in)
    norm("encoder.mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        conv("encoder.mid.attn_1." + n, block_in, block_in, 1)
    resblock("encoder.mid.block_2", block_in, block_in)
    norm("encoder.norm_out", block_in)
    conv("encoder.conv_out", cfg.z_channels, block_in, 3)
    conv("encoder.quant_conv", cfg.z_channels, cfg.z_channels, 1)
    return sh


def make_encoder_weights(cfg, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in encoder_param_shapes(cfg).items():
        if name.endswith(".weight") and len(shape) == 4:
            sd[name] = torch.randn(shape, generator=g) * (1.0 / (shape[1] * shape[2] * shape[3])) ** 0.5
        elif len(shape) == 1 and ".norm" in name and name.endswith(".weight"):
            sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            sd[name] = 0.05 * torch.randn(shape, generator=g)
    return sd


@torch.no_grad()
def encoder_forward(x: torch.Tensor, w: Dict[str, torch.Tensor], cfg) -> torch.Tensor:
    """VQGANEncoder.forward (modeling_magvitv2.py:143-169): pixels [B, 3, H, W] -> pre-quantisation z [B, 13, H/16, W/16]."""
    nres = len(cfg.ch_mult)
    h = _conv(x, w, "encoder.conv_in", 1)
    for i_level in range(nres):
        for i_block in range(cfg.num_res_blocks[i_level]):
            h = _resblock(h, w, f"encoder.down.{i_level}.block.{i_block}")
        if i_level != nres - 1:
            h = F.pad(h, (0, 1, 0, 1), mode="constant", value=0)                 # Downsample, common_modules.py:83-87
            h = F.conv2d(h, w[f"encoder.down.{i_level}.downsample.conv.weight"], w[f"encoder.down.{i_level}.downsample.conv.bias"],
                         stride=2, padding=0)
    h = _resblock(h, w, "encoder.mid.block_1")
    h = _attn(h, w, "encoder.mid.attn_1")
    h = _resblock(h, w, "encoder.mid.block_2")
    h = nonlinearity(_norm(h, w, "encoder.norm_out"))
    h = _conv(h, w, "encoder.conv_out", 1)
    return _conv(h, w, "encoder.quant_conv", 0)


def lfq_indices(z: torch.Tensor, bits: int = 13) -> torch.Tensor:
    """LFQuantizer.forward sign quantisation + get_indices (modeling_magvitv2.py:201-206,238-241): [B, 13, h, w] -> int64 [B, h*w]."""
    power = 2 ** torch.arange(bits - 1, -1, -1)
    return (power.reshape(1, -1, 1, 1) * (z > 0).long()).sum(1).reshape(z.shape[0], -1)


@torch.no_grad()
def get_code(pixels: torch.Tensor, w: Dict[str, torch.Tensor], cfg) -> torch.Tensor:
    return lfq_indices(encoder_forward(pixels, w, cfg), cfg.z_channels)
