"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the CogVideoX 3-D causal-conv VAE decoder — SURVEY.md §8 row a-7.

Restates diffusers' `AutoencoderKLCogVideoX.decode` / `CogVideoXDecoder3D` as the reference drives it
(`generate/CogVideoX-5B.py:20-21,72-77`: tiling + slicing enabled, bf16). diffusers (>=0.31, requirements.txt:20) is NOT
installable in this image and no checkpoint is reachable, so this file follows SURVEY.md App. A.5 from recollection of the
library: **parity unpinned** against the real package. Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may
import it. Plain torch ops, channels-first [B, C, T, H, W] like the library.

Semantics restated
  * CausalConv3d (pad_mode "first"): zero pad H/W by 1; time is padded in front with the last 2 input frames of the
    previous frame batch (conv_cache) or, for the first batch, 2 copies of the first frame.
  * SpatialNorm3D(f, zq) = GroupNorm(32, eps 1e-6)(f) * conv_y(zq') + conv_b(zq'), zq' = nearest-resized zq (first frame
    handled separately when T > 1 is odd), conv_y / conv_b pointwise (1x1x1) with bias.
  * ResnetBlock3D: norm1 -> SiLU -> conv1 -> norm2 -> SiLU -> conv2, 1x1x1 shortcut when Cin != Cout.
  * Upsample3D: nearest x2 in H, W (and in T when compress_time; an odd T > 1 keeps the first frame single) then a per-frame
    Conv2d 3x3.
  * Frame batching: 2 latent frames per decoder call (the first call takes the remainder), conv caches carried over.
  * Tiling: latent tiles 30x45 with stride 25x36, outputs blended linearly over 40 / 72 px and cropped to 200 x 288 px;
    caches reset per tile; GroupNorm statistics are per (tile, frame batch).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch
import torch.nn.functional as F


@dataclass
class VAEConfig:
    latent_channels: int = 16
    out_channels: int = 3
    block_out_channels: tuple = (128, 256, 256, 512)
    layers_per_block: int = 3
    norm_num_groups: int = 32
    temporal_compression_ratio: int = 4
    scaling_factor: float = 0.7
    sample_height: int = 480
    sample_width: int = 720
    num_latent_frames_batch_size: int = 2
    tile_overlap_factor_height: float = 1 / 6
    tile_overlap_factor_width: float = 1 / 5

    @property
    def reversed_channels(self):
        return tuple(reversed(self.block_out_channels))

    @property
    def temporal_compress_level(self):
        import math
        return int(math.log2(self.temporal_compression_ratio))

    # tiling geometry (AutoencoderKLCogVideoX.__init__ / tiled_decode)
    @property
    def tile_sample_min_height(self): return self.sample_height // 2
    @property
    def tile_sample_min_width(self): return self.sample_width // 2
    @property
    def spatial_scale(self): return 2 ** (len(self.block_out_channels) - 1)
    @property
    def tile_latent_min_height(self): return int(self.tile_sample_min_height / self.spatial_scale)
    @property
    def tile_latent_min_width(self): return int(self.tile_sample_min_width / self.spatial_scale)


def random_state_dict(cfg: VAEConfig, seed: int = 99, std: float = 0.05, dtype=torch.float32) -> dict:
    """Seeded random decoder weights with the diffusers parameter names. GroupNorm affine randomised so parity
    tests see it; conv_y initialised around 1 so activations stay O(1) through the stack."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    zc = cfg.latent_channels

    def conv(name, co, ci, *k, scale=None):
        fan = ci
        for kk in k:
            fan *= kk
        s = scale if scale is not None else (1.0 / fan) ** 0.5
        sd[name + ".weight"] = (torch.randn(co, ci, *k, generator=g) * s).to(dtype)
        sd[name + ".bias"] = (torch.randn(co, generator=g) * 0.02).to(dtype)

    def snorm(name, c):
        sd[name + ".norm_layer.weight"] = (1.0 + 0.1 * torch.randn(c, generator=g)).to(dtype)
        sd[name + ".norm_layer.bias"] = (0.05 * torch.randn(c, generator=g)).to(dtype)
        conv(name + ".conv_y.conv", c, zc, 1, 1, 1, scale=0.05)
        sd[name + ".conv_y.conv.bias"] = (1.0 + 0.05 * torch.randn(c, generator=g)).to(dtype)
        conv(name + ".conv_b.conv", c, zc, 1, 1, 1, scale=0.05)

    def resnet(name, ci, co):
        snorm(name + ".norm1", ci)
        conv(name + ".conv1.conv", co, ci, 3, 3, 3)
        snorm(name + ".norm2", co)
        conv(name + ".conv2.conv", co, co, 3, 3, 3)
        if ci != co:
            conv(name + ".conv_shortcut", co, ci, 1, 1, 1)

    rc = cfg.reversed_channels
    conv("decoder.conv_in.conv", rc[0], zc, 3, 3, 3)
    for j in range(2):
        resnet(f"decoder.mid_block.resnets.{j}", rc[0], rc[0])
    cin = rc[0]
    for i, co in enumerate(rc):
        for j in range(cfg.layers_per_block + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}", cin if j == 0 else co, co)
        if i != len(rc) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", co, co, 3, 3)
        cin = co
    snorm("decoder.norm_out", rc[-1])
    conv("decoder.conv_out.conv", cfg.out_channels, rc[-1], 3, 3, 3)
    return sd


# ------------------------------------------------------------------ layers
class _Cache(dict):
    """conv_cache of one decoder pass: layer name -> last 2 time-padded input frames."""


def causal_conv3d(sd, name, x, cache_in: dict | None, cache_out: dict):
    w, b = sd[name + ".weight"], sd[name + ".bias"]
    kt = w.shape[2]
    if kt > 1:
        if cache_in is not None and name in cache_in:
            x = torch.cat([cache_in[name], x], dim=2)
        else:
            x = torch.cat([x[:, :, :1]] * (kt - 1) + [x], dim=2)
        cache_out[name] = x[:, :, -(kt - 1):].clone()
    pad = w.shape[3] // 2
    return F.conv3d(x, w.to(x.dtype), b.to(x.dtype), padding=(0, pad, pad))


def spatial_norm(sd, name, f, zq, groups, cache_in, cache_out):
    T = f.shape[2]
    if T > 1 and T % 2 == 1:
        z_first = F.interpolate(zq[:, :, :1], size=(1,) + tuple(f.shape[-2:]))
        z_rest = F.interpolate(zq[:, :, 1:], size=(T - 1,) + tuple(f.shape[-2:]))
        zq = torch.cat([z_first, z_rest], dim=2)
    else:
        zq = F.interpolate(zq, size=tuple(f.shape[-3:]))
    y = causal_conv3d(sd, name + ".conv_y.conv", zq, cache_in, cache_out)
    bb = causal_conv3d(sd, name + ".conv_b.conv", zq, cache_in, cache_out)
    n = F.group_norm(f, groups, sd[name + ".norm_layer.weight"].to(f.dtype), sd[name + ".norm_layer.bias"].to(f.dtype), eps=1e-6)
    return n * y + bb


def resnet_block(sd, name, x, zq, groups, cache_in, cache_out):
    h = spatial_norm(sd, name + ".norm1", x, zq, groups, cache_in, cache_out)
    h = F.silu(h)
    h = causal_conv3d(sd, name + ".conv1.conv", h, cache_in, cache_out)
    h = spatial_norm(sd, name + ".norm2", h, zq, groups, cache_in, cache_out)
    h = F.silu(h)
    h = causal_conv3d(sd, name + ".conv2.conv", h, cache_in, cache_out)
    if name + ".conv_shortcut.weight" in sd:
        x = F.conv3d(x, sd[name + ".conv_shortcut.weight"].to(x.dtype), sd[name + ".conv_shortcut.bias"].to(x.dtype))
    return x + h


def upsample3d(sd, name, x, compress_time: bool):
    B, C, T, H, W = x.shape
    if compress_time:
        if T > 1 and T % 2 == 1:
            first = F.interpolate(x[:, :, 0], scale_factor=2.0)
            rest = F.interpolate(x[:, :, 1:], scale_factor=2.0)
            x = torch.cat([first[:, :, None], rest], dim=2)
        elif T > 1:
            x = F.interpolate(x, scale_factor=2.0)
        else:
            x = F.interpolate(x.squeeze(2), scale_factor=2.0)[:, :, None]
    else:
        x = x.permute(0, 2, 1, 3, 4).reshape(B * T, C, H, W)
        x = F.interpolate(x, scale_factor=2.0)
        x = x.reshape(B, T, C, 2 * H, 2 * W).permute(0, 2, 1, 3, 4)
    B, C, T, H, W = x.shape
    y = x.permute(0, 2, 1, 3, 4).reshape(B * T, C, H, W)
    y = F.conv2d(y, sd[name + ".conv.weight"].to(x.dtype), sd[name + ".conv.bias"].to(x.dtype), padding=1)
    return y.reshape(B, T, -1, H, W).permute(0, 2, 1, 3, 4)


def decoder_forward(sd, cfg: VAEConfig, z, cache_in: dict | None):
    """CogVideoXDecoder3D.forward(sample=z, conv_cache) -> (frames, new conv_cache)."""
    cache_out = _Cache()
    g = cfg.norm_num_groups
    h = causal_conv3d(sd, "decoder.conv_in.conv", z, cache_in, cache_out)
    for j in range(2):
        h = resnet_block(sd, f"decoder.mid_block.resnets.{j}", h, z, g, cache_in, cache_out)
    rc = cfg.reversed_channels
    for i in range(len(rc)):
        for j in range(cfg.layers_per_block + 1):
            h = resnet_block(sd, f"decoder.up_blocks.{i}.resnets.{j}", h, z, g, cache_in, cache_out)
        if i != len(rc) - 1:
            h = upsample3d(sd, f"decoder.up_blocks.{i}.upsamplers.0", h, compress_time=i < cfg.temporal_compress_level)
    h = spatial_norm(sd, "decoder.norm_out", h, z, g, cache_in, cache_out)
    h = F.silu(h)
    h = causal_conv3d(sd, "decoder.conv_out.conv", h, cache_in, cache_out)
    return h, cache_out


def frame_batches(num_frames: int, fb: int):
    """[(start, end)] of AutoencoderKLCogVideoX._decode's frame-batch loop."""
    nb = max(num_frames // fb, 1)
    rem = num_frames % fb
    return [(fb * i + (0 if i == 0 else rem), fb * (i + 1) + rem) for i in range(nb)]


def decode_untiled(sd, cfg: VAEConfig, z):
    cache = None
    outs = []
    for (s, e) in frame_batches(z.shape[2], cfg.num_latent_frames_batch_size):
        o, cache = decoder_forward(sd, cfg, z[:, :, s:e], cache)
        outs.append(o)
    return torch.cat(outs, dim=2)


def blend_v(a, b, extent):
    extent = min(a.shape[3], b.shape[3], extent)
    for y in range(extent):
        b[:, :, :, y, :] = a[:, :, :, -extent + y, :] * (1 - y / extent) + b[:, :, :, y, :] * (y / extent)
    return b


def blend_h(a, b, extent):
    extent = min(a.shape[4], b.shape[4], extent)
    for x in range(extent):
        b[:, :, :, :, x] = a[:, :, :, :, -extent + x] * (1 - x / extent) + b[:, :, :, :, x] * (x / extent)
    return b


def tiling_geometry(cfg: VAEConfig):
    tlh, tlw = cfg.tile_latent_min_height, cfg.tile_latent_min_width
    return dict(
        tile_latent_h=tlh, tile_latent_w=tlw,
        overlap_h=int(tlh * (1 - cfg.tile_overlap_factor_height)), overlap_w=int(tlw * (1 - cfg.tile_overlap_factor_width)),
        blend_h=int(cfg.tile_sample_min_height * cfg.tile_overlap_factor_height),
        blend_w=int(cfg.tile_sample_min_width * cfg.tile_overlap_factor_width),
        limit_h=cfg.tile_sample_min_height - int(cfg.tile_sample_min_height * cfg.tile_overlap_factor_height),
        limit_w=cfg.tile_sample_min_width - int(cfg.tile_sample_min_width * cfg.tile_overlap_factor_width))


def decode(sd, cfg: VAEConfig, z, tiling: bool = True):
    """AutoencoderKLCogVideoX.decode(z).sample for z [B, C, T, h, w] (already divided by scaling_factor)."""
    B, C, T, H, W = z.shape
    geo = tiling_geometry(cfg)
    if not (tiling and (W > geo["tile_latent_w"] or H > geo["tile_latent_h"])):
        return decode_untiled(sd, cfg, z)
    rows = []
    for i in range(0, H, geo["overlap_h"]):
        row = []
        for j in range(0, W, geo["overlap_w"]):
            row.append(decode_untiled(sd, cfg, z[:, :, :, i:i + geo["tile_latent_h"], j:j + geo["tile_latent_w"]]))
        rows.append(row)
    result_rows = []
    for i, row in enumerate(rows):
        result_row = []
        for j, tile in enumerate(row):
            if i > 0:
                tile = blend_v(rows[i - 1][j], tile, geo["blend_h"])
            if j > 0:
                tile = blend_h(row[j - 1], tile, geo["blend_w"])
            result_row.append(tile[:, :, :, :geo["limit_h"], :geo["limit_w"]])
        result_rows.append(torch.cat(result_row, dim=4))
    return torch.cat(result_rows, dim=3)


# ====================================================================== encoder
# Restates diffusers' CogVideoXEncoder3D / AutoencoderKLCogVideoX.encode as the reference's latent encoders drive it
# (`train/CogVideoX-5B/02_encode.py:108-115`: `vae.encode(video).latent_dist.sample()`; the I2V pipeline encodes the first
# frame the same way). Like the decoder this is recalled from the library: **parity unpinned**.
#   * CogVideoXDownBlock3D: `layers_per_block` resnets with plain GroupNorm(32, eps 1e-6) (no spatial conditioning), then
#     CogVideoXDownsample3D on all but the last block: when compress_time, avg_pool1d(kernel 2, stride 2) over time (an odd
#     frame count keeps the first frame and pools the rest), then F.pad(0,1,0,1) + per-frame Conv2d 3x3 stride 2.
#   * mid block: 2 resnets; norm_out GroupNorm -> SiLU -> conv_out to 2 * latent_channels (mean, logvar).
#   * frame batching: 8 sample frames per encoder call, the first call takes the remainder (49 -> 9 + 5 x 8), conv caches
#     carried across calls; tiling: sample tiles 240 x 360 with stride 200 x 288, latent blend extents 5 / 9, crop 25 x 36.
#   * DiagonalGaussianDistribution: logvar clamped to [-30, 20]; sample = mean + exp(0.5 logvar) * noise.
NUM_SAMPLE_FRAMES_BATCH_SIZE = 8


def random_encoder_state_dict(cfg: VAEConfig, seed: int = 98, dtype=torch.float32, in_channels: int = 3) -> dict:
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, co, ci, *k):
        fan = ci
        for kk in k:
            fan *= kk
        sd[name + ".weight"] = (torch.randn(co, ci, *k, generator=g) * (1.0 / fan) ** 0.5).to(dtype)
        sd[name + ".bias"] = (torch.randn(co, generator=g) * 0.02).to(dtype)

    def gn(name, c):
        sd[name + ".weight"] = (1.0 + 0.1 * torch.randn(c, generator=g)).to(dtype)
        sd[name + ".bias"] = (0.05 * torch.randn(c, generator=g)).to(dtype)

    def resnet(name, ci, co):
        gn(name + ".norm1", ci)
        conv(name + ".conv1.conv", co, ci, 3, 3, 3)
        gn(name + ".norm2", co)
        conv(name + ".conv2.conv", co, co, 3, 3, 3)
        if ci != co:
            conv(name + ".conv_shortcut", co, ci, 1, 1, 1)

    boc = cfg.block_out_channels
    conv("encoder.conv_in.conv", boc[0], in_channels, 3, 3, 3)
    cin = boc[0]
    for i, co in enumerate(boc):
        for j in range(cfg.layers_per_block):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}", cin if j == 0 else co, co)
        if i != len(boc) - 1:
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", co, co, 3, 3)
        cin = co
    for j in range(2):
        resnet(f"encoder.mid_block.resnets.{j}", boc[-1], boc[-1])
    gn("encoder.norm_out", boc[-1])
    conv("encoder.conv_out.conv", 2 * cfg.latent_channels, boc[-1], 3, 3, 3)
    return sd


def gn_resnet_block(sd, name, x, groups, cache_in, cache_out):
    h = F.group_norm(x, groups, sd[name + ".norm1.weight"].to(x.dtype), sd[name + ".norm1.bias"].to(x.dtype), eps=1e-6)
    h = causal_conv3d(sd, name + ".conv1.conv", F.silu(h), cache_in, cache_out)
    h = F.group_norm(h, groups, sd[name + ".norm2.weight"].to(x.dtype), sd[name + ".norm2.bias"].to(x.dtype), eps=1e-6)
    h = causal_conv3d(sd, name + ".conv2.conv", F.silu(h), cache_in, cache_out)
    if name + ".conv_shortcut.weight" in sd:
        x = F.conv3d(x, sd[name + ".conv_shortcut.weight"].to(x.dtype), sd[name + ".conv_shortcut.bias"].to(x.dtype))
    return x + h


def downsample3d(sd, name, x, compress_time: bool):
    B, C, T, H, W = x.shape
    if compress_time:
        y = x.permute(0, 3, 4, 1, 2).reshape(B * H * W, C, T)
        if T % 2 == 1:
            first, rest = y[..., 0], y[..., 1:]
            if rest.shape[-1] > 0:
                rest = F.avg_pool1d(rest, kernel_size=2, stride=2)
            y = torch.cat([first[..., None], rest], dim=-1)
        else:
            y = F.avg_pool1d(y, kernel_size=2, stride=2)
        T = y.shape[-1]
        x = y.reshape(B, H, W, C, T).permute(0, 3, 4, 1, 2)
    x = F.pad(x, (0, 1, 0, 1), mode="constant", value=0)
    y = x.permute(0, 2, 1, 3, 4).reshape(B * T, C, H + 1, W + 1)
    y = F.conv2d(y, sd[name + ".conv.weight"].to(x.dtype), sd[name + ".conv.bias"].to(x.dtype), stride=2)
    return y.reshape(B, T, -1, y.shape[-2], y.shape[-1]).permute(0, 2, 1, 3, 4)


def encoder_forward(sd, cfg: VAEConfig, x, cache_in: dict | None):
    """CogVideoXEncoder3D.forward(sample=x, conv_cache) -> (moments [B, 2*latent, T', h, w], new conv_cache)."""
    cache_out = _Cache()
    g = cfg.norm_num_groups
    boc = cfg.block_out_channels
    h = causal_conv3d(sd, "encoder.conv_in.conv", x, cache_in, cache_out)
    for i in range(len(boc)):
        for j in range(cfg.layers_per_block):
            h = gn_resnet_block(sd, f"encoder.down_blocks.{i}.resnets.{j}", h, g, cache_in, cache_out)
        if i != len(boc) - 1:
            h = downsample3d(sd, f"encoder.down_blocks.{i}.downsamplers.0", h, compress_time=i < cfg.temporal_compress_level)
    for j in range(2):
        h = gn_resnet_block(sd, f"encoder.mid_block.resnets.{j}", h, g, cache_in, cache_out)
    h = F.group_norm(h, g, sd["encoder.norm_out.weight"].to(h.dtype), sd["encoder.norm_out.bias"].to(h.dtype), eps=1e-6)
    h = causal_conv3d(sd, "encoder.conv_out.conv", F.silu(h), cache_in, cache_out)
    return h, cache_out


def encode_untiled(sd, cfg: VAEConfig, x):
    cache = None
    outs = []
    for (s, e) in frame_batches(x.shape[2], NUM_SAMPLE_FRAMES_BATCH_SIZE):
        o, cache = encoder_forward(sd, cfg, x[:, :, s:e], cache)
        outs.append(o)
    return torch.cat(outs, dim=2)


def encode_tiling_geometry(cfg: VAEConfig):
    th, tw = cfg.tile_sample_min_height, cfg.tile_sample_min_width
    bh = int(cfg.tile_latent_min_height * cfg.tile_overlap_factor_height)
    bw = int(cfg.tile_latent_min_width * cfg.tile_overlap_factor_width)
    return dict(tile_h=th, tile_w=tw,
                overlap_h=int(th * (1 - cfg.tile_overlap_factor_height)), overlap_w=int(tw * (1 - cfg.tile_overlap_factor_width)),
                blend_h=bh, blend_w=bw, limit_h=cfg.tile_latent_min_height - bh, limit_w=cfg.tile_latent_min_width - bw)


def encode(sd, cfg: VAEConfig, x, tiling: bool = True):
    """AutoencoderKLCogVideoX.encode(x).latent_dist.parameters for x [B, 3, T, H, W]: the moments (mean | logvar)."""
    B, C, T, H, W = x.shape
    geo = encode_tiling_geometry(cfg)
    if not (tiling and (W > geo["tile_w"] or H > geo["tile_h"])):
        return encode_untiled(sd, cfg, x)
    rows = []
    for i in range(0, H, geo["overlap_h"]):
        row = []
        for j in range(0, W, geo["overlap_w"]):
            row.append(encode_untiled(sd, cfg, x[:, :, :, i:i + geo["tile_h"], j:j + geo["tile_w"]]))
        rows.append(row)
    result_rows = []
    for i, row in enumerate(rows):
        result_row = []
        for j, tile in enumerate(row):
            if i > 0:
                tile = blend_v(rows[i - 1][j], tile, geo["blend_h"])
            if j > 0:
                tile = blend_h(row[j - 1], tile, geo["blend_w"])
            result_row.append(tile[:, :, :, :geo["limit_h"], :geo["limit_w"]])
        result_rows.append(torch.cat(result_row, dim=4))
    return torch.cat(result_rows, dim=3)


def gaussian_sample(moments, noise):
    """DiagonalGaussianDistribution(moments).sample() with the noise passed in (the library draws it with randn_tensor)."""
    mean, logvar = torch.chunk(moments, 2, dim=1)
    logvar = torch.clamp(logvar, -30.0, 20.0)
    return mean + torch.exp(0.5 * logvar) * noise
