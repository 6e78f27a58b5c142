"""CogVideoX 3-D causal-conv VAE decoder on the sm_100a kernels — drop-in for the part of diffusers'
AutoencoderKLCogVideoX the reference uses: `vae.enable_tiling()`, `vae.enable_slicing()`, `vae.decode(z).sample`,
`vae.config.scaling_factor`, `vae.dtype` (generate/CogVideoX-5B.py:20-21,72-77).

Every arithmetic step is a C-ABI kernel (include/videogpa_b200.h): the 3x3x3 / 1x3x3 convolutions are implicit GEMMs
on tcgen05 (vgpa_conv3d_causal_bf16), the 1x1x1 shortcuts and the conv_y / conv_b projections of SpatialNorm3D are
the DiT GEMM (vgpa_linear_bf16), GroupNorm statistics + SpatialNorm apply + SiLU, the nearest upsample and the tile
blend are HBM-bound kernels. torch is used for device memory, views and frame copies (conv_cache) only.

Layout: activations are channels-last [T, H, W, C] bf16 per latent tile and frame batch. The reference semantics
that change results are kept: frame batching (2 latent frames per decoder pass, the first pass takes the remainder),
conv_cache carry-over between passes, GroupNorm statistics per (tile, frame batch), 3x3 latent tiles of 30x45 with
stride 25x36 blended over 40 / 72 px and cropped to 200x288 px.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from types import SimpleNamespace

import torch

from . import _lib, dense
from ._lib import ComposeArgs, Conv3dArgs, SpatialNormArgs

BF16 = torch.bfloat16


@dataclass
class VAEDecoderConfig:
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
    num_sample_frames_batch_size: int = 8
    in_channels: int = 3
    tile_overlap_factor_height: float = 1 / 6
    tile_overlap_factor_width: float = 1 / 5


class DecoderOutput(SimpleNamespace):
    """`.sample` like diffusers' DecoderOutput."""


class _Conv:
    __slots__ = ("w", "b", "kt", "cin", "cout", "cout_pad")


class _SNorm:
    __slots__ = ("gamma", "beta", "c", "off")      # off: first column of this norm's conv_y block in the fused projection


def _pad_cout(c: int) -> int:
    if c % 256 == 0 or c in (64, 128):
        return c
    if c in (32, 48):                                # encoder conv_out: 2 * latent_channels moments
        return 64
    if c <= 16:
        return 16
    raise RuntimeError(f"VAE conv with {c} output channels is not supported (need 64, 128 or a multiple of 256)")


class AutoencoderKLCogVideoXDecoder:
    def __init__(self, state_dict: dict, config: VAEDecoderConfig | None = None, device="cuda"):
        self.config = config or VAEDecoderConfig()
        self.device = torch.device(device)
        self.dtype = BF16
        self.use_tiling = False
        self.use_slicing = False
        self.tile_streams = 9           # latent tiles are independent: one CUDA stream per tile of the 3x3 grid (512 ms per 49-frame clip against 524 at 4, 560 at 1)
        self._streams = None
        self.use_cuda_graph = False
        self._graphs = {}
        self.fuse_gn_stats = True        # GroupNorm statistics from the conv epilogues (False: separate two-kernel pass per norm)
        self._ws_conv = None
        self._capturing = False
        c = self.config
        self.rc = tuple(reversed(c.block_out_channels))
        self.temporal_compress_level = int(math.log2(c.temporal_compression_ratio))
        self.zc_pad = 64                                                    # latent channels padded to one K block
        if c.latent_channels > self.zc_pad:
            raise RuntimeError("latent_channels > 64 is not supported")
        self._norm_cols = 0
        self._yb_w, self._yb_b = [], []
        self._load(state_dict)
        self._ws = None
        # tiling geometry (AutoencoderKLCogVideoX.__init__)
        self.tile_sample_min_height = c.sample_height // 2
        self.tile_sample_min_width = c.sample_width // 2
        scale = 2 ** (len(c.block_out_channels) - 1)
        self.spatial_scale = scale
        self.tile_latent_min_height = int(self.tile_sample_min_height / scale)
        self.tile_latent_min_width = int(self.tile_sample_min_width / scale)

    # ------------------------------------------------------------------ construction
    @classmethod
    def random_init(cls, config: VAEDecoderConfig | None = None, seed: int = 5, device="cuda") -> "AutoencoderKLCogVideoXDecoder":
        """Synthetic decoder weights with the diffusers parameter names (no checkpoint is reachable):
        fan-in scaled normal conv weights, GroupNorm gamma ~ 1, conv_y bias ~ 1 so activations stay O(1)."""
        cfg = config or VAEDecoderConfig()
        dev = torch.device(device)
        g = torch.Generator(device=dev).manual_seed(seed)
        sd, zc = {}, cfg.latent_channels

        def conv(name, co, ci, *k, scale=None):
            fan = ci * math.prod(k)
            sd[name + ".weight"] = (torch.randn(co, ci, *k, generator=g, device=dev) * (scale or (1.0 / fan) ** 0.5)).to(BF16)
            sd[name + ".bias"] = (torch.randn(co, generator=g, device=dev) * 0.02).to(BF16)

        def snorm(name, ch):
            sd[name + ".norm_layer.weight"] = (1.0 + 0.1 * torch.randn(ch, generator=g, device=dev)).to(BF16)
            sd[name + ".norm_layer.bias"] = (0.05 * torch.randn(ch, generator=g, device=dev)).to(BF16)
            conv(name + ".conv_y.conv", ch, zc, 1, 1, 1, scale=0.05)
            sd[name + ".conv_y.conv.bias"] = (1.0 + 0.05 * torch.randn(ch, generator=g, device=dev)).to(BF16)
            conv(name + ".conv_b.conv", ch, zc, 1, 1, 1, scale=0.05)

        def resnet(name, ci, co):
            snorm(name + ".norm1", ci); conv(name + ".conv1.conv", co, ci, 3, 3, 3)
            snorm(name + ".norm2", co); conv(name + ".conv2.conv", co, co, 3, 3, 3)
            if ci != co:
                conv(name + ".conv_shortcut", co, ci, 1, 1, 1)

        rc = tuple(reversed(cfg.block_out_channels))
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
        return cls(sd, cfg, device=dev)

    # ------------------------------------------------------------------ weights
    def _conv(self, sd, name, pad_cin_to: int | None = None) -> _Conv:
        if name + ".weight" not in sd:
            raise RuntimeError(f"state dict is missing {name}.weight")
        w = sd[name + ".weight"].to(device=self.device, dtype=torch.float32)
        b = sd[name + ".bias"].to(device=self.device, dtype=torch.float32)
        if w.dim() == 4:                                                    # Conv2d of the upsampler: [co, ci, 3, 3]
            w = w[:, :, None]
        co, ci, kt, kh, kw = w.shape
        if (kh, kw) != (3, 3) or kt not in (1, 3):
            raise RuntimeError(f"{name}: unsupported kernel {tuple(w.shape)}")
        cin = ci if pad_cin_to is None else pad_cin_to
        if cin % 64 != 0:
            raise RuntimeError(f"{name}: input channels {cin} must be a multiple of 64")
        cp = _pad_cout(co)
        w2 = torch.zeros(cp, kt, kh, kw, cin, device=self.device, dtype=torch.float32)
        w2[:co, :, :, :, :ci] = w.permute(0, 2, 3, 4, 1)
        bb = torch.zeros(cp, device=self.device, dtype=torch.float32)
        bb[:co] = b
        cv = _Conv()
        cv.w = w2.reshape(cp, kt * kh * kw * cin).to(BF16).contiguous()
        cv.b = bb.to(BF16).contiguous()
        cv.kt, cv.cin, cv.cout, cv.cout_pad = kt, cin, co, cp
        return cv

    def _snorm(self, sd, name, ch: int) -> _SNorm:
        n = _SNorm()
        n.gamma = sd[name + ".norm_layer.weight"].to(device=self.device, dtype=BF16).contiguous()
        n.beta = sd[name + ".norm_layer.bias"].to(device=self.device, dtype=BF16).contiguous()
        n.c = ch
        n.off = self._norm_cols
        zc = self.config.latent_channels
        for part in ("conv_y", "conv_b"):                                   # [C, zc, 1, 1, 1] pointwise convs of zq
            w = torch.zeros(ch, self.zc_pad, device=self.device, dtype=torch.float32)
            w[:, :zc] = sd[f"{name}.{part}.conv.weight"].to(device=self.device, dtype=torch.float32).reshape(ch, zc)
            self._yb_w.append(w)
            self._yb_b.append(sd[f"{name}.{part}.conv.bias"].to(device=self.device, dtype=torch.float32))
        self._norm_cols += 2 * ch
        return n

    def _resnet(self, sd, name, ci, co):
        r = SimpleNamespace()
        r.norm1 = self._snorm(sd, name + ".norm1", ci)
        r.conv1 = self._conv(sd, name + ".conv1.conv")
        r.norm2 = self._snorm(sd, name + ".norm2", co)
        r.conv2 = self._conv(sd, name + ".conv2.conv")
        r.sc_w = r.sc_b = None
        if ci != co:
            w = sd[name + ".conv_shortcut.weight"]
            if w.shape[2:] != (1, 1, 1):
                raise RuntimeError(f"{name}.conv_shortcut: only the 1x1x1 shortcut of the released checkpoints is supported")
            r.sc_w = w.reshape(co, ci).to(device=self.device, dtype=BF16).contiguous()
            r.sc_b = sd[name + ".conv_shortcut.bias"].to(device=self.device, dtype=BF16).contiguous()
        r.ci, r.co = ci, co
        return r

    def _load(self, sd: dict) -> None:
        c, rc = self.config, self.rc
        self.conv_in = self._conv(sd, "decoder.conv_in.conv", pad_cin_to=self.zc_pad)
        self.mid = [self._resnet(sd, f"decoder.mid_block.resnets.{j}", rc[0], rc[0]) for j in range(2)]
        self.up = []
        cin = rc[0]
        for i, co in enumerate(rc):
            blk = SimpleNamespace()
            blk.resnets = [self._resnet(sd, f"decoder.up_blocks.{i}.resnets.{j}", cin if j == 0 else co, co)
                           for j in range(c.layers_per_block + 1)]
            blk.upsample = None
            if i != len(rc) - 1:
                blk.upsample = self._conv(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv")
                blk.compress_time = i < self.temporal_compress_level
            self.up.append(blk)
            cin = co
        self.norm_out = self._snorm(sd, "decoder.norm_out", rc[-1])
        self.conv_out = self._conv(sd, "decoder.conv_out.conv")
        # one fused projection for every conv_y / conv_b of the decoder: [sum 2C, 64]
        ncol = self._norm_cols
        ncol_pad = (ncol + 255) // 256 * 256
        W = torch.zeros(ncol_pad, self.zc_pad, device=self.device, dtype=torch.float32)
        B = torch.zeros(ncol_pad, device=self.device, dtype=torch.float32)
        W[:ncol] = torch.cat(self._yb_w, 0)
        B[:ncol] = torch.cat(self._yb_b, 0)
        self.yb_w, self.yb_b = W.to(BF16).contiguous(), B.to(BF16).contiguous()
        self._yb_w = self._yb_b = None

    # ------------------------------------------------------------------ diffusers-style switches
    def enable_tiling(self, *a, **k):
        self.use_tiling = True

    def disable_tiling(self):
        self.use_tiling = False

    def enable_slicing(self):
        self.use_slicing = True

    def disable_slicing(self):
        self.use_slicing = False

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    # ------------------------------------------------------------------ kernel wrappers
    def _gn_stats(self, x: torch.Tensor, C_: int) -> torch.Tensor:
        lib = _lib.load()
        need = lib.vgpa_groupnorm_workspace_bytes(C_)
        if self._ws is None:
            self._ws = {}
        key = torch.cuda.current_stream().cuda_stream            # one scratch buffer per stream: tiles decode concurrently
        ws = self._ws.get(key)
        if ws is None or ws.numel() < need:
            ws = self._ws[key] = torch.empty(max(need, 1 << 22), dtype=torch.uint8, device=self.device)
        out = torch.empty(2 * self.config.norm_num_groups, dtype=torch.float32, device=self.device)
        n_pix = x.numel() // C_
        _lib.check(lib.vgpa_groupnorm_stats_bf16(x.data_ptr(), n_pix, C_, self.config.norm_num_groups, 1e-6, ws.data_ptr(),
                                                 ws.numel(), out.data_ptr(), _lib.current_stream()), "vgpa_groupnorm_stats_bf16")
        return out

    @staticmethod
    def _tz_map(T: int, Tz: int):
        """Frame of zq seen by frame t of a [T]-frame feature map (CogVideoXSpatialNorm3D's nearest resize)."""
        if T > 1 and T % 2 == 1:
            rest = [1 + int(math.floor((k) * ((Tz - 1) / (T - 1)))) for k in range(T - 1)] if Tz > 1 else [0] * (T - 1)
            return [0] + rest
        return [int(math.floor(k * (Tz / T))) for k in range(T)]

    def _snorm_apply(self, x, norm: _SNorm, yb, zshape, out, silu: bool, stats=None):
        """x [T,H,W,C] -> out (same shape, may be a view 2 frames into a time-padded buffer). `stats` = the GroupNorm
        mean / rstd of x when the conv that wrote x already produced them in its epilogue."""
        lib = _lib.load()
        T, H, W, C_ = x.shape
        Tz, Hz, Wz = zshape
        if stats is None:
            stats = self._gn_stats(x, C_)
        a = SpatialNormArgs()
        a.x, a.out, a.mean_rstd = x.data_ptr(), out.data_ptr(), stats.data_ptr()
        a.gamma, a.beta = norm.gamma.data_ptr(), norm.beta.data_ptr()
        esz = yb.element_size()
        a.y_lat = yb.data_ptr() + norm.off * esz
        a.b_lat = yb.data_ptr() + (norm.off + C_) * esz
        a.ld_lat = yb.stride(0)
        a.T, a.H, a.W, a.C, a.groups = T, H, W, C_, self.config.norm_num_groups
        a.Hz, a.Wz = Hz, Wz
        shift = int(round(math.log2(H / Hz))) if H >= Hz else 0
        if (Hz << shift) < H or (Wz << shift) < W:
            raise RuntimeError(f"feature map {H}x{W} is not a power-of-two multiple of the latent tile {Hz}x{Wz}")
        a.shift = shift
        for i, tz in enumerate(self._tz_map(T, Tz)):
            a.tz_of_t[i] = tz
        a.silu = 1 if silu else 0
        _lib.check(lib.vgpa_spatialnorm_apply_bf16(C.byref(a), _lib.current_stream()), "vgpa_spatialnorm_apply_bf16")

    def _conv_call(self, cv: _Conv, xpad: torch.Tensor, T: int, out: torch.Tensor | None = None, residual=None, want_stats: bool = False):
        """-> out, or (out, mean_rstd [2 * groups] fp32) with want_stats: the GroupNorm statistics of `out`, accumulated by the
        conv epilogue (no separate pass over the tensor that the next norm would otherwise read once more)."""
        lib = _lib.load()
        Tp, H, W, Cin = xpad.shape
        if Tp != T + cv.kt - 1 or Cin != cv.cin:
            raise RuntimeError(f"conv input {tuple(xpad.shape)} does not match T={T}, kt={cv.kt}, Cin={cv.cin}")
        ldo = 16 if cv.cout_pad == 16 else cv.cout
        if out is None:
            out = torch.empty((T, H, W, ldo), dtype=BF16, device=self.device)
        a = Conv3dArgs()
        a.x, a.w, a.bias, a.out = xpad.data_ptr(), cv.w.data_ptr(), cv.b.data_ptr(), out.data_ptr()
        a.residual = residual.data_ptr() if residual is not None else None
        a.T, a.H, a.W, a.Cin, a.Cout, a.Cout_pad, a.KT = T, H, W, Cin, cv.cout, cv.cout_pad, cv.kt
        a.ldo = ldo
        a.ld_res = residual.shape[-1] if residual is not None else 0
        stats = None
        if want_stats and self.fuse_gn_stats and cv.cout_pad != 16 and cv.cout <= 512:
            need = lib.vgpa_conv3d_gn_workspace_bytes(cv.cout)
            if self._ws_conv is None:
                self._ws_conv = {}
            key = torch.cuda.current_stream().cuda_stream        # one scratch buffer per stream: tiles decode concurrently
            ws = self._ws_conv.get(key)
            if ws is None or ws.numel() < need:
                ws = self._ws_conv[key] = torch.empty(max(need, 1 << 20), dtype=torch.uint8, device=self.device)
            stats = torch.empty(2 * self.config.norm_num_groups, dtype=torch.float32, device=self.device)
            a.gn_mean_rstd, a.gn_workspace = stats.data_ptr(), ws.data_ptr()
            a.gn_groups, a.gn_eps = self.config.norm_num_groups, 1e-6
        _lib.check(lib.vgpa_conv3d_causal_bf16(C.byref(a), _lib.current_stream()), "vgpa_conv3d_causal_bf16")
        if want_stats:
            return out, stats
        return out

    def _finish_timepad(self, buf: torch.Tensor, key: str, cache_in: dict | None, cache_out: dict):
        """buf [T+2, H, W, C] with frames [2:] already written: fill the 2 leading frames from the conv_cache (or
        replicate the first frame) and record the new cache = last 2 frames of the padded input."""
        if cache_in is not None and key in cache_in:
            buf[:2].copy_(cache_in[key])
        else:
            buf[0].copy_(buf[2])
            buf[1].copy_(buf[2])
        cache_out[key] = buf[-2:].clone()

    def _norm_conv(self, x, norm, cv, key, yb, zshape, cache_in, cache_out, residual=None, out=None, x_stats=None, want_stats=False):
        """SpatialNorm -> SiLU -> causal conv (the repeated unit of CogVideoXResnetBlock3D and the output head). x_stats: the
        GroupNorm statistics of x if its producer made them; want_stats: also return those of the result."""
        T, H, W, C_ = x.shape
        buf = torch.empty((T + 2, H, W, C_), dtype=BF16, device=self.device)
        self._snorm_apply(x, norm, yb, zshape, buf[2:], silu=True, stats=x_stats)
        self._finish_timepad(buf, key, cache_in, cache_out)
        return self._conv_call(cv, buf, T, out=out, residual=residual, want_stats=want_stats)

    def _resnet_fwd(self, r, x, key, yb, zshape, cache_in, cache_out, x_stats=None):
        """-> (block output, its GroupNorm statistics): every tensor a norm reads is written by a conv epilogue that also
        accumulates its statistics (conv1 -> norm2, conv2 + residual -> the next block's norm1)."""
        T, H, W, _ = x.shape
        h, h_stats = self._norm_conv(x, r.norm1, r.conv1, key + ".conv1", yb, zshape, cache_in, cache_out, x_stats=x_stats, want_stats=True)
        res = x
        if r.sc_w is not None:
            res = dense.linear(x.view(-1, r.ci), r.sc_w, r.sc_b).view(T, H, W, r.co)
        return self._norm_conv(h, r.norm2, r.conv2, key + ".conv2", yb, zshape, cache_in, cache_out, residual=res, x_stats=h_stats,
                               want_stats=True)

    def _upsample_fwd(self, blk, x):
        lib = _lib.load()
        T, H, W, C_ = x.shape
        if blk.compress_time and T > 1:
            t_src = [0] + [1 + k // 2 for k in range(2 * (T - 1))] if T % 2 == 1 else [k // 2 for k in range(2 * T)]
        else:
            t_src = list(range(T))
        To = len(t_src)
        if To > 16:
            raise RuntimeError("more than 16 frames per decoder pass are not supported")
        up = torch.empty((To, 2 * H, 2 * W, C_), dtype=BF16, device=self.device)
        arr = (C.c_int32 * 16)(*(t_src + [0] * (16 - To)))
        _lib.check(lib.vgpa_upsample_nearest_bf16(x.data_ptr(), up.data_ptr(), To, H, W, C_, arr, _lib.current_stream()),
                   "vgpa_upsample_nearest_bf16")
        return self._conv_call(blk.upsample, up, To, want_stats=True)

    # ------------------------------------------------------------------ one decoder pass over one frame batch of one tile
    def _decoder_pass(self, zt: torch.Tensor, cache_in: dict | None, out: torch.Tensor):
        """zt [Tz, hz, wz, 64] channels-last padded latent; out [T_out, 8hz, 8wz, 16] receives the frames."""
        Tz, hz, wz, _ = zt.shape
        zshape = (Tz, hz, wz)
        cache_out: dict = {}
        yb = dense.linear(zt.view(-1, self.zc_pad), self.yb_w, self.yb_b)          # conv_y / conv_b of every SpatialNorm
        buf = torch.empty((Tz + 2, hz, wz, self.zc_pad), dtype=BF16, device=self.device)
        buf[2:].copy_(zt)
        self._finish_timepad(buf, "conv_in", cache_in, cache_out)
        h, st = self._conv_call(self.conv_in, buf, Tz, want_stats=True)
        for j, r in enumerate(self.mid):
            h, st = self._resnet_fwd(r, h, f"mid.{j}", yb, zshape, cache_in, cache_out, x_stats=st)
        for i, blk in enumerate(self.up):
            for j, r in enumerate(blk.resnets):
                h, st = self._resnet_fwd(r, h, f"up.{i}.{j}", yb, zshape, cache_in, cache_out, x_stats=st)
            if blk.upsample is not None:
                h, st = self._upsample_fwd(blk, h)
        if tuple(out.shape[:3]) != tuple(h.shape[:3]):
            raise RuntimeError(f"decoder pass produced {tuple(h.shape)}, expected {tuple(out.shape)}")
        self._norm_conv(h, self.norm_out, self.conv_out, "conv_out", yb, zshape, cache_in, cache_out, out=out, x_stats=st)
        return cache_out

    @staticmethod
    def frame_batches(num_frames: int, fb: int):
        nb = max(num_frames // fb, 1)
        rem = num_frames % fb
        return [(fb * i + (0 if i == 0 else rem), fb * (i + 1) + rem) for i in range(nb)]

    def _frames_out(self, n_latent: int, first: bool) -> int:
        """Output frames of a decoder pass over n_latent latent frames (first pass keeps the leading frame single)."""
        T = n_latent
        for _ in range(self.temporal_compress_level):
            T = (1 + 2 * (T - 1)) if (T > 1 and T % 2 == 1) else (2 * T if T > 1 else 1)
        return T

    def _decode_tile(self, zt: torch.Tensor) -> torch.Tensor:
        """zt [T, hz, wz, 64] -> [T_out, 8hz, 8wz, 16] (first 3 channels = RGB)."""
        T, hz, wz, _ = zt.shape
        batches = self.frame_batches(T, self.config.num_latent_frames_batch_size)
        touts = [self._frames_out(e - s, i == 0) for i, (s, e) in enumerate(batches)]
        s8 = self.spatial_scale
        out = torch.empty((sum(touts), hz * s8, wz * s8, 16), dtype=BF16, device=self.device)
        cache = None
        t0 = 0
        for (s, e), to in zip(batches, touts):
            cache = self._decoder_pass(zt[s:e].contiguous(), cache, out[t0:t0 + to])
            t0 += to
        return out

    # ------------------------------------------------------------------ public decode
    @torch.no_grad()
    def decode(self, z: torch.Tensor, return_dict: bool = True):
        """z [B, C, T, h, w] (already divided by scaling_factor) -> .sample [B, 3, T_out, 8h, 8w] bf16."""
        lib = _lib.load()
        if not isinstance(z, torch.Tensor) or not z.is_cuda:
            raise RuntimeError("decode needs a CUDA tensor (no CPU fallback exists)")
        if z.dim() != 5 or z.shape[1] != self.config.latent_channels:
            raise RuntimeError(f"z must be [B, {self.config.latent_channels}, T, h, w]")
        if self.use_cuda_graph:
            sample = self._decode_graphed(z)
        else:
            sample = self._decode_impl(z)
        if not return_dict:
            return (sample,)
        return DecoderOutput(sample=sample)

    def enable_cuda_graph(self, flag: bool = True) -> None:
        """Replay the whole decode (every tile, frame batch and stream fork/join) as one CUDA graph per latent shape.
        A tiled 49-frame decode is ~9000 launches of 50-500 us kernels; issued from Python the host is the limiter."""
        self.use_cuda_graph = bool(flag)

    def _decode_graphed(self, z: torch.Tensor) -> torch.Tensor:
        key = (tuple(z.shape), z.dtype, self.use_tiling, self.use_slicing, self.tile_streams)
        ent = self._graphs.get(key)
        if ent is None:
            self._decode_impl(z)                                 # eager pass: lazy buffers / streams exist before capture
            torch.cuda.synchronize(self.device)
            static_in = z.clone()
            g = torch.cuda.CUDAGraph()
            self._capturing = True
            try:
                with torch.cuda.graph(g):
                    static_out = self._decode_impl(static_in)
            finally:
                self._capturing = False
            ent = self._graphs[key] = (g, static_in, static_out)
        g, static_in, static_out = ent
        static_in.copy_(z)
        g.replay()
        return static_out.clone()

    def _decode_impl(self, z: torch.Tensor) -> torch.Tensor:
        lib = _lib.load()
        B, Cz, T, H, W = z.shape
        c = self.config
        zcl = torch.zeros((B, T, H, W, self.zc_pad), dtype=BF16, device=self.device)
        zcl[..., :Cz] = z.to(BF16).permute(0, 2, 3, 4, 1)
        tlh, tlw = self.tile_latent_min_height, self.tile_latent_min_width
        tiled = self.use_tiling and (W > tlw or H > tlh)
        if tiled:
            oh, ow = int(tlh * (1 - c.tile_overlap_factor_height)), int(tlw * (1 - c.tile_overlap_factor_width))
            bh = int(self.tile_sample_min_height * c.tile_overlap_factor_height)
            bw = int(self.tile_sample_min_width * c.tile_overlap_factor_width)
            lh, lw = self.tile_sample_min_height - bh, self.tile_sample_min_width - bw
            ys, xs = list(range(0, H, oh)), list(range(0, W, ow))
        else:
            oh = ow = bh = bw = 0
            lh, lw = H * self.spatial_scale, W * self.spatial_scale
            ys, xs = [0], [0]
            tlh, tlw = H, W
        if len(ys) > 4 or len(xs) > 4:
            raise RuntimeError("tiled decode supports at most 4x4 tiles")
        s8 = self.spatial_scale
        samples = []
        for b in range(B):
            tiles = []
            origins = [(y0, x0) for y0 in ys for x0 in xs]
            n_str = max(1, min(self.tile_streams, len(origins)))
            if n_str == 1:
                for (y0, x0) in origins:
                    tiles.append(self._decode_tile(zcl[b, :, y0:y0 + tlh, x0:x0 + tlw].contiguous()))
            else:
                # the low-resolution layers of one 30x45 tile fill only a fraction of the 148 SMs; independent tiles on
                # separate streams overlap them (results are unchanged: every tile's kernels run in order on its stream)
                if self._streams is None or len(self._streams) < n_str:
                    self._streams = [torch.cuda.Stream(device=self.device) for _ in range(n_str)]
                cur = torch.cuda.current_stream()
                for k, (y0, x0) in enumerate(origins):
                    st = self._streams[k % n_str]
                    st.wait_stream(cur)
                    with torch.cuda.stream(st):
                        tiles.append(self._decode_tile(zcl[b, :, y0:y0 + tlh, x0:x0 + tlw].contiguous()))
                for st in self._streams[:n_str]:
                    cur.wait_stream(st)
                if not self._capturing:                          # graph-pool memory is static: nothing to record
                    for t in tiles:
                        t.record_stream(cur)
            To = tiles[0].shape[0]
            out = torch.empty((3, To, H * s8, W * s8), dtype=BF16, device=self.device)
            a = ComposeArgs()
            for k, t in enumerate(tiles):
                a.tiles[k] = t.data_ptr()
            a.rows, a.cols = len(ys), len(xs)
            for i in range(len(ys)):
                a.th[i] = tiles[i * len(xs)].shape[1]
            for j in range(len(xs)):
                a.tw[j] = tiles[j].shape[2]
            a.T, a.H, a.W, a.ldc = To, H * s8, W * s8, 16
            a.blend_h, a.blend_w, a.limit_h, a.limit_w = bh, bw, lh, lw
            a.out = out.data_ptr()
            _lib.check(lib.vgpa_vae_compose_tiles_bf16(C.byref(a), _lib.current_stream()), "vgpa_vae_compose_tiles_bf16")
            samples.append(out)
        return torch.stack(samples, 0)

    # ------------------------------------------------------------------ bookkeeping for bench.py
    def conv_flops(self, T_lat: int, H: int, W: int) -> float:
        """Algorithmic conv FLOPs of an UNTILED decode of [T_lat, H, W] latents (2 * pixels * K * Cout per conv)."""
        rc = self.rc
        fl = 0.0
        Tcur, h, w = T_lat, H, W
        fl += 2.0 * Tcur * h * w * 27 * self.config.latent_channels * rc[0]
        res = lambda ci, co, px: 2.0 * px * 27 * (ci * co + co * co) + (2.0 * px * ci * co if ci != co else 0.0)
        fl += 2 * res(rc[0], rc[0], Tcur * h * w)
        cin = rc[0]
        for i, co in enumerate(rc):
            px = Tcur * h * w
            fl += res(cin, co, px) + self.config.layers_per_block * res(co, co, px)
            if i != len(rc) - 1:
                if i < self.temporal_compress_level:
                    Tcur = 1 + 2 * (Tcur - 1) if Tcur > 1 else 1
                h, w = 2 * h, 2 * w
                fl += 2.0 * Tcur * h * w * 9 * co * co
            cin = co
        fl += 2.0 * Tcur * h * w * 27 * rc[-1] * self.config.out_channels
        return fl


# ====================================================================================================== encoder
class AutoencoderKLOutput(SimpleNamespace):
    """`.latent_dist` like diffusers' AutoencoderKLOutput."""


class DiagonalGaussianDistribution:
    """diffusers' DiagonalGaussianDistribution over moments [B, 2C, T, h, w]: logvar clamped to [-30, 20];
    `sample(generator)` = mean + std * randn, `mode()` = mean (`train/CogVideoX-5B/02_encode.py:113-115`)."""

    def __init__(self, parameters: torch.Tensor):
        self.parameters = parameters
        self.mean, logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def sample(self, generator: torch.Generator | None = None) -> torch.Tensor:
        noise = torch.randn(self.mean.shape, generator=generator, device=self.mean.device, dtype=self.mean.dtype)
        return self.mean + self.std * noise

    def mode(self) -> torch.Tensor:
        return self.mean


class AutoencoderKLCogVideoXEncoder:
    """`vae.encode(x).latent_dist` of diffusers' AutoencoderKLCogVideoX (CogVideoXEncoder3D) on the same kernels as the
    decoder — SURVEY.md §8 row f-4; reference call sites `train/CogVideoX-5B/02_encode.py:108-115` (video latents) and the
    first-frame latent of the I2V pipeline (`generate/CogVideoX-5B-I2V.py`).

    Plain GroupNorm + SiLU runs on vgpa_spatialnorm_apply_bf16 with a constant conditioning pair (y = 1, b = 0: both extra
    roundings are exact); the stride-2 Conv2d of CogVideoXDownsample3D (F.pad(0,1,0,1), no left/top padding) is the
    stride-1 kernel evaluated at the odd pixel centres; the temporal average pool is two frame-slab adds. Frame batching
    (8 sample frames per pass, first pass takes the remainder), conv_cache carry-over and the tiled encode (sample tiles
    240x360, stride 200x288, latent blend 5 / 9, crop 25x36, GroupNorm statistics per tile) follow the library."""

    _conv = AutoencoderKLCogVideoXDecoder._conv
    _conv_call = AutoencoderKLCogVideoXDecoder._conv_call
    _gn_stats = AutoencoderKLCogVideoXDecoder._gn_stats
    _finish_timepad = AutoencoderKLCogVideoXDecoder._finish_timepad
    frame_batches = staticmethod(AutoencoderKLCogVideoXDecoder.frame_batches)
    enable_tiling = AutoencoderKLCogVideoXDecoder.enable_tiling
    disable_tiling = AutoencoderKLCogVideoXDecoder.disable_tiling
    enable_slicing = AutoencoderKLCogVideoXDecoder.enable_slicing
    disable_slicing = AutoencoderKLCogVideoXDecoder.disable_slicing

    def __init__(self, state_dict: dict, config: VAEDecoderConfig | None = None, device="cuda"):
        self.config = c = config or VAEDecoderConfig()
        self.device = torch.device(device)
        self.dtype = BF16
        self.use_tiling = False
        self.use_slicing = False
        self._ws = None
        self.cin_pad = 64
        if c.in_channels > self.cin_pad or 2 * c.latent_channels > 64:
            raise RuntimeError("in_channels > 64 or latent_channels > 32 are not supported")
        self.temporal_compress_level = int(math.log2(c.temporal_compression_ratio))
        boc = tuple(c.block_out_channels)
        sd = state_dict
        self.conv_in = self._conv(sd, "encoder.conv_in.conv", pad_cin_to=self.cin_pad)
        self.down = []
        cin = boc[0]
        for i, co in enumerate(boc):
            blk = SimpleNamespace()
            blk.resnets = [self._gn_resnet(sd, f"encoder.down_blocks.{i}.resnets.{j}", cin if j == 0 else co, co)
                           for j in range(c.layers_per_block)]
            blk.down = None
            if i != len(boc) - 1:
                blk.down = self._conv(sd, f"encoder.down_blocks.{i}.downsamplers.0.conv")
                blk.compress_time = i < self.temporal_compress_level
            self.down.append(blk)
            cin = co
        self.mid = [self._gn_resnet(sd, f"encoder.mid_block.resnets.{j}", boc[-1], boc[-1]) for j in range(2)]
        self.norm_out = self._gn(sd, "encoder.norm_out", boc[-1])
        self.conv_out = self._conv(sd, "encoder.conv_out.conv")
        cmax = max(boc)
        self._ones = torch.ones(cmax, dtype=BF16, device=self.device)
        self._zeros = torch.zeros(cmax, dtype=BF16, device=self.device)
        self.tile_sample_min_height = c.sample_height // 2
        self.tile_sample_min_width = c.sample_width // 2
        self.spatial_scale = 2 ** (len(boc) - 1)
        self.tile_latent_min_height = int(self.tile_sample_min_height / self.spatial_scale)
        self.tile_latent_min_width = int(self.tile_sample_min_width / self.spatial_scale)

    # ------------------------------------------------------------------ construction
    @classmethod
    def random_init(cls, config: VAEDecoderConfig | None = None, seed: int = 6, device="cuda") -> "AutoencoderKLCogVideoXEncoder":
        """Synthetic encoder weights with the diffusers parameter names (no checkpoint is reachable)."""
        cfg = config or VAEDecoderConfig()
        dev = torch.device(device)
        g = torch.Generator(device=dev).manual_seed(seed)
        sd = {}

        def conv(name, co, ci, *k):
            sd[name + ".weight"] = (torch.randn(co, ci, *k, generator=g, device=dev) * (1.0 / (ci * math.prod(k))) ** 0.5).to(BF16)
            sd[name + ".bias"] = (torch.randn(co, generator=g, device=dev) * 0.02).to(BF16)

        def gn(name, ch):
            sd[name + ".weight"] = (1.0 + 0.1 * torch.randn(ch, generator=g, device=dev)).to(BF16)
            sd[name + ".bias"] = (0.05 * torch.randn(ch, generator=g, device=dev)).to(BF16)

        def resnet(name, ci, co):
            gn(name + ".norm1", ci); conv(name + ".conv1.conv", co, ci, 3, 3, 3)
            gn(name + ".norm2", co); conv(name + ".conv2.conv", co, co, 3, 3, 3)
            if ci != co:
                conv(name + ".conv_shortcut", co, ci, 1, 1, 1)

        boc = tuple(cfg.block_out_channels)
        conv("encoder.conv_in.conv", boc[0], cfg.in_channels, 3, 3, 3)
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
        return cls(sd, cfg, device=device)

    def _gn(self, sd, name, ch):
        n = SimpleNamespace()
        if name + ".weight" not in sd:
            raise RuntimeError(f"state dict is missing {name}.weight")
        n.gamma = sd[name + ".weight"].to(device=self.device, dtype=BF16).contiguous()
        n.beta = sd[name + ".bias"].to(device=self.device, dtype=BF16).contiguous()
        n.c = ch
        return n

    def _gn_resnet(self, sd, name, ci, co):
        r = SimpleNamespace()
        r.norm1 = self._gn(sd, name + ".norm1", ci)
        r.conv1 = self._conv(sd, name + ".conv1.conv")
        r.norm2 = self._gn(sd, name + ".norm2", co)
        r.conv2 = self._conv(sd, name + ".conv2.conv")
        r.sc_w = r.sc_b = None
        if ci != co:
            w = sd[name + ".conv_shortcut.weight"]
            if tuple(w.shape[2:]) != (1, 1, 1):
                raise RuntimeError(f"{name}.conv_shortcut: only the 1x1x1 shortcut of the released checkpoints is supported")
            r.sc_w = w.reshape(co, ci).to(device=self.device, dtype=BF16).contiguous()
            r.sc_b = sd[name + ".conv_shortcut.bias"].to(device=self.device, dtype=BF16).contiguous()
        r.ci, r.co = ci, co
        return r

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    # ------------------------------------------------------------------ layers
    def _gn_silu(self, x, norm, out, silu: bool = True):
        """GroupNorm(32, eps 1e-6)(x) (+ SiLU) -> out, x [T,H,W,C] channels-last."""
        lib = _lib.load()
        T, H, W, C_ = x.shape
        stats = self._gn_stats(x, C_)
        a = SpatialNormArgs()
        a.x, a.out, a.mean_rstd = x.data_ptr(), out.data_ptr(), stats.data_ptr()
        a.gamma, a.beta = norm.gamma.data_ptr(), norm.beta.data_ptr()
        a.y_lat, a.b_lat, a.ld_lat = self._ones.data_ptr(), self._zeros.data_ptr(), self._ones.numel()
        a.T, a.H, a.W, a.C, a.groups = T, H, W, C_, self.config.norm_num_groups
        a.Hz = a.Wz = 1
        a.shift = 24                                                   # every pixel reads the single constant row
        a.silu = 1 if silu else 0
        _lib.check(lib.vgpa_spatialnorm_apply_bf16(C.byref(a), _lib.current_stream()), "vgpa_spatialnorm_apply_bf16")

    def _norm_conv(self, x, norm, cv, key, cache_in, cache_out, residual=None):
        T, H, W, C_ = x.shape
        buf = torch.empty((T + 2, H, W, C_), dtype=BF16, device=self.device)
        self._gn_silu(x, norm, buf[2:])
        self._finish_timepad(buf, key, cache_in, cache_out)
        return self._conv_call(cv, buf, T, residual=residual)

    def _resnet_fwd(self, r, x, key, cache_in, cache_out):
        T, H, W, _ = x.shape
        h = self._norm_conv(x, r.norm1, r.conv1, key + ".conv1", cache_in, cache_out)
        res = x
        if r.sc_w is not None:
            res = dense.linear(x.view(-1, r.ci), r.sc_w, r.sc_b).view(T, H, W, r.co)
        return self._norm_conv(h, r.norm2, r.conv2, key + ".conv2", cache_in, cache_out, residual=res)

    def _downsample_fwd(self, blk, x):
        T = x.shape[0]
        if blk.compress_time and T > 1:
            # avg_pool1d(kernel 2, stride 2) over time; an odd frame count keeps the first frame. (a + b) * 0.5 in bf16
            # rounds once, like the pooled fp32 sum rounded to bf16 (the halving is exact).
            if T % 2 == 1:
                x = torch.cat([x[:1], (x[1::2] + x[2::2]) * 0.5], dim=0)
            else:
                x = (x[0::2] + x[1::2]) * 0.5
        full = self._conv_call(blk.down, x.contiguous(), x.shape[0])       # KT = 1: no time padding
        return full[:, 1::2, 1::2].contiguous()                            # stride 2, window starting at even pixels

    def _encoder_pass(self, xt: torch.Tensor, cache_in: dict | None):
        """xt [T, H, W, 64] channels-last padded frames -> (moments [T', h, w, 2*latent], conv_cache)."""
        T = xt.shape[0]
        cache_out: dict = {}
        buf = torch.empty((T + 2,) + tuple(xt.shape[1:]), dtype=BF16, device=self.device)
        buf[2:].copy_(xt)
        self._finish_timepad(buf, "conv_in", cache_in, cache_out)
        h = self._conv_call(self.conv_in, buf, T)
        for i, blk in enumerate(self.down):
            for j, r in enumerate(blk.resnets):
                h = self._resnet_fwd(r, h, f"down.{i}.{j}", cache_in, cache_out)
            if blk.down is not None:
                h = self._downsample_fwd(blk, h)
        for j, r in enumerate(self.mid):
            h = self._resnet_fwd(r, h, f"mid.{j}", cache_in, cache_out)
        return self._norm_conv(h, self.norm_out, self.conv_out, "conv_out", cache_in, cache_out), cache_out

    def _encode_tile(self, xt: torch.Tensor) -> torch.Tensor:
        cache = None
        outs = []
        for (s, e) in self.frame_batches(xt.shape[0], self.config.num_sample_frames_batch_size):
            o, cache = self._encoder_pass(xt[s:e].contiguous(), cache)
            outs.append(o)
        return torch.cat(outs, dim=0)

    @staticmethod
    def _blend(a, b, extent, dim):
        """blend_v / blend_h of the library on channels-last latent tiles [T, h, w, C] (bf16 roundings as in eager)."""
        extent = min(a.shape[dim], b.shape[dim], extent)
        for y in range(extent):
            ia = [slice(None)] * 4
            ib = [slice(None)] * 4
            ia[dim] = a.shape[dim] - extent + y
            ib[dim] = y
            b[tuple(ib)] = a[tuple(ia)] * (1 - y / extent) + b[tuple(ib)] * (y / extent)
        return b

    # ------------------------------------------------------------------ public encode
    @torch.no_grad()
    def encode(self, x: torch.Tensor, return_dict: bool = True):
        """x [B, 3, T, H, W] -> AutoencoderKLOutput(latent_dist=DiagonalGaussianDistribution(moments [B, 2C, T', H/8, W/8]))."""
        _lib.load()
        if not isinstance(x, torch.Tensor) or not x.is_cuda:
            raise RuntimeError("encode needs a CUDA tensor (no CPU fallback exists)")
        c = self.config
        if x.dim() != 5 or x.shape[1] != c.in_channels:
            raise RuntimeError(f"x must be [B, {c.in_channels}, T, H, W]")
        B, Cx, T, H, W = x.shape
        if H % self.spatial_scale or W % self.spatial_scale:
            raise RuntimeError(f"height and width must be multiples of {self.spatial_scale}")
        th, tw = self.tile_sample_min_height, self.tile_sample_min_width
        tiled = self.use_tiling and (W > tw or H > th)
        moments = []
        for b in range(B):
            xcl = torch.zeros((T, H, W, self.cin_pad), dtype=BF16, device=self.device)
            xcl[..., :Cx] = x[b].to(BF16).permute(1, 2, 3, 0)
            if not tiled:
                m = self._encode_tile(xcl)
            else:
                oh, ow = int(th * (1 - c.tile_overlap_factor_height)), int(tw * (1 - c.tile_overlap_factor_width))
                bh = int(self.tile_latent_min_height * c.tile_overlap_factor_height)
                bw = int(self.tile_latent_min_width * c.tile_overlap_factor_width)
                lh, lw = self.tile_latent_min_height - bh, self.tile_latent_min_width - bw
                rows = [[self._encode_tile(xcl[:, i:i + th, j:j + tw].contiguous()) for j in range(0, W, ow)]
                        for i in range(0, H, oh)]
                out_rows = []
                for i, row in enumerate(rows):
                    rr = []
                    for j, tile in enumerate(row):
                        if i > 0:
                            tile = self._blend(rows[i - 1][j], tile, bh, 1)
                        if j > 0:
                            tile = self._blend(row[j - 1], tile, bw, 2)
                        rr.append(tile[:, :lh, :lw])
                    out_rows.append(torch.cat(rr, dim=2))
                m = torch.cat(out_rows, dim=1)
            moments.append(m.permute(3, 0, 1, 2))                          # [2C, T', h, w]
        dist = DiagonalGaussianDistribution(torch.stack(moments, 0))
        if not return_dict:
            return (dist,)
        return AutoencoderKLOutput(latent_dist=dist)
