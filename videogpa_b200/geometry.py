"""Point-cloud and reprojection utilities on the sm_100a kernels — drop-in for the reference's
utils/pointcloud_utils.py::get_colored_pointcloud, utils/projection_utils.py::batch_reproject and
the DA3 geometry helpers used by pipelines/process_video.py:151-156 (same names, argument meaning
and return types; numpy or tensor inputs are accepted as in utils/projection_utils.py:68-81).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def _cuda_f32(x, name: str) -> torch.Tensor:
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    if not isinstance(x, torch.Tensor):
        raise RuntimeError(f"{name} must be a tensor or numpy array")
    if not torch.cuda.is_available():
        raise RuntimeError("videogpa_b200 needs a CUDA device (no CPU fallback exists)")
    return x.to(device="cuda", dtype=torch.float32).contiguous()


def _workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(((nbytes + 255) // 256) * 256, dtype=torch.uint8, device=device)


def _e_rows(E: torch.Tensor, name="extrinsics") -> int:
    if E.dim() != 3 or E.shape[-1] != 4 or E.shape[-2] not in (3, 4):
        raise RuntimeError(f"{name} must be [T, 3, 4] or [T, 4, 4], got {tuple(E.shape)}")
    return E.shape[-2]


def get_colored_pointcloud(predictions: dict, mode: str = "pointmap", conf_thres=50):
    """-> (vertices [N', 3], colors [N', 3] float in 0..255), both CUDA fp32 (pointcloud_utils.py:10-80)."""
    lib = _lib.load()
    if "pointmap" in mode.lower() and "world_points" in predictions:
        points = predictions["world_points"]
        conf = predictions.get("world_points_conf", None)
    else:
        points = predictions["world_points_from_depth"]
        conf = predictions.get("depth_conf", None)
    points = _cuda_f32(points, "points")
    n = points.numel() // 3
    conf = torch.ones(n, device=points.device, dtype=torch.float32) if conf is None else _cuda_f32(conf, "conf").reshape(-1)
    if conf.numel() != n:
        raise RuntimeError(f"confidence has {conf.numel()} entries for {n} points")
    images = _cuda_f32(predictions["images"], "images")
    nhwc = 0 if (images.dim() == 4 and images.shape[1] == 3) else 1
    if images.numel() != n * 3:
        raise RuntimeError("images must hold one RGB colour per point")
    hw = images.shape[2] * images.shape[3] if nhwc == 0 else max(1, n // max(1, images.shape[0]))
    dev = points.device
    out_v = torch.empty((n, 3), dtype=torch.float32, device=dev)
    out_c = torch.empty((n, 3), dtype=torch.float32, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    ws_bytes = lib.vgpa_pointcloud_workspace_bytes(n)
    ws = _workspace(ws_bytes, dev)
    _lib.check(lib.vgpa_pointcloud_filter(points.data_ptr(), images.data_ptr(), conf.data_ptr(), n, hw, nhwc, float(conf_thres),
                                          ws.data_ptr(), ws.numel(), out_v.data_ptr(), out_c.data_ptr(), count.data_ptr(), None,
                                          _lib.current_stream()), "vgpa_pointcloud_filter")
    k = int(count.item())          # the reference's boolean-mask indexing synchronises here too
    return out_v[:k], out_c[:k]


def batch_reproject(pc, colors, intrinsics, extrinsics, H, W, save_path=None) -> torch.Tensor:
    """-> [T, 3, H, W] fp32 in [-1, 1] on CUDA (projection_utils.py:57-101)."""
    lib = _lib.load()
    pc, colors = _cuda_f32(pc, "pc").reshape(-1, 3), _cuda_f32(colors, "colors").reshape(-1, 3)
    K, E = _cuda_f32(intrinsics, "intrinsics"), _cuda_f32(extrinsics, "extrinsics")
    T = E.shape[0]
    if T == 0:
        return torch.zeros((0, 3, H, W), device="cuda", dtype=torch.float32)
    rows = _e_rows(E)
    if K.shape[-2:] != (3, 3) or K.shape[0] != T:
        raise RuntimeError(f"intrinsics must be [T, 3, 3], got {tuple(K.shape)}")
    if pc.shape[0] != colors.shape[0]:
        raise RuntimeError("pc and colors must have the same number of points")
    out = torch.empty((T, 3, H, W), dtype=torch.float32, device=pc.device)
    ws = _workspace(lib.vgpa_reproject_workspace_bytes(T, H, W), pc.device)
    _lib.check(lib.vgpa_reproject_batch(pc.data_ptr(), colors.data_ptr(), K.data_ptr(), E.data_ptr(), pc.shape[0], T, H, W, rows,
                                        ws.data_ptr(), ws.numel(), out.data_ptr(), _lib.current_stream()), "vgpa_reproject_batch")
    if save_path is not None:
        import os
        import cv2
        os.makedirs(save_path, exist_ok=True)
        u8 = ((out + 1.0) * 127.5).round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).cpu().numpy()
        for i in range(T):
            cv2.imwrite(os.path.join(save_path, f"{i:03d}.png"), cv2.cvtColor(u8[i], cv2.COLOR_RGB2BGR))
    return out


def unproject_depth(depth, intrinsics, extrinsics_w2c) -> torch.Tensor:
    """depth [T, H, W] + K [T, 3, 3] + w2c [T, 3|4, 4] -> world points [T, H, W, 3]
    (= unproject_depth(depth, K, affine_inverse(w2c)), pipelines/process_video.py:151-156)."""
    lib = _lib.load()
    d, K, E = _cuda_f32(depth, "depth"), _cuda_f32(intrinsics, "intrinsics"), _cuda_f32(extrinsics_w2c, "extrinsics")
    if d.dim() != 3:
        raise RuntimeError("depth must be [T, H, W]")
    T, H, W = d.shape
    out = torch.empty((T, H, W, 3), dtype=torch.float32, device=d.device)
    _lib.check(lib.vgpa_unproject_depth(d.data_ptr(), K.data_ptr(), E.data_ptr(), T, H, W, _e_rows(E), out.data_ptr(),
                                        _lib.current_stream()), "vgpa_unproject_depth")
    return out
