"""Scorer metrics on the sm_100a kernels — drop-in for the reference's metrics/{base,mvcs,epipolar,
consistency_score,mse}.py: same class names, constructor arguments, `compute(*, gt, rep, **kwargs)`
signatures, return types (python floats; -1.0 / None conventions of metrics/epipolar.py:172-216).

Additive batched, sync-free entry points (`mvcs_batch`, `epipolar_from_matches`) return device
tensors so a driver can score many clips per launch and synchronise once.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Optional

import numpy as np
import torch

from . import _lib
from .geometry import _cuda_f32, _e_rows, _workspace


class Metric(ABC):
    """metrics/base.py:4-35."""

    def __init__(self, name: str):
        self.name = name

    @abstractmethod
    def compute(self, *, gt, rep, **kwargs) -> float:
        raise NotImplementedError

    def __call__(self, *args: Any, **kwargs: Any) -> float:
        return self.compute(*args, **kwargs)


# ---------------------------------------------------------------------------------------------- MVCS
def mvcs_batch(depths: torch.Tensor, intrinsics: torch.Tensor, extrinsics: torch.Tensor, return_pairs: bool = False):
    """depths [N, T, H, W], intrinsics [N, T, 3|4, 3|4], extrinsics [N, T, 3|4, 4] (CUDA fp32) -> scores [N] fp64 (device).

    No host synchronisation. With return_pairs also (pair_mse [N, T-1] fp64, pair_cnt [N, T-1] int64).
    """
    lib = _lib.load()
    N, T, H, W = depths.shape
    k_dim, e_rows = intrinsics.shape[-1], extrinsics.shape[-2]
    dev = depths.device
    scores = torch.empty(N, dtype=torch.float64, device=dev)
    pair_mse = torch.zeros((N, max(T - 1, 0)), dtype=torch.float64, device=dev) if return_pairs else None
    pair_cnt = torch.zeros((N, max(T - 1, 0)), dtype=torch.int64, device=dev) if return_pairs else None
    ws = _workspace(lib.vgpa_mvcs_workspace_bytes(N, T, H, W), dev)
    _lib.check(lib.vgpa_mvcs_batch(depths.data_ptr(), intrinsics.data_ptr(), extrinsics.data_ptr(), N, T, H, W, k_dim, e_rows,
                                   ws.data_ptr(), ws.numel(), _lib.ptr(pair_mse), _lib.ptr(pair_cnt), scores.data_ptr(),
                                   _lib.current_stream()), "vgpa_mvcs_batch")
    if return_pairs:
        return scores, pair_mse, pair_cnt
    return scores


class MVCSMetric(Metric):
    """metrics/mvcs.py:6-114."""

    def __init__(self, device="cuda"):
        super().__init__(name="MVCS")
        self.device = device

    def compute(self, *, gt=None, rep=None, depths, intrinsics, extrinsics, **kwargs) -> float:
        d = _cuda_f32(depths, "depths")
        if d.dim() == 4:                                   # [T,1,H,W] or [T,H,W,1]  (mvcs.py:29-33)
            if d.shape[1] == 1:
                d = d[:, 0]
            elif d.shape[3] == 1:
                d = d[..., 0]
        if d.dim() != 3:
            raise RuntimeError(f"depths must be [T,H,W], [T,1,H,W] or [T,H,W,1], got {tuple(d.shape)}")
        K, E = _cuda_f32(intrinsics, "intrinsics"), _cuda_f32(extrinsics, "extrinsics")
        if K.shape[-2:] not in ((3, 3), (4, 4)):
            raise RuntimeError(f"intrinsics must be [T,3,3] or [T,4,4], got {tuple(K.shape)}")
        _e_rows(E)
        scores = mvcs_batch(d.contiguous()[None], K[None].contiguous(), E[None].contiguous())
        return float(scores.item())


# ---------------------------------------------------------------------------------------------- MSE / motion / consistency
def _describe(x):
    """-> (cuda tensor, kind, nhwc, is_numpy, N, C, H, W) following MSEMetric._to_tensor_01's layout rules (mse.py:31-54)."""
    is_numpy = isinstance(x, np.ndarray)
    t = torch.from_numpy(x) if is_numpy else x
    if not isinstance(t, torch.Tensor):
        raise RuntimeError("expected a tensor or numpy array")
    if t.dim() == 3:
        t = t.unsqueeze(0)
    if t.dim() != 4:
        raise RuntimeError(f"expected [N,C,H,W] / [N,H,W,C], got {tuple(t.shape)}")
    nhwc = 1 if t.shape[-1] == 3 else 0
    if t.dtype == torch.uint8:
        kind = 1
        t = t.to("cuda").contiguous()
    else:
        kind = 0
        t = t.to(device="cuda", dtype=torch.float32).contiguous()
    if nhwc:
        N, H, W, Cc = t.shape
    else:
        N, Cc, H, W = t.shape
    return t, kind, nhwc, int(is_numpy), N, Cc, H, W


class MSEMetric(Metric):
    """metrics/mse.py:6-54 (range heuristics included); differing spatial sizes are not supported."""

    def __init__(self):
        super().__init__(name="mse")

    def compute_device(self, *, gt, rep) -> torch.Tensor:
        lib = _lib.load()
        g, gk, gl, gn, N, Cc, H, W = _describe(gt)
        r, rk, rl, rn, N2, C2, H2, W2 = _describe(rep)
        if (N, Cc, H, W) != (N2, C2, H2, W2):
            raise RuntimeError(f"gt {N, Cc, H, W} and rep {N2, C2, H2, W2} must have the same shape (no resize path)")
        out = torch.empty(1, dtype=torch.float32, device=g.device)
        ws = _workspace(lib.vgpa_mse_workspace_bytes(), g.device)
        _lib.check(lib.vgpa_mse_range_normalized(g.data_ptr(), gk, gl, gn, r.data_ptr(), rk, rl, rn, N, Cc, H, W, ws.data_ptr(),
                                                 ws.numel(), out.data_ptr(), _lib.current_stream()), "vgpa_mse_range_normalized")
        return out

    def compute(self, *, gt, rep, **kwargs) -> float:
        return float(self.compute_device(gt=gt, rep=rep).item())


def compute_motion_score_vectorized(extrinsics, device="cuda") -> torch.Tensor:
    """metrics/consistency_score.py:8-38 -> 0-dim CUDA tensor."""
    lib = _lib.load()
    E = _cuda_f32(torch.as_tensor(np.asarray(extrinsics)) if not isinstance(extrinsics, torch.Tensor) else extrinsics, "extrinsics")
    rows = _e_rows(E)
    out = torch.empty(1, dtype=torch.float32, device=E.device)
    _lib.check(lib.vgpa_motion_score(E.data_ptr(), E.shape[0], rows, out.data_ptr(), _lib.current_stream()), "vgpa_motion_score")
    return out[0]


class Consistency_Score(Metric):
    """metrics/consistency_score.py:41-72: MSE + ratio * LPIPS, motion score returned separately.

    The LPIPS-VGG network is third-party weights: pass it as `lpips_net`, a callable
    (gt[-1,1], rep[-1,1]) -> per-frame distances. Without one the LPIPS term raises, it is never
    silently dropped.
    """

    def __init__(self, lpips_net=None, device="cuda"):
        super().__init__("Consistency_Score")
        self.device = device
        self.mse_metric = MSEMetric()
        self.lpips_net = lpips_net

    def _lpips(self, gt, rep) -> float:
        if self.lpips_net is None:
            raise RuntimeError("Consistency_Score needs lpips_net (the LPIPS-VGG weights are not part of this package)")

        def to_pm1(x):                                     # metrics/lpips.py:38-62
            t = torch.from_numpy(x).float() if isinstance(x, np.ndarray) else x.float()
            if t.dim() == 3:
                t = t.unsqueeze(0)
            if t.shape[-1] == 3:
                t = t.permute(0, 3, 1, 2)
            t = t.to(self.device)
            if isinstance(x, np.ndarray):                  # :56-62: numpy frames are always taken as 0-255
                return (t / 255.0) * 2.0 - 1.0
            if t.min() >= 0:
                if t.max() > 1.0:
                    t = t / 255.0
                t = t * 2.0 - 1.0
            return t

        gt_t, rep_t = to_pm1(gt), to_pm1(rep)
        if gt_t.shape[-2:] != rep_t.shape[-2:]:            # :30-31 (torch's own resampling, as in the reference; LPIPS is third-party here)
            rep_t = torch.nn.functional.interpolate(rep_t, size=gt_t.shape[-2:], mode="bilinear", align_corners=False)
        with torch.no_grad():
            return float(self.lpips_net(gt_t, rep_t).mean().item())

    def compute(self, *, gt, rep, extrinsics, ratio=1, **kwargs):
        val_mse = self.mse_metric.compute(gt=gt, rep=rep)
        # the reference always evaluates LPIPS (consistency_score.py:68-71), so a NaN there reaches the score even at ratio 0;
        # only without an injected network (the VGG weights are not part of this package) is ratio = 0 a way to skip the term
        val_lpips = 0.0 if (self.lpips_net is None and ratio == 0) else self._lpips(gt, rep)
        motion = compute_motion_score_vectorized(extrinsics, device=self.device)
        return float(val_mse + ratio * val_lpips), float(motion)


# ---------------------------------------------------------------------------------------------- Epipolar
def epipolar_from_matches(pts1: torch.Tensor, pts2: torch.Tensor, counts: Optional[torch.Tensor] = None):
    """pts1/pts2 [P, M, 2] CUDA fp32, counts [P] int32 -> (F [P,3,3], mean_dist [P], valid [P]) on the device."""
    lib = _lib.load()
    P, M, _ = pts1.shape
    dev = pts1.device
    Fm = torch.empty((P, 3, 3), dtype=torch.float32, device=dev)
    dist = torch.empty(P, dtype=torch.float32, device=dev)
    valid = torch.empty(P, dtype=torch.int32, device=dev)
    _lib.check(lib.vgpa_epipolar_batch(pts1.data_ptr(), pts2.data_ptr(), _lib.ptr(counts), P, M, Fm.data_ptr(), dist.data_ptr(),
                                       valid.data_ptr(), _lib.current_stream()), "vgpa_epipolar_batch")
    return Fm, dist, valid


class SIFTMatcher:
    """metrics/epipolar.py:22-69: cv2 SIFT + BFMatcher(k=2) + Lowe ratio test (CPU, third-party cv2)."""

    def __init__(self, ratio_thresh: float = 0.75, min_matches: int = 20):
        import cv2
        self.cv2 = cv2
        self.sift = cv2.SIFT_create()
        self.bf = cv2.BFMatcher()
        self.ratio_thresh, self.min_matches = ratio_thresh, min_matches

    def get_matched_points(self, frame1: np.ndarray, frame2: np.ndarray):
        cv2 = self.cv2
        g1 = cv2.cvtColor(frame1, cv2.COLOR_RGB2GRAY) if frame1.ndim == 3 else frame1
        g2 = cv2.cvtColor(frame2, cv2.COLOR_RGB2GRAY) if frame2.ndim == 3 else frame2
        kp1, des1 = self.sift.detectAndCompute(g1, None)
        kp2, des2 = self.sift.detectAndCompute(g2, None)
        if des1 is None or des2 is None or len(kp1) < self.min_matches or len(kp2) < self.min_matches:
            return None, None, 0, {}
        good = []
        for pair in self.bf.knnMatch(des1, des2, k=2):
            if len(pair) == 2 and pair[0].distance < self.ratio_thresh * pair[1].distance:
                good.append(pair[0])
        if len(good) < self.min_matches:
            return None, None, len(good), {}
        pts1 = np.float32([kp1[m.queryIdx].pt for m in good])
        pts2 = np.float32([kp2[m.trainIdx].pt for m in good])
        return pts1, pts2, len(good), {}


class EpipolarMetric(Metric):
    """metrics/epipolar.py:140-232. The matcher stays third-party (cv2 SIFT here; LightGlue is not
    installable offline and raises); the fundamental matrix and Sampson distances of ALL frame pairs of
    the clip run in one vgpa_epipolar_batch launch."""

    def __init__(self, descriptor_type: str = "sift", ratio_thresh: float = 0.75, min_matches: int = 20, device: str = None,
                 matcher=None):
        super().__init__(name="Epipolar")
        self.device = device or "cuda"
        self.descriptor_type = descriptor_type
        if matcher is not None:
            self.matcher = matcher
        elif descriptor_type == "sift":
            self.matcher = SIFTMatcher(ratio_thresh=ratio_thresh, min_matches=min_matches)
        elif descriptor_type == "lightglue":
            raise RuntimeError("lightglue is a third-party matcher that is not bundled; pass matcher=<object with get_matched_points>")
        else:
            raise ValueError(f"Unsupported descriptor type: {descriptor_type}")

    def compute(self, *, gt, rep=None, **kwargs) -> float:
        frames = self._to_numpy_thwc(gt)
        matches = []
        for i in range(len(frames) - 1):                    # consecutive pairs (epipolar.py:167-168)
            p1, p2, _, _ = self.matcher.get_matched_points(frames[i], frames[i + 1])
            if p1 is not None and p2 is not None and len(p1) >= 8:
                matches.append((p1, p2))
        if not matches:
            return -1.0
        M = max(len(m[0]) for m in matches)
        P = len(matches)
        a = np.zeros((P, M, 2), dtype=np.float32)
        b = np.zeros((P, M, 2), dtype=np.float32)
        cnt = np.zeros(P, dtype=np.int32)
        for k, (p1, p2) in enumerate(matches):
            a[k, :len(p1)], b[k, :len(p2)], cnt[k] = p1, p2, len(p1)
        _, dist, valid = epipolar_from_matches(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(cnt).cuda())
        dist, valid = dist.cpu().numpy(), valid.cpu().numpy().astype(bool)
        if not valid.any():
            return -1.0
        return float(np.mean(dist[valid].astype(np.float64)))

    @staticmethod
    def _to_numpy_thwc(x):                                   # epipolar.py:218-232
        if isinstance(x, torch.Tensor):
            x = x.detach().cpu().numpy()
        if not isinstance(x, np.ndarray):
            raise ValueError(f"Expected Tensor or numpy array, got {type(x)}")
        if x.ndim == 3:
            x = x[np.newaxis, ...]
        if x.shape[1] == 3 or x.shape[1] == 1:
            x = x.transpose(0, 2, 3, 1)
        if x.min() < 0:
            x = (x + 1.0) * 127.5
        elif x.max() <= 1.0:
            x = x * 255.0
        return np.clip(x, 0, 255).astype(np.uint8)

def sample_frame_indices(total_frames: int, num_frames: int):
    """Frame-index rule of the scorer (utils/video_utils.py:31-32): `np.linspace(0, total-1, n_eff).astype(int)` with
    n_eff = min(num_frames, total). Bit-exact part of the contract: 49 frames, n = 10 ->
    [0, 5, 10, 16, 21, 26, 32, 37, 42, 48]."""
    import numpy as np
    n_eff = min(int(num_frames), int(total_frames))
    return np.linspace(0, total_frames - 1, n_eff).astype(int)


def consecutive_pairs(T: int):
    """Frame pairs scored by MVCS / Epipolar: (i, i+1) (metrics/mvcs.py:59-60, metrics/epipolar.py:167-168)."""
    return [(i, i + 1) for i in range(max(0, T - 1))]
