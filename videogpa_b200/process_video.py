"""Scorer orchestration on the sm_100a kernels — the mirror of `pipelines/process_video.VideoProcessor`
(pipelines/process_video.py:17-201), from the backbone's predictions onwards.

The reference's `process(video_path, thresholds, num_frames)` = decode frames -> run VGGT / Depth-Anything-3 -> for every
confidence threshold: coloured point cloud -> reprojection -> metric dict. The video I/O and the two third-party
backbones (1 B-parameter networks whose weights are not reachable here) are out of scope,
so they are INJECTED: `backbone_fn(frames) -> predictions` returns what `run_model_gpu` / `DepthAnything3.inference` give
(`depth`, `depth_conf`, `extrinsic` w2c [T,3,4], `intrinsic` [T,3,3], `images` [T,3,H,W] in [0,1], optionally
`world_points_from_depth`). Everything after that runs on videogpa_b200 kernels with the reference's semantics:
  * DA3 path: world points = un-projected depth (pipelines/process_video.py:151-156); GT frames = float [T,3,H,W];
  * VGGT path: `world_points_from_depth` comes from the backbone (aliased to the point head, utils/model_utils.py:116-117);
    GT frames = the uint8 THWC numpy frames (pipelines/process_video.py:89-91);
  * metric dispatch by dict key: "Consistency_Score" -> (score, motion_norm), "MVCS" gets depths / intrinsics / extrinsics,
    anything else `compute(gt=, rep=)` (pipelines/process_video.py:168-196); `results["_extrinsic"]` = extrinsics as lists.
`process_batch` is the additive sync-free API: MVCS of many clips in one launch (f-1).
"""
from __future__ import annotations

import os

import torch

from .geometry import batch_reproject, get_colored_pointcloud, unproject_depth
from .metrics import mvcs_batch


class VideoProcessor:
    def __init__(self, metrics: dict, model_name=None, device=None, backbone=None, backbone_fn=None, frame_sampler=None):
        if not torch.cuda.is_available():
            raise RuntimeError("VideoProcessor needs a CUDA device: the scorer kernels have no CPU fallback")
        self.device = device or "cuda"
        self.metrics = metrics
        self.backbone = self._resolve_backbone(backbone, model_name)
        self.model_name = model_name
        self.backbone_fn = backbone_fn
        self.frame_sampler = frame_sampler

    @staticmethod
    def _resolve_backbone(backbone, model_name):
        """pipelines/process_video.py:31-41."""
        if backbone:
            return backbone.lower()
        env_backbone = os.getenv("VIDEO_PROCESSOR_BACKBONE")
        if env_backbone:
            return env_backbone.lower()
        if model_name and "depth-anything" in model_name.lower():
            return "da3"
        return "vggt"

    # ------------------------------------------------------------------ reference entry point
    def process(self, video_path, thresholds, num_frames, save_visuals=False, out_dir=None):
        if self.backbone_fn is None:
            raise RuntimeError("VideoProcessor.process needs an injected `backbone_fn(frames)`: the VGGT / DA3 backbones are "
                               "third-party and out of scope; use process_predictions(...) when predictions are already available")
        sampler = self.frame_sampler
        if sampler is None:                              # pipelines/process_video.py:70 -> utils.video_utils.sample_uniform_frames
            from .video_io import sample_uniform_frames as sampler
        frames = sampler(video_path, n_frames=num_frames)
        preds = self.backbone_fn(frames)
        return self.process_predictions(preds, thresholds, frames_np=frames, save_visuals=save_visuals, out_dir=out_dir)

    def _t(self, x):
        return torch.as_tensor(x).to(self.device)

    def process_predictions(self, preds: dict, thresholds, frames_np=None, save_visuals=False, out_dir=None):
        """The part of `_process_vggt` / `_process_da3` after the backbone forward."""
        extrinsics = self._t(preds["extrinsic"]).float()
        intrinsics = self._t(preds["intrinsic"]).float()
        depths = self._t(preds["depth"]).float()
        images = self._t(preds["images"]).float()
        if self.backbone == "da3":
            if images.dim() == 4 and images.shape[-1] == 3:                       # processed_images are THWC
                images = images.permute(0, 3, 1, 2).contiguous()
            if images.max() > 1.0:
                images = images / 255.0
            conf = self._t(preds["depth_conf"]).float() if preds.get("depth_conf") is not None else torch.ones_like(depths)
            d3 = depths.reshape(depths.shape[0], *depths.shape[-2:]) if depths.dim() != 3 else depths
            world_points = unproject_depth(d3, intrinsics, extrinsics)
            preds = dict(preds, world_points_from_depth=world_points, depth_conf=conf, images=images)
            gt_frames = images
        else:
            if frames_np is None:
                raise RuntimeError("the VGGT path compares against the uint8 THWC frames: pass frames_np")
            gt_frames = frames_np
            preds = dict(preds, images=images)
        height, width = images.shape[-2:]
        results = {}
        for th in thresholds:
            save_path = None
            if save_visuals:
                save_path = os.path.join(out_dir, f"th{th}", "reprojections")
                os.makedirs(save_path, exist_ok=True)
            vertices_3d, colors_rgb = get_colored_pointcloud(preds, mode="depth", conf_thres=th)
            reprojected = batch_reproject(vertices_3d, colors_rgb, intrinsics, extrinsics, height, width, save_path=save_path)
            results[th] = self.compute_metrics(gt_frames, reprojected, extrinsics, intrinsics=intrinsics, depths=depths)
        results["_extrinsic"] = extrinsics.detach().cpu().tolist()
        return results

    def compute_metrics(self, gt_frames, rep_frames, extrinsics, intrinsics=None, depths=None):
        """pipelines/process_video.py:168-196."""
        results = {}
        for name, metric_fn in self.metrics.items():
            if name == "Consistency_Score":
                final_score, motion_norm_val = metric_fn.compute(gt=gt_frames, rep=rep_frames, extrinsics=extrinsics)
                results[name] = final_score
                results["motion_norm"] = motion_norm_val
            elif name == "MVCS":
                results[name] = metric_fn.compute(gt=gt_frames, rep=rep_frames, depths=depths, intrinsics=intrinsics,
                                                  extrinsics=extrinsics)
            else:
                results[name] = metric_fn.compute(gt=gt_frames, rep=rep_frames)
        return results

    # ------------------------------------------------------------------ additive batched API (no per-clip host sync)
    def process_batch(self, depths, intrinsics, extrinsics):
        """MVCS for a batch of clips in one launch: depths [N,T,H,W], intrinsics [N,T,3,3], extrinsics [N,T,3|4,4]
        -> device tensor [N] float64 (no host synchronisation)."""
        return mvcs_batch(self._t(depths).float(), self._t(intrinsics).float(), self._t(extrinsics).float())
