"""The video-file edge either side of the path: decode an mp4 into the frames the scorer and the
latent encoder take.

Reference: `utils/video_utils.py:10-45` (`center_crop_and_resize`, `sample_uniform_frames`: decord `VideoReader`,
`np.linspace(0, total - 1, n_eff).astype(int)` with `n_eff = min(n, total)`, centre crop to a square, `cv2.resize` to 518 with
`INTER_LINEAR`, RGB uint8 `[T, 518, 518, 3]`) and `train/CogVideoX-5B/02_encode.py:55-63` (`load_video_frames_tensor`: all
frames when the clip is shorter than `num_frames`, else the linspace sample; float `[3, T, H, W]` in [0, 1]).

decord is not installable here; OpenCV (which the reference imports next to decord for the resize) decodes the file. Frames
are read sequentially up to the last requested index and the requested ones are kept, so the selected frame NUMBERS are
exactly the reference's (frame-accurate seeking is not relied on); pixel values are those of OpenCV's decoder, converted
BGR -> RGB. Host code: no GPU work happens here.
"""
from __future__ import annotations

import numpy as np
import torch


def center_crop_and_resize(frame: np.ndarray, size: int = 518) -> np.ndarray:
    """utils/video_utils.py:10-17."""
    import cv2
    h, w = frame.shape[:2]
    side = min(h, w)
    top, left = (h - side) // 2, (w - side) // 2
    return cv2.resize(frame[top:top + side, left:left + side], (size, size), interpolation=cv2.INTER_LINEAR)


def count_frames(video_path: str, decode: bool = False) -> int:
    """Number of frames: the container's count, or (decode=True, or when the container reports none) the number that actually
    decodes — `len(VideoReader(...))` of the reference is the decodable count."""
    import cv2
    cap = cv2.VideoCapture(str(video_path))
    if not cap.isOpened():
        raise RuntimeError(f"cannot read video {video_path}")
    n = 0 if decode else int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    if n <= 0:
        n = 0
        while cap.grab():
            n += 1
    cap.release()
    return n


def _read_sampled(video_path: str, index_rule) -> np.ndarray:
    """Frames at `index_rule(total)`; if the container over-reports its length the indices are re-derived from the decodable count."""
    total = count_frames(video_path)
    if total <= 0:
        raise RuntimeError(f"Video has 0 frames: {video_path}")
    try:
        return read_frames(video_path, index_rule(total))
    except RuntimeError:
        decodable = count_frames(video_path, decode=True)
        if decodable <= 0 or decodable == total:
            raise
        return read_frames(video_path, index_rule(decodable))


def read_frames(video_path: str, indices) -> np.ndarray:
    """RGB uint8 frames `[len(indices), H, W, 3]` at the given (non-decreasing) frame numbers, like
    `VideoReader(video_path).get_batch(indices).asnumpy()`; a repeated index repeats the frame."""
    import cv2
    idx = [int(i) for i in indices]
    if any(b < a for a, b in zip(idx, idx[1:])) or (idx and idx[0] < 0):
        raise RuntimeError("frame indices must be non-negative and non-decreasing")
    cap = cv2.VideoCapture(str(video_path))
    if not cap.isOpened():
        raise RuntimeError(f"cannot read video {video_path}")
    out, pos, want = [], -1, 0
    frame = None
    while want < len(idx):
        while pos < idx[want]:
            ok, bgr = cap.read()
            if not ok:
                cap.release()
                raise RuntimeError(f"video {video_path} ended at frame {pos + 1}, frame {idx[want]} was requested")
            pos += 1
            frame = bgr
        out.append(cv2.cvtColor(frame, cv2.COLOR_BGR2RGB))
        want += 1
    cap.release()
    return np.stack(out, axis=0) if out else np.zeros((0, 0, 0, 3), dtype=np.uint8)


def sample_uniform_frames(video_path: str, n_frames: int = 48, size: int = 518) -> np.ndarray:
    """utils/video_utils.py:20-45 -> `[T, size, size, 3]` uint8 RGB; the `frame_sampler` of process_video.VideoProcessor."""
    from .metrics import sample_frame_indices
    frames = _read_sampled(video_path, lambda total: sample_frame_indices(total, n_frames))
    return np.stack([center_crop_and_resize(f, size) for f in frames], axis=0)


def load_video_frames_tensor(video_path: str, num_frames: int = 49, device=None) -> torch.Tensor:
    """train/CogVideoX-5B/02_encode.py:55-63 -> float `[3, T, H, W]` in [0, 1] (every frame of a clip shorter than
    `num_frames`); the input of encode.encode_video_latent."""
    from .encode import select_frame_indices
    frames = _read_sampled(video_path, lambda total: select_frame_indices(total, num_frames))
    t = torch.from_numpy(frames).float() / 255.0
    t = t.permute(3, 0, 1, 2)
    return t.to(device) if device is not None else t
