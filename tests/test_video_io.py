"""CPU tests of the video-file edge: frame selection and layout of videogpa_b200.video_io against the
rules of utils/video_utils.py:20-45 and train/CogVideoX-5B/02_encode.py:55-63 on an mp4 written here with OpenCV."""
import numpy as np
import pytest
import torch

cv2 = pytest.importorskip("cv2")


def _write_clip(path, n, h=96, w=128, fps=8):
    """Frame k is a flat grey level 10 + 4 k with a brighter left half in the red channel (to pin RGB order)."""
    wr = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*"mp4v"), float(fps), (w, h))
    assert wr.isOpened()
    for k in range(n):
        f = np.full((h, w, 3), 10 + 4 * k, dtype=np.uint8)          # BGR
        f[:, : w // 2, 2] = np.minimum(10 + 4 * k + 60, 255)         # red, left half
        wr.write(f)
    wr.release()


def _level(frame_rgb):
    return float(frame_rgb[:, frame_rgb.shape[1] // 2 + 8:, 1].mean())          # green, right half: the grey level


def test_sample_uniform_frames_and_tensor_loader(tmp_path):
    from videogpa_b200 import video_io
    from videogpa_b200.encode import select_frame_indices
    from videogpa_b200.metrics import sample_frame_indices
    p = tmp_path / "clip.mp4"
    _write_clip(p, 49)
    assert video_io.count_frames(str(p)) == 49
    # scorer sampler: 10 of 49 frames -> the reference's index list, centre-cropped to a square and resized to 518
    frames = video_io.sample_uniform_frames(str(p), n_frames=10)
    assert frames.shape == (10, 518, 518, 3) and frames.dtype == np.uint8
    idx = sample_frame_indices(49, 10)
    assert idx.tolist() == [0, 5, 10, 16, 21, 26, 32, 37, 42, 48]
    levels = [_level(f) for f in frames]
    assert all(abs(lv - (10 + 4 * k)) < 3.0 for lv, k in zip(levels, idx)), (levels, idx)       # mp4v is lossy: within 3 grey levels
    assert frames[3][:, :200, 0].mean() > frames[3][:, :200, 2].mean() + 30                       # RGB order: red is brighter on the left
    # more frames requested than exist: every frame once
    assert video_io.sample_uniform_frames(str(p), n_frames=80, size=64).shape == (49, 64, 64, 3)
    # encoder loader: [3, T, H, W] float in [0, 1], all frames of a short clip / linspace sample of a long one
    t = video_io.load_video_frames_tensor(str(p), num_frames=49)
    assert t.shape == (3, 49, 96, 128) and t.dtype == torch.float32 and 0.0 <= float(t.min()) and float(t.max()) <= 1.0
    t13 = video_io.load_video_frames_tensor(str(p), num_frames=13)
    want = select_frame_indices(49, 13)
    got = [255.0 * float(t13[1, i, :, 72:].mean()) for i in range(13)]
    assert all(abs(g - (10 + 4 * k)) < 3.0 for g, k in zip(got, want)), (got, want)
    short = tmp_path / "short.mp4"
    _write_clip(short, 7)
    assert video_io.load_video_frames_tensor(str(short), num_frames=49).shape[1] == 7


def test_read_frames_errors_and_repeats(tmp_path):
    from videogpa_b200 import video_io
    p = tmp_path / "clip.mp4"
    _write_clip(p, 12)
    f = video_io.read_frames(str(p), [0, 0, 5, 11])
    assert f.shape[0] == 4 and np.array_equal(f[0], f[1])
    with pytest.raises(RuntimeError):
        video_io.read_frames(str(p), [3, 2])
    with pytest.raises(RuntimeError):
        video_io.read_frames(str(p), [0, 40])
    with pytest.raises(RuntimeError):
        video_io.count_frames(str(tmp_path / "missing.mp4"))
    crop = video_io.center_crop_and_resize(np.zeros((60, 100, 3), dtype=np.uint8), size=32)
    assert crop.shape == (32, 32, 3)
