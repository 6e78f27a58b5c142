"""CPU: the numpy oracle against golden vectors produced by the REFERENCE's own files
(tests/golden/make_golden.py ran metrics/mvcs.py, train/loss.py, utils/projection_utils.py,
utils/pointcloud_utils.py, metrics/consistency_score.py, metrics/mse.py and DA3 geometry.py in the
build container). This is what pins the oracle."""
import math

import numpy as np
import pytest

from oracle import scorer_np as o


@pytest.mark.parametrize("name", ["kat", "yaw_4x4", "big_motion_k4", "empty_pair", "single"])
def test_mvcs_matches_reference(golden, name):
    g = golden("mvcs")
    s = o.mvcs(g[name + "_depths"], g[name + "_K"], g[name + "_E"])
    assert abs(s - float(g[name + "_score"])) <= 1e-6          # fp32 reference vs fp64-accumulating oracle


def test_mvcs_survey_kat(golden):
    # 0.9537439236729838 measured with the reference file in the survey container
    g = golden("mvcs")
    assert abs(float(g["kat_score"]) - 0.9537439236729838) < 1e-6


def test_mvcs_empty_pairs_are_skipped(golden):
    g = golden("mvcs")
    score, mse, cnt = o.mvcs(g["empty_pair_depths"], g["empty_pair_K"], g["empty_pair_E"], return_pairs=True)
    assert (cnt == 0).sum() >= 1 and (cnt > 0).sum() >= 1
    assert abs(score - math.exp(-np.mean(mse[cnt > 0]))) < 1e-12


@pytest.mark.parametrize("frames,colors,E", [("frames", "colors", "E"), ("frames01", "colors01", "E"),
                                             ("frames_behind", "colors", "E_behind")])
def test_reproject_bit_exact(golden, frames, colors, E):
    g = golden("reproject")
    T = len(g[E])
    H, W = g[frames].shape[1:3]
    mine = np.stack([o.project_points(g["pc"], g[colors], g["K"][i], g[E][i], H, W) for i in range(T)])
    assert np.array_equal(mine, g[frames])                     # uint8 canvases, every pixel


def test_batch_reproject_range(golden):
    g = golden("reproject")
    H, W = g["frames"].shape[1:3]
    out = o.batch_reproject(g["pc"], g["colors"], g["K"], g["E"], H, W)
    ref = (g["frames"].transpose(0, 3, 1, 2).astype(np.float32) / np.float32(255)) * np.float32(2) - np.float32(1)
    assert out.shape == (len(g["E"]), 3, H, W) and np.array_equal(out, ref)
    assert o.batch_reproject(g["pc"], g["colors"], g["K"][:0], g["E"][:0], H, W).shape == (0, 3, H, W)


@pytest.mark.parametrize("th", [0, 30, 50, 97.5])
def test_pointcloud_exact(golden, th):
    g = golden("pointcloud")
    v, c, _ = o.get_colored_pointcloud(g["points"], g["conf"], g["images"], th)
    assert np.array_equal(v, g[f"v_{th}"]) and np.array_equal(c, g[f"c_{th}"])


def test_motion_and_mse(golden):
    g = golden("consistency")
    assert abs(o.motion_score(g["motion_E"]) - float(g["motion_kat"])) < 1e-7
    assert abs(float(g["motion_kat"]) - 0.020872879773378372) < 1e-9            # SURVEY §8c KAT
    assert abs(o.motion_score(g["motion_E2"]) - float(g["motion_2"])) < 1e-6
    assert o.motion_score(g["motion_E"][:1]) == 0.0 == float(g["motion_single"])
    assert abs(o.mse_metric(g["mse_gt"], g["mse_rep"]) - float(g["mse_kat"])) < 1e-6
    assert abs(float(g["mse_kat"]) - 0.16718168556690216) < 1e-9                 # SURVEY §8c KAT
    assert abs(o.mse_metric(g["mse_gt_u8"], g["mse_rep"], gt_is_numpy=True) - float(g["mse_u8"])) < 1e-6


def test_unproject_depth(golden):
    g = golden("geometry")
    assert np.abs(o.unproject_depth(g["depth"], g["K"], g["E4"]) - g["world_points"]).max() < 2e-6


def test_dpo_loss(golden):
    g = golden("loss")
    x = g["small_inputs"]
    r = o.dpo_loss(*x, beta=2.0)
    ref = g["small_out"]
    got = [r["loss"], r["reward_margin"], r["winner_reward"], r["loser_reward"], r["accuracy"]]
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-6)
    # SURVEY §8c KAT: loss 0.6963454484939575, margin -0.002503514289855957, accuracy 0
    assert abs(g["kat_b1"][0] - 0.6963454484939575) < 1e-7 and abs(g["kat_b1"][1] + 0.002503514289855957) < 1e-7


def test_frame_index_rule():
    # 49 frames, n = 10
    assert o.uniform_frame_indices(49, 10).tolist() == [0, 5, 10, 16, 21, 26, 32, 37, 42, 48]
    assert o.uniform_frame_indices(6, 10).tolist() == [0, 1, 2, 3, 4, 5]
    assert o.consecutive_pairs(4) == [(0, 1), (1, 2), (2, 3)]


def test_epipolar_geometry_selfcheck():
    """kornia is not installable here (parity unpinned): check the restatement on exact correspondences."""
    rng = np.random.default_rng(0)
    n, f = 200, 200.0
    X = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), rng.uniform(3, 6, n)], 1)
    a = math.radians(3.0)
    R = np.array([[math.cos(a), 0, math.sin(a)], [0, 1, 0], [-math.sin(a), 0, math.cos(a)]])
    t = np.array([0.2, 0.01, 0.02])
    K = np.array([[f, 0, 128], [0, f, 128], [0, 0, 1.0]])
    p1 = (K @ X.T).T; p1 = p1[:, :2] / p1[:, 2:]
    X2 = (R @ X.T).T + t
    p2 = (K @ X2.T).T; p2 = p2[:, :2] / p2[:, 2:]
    F = o.find_fundamental(p1, p2)
    d = o.sampson_mean_distance(F, p1, p2)
    assert d < 5e-3                                            # sqrt(1e-8) floor = 1e-4 plus fp32 noise
    noisy = p2 + rng.normal(0, 0.5, p2.shape)
    assert 0.05 < o.sampson_mean_distance(o.find_fundamental(p1, noisy), p1, noisy) < 1.0
    assert o.epipolar_metric_from_matches([None, (p1[:5], p2[:5])]) == -1.0


def test_epipolar_oracle_vs_opencv():
    """kornia (the library behind metrics/epipolar.py:194-216) is not installable here, but OpenCV — which the reference also imports —
    ships independent implementations of the same two published algorithms: `cv2.findFundamentalMat(..., FM_8POINT)` (Hartley's
    normalised 8-point: isotropic normalisation, least squares, rank-2 projection) and `cv2.sampsonDistance` (the squared first-order
    geometric error). The oracle's restatement of kornia's `find_fundamental` / `sampson_epipolar_distance` must agree with them:
    F up to scale and sign to 1e-6, the mean distance sqrt(d^2 + 1e-8) to 1e-5."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for p, n in enumerate([300, 120, 64, 30, 9]):
        X = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), rng.uniform(3, 6, n)], 1)
        ang = math.radians(2.0 + p)
        R = np.array([[math.cos(ang), 0, math.sin(ang)], [0, 1, 0], [-math.sin(ang), 0, math.cos(ang)]])
        t = np.array([0.2, 0.01 * p, 0.02])
        K = np.array([[200.0, 0, 128], [0, 200.0, 128], [0, 0, 1.0]])
        p1 = (K @ X.T).T; p1 = (p1[:, :2] / p1[:, 2:]).astype(np.float32)
        X2 = (R @ X.T).T + t
        p2 = (K @ X2.T).T; p2 = (p2[:, :2] / p2[:, 2:] + rng.normal(0, 0.4, (n, 2))).astype(np.float32)
        F = o.find_fundamental(p1, p2).astype(np.float64)
        Fc, _ = cv2.findFundamentalMat(p1.astype(np.float64), p2.astype(np.float64), cv2.FM_8POINT)
        Fn, Fcn = F / np.linalg.norm(F), Fc / np.linalg.norm(Fc)
        if np.sum(Fn * Fcn) < 0:
            Fcn = -Fcn
        assert np.abs(Fn - Fcn).max() < 1e-6, (n, np.abs(Fn - Fcn).max())
        d2 = [cv2.sampsonDistance(np.array([a[0], a[1], 1.0]), np.array([b[0], b[1], 1.0]), F)
              for a, b in zip(p1.astype(np.float64), p2.astype(np.float64))]
        d_cv = float(np.mean(np.sqrt(np.array(d2) + 1e-8)))
        assert abs(o.sampson_mean_distance(F, p1, p2) - d_cv) < 1e-5 * max(d_cv, 1.0), (n, d_cv)
