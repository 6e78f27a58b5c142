"""GPU parity: scorer + loss kernels (through the C-ABI via the Python mirror) vs the numpy oracle and
the reference-generated golden vectors. Integer / index / mask work must be bit-exact; floating-point
reductions carry the tolerance written at each assert."""
import math

import numpy as np
import pytest
import torch

from oracle import scorer_np as o

pytestmark = pytest.mark.gpu


def cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


# ------------------------------------------------------------------------------------------------ MVCS
@pytest.mark.parametrize("name", ["kat", "yaw_4x4", "big_motion_k4", "empty_pair", "single"])
def test_mvcs_golden_and_oracle(lib, golden, name):
    from videogpa_b200.metrics import MVCSMetric
    g = golden("mvcs")
    d, K, E = g[name + "_depths"], g[name + "_K"], g[name + "_E"]
    score = MVCSMetric(device="cuda").compute(gt=None, rep=None, depths=cuda(d), intrinsics=cuda(K), extrinsics=cuda(E))
    assert isinstance(score, float)
    assert abs(score - float(g[name + "_score"])) <= 1e-6           # vs the reference file's own output
    # vs the oracle: the mask is bit-exact (next test); the sampled values come from the merged-matrix FMA evaluation and
    # differ from the reference's fp32 operation order by a few ulp per pixel (mvcs.cu header)
    assert abs(score - o.mvcs(d, K, E)) <= 1e-6


@pytest.mark.parametrize("name", ["kat", "yaw_4x4", "big_motion_k4", "empty_pair"])
def test_mvcs_mask_counts_bit_exact(lib, golden, name):
    from videogpa_b200.metrics import mvcs_batch
    g = golden("mvcs")
    d = o._squeeze_depths(g[name + "_depths"])
    K, E = g[name + "_K"], g[name + "_E"]
    _, mse_ref, cnt_ref = o.mvcs(d, K, E, return_pairs=True)
    scores, mse, cnt = mvcs_batch(cuda(d)[None], cuda(K)[None], cuda(E)[None], return_pairs=True)
    assert np.array_equal(cnt.cpu().numpy()[0], cnt_ref)             # mask sizes (integer work): exact
    # per-pair MSE (floating point): same pixels, values within a few fp32 ulp of the reference's operation order
    np.testing.assert_allclose(mse.cpu().numpy()[0], mse_ref, rtol=2e-5, atol=1e-12)


def test_mvcs_batched_equals_single_and_production_size(lib):
    from videogpa_b200.metrics import mvcs_batch
    g = torch.Generator().manual_seed(0)
    N, T, H, W = 3, 10, 504, 504                                      # DA3 production size
    depth = (2.0 + 0.5 * torch.rand(N, T, H, W, generator=g)).numpy()
    K = np.tile(np.array([[0.8 * W, 0, W / 2], [0, 0.8 * W, H / 2], [0, 0, 1]], dtype=np.float32), (N, T, 1, 1))
    E = np.zeros((N, T, 3, 4), dtype=np.float32)
    for n in range(N):
        for i in range(T):
            a = math.radians(0.5 * i * (n + 1))
            E[n, i] = [[math.cos(a), 0, math.sin(a), 0.02 * i], [0, 1, 0, 0], [-math.sin(a), 0, math.cos(a), 0]]
    scores, mse, cnt = mvcs_batch(cuda(depth), cuda(K), cuda(E), return_pairs=True)
    scores = scores.cpu().numpy()
    for n in range(N):
        s1 = mvcs_batch(cuda(depth[n:n + 1]), cuda(K[n:n + 1]), cuda(E[n:n + 1])).cpu().numpy()[0]
        assert abs(s1 - scores[n]) < 1e-12                           # batching does not change a clip's score
    s_ref, mse_ref, cnt_ref = o.mvcs(depth[0], K[0], E[0], return_pairs=True)
    assert np.array_equal(cnt.cpu().numpy()[0], cnt_ref) and abs(scores[0] - s_ref) < 1e-6
    print("production-size per-pair MSE max rel diff vs oracle:", float(np.max(np.abs(mse.cpu().numpy()[0] - mse_ref) / mse_ref)))
    # size-independent property: identical cameras and depth-consistent frames -> error 0 -> score 1
    flat = np.full((1, 4, 64, 64), 3.0, dtype=np.float32)
    I = np.tile(np.eye(4, dtype=np.float32)[:3], (1, 4, 1, 1))
    Kf = np.tile(np.array([[50, 0, 32], [0, 50, 32], [0, 0, 1]], dtype=np.float32), (1, 4, 1, 1))
    assert abs(mvcs_batch(cuda(flat), cuda(Kf), cuda(I)).item() - 1.0) < 1e-12


# ------------------------------------------------------------------------------------------------ reprojection
@pytest.mark.parametrize("frames,colors,E", [("frames", "colors", "E"), ("frames01", "colors01", "E"),
                                             ("frames_behind", "colors", "E_behind"), ("frames", "colors", "E4")])
def test_reproject_bit_exact_vs_reference(lib, golden, frames, colors, E):
    from videogpa_b200.geometry import batch_reproject
    g = golden("reproject")
    H, W = g[frames].shape[1:3]
    out = batch_reproject(cuda(g["pc"]), cuda(g[colors]), cuda(g["K"]), cuda(g[E]), H, W)
    ref = (g[frames].transpose(0, 3, 1, 2).astype(np.float32) / np.float32(255)) * np.float32(2) - np.float32(1)
    assert out.shape == ref.shape and out.dtype == torch.float32
    assert np.array_equal(out.cpu().numpy(), ref)                    # every pixel of every view


def test_reproject_large_random_vs_oracle_and_edges(lib):
    from videogpa_b200.geometry import batch_reproject
    rng = np.random.default_rng(3)
    P, T, H, W = 300_000, 5, 120, 160
    pc = np.stack([rng.uniform(-2, 2, P), rng.uniform(-1.5, 1.5, P), rng.uniform(0.5, 6, P)], 1).astype(np.float32)
    col = rng.uniform(0, 255, (P, 3)).astype(np.float32)
    K = np.tile(np.array([[130, 0, 80], [0, 130, 60], [0, 0, 1]], dtype=np.float32), (T, 1, 1))
    E = np.zeros((T, 3, 4), dtype=np.float32)
    for i in range(T):
        a = math.radians(4.0 * i)
        E[i] = [[math.cos(a), 0, math.sin(a), 0.1 * i], [0, 1, 0, 0.02 * i], [-math.sin(a), 0, math.cos(a), 0.05 * i]]
    out = batch_reproject(pc, col, K, E, H, W).cpu().numpy()          # numpy inputs are accepted like the reference
    assert np.array_equal(out, o.batch_reproject(pc, col, K, E, H, W))
    # exact z ties: duplicate the cloud, lowest index must win -> same picture
    out2 = batch_reproject(np.concatenate([pc, pc]), np.concatenate([col, 255 - col]), K, E, H, W).cpu().numpy()
    assert np.array_equal(out2, out)
    # empty cloud and zero views
    assert (batch_reproject(pc[:0], col[:0], K, E, H, W) == -1).all()
    assert batch_reproject(pc, col, K[:0], E[:0], H, W).shape == (0, 3, H, W)


# ------------------------------------------------------------------------------------------------ point cloud
@pytest.mark.parametrize("th", [0, 30, 50, 97.5])
def test_pointcloud_exact_vs_reference(lib, golden, th):
    from videogpa_b200.geometry import get_colored_pointcloud
    g = golden("pointcloud")
    preds = {"world_points_from_depth": cuda(g["points"]), "depth_conf": cuda(g["conf"]), "images": cuda(g["images"])}
    v, c = get_colored_pointcloud(preds, mode="depth", conf_thres=th)
    assert np.array_equal(v.cpu().numpy(), g[f"v_{th}"]) and np.array_equal(c.cpu().numpy(), g[f"c_{th}"])


def test_pointcloud_production_size_and_edges(lib):
    from videogpa_b200.geometry import get_colored_pointcloud
    g = torch.Generator().manual_seed(11)
    T, H, W = 10, 504, 504                                            # 2.54 M points
    pts = torch.randn(T, H, W, 3, generator=g)
    conf = 1.0 + torch.rand(T, H, W, generator=g).exp()
    conf.view(-1)[::1001] = float("nan")
    img = torch.rand(T, 3, H, W, generator=g)
    for th in (0, 25.0):
        v, c = get_colored_pointcloud({"world_points_from_depth": pts.cuda(), "depth_conf": conf.cuda(), "images": img.cuda()},
                                      mode="depth", conf_thres=th)
        vr, cr, _ = o.get_colored_pointcloud(pts.numpy(), conf.numpy(), img.numpy(), th)
        assert np.array_equal(v.cpu().numpy(), vr) and np.array_equal(c.cpu().numpy(), cr)
    # nothing valid -> empty
    v, c = get_colored_pointcloud({"world_points_from_depth": pts[:1].cuda(), "depth_conf": torch.zeros(1, H, W).cuda(),
                                   "images": img[:1].cuda()}, mode="depth", conf_thres=50)
    assert v.shape == (0, 3) and c.shape == (0, 3)
    # pointmap mode picks world_points / world_points_conf (pointcloud_utils.py:17-19)
    v2, _ = get_colored_pointcloud({"world_points": pts[:1].cuda(), "world_points_conf": conf[:1].cuda(), "images": img[:1].cuda()},
                                   mode="pointmap", conf_thres=0)
    assert v2.shape[0] == int(torch.isfinite(conf[:1]).sum())


# ------------------------------------------------------------------------------------------------ consistency pieces
def test_motion_mse_unproject(lib, golden):
    from videogpa_b200.geometry import unproject_depth
    from videogpa_b200.metrics import Consistency_Score, MSEMetric, compute_motion_score_vectorized
    g = golden("consistency")
    assert abs(float(compute_motion_score_vectorized(cuda(g["motion_E"]))) - float(g["motion_kat"])) < 1e-6
    assert abs(float(compute_motion_score_vectorized(g["motion_E2"])) - float(g["motion_2"])) < 1e-6
    assert float(compute_motion_score_vectorized(cuda(g["motion_E"][:1]))) == 0.0
    m = MSEMetric()
    assert abs(m.compute(gt=cuda(g["mse_gt"]), rep=cuda(g["mse_rep"])) - float(g["mse_kat"])) < 1e-6
    assert abs(m.compute(gt=g["mse_gt_u8"], rep=cuda(g["mse_rep"])) - float(g["mse_u8"])) < 1e-6      # uint8 THWC numpy GT (VGGT path)
    cs = Consistency_Score(lpips_net=lambda a, b: (a - b).abs().mean(dim=(1, 2, 3)), device="cuda")
    val, motion = cs.compute(gt=cuda(g["mse_gt"]), rep=cuda(g["mse_rep"]), extrinsics=cuda(g["motion_E"]))
    assert isinstance(val, float) and abs(motion - float(g["motion_kat"])) < 1e-6 and val > float(g["mse_kat"])
    with pytest.raises(RuntimeError):
        Consistency_Score(device="cuda").compute(gt=cuda(g["mse_gt"]), rep=cuda(g["mse_rep"]), extrinsics=cuda(g["motion_E"]))
    # metrics/lpips.py:56-62: numpy frames are always divided by 255 (a dark uint8 clip with max <= 1 too), and
    # consistency_score.py:68-71 evaluates LPIPS whatever the ratio
    seen = []
    cs2 = Consistency_Score(lpips_net=lambda a, b: (seen.append((a, b)), (a - b).abs().mean(dim=(1, 2, 3)))[1], device="cuda")
    dark = np.zeros((2, 8, 8, 3), dtype=np.uint8); dark[0, 0, 0, 0] = 1
    rep_small = cuda(g["mse_rep"])[:2, :, :8, :8].contiguous()
    cs2.compute(gt=dark, rep=rep_small, extrinsics=cuda(g["motion_E"]), ratio=0)
    assert len(seen) == 1 and abs(float(seen[0][0].max()) - (2.0 / 255.0 - 1.0)) < 1e-7 and float(seen[0][0].min()) == -1.0
    gg = golden("geometry")
    wp = unproject_depth(cuda(gg["depth"]), cuda(gg["K"]), cuda(gg["E4"])).cpu().numpy()
    assert np.abs(wp - gg["world_points"]).max() < 2e-6               # vs DA3 geometry.py
    assert np.array_equal(wp, o.unproject_depth(gg["depth"], gg["K"], gg["E4"]))   # vs oracle: bit-exact


# ------------------------------------------------------------------------------------------------ epipolar
def test_epipolar_vs_oracle(lib):
    from videogpa_b200.metrics import epipolar_from_matches
    rng = np.random.default_rng(0)
    P, M = 6, 300
    a = np.zeros((P, M, 2), dtype=np.float32); b = np.zeros((P, M, 2), dtype=np.float32)
    cnt = np.array([300, 257, 8, 7, 120, 64], dtype=np.int32)
    for p in range(P):
        n = M
        X = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), rng.uniform(3, 6, n)], 1)
        ang = math.radians(2.0 + p)
        R = np.array([[math.cos(ang), 0, math.sin(ang)], [0, 1, 0], [-math.sin(ang), 0, math.cos(ang)]])
        t = np.array([0.2, 0.01 * p, 0.02])
        K = np.array([[200.0, 0, 128], [0, 200.0, 128], [0, 0, 1.0]])
        p1 = (K @ X.T).T; p1 = p1[:, :2] / p1[:, 2:]
        X2 = (R @ X.T).T + t
        p2 = (K @ X2.T).T; p2 = p2[:, :2] / p2[:, 2:] + rng.normal(0, 0.4, (n, 2))
        a[p], b[p] = p1, p2
    Fm, dist, valid = epipolar_from_matches(cuda(a), cuda(b), cuda(cnt))
    dist, valid, Fm = dist.cpu().numpy(), valid.cpu().numpy(), Fm.cpu().numpy()
    assert valid.tolist() == [1, 1, 1, 0, 1, 1]                         # fewer than 8 matches -> invalid
    for p in range(P):
        if not valid[p]:
            continue
        n = cnt[p]
        F = o.find_fundamental(a[p, :n], b[p, :n])
        d = o.sampson_mean_distance(F, a[p, :n], b[p, :n])
        if n > 8:   # minimal 8-point problems are exactly determined; F (hence d) is ill-conditioned there
            assert abs(dist[p] - d) <= 2e-3 * max(d, 1e-3), (p, dist[p], d)
            np.testing.assert_allclose(Fm[p], F, rtol=0, atol=2e-3 * np.abs(F).max())


def test_epipolar_metric_class(lib):
    from videogpa_b200.metrics import EpipolarMetric

    class FixedMatcher:
        def __init__(self):
            self.rng = np.random.default_rng(1)
        def get_matched_points(self, f1, f2):
            p1 = self.rng.uniform(20, 200, (50, 2)).astype(np.float32)
            return p1, p1 + np.float32([3.0, 0.0]), 50, {}

    m = EpipolarMetric(matcher=FixedMatcher())
    frames = torch.rand(4, 3, 32, 32)
    val = m.compute(gt=frames, rep=None)
    assert isinstance(val, float) and 0.0 <= val < 1e-2                 # pure translation: distances at the 1e-4 floor

    class NoMatch:
        def get_matched_points(self, f1, f2):
            return None, None, 0, {}
    assert EpipolarMetric(matcher=NoMatch()).compute(gt=frames, rep=None) == -1.0


# ------------------------------------------------------------------------------------------------ DPO loss
def test_dpo_loss_golden_and_grad(lib, golden):
    from videogpa_b200.loss import DPOLoss, LossOutput, create_loss_strategy
    g = golden("loss")
    torch.manual_seed(0)
    ts = [torch.randn(2, 13, 16, 60, 90) for _ in range(6)]            # the SURVEY §8c KAT tensors
    cu = [t.cuda() for t in ts]
    for tag, kw in {"kat_b1": dict(beta=1.0), "b500": dict(beta=500.0), "smooth": dict(beta=5.0, label_smoothing=0.1),
                    "hinge": dict(beta=5.0, loss_type="hinge")}.items():
        out = DPOLoss(**kw)(*cu)
        assert isinstance(out, LossOutput)
        got = [out.loss.item(), out.reward_margin.item(), out.winner_reward.item(), out.loser_reward.item(), out.accuracy.item()]
        # fp32 reference reductions vs fp64 partial sums; beta multiplies the ~1e-7 noise of the reference's
        # fp32 means into the logit, so the loss tolerance scales with beta
        np.testing.assert_allclose(got, g[tag], rtol=2e-5 * max(1.0, kw["beta"] / 50.0), atol=2e-6)
    x = [torch.from_numpy(a).cuda() for a in g["small_inputs"]]
    x[0].requires_grad_(True); x[1].requires_grad_(True)
    out = create_loss_strategy("dpo", beta=2.0)(*x)
    out.loss.backward()
    np.testing.assert_allclose(out.loss.item(), g["small_out"][0], rtol=1e-5)
    np.testing.assert_allclose(x[0].grad.cpu().numpy(), g["small_grad_win"], rtol=1e-4, atol=1e-7)     # reference autograd
    np.testing.assert_allclose(x[1].grad.cpu().numpy(), g["small_grad_lose"], rtol=1e-4, atol=1e-7)
    # bf16 predictions + fp32 targets (Lightning bf16-mixed) against the oracle on the same rounded values
    mix = [cu[0].bfloat16(), cu[1].bfloat16(), cu[2].bfloat16(), cu[3].bfloat16(), cu[4], cu[5]]
    ref = o.dpo_loss(*[t.float().cpu().numpy() for t in mix], beta=1.0)
    out = DPOLoss(beta=1.0)(*mix)
    assert abs(out.loss.item() - ref["loss"]) < 1e-5 and abs(out.reward_margin.item() - ref["reward_margin"]) < 1e-5
    # sft strategy = plain MSE, differentiable
    p = torch.randn(2, 3, 4, 5, 6, device="cuda", requires_grad=True)
    tgt = torch.randn(2, 3, 4, 5, 6, device="cuda")
    l = create_loss_strategy("sft")(p, tgt).loss
    l.backward()
    assert abs(l.item() - torch.nn.functional.mse_loss(p.detach(), tgt).item()) < 1e-6
    np.testing.assert_allclose(p.grad.cpu().numpy(), (2 * (p.detach() - tgt) / p.numel()).cpu().numpy(), rtol=1e-5, atol=1e-8)
    with pytest.raises(ValueError):
        create_loss_strategy("nope")


def test_cpu_tensor_is_rejected(lib):
    from videogpa_b200.loss import DPOLoss
    with pytest.raises(RuntimeError):
        DPOLoss()(*[torch.randn(1, 2, 2, 2, 2) for _ in range(6)])


# ------------------------------------------------------------------------------------------------ f-1: VideoProcessor
def _synthetic_da3_predictions(T=5, H=48, W=64, seed=0):
    rng = np.random.default_rng(seed)
    depth = (2.0 + 0.5 * rng.random((T, H, W))).astype(np.float32)
    conf = (1.0 + rng.random((T, H, W))).astype(np.float32)
    images = rng.random((T, H, W, 3)).astype(np.float32)                                   # processed_images are THWC
    K = np.tile(np.array([[0.8 * W, 0, W / 2], [0, 0.8 * W, H / 2], [0, 0, 1]], dtype=np.float32), (T, 1, 1))
    E = np.zeros((T, 3, 4), dtype=np.float32)
    for i in range(T):
        a = math.radians(0.7 * i)
        E[i] = np.array([[math.cos(a), 0, math.sin(a), 0.03 * i], [0, 1, 0, 0.0], [-math.sin(a), 0, math.cos(a), 0.01 * i]], dtype=np.float32)
    return dict(depth=depth, depth_conf=conf, images=images, intrinsic=K, extrinsic=E)


def test_video_processor_da3_path_vs_oracle_composition(lib):
    """pipelines/process_video.py:100-196 from the predictions onwards: un-project, point cloud per threshold, reprojection,
    metric dispatch — against the same chain built from the oracle functions."""
    from videogpa_b200.metrics import Consistency_Score, MSEMetric, MVCSMetric
    from videogpa_b200.process_video import VideoProcessor
    preds = _synthetic_da3_predictions()
    lp = lambda a, b: (a - b).abs().mean(dim=(1, 2, 3))                                     # stand-in for the LPIPS-VGG callable
    vp = VideoProcessor({"Consistency_Score": Consistency_Score(lpips_net=lp), "MVCS": MVCSMetric(), "MSE": MSEMetric()},
                        model_name="depth-anything/DA3-LARGE")
    assert vp.backbone == "da3"
    res = vp.process_predictions(preds, thresholds=[0, 40])
    assert set(res) == {0, 40, "_extrinsic"} and np.allclose(np.array(res["_extrinsic"]), preds["extrinsic"])
    T, H, W = preds["depth"].shape
    imgs = np.ascontiguousarray(preds["images"].transpose(0, 3, 1, 2))
    world = o.unproject_depth(preds["depth"], preds["intrinsic"], preds["extrinsic"])
    for th in (0, 40):
        verts, cols, _ = o.get_colored_pointcloud(world.reshape(-1, 3), preds["depth_conf"].reshape(-1), imgs, conf_thres=float(th))
        rep = o.batch_reproject(verts, cols, preds["intrinsic"], preds["extrinsic"], H, W)
        want_mse = o.mse_metric(imgs, rep)
        want_lp = float(np.abs((imgs * 2 - 1) - rep).mean())
        assert abs(res[th]["MSE"] - want_mse) <= 1e-5 * max(1.0, want_mse)
        assert abs(res[th]["Consistency_Score"] - (want_mse + want_lp)) <= 2e-5 * max(1.0, want_mse + want_lp)
        assert abs(res[th]["motion_norm"] - o.motion_score(preds["extrinsic"])) <= 1e-6
        assert abs(res[th]["MVCS"] - o.mvcs(preds["depth"], preds["intrinsic"], preds["extrinsic"])) <= 1e-6
    assert res[0]["MVCS"] == res[40]["MVCS"]                                                # MVCS ignores the threshold
    # additive batched API: one launch for many clips, result stays on the device
    d = torch.from_numpy(preds["depth"]).cuda()[None].repeat(3, 1, 1, 1)
    sc = vp.process_batch(d, torch.from_numpy(preds["intrinsic"])[None].repeat(3, 1, 1, 1), torch.from_numpy(preds["extrinsic"])[None].repeat(3, 1, 1, 1))
    assert sc.is_cuda and sc.shape == (3,) and abs(sc[1].item() - res[0]["MVCS"]) <= 1e-12
    # backbone resolution rules (pipelines/process_video.py:31-41) and the injected-backbone contract
    assert VideoProcessor({}, backbone="VGGT").backbone == "vggt" and VideoProcessor({}).backbone == "vggt"
    with pytest.raises(RuntimeError):
        VideoProcessor({}).process("x.mp4", [0], 10)
