"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN FILES on seeded inputs.

Run once in the build container (where /root/reference is mounted):
    python tests/golden/make_golden.py
The reference files that run on CPU here are imported by path: metrics/mvcs.py,
train/loss.py, utils/projection_utils.py::project_points, utils/pointcloud_utils.py (with a plyfile
stub), metrics/consistency_score.py::compute_motion_score_vectorized and metrics/mse.py (with
piq/lpips stubs), depth_anything_3/utils/geometry.py (affine_inverse, unproject_depth).
Nothing here is imported at test time; the tests only read the .npz files.
"""
import importlib.util
import math
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, REF / rel)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


sys.path.insert(0, str(REF))
_stub("plyfile", PlyData=object, PlyElement=object)
_stub("piq", ssim=lambda *a, **k: None)
_stub("lpips", LPIPS=object)
metrics_pkg = types.ModuleType("metrics")
metrics_pkg.__path__ = [str(REF / "metrics")]
sys.modules["metrics"] = metrics_pkg
_load("metrics.base", "metrics/base.py")
mvcs_mod = _load("metrics.mvcs", "metrics/mvcs.py")
mse_mod = _load("metrics.mse", "metrics/mse.py")
_load("metrics.lpips", "metrics/lpips.py")
cs_mod = _load("metrics.consistency_score", "metrics/consistency_score.py")
loss_mod = _load("ref_loss", "train/loss.py")
proj_mod = _load("ref_projection", "utils/projection_utils.py")
pc_mod = _load("ref_pointcloud", "utils/pointcloud_utils.py")
geo_mod = _load("ref_geometry", "depth_anything_3/utils/geometry.py")


def poses(T, yaw_deg=0.5, tx=0.02, rows=3):
    E = np.zeros((T, 4, 4), dtype=np.float32)
    for i in range(T):
        a = math.radians(yaw_deg * i)
        E[i] = np.array([[math.cos(a), 0, math.sin(a), tx * i], [0, 1, 0, 0.003 * i],
                         [-math.sin(a), 0, math.cos(a), -0.01 * i], [0, 0, 0, 1]], dtype=np.float32)
    return E[:, :rows]


def intr(T, H, W, f=0.8):
    K = np.zeros((T, 3, 3), dtype=np.float32)
    K[:] = np.array([[f * W, 0, W / 2], [0, f * W, H / 2], [0, 0, 1]], dtype=np.float32)
    return K


def gen_mvcs():
    cases = {}
    # the SURVEY.md §8c KAT (config 1): T=8, 256x256, depth = 2 + 0.5*rand(seed 0), f=200, t_x = 0.02 i
    g = torch.Generator().manual_seed(0)
    T, H, W = 8, 256, 256
    depth = (2.0 + 0.5 * torch.rand(T, H, W, generator=g)).numpy()
    K = np.zeros((T, 3, 3), dtype=np.float32); K[:] = [[200, 0, 128], [0, 200, 128], [0, 0, 1]]
    E = np.zeros((T, 3, 4), dtype=np.float32); E[:, :3, :3] = np.eye(3); E[:, 0, 3] = 0.02 * np.arange(T)
    cases["kat"] = (depth, K, E)
    # yaw + translation, non-square, [T,H,W,1] depth (VGGT layout), 4x4 extrinsics
    g = torch.Generator().manual_seed(1)
    T, H, W = 5, 96, 128
    depth = (1.5 + torch.rand(T, H, W, 1, generator=g) * 2.0).numpy()
    cases["yaw_4x4"] = (depth, intr(T, H, W), poses(T, 2.0, 0.05, rows=4))
    # large motion: part of the image leaves the frame, [T,1,H,W] depth, 4x4 intrinsics
    g = torch.Generator().manual_seed(2)
    T, H, W = 4, 64, 80
    depth = (1.0 + torch.rand(T, 1, H, W, generator=g)).numpy()
    K4 = np.zeros((T, 4, 4), dtype=np.float32); K4[:, :3, :3] = intr(T, H, W, 0.7); K4[:, 3, 3] = 1
    cases["big_motion_k4"] = (depth, K4, poses(T, 8.0, 0.4, rows=3))
    # one pair whose mask is empty (camera 1 looks away): that pair must be skipped
    T, H, W = 4, 32, 32
    depth = (2.0 + 0.25 * torch.rand(T, H, W, generator=g)).numpy()
    E = poses(T, 0.0, 0.01, rows=4).copy()
    E[1, :3, :3] = np.diag([-1.0, 1.0, -1.0]).astype(np.float32)
    cases["empty_pair"] = (depth, intr(T, H, W), E)
    # single frame: no pair -> 0.0
    cases["single"] = (np.full((1, 16, 16), 1.0, dtype=np.float32), intr(1, 16, 16), poses(1, rows=3))
    out = {}
    metric = mvcs_mod.MVCSMetric(device="cpu")
    for name, (d, K, E) in cases.items():
        score = metric.compute(gt=None, rep=None, depths=torch.from_numpy(d), intrinsics=torch.from_numpy(K),
                               extrinsics=torch.from_numpy(E))
        out[f"{name}_depths"], out[f"{name}_K"], out[f"{name}_E"] = d, K, E
        out[f"{name}_score"] = np.float64(score)
        print("mvcs", name, score)
    np.savez_compressed(OUT / "mvcs.npz", **out)


def gen_loss():
    out = {}
    torch.manual_seed(0)
    ts = [torch.randn(2, 13, 16, 60, 90) for _ in range(6)]               # SURVEY §8c KAT
    for tag, kw in {"kat_b1": dict(beta=1.0), "b500": dict(beta=500.0), "smooth": dict(beta=5.0, label_smoothing=0.1),
                    "hinge": dict(beta=5.0, loss_type="hinge")}.items():
        fn = loss_mod.DPOLoss(**kw)
        o = fn(*ts)
        out[tag] = np.array([o.loss.item(), o.reward_margin.item(), o.winner_reward.item(), o.loser_reward.item(),
                             o.accuracy.item()], dtype=np.float64)
        print("loss", tag, out[tag])
    # small case with gradients (autograd of the reference)
    torch.manual_seed(3)
    small = [torch.randn(3, 2, 4, 6, 10) * (0.5 + 0.3 * i) for i in range(6)]
    small[0].requires_grad_(True); small[1].requires_grad_(True)
    o = loss_mod.DPOLoss(beta=2.0)(*small)
    o.loss.backward()
    out["small_inputs"] = np.stack([t.detach().numpy() for t in small])
    out["small_out"] = np.array([o.loss.item(), o.reward_margin.item(), o.winner_reward.item(), o.loser_reward.item(),
                                 o.accuracy.item()], dtype=np.float64)
    out["small_grad_win"], out["small_grad_lose"] = small[0].grad.numpy(), small[1].grad.numpy()
    np.savez_compressed(OUT / "loss.npz", **out)


def gen_reproject():
    # project_points ends with `canvas[v, u] = c` over depth-sorted points (painter's algorithm,
    # projection_utils.py:36,50). With duplicate pixel indices torch's index_put_ is only
    # last-write-wins when deterministic algorithms are on (the default CPU kernel splits the work
    # across threads, the CUDA kernel races), so the goldens are taken in deterministic mode: that is
    # the semantics the reference's sort is written for, and the one videogpa_b200 implements.
    torch.use_deterministic_algorithms(True)
    out = {}
    g = torch.Generator().manual_seed(5)
    T, H, W = 4, 48, 64
    depth = 1.5 + torch.rand(T, H, W, generator=g)
    K, E = intr(T, H, W), poses(T, 3.0, 0.08, rows=3)
    c2w_pts = geo_mod.unproject_depth(depth[None, ..., None], torch.from_numpy(K)[None],
                                      geo_mod.affine_inverse(torch.from_numpy(poses(T, 3.0, 0.08, rows=4)))[None])[0]
    pc = c2w_pts.reshape(-1, 3).contiguous()
    colors = (torch.rand(T * H * W, 3, generator=g) * 255.0)
    frames = [proj_mod.project_points(pc, colors, torch.from_numpy(K[i]), torch.from_numpy(E[i]), H, W).numpy() for i in range(T)]
    out.update(depth=depth.numpy(), K=K, E=E, E4=poses(T, 3.0, 0.08, rows=4), pc=pc.numpy(), colors=colors.numpy(),
               frames=np.stack(frames))
    # colours in [0, 1]: the x255 branch
    colors01 = torch.rand(T * H * W, 3, generator=g)
    out["colors01"] = colors01.numpy()
    out["frames01"] = np.stack([proj_mod.project_points(pc, colors01, torch.from_numpy(K[i]), torch.from_numpy(E[i]), H, W).numpy()
                                for i in range(T)])
    # everything behind the camera: bare background
    Eb = E.copy(); Eb[:, 2, 3] = -100.0
    out["E_behind"] = Eb
    out["frames_behind"] = np.stack([proj_mod.project_points(pc, colors, torch.from_numpy(K[i]), torch.from_numpy(Eb[i]), H, W).numpy()
                                     for i in range(T)])
    print("reproject: nonzero px", (out["frames"].sum(-1) > 0).mean())
    torch.use_deterministic_algorithms(False)
    np.savez_compressed(OUT / "reproject.npz", **out)


def gen_pointcloud():
    out = {}
    g = torch.Generator().manual_seed(7)
    T, H, W = 3, 20, 24
    pts = torch.randn(T, H, W, 3, generator=g)
    conf = 1.0 + torch.rand(T, H, W, generator=g) * 5.0
    conf.view(-1)[::37] = float("nan"); conf.view(-1)[5::41] = 0.0; conf.view(-1)[3::53] = float("inf")
    conf.view(-1)[100:110] = 2.5                                           # ties around a threshold
    images = torch.rand(T, 3, H, W, generator=g)
    out.update(points=pts.numpy(), conf=conf.numpy(), images=images.numpy())
    for th in (0, 30, 50, 97.5):
        v, c = pc_mod.get_colored_pointcloud({"world_points_from_depth": pts, "depth_conf": conf, "images": images},
                                             mode="depth", conf_thres=th)
        out[f"v_{th}"], out[f"c_{th}"] = v.numpy(), c.numpy()
        print("pointcloud th", th, v.shape)
    np.savez_compressed(OUT / "pointcloud.npz", **out)


def gen_consistency():
    out = {}
    E = np.zeros((8, 4, 4), dtype=np.float32)                             # SURVEY §8c KAT: yaw 0.5 deg * i, t_x = 0.02 i
    for i in range(8):
        a = math.radians(0.5 * i)
        E[i] = [[math.cos(a), 0, math.sin(a), 0.02 * i], [0, 1, 0, 0], [-math.sin(a), 0, math.cos(a), 0], [0, 0, 0, 1]]
    out["motion_E"] = E
    out["motion_kat"] = np.float64(cs_mod.compute_motion_score_vectorized(torch.from_numpy(E), device="cpu").item())
    E2 = poses(10, 3.0, 0.1, rows=3)
    out["motion_E2"] = E2
    out["motion_2"] = np.float64(cs_mod.compute_motion_score_vectorized(torch.from_numpy(E2), device="cpu").item())
    out["motion_single"] = np.float64(cs_mod.compute_motion_score_vectorized(torch.from_numpy(E[:1]), device="cpu").item())
    g = torch.Generator().manual_seed(1)                                  # SURVEY §8c KAT for MSE
    gt = torch.rand(4, 3, 32, 32, generator=g)
    rep = torch.rand(4, 3, 32, 32, generator=g) * 2 - 1
    m = mse_mod.MSEMetric()
    out["mse_gt"], out["mse_rep"] = gt.numpy(), rep.numpy()
    out["mse_kat"] = np.float64(m.compute(gt=gt, rep=rep))
    gt_u8 = (torch.rand(4, 32, 32, 3, generator=g) * 255).to(torch.uint8).numpy()   # VGGT path: uint8 THWC numpy GT
    out["mse_gt_u8"] = gt_u8
    out["mse_u8"] = np.float64(m.compute(gt=gt_u8, rep=rep))
    print("consistency", out["motion_kat"], out["motion_2"], out["motion_single"], out["mse_kat"], out["mse_u8"])
    np.savez_compressed(OUT / "consistency.npz", **out)


def gen_geometry():
    g = torch.Generator().manual_seed(9)
    T, H, W = 3, 12, 16
    depth = 1.0 + torch.rand(T, H, W, generator=g)
    K, E4 = intr(T, H, W), poses(T, 4.0, 0.1, rows=4)
    c2w = geo_mod.affine_inverse(torch.from_numpy(E4))
    wp = geo_mod.unproject_depth(depth[None, ..., None], torch.from_numpy(K)[None], c2w[None])[0]
    np.savez_compressed(OUT / "geometry.npz", depth=depth.numpy(), K=K, E4=E4, world_points=wp.numpy())
    print("geometry", wp.shape)


if __name__ == "__main__":
    gen_mvcs(); gen_loss(); gen_reproject(); gen_pointcloud(); gen_consistency(); gen_geometry()
