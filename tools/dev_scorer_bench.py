"""Dev timing of the scorer kernels at the DA3 production size (10 frames 504x504): point cloud, reprojection, MVCS,
MSE / motion — per clip, device time (not part of the product)."""
import math, sys
import torch
sys.path.insert(0, ".")
from videogpa_b200.geometry import batch_reproject, get_colored_pointcloud, unproject_depth
from videogpa_b200.metrics import MSEMetric, MVCSMetric, compute_motion_score_vectorized, mvcs_batch

T, H, W = 10, 504, 504
g = torch.Generator(device="cuda").manual_seed(0)
depth = 2.0 + 0.5 * torch.rand(T, H, W, device="cuda", generator=g)
conf = 1.0 + torch.rand(T, H, W, device="cuda", generator=g)
images = torch.rand(T, 3, H, W, device="cuda", generator=g)
K = torch.tensor([[0.8 * W, 0, W / 2], [0, 0.8 * W, H / 2], [0, 0, 1]], device="cuda").expand(T, 3, 3).contiguous()
E = torch.zeros(T, 3, 4, device="cuda")
for i in range(T):
    a = math.radians(0.5 * i)
    E[i] = torch.tensor([[math.cos(a), 0, math.sin(a), 0.02 * i], [0, 1, 0, 0], [-math.sin(a), 0, math.cos(a), 0]], device="cuda")

def timeit(name, fn, n=10, bytes_=None):
    for _ in range(3): r = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): r = fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    extra = f"  {bytes_ / ms / 1e6:.0f} GB/s algorithmic" if bytes_ else ""
    print(f"{name:34s} {ms:8.3f} ms{extra}", flush=True)
    return r

N = T * H * W
world = timeit("unproject_depth", lambda: unproject_depth(depth, K, E), bytes_=N * 16)
preds = dict(world_points_from_depth=world, depth_conf=conf, images=images)
for th in (0, 50):
    v, c = timeit(f"get_colored_pointcloud th={th}", lambda: get_colored_pointcloud(preds, mode="depth", conf_thres=th), bytes_=N * 28)
    rep = timeit(f"batch_reproject th={th} ({v.shape[0]} pts)", lambda: batch_reproject(v, c, K, E, H, W), bytes_=v.shape[0] * T * 24 + T * H * W * 12)
timeit("MSEMetric", lambda: MSEMetric().compute(gt=images, rep=rep), bytes_=N * 3 * 8)
timeit("motion score", lambda: compute_motion_score_vectorized(E))
timeit("MVCS (1 clip, float result)", lambda: MVCSMetric().compute(gt=None, rep=None, depths=depth, intrinsics=K, extrinsics=E), bytes_=(T - 1) * H * W * 8)
d128 = depth[None].expand(128, T, H, W).contiguous()
timeit("MVCS batched 128 clips", lambda: mvcs_batch(d128, K[None].expand(128, T, 3, 3).contiguous(), E[None].expand(128, T, 3, 4).contiguous()), n=5, bytes_=128 * (T - 1) * H * W * 8)
