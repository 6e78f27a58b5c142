"""numpy restatement of the reference's geometry-consistency scorer and DPO loss (test infrastructure).

Every function cites the reference file:line it follows. Arithmetic is float32 with one rounding per
operation and the reference's operation order (numpy never contracts to FMA), so the CUDA kernels —
compiled with --fmad=false — can match the index / mask decisions bit for bit. Matrix inverses are
float64 (closed form / Gauss-Jordan) rounded to float32; the reference uses torch.inverse (LU, fp32),
which agrees to ~1e-7 relative.
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32


# ------------------------------------------------------------------------------------------------
# small linear algebra shared with the kernels (same operation order, float64)
# ------------------------------------------------------------------------------------------------
def inv3x3_f64(m: np.ndarray) -> np.ndarray:
    m = [float(x) for x in np.asarray(m, dtype=np.float64).reshape(9)]
    c00 = m[4] * m[8] - m[5] * m[7]
    c01 = m[5] * m[6] - m[3] * m[8]
    c02 = m[3] * m[7] - m[4] * m[6]
    det = m[0] * c00 + m[1] * c01 + m[2] * c02
    with np.errstate(all="ignore"):
        idet = np.float64(1.0) / np.float64(det)
    o = [c00 * idet, (m[2] * m[7] - m[1] * m[8]) * idet, (m[1] * m[5] - m[2] * m[4]) * idet,
         c01 * idet, (m[0] * m[8] - m[2] * m[6]) * idet, (m[2] * m[3] - m[0] * m[5]) * idet,
         c02 * idet, (m[1] * m[6] - m[0] * m[7]) * idet, (m[0] * m[4] - m[1] * m[3]) * idet]
    return np.array(o, dtype=np.float64).reshape(3, 3)


def inv4x4_f64(a: np.ndarray) -> np.ndarray:
    """Gauss-Jordan with partial pivoting on [A | I] — the loop order of mvcs.cu::inv4x4_d."""
    m = [[float(a[r][c]) for c in range(4)] + [1.0 if r == c else 0.0 for c in range(4)] for r in range(4)]
    for col in range(4):
        piv, best = col, abs(m[col][col])
        for r in range(col + 1, 4):
            if abs(m[r][col]) > best:
                best, piv = abs(m[r][col]), r
        if piv != col:
            m[col], m[piv] = m[piv], m[col]
        d = 1.0 / m[col][col]
        m[col] = [x * d for x in m[col]]
        for r in range(4):
            if r == col:
                continue
            f = m[r][col]
            m[r] = [m[r][c] - f * m[col][c] for c in range(8)]
    return np.array([row[4:] for row in m], dtype=np.float64)


def _to_4x4(E: np.ndarray) -> np.ndarray:
    E = np.asarray(E, dtype=np.float64)
    if E.shape[-2:] == (3, 4):
        bottom = np.zeros(E.shape[:-2] + (1, 4))
        bottom[..., 0, 3] = 1.0
        E = np.concatenate([E, bottom], axis=-2)
    return E


# ------------------------------------------------------------------------------------------------
# a-10  MVCS                                                   reference: metrics/mvcs.py:12-114
# ------------------------------------------------------------------------------------------------
def mvcs_pair(depth_i, depth_j, K_i, K_j, E_i, E_j):
    """One (i, j) pair -> (sum of squared errors [float64], mask count). mvcs.py:59-104."""
    H, W = depth_i.shape
    kinv = inv3x3_f64(np.asarray(K_i)[:3, :3]).astype(f32)                 # mvcs.py:65  torch.inverse(K_i)
    Ei, Ej = _to_4x4(E_i), _to_4x4(E_j)                                    # mvcs.py:43-45
    einv = inv4x4_f64(Ei)
    rel = np.zeros((3, 4), dtype=np.float64)
    for r in range(3):
        for c in range(4):
            s = 0.0
            for k in range(4):
                s = s + float(Ej[r, k]) * float(einv[k, c])
            rel[r, c] = s
    R, t = rel[:, :3].astype(f32), rel[:, 3].astype(f32)                   # mvcs.py:70-71
    Kj = np.asarray(K_j, dtype=f32)[:3, :3]
    ys, xs = np.meshgrid(np.arange(H, dtype=f32), np.arange(W, dtype=f32), indexing="ij")   # mvcs.py:50-56
    u, v, d = xs, ys, np.asarray(depth_i, dtype=f32)
    xi = ((kinv[0, 0] * u + kinv[0, 1] * v) + kinv[0, 2]) * d              # mvcs.py:66
    yi = ((kinv[1, 0] * u + kinv[1, 1] * v) + kinv[1, 2]) * d
    zi = ((kinv[2, 0] * u + kinv[2, 1] * v) + kinv[2, 2]) * d
    xj = ((R[0, 0] * xi + R[0, 1] * yi) + R[0, 2] * zi) + t[0]             # mvcs.py:72
    yj = ((R[1, 0] * xi + R[1, 1] * yi) + R[1, 2] * zi) + t[1]
    zj = ((R[2, 0] * xi + R[2, 1] * yi) + R[2, 2] * zi) + t[2]
    hx = (Kj[0, 0] * xj + Kj[0, 1] * yj) + Kj[0, 2] * zj                   # mvcs.py:75
    hy = (Kj[1, 0] * xj + Kj[1, 1] * yj) + Kj[1, 2] * zj
    hz = (Kj[2, 0] * xj + Kj[2, 1] * yj) + Kj[2, 2] * zj
    zc = np.maximum(hz, f32(1e-8))                                         # mvcs.py:79
    with np.errstate(all="ignore"):
        uj, vj = hx / zc, hy / zc                                          # mvcs.py:80-81
        gu = (f32(2.0) * uj) / f32(W - 1) - f32(1.0)                       # mvcs.py:85-86
        gv = (f32(2.0) * vj) / f32(H - 1) - f32(1.0)
        ix = ((gu + f32(1.0)) / f32(2.0)) * f32(W - 1)                     # grid_sample, align_corners=True
        iy = ((gv + f32(1.0)) / f32(2.0)) * f32(H - 1)
    mask = (uj >= 0) & (uj < f32(W)) & (vj >= 0) & (vj < f32(H)) & (zj > 0)   # mvcs.py:99
    if not mask.any():
        return 0.0, 0
    ixm, iym, zjm = ix[mask], iy[mask], zj[mask]
    fx, fy = np.floor(ixm), np.floor(iym)
    x0, y0 = fx.astype(np.int64), fy.astype(np.int64)
    tx, ty = ixm - fx, iym - fy
    dj = np.asarray(depth_j, dtype=f32)

    def fetch(x, y):
        ok = (x >= 0) & (x < W) & (y >= 0) & (y < H)
        out = np.zeros(x.shape, dtype=f32)
        out[ok] = dj[y[ok], x[ok]]
        return out

    one = f32(1.0)
    w_nw, w_ne = (one - tx) * (one - ty), tx * (one - ty)
    w_sw, w_se = (one - tx) * ty, tx * ty
    s = ((fetch(x0, y0) * w_nw + fetch(x0 + 1, y0) * w_ne) + fetch(x0, y0 + 1) * w_sw) + fetch(x0 + 1, y0 + 1) * w_se
    e = s - zjm
    return float(np.sum((e * e).astype(np.float64))), int(mask.sum())     # mvcs.py:103


def _squeeze_depths(depths):
    depths = np.asarray(depths, dtype=f32)
    if depths.ndim == 4:                                                   # mvcs.py:29-33
        if depths.shape[1] == 1:
            depths = depths[:, 0]
        elif depths.shape[3] == 1:
            depths = depths[..., 0]
    return depths


def mvcs(depths, intrinsics, extrinsics, return_pairs: bool = False):
    """MVCSMetric.compute for one clip. Returns the python float score (mvcs.py:108-113)."""
    depths = _squeeze_depths(depths)
    T = depths.shape[0]
    K = np.asarray(intrinsics, dtype=f32)
    E = np.asarray(extrinsics, dtype=f32)
    pair_mse, pair_cnt = [], []
    for i in range(T - 1):                                                 # mvcs.py:59-60
        s, c = mvcs_pair(depths[i], depths[i + 1], K[i], K[i + 1], E[i], E[i + 1])
        pair_mse.append(s / c if c > 0 else 0.0)
        pair_cnt.append(c)
    used = [m for m, c in zip(pair_mse, pair_cnt) if c > 0]                # empty pairs are skipped (mvcs.py:101)
    score = float(np.exp(-1.0 * np.mean(used))) if used else 0.0
    if return_pairs:
        return score, np.array(pair_mse), np.array(pair_cnt, dtype=np.int64)
    return score


# ------------------------------------------------------------------------------------------------
# a-14  reprojection renderer                     reference: utils/projection_utils.py:12-101
# ------------------------------------------------------------------------------------------------
def project_points(pc, colors, K, E, H, W):
    """One view -> uint8 canvas [H, W, 3]. Nearest z wins; exact ties -> lowest point index."""
    pc = np.asarray(pc, dtype=f32)
    colors = np.asarray(colors, dtype=f32)
    K = np.asarray(K, dtype=f32)
    E = np.asarray(E, dtype=f32)
    R, t = E[:3, :3], E[:3, 3]
    x, y, z = pc[:, 0], pc[:, 1], pc[:, 2]
    xc = ((x * R[0, 0] + y * R[0, 1]) + z * R[0, 2]) + t[0]               # projection_utils.py:19
    yc = ((x * R[1, 0] + y * R[1, 1]) + z * R[1, 2]) + t[1]
    zc = ((x * R[2, 0] + y * R[2, 1]) + z * R[2, 2]) + t[2]
    px = (xc * K[0, 0] + yc * K[0, 1]) + zc * K[0, 2]                     # projection_utils.py:20
    py = (xc * K[1, 0] + yc * K[1, 1]) + zc * K[1, 2]
    pz = (xc * K[2, 0] + yc * K[2, 1]) + zc * K[2, 2]
    with np.errstate(all="ignore"):
        den = pz + f32(1e-8)
        uf, vf = np.rint(px / den), np.rint(py / den)                     # projection_utils.py:23-24 (half-to-even)
    valid = (uf >= 0) & (uf < f32(W)) & (vf >= 0) & (vf < f32(H)) & (pz > 0)   # projection_utils.py:26
    canvas = np.zeros((H, W, 3), dtype=np.uint8)                          # bg = (0, 0, 0)
    idx = np.nonzero(valid)[0]
    if idx.size == 0:
        return canvas                                                     # projection_utils.py:33-34
    u, v, zz = uf[idx].astype(np.int64), vf[idx].astype(np.int64), pz[idx]
    # painter's algorithm (projection_utils.py:36,50) == nearest z per pixel; ties -> lowest index
    key = (zz.view(np.uint32).astype(np.uint64) << np.uint64(32)) | idx.astype(np.uint64)
    pix = v * W + u
    order = np.lexsort((key, pix))
    pix_s, key_s = pix[order], key[order]
    first = np.ones(pix_s.shape, dtype=bool)
    first[1:] = pix_s[1:] != pix_s[:-1]
    win_pix = pix_s[first]
    win_idx = (key_s[first] & np.uint64(0xFFFFFFFF)).astype(np.int64)
    c = colors[win_idx]
    cmax = np.max(colors[idx])                                            # max over the valid subset (:45)
    with np.errstate(all="ignore"):
        if cmax <= 1.0:
            c = np.clip(c * f32(255), 0, 255)
        else:
            c = np.clip(c, 0, 255)
    c = np.nan_to_num(c, nan=0.0).astype(np.uint8)                        # .to(torch.uint8) truncates
    canvas.reshape(-1, 3)[win_pix] = c
    return canvas


def batch_reproject(pc, colors, intrinsics, extrinsics, H, W):
    """-> [T, 3, H, W] float32 in [-1, 1] (projection_utils.py:84-101)."""
    T = len(extrinsics)
    if T == 0:
        return np.zeros((0, 3, H, W), dtype=f32)
    frames = [project_points(pc, colors, intrinsics[i], extrinsics[i], H, W) for i in range(T)]
    stack = np.stack(frames).transpose(0, 3, 1, 2).astype(f32)
    return (stack / f32(255.0)) * f32(2.0) - f32(1.0)


# ------------------------------------------------------------------------------------------------
# a-13  coloured point cloud                       reference: utils/pointcloud_utils.py:10-80
# ------------------------------------------------------------------------------------------------
def get_colored_pointcloud(points, conf, images, conf_thres=50.0):
    points = np.asarray(points, dtype=f32).reshape(-1, 3)                 # :31
    images = np.asarray(images, dtype=f32)
    if images.ndim == 4 and images.shape[1] == 3:                         # :38-41
        colors = images.transpose(0, 2, 3, 1)
    else:
        colors = images
    colors = colors.reshape(-1, 3) * f32(255)                             # :42
    vals = np.asarray(conf, dtype=f32).reshape(-1)
    valid = np.isfinite(vals) & (vals > f32(1e-5))                        # :47
    thr = None
    if conf_thres <= 0:                                                   # :50-52
        mask = valid
    else:
        N = int(valid.sum())
        if N == 0:
            mask = valid
        else:
            keep_frac = max(0.0, min(1.0, 1.0 - conf_thres / 100.0))      # :60
            k = max(1, int(np.ceil(N * keep_frac)))                       # :61
            vv = np.sort(vals[valid])[::-1]
            thr = vv[k - 1]                                               # :69-70
            mask = valid & (vals >= thr)                                  # :73
    return points[mask], colors[mask], thr


# ------------------------------------------------------------------------------------------------
# a-12  motion score and range-normalised MSE     reference: metrics/consistency_score.py:8-38, metrics/mse.py:14-54
# ------------------------------------------------------------------------------------------------
def motion_score(extrinsics) -> float:
    E = np.asarray(extrinsics, dtype=f32)
    T = E.shape[0]
    if T < 2:
        return 0.0                                                        # mean of empty -> NaN -> 0 (:36-37)
    R, t = E[:, :3, :3], E[:, :3, 3]
    dt = t[1:] - t[:-1]
    trans = np.sqrt((dt[:, 0] * dt[:, 0] + dt[:, 1] * dt[:, 1]) + dt[:, 2] * dt[:, 2])   # :23
    tr = np.zeros(T - 1, dtype=f32)
    for r in range(3):
        tr = tr + ((R[1:, r, 0] * R[:-1, r, 0] + R[1:, r, 1] * R[:-1, r, 1]) + R[1:, r, 2] * R[:-1, r, 2])   # :26-27
    c = np.clip((tr - f32(1.0)) / f32(2.0), f32(-1.0), f32(1.0))          # :28
    ang = np.arccos(c).astype(f32)                                        # :29
    st = f32(0.0)
    sr = f32(0.0)
    for i in range(T - 1):
        st = f32(st + trans[i])
        sr = f32(sr + ang[i])
    n = f32(T - 1)
    score = f32(st / n) + f32(0.1) * f32(sr / n)                          # :24,30,32
    return 0.0 if np.isnan(score) else float(score)


def _to01(x, is_numpy: bool):
    """MSEMetric._to_tensor_01 (metrics/mse.py:31-54): returns [N, C, H, W] float32."""
    t = np.asarray(x).astype(f32)
    if t.ndim == 3:
        t = t[None]
    if t.shape[-1] == 3:
        t = t.transpose(0, 3, 1, 2)
    if not is_numpy and t.min() < 0:
        t = (t + f32(1.0)) / f32(2.0)
    elif t.max() > 1.0:
        t = t / f32(255.0)
    return t


def mse_metric(gt, rep, gt_is_numpy=False, rep_is_numpy=False) -> float:
    a, b = _to01(gt, gt_is_numpy), _to01(rep, rep_is_numpy)
    d = (a - b).astype(np.float64)
    return float(np.mean(d * d))                                          # mse.py:28


# ------------------------------------------------------------------------------------------------
# DA3 geometry                reference: depth_anything_3/utils/geometry.py:54-59 (affine_inverse), 434-498 (unproject_depth)
# ------------------------------------------------------------------------------------------------
def unproject_depth(depth, intrinsics, extrinsics_w2c):
    depth = np.asarray(depth, dtype=f32)
    T, H, W = depth.shape
    out = np.zeros((T, H, W, 3), dtype=f32)
    ys, xs = np.meshgrid(np.arange(H, dtype=f32), np.arange(W, dtype=f32), indexing="ij")
    for v in range(T):
        kinv = inv3x3_f64(np.asarray(intrinsics[v], dtype=f32)).astype(f32)
        E = np.asarray(extrinsics_w2c[v], dtype=f32)
        C = np.zeros((3, 4), dtype=f32)
        for r in range(3):
            for c in range(3):
                C[r, c] = E[c, r]
            C[r, 3] = -((E[0, r] * E[0, 3] + E[1, r] * E[1, 3]) + E[2, r] * E[2, 3])
        d = depth[v]
        cx = ((kinv[0, 0] * xs + kinv[0, 1] * ys) + kinv[0, 2]) * d
        cy = ((kinv[1, 0] * xs + kinv[1, 1] * ys) + kinv[1, 2]) * d
        cz = ((kinv[2, 0] * xs + kinv[2, 1] * ys) + kinv[2, 2]) * d
        for r in range(3):
            out[v, ..., r] = ((C[r, 0] * cx + C[r, 1] * cy) + C[r, 2] * cz) + C[r, 3]
    return out


# ------------------------------------------------------------------------------------------------
# a-11  8-point fundamental matrix + Sampson       reference: metrics/epipolar.py:194-216 -> kornia (App. A.6) [kornia absent; pinned to OpenCV's FM_8POINT / sampsonDistance, tests/test_oracle_golden.py]
# ------------------------------------------------------------------------------------------------
def _normalize_points(p):
    mu = p.mean(axis=0)
    s = math.sqrt(2.0) / (np.sqrt(((p - mu) ** 2).sum(axis=1)).mean() + 1e-8)
    T = np.array([[s, 0, -s * mu[0]], [0, s, -s * mu[1]], [0, 0, 1.0]])
    return (p - mu) * s, T


def find_fundamental(pts1, pts2):
    p1, p2 = np.asarray(pts1, dtype=np.float64), np.asarray(pts2, dtype=np.float64)
    n1, T1 = _normalize_points(p1)
    n2, T2 = _normalize_points(p2)
    x1, y1, x2, y2 = n1[:, 0], n1[:, 1], n2[:, 0], n2[:, 1]
    X = np.stack([x2 * x1, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, np.ones_like(x1)], axis=1)
    w, V = np.linalg.eigh(X.T @ X)
    Fh = V[:, 0].reshape(3, 3)
    U, S, Vt = np.linalg.svd(Fh)
    S[2] = 0.0
    Fp = U @ np.diag(S) @ Vt
    F = T2.T @ Fp @ T1
    if abs(F[2, 2]) > 1e-8:
        F = F / (F[2, 2] + 1e-8)
    return F.astype(f32)


def sampson_mean_distance(F, pts1, pts2) -> float:
    F = np.asarray(F, dtype=f32)
    p1, p2 = np.asarray(pts1, dtype=f32), np.asarray(pts2, dtype=f32)
    x1, y1, x2, y2 = p1[:, 0], p1[:, 1], p2[:, 0], p2[:, 1]
    l0 = (F[0, 0] * x1 + F[0, 1] * y1) + F[0, 2]
    l1 = (F[1, 0] * x1 + F[1, 1] * y1) + F[1, 2]
    l2 = (F[2, 0] * x1 + F[2, 1] * y1) + F[2, 2]
    m0 = (F[0, 0] * x2 + F[1, 0] * y2) + F[2, 0]
    m1 = (F[0, 1] * x2 + F[1, 1] * y2) + F[2, 1]
    num = (x2 * l0 + y2 * l1) + l2
    den = ((l0 * l0 + l1 * l1) + m0 * m0) + m1 * m1
    d2 = (num * num) / den
    return float(np.mean(np.sqrt(d2 + f32(1e-8)).astype(np.float64)))     # epipolar.py:213, then np.mean (:192)


def epipolar_metric_from_matches(matches) -> float:
    """EpipolarMetric.compute over precomputed per-pair matches [(pts1, pts2) or None] (epipolar.py:161-175)."""
    errs = []
    for m in matches:
        if m is None or len(m[0]) < 8:
            continue
        F = find_fundamental(m[0], m[1])
        if np.isnan(F).any():
            continue
        errs.append(sampson_mean_distance(F, m[0], m[1]))
    return float(np.mean(errs)) if errs else -1.0


# ------------------------------------------------------------------------------------------------
# a-15  DPO loss                                         reference: train/loss.py:53-121
# ------------------------------------------------------------------------------------------------
def dpo_loss(v_win, v_lose, v_win_ref, v_lose_ref, v_win_target, v_lose_target, beta=500.0, label_smoothing=0.0,
             loss_type="sigmoid"):
    def err(a, b):
        d = np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)
        return (d * d).reshape(d.shape[0], -1).mean(axis=1)

    mw, ml = err(v_win, v_win_target), err(v_lose, v_lose_target)          # loss.py:73-74
    rw, rl = err(v_win_ref, v_win_target), err(v_lose_ref, v_lose_target)  # loss.py:76-77
    logits = beta * ((rw - mw) - (rl - ml))                                # loss.py:82-93
    softplus = lambda x: np.maximum(x, 0) + np.log1p(np.exp(-np.abs(x)))
    if loss_type == "sigmoid":
        if label_smoothing > 0:
            tgt = 1.0 - label_smoothing
            loss = np.mean((1 - tgt) * logits + softplus(-logits))        # BCE-with-logits (:97-103)
        else:
            loss = np.mean(softplus(-logits))                              # -logsigmoid (:105)
    elif loss_type == "hinge":
        loss = np.mean(np.maximum(1.0 - logits, 0.0))                      # :108
    else:
        raise ValueError(f"Unknown loss type: {loss_type}")
    wr, lr = -mw, -ml
    return {"loss": float(loss), "reward_margin": float(np.mean(wr - lr)), "winner_reward": float(np.mean(wr)),
            "loser_reward": float(np.mean(lr)), "accuracy": float(np.mean((wr > lr).astype(np.float64))),
            "errors": np.stack([mw, ml, rw, rl])}


# ------------------------------------------------------------------------------------------------
# a-17  frame / pair index rules (bit-exact)      reference: utils/video_utils.py:31-32, metrics/mvcs.py:59-60
# ------------------------------------------------------------------------------------------------
def uniform_frame_indices(total: int, n_frames: int) -> np.ndarray:
    n_eff = min(n_frames, total)
    return np.linspace(0, total - 1, n_eff).astype(int)


def consecutive_pairs(T: int):
    return [(i, i + 1) for i in range(T - 1)]
