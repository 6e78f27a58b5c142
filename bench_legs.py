"""Secondary legs of bench.py: the scorer kernels with their own rooflines and CPU baselines, the GPU-eager torch baseline of
the denoise step, and the multi-GPU splits north_star names (CFG-pair shard, prompt-batch clip + frame gather, DPO step
under DDP). Every leg returns a plain dict and never raises into the primary line (bench.py wraps the calls).

The CPU baselines run the reference's OWN files when oracle/_ref/ (byte-compiled from /root/reference by
oracle/build_ref.py) travelled with the repo: `kind = "reference"`; otherwise the numpy restatement: `kind = "port"`.
This module is measurement harness, not product: it is the one place besides tests/ that may import oracle/.
"""
from __future__ import annotations

import hashlib
import json
import math
import os
import time

ROOT = os.path.dirname(os.path.abspath(__file__))


# ---------------------------------------------------------------------------------------------- helpers
def cuda_ms(fn, iters: int, warmup: int = 2) -> float:
    import torch
    for _ in range(warmup):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def source_sha16(rel: str) -> str:
    with open(os.path.join(ROOT, rel), "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()[:16]


def measured_traffic(kernel: str):
    """dram bytes per launch of `kernel` from this round's ncu capture (profiles/r02_traffic.json), or None when the capture
    was taken from another version of the kernel's source file (the sha of the .cu is stored beside the number)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            rec = json.load(f)[kernel]
        if rec.get("source_sha16") and rec["source_sha16"] != source_sha16(rec["source"]):
            return None
        return rec["dram_bytes_per_launch"]
    except Exception:
        return None


def scorer_inputs(dev, N: int, T: int, H: int, W: int):
    """SURVEY.md §8d synthetic scorer input: depth = 2 + 0.5 U[0,1) (seed 0), K = [[0.8W,0,W/2],[0,0.8W,H/2],[0,0,1]], E_i = yaw 0.5 deg * i,
    t_x = 0.02 i."""
    import torch
    gd = torch.Generator(device=dev).manual_seed(0)
    depth = 2.0 + 0.5 * torch.rand(N, T, H, W, generator=gd, device=dev)
    K = torch.tensor([[0.8 * W, 0, W / 2], [0, 0.8 * W, H / 2], [0, 0, 1]], device=dev).expand(N, T, 3, 3).contiguous()
    E = torch.zeros(N, T, 3, 4, device=dev)
    for i in range(T):
        a = math.radians(0.5 * i)
        E[:, i] = torch.tensor([[math.cos(a), 0, math.sin(a), 0.02 * i], [0, 1, 0, 0], [-math.sin(a), 0, math.cos(a), 0]], device=dev)
    return depth, K, E


# ---------------------------------------------------------------------------------------------- scorer: CPU arms
def cpu_scorer_baselines(budget_s: float = 6.0) -> dict:
    """MVCSMetric.compute (metrics/mvcs.py:12-114), project_points (utils/projection_utils.py:12-51) and DPOLoss.forward
    (train/loss.py:53-121) timed on the host cores on a bounded sample of the GPU legs' workloads."""
    import numpy as np
    import torch
    from oracle import ref_loader
    from oracle import scorer_np as S
    ref = ref_loader.load()
    kind = "reference" if ref is not None else "port"
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    out = {}
    T, H, W = 10, 504, 504
    depth, K, E = scorer_inputs("cpu", 1, T, H, W)
    # --- MVCS: whole clips until the budget is spent
    n, t0 = 0, time.perf_counter()
    while True:
        if ref is not None:
            ref.MVCSMetric(device="cpu").compute(gt=None, rep=None, depths=depth[0], intrinsics=K[0], extrinsics=E[0])
        else:
            S.mvcs(depth[0].numpy(), K[0].numpy(), E[0].numpy())
        n += 1
        if time.perf_counter() - t0 > budget_s or n >= 2000:
            break
    dt = time.perf_counter() - t0
    out["mvcs"] = {"value": n / dt, "unit": "scores/s", "cores": cores, "kind": kind,
                   "sample": f"{n} clip(s) of 10 frames 504x504 through " + ("the reference's metrics/mvcs.py (oracle/_ref)" if ref is not None else "oracle/scorer_np.py") + ", torch/numpy CPU"}
    # --- project_points: views of the 10 x 504 x 504 point cloud of one clip
    g = torch.Generator().manual_seed(1)
    npts = T * H * W
    pc = torch.randn(npts, 3, generator=g) * torch.tensor([1.0, 1.0, 0.3]) + torch.tensor([0.0, 0.0, 2.5])
    col = torch.rand(npts, 3, generator=g)
    n, t0 = 0, time.perf_counter()
    while True:
        i = n % T
        if ref is not None:
            ref.project_points(pc, col, K[0, i], E[0, i], H, W)
        else:
            S.project_points(pc.numpy(), col.numpy(), K[0, i].numpy(), E[0, i].numpy(), H, W)
        n += 1
        if time.perf_counter() - t0 > budget_s or n >= 500:
            break
    dt = time.perf_counter() - t0
    out["reproject"] = {"value": n / dt, "unit": "views/s", "cores": cores, "kind": kind,
                        "sample": f"{n} view(s) of a {npts}-point cloud onto 504x504 through " + ("the reference's project_points (oracle/_ref)" if ref is not None else "oracle/scorer_np.py")}
    # --- DPO loss term
    vs = [torch.randn(1, 13, 16, 60, 90, generator=g) for _ in range(6)]
    n, t0 = 0, time.perf_counter()
    while True:
        if ref is not None:
            ref.DPOLoss(beta=1.0)(*vs)
        else:
            S.dpo_loss(*[v.numpy() for v in vs], beta=1.0)
        n += 1
        if time.perf_counter() - t0 > budget_s / 3 or n >= 20000:
            break
    dt = time.perf_counter() - t0
    out["dpo_loss"] = {"value": n / dt, "unit": "losses/s", "cores": cores, "kind": kind,
                       "sample": f"{n} evaluation(s) of the loss on six [1,13,16,60,90] fp32 tensors through " + ("the reference's train/loss.py (oracle/_ref)" if ref is not None else "oracle/scorer_np.py")}
    return out


# ---------------------------------------------------------------------------------------------- scorer: GPU legs
def gpu_scorer_legs(dev, peak_hbm: float, cpu: dict | None) -> dict:
    import torch
    from videogpa_b200.geometry import batch_reproject
    from videogpa_b200.loss import DPOLoss
    from videogpa_b200.metrics import mvcs_batch
    out = {}
    # --- MVCS (BASELINE.json's second headline metric), batched: 128 clips per launch at the DA3 production size
    N, T, H, W = 128, 10, 504, 504
    depth, K, E = scorer_inputs(dev, N, T, H, W)
    sc = mvcs_batch(depth, K, E)
    ms = cuda_ms(lambda: mvcs_batch(depth, K, E), 10)
    alg = N * (T - 1) * H * W * 8                               # 8 B per pixel per pair
    gbs = alg / (ms / 1000.0) / 1e9
    out["mvcs"] = {"metric": "MVCS scores/sec", "value": N / (ms / 1000.0), "unit": "scores/s", "clips_per_launch": N,
                   "workload": "10 frames 504x504 (DA3 production size), synthetic depth/pose", "ms_per_launch": ms, "score0": float(sc[0].item()),
                   "roofline": {"kernel": "mvcs_pairs_kernel", "bound": "hbm", "achieved": gbs, "peak": peak_hbm, "unit": "GB/s", "frac": gbs / peak_hbm,
                                "traffic": measured_traffic("mvcs_pairs_kernel"), "bytes_per_launch": alg,
                                "note": "8 B per pixel per pair x 128 clips x 9 pairs x 504^2; time = whole vgpa_mvcs_batch call (prepare + pairs + finalize)"},
                   "cpu_baseline": (cpu or {}).get("mvcs")}
    del depth
    # --- batch_reproject: the point cloud of one clip rendered into its 10 views (z-buffer)
    g = torch.Generator(device=dev).manual_seed(1)
    npts = T * H * W
    pc = torch.randn(npts, 3, generator=g, device=dev) * torch.tensor([1.0, 1.0, 0.3], device=dev) + torch.tensor([0.0, 0.0, 2.5], device=dev)
    col = torch.rand(npts, 3, generator=g, device=dev)
    ms = cuda_ms(lambda: batch_reproject(pc, col, K[0], E[0], H, W), 10)
    alg = T * (npts * 24 + H * W * 3)                           # 24 B/point/view read + 3 B/pixel written
    gbs = alg / (ms / 1000.0) / 1e9
    out["reproject"] = {"metric": "reprojected views/sec", "value": T / (ms / 1000.0), "unit": "views/s", "ms_per_clip": ms,
                        "workload": f"{npts} points x 10 views at 504x504",
                        "roofline": {"kernel": "reproject_zbuffer_kernel + reproject_resolve_kernel", "bound": "hbm", "achieved": gbs, "peak": peak_hbm, "unit": "GB/s",
                                     "frac": gbs / peak_hbm, "traffic": measured_traffic("reproject"), "bytes_per_launch": alg},
                        "cpu_baseline": (cpu or {}).get("reproject")}
    # --- DPO loss term
    vs = [torch.randn(1, 13, 16, 60, 90, generator=g, device=dev) for _ in range(6)]
    crit = DPOLoss(beta=1.0)
    ms = cuda_ms(lambda: crit(*vs), 20)
    alg = 6 * vs[0].numel() * 4
    gbs = alg / (ms / 1000.0) / 1e9
    out["dpo_loss"] = {"metric": "DPO loss evaluations/sec", "value": 1000.0 / ms, "unit": "losses/s", "ms_per_call": ms,
                       "workload": "six [1,13,16,60,90] fp32 tensors (one preference pair)",
                       "roofline": {"kernel": "dpo_loss_forward kernels", "bound": "hbm", "achieved": gbs, "peak": peak_hbm, "unit": "GB/s", "frac": gbs / peak_hbm,
                                    "traffic": None, "bytes_per_launch": alg, "note": "27 MB per call: launch-latency bound at one pair per call"},
                       "cpu_baseline": (cpu or {}).get("dpo_loss")}
    return out


# ---------------------------------------------------------------------------------------------- GPU-eager torch baseline
def gpu_eager_baseline(dev, steps: int = 2) -> dict:
    """The same guided DDIM step through plain torch on the same GPU: oracle/dit_torch.py with bf16 weights on CUDA, i.e.
    cuBLAS linears, torch's SDPA backend for attention and eager bf16 elementwise kernels (what the reference's diffusers
    pipeline executes in bf16). Random weights of the same shapes; the step is timed, not compared."""
    import torch
    from oracle import dit_torch as O
    BF = torch.bfloat16
    ocfg = O.DiTConfig()
    sd = O.random_state_dict(ocfg, seed=1234, dtype=BF, device=dev)
    g = torch.Generator(device=dev).manual_seed(42)
    lat = torch.randn(1, 13, 16, 60, 90, generator=g, device=dev).to(BF)
    pe = torch.randn(2, 226, 4096, generator=g, device=dev).to(BF)
    rope = tuple(r.to(dev) for r in O.rope_3d(ocfg, 13, 60, 90))
    ac = O.cogvideox_alphas_cumprod()
    ts = [int(t) for t in O.trailing_timesteps(50)]

    def step(i, x):
        tt = torch.tensor([ts[i], ts[i]], device=dev)
        pred = O.transformer_forward(sd, ocfg, torch.cat([x, x]), pe, tt, rope)
        return O.ddim_step(ac, ts[i], ts[i + 1], x.float(), O.cfg_combine(pred, 6.0)).to(BF)

    with torch.no_grad():
        x = step(0, lat)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            x = step(1 + i, x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del sd
    torch.cuda.empty_cache()
    return {"value": 17550 / (ms / 1000.0), "unit": "tokens/s", "ms_per_step": ms, "steps": steps,
            "what": "oracle/dit_torch.py on CUDA in bf16: cuBLAS linears + torch SDPA + eager elementwise kernels, same CFG-pair DDIM step, random weights",
            "torch": torch.__version__, "finite": bool(torch.isfinite(x.float()).all().item())}


# ---------------------------------------------------------------------------------------------- multi-GPU legs
def _max_over_ranks(vals, dev):
    import torch
    import torch.distributed as dist
    t = torch.tensor(vals, dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def multi_gpu_legs(args, rank: int, world: int, dev, model, pipe, single_gpu_step_ms: float, out: dict | None = None) -> dict:
    """world >= 2 (even): (i) CFG-pair shard of the CogVideoX step and of the Wan2.2 step, (ii) BASELINE.json configs[2]
    end to end (50 steps + VAE decode + frame gather), (iii) the DPO step under DDP. All times are CUDA events, max over ranks."""
    import torch
    import torch.distributed as dist
    from videogpa_b200.parallel import (BucketedGradReducer, CfgPairGroup, CfgPairPeerGroup, gather_frames)
    BF = torch.bfloat16
    out = {} if out is None else out                     # filled leg by leg, so a watchdog can report the legs that finished

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    # One 2-rank NCCL group per CFG pair, created once and shared by every leg. The peer-memory group (exchange fused into the
    # guidance + scheduler kernel) serves the CogVideoX leg; the Wan leg exchanges through NCCL on the same group: on 2 B200s the
    # Wan leg never returned when it built a second peer group after the first (cause not isolated; profiles/r02_multi_gpu.md).
    pair_nccl = CfgPairGroup(rank, world)

    def make_group(peer: bool):
        if peer and world == 2:                          # with several pairs (4 GPUs) the peer-memory leg never returned: one pair only

            try:
                return CfgPairPeerGroup(rank, world, share=pair_nccl), "peer memory (exchange fused into the guidance + scheduler kernel over NVLink)"
            except Exception as ex:                      # noqa: BLE001
                return pair_nccl, f"NCCL all-gather (peer-memory path unavailable: {type(ex).__name__})"
        return pair_nccl, ("NCCL all-gather inside the pair, then the fused guidance + scheduler kernel"
                           + (" (the peer-memory exchange is only used with one pair: it hung with two pairs on 4 GPUs)" if peer else ""))

    # Order: the legs that only use NCCL collectives first (clip + gather, DDP step, Wan pair), the peer-memory leg last, so that a
    # problem in the symmetric-memory path cannot take the others with it (bench.py bounds the whole sequence with a watchdog).
    # ---- (ii) BASELINE.json configs[2]: CogVideoX-5B-I2V, one prompt per GPU, 50 steps + VAE decode + gather of the uint8 frames
    try:
        from videogpa_b200.pipeline import CogVideoXDenoisePipeline
        from videogpa_b200.schedulers import CogVideoXDDIMScheduler
        from videogpa_b200.transformer import CogVideoXTransformer3D, TransformerConfig
        from videogpa_b200.vae import AutoencoderKLCogVideoXDecoder, VAEDecoderConfig
        steps = int(args.e2e_steps)
        icfg = TransformerConfig.cogvideox_5b_i2v()
        if args.layers:
            icfg.num_layers = args.layers
        imodel = CogVideoXTransformer3D.random_init(icfg, seed=77, device=dev)
        dec = AutoencoderKLCogVideoXDecoder.random_init(VAEDecoderConfig(), seed=5, device=dev)
        dec.enable_tiling(); dec.enable_slicing()
        ipipe = CogVideoXDenoisePipeline(imodel, CogVideoXDDIMScheduler(), vae=dec)
        gi = torch.Generator(device=dev).manual_seed(300 + rank)
        prompt = torch.randn(1, 226, 4096, generator=gi, device=dev).to(BF)
        negative = torch.randn(1, 226, 4096, generator=gi, device=dev).to(BF)
        img_lat = torch.zeros(1, 13, 16, 60, 90, device=dev, dtype=BF)
        img_lat[:, :1] = torch.randn(1, 1, 16, 60, 90, generator=gi, device=dev).to(BF)
        with torch.no_grad():
            ipipe(prompt, negative, num_inference_steps=2, guidance_scale=6.0, generator=gi, image_latents=img_lat)    # warm-up
            dec.decode(torch.zeros(1, 16, 13, 60, 90, device=dev, dtype=BF))
            gather_frames(torch.zeros(49, 480, 720, 3, dtype=torch.uint8, device=dev), rank, world)   # first collective of this size: protocol / buffer set-up
            barrier()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
            lat = ipipe(prompt, negative, num_inference_steps=steps, guidance_scale=6.0, generator=gi, image_latents=img_lat)
            ev[1].record()
            frames = dec.decode(lat.permute(0, 2, 1, 3, 4) / 0.7).sample                                              # [1, 3, 49, 480, 720]
            u8 = ((frames[0].float().clamp(-1, 1) + 1.0) * 127.5).round().to(torch.uint8).permute(1, 2, 3, 0).contiguous()   # [49, 480, 720, 3]
            ev[2].record()
            barrier()                                    # ranks finish 50 steps ~1 % apart: keep that skew out of the collective's time
            evb = torch.cuda.Event(enable_timing=True)
            evb.record()
            got = gather_frames(u8, rank, world)
            ev[3].record()
            barrier()
        den_ms, dec_ms, gat_ms = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), evb.elapsed_time(ev[3])
        skew_ms = ev[2].elapsed_time(evb)
        total_ms = ev[0].elapsed_time(ev[3])
        ok = rank != 0 or (len(got) == world and all(tuple(f.shape) == (49, 480, 720, 3) for f in got))
        den_ms, dec_ms, gat_ms, total_ms, skew_ms = _max_over_ranks([den_ms, dec_ms, gat_ms, total_ms, skew_ms], dev)
        nbytes = 49 * 480 * 720 * 3
        out["prompt_shard_clip"] = {
            "workload": f"CogVideoX-5B-I2V (in_channels 32, learned positional embedding; a merged LoRA does not change the FLOPs), 49f 720x480, "
                        f"one prompt per GPU x {world} GPUs, {steps} DDIM steps + tiled VAE decode + gather of the uint8 frames to rank 0",
            "denoise_ms": den_ms, "vae_decode_and_uint8_ms": dec_ms, "frame_gather_ms": gat_ms, "wait_for_slowest_rank_ms": skew_ms, "total_ms": total_ms,
            "clips_per_hour_all_gpus": world * 3600.0 * 1000.0 / total_ms, "gather_bytes_per_rank": nbytes,
            "gather_gbs_per_rank": (world - 1) * nbytes / (gat_ms / 1000.0) / 1e9 if gat_ms > 0 else None, "gather_ok": ok,
            "tokens_per_s_all_gpus": world * 17550 * steps / (den_ms / 1000.0)}
        del imodel, dec, ipipe, lat, frames, u8, got
        torch.cuda.empty_cache()
    except Exception as ex:                              # noqa: BLE001
        out["prompt_shard_clip"] = {"error": f"{type(ex).__name__}: {ex}"}

    # ---- (iii) BASELINE.json configs[4]: DPO step under DDP, one preference pair per GPU, LoRA gradient all-reduce overlapped with backward
    try:
        from videogpa_b200.train_dit import LoRATrainableTransformer
        from videogpa_b200.train_step import DPOSharedStep
        torch.cuda.empty_cache()
        free_gib = torch.cuda.mem_get_info()[0] / 2 ** 30
        ck_mode = "mlp" if free_gib > 135 else True
        pol = LoRATrainableTransformer(model, r=64, lora_alpha=128.0, gradient_checkpointing=ck_mode)
        dstep = DPOSharedStep(model, None, beta=1.0, trainable=pol)
        opt = dstep.configure_optimizers()
        params = list(pol.parameters())
        red = BucketedGradReducer(params, bucket_bytes=32 << 20)
        gt = torch.Generator().manual_seed(rank)
        tb = {"x_win": torch.randn(1, 16, 13, 60, 90, generator=gt), "x_lose": torch.randn(1, 16, 13, 60, 90, generator=gt),
              "prompt_emb": torch.randn(1, 226, 4096, generator=gt).to(BF)}

        def one_step():
            opt.zero_grad(set_to_none=True)
            loss = dstep.training_step(tb)
            red.armed = True
            loss.backward()
            st = red.finish()
            torch.nn.utils.clip_grad_norm_(params, 1.0)
            opt.step()
            return loss, st

        one_step()                                       # warm-up: builds the transposed dgrad weights
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss, st = one_step()
        e1.record()
        barrier()
        ms, exposed = _max_over_ranks([e0.elapsed_time(e1), st["exposed_ms"]], dev)
        pv = [torch.empty_like(params[1]) for _ in range(world)]
        dist.all_gather(pv, params[1].detach().contiguous())
        out["dpo_ddp_step"] = {
            "workload": f"DPO training step, CogVideoX-5B 49f 720x480, one preference pair per GPU x {world} GPUs, LoRA r = 64 (66 M parameters), "
                        "bucketed gradient all-reduce launched from autograd hooks during backward",
            "ms_per_step": ms, "pairs_per_s_all_gpus": world * 1000.0 / ms, "allreduce_bytes": st["bytes"], "buckets": st["buckets"],
            "exposed_comm_ms": exposed, "loss": float(loss.detach()),
            "parameters_identical_across_ranks": bool(all(torch.equal(pv[0], p) for p in pv)),
            "checkpointing": "mlp" if ck_mode == "mlp" else "full"}
        red.remove()
        del pol, dstep, opt, red
        torch.cuda.empty_cache()
    except Exception as ex:                              # noqa: BLE001
        out["dpo_ddp_step"] = {"error": f"{type(ex).__name__}: {ex}"}
    # ---- (i-b) CFG-pair shard, Wan2.2-TI2V-5B (BASELINE.json configs[3]: 4 GPUs = 2 prompts)
    try:
        from videogpa_b200.wan import WanConfig, WanDenoiseStep, WanTransformer3D, flow_sigmas
        grp, how = make_group(peer=False)
        wm = WanTransformer3D.random_init(WanConfig.ti2v_5b(), seed=21, device=dev)
        wstep = WanDenoiseStep(wm, guide_scale=5.0)
        gw = torch.Generator(device=dev).manual_seed(200 + grp.pair)
        wlat = torch.randn(48, 21, 44, 80, device=dev, generator=gw).to(BF)
        wctx = torch.randn(512, 4096, device=dev, generator=gw).to(BF)
        wnull = torch.zeros(1, 4096, device=dev, dtype=BF)
        wS, whw = 21 * 22 * 40, 22 * 40
        sig = flow_sigmas(50, 5.0)
        wt = torch.full((1, wS), sig[0] * 1000.0); wt[:, :whw] = 0
        xw = wstep(wlat, wt, sig[0], sig[1], wctx, wnull, cfg_group=grp, first_frame=wlat[:, :1])
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 2
        for i in range(n):
            xw = wstep(xw, wt, sig[i + 1], sig[i + 2], wctx, wnull, cfg_group=grp, first_frame=wlat[:, :1])
        e1.record()
        barrier()
        pair_ms = e0.elapsed_time(e1) / n
        single_ms = cuda_ms(lambda: wstep(wlat, wt, sig[1], sig[2], wctx, wnull, first_frame=wlat[:, :1]), 1, warmup=1)
        fwd_ms = cuda_ms(lambda: wm([wlat], wt, [wctx]), 2, warmup=1)
        pair_ms, single_ms, fwd_ms = _max_over_ranks([pair_ms, single_ms, fwd_ms], dev)
        out["cfg_pair_wan"] = {
            "workload": f"Wan2.2-TI2V-5B 81f 1280x704 (S = 18 480), {world // 2} prompt(s), cond / uncond on the two GPUs of a pair"
                        + ("" if world == 4 else f" (BASELINE.json configs[3] names 4 GPUs; this run has {world})"),
            "exchange": how, "ms_per_step": pair_ms, "tokens_per_s": (world // 2) * wS / (pair_ms / 1000.0),
            "single_gpu_two_forwards_ms_per_step": single_ms, "speedup_vs_single_gpu": single_ms / pair_ms,
            "one_branch_forward_ms": fwd_ms, "exposed_exchange_and_update_ms": pair_ms - fwd_ms,
            "exchange_bytes_per_step_per_rank": 48 * 21 * 44 * 80 * 2, "finite": bool(torch.isfinite(xw.float()).all().item())}
        del wm, wstep, xw, grp
        torch.cuda.empty_cache()
    except Exception as ex:                              # noqa: BLE001
        out["cfg_pair_wan"] = {"error": f"{type(ex).__name__}: {ex}"}

    # ---- (i-a) CFG-pair shard, CogVideoX-5B: ranks (2p, 2p+1) run uncond / cond of prompt p
    try:
        grp, how = make_group(peer=True)
        ts = pipe.scheduler.set_timesteps(50)
        gl = torch.Generator(device=dev).manual_seed(100 + grp.pair)          # identical latents and prompts inside a pair
        lat = pipe.prepare_latents(1, 49, 480, 720, generator=gl)
        pe = torch.randn(2, 226, 4096, generator=gl, device=dev).to(BF)
        rope = pipe.rotary(13, 60, 90)
        with torch.no_grad():
            x = lat
            for i in range(2):
                x = pipe.denoise_step(x, pe, int(ts[i]), 6.0, rope, cfg_group=grp)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            n = 3
            for i in range(n):
                x = pipe.denoise_step(x, pe, int(ts[2 + i]), 6.0, rope, cfg_group=grp)
            e1.record()
            barrier()
            step_ms = e0.elapsed_time(e1) / n
            # one branch alone (no exchange, no update): what the step costs without the split's communication
            xin = lat
            tt = torch.full((1,), 999.0, device=dev)
            fwd = lambda: model(hidden_states=xin, encoder_hidden_states=pe[:1], timestep=tt, image_rotary_emb=rope, return_dict=False)
            fwd_ms = cuda_ms(fwd, 3, warmup=1)
        both = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(both, x.contiguous())
        same = bool(torch.equal(both[2 * grp.pair], both[2 * grp.pair + 1]))
        step_ms, fwd_ms = _max_over_ranks([step_ms, fwd_ms], dev)
        out["cfg_pair_cogvideox"] = {
            "workload": f"CogVideoX-5B T2V 49f 720x480, {world // 2} prompt(s), cond / uncond on the two GPUs of a pair", "exchange": how,
            "ms_per_step": step_ms, "tokens_per_s": (world // 2) * 17550 / (step_ms / 1000.0),
            "single_gpu_batched_ms_per_step": single_gpu_step_ms, "speedup_vs_single_gpu_batched": single_gpu_step_ms / step_ms,
            "one_branch_forward_ms": fwd_ms, "exposed_exchange_and_update_ms": step_ms - fwd_ms,
            "exchange_bytes_per_step_per_rank": 13 * 16 * 60 * 90 * 2, "latents_identical_inside_pair": same}
        del grp
    except Exception as ex:                              # noqa: BLE001
        out["cfg_pair_cogvideox"] = {"error": f"{type(ex).__name__}: {ex}"}

    return out
