"""One-process-per-GPU sharding of the denoise-and-score path.

The reference shards independent work items over processes with no collective
(`items[i::num_gpus]`, replicate.py:119-133; contiguous chunks for scoring, replicate_scorer.py:244-250).
Here the processes are torch.distributed ranks (NCCL over NVLink on the GPU box, gloo in the CPU
tests) and the only exchanges are
  * CfgPairGroup.exchange — 2-rank all-gather of the noise prediction, once per denoise step, when the
    cond/uncond pair of one prompt is split over two GPUs (4.5 MB fp32-equivalent for CogVideoX);
  * gather_frames / gather_scores — the final gather of decoded uint8 frames or of scalar scores to rank 0.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun). Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_round_robin(items, rank: int, world: int):
    """Prompt shard: rank r takes items r, r+world, ... (replicate.py:120)."""
    return list(items)[rank::world]


def shard_padded(items, rank: int, world: int):
    """Training shard with the size rule of torch.utils.data.DistributedSampler (drop_last=False), which Lightning's DDP
    strategy wraps around the reference's train loader (train/CogVideoX-5B/03_train.py:251,258-266): the index list is padded
    to a multiple of `world` by wrapping around to its start, then rank r takes r, r+world, ... Every rank gets exactly
    ceil(n / world) items, so all ranks run the same number of micro-batches and issue the same number of all-reduces."""
    items = list(items)
    if not items:
        return []
    total = -(-len(items) // world) * world
    pad = total - len(items)
    items = items + (items * (pad // len(items) + 1))[:pad]
    return items[rank:total:world]


def shard_contiguous(items, rank: int, world: int):
    """Scorer shard: contiguous chunks, remainder spread over the first ranks (replicate_scorer.py:244-250)."""
    items = list(items)
    base, rem = divmod(len(items), world)
    start = rank * base + min(rank, rem)
    return items[start:start + base + (1 if rank < rem else 0)]


class CfgPairGroup:
    """Ranks (2p, 2p+1) hold the uncond / cond branch of prompt-group p."""

    def __init__(self, rank: int, world: int, share: "CfgPairGroup | None" = None):
        """share: reuse the 2-rank process group of an existing CfgPairGroup instead of creating new ones."""
        if world % 2 != 0:
            raise RuntimeError("CFG-pair sharding needs an even number of ranks")
        self.rank, self.world = rank, world
        self.pair = rank // 2
        self.branch = rank % 2            # 0 = uncond (negative prompt), 1 = cond
        self.group = None
        if share is not None:
            self.group = share.group
            return
        for p in range(world // 2):       # every rank must take part in every new_group call
            g = dist.new_group(ranks=[2 * p, 2 * p + 1])
            if p == self.pair:
                self.group = g

    def exchange(self, pred: torch.Tensor):
        """-> (pred_uncond, pred_cond), identical on both ranks of the pair."""
        pred = pred.contiguous()
        both = torch.empty((2,) + tuple(pred.shape), dtype=pred.dtype, device=pred.device)
        dist.all_gather_into_tensor(both.view(-1), pred.view(-1), group=self.group)
        return both[0], both[1]


class CfgPairPeerGroup(CfgPairGroup):
    """CFG-pair sharding with the exchange FUSED into the guidance + scheduler kernel over NVLink peer memory.

    Instead of an NCCL all-gather followed by the update kernel, every rank publishes its noise prediction in a symmetric
    (peer-mapped) buffer, and `vgpa_cfg_scheduler_step` reads the partner's half straight through the peer pointer while it
    combines guidance and applies the scheduler update: one kernel does the transfer and the math. Two device-side barriers
    on the symmetric-memory signal pads order "partner has written" before and "partner has read" after the kernel; no
    collective library call is on the data path. Falls back to nothing: construction raises if peer access is unavailable,
    and for more than one pair (world size > 2).
    """

    def __init__(self, rank: int, world: int, share: "CfgPairGroup | None" = None):
        if world != 2:
            # measured on 4 B200s (two pairs): the first fused step never returned (profiles/r02_multi_gpu.md); until the
            # symmetric-memory rendezvous on 2-rank subgroups is understood, several pairs exchange through CfgPairGroup (NCCL)
            raise RuntimeError("CfgPairPeerGroup is validated for one CFG pair (world size 2); use CfgPairGroup with more ranks")
        super().__init__(rank, world, share)
        self._buf = None
        self._hdl = None
        self._n = 0

    def _ensure(self, pred: torch.Tensor):
        if self._buf is not None and self._n == pred.numel():
            return
        import torch.distributed._symmetric_memory as symm_mem
        self._buf = symm_mem.empty(pred.numel(), dtype=pred.dtype, device=pred.device)
        self._hdl = symm_mem.rendezvous(self._buf, self.group)
        self._n = pred.numel()
        if self._hdl.world_size != 2:
            raise RuntimeError("CFG-pair peer exchange needs a 2-rank group")

    def peer_views(self, pred: torch.Tensor):
        """Publish `pred` and return (pred_uncond, pred_cond) where the partner's half is a tensor aliasing ITS memory."""
        self._ensure(pred)
        self._buf.copy_(pred.reshape(-1))
        self._hdl.barrier(channel=0)                                   # both halves are written
        me = self._hdl.rank
        mine = self._buf.view(pred.shape)
        theirs = self._hdl.get_buffer(1 - me, tuple(pred.shape), pred.dtype)
        return (mine, theirs) if self.branch == 0 else (theirs, mine)

    def release(self):
        self._hdl.barrier(channel=1)                                   # the partner has finished reading my half


def gather_frames(frames: torch.Tensor, rank: int, world: int, dst: int = 0):
    """Final decoded-frame gather: every rank contributes a [n_i, ...] uint8 tensor with identical trailing
    shape; rank `dst` receives the list ordered by rank, the others get None.

    One all-gather collective into a [world, n_max, ...] buffer (NCCL: ring / NVLS over NVSwitch; 50.8 MB per 49-frame clip):
    the send/recv pairs of a rooted gather set up point-to-point channels on first use and measured 0.2-0.9 s on 2-4 B200s,
    the all-gather runs on the communicator's existing rings."""
    if world == 1:
        return [frames]
    n = torch.tensor([frames.shape[0]], dtype=torch.int64, device=frames.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    nmax = max(counts)
    if frames.shape[0] == nmax:
        pad = frames.contiguous()
    else:
        pad = torch.zeros((nmax,) + tuple(frames.shape[1:]), dtype=frames.dtype, device=frames.device)
        pad[:frames.shape[0]] = frames
    if dist.get_backend() == "nccl":
        buf = torch.empty((world,) + tuple(pad.shape), dtype=pad.dtype, device=pad.device)
        dist.all_gather_into_tensor(buf, pad)
        bufs = list(buf.unbind(0))
    else:
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad)
    if rank != dst:
        return None
    return [b[:c] for b, c in zip(bufs, counts)]


def gather_scores(scores: torch.Tensor, rank: int, world: int):
    """All-gather of per-rank score vectors (variable length) -> concatenated tensor in rank order on every rank."""
    if world == 1:
        return scores
    n = torch.tensor([scores.numel()], dtype=torch.int64, device=scores.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    nmax = int(max(int(c.item()) for c in counts))
    pad = torch.zeros(nmax, dtype=scores.dtype, device=scores.device)
    pad[:scores.numel()] = scores.reshape(-1)
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[:int(c.item())] for b, c in zip(bufs, counts)])


def average_gradients(params, group=None, bucket_bytes: int = 256 << 20) -> int:
    """The data-path collective of the DPO training step (config 5: DDP over the GPUs of one box; Lightning's
    `strategy="ddp"` in train/CogVideoX-5B/03_train.py:258-266): all-reduce-average the `.grad` of `params` in place.

    Only the LoRA factors train (66 M fp32 values = 264 MB for CogVideoX-5B at r = 64), so the gradients are packed into
    flat fp32 buckets of at most `bucket_bytes` and reduced with one NCCL all-reduce per bucket (NVSwitch reduces in the
    switch when NVLS is available; the bucket is sized for launch latency, not link count). Parameters without a gradient
    contribute zeros so that every rank issues the same collectives. Returns the number of collectives issued."""
    import torch.distributed as dist
    params = [p for p in params]
    world = dist.get_world_size(group)
    if world == 1 or not params:
        return 0
    calls = 0
    i = 0
    while i < len(params):
        n, j = 0, i
        while j < len(params) and (j == i or (n + params[j].numel()) * 4 <= bucket_bytes):
            n += params[j].numel()
            j += 1
        ref = params[i]
        flat = torch.zeros(n, dtype=torch.float32, device=ref.device)
        off = 0
        for p in params[i:j]:
            if p.grad is not None:
                flat[off:off + p.numel()].copy_(p.grad.reshape(-1))
            off += p.numel()
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
        calls += 1
        off = 0
        for p in params[i:j]:
            g = flat[off:off + p.numel()].view_as(p).to(p.dtype)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += p.numel()
        i = j
    return calls


class BucketedGradReducer:
    """DDP gradient average overlapped with the backward pass (Lightning's `strategy="ddp"` of
    train/CogVideoX-5B/03_train.py:258-266 = torch DistributedDataParallel: bucketed all-reduce launched from autograd hooks).

    `params` are grouped into buckets of at most `bucket_bytes` in REVERSE order (the order in which backward produces
    gradients: last transformer block first). A post-accumulate-grad hook on every parameter marks it ready; when the last
    parameter of a bucket is ready — and `armed` is set, i.e. this backward is the last micro-batch of the accumulation
    window — the bucket's gradients are packed into one flat fp32 buffer and `dist.all_reduce(async_op=True)` is issued on
    the process group's own stream, while backward continues on the compute stream. `finish()` waits for the handles,
    divides by the world size and copies the averages back into `.grad`. Only the LoRA factors train (66 M values = 264 MB
    for CogVideoX-5B at r = 64), so a handful of buckets covers the step; the bucket size is chosen for launch latency and
    overlap, not link count (NVSwitch gives every peer full bandwidth).

    finish() returns {"buckets", "bytes", "exposed_ms"}: exposed_ms is the time the compute stream spent waiting for the
    collectives and unpacking after backward had finished (CUDA events; 0.0 on CPU / gloo)."""

    def __init__(self, params, group=None, bucket_bytes: int = 32 << 20):
        self.params = [p for p in params]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.armed = True
        self.buckets: list[list[int]] = []
        cur, n = [], 0
        for i in reversed(range(len(self.params))):
            sz = self.params[i].numel() * 4
            if cur and n + sz > bucket_bytes:
                self.buckets.append(cur)
                cur, n = [], 0
            cur.append(i)
            n += sz
        if cur:
            self.buckets.append(cur)
        self._bucket_of = {i: b for b, idxs in enumerate(self.buckets) for i in idxs}
        self._pending = [len(b) for b in self.buckets]
        self._inflight: dict[int, tuple] = {}
        self._hooks = []
        if self.world > 1:
            for i, p in enumerate(self.params):
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(i)))

    def _make_hook(self, i: int):
        def hook(_p):
            if not self.armed:
                return
            b = self._bucket_of[i]
            self._pending[b] -= 1
            if self._pending[b] == 0:
                self._launch(b)
        return hook

    def _launch(self, b: int):
        ps = [self.params[i] for i in self.buckets[b]]
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in ps])
        self._inflight[b] = (flat, dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self) -> dict:
        """Call after backward of the armed micro-batch, before clipping / optimizer.step()."""
        stats = {"buckets": 0, "bytes": 0, "exposed_ms": 0.0}
        if self.world == 1:
            return stats
        for b in range(len(self.buckets)):                  # parameters that received no gradient this step never fired
            if b not in self._inflight:
                self._launch(b)
        cuda = self.params[0].is_cuda
        if cuda:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        for b, (flat, handle) in sorted(self._inflight.items()):
            handle.wait()                                       # NCCL: the compute stream waits on the collective's stream
            flat.div_(self.world)
            off = 0
            for i in self.buckets[b]:
                p = self.params[i]
                g = flat[off:off + p.numel()].view_as(p).to(p.dtype)
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)
                off += p.numel()
            stats["buckets"] += 1
            stats["bytes"] += flat.numel() * 4
        if cuda:
            e1.record()
            e1.synchronize()
            stats["exposed_ms"] = float(e0.elapsed_time(e1))
        self._inflight.clear()
        self._pending = [len(b) for b in self.buckets]
        return stats

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
