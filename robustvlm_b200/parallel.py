"""Multi-GPU plumbing: one process per GPU (torchrun), batch-sharded.

The inner attack loop is per-sample: every rank attacks its own shard with
its own model replica and NO data-path collective.  The only exchange in the whole FARE /
TeCoA step is the all-reduce of the outer model gradients (replacing the reference's
torch.nn.DataParallel replicate/scatter/gather, train/adversarial_training_clip.py:184-190),
issued once per step over NCCL / NVLink on a flat fp32 bucket.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init(backend: str | None = None) -> tuple[int, int]:
    """(rank, world) from the torchrun environment; initialises the process group if N > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            local = int(os.environ.get("LOCAL_RANK", "0"))
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous shard [lo, hi) of n samples; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi]


def allreduce_mean_(flat_grad: torch.Tensor, local_count: int, global_count: int) -> torch.Tensor:
    """In-place: sum over ranks of (per-rank mean gradient * local_count) / global_count, i.e.
    the gradient of the mean loss over the GLOBAL batch (the loss is a batch mean,
    train/adversarial_training_clip.py:517,528) even when shards are uneven."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        world = dist.get_world_size()
        if local_count * world == global_count and dist.get_backend() == "nccl":
            dist.all_reduce(flat_grad, op=dist.ReduceOp.AVG)      # even shards: one pass, in NCCL
        else:
            flat_grad.mul_(float(local_count))
            dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
            flat_grad.div_(float(global_count))
    return flat_grad


def grad_buckets(layout, n_layers: int):
    """Contiguous [lo, hi) ranges of the flat gradient buffer: one per transformer block (its
    parameters are adjacent in `model.parameters()` order) plus the rest (class / positional
    embedding, proj, conv1, ln_pre in front; ln_post behind).  layout = checkpoint.param_layout."""
    per = [[None, None] for _ in range(n_layers)]
    total = 0
    for name, _, off, n in layout:
        total = max(total, off + n)
        if name.startswith("transformer.resblocks."):
            l = int(name.split(".")[2])
            per[l][0] = off if per[l][0] is None else min(per[l][0], off)
            per[l][1] = off + n if per[l][1] is None else max(per[l][1], off + n)
    lo_all = min(p[0] for p in per)
    hi_all = max(p[1] for p in per)
    for a, b in zip(per, per[1:]):
        assert a[1] == b[0], "block parameters are not contiguous in the flat buffer"
    rest = [(0, lo_all), (hi_all, total)]
    return [tuple(p) for p in per], [r for r in rest if r[1] > r[0]]


class BucketedAllReduce:
    """The outer-gradient all-reduce of the FARE / TeCoA step, bucketed per
    transformer block and overlapped with the backward: the engine records an event when a
    block's parameter gradients are enqueued (`rvlm_vit_train_grads_ev`); a side stream waits on
    it and NCCL averages that block's slice of the flat buffer while the backward of the blocks
    below is still running.  Even shards only (ReduceOp.AVG == the gradient of the global-batch
    mean loss); the caller falls back to `allreduce_mean_` otherwise."""

    def __init__(self, layout, n_layers, device):
        self.blocks, self.rest = grad_buckets(layout, n_layers)
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.side, self.events = None, None
        if self.cuda:
            self.side = torch.cuda.Stream(device=device)
            self.events = [torch.cuda.Event() for _ in range(n_layers + 1)]
            for e in self.events:        # creates the cudaEvent_t handles the C engine records on
                e.record(torch.cuda.current_stream(device))

    def launch(self, flat_grad: torch.Tensor):
        """enqueue every bucket's all-reduce behind its event; returns the work handles"""
        works = []
        op = dist.ReduceOp.AVG if dist.get_backend() == "nccl" else dist.ReduceOp.SUM
        if not self.cuda:                # host tensors (gloo tests): same buckets, no streams
            for lo, hi in list(reversed(self.blocks)) + self.rest:
                works.append(dist.all_reduce(flat_grad[lo:hi], op=op, async_op=True))
            return works
        with torch.cuda.stream(self.side):
            for l in range(len(self.blocks) - 1, -1, -1):
                self.side.wait_event(self.events[l])
                lo, hi = self.blocks[l]
                works.append(dist.all_reduce(flat_grad[lo:hi], op=op, async_op=True))
            self.side.wait_event(self.events[-1])
            for lo, hi in self.rest:
                works.append(dist.all_reduce(flat_grad[lo:hi], op=op, async_op=True))
        return works

    def finish(self, works, flat_grad):
        for w in works:
            w.wait()                     # the current stream waits for NCCL's
        if self.cuda:
            torch.cuda.current_stream(self.device).wait_stream(self.side)
        if dist.get_backend() != "nccl":
            flat_grad.div_(float(dist.get_world_size()))


def gather_batch(t: torch.Tensor) -> torch.Tensor:
    """All-gather shards along dim 0 (uneven shards allowed); used by tests / evaluation."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return t
    world = dist.get_world_size()
    sizes = [torch.zeros(1, dtype=torch.long, device=t.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([t.shape[0]], dtype=torch.long, device=t.device))
    mx = int(max(s.item() for s in sizes))
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    return torch.cat([o[: int(s.item())] for o, s in zip(outs, sizes)], dim=0)
