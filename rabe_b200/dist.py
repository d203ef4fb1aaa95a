"""Multi-GPU plumbing for the batch split: batch items are independent, so every
rank owns a contiguous slice (weak scaling: its own batch), there is no data-path collective, and
torch.distributed is used only for the barrier and for reducing timings (max over ranks) and
counts (sum).  Works with the nccl backend on GPUs and with gloo on CPU (tests)."""
import os

import torch
import torch.distributed as dist


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend, device=None):
    rank, world, _ = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world


def shard(total_items, rank, world):
    """Contiguous strong-scaling slice [lo, hi) of `total_items` for `rank` (remainder to the first ranks)."""
    base, rem = divmod(total_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def rank_seed(base_seed, rank):
    """Distinct, reproducible input stream per rank (weak scaling: every rank has its own batch)."""
    return base_seed * 1000003 + rank


def barrier(device=None):
    if dist.is_initialized():
        dist.barrier()
    if device is not None and device.type == "cuda":
        torch.cuda.synchronize(device)


def reduce_max(values, device):
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def reduce_sum(values, device):
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()]


def throughput(items_per_rank, steps, elapsed_ms_per_rank, device):
    """Whole-job items/s: all ranks' items over the slowest rank's time."""
    (t_max,) = reduce_max([elapsed_ms_per_rank], device)
    (n_total,) = reduce_sum([items_per_rank * steps], device)
    return n_total / (t_max / 1e3), t_max


def broadcast_bytes(data, device, src=0):
    """Rank `src` sends a byte string (keys, tables' bases ...) to every rank: one length broadcast + one payload
    broadcast (ncclBroadcast on GPUs; the only data-path-adjacent collective -- it runs once, at setup)."""
    if not dist.is_initialized():
        return bytes(data)
    rank = dist.get_rank()
    n = torch.tensor([len(data) if rank == src else 0], dtype=torch.int64, device=device)
    dist.broadcast(n, src)
    if rank == src:
        buf = torch.frombuffer(bytearray(data), dtype=torch.uint8).to(device)
    else:
        buf = torch.empty(int(n.item()), dtype=torch.uint8, device=device)
    dist.broadcast(buf, src)
    return bytes(buf.cpu().numpy().tobytes())


def gather_fixed(t, device):
    """All ranks contribute a tensor of the SAME shape (fixed-size per-item outputs of equal shards); every rank
    receives the concatenation in rank order (all_gather)."""
    if not dist.is_initialized():
        return t
    parts = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, t.contiguous())
    return torch.cat(parts)


def finalize():
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
