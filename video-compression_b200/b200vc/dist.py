"""One process per GPU; ``torch.distributed`` (NCCL over NVLink on the B200 box, gloo in CPU tests) is used for
exactly one thing: gathering the per-frame (gop, frame, bits, sse) records at the end of a GOP-sharded run.
The reference has no distributed code at all; GOP independence is the only parallelism.

Totals are formed by summing the gathered per-frame records in global frame order on every rank, so a 1-GPU
and an N-GPU run produce bit-identical fp64 totals (an all-reduce would re-associate the sum).
"""
import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init(backend=None):
    """Initialise the default process group from the torchrun environment (no-op for world size 1)."""
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def world_size():
    return dist.get_world_size() if dist.is_initialized() else 1


def barrier():
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(value, device):
    """Max of a python float over ranks (device timing: the slowest rank defines the step)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def gather_records(records):
    """records: [n_local, K] float64 tensor (rows = per-frame records, column 0/1 = global gop / frame index).
    Returns the concatenation over ranks sorted by (gop, frame) -- identical on every rank."""
    if records.dtype != torch.float64 or records.dim() != 2:
        raise ValueError("records must be a 2-D float64 tensor")
    if dist.is_initialized():
        W = dist.get_world_size()
        n = torch.tensor([records.shape[0]], dtype=torch.int64, device=records.device)
        counts = [torch.zeros_like(n) for _ in range(W)]
        dist.all_gather(counts, n)
        counts = [int(c.item()) for c in counts]
        cap = max(counts) if counts else 0
        padded = torch.zeros((cap, records.shape[1]), dtype=torch.float64, device=records.device)
        padded[: records.shape[0]] = records
        bufs = [torch.zeros_like(padded) for _ in range(W)]
        dist.all_gather(bufs, padded)
        records = torch.cat([b[:c] for b, c in zip(bufs, counts)], 0)
    if records.shape[0] == 0:
        return records
    key = records[:, 0] * 1e6 + records[:, 1]
    return records[torch.argsort(key)]


def totals(records):
    """Sequential fp64 sums in global frame order: columns 2.. of the sorted record table."""
    rec = records.cpu()
    out = torch.zeros(rec.shape[1] - 2, dtype=torch.float64)
    for row in rec:
        out += row[2:]
    return out
