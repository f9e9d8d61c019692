"""CPU, world_size 2 over gloo: GOP sharding + record gathering gives the same table and bit-identical
fp64 totals as a single process."""
import os
import socket
import sys

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _records_for(units, frames_per_gop=7):
    rows = []
    for u in units:
        g = torch.Generator().manual_seed(1000 + u)
        for f in range(1, frames_per_gop + 1):
            bits = 1e5 * (1.0 + torch.rand(1, generator=g, dtype=torch.float64).item())
            sse = float(int(1e6 * torch.rand(1, generator=g, dtype=torch.float64).item()))
            rows.append([float(u), float(f), bits, sse])
    return torch.tensor(rows, dtype=torch.float64).reshape(-1, 4)


def _worker(rank, world, port, units, out_dir):
    sys.path.insert(0, os.path.join(ROOT, "video-compression_b200"))
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from b200vc import dist as bd
    from b200vc import gop
    bd.init(backend="gloo")
    mine = gop.shard_units(units, world, rank)
    table = bd.gather_records(_records_for(mine))
    tot = bd.totals(table)
    slowest = bd.max_over_ranks(10.0 + rank, torch.device("cpu"))
    torch.save({"table": table, "totals": tot, "mine": list(mine), "slowest": slowest},
               os.path.join(out_dir, f"r{rank}.pt"))
    bd.barrier()
    torch.distributed.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_gather_equals_single_process(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "video-compression_b200"))
    from b200vc import dist as bd
    units = 11  # uneven split: 5 + 6
    mp.spawn(_worker, args=(2, _free_port(), units, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "r0.pt")
    r1 = torch.load(tmp_path / "r1.pt")
    assert r0["mine"] == list(range(0, 5)) and r1["mine"] == list(range(5, 11))
    single = bd.gather_records(_records_for(range(units)))
    assert torch.equal(r0["table"], single) and torch.equal(r1["table"], single)
    want = bd.totals(single)
    assert torch.equal(r0["totals"], want) and torch.equal(r1["totals"], want)  # bit-identical fp64
    assert r0["slowest"] == 11.0 and r1["slowest"] == 11.0


def test_gather_handles_an_empty_rank(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), 1, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "r0.pt")
    r1 = torch.load(tmp_path / "r1.pt")
    assert r0["mine"] == [] and r1["mine"] == [0]
    assert r0["table"].shape == (7, 4) and torch.equal(r0["table"], r1["table"])
