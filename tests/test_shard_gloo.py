"""N > 1 path on CPU: world_size-2 gloo processes partition contigs, produce their rows and gather
them on rank 0; the merged table must equal the single-process one."""
import os
import socket

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from igdetective_b200 import shard


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


LENGTHS = [500, 90_000, 1_200, 33_000, 64, 7_777, 250_000, 10]
IDS = ["ctg%d" % i for i in range(len(LENGTHS))]


def _rows_for(contigs):
    """stand-in for the per-rank scan + scoring: rows derived only from the contig itself"""
    rows = {"V": [], "D": [], "J": []}
    for c in contigs:
        L = LENGTHS[IDS.index(c)]
        for k in range(L % 5):
            for sd in "+-":
                rows["V"].append([c, sd, k * 11, k * 11 + 30, "CACAGTG", "ACAAAAACC", L + k])
        rows["D"].append([c, "+", L % 97, 0])
        rows["J"].append([c, "-", L % 13, 1])
    return rows


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = [IDS[i] for i in shard.partition_contigs(LENGTHS, world)[rank]]
    merged = shard.gather_rows(_rows_for(mine), IDS)
    if rank == 0:
        q.put(merged)
    dist.barrier()
    dist.destroy_process_group()


def test_partition_is_balanced_and_complete():
    for world in (1, 2, 4, 8):
        parts = shard.partition_contigs(LENGTHS, world)
        assert sorted(i for p in parts for i in p) == list(range(len(LENGTHS)))
        loads = [sum(LENGTHS[i] for i in p) for p in parts]
        assert max(loads) <= max(max(LENGTHS), sum(LENGTHS) / world * 1.5)


def test_two_rank_gather_equals_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = shard.merge_rows([_rows_for(IDS)], IDS)
    assert merged == single
