"""Sharding of the hot path across ranks (one process per GPU) and the deterministic merge.

The path has no exchange step: contigs (scan) and candidate rows (scoring) are independent units.
Contigs are assigned by longest-processing-time bin packing on length; every rank scans and scores
its own contigs; rank 0 gathers the per-rank rows and orders them by (contig index, strand, list
position), which does not depend on the number of ranks."""
import heapq


def partition_contigs(lengths, world_size):
    """LPT: returns [sorted contig indices of rank r]."""
    bins = [(0, r) for r in range(world_size)]
    heapq.heapify(bins)
    out = [[] for _ in range(world_size)]
    for i in sorted(range(len(lengths)), key=lambda i: (-lengths[i], i)):
        load, r = heapq.heappop(bins)
        out[r].append(i)
        heapq.heappush(bins, (load + lengths[i], r))
    return [sorted(x) for x in out]


def partition_by_cost(costs, world_size):
    """Scoring shards ("partition the pair list evenly by sum of cells"): contiguous
    blocks of the item list with near-equal total cost.  Returns [(lo, hi) of rank r]; the blocks keep
    the list order, so concatenating the per-rank results in rank order restores it."""
    total = float(sum(costs))
    cuts, acc, r = [0], 0.0, 1
    for i, c in enumerate(costs):
        acc += c
        while r < world_size and acc >= total * r / world_size:
            cuts.append(i + 1)
            r += 1
    while len(cuts) < world_size:
        cuts.append(len(costs))
    cuts.append(len(costs))
    return [(cuts[k], cuts[k + 1]) for k in range(world_size)]


def all_gather_lists(items, group=None):
    """Concatenation, in rank order, of every rank's list (torch.distributed all_gather_object; gloo or nccl).
    With partition_by_cost blocks this is the single-process list on every rank."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return list(items)
    bucket = [None] * dist.get_world_size(group)
    dist.all_gather_object(bucket, list(items), group=group)
    return [x for part in bucket for x in part]


def merge_rows(per_rank_rows, contig_order):
    """per_rank_rows: list over ranks of {gene type: rows}; every row starts with (contig id, strand).
    Returns {gene type: rows} ordered as a single-process run emits them: strand '+' before '-',
    contigs in input order, list order within a contig (stable)."""
    rank_of = {c: i for i, c in enumerate(contig_order)}
    out = {}
    for rows in per_rank_rows:
        for g, rs in rows.items():
            out.setdefault(g, []).extend(rs)
    for g in out:
        out[g].sort(key=lambda r: (0 if r[1] == "+" else 1, rank_of[r[0]]))
    return out


def gather_rows(rows, contig_order, group=None):
    """torch.distributed gather of row dicts to rank 0 (None elsewhere).  Works on gloo and nccl."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return merge_rows([rows], contig_order)
    world = dist.get_world_size(group)
    bucket = [None] * world if dist.get_rank(group) == 0 else None
    dist.gather_object(rows, bucket, dst=0, group=group)
    if dist.get_rank(group) != 0:
        return None
    return merge_rows(bucket, contig_order)
