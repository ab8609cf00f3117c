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


def plan_pieces(lengths, world_size, tile=32_000_000, halo=512):
    """Units of the de-novo stage for `world_size` ranks: contigs longer than 1.5 tiles are cut into pieces of
    about `tile` bases (the reference's own answer to long contigs is cropping,
    run_iterative_igdetective.py:41-49), whole contigs otherwise; pieces are assigned by LPT on length.
    A piece is (contig, own_start, own_end, load_start, load_end): the rank loads [load_start, load_end) -- the
    owned range plus `halo` bases of context on either side -- and keeps exactly the rows whose heptamer
    (right heptamer for a D pair) starts inside [own_start, own_end).  halo must cover what a row reads around
    its heptamer: the 350-nt V flank + 7 (py/IGDetective.py:268-286) and a D pair's 157 + the 40-base RSS span
    (:210-224); 512 does, so every owned row is computed from the same letters as on one GPU.
    Returns [pieces of rank r], each list in (contig, own_start) order."""
    if halo < 400:
        raise ValueError("halo must be at least 400 bases (V flank 350 + heptamer 7 + RSS span 40)")
    pieces = []
    for ci, L in enumerate(lengths):
        L = int(L)
        if L <= tile + tile // 2:
            pieces.append((ci, 0, L, 0, L))
            continue
        n = -(-L // tile)
        step = -(-L // n)
        for k in range(n):
            a, b = k * step, min(L, (k + 1) * step)
            pieces.append((ci, a, b, max(0, a - halo), min(L, b + halo)))
    bins = [(0, r) for r in range(world_size)]
    heapq.heapify(bins)
    out = [[] for _ in range(world_size)]
    for p in sorted(pieces, key=lambda p: (-(p[4] - p[3]), p[0], p[1])):
        load, r = heapq.heappop(bins)
        out[r].append(p)
        heapq.heappush(bins, (load + p[4] - p[3], r))
    return [sorted(x) for x in out]


def range_summary(fa):
    """What a rank tells the others about its byte range of the FASTA (fasta.NativeFasta in range form): the letters
    of the range in front of its first header (they continue a record that began in an earlier range) and, for every
    header inside the range, the id and the letters of that record inside the range."""
    lo, hi = fa.byte_range
    own_lo, own_hi = fa.own_letters
    off = [0]
    for n in fa.lengths:
        off.append(off[-1] + n)
    lead_end, recs = own_hi, []
    for i, hb in enumerate(fa.header_bytes):
        if hb is not None and lo <= hb < hi:
            if not recs:
                lead_end = off[i]
            recs.append((fa.ids[i], min(off[i + 1], own_hi) - off[i]))
    return {"lead": max(0, lead_end - own_lo), "recs": recs}


def contig_table(summaries):
    """All ranks' range_summary, in rank order -> (ids, lengths, [(contig open at the start of rank r's range or -1,
    letters of it in front of the range)]).  Letters in front of the file's first header belong to no record (the
    reference's reader drops them).  Record ids must be unique: the dict semantics of a repeated id (last sequence, first
    position) cannot be kept when ranks see different parts of the file."""
    ids, lens, starts, open_ci = [], [], [], -1
    for s in summaries:
        starts.append((open_ci, lens[open_ci] if open_ci >= 0 else 0))
        if open_ci >= 0:
            lens[open_ci] += s["lead"]
        for rid, n in s["recs"]:
            ids.append(rid)
            lens.append(n)
            open_ci = len(ids) - 1
    if len(set(ids)) != len(ids):
        raise ValueError("the FASTA repeats a record id: run it in one process (igdetective.run), which keeps the "
                         "reference's dict semantics for repeated ids")
    return ids, lens, starts


def range_pieces(fa, ids, lens, start, halo=512):
    """Pieces (contig, own_start, own_end, load_start, load_end) of a rank whose loaded contigs are the records of its
    range-form FASTA `fa`, one piece per local record, in contig coordinates.  A record owns the letters it holds inside
    the rank's range; what it holds in the margins is context.  `start` = contig_table(...)[2][rank]."""
    index = {rid: i for i, rid in enumerate(ids)}
    own_lo, own_hi = fa.own_letters
    pieces, off = [], 0
    for i, n in enumerate(fa.lengths):
        s0, s1 = max(off, own_lo), min(off + n, own_hi)
        if i == 0:                                  # the lead: continues the contig that is open at the range's start
            ci, before = start
            p0 = before - (own_lo - off)
            if s1 <= s0 or ci < 0:
                pieces.append((max(ci, 0), 0, 0, 0, n))
                off += n
                continue
        else:
            ci, p0 = index[fa.ids[i]], 0
        if s1 <= s0:
            pieces.append((ci, 0, 0, p0, p0 + n))
        else:
            a, b = p0 + (s0 - off), p0 + (s1 - off)
            p1 = p0 + n
            if (p0 > 0 and a - p0 < halo) or (p1 < lens[ci] and p1 - b < halo):
                raise ValueError("the margin around a byte range holds fewer than %d letters of context: "
                                 "open the ranges with a larger margin" % halo)
            pieces.append((ci, a, b, p0, p1))
        off += n
    return pieces


def own_rows(rows, piece, contig_id):
    """Rows computed on a loaded piece (coordinates relative to load_start, contig id = the piece's) -> the rows
    the piece owns, in contig coordinates.  rows: {gene type: [row]} in the layouts of genes_{V,D,J}.tsv."""
    _, a, b, p0, _ = piece
    out = {}
    for g, rs in rows.items():
        keep = []
        if g == "D":            # [contig, strand, lh idx, ln idx, lh, ln, rh idx, rn idx, rh, rn, gs, ge, gene]
            for r in rs:
                if a <= r[6] + p0 < b:
                    keep.append([contig_id, r[1], r[2] + p0, r[3] + p0, r[4], r[5], r[6] + p0, r[7] + p0, r[8], r[9],
                                 r[10] + p0, r[11] + p0, r[12]])
        else:                   # [contig, strand, hept idx, nona idx, hept, nona, gs, ge, ...]
            for r in rs:
                if a <= r[2] + p0 < b:
                    keep.append([contig_id, r[1], r[2] + p0, r[3] + p0, r[4], r[5], r[6] + p0, r[7] + p0] + list(r[8:]))
        out[g] = keep
    return out


def canonical_row_key(gene, rank_of):
    """Sort key that puts rows of any sharding into the order a single-process canonical run emits them: strand '+'
    before '-', contigs in input order, then the RSS list order for ascending lists -- ascending position of the list's
    first element on '+' (heptamer for V, nonamer for J; D: right heptamer, then left nonamer), descending on '-'."""
    if gene == "D":
        def key(r):
            sgn = 1 if r[1] == "+" else -1
            return (0 if r[1] == "+" else 1, rank_of[r[0]], sgn * r[6], sgn * r[3])
    else:
        col = 2 if gene == "V" else 3

        def key(r):
            return (0 if r[1] == "+" else 1, rank_of[r[0]], r[col] if r[1] == "+" else -r[col])
    return key


def merge_rows(per_rank_rows, contig_order):
    """per_rank_rows: list over ranks of {gene type: rows}; every row starts with (contig id, strand).
    Returns {gene type: rows} ordered as a single-process canonical run emits them (canonical_row_key); the
    result does not depend on how contigs or pieces were distributed."""
    rank_of = {c: i for i, c in enumerate(contig_order)}
    out = {}
    for rows in per_rank_rows:
        for g, rs in rows.items():
            out.setdefault(g, []).extend(rs)
    for g in out:
        out[g].sort(key=canonical_row_key(g, rank_of))
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
