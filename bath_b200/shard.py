"""Multi-GPU plumbing of the path: the target shards by sequence block across ranks with the profiles replicated, and no
collective runs on the DP path.  What crosses ranks is bookkeeping only: timings (max over ranks), residue
counts (sum -- the E-value search space, src/bathsearch.c:869-883) and hit lists (gathered to rank 0, merged as the
reference merges its worker threads, :886-921).  Works over any torch.distributed backend (NCCL on GPUs, gloo in tests)."""
import numpy as np
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_blocks(n_blocks, rank, world_size):
    """Blocks are dealt round-robin, as the reference's reader deals them to worker threads (src/bathsearch.c:1119-1222)."""
    return list(range(rank, n_blocks, world_size))


def reduce_scalar(x, op="max", device="cpu"):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op={"max": dist.ReduceOp.MAX, "sum": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN}[op])
    return float(t.item())


def gather_hits(hits):
    """Every rank's hit records on rank 0 (None elsewhere)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(hits)
    out = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(list(hits), out, dst=0)
    if dist.get_rank() != 0:
        return None
    return [h for part in out for h in part]


def merge_hits(hits, total_residues, max_length, E=10.0):
    """Rank 0: E-values over the whole search space, then the reference's ordering (src/p7_tophits.c:789-800, :262-284).
    Each hit carries the lnP its rank computed BEFORE the search-space correction ('lnP_raw')."""
    w = max_length * 3
    out = []
    for h in hits:
        h = dict(h)
        h["lnP"] = h["lnP_raw"] + float(np.log(np.float32(total_residues) / np.float32(w)))
        h["evalue"] = float(np.exp(h["lnP"]))
        if h["evalue"] <= E:
            out.append(h)
    out.sort(key=lambda h: (h["lnP"], h["name"], -h["strand"], h["ali_from"]))
    return out
