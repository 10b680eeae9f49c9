"""What crosses process ranks when bench.py runs one process per GPU (torchrun): nothing on the DP path -- a barrier and two
reductions of scalars (timings: max over ranks; cells, residues: sum).  The sharded SEARCH does not use ranks at all: one
process drives every device through bathhost_search_create_multi (include/bathhost.h), so under torchrun rank 0 runs it over all
N devices while the other ranks wait.  Works over any torch.distributed backend (NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def reduce_scalar(x, op="max", device="cpu"):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op={"max": dist.ReduceOp.MAX, "sum": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN}[op])
    return float(t.item())


def search_devices(rank, world_size, n_visible):
    """Devices the sharded search of this rank drives: rank 0 takes min(world_size, visible devices), the others none."""
    return list(range(min(world_size, n_visible))) if rank == 0 else []
