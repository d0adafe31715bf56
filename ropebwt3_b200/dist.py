"""Multi-GPU merge (first step): one process per GPU, the index replicated on every device, the rows of the
batch split among the ranks for the rank phase, one NCCL all-reduce (MAX) of the 8-byte-per-row interleave array, then
the same streaming merge on every rank.  torch.distributed is plumbing only; the compute is in librb3b200.so.

The engine is passed in so that the collective logic can be tested on CPU with gloo and a stand-in engine.
"""
import torch
import torch.distributed as dist


class DeviceEngine:
    """The C-ABI calls on device pointers (rb3b_mg_rank_part / rb3b_merge_with_ka)."""

    def __init__(self, index):
        from . import capi
        self.idx, self.capi, self.L = index, capi, capi.lib()

    def new_ka(self, n):
        return torch.empty(n, dtype=torch.int64, device="cuda")

    def rank_part(self, bwt, n, part, n_parts, ka):
        rc = self.L.rb3b_mg_rank_part(self.idx.h, n, int(bwt), part, n_parts, ka.data_ptr())
        return self.capi.check(rc)

    def merge_with_ka(self, bwt, n, ka):
        self.capi.check(self.L.rb3b_merge_with_ka(self.idx.h, n, int(bwt), ka.data_ptr()))


def merge_plain_sharded(engine, bwt, n, group=None):
    """rb3_fmi_merge_plain across the ranks of `group`.  Returns True when the sharded rank phase was used, False when
    a rank could not resolve its rows locally and every rank recomputed the whole array (still correct)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    ka = engine.new_ka(n)
    rc = engine.rank_part(bwt, n, rank, world, ka)
    sharded = True
    if world > 1:
        flag = torch.tensor([rc], dtype=torch.int64, device=ka.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
        if int(flag.item()) > 0:
            engine.rank_part(bwt, n, 0, 1, ka)   # rare: a match longer than the halo at a stretch boundary
            sharded = False
        else:
            dist.all_reduce(ka, op=dist.ReduceOp.MAX, group=group)  # rows nobody else resolved are -1
    engine.merge_with_ka(bwt, n, ka)
    return sharded
