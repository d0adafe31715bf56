"""Multi-GPU merge (first step): one process per GPU, the index replicated on every device, the rows of the
batch split among the ranks for the rank phase, the interleave positions combined over NVLink, then the same streaming
merge on every rank.

The product path is inside the C ABI: `init_library_comm()` makes this process's librb3b200 context a rank of an NCCL
communicator owned by the library (torch.distributed only carries the 128-byte NCCL id to the other ranks), after which
`rb3b_merge_plain_dist[_dev]` does the sharded rank phase, the NCCL exchange and the merge in one call.

`merge_plain_sharded()` below is the same orchestration written against an engine object, so that the collective logic
(partial arrays, MAX-combine, the fallback flag) can be tested on CPU with gloo and a stand-in engine.
"""
import torch
import torch.distributed as dist


def init_library_comm(group=None):
    """rb3b_dist_init for this process: rank 0 draws the NCCL unique id, every rank joins.  Call after rb3b_init(device)."""
    import ctypes as C
    from . import capi
    L = capi.lib()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    buf = C.create_string_buffer(128)
    if rank == 0:
        capi.check(L.rb3b_dist_unique_id(buf))
    box = [buf.raw]
    if world > 1:
        dist.broadcast_object_list(box, src=0, group=group)
    capi.check(L.rb3b_dist_init(rank, world, C.create_string_buffer(box[0], 128)))
    return rank, world


def merge_plain_dist_dev(index, d_bwt, n):
    """rb3_fmi_merge_plain across the library's communicator, batch in device memory -> True when sharded (no fallback)."""
    from . import capi
    return capi.check(capi.lib().rb3b_merge_plain_dist_dev(index.h, n, int(d_bwt))) == 0


def merge_plain_dist(index, h_bwt_ptr, n):
    """same, batch in (pinned) host memory: every rank copies 1/world of it, NVLink all-gathers the rest"""
    from . import capi
    return capi.check(capi.lib().rb3b_merge_plain_dist(index.h, n, int(h_bwt_ptr))) == 0


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
