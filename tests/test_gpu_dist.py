"""Multi-device parity (needs >= 2 GPUs; skipped on a single-GPU box): the sharded merge inside the C ABI
(rb3b_merge_plain_dist[_dev]: sharded rank phase + NCCL exchange + replicated merge) must leave on EVERY rank exactly the
index that the single-device merge builds, for world sizes 2..4, with one process per GPU (torch.distributed carries the
NCCL id) and with one host thread per GPU inside one process (explicit contexts)."""
import os
import socket
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _genomes():
    from ropebwt3_b200 import synth
    gs = synth.genomes(7, 60000, seed=11, sub=0.01, indel=0.001)
    gs.append(gs[2].copy())  # an exact duplicate: a match longer than any halo
    return gs


def _expected(rb3, gs, per):
    from ropebwt3_b200 import synth
    idx = None
    for b in range(0, len(gs), per):
        bwt = rb3.rb3_build_sais(synth.batch_text(gs[b:b + per]))
        if idx is None:
            idx = rb3.Index.from_plain(bwt)
        else:
            idx.merge_plain(bwt)
    return idx.export_runs()


def _proc(rank, world, port, q, pairs=0, dist_async=0):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ropebwt3_b200 as R
    from ropebwt3_b200 import synth, dist as rdist
    R.init(rank)
    R.set_param("dist_pairs", pairs)   # 1: (row, position) pairs routed by all-to-all + all-gather instead of the all-reduce
    R.set_param("dist_async", dist_async)   # 1: exchange + merge queued on the second stream / second communicator (default 0)
    rdist.init_library_comm()
    gs = _genomes()
    idx, sharded = None, []
    for b in range(0, len(gs), 2):
        bwt = R.rb3_build_sais(synth.batch_text(gs[b:b + 2]))
        if idx is None:
            idx = R.Index.from_plain(bwt)
        elif b % 4 == 0:   # host-buffer entry point
            sharded.append(rdist.merge_plain_dist(idx, bwt.ctypes.data, len(bwt)))
        else:
            d = torch.from_numpy(bwt).cuda(rank)
            torch.cuda.synchronize()
            sharded.append(rdist.merge_plain_dist_dev(idx, d.data_ptr(), len(bwt)))
    s, l = idx.export_runs()
    q.put((rank, s, l, sharded))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,pairs,dist_async", [(2, 0, 0), (2, 0, 1), (2, 1, 0), (4, 0, 0), (4, 1, 0), (4, 0, 1)])
def test_dist_merge_processes(rb3, world, pairs, dist_async):
    if _n_gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    s0, l0 = _expected(rb3, _genomes(), 2)
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_proc, args=(r, world, port, q, pairs, dist_async)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in ps:
        p.join(60)
    for rank, s, l, sharded in res:
        assert np.array_equal(s, s0) and np.array_equal(l, l0), "rank %d built a different index" % rank
        assert sharded[0], sharded   # (the batch with the duplicated genome may or may not need the exact fallback)


def test_dist_merge_threads_one_process(rb3):
    """N host threads, each with an explicit context on its own device, join one communicator (the CLI's RB3B_DEVICES mode)."""
    world = min(_n_gpus(), 4)
    if world < 2:
        pytest.skip("needs 2 GPUs")
    import ctypes as C
    from ropebwt3_b200 import capi, synth
    L = capi.lib()
    gs = _genomes()[:6]
    s0, l0 = _expected(rb3, gs, 2)
    bwts = [rb3.rb3_build_sais(synth.batch_text(gs[b:b + 2])) for b in range(0, len(gs), 2)]
    uid = C.create_string_buffer(128)
    capi.check(L.rb3b_dist_unique_id(uid))
    out, err = [None] * world, []

    def work(r):
        try:
            ctx = L.rb3b_ctx_create(r)
            assert ctx
            capi.check(L.rb3b_ctx_make_current(ctx))
            capi.check(L.rb3b_dist_init(r, world, uid))
            h = L.rb3b_index_create()
            capi.check(L.rb3b_index_from_plain(h, len(bwts[0]), bwts[0].ctypes.data))
            for b in bwts[1:]:
                capi.check(L.rb3b_merge_plain_dist(h, len(b), b.ctypes.data))
            n = capi.check(L.rb3b_export_runs(h, None, None, 0))
            s, l = np.empty(n, np.uint8), np.empty(n, np.int64)
            capi.check(L.rb3b_export_runs(h, s.ctypes.data, l.ctypes.data, n))
            out[r] = (s, l)
            L.rb3b_index_destroy(h)
            capi.check(L.rb3b_dist_finalize())
            capi.check(L.rb3b_ctx_make_current(None))
            L.rb3b_ctx_destroy(ctx)
        except Exception as e:  # noqa: BLE001
            err.append((r, repr(e)))

    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(600)
    assert not err, err
    for r in range(world):
        assert np.array_equal(out[r][0], s0) and np.array_equal(out[r][1], l0), "thread %d built a different index" % r
