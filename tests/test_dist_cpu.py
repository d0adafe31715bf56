"""World-size-2 gloo test of the multi-GPU orchestration (ropebwt3_b200/dist.py) with a CPU stand-in engine built on
the oracle: the collective logic (MAX-reduce of the partial interleave arrays, the fallback flag) is what is tested."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class OracleEngine:
    """Stand-in with the DeviceEngine interface: part p resolves the rows with (row % n_parts == p)."""

    def __init__(self, sym, ln, force_fallback=False):
        self.sym, self.ln, self.force = sym, ln, force_fallback
        self.calls = []

    def new_ka(self, n):
        return torch.empty(n, dtype=torch.int64)

    def rank_part(self, bwt, n, part, n_parts, ka):
        from oracle import oracle as O
        rb, _ = O.mg_rank_plain(self.sym, self.ln, bwt)
        full = (rb >> 6) - np.arange(n)
        mine = np.arange(n) % n_parts == part
        out = np.where(mine, full, -1)
        ka.copy_(torch.from_numpy(out))
        self.calls.append((part, n_parts))
        return 1 if (self.force and n_parts > 1 and part == 1) else 0

    def merge_with_ka(self, bwt, n, ka):
        from oracle import oracle as O
        k = ka.numpy()
        assert (k >= 0).all(), "holes left in the interleave array"
        rb = (k + np.arange(n)) << 6 | bwt.astype(np.int64) << 3
        self.sym, self.ln = O.merge_runs(self.sym, self.ln, rb)


class ModelEngine(OracleEngine):
    """Stand-in that shards like the device does (tests/algo_model.py): part p resolves its slices of walk order from a
    speculative halo and reports 1 when the halo did not give it an exact start (duplicated sequences)."""

    def rank_part(self, bwt, n, part, n_parts, ka):
        from oracle import oracle as O
        import algo_model as M
        out, unres = M.interleave(O.runs2plain(self.sym, self.ln), bwt, 64, part=part, n_parts=n_parts, halo=2)
        ka.copy_(torch.from_numpy(out))
        self.calls.append((part, n_parts))
        return 1 if unres else 0


def _worker(rank, world, port, force, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from ropebwt3_b200 import synth
    from ropebwt3_b200.dist import merge_plain_sharded
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    if force == "model":         # real slice sharding; the third genome is an exact copy of the second: no halo can resolve it
        gs = synth.genomes(2, 1500, seed=21)
        gs.append(gs[1].copy())
    else:
        gs = synth.genomes(3, 1500, seed=21)
    b0 = O.build_bwt(synth.batch_text(gs[:1]))
    sym, ln = O.plain2runs(b0)
    eng = ModelEngine(sym, ln) if force == "model" else OracleEngine(sym, ln, force)
    used = []
    for g in gs[1:]:
        bwt = O.build_bwt(synth.batch_text([g]))
        used.append(merge_plain_sharded(eng, bwt, len(bwt)))
    # expected: the plain sequential merge
    s0, l0 = sym, ln
    for g in gs[1:]:
        s0, l0 = O.merge_plain(s0, l0, O.build_bwt(synth.batch_text([g])))
    ok = np.array_equal(eng.sym, s0) and np.array_equal(eng.ln, l0)
    q.put((rank, ok, used, eng.calls))
    dist.destroy_process_group()


def _run(force):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, force, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(res)


def test_sharded_merge_world2_gloo():
    res = _run(False)
    for rank, ok, used, calls in res:
        assert ok and used == [True, True]
        assert calls == [(rank, 2), (rank, 2)]


def test_fallback_when_a_rank_is_incomplete():
    res = _run(True)
    for rank, ok, used, calls in res:
        assert ok and used == [False, False]            # every rank recomputed the whole array
        assert calls == [(rank, 2), (0, 1), (rank, 2), (0, 1)]


def test_slice_sharding_model_and_natural_fallback():
    """The device's sharding rule (slices + halo) behind the same collectives: a diverged genome is resolved by the two
    parts together, an exact duplicate makes part 1 report 'incomplete' and every rank falls back -- same result."""
    res = _run("model")
    for rank, ok, used, calls in res:
        assert ok and used == [True, False]
        assert calls == [(rank, 2), (rank, 2), (0, 1)]
