#!/usr/bin/env python
"""bench.py -- bases/sec indexed through the BWT-merge hot path (BASELINE.json configs[1]).

Workload: merge-build of synthetic 5 Mb bacterial genomes, one genome (both strands,
10^7 symbols) per merge, exactly the README multi-file form of `ropebwt3 build`.
A *step* is one rb3_fmi_merge_plain of one genome's partial BWT into the growing index.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our CUDA path
  python bench.py --impl reference [...]                        the reference's CPU path (oracle/_ref)

`value`  = bases/s with the partial BWTs already resident in HBM (device pointers in).
`e2e`    = the same merges through the host-buffer C-ABI call (H2D copy of every batch and
           the D2H reads of the result inside the timed region).
`roofline` is for the dominant kernel (k_walk_first), timed with CUDA events on its stream.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bases/sec indexed (build)"
UNIT = "bases/s"
GENOME_LEN = 5_000_000
SEED = 43  # 42 + config index
LF_STEP_BYTES = 160  # SURVEY 8d, whole LF-walk step: 128-B index cell + 8 B query + 8 B result + 8 B LF_B read + 8 B ka write
WALK_ROW_BYTES = 137  # what k_walk_first itself must move per row: one 128-B cell line + 1 B symbol in + 8 B position out


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=96)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--genome-len", type=int, default=GENOME_LEN)
    ap.add_argument("--seg-len", type=int, default=0)
    ap.add_argument("--genomes-per-merge", type=int, default=0,
                    help="genomes per batch (one batch = one partial BWT = one merge); default: the number of GPUs, i.e. weak scaling")
    ap.add_argument("--ref-budget-s", type=float, default=240.0, help="--impl reference: CPU seconds all steps together may take (the number of timed steps shrinks to fit, never the genomes)")
    ap.add_argument("--config", default="c1", choices=["c1", "c2s"],
                    help="c1: BASELINE configs[1] (default).  c2s: configs[2] scaled to one box's memory budget: --c2s-genomes x 5 Mb in batches of 100 genomes (10^9 symbols per merge)")
    ap.add_argument("--c2s-genomes", type=int, default=1000)
    ap.add_argument("--param", action="append", default=[], help="engine tuning knob key=value (rb3b_set_param), repeatable")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-prefetch", action="store_true", help="e2e leg: do not queue the H2D copy of batch i+1 (rb3b_prefetch_batch) before merging batch i")
    ap.add_argument("--no-rank-bench", action="store_true")
    ap.add_argument("--no-build", action="store_true", help="skip the text-in build_value leg")
    return ap.parse_args()


def apply_config(a):
    if a.config == "c2s":
        if a.genomes_per_merge <= 0:
            a.genomes_per_merge = 100
        n_batches = max(3, a.c2s_genomes // a.genomes_per_merge)
        a.warmup = 1
        a.steps = n_batches - 2
    return a


def gpm_of(a):
    """genomes per merge: 1 on one GPU (the README's one-file-per-genome form); N on N GPUs (weak scaling: the rows each
    GPU walks per step stay the same)"""
    return a.genomes_per_merge if a.genomes_per_merge > 0 else max(1, a.gpus)


def config_of(a, extra=None):
    g = gpm_of(a)
    c = {"workload": "configs[1]: merge-build of synthetic %.1f Mb bacterial genomes (0.5%% subst + 0.05%% indel from a random earlier genome), %d genome(s) = one batch = one merge of %d symbols (both strands); %d genomes in total" % (
        a.genome_len / 1e6, g, g * (2 * a.genome_len + 2), g * (1 + a.warmup + a.steps)),
        "genome_len": a.genome_len, "genomes": g * (1 + a.warmup + a.steps), "genomes_per_merge": g, "seed": SEED,
        "l2": "no explicit flush: every step works on a NEW batch (%d MB) and ~26 bytes per batch symbol of freshly written LF / walk-order / interleave arrays (%d MB per step, above the 126 MB L2), so nothing a step reads was left in L2 by the step before except index cells; the index itself (1 byte per symbol) grows from %d MB to %d MB during the timed steps, i.e. the walk's random cell reads are partly L2 hits while it is below ~126 MB -- the ncu capture in roofline.traffic says how many bytes really came from DRAM" % (
            g * (2 * a.genome_len + 2) // 1000000, 26 * g * (2 * a.genome_len + 2) // 1000000,
            g * (1 + a.warmup) * (2 * a.genome_len + 2) // 1000000, g * (1 + a.warmup + a.steps) * (2 * a.genome_len + 2) // 1000000)}
    if extra:
        c.update(extra)
    return c


class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile(suffix=".csv", delete=False)
        try:
            if os.environ.get("RB3B_BENCH_NO_SAMPLER"):   # debugging aid only: a line without clocks is not a valid bench line
                raise OSError("disabled")
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.f.name):
            t = [x.strip() for x in line.split(",")]
            if len(t) < 8 or not t[0].isdigit() or int(t[0]) != self.idx:
                continue
            try:
                sm.append(float(t[1])); mx.append(float(t[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], t[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def make_genomes(a):
    from ropebwt3_b200 import synth
    return synth.genomes(gpm_of(a) * (1 + a.warmup + a.steps), a.genome_len, seed=SEED, sub=0.005, indel=0.0005)


# --------------------------------------------------------------------------- our arm

def run_b200(a):
    import torch
    import torch.distributed as dist
    import ropebwt3_b200 as R
    from ropebwt3_b200 import synth, capi

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    R.init(local)
    stream = torch.cuda.Stream(priority=-1)   # above the library's second stream (queued merges): the next batch's small kernels go first
    torch.cuda.set_stream(stream)
    capi.check(capi.lib().rb3b_set_stream(stream.cuda_stream))
    from ropebwt3_b200 import dist as rdist
    if world > 1:
        rdist.init_library_comm()   # the library's own NCCL communicator; torch.distributed only carries the id
    if a.seg_len:
        R.set_param("seg_len", a.seg_len)
    for kv in a.param:
        k, v = kv.split("=")
        R.set_param(k, int(v))

    # ---- synthetic input, partial BWTs built on the device (untimed producer of the path's input)
    t0 = time.time()
    gs = make_genomes(a)
    G = gpm_of(a)
    n_g = len(gs) // G          # batches
    d_bwt, h_bwt, lens, bwt_ms = [], [], [], []
    t_bwt = 0.0
    for b in range(n_g):
        text = synth.batch_text(gs[b * G:(b + 1) * G])
        d_text = torch.from_numpy(text).cuda()
        out = torch.empty_like(d_text)
        torch.cuda.synchronize()
        t1 = time.time()
        capi.check(capi.lib().rb3b_build_bwt_dev(len(text), d_text.data_ptr(), out.data_ptr()))
        R.sync()
        t_bwt += time.time() - t1
        bwt_ms.append((time.time() - t1) * 1e3)
        d_bwt.append(out)
        hb = torch.empty(len(text), dtype=torch.uint8).pin_memory()
        hb.copy_(out)
        h_bwt.append(hb)
        lens.append(len(text))
    bases = [sum(len(g) for g in gs[b * G:(b + 1) * G]) for b in range(n_g)]
    # what the finished index must hold (both strands + two sentinels per genome); full parity lives in tests/
    expect = np.zeros(6, np.int64)
    for g in gs:
        cnt = np.bincount(g, minlength=6)[:6]
        expect += cnt + cnt[[0, 4, 3, 2, 1, 5]]
        expect[0] += 2
    if world > 1:
        del gs   # N genomes per step on N ranks: keep the host footprint of every rank down
    t_setup = time.time() - t0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident run
    idx = R.Index.from_plain_dev(d_bwt[0].data_ptr(), lens[0])
    idx.reserve(sum(lens))   # the CLI knows its input size too; see rb3b_index_reserve
    for i in range(1, 1 + a.warmup):
        if world > 1:
            rdist.merge_plain_dist_dev(idx, d_bwt[i].data_ptr(), lens[i])
        else:
            idx.merge_plain_dev(d_bwt[i].data_ptr(), lens[i])
    barrier()
    R.get_stat("reset")
    sampler = ClockSampler(local)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    e0.record(stream)
    n_sharded = 0
    for i in range(1 + a.warmup, n_g):
        if world > 1:   # rank phase split over the ranks + one NCCL all-reduce (MAX) of the interleave array
            n_sharded += int(rdist.merge_plain_dist_dev(idx, d_bwt[i].data_ptr(), lens[i]))
        else:
            idx.merge_plain_dev(d_bwt[i].data_ptr(), lens[i])
    R.sync()   # asynchronous merges run on the library's second stream: the timed region ends when they have
    e1.record(stream)
    barrier()
    wall_dev = time.time() - w0
    ms_dev = e0.elapsed_time(e1)
    st = {k: R.get_stat(k) for k in ["us_prep", "us_walk_first", "us_walk_fix", "us_merge", "us_scatter", "us_comm", "kernel_launches", "n_segments", "fix_rounds", "fix_rounds_total", "n_cells", "n_ovf_cells", "cell_shift", "fix_rows", "fix_wide_rows", "fix_longest_chain"]}
    acc_dev = idx.acc()
    index_bytes = idx.nbytes()
    timed_bases = sum(bases[1 + a.warmup:])
    timed_syms = sum(lens[1 + a.warmup:])

    # ---- end to end through the host-buffer C-ABI call
    ms_e2e, e2e_val, d2h = 0.0, 0.0, 0
    if not a.no_e2e and world > 1:
        # end to end at N > 1: the batch arrives in pinned host memory on every rank and is copied to the device inside
        # the timed region, then the sharded merge; the result read back is the new C[] of the index
        idx2 = R.Index.from_plain(h_bwt[0].numpy())
        idx2.reserve(sum(lens))
        for i in range(1, 1 + a.warmup):
            rdist.merge_plain_dist(idx2, h_bwt[i].data_ptr(), lens[i])
        barrier()
        e0.record(stream)
        w0 = time.time()
        for i in range(1 + a.warmup, n_g):
            rdist.merge_plain_dist(idx2, h_bwt[i].data_ptr(), lens[i])   # H2D of this rank's share + NVLink all-gather + sharded merge
            acc_host = idx2.acc()
        R.sync()   # asynchronous merges run on the library's second stream: the timed region ends when they have
        e1.record(stream)
        barrier()
        wall_e2e = time.time() - w0
        ms_e2e = max(e0.elapsed_time(e1), wall_e2e * 1e3)
        assert np.array_equal(acc_host, acc_dev), "host-buffer and device-pointer builds disagree"
        d2h = 8 * 40
        e2e_val = timed_bases / (ms_e2e / 1e3)
    elif not a.no_e2e:
        idx2 = R.Index.from_plain(h_bwt[0].numpy())
        idx2.reserve(sum(lens))
        for i in range(1, 1 + a.warmup):
            idx2.merge_plain(h_bwt[i].numpy())
        barrier()
        hp = [h.data_ptr() for h in h_bwt]
        L = capi.lib()
        acc_host = np.zeros(7, np.int64)
        e0.record(stream)
        w0 = time.time()
        if not a.no_prefetch:
            capi.check(L.rb3b_prefetch_batch(lens[1 + a.warmup], hp[1 + a.warmup]))
        for i in range(1 + a.warmup, n_g):
            if not a.no_prefetch and i + 1 < n_g:                          # like the reference's -p pipeline (read batch i+1 while merging batch i):
                capi.check(L.rb3b_prefetch_batch(lens[i + 1], hp[i + 1]))  # the H2D copy of the NEXT batch is queued on the library's copy stream
            capi.check(L.rb3b_merge_plain(idx2.h, lens[i], hp[i]))       # (H2D of the batch unless it was copied ahead) + merge + sync
            L.rb3b_get_acc(idx2.h, acc_host.ctypes.data)                   # the step's result: new C[] of the index
        R.sync()   # asynchronous merges run on the library's second stream: the timed region ends when they have
        e1.record(stream)
        barrier()
        wall_e2e = time.time() - w0
        ms_e2e = max(e0.elapsed_time(e1), wall_e2e * 1e3)
        assert np.array_equal(acc_host, acc_dev), "host-buffer and device-pointer builds disagree"
        # D2H per merge: 7 words (batch totals + flag), 2 (chain cover), 1 (worklist size), 4 per fix-up round, 1 (unresolved),
        # 1 (monotone flag), 7 (totals of the new index)
        d2h = 8 * (7 + 2 + 1 + 4 * max(1, st["fix_rounds_total"] // max(1, a.steps)) + 1 + 1 + 7)
        e2e_val = timed_bases / (ms_e2e / 1e3)
    clocks = sampler.stop()

    # ---- sanity of the result (full parity lives in tests/): totals must add up
    assert np.array_equal(np.diff(acc_dev), expect), "symbol totals of the built index are wrong"

    # max over ranks
    if world > 1:
        t = torch.tensor([ms_dev, ms_e2e], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev = float(t[0])
        if not a.no_e2e:
            ms_e2e = float(t[1])
            e2e_val = timed_bases / (ms_e2e / 1e3)

    value = timed_bases / (ms_dev / 1e3)
    # what the metric's name promises -- text in, merged index out: the same steps through the two-step form the CLI uses
    # (rb3b_batch_prepare_dev: device suffix sort, BWT and walk order; rb3b_merge_prepared: walk, fix-up, scatter, merge),
    # batch texts resident in HBM, one stream, nothing overlapped
    ms_build, build_value = None, None
    if world == 1 and not a.no_build:
        texts = [torch.from_numpy(synth.batch_text(gs[b * G:(b + 1) * G])).cuda() for b in range(n_g)]
        idx3 = R.Index()
        for i in range(0, 1 + a.warmup):
            bt = R.Batch.prepare_dev(texts[i].data_ptr(), lens[i]); R.merge_prepared(idx3, bt); bt.close()
        idx3.reserve(sum(lens))
        barrier()
        e0.record(stream)
        for i in range(1 + a.warmup, n_g):
            bt = R.Batch.prepare_dev(texts[i].data_ptr(), lens[i]); R.merge_prepared(idx3, bt); bt.close()
        R.sync()
        e1.record(stream)
        barrier()
        ms_build = e0.elapsed_time(e1)
        build_value = timed_bases / (ms_build / 1e3)
        assert np.array_equal(idx3.acc(), acc_dev), "prepared-batch build and seam build disagree"
        del idx3, texts
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    walk_s = st["us_walk_first"] / 1e6
    # per-launch figures are THIS rank's: at N > 1 a rank walks 1/N of the batch's rows (plus its halo)
    rank_syms = timed_syms / world
    achieved = WALK_ROW_BYTES * rank_syms / walk_s / 1e9 if walk_s > 0 else 0.0
    phase_s = (st["us_prep"] + st["us_walk_first"] + st["us_walk_fix"] + st["us_scatter"]) / 1e6
    phase_gbs = LF_STEP_BYTES * rank_syms / phase_s / 1e9 if phase_s > 0 else 0.0
    traffic, traffic_note = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "walk_first_traffic.json")))
        if world == 1 and G == 1 and a.genome_len == GENOME_LEN:   # the capture is of this workload; anything else has no measured traffic
            traffic = tj.get("dram_bytes_per_launch")
            traffic_note = "STATIC: from the committed ncu --set full capture %s (%s), not measured in this run" % (tj.get("source"), tj.get("what"))
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_dev / a.steps, "higher_is_better": True, "scaling": "strong" if (a.genomes_per_merge > 0 and a.config == "c1") else "weak", "vs_baseline": None,
        "dtype": "u8/int64", "data": "synthetic",
        "config": config_of(a, {"seg_len": a.seg_len or R.get_stat("seg_len_used"), "parallelism": "1 GPU" if world == 1 else "weak scaling: %d genomes per merge on %d GPUs; index replicated; rank phase of every merge split over the ranks (walk-order slices + halo, first walk of the pieces shared by an all-gather), one NCCL all-reduce of the partial interleave arrays (32-bit SUM while positions fit, else 64-bit MAX), merge replicated; %d of %d merges sharded" % (G, world, n_sharded, a.steps)}),
        "e2e": None if a.no_e2e else {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(np.mean(lens[1 + a.warmup:])), "d2h_bytes_per_step": int(d2h),
                                      "ms_per_step": ms_e2e / a.steps,
                                      "how": "every batch goes from pinned host memory through rb3b_merge_plain (host-pointer C-ABI call) and the new C[] of the index is read back, all inside the timed region" +
                                             ("" if (a.no_prefetch or world > 1) else "; the H2D copy of batch i+1 is queued with rb3b_prefetch_batch before the merge of batch i (the reference's -p pipeline overlaps reading and merging the same way)")},
        "gpu_launches": int(st["kernel_launches"]),
        "clocks": clocks,
        "build_value": None if build_value is None else {"value": build_value, "unit": UNIT, "ms_per_step": ms_build / a.steps,
                        "what": "text in, merged index out: the same timed steps through rb3b_batch_prepare_dev (device suffix sort -> partial BWT + walk order) + rb3b_merge_prepared (walk over the index, fix-up, scatter, merge), batch texts resident in HBM; this is what the CLI runs per batch"},
        "roofline": {"kernel": "k_walk_pair (sliced LF walk over the index, two lanes per walk: per row one symbol streamed in, one rank lookup per lane in the bitmap cells, one 8-byte position streamed out, a 16-byte transfer mask for rows still under a bracket)",
                     "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                     "traffic": traffic, "traffic_note": traffic_note, "bytes_per_unit": WALK_ROW_BYTES, "units_per_launch": rank_syms / a.steps,
                     "avg_launch_ms": walk_s * 1e3 / a.steps,
                     "rank_phase": {"bytes_per_unit": LF_STEP_BYTES, "achieved": phase_gbs, "frac": phase_gbs / peak, "ms_per_step": phase_s * 1e3 / a.steps,
                                    "what": "SURVEY 8d's 160 B per LF-walk step over ALL kernels of the rank phase (LF table, list ranking, walk-order rewrite, walk, fix-up, scatter)"},
                     "note": "one dependent random 64-B access per row per walk; with one genome per batch only len/seg_len = 26 k walks exist and the kernel is bound by the instruction + DRAM latency of a step, with ten genomes per batch by DRAM random-access throughput (see DESIGN.md 5); the rank primitive itself is measured under rank_kernel"},
        "phase_ms_per_step": {k[3:]: st[k] / 1e3 / a.steps for k in st if k.startswith("us_")},
        "wall_ms_per_step": wall_dev * 1e3 / a.steps,
        "setup": {"genomes_and_bwt_s": t_setup, "device_bwt_build_s": t_bwt, "bwt_build_bases_per_s": sum(bases) / t_bwt},
        "index": {"symbols": int(acc_dev[6]), "device_bytes": int(index_bytes), "cells": int(st["n_cells"]), "overflow_cells": int(st["n_ovf_cells"]), "cell_span": 1 << int(st["cell_shift"])},
        "walk": {"segments_per_step": st["n_segments"], "fix_rounds_total": st["fix_rounds_total"], "last_step_fix_rows": st["fix_rows"], "last_step_fix_wide_rows": st["fix_wide_rows"], "last_step_fix_longest_chain": st["fix_longest_chain"]},
    }
    if rank == 0 and world == 1 and not a.no_rank_bench:
        del idx, d_bwt
        torch.cuda.empty_cache()
        line["rank_kernel"] = rank_kernel_bench(R, capi, torch, stream, peak)
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(a, gs, idx_state_genomes=1 + a.warmup, budget_s=20.0)
    if rank == 0:
        sys.stdout.flush()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def rank_kernel_bench(R, capi, torch, stream, peak, n_symbols=4_000_000_000, nq=32_000_000):
    """BASELINE.json's second metric: achieved HBM GB/s of the batched rank kernel on independent random queries over an
    index far larger than L2 (4 G symbols = 4 GB of bitmap cells), 144 algorithmic bytes per query."""
    g = torch.Generator(device="cuda").manual_seed(1)
    n_runs = n_symbols // 16
    sym = torch.randint(0, 6, (n_runs,), device="cuda", dtype=torch.uint8, generator=g)
    ln = torch.randint(1, 32, (n_runs,), device="cuda", dtype=torch.int64, generator=g)
    idx = R.Index()
    torch.cuda.synchronize()
    capi.check(capi.lib().rb3b_index_from_runs_device(idx.h, n_runs, sym.data_ptr(), ln.data_ptr()))
    R.sync()
    n = len(idx)
    del sym, ln
    k = torch.randint(0, n, (nq,), device="cuda", dtype=torch.int64, generator=g)
    c = torch.randint(0, 6, (nq,), device="cuda", dtype=torch.uint8, generator=g)
    out = torch.empty(nq, dtype=torch.int64, device="cuda")
    for _ in range(3):
        idx.lf_dev(nq, k.data_ptr(), c.data_ptr(), out.data_ptr(), 0)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        idx.lf_dev(nq, k.data_ptr(), c.data_ptr(), out.data_ptr(), 0)
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    gbs = nq * 144 / (ms / 1e3) / 1e9
    res = {"kernel": "k_lf_bm (one thread per query: two 16-B loads + popcount)", "index_symbols": int(n), "index_bytes": idx.nbytes(),
           "queries": nq, "ms": ms, "gqueries_per_s": nq / ms / 1e6, "bytes_per_query": 144, "achieved": gbs, "unit": "GB/s",
           "peak": peak, "frac": gbs / peak, "frac_of_nominal_8TBps": gbs / 8000.0}
    idx.close()
    return res


# --------------------------------------------------------------------------- CPU reference

def ref_index_of(gs, n_threads):
    """Index of the given genomes built by the reference itself (libsais + rb3_enc_plain2fmr)."""
    from oracle import ref
    from ropebwt3_b200 import synth
    text = synth.batch_text(gs)
    bwt = ref.build_sais(text, 2 * len(gs), n_threads)
    return ref.Rope.from_plain(bwt, n_threads)


def cpu_baseline(a, gs, idx_state_genomes, budget_s):
    """The reference's rb3_fmi_merge_plain (oracle/_ref/librb3ref.so) on the host cores, on a bounded sample of the SAME
    workload: whole genomes, same batching, as many merges as fit the budget.  The .fmd of what it built is compared with
    the .fmd our engine builds from the same genomes (`parity_checked`)."""
    from oracle import ref
    from ropebwt3_b200 import synth
    import ropebwt3_b200 as R
    cores = os.cpu_count() or 1
    if not ref.available():
        return port_baseline(a, gs, budget_s)
    G = gpm_of(a)
    rope = ref_index_of(gs[:G], cores)
    done_bases, t_used, n_merge, last = 0, 0.0, 0, G
    for b in range(G, len(gs) - G + 1, G):
        bwt = ref.build_sais(synth.batch_text(gs[b:b + G]), 2 * G, cores)
        t0 = time.time()
        rope.merge_plain(bwt, cores)
        t_used += time.time() - t0
        done_bases += sum(len(x) for x in gs[b:b + G])
        n_merge += 1
        last = b + G
        if t_used > budget_s:
            break
    want = rope.to_fmd()   # consumes the rope
    # our engine on the same genomes, same batching
    idx = None
    for b in range(0, last, G):
        bwt = R.rb3_build_sais(synth.batch_text(gs[b:b + G]))
        if idx is None:
            idx = R.Index.from_plain(bwt)
        else:
            idx.merge_plain(bwt)
    s_, l_ = idx.export_runs()
    same = R.fmd_image(s_, l_) == want
    idx.close()
    return {"value": done_bases / t_used, "unit": UNIT, "cores": cores, "kind": "reference",
            "parity_checked": {"fmd_identical": bool(same), "genomes": last, "fmd_bytes": len(want),
                               "what": "the .fmd of the index the reference built in this sample vs the .fmd of our engine's index of the same genomes, byte for byte"},
            "sample": "%d merges (rb3_fmi_merge_plain, n_threads=%d) of %d whole genome(s) each (%.1f Mb, both strands) into the reference's index of the first %d genome(s); %.1f s of CPU wall time; only %d chains per merge exist, so the reference's rank phase cannot use more than %d threads here" % (
                n_merge, cores, G, a.genome_len / 1e6, G, t_used, 2 * G, 2 * G)}


def port_baseline(a, gs, budget_s):
    from oracle import oracle as O
    from ropebwt3_b200 import synth
    frag = 200_000
    sym, ln = O.plain2runs(O.build_bwt(synth.batch_text([gs[0][:frag]])))
    bwt = O.build_bwt(synth.batch_text([gs[1][:frag]]))
    t0 = time.time()
    O.merge_plain(sym, ln, bwt)
    dt = time.time() - t0
    return {"value": frag / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "one merge of a %d bp genome prefix into a %d bp one with the scalar oracle port (oracle/_ref absent)" % (frag, frag)}


def run_reference(a):
    """The reference arm: the unmodified reference's rb3_fmi_merge_plain (oracle/_ref) on the host cores, on the b200 arm's
    workload -- the same seeded genomes, the same batching, whole genomes, no genome inserted twice.  If the requested steps
    do not fit the time budget the number of TIMED steps shrinks (reported in `steps`), never the genomes."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle import ref
    from ropebwt3_b200 import synth
    cores = os.cpu_count() or 1
    G = gpm_of(a)
    if not ref.available():
        gs = synth.genomes(2, a.genome_len, seed=SEED, sub=0.005, indel=0.0005)
        b = port_baseline(a, gs, 20.0)
        line = {"impl": "reference", "metric": METRIC, "value": b["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int64", "data": "synthetic",
                "config": config_of(a), "cpu_baseline": b, "e2e": {"value": b["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return
    gen = synth.genome_stream(a.genome_len, seed=SEED, sub=0.005, indel=0.0005)   # the b200 arm's genomes, produced as needed
    take = lambda k: [next(gen) for _ in range(k)]
    rope = ref_index_of(take(G), cores)
    times, nb, t_all = [], 0, 0.0
    for i in range(a.warmup + a.steps):
        pieces = take(G)
        bwt = ref.build_sais(synth.batch_text(pieces), 2 * G, cores)
        t0 = time.time()
        rope.merge_plain(bwt, cores)
        dt = time.time() - t0
        t_all += dt
        if i >= a.warmup:
            times.append(dt)
            nb += sum(len(x) for x in pieces)
        # stop early when the remaining steps would not fit the budget (at least 3 timed steps)
        if len(times) >= 3 and t_all + dt > a.ref_budget_s:
            break
    rope.close()
    tot = sum(times)
    val = nb / tot
    sample = "each step = rb3_fmi_merge_plain(n_threads=%d) of the next %d whole genome(s) of the b200 arm's set (%.1f Mb each, both strands, %d symbols) into the reference's own index (first %d genome(s) + earlier steps); %d warm-up + %d timed steps%s" % (
        cores, G, a.genome_len / 1e6, G * (2 * a.genome_len + 2), G, a.warmup, len(times),
        "" if len(times) == a.steps else " (of the %d requested: the rest did not fit --ref-budget-s %d)" % (a.steps, int(a.ref_budget_s)))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": len(times), "warmup": a.warmup,
            "ms_per_step": tot * 1e3 / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/int64", "data": "synthetic", "config": config_of(a),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


if __name__ == "__main__":
    args = apply_config(parse())
    # stdout carries the one JSON line and nothing else: libraries that print to fd 1 (the NCCL banner) go to stderr
    sys.stdout.flush()
    _real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = _real_stdout
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
