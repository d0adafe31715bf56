"""Seeded synthetic genome sets: a uniform ACGT ancestor, every later
genome derived from a uniformly chosen earlier one with substitutions and 1-bp
indels.  Data generation only -- not part of the hot path."""
import numpy as np

COMP = np.array([0, 4, 3, 2, 1, 5], np.uint8)  # nt6 complement, io.c:30


def mutate(rng, g, sub=0.005, indel=0.0005):
    g = g.copy()
    n = len(g)
    m = rng.random(n) < sub
    g[m] = (g[m] - 1 + rng.integers(1, 4, int(m.sum()))) % 4 + 1  # a different base
    if indel > 0:
        d = rng.random(n) < indel / 2
        g = g[~d]
        ins = np.flatnonzero(rng.random(len(g)) < indel / 2)
        g = np.insert(g, ins, rng.integers(1, 5, len(ins)).astype(np.uint8))
    return g


def genomes(n_genomes, length, seed=43, sub=0.005, indel=0.0005):
    """-> list of uint8 nt6 arrays (values 1..4)."""
    rng = np.random.default_rng(seed)
    out = [rng.integers(1, 5, length).astype(np.uint8)]
    for i in range(1, n_genomes):
        out.append(mutate(rng, out[int(rng.integers(0, i))], sub, indel))
    return out


def genome_stream(length, seed=43, sub=0.005, indel=0.0005):
    """The same sequence of genomes as genomes(), one at a time (every earlier genome is kept: later ones derive from them)."""
    rng = np.random.default_rng(seed)
    out = [rng.integers(1, 5, length).astype(np.uint8)]
    yield out[0]
    i = 1
    while True:
        out.append(mutate(rng, out[int(rng.integers(0, i))], sub, indel))
        yield out[-1]
        i += 1


def batch_text(gs, both=True):
    """Concatenate genomes as the reference reader does (io.c:84-102): forward strand,
    0, reverse complement, 0."""
    parts = []
    for g in gs:
        parts += [g, np.zeros(1, np.uint8)]
        if both:
            parts += [COMP[g[::-1]], np.zeros(1, np.uint8)]
    return np.concatenate(parts)


def random_runs(rng, n_runs, max_len=40, big_every=0):
    """A random coalesced run list (sym, len) for index-level tests."""
    sym = rng.integers(0, 6, n_runs).astype(np.uint8)
    for i in range(1, n_runs):
        if sym[i] == sym[i - 1]:
            sym[i] = (sym[i] + 1 + rng.integers(0, 5)) % 6
    ln = rng.integers(1, max_len + 1, n_runs).astype(np.int64)
    if big_every:
        idx = np.arange(0, n_runs, big_every)
        ln[idx] = rng.integers(1, 1 << 26, len(idx))
    return sym, ln
