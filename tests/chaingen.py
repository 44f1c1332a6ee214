"""Synthetic stage-2 chaining problems: what Oligoindex_get_mappings +
Diag_compute_bounds hand to align_compute_lookback -- per query position the sorted genomic positions of its
k-mer, plus the active window [minactive, maxactive] -- built from a random genome and a spliced, mutated
transcript of it.  Pure numpy; used by the CPU and GPU parity tests and by bench.py."""
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _kmer_codes(seq, k):
    """integer code of every k-mer of a uint8 array of 0..3 (length len-k+1)"""
    n = len(seq) - k + 1
    if n <= 0:
        return np.zeros(0, dtype=np.int64)
    code = np.zeros(n, dtype=np.int64)
    for i in range(k):
        code = code * 4 + seq[i:i + n]
    return code


def make_problem(rng, glen=20000, nexons=4, exon_len=(60, 400), err=0.02, k=8, repeat=0, window="full",
                 localp=True, skip_repetitive_p=True, middlep=True, max_nalignments=10, subrange=False):
    """Returns a dict with the arrays of one problem.  `repeat` > 0 plants a low-complexity stretch that gives
    some k-mers >= 100 hits (exercises MAX_NACTIVE / MAX_SKIPPED, stage2.c:43,86)."""
    g = rng.integers(0, 4, size=glen).astype(np.uint8)
    if repeat:
        unit = rng.integers(0, 4, size=int(rng.integers(2, 6))).astype(np.uint8)
        for _ in range(repeat):
            a = int(rng.integers(0, glen - 400))
            ln = int(rng.integers(150, 400))
            g[a:a + ln] = np.resize(unit, ln)
    # exons: increasing starts
    lens = rng.integers(exon_len[0], exon_len[1] + 1, size=nexons)
    span = glen - int(lens.sum()) - 10
    gaps = np.sort(rng.integers(0, max(span, 1), size=nexons))
    starts, pos = [], 0
    for e in range(nexons):
        s = int(gaps[e]) + pos
        starts.append(s)
        pos += int(lens[e])
    q = np.concatenate([g[s:s + int(l)] for s, l in zip(starts, lens)])
    # mutate: substitutions, deletions, insertions (60/20/20)
    out = []
    r = rng.random(len(q))
    kind = rng.random(len(q))
    sub = rng.integers(0, 4, size=len(q))
    for i in range(len(q)):
        if r[i] < err:
            if kind[i] < 0.6:
                out.append(sub[i])
            elif kind[i] < 0.8:
                continue
            else:
                out.append(q[i]); out.append(sub[i])
        else:
            out.append(q[i])
    q = np.array(out, dtype=np.uint8)
    L = len(q)
    gc = _kmer_codes(g, k)
    order = np.argsort(gc, kind="stable")
    gs = gc[order]
    qc = _kmer_codes(q, k)
    lo = np.searchsorted(gs, qc, side="left")
    hi = np.searchsorted(gs, qc, side="right")
    npos = np.zeros(L, dtype=np.int32)
    npos[:len(qc)] = (hi - lo).astype(np.int32)
    chunks = [np.sort(order[a:b]) for a, b in zip(lo, hi)]
    positions = np.concatenate(chunks).astype(np.uint32) if chunks else np.zeros(0, dtype=np.uint32)
    if window == "full":
        mina = np.zeros(L, dtype=np.uint32)
        maxa = np.full(L, glen, dtype=np.uint32)
    else:   # a band around the transcript's span, as Diag_compute_bounds produces
        margin = int(window)
        mina = np.full(L, max(0, starts[0] - margin), dtype=np.uint32)
        maxa = np.full(L, min(glen, starts[-1] + int(lens[-1]) + margin), dtype=np.uint32)
    qs, qe = 0, L - 1
    if subrange and L > 40:
        qs = int(rng.integers(0, L // 4))
        qe = int(rng.integers(3 * L // 4, L))
    seq = ACGT[q].tobytes()
    return {"positions": positions, "npositions": npos, "minactive": mina, "maxactive": maxa, "querylength": L,
            "querystart": qs, "queryend": qe, "indexsize": k, "localp": int(localp), "skip_repetitive_p": int(skip_repetitive_p),
            "favor_right_p": 0, "middlep": int(middlep), "max_nalignments": max_nalignments, "queryseq": seq}


def make_set(seed, n, small=False):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        glen = int(rng.integers(3000, 8000)) if small else int(rng.integers(20000, 120000))
        out.append(make_problem(
            rng, glen=glen, nexons=int(rng.integers(1, 4 if small else 9)),
            exon_len=(40, 150) if small else (60, 500),
            err=float(rng.choice([0.0, 0.01, 0.03, 0.08])),
            k=int(rng.choice([8, 8, 7, 6])) if small else 8,
            repeat=int(rng.integers(0, 3)) if i % 3 == 0 else 0,
            window="full" if i % 2 == 0 else str(int(rng.integers(50, 3000))),
            localp=bool(i % 5 != 4), skip_repetitive_p=bool(i % 7 != 6), middlep=bool(i % 3 != 1),
            max_nalignments=int(rng.choice([1, 10])), subrange=bool(i % 4 == 3)))
    return out
