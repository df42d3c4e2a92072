"""Seeded synthetic inputs for the configs of BASELINE.json.

Everything here is numpy-vectorised so that the E. coli-sized config (4.6 Mbp, 1.15 M pairs) is
generated in seconds inside bench.py and the tests.  Nothing in this module touches the GPU or the
oracle; it only fabricates genomes, mutated genomes and paired-end reads:

* genome():        uniform ACGT contigs + copied segments / repeat families (multi-hit, tandem and
                   OCC_Thr paths of the reference, src/bwt_search.cpp:153-161, src/ReadMapping.cpp:209)
* mutate():        seeded restatement of the variant *rates* of the reference's stand-alone
                   simulator (reference src/sv_simulator/SVsim.cpp:14-21) - SNP, small/large indel,
                   inversion, translocation, tandem duplication
* simulate_pairs(): 2xL reads, fragment ~ N(mu, sd), uniform substitution errors, 50 % strand flip,
                   optional per-base indel errors (config 5)
* FASTA/FASTQ writers in the shape the reference's reader expects (LF, trailing newline,
                   src/GetData.cpp:32-83)
"""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_CODE = np.full(256, 4, dtype=np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _CODE[_c] = _i
    _CODE[_c + 32] = _i
_COMP = np.full(256, ord("N"), dtype=np.uint8)
for _a, _b in zip(b"ACGTacgt", b"TGCATGCA"):
    _COMP[_a] = _b


def revcomp(a: np.ndarray) -> np.ndarray:
    """Reverse complement of an ASCII uint8 array (last axis)."""
    return _COMP[a[..., ::-1]]


def encode(a: np.ndarray) -> np.ndarray:
    """ASCII -> 0..3 (4 = ambiguous), the nst_nt4_table mapping (reference src/BWT_Index/bntseq.c:40-57)."""
    return _CODE[a]


def genome(length: int, seed: int, n_dup: int = 0, dup_len=(1000, 3000), repeat_frac: float = 0.0,
           families=((300, 0.10), (1000, 0.08), (6000, 0.05)), tandem: int = 0) -> np.ndarray:
    """One contig of `length` ASCII bases."""
    rng = np.random.default_rng(seed)
    g = _ACGT[rng.integers(0, 4, size=length, dtype=np.uint8)]
    if repeat_frac > 0:
        # interspersed repeat families: diverged copies of a few consensus sequences
        target = int(length * repeat_frac)
        placed = 0
        cons = [_ACGT[rng.integers(0, 4, size=fl, dtype=np.uint8)] for fl, _ in families]
        while placed < target:
            k = int(rng.integers(0, len(families)))
            fl, div = families[k]
            pos = int(rng.integers(0, length - fl))
            copy = cons[k].copy()
            m = rng.random(fl) < div * rng.uniform(0.5, 1.5)
            copy[m] = _ACGT[rng.integers(0, 4, size=int(m.sum()), dtype=np.uint8)]
            g[pos:pos + fl] = copy
            placed += fl
    for _ in range(n_dup):
        ln = int(rng.integers(dup_len[0], dup_len[1] + 1))
        src = int(rng.integers(0, length - ln))
        dst = int(rng.integers(0, length - ln))
        seg = g[src:src + ln].copy()
        if rng.random() < 0.3:
            seg = revcomp(seg)
        g[dst:dst + ln] = seg
    for _ in range(tandem):
        unit = int(rng.integers(2, 60))
        copies = int(rng.integers(4, 40))
        pos = int(rng.integers(0, length - unit * copies))
        g[pos:pos + unit * copies] = np.tile(g[pos:pos + unit], copies)
    return g


def mutate(g: np.ndarray, seed: int, snp_per_mb: float = 3000, small_indel_per_mb: float = 200,
           large_indel_per_mb: float = 50, sv_per_mb: float = 1, sv_len=(1000, 2000)):
    """Returns (mutant, truth) where truth is a list of (pos0, kind, ref, alt) on the input genome."""
    rng = np.random.default_rng(seed)
    n = len(g)
    mb = n / 1e6
    out = g.copy()
    truth = []
    # SNPs in place (complement base, as the reference's simulator does)
    n_snp = rng.poisson(snp_per_mb * mb)
    snp_pos = np.unique(rng.integers(0, n, size=n_snp))
    out[snp_pos] = _COMP[g[snp_pos]]
    # structural edits, applied right-to-left so coordinates stay valid
    ev = []
    for _ in range(rng.poisson(small_indel_per_mb * mb)):
        ev.append((int(rng.integers(100, n - 100)), "indel", int(min(10, rng.geometric(0.5)))))
    for _ in range(rng.poisson(large_indel_per_mb * mb)):
        ev.append((int(rng.integers(100, n - 100)), "indel", int(rng.integers(11, 31))))
    for _ in range(rng.poisson(sv_per_mb * mb)):
        ev.append((int(rng.integers(5000, max(5001, n - 5000))), ("inv", "tnl", "dup")[int(rng.integers(0, 3))],
                   int(rng.integers(sv_len[0], sv_len[1] + 1))))
    ev.sort(key=lambda e: e[0])
    keep, last_end = [], -1
    for p, kind, ln in ev:  # drop overlapping events
        if p > last_end + 50 and p + ln + 50 < n:
            keep.append((p, kind, ln))
            last_end = p + ln
    pieces = []
    cursor = n
    moved = []
    for p, kind, ln in reversed(keep):
        if kind == "indel":
            if rng.random() < 0.5:  # deletion of ln bases after p
                pieces.append(out[p + ln:cursor]); cursor = p
                truth.append((p, "DEL", ln))
            else:
                ins = _ACGT[rng.integers(0, 4, size=ln, dtype=np.uint8)]
                pieces.append(out[p:cursor]); pieces.append(ins); cursor = p
                truth.append((p, "INS", ln))
        elif kind == "inv":
            pieces.append(out[p + ln:cursor]); pieces.append(revcomp(out[p:p + ln])); cursor = p
            truth.append((p, "INV", ln))
        elif kind == "dup":
            pieces.append(out[p:cursor]); pieces.append(out[p:p + ln].copy()); cursor = p
            truth.append((p, "DUP", ln))
        else:  # translocation: cut here, paste later at a random earlier cut point
            pieces.append(out[p + ln:cursor]); cursor = p
            moved.append(out[p:p + ln].copy())
            truth.append((p, "TNL", ln))
    pieces.append(out[0:cursor])
    pieces.reverse()
    for seg in moved:
        k = int(rng.integers(0, len(pieces) + 1))
        pieces.insert(k, seg)
    truth.extend((int(p), "SNP", 1) for p in snp_pos)
    return np.concatenate(pieces), truth


def mutate_fast(g: np.ndarray, seed: int, snp_per_mb: float = 1000, small_indel_per_mb: float = 100, large_indel_per_mb: float = 0) -> np.ndarray:
    """SNPs and indels at the rates of mutate(), vectorised for genomes of gigabases (no structural variants, no truth
    list): complement-base SNPs in place, then every indel site either loses the `len` bases after it or gains `len`
    random bases before it (small: 1 + geometric, capped at 10; large: 11-30); sites closer than 64 bases are thinned."""
    rng = np.random.default_rng(seed)
    n = len(g); mb = n / 1e6
    out = g.copy()
    snp = rng.integers(0, n, size=rng.poisson(snp_per_mb * mb))
    out[snp] = _COMP[out[snp]]
    n_small, n_large = rng.poisson(small_indel_per_mb * mb), rng.poisson(large_indel_per_mb * mb)
    pos = rng.integers(100, n - 100, size=n_small + n_large)
    ln = np.concatenate([np.minimum(10, rng.geometric(0.5, size=n_small)), rng.integers(11, 31, size=n_large)]).astype(np.int64)
    order = np.argsort(pos, kind="stable"); pos, ln = pos[order], ln[order]
    keep = np.concatenate([[True], np.diff(pos) > 64]); pos, ln = pos[keep], ln[keep]
    is_del = rng.random(len(pos)) < 0.5
    dpos, dlen = pos[is_del], ln[is_del]
    mask = np.ones(n, dtype=bool)
    if len(dpos):
        idx = np.repeat(dpos, dlen) + (np.arange(int(dlen.sum())) - np.repeat(np.cumsum(dlen) - dlen, dlen))
        mask[idx] = False
    ipos, ilen = pos[~is_del], ln[~is_del]
    if len(ipos):
        ins_at = np.repeat(ipos, ilen)
        bases = _ACGT[rng.integers(0, 4, size=len(ins_at), dtype=np.uint8)]
        # np.insert works on the pre-deletion coordinates; flag the inserted bases as kept
        out = np.insert(out, ins_at, bases); mask = np.insert(mask, ins_at, True)
    return out[mask]


def simulate_pairs(g: np.ndarray, n_pairs: int, read_len: int, seed: int, frag_mean: float = 400, frag_sd: float = 40,
                   sub_rate: float = 0.005, indel_rate: float = 0.0, n_rate: float = 0.0):
    """Returns (r1, r2): two uint8 arrays [n_pairs, read_len] of ASCII bases, mates as a sequencer
    reports them (mate 2 is the reverse strand of the fragment)."""
    rng = np.random.default_rng(seed)
    n = len(g)
    flen = np.clip(np.rint(rng.normal(frag_mean, frag_sd, size=n_pairs)), read_len + 10, None).astype(np.int64)
    flen = np.minimum(flen, n - 1)
    start = (rng.random(n_pairs) * (n - flen)).astype(np.int64)
    idx = np.arange(read_len, dtype=np.int64)
    pad = 0
    if indel_rate > 0:
        pad = 16
    L = read_len + pad
    left = g[np.minimum(start[:, None] + np.arange(L)[None, :], n - 1)]
    right_beg = start + flen - L
    right_beg = np.maximum(right_beg, 0)
    right = g[np.minimum(right_beg[:, None] + np.arange(L)[None, :], n - 1)]
    m1 = left
    m2 = revcomp(right)          # starts at the fragment's far end
    if indel_rate > 0:
        m1 = _apply_read_indels(m1, read_len, indel_rate, rng)
        m2 = _apply_read_indels(m2, read_len, indel_rate, rng)
    else:
        m1 = m1[:, :read_len]; m2 = m2[:, :read_len]
    flip = rng.random(n_pairs) < 0.5
    r1 = np.where(flip[:, None], m2, m1)
    r2 = np.where(flip[:, None], m1, m2)
    for r in (r1, r2):
        if sub_rate > 0:
            m = rng.random(r.shape) < sub_rate
            k = int(m.sum())
            # substitute with a different base
            r[m] = _ACGT[(encode(r[m]) + rng.integers(1, 4, size=k, dtype=np.uint8)) & 3]
        if n_rate > 0:
            r[rng.random(r.shape) < n_rate] = ord("N")
    del idx
    return np.ascontiguousarray(r1), np.ascontiguousarray(r2)


def _apply_read_indels(m: np.ndarray, read_len: int, rate: float, rng) -> np.ndarray:
    """Per-base insertion/deletion errors; rows are rebuilt one by one only for the affected reads."""
    out = np.ascontiguousarray(m[:, :read_len]).copy()
    hit = rng.random(m.shape) < rate
    rows = np.nonzero(hit.any(axis=1))[0]
    for r in rows:
        src = m[r]
        buf = []
        for j in range(len(src)):
            if hit[r, j]:
                if rng.random() < 0.5:
                    continue  # deletion
                buf.append(int(_ACGT[rng.integers(0, 4)]))  # insertion before j
            buf.append(int(src[j]))
            if len(buf) >= read_len:
                break
        row = np.array(buf[:read_len], dtype=np.uint8)
        if len(row) < read_len:
            row = np.concatenate([row, src[:read_len - len(row)]])
        out[r] = row
    return out


def write_fasta(path: str, contigs, width: int = 70) -> None:
    """contigs: list of (name, uint8 ASCII array)."""
    with open(path, "wb") as fh:
        for name, seq in contigs:
            fh.write(b">" + name.encode() + b"\n")
            n = len(seq)
            full = n // width * width
            if full:
                body = np.empty((full // width, width + 1), dtype=np.uint8)
                body[:, :width] = seq[:full].reshape(-1, width)
                body[:, width] = 10
                fh.write(body.tobytes())
            if n > full:
                fh.write(seq[full:].tobytes() + b"\n")


def read_fasta(path: str):
    contigs, name, chunks = [], None, []
    with open(path, "rb") as fh:
        for line in fh:
            if line.startswith(b">"):
                if name is not None:
                    contigs.append((name, np.frombuffer(b"".join(chunks), dtype=np.uint8).copy()))
                name, chunks = line[1:].split()[0].decode(), []
            else:
                chunks.append(line.strip())
    if name is not None:
        contigs.append((name, np.frombuffer(b"".join(chunks), dtype=np.uint8).copy()))
    return contigs


def fastq_text(reads: np.ndarray, mate: int, prefix: str = "r", first: int = 0) -> np.ndarray:
    """The bytes of a FASTQ file with fixed-width records `@r000000123/1` (record i is read number first + i), built as
    one numpy byte matrix."""
    n, L = reads.shape
    pre = np.frombuffer(("@" + prefix).encode(), dtype=np.uint8)
    w = len(pre) + 9 + 2
    rec = np.empty((n, w + 1 + L + 1 + 2 + L + 1), dtype=np.uint8)
    rec[:, :len(pre)] = pre
    idx = np.arange(first, first + n, dtype=np.int64)
    for k in range(9):                      # zero-padded decimal digits, most significant first
        rec[:, len(pre) + k] = (idx // 10 ** (8 - k)) % 10 + 48
    rec[:, w - 2] = ord("/")
    rec[:, w - 1] = 48 + mate
    rec[:, w] = 10
    rec[:, w + 1:w + 1 + L] = reads
    rec[:, w + 1 + L] = 10
    rec[:, w + 2 + L] = ord("+")
    rec[:, w + 3 + L] = 10
    rec[:, w + 4 + L:w + 4 + 2 * L] = ord("I")
    rec[:, w + 4 + 2 * L] = 10
    return rec.reshape(-1)


def simulate_pairs_fast(g: np.ndarray, n_pairs: int, read_len: int, seed: int, frag_mean: float = 400, frag_sd: float = 40,
                        sub_rate: float = 0.005, block: int = 250_000, first_block: int = 0, indel_read_frac: float = 0.0):
    """The same read model as simulate_pairs() without N errors, for libraries of millions of pairs: generated in
    independent blocks of `block` pairs (block b depends on (seed, b) only, so any prefix or slice of a library can be
    regenerated), rows gathered through a strided window view, substitution sites drawn by count instead of a mask.
    indel_read_frac: this fraction of the reads carries ONE sequencing indel (1-3 bases inserted or deleted at a random
    offset; the vectorised stand-in for simulate_pairs()'s per-base indel errors)."""
    n = len(g)
    pad = 4 if indel_read_frac > 0 else 0
    win = np.lib.stride_tricks.sliding_window_view(g, read_len + pad)
    r1 = np.empty((n_pairs, read_len), dtype=np.uint8); r2 = np.empty((n_pairs, read_len), dtype=np.uint8)
    for b0 in range(0, n_pairs, block):
        m = min(block, n_pairs - b0)
        rng = np.random.default_rng([seed, first_block + b0 // block])
        flen = np.clip(np.rint(rng.normal(frag_mean, frag_sd, size=m)), read_len + 10, None).astype(np.int64)
        flen = np.minimum(flen, n - 1)
        start = (rng.random(m) * (n - flen)).astype(np.int64)
        flen = np.minimum(flen, n - 1 - pad)
        start = np.minimum(start, n - flen - pad)
        left = win[start]
        right = _COMP[win[np.maximum(start + flen - read_len - pad, 0)][:, ::-1]]
        if pad:
            left, right = _one_indel_per_read(left, read_len, indel_read_frac, rng), _one_indel_per_read(right, read_len, indel_read_frac, rng)
        flip = rng.random(m) < 0.5
        a = np.where(flip[:, None], right, left); c = np.where(flip[:, None], left, right)
        for r in (a, c):
            k = int(rng.binomial(r.size, sub_rate)) if sub_rate > 0 else 0
            if k:
                pos = rng.integers(0, r.size, size=k)
                flat = r.reshape(-1)
                flat[pos] = _ACGT[(_CODE[flat[pos]] + rng.integers(1, 4, size=k, dtype=np.uint8)) & 3]
        r1[b0:b0 + m] = a; r2[b0:b0 + m] = c
    return r1, r2


def _one_indel_per_read(rows: np.ndarray, read_len: int, frac: float, rng) -> np.ndarray:
    """rows [m, read_len + pad]: a fraction of the rows gets 1-3 bases deleted or inserted at a random offset (index
    arithmetic on the padded window, no per-row loop); returns [m, read_len]."""
    m = rows.shape[0]
    hit = rng.random(m) < frac
    at = rng.integers(10, read_len - 10, size=m); ln = rng.integers(1, 4, size=m); dele = rng.random(m) < 0.5
    j = np.arange(read_len, dtype=np.int32)[None, :]
    shift = np.where(hit & dele, ln, np.where(hit & ~dele, -ln, 0)).astype(np.int32)[:, None]
    a = at.astype(np.int32)[:, None]
    # deletion: columns from `at` on read ln bases further right; insertion: columns at..at+ln-1 are new bases, later ones shift left
    src = np.where(j >= a, np.where(shift < 0, np.maximum(j + shift, a), j + shift), j)
    out = np.take_along_axis(rows, src.astype(np.int64), axis=1)
    new = (hit & ~dele)[:, None] & (j >= a) & (j < a + ln[:, None])
    k = int(new.sum())
    if k:
        out[new] = _ACGT[rng.integers(0, 4, size=k, dtype=np.uint8)]
    return out


def write_fastq(path: str, reads: np.ndarray, mate: int, prefix: str = "r") -> None:
    with open(path, "wb") as fh:
        fh.write(fastq_text(reads, mate, prefix).tobytes())


def interleave(r1: np.ndarray, r2: np.ndarray):
    """(seq_bytes, offsets) in the reference's chunk order: mates adjacent, mate 1 first
    (reference src/GetData.cpp:85-99)."""
    n, L = r1.shape
    both = np.empty((2 * n, L), dtype=np.uint8)
    both[0::2] = r1
    both[1::2] = r2
    off = np.arange(2 * n + 1, dtype=np.int64) * L
    return both.reshape(-1), off
