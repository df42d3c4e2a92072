"""Index builder / loader of the library (host code): on-disk format identical to the reference's builder, and
FM-index invariants that hold without any reference."""
import filecmp
import os
import subprocess

import numpy as np
import pytest

import parity_util as pu
import ref_oracle
from mapcaller_b200 import api, simulate as sim


def _fa(tmp_path, contigs):
    p = str(tmp_path / "g.fa")
    sim.write_fasta(p, contigs)
    return p


@pytest.mark.skipif(not os.path.exists(ref_oracle.BIN_PATH), reason="oracle/_ref not built")
@pytest.mark.parametrize("kind", ["single", "multi_with_N", "tiny", "repeats"])
def test_index_files_are_byte_identical_to_the_reference_builder(built, tmp_path, kind):
    if kind == "single":
        contigs = [("chr1", sim.genome(50_000, 1, n_dup=5))]
    elif kind == "multi_with_N":
        a, b, c = sim.genome(30_000, 2), sim.genome(12_345, 3), sim.genome(4_001, 4)
        a[100:160] = ord("N"); b[0:7] = ord("N"); b[5000] = ord("R"); c[-9:] = ord("N")
        contigs = [("ctgA", a), ("ctgB some comment", b), ("ctgC", c)]
    elif kind == "tiny":
        contigs = [("t", sim.genome(131, 5))]
    else:
        g = sim.genome(20_000, 6, tandem=30); g[3000:9000] = np.tile(g[3000:3003], 2000)
        contigs = [("rep", g)]
    fa = _fa(tmp_path, contigs)
    subprocess.check_call([ref_oracle.BIN_PATH, "index", fa, str(tmp_path / "ref")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    ix = api.Index.build_fasta(fa, threads=4)
    ix.save(str(tmp_path / "mine"))
    for ext in ("bwt", "sa", "pac", "ann", "amb"):
        assert filecmp.cmp(str(tmp_path / ("mine." + ext)), str(tmp_path / ("ref." + ext)), shallow=False), ext
    # and the loader reads the reference's files back into the same image
    v1, v2 = ix.view(), api.Index.load(str(tmp_path / "ref")).view()
    assert (v1.primary, list(v1.L2), v1.seq_len, v1.n_sa, v1.genome_size, v1.n_chrom) == (v2.primary, list(v2.L2), v2.seq_len, v2.n_sa, v2.genome_size, v2.n_chrom)


def _arrays(ix):
    import ctypes as C
    v = ix.view()
    bwt = np.ctypeslib.as_array(C.cast(v.bwt, C.POINTER(C.c_uint32)), shape=(v.bwt_size,)).copy()
    sa = np.ctypeslib.as_array(C.cast(v.sa, C.POINTER(C.c_uint64)), shape=(v.n_sa,)).copy()
    return v, bwt, sa


def test_fm_index_invariants(built):
    """Independent of any reference: the sampled SA values are exactly the suffix ranks of a naive sort."""
    g = sim.genome(3000, 11, tandem=3)
    codes = sim.encode(g)
    ix = api.Index.build(codes)
    v, bwt, sa = _arrays(ix)
    text = np.concatenate([codes, 3 - codes[::-1]])
    n = len(text)
    assert v.seq_len == n and v.genome_size == len(g)
    s = bytes(text.tolist())
    order = sorted(range(n), key=lambda i: s[i:])           # naive suffix array
    rows = [n] + order                                       # row 0 is the empty suffix
    assert sa[0] == np.uint64(2**64 - 1)
    for j in range(1, len(sa)):
        assert sa[j] == rows[32 * j]
    assert rows[v.primary] == 0
    # the BWT symbols: unpack the 64-byte blocks and compare with text[rows[r]-1]
    want = [text[r - 1] for r in rows if r != 0]
    got = []
    for k in range(n):
        blk = (k >> 7) * 16
        w = bwt[blk + 8 + ((k & 127) >> 4)]
        got.append((int(w) >> ((~k & 15) << 1)) & 3)
    assert got == [int(x) for x in want]
    counts = np.bincount(text, minlength=4)
    assert list(v.L2) == [0] + list(np.cumsum(counts))
    # running counts at the head of every block
    run = np.zeros(4, dtype=np.int64)
    for k in range(0, n, 128):
        blk = (k >> 7) * 16
        c = bwt[blk:blk + 8].view(np.uint64)
        assert list(c) == list(run)
        run += np.bincount(got[k:k + 128], minlength=4)


def test_save_load_roundtrip(built, tmp_path):
    case = pu.make_case(seed=2, n_pairs=10, genome_len=20000, contigs=2)
    ix = pu.build_index(case)
    ix.save(str(tmp_path / "i"))
    jx = api.Index.load(str(tmp_path / "i"))
    a, b = _arrays(ix), _arrays(jx)
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert a[0].primary == b[0].primary and list(a[0].L2) == list(b[0].L2)
