import numpy as np

from mapcaller_b200 import simulate as sim


def test_generators_are_seeded_and_shaped():
    g1, g2 = sim.genome(20000, 5, n_dup=3, tandem=2), sim.genome(20000, 5, n_dup=3, tandem=2)
    assert np.array_equal(g1, g2) and set(np.unique(g1)) <= set(b"ACGT")
    m1, t1 = sim.mutate(g1, 9); m2, _ = sim.mutate(g1, 9)
    assert np.array_equal(m1, m2) and len(t1) > 0
    a1, b1 = sim.simulate_pairs(m1, 500, 100, seed=3); a2, b2 = sim.simulate_pairs(m1, 500, 100, seed=3)
    assert np.array_equal(a1, a2) and np.array_equal(b1, b2) and a1.shape == (500, 100)
    seq, off = sim.interleave(a1, b1)
    assert len(off) == 1001 and np.array_equal(seq[off[1]:off[2]], b1[0])


def test_revcomp_and_encode():
    s = np.frombuffer(b"ACGTNacgtR", dtype=np.uint8)
    assert sim.revcomp(s).tobytes() == b"NACGTNACGT"
    assert sim.encode(s).tolist() == [0, 1, 2, 3, 4, 0, 1, 2, 3, 4]


def test_fastq_writer_shape(tmp_path):
    r = np.frombuffer(b"ACGTACGTAC" * 3, dtype=np.uint8).reshape(3, 10)
    p = str(tmp_path / "x.fq")
    sim.write_fastq(p, r, 1)
    lines = open(p).read().split("\n")
    assert lines[0] == "@r000000000/1" and lines[1] == "ACGTACGTAC" and lines[2] == "+" and lines[3] == "I" * 10 and lines[-1] == ""


def test_vectorised_mutant_and_read_indels():
    """The generators of the GRCh38-sized bench configs (no per-event Python loop): seeded, ACGT only, lengths and reads as specified."""
    g = sim.genome(400000, 6, n_dup=5, repeat_frac=0.15)
    m1, m2 = sim.mutate_fast(g, 3, snp_per_mb=1000, small_indel_per_mb=2000, large_indel_per_mb=500), sim.mutate_fast(g, 3, 1000, 2000, 500)
    assert np.array_equal(m1, m2) and set(np.unique(m1)) <= set(b"ACGT")
    assert m1.shape != g.shape and abs(len(m1) - len(g)) < 20000            # ~1000 indel events of 1..30 bases either way
    assert (m1[:90] == g[:90]).mean() > 0.9                                 # events start at position 100
    plain = sim.simulate_pairs_fast(m1, 2000, 250, seed=5, frag_mean=600, frag_sd=50, sub_rate=0.0, block=500)
    again = sim.simulate_pairs_fast(m1, 2000, 250, seed=5, frag_mean=600, frag_sd=50, sub_rate=0.0, block=500)
    assert np.array_equal(plain[0], again[0]) and plain[0].shape == (2000, 250)
    # any slice of a library can be regenerated: block b depends on (seed, b) only
    tail = sim.simulate_pairs_fast(m1, 500, 250, seed=5, frag_mean=600, frag_sd=50, sub_rate=0.0, block=500, first_block=3)
    assert np.array_equal(tail[0], plain[0][1500:]) and np.array_equal(tail[1], plain[1][1500:])
    hit = sim.simulate_pairs_fast(m1, 2000, 250, seed=5, frag_mean=600, frag_sd=50, sub_rate=0.0, block=500, indel_read_frac=0.4)
    assert hit[0].shape == (2000, 250) and set(np.unique(hit[0])) <= set(b"ACGT")
    # a read with a sequencing indel still starts like its fragment (the event lies at offset >= 10)
    win = np.lib.stride_tricks.sliding_window_view(m1, 10)
    starts = {w.tobytes() for w in win[::1]}
    assert sum(r[:10].tobytes() in starts or sim.revcomp(r)[-10:].tobytes() in starts for r in hit[0][:200]) >= 190


def test_bench_contig_table():
    import bench
    assert sum(bench.GRCH38) == 3_088_269_832 and len(bench.GRCH38) == 24
    bench.CFG = bench.CONFIGS[3]
    lens, names = bench.contig_lengths(1_000_000)
    assert sum(lens) == 1_000_000 and len(lens) == 24 and names[0] == "chr1" and names[-2:] == ["chrX", "chrY"]
    bench.CFG = bench.CONFIGS[2]
    assert bench.contig_lengths(1000) == (None, None)
