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
