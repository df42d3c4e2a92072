"""mc_index_build_gpu: BWT rows and suffix samples sorted on the GPU in chunks of rows (csrc/index_gpu.cu) must give the very
files the host builder writes - which tests/test_index.py pins byte for byte against the reference's own builder."""
import os
import time

import numpy as np
import pytest

from mapcaller_b200 import api, simulate as sim

pytestmark = pytest.mark.gpu


def _files(ix, prefix):
    ix.save(prefix)
    return {ext: open(prefix + ext, "rb").read() for ext in (".bwt", ".sa", ".pac", ".ann", ".amb")}


@pytest.mark.parametrize("kind", ["repeats_two_contigs", "low_complexity", "tiny", "five_megabases", "five_megabases_in_chunks", "repeats_in_chunks"])
def test_gpu_suffix_sort_gives_the_same_index_files(tmp_path, kind, monkeypatch):
    if kind.endswith("_in_chunks"):      # rows produced in many chunks, as for a text that does not fit one sort
        monkeypatch.setenv("MC_INDEX_CHUNK", "700000" if kind.startswith("five") else "20000")
        kind = "five_megabases" if kind.startswith("five") else "repeats_two_contigs"
    if kind == "repeats_two_contigs":
        g = np.concatenate([sim.genome(90000, 5, n_dup=12, dup_len=(300, 3000), tandem=6), sim.genome(70000, 6, n_dup=8, dup_len=(300, 900), tandem=3)])
        lens, names = [90000, 70000], ["ctgA", "ctgB"]
    elif kind == "low_complexity":
        rng = np.random.default_rng(3)
        g = np.concatenate([np.full(5000, ord("A"), dtype=np.uint8), np.frombuffer(b"ACG" * 4000, dtype=np.uint8), rng.choice(np.frombuffer(b"AT", dtype=np.uint8), 9000)])
        lens, names = [len(g)], ["lc"]
    elif kind == "tiny":
        g = np.frombuffer(b"GATTACAGATTACA", dtype=np.uint8)
        lens, names = [len(g)], ["t"]
    else:
        g = sim.genome(5_000_000, 8, n_dup=200, dup_len=(1000, 3000))
        lens, names = [len(g)], ["chr"]
    codes = sim.encode(g)
    t = time.time(); cpu = api.Index.build(codes, chrom_len=lens, chrom_name=names); t_cpu = time.time() - t
    t = time.time(); gpu = api.Index.build(codes, chrom_len=lens, chrom_name=names, gpu_device=0); t_gpu = time.time() - t
    a, b = _files(cpu, str(tmp_path / "cpu")), _files(gpu, str(tmp_path / "gpu"))
    for ext in a:
        assert a[ext] == b[ext], ext
    print("%s: %d bases, host %.2f s, GPU %.2f s" % (kind, len(g), t_cpu, t_gpu))
