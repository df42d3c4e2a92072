"""End-to-end witness of the drop-in: the reference's own CLI (oracle/_ref/MapCaller -t 1) against the same driver with the
hot path swapped for the GPU library (mapcaller_b200/dropin/_build/MapCaller_b200).  SAM must be byte-identical, the VCF
identical apart from the header lines that embed argv."""
import os
import subprocess

import pytest

import parity_util as pu
from mapcaller_b200 import simulate as sim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "MapCaller")
GPU_BIN = os.path.join(ROOT, "mapcaller_b200", "dropin", "_build", "MapCaller_b200")

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not (os.path.exists(REF_BIN) and os.path.exists(GPU_BIN)), reason="reference / drop-in binaries not on this box")]


def sam_records(path, drop_qual):
    """SAM lines as bytes.  In single-end mode the reference's (unchanged) SamReport.cpp leaves the first byte of a reversed
    quality string uninitialised (GetReverseQualityStr, src/SamReport.cpp:318-322): whatever the heap held ends up in the
    record (it may even be a NUL or a newline and cut the line), so only the columns before QUAL are comparable there."""
    lines = open(path, "rb").read().split(b"\n")
    if not drop_qual:
        return lines
    out = []
    for l in lines:
        f = l.split(b"\t")
        if l.startswith(b"@"):
            out.append(l)
        elif len(f) >= 10 and f[0].startswith(b"r"):
            out.append(b"\t".join(f[:10]))
    return out


def vcf_body(path):
    return [l for l in open(path) if not (l.startswith("##command_line") or l.startswith("##reference") or l.startswith("##fileDate"))]


@pytest.mark.parametrize("mode", ["pe_nw", "pe_ksw2_monomorphic", "se_nw", "pe_nw_vcfonly", "se_nw_vcfonly"])
def test_same_sam_and_vcf_as_the_reference_cli(tmp_path, mode):
    case = pu.make_case(seed=41, n_pairs=6000, genome_len=120000, contigs=2, sv=3.0)
    fa = str(tmp_path / "ref.fa")
    sim.write_fasta(fa, case["contigs"])
    f1, f2 = str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq")
    sim.write_fastq(f1, case["r1"], 1); sim.write_fastq(f2, case["r2"], 2)
    idx = str(tmp_path / "idx")
    subprocess.check_call([REF_BIN, "index", fa, idx], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    extra = []
    if "ksw2" in mode:
        extra += ["-alg", "ksw2"]
    if "monomorphic" in mode:
        extra += ["-monomorphic"]
    reads = ["-f", f1] + ([] if mode.startswith("se") else ["-f2", f2])
    outs = {}
    for tag, exe in (("ref", REF_BIN), ("gpu", GPU_BIN)):
        sam, vcf = str(tmp_path / (tag + ".sam")), str(tmp_path / (tag + ".vcf"))
        # without -sam the drop-in hands raw FASTQ blocks to the device parser (mc_ingest_fastq) instead of the reference's reader
        want_sam = "vcfonly" not in mode
        subprocess.check_call([exe, "-i", idx, "-t", "1"] + reads + (["-sam", sam] if want_sam else []) + ["-vcf", vcf, "-log", str(tmp_path / (tag + ".log"))] + extra,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=str(tmp_path))
        outs[tag] = (sam_records(sam, mode.startswith("se")) if want_sam else [], vcf_body(vcf))
    assert outs["gpu"][0] == outs["ref"][0], "SAM differs"
    assert outs["gpu"][1] == outs["ref"][1], "VCF differs"
    assert len(outs["ref"][1]) > 20
