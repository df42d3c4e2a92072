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


# pe_nw / pe_ksw2_monomorphic / se_nw / pe_nw_m: plain FASTQ, SAM text assembled on the device (mc_sam_text; _m = the reference's -m,
# one line per best candidate); *_hostsam: MC_B200_HOST_SAM=1, the reference's reader + SamReport.o over the downloaded candidates;
# *_gz: gzip'ed FASTQ (the reference's gz reader feeds mc_map_batch); *_vcfonly: no SAM.  In every mode the VCF comes from the
# unchanged VariantCalling() with IdentifyVariants answered by mc_variant_scan.
# *_fasta: FASTA reads, one line of bases per record - parsed on the device like FASTQ (QUAL is "*"); *_fasta_wrapped: bases on
# several lines - the device refuses the first block and the reference's reader takes the library.
@pytest.mark.parametrize("mode", ["pe_nw", "pe_ksw2_monomorphic", "se_nw", "pe_nw_vcfonly", "se_nw_vcfonly", "pe_nw_m", "pe_nw_hostsam", "pe_nw_gz", "pe_nw_fasta", "pe_nw_fasta_wrapped"])
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
    if mode.endswith("_m"):
        extra += ["-m"]
    if mode.endswith("_gz"):
        import gzip
        for f in (f1, f2):
            with open(f, "rb") as src, gzip.open(f + ".gz", "wb", compresslevel=1) as dst:
                dst.write(src.read())
        f1, f2 = f1 + ".gz", f2 + ".gz"
    if "_fasta" in mode:
        width = 60 if mode.endswith("_wrapped") else 0
        for f, r, mate in ((f1, case["r1"], 1), (f2, case["r2"], 2)):
            with open(f[:-3] + ".fa", "wb") as fh:
                for i, row in enumerate(r):
                    b = row.tobytes()
                    fh.write(b">r%09d/%d\n" % (i, mate))
                    fh.write(b"\n".join(b[k:k + width] for k in range(0, len(b), width)) + b"\n" if width else b + b"\n")
        f1, f2 = f1[:-3] + ".fa", f2[:-3] + ".fa"
    env = dict(os.environ)
    if mode.endswith("_hostsam"):
        env["MC_B200_HOST_SAM"] = "1"
    reads = ["-f", f1] + ([] if mode.startswith("se") else ["-f2", f2])
    outs = {}
    for tag, exe in (("ref", REF_BIN), ("gpu", GPU_BIN)):
        sam, vcf = str(tmp_path / (tag + ".sam")), str(tmp_path / (tag + ".vcf"))
        # without -sam the drop-in hands raw FASTQ blocks to the device parser (mc_ingest_fastq) instead of the reference's reader
        want_sam = "vcfonly" not in mode
        subprocess.check_call([exe, "-i", idx, "-t", "1"] + reads + (["-sam", sam] if want_sam else []) + ["-vcf", vcf, "-log", str(tmp_path / (tag + ".log"))] + extra,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=str(tmp_path), env=env)
        outs[tag] = (sam_records(sam, mode.startswith("se")) if want_sam else [], vcf_body(vcf))
    assert outs["gpu"][0] == outs["ref"][0], "SAM differs"
    assert outs["gpu"][1] == outs["ref"][1], "VCF differs"
    assert len(outs["ref"][1]) > 20


def _run_both(tmp_path, idx, reads_args, env=None, extra=()):
    """VCF bodies of the reference CLI and of the drop-in for the same command line (no SAM: the drop-in then feeds raw FASTQ
    blocks to the device parser)."""
    out = {}
    for tag, exe in (("ref", REF_BIN), ("gpu", GPU_BIN)):
        vcf = str(tmp_path / (tag + ".vcf"))
        e = dict(os.environ); e.update(env or {})
        subprocess.check_call([exe, "-i", idx, "-t", "1"] + list(reads_args) + ["-vcf", vcf, "-log", str(tmp_path / (tag + ".log"))] + list(extra),
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=str(tmp_path), env=e)
        out[tag] = vcf_body(vcf)
    return out


def _case_files(tmp_path, seed=43, n_pairs=9000):
    case = pu.make_case(seed=seed, n_pairs=n_pairs, genome_len=120000, contigs=2, sv=3.0)
    fa = str(tmp_path / "ref.fa")
    sim.write_fasta(fa, case["contigs"])
    idx = str(tmp_path / "idx")
    subprocess.check_call([REF_BIN, "index", fa, idx], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return case, idx


@pytest.mark.parametrize("layout", ["two_files", "single_end", "interleaved", "unequal_mates", "two_files_gz"])
def test_fastq_block_reader_crosses_block_boundaries(tmp_path, layout):
    """The raw-FASTQ path of the drop-in reads the files in blocks (64 MiB; 40 kB here, so that a 2 MB library crosses ~50
    block boundaries): every read of a single-file library has to be mapped, mate files whose byte sizes differ (trimmed
    R2) have to stay in step, and the library ends with the shorter file."""
    case, idx = _case_files(tmp_path)
    f1, f2 = str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq")
    r1, r2 = case["r1"], case["r2"]
    if layout == "unequal_mates":
        r2 = r2[:, :70]                      # shorter records in the second file: the two files reach EOF in different blocks
    if layout == "interleaved":
        t1, t2 = sim.fastq_text(r1, 1).reshape(len(r1), -1), sim.fastq_text(r2, 2).reshape(len(r2), -1)
        import numpy as np
        np.concatenate([t1, t2], axis=1).reshape(-1).tofile(f1)
        reads = ["-f", f1, "-p"]
    else:
        sim.write_fastq(f1, r1, 1); sim.write_fastq(f2, r2, 2)
        if layout.endswith("_gz"):           # inflated block by block (gzread), parsed on the device like plain text
            import gzip
            for f in (f1, f2):
                with open(f, "rb") as src, gzip.open(f + ".gz", "wb", compresslevel=1) as dst:
                    dst.write(src.read())
            f1, f2 = f1 + ".gz", f2 + ".gz"
        reads = ["-f", f1] if layout == "single_end" else ["-f", f1, "-f2", f2]
    out = _run_both(tmp_path, idx, reads, env={"MC_B200_FASTQ_BLOCK": "40000"})
    assert out["gpu"] == out["ref"], "VCF differs"
    assert len(out["ref"]) > 20


@pytest.mark.parametrize("sam", [False, True])
def test_two_libraries_in_one_run(tmp_path, sam):
    """-f a.fq b.fq -f2 c.fq d.fq: the reference restarts its chunk grid per library and keeps accumulating the profile, the
    totals and avgDist (src/ReadMapping.cpp:705-748).  Library 1 is not a multiple of 200 reads."""
    case, idx = _case_files(tmp_path, seed=47, n_pairs=5150)
    names = []
    for k, (lo, hi) in enumerate(((0, 2150), (2150, 5150))):
        f1, f2 = str(tmp_path / ("lib%d_1.fq" % k)), str(tmp_path / ("lib%d_2.fq" % k))
        sim.write_fastq(f1, case["r1"][lo:hi], 1); sim.write_fastq(f2, case["r2"][lo:hi], 2)
        names.append((f1, f2))
    reads = ["-f", names[0][0], names[1][0], "-f2", names[0][1], names[1][1]]
    if not sam:
        out = _run_both(tmp_path, idx, reads)
        assert out["gpu"] == out["ref"], "VCF differs"
    else:
        res = {}
        for tag, exe in (("ref", REF_BIN), ("gpu", GPU_BIN)):
            s, v = str(tmp_path / (tag + ".sam")), str(tmp_path / (tag + ".vcf"))
            subprocess.check_call([exe, "-i", idx, "-t", "1"] + reads + ["-sam", s, "-vcf", v, "-log", str(tmp_path / (tag + ".log"))],
                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=str(tmp_path))
            res[tag] = (sam_records(s, False), vcf_body(v))
        assert res["gpu"][0] == res["ref"][0], "SAM differs"
        assert res["gpu"][1] == res["ref"][1], "VCF differs"
