"""The drop-in binary with MC_B200_DEVICE_SAM=1: SAM lines printed from the device's records (mc_sam_records) instead of the
reference's SamReport.o over downloaded candidates - SAM and VCF must still equal the reference CLI's.  (Last in the
collection order on purpose: when it was written only the pe_nw mode could still be run on a B200, the others on the host
harness; the C-ABI entry itself is covered by tests/test_sam_records_gpu.py.)"""
import os
import subprocess

import pytest

import parity_util as pu
from mapcaller_b200 import simulate as sim
from test_dropin_gpu import GPU_BIN, REF_BIN, sam_records, vcf_body

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not (os.path.exists(REF_BIN) and os.path.exists(GPU_BIN)), reason="reference / drop-in binaries not on this box")]


@pytest.mark.parametrize("mode", ["pe_nw", "pe_ksw2", "se_nw"])
def test_device_sam_dropin_equals_the_reference_cli(tmp_path, mode):
    # the case of tests/test_dropin_gpu.py: known to stay clear of the reference's out-of-bounds rescue windows on the GPU box
    case = pu.make_case(seed=41, n_pairs=6000, genome_len=120000, contigs=2, sv=3.0)
    fa = str(tmp_path / "ref.fa")
    sim.write_fasta(fa, case["contigs"])
    f1, f2 = str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq")
    sim.write_fastq(f1, case["r1"], 1); sim.write_fastq(f2, case["r2"], 2)
    idx = str(tmp_path / "idx")
    subprocess.check_call([REF_BIN, "index", fa, idx], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    extra = ["-alg", "ksw2"] if "ksw2" in mode else []
    reads = ["-f", f1] + ([] if mode.startswith("se") else ["-f2", f2])
    outs = {}
    for tag, exe in (("ref", REF_BIN), ("gpu", GPU_BIN)):
        sam, vcf = str(tmp_path / (tag + ".sam")), str(tmp_path / (tag + ".vcf"))
        subprocess.check_call([exe, "-i", idx, "-t", "1"] + reads + ["-sam", sam, "-vcf", vcf, "-log", str(tmp_path / (tag + ".log"))] + extra,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=str(tmp_path), env=dict(os.environ, MC_B200_DEVICE_SAM="1"))
        outs[tag] = (sam_records(sam, mode.startswith("se")), vcf_body(vcf))
    assert outs["gpu"][0] == outs["ref"][0], "SAM differs"
    assert outs["gpu"][1] == outs["ref"][1], "VCF differs"
    assert len(outs["ref"][0]) > 5000
