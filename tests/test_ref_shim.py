"""The reference-side shim (oracle/ref_shim.cpp) drives the reference's operator functions in a loop written for the
tests; this checks that loop against the reference's own Mapping() (the real thread body) on FASTQ files."""
import os
import subprocess
import sys
import textwrap

import pytest

import parity_util as pu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not pu.have_ref(), reason="oracle/_ref not built")
def test_driver_loop_equals_reference_thread_body(built, tmp_path):
    code = textwrap.dedent("""
        import sys, numpy as np
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        import parity_util as pu, ref_oracle as ro
        from mapcaller_b200 import simulate as sim
        case = pu.make_case(seed=31, n_pairs=4000, genome_len=90000, contigs=2, sv=4.0)
        ix = pu.build_index(case); ix.save(%r)
        sim.write_fastq(%r, case['r1'], 1); sim.write_fastq(%r, case['r2'], 2)
        ro.load(%r)
        reads, est = ro.map_reads(case['seq'], case['off'], True); ro.lib().mcref_finish_sites()
        a = (ro.counters(), ro.profile().tobytes(), ro.indels(0), ro.indels(1), ro.breakpoints(), ro.sites(0), ro.sites(1))
        ro.lib().mcref_reset_state()
        ro.lib().mcref_run_mapping(%r.encode(), %r.encode(), b'', 1, 1)
        b = (ro.counters(), ro.profile().tobytes(), ro.indels(0), ro.indels(1), ro.breakpoints(), ro.sites(0), ro.sites(1))
        for k in ('reads', 'mapped', 'paired', 'dist_sum', 'len_sum'): assert a[0][k] == b[0][k], k
        assert a[1:] == b[1:]
        print('OK')
    """) % (ROOT, os.path.join(ROOT, "tests"), str(tmp_path / "idx"), str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq"), str(tmp_path / "idx"),
            str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(tmp_path))
    assert out.returncode == 0 and "OK" in out.stdout, out.stderr[-2000:]
