"""Subprocess worker: runs one case through oracle/_ref (the unmodified reference) and pickles what it
leaves behind.  A separate process per case because the reference keeps its index and profile in
process-wide globals.  Test infrastructure only."""
import os
import pickle
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_oracle as ro  # noqa: E402


def main(job_path: str, out_path: str) -> None:
    job = np.load(job_path, allow_pickle=True)
    prm = job["params"].item()
    ro.load(str(job["prefix"]))
    ro.set_params(max_pos_diff=prm.get("max_pos_diff", 30), max_clip=prm.get("max_clip", 5), max_dup=prm.get("max_dup", 5),
                  maxmm=prm.get("max_mismatch_rate", 0.05), nw=not prm.get("alg_ksw2", 0), unique=True, threads=1)
    seq, off = job["seq"], job["off"]
    paired = bool(prm.get("paired", 1))
    reads, est = ro.map_reads(seq, off, paired, True)
    ro.lib().mcref_finish_sites()
    out = dict(est=est, counters=ro.counters(), profile=ro.profile(), ins=ro.indels(0), dele=ro.indels(1),
               bp=ro.breakpoints(), inv=ro.sites(0), tnl=ro.sites(1))
    if "vc" in job.files:   # variant-calling scans of the profile just built, one per parameter set
        out["vc"] = [ro.variant_scan(**kw) for kw in job["vc"].item()["sets"]]
    if bool(job["want_reads"]):
        out["reads"] = reads
    with open(out_path, "wb") as fh:
        pickle.dump(out, fh, protocol=4)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
