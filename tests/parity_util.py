"""Shared helpers of the parity tests: seeded cases, the CUDA path through the C ABI, the checkers."""
from __future__ import annotations

import os
import pickle
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from mapcaller_b200 import simulate as sim  # noqa: E402


def have_ref() -> bool:
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libmcref.so"))


def make_case(kind: str = "small", seed: int = 1, n_pairs: int = 3000, read_len: int = 100, genome_len: int = 60000,
              contigs: int = 1, sub_rate: float = 0.005, indel_rate: float = 0.0, n_rate: float = 0.0, sv: float = 1.0,
              n_dup: int = 6, tandem: int = 3, frag_mean: float = 400, frag_sd: float = 40, lower_rate: float = 0.0,
              repeat_frac: float = 0.0, repeat_div: float = 0.02, **params):
    """A genome (possibly several contigs), a mutated copy the reads come from, and the reads."""
    parts, names = [], []
    for c in range(contigs):
        # repeat_frac > 0: interspersed families of slightly diverged copies (reads with 0 < score - sub_score <= 5: the MAPQ formula)
        parts.append(sim.genome(genome_len // contigs, seed * 100 + c, n_dup=n_dup, dup_len=(300, 900), tandem=tandem, repeat_frac=repeat_frac,
                                families=((300, repeat_div), (1000, repeat_div))))
        names.append("ctg%d" % (c + 1))
    ref = np.concatenate(parts)
    mut_parts = [sim.mutate(p, seed * 1000 + i, snp_per_mb=3000, small_indel_per_mb=400, large_indel_per_mb=150, sv_per_mb=sv * 30, sv_len=(400, 900))[0]
                 for i, p in enumerate(parts)]
    mut = np.concatenate(mut_parts)
    # Fragments stay clear of the first 3 kbp: for a mate anchored there the reference's rescue window starts before
    # RefSequence[0] (src/AlignmentRescue.cpp:87,93) and the unmodified reference reads out of bounds (it segfaults on
    # some hosts), so such reads cannot be pinned against it.
    r1, r2 = sim.simulate_pairs(mut[3000:], n_pairs, read_len, seed=seed + 7, frag_mean=frag_mean, frag_sd=frag_sd, sub_rate=sub_rate, indel_rate=indel_rate, n_rate=n_rate)
    if lower_rate > 0:   # soft-masked style input: the reference seeds through lower case but only counts upper-case bases in the profile
        rng = np.random.default_rng(seed + 99)
        for r in (r1, r2):
            rows = rng.random(len(r)) < lower_rate
            cols = rng.random(r.shape) < 0.3
            m = rows[:, None] & cols & (r != ord("N"))
            r[m] |= 0x20
    p = dict(paired=1, alg_ksw2=0, max_pos_diff=30, max_clip=5, max_dup=5, max_mismatch_rate=0.05)
    p.update(params)
    if p["paired"]:
        seq, off = sim.interleave(r1, r2)
    else:
        seq, off = r1.reshape(-1).copy(), np.arange(len(r1) + 1, dtype=np.int64) * r1.shape[1]
    return dict(contigs=list(zip(names, parts)), ref=ref, seq=seq, off=off, params=p, r1=r1, r2=r2)


def sparse_case(genome_len: int = 400000, covered: int = 60000, n_pairs: int = 6000, seed: int = 21, contigs: int = 1):
    """Reads from the first `covered` and the last 40 k bases only: the gap run in between is longer than 65535 columns
    (Variant_t::DP wraps, src/structure.h:187) and spans thousands of scan blocks."""
    case = make_case(seed=seed, n_pairs=10, genome_len=genome_len, contigs=contigs, n_dup=2, tandem=1)
    ref = case["ref"]
    kw = dict(snp_per_mb=3000, small_indel_per_mb=400, large_indel_per_mb=150, sv_per_mb=0, sv_len=(400, 900))
    head = sim.mutate(ref[:covered], seed * 31, **kw)[0]
    tail = sim.mutate(ref[-40000:], seed * 37, **kw)[0]
    r1, r2 = sim.simulate_pairs(head[3000:], n_pairs, 100, seed=seed + 7)
    q1, q2 = sim.simulate_pairs(tail[:-3000], n_pairs // 2, 100, seed=seed + 9)
    case["seq"], case["off"] = sim.interleave(np.concatenate([r1, q1]), np.concatenate([r2, q2]))
    return case


# parameter sets of the variant-calling scan used by the tests (defaults, gVCF, monomorphic, somatic, haploid / low thresholds)
VC_SETS = [dict(), dict(gvcf=1), dict(monomorphic=1), dict(somatic=1), dict(ploidy=1, min_allele_depth=2, frequency_thr=0.1),
           dict(min_cnv_size=5, min_unmapped_size=10, gvcf=1)]


def build_index(case, threads: int = 0):
    from mapcaller_b200 import api
    codes = sim.encode(case["ref"])
    assert codes.max() <= 3
    return api.Index.build(codes, [len(s) for _, s in case["contigs"]], [n for n, _ in case["contigs"]], threads)


def ref_results(case, index, want_reads: bool = True, vc=None):
    """Runs the case through oracle/_ref in a subprocess; `vc` = list of variant-scan parameter sets to run afterwards."""
    with tempfile.TemporaryDirectory() as td:
        prefix = os.path.join(td, "idx")
        index.save(prefix)
        job = os.path.join(td, "job.npz")
        extra = dict(vc=np.array({"sets": list(vc)}, dtype=object)) if vc else {}
        np.savez(job, prefix=prefix, seq=case["seq"], off=case["off"], params=np.array(case["params"], dtype=object), want_reads=want_reads, **extra)
        outp = os.path.join(td, "out.pkl")
        subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_worker.py"), job, outp], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        with open(outp, "rb") as fh:
            return pickle.load(fh)


def oracle_results(case, index, want_reads: bool = True):
    """Runs the case through the CPU restatement (oracle/libmcoracle.so)."""
    import cpu_oracle
    with tempfile.TemporaryDirectory() as td:
        prefix = os.path.join(td, "idx")
        index.save(prefix)
        orc = cpu_oracle.Oracle(prefix, **case["params"])
        paired = bool(case["params"]["paired"])
        reads, est = orc.map_reads(case["seq"], case["off"], paired, True)
        out = dict(est=est, counters=orc.counters(), profile=orc.profile(), ins=orc.indels(0), dele=orc.indels(1), bp=orc.breakpoints(),
                   inv=orc.sites(0), tnl=orc.sites(1), work=orc.work())
        if want_reads:
            out["reads"] = reads
        orc.close()
        return out


def cuda_results(case, index, batch_reads: int | None = None, want_reads: bool = True, device: int = 0, vc=None, **extra):
    from mapcaller_b200 import api
    seq, off = case["seq"], case["off"]
    n = len(off) - 1
    batch_reads = batch_reads or n
    out = dict(reads=[], est=[], replays=0)
    with api.Context(index, want_alignments=int(want_reads), update_profile=1, device=device, **case["params"], **extra) as ctx:
        for b in range(0, n, batch_reads):
            e = min(n, b + batch_reads)
            res = ctx.map_batch(seq[off[b]:off[e]], off[b:e + 1] - off[b])
            if want_reads:
                out["reads"] += api.unpack_reads(res)
            out["est"] += [int(x) for x in res["chunks"]["est_distance"]] if case["params"]["paired"] else []
            out["replays"] += res["replays"]
        t = ctx.totals()
        out["counters"] = dict(reads=t["total_reads"], mapped=t["total_mapped"], paired=t["total_paired"], dist_sum=t["total_distance"],
                               len_sum=t["read_length_sum"], avgDist=t["avg_dist"])
        out["profile"] = ctx.profile_columns()
        out["ins"], out["dele"] = ctx.indels()
        out["bp"] = ctx.breakpoints()
        out["inv"] = sorted(ctx.sites(0), key=lambda x: x[0])
        out["tnl"] = sorted(ctx.sites(1), key=lambda x: x[0])
        out["stats"] = ctx.stats()
        out["summary"] = ctx.profile_summary()
        if vc:
            out["vc"] = [ctx.variant_scan(**kw) for kw in vc]
    return out


def assert_same_variants(mine, ref) -> None:
    """mc_variant_scan against IdentifyVariants of the unmodified reference, one entry per parameter set."""
    import ref_oracle as ro
    assert len(mine["vc"]) == len(ref["vc"])
    for i, ((mv, md), (rv, rd)) in enumerate(zip(mine["vc"], ref["vc"])):
        assert np.array_equal(md, rd), "BlockDepthArr differs (set %d)" % i
        assert ro.variants_equal(mv, rv), "variant records differ (set %d): %d vs %d" % (i, len(mv), len(rv))


def _strip_paired(r):
    return dict(r, cands=[dict(c, paired=-1) for c in r["cands"]])


def first_read_diff(mine, ref, paired: bool = True):
    """Single-end runs never initialise PairedAlnCanIdx in the reference (src/ReadMapping.cpp:575-585) and nothing reads it."""
    for i, (a, b) in enumerate(zip(mine, ref)):
        if not paired:
            a, b = _strip_paired(a), _strip_paired(b)
        if a != b:
            return i, a, b
    return None


def assert_same(mine, ref, want_reads: bool = True, paired: bool = True) -> None:
    """Bit-exact comparison of everything the hot path leaves behind."""
    if want_reads and "reads" in ref:
        assert len(mine["reads"]) == len(ref["reads"])
        d = first_read_diff(mine["reads"], ref["reads"], paired=bool(ref.get("est")) or paired)
        assert d is None, "read %d differs:\n mine %r\n ref  %r" % d
    if ref.get("est"):
        assert mine["est"] == ref["est"], "EstiDistance trajectory differs"
    for k in ("reads", "mapped", "paired", "dist_sum", "len_sum", "avgDist"):
        assert mine["counters"][k] == ref["counters"][k], "counter %s: %r != %r" % (k, mine["counters"][k], ref["counters"][k])
    bad = np.nonzero((mine["profile"] != ref["profile"]).any(axis=1))[0]
    assert len(bad) == 0, "profile differs at %d columns, first %d: %r vs %r" % (len(bad), bad[0], mine["profile"][bad[0]], ref["profile"][bad[0]])
    assert mine["ins"] == ref["ins"], "InsertSeqMap differs"
    assert mine["dele"] == ref["dele"], "DeleteSeqMap differs"
    assert mine["bp"] == ref["bp"], "BreakPointMap differs"
    assert sorted(mine["inv"]) == sorted(ref["inv"]), "InversionSiteVec differs"
    assert sorted(mine["tnl"]) == sorted(ref["tnl"]), "TranslocationSiteVec differs"
    if "summary" in mine:   # CheckMappingCoverage / ReportDuplicationRate over the checker's profile
        p = ref["profile"].astype(np.int64); cov = p[:, :4].sum(axis=1)
        want = dict(aligned_bases=int((cov > 0).sum()), coverage_sum=int(cov.sum()), dup_sites=int((p[:, 5] > 0).sum()), dup_reads=int(p[:, 5].sum()))
        assert mine["summary"] == want, "coverage / duplication summary differs: %r vs %r" % (mine["summary"], want)


def _write_inputs(case, td):
    fa = os.path.join(td, "ref.fa")
    sim.write_fasta(fa, case["contigs"])
    paired = bool(case["params"]["paired"])
    f1, f2 = os.path.join(td, "r1.fq"), os.path.join(td, "r2.fq")
    sim.write_fastq(f1, case["r1"], 1)
    if paired:
        sim.write_fastq(f2, case["r2"], 2)
    return fa, f1, (f2 if paired else None)


def sam_lines_reference(case, td, all_best: bool = False):
    """SAM records (bytes, in file order, header lines dropped) of the unmodified reference CLI, -t 1 (all_best: with -m)."""
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "MapCaller")
    fa, f1, f2 = _write_inputs(case, td)
    idx = os.path.join(td, "idx")
    subprocess.check_call([ref_bin, "index", fa, idx], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    sam = os.path.join(td, "ref.sam")
    cmd = [ref_bin, "-i", idx, "-t", "1", "-f", f1] + (["-f2", f2] if f2 else []) + ["-sam", sam, "-no_vcf", "-log", os.path.join(td, "log")]
    prm = case["params"]
    if prm.get("alg_ksw2"):
        cmd += ["-alg", "ksw2"]
    if all_best:
        cmd += ["-m"]
    assert prm.get("max_pos_diff", 30) == 30, "MaxPosDiff has no command-line switch"
    cmd += ["-dup", str(prm.get("max_dup", 5)), "-maxclip", str(prm.get("max_clip", 5)), "-maxmm", repr(float(prm.get("max_mismatch_rate", 0.05)))]
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=td)
    return [l for l in open(sam, "rb").read().split(b"\n") if l and not l.startswith(b"@")]


def sam_lines_cuda(case, device: int = 0, batch_reads: int | None = None):
    """The same records printed from mc_sam_records (flag, position, MAPQ, CIGAR, mate fields and tags from the device;
    names, bases and qualities from the FASTQ text)."""
    from mapcaller_b200 import api
    paired = bool(case["params"]["paired"])
    texts = [sim.fastq_text(case["r1"], 1, "r").tobytes().split(b"\n")] + ([sim.fastq_text(case["r2"], 2, "r").tobytes().split(b"\n")] if paired else [])
    seq, off = case["seq"], case["off"]
    n = len(off) - 1
    batch_reads = batch_reads or n
    names = [nm for nm, _ in case["contigs"]]
    lines = []
    with api.Context(build_index(case), want_alignments=0, update_profile=0, device=device, **case["params"]) as ctx:
        for b in range(0, n, batch_reads):
            e = min(n, b + batch_reads)
            ctx.map_batch(seq[off[b]:off[e]], off[b:e + 1] - off[b])
            recs, cigars = ctx.sam_records()
            for k in range(e - b):
                r = b + k
                t, i = (texts[r & 1], r >> 1) if paired else (texts[0], r)
                head = t[4 * i][1:].split()[0]
                if head.endswith(b"/1") or head.endswith(b"/2"):   # IdentifyHeaderEndPos, src/GetData.cpp:15-30
                    head = head[:-2]
                l = api.format_sam_line(recs[k], cigars[k], head, t[4 * i + 1], t[4 * i + 3], [x.encode() for x in names])
                if l is not None:
                    lines.append(l)
    return lines


def sam_text_cuda(case, device: int = 0, batch_pairs: int | None = None, all_best: bool = False):
    """The SAM lines assembled on the device (mc_sam_text) from FASTQ text ingested on the device, batch by batch."""
    from mapcaller_b200 import api
    paired = bool(case["params"]["paired"])
    r1, r2 = case["r1"], case["r2"]
    n = len(r1)
    batch_pairs = batch_pairs or n
    lines = []
    with api.Context(build_index(case), want_alignments=0, update_profile=0, device=device, **case["params"]) as ctx:
        for b in range(0, n, batch_pairs):
            e = min(n, b + batch_pairs)
            t1 = sim.fastq_text(r1[b:e], 1, "r", first=b)
            t2 = sim.fastq_text(r2[b:e], 2, "r", first=b) if paired else None
            ctx.ingest_fastq(t1, t2, slot=1, final=(e == n))
            ctx.map_staged(1)
            lines += [l for l in ctx.sam_text(1, all_best).split(b"\n") if l]
    return lines


def sam_comparable(lines, paired: bool):
    """In single-end mode the reference's SamReport.cpp leaves the first byte of a reversed quality string uninitialised
    (GetReverseQualityStr, src/SamReport.cpp:318-322; it may even be a NUL or a newline and cut the line), so only the
    columns before QUAL are comparable there."""
    if paired:
        return lines
    out = []
    for l in lines:
        f = l.split(b"\t")
        if len(f) >= 10 and f[0].startswith(b"r"):
            out.append(b"\t".join(f[:10]))
    return out


def smoke_case() -> None:
    """One small invocation of the hot path on cuda:0, checked against the oracle."""
    case = make_case(seed=3, n_pairs=600, genome_len=40000)
    ix = build_index(case)
    mine = cuda_results(case, ix)
    assert_same(mine, oracle_results(case, ix))
    assert mine["stats"]["kernel_launches"] > 0
    # the variant-calling scan of the same profile against the committed output of the unmodified reference (the smoke case
    # is golden case pe_nw: same seed and sizes)
    import golden_util as gu
    gcase, _ = gu.load("pe_nw")
    sets, gold = gu.load_vc("pe_nw")
    assert_same_variants(cuda_results(gcase, build_index(gcase), want_reads=False, vc=sets), dict(vc=gold))
