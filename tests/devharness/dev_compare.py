"""Developer harness: run the stage bodies compiled for the host (tools/hostemu/build.sh) against the
reference (oracle/_ref) on a box without a GPU.  Not a test, not a fallback, never imported elsewhere."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from mapcaller_b200 import api, simulate as sim
api._LIB_PATH = os.path.join(ROOT, "tools", "hostemu", "_build", "libmc_hostemu.so")
import ref_oracle as ro


def diff_reads(mine, ref, limit=5):
    bad = 0
    for i, (a, b) in enumerate(zip(mine, ref)):
        if a != b:
            bad += 1
            if bad <= limit:
                print("read", i, "differs")
                for k in ("rlen", "score", "sub_score", "best"):
                    if a[k] != b[k]: print("   ", k, a[k], b[k])
                if len(a["cands"]) != len(b["cands"]): print("    ncand", len(a["cands"]), len(b["cands"]))
                for ci, (x, y) in enumerate(zip(a["cands"], b["cands"])):
                    if x != y:
                        print("    cand", ci, {k: x[k] for k in x if k != "frags"}, {k: y[k] for k in y if k != "frags"})
                        for fx, fy in zip(x["frags"], y["frags"]):
                            if fx != fy: print("       mine", fx, "\n       ref ", fy)
                        if len(x["frags"]) != len(y["frags"]): print("       nfrag", len(x["frags"]), len(y["frags"]))
    return bad


def run(fa, idx_prefix, r1, r2, paired=True, ksw2=False, batch=None):
    G = ro.load(idx_prefix)
    ro.set_params(nw=not ksw2)
    seq, off = sim.interleave(r1, r2) if paired else (r1.reshape(-1), np.arange(len(r1) + 1, dtype=np.int64) * r1.shape[1])
    t = time.time(); ref_reads, ref_est = ro.map_reads(seq, off, paired); print("reference %.2fs" % (time.time() - t))
    ro.lib().mcref_finish_sites()
    ix = api.Index.load(idx_prefix)
    ctx = api.Context(ix, paired=int(paired), alg_ksw2=int(ksw2), want_alignments=1)
    n = len(off) - 1
    batch = batch or n
    mine = []; est = []
    t = time.time()
    for b in range(0, n, batch):
        e = min(n, b + batch)
        res = ctx.map_batch(seq[off[b]:off[e]], off[b:e + 1] - off[b])
        mine += api.unpack_reads(res); est += [int(x) for x in res["chunks"]["est_distance"]]
        print("  batch", b, "replays", res["replays"])
    print("mine %.2fs" % (time.time() - t))
    print("reads differing:", diff_reads(mine, ref_reads), "of", n)
    if paired: print("est equal:", est == ref_est)
    rc = ro.counters(); mt = ctx.totals(); print(rc, mt)
    p_ref = ro.profile(); p_mine = ctx.profile_columns()
    d = np.nonzero((p_ref != p_mine).any(axis=1))[0]
    print("profile columns differing:", len(d), d[:10])
    for k in d[:5]: print("   ", k, p_mine[k], p_ref[k])
    ins, dele = ctx.indels()
    print("ins equal", ins == ro.indels(0), len(ins), "del equal", dele == ro.indels(1), len(dele))
    print("bp equal", ctx.breakpoints() == ro.breakpoints(), len(ro.breakpoints()))
    inv = sorted(ctx.sites(0), key=lambda x: x[0]); tnl = sorted(ctx.sites(1), key=lambda x: x[0])
    print("inv equal", sorted(inv) == sorted(ro.sites(0)), len(inv), "tnl equal", sorted(tnl) == sorted(ro.sites(1)), len(tnl))
    print(ctx.stats())


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "toy"
    ksw2 = "ksw2" in sys.argv
    os.makedirs("/tmp/toy", exist_ok=True)
    if which == "toy":
        mut = sim.read_fasta("/root/reference/test/mut.fa")
        r1, r2 = sim.simulate_pairs(mut[0][1], 10500, 100, seed=1)
        if not os.path.exists("/tmp/toy/idx.bwt"): ro.build_index("/root/reference/test/ref.fa", "/tmp/toy/idx")
        run("/root/reference/test/ref.fa", "/tmp/toy/idx", r1, r2, ksw2=ksw2, batch=int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else None)
