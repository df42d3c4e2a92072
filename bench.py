#!/usr/bin/env python
"""bench.py -- read pairs mapped / s of the MapCaller hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pairs P]

Workload (config.workload): BASELINE.json configs[1] -- E. coli-sized 4.6 Mbp synthetic reference (uniform ACGT +
200 copied 1-3 kbp segments), reads simulated from a mutant carrying the reference simulator's variant rates
(SNP 3000/Mb, small indel 200/Mb, large indel 50/Mb, SV 1/Mb), 2x100 bp at 50x = 1.15 M pairs, 0.5 % substitution
errors.  One STEP = one complete pass of the hot path over that library: seeding, locate, clustering, pairing,
rescue, gapped fills (nw), pair statistics and the pile-up profile update, starting from a freshly reset context
(empty profile, avgDist = 1000) exactly like one run of the reference.

  value   device-timed throughput with the reads already resident in HBM (mc_stage_batch + mc_map_staged); CUDA
          events on the library's stream bracket every step, summed over the K steps, max over ranks
  e2e     the same metric through the public C ABI (mc_map_batch) with HOST buffers: pinned staging + H2D copy of
          the reads and D2H copy of the per-pair / per-chunk results inside the timed region (wall clock)
  roofline  the seed-search kernel (mc_seed_kernel): algorithmic bytes = 64 B x occ blocks the reference algorithm
          touches (counted by the kernel, identical to the oracle's count) / CUDA-event time of that kernel alone
  cpu_baseline  the unmodified reference (oracle/_ref, all host threads) or, if absent, the CPU restatement, timed
          on a bounded prefix of the same reads

With --gpus N > 1 (launched by torchrun, one rank per GPU) every rank maps its own library of the same size
(weak scaling: reads shard across GPUs with a full index replica each, no data-path collective); value is the
total over ranks / max-over-ranks time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GENOME_LEN = 4_600_000
READ_LEN = 100
COVERAGE = 50
FULL_PAIRS = GENOME_LEN * COVERAGE // (2 * READ_LEN)   # 1.15 M


def make_workload(n_pairs: int, seed_shift: int = 0):
    from mapcaller_b200 import simulate as sim
    g = sim.genome(GENOME_LEN, 7, n_dup=200)
    mut, _ = sim.mutate(g, 8)
    r1, r2 = sim.simulate_pairs(mut, n_pairs, READ_LEN, seed=11 + seed_shift, frag_mean=400, frag_sd=40, sub_rate=0.005)
    seq, off = sim.interleave(r1, r2)
    return g, r1, r2, seq, off


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device: int):
        self.rows, self.proc, self.device = [], None, device

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic():
    """dram bytes per launch of the seed kernel from the committed ncu --set full capture (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "seed_kernel_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


# --------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline
# --------------------------------------------------------------------------------------------------------------
def time_reference(g, r1, r2, n_sample: int, repeats: int = 1):
    """Pairs/s of the reference's own CPU implementation on the first n_sample pairs, all host threads.
    Returns (pairs_per_s, kind, cores, sample description, seconds per run)."""
    from mapcaller_b200 import api, simulate as sim
    cores = os.cpu_count() or 1
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libmcref.so")
    td = tempfile.mkdtemp(prefix="mcbench_")
    prefix = os.path.join(td, "idx")
    ix = api.Index.build(sim.encode(g))
    ix.save(prefix)
    if os.path.exists(ref_so):
        f1, f2 = os.path.join(td, "r1.fq"), os.path.join(td, "r2.fq")
        sim.write_fastq(f1, r1[:n_sample], 1); sim.write_fastq(f2, r2[:n_sample], 2)
        code = ("import sys, time, os; sys.path.insert(0, %r); import ref_oracle as ro\n"
                "ro.load(%r); ro.set_params(threads=%d)\n"
                "ts = []\n"
                "for i in range(%d):\n"
                "    ro.lib().mcref_reset_state(); t = time.perf_counter(); ro.lib().mcref_run_mapping(%r.encode(), %r.encode(), b'', %d, 1); ts.append(time.perf_counter() - t)\n"
                "print('SECONDS', min(ts), ro.counters()['reads'])\n") % (os.path.join(ROOT, "tests"), prefix, cores, repeats, f1, f2, cores)
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
        line = [l for l in out.stdout.splitlines() if l.startswith("SECONDS")]
        if not line:
            raise RuntimeError("reference run failed: " + out.stderr[-500:])
        sec = float(line[0].split()[1])
        kind, used = "reference", cores
        what = "Mapping() of the unmodified reference (oracle/_ref), -t %d, VCF profile on, FASTQ parse included" % cores
    else:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import cpu_oracle
        seq, off = sim.interleave(r1[:n_sample], r2[:n_sample])
        orc = cpu_oracle.Oracle(prefix)
        t = time.perf_counter(); orc.map_reads(seq, off, True, True); sec = time.perf_counter() - t
        orc.close()
        kind, used = "port", 1
        what = "CPU restatement oracle/libmcoracle.so, single thread"
    return n_sample / sec, kind, used, "first %d of %d pairs; %s" % (n_sample, len(r1), what), sec


# --------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=FULL_PAIRS, help="library size per GPU (default: the full 50x library)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="pairs for the cpu_baseline leg (default: sized for ~10-30 s)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "configs[1]: E. coli-sized 4.6 Mbp synthetic reference, 2x100 bp simulated PE reads at 50x (%d pairs per GPU), -alg nw, VCF profile on" % args.pairs,
              "genome_bp": GENOME_LEN, "read_len": READ_LEN, "pairs_per_gpu": args.pairs, "alg": "nw", "parallelism": "reads sharded over %d GPU(s) in file order (one library: avgDist / dedup gate / discordant-pair state exchanged over NCCL), full index replica per GPU" % world,
              "l2_policy": "inputs larger than L2: 2x%d MB of reads per step stream through; the 6.9 MB index stays L2-resident" % (args.pairs * READ_LEN // 1_000_000)}

    if args.impl == "reference":
        if rank != 0:
            return
        g, r1, r2, _, _ = make_workload(args.pairs)
        n_sample = args.cpu_sample or min(args.pairs, 150_000 * max(1, (os.cpu_count() or 1) // 4))
        vals = []
        for _ in range(args.warmup + args.steps):
            v, kind, cores, what, sec = time_reference(g, r1, r2, n_sample)
            vals.append((v, sec))
        vals = vals[args.warmup:]
        v = float(np.mean([x[0] for x in vals]))
        print(json.dumps({"impl": "reference", "metric": "read pairs mapped/sec", "value": v, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": 1000 * float(np.mean([x[1] for x in vals])), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int64",
                          "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores, "kind": kind, "sample": what},
                          "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    from mapcaller_b200 import api, simulate as sim
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local)
        dist_mod.init_process_group("nccl")
        dist = dist_mod

    g, r1, r2, seq, off = make_workload(args.pairs, seed_shift=rank)
    ix = api.Index.build(sim.encode(g))
    n_pairs = len(r1)
    ctx = api.Context(ix, paired=1, alg_ksw2=0, update_profile=1, want_alignments=0, device=local, shard_rank=rank, shard_count=world)
    if dist is not None:
        # the library's own NCCL communicator (NVLink / NVSwitch) for the end-of-pass profile reduction
        uid = [api.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)

    def barrier():
        if dist is not None:
            import torch
            dist.barrier(); torch.cuda.synchronize()

    # page-locked copies of the inputs for the end-to-end legs (allocated up front: the clock sampler below should not
    # see an idle GPU between the legs)
    pseq = api.pinned_array(seq.shape, np.uint8); pseq[:] = seq
    poff = api.pinned_array(off.shape, np.int64); poff[:] = off
    # ---- resident leg (value) ----
    ctx.stage_batch(seq, off, 0)
    def one_pass():
        ctx.reset(); ctx.map_staged(0)
        if dist is not None:
            ctx.profile_allreduce()      # every rank ends the pass with the whole-library profile

    for _ in range(args.warmup):
        one_pass()
    ctx.reset_stats()
    sampler = ClockSampler(local); sampler.start()
    barrier(); t0 = time.perf_counter()
    for _ in range(args.steps):
        one_pass()
    barrier(); wall_resident = time.perf_counter() - t0
    st = ctx.stats()
    dev_s = st["ms_total"] / 1000.0
    totals = ctx.totals()

    # ---- end-to-end leg (host buffers through mc_map_batch) ----
    # the step's inputs sit in page-locked host memory (mc_host_alloc), as the contract asks; every step copies them to
    # the device and reads the per-pair / per-chunk results back
    # (a) the double-buffered feed of a host that streams batches (mc_stage_batch_async + mc_map_staged): the copy of step
    #     i+1's inputs is queued before step i is mapped, so it overlaps that mapping; every step's bytes cross PCIe inside
    #     the timed region (steps copies for steps steps, the first one is not overlapped with anything)
    def feed(steps):
        ctx.stage_batch_async(pseq, poff, 0)
        for i in range(steps):
            ctx.reset()
            if i + 1 < steps:
                ctx.stage_batch_async(pseq, poff, (i + 1) & 1)
            ctx.map_staged(i & 1, copy=False)
            if dist is not None:
                ctx.profile_allreduce()
    feed(max(2, args.warmup // 2))
    barrier(); t0 = time.perf_counter()
    feed(args.steps)
    barrier(); wall_e2e = time.perf_counter() - t0
    # (b) one synchronous mc_map_batch call per step (upload in four pieces, seeding overlapped)
    for _ in range(max(1, args.warmup // 2)):
        ctx.reset(); ctx.map_batch(pseq, poff, copy=False)
    barrier(); t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.reset(); res = ctx.map_batch(pseq, poff, copy=False)
        if dist is not None:
            ctx.profile_allreduce()
    barrier(); wall_sync = time.perf_counter() - t0
    clocks = sampler.stop()   # sampled every 20 ms across the resident and both end-to-end legs, back to back
    # same call with ordinary pageable numpy arrays (bounced through pinned buffers inside the library)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.reset(); ctx.map_batch(seq, off, copy=False)
    wall_pageable = time.perf_counter() - t0
    # (c) from FASTQ text: the raw bytes of the two mate files in page-locked memory, parsed on the device (mc_ingest_fastq)
    ft1, ft2 = sim.fastq_text(r1, 1), sim.fastq_text(r2, 2)
    pf1 = api.pinned_array(ft1.shape, np.uint8); pf1[:] = ft1
    pf2 = api.pinned_array(ft2.shape, np.uint8); pf2[:] = ft2
    del ft1, ft2
    for _ in range(2):
        ctx.reset(); ctx.ingest_fastq(pf1, pf2, slot=2); ctx.map_staged(2, copy=False)
    barrier(); t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.reset(); ctx.ingest_fastq(pf1, pf2, slot=2); ctx.map_staged(2, copy=False)
        if dist is not None:
            ctx.profile_allreduce()
    barrier(); wall_fastq = time.perf_counter() - t0
    n_reads = 2 * n_pairs
    h2d = int(seq.nbytes + (n_reads + 1) * 8 + 5 * ((n_reads + 199) // 200))
    d2h = int(((n_reads + 199) // 200) * (32 + 8) + 72 + 64 + 8 * 4 + 40 * 2048)   # per-chunk sums + intervals, counters, cursors, the list of discordant pairs (<= ~2 k records here)

    if dist is not None:
        import torch
        # multi-GPU: the pass includes the NCCL reduction, which the per-context CUDA events do not see -> the time between the
        # barriers (each with a device synchronize) is the step time, max over ranks
        # device time of the resident leg = CUDA events of the mapping (ms_total) + of the reduction (ms_reduce) on the
        # context's stream, max over ranks; the time between the barriers (wall, also max over ranks) is reported beside it
        t = torch.tensor([(st["ms_total"] + st["ms_reduce"]) / 1000.0, wall_e2e, wall_resident, wall_sync], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s, wall_e2e, wall_resident, wall_sync = [float(x) for x in t.tolist()]
    total_pairs = n_pairs * world * args.steps

    if rank == 0:
        peak, peak_src = measured_peak()
        seed_bytes = st["seed_blocks"] * 64 / args.steps
        seed_ms = st["ms_seed"] / args.steps
        achieved = seed_bytes / (seed_ms * 1e-3) / 1e9 if seed_ms > 0 else 0.0
        tr = recorded_traffic()
        line = {"metric": "read pairs mapped/sec", "value": total_pairs / dev_s, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1000 * dev_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int64",
                "data": "synthetic", "config": config, "clocks": clocks, "gpu_launches": int(st["kernel_launches"]),
                "e2e": {"value": total_pairs / wall_e2e, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": 1000 * wall_e2e / args.steps,
                        "how": "mc_stage_batch_async(step i+1) + mc_map_staged(step i) from page-locked host arrays: every step's inputs are copied inside the timed region, the copy overlaps the previous step's mapping",
                        "sync_call_pairs_per_s": total_pairs / wall_sync, "sync_call_ms_per_step": 1000 * wall_sync / args.steps},
                "roofline": {"kernel": "mc_seed_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": (tr or {}).get("dram_bytes_per_launch"), "algorithmic_bytes_per_launch": seed_bytes, "ms_per_launch": seed_ms,
                             "note": "the 4.6 MB compact index is L2-resident at this genome size, so achieved (64 B x reference blocks, SURVEY 8d) counts L2-served bytes against the HBM peak and can exceed it; hbm_regime = the same kernel on a 248 Mbp genome (committed ncu capture, tools/big_genome.py)",
                             "hbm_regime": {k: v for k, v in ((tr or {}).get("hbm_regime") or {}).items() if k != "metrics"} or None},
                "stages_ms_per_step": {k: st[k] / args.steps for k in ("ms_seed", "ms_locate", "ms_cluster", "ms_pair", "ms_align", "ms_profile", "ms_h2d", "ms_d2h", "ms_total")},
                "work_per_step": {k: st[k] / args.steps for k in ("seed_blocks", "locate_blocks", "sa_reads", "dp_cells", "dp_tasks", "profile_columns")},
                "locate_gbs": (st["locate_blocks"] * 64 + st["sa_reads"] * 8) / (st["ms_locate"] * 1e-3) / 1e9 if st["ms_locate"] > 0 else None,
                "dp_gcups": st["dp_cells"] / (st["ms_align"] * 1e-3) / 1e9 if st["ms_align"] > 0 else None,
                "wall_ms_per_step_resident": 1000 * wall_resident / args.steps, "e2e_pageable_pairs_per_s": n_pairs * args.steps / wall_pageable,
                "e2e_from_fastq_text": {"pairs_per_s_this_rank": n_pairs * args.steps / wall_fastq, "ms_per_step": 1000 * wall_fastq / args.steps, "fastq_bytes_per_step": int(pf1.nbytes + pf2.nbytes),
                                        "how": "raw bytes of the two mate FASTQ files (page-locked) -> mc_ingest_fastq (records found on the device) -> mc_map_staged"},
                "check": {"mapped_fraction": totals["total_mapped"] / max(1, totals["total_reads"]), "avg_dist": totals["avg_dist"]}}
        if world == 1:
            n_sample = args.cpu_sample or min(n_pairs, 150_000 * max(1, (os.cpu_count() or 1) // 4))
            try:
                v, kind, cores, what, sec = time_reference(g, r1, r2, n_sample)
                line["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": cores, "kind": kind, "sample": what, "seconds": sec}
            except Exception as e:  # the baseline is a reported number, never a reason to lose the bench line
                line["cpu_baseline"] = {"value": None, "unit": "pairs/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %s" % str(e)[:200]}
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
