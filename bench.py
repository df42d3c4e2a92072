#!/usr/bin/env python
"""bench.py -- read pairs mapped / s of the MapCaller hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pairs P] [--genome BP]

Workload (config.workload): BASELINE.json configs[2] (SURVEY.md 8d "chr1-size") -- one contig of 248,956,422 bp, 85 %
uniform ACGT + 15 % copies of three repeat families (300 bp / 1 kb / 6 kb consensus, 5-15 % divergence) + 2000 copied
segments, reads simulated from a mutant with SNP 1000/Mb and small indels 100/Mb: 10 M pairs of 2x150 bp per GPU,
fragment ~N(450, 50), 0.3 % substitution errors, -alg nw, VCF profile on.  The index is built on the GPU
(mc_index_build_gpu) and is 0.37 GB (+ 0.12 GB sampled SA), i.e. the FM-index gathers come from HBM, not from L2.

One STEP = one complete pass of the hot path over the library, as one run of the reference's Mapping(): reset context
(empty profile, avgDist = 1000), then the library in batches of 2 M pairs (seeding, locate, clustering, pairing, rescue,
gapped fills, pair statistics, pile-up profile update); with N > 1 GPUs the N ranks map ONE library in file order
(ordered exchange of avgDist / dedup gate / discordant-pair state inside every batch) and the pass ends with the NCCL
profile reduction.

  value     device-timed throughput, reads resident in HBM (mc_stage_batch once, step = mc_reset + mc_map_staged per
            batch): CUDA events on the library's stream around every reset / batch / reduction, max over ranks
  e2e       FASTQ TEXT in host memory -> result in host memory, through the C ABI, wall clock: every step copies the raw
            bytes of both mate files host -> device (page-locked), parses them on the device (mc_ingest_fastq, a feeder
            thread stages block i+1 while block i is mapped), maps them, runs the variant-calling scan on the resident
            profile (mc_variant_scan) and brings the variant records + block depths back.  Bytes are the library's own
            counters of what it copied.  e2e.sam = the same loop producing the SAM text of every read as well.
  roofline  mc_seed_kernel: algorithmic bytes = 64 B x occ blocks the reference algorithm touches (counted by the kernel,
            equal to the oracle's count) / CUDA-event time of the seed launches of the timed steps; `traffic` = DRAM bytes per
            launch from the committed ncu --set full capture of this command (profiles/), if present.
            roofline_locate / roofline_profile: the same for the locate and profile-update stages.
  cpu_baseline  the unmodified reference (oracle/_ref, Mapping() on all host threads) on a prefix of the same reads
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
ALL_CPUS = os.sched_getaffinity(0)      # before any NUMA pinning: the reference arm gets every host core it was given

GENOME_LEN = 248_956_422
READ_LEN = 150               # set from the chosen config in main()
FULL_PAIRS = 10_000_000
BATCH_PAIRS = 2_000_000
# GRCh38 primary assembly, chr1..22, X, Y (configs[3] / configs[4]: "24 contigs with GRCh38 primary lengths")
GRCH38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717, 133797422, 135086622, 133275309,
          114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616, 64444167, 46709983, 50818468, 156040895, 57227415]
# BASELINE.json configs the bench can run.  2 is the default (the driver's BENCH / SCALE lines); 3 and 4 are the GRCh38-sized
# runs (text of 6.2 G symbols: 64-bit rows, the wide index layout), committed as logs under profiles/.
CONFIGS = {
    2: dict(name="configs[2]: chr1-sized", genome=GENOME_LEN, contigs=None, read_len=150, ksw2=0, frag_mean=450, small_indel=100, large_indel=0, read_indel=0.0),
    3: dict(name="configs[3]: GRCh38-sized", genome=sum(GRCH38), contigs=GRCH38, read_len=150, ksw2=0, frag_mean=450, small_indel=100, large_indel=0, read_indel=0.0),
    # 250-bp reads with alignment strings for nearly every read: 500 k pairs per batch keep the batch's arenas below 2^31 bytes
    4: dict(name="configs[4]: GRCh38-sized, ksw2 stress", genome=sum(GRCH38), contigs=GRCH38, read_len=250, ksw2=1, frag_mean=600, small_indel=2000, large_indel=500, read_indel=0.4,
            batch=500_000, pairs=2_500_000),
}
CFG = CONFIGS[2]
SIM_BLOCK = 250_000          # simulate_pairs_fast generates independent blocks of this many pairs
CHECK_PAIRS = 500_000        # per rank, for the N-GPU == 1-GPU profile check


def make_genome(genome_len: int):
    from mapcaller_b200 import simulate as sim
    g = sim.genome(genome_len, 13, n_dup=2000 if genome_len > 50_000_000 else 200, repeat_frac=0.15)
    if genome_len > 1_000_000_000 or CFG["ksw2"]:      # the vectorised variant of the same model (no per-event Python loop)
        mut = sim.mutate_fast(g, 14, snp_per_mb=1000, small_indel_per_mb=CFG["small_indel"], large_indel_per_mb=CFG["large_indel"])
    else:
        mut, _ = sim.mutate(g, 14, snp_per_mb=1000, small_indel_per_mb=CFG["small_indel"], large_indel_per_mb=CFG["large_indel"], sv_per_mb=0)
    return g, mut


def make_reads(mut, n_pairs: int, rank: int, first_block: int = 0):
    from mapcaller_b200 import simulate as sim
    return sim.simulate_pairs_fast(mut, n_pairs, READ_LEN, seed=15 + 1000 * rank, frag_mean=CFG["frag_mean"], frag_sd=50, sub_rate=0.003,
                                   block=SIM_BLOCK, first_block=first_block, indel_read_frac=CFG["read_indel"])


def contig_lengths(genome_len: int):
    """The contig table of the chosen config, rescaled when --genome overrides the size (None = one contig)."""
    if not CFG["contigs"]:
        return None, None
    lens = [int(x * genome_len / CFG["genome"]) for x in CFG["contigs"]]
    lens[-1] += genome_len - sum(lens)
    return lens, ["chr%s" % (i + 1 if i < 22 else "XY"[i - 22]) for i in range(len(lens))]


def build_index(g, device):
    from mapcaller_b200 import api, simulate as sim
    codes = sim.encode(g)
    lens, names = contig_lengths(len(g))
    if device is not None:
        try:
            return api.Index.build(codes, chrom_len=lens, chrom_name=names, gpu_device=device), "mc_index_build_gpu"
        except api.McError as e:
            print("[bench] GPU index build failed, host build instead: %s" % e, file=sys.stderr)
    return api.Index.build(codes, chrom_len=lens, chrom_name=names), "mc_index_build (host)"


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


def recorded_traffic(genome_len: int):
    """DRAM bytes per launch of the kernels from the committed ncu --set full capture of this command (profiles/), or {}."""
    p = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    try:
        t = json.load(open(p))
        return t if int(t.get("genome_bp", 0)) == genome_len else {}
    except Exception:
        return {}


def bind_to_gpu_numa(device: int):
    """Pin this process (threads and the first-touch / page-locked host memory it allocates afterwards) to the CPUs next to its
    GPU: the NUMA node sysfs names for the GPU's PCI device, else the CPU set NVML reports as ideal for it.  Without this,
    eight ranks' FASTQ buffers land on whatever socket their process started on and half the GPUs pull their input across
    the inter-socket link.  Returns a description for the JSON line (None when the box exposes nothing)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device)
    except Exception:
        return None
    try:
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus.lower()[-12:]).read())
        if node >= 0:
            cpus = []
            for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
                a, _, b = part.partition("-")
                cpus += list(range(int(a), int(b or a) + 1))
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    try:
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * w + b for w, x in enumerate(words) for b in range(64) if (int(x) >> b) & 1 and 64 * w + b < n_cpu]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed and len(allowed) < n_cpu:
            os.sched_setaffinity(0, allowed)
            return "nvml cpu set %d-%d (%d cpus)" % (allowed[0], allowed[-1], len(allowed))
    except Exception:
        pass
    return None


# --------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: Mapping() of the unmodified reference on the host cores
# --------------------------------------------------------------------------------------------------------------
_REF_CODE = r"""
import sys, time, os
sys.path.insert(0, %(tests)r)
import ref_oracle as ro
ro.load(%(prefix)r); ro.set_params(threads=%(cores)d)
for i in range(%(runs)d):
    ro.lib().mcref_reset_state(); t = time.perf_counter()
    ro.lib().mcref_run_mapping(%(f1)r.encode(), %(f2)r.encode(), b'', %(cores)d, 1)
    print('SECONDS', time.perf_counter() - t, ro.counters()['reads'], flush=True)
"""


def time_reference(ix, text1, text2, n_sample: int, rec_bytes: int, runs: int):
    """Seconds per run of the reference's own CPU implementation on the first n_sample pairs, all host threads.
    Returns (list of seconds, kind, cores, description)."""
    cores = len(ALL_CPUS) or 1
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libmcref.so")
    td = tempfile.mkdtemp(prefix="mcbench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    prefix = os.path.join(td, "idx")
    try:
        ix.save(prefix)
        if os.path.exists(ref_so):
            f1, f2 = os.path.join(td, "r1.fq"), os.path.join(td, "r2.fq")
            text1[:n_sample * rec_bytes].tofile(f1); text2[:n_sample * rec_bytes].tofile(f2)
            code = _REF_CODE % dict(tests=os.path.join(ROOT, "tests"), prefix=prefix, cores=cores, runs=runs, f1=f1, f2=f2)
            out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, preexec_fn=lambda: os.sched_setaffinity(0, ALL_CPUS))
            secs = [float(l.split()[1]) for l in out.stdout.splitlines() if l.startswith("SECONDS")]
            if len(secs) != runs:
                raise RuntimeError("reference run failed: " + out.stderr[-500:])
            return secs, "reference", cores, "first %d pairs of the library; Mapping() of the unmodified reference (oracle/_ref), -t %d, VCF profile on, FASTQ parse included, index load excluded" % (n_sample, cores)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import cpu_oracle
        from mapcaller_b200 import simulate as sim
        L = READ_LEN
        w = rec_bytes - (2 * L + 5)
        m1 = text1[:n_sample * rec_bytes].reshape(n_sample, rec_bytes)[:, w + 1:w + 1 + L]
        m2 = text2[:n_sample * rec_bytes].reshape(n_sample, rec_bytes)[:, w + 1:w + 1 + L]
        seq, off = sim.interleave(m1, m2)
        orc = cpu_oracle.Oracle(prefix)
        secs = []
        for _ in range(runs):
            t = time.perf_counter(); orc.map_reads(seq, off, True, True); secs.append(time.perf_counter() - t)
        orc.close()
        return secs, "port", 1, "first %d pairs of the library; CPU restatement oracle/libmcoracle.so, single thread" % n_sample
    finally:
        import shutil
        shutil.rmtree(td, ignore_errors=True)


def cpu_sample_size(n_pairs: int, override: int, per_core: int = 60_000) -> int:
    n = override or min(n_pairs, per_core * (os.cpu_count() or 1))
    return max(100, n - n % 100)


# --------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json configs[] index (default 2: chr1-sized, 2x150, nw)")
    ap.add_argument("--pairs", type=int, default=0, help="library size per GPU (default: 10 M pairs; 2.5 M for configs[4])")
    ap.add_argument("--genome", type=int, default=0, help="genome size in bp (default: the config's)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="pairs for the cpu_baseline leg (default: sized for ~10-30 s)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sam", action="store_true", help="skip the SAM flavour of the end-to-end leg")
    ap.add_argument("--cache-dir", default="", help="profiling runs: keep the mutant genome and the index files here (e.g. /dev/shm/mc) and reuse them, so that a run under ncu does not profile the index build")
    ap.add_argument("--allreduce", action="store_true", help="N > 1: end the library with the whole-array all-reduce instead of the reduce-scatter")
    ap.add_argument("--trace-e2e", action="store_true", help="print a host-side timeline of one end-to-end step to stderr")
    ap.add_argument("--resident-only", action="store_true", help="profiling runs: only the resident leg (the JSON line then has no e2e)")
    args = ap.parse_args()
    global CFG, READ_LEN, BATCH_PAIRS
    CFG = CONFIGS[args.config]; READ_LEN = CFG["read_len"]
    BATCH_PAIRS = CFG.get("batch", BATCH_PAIRS)
    args.pairs = args.pairs or CFG.get("pairs", FULL_PAIRS)
    args.genome = args.genome or CFG["genome"]
    args.pairs -= args.pairs % SIM_BLOCK
    assert args.pairs >= SIM_BLOCK, "--pairs must be at least %d" % SIM_BLOCK

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    n_batches = (args.pairs + BATCH_PAIRS - 1) // BATCH_PAIRS
    config = {"workload": "%s %d bp synthetic reference (%s, 15 %% repeat families), 2x%d bp simulated PE reads, %d pairs per GPU in batches of %d, -alg %s, VCF profile on"
                          % (CFG["name"], args.genome, "24 contigs" if CFG["contigs"] else "one contig", READ_LEN, args.pairs, min(BATCH_PAIRS, args.pairs), "ksw2" if CFG["ksw2"] else "nw"),
              "genome_bp": args.genome, "read_len": READ_LEN, "pairs_per_gpu": args.pairs, "batch_pairs": min(BATCH_PAIRS, args.pairs), "alg": "ksw2" if CFG["ksw2"] else "nw",
              "parallelism": "reads sharded over %d GPU(s) in file order (one library: avgDist / dedup gate / discordant-pair state exchanged over NCCL inside every batch; the pass ends with %s of the profile), full index replica per GPU"
                             % (world, "ncclAllReduce" if args.allreduce else "ncclReduceScatter (every rank finishes one genome tile)")
                             if world > 1 else "one GPU, one index replica",
              "l2_policy": "inputs larger than L2: index %.2f GB + %.1f GB of reads per batch stream from HBM" % (args.genome * 1.5 / 1e9, 2 * min(BATCH_PAIRS, args.pairs) * READ_LEN / 1e9)}
    numa = bind_to_gpu_numa(local) if args.impl == "ours" else None

    if args.impl == "reference":
        if rank != 0:
            return
        from mapcaller_b200 import simulate as sim
        g, mut = make_genome(args.genome)
        n_sample = cpu_sample_size(args.pairs, args.cpu_sample)
        r1, r2 = make_reads(mut, (n_sample + SIM_BLOCK - 1) // SIM_BLOCK * SIM_BLOCK, 0)
        t1, t2 = sim.fastq_text(r1[:n_sample], 1), sim.fastq_text(r2[:n_sample], 2)
        has_gpu = subprocess.call(["nvidia-smi", "-L"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) == 0 if _which("nvidia-smi") else False
        ix, how = build_index(g, 0 if has_gpu else None)
        secs, kind, cores, what = time_reference(ix, t1, t2, n_sample, len(t1) // n_sample, args.warmup + args.steps)
        secs = secs[args.warmup:]
        v = n_sample * len(secs) / sum(secs)
        print(json.dumps({"impl": "reference", "metric": "read pairs mapped/sec", "value": v, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": 1000 * float(np.mean(secs)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int64",
                          "data": "synthetic", "config": dict(config, reference_sample_pairs_per_step=n_sample, index_built_by=how),
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

    t_setup = time.perf_counter()
    # GRCh38-sized runs on several GPUs: rank 0 makes the mutant genome and the index once and the other ranks load them from
    # /dev/shm (every rank generating 3.1 Gbp on its own costs a minute of box time and 15 GB of host memory per rank)
    own_cache = world > 1 and args.genome > 1_000_000_000 and not args.cache_dir
    if own_cache:
        args.cache_dir = "/dev/shm/mc_bench_%s" % os.environ.get("MASTER_PORT", "0")
    cache = os.path.join(args.cache_dir, "c%d_%d" % (args.config, args.genome)) if args.cache_dir else ""
    have = bool(cache) and os.path.exists(cache + ".bwt") and os.path.exists(cache + ".mut.npy") and not own_cache
    if not have and (rank == 0 or not cache):
        g, mut = make_genome(args.genome)
        ix, index_how = build_index(g, local)
        del g
        if cache:
            os.makedirs(args.cache_dir, exist_ok=True)
            np.save(cache + ".mut.npy", mut); ix.save(cache)
    if cache and dist is not None:
        dist.barrier()
    if have or (cache and rank != 0):
        mut = np.load(cache + ".mut.npy", mmap_mode="r")
        ix, index_how = api.Index.load(cache), "mc_index_load (files written by rank 0 / an earlier run with mc_index_build_gpu)"
    if own_cache:
        dist.barrier()
        if rank == 0:
            import shutil
            mut = np.array(mut)          # rank 0 keeps its copy in memory; the files go away
            shutil.rmtree(args.cache_dir, ignore_errors=True)
    t_index = time.perf_counter() - t_setup
    r1, r2 = make_reads(mut, args.pairs, rank)
    n_pairs = len(r1)
    big = args.genome > 1_000_000_000     # GRCh38-sized: a second context (102 GB of profile) does not fit beside the first
    if big or world == 1:
        del mut
    ctx = api.Context(ix, paired=1, alg_ksw2=CFG["ksw2"], update_profile=1, want_alignments=0, device=local, shard_rank=rank, shard_count=world)
    if dist is not None:
        # the library's own NCCL communicator (NVLink / NVSwitch): ordered exchange inside the batches + profile reduction
        uid = [api.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)

    def barrier():
        if dist is not None:
            import torch
            dist.barrier(); torch.cuda.synchronize()

    # inputs: packed batches staged once in HBM (resident leg), FASTQ text of every batch in page-locked host memory (e2e)
    bounds = [(b * BATCH_PAIRS, min(n_pairs, (b + 1) * BATCH_PAIRS)) for b in range(n_batches)]
    ptext = []
    for b, (p0, p1) in enumerate(bounds):
        seq, off = sim.interleave(r1[p0:p1], r2[p0:p1])
        ctx.stage_batch(seq, off, b)
        del seq, off
        pair = []
        for mate, r in ((1, r1), (2, r2)):
            t = sim.fastq_text(r[p0:p1], mate, first=p0)
            pt = api.pinned_array(t.shape, np.uint8); pt[:] = t
            pair.append(pt)
            del t
        ptext.append(pair)
    rec_bytes = len(ptext[0][0]) // (bounds[0][1] - bounds[0][0])
    chk1, chk2 = r1[:CHECK_PAIRS].copy(), r2[:CHECK_PAIRS].copy()
    del r1, r2
    t_setup = time.perf_counter() - t_setup
    S_IN0 = n_batches                            # first of the three device slots of the FASTQ feed
    assert n_batches + 3 <= 8

    # The N ranks end a library with mc_profile_reduce_scatter (each keeps the sums of one genome tile; the variant scan and the
    # checksums then run on the N tiles at once); --allreduce keeps the whole-array ncclAllReduce of round 1 for comparison.
    end_of_library = ctx.profile_allreduce if args.allreduce else ctx.profile_reduce_scatter

    # ---- resident leg (value) ----
    def one_pass():
        ctx.reset()
        for b in range(n_batches):
            ctx.map_staged(b)
        if dist is not None:
            end_of_library()             # every rank ends the pass with the library's counters of its genome tile

    for _ in range(args.warmup):
        one_pass()
    ctx.reset_stats()
    sampler = ClockSampler(local); sampler.start()
    barrier(); t0 = time.perf_counter()
    for _ in range(args.steps):
        one_pass()
    barrier(); wall_resident = time.perf_counter() - t0
    st = ctx.stats()
    dev_s = max((st["ms_total"] + st["ms_reduce"]) / 1000.0, 1e-12)
    totals = ctx.totals()
    resident_fp = ctx.profile_checksum()

    # ---- end-to-end leg: FASTQ text (host) -> variant records (host) ----
    # The end-to-end feed cuts the library's FASTQ text into blocks of its own: the first one small (the pipeline starts mapping
    # after 0.2 M pairs are on the device instead of 2 M), then the 2 M-pair batches, and a small last one; block boundaries are multiples of the
    # reference's 200-read chunk, so the result is the resident path's (check.fastq_path_equals_resident).
    FIRST = max(100, min(200_000, BATCH_PAIRS // 10) // 100 * 100)
    blocks = []
    for b, (p0, p1) in enumerate(bounds):
        cuts = [0, p1 - p0]
        if world > 1:
            pass   # N ranks on one library: file order = rank after rank inside every batch, so both legs must cut the library alike
        elif b == 0 and p1 - p0 > FIRST:
            cuts.insert(1, FIRST)                      # a small first block: mapping starts early
        if world == 1 and b == len(bounds) - 1 and cuts[-1] - cuts[-2] > FIRST:
            cuts.insert(len(cuts) - 1, p1 - p0 - FIRST)  # and a small last one: little is left to map when the last byte has arrived
        for c0, c1 in zip(cuts[:-1], cuts[1:]):
            blocks.append((ptext[b][0][c0 * rec_bytes:c1 * rec_bytes], ptext[b][1][c0 * rec_bytes:c1 * rec_bytes]))
    n_blocks = len(blocks)

    def fastq_pass(sam: bool):
        # three device slots in flight: block b+1 on the wire (mc_ingest_prefetch, plain DMA), block b being parsed
        # (mc_ingest_fastq) by the feeder thread, block b-1 being mapped by this thread
        T0 = time.perf_counter(); tl = []
        mark = (lambda what, b: tl.append((1000 * (time.perf_counter() - T0), what, b))) if args.trace_e2e else (lambda what, b: None)
        ctx.reset(); mark("reset done", -1)
        err = []
        ready = [threading.Semaphore(0) for _ in range(3)]; free = [threading.Semaphore(1) for _ in range(3)]
        slot = lambda b: S_IN0 + b % 3

        def feeder():
            try:
                free[0].acquire(); ctx.ingest_prefetch(blocks[0][0], blocks[0][1], slot=slot(0))
                for b in range(n_blocks):
                    if b + 1 < n_blocks:
                        free[(b + 1) % 3].acquire(); ctx.ingest_prefetch(blocks[b + 1][0], blocks[b + 1][1], slot=slot(b + 1))
                    mark("ingest begin", b)
                    ctx.ingest_fastq(blocks[b][0], blocks[b][1], slot=slot(b), final=True)
                    mark("ingest end", b)
                    ready[b % 3].release()
            except Exception as e:      # surfaces in the main thread
                err.append(e)
                for s_ in ready:
                    s_.release()
        th = threading.Thread(target=feeder); th.start()
        sam_bytes = 0
        for b in range(n_blocks):
            ready[b % 3].acquire()
            if err:
                break
            mark("map begin", b)
            ctx.map_staged(slot(b))
            mark("map end", b)
            if sam:
                sam_bytes += ctx.sam_text_raw(slot(b)); mark("sam text end", b)
            free[b % 3].release()
        th.join()
        if err:
            raise err[0]
        if dist is not None:
            end_of_library()
        mark("scan begin", -1)
        n_var, n_blk = ctx.variant_scan_raw()
        mark("scan end", -1)
        if args.trace_e2e and rank == 0:
            for t, what, b in sorted(tl):
                print("[e2e %s] %9.3f ms  %-14s batch %d" % ("sam" if sam else "vcf", t, what, b), file=sys.stderr)
        return n_var, sam_bytes

    def timed_fastq(sam: bool):
        for _ in range(max(2, args.warmup // 2)):
            fastq_pass(sam)
        s0 = ctx.stats()
        barrier(); t0 = time.perf_counter()
        for _ in range(args.steps):
            n_var, sam_bytes = fastq_pass(sam)
        barrier(); wall = time.perf_counter() - t0
        s1 = ctx.stats()
        return wall, (s1["h2d_bytes"] - s0["h2d_bytes"]) // args.steps, (s1["d2h_bytes"] - s0["d2h_bytes"]) // args.steps, n_var, sam_bytes

    if args.resident_only:
        wall_e2e, h2d, d2h, n_var, e2e_fp, e2e_totals = float("inf"), 0, 0, 0, resident_fp, totals
    else:
        wall_e2e, h2d, d2h, n_var, _ = timed_fastq(False)
        e2e_fp = ctx.profile_checksum(); e2e_totals = ctx.totals()
    sam_leg = None
    if not args.no_sam and not args.resident_only and world == 1:      # the SAM flavour is a single-GPU figure
        w, h, d, _, sam_bytes = timed_fastq(True)
        sam_leg = (w, h, d, sam_bytes)

    # ---- packed-read feed (round-1 definition, kept for continuity): pre-parsed reads in page-locked memory ----
    p0, p1 = bounds[0]
    seq0 = api.pinned_array(((p1 - p0) * 2 * READ_LEN,), np.uint8); off0 = api.pinned_array(((p1 - p0) * 2 + 1,), np.int64)
    w = rec_bytes - (2 * READ_LEN + 5)
    m1 = ptext[0][0].reshape(p1 - p0, rec_bytes)[:, w + 1:w + 1 + READ_LEN]; m2 = ptext[0][1].reshape(p1 - p0, rec_bytes)[:, w + 1:w + 1 + READ_LEN]
    seq0.reshape(-1, 2, READ_LEN)[:, 0] = m1; seq0.reshape(-1, 2, READ_LEN)[:, 1] = m2
    off0[:] = np.arange(len(off0), dtype=np.int64) * READ_LEN
    wall_packed = float("inf")
    if not args.resident_only:
        ctx.reset(); ctx.map_batch(seq0, off0, copy=False)
        barrier(); t0 = time.perf_counter()
        for _ in range(args.steps):
            ctx.reset(); ctx.map_batch(seq0, off0, copy=False)
        barrier(); wall_packed = time.perf_counter() - t0
    clocks = sampler.stop()   # sampled every 20 ms across the resident and the end-to-end legs, back to back

    # ---- N GPUs on one library == one GPU on the same reads (bit-exact profile and totals) ----
    check = {"mapped_fraction": totals["total_mapped"] / max(1, totals["total_reads"]), "avg_dist": totals["avg_dist"],
             "fastq_path_equals_resident": bool(e2e_fp == resident_fp and e2e_totals == totals), "variant_records": int(n_var)}
    if dist is not None and big:
        check["n_gpu_equals_1_gpu"] = "not run: the single-GPU control context does not fit beside a GRCh38-sized profile"
    elif dist is not None:
        cseq, coff = sim.interleave(chk1, chk2)
        ctx.reset(); ctx.map_batch(cseq, coff, copy=False); end_of_library()
        multi_fp, multi_tot = ctx.profile_checksum(), ctx.totals()
        if rank == 0:
            solo = api.Context(ix, paired=1, alg_ksw2=CFG["ksw2"], update_profile=1, want_alignments=0, device=local)
            for r in range(world):
                a1, a2 = (chk1, chk2) if r == 0 else make_reads(mut, CHECK_PAIRS, r)
                s_, o_ = sim.interleave(a1, a2)
                solo.map_batch(s_, o_, copy=False)
            check["n_gpu_equals_1_gpu"] = bool(solo.profile_checksum() == multi_fp and solo.totals() == multi_tot)
            check["n_gpu_check_pairs"] = CHECK_PAIRS * world
            solo.close()
        barrier()

    if dist is not None:
        import torch
        t = torch.tensor([dev_s, wall_e2e, wall_resident, wall_packed, sam_leg[0] if sam_leg else 0.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s, wall_e2e, wall_resident, wall_packed, wsam = [float(x) for x in t.tolist()]
        if sam_leg:
            sam_leg = (wsam,) + sam_leg[1:]
    total_pairs = n_pairs * world * args.steps

    if rank == 0:
        peak, peak_src = measured_peak()
        tr = recorded_traffic(args.genome)
        launches = n_batches * args.steps

        def roof(kernel, alg_bytes, ms, note):
            a = alg_bytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
            t = (tr.get("kernels") or {}).get(kernel) or {}
            o = {"kernel": kernel, "bound": "hbm", "achieved": a, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": a / peak,
                 "traffic": t.get("dram_bytes_per_launch"), "algorithmic_bytes_per_launch": alg_bytes / launches, "ms_per_launch": ms / launches, "launches_timed": launches, "note": note}
            if t.get("dram_bytes_per_launch") and ms > 0:
                o["dram_frac"] = t["dram_bytes_per_launch"] * launches / (ms * 1e-3) / 1e9 / peak
                o["traffic_source"] = tr.get("source")
            return o
        line = {"metric": "read pairs mapped/sec", "value": total_pairs / dev_s, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1000 * dev_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int64",
                "data": "synthetic", "config": config, "clocks": clocks, "gpu_launches": int(st["kernel_launches"]),
                "e2e": {"value": total_pairs / wall_e2e, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": 1000 * wall_e2e / args.steps,
                        "how": "FASTQ text of both mate files in page-locked host memory -> mc_ingest_fastq (H2D by mc_ingest_prefetch + record parse on the device; a feeder thread keeps block i+2 on the wire and parses block i+1 while block i is mapped) -> mc_map_staged -> "
                               + ("mc_profile_allreduce -> " if world > 1 else "") + "mc_variant_scan (variant records + block depths back in host memory); bytes = the library's copy counters (this rank)",
                        "variant_records": int(n_var)},
                "roofline": roof("mc_seed_kernel", st["seed_blocks"] * 64, st["ms_seed"], "algorithmic bytes = 64 B x occ blocks of the reference algorithm (kernel counter = oracle count); time = CUDA events around the seed launches of the timed steps"),
                "roofline_locate": roof("mc_locate_kernel", st["locate_blocks"] * 64 + st["sa_reads"] * 8, st["ms_locate"], "64 B per LF step + 8 B per sampled-SA read"),
                "roofline_profile": roof("profile stage", st["profile_columns"] * 32, st["ms_profile"], "reference formulation: 32 B (read + write of one MappingRecord_t) per column a counted read covers; the stage = gate sort + scatter + piece kernels"),
                "stages_ms_per_step": {k: st[k] / args.steps for k in ("ms_reset", "ms_seed", "ms_locate", "ms_cluster", "ms_pair", "ms_align", "ms_profile", "ms_h2d", "ms_d2h", "ms_reduce", "ms_total")},
                "work_per_step": {k: st[k] / args.steps for k in ("seed_blocks", "locate_blocks", "sa_reads", "dp_cells", "dp_tasks", "profile_columns", "profile_atomics")},
                "dp_gcups": st["dp_cells"] / (st["ms_align"] * 1e-3) / 1e9 if st["ms_align"] > 0 else None,
                "wall_ms_per_step_resident": 1000 * wall_resident / args.steps,
                "e2e_packed_reads": {"pairs_per_s_this_rank": (p1 - p0) * args.steps / wall_packed, "ms_per_batch": 1000 * wall_packed / args.steps,
                                     "how": "one batch of pre-parsed reads in page-locked memory through mc_map_batch (the round-1 e2e definition; no FASTQ parse, no result read-out)"},
                "setup_s": {"total": t_setup, "genome_and_index": t_index, "index": index_how}, "numa_node": numa, "check": check}
        if sam_leg:
            line["e2e"]["sam"] = {"value": total_pairs / sam_leg[0], "unit": "pairs/s", "ms_per_step": 1000 * sam_leg[0] / args.steps, "h2d_bytes_per_step": int(sam_leg[1]), "d2h_bytes_per_step": int(sam_leg[2]),
                                  "sam_text_bytes_per_step": int(sam_leg[3]), "how": "the same loop with the SAM text of every read assembled on the device from the resident FASTQ text and copied back (mc_sam_text)"}
        if world == 1 and not args.no_cpu:
            n_sample = cpu_sample_size(bounds[0][1], args.cpu_sample, 150_000)
            try:
                secs, kind, cores, what = time_reference(ix, ptext[0][0], ptext[0][1], n_sample, rec_bytes, 1)
                line["cpu_baseline"] = {"value": n_sample / secs[0], "unit": "pairs/s", "cores": cores, "kind": kind, "sample": what, "seconds": secs[0]}
            except Exception as e:  # the baseline is a reported number, never a reason to lose the bench line
                line["cpu_baseline"] = {"value": None, "unit": "pairs/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %s" % str(e)[:200]}
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def _which(x):
    import shutil
    return shutil.which(x)


if __name__ == "__main__":
    main()
