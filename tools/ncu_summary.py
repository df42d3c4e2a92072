"""Per-kernel summary of an `ncu --set full` capture (.ncu-rep): time, DRAM bytes, issue rate, lanes per instruction, occupancy.
Optionally writes profiles/kernel_traffic.json (DRAM bytes per launch of the LARGEST launch of each kernel), which bench.py
reports as roofline.traffic.
usage: python tools/ncu_summary.py <file.ncu-rep> [--traffic-json out.json --genome BP --source NAME]"""
import csv, json, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
COLS = [("gpu__time_duration.sum", "ms"), ("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr"), ("sm__inst_executed.avg.per_cycle_elapsed", "issue%"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("l1tex__t_sector_hit_rate.pct", "L1hit%"), ("lts__t_sector_hit_rate.pct", "L2hit%")]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}
def val(r, name):
    if name not in hdr: return None
    i = hdr.index(name)
    try: return float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)
    except ValueError: return None
kn = hdr.index("Kernel Name")
print("%-34s %9s %9s %9s %8s %7s %6s %6s %5s %7s %6s %7s %7s" % ("kernel", "ms", "DRAM rd GB", "wr GB", "TB/s", "IPC", "lanes", "occ%", "regs", "grid", "block", "L1hit%", "L2hit%"))
best = {}
for r in rows[2:]:
    name = r[kn].split("(")[0].replace("void ", "").split("<")[0]
    v = {k: val(r, m) for m, k in COLS}
    tb = ((v["rd"] or 0) + (v["wr"] or 0)) / (v["ms"] * 1e-3) / 1e12 if v["ms"] else 0
    print("%-34s %9.3f %9.3f %9.3f %8.2f %7.2f %6.1f %6.1f %5d %7d %6d %7.1f %7.1f" % (name[:34], v["ms"], (v["rd"] or 0) / 1e9, (v["wr"] or 0) / 1e9, tb, v["issue%"] or 0, v["lanes"] or 0,
          v["occ%"] or 0, v["regs"] or 0, v["grid"] or 0, v["block"] or 0, v["L1hit%"] or 0, v["L2hit%"] or 0))
    if name not in best or v["ms"] > best[name]["ms"]:
        best[name] = dict(ms=v["ms"], dram_bytes_per_launch=(v["rd"] or 0) + (v["wr"] or 0), dram_read=v["rd"], dram_write=v["wr"])
if "--traffic-json" in sys.argv:
    out = sys.argv[sys.argv.index("--traffic-json") + 1]
    g = int(sys.argv[sys.argv.index("--genome") + 1]) if "--genome" in sys.argv else 0
    src = sys.argv[sys.argv.index("--source") + 1] if "--source" in sys.argv else rep
    try: old = json.load(open(out))
    except Exception: old = {}
    if old.get("genome_bp") != g: old = {"genome_bp": g, "kernels": {}}
    old["source"] = src; old["kernels"].update(best)
    json.dump(old, open(out, "w"), indent=1)
