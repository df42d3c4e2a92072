"""One-off check of the GPU index builder beyond 2^32 text symbols: build the same genome with the host builder and with
mc_index_build_gpu and compare every array of the two images (run under gpurun; the log goes to profiles/).
usage: python tools/index_big_check.py [genome_bp] [contigs] [skip_host]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mapcaller_b200 import api, simulate as sim

G = int(sys.argv[1]) if len(sys.argv) > 1 else 2_200_000_000
NC = int(sys.argv[2]) if len(sys.argv) > 2 else 10
SKIP_HOST = len(sys.argv) > 3
t = time.time(); g = sim.genome(G, 21, n_dup=2000, repeat_frac=0.15); codes = sim.encode(g); del g
print("genome %d bp (text %d symbols, 2^32 = %d) %.1f s" % (G, 2 * G, 1 << 32, time.time() - t), flush=True)
lens = [G // NC] * (NC - 1); lens.append(G - sum(lens)); names = ["c%d" % i for i in range(NC)]
os.environ["MC_DEBUG"] = "1"
t = time.time(); gpu = api.Index.build(codes, chrom_len=lens, chrom_name=names, gpu_device=0); print("mc_index_build_gpu %.1f s" % (time.time() - t), flush=True)
va = gpu.view_arrays()
print("gpu image: primary %d, L2 %s, bwt words %d, sa entries %d" % (va["primary"], va["L2"], len(va["bwt"]), len(va["sa"])), flush=True)
if not SKIP_HOST:
    t = time.time(); cpu = api.Index.build(codes, chrom_len=lens, chrom_name=names); print("mc_index_build (host, %d threads) %.1f s" % (os.cpu_count(), time.time() - t), flush=True)
    vb = cpu.view_arrays()
    ok = va["primary"] == vb["primary"] and list(va["L2"]) == list(vb["L2"]) and np.array_equal(va["bwt"], vb["bwt"]) and np.array_equal(va["sa"], vb["sa"]) and np.array_equal(va["pac"], vb["pac"])
    print("IMAGES EQUAL" if ok else "IMAGES DIFFER", flush=True)
    sys.exit(0 if ok else 1)
