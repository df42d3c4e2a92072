"""Developer harness: walks bench.py's control flow (batching, feeder thread, checks, JSON assembly) on a box without a GPU by
pointing the ctypes binding at the host-compiled stage bodies (tools/hostemu/build.sh).  The numbers it prints mean nothing
(the harness has no CUDA events); not a test, not a fallback, never imported elsewhere.
usage: python tests/devharness/bench_dry_run.py [bench.py arguments]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mapcaller_b200 import api
api._LIB_PATH = os.path.join(ROOT, "tools", "hostemu", "_build", "libmc_hostemu.so")
import bench
bench.SIM_BLOCK = 2000; bench.BATCH_PAIRS = 6000; bench.CHECK_PAIRS = 2000
bench.ClockSampler.start = lambda self: None
sys.argv = ["bench.py"] + (sys.argv[1:] or ["--genome", "300000", "--pairs", "16000", "--steps", "2", "--warmup", "1", "--cpu-sample", "2000"])
bench.main()
