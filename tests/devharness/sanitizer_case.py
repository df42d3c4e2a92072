"""Two parity cases (repeats / SV / N bases with -alg nw, indel-rich with -alg ksw2) sized for a run under compute-sanitizer:
    compute-sanitizer --tool memcheck  python tests/devharness/sanitizer_case.py
    compute-sanitizer --tool racecheck python tests/devharness/sanitizer_case.py
Not collected by pytest (takes minutes under the sanitizer)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_util as pu

case = pu.make_case(seed=9, n_pairs=3000, genome_len=60000, sv=5.0, n_dup=20, tandem=10, n_rate=0.005)
ix = pu.build_index(case)
pu.assert_same(pu.cuda_results(case, ix), pu.oracle_results(case, ix))
case2 = pu.make_case(seed=6, n_pairs=1500, genome_len=50000, alg_ksw2=1, indel_rate=0.003)
ix2 = pu.build_index(case2)
pu.assert_same(pu.cuda_results(case2, ix2), pu.oracle_results(case2, ix2))
print("sanitizer cases: parity ok")
