import sys, os, ctypes as C, numpy as np
os.environ["CRGPU_VERIFY"] = "1"
sys.path.insert(0, '.')
from tests import helpers
import cellranger_b200 as cb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
prob = helpers.make_problem("cfg1", n)
gw = helpers.run_gpu(prob, annotate=False, run=False)
ref = None
for it in range(iters):
    gw.run()
    st = gw.stats()
    sig = (st["keys"], st["distinct_keys"], st["molecules"], st["nnz"], st["umi_corrected_keys"], st["low_support_keys"])
    if ref is None: ref = sig
    flag = "" if sig == ref and st["sort_violations"] == 0 and st["rle_violations"] == 0 else "  <<<<<< DIFF"
    print(it, sig, "sortviol", st["sort_violations"], "rleviol", st["rle_violations"], flag, flush=True)
