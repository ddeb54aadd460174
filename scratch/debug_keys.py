import sys, os, ctypes as C, numpy as np
sys.path.insert(0, '.')
from tests import helpers
import cellranger_b200 as cb
from cellranger_b200._lib import check, ptr
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
prob = helpers.make_problem("cfg1", n)
gw = helpers.run_gpu(prob, annotate=False, run=False)
kl = gw.key_layout(); print(kl)
runs = []
for it in range(iters):
    gw.make_shard(); gw.barcode_correction()
    p, nk = gw.keys_dev()
    k = np.zeros(nk, dtype=np.uint64); check(gw.L.crgpu_memcpy_d2h(gw.ctx, ptr(k), C.c_void_p(p), C.c_uint64(nk*8)))
    runs.append(np.sort(k))
# expected keys from per-read outputs of the last run
r = gw.reads(0)
ok = ((r["state"] == 1) | (r["state"] == 2)) & ((r["flags"] & 1) != 0) & (r["feature"] != 0xFFFFFFFF)
exp = (r["bc_rank"][ok].astype(np.uint64) << np.uint64(kl["rank_shift"])) | (r["feature"][ok].astype(np.uint64) << np.uint64(kl["feature_shift"])) | r["umi"][ok].astype(np.uint64)
exp = np.sort(exp)
print("expected keys", len(exp))
for it, k in enumerate(runs):
    same = len(k) == len(exp) and np.array_equal(k, exp)
    if not same:
        a = np.setdiff1d(k, exp); b = np.setdiff1d(exp, k)
        print(it, "n", len(k), "extra", len(a), "missing", len(b))
        for x in a[:6]:
            print("   extra  rank", int(x >> np.uint64(kl["rank_shift"])), "feat", int((x >> np.uint64(kl["feature_shift"])) & np.uint64(0x7FFF)), "umi", hex(int(x & np.uint64((1 << kl["umi_bits"]) - 1))))
        for x in b[:6]:
            print("   missing rank", int(x >> np.uint64(kl["rank_shift"])), "feat", int((x >> np.uint64(kl["feature_shift"])) & np.uint64(0x7FFF)), "umi", hex(int(x & np.uint64((1 << kl["umi_bits"]) - 1))))
    else:
        print(it, "ok")
