import sys, ctypes as C, numpy as np
sys.path.insert(0, '.')
from tests import helpers
import cellranger_b200 as cb
from cellranger_b200._lib import check, ptr
prob = helpers.make_problem("cfg1", 2_000_000)
gw = helpers.run_gpu(prob, annotate=True)
st = gw.stats(); print(st)
r = gw.reads(0)
print("rep reads", int(((r["flags"] & 16) != 0).sum()), "molecules", st["molecules"])
# sort check
gw2 = helpers.run_gpu(prob, annotate=False, run=False)
gw2.make_shard(); gw2.barcode_correction()
p, n = gw2.keys_dev()
before = np.zeros(n, dtype=np.uint64); check(gw2.L.crgpu_memcpy_d2h(gw2.ctx, ptr(before), C.c_void_p(p), C.c_uint64(n*8)))
cnt = gw2.keys_partition(np.array([0, gw2.n_content()], dtype=np.uint32))
p, n2 = gw2.keys_dev()
after = np.zeros(n2, dtype=np.uint64); check(gw2.L.crgpu_memcpy_d2h(gw2.ctx, ptr(after), C.c_void_p(p), C.c_uint64(n2*8)))
ref = np.sort(before)
print("n", n, n2, "sorted ok:", bool(np.array_equal(ref, after)), "monotone:", bool(np.all(after[1:] >= after[:-1])))
if not np.array_equal(ref, after):
    bad = np.nonzero(ref != after)[0]; print("first bad", bad[:10], len(bad))
o = helpers.run_oracle(prob, threads=8)
try:
    print(helpers.compare_all(o, gw, prob))
except AssertionError as e:
    print("MISMATCH:", e)
print("distinct keys host", len(np.unique(before)), "gpu", st["distinct_keys"])
ro = o.reads(); rg = gw.reads(0)
d = ro["flags"] != rg["flags"]
print("flag diffs", int(d.sum()))
idx = np.nonzero(d)[0][:12]
for i in idx:
    print(i, "oracle", ro["flags"][i], "gpu", rg["flags"][i], bytes(ro["umi"][i]), bytes(cb.unpack_2bit(rg["umi"][i:i+1], 10)[0]), "feat", ro["feature"][i], "state", ro["state"][i], bytes(ro["bc"][i]))
mo, mg = o.matrix(), gw.count_matrix()
print("matrix equal:", np.array_equal(mo["indptr"], mg.indptr), np.array_equal(mo["indices"], mg.indices), np.array_equal(mo["data"], mg.data))
so = o.stats(); print(so)
# look at one differing read's segment
if len(idx):
    i = idx[0]
    same = np.nonzero((ro["feature"] == ro["feature"][i]) & (ro["state"] != 3) & np.all(ro["bc"] == ro["bc"][i], axis=1) & ((ro["flags"] & 2) != 0))[0]
    print("segment reads", len(same))
    import collections
    raw = prob["gex"]["r1_seq"][same][:, 16:26]
    c = collections.Counter(bytes(x) for x in raw)
    print("distinct raw umis in segment", len(c))
