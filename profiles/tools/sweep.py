import os, sys, time, numpy as np
sys.path.insert(0, '.')
import cellranger_b200 as cb
from cellranger_b200 import synth, synth_device
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
cfg = synth.preset("cfg2", n); tables = synth.make_tables(cfg, n)
gw = cb.GemWell(); libs = bench.setup_problem(gw, cfg, tables)
d = synth_device.generate_device(gw, tables, 0, n, "gex")
gw.add_reads_device(libs[0], n, cfg.r1_len, d.r1_seq, d.r1_qual, d.feature)
def run(tag, reps=3):
    for _ in range(2): gw.run()
    acc = {}
    for _ in range(reps):
        gw.run()
        for k, v in gw.phase_times().items(): acc[k] = acc.get(k, 0) + v / reps
    print(tag, " ".join(f"{k.split('.')[-1]}={v:.2f}" for k, v in acc.items()), f"total={sum(acc.values()):.2f}", flush=True)
    return gw.stats()
ref = None
for p1 in range(1):
    os.environ["CRGPU_P1_CFG"] = str(p1); os.environ["CRGPU_SORT_CFG"] = "0"
    st = run(f"P1_CFG={p1}")
    sig = (st["keys"], st["distinct_keys"], st["molecules"], st["nnz"])
    ref = ref or sig; assert sig == ref, (sig, ref)
os.environ["CRGPU_P1_CFG"] = "0"
for sc in [0,4]:
    os.environ["CRGPU_SORT_CFG"] = str(sc)
    st = run(f"SORT_CFG={sc}")
    sig = (st["keys"], st["distinct_keys"], st["molecules"], st["nnz"]); assert sig == ref, (sig, ref)
