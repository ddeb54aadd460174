"""Phase times of the 200 M-read cfg2 step under different build-time / environment configurations."""
import os, sys
sys.path.insert(0, '.')
import cellranger_b200 as cb
from cellranger_b200 import synth, synth_device
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
envs = [dict(kv.split("=") for kv in arg.split(",")) for arg in sys.argv[2:]] or [{}]
cfg = synth.preset("cfg2", n); tables = synth.make_tables(cfg, n)
gw = cb.GemWell(); libs = bench.setup_problem(gw, cfg, tables)
d = synth_device.generate_device(gw, tables, 0, n, "gex")
gw.add_reads_device(libs[0], n, cfg.r1_len, d.r1_seq, d.r1_qual, d.feature)
ref = None
for env in envs:
    for k, v in env.items(): os.environ[k] = v
    for _ in range(2): gw.run()
    acc = {}
    reps = 3
    for _ in range(reps):
        gw.run()
        for k, v in gw.phase_times().items(): acc[k] = acc.get(k, 0) + v / reps
    st = gw.stats()
    sig = (st["keys"], st["distinct_keys"], st["molecules"], st["nnz"], st["umi_corrected_keys"])
    ref = ref or sig
    print(env, " ".join(f"{k.split('.')[-1]}={v:.2f}" for k, v in acc.items()), f"total={sum(acc.values()):.2f}",
          "" if sig == ref else f"RESULT DIFFERS {sig} vs {ref}", flush=True)
    for k in env: os.environ.pop(k, None)
