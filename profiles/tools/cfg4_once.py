import sys
sys.path.insert(0, '.')
import cellranger_b200 as cb
from cellranger_b200 import synth, synth_device
import bench
n = 50_000_000
cfg = synth.preset("cfg4", n); t = synth.make_tables(cfg, n)
gw = cb.GemWell(); libs = bench.setup_problem(gw, cfg, t)
n_fb = int(round(n * cfg.fb_frac))
d = synth_device.generate_device(gw, t, 0, n - n_fb, "gex")
gw.add_reads_device(libs[0], n - n_fb, cfg.r1_len, d.r1_seq, d.r1_qual, d.feature)
f = synth_device.generate_device(gw, t, 0, n_fb, "fb")
gw.add_reads_device(libs[1], n_fb, cfg.r1_len, f.r1_seq, f.r1_qual, 0, cfg.r2_len, f.r2_seq, f.r2_qual)
gw.run(); gw.run()
print(gw.phase_times())
