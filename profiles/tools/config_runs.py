"""The five BASELINE.json configurations on one GPU: phase times at full size, parity against the CPU oracle
at a size the oracle finishes in seconds. Prints one JSON line per configuration."""
import json, sys, time
sys.path.insert(0, '.')
import numpy as np
import cellranger_b200 as cb
from cellranger_b200 import synth, synth_device
from tests import helpers
import bench

def full_size(name, n):
    cfg = synth.preset(name, n); t = synth.make_tables(cfg, n)
    gw = cb.GemWell(); libs = bench.setup_problem(gw, cfg, t)
    n_fb = int(round(n * cfg.fb_frac)) if cfg.n_fb_features else 0
    d = synth_device.generate_device(gw, t, 0, n - n_fb, "gex")
    gw.add_reads_device(libs[0], n - n_fb, cfg.r1_len, d.r1_seq, d.r1_qual, d.feature)
    if n_fb:
        f = synth_device.generate_device(gw, t, 0, n_fb, "fb")
        gw.add_reads_device(libs[1], n_fb, cfg.r1_len, f.r1_seq, f.r1_qual, 0, cfg.r2_len, f.r2_seq, f.r2_qual)
    for _ in range(3): gw.run()
    acc = {}
    for _ in range(3):
        gw.run()
        for k, v in gw.phase_times().items(): acc[k] = acc.get(k, 0) + v / 3
    st = gw.stats(); m = gw.count_matrix()
    assert st["valid_before"] + st["corrected"] + st["invalid"] == n
    assert int(m.data.sum()) == st["molecules"] and np.all(np.diff(m.barcode_rank.astype(np.int64)) > 0)
    ms = sum(acc.values())
    gw.close()
    return dict(config=name, reads=n, ms_per_step=round(ms, 2), reads_per_s=n / ms * 1e3, phases_ms={k: round(v, 2) for k, v in acc.items()},
                counts={k: st[k] for k in ("valid_before", "corrected", "invalid", "keys", "distinct_keys", "umi_corrected_keys", "low_support_keys", "molecules", "nnz", "barcodes")})

def parity(name, n, **kw):
    prob = helpers.make_problem(name, n, **kw)
    t0 = time.time(); o = helpers.run_oracle(prob, threads=16); t_cpu = time.time() - t0
    gw = helpers.run_gpu(prob)
    info = helpers.compare_all(o, gw, prob)
    gw.close()
    return dict(config=name, parity_reads=n, parity="bit-exact", oracle_s=round(t_cpu, 1), nnz=info["nnz"])

sizes = {"cfg1": 1_000_000, "cfg2": 200_000_000, "cfg4": 200_000_000, "cfg5": 200_000_000}
if __name__ == "__main__":
    # --full-only: skip the 1 M-read parity run (tests/test_gpu_parity.py::test_full_path_matches_oracle_1m_full_whitelist
    # holds it since round 2)
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    for name in args or ["cfg1", "cfg2", "cfg4", "cfg5"]:
        if "--full-only" not in sys.argv:
            print(json.dumps(parity(name, 1_000_000)), flush=True)
        print(json.dumps(full_size(name, sizes[name])), flush=True)
