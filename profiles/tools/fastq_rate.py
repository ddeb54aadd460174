"""Throughput of the FASTQ front end (crgpu_fastq_extract) and of the BarcodeSummary reduction
(crgpu_barcode_summary) on one GPU. Prints one JSON line."""
import ctypes as C, json, sys, time
sys.path.insert(0, '.')
import numpy as np
import torch
import cellranger_b200 as cb
from cellranger_b200 import synth, synth_device
from cellranger_b200._lib import check, ptr
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
rl = 28
rng = np.random.default_rng(1)
# fixed-width records so that the text can be built with numpy: "@" + 39 header bytes, seq, "+", qual
head = np.frombuffer(b"@A00123:45:HXXXXXXXX:1:1101:00000:00000 1", dtype=np.uint8)
rec_len = len(head) + 1 + rl + 1 + 1 + 1 + rl + 1
text = np.empty((n, rec_len), dtype=np.uint8)
text[:, :len(head)] = head
text[:, len(head)] = 10
seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(n, rl))]
o = len(head) + 1
text[:, o:o + rl] = seq
text[:, o + rl] = 10
text[:, o + rl + 1] = ord("+")
text[:, o + rl + 2] = 10
text[:, o + rl + 3:o + rl + 3 + rl] = ord("I")
text[:, -1] = 10
text = text.reshape(-1)
gw = cb.GemWell()
ext = torch.cuda.ExternalStream(gw.stream())
def dalloc(nb):
    p = C.c_void_p(); check(gw.L.crgpu_dev_alloc(gw.ctx, C.c_uint64(nb), C.byref(p))); return p
d_text, d_seq, d_qual = dalloc(text.nbytes + 16), dalloc(n * rl + 16), dalloc(n * rl + 16)
check(gw.L.crgpu_memcpy_h2d(gw.ctx, d_text, ptr(text), C.c_uint64(text.nbytes)))
nr, ns, nb = C.c_uint64(), C.c_uint64(), C.c_uint64()
def run():
    check(gw.L.crgpu_fastq_extract(gw.ctx, d_text, C.c_uint64(text.nbytes), 1, rl, d_seq, d_qual, C.c_uint64(n),
                                   C.byref(nr), C.byref(ns), C.byref(nb)), "fastq")
for _ in range(3): run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(ext)
for _ in range(10): run()
e1.record(ext); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
got = gw.read_device(d_seq.value, (n, rl))
assert nr.value == n and ns.value == 0 and nb.value == 0 and np.array_equal(got, seq)
out = {"fastq_extract": {"records": n, "text_bytes": int(text.nbytes), "ms": ms, "records_per_s": n / ms * 1e3,
                          "text_GB_per_s": text.nbytes / ms / 1e6, "note": "includes the D2H of three counters and a stream sync per call"}}
gw.close()
# BarcodeSummary on the cfg2 step
N = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000_000
cfg = synth.preset("cfg2", N); tables = synth.make_tables(cfg, N)
gw = cb.GemWell(); libs = bench.setup_problem(gw, cfg, tables)
d = synth_device.generate_device(gw, tables, 0, N, "gex")
gw.add_reads_device(libs[0], N, cfg.r1_len, d.r1_seq, d.r1_qual, d.feature)
gw.run()
t0 = time.perf_counter(); s = gw.barcode_summary(0); t1 = time.perf_counter()
t0 = time.perf_counter(); s = gw.barcode_summary(0); t1 = time.perf_counter()
st = gw.stats()
assert int(s["umis"].sum()) == st["molecules"] and int(s["umi_corrected_reads"].sum()) == st["umi_corrected_reads"]
assert int(s["reads"].sum()) == st["valid_before"] + st["corrected"]
out["barcode_summary"] = {"reads": N, "rows": int(len(s)), "ms_incl_d2h_and_host_filter": (t1 - t0) * 1e3}
print(json.dumps(out))
