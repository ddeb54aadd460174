"""Experiment: how much of pass 2's time is the order of the invalid-read side list?

The side list follows the read order. Here the INPUT reads are reordered on the device (torch sort on the first
k barcode bases, k = 0, 1, 2, 4, 8) before the step runs, so that pass 2 meets its reads grouped by barcode
prefix - the upper bound of what binning the side list inside the library could buy. Prints phase times per k.
"""
import sys

sys.path.insert(0, '.')
import torch

import bench
import cellranger_b200 as cb
from cellranger_b200 import synth, synth_device


class _Dev:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}


n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
ks = [int(a) for a in sys.argv[2:]] or [0, 1, 2, 4, 8]
cfg = synth.preset("cfg2", n)
tables = synth.make_tables(cfg, n)
gw = cb.GemWell()
libs = bench.setup_problem(gw, cfg, tables)
d = synth_device.generate_device(gw, tables, 0, n, "gex")
gw.sync()
seq = torch.as_tensor(_Dev(d.r1_seq, (n, cfg.r1_len), "|u1"), device="cuda")
qual = torch.as_tensor(_Dev(d.r1_qual, (n, cfg.r1_len), "|u1"), device="cuda")
feat = torch.as_tensor(_Dev(d.feature, (n,), "<i4"), device="cuda")
gw.add_reads_device(libs[0], n, cfg.r1_len, d.r1_seq, d.r1_qual, d.feature)
for k in ks:
    if k:
        key = torch.zeros(n, dtype=torch.int64, device="cuda")
        for j in range(k):
            key = key * 256 + seq[:, j].to(torch.int64)
        perm = torch.argsort(key)
        del key
        for t in (seq, qual, feat):
            t.copy_(t[perm])
        del perm
        torch.cuda.synchronize()
    for _ in range(2):
        gw.run()
    acc, reps = {}, 3
    for _ in range(reps):
        gw.run()
        for kk, v in gw.phase_times().items():
            acc[kk] = acc.get(kk, 0) + v / reps
    st = gw.stats()
    print(f"sorted_on_first_{k}_bases", " ".join(f"{kk.split('.')[-1]}={v:.2f}" for kk, v in acc.items()),
          f"total={sum(acc.values()):.2f}", f"nnz={st['nnz']} molecules={st['molecules']}", flush=True)
