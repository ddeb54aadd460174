#!/usr/bin/env python
"""bench.py — reads/s through barcode correction + UMI correction/dedup + counting on N B200s.

  python bench.py --gpus 1 --steps K --warmup W              (our arm: the CUDA path)
  python bench.py --impl reference --steps K --warmup W      (the CPU restatement of the reference)
  torchrun ... bench.py --gpus N ...                         (one rank per GPU, NCCL)

A step is one pass of the whole hot path (pass 1 → priors → pass 2 → sort → dedup → matrix) over one
batch of synthetic reads. N=1 runs BASELINE.json configs[1] (200 M 3' v3 reads vs the 6 794 880-entry
whitelist, 30 k genes); N>1 runs configs[2] (200 M reads per GPU, weak scaling, barcode-owner all-to-all).
One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "reads/s through BC+UMI correction & UMI count"
FULL_READS_PER_GPU = 200_000_000
# algorithmic HBM bytes per read of the whole path for 3' v3 keys (62 bits, 8 passes)
PATH_BYTES_PER_READ = {"cfg1": 206.6, "cfg2": 226.6, "cfg3": 226.6, "cfg4": 232.6, "cfg5": 244.6}


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 7:
                    continue
                try:
                    sm.append(float(p[0]))
                    mx.append(float(p[1]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def setup_problem(gw, cfg, tables):
    """Whitelist, library types and feature reference of the workload on one GemWell."""
    import cellranger_b200 as cb

    wl = gw.add_whitelist(cb.Whitelist.plain(tables.whitelist))
    chem = cb.ChemistryDef(cfg.name, 0, cfg.bc_len, cfg.bc_len, cfg.umi_len)
    libs = [gw.add_library(wl, chem)]
    if cfg.n_fb_features:
        wl2 = gw.add_whitelist(cb.Whitelist.trans(tables.trans, tables.whitelist))
        libs.append(gw.add_library(wl2, chem, feature_type=1, fb_offset=cfg.fb_offset, fb_length=cfg.fb_len))
    fr = cb.FeatureReference(cfg.n_genes)
    for i in range(cfg.n_fb_features):
        fr.add_feature_barcode(f"FB{i}", bytes(tables.fb_seqs[i]).decode(), 1, "5P" + "N" * cfg.fb_offset + "(BC)")
    gw.set_feature_reference(fr)
    return libs


def cpu_oracle_rate(cfg, tables, sample_reads: int, threads: int, steps: int = 1, warmup: int = 0, reads=None):
    """reads/s of the CPU restatement (oracle/) on a bounded sample of the same workload: the first `sample_reads`
    reads, handed in (`reads`, already generated) or generated here with the numpy generator."""
    from cellranger_b200 import synth
    from oracle import cro

    o = cro.Oracle()
    wl = o.add_whitelist(tables.whitelist)
    lib = o.add_library(wl, 0, cfg.bc_len, cfg.bc_len, cfg.umi_len)
    o.set_features(np.zeros(cfg.n_genes, dtype=np.int32))
    if reads is None:
        reads = synth.generate_reads(tables, 0, sample_reads, "gex")
    times = []
    for it in range(warmup + steps):
        o.reset_reads()
        o.add_reads(lib, reads["r1_seq"], reads["r1_qual"], reads["feature"])
        t0 = time.perf_counter()
        o.run(threads)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    st = o.stats()
    o.close()
    return sample_reads / float(np.mean(times)), float(np.mean(times)), st


def csc_fingerprint(m, n_molecules: int) -> dict:
    """Order-independent 64-bit fingerprint of a (part of a) count matrix: a wrapping sum over the entries of a
    mix of (barcode content rank, feature, count), plus one over the barcode index. A barcode-owner sharded run
    is a partition of the single-GPU run, so the per-rank values add up (mod 2^64) to the single-GPU ones."""
    from cellranger_b200.synth import splitmix64

    cols = np.repeat(m.barcode_rank.astype(np.uint64), np.diff(m.indptr))
    with np.errstate(over="ignore"):
        e = splitmix64((cols << np.uint64(32)) | m.indices.astype(np.uint64)) * \
            (np.uint64(2) * m.data.astype(np.uint64) + np.uint64(1))
        h_entries = int(e.sum(dtype=np.uint64))
        h_barcodes = int(splitmix64(m.barcode_rank.astype(np.uint64) ^ np.uint64(0xB5AD4ECEDA1CE2A9)).sum(dtype=np.uint64))
    return {"entries": h_entries, "barcode_index": h_barcodes, "nnz": int(m.data.shape[0]),
            "n_barcodes": int(m.barcode_rank.shape[0]), "umis": int(m.data.sum(dtype=np.int64)),
            "molecules": int(n_molecules)}


def combine_fingerprints(parts) -> dict:
    out = {k: 0 for k in parts[0]}
    for p in parts:
        for k, v in p.items():
            out[k] = (out[k] + v) & ((1 << 64) - 1)
    return {"csc_hash": f"{out['entries']:016x}{out['barcode_index']:016x}", "nnz": out["nnz"],
            "n_barcodes": out["n_barcodes"], "umis": out["umis"], "molecules": out["molecules"]}


class _DevBytes:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


def parity_at_scale(gw, lib, cfg, tables, reads_dev, n_reads, n_targets=200, check_molecules=True, threads=None):
    """Bit-exact check of single output values at the full bench size. ~n_targets barcodes (cells and ambient)
    are drawn; every read of the batch whose raw barcode is one of them, one of their 48 Hamming-1 neighbours,
    or holds a non-ACGT base - a superset of the reads any of them can receive - goes through the CPU oracle
    with the GPU's GLOBAL priors pushed in (the posterior of a read depends on the priors of all its whitelist
    neighbours, Posterior::correct_barcode, barcode/src/corrector.rs:125-162). The oracle's matrix columns and
    molecule rows of the drawn barcodes must equal the GPU's; the priors of the drawn barcodes are recounted
    from the raw reads. The reads are selected with torch ops on the device arrays (checker plumbing)."""
    import torch

    from cellranger_b200 import synth
    from oracle import cro

    t0 = time.perf_counter()
    dev = torch.device("cuda", gw.device)
    m = gw.count_matrix()
    rng = np.random.default_rng(20240607)
    cells = np.unique(np.asarray(tables.cell_rank, dtype=np.int64))
    cells = cells[np.isin(cells, m.barcode_rank)]
    pick_cells = cells[np.linspace(0, len(cells) - 1, min(n_targets // 2, len(cells))).astype(np.int64)]
    others = m.barcode_rank[~np.isin(m.barcode_rank, cells)].astype(np.int64)
    pick_amb = rng.choice(others, size=min(n_targets - len(pick_cells), len(others)), replace=False)
    targets = np.unique(np.concatenate([pick_cells, pick_amb]))
    t_seq = gw.barcode_seqs(targets.astype(np.uint32))
    t_packed = synth.pack_2bit(t_seq)
    L = cfg.bc_len
    near = [t_packed]
    for pos in range(L):
        sh = np.uint64(2 * (L - 1 - pos))
        for b in range(4):
            near.append((t_packed & ~(np.uint64(3) << sh)) | (np.uint64(b) << sh))
    near = torch.as_tensor(np.unique(np.concatenate(near)).astype(np.int64), device=dev)
    r1_len = cfg.r1_len
    seq = torch.as_tensor(_DevBytes(reads_dev.r1_seq, n_reads * r1_len, "|u1"), device=dev).view(n_reads, r1_len)
    qual = torch.as_tensor(_DevBytes(reads_dev.r1_qual, n_reads * r1_len, "|u1"), device=dev).view(n_reads, r1_len)
    feat = torch.as_tensor(_DevBytes(reads_dev.feature, n_reads, "<i4"), device=dev)
    lut = torch.full((256,), 4, dtype=torch.int64, device=dev)
    for ch, v in zip(b"ACGT", range(4)):
        lut[ch] = v
    shifts = (torch.arange(L - 1, -1, -1, device=dev, dtype=torch.int64) * 2)
    picked = []
    step = 8_000_000
    for lo in range(0, n_reads, step):
        codes = lut[seq[lo:lo + step, :L].long()]
        has_n = (codes == 4).any(dim=1)
        packed = ((codes & 3) << shifts).sum(dim=1)
        sel = torch.isin(packed, near) | has_n
        picked.append(sel.nonzero(as_tuple=False).flatten() + lo)
        del codes, packed, sel, has_n
    idx = torch.cat(picked)
    sub_seq = seq[idx].cpu().numpy()
    sub_qual = qual[idx].cpu().numpy()
    sub_feat = feat[idx].cpu().numpy().view(np.uint32)
    n_sub = int(idx.numel())
    del idx, picked
    torch.cuda.empty_cache()
    t_select = time.perf_counter() - t0

    threads = threads or (os.cpu_count() or 1)
    o = cro.Oracle()
    wl = o.add_whitelist(tables.whitelist)
    olib = o.add_library(wl, 0, cfg.bc_len, cfg.bc_len, cfg.umi_len)
    o.set_features(np.zeros(cfg.n_genes, dtype=np.int32))
    o.add_reads(olib, sub_seq, sub_qual, sub_feat)
    o.pass1(threads)
    # the priors of the drawn barcodes, recounted from the raw reads: every exact read of a target is in the subset
    prior_gpu = gw.prior(lib)
    prior_sub = o.counts(olib, 0, t_seq)
    priors_equal = bool(np.array_equal(prior_sub, prior_gpu[targets].astype(np.int64)))
    # ... then the GLOBAL priors (all 200 M reads) replace the subset's
    nz = np.flatnonzero(prior_gpu)
    o.prior_clear(olib)
    o.prior_add(olib, gw.barcode_seqs(nz.astype(np.uint32)), prior_gpu[nz].astype(np.int64))
    o.pass2(threads)
    o.count(threads)
    mo = o.matrix()
    col_o = {bytes(b): i for i, b in enumerate(mo["barcodes"])}
    col_g = np.searchsorted(m.barcode_rank, targets)
    cols_ok, cols_bad, entries = 0, [], 0
    for t, s_, cg in zip(targets, t_seq, col_g):
        in_g = cg < len(m.barcode_rank) and m.barcode_rank[cg] == t
        co = col_o.get(bytes(s_))
        if not in_g or co is None:
            ok = (not in_g) and co is None
        else:
            a0, a1 = m.indptr[cg], m.indptr[cg + 1]
            b0, b1 = mo["indptr"][co], mo["indptr"][co + 1]
            ok = np.array_equal(m.indices[a0:a1], mo["indices"][b0:b1]) and np.array_equal(m.data[a0:a1], mo["data"][b0:b1])
            entries += int(a1 - a0)
        cols_ok += bool(ok)
        if not ok:
            cols_bad.append(int(t))
    mol_equal = None
    n_mol_rows = 0
    if check_molecules:
        mg = gw.molecules()                       # (column, library, feature, umi, read_count)
        keep = np.isin(mg[:, 0], col_g.astype(np.uint32))
        mg = mg[keep]
        mg[:, 0] = targets[np.searchsorted(col_g, mg[:, 0])]   # column -> content rank
        ma = o.molecules()
        rank_of_col = np.full(len(mo["barcodes"]), -1, dtype=np.int64)
        for t, s_ in zip(targets, t_seq):
            co = col_o.get(bytes(s_))
            if co is not None:
                rank_of_col[co] = t
        ra = rank_of_col[ma[:, 0].astype(np.int64)]
        ma = ma[ra >= 0].astype(np.int64)
        ma[:, 0] = ra[ra >= 0]
        mg = mg.astype(np.int64)
        ka = np.lexsort((ma[:, 5], ma[:, 4], ma[:, 3], ma[:, 2], ma[:, 1], ma[:, 0]))
        kg = np.lexsort((mg[:, 5], mg[:, 4], mg[:, 3], mg[:, 2], mg[:, 1], mg[:, 0]))
        mol_equal = bool(ma.shape == mg.shape and np.array_equal(ma[ka], mg[kg]))
        n_mol_rows = int(mg.shape[0])
    o.close()
    return {"barcodes_checked": int(len(targets)), "cells": int(len(pick_cells)), "reads_through_oracle": n_sub,
            "columns_equal": cols_ok == len(targets), "columns_ok": cols_ok, "columns_bad": cols_bad[:8],
            "entries_checked": entries, "priors_equal": priors_equal, "molecule_rows_equal": mol_equal,
            "molecule_rows_checked": n_mol_rows, "seconds": round(time.perf_counter() - t0, 1),
            "select_seconds": round(t_select, 1)}


def run_reference(args):
    """The reference arm: the reference's algorithm for this path on the host cores. The Rust crates cannot
    be compiled in this image (no cargo/rustc), so this is the C++ port in oracle/ (kind = "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cellranger_b200 import synth

    name = "cfg2" if args.gpus == 1 else "cfg3"
    n_total = args.reads * args.gpus
    cfg = synth.preset(name, n_total)
    tables = synth.make_tables(cfg, n_total)
    threads = os.cpu_count() or 1
    # every step runs the whole sample again: bounded so that the driver's --steps 20 --warmup 5 stays within minutes
    sample = min(args.cpu_sample, 8_000_000)
    rate, sec, st = cpu_oracle_rate(cfg, tables, sample, threads, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "reads/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/f64", "data": "synthetic",
        "config": workload_config(name, cfg, args.reads, args.gpus),
        "cpu_baseline": {"value": rate, "unit": "reads/s", "cores": threads, "kind": "port",
                         "sample": f"first {sample} reads of the workload per step, all {threads} host threads"},
        "e2e": {"value": rate, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(name, cfg, reads_per_gpu, gpus):
    desc = {
        "cfg2": "200M synthetic 3' v3 reads (16bp BC + 12bp UMI) vs 3M-february-2018-size whitelist (6794880), 30k genes, 1xB200",
        "cfg3": "synthetic 3' v3 reads at NovaSeq S4-lane scale, 200M per GPU, sharded by barcode owner (fused NVLink key exchange, NCCL all-to-all fallback)",
    }.get(name, name)
    return {"workload": f"{name}: {desc}", "reads_per_gpu": reads_per_gpu, "total_reads": reads_per_gpu * gpus,
            "whitelist": cfg.n_whitelist, "genes": cfg.n_genes, "cells": cfg.n_cells, "bc_len": cfg.bc_len,
            "umi_len": cfg.umi_len, "bc_err": cfg.bc_err, "parallelism": f"read-sharded x{gpus}, barcode-owner exchange",
            "l2": "inputs (60 B/read) far larger than L2; no flush needed"}


def run_ours(args):
    import torch

    import cellranger_b200 as cb
    from cellranger_b200 import dist as crdist
    from cellranger_b200 import synth, synth_device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        dist = None
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    n_per = args.reads
    name = "cfg2" if world == 1 else "cfg3"
    n_total = n_per * world
    cfg = synth.preset(name, n_total)
    tables = synth.make_tables(cfg, n_total)
    gw = cb.GemWell(device=local_rank)
    libs = setup_problem(gw, cfg, tables)
    reads_dev = synth_device.generate_device(gw, tables, rank * n_per, n_per, "gex")
    ext = torch.cuda.ExternalStream(gw.stream(), device=dev)
    sharded = None
    exchange = "none (one GPU)"
    if world > 1:
        if os.environ.get("CRGPU_NO_P2P") or os.environ.get("CRGPU_TORCH_DIST"):
            # the engine protocol of dist.py with torch.distributed collectives (second implementation; the
            # NCCL all-to-all route when CRGPU_NO_P2P is set)
            engine = crdist.TorchEngine(gw, len(libs))
            if not os.environ.get("CRGPU_NO_P2P"):
                engine.setup_peer_exchange(rank, world, capacity_keys=int(n_per * 1.25))
            sharded = crdist.ShardedGemWell(engine, rank, world)
            exchange = "torch.distributed all-reduces + " + ("peer stores" if engine.p2p else "NCCL all_to_all_single")
        else:
            # the product path: the whole step inside libcrgpu.so (crgpu_sharded_run); torch only carries the
            # communicator id to the ranks and the timing reduction
            box = [cb.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            sharded = crdist.NativeShardedGemWell(gw, rank, world, box[0], capacity_keys=int(n_per * 1.25))
            exchange = "crgpu_sharded_run: in-library NCCL all-reduces, device owner ranges, NVLink peer stores" + \
                       (", early scatter" if os.environ.get("CRGPU_EARLY_SCATTER") else "")

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_device():
        if world == 1:
            gw.run()
        else:
            sharded.run()

    # ---------------- device-resident timing (value) ----------------
    gw.add_reads_device(libs[0], n_per, cfg.r1_len, reads_dev.r1_seq, reads_dev.r1_qual, reads_dev.feature)
    for _ in range(args.warmup):
        step_device()
    launches0 = gw.stats()["kernel_launches"]
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(ext)
    for _ in range(args.steps):
        step_device()
    ev1.record(ext)
    barrier()
    clocks = sampler.stop() if rank == 0 else {}
    ms = ev0.elapsed_time(ev1) / args.steps
    if dist is not None:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    stats = gw.stats()
    launches = (stats["kernel_launches"] - launches0) // max(args.steps, 1)
    value = n_total / (ms * 1e-3)

    if world > 1 and getattr(sharded, "timing", False) and rank == 0:
        n_runs = args.warmup + args.steps
        print("dist phases (ms/step, host clock, rank 0): " +
              " ".join(f"{k}={v / n_runs:.2f}" for k, v in sharded.times.items()), file=sys.stderr)
    # per-phase device times of one more step (phase_times() synchronises, so outside the timed loop)
    step_device()
    phases = gw.phase_times()
    if os.environ.get("CRGPU_BENCH_PHASES_EARLY") and rank == 0:  # timing experiments whose results are not valid
        print(json.dumps({"ms_per_step": ms, "phases_ms": phases}), file=sys.stderr, flush=True)
    n_keys, n_distinct = stats["keys"], stats["distinct_keys"]

    # ---------------- what was computed: fingerprint, and single values checked against the oracle ----------------
    def fingerprint():
        fp = csc_fingerprint(gw.count_matrix(), gw.stats()["molecules"])
        if dist is None:
            return combine_fingerprints([fp])
        parts = [None] * world
        dist.all_gather_object(parts, fp)
        return combine_fingerprints(parts)

    result_fp = fingerprint()
    scale_check = None
    if world == 1 and not args.no_scale_check:
        try:
            scale_check = parity_at_scale(gw, libs[0], cfg, tables, reads_dev, n_per, n_targets=args.scale_check_barcodes)
        except Exception as e:  # the check must never take the measurement down with it
            scale_check = {"error": f"{type(e).__name__}: {e}"}
    # Strong-scaling identity: the 200 M reads of cfg2 (the N=1 workload) split N ways give the matrix of N=1 -
    # equal fingerprints on the N=1, 2, 4, 8 lines (columns concatenate by owner range, as the reference's chunk
    # outputs do: barcode_correction.rs:252-262, align_and_count.rs:519-524)
    identity = None
    if world == 1:
        identity = dict(result_fp, workload=f"cfg2, {n_per} reads on 1 GPU (this run)")
    elif not args.no_identity:
        n_id = args.reads
        cfg2 = synth.preset("cfg2", n_id)
        tables2 = synth.make_tables(cfg2, n_id)
        lo, hi = (n_id * rank) // world, (n_id * (rank + 1)) // world
        gw.clear_reads()
        id_reads = synth_device.generate_device(gw, tables2, lo, hi - lo, "gex")
        gw.add_reads_device(libs[0], hi - lo, cfg2.r1_len, id_reads.r1_seq, id_reads.r1_qual, id_reads.feature)
        step_device()
        identity = dict(fingerprint(), workload=f"cfg2, {n_id} reads split over {world} GPUs")
        gw.clear_reads()
        id_reads.close()
        gw.add_reads_device(libs[0], n_per, cfg.r1_len, reads_dev.r1_seq, reads_dev.r1_qual, reads_dev.feature)

    # ---------------- end to end through the public API with host buffers (e2e) ----------------
    e2e = None
    cpu_reads = None
    if not args.no_e2e:
        gw.clear_reads()
        host = {}
        keep_ptrs = []
        for key, devp, shape, dt in (("r1_seq", reads_dev.r1_seq, (n_per, cfg.r1_len), np.uint8),
                                     ("r1_qual", reads_dev.r1_qual, (n_per, cfg.r1_len), np.uint8),
                                     ("feature", reads_dev.feature, (n_per,), np.uint32)):
            nbytes = int(np.prod(shape)) * np.dtype(dt).itemsize
            p = C.c_void_p()
            cb._lib.check(gw.L.crgpu_host_alloc_pinned(C.c_uint64(nbytes), C.byref(p)), "pinned alloc")
            keep_ptrs.append(p)
            buf = (C.c_uint8 * nbytes).from_address(p.value)
            arr = np.frombuffer(buf, dtype=dt).reshape(shape)
            cb._lib.check(gw.L.crgpu_memcpy_d2h(gw.ctx, C.c_void_p(p.value), C.c_void_p(devp), C.c_uint64(nbytes)))
            host[key] = arr
        reads_dev.close()
        h2d = sum(a.nbytes for a in host.values())
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        d2h = 0

        def step_e2e():
            nonlocal d2h
            gw.clear_reads()
            gw.add_reads(libs[0], host["r1_seq"], host["r1_qual"], host["feature"])
            step_device()
            m = gw.count_matrix(pinned=True)  # device→host read of the step's result
            d2h = m.indptr.nbytes + m.indices.nbytes + m.data.nbytes + m.barcode_rank.nbytes
            return m

        step_e2e()  # warm-up (allocations of the H2D staging buffers)
        # the ceiling of this number: the same pinned buffers copied to the device and nothing else, on every rank
        # at once (PCIe link per GPU; at several GPUs also the host's memory and root complexes)
        gw.clear_reads()
        torch.cuda.empty_cache()
        sinks = {k: torch.empty(a.nbytes, dtype=torch.uint8, device=dev) for k, a in host.items()}
        srcs = {k: torch.from_numpy(a.reshape(-1).view(np.uint8)) for k, a in host.items()}
        h2d_ms = []
        for _ in range(2):
            barrier()
            t0 = time.perf_counter()
            for k in sinks:
                sinks[k].copy_(srcs[k], non_blocking=True)
            torch.cuda.synchronize(dev)
            h2d_ms.append((time.perf_counter() - t0) * 1e3)
        h2d_only = min(h2d_ms)
        if dist is not None:
            t = torch.tensor([h2d_only], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            h2d_only = float(t.item())
        del sinks, srcs
        torch.cuda.empty_cache()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_e2e()
        barrier()
        sec = (time.perf_counter() - t0) / e2e_steps
        if dist is not None:
            t = torch.tensor([sec], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        e2e = {"value": n_total / sec, "unit": "reads/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "steps": e2e_steps, "ms_per_step": sec * 1e3,
               # the bare host-to-device copy of one step's inputs, all ranks at once, max over ranks: what PCIe and
               # the host allow; the step overlaps pass 1 with the copy, pass 2 / count / D2H come after it
               "h2d_only_ms": h2d_only, "h2d_gbs_per_gpu": h2d / h2d_only / 1e6,
               "frac_of_h2d_ceiling": h2d_only / (sec * 1e3)}
        if rank == 0 and world == 1 and not args.no_cpu:  # the CPU baseline's sample: the first reads of this very batch
            ns = min(args.cpu_sample, n_per)
            cpu_reads = {k: np.array(v[:ns]) for k, v in host.items()}
        for p in keep_ptrs:
            gw.L.crgpu_host_free_pinned(p)

    if rank != 0:
        gw.close()
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel ----------------
    peak, peak_src = measured_peak_hbm()
    kern = {}
    sweep_name = next((k for k in phases if k.startswith("count.sort.onesweep_x")), None)
    if sweep_name:
        n_pass = int(sweep_name.rsplit("x", 1)[1])
        per_launch_ms = phases[sweep_name] / n_pass
        kern["radix_onesweep_kernel"] = {"launches_per_step": n_pass, "ms_per_launch": per_launch_ms,
                                         "total_ms": phases[sweep_name], "alg_bytes_per_launch": n_keys * 16}
    if "pass1" in phases:
        kern["pass1_staged_kernel"] = {"launches_per_step": 1, "ms_per_launch": phases["pass1"],
                                       "total_ms": phases["pass1"],
                                       "alg_bytes_per_launch": n_per * (2 * cfg.r1_len + 4 + 8) + n_keys * 8}
    dom = max(kern, key=lambda k: kern[k]["total_ms"]) if kern else None
    roofline = None
    if dom:
        k = kern[dom]
        achieved = k["alg_bytes_per_launch"] / (k["ms_per_launch"] * 1e-3) / 1e9
        # dram__bytes_read+write per launch: ncu cannot run inside this process, so the figure comes from the
        # committed `ncu --set full` capture of this very workload (same key count, checked) - see traffic_source
        traffic, traffic_source = None, None
        try:
            rec = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
            if dom in rec and rec.get("_n_keys") == n_keys and world == 1:
                traffic = rec[dom]["dram_bytes_per_launch"]
                traffic_source = "profiles/r02_traffic.json <- " + rec.get("_source", "ncu --set full")
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_source,
                    "peak_source": peak_src,
                    "alg_bytes_per_launch": k["alg_bytes_per_launch"], "ms_per_launch": k["ms_per_launch"],
                    "launches_per_step": k["launches_per_step"],
                    "share_of_step": k["total_ms"] / max(sum(phases.values()), 1e-9)}
    path_frac = (value / world) * PATH_BYTES_PER_READ[name] / 1e9 / peak

    # ---------------- CPU baseline on a bounded sample ----------------
    cpu = None
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        ns = args.cpu_sample if cpu_reads is not None else min(args.cpu_sample, 2_000_000)
        rate, sec, _ = cpu_oracle_rate(cfg, tables, ns, threads, reads=cpu_reads)
        cpu = {"value": rate, "unit": "reads/s", "cores": threads, "kind": "port",
               "sample": f"first {ns} reads of the workload, {sec:.1f} s on {threads} host threads",
               "note": "C++ restatement of the reference's hash-map algorithm (oracle/cr_oracle.cpp: inline "
                       "fixed-capacity sequences and open-addressing maps like the reference's SSeqGen / hashbrown), "
                       "not the Rust crates (no cargo/rustc in this image): a reported baseline only"}

    line = {
        "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 keys / f64 posterior", "data": "synthetic", "config": workload_config(name, cfg, n_per, world),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        "phases_ms": phases, "path_hbm_frac": path_frac,
        "counts": {k: stats[k] for k in ("reads", "valid_before", "corrected", "invalid", "keys", "distinct_keys",
                                         "molecules", "nnz", "barcodes")},
        "result": result_fp, "parity_at_scale": scale_check, "strong_identity": identity, "exchange": exchange,
    }
    print(json.dumps(line))
    gw.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=FULL_READS_PER_GPU, help="reads per GPU (default: the full config)")
    ap.add_argument("--cpu-sample", type=int, default=32_000_000,
                    help="reads of the CPU baseline's sample (about 10-30 s of CPU work at the default)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-scale-check", action="store_true", help="skip the sampled-barcode oracle check at N=1")
    ap.add_argument("--scale-check-barcodes", type=int, default=200)
    ap.add_argument("--no-identity", action="store_true", help="skip the strong-scaling identity run at N>1")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
