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


def cpu_oracle_rate(cfg, tables, sample_reads: int, threads: int, steps: int = 1, warmup: int = 0):
    """reads/s of the CPU restatement (oracle/) on a bounded sample of the same workload."""
    from cellranger_b200 import synth
    from oracle import cro

    o = cro.Oracle()
    wl = o.add_whitelist(tables.whitelist)
    lib = o.add_library(wl, 0, cfg.bc_len, cfg.bc_len, cfg.umi_len)
    o.set_features(np.zeros(cfg.n_genes, dtype=np.int32))
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
    sample = args.cpu_sample
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
    engine = crdist.TorchEngine(gw, len(libs))
    if world > 1 and not os.environ.get("CRGPU_NO_P2P"):
        engine.setup_peer_exchange(rank, world, capacity_keys=int(n_per * 1.25))
    sharded = crdist.ShardedGemWell(engine, rank, world)

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

    if world > 1 and sharded.timing and rank == 0:
        n_runs = args.warmup + args.steps
        print("dist phases (ms/step, host clock, rank 0): " +
              " ".join(f"{k}={v / n_runs:.2f}" for k, v in sharded.times.items()), file=sys.stderr)
    # per-phase device times of one more step (phase_times() synchronises, so outside the timed loop)
    step_device()
    phases = gw.phase_times()
    n_keys, n_distinct = stats["keys"], stats["distinct_keys"]

    # ---------------- end to end through the public API with host buffers (e2e) ----------------
    e2e = None
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
               "steps": e2e_steps, "ms_per_step": sec * 1e3}
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
        # dram__bytes_read+write per launch from the committed ncu --set full capture of the same workload
        traffic = None
        try:
            rec = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
            if dom in rec and rec.get("_n_keys") == n_keys and world == 1:
                traffic = rec[dom]["dram_bytes_per_launch"]
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "alg_bytes_per_launch": k["alg_bytes_per_launch"], "ms_per_launch": k["ms_per_launch"],
                    "launches_per_step": k["launches_per_step"],
                    "share_of_step": k["total_ms"] / max(sum(phases.values()), 1e-9)}
    path_frac = (value / world) * PATH_BYTES_PER_READ[name] / 1e9 / peak

    # ---------------- CPU baseline on a bounded sample ----------------
    cpu = None
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        rate, sec, _ = cpu_oracle_rate(cfg, tables, args.cpu_sample, threads)
        cpu = {"value": rate, "unit": "reads/s", "cores": threads, "kind": "port",
               "sample": f"first {args.cpu_sample} reads of the workload, {sec:.1f} s on {threads} host threads"}

    line = {
        "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 keys / f64 posterior", "data": "synthetic", "config": workload_config(name, cfg, n_per, world),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        "phases_ms": phases, "path_hbm_frac": path_frac,
        "counts": {k: stats[k] for k in ("reads", "valid_before", "corrected", "invalid", "keys", "distinct_keys",
                                         "molecules", "nnz", "barcodes")},
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
    ap.add_argument("--cpu-sample", type=int, default=2_000_000)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
