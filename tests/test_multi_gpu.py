"""Several real GPUs: the barcode-owner sharded path against a single-process oracle run, at world sizes 2, 4 and
8 (each skipped when the box has fewer GPUs; run with `gpurun --gpus 8 -- python -m pytest tests/test_multi_gpu.py
-m gpu`). Four routes must give the same matrix:

  native        crgpu_comm_init + crgpu_sharded_run: the whole step inside libcrgpu.so (NCCL all-reduces, device
                owner ranges, peer stores, on-stream barrier) - the product path
  native-early  the same with CRGPU_EARLY_SCATTER=1 (keys of pass 1 travel while pass 2 runs)
  torch-peer    the engine protocol of cellranger_b200/dist.py with torch.distributed collectives + peer stores
  torch-a2a     ... + NCCL all_to_all_single instead of peer stores
"""
import os
import socket
import tempfile

import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, outdir, name, n_reads, kw, mode, uid):
    import torch

    import cellranger_b200 as cb
    from cellranger_b200 import dist as crdist

    torch.cuda.set_device(rank)
    use_torch = mode.startswith("torch")
    if use_torch:
        import torch.distributed as dist

        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                                device_id=torch.device("cuda", rank))
    if mode == "native-early":
        os.environ["CRGPU_EARLY_SCATTER"] = "1"
    try:
        prob = helpers.make_problem(name, n_reads, **kw)
        cfg, t = prob["cfg"], prob["tables"]
        per = n_reads // world
        lo, hi = rank * per, (n_reads if rank == world - 1 else (rank + 1) * per)
        gw = cb.GemWell(device=rank)
        wl = gw.add_whitelist(cb.Whitelist.plain(t.whitelist))
        lib = gw.add_library(wl, cb.ChemistryDef(cfg.name, 0, cfg.bc_len, cfg.bc_len, cfg.umi_len))
        gw.set_feature_reference(cb.FeatureReference(cfg.n_genes))
        g = prob["gex"]
        gw.add_reads(lib, g["r1_seq"][lo:hi], g["r1_qual"][lo:hi], g["feature"][lo:hi])
        if use_torch:
            eng = crdist.TorchEngine(gw, 1)
            if mode == "torch-peer":
                assert eng.setup_peer_exchange(rank, world, capacity_keys=n_reads), "no peer access between the GPUs"
            sh = crdist.ShardedGemWell(eng, rank, world, early_scatter=False)
        else:
            sh = crdist.NativeShardedGemWell(gw, rank, world, uid, capacity_keys=n_reads)
        for _ in range(2):  # twice: the second run reuses every buffer
            sh.run()
        m = gw.count_matrix()
        st = gw.stats()
        np.savez(os.path.join(outdir, f"rank{rank}.npz"), barcodes=m.barcodes, indptr=m.indptr, indices=m.indices,
                 data=m.data, bounds=np.asarray(sh.bounds).astype(np.int64), rank_ids=m.barcode_rank,
                 states=np.array([st["valid_before"], st["corrected"], st["invalid"]]),
                 exchange=np.array([sh.exchange_bytes]))
        gw.close()
    finally:
        if use_torch:
            dist.destroy_process_group()


CASES = [(2, "native"), (2, "native-early"), (2, "torch-peer"), (2, "torch-a2a"),
         (4, "native"), (4, "native-early"), (8, "native"), (8, "native-early"), (8, "torch-peer")]


@pytest.mark.parametrize("world,mode", CASES, ids=[f"{w}gpu-{m}" for w, m in CASES])
@pytest.mark.parametrize("name,n,kw", [("cfg1", 400_000, {}), ("cfg2", 300_000, {"n_whitelist": 300_000, "n_cells": 300})],
                         ids=["cfg1", "cfg2"])
def test_sharded_matches_oracle(name, n, kw, world, mode):
    import torch
    import torch.multiprocessing as mp

    import cellranger_b200 as cb

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    uid = cb.comm_unique_id() if mode.startswith("native") else b""
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), d, name, n, kw, mode, uid), nprocs=world, join=True)
        parts = [dict(np.load(os.path.join(d, f"rank{r}.npz"))) for r in range(world)]
    prob = helpers.make_problem(name, n, **kw)
    o = helpers.run_oracle(prob, threads=8)
    mo = o.matrix()
    for p in parts[1:]:
        assert np.array_equal(parts[0]["bounds"], p["bounds"]), "every rank holds the same owner ranges"
    barcodes = np.concatenate([p["barcodes"] for p in parts])
    assert np.array_equal(mo["barcodes"], barcodes), "barcode index = concatenation of the owners' column blocks"
    indices = np.concatenate([p["indices"] for p in parts])
    data = np.concatenate([p["data"] for p in parts])
    indptr, off = [np.zeros(1, dtype=np.int64)], 0
    for p in parts:
        indptr.append(p["indptr"][1:] + off)
        off += int(p["indptr"][-1])
    assert np.array_equal(mo["indptr"], np.concatenate(indptr))
    assert np.array_equal(mo["indices"], indices)
    assert np.array_equal(mo["data"], data)
    so = o.stats()
    states = sum(p["states"] for p in parts)
    assert states.tolist() == [so["valid_before"], so["corrected"], so["invalid"]]
    assert all(int(p["exchange"][0]) > 0 for p in parts)
    for r, p in enumerate(parts):
        if len(p["rank_ids"]):
            assert p["rank_ids"].min() >= p["bounds"][r] and p["rank_ids"].max() < p["bounds"][r + 1]
    # the owner ranges of the device scan are the ones the host formula gives for the oracle's valid-read counts
    from cellranger_b200.dist import owner_bounds

    wl = prob["tables"].whitelist
    valid = o.counts(0, 0, wl) + o.counts(0, 1, wl)
    if mode != "native-early":  # the early variant balances on the valid-before counts alone
        assert np.array_equal(owner_bounds(torch.as_tensor(valid), world).astype(np.int64), parts[0]["bounds"])
    o.close()


def test_device_owner_bounds_match_host_formula():
    """crgpu_owner_bounds_compute (the device scan of the sharded run) against owner_bounds() of dist.py, the
    analogue of ShardReader::make_chunks' equal-read barcode ranges: empty vectors, zeros, one heavy barcode,
    sizes around the scan's chunk size, every world size."""
    import torch

    import cellranger_b200 as cb
    from cellranger_b200.dist import owner_bounds

    rng = np.random.default_rng(11)
    gw = cb.GemWell()
    vectors = [np.zeros(1000, dtype=np.uint32), np.ones(1, dtype=np.uint32), np.arange(1, 70_000, dtype=np.uint32),
               rng.integers(0, 5, size=32768).astype(np.uint32), rng.integers(0, 50, size=32769).astype(np.uint32),
               (rng.random(200_003) < 0.01).astype(np.uint32) * rng.integers(1, 40_000, size=200_003).astype(np.uint32),
               rng.integers(0, 3, size=737_280).astype(np.uint32)]
    heavy = np.zeros(100_000, dtype=np.uint32)
    heavy[77_777] = 4_000_000_000
    heavy[5] = 3
    vectors.append(heavy)
    for v in vectors:
        for g in (1, 2, 3, 4, 7, 8, 16):
            exp = owner_bounds(torch.as_tensor(v.astype(np.int64)), g)
            got = gw.owner_bounds_compute(v, g)
            assert np.array_equal(exp, got), (v.shape, g, exp, got)
    gw.close()
