"""Two real GPUs, NCCL: the barcode-owner sharded path against a single-process oracle run. Needs >= 2 CUDA
devices (skipped otherwise; run with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`)."""
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


def _worker(rank, world, port, outdir, name, n_reads, kw, p2p):
    import torch
    import torch.distributed as dist

    import cellranger_b200 as cb
    from cellranger_b200.dist import ShardedGemWell, TorchEngine

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
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
        eng = TorchEngine(gw, 1)
        if p2p:
            assert eng.setup_peer_exchange(rank, world, capacity_keys=n_reads), "no peer access between the GPUs"
        sh = ShardedGemWell(eng, rank, world, early_scatter=(p2p == "early"))
        for _ in range(2):  # twice: the second run reuses every buffer
            sh.run()
        m = gw.count_matrix()
        st = gw.stats()
        np.savez(os.path.join(outdir, f"rank{rank}.npz"), barcodes=m.barcodes, indptr=m.indptr, indices=m.indices,
                 data=m.data, bounds=sh.bounds.astype(np.int64), rank_ids=m.barcode_rank,
                 states=np.array([st["valid_before"], st["corrected"], st["invalid"]]),
                 exchange=np.array([sh.exchange_bytes]))
        gw.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("p2p", [True, "early", False], ids=["peer-stores", "peer-stores-early", "nccl-all-to-all"])
@pytest.mark.parametrize("name,n,kw", [("cfg1", 400_000, {}), ("cfg2", 300_000, {"n_whitelist": 300_000, "n_cells": 300})])
def test_two_gpu_sharded_matches_oracle(name, n, kw, p2p):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), d, name, n, kw, p2p), nprocs=world, join=True)
        parts = [dict(np.load(os.path.join(d, f"rank{r}.npz"))) for r in range(world)]
    prob = helpers.make_problem(name, n, **kw)
    o = helpers.run_oracle(prob, threads=8)
    mo = o.matrix()
    assert np.array_equal(parts[0]["bounds"], parts[1]["bounds"])
    barcodes = np.concatenate([p["barcodes"] for p in parts])
    assert np.array_equal(mo["barcodes"], barcodes), "barcode index = concatenation of the owners' column blocks"
    indices = np.concatenate([p["indices"] for p in parts])
    data = np.concatenate([p["data"] for p in parts])
    indptr = np.concatenate([parts[0]["indptr"], parts[1]["indptr"][1:] + parts[0]["indptr"][-1]])
    assert np.array_equal(mo["indptr"], indptr)
    assert np.array_equal(mo["indices"], indices)
    assert np.array_equal(mo["data"], data)
    so = o.stats()
    states = parts[0]["states"] + parts[1]["states"]
    assert states.tolist() == [so["valid_before"], so["corrected"], so["invalid"]]
    assert all(int(p["exchange"][0]) > 0 for p in parts)
    for r, p in enumerate(parts):
        if len(p["rank_ids"]):
            assert p["rank_ids"].min() >= p["bounds"][r] and p["rank_ids"].max() < p["bounds"][r + 1]
