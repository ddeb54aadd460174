"""world_size-2 gloo test of the barcode-owner sharding host logic (cellranger_b200/dist.py) on CPU.

The compute engine is a stand-in defined here: pass 1 / pass 2 by the C++ oracle on the rank's own reads
with the all-reduced priors pushed back in, dedup + counting on the exchanged keys by the Python
restatement. What is under test is ShardedGemWell: the all-reduces, the owner ranges, the split sizes and
the all-to-all, checked by comparing the concatenation of the two ranks' column blocks with a
single-process oracle run over all the reads."""
import os
import socket
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cellranger_b200 import synth
from cellranger_b200.dist import ShardedGemWell, owner_bounds
from tests import helpers

RANK_SHIFT, FEATURE_SHIFT = 35, 20  # cfg1: 20-bit UMI, 15-bit feature


class OracleEngine:
    def __init__(self, prob, lo, hi):
        from oracle import cro

        self.prob = prob
        cfg, t = prob["cfg"], prob["tables"]
        self.wl = t.whitelist
        self.wl_packed = t.wl_packed.astype(np.int64)
        self.o = cro.Oracle()
        w = self.o.add_whitelist(t.whitelist)
        self.lib = self.o.add_library(w, 0, cfg.bc_len, cfg.bc_len, cfg.umi_len)
        self.o.set_features(np.zeros(cfg.n_genes, dtype=np.int32))
        g = prob["gex"]
        self.o.add_reads(self.lib, g["r1_seq"][lo:hi], g["r1_qual"][lo:hi], g["feature"][lo:hi])
        self.prior = self.corrected = None
        self.recv = None
        self.own = (0, len(self.wl))
        self.result = None

    def make_shard(self):
        self.o.pass1(1)
        self.prior = torch.from_numpy(self.o.counts(self.lib, 0, self.wl).astype(np.int32))

    def prior_tensors(self):
        return [self.prior]

    def fb_counts_tensor(self):
        return None

    def barcode_correction(self):
        self.o.prior_clear(self.lib)
        self.o.prior_add(self.lib, self.wl, self.prior.numpy().astype(np.int64))
        self.o.pass2(1)
        self.corrected = torch.from_numpy(self.o.counts(self.lib, 1, self.wl).astype(np.int32))

    def corrected_tensors(self):
        return [self.corrected]

    def valid_count_tensors(self):
        return [self.prior + self.corrected]

    def _local_keys(self):
        r = self.o.reads()
        ok = ((r["state"] == 1) | (r["state"] == 2)) & ((r["flags"] & 1) != 0) & (r["feature"] != 0xFFFFFFFF)
        rank = np.searchsorted(self.wl_packed, synth.pack_2bit(r["bc"][ok]).astype(np.int64))
        umi = synth.pack_2bit(r["umi"][ok]).astype(np.int64)
        return (rank.astype(np.int64) << RANK_SHIFT) | (r["feature"][ok].astype(np.int64) << FEATURE_SHIFT) | umi

    def keys_partition(self, bounds):
        keys = np.sort(self._local_keys())
        cuts = np.searchsorted(keys, bounds.astype(np.int64) << RANK_SHIFT)
        return torch.from_numpy(keys), np.diff(cuts).astype(np.int64)

    def new_keys(self, n):
        return torch.empty(n, dtype=torch.int64)

    def keys_set(self, t):
        self.recv = t.numpy().copy()

    def set_owned_range(self, lo, hi):
        self.own = (lo, hi)

    def align_and_count(self):
        from oracle import pyref

        keys = self.recv if self.recv is not None else self._local_keys()
        tup = [(int(k >> RANK_SHIFT), 0, int((k >> FEATURE_SHIFT) & 0x7FFF), int(k & 0xFFFFF)) for k in keys]
        dd = pyref.dedup_count(tup, None, {0: True})
        valid = (self.prior + self.corrected).numpy()
        ranks = [r for r in range(self.own[0], self.own[1]) if valid[r] > 0]
        self.result = dict(ranks=np.array(ranks, dtype=np.int64), entries=np.array(dd["entries"], dtype=np.int64).reshape(-1, 3))

    def sync(self):
        pass


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, outdir, n_reads):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        prob = helpers.make_problem("cfg1", n_reads, n_whitelist=4000, n_cells=12)
        per = n_reads // world
        lo, hi = rank * per, (n_reads if rank == world - 1 else (rank + 1) * per)
        eng = OracleEngine(prob, lo, hi)
        sh = ShardedGemWell(eng, rank, world)
        sh.run()
        np.savez(os.path.join(outdir, f"rank{rank}.npz"), ranks=eng.result["ranks"], entries=eng.result["entries"],
                 bounds=sh.bounds.astype(np.int64), exchange=np.array([sh.exchange_bytes]))
    finally:
        dist.destroy_process_group()


def test_owner_bounds_balances_and_covers():
    v = torch.tensor([0, 5, 0, 0, 10, 1, 1, 1, 0, 2], dtype=torch.int32)
    for world in (1, 2, 3, 4):
        b = owner_bounds(v, world)
        assert b[0] == 0 and b[-1] == len(v) and np.all(np.diff(b.astype(np.int64)) >= 0)
    b = owner_bounds(v, 2)
    left = int(v[: b[1]].sum())
    assert 5 <= left <= 15  # the heavy barcode cannot be split; both sides get work
    assert owner_bounds(torch.zeros(0, dtype=torch.int32), 2).tolist() == [0, 0, 0]


def test_two_rank_sharding_matches_single_process():
    n_reads = 8000
    world = 2
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), d, n_reads), nprocs=world, join=True)
        parts = [np.load(os.path.join(d, f"rank{r}.npz")) for r in range(world)]
    assert np.array_equal(parts[0]["bounds"], parts[1]["bounds"])
    bounds = parts[0]["bounds"]
    # single-process reference
    prob = helpers.make_problem("cfg1", n_reads, n_whitelist=4000, n_cells=12)
    o = helpers.run_oracle(prob, threads=2)
    m = o.matrix()
    wl_packed = prob["tables"].wl_packed.astype(np.int64)
    ref_ranks = np.searchsorted(wl_packed, synth.pack_2bit(m["barcodes"]).astype(np.int64))
    got_ranks = np.concatenate([p["ranks"] for p in parts])
    assert np.array_equal(ref_ranks, got_ranks), "barcode index = concatenation of the owners' blocks"
    ent = np.concatenate([p["entries"] for p in parts])
    for r, p in enumerate(parts):  # every entry sits with its owner
        if len(p["entries"]):
            assert p["entries"][:, 0].min() >= bounds[r] and p["entries"][:, 0].max() < bounds[r + 1]
    cols = np.repeat(np.arange(len(ref_ranks)), np.diff(m["indptr"]))
    ref_ent = np.stack([ref_ranks[cols], m["indices"].astype(np.int64), m["data"].astype(np.int64)], axis=1)
    assert np.array_equal(ref_ent, ent), "matrix entries"
    assert all(int(p["exchange"][0]) > 0 for p in parts)
