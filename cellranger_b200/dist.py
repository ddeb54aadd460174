"""Barcode-owner sharding of the path across the GPUs of one box (one process per GPU).

The reference partitions the same way, by barcode range, but through files
(ShardReader::make_chunks, lib/rust/cr_lib/src/stages/barcode_correction.rs:252-262 and
stages/align_and_count.rs:519-524); priors are summed over every chunk in the MAKE_SHARD join
(stages/make_shard.rs:303-358). Here:

  1. every rank runs pass 1 on its own reads                      (local)
  2. all-reduce(sum) of the per-library prior histograms          (priors are global)
     + the exact feature-barcode counts
  3. every rank corrects its own invalid reads                    (local)
  4. all-reduce(sum) of the corrected-read counts; prior + corrected = the valid-barcode counts
     (barcode index) → owner ranges of contiguous content ranks balanced by read count
  5. the packed 64-bit keys go to the rank that owns their barcode: fused scatter into the owner's
     receive buffer through peer memory (NVLink stores, CUDA IPC mapping), or - when the GPUs have no peer
     access - grouping by owner plus one NCCL all-to-all
  6. dedup + counting, shard-local; the matrix is the concatenation of the ranks' column blocks

The product path is `NativeShardedGemWell`: every step above happens inside libcrgpu.so
(`crgpu_comm_init` / `crgpu_sharded_run`, cellranger_b200/csrc/shard.cu) on the context stream - NCCL
all-reduces, owner ranges from a device scan, peer stores, an on-stream barrier - and this module only calls it.

`ShardedGemWell` is the same protocol written against an abstract *engine*, with torch.distributed doing the
collectives: the world_size-2 gloo tests on CPU drive it with a CPU reference engine defined in the tests (host logic: owner ranges,
split sizes, the exchange), the GPU tests drive it through `TorchEngine` as a second, independent implementation
that the native path must agree with (and as the NCCL all-to-all route for GPUs without peer access).
"""
from __future__ import annotations

import os
import time
from typing import List, Sequence

import numpy as np
import torch
import torch.distributed as dist


class _DevArray:
    """A raw device pointer as a __cuda_array_interface__ object (zero copy)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


def dev_tensor(ptr: int, n: int, dtype: str, device) -> torch.Tensor:
    if n == 0 or not ptr:
        return torch.empty(0, dtype={"<i4": torch.int32, "<i8": torch.int64}[dtype], device=device)
    return torch.as_tensor(_DevArray(ptr, n, dtype), device=device)


class TorchEngine:
    """GemWell (GPU) behind the engine interface; tensors alias the library's device buffers."""

    def __init__(self, gw, n_libs: int):
        self.gw = gw
        self.n_libs = n_libs
        self.device = torch.device("cuda", gw.device)
        self._recv = None
        self.p2p = False

    def setup_peer_exchange(self, rank: int, world: int, capacity_keys: int, group=None) -> bool:
        """Map every rank's receive buffer into every other rank (CUDA IPC). Returns False - and leaves the
        NCCL all-to-all in place - when the GPUs cannot reach each other's memory."""
        if world < 2 or self.p2p:
            return self.p2p
        ok = all(torch.cuda.can_device_access_peer(self.gw.device, d) for d in range(torch.cuda.device_count())
                 if d != self.gw.device)
        flag = torch.tensor([1 if ok else 0], device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            return False
        mine = torch.frombuffer(bytearray(self.gw.exchange_init(capacity_keys)), dtype=torch.uint8).to(self.device)
        every = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(every, mine, group=group)
        self.gw.exchange_connect(world, rank, b"".join(bytes(x.cpu().numpy()) for x in every))
        self.p2p = True
        return True

    def exchange_reset(self):
        self.gw.exchange_reset()

    def keys_scatter_peers_begin(self, bounds: np.ndarray):
        self.gw.keys_scatter_peers_begin(bounds)

    def keys_scatter_peers(self, bounds: np.ndarray) -> np.ndarray:
        return self.gw.keys_scatter_peers(bounds)

    def exchange_finish(self) -> int:
        return self.gw.exchange_finish()

    def make_shard(self):
        self.gw.make_shard()
        self.gw.sync()

    def prior_tensors(self) -> List[torch.Tensor]:
        return [dev_tensor(*self.gw.prior_dev(l), "<i4", self.device) for l in range(self.n_libs)]

    def fb_counts_tensor(self):
        p, n = self.gw.fb_counts_dev()
        return dev_tensor(p, n, "<i8", self.device) if n else None

    def barcode_correction(self):
        self.gw.barcode_correction()
        self.gw.sync()

    def corrected_tensors(self) -> List[torch.Tensor]:
        return [dev_tensor(*self.gw.corrected_dev(l), "<i4", self.device) for l in range(self.n_libs)]

    def valid_count_tensors(self) -> List[torch.Tensor]:
        """valid = prior + corrected, recomputed from the (now global) vectors."""
        self.gw.valid_counts_refresh()
        self.gw.sync()
        return [dev_tensor(*self.gw.valid_counts_dev(l), "<i4", self.device) for l in range(self.n_libs)]

    def keys_partition(self, bounds: np.ndarray):
        counts = self.gw.keys_partition(bounds)
        p, n = self.gw.keys_dev()
        return dev_tensor(p, n, "<i8", self.device), counts.astype(np.int64)

    def new_keys(self, n: int) -> torch.Tensor:
        self._recv = torch.empty(max(n, 1), dtype=torch.int64, device=self.device)[:n]
        return self._recv

    def keys_set(self, t: torch.Tensor):
        torch.cuda.synchronize(self.device)
        self.gw.keys_set(t.data_ptr() if t.numel() else 0, int(t.numel()))

    def set_owned_range(self, lo: int, hi: int):
        self.gw.set_owned_range(lo, hi)

    def align_and_count(self):
        self.gw.align_and_count()

    def sync(self):
        torch.cuda.synchronize(self.device)


def owner_bounds(valid_total: torch.Tensor, world: int) -> np.ndarray:
    """Contiguous content-rank ranges [bounds[r], bounds[r+1]) holding ~equal numbers of valid reads:
    the GPU analogue of ShardReader::make_chunks' equal-read barcode ranges."""
    n = int(valid_total.numel())
    if n == 0:
        return np.zeros(world + 1, dtype=np.uint32)
    csum = torch.cumsum(valid_total.to(torch.int64), 0)
    total = int(csum[-1].item())
    targets = torch.tensor([(total * r) // world for r in range(1, world)], dtype=torch.int64, device=csum.device)
    cuts = torch.searchsorted(csum, targets, right=False) + 1 if world > 1 else torch.empty(0, dtype=torch.int64)
    b = np.zeros(world + 1, dtype=np.int64)
    b[world] = n
    if world > 1:
        b[1:world] = np.minimum(cuts.cpu().numpy(), n)
    b = np.maximum.accumulate(b)
    return b.astype(np.uint32)


class ShardedGemWell:
    def __init__(self, engine, rank: int, world: int, group=None, early_scatter=None):
        self.e = engine
        # Early part of the fused exchange (the keys of pass 1 travel while pass 2 runs; owner ranges then come
        # from the valid-before counts alone). Off by default: at 2 GPUs it measured 2 % slower than sending
        # everything after pass 2 (the scatter competes with pass 2 for the SMs and adds one host sync); it is
        # meant for 8 GPUs, where 7/8 of the keys are remote - CRGPU_EARLY_SCATTER=1 turns it on.
        self.early_scatter = bool(os.environ.get("CRGPU_EARLY_SCATTER")) if early_scatter is None else bool(early_scatter)
        self.rank, self.world, self.group = rank, world, group
        self.bounds = None
        self.exchange_bytes = 0
        # CRGPU_DIST_TIMING=1: host-clock time of every phase of run() (with a device sync at each boundary, so
        # only for diagnosis - the syncs cost a little themselves)
        self.timing = bool(os.environ.get("CRGPU_DIST_TIMING"))
        self.times = {}

    def _t(self, name, t0):
        if not self.timing:
            return t0
        self.e.sync()
        t1 = time.perf_counter()
        self.times[name] = self.times.get(name, 0.0) + (t1 - t0) * 1e3
        return t1

    def _allreduce(self, t: torch.Tensor):
        if self.world > 1 and t is not None and t.numel():
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def run(self):
        e = self.e
        t = time.perf_counter() if self.timing else 0.0
        p2p = self.world > 1 and getattr(e, "p2p", False)
        if p2p:
            e.exchange_reset()  # ordered before every peer's scatter by the all-reduces below
        e.make_shard()
        t = self._t("pass1", t)
        priors = e.prior_tensors()
        for x in priors:
            self._allreduce(x)
        self._allreduce(e.fb_counts_tensor())
        e.sync()
        t = self._t("allreduce.priors", t)
        early = p2p and self.early_scatter
        if early:
            # Owner ranges from the (global) valid-before counts alone: they are known one pass earlier than the
            # corrected counts, so the keys of pass 1 - 85 % of all keys - travel over NVLink while pass 2 runs.
            total = priors[0].to(torch.int64)
            for x in priors[1:]:
                total = total + x.to(torch.int64)
            self.bounds = owner_bounds(total, self.world)
            e.keys_scatter_peers_begin(self.bounds)
            t = self._t("owner_bounds", t)
        e.barcode_correction()
        t = self._t("pass2", t)
        for x in e.corrected_tensors():
            self._allreduce(x)
        e.sync()
        t = self._t("allreduce.corrected", t)
        valid = e.valid_count_tensors()  # identical on every rank; the barcode index of the count stage
        if not early:
            total = valid[0].to(torch.int64)
            for x in valid[1:]:
                total = total + x.to(torch.int64)
            self.bounds = owner_bounds(total, self.world)
            t = self._t("owner_bounds", t)
        if p2p:
            # fused: one pass writes every key into its owner's receive buffer over NVLink
            sent = e.keys_scatter_peers(self.bounds)
            self.exchange_bytes = int(sent.sum() - sent[self.rank]) * 8
            dist.barrier(group=self.group)  # every rank's stores have landed
            e.exchange_finish()
        elif self.world > 1:
            keys, send_counts = e.keys_partition(self.bounds)
            sc = torch.as_tensor(send_counts, dtype=torch.int64, device=keys.device)
            rc = torch.empty_like(sc)
            dist.all_to_all_single(rc, sc, group=self.group)
            recv_counts = rc.cpu().numpy()
            recv = e.new_keys(int(recv_counts.sum()))
            dist.all_to_all_single(recv, keys, output_split_sizes=[int(x) for x in recv_counts],
                                   input_split_sizes=[int(x) for x in send_counts], group=self.group)
            self.exchange_bytes = int(send_counts.sum() - send_counts[self.rank]) * 8
            e.keys_set(recv)
        t = self._t("exchange", t)
        e.set_owned_range(int(self.bounds[self.rank]), int(self.bounds[self.rank + 1]))
        e.align_and_count()
        t = self._t("count", t)


class NativeShardedGemWell:
    """The sharded run behind the C ABI: this class holds no logic of the step. `unique_id` comes from
    cellranger_b200.comm_unique_id() on rank 0 and reaches the other ranks through the host's own channel
    (torch.distributed.broadcast_object_list under torchrun, an argument of mp.spawn in the tests)."""

    def __init__(self, gw, rank: int, world: int, unique_id: bytes, capacity_keys: int):
        self.gw, self.rank, self.world = gw, rank, world
        gw.comm_init(unique_id, world, rank, capacity_keys)
        self.bounds = None
        self.exchange_bytes = 0

    def run(self):
        self.gw.sharded_run()
        self.bounds = self.gw.owner_bounds()
        self.exchange_bytes = self.gw.shard_stats()["sent_remote_keys"] * 8
