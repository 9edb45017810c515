"""Deterministic synthetic event logs for the BASELINE.json configs (there is no network for datasets).

Host generators use numpy's Philox bit generator (seeded per config);
the device generator builds the same record layout with torch on the GPU for sizes that
would take minutes to build and copy from the host (config 2 and up).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

from . import formats as F


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(seed))


def counter_csr(n_agg: int, events_per_agg, seed: int, p_incr: float = 0.45, p_decr: float = 0.45,
                p_throw: float = 0.0, by_max: int = 1 << 31) -> Tuple[np.ndarray, np.ndarray]:
    """Counter log in CSR order. events_per_agg: int or per-aggregate counts.
    type ~ {Incr p_incr, Decr p_decr, NoOp rest, Throw p_throw}; by uniform in [0, by_max) (forces wraparound);
    seq = 1..k per aggregate. Returns (records REC64[n], seg_offsets u64[n_agg+1])."""
    rng = _rng(seed)
    counts = np.full(n_agg, events_per_agg, dtype=np.int64) if np.isscalar(events_per_agg) else np.asarray(events_per_agg, dtype=np.int64)
    n = int(counts.sum())
    agg = np.repeat(np.arange(n_agg, dtype=np.uint64), counts)
    starts = np.zeros(n_agg + 1, dtype=np.int64)
    np.cumsum(counts, out=starts[1:])
    seq = (np.arange(n, dtype=np.int64) - np.repeat(starts[:-1], counts) + 1).astype(np.uint32)
    u = rng.random(n)
    types = np.full(n, F.NO_OP_EVENT, dtype=np.uint32)
    types[u < p_incr + p_decr] = F.COUNT_DECREMENTED
    types[u < p_incr] = F.COUNT_INCREMENTED
    if p_throw > 0:
        types[u > 1.0 - p_throw] = F.EXCEPTION_THROWING_EVENT
    by = rng.integers(0, by_max, size=n, dtype=np.int64).astype(np.int32)
    rec = F.counter_records(types, seq, agg, by)
    return rec, F.csr_offsets_from_counts(counts)


def interleave_arrival(records: np.ndarray, seed: int) -> np.ndarray:
    """Re-order CSR records into a plausible ARRIVAL order: aggregates interleaved, each
    aggregate's own order preserved (one key -> one Kafka partition -> log order)."""
    rng = _rng(seed)
    n = len(records)
    slots = rng.random(n)
    agg = records["agg"]
    idx = np.lexsort((slots, agg))          # grouped by agg (already ascending), slots ascending inside
    slots_in_order = slots[idx]
    perm = np.argsort(slots_in_order, kind="stable")
    return records[perm]


def counter_var_csr(n_agg: int, counts, seed: int, payload_min: int = 32, payload_max: int = 512, with_directory: bool = False):
    """Counter-with-payload (config 4): variable records, payload length uniform in [min,max],
    first 4 payload bytes = by, the rest is filler the fold must still read.
    Returns (bytes u8[total], seg_offsets u64[n_agg+1])."""
    rng = _rng(seed)
    counts = np.full(n_agg, counts, dtype=np.int64) if np.isscalar(counts) else np.asarray(counts, dtype=np.int64)
    n = int(counts.sum())
    plen = rng.integers(payload_min, payload_max + 1, size=n, dtype=np.int64)
    rlen = 16 + ((plen + 15) // 16) * 16
    rec_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(rlen, out=rec_off[1:])
    total = int(rec_off[-1])
    buf = rng.integers(0, 256, size=total, dtype=np.uint8)  # filler everywhere, then headers on top
    u = rng.random(n)
    types = np.full(n, F.NO_OP_EVENT, dtype=np.uint32)
    types[u < 0.9] = F.COUNT_DECREMENTED
    types[u < 0.45] = F.COUNT_INCREMENTED
    starts = np.zeros(n_agg + 1, dtype=np.int64)
    np.cumsum(counts, out=starts[1:])
    agg = np.repeat(np.arange(n_agg, dtype=np.uint32), counts)
    seq = (np.arange(n, dtype=np.int64) - np.repeat(starts[:-1], counts) + 1).astype(np.uint32)
    hdr = np.zeros((n, 4), dtype=np.uint32)
    hdr[:, 0], hdr[:, 1], hdr[:, 2], hdr[:, 3] = types, seq, plen.astype(np.uint32), agg
    hb = hdr.view(np.uint8).reshape(n, 16)
    pos = rec_off[:-1]
    for j in range(16):
        buf[pos + j] = hb[:, j]
    # zero the padding so the log is canonical
    pad = rlen - 16 - plen
    for j in range(1, 16):
        sel = pad >= j
        buf[(rec_off[1:] - j)[sel]] = 0
    seg = np.zeros(n_agg + 1, dtype=np.uint64)
    seg[1:] = rec_off[starts[1:]]
    if with_directory:
        return buf, seg, rec_off.astype(np.uint64)
    return buf, seg


def zipf_counts(n_agg: int, n_events: int, alpha: float, seed: int) -> np.ndarray:
    """Events per aggregate for keys ~ Zipf(alpha) over n_agg ranks (inverse-CDF sampling)."""
    rng = _rng(seed)
    w = 1.0 / np.power(np.arange(1, n_agg + 1, dtype=np.float64), alpha)
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    keys = np.searchsorted(cdf, rng.random(n_events), side="left")
    return np.bincount(keys, minlength=n_agg).astype(np.int64)


def counter_csr_device(n_agg: int, events_per_agg: int, seed: int, device: str = "cuda:0"):
    """Config-2-shaped Counter log built on the GPU with torch (uniform events per aggregate).
    Returns (records int32[n,16] CUDA tensor, seg_offsets int64[n_agg+1] CUDA tensor)."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    n = n_agg * events_per_agg
    rec = torch.zeros((n, 16), dtype=torch.int32, device=device)
    u = torch.rand(n, generator=g, device=device)
    types = torch.full((n,), F.NO_OP_EVENT, dtype=torch.int32, device=device)
    types[u < 0.9] = F.COUNT_DECREMENTED
    types[u < 0.45] = F.COUNT_INCREMENTED
    rec[:, 0] = types
    idx = torch.arange(n, device=device, dtype=torch.int64)
    rec[:, 1] = (idx % events_per_agg + 1).to(torch.int32)
    agg = idx // events_per_agg
    rec[:, 2] = (agg & 0xFFFFFFFF).to(torch.int32)
    rec[:, 3] = (agg >> 32).to(torch.int32)
    rec[:, 4] = torch.randint(0, 1 << 31, (n,), generator=g, device=device, dtype=torch.int64).to(torch.int32)
    # payload bytes the Counter fold ignores but must read: make them non-trivial
    rec[:, 8:16] = torch.randint(-(1 << 31), 1 << 31, (n, 8), generator=g, device=device, dtype=torch.int64).to(torch.int32)
    off = torch.arange(n_agg + 1, device=device, dtype=torch.int64) * (events_per_agg * 64)
    return rec, off


# ---------------------------------------------------------------- configs[2]: a log that is a pure function of (aggregate, event index)
# Every rank count sees the SAME logical log (so state hashes are comparable across N), and the CPU oracle can regenerate the
# events of any sampled aggregate without holding the 64 GB log.
_SM_A, _SM_B, _SM_G = 0xBF58476D1CE4E5B9, 0x94D049BB133111EB, 0x9E3779B97F4A7C15


def _routed_key(g, k, seed):
    return g * 1000003 + k * 7919 + seed * 0x51ED27


def routed_events_host(g_ids: np.ndarray, epa: int, seed: int) -> Tuple[np.ndarray, np.ndarray]:
    """Counter events of the given global aggregates, CSR order (numpy twin of routed_log_device): event k of aggregate g has
    type/by = f(splitmix64(g, k, seed)), seq = k + 1. Records carry agg = position in g_ids."""
    g = np.repeat(np.asarray(g_ids, dtype=np.uint64), epa)
    k = np.tile(np.arange(epa, dtype=np.uint64), len(g_ids))
    with np.errstate(over="ignore"):
        x = _routed_key(g, k, np.uint64(seed)) + np.uint64(_SM_G)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(_SM_A)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(_SM_B)
        x = x ^ (x >> np.uint64(31))
    u = (x & np.uint64(0xFFFF)).astype(np.int64)
    types = np.where(u < 29491, F.COUNT_INCREMENTED, np.where(u < 58982, F.COUNT_DECREMENTED, F.NO_OP_EVENT)).astype(np.uint32)
    by = ((x >> np.uint64(16)) & np.uint64(0x7FFFFFFF)).astype(np.int64).astype(np.int32)
    agg = np.repeat(np.arange(len(g_ids), dtype=np.uint64), epa)
    rec = F.counter_records(types, (k + np.uint64(1)).astype(np.uint32), agg, by)
    return rec, F.csr_offsets_from_counts(np.full(len(g_ids), epa, dtype=np.int64))


def routed_round(g, k: int, seed: int):
    """(type, by) int32 CUDA tensors of event k of the global aggregates g (int64 CUDA tensor): the torch form of the generator."""
    import torch

    def s64(c):
        return c - (1 << 64) if c >= (1 << 63) else c

    def lsr(x, n):
        return (x >> n) & ((1 << (64 - n)) - 1)

    x = g * 1000003 + (k * 7919 + seed * 0x51ED27) + s64(_SM_G)
    x = (x ^ lsr(x, 30)) * s64(_SM_A)
    x = (x ^ lsr(x, 27)) * s64(_SM_B)
    x = x ^ lsr(x, 31)
    u = x & 0xFFFF
    typ = torch.where(u < 29491, F.COUNT_INCREMENTED, torch.where(u < 58982, F.COUNT_DECREMENTED, F.NO_OP_EVENT)).to(torch.int32)
    return typ, (lsr(x, 16) & 0x7FFFFFFF).to(torch.int32)


def routed_log_device(rank: int, world: int, n_global: int, epa: int, seed: int, device: str):
    """This rank's share of the configs[2] log, on the device, in ARRIVAL order: the rank holds the source partitions of the
    aggregates g with g % world == rank; event k of every one of its aggregates comes before event k+1 of any (aggregates
    interleaved, every aggregate's own order kept). Records carry the GLOBAL aggregate index. int32[n, 16] CUDA tensor."""
    import torch

    g = torch.arange(rank, n_global, world, device=device, dtype=torch.int64)
    na = g.numel()
    rec = torch.zeros((na * epa, 16), dtype=torch.int32, device=device)
    g32 = g.to(torch.int32)
    for k in range(epa):   # one round of events at a time keeps the temporaries small (the full log is 64 GB)
        typ, by = routed_round(g, k, seed)
        blk = rec[k * na:(k + 1) * na]
        blk[:, 0] = typ
        blk[:, 1] = k + 1
        blk[:, 2] = g32
        blk[:, 4] = by
    return rec


def routed_partitions(n_global: int, num_partitions: int = 64) -> np.ndarray:
    """State-topic partition of every dense aggregate id for the synthetic configs[2] runs: ids are pre-hashed once on load
     — a multiplicative hash stands in for partitionForKey on 10^7 synthetic ids."""
    with np.errstate(over="ignore"):
        return ((np.arange(n_global, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(40)).astype(np.uint32) % np.uint32(num_partitions)
