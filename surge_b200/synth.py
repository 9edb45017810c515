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
