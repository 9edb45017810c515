"""Host-side plumbing of the multi-GPU replay: one process per GPU, torch.distributed only for the
rendezvous (NCCL unique id, CUDA IPC handles, barriers); the data path is the engine's own
route kernel + NCCL / peer-memory exchange (csrc/dist.cu).

Ownership mirrors the reference: aggregateId -> partition (KafkaPartitionProvider.partitionForKey,
modules/common/src/main/scala/surge/kafka/KafkaPartitioner.scala:7-9) -> owner = partition % nranks.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import native as N


def partitions_for_keys(keys: Sequence[str], num_partitions: int, up_to_colon: bool = True) -> np.ndarray:
    """partition_of[i] = abs(MurmurHash3.stringHash(keys[i].takeWhile(_ != ':')) % num_partitions)."""
    lib = N.load_library()
    enc = [k.encode("utf-8") for k in keys]
    offs = np.zeros(len(enc) + 1, dtype=np.uint32)
    np.cumsum([len(b) for b in enc], out=offs[1:])
    blob = np.frombuffer(b"".join(enc) or b"\0", dtype=np.uint8).copy()
    out = np.zeros(len(enc), dtype=np.uint32)
    rc = lib.sgr_partitions_for_keys(blob.ctypes.data, offs.ctypes.data, len(enc), num_partitions, 1 if up_to_colon else 0, out.ctypes.data)
    if rc != 0:
        raise ValueError(f"sgr_partitions_for_keys -> {rc}")
    return out


def owner_and_local_index(partition_of_agg: np.ndarray, nranks: int) -> Tuple[np.ndarray, np.ndarray, List[np.ndarray]]:
    """numpy mirror of the device tables (csrc/dist.cu): owner rank, local dense index on the owner, and per rank
    the global indices of its local slots. Used by the CPU (gloo) tests of the routing logic."""
    owner = (np.asarray(partition_of_agg, dtype=np.uint32) % np.uint32(nranks)).astype(np.uint8)
    local = np.zeros(len(owner), dtype=np.uint32)
    globals_of = []
    for r in range(nranks):
        idx = np.nonzero(owner == r)[0]
        local[idx] = np.arange(len(idx), dtype=np.uint32)
        globals_of.append(idx.astype(np.uint32))
    return owner, local, globals_of


def route_on_host(records: np.ndarray, owner: np.ndarray, local: np.ndarray, nranks: int) -> List[np.ndarray]:
    """Stable partition of REC64 records by owner with the agg field rewritten to the owner's local index:
    what K4 (route_scatter) produces, as numpy. Returns one array per destination rank."""
    agg = records["agg"].astype(np.int64)
    o = owner[agg]
    out = []
    for r in range(nranks):
        sel = records[o == r].copy()           # boolean mask keeps order: stable
        sel["agg"] = local[agg[o == r]]
        out.append(sel)
    return out


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def states_hash(states: np.ndarray, global_ids: Optional[np.ndarray] = None) -> int:
    """numpy twin of sgr_states_hash (csrc/bulk_fold.cu states_hash_kernel): sum over slots of
    mix(id, state bytes) mod 2^64 — order independent, so per-rank hashes of a routed table add up to the hash of the
    whole table under global aggregate indices."""
    st = np.ascontiguousarray(states).view(np.uint8)
    n = st.shape[0]
    if n == 0:
        return 0
    words = st.reshape(n, -1).view(np.uint64)
    ids = np.arange(n, dtype=np.uint64) if global_ids is None else np.asarray(global_ids).astype(np.uint64)
    with np.errstate(over="ignore"):
        h = _splitmix64(ids)
        for k in range(words.shape[1]):
            h = _splitmix64(h ^ words[:, k])
        return int(h.sum(dtype=np.uint64))


def partitions_of_rank(rank: int, nranks: int, num_partitions: int) -> List[int]:
    """Topic partitions a rank consumes when the store is fed from the topic itself (surge_b200/ingest.py): the broker has already
    done the shuffle — every record of an aggregate sits in partition partitionForKey(id) — so rank r decodes and folds the
    partitions p with p % nranks == r and NO exchange between GPUs is needed; the same owner rule as the routed path."""
    return [p for p in range(num_partitions) if p % nranks == rank]


def exchange_ids(engine, rank: int, nranks: int, recv_capacity_records: int, fused: bool = True) -> None:
    """Rendezvous over torch.distributed: NCCL unique id from rank 0, then (fused path) the IPC handles."""
    import torch.distributed as dist

    lib = N.load_library()
    uid = [None]
    if rank == 0 and nranks > 1:
        buf = C.create_string_buffer(128)
        rc = lib.sgr_dist_unique_id(buf)
        if rc != 0:
            raise N.SgrError(rc, (lib.sgr_last_error(None) or b"").decode())
        uid = [bytes(buf.raw)]
    if nranks > 1:
        dist.broadcast_object_list(uid, src=0)
    engine.dist_init(rank, nranks, uid[0], recv_capacity_records)
    if fused and nranks > 1:
        mine = engine.dist_ipc_export()
        handles: List[Optional[bytes]] = [None] * nranks
        dist.all_gather_object(handles, mine)
        engine.dist_ipc_import(handles)
