"""CPU tests of the N>1 host logic (gloo, world_size 2): ownership tables, stable routing, the exchange plan.
The device kernels (csrc/dist.cu) mirror surge_b200.dist.route_on_host; here the numpy mirror is driven through a real
2-process all-to-all and checked against the single-process oracle, so the routing contract is pinned without a GPU."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank: int, world: int, port: int, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from surge_b200 import dist as D
    from surge_b200 import formats as F
    from surge_b200 import synth as S

    n_global, n_src = 3000, 16
    rng = np.random.default_rng(1)
    counts = rng.integers(0, 12, size=n_global)
    rec, off = S.counter_csr(n_global, counts, seed=2, p_throw=0.01)
    want, _, _ = O.fold_packed(O.MODEL_COUNTER, O.REC_FIXED64, rec, off)
    arrival = S.interleave_arrival(rec, seed=3)
    mine = arrival[((arrival["agg"] % n_src) % world) == rank]        # this rank's source partitions, arrival order
    part = D.partitions_for_keys([f"agg-{g}" for g in range(n_global)], 32)
    owner, local, globals_of = D.owner_and_local_index(part, world)
    sends = D.route_on_host(mine, owner, local, world)
    # counts all-gather, then the all-to-all of packed records
    cnt = torch.tensor([len(s) for s in sends], dtype=torch.int64)
    allc = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allc, cnt)
    recv_counts = [int(allc[s][rank]) for s in range(world)]
    send_t = torch.from_numpy(np.concatenate(sends).view(np.uint8).reshape(-1).copy())
    recv_t = torch.zeros(sum(recv_counts) * 64, dtype=torch.uint8)
    dist.all_to_all_single(recv_t, send_t, [c * 64 for c in recv_counts], [len(s) * 64 for s in sends])
    got_rec = recv_t.numpy().view(F.REC64)
    n_local = len(globals_of[rank])
    grouped, goff = O.group_by_agg(got_rec, n_local)
    states, _, _ = O.fold_packed(O.MODEL_COUNTER, O.REC_FIXED64, grouped, goff)
    ok = np.array_equal(states, want[globals_of[rank].astype(np.int64)])
    total = torch.tensor([len(got_rec)], dtype=torch.int64)
    dist.all_reduce(total)
    ret[rank] = (bool(ok), int(total[0]) == len(rec), n_local)
    dist.destroy_process_group()


def test_two_rank_routing_matches_single_process_fold():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret[r][0] for r in range(world)), dict(ret)
    assert all(ret[r][1] for r in range(world))
    assert sum(ret[r][2] for r in range(world)) == 3000


def test_owner_tables_follow_the_reference_partitioner():
    from oracle import surge_model as M
    from surge_b200 import dist as D

    keys = [f"agg-{g}:{g % 7}" for g in range(200)]
    part = D.partitions_for_keys(keys, 32, up_to_colon=True)
    for k, p in zip(keys, part):
        assert p == M.partition_for_key(M.partition_string_up_to_colon(k), 32)
    owner, local, globals_of = D.owner_and_local_index(part, 4)
    assert (owner == part % 4).all()
    for r in range(4):
        assert (owner[globals_of[r]] == r).all() and (local[globals_of[r]] == np.arange(len(globals_of[r]))).all()


def test_topic_partitions_split_over_ranks_need_no_exchange():
    """Feeding from the topic: rank r takes the partitions p % nranks == r; because the partition of a record IS
    partitionForKey(aggregate id), every aggregate's records land on exactly one rank and ranks share no aggregate."""
    from surge_b200 import dist as D

    keys = [f"agg-{i}" for i in range(5000)]
    n_part, nranks = 32, 4
    part = D.partitions_for_keys(keys, n_part)
    owners = {}
    for r in range(nranks):
        mine = set(D.partitions_of_rank(r, nranks, n_part))
        assert mine == {p for p in range(n_part) if p % nranks == r}
        for k, p in zip(keys, part.tolist()):
            if p in mine:
                assert k not in owners
                owners[k] = r
    assert len(owners) == len(keys)
    owner, _, _ = D.owner_and_local_index(part, nranks)          # the routed path's owner table says the same
    assert [owners[k] for k in keys] == owner.tolist()
