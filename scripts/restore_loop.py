"""The whole restore loop on raw broker bytes (SURVEY §8 f1 + a1): poll -> sgr_ingest_record_batches_mt (host decode) ->
sgr_fold_ingested (GPU fold) -> committed offsets. Reports records/s end to end and the split between host decode,
H2D + fold, and key publication. Needs a GPU. Usage: python scripts/restore_loop.py [--partitions 8] [--polls 12] [--threads 8]
"""
import argparse
import os
import struct
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import kafka_batch as K  # noqa: E402  (encoder: scaffolding for building the input only)
from oracle import oracle as O  # noqa: E402  (checker)
from surge_b200 import formats as F  # noqa: E402
from surge_b200 import programs as P  # noqa: E402
from surge_b200.engine import ReplayEngine  # noqa: E402
from surge_b200.ingest import Ingest  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--partitions", type=int, default=8)
    ap.add_argument("--polls", type=int, default=12)
    ap.add_argument("--threads", type=int, default=8)
    ap.add_argument("--records-per-batch", type=int, default=500)
    ap.add_argument("--batches-per-fetch", type=int, default=320)
    ap.add_argument("--keys-per-partition", type=int, default=100_000)
    ap.add_argument("--compression", default="lz4")
    a = ap.parse_args()
    rng = np.random.default_rng(17)
    # 16 prototype batches per partition (the python lz4 encoder is slow); a fetch repeats them with patched base offsets
    protos = {}
    for p in range(a.partitions):
        protos[p] = []
        for _ in range(16):
            recs = []
            for d in range(a.records_per_batch):
                k = int(rng.integers(0, a.keys_per_partition))
                recs.append((d, f"p{p}-aggregate-{k:07d}:{d}".encode(), struct.pack("<IIi", int(rng.choice([0, 1, 2], p=[0.45, 0.45, 0.1])), d + 1, int(rng.integers(0, 1 << 31)))))
            protos[p].append(K.encode_record_batch(0, recs, compression=a.compression, producer_id=p, transactional=True))
    nxt = {p: 0 for p in range(a.partitions)}

    def poll():
        out = []
        for p in range(a.partitions):
            blob = bytearray()
            for b in range(a.batches_per_fetch):
                pb = protos[p][b % 16]
                blob += struct.pack(">q", nxt[p]) + pb[8:]
                nxt[p] += a.records_per_batch
            blob += K.encode_control_batch(nxt[p], p, K.COMMIT)
            nxt[p] += 1
            out.append((p, bytes(blob)))
        return out

    ing = Ingest()
    t_decode = t_fold = 0.0
    n_rec = wire = 0
    want = np.zeros((0, 16), np.uint8)
    check_polls = 2
    with ReplayEngine(0) as e:
        e.register_program(P.counter_program())
        for i in range(a.polls):
            fetches = poll()
            t0 = time.perf_counter()
            st = ing.record_batches_mt(fetches, threads=a.threads)
            t1 = time.perf_counter()
            batch = ing.pending() if i < check_polls else None
            t1b = time.perf_counter()
            e.fold_ingested(ing)
            t2 = time.perf_counter()
            if i < check_polls:   # the first polls are checked against the oracle (and warm the buffers); later ones are timed
                cap = e.n_aggregates()
                grown = np.zeros((cap, 16), np.uint8)
                grown[: want.shape[0]] = want
                want = O.fold_incremental(O.MODEL_COUNTER, batch.view(F.REC64).reshape(-1), grown)
                assert np.array_equal(e.export_states(), want), "state table differs from the oracle"
            else:
                t_decode += t1 - t0
                t_fold += t2 - t1b
                n_rec += sum(s["n_records"] for s in st)
                wire += sum(len(b) for _, b in fetches)
            for p in range(a.partitions):
                assert ing.offsets(p) == (nxt[p], nxt[p])
        stats = e.stats()
    total = t_decode + t_fold
    print({"records": n_rec, "wire_mb": round(wire / 1e6, 1), "wire_bytes_per_record": round(wire / n_rec, 2), "threads": a.threads,
           "decode_s": round(t_decode, 4), "h2d_fold_publish_s": round(t_fold, 4),
           "records_per_s_end_to_end": round(n_rec / total), "records_per_s_decode_only": round(n_rec / t_decode),
           "last_fold_device_ms": round(float(stats.ms_fold), 4), "keys": len(ing.keys()), "checked_polls_bit_exact": check_polls})


if __name__ == "__main__":
    main()
