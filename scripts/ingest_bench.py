"""Host decode throughput of sgr_ingest_record_batches: RecordBatch v2 bytes -> packed 64-byte records.

Builds a few batches with the test encoder (oracle/kafka_batch.py — measurement scaffolding only, the timed call is the
product's C decoder), replicates them with patched base offsets (the batch CRC does not cover baseOffset) and times
the decode on one host thread. Usage: python scripts/ingest_bench.py [--records-per-batch 500] [--copies 200]
"""
import argparse
import ctypes as C
import os
import struct
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import kafka_batch as K  # noqa: E402
from surge_b200 import native as N  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records-per-batch", type=int, default=500)
    ap.add_argument("--batches", type=int, default=16)
    ap.add_argument("--copies", type=int, default=200)
    ap.add_argument("--copies-per-fetch", type=int, default=20)
    ap.add_argument("--keys", type=int, default=1 << 20)
    a = ap.parse_args()
    lib = N.load_library()
    rng = np.random.default_rng(9)
    out = {}
    for compression in ("none", "lz4"):
        protos = []
        for _ in range(a.batches):
            recs = []
            for d in range(a.records_per_batch):
                k = int(rng.integers(0, a.keys))
                recs.append((d, f"aggregate-{k:08d}:{d}".encode(), struct.pack("<IIi", int(rng.integers(0, 3)), d, int(rng.integers(0, 100)))))
            protos.append(K.encode_record_batch(0, recs, compression=compression, producer_id=1, transactional=True,
                                                headers=[(b"aggregate_id", b"aggregate-00000000")]))
        blob = bytearray()
        fetches = []
        off = 0
        # where the 8 key digits sit in each prototype (uncompressed only: a copy gets fresh random keys and a fresh CRC,
        # computed with the library's own CRC routine, so that the id dictionary grows the way a real topic makes it grow)
        key_pos = []
        for b in protos:
            arr = np.frombuffer(b, dtype=np.uint8)
            hits = [i + 10 for i in range(len(b) - 19) if b[i:i + 10] == b"aggregate-" and b[i + 18:i + 19] == b":"]
            key_pos.append(np.asarray(hits, dtype=np.int64))
        for c in range(a.copies):
            for bi, b in enumerate(protos):
                if compression == "none" and c:
                    m = np.frombuffer(b, dtype=np.uint8).copy()
                    ks = rng.integers(0, a.keys, len(key_pos[bi]))
                    for dgt in range(8):
                        m[key_pos[bi] + dgt] = 48 + (ks // 10 ** (7 - dgt)) % 10
                    body = m[21:].tobytes()
                    b = b[:17] + struct.pack(">I", lib.sgr_crc32c(body, len(body))) + body
                blob += struct.pack(">q", off) + b[8:]
                off += a.records_per_batch
            if (c + 1) % a.copies_per_fetch == 0:
                fetches.append(bytes(blob))
                blob = bytearray()
        n_rec = len(fetches) * a.copies_per_fetch * a.batches * a.records_per_batch
        wire = sum(len(f) for f in fetches)
        best = None
        for _ in range(3):
            g = C.c_void_p()
            assert lib.sgr_ingest_create(C.byref(g)) == 0
            st = N.sgr_ingest_stats()
            # the restore loop: poll -> decode -> (fold) -> mark folded; the first poll warms the buffers and is not timed
            dt = 0.0
            for i, data in enumerate(fetches):
                t0 = time.perf_counter()
                rc = lib.sgr_ingest_record_batches(g, 0, data, len(data), C.byref(st))
                if i:
                    dt += time.perf_counter() - t0
                assert rc == 0 and st.n_records == n_rec // len(fetches), (rc, lib.sgr_ingest_last_error(g))
                lib.sgr_ingest_mark_folded(g)
            lib.sgr_ingest_destroy(g)
            best = dt if best is None else min(best, dt)
        timed = n_rec * (len(fetches) - 1) // len(fetches)
        out[compression] = dict(records=timed, fetch_mb=wire / len(fetches) / 1e6, wire_bytes_per_record=wire / n_rec, seconds=best,
                                records_per_s=timed / best, wire_mb_per_s=wire * (len(fetches) - 1) / len(fetches) / 1e6 / best)
    # partition-parallel decode: the same fetches spread over P partitions, one call
    for threads in (1, 2, 4, 8):
        P = 8
        per = [fetches[1 + (i % (len(fetches) - 1))] for i in range(P)]   # lz4 fetches from the loop above
        g = C.c_void_p()
        assert lib.sgr_ingest_create(C.byref(g)) == 0
        parts = (C.c_int32 * P)(*range(P))
        ptrs = (C.c_void_p * P)(*[C.cast(C.c_char_p(d), C.c_void_p) for d in per])
        lens = (C.c_uint64 * P)(*[len(d) for d in per])
        best = None
        for rep in range(5):   # distinct partition numbers per repeat: nothing is a duplicate, staging buffers stay warm
            parts = (C.c_int32 * P)(*range(rep * P, rep * P + P))
            t0 = time.perf_counter()
            rc = lib.sgr_ingest_record_batches_mt(g, P, parts, ptrs, lens, threads, None)
            dt = time.perf_counter() - t0
            assert rc == 0
            lib.sgr_ingest_mark_folded(g)
            if rep:
                best = dt if best is None else min(best, dt)
        n = P * a.copies_per_fetch * a.batches * a.records_per_batch
        out[f"lz4 x{P} partitions, {threads} threads"] = dict(records=n, seconds=best, records_per_s=n / best)
        lib.sgr_ingest_destroy(g)
    for k, v in out.items():
        print(k, {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items()})


if __name__ == "__main__":
    main()
