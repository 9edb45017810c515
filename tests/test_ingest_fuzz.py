"""Memory-safety fuzz of the record-batch decoder: the decoder reads bytes that came off a network.

Builds surge_b200/csrc/ingest.cpp + tests/fuzz/ingest_fuzz_main.cpp with AddressSanitizer and UBSan (host-only code, no
CUDA) and runs a few thousand mutated inputs through it. Mutants are re-sealed with a fresh CRC-32C so that the damage
reaches the record parser and the lz4 decoder instead of stopping at the checksum.
"""
import os
import struct
import subprocess

import numpy as np
import pytest

from oracle import kafka_batch as K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "oracle", "_build")
BIN = os.path.join(OUT, "ingest_fuzz_asan")


def _build():
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(ROOT, "surge_b200", "csrc", "ingest.cpp"), os.path.join(ROOT, "tests", "fuzz", "ingest_fuzz_main.cpp")]
    if os.path.exists(BIN) and os.path.getmtime(BIN) >= max(os.path.getmtime(s) for s in srcs):
        return
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer",
           *srcs, "-o", BIN, "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("sanitizer build unavailable: " + r.stderr[-300:])


def _reseal(batch: bytearray) -> bytes:
    """Recompute batchLength-consistent CRC of a single (mutated) batch."""
    if len(batch) >= 21:
        batch[17:21] = struct.pack(">I", K.crc32c(bytes(batch[21:])))
    return bytes(batch)


def _corpus(rng, n_cases):
    ev = lambda s, extra=b"": struct.pack("<IIi", s % 3, s, s * 7) + extra  # noqa: E731
    seeds = []
    for comp in ("none", "lz4"):
        recs = [(d, f"agg-{d % 7}:{d}".encode(), ev(d, bytes(d % 40))) for d in range(30)]
        seeds.append(K.encode_record_batch(10, recs, compression=comp, headers=[(b"h", b"v"), (b"n", None)]))
        seeds.append(K.encode_record_batch(0, [(0, b"", b""), (1, None, ev(1)), (2, b"k", None)], compression=comp, producer_id=3, transactional=True))
    seeds.append(K.encode_control_batch(5, 3, K.ABORT))
    frames = [K.lz4_frame_compress(bytes(rng.integers(97, 101, 3000, dtype=np.uint8)), **kw)
              for kw in (dict(), dict(block_checksum=True, content_checksum=True, content_size=True))]
    cases = [(0, s) for s in seeds] + [(1, f) for f in frames]
    while len(cases) < n_cases:
        if rng.random() < 0.75:
            b = bytearray(seeds[int(rng.integers(0, len(seeds)))])
            style = rng.random()
            for _ in range(int(rng.integers(1, 4))):
                pos = int(rng.integers(21, len(b))) if style < 0.8 else int(rng.integers(0, len(b)))
                b[pos] = int(rng.integers(0, 256)) if rng.random() < 0.5 else b[pos] ^ (1 << int(rng.integers(0, 8)))
            if rng.random() < 0.15:       # lie about recordsCount / lastOffsetDelta
                struct.pack_into(">i", b, 57 if rng.random() < 0.5 else 23, int(rng.integers(-5, 1 << 31)))
            data = _reseal(b) if style < 0.8 else bytes(b)
            if rng.random() < 0.2:
                data = data[: int(rng.integers(0, len(data) + 1))]
            if rng.random() < 0.2:
                data = data + seeds[int(rng.integers(0, len(seeds)))]
            cases.append((0, data))
        else:
            f = bytearray(frames[int(rng.integers(0, len(frames)))])
            for _ in range(int(rng.integers(1, 4))):
                pos = int(rng.integers(0, len(f)))
                f[pos] = int(rng.integers(0, 256))
            if rng.random() < 0.3:
                f = f[: int(rng.integers(0, len(f) + 1))]
            if rng.random() < 0.5:
                # wrap the damaged frame as the records section of a batch, sealed, so the batch path decodes it too
                tail = struct.pack(">hiqqqhii", 3, 0, 0, 0, -1, -1, -1, 1) + bytes(f)
                cases.append((0, struct.pack(">qiib", 0, 9 + len(tail), 0, 2) + struct.pack(">I", K.crc32c(tail)) + tail))
            else:
                cases.append((1, bytes(f)))
    return cases


def _value_corpus(rng, n_cases):
    """Well-framed batches whose record VALUES are hostile: mutated JSON objects (kind 2) and mutated protobuf Events (kind 3)."""
    import json

    seeds_json = [json.dumps({"_type": "Inc", "aggregateId": "a", "by": 5, "seq": 7, "w": 1.5e10, "x": {"y": [1, 2, {"z": None}], "s": 'q"\\' + "é"}}).encode(),
                  json.dumps({"_type": "Big\u00e9", "v": -2**63, "pad": "x" * 70}, ensure_ascii=False).encode("utf-8"),
                  json.dumps({"_type": "Big\u00e9", "v": 12}, ensure_ascii=True).encode(),
                  b'{ "_type" : "Other" , "n" : [ ] , "o" : { } }']
    ev = struct.pack("<IIi", 1, 2, 3) + bytes(20)
    seeds_pb = [b"\x0a\x03abc\x12" + bytes([len(ev)]) + ev, b"\x12" + bytes([len(ev)]) + ev + b"\x18\x05\x25\x01\x02\x03\x04", b"\x12\x08" + ev[:8]]
    cases = []
    while len(cases) < n_cases:
        kind = 2 if rng.random() < 0.7 else 3
        v = bytearray((seeds_json if kind == 2 else seeds_pb)[int(rng.integers(0, 4 if kind == 2 else 3))])
        if rng.random() < 0.9:
            for _ in range(int(rng.integers(1, 4))):
                op = rng.random()
                pos = int(rng.integers(0, len(v)))
                if op < 0.4:
                    v[pos] = int(rng.integers(0, 256))
                elif op < 0.6:
                    v[pos] = int(rng.choice(list(b'{}[]",:\\-.eE0u')))
                elif op < 0.8:
                    del v[pos:pos + int(rng.integers(1, 6))]
                else:
                    v[pos:pos] = bytes(rng.integers(0, 256, int(rng.integers(1, 5)), dtype=np.uint8))
            if rng.random() < 0.2:
                v = v[: int(rng.integers(0, len(v) + 1))]
        cases.append((kind, K.encode_record_batch(0, [(0, b"key:1", bytes(v))], compression="lz4" if rng.random() < 0.3 else "none")))
    return cases


def test_decoder_survives_mutated_input_under_asan_ubsan(tmp_path):
    _build()
    rng = np.random.default_rng(2024)
    cases = _corpus(rng, 3000) + _value_corpus(rng, 2500)
    path = tmp_path / "corpus.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("<I", len(cases)))
        for kind, data in cases:
            f.write(struct.pack("<BI", kind, len(data)) + data)
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")
    r = subprocess.run([BIN, str(path)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-3000:])
    assert f"cases={len(cases)}" in r.stdout
    ok = int(r.stdout.split("ok=")[1].split()[0])
    refused = int(r.stdout.split("refused=")[1].split()[0])
    assert ok > 200 and refused > 1000      # both outcomes are exercised


TSAN_BIN = os.path.join(OUT, "ingest_mt_tsan")


def test_multi_threaded_decode_is_race_free_and_deterministic_under_tsan(tmp_path):
    """The parallel phases (per-partition decode, sharded id probing, slot publication, placement) under ThreadSanitizer,
    against the single-threaded result of the same polls."""
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(ROOT, "surge_b200", "csrc", "ingest.cpp"), os.path.join(ROOT, "tests", "fuzz", "ingest_mt_main.cpp")]
    if not (os.path.exists(TSAN_BIN) and os.path.getmtime(TSAN_BIN) >= max(os.path.getmtime(s) for s in srcs)):
        r = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", *srcs, "-o", TSAN_BIN, "-lpthread"], capture_output=True, text=True)
        if r.returncode != 0:
            pytest.skip("thread sanitizer build unavailable: " + r.stderr[-300:])
    rng = np.random.default_rng(31)
    ev = lambda s: struct.pack("<IIi", s % 3, s, s)  # noqa: E731
    nxt = {p: 0 for p in range(6)}
    polls = []
    for poll in range(6):
        fetches = []
        for p in range(6):
            for _ in range(1 + (poll + p) % 2):       # sometimes two fetches of one partition in a poll: they chain
                blob = bytearray()
                for _b in range(3):
                    n = int(rng.integers(300, 600))      # ~12k records per poll: the id probing really runs on several workers
                    recs = [(d, f"p{p}-k{int(rng.integers(0, 4000))}:{d}".encode(), ev(nxt[p] + d)) for d in range(n)]
                    blob += K.encode_record_batch(nxt[p], recs, compression="lz4" if (p + _b) % 2 else "none")
                    nxt[p] += n
                fetches.append((p, bytes(blob)))
        if poll == 4:
            fetches.append((0, b"\x00" * 40))          # a malformed fetch: the whole poll is refused by both
        polls.append(fetches)
    path = tmp_path / "polls.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("<I", len(polls)))
        for fetches in polls:
            f.write(struct.pack("<I", len(fetches)))
            for p, data in fetches:
                f.write(struct.pack("<iI", p, len(data)) + data)
    for threads in (4, 13):
        r = subprocess.run([TSAN_BIN, str(path), str(threads)], capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, TSAN_OPTIONS="halt_on_error=1"))
        assert r.returncode == 0, (r.stdout[-300:], r.stderr[-3000:])
        assert "polls=6" in r.stdout and "records=" in r.stdout
