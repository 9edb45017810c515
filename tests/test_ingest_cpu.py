"""Kafka RecordBatch decode and the per-partition offsets the lag gate needs (f2) — host logic, no GPU.

The product decoder is surge_b200/csrc/ingest.cpp behind the C ABI; the checker is the independent Python
encoder/decoder in oracle/kafka_batch.py. Byte-level parity with a real broker is UNPINNED (the reference holds no broker
bytes); what is pinned here are the published known-answer vectors of CRC-32C (RFC 3720 B.4) and xxHash32 and the LZ4
frame header checksum bytes.
"""
import ctypes as C
import struct

import numpy as np
import pytest

from oracle import kafka_batch as K
from surge_b200 import native as N
from surge_b200.ingest import Ingest, IngestError


@pytest.fixture(scope="module")
def lib():
    return N.load_library()


def _crc(lib, b: bytes, portable=False) -> int:
    buf = C.create_string_buffer(b, len(b)) if b else None
    return (lib.sgr_crc32c_portable if portable else lib.sgr_crc32c)(buf, len(b))


# RFC 3720 appendix B.4
CRC_VECTORS = [
    (b"\x00" * 32, 0x8A9136AA),
    (b"\xff" * 32, 0x62A8AB43),
    (bytes(range(32)), 0x46DD794E),
    (bytes(range(31, -1, -1)), 0x113FDB5C),
    (b"123456789", 0xE3069283),
    (b"", 0),
]


@pytest.mark.parametrize("data,want", CRC_VECTORS)
def test_crc32c_known_answers(lib, data, want):
    assert K.crc32c(data) == want
    assert _crc(lib, data) == want
    assert _crc(lib, data, portable=True) == want


def test_crc32c_hw_and_table_agree_on_unaligned_lengths(lib):
    rng = np.random.default_rng(7)
    blob = rng.integers(0, 256, 5000, dtype=np.uint8).tobytes()
    for start in range(0, 9):
        for n in (0, 1, 7, 8, 9, 63, 64, 65, 1000, 4991):
            d = blob[start:start + n]
            assert _crc(lib, d) == _crc(lib, d, portable=True) == K.crc32c(d)


def test_xxh32_known_answers(lib):
    def x(b, seed=0):
        buf = C.create_string_buffer(b, len(b)) if b else None
        return lib.sgr_xxh32(buf, len(b), seed)

    assert K.xxh32(b"") == x(b"") == 0x02CC5D05
    assert K.xxh32(b"abc") == x(b"abc") == 0x32D153FF
    # LZ4 frame header checksum byte = (xxh32(descriptor) >> 8) & 0xff: the two headers every lz4 tool writes
    assert (x(bytes([0x60, 0x40])) >> 8) & 0xFF == 0x82
    assert (x(bytes([0x64, 0x40])) >> 8) & 0xFF == 0xA7
    rng = np.random.default_rng(3)
    for n in (1, 3, 4, 15, 16, 17, 31, 32, 33, 100, 1000):
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        for seed in (0, 1, 0x9E3779B1):
            assert x(d, seed) == K.xxh32(d, seed)


def _lz4_decode(lib, frame: bytes, cap: int = 1 << 22):
    out = C.create_string_buffer(cap)
    n = C.c_uint64()
    rc = lib.sgr_lz4_frame_decode(C.create_string_buffer(frame, len(frame)), len(frame), out, cap, C.byref(n))
    return rc, out.raw[: n.value]


def _compressible(rng, n):
    words = [bytes(rng.integers(97, 123, int(rng.integers(2, 9)), dtype=np.uint8)) for _ in range(20)]
    out = bytearray()
    while len(out) < n:
        out += words[int(rng.integers(0, len(words)))]
    return bytes(out[:n])


@pytest.mark.parametrize("kw", [dict(), dict(block_checksum=True), dict(content_checksum=True), dict(content_size=True),
                                dict(block_checksum=True, content_checksum=True, content_size=True, block_code=5)])
def test_lz4_frames_round_trip(lib, kw):
    rng = np.random.default_rng(11)
    cases = [b"", b"a", b"abcd" * 3, b"\x00" * 100_000, _compressible(rng, 70_000), _compressible(rng, 200_001),
             rng.integers(0, 256, 3000, dtype=np.uint8).tobytes()]  # incompressible -> stored block
    for data in cases:
        frame = K.lz4_frame_compress(data, **kw)
        assert K.lz4_frame_decompress(frame) == data
        rc, got = _lz4_decode(lib, frame)
        assert rc == 0 and got == data
    # the frame of a long run is much smaller than the run: matches (incl. overlapping ones) were really emitted
    assert len(K.lz4_frame_compress(b"\x00" * 100_000)) < 1000


def test_lz4_hand_built_overlapping_match(lib):
    # token 0x1F: 1 literal, match length 15+4(+ext 1) = 20, offset 1 -> "a" * 21, then a literal-only tail
    block = bytes([0x1F, ord("a"), 0x01, 0x00, 0x01]) + bytes([0x50]) + b"bcdef"
    desc = bytes([0x60, 0x40])
    frame = struct.pack("<I", 0x184D2204) + desc + bytes([(K.xxh32(desc) >> 8) & 0xFF]) + struct.pack("<I", len(block)) + block + struct.pack("<I", 0)
    rc, got = _lz4_decode(lib, frame)
    assert rc == 0 and got == b"a" * 21 + b"bcdef"
    assert K.lz4_frame_decompress(frame) == got


def test_lz4_rejects_corruption(lib):
    data = _compressible(np.random.default_rng(5), 10_000)
    frame = bytearray(K.lz4_frame_compress(data, block_checksum=True, content_checksum=True))
    assert _lz4_decode(lib, bytes(frame))[0] == 0
    for pos in (0, 4, 6, 20, len(frame) - 2):
        bad = bytearray(frame)
        bad[pos] ^= 0x40
        assert _lz4_decode(lib, bytes(bad))[0] == N.SGR_ERR_INVALID
    assert _lz4_decode(lib, bytes(frame[:-9]))[0] == N.SGR_ERR_INVALID        # no end mark
    # a match that reaches before the start of the output
    block = bytes([0x0F, 0x05, 0x00])
    desc = bytes([0x60, 0x40])
    f2 = struct.pack("<I", 0x184D2204) + desc + bytes([(K.xxh32(desc) >> 8) & 0xFF]) + struct.pack("<I", len(block)) + block + struct.pack("<I", 0)
    assert _lz4_decode(lib, f2)[0] == N.SGR_ERR_INVALID
    rc, _ = _lz4_decode(lib, K.lz4_frame_compress(data), cap=100)
    assert rc == N.SGR_ERR_CAPACITY


# ----------------------------------------------------------------------------- record batches
def _event(type_, seq, by=0, extra=b""):
    return struct.pack("<IIi", type_, seq, by) + extra


def _batch_stream(rng, n_batches, n_keys, compression, base=0):
    out, off = bytearray(), base
    for _ in range(n_batches):
        n = int(rng.integers(1, 40))
        recs = []
        for d in range(n):
            k = int(rng.integers(0, n_keys))
            key = f"agg-{k}:{off + d}".encode() if rng.random() < 0.7 else f"agg-{k}".encode()
            recs.append((d, key, _event(int(rng.integers(0, 3)), off + d, int(rng.integers(-2**31, 2**31)), bytes(int(rng.integers(0, 45))))))
        out += K.encode_record_batch(off, recs, compression=compression, headers=[(b"aggregate_id", b"x"), (b"n", None)])
        off += n
    return bytes(out), off


@pytest.mark.parametrize("compression", ["none", "lz4"])
def test_decode_matches_restatement(compression):
    rng = np.random.default_rng(42)
    fetches = []
    nxt = {0: 0, 1: 1000}
    for _ in range(6):
        for p in (0, 1):
            buf, nxt[p] = _batch_stream(rng, 5, 50, compression, nxt[p])
            fetches.append((p, buf, []))
    want_recs, want_keys, want_next = K.read_committed_pack(fetches)
    ing = Ingest()
    total = 0
    for p, buf, _ in fetches:
        st = ing.record_batches(p, buf)
        assert st["n_trailing_bytes"] == 0 and st["n_bytes"] == len(buf)
        total += st["n_records"]
    got = ing.pending()
    assert total == len(want_recs) == len(got)
    assert np.array_equal(got, want_recs)
    assert ing.keys() == [k.decode() for k in want_keys]
    for p in (0, 1):
        assert ing.offsets(p) == (want_next[p], 0)
    ing.mark_folded()
    assert len(ing.pending()) == 0
    for p in (0, 1):
        assert ing.offsets(p) == (want_next[p], want_next[p])
    if compression == "lz4":
        s = ing.stats()
        assert 0 < s["n_compressed_bytes"] < s["n_decompressed_bytes"]


def test_flush_markers_null_values_and_key_up_to_colon():
    recs = [(0, b"", b""),                        # the producer's flush record: empty key, empty value
            (1, None, _event(0, 1, 5)),            # null key
            (2, b"acct:7", _event(0, 2, 5)),
            (3, b"acct", _event(1, 3, 2)),         # same aggregate as "acct:7"
            (4, b"acct2:1", None),                 # null value
            (5, "zażółć:1".encode(), _event(0, 1, 1))]
    ing = Ingest()
    st = ing.record_batches(3, K.encode_record_batch(100, recs))
    assert (st["n_records"], st["n_markers"], st["n_null_values"], st["n_new_keys"]) == (3, 2, 1, 2)
    assert ing.keys() == ["acct", "zażółć"]
    p = ing.pending()
    assert p[:, 8:16].view(np.uint64).ravel().tolist() == [0, 0, 1]
    assert p[:, 0:4].view(np.uint32).ravel().tolist() == [0, 1, 0]
    assert p[:, 4:8].view(np.uint32).ravel().tolist() == [2, 3, 1]
    assert p[:, 16:20].view(np.int32).ravel().tolist() == [5, 2, 1]
    assert not p[:, 20:].any()
    assert ing.offsets(3) == (106, 0)
    assert ing.offsets(9) == (0, 0)


def test_trailing_partial_batch_is_left_for_the_next_fetch():
    b1 = K.encode_record_batch(0, [(0, b"a", _event(0, 1, 1))])
    b2 = K.encode_record_batch(1, [(0, b"b", _event(0, 1, 1)), (1, b"a", _event(0, 2, 1))], compression="lz4")
    ing = Ingest()
    for cut in (5, 12, 30, len(b2) - 1):
        g = Ingest()
        st = g.record_batches(0, b1 + b2[:cut])
        assert (st["n_batches"], st["n_records"], st["n_bytes"], st["n_trailing_bytes"]) == (1, 1, len(b1), cut)
        assert g.offsets(0)[0] == 1
    st = ing.record_batches(0, b1 + b2)
    assert st["n_records"] == 3 and ing.offsets(0)[0] == 3


def test_refetch_after_restart_skips_duplicates():
    rng = np.random.default_rng(1)
    buf, end = _batch_stream(rng, 4, 10, "none")
    ing = Ingest()
    n = ing.record_batches(0, buf)["n_records"]
    st = ing.record_batches(0, buf)          # the same bytes again
    assert st["n_records"] == 0 and st["n_duplicates"] == n
    assert len(ing.pending()) == n and ing.offsets(0)[0] == end
    # a batch that straddles the position: only its tail is new
    tail = K.encode_record_batch(end - 1, [(0, b"x", _event(0, 1, 1)), (1, b"y", _event(0, 1, 1))])
    st = ing.record_batches(0, tail)
    assert (st["n_records"], st["n_duplicates"]) == (1, 1)
    assert ing.keys()[-1] == "y"


def test_read_committed_skips_aborted_transactions_and_control_batches():
    ev = lambda s: _event(0, s, 1)  # noqa: E731
    log = b"".join([
        K.encode_record_batch(0, [(0, b"a", ev(1)), (1, b"b", ev(1))], producer_id=7, transactional=True),     # aborted
        K.encode_record_batch(2, [(0, b"c", ev(1))], producer_id=8, transactional=True, compression="lz4"),    # committed
        K.encode_record_batch(3, [(0, b"a", ev(2))], producer_id=7, transactional=True),                       # aborted
        K.encode_control_batch(4, 7, K.ABORT),
        K.encode_control_batch(5, 8, K.COMMIT),
        K.encode_record_batch(6, [(0, b"a", ev(3))], producer_id=7, transactional=True),                       # new txn, committed
        K.encode_control_batch(7, 7, K.COMMIT),
        K.encode_record_batch(8, [(0, b"d", ev(1))]),                                                          # non-transactional
        K.encode_record_batch(9, [(0, b"e", ev(1))], producer_id=9, transactional=True),                       # aborted later in the log
        K.encode_control_batch(10, 9, K.ABORT),
    ])
    aborted = [(7, 0), (9, 9)]
    want, want_keys, want_next = K.read_committed_pack([(0, log, aborted)])
    ing = Ingest()
    ing.set_aborted(0, aborted)
    st = ing.record_batches(0, log)
    assert (st["n_control_batches"], st["n_aborted_batches"], st["n_aborted_records"], st["n_records"]) == (4, 3, 4, 3)
    assert ing.keys() == ["c", "a", "d"] == [k.decode() for k in want_keys]
    assert np.array_equal(ing.pending(), want)
    assert ing.offsets(0)[0] == 11 == want_next[0]
    # without the aborted list (read_uncommitted view) everything is data
    g = Ingest()
    assert g.record_batches(0, log)["n_records"] == 7


def test_malformed_input_fails_loudly_and_leaves_state_untouched():
    good = K.encode_record_batch(0, [(0, b"a", _event(0, 1, 1))])
    nxt = K.encode_record_batch(1, [(0, b"b", _event(0, 1, 1)), (1, b"c", _event(0, 2, 1))], compression="lz4")
    ing = Ingest()
    ing.record_batches(0, good)

    def refused(buf, code, fragment):
        with pytest.raises(IngestError) as ei:
            ing.record_batches(0, buf)
        assert ei.value.code == code and fragment in str(ei.value)
        assert len(ing.pending()) == 1 and ing.offsets(0)[0] == 1

    bad = bytearray(nxt)
    bad[70] ^= 1
    refused(good[:0] + bytes(bad), N.SGR_ERR_INVALID, "CRC-32C mismatch")
    refused(K.encode_record_batch(1, [(0, b"b", b"x" * 8)], magic=1), N.SGR_ERR_UNSUPPORTED, "message format v1")
    for codec in ("gzip", "snappy", "zstd"):
        refused(K.encode_record_batch(1, [(0, b"b", _event(0, 1))], compression=codec), N.SGR_ERR_UNSUPPORTED, codec)
    refused(K.encode_record_batch(1, [(0, b"b", b"short")]), N.SGR_ERR_INVALID, "packed event value of 5 bytes")
    refused(K.encode_record_batch(1, [(0, b"b", bytes(57))]), N.SGR_ERR_INVALID, "packed event value of 57 bytes")
    # a good batch followed by a bad one: nothing of the call is kept
    refused(nxt + bytes(bad), N.SGR_ERR_INVALID, "CRC-32C mismatch")
    # records section shorter than recordsCount says (CRC re-computed so that only the structure is wrong)
    body = K.encode_record(0, b"b", _event(0, 1))
    tail = struct.pack(">hiqqqhii", 0, 1, 0, 1, -1, -1, -1, 2) + body
    lying = struct.pack(">qiib", 1, 9 + len(tail), 0, 2) + struct.pack(">I", K.crc32c(tail)) + tail
    refused(lying, N.SGR_ERR_INVALID, "record 1")
    st = ing.record_batches(0, nxt)
    assert st["n_records"] == 2 and ing.offsets(0)[0] == 3


def test_varint_extremes_round_trip():
    for v in (0, 1, -1, 63, 64, -64, -65, 2**31 - 1, -2**31):
        assert K.read_varint(K.varint(v), 0)[0] == v
    big_delta = 2**31 - 1
    b = K.encode_record_batch(5, [(0, b"k", _event(0, 1, 1)), (big_delta, b"k", _event(0, 2, 1))])
    ing = Ingest()
    assert ing.record_batches(0, b)["n_records"] == 2
    assert ing.offsets(0)[0] == 5 + big_delta + 1


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_multi_fetch_call_equals_single_calls(threads):
    rng = np.random.default_rng(77)
    nxt = {p: 100 * p for p in range(4)}
    fetches = []
    for rnd in range(3):                       # several fetches per partition inside ONE call: they chain
        for p in range(4):
            buf, nxt[p] = _batch_stream(rng, 3, 200, "lz4" if (p + rnd) % 2 else "none", nxt[p])
            if rnd == 1 and p == 2:
                buf = buf + buf[:40]           # a trailing partial batch
            fetches.append((p, buf))
    fetches.append((9, b""))                  # an empty fetch
    serial = Ingest()
    want_stats = [serial.record_batches(p, b) for p, b in fetches]
    par = Ingest()
    got_stats = par.record_batches_mt(fetches, threads=threads)
    assert got_stats == want_stats
    assert np.array_equal(par.pending(), serial.pending())
    assert par.keys() == serial.keys()
    for p in (0, 1, 2, 3, 9):
        assert par.offsets(p) == serial.offsets(p)
    assert par.stats() == serial.stats()
    # and again on warm staging buffers
    more = []
    for p in range(4):
        buf, nxt[p] = _batch_stream(rng, 2, 300, "lz4", nxt[p])
        more.append((p, buf))
    assert par.record_batches_mt(more, threads=threads) == [serial.record_batches(p, b) for p, b in more]
    assert np.array_equal(par.pending(), serial.pending()) and par.keys() == serial.keys()


def test_sharded_id_probing_keeps_first_seen_order():
    """Polls big enough for the id dictionary to be probed by several workers (one per ~1k ids): indices must still be
    first-seen ranks, provisional ids that recur inside the poll (same and other fetches) must resolve to one index."""
    rng = np.random.default_rng(123)
    nxt = {p: 0 for p in range(5)}
    serial, par = Ingest(), Ingest()
    for poll in range(3):
        fetches = []
        for p in range(5):
            blob = bytearray()
            for _ in range(4):
                n = 400
                recs = [(d, f"id-{int(rng.integers(0, 3000 * (poll + 1)))}:{d}".encode() if d % 50 else f"hot-{p % 2}".encode(),
                         _event(0, nxt[p] + d, 1)) for d in range(n)]
                blob += K.encode_record_batch(nxt[p], recs)
                nxt[p] += n
            fetches.append((p, bytes(blob)))
        for p, b in fetches:
            serial.record_batches(p, b)
        par.record_batches_mt(fetches, threads=7)
        assert np.array_equal(par.pending(), serial.pending())
        assert par.keys() == serial.keys()
        aggs = par.pending()[:, 8:16].copy().view(np.uint64).ravel()
        first_seen = {}
        for a in aggs.tolist():
            first_seen.setdefault(a, len(first_seen))
        if poll == 0:
            assert all(a == rank for a, rank in first_seen.items())       # dense indices ARE first-seen ranks
        if poll == 1:
            serial.mark_folded(); par.mark_folded()


def test_multi_fetch_call_is_all_or_nothing():
    rng = np.random.default_rng(78)
    good0, _ = _batch_stream(rng, 3, 50, "lz4", 0)
    good1, _ = _batch_stream(rng, 3, 50, "none", 0)
    bad = bytearray(good1)
    bad[len(bad) // 2] ^= 0x10
    ing = Ingest()
    ing.record_batches(5, K.encode_record_batch(0, [(0, b"seed", _event(0, 1, 1))]))
    with pytest.raises(IngestError) as ei:
        ing.record_batches_mt([(0, good0), (1, bytes(bad)), (2, good1)], threads=3)
    assert "partition 1" in str(ei.value) and "CRC-32C" in str(ei.value)
    assert ing.keys() == ["seed"] and len(ing.pending()) == 1
    assert ing.offsets(0) == (0, 0) and ing.offsets(2) == (0, 0)
    st = ing.record_batches_mt([(0, good0), (1, good1), (2, good1)], threads=3)
    assert sum(s["n_records"] for s in st) == len(ing.pending()) - 1


def test_id_dictionary_refuses_to_wrap_its_32_bit_fields():
    """ADVICE r1 (ingest.cpp ShardedDict): past 2^31 ids / 4 GiB of id bytes the dictionary would wrap silently and fold events
    into the wrong aggregates. The bound is checked before anything is applied; lowered here to make it reachable."""
    rng = np.random.default_rng(5)
    ing = Ingest()
    first, nxt = _batch_stream(rng, 2, 20, "none", 0)
    ing.record_batches(0, first)
    keys_before, pending_before, offs_before = ing.keys(), len(ing.pending()), ing.offsets(0)
    ing.set_dictionary_limits(len(keys_before) + 5, 1 << 32)          # room for 5 more ids; the next fetch carries dozens of records
    more, _ = _batch_stream(rng, 3, 500, "lz4", nxt)
    with pytest.raises(IngestError) as ei:
        ing.record_batches(0, more)
    assert ei.value.code == N.SGR_ERR_CAPACITY and "id dictionary full" in str(ei.value)
    assert ing.keys() == keys_before and len(ing.pending()) == pending_before and ing.offsets(0) == offs_before   # nothing applied
    ing.set_dictionary_limits(1 << 31, sum(len(k) for k in keys_before) + 16)   # the byte bound
    with pytest.raises(IngestError):
        ing.record_batches(0, more)
    ing.set_dictionary_limits(1 << 31, 1 << 32)
    ing.record_batches(0, more)                                      # and with the real bounds the same fetch goes through
    assert len(ing.pending()) > pending_before


def test_pending_log_moves_to_a_caller_supplied_allocator(lib):
    """sgr_fold_ingested installs page-locked memory this way; here: counting wrappers around libc malloc/free."""
    libc = C.CDLL(None)
    libc.malloc.restype, libc.malloc.argtypes = C.c_void_p, [C.c_size_t]
    libc.free.restype, libc.free.argtypes = None, [C.c_void_p]
    calls = {"alloc": 0, "free": 0}

    @C.CFUNCTYPE(C.c_void_p, C.c_size_t)
    def my_alloc(n):
        calls["alloc"] += 1
        return libc.malloc(n)

    @C.CFUNCTYPE(None, C.c_void_p)
    def my_free(p):
        calls["free"] += 1
        libc.free(p)

    rng = np.random.default_rng(5)
    ing = Ingest()
    buf, end = _batch_stream(rng, 5, 30, "none")
    ing.record_batches(0, buf)
    before = ing.pending()
    assert lib.sgr_ingest_set_allocator(ing.handle, my_alloc, my_free) == 0
    assert calls == {"alloc": 1, "free": 0} and np.array_equal(ing.pending(), before)      # content carried over
    assert lib.sgr_ingest_set_allocator(ing.handle, my_alloc, my_free) == 0 and calls["alloc"] == 1   # idempotent
    big, _ = _batch_stream(rng, 200, 30, "lz4", end)       # forces growth through the new allocator
    ing.record_batches(0, big)
    assert calls["alloc"] >= 2 and calls["free"] == calls["alloc"] - 1
    assert np.array_equal(ing.pending()[: len(before)], before)
    ing.close()
    assert calls["free"] == calls["alloc"]


def test_state_topic_tombstones_become_events():
    """A compacted state topic: a keyed record with a null value deletes the key (SurgeModel.scala:62-64). By default such
    records are dropped (events topics never hold them); in state-topic mode they become events of the chosen type."""
    snap = lambda c, v: struct.pack("<IIii", 0, 0, c, v)  # noqa: E731   type 0 = snapshot {count, version}
    recs = [(0, b"a", snap(1, 1)), (1, b"b", snap(7, 2)), (2, b"a", None), (3, None, None), (4, b"c", None)]
    batch = K.encode_record_batch(0, recs, compression="lz4")
    dropped = Ingest()
    st = dropped.record_batches(0, batch)
    assert (st["n_records"], st["n_null_values"], st["n_markers"]) == (2, 2, 1) and dropped.keys() == ["a", "b"]
    ing = Ingest()
    ing.set_null_value_type(1)
    st = ing.record_batches(0, batch)
    assert (st["n_records"], st["n_null_values"], st["n_markers"]) == (4, 2, 1) and ing.keys() == ["a", "b", "c"]
    p = ing.pending()
    assert p[:, 0:4].view(np.uint32).ravel().tolist() == [0, 0, 1, 1]
    assert p[:, 8:16].view(np.uint64).ravel().tolist() == [0, 1, 0, 2]
    assert not p[2:, 16:].any() and not p[2:, 4:8].any()
    with pytest.raises(IngestError):
        ing.set_null_value_type(16)
    ing.set_null_value_type(-1)
    assert ing.record_batches(1, batch)["n_records"] == 2


# ----------------------------------------------------------------------------- multilanguage protobuf framing
def _pb_class(name):
    """protobuf.State / protobuf.Event of modules/multilanguage-protocol/src/main/protobuf/multilanguage-protocol.proto:7-20, built
    at run time with the protobuf runtime that ships in this image (an independent implementation of the wire format)."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

    fdp = descriptor_pb2.FileDescriptorProto(name=f"{name}.proto", package="surge.multilanguage", syntax="proto3")
    m = fdp.message_type.add(name=name)
    m.field.add(name="aggregateId", number=1, type=descriptor_pb2.FieldDescriptorProto.TYPE_STRING, label=descriptor_pb2.FieldDescriptorProto.LABEL_OPTIONAL)
    m.field.add(name="payload", number=2, type=descriptor_pb2.FieldDescriptorProto.TYPE_BYTES, label=descriptor_pb2.FieldDescriptorProto.LABEL_OPTIONAL)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fdp)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName(f"surge.multilanguage.{name}"))


def test_multilanguage_protobuf_framing_matches_the_protobuf_runtime():
    from surge_b200 import formats as F

    State = _pb_class("State")
    rng = np.random.default_rng(8)
    for aid, n in [("", 0), ("a", 0), ("", 5), ("agg-1", 8), ("zażółć", 127), ("x" * 200, 128), ("id", 20000)]:
        payload = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        want = State(aggregateId=aid, payload=payload).SerializeToString()
        assert F.multilanguage_proto(aid, payload) == want
        assert F.parse_multilanguage_proto(want) == (aid, payload)
        back = State()
        back.ParseFromString(F.multilanguage_proto(aid, payload))
        assert (back.aggregateId, back.payload) == (aid, payload)


def test_events_wrapped_in_protobuf_are_unwrapped_natively():
    Event = _pb_class("Event")
    evs = [("a", _event(0, 1, 5)), ("b", _event(1, 2, 7, bytes(range(40)))), ("a", _event(2, 3, 0))]
    recs = [(d, aid.encode(), Event(aggregateId=aid, payload=p).SerializeToString()) for d, (aid, p) in enumerate(evs)]
    # unknown fields and a field order no canonical writer produces must still parse (protobuf semantics)
    odd = b"\x18\x07" + b"\x12\x0c" + _event(0, 4, 9) + b"\x0a\x01a" + b"\x25\x01\x02\x03\x04"
    recs.append((3, b"a", odd))
    plain = Ingest()
    plain.record_batches(0, K.encode_record_batch(0, [(d, aid.encode(), p) for d, (aid, p) in enumerate(evs)] + [(3, b"a", _event(0, 4, 9))]))
    ing = Ingest()
    ing.set_value_framing(1)
    st = ing.record_batches(0, K.encode_record_batch(0, recs, compression="lz4"))
    assert st["n_records"] == 4 and np.array_equal(ing.pending(), plain.pending()) and ing.keys() == ["a", "b"]
    for bad, why in [(b"\x12\x7f" + bytes(5), "not a protobuf Event"), (b"\x0a\x01a", "packed event value of 0 bytes"), (b"\x13", "not a protobuf Event")]:
        with pytest.raises(IngestError) as ei:
            ing.record_batches(0, K.encode_record_batch(4, [(0, b"k", bad)]))
        assert why in str(ei.value)
    with pytest.raises(IngestError):
        ing.set_value_framing(7)


# ----------------------------------------------------------------------------- pinned against real implementations in the image
def test_lz4_frames_from_liblz4_decode_and_ours_decode_with_liblz4(lib):
    """pyarrow's "lz4" codec is liblz4's frame API — and writes exactly the header Kafka's producer does (FLG 0x60, BD 0x40,
    HC 0x82). Frames made by the real library must decode through the product decoder and the Python restatement; frames made
    by the restatement's compressor must decode through the real library."""
    pa = pytest.importorskip("pyarrow")
    codec = pa.Codec("lz4")
    rng = np.random.default_rng(21)
    cases = [b"", b"x", b"hello hello hello hello hello hello hello", b"\x00" * 300_000, _compressible(rng, 1_000_000),
             rng.integers(0, 256, 70_000, dtype=np.uint8).tobytes(), _compressible(rng, 65_536), _compressible(rng, 65_537)]
    for data in cases:
        frame = codec.compress(data, asbytes=True)
        if data:
            # small inputs come out with Kafka's exact header; multi-block ones with linked blocks (FLG 0x40), which
            # exercises matches that reach back into the previous block
            assert frame[:4] == bytes.fromhex("04224d18") and frame[4] in (0x60, 0x40) and frame[5] == 0x40
        rc, got = _lz4_decode(lib, frame)
        assert rc == 0 and got == data
        assert K.lz4_frame_decompress(frame) == data
        ours = K.lz4_frame_compress(data)
        assert codec.decompress(ours, decompressed_size=len(data), asbytes=True) == data
    # a record batch whose records section was compressed by liblz4 (what a broker hands out) decodes like our own
    recs = [(d, f"k{d % 5}:{d}".encode(), _event(d % 3, d, d)) for d in range(400)]
    plain = K.encode_record_batch(0, recs)
    body = codec.compress(plain[61:], asbytes=True)
    tail = struct.pack(">h", 3) + plain[23:61] + body
    batch = struct.pack(">qiib", 0, 9 + len(tail), 0, 2) + struct.pack(">I", K.crc32c(tail)) + tail
    a, b = Ingest(), Ingest()
    a.record_batches(0, plain)
    assert b.record_batches(0, batch)["n_records"] == 400
    assert np.array_equal(a.pending(), b.pending()) and a.keys() == b.keys()


def test_xxh32_matches_the_xxhash_library(lib):
    xxhash = pytest.importorskip("xxhash")
    rng = np.random.default_rng(22)
    for n in list(range(0, 40)) + [63, 64, 65, 1000, 4096, 100_003]:
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        for seed in (0, 1, 0xDEADBEEF):
            want = xxhash.xxh32(d, seed=seed).intdigest()
            buf = C.create_string_buffer(d, len(d)) if d else None
            assert lib.sgr_xxh32(buf, len(d), seed) == want == K.xxh32(d, seed)


def test_kafka_varints_are_protobuf_zigzag_varints():
    """ByteUtils.writeVarint / writeVarlong (kafka-clients) are protobuf's sint32 / sint64 encoding; the protobuf runtime in the
    image pins the restatement's encoder, whose output the product decoder parses in every other test of this file."""
    enc = pytest.importorskip("google.protobuf.internal.encoder")
    wf = pytest.importorskip("google.protobuf.internal.wire_format")
    rng = np.random.default_rng(4)
    ints = [0, 1, -1, 63, 64, -64, -65, 127, 128, 300, 2**31 - 1, -2**31] + rng.integers(-2**31, 2**31, 200).tolist()
    for v in ints:
        assert K.varint(int(v)) == enc._VarintBytes(wf.ZigZagEncode(int(v)))
        assert K.read_varint(K.varint(int(v)), 0)[0] == int(v)
    longs = [2**63 - 1, -2**63, 2**40, -2**40] + rng.integers(-2**62, 2**62, 200).tolist()
    for v in ints + longs:
        assert K.varlong(int(v)) == enc._VarintBytes(wf.ZigZagEncode(int(v)))


# ----------------------------------------------------------------------------- JSON-valued events (play-json style)
import json  # noqa: E402

COUNTER_JSON = [("surge.core.TestBoundedContext.CountIncremented", 0, [("incrementBy", N.JSON_I32, 16), ("sequenceNumber", N.JSON_I32, 4)]),
                ("surge.core.TestBoundedContext.CountDecremented", 1, [("decrementBy", N.JSON_I32, 16), ("sequenceNumber", N.JSON_I32, 4)]),
                ("surge.core.TestBoundedContext.NoOpEvent", 2, [("sequenceNumber", N.JSON_I32, 4)])]


def _json_ingest(unknown_type=3):
    ing = Ingest()
    ing.set_json_packer("_type", COUNTER_JSON, unknown_type=unknown_type)
    ing.set_value_framing(N.VALUE_JSON)
    return ing


def test_json_events_pack_like_hand_packed_ones():
    """The reference's own test model writes Json.toJson(evt) (core TestBoundedContext.scala:159-161); whatever the member
    order or spacing, the packed record must equal the hand-packed one."""
    rng = np.random.default_rng(55)
    recs_json, recs_packed = [], []
    for d in range(300):
        t = int(rng.integers(0, 4))
        by, seq = int(rng.integers(-2**31, 2**31)), int(rng.integers(0, 2**31))
        aid = f"agg-{int(rng.integers(0, 40))}"
        name = ["CountIncremented", "CountDecremented", "NoOpEvent", "SomethingElse"][t]
        obj = {"_type": f"surge.core.TestBoundedContext.{name}", "aggregateId": aid, "sequenceNumber": seq}
        if t == 0:
            obj["incrementBy"] = by
        if t == 1:
            obj["decrementBy"] = by
        items = list(obj.items()) + [("extra", {"nested": [1, "two", {"x": None}], "s": 'a"b\\'}), ("flag", True)]
        rng.shuffle(items)                                        # member order must not matter
        text = json.dumps(dict(items), separators=((",", ":") if d % 2 else (", ", " : ")), ensure_ascii=bool(d % 3))
        recs_json.append((d, f"{aid}:{seq}".encode(), text.encode("utf-8")))
        packed = struct.pack("<II", t, seq if t < 3 else 0) + (struct.pack("<i", by) if t < 2 else b"")
        recs_packed.append((d, f"{aid}:{seq}".encode(), packed))
    a, b = _json_ingest(), Ingest()
    a.record_batches(0, K.encode_record_batch(0, recs_json, compression="lz4"))
    b.record_batches(0, K.encode_record_batch(0, recs_packed))
    assert np.array_equal(a.pending(), b.pending()) and a.keys() == b.keys()


def test_json_values_numbers_escapes_and_errors():
    f64 = Ingest()
    f64.set_json_packer("t", [("Upd\u00e9", 1, [("newBalance", N.JSON_F64, 32), ("big", N.JSON_I64, 40)])])    # class name "Updé"
    f64.set_value_framing(N.VALUE_JSON)
    cases = [0.1, -0.0, 1e300, 5e-324, 2.2250738585072014e-308, 123456789.12345679, 1.0]
    # ensure_ascii=True writes the class name as "Upd\\u00e9" inside the JSON text: it must match after unescaping
    recs = [(d, b"k", json.dumps({"t": "Upd\u00e9", "newBalance": v, "big": -2**63 + d}, ensure_ascii=bool(d % 2)).encode("utf-8")) for d, v in enumerate(cases)]
    assert b"\\u00e9" in recs[1][2] and "Upd\u00e9".encode("utf-8") in recs[0][2]
    f64.record_batches(0, K.encode_record_batch(0, recs))
    p = f64.pending()
    assert p[:, 32:40].copy().view("<f8").ravel().tolist() == cases
    assert struct.pack("<d", float(p[1, 32:40].copy().view("<f8")[0])) == struct.pack("<d", -0.0)
    assert p[:, 40:48].copy().view("<i8").ravel().tolist() == [-2**63 + d for d in range(len(cases))]
    assert p[:, 0:4].copy().view("<u4").ravel().tolist() == [1] * len(cases)

    ing = _json_ingest(unknown_type=-1)

    def refused(value: bytes, why: str):
        with pytest.raises(IngestError) as ei:
            ing.record_batches(0, K.encode_record_batch(0, [(0, b"k", value)]))
        assert why in str(ei.value), str(ei.value)
        assert len(ing.pending()) == 0

    T = "surge.core.TestBoundedContext.CountIncremented"
    refused(b'{"_type":"%s","incrementBy":1.5,"sequenceNumber":1}' % T.encode(), "fraction or an exponent")
    refused(b'{"_type":"%s","incrementBy":2147483648,"sequenceNumber":1}' % T.encode(), "does not fit an Int")
    refused(b'{"_type":"%s","incrementBy":"1","sequenceNumber":1}' % T.encode(), "missing or not a number")
    refused(b'{"_type":"%s","sequenceNumber":1}' % T.encode(), "missing or not a number")
    refused(b'{"_type":"nope","sequenceNumber":1}', "unknown event class")
    refused(b'{"sequenceNumber":1}', "discriminator member is missing")
    refused(b'[1,2]', "not a JSON object")
    refused(b'{"_type":"%s","incrementBy":1,"sequenceNumber":1} x' % T.encode(), "bytes after the JSON object")
    refused(b'{"_type":"%s","incrementBy":01,"sequenceNumber":1}' % T.encode(), "expected")
    refused(b'{"_type":"%s","incrementBy":1,"sequenceNumber":1' % T.encode(), "expected")
    refused(b'{"_type":"%s" "incrementBy":1}' % T.encode(), "expected")
    refused(b'{"a":"unterminated', "unterminated string")
    # a later duplicate member wins (JsObject semantics), nested look-alikes are not members
    ok = b'{"x":{"_type":"nope","incrementBy":9},"_type":"nope","_type":"%s","incrementBy":7,"incrementBy":8,"sequenceNumber":2}' % T.encode()
    ing.record_batches(0, K.encode_record_batch(0, [(0, b"k", ok)]))
    assert ing.pending()[0, 16:20].copy().view("<i4")[0] == 8 and ing.pending()[0, 4:8].copy().view("<u4")[0] == 2
    with pytest.raises(IngestError):
        Ingest().set_value_framing(N.VALUE_JSON)                  # no packer registered
    with pytest.raises(IngestError):
        Ingest().set_json_packer("_type", [("A", 0, [("x", N.JSON_I32, 8)])])   # would overwrite the aggregate index


def test_bank_account_json_events_pack_like_the_binary_formatting():
    """surge-docs BankAccountSurgeModel.scala:30-32 writes Json.toJson(evt)(Json.format[BankAccountEvent]); UUID and String
    members land in the packed record exactly as surge_b200/formats.py packs them by hand."""
    import uuid as _uuid

    from surge_b200 import formats as F

    spec = [("docs.command.BankAccountCreated", 0, [("accountNumber", N.JSON_UUID, 16), ("balance", N.JSON_F64, 32),
                                                     ("accountOwner", N.JSON_PSTR, 40, 16), ("securityCode", N.JSON_PSTR, 56, 8)]),
            ("docs.command.BankAccountUpdated", 1, [("accountNumber", N.JSON_UUID, 16), ("newBalance", N.JSON_F64, 32)])]
    ing = Ingest()
    ing.set_json_packer("_type", spec)
    ing.set_value_framing(N.VALUE_JSON)
    rng = np.random.default_rng(3)
    recs, want = [], []
    for d in range(60):
        acct = str(_uuid.UUID(int=int(rng.integers(0, 2**63)) << 64 | int(rng.integers(0, 2**63))))
        if d % 3:
            owner, code, bal = ["Jane Doe", "Zo\\u00eb", "", "x" * 15][d % 4], ["1234", "", "abcdefg"][d % 3], float(rng.integers(0, 10**6)) / 8
            obj = {"_type": "docs.command.BankAccountCreated", "accountNumber": acct.upper() if d % 5 == 0 else acct, "accountOwner": owner, "securityCode": code, "balance": bal}
            want.append(F.bank_created_record(d % 7, 0, acct, owner, code, bal))
        else:
            bal = [0.0, -0.0, 1e-3, 12345.678][d % 4]
            obj = {"_type": "docs.command.BankAccountUpdated", "accountNumber": acct, "newBalance": bal}
            want.append(F.bank_updated_record(d % 7, 0, acct, bal))
        recs.append((d, f"k{d % 7}".encode(), json.dumps(obj, ensure_ascii=bool(d % 2)).encode("utf-8")))
    ing.record_batches(0, K.encode_record_batch(0, recs))
    got = ing.pending()
    assert [bytes(r) for r in got] == want
    for bad, why in [({"_type": "docs.command.BankAccountUpdated", "accountNumber": "not-a-uuid", "newBalance": 1.0}, "8-4-4-4-12"),
                     ({"_type": "docs.command.BankAccountUpdated", "accountNumber": "0000000g-0000-0000-0000-000000000000", "newBalance": 1.0}, "non-hex"),
                     ({"_type": "docs.command.BankAccountCreated", "accountNumber": str(_uuid.UUID(int=1)), "accountOwner": "x" * 16, "securityCode": "", "balance": 1.0}, "does not fit"),
                     ({"_type": "docs.command.BankAccountCreated", "accountNumber": str(_uuid.UUID(int=1)), "accountOwner": 5, "securityCode": "", "balance": 1.0}, "not a string")]:
        with pytest.raises(IngestError) as ei:
            ing.record_batches(1, K.encode_record_batch(0, [(0, b"k", json.dumps(bad).encode())]))
        assert why in str(ei.value)
    with pytest.raises(IngestError):
        Ingest().set_json_packer("_type", [("A", 0, [("s", N.JSON_PSTR, 56, 12)])])      # runs past the record


def test_json_state_topic_snapshots_and_tombstones():
    """The reference test model writes its STATE topic as Json.toJson(agg) = {"aggregateId":..,"count":..,"version":..}
    (core TestBoundedContext.scala:151-157), no discriminator, null = deleted: one registered class + the tombstone type."""
    ing = Ingest()
    ing.set_json_packer("", [("State", 0, [("count", N.JSON_I32, 16), ("version", N.JSON_I32, 20)])])
    ing.set_value_framing(N.VALUE_JSON)
    ing.set_null_value_type(1)
    recs = [(0, b"a", b'{"aggregateId":"a","count":4,"version":4}'), (1, b"b", b'{"aggregateId":"b","count":-7,"version":2}'), (2, b"a", None),
            (3, b"a", b'{"version":9,"count":1,"aggregateId":"a"}')]
    ing.record_batches(0, K.encode_record_batch(0, recs))
    p = ing.pending()
    assert p[:, 0:4].copy().view("<u4").ravel().tolist() == [0, 0, 1, 0]
    assert p[:, 16:24].copy().view("<i4").reshape(-1, 2).tolist() == [[4, 4], [-7, 2], [0, 0], [1, 9]]
    assert ing.keys() == ["a", "b"]
    with pytest.raises(IngestError):
        Ingest().set_json_packer("", [("A", 0, []), ("B", 1, [])])
