"""Property test (hypothesis) of the record-batch decoder against the Python restatement (oracle/kafka_batch.py): arbitrary
mixes of plain / transactional / control batches, committed and aborted transactions, flush markers and null values,
lz4 or not, delivered as fetches cut at ARBITRARY byte positions (a fetch may end inside a batch; the next one starts at
the first undecoded batch) and occasionally re-delivered from an earlier batch (restart). No GPU."""
import struct

import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import kafka_batch as K
from surge_b200.ingest import Ingest

keys = st.sampled_from([b"a", b"b:1", b"b:2", b"c", b"dd:x:y", "é:1".encode(), b"", None])
values = st.one_of(st.none(), st.integers(0, 40).map(lambda n: struct.pack("<II", n % 3, n) + bytes(range(n))))
record = st.tuples(keys, values)
batch = st.fixed_dictionaries({
    "records": st.lists(record, min_size=1, max_size=6),
    "compression": st.sampled_from(["none", "lz4"]),
    "pid": st.sampled_from([-1, -1, 5, 6]),            # -1: not transactional
    "end": st.sampled_from(["open", "commit", "abort"]),  # what follows a transactional batch
})


def build_log(batches):
    """-> (bytes, batch boundaries, aborted [(pid, first_offset)])"""
    out, bounds, aborted = bytearray(), [0], []
    off = 0
    open_first = {}
    for b in batches:
        recs = [(d, k, v) for d, (k, v) in enumerate(b["records"])]
        txn = b["pid"] >= 0
        if txn:
            open_first.setdefault(b["pid"], off)
        out += K.encode_record_batch(off, recs, compression=b["compression"], producer_id=b["pid"], transactional=txn)
        bounds.append(len(out))
        off += len(recs)
        if txn and b["end"] != "open":
            if b["end"] == "abort":
                aborted.append((b["pid"], open_first[b["pid"]]))
            out += K.encode_control_batch(off, b["pid"], K.ABORT if b["end"] == "abort" else K.COMMIT)
            bounds.append(len(out))
            off += 1
            del open_first[b["pid"]]
    return bytes(out), bounds, aborted


@settings(max_examples=250, deadline=None, derandomize=True)
@given(st.lists(batch, min_size=1, max_size=8), st.lists(st.integers(0, 10_000), min_size=0, max_size=6), st.booleans())
def test_decoder_equals_restatement_under_arbitrary_fetch_cuts(batches, cuts, restart):
    log, bounds, aborted = build_log(batches)
    # fetch boundaries: arbitrary byte positions; each fetch starts at the first batch the previous one did not complete
    cut_at = sorted({c % (len(log) + 1) for c in cuts} | {len(log)})
    fetches, start = [], 0
    for c in cut_at:
        if c <= start:
            continue
        fetches.append(log[start:c])
        start = max(b for b in bounds if b <= c)       # decoded whole batches only
    if start < len(log):
        fetches.append(log[start:])
    if restart and len(bounds) > 2:                      # re-delivery from an earlier batch boundary
        fetches.append(log[bounds[len(bounds) // 2]:])
    feed = [(0, f, aborted) for f in fetches]
    # every fetch response of a read_committed consumer carries the aborted transactions that overlap it; announcing all of
    # them once, before the first fetch, is equivalent for the restatement and the decoder alike
    want_recs, want_keys, want_next = K.read_committed_pack([(0, feed[0][1], aborted)] + [(0, f, []) for _, f, _ in feed[1:]])
    ing = Ingest()
    ing.set_aborted(0, aborted)
    for _, f, _ in feed:
        ing.record_batches(0, f)
    got = ing.pending()
    assert got.shape == want_recs.shape and np.array_equal(got, want_recs)
    assert ing.keys() == [k.decode() for k in want_keys]
    assert ing.offsets(0)[0] == want_next.get(0, 0)


# ----------------------------------------------------------------------------- JSON values: the native scanner vs Python's json module
import json  # noqa: E402

from surge_b200 import native as N  # noqa: E402

json_leaf = st.one_of(st.none(), st.booleans(), st.integers(-2**53, 2**53), st.floats(allow_nan=False, allow_infinity=False), st.text(max_size=12))
json_tree = st.recursive(json_leaf, lambda kids: st.one_of(st.lists(kids, max_size=4), st.dictionaries(st.text(max_size=6), kids, max_size=4)), max_leaves=12)


@settings(max_examples=300, deadline=None, derandomize=True)
@given(st.dictionaries(st.text(max_size=8).filter(lambda k: k not in ("_t", "by", "seq", "w", "big")), json_tree, max_size=6),
       st.integers(-2**31, 2**31 - 1), st.integers(0, 2**31 - 1), st.floats(allow_nan=False, allow_infinity=False), st.integers(-2**63, 2**63 - 1),
       st.booleans(), st.booleans(), st.randoms(use_true_random=False))
def test_json_scanner_accepts_what_python_writes_and_finds_the_members(extra, by, seq, w, big, ascii_only, spaced, rnd):
    """Any JSON object Python's json module can write — arbitrary nesting, escapes, unicode, member order, spacing — must be
    accepted, and the registered members must be found by name and parsed to the same numbers."""
    members = list(extra.items()) + [("_t", "Evt"), ("by", by), ("seq", seq), ("w", w), ("big", big)]
    rnd.shuffle(members)
    text = json.dumps(dict(members), ensure_ascii=ascii_only, separators=((", ", " : ") if spaced else (",", ":")))
    assert json.loads(text)["by"] == by
    ing = Ingest()
    ing.set_json_packer("_t", [("Evt", 2, [("by", N.JSON_I32, 16), ("seq", N.JSON_I32, 4), ("w", N.JSON_F64, 24), ("big", N.JSON_I64, 32)])])
    ing.set_value_framing(N.VALUE_JSON)
    ing.record_batches(0, K.encode_record_batch(0, [(0, b"k", text.encode("utf-8"))]))
    rec = bytes(ing.pending()[0])
    assert struct.unpack_from("<II", rec, 0) == (2, seq)
    assert struct.unpack_from("<i", rec, 16)[0] == by
    assert struct.unpack_from("<d", rec, 24)[0] == w and struct.pack("<d", w) == rec[24:32]
    assert struct.unpack_from("<q", rec, 32)[0] == big
    assert rec[20:24] == bytes(4) and rec[40:] == bytes(24)
