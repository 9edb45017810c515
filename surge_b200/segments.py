"""Kafka log SEGMENT FILES as the input of a rebuild: the bytes of a partition directory go to the record-batch decoders as they are.

A partition directory (`<topic>-<partition>/`) holds, per segment, `<baseOffset>.log` — a plain sequence of RecordBatch v2
structures, byte-identical to what a fetch response carries (kafka `FileRecords`; the reference's topics are written by its
transactional producer, modules/command-engine/core/src/main/scala/surge/internal/kafka/KafkaProducerActorImpl.scala:321-329,
and read read_committed, modules/common/src/main/scala/surge/kafka/streams/SurgeStateStoreConsumer.scala:38) — and
`<baseOffset>.txnindex`, the aborted transactions that intersect the segment: 34-byte entries
`int16 version, int64 producerId, int64 firstOffset, int64 lastOffset, int64 lastStableOffset` (kafka `AbortedTxn`). That pair is
exactly what the decoders need: bytes for `record_batches` / `submit`, (producerId, firstOffset) for `set_aborted`.

`feed_partition` walks the segments in offset order and hands the bytes over in chunks cut at batch boundaries; it works with the
host decoder (`surge_b200.ingest.Ingest`) and the device decoder (`surge_b200.dingest.DeviceIngest`) alike — they share the three
calls it makes. Nothing here parses records: CRC, decompression, transactions and record parsing stay in the decoders.
"""
import os
import re
import struct
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

ABORTED_TXN_BYTES = 34
_SEGMENT = re.compile(r"^(\d{20})\.log$")


def read_txnindex(path: str) -> List[Tuple[int, int, int, int]]:
    """[(producer_id, first_offset, last_offset, last_stable_offset)] of a `.txnindex` file; a torn tail entry is ignored (the
    broker truncates it on recovery)."""
    with open(path, "rb") as f:
        raw = f.read()
    out = []
    for at in range(0, len(raw) - len(raw) % ABORTED_TXN_BYTES, ABORTED_TXN_BYTES):
        version, pid, first, last, lso = struct.unpack_from(">hqqqq", raw, at)
        if version != 0:
            raise ValueError(f"{path}: aborted-transaction entry version {version} at byte {at} (only 0 is known)")
        out.append((pid, first, last, lso))
    return out


def write_txnindex(path: str, aborted: Sequence[Tuple[int, int, int, int]]) -> None:
    """The inverse of read_txnindex (test fixtures)."""
    with open(path, "wb") as f:
        for pid, first, last, lso in aborted:
            f.write(struct.pack(">hqqqq", 0, pid, first, last, lso))


def partition_segments(directory: str) -> List[Tuple[int, str, Optional[str]]]:
    """[(base_offset, log path, txnindex path or None)] of a partition directory, by base offset."""
    segs = []
    for name in os.listdir(directory):
        m = _SEGMENT.match(name)
        if m:
            tx = os.path.join(directory, m.group(1) + ".txnindex")
            segs.append((int(m.group(1)), os.path.join(directory, name), tx if os.path.exists(tx) else None))
    segs.sort()
    return segs


def batch_chunks(data, chunk_bytes: int) -> Iterator[Tuple[int, int]]:
    """(begin, end) ranges of `data` that end on batch boundaries and hold about chunk_bytes each (at least one batch).
    Stops at the first position that cannot start a batch: a zero length (the preallocated tail of an active segment) or a
    batch that runs past the end (a torn write) — what lies behind it is not log."""
    n = len(data)
    begin = pos = 0
    while n - pos >= 12:
        (length,) = struct.unpack_from(">i", data, pos + 8)
        if length <= 0 or pos + 12 + length > n:
            break
        pos += 12 + length
        if pos - begin >= chunk_bytes:
            yield begin, pos
            begin = pos
    if pos > begin:
        yield begin, pos


def _last_batch_is_whole(data, begin: int, end: int) -> bool:
    """CRC-32C of the batch [begin, end) against its header field (what the broker's log recovery checks before it truncates)."""
    import ctypes as C

    from . import native as N

    if end - begin < 61:
        return False
    (stored,) = struct.unpack_from(">I", data, begin + 17)
    body = bytes(data[begin + 21:end])
    return int(N.load_library().sgr_crc32c(C.c_char_p(body), len(body))) == stored


def feed_partition(decoder, partition: int, directory: str, chunk_bytes: int = 64 << 20, from_offset: int = 0) -> Dict[str, int]:
    """All segments of `directory` whose records may lie at or above `from_offset`, in order, into `decoder`
    (Ingest: record_batches / DeviceIngest: submit — the caller folds afterwards). Returns the summed statistics.
    The ACTIVE (last) segment may end in a torn write: its final batch is dropped when its CRC does not hold, like the broker's
    recovery does; a bad CRC anywhere else is corruption and is left to the decoder to refuse."""
    import mmap

    segs = partition_segments(directory)
    # a segment is needed unless the NEXT one starts at or below from_offset
    keep = [s for i, s in enumerate(segs) if i + 1 == len(segs) or segs[i + 1][0] > from_offset]
    total: Dict[str, int] = {}
    push = getattr(decoder, "submit", None) or decoder.record_batches
    for _base, log, tx in keep:
        if tx is not None:
            aborted = [(pid, first) for pid, first, _last, _lso in read_txnindex(tx)]
            if aborted:
                decoder.set_aborted(partition, aborted)
        if os.path.getsize(log) == 0:
            continue
        with open(log, "rb") as f, mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ) as mm:
            chunks = list(batch_chunks(mm, chunk_bytes))
            if chunks and log == keep[-1][1]:
                # position of the final batch: walk the last chunk's boundaries once more
                b, e = chunks[-1]
                pos = last = b
                while pos < e:
                    last = pos
                    pos += 12 + struct.unpack_from(">i", mm, pos + 8)[0]
                if not _last_batch_is_whole(mm, last, e):
                    chunks[-1] = (b, last)
            for b, e in chunks:
                if e <= b:
                    continue
                st = push(partition, bytes(mm[b:e]))   # (a copy: the device decoder reads it asynchronously until its fold)
                for k, v in st.items():
                    total[k] = total.get(k, 0) + int(v)
    return total
