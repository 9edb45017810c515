"""TEST INFRASTRUCTURE — CPU restatement of the byte formats on the way INTO the fold.

Only tests/ may import this module. It is an independent encoder + decoder used to check the product decoder
(surge_b200/csrc/ingest.cpp) — never a fallback for it.

PARITY UNPINNED: the formats live in third-party dependencies that are not under the reference checkout, and the
reference's tests hold no broker bytes:
  * Kafka RecordBatch magic 2 / Record / control records — org.apache.kafka:kafka-clients:3.2.3
    (project/Dependencies.scala:42 of the reference; DefaultRecordBatch, DefaultRecord, ControlRecordType, ByteUtils);
    call sites on the reference side: the read_committed consumer that feeds the state store
    (modules/common/src/main/scala/surge/kafka/streams/SurgeStateStoreConsumer.scala:38), the lz4 default of the publisher
    (modules/common/src/main/resources/reference.conf:124), the empty-key flush record
    (modules/command-engine/core/src/main/scala/surge/internal/kafka/KafkaProducerActorImpl.scala:321-329).
  * LZ4 frame + block format (lz4 frame format 1.6.x) as written by KafkaLZ4BlockOutputStream for magic >= 2
    (FLG = 0x60: version 01, independent blocks, no checksums; BD = 0x40: 64 KiB blocks; correct header checksum).
  * CRC-32C (RFC 3720 appendix B.4 vectors) and xxHash32 (published vectors).
What IS pinned (tests/test_ingest_cpu.py): CRC-32C by the RFC 3720 vectors; xxHash32 by the `xxhash` package; the LZ4 frame
and block codec by liblz4 itself (pyarrow's "lz4" codec: frames made by the library decode here, frames made here decode in the
library); the record-batch FRAMING around them stays unpinned — no Kafka client exists in this image.
"""
from __future__ import annotations

import struct
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

M32 = 0xFFFFFFFF

# ----------------------------------------------------------------------------- CRC-32C
_CRC_TABLE: List[int] = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ 0x82F63B78 if _c & 1 else _c >> 1
    _CRC_TABLE.append(_c)


def crc32c(data: bytes) -> int:
    c = M32
    for b in data:
        c = _CRC_TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ M32


# ----------------------------------------------------------------------------- xxHash32
_P1, _P2, _P3, _P4, _P5 = 2654435761, 2246822519, 3266489917, 668265263, 374761393


def _rotl(x: int, r: int) -> int:
    return ((x << r) | (x >> (32 - r))) & M32


def xxh32(data: bytes, seed: int = 0) -> int:
    n = len(data)
    p = 0
    if n >= 16:
        v = [(seed + _P1 + _P2) & M32, (seed + _P2) & M32, seed & M32, (seed - _P1) & M32]
        while p + 16 <= n:
            for k in range(4):
                (w,) = struct.unpack_from("<I", data, p)
                v[k] = (_rotl((v[k] + w * _P2) & M32, 13) * _P1) & M32
                p += 4
        h = (_rotl(v[0], 1) + _rotl(v[1], 7) + _rotl(v[2], 12) + _rotl(v[3], 18)) & M32
    else:
        h = (seed + _P5) & M32
    h = (h + n) & M32
    while p + 4 <= n:
        (w,) = struct.unpack_from("<I", data, p)
        h = (_rotl((h + w * _P3) & M32, 17) * _P4) & M32
        p += 4
    while p < n:
        h = (_rotl((h + data[p] * _P5) & M32, 11) * _P1) & M32
        p += 1
    h ^= h >> 15
    h = (h * _P2) & M32
    h ^= h >> 13
    h = (h * _P3) & M32
    h ^= h >> 16
    return h


# ----------------------------------------------------------------------------- LZ4 block + frame
def lz4_block_compress(data: bytes) -> bytes:
    """Greedy single-probe matcher. Honours the format's end-of-block rules (last 5 bytes are literals, the last match
    starts at least 12 bytes before the end) so that any conforming decoder accepts it."""
    n = len(data)
    out = bytearray()
    table: Dict[bytes, int] = {}
    anchor = 0
    i = 0

    def emit(lit: bytes, mlen: int, off: int) -> None:
        ll = len(lit)
        ml = mlen - 4 if mlen else 0
        out.append((min(ll, 15) << 4) | (min(ml, 15) if mlen else 0))
        if ll >= 15:
            r = ll - 15
            while r >= 255:
                out.append(255)
                r -= 255
            out.append(r)
        out.extend(lit)
        if mlen:
            out.extend(struct.pack("<H", off))
            if ml >= 15:
                r = ml - 15
                while r >= 255:
                    out.append(255)
                    r -= 255
                out.append(r)

    while i + 12 < n:
        key = data[i:i + 4]
        cand = table.get(key)
        table[key] = i
        if cand is not None and i - cand <= 0xFFFF:
            m = 4
            limit = n - 5
            while i + m < limit and data[cand + m] == data[i + m]:
                m += 1
            emit(data[anchor:i], m, i - cand)
            i += m
            anchor = i
        else:
            i += 1
    emit(data[anchor:], 0, 0)
    return bytes(out)


def lz4_block_decompress(block: bytes, history: bytearray) -> None:
    """Appends to `history` (matches may reach back into it)."""
    ip = 0
    n = len(block)
    while True:
        token = block[ip]
        ip += 1
        lit = token >> 4
        if lit == 15:
            while True:
                s = block[ip]
                ip += 1
                lit += s
                if s != 255:
                    break
        history.extend(block[ip:ip + lit])
        ip += lit
        if ip >= n:
            return
        off = block[ip] | (block[ip + 1] << 8)
        ip += 2
        ml = token & 15
        if ml == 15:
            while True:
                s = block[ip]
                ip += 1
                ml += s
                if s != 255:
                    break
        ml += 4
        if off == 0 or off > len(history):
            raise ValueError("bad match offset")
        start = len(history) - off
        for k in range(ml):
            history.append(history[start + k])


def lz4_frame_compress(data: bytes, *, block_code: int = 4, block_checksum: bool = False, content_checksum: bool = False,
                       content_size: bool = False, store_incompressible: bool = True) -> bytes:
    flg = 0x40 | 0x20 | (0x10 if block_checksum else 0) | (0x08 if content_size else 0) | (0x04 if content_checksum else 0)
    desc = bytes([flg, block_code << 4]) + (struct.pack("<Q", len(data)) if content_size else b"")
    out = bytearray(struct.pack("<I", 0x184D2204) + desc + bytes([(xxh32(desc) >> 8) & 0xFF]))
    bmax = 1 << (8 + 2 * block_code)
    for s in range(0, len(data), bmax):
        raw = data[s:s + bmax]
        comp = lz4_block_compress(raw)
        if store_incompressible and len(comp) >= len(raw):
            out += struct.pack("<I", len(raw) | 0x80000000) + raw
            body = raw
        else:
            out += struct.pack("<I", len(comp)) + comp
            body = comp
        if block_checksum:
            out += struct.pack("<I", xxh32(body))
    out += struct.pack("<I", 0)
    if content_checksum:
        out += struct.pack("<I", xxh32(data))
    return bytes(out)


def lz4_frame_decompress(frame: bytes) -> bytes:
    if struct.unpack_from("<I", frame, 0)[0] != 0x184D2204:
        raise ValueError("bad magic")
    flg, bd = frame[4], frame[5]
    p = 6
    if flg & 0x08:
        p += 8
    if flg & 0x01:
        p += 4
    if ((xxh32(frame[4:p]) >> 8) & 0xFF) != frame[p]:
        raise ValueError("header checksum")
    p += 1
    out = bytearray()
    while True:
        (w,) = struct.unpack_from("<I", frame, p)
        p += 4
        if w == 0:
            break
        size = w & 0x7FFFFFFF
        body = frame[p:p + size]
        p += size
        if flg & 0x10:
            if xxh32(body) != struct.unpack_from("<I", frame, p)[0]:
                raise ValueError("block checksum")
            p += 4
        if w & 0x80000000:
            out += body
        else:
            lz4_block_decompress(body, out)
    if flg & 0x04 and xxh32(bytes(out)) != struct.unpack_from("<I", frame, p)[0]:
        raise ValueError("content checksum")
    return bytes(out)


# ----------------------------------------------------------------------------- varints (ByteUtils.writeVarint / writeVarlong)
def varint(v: int) -> bytes:
    z = ((v << 1) ^ (v >> 31)) & M32
    return _uvar(z)


def varlong(v: int) -> bytes:
    z = ((v << 1) ^ (v >> 63)) & 0xFFFFFFFFFFFFFFFF
    return _uvar(z)


def _uvar(z: int) -> bytes:
    out = bytearray()
    while z & ~0x7F:
        out.append((z & 0x7F) | 0x80)
        z >>= 7
    out.append(z)
    return bytes(out)


def read_varint(buf: bytes, p: int) -> Tuple[int, int]:
    z = 0
    shift = 0
    while True:
        b = buf[p]
        p += 1
        z |= (b & 0x7F) << shift
        if not b & 0x80:
            break
        shift += 7
    return (z >> 1) ^ -(z & 1), p


# ----------------------------------------------------------------------------- RecordBatch v2
Record = Tuple[int, Optional[bytes], Optional[bytes]]  # (offset_delta, key, value)
ABORT, COMMIT = 0, 1
COMPRESSION = {"none": 0, "gzip": 1, "snappy": 2, "lz4": 3, "zstd": 4}


def encode_record(offset_delta: int, key: Optional[bytes], value: Optional[bytes], headers: Sequence[Tuple[bytes, Optional[bytes]]] = (),
                  timestamp_delta: int = 0) -> bytes:
    body = bytearray(b"\x00") + varlong(timestamp_delta) + varint(offset_delta)
    body += varint(-1) if key is None else varint(len(key)) + key
    body += varint(-1) if value is None else varint(len(value)) + value
    body += varint(len(headers))
    for hk, hv in headers:
        body += varint(len(hk)) + hk
        body += varint(-1) if hv is None else varint(len(hv)) + hv
    return varint(len(body)) + bytes(body)


def encode_record_batch(base_offset: int, records: Sequence[Record], *, compression: str = "none", producer_id: int = -1,
                        producer_epoch: int = -1, base_sequence: int = -1, transactional: bool = False, control: bool = False,
                        base_timestamp: int = 1_600_000_000_000, headers: Sequence[Tuple[bytes, Optional[bytes]]] = (),
                        lz4_kwargs: Optional[dict] = None, magic: int = 2, last_offset_delta: Optional[int] = None) -> bytes:
    body = b"".join(encode_record(d, k, v, headers, timestamp_delta=d) for d, k, v in records)
    codec = COMPRESSION[compression]
    if codec == 3:
        body = lz4_frame_compress(body, **(lz4_kwargs or {}))
    elif codec != 0:
        body = b"\x00" * 8  # the product must refuse before looking inside
    attrs = codec | (0x10 if transactional else 0) | (0x20 if control else 0)
    lod = last_offset_delta if last_offset_delta is not None else (max(d for d, _, _ in records) if records else 0)
    tail = struct.pack(">hiqqqhii", attrs, lod, base_timestamp, base_timestamp + lod, producer_id, producer_epoch, base_sequence, len(records)) + body
    batch_length = 4 + 1 + 4 + len(tail)  # partitionLeaderEpoch + magic + crc + rest
    return struct.pack(">qiib", base_offset, batch_length, 0, magic) + struct.pack(">I", crc32c(tail)) + tail


def encode_control_batch(base_offset: int, producer_id: int, kind: int, producer_epoch: int = 0) -> bytes:
    key = struct.pack(">hh", 0, kind)
    value = struct.pack(">hi", 0, 0)  # version, coordinatorEpoch
    return encode_record_batch(base_offset, [(0, key, value)], producer_id=producer_id, producer_epoch=producer_epoch, transactional=True, control=True)


def decode_record_batches(buf: bytes) -> List[dict]:
    """Plain restatement of DefaultRecordBatch iteration; raises on CRC mismatch; stops at a trailing partial batch."""
    out = []
    p = 0
    while len(buf) - p >= 12:
        base_offset, batch_length = struct.unpack_from(">qi", buf, p)
        total = 12 + batch_length
        if len(buf) - p < total:
            break
        magic = struct.unpack_from(">b", buf, p + 16)[0]
        if magic != 2:
            raise ValueError("magic")
        (crc,) = struct.unpack_from(">I", buf, p + 17)
        if crc != crc32c(buf[p + 21:p + total]):
            raise ValueError("crc")
        attrs, lod, ts0, ts1, pid, pep, bseq, count = struct.unpack_from(">hiqqqhii", buf, p + 21)
        body = buf[p + 61:p + total]
        if attrs & 7 == 3:
            body = lz4_frame_decompress(body)
        elif attrs & 7:
            raise NotImplementedError("codec")
        recs = []
        q = 0
        for _ in range(count):
            ln, q = read_varint(body, q)
            end = q + ln
            q += 1
            _, q = read_varint(body, q)
            od, q = read_varint(body, q)
            kl, q = read_varint(body, q)
            key = None
            if kl >= 0:
                key = body[q:q + kl]
                q += kl
            vl, q = read_varint(body, q)
            val = None
            if vl >= 0:
                val = body[q:q + vl]
                q += vl
            q = end
            recs.append((od, key, val))
        out.append(dict(base_offset=base_offset, last_offset=base_offset + lod, attrs=attrs, producer_id=pid, records=recs,
                        transactional=bool(attrs & 0x10), control=bool(attrs & 0x20)))
        p += total
    return out


def read_committed_pack(fetches: Iterable[Tuple[int, bytes, Sequence[Tuple[int, int]]]]):
    """Restatement of what the product's ingest must produce for a sequence of fetches
    [(partition, bytes, aborted [(producer_id, first_offset)])]: packed 64-byte records in arrival order, the id
    dictionary in first-seen order, and next offsets per partition."""
    keys: Dict[bytes, int] = {}
    recs: List[bytes] = []
    nxt: Dict[int, int] = {}
    aborting: Dict[int, set] = {}
    pending: Dict[int, list] = {}
    for partition, buf, aborted in fetches:
        pend = pending.setdefault(partition, [])
        pend.extend((fo, pid) for pid, fo in aborted)
        pend.sort()
        act = aborting.setdefault(partition, set())
        for b in decode_record_batches(buf):
            while pend and pend[0][0] <= b["last_offset"]:
                act.add(pend.pop(0)[1])
            if b["control"]:
                k = b["records"][0][1]
                if struct.unpack(">hh", k[:4])[1] == ABORT:
                    act.discard(b["producer_id"])
            elif b["transactional"] and b["producer_id"] in act:
                pass
            else:
                for od, key, val in b["records"]:
                    off = b["base_offset"] + od
                    if partition in nxt and off < nxt[partition]:
                        continue
                    if not key or val is None:
                        continue
                    agg_id = key.split(b":", 1)[0]
                    idx = keys.setdefault(agg_id, len(keys))
                    recs.append(val[:8] + struct.pack("<Q", idx) + val[8:] + b"\x00" * (56 - len(val)))
            nxt[partition] = max(nxt.get(partition, 0), b["last_offset"] + 1)
    arr = np.frombuffer(b"".join(recs), dtype=np.uint8).reshape(-1, 64) if recs else np.zeros((0, 64), np.uint8)
    return arr, list(keys), nxt
