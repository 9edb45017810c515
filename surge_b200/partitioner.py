"""Host-side mirror of the reference's partitioners (modules/common/src/main/scala/surge/kafka/KafkaPartitioner.scala)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import native as N


class KafkaPartitionProvider:
    """trait KafkaPartitionProvider — partitionForKey = math.abs(MurmurHash3.stringHash(s) % n) (:7-9)."""

    def partitionForKey(self, partitionByString: str, numberOfPartitions: int) -> int:  # noqa: N802 - reference name
        lib = N.load_library()
        kb = partitionByString.encode("utf-8")
        out = C.c_int32()
        buf = C.create_string_buffer(kb, len(kb)) if kb else None
        rc = lib.sgr_partition_for_key_utf8(C.cast(buf, C.c_void_p) if buf else None, len(kb), numberOfPartitions, 0, C.byref(out))
        if rc != 0:
            raise ValueError(f"partitionForKey({partitionByString!r}, {numberOfPartitions}) -> {rc}")
        return int(out.value)


class PartitionStringUpToColon(KafkaPartitionProvider):
    """final class PartitionStringUpToColon — partitionBy = str.takeWhile(_ != ':') (:38-42)."""

    @staticmethod
    def partitionBy(s: str) -> str:  # noqa: N802
        i = s.find(":")
        return s if i < 0 else s[:i]

    def partition_of_record_key(self, key: str, numberOfPartitions: int) -> int:
        return self.partitionForKey(self.partitionBy(key), numberOfPartitions)


def string_hash(s: str) -> int:
    """scala.util.hashing.MurmurHash3.stringHash over the UTF-16 code units of s."""
    u = np.frombuffer(s.encode("utf-16-le"), dtype=np.uint16).copy()
    return int(N.load_library().sgr_string_hash_utf16(u.ctypes.data if len(u) else None, len(u)))
