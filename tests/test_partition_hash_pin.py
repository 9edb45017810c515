"""The partition hash (scala.util.hashing.MurmurHash3.stringHash, scala-library 2.13.8; call site
modules/common/src/main/scala/surge/kafka/KafkaPartitioner.scala:7-9) has no known-answer vector in the reference, and no Scala
runtime exists here. What exists is a real MurmurHash3_x86_32 (scikit-learn's Cython binding of Appleby's code), and
stringHash IS that function over a re-ordered byte stream with a different length word:

  * stringHash mixes one 32-bit word per PAIR of UTF-16 units, data = (s[i] << 16) + s[i+1]; the standard function reads its
    blocks little-endian, so the same word comes from the bytes le16(s[i+1]) + le16(s[i]);
  * an odd last unit goes through mixLast, which is exactly the standard 2-byte tail;
  * the standard function finishes with fmix32(h ^ byte_length), stringHash with fmix32(h ^ unit_length); fmix32 is a bijection,
    so h is recovered from the library's result and re-finished with the other length.

That pins the seed handling, the round constants, the rotations, the block order and the tail of every restatement in this
repository (oracle C, oracle Python, the product's partitioner.cpp) to a real implementation; only the bijection arithmetic is ours.
"""
import struct

import numpy as np
import pytest

from oracle import oracle as O
from oracle import surge_model as M
from surge_b200 import partitioner as PT

M32 = 0xFFFFFFFF
SEED = 0xF7CA7FD2          # MurmurHash3.stringSeed


def fmix32(h):
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & M32
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & M32
    h ^= h >> 16
    return h


def inv_fmix32(h):
    h ^= h >> 16
    h = (h * pow(0xC2B2AE35, -1, 1 << 32)) & M32
    h ^= (h >> 13) ^ (h >> 26)
    h = (h * pow(0x85EBCA6B, -1, 1 << 32)) & M32
    h ^= h >> 16
    return h


def scala_string_hash_via_library(s: str) -> int:
    murmurhash3_32 = pytest.importorskip("sklearn.utils").murmurhash3_32
    units = np.frombuffer(s.encode("utf-16-le"), dtype="<u2")
    stream = bytearray()
    for i in range(0, len(units) - 1, 2):
        stream += struct.pack("<HH", int(units[i + 1]), int(units[i]))
    if len(units) % 2:
        stream += struct.pack("<H", int(units[-1]))
    std = int(murmurhash3_32(bytes(stream), seed=SEED, positive=True))
    h = inv_fmix32(std) ^ len(stream)
    out = fmix32(h ^ len(units))
    return out - (1 << 32) if out & 0x80000000 else out


def test_fmix_inverse_is_an_inverse():
    rng = np.random.default_rng(1)
    for v in [0, 1, M32, 0x80000000] + rng.integers(0, 1 << 32, 200).tolist():
        assert inv_fmix32(fmix32(int(v))) == int(v) and fmix32(inv_fmix32(int(v))) == int(v)


def test_every_restatement_agrees_with_the_library_route():
    rng = np.random.default_rng(2)
    samples = ["", "a", "ab", "abc", "agg-1", "aggregate-00012345", "zażółć gęślą", "\U0001F600", "a\U0001F600b", " ", "x" * 257]
    for _ in range(300):
        n = int(rng.integers(0, 40))
        kinds = rng.integers(0, 3, size=n)
        chars = [chr(int(rng.integers(32, 127))) if k == 0 else chr(int(rng.integers(0x100, 0xD7FF))) if k == 1 else chr(int(rng.integers(0x10000, 0x10FFFF)))
                 for k in kinds]
        samples.append("".join(chars))
    provider = PT.KafkaPartitionProvider()
    for s in samples:
        want = scala_string_hash_via_library(s)
        assert O.scala_string_hash(s) == want, repr(s)
        assert M.scala_string_hash(s) == want, repr(s)
        assert PT.string_hash(s) == want, repr(s)                       # the product's partitioner.cpp, UTF-16 entry point
        for parts in (1, 7, 32, 1000):                                   # and its UTF-8 entry point incl. the abs(% n)
            assert provider.partitionForKey(s, parts) == abs(_jvm_rem(want, parts)), (repr(s), parts)


def _jvm_rem(a: int, n: int) -> int:
    """JVM % truncates toward zero."""
    r = abs(a) % n
    return -r if a < 0 else r
