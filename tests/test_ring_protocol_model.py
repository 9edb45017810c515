"""A model of the device decoder's input ring (surge_b200/csrc/dingest_kernels.cu, RingIn) under the memory system's actual
guarantee: asynchronous copies (cp.async) complete in groups IN ORDER OF ACCOUNTING — `wait_group N` returns once all but the N
most recent groups are done — but the bytes of two copies that are both still in flight may land in either order.

The ring has 8 slots (chunk c lives in slot c mod 8), is kept up to 6 chunks ahead, and a reader needs the chunk under the read
position and the next one. The protocol that shipped first issued k requests for an advance of k chunks and waited once; for
k >= 3 the third request goes to a slot whose previous copy may still be among the six allowed to be pending, so the slot's final
content depends on arrival order. 1 batch in ~7 800 failed on the GPU with 38-byte records (k = 3 on most records), none with the
bench's 34-byte records; this model shows the hazard for every step size >= 33 bytes and none for the protocol that waits after each
request (RingIn::advance as it is now). The GPU-side pin is tests/test_gpu_dingest.py::test_forty_byte_records_walk_through_the_ring."""
import random

R, AHEAD = 8, 6


class Ring:
    def __init__(self, wait_after_each_request: bool):
        self.fixed = wait_after_each_request
        self.slots = [None] * R       # chunk whose copy has certainly landed last in the slot
        self.pending = []             # requests accounted as possibly in flight, oldest first
        self.base = None
        self.hazards = 0              # requests aimed at a slot that another in-flight request also targets
        self.stale = 0                # reads of a chunk that is not guaranteed to be in its slot

    def _issue(self, chunk):
        if any(c % R == chunk % R for c in self.pending):
            self.hazards += 1
        self.pending.append(chunk)

    def _wait(self, n):
        while len(self.pending) > n:
            c = self.pending.pop(0)
            self.slots[c % R] = c

    def seek(self, p):
        self._wait(0)
        self.base = p // 16
        for k in range(R):
            self._issue(self.base + k)
        self._wait(AHEAD)

    def advance(self, p):
        if p < (self.base + 1) * 16:
            return
        if p >= (self.base + R) * 16:
            self.seek(p)
            return
        while True:
            self._issue(self.base + R)
            self.base += 1
            if self.fixed:
                self._wait(AHEAD)
            if p < (self.base + 1) * 16:
                break
        self._wait(AHEAD)

    def get64(self, p):
        assert self.base * 16 <= p < (self.base + 1) * 16
        a = p & ~7
        for w in (a, a + 8):
            c = w // 16
            if self.slots[c % R] != c or c in self.pending:
                self.stale += 1


def _run(fixed, steps, trials=300, length=300):
    rng = random.Random(7)
    hazards = stale = 0
    for _ in range(trials):
        r = Ring(fixed)
        p = rng.randrange(0, 4096)
        r.seek(p)
        for _ in range(length):
            r.get64(p)
            p += rng.choice(steps)
            r.advance(p)
        hazards += r.hazards
        stale += r.stale
    return hazards, stale


def test_waiting_after_every_request_leaves_no_two_copies_aimed_at_one_slot():
    for steps in ([1, 2, 3, 5, 8], [33, 34], [37, 38, 39, 40], [1, 7, 16, 33, 48, 64, 100, 127, 128, 129, 300]):
        assert _run(True, steps) == (0, 0), steps


def test_one_wait_per_advance_was_wrong_from_three_chunks_on():
    # advances of at most two chunks are fine either way (the lz4 token / literal / offset reads) ...
    assert _run(False, [1, 2, 3, 5, 8, 16, 17, 31]) == (0, 0)
    # ... three chunks (a 33..48-byte record from most alignments) are not
    hazards, _ = _run(False, [37, 38, 39, 40])
    assert hazards > 0
    hazards_bench_records, _ = _run(False, [33, 34])
    assert 0 < hazards_bench_records < hazards      # rarer with 34-byte records: why the bench never tripped over it
