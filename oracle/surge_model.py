"""Object-level restatement of the reference's replay path, mirroring the Scala line by line.

TEST INFRASTRUCTURE ONLY (see oracle/sgr_oracle.h): imported by tests/ to pin the C
oracle and the golden vectors; never by surge_b200/.

Where the C oracle works on packed bytes, this file keeps the reference's own shapes —
case classes, Option (None / object), Seq, foldLeft — so each golden vector from the
reference's tests can be written down exactly as the Scala spec
writes it. Pure-Python loops: small cases only.

Paths are relative to the reference checkout.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, fields, replace
from functools import reduce
from typing import Callable, Dict, List, Optional, Sequence, Tuple


def jvm_int(x: int) -> int:
    """Wrap to a JVM Int (two's complement, 32 bits)."""
    x &= 0xFFFFFFFF
    return x - (1 << 32) if x & 0x80000000 else x


# --------------------------------------------------------------------------- Counter
# modules/command-engine/scaladsl/src/test/scala/surge/scaladsl/TestBoundedContext.scala:13-89
@dataclass(frozen=True)
class State:
    aggregateId: str
    count: int
    version: int


@dataclass(frozen=True)
class CountIncremented:
    aggregateId: str
    incrementBy: int
    sequenceNumber: int


@dataclass(frozen=True)
class CountDecremented:
    aggregateId: str
    decrementBy: int
    sequenceNumber: int


@dataclass(frozen=True)
class NoOpEvent:
    aggregateId: str
    sequenceNumber: int


@dataclass(frozen=True)
class ExceptionThrowingEvent:
    aggregateId: str
    sequenceNumber: int
    throwable: Exception


def counter_handle_event(agg: Optional[State], evt) -> Optional[State]:
    """TestBoundedContext.scala:77-89 (BusinessLogicTrait.handleEvent)."""
    current = agg if agg is not None else State(evt.aggregateId, 0, 0)  # agg.getOrElse(State(id,0,0))
    if isinstance(evt, CountIncremented):
        new_state = replace(current, count=jvm_int(current.count + evt.incrementBy), version=evt.sequenceNumber)
    elif isinstance(evt, CountDecremented):
        new_state = replace(current, count=jvm_int(current.count - evt.decrementBy), version=evt.sequenceNumber)
    elif isinstance(evt, NoOpEvent):
        new_state = current
    elif isinstance(evt, ExceptionThrowingEvent):
        raise evt.throwable
    else:
        raise ValueError("scala.MatchError")
    return new_state  # Some(newState)


def counter_process_command(agg: Optional[State], cmd: Tuple[str, str]) -> List:
    """TestBoundedContext.scala:91-103 (processCommand): used to replay the spec's command vectors."""
    new_seq = (agg.version if agg is not None else 0) + 1
    kind, aggregate_id = cmd
    if kind == "Increment":
        return [CountIncremented(aggregate_id, 1, new_seq)]
    if kind == "Decrement":
        return [CountDecremented(aggregate_id, 1, new_seq)]
    if kind == "CreateNoOpEvent":
        return [NoOpEvent(aggregate_id, new_seq)]
    if kind == "DoNothing":
        return []
    raise ValueError(kind)


def ml_counter_apply_event(state: Optional[State], evt) -> Optional[State]:
    """modules/multilanguage/src/test/scala/com/ukg/surge/multilanguage/TestBoundedContext.scala:68-75."""
    current = state if state is not None else State(evt.aggregateId, 0, 0)
    if isinstance(evt, CountIncremented):
        return replace(current, count=jvm_int(current.count + evt.incrementBy), version=evt.sequenceNumber)
    if isinstance(evt, CountDecremented):
        return replace(current, count=jvm_int(current.count - evt.decrementBy), version=evt.sequenceNumber)
    raise ValueError("scala.MatchError")


# --------------------------------------------------------------------------- BankAccount
# modules/surge-docs/src/test/scala/docs/command/BankAccountCommandModel.scala:19,39,46,81-86
@dataclass(frozen=True)
class BankAccount:
    accountNumber: str  # UUID
    accountOwner: str
    securityCode: str
    balance: float


@dataclass(frozen=True)
class BankAccountCreated:
    accountNumber: str
    accountOwner: str
    securityCode: str
    balance: float


@dataclass(frozen=True)
class BankAccountUpdated:
    accountNumber: str
    newBalance: float


def bank_account_handle_event(aggregate: Optional[BankAccount], event) -> Optional[BankAccount]:
    """BankAccountCommandModel.scala:81-86."""
    if isinstance(event, BankAccountCreated):
        return BankAccount(event.accountNumber, event.accountOwner, event.securityCode, event.balance)
    if isinstance(event, BankAccountUpdated):
        return None if aggregate is None else replace(aggregate, balance=event.newBalance)  # aggregate.map(_.copy(..))
    raise ValueError("scala.MatchError")


def bank_account_process_command(aggregate: Optional[BankAccount], command: Tuple) -> List:
    """BankAccountCommandModel.scala:50-79 (processCommand) — f64 add/subtract lives HERE, not in the fold."""
    kind = command[0]
    if kind == "CreateAccount":
        _, n, owner, code, initial = command
        return [] if aggregate is not None else [BankAccountCreated(n, owner, code, initial)]
    if kind == "CreditAccount":
        _, n, amount = command
        if aggregate is None:
            raise KeyError("AccountDoesNotExistException")
        return [BankAccountUpdated(aggregate.accountNumber, aggregate.balance + amount)]
    if kind == "DebitAccount":
        _, n, amount = command
        if aggregate is None:
            raise KeyError("AccountDoesNotExistException")
        if aggregate.balance >= amount:
            return [BankAccountUpdated(aggregate.accountNumber, aggregate.balance - amount)]
        raise ValueError("InsufficientFundsException")
    raise ValueError(kind)


# --------------------------------------------------------------------------- IntBalance
# modules/multilanguage-scala-sdk-sample/src/main/scala/com/ukg/surge/multilanguage/scalasdk/sample/Main.scala:19-30
@dataclass(frozen=True)
class IntBankAccount:
    balance: int


@dataclass(frozen=True)
class MoneyDeposited:
    amount: int


def int_balance_event_handler(agg: Optional[IntBankAccount], evt: MoneyDeposited) -> Optional[IntBankAccount]:
    """Main.scala:25-30."""
    if agg is None:
        return IntBankAccount(evt.amount)
    return IntBankAccount(jvm_int(agg.balance + evt.amount))


# --------------------------------------------------------------------------- the fold and the actor rules
def fold_left(handle_event: Callable, state, events: Sequence):
    """events.foldLeft(state)((stateAccum, evt) => handleEvent(stateAccum, evt))
    modules/command-engine/scaladsl/src/main/scala/surge/scaladsl/command/CommandModels.scala:26"""
    return reduce(handle_event, events, state)


def scala_equals(a, b) -> bool:
    """`state.stateOpt != context.state` (PersistentActor.scala:257) on Option[case class]: scalac's generated equals is
    `(this eq that) || (same class && every field ==)`, and a Double field compares by VALUE (0.0 == -0.0, NaN != NaN).
    Python's dataclass __eq__ compares field tuples, whose element comparison short-cuts on identity — the same NaN
    object would compare equal — so the rule is spelled out here."""
    if a is b:
        return True
    if a is None or b is None or type(a) is not type(b):
        return False
    for f in fields(a):
        x, y = getattr(a, f.name), getattr(b, f.name)
        if not (x == y):
            return False
    return True


@dataclass
class Ack:
    success: bool
    state: object
    published_state: bool  # one record to the STATE topic (never to the events topic for ApplyEvents)
    error: Optional[str] = None


def apply_events(handle_event: Callable, state, events: Sequence) -> Ack:
    """PersistentActor.doApplyEvent, modules/command-engine/core/src/main/scala/surge/internal/persistence/
    PersistentActor.scala:245-264: fold; publish the serialized state iff it changed (:257);
    on exception reply ACKError and keep the previous state (:260-263)."""
    try:
        new_state = fold_left(handle_event, state, events)
    except Exception as e:  # noqa: BLE001 - mirrors .recover { case e => ACKError(e) }
        return Ack(False, state, False, str(e))
    return Ack(True, new_state, not scala_equals(new_state, state))


def handle_command(process_command: Callable, handle_event: Callable, state, cmd) -> Tuple[object, List]:
    """AggregateCommandModel.toCore.handle, CommandModels.scala:18-23: events then the same foldLeft."""
    events = process_command(state, cmd)
    return fold_left(handle_event, state, events), events


# --------------------------------------------------------------------------- KTable (today's recovery)
def ktable_restore(records: Sequence[Tuple[str, Optional[bytes]]]) -> Dict[str, bytes]:
    """builder.table(stateTopic) — last write wins per key, null value deletes.
    modules/common/src/main/scala/surge/kafka/streams/SurgeStateStoreConsumer.scala:57-76;
    null = None state, modules/command-engine/core/src/main/scala/surge/internal/SurgeModel.scala:62-64."""
    table: Dict[str, bytes] = {}
    for key, value in records:
        if value is None:
            table.pop(key, None)
        else:
            table[key] = value
    return table


# --------------------------------------------------------------------------- serialization of integer-only states
def play_json_counter_state(s: State) -> bytes:
    """Json.toJson(state).toString().getBytes with the macro format of State(aggregateId,count,version)
    (core TestBoundedContext.scala:153): field order = case-class order, no whitespace.
    Third-party (play-json 2.9.2) — byte parity unpinned in the reference (it compares parsed JSON)."""
    return json.dumps({"aggregateId": s.aggregateId, "count": s.count, "version": s.version},
                      separators=(",", ":"), ensure_ascii=False).encode("utf-8")


# --------------------------------------------------------------------------- partitioner
def scala_string_hash(s: str) -> int:
    """scala.util.hashing.MurmurHash3.stringHash (scala-library 2.13.8), over UTF-16 code units."""
    b = s.encode("utf-16-le")
    units = [b[i] | (b[i + 1] << 8) for i in range(0, len(b), 2)]
    m = 0xFFFFFFFF

    def rotl(x, r):
        return ((x << r) | (x >> (32 - r))) & m

    def mix_last(h, k):
        k = (k * 0xCC9E2D51) & m
        k = rotl(k, 15)
        k = (k * 0x1B873593) & m
        return h ^ k

    def mix(h, k):
        h = mix_last(h, k)
        h = rotl(h, 13)
        return (h * 5 + 0xE6546B64) & m

    h = 0xF7CA7FD2
    i = 0
    n = len(units)
    while i + 1 < n:
        h = mix(h, ((units[i] << 16) + units[i + 1]) & m)
        i += 2
    if i < n:
        h = mix_last(h, units[i])
    h ^= n
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & m
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & m
    h ^= h >> 16
    return jvm_int(h)


def partition_for_key(partition_by: str, number_of_partitions: int) -> int:
    """math.abs(MurmurHash3.stringHash(s) % n), modules/common/src/main/scala/surge/kafka/KafkaPartitioner.scala:8.
    Java % takes the sign of the dividend."""
    h = scala_string_hash(partition_by)
    r = abs(h) % number_of_partitions
    r = -r if h < 0 else r
    return abs(r)


def partition_string_up_to_colon(s: str) -> str:
    """str.takeWhile(_ != ':'), KafkaPartitioner.scala:38-42."""
    i = s.find(":")
    return s if i < 0 else s[:i]
