"""Executable model of the TMA ring protocol of the strip kernels (csrc/scope_kernels.cuh).

The kernels cannot be run without a GPU, and their riskiest part is not arithmetic but the
producer/consumer protocol: mbarrier phases named only by their parity, TMA loads that may
complete out of order, a ring that is refilled while slow warps are still around.  This model
transcribes that protocol (tma_produce, tma_consume's tile walk, tma_consume_groups' row-group
walk) into Python coroutines, runs them under randomly chosen interleavings with randomly delayed
and reordered TMA completions, and checks the properties the kernels rely on:

  * no deadlock;
  * every read of a stage sees the tile the reader expects (no refill before the last reader
    left, no wait that returns for the wrong phase);
  * every group of every tile of every strip is read exactly once;
  * every "empty" phase gets exactly the arrivals it was initialised for.

mbarrier semantics modelled: a barrier holds a phase counter, the arrivals still pending in the
current phase and the transaction bytes still pending; `wait(parity)` succeeds iff the parity of
the CURRENT phase differs from `parity` ("the phase with that parity has completed"), which is
exactly why a waiter two phases behind or ahead gets a wrong answer.

Used by tests/test_ring_protocol.py; `python tools/ring_model.py` runs a larger random sweep.
"""
from __future__ import annotations

import random
from dataclasses import dataclass, field


class ProtocolError(AssertionError):
    pass


class MBarrier:
    def __init__(self, count: int):
        self.count = count
        self.phase = 0          # number of completed phases
        self.pending = count    # arrivals still expected in the current phase
        self.tx = 0             # transaction bytes still expected in the current phase
        self.arrivals_in_phase = 0

    def _maybe_complete(self):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.count
            self.arrivals_in_phase = 0

    def arrive(self):
        if self.pending == 0:
            raise ProtocolError("arrival on a barrier whose phase already has all its arrivals")
        self.pending -= 1
        self.arrivals_in_phase += 1
        self._maybe_complete()

    def arrive_expect_tx(self, nbytes: int):
        self.tx += nbytes
        self.arrive()

    def complete_tx(self, nbytes: int):
        self.tx -= nbytes
        self._maybe_complete()

    def test(self, parity: int) -> bool:
        return (self.phase & 1) != parity


@dataclass
class Ring:
    stages: int
    consumers_per_phase: int
    full: list = field(default_factory=list)
    empty: list = field(default_factory=list)
    content: list = field(default_factory=list)   # tile id each stage currently holds
    in_flight: list = field(default_factory=list)  # issued TMA loads not yet landed: (stage, tile)

    def __post_init__(self):
        self.full = [MBarrier(1) for _ in range(self.stages)]
        self.empty = [MBarrier(self.consumers_per_phase) for _ in range(self.stages)]
        self.content = [None] * self.stages


TILE_BYTES = 8192


def producer(ring: Ring, chunks: list, tiles_per_strip: int, queue: list, kqueue: int, log: dict):
    """tma_produce: chunks = [(first_item, count), ...]; the end marker follows the last chunk."""
    stage, phase, qw = 0, 0, 0
    for first, count in chunks + [(None, 0)]:
        while not ring.empty[stage].test(phase ^ 1):
            yield "producer waits empty (chunk)"
        queue[qw % kqueue] = (first, count, qw)   # (the sequence number only exists in the model)
        qw += 1
        if count == 0:
            ring.full[stage].arrive()
            return
        for item in range(first, first + count):
            for t in range(tiles_per_strip):
                if not (item == first and t == 0):
                    while not ring.empty[stage].test(phase ^ 1):
                        yield "producer waits empty"
                tile = log["issued"]
                log["issued"] += 1
                ring.full[stage].arrive_expect_tx(TILE_BYTES)
                ring.in_flight.append((stage, tile))
                stage += 1
                if stage == ring.stages:
                    stage, phase = 0, phase ^ 1
                yield "producer issued"


def read_mailbox(queue: list, qr: int, kqueue: int):
    """a consumer's read of chunk announcement number qr: the entry must still be that chunk's (the
    producer must not have lapped the mailbox: kQueue >= kStages in scope_kernels.cuh)"""
    first, count, seq = queue[qr % kqueue]
    if seq != qr:
        raise ProtocolError(f"chunk mailbox entry {qr % kqueue} overwritten: expected announcement {qr}, found {seq}")
    return first, count


def land_one(ring: Ring, rng: random.Random):
    """one in-flight TMA load completes (any of them: completions may be reordered)"""
    stage, tile = ring.in_flight.pop(rng.randrange(len(ring.in_flight)))
    ring.content[stage] = tile
    ring.full[stage].complete_tx(TILE_BYTES)


def consumer_tiles(ring: Ring, warp: int, tiles_per_strip: int, queue: list, kqueue: int, reads: dict,
                   strip_barrier, peek_prob: float, rng: random.Random):
    """tma_consume: every warp reads its rows of every tile, in lock step with the ring."""
    stage, phase, qr = 0, 0, 0
    tile = 0
    while True:
        while not ring.full[stage].test(phase):
            yield "consumer waits chunk"
        first, count = read_mailbox(queue, qr, kqueue)
        qr += 1
        if count == 0:
            return
        for item in range(first, first + count):
            skip_wait = item == first
            landed = False
            for t in range(tiles_per_strip):
                if not skip_wait and not landed:
                    while not ring.full[stage].test(phase):
                        yield "consumer waits tile"
                skip_wait, landed = False, False
                if ring.content[stage] != tile:
                    raise ProtocolError(f"warp {warp} expected tile {tile} in stage {stage}, found {ring.content[stage]}")
                reads[(tile, warp)] = reads.get((tile, warp), 0) + 1
                yield "consumer read"
                ring.empty[stage].arrive()
                tile += 1
                stage += 1
                if stage == ring.stages:
                    stage, phase = 0, phase ^ 1
                if t + 1 < tiles_per_strip and rng.random() < peek_prob:
                    landed = ring.full[stage].test(phase)
                yield "consumer released"
            yield from strip_barrier(warp)


def consumer_groups(ring: Ring, warp: int, nwork: int, gpt: int, tiles_per_strip: int, queue: list, kqueue: int,
                    reads: dict, strip_barrier, peek_prob: float, rng: random.Random, sync_per_strip: bool):
    """tma_consume_groups: warp w reads groups w, w + NW, ... of each strip; visits every tile in order."""
    S = ring.stages
    st = {"next_tile": 0, "waited": 0, "landed": False}

    def ensure_waited(m):
        if st["waited"] <= m:
            if st["waited"] != m:
                raise ProtocolError("tiles waited for out of order")
            if not st["landed"]:
                while not ring.full[m % S].test((m // S) & 1):
                    yield "group consumer waits"
            st["landed"] = False
            st["waited"] = m + 1

    def pass_until(n):
        while st["next_tile"] < n:
            yield from ensure_waited(st["next_tile"])
            ring.empty[st["next_tile"] % S].arrive()
            st["next_tile"] += 1
            yield "passed"

    def advance_to(n):
        yield from pass_until(n)
        yield from ensure_waited(n)

    def peek():
        if st["waited"] == st["next_tile"]:
            m = st["waited"]
            st["landed"] = ring.full[m % S].test((m // S) & 1)

    tile_seq, qr = 0, 0
    groups = tiles_per_strip * gpt
    while True:
        yield from advance_to(tile_seq)
        first, count = read_mailbox(queue, qr, kqueue)
        qr += 1
        if count == 0:
            return
        for item in range(first, first + count):
            g = warp
            while g < groups:
                n = tile_seq + g // gpt
                yield from advance_to(n)
                st["next_tile"] = n + 1
                if ring.content[n % S] != n:
                    raise ProtocolError(f"warp {warp} expected tile {n} in stage {n % S}, found {ring.content[n % S]}")
                reads[(n, g % gpt)] = reads.get((n, g % gpt), 0) + 1
                yield "group read"
                ring.empty[n % S].arrive()
                if rng.random() < peek_prob:
                    peek()
                yield "group released"
                g += nwork
            tile_seq += tiles_per_strip
            yield from pass_until(tile_seq)
            if sync_per_strip:
                yield from strip_barrier(warp)


def run(kind: str, nwork: int, stages: int, tiles_per_strip: int, chunks: list, seed: int, gpt: int = 16,
        kqueue: int = 4, peek_prob: float = 0.7, sync_per_strip: bool = True, max_steps: int = 2_000_000):
    """Simulate one CTA.  kind = 'tiles' (tma_consume) or 'groups' (tma_consume_groups)."""
    rng = random.Random(seed)
    ring = Ring(stages, nwork)
    queue = [None] * kqueue
    log = {"issued": 0}
    reads: dict = {}
    # CTA-wide barrier among the consumer warps (emit_strip / flush_vscope)
    bar = {"gen": 0, "count": 0}

    def strip_barrier(_warp):
        gen = bar["gen"]
        bar["count"] += 1
        if bar["count"] == nwork:
            bar["count"] = 0
            bar["gen"] += 1
        while bar["gen"] == gen:
            yield "strip barrier"

    threads = [producer(ring, chunks, tiles_per_strip, queue, kqueue, log)]
    for w in range(nwork):
        if kind == "tiles":
            threads.append(consumer_tiles(ring, w, tiles_per_strip, queue, kqueue, reads, strip_barrier, peek_prob, rng))
        else:
            threads.append(consumer_groups(ring, w, nwork, gpt, tiles_per_strip, queue, kqueue, reads, strip_barrier,
                                           peek_prob, rng, sync_per_strip))
    alive = list(range(len(threads)))
    idle = 0
    for _ in range(max_steps):
        if not alive:
            break
        # a TMA completion is an event like any thread step; slow it down sometimes so that the
        # ring really runs dry and really fills up
        if ring.in_flight and rng.random() < rng.choice((0.02, 0.2, 0.6)):
            land_one(ring, rng)
            idle = 0
            continue
        i = rng.choice(alive)
        before = (log["issued"], len(reads), sum(b.phase for b in ring.full + ring.empty),
                  sum(b.pending for b in ring.empty), bar["gen"], bar["count"])
        try:
            next(threads[i])
        except StopIteration:
            alive.remove(i)
        after = (log["issued"], len(reads), sum(b.phase for b in ring.full + ring.empty),
                 sum(b.pending for b in ring.empty), bar["gen"], bar["count"])
        idle = 0 if before != after else idle + 1
        if idle > 50 * (nwork + 1) and not ring.in_flight:
            raise ProtocolError(f"deadlock: no progress, {len(alive)} threads alive")
    else:
        raise ProtocolError("step budget exhausted")
    if ring.in_flight:
        raise ProtocolError("TMA loads still in flight at exit")
    n_tiles = sum(c for _, c in chunks) * tiles_per_strip
    if log["issued"] != n_tiles:
        raise ProtocolError("producer issued the wrong number of tiles")
    want = {(t, x) for t in range(n_tiles) for x in range(nwork if kind == "tiles" else gpt)}
    if set(reads) != want or any(v != 1 for v in reads.values()):
        raise ProtocolError("not every (tile, reader) was read exactly once")
    return n_tiles


def make_chunks(n_strips: int, rng: random.Random, max_chunk: int = 4):
    out, first = [], 0
    while first < n_strips:
        c = min(rng.randint(1, max_chunk), n_strips - first)
        out.append((first, c))
        first += c
    return out


if __name__ == "__main__":
    import itertools
    import sys

    n = 0
    rng = random.Random(1)
    for seed, nwork, stages, tiles in itertools.product(range(int(sys.argv[1]) if len(sys.argv) > 1 else 6),
                                                        (16, 23, 24, 27, 31), (2, 3, 4), (1, 2, 3, 5, 9)):
        chunks = make_chunks(rng.randint(1, 7), rng)
        run("groups", nwork, stages, tiles, chunks, seed)
        run("groups", nwork, stages, tiles, chunks, seed, sync_per_strip=False)
        if nwork == 16:
            run("tiles", nwork, stages, tiles, chunks, seed)
        n += 1
    print(f"{n} configurations ok")
