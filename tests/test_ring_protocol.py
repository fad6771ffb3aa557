"""The TMA ring protocol of the strip kernels, checked on the CPU with the executable model in
tools/ring_model.py (producer / consumer coroutines under random interleavings, TMA completions
delayed and reordered): no deadlock, every stage read sees the expected tile, every group read
exactly once, barrier arrival counts exact.  The kernels themselves need a GPU; the protocol's
logic does not."""
import os
import random
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import ring_model as rm  # noqa: E402


@pytest.mark.parametrize("nwork", [16, 23, 24, 27, 31])
@pytest.mark.parametrize("stages", [2, 4])
def test_row_group_walk(nwork, stages):
    rng = random.Random(nwork * 10 + stages)
    for seed in range(6):
        for tiles in (1, 2, 5, 34):   # 34 tiles = a 2160-row strip
            chunks = rm.make_chunks(rng.randint(1, 5), rng)
            rm.run("groups", nwork, stages, tiles, chunks, seed)
            rm.run("groups", nwork, stages, tiles, chunks, seed, sync_per_strip=False)


@pytest.mark.parametrize("stages", [2, 4, 8])
def test_tile_walk(stages):
    rng = random.Random(stages)
    for seed in range(6):
        for tiles in (1, 3, 34):
            rm.run("tiles", 16, stages, tiles, rm.make_chunks(rng.randint(1, 5), rng), seed)


def test_chunk_mailbox_must_be_as_long_as_the_ring():
    """The producer announces a chunk once the stage of its first tile is free, so with single-tile
    strips in single-strip chunks it runs kStages announcements ahead of the slowest consumer: a
    4-entry mailbox is lapped by an 8-stage ring (SCOPE_DEEP_RING builds use kQueue = 8), never by a
    4-stage one."""
    chunks = [(i, 1) for i in range(40)]
    for seed in range(10):
        rm.run("tiles", 16, 4, 1, chunks, seed, kqueue=4)
        rm.run("tiles", 16, 8, 1, chunks, seed, kqueue=8)
        rm.run("tiles", 8, 8, 2, chunks, seed, kqueue=8)
    lapped = 0
    for seed in range(30):
        try:
            rm.run("tiles", 16, 8, 1, chunks, seed, kqueue=4)
        except rm.ProtocolError:
            lapped += 1
    assert lapped > 0


def test_model_rejects_a_walk_that_skips_tiles():
    """The first draft of the row-group walk let a warp wait only for the tiles it owns a group in
    (one arrival per group).  A warp can then fall two phases behind a barrier it does not take
    part in, and its parity wait never returns.  The model must find that."""

    def skipping(ring, warp, nwork, gpt, tiles_per_strip, queue, kqueue, reads, strip_barrier, peek_prob, rng,
                 sync_per_strip):
        S, tile_seq, qr = ring.stages, 0, 0
        while True:
            while not ring.full[tile_seq % S].test((tile_seq // S) & 1):
                yield
            first, count = rm.read_mailbox(queue, qr, kqueue)
            qr += 1
            if count == 0:
                return
            for _item in range(first, first + count):
                for g in range(warp, tiles_per_strip * gpt, nwork):
                    n = tile_seq + g // gpt
                    while not ring.full[n % S].test((n // S) & 1):
                        yield
                    if ring.content[n % S] != n:
                        raise rm.ProtocolError("wrong tile")
                    reads[(n, g % gpt)] = reads.get((n, g % gpt), 0) + 1
                    yield
                    ring.empty[n % S].arrive()
                    yield
                tile_seq += tiles_per_strip

    class Ring16(rm.Ring):
        def __post_init__(self):
            self.consumers_per_phase = 16
            super().__post_init__()

    good_consumer, good_ring = rm.consumer_groups, rm.Ring
    rm.consumer_groups, rm.Ring = skipping, Ring16
    try:
        failures = 0
        rng = random.Random(3)
        for seed in range(60):
            try:
                rm.run("groups", 24, 4, rng.choice((3, 5, 9)), rm.make_chunks(rng.randint(2, 7), rng), seed)
            except rm.ProtocolError:
                failures += 1
        assert failures > 0
    finally:
        rm.consumer_groups, rm.Ring = good_consumer, good_ring
