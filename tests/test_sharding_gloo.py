"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: band/shard partitions and the
all-reduce + saturate algebra of tile-sharded frames.  The per-band partial counts come from
the CPU oracle here (test infrastructure); on GPUs they come from scope_accumulate_partial."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_partitions(pkg):
    sh = pkg.sharding
    assert [list(sh.frame_shard(64, r, 8)) for r in range(8)] == [list(range(8 * r, 8 * r + 8)) for r in range(8)]
    assert sum(len(sh.frame_shard(10, r, 4)) for r in range(4)) == 10
    assert sh.row_bands(4320, 4) == [(0, 1080), (1080, 2160), (2160, 3240), (3240, 4320)]
    bands = sh.col_bands(7680, 4)
    assert bands == [(0, 1920), (1920, 3840), (3840, 5760), (5760, 7680)]
    odd = sh.col_bands(1000, 3)
    assert odd[0][0] == 0 and odd[-1][1] == 1000 and all(a % 32 == 0 for a, _ in odd)
    assert all(odd[i][1] == odd[i + 1][0] for i in range(2))


def _oracle_partials(orc, f, yuv, y0, y1, x0, x1, full_width):
    """what scope_accumulate_partial produces for the tile rows [y0,y1) x cols [x0,x1)"""
    tile = np.ascontiguousarray(f[y0:y1, x0:x1])
    tyuv = np.ascontiguousarray(yuv[y0:y1, x0:x1])
    hist = orc.histogram_counts(0x07, tile, tyuv).astype(np.int32)
    # unsaturated per-column counts, packed as u16 pairs (B,G | R,0)
    h, w = tile.shape[:2]
    pairs = np.zeros((2, 256, full_width), np.int32)
    for c, (word, shift) in enumerate([(0, 0), (0, 16), (1, 0)]):
        cnt = np.zeros((256, w), np.int64)
        a = tile[..., 3] != 0
        for x in range(w):
            cnt[:, x] = np.bincount(tile[a[:, x], x, c], minlength=256)
        pairs[word, ::-1, x0:x1] += (cnt << shift).astype(np.int32)   # row 0 = value 255
    vs = np.zeros(65536, np.int32)
    idx = tyuv[..., 0].astype(np.int64) + 256 * (255 - tyuv[..., 2].astype(np.int64))
    vs += np.bincount(idx.ravel(), minlength=65536).astype(np.int32)
    return hist, pairs, vs


def _worker(rank, world, port, mode, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import obs_color_monitor_b200 as pkg
    from oracle.oracle import Oracle
    orc = Oracle()
    w, h = 96, 300                                   # H > 255: the saturation matters
    f = pkg.frames.natural(w, h, seed=3)
    f[5:40, 10:14, 3] = 0
    yuv = orc.rgb_to_yuv(f, 2)
    bands = pkg.sharding.row_bands(h, world) if mode == "rows" else pkg.sharding.col_bands(w, world)
    a, b = bands[rank]
    y0, y1, x0, x1 = (a, b, 0, w) if mode == "rows" else (0, h, a, b)
    hist, pairs, vs = _oracle_partials(orc, f, yuv, y0, y1, x0, x1, w)
    partial = {"hist": torch.from_numpy(hist), "wave_pairs": torch.from_numpy(pairs), "vscope": torch.from_numpy(vs)}
    pkg.sharding.allreduce_partials(partial)
    # finalize exactly like wave_pairs_finalize_kernel / vscope_finalize_kernel
    pr = partial["wave_pairs"].numpy().view(np.uint32)
    wave = np.zeros((256, w, 4), np.uint8)
    wave[..., 0] = np.minimum(pr[0] & 0xFFFF, 255)
    wave[..., 1] = np.minimum(pr[0] >> 16, 255)
    wave[..., 2] = np.minimum(pr[1] & 0xFFFF, 255)
    vsc = np.minimum(partial["vscope"].numpy(), 255).astype(np.uint8).reshape(256, 256)
    ok = (np.array_equal(partial["hist"].numpy().view(np.uint32), orc.histogram_counts(0x07, f, yuv))
          and np.array_equal(wave, orc.waveform(0x07, f, yuv)) and np.array_equal(vsc, orc.vectorscope(yuv)))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["rows", "cols"])
def test_tile_sharded_allreduce_world2(mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
