"""N=2 GPU test (skipped with fewer than 2 GPUs): tile-sharded frame over NCCL == whole frame on
one GPU, for row bands and column bands; frame-sharded batch gathered on rank 0 == single-GPU batch."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import obs_color_monitor_b200 as pkg
    from obs_color_monitor_b200 import frames_torch
    eng = pkg.ScopeEngine(rank)
    ok = True
    w, h = 1000, 700
    full = frames_torch.mixed_batch(1, w, h, dev, first_index=3, content="natural")[0]   # same seed on every rank
    whole = eng.accumulate_device(full[None])
    w2 = 1024   # equal column bands (32 strips over 2 ranks): the all-gather form of the NCCL column mode
    full2 = frames_torch.mixed_batch(1, w2, h, dev, first_index=5, content="natural")[0]
    whole2 = eng.accumulate_device(full2[None])
    for mode, fw, ff, wh in (("rows", w, full, whole), ("cols", w, full, whole), ("cols", w2, full2, whole2)):
        tiled = pkg.sharding.TiledFrame(eng, fw, h, pkg.ScopeSettings(), mode=mode)
        a, b = tiled.my_band
        for rep in range(2):
            tiled.reset()
            if mode == "rows":
                tiled.accumulate(ff[a:b])
            else:
                band = torch.as_strided(ff.reshape(-1)[a * 4:], (h, (fw - a) * 4), (fw * 4, 1))
                tiled.accumulate(band, width=b - a)
            out = tiled.reduce_and_finalize()
            torch.cuda.synchronize()
            for k in ("hist", "wave", "vscope"):
                same = bool(torch.equal(out[k][0], wh[k][0]))
                if not same:
                    print(f"rank {rank}: TiledFrame {mode} w={fw} gather={tiled.gather_wave}: {k} differs", flush=True)
                ok = ok and same
    # frame sharding + optional gather
    n = 6
    batch = frames_torch.mixed_batch(n, 640, 360, dev)
    ref = eng.accumulate_device(batch)
    mine = pkg.sharding.frame_shard(n, rank, world)
    part = eng.accumulate_device(batch[mine.start:mine.stop].contiguous())
    got = pkg.sharding.gather_results({k: part[k] for k in ("hist", "wave", "vscope")}, dst=0)
    if rank == 0:
        for k in ("hist", "wave", "vscope"):
            cat = torch.cat([t.to(dev) for t in got[k]], dim=0)
            ok = ok and bool(torch.equal(cat, ref[k]))
    q.put((rank, ok))
    dist.destroy_process_group()
    eng.close()


def _peer_worker(rank, world, port, q):
    """PeerTiledFrame over real NVLink peer mappings (torch symmetric memory): one-shot and two-shot must equal
    the whole frame accumulated on one GPU, for two frames in a row (reset between them)."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import obs_color_monitor_b200 as pkg
    from obs_color_monitor_b200 import frames_torch
    eng = pkg.ScopeEngine(rank)
    ok = True
    w, h = 1000, 700
    st = pkg.ScopeSettings(vscope_intensity=25)
    import torch.distributed._symmetric_memory as symm_mem
    probe = symm_mem.rendezvous(symm_mem.empty(1024, dtype=torch.int32, device=dev), dist.group.WORLD)
    forms = [(False, False), (True, False)]
    if probe.multicast_ptr:                       # the NVLS forms only where the switch offers multicast
        forms += [(False, True), (True, True)]
    def band_of(full, tiled, mode):
        a, b = tiled.my_band
        if mode == "rows":
            return full[a:b], None
        return torch.as_strided(full.reshape(-1)[a * 4:], (h, (w - a) * 4), (w * 4, 1)), b - a

    for mode in ("rows", "cols"):
        for two_shot, nvls in forms:
            tiled = pkg.sharding.PeerTiledFrame(eng, w, h, st, mode=mode, two_shot=two_shot, nvls=nvls)
            for index in (3, 7):
                full = frames_torch.mixed_batch(1, w, h, dev, first_index=index, content="natural")[0]
                whole = eng.accumulate_device(full[None], settings=st)
                tiled.reset()
                band, bw = band_of(full, tiled, mode)
                tiled.accumulate(band, width=bw)
                out = tiled.reduce_and_finalize()
                torch.cuda.synchronize()
                dist.barrier()
                for k in ("hist", "hist_max", "wave", "vscope", "vscope_display"):
                    same = bool(torch.equal(out[k][0], whole[k][0]))
                    if not same:
                        print(f"rank {rank}: PeerTiledFrame {mode} two_shot={two_shot} nvls={nvls} frame {index}: {k} differs", flush=True)
                    ok = ok and same
    # BASELINE config 4 in small: luma waveform only (components 0x20); column bands = the strip kernel's own peer
    # stores + one barrier, row bands = exclusive pairs + the fused reduce
    st4 = pkg.ScopeSettings(scopes=pkg.SCOPE_WAVE, wave_components=0x20)
    for mode in ("rows", "cols"):
        tiled = pkg.sharding.PeerTiledFrame(eng, w, h, st4, mode=mode)
        for index in (3, 7):
            full = frames_torch.mixed_batch(1, w, h, dev, first_index=index, content="natural")[0]
            whole = eng.accumulate_device(full[None], settings=st4)
            tiled.reset()
            band, bw = band_of(full, tiled, mode)
            tiled.accumulate(band, width=bw)
            out = tiled.reduce_and_finalize()
            torch.cuda.synchronize()
            dist.barrier()
            same = bool(torch.equal(out["wave"][0], whole["wave"][0]))
            if not same:
                print(f"rank {rank}: config-4 {mode} frame {index}: waveform differs", flush=True)
            ok = ok and same
    q.put((rank, ok))
    dist.destroy_process_group()
    eng.close()


def test_two_gpus_peer_memory_reduce():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_two_gpus_tiled_and_sharded():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
