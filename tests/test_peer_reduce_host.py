"""The peer-memory reduce + saturate kernel (csrc/scope_peer_reduce.cuh, behind scope_finalize_peers) on the
CPU: its per-thread body is plain C++, tools/simt/peer_reduce_host.cpp runs it for every (block, thread) of
the grid with the slice / grid arithmetic of the host entry point, and the results must equal the oracle's
whole-frame outputs when the "ranks'" partials are the oracle's per-band counts.  No GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from test_sharding_gloo import _oracle_partials

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tools", "simt", "peer_reduce_host.cpp")
HDR = os.path.join(ROOT, "obs-color-monitor_b200", "csrc", "scope_peer_reduce.cuh")
LIB = os.path.join(ROOT, "tools", "simt", "_build", "libpeer_reduce_host.so")


@pytest.fixture(scope="module")
def host_lib():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-fPIC", "-shared",
                               "-I" + os.path.dirname(HDR), SRC, "-o", LIB])
    L = C.CDLL(LIB)
    pp = C.POINTER(C.c_void_p)
    L.peer_reduce_host.argtypes = [pp, pp, pp, C.c_uint32, pp, pp, pp, pp, C.c_uint32, C.c_void_p, C.c_uint32,
                                   C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_float]
    L.peer_reduce_host.restype = C.c_int
    return L


def _aligned(shape, dtype, fill=None):
    """16-byte aligned numpy array (the kernel uses 16-byte loads and stores)"""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    raw = np.empty(n + 16, np.uint8)
    off = (-raw.ctypes.data) % 16
    a = raw[off:off + n].view(dtype).reshape(shape)
    if fill is not None:
        a[...] = fill
    return a


def _ptrs(arrays):
    return (C.c_void_p * len(arrays))(*[a.ctypes.data if a is not None else None for a in arrays])


def _bands(h, n):
    edges = [h * k // n for k in range(n + 1)]
    return list(zip(edges[:-1], edges[1:]))


def _run(L, parts, outs, width, planes, do_vs, slice_index, slice_count, hist_out, max_blocks=1184,
         wave_k=0.0, vs_k=0.0):
    rc = L.peer_reduce_host(_ptrs([p[0] for p in parts]), _ptrs([p[1] for p in parts]), _ptrs([p[2] for p in parts]),
                            len(parts), _ptrs([o.get("wave") for o in outs]), _ptrs([o.get("wave_display") for o in outs]),
                            _ptrs([o.get("vscope") for o in outs]), _ptrs([o.get("vscope_display") for o in outs]),
                            len(outs), hist_out.ctypes.data if hist_out is not None else None, width, planes,
                            int(do_vs), slice_index, slice_count, max_blocks, wave_k, vs_k)
    assert rc == 0


def _frame_and_partials(pkg, oracle, w, h, n_ranks, kind):
    f = {"random": lambda: pkg.frames.random(w, h, 5), "solid": lambda: pkg.frames.solid(w, h, (10, 200, 90, 255)),
         "natural": lambda: pkg.frames.natural(w, h, 2), "alpha": lambda: pkg.frames.alpha_stripes(w, h, 3)}[kind]()
    yuv = oracle.rgb_to_yuv(f, 2)
    parts = []
    for y0, y1 in _bands(h, n_ranks):
        hist, pairs, vs = _oracle_partials(oracle, f, yuv, y0, y1, 0, w, w)
        parts.append((_aligned(hist.shape, np.int32, hist), _aligned(pairs.shape, np.int32, pairs),
                      _aligned(vs.shape, np.int32, vs)))
    return f, yuv, parts


def _alloc_out(w, display=False):
    o = {"wave": _aligned((256, w, 4), np.uint8, 0xEE), "vscope": _aligned((256, 256), np.uint8, 0xEE)}
    if display:
        o["wave_display"] = _aligned((256, w, 4), np.uint8, 0xEE)
        o["vscope_display"] = _aligned((256, 256), np.uint8, 0xEE)
    return o


@pytest.mark.parametrize("kind,w,h,n_ranks", [("random", 96, 300, 3), ("solid", 33, 700, 4), ("natural", 160, 90, 2),
                                              ("alpha", 64, 520, 8), ("random", 1, 40, 1)])
def test_one_shot_equals_whole_frame(host_lib, pkg, oracle, kind, w, h, n_ranks):
    f, yuv, parts = _frame_and_partials(pkg, oracle, w, h, n_ranks, kind)
    out = _alloc_out(w, display=True)
    hist = _aligned((1024,), np.uint32, 0)
    _run(host_lib, parts, [out], w, 2, True, 0, 1, hist, wave_k=7.0, vs_k=25.0)
    want_wave, want_vs = oracle.waveform(0x07, f, yuv), oracle.vectorscope(yuv)
    assert np.array_equal(hist, oracle.histogram_counts(0x07, f, yuv).ravel())
    assert np.array_equal(out["wave"], want_wave)
    assert np.array_equal(out["vscope"], want_vs)
    assert np.array_equal(out["wave_display"][..., :3], oracle.apply_intensity(want_wave, 7)[..., :3])
    assert np.array_equal(out["vscope_display"], oracle.apply_intensity(want_vs, 25))


@pytest.mark.parametrize("n_ranks,max_blocks", [(2, 1184), (3, 2), (8, 1184), (16, 1)])
def test_two_shot_every_rank_ends_with_the_whole_result(host_lib, pkg, oracle, n_ranks, max_blocks):
    """rank r reduces slice r of n and stores it into EVERY rank's images; after all ranks ran, every image is
    complete, every byte written exactly by one rank (the 0xEE canary is gone, nothing outside is touched)."""
    w, h = 75, 330   # 256 * 75 / 4 = 4800 quads: not divisible by 16 * 256, slices of unequal size
    f, yuv, parts = _frame_and_partials(pkg, oracle, w, h, n_ranks, "random")
    images = [_alloc_out(w) for _ in range(n_ranks)]
    hists = [_aligned((1024,), np.uint32, 0) for _ in range(n_ranks)]
    for r in range(n_ranks):
        outs = [images[r]] + [images[k] for k in range(n_ranks) if k != r]   # outs[0] = the rank's own
        _run(host_lib, parts, outs, w, 2, True, r, n_ranks, hists[r], max_blocks=max_blocks)
    want_wave, want_vs, want_hist = oracle.waveform(0x07, f, yuv), oracle.vectorscope(yuv), \
        oracle.histogram_counts(0x07, f, yuv).ravel()
    for r in range(n_ranks):
        assert np.array_equal(images[r]["wave"], want_wave)
        assert np.array_equal(images[r]["vscope"], want_vs)
        assert np.array_equal(hists[r], want_hist)


def test_slices_are_disjoint_and_single_plane_skips_plane_1(host_lib, pkg, oracle):
    """a slice touches only its own quads; with wave_planes = 1 (no R|V channel) plane 1 is never read"""
    w, n = 40, 4
    f, yuv, parts = _frame_and_partials(pkg, oracle, w, 64, 2, "random")
    for p in parts:
        p[1][1] = 0x7FFF7FFF     # poison plane 1: must not leak into the result
    out = _alloc_out(w)
    _run(host_lib, parts, [out], w, 1, False, 1, n, None)
    quads = 256 * w // 4
    q0, q1 = quads * 1 // n, quads * 2 // n
    flat = out["wave"].reshape(-1, 4)
    assert (flat[:q0 * 4] == 0xEE).all() and (flat[q1 * 4:] == 0xEE).all()
    want = oracle.waveform(0x07, f, yuv).reshape(-1, 4).copy()
    want[:, 2] = 0                                    # channel R lives in plane 1
    assert np.array_equal(flat[q0 * 4:q1 * 4], want[q0 * 4:q1 * 4])
    assert (out["vscope"] == 0xEE).all()              # vectorscope not requested


class _HostEngine:
    """stands in for ScopeEngine on the CPU: finalize_peers runs the kernel body through the host harness on the
    very addresses sharding.PeerTiledFrame hands to scope_finalize_peers"""

    def __init__(self, L, pkg):
        self.L, self.pkg, self.calls = L, pkg, []

    def alloc_device_out(self, n, width, st, dev=None):
        return self.pkg.ScopeEngine.alloc_device_out(None, n, width, st, dev)

    def finalize_peers(self, partials, outs, *, full_width, full_height, settings, slice_index=0, slice_count=1):
        self.calls.append((slice_index, slice_count, len(partials), len(outs)))

        def addr(v):
            return None if v is None else (v.data_ptr() if hasattr(v, "data_ptr") else int(v))

        def col(dicts, key):
            return (C.c_void_p * len(dicts))(*[addr(d.get(key)) for d in dicts])

        planes = 0 if "wave" not in outs[0] else (2 if settings.wave_components & 0x44 else 1)
        rc = self.L.peer_reduce_host(col(partials, "hist"), col(partials, "wave_pairs"), col(partials, "vscope"),
                                     len(partials), col(outs, "wave"), col(outs, "wave_display"), col(outs, "vscope"),
                                     col(outs, "vscope_display"), len(outs), addr(outs[0].get("hist")), full_width,
                                     planes, int("vscope" in outs[0]), slice_index, slice_count, 1184,
                                     float(settings.wave_intensity), float(settings.vscope_intensity))
        assert rc == 0


class _NoBarrier:
    def barrier(self, channel=0):
        pass


@pytest.mark.parametrize("two_shot", [False, True])
def test_peer_tiled_frame_host_logic(host_lib, pkg, oracle, two_shot):
    """sharding.PeerTiledFrame's layout of the symmetric allocation, the address lists it builds and its one- /
    two-shot choice, with three "ranks" in one process: their buffers stand in for the peer mappings, the
    barriers are trivial because the ranks run one after the other."""
    import torch
    w, h, n = 75, 330, 3
    st = pkg.ScopeSettings(vscope_intensity=25)
    f, yuv, parts = _frame_and_partials(pkg, oracle, w, h, n, "natural")
    eng = _HostEngine(host_lib, pkg)
    ranks = [pkg.sharding.PeerTiledFrame(eng, w, h, st, mode="rows", device=torch.device("cpu")) for _ in range(n)]
    bases = [t._buf.data_ptr() for t in ranks]
    assert all(b % 16 == 0 for b in bases)
    for r, t in enumerate(ranks):         # what the rendezvous would have set up
        t.rank, t.world, t._bases, t._hdl, t.two_shot = r, n, bases, _NoBarrier(), two_shot
        t.bands = pkg.sharding.row_bands(h, n)
    assert ranks[1].my_band == (110, 220)
    for frame in range(2):                # the second frame checks reset()
        for t, (hist, pairs, vs) in zip(ranks, parts):
            if frame:
                assert all(bool(v.any()) for v in t.partial.values())
            t.reset()     # zeroes what accumulate ADDS to; the waveform pairs are stored whole (SCOPE_BAND_EXCLUSIVE)
            assert not bool(t.partial["hist"].any()) and not bool(t.partial["vscope"].any())
            t.partial["hist"].copy_(torch.from_numpy(hist))
            t.partial["wave_pairs"].copy_(torch.from_numpy(pairs))
            t.partial["vscope"].copy_(torch.from_numpy(vs))
        for t in ranks:
            t.start_reduce()
        want_wave, want_vs = oracle.waveform(0x07, f, yuv), oracle.vectorscope(yuv)
        for t in ranks:
            out = t.finish()
            assert np.array_equal(out["wave"][0].numpy(), want_wave)
            assert np.array_equal(out["vscope"][0].numpy(), want_vs)
            assert np.array_equal(out["vscope_display"][0].numpy(), oracle.apply_intensity(want_vs, 25))
            assert np.array_equal(out["hist"][0].numpy().view(np.uint32), oracle.histogram_counts(0x07, f, yuv).ravel())
    assert eng.calls[:3] == ([(0, 3, 3, 3), (1, 3, 3, 3), (2, 3, 3, 3)] if two_shot else [(0, 1, 3, 1)] * 3)


def test_random_partials_against_numpy(host_lib):
    """arbitrary partial arrays (not derived from a frame): u16 halves up to the largest total a frame allows
    (65535 rows over all ranks), vectorscope and histogram counts up to 2^31 in total; every slicing"""
    rng = np.random.default_rng(11)
    for case in range(12):
        n = int(rng.integers(1, 17))
        w = int(rng.integers(1, 70))
        planes = int(rng.integers(1, 3))
        cap = 65535 // n
        parts = []
        for _ in range(n):
            small = rng.integers(0, 3, size=(2, 256, w, 2)) * rng.integers(0, 200, size=(2, 256, w, 2))
            big = rng.integers(0, cap + 1, size=(2, 256, w, 2))
            halves = np.where(rng.random((2, 256, w, 2)) < 0.1, big, small).astype(np.uint32)
            halves[1, ..., 1] = 0                                   # plane 1 holds R|V only
            pairs = halves[..., 0] | (halves[..., 1] << 16)
            vs = (rng.integers(0, 2, size=65536) * rng.integers(0, (2 ** 31) // n, size=65536)).astype(np.uint32)
            hist = rng.integers(0, (2 ** 31) // n, size=1024).astype(np.uint32)
            parts.append((_aligned(hist.shape, np.uint32, hist), _aligned(pairs.shape, np.uint32, pairs),
                          _aligned(vs.shape, np.uint32, vs)))
        tot = sum(p[1].astype(np.uint64) for p in parts)
        lo, hi, r = tot[0] & 0xFFFF, tot[0] >> 16, tot[1] & 0xFFFF
        want = np.zeros((256, w, 4), np.uint8)
        want[..., 0], want[..., 1] = np.minimum(lo, 255), np.minimum(hi, 255)
        if planes == 2:
            want[..., 2] = np.minimum(r, 255)
        want_vs = np.minimum(sum(p[2].astype(np.uint64) for p in parts), 255).astype(np.uint8).reshape(256, 256)
        want_hist = sum(p[0].astype(np.uint64) for p in parts).astype(np.uint32)
        slices = int(rng.integers(1, 6))
        out, hist_out = _alloc_out(w), _aligned((1024,), np.uint32, 0)
        for k in range(slices):
            _run(host_lib, parts, [out], w, planes, True, k, slices, hist_out, max_blocks=int(rng.integers(1, 9)))
        assert np.array_equal(out["wave"], want), case
        assert np.array_equal(out["vscope"], want_vs), case
        assert np.array_equal(hist_out, want_hist), case
