"""(Named _z_ so that it runs after the established GPU suites: it was written when this round's GPU budget was
already spent and meets the hardware for the first time in the round-end pass.)

GPU parity of scope_finalize_peers (reduce + saturate over peer memory in one kernel).
On one GPU the "peers" are separate allocations of the same device - the kernel only sees addresses - so the
arithmetic, the slicing and the stores into several receivers are checked bit-exact against the oracle here;
tests/test_gpu_multirank.py covers real peer mappings when more than one GPU is visible."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W, H = 260, 300
BANDS = [(0, 77), (77, 200), (200, 300)]


def _partials(engine, pkg, d, st):
    parts = []
    for y0, y1 in BANDS:
        part = engine.alloc_partial(W)
        engine.accumulate_partial(d[y0:y1], part, x_offset=0, full_width=W, settings=st)
        parts.append(part)
    return parts


def _frame(pkg):
    f = pkg.frames.natural(W, H, seed=2)
    f[:, :, 3] = 255
    f[10:40, 5:9, 3] = 0
    f[20:, 64:96, :3] = (10, 200, 97)      # a flat patch of 280 rows: column bins above 255 in the sum, below in every band
    return f


def _check(out, oracle, f, yuv, st):
    want_wave, want_vs = oracle.waveform(st.wave_components, f, yuv), oracle.vectorscope(yuv)
    assert np.array_equal(out["wave"][0].cpu().numpy(), want_wave)
    if "vscope" in out:
        assert np.array_equal(out["vscope"][0].cpu().numpy(), want_vs)
    if "wave_display" in out:
        assert np.array_equal(out["wave_display"][0].cpu().numpy(), oracle.apply_intensity(want_wave, st.wave_intensity))
    if "vscope_display" in out:
        assert np.array_equal(out["vscope_display"][0].cpu().numpy(), oracle.apply_intensity(want_vs, st.vscope_intensity))


def test_one_shot_equals_whole_frame_and_finalize_partial(engine, oracle, pkg):
    import torch
    f = _frame(pkg)
    d = torch.from_numpy(f).cuda()
    yuv = oracle.rgb_to_yuv(f, 2)
    st = pkg.ScopeSettings(wave_intensity=7, vscope_intensity=25)
    parts = _partials(engine, pkg, d, st)
    out = engine.alloc_device_out(1, W, st)
    for t in out.values():
        t.fill_(0x6E)
    before = engine.launch_count
    engine.finalize_peers(parts, [out], full_width=W, full_height=H, settings=st)
    torch.cuda.synchronize()
    assert engine.launch_count - before == 2          # the fused kernel + hist_max
    hist = oracle.histogram_counts(7, f, yuv)
    assert np.array_equal(out["hist"][0].cpu().numpy().view(np.uint32), hist.ravel())
    _, hi = oracle.histogram_post(7, W, H, hist)
    assert np.array_equal(out["hist_max"][0, :3].cpu().numpy().view(np.uint32), np.asarray(hi, np.uint32))
    _check(out, oracle, f, yuv, st)
    # and the same bytes as the NCCL-style path: sum the partials, then scope_finalize_partial
    summed = {k: sum(p[k] for p in parts) for k in parts[0]}
    ref = engine.finalize_partial(summed, full_width=W, full_height=H, settings=st)
    torch.cuda.synchronize()
    for k in ref:
        assert torch.equal(ref[k], out[k]), k


@pytest.mark.parametrize("wave_components", [0x07, 0x20])
def test_two_shot_every_receiver_complete(engine, oracle, pkg, wave_components):
    """rank r reduces slice r of 3 and stores into all three receivers; 0x20 (luma waveform alone, BASELINE
    config 4) takes the one-plane path"""
    import torch
    f = _frame(pkg)
    d = torch.from_numpy(f).cuda()
    yuv = oracle.rgb_to_yuv(f, 2)
    st = pkg.ScopeSettings(wave_components=wave_components,
                           scopes=pkg.SCOPE_ALL if wave_components == 0x07 else pkg.SCOPE_WAVE)
    parts = _partials(engine, pkg, d, st)
    n = len(parts)
    images = [engine.alloc_device_out(1, W, st) for _ in range(n)]
    for o in images:
        for t in o.values():
            t.fill_(0x6E)
    for r in range(n):
        outs = [images[r]] + [images[k] for k in range(n) if k != r]
        engine.finalize_peers(parts, outs, full_width=W, full_height=H, settings=st, slice_index=r, slice_count=n)
    torch.cuda.synchronize()
    hist = oracle.histogram_counts(7, f, yuv).ravel()
    for o in images:
        _check(o, oracle, f, yuv, st)
        if "hist" in o:
            assert np.array_equal(o["hist"][0].cpu().numpy().view(np.uint32), hist)


def test_addresses_instead_of_tensors_and_errors(engine, oracle, pkg):
    """plain device addresses (what a symmetric-memory handle's buffer_ptrs are) work like tensors; bad requests
    fail with the documented codes"""
    import torch
    f = _frame(pkg)
    d = torch.from_numpy(f).cuda()
    yuv = oracle.rgb_to_yuv(f, 2)
    st = pkg.ScopeSettings(scopes=pkg.SCOPE_WAVE | pkg.SCOPE_VSCOPE)
    parts = _partials(engine, pkg, d, st)
    as_addr = [{k: t.data_ptr() for k, t in p.items()} for p in parts]
    out = engine.alloc_device_out(1, W, st)
    engine.finalize_peers(as_addr, [{k: t.data_ptr() for k, t in out.items()}], full_width=W, full_height=H, settings=st)
    torch.cuda.synchronize()
    _check(out, oracle, f, yuv, st)
    E = pkg._ffi
    with pytest.raises(pkg.ScopeError) as e:
        engine.finalize_peers(parts * 6, [out], full_width=W, full_height=H, settings=st)       # 18 partials
    assert e.value.code == E.SCOPE_ERR_UNSUPPORTED
    with pytest.raises(pkg.ScopeError) as e:
        engine.finalize_peers(parts, [out], full_width=W, full_height=H, settings=st, slice_index=3, slice_count=3)
    assert e.value.code == E.SCOPE_ERR_INVALID
    with pytest.raises(pkg.ScopeError) as e:                                                      # misaligned partial
        bad = [dict(p) for p in as_addr]
        bad[1]["vscope"] += 4
        engine.finalize_peers(bad, [out], full_width=W, full_height=H, settings=st)
    assert e.value.code == E.SCOPE_ERR_INVALID


def test_peer_tiled_frame_single_rank(engine, oracle, pkg):
    """sharding.PeerTiledFrame without a process group: the same kernel on local memory, bands fed one by one"""
    import torch
    f = _frame(pkg)
    d = torch.from_numpy(f).cuda()
    yuv = oracle.rgb_to_yuv(f, 2)
    st = pkg.ScopeSettings(vscope_intensity=25)
    tiled = pkg.sharding.PeerTiledFrame(engine, W, H, st, mode="rows")
    assert tiled.world == 1 and tiled.my_band == (0, H) and not tiled.two_shot
    for _ in range(2):                                   # twice: reset() must give a clean second frame
        tiled.reset(additive=True)                       # accumulate_partial ADDS its waveform pairs
        for y0, y1 in BANDS:
            engine.accumulate_partial(d[y0:y1], tiled.partial, x_offset=0, full_width=W, settings=st)
        out = tiled.reduce_and_finalize()
        torch.cuda.synchronize()
        _check(out, oracle, f, yuv, st)
        assert np.array_equal(out["hist"][0].cpu().numpy().view(np.uint32), oracle.histogram_counts(7, f, yuv).ravel())


def test_accumulate_band_exclusive_and_direct_columns(engine, oracle, pkg):
    """scope_accumulate_band: (a) SCOPE_BAND_EXCLUSIVE - a band's waveform pairs are STORED into accumulators that
    hold garbage (no zero-fill), row bands then summed by scope_finalize_peers; (b) column bands spanning the full
    height write their FINAL waveform columns into several images at once (the peers' images on a multi-GPU box,
    two local allocations here) while histogram and vectorscope go through partials."""
    import torch
    f = _frame(pkg)
    d = torch.from_numpy(f).cuda()
    yuv = oracle.rgb_to_yuv(f, 2)
    st = pkg.ScopeSettings()
    # (a) row bands, each into its own garbage-filled partial
    parts = []
    for y0, y1 in BANDS:
        part = engine.alloc_partial(W)
        part["wave_pairs"].fill_(0x5A5A5A5A)
        engine.accumulate_band(d[y0:y1], part, x_offset=0, full_width=W, exclusive=True, settings=st)
        parts.append(part)
    out = engine.alloc_device_out(1, W, st)
    engine.finalize_peers(parts, [out], full_width=W, full_height=H, settings=st)
    torch.cuda.synchronize()
    _check(out, oracle, f, yuv, st)
    assert np.array_equal(out["hist"][0].cpu().numpy().view(np.uint32), oracle.histogram_counts(7, f, yuv).ravel())
    # (b) column bands -> two images at once; luma-only (config 4's components) and RGB
    for comp in (0x20, 0x07):
        stc = pkg.ScopeSettings(wave_components=comp)
        imgs = [torch.full((1, 256, W, 4), 0x6E, dtype=torch.uint8, device="cuda") for _ in range(2)]
        part = engine.alloc_partial(W)
        for x0, x1 in [(0, 64), (64, 96), (96, W)]:
            tile = torch.as_strided(d.reshape(-1)[x0 * 4:], (H, (W - x0) * 4), (W * 4, 1))
            engine.accumulate_band(tile, {"hist": part["hist"], "vscope": part["vscope"]}, x_offset=x0, full_width=W,
                                   wave_outs=imgs, settings=stc, width=x1 - x0)
        torch.cuda.synchronize()
        want = oracle.waveform(comp, f, yuv)
        for im in imgs:
            assert np.array_equal(im[0].cpu().numpy(), want), hex(comp)
        assert np.array_equal(part["hist"].cpu().numpy().view(np.uint32), oracle.histogram_counts(7, f, yuv).ravel())
        assert np.array_equal(np.minimum(part["vscope"].cpu().numpy().view(np.uint32), 255).astype(np.uint8).reshape(256, 256),
                              oracle.vectorscope(yuv))
