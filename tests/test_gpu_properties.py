"""Full-size GPU checks that need no oracle: size-independent properties of the scopes at the
sizes BASELINE.json names (4K, 8K), plus host-path == device-path and batch consistency."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _batch(pkg, n, w, h, content="mixed"):
    import torch
    from obs_color_monitor_b200 import frames_torch
    return frames_torch.mixed_batch(n, w, h, torch.device("cuda", 0), content=content)


@pytest.mark.parametrize("w,h", [(3840, 2160), (7680, 4320)])
def test_totals_and_saturation_full_size(engine, pkg, w, h):
    import torch
    d = _batch(pkg, 4, w, h)
    out = engine.accumulate_device(d)
    torch.cuda.synchronize()
    hist = out["hist"].to(torch.int64).reshape(4, 256, 4)
    assert bool((hist[..., :3].sum(dim=1) == w * h).all()) and bool((hist[..., 3] == 0).all())
    # waveform: every column saw h pixels per channel; bins saturate at 255; byte 3 stays 0
    wave = out["wave"]
    assert bool((wave[..., 3] == 0).all())
    col_sum = wave[..., :3].to(torch.int64).sum(dim=1)            # (n, W, 3)
    assert int(col_sum.max()) <= h and int(col_sum.min()) >= 255
    # solid frame (index 2): exactly one vectorscope bin, saturated; one waveform row per channel
    vs = out["vscope"]
    assert int(torch.count_nonzero(vs[2])) == 1 and int(vs[2].max()) == 255
    assert int(torch.count_nonzero(wave[2, :, :, 0].to(torch.int64).sum(dim=1))) == 1
    # histogram == column sums of the (unsaturated) waveform wherever nothing saturated: ramp frame
    ramp_wave = wave[1].to(torch.int64)
    unsat = ramp_wave.amax() < 255
    if bool(unsat):
        assert bool((ramp_wave[:, :, 0].sum(dim=1).flip(0) == hist[1, :, 2]).all())


def test_tiles_add_up_to_the_frame_4k(engine, pkg):
    """linearity: row bands + column bands accumulated as partial tiles == whole-frame call"""
    import torch
    w, h = 3840, 2160
    d = _batch(pkg, 1, w, h, "natural")[0]
    whole = engine.accumulate_device(d[None])
    for cuts in ([(0, 700), (700, 701), (701, 2160)],):
        part = engine.alloc_partial(w)
        for y0, y1 in cuts:
            engine.accumulate_partial(d[y0:y1], part, x_offset=0, full_width=w)
        out = engine.finalize_partial(part, full_width=w, full_height=h)
        torch.cuda.synchronize()
        for k in ("hist", "wave", "vscope"):
            assert torch.equal(out[k][0], whole[k][0]), k


def test_host_path_equals_device_path_4k(engine, pkg):
    import torch
    d = _batch(pkg, 2, 3840, 2160, "random")
    dev = engine.accumulate_device(d, settings=pkg.ScopeSettings(vscope_intensity=25, wave_intensity=51))
    torch.cuda.synchronize()
    for i in range(2):
        res = engine.accumulate_host(d[i].cpu().numpy(), settings=pkg.ScopeSettings(vscope_intensity=25, wave_intensity=51))
        assert np.array_equal(res["hist"], dev["hist"][i].cpu().numpy().view(np.uint32))
        assert np.array_equal(res["wave"], dev["wave"][i].cpu().numpy())
        assert np.array_equal(res["vscope"], dev["vscope"][i].cpu().numpy())
        assert np.array_equal(res["vscope_display"], dev["vscope_display"][i].cpu().numpy())
        assert np.array_equal(res["wave_display"], dev["wave_display"][i].cpu().numpy())
        assert np.array_equal(res["hist_max"], dev["hist_max"][i, :3].cpu().numpy().view(np.uint32))


def test_batch_order_and_repeat_invariance(engine, pkg):
    """frames are independent: permuting the batch permutes the outputs; a second pass over the
    same batch gives identical bytes (no state leaks between launches or frames)"""
    import torch
    d = _batch(pkg, 8, 1920, 1080)
    a = engine.accumulate_device(d)
    b = engine.accumulate_device(d)
    perm = torch.tensor([3, 0, 7, 1, 6, 2, 5, 4], device=d.device)
    c = engine.accumulate_device(d[perm].contiguous())
    torch.cuda.synchronize()
    for k in ("hist", "wave", "vscope"):
        assert torch.equal(a[k], b[k]), k
        assert torch.equal(a[k][perm], c[k]), k


def test_ring_stream_matches_sync(engine, pkg):
    fr = pkg.frames
    frames = [fr.mixed(640, 360, i) for i in range(7)]
    st = pkg.ScopeSettings()
    sync = [engine.accumulate_host(f, settings=st) for f in frames]
    got = []
    for i, f in enumerate(frames):
        sl = i % 3
        if i >= 3:
            got.append(engine.wait_host(sl))
        assert engine.submit_host(sl, f, settings=st) is True
    assert engine.submit_host((len(frames)) % 3, frames[0], settings=st) is False     # slot busy -> dropped
    for i in range(len(frames) - 3, len(frames)):
        got.append(engine.wait_host(i % 3))
    for a, b in zip(sync, got):
        for k in ("hist", "wave", "vscope", "hist_float", "hist_max"):
            assert np.array_equal(a[k], b[k]), k
