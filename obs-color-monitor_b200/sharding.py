"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing).

Two ways the path shards; the reference has neither (it is single threaded):

* frame sharding  — frames are independent (every frame re-zeroes its bins:
  histogram.c:363-365, waveform.c:225-226, vectorscope.c:219-220), so rank r takes a
  contiguous share of the batch and NO data-path collective is needed; `gather_results`
  is only for callers that want everything on one rank.
* tile sharding of ONE frame — row bands (or column bands) per rank; the partial bins are
  additive, so the ranks all-reduce them (NCCL sum over int32 lanes) and then apply the
  saturation the reference applies per increment: min(sum of partials, 255).
  With column bands the waveform needs no reduction at all (disjoint columns).

Nothing here touches the oracle; the accumulation itself is `ScopeEngine.accumulate_partial`.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

STRIP = 32  # the kernels work in strips of 32 columns; column bands are aligned to it


def frame_shard(n_frames: int, rank: int, world: int) -> range:
    """Contiguous share of `n_frames` for `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_frames, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def row_bands(height: int, world: int) -> List[Tuple[int, int]]:
    """[y0, y1) per rank; bands differ by at most one row."""
    return [(r.start, r.stop) for r in (frame_shard(height, k, world) for k in range(world))]


def col_bands(width: int, world: int) -> List[Tuple[int, int]]:
    """[x0, x1) per rank, boundaries on multiples of 32 columns (whole strips per rank)."""
    strips = (width + STRIP - 1) // STRIP
    out = []
    for k in range(world):
        s = frame_shard(strips, k, world)
        out.append((min(s.start * STRIP, width), min(s.stop * STRIP, width)))
    return out


def allreduce_partials(partial: Dict[str, "torch.Tensor"], keys=("hist", "wave_pairs", "vscope"), group=None,
                       wave_planes: int = 2):
    """In-place sum of the partial accumulators over all ranks.  The tensors are int32 views of
    u32 counts / u16x2 pairs: every lane stays below 2^31 (hist: <= W*H, pairs: each u16 half
    <= rows of the whole frame <= 65535, vscope: <= W*H), so integer addition is exact and carries
    never cross a u16 half."""
    import torch.distributed as dist

    works = []
    for k in keys:
        if k not in partial:
            continue
        # waveform plane 1 only holds the R|V channel: skip it when that channel is off
        t = partial[k][:wave_planes] if k == "wave_pairs" else partial[k]
        works.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=True))
    for w in works:
        w.wait()
    return partial


def gather_results(out: Dict[str, "torch.Tensor"], dst: int = 0, group=None) -> Optional[Dict[str, list]]:
    """Optional: collect every rank's per-frame outputs on `dst` (frame-sharded jobs)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    res = {}
    for k, t in out.items():
        bufs = [torch.empty_like(t) for _ in range(world)] if dist.get_rank(group) == dst else None
        dist.gather(t, bufs, dst=dst, group=group)
        if bufs is not None:
            res[k] = bufs
    return res if dist.get_rank(group) == dst else None


class TiledFrame:
    """One frame split across ranks, cross-rank step by NCCL.  Each rank calls `accumulate(band_tensor)` with ITS
    band (device tensor) and then `reduce_and_finalize()`; every rank ends with the full result.

    rows: every rank holds partial sums for all columns -> all-reduce (sum) of the partials, then saturate.
    cols: the waveform columns of a rank are FINAL (its band spans the full height): they are written as u8 and
          all-gathered (1/4 of the bytes of the u16-pair all-reduce, and no clamp pass); histogram and vectorscope
          partials (260 KB) are still all-reduced.  Needs bands of equal width (else the waveform falls back to the
          all-reduce of the pairs, which merges disjoint columns exactly as well)."""

    def __init__(self, engine, full_width: int, full_height: int, settings, mode: str = "rows", group=None):
        import torch.distributed as dist

        assert mode in ("rows", "cols")
        assert full_height <= 65535, "u16 halves of the partial waveform: a frame has at most 65535 rows"
        self.engine, self.settings, self.mode, self.group = engine, settings, mode, group
        self.width, self.height = full_width, full_height
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.bands = row_bands(full_height, self.world) if mode == "rows" else col_bands(full_width, self.world)
        self.partial = engine.alloc_partial(full_width)
        from ._ffi import SCOPE_WAVE
        widths = {b - a for a, b in self.bands}
        self.gather_wave = (mode == "cols" and bool(settings.scopes & SCOPE_WAVE) and len(widths) == 1
                            and settings.wave_intensity == 0)
        if self.gather_wave:
            import torch
            bw = self.bands[0][1] - self.bands[0][0]
            dev = self.partial["hist"].device
            self._band_img = torch.empty((256, bw, 4), dtype=torch.uint8, device=dev)
            self._gathered = torch.empty((self.world, 256, bw, 4), dtype=torch.uint8, device=dev)

    @property
    def my_band(self) -> Tuple[int, int]:
        return self.bands[self.rank]

    def _other_keys(self):
        from ._ffi import SCOPE_HIST, SCOPE_VSCOPE

        keys = []
        if self.settings.scopes & SCOPE_HIST:
            keys.append("hist")
        if self.settings.scopes & SCOPE_VSCOPE:
            keys.append("vscope")
        return keys

    def reset(self, additive: bool = False):
        """zero what the next accumulate() ADDS to.  A band accumulated in one call stores its waveform pairs
        (SCOPE_BAND_EXCLUSIVE) or writes final u8 columns, so the 2 x 256 x W pairs need no zero-fill.
        additive=True: the caller feeds several tiles per rank through `engine.accumulate_partial`, which ADDS its
        pairs: zero them as well."""
        for k in self._other_keys():
            self.partial[k].zero_()
        if additive:
            self.partial["wave_pairs"].zero_()

    def accumulate(self, band, width: Optional[int] = None):
        """band: this rank's rows (rows mode: (h_band, W, 4)) or columns (cols mode: a
        (H, linesize) byte view starting at column x0, with `width` = x1 - x0).  ONE call per frame and rank."""
        a, b = self.my_band
        others = {k: self.partial[k] for k in self._other_keys()}
        if self.mode == "rows":
            part = dict(others, wave_pairs=self.partial["wave_pairs"])
            self.engine.accumulate_band(band, part, x_offset=0, full_width=self.width, exclusive=True,
                                        settings=self.settings, width=width)
        elif self.gather_wave:
            # final u8 columns of this band into a band-wide image (row length = band width)
            self.engine.accumulate_band(band, others or None, x_offset=0, full_width=b - a,
                                        wave_outs=[self._band_img], settings=self.settings,
                                        width=(b - a) if width is None else width)
        else:
            part = dict(others, wave_pairs=self.partial["wave_pairs"])
            self.engine.accumulate_band(band, part, x_offset=a, full_width=self.width, exclusive=True,
                                        settings=self.settings, width=(b - a) if width is None else width)

    def _reduce_keys(self):
        from ._ffi import SCOPE_WAVE

        keys = self._other_keys()
        if (self.settings.scopes & SCOPE_WAVE) and not self.gather_wave:
            keys.append("wave_pairs")
        return keys

    def start_reduce(self):
        """Enqueue the collectives of this frame without waiting for them, so the caller
        can accumulate the next frame (into another TiledFrame) while NCCL runs."""
        import torch.distributed as dist

        self._works = []
        if self.world > 1:
            planes = 2 if (self.settings.wave_components & 0x44) else 1
            for k in self._reduce_keys():
                t = self.partial[k][:planes] if k == "wave_pairs" else self.partial[k]
                if k == "wave_pairs" and self.mode == "cols":
                    # stored (not added) pairs: columns of other ranks hold stale data -> zero them first
                    a, b = self.my_band
                    t[:, :, :a].zero_()
                    t[:, :, b:].zero_()
                self._works.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            if self.gather_wave:
                self._works.append(dist.all_gather_into_tensor(self._gathered, self._band_img, group=self.group,
                                                               async_op=True))
        elif self.gather_wave:
            self._gathered[0].copy_(self._band_img)

    def finish(self):
        """Wait for start_reduce() (stream-side) and saturate into the reference layouts."""
        for w in getattr(self, "_works", []):
            w.wait()
        self._works = []
        if not self.gather_wave:
            return self.engine.finalize_partial(self.partial, full_width=self.width, full_height=self.height,
                                                settings=self.settings)
        from dataclasses import replace
        from ._ffi import SCOPE_WAVE

        st = replace(self.settings, scopes=self.settings.scopes & ~SCOPE_WAVE)
        out = (self.engine.finalize_partial(self.partial, full_width=self.width, full_height=self.height, settings=st)
               if st.scopes else {})
        # [rank][256][bw][4] -> [256][rank * bw][4]: the bands side by side
        out["wave"] = self._gathered.permute(1, 0, 2, 3).reshape(1, 256, self.width, 4)
        return out

    def reduce_and_finalize(self):
        self.start_reduce()
        return self.finish()


class PeerTiledFrame(TiledFrame):
    """TiledFrame whose cross-rank step is ONE kernel over NVLink peer memory instead of an NCCL all-reduce
    followed by a clamp (`scope_finalize_peers`, csrc/scope_peer_reduce.cuh).

    The partial accumulators and the u8 result images of every rank live in one symmetric-memory allocation
    (`torch.distributed._symmetric_memory`: torch only provides the peer mappings and the device-side barrier).
    `reduce_and_finalize()` = barrier -> kernel -> barrier, all on the current stream:

    * two-shot (default for world > 2): rank r sums slice r of the bins from every peer's partials (16-byte peer
      loads), saturates and stores the u8 slice into every rank's images (peer stores) - 1/world of the reads;
    * one-shot: every rank reads all partials and writes only its own images.

    `nvls=True` lets the NVSwitch do the sum and the distribution (`scope_finalize_multicast`: `multimem.ld_reduce`
    on the allocation's multicast address, `multimem.st` for the two-shot stores).
    With one rank (or no process group) the same kernel runs on local memory."""

    def __init__(self, engine, full_width: int, full_height: int, settings, mode: str = "rows", group=None,
                 two_shot: Optional[bool] = None, device=None, nvls: bool = False):
        import torch
        import torch.distributed as dist

        assert mode in ("rows", "cols")
        assert full_height <= 65535, "u16 halves of the partial waveform: a frame has at most 65535 rows"
        self.engine, self.settings, self.mode, self.group = engine, settings, mode, group
        self.width, self.height = full_width, full_height
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.bands = row_bands(full_height, self.world) if mode == "rows" else col_bands(full_width, self.world)
        self.gather_wave = False
        from ._ffi import SCOPE_WAVE as _W
        # column bands: the strip kernel stores its final waveform columns straight into every rank's image
        self.direct_wave = mode == "cols" and bool(settings.scopes & _W) and settings.wave_intensity == 0
        self.two_shot = (self.world > 2) if two_shot is None else (bool(two_shot) and self.world > 1)
        self.nvls = bool(nvls) and self.world > 1
        self._mc_base = 0
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())

        # layout of the symmetric allocation, in int32 words; every section starts on a 16-byte boundary
        W = full_width
        sections = [("hist", 1024), ("vscope", 65536), ("wave_pairs", 2 * 256 * W),
                    ("out_wave", 256 * W), ("out_vscope", 65536 // 4)]
        if settings.wave_intensity > 0:
            sections.append(("out_wave_display", 256 * W))
        if settings.vscope_intensity > 0:
            sections.append(("out_vscope_display", 65536 // 4))
        self._off, words = {}, 0
        for name, n in sections:
            self._off[name] = words
            words += (n + 3) // 4 * 4
        if self.world > 1:
            import torch.distributed._symmetric_memory as symm_mem

            self._buf = symm_mem.empty(words, dtype=torch.int32, device=dev)
            self._hdl = symm_mem.rendezvous(self._buf, group if group is not None else dist.group.WORLD)
            self._bases = [int(p) for p in self._hdl.buffer_ptrs]
            if nvls:
                self._mc_base = int(self._hdl.multicast_ptr or 0)
                if not self._mc_base:
                    raise RuntimeError("PeerTiledFrame(nvls=True): the symmetric allocation has no multicast mapping "
                                       "(no NVSwitch multicast support on this system)")
        else:
            self._buf = torch.empty(words, dtype=torch.int32, device=dev)
            self._hdl = None
            self._bases = [self._buf.data_ptr()]
        self._buf.zero_()

        def view(name, n):
            return self._buf[self._off[name]:self._off[name] + n]

        self.partial = {"hist": view("hist", 1024), "wave_pairs": view("wave_pairs", 2 * 256 * W).view(2, 256, W),
                        "vscope": view("vscope", 65536)}
        # this rank's results, shaped like ScopeEngine.alloc_device_out(1, ...)
        self.out = engine.alloc_device_out(1, W, settings, dev)
        if "wave" in self.out:
            self.out["wave"] = view("out_wave", 256 * W).view(torch.uint8).view(1, 256, W, 4)
        if "vscope" in self.out:
            self.out["vscope"] = view("out_vscope", 65536 // 4).view(torch.uint8).view(1, 256, 256)
        if "wave_display" in self.out:
            self.out["wave_display"] = view("out_wave_display", 256 * W).view(torch.uint8).view(1, 256, W, 4)
        if "vscope_display" in self.out:
            self.out["vscope_display"] = view("out_vscope_display", 65536 // 4).view(torch.uint8).view(1, 256, 256)

    def _addresses(self, rank: int, names) -> Dict[str, int]:
        return {key: self._bases[rank] + 4 * self._off[sec] for key, sec in names if sec in self._off}

    def reset(self, additive: bool = False):
        """zero what the next accumulate() ADDS to: histogram and vectorscope partials (the first two sections of the
        allocation, one memset of 260 KB) when those scopes are on.  The waveform pairs are stored, not added
        (SCOPE_BAND_EXCLUSIVE), or not used at all (column bands): no zero-fill - unless the caller feeds several
        tiles per rank through `engine.accumulate_partial` (additive=True: one memset over all three sections)."""
        if additive:
            self._buf[:self._off["out_wave"]].zero_()
        elif self._other_keys():
            self._buf[:self._off["wave_pairs"]].zero_()

    def accumulate(self, band, width: Optional[int] = None):
        a, b = self.my_band
        others = {k: self.partial[k] for k in self._other_keys()}
        if self.mode == "rows" or not self.direct_wave:
            part = dict(others, wave_pairs=self.partial["wave_pairs"])
            self.engine.accumulate_band(band, part, x_offset=0 if self.mode == "rows" else a, full_width=self.width,
                                        exclusive=True, settings=self.settings,
                                        width=width if self.mode == "rows" else ((b - a) if width is None else width))
            return
        # column bands: my columns of EVERY rank's image, mine first (local), the others over NVLink
        off = 4 * self._off["out_wave"]
        outs = [self._bases[self.rank] + off] + [self._bases[r] + off for r in range(self.world) if r != self.rank]
        self.engine.accumulate_band(band, others or None, x_offset=a, full_width=self.width, wave_outs=outs,
                                    settings=self.settings, width=(b - a) if width is None else width)

    def finish(self):
        """results of the last start_reduce() (everything is stream-ordered; nothing to wait for on the host)"""
        return self.out

    def reduce_and_finalize(self):
        self.start_reduce()
        return self.out

    def start_reduce(self):
        """Enqueue barrier -> reduce/saturate/distribute kernel -> barrier on the current stream."""
        part_names = [("hist", "hist"), ("wave_pairs", "wave_pairs"), ("vscope", "vscope")]
        out_names = [("wave", "out_wave"), ("vscope", "out_vscope"), ("wave_display", "out_wave_display"),
                     ("vscope_display", "out_vscope_display")]
        if self._hdl is not None:
            self._hdl.barrier(channel=0)         # every rank's accumulate is complete and visible
        st = self.settings
        if self.direct_wave:
            # the waveform is already in every rank's image (the barrier above completes it); what is left is the
            # 260 KB of histogram / vectorscope partials, if those scopes are on
            from dataclasses import replace
            from ._ffi import SCOPE_WAVE
            st = replace(st, scopes=st.scopes & ~SCOPE_WAVE)
            if not st.scopes:
                return
            part_names = [pn for pn in part_names if pn[0] != "wave_pairs"]
            out_names = [on for on in out_names if not on[0].startswith("wave")]
        partials = [self._addresses(r, part_names) for r in range(self.world)]
        mine = dict(self._addresses(self.rank, out_names))
        for k in ("hist", "hist_max"):
            if k in self.out:
                mine[k] = self.out[k]
        mine = {k: v for k, v in mine.items() if k in self.out and not (self.direct_wave and k.startswith("wave"))}
        if self.nvls:
            # the switch sums (multimem.ld_reduce) and, two-shot, replicates the slice into every rank's images
            mc = self._mc_base
            mc_partial = {key: mc + 4 * self._off[sec] for key, sec in part_names}
            for key in ("hist", "wave_pairs", "vscope"):
                mc_partial.setdefault(key, None)
            mc_images = ({key: mc + 4 * self._off[sec] for key, sec in out_names if sec in self._off and key in self.out}
                         if self.two_shot else None)
            self.engine.finalize_multicast(mc_partial, mine, mc_images, full_width=self.width, full_height=self.height,
                                           settings=st, slice_index=self.rank if self.two_shot else 0,
                                           slice_count=self.world if self.two_shot else 1)
            self._hdl.barrier(channel=1)
            return
        outs = [mine]
        if self.two_shot:
            outs += [{k: v for k, v in self._addresses(r, out_names).items() if k in self.out}
                     for r in range(self.world) if r != self.rank]
        self.engine.finalize_peers(partials, outs, full_width=self.width, full_height=self.height,
                                   settings=st, slice_index=self.rank if self.two_shot else 0,
                                   slice_count=self.world if self.two_shot else 1)
        if self._hdl is not None:
            self._hdl.barrier(channel=1)         # peers are done reading my partials / writing my images
