#!/usr/bin/env python
"""bench.py — frames/s of the fused histogram+waveform+vectorscope pass on 3840x2160 BGRA.

    python bench.py --gpus N --steps K --warmup W            (this repo's CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  (the reference's CPU loops)

One step = one pass of the hot path over one batch of synthetic frames (default: the
64-frame mixed batch of BASELINE config 5 per GPU, 2.1 GB, larger than 2x L2 so every pass
streams from HBM).  Prints ONE JSON line (rank 0).

  value     whole-job frames/s with the batch already resident in HBM, timed with CUDA events
            on the launch stream, max over ranks
  e2e       frames/s through the host-buffer C-ABI (scope_submit_host / scope_wait_host ring):
            pinned host frames in, host results out, H2D and D2H copies inside the timed region
  roofline  the accumulation kernel alone: algorithmic bytes (W*H*4 per frame) / its CUDA-event
            duration (scope_profile_*), against MEASURED_PEAKS.json's hbm_gbs
  cpu_baseline  the reference's own loops (oracle/_ref) on the box's host cores, bounded sample

Under torchrun (N > 1) every rank owns one GPU and a disjoint share of the frames
(frame-sharded, no data-path collective): weak scaling.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WIDTH_4K, HEIGHT_4K = 3840, 2160


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames-per-gpu", type=int, default=64)
    ap.add_argument("--width", type=int, default=WIDTH_4K)
    ap.add_argument("--height", type=int, default=HEIGHT_4K)
    ap.add_argument("--content", default="mixed", choices=["mixed", "random", "natural", "ramp", "solid", "ui"])
    ap.add_argument("--colorspace", type=int, default=2)
    ap.add_argument("--e2e-frames", type=int, default=16, help="frames per step on the host-buffer path")
    ap.add_argument("--cpu-sample-frames", type=int, default=0, help="0 = one frame per host thread (bounded)")
    ap.add_argument("--workload", default="batch", choices=["batch", "roi-tiled-8k", "stream-vscope-4k"],
                    help="batch = the headline (BASELINE config 5 / metric); roi-tiled-8k = config 4; "
                         "stream-vscope-4k = config 3")
    ap.add_argument("--bands", default="rows", choices=["rows", "cols"])
    ap.add_argument("--in-flight", type=int, default=2, help="roi-tiled-8k: frames in flight (streams, accumulator sets)")
    ap.add_argument("--emulate-world", type=int, default=0,
                    help="roi-tiled-8k on ONE GPU: do rank 0's share of an M-rank run (a rank's pipeline without M GPUs)")
    ap.add_argument("--graph", action="store_true",
                    help="roi-tiled-8k: replay each frame's launches (reset, accumulate, cross-rank step) as one CUDA graph, "
                         "so that the number is the device's and not the Python harness's (not with --reduce nccl at N > 1)")
    ap.add_argument("--reduce", default="nccl", choices=["nccl", "peers", "peers-one-shot", "nvls", "nvls-one-shot"],
                    help="roi-tiled-8k: NCCL all-reduce + clamp, or the fused peer-memory kernel (scope_finalize_peers)")
    ap.add_argument("--scopes", default="hist,wave,vscope",
                    help="subset of hist,wave,vscope for the batch workload (default: all three = the headline)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-config4", action="store_true", help="skip the config4 sub-record of the default line")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# --------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe) during the timed region
# --------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def wait_first(self, timeout: float = 4.0):
        """block until nvidia-smi has delivered its first sample (its start-up can take longer than a whole timed
        region), then forget what was sampled so far: everything kept from here on is taken while the caller keeps
        the GPU busy"""
        t0 = time.time()
        while self.proc and not self.lines and time.time() - t0 < timeout:
            time.sleep(0.01)
        self.mark = len(self.lines)

    def n_samples(self):
        return len(self.lines) - getattr(self, "mark", 0)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[getattr(self, "mark", 0):]:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax = float(parts[2])
            except ValueError:
                continue
            for n, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------
# the reference's CPU path (oracle/_ref when present, else the oracle port)
# --------------------------------------------------------------------------------------
class CpuReference:
    """The reference's three draw loops (hist RGB + waveform RGB + vectorscope) on host threads,
    one frame per thread at a time: oracle/_ref (the reference's own compiled loops) when it
    exists, else the oracle port.  The YUV plane (made by the GPU shader in the reference) is
    prepared once, outside every timed region."""

    def __init__(self, width, height, n_frames, threads, colorspace, content):
        import obs_color_monitor_b200 as pkg
        from oracle.oracle import Oracle, Ref

        self.orc = Oracle()
        self.kind = "reference" if Ref.available() else "port"
        self.impl = Ref() if self.kind == "reference" else None
        self.threads, self.n = threads, n_frames
        offset = {"mixed": None, "random": 0, "ramp": 1, "solid": 2, "natural": 3, "ui": None}[content]
        self.frames = []
        for i in range(n_frames):
            if content == "ui":
                f = pkg.frames.ui(width, height, i)
            else:
                f = pkg.frames.mixed(width, height, i if offset is None else 4 * i + offset)
            self.frames.append((f, self.orc.rgb_to_yuv(f, colorspace)))

    def _work(self, i):
        f, yuv = self.frames[i % self.n]
        if self.impl is not None:
            self.impl.histogram(0x07, f, yuv)
            self.impl.waveform(0x07, f, yuv)
            self.impl.vectorscope(yuv)
        else:
            self.orc.histogram_counts(0x07, f, yuv)
            self.orc.waveform(0x07, f, yuv)
            self.orc.vectorscope(yuv)

    def one_thread(self, budget_s=3.0):
        """What the reference does in OBS: ONE "color-monitor" worker runs the three loops back to back
        (roi.c:329-341).  Frames/s of that, and - separately, because the reference gets the YUV plane from
        its GPU shader - of the oracle's CPU transform.  A few seconds, median of the passes."""
        def median_fps(fn):
            times, spent = [], 0.0
            while spent < budget_s / 2 and len(times) < 20:
                t0 = time.perf_counter()
                fn()
                times.append(time.perf_counter() - t0)
                spent += times[-1]
            return 1.0 / sorted(times)[len(times) // 2], len(times)
        try:
            loops, n1 = median_fps(lambda: self._work(0))
            f = self.frames[0][0]
            yuv, n2 = median_fps(lambda: self.orc.rgb_to_yuv(f, 2))
            return {"loops_frames_per_s": round(loops, 2), "loops_passes": n1, "yuv_transform_frames_per_s": round(yuv, 2),
                    "yuv_passes": n2, "note": "one thread, one frame at a time, median; the transform is the oracle's C code "
                                              "(a GPU shader in the reference) and is not part of any other number here"}
        except Exception as e:   # context only: never a reason to lose the bench line
            return {"error": repr(e)[:200]}

    def step(self):
        """one pass over the sample; returns seconds"""
        from concurrent.futures import ThreadPoolExecutor

        with ThreadPoolExecutor(max_workers=self.threads) as ex:
            t0 = time.perf_counter()
            list(ex.map(self._work, range(self.n)))
            return time.perf_counter() - t0


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = args.cpu_sample_frames or min(threads, 32)
    ref = CpuReference(args.width, args.height, n, threads, args.colorspace, args.content)
    # one step = `reps` passes over the n sample frames, sized from a first untimed pass so that a
    # step takes about 2 s and the whole run (warm-up + steps) stays inside the budget
    budget_s = 150.0
    dt0 = ref.step()
    total_steps = max(args.warmup, 1) + max(args.steps, 1)
    reps = max(1, min(int(2.0 / max(dt0, 1e-3)), int(budget_s / (total_steps * max(dt0, 1e-3)))))
    spent, times, warm = dt0, [], 0
    for _ in range(args.warmup):
        spent += sum(ref.step() for _ in range(reps))
        warm += 1
        if spent > budget_s / 3:
            break
    for _ in range(args.steps):
        dt = sum(ref.step() for _ in range(reps))
        times.append(dt)
        spent += dt
        if spent > budget_s:
            break
    value = n * reps * len(times) / sum(times)
    sample = (f"{reps} passes over {n} {args.content} {args.width}x{args.height} frames per step ({n * reps} frames), "
              f"the reference's hist RGB + waveform RGB + vectorscope loops, one frame per thread over {threads} threads, "
              f"YUV plane precomputed")
    line = {
        "impl": "reference", "metric": metric_name(args), "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": warm, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": bench_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": ref.kind, "sample": sample,
                         "one_thread": ref.one_thread()},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bench_config(args, world):
    """the `config` object of BOTH arms (same keys, same values: it names the workload, not the arm; what the
    reference arm samples of it is said in its cpu_baseline.sample)"""
    n, W, H = args.frames_per_gpu, args.width, args.height
    return {"workload": workload_name(args), "frames_per_gpu": n, "global_batch": n * world, "width": W, "height": H,
            "parallelism": f"frame-sharded x{world}",
            "l2": "inputs larger than L2 (batch = %.2f GB per GPU per step)" % (n * W * H * 4 / 1e9)}


def check_parity(args, pkg, batch, out, st, n_check=4):
    """frames 0..n_check-1 of the timed batch: GPU outputs vs the reference's own loops (bit-exact or counted)"""
    import numpy as np
    from oracle.oracle import Oracle, Ref

    try:
        orc = Oracle()
        ref = Ref() if Ref.available() else None
        host = batch[:n_check].cpu().numpy()
        bad, what = 0, []
        for i in range(n_check):
            f = np.ascontiguousarray(host[i])
            yuv = orc.rgb_to_yuv(f, args.colorspace)
            exp = {}
            if "hist" in out:
                exp["hist"] = orc.histogram_counts(0x07, f, yuv).ravel()
            if "wave" in out:
                exp["wave"] = (ref.waveform(0x07, f, yuv) if ref else orc.waveform(0x07, f, yuv))
            if "vscope" in out:
                exp["vscope"] = (ref.vectorscope(yuv) if ref else orc.vectorscope(yuv))
            for k, e in exp.items():
                got = out[k][i].cpu().numpy()
                got = got.view(np.uint32).ravel() if k == "hist" else got
                if not np.array_equal(got, np.asarray(e).reshape(got.shape)):
                    bad += 1
                    what.append(f"frame {i}: {k}")
        return {"checked_frames": n_check, "mismatches": bad, "against": "oracle/_ref (the reference's compiled loops)"
                if ref else "oracle port", "scopes": sorted(k for k in ("hist", "wave", "vscope") if k in out),
                "detail": what[:8]}
    except Exception as e:   # context: never a reason to lose the bench line
        return {"checked_frames": 0, "mismatches": None, "error": repr(e)[:200]}


def run_config4_record(args, eng, pkg, dev, world, rank):
    """the `config4` sub-record of the default line: 8K luma waveform over the ranks, CUDA-graphed per frame.
    At N > 1: column bands with the strip kernel's own peer stores (the best form) and row bands with the fused
    peer-memory reduce; if symmetric memory is not available the NCCL forms."""
    import torch.distributed as dist

    rec = {}
    plans = [("cols", "peers"), ("rows", "peers")] if world > 1 else [("rows", "peers")]
    for bands, red in plans:
        for attempt in (red, "nccl"):
            try:
                # column bands at N ranks: 240 / N strips = CTAs per frame and rank; eight frames in flight fill the GPU
                r = measure_config4(eng, pkg, dev, world, rank, bands=bands, reduce=attempt, graph=True,
                                    steps=400 if world > 1 else 200, warmup=24, colorspace=args.colorspace,
                                    in_flight=(8 if bands == "cols" else 4) if world > 1 else 2)
                err = None
            except Exception as e:
                r, err = None, repr(e)[:300]
            ok = 1 if r is not None else 0
            if world > 1:
                import torch
                flag = torch.tensor([ok], dtype=torch.int32, device=dev)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                ok = int(flag.item())
            if ok:
                peak = _hbm_peak()[0]
                r["roofline"].update({"peak": peak, "frac": r["roofline"]["achieved"] / peak})
                rec[bands] = r
                break
            rec[bands + "_error_" + attempt] = err
            if attempt == "nccl" or world == 1:
                break
    best = max((v for k, v in rec.items() if isinstance(v, dict) and "value" in v), key=lambda v: v["value"], default=None)
    if best is not None:
        rec["best"] = {"bands": best["bands"], "reduce": best["reduce"], "value": best["value"],
                       "ms_per_frame": best["ms_per_frame"], "frames_in_flight": best["frames_in_flight"]}
    return rec


def metric_name(args):
    return f"frames/sec fused scopes @{args.width}x{args.height} BGRA"


def workload_name(args):
    return (f"batch of {args.frames_per_gpu} independent {args.width}x{args.height} BGRA frames per GPU "
            f"({'mixed: random/ramp/solid/natural cycle' if args.content == 'mixed' else args.content + ' content'}), "
            f"fused histogram RGB + waveform RGB + "
            f"vectorscope BT.{'601' if args.colorspace == 1 else '709'}, frame-sharded (BASELINE config 5; "
            f"metric quoted @3840x2160)")


# --------------------------------------------------------------------------------------
# this repo's arm
# --------------------------------------------------------------------------------------
def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    import obs_color_monitor_b200 as pkg
    from obs_color_monitor_b200 import frames_torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the scope kernels have no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W, H, n = args.width, args.height, args.frames_per_gpu

    eng = pkg.ScopeEngine(local_rank)
    scope_bits = sum({"hist": pkg.SCOPE_HIST, "wave": pkg.SCOPE_WAVE, "vscope": pkg.SCOPE_VSCOPE}[x]
                     for x in args.scopes.split(",") if x)
    st = pkg.ScopeSettings(colorspace=args.colorspace, scopes=scope_bits)
    batch = frames_torch.mixed_batch(n, W, H, dev, first_index=rank * n, content=args.content)
    out = eng.alloc_device_out(n, W, st, dev)
    torch.cuda.synchronize()

    def step():
        eng.accumulate_device(batch, settings=st, out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    # sanity: every pixel of every frame was counted (alpha is 255 everywhere in the synthetic batch)
    diagnostic = bool(os.environ.get("SCOPE_BENCH_DIAGNOSTIC"))   # kernel A/B builds that compute nothing (ring-only timing)
    if "hist" in out and not diagnostic:
        hsum = out["hist"].to(torch.int64).sum(dim=1)
        assert bool((hsum == 3 * W * H).all()), "histogram totals are wrong"
    if "vscope" in out and not diagnostic:
        assert bool((out["vscope"].amax(dim=(1, 2)) > 0).all())
    # parity of the TIMED batch (outside the timed region): frames 0..3 = one of every content class of the mixed
    # batch, bit for bit against the reference's own compiled loops (oracle/_ref; the oracle port if absent)
    parity = check_parity(args, pkg, batch, out, st, n_check=min(4, n)) if rank == 0 else None

    # clocks DURING load: the sampler starts, then the same steps run untimed until it has seen the GPU busy for a
    # while (>= 0.3 s and >= 5 samples on rank 0; the other ranks keep their GPUs equally busy for the same time), and
    # the timed steps follow without a gap
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    t_load = time.time()
    while True:
        for _ in range(8):
            step()
        torch.cuda.synchronize()
        dt_load = time.time() - t_load
        if dt_load >= 0.3 and (rank != 0 or sampler.n_samples() >= 5 or dt_load >= 3.0):
            break
    eng.ctx.profile_enable(True)
    eng.ctx.profile_read()
    launches0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count - launches0
    kernel_ms = eng.ctx.profile_read()
    eng.ctx.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    total_frames = n * world * args.steps
    value = total_frames / (elapsed_ms * 1e-3)

    # ---- roofline of the accumulation kernel (rank-local, then max duration over ranks) ----
    k_ms = sum(kernel_ms) / max(len(kernel_ms), 1)
    kt = torch.tensor([k_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(kt, op=dist.ReduceOp.MAX)
    k_ms = float(kt.item())
    launches_per_step = max(len(kernel_ms) // max(args.steps, 1), 1)
    alg_bytes_per_launch = n * W * H * 4 / launches_per_step
    peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, peak_src = float(mp["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy, read+write)"
    except Exception:
        pass
    achieved = alg_bytes_per_launch / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None,
                "kernel": ("scope_fused_kernel_v3 (hist RGB + waveform RGB + vectorscope in one pass)"
                           if sorted(args.scopes.split(",")) == ["hist", "vscope", "wave"]
                           else "scope_strip_kernel_tma, the general strip kernel (scopes: %s)" % args.scopes),
                "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg_bytes_per_launch, "peak_source": peak_src,
                "kernel_share_of_step": (k_ms * launches_per_step) / (elapsed_ms / args.steps)}
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic_latest.json")))
        roofline["traffic"] = tr["dram_bytes_per_frame"] * n / launches_per_step
        roofline["traffic_note"] = tr.get("note")
    except Exception:
        pass

    # ---- e2e: host buffers through the C-ABI ring, copies inside the timed region ----
    e2e = None
    if not args.no_e2e and scope_bits == pkg.SCOPE_ALL:
        e2e = run_e2e(args, eng, st, batch, dev, world, rank)

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        ns = args.cpu_sample_frames or min(threads, 32)
        ref = CpuReference(W, H, ns, threads, args.colorspace, args.content)
        ref.step()
        # bounded sample: passes over the same ns frames until about 10 s of wall clock are spent (4 s at N > 1,
        # where the other ranks wait for rank 0)
        spent, passes = 0.0, 0
        while spent < (10.0 if world == 1 else 4.0) and passes < 1000:
            spent += ref.step()
            passes += 1
        cpu = {"value": ns * passes / spent, "unit": "frames/s", "cores": threads, "kind": ref.kind,
               "sample": f"{passes} passes over {ns} {args.content} {W}x{H} frames ({ns * passes} frames, {spent:.1f} s), "
                         f"the reference's hist RGB + waveform RGB + vectorscope loops, one frame per thread over "
                         f"{threads} threads, YUV plane precomputed",
               "one_thread": ref.one_thread()}

    # ---- BASELINE config 4 on the same ranks (one 8K frame, luma waveform, bands over the GPUs) ----
    config4 = None
    if not args.no_config4 and args.workload == "batch":
        del batch, out
        torch.cuda.empty_cache()
        config4 = run_config4_record(args, eng, pkg, dev, world, rank)

    if rank == 0:
        line = {
            "metric": metric_name(args), "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": bench_config(args, world),
            "clocks": clocks, "gpu_launches": launches, "roofline": roofline, "parity": parity, "cpu_baseline": cpu,
            "e2e": e2e, "config4": config4,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    eng.close()


def bind_to_gpu_numa_node(dev_index):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (sysfs), so that the page-locked frame buffers
    allocated next come from that node's memory (first touch) and the copies do not cross the socket interconnect.
    Returns what it did; a box with one node (or no sysfs entry) is left alone."""
    import torch

    info = {"gpu_node": None, "bound_cpus": None}
    try:
        pr = torch.cuda.get_device_properties(dev_index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        info["gpu_node"] = node
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        info["nodes"] = len(nodes)
        if node >= 0 and len(nodes) > 1:
            cpus = set()
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
            allowed = cpus & os.sched_getaffinity(0)
            if allowed:
                os.sched_setaffinity(0, allowed)
                info["bound_cpus"] = len(allowed)
    except Exception as e:   # context only
        info["error"] = repr(e)[:120]
    return info


def run_e2e_shim(args, eng, st, host_ptr, nf, fbytes, dev, world):
    """The same frames through the C shim a maintainer links (include/cm_shim.h): b200_cm_tick ->
    b200_cm_render_target (zero-copy staging of the page-locked frame into queue slot = ring slot) -> the worker
    thread -> b200_roi_surface_cb (one fused pass: submit frame n, file the results of frame n - 1 into the scopes'
    double buffers).  A render that finds the worker busy is DROPPED by the capture core (common.c:260-268); the
    driver below offers the frame again, so that every frame of the step is processed and the rate is comparable."""
    import ctypes as C

    import torch
    import torch.distributed as dist
    from obs_color_monitor_b200 import shim as S

    W, H = args.width, args.height
    lib, ctx = S.load(), eng.ctx.handle
    cm, roi = S.CmSource(), S.RoiSource()
    his, wvs, vss = S.HisSource(), S.WvsSource(), S.VssSource()
    lib.b200_cm_create(C.byref(cm))
    lib.b200_cm_attach_gpu(C.byref(cm), ctx, True)
    cm.colorspace = st.colorspace
    lib.b200_roi_init(C.byref(roi), ctx, 0)                 # SCOPE_MODE_FUSED
    lib.b200_his_init(C.byref(his), ctx, 0x07)
    lib.b200_wvs_init(C.byref(wvs), ctx, 0x07)
    lib.b200_vss_init(C.byref(vss), ctx)
    lib.b200_roi_register_his(C.byref(roi), C.byref(his))
    lib.b200_roi_register_wvs(C.byref(roi), C.byref(wvs))
    lib.b200_roi_register_vss(C.byref(roi), C.byref(vss))
    cm.flags = lib.b200_roi_capture_flags(C.byref(roi)) & ~S.CM_FLAG_ROI
    lib.b200_cm_request(C.byref(cm), C.cast(lib.b200_roi_surface_cb, C.c_void_p), C.cast(C.byref(roi), C.c_void_p))
    retries = 0

    def step():
        nonlocal retries
        for i in range(nf):
            while True:
                lib.b200_cm_tick(C.byref(cm))
                if lib.b200_cm_render_target(C.byref(cm), host_ptr + i * fbytes, None, W * 4, W, H):
                    break
                retries += 1
                time.sleep(0)
        lib.b200_cm_drain(C.byref(cm))

    def finish():
        lib.b200_cm_tick(C.byref(cm))                       # the queue hands a surface to the worker one render late
        while not lib.b200_cm_render_target(C.byref(cm), host_ptr, None, W * 4, W, H):
            lib.b200_cm_tick(C.byref(cm))
        lib.b200_cm_drain(C.byref(cm))
        lib.b200_roi_finish(C.byref(roi))

    for _ in range(2):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    filed0, retries = roi.frames_filed, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    finish()
    dt = time.perf_counter() - t0
    filed = roi.frames_filed - filed0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    import numpy as np
    r = his.w_tex_buf ^ 1
    hist = np.frombuffer((C.c_float * 1024).from_address(his.tex_buf[r]), dtype=np.float32)
    ok = int(hist.sum()) == 3 * W * H
    lib.b200_cm_destroy(C.byref(cm))
    lib.b200_roi_destroy(C.byref(roi))
    for d, src in ((lib.b200_his_destroy, his), (lib.b200_wvs_destroy, wvs), (lib.b200_vss_destroy, vss)):
        d(C.byref(src))
    return {"value": filed * world / dt, "unit": "frames/s", "frames_filed": int(filed), "renders_dropped_and_offered_again": retries,
            "histogram_total_ok": ok,
            "path": "b200_cm_tick -> b200_cm_render_target (zero-copy into queue slot = ring slot) -> worker thread -> "
                    "b200_roi_surface_cb (submit frame n, file frame n-1 into the scopes' double buffers); wall clock"}


def run_e2e(args, eng, st, batch, dev, world, rank):
    """Pinned host frames -> scope_submit_host (H2D + kernels + D2H on the slot's stream) ->
    scope_wait_host, three slots in flight like the reference's 3-deep stagesurface ring."""
    import ctypes as C

    import numpy as np
    import torch
    import torch.distributed as dist

    import obs_color_monitor_b200 as pkg

    W, H = args.width, args.height
    nf = min(args.e2e_frames, batch.shape[0])
    fbytes = W * H * 4
    lib = eng.lib
    numa = bind_to_gpu_numa_node(dev.index if dev.index is not None else 0)
    ptr = lib.scope_host_alloc(nf * fbytes)
    if not ptr:
        return None
    host = np.ctypeslib.as_array((C.c_uint8 * (nf * fbytes)).from_address(ptr)).reshape(nf, H, W, 4)
    host[:] = batch[:nf].cpu().numpy()
    wave_bytes = 256 * W * 4
    res_ptr = lib.scope_host_alloc(3 * (4096 + 12 + 65536 + wave_bytes))
    slots = pkg._ffi.RING_SLOTS
    outs = []
    for s in range(slots):
        base = res_ptr + s * (4096 + 12 + 65536 + wave_bytes)
        o = pkg._ffi.OutHost()
        o.hist_counts, o.hist_max, o.vscope, o.wave = base, base + 4096, base + 4096 + 12, base + 4096 + 12 + 65536
        outs.append(o)
    p = st.to_c()

    def surface(i):
        s = pkg._ffi.Surface()
        s.rgb_data = ptr + i * fbytes
        s.yuv_data = None
        s.linesize, s.width, s.height, s.colorspace = W * 4, W, H, st.colorspace
        return s

    surfs = [surface(i) for i in range(nf)]

    def step():
        for i in range(nf):
            sl = i % slots
            if i >= slots:
                eng.ctx.check(lib.scope_wait_host(eng.ctx.handle, sl, C.byref(outs[sl])))
            eng.ctx.check(lib.scope_submit_host(eng.ctx.handle, sl, C.byref(p), C.byref(surfs[i])))
        for i in range(max(nf - slots, 0), nf):
            eng.ctx.check(lib.scope_wait_host(eng.ctx.handle, i % slots, C.byref(outs[i % slots])))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(2):
        step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    hist = np.ctypeslib.as_array((C.c_uint32 * 1024).from_address(res_ptr))
    assert int(hist.sum()) == 3 * W * H, "e2e histogram total is wrong"
    shim_leg = None
    try:
        shim_leg = run_e2e_shim(args, eng, st, ptr, nf, fbytes, dev, world)
    except Exception as e:   # context: never a reason to lose the bench line
        shim_leg = {"error": repr(e)[:200]}
    lib.scope_host_free(ptr)
    lib.scope_host_free(res_ptr)
    out = {"value": nf * world * args.steps / dt, "unit": "frames/s", "h2d_bytes_per_step": nf * fbytes,
           "d2h_bytes_per_step": nf * (4096 + 16 + 65536 + wave_bytes), "frames_per_step": nf,
           "path": "scope_submit_host/scope_wait_host, 3-slot ring, pinned host frames, wall clock max over ranks",
           "shim": shim_leg, "numa": numa}
    # what bounds this path: the host->device copy of the frames.  Measure the box's pinned H2D bandwidth
    # (plain copy, nothing else running) so that the e2e number can be read against ITS roofline.
    try:
        n = 256 << 20
        src = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        dst = torch.empty(n, dtype=torch.uint8, device=dev)
        best = 0.0
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            dst.copy_(src, non_blocking=True)
            b.record()
            b.synchronize()
            best = max(best, n / (a.elapsed_time(b) * 1e-3) / 1e9)
        per_gpu = out["value"] / world * fbytes / 1e9
        # the same copy on ALL ranks at once (barrier, then 5 back-to-back copies each): what the host side of the
        # box gives every GPU when N of them pull frames at the same time - the ceiling the N-GPU e2e number has
        conc = best
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                dst.copy_(src, non_blocking=True)
            b.record()
            b.synchronize()
            tc = torch.tensor([5 * n / (a.elapsed_time(b) * 1e-3) / 1e9], dtype=torch.float64, device=dev)
            dist.all_reduce(tc, op=dist.ReduceOp.MIN)
            conc = float(tc.item())
        out["pcie"] = {"h2d_gbs_measured": round(best, 2), "h2d_gbs_concurrent": round(conc, 2),
                       "h2d_gbs_used": round(per_gpu, 2), "frac": round(per_gpu / conc, 4),
                       "frac_of_alone": round(per_gpu / best, 4),
                       "how": "256 MiB pinned -> device copy_, CUDA events, per GPU: `measured` = best of 5 with this "
                              "rank copying alone-ish (ranks are not synchronised), `concurrent` = all ranks copying at "
                              "the same time (min over ranks); frac = used / concurrent"}
    except Exception as e:  # the measurement is context, never a reason to lose the bench line
        out["pcie"] = {"error": repr(e)[:200]}
    return out


def measure_config4(eng, pkg, dev, world, rank, bands="rows", reduce="peers", graph=True, steps=200, warmup=20,
                    colorspace=2, check=True, in_flight=2, emulate_world=0):
    """BASELINE config 4: ONE 7680x4320 frame, waveform of luma (components 0x20), split into row bands (or
    column bands) over the ranks.  A step = one frame; two frames are in flight on two streams (double-buffered
    accumulators and images), so the cross-rank step of frame i overlaps the accumulation of frame i + 1.

      rows  every rank's strip kernel STORES u16-pair partials for its band (scope_accumulate_band, exclusive: no
            zero-fill, no atomics), then  nccl: all-reduce + clamp  |  peers / nvls: barrier -> ONE kernel that sums
            the ranks' partials over NVLink (or in the switch), saturates and stores the u8 image into every rank
            -> barrier
      cols  a rank's columns are final: its strip kernel stores them as u8 straight into every rank's image (peer
            stores; nccl: all-gather of the u8 bands) -> one barrier.  No reduce step at all.
    in_flight = F: F frames are in flight on F streams (F sets of accumulators and images).  A rank's band of a
    frame is at most 240 / N strips = CTAs of the strip kernel, far fewer than the GPU holds at N = 4: only several
    frames in flight fill it, which is what a stream of video frames offers anyway (the per-frame latency is reported
    next to the rate).
    emulate_world = M (one process, one GPU): this GPU does rank 0's share of an M-rank run - the same band kernel,
    the same number of launches, no peers - to see a rank's pipeline without paying for M GPUs.
    Returns a dict (every rank computes it; rank 0 reports)."""
    import torch
    import torch.distributed as dist

    from obs_color_monitor_b200 import frames_torch

    W, H = 7680, 4320
    F = max(1, int(in_flight))
    emulate = int(emulate_world) if world == 1 and emulate_world and emulate_world > 1 else 0
    st = pkg.ScopeSettings(scopes=pkg.SCOPE_WAVE, wave_components=pkg.COMP_Y, colorspace=colorspace)

    def new_tiled():
        if reduce == "nccl":
            return pkg.sharding.TiledFrame(eng, W, H, st, mode=bands)
        tf = pkg.sharding.PeerTiledFrame(eng, W, H, st, mode=bands, two_shot=reduce in ("peers", "nvls"),
                                         nvls=reduce.startswith("nvls"))
        if emulate:
            tf.bands = (pkg.sharding.row_bands(H, emulate) if bands == "rows" else pkg.sharding.col_bands(W, emulate))[:1]
        return tf

    ring = [new_tiled() for _ in range(F)]
    a, b = ring[0].my_band
    # 4 different frames so that successive steps do not hit L2 (an 8K frame is 133 MB > L2; a rank's band of 4
    # frames is 133 MB at N = 4)
    if bands == "rows":
        data = [frames_torch.mixed_batch(1, W, b - a, dev, first_index=4 * i + 3, content="natural")[0] for i in range(4)]
        width = None
    else:
        full = [frames_torch.mixed_batch(1, W, H, dev, first_index=4 * i + 3, content="natural")[0] for i in range(4)]
        data = [torch.as_strided(f.reshape(-1)[a * 4:], (H, (W - a) * 4), (W * 4, 1)) for f in full]
        width = b - a
    streams = [torch.cuda.Stream(dev) for _ in range(F)]
    check = check and not emulate

    def enqueue(j, k):
        tf = ring[j]
        tf.reset()
        tf.accumulate(data[k], width=width)
        tf.start_reduce()
        return tf.finish()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    outs = {}
    n_warm = (max(warmup, 4) + F - 1) // F * F
    for i in range(n_warm):
        with torch.cuda.stream(streams[i % F]):
            outs[i % F] = enqueue(i % F, i % 4)
    barrier()
    parity = None
    if check:
        # the last two frames against the unsharded pass on this GPU (which tests/ pins on the oracle at 8K)
        bad = 0
        for i in (n_warm - 2, n_warm - 1) if F > 1 else (n_warm - 1,):
            j, k = i % F, i % 4
            if bands == "rows":
                whole_src = frames_torch.mixed_batch(1, W, H, dev, first_index=4 * k + 3, content="natural") \
                    if world == 1 else None
            else:
                whole_src = full[k][None]
            if whole_src is not None:
                ref = eng.accumulate_device(whole_src, settings=st)
                torch.cuda.synchronize()
                bad += int(not torch.equal(outs[j]["wave"][0], ref["wave"][0]))
        parity = {"checked_frames": min(F, 2) if (bands == "cols" or world == 1) else 0, "mismatches": bad}
    launches_per_frame, big = None, None
    use_graph = graph and not (reduce == "nccl" and world > 1)
    graph_error = None
    G = max(8, 2 * F) // F * F   # frames per graph replay: all streams inside ONE graph (fork / join), so that the host's
             # graph-launch rate (a few tens of microseconds per replay from Python) cannot be what a 20-microsecond
             # frame waits for
    if use_graph:
        try:
            big = torch.cuda.CUDAGraph()
            l_before = eng.launch_count
            fork, graph_events = torch.cuda.Event(), []
            with torch.cuda.graph(big, stream=streams[0]):
                fork.record(streams[0])
                for s_ in streams[1:]:
                    s_.wait_event(fork)
                for i in range(G):
                    with torch.cuda.stream(streams[i % F]):
                        enqueue(i % F, i % 4)
                for s_ in streams[1:]:
                    graph_events.append(torch.cuda.Event())
                    graph_events[-1].record(s_)
                    streams[0].wait_event(graph_events[-1])
            launches_per_frame = (eng.launch_count - l_before) / G
            for _ in range(2):
                with torch.cuda.stream(streams[0]):
                    big.replay()
            barrier()
        except Exception as e:      # a capture that fails on one rank must not hang the others: fall back together
            graph_error = repr(e)[:200]
            use_graph = False
        if world > 1:
            flag = torch.tensor([0 if use_graph else 1], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            use_graph = use_graph and int(flag.item()) == 0
    if use_graph:
        steps = max(G, steps // G * G)

    def step(i):
        if use_graph:
            if i % G == 0:
                with torch.cuda.stream(streams[0]):
                    big.replay()
            return
        with torch.cuda.stream(streams[i % F]):
            enqueue(i % F, i % 4)

    # the accumulation kernel alone (this rank's band, stream 0, events around single launches): what the frame time
    # is made of besides it is the cross-rank step and whatever the two streams do not overlap
    eng.ctx.profile_enable(True)
    eng.ctx.profile_read()
    with torch.cuda.stream(streams[0]):
        for i in range(6):
            ring[0].accumulate(data[i % 4], width=width)
    barrier()
    kms = eng.ctx.profile_read()
    eng.ctx.profile_enable(False)
    kernel_us = 1e3 * sorted(kms)[len(kms) // 2] if kms else None

    # clocks DURING load: every frame has a cross-rank barrier inside, so all ranks run the SAME number of untimed
    # frames under the sampler (about 0.4 s worth, from rank 0's own estimate), then the timed ones without a gap
    sampler = ClockSampler(dev.index if dev.index is not None else 0)
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    t_est = time.time()
    for i in range(G if use_graph else 8):
        step(i)
    barrier()
    per_frame = max((time.time() - t_est) / (G if use_graph else 8), 5e-6)
    n_load = torch.tensor([min(int(0.4 / per_frame), 40000)], dtype=torch.int64, device=dev)
    if world > 1:
        dist.broadcast(n_load, src=0)
    n_load = max(int(n_load.item()) // G * G, G)
    for i in range(n_load):
        step(i)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    join = torch.cuda.Event()
    l0 = eng.launch_count
    barrier()
    cur = torch.cuda.current_stream(dev)
    ev0.record(cur)
    for s_ in streams:
        s_.wait_event(ev0)
    for i in range(steps):
        step(i)
    for s_ in streams[1:]:
        join = torch.cuda.Event()
        join.record(s_)
        streams[0].wait_event(join)
    ev1.record(streams[0])
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    n_bins = 256 * W
    if world == 1:
        nvlink = 0
    elif bands == "cols":
        nvlink = (world - 1) * n_bins * 4 // world                      # my u8 columns into every other image
    elif reduce == "nccl":
        nvlink = 2 * (world - 1) * n_bins * 4 // world                  # ring all-reduce of the u16-pair plane
    elif reduce in ("peers", "nvls"):
        nvlink = 2 * (world - 1) * n_bins * 4 // world                  # read my slice of every peer, store it to every peer
    else:
        nvlink = (world - 1) * n_bins * 4                               # one-shot: read everything from every peer
    return {
        "metric": "frames/sec ROI-tiled waveform (luma) @7680x4320 BGRA", "value": steps / (ms * 1e-3), "unit": "frames/s",
        "n_gpus": world, "steps": steps, "warmup": max(warmup, 4), "ms_per_frame": ms / steps, "bands": bands,
        "reduce": reduce if world > 1 else "none (one rank)", "graph": bool(use_graph), "graph_error": graph_error,
        "frames_in_flight": F, "frames_per_graph": G if use_graph else None, "emulated_world": emulate or None,
        "frame_latency_us": 1e3 * F * ms / steps,   # Little's law: F frames in flight at this rate
        "band_kernel_us": kernel_us,
        "band_bytes": W * H * 4 // (emulate or world),
        "bytes_over_nvlink_per_rank_per_frame": int(nvlink),
        "gpu_launches_per_frame": launches_per_frame if launches_per_frame is not None else (eng.launch_count - l0) / steps,
        "achieved_read_GBps_all_ranks": steps * W * H * 4 / (ms * 1e-3) / 1e9, "clocks": clocks, "parity": parity,
        "roofline": {"bound": "hbm", "unit": "GB/s", "note": "per rank: its band's pixel bytes / the frame time",
                     "achieved": steps * W * H * 4 / (emulate or world) / (ms * 1e-3) / 1e9},
    }


def run_roi_tiled(args):
    """`--workload roi-tiled-8k`: BASELINE config 4 alone (see measure_config4)."""
    import torch
    import torch.distributed as dist

    import obs_color_monitor_b200 as pkg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = pkg.ScopeEngine(local_rank)
    res = measure_config4(eng, pkg, dev, world, rank, bands=args.bands, reduce=args.reduce, graph=args.graph,
                          steps=args.steps, warmup=args.warmup, colorspace=args.colorspace, in_flight=args.in_flight,
                          emulate_world=args.emulate_world)
    if rank == 0:
        peak = _hbm_peak()[0]
        res["roofline"].update({"peak": peak, "frac": res["roofline"]["achieved"] / peak})
        res.update({"ms_per_step": res["ms_per_frame"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                    "dtype": "u8", "data": "synthetic",
                    "config": {"workload": f"BASELINE config 4: one 7680x4320 frame, luma waveform, {args.bands} bands over "
                                           f"{world} GPU(s), cross-rank step: {res['reduce']}", "bands": args.bands,
                               "reduce": res["reduce"]},
                    "gpu_launches": res["gpu_launches_per_frame"] * args.steps})
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()
    eng.close()


def _hbm_peak():
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(mp["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy, read+write)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def run_stream_vscope(args):
    """BASELINE config 3: a 3840x2160 stream through the 3-slot host ring, vectorscope only,
    with the intensity-applied display image (intensity 25).  Reports sustained fps and the
    latency of a single synchronous frame."""
    import ctypes as C

    import numpy as np
    import torch

    import obs_color_monitor_b200 as pkg
    from obs_color_monitor_b200 import frames_torch

    torch.cuda.set_device(0)
    eng = pkg.ScopeEngine(0)
    W, H, nf = 3840, 2160, 12
    st = pkg.ScopeSettings(scopes=pkg.SCOPE_VSCOPE, vscope_intensity=25, colorspace=args.colorspace)
    lib = eng.lib
    ptr = lib.scope_host_alloc(nf * W * H * 4)
    host = np.ctypeslib.as_array((C.c_uint8 * (nf * W * H * 4)).from_address(ptr)).reshape(nf, H, W, 4)
    host[:] = frames_torch.mixed_batch(nf, W, H, torch.device("cuda", 0)).cpu().numpy()
    lat = []
    for i in range(6):
        t0 = time.perf_counter()
        res = eng.accumulate_host(host[i % nf], settings=st)
        lat.append(time.perf_counter() - t0)
    n = 60 * max(args.steps, 1)
    sampler = ClockSampler(0)
    sampler.start()
    sampler.wait_first()
    eng.ctx.profile_enable(True)
    eng.ctx.profile_read()
    t0 = time.perf_counter()
    for i in range(n):
        sl = i % 3
        if i >= 3:
            eng.wait_host(sl)
        assert eng.submit_host(sl, host[i % nf], settings=st)
    for i in range(n - 3, n):
        res = eng.wait_host(i % 3)
    dt = time.perf_counter() - t0
    kernel_ms = eng.ctx.profile_read()
    eng.ctx.profile_enable(False)
    clocks = sampler.stop()
    assert res["vscope"].max() > 0 and res["vscope_display"].max() == 255
    # parity of the last frames against the oracle (outside the timed region)
    from oracle.oracle import Oracle
    orc = Oracle()
    bad = 0
    for i in range(2):
        f = np.ascontiguousarray(host[i])
        r = eng.accumulate_host(f, settings=st)
        vs = orc.vectorscope(orc.rgb_to_yuv(f, args.colorspace))
        bad += int(not np.array_equal(r["vscope"], vs)) + int(not np.array_equal(r["vscope_display"], orc.apply_intensity(vs, 25)))
    lib.scope_host_free(ptr)
    peak, peak_src = _hbm_peak()
    k_ms = sum(kernel_ms) / max(len(kernel_ms), 1)
    achieved = W * H * 4 / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    print(json.dumps({
        "clocks": clocks, "parity": {"checked_frames": 2, "mismatches": bad},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "kernel": "the strip kernel of the vectorscope-only pass, one frame per launch", "kernel_ms": k_ms,
                     "algorithmic_bytes_per_launch": W * H * 4, "peak_source": peak_src, "traffic": None,
                     "note": "one 33 MB frame per launch (148 CTAs x ~0.8 strips): launch-latency sized; the stream "
                             "itself is bound by the host->device copy, see value"},
        "metric": "frames/sec vectorscope+intensity stream @3840x2160 BGRA (host ring)", "value": n / dt,
        "unit": "frames/s", "n_gpus": 1, "steps": n, "warmup": 6, "ms_per_step": 1e3 * dt / n,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": "BASELINE config 3: 3840x2160 stream, 3-slot ring (CM_SURFACE_QUEUE_SIZE), vectorscope "
                               "256x256 + intensity 25 display image, pinned host frames in, host results out",
                   "sync_frame_latency_ms": 1e3 * min(lat), "realtime_60fps_headroom": (n / dt) / 60.0},
        "gpu_launches": eng.launch_count}), flush=True)
    eng.close()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.workload == "roi-tiled-8k":
        run_roi_tiled(args)
    elif args.workload == "stream-vscope-4k":
        run_stream_vscope(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
