"""Turn the captures of tools/run_ncu.sh into the summaries committed under profiles/ (run HERE, no GPU: `ncu -i` only
reads the report).   usage: python tools/ncu_summary.py TAG [frames_per_launch]
   gpurun_out/prof_TAG.ncu-rep   -> profiles/ncu_full_TAG.md, profiles/ncu_lines_TAG.md, profiles/traffic_latest.json
   gpurun_out/launches_TAG.csv   -> profiles/launches_TAG.md (+ a copy of the csv)"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 64
W, H = 3840, 2160
rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True, check=True).stdout


rows = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, vals))
u = dict(zip(hdr, units))
px_warps = frames * W * H / 32
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_op_shared_atom.sum", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active"]


def num(k):
    return float(d[k].replace(",", ""))


def to_bytes(k):
    v, un = num(k), u[k].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[un]


out = [f"# ncu --set full, dominant kernel, round 2 ({tag})", "",
       f"Capture: `ncu --set full --clock-control none --import-source on -k regex:\"scope_strip|scope_fused\" -s 3 -c 1 python bench.py "
       f"--steps 1 --warmup 3 --frames-per-gpu {frames} --no-e2e --no-cpu-baseline --no-config4` (tools/run_ncu.sh; one launch = the "
       f"bench's own {frames} mixed 3840x2160 frames = {frames * W * H * 4:,} algorithmic bytes; report gpurun_out/prof_{tag}.ncu-rep, "
       f"not committed).  Numbers taken under the profiler describe the kernel; the bench values are never taken from here.", "",
       f"Kernel: `{d['Kernel Name']}`", "", "| metric | value | unit |", "|---|---|---|"]
for k in keys:
    if k in d:
        out.append(f"| `{k}` | {d[k]} | {u[k]} |")
rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
inst, cyc = num("smsp__inst_executed.sum"), num("sm__cycles_active.avg")
wf, cf = num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"), num("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")
atoms = num("smsp__inst_executed_op_shared_atom.sum")
out += ["", f"DRAM traffic per launch: read {rd / 1e6:.1f} MB + write {wr / 1e6:.1f} MB = {(rd + wr) / 1e6:.1f} MB for "
        f"{frames * W * H * 4 / 1e6:.1f} MB of algorithmic reads -> read traffic = {rd / (frames * W * H * 4):.3f} x algorithmic.", "",
        f"Per 32 pixels (one pixel-warp; {px_warps / 1e6:.2f} M of them): **{inst / px_warps:.1f} warp instructions**, "
        f"**{cyc * 148 / px_warps:.1f} SM-cycles**, {wf / px_warps:.2f} shared-memory wavefronts of which {cf / px_warps:.2f} bank-conflict "
        f"replays, {atoms / px_warps:.2f} shared atomics.", "", "Warp stall reasons (warps per issue-active cycle):", ""]
stalls = sorted(((float(d[h]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                 for h in hdr if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")), reverse=True)
out += [f"* {n}: {v:.2f}" for v, n in stalls[:12]]
open(os.path.join(ROOT, "profiles", f"ncu_full_{tag}.md"), "w").write("\n".join(out) + "\n")
json.dump({"dram_bytes_per_frame": (rd + wr) / frames, "dram_read_bytes_per_frame": rd / frames,
           "dram_write_bytes_per_frame": wr / frames,
           "note": f"ncu --set full (profiles/ncu_full_{tag}.md), {frames} frames per captured launch; bench.py scales "
                   "dram_bytes_per_frame by the frames of its own launch"},
          open(os.path.join(ROOT, "profiles", "traffic_latest.json"), "w"), indent=1)

# ---- per SASS opcode (the report carries no CUDA-line correlation on this box; the SASS view does carry counts)
rows = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--print-source", "sass"))))
h2 = rows[1]
col = {n: i for i, n in enumerate(h2)}
agg, tot_s, tot_i = {}, 0, 0
for r in rows[2:]:
    try:
        smp, ins = int(r[col["# Samples"]] or 0), int(r[col["Instructions Executed"]] or 0)
    except Exception:
        continue
    toks = r[col["Source"]].split()
    if toks and toks[0].startswith("@"):
        toks = toks[1:]
    op = toks[0].rstrip(";") if toks else "?"
    op = ".".join(op.split(".")[:2]) if op.startswith(("ATOMS", "SYNCS", "LDSM", "IMAD", "VOTE", "SHFL", "BAR", "RED", "ATOMG", "LDS", "STS")) else op.split(".")[0]
    a_ = agg.setdefault(op, [0, 0])
    a_[0] += smp
    a_[1] += ins
    tot_s += smp
    tot_i += ins
lines = [f"# Executed instructions and stall samples per SASS opcode ({tag}: `ncu -i prof_{tag}.ncu-rep --page source --csv --print-source sass`)",
         "", f"Total: {tot_s} stall samples, {tot_i / 1e6:.1f} M warp instructions = {tot_i / px_warps:.1f} per 32 pixels.", "",
         "| opcode | instructions per 32 pixels | share of instructions | share of stall samples |", "|---|---|---|---|"]
for op, (smp, ins) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    lines.append(f"| `{op}` | {ins / px_warps:.2f} | {100 * ins / max(tot_i, 1):.1f} % | {100 * smp / max(tot_s, 1):.1f} % |")
open(os.path.join(ROOT, "profiles", f"ncu_lines_{tag}.md"), "w").write("\n".join(lines) + "\n")

# ---- launch list
lcsv = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(lcsv):
    shutil.copy(lcsv, os.path.join(ROOT, "profiles", f"launches_{tag}.csv"))
    rows = [r for r in csv.reader(open(lcsv)) if len(r) > 10 and r[0].isdigit()]
    by = {}
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "")
        by.setdefault(name, []).append(float(r[-1]) / 1e3)
    total = sum(sum(v) for v in by.values())
    md = [f"# ncu launch list, round 2 ({tag}; gpu__time_duration.sum, --clock-control none)", "",
          "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -k regex:\"scope_strip|scope_fused|finalize|hist_max\" -s 9 -c 24 "
          "--csv python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-config4` (tools/run_ncu.sh)", "",
          "(64 x 3840x2160 mixed frames per step; per-launch times are serialised and cold-cache: compare SHARES)", "",
          "| kernel | launches | avg us | share of captured GPU time |", "|---|---|---|---|"]
    for name, v in sorted(by.items(), key=lambda kv: -sum(kv[1])):
        md.append(f"| `{name}` | {len(v)} | {sum(v) / len(v):.1f} | {100 * sum(v) / total:.1f} % |")
    md += ["", f"Total captured: {total / 1e3:.3f} ms over {len(rows)} launches."]
    open(os.path.join(ROOT, "profiles", f"launches_{tag}.md"), "w").write("\n".join(md) + "\n")
print("wrote profiles/ncu_full_%s.md, ncu_lines_%s.md, launches_%s.md, traffic_latest.json" % (tag, tag, tag))
