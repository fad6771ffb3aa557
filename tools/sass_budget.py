#!/usr/bin/env python
"""Static instruction budget of a strip kernel's steady-state loop, read from the SASS.  No GPU needed.

The fused pass is bound by instruction issue and by the LSU, so the numbers that
predict its speed are "warp instructions per 32 pixels" of the software-pipelined loop in tma_consume,
split by the pipe they go to, and the LSU instructions among them.  This tool reads them off a built
library, so that a source change can be judged before any GPU time is spent on it:

    python tools/sass_budget.py [lib.so] [--kernel SUBSTR] [--roles] [--dump]

How the loop and its fast path are found (nothing is hard-coded to addresses):
  * the cubin is disassembled with `nvdisasm -gi -hex`: every instruction comes with its chain of
    source lines (innermost first, then the lines it was inlined at) and its two encoding words;
  * the steady-state loop is the backward branch of the for-statement marked `// sass-loop` in
    scope_kernels.cuh (its body is two tile visits of a warp, ping-pong; one LDSM.x4 per 4 rows);
  * an instruction is COLD if any line of its chain lies between `// sass-cold{` and `// sass-cold}`
    markers in scope_kernels.cuh (the transparent-pixel branch, the flat-block branch, the overflow
    undo) or if it sits outside the loop span (spin-wait loops ptxas moved out of line);
    everything else is the fast path: what a warp executes on an interior tile of ordinary content.
Also reported: the sum of the ptxas stall counts (bits 41..44 of the second encoding word) along the
fast path = cycles ONE warp needs to issue the loop body if it never waits on a scoreboard, and the
same budget by role (innermost enclosing function) with --roles.
"""
import argparse
import collections
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_LIB = os.path.join(ROOT, "obs-color-monitor_b200", "lib", "libscope_b200.so")
DEFAULT_KERNEL = "scope_strip_kernel_tmaILi1ELb1ELb0"   # <SRC_RGB, VSCOPE, fused>
SOURCE = os.path.join(ROOT, "obs-color-monitor_b200", "csrc", "scope_kernels.cuh")

# which issue port / pipe an opcode goes to on sm_100 (B300_MICROARCH.md "Pipe rates"; IMAD* only on
# the FMA-heavy half, FFMA with an immediate operand is full rate)
PIPES = [
    ("fma-heavy (IMAD, IMAD.HI, IDP)", ("IMAD", "IDP")),
    ("fma (FFMA, FADD, FMUL)", ("FFMA", "FADD", "FMUL", "HFMA2")),
    ("alu (PRMT, LOP3, SHF, ISETP, SEL, VIMNMX, ...)",
     ("PRMT", "LOP3", "SHF", "ISETP", "SEL", "LEA", "IADD3", "VIADD", "VIMNMX", "VIMNMX3", "PLOP3", "MOV", "IABS",
      "P2R", "R2P", "FLO", "POPC", "BREV", "IADD")),
    ("uniform datapath (U*)", ("UISETP", "UMOV", "ULOP3", "UIADD3", "USEL", "ULEA", "UIMAD", "UPRMT", "USHF", "UPLOP3",
                               "R2UR", "UFLO", "UPOPC", "VOTEU", "S2UR")),
    ("lsu: shared atomics (ATOMS)", ("ATOMS",)),
    ("lsu: other (LDSM, LDS, STS, LDC, SYNCS, LDG, STG)", ("LDS", "LDSM", "STS", "SYNCS", "LDC", "LDCU", "LDG", "STG", "RED",
                                                          "REDG", "ATOMG", "ATOM", "S2R")),
    ("warp-wide (SHFL, VOTE, REDUX, MATCH)", ("SHFL", "VOTE", "REDUX", "MATCH")),
    ("control (BRA, BSSY, BSYNC, WARPSYNC, NOP)", ("BRA", "BSSY", "BSYNC", "WARPSYNC", "NOP", "YIELD", "EXIT", "CALL", "RET",
                                                   "NANOSLEEP", "BAR", "BREAK")),
]
LSU_OPS = ("ATOMS", "LDS", "LDSM", "STS", "LDG", "STG", "RED", "REDG", "ATOMG", "SYNCS")


def pipe_of(op):
    for name, ops in PIPES:
        if op in ops:
            return name
    return "other: " + op


def disassemble(lib, kernel):
    """-> (mangled name, [instruction dicts]) of the first function whose name contains `kernel`"""
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
        cubins = [f for f in os.listdir(tmp) if f.endswith(".cubin")]
        if not cubins:
            raise SystemExit(f"no cubin inside {lib}")
        text = subprocess.run(["nvdisasm", "-gi", "-hex", os.path.join(tmp, cubins[0])], capture_output=True,
                              text=True).stdout
    lines = text.splitlines()
    start = name = None
    for i, ln in enumerate(lines):
        m = re.match(r"\.text\.(\S+):", ln)
        if m:
            if start is not None:
                end = i
                break
            if kernel in m.group(1):
                start, name = i, m.group(1)
    else:
        end = len(lines)
    if start is None:
        raise SystemExit(f"no kernel matching {kernel} in {lib}")
    ins, labels, chain, fresh = [], {}, [], False
    pending_labels = []
    for ln in lines[start + 1:end]:
        m = re.match(r"\s*(\.L_x_\d+):", ln)
        if m:
            pending_labels.append(m.group(1))
            continue
        m = re.match(r'\s*//## File "(.*?)", line (\d+)', ln)
        if m:
            if not fresh:
                chain, fresh = [], True
            # only lines of scope_kernels.cuh carry markers; lines of other files (the experiments header) get 0
            chain.append(int(m.group(2)) if os.path.basename(m.group(1)) == os.path.basename(SOURCE) else 0)
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/", ln)
        if m:
            fresh = False
            addr, txt = int(m.group(1), 16), m.group(2).strip()
            for lb in pending_labels:
                labels[lb] = addr
            pending_labels = []
            pred = re.match(r"@!?U?P\d+\s+", txt)
            body = txt[pred.end():] if pred else txt
            ins.append({"addr": addr, "text": txt, "pred": bool(pred), "op": body.split()[0].split(".")[0],
                        "full_op": body.split()[0], "w0": int(m.group(3), 16), "w1": 0, "lines": list(chain)})
            continue
        m = re.match(r"\s+/\* 0x([0-9a-f]{16}) \*/", ln)
        if m and ins:
            ins[-1]["w1"] = int(m.group(1), 16)
    for i in ins:
        m = re.search(r"`\((\.L_x_\d+)\)", i["text"])
        i["target"] = labels.get(m.group(1)) if m else None
        i["stall"] = (i["w1"] >> 41) & 0xF
    return name, ins


def cold_ranges(source):
    out, begin = [], None
    for n, ln in enumerate(open(source), 1):
        if "sass-cold{" in ln:
            begin = n
        elif "sass-cold}" in ln and begin is not None:
            out.append((begin, n))
            begin = None
    return out


def function_of_line(source):
    """line -> name of the enclosing device function or named lambda (a crude scan, good enough for a table)"""
    names, cur = {}, "?"
    for n, ln in enumerate(open(source), 1):
        m = re.match(r"\s*(?:__device__|__global__|template|static|inline).*?\b(\w+)\s*\([^;]*$", ln)
        if m and "__device__" in ln or (m and "__global__" in ln):
            cur = m.group(1)
        m = re.match(r"\s*auto (\w+) = \[&\]", ln)
        if m:
            cur_lambda = m.group(1)
            names[n] = cur_lambda
            cur = cur_lambda
            continue
        m = re.match(r"\s*(\w+)\((?:const )?StripParams", ln)  # kernel definitions split over two lines
        if m and m.group(1).startswith("scope_"):
            cur = m.group(1)
        names[n] = cur
    return names


def find_loop(ins, loop_lines):
    """the backward branch that closes the `// sass-loop` for-statement (the widest one, if ptxas split it)"""
    idx = {i["addr"]: k for k, i in enumerate(ins)}
    best = None
    for k, i in enumerate(ins):
        if i["op"] == "BRA" and i["target"] is not None and i["target"] <= i["addr"] and i["target"] in idx:
            lo = idx[i["target"]]
            if i["lines"] and i["lines"][0] in loop_lines and any(j["op"] == "LDSM" for j in ins[lo:k + 1]):
                if best is None or k - lo > best[1] - best[0]:
                    best = (lo, k)
    if best is None:
        raise SystemExit("no backward branch on a `// sass-loop` line that spans an LDSM")
    return best


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("lib", nargs="?", default=DEFAULT_LIB)
    ap.add_argument("--kernel", default=DEFAULT_KERNEL)
    ap.add_argument("--roles", action="store_true", help="budget by enclosing function of the innermost source line")
    ap.add_argument("--dump", action="store_true", help="print the fast path, one instruction per line")
    ap.add_argument("--source", default=SOURCE)
    a = ap.parse_args()

    name, ins = disassemble(a.lib, a.kernel)
    loop_lines = {n for n, ln in enumerate(open(a.source), 1) if "sass-loop" in ln}
    lo, hi = find_loop(ins, loop_lines)
    cold = cold_ranges(a.source)
    body = ins[lo:hi + 1]
    fast = [i for i in body if not any(b <= ln <= e for ln in i["lines"] for b, e in cold)]
    # pixel-warps (rows of 32 pixels) one warp handles per loop iteration: LDSM.x4 reads 4 rows, .x2 reads 2
    px_warps = sum(int(i["full_op"].rsplit(".", 1)[1]) for i in fast if i["op"] == "LDSM")
    by_pipe = collections.Counter(pipe_of(i["op"]) for i in fast)
    total = len(fast)
    print(f"{name}")
    print(f"loop {ins[lo]['addr']:#x}..{ins[hi]['addr']:#x}: {len(body)} instructions, {len(body) - total} of them cold; "
          f"fast path {total} per {px_warps} pixel-warps = {total / px_warps:.2f} per 32 pixels")
    for pipe, n in sorted(by_pipe.items(), key=lambda kv: -kv[1]):
        print(f"  {n / px_warps:6.2f}  {pipe}")
    lsu = sum(1 for i in fast if i["op"] in LSU_OPS)
    stall = sum(i["stall"] for i in fast)
    issue = total / px_warps / 4.0
    tb = lambda cyc: 148 * 128 * 1.965 / cyc / 1e3  # noqa: E731  TB/s at 1965 MHz for `cyc` cycles per 32 pixels per SM
    print(f"LSU instructions: {lsu / px_warps:.2f} per 32 pixels -> at 2.1 cycles each {lsu / px_warps * 2.1:.1f} cycles "
          f"per 32 pixels per SM = {tb(lsu / px_warps * 2.1):.2f} TB/s (bank conflicts of the vectorscope come on top)")
    print(f"issue: {issue:.2f} cycles per 32 pixels per SM at 4 x 1 instruction/clk = {tb(issue):.2f} TB/s")
    print(f"ptxas stall counts along the fast path: {stall} cycles per iteration of one warp "
          f"({stall / total:.2f} per instruction): with W warps per scheduler the loop cannot run faster than "
          f"max({total} W, {stall}) cycles per W iterations")
    if a.roles:
        fn = function_of_line(a.source)
        by_role = collections.Counter(fn.get(i["lines"][0], "?") if i["lines"] else "?" for i in fast)
        print("by role (enclosing function of the innermost source line), per 32 pixels:")
        for role, n in sorted(by_role.items(), key=lambda kv: -kv[1]):
            print(f"  {n / px_warps:6.2f}  {role}")
    if a.dump:
        for i in fast:
            print(f"  {i['addr']:#07x} s{i['stall']:<2d} {i['text']:60s} {i['lines'][:1]}")


if __name__ == "__main__":
    main()
