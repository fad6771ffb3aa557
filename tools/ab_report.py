"""Markdown table of an A/B run: python tools/ab_report.py [gpurun_out/ab] > profiles/ab_rNN.md
One row per bench line (tools/run_ab.sh, tools/run_multigpu.sh write one JSON line per file); the in-tree
library's `new_mixed.json` is the reference for the ratio column when present."""
import glob
import json
import os
import sys


def last_json_line(path):
    try:
        lines = [l for l in open(path).read().strip().splitlines() if l.startswith("{")]
        return json.loads(lines[-1]) if lines else None
    except (OSError, ValueError):
        return None


def main():
    d = sys.argv[1] if len(sys.argv) > 1 else os.path.join("gpurun_out", "ab")
    rows = []
    for f in sorted(glob.glob(os.path.join(d, "*.json"))):
        j = last_json_line(f)
        if not j or "value" not in j:
            rows.append((os.path.basename(f)[:-5], None))
            continue
        rows.append((os.path.basename(f)[:-5], j))
    base = dict(rows).get("new_mixed")
    print("| run | frames/s | vs in-tree (mixed) | kernel GB/s | of HBM peak | kernel ms | SM MHz | throttle | launches |")
    print("|---|---|---|---|---|---|---|---|---|")
    for name, j in rows:
        if j is None:
            print(f"| `{name}` | no result | | | | | | | |")
            continue
        r = j.get("roofline") or {}
        c = j.get("clocks") or {}
        ratio = f"{j['value'] / base['value']:.3f}" if base and name.endswith("_mixed") else ""
        print(f"| `{name}` | {j['value']:,.0f} | {ratio} | {r.get('achieved', 0):,.0f} | "
              f"{100 * r.get('frac', 0):.1f} % | {r.get('kernel_ms', 0):.3f} | {c.get('sm_mhz', '')} | "
              f"{','.join(c.get('reasons', []) or []) or '-'} | {j.get('gpu_launches', '')} |")
    for f in sorted(glob.glob(os.path.join(d, "pytest*.full"))):
        tail = [l for l in open(f).read().strip().splitlines() if l.strip()][-2:]
        print(f"\n`{os.path.basename(f)}`: {' / '.join(tail)}")


if __name__ == "__main__":
    main()
