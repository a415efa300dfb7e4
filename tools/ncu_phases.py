#!/usr/bin/env python
"""Phase profile of one kernel of an `ncu --set full --import-source on` capture (no GPU needed).

  python tools/ncu_phases.py rep.ncu-rep <kernel regex> [launch index, default 0] [bucket, default 32]

Splits the SASS stream into runs with the same execution count (= loop nests: once per step, per flush block, per TMA
issue, ...) and prints for each run: instructions, executions, issued warp-instructions, stall samples, shared-memory
wavefronts and the top stall reasons - which phase of the kernel costs what."""
import csv
import subprocess
import sys

REASONS = ["stall_long_sb", "stall_wait", "stall_short_sb", "stall_math", "stall_mio", "stall_lg", "stall_not_selected",
           "stall_selected", "stall_dispatch", "stall_branch_resolving", "stall_barrier", "stall_membar", "stall_drain", "stall_no_inst"]


def main(path, kernel, launch=0):
    cmd = ["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}", "--launch-skip", str(launch), "--launch-count", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr_i = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
    hdr = rows[hdr_i]
    ix = {h: i for i, h in enumerate(hdr)}
    data, seen = [], set()
    for r in rows[hdr_i + 1:]:  # the page repeats the listing (two views of the same SASS): keep the first row of every address
        if len(r) == len(hdr) and r[ix["Address"]] not in seen:
            seen.add(r[ix["Address"]])
            data.append(r)

    def num(r, k):
        try:
            return int(float(r[ix[k]] or 0))
        except (KeyError, ValueError):
            return 0

    print(rows[0][1] if rows and len(rows[0]) > 1 else "")
    tot_s = sum(num(r, "# Samples") for r in data)
    tot_i = sum(num(r, "Instructions Executed") for r in data)
    tot_w = sum(num(r, "L1 Wavefronts Shared") for r in data)
    print(f"{len(data)} SASS instructions, {tot_i} warp-instructions issued, {tot_s} samples, {tot_w} shared wavefronts")
    runs = []
    for r in data:
        ex = num(r, "Instructions Executed")
        if runs and (abs(runs[-1]["ex"] - ex) <= 0.02 * max(ex, 1)):
            runs[-1]["rows"].append(r)
        else:
            runs.append({"ex": ex, "rows": [r]})
    # merge tiny runs into neighbours for readability
    print(f"{'first':>8} {'n_ins':>6} {'exec':>10} {'issued':>12} {'%iss':>6} {'samples':>8} {'%smp':>6} {'smem_wf':>10} {'fp64':>5}  top stalls")
    for run in runs:
        rs = run["rows"]
        if len(rs) < 4 and run["ex"] < 0.001 * tot_i:
            continue
        iss = sum(num(r, "Instructions Executed") for r in rs)
        smp = sum(num(r, "# Samples") for r in rs)
        wf = sum(num(r, "L1 Wavefronts Shared") for r in rs)
        fp64 = sum(1 for r in rs if r[ix["Source"]].strip().lstrip("@!UP0123456789 ").startswith(("DFMA", "DMUL", "DADD")))
        st = sorted(((sum(num(r, k) for r in rs), k[6:]) for k in REASONS if k in ix), reverse=True)[:3]
        print(f"{rs[0][ix['Address']][-6:]:>8} {len(rs):6d} {run['ex']:10d} {iss:12d} {100.0 * iss / max(tot_i, 1):6.1f} {smp:8d} {100.0 * smp / max(tot_s, 1):6.1f} {wf:10d} {fp64:5d}  "
              + ", ".join(f"{n} {v}" for v, n in st if v))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
