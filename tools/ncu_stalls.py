#!/usr/bin/env python
"""Where a kernel stalls, from an `ncu --set full --import-source on` capture (no GPU needed).

  python tools/ncu_stalls.py gpurun_out/x.ncu-rep [top N, default 40]

Runs `ncu -i <rep> --page source --csv`, prints the N SASS instructions with the most warp-stall
samples (address, samples, executions, opcode, three largest stall reasons) and a coarse histogram of
samples over the instruction stream (64 instructions per bucket) to tell the phases of a kernel apart."""
import csv
import subprocess
import sys

REASONS = ["stall_long_sb", "stall_wait", "stall_short_sb", "stall_math", "stall_mio", "stall_lg", "stall_not_selected",
           "stall_selected", "stall_dispatch", "stall_branch_resolving", "stall_barrier", "stall_membar", "stall_drain"]


def main(path, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr_i = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
    hdr = rows[hdr_i]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]

    def num(r, k):
        try:
            return int(float(r[ix[k]] or 0))
        except (KeyError, ValueError):
            return 0

    total = sum(num(r, "# Samples") for r in data)
    print(f"{rows[0][1] if rows and len(rows[0]) > 1 else ''}\ntotal samples {total}, {len(data)} instructions")
    for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:top]:
        st = sorted(((num(r, k), k[6:]) for k in REASONS if k in ix), reverse=True)[:3]
        print(f"{r[ix['Address']][-6:]} {num(r, '# Samples'):7d} {num(r, 'Instructions Executed'):10d}  {r[ix['Source']][:64]:64s} "
              + ", ".join(f"{n} {v}" for v, n in st if v))
    print("\nsamples per 64 instructions (first address, samples, max executions):")
    for k in range(0, len(data), 64):
        ch = data[k:k + 64]
        print(f"{ch[0][ix['Address']][-6:]} {sum(num(r, '# Samples') for r in ch):8d} {max(num(r, 'Instructions Executed') for r in ch):10d}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
