#!/usr/bin/env python
"""tools/ncu_phase_table.py REPORT.ncu-rep KERNEL_REGEX > table.md — splits the SASS of a kernel at its CTA barriers and reports,
per barrier-delimited phase, the share of warp-stall samples, the executed warp instructions, the top stall reasons and the
opcode mix (from `ncu --page source --csv --print-source sass`).  How the phase attribution of DESIGN.md §4 was made."""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", f"regex:{pat}"],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    print(f"kernel: `{rows[start - 1][1][:110]}`\n")
    hdr, data = rows[start], []
    for r in rows[start + 1:]:
        if r and r[0] == "Kernel Name":
            break                          # first matching launch only
        if len(r) == len(hdr):
            data.append(r)
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    total = sum(int(r[ix["# Samples"]] or 0) for r in data)
    print("| phase (ends at) | samples | warp instr. | top stalls | opcode mix (warp instr.) |\n|---|---|---|---|---|")
    n = ex = 0
    st, ops = collections.Counter(), collections.Counter()
    phase = 0
    for r in data + [None]:
        if r is not None:
            src = r[ix["Source"]].split()
            op = (src[1] if src[0].startswith("@") else src[0]).split(".")[0]
            k = int(r[ix["Instructions Executed"]] or 0)
            n += int(r[ix["# Samples"]] or 0); ex += k; ops[op] += k
            for c in stall_cols:
                st[c] += int(r[ix[c]] or 0)
        if r is None or op == "BAR":
            if n:
                top = ", ".join(f"{k[6:]} {100 * v // max(1, n)}%" for k, v in st.most_common(4))
                mix = ", ".join(f"{k} {v}" for k, v in ops.most_common(6))
                print(f"| {phase} ({'BAR' if r is not None else 'end'}) | {100 * n / total:.1f}% | {ex} | {top} | {mix} |")
            phase += 1
            n = ex = 0
            st, ops = collections.Counter(), collections.Counter()


if __name__ == "__main__":
    main()
