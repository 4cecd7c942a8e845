"""Development aid: where a kernel's warp-stall samples and executed instructions are, per CUDA source line or per SASS
instruction, from an `ncu --set full --import-source on` report — read offline (no GPU needed).

    python scripts/ncu_source.py gpurun_out/r2_full.ncu-rep k_finish:2 [top N lines, default 40] [--sass]

The kernel is `<name>:<n-th launch of that name in the report>` (ncu's --kernel-id ::name:n).  Default view: one row per source
line of rlb_boost.cu — share of the kernel's samples, warp instructions executed, the line, its three largest stall reasons.
--sass: per instruction, plus the opcode mix weighted by execution count.  The kernel must have been built with -lineinfo.
"""
import csv
import io
import subprocess
import sys
from collections import Counter


def page(rep, kern, source):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", source, "--kernel-id", f"::{kern}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next((r for r in rows if len(r) > 10 and r[0] in ("Line No", "Address")), None)
    if hdr is None:
        raise SystemExit(f"no kernel {kern!r} in {rep} (ncu --page raw lists the launches)")
    ix = {}
    for i, h in enumerate(hdr):
        ix.setdefault(h, i)
    data = []
    for r in rows[rows.index(hdr) + 1:]:
        if len(r) == len(hdr) and r[0] == hdr[0]:
            break                                       # ncu prints the table of a matched launch twice: keep the first
        if len(r) == len(hdr):
            data.append(r)
    return hdr, ix, data


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    rep, kern = args[0], args[1]
    top = int(args[2]) if len(args) > 2 else 40
    sass = "--sass" in sys.argv
    hdr, ix, data = page(rep, kern, "sass" if sass else "cuda,sass")
    stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    if not sass:
        data = [r for r in data if r[0] != ""]          # the per-line rows (the SASS rows have an empty line number)
    tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
    ninst = sum(int(r[ix["Instructions Executed"]]) for r in data) or 1
    agg = {h: sum(int(r[ix[h]]) for r in data) for h in stall}
    print(f"{kern}: {tot} samples, {ninst} warp instructions")
    print("  " + ", ".join(f"{h[6:]} {100 * v / tot:.0f}%" for h, v in sorted(agg.items(), key=lambda x: -x[1])[:7]))
    if sass:
        mix = Counter()
        for r in data:
            p = r[ix["Source"]].split()
            mix[(p[1] if p[0].startswith("@") else p[0]).split(".")[0]] += int(r[ix["Instructions Executed"]])
        print("  " + ", ".join(f"{k} {100 * v / ninst:.1f}%" for k, v in mix.most_common(16)))
    order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:top]
    for i in sorted(order):
        r = data[i]
        st = sorted(((h[6:], int(r[ix[h]])) for h in stall if int(r[ix[h]]) > 0), key=lambda x: -x[1])[:3]
        what = r[ix["Source"]].strip() if sass else f"{r[0]:>5s} {r[1].strip()}"
        print(f"{100 * int(r[ix['# Samples']]) / tot:5.1f}% {r[ix['Instructions Executed']]:>9s}  {what[:100]:100s} {st}")


if __name__ == "__main__":
    main()
