"""Development aid: static issue-cycle count of a kernel's loops from its SASS control words — no GPU needed.

    python scripts/sass_stalls.py ranklib_b200/csrc/rlb_boost.o k_hist_rootILi1E [min loop length, default 60]

`cuobjdump -sass` prints two 64-bit words per instruction; bits 41-44 of the second one are the stall count ptxas assigned
(cycles before the same warp may issue its next instruction), bits 46-48 / 49-51 the write / read scoreboard the instruction
arms, bits 52-57 the scoreboards it waits for.  For a kernel that runs ONE warp per scheduler — the consumers of
k_hist_root / k_hist_child: 96 private histograms = 3 warps on 4 schedulers — the sum of the stall counts over a loop body
is the loop's issue time; ncu's `selected` + `wait` samples of the same lines add up to it (735 cycles per
stage against 981 measured for the root kernel of round 2; the rest is scoreboard waits).  The sum predicted every variant
of profiles/r2x_variants_hist_*.jsonl in sign and roughly in size, which is how changes were chosen between GPU calls.

Prints every backward-branch loop of at least the given length: instructions, stall-field sum, opcode mix, and the stall
cycles by (pipe of the instruction, pipe of its successor) — back-to-back ALU-pipe instructions cost 2 cycles each.
"""
import re
import subprocess
import sys
from collections import Counter, defaultdict

ALU = {"SEL", "IADD3", "ISETP", "LOP3", "SHF", "VIADD", "PLOP3", "LEA", "PRMT", "IABS", "MOV", "FSEL", "FSETP", "IMNMX", "VIADDMNMX"}
FMA = {"IMAD", "FFMA", "FMUL", "FADD", "HFMA2"}
LSU = {"LDS", "STS", "LDG", "STG", "LDGSTS", "SYNCS", "ATOMS", "REDG", "ATOMG", "UBLKCP"}


def parse(obj, pattern):
    text = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, errors="replace").stdout
    on, ins, cur = False, [], None
    for line in text.splitlines():
        if "Function :" in line:
            on = pattern in line
            continue
        if not on:
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?)\s*;\s*/\* (0x[0-9a-f]{16}) \*/", line)
        if m:
            cur = [int(m.group(1), 16), m.group(2), None]
            ins.append(cur)
            continue
        m = re.match(r"\s+/\* (0x[0-9a-f]{16}) \*/", line)
        if m and cur is not None and cur[2] is None:
            cur[2] = int(m.group(1), 16)
    return [dict(addr=a, text=t, stall=(hi >> 41) & 0xF, wbar=(hi >> 46) & 7, rbar=(hi >> 49) & 7, wait=(hi >> 52) & 0x3F)
            for a, t, hi in ins if hi is not None]


def opcode(text):
    p = text.split()
    return (p[1] if p[0].startswith("@") else p[0]).split(".")[0]


def pipe(text):
    op = opcode(text)
    return "alu" if op in ALU else "fma" if op in FMA else "lsu" if op in LSU else "other"


def main():
    obj, pattern = sys.argv[1], sys.argv[2]
    min_len = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    ins = parse(obj, pattern)
    if not ins:
        raise SystemExit(f"no function matching {pattern!r} in {obj}")
    index = {d["addr"]: i for i, d in enumerate(ins)}
    print(f"{pattern}: {len(ins)} instructions")
    loops = []
    for i, d in enumerate(ins):
        m = re.search(r"BRA(?:\.[A-Z.]+)?\s+(0x[0-9a-f]+)", d["text"])
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt <= d["addr"] and tgt in index and i - index[tgt] >= min_len:
            loops.append((index[tgt], i))
    # innermost loops only (the mbarrier retry branches wrap whole regions of the kernel)
    loops = [l for l in loops if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]
    for lo, hi in loops:
        body = ins[lo:hi + 1]
        stalls = sum(max(x["stall"], 1) for x in body)
        mix = Counter(opcode(x["text"]) for x in body)
        print(f"  loop {body[0]['addr']:#x}..{body[-1]['addr']:#x}: {len(body)} instructions, stall-field sum {stalls} cycles")
        print("     " + ", ".join(f"{k} {v}" for k, v in mix.most_common(14)))
        tr = defaultdict(lambda: [0, 0])
        for a, b in zip(body, body[1:]):
            k = (pipe(a["text"]), pipe(b["text"]))
            tr[k][0] += 1
            tr[k][1] += max(a["stall"], 1)
        print("     " + "; ".join(f"{a}->{b}: {n} x {c / n:.2f}" for (a, b), (n, c) in sorted(tr.items()) if n >= 10))


if __name__ == "__main__":
    main()
