"""Opcode histogram of a kernel's SASS and of its loops (backward branches).
Usage: python tools/sass_loops.py <lib.so|cubin> <kernel-name-substring> [--loops]
Evidence for DESIGN.md / profiles: counts DFMA/HMMA/LDGSTS/... per loop body without a GPU."""
import collections
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, body = None, {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            body[cur] = []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            body[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return body


def opcode(text):
    t = text.split()
    if t[0].startswith("@"):
        t = t[1:]
    return t[0]


def hist(instrs, full=False):
    c = collections.Counter()
    for _, t in instrs:
        op = opcode(t)
        c[op if full else op.split(".")[0]] += 1
    return c


def main():
    path, pat = sys.argv[1], sys.argv[2]
    show_loops = "--loops" in sys.argv
    full = "--full" in sys.argv
    for name, ins in kernels(path).items():
        if pat not in name or not ins:
            continue
        print(f"== {name}: {len(ins)} instructions")
        h = hist(ins, full)
        print("   " + ", ".join(f"{k} {v}" for k, v in h.most_common(40)))
        if not show_loops:
            continue
        addr = {a: i for i, (a, _) in enumerate(ins)}
        loops = []
        for i, (a, t) in enumerate(ins):
            m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", t)
            if m and "BRA" in opcode(t):
                tgt = int(m.group(1), 16)
                if tgt <= a and tgt in addr:
                    loops.append((addr[tgt], i))
        for lo, hi in sorted(loops):
            inner = [l for l in loops if l != (lo, hi) and lo <= l[0] and l[1] <= hi]
            body = ins[lo:hi + 1]
            h = hist(body, full)
            print(f" loop 0x{ins[lo][0]:x}..0x{ins[hi][0]:x}: {len(body)} instr, {len(inner)} nested loops")
            print("    " + ", ".join(f"{k} {v}" for k, v in h.most_common(30)))


if __name__ == "__main__":
    main()
