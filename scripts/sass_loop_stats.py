#!/usr/bin/env python
"""Static look at the FIR-bank kernel's inner loop in the built library (no GPU needed).

Finds the FFMA2 body loop (the backward-branch loop with the highest FFMA2 density) of every fir_bank_kernel instantiation in `cuobjdump -sass` output and reports
its instruction mix, how many FFMA2 carry an operand-reuse flag, the yield hints ptxas placed inside the FFMA2
stream (control-word bit 45 clear), and the sum of the static stall counts (the schedule's own cycle estimate).

    python scripts/sass_loop_stats.py [path/to/librawboost_b200.so]
"""
import collections
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "..", "scl-deepfake-audio-detection_b200", "lib", "librawboost_b200.so")


def functions(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout.split("\n")
    name, cur = None, []
    for i, line in enumerate(txt):
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                yield name, cur
            name, cur = m.group(1), []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", line)
        if m and name:
            hi = re.search(r"/\* (0x[0-9a-f]+) \*/", txt[i + 1])
            cur.append((int(m.group(1), 16), m.group(2).strip(), int(hi.group(1), 16) if hi else 0))
    if name:
        yield name, cur


def body_loop(ins):
    """The FFMA2 body loop of a kernel: the backward-branch loop whose instructions are mostly FFMA2 (None if there is none)."""
    loops = []
    for a, t, _ in ins:
        m = re.search(r"BRA\S*\s+(?:\S+,\s+)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            loops.append((int(m.group(1), 16), a))
    if not loops:
        return None

    def density(loop):
        b = [x for x in ins if loop[0] <= x[0] <= loop[1]]
        f = sum("FFMA2" in t for _, t, _ in b)
        return (f / len(b) if f >= 100 else 0.0, f)

    lo, hi_addr = max(loops, key=density)
    return [x for x in ins if lo <= x[0] <= hi_addr]


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else LIB
    for name, ins in functions(path):
        if "fir_bank_kernel" not in name:
            continue
        body = body_loop(ins)
        if not body:
            continue
        lo, hi_addr = body[0][0], body[-1][0]
        mix = collections.Counter((t.split()[1] if t.startswith("@") else t.split()[0]) for _, t, _ in body)
        ff = [x for x in body if "FFMA2" in x[1]]
        reuse = sum("reuse" in t for _, t, _ in ff)
        yields = sum(((h >> 45) & 1) == 0 for _, _, h in ff)
        stalls = sum((h >> 41) & 0xF for _, _, h in body)
        tmpl = re.search(r"fir_bank_kernelILi(\d)E", name)
        print(f"fir_bank_kernel<{tmpl.group(1) if tmpl else '?'}>: loop {lo:#x}..{hi_addr:#x}, {len(body)} instructions, mix {dict(mix)}")
        print(f"  FFMA2 {len(ff)}: reuse-flagged {reuse} ({100.0 * reuse / max(1, len(ff)):.1f} %), yield hints inside the stream {yields}"
              f" (one per {len(ff) / max(1, yields):.1f} FFMA2)")
        print(f"  static stall sum {stalls} cycles for {2 * len(ff)} FFMA2 issue cycles -> schedule efficiency {200.0 * len(ff) / stalls:.1f} %")


if __name__ == "__main__":
    main()
