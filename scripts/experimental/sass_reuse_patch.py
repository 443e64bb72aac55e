#!/usr/bin/env python
"""EXPERIMENT (not part of the build): patch the control words of the FFMA2 body loop of fir_bank_kernel instantiations inside a
built library -- drop ptxas's every-sixth-FFMA2 yield hints and flag the tap operand for operand-cache reuse wherever the next
FFMA2 of the stream multiplies by the same tap pair -- to measure what the FFMA2 stream could do with a denser reuse pattern.

    python scripts/experimental/sass_reuse_patch.py <in.so> <out.so> [mode]     mode: yield | reuse | both (default)
Control word (high 64 bits of the 128-bit instruction, sm_70+): stall 41-44, yield 45 (0 = may yield), write barrier 46-48,
read barrier 49-51, wait mask 52-57, reuse flags 58-61 (58 = operand A)."""
import re
import subprocess
import sys

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import sass_loop_stats as S


def encodings(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout.split("\n")
    name, out = None, {}
    for i, line in enumerate(txt):
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            out[name] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", line)
        if m and name:
            hi = re.search(r"/\* (0x[0-9a-f]+) \*/", txt[i + 1])
            out[name].append((int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16), int(hi.group(1), 16)))
    return out


def opA(text):
    m = re.match(r"FFMA2 \S+, (U?R\d+)", text)
    return m.group(1) if m else None


def written(text):
    """Registers an instruction of the loop writes (FFMA2: a pair; LDS.128: a quad; others: conservatively their first operand)."""
    m = re.match(r"(\S+) (U?R)(\d+)", text)
    if not m or m.group(2) == "UR":
        return set()
    n = int(m.group(3))
    width = 4 if m.group(1).startswith("LDS.128") else 2 if m.group(1).startswith(("FFMA2", "LDS.64")) else 1
    return {f"R{n + i}" for i in range(width)}


def pair(reg):
    n = int(reg[1:])
    return {f"R{n}", f"R{n + 1}"}


def main():
    src, dst = sys.argv[1], sys.argv[2]
    mode = sys.argv[3] if len(sys.argv) > 3 else "both"
    which = sys.argv[4] if len(sys.argv) > 4 else "fir_bank_kernel"
    scope = sys.argv[5] if len(sys.argv) > 5 else "loop"
    blob = bytearray(open(src, "rb").read())
    enc = encodings(src)
    for name, ins in enc.items():
        if "fir_bank_kernel" not in name or which not in name:
            continue
        key = b"".join(lo.to_bytes(8, "little") + hi.to_bytes(8, "little") for _, _, lo, hi in ins[:8])
        assert blob.count(key) == 1, (name, blob.count(key))
        base = blob.find(key) - ins[0][0]
        if scope == "loop":
            loop = S.body_loop([(a, t, hi) for a, t, lo, hi in ins])
            lo_addr, hi_addr = loop[0][0], loop[-1][0]
            body = [x for x in ins if lo_addr <= x[0] <= hi_addr]
        else:
            body = ins
        # straight-line runs only: a pair of FFMA2 is considered when no branch, branch target or barrier lies between them
        targets = {int(m.group(1), 16) for _, t, _, _ in ins for m in [re.search(r"\b(?:BRA|BSSY\S*|CALL\S*)\b.*?(0x[0-9a-f]+)", t)] if m}
        flow = re.compile(r"^(@!?U?P\d+ )?(BRA|BSYNC|BSSY|EXIT|RET|CALL|BAR|WARPSYNC|NANOSLEEP|YIELD|BREAK|JMP|BRX)")
        ff = [k for k, x in enumerate(body) if x[1].startswith("FFMA2")]
        straight = {}
        for pos in range(len(ff) - 1):
            k, j = ff[pos], ff[pos + 1]
            straight[k] = not any(flow.match(body[i][1]) for i in range(k + 1, j)) and not any(body[i][0] in targets for i in range(k + 1, j + 1))
        n_y = n_r = n_ya = 0
        for pos, k in enumerate(ff):
            addr, text, lo, hi = body[k]
            new = hi
            same = False
            if pos + 1 < len(ff) and straight[k] and opA(text).startswith("R") and opA(body[ff[pos + 1]][1]) == opA(text):
                # the tap pair must still hold the same value when the next FFMA2 reads it: nothing from this instruction up to
                # (not including) the next FFMA2 may write it
                clobber = set().union(*[written(body[j][1]) for j in range(k, ff[pos + 1])])
                same = not (clobber & pair(opA(text)))
            had_yield = not (hi >> 45) & 1
            if mode == "yield":
                if had_yield:
                    new |= 1 << 45
                    n_y += 1
            elif mode == "bothnoY":
                if had_yield:
                    new |= 1 << 45
                    n_y += 1
                if same and not (new >> 58) & 1:
                    new |= 1 << 58
                    n_r += 1
            elif mode in ("both", "runend"):
                if same:
                    if had_yield:
                        new |= 1 << 45
                        n_y += 1
                    if not (new >> 58) & 1:
                        new |= 1 << 58
                        n_r += 1
                elif mode == "runend" and not had_yield:
                    new &= ~(1 << 45)   # a yield hint at the end of every tap run (no reuse is possible there anyway)
                    n_ya += 1
            elif mode == "reuseY":   # keep ptxas's yields, flag reuse regardless (relies on the hardware dropping the cache on a switch)
                if same and not (new >> 58) & 1:
                    new |= 1 << 58
                    n_r += 1
            if new != hi:
                off = base + addr + 8
                assert int.from_bytes(blob[off:off + 8], "little") == hi
                blob[off:off + 8] = new.to_bytes(8, "little")
        print(f"{name[-58:]}: {len(ff)} FFMA2 in the loop, {n_y} yield hints removed, {n_ya} added, {n_r} reuse flags added")
    open(dst, "wb").write(blob)


if __name__ == "__main__":
    main()
