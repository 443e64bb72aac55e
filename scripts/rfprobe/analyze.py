import re, subprocess, sys
from collections import Counter
sass = subprocess.run(['cuobjdump', '-sass', 'rfprobe'], capture_output=True, text=True).stdout
cur = None; per = {}
for l in sass.splitlines():
    m = re.search(r'Function : _Z\d*k(\d+)i', l)
    if m: cur = int(m.group(1)); per[cur] = []; continue
    if cur is not None and re.search(r'/\*[0-9a-f]{4}\*/', l): per[cur].append(l)
for v in sorted(per):
    lines = per[v]; recs = []
    for n, l in enumerate(lines):
        m = re.search(r'FFMA2 (R\d+), (R\d+)(\.reuse)?\.F32x2\.HI_LO, (R\d+)(\.reuse)?\.F32x2\.HI_LO, (R\d+)(\.reuse)?\.F32x2\.HI_LO', l)
        if m:
            d, a, ar, b, br, c, cr = m.groups(); recs.append((n, int(a[1:]), bool(ar), int(b[1:]), bool(br), int(c[1:]), bool(cr)))
    coll = sum(1 for r in recs if (r[3] % 4) // 2 == (r[5] % 4) // 2)
    reads3 = 0
    for i, r in enumerate(recs):
        served = 0
        if i > 0 and recs[i - 1][0] == r[0] - 1:
            p = recs[i - 1]
            served += (p[2] and p[1] == r[1]) + (p[4] and p[3] == r[3]) + (p[6] and p[5] == r[5])
        reads3 += (served == 0)
    print(f'k{v}: FFMA2={len(recs)} window/acc same-class={coll / len(recs):.2f} no-reuse={reads3 / len(recs):.2f} other_instrs={len(lines) - len(recs)}')
