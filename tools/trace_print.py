"""Print the pipelined 1-D kernel's debug timeline (gpurun_out/trace_{fwd,bwd}.txt) relative to the first stamp."""
import re, sys
def load(fn):
    d = {}
    for ln in open(fn):
        m = re.match(r'cta (\d+) it (\d+) role (\d+): (.*)', ln)
        d[(int(m[1]), int(m[2]), int(m[3]))] = [int(x) for x in m[4].split()]
    return d
fn = sys.argv[1]; ctas = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [1]
lo, hi = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (203, 209)
d = load(fn)
for c in ctas:
    base = d[(c, 200, 0)][0]
    print(fn, 'cta', c)
    for it in range(lo, hi):
        for r in range(4):
            print(' it', it, 'role', r, ' '.join('%6d' % (x - base) if x else '     .' for x in d[(c, it, r)]))
