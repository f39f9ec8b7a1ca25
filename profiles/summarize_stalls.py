"""Warp-stall sampling of one kernel of an ncu report, aggregated by role (SASS address ranges) and by instruction.
  ncu -i REPORT.ncu-rep --page source --csv --print-source sass > src.csv
  python profiles/summarize_stalls.py src.csv tcx_forward 1 "issuer:400:1800" "tile epilogue:3400:5800" """
import csv
import re
import sys


def kernels(path):
    out, cur = [], None
    for r in csv.reader(open(path)):
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'hdr': None, 'rows': []}
            out.append(cur)
        elif cur is not None and cur['hdr'] is None:
            cur['hdr'] = r
        elif cur is not None:
            cur['rows'].append(r)
    return out


def main():
    path, pattern, which = sys.argv[1], sys.argv[2], int(sys.argv[3])
    k = [k for k in kernels(path) if pattern in k['name']][which]
    h, R = k['hdr'], k['rows']
    si, ie, src = h.index('# Samples'), h.index('Instructions Executed'), h.index('Source')
    st = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
    total = sum(int(r[si]) for r in R)
    print('kernel %s: %d SASS instructions, %d warp samples, %d warp-level instructions executed' % (
        k['name'], len(R), total, sum(int(r[ie]) for r in R)))
    agg = sorted(((sum(int(r[h.index(n)] or 0) for r in R), n[6:]) for n in st), reverse=True)
    print('stall reasons (all warps): ' + ', '.join('%s %.1f %%' % (n, 100.0 * v / total) for v, n in agg[:8]))
    lm = [i for i, r in enumerate(R) if re.search(r'\b(LDL|STL)\b', r[src])]
    print('local-memory instructions: %d static, %d executed (%.2f %% of all), %d samples on themselves' % (
        len(lm), sum(int(R[i][ie]) for i in lm), 100.0 * sum(int(R[i][ie]) for i in lm) / sum(int(r[ie]) for r in R),
        sum(int(R[i][si]) for i in lm)))
    specs = sys.argv[4:] or ['whole kernel:0:%d' % len(R)]
    for spec in specs:
        name, lo, hi = spec.rsplit(':', 2)
        lo, hi = int(lo), int(hi)
        reg = R[lo:hi]
        tot = sum(int(r[si]) for r in reg)
        print('\n== %s (SASS index %d..%d): %d samples = %.1f %% of the kernel\'s' % (name, lo, hi, tot, 100.0 * tot / total))
        agg = sorted(((sum(int(r[h.index(n)] or 0) for r in reg), n[6:]) for n in st), reverse=True)
        print('   stall reasons: ' + ', '.join('%s %.1f %%' % (n, 100.0 * v / max(tot, 1)) for v, n in agg[:6]))
        # a register written by an LDL and waited for (long scoreboard) by a later instruction = time lost to the local-memory stack
        wait_local, barrier_wait = 0, 0
        pending = {}
        for i, r in enumerate(reg):
            text = r[src].strip()
            m = re.search(r'\bLDL(?:\.\S+)?\s+(R\d+)', text)
            if m:
                pending[m.group(1)] = i
                continue
            if 'NANOSLEEP' in text or 'SYNCS.PHASECHK' in text:
                barrier_wait += int(r[si])
                continue
            lsb = int(r[h.index('stall_long_sb')] or 0)
            if lsb:
                regs = re.findall(r'\bR\d+\b', text)
                if any(x in pending for x in regs[1:] if True):
                    wait_local += lsb
                    for x in regs[1:]:
                        pending.pop(x, None)
        print('   waiting on mbarriers (NANOSLEEP / PHASECHK): %.1f %%; long-scoreboard stalls of consumers of LDL results: %.1f %%' % (
            100.0 * barrier_wait / max(tot, 1), 100.0 * wait_local / max(tot, 1)))
        top = sorted(((int(r[si]), i + lo, r[src].strip()) for i, r in enumerate(reg)), reverse=True)[:12]
        for s, i, text in top:
            print('   %7d  [%5d]  %s' % (s, i, text[:90]))


if __name__ == '__main__':
    main()
