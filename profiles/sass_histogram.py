"""cuobjdump -sass of the shipped library -> opcode counts per kernel (the proof that the hot kernels issue tcgen05 / TMEM /
bulk-copy / TMA instructions and no legacy mma.sync), plus LDL / STL (local-memory traffic: at a 223 KB shared-memory carve-out
the L1 that backs the stack is nearly gone) and the registers / spills ptxas reported.
  python profiles/sass_histogram.py > profiles/r02_sass_opcode_histogram.txt"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'flowket_b200', 'libflowket_b200.so')
COLS = ['UTCHMMA', 'UTCQMMA', 'UTCBAR', 'LDTM', 'STTM', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'SYNCS', 'HMMA', 'LDGSTS', 'ATOMG', 'RED',
        'LDL', 'STL']


def demangle(names):
    out = subprocess.run(['c++filt'], input='\n'.join(names), stdout=subprocess.PIPE, universal_newlines=True).stdout.split('\n')
    return dict(zip(names, out))


def ptxas_info():
    """mangled name -> (registers, spill stores, spill loads, stack bytes) from the build logs of `make`"""
    info = {}
    for log in glob.glob(os.path.join(ROOT, 'flowket_b200', 'csrc', 'build', '*.ptxas.log')):
        cur = None
        for line in open(log):
            m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", line)
            if m:
                cur = m.group(1)
                info[cur] = [0, 0, 0, 0]
                continue
            if cur is None:
                continue
            m = re.search(r'(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads', line)
            if m:
                info[cur][3], info[cur][1], info[cur][2] = int(m.group(1)), int(m.group(2)), int(m.group(3))
            m = re.search(r'Used (\d+) registers', line)
            if m:
                info[cur][0] = int(m.group(1))
    return info


def collect():
    """-> (opcode counts per mangled kernel name, instruction totals, demangled names)"""
    sass = subprocess.run(['cuobjdump', '-sass', SO], stdout=subprocess.PIPE, universal_newlines=True).stdout
    counts, total, cur = collections.defaultdict(collections.Counter), collections.Counter(), None
    for line in sass.split('\n'):
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
        if m and cur:
            op = m.group(1)
            total[cur] += 1
            counts[cur][op] += 1
    return counts, total, demangle(list(total))


def main():
    counts, total, names = collect()
    pt = ptxas_info()
    print('# cuobjdump -sass flowket_b200/libflowket_b200.so: opcode counts per kernel (tcgen05.mma = UTCHMMA, tcgen05.ld/st = LDTM/STTM,')
    print('# cp.async.bulk = UBLKCP, cp.async.bulk.tensor = UTMALDG, tcgen05.commit = UTCBAR, mbarrier = SYNCS; no HMMA = no legacy mma.sync')
    print('# path; LDL / STL = local-memory loads / stores); regs / spill bytes / stack bytes from ptxas -v')
    print('%-64s %6s %s %5s %6s %6s' % ('kernel', 'insts', ' '.join('%7s' % c for c in COLS), 'regs', 'spill', 'stack'))
    for k in sorted(total, key=lambda k: -total[k]):
        name = re.sub(r'\(.*', '', names.get(k, k))
        r = pt.get(k, [0, 0, 0, 0])
        print('%-64s %6d %s %5d %6d %6d' % (name[:64], total[k], ' '.join('%7d' % counts[k][c] for c in COLS), r[0], r[1] + r[2], r[3]))


if __name__ == '__main__':
    main()
