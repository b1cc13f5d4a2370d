"""Dynamic instruction mix of one Poseidon permutation from the SASS of permute_kernel (rolled round loops:
full-round body x8, partial-round body x22, the rest x1).  usage: python tools/sass_mix.py lib.so"""
import collections, re, subprocess, sys

def classify(op):
    if op.startswith('IMAD.WIDE'): return 'WIDE'
    if op.startswith('IMAD.HI'): return 'IMADHI'
    if op.startswith('IMAD'): return 'IMAD'
    if op.startswith(('IADD3', 'LOP3', 'SHF', 'SEL', 'ISETP', 'VIADD', 'MOV', 'PLOP3', 'LEA', 'PRMT', 'IABS', 'FLO')): return 'ALU'
    if op.startswith(('LDC', 'ULDC', 'LDCU')): return 'LDC'
    if op.startswith('U') or op.startswith('BRA') or op.startswith('R2UR'): return 'UNI'
    return 'OTHER'

def mix(so, kernel='permute_kernel'):
    txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
    blocks = re.split(r'\n\s*Function : ', txt)
    body = [b for b in blocks if b.startswith('_ZN6merkle14' + kernel) or kernel in b.split('\n')[0]][0]
    ins = []
    for l in body.splitlines():
        m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)(.*)', l)
        if m: ins.append((int(m.group(1), 16), m.group(2), l))
    loops = []
    for a, op, l in ins:
        if op.startswith('BRA'):
            m = re.search(r'0x([0-9a-f]+)\s*;', l)
            if m and int(m.group(1), 16) < a: loops.append((int(m.group(1), 16), a))
    loops.sort(key=lambda t: t[1] - t[0])
    # innermost two loops: partial (shorter) and full (longer) -- identify by size
    inner = [lp for lp in loops if not any(o != lp and lp[0] <= o[0] and o[1] <= lp[1] for o in loops)]
    inner.sort(key=lambda t: t[1] - t[0])
    partial, full = inner[0], inner[1]
    outer = max(loops, key=lambda t: t[1] - t[0])
    tot = collections.Counter()
    for a, op, l in ins:
        if not (outer[0] <= a <= outer[1]):
            continue
        c = classify(op)
        w = 8 if full[0] <= a <= full[1] else 22 if partial[0] <= a <= partial[1] else 1
        # code between the two inner loops but inside the `half` loop body executes once (half==0 branch) except the
        # part before the full loop; both approximated as x1
        tot[c] += w
    return tot, (full[1] - full[0]) // 16 + 1, (partial[1] - partial[0]) // 16 + 1

if __name__ == '__main__':
    for so in sys.argv[1:]:
        t, f, p = mix(so)
        print(so, 'full-round body', f, 'partial body', p, dict(t), 'total', sum(t.values()))
