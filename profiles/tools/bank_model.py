"""Bank model of the face / LDG phases of k_btp_stage_pair<5,9> (warp per element): wavefronts of every 64-bit shared-memory access
for a given layout, against the conflict-free count (one wavefront per active half-warp).  python profiles/tools/bank_model.py"""
import itertools
G, Q = 5, 9
NP = G * G

def wf(addrs):
    tot = ideal = 0
    for h in range(2):
        banks = {}
        for lane, w in addrs:
            if lane // 16 == h: banks.setdefault(w % 16, set()).add(w)
        if banks:
            tot += max(len(v) for v in banks.values()); ideal += 1
    return tot, ideal

def face_node(s, n): return [n, (G - 1) * G + n, n * G, n * G + G - 1][s]

class Acc:
    def __init__(self): self.d = {}
    def add(self, name, addrs):
        t, i = wf(addrs); a = self.d.setdefault(name, [0, 0]); a[0] += t; a[1] += i

def model(L):
    A = Acc()
    NOD, X, T, LR = 0, 225, 225 + 712, 225 + 712 + 456      # region bases as in PairRec<5,9> (S_NOD, S_X, S_T, S_L)
    # ---- gradient lines: lanes -> (kind, f, l)
    def grad_lanes():
        out = []
        for lane in range(32):
            if L['gsplit']:
                kind, r = lane >> 4, lane & 15
                if r >= 2 * G: continue
            else:
                if lane >= 4 * G: continue
                kind, r = lane // (2 * G), lane % (2 * G)
            out.append((lane, kind, r // G, r % G))
        return out
    for name, src, dst in (('P2g', lambda k_, f: NOD + (7 + f) * NP, lambda k_, f: LR + (2 * k_ + f) * NP),
                           ('P6l', lambda k_, f: LR + (8 + 2 * k_ + f) * NP, lambda k_, f: LR + (2 * k_ + f) * NP),
                           ('P9g', lambda k_, f: NOD + (7 + f) * NP, lambda k_, f: LR + (2 * k_ + f) * NP)):
        for k in range(G):
            a, b = [], []
            for lane, kind, f, l in grad_lanes():
                off, st = (l, G) if kind else (l * G, 1)
                a.append((lane, src(kind, f) + off + k * st)); b.append((lane, dst(kind, f) + off + k * st))
            A.add(name, a); A.add(name, b)
    # ---- face lanes (s, n)
    fl = L['face_lanes']   # list of (lane, s, n)
    XFL = X + 3 * NP; XFR = XFL + 16 * G + L['fr_pad']; XLF = XFR + 16 * G; XFF = XLF + 8 * G
    rowb = L['fl_rowbase']   # base of row (var, s) inside a side block
    for f in range(4): A.add('P7a', [(l, NOD + f * NP + face_node(s, n)) for l, s, n in fl])
    for k in range(4):
        A.add('P7a', [(l, LR + (4 + k) * NP + face_node(s, n)) for l, s, n in fl])
        A.add('P7a', [(l, NOD + (4 + k) * NP + face_node(s, n)) for l, s, n in fl])
    for base in (XFL, XFR):
        for var in range(4): A.add('P7a', [(l, base + rowb(var, s) + n) for l, s, n in fl])
    for c in range(2): A.add('P7a', [(l, XLF + (s * 2 + c) * G + n) for l, s, n in fl])
    # ---- P7b: lane l = s*8 + side*4 + var
    for n in range(G): A.add('P7b', [(l, (XFL if not ((l >> 2) & 1) else XFR) + rowb(l & 3, l >> 3) + n) for l in range(32)])
    for i in range(Q): A.add('P7b', [(l, T + (l >> 3) * L['ts7'] + (l & 7) * L['tr7'] + i) for l in range(32)])
    # ---- P7c
    for it0 in (0, 32):
        lanes = [(l, it0 + l) for l in range(32) if it0 + l < 4 * Q]
        for k in range(8): A.add('P7c', [(l, T + (p // Q) * L['ts7'] + k * L['tr7'] + p % Q) for l, p in lanes])
        for k in range(3): A.add('P7c', [(l, XFF + (p // Q) * L['fs'] + k * L['fr'] + p % Q) for l, p in lanes])
    # ---- P7d: lane (s, k) reads ff rows, writes T rows tb(s, k)
    tb = L['tproj']
    for i in range(Q): A.add('P7d', [(l, XFF + (l // 3) * L['fs'] + (l % 3) * L['fr'] + i) for l in range(12)])
    for n in range(G): A.add('P7d', [(l, T + tb(l // 3, l % 3) + n) for l in range(12)])
    # ---- P8: node lanes I = m*G + n
    nodes = [(I, I // G, I % G) for I in range(NP)]
    for f in range(3): A.add('P8', [(I, X + f * NP + I) for I, m, n in nodes])
    for f in range(4): A.add('P8', [(I, LR + f * NP + I) for I, m, n in nodes])
    for ps in range(2):
        lanes = []
        for I, m, n in nodes:
            s = (0 if m == 0 else 1 if m == G - 1 else -1) if ps == 0 else (2 if n == 0 else 3 if n == G - 1 else -1)
            if s >= 0: lanes.append((I, s, n if ps == 0 else m))
        for k in range(3): A.add('P8', [(I, T + tb(s, k) + nf) for I, s, nf in lanes])
        for c in range(2): A.add('P8', [(I, XLF + (s * 2 + c) * G + nf) for I, s, nf in lanes])
    for f in range(3): A.add('P8', [(I, NOD + f * NP + I) for I, m, n in nodes])
    for f in (0, 1, 2, 7, 8): A.add('P8', [(I, NOD + f * NP + I) for I, m, n in nodes])
    # ---- P9 face gathers
    for f in range(3): A.add('P9', [(l, NOD + f * NP + face_node(s, n)) for l, s, n in fl])
    for f in range(4): A.add('P9', [(l, LR + f * NP + face_node(s, n)) for l, s, n in fl])
    return A.d

CUR = dict(gsplit=False, face_lanes=[(l, l // G, l % G) for l in range(4 * G)], fr_pad=0, fl_rowbase=lambda var, s: (var * 4 + s) * G,
           ts7=8 * Q, tr7=Q, fs=3 * Q, fr=Q, tproj=lambda s, k: (s * 3 + k) * G)

def show(name, L):
    d = model(L); t = sum(v[0] for v in d.values()); i = sum(v[1] for v in d.values())
    print(name, 'wavefronts', t, 'ideal', i, 'excess', t - i, {k: v[0] - v[1] for k, v in d.items()})
    return t - i

if __name__ == '__main__':
    show('current', CUR)
    N1 = dict(CUR, gsplit=True, fr_pad=2, ts7=89, tr7=10, fs=41, fr=10)
    show('strides + gradient-lane split', N1)

def search_face_lanes():
    """face lanes: which four (s, n) go to the second half-warp, and the row bases of the trace staging arrays"""
    N1 = dict(CUR, gsplit=True, fr_pad=2, ts7=89, tr7=10, fs=41, fr=10)
    allsn = [(s, n) for s in range(4) for n in range(G)]
    best = []
    for w20, w21, w19 in itertools.product((0, 1), repeat=3):
        hw1 = ([(1, 0), (2, 4)] if w20 else [(0, 4), (3, 0)]) + [(1, 1) if w21 else (2, 1)] + [(3, 3) if w19 else (0, 3)]
        hw0 = [x for x in allsn if x not in hw1]
        fl = [(i, s, n) for i, (s, n) in enumerate(hw0)] + [(16 + i, s, n) for i, (s, n) in enumerate(hw1)]
        for rs in itertools.product(range(16), repeat=3):
            r = (0,) + rs
            if len({(r[s] + n) % 16 for s, n in hw0}) < 16: continue      # stores of the first half-warp conflict-free
            for VS in range(4 * G, 4 * G + 24):
                # rows must not overlap: row (var, s) occupies [var*VS + R[s], +G)
                for lift in itertools.product((0, 16, 32), repeat=4):
                    R = [r[s] + lift[s] for s in range(4)]
                    rows = sorted(R)
                    if any(rows[i + 1] - rows[i] < G for i in range(3)) or rows[3] + G > VS: continue
                    for pad in range(0, 16):
                        L2 = dict(N1, face_lanes=fl, fl_rowbase=lambda var, s, R=R, VS=VS: var * VS + R[s], fr_pad=pad + 4 * VS - 16 * G if 4 * VS >= 16 * G else pad)
                        d = model(L2); e = sum(d[p][0] - d[p][1] for p in ('P7a', 'P7b', 'P9'))
                        if e == 0:
                            return (hw1, R, VS, L2['fr_pad'])
                        best.append((e, hw1, R, VS, pad))
    best.sort(key=lambda x: x[0])
    return best[:3]

if __name__ == '__main__':
    print(search_face_lanes())
