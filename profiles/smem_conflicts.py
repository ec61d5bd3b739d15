"""Bank-conflict model of the element-record stage kernel's shared-memory accesses (NE = 1: 64-bit words).

A 64-bit warp access is served in half-warps of 16 lanes; a half-warp needs as many wavefronts as the largest number
of DISTINCT words that fall into the same 8-byte bank (word index mod 16).  Used to choose the array strides of
PairRec (stage_pair.cuh) -- prints wavefronts per element for a stride set and searches small paddings."""
import itertools, sys

G, Q = 5, 9
NP, NQ2 = G * G, Q * Q

def wf(addrs):
    """addrs: list of (lane, word).  returns wavefronts for one 64-bit access"""
    tot = 0
    for h in range(2):
        banks = {}
        for lane, w in addrs:
            if lane // 16 == h:
                banks.setdefault(w % 16, set()).add(w)
        if banks:
            tot += max(len(v) for v in banks.values())
    return tot

def face_node(s, n):
    return [n, (G - 1) * G + n, n * G, n * G + G - 1][s]

def count(SN, SX, XJ, ST, TM, SL, FLAY, verbose=False):
    """SN nod field stride, SX X field stride, XJ X row stride, ST T field stride, TM T m-stride, SL L field stride"""
    NOD, X, T, L = 0, 1000, 3000, 5000      # bases (multiples of 16 assumed: only relative words matter mod 16)
    ph = {}
    def add(name, addrs):
        ph[name] = ph.get(name, 0) + wf(addrs)
    # P2 pass 1
    for grp, nl, f0 in ((0, 4 * G, 0), (1, 3 * G, 4)):
        for n in range(G):
            add('P2', [(l, NOD + (f0 + l // G) * SN + (l % G) * G + n) for l in range(nl)])
        for i in range(Q):
            add('P2', [(l, T + (f0 + l // G) * ST + (l % G) * TM + i) for l in range(nl)])
    def gradlines(name, srcbase, srcf, dstf):
        for k in range(G):
            a, b = [], []
            for l in range(4 * G):
                kind, r = l // (2 * G), l % (2 * G); f, ll = r // G, r % G
                off = ll if kind else ll * G; st = G if kind else 1
                a.append((l, srcbase + srcf(kind, f) + off + k * st)); b.append((l, L + dstf(kind, f) * SL + off + k * st))
            add(name, a); add(name, b)
    gradlines('P2g', NOD, lambda kind, f: (7 + f) * SN, lambda kind, f: 2 * kind + f)
    # P3 pass 2
    for it0 in (0, 32):
        lanes = [(l, it0 + l) for l in range(32) if it0 + l < 7 * Q]
        for m in range(G):
            add('P3', [(l, T + (it // Q) * ST + m * TM + it % Q) for l, it in lanes])
        for j in range(Q):
            add('P3', [(l, X + (it // Q) * SX + j * XJ + it % Q) for l, it in lanes])
    # P4 (consecutive; count plain)
    for it0 in (0, 32, 64):
        lanes = [l for l in range(32) if it0 + l < NQ2]
        for f in range(15):
            add('P4', [(l, X + (f % 8) * SX + ((it0 + l) // Q) * XJ + (it0 + l) % Q) for l in lanes])
    # P5: arrays order [Fk1 Fk2 Fk3 Fe1 Fe2 Fe3 S2 S3] -> Fk_f = f, Fe_f = 3+f, S_f = 5+f
    for arrs in ((0, 1, 2), (3, 4, 5), (None, 6, 7)):
        for j in range(Q):
            add('P5', [(l, X + arrs[l // Q] * SX + j * XJ + l % Q) for l in range(3 * Q) if arrs[l // Q] is not None])
    for half in (0, 3):
        for m in range(G):
            add('P5', [(l, T + (half + l // Q) * ST + m * TM + l % Q) for l in range(3 * Q)])
    # P6
    for half in (0, 3):
        for i in range(Q):
            add('P6', [(l, T + (half + l // G) * ST + (l % G) * TM + i) for l in range(3 * G)])
    for n in range(G):
        add('P6', [(l, X + (l // G) * NP + (l % G) * G + n) for l in range(3 * G)])
    gradlines('P6l', L, lambda kind, f: (8 + 2 * kind + f) * SL, lambda kind, f: 2 * kind + f)
    # P7a gathers + stores
    lanes = [(l, l // G, l % G) for l in range(4 * G)]
    for f in range(4):
        add('P7a', [(l, NOD + f * SN + face_node(s, n)) for l, s, n in lanes])
    for k in range(4):
        add('P7a', [(l, L + (4 + k) * SL + face_node(s, n)) for l, s, n in lanes])
        add('P7a', [(l, NOD + (4 + k) * SN + face_node(s, n)) for l, s, n in lanes])
    FLB = X + 3 * NP
    def fl_addr(side, s, var, n):
        return FLB + (side * 16 * G + ((s * 4 + var) * G + n if FLAY == 0 else var * 4 * G + s * G + n))
    for side in range(2):
        for var in range(4):
            add('P7a', [(l, fl_addr(side, s, var, n)) for l, s, n in lanes])
    # P7b
    for n in range(G):
        add('P7b', [(l, fl_addr((l >> 2) & 1, l >> 3, l & 3, n)) for l in range(32)])
    for iq in range(Q):
        add('P7b', [(l, T + l * Q + iq) for l in range(32)])
    # P7c
    for it0 in (0, 32):
        lanes = [(l, it0 + l) for l in range(32) if it0 + l < 4 * Q]
        for k in range(8):
            add('P7c', [(l, T + ((p // Q) * 8 + k) * Q + p % Q) for l, p in lanes])
        for k in range(3):
            add('P7c', [(l, X + 400 + ((p // Q) * 3 + k) * Q + p % Q) for l, p in lanes])
    tot = sum(ph.values())
    if verbose:
        print(ph)
    return tot

if __name__ == "__main__":
    base = count(NP, NQ2, Q, G * Q, Q, NP, 0, verbose=True)
    print("current layout: wavefronts (modelled phases)", base)
    best = []
    for SN, SX, XJ, ST, TM, SL, FLAY in itertools.product(range(25, 30), range(81, 98), (9,), range(45, 62), (9, 10, 11), range(25, 30), (0, 1)):
        if ST < G * TM - (TM - Q):
            continue
        c = count(SN, SX, XJ, ST, TM, SL, FLAY)
        extra = 9 * (SN - 25) + 8 * (SX - 81) + 8 * (ST - 45) + 12 * (SL - 25)
        best.append((c, extra, SN, SX, XJ, ST, TM, SL, FLAY))
    best.sort()
    for b in best[:15]:
        print(b)
    print("best with extra <= 80:")
    for b in [b for b in best if b[1] <= 80][:10]:
        print(b)
