"""CPU check of the array formulation of DistributeOctTree used by k_quadtree (csrc/orb.cu) against a direct restatement of
the reference's list algorithm (src/ORBextractor.cc:483-765, canonical tie-break: creation sequence).  Random candidate sets."""
import math, random, sys
import numpy as np


def f32(x):
    return float(np.float32(x))


def list_quadtree(cands, W, H, N):
    """cands: list of (x, y, score) with integer coords in [0,W)x[0,H).  Returns kept indices in list order."""
    nIni = max(int(round(f32(np.float32(W) / np.float32(H)))), 1)        # roundf
    # roundf rounds half away from zero; python round is banker's: handle .5
    v = f32(np.float32(W) / np.float32(H)); nIni = max(int(math.floor(v + 0.5)), 1)
    hX = np.float32(W) / np.float32(nIni)
    seq = 0
    nodes = []          # list order; each node: dict
    ini = []
    for i in range(nIni):
        nd = dict(x0=int(hX * np.float32(i)), y0=0, x1=int(hX * np.float32(i + 1)), y1=H, keys=[], seq=seq, nomore=False); seq += 1
        ini.append(nd); nodes.append(nd)
    for i, (x, y, s) in enumerate(cands):
        ini[min(int(np.float32(x) / hX), nIni - 1)]['keys'].append(i)
    nodes = [n for n in nodes if n['keys']]
    for n in nodes:
        if len(n['keys']) == 1: n['nomore'] = True

    def divide(n):
        hx = (n['x1'] - n['x0'] + 1) // 2; hy = (n['y1'] - n['y0'] + 1) // 2
        ch = [dict(x0=n['x0'], y0=n['y0'], x1=n['x0'] + hx, y1=n['y0'] + hy), dict(x0=n['x0'] + hx, y0=n['y0'], x1=n['x1'], y1=n['y0'] + hy),
              dict(x0=n['x0'], y0=n['y0'] + hy, x1=n['x0'] + hx, y1=n['y1']), dict(x0=n['x0'] + hx, y0=n['y0'] + hy, x1=n['x1'], y1=n['y1'])]
        for c in ch: c['keys'] = []; c['nomore'] = False
        for k in n['keys']:
            x, y, _ = cands[k]
            if x < ch[0]['x1']:
                (ch[0] if y < ch[0]['y1'] else ch[2])['keys'].append(k)
            elif y < ch[0]['y1']: ch[1]['keys'].append(k)
            else: ch[3]['keys'].append(k)
        for c in ch:
            if len(c['keys']) == 1: c['nomore'] = True
        return ch
    finish = False
    while not finish:
        prev = len(nodes); nexp = 0; expand = []
        new_front = []
        remaining = []
        for n in list(nodes):
            if n['nomore']: remaining.append(n); continue
            for c in divide(n):
                if not c['keys']: continue
                c['seq'] = seq; seq += 1
                new_front.insert(0, c)
                if len(c['keys']) > 1: nexp += 1; expand.append(c)
        nodes = new_front + remaining
        if len(nodes) >= N or len(nodes) == prev: finish = True
        elif len(nodes) + nexp * 3 > N:
            while not finish:
                prev = len(nodes)
                order = sorted(expand, key=lambda c: (len(c['keys']), c['seq']))
                expand = []
                for j in range(len(order) - 1, -1, -1):
                    par = order[j]
                    for c in divide(par):
                        if not c['keys']: continue
                        c['seq'] = seq; seq += 1
                        nodes.insert(0, c)
                        if len(c['keys']) > 1: expand.append(c)
                    nodes.remove(par)
                    if len(nodes) >= N: break
                if len(nodes) >= N or len(nodes) == prev: finish = True
    out = []
    for n in nodes:
        best = n['keys'][0]
        for k in n['keys'][1:]:
            if cands[k][2] > cands[best][2]: best = k
        out.append(best)
    return out


def array_quadtree(cands, W, H, N):
    """The formulation of k_quadtree: node arrays in list order, processing order proc[], prefix sums."""
    m = len(cands)
    v = f32(np.float32(W) / np.float32(H)); nIni = max(int(math.floor(v + 0.5)), 1)
    hX = np.float32(W) / np.float32(nIni)
    cc = [0] * nIni
    cnode = [0] * m
    for i, (x, y, s) in enumerate(cands):
        ni = min(int(np.float32(x) / hX), nIni - 1); cnode[i] = ni; cc[ni] += 1
    A = []; moved = [-1] * nIni
    for i in range(nIni):
        if cc[i] > 0:
            moved[i] = len(A); A.append([int(hX * np.float32(i)), 0, int(hX * np.float32(i + 1)), H, cc[i], i])
    cnode = [moved[c] for c in cnode]
    seq = nIni; phase = 1; expA = []

    def quad(x, y, nd):
        hx = (nd[2] - nd[0] + 1) >> 1; hy = (nd[3] - nd[1] + 1) >> 1
        return (0 if x < nd[0] + hx else 1) + (0 if y < nd[1] + hy else 2)

    def child(nd, q, cnt, sq):
        hx = (nd[2] - nd[0] + 1) >> 1; hy = (nd[3] - nd[1] + 1) >> 1
        x0 = nd[0] + hx if q & 1 else nd[0]; x1 = nd[2] if q & 1 else nd[0] + hx
        y0 = nd[1] + hy if q & 2 else nd[1]; y1 = nd[3] if q & 2 else nd[1] + hy
        return [x0, y0, x1, y1, cnt, sq]
    while True:
        L = len(A)
        if phase == 1: proc = [p for p in range(L) if A[p][4] >= 2]
        else: proc = [k[2] for k in sorted(expA, key=lambda k: (k[0], k[1]), reverse=True)]
        nP = len(proc)
        if nP == 0: break
        rof = [0] * L
        for r, p in enumerate(proc): rof[p] = r + 1
        cc = [0] * (4 * nP)
        for i, (x, y, s) in enumerate(cands):
            r = rof[cnode[i]]
            if r: cc[4 * (r - 1) + quad(x, y, A[cnode[i]])] += 1
        E = []; run = 0; R = nP - 1
        for r in range(nP):
            k = sum(1 for q in range(4) if cc[4 * r + q] > 0)
            E.append(run); run += k
            if phase == 2 and L + run - (r + 1) >= N and r < R: R = r
        K = E[R] + sum(1 for q in range(4) if cc[4 * R + q] > 0)
        Bn = [None] * (K + L)
        moved = [0] * L; run = 0
        for p in range(L):
            if not (rof[p] and rof[p] - 1 <= R): moved[p] = K + run; Bn[K + run] = A[p]; run += 1
        childpos = [0] * (4 * nP); expB = []
        for r in range(R + 1):
            e = E[r]
            for q in range(4):
                cnt = cc[4 * r + q]
                if cnt <= 0: continue
                pos = K - 1 - e
                Bn[pos] = child(A[proc[r]], q, cnt, seq + e); childpos[4 * r + q] = pos
                if cnt > 1: expB.append((cnt, seq + e, pos))
                e += 1
        for i, (x, y, s) in enumerate(cands):
            p = cnode[i]; r = rof[p]
            cnode[i] = childpos[4 * (r - 1) + quad(x, y, A[p])] if (r and r - 1 <= R) else moved[p]
        Ln = K + (L - (R + 1)); seq += K
        A = Bn[:Ln]; expA = expB
        if Ln >= N or Ln == L: break
        if phase == 1 and Ln + 3 * len(expB) > N: phase = 2
    best = [(-1, 0)] * len(A)
    for i, (x, y, s) in enumerate(cands):
        key = (s, -i)
        if best[cnode[i]] == (-1, 0) or key > best[cnode[i]]: best[cnode[i]] = key
    return [-b[1] for b in best]


if __name__ == "__main__":
    rng = random.Random(1)
    nruns = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    for run in range(nruns):
        W = rng.choice([100, 608, 1248, 1004, 321, 57]); H = rng.choice([88, 448, 688, 568, 200, 40])
        m = rng.choice([1, 2, 5, 40, 300, 1500, 4000]); N = rng.choice([0, 1, 3, 17, 120, 434, 868])
        clustered = rng.random() < 0.4
        cands = []; seen = set()
        while len(cands) < m:
            if clustered: x = min(W - 1, max(0, int(rng.gauss(W / 3, W / 12)))); y = min(H - 1, max(0, int(rng.gauss(H / 2, H / 10))))
            else: x = rng.randrange(W); y = rng.randrange(H)
            if (x, y) in seen and rng.random() < 0.9: continue
            seen.add((x, y)); cands.append((x, y, rng.randrange(7, 60)))
        # reference order: cells row-major, pixels row-major in cell -> any fixed order works for the equivalence
        a = list_quadtree(cands, W, H, N); b = array_quadtree(cands, W, H, N)
        if a != b:
            print("MISMATCH run", run, W, H, m, N, len(a), len(b)); sys.exit(1)
    print("ok", nruns, "runs")
