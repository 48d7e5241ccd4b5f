"""CPU model of the packet kNN traversals (work counters only, no timing): how many leaf scans, node expansions and
insertion rounds a 32-query warp performs under (A) the round-1 lockstep packet search and (B) the lane-autonomous
search (leaves collected per packet, every lane scans only the leaves ITS bound still needs).
Used to choose the kernel design before spending GPU time; nothing here is on the product path.

  python tools/sim_knn.py [k] [npackets] [self|cross]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import semantic_icp_b200 as pkg  # noqa: E402

synth = pkg.synth
K = int(sys.argv[1]) if len(sys.argv) > 1 else 4
NP = int(sys.argv[2]) if len(sys.argv) > 2 else 200
MODE = sys.argv[3] if len(sys.argv) > 3 else "cross"
SEED = int(os.environ.get("SEED", "1" if K >= 8 else "0"))
LEAF, AR = 32, 8


def spread10(v):
    v = v & 0x3FF
    v = (v | (v << 16)) & 0x030000FF
    v = (v | (v << 8)) & 0x0300F00F
    v = (v | (v << 4)) & 0x030C30C3
    v = (v | (v << 2)) & 0x09249249
    return v


def morton(p, lo, ext):
    f = np.clip((p - lo) * (1023.0 / ext), 0, 1023).astype(np.uint32)
    return spread10(f[:, 0]) | (spread10(f[:, 1]) << 1) | (spread10(f[:, 2]) << 2)


class Tree:
    def __init__(self, xyz):
        self.lo = xyz.min(0)
        self.ext = float((xyz.max(0) - self.lo).max())
        key = morton(xyz, self.lo, self.ext)
        order = np.argsort(key, kind="stable")
        self.pts = xyz[order]
        self.key = key[order]
        n = len(xyz)
        self.nleaf = (n + LEAF - 1) // LEAF
        pad = self.nleaf * LEAF - n
        P = np.concatenate([self.pts, np.full((pad, 3), np.nan, np.float32)])
        L = P.reshape(self.nleaf, LEAF, 3)
        self.leafpts = L
        lo, hi = [np.nanmin(L, 1)], [np.nanmax(L, 1)]
        while len(lo[-1]) > AR:
            m = len(lo[-1])
            g = (m + AR - 1) // AR
            a = np.full((g * AR, 3), np.inf, np.float32); a[:m] = lo[-1]
            b = np.full((g * AR, 3), -np.inf, np.float32); b[:m] = hi[-1]
            lo.append(a.reshape(g, AR, 3).min(1)); hi.append(b.reshape(g, AR, 3).max(1))
        self.blo, self.bhi = lo, hi
        self.leafkey = self.key[::LEAF]

    def home(self, q):
        k = morton(q, self.lo, self.ext)
        return np.clip(np.searchsorted(self.leafkey, k, side="right") - 1, 0, self.nleaf - 1)


def box_lb(q, lo, hi):  # q [32,3], lo/hi [3] -> [32]
    d = np.maximum(np.maximum(lo - q, q - hi), 0)
    return (d * d).sum(-1)


class Lists:
    """per-lane sorted k-best distances"""

    def __init__(self):
        self.d = np.full((32, K), np.inf, np.float32)

    def worst(self):
        return self.d[:, K - 1]

    def scan(self, q, pts, lanes):
        """lanes [32] bool scan candidate block pts [m,3] (same for all lanes) -> max insertions over lanes"""
        d = ((q[:, None, :] - pts[None, :, :]) ** 2).sum(-1)
        d = np.where(np.isnan(d), np.inf, d)
        return self._merge(d, lanes)

    def scan_own(self, q, leafpts):
        """every lane scans its own block leafpts [32,m,3] (nan rows = lane idle)"""
        d = ((q[:, None, :] - leafpts) ** 2).sum(-1)
        lanes = ~np.isnan(leafpts[:, 0, 0]) | ~np.all(np.isnan(d), 1)
        d = np.where(np.isnan(d), np.inf, d)
        return self._merge(d, np.ones(32, bool))

    def _merge(self, d, lanes):
        w = self.worst()[:, None]
        passed = (d <= w) & lanes[:, None]
        # sequential insertion count: candidate j inserts if it beats the bound as tightened by earlier ones; model with exact sequential loop
        ins = np.zeros(32, int)
        for j in range(d.shape[1]):
            c = d[:, j]
            ok = (c < self.d[:, K - 1]) & passed[:, j]
            if ok.any():
                ins += ok
                nd = np.sort(np.concatenate([self.d, np.where(ok, c, np.inf)[:, None]], 1), 1)[:, :K]
                self.d = nd
        return ins


def sim_A(T, q):
    """round-1 design: seeds scanned by ALL lanes, nearest-first DFS, a leaf is scanned by all lanes if any lane needs it"""
    L = Lists()
    scans = exps = ins_rounds = 0
    home = T.home(q)
    seeds = []
    for h in dict.fromkeys(home.tolist()):
        for dl in range(SEED, -SEED - 1, -1):
            lf = h + dl
            if 0 <= lf < T.nleaf and lf not in seeds:
                seeds.append(lf)
    for lf in seeds[:24]:
        ins = L.scan(q, T.leafpts[lf], np.ones(32, bool)); scans += 1; ins_rounds += ins.max()
    plo, phi = q.min(0), q.max(0)
    top = len(T.blo) - 1
    stack = []

    def expand(level, idx):
        nonlocal exps
        exps += 1
        cl = level - 1
        c0 = idx * AR
        nc = min(AR, len(T.blo[cl]) - c0)
        out = []
        wmax = L.worst().max()
        for j in range(nc):
            lo, hi = T.blo[cl][c0 + j], T.bhi[cl][c0 + j]
            g = np.maximum(np.maximum(lo - phi, plo - hi), 0)
            if (g * g).sum() > wmax:
                continue
            lb = box_lb(q, lo, hi)
            need = lb <= L.worst()
            if need.any():
                out.append((lb[need].min(), cl, c0 + j))
        out.sort(key=lambda e: -e[0])
        stack.extend(out)

    expand(top + 1, 0)
    while stack:
        key, level, idx = stack.pop()
        if key > L.worst().max():
            continue
        if level == 0:
            if idx in seeds:
                continue
            lb = box_lb(q, T.blo[0][idx], T.bhi[0][idx])
            if not (lb <= L.worst()).any():
                continue
            ins = L.scan(q, T.leafpts[idx], np.ones(32, bool)); scans += 1; ins_rounds += ins.max()
        else:
            expand(level, idx)
    return dict(scans=scans, exps=exps, ins_rounds=ins_rounds), L


def sim_B(T, q, cap=32, lane_test=False):
    """lane-autonomous: own home leaf (+-SEED) per lane in lockstep rounds; then DFS collects leaves that pass the PACKET test into
    a pending list (<= cap); flush: every lane tests every entry against ITS bound, then rounds in which each lane scans its own next needed leaf"""
    L = Lists()
    c = dict(rounds=0, lane_scans=0, exps=0, entries=0, flushes=0, ins_rounds=0, lane_tests=0, distinct=0)
    home = T.home(q)
    mine = [set() for _ in range(32)]
    for dl in [0] + [s * d for d in range(1, SEED + 1) for s in (-1, 1)]:
        lf = np.clip(home + dl, 0, T.nleaf - 1)
        blocks = T.leafpts[lf].copy()
        for i in range(32):
            if lf[i] in mine[i]:
                blocks[i] = np.nan
            mine[i].add(int(lf[i]))
        ins = L.scan_own(q, blocks); c["rounds"] += 1; c["lane_scans"] += 32; c["ins_rounds"] += ins.max(); c["distinct"] += len(set(lf.tolist()))
    plo, phi = q.min(0), q.max(0)
    top = len(T.blo) - 1
    stack = [(top + 1, 0)]
    pending = []

    def flush():
        if not pending:
            return
        c["flushes"] += 1; c["entries"] += len(pending)
        need = np.zeros((32, len(pending)), bool)
        for e, lf in enumerate(pending):
            lb = box_lb(q, T.blo[0][lf], T.bhi[0][lf])
            need[:, e] = lb <= L.worst()
            for i in range(32):
                if lf in mine[i]:
                    need[i, e] = False
        c["lane_tests"] += len(pending)
        while need.any():
            blocks = np.full((32, LEAF, 3), np.nan, np.float32)
            chosen = set()
            for i in range(32):
                while need[i].any():
                    e = int(np.argmax(need[i])); need[i, e] = False
                    lf = pending[e]
                    lb = box_lb(q[i:i + 1], T.blo[0][lf], T.bhi[0][lf])[0]
                    if lb <= L.worst()[i]:
                        blocks[i] = T.leafpts[lf]; c["lane_scans"] += 1; chosen.add(lf)
                        break
            if not chosen:
                break
            ins = L.scan_own(q, blocks); c["rounds"] += 1; c["ins_rounds"] += ins.max(); c["distinct"] += len(chosen)
        pending.clear()

    while stack:
        level, idx = stack.pop()
        c["exps"] += 1
        cl = level - 1
        c0 = idx * AR
        nc = min(AR, len(T.blo[cl]) - c0)
        wmax = L.worst().max()
        kids = []
        for j in range(nc):
            lo, hi = T.blo[cl][c0 + j], T.bhi[cl][c0 + j]
            g = np.maximum(np.maximum(lo - phi, plo - hi), 0)
            gd = (g * g).sum()
            if gd > wmax:
                continue
            if lane_test and not (box_lb(q, lo, hi) <= L.worst()).any():
                continue
            kids.append((gd, c0 + j))
        kids.sort(key=lambda e: -e[0])
        for gd, ci in kids:
            if cl == 0:
                pending.append(ci)
                if len(pending) >= cap:
                    flush()
            else:
                stack.append((cl, ci))
    flush()
    return c, L


def main():
    p = synth.kitti_pair(0)
    tgt = p["tgt_xyz"].astype(np.float32)
    src = tgt if MODE == "self" else p["src_xyz"].astype(np.float32)
    T = Tree(tgt)
    Q = Tree(src)
    rng = np.random.default_rng(0)
    picks = rng.choice(Q.nleaf - 1, NP, replace=False)
    accA, accB, accB2 = {}, {}, {}
    for lf in picks:
        q = Q.leafpts[lf]
        if np.isnan(q).any():
            continue
        a, LA = sim_A(T, q)
        b, LB = sim_B(T, q)
        b2, LB2 = sim_B(T, q, lane_test=True)
        assert np.array_equal(LA.d, LB.d) and np.array_equal(LA.d, LB2.d), "designs disagree"
        for acc, r in ((accA, a), (accB, b), (accB2, b2)):
            for k, v in r.items():
                acc.setdefault(k, []).append(v)
    print(f"k={K} mode={MODE} seed={SEED} packets={len(accA['scans'])}")
    for name, acc in (("A lockstep packet", accA), ("B lane-autonomous (packet test only)", accB), ("B2 lane-autonomous (+ any-lane test)", accB2)):
        print(name, {k: round(float(np.mean(v)), 2) for k, v in acc.items()}, "p95", {k: float(np.percentile(v, 95)) for k, v in acc.items() if k in ("scans", "rounds", "entries")})


if __name__ == "__main__":
    main()


def ideal(T, q, L):
    """lower bound for any exact search on these leaf boxes: leaves whose box is within the FINAL k-th distance of the query"""
    w = L.worst()
    lo, hi = T.blo[0], T.bhi[0]
    d = np.maximum(np.maximum(lo[None] - q[:, None], q[:, None] - hi[None]), 0)
    lb = (d * d).sum(-1)
    need = lb <= w[:, None]
    return need.sum(1), need.any(0).sum()
