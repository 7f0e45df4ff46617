"""
ORACLE — CPU restatement of Whale.jl's ALE/DLWGD hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  The product (``whale.jl_b200``) never does.

Parity status: PINNED.  The reference (pure Julia) cannot run here (no ``julia`` binary), so this is
a restatement; it is pinned by the reference's own known-answer tests
(``test/runtests.jl:19`` -> -60.96367806571888 and ``test/runtests.jl:32-34`` -> -592.0185620440255),
see ``tests/test_oracle.py``.  Gradients (ForwardDiff.jl 0.10.38, Manifest.toml:630-634) are restated
as forward-mode dual numbers (`Dual` below) and cross-checked with central finite differences;
backtracking (global ``rand()`` in the reference) is restated with an explicit uniform stream.

Everything is written 1-based like the Julia source (index 0 of id-indexed lists is a dummy) so that
each function can be read side by side with the file:line it cites (paths relative to /root/reference).

Third-party semantics restated here (not vendored in the reference tree):
  * NewickTree.jl 0.3.1 (Manifest.toml:1061-1065): ``readnw``, ``getleaves`` (left-to-right),
    ``postwalk``, ``getlca`` and ``insertnode!``.  ``insertnode!(n; name)`` places the new node halfway
    along n's branch and *appends* it as the LAST child of n's former parent (delete!+push!).  The
    append behaviour is pinned by the reference's own test ``test/runtests.jl:54-55``:
    ``insertnode!(t[1][1]); insertnode!(t[1][2][1])`` only works if, after the first call, ``t[1][2]``
    is the freshly inserted (non-leaf) node.
  * ``Base.isapprox(a, b, atol=1e-6)`` with rtol=0  =>  |a-b| <= 1e-6 on the *values* of duals.
"""
from __future__ import annotations

import math
import os
import re
from dataclasses import dataclass, field

import numpy as np

NaN = float("nan")
LMATOL = 1e-6  # src/bdputil.jl:3


# --------------------------------------------------------------------------------------------
# forward-mode dual numbers (ForwardDiff.Dual restated); value `v`, partials `d` (numpy vector)
# --------------------------------------------------------------------------------------------
class Dual:
    __slots__ = ("v", "d")

    def __init__(self, v, d):
        self.v = float(v)
        self.d = d

    @staticmethod
    def _lift(x, like):
        return x if isinstance(x, Dual) else Dual(x, np.zeros_like(like.d))

    def __add__(self, o):
        if isinstance(o, Dual):
            return Dual(self.v + o.v, self.d + o.d)
        return Dual(self.v + o, self.d)

    __radd__ = __add__

    def __neg__(self):
        return Dual(-self.v, -self.d)

    def __sub__(self, o):
        if isinstance(o, Dual):
            return Dual(self.v - o.v, self.d - o.d)
        return Dual(self.v - o, self.d)

    def __rsub__(self, o):
        return Dual(o - self.v, -self.d)

    def __mul__(self, o):
        if isinstance(o, Dual):
            return Dual(self.v * o.v, self.d * o.v + self.v * o.d)
        return Dual(self.v * o, self.d * o)

    __rmul__ = __mul__

    def __truediv__(self, o):
        if isinstance(o, Dual):
            q = self.v / o.v
            return Dual(q, (self.d - q * o.d) / o.v)
        return Dual(self.v / o, self.d / o)

    def __rtruediv__(self, o):
        q = o / self.v
        return Dual(q, (-q / self.v) * self.d)

    def __pow__(self, n):
        assert isinstance(n, int)
        return Dual(self.v ** n, (n * self.v ** (n - 1)) * self.d)

    def __lt__(self, o):
        return self.v < (o.v if isinstance(o, Dual) else o)

    def __gt__(self, o):
        return self.v > (o.v if isinstance(o, Dual) else o)

    def __le__(self, o):
        return self.v <= (o.v if isinstance(o, Dual) else o)

    def __ge__(self, o):
        return self.v >= (o.v if isinstance(o, Dual) else o)

    def __repr__(self):
        return f"Dual({self.v}, {self.d})"


def value(x):
    return x.v if isinstance(x, Dual) else x


def _exp(x):
    if isinstance(x, Dual):
        e = math.exp(x.v)
        return Dual(e, e * x.d)
    return math.exp(x)


def _log(x):
    if isinstance(x, Dual):
        return Dual(math.log(x.v), x.d / x.v)
    return math.log(x)


def _isfinite(x):
    return math.isfinite(value(x))


# --------------------------------------------------------------------------------------------
# NewickTree.jl restated (structure only)
# --------------------------------------------------------------------------------------------
class TNode:
    """A Newick tree node: name, distance to parent (NaN for the root), ordered children."""

    def __init__(self, name="", dist=NaN):
        self.name = name
        self.dist = dist
        self.children: list[TNode] = []
        self.parent: TNode | None = None

    def isleaf(self):
        return not self.children

    def isroot(self):
        return self.parent is None

    def push(self, c):
        c.parent = self
        self.children.append(c)

    def __getitem__(self, i):  # 1-based child access like NewickTree's n[i]
        return self.children[i - 1]


def readnw(s: str) -> TNode:
    """Parse a Newick string (names, branch lengths; support values ignored)."""
    s = s.strip()
    assert s.endswith(";"), "newick string must end with ';'"
    pos = 0

    def parse() -> TNode:
        nonlocal pos
        node = TNode()
        if s[pos] == "(":
            pos += 1
            while True:
                node.push(parse())
                if s[pos] == ",":
                    pos += 1
                    continue
                assert s[pos] == ")"
                pos += 1
                break
        m = re.match(r"[^,:();]*", s[pos:])
        label = m.group(0)
        pos += len(label)
        node.name = label.strip() if node.isleaf() else ""  # internal labels = support: ignored
        if node.children and label.strip().startswith("wgd"):
            node.name = label.strip()
        if s[pos] == ":":
            pos += 1
            m = re.match(r"[^,();]*", s[pos:])
            node.dist = float(m.group(0))
            pos += len(m.group(0))
        return node

    root = parse()
    assert s[pos] == ";"
    return root


def getleaves(n: TNode) -> list[TNode]:
    if n.isleaf():
        return [n]
    out = []
    for c in n.children:
        out.extend(getleaves(c))
    return out


def postwalk(n: TNode) -> list[TNode]:
    out = []
    for c in n.children:
        out.extend(postwalk(c))
    out.append(n)
    return out


def prewalk(n: TNode) -> list[TNode]:
    out = [n]
    for c in n.children:
        out.extend(prewalk(c))
    return out


def getroot(n: TNode) -> TNode:
    while n.parent is not None:
        n = n.parent
    return n


def getlca(t: TNode, a: str, b: str) -> TNode:
    """Last common ancestor of the leaves named a and b (a == b: the leaf itself)."""
    leaves = {l.name: l for l in getleaves(t)}
    x = leaves[a]
    anc = []
    while x is not None:
        anc.append(x)
        x = x.parent
    y = leaves[b]
    while y not in anc:
        y = y.parent
    return y


def insertnode(n: TNode, name="") -> TNode:
    """NewickTree.insertnode!(n; name): new node halfway on n's branch; appended as the LAST child
    of n's former parent (pinned by test/runtests.jl:54-55, see module docstring)."""
    p = n.parent
    assert p is not None
    half = n.dist / 2
    m = TNode(name=name, dist=n.dist - half)
    n.dist = half
    p.children.remove(n)
    p.push(m)
    m.push(n)
    return m


def nwstr(n: TNode) -> str:
    if n.isleaf():
        return n.name
    return "(" + ",".join(nwstr(c) for c in n.children) + ")" + (n.name if n.name else "")


# src/Whale.jl:34-37
EXTREE = ("((MPOL:4.752,PPAT:4.752):0.292,(SMOE:4.457,(((OSAT:1.555,(ATHA:0.5548,CPAP:0.5548):1.0002):0"
          ".738,ATRI:2.293):1.225,(GBIL:3.178,PABI:3.178):0.34):0.939):0.587);")


# --------------------------------------------------------------------------------------------
# rate models (src/rmodels.jl)
# --------------------------------------------------------------------------------------------
@dataclass
class ConstantDLWGD:  # src/rmodels.jl:23-29
    lam: object
    mu: object
    q: list = field(default_factory=list)
    p: list = field(default_factory=list)
    eta: object = 0.66


@dataclass
class DLWGD:  # src/rmodels.jl:47-53 (lam, mu on log scale)
    lam: list
    mu: list
    q: list = field(default_factory=list)
    p: list = field(default_factory=list)
    eta: object = 0.66


# --------------------------------------------------------------------------------------------
# WhaleModel structure (src/model.jl)
# --------------------------------------------------------------------------------------------
class MNode:
    """ModelNode (src/model.jl:6-14): slices matrix rows 1..n+1, cols [dt, eps, phi, psi]."""

    def __init__(self, name, dist, dt, minn, maxn, wgdid, leafP):
        self.name = name
        self.dist = dist
        # src/model.jl:16-22
        n = 0 if math.isnan(dist) else min(maxn, max(minn, math.ceil(dist / dt)))
        self.n = n
        self.dts = [0.0] + [0.0 if n == 0 else dist / n] * n  # column 1
        self.eps = [1.0] * (n + 1)
        self.phi = [1.0] * (n + 1)
        self.psi = [1.0] * (n + 1)
        self.wgdid = wgdid
        self.leafP = leafP
        self.id = 0
        self.children: list[MNode] = []
        self.parent: MNode | None = None
        self.clade: set[int] = set()

    def isleaf(self):
        return not self.children

    def isroot(self):
        return self.parent is None

    def iswgd(self):  # src/model.jl:54
        return self.name.startswith("wgd")

    def __len__(self):  # src/model.jl:53  (number of rows)
        return self.n + 1


def nonwgdchild(n: MNode) -> MNode:  # src/model.jl:200-203
    while n.iswgd():
        n = n.children[0]
    return n


class WhaleModel:
    """src/model.jl:71-145.  `order`: leaves (left-to-right) then the remaining nodes in postorder;
    `index[id]` = 1-based position in order."""

    def __init__(self, rates, tree: TNode, dt, minn=5, maxn=10000, condition="root"):
        self.rates = rates
        self.condition = condition
        self.dt, self.minn, self.maxn = dt, minn, maxn
        leafnames = [l.name for l in getleaves(tree)]
        counts = {}
        for nm in leafnames:
            counts[nm] = counts.get(nm, 0) + 1
        mulgroups = {k: v for k, v in counts.items() if v > 1}  # src/model.jl:101
        nonwgd = 0
        wgdid = 0
        order: list[MNode] = []

        def walk(x: TNode, y):  # src/model.jl:102-120 (wgdid in PREORDER, push in POSTORDER)
            nonlocal nonwgd, wgdid
            if x.name.startswith("wgd"):
                wgdid += 1
                i = wgdid
            else:
                nonwgd += 1
                i = 0
            P = 1.0 / mulgroups[x.name] if x.name in mulgroups else (1.0 if x.isleaf() else 0.0)
            y2 = MNode(x.name, x.dist, dt, minn, maxn, i, P)
            if y is not None:
                y2.parent = y
                y.children.append(y2)
            for c in x.children:
                walk(c, y2)
            order.append(y2)
            return y2

        root = walk(tree, None)

        def mleaves(n):
            if n.isleaf():
                return [n]
            return [l for c in n.children for l in mleaves(c)]

        leaves = mleaves(root)
        order = leaves + [n for n in order if not n.isleaf()]  # src/model.jl:124 (union)
        i = nonwgd + 1
        j = 1
        self.index = [0] * (len(order) + 1)
        mulid = {}
        for k, n in enumerate(order, start=1):  # src/model.jl:127-139
            if n.iswgd():
                n.id = i
                i += 1
            else:
                n.id = j
                if n.name in mulgroups:
                    mulid[n.name] = n.id
                j += 1
            self.index[n.id] = k
        self.order = order
        self.nwgd = wgdid

        def setclade(n):  # src/model.jl:32-42
            if n.isleaf():
                n.clade |= {mulid[n.name]} if n.name in mulid else {n.id}
            else:
                for c in n.children:
                    setclade(c)
                    n.clade |= c.clade

        setclade(order[-1])
        setmodel(self)

    def __getitem__(self, i) -> MNode:  # src/model.jl:79
        return self.order[self.index[i] - 1]

    def __len__(self):
        return len(self.order)

    @property
    def root(self) -> MNode:
        return self.order[-1]

    def with_rates(self, rates) -> "WhaleModel":
        """(m::WhaleModel)(rates) src/model.jl:148-160: same structure, new parameters."""
        import copy
        m = copy.copy(self)
        m.rates = rates
        # fresh slice storage, shared structure
        old2new = {}
        new_order = []
        for n in self.order:
            c = copy.copy(n)
            c.eps = [1.0] * len(n)
            c.phi = [1.0] * len(n)
            c.psi = [1.0] * len(n)
            old2new[id(n)] = c
            new_order.append(c)
        for n, c in zip(self.order, new_order):
            c.children = [old2new[id(x)] for x in n.children]
            c.parent = old2new[id(n.parent)] if n.parent is not None else None
        m.order = new_order
        setmodel(m)
        return m

    def show(self) -> str:  # src/model.jl:214-228 (structure lines only)
        lines = ["node_id,wgd_id,distance,dt,n,subtree"]
        for n in self.order:
            lines.append(f"{n.id},{n.wgdid},{n.dist if math.isnan(n.dist) else round(n.dist,4)},"
                         f"{round(n.dts[-1],4)},{n.n},\"{_mnwstr(n)};\"")
        return "\n".join(lines)


def _mnwstr(n: MNode) -> str:
    if n.isleaf():
        return n.name
    return "(" + ",".join(_mnwstr(c) for c in n.children) + ")"


def getp(m, n: MNode):  # src/rmodels.jl:14
    return m.p[n.id - 1] if len(m.p) > 0 and n.isleaf() else 0.0


def gettheta(m, n: MNode):
    """getθ: src/rmodels.jl:31-33 (ConstantDLWGD), :55-64 (DLWGD)."""
    if isinstance(m, ConstantDLWGD):
        return dict(lam=m.lam, mu=m.mu, eta=m.eta, p=getp(m, n),
                    q=m.q[n.wgdid - 1] if n.iswgd() else NaN)
    if n.iswgd():
        c = nonwgdchild(n)
        return dict(lam=_exp(m.lam[c.id - 1]), mu=_exp(m.mu[c.id - 1]), q=m.q[n.wgdid - 1])
    if n.id > len(m.lam):
        return dict(lam=NaN, mu=NaN, p=getp(m, n), eta=m.eta)
    return dict(lam=_exp(m.lam[n.id - 1]), mu=_exp(m.mu[n.id - 1]), p=getp(m, n), eta=m.eta)


# src/bdputil.jl:6-11
def getalpha(lam, mu, t):
    if abs(value(lam) - value(mu)) <= LMATOL:  # isapprox(λ, μ, atol=ΛMATOL) on values
        return lam * t / (1.0 + lam * t)
    e = _exp(t * (lam - mu))
    return mu * (e - 1.0) / (lam * e - mu)


def _eps(a, b, e):
    return (a + (1.0 - a - b) * e) / (1.0 - b * e)


def _phi(a, b, e):
    return (1.0 - a) * (1.0 - b) / (1.0 - b * e) ** 2


def _psi(a, b, e):
    return (1.0 - a) * (1.0 - b) * b / (1.0 - b * e) ** 3


def setslices(n: MNode, lam, mu):  # src/model.jl:182-191
    for i in range(1, len(n)):  # rows 2..n+1 (0-based i = row-1)
        a = getalpha(lam, mu, n.dts[i])
        b = (lam / mu) * a
        e = n.eps[i - 1]
        n.eps[i] = _eps(a, b, e)
        n.phi[i] = _phi(a, b, e)
        n.psi[i] = _psi(a, b, e)


def setmodel(model: WhaleModel):  # src/model.jl:162-180
    for n in model.order:
        th = gettheta(model.rates, n)
        if n.iswgd():
            e = n.children[0].eps[-1]
            n.eps[0] = th["q"] * e ** 2 + (1.0 - th["q"]) * e
        else:
            if n.isleaf():
                n.eps[0] = th["p"]
            else:
                pr = 1.0
                for c in n.children:
                    pr = pr * c.eps[-1]
                n.eps[0] = pr
            n.phi[0] = 1.0
        setslices(n, th["lam"], th["mu"])


# --------------------------------------------------------------------------------------------
# conditioning (src/condition.jl, src/bdputil.jl:67)
# --------------------------------------------------------------------------------------------
def geompgf(p, s):
    return p * s / (1.0 - (1.0 - p) * s)


def condition(wm: WhaleModel, kind=None):
    kind = kind or wm.condition
    if kind == "none":  # NoCondition src/condition.jl:12
        return 0.0
    eta = gettheta(wm.rates, wm.root)["eta"]
    if kind == "nonextinct":  # src/condition.jl:15-18
        return _log(1.0 - geompgf(eta, wm.root.eps[-1]))
    if kind == "root":  # src/condition.jl:21-29
        f, g = wm.root.children
        er = geompgf(eta, wm.root.eps[-1])
        ef = geompgf(eta, f.eps[-1])
        eg = geompgf(eta, g.eps[-1])
        p = 1.0 - ef - eg + er
        return _log(p) if p > 0.0 else -math.inf
    if kind == "nowhere":  # NowhereExtinctCondition src/condition.jl:5-9,31-36
        p = treepgf_allbinary(wm)
        s = treepgf_allbinary_sign(wm)
        c = 1.0
        for pi, si in zip(p[:-1], s[:-1]):  # sum((p .* e.s)[1:end-1])
            c = c - pi * si
        return _log(c) if 0.0 < c < 1.0 else -math.inf
    raise ValueError(kind)


def bdp_pgf(lam, mu, t, s):
    """pgf(::LinearBDP, s) src/bdputil.jl:58-64 (ρ = exp((μ−λ)t), :42)."""
    if abs(value(lam) - value(mu)) <= LMATOL:
        return (1.0 - (lam * t - 1.0) * (s - 1.0)) / (1.0 - lam * t * (s - 1.0))
    rho = _exp((mu - lam) * t)
    return (rho * (lam * s - mu) - mu * (s - 1.0)) / (rho * (lam * s - mu) - lam * (s - 1.0))


def wgdpgf(q, s):  # src/bdputil.jl:73
    return s * (1.0 - q + s * q)


def treepgf_allbinary(wm: WhaleModel):
    """src/bdputil.jl:109-121: the tree pgf at every binary argument f(0..0), f(1,0..0), …, f(1..1); children's
    vectors combine as vec(prod.(Iterators.product(down...))) — the FIRST child's index runs fastest."""
    def walk(n: MNode):
        th = gettheta(wm.rates, n)
        if n is wm.root:
            f = lambda x: geompgf(th["eta"], x)
        else:
            f = lambda x: bdp_pgf(th["lam"], th["mu"], n.dist, x)
        if n.isleaf():
            return [f(0.0), f(1.0)]
        down = [walk(c) for c in n.children]
        xs = down[0]
        for d in down[1:]:
            xs = [a * b for b in d for a in xs]
        return [f(wgdpgf(th["q"], x)) if n.iswgd() else f(x) for x in xs]
    return walk(wm.root)


def treepgf_allbinary_sign(wm: WhaleModel):
    """src/bdputil.jl:125-133."""
    def walk(n: MNode):
        if n.isleaf():
            return [1.0, -1.0]
        down = [walk(c) for c in n.children]
        xs = down[0]
        for d in down[1:]:
            xs = [a * b for b in d for a in xs]
        return xs
    return [math.copysign(1.0, v) for v in walk(wm.root)]


# --------------------------------------------------------------------------------------------
# .ale parser and CCD (src/ccd.jl)
# --------------------------------------------------------------------------------------------
def _tryparse(x):
    try:
        return int(x)
    except ValueError:
        try:
            return float(x)
        except ValueError:
            return x


def parse_aleobserve(fname):
    """src/ccd.jl:147-248 (parse_aleobserve, parse_body, addleafclades!, addubiquitous!)."""
    with open(fname) as fh:
        s = "\n".join(fh.read().splitlines())
    sections = s.split("#")[1:-1]
    assert len(sections) == 8, f"Not a valid .ale file {fname}"
    d = {}
    for sec in sections:
        sec = sec.replace(":\t", "")
        x = sec.split("\n")
        header = x[0].replace("-", "_")
        xs = [l for l in x[1:] if l != ""]
        if header in ("constructor_string", "observations", "last_leafset_id"):
            d[header] = _tryparse(xs[0])  # single-line sections (parse_body :168-169)
        elif header == "Dip_counts":
            dd = {}
            for l in xs:
                y = l.split()
                dd.setdefault(int(y[0]), []).append((int(y[1]), int(y[2]), _tryparse(y[3])))
            d[header] = dd
        elif header == "set_id":
            d[header] = {int(l.split()[0]): [int(t) for t in l.split()[1:]] for l in xs}
        elif header == "leaf_id":
            d[header] = {l.split()[0]: int(l.split()[1]) for l in xs}
        else:
            d[header] = {int(l.split()[0]): _tryparse(l.split()[1]) for l in xs}
    d["leaf_id"] = {v: k for k, v in d["leaf_id"].items()}  # invert: leaf id -> name
    d.setdefault("Dip_counts", {})
    # addleafclades! src/ccd.jl:199-216
    leafclades = {}
    themap = {}
    for k in sorted(d["set_id"]):
        v = d["set_id"][k]
        if len(v) == 1:
            d["Bip_counts"][k] = d["observations"]
            d["Dip_counts"][k] = []
            leafclades[k] = d["leaf_id"][v[0]]
            themap[v[0]] = k
        else:
            d["set_id"][k] = [themap[i] for i in v]
    for k, v in themap.items():
        d["set_id"][v] = [v]
    d["leaf_id"] = leafclades
    # addubiquitous! src/ccd.jl:219-248
    n = len(d["leaf_id"])
    l = d["set_id"]
    nc = len(l)
    G = nc + 1
    d["Dip_counts"][G] = []
    N = 0
    lsets = {k: set(v) for k, v in l.items()}
    for i in range(1, nc + 1):
        for j in range(i + 1, nc + 1):
            if len(l[i]) + len(l[j]) != n:
                continue
            if not (lsets[i] & lsets[j]):
                assert d["Bip_counts"][i] == d["Bip_counts"][j]
                N += d["Bip_counts"][i]
                d["Dip_counts"][G].append((i, j, d["Bip_counts"][j]))
    t = d["Dip_counts"][G][-1]
    d["Bip_counts"][G] = N
    d["set_id"][G] = sorted(set(l[t[0]]) | set(l[t[1]]))
    d["Bip_bls"][G] = 0.0
    d["fname"] = os.path.basename(fname)
    return d


@dataclass
class Clade:  # src/ccd.jl:25-33
    id: int
    count: int
    splits: list  # [(g1, g2, p)] new ids, file order
    leaves: frozenset
    species: frozenset

    def isleaf(self):
        return len(self.leaves) == 1


class CCD:
    """src/ccd.jl:80-121.  clades[γ] (1-based; clades[0] dummy) sorted by (size, old id);
    compat[e] ascending clade ids; index[γ][e] = 1-based column in ℓ[e] or 0."""

    def __init__(self, ale: dict, wm: WhaleModel, spmap: dict):
        set_id = ale["set_id"]
        order = sorted((len(v), k) for k, v in set_id.items())
        idmap = {k: i for i, (_, k) in enumerate(order, start=1)}
        self.clades = [None]
        for i, (_, k) in enumerate(order, start=1):
            v = ale["Bip_counts"][k]
            spl = [(idmap[t[0]], idmap[t[1]], t[2] / v) for t in ale["Dip_counts"].get(k, [])]
            sp = frozenset(spmap[ale["leaf_id"][g].split("_")[0]] for g in set_id[k])
            self.clades.append(Clade(i, v, spl, frozenset(set_id[k]), sp))
        self.leaves = [ale["leaf_id"][k] for k in sorted(ale["leaf_id"])]
        self.total = ale["observations"]
        self.fname = ale["fname"]
        # index_and_getℓ src/ccd.jl:59-75
        G = len(self.clades) - 1
        nn = len(wm)
        self.index = [[0] * (nn + 1) for _ in range(G + 1)]
        self.compat = [[] for _ in range(nn + 1)]
        for n in wm.order:
            i = 1
            for c in self.clades[1:]:
                if not (c.species <= n.clade):  # iscompatible src/ccd.jl:39
                    continue
                self.compat[n.id].append(c.id)
                self.index[c.id][n.id] = i
                i += 1
        self.ell = None  # filled by logpdf_inplace: ell[e][row][col], 0-based rows/cols

    def __len__(self):
        return len(self.clades) - 1


def read_ale(path: str, wm: WhaleModel) -> list[CCD]:  # src/ccd.jl:126-137
    spmap = {}
    for l in [n for n in wm.order if n.isleaf()]:
        spmap[l.name] = l.id
    if os.path.isfile(path) and path.endswith(".ale"):
        return [CCD(parse_aleobserve(path), wm, spmap)]
    if os.path.isfile(path):
        fs = [l.strip() for l in open(path)]
    else:
        fs = [os.path.join(path, f) for f in sorted(os.listdir(path))]  # readdir is sorted
    fs = [f for f in fs if not f.startswith("#")]
    return [CCD(parse_aleobserve(f), wm, spmap) for f in fs]


# --------------------------------------------------------------------------------------------
# the DP (src/core.jl)
# --------------------------------------------------------------------------------------------
def _getl(x: CCD, ell, e, g, row=None):  # src/ccd.jl:41-49  (row 1-based; None = last)
    i = x.index[g][e]
    if i == 0:
        return 0.0
    return ell[e][-1][i - 1] if row is None else ell[e][row - 1][i - 1]


def _alloc(x: CCD, wm: WhaleModel):
    ell = [None] * (len(wm) + 1)
    for n in wm.order:
        ell[n.id] = [[0.0] * len(x.compat[n.id]) for _ in range(len(n))]
    return ell


def _Pspeciation(x, c: Clade, ell, n: MNode):  # src/core.jl:160-170
    f, g = n.children[0].id, n.children[1].id
    p = 0.0
    for (g1, g2, pr) in c.splits:
        p = p + pr * (_getl(x, ell, f, g1) * _getl(x, ell, g, g2) +
                      _getl(x, ell, g, g1) * _getl(x, ell, f, g2))
    return p


def _Ploss(x, c: Clade, ell, n: MNode):  # src/core.jl:172-176
    f, g = n.children
    return _getl(x, ell, f.id, c.id) * g.eps[-1] + _getl(x, ell, g.id, c.id) * f.eps[-1]


def _within_branch(n: MNode, c: Clade, ell, x, e, j, leaf):  # src/core.jl:121-128,178-185
    L = ell[e]
    for i in range(2, len(n) + 1):
        L[i - 1][j - 1] = L[i - 1][j - 1] + n.phi[i - 1] * L[i - 2][j - 1]
        if not leaf:
            p = 0.0
            for (g1, g2, pr) in c.splits:
                p = p + pr * _getl(x, ell, e, g1, i - 1) * _getl(x, ell, e, g2, i - 1)
            L[i - 1][j - 1] = L[i - 1][j - 1] + n.psi[i - 1] * p


def _whale(n: MNode, ell, x: CCD, wm: WhaleModel):
    e = n.id
    if n.iswgd():  # whalewgd! src/core.jl:103-119
        q = gettheta(wm.rates, n)["q"]
        f = n.children[0]
        for c_id in x.compat[e]:
            c = x.clades[c_id]
            j = x.index[c_id][e]
            p = 0.0
            leaf = c.isleaf()
            if not leaf:
                s = 0.0
                for (g1, g2, pr) in c.splits:  # Πwgdretention :187-194
                    s = s + pr * _getl(x, ell, f.id, g1) * _getl(x, ell, f.id, g2)
                p = p + s * q
            # Πwgdloss :196-199
            p = p + ((1.0 - q) * _getl(x, ell, f.id, c_id) + 2.0 * q * f.eps[-1] * _getl(x, ell, f.id, c_id))
            ell[e][0][j - 1] = p
            _within_branch(n, c, ell, x, e, j, leaf)
        return
    if n.isroot():  # whaleroot! src/core.jl:130-149
        eta = gettheta(wm.rates, n)["eta"]
        eps = n.eps[-1]
        xi = 1.0 - (1.0 - eta) * eps
        for c in x.clades[1:]:
            leaf = c.isleaf()
            a = 0.0
            b = 0.0
            cc = _Ploss(x, c, ell, n)
            if not leaf:
                for (g1, g2, pr) in c.splits:  # Πroot :151-158
                    a = a + pr * _getl(x, ell, e, g1, 1) * _getl(x, ell, e, g2, 1)
                b = b + _Pspeciation(x, c, ell, n)
            ell[e][0][c.id - 1] = (1.0 - eta) * xi * a / eta + (eta * (1.0 - eps) / xi ** 2) * (b + cc)
        return
    # whale! src/core.jl:83-101
    for c_id in x.compat[e]:
        c = x.clades[c_id]
        j = x.index[c_id][e]
        leaf = c.isleaf()
        if leaf and n.isleaf():
            ell[e][0][j - 1] = n.leafP
        elif not n.isleaf():
            ell[e][0][j - 1] = ell[e][0][j - 1] + (_Pspeciation(x, c, ell, n) + _Ploss(x, c, ell, n))
        _within_branch(n, c, ell, x, e, j, leaf)


def logpdf_single(wm: WhaleModel, x: CCD, keep=False):
    """logpdf(wm, x::CCD) src/core.jl:31-43 (fresh ℓ; `keep` stores it in x.ell like logpdf!)."""
    ell = _alloc(x, wm)
    for n in wm.order:
        _whale(n, ell, x, wm)
    if keep:
        x.ell = ell
    L = ell[wm.root.id][0][-1]
    return _log(L) if L > 0.0 else -math.inf


def logpdf(wm: WhaleModel, xs, keep=False):
    """src/core.jl:46-64: Σ_i ℓ_i − N·condition(wm), through ℓhood (:15)."""
    if isinstance(xs, CCD):
        return logpdf_single(wm, xs, keep)
    tot = 0.0
    for x in xs:
        tot = tot + logpdf_single(wm, x, keep)
    tot = tot - len(xs) * condition(wm)
    return tot if _isfinite(tot) else -math.inf


# --------------------------------------------------------------------------------------------
# raw-parameter vector <-> rates (the gradient contract of SURVEY §8b)
# --------------------------------------------------------------------------------------------
def rates_from_vector(template, x):
    """x = [lam..., mu..., q..., eta] in the scale stored by the rates struct."""
    if isinstance(template, ConstantDLWGD):
        nq = len(template.q)
        return ConstantDLWGD(lam=x[0], mu=x[1], q=list(x[2:2 + nq]), p=template.p, eta=x[2 + nq])
    n = len(template.lam)
    nq = len(template.q)
    return DLWGD(lam=list(x[:n]), mu=list(x[n:2 * n]), q=list(x[2 * n:2 * n + nq]), p=template.p,
                 eta=x[2 * n + nq])


def vector_from_rates(r):
    if isinstance(r, ConstantDLWGD):
        return [r.lam, r.mu] + list(r.q) + [r.eta]
    return list(r.lam) + list(r.mu) + list(r.q) + [r.eta]


def logpdf_and_gradient(wm: WhaleModel, xs, x0=None):
    """ForwardDiff.gradient(x -> logpdf(wm(rates(x)), xs), x0) restated (test/runtests.jl:36-38)."""
    x0 = [float(v) for v in (x0 if x0 is not None else vector_from_rates(wm.rates))]
    P = len(x0)
    duals = [Dual(v, np.eye(P)[i]) for i, v in enumerate(x0)]
    m = wm.with_rates(rates_from_vector(wm.rates, duals))
    out = logpdf(m, xs)
    if isinstance(out, Dual):
        return out.v, out.d.copy()
    return out, np.zeros(P)


# --------------------------------------------------------------------------------------------
# stochastic backtracking (src/track.jl:123-414) with an explicit uniform stream
# --------------------------------------------------------------------------------------------
class BacktrackFailed(Exception):
    pass


def backtrack(wm: WhaleModel, x: CCD, uniforms):
    """backtrack(wm, ccd) src/track.jl:190-194 on the ℓ left in x.ell by logpdf(..., keep=True).

    Returns (nodes, n_uniforms_used); nodes = list of (γ, e, t, parent_index) in creation (DFS)
    order, parent_index = -1 for the root; loss nodes have γ = 0, t = 0 (src/track.jl:132).
    A uniform is consumed exactly where the reference calls rand() (:217, :274)."""
    ell = x.ell
    u = iter(uniforms)
    used = 0
    nodes = []

    def draw():
        nonlocal used
        used += 1
        return float(next(u))

    G = len(x)
    root = wm.root
    nodes.append((G, root.id, 1, -1))  # BackTracker(model, ccd) :157-160

    def step(state, node):  # backtrack!(b) :205-210 ; state = (e, γ, t)
        stack = [(state, node)]
        # explicit DFS stack, children pushed in reverse so they are processed in listed order
        while stack:
            (e, g, t), node = stack.pop()
            # b(newstate) :162-170 — new tree node whenever (γ, e) changes
            pg, pe = nodes[node][0], nodes[node][1]
            if g != pg or e != pe:
                nodes.append((g, e, t, node))
                node = len(nodes) - 1
            if g == 0:  # loss node :206-207
                continue
            n = wm[e]
            if t == 1:  # _backtrack!(b, n) :213-225
                if n.isleaf():
                    continue
                r = draw() * _getl(x, ell, e, g, t)
                nxt = (_bt_root(r, wm, x, ell, n, e, g, t) if n.isroot() else
                       _bt_wgd(r, wm, x, ell, n, e, g) if n.iswgd() else
                       _bt_internal(r, x, ell, n, e, g))
            else:  # _backtrack!(b) :268-281
                if x.clades[g].isleaf():
                    nxt = [(e, g, 1)]
                else:
                    r = draw() * _getl(x, ell, e, g, t)
                    r -= n.phi[t - 1] * _getl(x, ell, e, g, t - 1)
                    if r < 0.0:
                        nxt = [(e, g, t - 1)]
                    else:
                        nxt = None
                        for (g1, g2, p) in x.clades[g].splits:  # duplication :286-302
                            r -= p * _getl(x, ell, e, g1, t - 1) * _getl(x, ell, e, g2, t - 1) * n.psi[t - 1]
                            if r < 0.0:
                                nxt = [(e, g1, t - 1), (e, g2, t - 1)]
                                break
                        if nxt is None:
                            raise BacktrackFailed(f"r={r} at {(e, g, t)}")
            for s in reversed(nxt):
                stack.append((s, node))

    step((root.id, G, 1), 0)
    return nodes, used


def _bt_internal(r, x, ell, n, e, g):  # :234-242 with sploss :326-340, speciation :304-324
    f, h = n.children
    tf, th = len(f), len(h)
    r -= _getl(x, ell, f.id, g) * h.eps[-1]
    if r < 0.0:
        return [(f.id, g, tf), (h.id, 0, 0)]
    r -= _getl(x, ell, h.id, g) * f.eps[-1]
    if r < 0.0:
        return [(h.id, g, th), (f.id, 0, 0)]
    for (g1, g2, p) in x.clades[g].splits:
        r -= p * _getl(x, ell, f.id, g1) * _getl(x, ell, h.id, g2)
        if r < 0.0:
            return [(f.id, g1, tf), (h.id, g2, th)]
        r -= p * _getl(x, ell, h.id, g1) * _getl(x, ell, f.id, g2)
        if r < 0.0:
            return [(h.id, g1, th), (f.id, g2, tf)]
    raise BacktrackFailed(f"r={r} at internal {(e, g)}")


def _bt_wgd(r, wm, x, ell, n, e, g):  # :258-266 with wgdloss :342-350, wgdretention :352-367
    q = gettheta(wm.rates, n)["q"]
    f = n.children[0]
    tf = len(f)
    r -= (1.0 - q + 2.0 * q * f.eps[-1]) * _getl(x, ell, f.id, g)
    if r < 0.0:
        return [(f.id, g, tf)]
    for (g1, g2, p) in x.clades[g].splits:
        r -= q * p * _getl(x, ell, f.id, g1) * _getl(x, ell, f.id, g2)
        if r < 0.0:
            return [(f.id, g1, tf), (f.id, g2, tf)]
    raise BacktrackFailed(f"r={r} at wgd {(e, g)}")


def _bt_root(r, wm, x, ell, n, e, g, t):  # :245-256 with rootbifurcation :369-397, rootloss :399-414
    eta = gettheta(wm.rates, n)["eta"]
    eps = n.eps[-1]
    xi = 1.0 - (1.0 - eta) * eps
    f, h = n.children
    tf, th = len(f), len(h)
    for (g1, g2, p) in x.clades[g].splits:
        r -= p * _getl(x, ell, e, g1, t) * _getl(x, ell, e, g2, t) * xi * (1.0 - eta) / eta
        if r < 0.0:
            return [(e, g1, t), (e, g2, t)]
        r -= p * _getl(x, ell, f.id, g1) * _getl(x, ell, h.id, g2) * eta * (1.0 - eps) / xi ** 2
        if r < 0.0:
            return [(f.id, g1, tf), (h.id, g2, th)]
        r -= p * _getl(x, ell, h.id, g1) * _getl(x, ell, f.id, g2) * eta * (1.0 - eps) / xi ** 2
        if r < 0.0:
            return [(h.id, g1, th), (f.id, g2, tf)]
    r -= _getl(x, ell, f.id, g) * h.eps[-1] * eta * (1.0 - eps) / xi ** 2
    if r < 0.0:
        return [(f.id, g, tf), (h.id, 0, 0)]
    r -= _getl(x, ell, h.id, g) * f.eps[-1] * eta * (1.0 - eps) / xi ** 2
    if r < 0.0:
        return [(h.id, g, th), (f.id, 0, 0)]
    raise BacktrackFailed(f"r={r} at root {(e, g)}")


# --------------------------------------------------------------------------------------------
# the reference's own test models (test/runtests.jl)
# --------------------------------------------------------------------------------------------
def c1_tree() -> TNode:
    """test/runtests.jl:8-11."""
    t = readnw(EXTREE)
    insertnode(getlca(t, "ATHA", "ATHA"), name="wgd_1")
    insertnode(getlca(t, "ATHA", "ATRI"), name="wgd_2")
    return t


def c1_model(maxn=10000, dt=0.05, minn=5, condition="root") -> WhaleModel:
    """test/runtests.jl:12-21: DLWGD(λ=μ=ones(17), q=[0.2,0.1], η=0.9)."""
    r = DLWGD(lam=[1.0] * 17, mu=[1.0] * 17, q=[0.2, 0.1], eta=0.9)
    return WhaleModel(r, c1_tree(), dt, minn=minn, maxn=maxn, condition=condition)
