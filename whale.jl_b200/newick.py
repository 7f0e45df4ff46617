"""Species-tree handling the reference delegates to NewickTree.jl 0.3.1 (Manifest.toml:1061-1065).

Host prep only (runs once per model): Newick parsing, `getlca`, and `insertnode` with NewickTree's
semantics — the new node sits halfway along the branch and is appended as the LAST child of the former
parent (pinned by the reference's own test, test/runtests.jl:54-55).
"""
from __future__ import annotations

import math


class Node:
    __slots__ = ("name", "distance", "children", "parent")

    def __init__(self, name: str = "", distance: float = math.nan):
        self.name = name
        self.distance = distance
        self.children: list[Node] = []
        self.parent: Node | None = None

    def add(self, child: "Node") -> "Node":
        child.parent = self
        self.children.append(child)
        return child

    @property
    def isleaf(self) -> bool:
        return len(self.children) == 0

    @property
    def isroot(self) -> bool:
        return self.parent is None

    def __getitem__(self, i: int) -> "Node":
        """1-based child access, like NewickTree's `n[i]`."""
        return self.children[i - 1]

    def __repr__(self):
        return f"Node({nwstr(self)})"


def readnw(text: str) -> Node:
    """Parse one Newick string; internal labels are kept only when they name a WGD (`wgd...`)."""
    text = text.strip()
    if not text.endswith(";"):
        raise ValueError("Newick string must end with ';'")
    stack: list[Node] = []
    cur = Node()
    root = cur
    i, n = 0, len(text)
    while i < n:
        ch = text[i]
        if ch == "(":
            child = Node()
            cur.add(child)
            stack.append(cur)
            cur = child
            i += 1
        elif ch == ",":
            parent = stack[-1]
            cur = parent.add(Node())
            i += 1
        elif ch == ")":
            cur = stack.pop()
            i += 1
        elif ch == ";":
            break
        elif ch == ":":
            j = i + 1
            while j < n and text[j] not in ",();":
                j += 1
            cur.distance = float(text[i + 1:j])
            i = j
        else:
            j = i
            while j < n and text[j] not in ",():;":
                j += 1
            label = text[i:j].strip()
            if cur.isleaf or label.startswith("wgd"):
                cur.name = label
            i = j
    if stack:
        raise ValueError("unbalanced parentheses in Newick string")
    return root


def getleaves(n: Node) -> list[Node]:
    out, todo = [], [n]
    while todo:
        x = todo.pop()
        if x.isleaf:
            out.append(x)
        else:
            todo.extend(reversed(x.children))
    return out


def postwalk(n: Node) -> list[Node]:
    out: list[Node] = []

    def rec(x):
        for c in x.children:
            rec(c)
        out.append(x)

    rec(n)
    return out


def getroot(n: Node) -> Node:
    while n.parent is not None:
        n = n.parent
    return n


def getlca(tree: Node, a: str, b: str) -> Node:
    byname = {}
    for l in getleaves(tree):
        byname[l.name] = l
    path = set()
    x = byname[a]
    while x is not None:
        path.add(id(x))
        x = x.parent
    y = byname[b]
    while id(y) not in path:
        y = y.parent
    return y


def insertnode(n: Node, name: str = "", dist: float = math.nan) -> Node:
    """`insertnode!(n; name, dist)`: put a node on the branch above `n`, `dist` above it
    (default: halfway); it becomes the last child of n's former parent."""
    p = n.parent
    if p is None:
        raise ValueError("cannot insert a node above the root")
    d = n.distance / 2 if math.isnan(dist) else dist
    new = Node(name=name, distance=n.distance - d)
    n.distance = d
    p.children.remove(n)
    p.add(new)
    new.add(n)
    return new


def nwstr(n: Node, dist: bool = False) -> str:
    s = n.name if n.isleaf else "(" + ",".join(nwstr(c, dist) for c in n.children) + ")" + n.name
    if dist and not math.isnan(n.distance):
        s += f":{n.distance}"
    return s


def extree() -> Node:
    """`Whale.extree` (src/Whale.jl:34-37): a fresh copy of the bundled 9-taxon land-plant tree."""
    return readnw("((MPOL:4.752,PPAT:4.752):0.292,(SMOE:4.457,(((OSAT:1.555,(ATHA:0.5548,CPAP:0.5548):1.0002):0"
                  ".738,ATRI:2.293):1.225,(GBIL:3.178,PABI:3.178):0.34):0.939):0.587);")
