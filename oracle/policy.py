"""TEST INFRASTRUCTURE ONLY -- Python restatement of rabe's policy layer (strings + Fr bigints).

Restates, function by function:
  parse()              /root/reference/src/utils/policy/pest/mod.rs:40-66 with the PEG grammars
                       src/human.policy.pest and src/json.policy.pest and the tree builders
                       pest/human.rs:8-49, pest/json.rs:8-48 (leaf = (text, pest column))
  calculate_msp / lw   src/utils/policy/msp.rs:78-147
  gen_shares_policy    src/utils/secretsharing/mod.rs:82-141,215-221
  calc_coefficients    src/utils/secretsharing/mod.rs:9-72
  calc_pruned          src/utils/secretsharing/mod.rs:143-201
  traverse_policy      src/utils/tools/mod.rs:31-61
These ARE pinned by the reference's own fixtures (msp.rs:157-199, secretsharing/mod.rs:229-324,
pest/mod.rs:119-149, tools/mod.rs:77-129); tests/test_oracle_policy.py replays them.

Tree representation: ("and", [children]) | ("or", [children]) | ("leaf", name, column).
Only tests/, bench.py's CPU legs and smoke() may import this module.
"""
from .pyref import R

HUMAN, JSON = "human", "json"


class PolicyError(ValueError):
    pass


# ------------------------------------------------------------------------------------ PEG helpers
class _P:
    def __init__(self, s):
        self.s = s

    def col(self, pos):
        nl = self.s.rfind("\n", 0, pos)
        return pos - nl                     # 1-based, counted in characters

    def skip(self, pos):
        s = self.s
        while True:
            if pos < len(s) and s[pos] in " \t\r\n":
                pos += 1
            elif s.startswith("/*", pos):
                end = s.find("*/", pos + 2)
                if end < 0:
                    return pos
                pos = end + 2
            else:
                return pos

    def lit(self, pos, options):
        for o in options:
            if self.s.startswith(o, pos):
                return pos + len(o)
        return None

    def string(self, pos):
        s = self.s
        if not s.startswith('"', pos):
            return None
        i = pos + 1
        start = i
        while i < len(s):
            ch = s[i]
            if ch == '"':
                return ("leaf", s[start:i], self.col(start)), i + 1
            if ch == "\\":
                if i + 1 < len(s) and s[i + 1] in '"\\/bfnrt':
                    i += 2
                elif s.startswith("u", i + 1) and len(s) >= i + 6 and all(c in "0123456789abcdefABCDEF" for c in s[i + 2:i + 6]):
                    i += 6
                else:
                    return None
            else:
                i += 1
        return None

    def number(self, pos):
        s = self.s
        i = pos
        if i < len(s) and s[i] == "-":
            i += 1
        if i < len(s) and s[i] == "0":
            i += 1
        elif i < len(s) and s[i] in "123456789":
            while i < len(s) and s[i].isdigit() and s[i].isascii():
                i += 1
        else:
            return None
        if i < len(s) and s[i] == ".":
            i += 1
            while i < len(s) and s[i].isdigit() and s[i].isascii():
                i += 1
        if i < len(s) and s[i] in "eE":
            j = i + 1
            if j < len(s) and s[j] in "+-":
                j += 1
            if j < len(s) and s[j].isdigit():
                while j < len(s) and s[j].isdigit() and s[j].isascii():
                    j += 1
                i = j
        # the reference builders call `pair.into_inner().next().unwrap()` on an atomic rule with
        # no inner pairs (human.rs:15-17, json.rs:15-17) => panic.  Reported as an error here.
        raise PolicyError("number leaves are not supported (the reference panics on them)")

    def inner_kw(self, pos, words):
        """andinner / orinner: bare keyword, or QUOTE ~ keyword ~ QUOTE (implicit whitespace allowed)."""
        e = self.lit(pos, words)
        if e is not None:
            return e
        if self.s.startswith('"', pos):
            p = self.skip(pos + 1)
            e = self.lit(p, words)
            if e is not None:
                p = self.skip(e)
                if self.s.startswith('"', p):
                    return p + 1
        return None


_AND = ("and", "AND", "&&")
_OR = ("or", "OR", "||")


class _Human(_P):
    def value(self, pos):
        r = self.string(pos)
        if r:
            return r
        self.number(pos) if (pos < len(self.s) and (self.s[pos] == "-" or self.s[pos].isdigit())) else None
        if pos < len(self.s) and self.s[pos] in "([{":
            p = self.skip(pos + 1)
            r = self.node(p)
            if r:
                p = self.skip(r[1])
                if p < len(self.s) and self.s[p] in ")]}":
                    return r[0], p + 1
        return None

    def term(self, pos):
        r = self.value(pos)
        if r:
            return r
        if self.s.startswith("(", pos):
            p = self.skip(pos + 1)
            r = self.node(p)
            if r:
                p = self.skip(r[1])
                if self.s.startswith(")", p):
                    return r[0], p + 1
        return None

    def gate(self, pos, words, kind):
        r = self.term(pos)
        if not r:
            return None
        kids, p = [r[0]], r[1]
        while True:
            q = self.skip(p)
            e = self.inner_kw(q, words)
            if e is None:
                break
            q = self.skip(e)
            r = self.term(q)
            if not r:
                break
            kids.append(r[0]); p = r[1]
        if len(kids) < 2:
            return None
        return (kind, kids), p

    def node(self, pos):
        return self.gate(pos, _AND, "and") or self.gate(pos, _OR, "or") or self.term(pos)


class _Json(_P):
    _NAME = ("name", "NAME")
    _CHILDREN = ("children", "CHILDREN")

    def key(self, pos, words):
        e = self.lit(pos, words)
        if e is not None:
            return e
        if self.s.startswith('"', pos):
            p = self.skip(pos + 1)
            e = self.lit(p, words)
            if e is not None:
                p = self.skip(e)
                if self.s.startswith('"', p):
                    return p + 1
        return None

    def tok(self, pos, ch):
        return pos + 1 if self.s.startswith(ch, pos) else None

    def gate(self, pos, words, kind):
        e = self.inner_kw(pos, words)
        if e is None: return None
        p = self.tok(self.skip(e), ",")
        if p is None: return None
        p = self.key(self.skip(p), self._CHILDREN)
        if p is None: return None
        p = self.tok(self.skip(p), ":")
        if p is None: return None
        p = self.tok(self.skip(p), "[")
        if p is None: return None
        q = self.tok(self.skip(p), "]")
        if q is not None:
            return (kind, []), q
        r = self.node(self.skip(p))
        if not r: return None
        kids, p = [r[0]], r[1]
        while True:
            q = self.tok(self.skip(p), ",")
            if q is None: break
            r = self.node(self.skip(q))
            if not r: break
            kids.append(r[0]); p = r[1]
        p = self.tok(self.skip(p), "]")
        if p is None: return None
        return (kind, kids), p

    def node(self, pos):
        p = self.tok(pos, "{")
        if p is None: return None
        p = self.key(self.skip(p), self._NAME)
        if p is None: return None
        p = self.tok(self.skip(p), ":")
        if p is None: return None
        body = self.skip(p)
        for alt in (self.string, lambda q: self.gate(q, _AND, "and"), lambda q: self.gate(q, _OR, "or")):
            r = alt(body)
            if r:
                e = self.tok(self.skip(r[1]), "}")
                if e is not None:
                    return r[0], e
        if body < len(self.s) and (self.s[body] == "-" or self.s[body].isdigit()):
            self.number(body)
        return None


def parse(policy: str, language: str):
    """pest/mod.rs:40-66."""
    p = _Human(policy) if language == HUMAN else _Json(policy)
    start = p.skip(0)
    r = p.node(start)
    if not r or p.skip(r[1]) != len(policy):
        raise PolicyError(f"{language} policy parse error")
    return r[0]


def serialize_policy(node, language, parent=None):
    """pest/mod.rs:68-111."""
    kind = node[0]
    if language == JSON:
        if kind == "leaf":
            return '{"name": "%s"}' % node[1]
        inner = ", ".join(serialize_policy(c, language) for c in node[1])
        return '{"name": "%s", "children": [%s]}' % (kind, inner)
    if kind == "leaf":
        return node[1]
    return "(" + (" %s " % kind).join(serialize_policy(c, language) for c in node[1]) + ")"


# ------------------------------------------------------------------------------------ MSP
def calculate_msp(tree):
    """msp.rs:78-147.  Returns (m rows, pi, c)."""
    state = {"m": [], "pi": [], "c": 1}

    def lw(p, v):
        if p[0] == "leaf":
            state["m"].insert(0, list(v)); state["pi"].insert(0, p[1])
            return True
        kids = p[1]
        if len(kids) < 2:
            raise PolicyError("lw: policy with just a single attribute is not allowed")
        if p[0] == "or":
            ret = True
            for k in kids:
                ret &= lw(k, v)
            return ret
        if len(kids) != 2:
            raise PolicyError("lw: Invalid policy. Number of arguments under AND != 2")
        right = list(v) + [0] * (state["c"] - len(v))
        right = right[:state["c"]] + [1]
        left = [0] * state["c"] + [-1]
        state["c"] += 1
        return lw(kids[0], right) and lw(kids[1], left)

    if not lw(tree, [1]):
        raise PolicyError("lewko waters algorithm failed =(")
    c = state["c"]
    rows = [(r + [0] * c)[:c] for r in state["m"]]
    order = sorted(range(len(rows)), key=lambda i: state["pi"][i].encode())     # stable, byte order
    return [rows[i] for i in order], [state["pi"][i] for i in order], c


# ------------------------------------------------------------------------------------ sharing
def node_index(leaf):
    return "%s_%d" % (leaf[1], leaf[2])


def remove_index(label):
    return label.split("_")[0]


def polynomial(coeff, x):
    """secretsharing/mod.rs:215-221."""
    return sum(c * pow(x, i, R) for i, c in enumerate(coeff)) % R


def gen_shares_policy(secret, tree, rnd):
    """secretsharing/mod.rs:82-141.  `rnd` is an iterator yielding the Fr values the reference draws
    (k-1 per AND gate with k children, none for OR), in pre-order."""
    if tree[0] == "leaf":
        return [(node_index(tree), secret % R)]
    kids = tree[1]
    n = len(kids)
    k = n if tree[0] == "and" else 1
    a = [secret % R] + [next(rnd) % R for _ in range(1, k)]
    shares = [polynomial(a, i) for i in range(n + 1)]
    out = []
    for i in range(n):
        out.extend(gen_shares_policy(shares[i + 1], kids[i], rnd))
    return out


def count_share_randomness(tree):
    if tree[0] == "leaf":
        return 0
    own = len(tree[1]) - 1 if tree[0] == "and" else 0
    return own + sum(count_share_randomness(k) for k in tree[1])


def recover_coefficients(points):
    """secretsharing/mod.rs:60-72 (Lagrange at 0)."""
    out = []
    for i in points:
        res = 1
        for j in points:
            if i != j:
                res = res * ((-j) % R) * pow((i - j) % R, -1, R) % R
        out.append(res)
    return out


def calc_coefficients(tree, coeff=1):
    """secretsharing/mod.rs:9-58.  Returns [(node_index, coeff)] in DFS order."""
    if tree[0] == "leaf":
        return [(node_index(tree), coeff % R)]
    kids = tree[1]
    if tree[0] == "and":
        this = recover_coefficients(list(range(1, len(kids) + 1)))
    else:
        this = [1] * len(kids)
    out = []
    for i, k in enumerate(kids):
        out.extend(calc_coefficients(k, coeff * this[i] % R))
    return out


def calc_pruned(attrs, tree):
    """secretsharing/mod.rs:143-201.  Returns (match, [(name, node_index)])."""
    if tree[0] == "leaf":
        if tree[1] in attrs:
            return True, [(tree[1], node_index(tree))]
        return False, []
    kids = tree[1]
    if len(kids) < 2:
        raise PolicyError("Invalid policy (gate with just a single child)")
    if tree[0] == "and":
        ok, acc = True, []
        for k in kids:
            found, lst = calc_pruned(attrs, k)
            ok = ok and found
            if ok:
                acc.extend(lst)
        return (ok, acc if ok else [])
    for k in kids:
        found, lst = calc_pruned(attrs, k)
        if found:
            return True, lst
    return False, []


def traverse_policy(attrs, tree):
    """tools/mod.rs:31-61 called with PolicyType::Leaf at the root."""
    if len(attrs) == 0:
        return False
    if tree[0] == "leaf":
        return tree[1] in attrs
    if tree[0] == "and":
        ret = True
        for k in tree[1]:
            ret &= traverse_policy(attrs, k)
        return ret
    ret = False
    for k in tree[1]:
        ret |= traverse_policy(attrs, k)
    return ret
