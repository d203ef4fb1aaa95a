"""The product's C++ policy layer (rabe_b200/csrc/host_policy.cpp, through the C ABI) against the
reference's own L2 fixtures and, on random trees, against the oracle's independent Python
restatement (oracle/policy.py)."""
import hashlib
import random

import pytest

from oracle import policy as OP
from oracle.pyref import R
from rabe_b200.error import RabeError
from rabe_b200.policy import Policy, PolicyLanguage, sha3_hash_fr

J, H = PolicyLanguage.JsonPolicy, PolicyLanguage.HumanPolicy


def test_msp_golden():                      # msp.rs:157-199
    pol = '{name:"and", children:[{name:"A"}, {name:"or", "children":[{name:"D"}, {name:"and", "children":[{name:"B"},{name:"C"}]}]} ]}'
    assert Policy(pol, J).msp() == ([[1, 1, 0], [0, -1, 1], [0, 0, -1], [0, -1, 0]], ["A", "B", "C", "D"], 3)


def test_pruning_golden():                  # secretsharing/mod.rs:286-324
    attrs = ["A", "B", "C"]
    pol1 = '{"name": "or", "children": [{"name": "and", "children": [{"name": "A"}, {"name": "B"}]}, {"name": "and", "children": [{"name": "C"}, {"name": "D"}]}]}'
    pol2 = '{"name": "or", "children": [{"name": "C"}, {"name": "and", "children": [{"name": "A"}, {"name": "E"}]}]}'
    pol3 = '{"name": "or", "children": [{"name": "and", "children": [{"name": "A"}, {"name": "C"}]}, {"name": "and", "children": [{"name": "C"}, {"name": "A"}]}]}'
    assert Policy(pol1, J).prune(attrs) == (True, [("A", "A_68"), ("B", "B_83")])
    assert Policy(pol2, J).prune(attrs) == (True, [("C", "C_39")])
    assert Policy(pol3, J).prune(attrs) == (True, [("A", "A_68"), ("C", "C_83")])


def test_parse_serialize():                 # pest/mod.rs:119-149
    for pol, human in ((r'{"name": "A"}', "A"),
                       (r'{"name": "and", "children": [{"name": "B"}, {"name": "C"}]}', "(B and C)"),
                       (r'{"name": "or", "children": [{"name": "A"}, {"name": "and", "children": [{"name": "B"}, {"name": "C"}]}]}', "(A or (B and C))")):
        p = Policy(pol, J)
        assert p.serialize(J) == pol and p.serialize(H) == human


def test_traverse_truth_table():            # tools/mod.rs:77-129
    p1 = Policy('{"name": "and", "children": [{"name": "A"}, {"name": "B"}]}', J)
    p2 = Policy('{"name": "or", "children": [{"name": "A"}, {"name": "B"}]}', J)
    p3 = Policy('{"name": "and", "children": [{"name":"or", "children": [{"name": "C"}, {"name": "D"}]}, {"name": "B"}]}', J)
    s0, s1, s2, s3 = ["X", "Y"], ["A", "B"], ["C", "D"], ["A", "B", "C", "D"]
    with pytest.raises(RabeError):
        Policy("what-the-heck?", J)
    assert [p1.satisfied(s) for s in (s0, s1, s2, s3)] == [False, True, False, True]
    assert [p2.satisfied(s) for s in (s1, s2, s3)] == [True, False, True]
    assert [p3.satisfied(s) for s in (s1, s2, s3)] == [False, False, True]
    assert p2.satisfied([]) is False


def test_errors_instead_of_panics():
    with pytest.raises(RabeError):
        Policy('"A" and "B" and "C"', H).msp()          # msp.rs:132 panics
    with pytest.raises(RabeError):
        Policy('{"name": "and", "children": [{"name": "A"}]}', J).msp()   # msp.rs:121 panics
    with pytest.raises(RabeError):
        Policy('{"name": "and", "children": [{"name": "A"}]}', J).prune(["A"])   # secretsharing:167 panics
    for bad in ('"A" and "B" or "C"', '"A" and', "", '("A" and "B"', '"A" "B"', "42"):
        with pytest.raises(RabeError):
            Policy(bad, H)


def test_hash_to_fr():
    assert int.from_bytes(sha3_hash_fr("A00"), "big") == \
        10390014792917408443610864756208359696845607054198933325498802947006247339737
    rng = random.Random(1)
    for _ in range(200):
        s = "".join(rng.choice("abcXYZ019_:") for _ in range(rng.randrange(0, 300)))
        exp = int.from_bytes(hashlib.sha3_256(s.encode()).digest(), "big") % R
        assert int.from_bytes(sha3_hash_fr(s), "big") == exp


def _rand_tree(rng, names, depth=0):
    if len(names) == 1 or (depth > 0 and rng.random() < 0.15):
        return ("leaf", rng.choice(names))
    k = rng.randrange(1, len(names)) if len(names) > 1 else 1
    kind = rng.choice(["and", "or"])
    if kind == "and" or rng.random() < 0.5:
        return (kind, [_rand_tree(rng, names[:k], depth + 1), _rand_tree(rng, names[k:], depth + 1)])
    cuts = sorted(rng.sample(range(1, len(names)), min(len(names) - 1, rng.randrange(1, 4))))
    parts = [names[a:b] for a, b in zip([0] + cuts, cuts + [len(names)])]
    return (kind, [_rand_tree(rng, p, depth + 1) for p in parts])


def _human(t, rng):
    if t[0] == "leaf":
        return f'"{t[1]}"'
    op = rng.choice({"and": ["and", "AND", "&&"], "or": ["or", "OR", "||"]}[t[0]])
    ws = rng.choice([" ", "  ", "\n ", " /* c */ "])
    o, c = rng.choice(["()", "[]", "{}"])
    return o + (ws + op + ws).join(_human(k, rng) for k in t[1]) + c


def _json(t, rng):
    if t[0] == "leaf":
        return '{"name": "%s"}' % t[1]
    return '{"name": "%s", %s: [%s]}' % (t[0], rng.choice(['"children"', "children", "CHILDREN"]), ", ".join(_json(k, rng) for k in t[1]))


def test_random_trees_match_oracle_restatement():
    rng = random.Random(77)
    names = [f"a{i}" for i in range(12)] + ["b", "a1"]          # duplicates allowed
    for it in range(300):
        tree = _rand_tree(rng, names[:rng.randrange(1, len(names))])
        for text, lang, olang in ((_human(tree, rng), H, OP.HUMAN), (_json(tree, rng), J, OP.JSON)):
            ot = OP.parse(text, olang)
            p = Policy(text, lang)
            assert p.serialize(J) == OP.serialize_policy(ot, OP.JSON), text
            try:
                om = OP.calculate_msp(ot)
            except OP.PolicyError:
                om = None
            if om is None:
                with pytest.raises(RabeError):
                    p.msp()
            else:
                assert p.msp() == (om[0], om[1], om[2]), text
            attrs = [n for n in names if rng.random() < 0.6]
            assert p.satisfied(attrs) == OP.traverse_policy(attrs, ot)
            assert p.prune(attrs) == OP.calc_pruned(attrs, ot), text
