"""Replays the reference's own L2 fixtures against the oracle's policy layer (oracle/policy.py):
msp.rs:157-199, secretsharing/mod.rs:229-324, pest/mod.rs:119-149, tools/mod.rs:77-129."""
import random

import pytest

from oracle import policy as P
from oracle.pyref import R


def test_msp_golden():           # msp.rs:157-199
    pol = '{name:"and", children:[{name:"A"}, {name:"or", "children":[{name:"D"}, {name:"and", "children":[{name:"B"},{name:"C"}]}]} ]}'
    m, pi, c = P.calculate_msp(P.parse(pol, P.JSON))
    assert m == [[1, 1, 0], [0, -1, 1], [0, 0, -1], [0, -1, 0]] and pi == ["A", "B", "C", "D"] and c == 3


def test_msp_config1_hand_derived():
    m, pi, c = P.calculate_msp(P.parse('("A" and "B") and ("C" and "D")', P.HUMAN))
    assert m == [[1, 1, 1, 0], [0, 0, -1, 0], [0, -1, 0, 1], [0, 0, 0, -1]] and c == 4
    assert [sum(col) for col in zip(*m)] == [1, 0, 0, 0]


def test_msp_rejects_nary_and():         # msp.rs:132 panics
    with pytest.raises(P.PolicyError):
        P.calculate_msp(P.parse('"A" and "B" and "C"', P.HUMAN))


def test_pruning_golden():       # secretsharing/mod.rs:286-324
    attrs = ["A", "B", "C"]
    pol1 = '{"name": "or", "children": [{"name": "and", "children": [{"name": "A"}, {"name": "B"}]}, {"name": "and", "children": [{"name": "C"}, {"name": "D"}]}]}'
    pol2 = '{"name": "or", "children": [{"name": "C"}, {"name": "and", "children": [{"name": "A"}, {"name": "E"}]}]}'
    pol3 = '{"name": "or", "children": [{"name": "and", "children": [{"name": "A"}, {"name": "C"}]}, {"name": "and", "children": [{"name": "C"}, {"name": "A"}]}]}'
    assert P.calc_pruned(attrs, P.parse(pol1, P.JSON)) == (True, [("A", "A_68"), ("B", "B_83")])
    assert P.calc_pruned(attrs, P.parse(pol2, P.JSON)) == (True, [("C", "C_39")])
    assert P.calc_pruned(attrs, P.parse(pol3, P.JSON)) == (True, [("A", "A_68"), ("C", "C_83")])


def test_parse_serialize():      # pest/mod.rs:119-149
    for pol, human in ((r'{"name": "A"}', "A"),
                       (r'{"name": "and", "children": [{"name": "B"}, {"name": "C"}]}', "(B and C)"),
                       (r'{"name": "or", "children": [{"name": "A"}, {"name": "and", "children": [{"name": "B"}, {"name": "C"}]}]}', "(A or (B and C))")):
        t = P.parse(pol, P.JSON)
        assert P.serialize_policy(t, P.JSON) == pol and P.serialize_policy(t, P.HUMAN) == human


def test_traverse_truth_table():  # tools/mod.rs:77-129
    p1 = P.parse('{"name": "and", "children": [{"name": "A"}, {"name": "B"}]}', P.JSON)
    p2 = P.parse('{"name": "or", "children": [{"name": "A"}, {"name": "B"}]}', P.JSON)
    p3 = P.parse('{"name": "and", "children": [{"name":"or", "children": [{"name": "C"}, {"name": "D"}]}, {"name": "B"}]}', P.JSON)
    s0, s1, s2, s3 = ["X", "Y"], ["A", "B"], ["C", "D"], ["A", "B", "C", "D"]
    with pytest.raises(P.PolicyError):
        P.parse("what-the-heck?", P.JSON)
    assert [P.traverse_policy(s, p1) for s in (s0, s1, s2, s3)] == [False, True, False, True]
    assert [P.traverse_policy(s, p2) for s in (s1, s2, s3)] == [True, False, True]
    assert [P.traverse_policy(s, p3) for s in (s1, s2, s3)] == [False, False, True]
    assert P.traverse_policy([], p2) is False


def test_share_recover():        # secretsharing/mod.rs:229-283
    rng = random.Random(9)
    for pol in ('{"name":"or", "children": [{"name": "A"}, {"name": "B"}]}',
                '{"name": "and", "children": [{"name": "A"}, {"name": "B"}]}',
                '{"name": "and", "children": [{"name": "A"}, {"name": "B"}, {"name": "C"}, {"name": "D"}]}',
                '{"name": "and", "children": [{"name": "A"}, {"name": "or", "children": [{"name": "B"}, {"name": "and", "children": [{"name": "C"}, {"name": "D"}, {"name": "E"}]}]}]}'):
        tree = P.parse(pol, P.JSON)
        secret = rng.randrange(R)
        rnd = iter([rng.randrange(R) for _ in range(P.count_share_randomness(tree))])
        shares = dict(P.gen_shares_policy(secret, tree, rnd))
        coeffs = dict(P.calc_coefficients(tree))
        assert len(coeffs) == len(shares)
        names = [l[0] for l in shares]
        ok, pruned = P.calc_pruned([n.split("_")[0] for n in names], tree)
        assert ok
        assert sum(coeffs[idx] * shares[idx] for _, idx in pruned) % R == secret


def test_human_grammar_edges():
    assert P.parse('"A"  /* c */ AND ("B" || "C")', P.HUMAN) == ("and", [("leaf", "A", 2), ("or", [("leaf", "B", 20), ("leaf", "C", 27)])])
    assert P.parse('["A" && "B"]', P.HUMAN)[0] == "and"
    for bad in ('"A" and "B" or "C"', '"A" and', "", '("A" and "B"', '"A" "B"'):
        with pytest.raises(P.PolicyError):
            P.parse(bad, P.HUMAN)
    # column is per line, 1-based, counted in characters
    assert P.parse('"A" and\n  "B"', P.HUMAN) == ("and", [("leaf", "A", 2), ("leaf", "B", 4)])
