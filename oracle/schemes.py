"""TEST INFRASTRUCTURE ONLY -- reference-sequence restatements of rabe's BSW, LSW and AW11 schemes
in Python over the oracle's primitives (oracle/liboracle.so), statement by statement:

    BSW   /root/reference/src/schemes/bsw/mod.rs:92-318
    LSW   /root/reference/src/schemes/lsw/mod.rs:86-290
    AW11  /root/reference/src/schemes/aw11/mod.rs:100-390
    GHW11 /root/reference/src/schemes/ghw11/mod.rs:92-305 (outsourced decryption: tkgen / transform / decrypt_out)

Every value rabe draws from rand::thread_rng() is taken, in the reference's draw order, from the
iterator `rnd` (ints); random group elements are generator multiples (G1gen*rho), and the random
Gt `msg` is an explicit argument.  Group elements are canonical byte strings.
PARITY UNPINNED against rabe itself (see oracle/__init__.py).
"""
import oracle as o
from . import policy as P
from .pyref import R

G1_ZERO, G2_ZERO = b"\0" * 64, b"\0" * 128


def _fr(x):
    return int(x % R).to_bytes(32, "big")


def _int(b):
    return int.from_bytes(b, "big")


def _hash_fr(s):
    return _int(o.sha3_fr(s))


def _hash_g1(g, s):          # hash/mod.rs:10-20 with T = G1
    return o.g1_mul(g, o.sha3_fr(s))


def _hash_g2(g, s):
    return o.g2_mul(g, o.sha3_fr(s))


def gt_random(rho):
    return o.gt_pow(o.pairing(o.g1_generator(), o.g2_generator()), _fr(rho))


# ======================================================================================= BSW
def bsw_setup(rnd):
    """bsw/mod.rs:92-115.  draws: g1, g2, beta, alpha."""
    g1 = o.g1_mul(o.g1_generator(), _fr(next(rnd)))
    g2 = o.g2_mul(o.g2_generator(), _fr(next(rnd)))
    beta, alpha = next(rnd) % R, next(rnd) % R
    h = o.g1_mul(g1, _fr(beta))
    f = o.g2_mul(g2, _fr(pow(beta, -1, R)))
    g2_alpha = o.g2_mul(g2, _fr(alpha))
    e_gg_alpha = o.pairing(g1, g2_alpha)
    return {"g1": g1, "g2": g2, "h": h, "f": f, "e_gg_alpha": e_gg_alpha}, {"beta": beta, "g2_alpha": g2_alpha}


def bsw_keygen(pk, msk, attributes, rnd):
    """bsw/mod.rs:125-152.  draws: r, then r_j per attribute."""
    if len(attributes) == 0:
        return None
    r = next(rnd) % R
    g2_r = o.g2_mul(pk["g2"], _fr(r))
    d = o.g2_mul(o.g2_add(msk["g2_alpha"], g2_r), _fr(pow(msk["beta"], -1, R)))
    d_j = []
    for j in attributes:
        r_j = next(rnd) % R
        d_j.append((j, o.g1_mul(pk["g1"], _fr(r_j)), o.g2_add(g2_r, o.g2_mul(_hash_g2(pk["g2"], j), _fr(r_j)))))
    return {"d": d, "d_j": d_j}


def bsw_delegate(pk, sk, subset, rnd):
    """bsw/mod.rs:162-206.  draws: r, then r_j per attribute of the subset."""
    names = [x[0] for x in sk["d_j"]]
    if not set(subset) <= set(names) or len(subset) == 0:
        return None
    r = next(rnd) % R
    d_j = []
    for attr in subset:
        r_j = next(rnd) % R
        _, g1v, g2v = next(x for x in sk["d_j"] if x[0] == attr)
        d_j.append((attr, o.g1_add(g1v, o.g1_mul(pk["g1"], _fr(r_j))),
                    o.g2_add(o.g2_add(g2v, o.g2_mul(_hash_g2(pk["g2"], attr), _fr(r_j))), o.g2_mul(pk["g2"], _fr(r)))))
    return {"d": o.g2_add(sk["d"], o.g2_mul(pk["f"], _fr(r))), "d_j": d_j}


def bsw_encrypt(pk, policy, language, msg, rnd):
    """bsw/mod.rs:217-251.  draws: secret, (msg,) then the gen_shares coefficients."""
    secret = next(rnd) % R
    tree = P.parse(policy, language)
    shares = P.gen_shares_policy(secret, tree, rnd)
    c = o.g1_mul(pk["h"], _fr(secret))
    c_p = o.gt_mul(o.gt_pow(pk["e_gg_alpha"], _fr(secret)), msg)
    c_y = []
    for node, val in shares:
        j = P.remove_index(node)
        c_y.append((node, o.g1_mul(pk["g1"], _fr(val)), o.g2_mul(_hash_g2(pk["g2"], j), _fr(val))))
    return {"policy": (policy, language), "c": c, "c_p": c_p, "c_y": c_y}


def bsw_decrypt(sk, ct):
    """bsw/mod.rs:260-318.  Returns the Gt `_msg` (or None where rabe returns Err)."""
    attr = [x[0] for x in sk["d_j"]]
    tree = P.parse(*ct["policy"])
    if not P.traverse_policy(attr, tree):
        return None
    ok, pruned = P.calc_pruned(attr, tree)
    if not ok:
        return None
    z = P.calc_coefficients(tree)
    a = o.GT_ONE
    for k, j in pruned:
        c_y = next((x for x in ct["c_y"] if x[0] == j), None)
        d_j = next((x for x in sk["d_j"] if x[0] == k), None)
        if c_y is None or d_j is None:
            continue
        for label, zc in z:
            if label == j:
                t = o.gt_mul(o.pairing(c_y[1], d_j[2]), o.gt_inverse(o.pairing(d_j[1], c_y[2])))
                a = o.gt_mul(a, o.gt_pow(t, _fr(zc)))
    return o.gt_mul(ct["c_p"], o.gt_inverse(o.gt_mul(o.pairing(ct["c"], sk["d"]), o.gt_inverse(a))))


# ======================================================================================= LSW
def lsw_setup(rnd):
    """lsw/mod.rs:86-110.  draws: alpha1, alpha2, b, g1, g2, h_g1, h_g2."""
    alpha1, alpha2, b = next(rnd) % R, next(rnd) % R, next(rnd) % R
    g1 = o.g1_mul(o.g1_generator(), _fr(next(rnd)))
    g2 = o.g2_mul(o.g2_generator(), _fr(next(rnd)))
    h_g1 = o.g1_mul(o.g1_generator(), _fr(next(rnd)))
    h_g2 = o.g2_mul(o.g2_generator(), _fr(next(rnd)))
    g1_b = o.g1_mul(g1, _fr(b))
    g1_b2 = o.g1_mul(g1_b, _fr(b))
    h_b = o.g1_mul(h_g1, _fr(b))
    e_gg_alpha = o.gt_pow(o.pairing(g1, g2), _fr(alpha1 * alpha2))
    return ({"g1": g1, "g2": g2, "g1_b": g1_b, "g1_b2": g1_b2, "h_b": h_b, "e_gg_alpha": e_gg_alpha},
            {"alpha1": alpha1, "alpha2": alpha2, "b": b, "h_g1": h_g1, "h_g2": h_g2})


def lsw_keygen(pk, msk, policy, language, rnd):
    """lsw/mod.rs:121-170.  draws: the gen_shares coefficients, then `random` per leaf."""
    tree = P.parse(policy, language)
    shares = P.gen_shares_policy(msk["alpha1"], tree, rnd)
    dj = []
    for label, val in shares:
        striped = P.remove_index(label)
        rand = next(rnd) % R
        if striped.startswith("!"):
            h = _hash_fr(striped)
            dj.append((striped, G1_ZERO, G2_ZERO,
                       o.g1_add(o.g1_mul(pk["g1"], _fr(val)), o.g1_mul(pk["g1_b2"], _fr(rand))),
                       o.g1_add(o.g1_mul(pk["g1_b"], _fr(h * rand)), o.g1_mul(msk["h_g1"], _fr(rand))),
                       o.g1_mul(pk["g1"], _fr(-rand))))
        else:
            dj.append((striped,
                       o.g1_add(o.g1_mul(pk["g1"], _fr(msk["alpha2"] * val)), o.g1_mul(_hash_g1(pk["g1"], striped), _fr(rand))),
                       o.g2_mul(pk["g2"], _fr(rand)), G1_ZERO, G1_ZERO, G1_ZERO))
    return {"policy": (policy, language), "dj": dj}


def lsw_encrypt(pk, attributes, msg, rnd):
    """lsw/mod.rs:180-219.  draws: secret, then one Fr per attribute (the `sx` quirk at :197-200 is
    reproduced: sx[0] ends up as minus the sum of sx[1..n-1])."""
    if len(attributes) == 0:
        return None
    secret = next(rnd) % R
    sx = [secret]
    for i, _ in enumerate(attributes):
        sx.append(next(rnd) % R)
        sx[0] = (sx[0] - sx[i]) % R
    ej = []
    for i, attr in enumerate(attributes):
        ej.append((attr, o.g1_mul(_hash_g1(pk["g1"], attr), _fr(secret)), o.g1_mul(pk["g1_b"], _fr(sx[i])),
                   o.g1_add(o.g1_mul(pk["g1_b2"], _fr(sx[i] * _hash_fr(attr))), o.g1_mul(pk["h_b"], _fr(sx[i])))))
    e1 = o.gt_mul(o.gt_pow(pk["e_gg_alpha"], _fr(secret)), msg)
    e2 = o.g2_mul(pk["g2"], _fr(secret))
    return {"e1": e1, "e2": e2, "ej": ej}


def lsw_decrypt(sk, ct):
    """lsw/mod.rs:228-290 (positive attributes; the negative branch is a TODO in the reference)."""
    attr = [x[0] for x in ct["ej"]]
    tree = P.parse(*sk["policy"])
    ok, pruned = P.calc_pruned(attr, tree)
    if not ok:
        return None
    coeffs = P.calc_coefficients(tree)
    prod_t, z_y = o.GT_ONE, o.GT_ONE
    for name, label in pruned:
        sk_attr = next(x for x in sk["dj"] if x[0] == name)
        ct_attr = next(x for x in ct["ej"] if x[0] == name)
        coeff = next(c for l, c in coeffs if l == label)
        if not name.startswith("!"):
            z_y = o.gt_mul(o.pairing(sk_attr[1], ct["e2"]), o.gt_inverse(o.pairing(ct_attr[1], sk_attr[2])))
        prod_t = o.gt_mul(prod_t, o.gt_pow(z_y, _fr(coeff)))
    return o.gt_mul(ct["e1"], o.gt_inverse(prod_t))


# ======================================================================================= AW11
def aw11_setup(rnd):
    """aw11/mod.rs:100-108.  draws: g1, g2."""
    return {"g1": o.g1_mul(o.g1_generator(), _fr(next(rnd))), "g2": o.g2_mul(o.g2_generator(), _fr(next(rnd)))}


def aw11_authgen(gk, attributes, rnd):
    """aw11/mod.rs:121-151.  draws: alpha_i, y_i per attribute."""
    if len(attributes) == 0:
        return None
    sk, pk = [], []
    for attr in attributes:
        name = attr.upper()
        alpha_i, y_i = next(rnd) % R, next(rnd) % R
        sk.append((name, alpha_i, y_i))
        pk.append((name, o.gt_pow(o.pairing(gk["g1"], gk["g2"]), _fr(alpha_i)), o.g2_mul(gk["g2"], _fr(y_i))))
    return {"attr": pk}, {"attr": sk}


def aw11_keygen(gk, msk, name, attributes):
    """aw11/mod.rs:165-232 (no randomness)."""
    if len(attributes) == 0 or len(name) == 0:
        return None
    sk = {"gid": name, "attr": []}
    for attribute in attributes:
        h = _hash_g1(gk["g1"], sk["gid"])
        auth = next(x for x in msk["attr"] if x[0] == attribute)
        sk["attr"].append((auth[0].upper(), o.g1_add(o.g1_mul(gk["g1"], _fr(auth[1])), o.g1_mul(h, _fr(auth[2])))))
    return sk


def aw11_encrypt(gk, pks, policy, language, msg, rnd):
    """aw11/mod.rs:241-289.  draws: s, coefficients of the s-sharing, coefficients of the
    zero-sharing, (msg,) then r_x per share (drawn even when no authority holds the attribute)."""
    tree = P.parse(policy, language)
    P.calculate_msp(tree)                                   # :253 -- panics unless every AND is binary
    s = next(rnd) % R
    s_shares = P.gen_shares_policy(s, tree, rnd)
    w_shares = P.gen_shares_policy(0, tree, rnd)
    e_gg = o.pairing(gk["g1"], gk["g2"])
    c_0 = o.gt_mul(msg, o.gt_pow(e_gg, _fr(s)))
    c = []
    for i, (label, share) in enumerate(s_shares):
        r_x = next(rnd) % R
        want = P.remove_index(label.upper())
        found = None
        for pk in pks:
            found = next((x for x in pk["attr"] if x[0] == want), None)
            if found is not None:
                break
        if found is None:
            continue
        c.append((label.upper(),
                  o.gt_mul(o.gt_pow(o.pairing(gk["g1"], gk["g2"]), _fr(share)), o.gt_pow(found[1], _fr(r_x))),
                  o.g2_mul(gk["g2"], _fr(r_x)),
                  o.g2_add(o.g2_mul(found[2], _fr(r_x)), o.g2_mul(gk["g2"], _fr(w_shares[i][1])))))
    return {"policy": (policy, language), "c_0": c_0, "c": c}


def aw11_decrypt(gk, sk, ct):
    """aw11/mod.rs:298-372.  Returns the Gt `_msg` (or None where rabe returns Err)."""
    str_attr = [x[0] for x in sk["attr"]]
    tree = P.parse(*ct["policy"])
    if not P.traverse_policy(str_attr, tree):
        return None
    ok, pruned = P.calc_pruned(str_attr, tree)
    coeffs = P.calc_coefficients(tree)
    if not ok:
        return None
    h = _hash_g1(gk["g1"], sk["gid"])
    egg_s = o.GT_ONE
    for name, label in pruned:
        sk_attr = next(x for x in sk["attr"] if x[0] == name)
        ct_attr = next(x for x in ct["c"] if x[0] == label)
        num = o.gt_mul(ct_attr[1], o.pairing(h, ct_attr[3]))
        dem = o.pairing(sk_attr[1], ct_attr[2])
        coeff = next(c for l, c in coeffs if l == label)
        egg_s = o.gt_mul(egg_s, o.gt_pow(o.gt_mul(num, o.gt_inverse(dem)), _fr(coeff)))
    return o.gt_mul(ct["c_0"], o.gt_inverse(egg_s))


# ======================================================================================= AC17 KP
def ac17_kp_keygen(msk_bytes, policy, language, rnd):
    """ac17/mod.rs:439-546.  draws: r0, r1, sigma'[0..c-2], then sigma_attr per row.  Note the
    reference's `_temp` (declared at :491, before the `_j` loop) accumulates ACROSS j -- reproduced."""
    g, h = msk_bytes[:64], msk_bytes[64:192]
    g_k = [msk_bytes[192 + 64 * i:256 + 64 * i] for i in range(3)]
    a = [_int(msk_bytes[384 + 32 * i:416 + 32 * i]) for i in range(2)]
    b = [_int(msk_bytes[448 + 32 * i:480 + 32 * i]) for i in range(2)]
    tree = P.parse(policy, language)
    m, pi, c = P.calculate_msp(tree)
    r = [next(rnd) % R for _ in range(2)]
    br = [b[0] * r[0] % R, b[1] * r[1] % R, (r[0] + r[1]) % R]
    k_0 = [o.g2_mul(h, _fr(x)) for x in br]
    sigma_p = [next(rnd) % R for _ in range(c - 1)]
    k = []
    for i in range(len(m)):
        key = []
        sigma_attr = next(rnd) % R
        for t in range(2):
            prod = b"\0" * 64
            a_t = pow(a[t], -1, R)
            for l in range(3):
                prod = o.g1_add(prod, o.g1_mul(_hash_g1(g, "%s%d%d" % (pi[i], l, t)), _fr(br[l] * a_t)))
            prod = o.g1_add(prod, o.g1_mul(g, _fr(sigma_attr * a_t)))
            if m[i][0] == 1:
                prod = o.g1_add(prod, g_k[t])
            elif m[i][0] == -1:
                prod = o.g1_add(prod, o.g1_neg(g_k[t]))
            temp = b"\0" * 64
            for j in range(1, c):
                for l in range(3):
                    temp = o.g1_add(temp, o.g1_mul(_hash_g1(g, "0%d%d%d" % (j, l, t)), _fr(br[l] * a_t)))
                temp = o.g1_add(temp, o.g1_mul(g, _fr(-sigma_p[j - 1])))
                if m[i][j] == 1:
                    prod = o.g1_add(prod, temp)
                elif m[i][j] == -1:
                    prod = o.g1_add(prod, o.g1_neg(temp))
            key.append(prod)
        sk3 = o.g1_mul(g, _fr(-sigma_attr))
        if m[i][0] == 1:
            sk3 = o.g1_add(sk3, g_k[2])
        elif m[i][0] == -1:
            sk3 = o.g1_add(sk3, o.g1_neg(g_k[2]))
        for j in range(1, c):
            if m[i][j] == 1:
                sk3 = o.g1_add(sk3, o.g1_mul(g, _fr(-sigma_p[j - 1])))
            elif m[i][j] == -1:
                sk3 = o.g1_add(sk3, o.g1_neg(o.g1_mul(g, _fr(-sigma_p[j - 1]))))
        key.append(sk3)
        k.append((pi[i], key))
    return {"policy": (policy, language), "k_0": k_0, "k": k}


def ac17_kp_encrypt(pk_bytes, attributes, msg, rnd):
    """ac17/mod.rs:556-617.  draws: s0, s1."""
    g = pk_bytes[:64]
    h_a = [pk_bytes[64 + 128 * i:192 + 128 * i] for i in range(3)]
    e = [pk_bytes[448 + 384 * i:832 + 384 * i] for i in range(2)]
    s = [next(rnd) % R for _ in range(2)]
    c_0 = [o.g2_mul(h_a[0], _fr(s[0])), o.g2_mul(h_a[1], _fr(s[1])), o.g2_mul(h_a[2], _fr(s[0] + s[1]))]
    c = []
    for attr in attributes:
        ct = []
        for l in range(3):
            prod = b"\0" * 64
            for t in range(2):
                prod = o.g1_add(prod, o.g1_mul(_hash_g1(g, "%s%d%d" % (attr, l, t)), _fr(s[t])))
            ct.append(prod)
        c.append((attr, ct))
    c_p = o.gt_mul(o.gt_mul(o.gt_pow(e[0], _fr(s[0])), o.gt_pow(e[1], _fr(s[1]))), msg)
    return {"attr": list(attributes), "c_0": c_0, "c": c, "c_p": c_p}


def ac17_kp_decrypt(sk, ct):
    """ac17/mod.rs:625-680.  Returns the Gt `_msg` (None where rabe returns Err)."""
    tree = P.parse(*sk["policy"])
    if not P.traverse_policy(ct["attr"], tree):
        return None
    ok, lst = P.calc_pruned(ct["attr"], tree)
    if not ok:
        return None
    prod1, prod2 = o.GT_ONE, o.GT_ONE
    for i in range(3):
        prod_h, prod_g = b"\0" * 64, b"\0" * 64
        for cur, _ in lst:
            for name, v in ct["c"]:
                if name == cur:
                    prod_g = o.g1_add(prod_g, v[i])
            for name, v in sk["k"]:
                if name == cur:
                    prod_h = o.g1_add(prod_h, v[i])
        prod1 = o.gt_mul(prod1, o.pairing(prod_h, ct["c_0"][i]))
        prod2 = o.gt_mul(prod2, o.pairing(prod_g, sk["k_0"][i]))
    return o.gt_mul(ct["c_p"], o.gt_mul(prod2, o.gt_inverse(prod1)))


# ======================================================================================= GHW11
def ghw11_setup(rnd):
    """ghw11/mod.rs:92-111.  draws: g1, g2, a, alpha."""
    g1 = o.g1_mul(o.g1_generator(), _fr(next(rnd)))
    g2 = o.g2_mul(o.g2_generator(), _fr(next(rnd)))
    a = next(rnd) % R
    alpha = next(rnd) % R
    pk = {"g1": g1, "g2": g2, "g1_a": o.g1_mul(g1, _fr(a)), "g2_a": o.g2_mul(g2, _fr(a)),
          "e_gg_alpha": o.gt_pow(o.pairing(g1, g2), _fr(alpha))}
    return pk, {"g2_alpha": o.g2_mul(g2, _fr(alpha)), "pk": pk}


def ghw11_keygen(pk, msk, attributes, rnd):
    """ghw11/mod.rs:121-151.  draws: r."""
    if len(attributes) == 0:
        return None
    r = next(rnd) % R
    return {"k": o.g2_add(msk["g2_alpha"], o.g2_mul(pk["g2_a"], _fr(r))), "l": o.g2_mul(pk["g2"], _fr(r)),
            "attr_key": [(j, o.g2_mul(_hash_g2(pk["g2"], j), _fr(r))) for j in attributes]}


def ghw11_tkgen(sk, rnd):
    """ghw11/mod.rs:156-179.  draws: z."""
    z = next(rnd) % R
    zi = _fr(pow(z, -1, R))
    return ({"k_z": o.g2_mul(sk["k"], zi), "l_z": o.g2_mul(sk["l"], zi), "attr_key_z": [(n, o.g2_mul(kx, zi)) for n, kx in sk["attr_key"]]},
            {"z": z})


def ghw11_encrypt(pk, policy, language, msg, rnd):
    """ghw11/mod.rs:190-224.  draws: secret, (msg,) the gen_shares coefficients, then t_i per share."""
    secret = next(rnd) % R
    tree = P.parse(policy, language)
    shares = P.gen_shares_policy(secret, tree, rnd)
    c = o.gt_mul(o.gt_pow(pk["e_gg_alpha"], _fr(secret)), msg)
    c1 = o.g1_mul(pk["g1"], _fr(secret))
    ci_di = []
    for node, val in shares:
        t_i = next(rnd) % R
        j = P.remove_index(node)
        ci_di.append((node, o.g1_add(o.g1_mul(pk["g1_a"], _fr(val)), o.g1_mul(_hash_g1(pk["g1"], j), _fr(-t_i))), o.g1_mul(pk["g1"], _fr(t_i))))
    return {"policy": (policy, language), "c": c, "c1": c1, "ci_di": ci_di}


def ghw11_transform(ct, tk):
    """ghw11/mod.rs:227-294.  Returns {"c", "t"} (or None where rabe returns Err)."""
    attr = [n for n, _ in tk["attr_key_z"]]
    tree = P.parse(*ct["policy"])
    if not P.traverse_policy(attr, tree):
        return None
    ok, pruned = P.calc_pruned(attr, tree)
    if not ok:
        return None
    coeffs = P.calc_coefficients(tree)
    t, ci_wi = o.GT_ONE, G1_ZERO
    for name, label in pruned:
        coeff = next(cv for l, cv in coeffs if l == label)
        kx = next(k for n, k in tk["attr_key_z"] if n == name)
        _, ci, di = next(x for x in ct["ci_di"] if x[0] == label)
        ci_wi = o.g1_add(ci_wi, o.g1_mul(ci, _fr(coeff)))
        t = o.gt_mul(t, o.pairing(o.g1_mul(di, _fr(coeff)), kx))
    t = o.gt_mul(t, o.pairing(ci_wi, tk["l_z"]))
    t = o.gt_mul(o.pairing(ct["c1"], tk["k_z"]), o.gt_inverse(t))
    return {"c": ct["c"], "t": t}


def ghw11_decrypt_out(pct, rk):
    """ghw11/mod.rs:297-305 up to the KEM: the Gt `msg`."""
    return o.gt_mul(pct["c"], o.gt_inverse(o.gt_pow(pct["t"], _fr(rk["z"]))))
