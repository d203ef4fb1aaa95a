"""TEST INFRASTRUCTURE ONLY -- independent pure-Python BN254 known-answer generator.

This file is part of the oracle (see oracle/README.md).  Only tests/, bench.py's
cpu_baseline leg and __graft_entry__.smoke() may import it; the product package
`rabe_b200` never does.

It is a *second*, structurally different statement of the arithmetic the C++
oracle (oracle/bn254_oracle.cpp) implements, used to pin that oracle:

  * Fp12 is the flat extension Fp[w]/(w^12 - 18 w^6 + 82) (w^6 = 9+i), not a
    2-3-2 tower; points of the twist are mapped into E(Fp12) and the Miller loop
    uses affine chord/tangent lines over Fp12 (textbook optimal ate), not
    projective twist formulas with sparse multiplications.
  * The final exponentiation is one big modular power by LAMBDA (below), not an
    addition chain.

rabe delegates all of this to the external crate `rabe-bn 0.4.23`
(/root/reference/Cargo.toml:33), a fork of zcash `bn`, whose source is not in
/root/reference.  That lineage's hard part raises to

    LAMBDA = K * (p^4 - p^2 + 1)/r,   K = 2u(6u^2+3u+1)

(SURVEY.md section 8c), i.e. its pairing() is the reduced optimal-ate pairing to
the power K.  PARITY UNPINNED: no reference test asserts a Gt value, so the
exponent (and the element encodings) cannot be checked against rabe here.
"""
import hashlib

U = 4965661367192848881
P = 36 * U**4 + 36 * U**3 + 24 * U**2 + 6 * U + 1
R = 36 * U**4 + 36 * U**3 + 18 * U**2 + 6 * U + 1
ATE = 6 * U + 2
K_COFACTOR = 2 * U * (6 * U * U + 3 * U + 1)
HARD = (P**4 - P**2 + 1) // R
LAMBDA = (P**3 * (12 * U**3 + 6 * U**2 + 4 * U - 1) + P**2 * (12 * U**3 + 6 * U**2 + 6 * U)
          + P * (12 * U**3 + 6 * U**2 + 4 * U) + (12 * U**3 + 12 * U**2 + 6 * U + 1))
assert (P**4 - P**2 + 1) % R == 0 and LAMBDA == K_COFACTOR * HARD

G1_GEN = (1, 2)
G2_GEN = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
           11559732032986387107991004021392285783925812861821192530917403151452391805634),
          (8495653923123431417604973247489272438418190587263600148770280649306958101930,
           4082367875863433681332203403145435568316851327593401208105741076214120093531))


# ----------------------------------------------------------------- Fp2 = Fp[i]/(i^2+1)
def f2_add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
def f2_sub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
def f2_neg(a): return (-a[0] % P, -a[1] % P)
def f2_mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
def f2_inv(a):
    n = pow(a[0] * a[0] + a[1] * a[1], -1, P)
    return (a[0] * n % P, -a[1] * n % P)
F2_ZERO, F2_ONE = (0, 0), (1, 0)
XI = (9, 1)
B2 = f2_mul((3, 0), f2_inv(XI))          # twist curve coefficient 3/(9+i)


# ----------------------------------------------------------------- G1 / G2 affine (None = infinity)
def g1_add(a, b):
    if a is None: return b
    if b is None: return a
    if a[0] == b[0]:
        if (a[1] + b[1]) % P == 0: return None
        l = 3 * a[0] * a[0] * pow(2 * a[1], -1, P) % P
    else:
        l = (b[1] - a[1]) * pow(b[0] - a[0], -1, P) % P
    x = (l * l - a[0] - b[0]) % P
    return (x, (l * (a[0] - x) - a[1]) % P)

def g1_neg(a): return None if a is None else (a[0], -a[1] % P)

def g1_mul(a, k):
    k %= R
    acc = None
    while k:
        if k & 1: acc = g1_add(acc, a)
        a = g1_add(a, a); k >>= 1
    return acc

def g2_add(a, b):
    if a is None: return b
    if b is None: return a
    if a[0] == b[0]:
        if f2_add(a[1], b[1]) == F2_ZERO: return None
        l = f2_mul(f2_mul((3, 0), f2_mul(a[0], a[0])), f2_inv(f2_add(a[1], a[1])))
    else:
        l = f2_mul(f2_sub(b[1], a[1]), f2_inv(f2_sub(b[0], a[0])))
    x = f2_sub(f2_sub(f2_mul(l, l), a[0]), b[0])
    return (x, f2_sub(f2_mul(l, f2_sub(a[0], x)), a[1]))

def g2_neg(a): return None if a is None else (a[0], f2_neg(a[1]))

def g2_mul(a, k):
    k %= R
    acc = None
    while k:
        if k & 1: acc = g2_add(acc, a)
        a = g2_add(a, a); k >>= 1
    return acc

def g1_on_curve(a): return a is None or (a[1] * a[1] - a[0]**3 - 3) % P == 0
def g2_on_curve(a):
    return a is None or f2_sub(f2_mul(a[1], a[1]), f2_add(f2_mul(a[0], f2_mul(a[0], a[0])), B2)) == F2_ZERO


# ----------------------------------------------------------------- flat Fp12 = Fp[w]/(w^12 - 18w^6 + 82)
def f12_mul(a, b):
    t = [0] * 23
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                t[i + j] += x * y
    for k in range(22, 11, -1):            # w^12 = 18 w^6 - 82
        c = t[k]
        if c:
            t[k - 6] += 18 * c
            t[k - 12] -= 82 * c
    return [x % P for x in t[:12]]

F12_ONE = [1] + [0] * 11

def f12_pow(a, e):
    acc = F12_ONE
    for bit in bin(e)[2:]:
        acc = f12_mul(acc, acc)
        if bit == '1': acc = f12_mul(acc, a)
    return acc

def f12_inv(a):
    # a^(p^12-2) is far too slow; solve the 12x12 linear system a*x = 1 instead.
    n = 12
    M = []
    for col in range(n):
        e = [0] * n; e[col] = 1
        M.append(f12_mul(a, e))
    # M[col][row]; build augmented rows
    A = [[M[c][r] for c in range(n)] + [1 if r == 0 else 0] for r in range(n)]
    for c in range(n):
        piv = next(r for r in range(c, n) if A[r][c] % P)
        A[c], A[piv] = A[piv], A[c]
        inv = pow(A[c][c], -1, P)
        A[c] = [x * inv % P for x in A[c]]
        for r in range(n):
            if r != c and A[r][c]:
                f = A[r][c]
                A[r] = [(x - f * y) % P for x, y in zip(A[r], A[c])]
    return [A[r][n] for r in range(n)]

def f2_to_f12(a, k):
    """(a0 + a1 i) * w^k with i = w^6 - 9, 0 <= k < 6."""
    out = [0] * 12
    out[k] = (a[0] - 9 * a[1]) % P
    out[k + 6] = a[1]
    return out

def tower_to_flat(c):
    """c = 12 Fp coefficients in tower order
    [c0.c0.(re,im), c0.c1.(re,im), c0.c2.(re,im), c1.c0.(re,im), c1.c1.(re,im), c1.c2.(re,im)]
    for Fq12 = Fq6[w]/(w^2 - v), Fq6 = Fq2[v]/(v^3 - xi): element = sum_k z_k w^k with
    z_0=c0.c0, z_1=c1.c0, z_2=c0.c1, z_3=c1.c1, z_4=c0.c2, z_5=c1.c2."""
    z = {0: (c[0], c[1]), 2: (c[2], c[3]), 4: (c[4], c[5]), 1: (c[6], c[7]), 3: (c[8], c[9]), 5: (c[10], c[11])}
    out = [0] * 12
    for k, a in z.items():
        t = f2_to_f12(a, k)
        out = [(x + y) % P for x, y in zip(out, t)]
    return out

def flat_to_tower(f):
    z = {}
    for k in range(6):
        im = f[k + 6] % P
        re = (f[k] + 9 * im) % P
        z[k] = (re, im)
    order = [0, 2, 4, 1, 3, 5]
    out = []
    for k in order:
        out += [z[k][0], z[k][1]]
    return out


# ----------------------------------------------------------------- textbook optimal ate over E(Fp12)
def _untwist(q):
    """psi: E'(Fp2) -> E(Fp12), (x, y) -> (x w^2, y w^3)."""
    return (f2_to_f12(q[0], 2), f2_to_f12(q[1], 3))

def _e12_sub(a, b): return [(x - y) % P for x, y in zip(a, b)]
def _e12_add(a, b): return [(x + y) % P for x, y in zip(a, b)]

def _line(a, b, t):
    """Value at t of the line through a and b (tangent if a == b) on y^2 = x^3 + 3 over Fp12."""
    (x1, y1), (x2, y2), (xt, yt) = a, b, t
    if x1 != x2:
        m = f12_mul(_e12_sub(y2, y1), f12_inv(_e12_sub(x2, x1)))
    elif y1 == y2:
        m = f12_mul([3 * c % P for c in f12_mul(x1, x1)], f12_inv([2 * c % P for c in y1]))
    else:
        return _e12_sub(xt, x1)
    return _e12_sub(f12_mul(m, _e12_sub(xt, x1)), _e12_sub(yt, y1))

def _e12_pt_add(a, b):
    (x1, y1), (x2, y2) = a, b
    if x1 == x2 and y1 == y2:
        m = f12_mul([3 * c % P for c in f12_mul(x1, x1)], f12_inv([2 * c % P for c in y1]))
    else:
        m = f12_mul(_e12_sub(y2, y1), f12_inv(_e12_sub(x2, x1)))
    x3 = _e12_sub(_e12_sub(f12_mul(m, m), x1), x2)
    y3 = _e12_sub(f12_mul(m, _e12_sub(x1, x3)), y1)
    return (x3, y3)

def _frob_pt(a):
    return (f12_pow(a[0], P), f12_pow(a[1], P))

def miller_textbook(p1, q2):
    """f_{6u+2,Q}(P) * l_{[6u+2]Q,pi(Q)}(P) * l_{.., -pi^2(Q)}(P), flat Fp12."""
    Pt = ([p1[0]] + [0] * 11, [p1[1]] + [0] * 11)
    Q = _untwist(q2)
    T = Q
    f = F12_ONE
    for bit in bin(ATE)[3:]:
        f = f12_mul(f12_mul(f, f), _line(T, T, Pt))
        T = _e12_pt_add(T, T)
        if bit == '1':
            f = f12_mul(f, _line(T, Q, Pt))
            T = _e12_pt_add(T, Q)
    Q1 = _frob_pt(Q)
    Q2 = _frob_pt(Q1)
    nQ2 = (Q2[0], [-c % P for c in Q2[1]])
    f = f12_mul(f, _line(T, Q1, Pt))
    T = _e12_pt_add(T, Q1)
    f = f12_mul(f, _line(T, nQ2, Pt))
    return f

def pairing_textbook(p1, q2):
    """Reduced optimal ate pairing e(P,Q) = miller^((p^12-1)/r), flat Fp12."""
    if p1 is None or q2 is None: return F12_ONE
    return f12_pow(miller_textbook(p1, q2), (P**12 - 1) // R)

def pairing_lineage(p1, q2):
    """What the zcash-bn lineage final exponentiation yields: e(P,Q)^K, as 12 tower coefficients."""
    if p1 is None or q2 is None: return flat_to_tower(F12_ONE)
    f = miller_textbook(p1, q2)
    easy = (P**6 - 1) * (P**2 + 1)
    return flat_to_tower(f12_pow(f12_pow(f, easy), LAMBDA))

def gt_pow_tower(c, e):
    return flat_to_tower(f12_pow(tower_to_flat(c), e))

def gt_mul_tower(a, b):
    return flat_to_tower(f12_mul(tower_to_flat(a), tower_to_flat(b)))


# ----------------------------------------------------------------- hashing (src/utils/hash/mod.rs:10-32)
def sha3_fr(s: str) -> int:
    """Fr::from_slice(SHA3-256(utf8)) -- big-endian 256-bit integer reduced mod r."""
    return int.from_bytes(hashlib.sha3_256(s.encode()).digest(), 'big') % R
