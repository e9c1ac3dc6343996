"""Prototype (big ints) of the homogeneous-projective Miller steps used by miller6.cuh, checked against the oracle pairing."""
import sys, os, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import bls_oracle as O
P = O.P
mul, add, sub, sqr, neg, muls = O.f2_mul, O.f2_add, O.f2_sub, O.f2_sqr, O.f2_neg, O.f2_muls
XI = O.XI
Z2 = O.F2_ZERO

def dbl(T, px, py, pz):
    X, Y, Z = T
    B = mul(Y, Y); C = mul(Z, Z); J = mul(X, X); XY = mul(X, Y); YZ = mul(Y, Z)
    E = muls(mul(XI, C), 12); F = muls(E, 3)
    X3 = mul(muls(XY, 2), sub(B, F))
    Y3 = sub(mul(add(B, F), add(B, F)), mul(muls(E, 12), E))
    Z3 = mul(muls(B, 8), YZ)
    c0 = muls(sub(B, E), pz)
    c2 = muls(muls(J, -3 % P), px)
    c3 = muls(muls(YZ, 2), py)
    return (c0, Z2, c2, c3, Z2, Z2), (X3, Y3, Z3)

def addq(T, Q, px, py, pz):
    X, Y, Z = T
    x2, y2 = Q
    u = sub(mul(y2, Z), Y); v = sub(mul(x2, Z), X)
    vv = mul(v, v); vvv = mul(v, vv); R_ = mul(vv, X); uu = mul(u, u)
    A = sub(sub(mul(uu, Z), vvv), muls(R_, 2))
    X3 = mul(v, A); Y3 = sub(mul(u, sub(R_, A)), mul(vvv, Y)); Z3 = mul(vvv, Z)
    c0 = muls(sub(mul(u, x2), mul(v, y2)), pz)
    c2 = muls(neg(u), px)
    c3 = muls(v, py)
    return (c0, Z2, c2, c3, Z2, Z2), (X3, Y3, Z3)

def miller(Pj, Q):
    Xp, Yp, Zp = Pj  # Jacobian G1
    px, py, pz = Xp * Zp % P, Yp, pow(Zp, 3, P)
    f = O.F12_ONE
    T = (Q[0], Q[1], O.F2_ONE)
    for bit in bin(O.X_ABS)[3:]:
        l, T = dbl(T, px, py, pz)
        f = O.f12_mul(O.f12_mul(f, f), l)
        if bit == "1":
            l, T = addq(T, Q, px, py, pz)
            f = O.f12_mul(f, l)
    return O.f12_conj6(f)

rnd = random.Random(3)
p = O.g1_mul(O.G1_GEN, rnd.randrange(O.R)); q = O.g2_mul(O.G2_GEN, rnd.randrange(O.R))
z = rnd.randrange(1, P)
pj = (p[0] * z * z % P, p[1] * z * z * z % P, z)
assert O.final_exponentiation(miller(pj, q)) == O.pairing(p, q)
assert O.final_exponentiation(miller((p[0], p[1], 1), q)) == O.pairing(p, q)
print("ok")
