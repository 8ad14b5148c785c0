"""A second, independent restatement of the reference's strict-mode phase arithmetic in numpy float32
(vectorised), used to cross-check the C oracle.  Follows Duke/mfreconstruct.cpp:231-269."""
import math

import numpy as np

F32 = np.float32
PI = F32(3.1416)            # Duke/mfreconstruct.cpp:5


def wrapped_phase(a, b):
    """a = G4-G2, b = G1-G3 (int arrays). Returns (P float32, ok bool)."""
    a = np.asarray(a, np.int64)
    b = np.asarray(b, np.int64)
    P = np.zeros(a.shape, F32)
    ok = np.ones(a.shape, bool)
    bs = np.where(b == 0, 1, b)
    q = (np.abs(a) // np.abs(bs)) * np.sign(a) * np.sign(bs)          # C++ truncating division
    at = np.array([math.atan(float(v)) for v in q.ravel()], np.float64).astype(F32).reshape(q.shape)
    two_pi = F32(2.0) * PI
    gen = np.where(b < 0, at + PI, np.where((b > 0) & (a > 0), at + two_pi, at)).astype(F32)
    P = gen
    P = np.where((b == 0) & (a < 0), PI / F32(2), P)
    P = np.where((b == 0) & (a > 0), F32(3) * PI / F32(2), P)
    P = np.where((a == 0) & (b < 0), PI, P)
    P = np.where((a == 0) & (b > 0), F32(0), P)
    ok = ~((a == 0) & (b == 0))
    return P.astype(F32), ok


def get_phase(G):
    """G: int array [..., 12] -> (phase float32, ok)."""
    G = np.asarray(G, np.int64)
    Ps, oks = [], []
    for f in range(3):
        P, ok = wrapped_phase(G[..., 4 * f + 3] - G[..., 4 * f + 1], G[..., 4 * f + 0] - G[..., 4 * f + 2])
        Ps.append(P.astype(np.float64))
        oks.append(ok)
    two_pi = F32(2.0) * PI
    P12 = np.where(Ps[0] > Ps[1], Ps[0] - Ps[1], Ps[0] - Ps[1] + np.float64(two_pi)).astype(F32)
    P23 = np.where(Ps[1] > Ps[2], Ps[1] - Ps[2], Ps[1] - Ps[2] + np.float64(two_pi)).astype(F32)
    d = (P12 - P23).astype(F32)
    P123 = np.where(P12 > P23, d, (d + two_pi).astype(F32)).astype(F32)
    phase = ((P123 / two_pi).astype(F32) * F32(255)).astype(F32)
    return phase, oks[0] & oks[1] & oks[2]
