"""Oracle: explicit central-difference loop with mass-proportional damping and the
power-iteration estimate of the largest eigenfrequency.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
Follows examples/shells/dynamics/homogeneous/explicit/plate_expl_examples.jl:61-94
(`_cd_loop!`) and .../spherical_cap/spherical_cap_expl_examples.jl:160-173.
"""
from __future__ import annotations

import numpy as np


def cd_loop(Mdiag, K, ksi_2omegad, U0, V0, nsteps, dt, force, peek=None):
    """Central differences: `K` is a scipy CSR matrix (the free-free block), `Mdiag`
    the lumped-mass diagonal, C = (ksi*2*omegad) * diag(M), `force(t) -> F`.
    plate_expl_examples.jl:61-94 (statement order preserved)."""
    C = ksi_2omegad * Mdiag
    invMC = 1.0 / (Mdiag + (dt / 2) * C)
    dt2_2 = (dt**2) / 2
    dt_2 = dt / 2
    t = 0.0
    U = U0.copy()
    V = V0.copy()
    A = invMC * force(t)
    if peek:
        peek(0, U, V, t)
    for step in range(1, nsteps + 1):
        t = t + dt
        U += dt * V + dt2_2 * A
        F = force(t).copy()
        E = K @ U
        F -= E + C * (V + dt_2 * A)
        V += dt_2 * A
        A = invMC * F
        V += dt_2 * A
        if peek:
            peek(step, U, V, t)
    return U, V, A


def pwr_largest(K, Mdiag, maxit=30, seed=0):
    """Power iteration for the largest omega^2 of K x = omega^2 M x (lumped M);
    GEPHelpers.pwr_largest as used at spherical_cap_expl_examples.jl:166."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1, 1, Mdiag.shape[0])
    lam = 0.0
    for _ in range(maxit):
        y = (K @ x) / Mdiag
        lam = float(np.dot(x * Mdiag, y) / np.dot(x * Mdiag, x))
        x = y / np.linalg.norm(y)
    return lam
