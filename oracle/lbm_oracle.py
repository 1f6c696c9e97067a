"""CPU oracle for the D2Q9 lid-driven-cavity collide-and-stream step.

TEST INFRASTRUCTURE ONLY.  Nothing under ``latticeboltzmannsimulations_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` do, and there only as the checker / the timed CPU baseline.

It is a plain NumPy fp64 restatement of the reference's arithmetic (citations are file:line into
the upstream repo RaghuvirJonnagiri/LatticeBoltzmannSimulations):

* semantics ``"A"``  -- the NumPy solver, ``MRT.py:286-453`` (SRT collision, slice streaming with the
  exclusive ``xsize_max`` bound, "= feq" left wall).  Pinned: ``tests/test_oracle.py`` and
  ``tests/golden/make_golden.py`` run the real ``MRT.py`` under import stubs and require a 0.0 difference.
* semantics ``"C"``  -- the PyCUDA solver, ``MRT_GPU.py:336-703`` (``funRT`` in its SRT ``:338-422``,
  TRT ``:426-531`` and MRT ``:535-662`` forms, optional Smagorinsky ``:570-589``, then ``funBC``
  ``:664-699``), evaluated in fp64.  Pinned in two parts: moments, overrides, equilibrium, SRT collision
  and push streaming agree with the *compiled* reference ``functions.allfunc`` (``functions.pyx:45-222``)
  to 2.2e-16 on one step (same tests).  The MRT relaxation and ``funBC`` cannot be executed on a CPU (PyCUDA kernel
  strings); they are pinned on the GPU box instead: ``oracle/build_ref_kernels.py`` compiles the reference's own
  ``funRT`` (SRT/TRT/MRT, +-Smagorinsky) and ``funBC`` strings with nvcc and
  ``tests/test_gpu_reference_kernels.py`` runs them on the B200 against this oracle -- agreement at fp32 round-off
  (rho 1e-6..5e-6, u 4e-6..1.8e-5 of uLB; the reference kernels are fp32), where any semantic slip would show at
  1e-3..1e-2.  Every line below that restates those pieces carries its citation.

``step_C`` is the literal two-kernel push form.  ``PullState``/``step_C_pull`` is the same update
re-expressed as the single fused *pull* pass that the CUDA kernel implements (state kept between steps
= post-collision populations, wall rule applied on read, lid density and the four doubly-orphaned
corner populations carried in side buffers); ``tests/test_oracle.py`` asserts the two forms agree.

Array convention (all variants, ``MRT.py:252-261``): arrays are ``[k, x, y]``; ``y == 0`` is the
moving lid, ``y == ny-1`` the bottom wall; a population with ``c_y = +1`` moves to ``y-1``.
"""
from __future__ import annotations

import dataclasses
import os
from typing import Optional, Tuple

import numpy as np

# --------------------------------------------------------------------------------------------
# Lattice constants -- MRT.py:138-160 (same tables re-typed inside every kernel, MRT_GPU.py:363-365)
# --------------------------------------------------------------------------------------------
Q = 9
C = np.array([[0, 0], [1, 0], [0, 1], [-1, 0], [0, -1], [1, 1], [-1, 1], [-1, -1], [1, -1]])
T = 1.0 / 36.0 * np.ones(Q)          # MRT.py:144-146
T[1:5] = 1.0 / 9.0
T[0] = 4.0 / 9.0
BOUNCE = [0, 3, 4, 1, 2, 7, 8, 5, 6]  # MRT.py:152
RIGHT = np.array([1, 5, 8])           # c_x > 0   MRT.py:155-160
LEFT = np.array([3, 6, 7])            # c_x < 0
TOP = np.array([2, 5, 6])             # c_y > 0
BOT = np.array([4, 7, 8])             # c_y < 0
CENTH = np.array([0, 1, 3])           # c_y == 0

# Gram-Schmidt moment basis and its inverse -- MRT.py:163-183 == MRT_GPU.py:593-612
M_GS = np.array([
    [1, 1, 1, 1, 1, 1, 1, 1, 1],
    [-4, -1, -1, -1, -1, 2, 2, 2, 2],
    [4, -2, -2, -2, -2, 1, 1, 1, 1],
    [0, 1, 0, -1, 0, 1, -1, -1, 1],
    [0, -2, 0, 2, 0, 1, -1, -1, 1],
    [0, 0, 1, 0, -1, 1, 1, -1, -1],
    [0, 0, -2, 0, 2, 1, 1, -1, -1],
    [0, 1, -1, 1, -1, 0, 0, 0, 0],
    [0, 0, 0, 0, 0, 1, -1, 1, -1]], dtype=np.float64)
M_GS_INV = np.array([
    [1.0 / 9, -1.0 / 9, 1.0 / 9, 0, 0, 0, 0, 0, 0],
    [1.0 / 9, -1.0 / 36, -1.0 / 18, 1.0 / 6, -1.0 / 6, 0, 0, 1.0 / 4, 0],
    [1.0 / 9, -1.0 / 36, -1.0 / 18, 0, 0, 1.0 / 6, -1.0 / 6, -1.0 / 4, 0],
    [1.0 / 9, -1.0 / 36, -1.0 / 18, -1.0 / 6, 1.0 / 6, 0, 0, 1.0 / 4, 0],
    [1.0 / 9, -1.0 / 36, -1.0 / 18, 0, 0, -1.0 / 6, 1.0 / 6, -1.0 / 4, 0],
    [1.0 / 9, 1.0 / 18, 1.0 / 36, 1.0 / 6, 1.0 / 12, 1.0 / 6, 1.0 / 12, 0, 1.0 / 4],
    [1.0 / 9, 1.0 / 18, 1.0 / 36, -1.0 / 6, -1.0 / 12, 1.0 / 6, 1.0 / 12, 0, -1.0 / 4],
    [1.0 / 9, 1.0 / 18, 1.0 / 36, -1.0 / 6, -1.0 / 12, -1.0 / 6, -1.0 / 12, 0, 1.0 / 4],
    [1.0 / 9, 1.0 / 18, 1.0 / 36, 1.0 / 6, 1.0 / 12, -1.0 / 6, -1.0 / 12, 0, -1.0 / 4]], dtype=np.float64)


@dataclasses.dataclass
class Params:
    """Run parameters -- MRT.py:41-75, MRT_GPU.py:45-93, functions.pyx:38-43."""
    nx: int
    ny: int
    uLB: float = 0.08
    Re: float = 100.0
    collision: str = "MRT"        # 'SRT' | 'TRT' | 'MRT'   (MRT_GPU.py:48 `RT`)
    turb: int = 0                 # Smagorinsky switch       (MRT_GPU.py:49)
    omega_e: float = 1.0          # MRT_GPU.py:89
    omega_eps: float = 1.2        # MRT_GPU.py:90   (MRT.py:72 has 1.0; the GPU values are the live ones)
    omega_q: float = 1.2          # MRT_GPU.py:90
    delTRT: float = 1.0 / 3.5     # MRT_GPU.py:83
    omega: Optional[float] = None  # override; default derived from Re

    def __post_init__(self):
        if self.omega is None:
            self.omega = omega_from_re(self.uLB, self.Re, self.ny)

    @property
    def omegam(self) -> float:     # MRT_GPU.py:82-84
        omegap = self.omega
        return 1.0 / (0.5 + (self.delTRT / ((1 / omegap) - 0.5)))


def omega_from_re(uLB: float, Re: float, ysize: int) -> float:
    """nuLB = uLB*ysize/Re ; omega = 2/(6 nuLB + 1) -- MRT.py:53-55, functions.pyx:38-43, MRT_GPU.py:63-65."""
    nuLB = uLB * ysize / Re
    return 2.0 / (6. * nuLB + 1)


def equ(rho: np.ndarray, u: np.ndarray) -> np.ndarray:
    """Second-order equilibrium, operation order of MRT.py:213-231 (== functions.pyx:229-267)."""
    cu = [C[k, 0] * u[0] + C[k, 1] * u[1] for k in range(Q)]
    usqr = (u[0] * u[0] + u[1] * u[1])
    feq = np.empty((Q,) + rho.shape)
    for i in range(Q):
        feq[i] = rho * T[i] * (1. + 3.0 * cu[i] + 9 * 0.5 * cu[i] * cu[i] - 3.0 * 0.5 * usqr)
    return feq


def init_fields(nx: int, ny: int, uLB: float) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """rho = 1, u = (uLB, 0) on the whole row y == 0, fin = feq -- MRT.py:252-268, MRT_GPU.py:259-267."""
    rho = np.ones((nx, ny))
    vel = np.zeros((2, nx, ny))
    vel[0, :, 0] = uLB
    return rho, vel, equ(rho, vel)


# --------------------------------------------------------------------------------------------
# Semantics A : MRT.py:286-453, literal.
# --------------------------------------------------------------------------------------------
def step_A(fin: np.ndarray, p: Params) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """One iteration of the MRT.py time loop; ``fin`` is updated in place.  Returns (rho, u, feq)."""
    nx, ny = p.nx, p.ny
    xsize_max, ysize_max = nx - 1, ny - 1
    omega, uLB = p.omega, p.uLB
    c = C
    rho = np.sum(fin, axis=0)                                                        # MRT.py:292
    u = np.empty((2, nx, ny))
    u[0] = (c[0, 0] * fin[0] + c[1, 0] * fin[1] + c[2, 0] * fin[2] + c[3, 0] * fin[3] + c[4, 0] * fin[4]
            + c[5, 0] * fin[5] + c[6, 0] * fin[6] + c[7, 0] * fin[7] + c[8, 0] * fin[8]) / rho   # :320
    u[1] = (c[0, 1] * fin[0] + c[1, 1] * fin[1] + c[2, 1] * fin[2] + c[3, 1] * fin[3] + c[4, 1] * fin[4]
            + c[5, 1] * fin[5] + c[6, 1] * fin[6] + c[7, 1] * fin[7] + c[8, 1] * fin[8]) / rho   # :321
    rho[:, 0] = np.sum(fin[CENTH, :, 0], axis=0) + 2. * np.sum(fin[TOP, :, 0], axis=0)  # :337
    u[:, 0, 1:] = 0; u[:, xsize_max, 1:] = 0; u[:, :, ysize_max] = 0                 # :341
    u[0, :, 0] = uLB; u[1, :, 0] = 0                                                 # :342
    feq = equ(rho, u)                                                                # :344
    fpost = fin - omega * (fin - feq)                                                # :396 (numexpr)
    # streaming :404-414 -- xsize_max/ysize_max used as *exclusive* slice bounds (reference behaviour)
    fin[0, :, :] = fpost[0, :, :]
    fin[1, 1:xsize_max, :] = fpost[1, 0:xsize_max - 1, :]
    fin[2, :, 0:ysize_max - 1] = fpost[2, :, 1:ysize_max]
    fin[3, 0:xsize_max - 1, :] = fpost[3, 1:xsize_max, :]
    fin[4, :, 1:ysize_max] = fpost[4, :, 0:ysize_max - 1]
    fin[5, 1:xsize_max, 0:ysize_max - 1] = fpost[5, 0:xsize_max - 1, 1:ysize_max]
    fin[6, 0:xsize_max - 1, 0:ysize_max - 1] = fpost[6, 1:xsize_max, 1:ysize_max]
    fin[7, 0:xsize_max - 1, 1:ysize_max] = fpost[7, 1:xsize_max, 0:ysize_max - 1]
    fin[8, 1:xsize_max, 1:ysize_max] = fpost[8, 0:xsize_max - 1, 0:ysize_max - 1]
    # boundary :450-453
    fin[RIGHT, 0, :] = feq[RIGHT, 0, :]
    fin[LEFT, xsize_max, :] = - feq[RIGHT, xsize_max, :] + (feq[LEFT, xsize_max, :] + fin[RIGHT, xsize_max, :])
    fin[TOP, :, ysize_max] = - feq[BOT, :, ysize_max] + (feq[TOP, :, ysize_max] + fin[BOT, :, ysize_max])
    fin[BOT, :, 0] = - feq[TOP, :, 0] + (feq[BOT, :, 0] + fin[TOP, :, 0])
    return rho, u, feq


# --------------------------------------------------------------------------------------------
# Semantics C : MRT_GPU.py funRT + funBC, literal two-pass push form, fp64.
# --------------------------------------------------------------------------------------------
@dataclasses.dataclass
class StateC:
    """Device arrays of MRT_GPU.py:309-328, held in the host ``[k,x,y]`` convention."""
    fin: np.ndarray
    ftemp: np.ndarray
    feq: np.ndarray
    rho: np.ndarray
    u: np.ndarray
    taus: np.ndarray

    @staticmethod
    def initial(p: Params, fin0: Optional[np.ndarray] = None) -> "StateC":
        """MRT_GPU.py:259-267, 323-328: fin = ftemp = feq_g = feq(1,InitVel); u_g = 0; rho_g = 1; taus = 1/omega."""
        if fin0 is None:
            _, _, fin0 = init_fields(p.nx, p.ny, p.uLB)
        fin0 = np.array(fin0, dtype=np.float64)
        return StateC(fin=fin0.copy(), ftemp=fin0.copy(), feq=fin0.copy(),
                      rho=np.ones((p.nx, p.ny)), u=np.zeros((2, p.nx, p.ny)),
                      taus=np.full((p.nx, p.ny), 1.0 / p.omega))


def _lsum(terms):
    """Strict left-to-right sum, the order the kernels write (MRT_GPU.py:615)."""
    acc = terms[0]
    for t_ in terms[1:]:
        acc = acc + t_
    return acc


def _smagorinsky_tau(f, feq_prev, rho_prev, tau0):
    """MRT_GPU.py:570-589.  Cs2 is hard-overridden to 0.025 (:578); the Van-Driest lines are dead."""
    Cs2 = 0.025
    product1 = 0.0
    product2 = 0.0
    for k in range(Q):
        cc = C[k, 0] * C[k, 1]
        product1 = cc * f[k] + product1
        product2 = cc * feq_prev[k] + product2
    Qmf = product1 - product2
    tau = 0.5 * (tau0 + np.sqrt((tau0 * tau0 + (18 * 1.4142 * Cs2 * np.abs(Qmf)) / rho_prev)))
    return tau


def _moments_overrides(f, p: Params):
    """rho, u with the wall/lid overrides -- MRT_GPU.py:615-631 (== MRT.py:292,320-342, functions.pyx:73-81)."""
    nx, ny = p.nx, p.ny
    c = C
    rho_l = _lsum([f[k] for k in range(Q)])
    ux = _lsum([c[k, 0] * f[k] for k in range(Q)]) / rho_l
    uy = _lsum([c[k, 1] * f[k] for k in range(Q)]) / rho_l
    ux[0, :] = 0; uy[0, :] = 0
    ux[nx - 1, :] = 0; uy[nx - 1, :] = 0
    ux[:, ny - 1] = 0; uy[:, ny - 1] = 0
    # lid test comes last and wins, corners included (MRT_GPU.py:626-631)
    rho_l = rho_l.copy()
    rho_l[:, 0] = f[0, :, 0] + f[1, :, 0] + f[3, :, 0] + 2 * (f[2, :, 0] + f[5, :, 0] + f[6, :, 0])
    ux[:, 0] = p.uLB
    uy[:, 0] = 0
    return rho_l, ux, uy


def _feq_kernel(rho_l, ux, uy):
    """In-kernel equilibrium, MRT_GPU.py:649-652 (same expression as ``equ``)."""
    usqr = ux * ux + uy * uy
    feq = np.empty((Q,) + rho_l.shape)
    for k in range(Q):
        cu = (C[k, 0] * ux + C[k, 1] * uy)
        feq[k] = rho_l * T[k] * (1. + 3.0 * cu + 9 * 0.5 * cu * cu - 3.0 * 0.5 * usqr)
    return feq


def _collide(f, rho_l, feq, p: Params, omega_nu):
    """Post-collision populations for the three `RT` choices; ``omega_nu`` may be a per-node array (turb)."""
    if p.collision == "SRT":                                   # MRT_GPU.py:413
        return f - omega_nu * (f - feq)
    if p.collision == "TRT":                                   # MRT_GPU.py:455-462, 514-527
        b = BOUNCE
        fplus = np.empty_like(f); fminus = np.empty_like(f)
        feplus = np.empty_like(f); feminus = np.empty_like(f)
        for a_, o_ in ((2, 4), (5, 7), (6, 8), (1, 3)):
            fplus[a_] = 0.5 * (f[a_] + f[o_]); fplus[o_] = fplus[a_]
            fminus[a_] = 0.5 * (f[a_] - f[o_]); fminus[o_] = -fminus[a_]
            feplus[a_] = 0.5 * (feq[a_] + feq[o_]); feplus[o_] = feplus[a_]
            feminus[a_] = 0.5 * (feq[a_] - feq[o_]); feminus[o_] = -feminus[a_]
        fplus[0] = f[0]; fminus[0] = 0
        feplus[0] = feq[0]; feminus[0] = 0
        del b
        return f - omega_nu * (fplus - feplus) - p.omegam * (fminus - feminus)
    if p.collision == "MRT":                                   # MRT_GPU.py:633-648, 655
        m = [_lsum([M_GS[k, j] * f[j] for j in range(Q)]) for k in range(Q)]
        jx = m[3]; jy = m[5]
        meq = [None] * Q
        meq[0] = rho_l
        meq[1] = -2.0 * rho_l + 3.0 * (jx * jx + jy * jy)
        meq[2] = - 3.0 * (jx * jx + jy * jy) + rho_l + 9.0 * (jx * jx * jy * jy)
        meq[4] = - jx + 3.0 * (jx * jx * jx)
        meq[6] = - jy + 3.0 * (jy * jy * jy)
        meq[7] = jx * jx - jy * jy
        meq[8] = jx * jy
        meq[3] = m[3]; meq[5] = m[5]
        omega_vec = [0.0, p.omega_e, p.omega_eps, 0.0, p.omega_q, 0.0, p.omega_q, omega_nu, omega_nu]
        m = [m[k] - omega_vec[k] * (m[k] - meq[k]) for k in range(Q)]
        out = np.empty_like(f)
        for k in range(Q):
            out[k] = _lsum([M_GS_INV[k, j] * m[j] for j in range(Q)])
        return out
    raise ValueError(p.collision)


def funRT(st: StateC, p: Params) -> None:
    """MRT_GPU.py ``funRT`` (SRT :338-422 / TRT :426-531 / MRT :535-662): moments, collide, push into ftemp."""
    nx, ny = p.nx, p.ny
    f = st.fin
    omega_nu = p.omega
    if p.turb == 1:
        tau = _smagorinsky_tau(f, st.feq, st.rho, 1.0 / p.omega)   # uses feq_g / rho_g of the PREVIOUS step
        omega_nu = 1.0 / tau
        st.taus = tau
    rho_l, ux, uy = _moments_overrides(f, p)
    st.rho = rho_l
    st.u = np.stack([ux, uy])
    feq = _feq_kernel(rho_l, ux, uy)
    st.feq = feq
    fpost = _collide(f, rho_l, feq, p, omega_nu)
    # push, bounds-checked; slots that receive nothing keep their previous ftemp content (:654-656)
    for k in range(Q):
        cx, cy = int(C[k, 0]), int(C[k, 1])
        xs = slice(max(0, -cx), nx - max(0, cx))           # source x with 0 <= x+cx < nx
        ys = slice(max(0, cy), ny - max(0, -cy))           # source y with 0 <= y-cy < ny
        xd = slice(xs.start + cx, xs.stop + cx)
        yd = slice(ys.start - cy, ys.stop - cy)
        st.ftemp[k, xd, yd] = fpost[k, xs, ys]


def funBC(st: StateC, p: Params) -> None:
    """MRT_GPU.py:664-699: NEBB, x-block then y-block per node, then fin = ftemp."""
    nx, ny = p.nx, p.ny
    ft, fe = st.ftemp, st.feq
    x = 0                                                   # :674-677
    ft[1, x, :] = fe[1, x, :] - fe[3, x, :] + ft[3, x, :]
    ft[5, x, :] = fe[5, x, :] - fe[7, x, :] + ft[7, x, :]
    ft[8, x, :] = fe[8, x, :] - fe[6, x, :] + ft[6, x, :]
    x = nx - 1                                              # :678-682
    ft[3, x, :] = -fe[1, x, :] + fe[3, x, :] + ft[1, x, :]
    ft[6, x, :] = -fe[8, x, :] + fe[6, x, :] + ft[8, x, :]
    ft[7, x, :] = -fe[5, x, :] + fe[7, x, :] + ft[5, x, :]
    y = ny - 1                                              # :684-687
    ft[2, :, y] = -fe[4, :, y] + fe[2, :, y] + ft[4, :, y]
    ft[5, :, y] = -fe[7, :, y] + fe[5, :, y] + ft[7, :, y]
    ft[6, :, y] = -fe[8, :, y] + fe[6, :, y] + ft[8, :, y]
    y = 0                                                   # :688-692
    ft[4, :, y] = -fe[2, :, y] + fe[4, :, y] + ft[2, :, y]
    ft[7, :, y] = -fe[5, :, y] + fe[7, :, y] + ft[5, :, y]
    ft[8, :, y] = -fe[6, :, y] + fe[8, :, y] + ft[6, :, y]
    st.fin = ft.copy()                                      # :694-696


def step_C(st: StateC, p: Params) -> None:
    """One iteration of the MRT_GPU.py time loop (:724, :732)."""
    funRT(st, p)
    funBC(st, p)


# --------------------------------------------------------------------------------------------
# Semantics C re-expressed as ONE fused pull pass (the specification of the CUDA kernel).
# --------------------------------------------------------------------------------------------
# corner-carry slots: population that is orphaned twice at each corner (see SURVEY.md 8a / DESIGN.md)
CARRY_TL, CARRY_TR, CARRY_BL, CARRY_BR = 0, 1, 2, 3


@dataclasses.dataclass
class PullState:
    """State the fused kernel keeps between steps.

    ``g``       populations [9,nx,ny]: pre-collision ``fin`` if ``kind == 'pre'`` (just uploaded),
                else post-collision f* of the previous launch.
    ``rho_lid`` lid density of row y == 0 computed by the previous launch (nx values).
    ``carry``   previous final value of f7@(0,0), f8@(nx-1,0), f6@(0,ny-1), f5@(nx-1,ny-1).
    ``pi_eq``/``rho_prev`` previous-step sum_k cx cy feq_k and rho (only used with turb == 1).
    """
    g: np.ndarray
    kind: str
    rho_lid: np.ndarray
    carry: np.ndarray
    pi_eq: np.ndarray
    rho_prev: np.ndarray
    rho: Optional[np.ndarray] = None     # lagged outputs, as the reference stores them
    u: Optional[np.ndarray] = None

    @staticmethod
    def from_fin(fin0: np.ndarray, p: Params) -> "PullState":
        nx, ny = p.nx, p.ny
        fin0 = np.array(fin0, dtype=np.float64)
        carry = np.array([fin0[7, 0, 0], fin0[8, nx - 1, 0], fin0[6, 0, ny - 1], fin0[5, nx - 1, ny - 1]])
        pi = 0.0
        for k in range(Q):
            pi = (C[k, 0] * C[k, 1]) * fin0[k] + pi       # feq_g := fin at upload, MRT_GPU.py:325
        return PullState(g=fin0.copy(), kind="pre", rho_lid=np.zeros(nx), carry=carry,
                         pi_eq=pi, rho_prev=np.ones((nx, ny)))


def _gather_bc(ps: PullState, p: Params) -> Tuple[np.ndarray, np.ndarray]:
    """Pull streaming + wall rule: returns (fin entering this step, updated corner carry)."""
    nx, ny = p.nx, p.ny
    if ps.kind == "pre":
        return ps.g.copy(), ps.carry.copy()
    g = ps.g
    h = np.full_like(g, np.nan)                      # NaN marks orphan slots: must all be overwritten
    for k in range(Q):
        cx, cy = int(C[k, 0]), int(C[k, 1])
        xs = slice(max(0, -cx), nx - max(0, cx))
        ys = slice(max(0, cy), ny - max(0, -cy))
        xd = slice(xs.start + cx, xs.stop + cx)
        yd = slice(ys.start - cy, ys.stop - cy)
        h[k, xd, yd] = g[k, xs, ys]
    # doubly-orphaned corner slots hold last step's final value (single persistent ftemp in the reference)
    h[7, 0, 0] = ps.carry[CARRY_TL]
    h[8, nx - 1, 0] = ps.carry[CARRY_TR]
    h[6, 0, ny - 1] = ps.carry[CARRY_BL]
    h[5, nx - 1, ny - 1] = ps.carry[CARRY_BR]
    # feq of the node's own previous-step state: resting walls u = 0, lid row rho_lid / (uLB, 0)
    # -> only the lid row has non-zero feq_k - feq_opp(k); build it exactly as the kernel stored it.
    rho_w = np.ones((nx, ny))                        # value irrelevant where u == 0 (differences are exactly 0)
    ux = np.zeros((nx, ny)); uy = np.zeros((nx, ny))
    rho_w[:, 0] = ps.rho_lid
    ux[:, 0] = p.uLB
    fe = _feq_kernel(rho_w, ux, uy)
    x = 0
    h[1, x, :] = fe[1, x, :] - fe[3, x, :] + h[3, x, :]
    h[5, x, :] = fe[5, x, :] - fe[7, x, :] + h[7, x, :]
    h[8, x, :] = fe[8, x, :] - fe[6, x, :] + h[6, x, :]
    x = nx - 1
    h[3, x, :] = -fe[1, x, :] + fe[3, x, :] + h[1, x, :]
    h[6, x, :] = -fe[8, x, :] + fe[6, x, :] + h[8, x, :]
    h[7, x, :] = -fe[5, x, :] + fe[7, x, :] + h[5, x, :]
    y = ny - 1
    h[2, :, y] = -fe[4, :, y] + fe[2, :, y] + h[4, :, y]
    h[5, :, y] = -fe[7, :, y] + fe[5, :, y] + h[7, :, y]
    h[6, :, y] = -fe[8, :, y] + fe[6, :, y] + h[8, :, y]
    y = 0
    h[4, :, y] = -fe[2, :, y] + fe[4, :, y] + h[2, :, y]
    h[7, :, y] = -fe[5, :, y] + fe[7, :, y] + h[5, :, y]
    h[8, :, y] = -fe[6, :, y] + fe[8, :, y] + h[6, :, y]
    assert not np.isnan(h).any()
    carry = np.array([h[7, 0, 0], h[8, nx - 1, 0], h[6, 0, ny - 1], h[5, nx - 1, ny - 1]])
    return h, carry


def step_C_pull(ps: PullState, p: Params) -> None:
    """One launch of the fused kernel: gather + wall rule -> moments/overrides -> collide -> store f*."""
    h, carry = _gather_bc(ps, p)
    omega_nu = p.omega
    if p.turb == 1:
        Cs2 = 0.025
        product1 = 0.0
        for k in range(Q):
            product1 = (C[k, 0] * C[k, 1]) * h[k] + product1
        Qmf = product1 - ps.pi_eq
        tau0 = 1.0 / p.omega
        tau = 0.5 * (tau0 + np.sqrt((tau0 * tau0 + (18 * 1.4142 * Cs2 * np.abs(Qmf)) / ps.rho_prev)))
        omega_nu = 1.0 / tau
    rho_l, ux, uy = _moments_overrides(h, p)
    feq = _feq_kernel(rho_l, ux, uy)
    ps.g = _collide(h, rho_l, feq, p, omega_nu)
    ps.kind = "post"
    ps.rho_lid = rho_l[:, 0].copy()
    ps.carry = carry
    pi = 0.0
    for k in range(Q):
        pi = (C[k, 0] * C[k, 1]) * feq[k] + pi
    ps.pi_eq = pi
    ps.rho_prev = rho_l
    ps.rho = rho_l
    ps.u = np.stack([ux, uy])


def fin_from_pull(ps: PullState, p: Params) -> np.ndarray:
    """The reference's ``fin`` for the current step count (gather + wall rule, no collision)."""
    return _gather_bc(ps, p)[0]


# --------------------------------------------------------------------------------------------
# y-strip form of the fused pull pass (specification of the multi-GPU decomposition).
# --------------------------------------------------------------------------------------------
@dataclasses.dataclass
class StripState:
    """One y-strip [y0, y0+nyl) of a cavity in the fused-kernel representation.

    ``g`` is ``[9, nx, nyl+2]``: stored row r holds local row r-1, rows 0 and nyl+1 are ghost rows that the halo
    exchange fills with the neighbour strip's edge rows (only populations 4,7,8 of the row above and 2,5,6 of the
    row below are ever read).  Everything else as ``PullState``.
    """
    g: np.ndarray
    kind: str
    y0: int
    nyl: int
    rho_lid: np.ndarray
    carry: np.ndarray
    rho: Optional[np.ndarray] = None
    u: Optional[np.ndarray] = None

    @staticmethod
    def from_fin(fin0: np.ndarray, p: Params, y0: int, nyl: int) -> "StripState":
        nx, ny = p.nx, p.ny
        g = np.zeros((Q, nx, nyl + 2))
        g[:, :, 1:nyl + 1] = fin0[:, :, y0:y0 + nyl]
        carry = np.array([fin0[7, 0, 0], fin0[8, nx - 1, 0], fin0[6, 0, ny - 1], fin0[5, nx - 1, ny - 1]])
        return StripState(g=g, kind="pre", y0=y0, nyl=nyl, rho_lid=np.zeros(nx), carry=carry)


def _strip_gather_bc(ss: StripState, p: Params):
    nx, ny, y0, nyl = p.nx, p.ny, ss.y0, ss.nyl
    if ss.kind == "pre":
        return ss.g[:, :, 1:nyl + 1].copy(), ss.carry.copy()
    g = ss.g
    has_lid, has_bot = (y0 == 0), (y0 + nyl == ny)
    h = np.full((Q, nx, nyl), np.nan)
    for k in range(Q):
        cx, cy = int(C[k, 0]), int(C[k, 1])
        xd = slice(max(0, cx), nx - max(0, -cx))            # destination x with 0 <= x - cx < nx
        xs = slice(xd.start - cx, xd.stop - cx)
        j0, j1 = 0, nyl                                     # destination local rows with 0 <= y + cy < ny
        if cy < 0 and has_lid:
            j0 = 1
        if cy > 0 and has_bot:
            j1 = nyl - 1
        h[k, xd, j0:j1] = g[k, xs, j0 + cy + 1:j1 + cy + 1]
    if has_lid:
        h[7, 0, 0] = ss.carry[CARRY_TL]
        h[8, nx - 1, 0] = ss.carry[CARRY_TR]
    if has_bot:
        h[6, 0, nyl - 1] = ss.carry[CARRY_BL]
        h[5, nx - 1, nyl - 1] = ss.carry[CARRY_BR]
    rho_w = np.ones((nx, nyl)); ux = np.zeros((nx, nyl)); uy = np.zeros((nx, nyl))
    if has_lid:
        rho_w[:, 0] = ss.rho_lid
        ux[:, 0] = p.uLB
    fe = _feq_kernel(rho_w, ux, uy)
    x = 0
    h[1, x, :] = fe[1, x, :] - fe[3, x, :] + h[3, x, :]
    h[5, x, :] = fe[5, x, :] - fe[7, x, :] + h[7, x, :]
    h[8, x, :] = fe[8, x, :] - fe[6, x, :] + h[6, x, :]
    x = nx - 1
    h[3, x, :] = -fe[1, x, :] + fe[3, x, :] + h[1, x, :]
    h[6, x, :] = -fe[8, x, :] + fe[6, x, :] + h[8, x, :]
    h[7, x, :] = -fe[5, x, :] + fe[7, x, :] + h[5, x, :]
    if has_bot:
        y = nyl - 1
        h[2, :, y] = -fe[4, :, y] + fe[2, :, y] + h[4, :, y]
        h[5, :, y] = -fe[7, :, y] + fe[5, :, y] + h[7, :, y]
        h[6, :, y] = -fe[8, :, y] + fe[6, :, y] + h[8, :, y]
    if has_lid:
        y = 0
        h[4, :, y] = -fe[2, :, y] + fe[4, :, y] + h[2, :, y]
        h[7, :, y] = -fe[5, :, y] + fe[7, :, y] + h[5, :, y]
        h[8, :, y] = -fe[6, :, y] + fe[8, :, y] + h[6, :, y]
    assert not np.isnan(h).any()
    carry = ss.carry.copy()
    if has_lid:
        carry[CARRY_TL], carry[CARRY_TR] = h[7, 0, 0], h[8, nx - 1, 0]
    if has_bot:
        carry[CARRY_BL], carry[CARRY_BR] = h[6, 0, nyl - 1], h[5, nx - 1, nyl - 1]
    return h, carry


def step_C_pull_strip(ss: StripState, p: Params) -> None:
    """One launch on one strip; the caller must then exchange halo rows (see ``exchange_halo_local``)."""
    nx, ny, y0, nyl = p.nx, p.ny, ss.y0, ss.nyl
    h, carry = _strip_gather_bc(ss, p)
    # moments / overrides on the strip: embed the wall tests in global coordinates
    rho_l = _lsum([h[k] for k in range(Q)])
    ux = _lsum([C[k, 0] * h[k] for k in range(Q)]) / rho_l
    uy = _lsum([C[k, 1] * h[k] for k in range(Q)]) / rho_l
    ux[0, :] = 0; uy[0, :] = 0; ux[nx - 1, :] = 0; uy[nx - 1, :] = 0
    if y0 + nyl == ny:
        ux[:, nyl - 1] = 0; uy[:, nyl - 1] = 0
    if y0 == 0:
        rho_l = rho_l.copy()
        rho_l[:, 0] = h[0, :, 0] + h[1, :, 0] + h[3, :, 0] + 2 * (h[2, :, 0] + h[5, :, 0] + h[6, :, 0])
        ux[:, 0] = p.uLB; uy[:, 0] = 0
    feq = _feq_kernel(rho_l, ux, uy)
    ss.g[:, :, 1:nyl + 1] = _collide(h, rho_l, feq, p, p.omega)
    ss.kind = "post"
    if y0 == 0:
        ss.rho_lid = rho_l[:, 0].copy()
    ss.carry = carry
    ss.rho, ss.u = rho_l, np.stack([ux, uy])


def exchange_halo_local(strips) -> None:
    """In-process halo exchange between consecutive strips (what NCCL send/recv does between ranks)."""
    for a, b in zip(strips[:-1], strips[1:]):           # a above b
        for k in (4, 7, 8):                             # c_y = -1: cross towards larger y
            b.g[k, :, 0] = a.g[k, :, a.nyl]
        for k in (2, 5, 6):                             # c_y = +1: cross towards smaller y
            a.g[k, :, a.nyl + 1] = b.g[k, :, 1]


def strip_fin(ss: StripState, p: Params) -> np.ndarray:
    return _strip_gather_bc(ss, p)[0]


# --------------------------------------------------------------------------------------------
# Drivers
# --------------------------------------------------------------------------------------------
def run(p: Params, steps: int, semantics: str = "C", fin0: Optional[np.ndarray] = None, form: str = "push"):
    """Run ``steps`` iterations from the equilibrium start (or ``fin0``).

    Returns ``(rho, u, fin)`` exactly as the reference scripts leave them: ``fin`` after ``steps`` steps,
    ``rho``/``u`` = overridden moments of the state that *entered* the last step.
    """
    if fin0 is None:
        _, _, fin0 = init_fields(p.nx, p.ny, p.uLB)
    if semantics == "A":
        fin = np.array(fin0, dtype=np.float64)
        rho = np.ones((p.nx, p.ny)); u = np.zeros((2, p.nx, p.ny))
        for _ in range(steps):
            rho, u, _ = step_A(fin, p)
        return rho, u, fin
    if semantics != "C":
        raise ValueError(semantics)
    if form == "push":
        st = StateC.initial(p, fin0)
        for _ in range(steps):
            step_C(st, p)
        return st.rho, st.u, st.fin
    ps = PullState.from_fin(fin0, p)
    for _ in range(steps):
        step_C_pull(ps, p)
    rho = ps.rho if ps.rho is not None else np.ones((p.nx, p.ny))
    u = ps.u if ps.u is not None else np.zeros((2, p.nx, p.ny))
    return rho, u, fin_from_pull(ps, p)


def random_state(nx: int, ny: int, seed: int = 1234) -> np.ndarray:
    """Generic near-equilibrium state exercising every moment (SURVEY.md 8d 'value distributions')."""
    rng = np.random.default_rng(seed)
    rho = 1 + 0.01 * rng.uniform(-1, 1, (nx, ny))
    u = 0.05 * rng.uniform(-1, 1, (2, nx, ny))
    return equ(rho, u) * (1 + 1e-3 * rng.uniform(-1, 1, (Q, nx, ny)))


# --------------------------------------------------------------------------------------------
# Ghia et al. (1982) centre-line comparison -- fixture GhiaData.csv, Re = 100 columns only are clean
# --------------------------------------------------------------------------------------------
def load_ghia(path: Optional[str] = None):
    """Ghia Re = 100 centre-line stations: returns (Y, Ux(Y) at x = 0.5, X, Uy(X) at y = 0.5).

    ``tests/golden/ghia_re100.json`` is extracted from the reference fixture ``GhiaData.csv`` rows 7-23
    (the rows ``MRT.py:104`` slices) by ``tests/golden/make_golden.py``; only the Re = 100 columns are
    kept because the other columns of that file carry transcription errors.
    """
    import json
    with open(path or GHIA_JSON) as fh:
        d = json.load(fh)
    return (np.array(d["Y"]), np.array(d["Ux"]), np.array(d["X"]), np.array(d["Uy"]))


def ghia_errors(u: np.ndarray, uLB: float, ghia) -> Tuple[float, float]:
    """max |u_x(centre column) - Ghia| and max |u_y(centre row) - Ghia|, both / uLB.

    Own, correct station mapping (y_phys = 1 - j/(ny-1), x = i/(nx-1), linear interpolation); the
    reference's pairing (MRT.py:119-120, 559-561) is wrong and is deliberately not reused.
    """
    Y, Ux, X, Uy = ghia
    nx, ny = u.shape[1], u.shape[2]
    yphys = 1.0 - np.arange(ny) / (ny - 1.0)
    ux_col = 0.5 * (u[0, (nx - 1) // 2, :] + u[0, nx // 2, :]) / uLB
    ex = np.max(np.abs(np.interp(Y, yphys[::-1], ux_col[::-1]) - Ux))
    xphys = np.arange(nx) / (nx - 1.0)
    uy_row = 0.5 * (u[1, :, (ny - 1) // 2] + u[1, :, ny // 2]) / uLB
    ey = np.max(np.abs(np.interp(X, xphys, uy_row) - Uy))
    return float(ex), float(ey)


GHIA_JSON = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ghia_re100.json")


# --------------------------------------------------------------------------------------------
# Fast path: the same step_C compiled from oracle/lbm_oracle_c.c (bit-identical, multi-threaded)
# --------------------------------------------------------------------------------------------
def run_fast(p: Params, steps: int, fin0: Optional[np.ndarray] = None):
    """``run(p, steps, semantics='C', form='push')`` through the C restatement; falls back to NumPy if gcc is absent."""
    import ctypes as C_
    try:
        from . import build_oracle_c
        lib = C_.CDLL(build_oracle_c.build())
    except Exception:
        return run(p, steps, semantics="C", fin0=fin0, form="push")

    class _P(C_.Structure):
        _fields_ = [("nx", C_.c_int), ("ny", C_.c_int), ("collision", C_.c_int), ("turb", C_.c_int),
                    ("uLB", C_.c_double), ("omega", C_.c_double), ("omega_e", C_.c_double),
                    ("omega_eps", C_.c_double), ("omega_q", C_.c_double), ("omegam", C_.c_double)]

    st = StateC.initial(p, fin0)
    cp = _P(p.nx, p.ny, {"SRT": 0, "TRT": 1, "MRT": 2}[p.collision], int(p.turb), p.uLB, p.omega, p.omega_e,
            p.omega_eps, p.omega_q, p.omegam)
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (st.fin, st.ftemp, st.feq, st.rho, st.u)]
    dp = C_.POINTER(C_.c_double)
    lib.oracle_step_C.argtypes = [C_.POINTER(_P)] + [dp] * 5 + [C_.c_int]
    lib.oracle_step_C.restype = None
    lib.oracle_step_C(C_.byref(cp), *[a.ctypes.data_as(dp) for a in arrs], int(steps))
    return arrs[3], arrs[4], arrs[0]
