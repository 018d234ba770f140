#!/usr/bin/env python
"""Exact diagonalisation of small S=1/2 XXZ chains (numpy restatement of the thermal averages of
diag.C:376-468: energy, uniform susceptibility, staggered magnetisation^2 and staggered
susceptibility with the kernel of diag.C:272-277).  Writes tests/golden/ed_chain.json.
H = sum_b [ Jz Sz Sz + Jxy/2 (S+S- + S-S+) ], periodic chain of L sites."""
import json
import os

import numpy as np


def ed_chain(L, jxy, jz, T):
    dim = 1 << L
    H = np.zeros((dim, dim))
    sz = lambda s, i: 0.5 - ((s >> i) & 1)
    for s in range(dim):
        for i in range(L):
            j = (i + 1) % L
            H[s, s] += jz * sz(s, i) * sz(s, j)
            if ((s >> i) & 1) != ((s >> j) & 1):
                s2 = s ^ (1 << i) ^ (1 << j)
                H[s2, s] += 0.5 * jxy
    E, V = np.linalg.eigh(H)
    beta = 1.0 / T
    w = np.exp(-beta * (E - E.min()))
    Z = w.sum()
    ene = (w * E).sum() / Z
    states = np.arange(dim)
    mu = sum(0.5 - ((states >> i) & 1) for i in range(L))
    ms = sum((1 - 2 * (i % 2)) * (0.5 - ((states >> i) & 1)) for i in range(L))
    P = V ** 2                                   # |<s|n>|^2
    umag2 = (w * (P * (mu ** 2)[:, None]).sum(0)).sum() / Z
    smag2 = (w * (P * (ms ** 2)[:, None]).sum(0)).sum() / Z
    # staggered susceptibility: sum_nm |<n|Ms|m>|^2 (w_m - w_n)/(E_n - E_m), beta*w_n on the diagonal
    M = V.T @ (ms[:, None] * V)
    dE = E[:, None] - E[None, :]
    wn, wm = w[:, None], w[None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        K = np.where(np.abs(dE) > 1e-12, (wm - wn) / dE, beta * wn)
    ssus = (M ** 2 * K).sum() / Z
    return dict(L=L, jxy=jxy, jz=jz, T=T, energy_density=ene / L, usus_density=beta * umag2 / L,
                smag2=smag2, ssus_density=ssus / L)


def ed_bonds(n, bonds, gauge, jxy, jz, gamma, T):
    """Same thermal averages as ed_tfi for an arbitrary bond list (n <= 10 sites): used for the 4 x 2
    ladder, the smallest two-dimensional case of the tests."""
    dim = 1 << n
    H = np.zeros((dim, dim))
    SX = np.zeros((dim, dim))
    sz = lambda s, i: 0.5 - ((s >> i) & 1)
    for s in range(dim):
        for (i, j) in bonds:
            H[s, s] += jz * sz(s, i) * sz(s, j)
            if ((s >> i) & 1) != ((s >> j) & 1):
                H[s ^ (1 << i) ^ (1 << j), s] += 0.5 * jxy
        for i in range(n):
            SX[s ^ (1 << i), s] += 0.5
    H -= gamma * SX
    E, V = np.linalg.eigh(H)
    beta = 1.0 / T
    w = np.exp(-beta * (E - E.min()))
    Z = w.sum()
    states = np.arange(dim)
    mu = sum(0.5 - ((states >> i) & 1) for i in range(n))
    ms = sum(gauge[i] * (0.5 - ((states >> i) & 1)) for i in range(n))
    P = V ** 2
    dE = E[:, None] - E[None, :]
    wn, wm = w[:, None], w[None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        K = np.where(np.abs(dE) > 1e-12, (wm - wn) / dE, beta * wn)

    def kubo(diag):
        M = V.T @ (diag[:, None] * V)
        return (M ** 2 * K).sum() / Z

    return dict(n=n, bonds=[list(b) for b in bonds], gauge=list(gauge), jxy=jxy, jz=jz, gamma=gamma, T=T,
                energy_density=(w * E).sum() / Z / n,
                transmag_density=(w * np.einsum("sn,st,tn->n", V, SX, V)).sum() / Z / n,
                umag2=(w * (P * (mu ** 2)[:, None]).sum(0)).sum() / Z,
                smag2=(w * (P * (ms ** 2)[:, None]).sum(0)).sum() / Z,
                usus_density=kubo(mu) / n, ssus_density=kubo(ms) / n)


def ladder_4x2():
    """hypercubic_lattice((4, 2)) of tests/looper_lattices.py: x-bonds on both legs, one rung per x."""
    n = 8
    bonds = [(s, (s % 4 + 1) % 4 + 4 * (s // 4)) for s in range(8)] + [(x, x + 4) for x in range(4)]
    gauge = [1.0 if ((s % 4) + (s // 4)) % 2 == 0 else -1.0 for s in range(8)]
    return n, bonds, gauge


def ed_tfi(L, jxy, jz, gamma, T):
    """Transverse field: H = sum_b [Jz Sz Sz + Jxy/2 (S+S- + S-S+)] - Gamma sum_i Sx_i
    (site term of weight_impl.h:62-88: offdiagonal element Hx/2).  Thermal averages of the energy,
    of sum_i Sx_i (transmag.h:106-109), of the equal-time Mz^2 / staggered Mz^2 and the Kubo
    susceptibilities of Mz and staggered Mz (susceptibility.h:199-254)."""
    dim = 1 << L
    H = np.zeros((dim, dim))
    SX = np.zeros((dim, dim))
    sz = lambda s, i: 0.5 - ((s >> i) & 1)
    for s in range(dim):
        for i in range(L):
            j = (i + 1) % L
            if L > 2 or i == 0:
                H[s, s] += jz * sz(s, i) * sz(s, j)
                if ((s >> i) & 1) != ((s >> j) & 1):
                    H[s ^ (1 << i) ^ (1 << j), s] += 0.5 * jxy
            SX[s ^ (1 << i), s] += 0.5
    H -= gamma * SX
    E, V = np.linalg.eigh(H)
    beta = 1.0 / T
    w = np.exp(-beta * (E - E.min()))
    Z = w.sum()
    states = np.arange(dim)
    mu = sum(0.5 - ((states >> i) & 1) for i in range(L))
    ms = sum((1 - 2 * (i % 2)) * (0.5 - ((states >> i) & 1)) for i in range(L))
    P = V ** 2
    dE = E[:, None] - E[None, :]
    wn, wm = w[:, None], w[None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        K = np.where(np.abs(dE) > 1e-12, (wm - wn) / dE, beta * wn)

    def kubo(diag):
        M = V.T @ (diag[:, None] * V)
        return (M ** 2 * K).sum() / Z

    return dict(L=L, jxy=jxy, jz=jz, gamma=gamma, T=T,
                energy_density=(w * E).sum() / Z / L,
                transmag_density=(w * np.einsum("sn,st,tn->n", V, SX, V)).sum() / Z / L,
                umag2=(w * (P * (mu ** 2)[:, None]).sum(0)).sum() / Z,
                smag2=(w * (P * (ms ** 2)[:, None]).sum(0)).sum() / Z,
                usus_density=kubo(mu) / L, ssus_density=kubo(ms) / L)


def ed_stiffness(n, bonds, rvec, dim, jxy, jz, T):
    """Spin stiffness as the reference reports it (looper/stiffness.h:126-129,160-167: <W^2> / (beta dim),
    W = sum over the off-diagonal operators of +-bond_vector_relative, an integer per direction).
    A twist phi turns the hopping term of bond b into Jxy/2 (e^{i phi r_b} S+_src S-_dst + h.c.), so that
    Z(phi) = sum_W Z_W e^{i phi W} and <W^2> = -d^2 ln Z / d phi^2 = beta F''(0):
      F''(0) = <d^2H/dphi^2> - sum_nm |<n|dH/dphi|m>|^2 K_nm / Z   (second-order perturbation theory,
    K_nm = (w_m - w_n)/(E_n - E_m), beta w_n on degenerate pairs).  Returned per direction and summed;
    `fd` is the same number from central differences of ln Z(phi) with the complex Hamiltonian."""
    dimH = 1 << n
    sz = lambda s, i: 0.5 - ((s >> i) & 1)

    def ham(phi, k):
        H = np.zeros((dimH, dimH), dtype=complex)
        for s in range(dimH):
            for b, (i, j) in enumerate(bonds):
                H[s, s] += jz * sz(s, i) * sz(s, j)
                bi, bj = (s >> i) & 1, (s >> j) & 1
                if bi != bj:   # bit clear = up: S+_i S-_j moves an up spin from j to i
                    ph = np.exp(1j * phi * rvec[b][k] * (1 if bi == 1 else -1))
                    H[s ^ (1 << i) ^ (1 << j), s] += 0.5 * jxy * ph
        return H

    beta = 1.0 / T
    H0 = ham(0.0, 0).real
    E, V = np.linalg.eigh(H0)
    w = np.exp(-beta * (E - E.min()))
    Z = w.sum()
    dE = E[:, None] - E[None, :]
    wn, wm = w[:, None], w[None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        K = np.where(np.abs(dE) > 1e-10, (wm - wn) / dE, beta * wn)
    w2, fd = 0.0, 0.0
    for k in range(dim):
        J = np.zeros((dimH, dimH), dtype=complex)
        T2 = np.zeros((dimH, dimH))
        for s in range(dimH):
            for b, (i, j) in enumerate(bonds):
                bi, bj = (s >> i) & 1, (s >> j) & 1
                if bi != bj:
                    sg = 1 if bi == 1 else -1
                    s2 = s ^ (1 << i) ^ (1 << j)
                    J[s2, s] += 0.5 * jxy * 1j * rvec[b][k] * sg
                    T2[s2, s] -= 0.5 * jxy * rvec[b][k] ** 2
        M = V.T @ J @ V
        t2 = (w * np.einsum("sn,st,tn->n", V, T2, V)).sum() / Z
        f2 = t2 - ((np.abs(M) ** 2) * K).sum() / Z
        w2 += beta * f2
        h = 1e-3
        lz = []
        for phi in (-h, 0.0, h):
            Ep = np.linalg.eigvalsh(ham(phi, k))
            lz.append(np.log(np.exp(-beta * (Ep - E.min())).sum()))
        fd += -(lz[0] - 2 * lz[1] + lz[2]) / h ** 2
    return dict(n=n, bonds=[list(b) for b in bonds], rvec=[list(r) for r in rvec], dim=dim, jxy=jxy, jz=jz, T=T,
                w2=w2, w2_fd=fd, stiffness=w2 / (beta * dim))


if __name__ == "__main__":
    chain = lambda L: ([(i, (i + 1) % L) for i in range(L)], [[1.0 / L, 0, 0]] * L)
    n, lbonds, _ = ladder_4x2()
    lvec = [[0.25, 0, 0]] * 8 + [[0, 0.5, 0]] * 4
    st = [ed_stiffness(8, *chain(8), 1, 1.0, 1.0, 0.25), ed_stiffness(8, *chain(8), 1, 1.0, 0.5, 0.5),
          ed_stiffness(6, *chain(6), 1, -1.0, -1.0, 0.5), ed_stiffness(n, lbonds, lvec, 2, 1.0, 1.0, 0.4)]
    json.dump(st, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ed_stiffness.json"), "w"), indent=1)
    for o in st:
        print({k: v for k, v in o.items() if k not in ("bonds", "rvec")})
    tfi = [ed_tfi(8, 0.0, 1.0, 0.7, 0.5), ed_tfi(8, 0.0, -1.0, 0.5, 0.4), ed_tfi(8, -1.0, 0.5, 0.6, 0.4),
           ed_tfi(6, -1.0, -1.0, 1.0, 0.25)]   # Jxy <= 0 with a field: no sign problem
    json.dump(tfi, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ed_tfi.json"), "w"), indent=1)
    n, bonds, gauge = ladder_4x2()
    lad = [ed_bonds(n, bonds, gauge, 1.0, 1.0, 0.0, 0.3), ed_bonds(n, bonds, gauge, -1.0, 0.5, 0.6, 0.4),
           ed_bonds(n, bonds, gauge, 0.0, 1.0, 0.8, 0.5)]
    json.dump(lad, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ed_ladder.json"), "w"), indent=1)
    for o in lad:
        print({k: v for k, v in o.items() if k not in ("bonds", "gauge")})
    for o in tfi:
        print(o)
    out = [ed_chain(8, 1.0, 1.0, 0.2), ed_chain(8, 1.0, 0.5, 0.25), ed_chain(8, 1.0, 2.0, 0.5),
           ed_chain(8, 1.0, 0.0, 0.2), ed_chain(10, 1.0, 0.5, 0.2)]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ed_chain.json")
    json.dump(out, open(path, "w"), indent=1)
    for o in out:
        print(o)
