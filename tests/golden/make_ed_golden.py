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


if __name__ == "__main__":
    out = [ed_chain(8, 1.0, 1.0, 0.2), ed_chain(8, 1.0, 0.5, 0.25), ed_chain(8, 1.0, 2.0, 0.5),
           ed_chain(8, 1.0, 0.0, 0.2), ed_chain(10, 1.0, 0.5, 0.2)]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ed_chain.json")
    json.dump(out, open(path, "w"), indent=1)
    for o in out:
        print(o)
