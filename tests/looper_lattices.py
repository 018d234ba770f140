"""Lattice helpers for CPU-only tests (no liblq.so needed): same generators as the package."""
import numpy as np


def chain_lattice(L):
    src = np.arange(L, dtype=np.int32)
    dst = ((src + 1) % L).astype(np.int32)
    gauge = np.where(np.arange(L) % 2 == 0, 1.0, -1.0)
    vec = np.zeros((L, 3))
    vec[:, 0] = 1.0 / L     # ALPS bond_vector_relative: bond vector over the lattice extent
    return dict(num_sites=L, src=src, dst=dst, gauge=gauge, dims=(L, 0, 0), bond_vectors=vec, vector_dim=1)


def hypercubic_lattice(dims):
    dims = [int(d) for d in dims if int(d) > 1]
    n = int(np.prod(dims))
    idx = np.arange(n)
    coords, rem = [], idx.copy()
    for d in dims:
        coords.append(rem % d)
        rem = rem // d
    src, dst, vecs, stride = [], [], [], 1
    for k, d in enumerate(dims):
        nxt = idx + stride * (((coords[k] + 1) % d) - coords[k])
        keep = (coords[k] == 0) if d == 2 else np.ones(n, dtype=bool)
        src.append(idx[keep]); dst.append(nxt[keep])
        v = np.zeros((int(keep.sum()), 3))
        if k < 3:
            v[:, k] = 1.0 / d
        vecs.append(v)
        stride *= d
    parity = sum(coords)
    bip = all(d % 2 == 0 for d in dims)
    gauge = np.where(parity % 2 == 0, 1.0, -1.0) if bip else np.zeros(n)
    return dict(num_sites=n, src=np.concatenate(src).astype(np.int32),
                dst=np.concatenate(dst).astype(np.int32), gauge=gauge,
                dims=tuple(dims + [0] * (3 - len(dims))),
                bond_vectors=np.concatenate(vecs), vector_dim=min(len(dims), 3))
