"""Host logic of the engine that runs without a GPU: the spatial tiling behind lq_create
(include/lq.h lq_tiling_info; tiles of neighbouring sites, bonds owned by the tile of their source
site, halo = foreign bonds touching a site of an owned bond, stencils shared per tile shape)."""
import numpy as np
import pytest

import looper_lattices as ll


def _lq():
    import looper_b200 as lq
    return lq


def test_square_lattice_tiles_share_a_handful_of_stencils():
    lq = _lq()
    t = lq.tiling_info(ll.hypercubic_lattice((64, 64)), 256)
    # 4 x 4 tiles of 16 x 16 sites; the stencil depends on the order in which the neighbouring bonds
    # are numbered, which differs at the periodic seam: interior / edge / corner = 9 shapes, for any size
    assert t["num_tiles"] == 16 and t["num_classes"] == 9
    assert lq.tiling_info(ll.hypercubic_lattice((128, 128)), 256)["num_classes"] == 9
    assert t["max_sites"] == 256 and t["max_bonds"] == 512 and t["max_degree"] == 4
    assert t["owned_bonds"] == 2 * 64 * 64                         # every bond owned exactly once
    assert t["max_ksites"] == 256 + 16 + 16                        # own sites + far ends of the right / top bonds
    assert t["max_walk_halo"] == 32                                # bonds entering the left column and the bottom row
    # + per far-end column / row: its 16 outward bonds and the 17 bonds along it (neighbours share them)
    assert t["max_halo_buckets"] == 32 + 2 * (16 + 17)
    assert t["halo_buckets"] == 16 * 98


def test_site_pseudo_bonds_join_the_tiling():
    lq = _lq()
    a = lq.tiling_info(ll.hypercubic_lattice((64, 64)), 256)
    b = lq.tiling_info(ll.hypercubic_lattice((64, 64)), 256, with_sites=True)
    assert b["owned_bonds"] == a["owned_bonds"] + 64 * 64          # one one-ended pseudo-bond per site
    assert b["max_bonds"] == a["max_bonds"] + 256 and b["max_degree"] == a["max_degree"] + 1
    assert b["max_halo_buckets"] == a["max_halo_buckets"] + 32    # the pseudo-bonds of the 32 far-end sites
    assert b["max_walk_halo"] == a["max_walk_halo"]                # pseudo-bonds of own sites are owned, not halo
    assert b["num_tiles"] == a["num_tiles"] and b["num_classes"] == a["num_classes"]


@pytest.mark.parametrize("dims,tile,tiles", [((100,), 16, 7), ((8, 8, 8), 64, 8), ((6, 10), 16, 6), ((16,), 64, 1)])
def test_ragged_and_small_lattices(dims, tile, tiles):
    lq = _lq()
    lat = ll.chain_lattice(dims[0]) if len(dims) == 1 else ll.hypercubic_lattice(dims)
    t = lq.tiling_info(lat, tile)
    assert t["num_tiles"] == tiles
    assert t["owned_bonds"] == len(lat["src"])
    assert t["max_sites"] <= tile and t["max_bonds"] <= t["max_sites"] * len(dims)
    assert t["max_degree"] == 2 * len(dims)
    assert t["max_walk_halo"] <= t["max_halo_buckets"] <= 1024
    if tiles == 1:
        assert t["halo_buckets"] == 0 and t["max_ksites"] == t["max_sites"]


def test_bad_lattices_are_rejected():
    lq = _lq()
    lat = ll.chain_lattice(8)
    bad = dict(lat)
    bad["dst"] = lat["dst"].copy()
    bad["dst"][3] = 99
    with pytest.raises(lq.LqError):
        lq.tiling_info(bad, 4)
    loop = dict(lat)
    loop["dst"] = lat["dst"].copy()
    loop["dst"][2] = lat["src"][2]
    with pytest.raises(lq.LqError):
        lq.tiling_info(loop, 4)
