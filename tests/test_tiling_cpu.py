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


# ---- spatial cut (lq_options.cut = LQ_CUT_SPACE): the host-side plan, lq_space_plan_info -------------
@pytest.mark.parametrize("P", [2, 4])
def test_space_plan_of_the_headline_lattice_shape(P):
    """square 64 x 64, tiles of 16 x 16 sites (the shape of the 1024-wide benchmark lattice scaled down):
    every rank owns a strip of tile rows, mirrors the tile row above (walked: it holds the far-end sites)
    and the tile row below, and lists per cut 32 + 16 boundary bonds per tile and 16 boundary sites."""
    lq = _lq()
    lat = ll.hypercubic_lattice((64, 64))
    owned = 0
    for r in range(P):
        p = lq.space_plan_info(lat, P, r, tile_sites=256)
        owned += p["owned_sites"]
        assert p["owned_tiles"] == 16 // P and p["owned_sites"] == 64 * 64 // P
        # above: one row of 4 tiles, walked; below: one row of 4 tiles, not walked.  With two ranks of two tile rows
        # each both neighbours are the same rank, with P = 4 a rank owns ONE tile row and the rows differ
        assert p["walked_ghost_tiles"] == 4 and p["walked_sites"] == p["owned_sites"] + 4 * 256
        assert p["ghost_tiles"] == (8 if 16 // P >= 1 else 4)
        assert p["local_sites"] == p["owned_sites"] + p["ghost_tiles"] * 256
        assert p["neighbours"] == (1 if P == 2 else 2)
        assert p["segments"] == 2 * p["neighbours"] if P > 2 else p["segments"] == 2
        # as OWNER: towards the rank below, the 64 horizontal + 64 vertical bonds of the first site row and its 64
        # sites (they are the far-end sites of the rank below); towards the rank above, the 64 vertical bonds that
        # leave the last site row.  As USER: the mirror image.
        assert p["owner_bonds"] == 192 and p["owner_sites"] == 64
        assert p["user_bonds"] == 192 and p["user_sites"] == 64
    assert owned == 64 * 64


@pytest.mark.parametrize("name,dims,ts,P", [("chain", (16,), 2, 4), ("rect", (12, 8), 8, 3), ("cubic", (4, 4, 4), 8, 2),
                                            ("square_sites", (8, 8), 4, 4), ("square_half_rows", (64, 64), 256, 8)])
def test_space_plan_sides_agree(name, dims, ts, P):
    """the user side of rank r for owner q lists exactly what q's owner side lists for r, in the same order
    (the merge identifies entry i of one with entry i of the other); with site pseudo-bonds too."""
    lq = _lq()
    lat = ll.chain_lattice(dims[0]) if len(dims) == 1 else ll.hypercubic_lattice(dims)
    ws = name == "square_sites"
    owned = 0
    for q in range(P):
        owned += lq.space_plan_info(lat, P, q, tile_sites=ts, with_sites=ws)["owned_sites"]
        for r in range(P):
            if q == r:
                continue
            a = lq.space_plan_info(lat, P, q, tile_sites=ts, with_sites=ws, peer=r)
            b = lq.space_plan_info(lat, P, r, tile_sites=ts, with_sites=ws, peer=q)
            assert a["checksum_owner"] == b["checksum_user"]
            assert a["checksum_user"] == b["checksum_owner"]
    assert owned == lat["num_sites"]


def test_space_plan_rejects_more_ranks_than_tiles():
    lq = _lq()
    with pytest.raises(lq.LqError):
        lq.space_plan_info(ll.chain_lattice(8), 4, 0, tile_sites=4)     # two tiles, four ranks
