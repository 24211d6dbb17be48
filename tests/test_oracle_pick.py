"""Picking ray (World::RaycastDetect, Core/World.cpp:496-546): oracle self-checks, oracle == the reference's own C++
(lifted in place into oracle/_ref), golden fixture.  CPU only."""
from pathlib import Path

import numpy as np
import pytest

import pick_util as pu
from oracle import ref_binding as rb

GOLD = Path(__file__).resolve().parent / "golden" / "pick_ref.npz"


def test_pick_hits_are_solid_inside_reach_and_face_the_ray(plains0, plains0_oracle):
    o, d = pu.pick_rays(50_000, 1)
    r = plains0_oracle.raycast_detect(o, d)
    found = r[:, 7] == 1
    assert 0.1 < found.mean() < 0.9
    h = r[found]
    assert (plains0[h[:, 2], h[:, 1], h[:, 0]] == h[:, 3]).all() and (h[:, 3] > 0).all()
    assert (h[:, :3] > 0).all()                                   # index 0 is outside for the picking ray
    centre = h[:, :3] + 0.5
    assert (np.linalg.norm(centre - o[found], axis=1) < 48 * 1.8 + 2).all()
    n = h[:, 4:7].astype(np.float32)
    assert ((n * d[found]).sum(1) <= 0).all()                     # the face normal opposes the ray
    assert (r[~found, :4] == -1).all() and (r[~found, 4:7] == 0).all()


def test_oracle_pick_matches_reference_golden(plains0_oracle):
    g = np.load(GOLD)
    r = plains0_oracle.raycast_detect(g["positions"], g["directions"])
    found = r[:, 7] == 1
    assert np.array_equal(r[found, :4], g["hits"][found]) and (g["hits"][~found] == -2).all()
    assert found.sum() > 1000


@pytest.mark.skipif(not rb.available("raycast"), reason="oracle/_ref not built on this box")
def test_oracle_pick_equals_the_reference_function(plains0, plains0_oracle):
    o, d = pu.pick_rays(200_000, 3)
    a, b = plains0_oracle.raycast_detect(o, d), rb.raycast_detect(plains0, o, d)
    found = a[:, 7] == 1
    assert np.array_equal(a[found, :4], b[found])
    assert (b[~found] == -2).all()          # the return the generated copy adds where the reference falls off the end
