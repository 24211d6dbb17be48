"""Pin of the oracle for SURVEY 8 rows a8 - a11 (diffuse GI, reflections, hit-material fetch, Cook-Torrance direct term):

(1) against tests/golden/shade_ref.npz — outputs of the reference's own GenerateGBuffer.glsl:351-578, ColorPassFrag.glsl:394-451,
    DiffuseRayTraceFrag.glsl:535-664,910-1021 and ReflectionTraceFrag.glsl:717-1038 compiled through oracle/build_ref.py
    (tests/golden/make_golden_shade.py) — byte for byte, on every box;
(2) live against oracle/_ref where that build is present, on more frames (other poses, frames, spp 1-4, checkerboard,
    reprojection + temporal jitter, both worlds).
CPU only."""
import numpy as np
import pytest

import scene_util as su
import shade_golden_util as sg
from oracle import binding as ob
from oracle import ref_binding as rb
from voxeltracing_b200 import host_api


class OracleBackend:
    def __init__(self, blocks, inputs, df=None):
        self.blocks = blocks
        self.ow = ob.OracleWorld(blocks, df)
        self.sc = ob.OracleScene(self.ow)
        inputs.apply_to_oracle(self.sc)

    def initial_trace(self, p):
        return self.ow.initial_trace(p)

    def shadow_trace(self, p, g_t, g_n):
        return self.ow.shadow_trace(p, g_t, g_n, None)

    def generate_gbuffer(self, *a):
        return self.sc.generate_gbuffer(*a)

    def shade_direct(self, *a):
        return self.sc.shade_direct(*a)

    def diffuse_trace(self, *a):
        return self.sc.diffuse_trace(*a)

    def reflection_trace(self, *a):
        return self.sc.reflection_trace(*a)

    def set_lpv(self, level, btype, avg):
        self.sc.set_lpv(level, btype, avg)


@pytest.fixture(scope="module")
def inputs():
    return sg.inputs()


@pytest.mark.parametrize("wname", sg.WORLDS)
def test_oracle_shade_passes_match_reference_golden(wname, inputs):
    g = sg.golden()
    be = OracleBackend(sg.world(wname), inputs)
    n = 0
    for case in sg.CASES:
        if case["world"] != wname:
            continue
        res = be_res = sg.run_case(be, case, inputs)
        for k, v in be_res.items():
            if not isinstance(v, np.ndarray):
                continue
            assert sg.same_bits(v, g[f"{case['name']}_{k}"]), (case["name"], k)
            n += 1
        # the fixture is not degenerate: lit albedo, some GI radiance, some reflection hits
        assert res["gb_albedo"].astype(np.float32).std() > 0.01
        assert res["gi0_sh"].astype(np.float32).std() > 1e-3
        assert (res["refl0_hitdist"].astype(np.float32) > 0).mean() > 0.02
    assert n >= 15


def test_golden_covers_spp4_and_checkerboard():
    """Blue-noise quirk of spp = 4 (sample index past the ranking tile, DiffuseRayTraceFrag.glsl:139-148) and the checkerboard
    path are in the fixture and differ from the lower-spp frames."""
    g = sg.golden()
    assert not sg.same_bits(g["rooms_c_gi0_sh"], g["rooms_c_gi1_sh"])
    assert g["rooms_c_gi1_sh"].astype(np.float32).std() > 1e-3
    assert [c for c in sg.CASES if any(k["spp"] == 4 for k in c["gi"])]


needs_ref = pytest.mark.skipif(not all(rb.available(w) for w in ("df", "initial", "shadow", "gbuffer", "diffuse", "reflection", "color")),
                               reason="oracle/_ref not built on this box")


class RefBackend:
    def __init__(self, blocks, df, inputs):
        self.blocks, self.df = blocks, df
        rb.set_scene(blocks, df, inputs.table, inputs.blue, inputs.textures, inputs.sky)

    def initial_trace(self, p):
        return rb.initial_trace(self.blocks, self.df, p)

    def shadow_trace(self, p, g_t, g_n):
        return rb.shadow_trace(self.blocks, self.df, p, g_t, g_n, None)

    generate_gbuffer = staticmethod(rb.generate_gbuffer)
    shade_direct = staticmethod(rb.shade_direct)
    diffuse_trace = staticmethod(rb.diffuse_trace)
    reflection_trace = staticmethod(rb.reflection_trace)
    set_lpv = staticmethod(rb.set_lpv)


LIVE_CASES = [
    dict(name="live_rooms", world="rooms2", pos=[188.3, 61.0, 172.9], yaw=115.0, pitch=-8.0, sun_ticks=(50.0, 130.0),
         gi=[dict(frame=17, spp=1, checkerboard=True), dict(frame=130, spp=2, checkerboard=False), dict(frame=3, spp=4, checkerboard=True)],
         refl=[dict(frame=4, spp=1, reproject=True), dict(frame=200, spp=3, temporal=True), dict(frame=1, spp=4)]),
    dict(name="live_plains", world="plains1", pos=[120.7, 95.0, 260.1], yaw=290.0, pitch=-35.0, sun_ticks=(20.0,),
         gi=[dict(frame=64, spp=3, checkerboard=False), dict(frame=7, spp=4, checkerboard=False)],
         refl=[dict(frame=11, spp=2, reproject=True, temporal=True)]),
]


def test_oracle_lpv_gi_matches_reference_golden(inputs):
    """ApproximateGILPV inside the reflection pass: the oracle against the compiled shader's outputs with u_LPVGI on."""
    g = sg.golden_lpv()
    case = sg.LPV_CASE
    res = sg.run_case(OracleBackend(sg.world(case["world"]), inputs), case, inputs)
    n = 0
    for k, v in res.items():
        if k.startswith("refl"):
            assert sg.same_bits(v, g[f"{case['name']}_{k}"]), k
            n += 1
    assert n == 3 * len(case["refl"])
    # the volume changes the picture: same frame / spp with and without the LPV term differ
    plain = sg.golden()
    assert not sg.same_bits(g["rooms_lpv_refl0_color"], plain["rooms_a_refl0_color"])


@needs_ref
def test_oracle_lpv_gi_equals_compiled_reference_shader_live(inputs):
    case = dict(sg.LPV_CASE, name="live_lpv", pos=[188.3, 61.0, 172.9], yaw=115.0, pitch=-8.0, lpv_limit=4,
                refl=[dict(frame=5, spp=2, lpv_gi=True, reproject=True), dict(frame=1, spp=1, lpv_gi=True, decoupled=True, ss_sky_valid=True, temporal=True)])
    blocks = sg.world(case["world"])
    df = ob.distance_field(blocks)
    a = sg.run_case(OracleBackend(blocks, inputs, df), case, inputs)
    b = sg.run_case(RefBackend(blocks, df, inputs), case, inputs)
    for k in a:
        assert sg.same_bits(a[k], b[k]), k


@needs_ref
@pytest.mark.parametrize("case", LIVE_CASES, ids=lambda c: c["name"])
def test_oracle_equals_compiled_reference_shaders_live(case, inputs):
    blocks = sg.world(case["world"])
    df = ob.distance_field(blocks)
    assert np.array_equal(rb.distance_field(blocks), df)
    a = sg.run_case(OracleBackend(blocks, inputs, df), case, inputs)
    b = sg.run_case(RefBackend(blocks, df, inputs), case, inputs)
    assert a.keys() == b.keys()
    for k in a:
        assert sg.same_bits(a[k], b[k]), (case["name"], k)
