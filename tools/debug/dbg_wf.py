import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import scene_util as su
from voxeltracing_b200 import abi, engine, host_api
W,H=320,180
inp = su.SceneInputs(128)
blocks = host_api.gen_world("rooms", 2)
c = engine.Context(0); c.upload_world(blocks); c.generate_distance_field(); inp.apply_to_context(c)
cam = host_api.camera([200,58,200], 75.0, -12.0, W/H)
c.initial_trace(cam, W, H)
atts = (abi.ATT_GI_SH, abi.ATT_GI_COCG, abi.ATT_GI_UTILITY, abi.ATT_GI_AOSKY)
for order in ("gi-first", "refl-first"):
    if order == "refl-first":
        c.shadow_trace(cam, W, H, host_api.sun_direction(50.0)[2], soft=False)
        c.generate_gbuffer(su.gbuffer_params(cam, W, H, inp))
        c.diffuse_trace(su.gi_params(cam, W, H, frame=1, spp=1))
        c.reflection_trace(su.reflection_params(cam, W, H, inputs=inp, frame=1, spp=2))
    for spp, chk in ((1, False), (3, True), (3, False)):
        ip = su.gi_params(cam, W, H, frame=4, spp=spp, checkerboard=chk)
        res = {}
        for mode in (0, 1, 1):
            c.set_option("wavefront", mode)
            c.diffuse_trace(ip)
            res.setdefault(mode, []).append([c.read_attachment(a).copy() for a in atts])
        c.set_option("wavefront", 1)
        d01 = [int((a.view(np.uint8) != b.view(np.uint8)).sum()) for a, b in zip(res[0][0], res[1][0])]
        d11 = [int((a.view(np.uint8) != b.view(np.uint8)).sum()) for a, b in zip(res[1][0], res[1][1])]
        print(order, spp, chk, "mega-vs-wf", d01, "wf-vs-wf", d11)
