import sys
import numpy as np
sys.path.insert(0, "tests")
import scene_util as su
from voxeltracing_b200 import engine, host_api
c = engine.Context(0)
blocks = host_api.gen_world("rooms", 2)
rng = np.random.default_rng(4)
nz, ny, nx = blocks.shape
blocks[rng.integers(1, nz, 2000), rng.integers(1, ny, 2000), rng.integers(1, nx, 2000)] = 12
t = np.full((6, 128), -1, dtype=np.int32); t[3, 12] = 0; t[0, :40] = np.arange(40)
c.set_block_data(t); c.upload_world(blocks)
su.SceneInputs(512).apply_to_context(c)
c.set_block_data(t)
c.lpv_average_colors()
c.lpv_repropagate(None, 8)
level = c.lpv_download()[0]
lit = np.argwhere(level > 0)
pts = (lit[rng.integers(0, len(lit), 1 << 21)][:, ::-1] + rng.random((1 << 21, 3)) * 1.5 - 0.25).astype(np.float32)
for _ in range(2):
    c.lpv_sample(pts, (0.001, 0.003, 0.001))
c.close()
