"""Ray batches for the picking-ray tests (World::RaycastDetect)."""
import numpy as np


def pick_rays(n, seed, dims=(384, 128, 384)):
    """Random rays from inside / just outside the world, plus axis-aligned directions (inf / NaN in tvec) and origins on
    the voxel lattice (t == 0 ties on several axes)."""
    rng = np.random.default_rng(seed)
    nx, ny, nz = dims
    o = np.stack([rng.uniform(-5, nx + 5, n), rng.uniform(-5, ny + 5, n), rng.uniform(-5, nz + 5, n)], 1).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = d.astype(np.float32)
    k = n // 10
    d[:k] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, k)] * rng.choice([-1.0, 1.0], k).astype(np.float32)[:, None]
    o[k:2 * k] = np.floor(o[k:2 * k])
    # the player's view: eye height over the terrain looking slightly down (what Pipeline.cpp:2044 casts every frame)
    o[2 * k:3 * k, 1] = rng.uniform(30, 70, k).astype(np.float32)
    return o, d
