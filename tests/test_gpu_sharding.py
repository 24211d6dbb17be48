"""Multi-GPU paths on real devices: the z-slab kernels (simulated on one GPU, then over NCCL when two GPUs
are visible) and the screen-tile band gather."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle import binding as ob
from voxeltracing_b200 import engine, host_api, sharding

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("nslabs", [2, 3, 8])
def test_slab_kernels_reproduce_the_full_field_on_one_gpu(nslabs):
    import torch

    blocks = host_api.gen_world("rooms", 2)
    c = engine.Context(0)
    c.upload_world(blocks)
    z0 = sharding.slab_bounds(384, nslabs)
    df = torch.as_tensor(c.df_device_array(), device="cuda:0")
    for s in range(nslabs):
        c.df_slab_phase_a(s, z0)
    c.synchronize()
    firsts = torch.stack([df[z0[s]].clone() for s in range(nslabs)])
    lasts = torch.stack([df[z0[s + 1] - 1].clone() for s in range(nslabs)])
    torch.cuda.synchronize()
    for s in range(nslabs):
        c.df_slab_phase_b(s, z0, firsts.data_ptr(), lasts.data_ptr())
    with pytest.raises(engine.VxrtError):
        c.download_distance_field()  # not committed yet
    c.df_commit()
    assert np.array_equal(c.download_distance_field(), ob.distance_field(blocks))
    c.close()


def test_slab_argument_checks():
    c = engine.Context(0)
    c.upload_world(np.zeros((384, 128, 384), np.uint8))
    with pytest.raises(engine.VxrtError):
        c.df_slab_phase_a(0, [0, 100, 300])       # does not cover the grid
    with pytest.raises(engine.VxrtError):
        c.df_slab_phase_a(2, [0, 192, 384])       # slab index out of range
    with pytest.raises(engine.VxrtError):
        c.df_slab_phase_a(0, [0, 384, 384])       # empty slab
    c.close()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_gpu_nccl_sharded_df_and_band_gather():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(ROOT / "tests" / "dist_worker_gpu.py")], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "RESULT ok" in outs[0], outs[0]
