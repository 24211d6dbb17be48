"""A/B timing of the distance-field kernels for every kernel version selected by set_option("df_xyver" / "df_zver"), each checked
bit for bit against the oracle.  CUDA events on this box tick in ~2 us steps, so one event pair brackets a ROUND of NCTX regenerations,
each on its own context (own grids): with 8 contexts x 37.7 MB cycling through a 126 MB L2, every regeneration finds its block grid
evicted (cold, like the L2-flushed single measurement) and the per-regeneration time has sub-microsecond resolution."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
from oracle import binding as ob
from voxeltracing_b200 import engine, host_api

xyvers = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1").split(",")]
zvers = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "1,2").split(",")]
NCTX = 8
dbgs = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "0").split(",")]
worlds = {"plains0": host_api.gen_world("plains", 0), "rooms2": host_api.gen_world("rooms", 2)}
want = {k: ob.distance_field(v) for k, v in worlds.items()}
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctxs = [engine.Context(0) for _ in range(NCTX)]
for c in ctxs:
    c.set_stream(stream.cuda_stream)
ctx = ctxs[0]


def timed(rounds=12, reps=4):
    """The round (reps x NCTX regenerations) is captured into a CUDA graph, so the host's launch rate (about 5 us per launch through
    ctypes) is out of the measurement and the kernels run back to back."""
    for c in ctxs:
        c.generate_distance_field()   # attributes / lazy loading outside the capture
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream):
        for _ in range(reps):
            for c in ctxs:
                c.generate_distance_field()
    ts = []
    for _ in range(rounds):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        g.replay()
        b.record(stream); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / (NCTX * reps))
    return np.array(ts[2:])


for xv in xyvers:
  for dbg in dbgs:
    for zv in zvers:
        ok = True
        for c in ctxs:
            c.set_option("df_xyver", xv); c.set_option("df_zver", zv); c.set_option("df_stage", 0); c.set_option("df_dbg", dbg)
        for name, w in worlds.items():
            ctx.upload_world(w); ctx.generate_distance_field()
            ok = ok and np.array_equal(ctx.download_distance_field(), want[name])
        for c in ctxs:
            c.upload_world(worlds["plains0"])
        res = {}
        for stage in (0, 1):
            for c in ctxs:
                c.set_option("df_stage", stage)
            res[stage] = timed()
        for c in ctxs:
            c.set_option("df_stage", 0)
        t = res[0]
        print(f"xyver={xv} dbg={dbg} zver={zv} bit-exact={ok}: both {np.median(t):.2f} us (min {t.min():.2f})  XY alone {np.median(res[1]):.2f}"
              f"  -> {2 * worlds['plains0'].size / np.median(t) / 1e3:.0f} GB/s algorithmic", flush=True)
for c in ctxs:
    c.close()
