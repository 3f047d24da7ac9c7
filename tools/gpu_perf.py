"""Development tool (run under gpurun): times trace_samples on one of the BASELINE configs.
usage: gpu_perf.py [config] [resolution] [spp] [repeats] [name=value ...]
  name=value: a scheduling option of the context (ygl_context_set_option), or
  mode=persistent | tile=rank,nranks (emulate one rank of an N-GPU run) | profile=1 (CUDA-event timing of extend)
  lib=<dir under yocto-gl_b200/> is read by run scripts through YGL_B200_LIB (python binding), not here."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("yocto-gl_b200", "oracle", "tests", "."):
    sys.path.insert(0, os.path.join(ROOT, p))
from ygl_b200 import abi, lib, scenes  # noqa: E402

pos = [a for a in sys.argv[1:] if "=" not in a]
opts = dict(a.split("=", 1) for a in sys.argv[1:] if "=" in a)
config = pos[0] if len(pos) > 0 else "c3"
res = int(pos[1]) if len(pos) > 1 else 1920
spp = int(pos[2]) if len(pos) > 2 else 8
reps = int(pos[3]) if len(pos) > 3 else 3

t0 = time.time()
if config == "c3":
    scene, bounces = scenes.instanced_spheres(10), 8
elif config == "c1":
    scene, bounces = scenes.cornellbox(), 4
elif config == "c2":
    scene, bounces = scenes.bunny_like(6), 8
elif config == "c5":
    scene, bounces = scenes.hair_stress(), 12
else:
    raise SystemExit("unknown config")
print(f"scene built in {time.time() - t0:.2f}s", flush=True)
ctx = lib.Context(0)
mode, tile, profile = opts.pop("mode", None), opts.pop("tile", None), opts.pop("profile", None)
if mode:
    ctx.set_mode(mode)
for k, v in opts.items():
    ctx.set_option(k, float(v))
t0 = time.time()
ds = lib.DeviceScene(ctx, scene)
print(f"bvh+lights+upload in {time.time() - t0:.2f}s", flush=True)
params = abi.trace_params(resolution=res, samples=spp * (reps + 1), bounces=bounces, batch=spp)
state = ds.make_state(params, interleave=tuple(int(x) for x in tile.split(","))) if tile else ds.make_state(params)
ctx.set_profiling(bool(profile), False)
for r in range(reps + 1):
    t0 = time.time()
    ds.trace_samples(state, params)
    ctx.synchronize()
    dt = time.time() - t0
    c = ctx.counters()
    n = state.width * state.num_rows * spp
    print(f"{config} {state.width}x{state.height} {spp}spp: {dt * 1e3:.1f} ms  {n / dt / 1e6:.2f} Msamples/s  "
          f"{(c['scene_rays'] + c['instance_rays']) / dt / 1e6:.1f} Mrays/s  rays/sample "
          f"{c['scene_rays'] / n:.2f}+{c['instance_rays'] / n:.2f}  iters {c['iterations']} launches {c['launches']}",
          ctx.timings() if profile else "", " ".join(sys.argv[5:]), flush=True)
