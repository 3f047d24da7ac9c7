"""Exploratory GPU check (development tool, run under gpurun): traversal + render parity against
the oracle and a first timing. Writes gpurun_out/gpu_check.json."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("yocto-gl_b200", "oracle", "tests", "."):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np  # noqa: E402
import refbind  # noqa: E402
from parity_util import axis_rays, compare_hits, image_stats, random_rays  # noqa: E402
from ygl_b200 import abi, lib, scenes  # noqa: E402

out = {}
ctx = lib.Context(0)
ref, refd = refbind.Ref(), refbind.Ref("_dlibm")

cases = {
    "cornell": scenes.cornellbox(),
    "cornellq": scenes.cornellbox_quads(),
    "inst4": scenes.instanced_spheres(4),
    "features": scenes.features(),
    "hair": scenes.hair_scene(4000, 8, 3),
}
for name, sc in cases.items():
    ds = lib.DeviceScene(ctx, sc)
    rs = ref.scene(sc)
    n = 200000
    rays = np.concatenate([random_rays(sc, n), axis_rays(sc, n // 10)])
    t = time.time()
    a = ds.intersect(rays)
    tg = time.time() - t
    b = rs.intersect(rays)
    bad = compare_hits(b, a)
    res = {"rays": len(rays), "hit_frac": float((b["hit"] != 0).mean()), "mismatch": bad, "gpu_s": tg}
    li = [i for i, n_ in enumerate(sc.instances) if any(np.asarray(sc.materials[n_["material"]]["emission"]) != 0)]
    if li:
        a2, b2 = ds.intersect(rays, instance=li[0]), rs.intersect(rays, instance=li[0])
        res["instance_mismatch"] = compare_hits(b2, a2)
    a3, b3 = ds.intersect(rays, find_any=True), rs.intersect(rays, find_any=True)
    res["any_mismatch"] = compare_hits(b3, a3)
    out["rays_" + name] = res
    print(name, res, flush=True)

renders = [
    ("cornell", dict(resolution=64, samples=4, bounces=4)),
    ("cornell", dict(resolution=128, samples=16, bounces=8)),
    ("cornellq", dict(resolution=64, samples=4, bounces=4)),
    ("inst4", dict(resolution=96, samples=4, bounces=8)),
    ("features", dict(resolution=128, samples=4, bounces=8)),
    ("features", dict(resolution=128, samples=4, bounces=8, camera=1)),
    ("features", dict(resolution=128, samples=2, bounces=8, sampler=abi.SAMPLER_EYELIGHT)),
    ("features", dict(resolution=128, samples=1, bounces=8, sampler=abi.SAMPLER_FALSECOLOR, falsecolor=abi.FC_NORMAL)),
    ("features", dict(resolution=128, samples=1, bounces=8, sampler=abi.SAMPLER_FALSECOLOR, falsecolor=abi.FC_ELEMENT)),
    ("features", dict(resolution=128, samples=1, bounces=8, sampler=abi.SAMPLER_FALSECOLOR, falsecolor=abi.FC_COLOR)),
    ("hair", dict(resolution=128, samples=4, bounces=8)),
    ("features", dict(resolution=128, samples=3, bounces=8, sampler=abi.SAMPLER_NAIVE)),
    ("features", dict(resolution=128, samples=3, bounces=8, sampler=abi.SAMPLER_FURNACE)),
    ("features", dict(resolution=128, samples=3, bounces=8, sampler=abi.SAMPLER_PATHDIRECT)),
    ("cornell", dict(resolution=64, samples=4, bounces=8, sampler=abi.SAMPLER_PATHDIRECT)),
    ("features", dict(resolution=128, samples=3, bounces=8, sampler=abi.SAMPLER_PATHMIS)),
    ("cornell", dict(resolution=64, samples=4, bounces=8, sampler=abi.SAMPLER_PATHMIS)),
]
for name, kw in renders:
    sc = cases[name]
    p = abi.trace_params(**kw)
    t = time.time()
    img = ctx.trace_image(sc, p)
    tg = time.time() - t
    cnt = ctx.counters()
    r1 = ref.scene(sc).trace_image(p)
    r2 = refd.scene(sc).trace_image(p)
    res = {"kw": {k: int(v) for k, v in kw.items()}, "gpu_s": tg, "ref_s": r1["seconds"], "vs_ref": image_stats(r1["image"], img),
           "vs_ref_dlibm": image_stats(r2["image"], img), "ref_vs_dlibm": image_stats(r1["image"], r2["image"]),
           "counters": cnt}
    out.setdefault("render_" + name, []).append(res)
    print(name, json.dumps(res), flush=True)

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w") as f:
    json.dump(out, f, indent=1)
