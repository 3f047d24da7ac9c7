"""Generates the committed golden fixtures from the REAL reference (oracle/_ref, built from
/root/reference in the development container). Run: python tests/golden/make_golden.py
  cornellbox_scene.npz  make_cornellbox() (yocto_scene.cpp:970) as flat arrays
  rays_features.npz     seeded rays on scenes.features() + reference scene_intersections
  renders.npz           small renders by the reference (glibc libm) and its double-libm twin
  kat.json              known-answer values (PCG, pixel seeds, Cornell checksum)
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in ("yocto-gl_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np  # noqa: E402
import refbind  # noqa: E402
from parity_util import axis_rays, random_rays  # noqa: E402
from ygl_b200 import abi, scenes  # noqa: E402

ref, refd = refbind.Ref(), refbind.Ref("_dlibm")
np.savez_compressed(os.path.join(HERE, "cornellbox_scene.npz"), **ref.cornellbox().to_npz_dict())

sc = scenes.features()
rays = np.concatenate([random_rays(sc, 6000), axis_rays(sc, 1000)])
hits = ref.scene(sc).intersect(rays)
np.savez_compressed(os.path.join(HERE, "rays_features.npz"), rays=rays.view(np.uint8), hits=hits.view(np.uint8))

cases = {
    "cornell_path": ("cornell", dict(resolution=48, samples=4, bounces=4)),
    "features_path": ("features", dict(resolution=64, samples=2, bounces=6)),
    "features_falsecolor": ("features", dict(resolution=64, samples=1, sampler=abi.SAMPLER_FALSECOLOR,
                                             falsecolor=abi.FC_ELEMENT)),
    # one case per remaining sampler of get_trace_sampler_func (yocto_trace.cpp:1422-1438)
    "features_pathdirect": ("features", dict(resolution=56, samples=2, bounces=6, sampler=abi.SAMPLER_PATHDIRECT)),
    "features_pathmis": ("features", dict(resolution=56, samples=2, bounces=6, sampler=abi.SAMPLER_PATHMIS)),
    "features_pathtest": ("features", dict(resolution=56, samples=2, bounces=6, sampler=abi.SAMPLER_PATHTEST)),
    "features_naive": ("features", dict(resolution=56, samples=2, bounces=6, sampler=abi.SAMPLER_NAIVE)),
    "features_eyelight": ("features", dict(resolution=56, samples=2, bounces=6, sampler=abi.SAMPLER_EYELIGHT)),
    "features_diagram": ("features", dict(resolution=56, samples=2, bounces=6, sampler=abi.SAMPLER_DIAGRAM)),
    "features_furnace": ("features", dict(resolution=56, samples=2, bounces=6, sampler=abi.SAMPLER_FURNACE)),
    "cornell_pathmis": ("cornell", dict(resolution=40, samples=3, bounces=5, sampler=abi.SAMPLER_PATHMIS)),
    "instanced_path": ("instanced4", dict(resolution=64, samples=2, bounces=8)),
    "hair_path": ("hair", dict(resolution=48, samples=2, bounces=8)),
}
factories = {"cornell": scenes.cornellbox, "features": scenes.features,
             "instanced4": lambda: scenes.instanced_spheres(4), "hair": lambda: scenes.hair_scene(4000, 8, 3)}
out = {}
for name, (scene_name, kw) in cases.items():
    scene = factories[scene_name]()
    p = abi.trace_params(**kw)
    out[f"{name}.image"] = ref.scene(scene).trace_image(p)["image"]
    out[f"{name}.image_dlibm"] = refd.scene(scene).trace_image(p)["image"]
    out[f"{name}.scene"] = np.array(scene_name)
    out[f"{name}.param_names"] = np.array(list(kw))
    out[f"{name}.param_values"] = np.array([int(v) for v in kw.values()])
np.savez_compressed(os.path.join(HERE, "renders.npz"), **out)

cornell = scenes.cornellbox()
p = abi.trace_params(resolution=256, samples=16, bounces=4)
img = ref.scene(cornell).trace_image(p)["image"]
f, st = ref.rng_floats(961748941, 1, 4)
_, _, rngs = ref.scene(cornell).state_rngs(abi.trace_params(resolution=8))
kat = {
    "pcg_make_rng_961748941_1": {"state": int(st[0]), "inc": int(st[1]), "floats": [float(x) for x in f]},
    "pixel_seq_ids": [int(x) >> 1 for x in rngs[:3, 1]],
    "cornell_256_16spp_4b_sum_rgb": float(img[..., :3].astype(np.float64).sum()),
}
json.dump(kat, open(os.path.join(HERE, "kat.json"), "w"), indent=1)
print(kat)
