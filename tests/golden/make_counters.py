"""Generates tests/golden/traversal_counters.json from the INSTRUMENTED reference
(oracle/_ref/libyocto_ref_count.so: the reference with counter statements at yocto_bvh.cpp:466,487,506-545,560,581,
621 — oracle/make_counted_bvh.py). One entry per BASELINE config rendered at the sample count of
tests/test_gpu_parity.py::FULL_SIZE; also digests of the reference render of those frames.
Run in the development container: python tests/golden/make_counters.py"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in ("yocto-gl_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import refbind  # noqa: E402
from ygl_b200 import abi, scenes  # noqa: E402

CONFIGS = {
    "c3": (lambda: scenes.instanced_spheres(10), dict(resolution=1920, samples=2, bounces=8, batch=2)),
    "c2": (lambda: scenes.bunny_like(6), dict(resolution=1280, samples=2, bounces=8, batch=2)),
    "c5": (scenes.hair_stress, dict(resolution=1920, samples=1, bounces=12, batch=1)),
}
ref = refbind.Ref("_count")
out = {}
for name, (factory, kw) in CONFIGS.items():
    scene = factory()
    rs = ref.scene(scene)
    ref.counters_reset()
    img = rs.trace_image(abi.trace_params(**kw), full=True)
    c = ref.counters()
    c["params"] = kw
    c["image_sha256"] = hashlib.sha256(img["image"].tobytes()).hexdigest()
    c["rngs_sha256"] = hashlib.sha256(img["rngs"].tobytes()).hexdigest()
    n = c["scene"]["rays"]
    c["per_scene_ray"] = {k: v / n for k, v in c["scene"].items()}
    out[name] = c
    print(name, c["per_scene_ray"])
json.dump(out, open(os.path.join(HERE, "traversal_counters.json"), "w"), indent=1)
