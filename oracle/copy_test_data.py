#!/usr/bin/env python
"""TEST INFRASTRUCTURE. Copies the reference's own test scenes that the ingestion tests render
(tests/_version43/<scene>/<scene>.json with their PLY shapes and PNG / HDR textures, and the two shapes the
BASELINE configs C2 / C5 name: bunny.ply, hairball1.ply) from /root/reference into oracle/_ref/data/ — a build
directory: git-ignored, so nothing of the reference enters the history, but not gpurun-ignored, so the files
travel to the GPU box like the other oracle build products. Files shared by several scenes are stored once
(pool/shapes, pool/textures); tests/scene_data.py lays a scene out again with symlinks.
usage: copy_test_data.py /root/reference oracle/_ref/data"""
import json
import os
import shutil
import sys

SCENES = ["features1", "materials1", "materials2", "materials3", "materials4", "shapes4", "cornellbox"]
EXTRA_SHAPES = ["bunny.ply", "hairball1.ply", "sphere.ply", "floor.ply", "arealight1.ply", "arealight2.ply"]
EXTRA_TEXTURES = ["sky.hdr", "floor.png"]


def main(ref, out):
    tests = os.path.join(ref, "tests")
    for sub in ("pool/shapes", "pool/textures", "scenes"):
        os.makedirs(os.path.join(out, sub), exist_ok=True)

    def copy(src, dst):
        if not os.path.exists(dst) or os.path.getsize(dst) != os.path.getsize(src):
            shutil.copyfile(src, dst)

    for name in SCENES:
        src_dir = os.path.join(tests, "_version43", name)
        src = os.path.join(src_dir, name + ".json")
        scene = json.load(open(src))
        copy(src, os.path.join(out, "scenes", name + ".json"))
        for group in ("shapes", "textures"):
            for item in scene.get(group, []):
                uri = item["uri"]
                assert uri.startswith(group + "/"), uri
                copy(os.path.join(src_dir, uri), os.path.join(out, "pool", uri))
    for f in EXTRA_SHAPES:
        copy(os.path.join(tests, "_data", "shapes", f), os.path.join(out, "pool", "shapes", f))
    for f in EXTRA_TEXTURES:
        copy(os.path.join(tests, "_data", "textures", f), os.path.join(out, "pool", "textures", f))
    total = sum(os.path.getsize(os.path.join(d, f)) for d, _, fs in os.walk(out) for f in fs)
    print(f"copy_test_data: {total / 1e6:.1f} MB under {out}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
