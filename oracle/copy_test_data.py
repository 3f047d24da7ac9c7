#!/usr/bin/env python
"""TEST INFRASTRUCTURE. Copies the reference's own test scenes that the ingestion tests render
(tests/_version43/<scene>/<scene>.json with their PLY shapes and PNG / HDR textures, and the two shapes the
BASELINE configs C2 / C5 name: bunny.ply, hairball1.ply) from /root/reference into oracle/_ref/data/ — a build
directory: git-ignored, so nothing of the reference enters the history, but not gpurun-ignored, so the files
travel to the GPU box like the other oracle build products. Files shared by several scenes are stored once
(pool/shapes, pool/textures, under a content hash; scenes/<name>.files maps uris to them); tests/scene_data.py lays a
scene out again with symlinks.
usage: copy_test_data.py /root/reference oracle/_ref/data"""
import hashlib
import json
import os
import shutil
import sys

SCENES = ["features1", "materials1", "materials2", "materials3", "materials4", "materials5", "shapes1", "shapes4", "cornellbox",
          "arealights1", "environments1", "environments2", "furnace1", "furnace2", "instances1", "features2", "shapes2"]
EXTRA_SHAPES = ["bunny.ply", "hairball1.ply", "sphere.ply", "floor.ply", "arealight1.ply", "arealight2.ply"]
EXTRA_TEXTURES = ["sky.hdr", "floor.png"]


def main(ref, out):
    tests = os.path.join(ref, "tests")
    for sub in ("pool/shapes", "pool/textures", "pool/subdivs", "scenes"):
        os.makedirs(os.path.join(out, sub), exist_ok=True)

    def copy(src, dst):
        if not os.path.exists(dst) or os.path.getsize(dst) != os.path.getsize(src):
            shutil.copyfile(src, dst)

    # the files the BASELINE configs name keep their plain names; scene files with the same content reuse them
    by_hash = {}
    for kind, names, sub in (("shapes", EXTRA_SHAPES, "shapes"), ("textures", EXTRA_TEXTURES, "textures")):
        for f in names:
            src = os.path.join(tests, "_data", sub, f)
            copy(src, os.path.join(out, "pool", kind, f))
            by_hash[(kind, hashlib.sha1(open(src, "rb").read()).hexdigest())] = f

    def pooled(src, kind):
        """scene files go into the pool under a content hash (several scenes ship different files of one name)"""
        data = open(src, "rb").read()
        digest = hashlib.sha1(data).hexdigest()
        if (kind, digest) in by_hash:
            return by_hash[(kind, digest)]
        name = digest[:12] + "_" + os.path.basename(src)
        dst = os.path.join(out, "pool", kind, name)
        if not os.path.exists(dst):
            open(dst, "wb").write(data)
        return name

    for name in SCENES:
        src_dir = os.path.join(tests, "_version43", name)
        src = os.path.join(src_dir, name + ".json")
        scene = json.load(open(src))
        copy(src, os.path.join(out, "scenes", name + ".json"))
        files = {}
        for group in ("shapes", "textures", "subdivs"):
            for item in scene.get(group, []):
                uri = item["uri"]
                assert uri.startswith(group + "/"), uri
                files[uri] = pooled(os.path.join(src_dir, uri), group)
        json.dump(files, open(os.path.join(out, "scenes", name + ".files"), "w"), indent=0)
    total = sum(os.path.getsize(os.path.join(d, f)) for d, _, fs in os.walk(out) for f in fs)
    print(f"copy_test_data: {total / 1e6:.1f} MB under {out}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
