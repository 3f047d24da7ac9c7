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
# the same scenes in the older scene format 4.0 (objects keyed by name, files found by name): stored as v40_<name>
SCENES_V40 = ["arealights1", "cornellbox", "environments1", "environments2", "features1", "features2", "furnace1", "instances1",
              "materials1", "materials2", "materials3", "materials4", "materials5", "shapes1", "shapes2", "shapes3"]
EXTRA_SHAPES = ["bunny.ply", "hairball1.ply", "sphere.ply", "floor.ply", "arealight1.ply", "arealight2.ply"]
EXTRA_TEXTURES = ["sky.hdr", "floor.png"]


def main(ref, out):
    tests = os.path.join(ref, "tests")
    out40 = out.rstrip("/") + "_v40"    # format-4.0 scenes: host-only loader tests, listed in .gpurunignore (76 MB)
    for sub in ("pool/shapes", "pool/textures", "pool/subdivs", "scenes"):
        os.makedirs(os.path.join(out, sub), exist_ok=True)
    os.makedirs(os.path.join(out40, "scenes"), exist_ok=True)

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

    def pooled(src, kind, root=None):
        """scene files go into the pool under a content hash (several scenes ship different files of one name)"""
        data = open(src, "rb").read()
        digest = hashlib.sha1(data).hexdigest()
        if (kind, digest) in by_hash and root is None:
            return by_hash[(kind, digest)]
        name = digest[:12] + "_" + os.path.basename(src)
        dst = os.path.join(root or out, "pool", kind, name)
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
    for name in SCENES_V40:
        src_dir = os.path.join(tests, "_version40", name)
        src = os.path.join(src_dir, name + ".json")
        scene = json.load(open(src))
        copy(src, os.path.join(out40, "scenes", name + ".json"))
        mentioned = set()

        def walk(x):
            if isinstance(x, dict):
                for v in x.values():
                    walk(v)
            elif isinstance(x, str):
                mentioned.add(x)
        walk(scene)
        mentioned |= set(scene.get("subdivs", {}))          # a subdiv's file is named after its key
        files = {}
        for group in ("shapes", "textures", "subdivs", "instances"):
            os.makedirs(os.path.join(out40, "pool", group), exist_ok=True)
            d = os.path.join(src_dir, group)
            for f in sorted(os.listdir(d)) if os.path.isdir(d) else []:
                if os.path.splitext(f)[0] in mentioned and not f.endswith(".py"):
                    files[group + "/" + f] = pooled(os.path.join(d, f), group, out40)
        json.dump(files, open(os.path.join(out40, "scenes", name + ".files"), "w"), indent=0)
    # the reference's own renderings of six of these scenes (tests/_renderings, Radiance RGBE): coarse radiometric check
    os.makedirs(os.path.join(out40, "renderings"), exist_ok=True)
    for f in sorted(os.listdir(os.path.join(tests, "_renderings"))):
        if f.endswith("-mst.hdr"):
            copy(os.path.join(tests, "_renderings", f), os.path.join(out40, "renderings", f))
    for root in (out, out40):
        total = sum(os.path.getsize(os.path.join(d, f)) for d, _, fs in os.walk(root) for f in fs)
        print(f"copy_test_data: {total / 1e6:.1f} MB under {root}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
