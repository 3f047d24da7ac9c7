"""Differential test of the glTF 2.0 reader (load_gltf_scene in ygl_sceneio.cpp) against the reference's
(yocto_sceneio.cpp:4430, over cgltf): seeded random files from the writer below - .gltf with an external .bin, with a
base64 data uri, and .glb; attributes as float or normalised / plain integers of every component type, interleaved with a
byteStride or packed, accessors without data; indexed (u8 / u16 / u32) and non-indexed triangles, strips, fans, lines,
line strips and loops; node hierarchies with TRS or matrix transforms; perspective and orthographic cameras; materials
with base colour / alpha, metallic, roughness, emission and its strength, transmission, textures - must give
bit-identical scenes (every array, including the procedural sky both loaders append) or be refused by both.
Host-only; the files are read in a worker process (tests/loader_worker.py)."""
import base64
import json
import os
import random
import struct

import numpy as np
import pytest

from loader_worker import LoaderPair
from test_texture_fuzz import make_png

COMPONENT = {5120: "b", 5121: "B", 5122: "h", 5123: "H", 5125: "I", 5126: "f"}


class Writer:
    def __init__(self, rng):
        self.rng, self.blob, self.views, self.accessors = rng, bytearray(), [], []

    def accessor(self, values, ctype, kind, normalized=False, interleave_pad=0):
        """values: (count, components) array already in the component's range"""
        values = np.asarray(values)
        count, comps = values.shape
        fmt = "<" + COMPONENT[ctype] * comps
        row = struct.calcsize(fmt)
        while len(self.blob) % 4:
            self.blob.append(0)
        offset = len(self.blob)
        stride = row + interleave_pad
        for r in values:
            self.blob += struct.pack(fmt, *(float(v) if ctype == 5126 else int(v) for v in r)) + bytes(interleave_pad)
        view = {"buffer": 0, "byteOffset": offset, "byteLength": len(self.blob) - offset}
        if interleave_pad:
            view["byteStride"] = stride
        self.views.append(view)
        acc = {"bufferView": len(self.views) - 1, "componentType": ctype, "count": count, "type": kind}
        if self.rng.random() < 0.3:
            acc["byteOffset"] = 0
        if normalized:
            acc["normalized"] = True
        self.accessors.append(acc)
        return len(self.accessors) - 1


def make_gltf(rng, d, stem):
    w = Writer(rng)
    doc = {"asset": {"version": "2.0"}}
    nimages = rng.randint(0, 2)
    if nimages:
        os.makedirs(d / "tex", exist_ok=True)
        doc["images"], doc["textures"] = [], []
        for k in range(nimages):
            name = f"{stem} img{k}.png" if rng.random() < 0.3 else f"{stem}_img{k}.png"
            data = make_png(rng, rng.randint(1, 6), rng.randint(1, 5), rng.choice([2, 6]), 8, 0)
            if rng.random() < 0.5:      # glTF assets mostly carry JPEG textures
                try:
                    import io
                    from PIL import Image
                    out = io.BytesIO()
                    jw, jh = rng.randint(1, 20), rng.randint(1, 20)
                    pixels = np.array([[[rng.randrange(256) for _ in range(3)] for _ in range(jw)] for _ in range(jh)], np.uint8)
                    Image.fromarray(pixels, "RGB").save(out, "JPEG", quality=rng.choice([30, 80, 95]), progressive=rng.random() < 0.3,
                                                        subsampling=rng.choice([0, 1, 2]))
                    name, data = name[:-4] + ".jpg", out.getvalue()
                except ImportError:
                    pass
            (d / "tex" / name).write_bytes(data)
            doc["images"].append({"uri": "tex/" + name.replace(" ", "%20")})
            doc["textures"].append({"source": k})
    view = lambda: {"index": rng.randrange(nimages)} if nimages and rng.random() < 0.5 else None
    materials = []
    for _ in range(rng.randint(0, 3)):
        m = {}
        if rng.random() < 0.8:
            pbr = {}
            if rng.random() < 0.7:
                pbr["baseColorFactor"] = [round(rng.random(), 3) for _ in range(4)]
            if rng.random() < 0.5:
                pbr["metallicFactor"] = round(rng.random(), 3)
            if rng.random() < 0.5:
                pbr["roughnessFactor"] = round(rng.random(), 3)
            for key in ("baseColorTexture", "metallicRoughnessTexture"):
                if (v := view()) is not None:
                    pbr[key] = v
            m["pbrMetallicRoughness"] = pbr
        if rng.random() < 0.5:
            m["emissiveFactor"] = [round(rng.random(), 3) for _ in range(3)]
        for key in ("emissiveTexture", "normalTexture"):
            if (v := view()) is not None:
                m[key] = v
        ext = {}
        if rng.random() < 0.3:
            ext["KHR_materials_emissive_strength"] = {"emissiveStrength": rng.choice([2, 5.5, 0.25])} if rng.random() < 0.8 else {}
        if rng.random() < 0.3:
            t = {"transmissionFactor": rng.choice([0, 0.5, 1])}
            if (v := view()) is not None:
                t["transmissionTexture"] = v
            ext["KHR_materials_transmission"] = t
        if ext:
            m["extensions"] = ext
        materials.append(m)
    if materials:
        doc["materials"] = materials
    meshes = []
    for _ in range(rng.randint(1, 3)):
        primitives = []
        for _ in range(rng.randint(1, 2)):
            nv = rng.randint(3, 12)
            pad = rng.choice([0, 0, 4, 8])
            attributes = {}
            kind = rng.random()
            if kind < 0.6:
                attributes["POSITION"] = w.accessor(np.round(np.array([[rng.uniform(-2, 2) for _ in range(3)] for _ in range(nv)]), 4), 5126, "VEC3", interleave_pad=pad)
            elif kind < 0.8:
                attributes["POSITION"] = w.accessor([[rng.randint(-32767, 32767) for _ in range(3)] for _ in range(nv)], 5122, "VEC3", normalized=rng.random() < 0.7)
            else:
                attributes["POSITION"] = w.accessor([[rng.randint(0, 255) for _ in range(3)] for _ in range(nv)], rng.choice([5121, 5120 if False else 5121]), "VEC3", normalized=rng.random() < 0.5)
            if rng.random() < 0.5:
                attributes["NORMAL"] = w.accessor(np.round(np.array([[rng.uniform(-1, 1) for _ in range(3)] for _ in range(nv)]), 4), 5126, "VEC3")
            if rng.random() < 0.5:
                key = rng.choice(["TEXCOORD_0", "TEXCOORD_0", "TEXCOORD"])
                if rng.random() < 0.5:
                    attributes[key] = w.accessor(np.round(np.array([[rng.random() for _ in range(2)] for _ in range(nv)]), 4), 5126, "VEC2", interleave_pad=rng.choice([0, 4]))
                else:
                    ctype = rng.choice([5121, 5123])
                    attributes[key] = w.accessor([[rng.randint(0, 255 if ctype == 5121 else 65535) for _ in range(2)] for _ in range(nv)], ctype, "VEC2", normalized=True)
            if rng.random() < 0.4:
                comps = rng.choice([3, 4])
                ctype = rng.choice([5126, 5121, 5123])
                hi = {5126: 1, 5121: 255, 5123: 65535}[ctype]
                vals = [[round(rng.random(), 3) if ctype == 5126 else rng.randint(0, hi) for _ in range(comps)] for _ in range(nv)]
                attributes[rng.choice(["COLOR_0", "COLOR"])] = w.accessor(vals, ctype, "VEC%d" % comps, normalized=ctype != 5126)
            if rng.random() < 0.2:
                attributes["RADIUS"] = w.accessor([[round(rng.uniform(0.001, 0.1), 4)] for _ in range(nv)], 5126, "SCALAR")
            if rng.random() < 0.2:
                attributes["TANGENT"] = w.accessor([[0.0, 1.0, 0.0, 1.0] for _ in range(nv)], 5126, "VEC4")
            if rng.random() < 0.2:
                attributes["_CUSTOM"] = w.accessor([[1.0] for _ in range(nv)], 5126, "SCALAR")
            if rng.random() < 0.1:      # an accessor without data reads as zeros
                w.accessors.append({"componentType": 5126, "count": nv, "type": "VEC3"})
                attributes["NORMAL"] = len(w.accessors) - 1
            prim = {"attributes": attributes}
            mode = rng.choice([None, 4, 4, 5, 6, 1, 3, 2])
            if mode is not None:
                prim["mode"] = mode
            if rng.random() < 0.7 and mode != 2:
                ni = rng.randint(3, 15)
                ctype = rng.choice([5121, 5123, 5125])
                prim["indices"] = w.accessor([[rng.randrange(nv)] for _ in range(ni)], ctype, "SCALAR")
            if materials and rng.random() < 0.7:
                prim["material"] = rng.randrange(len(materials))
            primitives.append(prim)
        meshes.append({"primitives": primitives})
    doc["meshes"] = meshes
    cameras = []
    for _ in range(rng.randint(0, 2)):
        if rng.random() < 0.7:
            p = {"yfov": round(rng.uniform(0.3, 1.5), 4), "znear": 0.1}
            if rng.random() < 0.6:
                p["aspectRatio"] = rng.choice([1.0, 1.7778, 0.5, 2.4])
            cameras.append({"type": "perspective", "perspective": p})
        else:
            cameras.append({"type": "orthographic", "orthographic": {"xmag": rng.choice([1.0, 2.5]), "ymag": rng.choice([1.0, 0.75]), "znear": 0.1, "zfar": 10}})
    if cameras:
        doc["cameras"] = cameras
    nodes = []
    nnodes = rng.randint(1, 7)
    for k in range(nnodes):
        node = {}
        r = rng.random()
        if r < 0.3:
            c, s = np.cos(a := rng.uniform(0, 6.28)), np.sin(a)
            node["matrix"] = [round(float(v), 5) for v in [c, 0, -s, 0, 0, rng.choice([1, 2]), 0, 0, s, 0, c, 0, rng.uniform(-3, 3), rng.uniform(-3, 3), rng.uniform(-3, 3), 1]]
        elif r < 0.85:
            if rng.random() < 0.7:
                node["translation"] = [round(rng.uniform(-3, 3), 4) for _ in range(3)]
            if rng.random() < 0.6:
                q = np.array([rng.uniform(-1, 1) for _ in range(4)])
                q /= np.linalg.norm(q)
                node["rotation"] = [round(float(v), 6) for v in q]
            if rng.random() < 0.4:
                node["scale"] = [rng.choice([1, 2, 0.5, -1]) for _ in range(3)]
        if rng.random() < 0.75:
            node["mesh"] = rng.randrange(len(meshes))
        if cameras and rng.random() < 0.4:
            node["camera"] = rng.randrange(len(cameras))
        nodes.append(node)
    for k in range(1, nnodes):              # a forest: every node but the first may hang under an earlier one
        if rng.random() < 0.6:
            nodes[rng.randrange(k)].setdefault("children", []).append(k)
    doc["nodes"] = nodes
    doc["scenes"], doc["scene"] = [{"nodes": [0]}], 0
    doc["accessors"], doc["bufferViews"] = w.accessors, w.views
    blob = bytes(w.blob)
    container = rng.choice(["bin", "base64", "glb"])
    if container == "glb":
        doc["buffers"] = [{"byteLength": len(blob)}]
        text = json.dumps(doc).encode()
        text += b" " * (-len(text) % 4)
        binary = blob + bytes(-len(blob) % 4)
        data = struct.pack("<III", 0x46546C67, 2, 12 + 8 + len(text) + 8 + len(binary)) + struct.pack("<II", len(text), 0x4E4F534A) + text \
            + struct.pack("<II", len(binary), 0x004E4942) + binary
        path = d / (stem + ".gltf")     # (the reference picks the loader by extension and knows no ".glb"; cgltf goes by the magic)
        path.write_bytes(data)
        return path
    if container == "bin":
        name = stem + ".bin"
        (d / name).write_bytes(blob)
        doc["buffers"] = [{"byteLength": len(blob), "uri": name}]
    else:
        doc["buffers"] = [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}]
    path = d / (stem + ".gltf")
    path.write_text(json.dumps(doc))
    return path


@pytest.mark.parametrize("seed", [1, 2])
def test_random_gltf_files_load_like_the_reference(ref, seed, tmp_path):
    rng = random.Random(seed)
    counts = {"same": 0, "refused": 0, "reference crashed": 0}
    loaders = LoaderPair()
    for k in range(100):
        path = make_gltf(rng, tmp_path, f"s{k}")
        verdict = loaders.verdict(path)
        assert verdict in counts, f"file {k} (seed {seed}): {verdict}\n{path.read_bytes()[:3000]!r}"
        counts[verdict] += 1
    loaders.close()
    assert counts["same"] >= 60, counts


def test_gltf_files_the_reference_refuses(ref, tmp_path):
    """sparse accessors, point primitives, a camera without a known type, a missing buffer file, broken JSON: refused by
    both loaders. And the reference-side drop-in (oracle/shim_load_demo) on a few accepted files: scene_data identical."""
    import subprocess
    import scene_data
    rng = random.Random(4)
    loaders = LoaderPair()
    good = make_gltf(rng, tmp_path, "good")
    assert loaders.verdict(good) == "same"
    base = {"asset": {"version": "2.0"}, "buffers": [{"byteLength": 36, "uri": "tri.bin"}],
            "bufferViews": [{"buffer": 0, "byteLength": 36}],
            "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3"}],
            "meshes": [{"primitives": [{"attributes": {"POSITION": 0}}]}], "nodes": [{"mesh": 0}]}
    (tmp_path / "tri.bin").write_bytes(struct.pack("<9f", 0, 0, 0, 1, 0, 0, 0, 1, 0))

    def variant(name, edit):
        doc = json.loads(json.dumps(base))
        edit(doc)
        path = tmp_path / (name + ".gltf")
        path.write_text(json.dumps(doc) if name != "broken" else json.dumps(doc)[:-3])
        return path
    assert loaders.verdict(variant("plain", lambda d: None)) == "same"
    cases = {
        "sparse": lambda d: d["accessors"][0].update(sparse={"count": 1, "indices": {"bufferView": 0, "componentType": 5121}, "values": {"bufferView": 0}}),
        "points": lambda d: d["meshes"][0]["primitives"][0].update(mode=0),
        "camera": lambda d: (d.update(cameras=[{"type": "fisheye"}]), d["nodes"][0].update(camera=0)),
        "nofile": lambda d: d["buffers"][0].update(uri="missing.bin"),
        "vec2pos": lambda d: d["accessors"][0].update(type="VEC2"),
        "broken": lambda d: None,
    }
    for name, edit in cases.items():
        assert loaders.verdict(variant(name, edit)) == "refused", name
    loaders.close()
    exe = os.path.join(scene_data.ROOT, "oracle", "_ref", "shim_load_demo")
    if os.path.exists(exe):
        files = [str(good), str(tmp_path / "plain.gltf")] + [str(make_gltf(rng, tmp_path, f"shim{k}")) for k in range(4)]
        res = subprocess.run([exe] + files, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
        assert res.returncode == 0 and res.stdout.count("identical") == len(files), res.stdout[-2000:]
