"""Scene ingestion (ygl_scene_load: JSON v4.2 + PLY + PNG / HDR, yocto-gl_b200/csrc/ygl_sceneio.cpp) against the
reference's own load_scene (yocto_sceneio.cpp:2761) on the reference's own test scenes: every array of scene_data
must come out bit-identical — cameras, instances, materials, environments, texture pixels, shape elements and vertex
data. Host-only: runs without a GPU. Needs oracle/_ref/data (oracle/copy_test_data.py) and the reference oracle."""
import json
import os

import numpy as np
import pytest

import scene_data
from ygl_b200 import lib

pytestmark = pytest.mark.skipif(not scene_data.available(), reason="oracle/_ref/data not present (run `make -C oracle data`)")


def assert_scenes_identical(a, b):
    assert (len(a.cameras), len(a.instances), len(a.materials), len(a.environments), len(a.shapes), len(a.textures)) == \
           (len(b.cameras), len(b.instances), len(b.materials), len(b.environments), len(b.shapes), len(b.textures))

    def same(x, y, what):
        if isinstance(x, dict):
            assert x.keys() == y.keys(), what
            for k in x:
                same(x[k], y[k], f"{what}.{k}")
        elif isinstance(x, np.ndarray):
            assert x.shape == y.shape and x.dtype == y.dtype and x.tobytes() == y.tobytes(), what
        elif isinstance(x, (tuple, list)):
            assert np.asarray(x, np.float32).tobytes() == np.asarray(y, np.float32).tobytes(), what
        elif isinstance(x, float):
            assert np.float32(x).tobytes() == np.float32(y).tobytes(), what
        else:
            assert x == y, what

    for group in ("cameras", "instances", "materials", "environments", "shapes", "textures"):
        for k, (x, y) in enumerate(zip(getattr(a, group), getattr(b, group))):
            same(x, y, f"{group}[{k}]")


@pytest.mark.parametrize("name", scene_data.names())
def test_loader_matches_reference_loader(ref, name, tmp_path):
    path = scene_data.scene_file(name, tmp_path)
    ours = lib.load_scene(path)
    theirs = ref.load_scene(path)
    assert_scenes_identical(ours, theirs)
    want = [c.get("name", "") for c in json.load(open(path)).get("cameras", [])]
    assert ours.camera_names == (want or ["camera"])


@pytest.mark.parametrize("name", scene_data.names_v40() or ["none"])
def test_loader_matches_reference_loader_format40(ref, name, tmp_path):
    """Scene format 4.0 (tests/_version40: groups keyed by name, shapes and textures created by mention, files found by
    trying extensions, the older material labels, lookat without the x / z flip)."""
    if name == "none":
        pytest.skip("oracle/_ref/data_v40 not present (run `make -C oracle data`)")
    path = scene_data.scene_file(name, tmp_path, scene_data.DATA_V40)
    ours = lib.load_scene(path)
    theirs = ref.load_scene(path)
    assert_scenes_identical(ours, theirs)
    assert ours.camera_names == list(json.load(open(path))["cameras"])


def test_loader_errors(tmp_path):
    with pytest.raises(lib.YglError):
        lib.load_scene(tmp_path / "missing.json")
    bad = tmp_path / "bad.json"
    bad.write_text('{"asset": {"version": "4.2"}, "cameras": [{"lens": "wide"}]}')
    with pytest.raises(lib.YglError):  # a key of the wrong type is a parse error, as in the reference
        lib.load_scene(bad)
    old = tmp_path / "old.json"
    old.write_text('{"asset": {"version": "4.1"}}')
    with pytest.raises(lib.YglError):
        lib.load_scene(old)
    sub = tmp_path / "sub.json"
    sub.write_text('{"asset": {"version": "4.2"}, "subdivs": [{"uri": "subdivs/x.obj"}]}')
    with pytest.raises(lib.YglError):
        lib.load_scene(sub)
    noshape = tmp_path / "noshape.json"
    noshape.write_text('{"asset": {"version": "4.2"}, "shapes": [{"uri": "shapes/none.ply"}]}')
    with pytest.raises(lib.YglError):
        lib.load_scene(noshape)


def test_ascii_and_big_endian_ply_and_lookat(ref, tmp_path):
    """Formats the reference's test scenes do not use: ASCII and big-endian PLY, polygons (fans), mixed triangle /
    quad faces, polylines, vertex colours without alpha, `lookat` cameras, a scene without camera (add_missing_camera)
    and lines without radius (add_missing_radius) — all against the reference loader."""
    import struct
    d = tmp_path / "mini"
    (d / "shapes").mkdir(parents=True)
    (d / "shapes" / "poly.ply").write_text(
        "ply\nformat ascii 1.0\ncomment test\nelement vertex 6\nproperty float x\nproperty float y\nproperty float z\n"
        "property uchar red\nproperty uchar green\nproperty uchar blue\nproperty float s\nproperty float t\n"
        "element face 3\nproperty list uchar int vertex_indices\nend_header\n"
        "0 0 0 255 0 0 0 0\n1 0 0 0 255 0 1 0\n1 1 0 0 0 255 1 1\n0 1 0 10 20 30 0 1\n0.5 1.5 0.25 1 2 3 0.5 0.75\n-0.5 0.5 1e-3 4 5 6 0.125 0.0625\n"
        "3 0 1 2\n4 0 1 2 3\n5 0 1 2 4 5\n")
    verts = [(0.0, 0.0, 0.0), (1.0, 0.5, 0.25), (2.0, 1.5, -0.125), (3.0, 0.0, 7.5)]
    with open(d / "shapes" / "lines.ply", "wb") as f:
        f.write(b"ply\nformat binary_big_endian 1.0\nelement vertex 4\nproperty double x\nproperty double y\nproperty double z\n"
                b"element line 2\nproperty list uchar ushort vertex_indices\nend_header\n")
        for v in verts:
            f.write(struct.pack(">ddd", *v))
        f.write(struct.pack(">BHHH", 3, 0, 1, 2) + struct.pack(">BHH", 2, 2, 3))
    (d / "mini.json").write_text(json.dumps({
        "asset": {"version": "4.2", "copyright": "test"},
        "materials": [{"name": "m", "type": "glossy", "color": [0.1, 0.2, 0.3], "roughness": 0.25}, {"type": "nonsense"}],
        "shapes": [{"name": "poly", "uri": "shapes/poly.ply"}, {"name": "lines", "uri": "shapes/lines.ply"}],
        "instances": [{"shape": 0, "material": 0}, {"shape": 1, "material": 1, "lookat": [1, 2, 3, 0, 0.5, 0, 0, 1, 0]}],
        "environments": [{"emission": [1, 2, 3], "lookat": [0, 0, 0, 0, 0, -1, 0, 1, 0]}]}))
    assert_scenes_identical(lib.load_scene(d / "mini.json"), ref.load_scene(d / "mini.json"))
    scene = json.loads((d / "mini.json").read_text())
    scene["cameras"] = [{"name": "c", "lookat": [3, 2, 5, 0.5, 0.5, 0, 0, 1, 0], "aspect": 2.0, "orthographic": True}]
    (d / "mini2.json").write_text(json.dumps(scene))
    os.symlink(d / "shapes", tmp_path / "shapes2")
    assert_scenes_identical(lib.load_scene(d / "mini2.json"), ref.load_scene(d / "mini2.json"))


OBJ_MIXED = """# quads and a pentagon, shared and split vertices, negative indices, a polyline and points
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0.25
v 0.5 1.5 0.125
v 2 0 -1
vn 0 0 1
vn 0 0.6 0.8
vt 0 0
vt 1 0
vt 1 1
vt 0.25 0.75
f 1/1/1 2/2/1 3/3/1 4/4/2
f 4/4/2 3/3/1 5/1/2
f -6/1/1 -5/2/1 -1/2/2 3/3/1 5/4/2
l 1 2 6
p 6 1
"""
OBJ_TRIS = """v -1 0 0.5
v 1 0 0.5
v 0 1.0e0 0.5
v 0 -1 .5
vt 0.1 0.9
f 1 2 3
f 1//1 4//1 2//1
vn 0 0 1
f 2/1 3/1 1/1
"""


@pytest.mark.parametrize("name, text", [("mixed", OBJ_MIXED), ("tris", OBJ_TRIS)])
def test_obj_shapes_match_reference_loader(ref, tmp_path, name, text):
    """.obj shapes (load_shape, yocto_sceneio.cpp:1036-1051): vertex de-duplication in order of first use, quads as soon
    as one face has four corners, fans, the cursor rule of get_lines / get_points, texcoord flip."""
    os.makedirs(tmp_path / "shapes")
    (tmp_path / "shapes" / "a.obj").write_text(text)
    path = tmp_path / "scene.json"
    path.write_text(json.dumps({"asset": {"version": "4.2"}, "shapes": [{"name": "a", "uri": "shapes/a.obj"}],
                                "materials": [{"name": "m", "color": [0.5, 0.5, 0.5]}],
                                "instances": [{"name": "i", "shape": 0, "material": 0}]}))
    ours, theirs = lib.load_scene(str(path)), ref.load_scene(str(path))
    assert_scenes_identical(ours, theirs)
    assert len(ours.shapes[0]["positions"]) > 0


def _write_tri_ply(path):
    path.write_text("ply\nformat ascii 1.0\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\n"
                    "element face 1\nproperty list uchar int vertex_indices\nend_header\n0 0 0\n1 0 0.5\n0 1 0.25\n3 0 1 2\n")


def test_format40_objects_and_instance_lists(ref, tmp_path):
    """What the reference's 4.0 test scenes do not use: an "objects" group whose members carry an `instance` list
    (instances/<name>.ply: one frame per row, new frame = row * object frame, names <object>_<k>), `ortho`, a material
    named before it is... never defined (an error), groups given as arrays (keys are the indices), booleans where numbers
    are expected, arrays longer than needed, lookat on every kind of element."""
    import random
    d = tmp_path / "old"
    for sub in ("shapes", "instances"):
        (d / sub).mkdir(parents=True)
    _write_tri_ply(d / "shapes" / "tri.ply")
    (d / "shapes" / "quad.obj").write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nf 1 2 3 4\n")
    rng = random.Random(5)
    rows = [" ".join(repr(round(rng.uniform(-2, 2), 4)) for _ in range(12)) for _ in range(7)]
    (d / "instances" / "grid.ply").write_text(
        "ply\nformat ascii 1.0\nelement instance 7\n" + "".join(f"property float {n}\n" for n in
        ("xx", "xy", "xz", "yx", "yy", "yz", "zx", "zy", "zz", "ox", "oy", "oz")) + "end_header\n" + "\n".join(rows) + "\n")
    doc = {
        "asset": {"copyright": "old"},
        "objects": {
            "many": {"shape": "tri", "material": "red", "instance": "grid", "frame": [0.5, 0.1, 0, -0.1, 0.5, 0, 0, 0, 2, 1, 2, 3]},
            "one": {"shape": "quad", "material": "vol", "lookat": [1, 2, 3, 0, 0.5, 0, 0, 1, 0, 9, 9]},
            "again": {"shape": "tri", "instance": ""}},
        "instances": {"plain": {"shape": "quad", "material": "red", "instance": "grid"}},
        "materials": {"red": {"type": "metallic", "color": [1, 0, 0], "roughness": True},
                      "vol": {"type": "volume", "scattering": [0.5, 0.25, 0.125], "trdepth": 0.5},
                      "odd": {"type": "reflective"}},
        "cameras": [{"ortho": True, "lookat": [3, 2, 5, 0.5, 0.5, 0, 0, 1, 0], "lens": 0.1}, {"film": 0.024}],
        "environments": {"sky": {"emission": [1, 2, 3], "lookat": [0, 0, 0, 0, 0, -1, 0, 1, 0]}}}
    path = d / "old.json"
    path.write_text(json.dumps(doc))
    ours, theirs = lib.load_scene(path), ref.load_scene(path)
    assert_scenes_identical(ours, theirs)
    assert len(ours.instances) == 1 + 7 + 1 + 1 and ours.camera_names == ["0", "1"]
    assert [m["type"] for m in ours.materials] == [2, 6, 0]
    # an instance that names a material the file does not define cannot be parsed - there as here
    doc["objects"]["one"]["material"] = "missing"
    bad = d / "bad.json"
    bad.write_text(json.dumps(doc))
    with pytest.raises(lib.YglError):
        lib.load_scene(bad)
    with pytest.raises(RuntimeError):
        ref.load_scene(bad)
    doc["objects"]["one"]["material"] = "vol"
    doc["objects"]["many"]["instance"] = "nofile"
    bad.write_text(json.dumps(doc))
    with pytest.raises(lib.YglError):
        lib.load_scene(bad)
    with pytest.raises(RuntimeError):
        ref.load_scene(bad)


def test_ply_file_as_scene_with_procedural_sky(ref, tmp_path):
    """`load_scene("x.ply")` (load_ply_scene, yocto_sceneio.cpp:4364): the shape, the default matte material, the framing
    camera and add_sky's 1024 x 512 Preetham sky (make_sunsky) as environment - every texel bit-identical."""
    src = scene_data.pool("shapes", "bunny.ply")
    path = tmp_path / "bunny.ply"
    os.symlink(src, path)
    ours, theirs = lib.load_scene(path), ref.load_scene(path)
    assert_scenes_identical(ours, theirs)
    sky = ours.textures[0]
    assert sky["pixels"].shape == (512, 1024, 4) and sky["pixels"].dtype == np.float32 and sky["linear"]
    assert float(sky["pixels"][:256, :, :3].max()) > 0.05 and np.ptp(sky["pixels"][256:, :, :3], axis=(0, 1)).max() == 0
    tri = tmp_path / "tri.ply"
    _write_tri_ply(tri)
    assert_scenes_identical(lib.load_scene(tri), ref.load_scene(tri))
    with pytest.raises(lib.YglError):
        lib.load_scene(tmp_path / "scene.gltf")


@pytest.mark.parametrize("name", ["features1", "features2", "materials1", "materials2", "materials3"])
def test_oracle_agrees_with_the_references_published_renderings(ref, name, tmp_path):
    """The only image-level fixtures the reference ships for this path: tests/_renderings/<scene>-mst.hdr,
    1280 x 533 RGBE of an unknown sample count, made with a slightly different camera than the scene files hold today
    (the objects are a few percent larger), so pixels cannot be compared. What can: the mean radiance of the frame, which
    pins light intensities, the environment and the exposure of the whole pipeline. The oracle (the reference's code,
    compiled here) renders the scene file our loader read, 4 spp, and must land within 4 % of the published frame.
    (materials4-mst.hdr is left out: it shows two emissive volumes that today's materials4.json no longer contains.)"""
    from ygl_b200 import abi
    golden = os.path.join(scene_data.DATA_V40, "renderings", name + "-mst.hdr")
    if not os.path.exists(golden):
        pytest.skip("oracle/_ref/data_v40/renderings not present (run `make -C oracle data`)")
    os.makedirs(tmp_path / "g" / "textures")
    os.symlink(golden, tmp_path / "g" / "textures" / "g.hdr")
    (tmp_path / "g" / "g.json").write_text(json.dumps({"asset": {"version": "4.2"}, "textures": [{"name": "g", "uri": "textures/g.hdr"}]}))
    published = lib.load_scene(tmp_path / "g" / "g.json").textures[0]["pixels"][..., :3]     # our own RGBE reader
    scene = lib.load_scene(scene_data.scene_file(name, tmp_path))                                # our own scene reader
    image = ref.scene(scene).trace_image(abi.trace_params(resolution=1280, samples=4, bounces=8))["image"][..., :3]
    assert image.shape == published.shape == (533, 1280, 3)
    ours, theirs = float(np.minimum(image, 10).mean()), float(np.minimum(published, 10).mean())
    assert abs(ours - theirs) / theirs < 0.04, (ours, theirs)


def test_reference_side_load_scene_shim(tmp_path):
    """The C++ drop-in for `yocto::load_scene` (yocto::b200::load_scene, yocto-gl_b200/host/yocto_b200trace.h) in a
    reference-side program (oracle/shim_load_demo.cpp) next to the reference's own loader + tesselate_subdivs: the two
    scene_data objects - every array, every name - must be identical. Format 4.2 with subdivs, format 4.0, a .ply scene."""
    import subprocess
    exe = os.path.join(scene_data.ROOT, "oracle", "_ref", "shim_load_demo")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/shim_load_demo not built (needs the reference headers)")
    files = [scene_data.scene_file(name, tmp_path) for name in ("features2", "materials3", "shapes1")]
    if scene_data.names_v40():
        files += [scene_data.scene_file(name, tmp_path / "v40", scene_data.DATA_V40) for name in ("shapes2", "instances1")]
    files.append(scene_data.pool("shapes", "bunny.ply"))
    res = subprocess.run([exe] + [str(f) for f in files], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:]
    assert res.stdout.count("identical") == len(files), res.stdout[-2000:]


@pytest.mark.parametrize("name, low", [("furnace1", 0.48), ("furnace2", 0.44)])
def test_oracle_passes_the_references_furnace_scenes(ref, name, low, tmp_path):
    """The reference's analytic fixtures: furnace1 / furnace2 put energy-conserving materials inside a uniform
    0.5 environment and are meant to be rendered with `--sampler furnace`. No surface may add energy: every pixel stays at or
    below the environment radiance up to Monte-Carlo noise, rays that miss everything return exactly 0.5, and the frame mean
    sits just below 0.5 (what rough microfacet lobes lose to single scattering) - for the scene files as our loader read
    them, rendered by the oracle."""
    from ygl_b200 import abi
    scene = lib.load_scene(scene_data.scene_file(name, tmp_path))
    image = ref.scene(scene).trace_image(abi.trace_params(resolution=96, samples=64, bounces=32, sampler=abi.SAMPLER_FURNACE))["image"][..., :3]
    assert np.isfinite(image).all()
    assert low < float(image.mean()) <= 0.5005, float(image.mean())
    assert float(image.max()) < 0.75                       # 64 spp of noise above 0.5, never a gain
    corner = image[:4, :4]                                 # the corners of both scenes see only the environment
    assert np.all(corner == np.float32(0.5))
