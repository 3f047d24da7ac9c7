"""Pins the plain-C restatement (oracle/restate/ygl_oracle.c) against the REAL reference
(oracle/_ref) and against the committed golden fixtures: trees, hits and rendered images must be
bit-identical (same host libm, same operation order). No GPU involved."""
import json
import os

import numpy as np
import pytest

import restate
from parity_util import axis_rays, compare_hits, random_rays
from ygl_b200 import abi, scenes

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
pytestmark = pytest.mark.skipif(not restate.available(), reason="oracle/_build/libygl_oracle.so not built")

SCENES = {
    "cornell": scenes.cornellbox,
    "cornell_quads": scenes.cornellbox_quads,
    "instanced3": lambda: scenes.instanced_spheres(3),
    "features": scenes.features,
    "hair": lambda: scenes.hair_scene(1500, 6, 2),
}


@pytest.mark.parametrize("name", list(SCENES))
@pytest.mark.parametrize("hq", [False, True])
def test_trees_match_reference(ref, name, hq):
    if hq and name == "hair":
        pytest.skip("covered by the smaller scenes")
    scene = SCENES[name]()
    rs, mine = ref.scene(scene), restate.OracleScene(scene, hq)
    for shape in [-1] + list(range(len(scene.shapes))):
        a, b = rs.bvh_tree(shape, hq), mine.tree(shape)
        assert a[0].tobytes() == b[0].tobytes() and a[1].tobytes() == b[1].tobytes(), (name, shape)


@pytest.mark.parametrize("name", list(SCENES))
def test_hits_match_reference(ref, name):
    scene = SCENES[name]()
    rs, mine = ref.scene(scene), restate.OracleScene(scene)
    rays = np.concatenate([random_rays(scene, 20000), axis_rays(scene, 4000)])
    assert compare_hits(rs.intersect(rays), mine.intersect(rays)) == 0
    assert compare_hits(rs.intersect(rays, find_any=True), mine.intersect(rays, find_any=True)) == 0
    assert compare_hits(rs.intersect(rays[:5000], instance=0), mine.intersect(rays[:5000], instance=0)) == 0


def test_golden_rays():
    g = np.load(os.path.join(GOLDEN, "rays_features.npz"))
    mine = restate.OracleScene(scenes.features())
    out = mine.intersect(g["rays"].view(abi.RAY_DTYPE).reshape(-1))
    assert compare_hits(g["hits"].view(abi.ISEC_DTYPE).reshape(-1), out) == 0


def test_rng_table_and_known_answers(ref):
    kat = json.load(open(os.path.join(GOLDEN, "kat.json")))
    scene = scenes.cornellbox()
    mine = restate.OracleScene(scene)
    p = abi.trace_params(resolution=8)
    _, _, r = mine.state_rngs(p)
    assert [int(x) >> 1 for x in r[:3, 1]] == kat["pixel_seq_ids"]
    w, h, rr = ref.scene(scene).state_rngs(abi.trace_params(resolution=40, seed=77))
    w2, h2, r2 = mine.state_rngs(abi.trace_params(resolution=40, seed=77))
    assert (w, h) == (w2, h2) and np.array_equal(rr, r2)


@pytest.mark.parametrize("name,kw", [
    ("cornell", dict(resolution=64, samples=4, bounces=4)),
    ("cornell_quads", dict(resolution=48, samples=3, bounces=8)),
    ("instanced3", dict(resolution=64, samples=3, bounces=8)),
])
def test_images_bit_identical_to_reference(ref, name, kw):
    scene = SCENES[name]()
    mine = restate.OracleScene(scene)
    assert mine.supported()
    p = abi.trace_params(**kw)
    assert mine.trace_image(p).tobytes() == ref.scene(scene).trace_image(p)["image"].tobytes()


@pytest.mark.parametrize("sampler", [abi.SAMPLER_PATHDIRECT, abi.SAMPLER_PATHMIS, abi.SAMPLER_PATHTEST, abi.SAMPLER_NAIVE,
                                     abi.SAMPLER_EYELIGHT, abi.SAMPLER_DIAGRAM, abi.SAMPLER_FURNACE])
@pytest.mark.parametrize("name,kw", [
    ("cornell", dict(resolution=56, samples=3, bounces=6)),
    ("instanced3", dict(resolution=56, samples=2, bounces=8, envhidden=1)),
])
def test_other_samplers_bit_identical_to_reference(ref, name, kw, sampler):
    """The restated trace_pathdirect / pathmis / pathtest / naive / eyelight / diagram / furnace
    (yocto_trace.cpp:599-1338) against the reference's own, on the scenes the restatement supports."""
    scene = SCENES[name]()
    mine = restate.OracleScene(scene)
    assert mine.supported()
    p = abi.trace_params(sampler=sampler, **kw)
    assert mine.trace_image(p).tobytes() == ref.scene(scene).trace_image(p)["image"].tobytes()


@pytest.mark.parametrize("fc", range(18))
def test_falsecolor_bit_identical_to_reference(ref, fc):
    """trace_falsecolor (yocto_trace.cpp:1341-1419), every trace_falsecolor_type."""
    for name in ("cornell_quads", "instanced3"):
        scene = SCENES[name]()
        p = abi.trace_params(resolution=48, samples=1, sampler=abi.SAMPLER_FALSECOLOR, falsecolor=fc)
        assert restate.OracleScene(scene).trace_image(p).tobytes() == ref.scene(scene).trace_image(p)["image"].tobytes(), name


@pytest.mark.parametrize("sampler", [abi.SAMPLER_PATH, abi.SAMPLER_PATHDIRECT, abi.SAMPLER_PATHMIS, abi.SAMPLER_NAIVE,
                                     abi.SAMPLER_EYELIGHT, abi.SAMPLER_DIAGRAM, abi.SAMPLER_FURNACE])
def test_opacity_nocaustics_tentfilter_bit_identical_to_reference(ref, sampler):
    """Opacity pass-through (rng drawn only when opacity < 1, bounce not counted, <= 128 passes), the nocaustics
    roughness clamp and the tent pixel filter, for every sampler that has them."""
    scene = scenes.instanced_spheres(3)
    for k, m in enumerate(scene.materials):
        if k % 2 == 0 and not np.any(np.asarray(m["emission"])):
            m["opacity"] = 0.35 + 0.2 * (k % 3)
    mine = restate.OracleScene(scene)
    assert mine.supported()
    for extra in (dict(), dict(nocaustics=1, tentfilter=1, envhidden=1)):
        p = abi.trace_params(resolution=56, samples=3, bounces=8, sampler=sampler, **extra)
        assert mine.trace_image(p).tobytes() == ref.scene(scene).trace_image(p)["image"].tobytes(), extra


def _transmissive_scene(offset=0):
    """instanced_spheres(3) with its non-emissive materials turned into every transmission / volume flavour."""
    scene = scenes.instanced_spheres(3)
    flavours = [
        dict(type=abi.REFRACTIVE, roughness=0.0, ior=1.5, color=(0.9, 1.0, 0.9), trdepth=1.0),
        dict(type=abi.TRANSPARENT, roughness=0.2, ior=1.5, color=(1.0, 0.9, 0.8)),
        dict(type=abi.SUBSURFACE, roughness=0.2, ior=1.4, color=(0.9, 0.6, 0.5), scattering=(0.6, 0.4, 0.3),
             scanisotropy=0.3, trdepth=0.2),
        dict(type=abi.VOLUMETRIC, color=(0.7, 0.8, 0.9), scattering=(0.8, 0.8, 0.8), scanisotropy=-0.2, trdepth=0.6),
        dict(type=abi.REFRACTIVE, roughness=0.15, ior=1.33, color=(0.8, 0.9, 1.0), scattering=(0.2, 0.2, 0.2), trdepth=0.5),
        dict(type=abi.TRANSPARENT, roughness=0.0, ior=1.5, color=(0.9, 0.9, 1.0)),
    ]
    k = 0
    for m in scene.materials:
        if np.any(np.asarray(m["emission"])):
            continue
        for key, value in flavours[(k + offset) % len(flavours)].items():
            m[key] = value
        k += 1
    return scene


@pytest.mark.parametrize("sampler", [abi.SAMPLER_PATH, abi.SAMPLER_PATHDIRECT, abi.SAMPLER_PATHMIS, abi.SAMPLER_PATHTEST,
                                     abi.SAMPLER_NAIVE, abi.SAMPLER_EYELIGHT, abi.SAMPLER_DIAGRAM, abi.SAMPLER_FURNACE])
def test_transmission_and_volumes_bit_identical_to_reference(ref, sampler):
    """Transparent / refractive / subsurface / volumetric materials (rough and delta), the volume slot, transmittance
    sampling and phase-function scattering of trace_path / pathdirect / pathmis, and the in_volume flag of furnace."""
    for offset in (0, 3):  # 5 non-emissive materials, 6 flavours: two rotations cover them all
        scene = _transmissive_scene(offset)
        mine = restate.OracleScene(scene)
        assert mine.supported()
        p = abi.trace_params(resolution=64, samples=3, bounces=10, sampler=sampler)
        assert mine.trace_image(p).tobytes() == ref.scene(scene).trace_image(p)["image"].tobytes(), offset


@pytest.mark.parametrize("sampler", list(range(9)))
def test_features_scene_bit_identical_to_reference(ref, sampler):
    """scenes.features(): triangles, quads, lines, points; all 8 material types; opacity; vertex colors; color /
    roughness / normal / emission textures (byte sRGB and float); textured rotated environment; thin-lens and
    orthographic cameras - every sampler, bit for bit against the unmodified reference."""
    scene = scenes.features()
    mine = restate.OracleScene(scene)
    assert mine.supported()
    for extra in (dict(), dict(camera=1, nocaustics=1, tentfilter=1)):
        p = abi.trace_params(resolution=72, samples=2, bounces=8, sampler=sampler, **extra)
        assert mine.trace_image(p).tobytes() == ref.scene(scene).trace_image(p)["image"].tobytes(), extra


def test_restatement_rejects_what_it_does_not_cover():
    scene = SCENES["cornell"]()
    mine = restate.OracleScene(scene)
    for kw in (dict(sampler=9), dict(sampler=-1)):
        with pytest.raises(NotImplementedError):
            mine.trace_image(abi.trace_params(resolution=16, samples=1, **kw))


GOLDEN_SCENES = {"cornell": scenes.cornellbox, "features": scenes.features,
                 "instanced4": lambda: scenes.instanced_spheres(4), "hair": lambda: scenes.hair_scene(4000, 8, 3)}


def test_every_golden_render_and_checksum():
    """The committed golden renders (made by the real reference, tests/golden/make_golden.py: every sampler, all four
    test scenes) are reproduced bit for bit by the restatement - this pins it even where oracle/_ref is absent."""
    g = np.load(os.path.join(GOLDEN, "renders.npz"))
    cache = {}
    names = [k.split(".")[0] for k in g.files if k.endswith(".image")]
    assert len(names) >= 13
    for name in names:
        scene_name = str(g[f"{name}.scene"])
        if scene_name not in cache:
            cache[scene_name] = restate.OracleScene(GOLDEN_SCENES[scene_name]())
        kw = {k: int(v) for k, v in zip(g[f"{name}.param_names"], g[f"{name}.param_values"])}
        assert cache[scene_name].trace_image(abi.trace_params(**kw)).tobytes() == g[f"{name}.image"].tobytes(), name
    kat = json.load(open(os.path.join(GOLDEN, "kat.json")))
    img = cache["cornell"].trace_image(abi.trace_params(resolution=256, samples=16, bounces=4))
    assert abs(float(img[..., :3].astype(np.float64).sum()) - kat["cornell_256_16spp_4b_sum_rgb"]) < 1e-9
