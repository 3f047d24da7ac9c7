"""Differential tests of the host-side preparation on seeded random, mostly degenerate scenes, against the reference
itself (oracle/_ref): make_scene_bvh in both split modes (yocto_bvh.cpp:108-302, 364-396) - nodes and primitive order of
the instance tree and of every shape tree -, and make_trace_lights (yocto_trace.cpp:1528-1581) - light list and CDFs -
must be bit-identical. The generators aim at what decides node order and CDF bits: coincident and lattice centroids,
signed zeros, planar sets, magnitudes from 1e-44 to 3e38, NaN and infinite coordinates, zero-area elements, zero and
scaled-to-nothing instance frames, emission on every element type, byte and float environment textures. Host-only."""
import numpy as np
import pytest

from ygl_b200 import abi, lib


def _positions(rng, nv):
    mode = rng.choice(["uniform", "lattice", "coincident", "plane", "huge", "tiny", "signedzero", "clusters"])
    if mode == "uniform":
        pos = rng.uniform(-1, 1, (nv, 3))
    elif mode == "lattice":
        pos = rng.integers(-2, 3, (nv, 3)).astype(np.float64) * 0.5
    elif mode == "coincident":
        pos = np.tile(rng.uniform(-1, 1, (1, 3)), (nv, 1))
    elif mode == "plane":
        pos = rng.uniform(-1, 1, (nv, 3))
        pos[:, int(rng.integers(0, 3))] = 0.25
    elif mode == "huge":
        pos = rng.uniform(-1, 1, (nv, 3)) * rng.choice([1e18, 1e30, 3e38])
    elif mode == "tiny":
        pos = rng.uniform(-1, 1, (nv, 3)) * rng.choice([1e-30, 1e-40, 1e-44])
    elif mode == "signedzero":
        pos = rng.choice([0.0, -0.0, 1.0, -1.0], (nv, 3))
    else:
        pos = rng.integers(0, 2, (nv, 1)) * 10.0 + rng.uniform(-0.01, 0.01, (nv, 3))
    if rng.random() < 0.25:
        pos[rng.integers(0, nv), rng.integers(0, 3)] = rng.choice([np.nan, np.inf, -np.inf])
    with np.errstate(over="ignore"):
        return pos.astype(np.float32)


def _add_shape(rng, sc, radius_choices):
    nv, ne = int(rng.integers(3, 60)), int(rng.integers(1, 80))
    pos = _positions(rng, nv)
    kind = rng.choice(["tri", "quad", "line", "point"])
    if kind == "tri":
        sc.add_shape(triangles=rng.integers(0, nv, (ne, 3)), positions=pos)
    elif kind == "quad":
        quads = rng.integers(0, nv, (ne, 4))
        as_triangle = rng.random(ne) < 0.3
        quads[as_triangle, 3] = quads[as_triangle, 2]
        sc.add_shape(quads=quads, positions=pos)
    elif kind == "line":
        sc.add_shape(lines=rng.integers(0, nv, (ne, 2)), positions=pos, radius=rng.choice(radius_choices, nv).astype(np.float32))
    else:
        sc.add_shape(points=rng.integers(0, nv, ne), positions=pos, radius=rng.choice(radius_choices, nv).astype(np.float32))


def _frame(rng):
    frame = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], np.float32)
    r = rng.random()
    if r < 0.5:
        frame[9:] = rng.uniform(-3, 3, 3)
    elif r < 0.7:
        frame[9:] = rng.integers(-1, 2, 3)            # centroids that tie
    elif r < 0.85:
        c, s = np.cos(a := rng.uniform(0, 6.28)), np.sin(a)
        frame[:9] = [c, 0, s, 0, 1, 0, -s, 0, c]
        frame[9:] = rng.uniform(-3, 3, 3)
    else:
        frame[:9] *= rng.choice([0.0, 2.0, -1.0, 1e-20])
    return frame


@pytest.mark.parametrize("seed", [1, 2])
def test_bvh_build_matches_reference_on_random_degenerate_scenes(ref, seed):
    for it in range(60):
        rng = np.random.default_rng(seed * 100000 + it)
        sc = abi.Scene()
        sc.add_camera()
        material = sc.add_material(color=(0.5, 0.5, 0.5))
        nshapes = int(rng.integers(1, 4))
        for _ in range(nshapes):
            _add_shape(rng, sc, [0.0, 0.001, 0.1, 1.0])
        for _ in range(int(rng.integers(1, 40))):
            sc.add_instance(int(rng.integers(0, nshapes)), material, frame=_frame(rng))
        rs = ref.scene(sc)
        for highquality in (False, True):
            mine = lib.Bvh(sc, highquality)
            for shape in [-1] + list(range(nshapes)):
                n_ref, p_ref = rs.bvh_tree(shape, highquality)
                n_my, p_my = mine.tree(shape)
                assert n_ref.tobytes() == n_my.tobytes() and p_ref.tobytes() == p_my.tobytes(), (seed, it, highquality, shape)


@pytest.mark.parametrize("seed", [1, 2])
def test_lights_match_reference_on_random_scenes(ref, seed):
    for it in range(80):
        rng = np.random.default_rng(seed * 100000 + it)
        sc = abi.Scene()
        sc.add_camera()
        ntex = int(rng.integers(0, 3))
        for _ in range(ntex):
            w, h = int(rng.integers(1, 40)), int(rng.integers(1, 30))
            if rng.random() < 0.5:
                sc.add_texture(rng.uniform(0, rng.choice([1, 50, 1e6]), (h, w, 4)).astype(np.float32))
            else:
                sc.add_texture(rng.integers(0, 256, (h, w, 4)).astype(np.uint8))
        materials = []
        for _ in range(int(rng.integers(1, 5))):
            emission = rng.choice([0.0, 0.0, 1.0, 5.0, 1e-30], 3) if rng.random() < 0.7 else (0, 0, 0)
            materials.append(sc.add_material(color=(0.5, 0.5, 0.5), emission=tuple(float(x) for x in emission)))
        nshapes = int(rng.integers(1, 4))
        for _ in range(nshapes):
            _add_shape(rng, sc, [0.01])
        for _ in range(int(rng.integers(1, 10))):
            sc.add_instance(int(rng.integers(0, nshapes)), materials[int(rng.integers(0, len(materials)))])
        for _ in range(int(rng.integers(0, 3))):
            emission = rng.choice([0.0, 1.0, 2.5], 3) if rng.random() < 0.8 else (0, 0, 0)
            sc.add_environment(emission=tuple(float(x) for x in emission), emission_tex=int(rng.integers(-1, ntex)) if ntex else -1)
        l_ref, l_my = ref.scene(sc).lights(), lib.Lights(sc).items()
        assert len(l_ref) == len(l_my), (seed, it)
        for a, b in zip(l_ref, l_my):
            assert a[0] == b[0] and a[1] == b[1] and a[2].tobytes() == b[2].tobytes(), (seed, it)


def test_bvh_refit_matches_reference_on_random_degenerate_scenes(ref):
    """ygl_bvh_update against update_scene_bvh (yocto_bvh.cpp:303-318, 398-451): the trees of a random scene refitted to
    an edited copy (some shapes get new vertices of a new distribution - possibly NaN -, every frame moves) must equal
    the reference's refit of its own trees, node for node."""
    import copy
    for it in range(60):
        rng = np.random.default_rng(7000 + it)
        sc = abi.Scene()
        sc.add_camera()
        material = sc.add_material(color=(0.5, 0.5, 0.5))
        nshapes = int(rng.integers(1, 4))
        for _ in range(nshapes):
            _add_shape(rng, sc, [0.0, 0.001, 0.1, 1.0])
        for _ in range(int(rng.integers(1, 40))):
            sc.add_instance(int(rng.integers(0, nshapes)), material, frame=_frame(rng))
        moved = copy.copy(sc)
        moved._keep = None
        moved.shapes = [dict(s) for s in sc.shapes]
        moved.instances = [dict(n) for n in sc.instances]
        updated = [k for k in range(nshapes) if rng.random() < 0.6]
        for k in updated:
            moved.shapes[k]["positions"] = _positions(rng, len(moved.shapes[k]["positions"]))
            if len(moved.shapes[k]["radius"]):
                moved.shapes[k]["radius"] = rng.choice([0.0, 0.02, 0.5], len(moved.shapes[k]["radius"])).astype(np.float32)
        for n in moved.instances:
            n["frame"] = _frame(rng)
        for highquality in (False, True):
            mine = lib.Bvh(sc, highquality)
            mine.update(moved, updated)
            rs_old, rs_new = ref.scene(sc), ref.scene(moved)
            rs_new.adopt_updated_bvh(rs_old, updated, highquality)
            for shape in [-1] + list(range(nshapes)):
                n_ref, p_ref = rs_new.bvh_tree(shape, highquality)
                n_my, p_my = mine.tree(shape)
                assert n_ref.tobytes() == n_my.tobytes() and p_ref.tobytes() == p_my.tobytes(), (it, highquality, shape)


@pytest.mark.parametrize("mode", ["uniform", "lattice", "coincident", "plane", "signedzero", "clusters", "nan"])
def test_large_trees_built_on_several_cores_match_reference(ref, mode):
    """From 32768 primitives up one tree is built on all host cores (make_tree_parallel, ygl_build.cpp: top nodes one by
    one with the SAH candidates spread over the cores, subtrees independently, node slots handed out by replaying the
    reference's walk): shape trees of 33-70 K elements and an instance tree of 40 K instances whose centroids tie, are
    coincident, planar, carry signed zeros or NaN - nodes and primitive order bit-identical in both split modes."""
    rng = np.random.default_rng(["uniform", "lattice", "coincident", "plane", "signedzero", "clusters", "nan"].index(mode))
    nv = 20000
    if mode == "uniform":
        pos = rng.uniform(-1, 1, (nv, 3))
    elif mode == "lattice":
        pos = rng.integers(-6, 7, (nv, 3)).astype(np.float64) * 0.25
    elif mode == "coincident":
        pos = np.tile(rng.uniform(-1, 1, (1, 3)), (nv, 1))
    elif mode == "plane":
        pos = rng.uniform(-1, 1, (nv, 3))
        pos[:, 1] = 0.25
    elif mode == "signedzero":
        pos = rng.choice([0.0, -0.0, 1.0, -1.0], (nv, 3))
    elif mode == "clusters":
        pos = rng.integers(0, 2, (nv, 1)) * 10.0 + rng.uniform(-0.01, 0.01, (nv, 3))
    else:
        pos = rng.uniform(-1, 1, (nv, 3))
        pos[rng.integers(0, nv, 50), rng.integers(0, 3, 50)] = np.nan
        pos[rng.integers(0, nv, 20), rng.integers(0, 3, 20)] = np.inf
    pos = pos.astype(np.float32)
    sc = abi.Scene()
    sc.add_camera()
    material = sc.add_material(color=(0.5, 0.5, 0.5))
    sc.add_shape(triangles=rng.integers(0, nv, (int(rng.integers(33000, 70000)), 3)), positions=pos)
    sc.add_shape(lines=rng.integers(0, nv, (int(rng.integers(33000, 50000)), 2)), positions=pos,
                 radius=rng.choice([0.0, 0.001, 0.1], nv).astype(np.float32))
    sc.add_shape(triangles=[[0, 1, 2]], positions=[[0, 0, 0], [1, 0, 0], [0, 1, 0]])
    for k in range(40000 if mode in ("uniform", "lattice", "nan") else 50):
        frame = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], np.float32)
        frame[9:] = rng.integers(-8, 9, 3) if mode == "lattice" else rng.uniform(-30, 30, 3)
        if mode == "nan" and k % 997 == 0:
            frame[9 + k % 3] = np.nan
        sc.add_instance(2 if k >= 2 else k, material, frame=frame)
    rs = ref.scene(sc)
    for highquality in (False, True):
        mine = lib.Bvh(sc, highquality)
        for shape in [-1, 0, 1, 2]:
            n_ref, p_ref = rs.bvh_tree(shape, highquality)
            n_my, p_my = mine.tree(shape)
            assert n_ref.tobytes() == n_my.tobytes() and p_ref.tobytes() == p_my.tobytes(), (mode, highquality, shape)


@pytest.mark.parametrize("name", ["c2", "c5"])
def test_reference_assets_built_on_several_cores_match_reference(ref, name, tmp_path):
    """The BASELINE shapes themselves - bunny.ply (144 K triangles), hairball1.ply (262 K lines) + two bunnies - through
    the multi-core host build, both split modes (the SAH build of the C5 scene: 0.26 s against the reference's 1.07 s on
    the 8 cores of the development container)."""
    import os
    import scene_data
    from ygl_b200 import scenes
    if not scene_data.available():
        pytest.skip("oracle/_ref/data not present")
    pool = os.path.join(scene_data.DATA, "pool")
    sc = scenes.bunny_file_scene(tmp_path, pool) if name == "c2" else scenes.hairball_file_scene(tmp_path, pool)
    rs = ref.scene(sc)
    for highquality in (False, True):
        mine = lib.Bvh(sc, highquality)
        for shape in [-1] + list(range(len(sc.shapes))):
            n_ref, p_ref = rs.bvh_tree(shape, highquality)
            n_my, p_my = mine.tree(shape)
            assert n_ref.tobytes() == n_my.tobytes() and p_ref.tobytes() == p_my.tobytes(), (name, highquality, shape)
