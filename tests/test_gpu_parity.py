"""GPU parity tests (run on the B200 box): the CUDA path through the C ABI against the oracle —
the real, unmodified reference CPU renderer (oracle/_ref) and the committed golden fixtures.
Bit-exact for hit ids / uv / distance AND for every rendered pixel (image, albedo, normal, hits,
rng state): the device restates glibc's float libm bit for bit (ygl_glibm.cuh, checked against the
box's host libm in test_device_libm_matches_host_glibc). The double-rounded-libm twin of the
reference is kept as a cross-check within parity_util.assert_close_to_reference."""
import os

import numpy as np
import pytest

from parity_util import assert_close_to_reference, axis_rays, compare_hits, edited_copy, image_stats, random_rays
from ygl_b200 import abi, lib, scenes

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCENES = {
    "cornell": scenes.cornellbox,
    "cornell_quads": scenes.cornellbox_quads,
    "instanced4": lambda: scenes.instanced_spheres(4),
    "features": scenes.features,
    "hair": lambda: scenes.hair_scene(4000, 8, 3),
}
_cache = {}


def get_scene(name):
    if name not in _cache:
        _cache[name] = SCENES[name]()
    return _cache[name]


@pytest.mark.parametrize("name", list(SCENES))
def test_intersect_rays_bit_exact(ctx, ref, name):
    scene = get_scene(name)
    ds, rs = lib.DeviceScene(ctx, scene), ref.scene(scene)
    rays = np.concatenate([random_rays(scene, 100000), axis_rays(scene, 20000),
                           random_rays(scene, 20000, seed=11, tmax=1.5)])
    assert compare_hits(rs.intersect(rays), ds.intersect(rays)) == 0
    assert compare_hits(rs.intersect(rays, find_any=True), ds.intersect(rays, find_any=True)) == 0
    for inst in range(0, len(scene.instances), max(1, len(scene.instances) // 5)):
        assert compare_hits(rs.intersect(rays[:30000], instance=inst),
                            ds.intersect(rays[:30000], instance=inst)) == 0


def test_intersect_rays_edge_cases(ctx, ref):
    scene = get_scene("features")
    ds, rs = lib.DeviceScene(ctx, scene), ref.scene(scene)
    assert len(ds.intersect(np.zeros(0, abi.RAY_DTYPE))) == 0  # empty batch
    rays = random_rays(scene, 1000)
    odd = rays[:37].copy()  # ragged (not a multiple of the warp size)
    assert compare_hits(rs.intersect(odd), ds.intersect(odd)) == 0
    weird = rays[:64].copy()
    weird["d"][:16] = 0  # zero directions: infinite slabs / NaN paths
    weird["tmax"][16:32] = 0  # empty range
    weird["tmin"][32:48] = np.float32(np.inf)
    weird["d"][48:64] *= np.float32(1e-30)  # denormal-ish directions
    assert compare_hits(rs.intersect(weird), ds.intersect(weird)) == 0


def test_golden_ray_fixture(ctx):
    """Committed rays + reference intersections (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLDEN, "rays_features.npz"))
    ds = lib.DeviceScene(ctx, get_scene("features"))
    out = ds.intersect(g["rays"].view(abi.RAY_DTYPE).reshape(-1))
    assert compare_hits(g["hits"].view(abi.ISEC_DTYPE).reshape(-1), out) == 0


RENDERS = [
    ("cornell", dict(resolution=64, samples=4, bounces=4)),
    ("cornell", dict(resolution=96, samples=8, bounces=8, tentfilter=1)),
    ("cornell_quads", dict(resolution=64, samples=4, bounces=4)),
    ("instanced4", dict(resolution=96, samples=4, bounces=8)),
    ("instanced4", dict(resolution=64, samples=3, bounces=8, highqualitybvh=1)),
    ("features", dict(resolution=128, samples=4, bounces=8)),
    ("features", dict(resolution=96, samples=3, bounces=8, camera=1)),
    ("features", dict(resolution=96, samples=3, bounces=6, nocaustics=1, envhidden=1)),
    ("features", dict(resolution=96, samples=2, bounces=8, sampler=abi.SAMPLER_EYELIGHT)),
    ("features", dict(resolution=96, samples=1, sampler=abi.SAMPLER_FALSECOLOR, falsecolor=abi.FC_NORMAL)),
    ("features", dict(resolution=96, samples=1, sampler=abi.SAMPLER_FALSECOLOR, falsecolor=abi.FC_ELEMENT)),
    ("features", dict(resolution=96, samples=1, sampler=abi.SAMPLER_FALSECOLOR, falsecolor=abi.FC_INSTANCE)),
    ("features", dict(resolution=96, samples=1, sampler=abi.SAMPLER_FALSECOLOR, falsecolor=abi.FC_TEXCOORD)),
    ("hair", dict(resolution=96, samples=4, bounces=8)),
    ("features", dict(resolution=96, samples=3, bounces=8, sampler=abi.SAMPLER_NAIVE)),
    ("cornell", dict(resolution=64, samples=4, bounces=6, sampler=abi.SAMPLER_NAIVE)),
    ("features", dict(resolution=96, samples=3, bounces=8, sampler=abi.SAMPLER_FURNACE)),
    ("features", dict(resolution=96, samples=2, bounces=8, sampler=abi.SAMPLER_FURNACE, camera=1, envhidden=1)),
    ("features", dict(resolution=96, samples=3, bounces=8, sampler=abi.SAMPLER_PATHDIRECT)),
    ("features", dict(resolution=96, samples=2, bounces=6, sampler=abi.SAMPLER_PATHDIRECT, camera=1, nocaustics=1, envhidden=1)),
    ("cornell", dict(resolution=64, samples=4, bounces=8, sampler=abi.SAMPLER_PATHDIRECT)),
    ("hair", dict(resolution=64, samples=2, bounces=8, sampler=abi.SAMPLER_PATHDIRECT)),
    ("features", dict(resolution=96, samples=3, bounces=8, sampler=abi.SAMPLER_PATHMIS)),
    ("features", dict(resolution=96, samples=2, bounces=6, sampler=abi.SAMPLER_PATHMIS, camera=1, nocaustics=1, envhidden=1)),
    ("cornell", dict(resolution=64, samples=4, bounces=8, sampler=abi.SAMPLER_PATHMIS)),
    ("instanced4", dict(resolution=64, samples=2, bounces=8, sampler=abi.SAMPLER_PATHMIS)),
    ("features", dict(resolution=96, samples=3, bounces=8, sampler=abi.SAMPLER_PATHTEST)),
    ("cornell", dict(resolution=64, samples=3, bounces=5, sampler=abi.SAMPLER_PATHTEST, envhidden=1)),
    ("features", dict(resolution=96, samples=3, bounces=8, sampler=abi.SAMPLER_DIAGRAM)),
    ("features", dict(resolution=96, samples=2, bounces=2, sampler=abi.SAMPLER_DIAGRAM, camera=1)),
    ("cornell", dict(resolution=64, samples=2, bounces=4, sampler=abi.SAMPLER_DIAGRAM)),
]


@pytest.mark.parametrize("name,kw", RENDERS)
def test_render_matches_oracle(ctx, ref, ref_dlibm, name, kw):
    scene = get_scene(name)
    params = abi.trace_params(**kw)
    image = ctx.trace_image(scene, params)
    exact = ref.scene(scene).trace_image(params)["image"]
    # the UNMODIFIED reference: same arithmetic order, same libm bits (ygl_glibm.cuh) -> every pixel identical
    assert image.tobytes() == exact.tobytes(), image_stats(exact, image)
    # cross-check: the double-rounded-libm twin of the reference stays within the stated float tolerance
    assert_close_to_reference(image_stats(ref_dlibm.scene(scene).trace_image(params)["image"], image))


@pytest.mark.parametrize("sampler", [abi.SAMPLER_PATHDIRECT, abi.SAMPLER_PATHMIS, abi.SAMPLER_PATHTEST, abi.SAMPLER_NAIVE,
                                     abi.SAMPLER_EYELIGHT, abi.SAMPLER_DIAGRAM, abi.SAMPLER_FURNACE, abi.SAMPLER_FALSECOLOR])
def test_full_state_every_sampler(ctx, ref, sampler):
    """Every sampler of get_trace_sampler_func (yocto_trace.cpp:1422-1438): image, denoise guides (albedo,
    normal), hit counts and the advanced rng streams all match the unmodified reference bit for bit."""
    scene = get_scene("features")
    params = abi.trace_params(resolution=72, samples=2, bounces=6, batch=2, sampler=sampler)
    want = ref.scene(scene).trace_image(params, full=True)
    ds = lib.DeviceScene(ctx, scene)
    st = ds.make_state(params)
    ds.trace_samples(st, params)
    got = st.download(full=True)
    for k in ("image", "albedo", "normal", "hits"):
        assert got[k].tobytes() == want[k].tobytes(), k
    np.testing.assert_array_equal(got["rngs"], want["rngs"])


@pytest.mark.parametrize("options", [
    dict(suspend=8, suspend_rounds=8, lone=0),      # park the last lanes of a drained warp early and often
    dict(suspend=0, lone=32, lone_steps=40),        # vote-free tail walk, parked after 40 steps
    dict(suspend=0, lone=0, fuse=0),                # plain drain, finished paths go through k_finish
    dict(fuse=1, refill=4, node_reps=1),
    dict(bin=0, fuse=0),                            # one shade queue, the unspecialised kernel
    dict(bin=0, fuse=1),
    dict(bin=1, fuse=0, suspend=8, suspend_rounds=4),  # class-binned queues + parked rays re-entering the extend queue
    dict(bin=1, fuse=1, pipes=2),
    dict(ext_blocks_per_sm=2, bin=1),
    dict(graph=1, fuse=1, bin=1),                   # rounds of iterations submitted as one CUDA graph
    dict(graph=1, fuse=0, bin=0, lone=8),
    dict(graph=0, top_smem=1, suspend=8, lone=0),   # instance-level tree staged in shared memory by a bulk async copy
])
def test_scheduling_options_are_bit_exact(ref, options):
    """Every scheduling knob of a context (ygl_context_set_option: the extend kernel's tail strategies, path-end
    fusion, class-binned shade queues and per-class kernels, pipelines, grid size) is a scheduling change only:
    the image of two resumed batches equals the reference bit for bit whatever the setting."""
    octx = lib.Context(0)
    for k, v in options.items():
        octx.set_option(k, v)
        assert octx.get_option(k) == v
    with pytest.raises(lib.YglError):
        octx.set_option("no_such_option", 1)
    scene = get_scene("features")
    params = abi.trace_params(resolution=160, samples=4, bounces=8, batch=2)
    ds = lib.DeviceScene(octx, scene)
    st = ds.make_state(params)
    ds.trace_samples(st, params)
    ds.trace_samples(st, params)
    got = st.download(full=True)
    want = ref.scene(scene).trace_image(params, full=True)
    for k in ("image", "albedo", "normal", "hits"):
        assert got[k].tobytes() == want[k].tobytes(), (k, options)
    np.testing.assert_array_equal(got["rngs"], want["rngs"])


def test_state_reset_trace_sample_and_adopted_bvh(ctx, ref):
    """The rest of the low-level API (yocto_trace.h:160-190, yocto_cutrace.h:119): ygl_state_reset re-seeds and
    renders the same bits again; per-pixel ygl_trace_sample calls in the reference's loop order reproduce
    trace_samples; a bvh adopted verbatim from the reference's make_scene_bvh (ygl_bvh_create_from_host) renders
    the same image, and a corrupted tree is refused."""
    scene = get_scene("cornell")
    params = abi.trace_params(resolution=24, samples=2, bounces=4, batch=2)
    rs = ref.scene(scene)
    want = rs.trace_image(params, full=True)
    ds = lib.DeviceScene(ctx, scene)
    st = ds.make_state(params)
    ds.trace_samples(st, params)
    assert st.download()["image"].tobytes() == want["image"].tobytes()
    st.reset(params)
    assert st.samples == 0 and not st.download()["image"].any()
    ds.trace_samples(st, params)
    assert st.download()["image"].tobytes() == want["image"].tobytes()
    # trace_sample, pixel by pixel (yocto_trace.cpp:1600-1606: the noparallel loop of trace_samples)
    st.reset(params)
    for j in range(0, st.height, 5):
        for i in range(st.width):
            for sample in range(2):
                ds.trace_sample(st, i, j, sample, params)
    assert st.samples == 0
    got = st.download(full=True)
    assert got["image"][::5].tobytes() == want["image"][::5].tobytes()
    assert got["hits"][::5].tobytes() == want["hits"][::5].tobytes()
    assert not got["image"][1::5].any()
    with pytest.raises(lib.YglError):
        ds.trace_sample(st, st.width, 0, 0, params)
    # the reference's own trees, adopted verbatim
    trees = (rs.bvh_tree(-1), [rs.bvh_tree(k) for k in range(len(scene.shapes))])
    ds2 = lib.DeviceScene(ctx, scene, trees=trees)
    st2 = ds2.make_state(params)
    ds2.trace_samples(st2, params)
    assert st2.download()["image"].tobytes() == want["image"].tobytes()
    rays = random_rays(scene, 20000)
    assert compare_hits(rs.intersect(rays), ds2.intersect(rays)) == 0
    bad_nodes, bad_prims = trees[1][0][0].copy(), trees[1][0][1].copy()
    bad_prims[0] = 10 ** 6
    with pytest.raises(lib.YglError):
        lib.Bvh(scene, trees=(trees[0], [(bad_nodes, bad_prims)] + trees[1][1:]))


def test_full_state_and_resume(ctx, ref):
    ref_dlibm = ref  # (name kept below) the unmodified reference
    """trace_state semantics: albedo/normal/hits/rngs match, and batches of 1,2,3 samples resume to
    the same bits as one batch of 6 (yocto_trace.cpp:1595-1619)."""
    scene = get_scene("features")
    params = abi.trace_params(resolution=80, samples=6, bounces=6, batch=6)
    out_ref = ref_dlibm.scene(scene).trace_image(params, full=True)
    ds = lib.DeviceScene(ctx, scene)
    st = ds.make_state(params)
    ds.trace_samples(st, params)
    a = st.download(full=True)
    for k in ("image", "albedo", "normal", "hits"):
        assert a[k].tobytes() == out_ref[k].tobytes(), k
    np.testing.assert_array_equal(a["rngs"], out_ref["rngs"])
    st2 = ds.make_state(params)
    for batch in (1, 2, 3):
        p = abi.trace_params(resolution=80, samples=6, bounces=6, batch=batch)
        ds.trace_samples(st2, p)
    assert st2.samples == 6
    b = st2.download(full=True)
    for k in a:
        assert a[k].tobytes() == b[k].tobytes(), k
    ds.trace_samples(st2, params)  # no-op once samples >= params.samples
    assert st2.samples == 6
    # checkpoint / restore through host memory
    st3 = ds.make_state(params)
    p3 = abi.trace_params(resolution=80, samples=6, bounces=6, batch=3)
    ds.trace_samples(st3, p3)
    half = st3.download(full=True)
    st4 = ds.make_state(params)
    st4.upload(3, half["image"], half["albedo"], half["normal"], half["hits"], half["rngs"])
    ds.trace_samples(st4, p3)
    assert st4.download()["image"].tobytes() == a["image"].tobytes()


def test_tiles_equal_full_image(ctx):
    """Row-tile partition (the multi-GPU decomposition) reproduces the single-state image bitwise."""
    scene = get_scene("instanced4")
    params = abi.trace_params(resolution=72, samples=3, bounces=6, batch=3)
    ds = lib.DeviceScene(ctx, scene)
    full = ds.make_state(params)
    ds.trace_samples(full, params)
    want = full.download()["image"]
    got = np.zeros_like(want)
    for r in range(3):
        rows = lib.tile_rows(full.height, r, 3)
        st = ds.make_state(params, rows=rows)
        ds.trace_samples(st, params)
        got[rows[0]:rows[1]] = st.download()["image"]
    assert want.tobytes() == got.tobytes()
    assert full.gather_image().tobytes() == want.tobytes()  # single rank: gather == download
    # interleaved tiling (row j -> rank j % n), the layout bench.py uses for N > 1
    got2 = np.zeros_like(want)
    for r in range(4):
        st = ds.make_state(params, interleave=(r, 4))
        assert (st.row_first, st.row_step) == (r, 4)
        ds.trace_samples(st, params)
        got2[r::4] = st.download()["image"]
    assert want.tobytes() == got2.tobytes()


def test_golden_images(ctx):
    """Committed golden renders produced by the real reference (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLDEN, "renders.npz"))
    for key in [k for k in g.files if k.endswith(".image")]:
        name = key.split(".")[0]
        kw = {k: int(v) for k, v in zip(g[f"{name}.param_names"], g[f"{name}.param_values"])}
        scene = get_scene(str(g[f"{name}.scene"]))
        image = ctx.trace_image(scene, abi.trace_params(**kw))
        assert image.tobytes() == g[key].tobytes(), (name, image_stats(g[key], image))
        assert_close_to_reference(image_stats(g[f"{name}.image_dlibm"], image))


# ---- parity AT THE BASELINE SIZES (BASELINE.json configs[1..4]) against the unmodified reference ----
FULL_SIZE = {
    # name: (scene factory, trace_params of the config at a sample count the CPU reference renders in seconds)
    "c3": (lambda: scenes.instanced_spheres(10), dict(resolution=1920, samples=2, bounces=8, batch=2)),
    "c2": (lambda: scenes.bunny_like(6), dict(resolution=1280, samples=2, bounces=8, batch=2)),
    "c5": (scenes.hair_stress, dict(resolution=1920, samples=1, bounces=12, batch=1)),
}
_full_cache = {}


def full_scene(name):
    if name not in _full_cache:
        _full_cache[name] = FULL_SIZE[name][0]()
    return _full_cache[name]


@pytest.mark.parametrize("name", list(FULL_SIZE))
def test_full_size_render_bit_exact(ctx, ref, name):
    """C3 at 1920x1080, C2 at 1280x720, C5 at 1920x1080: image, denoise guides, hit counts and rng streams of the
    whole frame identical to the reference CPU render, bit for bit (tolerance 0 => per-pixel RMSE 0 < 1e-5)."""
    scene = full_scene(name)
    params = abi.trace_params(**FULL_SIZE[name][1])
    want = ref.scene(scene).trace_image(params, full=True)
    ds = lib.DeviceScene(ctx, scene)
    st = ds.make_state(params)
    ds.trace_samples(st, params)
    got = st.download(full=True)
    assert got["image"].shape == want["image"].shape
    for k in ("image", "albedo", "normal", "hits"):
        assert got[k].tobytes() == want[k].tobytes(), (name, k, image_stats(want["image"], got["image"]))
    np.testing.assert_array_equal(got["rngs"], want["rngs"])
    # ... and to the committed digest of the reference render made in the development container
    import hashlib
    import json
    fixture = json.load(open(os.path.join(GOLDEN, "traversal_counters.json")))[name]
    assert hashlib.sha256(got["image"].tobytes()).hexdigest() == fixture["image_sha256"]
    assert hashlib.sha256(got["rngs"].tobytes()).hexdigest() == fixture["rngs_sha256"]
    # size-independent properties of the same frame
    assert np.isfinite(got["image"]).all() and got["image"][..., :3].max() <= 10.0 + 1e-4  # clamp respected
    spp = params.samples
    assert ((got["hits"] == spp) == (got["image"][..., 3] == 1.0)).all()
    assert ctx.counters()["camera_samples"] == got["hits"].size * spp
    # tile invariance at size: a band of rows rendered alone equals the same rows of the frame
    H = got["image"].shape[0]
    rows = (H // 2 - 20, H // 2 + 20)
    st2 = ds.make_state(params, rows=rows)
    ds.trace_samples(st2, params)
    assert st2.download()["image"].tobytes() == got["image"][rows[0]:rows[1]].tobytes()


@pytest.mark.parametrize("name", list(FULL_SIZE))
def test_full_size_rays_bit_exact(ctx, ref, name):
    """>= 10^6 rays per config (all-direction rays from the scene box + primary rays of the config's camera)
    through ygl_intersect_rays: instance, element, uv and distance identical to intersect_scene_bvh."""
    scene = full_scene(name)
    params = abi.trace_params(**FULL_SIZE[name][1])
    ds, rs = lib.DeviceScene(ctx, scene), ref.scene(scene)
    rays = np.concatenate([random_rays(scene, 600000), scenes.camera_rays(scene, params, 500000),
                           axis_rays(scene, 50000)])
    threads = max(8, os.cpu_count() or 8)
    assert compare_hits(rs.intersect(rays, nthreads=threads), ds.intersect(rays)) == 0
    some = rays[:200000]
    assert compare_hits(rs.intersect(some, find_any=True, nthreads=threads), ds.intersect(some, find_any=True)) == 0


# ---- the reference's own scene files through the library's own loader ----
import scene_data  # noqa: E402


@pytest.mark.skipif(not scene_data.available(), reason="oracle/_ref/data not present")
@pytest.mark.parametrize("name", [n for n in scene_data.names() if n != "cornellbox"] or ["none"])
def test_reference_scene_files_render_bit_exact(ctx, ref, name, tmp_path):
    """The reference's own test scenes (tests/_version43, incl. the two with subdivs): loaded with ygl_scene_load, rendered
    on the GPU, compared bit for bit with the reference rendering the scene IT loaded from the same file — every
    camera of the file for one of them, the default camera for the rest."""
    path = scene_data.scene_file(name, tmp_path)
    ours, theirs = lib.load_scene(path), ref.load_scene(path)
    rs, ds = ref.scene(theirs), lib.DeviceScene(ctx, ours)
    cameras = range(len(ours.cameras)) if name == "features1" else [0]
    for camera in cameras:
        params = abi.trace_params(resolution=360, samples=2, bounces=8, batch=2, camera=camera)
        want = rs.trace_image(params, full=True)
        st = ds.make_state(params)
        ds.trace_samples(st, params)
        got = st.download(full=True)
        for k in ("image", "albedo", "normal", "hits"):
            assert got[k].tobytes() == want[k].tobytes(), (name, camera, k, image_stats(want["image"], got["image"]))
    rays = np.concatenate([random_rays(ours, 200000), scenes.camera_rays(ours, abi.trace_params(resolution=360), 100000)])
    assert compare_hits(rs.intersect(rays), ds.intersect(rays)) == 0


@pytest.mark.skipif(not scene_data.available(), reason="oracle/_ref/data not present")
@pytest.mark.parametrize("which", ["c2_bunny", "c5_hairball"])
def test_baseline_assets_full_size_bit_exact(ctx, ref, which, tmp_path):
    """C2 and C5 with the assets BASELINE.json names (the reference's bunny.ply: 144,046 triangles; hairball1.ply:
    2 x 262,144 line segments + 2 bunnies) at the configs' resolution: whole frame bit-identical to the reference."""
    pool = os.path.join(scene_data.DATA, "pool")
    if which == "c2_bunny":
        scene, kw = scenes.bunny_file_scene(tmp_path, pool), dict(resolution=1280, samples=2, bounces=8, batch=2)
    else:
        scene, kw = scenes.hairball_file_scene(tmp_path, pool), dict(resolution=1920, samples=1, bounces=12, batch=1)
    params = abi.trace_params(**kw)
    want = ref.scene(scene).trace_image(params, full=True)
    ds = lib.DeviceScene(ctx, scene)
    st = ds.make_state(params)
    ds.trace_samples(st, params)
    got = st.download(full=True)
    assert got["image"].shape[1] == kw["resolution"]
    for k in ("image", "albedo", "normal", "hits"):
        assert got[k].tobytes() == want[k].tobytes(), (which, k, image_stats(want["image"], got["image"]))
    np.testing.assert_array_equal(got["rngs"], want["rngs"])


def _stress_soup(n=60000, seed=21):
    """Triangles built to hit the corners of make_bvh: thousands of coincident centroids (csize == 0 fallback), long
    runs of equal coordinates (predicate splits nothing -> midpoint fallback), signed zeros, a few huge triangles."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-1, 1, size=(n, 3))
    c[: n // 6] = c[0]                                # identical centroids
    c[n // 6: n // 3, 0] = 0.25                       # a plane of equal x
    c[n // 3: n // 2] = np.round(c[n // 3: n // 2] * 4) / 4  # a coarse lattice: many ties
    c[n // 2: n // 2 + 500] *= 0.0                    # zeros ...
    c[n // 2 + 250: n // 2 + 500] *= -1.0             # ... of both signs
    size = rng.uniform(0.001, 0.02, size=(n, 1, 1))
    size[:50] = 3.0
    p = c[:, None, :] + rng.normal(size=(n, 3, 3)) * size
    p[n // 2: n // 2 + 500] = np.where(rng.random((500, 3, 3)) < 0.5, 0.0, -0.0)  # degenerate, signed-zero boxes
    sc = scenes.Scene()
    sc.add_camera(scenes.lookat_frame((0, 0, 4), (0, 0, 0)), lens=0.05, film=0.036, aspect=1.0, focus=4.0, aperture=0.0)
    m = sc.add_material(abi.MATTE, color=(0.7, 0.7, 0.7))
    sc.add_instance(sc.add_shape(triangles=np.arange(3 * n, dtype=np.int32).reshape(-1, 3),
                                 positions=p.reshape(-1, 3).astype(np.float32)), m)
    return sc


@pytest.mark.parametrize("name", ["c2", "c5", "soup", "instances", "chain"])
def test_device_bvh_build_matches_host_build(ctx, name, tmp_path):
    """ygl_bvh_build_device (level-parallel split_middle on the GPU, ygl_bvh_device.cu) against the host build - which
    tests/test_host_parity.py pins to the reference's make_scene_bvh: every node (box bits, start, num, axis, internal)
    at the same index and the same `primitives` permutation, for every tree of the scene."""
    import time
    if name == "soup":
        scene = _stress_soup()
    elif name == "instances":
        scene = scenes.instanced_spheres(17)   # 4913 + 2 instances: the instance tree is built on the device too
    elif name == "chain":
        scene = scenes.sliver_chain(100)       # small: stays on the host (and must still be accepted)
    elif scene_data.available() and name == "c2":
        scene = scenes.bunny_file_scene(tmp_path, os.path.join(scene_data.DATA, "pool"))
    elif scene_data.available() and name == "c5":
        scene = scenes.hairball_file_scene(tmp_path, os.path.join(scene_data.DATA, "pool"))
    else:
        scene = full_scene(name)
    t0 = time.time()
    host = lib.Bvh(scene)
    t1 = time.time()
    dev = lib.Bvh(scene, device_ctx=ctx)
    t2 = time.time()
    dev2 = lib.Bvh(scene, device_ctx=ctx)
    t3 = time.time()
    print(f"bvh build {name}: host {1e3 * (t1 - t0):.1f} ms, device {1e3 * (t2 - t1):.1f} ms (first), {1e3 * (t3 - t2):.1f} ms")
    for shape in range(-1, len(scene.shapes)):
        hn, hp = host.tree(shape)
        dn, dp = dev.tree(shape)
        assert hp.tobytes() == dp.tobytes(), (name, shape, "primitives")
        assert hn.tobytes() == dn.tobytes(), (name, shape, "nodes")
    # and the trees render: rays through the device-built bvh equal rays through the host-built one
    ds = lib.DeviceScene(ctx, scene, device_build=True)
    rays = random_rays(scene, 100000)
    assert compare_hits(lib.DeviceScene(ctx, scene).intersect(rays), ds.intersect(rays)) == 0


COUNTER_KEYS = ("top_nodes", "bottom_nodes", "instance_visits", "triangle_tests", "quad_tests", "line_tests", "point_tests")


@pytest.mark.parametrize("name", ["c3", "c5"])
def test_traversal_counters_match_instrumented_oracle(ctx, ref_count, name):
    """The numerator of bench.py's roofline: the counting variant of k_extend must reproduce, exactly, the
    per-query counters of the INSTRUMENTED reference (oracle/ref_counters.h: counter sites at
    yocto_bvh.cpp:466,487,506-545,560,581) on the BASELINE config, and the committed fixture of the same run."""
    scene = full_scene(name)
    params = abi.trace_params(**FULL_SIZE[name][1])
    rs = ref_count.scene(scene)
    ref_count.counters_reset()
    rs.trace_image(params)
    want = ref_count.counters()["scene"]
    ctx.set_profiling(False, True)
    try:
        ds = lib.DeviceScene(ctx, scene)
        st = ds.make_state(params)
        ds.trace_samples(st, params)
        got = ctx.counters()
    finally:
        ctx.set_profiling(False, False)
    assert got["scene_rays"] == want["rays"]
    for k in COUNTER_KEYS:
        assert got[k] == want[k], (k, got[k], want[k])
    assert got["instance_rays"] == ref_count.counters()["instance"]["rays"]
    import json
    fixture = json.load(open(os.path.join(GOLDEN, "traversal_counters.json")))[name]
    assert {k: got[k] for k in COUNTER_KEYS} == {k: fixture["scene"][k] for k in COUNTER_KEYS}
    assert got["scene_rays"] == fixture["scene"]["rays"]


def test_deep_tree_uses_the_full_reference_stack(ctx, ref):
    """A degenerate (sorted-sliver) mesh whose split_middle tree is a 97-level chain: deeper than the 64 entries
    the round-1 build accepted, inside the reference's 128 (yocto_bvh.cpp:469). Rays and a render match bit for
    bit; a 167-level chain, which would overflow the reference's own stack, is rejected with an error."""
    scene = scenes.sliver_chain(100)
    ds, rs = lib.DeviceScene(ctx, scene), ref.scene(scene)
    nodes, _ = ds.bvh.tree(0)
    depth, todo = 0, [(0, 1)]
    while todo:
        i, d = todo.pop()
        depth = max(depth, d)
        if nodes[i]["internal"]:
            todo += [(int(nodes[i]["start"]), d + 1), (int(nodes[i]["start"]) + 1, d + 1)]
    assert 64 < depth <= 128
    rng = np.random.default_rng(9)
    rays = random_rays(scene, 50000)
    # aim a third of the rays straight at sliver centres, from +x and from -x (both child orders of the chain)
    k = rng.integers(0, 100, 30000)
    tgt = scene.shapes[0]["positions"].reshape(-1, 3, 3).mean(1)[k]
    org = tgt + np.where(rng.random((30000, 1)) < 0.5, 1.0, -1.0) * np.array([[2.0, 0.3, 0.2]], np.float32)
    d = tgt - org
    rays["o"][:30000] = org.astype(np.float32)
    rays["d"][:30000] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    assert compare_hits(rs.intersect(rays), ds.intersect(rays)) == 0
    params = abi.trace_params(resolution=128, samples=4, bounces=4)
    assert ctx.trace_image(scene, params).tobytes() == rs.trace_image(params)["image"].tobytes()
    with pytest.raises(lib.YglError):
        lib.DeviceScene(ctx, scenes.sliver_chain(170))


def test_progressive_api(ref):
    """trace_start / trace_done / trace_cancel / trace_preview (yocto_trace.cpp:1627-1676): batches started
    asynchronously and waited for give the reference's bits; a cancelled batch returns promptly and leaves a state
    that ygl_state_reset makes good again; the preview is the 1-spp render at resolution / pratio, replicated."""
    import time
    pctx = lib.Context(0)
    scene = get_scene("features")
    params = abi.trace_params(resolution=192, samples=4, bounces=6, batch=2)
    rs = ref.scene(scene)
    want = rs.trace_image(params, full=True)
    ds = lib.DeviceScene(pctx, scene)
    st = ds.make_state(params)
    while st.samples < params.samples:
        ds.trace_start(st, params)
        t0 = time.time()
        while not pctx.trace_done():
            assert time.time() - t0 < 60
            time.sleep(0.001)
        pctx.trace_wait()
    assert st.download()["image"].tobytes() == want["image"].tobytes()
    # cancel a long batch right after starting it
    big = abi.trace_params(resolution=192, samples=4096, bounces=6, batch=4096)
    st2 = ds.make_state(big)
    ds.trace_start(st2, big)
    time.sleep(0.05)
    t0 = time.time()
    pctx.trace_cancel()
    assert time.time() - t0 < 20 and not pctx.trace_done()
    assert st2.samples == 4096  # yocto_trace.cpp:1641: the batch is counted even when abandoned
    st2.reset(params)
    ds.trace_samples(st2, params)
    ds.trace_samples(st2, params)
    assert st2.download()["image"].tobytes() == want["image"].tobytes()
    # preview
    pv = abi.trace_params(resolution=192, samples=4, bounces=6, batch=2, pratio=4)
    W, H = st.width, st.height
    got = ds.trace_preview(pv, W, H)
    small = rs.trace_image(abi.trace_params(resolution=192 // 4, samples=1, bounces=6, batch=1))["image"]
    jj = np.clip(np.arange(H) // 4, 0, small.shape[0] - 1)
    ii = np.clip(np.arange(W) // 4, 0, small.shape[1] - 1)
    assert got.tobytes() == small[jj][:, ii].tobytes()


def test_unsupported_sampler_and_errors(ctx):
    scene = get_scene("cornell")
    with pytest.raises(lib.YglError):
        ctx.trace_image(scene, abi.trace_params(resolution=16, samples=1, sampler=99))
    with pytest.raises(lib.YglError):
        ctx.trace_image(scene, abi.trace_params(resolution=16, samples=1, camera=3))
    with pytest.raises(lib.YglError):
        lib.Context(99)


def test_reference_side_shim_runs():
    """The C++ drop-in (yocto::b200::trace_image through yocto_b200trace.h) next to yocto::trace_image
    in one reference-side program (oracle/shim_demo.cpp, prebuilt where the reference headers exist)."""
    import subprocess
    exe = os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "shim_demo")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/shim_demo not built")
    out = subprocess.run([exe, "96", "8"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr


def test_device_libm_matches_host_glibc(ctx):
    """The device libm must reproduce THIS box's glibc float routines bit for bit (the reference calls
    them at run time): 4M inputs per function over the ranges the path tracer produces, plus edges."""
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    rng = np.random.default_rng(5)
    n = 4_000_000

    def host(name, *args):
        f = getattr(libm, name)
        f.restype = ctypes.c_float
        f.argtypes = [ctypes.c_float] * len(args)
        return np.array([f(*[float(v) for v in row]) for row in zip(*args)], np.float32)

    def check(fn, name, x, y=None):
        dev = ctx.libm(fn, x, y)
        idx = rng.choice(len(x), 60000, replace=False)  # ctypes calls are slow: verify a random subset ...
        want = host(name, x[idx]) if y is None else host(name, x[idx], y[idx])
        got = dev[idx]
        same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
        assert same.all(), ("device libm differs from this host's glibc (the tables restate glibc 2.39's x86-64 FMA builds; "
                            "on another glibc or a CPU without FMA the bit-exact render tests cannot hold either)",
                            name, x[idx][~same][:5], got[~same][:5], want[~same][:5])

    u = rng.random(n, dtype=np.float32)
    edges = np.array([0.0, -0.0, 1.0, -1.0, 0.5, 1e-8, 1e-30, 3.1415927, 6.2831855, 100.0, 119.9, 120.5, 1e10], np.float32)
    pad = lambda a: np.concatenate([a.astype(np.float32), edges])
    check(0, "sinf", pad(u * 6.2831855))
    check(1, "cosf", pad(u * 6.2831855))
    check(0, "sinf", pad((u - 0.5) * 200))
    check(2, "expf", pad(-u * 60))
    check(3, "logf", pad(u))
    check(4, "atanf", pad(u * 50))
    check(5, "acosf", np.clip(pad(u * 2 - 1), -1, 1))
    v = rng.random(n + len(edges), dtype=np.float32) * 2 - 1
    check(6, "atan2f", pad(u * 2 - 1), v)
    for e in (5.0, 6.0, 2.2, 2.4, 0.75, 2.0):
        check(7, "powf", pad(u), np.full(n + len(edges), e, np.float32))
    # negative bases (the refraction lobes square a signed quantity), tiny bases (underflow), deep expf tails
    check(7, "powf", pad(u * 4 - 2), np.full(n + len(edges), 2.0, np.float32))
    check(7, "powf", pad(u * 1e-7), np.full(n + len(edges), 5.0, np.float32))
    check(2, "expf", pad(-80 - u * 30))
    check(3, "logf", pad(u * 1e-38))


@pytest.mark.parametrize("name", ["instanced4", "features", "hair"])
def test_refitted_bvh(ctx, ref, name):
    """ygl_bvh_update (update_scene_bvh, yocto_bvh.cpp:434-451) after the trees were already uploaded and bound: the
    device copy is replaced, hits through the refitted trees equal the reference's through ITS refitted trees, and the
    frame rendered with them equals the reference's frame of the edited scene."""
    scene = get_scene(name)
    moved, updated = edited_copy(scene)
    ds = lib.DeviceScene(ctx, scene)
    rays = random_rays(moved, 60000)
    ds.intersect(rays)  # uploads and binds the trees of the scene before the edit
    ds.bvh.update(moved, updated)
    ds2 = lib.DeviceScene(ctx, moved)
    ds2.bvh = ds.bvh
    rs_old, rs_new = ref.scene(scene), ref.scene(moved)
    rs_new.adopt_updated_bvh(rs_old, updated)
    assert compare_hits(rs_new.intersect(rays), ds2.intersect(rays)) == 0
    assert compare_hits(rs_new.intersect(rays[:20000], find_any=True), ds2.intersect(rays[:20000], find_any=True)) == 0
    params = abi.trace_params(resolution=96, samples=4, bounces=4, batch=4)
    want = rs_new.trace_image(params)
    st = ds2.make_state(params)
    ds2.trace_samples(st, params)
    assert st.download()["image"].tobytes() == want["image"].tobytes()


def test_tonemap_matches_reference(ctx, ref):
    """ygl_tonemap_image / ygl_state_tonemap against tonemap_image (yocto_image.cpp:911-922): vec4f and vec4b outputs bit
    for bit, on a rendered frame and on a sweep that covers the sRGB knee, denormals, huge values, negatives, inf, NaN."""
    rng = np.random.default_rng(11)
    special = np.array([0.0, -0.0, 1e-45, 1e-8, 0.0031307, 0.0031308, 0.0031309, 0.18, 0.5, 1.0, 1.0000001, 3.9, 255.0 / 256,
                        1e4, 3e38, np.inf, -1e-3, -2.5, -np.inf, np.nan], np.float32)
    sweep = np.concatenate([special, np.exp(rng.uniform(-30, 12, 60000)).astype(np.float32),
                            rng.uniform(0, 1.2, 60000).astype(np.float32)])
    sweep = np.resize(sweep, (len(sweep) // 4 + 1) * 4).reshape(-1, 4)
    scene = get_scene("features")
    params = abi.trace_params(resolution=96, samples=4, bounces=4, batch=4)
    ds = lib.DeviceScene(ctx, scene)
    st = ds.make_state(params)
    ds.trace_samples(st, params)
    frame = st.download()["image"]
    for exposure in (0.0, 1.5, -2.25):
        for filmic in (False, True):
            for srgb in (False, True):
                for hdr in (sweep, frame):
                    want_f, want_b = ref.tonemap_image(hdr, exposure, filmic, srgb)
                    got_f, got_b = ctx.tonemap_image(hdr, exposure, filmic, srgb)
                    key = (exposure, filmic, srgb, hdr.shape)
                    assert got_b.tobytes() == want_b.tobytes(), key
                    ok = (got_f.view(np.uint32) == want_f.view(np.uint32)) | (np.isnan(got_f) & np.isnan(want_f))
                    assert ok.all(), (key, hdr.reshape(-1)[~ok.reshape(-1)][:8])
                want_f, want_b = ref.tonemap_image(frame, exposure, filmic, srgb)
                got_f, got_b = st.tonemap(exposure, filmic, srgb)
                assert got_f.tobytes() == want_f.tobytes() and got_b.tobytes() == want_b.tobytes(), (exposure, filmic, srgb)
    with pytest.raises(lib.YglError):
        lib._check(ctx.lib.ygl_tonemap_image(ctx.h, None, 4, 0.0, 0, 1, None, None))
