"""Synthetic scene_data generators for the parity tests and bench.py (BASELINE.json configs).

All geometry is produced deterministically in numpy (float32) from fixed seeds; the SAME arrays
feed the CUDA path and the test oracle. Config names follow SURVEY.md §8d:
  C1 cornellbox()            — the reference's make_cornellbox() values (tests/golden fixture)
  C2 bunny_like()            — ~82K-triangle blob + floor + sky-like environment texture
  C3 instanced_spheres()     — 1000 x 1024-triangle sphere instances (1,024,002 instanced tris)
  C5 hair_scene()            — line segments + triangles, glossy + subsurface materials
plus features() — a small scene touching every primitive/material/texture code path.
"""
import os

import numpy as np

from . import abi
from .abi import Scene

_GOLDEN = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..",
                                        "tests", "golden"))


def _normalize(v):
    v = np.asarray(v, np.float64)
    return v / np.linalg.norm(v)


def lookat_frame(eye, center, up=(0, 1, 0)):
    w = _normalize(np.subtract(eye, center))
    u = _normalize(np.cross(up, w))
    v = _normalize(np.cross(w, u))
    return np.array([u, v, w, eye], np.float32)


def translation(t):
    f = abi.IDENTITY_FRAME.copy()
    f[3] = t
    return f


def rotation(axis, angle):
    a = _normalize(axis)
    c, s = np.cos(angle), np.sin(angle)
    x, y, z = a
    m = np.array([[c + (1 - c) * x * x, (1 - c) * x * y + s * z, (1 - c) * x * z - s * y],
                  [(1 - c) * x * y - s * z, c + (1 - c) * y * y, (1 - c) * y * z + s * x],
                  [(1 - c) * x * z + s * y, (1 - c) * y * z - s * x, c + (1 - c) * z * z]])
    f = np.zeros((4, 3), np.float32)
    f[:3] = m
    return f


def uvsphere(steps=(32, 16), radius=1.0, triangles=True):
    """Lat-long sphere: steps[0] x steps[1] quads (x2 triangles), with normals and texcoords."""
    nu, nv = steps
    u = np.linspace(0, 1, nu + 1)
    v = np.linspace(0, 1, nv + 1)
    uu, vv = np.meshgrid(u, v)
    phi, theta = uu * 2 * np.pi, vv * np.pi
    n = np.stack([np.cos(phi) * np.sin(theta), np.cos(theta), np.sin(phi) * np.sin(theta)], -1)
    normals = n.reshape(-1, 3).astype(np.float32)
    positions = (normals * radius).astype(np.float32)
    texcoords = np.stack([uu, vv], -1).reshape(-1, 2).astype(np.float32)
    idx = lambda i, j: j * (nu + 1) + i
    quads = np.array([[idx(i, j), idx(i + 1, j), idx(i + 1, j + 1), idx(i, j + 1)]
                      for j in range(nv) for i in range(nu)], np.int32)
    if not triangles:
        return dict(quads=quads, positions=positions, normals=normals, texcoords=texcoords)
    tris = np.concatenate([quads[:, [0, 1, 3]], quads[:, [2, 3, 1]]], 1).reshape(-1, 3)
    return dict(triangles=tris, positions=positions, normals=normals, texcoords=texcoords)


def rect_y(scale=1.0, y=0.0, quads=True):
    """Horizontal rectangle facing +y."""
    p = np.array([[-1, 0, 1], [1, 0, 1], [1, 0, -1], [-1, 0, -1]], np.float32) * scale
    p[:, 1] = y
    n = np.tile(np.array([[0, 1, 0]], np.float32), (4, 1))
    t = np.array([[0, 1], [1, 1], [1, 0], [0, 0]], np.float32)
    if quads:
        return dict(quads=np.array([[0, 1, 2, 3]], np.int32), positions=p, normals=n, texcoords=t)
    return dict(triangles=np.array([[0, 1, 2], [2, 3, 0]], np.int32), positions=p, normals=n,
                texcoords=t)


def geoblob(subdiv=6, radius=1.0, bump=0.08, seed=7):
    """Icosphere subdivided `subdiv` times (20*4^subdiv triangles), radially perturbed — the
    stand-in for the Stanford bunny (no file IO on the path)."""
    t = (1 + 5 ** 0.5) / 2
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t),
         (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4),
         (11, 10, 2), (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9),
         (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    verts = np.array(v, np.float64)
    verts /= np.linalg.norm(verts, axis=1, keepdims=True)
    faces = np.array(f, np.int64)
    for _ in range(subdiv):
        e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]])
        e.sort(axis=1)
        ue, inv = np.unique(e, axis=0, return_inverse=True)
        mid = verts[ue[:, 0]] + verts[ue[:, 1]]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        base = len(verts)
        verts = np.concatenate([verts, mid])
        nf = len(faces)
        a, b, c = inv[:nf] + base, inv[nf:2 * nf] + base, inv[2 * nf:] + base
        faces = np.concatenate([np.stack([faces[:, 0], a, c], 1), np.stack([faces[:, 1], b, a], 1),
                                np.stack([faces[:, 2], c, b], 1), np.stack([a, b, c], 1)])
    rng = np.random.default_rng(seed)
    k = rng.normal(size=(6, 3))
    disp = sum(np.sin(verts @ kk * (2 + i)) for i, kk in enumerate(k)) / 6
    pos = verts * (radius * (1 + bump * disp))[:, None]
    return dict(triangles=faces.astype(np.int32), positions=pos.astype(np.float32))


def sky_texture(width=512, height=256):
    """Float lat-long environment: gradient sky + bright sun disk (stands in for make_sunsky)."""
    v = (np.arange(height) + 0.5) / height
    u = (np.arange(width) + 0.5) / width
    uu, vv = np.meshgrid(u, v)
    th, ph = vv * np.pi, uu * 2 * np.pi
    d = np.stack([np.cos(ph) * np.sin(th), np.cos(th), np.sin(ph) * np.sin(th)], -1)
    sun = _normalize([0.4, 0.7, 0.55])
    cosang = d @ sun
    sky = np.clip(d[..., 1], 0, 1)[..., None] * np.array([0.3, 0.5, 1.0]) + 0.15
    sky = np.where(d[..., 1:2] < 0, 0.1, sky)
    sky = sky + (cosang[..., None] > 0.995) * np.array([40.0, 36.0, 30.0])
    return np.concatenate([sky, np.ones_like(sky[..., :1])], -1).astype(np.float32)


def checker_texture(size=64, tiles=8, srgb=True):
    i, j = np.meshgrid(np.arange(size), np.arange(size))
    c = ((i * tiles // size + j * tiles // size) % 2).astype(np.float32)
    img = np.stack([0.2 + 0.6 * c, 0.3 + 0.4 * c, 0.8 - 0.5 * c, np.ones_like(c)], -1)
    return (img * 255).astype(np.uint8) if srgb else img.astype(np.float32)


# ---------------------------------------------------------------------------------------------
def cornellbox():
    """C1: the reference's procedural Cornell box (yocto_scene.cpp:970), loaded from the golden
    fixture generated from the reference itself (tests/golden/make_golden.py)."""
    with np.load(os.path.join(_GOLDEN, "cornellbox_scene.npz")) as d:
        return Scene.from_npz_dict(dict(d))


def cornellbox_quads():
    """Cornell box with every wall/box face as ONE quad (BASELINE configs[0] wording) to exercise
    intersect_quad; same camera/materials as cornellbox()."""
    sc = cornellbox()
    out = Scene()
    out.cameras, out.materials, out.environments = sc.cameras, sc.materials, sc.environments
    for s in sc.shapes:
        tris = s["triangles"]
        quads = []
        for k in range(0, len(tris), 2):
            a, b = tris[k], tris[k + 1]
            # make_cornellbox walls are quads split as (0,1,2),(2,3,0)
            quads.append([a[0], a[1], a[2], b[1]])
        out.add_shape(quads=np.array(quads, np.int32), positions=s["positions"])
    out.instances = sc.instances
    return out


def instanced_spheres(n=10, seed=11):
    """C3: n^3 instances of one 1024-triangle sphere (+ floor, area light, constant environment)."""
    sc = Scene()
    span = float(n)
    sc.add_camera(lookat_frame((0.9 * n, 1.1 * n, 1.6 * n), (n / 2 - 0.5, n / 2 - 0.5, n / 2 - 0.5)),
                  lens=0.05, film=0.036, aspect=16 / 9, focus=2.0 * n, aperture=0.0)
    m_matte = sc.add_material(abi.MATTE, color=(0.7, 0.25, 0.2))
    m_glossy = sc.add_material(abi.GLOSSY, color=(0.2, 0.6, 0.25), roughness=0.2)
    m_metal = sc.add_material(abi.REFLECTIVE, color=(0.8, 0.7, 0.4), roughness=0.15)
    m_white = sc.add_material(abi.MATTE, color=(0.75, 0.75, 0.75))
    m_floor = sc.add_material(abi.MATTE, color=(0.5, 0.5, 0.5))
    m_light = sc.add_material(abi.MATTE, emission=(30, 30, 30))
    sphere = sc.add_shape(**{k: v for k, v in uvsphere((32, 16), 0.4).items() if k != "texcoords"})
    floor = sc.add_shape(**rect_y(2.0 * span, -0.6))
    light = rect_y(0.35 * span, 0.0)
    light["positions"] = light["positions"][::-1].copy()  # face down
    light["normals"] = -light["normals"]
    lshape = sc.add_shape(**light)
    rng = np.random.default_rng(seed)
    mats = [m_matte, m_glossy, m_metal, m_white]
    for k in range(n):
        for j in range(n):
            for i in range(n):
                idx = (k * n + j) * n + i
                if idx % 7 == 3:  # rotated + non-uniformly scaled subset: non-trivial inverse(frame)
                    f = rotation(rng.normal(size=3), rng.uniform(0, np.pi))
                    f[:3] *= rng.uniform(0.7, 1.2, size=(3, 1)).astype(np.float32)
                else:
                    f = abi.IDENTITY_FRAME.copy()
                f[3] = (i, j, k)
                sc.add_instance(sphere, mats[idx % 4], f)
    sc.add_instance(floor, m_floor, translation((n / 2 - 0.5, 0, n / 2 - 0.5)))
    sc.add_instance(lshape, m_light, translation((n / 2 - 0.5, n + 1.5, n / 2 - 0.5)))
    sc.add_environment(emission=(0.5, 0.5, 0.5))
    return sc


def bunny_like(subdiv=6):
    """C2: ~82K-triangle blob, glossy, on a floor quad, lit by an environment TEXTURE (exercises
    the env-map CDF sampling) — 1280x720 with resolution=1280."""
    sc = Scene()
    sc.add_camera(lookat_frame((0.0, 1.2, 3.2), (0, 0.55, 0)), lens=0.05, film=0.036, aspect=16 / 9,
                  focus=3.3, aperture=0.0)
    tex = sc.add_texture(sky_texture(512, 256))
    m_obj = sc.add_material(abi.GLOSSY, color=(0.8, 0.5, 0.3), roughness=0.25)
    m_floor = sc.add_material(abi.MATTE, color=(0.6, 0.6, 0.6))
    blob = sc.add_shape(**geoblob(subdiv, 0.6))
    floor = sc.add_shape(**rect_y(4.0, 0.0))
    sc.add_instance(blob, m_obj, translation((0, 0.62, 0)))
    sc.add_instance(floor, m_floor)
    sc.add_environment(emission=(1, 1, 1), emission_tex=tex)
    return sc


def hair_strands(num_strands, steps, seed=7, radius=0.5, length=0.25, thickness=(0.004, 0.001)):
    rng = np.random.default_rng(seed)
    root = rng.normal(size=(num_strands, 3))
    root /= np.linalg.norm(root, axis=1, keepdims=True)
    t = np.linspace(0, 1, steps + 1)[None, :, None]
    noise = rng.normal(scale=0.15, size=(num_strands, 1, 3))
    grav = np.array([0, -0.35, 0])[None, None, :]
    pos = root[:, None, :] * (radius + length * t) + (noise + grav) * (length * t ** 2)
    rad = thickness[0] + (thickness[1] - thickness[0]) * np.broadcast_to(t[..., 0], pos.shape[:2])
    base = (np.arange(num_strands) * (steps + 1))[:, None] + np.arange(steps)[None, :]
    lines = np.stack([base, base + 1], -1).reshape(-1, 2)
    pos = pos.reshape(-1, 3)
    tang = np.gradient(pos.reshape(num_strands, steps + 1, 3), axis=1).reshape(-1, 3)
    tang /= np.maximum(np.linalg.norm(tang, axis=1, keepdims=True), 1e-12)
    return dict(lines=lines.astype(np.int32), positions=pos.astype(np.float32),
                normals=tang.astype(np.float32), radius=rad.reshape(-1).astype(np.float32))


def hair_scene(num_strands=65536, steps=8, tri_subdiv=6):
    """C5: num_strands*steps line segments (524,288 by default) + 20*4^subdiv + 2 triangles,
    glossy hair, subsurface scalp, refractive blob, area light + environment."""
    sc = Scene()
    sc.add_camera(lookat_frame((0.0, 0.6, 2.6), (0, 0.05, 0)), lens=0.05, film=0.036, aspect=16 / 9,
                  focus=2.6, aperture=0.0)
    m_hair = sc.add_material(abi.GLOSSY, color=(0.35, 0.2, 0.1), roughness=0.3)
    m_skin = sc.add_material(abi.SUBSURFACE, color=(0.8, 0.55, 0.45), roughness=0.3,
                             scattering=(0.5, 0.3, 0.25), trdepth=0.05, ior=1.4)
    m_glass = sc.add_material(abi.REFRACTIVE, color=(0.85, 0.95, 0.9), roughness=0.0, ior=1.5,
                              trdepth=0.5)
    m_floor = sc.add_material(abi.MATTE, color=(0.55, 0.55, 0.55))
    m_light = sc.add_material(abi.MATTE, emission=(25, 25, 25))
    hair = sc.add_shape(**hair_strands(num_strands, steps))
    scalp = sc.add_shape(**geoblob(tri_subdiv, 0.5, bump=0.0))
    glass = sc.add_shape(**{k: v for k, v in uvsphere((64, 32), 0.3).items() if k != "texcoords"})
    floor = sc.add_shape(**rect_y(4.0, -0.85))
    light = rect_y(0.8, 0.0)
    light["positions"] = light["positions"][::-1].copy()
    lshape = sc.add_shape(**{k: v for k, v in light.items() if k != "normals"})
    sc.add_instance(hair, m_hair)
    sc.add_instance(scalp, m_skin)
    sc.add_instance(glass, m_glass, translation((1.1, -0.5, 0.4)))
    sc.add_instance(floor, m_floor)
    sc.add_instance(lshape, m_light, translation((0, 2.2, 0.5)))
    sc.add_environment(emission=(0.3, 0.35, 0.45))
    return sc


def hair_stress():
    """C5 at the size BASELINE.json states: 524,288 line segments (65,536 strands x 8) + 200,704 triangles
    (subsurface scalp 81,920 + refractive blob 81,920 + glossy sphere 36,864) + floor and light quads; glossy hair,
    subsurface scalp (volume path), rough refractive blob, area light + environment."""
    sc = hair_scene(65536, 8, 6)
    m_glass2 = sc.add_material(abi.REFRACTIVE, color=(0.8, 0.9, 0.95), roughness=0.1, ior=1.45, trdepth=0.4,
                               scattering=(0.1, 0.1, 0.1))
    m_ball = sc.add_material(abi.GLOSSY, color=(0.25, 0.3, 0.7), roughness=0.15)
    blob = sc.add_shape(**geoblob(6, 0.32, bump=0.1, seed=13))
    ball = sc.add_shape(**{k: v for k, v in uvsphere((192, 96), 0.3).items() if k != "texcoords"})
    sc.add_instance(blob, m_glass2, translation((-1.15, -0.5, 0.35)))
    sc.add_instance(ball, m_ball, translation((0.2, -0.55, 1.1)))
    return sc


def sliver_chain(n=100):
    """A mesh whose split_middle tree degenerates into a chain: triangle k sits at x = 0.45^k (then, once float32
    runs out of exponent, along y) with a size proportional to its coordinate, so the midpoint of the centroid box
    peels ONE triangle off per level (yocto_bvh.cpp:202-232). Depth ~ n - 3: n = 100 needs more traversal stack than
    64 entries but fits the reference's 128 (yocto_bvh.cpp:469); n = 170 does not fit either."""
    nx = min(n, 105)
    k = np.arange(n, dtype=np.float64)
    x = 0.45 ** np.minimum(k, nx - 1)
    y = np.where(k >= nx, 0.9 * 0.45 ** np.maximum(k - nx, 0), 0.0)
    s = np.where(k >= nx, y, x) / 8
    c = np.stack([x, y, np.zeros(n)], 1)
    p = np.zeros((n, 3, 3))
    p[:, 0] = c + np.stack([np.zeros(n), -s, -s], 1)
    p[:, 1] = c + np.stack([np.zeros(n), s, -s], 1)
    p[:, 2] = c + np.stack([np.zeros(n), np.zeros(n), 2 * s], 1)
    sc = Scene()
    sc.add_camera(lookat_frame((3.0, 0.2, 0.1), (0, 0, 0)), lens=0.05, film=0.036, aspect=1.0, focus=3.0, aperture=0.0)
    m = sc.add_material(abi.MATTE, color=(0.7, 0.7, 0.7))
    ml = sc.add_material(abi.MATTE, emission=(5, 5, 5))
    sc.add_instance(sc.add_shape(triangles=np.arange(3 * n, dtype=np.int32).reshape(-1, 3),
                                 positions=p.reshape(-1, 3).astype(np.float32)), m)
    light = rect_y(0.5, 0.0)
    light["positions"] = light["positions"][::-1].copy()
    sc.add_instance(sc.add_shape(**{k_: v for k_, v in light.items() if k_ != "normals"}), ml, translation((0.5, 1.5, 0)))
    sc.add_environment(emission=(0.4, 0.4, 0.4))
    return sc


def camera_rays(scene, params, n, seed=5):
    """n primary rays of the camera the config renders with (pixel positions drawn uniformly; thin-lens centre):
    input for the batch-intersection parity tests (SURVEY.md §8d "all camera rays of the config")."""
    cam = scene.cameras[params.camera]
    rng = np.random.default_rng(seed)
    uv = rng.random((n, 2)).astype(np.float32)
    film = np.float32(cam["film"])
    aspect = np.float32(cam["aspect"])
    fw, fh = (film, film / aspect) if aspect >= 1 else (film * aspect, film)
    f = cam["frame"].astype(np.float32)
    q = np.stack([fw * (np.float32(0.5) - uv[:, 0]), fh * (uv[:, 1] - np.float32(0.5)),
                  np.full(n, cam["lens"], np.float32)], 1)
    d = -(q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    rays = np.zeros(n, abi.RAY_DTYPE)
    rays["d"] = (d @ f[:3]).astype(np.float32)
    rays["o"] = f[3]
    rays["tmin"] = 1e-4
    rays["tmax"] = np.finfo(np.float32).max
    return rays


def _write_scene(tmpdir, name, doc, pool_dir):
    """Writes <tmpdir>/<name>/<name>.json next to shapes/ and textures/ links into a pool of reference test assets and
    loads it through the library's own loader (ygl_scene_load)."""
    import json
    from . import lib
    d = os.path.join(str(tmpdir), name)
    os.makedirs(d, exist_ok=True)
    for kind in ("shapes", "textures"):
        link = os.path.join(d, kind)
        if not os.path.lexists(link):
            os.symlink(os.path.join(pool_dir, kind), link)
    path = os.path.join(d, name + ".json")
    with open(path, "w") as f:
        json.dump(doc, f)
    return lib.load_scene(path)


def _frame(f):
    return [float(x) for x in np.asarray(f, np.float32).reshape(-1)]


def bunny_file_scene(tmpdir, pool_dir):
    """C2 with the asset BASELINE.json names: the Stanford bunny of the reference's test data (bunny.ply, 144,046
    triangles) + floor, glossy, lit by the reference's sky.hdr environment texture; 16:9 camera."""
    doc = {
        "asset": {"version": "4.2"},
        "cameras": [{"name": "default", "frame": _frame(lookat_frame((0.0, 0.12, 0.38), (0, 0.07, 0))), "lens": 0.05,
                     "film": 0.036, "aspect": 16 / 9, "focus": 0.4}],
        "textures": [{"name": "sky", "uri": "textures/sky.hdr"}, {"name": "floor", "uri": "textures/floor.png"}],
        "materials": [{"name": "floor", "color": [0.7, 0.7, 0.7], "color_tex": 1},
                      {"name": "bunny", "type": "glossy", "color": [0.8, 0.5, 0.3], "roughness": 0.25}],
        "shapes": [{"name": "floor", "uri": "shapes/floor.ply"}, {"name": "bunny", "uri": "shapes/bunny.ply"}],
        "instances": [{"name": "floor", "shape": 0, "material": 0}, {"name": "bunny", "shape": 1, "material": 1}],
        "environments": [{"name": "sky", "emission": [0.5, 0.5, 0.5], "emission_tex": 0}],
    }
    return _write_scene(tmpdir, "c2_bunny", doc, pool_dir)


def hairball_file_scene(tmpdir, pool_dir):
    """C5 with the reference's assets: two instances of hairball1.ply (2 x 262,144 line segments), two bunnies
    (2 x 144,046 triangles, subsurface and rough refractive), a glossy sphere, floor, area light, sky."""
    doc = {
        "asset": {"version": "4.2"},
        "cameras": [{"name": "default", "frame": _frame(lookat_frame((0.0, 0.2, 0.95), (0, 0.07, 0))), "lens": 0.05,
                     "film": 0.036, "aspect": 16 / 9, "focus": 0.96}],
        "textures": [{"name": "sky", "uri": "textures/sky.hdr"}],
        "materials": [{"name": "floor", "color": [0.55, 0.55, 0.55]},
                      {"name": "hair", "type": "glossy", "color": [0.35, 0.2, 0.1], "roughness": 0.3},
                      {"name": "skin", "type": "subsurface", "color": [0.8, 0.55, 0.45], "roughness": 0.3,
                       "scattering": [0.5, 0.3, 0.25], "trdepth": 0.05, "ior": 1.4},
                      {"name": "glass", "type": "refractive", "color": [0.85, 0.95, 0.9], "roughness": 0.1, "trdepth": 0.5},
                      {"name": "ball", "type": "glossy", "color": [0.25, 0.3, 0.7], "roughness": 0.15},
                      {"name": "light", "emission": [25, 25, 25]}],
        "shapes": [{"uri": "shapes/floor.ply"}, {"uri": "shapes/hairball1.ply"}, {"uri": "shapes/bunny.ply"},
                   {"uri": "shapes/sphere.ply"}, {"uri": "shapes/arealight1.ply"}],
        "instances": [{"shape": 0, "material": 0},
                      {"shape": 1, "material": 1, "frame": _frame(translation((-0.09, 0.0, 0.0)))},
                      {"shape": 1, "material": 1, "frame": _frame(np.vstack([rotation((0, 1, 0), 1.3)[:3], [[0.09, 0.0, 0.02]]]))},
                      {"shape": 2, "material": 2, "frame": _frame(translation((-0.24, 0.0, 0.12)))},
                      {"shape": 2, "material": 3, "frame": _frame(translation((0.25, 0.0, 0.1)))},
                      {"shape": 3, "material": 4, "frame": _frame(np.vstack([np.eye(3) * 0.45, [[0.0, 0.0, 0.22]]]))},
                      {"shape": 4, "material": 5}],
        "environments": [{"emission": [0.4, 0.45, 0.55], "emission_tex": 0}],
    }
    return _write_scene(tmpdir, "c5_hairball", doc, pool_dir)


def features(seed=3):
    """Small scene that touches every code path: triangles, quads, lines, points; all 8 material
    types incl. delta / rough variants, opacity, vertex colors, color/roughness/normal/emission
    textures (byte sRGB and float), environment texture, rotated+scaled instances; 2 cameras
    (thin-lens with aperture, orthographic)."""
    sc = Scene()
    sc.add_camera(lookat_frame((0, 1.6, 5.0), (0, 0.5, 0)), lens=0.05, film=0.036, aspect=2.0,
                  focus=5.0, aperture=0.05)
    sc.add_camera(lookat_frame((0, 2.5, 4.0), (0, 0.4, 0)), orthographic=True, lens=0.02, film=0.036,
                  aspect=2.0, focus=4.5, aperture=0.0)
    t_checker = sc.add_texture(checker_texture(64, 8, True))
    t_rough = sc.add_texture(checker_texture(32, 4, False), linear=True)
    rng = np.random.default_rng(seed)
    nm = rng.normal(scale=0.25, size=(32, 32, 3))
    nm[..., 2] = 1
    nm /= np.linalg.norm(nm, axis=-1, keepdims=True)
    nmap = np.concatenate([nm * 0.5 + 0.5, np.ones((32, 32, 1))], -1)
    t_normal = sc.add_texture((nmap * 255).astype(np.uint8), linear=True)
    t_sky = sc.add_texture(sky_texture(128, 64))
    t_emit = sc.add_texture(checker_texture(16, 2, True), nearest=True, clamp=True)

    mats = [
        sc.add_material(abi.MATTE, color=(0.8, 0.8, 0.8), color_tex=t_checker),
        sc.add_material(abi.GLOSSY, color=(0.7, 0.3, 0.3), roughness=0.3, roughness_tex=t_rough),
        sc.add_material(abi.REFLECTIVE, color=(0.9, 0.8, 0.5), roughness=0.0),
        sc.add_material(abi.REFLECTIVE, color=(0.6, 0.7, 0.9), roughness=0.25, normal_tex=t_normal),
        sc.add_material(abi.TRANSPARENT, color=(0.9, 0.9, 1.0), roughness=0.0, ior=1.5),
        sc.add_material(abi.TRANSPARENT, color=(1.0, 0.9, 0.8), roughness=0.2, ior=1.5),
        sc.add_material(abi.REFRACTIVE, color=(0.9, 1.0, 0.9), roughness=0.0, ior=1.5, trdepth=1.0),
        sc.add_material(abi.REFRACTIVE, color=(0.8, 0.9, 1.0), roughness=0.15, ior=1.33,
                        scattering=(0.2, 0.2, 0.2), trdepth=0.5),
        sc.add_material(abi.SUBSURFACE, color=(0.9, 0.6, 0.5), roughness=0.2,
                        scattering=(0.6, 0.4, 0.3), scanisotropy=0.3, trdepth=0.1),
        sc.add_material(abi.VOLUMETRIC, color=(0.7, 0.8, 0.9), scattering=(0.8, 0.8, 0.8),
                        scanisotropy=-0.2, trdepth=0.6),
        sc.add_material(abi.GLTFPBR, color=(0.8, 0.5, 0.2), roughness=0.4, metallic=0.7),
        sc.add_material(abi.MATTE, color=(0.3, 0.8, 0.4), opacity=0.5),
    ]
    m_floor = sc.add_material(abi.MATTE, color=(0.6, 0.6, 0.6), color_tex=t_checker)
    m_light = sc.add_material(abi.MATTE, emission=(12, 12, 12), emission_tex=t_emit)
    m_hair = sc.add_material(abi.GLOSSY, color=(0.4, 0.25, 0.15), roughness=0.3)
    m_pts = sc.add_material(abi.MATTE, color=(0.9, 0.2, 0.2))

    sph_t = sc.add_shape(**uvsphere((24, 12), 0.35))
    sph_q = sc.add_shape(**uvsphere((24, 12), 0.35, triangles=False))
    colored = uvsphere((16, 8), 0.35)
    colored["colors"] = np.concatenate(
        [0.5 + 0.5 * colored["normals"], np.ones((len(colored["normals"]), 1), np.float32)], 1)
    sph_c = sc.add_shape(**colored)
    for k, m in enumerate(mats):
        row, col = divmod(k, 6)
        f = abi.IDENTITY_FRAME.copy() if k % 3 else rotation((1, 1, 0), 0.6)
        if k == 3:
            f[:3] *= np.array([[1.0], [0.8], [1.1]], np.float32)
        f[3] = (-2.25 + 0.9 * col, 0.36 + 0.9 * row, -0.4 * row)
        shape = sph_q if k % 2 else sph_t
        if k == 0:
            shape = sph_c
        sc.add_instance(shape, m, f)
    floor = sc.add_shape(**rect_y(5.0, 0.0))
    sc.add_instance(floor, m_floor)
    light = rect_y(0.6, 0.0, quads=False)
    light["positions"] = light["positions"][::-1].copy()
    light["normals"] = -light["normals"]
    sc.add_instance(sc.add_shape(**light), m_light, translation((0.5, 3.2, 1.0)))
    hair = hair_strands(400, 4, seed=5, radius=0.3, length=0.25, thickness=(0.01, 0.003))
    sc.add_instance(sc.add_shape(**hair), m_hair, translation((2.6, 0.45, 1.2)))
    npts = 300
    ppos = rng.uniform(-0.4, 0.4, size=(npts, 3)).astype(np.float32)
    pts = dict(points=np.arange(npts, dtype=np.int32), positions=ppos,
               radius=np.full(npts, 0.02, np.float32))
    sc.add_instance(sc.add_shape(**pts), m_pts, translation((-2.6, 0.5, 1.3)))
    sc.add_environment(emission=(0.6, 0.6, 0.6), emission_tex=t_sky, frame=rotation((0, 1, 0), 0.7))
    sc.add_environment(emission=(0.05, 0.05, 0.08))
    return sc
