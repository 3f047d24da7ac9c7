"""Shared helpers for the parity tests: seeded ray batches and image comparison."""
import numpy as np

from ygl_b200 import abi


def scene_bounds(scene):
    lo, hi = np.full(3, np.inf), np.full(3, -np.inf)
    for inst in scene.instances:
        p = scene.shapes[inst["shape"]]["positions"].astype(np.float64)
        if not len(p):
            continue
        f = inst["frame"].astype(np.float64)
        w = p @ f[:3] + f[3]
        lo, hi = np.minimum(lo, w.min(0)), np.maximum(hi, w.max(0))
    return lo, hi


def random_rays(scene, n, seed=7, tmax=None):
    """Origins uniform in 1.2x the scene box, uniform directions."""
    rng = np.random.default_rng(seed)
    lo, hi = scene_bounds(scene)
    c, e = (lo + hi) / 2, (hi - lo) / 2 * 1.2 + 1e-3
    rays = np.zeros(n, abi.RAY_DTYPE)
    rays["o"] = (c + rng.uniform(-1, 1, size=(n, 3)) * e).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["d"] = d.astype(np.float32)
    rays["tmin"] = 1e-4
    rays["tmax"] = np.finfo(np.float32).max if tmax is None else tmax
    return rays


def axis_rays(scene, n, seed=3):
    """Axis-aligned rays (zero direction components -> inf/NaN slabs) starting on box faces."""
    rng = np.random.default_rng(seed)
    lo, hi = scene_bounds(scene)
    rays = random_rays(scene, n, seed)
    axis = rng.integers(0, 3, n)
    sign = rng.choice([-1.0, 1.0], n)
    d = np.zeros((n, 3), np.float32)
    d[np.arange(n), axis] = sign
    rays["d"] = d
    snap = rng.random(n) < 0.5
    o = rays["o"].copy()
    k = rng.integers(0, 3, n)
    o[snap, k[snap]] = np.where(rng.random(snap.sum()) < 0.5, lo[k[snap]], hi[k[snap]]).astype(np.float32)
    rays["o"] = o
    return rays


def compare_hits(a, b):
    """Returns the number of rays whose scene_intersection differs in any bit."""
    hit = a["hit"] != 0
    bad = a["hit"] != b["hit"]
    for f in ("instance", "element"):
        bad |= hit & (a[f] != b[f])
    bad |= hit & (a["distance"].view(np.uint32) != b["distance"].view(np.uint32))
    bad |= hit & (a["uv"].view(np.uint32) != b["uv"].view(np.uint32)).any(axis=1)
    return int(bad.sum())


def image_stats(a, b):
    a64, b64 = a[..., :3].astype(np.float64), b[..., :3].astype(np.float64)
    diff = a64 - b64
    rmse = float(np.sqrt(np.mean(diff ** 2)))
    per_pixel = np.abs(diff).max(axis=-1)
    exact = (a.view(np.uint32) == b.view(np.uint32)).all(axis=-1)
    keep = per_pixel <= 1e-4  # pixels without a diverged path (see tolerance note in test_gpu_parity)
    rmse_same_paths = float(np.sqrt(np.mean(diff[keep] ** 2))) if keep.any() else 0.0
    return dict(rmse=rmse, rmse_same_paths=rmse_same_paths, max_abs=float(per_pixel.max()),
                frac_exact=float(exact.mean()), frac_gt_1e4=float((per_pixel > 1e-4).mean()))


def assert_close_to_reference(stats):
    """Stated float tolerance against the UNMODIFIED reference (glibc float libm): pixels whose
    paths took the same decisions agree to RMSE < 1e-5; a last-bit libm difference may flip a
    discrete decision of a path (light pick, lobe choice, russian roulette), which we bound at
    < 0.2 % of pixels differing by more than 1e-4 at these low sample counts."""
    assert stats["rmse_same_paths"] < 1e-5, stats
    assert stats["frac_gt_1e4"] < 2e-3, stats


def edited_copy(scene, seed=5):
    """the same scene after an animation step: every vertex / radius of every other shape moved, every instance frame
    rotated a little and shifted (topology untouched) - what update_scene_bvh is for"""
    import copy
    rng = np.random.default_rng(seed)
    out = copy.copy(scene)
    out._keep = None
    out.shapes = [dict(s) for s in scene.shapes]
    out.instances = [dict(n) for n in scene.instances]
    updated = []
    for si, s in enumerate(out.shapes):
        if si % 2 == 1 and len(out.shapes) > 1:
            continue
        s["positions"] = (s["positions"] * np.float32(1.03) +
                          rng.uniform(-0.01, 0.01, s["positions"].shape).astype(np.float32))
        if len(s["radius"]):
            s["radius"] = s["radius"] * np.float32(1.25)
        updated.append(si)
    for n in out.instances:
        f = np.array(n["frame"], np.float32).reshape(4, 3).copy()
        a = np.float32(rng.uniform(-0.2, 0.2))
        rot = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], np.float32)
        f[:3] = f[:3] @ rot
        f[3] += rng.uniform(-0.05, 0.05, 3).astype(np.float32)
        n["frame"] = f.reshape(np.shape(n["frame"]))
    return out, updated
