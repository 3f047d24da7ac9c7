"""TEST INFRASTRUCTURE — ctypes binding of oracle/_ref/libyocto_ref.so (the real reference CPU
renderer behind oracle/ref_shim.cpp). Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg may import this; nothing under yocto-gl_b200/ does."""
import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(_HERE, "..", "yocto-gl_b200"))
from ygl_b200 import abi  # noqa: E402


def available(variant=""):
    return os.path.exists(os.path.join(_HERE, "_ref", f"libyocto_ref{variant}.so"))


class Ref:
    """One loaded reference library. variant '' = glibc float libm, '_dlibm' = double-rounded."""

    def __init__(self, variant=""):
        path = os.path.join(_HERE, "_ref", f"libyocto_ref{variant}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path}: run `make -C oracle ref` where /root/reference exists")
        self.lib = lib = C.CDLL(path)
        vp = C.c_void_p
        lib.ref_scene_create.restype = vp
        lib.ref_scene_create.argtypes = [C.POINTER(abi.SceneDesc)]
        lib.ref_scene_cornellbox.restype = vp
        lib.ref_scene_destroy.argtypes = [vp]
        lib.ref_scene_load.restype = vp
        lib.ref_scene_load.argtypes = [C.c_char_p]
        lib.ref_last_error.restype = C.c_char_p
        lib.ref_scene_describe.argtypes = [vp, C.POINTER(abi.SceneDesc)]
        lib.ref_bvh_build.restype = vp
        lib.ref_bvh_build.argtypes = [vp, C.c_int]
        lib.ref_bvh_destroy.argtypes = [vp]
        lib.ref_bvh_update.argtypes = [vp, vp, vp, C.c_int]
        lib.ref_tonemap_image.argtypes = [vp, C.c_int64, C.c_float, C.c_int, C.c_int, vp, vp]
        lib.ref_bvh_tree_size.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.ref_bvh_tree_get.argtypes = [vp, C.c_int, vp, vp]
        lib.ref_intersect_rays.argtypes = [vp, vp, vp, C.c_int64, C.c_int, C.c_int, vp, C.c_int]
        lib.ref_lights_create.restype = vp
        lib.ref_lights_create.argtypes = [vp, C.POINTER(abi.TraceParams)]
        lib.ref_lights_destroy.argtypes = [vp]
        lib.ref_lights_count.argtypes = [vp]
        lib.ref_lights_get.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                       C.POINTER(C.c_int), vp]
        lib.ref_state_rngs.argtypes = [vp, C.POINTER(abi.TraceParams), C.POINTER(C.c_int),
                                       C.POINTER(C.c_int), vp]
        lib.ref_trace_image.restype = C.c_double
        lib.ref_trace_image.argtypes = [vp, C.POINTER(abi.TraceParams), C.c_int,
                                        C.POINTER(C.c_int), C.POINTER(C.c_int), vp, vp, vp, vp, vp]
        lib.ref_rng_floats.argtypes = [C.c_uint64, C.c_uint64, C.c_int, vp, vp]
        lib.ref_sizeof.argtypes = [C.c_char_p]
        lib.ref_counters_read.argtypes = [vp]

    COUNTER_NAMES = ("top_nodes", "bottom_nodes", "instance_visits", "point_tests", "line_tests", "triangle_tests",
                     "quad_tests", "rays")

    def counters_reset(self):
        """Instrumented oracle only (variant '_count'): zero the traversal counters (oracle/ref_counters.h)."""
        if not self.lib.ref_counters_available():
            raise RuntimeError("this reference build carries no traversal counters: use Ref('_count')")
        self.lib.ref_counters_reset()

    def counters(self):
        """{'scene': {...}, 'instance': {...}}: totals since the last reset, split by the kind of query
        (intersect_scene_bvh / intersect_instance_bvh) that was running."""
        out = np.zeros(16, np.uint64)
        self.lib.ref_counters_read(out.ctypes.data)
        return {mode: {n: int(out[8 * m + k]) for k, n in enumerate(self.COUNTER_NAMES)}
                for m, mode in enumerate(("scene", "instance"))}

    def sizeof(self, name):
        return self.lib.ref_sizeof(name.encode())

    def scene(self, scene):
        return RefScene(self, scene)

    def cornellbox(self):
        """The reference's make_cornellbox() (yocto_scene.cpp:970) as an abi.Scene."""
        h = self.lib.ref_scene_cornellbox()
        d = abi.SceneDesc()
        self.lib.ref_scene_describe(h, C.byref(d))
        sc = abi.Scene.from_desc(d)
        self.lib.ref_scene_destroy(h)
        return sc

    def load_scene(self, filename):
        """The reference's load_scene (yocto_sceneio.cpp:2761) as an abi.Scene (arrays copied out)."""
        h = self.lib.ref_scene_load(str(filename).encode())
        if not h:
            raise RuntimeError(self.lib.ref_last_error().decode())
        d = abi.SceneDesc()
        self.lib.ref_scene_describe(h, C.byref(d))
        sc = abi.Scene.from_desc(d)
        self.lib.ref_scene_destroy(h)
        return sc

    def tonemap_image(self, hdr, exposure=0.0, filmic=False, srgb=True):
        """tonemap_image (yocto_image.cpp:911-922): (vec4f image, vec4b image)"""
        hdr = np.ascontiguousarray(hdr, np.float32)
        ldr, ldr_b = np.zeros_like(hdr), np.zeros(hdr.shape, np.uint8)
        self.lib.ref_tonemap_image(hdr.ctypes.data, hdr.size // 4, exposure, int(filmic), int(srgb), ldr.ctypes.data,
                                   ldr_b.ctypes.data)
        return ldr, ldr_b

    def rng_floats(self, seed, seq, n):
        out = np.zeros(n, np.float32)
        st = np.zeros(2, np.uint64)
        self.lib.ref_rng_floats(seed, seq, n, out.ctypes.data, st.ctypes.data)
        return out, st


class RefScene:
    def __init__(self, ref, scene):
        self.ref, self.lib = ref, ref.lib
        self.scene = scene
        self.desc = scene.desc()
        self.h = self.lib.ref_scene_create(C.byref(self.desc))
        self._bvh = {}

    def __del__(self):
        try:
            for b in self._bvh.values():
                self.lib.ref_bvh_destroy(b)
            self.lib.ref_scene_destroy(self.h)
        except Exception:
            pass

    def bvh(self, highquality=False):
        if highquality not in self._bvh:
            self._bvh[highquality] = self.lib.ref_bvh_build(self.h, int(highquality))
        return self._bvh[highquality]

    def adopt_updated_bvh(self, other, updated_shapes, highquality=False):
        """update_scene_bvh: take `other`'s trees (a RefScene of the scene before the edit) refitted to this scene."""
        b = other._bvh.pop(highquality, None) or self.lib.ref_bvh_build(other.h, int(highquality))
        shapes = np.ascontiguousarray(list(updated_shapes), np.int32)
        self.lib.ref_bvh_update(b, self.h, shapes.ctypes.data, len(shapes))
        self._bvh[highquality] = b

    def bvh_tree(self, shape, highquality=False):
        b = self.bvh(highquality)
        nn, npr = C.c_int(), C.c_int()
        self.lib.ref_bvh_tree_size(b, shape, C.byref(nn), C.byref(npr))
        nodes = np.zeros(nn.value, abi.NODE_DTYPE)
        prims = np.zeros(npr.value, np.int32)
        self.lib.ref_bvh_tree_get(b, shape, nodes.ctypes.data, prims.ctypes.data)
        return nodes, prims

    def intersect(self, rays, instance=-1, find_any=False, highquality=False, nthreads=8):
        rays = np.ascontiguousarray(rays, abi.RAY_DTYPE)
        out = np.zeros(len(rays), abi.ISEC_DTYPE)
        self.lib.ref_intersect_rays(self.h, self.bvh(highquality), rays.ctypes.data, len(rays),
                                    instance, int(find_any), out.ctypes.data, nthreads)
        return out

    def lights(self, params=None):
        params = params or abi.trace_params()
        h = self.lib.ref_lights_create(self.h, C.byref(params))
        res = []
        for i in range(self.lib.ref_lights_count(h)):
            inst, env, n = C.c_int(), C.c_int(), C.c_int()
            self.lib.ref_lights_get(h, i, C.byref(inst), C.byref(env), C.byref(n), None)
            cdf = np.zeros(n.value, np.float32)
            self.lib.ref_lights_get(h, i, C.byref(inst), C.byref(env), C.byref(n), cdf.ctypes.data)
            res.append((inst.value, env.value, cdf))
        self.lib.ref_lights_destroy(h)
        return res

    def state_rngs(self, params):
        w, h = C.c_int(), C.c_int()
        self.lib.ref_state_rngs(self.h, C.byref(params), C.byref(w), C.byref(h), None)
        rngs = np.zeros((h.value * w.value, 2), np.uint64)
        self.lib.ref_state_rngs(self.h, C.byref(params), C.byref(w), C.byref(h), rngs.ctypes.data)
        return w.value, h.value, rngs

    def trace_image(self, params, samples=0, full=False):
        """Returns dict(image (h,w,4), seconds, [albedo, normal, hits, rngs])."""
        w, h = C.c_int(), C.c_int()
        self.lib.ref_state_rngs(self.h, C.byref(params), C.byref(w), C.byref(h), None)
        W, H = w.value, h.value
        image = np.zeros((H, W, 4), np.float32)
        albedo = np.zeros((H, W, 3), np.float32) if full else None
        normal = np.zeros((H, W, 3), np.float32) if full else None
        hits = np.zeros((H, W), np.int32) if full else None
        rngs = np.zeros((H * W, 2), np.uint64) if full else None
        p = lambda a: a.ctypes.data if a is not None else None
        secs = self.lib.ref_trace_image(self.h, C.byref(params), samples, C.byref(w), C.byref(h),
                                        p(image), p(albedo), p(normal), p(hits), p(rngs))
        return dict(image=image, albedo=albedo, normal=normal, hits=hits, rngs=rngs,
                    seconds=secs, width=W, height=H)
