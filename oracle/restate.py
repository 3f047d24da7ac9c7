"""TEST INFRASTRUCTURE — ctypes binding of oracle/_build/libygl_oracle.so (the plain-C restatement
in oracle/restate/ygl_oracle.c). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg may import this."""
import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(_HERE, "..", "yocto-gl_b200"))
from ygl_b200 import abi  # noqa: E402

_PATH = os.path.join(_HERE, "_build", "libygl_oracle.so")
_lib = None


def available():
    return os.path.exists(_PATH)


def load():
    global _lib
    if _lib is None:
        if not available():
            raise FileNotFoundError(f"{_PATH}: run `make -C oracle restate`")
        lib = C.CDLL(_PATH)
        vp = C.c_void_p
        lib.oracle_scene_create.restype = vp
        lib.oracle_scene_create.argtypes = [C.POINTER(abi.SceneDesc), C.c_int]
        lib.oracle_scene_destroy.argtypes = [vp]
        lib.oracle_supported.argtypes = [C.POINTER(abi.SceneDesc)]
        lib.oracle_tree_size.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.oracle_tree_get.argtypes = [vp, C.c_int, vp, vp]
        lib.oracle_intersect_rays.argtypes = [vp, vp, C.c_int64, C.c_int, C.c_int, vp]
        lib.oracle_state_size.argtypes = [C.POINTER(abi.SceneDesc), C.POINTER(abi.TraceParams),
                                          C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.oracle_state_rngs.argtypes = [C.POINTER(abi.TraceParams), C.c_int, C.c_int, vp]
        lib.oracle_trace_image.argtypes = [vp, C.POINTER(abi.TraceParams), C.c_int, vp]
        _lib = lib
    return _lib


class OracleScene:
    def __init__(self, scene, highquality=False):
        self.lib = load()
        self.scene = scene
        self.desc = scene.desc()
        self.h = self.lib.oracle_scene_create(C.byref(self.desc), int(highquality))

    def __del__(self):
        try:
            self.lib.oracle_scene_destroy(self.h)
        except Exception:
            pass

    def supported(self):
        return bool(self.lib.oracle_supported(C.byref(self.desc)))

    def tree(self, shape):
        nn, npr = C.c_int(), C.c_int()
        self.lib.oracle_tree_size(self.h, shape, C.byref(nn), C.byref(npr))
        nodes, prims = np.zeros(nn.value, abi.NODE_DTYPE), np.zeros(npr.value, np.int32)
        self.lib.oracle_tree_get(self.h, shape, nodes.ctypes.data, prims.ctypes.data)
        return nodes, prims

    def intersect(self, rays, instance=-1, find_any=False):
        rays = np.ascontiguousarray(rays, abi.RAY_DTYPE)
        out = np.zeros(len(rays), abi.ISEC_DTYPE)
        self.lib.oracle_intersect_rays(self.h, rays.ctypes.data, len(rays), instance, int(find_any),
                                       out.ctypes.data)
        return out

    def state_rngs(self, params):
        w, h = C.c_int(), C.c_int()
        self.lib.oracle_state_size(C.byref(self.desc), C.byref(params), C.byref(w), C.byref(h))
        rngs = np.zeros((w.value * h.value, 2), np.uint64)
        self.lib.oracle_state_rngs(C.byref(params), w.value, h.value, rngs.ctypes.data)
        return w.value, h.value, rngs

    def trace_image(self, params, samples=0):
        w, h = C.c_int(), C.c_int()
        self.lib.oracle_state_size(C.byref(self.desc), C.byref(params), C.byref(w), C.byref(h))
        image = np.zeros((h.value, w.value, 4), np.float32)
        rc = self.lib.oracle_trace_image(self.h, C.byref(params), samples or params.samples,
                                         image.ctypes.data)
        if rc != 0:
            raise NotImplementedError("scene/params use features outside the C restatement")
        return image


def trace_image(scene, params):
    return OracleScene(scene).trace_image(params)
