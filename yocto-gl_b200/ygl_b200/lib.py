"""ctypes binding of yocto-gl_b200/lib/libygl_b200.so — the C ABI of include/ygl_b200.h.

The Python surface mirrors the reference's free functions (yocto_trace.h:116-190):
trace_image / make_trace_bvh / make_trace_lights / make_trace_state / trace_samples / get_image,
plus the batch form of intersect_scene_bvh / intersect_instance_bvh. There is no fallback: if
the extension is missing or no CUDA device is usable, calls raise.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("YGL_B200_LIB") or os.path.normpath(os.path.join(_HERE, "..", "lib", "libygl_b200.so"))
_lib = None

EXPORTS = [
    "ygl_last_error", "ygl_version", "ygl_trace_params_default", "ygl_context_create",
    "ygl_context_destroy", "ygl_context_synchronize", "ygl_context_stream", "ygl_scene_create",
    "ygl_scene_update_cameras", "ygl_scene_destroy", "ygl_bvh_build", "ygl_bvh_tree_size",
    "ygl_bvh_tree_get", "ygl_bvh_destroy", "ygl_lights_create", "ygl_lights_count",
    "ygl_lights_get", "ygl_lights_destroy", "ygl_state_create", "ygl_state_create_tile",
    "ygl_state_create_interleaved", "ygl_state_layout", "ygl_state_size", "ygl_state_rows", "ygl_state_download", "ygl_state_upload",
    "ygl_state_destroy", "ygl_make_state_rngs", "ygl_trace_samples", "ygl_trace_image",
    "ygl_trace_counters", "ygl_context_set_profiling", "ygl_context_set_mode", "ygl_trace_timings", "ygl_intersect_rays", "ygl_intersect_rays_device", "ygl_debug_libm", "ygl_comm_id_size",
    "ygl_comm_create_id", "ygl_comm_init", "ygl_tile_rows", "ygl_gather_image", "ygl_comm_destroy",
    "ygl_bvh_build_device", "ygl_bvh_update", "ygl_tonemap_image", "ygl_state_tonemap", "ygl_scene_load", "ygl_loaded_scene_desc", "ygl_loaded_scene_name", "ygl_loaded_scene_destroy",
    "ygl_trace_start", "ygl_trace_cancel", "ygl_trace_wait", "ygl_trace_done", "ygl_trace_preview",
    "ygl_context_set_option", "ygl_context_get_option", "ygl_state_reset", "ygl_trace_sample", "ygl_bvh_create_from_host",
]


class YglError(RuntimeError):
    pass


def load():
    """Loads the extension; raises (loudly) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise YglError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; "
                       "g.build()'` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    P = C.POINTER
    lib.ygl_last_error.restype = C.c_char_p
    lib.ygl_version.restype = C.c_char_p
    lib.ygl_context_stream.restype = vp
    lib.ygl_context_stream.argtypes = [vp]
    lib.ygl_context_create.argtypes = [i32, P(vp)]
    lib.ygl_context_destroy.argtypes = [vp]
    lib.ygl_context_synchronize.argtypes = [vp]
    lib.ygl_scene_create.argtypes = [vp, P(abi.SceneDesc), P(vp)]
    lib.ygl_scene_update_cameras.argtypes = [vp, P(abi.Camera), i32]
    lib.ygl_scene_destroy.argtypes = [vp]
    lib.ygl_bvh_build.argtypes = [P(abi.SceneDesc), i32, P(vp)]
    lib.ygl_bvh_build_device.argtypes = [vp, P(abi.SceneDesc), i32, P(vp)]
    lib.ygl_tonemap_image.argtypes = [vp, vp, C.c_int64, C.c_float, i32, i32, vp, vp]
    lib.ygl_state_tonemap.argtypes = [vp, C.c_float, i32, i32, vp, vp]
    lib.ygl_bvh_update.argtypes = [vp, P(abi.SceneDesc), vp, i32, vp, i32]
    lib.ygl_bvh_tree_size.argtypes = [vp, i32, P(i32), P(i32)]
    lib.ygl_bvh_tree_get.argtypes = [vp, i32, vp, vp]
    lib.ygl_bvh_destroy.argtypes = [vp]
    lib.ygl_lights_create.argtypes = [P(abi.SceneDesc), P(vp)]
    lib.ygl_lights_count.argtypes = [vp]
    lib.ygl_lights_get.argtypes = [vp, i32, P(i32), P(i32), P(i32), vp]
    lib.ygl_lights_destroy.argtypes = [vp]
    lib.ygl_state_create.argtypes = [vp, P(abi.SceneDesc), P(abi.TraceParams), P(vp)]
    lib.ygl_state_create_tile.argtypes = [vp, P(abi.SceneDesc), P(abi.TraceParams), i32, i32, P(vp)]
    lib.ygl_state_create_interleaved.argtypes = [vp, P(abi.SceneDesc), P(abi.TraceParams), i32, i32, P(vp)]
    lib.ygl_state_layout.argtypes = [vp, P(i32), P(i32), P(i32)]
    lib.ygl_state_size.argtypes = [vp, P(i32), P(i32), P(i32)]
    lib.ygl_state_rows.argtypes = [vp, P(i32), P(i32)]
    lib.ygl_state_download.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.ygl_state_upload.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    lib.ygl_state_destroy.argtypes = [vp]
    lib.ygl_make_state_rngs.argtypes = [P(abi.SceneDesc), P(abi.TraceParams), P(i32), P(i32), vp]
    lib.ygl_trace_samples.argtypes = [vp, vp, vp, vp, vp, P(abi.TraceParams)]
    lib.ygl_trace_image.argtypes = [vp, P(abi.SceneDesc), P(abi.TraceParams), P(i32), P(i32), vp]
    lib.ygl_trace_counters.argtypes = [vp, vp]
    lib.ygl_context_set_profiling.argtypes = [vp, i32, i32]
    lib.ygl_context_set_mode.argtypes = [vp, i32]
    lib.ygl_trace_timings.argtypes = [vp, vp]
    lib.ygl_scene_load.argtypes = [C.c_char_p, P(vp)]
    lib.ygl_loaded_scene_desc.argtypes = [vp]
    lib.ygl_loaded_scene_desc.restype = P(abi.SceneDesc)
    lib.ygl_loaded_scene_name.argtypes = [vp, i32, i32]
    lib.ygl_loaded_scene_name.restype = C.c_char_p
    lib.ygl_loaded_scene_destroy.argtypes = [vp]
    lib.ygl_trace_start.argtypes = [vp, vp, vp, vp, vp, P(abi.TraceParams)]
    lib.ygl_trace_cancel.argtypes = [vp]
    lib.ygl_trace_wait.argtypes = [vp]
    lib.ygl_trace_done.argtypes = [vp]
    lib.ygl_trace_preview.argtypes = [vp, vp, vp, vp, P(abi.TraceParams), i32, i32, vp]
    lib.ygl_context_set_option.argtypes = [vp, C.c_char_p, C.c_double]
    lib.ygl_context_get_option.argtypes = [vp, C.c_char_p, P(C.c_double)]
    lib.ygl_state_reset.argtypes = [vp, P(abi.TraceParams)]
    lib.ygl_trace_sample.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, P(abi.TraceParams)]
    lib.ygl_bvh_create_from_host.argtypes = [P(abi.SceneDesc), vp, i32, vp, i32, vp, vp, vp, vp, P(vp)]
    lib.ygl_intersect_rays.argtypes = [vp, vp, vp, vp, i64, i32, i32, vp]
    lib.ygl_intersect_rays_device.argtypes = [vp, vp, vp, vp, i64, i32, i32, vp, vp]
    lib.ygl_debug_libm.argtypes = [vp, i32, vp, vp, i64, vp]
    lib.ygl_comm_create_id.argtypes = [vp]
    lib.ygl_comm_init.argtypes = [vp, vp, i32, i32]
    lib.ygl_tile_rows.argtypes = [i32, i32, i32, P(i32), P(i32)]
    lib.ygl_tile_rows.restype = None
    lib.ygl_gather_image.argtypes = [vp, vp, vp]
    lib.ygl_comm_destroy.argtypes = [vp]
    lib.ygl_comm_destroy.restype = None
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        raise YglError(f"ygl error {rc}: {load().ygl_last_error().decode()}")


def _p(a):
    return a.ctypes.data if a is not None else None


def load_scene(filename):
    """load_scene (yocto_sceneio.h:93) for JSON v4.2 + PLY + PNG/HDR scenes -> abi.Scene (arrays copied out of the
    library's loaded-scene object), with `camera_names` etc. attached."""
    lib = load()
    h = C.c_void_p()
    _check(lib.ygl_scene_load(str(filename).encode(), C.byref(h)))
    try:
        scene = abi.Scene.from_desc(lib.ygl_loaded_scene_desc(h).contents)
        kinds = ("camera", "texture", "material", "shape", "instance", "environment")
        counts = (len(scene.cameras), len(scene.textures), len(scene.materials), len(scene.shapes),
                  len(scene.instances), len(scene.environments))
        for k, (kind, n) in enumerate(zip(kinds, counts)):
            setattr(scene, kind + "_names", [(lib.ygl_loaded_scene_name(h, k, i) or b"").decode() for i in range(n)])
    finally:
        lib.ygl_loaded_scene_destroy(h)
    return scene


class Bvh:
    """make_trace_bvh (host build in the reference's node order), or — trees=(top, [per shape]) with each tree a
    (nodes, primitives) pair in the reference's bvh_tree layout — trees built elsewhere adopted verbatim."""

    def __init__(self, scene, highquality=False, trees=None, device_ctx=None):
        self.lib = load()
        self.desc = scene.desc()
        self.h = C.c_void_p()
        if device_ctx is not None:  # large trees on the GPU (ygl_bvh_build_device), bit-identical to the host build
            _check(self.lib.ygl_bvh_build_device(device_ctx.h, C.byref(self.desc), int(highquality), C.byref(self.h)))
            return
        if trees is None:
            _check(self.lib.ygl_bvh_build(C.byref(self.desc), int(highquality), C.byref(self.h)))
            return
        (top_nodes, top_prims), shape_trees = trees
        keep = [np.ascontiguousarray(top_nodes, abi.NODE_DTYPE), np.ascontiguousarray(top_prims, np.int32)]
        n = len(shape_trees)
        node_ptrs, prim_ptrs = (C.c_void_p * max(1, n))(), (C.c_void_p * max(1, n))()
        num_nodes, num_prims = (C.c_int * max(1, n))(), (C.c_int * max(1, n))()
        for k, (nodes, prims) in enumerate(shape_trees):
            nodes, prims = np.ascontiguousarray(nodes, abi.NODE_DTYPE), np.ascontiguousarray(prims, np.int32)
            keep += [nodes, prims]
            node_ptrs[k], prim_ptrs[k] = nodes.ctypes.data, prims.ctypes.data
            num_nodes[k], num_prims[k] = len(nodes), len(prims)
        _check(self.lib.ygl_bvh_create_from_host(C.byref(self.desc), keep[0].ctypes.data, len(keep[0]),
                                                 keep[1].ctypes.data, len(keep[1]), node_ptrs, num_nodes,
                                                 prim_ptrs, num_prims, C.byref(self.h)))

    def update(self, scene, updated_shapes=(), updated_instances=()):
        """update_scene_bvh: refit to the edited scene (same topology)."""
        self.desc = scene.desc()
        shapes = np.ascontiguousarray(list(updated_shapes), np.int32)
        insts = np.ascontiguousarray(list(updated_instances), np.int32)
        _check(self.lib.ygl_bvh_update(self.h, C.byref(self.desc), insts.ctypes.data, len(insts),
                                       shapes.ctypes.data, len(shapes)))

    def tree(self, shape):
        nn, npr = C.c_int(), C.c_int()
        _check(self.lib.ygl_bvh_tree_size(self.h, shape, C.byref(nn), C.byref(npr)))
        nodes = np.zeros(nn.value, abi.NODE_DTYPE)
        prims = np.zeros(npr.value, np.int32)
        _check(self.lib.ygl_bvh_tree_get(self.h, shape, nodes.ctypes.data, prims.ctypes.data))
        return nodes, prims

    def __del__(self):
        try:
            self.lib.ygl_bvh_destroy(self.h)
        except Exception:
            pass


class Lights:
    """make_trace_lights."""

    def __init__(self, scene):
        self.lib = load()
        self.desc = scene.desc()
        self.h = C.c_void_p()
        _check(self.lib.ygl_lights_create(C.byref(self.desc), C.byref(self.h)))

    def items(self):
        res = []
        for i in range(self.lib.ygl_lights_count(self.h)):
            inst, env, n = C.c_int(), C.c_int(), C.c_int()
            _check(self.lib.ygl_lights_get(self.h, i, C.byref(inst), C.byref(env), C.byref(n), None))
            cdf = np.zeros(n.value, np.float32)
            _check(self.lib.ygl_lights_get(self.h, i, C.byref(inst), C.byref(env), C.byref(n),
                                           cdf.ctypes.data))
            res.append((inst.value, env.value, cdf))
        return res

    def __del__(self):
        try:
            self.lib.ygl_lights_destroy(self.h)
        except Exception:
            pass


def state_size(scene, params):
    """Image size make_trace_state would choose (yocto_trace.cpp:1499-1505)."""
    lib = load()
    desc = scene.desc()
    w, h = C.c_int(), C.c_int()
    _check(lib.ygl_make_state_rngs(C.byref(desc), C.byref(params), C.byref(w), C.byref(h), None))
    return w.value, h.value


def make_state_rngs(scene, params):
    """Image size and rng table of make_trace_state (host only)."""
    lib = load()
    desc = scene.desc()
    w, h = C.c_int(), C.c_int()
    _check(lib.ygl_make_state_rngs(C.byref(desc), C.byref(params), C.byref(w), C.byref(h), None))
    rngs = np.zeros((w.value * h.value, 2), np.uint64)
    _check(lib.ygl_make_state_rngs(C.byref(desc), C.byref(params), C.byref(w), C.byref(h),
                                   rngs.ctypes.data))
    return w.value, h.value, rngs


class Context:
    def __init__(self, device=0):
        self.lib = load()
        self.h = C.c_void_p()
        _check(self.lib.ygl_context_create(device, C.byref(self.h)))

    def tonemap_image(self, hdr, exposure=0.0, filmic=False, srgb=True):
        """tonemap_image (yocto_image.h:242-245) of host pixels (..., 4) float32: (vec4f image, vec4b image)."""
        hdr = np.ascontiguousarray(hdr, np.float32)
        assert hdr.shape[-1] == 4
        ldr, ldr_b = np.zeros_like(hdr), np.zeros(hdr.shape, np.uint8)
        _check(self.lib.ygl_tonemap_image(self.h, _p(hdr), hdr.size // 4, exposure, int(filmic), int(srgb), _p(ldr),
                                          _p(ldr_b)))
        return ldr, ldr_b

    def synchronize(self):
        _check(self.lib.ygl_context_synchronize(self.h))

    @property
    def stream(self):
        return self.lib.ygl_context_stream(self.h)

    def counters(self):
        c = np.zeros(16, np.uint64)
        _check(self.lib.ygl_trace_counters(self.h, c.ctypes.data))
        names = ["camera_samples", "scene_rays", "instance_rays", "iterations", "launches",
                 "extend_launches", "top_nodes", "bottom_nodes", "instance_visits", "triangle_tests",
                 "quad_tests", "line_tests", "point_tests"]
        return {k: int(v) for k, v in zip(names, c)}

    def set_mode(self, mode):
        """Kept for source compatibility: round 1's 'persistent' scheduler was removed, both names select the wavefront."""
        _check(self.lib.ygl_context_set_mode(self.h, {"wavefront": 0, "persistent": 1}[mode]))

    def trace_cancel(self):
        _check(self.lib.ygl_trace_cancel(self.h))

    def trace_wait(self):
        _check(self.lib.ygl_trace_wait(self.h))

    def trace_done(self):
        return bool(self.lib.ygl_trace_done(self.h))

    def set_option(self, name, value):
        """Scheduling knob by name (ygl_context_set_option); never changes a result bit."""
        _check(self.lib.ygl_context_set_option(self.h, name.encode(), float(value)))

    def get_option(self, name):
        v = C.c_double()
        _check(self.lib.ygl_context_get_option(self.h, name.encode(), C.byref(v)))
        return v.value

    def set_profiling(self, time_kernels=False, count_traversal=False):
        _check(self.lib.ygl_context_set_profiling(self.h, int(time_kernels), int(count_traversal)))

    def timings(self):
        t = np.zeros(4, np.float64)
        _check(self.lib.ygl_trace_timings(self.h, t.ctypes.data))
        return dict(extend_ms=float(t[0]), loop_ms=float(t[1]), extend_launches=int(t[2]))

    def libm(self, fn, x, y=None):
        """Device libm on arrays (test hook); fn as in ygl_debug_libm."""
        x = np.ascontiguousarray(x, np.float32)
        y = None if y is None else np.ascontiguousarray(y, np.float32)
        out = np.zeros_like(x)
        _check(self.lib.ygl_debug_libm(self.h, fn, x.ctypes.data, _p(y), x.size, out.ctypes.data))
        return out

    def trace_image(self, scene, params):
        """trace_image: host scene -> host rgba image, everything inside one call."""
        desc = scene.desc()
        w, h = C.c_int(), C.c_int()
        _check(self.lib.ygl_trace_image(self.h, C.byref(desc), C.byref(params), C.byref(w),
                                        C.byref(h), None))
        image = np.zeros((h.value, w.value, 4), np.float32)
        _check(self.lib.ygl_trace_image(self.h, C.byref(desc), C.byref(params), C.byref(w),
                                        C.byref(h), image.ctypes.data))
        return image

    def __del__(self):
        try:
            self.lib.ygl_context_destroy(self.h)
        except Exception:
            pass


class DeviceScene:
    """Device-resident scene + bvh + lights (make_cutrace_scene-style upload)."""

    def __init__(self, ctx, scene, highquality=False, trees=None, device_build=False):
        self.ctx, self.lib, self.scene = ctx, ctx.lib, scene
        self.desc = scene.desc()
        self.h = C.c_void_p()
        _check(self.lib.ygl_scene_create(ctx.h, C.byref(self.desc), C.byref(self.h)))
        self.bvh = Bvh(scene, highquality, trees, device_ctx=ctx if device_build else None)
        self.lights = Lights(scene)

    def trace_start(self, state, params):
        """trace_start (yocto_trace.h:212): one batch on a worker thread; poll ctx.trace_done(), or cancel / wait."""
        self._async_params = params  # keep alive; the library copies it, this is belt and braces
        _check(self.lib.ygl_trace_start(self.ctx.h, state.h, self.h, self.bvh.h, self.lights.h, C.byref(params)))

    def trace_preview(self, params, width, height):
        """trace_preview (yocto_trace.h:220): 1 spp at resolution / pratio, replicated to width x height."""
        image = np.zeros((height, width, 4), np.float32)
        _check(self.lib.ygl_trace_preview(self.ctx.h, self.h, self.bvh.h, self.lights.h, C.byref(params), width,
                                          height, image.ctypes.data))
        return image

    def trace_sample(self, state, i, j, sample, params):
        """trace_sample (yocto_trace.h:173): one sample of one pixel; state.samples is unchanged."""
        _check(self.lib.ygl_trace_sample(self.ctx.h, state.h, self.h, self.bvh.h, self.lights.h, i, j, sample,
                                         C.byref(params)))

    def make_state(self, params, rows=None, interleave=None):
        """rows=(begin, end): contiguous tile; interleave=(rank, nranks): rows rank, rank+nranks, ..."""
        return State(self, params, rows, interleave)

    def trace_samples(self, state, params):
        _check(self.lib.ygl_trace_samples(self.ctx.h, state.h, self.h, self.bvh.h, self.lights.h,
                                          C.byref(params)))

    def intersect(self, rays, instance=-1, find_any=False):
        rays = np.ascontiguousarray(rays, abi.RAY_DTYPE)
        out = np.zeros(len(rays), abi.ISEC_DTYPE)
        _check(self.lib.ygl_intersect_rays(self.ctx.h, self.h, self.bvh.h, rays.ctypes.data,
                                           len(rays), instance, int(find_any), out.ctypes.data))
        return out

    def intersect_device(self, d_rays, n, d_out, instance=-1, find_any=False, d_counters=None):
        _check(self.lib.ygl_intersect_rays_device(self.ctx.h, self.h, self.bvh.h, d_rays, n,
                                                  instance, int(find_any), d_out, d_counters))

    def __del__(self):
        try:
            self.lib.ygl_scene_destroy(self.h)
        except Exception:
            pass


class State:
    """trace_state: resumable per-pixel accumulators + rng streams, resident on the device."""

    def __init__(self, dscene, params, rows=None, interleave=None):
        self.lib, self.ctx = dscene.lib, dscene.ctx
        self.h = C.c_void_p()
        if interleave is not None:
            _check(self.lib.ygl_state_create_interleaved(self.ctx.h, C.byref(dscene.desc),
                                                         C.byref(params), interleave[0],
                                                         interleave[1], C.byref(self.h)))
        elif rows is None:
            _check(self.lib.ygl_state_create(self.ctx.h, C.byref(dscene.desc), C.byref(params),
                                             C.byref(self.h)))
        else:
            _check(self.lib.ygl_state_create_tile(self.ctx.h, C.byref(dscene.desc),
                                                  C.byref(params), rows[0], rows[1],
                                                  C.byref(self.h)))
        w, h, s = C.c_int(), C.c_int(), C.c_int()
        _check(self.lib.ygl_state_size(self.h, C.byref(w), C.byref(h), C.byref(s)))
        rb, re = C.c_int(), C.c_int()
        _check(self.lib.ygl_state_rows(self.h, C.byref(rb), C.byref(re)))
        self.width, self.height, self.rows = w.value, h.value, (rb.value, re.value)
        first, step, nrows = C.c_int(), C.c_int(), C.c_int()
        _check(self.lib.ygl_state_layout(self.h, C.byref(first), C.byref(step), C.byref(nrows)))
        self.row_first, self.row_step, self.num_rows = first.value, step.value, nrows.value

    @property
    def samples(self):
        s = C.c_int()
        _check(self.lib.ygl_state_size(self.h, None, None, C.byref(s)))
        return s.value

    def download(self, full=False):
        nrows = self.num_rows
        image = np.zeros((nrows, self.width, 4), np.float32)
        if not full:
            _check(self.lib.ygl_state_download(self.h, _p(image), None, None, None, None))
            return dict(image=image)
        albedo = np.zeros((nrows, self.width, 3), np.float32)
        normal = np.zeros((nrows, self.width, 3), np.float32)
        hits = np.zeros((nrows, self.width), np.int32)
        rngs = np.zeros((nrows * self.width, 2), np.uint64)
        _check(self.lib.ygl_state_download(self.h, _p(image), _p(albedo), _p(normal), _p(hits),
                                           _p(rngs)))
        return dict(image=image, albedo=albedo, normal=normal, hits=hits, rngs=rngs)

    def tonemap(self, exposure=0.0, filmic=False, srgb=True):
        """tonemap_image of the state's device image: (vec4f image, vec4b image)."""
        ldr = np.zeros((self.num_rows, self.width, 4), np.float32)
        ldr_b = np.zeros((self.num_rows, self.width, 4), np.uint8)
        _check(self.lib.ygl_state_tonemap(self.h, exposure, int(filmic), int(srgb), _p(ldr), _p(ldr_b)))
        return ldr, ldr_b

    def reset(self, params):
        """reset_cutrace_state-style: zero accumulators, samples = 0, rng streams re-seeded from params.seed."""
        _check(self.lib.ygl_state_reset(self.h, C.byref(params)))

    def upload(self, samples, image=None, albedo=None, normal=None, hits=None, rngs=None):
        _check(self.lib.ygl_state_upload(self.h, samples, _p(image), _p(albedo), _p(normal),
                                         _p(hits), _p(rngs)))

    def gather_image(self):
        """ncclAllGather of the tiles (identity on one rank) -> full host image."""
        image = np.zeros((self.height, self.width, 4), np.float32)
        _check(self.lib.ygl_gather_image(self.ctx.h, self.h, image.ctypes.data))
        return image

    def __del__(self):
        try:
            self.lib.ygl_state_destroy(self.h)
        except Exception:
            pass


def tile_rows(height, rank, nranks):
    rb, re = C.c_int(), C.c_int()
    load().ygl_tile_rows(height, rank, nranks, C.byref(rb), C.byref(re))
    return rb.value, re.value
