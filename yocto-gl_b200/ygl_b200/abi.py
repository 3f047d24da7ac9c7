"""ctypes mirror of include/ygl_b200.h plus a small host-side scene container.

The container (`Scene`) plays the role of the reference's in-memory `scene_data`
(yocto_scene.h:171-195): plain arrays owned by the caller. `Scene.desc()` produces the
`ygl_scene_desc` flat views that both the CUDA path (libygl_b200.so) and the test oracle
(oracle/_ref/libyocto_ref.so) consume, so both sides see bit-identical inputs.
"""
import ctypes as C

import numpy as np

INVALID_ID = -1

# material_type (yocto_scene.h:107-112)
MATTE, GLOSSY, REFLECTIVE, TRANSPARENT, REFRACTIVE, SUBSURFACE, VOLUMETRIC, GLTFPBR = range(8)
# trace_sampler_type (yocto_trace.h:71-81)
(SAMPLER_PATH, SAMPLER_PATHDIRECT, SAMPLER_PATHMIS, SAMPLER_PATHTEST, SAMPLER_NAIVE,
 SAMPLER_EYELIGHT, SAMPLER_DIAGRAM, SAMPLER_FURNACE, SAMPLER_FALSECOLOR) = range(9)
# trace_falsecolor_type (yocto_trace.h:83-89)
(FC_POSITION, FC_NORMAL, FC_FRONTFACING, FC_GNORMAL, FC_GFRONTFACING, FC_TEXCOORD, FC_MTYPE,
 FC_COLOR, FC_EMISSION, FC_ROUGHNESS, FC_OPACITY, FC_METALLIC, FC_DELTA, FC_INSTANCE, FC_SHAPE,
 FC_MATERIAL, FC_ELEMENT, FC_HIGHLIGHT) = range(18)
DEFAULT_SEED = 961748941


class Frame(C.Structure):
    _fields_ = [("x", C.c_float * 3), ("y", C.c_float * 3), ("z", C.c_float * 3),
                ("o", C.c_float * 3)]


class Camera(C.Structure):
    _fields_ = [("frame", Frame), ("orthographic", C.c_int32), ("lens", C.c_float),
                ("film", C.c_float), ("aspect", C.c_float), ("focus", C.c_float),
                ("aperture", C.c_float)]


class Material(C.Structure):
    _fields_ = [("type", C.c_int32), ("emission", C.c_float * 3), ("color", C.c_float * 3),
                ("roughness", C.c_float), ("metallic", C.c_float), ("ior", C.c_float),
                ("scattering", C.c_float * 3), ("scanisotropy", C.c_float),
                ("trdepth", C.c_float), ("opacity", C.c_float), ("emission_tex", C.c_int32),
                ("color_tex", C.c_int32), ("roughness_tex", C.c_int32),
                ("scattering_tex", C.c_int32), ("normal_tex", C.c_int32)]


class Instance(C.Structure):
    _fields_ = [("frame", Frame), ("shape", C.c_int32), ("material", C.c_int32)]


class Environment(C.Structure):
    _fields_ = [("frame", Frame), ("emission", C.c_float * 3), ("emission_tex", C.c_int32)]


class Texture(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("linear", C.c_int32),
                ("nearest", C.c_int32), ("clamp", C.c_int32), ("pixelsf", C.c_void_p),
                ("pixelsb", C.c_void_p)]


class Shape(C.Structure):
    _fields_ = [("num_points", C.c_int32), ("num_lines", C.c_int32),
                ("num_triangles", C.c_int32), ("num_quads", C.c_int32),
                ("points", C.c_void_p), ("lines", C.c_void_p), ("triangles", C.c_void_p),
                ("quads", C.c_void_p),
                ("num_positions", C.c_int32), ("num_normals", C.c_int32),
                ("num_texcoords", C.c_int32), ("num_colors", C.c_int32),
                ("num_radius", C.c_int32),
                ("positions", C.c_void_p), ("normals", C.c_void_p), ("texcoords", C.c_void_p),
                ("colors", C.c_void_p), ("radius", C.c_void_p)]


class SceneDesc(C.Structure):
    _fields_ = [("num_cameras", C.c_int32), ("num_instances", C.c_int32),
                ("num_environments", C.c_int32), ("num_shapes", C.c_int32),
                ("num_textures", C.c_int32), ("num_materials", C.c_int32),
                ("cameras", C.POINTER(Camera)), ("instances", C.POINTER(Instance)),
                ("environments", C.POINTER(Environment)), ("shapes", C.POINTER(Shape)),
                ("textures", C.POINTER(Texture)), ("materials", C.POINTER(Material))]


class TraceParams(C.Structure):
    _fields_ = [("camera", C.c_int32), ("resolution", C.c_int32), ("sampler", C.c_int32),
                ("falsecolor", C.c_int32), ("samples", C.c_int32), ("bounces", C.c_int32),
                ("clamp", C.c_float), ("nocaustics", C.c_int32), ("envhidden", C.c_int32),
                ("tentfilter", C.c_int32), ("seed", C.c_uint64), ("embreebvh", C.c_int32),
                ("highqualitybvh", C.c_int32), ("noparallel", C.c_int32), ("pratio", C.c_int32),
                ("denoise", C.c_int32), ("batch", C.c_int32)]


def trace_params(**kw):
    """trace_params with the reference defaults (yocto_trace.h:95-113)."""
    p = TraceParams(camera=0, resolution=1280, sampler=SAMPLER_PATH, falsecolor=FC_COLOR,
                    samples=512, bounces=8, clamp=10.0, nocaustics=0, envhidden=0, tentfilter=0,
                    seed=DEFAULT_SEED, embreebvh=0, highqualitybvh=0, noparallel=0, pratio=8,
                    denoise=0, batch=1)
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


class Ray(C.Structure):
    _fields_ = [("o", C.c_float * 3), ("d", C.c_float * 3), ("tmin", C.c_float),
                ("tmax", C.c_float)]


class Intersection(C.Structure):
    _fields_ = [("instance", C.c_int32), ("element", C.c_int32), ("uv", C.c_float * 2),
                ("distance", C.c_float), ("hit", C.c_int32)]


class BvhNode(C.Structure):
    _fields_ = [("bbox_min", C.c_float * 3), ("bbox_max", C.c_float * 3), ("start", C.c_int32),
                ("num", C.c_int16), ("axis", C.c_int8), ("internal", C.c_uint8)]


RAY_DTYPE = np.dtype([("o", "<f4", 3), ("d", "<f4", 3), ("tmin", "<f4"), ("tmax", "<f4")])
ISEC_DTYPE = np.dtype([("instance", "<i4"), ("element", "<i4"), ("uv", "<f4", 2),
                       ("distance", "<f4"), ("hit", "<i4")])
NODE_DTYPE = np.dtype([("bbox_min", "<f4", 3), ("bbox_max", "<f4", 3), ("start", "<i4"),
                       ("num", "<i2"), ("axis", "i1"), ("internal", "u1")])
assert RAY_DTYPE.itemsize == 32 and ISEC_DTYPE.itemsize == 24 and NODE_DTYPE.itemsize == 32
assert C.sizeof(Camera) == 72 and C.sizeof(Material) == 84 and C.sizeof(Instance) == 56
assert C.sizeof(Environment) == 64 and C.sizeof(BvhNode) == 32

IDENTITY_FRAME = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [0, 0, 0]], np.float32)


def _frame(a):
    a = np.asarray(a, np.float32).reshape(4, 3)
    f = Frame()
    for name, row in zip("xyzo", a):
        setattr(f, name, (C.c_float * 3)(*[float(v) for v in row]))
    return f


def _ptr(a):
    return a.ctypes.data if a is not None and a.size else None


class Scene:
    """Host scene: lists of plain dicts / numpy arrays (the caller-owned scene_data)."""

    def __init__(self):
        self.cameras, self.instances, self.environments = [], [], []
        self.shapes, self.textures, self.materials = [], [], []
        self._keep = None

    # -- builders -----------------------------------------------------------
    def add_camera(self, frame=IDENTITY_FRAME, orthographic=False, lens=0.050, film=0.036,
                   aspect=1.5, focus=10000.0, aperture=0.0):
        self.cameras.append(dict(frame=np.asarray(frame, np.float32), orthographic=orthographic,
                                 lens=lens, film=film, aspect=aspect, focus=focus,
                                 aperture=aperture))
        return len(self.cameras) - 1

    def add_material(self, type=MATTE, emission=(0, 0, 0), color=(0, 0, 0), roughness=0.0,
                     metallic=0.0, ior=1.5, scattering=(0, 0, 0), scanisotropy=0.0,
                     trdepth=0.01, opacity=1.0, emission_tex=-1, color_tex=-1,
                     roughness_tex=-1, scattering_tex=-1, normal_tex=-1):
        self.materials.append(dict(type=type, emission=emission, color=color,
                                   roughness=roughness, metallic=metallic, ior=ior,
                                   scattering=scattering, scanisotropy=scanisotropy,
                                   trdepth=trdepth, opacity=opacity, emission_tex=emission_tex,
                                   color_tex=color_tex, roughness_tex=roughness_tex,
                                   scattering_tex=scattering_tex, normal_tex=normal_tex))
        return len(self.materials) - 1

    def add_shape(self, points=None, lines=None, triangles=None, quads=None, positions=None,
                  normals=None, texcoords=None, colors=None, radius=None):
        def arr(a, dt, width):
            if a is None:
                return np.zeros((0, width) if width > 1 else (0,), dt)
            a = np.ascontiguousarray(a, dt)
            return a.reshape(-1, width) if width > 1 else a.reshape(-1)
        self.shapes.append(dict(
            points=arr(points, np.int32, 1), lines=arr(lines, np.int32, 2),
            triangles=arr(triangles, np.int32, 3), quads=arr(quads, np.int32, 4),
            positions=arr(positions, np.float32, 3), normals=arr(normals, np.float32, 3),
            texcoords=arr(texcoords, np.float32, 2), colors=arr(colors, np.float32, 4),
            radius=arr(radius, np.float32, 1)))
        return len(self.shapes) - 1

    def add_instance(self, shape, material, frame=IDENTITY_FRAME):
        self.instances.append(dict(frame=np.asarray(frame, np.float32), shape=shape,
                                   material=material))
        return len(self.instances) - 1

    def add_environment(self, emission=(0, 0, 0), emission_tex=-1, frame=IDENTITY_FRAME):
        self.environments.append(dict(frame=np.asarray(frame, np.float32), emission=emission,
                                      emission_tex=emission_tex))
        return len(self.environments) - 1

    def add_texture(self, pixels, linear=None, nearest=False, clamp=False):
        """pixels: (h, w, 4) float32 (linear HDR) or uint8 (sRGB LDR by default)."""
        pixels = np.ascontiguousarray(pixels)
        assert pixels.ndim == 3 and pixels.shape[2] == 4
        if pixels.dtype == np.uint8:
            linear = False if linear is None else linear
        else:
            pixels = pixels.astype(np.float32)
            linear = True if linear is None else linear
        self.textures.append(dict(pixels=pixels, linear=linear, nearest=nearest, clamp=clamp))
        return len(self.textures) - 1

    # -- flat views ------------------------------------------------------------
    def desc(self):
        cams = (Camera * max(1, len(self.cameras)))()
        for i, c in enumerate(self.cameras):
            cams[i] = Camera(_frame(c["frame"]), int(c["orthographic"]), c["lens"], c["film"],
                             c["aspect"], c["focus"], c["aperture"])
        insts = (Instance * max(1, len(self.instances)))()
        for i, n in enumerate(self.instances):
            insts[i] = Instance(_frame(n["frame"]), n["shape"], n["material"])
        envs = (Environment * max(1, len(self.environments)))()
        for i, e in enumerate(self.environments):
            envs[i] = Environment(_frame(e["frame"]),
                                  (C.c_float * 3)(*[float(v) for v in e["emission"]]),
                                  e["emission_tex"])
        mats = (Material * max(1, len(self.materials)))()
        for i, m in enumerate(self.materials):
            f3 = lambda v: (C.c_float * 3)(*[float(x) for x in v])
            mats[i] = Material(m["type"], f3(m["emission"]), f3(m["color"]), m["roughness"],
                               m["metallic"], m["ior"], f3(m["scattering"]), m["scanisotropy"],
                               m["trdepth"], m["opacity"], m["emission_tex"], m["color_tex"],
                               m["roughness_tex"], m["scattering_tex"], m["normal_tex"])
        texs = (Texture * max(1, len(self.textures)))()
        for i, t in enumerate(self.textures):
            px = t["pixels"]
            isf = px.dtype != np.uint8
            texs[i] = Texture(px.shape[1], px.shape[0], int(t["linear"]), int(t["nearest"]),
                              int(t["clamp"]), _ptr(px) if isf else None,
                              None if isf else _ptr(px))
        shps = (Shape * max(1, len(self.shapes)))()
        for i, s in enumerate(self.shapes):
            shps[i] = Shape(len(s["points"]), len(s["lines"]), len(s["triangles"]),
                            len(s["quads"]), _ptr(s["points"]), _ptr(s["lines"]),
                            _ptr(s["triangles"]), _ptr(s["quads"]), len(s["positions"]),
                            len(s["normals"]), len(s["texcoords"]), len(s["colors"]),
                            len(s["radius"]), _ptr(s["positions"]), _ptr(s["normals"]),
                            _ptr(s["texcoords"]), _ptr(s["colors"]), _ptr(s["radius"]))
        d = SceneDesc(len(self.cameras), len(self.instances), len(self.environments),
                      len(self.shapes), len(self.textures), len(self.materials), cams, insts,
                      envs, shps, texs, mats)
        self._keep = (cams, insts, envs, mats, texs, shps)
        return d

    # -- (de)serialisation for tests/golden ---------------------------------------
    def to_npz_dict(self):
        out = {}
        for kind in ("cameras", "instances", "environments", "materials"):
            items = getattr(self, kind)
            out[f"n_{kind}"] = np.int32(len(items))
            for i, it in enumerate(items):
                for k, v in it.items():
                    out[f"{kind}.{i}.{k}"] = np.asarray(v)
        out["n_shapes"] = np.int32(len(self.shapes))
        for i, s in enumerate(self.shapes):
            for k, v in s.items():
                if v.size:
                    out[f"shapes.{i}.{k}"] = v
        out["n_textures"] = np.int32(len(self.textures))
        for i, t in enumerate(self.textures):
            for k, v in t.items():
                out[f"textures.{i}.{k}"] = np.asarray(v)
        return out

    @staticmethod
    def from_npz_dict(d):
        sc = Scene()

        def items(kind):
            res = []
            for i in range(int(d[f"n_{kind}"])):
                prefix = f"{kind}.{i}."
                res.append({k[len(prefix):]: d[k] for k in d if k.startswith(prefix)})
            return res

        def py(v):
            v = np.asarray(v)
            return v.item() if v.ndim == 0 else v

        for c in items("cameras"):
            sc.add_camera(**{k: py(v) for k, v in c.items()})
        for m in items("materials"):
            sc.add_material(**{k: py(v) for k, v in m.items()})
        for s in items("shapes"):
            sc.add_shape(**s)
        for t in items("textures"):
            sc.add_texture(t["pixels"], linear=bool(t["linear"]), nearest=bool(t["nearest"]),
                           clamp=bool(t["clamp"]))
        for n in items("instances"):
            sc.add_instance(int(n["shape"]), int(n["material"]), n["frame"])
        for e in items("environments"):
            sc.add_environment(py(e["emission"]), int(e["emission_tex"]), e["frame"])
        return sc

    @staticmethod
    def from_desc(desc):
        """Deep-copies a ygl_scene_desc (e.g. one exported by the oracle) into a Scene."""
        sc = Scene()

        def fr(f):
            return np.array([list(f.x), list(f.y), list(f.z), list(f.o)], np.float32)

        def view(ptr, n, dt, width):
            if not ptr or n == 0:
                return None
            count = n * width
            buf = (C.c_byte * (count * np.dtype(dt).itemsize)).from_address(ptr)
            a = np.frombuffer(buf, dt, count).copy()
            return a.reshape(n, width) if width > 1 else a

        for i in range(desc.num_cameras):
            c = desc.cameras[i]
            sc.add_camera(fr(c.frame), bool(c.orthographic), c.lens, c.film, c.aspect, c.focus,
                          c.aperture)
        for i in range(desc.num_materials):
            m = desc.materials[i]
            sc.add_material(m.type, tuple(m.emission), tuple(m.color), m.roughness, m.metallic,
                            m.ior, tuple(m.scattering), m.scanisotropy, m.trdepth, m.opacity,
                            m.emission_tex, m.color_tex, m.roughness_tex, m.scattering_tex,
                            m.normal_tex)
        for i in range(desc.num_shapes):
            s = desc.shapes[i]
            sc.add_shape(view(s.points, s.num_points, np.int32, 1),
                         view(s.lines, s.num_lines, np.int32, 2),
                         view(s.triangles, s.num_triangles, np.int32, 3),
                         view(s.quads, s.num_quads, np.int32, 4),
                         view(s.positions, s.num_positions, np.float32, 3),
                         view(s.normals, s.num_normals, np.float32, 3),
                         view(s.texcoords, s.num_texcoords, np.float32, 2),
                         view(s.colors, s.num_colors, np.float32, 4),
                         view(s.radius, s.num_radius, np.float32, 1))
        for i in range(desc.num_textures):
            t = desc.textures[i]
            if t.pixelsf:
                px = view(t.pixelsf, t.width * t.height, np.float32, 4).reshape(t.height, t.width, 4)
            else:
                px = view(t.pixelsb, t.width * t.height, np.uint8, 4).reshape(t.height, t.width, 4)
            sc.add_texture(px, bool(t.linear), bool(t.nearest), bool(t.clamp))
        for i in range(desc.num_instances):
            n = desc.instances[i]
            sc.add_instance(n.shape, n.material, fr(n.frame))
        for i in range(desc.num_environments):
            e = desc.environments[i]
            sc.add_environment(tuple(e.emission), e.emission_tex, fr(e.frame))
        return sc
