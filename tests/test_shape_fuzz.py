"""Differential tests of the shape readers (PLY and OBJ in ygl_sceneio.cpp) against the reference's load_shape
(yocto_sceneio.cpp:1008-1051 over yocto_modelio's load_ply / load_obj): seeded random files - PLY in ascii and both binary
byte orders, properties of every scalar type in any order, list properties with any count / index type, extra elements and
properties, polygons of 0..7 corners, polylines, points, radius / colour / texcoord variants; OBJ with shared and split
vertices, negative indices, missing components, polygons, lines and points, comments - must load to bit-identical arrays
(or be refused by both). Host-only; the files are read in a worker process (tests/loader_worker.py) so that a crash of
the reference is survived."""
import json
import os
import random
import struct

import pytest

from loader_worker import LoaderPair

SCALARS = {"char": "b", "uchar": "B", "short": "h", "ushort": "H", "int": "i", "uint": "I", "float": "f", "double": "d",
           "int8": "b", "uint8": "B", "int16": "h", "uint16": "H", "int32": "i", "uint32": "I", "float32": "f", "float64": "d"}


def make_ply(rng):
    fmt = rng.choice(["ascii", "binary_little_endian", "binary_big_endian"])
    nverts = rng.randint(3, 12)
    groups = [["x", "y", "z"]]
    for extra in (["nx", "ny", "nz"], rng.choice([["u", "v"], ["s", "t"]]), rng.choice([["red", "green", "blue"], ["red", "green", "blue", "alpha"]]),
                  ["radius"], ["quality"]):
        if rng.random() < 0.5:
            groups.append(extra)
    if rng.random() < 0.15:
        groups[0] = ["x", "y"]                    # an incomplete triple: the reference then reads no positions at all
    props = [p for g in groups for p in g]
    if rng.random() < 0.5:
        rng.shuffle(props)
    elements = []   # (name, count, [(prop name, type or (count type, item type), values per row)])

    def scalar_type(name):
        if name in ("red", "green", "blue", "alpha"):
            return rng.choice(["uchar", "uint8", "float", "ushort"])
        return rng.choice(["float", "float32", "double", "float64"] if rng.random() < 0.9 else ["int", "short"])
    vertex = []
    for p in props:
        t = scalar_type(p)
        code = SCALARS[t]
        if code in "fd":
            vals = [rng.choice([0.0, 1.0, -1.0, 0.5]) if rng.random() < 0.2 else round(rng.uniform(-2, 2), rng.randint(1, 7)) for _ in range(nverts)]
        else:
            hi = {"b": 127, "B": 255, "h": 32767, "H": 65535, "i": 100000, "I": 100000}[code]
            vals = [rng.randint(0 if code.isupper() else -hi, hi) for _ in range(nverts)]
        vertex.append((p, t, vals))
    elements.append(("vertex", nverts, vertex))

    def index_list(name, sizes_choices, count):
        ctype = rng.choice(["uchar", "uint8"]) if rng.random() < 0.97 else rng.choice(["int", "ushort"])   # only bytes are accepted
        itype = rng.choice(["int", "uint", "int32", "ushort", "uchar", "short"])
        rows = [[rng.randrange(nverts) for _ in range(rng.choice(sizes_choices))] for _ in range(count)]
        return (name, (ctype, itype), rows)
    kinds = rng.sample(["face", "line", "point"], rng.randint(1, 3))
    if rng.random() < 0.3:
        elements.append(("material", 2, [("ambient", "float", [0.5, 0.25])]))       # an element nobody reads
    for kind in kinds:
        n = rng.randint(1, 6)
        if kind == "face":
            shapes = rng.choice([[3], [4], [3, 4], [3, 4, 5, 6, 7], [0, 1, 2, 3], [3, 3, 3, 4]])
            cols = [index_list(rng.choice(["vertex_indices", "vertex_indices", "vertex_index"]), shapes, n)]
            if rng.random() < 0.3:
                cols.append(("flags", "uchar", [rng.randrange(4) for _ in range(n)]))
                rng.shuffle(cols)
        elif kind == "line":
            cols = [index_list("vertex_indices", [2, 2, 3, 5, 1, 0], n)]
        else:
            cols = [index_list("vertex_indices", [1, 1, 2, 0], n)]
        elements.append((kind, n, cols))
    if rng.random() < 0.3:
        rng.shuffle(elements)
    nl = b"\n"
    head = [b"ply", b"format " + fmt.encode() + b" 1.0"]
    if rng.random() < 0.5:
        head.append(b"comment made by a test")
    if rng.random() < 0.2:
        head.append(b"obj_info whatever 1 2 3")
    for name, count, cols in elements:
        head.append(b"element %s %d" % (name.encode(), count))
        for pname, ptype, _ in cols:
            if isinstance(ptype, tuple):
                head.append(b"property list %s %s %s" % (ptype[0].encode(), ptype[1].encode(), pname.encode()))
            else:
                head.append(b"property %s %s" % (ptype.encode(), pname.encode()))
        if rng.random() < 0.1:
            head.append(b"comment in between")
    head.append(b"end_header")
    out = nl.join(head) + nl
    order = {"ascii": None, "binary_little_endian": "<", "binary_big_endian": ">"}[fmt]
    for name, count, cols in elements:
        for r in range(count):
            if order is None:
                words = []
                for pname, ptype, vals in cols:
                    if isinstance(ptype, tuple):
                        words += [str(len(vals[r]))] + [str(v) for v in vals[r]]
                    else:
                        words.append(repr(vals[r]) if isinstance(vals[r], float) else str(vals[r]))
                out += (" " * rng.randint(0, 1) + (" " * rng.randint(1, 2)).join(words)).encode() + nl
            else:
                for pname, ptype, vals in cols:
                    if isinstance(ptype, tuple):
                        out += struct.pack(order + SCALARS[ptype[0]], len(vals[r]))
                        out += b"".join(struct.pack(order + SCALARS[ptype[1]], v) for v in vals[r])
                    else:
                        out += struct.pack(order + SCALARS[ptype], vals[r])
    return out


def make_obj(rng):
    nv, nn, nt = rng.randint(3, 9), rng.choice([0, 0, 2, 4]), rng.choice([0, 0, 3, 5])
    lines = ["# a test file"] if rng.random() < 0.5 else []
    number = lambda: rng.choice(["0", "1", "-1", "0.5", "1e-3", "-2.5E1", ".25", "3."]) if rng.random() < 0.3 else repr(round(rng.uniform(-3, 3), rng.randint(0, 6)))
    body = [f"v {number()} {number()} {number()}" for _ in range(nv)]
    body += [f"vn {number()} {number()} {number()}" for _ in range(nn)]
    body += [f"vt {number()} {number()}" for _ in range(nt)]

    def vert():
        i = rng.randrange(nv)
        v = str(i + 1) if rng.random() < 0.8 else str(i - nv)
        t = (str(rng.randrange(nt) + 1) if rng.random() < 0.8 else str(-1 - rng.randrange(nt))) if nt and rng.random() < 0.8 else ""
        n = (str(rng.randrange(nn) + 1) if rng.random() < 0.8 else str(-1 - rng.randrange(nn))) if nn and rng.random() < 0.8 else ""
        return v + ("/" + t + ("/" + n if n else "") if t or n else "")
    elems = []
    kinds = rng.choice([["f"], ["f"], ["f", "l"], ["l"], ["p"], ["f", "l", "p"], ["l", "p"]])
    for _ in range(rng.randint(1, 7)):
        kind = rng.choice(kinds)
        n = {"f": rng.choice([3, 3, 4, 4, 5, 6]), "l": rng.choice([2, 3, 4]), "p": rng.choice([1, 1, 2])}[kind]
        elems.append(kind + " " + " ".join(vert() for _ in range(n)))
    if rng.random() < 0.3:
        elems.insert(rng.randrange(len(elems) + 1), "g part" if rng.random() < 0.5 else "s 1")
    if rng.random() < 0.3:
        elems.insert(rng.randrange(len(elems) + 1), "usemtl stuff")
    if rng.random() < 0.3:      # vertices declared after their use by negative index resolve differently: keep order, but interleave some
        rng.shuffle(body)
    text = "\n".join(lines + body + elems) + ("\n" if rng.random() < 0.9 else "")
    if rng.random() < 0.2:
        text = text.replace("\n", "\r\n")
    return text.encode()


def _run(loaders, tmp_path, name, data):
    os.makedirs(tmp_path / "shapes", exist_ok=True)
    (tmp_path / "shapes" / name).write_bytes(data)
    scene = tmp_path / (name + ".json")
    scene.write_text(json.dumps({"asset": {"version": "4.2"}, "cameras": [{"name": "c"}], "shapes": [{"name": "s", "uri": "shapes/" + name}]}))
    return loaders.verdict(scene)


@pytest.mark.parametrize("seed", [1, 2])
def test_random_ply_files_load_like_the_reference(ref, seed, tmp_path):
    rng = random.Random(seed)
    counts = {"same": 0, "refused": 0, "reference crashed": 0}
    loaders = LoaderPair()
    for k in range(150):
        data = make_ply(rng)
        verdict = _run(loaders, tmp_path, f"s{k}.ply", data)
        assert verdict in counts, f"file {k} (seed {seed}): {verdict}\n{data[:1200]!r}"
        counts[verdict] += 1
    loaders.close()
    assert counts["same"] >= 100, counts


@pytest.mark.parametrize("seed", [1, 2])
def test_random_obj_files_load_like_the_reference(ref, seed, tmp_path):
    rng = random.Random(seed)
    counts = {"same": 0, "refused": 0, "reference crashed": 0}
    loaders = LoaderPair()
    for k in range(150):
        data = make_obj(rng)
        verdict = _run(loaders, tmp_path, f"s{k}.obj", data)
        assert verdict in counts, f"file {k} (seed {seed}): {verdict}\n{data.decode()}"
        counts[verdict] += 1
    loaders.close()
    assert counts["same"] >= 100, counts


def make_subdiv_scene(rng, d):
    """a quad-grid control mesh with holes, triangles, degenerate pentagons, stray texture coordinates and normals (or a
    closed cube), subdivided 0-3 times with either scheme, smooth or faceted, displaced or not"""
    nx, ny = rng.randint(1, 4), rng.randint(1, 4)
    verts = [(i + rng.uniform(-.2, .2), j + rng.uniform(-.2, .2), rng.uniform(-.5, .5)) for j in range(ny + 1) for i in range(nx + 1)]
    faces = []
    for j in range(ny):
        for i in range(nx):
            if rng.random() < 0.15:
                continue
            a = j * (nx + 1) + i
            q = [a, a + 1, a + nx + 2, a + nx + 1]
            r = rng.random()
            faces += [[q[0], q[1], q[2]]] if r < 0.15 else [[q[0], q[1], q[2]], [q[0], q[2], q[3]]] if r < 0.2 else [q + [a]] if r < 0.25 else [q]
    faces = faces or [[0, 1, nx + 2, nx + 1]]
    if rng.random() < 0.2:
        verts = [(-1, -1, -1), (1, -1, -1), (1, 1, -1), (-1, 1, -1), (-1, -1, 1), (1, -1, 1), (1, 1, 1), (-1, 1, 1)]
        faces = [[0, 3, 2, 1], [4, 5, 6, 7], [0, 1, 5, 4], [2, 3, 7, 6], [1, 2, 6, 5], [0, 4, 7, 3]]
    has_t, has_n = rng.random() < 0.7, rng.random() < 0.4
    lines = [f"v {x:.5f} {y:.5f} {z:.5f}" for x, y, z in verts]
    ntex = (len(verts) + (3 if rng.random() < 0.5 else 0)) if has_t else 0
    lines += [f"vt {rng.random():.4f} {rng.random():.4f}" for _ in range(ntex)]
    lines += ["vn 0 0 1", "vn 0 1 0", "vn 1 0 0"] if has_n else []
    for f in faces:
        # the three topologies must agree on which faces are triangles (last two corners equal): what the reference does
        # otherwise is undefined (it reads past the shorter per-face array) and we refuse such a mesh
        tex = [v + 1 if rng.random() < 0.8 else rng.randint(1, ntex) for v in f] if has_t else []
        nrm = [rng.randint(1, 3) for _ in f] if has_n else []
        if len(f) == 4:
            if has_t and tex[2] == tex[3]:
                tex[2], tex[3] = f[2] + 1, f[3] + 1
            if has_n and nrm[2] == nrm[3]:
                nrm[3] = nrm[2] % 3 + 1
        corners = []
        for c, v in enumerate(f):
            t, n = (str(tex[c]) if has_t else ""), (str(nrm[c]) if has_n else "")
            corners.append(str(v + 1) + ("/" + t + ("/" + n if n else "") if t or n else ""))
        lines.append("f " + " ".join(corners))
    (d / "subdivs" / "s.obj").write_text("\n".join(lines) + "\n")
    subdiv = {"name": "s", "uri": "subdivs/s.obj", "shape": 0, "subdivisions": rng.choice([0, 1, 2, 3]),
              "catmullclark": rng.random() < 0.7, "smooth": rng.random() < 0.6}
    if has_t and rng.random() < 0.6:
        subdiv.update(displacement=rng.choice([0.05, 0.2, -0.1]), displacement_tex=0)
    return {"asset": {"version": "4.2"}, "cameras": [{"name": "c"}], "textures": [{"name": "t", "uri": "textures/t.png"}],
            "shapes": [{"name": "x", "uri": "shapes/tri.ply"}], "subdivs": [subdiv], "instances": [{"shape": 0}]}


def test_random_subdivs_tesselate_like_the_reference(ref, tmp_path):
    """load_subdiv + tesselate_subdivs (yocto_scene.cpp:739-813): Catmull-Clark / linear subdivision, boundary rules,
    smoothing, displacement by a byte texture, split_facevarying - every vertex and index of the result."""
    import struct
    import zlib
    from test_sceneio import _write_tri_ply
    rng = random.Random(5)
    for sub in ("subdivs", "shapes", "textures"):
        os.makedirs(tmp_path / sub)
    _write_tri_ply(tmp_path / "shapes" / "tri.ply")
    chunk = lambda tag, body: struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body) & 0xffffffff)
    rows = b"".join(b"\0" + bytes(rng.randrange(256) for _ in range(8 * 4)) for _ in range(8))
    (tmp_path / "textures" / "t.png").write_bytes(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 8, 8, 8, 6, 0, 0, 0))
                                                  + chunk(b"IDAT", zlib.compress(rows)) + chunk(b"IEND", b""))
    counts = {"same": 0, "refused": 0, "reference crashed": 0}
    loaders = LoaderPair()
    for k in range(100):
        scene = tmp_path / "scene.json"
        scene.write_text(json.dumps(make_subdiv_scene(rng, tmp_path)))
        verdict = loaders.verdict(scene)
        assert verdict in counts, f"mesh {k}: {verdict}\n{(tmp_path / 'subdivs' / 's.obj').read_text()}\n{scene.read_text()}"
        counts[verdict] += 1
    loaders.close()
    assert counts["same"] >= 80, counts     # (the reference itself crashes on a few of the meshes with stray indices)


def test_stl_shapes_load_like_the_reference(ref, tmp_path):
    """Binary .stl shapes (load_stl, yocto_modelio.cpp:2164): vertices merged by float equality in order of first use
    (-0 joins +0, a NaN joins nothing), a header that says "solid" but whose length fits the binary layout, two solids in
    one file (refused), a truncated file (refused), ascii text (refused: the reference's ascii branch cannot pass an
    "outer loop" line)."""
    rng = random.Random(9)
    nan = struct.unpack("<f", b"\x00\x00\xc0\x7f")[0]

    def solid(ntri, pool):
        out = struct.pack("<I", ntri)
        for _ in range(ntri):
            out += struct.pack("<3f", 0, 0, 1) + b"".join(struct.pack("<3f", *rng.choice(pool)) for _ in range(3)) + struct.pack("<H", rng.randrange(4))
        return out
    pool = [(rng.choice([0.0, -0.0, 1.0, 0.5, -2.25, nan]), rng.uniform(-1, 1) if rng.random() < 0.5 else 0.0, rng.choice([0.0, -0.0, 3.0])) for _ in range(12)]
    files = {
        "plain.stl": (b"made by a test".ljust(80, b"\0") + solid(40, pool), "same"),
        "solid_header.stl": (b"solid but binary".ljust(80, b" ") + solid(7, pool), "same"),
        "empty_solid.stl": (b"x".ljust(80, b"\0") + solid(0, pool), "same"),
        "two_solids.stl": (b"x".ljust(80, b"\0") + solid(3, pool) + solid(2, pool), "refused"),
        "truncated.stl": ((b"x".ljust(80, b"\0") + solid(5, pool))[:-20], "refused"),
        "no_solid.stl": (b"x".ljust(80, b"\0"), "refused"),
        "ascii.stl": (b"solid s\nfacet normal 0 0 1\nouter loop\nvertex 0 0 0\nvertex 1 0 0\nvertex 0 1 0\nendloop\nendfacet\nendsolid s\n".ljust(120, b"\n"), "refused"),
    }
    loaders = LoaderPair()
    for name, (data, expected) in files.items():
        assert _run(loaders, tmp_path, name, data) == expected, name
    loaders.close()


def make_obj_scene(rng, d, stem):
    """an .obj with objects / groups, several materials from one or two .mtl files (all the colour, exponent, opacity and
    texture statements), elements before the first usemtl, repeated usemtl, and sometimes the .obx side file with cameras
    and environments"""
    import struct
    import zlib
    nmat = rng.randint(0, 4)
    names = [f"mat{k}" for k in range(nmat)]
    number = lambda: repr(round(rng.uniform(0, 1), rng.randint(1, 4)))
    triple = lambda: " ".join(number() for _ in range(3))
    chunk = lambda tag, body: struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body) & 0xffffffff)
    texture_names = []
    for k in range(rng.randint(0, 2)):
        name = f"{stem}_t{k}.png"
        rows = b"".join(b"\0" + bytes(rng.randrange(256) for _ in range(3 * 3)) for _ in range(2))
        (d / name).write_bytes(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 3, 2, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(rows)) + chunk(b"IEND", b""))
        texture_names.append(name)
    libs = [names[:len(names) // 2 + 1], names[len(names) // 2 + 1:]] if nmat > 2 and rng.random() < 0.4 else [names]
    for k, lib in enumerate(libs):
        lines = ["# materials"]
        if rng.random() < 0.2:
            lines.append("Kd 0.1 0.2 0.3")           # before the first newmtl: goes to the placeholder
        for name in lib:
            lines.append(f"newmtl {name}")
            for key in ("Ke", "Ka", "Kd", "Ks", "Kt", "Tf"):
                if rng.random() < 0.45:
                    lines.append(f"{key} " + (triple() if rng.random() < 0.8 else "0 0 0"))
            for key in ("Ns", "d", "Tr", "illum"):
                if rng.random() < 0.4:
                    lines.append(f"{key} " + (str(rng.choice([0, 1, 2, 10, 200, 1500])) if key in ("Ns", "illum") else number()))
            for key in ("map_Kd", "map_Ks", "map_Ke", "map_Tr", "map_d", "map_bump", "norm", "map_Ka"):
                if texture_names and rng.random() < 0.25:
                    option = "-bm 0.5 " if key == "map_bump" and rng.random() < 0.5 else ""
                    lines.append(f"{key} {option}{rng.choice(texture_names)}")
            if rng.random() < 0.2:
                lines.append("Ni 1.33")              # not read by the reference
        (d / f"{stem}_{k}.mtl").write_text("\n".join(lines) + "\n")
    nv = rng.randint(4, 12)
    out = [f"v {number()} {number()} {number()}" for _ in range(nv)]
    nn, nt = rng.choice([0, 3]), rng.choice([0, 4])
    out += [f"vn {number()} {number()} {number()}" for _ in range(nn)] + [f"vt {number()} {number()}" for _ in range(nt)]

    def vert():
        v = rng.randint(1, nv) if rng.random() < 0.85 else -rng.randint(1, nv)
        t = str(rng.randint(1, nt)) if nt and rng.random() < 0.7 else ""
        n = str(rng.randint(1, nn)) if nn and rng.random() < 0.7 else ""
        return str(v) + ("/" + t + ("/" + n if n else "") if t or n else "")
    body = []
    if rng.random() < 0.4:
        body.append("f " + " ".join(vert() for _ in range(3)))         # before any mtllib / usemtl: the grey default
    for k in range(len(libs)):
        body.append(f"mtllib {stem}_{k}.mtl")
    if rng.random() < 0.2 and libs:
        body.append(f"mtllib {stem}_0.mtl")                             # a repeated library is read once
    for _ in range(rng.randint(1, 10)):
        r = rng.random()
        if r < 0.2:
            body.append(rng.choice(["o", "g"]) + (" part%d" % rng.randint(0, 3) if rng.random() < 0.8 else ""))
        elif r < 0.45 and names:
            body.append("usemtl " + rng.choice(names))
        else:
            kind = rng.choice(["f", "f", "f", "l", "p"])
            n = {"f": rng.choice([3, 4, 5]), "l": rng.choice([2, 3]), "p": 1}[kind]
            body.append(kind + " " + " ".join(vert() for _ in range(n)))
    if not any(line[0] in "flp" for line in body):
        body.append("f 1 2 3")
    (d / f"{stem}.obj").write_text("\n".join(out + body) + "\n")
    if rng.random() < 0.4:
        side = []
        for k in range(rng.randint(0, 2)):
            side += [f"newCam cam{k}", f"Ca {rng.choice([1.0, 1.5, 2.4])}", f"Cl {number()}"]
            side += ["Ct 1 2 3 0 0.5 0 0 1 0"] if rng.random() < 0.6 else ["Cx 1 0 0 0 1 0 0 0 1 0.5 1 4", "Cf 2.5", "Co 1"]
        for k in range(rng.randint(0, 2)):
            side += [f"newEnv env{k}", f"Ee {triple()}"] + ([f"map_Ee {rng.choice(texture_names)}"] if texture_names and rng.random() < 0.5 else [])
            side += ["Et 0 0 0 0 0 -1 0 1 0"] if rng.random() < 0.5 else []
        (d / f"{stem}.obx").write_text("\n".join(side) + "\n")
    return d / f"{stem}.obj"


@pytest.mark.parametrize("seed", [1, 2])
def test_random_obj_scenes_load_like_the_reference(ref, seed, tmp_path):
    """`load_scene("x.obj")` (load_obj_scene, yocto_sceneio.cpp:4111): shapes split by object / group and material,
    .mtl materials classified and converted, textures in order of first mention, .obx cameras and environments."""
    rng = random.Random(seed)
    counts = {"same": 0, "refused": 0, "reference crashed": 0}
    loaders = LoaderPair()
    for k in range(120):
        path = make_obj_scene(rng, tmp_path, f"s{k}")
        verdict = loaders.verdict(path)
        assert verdict in counts, f"scene {k} (seed {seed}): {verdict}\n{path.read_text()}\n" + "\n".join(
            p.read_text() for p in sorted(tmp_path.glob(f"s{k}_*.mtl")))
        counts[verdict] += 1
    loaders.close()
    assert counts["same"] >= 80, counts
