"""Differential fuzzing of the scene-file reader (ygl_scene_load) against the reference's load_scene
(yocto_sceneio.cpp:2761): seeded random Yocto/GL JSON documents of format 4.2 / 5.0 and of format 4.0 - mostly valid,
with a few mutated values per document (wrong types, short and long arrays, booleans as numbers, the numbers nlohmann
reads as integers or rejects: "-0", 2^64, 1e400; groups given as the other container kind; repeated keys; \\u escapes,
raw UTF-8, ill-formed UTF-8, a BOM; truncated text) - must be refused by both loaders or give bit-identical scenes.
Every document is loaded in a worker process (tests/loader_worker.py): the reference reads out of bounds on some inputs it
accepts (an instance without a valid shape in a scene it has to frame), which the generator avoids, and a crash of the
reference must not take the test down. Host-only."""
import json
import os
import random

import pytest

from ygl_b200 import lib
from loader_worker import LoaderPair
from test_sceneio import _write_tri_ply

RAW_NUMBERS = ["1e400", "-1e400", "1E5", "-0", "-0.0", "0.1e-46", "123456789012345678901234567890", "9007199254740993",
               "18446744073709551615", "18446744073709551616", "-9223372036854775808", "-9223372036854775809", "1.0e+2", "0e0",
               "3.4028235e38", "3.4028236e38", "1e39", "+1", ".5", "1.", "01", "inf", "0x10"]
KEYS = {
    "cameras": ["name", "frame", "orthographic", "ortho", "lens", "aspect", "film", "focus", "aperture", "lookat"],
    "materials": ["name", "type", "emission", "color", "metallic", "roughness", "ior", "trdepth", "scattering", "scanisotropy",
                  "opacity", "emission_tex", "color_tex", "roughness_tex", "scattering_tex", "normal_tex"],
    "instances": ["name", "frame", "shape", "material", "lookat"],
    "objects": ["frame", "shape", "material", "lookat", "instance"],
    "environments": ["name", "frame", "emission", "emission_tex", "lookat"],
}
GOOD = {"frame": [1, 0, 0, 0, 1, 0, 0, 0, 1, 0.5, 1, 2], "lookat": [1, 2, 3, 0, 0, 0, 0, 1, 0], "lens": 0.05, "aspect": 1.5,
        "film": 0.036, "focus": 3, "aperture": 0.1, "orthographic": True, "ortho": False, "type": "glossy", "emission": [1, 2, 3],
        "color": [0.5, 0.25, 0.125], "metallic": 0.5, "roughness": 0.25, "ior": 1.3, "trdepth": 0.02, "scattering": [0.1, 0.2, 0.3],
        "scanisotropy": 0.1, "opacity": 0.5, "name": "n"}
WEIRD_STRINGS = ['"a\\u00e9b"', '"\\ud83d\\ude00"', '"\\ud83d"', '"\\udc00x"', '"tab\there"', '"caf\xc3\xa9"', '"bad\xe9"', '"\\x"',
                 '"q\\/\\b"', '"\xf0\x9f\x98\x80"', '"\xed\xa0\x80"', '"\xc0\xaf"']


class Raw(str):
    """a token written into the document as it is"""


def dumps(x):
    if isinstance(x, Raw):
        return str(x)
    if isinstance(x, dict):
        return "{" + ",".join(json.dumps(k) + ":" + dumps(v) for k, v in x.items()) + "}"
    if isinstance(x, list):
        return "[" + ",".join(dumps(v) for v in x) + "]"
    return json.dumps(x)


def make_document(rng):
    v40 = rng.random() < 0.5
    mutate = rng.choice([0, 0, 0.02, 0.05, 0.2])     # chance of a mutated value per key

    def any_value(depth=0):
        r = rng.random()
        if r < 0.25:
            return rng.choice([0, 1, -1, 2, 0.5, 1e-3, 3.75, 1e10, -2.5, 7])
        if r < 0.35:
            return rng.choice([True, False])
        if r < 0.45:
            return None
        if r < 0.55:
            return rng.choice(["", "matte", "glossy", "metallic", "tri", "x", "volume", "reflective"])
        if r < 0.60:
            return Raw(rng.choice(RAW_NUMBERS))
        if r < 0.90:
            return [any_value(1) if rng.random() < 0.15 else rng.choice([0, 1, 0.5, -1, 2.25, True])
                    for _ in range(rng.choice([0, 2, 3, 3, 4, 9, 9, 12, 12, 13]))]
        return {} if depth else {"a": 1}

    def element(group):
        e = {}
        for k in KEYS[group]:
            if rng.random() < 0.5:
                continue
            bad = rng.random() < mutate
            if not bad and k in GOOD:
                plain = rng.random() < 0.7 or isinstance(GOOD[k], (bool, list, str))
                e[k] = GOOD[k] if plain else rng.choice([0, 1, -1, 2.5, Raw(rng.choice(RAW_NUMBERS[2:6] + RAW_NUMBERS[12:14]))])
            elif k in ("shape", "material"):
                good = rng.choice(["tri", "m0", "m1", ""]) if v40 else rng.choice([0, 0, -1])
                e[k] = good if not bad else rng.choice([0, 1, -1, "tri", "m0", "", "zz", 0.7, True, None, [1]])
            elif k.endswith("_tex"):
                e[k] = ("" if v40 else -1) if not bad else rng.choice([-1, 0, "", 3, None, True, "tex"])
            elif k == "instance":
                e[k] = ""
            else:
                e[k] = any_value()
        return e

    doc = {}
    if v40:
        if rng.random() < 0.7:
            doc["asset"] = rng.choice([{"copyright": "c"}, {}, {"copyright": 5}, 3, None, [1]]) if mutate else {"copyright": "c"}
        for g in ("cameras", "materials", "instances", "objects", "environments"):
            if rng.random() < 0.6:
                n = rng.randint(0, 3)
                if rng.random() < (0.85 if mutate else 1):
                    doc[g] = {("m%d" % k if g == "materials" else "%s%d" % (g[0], k)): element(g) for k in range(n)}
                else:
                    doc[g] = rng.choice([[element(g) for _ in range(n)], None, 5, "s", [1, 2]])
    else:
        versions = ["4.2", "4.2", "4.2", "5.0", "4.0", "4.1", "", 4.2, None] if mutate else ["4.2", "5.0"]
        doc["asset"] = {"version": rng.choice(versions)}
        if rng.random() < 0.3:
            doc["asset"]["copyright"] = rng.choice(["c", 1, None])
        for g in ("cameras", "materials", "instances", "environments", "shapes"):
            if rng.random() < 0.6:
                n = rng.randint(0, 3)
                if g == "shapes":
                    doc[g] = [{"name": "s", "uri": "shapes/tri.ply"} for _ in range(n)]
                elif rng.random() < (0.9 if mutate else 1):
                    doc[g] = [element(g) for _ in range(n)]
                else:
                    doc[g] = rng.choice([{"a": element(g)}, None, 5, "s", [1, 2]])
    # keep the reference inside its defined behaviour: an instance without a valid shape needs a camera in the file
    nshapes = len(doc["shapes"]) if isinstance(doc.get("shapes"), list) else 0

    def has_shape(e):
        if not isinstance(e, dict):
            return True
        shape = e.get("shape", "" if v40 else -1)
        return shape == "tri" if v40 else (isinstance(shape, int) and not isinstance(shape, bool) and 0 <= shape < nshapes)
    members = []
    for g in ("instances", "objects"):
        group = doc.get(g)
        members += list(group.values()) if isinstance(group, dict) else group if isinstance(group, list) else []
    if not all(has_shape(e) for e in members) or any(isinstance(doc.get(g), (int, str)) for g in ("instances", "objects")):
        camera = {"frame": GOOD["frame"], "lens": 0.05}
        doc["cameras"] = {"c": camera} if v40 else [camera]
    text = dumps(doc)
    if rng.random() < 0.15:
        text = text.replace('"n"', rng.choice(WEIRD_STRINGS), 1)
    if rng.random() < 0.1:
        text = text.replace('"lens":', '"lens":7,"lens":', 1).replace('"m0":', '"m0":{"color":[9,9,9]},"m1":{},"m0":', 1)
    if rng.random() < mutate:
        text = text[:-1]
    if rng.random() < mutate:
        text += rng.choice([" ", "x", "\n\n", ",", "{}"])
    data = text.encode("latin-1")          # the weird strings hold raw bytes
    if rng.random() < 0.05:
        data = b"\xef\xbb\xbf" + data
    return data


@pytest.mark.parametrize("seed", [101, 202, 303])
def test_random_scene_documents_load_like_the_reference(ref, seed, tmp_path):
    rng = random.Random(seed)
    os.makedirs(tmp_path / "shapes")
    _write_tri_ply(tmp_path / "shapes" / "tri.ply")
    verdicts = {"same": 0, "refused": 0, "reference crashed": 0}
    loaders = LoaderPair()
    for k in range(150):
        path = tmp_path / f"doc{k}.json"
        data = make_document(rng)
        path.write_bytes(data)
        verdict = loaders.verdict(path)
        assert verdict in verdicts, f"document {k} (seed {seed}): {verdict}\n{data[:1500]!r}"
        verdicts[verdict] += 1
        if verdict == "reference crashed":       # then ours alone must still come back
            try:
                lib.load_scene(path)
            except lib.YglError:
                pass
    loaders.close()
    assert verdicts["same"] >= 30 and verdicts["refused"] >= 30, verdicts
