"""Test helper: the reference's own test scenes (copied by oracle/copy_test_data.py into oracle/_ref/data, a
git-ignored build directory that travels to the GPU box) laid out again as <dir>/<name>/<name>.json + shapes/ +
textures/, the layout load_scene expects (uris are relative to the JSON file)."""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "oracle", "_ref", "data")
DATA_V40 = DATA + "_v40"   # the same scenes in scene format 4.0: host-only loader tests, does not travel to the GPU box


def available():
    return os.path.isdir(os.path.join(DATA, "scenes"))


def names():
    return sorted(f[:-5] for f in os.listdir(os.path.join(DATA, "scenes")) if f.endswith(".json")) if available() else []


def pool(kind, name):
    return os.path.join(DATA, "pool", kind, name)


def names_v40():
    d = os.path.join(DATA_V40, "scenes")
    return sorted(f[:-5] for f in os.listdir(d) if f.endswith(".json")) if os.path.isdir(d) else []


def scene_file(name, tmpdir, data=None):
    """Path of <tmpdir>/<name>/<name>.json with every shape / texture uri symlinked to its pooled file."""
    import json
    data = data or DATA
    d = os.path.join(str(tmpdir), name)
    dst = os.path.join(d, name + ".json")
    if not os.path.exists(dst):
        os.makedirs(d, exist_ok=True)
        os.symlink(os.path.join(data, "scenes", name + ".json"), dst)
        files = json.load(open(os.path.join(data, "scenes", name + ".files")))
        for uri, pooled in files.items():
            kind = uri.split("/")[0]
            os.makedirs(os.path.join(d, kind), exist_ok=True)
            os.symlink(os.path.join(data, "pool", kind, pooled), os.path.join(d, uri))
    return dst
