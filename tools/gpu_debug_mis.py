"""Development tool: locate pixels where a sampler differs from the unmodified reference."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("yocto-gl_b200", "oracle", "tests", "."):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np  # noqa: E402
import refbind  # noqa: E402
from ygl_b200 import abi, lib, scenes  # noqa: E402

ctx = lib.Context(0)
ref = refbind.Ref()
sc = scenes.features()
rs = ref.scene(sc)
sampler = int(sys.argv[1]) if len(sys.argv) > 1 else abi.SAMPLER_PATHMIS
for res in (128, 160, 200):
    for spp in (1, 4):
        for bounces in (1, 2, 3, 4, 8):
            p = abi.trace_params(resolution=res, samples=spp, bounces=bounces, sampler=sampler)
            a = ctx.trace_image(sc, p)
            b = rs.trace_image(p)["image"]
            d = np.argwhere((a.view(np.uint32) != b.view(np.uint32)).any(axis=-1))
            if len(d):
                print(res, spp, bounces, "mismatch pixels", len(d), flush=True)
                for (j, i) in d[:4]:
                    print("   ", j, i, a[j, i].tolist(), b[j, i].tolist(), [float(x).hex() for x in a[j, i]], [float(x).hex() for x in b[j, i]])
print("done")
