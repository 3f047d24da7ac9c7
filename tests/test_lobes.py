"""Function-level parity of the shading code (a24-a25 of SURVEY.md §8a): the product's BSDF / phase-function code
(yocto-gl_b200/csrc/ygl_shading.cuh, compiled for the host by tests/cpp/lobes_host.cpp) against the reference's own
static dispatchers (yocto_trace.cpp:166-335 through oracle/ref_lobes.cpp) — eval_bsdfcos, sample_bsdfcos,
sample_bsdfcos_pdf, eval_delta, sample_delta, sample_delta_pdf, eval/sample_scattering(+pdf) — bit for bit on random
and edge-case inputs for all eight material types. Host-only (the device runs the same header with glibc's libm
restated bit for bit). Needs the reference oracle (oracle/_ref/libyocto_ref_lobes.so) and g++."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libyocto_ref_lobes.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/libyocto_ref_lobes.so not built")

OUT_NAMES = (["eval_bsdfcos"] * 3 + ["sample_bsdfcos"] * 3 + ["sample_bsdfcos_pdf"] + ["eval_delta"] * 3 +
             ["sample_delta"] * 3 + ["sample_delta_pdf"] + ["eval_scattering"] * 3 + ["sample_scattering"] * 3 +
             ["sample_scattering_pdf"] + ["bsdfcos_at_sampled"] * 3 + ["pdf_at_sampled"])


@pytest.fixture(scope="module")
def libs(tmp_path_factory):
    so = tmp_path_factory.mktemp("lobes") / "liblobes_host.so"
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so),
                    os.path.join(ROOT, "tests", "cpp", "lobes_host.cpp")], check=True)
    ours, ref = C.CDLL(str(so)), C.CDLL(REF)
    for f in (ours.ygl_host_lobes, ref.ref_lobes):
        f.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p]
        f.restype = None
    return ours.ygl_host_lobes, ref.ref_lobes


def unit(rng, n):
    v = rng.normal(size=(n, 3))
    return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


def make_inputs(rng, n, mtype):
    a = np.zeros((n, 26), np.float32)
    a[:, 0] = mtype
    a[:, 1:4] = rng.random((n, 3))
    rough = rng.random(n) ** 2
    rough[rng.random(n) < 0.15] = 0.0                     # delta lobes
    rough[rng.random(n) < 0.05] = 0.03 ** 2               # the clamp value of eval_material
    a[:, 4] = rough
    a[:, 5] = rng.random(n)
    a[:, 6] = np.where(rng.random(n) < 0.1, 1.0, 1.0 + rng.random(n))  # ior, incl. the |ior - 1| < 1e-3 branch
    a[:, 7:10] = rng.random((n, 3)) * 3 * (rng.random((n, 1)) < 0.8)    # density (zero sometimes)
    a[:, 10:13] = rng.random((n, 3))
    a[:, 13] = np.where(rng.random(n) < 0.2, 0.0, rng.uniform(-0.9, 0.9, n))
    nrm, out, inc = unit(rng, n), unit(rng, n), unit(rng, n)
    k = n // 10                                           # geometry edge cases
    out[:k] = nrm[:k]                                     # normal incidence
    inc[k:2 * k] = -out[k:2 * k]                          # straight through
    inc[2 * k:3 * k] = out[2 * k:3 * k]                   # retro-reflection (halfway = outgoing)
    t = unit(rng, k)
    out[3 * k:4 * k] = np.cross(nrm[3 * k:4 * k], t)      # grazing: dot(n, o) ~ 0
    out[3 * k:4 * k] /= np.linalg.norm(out[3 * k:4 * k], axis=1, keepdims=True)
    a[:, 14:17], a[:, 17:20], a[:, 20:23] = nrm, out, inc
    a[:, 23] = rng.random(n)
    a[:, 24:26] = rng.random((n, 2))
    a[: n // 50, 24:26] = 0.0                             # rn = 0: sample_microfacet at the pole
    return a


@pytest.mark.parametrize("mtype", range(8))
def test_lobes_match_reference_bit_for_bit(libs, mtype):
    ours, ref = libs
    rng = np.random.default_rng(100 + mtype)
    n = 400000
    a = make_inputs(rng, n, mtype)
    got, want = np.zeros((n, 25), np.float32), np.zeros((n, 25), np.float32)
    ours(a.ctypes.data, n, got.ctypes.data)
    ref(a.ctypes.data, n, want.ctypes.data)
    same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
    bad = np.argwhere(~same)
    assert len(bad) == 0, [(OUT_NAMES[c], a[r].tolist(), float(got[r, c]), float(want[r, c])) for r, c in bad[:3]]
