"""Host model of the persistent mode's ticket ring queues (tests/cpp/ring_stress.cpp): the protocol of
ygl_kernels.cu (ring_reserve / ring_publish / ring_take) on std::atomic, driven by threads that play traversal,
shading and light-pdf warps. No entry may be lost or taken twice, tickets may run ahead of the producers, and the
job must terminate - for a large frame, a long-running small one and a tiny one."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def ring_stress(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("ring") / "ring_stress")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", os.path.join(HERE, "cpp", "ring_stress.cpp"), "-o", exe],
                   check=True)
    return exe


@pytest.mark.parametrize("lanes,hops,threads", [(20000, 30, 2), (2000, 200, 2), (64, 1500, 1)])
def test_ticket_rings_lose_and_duplicate_nothing(ring_stress, lanes, hops, threads):
    r = subprocess.run([ring_stress, str(lanes), str(hops), str(threads)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"done {lanes} abort 0 double_take 0 short 0" in r.stdout
