"""The device entry points from plain C (tests/c_abi/abi_device.c: gcc -std=c11 against the header alone): create on
cuda:0, step, read back, destroy -- and the same call sequence through the Python mirror must give the same bits."""
import os
import re
import subprocess

import numpy as np
import pytest

from softbodyunity_b200 import SoftBody, ingest, lib_path, load

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CUBE_POS = np.array([0, 0.05, 0, 1, 0.05, 0, 0, 1.05, 0, 1, 1.05, 0, 0, 0.05, 1, 1, 0.05, 1, 0, 1.05, 1, 1, 1.05, 1], np.float32).reshape(8, 3)
CUBE_TRI = np.array([0, 2, 1, 1, 2, 3, 4, 5, 6, 5, 7, 6, 0, 1, 4, 1, 5, 4, 2, 6, 3, 3, 6, 7, 0, 4, 2, 2, 4, 6, 1, 3, 5, 3, 7, 5], np.int32).reshape(12, 3)


def fnv1a(*arrays):
    h = 1469598103934665603
    for a in arrays:
        for byte in np.ascontiguousarray(a).view(np.uint8).ravel().tolist():
            h = ((h ^ byte) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def test_a_plain_c_host_steps_a_body_on_the_device(tmp_path):
    load()
    lib = lib_path()
    exe = tmp_path / "abi_device"
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_abi", "abi_device.c"), "-o", str(exe), lib, "-lm", "-Wl,-rpath," + os.path.dirname(lib)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    frames, spacing = 12, 0.125
    r = subprocess.run([str(exe), str(frames), str(spacing)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr
    m = re.search(r"V=(\d+) ns=(\d+) frames=\d+ state=([0-9a-f]+) surface=([0-9a-f]+) min_y=([-0-9.]+)", r.stdout)
    assert m, r.stdout
    # the same sequence through ctypes
    pos, tets, tris = ingest.tetrahedralize_surface(CUBE_POS, CUBE_TRI, spacing)
    sb = SoftBody(pos, tets, tris, tile_cap=256, stiffness=2.0e5)
    sb.step(frames=frames)
    x4, v4 = sb.get_state()
    sp, sn = sb.read_surface()
    assert int(m.group(1)) == len(pos) and int(m.group(2)) == len(sp)
    assert int(m.group(3), 16) == fnv1a(x4, v4)
    assert int(m.group(4), 16) == fnv1a(sp, sn)
    assert abs(float(m.group(5)) - float(x4[:, 1].min())) < 1e-6 and float(m.group(5)) >= 0.0
