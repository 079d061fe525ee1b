"""The boundary used from plain C: tests/c_abi/abi_host.c is compiled with gcc -std=c11 against
include/softbody_b200.h alone, linked to libsoftbody_b200.so and run (host-only entry points; no GPU)."""
import os
import subprocess

from softbodyunity_b200 import lib_path, load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_is_c_and_the_library_links_and_runs_from_c(tmp_path):
    load()  # builds the library if it is stale
    lib = lib_path()
    exe = tmp_path / "abi_host"
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_abi", "abi_host.c"), "-o", str(exe), lib, "-lm", "-Wl,-rpath," + os.path.dirname(lib)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([str(exe), str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr


def test_every_declared_entry_point_is_named_in_the_integration_guide():
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "softbody_b200.h")).read()
    guide = open(os.path.join(root, "INTEGRATION.md")).read()
    declared = set(re.findall(r"^(?:int|const char \*|uint64_t|void)\s*\*?\s*(sb_[a-z0-9_]+)\s*\(", header, re.M))
    assert len(declared) > 60
    # (families are written as `sb_halo_set/alloc/connect/pack/unpack/error`, `sb_ipc_export/open`, `sb_tetmesh_load` / `save`)
    def named(f):
        if f in guide:
            return True
        stem, _, last = f.rpartition("_")
        return re.search(re.escape(stem) + r"_[a-z/` ]*\b" + re.escape(last) + r"\b", guide) is not None
    missing = sorted(f for f in declared if not named(f))
    assert not missing, missing
