"""The managed shim cannot be compiled here (no C# toolchain in the image), so it is checked by parsing: every
[DllImport] it declares must be an export of the library with the same number of parameters as the C prototype,
and its blittable structs must have the field count / byte size of their C counterparts."""
import ctypes as C
import os
import re

from softbodyunity_b200 import _abi, load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CS = open(os.path.join(ROOT, "softbodyunity_b200", "unity", "SoftbodyB200.cs")).read()
HDR = open(os.path.join(ROOT, "include", "softbody_b200.h")).read()


def _c_prototypes():
    """name -> number of parameters, from the header."""
    text = re.sub(r"/\*.*?\*/", "", HDR, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|void|uint64_t|const char \*)\s*\*?\s*(sb_\w+)\s*\(([^;{}]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def test_every_dllimport_matches_a_c_prototype_and_an_export():
    lib = load()
    protos = _c_prototypes()
    imports = re.findall(r"\[DllImport\(Lib\)\]\s*public static extern\s+\w+\s+(sb_\w+)\s*\(([^)]*)\)", CS)
    assert len(imports) >= 20
    for name, args in imports:
        assert hasattr(lib, name), f"{name} is not exported"
        assert name in protos, f"{name} has no prototype in the header"
        n = 0 if not args.strip() else args.count(",") + 1
        assert n == protos[name], f"{name}: {n} parameters in the shim, {protos[name]} in the header"


def _cs_struct_bytes(name):
    body = re.search(r"public struct %s\b[^{]*\{(.*?)\n\}" % name, CS, flags=re.S).group(1)
    body = re.sub(r"public static .*", "", body, flags=re.S)  # helper methods come after the fields
    size = 0
    for typ, names in re.findall(r"public\s+(float|int|uint|IntPtr)\s+([^;]+);", body):
        size += (8 if typ == "IntPtr" else 4) * (names.count(",") + 1)
    return size


def test_managed_structs_have_the_c_sizes():
    assert _cs_struct_bytes("SbParams") == C.sizeof(_abi.SbParams) == 48
    assert _cs_struct_bytes("SbMeshDesc") == C.sizeof(_abi.SbMeshDesc) == 112
    assert _cs_struct_bytes("SbCollider") == 48
    # the version the shim insists on is the header's
    assert re.search(r"ver != (\d+)", CS).group(1) == re.search(r"#define SB_ABI_VERSION (\d+)u", HDR).group(1)


def _c_params():
    """name -> (return type, [parameter types]) from the header, comments stripped."""
    text = re.sub(r"/\*.*?\*/", "", HDR, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(int|void|uint64_t|const char \*)\s*(\*?)\s*(sb_\w+)\s*\(([^;{}]*?)\)\s*;", text, flags=re.S):
        args = m.group(4).strip()
        types = []
        if args not in ("", "void"):
            for a in args.split(","):
                a = " ".join(a.split())
                types.append(re.sub(r"\s*\w+$", "", a) if not a.endswith("*") else a)  # drop the parameter name
        out[m.group(3)] = ((m.group(1) + m.group(2)).strip(), types)
    return out


def test_every_dllimport_parameter_has_the_c_type():
    # by position: scalars must be the same width and signedness, everything the C side takes by pointer (handles, arrays,
    # structs, strings, out values) must be an IntPtr / ref / out / array / string on the managed side -- and vice versa
    scalars = {"float": "float", "uint32_t": "uint", "int32_t": "int", "int": "int", "uint64_t": "ulong"}
    protos = _c_params()
    imports = re.findall(r"\[DllImport\(Lib\)\]\s*public static extern\s+(\w+)\s+(sb_\w+)\s*\(([^)]*)\)", CS)
    checked = 0
    for ret, name, args in imports:
        c_ret, c_types = protos[name]
        assert {"int": "int", "void": "void", "const char *": "IntPtr", "uint64_t": "ulong"}[c_ret] == ret, name
        cs_types = [re.sub(r"\s*\w+$", "", " ".join(a.split())) for a in args.split(",")] if args.strip() else []
        assert len(cs_types) == len(c_types), name
        for ct, st in zip(c_types, cs_types):
            by_pointer = "*" in ct or ct in ("sb_handle", "sb_tetmesh_handle")
            if by_pointer:
                assert st == "IntPtr" or st == "string" or st.startswith(("ref ", "out ", "[In] ")) or st.endswith("[]"), (name, ct, st)
                if st == "string":
                    assert ct.replace(" ", "") == "constchar*", (name, ct, st)
                if st.startswith("out ") and st.split()[1] in ("uint", "int", "float", "ulong"):
                    base = ct.replace("*", "").replace("const", "").strip()
                    assert scalars[base] == st.split()[1], (name, ct, st)
            else:
                assert scalars[ct.replace("const", "").strip()] == st, (name, ct, st)
            checked += 1
    assert checked >= 50
