"""CPU: the C-ABI library loads and exports every symbol include/ct3d.h declares (no compute calls)."""
import ctypes
import os
import re

from conftest import ROOT, load_pkg


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ct3d.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ct_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    pkg = load_pkg()
    lib_mod = __import__("importlib").import_module("3deecelltracker_b200._lib")
    assert os.path.isfile(lib_mod.LIB_PATH), "libct3d.so missing: run __graft_entry__.build()"
    handle = ctypes.CDLL(lib_mod.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 25
    for sym in declared:
        assert hasattr(handle, sym), f"{sym} declared in ct3d.h but not exported"
    # and the Python binding table covers exactly the header
    assert sorted(lib_mod.SIGNATURES) == declared
    assert lib_mod.lib().ct_abi_version() == 1
    assert pkg is not None


def test_host_side_queries_without_gpu():
    L = __import__("importlib").import_module("3deecelltracker_b200._lib")
    lib = L.lib()
    assert lib.ct_ffn_weight_count() == 560129                       # SURVEY a-5
    u = __import__("importlib").import_module("3deecelltracker_b200.unet3d")
    spec = _spec(L, u._SPECS["a"])
    assert lib.ct_unet_weight_count(ctypes.byref(spec)) == 512025    # SURVEY a-2
    assert lib.ct_prgls_workspace_bytes(164, 170, 150) > 164 * 164 * 8
    assert lib.ct_normalize_workspace_bytes(64, 64, 16) >= 2 * 64 * 64 * 16 * 4


def _spec(L, sp):
    s = L.CtUNetSpec()
    s.in_x, s.in_y, s.in_z = sp["input"]
    s.pool_x, s.pool_y, s.pool_z = sp["pool"]
    s.act_relu = sp["relu"]
    s.levels = len(sp["down"])
    for i, (a, b) in enumerate(sp["down"]):
        s.down[i][0], s.down[i][1] = a, b
    for i, (a, b) in enumerate(sp["up"]):
        s.up[i][0], s.up[i][1] = a, b
    s.out[0], s.out[1] = sp["out"]
    return s


def test_no_product_import_of_oracle():
    """The product package must never import the oracle (it is test infrastructure)."""
    pkg_dir = os.path.join(ROOT, "3deecelltracker_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src, f"{f} mentions the oracle"


def test_host_logic_schedules():
    import numpy as np
    from conftest import golden
    track = __import__("importlib").import_module("3deecelltracker_b200.track")
    lite = __import__("importlib").import_module("3deecelltracker_b200.trackerlite")
    g = golden("schedules.npz")
    for key in g.files:
        parts = key.split("_")
        if parts[0] == "ref":
            got = track.get_reference_vols(int(parts[1]), int(parts[2]), adjacent=bool(int(parts[3])))
        else:
            got = lite.get_volumes_list(int(parts[1]), [4, 7], int(parts[2]), bool(int(parts[3])), int(parts[4]))
        assert list(g[key]) == list(got), key
    ffn = __import__("importlib").import_module("3deecelltracker_b200.ffn")
    e = golden("trackerlite_em.npz")
    norm, (mean, scale) = ffn.normalize_points(e["points"], return_para=True)
    np.testing.assert_allclose(norm, e["ref_norm"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(scale, e["scale"], rtol=1e-11)


def test_header_is_plain_c(tmp_path):
    """include/ct3d.h is the C-ABI contract: it must compile as C99 (no C++ or torch types in the signatures)."""
    import subprocess
    src = tmp_path / "t.c"
    src.write_text('#include "ct3d.h"\nint main(void) { return ct_abi_version() == CT3D_ABI_VERSION ? 0 : 1; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-fsyntax-only",
                        "-I", os.path.join(ROOT, "include"), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
