"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/kcbs_b200.h declares, validates arguments, and refuses to run without a B200 (no fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "kcbs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"KCBS_API\s+[\w\s\*]+?\b(kcbs_\w+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ("kcbs_create", "kcbs_destroy", "kcbs_propagate_batch", "kcbs_propagate_batch_dev", "kcbs_validity_batch",
              "kcbs_validity_batch_dev", "kcbs_first_conflict", "kcbs_first_conflict_dev", "kcbs_conflict_keys_dev",
              "kcbs_conflict_finish_dev", "kcbs_last_error", "kcbs_sync", "kcbs_propagate_pair_batch", "kcbs_validity_pair_batch",
              "kcbs_first_conflict_groups", "kcbs_rrt_plan", "kcbs_kcbs_solve"):
        assert s in syms
    assert len(syms) >= 30


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.load_library()
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/kcbs_b200.h but not exported"
    from kcbs_b200 import capi
    assert sorted(capi.EXPORTS) == declared_symbols()
    assert lib.kcbs_abi_version() == 1


def test_status_strings_and_defaults(pkg):
    from kcbs_b200 import capi
    lib = pkg.load_library()
    assert lib.kcbs_status_string(0) == b"ok"
    assert b"fallback" in lib.kcbs_status_string(3)
    w = capi._CWorld()
    lib.kcbs_world_defaults(C.byref(w), 32.0, 10.0)
    # StateSpaceDatabase.h:51-58, demo_*.cpp:115, StatePropagatorDatabase.h:53
    assert list(w.lo) == pytest.approx([0, 0, -1, -3.141592653589793 / 3])
    assert list(w.hi) == pytest.approx([32, 10, 1, 3.141592653589793 / 3])
    assert (w.step_size, w.integration_step, w.car_length) == (0.1, 1e-2, 0.5)


def test_create_rejects_bad_worlds(pkg):
    lib = pkg.load_library()
    from kcbs_b200 import capi
    h = C.c_void_p()
    assert lib.kcbs_create(None, 0, C.byref(h)) == 1
    w = capi._CWorld()
    lib.kcbs_world_defaults(C.byref(w), 10.0, 10.0)
    w.n_robots = 0
    assert lib.kcbs_create(C.byref(w), 0, C.byref(h)) == 1      # no robots
    assert b"n_robots" in lib.kcbs_last_error(None)
    dims = (C.c_double * 2)(1.0, 1.0)
    w.n_robots = 1
    w.robot_dims = C.cast(dims, C.POINTER(C.c_double))
    w.n_obstacles = 65
    assert lib.kcbs_create(C.byref(w), 0, C.byref(h)) == 1      # too many obstacles
    w.n_obstacles = 0
    w.car_length = 0.0
    assert lib.kcbs_create(C.byref(w), 0, C.byref(h)) == 1
    assert h.value is None


def test_no_cpu_fallback(pkg):
    """Without a CUDA device a context cannot be created: the product never computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.KcbsError) as e:
        pkg.Context(pkg.scenes.world("Empty32x32_10robots"))
    assert e.value.status == 3
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under the product package may reference it."""
    pkg_dir = os.path.join(ROOT, "k-cbs-demos_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".c", "Makefile")):
                text = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "kcbs_oracle" not in text and "oracle." not in text.replace("oracle.h", ""), os.path.join(dirpath, fn)


def test_scene_fixtures(pkg):
    s = pkg.scenes.SCENES
    assert len(s) == 13
    assert len(s["Benchmark_Empty32x32_30robots"]["robots"]) == 30
    assert s["Corridor10x10_4robots"]["obstacles"] == [[0.5, 1.8, 9.0, 1.0], [4.0, 5.0, 2.0, 2.0]]
    assert s["Benchmark_Corridor10x10_4robots"]["obstacles"][0] == [0.5, 2.0, 9.0, 0.5]
    assert len(s["Congested10x10_7robots"]["obstacles"]) == 9
    assert s["Corridor10x10_4robots"]["merge_bound"] == 10 and s["Congested10x10_7robots"]["merge_bound"] == 20
    assert (s["Empty32x32_10robots"]["robot_length"], s["Empty32x32_10robots"]["robot_width"]) == (0.7, 0.5)
    st = pkg.scenes.start_states("Corridor10x10_4robots")
    assert st[:2].T.tolist() == [[1.0, 0.5], [1.0, 3.5], [9.0, 0.5], [9.0, 3.5]]


def test_missing_library_fails_loudly(pkg, monkeypatch):
    """No silent fallback: without libkcbs_b200.so the binding raises instead of computing anything."""
    from kcbs_b200 import capi
    monkeypatch.setattr(capi, "_LIB", None)
    monkeypatch.setattr(capi, "lib_path", lambda: "/nonexistent/libkcbs_b200.so")
    with pytest.raises(ImportError):
        capi.load_library()


def test_header_is_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: the header must compile as C99 (no C++-isms), and as C++ too."""
    import subprocess
    src = tmp_path / "use.c"
    src.write_text('#include <kcbs_b200.h>\nint main(void) { kcbs_world w; kcbs_world_defaults(&w, 10.0, 10.0); return kcbs_abi_version() == KCBS_ABI_VERSION ? 0 : 1; }\n')
    inc = os.path.join(ROOT, "include")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", inc, str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c++", "-I", inc, str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # and it links and runs against the built library without a GPU (life-cycle entry points only)
    exe = tmp_path / "use"
    lib_dir = os.path.join(ROOT, "k-cbs-demos_b200")
    r = subprocess.run(["gcc", "-std=c99", "-I", inc, str(src), "-L", lib_dir, "-lkcbs_b200", f"-Wl,-rpath,{lib_dir}", "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert subprocess.run([str(exe)]).returncode == 0
