"""The C++ host-side mirror of the reference's plugin headers (k-cbs-demos_b200/includes/) on top of the C ABI.

CPU part: it builds, and every demo main of the reference compiles UNCHANGED against it (drop-in surface; only where
the reference tree is present, i.e. in the build container).  GPU part: the mirror's own unit test and demo driver run
end to end (scalar virtuals, batched extras, oc::RRT, omrc::KCBS, omrc::PP, Benchmark, SystemMerger)."""
import glob
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "k-cbs-demos_b200")
DEMOS = os.path.join(PKG, "demos")
REF = "/root/reference"
INC = ["-I", os.path.join(PKG, "includes"), "-I", os.path.join(PKG, "ompl_shim"), "-I", os.path.join(ROOT, "include")]


def test_mirror_builds(pkg):
    pkg.load_library()
    subprocess.check_call(["make", "-C", DEMOS, "-s", "-j", "4"])
    assert os.path.exists(os.path.join(DEMOS, "kcbs_demo")) and os.path.exists(os.path.join(DEMOS, "test_mirror"))


def test_mirror_keeps_the_reference_file_and_symbol_names():
    names = sorted(os.path.basename(p) for p in glob.glob(os.path.join(PKG, "includes", "*.h")))
    for h in ("Benchmark.h", "ControlSpaceDatabase.h", "GoalRegionDatabase.h", "Obstacle.h", "PlannerAllocatorDatabase.h", "Robot.h",
              "StatePropagatorDatabase.h", "StateSpaceDatabase.h", "StateValidityCheckerDatabase.h", "SystemMergerDatabase.h"):
        assert h in names
    text = "".join(open(p).read() for p in glob.glob(os.path.join(PKG, "includes", "*.h")))
    for sym in ("SecondOrderCarODE", "SecondOrderCarODEPostIntegration", "TwoSecondOrderCarsODE", "TwoSecondOrderCarODEPostIntegration",
                "TwoSecondOrderCarStatePropagator", "homogeneous2ndOrderCarSystemSVC", "homogeneousTwo2ndOrderCarSystemSVC",
                "homogeneous2ndOrderCarSystemMerger", "GoalRegion2ndOrderCar", "GoalRegionTwo2ndOrderCars", "allocateControlRRT",
                "createBounded2ndOrderCarStateSpace", "createBoundedTwo2ndOrderCarStateSpace", "createUniform2DRealVectorControlSpace",
                "createUniform4DRealVectorControlSpace", "RectangularRobot", "CompoundRobot", "RectangularObstacle", "setPoroblemDefinition",
                "runKCBS", "runPP", "writeCSV", "updateRobotMap", "areStatesValid", "isValid"):
        assert sym in text, sym


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference tree not present (GPU box)")
def test_reference_demo_mains_compile_unchanged():
    srcs = sorted(glob.glob(os.path.join(REF, "src", "demo_*.cpp")))
    assert len(srcs) == 13
    for src in srcs:
        r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-w"] + INC + [src], capture_output=True, text=True)
        assert r.returncode == 0, f"{os.path.basename(src)}:\n{r.stderr[-2000:]}"


@pytest.mark.gpu
def test_mirror_unit_test_on_gpu(tmp_path):
    exe = os.path.join(DEMOS, "test_mirror")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", DEMOS, "-s"])
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "all checks passed" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("scene,planner", [("Corridor10x10_4robots", "K-CBS"), ("Empty32x32_10robots", "K-CBS"), ("Empty32x32_10robots", "PP"),
                                           ("Congested10x10_7robots", "K-CBS")])
def test_demo_driver_on_gpu(tmp_path, scene, planner):
    exe = os.path.join(DEMOS, "kcbs_demo")
    r = subprocess.run([exe, scene, planner, "120"], cwd=tmp_path, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "Found Solution" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    plan = open(tmp_path / "plan.txt").read()
    assert plan.startswith("Robot 0")
    if planner == "K-CBS":
        assert "nodes expanded" in open(tmp_path / "tree.txt").read()


REF_BINS = os.path.join(DEMOS, "_ref")


@pytest.mark.gpu
@pytest.mark.parametrize("demo", ["demo_Corridor10x10_4robots_dyn2ndOrderCars", "demo_Congested10x10_7robots_dyn2ndOrderCars",
                                  "demo_Empty32x32_10robots_dyn2ndOrderCars"])
def test_reference_demo_main_runs_on_gpu(tmp_path, demo):
    """The reference's OWN demo main (compiled where it lies under /root/reference/src by demos/Makefile, linked against
    the mirror headers + libkcbs_b200.so) plans on the GPU and writes the files the reference writes."""
    exe = os.path.join(REF_BINS, demo)
    if not os.path.exists(exe):
        pytest.skip("demos/_ref/ not built (needs /root/reference at build time)")
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "Found Solution" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    assert open(tmp_path / "plan.txt").read().startswith("Robot 0")       # demo_*.cpp:176-180
    assert os.path.exists(tmp_path / "tree.txt")                          # demo_*.cpp:182-186


@pytest.mark.gpu
def test_reference_benchmark_main_runs_on_gpu(tmp_path):
    """demo_Benchmark_Empty32x32_30robots: Benchmark::runKCBS (Benchmark.h:151-173), one solve of the 30-robot scene,
    a header and one row appended to the CSV (Benchmark.h:200-219)."""
    exe = os.path.join(REF_BINS, "demo_Benchmark_Empty32x32_30robots_dyn2ndOrderCars")
    if not os.path.exists(exe):
        pytest.skip("demos/_ref/ not built (needs /root/reference at build time)")
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    csvs = [f for f in os.listdir(tmp_path) if f.endswith(".csv")]
    assert csvs, os.listdir(tmp_path)
    rows = open(tmp_path / csvs[0]).read().strip().splitlines()
    assert len(rows) >= 2 and "Success" in rows[0] and "Computation Time" in rows[0]
    cols = dict(zip(rows[0].split(","), rows[1].split(",")))
    assert [v for k, v in cols.items() if "Success" in k] == ["1"], cols
