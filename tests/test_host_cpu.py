"""CPU tests of the host layer (no compute kernels run): the C-ABI library loads and exports
every symbol include/ccg.h declares, the pycalibrator module mirrors the reference binding, and
the ExtrinsicsCalibrator bookkeeping (ids, frame removal, JSON wire format, the reference's Parse
quirk) behaves like reference src/extrinsics_calibrator.cpp."""
import ctypes
import json
import math
import os
import re

import numpy as np
import pytest

import camera_calibrator_b200 as cc
from camera_calibrator_b200 import host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pc():
    cc.build()
    host.build()
    return host.load()


def test_cabi_exports_every_declared_symbol():
    cc.build()
    header = open(os.path.join(ROOT, "include", "ccg.h")).read()
    names = sorted(set(re.findall(r"\b(ccg_[a-z0-9_]+)\s*\(", header)))
    assert len(names) >= 25
    lib = ctypes.CDLL(cc._ccg.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_cabi_defaults_and_status_strings():
    o = cc.default_opts()
    assert o.max_iterations == 100 and o.initial_radius == 1e4 and o.min_relative_decrease == 1e-3
    assert abs(o.huber_a - float(np.float32(3.0) / np.float32(500.0))) == 0.0  # float literal, :176
    assert cc._ccg.status_string(0) == "ok"
    assert "CUDA" in cc._ccg.status_string(-2)


def test_module_surface_matches_reference_binding(pc):
    for cls, methods in {
        "Calibrator": ["EstimateOpenCv", "Estimate", "Optimize", "GetK", "SetK", "GetDistortion", "SetDistortion",
                       "ForceDistortionToConstant", "Undistort", "Distort"],
        "ExtrinsicsCalibrator": ["AddCameraTRig", "GetCameraTRig", "AddObservationFrame", "GetObservationFrame",
                                 "AddWorldPoint", "AddObservation", "Optimize", "Serialize", "Parse",
                                 "RemoveObservationFrame", "RemoveObservationFrames"],
    }.items():
        for m in methods:
            assert hasattr(getattr(pc, cls), m), (cls, m)
    for fn in ["EstimatePlaneFinite", "PlaneNormal", "RotationMatrixFromPlane", "ProjectToPlane",
               "EstimateHomography", "EstimateKFromHomographies", "RecoverExtrinsics", "FixRotationMatrix"]:
        assert hasattr(pc, fn), fn


def test_calibrator_accessors(pc):
    c = pc.Calibrator(1600, 1000)
    assert np.array_equal(c.GetK(), np.eye(3, dtype=np.float32))
    assert np.array_equal(c.GetDistortion(), np.zeros(5, np.float32))
    K = np.array([[1000, 0, 800], [0, 1000, 500], [0, 0, 1]], np.float32)
    dist = np.array([-4e-2, 5e-4, 1e-3, 2e-5, -3e-4], np.float32)
    c.SetK(K=K)
    c.SetDistortion(distortion=dist)
    assert np.array_equal(c.GetK(), K) and np.array_equal(c.GetDistortion(), dist)
    with pytest.raises(RuntimeError):
        c.EstimateOpenCv([], [])  # no views (and, on a box without a GPU, no device)
    import torch

    if not torch.cuda.is_available():
        # Distort / Undistort / Estimate run on the device; without one they fail loudly (no CPU fallback)
        with pytest.raises(RuntimeError):
            c.Distort([[0.1, 0.2]])


def test_geometry_known_answers(pc):
    rng = np.random.default_rng(1)
    for _ in range(10):  # reference test_geometry.cpp:172-194
        H = rng.uniform(-1, 1, (3, 3)).astype(np.float32)
        H /= H[2, 2]
        p1 = rng.uniform(-1, 1, (10, 2)).astype(np.float32)
        p2 = (H @ np.c_[p1, np.ones(10)].T).T
        p2 = (p2[:, :2] / p2[:, 2:]).astype(np.float32)
        H2 = pc.EstimateHomography(p1, p2)
        H2 = H2 / H2[2, 2]
        assert np.max(np.abs(H - H2)) <= 1e-4
    for _ in range(10):  # reference test_geometry.cpp: FixRotationMatrix returns a rotation
        R = pc.FixRotationMatrix(rng.uniform(-1, 1, (3, 3)).astype(np.float32))
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-5)
    pl = pc.EstimatePlaneFinite([0, 0, 1.0], [1, 0, 1.0], [0, 1, 1.0])
    assert np.allclose(pl, [0, 0, 1, -1], atol=1e-6)
    assert np.allclose(pc.PlaneNormal(pl), [0, 0, 1])
    assert np.allclose(pc.ProjectToPlane(pl, [0.3, 0.2, 2.0]), [0.3, 0.2, 1.0], atol=1e-6)
    assert np.allclose(pc.ProjectToPlane(pl, [0.3, 0.2, 2.0], [0.3, 0.2, 2.0]), [0.15, 0.1, 1.0], atol=1e-6)


def _populate(pc, n_frames=3, n_pts=4, n_cams=2):
    e = pc.ExtrinsicsCalibrator()
    T = np.eye(4, dtype=np.float32)
    for c in range(n_cams):
        Tc = T.copy()
        Tc[0, 3] = 0.01 * c
        assert e.AddCameraTRig(camera_T_rig=Tc, freeze=(c == 0)) == c
    wid = 0
    for f in range(n_frames):
        Tf = T.copy()
        Tf[:3, 3] = [0.5 + f, 0.6, 0.7]
        assert e.AddObservationFrame(rig_T_world=Tf) == f
        for p in range(n_pts):
            assert e.AddWorldPoint(frame_id=f, world_point=[0.1 * p, 0.2 * f, 0.0]) == wid
            for c in range(n_cams):
                e.AddObservation(camera_id=c, world_point_id=wid, image_point=[0.01 * wid, 0.02 * c])
            wid += 1
    return e


def test_ids_and_observation_filing(pc):
    """world_point_id is the global push order, world_point_idx the per-frame index, observations
    are filed under the world point's frame and start with cost = NaN (extrinsics_calibrator.cpp:25-49)."""
    e = _populate(pc)
    assert (e.NumCameras(), e.NumObservationFrames(), e.NumWorldPoints()) == (2, 3, 12)
    assert e.IsFrozen(0) and not e.IsFrozen(1)
    obs = e.GetObservations(1)
    assert [o["world_point_id"] for o in obs] == [4, 4, 5, 5, 6, 6, 7, 7]
    assert [o["world_point_idx"] for o in obs] == [0, 0, 1, 1, 2, 2, 3, 3]
    assert [o["camera_id"] for o in obs] == [0, 1] * 4
    assert all(math.isnan(o["cost"]) for o in obs)
    # a world point added late to an early frame keeps its global id and gets the next local index
    assert e.AddWorldPoint(0, [9, 9, 9]) == 12
    e.AddObservation(1, 12, [0.5, 0.5])
    last = e.GetObservations(0)[-1]
    assert (last["world_point_id"], last["world_point_idx"], last["camera_id"]) == (12, 4, 1)


def test_remove_frames_reindexes_like_the_reference(pc):
    e = _populate(pc, n_frames=4)
    e.RemoveObservationFrame(1)
    assert (e.NumObservationFrames(), e.NumWorldPoints()) == (3, 12)
    # old frame 2 is now frame 1 and its world point ids moved down by the removed frame's 4 points
    assert [o["world_point_id"] for o in e.GetObservations(1)] == [4, 4, 5, 5, 6, 6, 7, 7]
    assert np.allclose(e.GetObservationFrame(1)[:3, 3], [2.5, 0.6, 0.7])
    e2 = _populate(pc, n_frames=5)
    e2.RemoveObservationFrames([1, 3])  # any order in, larger ids removed first
    assert e2.NumObservationFrames() == 3 and e2.NumWorldPoints() == 12
    assert [o["world_point_id"] for o in e2.GetObservations(2)][0] == 8
    assert np.allclose(e2.GetObservationFrame(2)[:3, 3], [4.5, 0.6, 0.7])


def test_json_wire_format_round_trip(pc, tmp_path):
    e = _populate(pc)
    fn = str(tmp_path / "calib.json")
    e.Serialize(fname=fn)
    root = json.load(open(fn))
    assert sorted(root) == ["camera_T_rigs", "observation_frames", "world_points"]
    assert root["camera_T_rigs"][0]["frozen"] is True and len(root["camera_T_rigs"][0]["camera_T_rig"]) == 16
    # column-major 4x4: translation x of camera 1 at index 12
    assert abs(root["camera_T_rigs"][1]["camera_T_rig"][12] - 0.01) < 1e-8
    assert root["world_points"][5] == {"frame_id": 1, "world_point": pytest.approx([0.1, 0.2, 0.0], abs=1e-7)}
    ob = root["observation_frames"][2]["observations"][3]
    assert sorted(ob) == ["camera_id", "cost", "image_point", "world_point_id"] and ob["cost"] is None
    fresh = pc.ExtrinsicsCalibrator()
    fresh.Parse(fname=fn)
    assert (fresh.NumCameras(), fresh.NumObservationFrames(), fresh.NumWorldPoints()) == (2, 3, 12)
    assert fresh.IsFrozen(0) and not fresh.IsFrozen(1)
    assert np.array_equal(fresh.GetCameraTRig(1), e.GetCameraTRig(1))
    for f in range(3):
        a, b = e.GetObservations(f), fresh.GetObservations(f)
        assert [(o["camera_id"], o["world_point_id"], o["world_point_idx"], o["image_point"]) for o in a] == \
               [(o["camera_id"], o["world_point_id"], o["world_point_idx"], o["image_point"]) for o in b]


def test_parse_on_populated_object_keeps_old_cameras(pc, tmp_path):
    """Reference quirk (extrinsics_calibrator.cpp:349-351): Parse does not clear the
    camera list, so cameras are appended and the ORIGINAL camera 0 is no longer frozen."""
    e = _populate(pc)
    fn = str(tmp_path / "calib.json")
    e.Serialize(fn)
    e.Parse(fn)
    assert e.NumCameras() == 4
    assert not e.IsFrozen(0) and e.IsFrozen(2)


def test_bulk_add_equals_point_by_point(pc):
    a = _populate(pc, n_frames=1, n_pts=0)
    b = _populate(pc, n_frames=1, n_pts=0)
    pts = np.array([[0.1, 0.2, 0.0], [0.3, 0.4, 0.0], [0.5, 0.6, 0.0]], np.float32)
    img = np.arange(12, dtype=np.float32).reshape(3, 2, 2) / 100
    assert a.AddObservations(frame_id=0, world_points=pts, camera_ids=[0, 1], image_points=img) == 0
    for i in range(3):
        wid = b.AddWorldPoint(0, pts[i])
        for c in range(2):
            b.AddObservation(c, wid, img[i, c])
    oa, ob = a.GetObservations(0), b.GetObservations(0)
    assert [(o["camera_id"], o["world_point_id"], o["image_point"]) for o in oa] == \
           [(o["camera_id"], o["world_point_id"], o["image_point"]) for o in ob]


def test_product_code_never_imports_or_links_the_oracle():
    """oracle/ is test infrastructure: nothing under camera_calibrator_b200/ or include/ may import, include, link or
    execute it (comments that mention it are fine), and the C-ABI library must not depend on liboracle.so."""
    import re
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bad = []
    for top in ("camera_calibrator_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(root, top)):
            if "variants" in dirpath:
                continue
            for f in files:
                if not f.endswith((".py", ".cpp", ".hh", ".h", ".cu", ".cuh")) and f != "Makefile":
                    continue
                for i, line in enumerate(open(os.path.join(dirpath, f), errors="replace"), 1):
                    code = line.split("//")[0].split("#")[0] if not f.endswith(".py") else line.split("#")[0]
                    if f.endswith((".cpp", ".hh", ".h", ".cu", ".cuh")):
                        code = line if line.lstrip().startswith("#include") else code
                    if re.search(r"(import\s+oracle|from\s+oracle|#include\s*[\"<][^\">]*oracle|liboracle|-loracle|oracle/)", code):
                        bad.append(f"{os.path.join(dirpath, f)}:{i}: {line.strip()}")
    bad = [b for b in bad if "restatement lives in oracle/datagen.cpp" not in b]  # a docstring sentence in synth.py
    assert not bad, "\n".join(bad)
    lib = os.path.join(root, "camera_calibrator_b200", "libccg.so")
    needed = subprocess.run(["readelf", "-d", lib], capture_output=True, text=True).stdout
    assert "oracle" not in needed
