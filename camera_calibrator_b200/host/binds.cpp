// pybind11 module `pycalibrator`: same module, class, method and keyword names as
// the reference binding (reference src/binds.cpp:47-92).  The reference relies on
// pybind11/eigen.h; the casters below map the Eigen-free boundary types to numpy
// the same way (Affine3f <-> 4x4, Matrix3 <-> 3x3, vectors <-> 1-D arrays, point
// lists <-> (N,2)/(N,3) arrays or lists of rows).
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "calibrator.hh"
#include "extrinsics_calibrator.hh"
#include "geometry.hh"

namespace py = pybind11;
using namespace calibrator;

namespace pybind11 {
namespace detail {

using farray = py::array_t<float, py::array::c_style | py::array::forcecast>;

template <typename P, int N>
struct point_caster {
  PYBIND11_TYPE_CASTER(P, const_name("numpy.ndarray[float32[") + const_name<N>() + const_name("]]"));
  bool load(handle src, bool) {
    if (!src) return false;
    farray a;
    try {
      a = farray::ensure(src);
    } catch (...) {
      return false;
    }
    if (!a) {
      PyErr_Clear();
      return false;
    }
    if (a.size() != N) return false;
    const float* d = a.data();
    for (int i = 0; i < N; ++i) value(i) = d[i];
    return true;
  }
  static handle cast(const P& src, return_value_policy, handle) {
    farray a(N);
    for (int i = 0; i < N; ++i) a.mutable_data()[i] = src(i);
    return a.release();
  }
};
template <> struct type_caster<Point2D> : point_caster<Point2D, 2> {};
template <> struct type_caster<Point3D> : point_caster<Point3D, 3> {};
template <> struct type_caster<Plane> : point_caster<Plane, 4> {};

// (N,D) float arrays (or anything numpy can turn into one, e.g. a list of rows)
template <typename P, int D>
struct points_caster {
  using Vec = std::vector<P>;
  PYBIND11_TYPE_CASTER(Vec, const_name("numpy.ndarray[float32[n,") + const_name<D>() + const_name("]]"));
  bool load(handle src, bool) {
    if (!src) return false;
    farray a;
    try {
      a = farray::ensure(src);
    } catch (...) {
      return false;
    }
    if (!a) {
      PyErr_Clear();
      return false;
    }
    if (a.ndim() == 1 && a.size() == 0) {
      value.clear();
      return true;
    }
    if (a.ndim() != 2 || a.shape(1) != D) return false;
    const py::ssize_t n = a.shape(0);
    value.resize((size_t)n);
    const float* d = a.data();
    for (py::ssize_t i = 0; i < n; ++i)
      for (int k = 0; k < D; ++k) value[(size_t)i](k) = d[i * D + k];
    return true;
  }
  static handle cast(const Vec& src, return_value_policy, handle) {
    farray a({(py::ssize_t)src.size(), (py::ssize_t)D});
    float* d = a.mutable_data();
    for (size_t i = 0; i < src.size(); ++i)
      for (int k = 0; k < D; ++k) d[i * D + k] = src[i](k);
    return a.release();
  }
};
template <> struct type_caster<Points2D> : points_caster<Point2D, 2> {};
template <> struct type_caster<Points3D> : points_caster<Point3D, 3> {};

template <> struct type_caster<Matrix3> {
  PYBIND11_TYPE_CASTER(Matrix3, const_name("numpy.ndarray[float32[3,3]]"));
  bool load(handle src, bool) {
    farray a = farray::ensure(src);
    if (!a) {
      PyErr_Clear();
      return false;
    }
    if (a.ndim() != 2 || a.shape(0) != 3 || a.shape(1) != 3) return false;
    for (int i = 0; i < 9; ++i) value.m[i] = a.data()[i];
    return true;
  }
  static handle cast(const Matrix3& src, return_value_policy, handle) {
    farray a({3, 3});
    for (int i = 0; i < 9; ++i) a.mutable_data()[i] = src.m[i];
    return a.release();
  }
};

template <> struct type_caster<DynamicVector> {
  PYBIND11_TYPE_CASTER(DynamicVector, const_name("numpy.ndarray[float32[n]]"));
  bool load(handle src, bool) {
    farray a = farray::ensure(src);
    if (!a) {
      PyErr_Clear();
      return false;
    }
    value.d.assign(a.data(), a.data() + a.size());
    return true;
  }
  static handle cast(const DynamicVector& src, return_value_policy, handle) {
    farray a((py::ssize_t)src.size());
    for (size_t i = 0; i < src.size(); ++i) a.mutable_data()[i] = src(i);
    return a.release();
  }
};

// Affine3f <-> 4x4 numpy (reference binds.cpp:10-45)
template <> struct type_caster<Affine3f> {
  PYBIND11_TYPE_CASTER(Affine3f, const_name("numpy.ndarray[float32[4,4]]"));
  bool load(handle src, bool) {
    farray a = farray::ensure(src);
    if (!a) {
      PyErr_Clear();
      return false;
    }
    if (a.ndim() != 2 || a.shape(0) != 4 || a.shape(1) != 4) return false;
    const float* d = a.data();
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) value.R(r, c) = d[r * 4 + c];
      value.t(r) = d[r * 4 + 3];
    }
    return true;
  }
  static handle cast(const Affine3f& src, return_value_policy, handle) {
    farray a({4, 4});
    float* d = a.mutable_data();
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) d[r * 4 + c] = src.R(r, c);
      d[r * 4 + 3] = src.t(r);
    }
    d[12] = d[13] = d[14] = 0.f;
    d[15] = 1.f;
    return a.release();
  }
};

}  // namespace detail
}  // namespace pybind11

static py::dict report_dict(const SolveReport& r) {
  py::dict d;
  d["iterations"] = r.iterations;
  d["accepted"] = r.accepted;
  d["rejected"] = r.rejected;
  d["sweeps"] = r.sweeps;
  d["termination"] = r.termination;
  d["initial_cost"] = r.initial_cost;
  d["final_cost"] = r.final_cost;
  d["gradient_norm"] = r.gradient_norm;
  d["rms"] = r.rms;
  return d;
}

PYBIND11_MODULE(pycalibrator, m) {
  m.doc() = "Camera calibrator (B200-native LM hot path behind the reference API)";

  py::class_<Calibrator>(m, "Calibrator")
      .def(py::init<int, int>())
      .def("EstimateOpenCv", &Calibrator::EstimateOpenCv, py::arg("img_points"), py::arg("world_points"))
      .def("Estimate", &Calibrator::Estimate, py::arg("img_points"), py::arg("world_points"),
           py::call_guard<py::gil_scoped_release>())
      // The reference binds Optimize with std::vector<Eigen::Quaternionf>&, which has no caster and
      // is uncallable from Python; here qs is an (N,4) array (w,x,y,z) and ts an (N,3) array.
      .def("Optimize",
           [](Calibrator& self, const std::vector<Points2D>& img, const std::vector<Points3D>& world,
              py::array_t<float, py::array::c_style | py::array::forcecast> qs,
              py::array_t<float, py::array::c_style | py::array::forcecast> ts) {
             if (qs.ndim() != 2 || qs.shape(1) != 4 || ts.ndim() != 2 || ts.shape(1) != 3)
               throw std::invalid_argument("Optimize: qs must be (N,4), ts (N,3)");
             std::vector<Quaternion> q((size_t)qs.shape(0));
             std::vector<Point3D> t((size_t)ts.shape(0));
             for (size_t i = 0; i < q.size(); ++i)
               q[i] = Quaternion(qs.at(i, 0), qs.at(i, 1), qs.at(i, 2), qs.at(i, 3));
             for (size_t i = 0; i < t.size(); ++i) t[i] = Point3D(ts.at(i, 0), ts.at(i, 1), ts.at(i, 2));
             self.Optimize(img, world, q, t);
           },
           py::arg("img_points"), py::arg("world_points"), py::arg("qs"), py::arg("ts"))
      .def("GetK", &Calibrator::GetK)
      .def("SetK", &Calibrator::SetK, py::arg("K"))
      .def("GetDistortion", &Calibrator::GetDistortion)
      .def("SetDistortion", &Calibrator::SetDistortion, py::arg("distortion"))
      .def("ForceDistortionToConstant", &Calibrator::ForceDistortionToConstant)
      .def("Undistort", &Calibrator::Undistort, py::arg("img_points"))
      .def("Distort", &Calibrator::Distort, py::arg("normalized_points"))
      .def("GetLastReport", [](const Calibrator& c) { return report_dict(c.GetLastReport()); })
      .def("GetOptimizedPoses", [](const Calibrator& c) {
        std::vector<std::array<double, 4>> q;
        std::vector<std::array<double, 3>> t;
        c.GetOptimizedPoses(&q, &t);
        py::array_t<double> qa({(py::ssize_t)q.size(), (py::ssize_t)4}), ta({(py::ssize_t)t.size(), (py::ssize_t)3});
        for (size_t i = 0; i < q.size(); ++i)
          for (int k = 0; k < 4; ++k) qa.mutable_at(i, k) = q[i][k];
        for (size_t i = 0; i < t.size(); ++i)
          for (int k = 0; k < 3; ++k) ta.mutable_at(i, k) = t[i][k];
        return py::make_tuple(qa, ta);
      });

  py::class_<ExtrinsicsCalibrator>(m, "ExtrinsicsCalibrator")
      .def(py::init<>())
      .def("AddCameraTRig", &ExtrinsicsCalibrator::AddCameraTRig, py::arg("camera_T_rig"), py::arg("freeze") = false)
      .def("GetCameraTRig", &ExtrinsicsCalibrator::GetCameraTRig, py::arg("id"))
      .def("AddObservationFrame", &ExtrinsicsCalibrator::AddObservationFrame, py::arg("rig_T_world"))
      .def("GetObservationFrame", &ExtrinsicsCalibrator::GetObservationFrame, py::arg("id"))
      .def("AddWorldPoint", &ExtrinsicsCalibrator::AddWorldPoint, py::arg("frame_id"), py::arg("world_point"))
      .def("AddObservation", &ExtrinsicsCalibrator::AddObservation, py::arg("camera_id"),
           py::arg("world_point_id"), py::arg("image_point"))
      .def("AddObservations",
           [](ExtrinsicsCalibrator& self, size_t frame_id,
              py::array_t<float, py::array::c_style | py::array::forcecast> world_points,
              std::vector<size_t> camera_ids,
              py::array_t<float, py::array::c_style | py::array::forcecast> image_points) {
             if (world_points.ndim() != 2 || world_points.shape(1) != 3)
               throw std::invalid_argument("AddObservations: world_points must be (n,3)");
             const size_t n = (size_t)world_points.shape(0);
             if (image_points.ndim() != 3 || (size_t)image_points.shape(0) != n ||
                 (size_t)image_points.shape(1) != camera_ids.size() || image_points.shape(2) != 2)
               throw std::invalid_argument("AddObservations: image_points must be (n, n_cams, 2)");
             return self.AddObservations(frame_id, world_points.data(), n, camera_ids.data(), camera_ids.size(),
                                         image_points.data());
           },
           py::arg("frame_id"), py::arg("world_points"), py::arg("camera_ids"), py::arg("image_points"))
      .def("Optimize", &ExtrinsicsCalibrator::Optimize, py::call_guard<py::gil_scoped_release>())
      .def("Serialize", &ExtrinsicsCalibrator::Serialize, py::arg("fname"))
      .def("Parse", &ExtrinsicsCalibrator::Parse, py::arg("fname"))
      .def("RemoveObservationFrame", &ExtrinsicsCalibrator::RemoveObservationFrame, py::arg("observation_frame_id"))
      .def("RemoveObservationFrames", &ExtrinsicsCalibrator::RemoveObservationFrames,
           py::arg("observation_frame_ids"))
      .def("NumCameras", &ExtrinsicsCalibrator::NumCameras)
      .def("NumObservationFrames", &ExtrinsicsCalibrator::NumObservationFrames)
      .def("NumWorldPoints", &ExtrinsicsCalibrator::NumWorldPoints)
      .def("IsFrozen", &ExtrinsicsCalibrator::IsFrozen, py::arg("camera_id"))
      .def("GetObservations",
           [](const ExtrinsicsCalibrator& self, size_t frame_id) {
             py::list out;
             for (const auto& o : self.GetObservations(frame_id)) {
               py::dict d;
               d["camera_id"] = o.camera_id;
               d["world_point_idx"] = o.world_point_idx;
               d["world_point_id"] = o.world_point_id;
               d["image_point"] = py::make_tuple(o.image_point.x(), o.image_point.y());
               d["cost"] = o.cost;
               out.append(d);
             }
             return out;
           },
           py::arg("frame_id"))
      .def("GetLastReport", [](const ExtrinsicsCalibrator& c) { return report_dict(c.GetLastReport()); });

  m.def("EstimatePlaneFinite", &EstimatePlaneFinite, py::arg("p1"), py::arg("p2"), py::arg("p3"));
  m.def("PlaneNormal", &PlaneNormal, py::arg("plane"));
  m.def("RotationMatrixFromPlane", &RotationMatrixFromPlane, py::arg("plane"),
        py::arg("new_normal") = Point3D(0.f, 0.f, 1.f));
  m.def("ProjectToPlane", &ProjectToPlane, py::arg("plane"), py::arg("p"),
        py::arg("projection_direction") = std::nullopt);
  m.def("EstimateHomography", py::overload_cast<const Points2D&, const Points2D&>(&EstimateHomography),
        "Estimate homography", py::arg("p1"), py::arg("p2"));
  m.def("EstimateKFromHomographies", &EstimateKFromHomographies, py::arg("Hs"));
  m.def("RecoverExtrinsics", &RecoverExtrinsics,
        "Recover extrinsics from inverted calibration matrix and homography", py::arg("K_inv"), py::arg("H"));
  m.def("FixRotationMatrix", &FixRotationMatrix, py::arg("R"));
}
