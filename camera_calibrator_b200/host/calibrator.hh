// calibrator::Calibrator — same class, methods and argument meaning as the
// reference (reference src/calibrator.hh:8-52); Optimize / Estimate run the LM
// solve on the GPU through the ccg C-ABI (include/ccg.h) instead of Ceres.
#pragma once
#include <set>

#include "types.hh"

struct ccg_ctx;

namespace calibrator {

// Result of the last Optimize()/Estimate() call (the reference discards Ceres'
// summary, calibrator.cpp:323-324; this is an addition, not a behaviour change).
struct SolveReport {
  int iterations = 0, accepted = 0, rejected = 0, sweeps = 0, termination = 0;
  double initial_cost = 0.0, final_cost = 0.0, gradient_norm = 0.0, rms = 0.0;
};

// Thread-local ccg context of the host classes (device: env CCG_DEVICE, default 0).
// Throws std::runtime_error when no CUDA device / library is usable: no CPU fallback.
ccg_ctx* HostContext();

class Calibrator {
 public:
  Calibrator(const int img_width, const int img_height);

  // reference: delegates to cv::calibrateCamera with flags = 0 (calibrator.cpp:16-45), i.e. own
  // initialisation, nothing frozen, same reprojection cost.  Here: the device Estimate path with an
  // empty frozen set (same minimum; OpenCV itself is not linked).
  void EstimateOpenCv(const std::vector<Points2D>& in_img_points, const std::vector<Points3D>& in_world_points);

  // closed-form initialisation (homographies -> K -> per-view pose) then Optimize (calibrator.cpp:47-68);
  // both run on the device behind ccg_intrinsics_estimate (host closed forms only when a distortion was
  // set beforehand, which the reference would use as the starting point)
  void Estimate(const std::vector<Points2D>& in_img_points, const std::vector<Points3D>& in_world_points);

  // LM over 9 intrinsics + one pose per view (calibrator.cpp:221-336).  As in the reference,
  // qs / ts are read as the starting poses and are NOT written back; the optimised poses of
  // the last call are available from GetOptimizedPoses().
  void Optimize(const std::vector<Points2D>& in_img_points, const std::vector<Points3D>& in_world_points,
                std::vector<Quaternion>& qs, std::vector<Point3D>& ts);

  Matrix3 GetK() const { return K_; }
  DynamicVector GetDistortion() const { return dist_; }
  void SetK(const Matrix3& K) { K_ = K; }
  void SetDistortion(const DynamicVector& dist) { dist_ = dist; }

  /// Freeze distortion coefficient `distortion_idx` (0..4 = k1,k2,p1,p2,k3), calibrator.cpp:338-340
  void ForceDistortionToConstant(const int distortion_idx);

  /// pixels -> undistorted normalised coordinates (calibrator.cpp:118-155)
  Points2D Undistort(const Points2D& img_points);
  /// normalised coordinates -> distorted pixels (calibrator.cpp:157-166)
  Points2D Distort(const Points2D& normalized_points);

  // additions (not in the reference)
  const SolveReport& GetLastReport() const { return report_; }
  void GetOptimizedPoses(std::vector<std::array<double, 4>>* qs, std::vector<std::array<double, 3>>* ts) const;

 private:
  void StoreResult(const SolveReport& rep, const double* intr, const double* q, const double* t, size_t n_img);
  int width_;
  int height_;
  Matrix3 K_{Matrix3::Identity()};
  DynamicVector dist_{DynamicVector::Zero(5)};
  std::set<int> constant_intrinsics_;
  SolveReport report_;
  std::vector<std::array<double, 4>> opt_q_;
  std::vector<std::array<double, 3>> opt_t_;
};

}  // namespace calibrator
