// Closed-form initialisation geometry used by Calibrator::Estimate, same free
// functions as reference src/geometry.hh:9-48 (host side, double precision SVDs).
#pragma once
#include <optional>
#include <tuple>

#include "types.hh"

namespace calibrator {

struct Plane {
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  float operator()(int i) const { return v[i]; }
  float& operator()(int i) { return v[i]; }
};

// reference geometry.hh:13-29 (plane helpers of the synthetic generator)
Plane EstimatePlaneFinite(const Point3D& p1, const Point3D& p2, const Point3D& p3);
Point3D PlaneNormal(const Plane& plane);
Matrix3 RotationMatrixFromPlane(const Plane& plane, const Point3D& new_normal = Point3D(0.f, 0.f, 1.f));
Point3D ProjectToPlane(const Plane& plane, const Point3D& p,
                       const std::optional<Point3D>& projection_direction = std::nullopt);

// reference geometry.hh:31-35 — p2 ~ H p1 (DLT; the z of 3-D points is ignored: planar target)
Matrix3 EstimateHomography(const Points2D& p1, const Points2D& p2);
Matrix3 EstimateHomography(const Points2D& p1, const Points3D& p2);
Matrix3 EstimateHomography(const Points3D& p1, const Points2D& p2);
Matrix3 EstimateHomography(const Points3D& p1, const Points3D& p2);

// reference geometry.hh:37-48
Matrix3 EstimateKFromHomographies(const std::vector<Matrix3>& Hs);
std::tuple<Matrix3, Point3D> RecoverExtrinsics(const Matrix3& K_inv, const Matrix3& H);
Matrix3 FixRotationMatrix(const Matrix3& R);

}  // namespace calibrator
