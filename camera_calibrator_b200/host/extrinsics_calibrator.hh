// calibrator::ExtrinsicsCalibrator — same class and methods as the reference
// (reference src/extrinsics_calibrator.hh:8-70); Optimize runs on the GPU through
// ccg_rig_solve.  The scene-graph bookkeeping (camera / frame / world-point ids,
// frame removal re-indexing) is host logic kept semantically identical.
#pragma once
#include <limits>
#include <set>
#include <string>

#include "calibrator.hh"
#include "types.hh"

namespace calibrator {

// Limit of this build: at most 16 cameras (AddCameraTRig accepts more, Optimize then throws with
// CCG_ERR_UNSUPPORTED; the reference has no bound, extrinsics_calibrator.cpp:9-17).
class ExtrinsicsCalibrator {
 public:
  /// \return Camera ID
  size_t AddCameraTRig(const Affine3f& camera_T_rig, const bool freeze = false);
  Affine3f GetCameraTRig(const size_t id) const;

  /// \return Frame ID
  size_t AddObservationFrame(const Affine3f& rig_T_world);
  Affine3f GetObservationFrame(const size_t id) const;

  /// \return World point ID (global, in insertion order)
  size_t AddWorldPoint(const size_t frame_id, const Point3D& world_point);

  /// Observation of a world point in NORMALISED image coordinates
  void AddObservation(const size_t camera_id, const size_t world_point_id, const Point2D& image_point);

  /// Bulk form of AddWorldPoint + AddObservation: adds `n` world points to
  /// `frame_id` and, for every camera id in `camera_ids`, one observation per point.  Semantically
  /// identical to calling AddWorldPoint / AddObservation point by point, camera by camera.
  /// image_points: n x n_cams x 2 (point-major).  Returns the id of the first new world point.
  size_t AddObservations(const size_t frame_id, const float* world_points_xyz, size_t n,
                         const size_t* camera_ids, size_t n_cams, const float* image_points);

  void Optimize();

  void Serialize(const std::string& fname) const;
  void Parse(const std::string& fname);

  void RemoveObservationFrame(const size_t observation_frame_id);
  void RemoveObservationFrames(const std::vector<size_t> observation_frame_ids);

  // ---- additions (introspection; not in the reference)
  struct Observation {
    size_t camera_id;
    size_t world_point_idx;
    size_t world_point_id;
    Point2D image_point;
    double cost{0.0};
  };
  size_t NumCameras() const { return camera_T_rigs_.size(); }
  size_t NumObservationFrames() const { return observation_frames_.size(); }
  size_t NumWorldPoints() const { return world_point_infos_.size(); }
  bool IsFrozen(size_t camera_id) const { return frozen_camera_T_rigs_.count(camera_id) != 0; }
  const std::vector<Observation>& GetObservations(size_t frame_id) const {
    return observation_frames_[frame_id].observations;
  }
  const SolveReport& GetLastReport() const { return report_; }

 private:
  std::vector<Affine3f> camera_T_rigs_;

  struct ObservationFrame {
    explicit ObservationFrame(const Affine3f& rig_T_world) : rig_T_world(rig_T_world) {}
    Affine3f rig_T_world;
    Points3D world_points;
    std::vector<Observation> observations;
  };
  std::vector<ObservationFrame> observation_frames_;

  struct WorldPointInfo {
    size_t observation_frame_id;
    size_t world_point_idx;
  };
  std::vector<WorldPointInfo> world_point_infos_;
  std::set<size_t> frozen_camera_T_rigs_;
  SolveReport report_;
};

}  // namespace calibrator
