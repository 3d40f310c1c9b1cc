// Host side of calibrator::ExtrinsicsCalibrator (reference src/extrinsics_calibrator.cpp).
//   id bookkeeping ........ reference :9-49      (kept semantically identical, integer-exact)
//   Optimize packing ...... reference :105-137   (rotation -> Quaterniond in double, float t)
//   solve ................. ccg_rig_solve instead of the Ceres problem of :92-216
//   cost read-back ........ reference :219-225   (0.5*rho per observation)
//   parameter write-back .. reference :228-256   (float-rounded q, normalised, rotation matrix)
//   JSON wire format ...... reference :268-413   (same keys, column-major 4x4 transforms)
//   frame removal ......... reference :415-452
#include "extrinsics_calibrator.hh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <functional>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>

#include "../../include/ccg.h"

namespace calibrator {

size_t ExtrinsicsCalibrator::AddCameraTRig(const Affine3f& camera_T_rig, const bool freeze) {
  camera_T_rigs_.push_back(camera_T_rig);
  const size_t id = camera_T_rigs_.size() - 1;
  if (freeze) frozen_camera_T_rigs_.insert(id);
  return id;
}

Affine3f ExtrinsicsCalibrator::GetCameraTRig(const size_t id) const { return camera_T_rigs_.at(id); }

size_t ExtrinsicsCalibrator::AddObservationFrame(const Affine3f& rig_T_world) {
  observation_frames_.emplace_back(rig_T_world);
  return observation_frames_.size() - 1;
}

Affine3f ExtrinsicsCalibrator::GetObservationFrame(const size_t id) const {
  return observation_frames_.at(id).rig_T_world;
}

size_t ExtrinsicsCalibrator::AddWorldPoint(const size_t frame_id, const Point3D& world_point) {
  ObservationFrame& frame = observation_frames_.at(frame_id);
  WorldPointInfo info;
  info.observation_frame_id = frame_id;
  info.world_point_idx = frame.world_points.size();
  world_point_infos_.push_back(info);
  frame.world_points.push_back(world_point);
  return world_point_infos_.size() - 1;
}

void ExtrinsicsCalibrator::AddObservation(const size_t camera_id, const size_t world_point_id,
                                          const Point2D& image_point) {
  const WorldPointInfo& info = world_point_infos_.at(world_point_id);
  Observation o;
  o.camera_id = camera_id;
  o.image_point = image_point;
  o.world_point_idx = info.world_point_idx;
  o.world_point_id = world_point_id;
  o.cost = std::numeric_limits<double>::quiet_NaN();
  observation_frames_[info.observation_frame_id].observations.push_back(o);
}

size_t ExtrinsicsCalibrator::AddObservations(const size_t frame_id, const float* xyz, size_t n,
                                             const size_t* camera_ids, size_t n_cams,
                                             const float* image_points) {
  const size_t first = world_point_infos_.size();
  ObservationFrame& frame = observation_frames_.at(frame_id);
  frame.world_points.reserve(frame.world_points.size() + n);
  frame.observations.reserve(frame.observations.size() + n * n_cams);
  for (size_t i = 0; i < n; ++i) {
    const size_t id = AddWorldPoint(frame_id, Point3D(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
    for (size_t c = 0; c < n_cams; ++c) {
      const float* ip = image_points + 2 * (i * n_cams + c);
      AddObservation(camera_ids[c], id, Point2D(ip[0], ip[1]));
    }
  }
  return first;
}

namespace {

// [ext] Eigen::Quaterniond(Matrix3d): trace test, else the largest diagonal entry
void rotation_to_quat(const Matrix3& Rf, double q[4]) {
  double R[3][3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) R[r][c] = (double)Rf(r, c);
  const double tr = R[0][0] + R[1][1] + R[2][2];
  if (tr > 0.0) {
    double s = std::sqrt(tr + 1.0);
    q[0] = 0.5 * s;
    s = 0.5 / s;
    q[1] = (R[2][1] - R[1][2]) * s;
    q[2] = (R[0][2] - R[2][0]) * s;
    q[3] = (R[1][0] - R[0][1]) * s;
    return;
  }
  int i = 0;
  if (R[1][1] > R[0][0]) i = 1;
  if (R[2][2] > R[i][i]) i = 2;
  const int j = (i + 1) % 3, k = (j + 1) % 3;
  double s = std::sqrt(R[i][i] - R[j][j] - R[k][k] + 1.0);
  q[1 + i] = 0.5 * s;
  s = 0.5 / s;
  q[0] = (R[k][j] - R[j][k]) * s;
  q[1 + j] = (R[j][i] + R[i][j]) * s;
  q[1 + k] = (R[k][i] + R[i][k]) * s;
}

// reference :228-256: components rounded to float, held in a double quaternion,
// normalised, converted to a rotation matrix, cast to float
Matrix3 quat_to_rotation_f(const double* qd) {
  double w = (double)(float)qd[0], x = (double)(float)qd[1], y = (double)(float)qd[2], z = (double)(float)qd[3];
  const double n = std::sqrt(w * w + x * x + y * y + z * z);
  w /= n; x /= n; y /= n; z /= n;
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x,
               tyy = ty * y, tyz = tz * y, tzz = tz * z;
  Matrix3 R;
  R(0, 0) = (float)(1.0 - (tyy + tzz)); R(0, 1) = (float)(txy - twz);         R(0, 2) = (float)(txz + twy);
  R(1, 0) = (float)(txy + twz);         R(1, 1) = (float)(1.0 - (txx + tzz)); R(1, 2) = (float)(tyz - twx);
  R(2, 0) = (float)(txz - twy);         R(2, 1) = (float)(tyz + twx);         R(2, 2) = (float)(1.0 - (txx + tyy));
  return R;
}

}  // namespace

void ExtrinsicsCalibrator::Optimize() {
  const size_t C = camera_T_rigs_.size(), F = observation_frames_.size(), NW = world_point_infos_.size();
  std::vector<double> qc(4 * C), tc(3 * C), qf(4 * F), tf(3 * F);
  std::vector<uint8_t> frozen(C, 0);
  for (size_t i = 0; i < C; ++i) {
    rotation_to_quat(camera_T_rigs_[i].rotation(), &qc[4 * i]);
    for (int k = 0; k < 3; ++k) tc[3 * i + k] = camera_T_rigs_[i].translation()(k);
    frozen[i] = frozen_camera_T_rigs_.count(i) ? 1 : 0;
  }
  for (size_t i = 0; i < F; ++i) {
    rotation_to_quat(observation_frames_[i].rig_T_world.rotation(), &qf[4 * i]);
    for (int k = 0; k < 3; ++k) tf[3 * i + k] = observation_frames_[i].rig_T_world.translation()(k);
  }
  std::vector<float> Xw(3 * NW);
  for (size_t i = 0; i < NW; ++i) {
    const WorldPointInfo& info = world_point_infos_[i];
    const Point3D& p = observation_frames_[info.observation_frame_id].world_points[info.world_point_idx];
    Xw[3 * i] = p(0);
    Xw[3 * i + 1] = p(1);
    Xw[3 * i + 2] = p(2);
  }
  std::vector<int64_t> foff(F + 1, 0);
  for (size_t f = 0; f < F; ++f) foff[f + 1] = foff[f] + (int64_t)observation_frames_[f].observations.size();
  const size_t N = (size_t)foff[F];
  std::vector<float> u(N), v(N);
  std::vector<int32_t> cam(N);
  std::vector<int64_t> pt(N);
  size_t o = 0;
  for (size_t f = 0; f < F; ++f)
    for (const Observation& ob : observation_frames_[f].observations) {
      u[o] = ob.image_point.x();
      v[o] = ob.image_point.y();
      cam[o] = (int32_t)ob.camera_id;
      pt[o] = (int64_t)ob.world_point_id;
      ++o;
    }
  std::vector<double> cost(N, 0.0);
  ccg_opts opts;
  ccg_default_opts(&opts);
  opts.max_iterations = 1000;                    // extrinsics_calibrator.cpp:211
  opts.huber_a = (double)(3.0f / 500.0f);        // extrinsics_calibrator.cpp:176
  ccg_report rep;
  ccg_ctx* ctx = HostContext();
  const int rc = ccg_rig_solve(ctx, u.data(), v.data(), cam.data(), pt.data(), foff.data(), (int64_t)F, Xw.data(),
                               (int64_t)NW, (int32_t)C, qc.data(), tc.data(), frozen.data(), qf.data(), tf.data(),
                               cost.data(), &opts, &rep);
  if (rc != CCG_OK)
    throw std::runtime_error(std::string("ExtrinsicsCalibrator::Optimize: ") + ccg_status_string(rc) + " " +
                             ccg_last_error(ctx));
  report_.iterations = rep.iterations;
  report_.accepted = rep.accepted;
  report_.rejected = rep.rejected;
  report_.sweeps = rep.sweeps;
  report_.termination = rep.termination;
  report_.initial_cost = rep.initial_cost;
  report_.final_cost = rep.final_cost;
  report_.gradient_norm = rep.gradient_norm;
  report_.rms = rep.rms;
  // per-observation cost
  o = 0;
  for (size_t f = 0; f < F; ++f)
    for (Observation& ob : observation_frames_[f].observations) ob.cost = cost[o++];
  // parameters
  for (size_t i = 0; i < C; ++i) {
    camera_T_rigs_[i].linear() = quat_to_rotation_f(&qc[4 * i]);
    camera_T_rigs_[i].translation() = Point3D((float)tc[3 * i], (float)tc[3 * i + 1], (float)tc[3 * i + 2]);
  }
  for (size_t i = 0; i < F; ++i) {
    Affine3f& T = observation_frames_[i].rig_T_world;
    T.linear() = quat_to_rotation_f(&qf[4 * i]);
    T.translation() = Point3D((float)tf[3 * i], (float)tf[3 * i + 1], (float)tf[3 * i + 2]);
  }
}

// ------------------------------------------------------------------ JSON ----
namespace {

void put_num(std::ostream& os, double x) {
  if (std::isnan(x) || std::isinf(x)) {
    os << "null";  // nlohmann::json dumps non-finite numbers as null
    return;
  }
  char buf[40];
  std::snprintf(buf, sizeof(buf), "%.17g", x);
  os << buf;
}

void put_transform(std::ostream& os, const Affine3f& T) {
  const auto m = T.matrix_colmajor();
  os << "[";
  for (int i = 0; i < 16; ++i) {
    if (i) os << ",";
    put_num(os, m[i]);
  }
  os << "]";
}

// minimal JSON reader (objects, arrays, numbers, strings, true/false/null)
struct JValue {
  enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
  bool b = false;
  double num = 0.0;
  std::string str;
  std::vector<JValue> arr;
  std::map<std::string, JValue> obj;
  const JValue& at(const std::string& k) const {
    auto it = obj.find(k);
    if (it == obj.end()) throw std::runtime_error("JSON: missing key '" + k + "'");
    return it->second;
  }
};

struct JParser {
  const std::string& s;
  size_t i = 0;
  explicit JParser(const std::string& src) : s(src) {}
  void ws() {
    while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) ++i;
  }
  [[noreturn]] void fail(const char* what) { throw std::runtime_error(std::string("JSON parse error: ") + what); }
  JValue parse() {
    ws();
    if (i >= s.size()) fail("unexpected end");
    JValue v;
    const char c = s[i];
    if (c == '{') {
      v.kind = JValue::Obj;
      ++i;
      ws();
      if (s[i] == '}') { ++i; return v; }
      for (;;) {
        ws();
        JValue key = parse();
        if (key.kind != JValue::Str) fail("object key");
        ws();
        if (s[i] != ':') fail("':' expected");
        ++i;
        v.obj[key.str] = parse();
        ws();
        if (s[i] == ',') { ++i; continue; }
        if (s[i] == '}') { ++i; break; }
        fail("',' or '}' expected");
      }
    } else if (c == '[') {
      v.kind = JValue::Arr;
      ++i;
      ws();
      if (s[i] == ']') { ++i; return v; }
      for (;;) {
        v.arr.push_back(parse());
        ws();
        if (s[i] == ',') { ++i; continue; }
        if (s[i] == ']') { ++i; break; }
        fail("',' or ']' expected");
      }
    } else if (c == '"') {
      v.kind = JValue::Str;
      ++i;
      while (i < s.size() && s[i] != '"') {
        if (s[i] == '\\' && i + 1 < s.size()) ++i;
        v.str.push_back(s[i++]);
      }
      ++i;
    } else if (s.compare(i, 4, "true") == 0) {
      v.kind = JValue::Bool; v.b = true; i += 4;
    } else if (s.compare(i, 5, "false") == 0) {
      v.kind = JValue::Bool; v.b = false; i += 5;
    } else if (s.compare(i, 4, "null") == 0) {
      v.kind = JValue::Null; i += 4;
    } else {
      char* end = nullptr;
      v.kind = JValue::Num;
      v.num = std::strtod(s.c_str() + i, &end);
      if (end == s.c_str() + i) fail("value expected");
      i = (size_t)(end - s.c_str());
    }
    return v;
  }
};

Affine3f get_transform(const JValue& a) {
  float m[16] = {0};
  for (size_t i = 0; i < a.arr.size() && i < 16; ++i) m[i] = (float)a.arr[i].num;
  return Affine3f::from_colmajor(m);
}

}  // namespace

void ExtrinsicsCalibrator::Serialize(const std::string& fname) const {
  std::ofstream os(fname.c_str());
  if (!os) throw std::runtime_error("ExtrinsicsCalibrator::Serialize: cannot open " + fname);
  os << "{\"camera_T_rigs\":[";
  for (size_t i = 0; i < camera_T_rigs_.size(); ++i) {
    if (i) os << ",";
    os << "{\"camera_T_rig\":";
    put_transform(os, camera_T_rigs_[i]);
    os << ",\"frozen\":" << (frozen_camera_T_rigs_.count(i) ? "true" : "false") << "}";
  }
  os << "],\"observation_frames\":[";
  for (size_t f = 0; f < observation_frames_.size(); ++f) {
    if (f) os << ",";
    const ObservationFrame& fr = observation_frames_[f];
    os << "{\"observations\":[";
    for (size_t k = 0; k < fr.observations.size(); ++k) {
      const Observation& ob = fr.observations[k];
      if (k) os << ",";
      os << "{\"camera_id\":" << ob.camera_id << ",\"cost\":";
      put_num(os, ob.cost);
      os << ",\"image_point\":[";
      put_num(os, ob.image_point.x());
      os << ",";
      put_num(os, ob.image_point.y());
      os << "],\"world_point_id\":" << ob.world_point_id << "}";
    }
    os << "],\"rig_T_world\":";
    put_transform(os, fr.rig_T_world);
    os << "}";
  }
  os << "],\"world_points\":[";
  for (size_t i = 0; i < world_point_infos_.size(); ++i) {
    const WorldPointInfo& info = world_point_infos_[i];
    const Point3D& p = observation_frames_[info.observation_frame_id].world_points[info.world_point_idx];
    if (i) os << ",";
    os << "{\"frame_id\":" << info.observation_frame_id << ",\"world_point\":[";
    put_num(os, p(0));
    os << ",";
    put_num(os, p(1));
    os << ",";
    put_num(os, p(2));
    os << "]}";
  }
  os << "]}";
}

void ExtrinsicsCalibrator::Parse(const std::string& fname) {
  // As in the reference (:349-351) the camera list is NOT cleared: parsing into a populated
  // object appends the parsed cameras after the existing ones.
  frozen_camera_T_rigs_.clear();
  world_point_infos_.clear();
  observation_frames_.clear();
  std::ifstream in(fname.c_str());
  if (!in) throw std::runtime_error("ExtrinsicsCalibrator::Parse: cannot open " + fname);
  std::stringstream ss;
  ss << in.rdbuf();
  const std::string text = ss.str();
  JParser parser(text);
  const JValue root = parser.parse();
  for (const JValue& c : root.at("camera_T_rigs").arr)
    AddCameraTRig(get_transform(c.at("camera_T_rig")), c.at("frozen").b);
  for (const JValue& f : root.at("observation_frames").arr) AddObservationFrame(get_transform(f.at("rig_T_world")));
  for (const JValue& w : root.at("world_points").arr) {
    const JValue& p = w.at("world_point");
    AddWorldPoint((size_t)w.at("frame_id").num, Point3D((float)p.arr.at(0).num, (float)p.arr.at(1).num, (float)p.arr.at(2).num));
  }
  for (const JValue& f : root.at("observation_frames").arr)
    for (const JValue& ob : f.at("observations").arr) {
      const JValue& ip = ob.at("image_point");
      AddObservation((size_t)ob.at("camera_id").num, (size_t)ob.at("world_point_id").num,
                     Point2D((float)ip.arr.at(0).num, (float)ip.arr.at(1).num));
    }
}

void ExtrinsicsCalibrator::RemoveObservationFrame(const size_t observation_frame_id) {
  const size_t num_world_points = observation_frames_.at(observation_frame_id).world_points.size();
  observation_frames_.erase(observation_frames_.begin() + (std::ptrdiff_t)observation_frame_id);
  // world point ids of later frames move down by the removed frame's point count
  for (size_t f = observation_frame_id; f < observation_frames_.size(); ++f)
    for (Observation& o : observation_frames_[f].observations) o.world_point_id -= num_world_points;
  world_point_infos_.erase(std::remove_if(world_point_infos_.begin(), world_point_infos_.end(),
                                          [=](const WorldPointInfo& wp) {
                                            return wp.observation_frame_id == observation_frame_id;
                                          }),
                           world_point_infos_.end());
  for (WorldPointInfo& wp : world_point_infos_)
    if (wp.observation_frame_id >= observation_frame_id) --wp.observation_frame_id;
}

void ExtrinsicsCalibrator::RemoveObservationFrames(const std::vector<size_t> observation_frame_ids) {
  std::vector<size_t> ids = observation_frame_ids;
  std::sort(ids.begin(), ids.end(), std::greater<size_t>());  // larger ids first: the others stay valid
  for (size_t id : ids) RemoveObservationFrame(id);
}

}  // namespace calibrator
