// Pieces shared by the drop-in command line tools: gflags-style parsing, the camera calibration JSON reader and the corner file
// (scene) reader.  Header-only, host C++.
#pragma once
#include "../../include/icc_b200.h"
#include "icc_json.hpp"

#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace icccli {
using iccjson::Value;

struct Flags {   // every tool fills in its own names and defaults
  std::map<std::string, std::string> str;
  std::map<std::string, bool> boolean;
  std::map<std::string, double> num;
};

inline bool parse_bool(const std::string& v) { return v == "true" || v == "1" || v == "t" || v == "yes" || v == "y" || v == "True"; }

// gflags syntax: --name=value, --name value, -name..., --boolflag, --noboolflag
inline void parse_flags(int argc, char** argv, Flags& f) {
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a.size() < 2 || a[0] != '-') throw std::runtime_error("unexpected argument: " + a);
    a = a.substr(a[1] == '-' ? 2 : 1);
    std::string name = a, value; bool has_value = false;
    const size_t eq = a.find('=');
    if (eq != std::string::npos) { name = a.substr(0, eq); value = a.substr(eq + 1); has_value = true; }
    if (f.boolean.count(name)) { f.boolean[name] = has_value ? parse_bool(value) : true; continue; }
    if (name.rfind("no", 0) == 0 && f.boolean.count(name.substr(2)) && !has_value) { f.boolean[name.substr(2)] = false; continue; }
    if (!has_value) { if (i + 1 >= argc) throw std::runtime_error("flag --" + name + " needs a value"); value = argv[++i]; }
    if (f.str.count(name)) f.str[name] = value;
    else if (f.num.count(name)) f.num[name] = std::stod(value);
    else throw std::runtime_error("unknown command line flag '" + name + "'");
  }
}


inline // src/io/read_camera_calibration.cc:35-119 -> (model id, Theia-ordered intrinsics).  Like the reference, `skew` is never read.
int read_camera(const Value& j, std::vector<double>& k, int& width, int& height, double& fps) {
  const std::string type = j.at("intrinsic_type").str();
  const Value& in = j.at("intrinsics");
  width = (int)j.at("image_width").num(); height = (int)j.at("image_height").num(); fps = j.at("fps").num();
  const double f = in.at("focal_length").num(), cx = in.at("principal_pt_x").num(), cy = in.at("principal_pt_y").num();
  auto ar = [&]() { return in.at("aspect_ratio").num(); };
  if (type == "DIVISION_UNDISTORTION") { k = {f, ar(), cx, cy, in.at("div_undist_distortion").num()}; return ICC_CAM_DIVISION_UNDISTORTION; }
  if (type == "DOUBLE_SPHERE") { k = {f, ar(), 0.0, cx, cy, in.at("xi").num(), in.at("alpha").num()}; return ICC_CAM_DOUBLE_SPHERE; }
  if (type == "EXTENDED_UNIFIED") { k = {f, ar(), 0.0, cx, cy, in.at("alpha").num(), in.at("beta").num()}; return ICC_CAM_EXTENDED_UNIFIED; }
  if (type == "FISHEYE") { k = {f, ar(), 0.0, cx, cy, in.at("radial_distortion_1").num(), in.at("radial_distortion_2").num(), in.at("radial_distortion_3").num(), in.at("radial_distortion_4").num()}; return ICC_CAM_FISHEYE; }
  if (type == "PINHOLE_RADIAL_TANGENTIAL") { k = {f, ar(), 0.0, cx, cy, in.at("radial_distortion_1").num(), in.at("radial_distortion_2").num(), in.at("radial_distortion_3").num(), in.at("tangential_distortion_1").num(), in.at("tangential_distortion_2").num()}; return ICC_CAM_PINHOLE_RADIAL_TANGENTIAL; }
  if (type == "PINHOLE") { k = {f, ar(), 0.0, cx, cy, 0.0, 0.0}; return ICC_CAM_PINHOLE; }
  if (type == "FOV") { k = {f, in.contains("aspect_ratio") ? ar() : 1.0, cx, cy, in.at("radial_distortion_1").num()}; return ICC_CAM_FOV; }
  throw std::runtime_error("unknown intrinsic_type " + type);
}

// Corner file (src/core/board_extractor.cc:294-296,325-333,375-380): scene_pts {"id": [x,y,z]} -> homogeneous board points by id.
inline std::vector<double> read_scene_points(const Value& scene_json, int& n_points) {
  const Value& pts = scene_json.at("scene_pts");
  int max_id = -1; for (const auto& kv : *pts.o) max_id = std::max(max_id, std::stoi(kv.first));
  std::vector<double> board(4 * (size_t)(max_id + 1), 0.0);
  for (int i = 0; i <= max_id; ++i) board[4 * i + 3] = 1.0;
  for (const auto& kv : *pts.o) { const int id = std::stoi(kv.first); for (int d = 0; d < 3; ++d) board[4 * id + d] = kv.second.at(d).num(); }
  n_points = max_id + 1;
  return board;
}

// Views of the corner file in file order: timestamps [us, as written], CSR corner lists.
struct SceneViews { std::vector<double> timestamp_us; std::vector<int32_t> off{0}, ids; std::vector<double> uv; };
inline SceneViews read_scene_views(const Value& scene_json) {
  SceneViews v;
  for (const auto& kv : *scene_json.at("views").o) {
    v.timestamp_us.push_back(std::stod(kv.first));
    for (const auto& ip : *kv.second.at("image_points").o) { v.ids.push_back(std::stoi(ip.first)); v.uv.push_back(ip.second.at(0).num()); v.uv.push_back(ip.second.at(1).num()); }
    v.off.push_back((int32_t)v.ids.size());
  }
  return v;
}
// view name in the pose dataset: std::to_string((uint64_t)(timestamp_s * S_TO_US)) with timestamp_s = timestamp_us * US_TO_S (pose_estimator.cc:119-120,146)
inline std::string pose_view_name(double timestamp_us) { const double timestamp_s = timestamp_us * 1e-6; return std::to_string((uint64_t)(timestamp_s * 1e6)); }

}  // namespace icccli
