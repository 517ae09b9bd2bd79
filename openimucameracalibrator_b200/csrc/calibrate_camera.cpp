// Drop-in host program for applications/calibrate_camera.cc of urbste/OpenImuCameraCalibrator: same
// gflags names and defaults (app :25-45), same input (UBJSON corner file of extract_board_to_json), same outputs:
//   <save_path_calib_dataset>.json          camera calibration in the layout of io::write_camera_calibration
//                                           (src/io/write_camera_calibration.cc:33-133) -- what every other tool reads back
//   <save_path_calib_dataset>.calibdata     the calibration dataset (views + board points)
// The calibration itself -- CameraCalibrator::CalibrateCameraFromJson / RunCalibration (src/core/camera_calibrator.cc:131-389) --
// runs on the B200 through icc_calibrate_camera; PrintResult (:391-458) is reproduced on stdout.
//
// Deliberate differences: the .calibdata file is the JSON pose dataset this repository's tools exchange ({"views": {"<name>":
// {"q_wc": [w,x,y,z], "p_wc": [...], "timestamp_s": t, "mean_reproj_error": e}}, "tracks": {...}}), not Theia's cereal binary
// (unreadable without Theia); the *_ransac_poses.ply / *_final_poses.ply point clouds are not written.  --optimize_board_points
// (camera_calibrator.cc:207-216) refines the board points on the GPU and writes them as the dataset's tracks.
// Extra flag: --device (CUDA ordinal, default 0).
#include "../../include/icc_b200.h"
#include "icc_cli_common.hpp"

#include <cmath>
#include <fstream>
#include <iostream>

using iccjson::Value;

namespace {
int model_from_string(const std::string& s) {   // theia::StringToCameraIntrinsicsModelType for the names the app documents (:27-31)
  if (s == "PINHOLE") return ICC_CAM_PINHOLE;
  if (s == "PINHOLE_RADIAL_TANGENTIAL") return ICC_CAM_PINHOLE_RADIAL_TANGENTIAL;
  if (s == "DIVISION_UNDISTORTION") return ICC_CAM_DIVISION_UNDISTORTION;
  if (s == "DOUBLE_SPHERE") return ICC_CAM_DOUBLE_SPHERE;
  if (s == "EXTENDED_UNIFIED") return ICC_CAM_EXTENDED_UNIFIED;
  if (s == "FISHEYE") return ICC_CAM_FISHEYE;
  if (s == "FOV") return ICC_CAM_FOV;
  return -1;
}
}  // namespace

int main(int argc, char** argv) {
  icccli::Flags F;
  F.str = {{"input_corners", ""}, {"camera_model_to_calibrate", "DOUBLE_SPHERE"}, {"save_path_calib_dataset", ""}};
  F.boolean = {{"optimize_board_points", false}, {"verbose", false}};
  F.num = {{"grid_size", 0.04}, {"device", 0.0}};
  try { icccli::parse_flags(argc, argv, F); } catch (const std::exception& e) { std::cerr << "ERROR: " << e.what() << std::endl; return 1; }
  try {
    Value scene_json; icccli::SceneViews sv;
    try { scene_json = icccli::load_scene(F.str["input_corners"], sv); } catch (const std::exception& e) { std::cerr << "Check failed: Failed to load " << F.str["input_corners"] << ": " << e.what() << std::endl; return 1; }
    const std::string model_name = F.str["camera_model_to_calibrate"];
    const int model = model_from_string(model_name);
    if (model < 0) { std::cerr << "ERROR: unknown camera model '" << model_name << "'" << std::endl; return 1; }
    const int width = (int)scene_json.at("image_width").num(), height = (int)scene_json.at("image_height").num();
    const double fps = scene_json.at("camera_fps").num();
    int np = 0;
    const std::vector<double> board = icccli::read_scene_points(scene_json, np);
    const int nv = (int)sv.timestamp_us.size();
    if (nv == 0) { std::cerr << "Check failed: the corner file holds no views" << std::endl; return 1; }
    icc_handle* h = nullptr;
    icc_status st = icc_create(&h, (int)F.num["device"]);
    if (st != ICC_OK) { std::cerr << "icc_create failed (" << st << "): " << icc_last_error(h) << std::endl; return 2; }
    std::vector<double> q(4 * (size_t)nv), p(3 * (size_t)nv), err(nv), intr(10, 0.0); std::vector<int32_t> used(nv);
    icc_camcal_options opt{}; opt.grid_size = F.num["grid_size"]; opt.optimize_board_points = F.boolean["optimize_board_points"] ? 1 : 0;
    icc_camcal_summary S{};
    st = icc_set_board_points(h, np, board.data());
    if (st == ICC_OK) st = icc_calibrate_camera(h, model, width, height, nv, sv.off.data(), sv.ids.data(), sv.uv.data(), nullptr, nullptr, nullptr, 0.0, 0.0, &opt,
                                                intr.data(), q.data(), p.data(), err.data(), used.data(), &S);
    if (st != ICC_OK) { std::cerr << "camera calibration failed (" << st << "): " << icc_last_error(h) << std::endl; return 2; }
    std::vector<double> board_out(board);
    if (opt.optimize_board_points) icc_get_board_points(h, board_out.data(), np);
    icc_destroy(h);
    std::cout << "Using " << S.n_views_selected << " views for camera calibration.\n";
    if (F.boolean["verbose"]) for (int i = 0; i < nv; ++i) if (used[i]) std::cout << "View: " << icccli::pose_view_name(sv.timestamp_us[i]) << " RMSE reprojection error: " << err[i] << "\n";
    if (opt.optimize_board_points && S.success) std::cout << "Optimized " << S.n_points_optimized << " board points.\n";
    if (!S.success) { std::cerr << "Not enough views for proper calibration!\nCalibration failed.\n"; return 3; }
    std::cout << "Final camera calibration reprojection error: " << S.final_reproj_error << " from " << S.n_views_used << " view." << std::endl;
    const bool noskew = model == ICC_CAM_FOV || model == ICC_CAM_DIVISION_UNDISTORTION;
    const double cx = noskew ? intr[2] : intr[3], cy = noskew ? intr[3] : intr[4];
    const std::string out = F.str["save_path_calib_dataset"];
    if (!out.empty()) {
      // io::write_camera_calibration (src/io/write_camera_calibration.cc:33-133)
      Value j = Value::object(), in = Value::object();
      j["stabelized"] = Value(false); j["fps"] = Value(fps); j["nr_calib_images"] = Value((int64_t)S.n_views_used); j["final_reproj_error"] = Value(S.final_reproj_error);
      j["image_width"] = Value((int64_t)width); j["image_height"] = Value((int64_t)height);   // integers, like io::write_camera_calibration
      j["intrinsic_type"] = Value(model_name);
      in["skew"] = Value(0.0); in["principal_pt_x"] = Value(cx); in["principal_pt_y"] = Value(cy); in["aspect_ratio"] = Value(intr[1]); in["focal_length"] = Value(intr[0]);
      switch (model) {
        case ICC_CAM_DIVISION_UNDISTORTION: in["div_undist_distortion"] = Value(intr[4]); break;
        case ICC_CAM_DOUBLE_SPHERE: in["xi"] = Value(intr[5]); in["alpha"] = Value(intr[6]); break;
        case ICC_CAM_EXTENDED_UNIFIED: in["alpha"] = Value(intr[5]); in["beta"] = Value(intr[6]); break;
        case ICC_CAM_FISHEYE: for (int i = 0; i < 4; ++i) in["radial_distortion_" + std::to_string(i + 1)] = Value(intr[5 + i]); break;
        case ICC_CAM_PINHOLE_RADIAL_TANGENTIAL: for (int i = 0; i < 3; ++i) in["radial_distortion_" + std::to_string(i + 1)] = Value(intr[5 + i]);
                                                in["tangential_distortion_1"] = Value(intr[8]); in["tangential_distortion_2"] = Value(intr[9]); break;
        case ICC_CAM_FOV: in["radial_distortion_1"] = Value(intr[4]); break;   // extension: the layout this repository's readers use for FOV
        default: break;
      }
      j["intrinsics"] = in;
      std::ofstream f(out + ".json");
      if (!f.is_open()) { std::cerr << "Could not open: " << out << ".json\nCheck failed: Could not write calibration file.\n"; return 1; }
      f << iccjson::dump(j, 2) << std::endl;
      Value ds = Value::object(); ds["views"] = Value::object(); ds["tracks"] = Value::object();
      for (int i = 0; i < nv; ++i) {
        if (!used[i]) continue;
        Value v = Value::object();
        v["q_wc"] = Value::array(); v["q_wc"].push_back(Value(q[4 * i + 3])); for (int d = 0; d < 3; ++d) v["q_wc"].push_back(Value(q[4 * i + d]));
        v["p_wc"] = Value::array(); for (int d = 0; d < 3; ++d) v["p_wc"].push_back(Value(p[3 * i + d]));
        v["timestamp_s"] = Value(sv.timestamp_us[i] * 1e-6); v["mean_reproj_error"] = Value(err[i]);
        ds["views"][icccli::pose_view_name(sv.timestamp_us[i])] = v;
      }
      for (int i = 0; i < np; ++i) { Value t = Value::array(); for (int d = 0; d < 4; ++d) t.push_back(Value(board_out[4 * i + d])); ds["tracks"][std::to_string(i)] = t; }
      std::ofstream g(out + ".calibdata");
      if (!g.is_open()) { std::cerr << "could not write " << out << ".calibdata" << std::endl; return 1; }
      g << iccjson::dump(ds, 1) << std::endl;
      // theia::WritePlyFile(output_path + "_final_poses.ply", recon, camera colour (255, 0, 0), min 1 observation) (camera_calibrator.cc:381-384):
      // the camera positions of the views that took part, then the board points.  (The reference also dumps "_ransac_poses.ply" before the
      // calibration, :342-345; there is no RANSAC stage here -- the initial poses never leave the device.)
      std::ofstream ply(out + "_final_poses.ply");
      if (!ply.is_open()) { std::cerr << "could not write " << out << "_final_poses.ply" << std::endl; return 1; }
      int n_cam = 0; for (int i = 0; i < nv; ++i) n_cam += used[i] ? 1 : 0;
      ply << "ply\nformat ascii 1.0\nelement vertex " << n_cam + np << "\nproperty float x\nproperty float y\nproperty float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n";
      for (int i = 0; i < nv; ++i) if (used[i]) ply << p[3 * i] << " " << p[3 * i + 1] << " " << p[3 * i + 2] << " 255 0 0\n";
      for (int i = 0; i < np; ++i) { const double w = board_out[4 * i + 3]; ply << board_out[4 * i] / w << " " << board_out[4 * i + 1] / w << " " << board_out[4 * i + 2] / w << " 255 255 255\n"; }
    }
    // CameraCalibrator::PrintResult (:391-458)
    std::cout << "Focal Length:" << intr[0] << "px Principal Point: " << cx << "/" << cy << "px.\n";
    if (model == ICC_CAM_DIVISION_UNDISTORTION) std::cout << "DIVISION_UNDISTORTION model: Distortion: " << intr[4] << "\n";
    else if (model == ICC_CAM_DOUBLE_SPHERE) std::cout << "DOUBLE_SPHERE model: XI: " << intr[5] << " ALPHA: " << intr[6] << "\n";
    else if (model == ICC_CAM_EXTENDED_UNIFIED) std::cout << "EXTENDED_UNIFIED model: " << intr[5] << " BETA: " << intr[6] << "\n";
    else if (model == ICC_CAM_FISHEYE) std::cout << "FISHEYE model: Radial distortion 1: " << intr[5] << " Radial distortion 2: " << intr[6] << " Radial distortion 3: " << intr[7] << " Radial distortion 4: " << intr[8] << "\n";
  } catch (const std::exception& e) { std::cerr << "ERROR: " << e.what() << std::endl; return 1; }
  return 0;
}
