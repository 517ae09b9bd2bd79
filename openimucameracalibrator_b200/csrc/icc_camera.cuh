// Device camera-model "plugin surface": project a camera-frame point to pixels and return the analytic 2x3 Jacobian.
//
// Replaces theia::XCameraModel::CameraToPixelCoordinates<T>(intr, pt3, px) (pyTheiaSfM@69c3d37, external to
// /root/reference) as dispatched at include/OpenCameraCalibrator/basalt_spline/ceres_calib_split_residuals.h:366-389,
// where the derivative came from Ceres Jets.  Intrinsic index order = Theia's InternalParametersIndex (see icc_b200.h).
// Return value = Theia's `bool` (false => the reference writes the constant 1e10 residual, quirk q12).
#pragma once
#include "icc_device_math.cuh"

namespace icc {

enum CameraModel { CAM_PINHOLE = 0, CAM_PINHOLE_RADTAN = 1, CAM_FISHEYE = 2, CAM_FOV = 3, CAM_DIVISION_UNDISTORTION = 4, CAM_DOUBLE_SPHERE = 5, CAM_EXTENDED_UNIFIED = 6 };

ICC_HD int camera_num_params(int model) {
  switch (model) { case CAM_PINHOLE: return 7; case CAM_PINHOLE_RADTAN: return 10; case CAM_FISHEYE: return 9; case CAM_FOV: return 5;
                   case CAM_DIVISION_UNDISTORTION: return 5; case CAM_DOUBLE_SPHERE: return 7; case CAM_EXTENDED_UNIFIED: return 7; default: return -1; }
}

// J is row-major 2x3: J[0..2] = d px / d(x,y,z), J[3..5] = d py / d(x,y,z).
// Jk (only filled when requested) is row-major 2 x 10: d(u,v)/d intrinsics in Theia's parameter order, unused entries zero.
struct Proj { double u, v; double J[6]; bool ok; };
struct ProjK { double Jk[20]; };

// final affine of the skewed models: u = f dx + s dy + cx, v = f ar dy + cy, given d(dx,dy)/dp
ICC_HD void affine_skew(const double* k, double dx, double dy, const double* Dd /*2x3*/, Proj& o) {
  const double f = k[0], fy = k[0] * k[1], s = k[2];
  o.u = f * dx + s * dy + k[3];
  o.v = fy * dy + k[4];
  for (int j = 0; j < 3; ++j) { o.J[j] = f * Dd[j] + s * Dd[3 + j]; o.J[3 + j] = fy * Dd[3 + j]; }
}
// intrinsic Jacobian of the skewed models: [f, ar, skew, cx, cy] columns + distortion columns idx[q] with d(dx,dy)/dk_q = (ddx[q], ddy[q])
ICC_HD void affine_skew_k(const double* k, double dx, double dy, int nd, const int* idx, const double* ddx, const double* ddy, ProjK& o) {
  for (int j = 0; j < 20; ++j) o.Jk[j] = 0.0;
  o.Jk[0] = dx; o.Jk[10] = k[1] * dy;          // f
  o.Jk[11] = k[0] * dy;                        // aspect ratio
  o.Jk[2] = dy;                                // skew
  o.Jk[3] = 1.0; o.Jk[14] = 1.0;               // principal point
  for (int q = 0; q < nd; ++q) { o.Jk[idx[q]] = k[0] * ddx[q] + k[2] * ddy[q]; o.Jk[10 + idx[q]] = k[0] * k[1] * ddy[q]; }
}

ICC_HD double unified_w(double alpha) { return alpha > 0.5 ? (1.0 - alpha) / alpha : alpha / (1.0 - alpha); }

template <int MODEL, bool WITH_K = false>
ICC_HD Proj project_model(const double* k, V3 p, bool dispatch_fov, ProjK* pk = nullptr) {
  Proj o; o.ok = true;
  if (WITH_K) for (int j = 0; j < 20; ++j) pk->Jk[j] = 0.0;
  if (MODEL == CAM_PINHOLE || MODEL == CAM_PINHOLE_RADTAN) {
    const double iz = 1.0 / p.z, xn = p.x * iz, yn = p.y * iz;
    const double r2 = xn * xn + yn * yn;
    double d, dd, dx, dy, a00, a01, a10, a11;   // a = d(dx,dy)/d(xn,yn)
    if (MODEL == CAM_PINHOLE) {
      d = 1.0 + r2 * (k[5] + k[6] * r2); dd = k[5] + 2.0 * k[6] * r2;
      dx = xn * d; dy = yn * d;
      a00 = d + 2.0 * xn * xn * dd; a01 = 2.0 * xn * yn * dd; a10 = a01; a11 = d + 2.0 * yn * yn * dd;
    } else {
      const double k1 = k[5], k2 = k[6], k3 = k[7], t1 = k[8], t2 = k[9];
      d = 1.0 + r2 * (k1 + r2 * (k2 + r2 * k3)); dd = k1 + r2 * (2.0 * k2 + 3.0 * k3 * r2);
      dx = xn * d + 2.0 * t1 * xn * yn + t2 * (r2 + 2.0 * xn * xn);
      dy = yn * d + 2.0 * t2 * xn * yn + t1 * (r2 + 2.0 * yn * yn);
      a00 = d + 2.0 * xn * xn * dd + 2.0 * t1 * yn + 6.0 * t2 * xn;
      a01 = 2.0 * xn * yn * dd + 2.0 * t1 * xn + 2.0 * t2 * yn;
      a10 = 2.0 * xn * yn * dd + 2.0 * t2 * yn + 2.0 * t1 * xn;
      a11 = d + 2.0 * yn * yn * dd + 2.0 * t2 * xn + 6.0 * t1 * yn;
    }
    // d(xn,yn)/dp = [[iz,0,-xn iz],[0,iz,-yn iz]]
    double Dd[6] = {a00 * iz, a01 * iz, -(a00 * xn + a01 * yn) * iz, a10 * iz, a11 * iz, -(a10 * xn + a11 * yn) * iz};
    affine_skew(k, dx, dy, Dd, o);
    if (WITH_K) {
      if (MODEL == CAM_PINHOLE) { const int idx[2] = {5, 6}; const double ddx[2] = {xn * r2, xn * r2 * r2}, ddy[2] = {yn * r2, yn * r2 * r2}; affine_skew_k(k, dx, dy, 2, idx, ddx, ddy, *pk); }
      else { const int idx[5] = {5, 6, 7, 8, 9};
             const double ddx[5] = {xn * r2, xn * r2 * r2, xn * r2 * r2 * r2, 2.0 * xn * yn, r2 + 2.0 * xn * xn}, ddy[5] = {yn * r2, yn * r2 * r2, yn * r2 * r2 * r2, r2 + 2.0 * yn * yn, 2.0 * xn * yn};
             affine_skew_k(k, dx, dy, 5, idx, ddx, ddy, *pk); }
    }
  } else if (MODEL == CAM_FISHEYE) {
    const double r2 = p.x * p.x + p.y * p.y;
    if (r2 < 1e-8) {
      double Dd[6] = {1, 0, 0, 0, 1, 0};
      affine_skew(k, p.x, p.y, Dd, o);
      if (WITH_K) affine_skew_k(k, p.x, p.y, 0, nullptr, nullptr, nullptr, *pk);
    } else {
      const double r = sqrt(r2), az = fabs(p.z), sg = p.z < 0.0 ? -1.0 : 1.0;
      const double th = atan2(r, az), th2 = th * th;
      const double thd = th * (1.0 + th2 * (k[5] + th2 * (k[6] + th2 * (k[7] + th2 * k[8]))));
      const double dthd = 1.0 + th2 * (3.0 * k[5] + th2 * (5.0 * k[6] + th2 * (7.0 * k[7] + th2 * 9.0 * k[8])));
      const double rho2 = r2 + p.z * p.z;
      const double g = thd / r;
      // dg/dx = (dthd * az/rho2 * x/r) / r - thd x / r^3 ; dg/dz = -dthd * sign(z) / rho2
      const double common = (dthd * az / rho2 - g) / r2;
      const double gx = common * p.x, gy = common * p.y, gz = -dthd * sg / rho2;
      double Dd[6] = {sg * (g + p.x * gx), sg * (p.x * gy), sg * (p.x * gz), sg * (p.y * gx), sg * (g + p.y * gy), sg * (p.y * gz)};
      affine_skew(k, sg * g * p.x, sg * g * p.y, Dd, o);
      if (WITH_K) { const int idx[4] = {5, 6, 7, 8}; const double t3 = th * th2, ex = sg * p.x / r, ey = sg * p.y / r;   // d theta_d / d k_i = theta^(2i+1)
                    const double ddx[4] = {t3 * ex, t3 * th2 * ex, t3 * th2 * th2 * ex, t3 * th2 * th2 * th2 * ex}, ddy[4] = {t3 * ey, t3 * th2 * ey, t3 * th2 * th2 * ey, t3 * th2 * th2 * th2 * ey};
                    affine_skew_k(k, sg * g * p.x, sg * g * p.y, 4, idx, ddx, ddy, *pk); }
    }
  } else if (MODEL == CAM_FOV) {
    if (!dispatch_fov) { o.ok = false; o.u = o.v = 0; for (int j = 0; j < 6; ++j) o.J[j] = 0; return o; }
    const double iz = 1.0 / p.z, xn = p.x * iz, yn = p.y * iz;
    const double r2 = xn * xn + yn * yn, om = k[4];
    double s, ds_r;  // s(r), (ds/dr)/r
    if (om * om < 1e-10) { s = 1.0; ds_r = 0.0; }
    else if (r2 < 1e-10) { s = 2.0 * tan(0.5 * om) / om; ds_r = 0.0; }
    else {
      const double r = sqrt(r2), T = tan(0.5 * om), at = atan(2.0 * r * T);
      s = at / (om * r);
      ds_r = ((2.0 * T / (1.0 + 4.0 * r2 * T * T)) * r - at) / (om * r2) / r;
    }
    const double a00 = s + xn * xn * ds_r, a01 = xn * yn * ds_r, a11 = s + yn * yn * ds_r;
    const double f = k[0], fy = k[0] * k[1];
    o.u = f * s * xn + k[2]; o.v = fy * s * yn + k[3];
    o.J[0] = f * a00 * iz; o.J[1] = f * a01 * iz; o.J[2] = -f * (a00 * xn + a01 * yn) * iz;
    o.J[3] = fy * a01 * iz; o.J[4] = fy * a11 * iz; o.J[5] = -fy * (a01 * xn + a11 * yn) * iz;
    if (WITH_K) {   // [f, ar, cx, cy, omega]
      double ds_om = 0.0;
      if (!(om * om < 1e-10)) {
        const double T = tan(0.5 * om), dT = 0.5 * (1.0 + T * T);
        if (r2 < 1e-10) ds_om = 2.0 * dT / om - 2.0 * T / (om * om);
        else { const double r = sqrt(r2), at = atan(2.0 * r * T); ds_om = (2.0 * r * dT / (1.0 + 4.0 * r2 * T * T)) / (om * r) - at / (om * om * r); }
      }
      pk->Jk[0] = s * xn; pk->Jk[10] = k[1] * s * yn; pk->Jk[11] = f * s * yn; pk->Jk[2] = 1.0; pk->Jk[13] = 1.0; pk->Jk[4] = f * ds_om * xn; pk->Jk[14] = fy * ds_om * yn;
    }
  } else if (MODEL == CAM_DIVISION_UNDISTORTION) {
    const double iz = 1.0 / p.z, f = k[0], fy = k[0] * k[1], kd = k[4];
    const double xu = f * p.x * iz, yu = fy * p.y * iz;
    const double r2 = xu * xu + yu * yu;
    const double den = 2.0 * kd * r2, inner = 1.0 - 4.0 * kd * r2;
    double s = 1.0, ds = 0.0;   // scale and d(scale)/d(r2)
    if (!(fabs(den) < 1e-15 || inner < 0.0)) {
      const double w = sqrt(inner);
      s = (1.0 - w) / den;
      ds = (den / w - (1.0 - w)) / (den * r2);
    }
    o.u = xu * s + k[2]; o.v = yu * s + k[3];
    const double a00 = s + 2.0 * xu * xu * ds, a01 = 2.0 * xu * yu * ds, a11 = s + 2.0 * yu * yu * ds;
    // d(xu)/dp = f [iz, 0, -x iz^2], d(yu)/dp = fy [0, iz, -y iz^2]
    const double xz = -xu * iz, yz = -yu * iz;
    o.J[0] = a00 * f * iz; o.J[1] = a01 * fy * iz; o.J[2] = a00 * xz + a01 * yz;
    o.J[3] = a01 * f * iz; o.J[4] = a11 * fy * iz; o.J[5] = a01 * xz + a11 * yz;
    if (WITH_K) {   // [f, ar, cx, cy, k]: xu, yu scale with f, yu with ar; d(scale)/dk = (den/w - (1-w)) / (k den)
      double dsk = 0.0;
      if (!(fabs(den) < 1e-15 || inner < 0.0)) { const double w = sqrt(inner); dsk = (den / w - (1.0 - w)) / (kd * den); }
      pk->Jk[0] = (a00 * xu + a01 * yu) / f; pk->Jk[10] = (a01 * xu + a11 * yu) / f;
      pk->Jk[1] = a01 * yu / k[1]; pk->Jk[11] = a11 * yu / k[1];
      pk->Jk[2] = 1.0; pk->Jk[13] = 1.0;
      pk->Jk[4] = xu * dsk; pk->Jk[14] = yu * dsk;
    }
  } else if (MODEL == CAM_DOUBLE_SPHERE) {
    const double xi = k[5], al = k[6];
    const double r2 = p.x * p.x + p.y * p.y;
    const double d1 = sqrt(r2 + p.z * p.z);
    const double w1 = unified_w(al), w2 = (w1 + xi) / sqrt(2.0 * w1 * xi + xi * xi + 1.0);
    if (p.z <= -w2 * d1) { o.ok = false; o.u = o.v = 0; for (int j = 0; j < 6; ++j) o.J[j] = 0; return o; }
    const double kk = xi * d1 + p.z;
    const double d2 = sqrt(r2 + kk * kk);
    const double nrm = al * d2 + (1.0 - al) * kk, in = 1.0 / nrm;
    // dk/dp = xi p/d1 + ez ; dd2/dp = (x, y, 0)/d2 + kk/d2 dk/dp ; dn = al dd2 + (1-al) dk
    const double kx = xi * p.x / d1, ky = xi * p.y / d1, kz = xi * p.z / d1 + 1.0;
    const double c = al * kk / d2 + (1.0 - al);
    const double nx = al * p.x / d2 + c * kx, ny = al * p.y / d2 + c * ky, nz = c * kz;
    const double dx = p.x * in, dy = p.y * in;
    double Dd[6] = {in - dx * nx * in, -dx * ny * in, -dx * nz * in, -dy * nx * in, in - dy * ny * in, -dy * nz * in};
    affine_skew(k, dx, dy, Dd, o);
    if (WITH_K) {   // d norm / d xi = d1 (al kk / d2 + 1 - al),  d norm / d alpha = d2 - kk ;  d(dx)/d. = -dx / norm * d norm
      const int idx[2] = {5, 6};
      const double dn[2] = {d1 * c, d2 - kk};
      const double ddx[2] = {-dx * in * dn[0], -dx * in * dn[1]}, ddy[2] = {-dy * in * dn[0], -dy * in * dn[1]};
      affine_skew_k(k, dx, dy, 2, idx, ddx, ddy, *pk);
    }
  } else {  // CAM_EXTENDED_UNIFIED
    const double al = k[5], be = k[6];
    const double r2 = p.x * p.x + p.y * p.y;
    const double rho = sqrt(be * r2 + p.z * p.z);
    const double nrm = al * rho + (1.0 - al) * p.z, in = 1.0 / nrm;
    if (p.z <= -unified_w(al) * rho) { o.ok = false; o.u = o.v = 0; for (int j = 0; j < 6; ++j) o.J[j] = 0; return o; }
    const double nx = al * be * p.x / rho, ny = al * be * p.y / rho, nz = al * p.z / rho + (1.0 - al);
    const double dx = p.x * in, dy = p.y * in;
    double Dd[6] = {in - dx * nx * in, -dx * ny * in, -dx * nz * in, -dy * nx * in, in - dy * ny * in, -dy * nz * in};
    affine_skew(k, dx, dy, Dd, o);
    if (WITH_K) {   // d norm / d alpha = rho - z,  d norm / d beta = alpha r2 / (2 rho)
      const int idx[2] = {5, 6};
      const double dn[2] = {rho - p.z, al * r2 / (2.0 * rho)};
      const double ddx[2] = {-dx * in * dn[0], -dx * in * dn[1]}, ddy[2] = {-dy * in * dn[0], -dy * in * dn[1]};
      affine_skew_k(k, dx, dy, 2, idx, ddx, ddy, *pk);
    }
  }
  return o;
}

// Same dispatch, additionally returning d(u,v)/d(intrinsics) (CAM_INTRINSICS extension).
ICC_HD Proj project_with_k(int model, const double* k, V3 p, bool dispatch_fov, ProjK* pk) {
  switch (model) {
    case CAM_DIVISION_UNDISTORTION: return project_model<CAM_DIVISION_UNDISTORTION, true>(k, p, dispatch_fov, pk);
    case CAM_DOUBLE_SPHERE: return project_model<CAM_DOUBLE_SPHERE, true>(k, p, dispatch_fov, pk);
    case CAM_PINHOLE: return project_model<CAM_PINHOLE, true>(k, p, dispatch_fov, pk);
    case CAM_FISHEYE: return project_model<CAM_FISHEYE, true>(k, p, dispatch_fov, pk);
    case CAM_EXTENDED_UNIFIED: return project_model<CAM_EXTENDED_UNIFIED, true>(k, p, dispatch_fov, pk);
    case CAM_PINHOLE_RADTAN: return project_model<CAM_PINHOLE_RADTAN, true>(k, p, dispatch_fov, pk);
    case CAM_FOV: return project_model<CAM_FOV, true>(k, p, dispatch_fov, pk);
    default: { Proj o; o.ok = false; o.u = o.v = 0; for (int j = 0; j < 6; ++j) o.J[j] = 0; for (int j = 0; j < 20; ++j) pk->Jk[j] = 0; return o; }
  }
}

// Runtime dispatch in the reference's order (residuals.h:366-389).
ICC_HD Proj project(int model, const double* k, V3 p, bool dispatch_fov) {
  switch (model) {
    case CAM_DIVISION_UNDISTORTION: return project_model<CAM_DIVISION_UNDISTORTION>(k, p, dispatch_fov);
    case CAM_DOUBLE_SPHERE: return project_model<CAM_DOUBLE_SPHERE>(k, p, dispatch_fov);
    case CAM_PINHOLE: return project_model<CAM_PINHOLE>(k, p, dispatch_fov);
    case CAM_FISHEYE: return project_model<CAM_FISHEYE>(k, p, dispatch_fov);
    case CAM_EXTENDED_UNIFIED: return project_model<CAM_EXTENDED_UNIFIED>(k, p, dispatch_fov);
    case CAM_PINHOLE_RADTAN: return project_model<CAM_PINHOLE_RADTAN>(k, p, dispatch_fov);
    case CAM_FOV: return project_model<CAM_FOV>(k, p, dispatch_fov);
    default: { Proj o; o.ok = false; o.u = o.v = 0; for (int j = 0; j < 6; ++j) o.J[j] = 0; return o; }
  }
}

// theia::Camera::PixelToNormalizedCoordinates(px) / z : solve project(x, y, 1) = px by damped Gauss-Newton
ICC_HD bool unproject_gn(int model, const double* k, double px, double py, double& xo, double& yo) {
  const Proj p0 = project(model, k, v3(0.0, 0.0, 1.0), true);
  if (!p0.ok) return false;
  double x, y;
  {
    const double a = p0.J[0], b = p0.J[1], c = p0.J[3], d = p0.J[4], det = a * d - b * c;
    if (!(fabs(det) > 0.0)) return false;
    const double du = px - p0.u, dv = py - p0.v;
    x = (d * du - b * dv) / det; y = (-c * du + a * dv) / det;
  }
  Proj p = project(model, k, v3(x, y, 1.0), true);
  for (int s = 0; s < 60 && !p.ok; ++s) { x *= 0.5; y *= 0.5; p = project(model, k, v3(x, y, 1.0), true); }   // outside the model's domain
  if (!p.ok) return false;
  double ru = p.u - px, rv = p.v - py, e = ru * ru + rv * rv;
  for (int it = 0; it < 50; ++it) {
    if (e < 1e-26) break;
    const double a = p.J[0], b = p.J[1], c = p.J[3], d = p.J[4], det = a * d - b * c;
    if (!(fabs(det) > 1e-300)) break;
    const double dx = -(d * ru - b * rv) / det, dy = -(-c * ru + a * rv) / det;
    double t = 1.0; bool moved = false;
    for (int bt = 0; bt < 30; ++bt, t *= 0.5) {
      const double xn = x + t * dx, yn = y + t * dy;
      const Proj pn = project(model, k, v3(xn, yn, 1.0), true);
      if (!pn.ok) continue;
      const double r0 = pn.u - px, r1 = pn.v - py, en = r0 * r0 + r1 * r1;
      if (en < e) { x = xn; y = yn; p = pn; ru = r0; rv = r1; e = en; moved = true; break; }
    }
    if (!moved) break;
  }
  xo = x; yo = y;
  return e < 1e-12;      // (1e-6 px)^2: anything worse did not converge
}

}  // namespace icc
