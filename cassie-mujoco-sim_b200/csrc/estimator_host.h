// estimator_host.h -- host-side part of the reference's estimator (state_output_step, closed source; semantics recovered by probing the
// archive): the toe / heel force of one leg.  Pure function of the measured angles and the IMU quaternion; runs on the host while
// state_out_t rows are unpacked (the kernel's observation row does not carry it in this round).
//
//   toeForce = heelForce = Rz(yaw)' R(q) [f_x, 0, f_z],   (f_x, f_z) = -1/2 J^-T [k_s shin; k_h (H - H0)],   k = (1500, 1250) N m / rad
//
// H = the heel-spring angle that closes the achilles rod (length 0.5012 between (0, 0, 0.045) on the hip-pitch link and the rod end
// (0.11877, -0.01, 0) of the heel-spring frame, which rides on the tarsus as in model/cassie.xml:132); J = Jacobian (pelvis x, z) of the
// foot point w.r.t. the two spring deflections under that closure.  The archive evaluates this in single precision; agreement ~1e-4 relative.
#pragma once
#include <cmath>

namespace cassie {

struct V3 { double x, y, z; };
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 crs(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
struct M3 { V3 c0, c1, c2; };   // columns
inline V3 operator*(const M3 &m, V3 v) { return v.x * m.c0 + v.y * m.c1 + v.z * m.c2; }
inline M3 operator*(const M3 &a, const M3 &b) { return {a * b.c0, a * b.c1, a * b.c2}; }
inline M3 rotz(double t) { const double c = std::cos(t), s = std::sin(t); return {{c, s, 0}, {-s, c, 0}, {0, 0, 1}}; }

// ang: hipRoll, hipYaw, hipPitch, knee (motor positions), shin, tarsus (joint encoders), foot (motor position); quat: IMU quaternion (w, x, y, z)
inline void estimator_leg_force(int side, const double ang[7], const double quat[4], double force[3]) {
  const double sg = side ? -1.0 : 1.0, kn = ang[3], sh = ang[4], ta = ang[5];
  // ---- planar part in the hip-pitch frame (z = common axis of knee, shin, tarsus): the four-bar closure
  const V3 A{0, 0, 0.045 * sg}, k0{0.12, 0, 0.0045 * sg}, o4{0.06068, 0.04741, 0}, o5{0.43476, 0.02, 0}, hsp{-0.01269, -0.03059, 0.00092 * sg}, Bl{0.11877, -0.01, 0}, ez{0, 0, 1};
  V3 hx{-0.91211, 0.40829, 0.036948 * sg}, hy{-0.40992, -0.90952, -0.068841 * sg};
  hx = (1.0 / std::sqrt(dot(hx, hx))) * hx; hy = hy - dot(hx, hy) * hx; hy = (1.0 / std::sqrt(dot(hy, hy))) * hy;
  const M3 HF{hx, hy, crs(hx, hy)};
  const V3 s0 = k0 + rotz(kn) * o4, t0 = s0 + rotz(kn + sh) * o5;
  const M3 R3 = rotz(kn + sh + ta), RH = R3 * HF;
  const V3 hs0 = t0 + R3 * hsp, axh = RH.c2;
  double H = 0, gd = 1; V3 B{}, dB{};
  for (int it = 0; it < 5; it++) {   // Newton on |B - A|^2 = L^2 (quadratic: 5 steps from 0 reach the last bit for |H| < 0.3)
    B = hs0 + RH * (rotz(H) * Bl); dB = B - A;
    gd = 2 * dot(dB, crs(axh, B - hs0));
    if (it < 4) H -= (dot(dB, dB) - 0.5012 * 0.5012) / gd;
  }
  const double a = -2 * dot(dB, crs(ez, B - t0)) / gd, b = -2 * dot(dB, crs(ez, B - s0)) / gd;   // dH/dtarsus, dH/dshin
  // ---- serial chain pelvis -> hip roll -> hip yaw -> hip pitch: frame A2 and origin p2 of the hip-pitch link (exact quarter turns)
  const M3 F0{{0, 0, -1}, {0, 1, 0}, {1, 0, 0}}, F1{{0, 0, 1}, {0, 1, 0}, {-1, 0, 0}}, F2{{0, 0, -1}, {1, 0, 0}, {0, -1, 0}};
  const M3 A0 = F0 * rotz(ang[0]); const V3 p1 = V3{0.021, 0.135 * sg, 0} + A0 * V3{0, 0, -0.07};
  const M3 A1 = A0 * F1 * rotz(ang[1]); const V3 p2 = p1 + A1 * V3{0, 0, -0.09};
  const M3 A2 = A1 * F2 * rotz(ang[2]);
  // foot point in the hip-pitch frame and its partials w.r.t. shin / tarsus (rotations about z through s0 / t0)
  const V3 f0 = t0 + R3 * V3{0.408, -0.04, 0}, u = f0 + rotz(kn + sh + ta + ang[6]) * V3{0.01762, 0.05219, 0};
  const V3 ps = A2 * crs(ez, u - s0), pt = A2 * crs(ez, u - t0);
  (void)p2;
  const double j00 = ps.x - pt.x * b / a, j10 = ps.z - pt.z * b / a, j01 = pt.x / a, j11 = pt.z / a, det = j00 * j11 - j10 * j01;
  const double t0_ = 1500.0 * sh, t1_ = 1250.0 * (H - 2.586e-6);
  const double fx = -0.5 * (t0_ * j11 - t1_ * j10) / det, fz = -0.5 * (-t0_ * j01 + t1_ * j00) / det;
  // ---- heading-free world frame
  const double w = quat[0], x = quat[1], y = quat[2], z = quat[3];
  const double wx = (1 - 2 * (y * y + z * z)) * fx + 2 * (x * z + w * y) * fz, wy = 2 * (x * y + w * z) * fx + 2 * (y * z - w * x) * fz, wz = 2 * (x * z - w * y) * fx + (1 - 2 * (x * x + y * y)) * fz;
  const double yaw = std::atan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z)), cy = std::cos(yaw), sy = std::sin(yaw);
  force[0] = cy * wx + sy * wy; force[1] = -sy * wx + cy * wy; force[2] = wz;
}

}  // namespace cassie
