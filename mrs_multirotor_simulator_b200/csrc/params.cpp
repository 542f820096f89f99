// params.cpp — host-side derivation of the device parameter set from the public structs.
// Mirrors what the reference does once at construction time, on the host:
//   ModelParams defaults                       MM:26-66
//   inertia from mass/arm/body height          ROSW:664-671
//   allocation-matrix scaling                  ROSW:98-103 (MM:59-62)
//   Mixer::calculateAllocation                 CTL/mixer.hpp:72-101
//   RateController gains scaled by J_ii        CTL/rate_controller.hpp:56-65
//   J.inverse() (MM:350, every derivative call in the reference; once here)
#include <cmath>
#include <cstring>

#include "internal.h"
#include "params.h"

extern "C" void mrsb_model_params_finalize(mrsb_model_params* p) {
  const double m = p->mass, a = p->arm_length, bh = p->body_height;
  std::memset(p->J, 0, sizeof(p->J));
  p->J[0] = m * (3.0 * a * a + bh * bh) / 12.0;
  p->J[4] = m * (3.0 * a * a + bh * bh) / 12.0;
  p->J[8] = (m * a * a) / 2.0;
  const double s01 = p->arm_length * p->kf;
  const double s2  = p->km * (3.0 * p->prop_radius) * p->kf;
  const double s3  = p->kf;
  for (int c = 0; c < MRSB_MAX_MOTORS; c++) {
    p->allocation_matrix[0 * MRSB_MAX_MOTORS + c] *= s01;
    p->allocation_matrix[1 * MRSB_MAX_MOTORS + c] *= s01;
    p->allocation_matrix[2 * MRSB_MAX_MOTORS + c] *= s2;
    p->allocation_matrix[3 * MRSB_MAX_MOTORS + c] *= s3;
  }
}

extern "C" void mrsb_model_params_default(mrsb_model_params* p) {
  std::memset(p, 0, sizeof(*p));
  p->n_motors              = 4;
  p->g                     = 9.81;
  p->mass                  = 2.0;
  p->kf                    = 0.00000027087;
  p->km                    = 0.07;
  p->prop_radius           = 0.15;
  p->arm_length            = 0.25;
  p->body_height           = 0.1;
  p->motor_time_constant   = 0.03;
  p->max_rpm               = 7800;
  p->min_rpm               = 1170;
  p->air_resistance_coeff  = 0.30;
  p->ground_enabled        = 0;
  p->ground_z              = 0.0;
  p->takeoff_patch_enabled = 1;
  const double quad[4][4]  = {{-0.707, 0.707, 0.707, -0.707}, {-0.707, 0.707, -0.707, 0.707}, {-1, -1, 1, 1}, {1, 1, 1, 1}};
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) p->allocation_matrix[r * MRSB_MAX_MOTORS + c] = quad[r][c];
  mrsb_model_params_finalize(p);
}

extern "C" void mrsb_controller_params_default(mrsb_controller_params* c) {
  std::memset(c, 0, sizeof(*c));
  c->mixer_desaturation = 1;
  c->rate_kp = 4.0, c->rate_kd = 0.04, c->rate_ki = 0.0;
  c->att_kp = 6.0, c->att_kd = 0.05, c->att_ki = 0.01, c->att_max_rate_roll_pitch = 10.0, c->att_max_rate_yaw = 1.0;
  c->vel_kp = 2.0, c->vel_kd = 0.05, c->vel_ki = 0.01, c->vel_max_acceleration = 4.0;
  c->pos_kp = 2.0, c->pos_kd = 0.15, c->pos_ki = 0.2, c->pos_max_velocity = 6.0;
}

// Gauss-Jordan inverse with partial pivoting, n <= 4
static bool invert_small(const double* a, int n, double* out) {
  double w[4][8];
  for (int r = 0; r < n; r++)
    for (int c = 0; c < n; c++) {
      w[r][c]     = a[r * n + c];
      w[r][n + c] = (r == c) ? 1.0 : 0.0;
    }
  for (int k = 0; k < n; k++) {
    int piv = k;
    for (int r = k + 1; r < n; r++)
      if (std::fabs(w[r][k]) > std::fabs(w[piv][k])) piv = r;
    if (w[piv][k] == 0.0) return false;
    if (piv != k)
      for (int c = 0; c < 2 * n; c++) {
        const double t = w[k][c];
        w[k][c]        = w[piv][c];
        w[piv][c]      = t;
      }
    const double d = w[k][k];
    for (int c = 0; c < 2 * n; c++) w[k][c] /= d;
    for (int r = 0; r < n; r++) {
      if (r == k) continue;
      const double f = w[r][k];
      if (f == 0.0) continue;
      for (int c = 0; c < 2 * n; c++) w[r][c] -= f * w[k][c];
    }
  }
  for (int r = 0; r < n; r++)
    for (int c = 0; c < n; c++) out[r * n + c] = w[r][n + c];
  return true;
}

void mrsb_mixer_allocation(const mrsb_model_params& mp, double mix[MRSB_MAX_MOTORS][4]) {
  const int n = mp.n_motors;
  double    AAt[16];
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) {
      double s = 0.0;
      for (int m = 0; m < n; m++) s += mp.allocation_matrix[r * MRSB_MAX_MOTORS + m] * mp.allocation_matrix[c * MRSB_MAX_MOTORS + m];
      AAt[r * 4 + c] = s;
    }
  double inv[16];
  if (!invert_small(AAt, 4, inv)) {
    for (int k = 0; k < 16; k++) inv[k] = NAN;
  }
  for (int m = 0; m < MRSB_MAX_MOTORS; m++)
    for (int c = 0; c < 4; c++) mix[m][c] = 0.0;
  for (int m = 0; m < n; m++) {
    for (int c = 0; c < 4; c++) {
      double s = 0.0;
      for (int k = 0; k < 4; k++) s += mp.allocation_matrix[k * MRSB_MAX_MOTORS + m] * inv[k * 4 + c];
      mix[m][c] = s;
    }
    // PX4-style normalisation (CTL/mixer.hpp:82-100)
    const double z = mix[m][0] * mix[m][0] + mix[m][1] * mix[m][1];
    if (z > 0.0) {
      const double s = std::sqrt(z);
      mix[m][0] /= s;
      mix[m][1] /= s;
    }
    mix[m][2] = mix[m][2] > 1e-2 ? 1.0 : (mix[m][2] < -1e-2 ? -1.0 : 0.0);
    mix[m][3] = 1.0;
  }
}

extern "C" void mrsb_mixer_allocation_of(const mrsb_model_params* params, double* out) {
  double mix[MRSB_MAX_MOTORS][4];
  mrsb_mixer_allocation(*params, mix);
  std::memcpy(out, mix, sizeof(mix));
}

void mrsb_derive(const mrsb_model_params& mp, const mrsb_controller_params& cp, DevParams* d) {
  std::memset(d, 0, sizeof(*d));
  d->n_motors           = mp.n_motors;
  d->ground_enabled     = mp.ground_enabled;
  d->mixer_desaturation = cp.mixer_desaturation;
  d->g                  = mp.g;
  d->mass               = mp.mass;
  d->inv_mass           = 1.0 / mp.mass;
  d->inv_kf_n           = 1.0 / (mp.kf * mp.n_motors);
  d->min_rpm            = mp.min_rpm;
  d->rpm_range          = mp.max_rpm - mp.min_rpm;
  d->inv_rpm_range      = 1.0 / (mp.max_rpm - mp.min_rpm);
  d->neg_inv_tau        = -1.0 / mp.motor_time_constant;
  d->air_k              = mp.air_resistance_coeff * M_PI * mp.arm_length * mp.arm_length;
  d->ground_z           = mp.ground_z;
  d->takeoff_rpm        = 0.90 * std::sqrt((mp.mass * mp.g) / (mp.n_motors * mp.kf));
  d->arm_length         = mp.arm_length;
  d->prop_radius        = mp.prop_radius;
  std::memcpy(d->J, mp.J, sizeof(d->J));
  d->j_diagonal = (mp.J[1] == 0 && mp.J[2] == 0 && mp.J[3] == 0 && mp.J[5] == 0 && mp.J[6] == 0 && mp.J[7] == 0);
  if (d->j_diagonal) {
    d->Jinv[0] = 1.0 / mp.J[0];
    d->Jinv[4] = 1.0 / mp.J[4];
    d->Jinv[8] = 1.0 / mp.J[8];
  } else if (!invert_small(mp.J, 3, d->Jinv)) {
    for (int k = 0; k < 9; k++) d->Jinv[k] = NAN;
  }
  for (int r = 0; r < 4; r++)
    for (int m = 0; m < MRSB_MAX_MOTORS; m++) d->alloc[r][m] = m < mp.n_motors ? mp.allocation_matrix[r * MRSB_MAX_MOTORS + m] : 0.0;
  mrsb_mixer_allocation(mp, d->mix);
  d->pos_kp = cp.pos_kp, d->pos_kd = cp.pos_kd, d->pos_ki = cp.pos_ki, d->pos_sat = cp.pos_max_velocity;
  d->vel_kp = cp.vel_kp, d->vel_kd = cp.vel_kd, d->vel_ki = cp.vel_ki, d->vel_sat = cp.vel_max_acceleration;
  d->att_kp = cp.att_kp, d->att_kd = cp.att_kd, d->att_ki = cp.att_ki, d->att_sat_rp = cp.att_max_rate_roll_pitch,
  d->att_sat_yaw = cp.att_max_rate_yaw;
  for (int k = 0; k < 3; k++) {
    const double Jii = mp.J[4 * k];
    d->rate_kp[k]    = cp.rate_kp * Jii;
    d->rate_kd[k]    = cp.rate_kd * Jii;
    d->rate_ki[k]    = cp.rate_ki * Jii;
  }
}
