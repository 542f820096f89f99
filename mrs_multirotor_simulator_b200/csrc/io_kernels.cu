// io_kernels.cu — row-per-UAV (caller side) <-> structure-of-arrays (device side) movers behind the
// setInput / getState / applyForce ... entry points of the C ABI.  One thread per addressed UAV;
// idx == nullptr means "UAV k" (identity).
#include "internal.h"

namespace {

#define DEV __device__ __forceinline__
// k-th addressed UAV: its (external) local index ...
DEV int64_t ext(const int32_t* idx, int64_t k) { return idx ? int64_t(idx[k]) : k; }
// ... and the slot of the tiled arrays it lives in: the identity unless the batch was bucketed by airframe at create
// (DevState::perm, api.cu) so that every 128-UAV tile holds one airframe
DEV int64_t at(const int32_t* perm, const int32_t* idx, int64_t k) {
  const int64_t e = ext(idx, k);
  return perm ? int64_t(perm[e]) : e;
}

inline unsigned nblk(int64_t n, int t = 256) { return unsigned((n + t - 1) / t); }

// UavSystem::setInput overloads (US:175-248): store the command, switch the mode.  cos/sin of the
// commanded heading (used by CTL/acceleration_controller.hpp:50) are cached once per command
// instead of being re-evaluated every step.
__global__ void scatter_input_kernel(DevState s, int mode, int64_t n, const int32_t* __restrict__ idx, const double* __restrict__ payload, int stride) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int64_t e = ext(idx, k);
  if (e < 0 || e >= s.n) return;
  const int64_t i = s.perm ? int64_t(s.perm[e]) : e;
  const double* p    = payload + k * stride;
  const int     rows = mode == MRSB_ACTUATOR_CMD ? MRSB_NM : (mode == MRSB_ATTITUDE_CMD ? 10 : (mode == MRSB_TILT_HDG_RATE_CMD ? 5 : 4));
  for (int r = 0; r < rows; r++) s.cmd[tix(CMD_ROWS, r, i)] = r < stride ? p[r] : 0.0;
  if (mode == MRSB_POSITION_CMD || mode == MRSB_VELOCITY_HDG_CMD || mode == MRSB_ACCELERATION_HDG_CMD) {
    double sn, cs;
    sincos(p[3], &sn, &cs);
    s.cmd[tix(CMD_ROWS, CMD_COS, i)] = cs;
    s.cmd[tix(CMD_ROWS, CMD_SIN, i)] = sn;
  }
  s.mode[i] = uint8_t(mode);
  if (!(s.flags[i] & FLAG_HAD_INPUT)) s.flags[i] |= FLAG_HAD_INPUT;  // time_last_input_ > 0 from now on (ROSW:265)
}

__global__ void set_mode_kernel(DevState s, int64_t n, const int32_t* __restrict__ idx, int mode) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int64_t e = ext(idx, k);
  if (e < 0 || e >= s.n) return;
  const int64_t i = s.perm ? int64_t(s.perm[e]) : e;
  s.mode[i] = uint8_t(mode);
}

// payload[k][0 .. rows) -> rows [row0, row0+rows) of UAV i
__global__ void scatter_rows_kernel(double* __restrict__ dst, int rows_total, int row0, int rows, int64_t n, const int32_t* __restrict__ idx,
                                    const double* __restrict__ payload, int stride, const int32_t* __restrict__ perm) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int64_t i = at(perm, idx, k);
  for (int r = 0; r < rows; r++) dst[tix(rows_total, row0 + r, i)] = payload[k * stride + r];
}

__global__ void gather_rows_kernel(const double* __restrict__ src, int rows_total, int row0, int rows, int64_t n, const int32_t* __restrict__ idx,
                                   double* __restrict__ out, int stride, const int32_t* __restrict__ perm) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int64_t i = at(perm, idx, k);
  for (int r = 0; r < rows; r++) out[k * stride + r] = src[tix(rows_total, row0 + r, i)];
}

__global__ void flag_update_kernel(DevState s, int64_t n, const int32_t* __restrict__ idx, uint32_t and_mask, uint32_t or_mask) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int64_t i = at(s.perm, idx, k);
  s.flags[i]      = (s.flags[i] & and_mask) | or_mask;
}

__global__ void gather_u32_kernel(const uint32_t* __restrict__ src, int64_t n, const int32_t* __restrict__ idx, uint32_t* __restrict__ out,
                                  const int32_t* __restrict__ perm) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n) return;
  out[k] = src[at(perm, idx, k)];
}

__global__ void gather_u8_kernel(const uint8_t* __restrict__ src, int64_t n, const int32_t* __restrict__ idx, int32_t* __restrict__ out,
                                 const int32_t* __restrict__ perm) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n) return;
  out[k] = int32_t(src[at(perm, idx, k)]);
}

// MultirotorModel::setStatePos (MM:439-446): x, _initial_pos_, R = AngleAxis(-heading, z)
__global__ void set_state_pos_kernel(DevState s, int64_t n, const int32_t* __restrict__ idx, const double* __restrict__ xyz,
                                     const double* __restrict__ hdg) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int64_t e  = ext(idx, k);
  const int64_t i  = s.perm ? int64_t(s.perm[e]) : e;
  const double  px = xyz ? xyz[3 * k] : 0.0, py = xyz ? xyz[3 * k + 1] : 0.0, pz = xyz ? xyz[3 * k + 2] : 0.0;
  double        sn, cs;
  sincos(hdg ? -hdg[k] : -0.0, &sn, &cs);
  s.st[tix(ST_ROWS, 0, i)] = px;
  s.st[tix(ST_ROWS, 1, i)] = py;
  s.st[tix(ST_ROWS, 2, i)] = pz;
  s.initz[i]       = pz;
  // column-major R = [[c,-s,0],[s,c,0],[0,0,(1-c)+c]] for angle -heading (Eigen AngleAxis::toRotationMatrix)
  s.st[tix(ST_ROWS, 6, i)]  = cs;
  s.st[tix(ST_ROWS, 7, i)]  = sn;
  s.st[tix(ST_ROWS, 8, i)]  = 0.0;
  s.st[tix(ST_ROWS, 9, i)]  = -sn;
  s.st[tix(ST_ROWS, 10, i)] = cs;
  s.st[tix(ST_ROWS, 11, i)] = 0.0;
  s.st[tix(ST_ROWS, 12, i)] = 0.0;
  s.st[tix(ST_ROWS, 13, i)] = 0.0;
  s.st[tix(ST_ROWS, 14, i)] = (1.0 - cs) + cs;
  double* gp        = s.gpos + 3 * (s.shard_begin + e);
  gp[0]             = px;
  gp[1]             = py;
  gp[2]             = pz;
}

// before set_state overwrites v: remember the old v as v_prev (MM:424-433 leaves v_prev alone)
__global__ void stash_vprev_kernel(DevState s, int64_t n, const int32_t* __restrict__ idx) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int64_t i = at(s.perm, idx, k);
  if (s.flags[i] & FLAG_VPREV) return;
  for (int r = 0; r < 3; r++) s.vprev[tix(VPREV_ROWS, r, i)] = s.st[tix(ST_ROWS, 3 + r, i)];
  s.flags[i] |= FLAG_VPREV;
}

__global__ void gather_vprev_kernel(DevState s, int64_t n, const int32_t* __restrict__ idx, double* __restrict__ out) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int64_t i  = at(s.perm, idx, k);
  const bool    ov = s.flags[i] & FLAG_VPREV;
  for (int r = 0; r < 3; r++) out[3 * k + r] = ov ? s.vprev[tix(VPREV_ROWS, r, i)] : s.st[tix(ST_ROWS, 3 + r, i)];
}

__global__ void reset_pid_kernel(DevState s, int64_t n, const int32_t* __restrict__ idx, int pid0, int n_pids) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int64_t i = at(s.perm, idx, k);
  for (int r = pid0; r < pid0 + n_pids; r++) {
    s.pid[tix(PID_ROWS, r, i)]      = 0.0;  // last error
    s.pid[tix(PID_ROWS, 12 + r, i)] = 0.0;  // integral
  }
}

__global__ void set_pset_kernel(int32_t* __restrict__ pset, int64_t n, const int32_t* __restrict__ idx, int64_t offset,
                                const int32_t* __restrict__ values) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n) return;
  pset[offset + ext(idx, k)] = values[k];
}

// UavSystemRos::callbackTrackerCmd (ROSW:987-1022): one tracker command row -> the four sticky feed-forwards.
// row[11] = velocity xyz | acceleration xyz | heading_rate | use_velocity_horizontal | use_velocity_vertical | use_heading_rate | use_acceleration
__global__ void tracker_cmd_kernel(DevState s, int64_t n, const int32_t* __restrict__ idx, const double* __restrict__ rows) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int64_t i  = at(s.perm, idx, k);
  const double* r  = rows + MRSB_TRACKER_CMD_STRIDE * k;
  const bool    uh = r[7] != 0.0, uv = r[8] != 0.0, ur = r[9] != 0.0, ua = r[10] != 0.0;
  const double  v[3] = {uh ? r[0] : 0.0, uh ? r[1] : 0.0, uv ? r[2] : 0.0};
  const double  a[3] = {ua ? r[3] : 0.0, ua ? r[4] : 0.0, ua ? r[5] : 0.0};
  for (int c = 0; c < 3; c++) {
    s.ff[tix(FF_ROWS, FF_VEL_HDG + c, i)]      = v[c];
    s.ff[tix(FF_ROWS, FF_VEL_HDG_RATE + c, i)] = v[c];
    s.ff[tix(FF_ROWS, FF_ACC_HDG + c, i)]      = a[c];
    s.ff[tix(FF_ROWS, FF_ACC_HDG_RATE + c, i)] = a[c];
  }
  s.ff[tix(FF_ROWS, FF_ACC_HDG_RATE + 3, i)] = ur ? r[6] : 0.0;
  s.flags[i] |= FLAG_FF_VEL_HDG | FLAG_FF_VEL_HDG_RATE | FLAG_FF_ACC_HDG | FLAG_FF_ACC_HDG_RATE;
}

// packed positions of a subset: out[k] = xyz[idx[k]] (both packed [.][3])
__global__ void gather_xyz_kernel(const double* __restrict__ xyz, int64_t n, const int32_t* __restrict__ idx, double* __restrict__ out) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const double* p = xyz + 3 * int64_t(idx[k]);
  out[3 * k] = p[0], out[3 * k + 1] = p[1], out[3 * k + 2] = p[2];
}

// collision geometry of the addressed UAVs from their parameter set
__global__ void set_geom_kernel(double* __restrict__ geom, int64_t n, const int32_t* __restrict__ idx, int64_t offset, const int32_t* __restrict__ pset,
                                const DevParams* __restrict__ params) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int64_t    j = offset + ext(idx, k);
  const DevParams& P = params[pset[j]];
  geom[4 * j + 0]    = P.arm_length;
  geom[4 * j + 1]    = P.prop_radius;
  geom[4 * j + 2]    = P.mass;
  geom[4 * j + 3]    = 0.0;
}

// UavSystemRos::timeoutInput (ROSW:474-647): the active command becomes its "hover" version
__global__ void timeout_input_kernel(DevState s, int64_t n, const int32_t* __restrict__ idx) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int64_t i    = at(s.perm, idx, k);
  const int     mode = s.mode[i];
  s.flags[i] &= ~FLAG_HAD_INPUT;  // time_last_input_ = 0 (ROSW:256-259): with iterate_without_input off the UAV stops until the next command
  auto          C    = [&](int row) -> double& { return s.cmd[tix(CMD_ROWS, row, i)]; };
  auto          S    = [&](int row) { return s.st[tix(ST_ROWS, row, i)]; };
  const double  hdg  = atan2(S(7), S(6));  // AttitudeConverter(R).getHeading(): atan2(R10, R00)
  double        sh, ch;
  sincos(hdg, &sh, &ch);
  switch (mode) {
    case MRSB_POSITION_CMD:
      C(0) = S(0), C(1) = S(1), C(2) = S(2), C(3) = hdg, C(CMD_COS) = ch, C(CMD_SIN) = sh;
      break;
    case MRSB_VELOCITY_HDG_CMD:
    case MRSB_ACCELERATION_HDG_CMD:
      C(0) = 0.0, C(1) = 0.0, C(2) = 0.0, C(3) = hdg, C(CMD_COS) = ch, C(CMD_SIN) = sh;
      break;
    case MRSB_VELOCITY_HDG_RATE_CMD:
    case MRSB_ACCELERATION_HDG_RATE_CMD:
    case MRSB_ATTITUDE_RATE_CMD:
    case MRSB_CONTROL_GROUP_CMD:
      C(0) = 0.0, C(1) = 0.0, C(2) = 0.0, C(3) = 0.0;
      break;
    case MRSB_ATTITUDE_CMD:  // AttitudeConverter(0, 0, heading) -> Rz(heading), column-major; throttle 0
      C(0) = ch, C(1) = sh, C(2) = 0.0, C(3) = -sh, C(4) = ch, C(5) = 0.0, C(6) = 0.0, C(7) = 0.0, C(8) = 1.0, C(9) = 0.0;
      break;
    case MRSB_TILT_HDG_RATE_CMD:
      C(0) = 0.0, C(1) = 0.0, C(2) = 1.0, C(3) = 0.0, C(4) = 0.0;
      break;
    case MRSB_ACTUATOR_CMD:
      for (int m = 0; m < MRSB_NM; m++) C(m) = 0.0;
      break;
    default:
      break;
  }
}

// Eigen::Quaterniond(Matrix3d) as used by mrs_lib::AttitudeConverter(R); q = x y z w
DEV void quaternion_of(const double* m /* column-major */, double* q) {
  auto   M = [&](int r, int c) { return m[3 * c + r]; };
  double t = M(0, 0) + (M(1, 1) + M(2, 2));
  if (t > 0.0) {
    t    = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t    = 0.5 / t;
    q[0] = (M(2, 1) - M(1, 2)) * t;
    q[1] = (M(0, 2) - M(2, 0)) * t;
    q[2] = (M(1, 0) - M(0, 1)) * t;
  } else {
    int i = 0;
    if (M(1, 1) > M(0, 0)) i = 1;
    if (M(2, 2) > M(i, i)) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t    = sqrt(M(i, i) - M(j, j) - M(k, k) + 1.0);
    q[i] = 0.5 * t;
    t    = 0.5 / t;
    q[3] = (M(k, j) - M(j, k)) * t;
    q[j] = (M(j, i) + M(i, j)) * t;
    q[k] = (M(k, i) + M(i, k)) * t;
  }
}

// what = 0: publishOdometry (ROSW:340-368) rows [13]; 1: publishIMU (ROSW:374-395) rows [10];
// 2: publishRangefinder (ROSW:401-420) rows [1]; 3: all of them packed for device-resident callers, rows [17] =
// odometry 13 | IMU linear acceleration 3 | range 1
__global__ void observe_kernel(DevState s, int what, int64_t n, const int32_t* __restrict__ idx, double* __restrict__ out, int stride) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int64_t e = ext(idx, k);
  const int64_t i = s.perm ? int64_t(s.perm[e]) : e;
  double        x[3], v[3], R[9], w[3], q[4];
  for (int r = 0; r < 3; r++) {
    x[r] = s.st[tix(ST_ROWS, r, i)];
    v[r] = s.st[tix(ST_ROWS, 3 + r, i)];
    w[r] = s.st[tix(ST_ROWS, 15 + r, i)];
  }
  for (int r = 0; r < 9; r++) R[r] = s.st[tix(ST_ROWS, 6 + r, i)];
  double* o = out + k * stride;
  if (what == 0 || what == 1 || what == 3) quaternion_of(R, q);
  double range = 0.0;
  if (what == 2 || what == 3) {
    const double bz   = R[8];
    const double tilt = acos((-R[6]) * 0.0 + ((-R[7]) * 0.0 + (-bz) * -1.0));
    range             = bz > 0.0 ? (x[2] - s.params[s.pset[s.shard_begin + e]].ground_z) / cos(tilt) + 0.01 : 1.7976931348623157e308;
    if (range > 40.0) range = 41.0;
  }
  if (what == 0 || what == 3) {
    for (int r = 0; r < 3; r++) o[r] = x[r];
    for (int r = 0; r < 4; r++) o[3 + r] = q[r];
    for (int r = 0; r < 3; r++) {
      o[7 + r]  = R[3 * r] * v[0] + (R[3 * r + 1] * v[1] + R[3 * r + 2] * v[2]);  // R^T v
      o[10 + r] = w[r];
    }
    if (what == 3) {
      for (int r = 0; r < 3; r++) o[13 + r] = s.imu[tix(F3_ROWS, r, i)];
      o[16] = range;
    }
  } else if (what == 1) {
    for (int r = 0; r < 3; r++) {
      o[r]     = w[r];
      o[3 + r] = s.imu[tix(F3_ROWS, r, i)];
    }
    for (int r = 0; r < 4; r++) o[6 + r] = q[r];
  } else {
    o[0] = range;
  }
}

}  // namespace

int launch_timeout_input(const DevState& s, int64_t n, const int32_t* idx, cudaStream_t st) {
  if (n <= 0) return 0;
  timeout_input_kernel<<<nblk(n), 256, 0, st>>>(s, n, idx);
  return 1;
}
int launch_observe(const DevState& s, int what, int64_t n, const int32_t* idx, double* out, int stride, cudaStream_t st) {
  if (n <= 0) return 0;
  observe_kernel<<<nblk(n), 256, 0, st>>>(s, what, n, idx, out, stride);
  return 1;
}

int launch_scatter_input(const DevState& s, int mode, int64_t n, const int32_t* idx, const double* payload, int stride, cudaStream_t st) {
  if (n <= 0) return 0;
  scatter_input_kernel<<<nblk(n), 256, 0, st>>>(s, mode, n, idx, payload, stride);
  return 1;
}
int launch_set_mode(const DevState& s, int64_t n, const int32_t* idx, int mode, cudaStream_t st) {
  if (n <= 0) return 0;
  set_mode_kernel<<<nblk(n), 256, 0, st>>>(s, n, idx, mode);
  return 1;
}
int launch_scatter_rows(double* dst, int rows_total, int row0, int rows, int64_t n, const int32_t* idx, const double* payload, int stride,
                        const int32_t* perm, cudaStream_t st) {
  if (n <= 0) return 0;
  scatter_rows_kernel<<<nblk(n), 256, 0, st>>>(dst, rows_total, row0, rows, n, idx, payload, stride, perm);
  return 1;
}
int launch_gather_rows(const double* src, int rows_total, int row0, int rows, int64_t n, const int32_t* idx, double* out, int stride, const int32_t* perm,
                       cudaStream_t st) {
  if (n <= 0) return 0;
  gather_rows_kernel<<<nblk(n), 256, 0, st>>>(src, rows_total, row0, rows, n, idx, out, stride, perm);
  return 1;
}
int launch_flag_update(const DevState& s, int64_t n, const int32_t* idx, uint32_t and_mask, uint32_t or_mask, cudaStream_t st) {
  if (n <= 0) return 0;
  flag_update_kernel<<<nblk(n), 256, 0, st>>>(s, n, idx, and_mask, or_mask);
  return 1;
}
int launch_gather_u32(const uint32_t* src, int64_t n, const int32_t* idx, uint32_t* out, const int32_t* perm, cudaStream_t st) {
  if (n <= 0) return 0;
  gather_u32_kernel<<<nblk(n), 256, 0, st>>>(src, n, idx, out, perm);
  return 1;
}
int launch_gather_u8(const uint8_t* src, int64_t n, const int32_t* idx, int32_t* out, const int32_t* perm, cudaStream_t st) {
  if (n <= 0) return 0;
  gather_u8_kernel<<<nblk(n), 256, 0, st>>>(src, n, idx, out, perm);
  return 1;
}
int launch_set_state_pos(const DevState& s, int64_t n, const int32_t* idx, const double* xyz, const double* hdg, cudaStream_t st) {
  if (n <= 0) return 0;
  set_state_pos_kernel<<<nblk(n), 256, 0, st>>>(s, n, idx, xyz, hdg);
  return 1;
}
int launch_stash_vprev(const DevState& s, int64_t n, const int32_t* idx, cudaStream_t st) {
  if (n <= 0) return 0;
  stash_vprev_kernel<<<nblk(n), 256, 0, st>>>(s, n, idx);
  return 1;
}
int launch_gather_vprev(const DevState& s, int64_t n, const int32_t* idx, double* out, cudaStream_t st) {
  if (n <= 0) return 0;
  gather_vprev_kernel<<<nblk(n), 256, 0, st>>>(s, n, idx, out);
  return 1;
}
int launch_reset_pid(const DevState& s, int64_t n, const int32_t* idx, int pid0, int n_pids, cudaStream_t st) {
  if (n <= 0) return 0;
  reset_pid_kernel<<<nblk(n), 256, 0, st>>>(s, n, idx, pid0, n_pids);
  return 1;
}
int launch_gather_xyz(const double* xyz, int64_t n, const int32_t* idx, double* out, cudaStream_t st) {
  if (n <= 0) return 0;
  gather_xyz_kernel<<<nblk(n), 256, 0, st>>>(xyz, n, idx, out);
  return 1;
}
int launch_tracker_cmd(const DevState& s, int64_t n, const int32_t* idx, const double* rows, cudaStream_t st) {
  if (n <= 0) return 0;
  tracker_cmd_kernel<<<nblk(n), 256, 0, st>>>(s, n, idx, rows);
  return 1;
}
int launch_set_geom(double* geom, int64_t n, const int32_t* idx, int64_t offset, const int32_t* pset, const DevParams* params, cudaStream_t st) {
  if (n <= 0) return 0;
  set_geom_kernel<<<nblk(n), 256, 0, st>>>(geom, n, idx, offset, pset, params);
  return 1;
}
int launch_set_pset(int32_t* pset, int64_t n, const int32_t* idx, int64_t offset, const int32_t* values, cudaStream_t st) {
  if (n <= 0) return 0;
  set_pset_kernel<<<nblk(n), 256, 0, st>>>(pset, n, idx, offset, values);
  return 1;
}
