// io_kernels.cu — row-per-UAV (caller side) <-> structure-of-arrays (device side) movers behind the
// setInput / getState / applyForce ... entry points of the C ABI.  One thread per addressed UAV;
// idx == nullptr means "UAV k" (identity).
#include "internal.h"

namespace {

#define DEV __device__ __forceinline__
DEV int64_t at(const int32_t* idx, int64_t k) { return idx ? int64_t(idx[k]) : k; }

inline unsigned nblk(int64_t n, int t = 256) { return unsigned((n + t - 1) / t); }

// UavSystem::setInput overloads (US:175-248): store the command, switch the mode.  cos/sin of the
// commanded heading (used by CTL/acceleration_controller.hpp:50) are cached once per command
// instead of being re-evaluated every step.
__global__ void scatter_input_kernel(DevState s, int mode, int64_t n, const int32_t* __restrict__ idx, const double* __restrict__ payload, int stride) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int64_t i = at(idx, k);
  if (i < 0 || i >= s.n) return;
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
}

__global__ void set_mode_kernel(DevState s, int64_t n, const int32_t* __restrict__ idx, int mode) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int64_t i = at(idx, k);
  if (i < 0 || i >= s.n) return;
  s.mode[i] = uint8_t(mode);
}

// payload[k][0 .. rows) -> rows [row0, row0+rows) of UAV i
__global__ void scatter_rows_kernel(double* __restrict__ dst, int rows_total, int row0, int rows, int64_t n, const int32_t* __restrict__ idx,
                                    const double* __restrict__ payload, int stride) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int64_t i = at(idx, k);
  for (int r = 0; r < rows; r++) dst[tix(rows_total, row0 + r, i)] = payload[k * stride + r];
}

__global__ void gather_rows_kernel(const double* __restrict__ src, int rows_total, int row0, int rows, int64_t n, const int32_t* __restrict__ idx,
                                   double* __restrict__ out, int stride) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int64_t i = at(idx, k);
  for (int r = 0; r < rows; r++) out[k * stride + r] = src[tix(rows_total, row0 + r, i)];
}

__global__ void flag_update_kernel(DevState s, int64_t n, const int32_t* __restrict__ idx, uint32_t and_mask, uint32_t or_mask) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int64_t i = at(idx, k);
  s.flags[i]      = (s.flags[i] & and_mask) | or_mask;
}

__global__ void gather_u32_kernel(const uint32_t* __restrict__ src, int64_t n, const int32_t* __restrict__ idx, uint32_t* __restrict__ out) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n) return;
  out[k] = src[at(idx, k)];
}

__global__ void gather_u8_kernel(const uint8_t* __restrict__ src, int64_t n, const int32_t* __restrict__ idx, int32_t* __restrict__ out) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n) return;
  out[k] = int32_t(src[at(idx, k)]);
}

// MultirotorModel::setStatePos (MM:439-446): x, _initial_pos_, R = AngleAxis(-heading, z)
__global__ void set_state_pos_kernel(DevState s, int64_t n, const int32_t* __restrict__ idx, const double* __restrict__ xyz,
                                     const double* __restrict__ hdg) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int64_t i  = at(idx, k);
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
  double* gp        = s.gpos + 3 * (s.shard_begin + i);
  gp[0]             = px;
  gp[1]             = py;
  gp[2]             = pz;
}

// before set_state overwrites v: remember the old v as v_prev (MM:424-433 leaves v_prev alone)
__global__ void stash_vprev_kernel(DevState s, int64_t n, const int32_t* __restrict__ idx) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int64_t i = at(idx, k);
  if (s.flags[i] & FLAG_VPREV) return;
  for (int r = 0; r < 3; r++) s.vprev[tix(VPREV_ROWS, r, i)] = s.st[tix(ST_ROWS, 3 + r, i)];
  s.flags[i] |= FLAG_VPREV;
}

__global__ void gather_vprev_kernel(DevState s, int64_t n, const int32_t* __restrict__ idx, double* __restrict__ out) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int64_t i  = at(idx, k);
  const bool    ov = s.flags[i] & FLAG_VPREV;
  for (int r = 0; r < 3; r++) out[3 * k + r] = ov ? s.vprev[tix(VPREV_ROWS, r, i)] : s.st[tix(ST_ROWS, 3 + r, i)];
}

__global__ void reset_pid_kernel(DevState s, int64_t n, const int32_t* __restrict__ idx, int row0, int rows) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int64_t i = at(idx, k);
  for (int r = row0; r < row0 + rows; r++) s.pid[tix(PID_ROWS, r, i)] = 0.0;
}

__global__ void set_pset_kernel(int32_t* __restrict__ pset, int64_t n, const int32_t* __restrict__ idx, int64_t offset,
                                const int32_t* __restrict__ values) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n) return;
  pset[offset + at(idx, k)] = values[k];
}

}  // namespace

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
                        cudaStream_t st) {
  if (n <= 0) return 0;
  scatter_rows_kernel<<<nblk(n), 256, 0, st>>>(dst, rows_total, row0, rows, n, idx, payload, stride);
  return 1;
}
int launch_gather_rows(const double* src, int rows_total, int row0, int rows, int64_t n, const int32_t* idx, double* out, int stride, cudaStream_t st) {
  if (n <= 0) return 0;
  gather_rows_kernel<<<nblk(n), 256, 0, st>>>(src, rows_total, row0, rows, n, idx, out, stride);
  return 1;
}
int launch_flag_update(const DevState& s, int64_t n, const int32_t* idx, uint32_t and_mask, uint32_t or_mask, cudaStream_t st) {
  if (n <= 0) return 0;
  flag_update_kernel<<<nblk(n), 256, 0, st>>>(s, n, idx, and_mask, or_mask);
  return 1;
}
int launch_gather_u32(const uint32_t* src, int64_t n, const int32_t* idx, uint32_t* out, cudaStream_t st) {
  if (n <= 0) return 0;
  gather_u32_kernel<<<nblk(n), 256, 0, st>>>(src, n, idx, out);
  return 1;
}
int launch_gather_u8(const uint8_t* src, int64_t n, const int32_t* idx, int32_t* out, cudaStream_t st) {
  if (n <= 0) return 0;
  gather_u8_kernel<<<nblk(n), 256, 0, st>>>(src, n, idx, out);
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
int launch_reset_pid(const DevState& s, int64_t n, const int32_t* idx, int row0, int rows, cudaStream_t st) {
  if (n <= 0) return 0;
  reset_pid_kernel<<<nblk(n), 256, 0, st>>>(s, n, idx, row0, rows);
  return 1;
}
int launch_set_pset(int32_t* pset, int64_t n, const int32_t* idx, int64_t offset, const int32_t* values, cudaStream_t st) {
  if (n <= 0) return 0;
  set_pset_kernel<<<nblk(n), 256, 0, st>>>(pset, n, idx, offset, values);
  return 1;
}
