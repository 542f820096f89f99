// step_kernel.cu — dispatch of the stepping kernels (step_kernel.cuh, instantiated per motor count in
// step_kernel_nm*.cu) plus the small kernels around them: position publishing and the cross-GPU
// hand-shake of the fused exchange.
#include <cstdlib>

#include "internal.h"

template <int NM_T>
void launch_step_nm(const DevState& s, const DevParams* uniform_params, double dt, int k_substeps, int uniform_mode, bool any_moment, cudaStream_t stream, int* info);

int launch_step(const DevState& s, const DevParams* uniform_params, double dt, int k_substeps, int uniform_mode, int uniform_nm, bool any_moment,
                cudaStream_t stream, int* info) {
  int scratch[4];
  if (!info) info = scratch;
  info[0] = info[1] = info[2] = info[3] = 0;
  if (s.n <= 0) return 0;
  switch (uniform_nm) {
    case 4:
      launch_step_nm<4>(s, uniform_params, dt, k_substeps, uniform_mode, any_moment, stream, info);
      break;
    case 6:
      launch_step_nm<6>(s, uniform_params, dt, k_substeps, uniform_mode, any_moment, stream, info);
      break;
    case 8:
      launch_step_nm<8>(s, uniform_params, dt, k_substeps, uniform_mode, any_moment, stream, info);
      break;
    default:
      launch_step_nm<0>(s, uniform_params, dt, k_substeps, uniform_mode, any_moment, stream, info);
      break;
  }
  return 1;
}

namespace {
__global__ void publish_positions_kernel(DevState s) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= s.n) return;
  double* gp = s.gpos + 3 * (s.shard_begin + i);
  gp[0]      = s.st[tix(ST_ROWS, 0, i)];
  gp[1]      = s.st[tix(ST_ROWS, 1, i)];
  gp[2]      = s.st[tix(ST_ROWS, 2, i)];
}

// Flag block of a rank: [0, G) epochs written by the peers; [G, 3G) the peers' displacement words
// (float bits of the largest squared displacement of their stepping launch), double-buffered by epoch
// parity — a peer is at most one tick ahead, so the slot of epoch e is not overwritten before e + 2.
__global__ void p2p_signal_kernel(unsigned long long* const* peer_flags, int n_ranks, int rank, unsigned long long epoch, const uint32_t* disp,
                                  uint32_t disp_if_untracked) {
  const int r = threadIdx.x;
  if (r < n_ranks && r != rank)
    *reinterpret_cast<volatile unsigned long long*>(peer_flags[r] + n_ranks + 2 * rank + int(epoch & 1ull)) =
        (unsigned long long)(disp ? *disp : disp_if_untracked);
  __threadfence_system();
  if (r < n_ranks && r != rank) *reinterpret_cast<volatile unsigned long long*>(peer_flags[r] + rank) = epoch;
}
// one lane per peer spins (bounded) on this rank's own flag slots, which the peers write over NVLink
__global__ void p2p_wait_kernel(const unsigned long long* flags, int n_ranks, int rank, unsigned long long epoch, int* status, long long budget,
                                uint32_t* disp) {
  const int r = threadIdx.x;
  if (r < n_ranks && r != rank) {
    const long long t0 = clock64();
    bool            ok = true;
    while (*reinterpret_cast<const volatile unsigned long long*>(flags + r) < epoch) {
      if (clock64() - t0 > budget) {  // a peer is gone; report instead of hanging the GPU
        *status = 1;
        ok      = false;
        break;
      }
      __nanosleep(64);
    }
    if (disp) {
      // the swarm-wide displacement bound: the largest of every rank's (a lost peer counts as "unbounded")
      __threadfence_system();
      const uint32_t theirs = ok ? uint32_t(*reinterpret_cast<const volatile unsigned long long*>(flags + n_ranks + 2 * r + int(epoch & 1ull))) : 0xFFFFFFFFu;
      atomicMax(disp, theirs);
    }
  }
}
}  // namespace

int launch_p2p_signal(unsigned long long* const* peer_flags, int n_ranks, int rank, unsigned long long epoch, const uint32_t* disp,
                      uint32_t disp_if_untracked, cudaStream_t stream) {
  p2p_signal_kernel<<<1, 32, 0, stream>>>(peer_flags, n_ranks, rank, epoch, disp, disp_if_untracked);
  return 1;
}
int launch_p2p_wait(const unsigned long long* flags, int n_ranks, int rank, unsigned long long epoch, int* status, uint32_t* disp, cudaStream_t stream) {
  static long long budget = 0;
  if (!budget) {
    const char* e = getenv("MRSB_P2P_TIMEOUT_MS");
    budget        = (e ? atoll(e) : 20000LL) * 2000000LL;  // default 20 s at ~2 GHz SM clock
  }
  p2p_wait_kernel<<<1, 32, 0, stream>>>(flags, n_ranks, rank, epoch, status, budget, disp);
  return 1;
}

int launch_publish_positions(const DevState& s, cudaStream_t stream) {
  if (s.n <= 0) return 0;
  publish_positions_kernel<<<unsigned((s.n + 255) / 256), 256, 0, stream>>>(s);
  return 1;
}
