// api.cu — the C ABI of include/mrsb.h: host bookkeeping around the kernels.
// Every entry point cites the reference member it replaces in include/mrsb.h.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "internal.h"
#include "params.h"

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CU(call)                                                                                              \
  do {                                                                                                        \
    cudaError_t e_ = (call);                                                                                  \
    if (e_ != cudaSuccess) return fail(MRSB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// ------------------------------------------------------------------------------------------
// NCCL, loaded lazily so single-GPU users need no libnccl
// ------------------------------------------------------------------------------------------
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*)                                                         = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int)                                  = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t)                                                            = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t)    = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)()                                                                       = nullptr;
  ncclResult_t (*GroupEnd)()                                                                         = nullptr;
  const char* (*GetErrorString)(ncclResult_t)                                                        = nullptr;
};
static NcclApi g_nccl;

static int load_nccl() {
  if (g_nccl.lib) return MRSB_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void*       lib     = nullptr;
  for (const char* n : names) {
    lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) return fail(MRSB_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                          \
  *(void**)(&g_nccl.field) = dlsym(lib, name);                                    \
  if (!g_nccl.field) return fail(MRSB_ERR_NCCL, "libnccl lacks symbol %s", name);
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllGather, "ncclAllGather");
  SYM(Broadcast, "ncclBroadcast");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  g_nccl.lib = lib;
  return MRSB_OK;
}

#define NC(call)                                                                                                  \
  do {                                                                                                            \
    ncclResult_t r_ = (call);                                                                                     \
    if (r_ != ncclSuccess) return fail(MRSB_ERR_NCCL, "%s failed: %s", #call, g_nccl.GetErrorString(r_));          \
  } while (0)

// ------------------------------------------------------------------------------------------
// the handle
// ------------------------------------------------------------------------------------------
struct ParamSet {
  mrsb_model_params      mp;
  mrsb_controller_params cp;
};

struct mrsb_sim {
  int          device = 0;
  cudaStream_t stream = nullptr;
  DevState     ds{};
  DevGrid      grid{};

  std::vector<ParamSet>      sets;
  std::map<std::string, int> set_index;
  std::vector<int32_t>       pset_host;  // [n_global]
  DevParams*                 d_params     = nullptr;
  int                        d_params_cap = 0;
  int32_t*                   d_pset       = nullptr;
  bool                       params_dirty = true;

  char*    d_stage       = nullptr;  // staging for payload transposes
  size_t   d_stage_bytes = 0;
  int32_t* d_idx         = nullptr;
  size_t   d_idx_cap     = 0;

  int  uniform_mode = MRSB_INPUT_UNKNOWN;  // INPUT_MODE shared by all UAVs, or -1 if mixed
  int  uniform_nm   = 0;                   // n_motors shared by all local UAVs, or 0 if mixed
  int  uniform_pset = -1;                  // parameter set shared by all local UAVs, or -1 if mixed
  DevParams uniform_params{};              // host copy of that set (goes to the kernel by value)
  bool any_moment   = false;

  int    coll_enabled = 0, coll_crash = 0;
  double coll_rebounce = 0.0;
  void*  cub_tmp       = nullptr;
  size_t cub_tmp_bytes = 0;

  ncclComm_t           comm    = nullptr;
  int                  n_ranks = 1, rank = 0;
  std::vector<int64_t> shard_begin_of, shard_count_of;
  bool                 equal_shards = true;

  // fused position exchange over peer memory (set up by mrsb_comm_init_nccl when every peer is reachable)
  bool                 p2p       = false;
  double*              gbuf[2]   = {nullptr, nullptr};  // double-buffered gather buffer (parity flips every step)
  int                  parity    = 0;
  bool                 pushed    = false;  // the current buffer was last written by a pushing step kernel
  unsigned long long   epoch     = 0;
  double**             d_peers[2] = {nullptr, nullptr};  // device arrays [n_ranks] of the peers' gbuf[parity]
  unsigned long long*  d_flags   = nullptr;              // [n_ranks] written by the peers
  unsigned long long** d_peer_flags = nullptr;           // device array [n_ranks] of the peers' d_flags
  std::vector<void*>   ipc_opened;
  int*                 h_status  = nullptr;              // pinned, mapped: set by the wait kernel on time-out

  // pipelined host I/O (mrsb_set_input_async / mrsb_get_positions_async)
  cudaStream_t up_stream = nullptr, down_stream = nullptr;
  double*      d_up[2]   = {nullptr, nullptr};  // staged command rows
  size_t       d_up_bytes = 0;
  double*      d_snap[2] = {nullptr, nullptr};  // position snapshots [n_local][3]
  cudaEvent_t  ev_up[2] = {nullptr, nullptr}, ev_applied[2] = {nullptr, nullptr}, ev_snap[2] = {nullptr, nullptr}, ev_down[2] = {nullptr, nullptr};
  uint64_t     n_up = 0, n_down = 0;

  // the collision pass is a fixed set of launches with fixed arguments: replayed as a CUDA graph (one per
  // gather-buffer parity; with neighbour lists the table rebuild sits in a conditional node of it);
  // invalidated when its arguments change
  cudaGraphExec_t coll_graph[2]     = {nullptr, nullptr};
  int             coll_graph_own[2] = {0, 0};  // own kernels per replay (for the launch counter)

  // neighbour lists (single-shard handles): the table rebuild sits in a conditional node of the graph
  bool     lists_on            = false;
  int      steps_since_pass    = 0;     // stepping launches since the last collision pass
  bool     positions_touched   = true;  // positions were written by something else than ONE stepping launch
  int      rebuild_own         = 0;     // own kernels of one rebuild (for the launch counter)
  int64_t  list_passes         = 0;     // passes that went through decide_kernel (it counts them too: NlCtl::n_passes)
  bool     list_graph_failed   = false; // conditional graph nodes unavailable: do not try again on every pass
  int64_t  rebuilds_counted    = 0;
  uint32_t* h_one              = nullptr;  // pinned constant 1 (source of the async "force rebuild" copy)

  int64_t n_steps = 0, n_passes = 0, n_launches = 0;
  int     step_info[4] = {0, 0, 0, 0};  // last stepping launch: variant, grid, NM_T, MODE_T (mrsb_get_step_info)
};

static void drop_collision_graphs(mrsb_sim* h) {
  for (int k = 0; k < 2; k++) {
    if (h->coll_graph[k]) cudaGraphExecDestroy(h->coll_graph[k]);
    h->coll_graph[k] = nullptr;
  }
}

static std::string set_key(const ParamSet& s) {
  return std::string(reinterpret_cast<const char*>(&s), sizeof(ParamSet));
}

static int intern_set(mrsb_sim* h, const ParamSet& s) {
  const std::string key = set_key(s);
  auto              it  = h->set_index.find(key);
  if (it != h->set_index.end()) return it->second;
  const int id = int(h->sets.size());
  h->sets.push_back(s);
  h->set_index[key] = id;
  h->params_dirty   = true;
  return id;
}

static int check_params(const mrsb_model_params& p) {
  if (p.n_motors < 1 || p.n_motors > MRSB_MAX_MOTORS) return fail(MRSB_ERR_INVALID, "n_motors=%d outside 1..%d", p.n_motors, MRSB_MAX_MOTORS);
  return MRSB_OK;
}

static ParamSet canonical(const mrsb_model_params& mp, const mrsb_controller_params& cp) {
  ParamSet s;
  std::memset(&s, 0, sizeof(s));  // padding bytes take part in the key
  s.mp.n_motors              = mp.n_motors;
  s.mp.ground_enabled        = mp.ground_enabled != 0;
  s.mp.takeoff_patch_enabled = mp.takeoff_patch_enabled != 0;
  s.mp.g = mp.g, s.mp.mass = mp.mass, s.mp.kf = mp.kf, s.mp.km = mp.km, s.mp.prop_radius = mp.prop_radius, s.mp.arm_length = mp.arm_length;
  s.mp.body_height = mp.body_height, s.mp.motor_time_constant = mp.motor_time_constant, s.mp.max_rpm = mp.max_rpm, s.mp.min_rpm = mp.min_rpm;
  s.mp.air_resistance_coeff = mp.air_resistance_coeff, s.mp.ground_z = mp.ground_z;
  std::memcpy(s.mp.J, mp.J, sizeof(mp.J));
  for (int r = 0; r < 4; r++)
    for (int m = 0; m < mp.n_motors; m++) s.mp.allocation_matrix[r * MRSB_MAX_MOTORS + m] = mp.allocation_matrix[r * MRSB_MAX_MOTORS + m];
  s.cp.mixer_desaturation = cp.mixer_desaturation != 0;
  s.cp.rate_kp = cp.rate_kp, s.cp.rate_kd = cp.rate_kd, s.cp.rate_ki = cp.rate_ki;
  s.cp.att_kp = cp.att_kp, s.cp.att_kd = cp.att_kd, s.cp.att_ki = cp.att_ki, s.cp.att_max_rate_roll_pitch = cp.att_max_rate_roll_pitch,
  s.cp.att_max_rate_yaw = cp.att_max_rate_yaw;
  s.cp.vel_kp = cp.vel_kp, s.cp.vel_kd = cp.vel_kd, s.cp.vel_ki = cp.vel_ki, s.cp.vel_max_acceleration = cp.vel_max_acceleration;
  s.cp.pos_kp = cp.pos_kp, s.cp.pos_kd = cp.pos_kd, s.cp.pos_ki = cp.pos_ki, s.cp.pos_max_velocity = cp.pos_max_velocity;
  return s;
}

// upload the parameter table if it changed; recompute the launch specialisation hints
static int flush_params(mrsb_sim* h) {
  if (!h->params_dirty) return MRSB_OK;
  const int n_sets = int(h->sets.size());
  if (n_sets > h->d_params_cap) {
    CU(cudaStreamSynchronize(h->stream));
    if (h->d_params) CU(cudaFree(h->d_params));
    h->d_params_cap = std::max(16, 2 * n_sets);
    CU(cudaMalloc(&h->d_params, sizeof(DevParams) * h->d_params_cap));
    h->ds.params = h->d_params;
    drop_collision_graphs(h);
  }
  std::vector<DevParams> host(n_sets);
  for (int k = 0; k < n_sets; k++) mrsb_derive(h->sets[k].mp, h->sets[k].cp, &host[k]);
  CU(cudaMemcpyAsync(h->d_params, host.data(), sizeof(DevParams) * n_sets, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));  // `host` goes out of scope
  int nm = -1, ps = -2;
  for (int64_t i = 0; i < h->ds.n; i++) {
    const int id = h->pset_host[h->ds.shard_begin + i];
    const int m  = h->sets[id].mp.n_motors;
    if (nm == -1) nm = m;
    if (nm != m) nm = 0;
    if (ps == -2) ps = id;
    if (ps != id) ps = -1;
    if (nm == 0 && ps == -1) break;
  }
  h->uniform_nm   = nm > 0 ? nm : 0;
  h->uniform_pset = ps >= 0 ? ps : -1;
  if (h->uniform_pset >= 0) h->uniform_params = host[size_t(h->uniform_pset)];
  h->params_dirty = false;
  return MRSB_OK;
}

static int ensure_stage(mrsb_sim* h, size_t bytes) {
  if (bytes <= h->d_stage_bytes) return MRSB_OK;
  CU(cudaStreamSynchronize(h->stream));
  if (h->d_stage) CU(cudaFree(h->d_stage));
  h->d_stage_bytes = std::max<size_t>(bytes, 1 << 16);
  CU(cudaMalloc(&h->d_stage, h->d_stage_bytes));
  return MRSB_OK;
}

// copy an index list to the device (nullptr stays nullptr = identity) after validating it
static int stage_idx(mrsb_sim* h, int64_t n, const int32_t* idx, const int32_t** out) {
  *out = nullptr;
  if (n < 0 || n > h->ds.n && !idx) return fail(MRSB_ERR_INVALID, "n=%lld outside 0..%lld", (long long)n, (long long)h->ds.n);
  if (!idx || n == 0) return MRSB_OK;
  for (int64_t k = 0; k < n; k++)
    if (idx[k] < 0 || idx[k] >= h->ds.n) return fail(MRSB_ERR_INVALID, "idx[%lld]=%d outside 0..%lld", (long long)k, idx[k], (long long)h->ds.n - 1);
  if (size_t(n) > h->d_idx_cap) {
    CU(cudaStreamSynchronize(h->stream));
    if (h->d_idx) CU(cudaFree(h->d_idx));
    h->d_idx_cap = std::max<size_t>(size_t(n), 1024);
    CU(cudaMalloc(&h->d_idx, sizeof(int32_t) * h->d_idx_cap));
  }
  CU(cudaMemcpyAsync(h->d_idx, idx, sizeof(int32_t) * n, cudaMemcpyHostToDevice, h->stream));
  *out = h->d_idx;
  return MRSB_OK;
}

#define GUARD(h)                                              \
  if (!(h)) return fail(MRSB_ERR_INVALID, "null handle");     \
  CU(cudaSetDevice((h)->device));

static int64_t round_up(int64_t v, int64_t m) {
  return (v + m - 1) / m * m;
}

template <class T>
static int dalloc(T** p, size_t count) {
  CU(cudaMalloc(p, sizeof(T) * std::max<size_t>(count, 1)));
  CU(cudaMemset(*p, 0, sizeof(T) * std::max<size_t>(count, 1)));
  return MRSB_OK;
}

static void note_mode(mrsb_sim* h, int mode, int64_t n, bool all) {
  if (all && n == h->ds.n) {
    h->uniform_mode = mode;
  } else if (n > 0 && h->uniform_mode != mode) {
    h->uniform_mode = -1;
  }
}

static int stride_of(int mode) {
  switch (mode) {
    case MRSB_ACTUATOR_CMD:
      return MRSB_MAX_MOTORS;
    case MRSB_ATTITUDE_CMD:
      return 10;
    case MRSB_TILT_HDG_RATE_CMD:
      return 5;
    default:
      return 4;
  }
}

// re-point the addressed UAVs to (possibly new) parameter sets produced by `edit`
template <class Edit>
static int repoint(mrsb_sim* h, int64_t n, const int32_t* idx, Edit edit) {
  const int32_t* d_idx = nullptr;
  int            rc    = stage_idx(h, n, idx, &d_idx);
  if (rc) return rc;
  if (n == 0) return MRSB_OK;
  std::vector<int32_t> ids(static_cast<size_t>(n), 0);
  std::map<int, int>   memo;
  for (int64_t k = 0; k < n; k++) {
    const int64_t i   = idx ? idx[k] : k;
    const int     old = h->pset_host[size_t(h->ds.shard_begin + i)];
    auto          it  = memo.find(old);
    int           id;
    if (it != memo.end()) {
      id = it->second;
    } else {
      ParamSet s = h->sets[old];
      edit(s);
      id        = intern_set(h, canonical(s.mp, s.cp));
      memo[old] = id;
    }
    ids[size_t(k)]                                 = id;
    h->pset_host[size_t(h->ds.shard_begin + i)] = id;
  }
  h->params_dirty = true;
  rc              = ensure_stage(h, sizeof(int32_t) * size_t(n));
  if (rc) return rc;
  CU(cudaMemcpyAsync(h->d_stage, ids.data(), sizeof(int32_t) * size_t(n), cudaMemcpyHostToDevice, h->stream));
  h->n_launches += launch_set_pset(h->d_pset, n, d_idx, h->ds.shard_begin, reinterpret_cast<const int32_t*>(h->d_stage), h->stream);
  CU(cudaStreamSynchronize(h->stream));  // `ids` goes out of scope
  return MRSB_OK;
}

// Fused exchange set-up: a second gather buffer, this rank's flag slots, and IPC mappings of every
// peer's two buffers and flags (handles travel through one NCCL all-gather of raw bytes).
// Cell geometry of the spatial hash.  Handles that run the full pass every tick use the smallest cell
// whose half covers the search radius sqrt(3) (4 m: fewest candidates).  Handles that keep neighbour
// lists between rebuilds use a larger cell: the list radius is just under cell / 2, so a larger cell
// buys a larger skin, i.e. fewer rebuilds for more candidates per rebuild.
static void set_collision_geometry(mrsb_sim* h, bool lists) {
  DevGrid& g  = h->grid;
  h->lists_on = lists;
  double cell = lists ? 6.0 : 4.0;
  if (const char* e = getenv("MRSB_COLLISION_CELL")) cell = std::max(3.5, atof(e));
  g.inv_cell = 1.0 / cell;
  g.reach    = 0.5 * cell;
  const double r_list = g.reach * (1.0 - 1e-6);
  g.list_r2  = r_list * r_list;
  g.skin     = (r_list - 1.7320508075688775) * (1.0 - 1e-6);
  h->ds.disp_max       = lists ? &g.ctl->disp_max_bits : nullptr;
  h->positions_touched = true;
}

static int setup_p2p(mrsb_sim* h) {
  const int    G     = h->n_ranks;
  const size_t bytes = sizeof(double) * 3 * size_t(h->ds.n_global);
  h->gbuf[0]         = h->ds.gpos;
  CU(cudaMalloc(&h->gbuf[1], std::max<size_t>(bytes, 16)));
  CU(cudaMemcpy(h->gbuf[1], h->gbuf[0], bytes, cudaMemcpyDeviceToDevice));
  CU(cudaMalloc(&h->d_flags, sizeof(unsigned long long) * 3 * G));  // epochs [G] + displacement words [2G] (step_kernel.cu)
  CU(cudaMemset(h->d_flags, 0, sizeof(unsigned long long) * 3 * G));
  CU(cudaHostAlloc(&h->h_status, sizeof(int), cudaHostAllocMapped));
  *h->h_status = 0;
  struct Handles {
    cudaIpcMemHandle_t buf[2], flags;
  };
  Handles mine;
  CU(cudaIpcGetMemHandle(&mine.buf[0], h->gbuf[0]));
  CU(cudaIpcGetMemHandle(&mine.buf[1], h->gbuf[1]));
  CU(cudaIpcGetMemHandle(&mine.flags, h->d_flags));
  Handles* d_all = nullptr;
  CU(cudaMalloc(&d_all, sizeof(Handles) * G));
  CU(cudaMemcpyAsync(d_all + h->rank, &mine, sizeof(Handles), cudaMemcpyHostToDevice, h->stream));
  NC(g_nccl.AllGather(d_all + h->rank, d_all, sizeof(Handles), ncclChar, h->comm, h->stream));
  std::vector<Handles> all;
  all.resize(static_cast<size_t>(G));
  CU(cudaMemcpyAsync(all.data(), d_all, sizeof(Handles) * G, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaFree(d_all));
  std::vector<double*>             peers0(static_cast<size_t>(G), nullptr), peers1(static_cast<size_t>(G), nullptr);
  std::vector<unsigned long long*> pflags(static_cast<size_t>(G), nullptr);
  for (int r = 0; r < G; r++) {
    if (r == h->rank) {
      peers0[size_t(r)] = h->gbuf[0];
      peers1[size_t(r)] = h->gbuf[1];
      pflags[size_t(r)] = h->d_flags;
      continue;
    }
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, all[size_t(r)].buf[0], cudaIpcMemLazyEnablePeerAccess));
    h->ipc_opened.push_back(p);
    peers0[size_t(r)] = static_cast<double*>(p);
    CU(cudaIpcOpenMemHandle(&p, all[size_t(r)].buf[1], cudaIpcMemLazyEnablePeerAccess));
    h->ipc_opened.push_back(p);
    peers1[size_t(r)] = static_cast<double*>(p);
    CU(cudaIpcOpenMemHandle(&p, all[size_t(r)].flags, cudaIpcMemLazyEnablePeerAccess));
    h->ipc_opened.push_back(p);
    pflags[size_t(r)] = static_cast<unsigned long long*>(p);
  }
  CU(cudaMalloc(&h->d_peers[0], sizeof(double*) * G));
  CU(cudaMalloc(&h->d_peers[1], sizeof(double*) * G));
  CU(cudaMalloc(&h->d_peer_flags, sizeof(unsigned long long*) * G));
  CU(cudaMemcpy(h->d_peers[0], peers0.data(), sizeof(double*) * G, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(h->d_peers[1], peers1.data(), sizeof(double*) * G, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(h->d_peer_flags, pflags.data(), sizeof(unsigned long long*) * G, cudaMemcpyHostToDevice));
  // everybody must have opened everybody's memory before the first push: one more (tiny) collective
  unsigned long long* d_tmp = nullptr;
  CU(cudaMalloc(&d_tmp, sizeof(unsigned long long) * G));
  NC(g_nccl.AllGather(d_tmp + h->rank, d_tmp, sizeof(unsigned long long), ncclChar, h->comm, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaFree(d_tmp));
  h->parity   = 0;
  h->ds.peers = nullptr;  // set per step
  h->p2p      = true;
  // the hand-shake also carries every rank's displacement bound: neighbour lists work across shards
  set_collision_geometry(h, h->grid.nl_count != nullptr);
  drop_collision_graphs(h);
  return MRSB_OK;
}

extern "C" {

const char* mrsb_last_error(void) {
  return g_err;
}
int mrsb_version(void) {
  return MRSB_VERSION_MAJOR * 1000 + MRSB_VERSION_MINOR;
}

// ------------------------------------------------------------------------------------------
// lifetime
// ------------------------------------------------------------------------------------------
int mrsb_destroy(mrsb_handle h) {
  if (!h) return MRSB_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->up_stream) cudaStreamSynchronize(h->up_stream);
  if (h->down_stream) cudaStreamSynchronize(h->down_stream);
  for (int k = 0; k < 2; k++) {
    if (h->d_up[k]) cudaFree(h->d_up[k]);
    if (h->d_snap[k]) cudaFree(h->d_snap[k]);
    for (cudaEvent_t e : {h->ev_up[k], h->ev_applied[k], h->ev_snap[k], h->ev_down[k]})
      if (e) cudaEventDestroy(e);
  }
  if (h->up_stream) cudaStreamDestroy(h->up_stream);
  if (h->down_stream) cudaStreamDestroy(h->down_stream);
  drop_collision_graphs(h);
  for (void* p : h->ipc_opened) cudaIpcCloseMemHandle(p);
  if (h->gbuf[1] && h->gbuf[1] != h->ds.gpos) cudaFree(h->gbuf[1]);
  if (h->gbuf[0] && h->gbuf[0] != h->ds.gpos) cudaFree(h->gbuf[0]);
  for (void* p : {(void*)h->d_peers[0], (void*)h->d_peers[1], (void*)h->d_flags, (void*)h->d_peer_flags})
    if (p) cudaFree(p);
  if (h->h_status) cudaFreeHost(h->h_status);
  if (h->h_one) cudaFreeHost(h->h_one);
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  void* ptrs[] = {h->ds.st,     h->ds.vprev,  h->ds.rpm,      h->ds.pid,      h->ds.fext,         h->ds.mext,       h->ds.imu,    h->ds.initz,
                  h->ds.cmd,    h->ds.ff,     h->ds.flags,    h->ds.mode,     h->ds.gpos,         h->d_params,      h->d_pset,    h->d_stage,
                  h->d_idx,     h->grid.bucket, h->grid.rank, h->grid.count, h->grid.aabb, h->grid.begin, h->grid.rec, h->grid.pairs,
                  h->grid.counters, h->cub_tmp, h->grid.nl_count, h->grid.nl_items, h->grid.nl_active, h->grid.ctl};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return MRSB_OK;
}

int mrsb_create(const mrsb_create_info* info, mrsb_handle* out) {
  if (!info || !out) return fail(MRSB_ERR_INVALID, "null argument");
  *out = nullptr;
  if (info->n_types < 1 || !info->types) return fail(MRSB_ERR_INVALID, "need at least one airframe type");
  if (info->n_local < 0 || info->n_global < info->n_local || info->shard_begin < 0 || info->shard_begin + info->n_local > info->n_global)
    return fail(MRSB_ERR_INVALID, "inconsistent shard: n_local=%lld n_global=%lld shard_begin=%lld", (long long)info->n_local,
                (long long)info->n_global, (long long)info->shard_begin);
  if (info->n_global > 0x7fffffffLL) return fail(MRSB_ERR_INVALID, "n_global exceeds int32 indices");
  for (int t = 0; t < info->n_types; t++) {
    const int rc = check_params(info->types[t]);
    if (rc) return rc;
  }
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return fail(MRSB_ERR_CUDA, "no CUDA device available (libmrsb has no CPU fallback)");
  if (info->device < 0 || info->device >= n_dev) return fail(MRSB_ERR_INVALID, "device %d outside 0..%d", info->device, n_dev - 1);

  mrsb_sim* h = new mrsb_sim();
  h->device   = info->device;
#define CREATE_CU(call)                      \
  do {                                       \
    int rc_ = [&]() -> int {                 \
      CU(call);                              \
      return MRSB_OK;                        \
    }();                                     \
    if (rc_) {                               \
      mrsb_destroy(h);                       \
      return rc_;                            \
    }                                        \
  } while (0)
#define CREATE_RC(expr)  \
  do {                   \
    int rc_ = (expr);    \
    if (rc_) {           \
      mrsb_destroy(h);   \
      return rc_;        \
    }                    \
  } while (0)

  CREATE_CU(cudaSetDevice(h->device));
  CREATE_CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));

  DevState& s   = h->ds;
  s.n           = info->n_local;
  s.ld          = std::max<int64_t>(128, round_up(info->n_local, 128));
  s.n_global    = info->n_global;
  s.shard_begin = info->shard_begin;
  const size_t ld = size_t(s.ld);
  CREATE_RC(dalloc(&s.st, 18 * ld));
  CREATE_RC(dalloc(&s.vprev, 3 * ld));
  CREATE_RC(dalloc(&s.rpm, MRSB_NM * ld));
  CREATE_RC(dalloc(&s.pid, PID_ROWS * ld));
  CREATE_RC(dalloc(&s.fext, 3 * ld));
  CREATE_RC(dalloc(&s.mext, 3 * ld));
  CREATE_RC(dalloc(&s.imu, 3 * ld));
  CREATE_RC(dalloc(&s.initz, ld));
  CREATE_RC(dalloc(&s.cmd, CMD_ROWS * ld));
  CREATE_RC(dalloc(&s.ff, FF_ROWS * ld));
  CREATE_RC(dalloc(&s.flags, ld));
  CREATE_RC(dalloc(&s.mode, ld));
  CREATE_RC(dalloc(&s.gpos, 3 * size_t(s.n_global)));
  CREATE_RC(dalloc(&h->d_pset, size_t(s.n_global)));
  s.pset = h->d_pset;

  // parameter sets: one per airframe type with default controller gains (US:159-169)
  mrsb_controller_params cp;
  mrsb_controller_params_default(&cp);
  std::vector<int> set_of_type(info->n_types);
  for (int t = 0; t < info->n_types; t++) set_of_type[t] = intern_set(h, canonical(info->types[t], cp));
  h->pset_host.resize(size_t(s.n_global));
  for (int64_t j = 0; j < s.n_global; j++) {
    const int t = info->type_of_uav ? info->type_of_uav[j] : 0;
    if (t < 0 || t >= info->n_types) {
      mrsb_destroy(h);
      return fail(MRSB_ERR_INVALID, "type_of_uav[%lld]=%d outside 0..%d", (long long)j, t, info->n_types - 1);
    }
    h->pset_host[size_t(j)] = set_of_type[t];
  }
  CREATE_CU(cudaMemcpy(h->d_pset, h->pset_host.data(), sizeof(int32_t) * size_t(s.n_global), cudaMemcpyHostToDevice));
  CREATE_RC(flush_params(h));

  // initial state (MM:183-198 + setStatePos MM:439-446): R = Rz(-heading), flags from the airframe
  {
    std::vector<uint32_t> fl(ld, 0u);
    for (int64_t i = 0; i < s.n; i++)
      fl[size_t(i)] = h->sets[h->pset_host[size_t(s.shard_begin + i)]].mp.takeoff_patch_enabled ? FLAG_TAKEOFF : 0u;
    CREATE_CU(cudaMemcpy(s.flags, fl.data(), sizeof(uint32_t) * ld, cudaMemcpyHostToDevice));
    const size_t bytes = sizeof(double) * 4 * size_t(std::max<int64_t>(s.n, 1));
    CREATE_RC(ensure_stage(h, bytes));
    double* d_xyz = nullptr;
    double* d_hdg = nullptr;
    if (info->spawn_xyz && s.n) {
      d_xyz = reinterpret_cast<double*>(h->d_stage);
      CREATE_CU(cudaMemcpyAsync(d_xyz, info->spawn_xyz, sizeof(double) * 3 * s.n, cudaMemcpyHostToDevice, h->stream));
    }
    if (info->spawn_heading && s.n) {
      d_hdg = reinterpret_cast<double*>(h->d_stage) + 3 * s.n;
      CREATE_CU(cudaMemcpyAsync(d_hdg, info->spawn_heading, sizeof(double) * s.n, cudaMemcpyHostToDevice, h->stream));
    }
    h->n_launches += launch_set_state_pos(s, s.n, nullptr, d_xyz, d_hdg, h->stream);
    CREATE_CU(cudaStreamSynchronize(h->stream));
  }

  // collision workspace
  {
    DevGrid&     g  = h->grid;
    const size_t ng = size_t(std::max<int64_t>(s.n_global, 1));
    // >= 2 buckets per inserted UAV.  A shard inserts its own UAVs plus the halo of remote ones near its
    // bounding box; the table is sized for a halo of up to 3x the shard (a fuller table only means
    // longer bucket lists, never wrong results)
    const size_t expect = std::min(ng, std::max<size_t>(4 * size_t(std::max<int64_t>(s.n, 1)), 512));
    uint32_t     bits   = 10;
    while ((size_t(1) << bits) < 2 * expect && bits < 30) bits++;
    g.bits      = bits;
    g.n_buckets = 1u << bits;
    CREATE_RC(dalloc(&g.bucket, ng));
    CREATE_RC(dalloc(&g.rank, ng));
    CREATE_RC(dalloc(&g.count, size_t(g.n_buckets) + 4));  // + mirrors of buckets 0 and 1 + sentinel
    CREATE_RC(dalloc(&g.begin, size_t(g.n_buckets) + 4));
    CREATE_RC(dalloc(&g.aabb, 6));
    CREATE_RC(dalloc(&g.rec, 2 * ng));  // worst case: every UAV in bucket 0 and mirrored
    g.pair_cap = int64_t(std::max<size_t>(4096, 4 * size_t(std::max<int64_t>(s.n, 1))));
    CREATE_RC(dalloc(&g.pairs, 2 * size_t(g.pair_cap)));
    CREATE_RC(dalloc(&g.counters, 4));
    h->cub_tmp_bytes = collide_tmp_bytes(int64_t(g.n_buckets) + 3, s.n);
    CREATE_CU(cudaMalloc(&h->cub_tmp, std::max<size_t>(h->cub_tmp_bytes, 16)));
    // neighbour lists: single-shard handles now, sharded ones once the fused exchange is up (setup_p2p)
    if (s.n > 0 && !getenv("MRSB_NO_NEIGHBOUR_LISTS")) {
      g.nl_ld = s.ld;
      CREATE_RC(dalloc(&g.nl_count, size_t(g.nl_ld)));
      CREATE_RC(dalloc(&g.nl_items, size_t(MRSB_NL_CAP) * size_t(g.nl_ld)));
      CREATE_RC(dalloc(&g.nl_active, size_t(g.nl_ld)));
      CREATE_RC(dalloc(&g.ctl, 1));
      CREATE_CU(cudaMemsetAsync(g.ctl, 0, sizeof(NlCtl), h->stream));
      CREATE_CU(cudaHostAlloc(&h->h_one, 2 * sizeof(uint32_t), cudaHostAllocDefault));
      h->h_one[0] = 1u;
      h->h_one[1] = 0xFFFFFFFFu;  // "unbounded displacement"
    }
    set_collision_geometry(h, g.nl_count != nullptr && s.n_global == s.n);
  }
  h->shard_begin_of = {s.shard_begin};
  h->shard_count_of = {s.n};
  CREATE_CU(cudaStreamSynchronize(h->stream));
  *out = h;
  return MRSB_OK;
}

int mrsb_sync(mrsb_handle h) {
  GUARD(h);
  CU(cudaStreamSynchronize(h->stream));
  if (h->up_stream) CU(cudaStreamSynchronize(h->up_stream));
  if (h->down_stream) CU(cudaStreamSynchronize(h->down_stream));
  return MRSB_OK;
}

static int ensure_pipeline(mrsb_sim* h) {
  if (h->up_stream) return MRSB_OK;
  CU(cudaStreamCreateWithFlags(&h->up_stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&h->down_stream, cudaStreamNonBlocking));
  for (int k = 0; k < 2; k++) {
    CU(cudaEventCreateWithFlags(&h->ev_up[k], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_applied[k], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_snap[k], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_down[k], cudaEventDisableTiming));
    CU(cudaMalloc(&h->d_snap[k], sizeof(double) * 3 * size_t(std::max<int64_t>(h->ds.n, 1))));
  }
  return MRSB_OK;
}

int mrsb_set_input_async(mrsb_handle h, int32_t mode, const double* payload, int32_t stride) {
  GUARD(h);
  if (mode <= MRSB_INPUT_UNKNOWN || mode > MRSB_POSITION_CMD) return fail(MRSB_ERR_INVALID, "mode %d carries no payload", mode);
  if (!payload) return fail(MRSB_ERR_INVALID, "null payload");
  if (stride < (mode == MRSB_ACTUATOR_CMD ? 1 : stride_of(mode))) return fail(MRSB_ERR_INVALID, "stride %d too small for mode %d", stride, mode);
  int rc = ensure_pipeline(h);
  if (rc) return rc;
  const int64_t n     = h->ds.n;
  const size_t  bytes = sizeof(double) * size_t(n) * size_t(stride);
  if (bytes > h->d_up_bytes) {
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaStreamSynchronize(h->up_stream));
    for (int k = 0; k < 2; k++) {
      if (h->d_up[k]) CU(cudaFree(h->d_up[k]));
      CU(cudaMalloc(&h->d_up[k], std::max<size_t>(bytes, 256)));
    }
    h->d_up_bytes = bytes;
    h->n_up       = 0;
  }
  const int k = int(h->n_up & 1);
  if (h->n_up >= 2) CU(cudaStreamWaitEvent(h->up_stream, h->ev_applied[k], 0));  // the staging buffer was consumed two uploads ago
  CU(cudaMemcpyAsync(h->d_up[k], payload, bytes, cudaMemcpyHostToDevice, h->up_stream));
  CU(cudaEventRecord(h->ev_up[k], h->up_stream));
  CU(cudaStreamWaitEvent(h->stream, h->ev_up[k], 0));
  h->n_launches += launch_scatter_input(h->ds, mode, n, nullptr, h->d_up[k], stride, h->stream);
  CU(cudaEventRecord(h->ev_applied[k], h->stream));
  h->n_up++;
  note_mode(h, mode, n, true);
  CU(cudaGetLastError());
  return MRSB_OK;
}

int mrsb_get_positions_async(mrsb_handle h, double* out_xyz) {
  GUARD(h);
  if (!out_xyz) return fail(MRSB_ERR_INVALID, "null output");
  int rc = ensure_pipeline(h);
  if (rc) return rc;
  const int    k     = int(h->n_down & 1);
  const size_t bytes = sizeof(double) * 3 * size_t(h->ds.n);
  if (h->n_down >= 2) CU(cudaStreamWaitEvent(h->stream, h->ev_down[k], 0));  // the snapshot buffer was downloaded two reads ago
  CU(cudaMemcpyAsync(h->d_snap[k], h->ds.gpos + 3 * h->ds.shard_begin, bytes, cudaMemcpyDeviceToDevice, h->stream));
  CU(cudaEventRecord(h->ev_snap[k], h->stream));
  CU(cudaStreamWaitEvent(h->down_stream, h->ev_snap[k], 0));
  CU(cudaMemcpyAsync(out_xyz, h->d_snap[k], bytes, cudaMemcpyDeviceToHost, h->down_stream));
  CU(cudaEventRecord(h->ev_down[k], h->down_stream));
  h->n_down++;
  return MRSB_OK;
}

int mrsb_wait_uploads(mrsb_handle h) {
  GUARD(h);
  if (h->up_stream) CU(cudaStreamSynchronize(h->up_stream));
  return MRSB_OK;
}
int mrsb_wait_downloads(mrsb_handle h) {
  GUARD(h);
  if (h->down_stream) CU(cudaStreamSynchronize(h->down_stream));
  return MRSB_OK;
}
int64_t mrsb_n_local(mrsb_handle h) {
  return h ? h->ds.n : -1;
}
int64_t mrsb_n_global(mrsb_handle h) {
  return h ? h->ds.n_global : -1;
}
void* mrsb_get_stream(mrsb_handle h) {
  return h ? (void*)h->stream : nullptr;
}

// ------------------------------------------------------------------------------------------
// commands
// ------------------------------------------------------------------------------------------
int mrsb_set_input_device(mrsb_handle h, int32_t mode, int64_t n, const int32_t* idx_dev, const double* payload_dev, int32_t stride) {
  GUARD(h);
  if (mode < MRSB_INPUT_UNKNOWN || mode > MRSB_POSITION_CMD) return fail(MRSB_ERR_INVALID, "mode %d is not an INPUT_MODE", mode);
  if (n < 0 || (!idx_dev && n > h->ds.n)) return fail(MRSB_ERR_INVALID, "n=%lld outside 0..%lld", (long long)n, (long long)h->ds.n);
  if (mode == MRSB_INPUT_UNKNOWN) {
    h->n_launches += launch_set_mode(h->ds, n, idx_dev, mode, h->stream);
  } else {
    if (!payload_dev) return fail(MRSB_ERR_INVALID, "null payload");
    if (stride < (mode == MRSB_ACTUATOR_CMD ? 1 : stride_of(mode))) return fail(MRSB_ERR_INVALID, "stride %d too small for mode %d", stride, mode);
    h->n_launches += launch_scatter_input(h->ds, mode, n, idx_dev, payload_dev, stride, h->stream);
  }
  note_mode(h, mode, n, idx_dev == nullptr);
  CU(cudaGetLastError());
  return MRSB_OK;
}

int mrsb_set_input(mrsb_handle h, int32_t mode, int64_t n, const int32_t* idx, const double* payload, int32_t stride) {
  GUARD(h);
  if (mode < MRSB_INPUT_UNKNOWN || mode > MRSB_POSITION_CMD) return fail(MRSB_ERR_INVALID, "mode %d is not an INPUT_MODE", mode);
  const int32_t* d_idx = nullptr;
  int            rc    = stage_idx(h, n, idx, &d_idx);
  if (rc) return rc;
  if (mode == MRSB_INPUT_UNKNOWN) {
    h->n_launches += launch_set_mode(h->ds, n, d_idx, mode, h->stream);
    note_mode(h, mode, n, idx == nullptr);
    return MRSB_OK;
  }
  if (!payload && n) return fail(MRSB_ERR_INVALID, "null payload");
  if (stride < (mode == MRSB_ACTUATOR_CMD ? 1 : stride_of(mode))) return fail(MRSB_ERR_INVALID, "stride %d too small for mode %d", stride, mode);
  if (n == 0) return MRSB_OK;
  const size_t bytes = sizeof(double) * size_t(n) * size_t(stride);
  rc                 = ensure_stage(h, bytes);
  if (rc) return rc;
  CU(cudaMemcpyAsync(h->d_stage, payload, bytes, cudaMemcpyHostToDevice, h->stream));
  h->n_launches += launch_scatter_input(h->ds, mode, n, d_idx, reinterpret_cast<const double*>(h->d_stage), stride, h->stream);
  note_mode(h, mode, n, idx == nullptr);
  CU(cudaGetLastError());
  return MRSB_OK;
}

#define SET_INPUT(name, MODE, STRIDE)                                                                        \
  int mrsb_set_input_##name(mrsb_handle h, int64_t n, const int32_t* idx, const double* payload) {           \
    return mrsb_set_input(h, MODE, n, idx, payload, STRIDE);                                                 \
  }
SET_INPUT(actuators, MRSB_ACTUATOR_CMD, MRSB_MAX_MOTORS)
SET_INPUT(control_group, MRSB_CONTROL_GROUP_CMD, 4)
SET_INPUT(attitude_rate, MRSB_ATTITUDE_RATE_CMD, 4)
SET_INPUT(attitude, MRSB_ATTITUDE_CMD, 10)
SET_INPUT(tilt_hdg_rate, MRSB_TILT_HDG_RATE_CMD, 5)
SET_INPUT(acceleration_hdg_rate, MRSB_ACCELERATION_HDG_RATE_CMD, 4)
SET_INPUT(acceleration_hdg, MRSB_ACCELERATION_HDG_CMD, 4)
SET_INPUT(velocity_hdg_rate, MRSB_VELOCITY_HDG_RATE_CMD, 4)
SET_INPUT(velocity_hdg, MRSB_VELOCITY_HDG_CMD, 4)
SET_INPUT(position, MRSB_POSITION_CMD, 4)
#undef SET_INPUT

int mrsb_clear_input(mrsb_handle h, int64_t n, const int32_t* idx) {
  return mrsb_set_input(h, MRSB_INPUT_UNKNOWN, n, idx, nullptr, 0);
}

// host rows -> SoA rows [row0, row0+rows) of `dst`; then OR `flag` into the per-UAV flags
static int put_rows(mrsb_sim* h, double* dst, int rows_total, int row0, int rows, int64_t n, const int32_t* idx, const double* payload, int stride,
                    uint32_t or_flag) {
  const int32_t* d_idx = nullptr;
  int            rc    = stage_idx(h, n, idx, &d_idx);
  if (rc) return rc;
  if (n == 0) return MRSB_OK;
  if (!payload) return fail(MRSB_ERR_INVALID, "null payload");
  const size_t bytes = sizeof(double) * size_t(n) * size_t(stride);
  rc                 = ensure_stage(h, bytes);
  if (rc) return rc;
  CU(cudaMemcpyAsync(h->d_stage, payload, bytes, cudaMemcpyHostToDevice, h->stream));
  h->n_launches += launch_scatter_rows(dst, rows_total, row0, rows, n, d_idx, reinterpret_cast<const double*>(h->d_stage), stride, h->stream);
  if (or_flag) h->n_launches += launch_flag_update(h->ds, n, d_idx, 0xffffffffu, or_flag, h->stream);
  CU(cudaGetLastError());
  return MRSB_OK;
}

static int get_rows(mrsb_sim* h, const double* src, int rows_total, int row0, int rows, int64_t n, const int32_t* idx, double* out) {
  const int32_t* d_idx = nullptr;
  int            rc    = stage_idx(h, n, idx, &d_idx);
  if (rc) return rc;
  if (n == 0) return MRSB_OK;
  const size_t bytes = sizeof(double) * size_t(n) * size_t(rows);
  rc                 = ensure_stage(h, bytes);
  if (rc) return rc;
  h->n_launches += launch_gather_rows(src, rows_total, row0, rows, n, d_idx, reinterpret_cast<double*>(h->d_stage), rows, h->stream);
  CU(cudaMemcpyAsync(out, h->d_stage, bytes, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return MRSB_OK;
}

int mrsb_set_feedforward_acceleration_hdg_rate(mrsb_handle h, int64_t n, const int32_t* idx, const double* payload) {
  GUARD(h);
  return put_rows(h, h->ds.ff, FF_ROWS, FF_ACC_HDG_RATE, 4, n, idx, payload, 4, FLAG_FF_ACC_HDG_RATE);
}
int mrsb_set_feedforward_acceleration_hdg(mrsb_handle h, int64_t n, const int32_t* idx, const double* payload) {
  GUARD(h);
  return put_rows(h, h->ds.ff, FF_ROWS, FF_ACC_HDG, 3, n, idx, payload, 4, FLAG_FF_ACC_HDG);
}
int mrsb_set_feedforward_velocity_hdg(mrsb_handle h, int64_t n, const int32_t* idx, const double* payload) {
  GUARD(h);
  return put_rows(h, h->ds.ff, FF_ROWS, FF_VEL_HDG, 3, n, idx, payload, 4, FLAG_FF_VEL_HDG);
}
int mrsb_set_feedforward_velocity_hdg_rate(mrsb_handle h, int64_t n, const int32_t* idx, const double* payload) {
  GUARD(h);
  return put_rows(h, h->ds.ff, FF_ROWS, FF_VEL_HDG_RATE, 3, n, idx, payload, 4, FLAG_FF_VEL_HDG_RATE);
}
int mrsb_set_tracker_cmd(mrsb_handle h, int64_t n, const int32_t* idx, const double* rows) {
  GUARD(h);
  if (n < 0 || (n > 0 && !rows)) return fail(MRSB_ERR_INVALID, "null payload");
  // uav_system_ros.cpp:995-1021
  std::vector<double> vel_hdg(4 * size_t(n)), vel_hdg_rate(4 * size_t(n)), acc_hdg(4 * size_t(n)), acc_hdg_rate(4 * size_t(n));
  for (int64_t k = 0; k < n; k++) {
    const double* r  = rows + MRSB_TRACKER_CMD_STRIDE * k;
    const bool    uh = r[7] != 0.0, uv = r[8] != 0.0, ur = r[9] != 0.0, ua = r[10] != 0.0;
    const double  v[3]  = {uh ? r[0] : 0.0, uh ? r[1] : 0.0, uv ? r[2] : 0.0};
    const double  a[3]  = {ua ? r[3] : 0.0, ua ? r[4] : 0.0, ua ? r[5] : 0.0};
    const double  rate  = ur ? r[6] : 0.0;
    for (int c = 0; c < 3; c++) {
      vel_hdg[4 * k + c] = vel_hdg_rate[4 * k + c] = v[c];
      acc_hdg[4 * k + c] = acc_hdg_rate[4 * k + c] = a[c];
    }
    vel_hdg[4 * k + 3] = acc_hdg[4 * k + 3] = 0.0;
    vel_hdg_rate[4 * k + 3] = acc_hdg_rate[4 * k + 3] = rate;
  }
  int rc = mrsb_set_feedforward_velocity_hdg(h, n, idx, vel_hdg.data());
  if (!rc) rc = mrsb_set_feedforward_velocity_hdg_rate(h, n, idx, vel_hdg_rate.data());
  if (!rc) rc = mrsb_set_feedforward_acceleration_hdg(h, n, idx, acc_hdg.data());
  if (!rc) rc = mrsb_set_feedforward_acceleration_hdg_rate(h, n, idx, acc_hdg_rate.data());
  return rc;
}
int mrsb_clear_feedforward(mrsb_handle h, int64_t n, const int32_t* idx) {
  GUARD(h);
  const int32_t* d_idx = nullptr;
  int            rc    = stage_idx(h, n, idx, &d_idx);
  if (rc) return rc;
  h->n_launches += launch_flag_update(h->ds, n, d_idx, ~(FLAG_FF_ACC_HDG | FLAG_FF_ACC_HDG_RATE | FLAG_FF_VEL_HDG | FLAG_FF_VEL_HDG_RATE), 0u, h->stream);
  return MRSB_OK;
}

// ------------------------------------------------------------------------------------------
// stepping
// ------------------------------------------------------------------------------------------
static int exchange_positions(mrsb_sim* h) {
  if (h->n_ranks <= 1) return MRSB_OK;
  if (!h->comm) return fail(MRSB_ERR_STATE, "sharded handle (n_global > n_local) without a communicator: call mrsb_comm_init_nccl, or use mrsb_gather_buffer + mrsb_handle_collisions_gathered");
  if (h->p2p && h->pushed) {
    // the step kernel already stored this shard's positions into every peer's buffer: only the
    // hand-shake "my epoch has landed" / "everybody's has" is left
    uint32_t* disp = h->lists_on ? &h->grid.ctl->disp_max_bits : nullptr;
    h->n_launches += launch_p2p_signal(h->d_peer_flags, h->n_ranks, h->rank, h->epoch, disp, h->ds.n > 0 ? 0xFFFFFFFFu : 0u, h->stream);
    h->n_launches += launch_p2p_wait(h->d_flags, h->n_ranks, h->rank, h->epoch, h->h_status, disp, h->stream);
    return MRSB_OK;
  }
  h->positions_touched = true;  // all-gather without the hand-shake: no swarm-wide displacement bound for this pass
  double* buf = h->ds.gpos;
  if (h->equal_shards) {
    NC(g_nccl.AllGather(buf + 3 * h->ds.shard_begin, buf, size_t(3 * h->ds.n), ncclDouble, h->comm, h->stream));
  } else {
    NC(g_nccl.GroupStart());
    for (int r = 0; r < h->n_ranks; r++) {
      double* p = buf + 3 * h->shard_begin_of[r];
      NC(g_nccl.Broadcast(p, p, size_t(3 * h->shard_count_of[r]), ncclDouble, r, h->comm, h->stream));
    }
    NC(g_nccl.GroupEnd());
  }
  return MRSB_OK;
}

// The pass with neighbour lists as ONE graph:  decide -> IF (rebuild) { table, lists } -> check.
// The IF node's condition is set on the device by decide_kernel (cudaGraphSetConditional).
static cudaGraphExec_t build_list_graph(mrsb_sim* h, int* own_fixed, int* own_rebuild) {
  cudaGraph_t     graph = nullptr;
  cudaGraphExec_t exec  = nullptr;
  cudaStream_t    side  = nullptr;
  bool            ok    = false;
  if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  do {
    cudaStreamCaptureStatus status;
    const cudaGraphNode_t*  deps  = nullptr;
    size_t                  n_dep = 0;
    if (cudaStreamGetCaptureInfo_v2(h->stream, &status, nullptr, &graph, &deps, &n_dep) != cudaSuccess || !graph) break;
    cudaGraphConditionalHandle handle;
    if (cudaGraphConditionalHandleCreate(&handle, graph, 0, cudaGraphCondAssignDefault) != cudaSuccess) break;
    *own_fixed = launch_collide_decide(h->grid, 0, handle, 1, h->stream);
    if (cudaStreamGetCaptureInfo_v2(h->stream, &status, nullptr, &graph, &deps, &n_dep) != cudaSuccess) break;
    cudaGraphNodeParams cp = {};
    cp.type                = cudaGraphNodeTypeConditional;
    cp.conditional.handle  = handle;
    cp.conditional.type    = cudaGraphCondTypeIf;
    cp.conditional.size    = 1;
    cudaGraphNode_t cond   = nullptr;
    if (cudaGraphAddNode(&cond, graph, deps, n_dep, &cp) != cudaSuccess) break;
    cudaGraph_t body = cp.conditional.phGraph_out[0];
    if (cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking) != cudaSuccess) break;
    if (cudaStreamBeginCaptureToGraph(side, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) != cudaSuccess) break;
    *own_rebuild = launch_collide_rebuild(h->ds, h->grid, h->cub_tmp, h->cub_tmp_bytes, side);
    cudaGraph_t body_out = nullptr;
    if (cudaStreamEndCapture(side, &body_out) != cudaSuccess) break;
    if (cudaStreamUpdateCaptureDependencies(h->stream, &cond, 1, cudaStreamSetCaptureDependencies) != cudaSuccess) break;
    *own_fixed += launch_collide_check(h->ds, h->grid, h->coll_crash, h->coll_rebounce, h->stream);
    ok = true;
  } while (false);
  cudaGraph_t captured = nullptr;
  const bool  ended    = cudaStreamEndCapture(h->stream, &captured) == cudaSuccess && captured;
  if (ok && ended && cudaGraphInstantiate(&exec, captured, 0) != cudaSuccess) exec = nullptr;
  if (captured) cudaGraphDestroy(captured);
  if (side) cudaStreamDestroy(side);
  cudaGetLastError();
  return exec;
}

static int collide_local(mrsb_sim* h) {
  const int k = h->p2p ? h->parity : 0;
  if (h->lists_on) {
    // anything but exactly one stepping launch since the last pass: the displacement bound does not cover it
    if (h->positions_touched || h->steps_since_pass != 1)
      CU(cudaMemcpyAsync(&h->grid.ctl->force, h->h_one, sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    h->positions_touched = false;
    h->steps_since_pass  = 0;
    h->list_passes++;
    if (!h->coll_graph[k] && !h->list_graph_failed && !getenv("MRSB_NO_GRAPH")) {
      h->coll_graph[k]     = build_list_graph(h, &h->coll_graph_own[k], &h->rebuild_own);
      h->list_graph_failed = h->coll_graph[k] == nullptr;
    }
    if (h->coll_graph[k]) {
      CU(cudaGraphLaunch(h->coll_graph[k], h->stream));
      h->n_launches += h->coll_graph_own[k];  // the rebuilds are added from the device-side count (mrsb_get_counters)
    } else {
      // no graph (MRSB_NO_GRAPH, or conditional nodes unavailable): rebuild every pass
      h->n_launches += launch_collide_decide(h->grid, 1, cudaGraphConditionalHandle{}, 0, h->stream);
      h->rebuild_own = launch_collide_rebuild(h->ds, h->grid, h->cub_tmp, h->cub_tmp_bytes, h->stream);
      h->n_launches += launch_collide_check(h->ds, h->grid, h->coll_crash, h->coll_rebounce, h->stream);
    }
    h->n_passes++;
    CU(cudaGetLastError());
    return MRSB_OK;
  }
  if (!h->coll_graph[k] && !getenv("MRSB_NO_GRAPH")) {
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      const int own = launch_collide(h->ds, h->grid, h->coll_crash, h->coll_rebounce, h->cub_tmp, h->cub_tmp_bytes, h->stream);
      if (cudaStreamEndCapture(h->stream, &graph) == cudaSuccess && graph) {
        if (cudaGraphInstantiate(&h->coll_graph[k], graph, 0) != cudaSuccess) h->coll_graph[k] = nullptr;
        cudaGraphDestroy(graph);
        h->coll_graph_own[k] = own;
      }
    }
    cudaGetLastError();
  }
  if (h->coll_graph[k]) {
    CU(cudaGraphLaunch(h->coll_graph[k], h->stream));
    h->n_launches += h->coll_graph_own[k];
  } else {
    h->n_launches += launch_collide(h->ds, h->grid, h->coll_crash, h->coll_rebounce, h->cub_tmp, h->cub_tmp_bytes, h->stream);
  }
  h->n_passes++;
  CU(cudaGetLastError());
  return MRSB_OK;
}

int mrsb_make_step(mrsb_handle h, double dt, int32_t k_substeps) {
  GUARD(h);
  if (k_substeps < 1) return fail(MRSB_ERR_INVALID, "k_substeps must be >= 1");
  int rc = flush_params(h);
  if (rc) return rc;
  if (h->p2p) {
    if (*h->h_status) return fail(MRSB_ERR_STATE, "peer position exchange timed out (a rank stopped stepping)");
    h->parity ^= 1;
    h->ds.gpos  = h->gbuf[h->parity];
    h->ds.peers = h->d_peers[h->parity];
    h->epoch++;
    h->pushed = true;
  }
  h->n_launches += launch_step(h->ds, h->uniform_pset >= 0 ? &h->uniform_params : nullptr, dt, k_substeps, h->uniform_mode, h->uniform_nm,
                               h->any_moment, h->stream, h->step_info);
  h->n_steps += k_substeps;
  h->steps_since_pass++;
  CU(cudaGetLastError());
  return MRSB_OK;
}

int mrsb_handle_collisions(mrsb_handle h) {
  GUARD(h);
  if (!(h->coll_crash || h->coll_enabled)) return MRSB_OK;  // SIM:299-301
  int rc = flush_params(h);
  if (rc) return rc;
  if (h->ds.n_global > h->ds.n && h->n_ranks <= 1) return fail(MRSB_ERR_STATE, "sharded handle without communicator");
  if (h->lists_on && (h->positions_touched || h->steps_since_pass != 1))
    // anything but exactly one stepping launch since the last pass: this rank's displacement is unbounded —
    // said through the displacement word, so that every peer rebuilds as well
    CU(cudaMemcpyAsync(&h->grid.ctl->disp_max_bits, h->h_one + 1, sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
  rc = exchange_positions(h);
  if (rc) return rc;
  return collide_local(h);
}

int mrsb_handle_collisions_gathered(mrsb_handle h) {
  GUARD(h);
  if (!(h->coll_crash || h->coll_enabled)) return MRSB_OK;
  int rc = flush_params(h);
  if (rc) return rc;
  return collide_local(h);
}

int mrsb_run(mrsb_handle h, double dt, int32_t k_substeps, int32_t n_ticks, int32_t with_collisions) {
  GUARD(h);
  for (int t = 0; t < n_ticks; t++) {
    int rc = mrsb_make_step(h, dt, k_substeps);
    if (rc) return rc;
    if (with_collisions) {
      rc = mrsb_handle_collisions(h);
      if (rc) return rc;
    }
  }
  return MRSB_OK;
}

// ------------------------------------------------------------------------------------------
// state access
// ------------------------------------------------------------------------------------------
int mrsb_get_state(mrsb_handle h, int64_t n, const int32_t* idx, double* x, double* v, double* R, double* omega, double* motor_rpm) {
  GUARD(h);
  int rc = MRSB_OK;
  if (x && !rc) rc = get_rows(h, h->ds.st, ST_ROWS, 0, 3, n, idx, x);
  if (v && !rc) rc = get_rows(h, h->ds.st, ST_ROWS, 3, 3, n, idx, v);
  if (R && !rc) rc = get_rows(h, h->ds.st, ST_ROWS, 6, 9, n, idx, R);
  if (omega && !rc) rc = get_rows(h, h->ds.st, ST_ROWS, 15, 3, n, idx, omega);
  if (motor_rpm && !rc) rc = get_rows(h, h->ds.rpm, MRSB_NM, 0, MRSB_NM, n, idx, motor_rpm);
  return rc;
}

int mrsb_get_v_prev(mrsb_handle h, int64_t n, const int32_t* idx, double* v_prev) {
  GUARD(h);
  const int32_t* d_idx = nullptr;
  int            rc    = stage_idx(h, n, idx, &d_idx);
  if (rc) return rc;
  if (n == 0) return MRSB_OK;
  rc = ensure_stage(h, sizeof(double) * 3 * size_t(n));
  if (rc) return rc;
  h->n_launches += launch_gather_vprev(h->ds, n, d_idx, reinterpret_cast<double*>(h->d_stage), h->stream);
  CU(cudaMemcpyAsync(v_prev, h->d_stage, sizeof(double) * 3 * size_t(n), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return MRSB_OK;
}

int mrsb_get_imu_acceleration(mrsb_handle h, int64_t n, const int32_t* idx, double* acc) {
  GUARD(h);
  return get_rows(h, h->ds.imu, F3_ROWS, 0, 3, n, idx, acc);
}

int mrsb_set_state(mrsb_handle h, int64_t n, const int32_t* idx, const double* x, const double* v, const double* R, const double* omega,
                   const double* motor_rpm) {
  GUARD(h);
  int rc = MRSB_OK;
  if (v) {
    const int32_t* d_idx = nullptr;
    rc                   = stage_idx(h, n, idx, &d_idx);
    if (rc) return rc;
    h->n_launches += launch_stash_vprev(h->ds, n, d_idx, h->stream);
  }
  if (x && !rc) rc = put_rows(h, h->ds.st, ST_ROWS, 0, 3, n, idx, x, 3, 0);
  if (v && !rc) rc = put_rows(h, h->ds.st, ST_ROWS, 3, 3, n, idx, v, 3, 0);
  if (R && !rc) rc = put_rows(h, h->ds.st, ST_ROWS, 6, 9, n, idx, R, 9, 0);
  if (omega && !rc) rc = put_rows(h, h->ds.st, ST_ROWS, 15, 3, n, idx, omega, 3, 0);
  if (motor_rpm && !rc) rc = put_rows(h, h->ds.rpm, MRSB_NM, 0, MRSB_NM, n, idx, motor_rpm, MRSB_NM, 0);
  if (x && !rc) {
    h->n_launches += launch_publish_positions(h->ds, h->stream);
    h->pushed            = false;
    h->positions_touched = true;
  }
  return rc;
}

int mrsb_set_state_pos(mrsb_handle h, int64_t n, const int32_t* idx, const double* xyz, const double* heading) {
  GUARD(h);
  const int32_t* d_idx = nullptr;
  int            rc    = stage_idx(h, n, idx, &d_idx);
  if (rc) return rc;
  if (n == 0) return MRSB_OK;
  if (!xyz || !heading) return fail(MRSB_ERR_INVALID, "null payload");
  rc = ensure_stage(h, sizeof(double) * 4 * size_t(n));
  if (rc) return rc;
  double* d_xyz = reinterpret_cast<double*>(h->d_stage);
  double* d_hdg = d_xyz + 3 * n;
  CU(cudaMemcpyAsync(d_xyz, xyz, sizeof(double) * 3 * size_t(n), cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(d_hdg, heading, sizeof(double) * size_t(n), cudaMemcpyHostToDevice, h->stream));
  h->n_launches += launch_set_state_pos(h->ds, n, d_idx, d_xyz, d_hdg, h->stream);
  h->positions_touched = true;
  h->pushed = false;
  CU(cudaGetLastError());
  return MRSB_OK;
}

int mrsb_get_input_mode(mrsb_handle h, int64_t n, const int32_t* idx, int32_t* mode) {
  GUARD(h);
  const int32_t* d_idx = nullptr;
  int            rc    = stage_idx(h, n, idx, &d_idx);
  if (rc) return rc;
  if (n == 0) return MRSB_OK;
  rc = ensure_stage(h, sizeof(int32_t) * size_t(n));
  if (rc) return rc;
  h->n_launches += launch_gather_u8(h->ds.mode, n, d_idx, reinterpret_cast<int32_t*>(h->d_stage), h->stream);
  CU(cudaMemcpyAsync(mode, h->d_stage, sizeof(int32_t) * size_t(n), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return MRSB_OK;
}

static int get_flags(mrsb_sim* h, int64_t n, const int32_t* idx, std::vector<uint32_t>& out) {
  const int32_t* d_idx = nullptr;
  int            rc    = stage_idx(h, n, idx, &d_idx);
  if (rc) return rc;
  out.resize(size_t(n));
  if (n == 0) return MRSB_OK;
  rc = ensure_stage(h, sizeof(uint32_t) * size_t(n));
  if (rc) return rc;
  h->n_launches += launch_gather_u32(h->ds.flags, n, d_idx, reinterpret_cast<uint32_t*>(h->d_stage), h->stream);
  CU(cudaMemcpyAsync(out.data(), h->d_stage, sizeof(uint32_t) * size_t(n), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return MRSB_OK;
}

int mrsb_crash(mrsb_handle h, int64_t n, const int32_t* idx) {
  GUARD(h);
  const int32_t* d_idx = nullptr;
  int            rc    = stage_idx(h, n, idx, &d_idx);
  if (rc) return rc;
  h->n_launches += launch_flag_update(h->ds, n, d_idx, 0xffffffffu, FLAG_CRASHED, h->stream);
  return MRSB_OK;
}

int mrsb_has_crashed(mrsb_handle h, int64_t n, const int32_t* idx, int32_t* crashed) {
  GUARD(h);
  std::vector<uint32_t> fl;
  int                   rc = get_flags(h, n, idx, fl);
  if (rc) return rc;
  for (int64_t k = 0; k < n; k++) crashed[k] = (fl[size_t(k)] & FLAG_CRASHED) ? 1 : 0;
  return MRSB_OK;
}

// forces were written by something else than the collision pass: the next pass must replace every UAV's
// force (SIM:356-358), not only the ones it knows to be non-zero
static int forces_written(mrsb_sim* h) {
  if (!h->grid.ctl) return MRSB_OK;
  const unsigned long long until = (unsigned long long)h->list_passes + 1ull;  // index of the next pass that goes through decide_kernel
  CU(cudaMemcpyAsync(&h->grid.ctl->write_all_until, &until, sizeof(until), cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));  // `until` lives on this stack frame
  h->positions_touched = true;           // that pass rebuilds: UAVs without candidates are only visited by a rebuild
  return MRSB_OK;
}

int mrsb_forces_written(mrsb_handle h) {
  GUARD(h);
  return forces_written(h);
}

int mrsb_apply_force(mrsb_handle h, int64_t n, const int32_t* idx, const double* force) {
  GUARD(h);
  int rc = forces_written(h);
  if (rc) return rc;
  return put_rows(h, h->ds.fext, F3_ROWS, 0, 3, n, idx, force, 3, 0);
}
int mrsb_get_external_force(mrsb_handle h, int64_t n, const int32_t* idx, double* force) {
  GUARD(h);
  return get_rows(h, h->ds.fext, F3_ROWS, 0, 3, n, idx, force);
}
int mrsb_set_external_moment(mrsb_handle h, int64_t n, const int32_t* idx, const double* moment) {
  GUARD(h);
  h->any_moment = true;
  return put_rows(h, h->ds.mext, F3_ROWS, 0, 3, n, idx, moment, 3, 0);
}

// ------------------------------------------------------------------------------------------
// parameters
// ------------------------------------------------------------------------------------------
static int check_uav(mrsb_sim* h, int64_t uav) {
  if (uav < 0 || uav >= h->ds.n) return fail(MRSB_ERR_INVALID, "uav %lld outside 0..%lld", (long long)uav, (long long)h->ds.n - 1);
  return MRSB_OK;
}

int mrsb_get_params(mrsb_handle h, int64_t uav, mrsb_model_params* out) {
  GUARD(h);
  int rc = check_uav(h, uav);
  if (rc) return rc;
  *out = h->sets[h->pset_host[size_t(h->ds.shard_begin + uav)]].mp;
  std::vector<uint32_t> fl;
  const int32_t         one = int32_t(uav);
  rc                        = get_flags(h, 1, &one, fl);
  if (rc) return rc;
  out->takeoff_patch_enabled = (fl[0] & FLAG_TAKEOFF) ? 1 : 0;
  return MRSB_OK;
}

int mrsb_get_controller_params(mrsb_handle h, int64_t uav, mrsb_controller_params* out) {
  GUARD(h);
  int rc = check_uav(h, uav);
  if (rc) return rc;
  *out = h->sets[h->pset_host[size_t(h->ds.shard_begin + uav)]].cp;
  return MRSB_OK;
}

int mrsb_get_mixer_allocation(mrsb_handle h, int64_t uav, double* out) {
  GUARD(h);
  int rc = check_uav(h, uav);
  if (rc) return rc;
  double mix[MRSB_MAX_MOTORS][4];
  mrsb_mixer_allocation(h->sets[h->pset_host[size_t(h->ds.shard_begin + uav)]].mp, mix);
  std::memcpy(out, mix, sizeof(mix));
  return MRSB_OK;
}

int mrsb_set_params(mrsb_handle h, int64_t n, const int32_t* idx, const mrsb_model_params* params) {
  GUARD(h);
  if (!params) return fail(MRSB_ERR_INVALID, "null params");
  int rc = check_params(*params);
  if (rc) return rc;
  mrsb_controller_params def;
  mrsb_controller_params_default(&def);
  rc = repoint(h, n, idx, [&](ParamSet& s) {  // US:404-409: new model params, controllers re-created with default gains
    s.mp = *params;
    s.cp = def;
  });
  if (rc) return rc;
  const int32_t* d_idx = nullptr;
  rc                   = stage_idx(h, n, idx, &d_idx);
  if (rc) return rc;
  h->n_launches += launch_reset_pid(h->ds, n, d_idx, 0, PID_ROWS, h->stream);
  h->n_launches += launch_flag_update(h->ds, n, d_idx, ~FLAG_TAKEOFF, params->takeoff_patch_enabled ? FLAG_TAKEOFF : 0u, h->stream);
  return MRSB_OK;
}

static int set_ctrl(mrsb_sim* h, int64_t n, const int32_t* idx, int pid_row0, int pid_rows, void (*edit)(ParamSet&, const double*), const double* v) {
  int rc = repoint(h, n, idx, [&](ParamSet& s) { edit(s, v); });
  if (rc) return rc;
  if (pid_rows) {
    const int32_t* d_idx = nullptr;
    rc                   = stage_idx(h, n, idx, &d_idx);
    if (rc) return rc;
    h->n_launches += launch_reset_pid(h->ds, n, d_idx, pid_row0, pid_rows, h->stream);
  }
  return MRSB_OK;
}

int mrsb_set_mixer_params(mrsb_handle h, int64_t n, const int32_t* idx, int32_t desaturation) {
  GUARD(h);
  const double v[1] = {double(desaturation)};
  return set_ctrl(h, n, idx, 0, 0, [](ParamSet& s, const double* v) { s.cp.mixer_desaturation = v[0] != 0.0; }, v);
}
int mrsb_set_rate_controller_params(mrsb_handle h, int64_t n, const int32_t* idx, double kp, double kd, double ki) {
  GUARD(h);
  const double v[3] = {kp, kd, ki};
  return set_ctrl(h, n, idx, 18, 6, [](ParamSet& s, const double* v) { s.cp.rate_kp = v[0], s.cp.rate_kd = v[1], s.cp.rate_ki = v[2]; }, v);
}
int mrsb_set_attitude_controller_params(mrsb_handle h, int64_t n, const int32_t* idx, double kp, double kd, double ki, double max_rate_roll_pitch,
                                        double max_rate_yaw) {
  GUARD(h);
  const double v[5] = {kp, kd, ki, max_rate_roll_pitch, max_rate_yaw};
  return set_ctrl(h, n, idx, 12, 6, [](ParamSet& s, const double* v) {
    s.cp.att_kp = v[0], s.cp.att_kd = v[1], s.cp.att_ki = v[2], s.cp.att_max_rate_roll_pitch = v[3], s.cp.att_max_rate_yaw = v[4];
  }, v);
}
int mrsb_set_velocity_controller_params(mrsb_handle h, int64_t n, const int32_t* idx, double kp, double kd, double ki, double max_acceleration) {
  GUARD(h);
  const double v[4] = {kp, kd, ki, max_acceleration};
  return set_ctrl(h, n, idx, 6, 6, [](ParamSet& s, const double* v) {
    s.cp.vel_kp = v[0], s.cp.vel_kd = v[1], s.cp.vel_ki = v[2], s.cp.vel_max_acceleration = v[3];
  }, v);
}
int mrsb_set_position_controller_params(mrsb_handle h, int64_t n, const int32_t* idx, double kp, double kd, double ki, double max_velocity) {
  GUARD(h);
  const double v[4] = {kp, kd, ki, max_velocity};
  return set_ctrl(h, n, idx, 0, 6, [](ParamSet& s, const double* v) {
    s.cp.pos_kp = v[0], s.cp.pos_kd = v[1], s.cp.pos_ki = v[2], s.cp.pos_max_velocity = v[3];
  }, v);
}

// ------------------------------------------------------------------------------------------
// ROS-wrapper arithmetic around the path
// ------------------------------------------------------------------------------------------
int mrsb_timeout_input(mrsb_handle h, int64_t n, const int32_t* idx) {
  GUARD(h);
  const int32_t* d_idx = nullptr;
  int            rc    = stage_idx(h, n, idx, &d_idx);
  if (rc) return rc;
  h->n_launches += launch_timeout_input(h->ds, n, d_idx, h->stream);
  CU(cudaGetLastError());
  return MRSB_OK;
}

static int observe(mrsb_sim* h, int what, int width, int64_t n, const int32_t* idx, double* out) {
  if (!out) return fail(MRSB_ERR_INVALID, "null output");
  int rc = flush_params(h);
  if (rc) return rc;
  const int32_t* d_idx = nullptr;
  rc                   = stage_idx(h, n, idx, &d_idx);
  if (rc) return rc;
  if (n == 0) return MRSB_OK;
  const size_t bytes = sizeof(double) * size_t(n) * size_t(width);
  rc                 = ensure_stage(h, bytes);
  if (rc) return rc;
  h->n_launches += launch_observe(h->ds, what, n, d_idx, reinterpret_cast<double*>(h->d_stage), width, h->stream);
  CU(cudaMemcpyAsync(out, h->d_stage, bytes, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return MRSB_OK;
}
int mrsb_get_odometry(mrsb_handle h, int64_t n, const int32_t* idx, double* out13) {
  GUARD(h);
  return observe(h, 0, 13, n, idx, out13);
}
int mrsb_get_imu(mrsb_handle h, int64_t n, const int32_t* idx, double* out10) {
  GUARD(h);
  return observe(h, 1, 10, n, idx, out10);
}
int mrsb_get_rangefinder(mrsb_handle h, int64_t n, const int32_t* idx, double* out1) {
  GUARD(h);
  return observe(h, 2, 1, n, idx, out1);
}
int mrsb_pack_observations_device(mrsb_handle h, double* out_dev, int32_t stride) {
  GUARD(h);
  if (!out_dev || stride < 17) return fail(MRSB_ERR_INVALID, "need a device buffer with rows of >= 17 doubles");
  int rc = flush_params(h);
  if (rc) return rc;
  h->n_launches += launch_observe(h->ds, 3, h->ds.n, nullptr, out_dev, stride, h->stream);
  CU(cudaGetLastError());
  return MRSB_OK;
}

// set_mass / set_ground_z services: getParams -> edit -> setParams, per UAV (ROSW:1028-1080)
static int edit_params_each(mrsb_sim* h, int64_t n, const int32_t* idx, const std::function<void(mrsb_model_params&, int64_t)>& edit) {
  if (n < 0 || (!idx && n > h->ds.n)) return fail(MRSB_ERR_INVALID, "n=%lld outside 0..%lld", (long long)n, (long long)h->ds.n);
  std::vector<uint32_t> fl;
  int                   rc = get_flags(h, n, idx, fl);
  if (rc) return rc;
  for (int64_t k = 0; k < n; k++) {
    const int32_t     i = idx ? idx[k] : int32_t(k);
    mrsb_model_params p = h->sets[h->pset_host[size_t(h->ds.shard_begin + i)]].mp;
    p.takeoff_patch_enabled = (fl[size_t(k)] & FLAG_TAKEOFF) ? 1 : 0;  // getParams returns the live flag (MM:275)
    edit(p, k);
    rc = mrsb_set_params(h, 1, &i, &p);
    if (rc) return rc;
  }
  return MRSB_OK;
}

int mrsb_set_mass(mrsb_handle h, int64_t n, const int32_t* idx, const double* mass) {
  GUARD(h);
  if (!mass && n) return fail(MRSB_ERR_INVALID, "null mass");
  return edit_params_each(h, n, idx, [&](mrsb_model_params& p, int64_t k) {
    const double original = p.mass;
    p.mass                = mass[k];
    for (int m = 0; m < p.n_motors; m++)
      p.allocation_matrix[2 * MRSB_MAX_MOTORS + m] = p.mass * (p.allocation_matrix[2 * MRSB_MAX_MOTORS + m] / original);
    std::memset(p.J, 0, sizeof(p.J));
    p.J[0] = p.mass * (3.0 * p.arm_length * p.arm_length + p.body_height * p.body_height) / 12.0;
    p.J[4] = p.mass * (3.0 * p.arm_length * p.arm_length + p.body_height * p.body_height) / 12.0;
    p.J[8] = (p.mass * p.arm_length * p.arm_length) / 2.0;
  });
}

int mrsb_set_ground_z(mrsb_handle h, int64_t n, const int32_t* idx, const double* ground_z) {
  GUARD(h);
  if (!ground_z && n) return fail(MRSB_ERR_INVALID, "null ground_z");
  return edit_params_each(h, n, idx, [&](mrsb_model_params& p, int64_t k) { p.ground_z = ground_z[k]; });
}

// ------------------------------------------------------------------------------------------
// collisions
// ------------------------------------------------------------------------------------------
int mrsb_set_collisions(mrsb_handle h, int32_t enabled, int32_t crash, double rebounce) {
  GUARD(h);
  h->coll_enabled  = enabled != 0;
  h->coll_crash    = crash != 0;
  h->coll_rebounce = rebounce;
  drop_collision_graphs(h);
  return MRSB_OK;
}

int mrsb_get_collision_pairs(mrsb_handle h, int32_t* ij, int64_t cap, int64_t* count) {
  GUARD(h);
  unsigned long long found = 0;
  CU(cudaMemcpyAsync(&found, h->grid.counters, sizeof(found), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (count) *count = int64_t(found);
  if (int64_t(found) > h->grid.pair_cap)
    return fail(MRSB_ERR_CAPACITY, "%llu pairs found but the device pair buffer holds %lld", found, (long long)h->grid.pair_cap);
  if (!ij || found == 0) return MRSB_OK;
  std::vector<int32_t> tmp(2 * size_t(found));
  CU(cudaMemcpy(tmp.data(), h->grid.pairs, sizeof(int32_t) * tmp.size(), cudaMemcpyDeviceToHost));
  std::vector<std::pair<int32_t, int32_t>> pr(static_cast<size_t>(found), std::pair<int32_t, int32_t>(0, 0));
  for (size_t k = 0; k < pr.size(); k++) pr[k] = {tmp[2 * k], tmp[2 * k + 1]};
  std::sort(pr.begin(), pr.end());
  const int64_t m = std::min<int64_t>(cap, int64_t(found));
  for (int64_t k = 0; k < m; k++) {
    ij[2 * k]     = pr[size_t(k)].first;
    ij[2 * k + 1] = pr[size_t(k)].second;
  }
  if (int64_t(found) > cap) return fail(MRSB_ERR_CAPACITY, "%llu pairs found, caller buffer holds %lld", found, (long long)cap);
  return MRSB_OK;
}

int mrsb_set_pair_capacity(mrsb_handle h, int64_t max_pairs) {
  GUARD(h);
  if (max_pairs < 1) return fail(MRSB_ERR_INVALID, "max_pairs must be >= 1");
  CU(cudaStreamSynchronize(h->stream));
  int32_t* fresh = nullptr;
  CU(cudaMalloc(&fresh, sizeof(int32_t) * 2 * size_t(max_pairs)));
  if (h->grid.pairs) CU(cudaFree(h->grid.pairs));
  h->grid.pairs    = fresh;
  h->grid.pair_cap = max_pairs;
  drop_collision_graphs(h);
  CU(cudaMemsetAsync(h->grid.counters, 0, sizeof(unsigned long long), h->stream));
  return MRSB_OK;
}

int mrsb_get_counters(mrsb_handle h, int64_t* out5) {
  GUARD(h);
  unsigned long long found = 0;
  CU(cudaMemcpyAsync(&found, h->grid.counters, sizeof(found), cudaMemcpyDeviceToHost, h->stream));
  std::vector<uint32_t> fl;
  int                   rc = get_flags(h, h->ds.n, nullptr, fl);
  if (rc) return rc;
  int64_t crashed = 0;
  for (uint32_t f : fl) crashed += (f & FLAG_CRASHED) ? 1 : 0;
  if (h->lists_on) {
    // table rebuilds happen inside the graph's conditional node: count their kernels from the device-side tally
    NlCtl ctl;
    CU(cudaMemcpy(&ctl, h->grid.ctl, sizeof(ctl), cudaMemcpyDeviceToHost));
    h->n_launches += (int64_t(ctl.n_rebuilds) - h->rebuilds_counted) * h->rebuild_own;
    h->rebuilds_counted = int64_t(ctl.n_rebuilds);
  }
  out5[0] = h->n_steps;
  out5[1] = h->n_passes;
  out5[2] = int64_t(found);
  out5[3] = crashed;
  out5[4] = h->n_launches;
  return MRSB_OK;
}

int mrsb_get_step_info(mrsb_handle h, int32_t* out4) {
  GUARD(h);
  if (!out4) return fail(MRSB_ERR_INVALID, "null output");
  for (int k = 0; k < 4; k++) out4[k] = h->step_info[k];
  return MRSB_OK;
}

int mrsb_get_collision_info(mrsb_handle h, double* out8) {
  GUARD(h);
  CU(cudaStreamSynchronize(h->stream));
  NlCtl ctl{};
  if (h->lists_on) CU(cudaMemcpy(&ctl, h->grid.ctl, sizeof(ctl), cudaMemcpyDeviceToHost));
  out8[0] = 1.0 / h->grid.inv_cell;
  out8[1] = h->lists_on ? 1.0 : 0.0;
  out8[2] = h->lists_on ? std::sqrt(h->grid.list_r2) : 0.0;
  out8[3] = h->lists_on ? h->grid.skin : 0.0;
  out8[4] = double(ctl.n_passes);
  out8[5] = double(ctl.n_rebuilds);
  out8[6] = double(ctl.n_crowded);
  out8[7] = double(h->grid.n_buckets);
  return MRSB_OK;
}

// ------------------------------------------------------------------------------------------
// sharded operation
// ------------------------------------------------------------------------------------------
int mrsb_nccl_unique_id(void* out128) {
  int rc = load_nccl();
  if (rc) return rc;
  ncclUniqueId id;
  NC(g_nccl.GetUniqueId(&id));
  std::memcpy(out128, &id, sizeof(id));
  return MRSB_OK;
}

int mrsb_comm_init_nccl(mrsb_handle h, int32_t n_ranks, int32_t rank, const void* unique_id128) {
  GUARD(h);
  int rc = load_nccl();
  if (rc) return rc;
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(MRSB_ERR_INVALID, "bad rank %d of %d", rank, n_ranks);
  ncclUniqueId id;
  std::memcpy(&id, unique_id128, sizeof(id));
  NC(g_nccl.CommInitRank(&h->comm, n_ranks, id, rank));
  h->n_ranks = n_ranks;
  h->rank    = rank;
  // learn every rank's shard with one tiny all-gather
  int64_t* d_tab = nullptr;
  CU(cudaMalloc(&d_tab, sizeof(int64_t) * 2 * n_ranks));
  const int64_t mine[2] = {h->ds.shard_begin, h->ds.n};
  CU(cudaMemcpyAsync(d_tab + 2 * rank, mine, sizeof(mine), cudaMemcpyHostToDevice, h->stream));
  NC(g_nccl.AllGather(d_tab + 2 * rank, d_tab, 2, ncclInt64, h->comm, h->stream));
  std::vector<int64_t> tab(2 * size_t(n_ranks));
  CU(cudaMemcpyAsync(tab.data(), d_tab, sizeof(int64_t) * tab.size(), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaFree(d_tab));
  h->shard_begin_of.assign(size_t(n_ranks), 0);
  h->shard_count_of.assign(size_t(n_ranks), 0);
  h->equal_shards = true;
  int64_t covered = 0;
  for (int r = 0; r < n_ranks; r++) {
    h->shard_begin_of[size_t(r)] = tab[2 * size_t(r)];
    h->shard_count_of[size_t(r)] = tab[2 * size_t(r) + 1];
    covered += tab[2 * size_t(r) + 1];
    if (tab[2 * size_t(r) + 1] != h->ds.n || tab[2 * size_t(r)] != int64_t(r) * h->ds.n) h->equal_shards = false;
  }
  if (covered != h->ds.n_global) return fail(MRSB_ERR_INVALID, "shards cover %lld UAVs, n_global is %lld", (long long)covered, (long long)h->ds.n_global);
  h->ds.n_ranks = n_ranks;
  h->ds.rank    = rank;
  if (n_ranks > 1 && n_ranks <= 32 && !getenv("MRSB_NO_P2P")) {
    if (setup_p2p(h) != MRSB_OK) {  // no peer access (or IPC refused): the NCCL all-gather stays
      cudaGetLastError();
      h->p2p      = false;
      h->ds.peers = nullptr;
    }
  }
  return MRSB_OK;
}

int mrsb_exchange_mode(mrsb_handle h) {
  if (!h) return -1;
  return h->n_ranks <= 1 ? 0 : (h->p2p ? 2 : 1);
}

int mrsb_gather_buffer(mrsb_handle h, void** device_ptr, size_t* bytes) {
  GUARD(h);
  if (device_ptr) *device_ptr = h->ds.gpos;
  if (bytes) *bytes = sizeof(double) * 3 * size_t(h->ds.n_global);
  return MRSB_OK;
}

int mrsb_publish_positions(mrsb_handle h) {
  GUARD(h);
  h->pushed            = false;
  h->positions_touched = true;
  h->n_launches += launch_publish_positions(h->ds, h->stream);
  CU(cudaGetLastError());
  return MRSB_OK;
}

int mrsb_get_device_view(mrsb_handle h, mrsb_device_view* out) {
  GUARD(h);
  out->tile        = MRSB_TILE;
  out->state       = h->ds.st;
  out->state_rows  = ST_ROWS;
  out->motor_rpm   = h->ds.rpm;
  out->imu_acc     = h->ds.imu;
  out->ext_force   = h->ds.fext;
  out->flags       = h->ds.flags;
  out->input_mode  = h->ds.mode;
  return MRSB_OK;
}

}  // extern "C"
