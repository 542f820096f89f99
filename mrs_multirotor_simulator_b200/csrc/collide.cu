// collide.cu — K2/K3: MultirotorSimulator::handleCollisions (SIM:295-359) as a uniform-grid spatial
// hash over the packed positions of the swarm (after the cross-shard all-gather), queried for this
// shard's UAVs only.
//
// The reference rebuilds a nanoflann KD-tree every tick and runs one radius query per UAV with
// squared "radius" 3.0 (SIM:309-328).  Here, per pass:
//   K2a  box      (sharded runs only) bounding box of this shard's positions; a remote UAV further
//                 than 2 m (> sqrt 3) outside it cannot be a neighbour of any local UAV and is not
//                 inserted, so the table holds n_local + halo entries instead of n_global.
//   K2b  count    cell = floor(p / 4 m); bucket = (mix(cy,cz) + cx) mod B, B = 2^bits >= 2 n_global;
//                 rank = atomicAdd(count[bucket], 1).  Cells adjacent in x land in adjacent buckets,
//                 so one stencil row is ONE contiguous range of the grouped records.
//   K2c  scan     begin = exclusive prefix sum of count (CUB DeviceScan).
//   K2d  scatter  rec[begin[bucket] + rank] = {x, y, z, index}  (32-byte records: a candidate costs
//                 exactly one DRAM sector).  This is a counting sort: no radix passes.
//                 Bucket B mirrors bucket 0, so a row never straddles the end of the table.
//   K3   collide  one thread per record; the search ball of radius sqrt(3) < 2 m around p touches at
//                 most 2 cells per axis (interval [p-2, p+2] has the length of one 4 m cell), i.e.
//                 <= 4 stencil rows (cy0..cy1 x cz0..cz1), each one contiguous range cx0..cx1; per
//                 candidate the EXACT reference predicate, evaluated with explicit round-to-nearest
//                 multiplies and adds (no FMA contraction) in nanoflann's order
//                 d2 = ((dx*dx) + dy*dy) + dz*dz, dx = q - p (NF:479-484), accepted iff d2 < 3.0
//                 (NF:305-309, strict) and j != i (SIM:335) and
//                 d2 < ((arm_i+prop_i)+arm_j)+prop_j (SIM:342,346: squared metres against metres —
//                 reproduced as is).
// Bucket aliasing (two cells sharing a bucket) only adds candidates, which fail d2 < 3.0 unless they
// are true neighbours; a true neighbour is accepted only in the probe whose (cy,cz) row is its own
// cell row, so nothing is counted twice.  The arrival order inside a bucket is not deterministic;
// results are: pair lists are sorted on retrieval, and force sums of >= 3 terms are re-accumulated
// in ascending j (sums of <= 2 terms are order-independent bit for bit).
//
// Neighbour lists (single-shard handles).  A UAV moves centimetres per tick, so the table is not
// rebuilt every tick: a rebuild also records, for every UAV, the indices of all UAVs within
// R_list = sqrt(3) + skin of it (<= NL_CAP of them, slot-major [slot][uav]).  The stepping kernel
// reports the largest displacement of any UAV per launch (DevState::disp_max); `decide_kernel` sums
// these bounds into D and the lists stay valid while 2 D <= skin (two UAVs now closer than sqrt(3)
// were closer than sqrt(3) + 2 D when the lists were built).  On such ticks the pass is ONE kernel
// (`check_kernel`: the exact predicate on the current positions of the listed candidates); the
// rebuild (count/scan/scatter/build) sits in the body of a CUDA-graph conditional IF node whose
// condition `decide_kernel` sets on the device, so no host round trip is involved.  Results are the
// same bits either way: the lists only choose which pairs are tested.  Anything that moves UAVs
// other than one stepping launch (set_state, publish_positions, several launches between passes)
// forces a rebuild.  A UAV with more than NL_CAP candidates keeps no list: it remembers where its
// record sits in the (now ageing) table and, every pass, walks the stencil of its build-time cell —
// the records there are a superset of its possible neighbours for as long as the lists are valid —
// testing each candidate's CURRENT position (`check_crowded`).  Only the crowded UAVs pay for that.  Sharded handles use the lists when the fused exchange
// carries every rank's displacement bound (api.cu), the full pass otherwise.
// After a rebuild the UAVs that have anything to check (a candidate, or the crowded mark) are
// compacted, in index order, into `nl_active` (CUB DeviceSelect): a list-only pass runs over those
// only — typically a third of the swarm — with every lane busy.
// Bit 31 of a UAV's list-count word says "its external force may be non-zero": a pass must REPLACE every
// UAV's force (SIM:356-358), but writing 24 zero bytes over 24 zero bytes for the (vast) majority without a
// neighbour is most of a list-only pass — a UAV with an empty list and a clear bit is left alone.  Whoever
// writes forces elsewhere (mrsb_apply_force, mrsb_forces_written) makes the next pass write them all.
//
// Crash mode (SIM:347-348) marks the NEIGHBOUR crashed.  A shard must not write remote state, so the
// owner of i evaluates the mirrored test d2 < ((arm_j+prop_j)+arm_i)+prop_i — the exact threshold
// the owner of j uses for the directed pair (j,i) — and marks i itself.
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "internal.h"

namespace {

#define DEV __device__ __forceinline__

DEV int cell_of(double v, double inv_cell) {
  // floor(v / cell): one F2I.FLOOR — saturates to INT_MIN / INT_MAX, NaN -> 0.  Monotone in v, and the
  // SAME function files a record (count_kernel) and looks for it, which is all the stencil needs.
  return __double2int_rd(__dmul_rn(v, inv_cell));
}

DEV uint32_t row_hash(int cy, int cz) {
  uint32_t h = uint32_t(cy) * 0x9E3779B1u ^ (uint32_t(cz) * 0x85EBCA77u + 0x165667B1u);
  h ^= h >> 15;
  h *= 0x2C1B3C6Du;
  h ^= h >> 12;
  return h;
}

// order-preserving map double -> uint64 (for atomicMin / atomicMax)
DEV unsigned long long enc(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
DEV double dec(unsigned long long e) {
  const unsigned long long b = (e >> 63) ? (e & 0x7fffffffffffffffull) : ~e;
  return __longlong_as_double((long long)b);
}

__global__ void box_reset_kernel(unsigned long long* aabb) {
  if (threadIdx.x < 3) aabb[threadIdx.x] = ~0ull;
  if (threadIdx.x >= 3 && threadIdx.x < 6) aabb[threadIdx.x] = 0ull;
}

__global__ void __launch_bounds__(256) box_kernel(const double* __restrict__ gpos, int64_t begin, int64_t n, unsigned long long* aabb) {
  unsigned long long lo[3] = {~0ull, ~0ull, ~0ull}, hi[3] = {0ull, 0ull, 0ull};
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const double* p = gpos + 3 * (begin + i);
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const unsigned long long e = enc(p[c]);
      lo[c]                      = min(lo[c], e);
      hi[c]                      = max(hi[c], e);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; c++) {
    for (int o = 16; o > 0; o >>= 1) {
      lo[c] = min(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
      hi[c] = max(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&aabb[c], lo[c]);
      atomicMax(&aabb[3 + c], hi[c]);
    }
  }
}

__global__ void __launch_bounds__(256) count_kernel(const double* __restrict__ gpos, int64_t n, int64_t shard_begin, int64_t n_local, int filter,
                                                    const unsigned long long* __restrict__ aabb, uint32_t mask, double inv_cell, double reach,
                                                    uint32_t* __restrict__ count, uint32_t* __restrict__ bucket, uint32_t* __restrict__ rank) {
  const int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const double* p = gpos + 3 * j;
  const double  x = p[0], y = p[1], z = p[2];
  if (filter) {
    const int64_t l = j - shard_begin;
    if (l < 0 || l >= n_local) {
      // remote: keep it only if it can reach this shard's box (NaN compares false -> kept)
      const bool out = x < dec(aabb[0]) - reach || y < dec(aabb[1]) - reach || z < dec(aabb[2]) - reach || x > dec(aabb[3]) + reach ||
                       y > dec(aabb[4]) + reach || z > dec(aabb[5]) + reach;
      if (out) {
        bucket[j] = 0xFFFFFFFFu;
        return;
      }
    }
  }
  const uint32_t b = (row_hash(cell_of(y, inv_cell), cell_of(z, inv_cell)) + uint32_t(cell_of(x, inv_cell))) & mask;
  bucket[j]        = b;
  rank[j]          = atomicAdd(&count[b], 1u);
  // buckets B and B+1 (B = mask+1) mirror buckets 0 and 1 (same records, same ranks), so that the
  // x-adjacent cells of a stencil row are ALWAYS adjacent buckets, also across the end of the table
  if (b <= 1u) atomicAdd(&count[mask + 1u + b], 1u);
}

__global__ void __launch_bounds__(256) scatter_kernel(const double* __restrict__ gpos, int64_t n, const uint32_t* __restrict__ bucket,
                                                      const uint32_t* __restrict__ rank, const uint32_t* __restrict__ begin, uint32_t n_buckets,
                                                      double4* __restrict__ rec) {
  const int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const uint32_t b = bucket[j];
  if (b == 0xFFFFFFFFu) return;
  const double*  q = gpos + 3 * j;
  const double4  r = make_double4(q[0], q[1], q[2], __longlong_as_double((long long)j));
  const uint32_t k = rank[j];
  rec[begin[b] + k] = r;
  if (b <= 1u) rec[begin[n_buckets + b] + k] = r;  // mirrors of buckets 0 and 1
}

// nanoflann L2 metric for dim 3, no contraction
DEV double nf_dist2(double ax, double ay, double az, double bx, double by, double bz) {
  const double d0 = __dsub_rn(ax, bx), d1 = __dsub_rn(ay, by), d2 = __dsub_rn(az, bz);
  double       r  = __dmul_rn(d0, d0);
  r               = __dadd_rn(r, __dmul_rn(d1, d1));
  r               = __dadd_rn(r, __dmul_rn(d2, d2));
  return r;
}

// The <= 4 stencil rows around q.  The search ball (radius < reach = cell / 2) fits into two cells
// per axis: {c0, c0 + 1} with c0 = cell_of(q - reach) — for any p with |p - q| < r_search,
// q - reach <= p < (q - reach) + cell, and cell_of is monotone.  Each row (cy, cz) is ONE record
// range: buckets b0, b0 + 1 with b0 = bucket of (cx0, cy, cz).
struct Stencil {
  int      rcy[4], rcz[4];
  uint32_t lo[4], hi[4];
};
DEV Stencil stencil_of(const DevGrid& g, double qx, double qy, double qz) {
  Stencil        st;
  const uint32_t mask = g.n_buckets - 1;
  const int      cx0 = cell_of(qx - g.reach, g.inv_cell), cy0 = cell_of(qy - g.reach, g.inv_cell), cz0 = cell_of(qz - g.reach, g.inv_cell);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    st.rcy[k]         = cy0 + (k & 1);
    st.rcz[k]         = cz0 + ((k >> 1) & 1);
    const uint32_t b0 = (row_hash(st.rcy[k], st.rcz[k]) + uint32_t(cx0)) & mask;
    st.lo[k]          = g.begin[b0];
    st.hi[k]          = g.begin[b0 + 2];  // b0 + 2 <= mask + 2: begin[] has the two mirror buckets and a sentinel
  }
  return st;
}

// One accepted candidate (d2 < 3.0 already established): SIM:342-353 for the directed pair (i, j).
struct PairAcc {
  double fx = 0.0, fy = 0.0, fz = 0.0;
  bool   crashed_me = false;
};
DEV void process_pair(const DevState& s, const DevGrid& g, int crash_mode, double rebounce, int64_t gi, double qx, double qy, double qz,
                      const DevParams* __restrict__ Pi, int64_t gj, double rx_, double ry_, double rz_, PairAcc& acc) {
  const double ai = Pi->arm_length, pi_ = Pi->prop_radius, mi = Pi->mass;
  const double api = __dadd_rn(ai, pi_);
  const double d2  = nf_dist2(qx, qy, qz, rx_, ry_, rz_);
  const DevParams* __restrict__ Pj = s.params + s.pset[gj];
  const double aj = Pj->arm_length, pj = Pj->prop_radius;
  const double crit_ij = __dadd_rn(__dadd_rn(api, aj), pj);  // SIM:342
  if (d2 < crit_ij) {                                         // SIM:346
    const unsigned long long slot = atomicAdd(g.counters, 1ull);
    if (slot < (unsigned long long)g.pair_cap) {
      g.pairs[2 * slot]     = int32_t(gi);
      g.pairs[2 * slot + 1] = int32_t(gj);
    }
    if (!crash_mode) {
      // rebounce * normalized(x_i - x_j) * m_i * (m_j / (m_i + m_j))   (SIM:350), Eigen evaluation order
      const double rx = __dsub_rn(qx, rx_), ry = __dsub_rn(qy, ry_), rz = __dsub_rn(qz, rz_);
      const double z  = __dadd_rn(__dmul_rn(rx, rx), __dadd_rn(__dmul_rn(ry, ry), __dmul_rn(rz, rz)));
      double       nx = rx, ny = ry, nz = rz;
      if (z > 0.0) {
        const double sq = __dsqrt_rn(z);
        nx              = __ddiv_rn(rx, sq);
        ny              = __ddiv_rn(ry, sq);
        nz              = __ddiv_rn(rz, sq);
      }
      const double mj = Pj->mass;
      const double wt = __ddiv_rn(mj, __dadd_rn(mi, mj));
      acc.fx          = __dadd_rn(acc.fx, __dmul_rn(__dmul_rn(__dmul_rn(rebounce, nx), mi), wt));
      acc.fy          = __dadd_rn(acc.fy, __dmul_rn(__dmul_rn(__dmul_rn(rebounce, ny), mi), wt));
      acc.fz          = __dadd_rn(acc.fz, __dmul_rn(__dmul_rn(__dmul_rn(rebounce, nz), mi), wt));
    }
  }
  if (crash_mode) {
    const double crit_ji = __dadd_rn(__dadd_rn(__dadd_rn(aj, pj), ai), pi_);  // threshold of the directed pair (j,i)
    if (d2 < crit_ji) acc.crashed_me = true;
  }
}

// SIM:356-358: forces replace external_force_ for the next tick (zero in crash mode, SIM:315-319)
#define NL_LIVE 0x80000000u     // bit 31 of nl_count[]: the UAV's external force may be non-zero
#define NL_CROWDED 0x40000000u  // bit 30: more than MRSB_NL_CAP candidates; nl_items[0][uav] = its record in the table
DEV void store_result(const DevState& s, int64_t li, const PairAcc& acc) {
  s.fext[tix(F3_ROWS, 0, li)] = acc.fx;
  s.fext[tix(F3_ROWS, 1, li)] = acc.fy;
  s.fext[tix(F3_ROWS, 2, li)] = acc.fz;
  if (acc.crashed_me) s.flags[li] |= FLAG_CRASHED;
}
DEV bool nonzero(const PairAcc& acc) {
  return !(acc.fx == 0.0 && acc.fy == 0.0 && acc.fz == 0.0);  // NaN counts as non-zero
}

#ifndef MRSB_COLLIDE_MINB
#define MRSB_COLLIDE_MINB 7  // 71 registers, no spills; 8 and 10 (64 / 48 registers) measured no faster
#endif
// The full pass: every record against its stencil (handles without neighbour lists).
__global__ void __launch_bounds__(128, MRSB_COLLIDE_MINB) collide_kernel(DevState s, DevGrid g, int crash_mode, double rebounce) {
  const int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= int64_t(g.begin[g.n_buckets])) return;  // beyond the primary records (the mirror buckets hold copies)
  const double4 q  = g.rec[p];
  const int64_t gi = __double_as_longlong(q.w);
  const int64_t li = gi - s.shard_begin;
  if (li < 0 || li >= s.n) return;  // halo record: its owner handles it

  const Stencil st = stencil_of(g, q.x, q.y, q.z);
  double4       first[4];
#pragma unroll
  for (int k = 0; k < 4; k++) first[k] = st.lo[k] < st.hi[k] ? g.rec[st.lo[k]] : q;  // the four leading candidates are fetched together

  // ---- phase 1 (cheap, unrolled): which records are in the search ball?  A record r found in row k
  // is a genuine, not-yet-seen neighbour iff d2 < 3.0 AND its own cell row is row k (rejects bucket
  // aliases and duplicates).  Almost every UAV has none; keep the first two, count the rest.
  auto in_ball = [&](const double4& r, int k) {
    if (__double_as_longlong(r.w) == gi) return false;  // SIM:335
    const double d2 = nf_dist2(q.x, q.y, q.z, r.x, r.y, r.z);
    if (!(d2 < 3.0)) return false;  // NF:305-309
    return cell_of(r.y, g.inv_cell) == st.rcy[k] && cell_of(r.z, g.inv_cell) == st.rcz[k];
  };
  int      n_ball = 0;
  uint32_t h0 = 0, h1 = 0;
  auto     note = [&](uint32_t t) {
    if (n_ball == 0) h0 = t;
    if (n_ball == 1) h1 = t;
    n_ball++;
  };
#pragma unroll
  for (int k = 0; k < 4; k++) {
    if (st.lo[k] < st.hi[k]) {
      if (in_ball(first[k], k)) note(st.lo[k]);
      for (uint32_t t = st.lo[k] + 1; t < st.hi[k]; t++)
        if (in_ball(g.rec[t], k)) note(t);
    }
  }

  PairAcc acc;
  if (n_ball > 0) {
    // ---- phase 2 (rare, one copy of the heavy code): thresholds, pair list, force, crash flag
    const DevParams* __restrict__ Pi = s.params + s.pset[gi];
    auto process = [&](const double4& r) { process_pair(s, g, crash_mode, rebounce, gi, q.x, q.y, q.z, Pi, __double_as_longlong(r.w), r.x, r.y, r.z, acc); };
    if (n_ball <= 2) {
      // sums of <= 2 terms do not depend on the order
      for (int c = 0; c < n_ball; c++) {
        process(g.rec[h0]);
        h0 = h1;
      }
    } else {
      // >= 3 neighbours in the ball: visit them in ascending j so that the force sum does not
      // depend on the arrival order inside the buckets (selection by repeated scan; rare)
      int64_t last = -1;
      for (int c = 0; c < n_ball; c++) {
        int64_t  best = INT64_MAX;
        uint32_t bt   = 0;
        for (int k = 0; k < 4; k++) {
          for (uint32_t t = st.lo[k]; t < st.hi[k]; t++) {
            const double4 r  = g.rec[t];
            const int64_t gj = __double_as_longlong(r.w);
            if (gj <= last || gj >= best || !in_ball(r, k)) continue;
            best = gj;
            bt   = t;
          }
        }
        if (best == INT64_MAX) break;
        process(g.rec[bt]);
        last = best;
      }
    }
  }
  store_result(s, li, acc);
  if (g.nl_count) g.nl_count[li] = (g.nl_count[li] & ~NL_LIVE) | (nonzero(acc) ? NL_LIVE : 0u);
}

// ---- neighbour lists ---------------------------------------------------------------------------

// One thread per record: every other record within the list radius goes into the UAV's list.
__global__ void __launch_bounds__(128, MRSB_COLLIDE_MINB) build_lists_kernel(DevState s, DevGrid g) {
  const int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= int64_t(g.begin[g.n_buckets])) return;
  const double4 q  = g.rec[p];
  const int64_t gi = __double_as_longlong(q.w);
  const int64_t li = gi - s.shard_begin;
  if (li < 0 || li >= s.n) return;
  const Stencil st = stencil_of(g, q.x, q.y, q.z);
  double4       first[4];
#pragma unroll
  for (int k = 0; k < 4; k++) first[k] = st.lo[k] < st.hi[k] ? g.rec[st.lo[k]] : q;
  uint32_t cnt  = 0;
  auto     take = [&](const double4& r, int k) {
    const int64_t gj = __double_as_longlong(r.w);
    if (gj == gi) return;
    const double d2 = nf_dist2(q.x, q.y, q.z, r.x, r.y, r.z);
    if (!(d2 < g.list_r2)) return;
    if (cell_of(r.y, g.inv_cell) != st.rcy[k] || cell_of(r.z, g.inv_cell) != st.rcz[k]) return;  // bucket alias / duplicate
    if (cnt < MRSB_NL_CAP) g.nl_items[int64_t(cnt) * g.nl_ld + li] = int32_t(gj);
    cnt++;
  };
#pragma unroll
  for (int k = 0; k < 4; k++) {
    if (st.lo[k] < st.hi[k]) {
      take(first[k], k);
      for (uint32_t t = st.lo[k] + 1; t < st.hi[k]; t++) take(g.rec[t], k);
    }
  }
  uint32_t live = g.nl_count[li] & NL_LIVE;
  if (cnt == 0u && (live || g.ctl->n_passes <= g.ctl->write_all_until)) {
    // nobody within the list radius: this UAV is not visited again until the next rebuild, so its force
    // (left from an earlier collision, or written from outside) is replaced by zero right here (SIM:356-358)
    s.fext[tix(F3_ROWS, 0, li)] = 0.0;
    s.fext[tix(F3_ROWS, 1, li)] = 0.0;
    s.fext[tix(F3_ROWS, 2, li)] = 0.0;
    live                        = 0u;
  }
  if (cnt > MRSB_NL_CAP) {
    // too crowded for a list: remember the record instead (check_crowded walks its stencil every pass)
    g.nl_items[li] = int32_t(uint32_t(p));
    g.nl_count[li] = NL_CROWDED | live;
    atomicAdd(&g.ctl->n_crowded, 1u);
  } else {
    g.nl_count[li] = cnt | live;
  }
}

// A UAV without a list: every record of the stencil around its BUILD-TIME position, tested at the
// candidates' CURRENT positions (the table only says who they are).  Same structure as collide_kernel.
DEV void check_crowded(const DevState& s, const DevGrid& g, int crash_mode, double rebounce, int64_t gi, uint32_t rec_index, PairAcc& acc) {
  const double4 qb = g.rec[rec_index];
  const Stencil st = stencil_of(g, qb.x, qb.y, qb.z);
  const double* qp = s.gpos + 3 * gi;
  const double  qx = qp[0], qy = qp[1], qz = qp[2];
  auto in_ball = [&](const double4& r, int k) {  // r: the candidate's record (build-time position: decides the row it was filed under)
    const int64_t gj = __double_as_longlong(r.w);
    if (gj == gi) return false;
    if (cell_of(r.y, g.inv_cell) != st.rcy[k] || cell_of(r.z, g.inv_cell) != st.rcz[k]) return false;  // bucket alias / duplicate
    const double* rp = s.gpos + 3 * gj;
    return nf_dist2(qx, qy, qz, rp[0], rp[1], rp[2]) < 3.0;
  };
  const DevParams* __restrict__ Pi = s.params + s.pset[gi];
  auto process = [&](int64_t gj) {
    const double* rp = s.gpos + 3 * gj;
    process_pair(s, g, crash_mode, rebounce, gi, qx, qy, qz, Pi, gj, rp[0], rp[1], rp[2], acc);
  };
  int     n_ball = 0;
  int64_t h0 = 0, h1 = 0;
  for (int k = 0; k < 4; k++)
    for (uint32_t t = st.lo[k]; t < st.hi[k]; t++) {
      const double4 r = g.rec[t];
      if (!in_ball(r, k)) continue;
      if (n_ball == 0) h0 = __double_as_longlong(r.w);
      if (n_ball == 1) h1 = __double_as_longlong(r.w);
      n_ball++;
    }
  if (n_ball <= 2) {
    if (n_ball >= 1) process(h0);
    if (n_ball == 2) process(h1);
    return;
  }
  int64_t last = -1;  // >= 3: ascending j
  for (int c = 0; c < n_ball; c++) {
    int64_t best = INT64_MAX;
    for (int k = 0; k < 4; k++)
      for (uint32_t t = st.lo[k]; t < st.hi[k]; t++) {
        const double4 r  = g.rec[t];
        const int64_t gj = __double_as_longlong(r.w);
        if (gj <= last || gj >= best || !in_ball(r, k)) continue;
        best = gj;
      }
    if (best == INT64_MAX) break;
    process(best);
    last = best;
  }
}

// One UAV that has something to check: the exact predicate on the CURRENT positions of its listed candidates.
// (Tried and dropped: remembering each candidate's distance at build time and skipping the fetch while
// d_build - 2 D is still above sqrt(3) — the extra dependent load cost more than the skipped gathers.)
DEV void check_one(const DevState& s, const DevGrid& g, int crash_mode, double rebounce, int64_t li) {
  const uint32_t word = g.nl_count[li];
  const uint32_t cnt  = word & ~(NL_LIVE | NL_CROWDED);
  // forces were written from outside since the last pass (this pass' index <= write_all_until): replace them all
  const bool     live = (word & NL_LIVE) || g.ctl->n_passes <= g.ctl->write_all_until;
  PairAcc acc;
  if (word & NL_CROWDED) {
    check_crowded(s, g, crash_mode, rebounce, li + s.shard_begin, uint32_t(g.nl_items[li]), acc);
  } else if (cnt) {
    const int64_t gi = li + s.shard_begin;
    const double* qp = s.gpos + 3 * gi;
    const double  qx = qp[0], qy = qp[1], qz = qp[2];
    uint32_t      hits = 0;
    for (uint32_t base = 0; base < cnt; base += 4) {
      // four candidates at a time: indices, then positions, then distances (independent loads in flight)
      int32_t gj[4];
      double  r[4][3];
#pragma unroll
      for (int u = 0; u < 4; u++) gj[u] = base + u < cnt ? g.nl_items[int64_t(base + u) * g.nl_ld + li] : -1;
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const double* rp = s.gpos + 3 * int64_t(max(gj[u], 0));
        r[u][0] = rp[0], r[u][1] = rp[1], r[u][2] = rp[2];
      }
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (gj[u] >= 0 && nf_dist2(qx, qy, qz, r[u][0], r[u][1], r[u][2]) < 3.0) hits |= 1u << (base + u);  // NF:305-309
    }
    if (hits) {
      const DevParams* __restrict__ Pi = s.params + s.pset[gi];
      auto process = [&](uint32_t c) {
        const int64_t gj = g.nl_items[int64_t(c) * g.nl_ld + li];
        const double* rp = s.gpos + 3 * gj;
        process_pair(s, g, crash_mode, rebounce, gi, qx, qy, qz, Pi, gj, rp[0], rp[1], rp[2], acc);
      };
      if (__popc(hits) <= 2) {
        // sums of <= 2 terms do not depend on the order
        while (hits) {
          process(uint32_t(__ffs(int(hits)) - 1));
          hits &= hits - 1;
        }
      } else {
        // >= 3: ascending j, like collide_kernel
        int64_t last = -1;
        while (hits) {
          int64_t  best = INT64_MAX;
          uint32_t bc   = 0;
          for (uint32_t m = hits; m; m &= m - 1) {
            const uint32_t c  = uint32_t(__ffs(int(m)) - 1);
            const int64_t  gj = g.nl_items[int64_t(c) * g.nl_ld + li];
            if (gj > last && gj < best) best = gj, bc = c;
          }
          if (best == INT64_MAX) break;
          process(bc);
          last = best;
          hits &= ~(1u << bc);
        }
      }
    }
  }
  const bool nz = nonzero(acc);
  if (live || nz) store_result(s, li, acc);
  else if (acc.crashed_me) s.flags[li] |= FLAG_CRASHED;
  if (nz != bool(word & NL_LIVE)) g.nl_count[li] = (word & ~NL_LIVE) | (nz ? NL_LIVE : 0u);
}


// A list-only pass: the compacted UAVs that have something to check, grid-stride.
__global__ void __launch_bounds__(256, 4) check_kernel(DevState s, DevGrid g, int crash_mode, double rebounce) {
  const uint32_t n_active = g.ctl->n_active;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n_active; k += gridDim.x * blockDim.x) check_one(s, g, crash_mode, rebounce, g.nl_active[k]);
}

struct HasWork {
  const uint32_t* nl_count;
  __device__ bool operator()(int32_t li) const { return (nl_count[li] & ~NL_LIVE) != 0u; }
};

// Are the lists still good for the positions of this pass?  One thread.
__global__ void decide_kernel(NlCtl* __restrict__ c, unsigned long long* __restrict__ pair_counter, double skin, int always,
                              cudaGraphConditionalHandle handle, int has_handle) {
  // every load first (they are independent: one round trip), then the decision, then the stores
  const uint32_t           bits  = c->disp_max_bits;  // largest squared displacement of the stepping launch since the last pass (float, rounded up)
  const double             D_old = c->D_total;
  const uint32_t           force = c->force, valid = c->valid;
  const unsigned long long n_rebuilds = c->n_rebuilds, n_passes = c->n_passes;
  const double d       = __dsqrt_ru(double(__uint_as_float(bits)));  // NaN stays NaN
  double       D       = __dadd_ru(D_old, d);
  const bool   rebuild = always || force || !valid || !(__dmul_ru(2.0, D) <= skin);
  if (has_handle) cudaGraphSetConditional(handle, rebuild ? 1u : 0u);
  *pair_counter    = 0ull;  // pairs found by this pass
  c->disp_max_bits = 0u;
  if (rebuild) {
    D            = 0.0;
    c->force     = 0u;
    c->valid     = 1u;
    c->n_crowded = 0u;  // build_lists_kernel counts them again
    c->n_active  = 0u;  // ... and the compaction after it
    c->n_rebuilds = n_rebuilds + 1ull;
  }
  c->D_total  = D;
  c->rebuild  = rebuild ? 1u : 0u;
  c->n_passes = n_passes + 1ull;
}

}  // namespace

size_t collide_tmp_bytes(int64_t n_items, int64_t n_local) {
  size_t scan = 0, select = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan, (const uint32_t*)nullptr, (uint32_t*)nullptr, int(n_items));
  cub::DeviceSelect::If(nullptr, select, thrust::counting_iterator<int32_t>(0), (int32_t*)nullptr, (uint32_t*)nullptr, int(std::max<int64_t>(n_local, 1)),
                        HasWork{nullptr});
  return std::max(scan, select);
}

// table build: [box_reset, box,] count, scan, scatter.  Returns the number of own kernels.
static int launch_table(const DevState& s, const DevGrid& g, void* cub_tmp, size_t cub_tmp_bytes, cudaStream_t stream) {
  const int64_t  n      = s.n_global;
  const int      T      = 256;
  const unsigned nb     = unsigned((n + T - 1) / T);
  const int      filter = s.n_global > s.n;
  int            own    = 0;
  cudaMemsetAsync(g.count, 0, sizeof(uint32_t) * (size_t(g.n_buckets) + 3), stream);
  if (filter) {
    box_reset_kernel<<<1, 32, 0, stream>>>(g.aabb);
    if (s.n > 0) box_kernel<<<unsigned(std::min<int64_t>((s.n + T - 1) / T, 296)), T, 0, stream>>>(s.gpos, s.shard_begin, s.n, g.aabb);
    own += 2;
  }
  count_kernel<<<nb, T, 0, stream>>>(s.gpos, n, s.shard_begin, s.n, filter, g.aabb, g.n_buckets - 1, g.inv_cell, g.reach, g.count, g.bucket, g.rank);
  cub::DeviceScan::ExclusiveSum(cub_tmp, cub_tmp_bytes, g.count, g.begin, int(g.n_buckets) + 3, stream);
  scatter_kernel<<<nb, T, 0, stream>>>(s.gpos, n, g.bucket, g.rank, g.begin, g.n_buckets, g.rec);
  return own + 2;
}

// The full pass of every tick (sharded handles, or MRSB_NO_NEIGHBOUR_LISTS): table + collide_kernel.
int launch_collide(const DevState& s, const DevGrid& g, int crash_mode, double rebounce, void* cub_tmp, size_t cub_tmp_bytes, cudaStream_t stream) {
  const int64_t n = s.n_global;
  if (n <= 0) return 0;
  cudaMemsetAsync(g.counters, 0, sizeof(unsigned long long), stream);
  const int own = launch_table(s, g, cub_tmp, cub_tmp_bytes, stream);
  collide_kernel<<<unsigned((n + 127) / 128), 128, 0, stream>>>(s, g, crash_mode, rebounce);
  return own + 1;  // CUB's scan kernels and the memsets are not counted
}

// ---- the pass with neighbour lists, in three pieces so that api.cu can put the middle one into the
// body of a conditional graph node -----------------------------------------------------------------
int launch_collide_decide(const DevGrid& g, int always, cudaGraphConditionalHandle handle, int has_handle, cudaStream_t stream) {
  decide_kernel<<<1, 1, 0, stream>>>(g.ctl, g.counters, g.skin, always, handle, has_handle);
  return 1;
}
int launch_collide_rebuild(const DevState& s, const DevGrid& g, void* cub_tmp, size_t cub_tmp_bytes, cudaStream_t stream) {
  const int64_t n = s.n_global;
  if (n <= 0) return 0;
  const int own = launch_table(s, g, cub_tmp, cub_tmp_bytes, stream);
  build_lists_kernel<<<unsigned((n + 127) / 128), 128, 0, stream>>>(s, g);
  // the UAVs with something to check, in index order
  cub::DeviceSelect::If(cub_tmp, cub_tmp_bytes, thrust::counting_iterator<int32_t>(0), g.nl_active, &g.ctl->n_active, int(s.n), HasWork{g.nl_count}, stream);
  return own + 1;
}
int launch_collide_check(const DevState& s, const DevGrid& g, int crash_mode, double rebounce, cudaStream_t stream) {
  if (s.n <= 0) return 0;
  check_kernel<<<unsigned(std::min<int64_t>((s.n + 255) / 256, 148 * 4)), 256, 0, stream>>>(s, g, crash_mode, rebounce);
  return 1;
}
