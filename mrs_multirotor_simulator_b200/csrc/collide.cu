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
// Crash mode (SIM:347-348) marks the NEIGHBOUR crashed.  A shard must not write remote state, so the
// owner of i evaluates the mirrored test d2 < ((arm_j+prop_j)+arm_i)+prop_i — the exact threshold
// the owner of j uses for the directed pair (j,i) — and marks i itself.
#include <cub/device/device_scan.cuh>

#include "internal.h"

namespace {

#define DEV __device__ __forceinline__

constexpr double kInvCell = 0.25;  // 4 m cells
constexpr double kReach   = 2.0;   // > sqrt(3.0), exactly representable

DEV int cell_of(double v) {
  // floor(v / 4 m) saturated to +-2^29 (NaN -> 0); v * 0.25 is exact
  double c = floor(v * kInvCell);
  c        = fmin(fmax(c, -536870912.0), 536870912.0);
  return (c == c) ? int(c) : 0;
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
                                                    const unsigned long long* __restrict__ aabb, uint32_t mask, uint32_t* __restrict__ count,
                                                    uint32_t* __restrict__ bucket, uint32_t* __restrict__ rank) {
  const int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const double* p = gpos + 3 * j;
  const double  x = p[0], y = p[1], z = p[2];
  if (filter) {
    const int64_t l = j - shard_begin;
    if (l < 0 || l >= n_local) {
      // remote: keep it only if it can reach this shard's box (NaN compares false -> kept)
      const bool out = x < dec(aabb[0]) - kReach || y < dec(aabb[1]) - kReach || z < dec(aabb[2]) - kReach || x > dec(aabb[3]) + kReach ||
                       y > dec(aabb[4]) + kReach || z > dec(aabb[5]) + kReach;
      if (out) {
        bucket[j] = 0xFFFFFFFFu;
        return;
      }
    }
  }
  const uint32_t b = (row_hash(cell_of(y), cell_of(z)) + uint32_t(cell_of(x))) & mask;
  bucket[j]        = b;
  rank[j]          = atomicAdd(&count[b], 1u);
  // bucket B = mask+1 mirrors bucket 0 (same records, same ranks), so that the two x-adjacent cells
  // of a stencil row are ALWAYS adjacent buckets, also across the end of the table
  if (b == 0u) atomicAdd(&count[mask + 1u], 1u);
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
  if (b == 0u) rec[begin[n_buckets] + k] = r;  // mirror of bucket 0
}

// nanoflann L2 metric for dim 3, no contraction
DEV double nf_dist2(double ax, double ay, double az, double bx, double by, double bz) {
  const double d0 = __dsub_rn(ax, bx), d1 = __dsub_rn(ay, by), d2 = __dsub_rn(az, bz);
  double       r  = __dmul_rn(d0, d0);
  r               = __dadd_rn(r, __dmul_rn(d1, d1));
  r               = __dadd_rn(r, __dmul_rn(d2, d2));
  return r;
}

#ifndef MRSB_COLLIDE_MINB
#define MRSB_COLLIDE_MINB 7  // 71 registers, no spills; 8 and 10 (64 / 48 registers) measured no faster
#endif
__global__ void __launch_bounds__(128, MRSB_COLLIDE_MINB) collide_kernel(DevState s, DevGrid g, int crash_mode, double rebounce) {
  const int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= int64_t(g.begin[g.n_buckets])) return;  // beyond the primary records (the mirror bucket holds copies)
  const double4 q  = g.rec[p];
  const int64_t gi = __double_as_longlong(q.w);
  const int64_t li = gi - s.shard_begin;
  if (li < 0 || li >= s.n) return;  // halo record: its owner handles it

  // the <= 4 stencil rows (cy0..cy1) x (cz0..cz1) around q; each is ONE record range cx0..cx1
  const uint32_t mask = g.n_buckets - 1;
  const int      cx0 = cell_of(q.x - kReach), cx1 = cell_of(q.x + kReach);
  const int      cy0 = cell_of(q.y - kReach), cy1 = cell_of(q.y + kReach);
  const int      cz0 = cell_of(q.z - kReach), cz1 = cell_of(q.z + kReach);
  const uint32_t w   = uint32_t(cx1 - cx0) + 1u;  // 1 or 2 buckets
  int            rcy[4], rcz[4];
  uint32_t       lo[4], hi[4];
  double4        first[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    rcy[k]        = (k & 1) ? cy1 : cy0;
    rcz[k]        = (k & 2) ? cz1 : cz0;
    const bool on = (!(k & 1) || cy1 != cy0) && (!(k & 2) || cz1 != cz0);
    lo[k] = hi[k] = 0u;
    if (on) {
      const uint32_t b0 = (row_hash(rcy[k], rcz[k]) + uint32_t(cx0)) & mask;
      lo[k]             = g.begin[b0];
      hi[k]             = g.begin[b0 + w];  // b0 + w <= mask + 2: begin[] has the mirror bucket and a sentinel
    }
  }
#pragma unroll
  for (int k = 0; k < 4; k++) first[k] = lo[k] < hi[k] ? g.rec[lo[k]] : q;  // the four leading candidates are fetched together

  // ---- phase 1 (cheap, unrolled): which records are in the search ball?  A record r found in row k
  // is a genuine, not-yet-seen neighbour iff d2 < 3.0 AND its own cell row is row k (rejects bucket
  // aliases and duplicates).  Almost every UAV has none; keep the first two, count the rest.
  auto in_ball = [&](const double4& r, int k) {
    if (__double_as_longlong(r.w) == gi) return false;  // SIM:335
    const double d2 = nf_dist2(q.x, q.y, q.z, r.x, r.y, r.z);
    if (!(d2 < 3.0)) return false;  // NF:305-309
    return cell_of(r.y) == rcy[k] && cell_of(r.z) == rcz[k];
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
    if (lo[k] < hi[k]) {
      if (in_ball(first[k], k)) note(lo[k]);
      for (uint32_t t = lo[k] + 1; t < hi[k]; t++)
        if (in_ball(g.rec[t], k)) note(t);
    }
  }

  double fx = 0.0, fy = 0.0, fz = 0.0;
  bool   crashed_me = false;
  if (n_ball > 0) {
    // ---- phase 2 (rare, one copy of the heavy code): thresholds, pair list, force, crash flag
    const DevParams* __restrict__ Pi = s.params + s.pset[gi];
    const double ai = Pi->arm_length, pi_ = Pi->prop_radius, mi = Pi->mass;
    const double api = __dadd_rn(ai, pi_);
    auto process = [&](const double4& r) {
      const int64_t gj = __double_as_longlong(r.w);
      const double  d2 = nf_dist2(q.x, q.y, q.z, r.x, r.y, r.z);
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
          const double rx = __dsub_rn(q.x, r.x), ry = __dsub_rn(q.y, r.y), rz = __dsub_rn(q.z, r.z);
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
          fx              = __dadd_rn(fx, __dmul_rn(__dmul_rn(__dmul_rn(rebounce, nx), mi), wt));
          fy              = __dadd_rn(fy, __dmul_rn(__dmul_rn(__dmul_rn(rebounce, ny), mi), wt));
          fz              = __dadd_rn(fz, __dmul_rn(__dmul_rn(__dmul_rn(rebounce, nz), mi), wt));
        }
      }
      if (crash_mode) {
        const double crit_ji = __dadd_rn(__dadd_rn(__dadd_rn(aj, pj), ai), pi_);  // threshold of the directed pair (j,i)
        if (d2 < crit_ji) crashed_me = true;
      }
    };
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
          for (uint32_t t = lo[k]; t < hi[k]; t++) {
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

  // SIM:356-358: forces replace external_force_ for the next tick (zero in crash mode, SIM:315-319)
  s.fext[tix(F3_ROWS, 0, li)] = fx;
  s.fext[tix(F3_ROWS, 1, li)] = fy;
  s.fext[tix(F3_ROWS, 2, li)] = fz;
  if (crashed_me) s.flags[li] |= FLAG_CRASHED;
}

}  // namespace

size_t collide_tmp_bytes(int64_t n_items) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, int(n_items));
  return bytes;
}

int launch_collide(const DevState& s, const DevGrid& g, int crash_mode, double rebounce, void* cub_tmp, size_t cub_tmp_bytes, cudaStream_t stream) {
  const int64_t n = s.n_global;
  if (n <= 0) return 0;
  const int      T      = 256;
  const unsigned nb     = unsigned((n + T - 1) / T);
  const int      filter = s.n_global > s.n;
  int            own    = 0;
  cudaMemsetAsync(g.counters, 0, sizeof(unsigned long long), stream);
  cudaMemsetAsync(g.count, 0, sizeof(uint32_t) * (size_t(g.n_buckets) + 2), stream);
  if (filter) {
    box_reset_kernel<<<1, 32, 0, stream>>>(g.aabb);
    if (s.n > 0) box_kernel<<<unsigned(std::min<int64_t>((s.n + T - 1) / T, 296)), T, 0, stream>>>(s.gpos, s.shard_begin, s.n, g.aabb);
    own += 2;
  }
  count_kernel<<<nb, T, 0, stream>>>(s.gpos, n, s.shard_begin, s.n, filter, g.aabb, g.n_buckets - 1, g.count, g.bucket, g.rank);
  cub::DeviceScan::ExclusiveSum(cub_tmp, cub_tmp_bytes, g.count, g.begin, int(g.n_buckets) + 2, stream);
  scatter_kernel<<<nb, T, 0, stream>>>(s.gpos, n, g.bucket, g.rank, g.begin, g.n_buckets, g.rec);
  collide_kernel<<<unsigned((n + 127) / 128), 128, 0, stream>>>(s, g, crash_mode, rebounce);
  return own + 3;  // own kernels: [box_reset, box,] count, scatter, collide (CUB's scan kernels and the memsets are not counted)
}
