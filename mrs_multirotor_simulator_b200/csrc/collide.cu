// collide.cu — K2/K3: MultirotorSimulator::handleCollisions (SIM:295-359) as a uniform-grid spatial
// hash over the packed positions of the WHOLE swarm (after the cross-shard all-gather), queried for
// this shard's UAVs only.
//
// The reference rebuilds a nanoflann KD-tree every tick and runs one radius query per UAV with
// squared "radius" 3.0 (SIM:309-328).  Here:
//   K2a  hash:     cell = floor(p / 2 m) (2 m > sqrt(3) so the 3x3x3 stencil is a superset of the
//                  search ball); bucket = (mix(cy,cz) + cx) mod B, B = 2^bits >= 2N.  Cells adjacent
//                  in x land in adjacent buckets, so one stencil row is ONE contiguous range.
//   K2b  sort:     CUB radix sort of (bucket, index) on `bits` bits (stable: equal buckets stay in
//                  ascending index order -> deterministic traversal).
//   K2c  ranges:   begin[b] = first sorted slot with bucket >= b (gap-filling scan of the sorted keys).
//   K2d  records:  rec[slot] = {x, y, z, index} gathered in sorted order (32-byte records: a
//                  candidate costs exactly one DRAM sector).
//   K3   collide:  one thread per sorted slot; 9 range probes; per candidate the EXACT reference
//                  predicate, evaluated with explicit round-to-nearest multiplies and adds (no FMA
//                  contraction) in nanoflann's order  d2 = ((dx*dx) + dy*dy) + dz*dz, dx = q - p
//                  (NF:479-484), accepted iff d2 < 3.0 (NF:305-309, strict) and j != i (SIM:335)
//                  and d2 < ((arm_i+prop_i)+arm_j)+prop_j (SIM:342,346: squared metres against
//                  metres — reproduced as is).
// Bucket aliasing (two cells sharing a bucket) only adds candidates; a candidate is accepted in the
// one probe whose (cy,cz) row matches its own cell, so nothing is counted twice.
//
// Crash mode (SIM:347-348) marks the NEIGHBOUR crashed.  A shard must not write remote state, so the
// owner of i evaluates the mirrored test d2 < ((arm_j+prop_j)+arm_i)+prop_i — the exact threshold
// the owner of j uses for the directed pair (j,i) — and marks i itself.  Rebounce mode (SIM:350)
// accumulates F_i in ascending j (the reference's KD-tree order is not reproducible without the
// tree; sums of <= 2 terms are order-independent bit for bit, longer ones agree to rounding).
#include <cub/device/device_radix_sort.cuh>

#include "internal.h"

namespace {

#define DEV __device__ __forceinline__

DEV int cell_of(double v) {
  // floor(v/2) saturated to +-2^29 (NaN -> 0); x*0.5 is exact
  double c = floor(v * 0.5);
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

__global__ void hash_kernel(const double* __restrict__ gpos, int64_t n, uint32_t mask, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const double* p  = gpos + 3 * j;
  const int     cx = cell_of(p[0]), cy = cell_of(p[1]), cz = cell_of(p[2]);
  keys[j]          = (row_hash(cy, cz) + uint32_t(cx)) & mask;
  vals[j]          = uint32_t(j);
}

// begin[b] = first slot whose key >= b ; begin[n_buckets] = n
__global__ void ranges_kernel(const uint32_t* __restrict__ keys_sorted, int64_t n, uint32_t n_buckets, uint32_t* __restrict__ begin) {
  const int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p > n) return;
  const uint32_t hi = (p == n) ? n_buckets : keys_sorted[p];
  const int64_t  lo = (p == 0) ? -1 : int64_t(keys_sorted[p - 1]);
  for (int64_t b = lo + 1; b <= int64_t(hi); b++) begin[b] = uint32_t(p);
}

__global__ void records_kernel(const double* __restrict__ gpos, const uint32_t* __restrict__ vals_sorted, int64_t n, double4* __restrict__ rec) {
  const int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const uint32_t j = vals_sorted[p];
  const double*  q = gpos + 3 * int64_t(j);
  rec[p]           = make_double4(q[0], q[1], q[2], __longlong_as_double((long long)j));
}

// nanoflann L2 metric for dim 3, no contraction
DEV double nf_dist2(double ax, double ay, double az, double bx, double by, double bz) {
  const double d0 = __dsub_rn(ax, bx), d1 = __dsub_rn(ay, by), d2 = __dsub_rn(az, bz);
  double       r  = __dmul_rn(d0, d0);
  r               = __dadd_rn(r, __dmul_rn(d1, d1));
  r               = __dadd_rn(r, __dmul_rn(d2, d2));
  return r;
}

struct Hit {
  int    count;
  double fx, fy, fz;
};

// Visit every neighbour j of UAV i (record q) that passes the reference predicate.
//   only_above: visit only j > after (ordered re-scan); returns the smallest such j in *next.
template <class F>
DEV void for_each_candidate(const DevGrid& g, const double4 q, int cx, int cy, int cz, F f) {
  const uint32_t mask = g.n_buckets - 1;
#pragma unroll 1
  for (int dz = -1; dz <= 1; dz++) {
#pragma unroll 1
    for (int dy = -1; dy <= 1; dy++) {
      const uint32_t b1 = (row_hash(cy + dy, cz + dz) + uint32_t(cx)) & mask;
      uint32_t       lo, hi;
      if (b1 >= 1 && b1 + 1 <= mask) {
        lo = g.begin[b1 - 1];
        hi = g.begin[b1 + 2];
        for (uint32_t p = lo; p < hi; p++) {
          const double4 r = g.rec[p];
          if (cell_of(r.y) == cy + dy && cell_of(r.z) == cz + dz && abs(cell_of(r.x) - cx) <= 1) f(r);
        }
      } else {  // stencil row wraps around the bucket table: probe the three buckets one by one
        for (int dx = -1; dx <= 1; dx++) {
          const uint32_t b = (b1 + uint32_t(dx)) & mask;
          lo               = g.begin[b];
          hi               = g.begin[b + 1];
          for (uint32_t p = lo; p < hi; p++) {
            const double4 r = g.rec[p];
            if (cell_of(r.y) == cy + dy && cell_of(r.z) == cz + dz && cell_of(r.x) == cx + dx) f(r);
          }
        }
      }
    }
  }
}

__global__ void __launch_bounds__(128) collide_kernel(DevState s, DevGrid g, int crash_mode, double rebounce) {
  const int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= s.n_global) return;
  const double4 q  = g.rec[p];
  const int64_t gi = __double_as_longlong(q.w);
  const int64_t li = gi - s.shard_begin;
  if (li < 0 || li >= s.n) return;  // not ours: its owner handles it

  const int cx = cell_of(q.x), cy = cell_of(q.y), cz = cell_of(q.z);
  const DevParams* __restrict__ Pi = s.params + s.pset[gi];
  const double ai = Pi->arm_length, pi_ = Pi->prop_radius, mi = Pi->mass;
  const double api = __dadd_rn(ai, pi_);

  int    hits = 0;
  bool   crashed_me = false;
  double fx = 0.0, fy = 0.0, fz = 0.0;

  auto contribution = [&](const double4& r, const DevParams* __restrict__ Pj, double& cx_, double& cy_, double& cz_) {
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
    cx_             = __dmul_rn(__dmul_rn(__dmul_rn(rebounce, nx), mi), wt);
    cy_             = __dmul_rn(__dmul_rn(__dmul_rn(rebounce, ny), mi), wt);
    cz_             = __dmul_rn(__dmul_rn(__dmul_rn(rebounce, nz), mi), wt);
  };

  for_each_candidate(g, q, cx, cy, cz, [&](const double4& r) {
    const int64_t gj = __double_as_longlong(r.w);
    if (gj == gi) return;  // SIM:335
    const double d2 = nf_dist2(q.x, q.y, q.z, r.x, r.y, r.z);
    if (!(d2 < 3.0)) return;  // NF:305-309
    const DevParams* __restrict__ Pj = s.params + s.pset[gj];
    const double aj = Pj->arm_length, pj = Pj->prop_radius;
    const double crit_ij = __dadd_rn(__dadd_rn(api, aj), pj);  // SIM:342
    if (d2 < crit_ij) {                                         // SIM:346
      hits++;
      const unsigned long long slot = atomicAdd(g.counters, 1ull);
      if (slot < (unsigned long long)g.pair_cap) {
        g.pairs[2 * slot]     = int32_t(gi);
        g.pairs[2 * slot + 1] = int32_t(gj);
      }
      if (!crash_mode) {
        double ax, ay, az;
        contribution(r, Pj, ax, ay, az);
        fx = __dadd_rn(fx, ax);
        fy = __dadd_rn(fy, ay);
        fz = __dadd_rn(fz, az);
      }
    }
    if (crash_mode) {
      const double crit_ji = __dadd_rn(__dadd_rn(__dadd_rn(aj, pj), ai), pi_);  // threshold of the directed pair (j,i)
      if (d2 < crit_ji) crashed_me = true;
    }
  });

  if (!crash_mode && hits >= 3) {
    // >= 3 simultaneous neighbours: redo the sum in ascending j so the result does not depend on
    // bucket order (selection by repeated scan; such clusters are rare and small)
    fx = fy = fz   = 0.0;
    int64_t last   = -1;
    for (int k = 0; k < hits; k++) {
      int64_t best = INT64_MAX;
      double  bx = 0, by = 0, bz = 0;
      for_each_candidate(g, q, cx, cy, cz, [&](const double4& r) {
        const int64_t gj = __double_as_longlong(r.w);
        if (gj == gi || gj <= last || gj >= best) return;
        const double d2 = nf_dist2(q.x, q.y, q.z, r.x, r.y, r.z);
        if (!(d2 < 3.0)) return;
        const DevParams* __restrict__ Pj = s.params + s.pset[gj];
        if (d2 < __dadd_rn(__dadd_rn(api, Pj->arm_length), Pj->prop_radius)) {
          best = gj;
          contribution(r, Pj, bx, by, bz);
        }
      });
      if (best == INT64_MAX) break;
      fx   = __dadd_rn(fx, bx);
      fy   = __dadd_rn(fy, by);
      fz   = __dadd_rn(fz, bz);
      last = best;
    }
  }

  // SIM:356-358: forces replace external_force_ for the next tick (zero in crash mode, SIM:315-319)
  s.fext[0 * s.ld + li] = fx;
  s.fext[1 * s.ld + li] = fy;
  s.fext[2 * s.ld + li] = fz;
  if (crashed_me) s.flags[li] |= FLAG_CRASHED;
}

}  // namespace

size_t collide_tmp_bytes(int64_t n_global) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  int(n_global), 0, 32);
  return bytes;
}

int launch_collide(const DevState& s, const DevGrid& g, int crash_mode, double rebounce, void* cub_tmp, size_t cub_tmp_bytes, cudaStream_t stream) {
  const int64_t n = s.n_global;
  if (n <= 0) return 0;
  const int      T  = 256;
  const unsigned nb = unsigned((n + T - 1) / T);
  cudaMemsetAsync(g.counters, 0, sizeof(unsigned long long), stream);
  hash_kernel<<<nb, T, 0, stream>>>(s.gpos, n, g.n_buckets - 1, g.keys, g.vals);
  cub::DeviceRadixSort::SortPairs(cub_tmp, cub_tmp_bytes, g.keys, g.keys_sorted, g.vals, g.vals_sorted, int(n), 0, int(g.bits), stream);
  ranges_kernel<<<unsigned((n + 1 + T - 1) / T), T, 0, stream>>>(g.keys_sorted, n, g.n_buckets, g.begin);
  records_kernel<<<nb, T, 0, stream>>>(s.gpos, g.vals_sorted, n, g.rec);
  collide_kernel<<<unsigned((n + 127) / 128), 128, 0, stream>>>(s, g, crash_mode, rebounce);
  return 6;  // hash + (>=1) sort + ranges + records + collide (+ memset); sort passes counted as one
}
