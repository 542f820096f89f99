// oracle/ref_nanoflann.cpp — the reference's collision pass on the REAL vendored nanoflann.
// TEST INFRASTRUCTURE ONLY.  Compiled by oracle/Makefile from the reference sources where they
// lie (-I$(REFERENCE)/include: nanoflann.hpp v1.5.0 and KDTreeVectorOfVectorsAdaptor.h are
// included, never copied) into oracle/_ref/libref_nanoflann.so, which is git-ignored and travels
// to the GPU box.  This file contributes only the loop around the library — the statement-for-
// statement shape of MultirotorSimulator::handleCollisions (src/multirotor_simulator.cpp:303-358)
// with Eigen::VectorXd replaced by std::vector<double> (the reference's UavSystem needs Eigen, absent here):
// same tree type (KDTreeVectorOfVectorsAdaptor<vector-of-vectors, double>, dim 3, leaf 10,
// SIM:309-311), same result set (RadiusResultSet<double,int>(3.0), SIM:326), same traversal.
#include <nanoflann.hpp>
#include <KDTreeVectorOfVectorsAdaptor.h>

#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>

typedef std::vector<std::vector<double>>                           points_t;  // one heap vector per point, like vector<Eigen::VectorXd> (SIM:26)
typedef KDTreeVectorOfVectorsAdaptor<points_t, double>             kd_tree_t;
typedef std::vector<nanoflann::ResultItem<int, double>>            results_t;

extern "C" int64_t ref_nanoflann_collide(int64_t n, const double* xyz, const double* arm, const double* prop, const double* mass, int32_t crash_mode,
                                         double rebounce, double* forces, uint8_t* crashed, int32_t* pairs, int64_t cap, int32_t n_threads) {
  if (n <= 0) return 0;  // the adaptor asserts on an empty set (KDA:75)
  points_t poses(n);
  for (int64_t i = 0; i < n; i++) poses[i] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};

  kd_tree_t mat_index(3, poses, 10);

  for (int64_t i = 0; i < 3 * n; i++) forces[i] = 0.0;

  const int                         nt    = n_threads > 1 ? n_threads : 1;
  const int64_t                     chunk = (n + nt - 1) / nt;
  std::vector<std::vector<int32_t>> found(nt);

  auto work = [&](int t) {
    results_t     indices_dists;
    const int64_t b = t * chunk, e = std::min<int64_t>(n, b + chunk);
    for (int64_t i = b; i < e; i++) {
      nanoflann::RadiusResultSet<double, int> resultSet(3.0, indices_dists);
      mat_index.index->findNeighbors(resultSet, poses[i].data());
      for (size_t k = 0; k < resultSet.m_indices_dists.size(); k++) {
        const size_t idx  = resultSet.m_indices_dists.at(k).first;
        const double dist = resultSet.m_indices_dists.at(k).second;
        if (idx == size_t(i)) continue;
        const double crit_dist = arm[i] + prop[i] + arm[idx] + prop[idx];
        if (dist < crit_dist) {
          found[t].push_back(int32_t(i));
          found[t].push_back(int32_t(idx));
          if (crash_mode) {
            __atomic_store_n(&crashed[idx], uint8_t(1), __ATOMIC_RELAXED);
          } else {
            const double rx = poses[i][0] - poses[idx][0], ry = poses[i][1] - poses[idx][1], rz = poses[i][2] - poses[idx][2];
            const double z  = rx * rx + (ry * ry + rz * rz);  // Eigen fixed-size squaredNorm
            double       nx = rx, ny = ry, nz = rz;
            if (z > 0.0) {
              const double s = std::sqrt(z);
              nx /= s;
              ny /= s;
              nz /= s;
            }
            const double w = mass[idx] / (mass[i] + mass[idx]);
            forces[3 * i + 0] += ((rebounce * nx) * mass[i]) * w;
            forces[3 * i + 1] += ((rebounce * ny) * mass[i]) * w;
            forces[3 * i + 2] += ((rebounce * nz) * mass[i]) * w;
          }
        }
      }
    }
  };
  if (nt == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back(work, t);
    for (auto& t : th) t.join();
  }

  int64_t count = 0;
  for (int t = 0; t < nt; t++)
    for (size_t k = 0; k + 1 < found[t].size(); k += 2) {
      if (pairs && count < cap) {
        pairs[2 * count]     = found[t][k];
        pairs[2 * count + 1] = found[t][k + 1];
      }
      count++;
    }
  return count;
}

// candidate search only (tree build + N radius queries), for timing the KD-tree share
extern "C" int64_t ref_nanoflann_count_neighbours(int64_t n, const double* xyz) {
  if (n <= 0) return 0;
  points_t poses(n);
  for (int64_t i = 0; i < n; i++) poses[i] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
  kd_tree_t mat_index(3, poses, 10);
  results_t indices_dists;
  int64_t   total = 0;
  for (int64_t i = 0; i < n; i++) {
    nanoflann::RadiusResultSet<double, int> resultSet(3.0, indices_dists);
    mat_index.index->findNeighbors(resultSet, poses[i].data());
    total += int64_t(indices_dists.size());
  }
  return total;
}
