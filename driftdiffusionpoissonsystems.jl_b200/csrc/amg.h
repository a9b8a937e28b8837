// Smoothed-aggregation multigrid preconditioner: host-side hierarchy structures
// shared by amg_host.cu (setup, no device code) and amg.cu (device kernels + the AMG-preconditioned CG).
#pragma once

#include <stdint.h>

#include <atomic>
#include <vector>

#include "../../include/ddps.h"

namespace ddps {

struct AmgParams {
  double theta = 0.25;      // strength threshold
  double omega = 2.0 / 3.0; // prolongator smoothing + Jacobi smoother damping
  int coarse_max = 128;     // stop coarsening at this size (dense inverse on the device: shared memory up to 128 rows)
  int max_levels = 12;
  int small_level = 512;    // levels with at most this many rows are coarsened gently (aggregate_small) ...
  int small_cap = 2;       // ... a root takes at most this many neighbours (0: off); 228 -> 77 rows instead of 32 at 16.8M triangles: -2..4 CG iterations per solve
  int nthreads = 0;         // 0: min(16, hardware concurrency)
};

struct AmgHostLevel {
  int n = 0;
  std::vector<int> rowptr, col;   // CSR pattern, sorted columns (level 0: the fixed P1 pattern)
  std::vector<double> val;        // BASE operator values at this level (strength + prolongator smoothing)
  int nc = 0;                     // rows of the next level (0 on the coarsest level)
  std::vector<int> agg;           // [n] aggregate id or -1
  std::vector<int> Pptr, Pcol;    // prolongator [n x nc], CSR
  std::vector<double> Pval;
  std::vector<int> Rptr, Rrow;    // its transpose [nc x n], CSR, fine indices increasing inside a row
  std::vector<double> Rval;
};

struct AmgHost {
  std::vector<AmgHostLevel> lv;
};

inline int amg_small_cap(const AmgParams& prm, int n) { return (n <= prm.small_level) ? prm.small_cap : 0; }
// aggregation on sorted strong-neighbour lists (shared by the host and the device setup paths); returns #aggregates
// rows_ready (optional): the lists of rows [0, *rows_ready) have arrived (a producer thread is still copying the rest from
// the device); the first, sequential pass consumes the rows as they come.
int amg_aggregate_lists(int n, const int* sptr, const int* scol, int cap, std::vector<int>& agg,
                        const std::atomic<long long>* rows_ready = nullptr);

int amg_host_build(AmgHost& H, int n, std::vector<int>&& rowptr, std::vector<int>&& col, std::vector<double>&& val,
                   const unsigned char* fixed, const AmgParams& prm);

}  // namespace ddps
