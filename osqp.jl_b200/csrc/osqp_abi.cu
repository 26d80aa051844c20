// osqp_abi.cu -- the C ABI of include/osqp.h on top of the CUDA engine (kernels.cu).
//
// Host-side responsibilities only: validate, marshal the caller's Int64 CSC into the int32 CSR
// layouts the kernels stream, keep the five host-visible Workspace fields Julia dereferences
// (src/interface.jl:176-205) in pinned memory, turn CUDA errors into non-zero exit codes.  No
// arithmetic of the solver runs on the host; without a usable GPU every entry point fails loudly.
#include "engine.cuh"
#include "osqp.h"
#include "osqp_b200.h"
#ifdef OSQP_B200_DEVTOOLS
#include "osqp_b200_dev.h"
#endif

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

using namespace osqpb200;

namespace {

constexpr double kRhoMin = 1e-6, kRhoMax = 1e6;

double now_s() {
  using namespace std::chrono;
  return duration<double>(steady_clock::now().time_since_epoch()).count();
}

#define CU_OK(expr)                                                                              \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      fprintf(stderr, "ERROR in %s: CUDA failure '%s' at %s:%d (%s)\n", __func__,                \
              cudaGetErrorString(_e), __FILE__, __LINE__, #expr);                                \
      return 100 + (c_int)_e;                                                                    \
    }                                                                                            \
  } while (0)

struct Engine {
  OSQPWorkspace pub;  // MUST be first: the ABI pointer is &pub
  OSQPData data_pub;
  OSQPSettings st;
  OSQPInfo info;
  OSQPSolution sol_pub;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
  DevPtrs d;
  LaunchGeom geom;
  std::vector<void *> dev_allocs, host_allocs;
  // pinned host mirrors
  double *h_sol_x = nullptr, *h_sol_y = nullptr, *h_dx = nullptr, *h_dy = nullptr;
  DevInfo *h_info = nullptr;
  PolishOut *h_pol = nullptr;
  DevState *h_state = nullptr;
  PolishOut *d_pol = nullptr;
  // host copies of the bounds as passed (l <= u validation on updates)
  std::vector<double> l0, u0;
  // csc position -> device position maps
  int *mapA = nullptr, *mapP1 = nullptr, *mapP2 = nullptr;
  long long nnzA = 0, nnzPtriu = 0;
  // staging for value / index uploads
  double *stage_val = nullptr;
  long long *stage_idx = nullptr;
  long long stage_cap = 0;
  // engine controls
  double pcg_eta = 1e-3, pcg_floor = 1e-13, pcg_eta_e = -1.0;  // pcg_eta_e < 0: automatic (kernels.cu admm_kernel)
  int pcg_max_iter = 0, refresh_every = 25;
  double polish_penalty = 1e4;
  bool first_run = true, clear_update_time = false;
  // ---- tiny mode (n, m <= kTinyMax): the same QP also lives in the batched engine (batch.cu, a batch of one: dense /
  // sparse KKT in shared memory, exact Cholesky solve, ~1 us per ADMM iteration instead of the ~60 us of a
  // grid-synchronised PCG step on a problem with a handful of non-zeros).  osqp_solve uses it whenever the solve needs
  // nothing the batched engine lacks (polish, time limit, verbose log); every update is applied to both engines.
  OSQPB200Batch *tiny = nullptr;
  std::vector<c_int> tP_p, tP_i, tA_p, tA_i;  // host copies of the problem as passed (re-setup after P / A / rho updates)
  std::vector<double> tP_x, tA_x, tq;
  int last_path = 0;  // 0: none yet, 1: general engine, 2: tiny engine -- iterates are handed over on a switch
  bool wood_dirty = true;  // WoodDev data must be rebuilt by the next launch (setup, re-scaling, rho / bound updates, polish)
  OSQPB200Profile prof;
};

Engine *E(OSQPWorkspace *w) { return reinterpret_cast<Engine *>(w); }
const Engine *E(const OSQPWorkspace *w) { return reinterpret_cast<const Engine *>(w); }

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

template <typename T>
cudaError_t dalloc(Engine &e, T **p, size_t count) {
  *p = nullptr;
  if (count == 0) count = 1;
  void *q = nullptr;
  cudaError_t err = cudaMalloc(&q, count * sizeof(T));
  if (err != cudaSuccess) return err;
  e.dev_allocs.push_back(q);
  *p = reinterpret_cast<T *>(q);
  return cudaMemsetAsync(q, 0, count * sizeof(T), e.stream);
}
template <typename T>
cudaError_t halloc(Engine &e, T **p, size_t count) {
  *p = nullptr;
  if (count == 0) count = 1;
  void *q = nullptr;
  cudaError_t err = cudaMallocHost(&q, count * sizeof(T));
  if (err != cudaSuccess) return err;
  e.host_allocs.push_back(q);
  memset(q, 0, count * sizeof(T));
  *p = reinterpret_cast<T *>(q);
  return cudaSuccess;
}

void destroy(Engine *e) {
  if (!e) return;
  DeviceGuard guard(e->device);
  if (e->tiny) osqp_batch_cleanup(e->tiny);
  e->tiny = nullptr;
  if (e->stream) cudaStreamSynchronize(e->stream);
  for (void *p : e->dev_allocs) cudaFree(p);
  for (void *p : e->host_allocs) cudaFreeHost(p);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  if (e->ev2) cudaEventDestroy(e->ev2);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

void update_status(OSQPInfo &info, c_int v) {
  info.status_val = v;
  const char *s = "unsolved";
  switch (v) {
    case OSQP_SOLVED: s = "solved"; break;
    case OSQP_SOLVED_INACCURATE: s = "solved inaccurate"; break;
    case OSQP_PRIMAL_INFEASIBLE: s = "primal infeasible"; break;
    case OSQP_PRIMAL_INFEASIBLE_INACCURATE: s = "primal infeasible inaccurate"; break;
    case OSQP_DUAL_INFEASIBLE: s = "dual infeasible"; break;
    case OSQP_DUAL_INFEASIBLE_INACCURATE: s = "dual infeasible inaccurate"; break;
    case OSQP_MAX_ITER_REACHED: s = "maximum iterations reached"; break;
    case OSQP_TIME_LIMIT_REACHED: s = "run time limit reached"; break;
    case OSQP_SIGINT: s = "interrupted"; break;
    case OSQP_NON_CVX: s = "problem non convex"; break;
    default: break;
  }
  memset(info.status, 0, sizeof(info.status));
  strncpy(info.status, s, sizeof(info.status) - 1);
}

void reset_info(Engine &e) {
  e.info.solve_time = 0;
  e.info.polish_time = 0;
  update_status(e.info, OSQP_UNSOLVED);
  e.info.rho_updates = 0;
  e.h_state->rho_updates = 0;
}

void begin_update(Engine &e) {
  if (e.clear_update_time) {
    e.clear_update_time = false;
    e.info.update_time = 0.0;
  }
}

int validate_data(const OSQPData *d) {
  if (!d) { fprintf(stderr, "ERROR in osqp_setup: missing data\n"); return 1; }
  if (!d->P || !d->A || (!d->q && d->n > 0)) { fprintf(stderr, "ERROR in osqp_setup: missing matrix/vector\n"); return 1; }
  if (d->n <= 0 || d->m < 0) { fprintf(stderr, "ERROR in osqp_setup: n must be positive and m nonnegative\n"); return 1; }
  if (d->P->m != d->n || d->P->n != d->n) { fprintf(stderr, "ERROR in osqp_setup: P does not have dimension n x n\n"); return 1; }
  for (c_int j = 0; j < d->n; j++)
    for (c_int k = d->P->p[j]; k < d->P->p[j + 1]; k++)
      if (d->P->i[k] > j) { fprintf(stderr, "ERROR in osqp_setup: P is not upper triangular\n"); return 1; }
  if (d->A->m != d->m || d->A->n != d->n) { fprintf(stderr, "ERROR in osqp_setup: A does not have dimension m x n\n"); return 1; }
  for (c_int i = 0; i < d->m; i++)
    if (d->l[i] > d->u[i]) {
      fprintf(stderr, "ERROR in osqp_setup: lower bound at index %lld is greater than upper bound\n", (long long)i);
      return 1;
    }
  const long long lim = 2147483647LL - 64;
  if (d->n > lim || d->m > lim || d->A->p[d->n] > lim || 2 * d->P->p[d->n] > lim) {
    fprintf(stderr, "ERROR in osqp_setup: problem exceeds the engine's int32 index range\n");
    return 1;
  }
  return 0;
}
int validate_settings(const OSQPSettings *s) {
  if (!s) return 1;
  bool bad = s->scaling < 0 || (s->adaptive_rho != 0 && s->adaptive_rho != 1) || s->adaptive_rho_interval < 0 ||
             s->adaptive_rho_fraction <= 0 || s->adaptive_rho_tolerance < 1.0 || s->polish_refine_iter < 0 ||
             s->rho <= 0 || s->sigma <= 0 || s->delta <= 0 || s->max_iter <= 0 || s->eps_abs < 0 ||
             s->eps_rel < 0 || (s->eps_rel == 0 && s->eps_abs == 0) || s->eps_prim_inf <= 0 ||
             s->eps_dual_inf <= 0 || s->alpha <= 0 || s->alpha >= 2 ||
             (s->linsys_solver != QDLDL_SOLVER && s->linsys_solver != MKL_PARDISO_SOLVER) ||
             (s->verbose != 0 && s->verbose != 1) || (s->scaled_termination != 0 && s->scaled_termination != 1) ||
             s->check_termination < 0 || (s->warm_start != 0 && s->warm_start != 1) || s->time_limit < 0;
  if (bad) fprintf(stderr, "ERROR in osqp_setup: invalid settings\n");
  return bad ? 1 : 0;
}

double env_double(const char *name, double dflt) {
  const char *s = getenv(name);
  return (s && *s) ? atof(s) : dflt;
}
int env_int(const char *name, int dflt) {
  const char *s = getenv(name);
  return (s && *s) ? atoi(s) : dflt;
}

// Host-side setup work (index transposition, tile-stream construction) is split over a few threads: fn(lo, hi) on
// contiguous ranges of [0, count).  Every use below writes disjoint outputs per index, so the result does not
// depend on the number of threads.
int host_threads() {
  const int forced = env_int("OSQP_B200_HOST_THREADS", 0);
  if (forced > 0) return forced;
  const unsigned hw = std::thread::hardware_concurrency();
  return (int)std::max(1u, std::min(16u, hw));
}
template <typename F>
void par_ranges(long long count, long long min_per_thread, F fn) {
  int T = host_threads();
  if (count / std::max<long long>(1, min_per_thread) < T) T = (int)std::max<long long>(1, count / std::max<long long>(1, min_per_thread));
  if (T <= 1) { fn(0LL, count); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < T; t++) {
    const long long lo = count * t / T, hi = count * (t + 1) / T;
    th.emplace_back([=]() { fn(lo, hi); });
  }
  for (std::thread &x : th) x.join();
}

int pow2_lanes(double avg_nnz_per_row) {
  int forced = env_int("OSQP_B200_LANES", 0);
  if (forced > 0) return forced;
  int l = 1;
  while (l < 32 && (double)l * 4.0 < avg_nnz_per_row) l <<= 1;
  return l;
}

// contiguous, weight-balanced split of `rows` rows into `parts` ranges
void balanced_split(const std::vector<long long> &w_prefix, int rows, int parts, std::vector<int> &start) {
  start.assign(parts + 1, 0);
  const long long total = w_prefix[rows];
  int r = 0;
  for (int b = 1; b < parts; b++) {
    const long long target = (long long)((double)total * (double)b / (double)parts);
    while (r < rows && w_prefix[r] < target) r++;
    start[b] = r;
  }
  start[parts] = rows;
}

// Host-side construction of a tile stream (engine.cuh TileStreamDev) from row-stacked CSR matrices that share the
// column space.  `mats[k]` = (rowptr, col, rows) of the k-th stacked matrix; CSR positions are stacked likewise.
struct CsrRef {
  const std::vector<int> *rowptr, *col;
  int rows;
};
// Values are stored chunk-interleaved: entry t of quad (lane) l of a chunk sits at (t / 2) * 64 + 2 * l + t % 2, so a
// warp reads a chunk's values as two contiguous 512 B rows (16 B per lane each); the column words stay in entry order.
inline int stream_val_pos(int p) {
  const int chunk = p >> 7, l = (p >> 2) & 31, t = p & 3;
  return (chunk << 7) + ((t >> 1) << 6) + 2 * l + (t & 1);
}

struct TileStreamHost {
  int rows = 0, cols = 0, ngroups = 0;
  long long nelem = 0, nnz = 0;
  std::vector<unsigned short> cf;
  std::vector<int> from_csr, blk_group, blk_row0, blk_row1, grp_col0, w_row0, w_q0, w_qn, sr_ptr;
  std::vector<int> w_s0, sl_len, sl_row;  // lane-row layout (engine.cuh TileStreamDev): slices of every warp
  int max_slice = 0, max_block_rows = 0, paired = 0, split = 0, srows = 0, lane_rows = 0;
};

// false: the stream does not pay off for this matrix (too much padding) -> CSR path
// paired: thread-block clusters of two (DSMEM combine of the two column-group partials); needs ngroups == 2.
bool build_tile_stream(const std::vector<CsrRef> &mats, int cols, int grid, int ngroups, bool paired,
                       TileStreamHost &T, int max_pair_rows = 1 << 30, bool lane_rows = true) {
  if (paired && (ngroups != 2 || (grid & 1))) paired = false;
  T.sl_len.clear(); T.sl_row.clear(); T.w_s0.clear();
  const bool dbg = getenv("OSQP_B200_DEBUG") != nullptr;
  double t_mark = now_s();
  auto mark = [&](const char *what) {
    if (dbg) {
      const double t = now_s();
      fprintf(stderr, "[osqp_b200]   stream build: %-24s %7.1f ms\n", what, (t - t_mark) * 1e3);
      t_mark = t;
    }
  };
  int rows = 0;
  long long nnz = 0;
  for (const CsrRef &M : mats) { rows += M.rows; nnz += (*M.rowptr)[M.rows]; }
  T.rows = rows; T.cols = cols; T.ngroups = ngroups; T.nnz = nnz;
  if (ngroups > grid || ngroups < 1) return false;
  // column groups (boundaries are multiples of 32 so every TMA source stays 16 B aligned)
  const int Wg = (((cols + ngroups - 1) / ngroups) + 31) & ~31;
  if (Wg > kSliceMax || Wg > 32768) return false;
  T.max_slice = Wg;
  T.grp_col0.resize(ngroups + 1);
  for (int g = 0; g <= ngroups; g++) T.grp_col0[g] = std::min(cols, g * Wg);
  // entries per (row, group)
  std::vector<int> cnt((size_t)rows * ngroups, 0);
  {
    int r0 = 0;
    for (const CsrRef &M : mats) {
      par_ranges(M.rows, 8192, [&, r0](long long ra, long long rb) {
        for (long long r = ra; r < rb; r++)
          for (int k = (*M.rowptr)[r]; k < (*M.rowptr)[r + 1]; k++) cnt[(size_t)(r0 + r) * ngroups + (*M.col)[k] / Wg]++;
      });
      r0 += M.rows;
    }
  }
  mark("count per (row, group)");
  // quads of a row inside a group (a row without entries there still costs one zero quad)
  auto quads_of = [](int c) { return std::max(1, (c + 3) >> 2); };
  // Stream rows.  A row with more than `split_quads` quads in a group is cut into pieces of that many quads, each
  // piece a row of its own for the stream (own row-end flag, own partial sum): a warp owns whole stream rows, so
  // without the cut one dense row would serialise on a single warp.  sr_ptr[g][r] = first stream row of row r.
  // Pieces are only as small as balance needs: a quarter of the mean load of a warp, at least kSplitQuads.
  T.sr_ptr.assign((size_t)ngroups * (rows + 1), 0);
  T.split = 0;
  T.srows = rows;
  long long all_quads = 0;
  for (size_t i = 0; i < cnt.size(); i++) all_quads += quads_of(cnt[i]);
  // lane rows: a slice (32 stream rows) is as long as its longest row and goes to ONE warp, so pieces must be short
  // enough that every warp gets several slices: at most 1/128 of a warp's mean load (at least 16 quads).  A row that
  // would fall into more than 32 such pieces (a few very dense rows next to many short ones: portfolio) makes the
  // owners' sums over the pieces the bottleneck instead -- measured 13 % slower than the scan layout on config 4 --
  // so such matrices keep the scan layout.
  if (lane_rows) {
    const long long sq = std::max<long long>(16, all_quads / ((long long)grid * kWarps) / 128);
    int maxq = 0;
    for (size_t i = 0; i < cnt.size(); i++) maxq = std::max(maxq, quads_of(cnt[i]));
    if ((long long)maxq > 32 * sq) lane_rows = false;
  }
  const int split_quads = lane_rows ? (int)std::max<long long>(16, all_quads / ((long long)grid * kWarps) / 128)
                                    : (int)std::max<long long>(kSplitQuads, all_quads / ((long long)grid * kWarps) / 4);
  for (int g = 0; g < ngroups; g++) {
    int *sp = T.sr_ptr.data() + (size_t)g * (rows + 1);
    for (int r = 0; r < rows; r++) sp[r + 1] = sp[r] + (quads_of(cnt[(size_t)r * ngroups + g]) + split_quads - 1) / split_quads;
    if (sp[rows] != rows) T.split = 1;
    T.srows = std::max(T.srows, sp[rows]);
  }
  if (T.split) paired = false;  // the DSMEM combine of a pair works row by row
  // weight prefix per group over its stream rows
  std::vector<std::vector<long long>> wpre(ngroups);
  std::vector<long long> gw(ngroups, 0);
  long long stored = 0;
  for (int g = 0; g < ngroups; g++) {
    const int *sp = T.sr_ptr.data() + (size_t)g * (rows + 1);
    std::vector<long long> &wp = wpre[g];
    wp.assign((size_t)sp[rows] + 1, 0);
    for (int r = 0; r < rows; r++) {
      int q = quads_of(cnt[(size_t)r * ngroups + g]);
      for (int s = sp[r]; s < sp[r + 1]; s++) {
        const int piece = std::min(q, split_quads);
        wp[s + 1] = wp[s] + piece;
        q -= piece;
      }
    }
    gw[g] = wp[sp[rows]];
    stored += 4 * gw[g];
  }
  if (nnz > 0 && (double)stored > 1.35 * (double)nnz + 4096.0) return false;  // padding would dominate
  if (stored > 2147483000LL) return false;
  mark("stream rows, weights");
  // thread blocks per group, proportional to the stored quads (every group gets at least one)
  std::vector<int> nblk(ngroups, 1);
  {
    const int left = grid - ngroups;
    std::vector<double> frac(ngroups);
    int used = 0;
    for (int g = 0; g < ngroups; g++) {
      const double want = (double)left * (double)gw[g] / (double)std::max<long long>(1, stored / 4);
      nblk[g] += (int)want;
      used += (int)want;
      frac[g] = want - (int)want;
    }
    for (int k = used; k < left; k++) {
      int best = 0;
      for (int g = 1; g < ngroups; g++) if (frac[g] > frac[best]) best = g;
      nblk[best]++;
      frac[best] = -1.0;
    }
  }
  T.blk_group.assign(grid, 0);
  T.blk_row0.assign(grid, 0);
  T.blk_row1.assign(grid, 0);
  T.w_row0.assign((size_t)grid * kWarps, 0);
  T.w_q0.assign((size_t)grid * kWarps + 1, 0);
  // contiguous split of (stream) rows [ra, rb) into `parts` ranges balanced on a weight prefix (every range that
  // still has rows gets at least one)
  auto split = [&](const long long *wp, int ra, int rb, int parts, std::vector<int> &cut, int max_rows = 1 << 30) {
    cut.assign(parts + 1, rb);
    int r = ra;
    for (int k = 0; k < parts; k++) {
      cut[k] = r;
      const int left_k = parts - k;
      const long long target = wp[r] + (wp[rb] - wp[r] + left_k - 1) / left_k;
      // rows this range must take so that the remaining ranges can still hold the rest under the cap
      const long long must = (long long)(rb - r) - (long long)(left_k - 1) * max_rows;
      int r1 = r;
      while (r1 < rb && r1 - r < max_rows && (r1 == r || wp[r1 + 1] <= target || r1 - r < must)) r1++;
      r = (k == parts - 1) ? rb : r1;
    }
    cut[parts] = rb;
  };
  T.paired = paired ? 1 : 0;
  std::vector<int> cut;
  if (paired) {
    // cluster pairs: blocks 2p (group 0) and 2p+1 (group 1) stream the SAME row range p; ranges are balanced on the
    // combined weight so both halves of a pair finish together  (no split rows here: stream rows == rows)
    std::vector<long long> wsum(rows + 1, 0);
    for (int r = 0; r <= rows; r++) wsum[r] = wpre[0][r] + wpre[1][r];
    if ((long long)(grid / 2) * max_pair_rows < rows) return false;
    split(wsum.data(), 0, rows, grid / 2, cut, max_pair_rows);
    for (int b = 0; b < grid; b++) {
      T.blk_group[b] = b & 1;
      T.blk_row0[b] = cut[b >> 1];
      T.blk_row1[b] = cut[(b >> 1) + 1];
    }
  } else {
    int b0 = 0;
    for (int g = 0; g < ngroups; g++) {
      split(wpre[g].data(), 0, (int)wpre[g].size() - 1, nblk[g], cut);
      for (int bb = 0; bb < nblk[g]; bb++) {
        T.blk_group[b0 + bb] = g;
        T.blk_row0[b0 + bb] = cut[bb];
        T.blk_row1[b0 + bb] = cut[bb + 1];
      }
      b0 += nblk[g];
    }
  }
  T.max_block_rows = 0;
  for (int b = 0; b < grid; b++) T.max_block_rows = std::max(T.max_block_rows, T.blk_row1[b] - T.blk_row0[b]);
  T.lane_rows = lane_rows ? 1 : 0;
  if (getenv("OSQP_B200_DEBUG"))
    fprintf(stderr, "[osqp_b200] stream %dx%d groups=%d paired=%d split=%d stream rows=%d max_block_rows=%d lane_rows=%d\n",
            rows, cols, ngroups, T.paired, T.split, T.srows, T.max_block_rows, T.lane_rows);
  // first quad of every stream row and the stride between its quads (in quads)
  std::vector<std::vector<int>> sr_q0(ngroups);
  for (int g = 0; g < ngroups; g++) sr_q0[g].assign(wpre[g].size(), 0);
  int qstride = 1;
  T.w_qn.assign((size_t)grid * kWarps, 0);
  long long pos = 0;  // in entries (4 per quad)
  if (lane_rows) {
    // Lane-row layout.  The stream rows of a block are sorted by length (quads, descending; ties by index) and dealt
    // 32 at a time into slices: lane l of a slice owns one stream row and walks its quads one per chunk, so a row sum
    // is a private accumulator -- no segmented scan, no shuffles.  A slice is as long as its longest row (the sort
    // keeps the rows of a slice within a quad of each other; shorter rows are padded with zero quads).  Slices go to
    // the warps of the block longest-first onto the least loaded warp; a warp's slices are stored back to back.
    qstride = 32;
    T.w_s0.assign((size_t)grid * kWarps + 1, 0);
    std::vector<int> order, load(kWarps);
    std::vector<std::vector<int>> mine(kWarps);
    for (int b = 0; b < grid; b++) {
      const int g = T.blk_group[b], ra = T.blk_row0[b], rb = T.blk_row1[b];
      const long long *wp = wpre[g].data();
      order.resize(rb - ra);
      for (int i = 0; i < rb - ra; i++) order[i] = ra + i;
      std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return wp[x + 1] - wp[x] > wp[y + 1] - wp[y]; });
      const int nsl = (rb - ra + 31) / 32;
      std::fill(load.begin(), load.end(), 0);
      for (auto &v : mine) v.clear();
      for (int k = 0; k < nsl; k++) {  // slices come out longest first
        int w = 0;
        for (int j = 1; j < kWarps; j++) if (load[j] < load[w]) w = j;
        mine[w].push_back(k);
        load[w] += (int)(wp[order[(size_t)k * 32] + 1] - wp[order[(size_t)k * 32]]);
      }
      for (int w = 0; w < kWarps; w++) {
        const size_t wid = (size_t)b * kWarps + w;
        T.w_q0[wid] = (int)(pos / 4);
        T.w_s0[wid] = (int)T.sl_len.size();
        for (int k : mine[w]) {
          const int first = k * 32, cnt = std::min(32, rb - ra - first);
          const int L = (int)(wp[order[first] + 1] - wp[order[first]]);
          T.sl_len.push_back(L);
          for (int l = 0; l < 32; l++) {
            if (l < cnt) {
              const int sr = order[first + l];
              T.sl_row.push_back(sr);
              sr_q0[g][sr] = (int)(pos / 4) + l;
            } else {
              T.sl_row.push_back(-1);
            }
          }
          pos += 128LL * L;
        }
        T.w_qn[wid] = (int)(pos / 4) - T.w_q0[wid];
      }
    }
    T.w_s0[(size_t)grid * kWarps] = (int)T.sl_len.size();
    if (dbg) {  // how well the slices fill the warps: slowest warp of a block against the block's mean, and block against grid
      double worst = 0.0, sum_max = 0.0, gmax = 0.0, gsum = 0.0;
      for (int b = 0; b < grid; b++) {
        long long tot = 0, mx = 0;
        for (int w = 0; w < kWarps; w++) { const long long q = T.w_qn[(size_t)b * kWarps + w]; tot += q; mx = std::max(mx, q); }
        const double r = tot > 0 ? (double)mx * kWarps / (double)tot : 1.0;
        worst = std::max(worst, r); sum_max += r; gmax = std::max(gmax, (double)mx); gsum += (double)mx;
      }
      fprintf(stderr, "[osqp_b200]   lane rows: slowest warp / mean warp of a block: avg %.3f worst %.3f; slowest block / mean block %.3f\n",
              sum_max / grid, worst, gmax * grid / std::max(1.0, gsum));
    }
    // Few stream rows per block (or wildly different lengths inside a slice) leave lanes idle: past 25 % more stored
    // entries than the scan layout needs, use the scan layout
    if ((double)pos > 1.25 * (double)stored) {
      T.sl_len.clear(); T.sl_row.clear(); T.w_s0.clear();
      return build_tile_stream(mats, cols, grid, ngroups, paired, T, max_pair_rows, false);
    }
  } else {
    // scan layout: every warp streams a contiguous range of stream rows, quad by quad across its lanes
    std::vector<int> w_row1((size_t)grid * kWarps, 0);
    for (int b = 0; b < grid; b++) {
      const int g = T.blk_group[b];
      split(wpre[g].data(), T.blk_row0[b], T.blk_row1[b], kWarps, cut);
      for (int w = 0; w < kWarps; w++) {
        T.w_row0[(size_t)b * kWarps + w] = cut[w];
        w_row1[(size_t)b * kWarps + w] = cut[w + 1];
      }
    }
    // positions: warp by warp, stream row by stream row; every stream row is a whole number of quads and every warp's
    // stream starts on a chunk (32 quads) so that the value loads of a chunk are two fully coalesced 512 B rows
    for (int wid = 0; wid < grid * kWarps; wid++) {
      const int g = T.blk_group[wid / kWarps];
      T.w_q0[wid] = (int)(pos / 4);
      for (int sr = T.w_row0[wid]; sr < w_row1[wid]; sr++) {
        sr_q0[g][sr] = (int)(pos / 4);
        pos += 4 * (wpre[g][sr + 1] - wpre[g][sr]);
      }
      T.w_qn[wid] = (int)(pos / 4) - T.w_q0[wid];
      pos = (pos + 127) & ~127LL;
    }
  }
  T.w_q0[(size_t)grid * kWarps] = (int)(pos / 4);
  if (pos > 2147483000LL) return false;
  T.nelem = pos;
  mark("block / warp ranges, positions");
  T.cf.assign((size_t)pos + 8, 0);
  T.from_csr.resize(nnz);
  std::vector<int> cursor((size_t)rows * ngroups, 0);  // entries of (row, group) placed so far
  {
    int r0 = 0;
    long long k0 = 0;
    for (const CsrRef &M : mats) {
      par_ranges(M.rows, 8192, [&, r0, k0](long long ra, long long rb) {  // rows are independent of each other
        for (long long r = ra; r < rb; r++)
          for (int k = (*M.rowptr)[r]; k < (*M.rowptr)[r + 1]; k++) {
            const int c = (*M.col)[k], g = c / Wg;
            const int e = cursor[(size_t)(r0 + r) * ngroups + g]++;
            const int piece = (e >> 2) / split_quads;
            const int el = e - 4 * split_quads * piece;  // entry inside its stream row
            const int p = 4 * (sr_q0[g][T.sr_ptr[(size_t)g * (rows + 1) + r0 + r] + piece] + (el >> 2) * qstride) + (el & 3);
            T.cf[p] = (unsigned short)(c - g * Wg);
            T.from_csr[k0 + k] = stream_val_pos(p);
          }
      });
      r0 += M.rows;
      k0 += (*M.rowptr)[M.rows];
    }
  }
  mark("place entries");
  if (!lane_rows) {
    for (int g = 0; g < ngroups; g++)
      for (size_t sr = 0; sr + 1 < wpre[g].size(); sr++)
        T.cf[4 * sr_q0[g][sr] + 4 * (wpre[g][sr + 1] - wpre[g][sr]) - 1] |= 0x8000u;
    mark("row-end flags");
  }
  return true;
}

c_int upload_tile_stream(Engine &e, const TileStreamHost &h, TileStreamDev &t) {
  t.rows = h.rows; t.cols = h.cols; t.ngroups = h.ngroups; t.nelem = h.nelem;
  t.split = h.split; t.srows = h.srows;
  t.pf_chunks = std::max(0, std::min(32, env_int("OSQP_B200_PF", 4))) & ~3;
  CU_OK(dalloc(e, &t.val, (size_t)h.nelem + 8));
  CU_OK(dalloc(e, &t.cf, (size_t)h.nelem + 8));
  CU_OK(dalloc(e, &t.from_csr, (size_t)h.nnz));
  CU_OK(dalloc(e, &t.part, (size_t)h.ngroups * h.srows + 8));
#define UP(dst, vec)                                                                                         \
  CU_OK(dalloc(e, &dst, (vec).size()));                                                                      \
  CU_OK(cudaMemcpyAsync(dst, (vec).data(), (vec).size() * sizeof((vec)[0]), cudaMemcpyHostToDevice, e.stream))
  UP(t.blk_group, h.blk_group); UP(t.grp_col0, h.grp_col0); UP(t.w_row0, h.w_row0); UP(t.w_q0, h.w_q0);
  UP(t.w_qn, h.w_qn); UP(t.blk_row0, h.blk_row0); UP(t.blk_row1, h.blk_row1); UP(t.sr_ptr, h.sr_ptr);
  t.lane_rows = h.lane_rows;
  if (h.lane_rows) { UP(t.w_s0, h.w_s0); UP(t.sl_len, h.sl_len); UP(t.sl_row, h.sl_row); }
  t.paired = h.paired;
#undef UP
  CU_OK(cudaMemcpyAsync(t.cf, h.cf.data(), (size_t)h.nelem * sizeof(unsigned short), cudaMemcpyHostToDevice, e.stream));
  if (h.nnz > 0)
    CU_OK(cudaMemcpyAsync(t.from_csr, h.from_csr.data(), (size_t)h.nnz * sizeof(int), cudaMemcpyHostToDevice, e.stream));
  CU_OK(cudaStreamSynchronize(e.stream));  // the host vectors go out of scope
  return 0;
}

c_int upload_partition(Engine &e, const std::vector<int> &A_rowptr, const std::vector<int> &At_rowptr,
                       const std::vector<int> &P_rowptr, std::vector<int> &ms, std::vector<int> &ns) {
  const int n = e.d.n, m = e.d.m, grid = e.geom.grid;
  std::vector<long long> wm(m + 1, 0), wn(n + 1, 0);
  // Ownership serves two kinds of work: the element-wise owner phases of every PCG iteration (cost per ROW: one L2
  // round trip per 512 rows of a block) and the CSR products of update_info / the residual refresh (cost per
  // NON-ZERO, once per 25 ADMM iterations).  A row therefore weighs its non-zeros plus a constant that dominates for
  // short rows: a block that owns only one-entry rows (identity blocks of a Lasso / MPC matrix) must not own 100x
  // more rows than the others.
  const long long kRowCost = 256;
  for (int i = 0; i < m; i++) wm[i + 1] = wm[i] + (A_rowptr[i + 1] - A_rowptr[i]) + kRowCost;
  for (int j = 0; j < n; j++)
    wn[j + 1] = wn[j] + (P_rowptr[j + 1] - P_rowptr[j]) + (m > 0 ? At_rowptr[j + 1] - At_rowptr[j] : 0) + kRowCost;
  balanced_split(wm, m, grid, ms);
  balanced_split(wn, n, grid, ns);
  CU_OK(cudaMemcpyAsync(e.d.m_start, ms.data(), (grid + 1) * sizeof(int), cudaMemcpyHostToDevice, e.stream));
  CU_OK(cudaMemcpyAsync(e.d.n_start, ns.data(), (grid + 1) * sizeof(int), cudaMemcpyHostToDevice, e.stream));
  CU_OK(cudaStreamSynchronize(e.stream));  // ms/ns go out of scope
  return 0;
}

// The cluster-pair layout of the [A; P] stream is also a valid unpaired layout (every block still writes its partial
// row sums to part[group][row] and the owners add them), so a clustered cooperative launch that the runtime refuses
// (seen under ncu's launch interception) falls back to the plain cooperative launch for the rest of the workspace.
template <typename Launch>
cudaError_t launch_with_pair_fallback(Engine &e, Launch launch) {
  cudaError_t err = launch();
  if (err != cudaSuccess && e.geom.cluster > 1) {
    cudaGetLastError();
    fprintf(stderr, "WARNING osqp_b200: clustered cooperative launch refused (%s); continuing without cluster pairs\n",
            cudaGetErrorString(err));
    e.geom.cluster = 1;
    e.d.SA.paired = 0;
    e.prof.paired = 0;
    err = launch();
  }
  return err;
}

c_int push_state(Engine &e) {
  CU_OK(cudaMemcpyAsync(e.d.state, e.h_state, sizeof(DevState), cudaMemcpyHostToDevice, e.stream));
  return 0;
}
c_int pull_state(Engine &e) {
  CU_OK(cudaMemcpyAsync(e.h_state, e.d.state, sizeof(DevState), cudaMemcpyDeviceToHost, e.stream));
  CU_OK(cudaStreamSynchronize(e.stream));
  return 0;
}

// scale_data (a2) + set_rho_vec (a3) + preconditioner; `keep_rho_types`: libosqp's update_P/A keeps rho_vec
c_int rescale_and_refresh(Engine &e, bool reset_rho_types) {
  CU_OK(launch_scale_data(e.d, (int)e.st.scaling, e.st.sigma, e.stream));
  e.prof.launches += 2 + 5 * (int)e.st.scaling;
  if (reset_rho_types) {
    CU_OK(launch_set_rho_vec(e.d, e.st.rho, 0, e.stream));
    e.prof.launches += 1;
  }
  CU_OK(launch_precond(e.d, e.st.sigma, e.stream));
  CU_OK(launch_fill_blocked(e.d, e.stream));
  CU_OK(launch_fill_wood(e.d, e.stream));
  e.wood_dirty = true;
  e.prof.launches += 2 + (e.d.W.w > 0 ? 1 : 0);
  return 0;
}

// Convexity check of osqp_setup / osqp_update_P (libosqp: the LDL' of the KKT matrix must have n positive pivots, i.e.
// the scaled P + sigma I must be positive definite; test/non_convex.jl:13-21 needs setup to FAIL otherwise).  Three
// tiers, cheapest first: (1) Gershgorin certificate on the device, one pass over P; (2) n <= kDenseCholMax: exact --
// dense Cholesky of the scaled P + sigma I on the host (the verdict of a factorisation, like libosqp's); (3) larger n:
// CG / Lanczos curvature probe on the device from several start vectors with a residual-based stopping rule and an
// iteration budget that grows with n.  (3) cannot prove definiteness -- an indefinite direction that none of the
// start vectors excites above 1e-10 escapes it; PCG breakdown inside the ADMM loop (Non_convex) is the second net.
constexpr int kDenseCholMax = 640;

c_int convexity_check(Engine &e, int *nonconvex) {
  DevPtrs &d = e.d;
  const int n = d.n;
  *nonconvex = 0;
  { c_int rc = pull_state(e); if (rc) return rc; }  // the cost scaling c was just written on the device
  e.h_state->pd_certified = 1;
  e.h_state->pd_check_failed = 0;
  { c_int rc = push_state(e); if (rc) return rc; }
  CU_OK(launch_gershgorin(d, e.st.sigma, e.stream));
  e.prof.launches += 1;
  { c_int rc = pull_state(e); if (rc) return rc; }
  if (e.h_state->pd_certified && !env_int("OSQP_B200_NO_GERSHGORIN", 0)) return 0;
  if (n <= env_int("OSQP_B200_DENSE_CHOL_MAX", kDenseCholMax)) {
    const long long nnzP = d.P.nnz;
    std::vector<int> rp(n + 1), ci(nnzP);
    std::vector<double> val(nnzP);
    CU_OK(cudaMemcpyAsync(rp.data(), d.P.rowptr, (size_t)(n + 1) * sizeof(int), cudaMemcpyDeviceToHost, e.stream));
    if (nnzP > 0) {
      CU_OK(cudaMemcpyAsync(ci.data(), d.P.col, (size_t)nnzP * sizeof(int), cudaMemcpyDeviceToHost, e.stream));
      CU_OK(cudaMemcpyAsync(val.data(), d.P.val, (size_t)nnzP * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
    }
    CU_OK(cudaStreamSynchronize(e.stream));
    std::vector<double> L((size_t)n * n, 0.0);  // row-major, lower triangle
    for (int i = 0; i < n; i++) {
      for (int k = rp[i]; k < rp[i + 1]; k++)
        if (ci[k] <= i) L[(size_t)i * n + ci[k]] += val[k];
      L[(size_t)i * n + i] += e.st.sigma;
    }
    for (int j = 0; j < n && !*nonconvex; j++) {  // left-looking: the inner loops are contiguous dot products
      const double *Lj = &L[(size_t)j * n];
      double dj = Lj[j];
      for (int k = 0; k < j; k++) dj -= Lj[k] * Lj[k];
      if (!(dj > 0.0)) { *nonconvex = 1; break; }
      const double r = std::sqrt(dj), rinv = 1.0 / r;
      L[(size_t)j * n + j] = r;
      for (int i = j + 1; i < n; i++) {
        double *Li = &L[(size_t)i * n];
        double a = Li[j];
        for (int k = 0; k < j; k++) a -= Li[k] * Lj[k];
        Li[j] = a * rinv;
      }
    }
    return 0;
  }
  const int budget = env_int("OSQP_B200_PD_PROBE_ITERS", std::min(n, 2000));
  CU_OK(launch_pd_probe(d, e.geom, e.st.sigma, budget, env_int("OSQP_B200_PD_PROBE_TRIALS", 3), e.stream));
  e.prof.launches += 1;
  { c_int rc = pull_state(e); if (rc) return rc; }
  *nonconvex = e.h_state->pd_check_failed;
  e.h_state->pd_check_failed = 0;
  return push_state(e);
}

void publish(Engine &e) {
  OSQPWorkspace &p = e.pub;
  memset(&p, 0, sizeof(p));
  e.data_pub.n = e.d.n;
  e.data_pub.m = e.d.m;
  e.data_pub.P = nullptr;
  e.data_pub.A = nullptr;
  e.data_pub.q = nullptr;
  e.data_pub.l = nullptr;
  e.data_pub.u = nullptr;
  e.sol_pub.x = e.h_sol_x;
  e.sol_pub.y = e.h_sol_y;
  p.data = &e.data_pub;
  p.delta_y = e.h_dy;
  p.delta_x = e.h_dx;
  p.settings = &e.st;
  p.solution = &e.sol_pub;
  p.info = &e.info;
  p.first_run = e.first_run ? 1 : 0;
  p.summary_printed = 0;
}

void print_setup_header(const Engine &e) {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, e.device);
  printf("-----------------------------------------------------------------\n");
  printf("     OSQP ADMM engine for NVIDIA B200 (sm_100a) -- libosqp ABI\n");
  printf("-----------------------------------------------------------------\n");
  printf("problem:  variables n = %d, constraints m = %d\n", e.d.n, e.d.m);
  printf("          nnz(P) + nnz(A) = %lld\n", (long long)(e.nnzPtriu + e.nnzA));
  printf("device:   %s (#%d), persistent grid %d x %d threads\n", prop.name, e.device, e.geom.grid, e.geom.block);
  printf("settings: linear system solver = reduced-KKT Jacobi-PCG on CSR SpMV,\n");
  printf("          eps_abs = %.1e, eps_rel = %.1e,\n", e.st.eps_abs, e.st.eps_rel);
  printf("          eps_prim_inf = %.1e, eps_dual_inf = %.1e,\n", e.st.eps_prim_inf, e.st.eps_dual_inf);
  printf("          rho = %.2e %s, sigma = %.2e, alpha = %.2f, max_iter = %lld\n", e.st.rho,
         e.st.adaptive_rho ? "(adaptive)" : "", e.st.sigma, e.st.alpha, (long long)e.st.max_iter);
  printf("          scaling: %s, polish: %s, warm start: %s\n\n", e.st.scaling ? "on" : "off",
         e.st.polish ? "on" : "off", e.st.warm_start ? "on" : "off");
}

// algorithmic bytes of one SpMV in the engine's own storage (DESIGN.md "bytes model"): fp64 value + 16-bit local
// column per non-zero, the gathered vector read once, the result written once
double spmv_bytes(long long nnz, long long rows, long long cols) {
  return 10.0 * (double)nnz + 8.0 * (double)cols + 8.0 * (double)rows;
}

c_int upload_vector(Engine &e, double *dst, const c_float *src, long long count) {
  if (count > 0) CU_OK(cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  return 0;
}

constexpr int kTinyMax = 256;

// (Re-)create the tiny engine from the host copies of the problem; `rho` is the rho to start from.  Iterates are not
// carried over by this call (the caller warm-starts from the last solution when there is one).
c_int tiny_setup(Engine &e, double rho) {
  if (e.tiny) { osqp_batch_cleanup(e.tiny); e.tiny = nullptr; }
  const c_int n = e.d.n, m = e.d.m;
  csc Pc{}, Ac{};
  Pc.nzmax = (c_int)e.tP_x.size(); Pc.m = n; Pc.n = n; Pc.p = e.tP_p.data(); Pc.i = e.tP_i.data(); Pc.x = e.tP_x.data(); Pc.nz = -1;
  Ac.nzmax = (c_int)e.tA_x.size(); Ac.m = m; Ac.n = n; Ac.p = e.tA_p.data(); Ac.i = e.tA_i.data(); Ac.x = e.tA_x.data(); Ac.nz = -1;
  OSQPData pat{};
  pat.n = n; pat.m = m; pat.P = &Pc; pat.A = &Ac;
  OSQPSettings st = e.st;
  st.rho = rho;
  st.verbose = 0;
  st.polish = 0;
  const c_int rc = osqp_batch_setup(&e.tiny, 1, &pat, e.tP_x.data(), e.tA_x.data(), e.tq.data(), e.l0.data(), e.u0.data(), &st);
  if (rc != 0) e.tiny = nullptr;  // the general engine still serves the workspace
  return 0;
}

bool use_tiny(const Engine &e) {
  return e.tiny != nullptr && !e.st.polish && e.st.time_limit == 0 && !e.st.verbose;
}

bool has_solution(c_int sv) {
  return sv == OSQP_SOLVED || sv == OSQP_SOLVED_INACCURATE || sv == OSQP_MAX_ITER_REACHED;
}

c_int tiny_solve(Engine &e) {
  if (e.clear_update_time) e.info.update_time = 0.0;
  const double t0 = now_s();
  const int n = e.d.n, m = e.d.m;
  OSQPB200Batch *b = e.tiny;
  const OSQPSettings &s = e.st;
  osqp_batch_update_setting(b, "max_iter", (double)s.max_iter);
  osqp_batch_update_setting(b, "eps_abs", s.eps_abs);
  osqp_batch_update_setting(b, "eps_rel", s.eps_rel);
  osqp_batch_update_setting(b, "eps_prim_inf", s.eps_prim_inf);
  osqp_batch_update_setting(b, "eps_dual_inf", s.eps_dual_inf);
  osqp_batch_update_setting(b, "alpha", s.alpha);
  osqp_batch_update_setting(b, "check_termination", (double)s.check_termination);
  osqp_batch_update_setting(b, "warm_start", (double)s.warm_start);
  osqp_batch_update_setting(b, "scaled_termination", (double)s.scaled_termination);
  if (s.adaptive_rho && s.adaptive_rho_interval == 0) {  // the same deterministic rule as solve_impl
    const long long N = s.check_termination > 0 ? s.check_termination : 25;
    long long iv = ((50 + N / 2) / N) * N;
    if (iv < N) iv = N;
    e.st.adaptive_rho_interval = iv;
  }
  osqp_batch_update_setting(b, "adaptive_rho_interval", (double)s.adaptive_rho_interval);
  if (e.last_path == 1 && s.warm_start && has_solution(e.info.status_val))  // the general engine solved last: hand over
    osqp_batch_warm_start(b, e.h_sol_x, m > 0 ? e.h_sol_y : nullptr);
  const c_float *x = nullptr, *y = nullptr;
  const OSQPB200BatchInfo *bi = nullptr;
  const c_int rc = osqp_batch_solve_view(b, &x, &y, &bi);
  if (rc != 0) return rc;
  const c_int sv = bi->status_val;
  e.info.iter = bi->iter;
  update_status(e.info, sv);
  e.info.status_polish = 0;
  e.info.obj_val = bi->obj_val;
  e.info.pri_res = bi->pri_res;
  e.info.dua_res = bi->dua_res;
  e.info.rho_updates = bi->rho_updates;
  e.info.rho_estimate = bi->rho_estimate;
  const bool pinf = sv == OSQP_PRIMAL_INFEASIBLE || sv == OSQP_PRIMAL_INFEASIBLE_INACCURATE;
  const bool dinf = sv == OSQP_DUAL_INFEASIBLE || sv == OSQP_DUAL_INFEASIBLE_INACCURATE;
  if (pinf || dinf) {  // the batched engine returns the certificate in place of the solution
    if (pinf) memcpy(e.h_dy, y, (size_t)m * sizeof(double));
    if (dinf) memcpy(e.h_dx, x, (size_t)n * sizeof(double));
    for (int j = 0; j < n; j++) e.h_sol_x[j] = NAN;
    for (int i = 0; i < m; i++) e.h_sol_y[i] = NAN;
  } else {
    memcpy(e.h_sol_x, x, (size_t)n * sizeof(double));
    if (m > 0) memcpy(e.h_sol_y, y, (size_t)m * sizeof(double));
  }
  e.prof.kernel_ms = osqp_batch_last_kernel_ms(b);
  e.prof.polish_ms = 0;
  e.prof.admm_iters = bi->iter;
  e.prof.pcg_iters = 0;
  e.prof.launches += 1;
  e.last_path = 2;
  const double base = e.first_run ? e.info.setup_time : e.info.update_time;
  e.info.solve_time = now_s() - t0;
  e.info.polish_time = 0;
  e.info.run_time = base + e.info.solve_time;
  e.first_run = false;
  e.clear_update_time = true;
  e.pub.first_run = 0;
  return 0;
}

// After a change that the batched engine cannot take in place (P / A values, rho): rebuild it and restart it from the
// last solution, as libosqp keeps its iterates across osqp_update_P / osqp_update_rho.
c_int tiny_rebuild(Engine &e) {
  if (!e.tiny) return 0;
  const bool carry = e.last_path == 2 && has_solution(e.info.status_val);
  std::vector<double> x, y;
  if (carry) { x.assign(e.h_sol_x, e.h_sol_x + e.d.n); y.assign(e.h_sol_y, e.h_sol_y + e.d.m); }
  tiny_setup(e, e.st.rho);
  if (e.tiny && carry) osqp_batch_warm_start(e.tiny, x.data(), e.d.m > 0 ? y.data() : nullptr);
  return 0;
}

}  // namespace

// =================================================================== C ABI
extern "C" {

void osqp_set_default_settings(OSQPSettings *s) {  // src/types.jl:138-143; SURVEY Appendix A
  memset(s, 0, sizeof(*s));
  s->rho = 0.1; s->sigma = 1e-6; s->scaling = 10;
  s->adaptive_rho = 1; s->adaptive_rho_interval = 0; s->adaptive_rho_tolerance = 5; s->adaptive_rho_fraction = 0.4;
  s->max_iter = 4000; s->eps_abs = 1e-3; s->eps_rel = 1e-3; s->eps_prim_inf = 1e-4; s->eps_dual_inf = 1e-4;
  s->alpha = 1.6; s->linsys_solver = QDLDL_SOLVER; s->delta = 1e-6; s->polish = 0; s->polish_refine_iter = 3;
  s->verbose = 1; s->scaled_termination = 0; s->check_termination = 25; s->warm_start = 1; s->time_limit = 0;
}

const char *osqp_version(void) { return "0.6.2-b200"; }  // src/interface.jl:220

c_int osqp_b200_debug_read(OSQPWorkspace *work, unsigned long long *out, c_int count) {
  if (!work) return 1;
  Engine &e = *E(work);
  DeviceGuard guard(e.device);
  const c_int have = 16 * (c_int)e.geom.grid;
  CU_OK(cudaMemcpy(out, e.d.dbg, (size_t)std::min(count, have) * 8, cudaMemcpyDeviceToHost));
  return 0;
}

// Host-only check of the tile-stream format (no GPU needed): builds the stream of one CSR matrix exactly as
// osqp_setup does and replays the device algorithm of kernels.cu stream_phase lane by lane (quads, per-lane row-end
// flags, rank by flag count, segmented scan with the carry across chunks) on the CPU.  y_out = M x.
c_int osqp_b200_stream_selftest(c_int rows, c_int cols, const c_int *rowptr, const c_int *col, const c_float *val,
                                const c_float *x, c_int grid, c_int ngroups, c_int paired, c_float *y_out,
                                c_float *padding_ratio) {
  std::vector<int> rp(rows + 1), ci(rowptr[rows]);
  for (c_int r = 0; r <= rows; r++) rp[r] = (int)rowptr[r];
  for (c_int k = 0; k < rowptr[rows]; k++) ci[k] = (int)col[k];
  TileStreamHost T;
  std::vector<CsrRef> mats{CsrRef{&rp, &ci, (int)rows}};
  if (!build_tile_stream(mats, (int)cols, (int)grid, (int)ngroups, paired != 0, T, 1 << 30, env_int("OSQP_B200_LANE_ROWS", 1) != 0))
    return 2;
  if (paired && !T.paired) return 2;
  if (T.paired)  // both blocks of a cluster pair must cover the same rows
    for (int b = 0; b < (int)grid; b += 2)
      if (T.blk_row0[b] != T.blk_row0[b + 1] || T.blk_row1[b] != T.blk_row1[b + 1] || T.blk_group[b] != 0 ||
          T.blk_group[b + 1] != 1)
        return 10;
  if (padding_ratio) {  // entries actually streamed (quad padding, zero quads) per non-zero
    long long quads = 0;
    for (int q : T.w_qn) quads += q;
    *padding_ratio = T.nnz ? 4.0 * (double)quads / (double)T.nnz : 1.0;
  }
  std::vector<double> sv((size_t)T.nelem + 8, 0.0);
  for (long long k = 0; k < T.nnz; k++) sv[T.from_csr[k]] = val[k];
  std::vector<double> part((size_t)T.ngroups * T.srows, 0.0);
  std::vector<char> written((size_t)T.ngroups * T.srows, 0);
  for (int wid = 0; wid < (int)grid * kWarps && T.lane_rows; wid++) {
    // lane-row layout: replay of kernels.cu stream_phase_lr -- every lane accumulates its own stream row over the
    // chunks of a slice and stores it when the slice ends
    const int grp = T.blk_group[wid / kWarps];
    const int q0 = T.w_q0[wid], L = T.w_qn[wid];
    if (q0 % 32 != 0 || L % 32 != 0) return 3;
    const double *xs = x + T.grp_col0[grp];
    const int slice = T.grp_col0[grp + 1] - T.grp_col0[grp];
    int chunk = 0;
    for (int sl = T.w_s0[wid]; sl < T.w_s0[wid + 1]; sl++) {
      for (int lane = 0; lane < 32; lane++) {
        double acc = 0.0;
        for (int t = 0; t < T.sl_len[sl]; t++) {
          const long long e = 4ll * (q0 + 32 * (chunk + t) + lane);
          double inc = 0.0;
          for (int k = 0; k < 4; k++) {
            const unsigned w = T.cf[e + k];
            if (w & 0x8000u) return 4;  // no flags in this layout
            const double v = sv[stream_val_pos((int)(e + k))];
            if (v != 0.0 && (int)w >= slice) return 5;
            inc = (k == 0) ? v * xs[w] : std::fma(v, xs[w], inc);
          }
          acc += inc;
        }
        const int sr = T.sl_row[(size_t)sl * 32 + lane];
        if (sr < 0) { if (acc != 0.0) return 8; continue; }
        if (sr < T.blk_row0[wid / kWarps] || sr >= T.blk_row1[wid / kWarps]) return 11;
        part[(size_t)grp * T.srows + sr] = acc;
        written[(size_t)grp * T.srows + sr]++;
      }
      chunk += T.sl_len[sl];
    }
    if (32 * chunk != L) return 6;
  }
  for (int wid = 0; wid < (int)grid * kWarps && !T.lane_rows; wid++) {
    const int grp = T.blk_group[wid / kWarps];
    const int q0 = T.w_q0[wid], L = T.w_qn[wid];
    if (L <= 0) continue;
    if (q0 % 32 != 0) return 3;
    if (T.w_row0[wid] < T.blk_row0[wid / kWarps] || T.w_row0[wid] > T.blk_row1[wid / kWarps]) return 11;
    const double *xs = x + T.grp_col0[grp];
    const int slice = T.grp_col0[grp + 1] - T.grp_col0[grp];
    double *out = part.data() + (size_t)grp * T.srows + T.w_row0[wid];
    char *wr = written.data() + (size_t)grp * T.srows + T.w_row0[wid];
    int rdone = 0;
    double carry = 0.0;
    for (int cc = 0; cc < L; cc += 32) {
      double run = carry;
      int nflag = 0;
      for (int lane = 0; lane < 32; lane++) {
        double acc = 0.0;
        bool flag = false;
        if (cc + lane < L) {
          const long long e = 4ll * (q0 + cc + lane);
          for (int t = 0; t < 4; t++) {
            const unsigned w = T.cf[e + t];
            const int lc = (int)(w & 0x7fffu);
            if ((w & 0x8000u) && t != 3) return 4;  // only the last word of a quad may carry the flag
            const double v = sv[stream_val_pos((int)(e + t))];
            if (v != 0.0 && lc >= slice) return 5;
            acc = (t == 0) ? v * xs[lc] : std::fma(v, xs[lc], acc);
          }
          flag = (T.cf[e + 3] & 0x8000u) != 0;
        }
        run += acc;
        if (flag) {
          if (T.w_row0[wid] + rdone + nflag >= T.srows) return 6;
          out[rdone + nflag] = run;
          wr[rdone + nflag]++;
          nflag++;
          run = 0.0;
        }
      }
      carry = run;
      rdone += nflag;
    }
    if (carry != 0.0) return 8;  // the last quad of a warp's stream must close its row
  }
  for (int g = 0; g < T.ngroups; g++)
    for (int sr = 0; sr < T.sr_ptr[(size_t)g * (rows + 1) + rows]; sr++)
      if (written[(size_t)g * T.srows + sr] != 1) return 9;
  for (c_int r = 0; r < rows; r++) {  // part_sum of kernels.cu
    double a = 0.0;
    for (int g = 0; g < T.ngroups; g++)
      for (int sr = T.sr_ptr[(size_t)g * (rows + 1) + r]; sr < T.sr_ptr[(size_t)g * (rows + 1) + r + 1]; sr++)
        a += part[(size_t)g * T.srows + sr];
    y_out[r] = a;
  }
  return 0;
}

#ifdef OSQP_B200_DEVTOOLS  // measurement / self-test entry points: lib/libosqp_dev.so only
// Stream micro-benchmark (kernels.cu membench_kernel): GB/s of reading `mbytes` MB with the load shape of the
// tile-stream phase.  pattern 0/1/2, depth = chunks in flight per lane.  Standalone: needs no workspace.
c_float osqp_b200_membench(c_int mbytes, c_int pattern, c_int depth, c_int reps) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return -1.0;
  cudaDeviceProp prop;
  int dev = 0;
  cudaGetDevice(&dev);
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return -1.0;
  const int grid = prop.multiProcessorCount;
  long long bytes = (long long)mbytes * 1000000LL;
  bytes -= bytes % (1280LL * grid * kWarps);
  char *buf = nullptr;
  double *sink = nullptr;
  if (cudaMalloc(&buf, bytes + 4096) != cudaSuccess) return -1.0;
  cudaMalloc(&sink, 64);
  cudaMemset(buf, 0, bytes + 4096);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch_membench(buf, bytes, (int)pattern, (int)depth, grid, sink, 0);
  cudaEventRecord(e0, 0);
  for (c_int r = 0; r < reps; r++) launch_membench(buf, bytes, (int)pattern, (int)depth, grid, sink, 0);
  cudaEventRecord(e1, 0);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  cudaFree(sink);
  if (err != cudaSuccess || ms <= 0.f) return -1.0;
  return (double)bytes * (double)reps / ((double)ms * 1e-3) / 1e9;
}

// Grid-barrier micro-benchmark on a workspace's persistent grid: ns per barrier (kernels.cu barrier_bench_kernel).
c_float osqp_b200_barrier_bench(OSQPWorkspace *work, c_int iters, c_int mode) {
  if (!work || iters <= 0) return -1.0;
  Engine &e = *E(work);
  DeviceGuard guard(e.device);
  double *sink = nullptr;
  unsigned long long *ns = nullptr, h = 0;
  if (cudaMalloc(&sink, (100000 + 128) * sizeof(double)) != cudaSuccess) return -1.0;
  cudaMalloc(&ns, sizeof(unsigned long long));
  cudaError_t err = launch_barrier_bench(e.d, e.geom, (int)iters, (int)mode, sink, ns, e.stream);
  if (err == cudaSuccess) err = cudaMemcpyAsync(&h, ns, sizeof(h), cudaMemcpyDeviceToHost, e.stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize(e.stream);
  cudaFree(sink);
  cudaFree(ns);
  e.h_state->needs_refresh = 1;
  return err == cudaSuccess ? (double)h / (double)iters : -1.0;
}

c_int osqp_b200_cluster_probe(OSQPWorkspace *work, c_int csize) {
  if (!work) return -1;
  Engine &e = *E(work);
  DeviceGuard guard(e.device);
  return max_active_clusters((int)csize, e.geom.block, e.geom.dyn_smem);
}

// Cross-block reductions of the persistent grid on known data (kernels.cu reduce_selftest_kernel): out[6] =
// {tree sum, tree max, fixed-point sum, fixed-point max, second fixed-point sum, max}.
c_int osqp_b200_reduce_selftest(OSQPWorkspace *work, c_float ref, c_float *out) {
  if (!work || !out) return 1;
  Engine &e = *E(work);
  DeviceGuard guard(e.device);
  double *dout = nullptr;
  CU_OK(cudaMalloc(&dout, 8 * sizeof(double)));
  cudaError_t err = launch_reduce_selftest(e.d, e.geom, ref, dout, e.stream);
  if (err == cudaSuccess) err = cudaMemcpyAsync(out, dout, 6 * sizeof(double), cudaMemcpyDeviceToHost, e.stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize(e.stream);
  cudaFree(dout);
  e.h_state->needs_refresh = 1;
  CU_OK(err);
  return 0;
}

#endif  // OSQP_B200_DEVTOOLS

c_int osqp_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

c_int osqp_cleanup(OSQPWorkspace *work) {  // src/interface.jl:224-229 -- NULL and finalizer-thread safe
  if (work) destroy(E(work));
  return 0;
}

// The fp32 copies of the matrix values (DevPtrs::mat32) hold SCALED entries: Ruiz scaling multiplies an entry by at most
// 1e4 * 1e4 (the clamp on D, E) and the cost by at most 1e4 more, so raw entries below 1e25 stay far inside the fp32
// range.  Data beyond that (nothing a QP solver with OSQP_INFTY = 1e30 is meant for) keeps fp64 values everywhere.
static bool fits_fp32_copy(const c_float *x, long long k) {
  for (long long t = 0; t < k; t++)
    if (!(std::fabs(x[t]) < 1e25)) return false;
  return true;
}

c_int osqp_setup(OSQPWorkspace **workp, const OSQPData *data, const OSQPSettings *settings) {
  if (workp) *workp = nullptr;
  if (!workp) return 1;
  if (validate_data(data)) return 1;
  if (validate_settings(settings)) return 1;
  const double t0 = now_s();
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    fprintf(stderr, "ERROR in osqp_setup: no CUDA device available -- this engine has no CPU fallback\n");
    return 10;
  }
  Engine *ep = new Engine();
  Engine &e = *ep;
  memset(&e.prof, 0, sizeof(e.prof));
  int cur = 0;
  cudaGetDevice(&cur);
  e.device = env_int("OSQP_B200_DEVICE", cur);
  struct Fail {
    Engine *p;
    ~Fail() { if (p) destroy(p); }
  } fail{ep};
  DeviceGuard guard(e.device);
  if (!guard.ok) { fprintf(stderr, "ERROR in osqp_setup: cannot select CUDA device %d\n", e.device); return 11; }
  CU_OK(cudaStreamCreateWithFlags(&e.stream, cudaStreamNonBlocking));
  CU_OK(cudaEventCreate(&e.ev0));
  CU_OK(cudaEventCreate(&e.ev1));
  CU_OK(cudaEventCreate(&e.ev2));
  e.st = *settings;
  e.pcg_eta = env_double("OSQP_B200_PCG_ETA", e.pcg_eta);
  e.pcg_floor = env_double("OSQP_B200_PCG_FLOOR", e.pcg_floor);
  e.pcg_eta_e = env_double("OSQP_B200_PCG_ETA_E", e.pcg_eta_e);
  e.refresh_every = env_int("OSQP_B200_REFRESH_EVERY", e.refresh_every);
  e.polish_penalty = env_double("OSQP_B200_POLISH_PENALTY", e.polish_penalty);
  const int n = (int)data->n, m = (int)data->m;
  e.pcg_max_iter = env_int("OSQP_B200_PCG_MAX_ITER", std::max(20, std::min(10 * n, 2000)));
  DevPtrs &d = e.d;
  d.n = n;
  d.m = m;
  const csc *Pc = data->P, *Ac = data->A;
  const long long nnzA = Ac->p[n], nnzPt = Pc->p[n];
  e.nnzA = nnzA;
  e.nnzPtriu = nnzPt;

  const bool dbg_time = env_int("OSQP_B200_DEBUG", 0) != 0;
  double t_mark = now_s();
  auto mark = [&](const char *what) {
    if (dbg_time) {
      const double t = now_s();
      fprintf(stderr, "[osqp_b200] setup %-28s %8.1f ms\n", what, (t - t_mark) * 1e3);
      t_mark = t;
    }
  };
  // ---- host index work: A' CSR = caller's CSC; A CSR by counting sort; P full symmetric CSR
  std::vector<int> At_rowptr(n + 1), At_col(nnzA), A_rowptr(m + 1, 0), A_col(nnzA), mapA(nnzA);
  std::vector<double> A_val(nnzA);
  for (int j = 0; j <= n; j++) At_rowptr[j] = (int)Ac->p[j];
  {
    bool bad = false;
    for (long long k = 0; k < nnzA; k++) {
      const c_int r = Ac->i[k];
      if (r < 0 || r >= m) { bad = true; break; }
      At_col[k] = (int)r;
    }
    if (bad) { fprintf(stderr, "ERROR in osqp_setup: row index out of range in A\n"); return 1; }
  }
  // Stable counting sort by row in O(nnz + T m): thread t takes the t-th contiguous range of columns (balanced on
  // non-zeros), counts its entries per row into its own histogram, the histograms are turned into per-thread write
  // cursors (row start + what the threads before it hold of that row), and every thread places its entries.  Threads
  // are ordered like the columns, so the order inside a row (ascending column) and every position are those of the
  // sequential sort, whatever the number of threads.
  const int T = (int)std::max<long long>(1, std::min<long long>(host_threads(), (nnzA + nnzPt) / 65536));
  auto col_cuts = [&](const c_int *colptr, std::vector<int> &cut) {
    cut.assign(T + 1, n);
    cut[0] = 0;
    int j = 0;
    for (int t = 1; t < T; t++) {
      const long long target = colptr[n] * (long long)t / T;
      while (j < n && colptr[j] < target) j++;
      cut[t] = j;
    }
  };
  auto run_threads = [&](auto fn) {
    if (T == 1) { fn(0); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++) th.emplace_back([&fn, t]() { fn(t); });
    for (std::thread &x : th) x.join();
  };
  // hist[t][r] -> cursors; rowptr gets the totals' prefix sum
  auto cursors = [&](std::vector<int> &hist, int rows, std::vector<int> &rowptr) {
    par_ranges(rows, 16384, [&](long long r0, long long r1) {
      for (long long r = r0; r < r1; r++) {
        int tot = 0;
        for (int t = 0; t < T; t++) tot += hist[(size_t)t * rows + r];
        rowptr[r + 1] = tot;
      }
    });
    rowptr[0] = 0;
    for (int r = 0; r < rows; r++) rowptr[r + 1] += rowptr[r];
    par_ranges(rows, 16384, [&](long long r0, long long r1) {
      for (long long r = r0; r < r1; r++) {
        int run = rowptr[r];
        for (int t = 0; t < T; t++) {
          int &h = hist[(size_t)t * rows + r];
          const int c = h;
          h = run;
          run += c;
        }
      }
    });
  };
  if (m > 0) {
    std::vector<int> cut, hist((size_t)T * m, 0);
    col_cuts(Ac->p, cut);
    run_threads([&](int t) {
      int *h = hist.data() + (size_t)t * m;
      for (long long k = Ac->p[cut[t]]; k < Ac->p[cut[t + 1]]; k++) h[At_col[k]]++;
    });
    cursors(hist, m, A_rowptr);
    run_threads([&](int t) {
      int *h = hist.data() + (size_t)t * m;
      for (int j = cut[t]; j < cut[t + 1]; j++)
        for (c_int k = Ac->p[j]; k < Ac->p[j + 1]; k++) {
          const int pos = h[At_col[k]]++;
          A_col[pos] = j;
          A_val[pos] = Ac->x[k];
          mapA[k] = pos;
        }
    });
  }
  // full symmetric CSR of P from the upper triangle, same scheme: entry (i, j) of column j goes to row i and, off the
  // diagonal, its mirror (j, i) to row j; rows come out sorted by column
  std::vector<int> P_rowptr(n + 1, 0), mapP1(nnzPt), mapP2(nnzPt);
  {
    bool bad = false;
    for (long long k = 0; k < nnzPt && !bad; k++) bad = Pc->i[k] < 0 || Pc->i[k] >= n;
    if (bad) { fprintf(stderr, "ERROR in osqp_setup: row index out of range in P\n"); return 1; }
  }
  std::vector<int> P_col;
  std::vector<double> P_val;
  {
    std::vector<int> cut, hist((size_t)T * n, 0);
    col_cuts(Pc->p, cut);
    run_threads([&](int t) {
      int *h = hist.data() + (size_t)t * n;
      for (int j = cut[t]; j < cut[t + 1]; j++)
        for (c_int k = Pc->p[j]; k < Pc->p[j + 1]; k++) {
          const int i = (int)Pc->i[k];
          h[i]++;
          if (i != j) h[j]++;
        }
    });
    cursors(hist, n, P_rowptr);
    P_col.resize(P_rowptr[n]);
    P_val.resize(P_rowptr[n]);
    run_threads([&](int t) {
      int *h = hist.data() + (size_t)t * n;
      for (int j = cut[t]; j < cut[t + 1]; j++)
        for (c_int k = Pc->p[j]; k < Pc->p[j + 1]; k++) {
          const int i = (int)Pc->i[k];
          int pos = h[i]++;
          P_col[pos] = j;
          P_val[pos] = Pc->x[k];
          mapP1[k] = pos;
          mapP2[k] = -1;
          if (i != j) {
            pos = h[j]++;
            P_col[pos] = i;
            P_val[pos] = Pc->x[k];
            mapP2[k] = pos;
          }
        }
    });
  }
  const long long nnzP = P_rowptr[n];

  mark("CSR index work (host)");
  // ---- launch geometry
  cudaDeviceProp prop;
  CU_OK(cudaGetDeviceProperties(&prop, e.device));
  e.geom.block = std::min(coop_threads(), env_int("OSQP_B200_BLOCK", coop_threads()));
  const int per_sm = max_coop_blocks_per_sm(e.geom.block, 0);
  if (per_sm <= 0) {
    fprintf(stderr, "ERROR in osqp_setup: the sm_100a kernels cannot run on device %d (%s, sm_%d%d)\n", e.device,
            prop.name, prop.major, prop.minor);
    return 12;
  }
  const int max_grid = std::min(kMaxBlocks, prop.multiProcessorCount * std::min(per_sm, env_int("OSQP_B200_BLOCKS_PER_SM", 1)));
  const long long work = 2 * nnzA + nnzP + 4LL * (n + m);
  // measured on B200 (profiles/latency_small.py): one block up to a few thousand units of work, then about one block
  // per 3k units -- the CSR products of small problems are latency-bound and scale with the number of blocks until
  // the grid barrier (1.5 us) takes over
  long long want = (work + 3071) / 3072;
  int grid = (int)std::max(1LL, std::min<long long>(want, max_grid));
  if (env_int("OSQP_B200_GRID", 0) > 0) grid = env_int("OSQP_B200_GRID", grid);
  e.geom.grid = std::max(1, std::min(grid, max_grid));
  d.A.rows = m; d.A.cols = n; d.A.nnz = nnzA;
  d.At.rows = n; d.At.cols = m; d.At.nnz = nnzA;
  d.P.rows = n; d.P.cols = n; d.P.nnz = nnzP;
  d.A.lanes = pow2_lanes(m > 0 ? (double)nnzA / m : 1.0);
  d.At.lanes = d.P.lanes = pow2_lanes((double)(nnzA + nnzP) / n);
  if (d.A.lanes > e.geom.block) d.A.lanes = e.geom.block;

  // ---- device memory
  CU_OK(dalloc(e, &d.A.rowptr, (size_t)m + 1)); CU_OK(dalloc(e, &d.A.col, nnzA));
  CU_OK(dalloc(e, &d.A.val, nnzA)); CU_OK(dalloc(e, &d.A.val0, nnzA));
  CU_OK(dalloc(e, &d.At.rowptr, (size_t)n + 1)); CU_OK(dalloc(e, &d.At.col, nnzA));
  CU_OK(dalloc(e, &d.At.val, nnzA)); CU_OK(dalloc(e, &d.At.val0, nnzA));
  CU_OK(dalloc(e, &d.P.rowptr, (size_t)n + 1)); CU_OK(dalloc(e, &d.P.col, nnzP));
  CU_OK(dalloc(e, &d.P.val, nnzP)); CU_OK(dalloc(e, &d.P.val0, nnzP));
  double **nvecs[] = {&d.q, &d.q0, &d.Pdiag, &d.D, &d.Dinv, &d.dtmp, &d.x, &d.xt, &d.dx, &d.r, &d.b, &d.uu,
                      &d.p, &d.s, &d.w, &d.Minv, &d.pol_x, &d.pol_rhs, &d.sol_x};
  for (double **p : nvecs) CU_OK(dalloc(e, p, (size_t)n + 8));  // +8: TMA tile copies round up to 16 B
  double **mvecs[] = {&d.l, &d.u, &d.l0, &d.u0, &d.rho_vec, &d.rho_inv, &d.E, &d.Einv, &d.etmp, &d.z, &d.y, &d.zt,
                      &d.dy, &d.wv, &d.t, &d.tr, &d.Ap, &d.pol_y, &d.pol_z, &d.pol_rho, &d.pol_b, &d.sol_y};
  for (double **p : mvecs) CU_OK(dalloc(e, p, (size_t)m + 8));
  CU_OK(dalloc(e, &d.ctype, m));
  CU_OK(dalloc(e, &d.m_start, e.geom.grid + 1));
  CU_OK(dalloc(e, &d.n_start, e.geom.grid + 1));
  CU_OK(dalloc(e, &d.bar, kBarBytes / sizeof(unsigned)));
  CU_OK(dalloc(e, &d.red, std::max<size_t>((size_t)2 * kRedSlots * e.geom.grid, 512)));  // >= 2 x 148: k_ruiz_cost_partial
  CU_OK(dalloc(e, &d.dbg, (size_t)16 * e.geom.grid));
  CU_OK(dalloc(e, &d.state, 1));
  CU_OK(dalloc(e, &d.info, 1));
  CU_OK(dalloc(e, &e.d_pol, 1));
  CU_OK(dalloc(e, &e.mapA, nnzA));
  CU_OK(dalloc(e, &e.mapP1, nnzPt));
  CU_OK(dalloc(e, &e.mapP2, nnzPt));
  CU_OK(halloc(e, &e.h_sol_x, n)); CU_OK(halloc(e, &e.h_sol_y, m));
  CU_OK(halloc(e, &e.h_dx, n)); CU_OK(halloc(e, &e.h_dy, m));
  CU_OK(halloc(e, &e.h_info, 1)); CU_OK(halloc(e, &e.h_pol, 1)); CU_OK(halloc(e, &e.h_state, 1));

  // ---- uploads (the caller's buffers may be freed right after we return: everything is copied now)
#define H2D(dst, src, count, T) \
  if ((count) > 0) CU_OK(cudaMemcpyAsync(dst, src, (size_t)(count) * sizeof(T), cudaMemcpyHostToDevice, e.stream))
  H2D(d.A.rowptr, A_rowptr.data(), m + 1, int); H2D(d.A.col, A_col.data(), nnzA, int);
  H2D(d.A.val0, A_val.data(), nnzA, double);
  H2D(d.At.rowptr, At_rowptr.data(), n + 1, int); H2D(d.At.col, At_col.data(), nnzA, int);
  H2D(d.At.val0, Ac->x, nnzA, double);
  H2D(d.P.rowptr, P_rowptr.data(), n + 1, int); H2D(d.P.col, P_col.data(), nnzP, int);
  H2D(d.P.val0, P_val.data(), nnzP, double);
  H2D(d.q0, data->q, n, double); H2D(d.l0, data->l, m, double); H2D(d.u0, data->u, m, double);
  H2D(e.mapA, mapA.data(), nnzA, int); H2D(e.mapP1, mapP1.data(), nnzPt, int); H2D(e.mapP2, mapP2.data(), nnzPt, int);
#undef H2D
  e.l0.assign(data->l, data->l + m);
  e.u0.assign(data->u, data->u + m);
  std::vector<int> part_m, part_n;
  {
    c_int rc = upload_partition(e, A_rowptr, At_rowptr, P_rowptr, part_m, part_n);
    if (rc) return rc;
  }

  mark("device alloc + CSR upload");
  // ---- tile streams for the hot phases (engine.cuh TileStreamDev)
  {
    d.blocked = 0;
    const long long smem_budget = (long long)prop.sharedMemPerBlockOptin - (long long)coop_static_smem() - 384;
    const int slice_cap = std::min<long long>(std::min(kSliceMax, env_int("OSQP_B200_SLICE", kSliceMax)),
                                              (smem_budget - 128) / 8);
    auto groups_for = [&](int cols) {
      int gmin = std::max(1, (cols + slice_cap - 1) / slice_cap);
      return std::max(gmin, env_int("OSQP_B200_GROUPS", 0));
    };
    const bool want = env_int("OSQP_B200_BLOCKED", 1) != 0 && e.geom.grid >= 8 &&
                      (nnzA + nnzP) >= env_int("OSQP_B200_STREAM_MIN_NNZ", 200000);
    if (want) {
      TileStreamHost hA, hT;
      std::vector<CsrRef> matsA;
      if (m > 0) matsA.push_back(CsrRef{&A_rowptr, &A_col, m});
      matsA.push_back(CsrRef{&P_rowptr, &P_col, n});
      std::vector<CsrRef> matsT{CsrRef{&At_rowptr, &At_col, n}};
      bool want_pairs = env_int("OSQP_B200_PAIRS", 1) != 0 && groups_for(n) == 2 && (e.geom.grid & 1) == 0;
      if (want_pairs) {  // every block of the persistent grid must stay co-resident when launched as clusters of two
        CU_OK(configure_dyn_smem((size_t)smem_budget));
        const int nc = max_active_clusters(2, e.geom.block, (size_t)smem_budget);
        want_pairs = nc * 2 >= e.geom.grid;
        if (env_int("OSQP_B200_DEBUG", 0)) fprintf(stderr, "[osqp_b200] clusters of 2 co-resident: %d (grid %d)\n", nc, e.geom.grid);
      }
      bool ok = false;
      for (int attempt = want_pairs ? 0 : 1; attempt < 2 && !ok; attempt++) {
        const bool paired = attempt == 0;
        // a pair keeps the row sums of its range in shared memory next to the slice: cap the rows per pair
        const int Wg = (((n + 1) / 2) + 31) & ~31;
        const long long room = (smem_budget - 64 - 8LL * (((Wg + 2 + 15) & ~15))) / 8 - 16;
        if (paired && room < 64) continue;
        const bool lane_rows = env_int("OSQP_B200_LANE_ROWS", 1) != 0;
        ok = build_tile_stream(matsA, n, e.geom.grid, groups_for(n), paired, hA, (int)std::min<long long>(room, 1 << 30), lane_rows);
        if (ok && m > 0 && hT.nelem == 0) ok = build_tile_stream(matsT, m, e.geom.grid, groups_for(m), false, hT, 1 << 30, lane_rows);
        if (!ok) break;
        const int slice = std::max(hA.max_slice, m > 0 ? hT.max_slice : 0);
        d.smem_x_elems = (slice + 2 + 15) & ~15;
        d.smem_rows = hA.paired ? ((hA.max_block_rows + 15) & ~15) : 0;
        e.geom.dyn_smem = 8ULL * ((size_t)d.smem_x_elems + d.smem_rows) + 64;
        ok = (long long)e.geom.dyn_smem <= smem_budget;  // pairs need room for the row accumulators: else retry without
      }
      e.geom.cluster = (ok && hA.paired) ? 2 : 1;
      e.geom.fast = env_int("OSQP_B200_FAST_KERNELS", 1) != 0;
      if (env_int("OSQP_B200_DEBUG", 0))
        fprintf(stderr, "[osqp_b200] tile streams ok=%d paired=%d groups=%d/%d slice=%d rows=%d dyn_smem=%zu budget=%lld\n",
                (int)ok, hA.paired, hA.ngroups, hT.ngroups, d.smem_x_elems, d.smem_rows, e.geom.dyn_smem, smem_budget);
      if (ok) {
        { c_int rc = upload_tile_stream(e, hA, d.SA); if (rc) return rc; }
        if (m > 0) { c_int rc = upload_tile_stream(e, hT, d.ST); if (rc) return rc; }
        CU_OK(dalloc(e, &d.Pu, (size_t)n + 8));
        d.f32_slices = env_int("OSQP_B200_F32_SLICES", 1) != 0;
        if (d.f32_slices) {
          CU_OK(dalloc(e, &d.uu32, (size_t)n + 32));
          CU_OK(dalloc(e, &d.tr32, (size_t)m + 32));
        }
        CU_OK(configure_dyn_smem(e.geom.dyn_smem));
        if (max_coop_blocks_per_sm(e.geom.block, e.geom.dyn_smem) < 1) {
          fprintf(stderr, "ERROR in osqp_setup: shared-memory slice of %zu bytes does not fit\n", e.geom.dyn_smem);
          return 13;
        }
        d.blocked = 1;
        d.info_streams = env_int("OSQP_B200_INFO_STREAMS", 1) != 0;
      } else {
        e.geom.dyn_smem = 0;
      }
    }
  }

  // ---- low-rank (Woodbury) part of the preconditioner (engine.cuh WoodDev): the few coupling equality rows.
  // Membership is fixed here from the bounds as passed; it is a choice of preconditioner, not of algorithm -- a row
  // that later stops (or starts) being an equality only costs PCG iterations.
  if (d.blocked && m > 0 && env_int("OSQP_B200_WOODBURY", 1) != 0) {
    std::vector<int> rows;
    long long nnzW = 0;
    for (int i = 0; i < m && (int)rows.size() <= kWoodMax; i++)
      if (data->u[i] - data->l[i] < 1e-4 && A_rowptr[i + 1] - A_rowptr[i] >= 2) {
        rows.push_back(i);
        nnzW += A_rowptr[i + 1] - A_rowptr[i];
      }
    const int w = (int)rows.size();
    if (w > 0 && w <= kWoodMax) {
      WoodDev &W = d.W;
      std::vector<int> idx(m, -1), rp(w + 1, 0), ci(nnzW), src(nnzW), trp(n + 1, 0), tci(nnzW), tsrc(nnzW);
      for (int a = 0; a < w; a++) {
        idx[rows[a]] = a;
        rp[a + 1] = rp[a] + (A_rowptr[rows[a] + 1] - A_rowptr[rows[a]]);
        for (int k = A_rowptr[rows[a]], o = rp[a]; k < A_rowptr[rows[a] + 1]; k++, o++) { ci[o] = A_col[k]; src[o] = k; }
      }
      {
        int o = 0;
        for (int j = 0; j < n; j++) {
          for (int k = At_rowptr[j]; k < At_rowptr[j + 1]; k++)
            if (idx[At_col[k]] >= 0) { tci[o] = idx[At_col[k]]; tsrc[o] = mapA[k]; o++; }
          trp[j + 1] = o;
        }
      }
      W.ld = (w + 7) & ~7;
#define UPW(dst, vec)                                                                                        \
  CU_OK(dalloc(e, &dst, (vec).size()));                                                                      \
  CU_OK(cudaMemcpyAsync(dst, (vec).data(), (vec).size() * sizeof((vec)[0]), cudaMemcpyHostToDevice, e.stream))
      UPW(W.rows, rows); UPW(W.idx, idx); UPW(W.rp, rp); UPW(W.ci, ci); UPW(W.src, src);
      UPW(W.trp, trp); UPW(W.tci, tci); UPW(W.tsrc, tsrc);
#undef UPW
      CU_OK(dalloc(e, &W.val, (size_t)nnzW)); CU_OK(dalloc(e, &W.tval, (size_t)nnzW));
      CU_OK(dalloc(e, &W.C, (size_t)W.ld * W.ld)); CU_OK(dalloc(e, &W.Cinv, (size_t)W.ld * W.ld));
      CU_OK(dalloc(e, &W.s, (size_t)W.ld)); CU_OK(dalloc(e, &W.t, (size_t)W.ld)); CU_OK(dalloc(e, &W.g, (size_t)W.ld));
      CU_OK(dalloc(e, &W.v, (size_t)n + 8)); CU_OK(dalloc(e, &W.vk, (size_t)kWoodCols * ((size_t)n + 8)));
      CU_OK(cudaStreamSynchronize(e.stream));  // the host vectors go out of scope
      W.w = w;
      if (env_int("OSQP_B200_DEBUG", 0))
        fprintf(stderr, "[osqp_b200] Woodbury preconditioner: %d coupling equality rows, %lld non-zeros\n", w, nnzW);
    }
  }

  // ---- slack elimination in the preconditioner (engine.cuh SlackDev): equality rows with a private slack column,
  // when they are too many for the Woodbury correction
  if (d.blocked && m > 0 && d.W.w == 0 && env_int("OSQP_B200_SLACK", 1) != 0) {
    std::vector<int> scol(m, -1), spos(m, -1);
    int cnt = 0;
    for (int j = 0; j < n; j++) {
      if (At_rowptr[j + 1] - At_rowptr[j] != 1) continue;
      bool pdiag = true;
      for (int k = P_rowptr[j]; k < P_rowptr[j + 1] && pdiag; k++) pdiag = P_col[k] == j;
      if (!pdiag) continue;
      const int i = At_col[At_rowptr[j]];
      if (data->u[i] - data->l[i] < 1e-4 && scol[i] < 0 && A_rowptr[i + 1] - A_rowptr[i] >= 2) {
        scol[i] = j;
        spos[i] = mapA[At_rowptr[j]];
        cnt++;
      }
    }
    if (cnt >= 256) {
      SlackDev &S = d.SL;
      CU_OK(dalloc(e, &S.col, (size_t)m)); CU_OK(dalloc(e, &S.pos, (size_t)m));
      CU_OK(dalloc(e, &S.rho_eff, (size_t)m + 8)); CU_OK(dalloc(e, &S.g32, (size_t)m + 32));
      CU_OK(cudaMemcpyAsync(S.col, scol.data(), (size_t)m * sizeof(int), cudaMemcpyHostToDevice, e.stream));
      CU_OK(cudaMemcpyAsync(S.pos, spos.data(), (size_t)m * sizeof(int), cudaMemcpyHostToDevice, e.stream));
      CU_OK(cudaStreamSynchronize(e.stream));
      S.rows = cnt;
      if (!d.f32_slices) {  // the extra phase gathers an fp32 vector
        d.f32_slices = 1;
        CU_OK(dalloc(e, &d.uu32, (size_t)n + 32));
        CU_OK(dalloc(e, &d.tr32, (size_t)m + 32));
      }
      if (env_int("OSQP_B200_DEBUG", 0))
        fprintf(stderr, "[osqp_b200] slack elimination in the preconditioner: %d equality rows with a private column\n", cnt);
    }
  }

  // fp32 copies of the stream values for the PCG phases of the fixed-mode kernels (engine.cuh DevPtrs::mat32)
  d.mat32 = 0;
  if (env_int("OSQP_B200_MAT32", 1) != 0 && fits_fp32_copy(data->P->x, data->P->p[data->n]) &&
      fits_fp32_copy(data->A->x, data->A->p[data->n])) {
    d.mat32 = 1;  // provisional: fast_mode() looks at it
    if (fast_mode(d, e.geom)) {
      CU_OK(dalloc(e, &d.SA.val32, (size_t)d.SA.nelem + 8));
      if (m > 0) CU_OK(dalloc(e, &d.ST.val32, (size_t)d.ST.nelem + 8));
    } else {
      d.mat32 = 0;
    }
  }

  mark("tile streams (host build + upload)");
  // ---- state, scaling (a2), rho vector (a3), preconditioner, convexity probe
  e.st.rho = std::min(std::max(e.st.rho, kRhoMin), kRhoMax);
  memset(e.h_state, 0, sizeof(DevState));
  e.h_state->rho = e.st.rho;
  e.h_state->c = 1.0;
  e.h_state->cinv = 1.0;
  e.h_state->adaptive_interval = e.st.adaptive_rho_interval;
  e.h_state->needs_refresh = 1;
  { c_int rc = push_state(e); if (rc) return rc; }
  { c_int rc = rescale_and_refresh(e, true); if (rc) return rc; }
  {
    int nonconvex = 0;
    c_int rc = convexity_check(e, &nonconvex);
    if (rc) return rc;
    if (nonconvex) {
      fprintf(stderr, "ERROR in osqp_setup: P + sigma*I is not positive definite (the problem seems to be non-convex)\n");
      return 7;
    }
  }

  mark("scaling, rho, precond, probe");
  memset(&e.info, 0, sizeof(e.info));
  update_status(e.info, OSQP_UNSOLVED);
  e.info.rho_estimate = e.st.rho;
  e.first_run = true;
  e.prof.device = e.device;
  e.prof.grid = e.geom.grid;
  e.prof.block = e.geom.block;
  e.prof.lanes_A = d.A.lanes;
  e.prof.lanes_N = d.At.lanes;
  e.prof.nnz_A = nnzA;
  e.prof.nnz_P_full = nnzP;
  e.prof.spmv_bytes_A = spmv_bytes(nnzA, m, n);
  e.prof.spmv_bytes_At = spmv_bytes(nnzA, n, m);
  e.prof.spmv_bytes_P = spmv_bytes(nnzP, n, n);
  e.prof.streams = d.blocked;
  e.prof.groups_A = d.blocked ? d.SA.ngroups : 0;
  e.prof.groups_At = (d.blocked && m > 0) ? d.ST.ngroups : 0;
  e.prof.paired = d.blocked ? d.SA.paired : 0;
  // ---- tiny mode: the same QP in the batched engine (a batch of one)
  if (n <= kTinyMax && m <= kTinyMax && env_int("OSQP_B200_TINY", 1) != 0) {
    e.tP_p.assign(Pc->p, Pc->p + n + 1); e.tP_i.assign(Pc->i, Pc->i + nnzPt); e.tP_x.assign(Pc->x, Pc->x + nnzPt);
    e.tA_p.assign(Ac->p, Ac->p + n + 1); e.tA_i.assign(Ac->i, Ac->i + nnzA); e.tA_x.assign(Ac->x, Ac->x + nnzA);
    e.tq.assign(data->q, data->q + n);
    tiny_setup(e, e.st.rho);
  }
  publish(e);
  e.info.setup_time = now_s() - t0;
  if (e.st.verbose) print_setup_header(e);
  fail.p = nullptr;
  *workp = &e.pub;
  return 0;
}

static c_int solve_impl(Engine &e);
static c_int warm(Engine &e, const c_float *x, const c_float *y);

// The reference ignores osqp_solve's return value (src/interface.jl:170-175) and reads info / solution straight from
// the workspace.  libosqp cannot fail between a successful setup and the end of a solve; a GPU engine can (launch
// refused, sticky CUDA error), so every error exit leaves an unmistakably unsolved workspace behind: status
// Unsolved, iter 0, NaN solution -- never the results of the previous solve.
c_int osqp_solve(OSQPWorkspace *work) {
  if (!work) { fprintf(stderr, "ERROR in osqp_solve: workspace not initialized\n"); return 1; }
  Engine &e = *E(work);
  DeviceGuard guard(e.device);
  c_int rc;
  if (use_tiny(e)) {
    rc = tiny_solve(e);
  } else {
    if (e.last_path == 2 && e.st.warm_start && has_solution(e.info.status_val))  // the tiny engine solved last: hand over
      warm(e, e.h_sol_x, e.d.m > 0 ? e.h_sol_y : nullptr);
    rc = solve_impl(e);
    e.last_path = 1;
  }
  if (rc != 0) {
    update_status(e.info, OSQP_UNSOLVED);
    e.info.iter = 0;
    e.info.status_polish = 0;
    e.info.obj_val = e.info.pri_res = e.info.dua_res = NAN;
    for (int j = 0; j < e.d.n; j++) e.h_sol_x[j] = NAN;
    for (int i = 0; i < e.d.m; i++) e.h_sol_y[i] = NAN;
    e.h_state->needs_refresh = 1;
  }
  return rc;
}

static c_int solve_impl(Engine &e) {
  if (e.clear_update_time) e.info.update_time = 0.0;
  const double t0 = now_s();
  SolveCfg c;
  memset(&c, 0, sizeof(c));
  c.sigma = e.st.sigma; c.alpha = e.st.alpha;
  c.eps_abs = e.st.eps_abs; c.eps_rel = e.st.eps_rel;
  c.eps_prim_inf = e.st.eps_prim_inf; c.eps_dual_inf = e.st.eps_dual_inf;
  c.max_iter = e.st.max_iter; c.check_termination = e.st.check_termination;
  c.scaling = e.st.scaling != 0; c.scaled_termination = (int)e.st.scaled_termination;
  c.adaptive_rho = (int)e.st.adaptive_rho; c.adaptive_rho_tolerance = e.st.adaptive_rho_tolerance;
  c.adaptive_time_s = e.st.adaptive_rho_fraction * e.info.setup_time;
  c.warm_start = (int)e.st.warm_start; c.verbose = (int)e.st.verbose;
  const double base = e.first_run ? e.info.setup_time : e.info.update_time;
  c.time_limit_s = e.st.time_limit > 0 ? e.st.time_limit - base : -1e30;
  c.pcg_eta = e.pcg_eta; c.pcg_floor = e.pcg_floor; c.pcg_eta_e = e.pcg_eta_e; c.pcg_max_iter = e.pcg_max_iter;
  c.refresh_every = e.refresh_every;
  c.wood_refresh = ((e.d.W.w > 0 || e.d.SL.rows > 0) && e.wood_dirty) ? 1 : 0;
  if (e.st.verbose) printf("iter   objective    pri res    dua res    rho        time\n");
  // Automatic adaptive-rho interval (settings.adaptive_rho_interval = 0).  libosqp derives it from wall-clock:
  // the first iteration after 0.4 * setup_time, rounded to a multiple of check_termination -- a trade between the
  // cost of a refactorisation and of an iteration.  Here a rho update costs two stream phases, and setup_time is
  // host index work + CUDA initialisation, unrelated to either: the rule would fire at a different iteration in
  // every process.  The engine therefore fixes the interval deterministically at the multiple of check_termination
  // nearest to 50 (the batched engine uses the same number); OSQP_B200_ADAPTIVE_WALLCLOCK=1 restores libosqp's rule.
  if (e.st.adaptive_rho && e.h_state->adaptive_interval == 0 && !env_int("OSQP_B200_ADAPTIVE_WALLCLOCK", 0)) {
    const long long N = e.st.check_termination > 0 ? e.st.check_termination : 25;
    long long iv = ((50 + N / 2) / N) * N;
    if (iv < N) iv = N;
    if (e.st.adaptive_rho_interval > 0) iv = e.st.adaptive_rho_interval;  // already fixed by a tiny-mode solve
    e.h_state->adaptive_interval = iv;
    e.st.adaptive_rho_interval = iv;
    { c_int rc = push_state(e); if (rc) return rc; }
  }

  CU_OK(cudaEventRecord(e.ev0, e.stream));
  CU_OK(launch_with_pair_fallback(e, [&]() { return launch_solve(e.d, c, e.geom, e.stream); }));
  e.prof.fast_kernels = fast_mode(e.d, e.geom);
  CU_OK(cudaEventRecord(e.ev1, e.stream));
  e.prof.launches += 1;
  e.wood_dirty = false;  // the launch leaves the Woodbury data consistent with the rho it ends on
  const int n = e.d.n, m = e.d.m;
  CU_OK(cudaMemcpyAsync(e.h_info, e.d.info, sizeof(DevInfo), cudaMemcpyDeviceToHost, e.stream));
  CU_OK(cudaMemcpyAsync(e.h_sol_x, e.d.sol_x, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  if (m > 0) CU_OK(cudaMemcpyAsync(e.h_sol_y, e.d.sol_y, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  CU_OK(cudaStreamSynchronize(e.stream));
  const DevInfo &I = *e.h_info;
  e.info.iter = I.iter;
  update_status(e.info, I.status_val);
  e.info.obj_val = I.obj_val;
  e.info.pri_res = I.pri_res;
  e.info.dua_res = I.dua_res;
  e.info.rho_updates = I.rho_updates;
  e.info.rho_estimate = I.rho_estimate;
  e.st.rho = I.rho;
  e.st.adaptive_rho_interval = I.adaptive_interval;
  e.h_state->rho = I.rho;
  e.h_state->rho_updates = I.rho_updates;
  e.h_state->adaptive_interval = I.adaptive_interval;
  const c_int sv = I.status_val;
  if (sv == OSQP_PRIMAL_INFEASIBLE || sv == OSQP_PRIMAL_INFEASIBLE_INACCURATE) {
    if (m > 0) CU_OK(cudaMemcpyAsync(e.h_dy, e.d.dy, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
    CU_OK(cudaStreamSynchronize(e.stream));
  } else if (sv == OSQP_DUAL_INFEASIBLE || sv == OSQP_DUAL_INFEASIBLE_INACCURATE) {
    CU_OK(cudaMemcpyAsync(e.h_dx, e.d.dx, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
    CU_OK(cudaStreamSynchronize(e.stream));
  }
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e.ev0, e.ev1);
  e.prof.kernel_ms = ms;
  e.prof.polish_ms = 0;
  e.prof.admm_iters = I.cg_solves;
  e.prof.pcg_iters = I.cg_iters;
  e.prof.info_evals = I.checks;
  e.prof.refreshes = I.refreshes;
  for (int k = 0; k < kPhases; k++) e.prof.phase_us[k] = I.phase_us[k];
  {
    // algorithmic bytes of this launch (DESIGN.md "bytes model"): matrix streams + dense vector passes
    const double bA = e.prof.spmv_bytes_A, bAt = e.prof.spmv_bytes_At, bP = e.prof.spmv_bytes_P;
    const double k = (double)I.cg_iters, it = (double)I.cg_solves, ck = (double)I.checks, rf = (double)I.refreshes;
    const double vn = 8.0 * n, vm = 8.0 * m;
    e.prof.alg_bytes = k * (bA + bAt + bP + 11 * vn + 5 * vm) + it * (bAt + 6 * vn + 12 * vm + 4 * vn) +
                       rf * (bA + bP) + ck * (bA + bAt + bP);
    if (e.d.SL.rows > 0) e.prof.alg_bytes += k * (bAt + 2 * vn + 2 * vm);  // the extra A' phase of the slack preconditioner
  }
  if (e.st.verbose) {
    for (long long r = 0; r < I.log_rows; r++)
      printf("%4lld  %11.4e  %9.2e  %9.2e  %9.2e  %9.2es\n", (long long)I.log[r][0], I.log[r][1], I.log[r][2],
             I.log[r][3], I.log[r][4], I.log[r][5]);
    if (I.log_rows == 0 || (long long)I.log[I.log_rows - 1][0] != I.iter)
      printf("%4lld  %11.4e  %9.2e  %9.2e  %9.2e  %9.2es\n", (long long)I.iter, I.obj_val, I.pri_res, I.dua_res,
             I.rho, I.elapsed_s);
  }
  e.info.solve_time = now_s() - t0;

  // ---- polish (a12)
  if (e.st.polish && sv == OSQP_SOLVED) {
    const double tp = now_s();
    PolishCfg pc;
    memset(&pc, 0, sizeof(pc));
    pc.delta = e.st.delta;
    pc.penalty = e.polish_penalty;
    pc.refine_iter = 1 + 2 * (int)e.st.polish_refine_iter;
    pc.pcg_rel_tol = 1e-13;
    pc.pcg_max_iter = std::max(50, std::min(20 * n, 20000));
    pc.scaling = c.scaling;
    pc.scaled_termination = c.scaled_termination;
    CU_OK(launch_with_pair_fallback(e, [&]() { return launch_polish(e.d, pc, c, e.d_pol, e.geom, e.stream); }));
    CU_OK(cudaEventRecord(e.ev2, e.stream));
    e.wood_dirty = true;  // polish rebuilt the Woodbury data for its own penalty vector
    e.prof.launches += 1;
    CU_OK(cudaMemcpyAsync(e.h_pol, e.d_pol, sizeof(PolishOut), cudaMemcpyDeviceToHost, e.stream));
    CU_OK(cudaMemcpyAsync(e.h_sol_x, e.d.sol_x, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
    if (m > 0) CU_OK(cudaMemcpyAsync(e.h_sol_y, e.d.sol_y, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
    CU_OK(cudaStreamSynchronize(e.stream));
    cudaEventElapsedTime(&ms, e.ev1, e.ev2);
    e.prof.polish_ms = ms;
    if (e.h_pol->success) {
      e.info.obj_val = e.h_pol->obj_val;
      e.info.pri_res = e.h_pol->pri_res;
      e.info.dua_res = e.h_pol->dua_res;
      e.info.status_polish = 1;
    } else {
      e.info.status_polish = -1;
    }
    e.info.polish_time = now_s() - tp;
    if (e.st.verbose && e.h_pol->success)
      printf("plsh  %11.4e  %9.2e  %9.2e   --------  %9.2es\n", e.info.obj_val, e.info.pri_res, e.info.dua_res,
             e.info.polish_time);
  }
  e.info.run_time = base + e.info.solve_time + e.info.polish_time;
  e.first_run = false;
  e.clear_update_time = true;
  e.pub.first_run = 0;
  if (e.st.verbose) {
    printf("\nstatus:               %s\n", e.info.status);
    if (e.st.polish && sv == OSQP_SOLVED)
      printf("solution polish:      %s\n", e.info.status_polish == 1 ? "successful" : "unsuccessful");
    printf("number of iterations: %lld\n", (long long)e.info.iter);
    if (sv == OSQP_SOLVED || sv == OSQP_SOLVED_INACCURATE) printf("optimal objective:    %.4f\n", e.info.obj_val);
    printf("run time:             %.2es\n", e.info.run_time);
    printf("optimal rho estimate: %.2e\n\n", e.info.rho_estimate);
  }
  return 0;
}

// ---- updates (a13)
c_int osqp_update_lin_cost(OSQPWorkspace *work, const c_float *q_new) {
  if (!work) return 1;
  Engine &e = *E(work);
  DeviceGuard guard(e.device);
  begin_update(e);
  const double t0 = now_s();
  { c_int rc = upload_vector(e, e.d.q0, q_new, e.d.n); if (rc) return rc; }
  CU_OK(launch_scale_vectors(e.d, 1, 0, e.stream));
  e.prof.launches += 1;
  CU_OK(cudaStreamSynchronize(e.stream));  // q_new is caller-owned
  reset_info(e);
  { c_int rc = push_state(e); if (rc) return rc; }
  if (e.tiny) {
    e.tq.assign(q_new, q_new + e.d.n);
    osqp_batch_update(e.tiny, q_new, nullptr, nullptr);
  }
  e.info.update_time += now_s() - t0;
  return 0;
}

static c_int bounds_changed(Engine &e) {
  CU_OK(launch_scale_vectors(e.d, 0, 1, e.stream));
  CU_OK(launch_set_rho_vec(e.d, e.st.rho, 1, e.stream));
  CU_OK(launch_precond(e.d, e.st.sigma, e.stream));
  e.wood_dirty = true;
  e.prof.launches += 3;
  CU_OK(cudaStreamSynchronize(e.stream));
  reset_info(e);
  e.h_state->needs_refresh = 1;
  if (e.tiny) osqp_batch_update(e.tiny, nullptr, e.l0.data(), e.u0.data());
  return push_state(e);
}

c_int osqp_update_bounds(OSQPWorkspace *work, const c_float *l_new, const c_float *u_new) {
  if (!work) return 1;
  Engine &e = *E(work);
  DeviceGuard guard(e.device);
  begin_update(e);
  const double t0 = now_s();
  const int m = e.d.m;
  for (int i = 0; i < m; i++)
    if (l_new[i] > u_new[i]) {
      fprintf(stderr, "ERROR in osqp_update_bounds: lower bound must be lower than or equal to upper bound\n");
      return 1;
    }
  e.l0.assign(l_new, l_new + m);
  e.u0.assign(u_new, u_new + m);
  { c_int rc = upload_vector(e, e.d.l0, l_new, m); if (rc) return rc; }
  { c_int rc = upload_vector(e, e.d.u0, u_new, m); if (rc) return rc; }
  c_int rc = bounds_changed(e);
  e.info.update_time += now_s() - t0;
  return rc;
}
c_int osqp_update_lower_bound(OSQPWorkspace *work, const c_float *l_new) {
  if (!work) return 1;
  Engine &e = *E(work);
  DeviceGuard guard(e.device);
  begin_update(e);
  const double t0 = now_s();
  const int m = e.d.m;
  for (int i = 0; i < m; i++)
    if (l_new[i] > e.u0[i]) {
      fprintf(stderr, "ERROR in osqp_update_lower_bound: upper bound must be greater than or equal to lower bound\n");
      return 1;
    }
  e.l0.assign(l_new, l_new + m);
  { c_int rc = upload_vector(e, e.d.l0, l_new, m); if (rc) return rc; }
  c_int rc = bounds_changed(e);
  e.info.update_time += now_s() - t0;
  return rc;
}
c_int osqp_update_upper_bound(OSQPWorkspace *work, const c_float *u_new) {
  if (!work) return 1;
  Engine &e = *E(work);
  DeviceGuard guard(e.device);
  begin_update(e);
  const double t0 = now_s();
  const int m = e.d.m;
  for (int i = 0; i < m; i++)
    if (e.l0[i] > u_new[i]) {
      fprintf(stderr, "ERROR in osqp_update_upper_bound: upper bound must be greater than or equal to lower bound\n");
      return 1;
    }
  e.u0.assign(u_new, u_new + m);
  { c_int rc = upload_vector(e, e.d.u0, u_new, m); if (rc) return rc; }
  c_int rc = bounds_changed(e);
  e.info.update_time += now_s() - t0;
  return rc;
}

static c_int stage_values(Engine &e, const c_float *vals, const c_int *idx, long long k) {
  if (k > e.stage_cap) {
    long long cap = std::max<long long>(k, 1024);
    CU_OK(dalloc(e, &e.stage_val, cap));
    CU_OK(dalloc(e, &e.stage_idx, cap));
    e.stage_cap = cap;
  }
  CU_OK(cudaMemcpyAsync(e.stage_val, vals, k * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  if (idx) CU_OK(cudaMemcpyAsync(e.stage_idx, idx, k * sizeof(long long), cudaMemcpyHostToDevice, e.stream));
  return 0;
}

static c_int update_PA(Engine &e, const c_float *Px_new, const c_int *Px_idx, c_int P_n, bool doP,
                       const c_float *Ax_new, const c_int *Ax_idx, c_int A_n, bool doA) {
  DeviceGuard guard(e.device);
  begin_update(e);
  const double t0 = now_s();
  if (doP && Px_idx && P_n > e.nnzPtriu) {
    fprintf(stderr, "ERROR in osqp_update_P: new number of elements (%lld) greater than elements in P (%lld)\n",
            (long long)P_n, (long long)e.nnzPtriu);
    return 1;
  }
  if (doA && Ax_idx && A_n > e.nnzA) {
    fprintf(stderr, "ERROR in osqp_update_A: new number of elements (%lld) greater than elements in A (%lld)\n",
            (long long)A_n, (long long)e.nnzA);
    return doP ? 2 : 1;
  }
  if (e.d.mat32 && ((doP && !fits_fp32_copy(Px_new, Px_idx ? P_n : e.nnzPtriu)) ||
                    (doA && !fits_fp32_copy(Ax_new, Ax_idx ? A_n : e.nnzA))))
    e.d.mat32 = 0;  // from now on the plain kernels (kernels.cu fast_mode)
  if (doP) {
    const long long k = Px_idx ? P_n : e.nnzPtriu;
    if (Px_idx)
      for (long long t = 0; t < k; t++)
        if (Px_idx[t] < 0 || Px_idx[t] >= e.nnzPtriu) { fprintf(stderr, "ERROR in osqp_update_P: index out of range\n"); return 1; }
    { c_int rc = stage_values(e, Px_new, Px_idx, k); if (rc) return rc; }
    CU_OK(launch_scatter_values(e.d.P.val0, e.stage_val, Px_idx ? e.stage_idx : nullptr, e.mapP1, k, e.stream));
    CU_OK(launch_scatter_values(e.d.P.val0, e.stage_val, Px_idx ? e.stage_idx : nullptr, e.mapP2, k, e.stream));
    e.prof.launches += 2;
    CU_OK(cudaStreamSynchronize(e.stream));
  }
  if (doA) {
    const long long k = Ax_idx ? A_n : e.nnzA;
    if (Ax_idx)
      for (long long t = 0; t < k; t++)
        if (Ax_idx[t] < 0 || Ax_idx[t] >= e.nnzA) { fprintf(stderr, "ERROR in osqp_update_A: index out of range\n"); return doP ? 2 : 1; }
    { c_int rc = stage_values(e, Ax_new, Ax_idx, k); if (rc) return rc; }
    CU_OK(launch_scatter_values(e.d.At.val0, e.stage_val, Ax_idx ? e.stage_idx : nullptr, nullptr, k, e.stream));
    CU_OK(launch_scatter_values(e.d.A.val0, e.stage_val, Ax_idx ? e.stage_idx : nullptr, e.mapA, k, e.stream));
    e.prof.launches += 2;
    CU_OK(cudaStreamSynchronize(e.stream));
  }
  // unscale -> overwrite -> scale of libosqp == re-equilibrate the stored originals
  { c_int rc = rescale_and_refresh(e, false); if (rc) return rc; }
  int failed = 0;
  { c_int rc = convexity_check(e, &failed); if (rc) return rc; }
  reset_info(e);
  e.h_state->needs_refresh = 1;
  { c_int rc = push_state(e); if (rc) return rc; }
  if (e.tiny) {  // same values into the host copies the tiny engine is rebuilt from
    if (doP) for (long long t = 0, k = Px_idx ? P_n : e.nnzPtriu; t < k; t++) e.tP_x[Px_idx ? Px_idx[t] : t] = Px_new[t];
    if (doA) for (long long t = 0, k = Ax_idx ? A_n : e.nnzA; t < k; t++) e.tA_x[Ax_idx ? Ax_idx[t] : t] = Ax_new[t];
    if (failed) { osqp_batch_cleanup(e.tiny); e.tiny = nullptr; }
    else tiny_rebuild(e);
  }
  e.info.update_time += now_s() - t0;
  if (failed) {
    fprintf(stderr, "ERROR in osqp_update_P/A: new KKT matrix is not quasidefinite\n");
    return -2;
  }
  return 0;
}
c_int osqp_update_P(OSQPWorkspace *work, const c_float *Px_new, const c_int *Px_new_idx, c_int P_new_n) {
  if (!work) return 1;
  return update_PA(*E(work), Px_new, Px_new_idx, P_new_n, true, nullptr, nullptr, 0, false);
}
c_int osqp_update_A(OSQPWorkspace *work, const c_float *Ax_new, const c_int *Ax_new_idx, c_int A_new_n) {
  if (!work) return 1;
  return update_PA(*E(work), nullptr, nullptr, 0, false, Ax_new, Ax_new_idx, A_new_n, true);
}
c_int osqp_update_P_A(OSQPWorkspace *work, const c_float *Px_new, const c_int *Px_new_idx, c_int P_new_n,
                      const c_float *Ax_new, const c_int *Ax_new_idx, c_int A_new_n) {
  if (!work) return 1;
  return update_PA(*E(work), Px_new, Px_new_idx, P_new_n, true, Ax_new, Ax_new_idx, A_new_n, true);
}

// ---- warm start (a14)
static c_int warm(Engine &e, const c_float *x, const c_float *y) {
  DeviceGuard guard(e.device);
  if (!e.st.warm_start) e.st.warm_start = 1;
  // stage in pol_x / pol_y (free outside a polish launch)
  if (x) { c_int rc = upload_vector(e, e.d.pol_x, x, e.d.n); if (rc) return rc; }
  if (y) { c_int rc = upload_vector(e, e.d.pol_y, y, e.d.m); if (rc) return rc; }
  CU_OK(launch_warm_start(e.d, x ? e.d.pol_x : nullptr, y ? e.d.pol_y : nullptr, e.st.scaling != 0, e.stream));
  e.prof.launches += x ? 2 : 1;
  CU_OK(cudaStreamSynchronize(e.stream));
  return 0;
}
// both engines of a tiny-mode workspace take the caller's start; neither owes the other a hand-over afterwards
static c_int warm_both(Engine &e, const c_float *x, const c_float *y) {
  const c_int rc = warm(e, x, y);
  if (rc == 0 && e.tiny) {
    osqp_batch_warm_start(e.tiny, x, e.d.m > 0 ? y : nullptr);
    e.last_path = 0;
  }
  return rc;
}
c_int osqp_warm_start(OSQPWorkspace *work, const c_float *x, const c_float *y) {
  if (!work) return 1;
  return warm_both(*E(work), x, y);
}
c_int osqp_warm_start_x(OSQPWorkspace *work, const c_float *x) {
  if (!work) return 1;
  return warm_both(*E(work), x, nullptr);
}
c_int osqp_warm_start_y(OSQPWorkspace *work, const c_float *y) {
  if (!work) return 1;
  return warm_both(*E(work), nullptr, y);
}

// ---- settings (a15)
#define ENGINE_SETTER(NAME, TYPE, FIELD, BADCOND, MSG)                            \
  c_int NAME(OSQPWorkspace *work, TYPE v) {                                       \
    if (!work) return 1;                                                          \
    if (BADCOND) { fprintf(stderr, "ERROR in " #NAME ": " MSG "\n"); return 1; }  \
    E(work)->st.FIELD = v;                                                        \
    return 0;                                                                     \
  }
ENGINE_SETTER(osqp_update_max_iter, c_int, max_iter, v <= 0, "max_iter must be positive")
ENGINE_SETTER(osqp_update_eps_abs, c_float, eps_abs, v < 0, "eps_abs must be nonnegative")
ENGINE_SETTER(osqp_update_eps_rel, c_float, eps_rel, v < 0, "eps_rel must be nonnegative")
ENGINE_SETTER(osqp_update_eps_prim_inf, c_float, eps_prim_inf, v < 0, "eps_prim_inf must be nonnegative")
ENGINE_SETTER(osqp_update_eps_dual_inf, c_float, eps_dual_inf, v < 0, "eps_dual_inf must be nonnegative")
ENGINE_SETTER(osqp_update_alpha, c_float, alpha, (v <= 0 || v >= 2), "alpha must be between 0 and 2")
ENGINE_SETTER(osqp_update_delta, c_float, delta, v <= 0, "delta must be positive")
ENGINE_SETTER(osqp_update_polish_refine_iter, c_int, polish_refine_iter, v < 0, "polish_refine_iter must be nonnegative")
ENGINE_SETTER(osqp_update_verbose, c_int, verbose, (v != 0 && v != 1), "verbose should be either 0 or 1")
ENGINE_SETTER(osqp_update_scaled_termination, c_int, scaled_termination, (v != 0 && v != 1), "scaled_termination should be either 0 or 1")
ENGINE_SETTER(osqp_update_check_termination, c_int, check_termination, v < 0, "check_termination should be nonnegative")
ENGINE_SETTER(osqp_update_warm_start, c_int, warm_start, (v != 0 && v != 1), "warm_start should be either 0 or 1")
ENGINE_SETTER(osqp_update_time_limit, c_float, time_limit, v < 0, "time_limit must be nonnegative")

c_int osqp_update_polish(OSQPWorkspace *work, c_int v) {
  if (!work) return 1;
  if (v != 0 && v != 1) { fprintf(stderr, "ERROR in osqp_update_polish: polish should be either 0 or 1\n"); return 1; }
  E(work)->st.polish = v;
  E(work)->info.polish_time = 0.0;
  return 0;
}

c_int osqp_update_rho(OSQPWorkspace *work, c_float rho_new) {
  if (!work) return 1;
  Engine &e = *E(work);
  if (rho_new <= 0) { fprintf(stderr, "ERROR in osqp_update_rho: rho must be positive\n"); return 1; }
  DeviceGuard guard(e.device);
  begin_update(e);
  const double t0 = now_s();
  e.st.rho = std::min(std::max(rho_new, kRhoMin), kRhoMax);
  e.h_state->rho = e.st.rho;
  e.h_state->needs_refresh = 1;
  { c_int rc = push_state(e); if (rc) return rc; }
  CU_OK(launch_apply_rho(e.d, e.st.rho, e.stream));
  CU_OK(launch_precond(e.d, e.st.sigma, e.stream));
  e.wood_dirty = true;
  e.prof.launches += 2;
  CU_OK(cudaStreamSynchronize(e.stream));
  tiny_rebuild(e);
  e.info.update_time += now_s() - t0;
  return 0;
}

// ---- engine extensions (include/osqp_b200.h)
c_int osqp_b200_get_profile(const OSQPWorkspace *work, OSQPB200Profile *out) {
  if (!work || !out) return 1;
  *out = E(work)->prof;
  return 0;
}

c_int osqp_b200_set_pcg(OSQPWorkspace *work, c_float eta, c_float floor_rel, c_int max_iter, c_int refresh_every) {
  if (!work) return 1;
  Engine &e = *E(work);
  if (eta > 0) e.pcg_eta = eta;
  if (floor_rel > 0) e.pcg_floor = floor_rel;
  if (max_iter > 0) e.pcg_max_iter = (int)max_iter;
  if (refresh_every >= 0) e.refresh_every = (int)refresh_every;
  return 0;
}

c_int osqp_b200_spmv(OSQPWorkspace *work, c_int which, const c_float *in_host, c_float *out_host, c_int reps,
                     c_float *ms_per_rep) {
  if (!work || which < 0 || (which > 2 && which < 10) || which > 12 || reps < 1) return 1;
  Engine &e = *E(work);
  DeviceGuard guard(e.device);
  const int n = e.d.n, m = e.d.m;
  const int wm = (int)(which % 10);
  if (m == 0 && wm != 2) return 1;
  const int in_len = (wm == 1) ? m : n, out_len = (wm == 0) ? m : n;
  // scratch: PCG vectors are free between solves (uu/w are n, t/tr are m)
  double *din = (wm == 1) ? e.d.tr : e.d.uu;
  double *dout = (wm == 0) ? e.d.t : e.d.w;
  { c_int rc = upload_vector(e, din, in_host, in_len); if (rc) return rc; }
  CU_OK(launch_with_pair_fallback(e, [&]() {  // warm-up
    return launch_spmv(e.d, (int)which, din, dout, e.st.sigma, e.geom, e.stream);
  }));
  CU_OK(cudaEventRecord(e.ev0, e.stream));
  for (c_int r = 0; r < reps; r++) CU_OK(launch_spmv(e.d, (int)which, din, dout, e.st.sigma, e.geom, e.stream));
  CU_OK(cudaEventRecord(e.ev1, e.stream));
  e.prof.launches += (c_int)reps + 1;
  if (out_host && out_len > 0)
    CU_OK(cudaMemcpyAsync(out_host, dout, (size_t)out_len * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  CU_OK(cudaStreamSynchronize(e.stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e.ev0, e.ev1);
  if (ms_per_rep) *ms_per_rep = (double)ms / (double)reps;
  e.h_state->needs_refresh = 1;
  return 0;
}

c_int osqp_b200_get_scaling(OSQPWorkspace *work, c_float *D, c_float *Ev, c_float *c) {
  if (!work) return 1;
  Engine &e = *E(work);
  DeviceGuard guard(e.device);
  if (D) CU_OK(cudaMemcpyAsync(D, e.d.D, (size_t)e.d.n * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  if (Ev && e.d.m > 0) CU_OK(cudaMemcpyAsync(Ev, e.d.E, (size_t)e.d.m * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  { c_int rc = pull_state(e); if (rc) return rc; }
  if (c) *c = e.h_state->c;
  return 0;
}

}  // extern "C"
