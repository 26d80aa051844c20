// kernels.cu -- hand-written sm_100a kernels of the B200 OSQP engine.
//
// One osqp_solve = ONE persistent cooperative launch (admm_kernel): the whole ADMM loop of
// libosqp 0.6.2 runs on the device -- reduced-KKT Jacobi-PCG
// built on CSR SpMV with A, A' and P+sigma I, fused with the [l,u] projection, the dual update,
// the infinity-norm residuals, termination / infeasibility tests and adaptive rho.  Phases are
// separated by a hand-written grid barrier; dot products and norms ride on the same barrier
// through a fixed-order (deterministic) two-level reduction.  Nothing here is a dense
// contraction, so tensor cores are not used: the roofline is HBM.
//
// Work split: block b owns a contiguous, nnz-balanced range of rows of A (m-range) and of
// P / A' (n-range); all element-wise vector work on an index is done by the block owning it.
#include "engine.cuh"
#include "admm_rules.cuh"

#include <cooperative_groups.h>
#include <math.h>
#include <string.h>

#include <map>
#include <mutex>
#include <type_traits>
#include <utility>

// This file is compiled FOUR times.  The plain compilation holds every kernel and decides the storage mode of a
// workspace at run time (CSR or tile streams, scan or lane-row layout, cluster pairs or not, fp64 or fp32 slices,
// Jacobi / Woodbury / slack-elimination preconditioner).  kernels_fast.cu, kernels_fast2.cu and kernels_fast3.cu
// include it again with OSQP_B200_FAST = 1, 2, 3: the same admm_kernel / polish_kernel with one common mode
// (fast_mode() below) fixed at compile time.  Less than half the code and fewer spills: measured 5 % faster on every
// phase of config 2 (profiles/r2_ncu_admm.md section 5).  MODE(x, v) is x in the plain compilation and the constant v
// in the fixed-mode ones; only those stream fp32 copies of the matrix values in their PCG phases (DevPtrs::mat32).
#ifdef OSQP_B200_FAST
#define MODE(x, v) (v)
#define FAST_PAIRED (OSQP_B200_FAST == 1)  // fixed mode 1: [A; P] in cluster pairs; fixed mode 2 (kernels_fast2.cu): no pairs
#define FAST_SLACK (OSQP_B200_FAST == 3)   // fixed mode 3 (kernels_fast3.cu): no pairs, slack-elimination preconditioner
#else
#define MODE(x, v) (x)
#define FAST_PAIRED 0
#define FAST_SLACK 0
#endif

namespace osqpb200 {

namespace {

constexpr int kThreads = 32 * kWarps;  // threads per block of the cooperative kernels (512: 128 registers per thread)


// ------------------------------------------------------------------ memory helpers
// Matrix streams are read once per phase and never written inside a launch: non-coherent path,
// no L1 allocation, so L1 stays free for the gathered dense vector.
__device__ __forceinline__ double ld_stream(const double *p) {
  double v;
  asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_stream(const int *p) {
  int v;
  asm("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Block 0 / thread 0 attributes its wall time to phase classes (shared-memory accumulators, copied to DevInfo at exit).
//  0 stream [A;P]  1 barrier  2 combine t,Pu (+delta)  3 reduce+barrier  4 stream A'  5 barrier  6 vector update
//  7 reduce+barrier | 8 z,y,x update + rhs vector + barrier  9 stream A' rhs + barrier  10 rhs/residual + reduce
//  11 residual refresh  12 update_info  13 rho update  14 other
struct PhaseClock {
  double *acc;
  unsigned long long t;
  bool on;
  __device__ __forceinline__ void start(double *a) {
    acc = a;
    on = (blockIdx.x == 0 && threadIdx.x == 0);
    if (on) {
      for (int k = 0; k < kPhases; k++) acc[k] = 0.0;
      t = globaltimer_ns();
    }
  }
  __device__ __forceinline__ void tick(int k) {
    if (on) {
      const unsigned long long now = globaltimer_ns();
      acc[k] += (double)(now - t) * 1e-3;
      t = now;
    }
  }
};

// ------------------------------------------------------------------ tile streams (engine.cuh TileStreamDev)
// Shared-memory staging of the gathered vector slice: one elected thread arms an mbarrier with the byte count and
// issues 1-D bulk async copies (cp.async.bulk.shared.global -> UBLKCP); every warp waits on the mbarrier right
// before its first gather, after its first matrix loads are already in flight.
struct Slice {
  unsigned xs;       // shared-space address of the staged slice
  unsigned ys;       // shared-space address of the row sums of a cluster pair (paired streams only)
  unsigned mbar;     // shared-space address of its mbarrier
  unsigned parity;
  unsigned long long *probe;  // spmv_stream_kernel only: globaltimer when the slice has landed (else nullptr)
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void slice_init(Slice &S, const DevPtrs &d) {
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  S.xs = smem_u32(dyn_smem);
  S.ys = S.xs + 8u * (unsigned)d.smem_x_elems;
  S.mbar = S.ys + 8u * (unsigned)d.smem_rows;
  S.parity = 0;
  S.probe = nullptr;
  if (MODE(d.blocked, 1)) {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(S.mbar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }
}

__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
  unsigned done = 0;
  while (!done) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(mbar), "r"(parity)
        : "memory");
  }
}

__device__ __forceinline__ double lds_f64(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds_f32(unsigned addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f64(unsigned addr, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
// distributed shared memory: the same offset in the shared memory of block `rank` of this cluster
__device__ __forceinline__ double ld_dsmem_f64(unsigned addr, unsigned rank) {
  unsigned raddr;
  double v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(addr), "r"(rank));
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(raddr) : "memory");
  return v;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Fire-and-forget bulk prefetch of [ptr, ptr + bytes) into L2 (16 B aligned pieces).
__device__ __forceinline__ void l2_prefetch(const void *ptr, unsigned bytes) {
  const unsigned long long a0 = (unsigned long long)ptr & ~15ull;
  const unsigned long long a1 = ((unsigned long long)ptr + bytes + 15ull) & ~15ull;
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"((unsigned)(a1 - a0)) : "memory");
}

#ifndef OSQP_B200_DEPTH
#define OSQP_B200_DEPTH 4
#endif
#ifndef OSQP_B200_ILP
#define OSQP_B200_ILP 2
#endif
constexpr int kILP = OSQP_B200_ILP;      // chunks whose segmented scans are interleaved (stream_phase_impl consume2)
constexpr int kDepth = OSQP_B200_DEPTH;  // chunks (one quad per lane, 40 B) of matrix stream in flight per lane

struct QuadSlot {
  double v0, v1, v2, v3;
  unsigned c01, c23;
};

// pv / pc: this lane's byte addresses inside the value / column streams
// pv / pc: this lane's byte addresses inside the value / column streams.  Explicit L2 eviction priorities
// (createpolicy + .L2::cache_hint: evict_last on some streams, evict_first on the others) were measured on B200 and
// made every phase slower (profiles/r1_notes.md), so the loads carry no hint.
__device__ __forceinline__ void quad_load(QuadSlot &q, const char *pv, const char *pc, bool active) {
  q.v0 = q.v1 = q.v2 = q.v3 = 0.0;
  q.c01 = q.c23 = 0u;
  if (active) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(q.v0), "=d"(q.v1) : "l"(pv));
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(q.v2), "=d"(q.v3) : "l"(pv + 512));
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(q.c01), "=r"(q.c23) : "l"(pc));
  }
}

// The same for the fp32 copy of the values (TileStreamDev::val32, entry order): one 16 B load per lane and chunk.
struct QuadSlot32 {
  float v0, v1, v2, v3;
  unsigned c01, c23;
};
__device__ __forceinline__ void quad_load(QuadSlot32 &q, const char *pv, const char *pc, bool active) {
  q.v0 = q.v1 = q.v2 = q.v3 = 0.f;
  q.c01 = q.c23 = 0u;
  if (active) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(q.v0), "=f"(q.v1), "=f"(q.v2), "=f"(q.v3)
                 : "l"(pv));
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(q.c01), "=r"(q.c23) : "l"(pc));
  }
}

// Warm L2 with the head of this warp's stream of T; call before the grid barrier that precedes the phase.
template <bool kM32 = false>
__device__ __forceinline__ void stream_prefetch_head(const TileStreamDev &T) {
  const int lane = threadIdx.x & 31, wid = blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int q0 = __ldg(T.w_q0 + wid), total = __ldg(T.w_qn + wid);
  if (lane < T.pf_chunks && lane * 32 < total) {
    const unsigned quads = (unsigned)min(32, total - lane * 32);
    if (kM32) l2_prefetch(T.val32 + 4ll * (q0 + lane * 32), 512u);
    else l2_prefetch(T.val + 4ll * (q0 + lane * 32), 1024u);
    l2_prefetch(T.cf + 4ll * (q0 + lane * 32), quads * 8u);
  }
}

// One phase of  part[group][row] = sum over the block's column group of M[row, :] vec  for the rows of every warp.
// Must be entered by all threads of the block after a grid barrier (the previous users of the slice are done and
// `vec` is complete and visible).
// kD: chunks in flight per lane.  kPair: the row sums go to this block's shared-memory accumulators (cluster pairs,
// stream_phase_paired) instead of part[group][row].
// kF32: `vec` is an array of float (the fp32 shadows DevPtrs::uu32 / tr32); the slice is staged and gathered as fp32.
template <int kD, bool kPair, bool kF32>
__device__ __noinline__ void stream_phase_impl(Slice &S, const TileStreamDev &T, const void *__restrict__ vec) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.x;
  const int grp = __ldg(T.blk_group + b);
  const unsigned parity = S.parity;
  S.parity ^= 1u;
  if (tid == 0) {
    // the vector was written with ordinary stores by other blocks before the grid barrier: order those
    // generic-proxy writes before the async-proxy (TMA) reads
    asm volatile("fence.proxy.async;" ::: "memory");
    const int col0 = __ldg(T.grp_col0 + grp), ncols = __ldg(T.grp_col0 + grp + 1) - col0;
    const unsigned bytes = ((unsigned)ncols * (kF32 ? 4u : 8u) + 15u) & ~15u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(S.mbar), "r"(bytes) : "memory");
    const char *g = reinterpret_cast<const char *>(vec) + (size_t)col0 * (kF32 ? 4u : 8u);
    for (unsigned off = 0; off < bytes; off += 32768u) {
      const unsigned chunk = (bytes - off < 32768u) ? (bytes - off) : 32768u;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       S.xs + off),
                   "l"(g + off), "r"(chunk), "r"(S.mbar)
                   : "memory");
    }
  }
  const int wid = b * kWarps + warp;
  const int q0 = __ldg(T.w_q0 + wid), L = __ldg(T.w_qn + wid);  // first quad (chunk aligned), quads of this warp
  if (L <= 0) {
    // warp 0 still drains the mbarrier it armed so that the next phase never re-arms a pending one
    if (warp == 0) mbar_wait(S.mbar, parity);
    return;
  }
  double *__restrict__ out = T.part + (size_t)grp * T.srows + __ldg(T.w_row0 + wid);
  const unsigned out_s = kPair ? S.ys + 8u * (unsigned)(__ldg(T.w_row0 + wid) - __ldg(T.blk_row0 + b)) : 0u;
  const char *pv = reinterpret_cast<const char *>(T.val) + 32ll * q0 + 16 * lane;
  const char *pc = reinterpret_cast<const char *>(T.cf) + 8ll * (q0 + lane);
  const char *pf_v = reinterpret_cast<const char *>(T.val) + 32ll * q0;
  const char *pf_c = reinterpret_cast<const char *>(T.cf) + 8ll * q0;
  const int pf = T.pf_chunks;
  const unsigned xs = S.xs;
  const unsigned lt = (1u << lane) - 1u;
  int left = L - lane;   // > 0: this lane still has a quad in the chunk being issued
  int issued = 0;        // chunks issued so far
  int rdone = 0;
  double carry = 0.0;

  auto issue = [&](QuadSlot &q) {
    quad_load(q, pv, pc, left > 0);
    pv += 1024;
    pc += 256;
    left -= 32;
    if ((issued & 3) == 0 && lane == 0 && pf > 0) {  // every 4th chunk: the 4 chunks `pf` ahead go to L2
      const int pq = (issued + pf) * 32;
      if (pq < L) {
        const unsigned quads = (unsigned)min(128, L - pq);
        l2_prefetch(pf_v + 32ll * pq, ((quads + 31u) & ~31u) * 32u);
        l2_prefetch(pf_c + 8ll * pq, quads * 8u);
      }
    }
    issued++;
  };
  auto consume = [&](const QuadSlot &q) {
    double x0, x1, x2, x3;
    if (kF32) {
      x0 = (double)lds_f32(xs + ((q.c01 & 0x7fffu) << 2)); x1 = (double)lds_f32(xs + ((q.c01 >> 14) & 0x1fffcu));
      x2 = (double)lds_f32(xs + ((q.c23 & 0x7fffu) << 2)); x3 = (double)lds_f32(xs + ((q.c23 >> 14) & 0x1fffcu));
    } else {
      x0 = lds_f64(xs + ((q.c01 & 0x7fffu) << 3)); x1 = lds_f64(xs + ((q.c01 >> 13) & 0x3fff8u));
      x2 = lds_f64(xs + ((q.c23 & 0x7fffu) << 3)); x3 = lds_f64(xs + ((q.c23 >> 13) & 0x3fff8u));
    }
    double inc = q.v0 * x0;
    inc = fma(q.v1, x1, inc);
    inc = fma(q.v2, x2, inc);
    inc = fma(q.v3, x3, inc);
    const bool flag = (q.c23 >> 31) != 0u;  // a row ends with this lane's quad
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    const unsigned below = bal & lt;
    const int h = 32 - __clz(below);  // first lane of the row this lane's quad belongs to (0 if none ended below)
#pragma unroll
    for (int dlt = 1; dlt < 32; dlt <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, inc, dlt);
      if (lane - dlt >= h) inc += t;
    }
    if (below == 0u) inc += carry;  // the row started in an earlier chunk
    if (flag) {
      if (kPair) sts_f64(out_s + 8u * (unsigned)(rdone + __popc(below)), inc);
      else out[rdone + __popc(below)] = inc;
    }
    const double last = __shfl_sync(0xffffffffu, inc, 31);
    carry = (bal >> 31) ? 0.0 : last;
    rdone += __popc(bal);
  };

  // Two chunks at a time: their gathers and the five steps of their segmented scans are independent until the carry
  // is handed from the first to the second, which doubles the instruction-level parallelism of the scan -- with 16
  // warps per SM the phase is bound by the dependent shuffle / add chain of the scan, not by memory
  // (profiles/r2_ncu_admm.md).
  auto gather = [&](const QuadSlot &q) {
    double x0, x1, x2, x3;
    if (kF32) {
      x0 = (double)lds_f32(xs + ((q.c01 & 0x7fffu) << 2)); x1 = (double)lds_f32(xs + ((q.c01 >> 14) & 0x1fffcu));
      x2 = (double)lds_f32(xs + ((q.c23 & 0x7fffu) << 2)); x3 = (double)lds_f32(xs + ((q.c23 >> 14) & 0x1fffcu));
    } else {
      x0 = lds_f64(xs + ((q.c01 & 0x7fffu) << 3)); x1 = lds_f64(xs + ((q.c01 >> 13) & 0x3fff8u));
      x2 = lds_f64(xs + ((q.c23 & 0x7fffu) << 3)); x3 = lds_f64(xs + ((q.c23 >> 13) & 0x3fff8u));
    }
    double inc = q.v0 * x0;
    inc = fma(q.v1, x1, inc);
    inc = fma(q.v2, x2, inc);
    return fma(q.v3, x3, inc);
  };
  auto consume2 = [&](const QuadSlot &qa, const QuadSlot &qb) {
    double ia = gather(qa), ib = gather(qb);
    const bool fa = (qa.c23 >> 31) != 0u, fb = (qb.c23 >> 31) != 0u;
    const unsigned bala = __ballot_sync(0xffffffffu, fa), balb = __ballot_sync(0xffffffffu, fb);
    const unsigned bela = bala & lt, belb = balb & lt;
    const int ha = 32 - __clz(bela), hb = 32 - __clz(belb);
#pragma unroll
    for (int dlt = 1; dlt < 32; dlt <<= 1) {
      const double ta = __shfl_up_sync(0xffffffffu, ia, dlt), tb = __shfl_up_sync(0xffffffffu, ib, dlt);
      if (lane - dlt >= ha) ia += ta;
      if (lane - dlt >= hb) ib += tb;
    }
    if (bela == 0u) ia += carry;
    const int ra = rdone + __popc(bela);
    if (fa) {
      if (kPair) sts_f64(out_s + 8u * (unsigned)ra, ia);
      else out[ra] = ia;
    }
    const double lasta = __shfl_sync(0xffffffffu, ia, 31);
    const double ca = (bala >> 31) ? 0.0 : lasta;
    rdone += __popc(bala);
    if (belb == 0u) ib += ca;
    const int rb = rdone + __popc(belb);
    if (fb) {
      if (kPair) sts_f64(out_s + 8u * (unsigned)rb, ib);
      else out[rb] = ib;
    }
    const double lastb = __shfl_sync(0xffffffffu, ib, 31);
    carry = (balb >> 31) ? 0.0 : lastb;
    rdone += __popc(balb);
  };

  QuadSlot slot[kD];
#pragma unroll
  for (int k = 0; k < kD; k++) issue(slot[k]);
  mbar_wait(S.mbar, parity);  // the slice has landed (the first matrix loads are already in flight)
  if (S.probe != nullptr && tid == 0) S.probe[6] = globaltimer_ns();
  const int nchunks = (L + 31) >> 5;
  for (int base = 0; base < nchunks; base += kD) {
    if (kILP == 2 && (kD % 2) == 0) {
#pragma unroll
      for (int k = 0; k < kD; k += 2) {
        consume2(slot[k], slot[k + 1]);  // slots past the end hold zeros without flags: a no-op
        issue(slot[k]);
        issue(slot[k + 1]);
      }
    } else {
#pragma unroll
      for (int k = 0; k < kD; k++) {
        consume(slot[k]);  // slots past the end hold zeros without flags: a no-op that keeps carry and rdone
        issue(slot[k]);
      }
    }
  }
}

// The same phase on the lane-row layout (TileStreamDev::lane_rows): lane l of a slice owns one stream row and meets
// quad t of it in chunk t of the slice, so the row sum is a private accumulator -- per chunk three loads, four gathers
// and four fused multiply-adds per lane, no ballot, no shuffles, no carry between chunks.  The slice descriptors
// (length, row of this lane) of the NEXT slice are loaded while the current one streams.
#ifndef OSQP_B200_DEPTH_LR
#define OSQP_B200_DEPTH_LR 4
#endif
constexpr int kDepthLR = OSQP_B200_DEPTH_LR;
#ifndef OSQP_B200_DEPTH_LR32
#define OSQP_B200_DEPTH_LR32 4
#endif
constexpr int kDepthLR32 = OSQP_B200_DEPTH_LR32;  // chunks in flight per lane on the fp32 value streams (6 registers a slot)

// kM32: the values come from the fp32 copy T.val32 (engine.cuh DevPtrs::mat32): 6 instead of 10 bytes per entry.
template <int kD, bool kPair, bool kF32, bool kM32 = false>
__device__ __noinline__ void stream_phase_lr(Slice &S, const TileStreamDev &T, const void *__restrict__ vec) {
  using Slot = typename std::conditional<kM32, QuadSlot32, QuadSlot>::type;
  constexpr int kValBytes = kM32 ? 512 : 1024;  // value bytes of a chunk
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.x;
  const int grp = __ldg(T.blk_group + b);
  const unsigned parity = S.parity;
  S.parity ^= 1u;
  if (tid == 0) {
    asm volatile("fence.proxy.async;" ::: "memory");
    const int col0 = __ldg(T.grp_col0 + grp), ncols = __ldg(T.grp_col0 + grp + 1) - col0;
    const unsigned bytes = ((unsigned)ncols * (kF32 ? 4u : 8u) + 15u) & ~15u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(S.mbar), "r"(bytes) : "memory");
    const char *g = reinterpret_cast<const char *>(vec) + (size_t)col0 * (kF32 ? 4u : 8u);
    for (unsigned off = 0; off < bytes; off += 32768u) {
      const unsigned chunk = (bytes - off < 32768u) ? (bytes - off) : 32768u;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       S.xs + off),
                   "l"(g + off), "r"(chunk), "r"(S.mbar)
                   : "memory");
    }
  }
  const int wid = b * kWarps + warp;
  const int q0 = __ldg(T.w_q0 + wid), nchunks = __ldg(T.w_qn + wid) >> 5;  // whole chunks only
  if (nchunks <= 0) {
    if (warp == 0) mbar_wait(S.mbar, parity);  // keep the mbarrier phases in step
    return;
  }
  double *__restrict__ out = T.part + (size_t)grp * T.srows;
  const unsigned out_s = kPair ? S.ys - 8u * (unsigned)__ldg(T.blk_row0 + b) : 0u;
  int sl = __ldg(T.w_s0 + wid);
  const int sl_end = __ldg(T.w_s0 + wid + 1);
  int left = __ldg(T.sl_len + sl), myrow = __ldg(T.sl_row + (size_t)sl * 32 + lane);
  int nleft = 0, nrow = -1;
  if (sl + 1 < sl_end) { nleft = __ldg(T.sl_len + sl + 1); nrow = __ldg(T.sl_row + (size_t)(sl + 1) * 32 + lane); }
  const char *vbase = kM32 ? reinterpret_cast<const char *>(T.val32) : reinterpret_cast<const char *>(T.val);
  const char *pv = vbase + (kM32 ? 16ll : 32ll) * q0 + 16 * lane;
  const char *pc = reinterpret_cast<const char *>(T.cf) + 8ll * (q0 + lane);
  const char *pf_v = vbase + (kM32 ? 16ll : 32ll) * q0;
  const char *pf_c = reinterpret_cast<const char *>(T.cf) + 8ll * q0;
  const int pf = T.pf_chunks;
  const unsigned xs = S.xs;
  int issued = 0;
  double acc = 0.0;

  auto issue = [&](Slot &q) {
    quad_load(q, pv, pc, issued < nchunks);
    pv += kValBytes;
    pc += 256;
    if ((issued & 3) == 0 && lane == 0 && pf > 0) {  // every 4th chunk: the 4 chunks `pf` ahead go to L2
      const int pq = issued + pf;
      if (pq < nchunks) {
        const unsigned ch = (unsigned)min(4, nchunks - pq);
        l2_prefetch(pf_v + (long long)kValBytes * pq, ch * (unsigned)kValBytes);
        l2_prefetch(pf_c + 256ll * pq, ch * 256u);
      }
    }
    issued++;
  };
  auto consume = [&](const Slot &q) {
    double x0, x1, x2, x3;
    if (kF32) {
      x0 = (double)lds_f32(xs + ((q.c01 & 0x7fffu) << 2)); x1 = (double)lds_f32(xs + ((q.c01 >> 14) & 0x1fffcu));
      x2 = (double)lds_f32(xs + ((q.c23 & 0x7fffu) << 2)); x3 = (double)lds_f32(xs + ((q.c23 >> 14) & 0x1fffcu));
    } else {
      x0 = lds_f64(xs + ((q.c01 & 0x7fffu) << 3)); x1 = lds_f64(xs + ((q.c01 >> 13) & 0x3fff8u));
      x2 = lds_f64(xs + ((q.c23 & 0x7fffu) << 3)); x3 = lds_f64(xs + ((q.c23 >> 13) & 0x3fff8u));
    }
    double inc = (double)q.v0 * x0;  // (fp32 values and fp32 slice entries: the products are exact in fp64)
    inc = fma((double)q.v1, x1, inc);
    inc = fma((double)q.v2, x2, inc);
    inc = fma((double)q.v3, x3, inc);
    acc += inc;
    if (--left == 0) {  // warp-uniform: the slice ends with this chunk
      if (myrow >= 0) {
        if (kPair) sts_f64(out_s + 8u * (unsigned)myrow, acc);
        else out[myrow] = acc;
      }
      acc = 0.0;
      sl++;
      left = nleft;
      myrow = nrow;
      if (sl + 1 < sl_end) { nleft = __ldg(T.sl_len + sl + 1); nrow = __ldg(T.sl_row + (size_t)(sl + 1) * 32 + lane); }
    }
  };

  Slot slot[kD];
#pragma unroll
  for (int k = 0; k < kD; k++) issue(slot[k]);
  mbar_wait(S.mbar, parity);  // the slice has landed (the first matrix loads are already in flight)
  if (S.probe != nullptr && tid == 0) S.probe[6] = globaltimer_ns();
  for (int base = 0; base < nchunks; base += kD) {
#pragma unroll
    for (int k = 0; k < kD; k++) {
      if (base + k < nchunks) consume(slot[k]);
      issue(slot[k]);
    }
  }
}

__device__ __forceinline__ void stream_phase(Slice &S, const TileStreamDev &T, const double *__restrict__ vec) {
  if (MODE(T.lane_rows, 1)) stream_phase_lr<kDepthLR, false, false>(S, T, vec);
  else stream_phase_impl<kDepth, false, false>(S, T, vec);
}
template <bool kM32 = false>
__device__ __forceinline__ void stream_phase_f32(Slice &S, const TileStreamDev &T, const float *__restrict__ vec) {
  if (MODE(T.lane_rows, 1)) stream_phase_lr<kM32 ? kDepthLR32 : kDepthLR, false, true, kM32>(S, T, vec);
  else stream_phase_impl<kDepth, false, true>(S, T, vec);
}

// Paired stream (TileStreamDev::paired): blocks 2p / 2p+1 of a cluster stream column groups 0 / 1 of the same row
// range into their shared-memory accumulators; after a cluster barrier each block finalises one half of the rows,
// reading its partner's partial sums through distributed shared memory, and hands the complete row sum to
// fin(stacked_row, sum).  No partial vector goes through global memory and no grid barrier is needed for the combine.
// Every thread of both blocks must call this; the accumulators are next written after at least one grid barrier.
template <bool kF32 = false, bool kM32 = false, typename Fin>
__device__ __forceinline__ void stream_phase_paired(Slice &S, const TileStreamDev &T, const void *__restrict__ vec,
                                                    Fin fin) {
  if (MODE(T.lane_rows, 1)) stream_phase_lr<kM32 ? kDepthLR32 : kDepthLR, true, kF32, kM32>(S, T, vec);
  else stream_phase_impl<kDepth, true, kF32>(S, T, vec);
  cluster_sync();
  const int b = blockIdx.x, r0 = __ldg(T.blk_row0 + b), r1 = __ldg(T.blk_row1 + b);
  const unsigned rank = (unsigned)b & 1u;
  const int mid = r0 + ((r1 - r0 + 1) >> 1);
  const int lo = rank ? mid : r0, hi = rank ? r1 : mid;
  for (int r = lo + (int)threadIdx.x; r < hi; r += (int)blockDim.x) {
    const unsigned a = S.ys + 8u * (unsigned)(r - r0);
    const double mine = lds_f64(a), theirs = ld_dsmem_f64(a, rank ^ 1u);
    fin(r, rank ? theirs + mine : mine + theirs);  // group 0 first, as in part_sum
  }
  // the partner may still be reading this block's accumulators: keep both blocks together until the reads are done
  // (the grid barrier that follows every call would also guarantee it; this one makes the lifetime rule of
  // distributed shared memory local and keeps compute-sanitizer quiet)
  cluster_sync();
}

// ------------------------------------------------------------------ grid barrier + reductions
// One monotonically increasing arrival counter (reset to 0 by the host before every cooperative launch): barrier
// k is passed once the counter reaches k * nblocks.  Per block one thread arrives with a release reduction (no
// return value, no reset, no second atomic) and spins on an acquire load; critical path = one L2 round trip + one
// poll.  The block's other threads are ordered through the two __syncthreads().
struct Grid {
  unsigned *count;
  unsigned nblocks, target;
  int bank, fbank;
  double *red;
  unsigned long long *facc;  // [3][kFxSlots] fixed-point accumulators of reduce_and_barrier_fx (zero at launch)
};

__device__ __forceinline__ void grid_init(Grid &g, const DevPtrs &d) {
  g.count = d.bar;
  g.nblocks = gridDim.x;
  g.target = 0u;
  g.bank = 0;
  g.fbank = 0;
  g.red = d.red;
  g.facc = reinterpret_cast<unsigned long long *>(d.bar + 4);
}

__device__ __forceinline__ void grid_barrier(Grid &g) {
  __syncthreads();
  if (g.nblocks > 1) {
    g.target += g.nblocks;
    if (threadIdx.x == 0) {
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(g.count) : "memory");
      while ((int)(ld_acquire(g.count) - g.target) < 0) {
      }
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
  }
}

// op mask bit k = 1: slot k is a max-reduction, else a sum.
struct RedSmem {
  double part[kWarps][kRedSlots];  // the cooperative kernels run kThreads = 32 * kWarps threads per block
  double res[kRedSlots];
  unsigned long long fx[kFxSlots];
};

template <int NV>
__device__ __forceinline__ void reduce_and_barrier(Grid &g, RedSmem &sm, double (&v)[NV], unsigned maxmask) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; k++) {
    double a = v[k];
    if (maxmask & (1u << k)) {
#pragma unroll
      for (int o = 16; o; o >>= 1) a = fmax(a, __shfl_xor_sync(0xffffffffu, a, o));
    } else {
#pragma unroll
      for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    }
    if (lane == 0) sm.part[warp][k] = a;
  }
  __syncthreads();
  for (int k = warp; k < NV; k += nwarps) {  // one warp per slot: fixed shuffle tree over the warps' partials
    const bool ismax = maxmask & (1u << k);
    double a = lane < nwarps ? sm.part[lane][k] : (ismax ? -INFINITY : 0.0);
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const double x = __shfl_xor_sync(0xffffffffu, a, o);
      a = ismax ? fmax(a, x) : a + x;
    }
    if (lane == 0) {
      if (g.nblocks > 1) g.red[((size_t)g.bank * kRedSlots + k) * g.nblocks + blockIdx.x] = a;
      else sm.res[k] = a;
    }
  }
  grid_barrier(g);
  if (g.nblocks > 1) {
    for (int k = warp; k < NV; k += nwarps) {
      const double *src = g.red + ((size_t)g.bank * kRedSlots + k) * g.nblocks;
      const bool ismax = maxmask & (1u << k);
      double a = ismax ? -INFINITY : 0.0;
      for (unsigned i = lane; i < g.nblocks; i += 32) {
        double x = __ldcg(src + i);
        a = ismax ? fmax(a, x) : a + x;
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        double x = __shfl_xor_sync(0xffffffffu, a, o);
        a = ismax ? fmax(a, x) : a + x;
      }
      if (lane == 0) sm.res[k] = a;
    }
    g.bank ^= 1;
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < NV; k++) v[k] = sm.res[k];
  __syncthreads();  // sm.res may be rewritten by the next reduction
}

// Same contract, for the reductions inside the PCG iteration: the cross-block stage is done by the L2 atomic units
// instead of 148 loads per block after the barrier.  Sums are accumulated as 64-bit fixed point (integer addition is
// associative: bit-identical from run to run, like the fixed-order tree), scaled so that `ref` -- a positive value
// every thread of the grid holds identically and that is within a few binary orders of the result -- sits at 2^50;
// maxima of non-negative values use atomicMax on the bit pattern.  A block whose partial does not fit (|.| >= 2^55
// after scaling, or not finite) raises a flag, and a sum that comes out with fewer than 36 significant bits means
// `ref` was far too large: in both cases every block finishes through the plain two-level path on the fp64
// partials that are always written as well.  Three accumulator banks: call k uses bank k % 3 and block 0
// clears bank (k + 2) % 3 after the barrier of call k (its last readers finished before arriving at that barrier,
// its next writers start after the barrier of call k + 1).
template <int NV>
__device__ __forceinline__ void reduce_and_barrier_fx(Grid &g, RedSmem &sm, double (&v)[NV], unsigned maxmask,
                                                      double ref) {
  static_assert(NV < kFxSlots, "one slot is the overflow flag");
  if (!(g.nblocks > 1 && isfinite(ref) && ref > 0.0)) {  // uniform over the grid
    reduce_and_barrier<NV>(g, sm, v, maxmask);
    return;
  }
  const double scale = scalbn(1.0, 50 - ilogb(ref));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; k++) {
    double a = v[k];
    if (maxmask & (1u << k)) {
#pragma unroll
      for (int o = 16; o; o >>= 1) a = fmax(a, __shfl_xor_sync(0xffffffffu, a, o));
    } else {
#pragma unroll
      for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    }
    if (lane == 0) sm.part[warp][k] = a;
  }
  __syncthreads();
  unsigned long long *acc = g.facc + g.fbank * kFxSlots;
  if (warp < NV) {  // one warp per slot: fixed shuffle tree over the warps' partials, lane 0 feeds the L2 atomic
    const int k = warp;
    const bool ismax = maxmask & (1u << k);
    double a = lane < nwarps ? sm.part[lane][k] : (ismax ? -INFINITY : 0.0);
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const double x = __shfl_xor_sync(0xffffffffu, a, o);
      a = ismax ? fmax(a, x) : a + x;
    }
    if (lane == 0) {
      g.red[((size_t)g.bank * kRedSlots + k) * g.nblocks + blockIdx.x] = a;  // for the fallback
      if (ismax) {
        if (a >= 0.0 && isfinite(a)) atomicMax(acc + k, (unsigned long long)__double_as_longlong(a));
        else atomicMax(acc + kFxSlots - 1, 1ull);
      } else {
        const double sc = a * scale;
        if (fabs(sc) < 36028797018963968.0) atomicAdd(acc + k, (unsigned long long)__double2ll_rn(sc));
        else atomicMax(acc + kFxSlots - 1, 1ull);
      }
    }
  }
  grid_barrier(g);
  if (threadIdx.x < kFxSlots) {
    sm.fx[threadIdx.x] = __ldcg(acc + threadIdx.x);
    if (blockIdx.x == 0) g.facc[((g.fbank + 2) % 3) * kFxSlots + threadIdx.x] = 0ull;
  }
  __syncthreads();
  bool fallback = sm.fx[kFxSlots - 1] != 0ull;  // somebody's partial did not fit
#pragma unroll
  for (int k = 0; k < NV; k++) {  // or `ref` was far too large: fewer than 36 significant bits in a sum
    const long long t = (long long)sm.fx[k];
    if (!(maxmask & (1u << k)) && (t < 0 ? -t : t) < (1ll << 36)) fallback = true;
  }
  if (fallback) {  // uniform over the grid: every block read the same accumulators
    for (int k = warp; k < NV; k += nwarps) {
      const double *src = g.red + ((size_t)g.bank * kRedSlots + k) * g.nblocks;
      const bool ismax = maxmask & (1u << k);
      double a = ismax ? -INFINITY : 0.0;
      for (unsigned i = lane; i < g.nblocks; i += 32) {
        double x = __ldcg(src + i);
        a = ismax ? fmax(a, x) : a + x;
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        double x = __shfl_xor_sync(0xffffffffu, a, o);
        a = ismax ? fmax(a, x) : a + x;
      }
      if (lane == 0) sm.res[k] = a;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; k++) v[k] = sm.res[k];
  } else {
    const double inv = 1.0 / scale;  // a power of two: exact
#pragma unroll
    for (int k = 0; k < NV; k++)
      v[k] = (maxmask & (1u << k)) ? __longlong_as_double((long long)sm.fx[k]) : (double)(long long)sm.fx[k] * inv;
  }
  g.bank ^= 1;
  g.fbank = (g.fbank + 1) % 3;
  __syncthreads();  // sm.fx / sm.res may be rewritten by the next reduction
}

// ------------------------------------------------------------------ row-parallel SpMV building block
// `lanes` consecutive threads cooperate on one row (128-bit friendly: consecutive lanes read
// consecutive (val, col) pairs -> fully coalesced 8 B + 4 B streams), partial sums are combined
// with width-limited warp shuffles, and lane 0 of the group runs the epilogue.  The loop trip
// count is uniform across the block so the shuffles are always convergent.
template <int NV>
struct Acc {
  double a[NV];
};

template <int NV, typename Term>
__device__ __forceinline__ void row_accumulate(const int *__restrict__ rowptr, const int *__restrict__ col,
                                               const double *__restrict__ val, int row, bool valid, int sub, int lanes,
                                               Acc<NV> &acc, Term term) {
  int k0 = 0, k1 = 0;
  if (valid) {
    k0 = ld_stream(rowptr + row);
    k1 = ld_stream(rowptr + row + 1);
  }
#pragma unroll 4
  for (int k = k0 + sub; k < k1; k += lanes) {
    const int c = ld_stream(col + k);
    const double a = ld_stream(val + k);
    term(acc, c, a);
  }
}

template <int NV>
__device__ __forceinline__ void group_reduce(Acc<NV> &acc, int lanes) {
#pragma unroll
  for (int v = 0; v < NV; v++) {
    double a = acc.a[v];
    for (int o = lanes >> 1; o; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o, lanes);
    acc.a[v] = a;
  }
}

// PCG breakdown (alpha <= 0 or not finite while the residual is still above its threshold).  delta = u'Ku is a plain
// sum of products and denom = p'Kp follows from it by the Chronopoulos-Gear recurrence: a non-positive delta, or a
// denom that is negative by far more than the cancellation in that recurrence can explain, proves that K is not
// positive definite -- the solve ends with Non_convex instead of carrying on with truncated directions.  Anything
// else (stagnation at round-off level) just ends this inner solve, as before.
__device__ __forceinline__ bool pcg_negative_curvature(double delta, double denom) {
  return delta <= 0.0 || denom < -1e-8 * fabs(delta);
}

// ------------------------------------------------------------------ PCG on K = P + sigma I + A' diag(rho) A
// Return value: iterations run; -(iterations + 1) when the run ended on proven negative curvature.
// Chronopoulos-Gear single-reduction variant: 3 grid barriers per iteration
// (after t = A u | after w = K u with delta = w.u | after the vector updates with gamma, |r|inf).
// On entry: r, uu = Minv r are set on every n-range, gamma = r.uu and rn = |r|inf are reduced.
// x (n) is advanced in place; if zvec != nullptr it is advanced with A p (z = A x recurrence).
struct PcgVecs {
  double *r, *uu, *p, *s, *w, *t, *tr, *Ap;
};

// u = M^{-1} r and tr = rho .* (A u) as the PCG phases gather them: with fp32 slices the value is rounded to fp32 ONCE,
// here, and the rounded value is what every recurrence and dot product sees (engine.cuh DevPtrs::f32_slices).
__device__ __forceinline__ double store_u(const DevPtrs &d, double *uvec, int j, double u) {
  if (MODE(d.f32_slices, 1)) {
    const float f = (float)u;
    d.uu32[j] = f;
    u = (double)f;
  }
  uvec[j] = u;
  return u;
}
__device__ __forceinline__ double store_tr(const DevPtrs &d, double *trvec, int i, double tr) {
  if (MODE(d.f32_slices, 1)) {
    const float f = (float)tr;
    d.tr32[i] = f;
    tr = (double)f;
  }
  trvec[i] = tr;
  return tr;
}

// Woodbury part of the preconditioner (defined further down)
__device__ __noinline__ double wood_apply(Grid &g, RedSmem &sm, const DevPtrs &d, const double *Dinv,
                                          const double *rvec, double *uvec, int n0, int n1);

__device__ __noinline__ int pcg_run(Grid &g, RedSmem &sm, PhaseClock &pc, const DevPtrs &d, const PcgVecs &v,
                                    const double *rho_vec,
                                    const double *Minv, double sigma, double *xvec, double *zvec, double gamma,
                                    double rn, double thresh, int max_it, int m0, int m1, int n0, int n1,
                                    double thresh_floor = 0.0, double eta_e2 = 0.0) {
  const int tid = threadIdx.x, nth = blockDim.x;
  const int lanesA = d.A.lanes, lanesN = d.At.lanes;
  const int subA = tid & (lanesA - 1), grpA = tid / lanesA, ngrpA = nth / lanesA;
  const int subN = tid & (lanesN - 1), grpN = tid / lanesN, ngrpN = nth / lanesN;
  double a_old = 1.0, gamma_old = 1.0;
  int it = 0;
  bool broke = false;
  // Second stopping rule (eta_e2 > 0): the energy-norm error of CG falls by alpha_k gamma_k per iteration
  // (|e_k|_K^2 = |e_0|_K^2 - sum_{i<k} alpha_i gamma_i), so the solve also goes on while the last decrement is more
  // than eta_e^2 of everything gained so far -- a residual that is small only because K is badly conditioned does
  // not end the solve.  Below the round-off floor nothing goes on.
  double esum = 0.0, elast = 0.0;
  while (rn > thresh_floor && (rn > thresh || (eta_e2 > 0.0 && elast > eta_e2 * esum)) && it < max_it) {
    // ---- phase A: t = A uu, tr = rho .* t
    if (d.m > 0) {
      for (int base = m0; base < m1; base += ngrpA) {
        const int row = base + grpA;
        const bool valid = row < m1;
        Acc<1> acc;
        acc.a[0] = 0.0;
        row_accumulate<1>(d.A.rowptr, d.A.col, d.A.val, row, valid, subA, lanesA, acc,
                          [&](Acc<1> &ac, int c, double a) { ac.a[0] += a * v.uu[c]; });
        group_reduce<1>(acc, lanesA);
        if (valid && subA == 0) {
          v.t[row] = acc.a[0];
          v.tr[row] = rho_vec[row] * acc.a[0];
        }
      }
      grid_barrier(g);
    }
    // ---- phase B: w = P uu + sigma uu + A' tr ; delta = w . uu
    double red1[1] = {0.0};
    for (int base = n0; base < n1; base += ngrpN) {
      const int row = base + grpN;
      const bool valid = row < n1;
      Acc<1> acc;
      acc.a[0] = 0.0;
      row_accumulate<1>(d.P.rowptr, d.P.col, d.P.val, row, valid, subN, lanesN, acc,
                        [&](Acc<1> &ac, int c, double a) { ac.a[0] += a * v.uu[c]; });
      if (d.m > 0)
        row_accumulate<1>(d.At.rowptr, d.At.col, d.At.val, row, valid, subN, lanesN, acc,
                          [&](Acc<1> &ac, int c, double a) { ac.a[0] += a * v.tr[c]; });
      group_reduce<1>(acc, lanesN);
      if (valid && subN == 0) {
        const double uj = v.uu[row];
        const double wj = acc.a[0] + sigma * uj;
        v.w[row] = wj;
        red1[0] += wj * uj;
      }
    }
    reduce_and_barrier<1>(g, sm, red1, 0u);
    const double delta = red1[0];
    double beta, denom;
    if (it == 0) {
      beta = 0.0;
      denom = delta;
    } else {
      beta = gamma / gamma_old;
      denom = delta - beta * gamma / a_old;
    }
    const double alpha = gamma / denom;
    if (!(alpha > 0.0) || !isfinite(alpha)) {  // breakdown: p'Kp <= 0 or exact convergence
      broke = pcg_negative_curvature(delta, denom);
      break;
    }
    // ---- phase V: vector recurrences
    double red2[2] = {0.0, 0.0};
    for (int j = n0 + tid; j < n1; j += nth) {
      const double uj = v.uu[j], wj = v.w[j];
      const double pj = (it == 0) ? uj : uj + beta * v.p[j];
      const double sj = (it == 0) ? wj : wj + beta * v.s[j];
      v.p[j] = pj;
      v.s[j] = sj;
      xvec[j] += alpha * pj;
      const double rj = v.r[j] - alpha * sj;
      v.r[j] = rj;
      const double un = Minv[j] * rj;
      v.uu[j] = un;
      red2[0] += rj * un;
      red2[1] = fmax(red2[1], fabs(rj));
    }
    if (zvec != nullptr) {
      for (int i = m0 + tid; i < m1; i += nth) {
        const double api = (it == 0) ? v.t[i] : v.t[i] + beta * v.Ap[i];
        v.Ap[i] = api;
        zvec[i] += alpha * api;
      }
    }
    pc.tick(6);
    reduce_and_barrier<2>(g, sm, red2, 0x2u);
    pc.tick(7);
    gamma_old = gamma;
    gamma = red2[0];
    rn = red2[1];
    a_old = alpha;
    elast = alpha * gamma_old;  // alpha_k gamma_k of the iteration just finished
    esum += elast;
    it++;
  }
  return broke ? -(it + 1) : it;
}

// Owner-side sum of the per-group partial row sums of a tile-stream phase (fixed order).
__device__ __forceinline__ double part_sum(const TileStreamDev &T, int row) {
  if (!T.split) {
    double a = T.part[row];
    for (int g = 1; g < T.ngroups; g++) a += T.part[(size_t)g * T.srows + row];
    return a;
  }
  double a = 0.0;  // rows cut into pieces: add the pieces of every group in order
  for (int g = 0; g < T.ngroups; g++) {
    const int *sp = T.sr_ptr + (size_t)g * (T.rows + 1) + row;
    for (int sr = __ldg(sp); sr < __ldg(sp + 1); sr++) a += T.part[(size_t)g * T.srows + sr];
  }
  return a;
}

// Same recurrence on the tile streams (engine.cuh TileStreamDev), 4 grid barriers per iteration:
//   phase A : [A; P] uu  -> partial row sums per column group                      | barrier
//   phase C : owners: t = sum of partials, tr = rho .* t, Pu; delta = uu'P uu + sigma |uu|^2 + t'tr | reduce + barrier
//   phase B : A' tr      -> partial row sums per column group                      | barrier
//   phase V : owners: w = Pu + sigma uu + sum of partials, vector recurrences      | reduce + barrier
// kM32: the two stream phases read the fp32 copies of the matrix values (engine.cuh DevPtrs::mat32).
template <bool kM32>
__device__ __noinline__ int pcg_run_stream(Grid &g, RedSmem &sm, Slice &S, PhaseClock &pc, const DevPtrs &d,
                                           const PcgVecs &v, const double *rho_vec, const double *Minv, double sigma, double *xvec,
                                           double *zvec, double gamma, double rn, double thresh, int max_it, int m0,
                                           int m1, int n0, int n1, double thresh_floor = 0.0, double eta_e2 = 0.0) {
  const bool wood = MODE(d.W.w > 0, false);  // Minv is then D^{-1} of the Woodbury preconditioner (wood_refresh, wood_apply)
  const int tid = threadIdx.x, nth = blockDim.x;
  const int m = d.m;
  double a_old = 1.0, gamma_old = 1.0;
  int it = 0;
  bool broke = false;
  // Second stopping rule (eta_e2 > 0): the energy-norm error of CG falls by alpha_k gamma_k per iteration
  // (|e_k|_K^2 = |e_0|_K^2 - sum_{i<k} alpha_i gamma_i), so the solve also goes on while the last decrement is more
  // than eta_e^2 of everything gained so far -- a residual that is small only because K is badly conditioned does
  // not end the solve.  Below the round-off floor nothing goes on.
  double esum = 0.0, elast = 0.0;
  while (rn > thresh_floor && (rn > thresh || (eta_e2 > 0.0 && elast > eta_e2 * esum)) && it < max_it) {
    // beta only needs the gammas of the previous reductions
    const double beta = (it == 0) ? 0.0 : gamma / gamma_old;
    double red1[1] = {0.0};
    if (MODE(d.SA.paired, FAST_PAIRED)) {
      // ---- phase A with the combine fused in (cluster pairs): no partials, no extra grid barrier
      const bool rec = zvec != nullptr && it > 0;
      auto fin = [&](int r, double sum) {
        if (r < m) {
          const double tr = store_tr(d, v.tr, r, rho_vec[r] * sum);
          if (zvec != nullptr) v.Ap[r] = rec ? sum + beta * v.Ap[r] : sum;
          red1[0] += sum * tr;
        } else {
          const double uj = v.uu[r - m];
          d.Pu[r - m] = sum;
          red1[0] += uj * (sum + sigma * uj);
        }
      };
      if (MODE(d.f32_slices, 1)) stream_phase_paired<true, kM32>(S, d.SA, d.uu32, fin);
      else stream_phase_paired<false>(S, d.SA, v.uu, fin);
      if (m > 0) stream_prefetch_head<kM32>(d.ST);
      pc.tick(0);
    } else {
    // ---- phase A
    if (MODE(d.f32_slices, 1)) stream_phase_f32<kM32>(S, d.SA, d.uu32);
    else stream_phase(S, d.SA, v.uu);
    if (m > 0) stream_prefetch_head<kM32>(d.ST);
    pc.tick(0);
    grid_barrier(g);
    pc.tick(1);
    // ---- phase C
    {
      // a block owns ~2 rows of A and ~1 row of P per thread: every load of a round (two rows of A, one of P) is
      // issued before the first dependent store, so the phase costs about one L2 round trip
      const bool rec = zvec != nullptr && it > 0;
      int j = n0 + tid;
      for (int i = m0 + tid; i < m1 || j < n1; i += 2 * nth, j += nth) {
        const int i2 = i + nth;
        const bool h1 = i < m1, h2 = i2 < m1, hj = j < n1;
        double t1 = 0.0, t2 = 0.0, r1 = 0.0, r2 = 0.0, a1 = 0.0, a2 = 0.0, pu = 0.0, uj = 0.0;
        if (h1) { t1 = part_sum(d.SA, i); r1 = rho_vec[i]; if (rec) a1 = v.Ap[i]; }
        if (h2) { t2 = part_sum(d.SA, i2); r2 = rho_vec[i2]; if (rec) a2 = v.Ap[i2]; }
        if (hj) { pu = part_sum(d.SA, m + j); uj = v.uu[j]; }
        if (h1) {
          const double tr = store_tr(d, v.tr, i, r1 * t1);
          if (zvec != nullptr) v.Ap[i] = t1 + beta * a1;  // A p, for the z = A x recurrence (beta = 0 at it 0)
          red1[0] += t1 * tr;
        }
        if (h2) {
          const double tr = store_tr(d, v.tr, i2, r2 * t2);
          if (zvec != nullptr) v.Ap[i2] = t2 + beta * a2;
          red1[0] += t2 * tr;
        }
        if (hj) {
          d.Pu[j] = pu;
          red1[0] += uj * (pu + sigma * uj);
        }
      }
    }
    }
    pc.tick(2);
    if (m > 0) {
      reduce_and_barrier_fx<1>(g, sm, red1, 0u, gamma);  // delta / gamma is a Rayleigh quotient of M^-1 K
      pc.tick(3);
      // ---- phase B
      if (MODE(d.f32_slices, 1)) stream_phase_f32<kM32>(S, d.ST, d.tr32);
      else stream_phase(S, d.ST, v.tr);
      stream_prefetch_head<kM32>(d.SA);
      pc.tick(4);
      grid_barrier(g);
      pc.tick(5);
    } else {
      reduce_and_barrier<1>(g, sm, red1, 0u);
      pc.tick(3);
    }
    const double delta = red1[0];
    const double denom = (it == 0) ? delta : delta - beta * gamma / a_old;
    const double alpha = gamma / denom;
    if (!(alpha > 0.0) || !isfinite(alpha)) {  // breakdown: p'Kp <= 0 or exact convergence
      broke = pcg_negative_curvature(delta, denom);
      break;
    }
    // ---- phase V
    double red2[2] = {0.0, 0.0};
    {
      int i = m0 + tid;
      for (int j = n0 + tid; j < n1 || i < m1; j += nth, i += 2 * nth) {
        const int i2 = i + nth;
        const bool hj = j < n1, h1 = zvec != nullptr && i < m1, h2 = zvec != nullptr && i2 < m1;
        double uj = 0.0, wj = 0.0, pj = 0.0, sj = 0.0, xj = 0.0, rj = 0.0, mj = 0.0;
        double ap1 = 0.0, z1 = 0.0, ap2 = 0.0, z2 = 0.0;
        if (hj) {
          uj = v.uu[j];
          wj = d.Pu[j] + sigma * uj;
          if (m > 0) wj += part_sum(d.ST, j);
          if (it > 0) { pj = v.p[j]; sj = v.s[j]; }
          xj = xvec[j];
          rj = v.r[j];
          mj = Minv[j];
        }
        if (h1) { ap1 = v.Ap[i]; z1 = zvec[i]; }
        if (h2) { ap2 = v.Ap[i2]; z2 = zvec[i2]; }
        if (hj) {
          pj = uj + beta * pj;  // beta = 0 at it 0
          sj = wj + beta * sj;
          v.p[j] = pj;
          v.s[j] = sj;
          xvec[j] = xj + alpha * pj;
          rj -= alpha * sj;
          v.r[j] = rj;
          double un = mj * rj;
          if (wood) d.W.v[j] = un;
          else un = store_u(d, v.uu, j, un);
          red2[0] += rj * un;
          red2[1] = fmax(red2[1], fabs(rj));
        }
        if (h1) zvec[i] = z1 + alpha * ap1;
        if (h2) zvec[i2] = z2 + alpha * ap2;
      }
    }
    pc.tick(6);
    reduce_and_barrier_fx<2>(g, sm, red2, 0x2u, gamma);
    if (wood) {  // the low-rank correction of u = M^{-1} r and the gamma that goes with it
      double redg[1] = {wood_apply(g, sm, d, Minv, v.r, v.uu, n0, n1)};
      reduce_and_barrier<1>(g, sm, redg, 0u);
      red2[0] = redg[0];
    }
    pc.tick(7);
    gamma_old = gamma;
    gamma = red2[0];
    rn = red2[1];
    a_old = alpha;
    elast = alpha * gamma_old;  // alpha_k gamma_k of the iteration just finished
    esum += elast;
    it++;
  }
  return broke ? -(it + 1) : it;
}

// ------------------------------------------------------------------ PCG with slack elimination in the preconditioner
// (engine.cuh SlackDev).  Same Chronopoulos-Gear recurrences as pcg_run_stream; what differs is u = M^-1 r:
//   G : slack rows i (column j):  g_i = rho_i a_ij r_j / K_yy,j                         (fp32 gather vector)   | barrier
//   T2: stream A' against g  -> partials                                                                        | barrier
//   U : x-part:  u_x = diag(S)^-1 (r_x - A_x' g),  gamma_x = r_x'u_x                  (Minv holds diag(S)^-1, 0 on Y) | barrier
//   A : stream [A; P] against u_x (slack entries of the gather vector are zero); then per row
//         slack row i:  u_j = (r_j - rho_i a_ij t_i) / K_yy,j ;  t_i += a_ij u_j      (u_y by back-substitution)
//         all rows:     tr = rho .* t ;  delta = u'Pu + sigma |u|^2 + t'tr ;  gamma += r_y'u_y              | reduce + barrier
//   B : stream A' against tr -> partials                                                                      | barrier
//   V : w = Pu + sigma u + partials ;  p, s, x, r ;  Ap = t + beta Ap ;  z += alpha Ap ;  |r|inf            | reduce + barrier
// The [A; P] phase of the preconditioner IS the A phase of the K-apply, so the preconditioner costs one extra A'
// phase.  Not used by the polish (its penalties are not rho).
template <bool kM32>
__device__ __noinline__ int pcg_run_stream_slack(Grid &g, RedSmem &sm, Slice &S, PhaseClock &pc, const DevPtrs &d,
                                                 const PcgVecs &v, const double *rho_vec, const double *Minv, double sigma,
                                                 double *xvec, double *zvec, double rn, double thresh, int max_it, int m0,
                                                 int m1, int n0, int n1, double thresh_floor = 0.0, double eta_e2 = 0.0) {
  const int tid = threadIdx.x, nth = blockDim.x;
  const int m = d.m;
  const SlackDev &L = d.SL;
  double a_old = 1.0, gamma_old = 1.0;
  int it = 0;
  bool broke = false;
  // per-row finish of the [A; P] phase (row sums complete): returns this row's share of {delta, gamma}
  auto finish = [&](int r, double sum, double &dl, double &gm) {
    if (r < m) {
      double t = sum;
      const int j = __ldg(L.col + r);
      if (j >= 0) {
        const double a = d.A.val[__ldg(L.pos + r)], ri = rho_vec[r], pjj = d.Pdiag[j];
        const double kyy = pjj + sigma + ri * a * a;
        const double rj = v.r[j];
        const double uy = (rj - ri * a * t) / kyy;
        v.uu[j] = uy;
        d.Pu[j] = pjj * uy;
        t = fma(a, uy, t);
        dl += uy * (pjj * uy + sigma * uy);
        gm += rj * uy;
      }
      const double tr = store_tr(d, v.tr, r, rho_vec[r] * t);
      v.t[r] = t;
      dl += t * tr;
    } else {
      const int j = r - m;
      if (Minv[j] != 0.0) {  // x-part; the slack columns were handled by their rows
        const double uj = v.uu[j];
        d.Pu[j] = sum;
        dl += uj * (sum + sigma * uj);
      }
    }
  };
  // Second stopping rule (eta_e2 > 0): the energy-norm error of CG falls by alpha_k gamma_k per iteration
  // (|e_k|_K^2 = |e_0|_K^2 - sum_{i<k} alpha_i gamma_i), so the solve also goes on while the last decrement is more
  // than eta_e^2 of everything gained so far -- a residual that is small only because K is badly conditioned does
  // not end the solve.  Below the round-off floor nothing goes on.
  double esum = 0.0, elast = 0.0;
  while (rn > thresh_floor && (rn > thresh || (eta_e2 > 0.0 && elast > eta_e2 * esum)) && it < max_it) {
    // ---- G
    for (int i = m0 + tid; i < m1; i += nth) {
      const int j = __ldg(L.col + i);
      if (j >= 0) {
        const double a = d.A.val[__ldg(L.pos + i)], ri = rho_vec[i];
        const double kyy = d.Pdiag[j] + sigma + ri * a * a;
        L.g32[i] = (float)(ri * a * v.r[j] / kyy);
      }
    }
    grid_barrier(g);
    // ---- T2
    stream_phase_f32<kM32>(S, d.ST, L.g32);
    grid_barrier(g);
    // ---- U
    double red[2] = {0.0, 0.0};  // delta, gamma
    for (int j = n0 + tid; j < n1; j += nth) {
      const double mj = Minv[j];
      if (mj != 0.0) {
        const double rj = v.r[j];
        const double uj = store_u(d, v.uu, j, mj * (rj - part_sum(d.ST, j)));
        red[1] += rj * uj;
      }
    }
    stream_prefetch_head<kM32>(d.SA);
    grid_barrier(g);
    pc.tick(7);
    // ---- A
    if (MODE(d.SA.paired, FAST_PAIRED)) {
      auto fin = [&](int r, double sum) { finish(r, sum, red[0], red[1]); };
      stream_phase_paired<true, kM32>(S, d.SA, d.uu32, fin);
      pc.tick(0);
    } else {
      stream_phase_f32<kM32>(S, d.SA, d.uu32);
      pc.tick(0);
      grid_barrier(g);
      pc.tick(1);
      for (int i = m0 + tid; i < m1; i += nth) finish(i, part_sum(d.SA, i), red[0], red[1]);
      for (int j = n0 + tid; j < n1; j += nth) finish(m + j, part_sum(d.SA, m + j), red[0], red[1]);
    }
    stream_prefetch_head<kM32>(d.ST);
    pc.tick(2);
    reduce_and_barrier<2>(g, sm, red, 0u);
    pc.tick(3);
    const double delta = red[0], gamma = red[1];
    const double beta = (it == 0) ? 0.0 : gamma / gamma_old;
    // ---- B
    stream_phase_f32<kM32>(S, d.ST, d.tr32);
    stream_prefetch_head<kM32>(d.SA);
    pc.tick(4);
    grid_barrier(g);
    pc.tick(5);
    const double denom = (it == 0) ? delta : delta - beta * gamma / a_old;
    const double alpha = gamma / denom;
    if (!(alpha > 0.0) || !isfinite(alpha)) {
      broke = pcg_negative_curvature(delta, denom);
      break;
    }
    // ---- V
    double red2[1] = {0.0};
    for (int j = n0 + tid; j < n1; j += nth) {
      const double uj = v.uu[j];
      const double wj = d.Pu[j] + sigma * uj + part_sum(d.ST, j);
      const double pj = (it > 0) ? uj + beta * v.p[j] : uj;
      const double sj = (it > 0) ? wj + beta * v.s[j] : wj;
      v.p[j] = pj;
      v.s[j] = sj;
      xvec[j] += alpha * pj;
      const double rj = v.r[j] - alpha * sj;
      v.r[j] = rj;
      red2[0] = fmax(red2[0], fabs(rj));
    }
    if (zvec != nullptr)
      for (int i = m0 + tid; i < m1; i += nth) {
        const double api = (it > 0) ? v.t[i] + beta * v.Ap[i] : v.t[i];
        v.Ap[i] = api;
        zvec[i] += alpha * api;
      }
    pc.tick(6);
    reduce_and_barrier<1>(g, sm, red2, 0x1u);
    gamma_old = gamma;
    rn = red2[0];
    a_old = alpha;
    elast = alpha * gamma_old;  // alpha_k gamma_k of the iteration just finished
    esum += elast;
    it++;
  }
  return broke ? -(it + 1) : it;
}

// Data of the slack preconditioner for the current rho: reduced weights, diag(S)^-1 (zero on the slack columns), and a
// clean fp32 gather copy of u on the slack columns.  Entered and left by every thread of the grid.
__device__ __noinline__ void slack_refresh(Grid &g, const DevPtrs &d, const double *rho_vec, double sigma, double *Minv,
                                           int m0, int m1, int n0, int n1);

// ------------------------------------------------------------------ update_info (row a9) + infeasibility products (row a10)
// One phase streams A, P and A' once each with two gathered vectors per matrix:
//   A:(x, dx) -> Ax, A dx | P:(x, dx) -> Px, P dx | A':(y, dy) -> A'y, A'dy
__device__ __noinline__ void compute_info(Grid &g, RedSmem &sm, const DevPtrs &d, const SolveCfg &c, double cost_c,
                                          double cost_cinv, const double *xv, const double *zv, const double *yv,
                                          int m0, int m1, int n0, int n1, InfoScalars &S) {
  const int tid = threadIdx.x, nth = blockDim.x;
  const bool unscale = c.scaling && !c.scaled_termination;
  const int lanesA = d.A.lanes, lanesN = d.At.lanes;
  const int subA = tid & (lanesA - 1), grpA = tid / lanesA, ngrpA = nth / lanesA;
  const int subN = tid & (lanesN - 1), grpN = tid / lanesN, ngrpN = nth / lanesN;
  double v[23];
#pragma unroll
  for (int k = 0; k < 23; k++) v[k] = 0.0;
  v[8] = -INFINITY;
  v[9] = -INFINITY;
  // ---- m rows
  for (int base = m0; base < m1; base += ngrpA) {
    const int row = base + grpA;
    const bool valid = row < m1;
    Acc<2> acc;
    acc.a[0] = acc.a[1] = 0.0;
    row_accumulate<2>(d.A.rowptr, d.A.col, d.A.val, row, valid, subA, lanesA, acc, [&](Acc<2> &ac, int cc, double a) {
      ac.a[0] += a * xv[cc];
      ac.a[1] += a * d.dx[cc];
    });
    group_reduce<2>(acc, lanesA);
    if (valid && subA == 0) {
      const double Ax = acc.a[0], Adx = acc.a[1];
      const double zi = zv[row], ei = unscale ? d.Einv[row] : 1.0, Ei = unscale ? d.E[row] : 1.0;
      const double li = d.l[row], ui = d.u[row], dyi = d.dy[row];
      const double pr = fabs(Ax - zi);
      v[0] = fmax(v[0], ei * pr);
      v[1] = fmax(v[1], pr);
      v[2] = fmax(v[2], ei * fabs(zi));
      v[3] = fmax(v[3], fabs(zi));
      v[4] = fmax(v[4], ei * fabs(Ax));
      v[5] = fmax(v[5], fabs(Ax));
      v[6] = fmax(v[6], Ei * fabs(dyi));
      v[7] += ui * fmax(dyi, 0.0) + li * fmin(dyi, 0.0);
      const double adx = ei * Adx;
      if (ui < kInfty * kMinScaling) v[8] = fmax(v[8], adx);
      if (li > -kInfty * kMinScaling) v[9] = fmax(v[9], -adx);
    }
  }
  // ---- n rows
  for (int base = n0; base < n1; base += ngrpN) {
    const int row = base + grpN;
    const bool valid = row < n1;
    Acc<4> acc;
    acc.a[0] = acc.a[1] = acc.a[2] = acc.a[3] = 0.0;
    row_accumulate<4>(d.P.rowptr, d.P.col, d.P.val, row, valid, subN, lanesN, acc, [&](Acc<4> &ac, int cc, double a) {
      ac.a[0] += a * xv[cc];
      ac.a[1] += a * d.dx[cc];
    });
    if (d.m > 0)
      row_accumulate<4>(d.At.rowptr, d.At.col, d.At.val, row, valid, subN, lanesN, acc,
                        [&](Acc<4> &ac, int cc, double a) {
                          ac.a[2] += a * yv[cc];
                          ac.a[3] += a * d.dy[cc];
                        });
    group_reduce<4>(acc, lanesN);
    if (valid && subN == 0) {
      const double Px = acc.a[0], Pdx = acc.a[1], Aty = acc.a[2], Atdy = acc.a[3];
      const double qj = d.q[row], xj = xv[row], dxj = d.dx[row];
      const double di = unscale ? d.Dinv[row] : 1.0, Dj = unscale ? d.D[row] : 1.0;
      const double dr = fabs(qj + Px + Aty);
      v[10] = fmax(v[10], di * dr);
      v[11] = fmax(v[11], dr);
      v[12] = fmax(v[12], di * fabs(qj));
      v[13] = fmax(v[13], fabs(qj));
      v[14] = fmax(v[14], di * fabs(Aty));
      v[15] = fmax(v[15], fabs(Aty));
      v[16] = fmax(v[16], di * fabs(Px));
      v[17] = fmax(v[17], fabs(Px));
      v[18] += xj * (0.5 * Px + qj);
      v[19] = fmax(v[19], Dj * fabs(dxj));
      v[20] += qj * dxj;
      v[21] = fmax(v[21], di * fabs(Pdx));
      v[22] = fmax(v[22], di * fabs(Atdy));
    }
  }
  // sums: 7 (lhs), 18 (obj), 20 (q'dx); everything else is a max
  reduce_and_barrier<23>(g, sm, v, 0x7FFFFFu & ~((1u << 7) | (1u << 18) | (1u << 20)));
  S.pri_t = v[0]; S.pri_r = v[1]; S.nz_t = v[2]; S.nz_r = v[3]; S.nAx_t = v[4]; S.nAx_r = v[5];
  S.ndy_t = v[6]; S.lhs = v[7]; S.maxU_t = v[8]; S.maxNegL_t = v[9];
  S.dua_t = v[10]; S.dua_r = v[11]; S.nq_t = v[12]; S.nq_r = v[13]; S.nAty_t = v[14]; S.nAty_r = v[15];
  S.nPx_t = v[16]; S.nPx_r = v[17]; S.obj = v[18]; S.ndx_t = v[19]; S.qdx = v[20]; S.nPdx_t = v[21];
  S.nAtdy_t = v[22];
  S.obj_val = c.scaling ? S.obj * cost_cinv : S.obj;
  S.pri_res = (d.m == 0) ? 0.0 : S.pri_t;
  S.dua_res = unscale ? cost_cinv * S.dua_t : S.dua_t;
}

// The same scalars on the tile streams (problems large enough to have them): four stream phases -- [A; P] x, A' y,
// [A; P] dx, A' dy -- leave the six products in PCG scratch vectors that are free between two solves (t, Ap: m;
// p, s, w, uu: n), then the owners form the 23 scalars.  Must be entered after a grid barrier that made xv, zv, yv,
// dx and dy visible.
__device__ __noinline__ void compute_info_stream(Grid &g, RedSmem &sm, Slice &SG, const DevPtrs &d, const SolveCfg &c,
                                                 double cost_c, double cost_cinv, const double *xv, const double *zv,
                                                 const double *yv, int m0, int m1, int n0, int n1, InfoScalars &S) {
  const int tid = threadIdx.x, nth = blockDim.x, m = d.m;
  const bool unscale = c.scaling && !c.scaled_termination;
  auto products = [&](const double *vn, const double *vm, double *outA, double *outP, double *outT) {
    // outA = A vn (m), outP = P vn (n), outT = A' vm (n)
    if (MODE(d.SA.paired, FAST_PAIRED)) {
      stream_phase_paired(SG, d.SA, vn, [&](int r, double sum) {
        if (r < m) outA[r] = sum;
        else outP[r - m] = sum;
      });
    } else {
      stream_phase(SG, d.SA, vn);
      grid_barrier(g);
      for (int i = m0 + tid; i < m1; i += nth) outA[i] = part_sum(d.SA, i);
      for (int j = n0 + tid; j < n1; j += nth) outP[j] = part_sum(d.SA, m + j);
    }
    grid_barrier(g);
    if (m > 0) {
      stream_phase(SG, d.ST, vm);
      grid_barrier(g);
      for (int j = n0 + tid; j < n1; j += nth) outT[j] = part_sum(d.ST, j);
      grid_barrier(g);  // the partials are rewritten by the next A' phase
    }
  };
  products(xv, yv, d.t, d.p, d.w);
  products(d.dx, d.dy, d.Ap, d.s, d.uu);
  double v[23];
#pragma unroll
  for (int k = 0; k < 23; k++) v[k] = 0.0;
  v[8] = -INFINITY;
  v[9] = -INFINITY;
  for (int row = m0 + tid; row < m1; row += nth) {
    const double Ax = d.t[row], Adx = d.Ap[row];
    const double zi = zv[row], ei = unscale ? d.Einv[row] : 1.0, Ei = unscale ? d.E[row] : 1.0;
    const double li = d.l[row], ui = d.u[row], dyi = d.dy[row];
    const double pr = fabs(Ax - zi);
    v[0] = fmax(v[0], ei * pr);
    v[1] = fmax(v[1], pr);
    v[2] = fmax(v[2], ei * fabs(zi));
    v[3] = fmax(v[3], fabs(zi));
    v[4] = fmax(v[4], ei * fabs(Ax));
    v[5] = fmax(v[5], fabs(Ax));
    v[6] = fmax(v[6], Ei * fabs(dyi));
    v[7] += ui * fmax(dyi, 0.0) + li * fmin(dyi, 0.0);
    const double adx = ei * Adx;
    if (ui < kInfty * kMinScaling) v[8] = fmax(v[8], adx);
    if (li > -kInfty * kMinScaling) v[9] = fmax(v[9], -adx);
  }
  for (int row = n0 + tid; row < n1; row += nth) {
    const double Px = d.p[row], Pdx = d.s[row], Aty = (m > 0) ? d.w[row] : 0.0, Atdy = (m > 0) ? d.uu[row] : 0.0;
    const double qj = d.q[row], xj = xv[row], dxj = d.dx[row];
    const double di = unscale ? d.Dinv[row] : 1.0, Dj = unscale ? d.D[row] : 1.0;
    const double dr = fabs(qj + Px + Aty);
    v[10] = fmax(v[10], di * dr);
    v[11] = fmax(v[11], dr);
    v[12] = fmax(v[12], di * fabs(qj));
    v[13] = fmax(v[13], fabs(qj));
    v[14] = fmax(v[14], di * fabs(Aty));
    v[15] = fmax(v[15], fabs(Aty));
    v[16] = fmax(v[16], di * fabs(Px));
    v[17] = fmax(v[17], fabs(Px));
    v[18] += xj * (0.5 * Px + qj);
    v[19] = fmax(v[19], Dj * fabs(dxj));
    v[20] += qj * dxj;
    v[21] = fmax(v[21], di * fabs(Pdx));
    v[22] = fmax(v[22], di * fabs(Atdy));
  }
  reduce_and_barrier<23>(g, sm, v, 0x7FFFFFu & ~((1u << 7) | (1u << 18) | (1u << 20)));
  S.pri_t = v[0]; S.pri_r = v[1]; S.nz_t = v[2]; S.nz_r = v[3]; S.nAx_t = v[4]; S.nAx_r = v[5];
  S.ndy_t = v[6]; S.lhs = v[7]; S.maxU_t = v[8]; S.maxNegL_t = v[9];
  S.dua_t = v[10]; S.dua_r = v[11]; S.nq_t = v[12]; S.nq_r = v[13]; S.nAty_t = v[14]; S.nAty_r = v[15];
  S.nPx_t = v[16]; S.nPx_r = v[17]; S.obj = v[18]; S.ndx_t = v[19]; S.qdx = v[20]; S.nPdx_t = v[21];
  S.nAtdy_t = v[22];
  S.obj_val = c.scaling ? S.obj * cost_cinv : S.obj;
  S.pri_res = (d.m == 0) ? 0.0 : S.pri_t;
  S.dua_res = unscale ? cost_cinv * S.dua_t : S.dua_t;
}

// Minv = 1 / (P_jj + sigma + sum_i rho_i A_ij^2) on rows [n0, n1); rows of A with skip[i] >= 0 (members of the Woodbury
// set, engine.cuh WoodDev) are left out of the sum when `skip` is given
__device__ __forceinline__ void precond_rows(const DevPtrs &d, const double *rho_vec, double sigma, double *Minv,
                                             int n0, int n1, const int *skip = nullptr) {
  const int tid = threadIdx.x, nth = blockDim.x;
  const int lanesN = d.At.lanes;
  const int subN = tid & (lanesN - 1), grpN = tid / lanesN, ngrpN = nth / lanesN;
  for (int base = n0; base < n1; base += ngrpN) {
    const int row = base + grpN;
    const bool valid = row < n1;
    Acc<1> acc;
    acc.a[0] = 0.0;
    if (d.m > 0) {
      if (skip == nullptr)
        row_accumulate<1>(d.At.rowptr, d.At.col, d.At.val, row, valid, subN, lanesN, acc,
                          [&](Acc<1> &ac, int c, double a) { ac.a[0] += rho_vec[c] * a * a; });
      else
        row_accumulate<1>(d.At.rowptr, d.At.col, d.At.val, row, valid, subN, lanesN, acc,
                          [&](Acc<1> &ac, int c, double a) { if (__ldg(skip + c) < 0) ac.a[0] += rho_vec[c] * a * a; });
    }
    group_reduce<1>(acc, lanesN);
    if (valid && subN == 0) Minv[row] = 1.0 / (d.Pdiag[row] + sigma + acc.a[0]);
  }
}

__device__ __noinline__ void slack_refresh(Grid &g, const DevPtrs &d, const double *rho_vec, double sigma, double *Minv,
                                           int m0, int m1, int n0, int n1) {
  const int tid = threadIdx.x, nth = blockDim.x;
  const SlackDev &L = d.SL;
  grid_barrier(g);  // rho_vec is complete
  for (int i = m0 + tid; i < m1; i += nth) {
    const int j = L.col[i];
    double re = rho_vec[i];
    if (j >= 0) {
      const double a = d.A.val[L.pos[i]], pj = d.Pdiag[j] + sigma;
      re = re * pj / (pj + re * a * a);
    } else {
      L.g32[i] = 0.0f;
    }
    L.rho_eff[i] = re;
  }
  grid_barrier(g);
  precond_rows(d, L.rho_eff, sigma, Minv, n0, n1);
  grid_barrier(g);
  for (int i = m0 + tid; i < m1; i += nth) {
    const int j = L.col[i];
    if (j >= 0) {
      Minv[j] = 0.0;      // marks the slack columns for pcg_run_stream_slack
      d.uu32[j] = 0.0f;   // and they never enter the gathered u
    }
  }
  grid_barrier(g);
}

// ------------------------------------------------------------------ Woodbury part of the preconditioner (engine.cuh WoodDev)
// Sum of one value per thread over the block, fixed order (warp shuffle tree, then the warps in order).
__device__ __forceinline__ double block_sum(RedSmem &sm, double a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) sm.part[warp][0] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < nwarps; k++) t += sm.part[k][0];
    sm.res[0] = t;
  }
  __syncthreads();
  const double r = sm.res[0];
  __syncthreads();
  return r;
}

// Cholesky factor of C (w x w, lower triangle, in place) and the explicit inverse Cinv = C^{-1}; ONE thread block.
// C = I + (positive semidefinite), so the pivots are >= 1 up to rounding; a pivot that is not positive (never seen)
// is replaced by 1, which only weakens the preconditioner.
__device__ __noinline__ void wood_factor(const WoodDev &W) {
  const int w = W.w, ld = W.ld, tid = threadIdx.x, nth = blockDim.x;
  double *C = W.C, *X = W.Cinv;
  __shared__ double dk;
  for (int k = 0; k < w; k++) {
    if (tid == 0) {
      const double p = C[k * ld + k];
      dk = p > 0.0 ? sqrt(p) : 1.0;
      C[k * ld + k] = dk;
    }
    __syncthreads();
    const double inv = 1.0 / dk;
    for (int i = k + 1 + tid; i < w; i += nth) C[i * ld + k] *= inv;
    __syncthreads();
    const int R = w - k - 1;
    for (int e = tid; e < R * R; e += nth) {
      const int i = k + 1 + e / R, j = k + 1 + e % R;
      if (j <= i) C[i * ld + j] -= C[i * ld + k] * C[j * ld + k];
    }
    __syncthreads();
  }
  // column a of the inverse by thread a: L y = e_a, L' x = y, in place in X[:, a]; every thread walks the same
  // (i, k), so the factor loads are broadcasts and the X loads are coalesced
  for (int a = tid; a < w; a += nth) {
    for (int i = 0; i < w; i++) {
      double acc = (i == a) ? 1.0 : 0.0;
      for (int k = 0; k < i; k++) acc -= C[i * ld + k] * X[k * ld + a];
      X[i * ld + a] = acc / C[i * ld + i];
    }
    for (int i = w - 1; i >= 0; i--) {
      double acc = X[i * ld + a];
      for (int k = i + 1; k < w; k++) acc -= C[k * ld + i] * X[k * ld + a];
      X[i * ld + a] = acc / C[i * ld + i];
    }
  }
  __syncthreads();
}

// Rebuild the Woodbury data for the penalty vector `pen` (rho_vec of the ADMM loop, or the polish penalties) and the
// diagonal shift `sig`: Dinv (without the member rows), s = sqrt(pen_W), C = I + S A_W Dinv A_W' S, Cinv.  C is
// assembled kWoodCols columns per round: the scaled member rows are scattered into dense scratch vectors and every
// member row takes its dot products with them (one thread block per row, fixed-order block sums).
// Entered and left by every thread of the grid; starts and ends with a grid barrier.
__device__ __noinline__ void wood_refresh(Grid &g, RedSmem &sm, const DevPtrs &d, const double *pen, double sig,
                                          double *Dinv, int n0, int n1) {
  const WoodDev &W = d.W;
  const int tid = threadIdx.x, nth = blockDim.x, b = blockIdx.x, nb = gridDim.x;
  const int w = W.w, ld = W.ld;
  const size_t np = (size_t)d.n + 8;
  grid_barrier(g);  // pen is complete
  for (int a = b * nth + tid; a < w; a += nb * nth) W.s[a] = sqrt(fmax(pen[W.rows[a]], 0.0));
  precond_rows(d, pen, sig, Dinv, n0, n1, W.idx);
  grid_barrier(g);
  for (int a0 = 0; a0 < w; a0 += kWoodCols) {
    const int nc = min(kWoodCols, w - a0);
    for (int c = 0; c < nc; c++)
      for (int j = n0 + tid; j < n1; j += nth) W.vk[c * np + j] = 0.0;
    grid_barrier(g);
    for (int c = 0; c < nc; c++) {
      const int a = a0 + c;
      if (a % nb != b) continue;
      const double sa = W.s[a];
      for (int k = W.rp[a] + tid; k < W.rp[a + 1]; k += nth) {
        const int col = W.ci[k];
        W.vk[c * np + col] = Dinv[col] * W.val[k] * sa;
      }
    }
    grid_barrier(g);
    for (int r = b; r < w; r += nb) {
      double acc[kWoodCols];
#pragma unroll
      for (int c = 0; c < kWoodCols; c++) acc[c] = 0.0;
      for (int k = W.rp[r] + tid; k < W.rp[r + 1]; k += nth) {
        const int col = W.ci[k];
        const double a = W.val[k];
#pragma unroll
        for (int c = 0; c < kWoodCols; c++)
          if (c < nc) acc[c] = fma(a, W.vk[c * np + col], acc[c]);
      }
      const double sr = W.s[r];
      for (int c = 0; c < nc; c++) {
        const double sum = block_sum(sm, acc[c]);
        if (tid == 0) W.C[r * ld + a0 + c] = sr * sum + (r == a0 + c ? 1.0 : 0.0);
      }
    }
    grid_barrier(g);
  }
  if (b == 0) wood_factor(W);
  grid_barrier(g);
}

// u = M^{-1} r for the Woodbury preconditioner, given v = Dinv .* r in W.v (written by the caller, made visible by a
// grid barrier).  Returns this thread's share of gamma = r'u; the caller reduces it (its barrier also publishes u).
__device__ __noinline__ double wood_apply(Grid &g, RedSmem &sm, const DevPtrs &d, const double *Dinv,
                                          const double *rvec, double *uvec, int n0, int n1) {
  const WoodDev &W = d.W;
  const int tid = threadIdx.x, nth = blockDim.x, b = blockIdx.x, nb = gridDim.x;
  const int lane = tid & 31, warp = tid >> 5, w = W.w, ld = W.ld;
  // t = S A_W v: one thread block per member row (these rows are long)
  for (int r = b; r < w; r += nb) {
    double acc = 0.0;
    for (int k = W.rp[r] + tid; k < W.rp[r + 1]; k += nth) acc = fma(W.val[k], W.v[W.ci[k]], acc);
    const double sum = block_sum(sm, acc);
    if (tid == 0) W.t[r] = W.s[r] * sum;
  }
  grid_barrier(g);
  // g = S Cinv t: one warp per entry, spread over the grid
  for (int a = b * kWarps + warp; a < w; a += nb * kWarps) {
    double acc = 0.0;
    for (int k = lane; k < w; k += 32) acc = fma(W.Cinv[a * ld + k], __ldcg(W.t + k), acc);
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) W.g[a] = W.s[a] * acc;
  }
  grid_barrier(g);
  // u = v - Dinv .* (A_W' g) on the rows this block owns: 16 lanes per row
  constexpr int kL = 16;
  const int sub = tid & (kL - 1), grp = tid / kL, ngrp = nth / kL;
  double gam = 0.0;
  for (int base = n0; base < n1; base += ngrp) {
    const int row = base + grp;
    const bool valid = row < n1;
    double acc = 0.0;
    if (valid)
      for (int k = W.trp[row] + sub; k < W.trp[row + 1]; k += kL) acc = fma(W.tval[k], __ldcg(W.g + W.tci[k]), acc);
    for (int o = kL >> 1; o; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o, kL);
    if (valid && sub == 0) {
      const double u = store_u(d, uvec, row, W.v[row] - Dinv[row] * acc);
      gam += rvec[row] * u;
    }
  }
  return gam;
}

// Residual refresh on the tile streams: z_tilde = A x_tilde, tr = rho .* z_tilde and
// K x_tilde = P x_tilde + sigma x_tilde + A' tr (left in d.w).  Kept out of line so that its registers do not
// weigh on the loop body of admm_kernel.
__device__ __noinline__ void refresh_products_stream(Grid &g, Slice &SG, const DevPtrs &d, double sigma, int m0, int m1,
                                                     int n0, int n1) {
  const int tid = threadIdx.x, nth = blockDim.x, m = d.m;
  if (MODE(d.SA.paired, FAST_PAIRED)) {
    stream_phase_paired(SG, d.SA, d.xt, [&](int r, double sum) {
      if (r < m) {
        d.zt[r] = sum;
        d.tr[r] = d.rho_vec[r] * sum;
      } else {
        d.Pu[r - m] = sum;
      }
    });
  } else {
    stream_phase(SG, d.SA, d.xt);
    grid_barrier(g);
    for (int i = m0 + tid; i < m1; i += nth) {
      const double ti = part_sum(d.SA, i);
      d.zt[i] = ti;
      d.tr[i] = d.rho_vec[i] * ti;
    }
    for (int j = n0 + tid; j < n1; j += nth) d.Pu[j] = part_sum(d.SA, m + j);
  }
  grid_barrier(g);
  if (m > 0) {
    stream_phase(SG, d.ST, d.tr);
    grid_barrier(g);
  }
  for (int j = n0 + tid; j < n1; j += nth) d.w[j] = d.Pu[j] + sigma * d.xt[j] + (m > 0 ? part_sum(d.ST, j) : 0.0);
}

// ------------------------------------------------------------------ the ADMM kernel
__global__ void __launch_bounds__(kThreads, 1) admm_kernel(const __grid_constant__ DevPtrs d, const __grid_constant__ SolveCfg c) {
  __shared__ RedSmem sm;
  __shared__ double phase_acc[kPhases];
  PhaseClock pc;
  pc.start(phase_acc);
  Grid g;
  grid_init(g, d);
  Slice SG;
  slice_init(SG, d);
  const int tid = threadIdx.x, nth = blockDim.x, b = blockIdx.x;
  const int m0 = d.m_start[b], m1 = d.m_start[b + 1], n0 = d.n_start[b], n1 = d.n_start[b + 1];
  const int lanesA = d.A.lanes, lanesN = d.At.lanes;
  const int subA = tid & (lanesA - 1), grpA = tid / lanesA, ngrpA = nth / lanesA;
  const int subN = tid & (lanesN - 1), grpN = tid / lanesN, ngrpN = nth / lanesN;

  double rho = d.state->rho;
  const double cost_c = d.state->c, cost_cinv = d.state->cinv;
  long long interval = d.state->adaptive_interval;
  long long rho_updates = d.state->rho_updates;
  int refresh = 1;  // first PCG of a launch always rebuilds z_tilde and the residual from scratch
  const PcgVecs pv{d.r, d.uu, d.p, d.s, d.w, d.t, d.tr, d.Ap};
  const unsigned long long t_start = globaltimer_ns();

  if (!c.warm_start) {  // cold_start
    for (int j = n0 + tid; j < n1; j += nth) { d.x[j] = 0.0; d.xt[j] = 0.0; }
    for (int i = m0 + tid; i < m1; i += nth) { d.z[i] = 0.0; d.y[i] = 0.0; d.zt[i] = 0.0; }
  }
  grid_barrier(g);

  const bool wood = MODE(d.blocked && d.W.w > 0, false);
  const bool slack = MODE(d.blocked && d.SL.rows > 0, FAST_SLACK);
  if (wood && c.wood_refresh) wood_refresh(g, sm, d, d.rho_vec, c.sigma, d.Minv, n0, n1);
  if (slack && c.wood_refresh) slack_refresh(g, d, d.rho_vec, c.sigma, d.Minv, m0, m1, n0, n1);

  InfoScalars S;
  S.pri_res = S.dua_res = S.obj_val = 0.0;
  long long status = ST_UNSOLVED, info_iter = 0, cg_total = 0, cg_solves = 0, checks = 0, log_rows = 0, refreshes = 0;
  double rho_est = rho, elapsed = 0.0;
  bool can_check = false, can_print = false, wv_valid = false, pcg_broke = false;
  long long it;
  for (it = 1; it <= c.max_iter; it++) {
    // ---- P1: wv = rho .* z - y   (+ refresh: z_tilde = A x_tilde).  In steady state wv was already written by
    //         the Z phase of the previous iteration and its reduce + barrier made it visible: nothing to do here.
    const bool refresh_on_streams = refresh && MODE(d.blocked, 1) && MODE(d.info_streams, 1);
    if (refresh || !wv_valid) {
    for (int i = m0 + tid; i < m1; i += nth) d.wv[i] = d.rho_vec[i] * d.z[i] - d.y[i];
    wv_valid = true;
    if (refresh_on_streams) {
      refresh_products_stream(g, SG, d, c.sigma, m0, m1, n0, n1);
    } else if (refresh && d.m > 0) {
      for (int base = m0; base < m1; base += ngrpA) {
        const int row = base + grpA;
        const bool valid = row < m1;
        Acc<1> acc;
        acc.a[0] = 0.0;
        row_accumulate<1>(d.A.rowptr, d.A.col, d.A.val, row, valid, subA, lanesA, acc,
                          [&](Acc<1> &ac, int cc, double a) { ac.a[0] += a * d.xt[cc]; });
        group_reduce<1>(acc, lanesA);
        if (valid && subA == 0) {
          d.zt[row] = acc.a[0];
          d.tr[row] = d.rho_vec[row] * acc.a[0];
        }
      }
    }
    grid_barrier(g);
    pc.tick(refresh ? 11 : 8);
    }
    // ---- P2: b = sigma x - q + A' wv ; r = b - K x_tilde (refresh) or r += b - b_old
    double red3[3] = {0.0, 0.0, 0.0};
    if (MODE(d.blocked, 1) && (!refresh || refresh_on_streams)) {
      // A' wv through the staged tiles, then the element-wise part on the owner block (a refresh rebuilds the
      // residual from K x_tilde in d.w, the steady state carries it by recurrence)
      if (d.m > 0) {
        stream_phase(SG, d.ST, d.wv);
        stream_prefetch_head<MODE(false, true)>(d.SA);
        grid_barrier(g);
        pc.tick(9);
      }
      for (int j = n0 + tid; j < n1; j += nth) {
        const double acc = (d.m > 0) ? part_sum(d.ST, j) : 0.0;
        const double bj = c.sigma * d.x[j] - d.q[j] + acc;
        const double rj = refresh ? bj - d.w[j] : d.r[j] + (bj - d.b[j]);
        d.b[j] = bj;
        d.r[j] = rj;
        double uj = d.Minv[j] * rj;
        if (wood) d.W.v[j] = uj;
        else if (!slack) uj = store_u(d, d.uu, j, uj);  // slack mode: u is formed at the top of the PCG iteration
        red3[0] += rj * uj;
        red3[1] = fmax(red3[1], fabs(rj));
        red3[2] = fmax(red3[2], fabs(bj));
      }
    } else
    for (int base = n0; base < n1; base += ngrpN) {
      const int row = base + grpN;
      const bool valid = row < n1;
      Acc<2> acc;
      acc.a[0] = acc.a[1] = 0.0;
      if (refresh) {
        row_accumulate<2>(d.P.rowptr, d.P.col, d.P.val, row, valid, subN, lanesN, acc,
                          [&](Acc<2> &ac, int cc, double a) { ac.a[1] += a * d.xt[cc]; });
        if (d.m > 0)
          row_accumulate<2>(d.At.rowptr, d.At.col, d.At.val, row, valid, subN, lanesN, acc,
                            [&](Acc<2> &ac, int cc, double a) {
                              ac.a[0] += a * d.wv[cc];
                              ac.a[1] += a * d.tr[cc];
                            });
      } else if (d.m > 0) {
        row_accumulate<2>(d.At.rowptr, d.At.col, d.At.val, row, valid, subN, lanesN, acc,
                          [&](Acc<2> &ac, int cc, double a) { ac.a[0] += a * d.wv[cc]; });
      }
      group_reduce<2>(acc, lanesN);
      if (valid && subN == 0) {
        const double bj = c.sigma * d.x[row] - d.q[row] + acc.a[0];
        double rj;
        if (refresh) rj = bj - (acc.a[1] + c.sigma * d.xt[row]);
        else rj = d.r[row] + (bj - d.b[row]);
        d.b[row] = bj;
        d.r[row] = rj;
        double uj = d.Minv[row] * rj;
        if (wood) d.W.v[row] = uj;
        else if (!slack) uj = store_u(d, d.uu, row, uj);
        red3[0] += rj * uj;
        red3[1] = fmax(red3[1], fabs(rj));
        red3[2] = fmax(red3[2], fabs(bj));
      }
    }
    reduce_and_barrier<3>(g, sm, red3, 0x6u);
    if (wood) {
      double redg[1] = {wood_apply(g, sm, d, d.Minv, d.r, d.uu, n0, n1)};
      reduce_and_barrier<1>(g, sm, redg, 0u);
      red3[0] = redg[0];
    }
    pc.tick(refresh ? 11 : 10);
    refreshes += refresh;
    refresh = 0;
    {
      // stop when the residual has dropped by pcg_eta relative to where this ADMM step started
      // (r0 measures how far the system moved since the last solve), floored at roundoff level
      const double thresh = fmax(c.pcg_eta * red3[1], c.pcg_floor * red3[2]);
      // the energy-norm rule rides on gamma = r'M^-1 r, which measures the error only when M is close to K: on by default
      // (eta_e = 1e-3) with the slack-elimination preconditioner, where it restores the oracle's iteration counts on the
      // Lasso config at no extra PCG iterations; off with plain Jacobi, where it costs 26 % more PCG iterations on config 2
      // for no change in parity, and with the Woodbury correction (M = K on the portfolio: one iteration is exact)
      const double ee = c.pcg_eta_e >= 0.0 ? c.pcg_eta_e : (slack ? 1e-3 : 0.0);
      const double tfl = c.pcg_floor * red3[2], ee2 = ee * ee;
      int ncg = slack     ? pcg_run_stream_slack<MODE(false, true)>(g, sm, SG, pc, d, pv, d.rho_vec, d.Minv, c.sigma, d.xt, d.zt, red3[1],
                                                 thresh, c.pcg_max_iter, m0, m1, n0, n1, tfl, ee2)
                : MODE(d.blocked, 1) ? pcg_run_stream<MODE(false, true)>(g, sm, SG, pc, d, pv, d.rho_vec, d.Minv, c.sigma, d.xt, d.zt, red3[0],
                                           red3[1], thresh, c.pcg_max_iter, m0, m1, n0, n1, tfl, ee2)
                          : pcg_run(g, sm, pc, d, pv, d.rho_vec, d.Minv, c.sigma, d.xt, d.zt, red3[0], red3[1],
                                    thresh, c.pcg_max_iter, m0, m1, n0, n1, tfl, ee2);
      if (ncg < 0) {  // K = P + sigma I + A' rho A has a direction of non-positive curvature (uniform over the grid)
        ncg = -ncg - 1;
        pcg_broke = true;
      }
      cg_total += ncg;
      cg_solves++;
    }
    if (pcg_broke) {
      status = ST_NON_CVX;
      info_iter = it;
      can_check = true;  // nothing left to evaluate: the iterate is meaningless
      can_print = false;
      break;
    }
    // ---- Z: x, z, y updates (rows a6-a8); dy is stored already projected on the polar of the
    //         recession cone of [l,u] (is_primal_infeasible does that projection in place)
    for (int i = m0 + tid; i < m1; i += nth) {
      const double zti = d.zt[i], zi = d.z[i], yi = d.y[i], ri = d.rho_vec[i], rinv = d.rho_inv[i];
      const double li = d.l[i], ui = d.u[i];
      const double zh = c.alpha * zti + (1.0 - c.alpha) * zi;
      const double zn = fmin(fmax(zh + rinv * yi, li), ui);
      double dyi = ri * (zh - zn);
      const double yn = yi + dyi;
      d.y[i] = yn;
      d.z[i] = zn;
      d.wv[i] = ri * zn - yn;  // rhs vector of the next step (P1)
      if (ui > kInfty * kMinScaling) {
        if (li < -kInfty * kMinScaling) dyi = 0.0;
        else dyi = fmin(dyi, 0.0);
      } else if (li < -kInfty * kMinScaling) {
        dyi = fmax(dyi, 0.0);
      }
      d.dy[i] = dyi;
    }
    for (int j = n0 + tid; j < n1; j += nth) {
      const double xj = d.x[j];
      const double xn = c.alpha * d.xt[j] + (1.0 - c.alpha) * xj;
      d.dx[j] = xn - xj;
      d.x[j] = xn;
    }
    if (MODE(d.blocked, 1) && d.m > 0) stream_prefetch_head(d.ST);
    {
      double redt[1] = {0.0};
      if (b == 0 && tid == 0) redt[0] = (double)(globaltimer_ns() - t_start) * 1e-9;
      reduce_and_barrier<1>(g, sm, redt, 0x1u);
      elapsed = redt[0];
    }
    pc.tick(8);
    if (c.time_limit_s > -1e29 && elapsed >= c.time_limit_s) {
      status = ST_TIME_LIMIT;
      can_check = false;
      can_print = false;
      break;
    }
    can_check = c.check_termination && (it % c.check_termination == 0);
    can_print = c.verbose && ((it % kPrintInterval == 0) || it == 1);
    if (can_check || can_print) {
      if (MODE(d.blocked, 1) && MODE(d.info_streams, 1)) compute_info_stream(g, sm, SG, d, c, cost_c, cost_cinv, d.x, d.z, d.y, m0, m1, n0, n1, S);
      else compute_info(g, sm, d, c, cost_c, cost_cinv, d.x, d.z, d.y, m0, m1, n0, n1, S);
      pc.tick(12);
      info_iter = it;
      checks++;
      if (can_print && b == 0 && tid == 0 && log_rows < kLogRows) {
        double *row = d.info->log[log_rows];
        row[0] = (double)it; row[1] = S.obj_val; row[2] = S.pri_res; row[3] = S.dua_res; row[4] = rho; row[5] = elapsed;
      }
      if (can_print && log_rows < kLogRows) log_rows++;
      if (can_check) {
        status = check_termination(S, c, d.m, cost_c, cost_cinv, false);
        if (status != ST_UNSOLVED) break;
      }
    }
    if (c.adaptive_rho && interval == 0 && elapsed > c.adaptive_time_s) {
      const long long N = c.check_termination ? c.check_termination : 25;
      const double xx = (double)it + 0.5 * (double)N;
      interval = (long long)(xx - fmod(xx, (double)N));
      if (interval < c.check_termination) interval = c.check_termination;
    }
    if (c.adaptive_rho && interval && (it % interval == 0)) {
      if (!can_check && !can_print) {
        if (MODE(d.blocked, 1) && MODE(d.info_streams, 1)) compute_info_stream(g, sm, SG, d, c, cost_c, cost_cinv, d.x, d.z, d.y, m0, m1, n0, n1, S);
      else compute_info(g, sm, d, c, cost_c, cost_cinv, d.x, d.z, d.y, m0, m1, n0, n1, S);
        info_iter = it;
        checks++;
      }
      const double rho_new = rho_estimate(S, rho);
      rho_est = rho_new;
      if (rho_new > rho * c.adaptive_rho_tolerance || rho_new < rho / c.adaptive_rho_tolerance) {
        rho = fmin(fmax(rho_new, kRhoMin), kRhoMax);
        rho_updates++;
        for (int i = m0 + tid; i < m1; i += nth) {
          const int ct = d.ctype[i];
          if (ct == 0) { d.rho_vec[i] = rho; d.rho_inv[i] = 1.0 / rho; }
          else if (ct == 1) { d.rho_vec[i] = kRhoEqOverIneq * rho; d.rho_inv[i] = 1.0 / (kRhoEqOverIneq * rho); }
        }
        if (wood) {
          wood_refresh(g, sm, d, d.rho_vec, c.sigma, d.Minv, n0, n1);
        } else if (slack) {
          slack_refresh(g, d, d.rho_vec, c.sigma, d.Minv, m0, m1, n0, n1);
        } else {
          grid_barrier(g);
          precond_rows(d, d.rho_vec, c.sigma, d.Minv, n0, n1);
        }
        pc.tick(13);
        refresh = 1;  // K changed: the residual recurrence is void
      }
    }
    if (c.refresh_every > 0 && (it % c.refresh_every == 0)) refresh = 1;
  }
  if (!can_check) {
    if (!can_print) {
      if (MODE(d.blocked, 1) && MODE(d.info_streams, 1)) compute_info_stream(g, sm, SG, d, c, cost_c, cost_cinv, d.x, d.z, d.y, m0, m1, n0, n1, S);
      else compute_info(g, sm, d, c, cost_c, cost_cinv, d.x, d.z, d.y, m0, m1, n0, n1, S);
      info_iter = it - 1;
      checks++;
    }
    const long long s2 = check_termination(S, c, d.m, cost_c, cost_cinv, false);
    if (s2 != ST_UNSOLVED) status = s2;
  }
  rho_est = pcg_broke ? rho : rho_estimate(S, rho);
  if (status == ST_UNSOLVED) {
    const long long s2 = check_termination(S, c, d.m, cost_c, cost_cinv, true);
    status = (s2 != ST_UNSOLVED) ? s2 : ST_MAX_ITER;
  } else if (status == ST_TIME_LIMIT) {
    const long long s2 = check_termination(S, c, d.m, cost_c, cost_cinv, true);
    if (s2 != ST_UNSOLVED) status = s2;
  }
  double obj_val = S.obj_val;
  if (status == ST_NON_CVX) obj_val = nan("");
  if (status == ST_PINF || status == ST_PINF_INACC) obj_val = kInfty;
  if (status == ST_DINF || status == ST_DINF_INACC) obj_val = -kInfty;

  // ---- store_solution (row a16)
  const bool unscale = c.scaling && !c.scaled_termination;
  const bool pinf = (status == ST_PINF || status == ST_PINF_INACC), dinf = (status == ST_DINF || status == ST_DINF_INACC);
  const bool has_sol = !(pinf || dinf || status == ST_NON_CVX);
  if (has_sol) {
    for (int j = n0 + tid; j < n1; j += nth) d.sol_x[j] = c.scaling ? d.D[j] * d.x[j] : d.x[j];
    for (int i = m0 + tid; i < m1; i += nth) d.sol_y[i] = c.scaling ? cost_cinv * d.E[i] * d.y[i] : d.y[i];
  } else {
    const double qnan = nan("");
    for (int j = n0 + tid; j < n1; j += nth) {
      d.sol_x[j] = qnan;
      if (dinf) d.dx[j] = (unscale ? d.D[j] : 1.0) * d.dx[j] / S.ndx_t;
      d.x[j] = 0.0;
      d.xt[j] = 0.0;
    }
    for (int i = m0 + tid; i < m1; i += nth) {
      d.sol_y[i] = qnan;
      if (pinf) d.dy[i] = (unscale ? d.E[i] : 1.0) * d.dy[i] / S.ndy_t;
      d.z[i] = 0.0;
      d.y[i] = 0.0;
      d.zt[i] = 0.0;
    }
  }
  if (b == 0 && tid == 0) {
    DevInfo *I = d.info;
    I->iter = info_iter;
    I->status_val = status;
    I->obj_val = obj_val;
    I->pri_res = S.pri_res;
    I->dua_res = S.dua_res;
    I->rho_estimate = rho_est;
    I->rho_updates = rho_updates;
    I->rho = rho;
    I->adaptive_interval = interval;
    I->cg_iters = cg_total;
    I->cg_solves = cg_solves;
    I->checks = checks;
    I->refreshes = refreshes;
    I->elapsed_s = (double)(globaltimer_ns() - t_start) * 1e-9;
    I->log_rows = log_rows;
    pc.tick(14);
    for (int k = 0; k < kPhases; k++) I->phase_us[k] = phase_acc[k];
    d.state->rho = rho;
    d.state->rho_updates = rho_updates;
    d.state->adaptive_interval = interval;
    d.state->needs_refresh = 0;
  }
}

#ifndef OSQP_B200_FAST  // (only the ADMM and polish kernels are compiled a second time)
// ------------------------------------------------------------------ setup-time curvature probe
// libosqp fails osqp_setup when the LDL' of the KKT matrix has a wrong-sign pivot, i.e. when
// P + sigma I is not positive definite (test/non_convex.jl:13-21).  Without a factorisation we
// run CG on (P + sigma I) v = b for a pseudo-random b: CG's pivots p'(P+sigma I)p are the pivots
// of the Lanczos tridiagonal, so a non-positive one appears as soon as the smallest Ritz value
// crosses zero (exact after n steps; extreme eigenvalues converge first).
__global__ void __launch_bounds__(kThreads, 1) pd_probe_kernel(const __grid_constant__ DevPtrs d, double sigma, int max_it,
                                                               int trials) {
  __shared__ RedSmem sm;
  Grid g;
  grid_init(g, d);
  const int tid = threadIdx.x, nth = blockDim.x, b = blockIdx.x;
  const int n0 = d.n_start[b], n1 = d.n_start[b + 1];
  const int lanesN = d.At.lanes;
  const int subN = tid & (lanesN - 1), grpN = tid / lanesN, ngrpN = nth / lanesN;
  int failed = 0;
  // Every trial runs CG from its own pseudo-random right-hand side until the residual has dropped by 1e-10 (the
  // Krylov space then holds every eigen-direction the start vector excites above that level, and a negative one
  // would have shown up as a non-positive pivot on the way) or the iteration budget is spent.
  for (int trial = 0; trial < trials && !failed; trial++) {
    double red[1] = {0.0};
    for (int j = n0 + tid; j < n1; j += nth) {
      unsigned h = (unsigned)j * 2654435761u + 12345u + 977u * (unsigned)trial;
      h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
      const double v = ((double)(h & 0xFFFFFF) / 16777216.0) - 0.5 + 1e-3;
      d.r[j] = v;
      d.p[j] = v;
      red[0] += v * v;
    }
    reduce_and_barrier<1>(g, sm, red, 0u);
    double rr = red[0];
    const double rr0 = rr;
    for (int it = 0; it < max_it && rr > 1e-20 * rr0; it++) {
      red[0] = 0.0;
      for (int base = n0; base < n1; base += ngrpN) {
        const int row = base + grpN;
        const bool valid = row < n1;
        Acc<1> acc;
        acc.a[0] = 0.0;
        row_accumulate<1>(d.P.rowptr, d.P.col, d.P.val, row, valid, subN, lanesN, acc,
                          [&](Acc<1> &ac, int c, double a) { ac.a[0] += a * d.p[c]; });
        group_reduce<1>(acc, lanesN);
        if (valid && subN == 0) {
          const double pj = d.p[row];
          const double wj = acc.a[0] + sigma * pj;
          d.w[row] = wj;
          red[0] += wj * pj;
        }
      }
      reduce_and_barrier<1>(g, sm, red, 0u);
      const double pw = red[0];
      if (!(pw > 0.0)) {
        failed = 1;
        break;
      }
      const double a = rr / pw;
      red[0] = 0.0;
      for (int j = n0 + tid; j < n1; j += nth) {
        const double rj = d.r[j] - a * d.w[j];
        d.r[j] = rj;
        red[0] += rj * rj;
      }
      reduce_and_barrier<1>(g, sm, red, 0u);
      const double beta = red[0] / rr;
      rr = red[0];
      for (int j = n0 + tid; j < n1; j += nth) d.p[j] = d.r[j] + beta * d.p[j];
      grid_barrier(g);
    }
  }
  if (b == 0 && tid == 0) d.state->pd_check_failed = failed;
}

// Gershgorin certificate on the scaled P + sigma I: strictly diagonally dominant rows with a positive diagonal prove
// positive definiteness in one pass over P (diagonal, banded-dominant and regularised-Laplacian Hessians all pass).
// One warp per row; state->pd_certified must be 1 on entry and is cleared by any row that fails.
__global__ void k_gershgorin(const DevPtrs d, double sigma) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < d.n; r += nwarps) {
    double off = 0.0;
    for (int k = d.P.rowptr[r] + lane; k < d.P.rowptr[r + 1]; k += 32)
      if (d.P.col[k] != r) off += fabs(d.P.val[k]);
    for (int o = 16; o; o >>= 1) off += __shfl_xor_sync(0xffffffffu, off, o);
    if (lane == 0 && !(d.Pdiag[r] + sigma > off)) d.state->pd_certified = 0;
  }
}

#endif  // !OSQP_B200_FAST

// ------------------------------------------------------------------ polish (row a12)
// Active-set guess as in libosqp (z - l < -y lower-active, u - z < y upper-active), then the
// equality-constrained QP  min 1/2 x'Px + q'x  s.t. A_act x = b_act  is solved by a proximal
// method of multipliers whose inner systems  (P + delta I + penalty A_act'A_act) x = rhs  reuse
// the PCG above (rho := penalty on active rows, 0 elsewhere).  libosqp factorises the
// delta-regularised KKT and applies `polish_refine_iter` refinement steps; both converge to
// the same KKT point of the active-set QP.
__global__ void __launch_bounds__(kThreads, 1) polish_kernel(const __grid_constant__ DevPtrs d, const __grid_constant__ PolishCfg c,
                                                         const __grid_constant__ SolveCfg sc,
                                                         PolishOut *out) {
  __shared__ RedSmem sm;
  __shared__ double phase_acc[kPhases];
  PhaseClock pc;
  pc.start(phase_acc);
  Grid g;
  grid_init(g, d);
  Slice SG;
  slice_init(SG, d);
  const int tid = threadIdx.x, nth = blockDim.x, b = blockIdx.x;
  const int m0 = d.m_start[b], m1 = d.m_start[b + 1], n0 = d.n_start[b], n1 = d.n_start[b + 1];
  const int lanesA = d.A.lanes, lanesN = d.At.lanes;
  const int subA = tid & (lanesA - 1), grpA = tid / lanesA, ngrpA = nth / lanesA;
  const int subN = tid & (lanesN - 1), grpN = tid / lanesN, ngrpN = nth / lanesN;
  const double cost_c = d.state->c, cost_cinv = d.state->cinv;
  const PcgVecs pv{d.r, d.uu, d.p, d.s, d.w, d.t, d.tr, d.Ap};

  // ---- active sets, start point (ADMM x, multipliers y on active rows)
  double cnt[1] = {0.0};
  for (int i = m0 + tid; i < m1; i += nth) {
    const double zi = d.z[i], yi = d.y[i], li = d.l[i], ui = d.u[i];
    const bool low = (zi - li < -yi), upp = (ui - zi < yi);
    // libosqp gives lower-active precedence when both tests fire
    const bool act = low || upp;
    d.pol_rho[i] = act ? c.penalty : 0.0;
    d.pol_b[i] = low ? li : (upp ? ui : 0.0);
    d.pol_y[i] = act ? yi : 0.0;
    cnt[0] += act ? 1.0 : 0.0;
  }
  for (int j = n0 + tid; j < n1; j += nth) d.pol_x[j] = d.x[j];
  reduce_and_barrier<1>(g, sm, cnt, 0u);
  const long long n_active = (long long)(cnt[0] + 0.5);
  // preconditioner for K_pol
  double *Minv_pol = d.pol_rhs;  // pol_rhs is free here
  const bool wood = MODE(d.blocked && d.W.w > 0, false);
  if (wood) wood_refresh(g, sm, d, d.pol_rho, c.delta, Minv_pol, n0, n1);
  else precond_rows(d, d.pol_rho, c.delta, Minv_pol, n0, n1);
  long long cg_total = 0;
  for (int outer = 0; outer <= c.refine_iter; outer++) {
    // z = A x (fresh), wv = penalty*(b - z) - y on active rows : K dx = -(P x + q + delta*0) + A'(wv) ...
    // Solve for the full new x with warm start x:  r = rhs - K x where
    //   rhs = -q + delta x_k + A'(penalty b - y),  K = P + delta I + A' penalty A
    //   => r = -q - P x + A'(penalty (b - A x) - y)
    if (d.m > 0) {
      for (int base = m0; base < m1; base += ngrpA) {
        const int row = base + grpA;
        const bool valid = row < m1;
        Acc<1> acc;
        acc.a[0] = 0.0;
        row_accumulate<1>(d.A.rowptr, d.A.col, d.A.val, row, valid, subA, lanesA, acc,
                          [&](Acc<1> &ac, int cc, double a) { ac.a[0] += a * d.pol_x[cc]; });
        group_reduce<1>(acc, lanesA);
        if (valid && subA == 0) {
          d.pol_z[row] = acc.a[0];
          d.wv[row] = d.pol_rho[row] * (d.pol_b[row] - acc.a[0]) - d.pol_y[row];
        }
      }
    }
    grid_barrier(g);
    double red3[3] = {0.0, 0.0, 0.0};
    for (int base = n0; base < n1; base += ngrpN) {
      const int row = base + grpN;
      const bool valid = row < n1;
      Acc<2> acc;
      acc.a[0] = acc.a[1] = 0.0;
      row_accumulate<2>(d.P.rowptr, d.P.col, d.P.val, row, valid, subN, lanesN, acc,
                        [&](Acc<2> &ac, int cc, double a) { ac.a[1] += a * d.pol_x[cc]; });
      if (d.m > 0)
        row_accumulate<2>(d.At.rowptr, d.At.col, d.At.val, row, valid, subN, lanesN, acc,
                          [&](Acc<2> &ac, int cc, double a) { ac.a[0] += a * d.wv[cc]; });
      group_reduce<2>(acc, lanesN);
      if (valid && subN == 0) {
        const double rj = -d.q[row] - acc.a[1] + acc.a[0];
        d.r[row] = rj;
        double uj = Minv_pol[row] * rj;
        if (wood) d.W.v[row] = uj;
        else uj = store_u(d, d.uu, row, uj);
        red3[0] += rj * uj;
        red3[1] = fmax(red3[1], fabs(rj));
        red3[2] = fmax(red3[2], fabs(d.q[row]) + fabs(acc.a[1]));  // scale of the stationarity terms, penalty-free
      }
    }
    reduce_and_barrier<3>(g, sm, red3, 0x6u);
    if (wood) {
      double redg[1] = {wood_apply(g, sm, d, Minv_pol, d.r, d.uu, n0, n1)};
      reduce_and_barrier<1>(g, sm, redg, 0u);
      red3[0] = redg[0];
    }
    const double thresh = c.pcg_rel_tol * fmax(red3[2], 1e-3);
    {
      const int ncg = MODE(d.blocked, 1) ? pcg_run_stream<false>(g, sm, SG, pc, d, pv, d.pol_rho, Minv_pol, c.delta, d.pol_x, d.pol_z,
                                                 red3[0], red3[1], thresh, c.pcg_max_iter, m0, m1, n0, n1)
                                : pcg_run(g, sm, pc, d, pv, d.pol_rho, Minv_pol, c.delta, d.pol_x, d.pol_z, red3[0],
                                          red3[1], thresh, c.pcg_max_iter, m0, m1, n0, n1);
      cg_total += ncg < 0 ? -ncg - 1 : ncg;
    }
    // multiplier step on active rows: y += penalty (A x - b)
    for (int i = m0 + tid; i < m1; i += nth)
      if (d.pol_rho[i] > 0.0) d.pol_y[i] += d.pol_rho[i] * (d.pol_z[i] - d.pol_b[i]);
    grid_barrier(g);
  }
  // ---- z = A x exactly, then project (z, y) on the normal cone of [l, u]
  if (d.m > 0) {
    for (int base = m0; base < m1; base += ngrpA) {
      const int row = base + grpA;
      const bool valid = row < m1;
      Acc<1> acc;
      acc.a[0] = 0.0;
      row_accumulate<1>(d.A.rowptr, d.A.col, d.A.val, row, valid, subA, lanesA, acc,
                        [&](Acc<1> &ac, int cc, double a) { ac.a[0] += a * d.pol_x[cc]; });
      group_reduce<1>(acc, lanesA);
      if (valid && subA == 0) {
        const double tt = acc.a[0] + d.pol_y[row];
        const double zn = fmin(fmax(tt, d.l[row]), d.u[row]);
        d.pol_z[row] = zn;
        d.pol_y[row] = tt - zn;
      }
    }
  }
  // compute_info reads dx / dy for the infeasibility products; they are irrelevant here
  grid_barrier(g);
  InfoScalars S;
  compute_info(g, sm, d, sc, cost_c, cost_cinv, d.pol_x, d.pol_z, d.pol_y, m0, m1, n0, n1, S);
  // residuals of the ADMM iterate were published by admm_kernel
  const double pri0 = d.info->pri_res, dua0 = d.info->dua_res;
  const bool ok = (S.pri_res < pri0 && S.dua_res < dua0) || (S.pri_res < pri0 && dua0 < 1e-10) ||
                  (S.dua_res < dua0 && pri0 < 1e-10);
  if (ok) {
    for (int j = n0 + tid; j < n1; j += nth) {
      const double xj = d.pol_x[j];
      d.x[j] = xj;
      d.xt[j] = xj;
      d.sol_x[j] = sc.scaling ? d.D[j] * xj : xj;
    }
    for (int i = m0 + tid; i < m1; i += nth) {
      d.z[i] = d.pol_z[i];
      d.zt[i] = d.pol_z[i];
      d.y[i] = d.pol_y[i];
      d.sol_y[i] = sc.scaling ? cost_cinv * d.E[i] * d.pol_y[i] : d.pol_y[i];
    }
  }
  grid_barrier(g);
  if (b == 0 && tid == 0) {
    out->n_active = n_active;
    out->obj_val = S.obj_val;
    out->pri_res = S.pri_res;
    out->dua_res = S.dua_res;
    out->cg_iters = cg_total;
    out->success = ok ? 1 : 0;
    d.state->needs_refresh = 1;
  }
}

#ifndef OSQP_B200_FAST
// ------------------------------------------------------------------ standalone SpMV (profiling / parity)
// which: 0  out = A in (m) | 1  out = A' in (n) | 2  out = (P + sigma I) in (n)
__global__ void __launch_bounds__(kThreads, 1) spmv_kernel(const DevPtrs d, int which, const double *in, double *out,
                                                       double sigma) {
  const int tid = threadIdx.x, nth = blockDim.x, b = blockIdx.x;
  const int m0 = d.m_start[b], m1 = d.m_start[b + 1], n0 = d.n_start[b], n1 = d.n_start[b + 1];
  const int lanesA = d.A.lanes, lanesN = d.At.lanes;
  const int subA = tid & (lanesA - 1), grpA = tid / lanesA, ngrpA = nth / lanesA;
  const int subN = tid & (lanesN - 1), grpN = tid / lanesN, ngrpN = nth / lanesN;
  if (which == 0) {
    for (int base = m0; base < m1; base += ngrpA) {
      const int row = base + grpA;
      const bool valid = row < m1;
      Acc<1> acc;
      acc.a[0] = 0.0;
      row_accumulate<1>(d.A.rowptr, d.A.col, d.A.val, row, valid, subA, lanesA, acc,
                        [&](Acc<1> &ac, int c, double a) { ac.a[0] += a * in[c]; });
      group_reduce<1>(acc, lanesA);
      if (valid && subA == 0) out[row] = acc.a[0];
    }
  } else {
    for (int base = n0; base < n1; base += ngrpN) {
      const int row = base + grpN;
      const bool valid = row < n1;
      Acc<1> acc;
      acc.a[0] = 0.0;
      if (which == 1)
        row_accumulate<1>(d.At.rowptr, d.At.col, d.At.val, row, valid, subN, lanesN, acc,
                          [&](Acc<1> &ac, int c, double a) { ac.a[0] += a * in[c]; });
      else
        row_accumulate<1>(d.P.rowptr, d.P.col, d.P.val, row, valid, subN, lanesN, acc,
                          [&](Acc<1> &ac, int c, double a) { ac.a[0] += a * in[c]; });
      group_reduce<1>(acc, lanesN);
      if (valid && subN == 0) out[row] = acc.a[0] + (which == 2 ? sigma * in[row] : 0.0);
    }
  }
}

// Standalone SpMV on the tile streams: exactly the phase code of pcg_run_stream (stream phase -> grid barrier ->
// owner sum of the partials).  which 0 and 2 both run the [A; P] stream; the requested half is written out.
__global__ void __launch_bounds__(kThreads, 1) spmv_stream_kernel(const __grid_constant__ DevPtrs d, int which, const double *in,
                                                                  double *out, double sigma) {
  Grid g;
  grid_init(g, d);
  Slice SG;
  slice_init(SG, d);
  const int tid = threadIdx.x, nth = blockDim.x, b = blockIdx.x;
  const int m0 = d.m_start[b], m1 = d.m_start[b + 1], n0 = d.n_start[b], n1 = d.n_start[b + 1];
  const TileStreamDev &T = (which == 1) ? d.ST : d.SA;
  // globaltimer probes per block: 0 start | 1 after the first grid barrier | 2 first / 3 last warp done with its
  // stream | 4 block done | 5 after the second grid barrier
  unsigned long long *probe = d.dbg + (size_t)b * 16;
  if (tid == 0) { probe[0] = globaltimer_ns(); probe[2] = ~0ull; probe[3] = 0ull; }
  stream_prefetch_head(T);
  grid_barrier(g);
  if (tid == 0) probe[1] = globaltimer_ns();
  SG.probe = probe;
  if (which != 1 && MODE(d.SA.paired, FAST_PAIRED)) {
    stream_phase_paired(SG, d.SA, in, [&](int r, double sum) {
      if (which == 0 && r < d.m) out[r] = sum;
      if (which == 2 && r >= d.m) out[r - d.m] = sum + sigma * in[r - d.m];
    });
  } else {
    stream_phase(SG, T, in);
  }
  if ((tid & 31) == 0) {
    const unsigned long long t = globaltimer_ns();
    atomicMin(probe + 2, t);
    atomicMax(probe + 3, t);
  }
  __syncthreads();
  if (tid == 0) probe[4] = globaltimer_ns();
  grid_barrier(g);
  if (tid == 0) probe[5] = globaltimer_ns();
  if (which == 1) {
    for (int j = n0 + tid; j < n1; j += nth) out[j] = part_sum(d.ST, j);
  } else if (!MODE(d.SA.paired, FAST_PAIRED)) {
    if (which == 0)
      for (int i = m0 + tid; i < m1; i += nth) out[i] = part_sum(d.SA, i);
    else
      for (int j = n0 + tid; j < n1; j += nth) out[j] = part_sum(d.SA, d.m + j) + sigma * in[j];
  }
}

#ifdef OSQP_B200_DEVTOOLS  // measurement / self-test kernels: only in lib/libosqp_dev.so, never in the product
// ------------------------------------------------------------------ stream micro-benchmark (profiles/membench.py)
// Reads `bytes` of `buf` with the access shape of stream_phase (per lane and chunk: 2 x 16 B + 8 B loads, kD chunks in
// flight) and nothing else, so the memory system's ceiling for that shape can be separated from the reduction code.
//   pattern 0: every warp walks its own contiguous 1/(grid*16) share            (the layout of engine v3)
//   pattern 1: the 16 warps of a block interleave chunk by chunk in one share   (block-contiguous windows)
//   pattern 2: as 1, and a chunk is one contiguous 1280 B record (values then columns)
template <int kD>
__global__ void __launch_bounds__(kThreads, 1) membench_kernel(const char *buf, long long bytes, int pattern,
                                                               double *sink) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long nwarps = (long long)gridDim.x * kWarps;
  const long long chunks_total = bytes / 1280;
  const long long per_warp = chunks_total / nwarps;
  const char *vbase = buf, *cbase = buf + chunks_total * 1024;
  double acc = 0.0;
  unsigned cacc = 0u;
  QuadSlot slot[kD];
  long long issued = 0;
  auto issue = [&](QuadSlot &q) {
    if (issued < per_warp) {
      long long c;  // global chunk index
      if (pattern == 0) c = ((long long)blockIdx.x * kWarps + warp) * per_warp + issued;
      else c = (long long)blockIdx.x * kWarps * per_warp + issued * kWarps + warp;
      const char *pv = (pattern == 2) ? buf + c * 1280 + lane * 16 : vbase + c * 1024 + lane * 16;
      const char *pc = (pattern == 2) ? buf + c * 1280 + 1024 + lane * 8 : cbase + c * 256 + lane * 8;
      quad_load(q, pv, pc, true);
    }
    issued++;
  };
#pragma unroll
  for (int k = 0; k < kD; k++) issue(slot[k]);
  for (long long base = 0; base < per_warp; base += kD) {
#pragma unroll
    for (int k = 0; k < kD; k++) {
      acc += slot[k].v0 + slot[k].v1 + slot[k].v2 + slot[k].v3;
      cacc ^= slot[k].c01 ^ slot[k].c23;
      issue(slot[k]);
    }
  }
  if (acc == 1.2345 && cacc == 77u) sink[0] = acc;  // keep the loads alive
}

// ------------------------------------------------------------------ barrier micro-benchmark (profiles/membench.py)
// mode 0: `iters` bare grid barriers | 1: reduce_and_barrier<2> | 2: barrier after scattered 8 B stores | 3: reduce_and_barrier_fx<2>
__global__ void __launch_bounds__(kThreads, 1) barrier_bench_kernel(const DevPtrs d, int iters, int mode, double *sink,
                                                                    unsigned long long *ns_out) {
  __shared__ RedSmem sm;
  Grid g;
  grid_init(g, d);
  grid_barrier(g);
  const unsigned long long t0 = globaltimer_ns();
  double acc = 0.0;
  for (int it = 0; it < iters; it++) {
    if (mode == 1) {
      double v[2] = {1.0 + acc, (double)threadIdx.x};
      reduce_and_barrier<2>(g, sm, v, 0x2u);
      acc = v[0] * 1e-9;
    } else if (mode == 3) {
      double v[2] = {1.0 + acc, (double)threadIdx.x};
      reduce_and_barrier_fx<2>(g, sm, v, 0x2u, 1.0e5);
      acc = v[0] * 1e-9;
    } else {
      if (mode == 2 && (threadIdx.x & 31) < 2) sink[64 + ((size_t)blockIdx.x * 1024 + threadIdx.x * 17 + it) % 100000] = acc;
      grid_barrier(g);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    ns_out[0] = globaltimer_ns() - t0;
    sink[0] = acc;
  }
}

// Self-test of the two cross-block reductions (tests/test_engine_parity.py): every thread contributes its global
// index + 1 to a sum and a max; out = {tree sum, tree max, fixed-point sum, fixed-point max}.  A `ref` that is far
// too small makes the scaled partials overflow, which must route reduce_and_barrier_fx through its fp64 fallback.
__global__ void __launch_bounds__(kThreads, 1) reduce_selftest_kernel(const DevPtrs d, double ref, double *out) {
  __shared__ RedSmem sm;
  Grid g;
  grid_init(g, d);
  const double mine = (double)(blockIdx.x * blockDim.x + threadIdx.x + 1);
  double a[2] = {mine, mine}, b[2] = {mine, mine};
  reduce_and_barrier<2>(g, sm, a, 0x2u);
  reduce_and_barrier_fx<2>(g, sm, b, 0x2u, ref);
  double c[2] = {0.5 * mine, mine};
  reduce_and_barrier_fx<2>(g, sm, c, 0x2u, ref);  // a second call: bank rotation
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    out[0] = a[0]; out[1] = a[1]; out[2] = b[0]; out[3] = b[1]; out[4] = c[0]; out[5] = c[1];
  }
}

#endif  // OSQP_B200_DEVTOOLS

// ------------------------------------------------------------------ setup kernels (row a2, a3): simple grid-stride
__device__ __forceinline__ double limit_scaling(double a) {
  a = a < kMinScaling ? 1.0 : a;
  return a > kMaxScaling ? kMaxScaling : a;
}

__global__ void k_copy_vals(const DevPtrs d) {
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  for (long long k = tid; k < d.A.nnz; k += nth) { d.A.val[k] = d.A.val0[k]; d.At.val[k] = d.At.val0[k]; }
  for (long long k = tid; k < d.P.nnz; k += nth) d.P.val[k] = d.P.val0[k];
  for (long long j = tid; j < d.n; j += nth) { d.q[j] = d.q0[j]; d.D[j] = 1.0; d.Dinv[j] = 1.0; }
  for (long long i = tid; i < d.m; i += nth) { d.E[i] = 1.0; d.Einv[i] = 1.0; }
  if (tid == 0) { d.state->c = 1.0; d.state->cinv = 1.0; }
}

// one warp per row: infinity norms of the KKT columns -> dtmp, etmp = 1/sqrt(limit(norm))
__global__ void k_ruiz_norms(const DevPtrs d) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < (long long)d.n + d.m; r += nwarps) {
    double mx = 0.0;
    if (r < d.n) {
      for (int k = d.P.rowptr[r] + lane; k < d.P.rowptr[r + 1]; k += 32) mx = fmax(mx, fabs(d.P.val[k]));
      if (d.m > 0)
        for (int k = d.At.rowptr[r] + lane; k < d.At.rowptr[r + 1]; k += 32) mx = fmax(mx, fabs(d.At.val[k]));
    } else {
      const long long i = r - d.n;
      for (int k = d.A.rowptr[i] + lane; k < d.A.rowptr[i + 1]; k += 32) mx = fmax(mx, fabs(d.A.val[k]));
    }
    for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) {
      const double s = 1.0 / sqrt(limit_scaling(mx));
      if (r < d.n) d.dtmp[r] = s;
      else d.etmp[r - d.n] = s;
    }
  }
}

// P <- Dt P Dt, A <- Et A Dt (same rounding order as premult then postmult), q, D, E
__global__ void k_ruiz_apply(const DevPtrs d) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < (long long)d.n + d.m; r += nwarps) {
    if (r < d.n) {
      const double dr = d.dtmp[r];
      for (int k = d.P.rowptr[r] + lane; k < d.P.rowptr[r + 1]; k += 32) {
        // symmetric full storage: entry (r, c) must get the same bits as (c, r): scale by the
        // smaller index first, then the larger (the upper-triangle order row <= col of the oracle)
        const int cidx = d.P.col[k];
        const double dc = d.dtmp[cidx];
        const double first = (r <= cidx) ? dr : dc, second = (r <= cidx) ? dc : dr;
        d.P.val[k] = (d.P.val[k] * first) * second;
      }
      if (d.m > 0)
        for (int k = d.At.rowptr[r] + lane; k < d.At.rowptr[r + 1]; k += 32)
          d.At.val[k] = (d.At.val[k] * d.etmp[d.At.col[k]]) * dr;
      if (lane == 0) { d.q[r] *= dr; d.D[r] *= dr; }
    } else {
      const long long i = r - d.n;
      const double er = d.etmp[i];
      for (int k = d.A.rowptr[i] + lane; k < d.A.rowptr[i + 1]; k += 32)
        d.A.val[k] = (d.A.val[k] * er) * d.dtmp[d.A.col[k]];
      if (lane == 0) d.E[i] *= er;
    }
  }
}

// cost normalisation: c_temp = 1 / limit(max(mean_j |P_:j|inf, limit(|q|inf))).  Two fixed-order stages (the value
// must be bit-identical from run to run): every block reduces a contiguous slice of the columns into red[], one
// block combines the per-block values.
constexpr int kCostBlocks = 148;
__global__ void __launch_bounds__(256) k_ruiz_cost_partial(const DevPtrs d) {
  __shared__ double ssum[8], smax[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int per = (d.n + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * per, r1 = min(d.n, r0 + per);
  double sum = 0.0, qmax = 0.0;
  for (int r = r0 + warp; r < r1; r += nwarps) {
    double mx = 0.0;
    for (int k = d.P.rowptr[r] + lane; k < d.P.rowptr[r + 1]; k += 32) mx = fmax(mx, fabs(d.P.val[k]));
    for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) { sum += mx; qmax = fmax(qmax, fabs(d.q[r])); }
  }
  if (lane == 0) { ssum[warp] = sum; smax[warp] = qmax; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0, q = 0.0;
    for (int w = 0; w < nwarps; w++) { s += ssum[w]; q = fmax(q, smax[w]); }
    d.red[blockIdx.x] = s;               // d.red is free outside the cooperative kernels
    d.red[kCostBlocks + blockIdx.x] = q;
  }
}
__global__ void k_ruiz_cost(const DevPtrs d, int nblocks) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0, q = 0.0;
    for (int b = 0; b < nblocks; b++) { s += d.red[b]; q = fmax(q, d.red[kCostBlocks + b]); }
    double ct = s / (double)d.n;
    ct = limit_scaling(fmax(ct, limit_scaling(q)));
    ct = 1.0 / ct;
    d.dtmp[0] = ct;  // dtmp is free between Ruiz passes
    d.state->c *= ct;
  }
}

__global__ void k_ruiz_cost_apply(const DevPtrs d) {
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  const double ct = d.dtmp[0];
  for (long long k = tid; k < d.P.nnz; k += nth) d.P.val[k] *= ct;
  for (long long j = tid; j < d.n; j += nth) d.q[j] *= ct;
}

// Dinv, Einv, cinv, P diagonal, scaled bounds
__global__ void k_scale_finish(const DevPtrs d, int scaling) {
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  for (long long j = tid; j < d.n; j += nth) {
    if (scaling) d.Dinv[j] = 1.0 / d.D[j];
    double dg = 0.0;
    for (int k = d.P.rowptr[j]; k < d.P.rowptr[j + 1]; k++)
      if (d.P.col[k] == j) dg += d.P.val[k];
    d.Pdiag[j] = dg;
  }
  for (long long i = tid; i < d.m; i += nth) {
    if (scaling) d.Einv[i] = 1.0 / d.E[i];
    d.l[i] = scaling ? d.E[i] * d.l0[i] : d.l0[i];
    d.u[i] = scaling ? d.E[i] * d.u0[i] : d.u0[i];
  }
  if (tid == 0) d.state->cinv = 1.0 / d.state->c;
}

// q <- c D q0 ; l,u <- E l0, E u0
__global__ void k_scale_vectors(const DevPtrs d, int do_q, int do_bounds, int scaling) {
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  const double c = d.state->c;
  if (do_q)
    for (long long j = tid; j < d.n; j += nth) d.q[j] = scaling ? (d.q0[j] * d.D[j]) * c : d.q0[j];
  if (do_bounds)
    for (long long i = tid; i < d.m; i += nth) {
      d.l[i] = scaling ? d.E[i] * d.l0[i] : d.l0[i];
      d.u[i] = scaling ? d.E[i] * d.u0[i] : d.u0[i];
    }
}

// set_rho_vec / update_rho_vec (row a3)
__global__ void k_set_rho_vec(const DevPtrs d, double rho) {
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  for (long long i = tid; i < d.m; i += nth) {
    const double li = d.l[i], ui = d.u[i];
    int t;
    double r;
    if (li < -kInfty * kMinScaling && ui > kInfty * kMinScaling) { t = -1; r = kRhoMin; }
    else if (ui - li < kRhoTol) { t = 1; r = kRhoEqOverIneq * rho; }
    else { t = 0; r = rho; }
    d.ctype[i] = t;
    d.rho_vec[i] = r;
    d.rho_inv[i] = 1.0 / r;
  }
}
__global__ void k_apply_rho(const DevPtrs d, double rho) {
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  for (long long i = tid; i < d.m; i += nth) {
    const int t = d.ctype[i];
    if (t == 0) { d.rho_vec[i] = rho; d.rho_inv[i] = 1.0 / rho; }
    else if (t == 1) { d.rho_vec[i] = kRhoEqOverIneq * rho; d.rho_inv[i] = 1.0 / (kRhoEqOverIneq * rho); }
  }
}

__global__ void k_precond(const DevPtrs d, double sigma) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long j = warp; j < d.n; j += nwarps) {
    double s = 0.0;
    if (d.m > 0)
      for (int k = d.At.rowptr[j] + lane; k < d.At.rowptr[j + 1]; k += 32) {
        const double a = d.At.val[k];
        s += d.rho_vec[d.At.col[k]] * a * a;
      }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) d.Minv[j] = 1.0 / (d.Pdiag[j] + sigma + s);
  }
}

// osqp_warm_start (row a14): x <- Dinv x, y <- c Einv y, z <- A x ; PCG guess x_tilde := x
__global__ void k_warm_start_xy(const DevPtrs d, const double *x_in, const double *y_in, int scaling) {
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  const double c = d.state->c;
  if (x_in)
    for (long long j = tid; j < d.n; j += nth) {
      const double v = scaling ? x_in[j] * d.Dinv[j] : x_in[j];
      d.x[j] = v;
      d.xt[j] = v;
    }
  if (y_in)
    for (long long i = tid; i < d.m; i += nth) d.y[i] = scaling ? (y_in[i] * d.Einv[i]) * c : y_in[i];
  if (tid == 0) d.state->needs_refresh = 1;
}
__global__ void k_warm_start_z(const DevPtrs d) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long i = warp; i < d.m; i += nwarps) {
    double s = 0.0;
    for (int k = d.A.rowptr[i] + lane; k < d.A.rowptr[i + 1]; k += 32) s += d.A.val[k] * d.x[d.A.col[k]];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) { d.z[i] = s; d.zt[i] = s; }
  }
}
__global__ void k_cold_start(const DevPtrs d) {
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  for (long long j = tid; j < d.n; j += nth) { d.x[j] = 0.0; d.xt[j] = 0.0; d.dx[j] = 0.0; }
  for (long long i = tid; i < d.m; i += nth) { d.z[i] = 0.0; d.y[i] = 0.0; d.zt[i] = 0.0; d.dy[i] = 0.0; }
}

// dst[map ? map[idx[k]] : idx[k]] = vals[k]   (idx == nullptr: identity)
__global__ void k_scatter(double *dst, const double *vals, const long long *idx, const int *map, long long k) {
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  for (long long e = tid; e < k; e += nth) {
    const long long src = idx ? idx[e] : e;
    const long long pos = map ? (long long)map[src] : src;
    if (pos >= 0) dst[pos] = vals[e];
  }
}

// tile streams <- scaled CSR values (padding entries stay zero)
__global__ void k_fill_blocked(const DevPtrs d) {
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  // (val is chunk-interleaved, osqp_abi.cu stream_val_pos; val32 is in entry order: invert the interleave)
  auto entry_of = [](int vp) { return (vp & ~127) + (((vp & 63) >> 1) << 2) + (((vp >> 6) & 1) << 1) + (vp & 1); };
  for (long long k = tid; k < d.A.nnz; k += nth) {
    const int pa = d.SA.from_csr[k], pt = d.ST.from_csr[k];
    const double a = d.A.val[k], t = d.At.val[k];
    d.SA.val[pa] = a;
    d.ST.val[pt] = t;
    if (d.mat32) {
      d.SA.val32[entry_of(pa)] = (float)a;
      d.ST.val32[entry_of(pt)] = (float)t;
    }
  }
  for (long long k = tid; k < d.P.nnz; k += nth) {
    const int pp = d.SA.from_csr[d.A.nnz + k];
    const double v = d.P.val[k];
    d.SA.val[pp] = v;
    if (d.mat32) d.SA.val32[entry_of(pp)] = (float)v;
  }
}

// compact copies of the Woodbury member rows <- scaled CSR values of A
__global__ void k_fill_wood(const DevPtrs d) {
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  const long long nnzW = d.W.w > 0 ? d.W.rp[d.W.w] : 0;
  for (long long k = tid; k < nnzW; k += nth) {
    d.W.val[k] = d.A.val[d.W.src[k]];
    d.W.tval[k] = d.A.val[d.W.tsrc[k]];
  }
}

inline int ew_grid(long long work) {
  long long g = (work + 255) / 256;
  if (g < 1) g = 1;
  if (g > 148 * 8) g = 148 * 8;
  return (int)g;
}

#endif  // !OSQP_B200_FAST

template <typename... Args>
cudaError_t coop_launch(void (*kernel)(Args...), unsigned *bar, LaunchGeom g, cudaStream_t st, Args... args) {
  cudaError_t e = cudaMemsetAsync(bar, 0, kBarBytes, st);  // grid barrier arrival counter + fixed-point accumulators
  if (e != cudaSuccess) return e;
  void *params[] = {(void *)&args...};
  if (g.cluster <= 1)
    return cudaLaunchCooperativeKernel((const void *)kernel, dim3(g.grid), dim3(g.block), params, g.dyn_smem, st);
  // cooperative (all blocks co-resident: the hand-written grid barrier) AND clustered (pairs share DSMEM)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(g.grid);
  cfg.blockDim = dim3(g.block);
  cfg.dynamicSmemBytes = g.dyn_smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  at[1].id = cudaLaunchAttributeClusterDimension;
  at[1].val.clusterDim.x = g.cluster;
  at[1].val.clusterDim.y = 1;
  at[1].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 2;
  return cudaLaunchKernelExC(&cfg, (const void *)kernel, params);
}

}  // namespace

#ifdef OSQP_B200_FAST
// ------------------------------------------------------------------ host wrappers of the fixed-mode kernels
#if OSQP_B200_FAST == 1
#define FASTNAME(x) x##_fast
#elif OSQP_B200_FAST == 2
#define FASTNAME(x) x##_fast2
#else
#define FASTNAME(x) x##_fast3
#endif
cudaError_t FASTNAME(launch_solve)(const DevPtrs &d, const SolveCfg &cfg, LaunchGeom g, cudaStream_t st) {
  return coop_launch(admm_kernel, d.bar, g, st, d, cfg);
}
cudaError_t FASTNAME(launch_polish)(const DevPtrs &d, const PolishCfg &cfg, const SolveCfg &sc, PolishOut *out, LaunchGeom g,
                                    cudaStream_t st) {
  return coop_launch(polish_kernel, d.bar, g, st, d, cfg, sc, out);
}
void FASTNAME(kernels)(const void **admm, const void **polish) {
  *admm = (const void *)admm_kernel;
  *polish = (const void *)polish_kernel;
}
#else
// ------------------------------------------------------------------ host wrappers
// The modes the other compilations fix (see the top of this file): tile streams in the lane-row layout, fp32 slices and
// fp32 copies of the matrix values in the PCG phases (DevPtrs::mat32), update_info on the streams, no Woodbury rows;
// 1 = [A; P] in cluster pairs, Jacobi (kernels_fast.cu), 2 = no pairs, Jacobi (kernels_fast2.cu: one column group, or
// more than two), 3 = no pairs, slack-elimination preconditioner (kernels_fast3.cu); 0 = none of them, the plain
// kernels.  Evaluated at every launch, so a workspace that loses its cluster pairs (osqp_abi.cu
// launch_with_pair_fallback) moves on.
int fast_mode(const DevPtrs &d, const LaunchGeom &g) {
  if (!(g.fast && d.blocked && d.SA.lane_rows && (d.m == 0 || d.ST.lane_rows) && d.f32_slices && d.info_streams &&
        d.W.w == 0 && d.mat32))
    return 0;
  if (d.SL.rows > 0) return (!d.SA.paired && g.cluster == 1) ? 3 : 0;
  if (d.SA.paired) return g.cluster == 2 ? 1 : 0;
  return g.cluster == 1 ? 2 : 0;
}

// the cooperative kernels of all compilations (attributes and occupancy are set / taken over all of them)
constexpr int kCoopKernels = 8;
int coop_kernel_list(const void **f) {
  f[0] = (const void *)admm_kernel;
  f[1] = (const void *)polish_kernel;
  kernels_fast(&f[2], &f[3]);
  kernels_fast2(&f[4], &f[5]);
  kernels_fast3(&f[6], &f[7]);
  return kCoopKernels;
}

cudaError_t launch_scale_data(const DevPtrs &d, int scaling_iters, double sigma, cudaStream_t st) {
  const long long rows = (long long)d.n + d.m;
  const int gw = ew_grid(rows * 32), ge = ew_grid(d.A.nnz + d.P.nnz + rows);
  k_copy_vals<<<ge, 256, 0, st>>>(d);
  for (int it = 0; it < scaling_iters; it++) {
    k_ruiz_norms<<<gw, 256, 0, st>>>(d);
    k_ruiz_apply<<<gw, 256, 0, st>>>(d);
    k_ruiz_cost_partial<<<kCostBlocks, 256, 0, st>>>(d);
    k_ruiz_cost<<<1, 32, 0, st>>>(d, kCostBlocks);
    k_ruiz_cost_apply<<<ge, 256, 0, st>>>(d);
  }
  k_scale_finish<<<ew_grid(rows), 256, 0, st>>>(d, scaling_iters > 0 ? 1 : 0);
  return cudaGetLastError();
}

cudaError_t launch_scale_vectors(const DevPtrs &d, int do_q, int do_bounds, cudaStream_t st) {
  // `scaling` on/off is encoded in D/E/c being 1, so the scaled formulas are always right
  k_scale_vectors<<<ew_grid((long long)d.n + d.m), 256, 0, st>>>(d, do_q, do_bounds, 1);
  return cudaGetLastError();
}

cudaError_t launch_set_rho_vec(const DevPtrs &d, double rho, int, cudaStream_t st) {
  if (d.m > 0) k_set_rho_vec<<<ew_grid(d.m), 256, 0, st>>>(d, rho);
  return cudaGetLastError();
}

cudaError_t launch_apply_rho(const DevPtrs &d, double rho, cudaStream_t st) {
  if (d.m > 0) k_apply_rho<<<ew_grid(d.m), 256, 0, st>>>(d, rho);
  return cudaGetLastError();
}

cudaError_t launch_precond(const DevPtrs &d, double sigma, cudaStream_t st) {
  k_precond<<<ew_grid((long long)d.n * 32), 256, 0, st>>>(d, sigma);
  return cudaGetLastError();
}

cudaError_t launch_pd_probe(const DevPtrs &d, LaunchGeom g, double sigma, int max_it, int trials, cudaStream_t st) {
  g.dyn_smem = 0;
  g.cluster = 1;
  return coop_launch(pd_probe_kernel, d.bar, g, st, d, sigma, max_it, trials);
}

cudaError_t launch_gershgorin(const DevPtrs &d, double sigma, cudaStream_t st) {
  k_gershgorin<<<ew_grid((long long)d.n * 32), 256, 0, st>>>(d, sigma);
  return cudaGetLastError();
}

cudaError_t launch_warm_start(const DevPtrs &d, const double *x_in, const double *y_in, int scaling, cudaStream_t st) {
  k_warm_start_xy<<<ew_grid((long long)d.n + d.m), 256, 0, st>>>(d, x_in, y_in, scaling);
  if (x_in && d.m > 0) k_warm_start_z<<<ew_grid((long long)d.m * 32), 256, 0, st>>>(d);
  return cudaGetLastError();
}

cudaError_t launch_cold_start(const DevPtrs &d, cudaStream_t st) {
  k_cold_start<<<ew_grid((long long)d.n + d.m), 256, 0, st>>>(d);
  return cudaGetLastError();
}

cudaError_t launch_scatter_values(double *dst, const double *vals, const long long *idx, const int *map, long long k,
                                  cudaStream_t st) {
  if (k > 0) k_scatter<<<ew_grid(k), 256, 0, st>>>(dst, vals, idx, map, k);
  return cudaGetLastError();
}

cudaError_t launch_solve(const DevPtrs &d, const SolveCfg &cfg, LaunchGeom g, cudaStream_t st) {
  switch (fast_mode(d, g)) {
    case 1: return launch_solve_fast(d, cfg, g, st);
    case 2: return launch_solve_fast2(d, cfg, g, st);
    case 3: return launch_solve_fast3(d, cfg, g, st);
    default: return coop_launch(admm_kernel, d.bar, g, st, d, cfg);
  }
}

cudaError_t launch_spmv(const DevPtrs &d, int which, const double *in, double *out, double sigma, LaunchGeom g,
                        cudaStream_t st) {
  if (which >= 10 || !d.blocked) {  // CSR + L1-gather path
    spmv_kernel<<<g.grid, g.block, 0, st>>>(d, which % 10, in, out, sigma);
    return cudaGetLastError();
  }
  return coop_launch(spmv_stream_kernel, d.bar, g, st, d, which, in, out, sigma);
}

#ifdef OSQP_B200_DEVTOOLS
cudaError_t launch_reduce_selftest(const DevPtrs &d, LaunchGeom g, double ref, double *out, cudaStream_t st) {
  g.dyn_smem = 0;
  g.cluster = 1;
  return coop_launch(reduce_selftest_kernel, d.bar, g, st, d, ref, out);
}

cudaError_t launch_barrier_bench(const DevPtrs &d, LaunchGeom g, int iters, int mode, double *sink,
                                 unsigned long long *ns_out, cudaStream_t st) {
  g.dyn_smem = 0;
  g.cluster = 1;
  return coop_launch(barrier_bench_kernel, d.bar, g, st, d, iters, mode, sink, ns_out);
}

#endif  // OSQP_B200_DEVTOOLS

// How many clusters of `csize` thread blocks of admm_kernel (with `dyn_smem` bytes each) can be co-resident.
int max_active_clusters(int csize, int block, size_t dyn_smem) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(csize * 148);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = dyn_smem;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = csize;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (csize > 8) cudaFuncSetAttribute(admm_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, (const void *)admm_kernel, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  return n;
}

#ifdef OSQP_B200_DEVTOOLS
cudaError_t launch_membench(const void *buf, long long bytes, int pattern, int depth, int grid, double *sink,
                            cudaStream_t st) {
  const char *b = reinterpret_cast<const char *>(buf);
  if (depth <= 2) membench_kernel<2><<<grid, kThreads, 0, st>>>(b, bytes, pattern, sink);
  else if (depth <= 4) membench_kernel<4><<<grid, kThreads, 0, st>>>(b, bytes, pattern, sink);
  else if (depth <= 6) membench_kernel<6><<<grid, kThreads, 0, st>>>(b, bytes, pattern, sink);
  else membench_kernel<8><<<grid, kThreads, 0, st>>>(b, bytes, pattern, sink);
  return cudaGetLastError();
}

#endif  // OSQP_B200_DEVTOOLS

// cudaFuncAttributeMaxDynamicSharedMemorySize is per function and per device for the whole process and the last
// call wins: a second workspace with a smaller slice must never lower the cap under an older workspace that still
// launches with its larger one.  The attribute is therefore only ever raised (process-wide bookkeeping per device).
cudaError_t raise_dyn_smem(const void *func, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<int, const void *>, size_t> have;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  size_t &cur = have[std::make_pair(dev, func)];
  if (bytes <= cur) return cudaSuccess;
  e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) cur = bytes;
  return e;
}

cudaError_t configure_dyn_smem(size_t dyn_smem) {
  const void *f[kCoopKernels];
  const int nf = coop_kernel_list(f);
  for (int k = 0; k < nf; k++) {
    cudaError_t e = raise_dyn_smem(f[k], dyn_smem);
    if (e != cudaSuccess) return e;
  }
  return raise_dyn_smem((const void *)spmv_stream_kernel, dyn_smem);
}

int coop_threads() { return kThreads; }

size_t coop_static_smem() {
  const void *f[kCoopKernels];
  const int nf = coop_kernel_list(f);
  size_t most = 0;
  for (int k = 0; k < nf; k++) {
    cudaFuncAttributes a{};
    if (cudaFuncGetAttributes(&a, f[k]) != cudaSuccess) {
      cudaGetLastError();
      return 16384;
    }
    most = a.sharedSizeBytes > most ? a.sharedSizeBytes : most;
  }
  return most;
}

int max_coop_blocks_per_sm(int block, size_t dyn_smem) {
  const void *f[kCoopKernels];
  const int nf = coop_kernel_list(f);
  int least = 1 << 30;
  for (int k = 0; k < nf; k++) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, f[k], block, dyn_smem) != cudaSuccess) return 0;
    least = nb < least ? nb : least;
  }
  return least;
}

cudaError_t launch_fill_blocked(const DevPtrs &d, cudaStream_t st) {
  if (d.blocked) k_fill_blocked<<<ew_grid(d.A.nnz + d.P.nnz), 256, 0, st>>>(d);
  return cudaGetLastError();
}

cudaError_t launch_fill_wood(const DevPtrs &d, cudaStream_t st) {
  if (d.W.w > 0) k_fill_wood<<<148 * 4, 256, 0, st>>>(d);
  return cudaGetLastError();
}

cudaError_t launch_polish(const DevPtrs &d, const PolishCfg &cfg, const SolveCfg &sc, PolishOut *out, LaunchGeom g,
                          cudaStream_t st) {
  switch (fast_mode(d, g)) {
    case 1: return launch_polish_fast(d, cfg, sc, out, g, st);
    case 2: return launch_polish_fast2(d, cfg, sc, out, g, st);
    case 3: return launch_polish_fast3(d, cfg, sc, out, g, st);
    default: return coop_launch(polish_kernel, d.bar, g, st, d, cfg, sc, out);
  }
}
#endif  // OSQP_B200_FAST

}  // namespace osqpb200
