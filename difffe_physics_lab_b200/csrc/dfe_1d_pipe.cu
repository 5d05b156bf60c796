// Software-pipelined single-pass fused 1-D solve / adjoint for chain meshes (one Neumann sweep), sm_100a.
//
// Same mathematics as dfe_1d.cu / dfe_1d_split.cu (reference: diffhe/solver.py:73-98, :153-183 and their
// autograd backward; K = M + E, x = M^{-1}F - M^{-1}E M^{-1}F with flux-form prefix sums), organised so that
//   * every row is read from HBM once and written once (16 B/node forward, 24 B/node adjoint: the
//     algorithmic bytes), like the persistent kernel k_solve1d, and
//   * no CTA ever waits for another one in its critical path, like the split kernels:
// a sample spans G CTAs (chunk c of the mesh is owned by CTA c of a group for the whole batch, its mesh
// constants live in REGISTERS), and every CTA works on three different samples per iteration —
//     phase A (sample it)          rhs, local prefix sums, publish the chunk totals of sweep 0
//     phase B (sample it-LB)       x0 from the folded totals, rhs1 = -delta*x0, local sums, publish sweep 1
//     phase C (sample it-LB-LC)    x = x0 + x1, output
// so the cross-CTA exchange of sample s (two doubles per chunk and sweep through L2) has LB / LC whole
// iterations to complete.  Intermediate per-node state (W, then z = x0 - c*W1) stays in the shared-memory
// ring slot the row chunk was loaded into by TMA; the slot is finally stored back with one bulk store.
//
// Prefix sums are kept in MOMENT form: a block of nodes contributes (s, m) with m = w - s*X_end, which makes
// every combine a plain addition: W_i = W^thread_i + sum_before(m) + X_i * sum_before(s).  Scans are therefore
// two-component add-scans, chunk folds are plain sums; the affine part a + b*X_i is carried along the thread's nodes
// incrementally (X_i = coordinate in half element lengths, from the mesh handle).  The scan over the threads of a CTA
// is done by the fold warps: a compute thread stores its (s, m) in shared memory and reads its CTA-wide exclusive
// prefix back one iteration later (ncu: the two warp-shuffle scans per iteration were 29 % of the compute warps'
// stall samples).
//
// Roles inside a CTA: W compute warps (never touch global memory); one fold warp per sweep (publishes the CTA's
// totals, polls/folds the other chunks' totals — data-as-flag: the exchange buffer is pre-set to an all-ones
// sentinel, no counters, no fences); one I/O warp (TMA bulk loads and stores).  The roles are coupled only through
// shared-memory mbarriers (slot full / slot written / totals ready / constants ready): there is no CTA-wide barrier
// in the main loop, so a role waits exactly for the event it depends on and the fold / I/O latencies stay off the
// compute warps' critical path.
#include <cstdint>
#include <cstdlib>

#include "dfe_internal.h"

namespace {

#include "dfe_1d_common.cuh"

constexpr unsigned long long SENT = 0xFFFFFFFFFFFFFFFFull;   // "not yet published"
constexpr int NLMAX = 4;                                       // chunks per sample <= 32*NLMAX
constexpr int MISC_BYTES = 4096;

struct PP {
  int nn, G, chg, NG, slotd;
  long long B;
  const double* hs;      // h_e/2
  const double* rh;      // RN(1/(h_e/2))
  const double* X;       // X_i = sum_{e<i} h_e/2
  const double* in0;     // forward: f ; backward: gbar
  long long ld0;
  const double* in1;     // backward: u
  long long ld1;
  double* out;           // forward: u ; backward: dL/df (may be null)
  long long ldo;
  const double* kappa;
  const double* ck;      // (2/kappa, kappa/2) per sample (or one pair, shared kappa)
  int per_sample;
  int bcL, bcR;
  double gL, gR, Xtot;
  unsigned long long* part;   // [B][2][G][2] chunk totals (s, m) as bit patterns, pre-set to SENT
  double* gkpart;             // backward: [B][G+2] partial dL/dkappa (chunks, boundary terms of sweep 0 and 1)
  double* losspart;           // misfit adjoint: [B][G] partial sum_i (u_i - u_data_i)^2
  double mf_scale;            // misfit adjoint: gbar = mf_scale * (u - u_data), formed on the fly (in0 = u_data)
  int* err;                   // mesh handle's device fault word: set to 1 if a wait exceeded its bound
  int backoff;                // cycles a fold-warp lane waits between two polls of a chunk total (DFE_PIPE_BACKOFF)
};

__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void publish(unsigned long long* slot, double s, double m) {
  unsigned long long a = static_cast<unsigned long long>(__double_as_longlong(s));
  unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(m));
  if (a == SENT) a = 0x7FF8000000000000ull;   // a NaN that happens to carry the sentinel payload
  if (b == SENT) b = 0x7FF8000000000000ull;
  st_relaxed_u64(slot, a);
  st_relaxed_u64(slot + 1, b);
}
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// k_i = fl(kappa/h_i) from the stored reciprocal (Markstein correction: correctly rounded)
__device__ __forceinline__ double kdiv(double kaph, double h, double y) {
  const double q0 = kaph * y;
  return fma(fma(-h, q0, kaph), y, q0);
}
// err = (a + b) - fl(a + b), exact (TwoSum)
__device__ __forceinline__ double two_sum_err(double a, double b) {
  const double d = __dadd_rn(a, b);
  const double bb = __dsub_rn(d, a);
  return __dadd_rn(__dsub_rn(a, __dsub_rn(d, bb)), __dsub_rn(b, bb));
}
// x += (value of lane - d), only where that lane exists: the shuffle's own predicate guards the add (3 instructions)
__device__ __forceinline__ void scan_step(double& x, int d) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b32 lo, hi, olo, ohi;\n\t.reg .f64 o;\n\t"
      "mov.b64 {lo, hi}, %0;\n\t"
      "shfl.sync.up.b32 olo|p, lo, %1, 0, 0xffffffff;\n\t"
      "shfl.sync.up.b32 ohi, hi, %1, 0, 0xffffffff;\n\t"
      "mov.b64 o, {olo, ohi};\n\t"
      "@p add.rn.f64 %0, %0, o;\n\t}"
      : "+d"(x)
      : "r"(d));
}
// 16-byte poll of one (s, m) pair; tearing is harmless: each word is individually "sentinel or final"
__device__ __forceinline__ void ld_pair(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void poll_pair(const unsigned long long* p, unsigned long long a, unsigned long long b,
                                          double& s, double& m, int* err, int* dead, int backoff) {
  if (a == SENT || b == SENT) {
    int spins = 0;
    while (true) {
      ld_pair(p, a, b);
      if (a != SENT && b != SENT) break;
      if (*reinterpret_cast<volatile int*>(dead)) { a = b = 0; break; }
      if (++spins > (1 << 22)) {
        *reinterpret_cast<volatile int*>(dead) = 1;
        *reinterpret_cast<volatile int*>(err) = 1;
        a = b = 0;
        break;
      }
      if (backoff > 0) {   // busy wait on the clock: keeps the spinning lanes off the L2 lines the publishers store to
        const long long t0 = clock64();
        while (clock64() - t0 < backoff) {}
      }
      // no __nanosleep here: measured on B200, a sleeping poller occasionally oversleeps by ~4.5 us, and every such
      // hiccup stalls the whole group of CTAs two iterations later (1.7x on the config-2 step); the L2 round trip of
      // the load itself paces the loop
    }
  }
  s = __longlong_as_double(static_cast<long long>(a));
  m = __longlong_as_double(static_cast<long long>(b));
}

// ---- bounded waits: a protocol bug must never hang the GPU.  After LIMIT cycles a waiter raises the CTA-wide
// `dead` flag (and the global error word); every later wait returns at once, the results are garbage and
// dfe_solve1d reports the error.
constexpr long long WAIT_LIMIT = 4000000000ll;   // ~2 s at 1.965 GHz
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __noinline__ void mbar_wait_slow(uint64_t* bar, uint32_t parity, int* err, int* dead) {
  const long long t0 = clock64();
  int spins = 0;
  while (!mbar_try(bar, parity)) {
    if ((++spins & 63) == 0) {
      if (*reinterpret_cast<volatile int*>(dead)) return;
      if (clock64() - t0 > WAIT_LIMIT) {
        *reinterpret_cast<volatile int*>(dead) = 1;
        *reinterpret_cast<volatile int*>(err) = 1;
        return;
      }
    }
  }
}
__device__ __forceinline__ void mbar_wait_b(uint64_t* bar, uint32_t parity, int* err, int* dead) {
  if (!mbar_test(bar, parity)) mbar_wait_slow(bar, parity, err, dead);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- the three per-thread phases.  FULL: every one of the R nodes exists and none is a Dirichlet node (no masks).
// The coordinate X_j of node j (in half element lengths) is never held per node: a + b*X_j is carried along as
// t_{j+1} = t_j + b*hs[j+1] (one fma per node, the one the direct evaluation would need as well).
// MF (misfit adjoint): the ring slot holds u_data; gbar_j = sc (u_j - u_data_j) is formed here, kept in the slot for
// phase B, and sum_j (u_j - u_data_j)^2 over every node of the chunk (Dirichlet nodes included) goes to Lq.
template <bool BWD, int R, bool FULL, bool MF>
__device__ __forceinline__ void phase_a(double* buf, const double* ub, const double (&hs)[R + 1], int nin, int nst,
                                        bool ownsL, double sc, double& S, double& Wc, double& Lq) {
#pragma unroll
  for (int j = 0; j < R; ++j) {
    double in;
    if (MF) {
      const double diff = (FULL || j < nst) ? ub[j] - buf[j] : 0.0;
      Lq = fma(diff, diff, Lq);
      in = (FULL || j < nin) ? sc * diff : 0.0;
      if (!FULL && j == 0 && ownsL) in = 0.0;
      if (FULL || j < nst) buf[j] = in;
    } else {
      in = (FULL || j < nin) ? buf[j] : 0.0;
    }
    // forward: F_i = h_{i-1}/2 f_i + h_i/2 f_i (solver.py:95-96); backward: gbar on the free rows
    double F = BWD ? in : fma(in, hs[j + 1], in * hs[j]);
    if (!FULL && j == 0 && ownsL) F = 0.0;
    if (!BWD && (FULL || j < nst)) buf[j] = Wc;   // thread-local W at node j, kept in place for phase B
    S += F;
    Wc = fma(hs[j + 1], S, Wc);
  }
}

// SK (kappa shared by the batch): delta_i does not depend on the sample; it was computed once per thread and sits in
// rh[i + 1] (the reciprocals are not needed any more), which takes 9 of the 15 flops per node out of this phase.
// KEEP = false (no row is written by the call): z is not stored.
template <bool BWD, int R, bool FULL, bool SK, bool KEEP>
__device__ __forceinline__ void phase_b(double* buf, const double* ub, const double (&hs)[R + 1], const double (&rh)[R + 1],
                                        double X0, int nin, int nst, bool ownsL, double c0, double kaph, double a0,
                                        double b0, double& S1, double& W1, double& D) {
  double S = 0.0, Wc = 0.0;
  double kp = SK ? 0.0 : kdiv(kaph, hs[0], rh[0]);
  double t = fma(b0, X0, a0);
#pragma unroll
  for (int j = 0; j < R; ++j) {
    double g = 0.0, Wt;
    if (BWD) {
      g = (FULL || j < nin) ? buf[j] : 0.0;
      if (!FULL && j == 0 && ownsL) g = 0.0;
      Wt = Wc;
      S += g;
      Wc = fma(hs[j + 1], S, Wc);
    } else {
      Wt = (FULL || j < nst) ? buf[j] : 0.0;
    }
    const double x0 = fma(-c0, Wt, t);
    t = fma(b0, hs[j + 1], t);
    // k_i = fl(kappa/h_i) bit-exactly (solver.py:88); err = (k_{i-1}+k_i) - fl(k_{i-1}+k_i) is minus the rounding of
    // the reference's diagonal accumulation (solver.py:89-92); it is 0 on Dirichlet and padding rows.
    double v1;
    if (SK) {
      v1 = __dmul_rn(rh[j + 1], x0);
    } else {
      const double ki = kdiv(kaph, hs[j + 1], rh[j + 1]);
      v1 = __dmul_rn(two_sum_err(kp, ki), x0);
      kp = ki;
    }
    if (BWD) {
      const double uj = (FULL || j < nst) ? ub[j] : 0.0;
      D = fma(g + v1, uj, D);
    }
    if (KEEP && (FULL || j < nst)) buf[j] = fma(-c0, W1, x0);   // z = x0 - c*W1(thread-local), finished in phase C
    S1 += v1;
    W1 = fma(hs[j + 1], S1, W1);
  }
}

template <bool BWD, int R, bool FULL>
__device__ __forceinline__ void phase_c(double* buf, const double (&hs)[R + 1], double X0, int nst, double a1, double b1) {
  double t = fma(b1, X0, a1);
#pragma unroll
  for (int j = 0; j < R; ++j) {
    if (FULL || j < nst) {
      const double xv = buf[j] + t;
      // forward: u[free] = x (solver.py:180-181); backward: dL/df_i = lambda_i (h_{i-1}/2 + h_i/2)
      buf[j] = BWD ? fma(xv, hs[j], xv * hs[j + 1]) : xv;
    }
    t = fma(b1, hs[j + 1], t);
  }
}

// Shared-memory control block of one CTA.  All synchronisation between the roles goes through these mbarriers
// (no CTA-wide barrier inside the main loop): a role only ever waits for the event it really depends on.
template <int W, int NR>
struct Ctl {
  uint64_t full[NR];    // TMA load of ring slot k landed                       (tx bytes)     I/O -> compute
  uint64_t outr[NR];    // phase C finished writing ring slot k                 (W arrivals)   compute -> I/O
  uint64_t ufull[4];    // backward: u-row slot landed                          (tx bytes)     I/O -> compute
  uint64_t ufree[4];    // backward: phase B finished reading the u-row slot    (W arrivals)   compute -> I/O
  uint64_t tot[2][2];   // [sweep][it & 1] per-thread totals of this iteration  (W arrivals)   compute -> fold warp
  uint64_t cf[2][2];    // [sweep][it & 1] folded constants for iteration `it`  (1 arrival)    fold warp -> compute
  double cfa[2][2][2];      // [sweep][it & 1] (A, B) constants of the CTA
  double red[2][W];         // backward: [it & 1][warp] dot partials of phase B
  double redL[2][W];        // misfit adjoint: [it & 1][warp] partial sums of (u - u_data)^2 of phase A
  double sc[2][4];          // [it & 1] c0, kappa/2, c1
  int dead;
};

// One fold warp per sweep ST (0 or 1).  In iteration `it` it
//   * waits for the compute threads' (s, m) totals of this iteration's phase (A for sweep 0, B for sweep 1), turns
//     them IN PLACE into CTA-wide exclusive prefixes (lane l scans the W consecutive threads l*W .. l*W+W-1 serially,
//     one shuffle scan over the 32 lane totals — the compute warps execute no shuffle at all) and publishes the CTA
//     total of that sample;
//   * folds the G chunk totals of the sample the NEXT iteration consumes (published about one iteration ago by
//     every CTA of the group; its loads are issued before the wait): boundary constants -> (A, B) of the CTA ->
//     shared memory, then signals cf[ST][(it + 1) & 1].
// ST is a run-time value: both fold warps share ONE copy of this code (instruction-cache footprint).
template <bool BWD, int W, int NR, int LB, int LC, bool MF>
__device__ __forceinline__ void fold_warp(const PP& p, Ctl<W, NR>* ctl, double2* pt, int lane, int c, int grp, int nIt,
                                          int nTot, const int ST) {
  constexpr int WP = W | 1;                     // odd row pitch: conflict-free 16-byte accesses from both sides
  const int LAG = ST == 0 ? LB : LB + LC;       // the sample folded in iteration it is it + 1 - LAG
  const int G = p.G, GP = G + 2;
  const bool LR = p.bcL && p.bcR;
  const double rXtot = 1.0 / p.Xtot;
  const bool lift = (!BWD && ST == 0);   // harmonic interpolant of the Dirichlet data: forward, sweep 0 only
  const double ga = (lift && LR) ? (p.gR - p.gL) * rXtot : 0.0;
  const double gb = lift ? (p.bcL ? p.gL : p.gR) : 0.0;
  const long long sstep = p.NG;
  long long sF = grp + static_cast<long long>(1 - LAG) * sstep;   // sample folded in iteration 0 (may be negative)
  long long sP = grp + static_cast<long long>(ST == 0 ? 0 : -LB) * sstep;   // sample published in iteration 0
  const bool ends = BWD && c == 0;
  // constants of the sample folded in the current iteration, fetched one iteration ahead
  double ccur = 0.0, khcur = 0.0, uLcur = 0.0, uRcur = 0.0;
  {
    const int f = 1 - LAG;
    if (f >= 0 && f < nIt) {
      const double* ck = p.ck + (p.per_sample ? 2 * sF : 0);
      ccur = ck[0];
      khcur = ck[1];
      if (ends) { uLcur = p.in1[sF * p.ld1]; uRcur = p.in1[sF * p.ld1 + p.nn - 1]; }
    }
  }
  if (lane == 0) mbar_arrive(&ctl->cf[ST][0]);   // iteration 0 consumes no folded constants

  for (int it = 0; it < nTot; ++it) {
    const int par = it & 1, nxt = par ^ 1;
    const uint32_t ph = (it >> 1) & 1;
    const int f = it + 1 - LAG;
    const bool vf = f >= 0 && f < nIt;
    const unsigned long long* base = p.part + ((sF * 2 + ST) * G) * 2;
    // ---- loads first: chunk totals of sample f, constants of sample f+1
    unsigned long long ra[3], rb[3];
#pragma unroll
    for (int l = 0; l < 3; ++l) {
      const int ci = lane + 32 * l;
      ra[l] = rb[l] = SENT;
      if (vf && ci < G) ld_pair(base + 2 * ci, ra[l], rb[l]);
    }
    double cnx = 0.0, khnx = 0.0, uLnx = 0.0, uRnx = 0.0;
    if (f + 1 >= 0 && f + 1 < nIt) {
      const long long sN = sF + sstep;
      const double* ck = p.ck + (p.per_sample ? 2 * sN : 0);
      cnx = ck[0];
      khnx = ck[1];
      if (ends) { uLnx = p.in1[sN * p.ld1]; uRnx = p.in1[sN * p.ld1 + p.nn - 1]; }
    }
    // ---- totals of this iteration's phase: exclusive prefixes in place, publish the CTA total
    mbar_wait_b(&ctl->tot[ST][par], ph, p.err, &ctl->dead);
    {
      double2* P = pt + static_cast<size_t>(ST * 2 + par) * (32 * WP) + lane * WP;
      double es[W], em[W];
      double rs = 0.0, rm = 0.0;
#pragma unroll
      for (int k = 0; k < W; ++k) {
        const double2 v = P[k];
        es[k] = rs;
        em[k] = rm;
        rs += v.x;
        rm += v.y;
      }
      double is = rs, im = rm;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        scan_step(is, d);
        scan_step(im, d);
      }
      double xs = __shfl_up_sync(0xffffffffu, is, 1);
      double xm = __shfl_up_sync(0xffffffffu, im, 1);
      if (lane == 0) { xs = 0.0; xm = 0.0; }
#pragma unroll
      for (int k = 0; k < W; ++k) P[k] = make_double2(es[k] + xs, em[k] + xm);
      const int jp = ST == 0 ? it : it - LB;
      if (lane == 31 && jp >= 0 && jp < nIt) publish(p.part + ((sP * 2 + ST) * G + c) * 2, is, im);
      if (MF && ST == 0) {    // misfit partial of the sample phase A handled in this iteration
        double a = (lane < W) ? ctl->redL[par][lane < W ? lane : 0] : 0.0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
        if (jp >= 0 && jp < nIt && lane == 0) p.losspart[sP * G + c] = a;
      }
      if (BWD && ST == 1) {   // dot partial of the sample phase B handled in this iteration
        double a = (lane < W) ? ctl->red[par][lane < W ? lane : 0] : 0.0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
        if (jp >= 0 && jp < nIt && lane == 0) p.gkpart[sP * GP + c] = a;
      }
    }
    // ---- fold
    if (vf) {
      double Sx = 0.0, Mx = 0.0, St = 0.0, Mt = 0.0;
#pragma unroll
      for (int l = 0; l < 3; ++l) {
        const int ci = lane + 32 * l;
        if (ci < G) {
          double sv, mv;
          poll_pair(base + 2 * ci, ra[l], rb[l], sv, mv, p.err, &ctl->dead, p.backoff);
          St += sv; Mt += mv;
          if (ci < c) { Sx += sv; Mx += mv; }
        }
      }
      for (int l0 = 96; l0 < G; l0 += 32) {   // more than 96 chunks per sample (rare): extra trips
        const int ci = lane + l0;
        if (ci < G) {
          double sv, mv;
          poll_pair(base + 2 * ci, SENT, SENT, sv, mv, p.err, &ctl->dead, p.backoff);
          St += sv; Mt += mv;
          if (ci < c) { Sx += sv; Mx += mv; }
        }
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        Sx += __shfl_xor_sync(0xffffffffu, Sx, d);
        Mx += __shfl_xor_sync(0xffffffffu, Mx, d);
        St += __shfl_xor_sync(0xffffffffu, St, d);
        Mt += __shfl_xor_sync(0xffffffffu, Mt, d);
      }
      const double cc = ccur;
      const double Wtot = fma(St, p.Xtot, Mt);                   // sum_e h_e/2 S_e over the whole sample
      const double C = LR ? Wtot * rXtot : (p.bcL ? St : 0.0);    // flux constant fixed by the boundary conditions
      const double A = gb + (p.bcL ? 0.0 : cc * Wtot) - cc * Mx;
      const double Bc = ga + cc * (C - Sx);
      if (lane == 0) {
        ctl->cfa[ST][nxt][0] = A;
        ctl->cfa[ST][nxt][1] = Bc;
        ctl->sc[nxt][2 * ST] = cc;
        if (ST == 0) ctl->sc[nxt][1] = khcur;
        // backward, chunk 0: boundary terms of sum_e q_e (u_{e+1}-u_e) = C (u_R-u_L) - S_tot u_R + sum_i rhs_i u_i
        if (ends) p.gkpart[sF * GP + G + ST] = C * (uRcur - uLcur) - St * uRcur;
      }
    }
    ccur = cnx; khcur = khnx; uLcur = uLnx; uRcur = uRnx;
    sF += sstep;
    sP += sstep;
    __syncwarp();
    if (lane == 0) mbar_arrive(&ctl->cf[ST][nxt]);
  }
}

// Warps 0..W-1 compute; warp W / W+1 = fold warps of sweep 0 / 1; warp W+2 = I/O warp (TMA loads and stores).
// OUT = false (adjoint without dL/df): no row is written, phase C has nothing to do and a ring slot is released at the
// end of phase B — LC fewer ring slots, i.e. room for a larger chunk per CTA.
// PF = ring slots beyond the LB + LC + 1 a row occupies from phase A to phase C plus the one being stored / refilled:
// 1 (default) gives a refill two iterations to land, 0 gives it one and leaves room for a larger chunk.
template <bool BWD, int R, int W, int LB, int LC, bool SK, bool MF, bool OUT, int PF>
__global__ void __launch_bounds__(32 * (W + 3), (W >= 8 ? 1 : 2)) k1d_pipe(const PP p) {
  constexpr int LCO = OUT ? LC : 0;         // iterations a ring slot lives on after phase B
  constexpr int NR = LB + LCO + 2 + PF;     // ring slots of the in-place chain: prefetch, A..C, store drain
  // ring slots of the u row (backward): read by phase B only, or — misfit adjoint — by phase A and phase B
  constexpr int NRU = MF ? LB + 2 : 2;
  static_assert(W <= 16 && LB >= 1 && LC >= 1 && NRU <= 4 && (!MF || BWD) && (OUT || BWD), "layout");
  static_assert(sizeof(Ctl<W, NR>) <= MISC_BYTES, "misc region");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Ctl<W, NR>* ctl = reinterpret_cast<Ctl<W, NR>*>(smem_raw);
  constexpr int WP = W | 1;
  double* ring = reinterpret_cast<double*>(smem_raw + MISC_BYTES);
  double* uring = ring + static_cast<size_t>(NR) * p.slotd;
  // per-thread (s, m) totals -> exclusive prefixes, [sweep][it & 1][32 * WP]: thread t at (t / W) * WP + t % W
  double2* pt = reinterpret_cast<double2*>(uring + static_cast<size_t>(BWD ? NRU : 0) * p.slotd);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = p.G, nn = p.nn;
  const int c = blockIdx.x % G, grp = blockIdx.x / G;
  const int n0 = c * p.chg;
  const int len = min(nn, n0 + p.chg) - n0;
  const int nIt = static_cast<int>((p.B - grp + p.NG - 1) / p.NG);
  const int nTot = nIt + LB + LC;
  const bool have_out = OUT && (p.out != nullptr);
  const int slotd = p.slotd;
  // 16-byte phase of a row chunk: element j of the chunk lives at slot[mis + j]; mis depends on the sample only
  // through its parity (when the leading dimension is odd)
  const int mis0c = static_cast<int>(((reinterpret_cast<uintptr_t>(p.in0) >> 3) + n0) & 1);
  const int mis1c = BWD ? static_cast<int>(((reinterpret_cast<uintptr_t>(p.in1) >> 3) + n0) & 1) : 0;
  const int ld0p = static_cast<int>(p.ld0 & 1), ld1p = static_cast<int>(p.ld1 & 1), ngp = p.NG & 1, grpp = grp & 1;
  auto mis0_of = [&](int j) { return mis0c ^ (ld0p & (grpp ^ (j & ngp))); };
  auto mis1_of = [&](int j) { return mis1c ^ (ld1p & (grpp ^ (j & ngp))); };

  if (tid == 0) {
    for (int k = 0; k < NR; ++k) { mbar_init_raw(&ctl->full[k], 1); mbar_init_raw(&ctl->outr[k], W); }
    for (int k = 0; k < 4; ++k) {
      mbar_init_raw(&ctl->ufull[k], 1);
      mbar_init_raw(&ctl->ufree[k], W);
    }
    for (int k = 0; k < 2; ++k)
      for (int s = 0; s < 2; ++s) { mbar_init_raw(&ctl->tot[s][k], W); mbar_init_raw(&ctl->cf[s][k], 1); }
    fence_mbar_init();
    ctl->dead = 0;
  }
  __syncthreads();

  if (warp < W) {
    // ============================================================================ compute warps
    const int tb = tid * R;
    const int nin = max(0, min(R, min(len, nn - (p.bcR ? 1 : 0) - n0) - tb));   // nodes whose input is read
    const int nst = max(0, min(R, len - tb));                                  // nodes of the chunk held
    const bool ownsL = p.bcL && c == 0 && tid == 0;
    const bool ownsR = p.bcR && c == G - 1 && tb <= len - 1 && len - 1 < tb + R;
    // one variant per WARP (uniform branch): the mask-free code if every lane's R nodes are ordinary free nodes
    const bool full = __all_sync(0xffffffffu, (nin == R) && (nst == R) && !ownsL && !ownsR);
    // mesh constants of this thread's R nodes: hs[j] / rh[j] = element left of node j (hs[R]: right of the last)
    double hs[R + 1], rh[R + 1];
#pragma unroll
    for (int j = 0; j <= R; ++j) {
      const int e = n0 - 1 + tb + j;
      const bool ex = (e >= 0 && e < nn - 1 && tb + j <= len);
      hs[j] = ex ? p.hs[e] : 0.0;
      rh[j] = ex ? p.rh[e] : 0.0;
    }
    if (SK) {   // delta_i = (k_{i-1} + k_i) - fl(k_{i-1} + k_i) for the one kappa of the batch, kept in rh[i + 1]
      const double kaph = 0.5 * p.kappa[0];
      double kp = kdiv(kaph, hs[0], rh[0]);
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const double ki = kdiv(kaph, hs[j + 1], rh[j + 1]);
        rh[j + 1] = two_sum_err(kp, ki);
        kp = ki;
      }
    }
    const double X0 = p.X[min(n0 + tb, nn - 1)];
    const double Xe = p.X[min(n0 + min(tb + R, len), nn - 1)];   // where the block after this thread starts
    // this thread's CTA-wide exclusive prefixes in flight (computed by the fold warps, read back one iteration later)
    double q0s[LB - 1], q0m[LB - 1], q1s[LC - 1], q1m[LC - 1];
#pragma unroll
    for (int k = 0; k < LB - 1; ++k) q0s[k] = q0m[k] = 0.0;
#pragma unroll
    for (int k = 0; k < LC - 1; ++k) q1s[k] = q1m[k] = 0.0;
    const int pidx = (tid / W) * WP + tid % W;
    double2* pt0 = pt + pidx;                 // [it & 1] at pt0 + par * 32 * WP
    double2* pt1 = pt + 2 * 32 * WP + pidx;

    int slotA = 0, roundA = 0;   // ring slot of phase A's sample and its use count parity
    for (int it = 0; it < nTot; ++it) {
      const int par = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      // ---------------------------------------------------------------- phase A (sample it)
      {
        double S = 0.0, Wc = 0.0, Lq = 0.0;
        if (it < nIt) {
          double* buf = ring + slotA * slotd + mis0_of(it) + tb;
          mbar_wait_b(&ctl->full[slotA], roundA, p.err, &ctl->dead);
          const double* ub = nullptr;
          if (MF) {
            const int us = it % NRU;
            ub = uring + us * slotd + mis1_of(it) + tb;
            mbar_wait_b(&ctl->ufull[us], (it / NRU) & 1, p.err, &ctl->dead);
          }
          if (full) phase_a<BWD, R, true, MF>(buf, ub, hs, nin, nst, ownsL, p.mf_scale, S, Wc, Lq);
          else phase_a<BWD, R, false, MF>(buf, ub, hs, nin, nst, ownsL, p.mf_scale, S, Wc, Lq);
        }
        pt0[par * 32 * WP] = make_double2(S, fma(-S, Xe, Wc));
        if (MF) {
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) Lq += __shfl_xor_sync(0xffffffffu, Lq, d);
          if (lane == 0) ctl->redL[par][warp] = Lq;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl->tot[0][par]);
      }
      // ---------------------------------------------------------------- phase B (sample it-LB)
      {
        const int jB = it - LB;
        int slotB = slotA - LB;
        if (slotB < 0) slotB += NR;
        double S1 = 0.0, W1 = 0.0, D = 0.0;
        mbar_wait_b(&ctl->cf[0][par], ph, p.err, &ctl->dead);
        const bool act = jB >= 0 && jB < nIt;
        if (act) {
          double* buf = ring + slotB * slotd + mis0_of(jB) + tb;
          const double c0 = ctl->sc[par][0], kaph = ctl->sc[par][1];
          const double a0 = fma(-c0, q0m[LB - 2], ctl->cfa[0][par][0]);
          const double b0 = fma(-c0, q0s[LB - 2], ctl->cfa[0][par][1]);
          const double* ub = nullptr;
          if (BWD) {
            const int us = jB % NRU;
            ub = uring + us * slotd + mis1_of(jB) + tb;
            mbar_wait_b(&ctl->ufull[us], (jB / NRU) & 1, p.err, &ctl->dead);
          }
          if (full) phase_b<BWD, R, true, SK, OUT>(buf, ub, hs, rh, X0, nin, nst, ownsL, c0, kaph, a0, b0, S1, W1, D);
          else phase_b<BWD, R, false, SK, OUT>(buf, ub, hs, rh, X0, nin, nst, ownsL, c0, kaph, a0, b0, S1, W1, D);
          if (!OUT && MF) fence_async_smem();   // phase A wrote gbar into the slot the next TMA load overwrites
        }
        {   // the prefix of the sample phase A handled in the previous iteration is ready (cf[0] of this iteration
            // is signalled after the fold warp finished that iteration): keep it for phase B of iteration it-1+LB
          const double2 e0 = pt0[(par ^ 1) * 32 * WP];
#pragma unroll
          for (int k = LB - 2; k > 0; --k) { q0s[k] = q0s[k - 1]; q0m[k] = q0m[k - 1]; }
          q0s[0] = e0.x;
          q0m[0] = e0.y;
        }
        pt1[par * 32 * WP] = make_double2(S1, fma(-S1, Xe, W1));
        if (BWD) {
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) D += __shfl_xor_sync(0xffffffffu, D, d);
          if (lane == 0) ctl->red[par][warp] = D;
        }
        __syncwarp();   // every lane's shared-memory accesses of this phase are done
        if (lane == 0) {
          mbar_arrive(&ctl->tot[1][par]);
          if (BWD && act) mbar_arrive(&ctl->ufree[jB % NRU]);
          if (!OUT && act) mbar_arrive(&ctl->outr[slotB]);   // nothing is written back: the slot can be refilled
        }
      }
      // ---------------------------------------------------------------- phase C (sample it-LB-LC)
      {
        const int jC = it - LB - LC;
        mbar_wait_b(&ctl->cf[1][par], ph, p.err, &ctl->dead);
        if (OUT && jC >= 0 && jC < nIt) {
          int slotC = slotA - LB - LC;
          if (slotC < 0) slotC += NR;
          if (have_out) {
            double* buf = ring + slotC * slotd + mis0_of(jC) + tb;
            const double c1 = ctl->sc[par][2];
            const double a1 = fma(-c1, q1m[LC - 2], ctl->cfa[1][par][0]);
            const double b1 = fma(-c1, q1s[LC - 2], ctl->cfa[1][par][1]);
            if (full) {
              phase_c<BWD, R, true>(buf, hs, X0, nst, a1, b1);
            } else {
              phase_c<BWD, R, false>(buf, hs, X0, nst, a1, b1);
              if (ownsL) buf[0] = BWD ? 0.0 : p.gL;                 // u[d] = g (solver.py:177-179); dL/df = 0 there
              if (ownsR) buf[len - 1 - tb] = BWD ? 0.0 : p.gR;
            }
            fence_async_smem();
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&ctl->outr[slotC]);
        }
        {   // prefix of the sample phase B handled in the previous iteration (ready: cf[1] of this iteration was signalled
            // after the fold warp finished that iteration): keep it for phase C of iteration it-1+LC
          const double2 e1 = pt1[(par ^ 1) * 32 * WP];
#pragma unroll
          for (int k = LC - 2; k > 0; --k) { q1s[k] = q1s[k - 1]; q1m[k] = q1m[k - 1]; }
          q1s[0] = e1.x;
          q1m[0] = e1.y;
        }
      }
      if (++slotA == NR) { slotA = 0; roundA ^= 1; }
    }
  } else if (warp <= W + 1) {
    fold_warp<BWD, W, NR, LB, LC, MF>(p, ctl, pt, lane, c, grp, nIt, nTot, warp - W);
  } else {
    // ============================================================================ I/O warp
    // Row chunks are loaded as the 16-byte aligned superset [g - mis, g + len rounded up): the element before /
    // after the chunk belongs to the same array (the host checks the base alignment; the very last element of
    // the array is fetched separately), so a load is ONE bulk copy and element j lands at slot[mis + j].
    const long long last_s = p.B - 1;
    auto load_row = [&](const double* rowbase, long long ld, long long s, double* dst, uint64_t* mb) {
      const double* g = rowbase + s * ld + n0;
      const int mis = static_cast<int>((reinterpret_cast<uintptr_t>(g) >> 3) & 1);
      int cnt = (mis + len + 1) & ~1;
      const bool clip = (s == last_s && c == G - 1 && ((mis + len) & 1));   // would read past the end of the array
      if (clip) cnt -= 2;
      if (lane == 0) {
        if (clip) dst[mis + len - 1] = g[len - 1];   // before the arrive: its release orders this store
        mbar_arrive_expect_tx(mb, 8u * static_cast<uint32_t>(cnt));
        if (cnt) bulk_g2s(dst, g - mis, 8u * cnt, mb);
      }
    };
    for (int j = 0; j < NR && j < nIt; ++j)
      load_row(p.in0, p.ld0, grp + static_cast<long long>(j) * p.NG, ring + j * slotd, &ctl->full[j]);
    if (BWD)
      for (int j = 0; j < NRU && j < nIt; ++j)
        load_row(p.in1, p.ld1, grp + static_cast<long long>(j) * p.NG, uring + j * slotd, &ctl->ufull[j]);

    int slotC = 0, roundC = 0;
    for (int it = LB; it < nTot; ++it) {
      // ---- backward: the u-row slot phase B released in this iteration takes the row two samples ahead
      if (BWD) {
        const int jb = it - LB;
        if (jb < nIt) {
          const int us = jb % NRU;
          mbar_wait_b(&ctl->ufree[us], (jb / NRU) & 1, p.err, &ctl->dead);
          if (jb + NRU < nIt)
            load_row(p.in1, p.ld1, grp + static_cast<long long>(jb + NRU) * p.NG, uring + us * slotd, &ctl->ufull[us]);
        }
      }
      // ---- store the row phase C finished in this iteration, then refill its slot
      const int jc = it - LB - LCO;
      if (jc >= 0 && jc < nIt) {
        mbar_wait_b(&ctl->outr[slotC], roundC, p.err, &ctl->dead);
        if (have_out) {
          const long long s = grp + static_cast<long long>(jc) * p.NG;
          double* go = p.out + s * p.ldo + n0;
          const double* ssrc = ring + slotC * slotd;
          const Seg qo = make_seg(go, len);
          if (lane == 0) {
            if (qo.body) bulk_s2g(go + qo.head, ssrc + qo.mis + qo.head, 8u * qo.body);
            bulk_commit();
          } else if (lane == 1) {
            if (qo.head) go[0] = ssrc[qo.mis];
          } else if (lane == 2) {
            if (qo.tail) go[len - 1] = ssrc[qo.mis + len - 1];
          }
          if (lane == 0) bulk_wait_read0();   // the slot is reloaded right away
          __syncwarp();
        }
        if (jc + NR < nIt)
          load_row(p.in0, p.ld0, grp + static_cast<long long>(jc + NR) * p.NG, ring + slotC * slotd, &ctl->full[slotC]);
        if (++slotC == NR) { slotC = 0; roundC ^= 1; }
      }
    }
    if (lane == 0) bulk_wait_all0();
  }
}

// (2/kappa, kappa/2) per sample (or once, shared kappa): keeps every division out of the pipelined kernel.
// Also clears the ticket of k1d_pipe_gk (the workspace arrives with undefined contents).
__global__ void k1d_pipe_ck(const double* kappa, long long n, double* ck, unsigned* ticket) {
  const long long s = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (s == 0) *ticket = 0u;
  if (s < n) {
    const double k = kappa[s];
    ck[2 * s] = 2.0 / k;
    ck[2 * s + 1] = 0.5 * k;
  }
}

// dL/dkappa = -(1/kappa) * (sum of the per-(sample, chunk) partials + boundary terms) and, for the misfit adjoint,
// loss = lsc * sum of the per-(sample, chunk) squared misfits — fixed summation order, no float atomics:
//   per-sample kappa : one thread per sample;
//   shared kappa     : block b reduces samples [b*GK_SPB, (b+1)*GK_SPB) with a fixed tree into blk[b]; the block that
//                      draws the last ticket adds blk[0..nb) in index order (lane-strided, then a fixed butterfly) and
//                      writes out[0] (= sum dL/dkappa) and loss_out[0] — for config 5 these two doubles are adjacent
//                      words of the buffer the NCCL all-reduce runs on, so no copy kernel sits between the two.
constexpr int GK_SPB = 1024;
__global__ void __launch_bounds__(256) k1d_pipe_gk(const double* part, const double* lpart, const double* kappa,
                                                   long long B, int GP, int G, int per_sample, double lsc, double* out,
                                                   double* loss_out, double* blk, unsigned* ticket) {
  if (per_sample) {
    const long long s = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (s < B) {
      double a = 0.0;
      for (int g = 0; g < GP; ++g) a += part[s * GP + g];
      out[s] = -a / kappa[s];
      if (lpart) {
        double q = 0.0;
        for (int c = 0; c < G; ++c) q += lpart[s * G + c];
        loss_out[s] = lsc * q;
      }
    }
    return;
  }
  __shared__ double sh[2][256];
  __shared__ int is_last;
  const int tid = threadIdx.x;
  double a = 0.0, l = 0.0;
  for (int k = 0; k < GK_SPB / 256; ++k) {
    const long long s = static_cast<long long>(blockIdx.x) * GK_SPB + k * 256 + tid;
    if (s < B) {
      double r = 0.0;
      for (int g = 0; g < GP; ++g) r += part[s * GP + g];
      a += r;
      if (lpart) {
        double q = 0.0;
        for (int c = 0; c < G; ++c) q += lpart[s * G + c];
        l += q;
      }
    }
  }
  sh[0][tid] = a;
  sh[1][tid] = l;
  __syncthreads();
  for (int d = 128; d > 0; d >>= 1) {
    if (tid < d) { sh[0][tid] += sh[0][tid + d]; sh[1][tid] += sh[1][tid + d]; }
    __syncthreads();
  }
  if (tid == 0) {
    blk[2 * blockIdx.x] = sh[0][0];
    blk[2 * blockIdx.x + 1] = sh[1][0];
    __threadfence();
    is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (is_last && tid < 32) {
    __threadfence();
    double A = 0.0, Lt = 0.0;
    for (unsigned b = tid; b < gridDim.x; b += 32) {
      A += __ldcg(blk + 2 * b);
      Lt += __ldcg(blk + 2 * b + 1);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      A += __shfl_xor_sync(0xffffffffu, A, d);
      Lt += __shfl_xor_sync(0xffffffffu, Lt, d);
    }
    if (tid == 0) {
      out[0] = -A / kappa[0];
      if (loss_out) loss_out[0] = lsc * Lt;
      *ticket = 0u;
    }
  }
}

// A wait that exceeded its bound leaves garbage behind: make it loud.  Runs after every fused 1-D launch as ONE CTA;
// reads the handle's device fault word and, only if it is set, raises the host-visible word and overwrites every
// output of the call with NaN (slow, but this is the fault path).
__global__ void k1d_pipe_poison(const int* err, int* err_host, double* out, long long ldo, long long B, int nn,
                                double* gk, long long ngk, double* loss, long long nloss) {
  __shared__ int bad;
  if (threadIdx.x == 0) bad = *reinterpret_cast<const volatile int*>(err);
  __syncthreads();
  if (!bad) return;
  if (threadIdx.x == 0) *reinterpret_cast<volatile int*>(err_host) = 1;
  const double nan = __longlong_as_double(0x7FF8000000000000ll);
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long t0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (out)
    for (long long i = t0; i < B * nn; i += stride) out[(i / nn) * ldo + i % nn] = nan;
  if (gk)
    for (long long i = t0; i < ngk; i += stride) gk[i] = nan;
  if (loss)
    for (long long i = t0; i < nloss; i += stride) loss[i] = nan;
}

struct Geo {
  int G, chg, NG, slotd;
  size_t smem;
};

template <bool BWD, int R, int W, int LB, int LC, bool SK = false, bool MF = false, bool OUT = true, int PF = 1>
int run_cfg(const dfe_mesh* m, PP p, cudaStream_t st, int gbound, int* G_used, unsigned* ticket) {
  constexpr int NR = LB + (OUT ? LC : 0) + 2 + PF, NRU = BWD ? (MF ? LB + 2 : 2) : 0;
  constexpr int CAP = R * 32 * W, THREADS = 32 * (W + 3);
  auto kern = k1d_pipe<BWD, R, W, LB, LC, SK, MF, OUT, PF>;
  const int nn = p.nn;
  // per device / context, so set on every call (a process may drive several GPUs); it is a host-side table update
  DFE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  // geometry: the throughput is proportional to the number of groups NG resident at once (every CTA does the
  // same work per iteration), so take the largest NG whose chunk size fits the shared memory at the assumed
  // occupancy; within that NG use as many chunks as there are CTA slots (smaller slots, no idle SM).
  Geo g{};
  bool found = false;
  for (int per_sm = 2; per_sm >= 1 && !found; --per_sm) {
    const long long slots = static_cast<long long>(per_sm) * m->sm_count;
    const long long gcap = slots < gbound ? slots : gbound;   // gbound: what the workspace was sized for
    const long long gmin = (nn + CAP - 1) / CAP;
    if (gmin > gcap) continue;
    long long NGmax = slots / gmin;
    if (NGmax > p.B) NGmax = p.B;
    for (long long NG = NGmax; NG >= 1 && !found; --NG) {
      long long G = slots / NG;
      if (G > gcap) G = gcap;
      if (G > nn) G = nn;
      if (p.B <= NG) G = gmin;   // tiny batches: nothing to gain from more chunks
      // whole warps of fully populated threads: only the last chunk of a sample has a ragged tail, so every
      // other warp runs the mask-free code (a partially filled warp would run the slower masked variant in
      // EVERY CTA and set the pace of the whole group)
      g.chg = static_cast<int>((nn + G - 1) / G);
      g.chg = ((g.chg + 32 * R - 1) / (32 * R)) * (32 * R);
      if (g.chg > CAP) g.chg = CAP;
      g.G = (nn + g.chg - 1) / g.chg;   // drop empty trailing chunks
      g.NG = static_cast<int>(NG);
      g.slotd = ((g.chg + 2) + 1) & ~1;
      g.smem = MISC_BYTES + sizeof(double) * static_cast<size_t>(NR + NRU) * g.slotd + 4 * 32 * (W | 1) * 16;
      if (g.smem > 227 * 1024) continue;
      int occ = 0;
      DFE_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, g.smem));
      found = occ >= per_sm;
    }
  }
  if (!found) {
    dfe::set_error("dfe_solve1d (pipelined): no resident configuration for a mesh of %d nodes", nn);
    return DFE_ERR_UNSUPPORTED;
  }
  p.G = g.G;
  p.chg = g.chg;
  p.NG = g.NG;
  p.slotd = g.slotd;
  *G_used = g.G;
  // exchange buffer [B][2][G][2] <- "not yet published"
  DFE_CUDA_OK(cudaMemsetAsync(p.part, 0xFF, static_cast<size_t>(p.B) * 2 * g.G * 2 * sizeof(double), st));
  {
    const long long nk = p.per_sample ? p.B : 1;
    k1d_pipe_ck<<<static_cast<unsigned>((nk + 255) / 256), 256, 0, st>>>(p.kappa, nk, const_cast<double*>(p.ck), ticket);
  }
  // cooperative launch: the CTAs of a group wait for each other's chunk totals, so all NG*G CTAs must be co-resident —
  // the runtime checks exactly that and fails the launch (instead of hanging) under MPS / SM partitioning
  static const bool plain_launch = getenv("DFE_PIPE_PLAIN_LAUNCH") != nullptr;   // A/B switch for measurements
  if (plain_launch) {
    kern<<<static_cast<unsigned>(g.NG * g.G), THREADS, g.smem, st>>>(p);
    DFE_CUDA_OK(cudaGetLastError());
  } else {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(g.NG * g.G));
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = g.smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    DFE_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, p));
  }
  return DFE_OK;
}

int cfg_id() {
  static const int id = [] {
    const char* e = getenv("DFE_PIPE_CFG");
    return e ? atoi(e) : 0;
  }();
  return id;
}

}  // namespace

namespace dfe {

// upper bound of the chunks per sample over every configuration (smallest thread capacity 5*256, and the
// geometry pass never more than doubles the minimum chunk count)
static size_t g_bound(const dfe_mesh* m) {
  const size_t gmin = (static_cast<size_t>(m->info.n_nodes) + 1279) / 1280;
  const size_t b = 2 * gmin + 2;
  return b < 32 * NLMAX ? b : 32 * NLMAX;
}

// workspace layout: exchange buffer [B][2][gmax][2] | gkpart [B][gmax+2] | ck [B][2] | losspart [B][gmax] |
// blk [B/GK_SPB+1][2] | ticket
struct WsLayout {
  size_t part, gkpart, ck, losspart, blk, ticket, total;
};
static WsLayout ws_layout(const dfe_mesh* m, long long B) {
  const size_t gmax = g_bound(m), b = static_cast<size_t>(B), d = sizeof(double);
  WsLayout L{};
  L.part = 0;
  L.gkpart = L.part + b * 2 * gmax * 2 * d;
  L.ck = L.gkpart + b * (gmax + 2) * d;
  L.losspart = L.ck + b * 2 * d;
  L.blk = L.losspart + b * gmax * d;
  L.ticket = L.blk + (b / GK_SPB + 1) * 2 * d;
  L.total = L.ticket + 512;
  return L;
}

size_t pipe1d_workspace_bytes(const dfe_mesh* m, long long B) { return ws_layout(m, B).total; }

// Can the pipelined kernel take this call?  (chain mesh, one Neumann sweep, scalar / per-sample kappa, and the
// output row of every sample has the 16-byte phase of its input row — the chain works in place in shared memory.)
bool pipe1d_eligible(const dfe_mesh* m, long long B, int kappa_mode, int n_refine, const double* in0, long long ld0,
                     const double* in1, const double* out, long long ldo) {
  if (!m->chain || n_refine != 1) return false;
  if (kappa_mode != DFE_KAPPA_SCALAR && kappa_mode != DFE_KAPPA_PER_SAMPLE) return false;
  if (m->info.n_nodes > 5LL * 256 * 32 * NLMAX) return false;
  if ((reinterpret_cast<uintptr_t>(in0) | reinterpret_cast<uintptr_t>(in1)) & 15) return false;   // aligned-superset TMA loads (see the I/O warp)
  if (out) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(in0) >> 3, b = reinterpret_cast<uintptr_t>(out) >> 3;
    if ((a ^ b) & 1) return false;
    if (B > 1 && ((ld0 ^ ldo) & 1)) return false;
  }
  return true;
}

// misfit != nullptr (adjoint only): in0 = u_data, gbar = misfit->scale * (u - u_data) is formed inside the kernel and
// misfit->loss receives (scale / 2) * sum_i (u_i - u_data_i)^2 — per sample, or summed over the batch for a shared kappa.
int pipe1d_run(const dfe_mesh* m, long long B, bool bwd, const double* in0, long long ld0, const double* in1,
               long long ld1, const double* kappa, int kappa_mode, double* out, long long ldo, double* gkappa,
               void* ws, cudaStream_t st, const Misfit1D* misfit) {
  PP p{};
  p.nn = static_cast<int>(m->info.n_nodes);
  p.B = B;
  p.hs = m->d_hs;
  p.rh = m->d_rh;
  p.X = m->d_X;
  p.Xtot = m->x_total;
  p.in0 = in0; p.ld0 = ld0; p.in1 = in1; p.ld1 = ld1; p.out = out; p.ldo = ldo;
  p.kappa = kappa;
  p.per_sample = kappa_mode == DFE_KAPPA_PER_SAMPLE;
  p.bcL = m->bc_left; p.bcR = m->bc_right; p.gL = m->g_left; p.gR = m->g_right;
  const size_t gmax = g_bound(m);
  unsigned char* w = static_cast<unsigned char*>(ws);
  const WsLayout L = ws_layout(m, B);
  p.part = reinterpret_cast<unsigned long long*>(w + L.part);
  p.gkpart = reinterpret_cast<double*>(w + L.gkpart);
  p.ck = reinterpret_cast<double*>(w + L.ck);
  p.losspart = reinterpret_cast<double*>(w + L.losspart);
  p.mf_scale = misfit ? misfit->scale : 0.0;
  double* blk = reinterpret_cast<double*>(w + L.blk);
  unsigned* ticket = reinterpret_cast<unsigned*>(w + L.ticket);
  p.err = m->d_fault_dev;   // sticky; k1d_pipe_poison turns a fault into NaN outputs and raises the host-visible word
  int rc, G = 0;
  const int id = cfg_id();
  static const int backoff = [] { const char* e = getenv("DFE_PIPE_BACKOFF"); return e ? atoi(e) : 0; }();
  p.backoff = backoff;
  static const bool no_sk = getenv("DFE_PIPE_NOSK") != nullptr;   // tuning switch: generic kernel for a shared kappa too
  const bool sk = !p.per_sample && !no_sk;
  const int gb = static_cast<int>(gmax);
  // <BWD, R nodes per thread, W compute warps, LB, LC>.  Default: one large CTA per SM (the compute warps share the
  // instruction stream and one set of service warps); DFE_PIPE_CFG selects the alternatives kept for tuning.
  if (!bwd) {
    switch (id) {
      case 1: rc = run_cfg<false, 11, 10, 2, 2>(m, p, st, gb, &G, ticket); break;
      case 4: rc = run_cfg<false, 9, 10, 3, 2>(m, p, st, gb, &G, ticket); break;
      // (PF = 0, one ring slot less: measured slower — config 2 forward 1.44 -> 1.64 ms at the same (9, 12) geometry and
      // 1.74 ms with the larger (11, 12) chunk it makes room for: a refill needs more than one iteration to land)
      default:
        // (the shared-kappa specialisation is used by the adjoint only: measured on config 5a it makes the forward
        // kernel slower and erratic, 3.94 -> 4.4-5.1 ms — its shorter phase B moves the fold latency onto the critical path)
        rc = run_cfg<false, 9, 12, 2, 2>(m, p, st, gb, &G, ticket);
        break;
    }
  } else if (misfit) {
    // The misfit adjoint holds LB + 2 u-row slots next to the ring.  Without dL/df nothing is written back and a ring
    // slot is released after phase B (5 ring slots): 11 nodes per thread fit.  With dL/df (7 ring slots) the chunk that
    // still fits the shared memory at 11 nodes per thread fills only 6 of the 8 compute warps (measured on config 5a:
    // 6.0 ms per launch vs 5.4 ms with 9; the groups-per-iteration-time figure NG / (R + overhead) decides).
    if (!out) {
      switch (id) {
        case 1: rc = !sk ? run_cfg<true, 9, 10, 2, 2, false, true, false>(m, p, st, gb, &G, ticket)
                         : run_cfg<true, 9, 10, 2, 2, true, true, false>(m, p, st, gb, &G, ticket); break;
        case 2: rc = !sk ? run_cfg<true, 9, 8, 2, 2, false, true>(m, p, st, gb, &G, ticket)
                         : run_cfg<true, 9, 8, 2, 2, true, true>(m, p, st, gb, &G, ticket); break;
        default: rc = !sk ? run_cfg<true, 11, 8, 2, 2, false, true, false>(m, p, st, gb, &G, ticket)
                          : run_cfg<true, 11, 8, 2, 2, true, true, false>(m, p, st, gb, &G, ticket); break;
      }
    } else {
      rc = !sk ? run_cfg<true, 9, 8, 2, 2, false, true>(m, p, st, gb, &G, ticket)
               : run_cfg<true, 9, 8, 2, 2, true, true>(m, p, st, gb, &G, ticket);
    }
  } else if (!out) {
    switch (id) {
      case 2: rc = !sk ? run_cfg<true, 11, 8, 2, 2>(m, p, st, gb, &G, ticket)
                       : run_cfg<true, 11, 8, 2, 2, true>(m, p, st, gb, &G, ticket); break;
      case 1: rc = !sk ? run_cfg<true, 11, 10, 2, 2, false, false, false>(m, p, st, gb, &G, ticket)
                       : run_cfg<true, 11, 10, 2, 2, true, false, false>(m, p, st, gb, &G, ticket); break;
      default: rc = !sk ? run_cfg<true, 11, 8, 2, 2, false, false, false>(m, p, st, gb, &G, ticket)
                        : run_cfg<true, 11, 8, 2, 2, true, false, false>(m, p, st, gb, &G, ticket); break;
    }
  } else {
    switch (id) {
      case 1: rc = run_cfg<true, 9, 11, 2, 2>(m, p, st, gb, &G, ticket); break;
      default:
        rc = !sk ? run_cfg<true, 11, 8, 2, 2>(m, p, st, gb, &G, ticket)
                 : run_cfg<true, 11, 8, 2, 2, true>(m, p, st, gb, &G, ticket);
        break;
    }
  }
  if (rc != DFE_OK) return rc;
  if (bwd) {
    const unsigned nb = p.per_sample ? static_cast<unsigned>((B + 255) / 256) : static_cast<unsigned>((B + GK_SPB - 1) / GK_SPB);
    k1d_pipe_gk<<<nb, 256, 0, st>>>(p.gkpart, misfit ? p.losspart : nullptr, kappa, B, G + 2, G, p.per_sample,
                                    0.5 * p.mf_scale, gkappa, misfit ? misfit->loss : nullptr, blk, ticket);
    DFE_CUDA_OK(cudaGetLastError());
  }
  const long long nk = p.per_sample ? B : 1;
  return poison1d_launch(m, out, ldo, B, p.nn, bwd ? gkappa : nullptr, nk, misfit ? misfit->loss : nullptr, nk, st);
}

int poison1d_launch(const dfe_mesh* m, double* out, long long ldo, long long B, int nn, double* gk, long long ngk,
                    double* loss, long long nloss, cudaStream_t st) {
  k1d_pipe_poison<<<1, 1024, 0, st>>>(m->d_fault_dev, m->d_fault, out, ldo, B, nn, gk, ngk, loss, nloss);
  DFE_CUDA_OK(cudaGetLastError());
  return DFE_OK;
}

}  // namespace dfe
