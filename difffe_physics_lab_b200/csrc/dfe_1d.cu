// Fused batched 1-D P1 Poisson solve (forward) and adjoint (backward) for chain meshes, sm_100a.
//
// Replaces diffhe/solver.py:73-98 (_solve_1d), :153-183 (_apply_bc_and_solve) and their autograd
// backward for B independent samples on one mesh.  K is never formed: per sample the kernel reads
// the forcing row once and writes the solution row once (HBM traffic = 16 B/node forward).
//
// Algorithm.  The reference matrix is K = M + E with
//   M = structured stiffness  sum_e k_e [[1,-1],[-1,1]],  k_e = fl(kappa/h_e)      (solver.py:88-92)
//   E = diag(delta_i),        delta_i = fl(k_{i-1}+k_i) - (k_{i-1}+k_i)  (the rounding of the
//                             reference's diagonal accumulation, an exact float64 number).
// M^{-1} rhs is two prefix sums (flux form):  S_i = sum_{m<=i} rhs_m,  W_{i+1} = W_i + w_i S_i,
// X_{i+1} = X_i + w_i (w_e = 1/k_e),  x_i = x0 + C X_i - W_i, with (x0, C) fixed by the boundary
// conditions from the totals.  The solution of the float64 system K x = F is the Neumann series
// x = sum_s (-M^{-1}E)^s M^{-1}F; |M^{-1}E| ~ 1e-7 at 1e5 nodes, so one extra sweep reaches 1e-14.
// The prefix sums run as thread-serial runs of R nodes held in registers + warp-shuffle scans +
// one cross-CTA exchange of 3 doubles per chunk through global memory (a sample spans G CTAs).
//
// Kernel shape: persistent grid of NG groups x G CTAs, all co-resident.  CTA `c` of a group always
// owns chunk c of the mesh (its element lengths stay in shared memory for the whole batch) and
// walks samples s = group, group+NG, ...  Rows move global<->shared with 1-D TMA bulk copies
// (cp.async.bulk + mbarrier); thread t owns R consecutive nodes (R odd -> conflict-free LDS.64).
#include <cstdint>

#include "dfe_internal.h"

namespace {

constexpr int R = 17;           // nodes per thread (odd: stride-R shared-memory reads are conflict-free)
constexpr int T = 256;          // threads per CTA
constexpr int CHP = R * T;      // chunk capacity in nodes (4352)
constexpr int NW = T / 32;
constexpr int MAX_STAGES = 4;   // structured solve + up to 3 Neumann sweeps

struct Tri { double s, x, w; };  // (sum rhs, sum w, sum w*S) of a block of nodes

__device__ __forceinline__ Tri tri_id() { return Tri{0.0, 0.0, 0.0}; }
// a block followed by b block
__device__ __forceinline__ Tri combine(const Tri& a, const Tri& b) {
  Tri r;
  r.s = a.s + b.s;
  r.x = a.x + b.x;
  r.w = fma(a.s, b.x, a.w + b.w);
  return r;
}
__device__ __forceinline__ Tri shfl_up_tri(const Tri& t, int d) {
  Tri r;
  r.s = __shfl_up_sync(0xffffffffu, t.s, d);
  r.x = __shfl_up_sync(0xffffffffu, t.x, d);
  r.w = __shfl_up_sync(0xffffffffu, t.w, d);
  return r;
}
__device__ __forceinline__ Tri shfl_tri(const Tri& t, int src) {
  Tri r;
  r.s = __shfl_sync(0xffffffffu, t.s, src);
  r.x = __shfl_sync(0xffffffffu, t.x, src);
  r.w = __shfl_sync(0xffffffffu, t.w, src);
  return r;
}
__device__ __forceinline__ Tri warp_incl_scan(Tri t, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    Tri o = shfl_up_tri(t, d);
    if (lane >= d) t = combine(o, t);
  }
  return t;
}

// ---- PTX helpers: mbarrier + 1-D TMA bulk copies -------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(sdst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// A row segment [g, g+len) of doubles is moved as: head (0/1 element, when g is only 8-byte
// aligned) + 16-byte aligned body (TMA bulk) + tail (0/1 element).  Element j lives at sbuf[mis+j]
// so that global and shared addresses share their 16-byte phase.
struct Seg {
  int mis, head, body, tail;
};
__device__ __forceinline__ Seg make_seg(const double* g, int len) {
  Seg q;
  q.mis = static_cast<int>((reinterpret_cast<uintptr_t>(g) >> 3) & 1);
  q.head = (len > 0) ? q.mis : 0;
  q.body = (len - q.head) & ~1;
  q.tail = len - q.head - q.body;
  return q;
}

struct P1D {
  int nn, G, chg, NG, n_refine;
  long long B;
  const double* x;       // node coordinates (nn)
  const double* in0;     // forward: f ; backward: gbar
  long long ld0;
  const double* in1;     // backward: u (saved forward output)
  long long ld1;
  double* out;           // forward: u ; backward: gf (may be null)
  long long ldo;
  const double* kappa;
  int per_sample;        // kappa index = per_sample ? s : 0
  int bcL, bcR, lift_left_first;
  double gL, gR;
  int* cnt;              // [B] arrival counters (zeroed per call)
  double* summ;          // [B][MAX_STAGES][G][4] chunk summaries
  double* gkpart;        // backward: [B][G] partial dL/dkappa
};

// Exclusive prefix (carry) of chunk c and the total over the G chunk summaries of one stage.
__device__ __forceinline__ void fold_chunks(const double* summ, int G, int c, int lane, Tri& carry, Tri& total) {
  Tri base = tri_id();
  carry = tri_id();
  for (int g0 = 0; g0 < G; g0 += 32) {
    const int g = g0 + lane;
    Tri t = tri_id();
    if (g < G) {
      t.s = __ldcg(summ + 4 * g + 0);
      t.x = __ldcg(summ + 4 * g + 1);
      t.w = __ldcg(summ + 4 * g + 2);
    }
    t = warp_incl_scan(t, lane);
    t = combine(base, t);
    Tri ex = shfl_up_tri(t, 1);
    if (lane == 0) ex = base;
    if (c >= g0 && c < g0 + 32) carry = shfl_tri(ex, c - g0);
    base = shfl_tri(t, 31);
  }
  total = base;
}

template <bool BWD>
__global__ void __launch_bounds__(T, 2) k_solve1d(const P1D p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  Tri* wtot = reinterpret_cast<Tri*>(smem_raw + 16);                 // [2][NW]
  double* red = reinterpret_cast<double*>(smem_raw + 16 + 2 * NW * sizeof(Tri));  // [NW]
  double* hs = reinterpret_cast<double*>(smem_raw + 512);            // hs[j] = h_{n0-1+j}, j = 0..CHP
  double* b0 = hs + (CHP + 2);                                       // f / gbar row chunk, later output staging
  double* b1 = b0 + (CHP + 4);                                       // backward: u row chunk (+1 halo)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = p.G, nn = p.nn;
  const int grp = blockIdx.x / G, c = blockIdx.x % G;
  const int n0 = c * p.chg;
  const int n1 = min(nn, n0 + p.chg);
  const int len = n1 - n0;
  const int tb = tid * R;
  const int nref = p.n_refine;

  // element lengths of this chunk, bit-identical to the reference's h_e = x_j - x_i (solver.py:84-85)
  for (int j = tid; j <= CHP; j += T) {
    const int e = n0 - 1 + j;
    hs[j] = (e >= 0 && e < nn - 1 && j <= len) ? __dsub_rn(p.x[e + 1], p.x[e]) : 0.0;
  }
  if (tid == 0) mbar_init(bar, 1);
  __syncthreads();

  uint32_t phase = 0;
  for (long long s = grp; s < p.B; s += p.NG) {
    const double* g0 = p.in0 + s * p.ld0 + n0;
    const Seg q0 = make_seg(g0, len);
    const int len1 = BWD ? len + (n1 < nn ? 1 : 0) : 0;
    const double* g1 = BWD ? p.in1 + s * p.ld1 + n0 : nullptr;
    const Seg q1 = BWD ? make_seg(g1, len1) : Seg{0, 0, 0, 0};
    if (tid == 0) {
      bulk_wait_read0();  // the previous sample's bulk store has finished reading b0
      if (q0.head) b0[q0.mis] = g0[0];
      if (q0.tail) b0[q0.mis + len - 1] = g0[len - 1];
      if (BWD) {
        if (q1.head) b1[q1.mis] = g1[0];
        if (q1.tail) b1[q1.mis + len1 - 1] = g1[len1 - 1];
      }
      mbar_arrive_expect_tx(bar, 8u * static_cast<uint32_t>(q0.body + q1.body));
      if (q0.body) bulk_g2s(b0 + q0.mis + q0.head, g0 + q0.head, 8u * q0.body, bar);
      if (BWD && q1.body) bulk_g2s(b1 + q1.mis + q1.head, g1 + q1.head, 8u * q1.body, bar);
    }
    const double kap = p.kappa[p.per_sample ? s : 0];
    const double invk = 1.0 / kap;
    mbar_wait(bar, phase);
    phase ^= 1u;

    double v[R], xa[R];
    // ---- right-hand side of the free rows
    {
      double hp = hs[tb];
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int li = tb + j, i = n0 + li;
        const double hi = hs[tb + j + 1];
        const bool isfree = (li < len) && !(i == 0 && p.bcL) && !(i == nn - 1 && p.bcR);
        double val = 0.0;
        if (isfree) {
          const double in = b0[q0.mis + li];
          if (!BWD) {
            // F_i = (0 + h_{i-1}/2*f_i) + h_i/2*f_i   (solver.py:95-96, element i-1 then element i)
            val = __dadd_rn(__dmul_rn(__dmul_rn(hp, 0.5), in), __dmul_rn(__dmul_rn(hi, 0.5), in));
            // Lifting F_free -= K[free, d]*g in dict order (solver.py:166-169); K[1,0] = -k_0 etc.
            // The lifted load k*g is ~1/h times larger than F: pushing it through the prefix sums
            // would cost 3-4 digits.  So the reference's lifted value is formed bit-exactly and the
            // exact product k*g is then taken out again (TwoProduct / Sterbenz): the scans see only
            // the small remainder, and the k*g part is solved in closed form (x_g below).
            const bool liftL = p.bcL && i == 1 && p.gL != 0.0, liftR = p.bcR && i == nn - 2 && p.gR != 0.0;
            if (liftL || liftR) {
              const double kL = liftL ? __ddiv_rn(kap, hp) : 0.0, kR = liftR ? __ddiv_rn(kap, hi) : 0.0;
              const double pL = __dmul_rn(kL, p.gL), pR = __dmul_rn(kR, p.gR);   // = -fl(K[f,d]*g)
              if (liftL && p.lift_left_first) val = __dadd_rn(val, pL);
              if (liftR) val = __dadd_rn(val, pR);
              if (liftL && !p.lift_left_first) val = __dadd_rn(val, pL);
              // val is now the reference's F_free entry; remainder = val - kL*gL - kR*gR
              if (liftL && !p.lift_left_first) val = __dsub_rn(val, pL);
              if (liftR) val = __dsub_rn(val, pR);
              if (liftL && p.lift_left_first) val = __dsub_rn(val, pL);
              val = __dsub_rn(val, __dadd_rn(__fma_rn(kL, p.gL, -pL), __fma_rn(kR, p.gR, -pR)));
            }
          } else {
            val = in;  // gbar restricted to free rows (Dirichlet entries dropped)
          }
        }
        v[j] = val;
        hp = hi;
      }
    }

    double gk = 0.0;
    const bool has_g = (p.bcL && p.gL != 0.0) || (p.bcR && p.gR != 0.0);
    for (int st = 0; st <= nref; ++st) {
      // ---- local prefix of this thread's run
      Tri t = tri_id();
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const double w = hs[tb + j + 1] * invk;
        t.s += v[j];
        t.w = fma(w, t.s, t.w);
        t.x += w;
      }
      // ---- CTA scan
      const Tri inc = warp_incl_scan(t, lane);
      const int par = (st & 1) * NW;
      if (lane == 31) wtot[par + warp] = inc;
      __syncthreads();
      Tri wc = tri_id();
      for (int w = 0; w < warp; ++w) wc = combine(wc, wtot[par + w]);
      Tri ex = shfl_up_tri(inc, 1);
      if (lane == 0) ex = tri_id();
      const Tri texcl = combine(wc, ex);
      // ---- publish chunk summary, wait for the G chunks of this sample, fold
      double* stage_summ = p.summ + ((s * MAX_STAGES + st) * G) * 4;
      if (tid == T - 1) {
        const Tri tot = combine(wc, inc);
        double* slot = stage_summ + 4 * c;
        slot[0] = tot.s;
        slot[1] = tot.x;
        slot[2] = tot.w;
        __threadfence();
        atomicAdd(p.cnt + s, 1);
      }
      if (lane == 0) {
        const int target = (st + 1) * G;
        while (ld_acquire(p.cnt + s) < target) __nanosleep(40);
      }
      __syncwarp();
      Tri carry, total;
      fold_chunks(stage_summ, G, c, lane, carry, total);
      double C, x0c;
      if (p.bcL && p.bcR) {
        x0c = 0.0;
        C = total.w / total.x;
      } else if (p.bcL) {
        x0c = 0.0;
        C = total.s;
      } else {
        C = 0.0;
        x0c = total.w;
      }
      const Tri tc = combine(carry, texcl);
      // ---- apply: e = M^{-1} rhs on this run; accumulate; next rhs = -delta*e
      double S = tc.s, X = tc.x, W = tc.w;
      double hp = hs[tb];
      double kp = (st < nref && hp > 0.0) ? __ddiv_rn(kap, hp) : 0.0;
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int li = tb + j, i = n0 + li;
        const double hi = hs[tb + j + 1];
        const bool isfree = (li < len) && !(i == 0 && p.bcL) && !(i == nn - 1 && p.bcR);
        S += v[j];
        double xi = isfree ? (fma(C, X, x0c) - W) : 0.0;
        if (!BWD && st == 0 && isfree && has_g) {
          // x_g = M^{-1}(lifted boundary loads): the discrete harmonic interpolant of the Dirichlet data
          xi += (p.bcL && p.bcR) ? fma(p.gR - p.gL, X / total.x, p.gL) : (p.bcL ? p.gL : p.gR);
        }
        if (BWD) {
          // dL/dkappa = -(1/kappa) sum_e q_e (u_{e+1}-u_e),  q_e = C - S_e the flux of lambda
          if (li < len && i < nn - 1) gk = fma(C - S, b1[q1.mis + li + 1] - b1[q1.mis + li], gk);
        }
        const double w = hi * invk;
        W = fma(w, S, W);
        X += w;
        xa[j] = (st == 0) ? xi : xa[j] + xi;
        if (st < nref) {
          const double ki = (hi > 0.0) ? __ddiv_rn(kap, hi) : 0.0;
          // d_i = fl(k_{i-1}+k_i) (solver.py:89-92 accumulation); err = (k_{i-1}+k_i) - d_i exactly
          const double d = __dadd_rn(kp, ki);
          const double bb = __dsub_rn(d, kp);
          const double err = __dadd_rn(__dsub_rn(kp, __dsub_rn(d, bb)), __dsub_rn(ki, bb));
          v[j] = isfree ? __dmul_rn(err, xi) : 0.0;
          kp = ki;
        }
        hp = hi;
      }
    }

    // ---- epilogue
    const bool have_out = (p.out != nullptr);
    double* go = have_out ? p.out + s * p.ldo + n0 : nullptr;
    const Seg qo = have_out ? make_seg(go, len) : Seg{0, 0, 0, 0};
    if (have_out) {
      double hp = hs[tb];
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int li = tb + j, i = n0 + li;
        const double hi = hs[tb + j + 1];
        if (li < len) {
          double o;
          if (!BWD) {
            // u[d] = g ; u[free] = x   (solver.py:177-181)
            o = (i == 0 && p.bcL) ? p.gL : ((i == nn - 1 && p.bcR) ? p.gR : xa[j]);
          } else {
            // dL/df_i = lambda_i (h_{i-1}/2 + h_i/2)   (autograd of solver.py:95-96)
            o = fma(xa[j], hp * 0.5, xa[j] * (hi * 0.5));
          }
          b0[qo.mis + li] = o;
        }
        hp = hi;
      }
      fence_async_smem();
    }
    if (BWD) {
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) gk += __shfl_xor_sync(0xffffffffu, gk, d);
      if (lane == 0) red[warp] = gk;
    }
    __syncthreads();
    if (tid == 0) {
      if (BWD) {
        double a = 0.0;
        for (int w = 0; w < NW; ++w) a += red[w];
        p.gkpart[s * G + c] = -invk * a;
      }
      if (have_out) {
        if (qo.head) go[0] = b0[qo.mis];
        if (qo.tail) go[len - 1] = b0[qo.mis + len - 1];
        if (qo.body) bulk_s2g(go + qo.head, b0 + qo.mis + qo.head, 8u * qo.body);
        bulk_commit();
      }
    }
  }
  if (tid == 0) bulk_wait_read0();
}

// dL/dkappa from the per-(sample, chunk) partials, fixed summation order (no float atomics).
__global__ void k_reduce_gk(const double* part, long long B, int G, int per_sample, double* out) {
  if (per_sample) {
    const long long s = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (s < B) {
      double a = 0.0;
      for (int g = 0; g < G; ++g) a += part[s * G + g];
      out[s] = a;
    }
  } else {
    // single block: thread-strided partial sums, then a fixed tree
    __shared__ double sh[1024];
    double a = 0.0;
    const long long n = B * G;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) a += part[i];
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int d = blockDim.x / 2; d > 0; d >>= 1) {
      if (static_cast<int>(threadIdx.x) < d) sh[threadIdx.x] += sh[threadIdx.x + d];
      __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sh[0];
  }
}

constexpr size_t SMEM_FWD = 512 + sizeof(double) * ((CHP + 2) + (CHP + 4));
constexpr size_t SMEM_BWD = 512 + sizeof(double) * ((CHP + 2) + 2 * (CHP + 4));

struct Plan {
  int G, chg, NG;
  size_t off_cnt, off_summ, off_gk, total;
};

int make_plan(const dfe_mesh* m, long long B, bool bwd, Plan* pl) {
  const int nn = static_cast<int>(m->info.n_nodes);
  pl->G = (nn + CHP - 1) / CHP;
  pl->chg = (nn + pl->G - 1) / pl->G;
  pl->NG = 0;
  size_t off = 0;
  pl->off_cnt = off;
  off += ((static_cast<size_t>(B) * sizeof(int) + 255) / 256) * 256;
  pl->off_summ = off;
  off += static_cast<size_t>(B) * MAX_STAGES * pl->G * 4 * sizeof(double);
  pl->off_gk = off;
  off += static_cast<size_t>(B) * pl->G * sizeof(double);
  pl->total = off + 256;
  (void)bwd;
  return DFE_OK;
}

template <bool BWD>
int launch(const dfe_mesh* m, long long B, P1D p, const Plan& pl, cudaStream_t st) {
  auto kern = k_solve1d<BWD>;
  const size_t smem = BWD ? SMEM_BWD : SMEM_FWD;
  DFE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int per_sm = 0;
  DFE_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, T, smem));
  const long long resident = static_cast<long long>(per_sm) * m->sm_count;
  if (resident < pl.G) {
    dfe::set_error("dfe_solve1d: mesh of %d nodes needs %d co-resident CTAs, device holds %lld — not supported by the fused 1-D kernel",
                   p.nn, pl.G, resident);
    return DFE_ERR_UNSUPPORTED;
  }
  long long NG = resident / pl.G;
  if (NG > B) NG = B;
  p.NG = static_cast<int>(NG);
  k_solve1d<BWD><<<static_cast<unsigned>(NG * pl.G), T, smem, st>>>(p);
  DFE_CUDA_OK(cudaGetLastError());
  return DFE_OK;
}

int common_checks(const dfe_mesh* m, long long B, const void* a, const void* kappa, int kappa_mode, void* ws,
                  size_t ws_bytes, const Plan& pl, const char* who) {
  DFE_REQUIRE(m && a && kappa && ws, "%s: null argument", who);
  DFE_REQUIRE(B >= 1, "%s: B must be >= 1", who);
  if (m->info.device < 0) {
    dfe::set_error("%s: mesh handle is host-only; no CUDA device (this library has no CPU path)", who);
    return DFE_ERR_CUDA;
  }
  if (!m->chain) {
    dfe::set_error("%s: mesh is not a 1-D chain with Dirichlet nodes at its ends; use the general path", who);
    return DFE_ERR_UNSUPPORTED;
  }
  if (kappa_mode != DFE_KAPPA_SCALAR && kappa_mode != DFE_KAPPA_PER_SAMPLE) {
    dfe::set_error("%s: per-element kappa is not implemented in the fused 1-D path yet", who);
    return DFE_ERR_UNSUPPORTED;
  }
  if (ws_bytes < pl.total) {
    dfe::set_error("%s: workspace %zu bytes < required %zu", who, ws_bytes, pl.total);
    return DFE_ERR_WORKSPACE;
  }
  return DFE_OK;
}

int auto_refine(int n_refine, long long nn) {
  if (n_refine < 0) n_refine = nn <= 200000 ? 1 : (nn <= 2000000 ? 2 : 3);
  return n_refine > MAX_STAGES - 1 ? MAX_STAGES - 1 : n_refine;
}

P1D base_params(const dfe_mesh* m, long long B, const Plan& pl, const double* kappa, int kappa_mode, int n_refine,
                void* ws) {
  P1D p{};
  p.nn = static_cast<int>(m->info.n_nodes);
  p.G = pl.G;
  p.chg = pl.chg;
  p.n_refine = auto_refine(n_refine, p.nn);
  p.B = B;
  p.x = m->dev.nodes;
  p.kappa = kappa;
  p.per_sample = kappa_mode == DFE_KAPPA_PER_SAMPLE;
  p.bcL = m->bc_left;
  p.bcR = m->bc_right;
  p.lift_left_first = m->lift_left_first;
  p.gL = m->g_left;
  p.gR = m->g_right;
  unsigned char* w = static_cast<unsigned char*>(ws);
  p.cnt = reinterpret_cast<int*>(w + pl.off_cnt);
  p.summ = reinterpret_cast<double*>(w + pl.off_summ);
  p.gkpart = reinterpret_cast<double*>(w + pl.off_gk);
  return p;
}

}  // namespace

extern "C" size_t dfe_solve1d_workspace_bytes(const dfe_mesh* m, int64_t B) {
  if (!m || B < 1) return 0;
  Plan pl;
  make_plan(m, B, true, &pl);
  return pl.total;
}

extern "C" int dfe_solve1d_fwd(const dfe_mesh* m, int64_t B, const double* f, int64_t ldf, const double* kappa,
                               int kappa_mode, int n_refine, double* u, int64_t ldu, void* ws, size_t ws_bytes,
                               void* stream) {
  Plan pl{};
  if (m) make_plan(m, B, false, &pl);
  int rc = common_checks(m, B, f, kappa, kappa_mode, ws, ws_bytes, pl, "dfe_solve1d_fwd");
  if (rc != DFE_OK) return rc;
  DFE_REQUIRE(u && ldf >= m->info.n_nodes && ldu >= m->info.n_nodes, "dfe_solve1d_fwd: bad u / leading dimension");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int cur = -1;
  DFE_CUDA_OK(cudaGetDevice(&cur));
  if (cur != m->info.device) DFE_CUDA_OK(cudaSetDevice(m->info.device));
  P1D p = base_params(m, B, pl, kappa, kappa_mode, n_refine, ws);
  p.in0 = f;
  p.ld0 = ldf;
  p.out = u;
  p.ldo = ldu;
  DFE_CUDA_OK(cudaMemsetAsync(p.cnt, 0, static_cast<size_t>(B) * sizeof(int), st));
  rc = launch<false>(m, B, p, pl, st);
  if (cur != m->info.device) cudaSetDevice(cur);
  return rc;
}

extern "C" int dfe_solve1d_bwd(const dfe_mesh* m, int64_t B, const double* gbar, int64_t ldg, const double* u,
                               int64_t ldu, const double* kappa, int kappa_mode, int n_refine, double* gf,
                               int64_t ldgf, double* gkappa, void* ws, size_t ws_bytes, void* stream) {
  Plan pl{};
  if (m) make_plan(m, B, true, &pl);
  int rc = common_checks(m, B, gbar, kappa, kappa_mode, ws, ws_bytes, pl, "dfe_solve1d_bwd");
  if (rc != DFE_OK) return rc;
  DFE_REQUIRE(u && gkappa, "dfe_solve1d_bwd: null u / gkappa");
  DFE_REQUIRE(ldg >= m->info.n_nodes && ldu >= m->info.n_nodes && (!gf || ldgf >= m->info.n_nodes),
              "dfe_solve1d_bwd: leading dimension smaller than n_nodes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int cur = -1;
  DFE_CUDA_OK(cudaGetDevice(&cur));
  if (cur != m->info.device) DFE_CUDA_OK(cudaSetDevice(m->info.device));
  P1D p = base_params(m, B, pl, kappa, kappa_mode, n_refine, ws);
  p.in0 = gbar;
  p.ld0 = ldg;
  p.in1 = u;
  p.ld1 = ldu;
  p.out = gf;
  p.ldo = ldgf;
  DFE_CUDA_OK(cudaMemsetAsync(p.cnt, 0, static_cast<size_t>(B) * sizeof(int), st));
  rc = launch<true>(m, B, p, pl, st);
  if (rc == DFE_OK) {
    if (p.per_sample) {
      k_reduce_gk<<<static_cast<unsigned>((B + 255) / 256), 256, 0, st>>>(p.gkpart, B, pl.G, 1, gkappa);
    } else {
      k_reduce_gk<<<1, 1024, 0, st>>>(p.gkpart, B, pl.G, 0, gkappa);
    }
    if (cudaGetLastError() != cudaSuccess) {
      dfe::set_error("dfe_solve1d_bwd: reduce kernel launch failed");
      rc = DFE_ERR_CUDA;
    }
  }
  if (cur != m->info.device) cudaSetDevice(cur);
  return rc;
}
