// General (CSR) path: P1 assembly, Dirichlet elimination, scatter/gather and the element-gradient
// kernels for 2-D triangle meshes and for 1-D meshes that are not chains.
//
//   k_assemble   diffhe/solver.py:82-96 (1-D) and :112-145 (2-D): one thread per matrix row walks
//                the elements around its node in ascending element id — the reference's own
//                accumulation order — so K (structural CSR) and F are bit-identical to the dense
//                reference without float atomics.  All arithmetic uses explicit round-to-nearest
//                intrinsics in the reference's operation order (no FMA contraction).
//   k_eliminate  solver.py:162-171: F_free = F[free] - sum_d K[free,d] g_d in dict order,
//                1/diag(K_free), CSR values of K_free;  k_fill_sell copies K_free into SELL-32.
//   k_scatter / k_gather   solver.py:177-181 and the restriction of gbar to free rows.
//   k_grad_elem / k_grad_f the closed-form backward of the assembly.
#include "dfe_internal.h"

namespace {

using dfe::MeshDev;
constexpr double AREA_EPS = 1e-15;  // solver.py:120

struct Elem2D {
  double area, b[3], c[3];
};

// solver.py:114-134 in the reference's operation order
__device__ __forceinline__ Elem2D elem2d(const MeshDev& M, int e, int n[3]) {
  n[0] = M.elems[3 * e + 0];
  n[1] = M.elems[3 * e + 1];
  n[2] = M.elems[3 * e + 2];
  const double xi = M.nodes[2 * n[0]], yi = M.nodes[2 * n[0] + 1];
  const double xj = M.nodes[2 * n[1]], yj = M.nodes[2 * n[1] + 1];
  const double xk = M.nodes[2 * n[2]], yk = M.nodes[2 * n[2] + 1];
  Elem2D E;
  const double t1 = __dmul_rn(__dsub_rn(xj, xi), __dsub_rn(yk, yi));
  const double t2 = __dmul_rn(__dsub_rn(xk, xi), __dsub_rn(yj, yi));
  E.area = __dmul_rn(0.5, fabs(__dsub_rn(t1, t2)));
  E.b[0] = __dsub_rn(yj, yk);
  E.b[1] = __dsub_rn(yk, yi);
  E.b[2] = __dsub_rn(yi, yj);
  E.c[0] = __dsub_rn(xk, xj);
  E.c[1] = __dsub_rn(xi, xk);
  E.c[2] = __dsub_rn(xj, xi);
  return E;
}

__global__ void k_assemble(const MeshDev M, const double* __restrict__ kappa, int per_elem,
                           const double* __restrict__ f, double* __restrict__ vals, double* __restrict__ F) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= M.n_nodes) return;
  for (int k = M.rowptr[p]; k < M.rowptr[p + 1]; ++k) vals[k] = 0.0;
  double Fp = 0.0;
  for (int a = M.adj_ptr[p]; a < M.adj_ptr[p + 1]; ++a) {
    const int e = M.adj_elem[a];
    const int loc = M.adj_loc[a];
    const double kap = kappa[per_elem ? e : 0];
    if (M.dim == 1) {
      const int i = M.elems[2 * e], j = M.elems[2 * e + 1];
      const double h = __dsub_rn(M.nodes[j], M.nodes[i]);   // solver.py:84-85
      const double ke = __ddiv_rn(kap, h);                  // :88
      // row `loc` of k_e [[1,-1],[-1,1]]  (:89-92: K[i,i]+k, K[i,j]-k, K[j,i]-k, K[j,j]+k)
      const int s0 = M.adj_slot[2 * a], s1 = M.adj_slot[2 * a + 1];
      if (loc == 0) {
        vals[s0] = __dadd_rn(vals[s0], ke);
        vals[s1] = __dsub_rn(vals[s1], ke);
      } else {
        vals[s0] = __dsub_rn(vals[s0], ke);
        vals[s1] = __dadd_rn(vals[s1], ke);
      }
      Fp = __dadd_rn(Fp, __dmul_rn(__ddiv_rn(h, 2.0), f[p]));  // :95-96
    } else {
      int n[3];
      const Elem2D E = elem2d(M, e, n);
      if (E.area < AREA_EPS) continue;  // :120-121
      const double den = __dmul_rn(4.0, E.area);
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        // k_pq = kappa*(b_p b_q + c_p c_q)/(4 area)   (:139)
        const double num = __dmul_rn(kap, __dadd_rn(__dmul_rn(E.b[loc], E.b[q]), __dmul_rn(E.c[loc], E.c[q])));
        const int sl = M.adj_slot[3 * a + q];
        vals[sl] = __dadd_rn(vals[sl], __ddiv_rn(num, den));
      }
      // F_p += area/3 * (f_i+f_j+f_k)/3   (:143-145)
      const double fc = __ddiv_rn(__dadd_rn(__dadd_rn(f[n[0]], f[n[1]]), f[n[2]]), 3.0);
      Fp = __dadd_rn(Fp, __dmul_rn(__ddiv_rn(E.area, 3.0), fc));
    }
  }
  F[p] = Fp;
}

__global__ void k_eliminate(const MeshDev M, const double* __restrict__ vals, const double* __restrict__ F,
                            double* __restrict__ vals_free, double* __restrict__ F_free,
                            double* __restrict__ dinv) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= M.n_free) return;
  double Fr = F[M.free_nodes[r]];
  for (int t = M.lift_ptr[r]; t < M.lift_ptr[r + 1]; ++t)
    Fr = __dsub_rn(Fr, __dmul_rn(vals[M.lift_src[t]], M.lift_g[t]));  // solver.py:169
  F_free[r] = Fr;
  const int ds = M.diag_src[r];
  dinv[r] = 1.0 / (ds >= 0 ? vals[ds] : 0.0);
  if (vals_free)
    for (int k = M.rowptr_f[r]; k < M.rowptr_f[r + 1]; ++k) vals_free[k] = vals[M.src_f[k]];
}

__global__ void k_fill_sell(const MeshDev M, const double* __restrict__ vals, double* __restrict__ sell_vals) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M.sell_nnz; i += gridDim.x * blockDim.x) {
    const int s = M.sell_src[i];
    sell_vals[i] = s >= 0 ? vals[s] : 0.0;
  }
}

__global__ void k_scatter(const MeshDev M, const double* __restrict__ x, int zero_bc, double* __restrict__ u) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M.n_free) u[M.free_nodes[i]] = x[i];
  if (i < M.n_dir) u[M.dir_idx[i]] = zero_bc ? 0.0 : M.dir_val[i];
}

__global__ void k_gather(const MeshDev M, const double* __restrict__ v, double* __restrict__ vf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M.n_free) vf[i] = v[M.free_nodes[i]];
}

// dL/dkappa_e = -lam_e^T K_e^0 u_e  (K_e^0: element matrix at kappa = 1)
__global__ void k_grad_elem(const MeshDev M, const double* __restrict__ lam, const double* __restrict__ u,
                            double* __restrict__ gk) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= M.n_el) return;
  double g = 0.0;
  if (M.dim == 1) {
    const int i = M.elems[2 * e], j = M.elems[2 * e + 1];
    const double h = M.nodes[j] - M.nodes[i];
    g = -(lam[j] - lam[i]) * (u[j] - u[i]) / h;
  } else {
    int n[3];
    const Elem2D E = elem2d(M, e, n);
    if (!(E.area < AREA_EPS)) {
      double bl = 0, bu = 0, cl = 0, cu = 0;
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const double l = lam[n[q]], uu = u[n[q]];
        bl = fma(E.b[q], l, bl);
        bu = fma(E.b[q], uu, bu);
        cl = fma(E.c[q], l, cl);
        cu = fma(E.c[q], uu, cu);
      }
      g = -(bl * bu + cl * cu) / (4.0 * E.area);
    }
  }
  gk[e] = g;
}

// dL/df: 1-D  gf_p = sum_{e∋p} (h_e/2) lam_p ;  2-D  gf_q = sum_{e∋q} (area_e/9) sum_{p∈e} lam_p
__global__ void k_grad_f(const MeshDev M, const double* __restrict__ lam, double* __restrict__ gf) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= M.n_nodes) return;
  double g = 0.0;
  for (int a = M.adj_ptr[p]; a < M.adj_ptr[p + 1]; ++a) {
    const int e = M.adj_elem[a];
    if (M.dim == 1) {
      const int i = M.elems[2 * e], j = M.elems[2 * e + 1];
      g = fma((M.nodes[j] - M.nodes[i]) * 0.5, lam[p], g);
    } else {
      int n[3];
      const Elem2D E = elem2d(M, e, n);
      if (E.area < AREA_EPS) continue;
      g = fma(E.area / 9.0, (lam[n[0]] + lam[n[1]]) + lam[n[2]], g);
    }
  }
  gf[p] = g;
}

// deterministic sum of n doubles by ONE block (fixed order for a fixed n)
__global__ void k_sum(const double* __restrict__ in, long long n, double* __restrict__ out) {
  __shared__ double sh[1024];
  double a = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) a += in[i];
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int d = blockDim.x / 2; d > 0; d >>= 1) {
    if (static_cast<int>(threadIdx.x) < d) sh[threadIdx.x] += sh[threadIdx.x + d];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0];
}

int enter(const dfe_mesh* m, const char* who, int* prev) {
  if (!m) {
    dfe::set_error("%s: mesh is null", who);
    return DFE_ERR_INVALID;
  }
  if (m->info.device < 0) {
    dfe::set_error("%s: mesh handle is host-only; no CUDA device (this library has no CPU path)", who);
    return DFE_ERR_CUDA;
  }
  DFE_CUDA_OK(cudaGetDevice(prev));
  if (*prev != m->info.device) DFE_CUDA_OK(cudaSetDevice(m->info.device));
  return DFE_OK;
}
void leave(const dfe_mesh* m, int prev) {
  if (prev != m->info.device) cudaSetDevice(prev);
}
int check_launch(const char* who) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    dfe::set_error("%s: kernel launch failed: %s", who, cudaGetErrorString(e));
    return DFE_ERR_CUDA;
  }
  return DFE_OK;
}
inline unsigned blocks(long long n, int t) { return static_cast<unsigned>((n + t - 1) / t > 0 ? (n + t - 1) / t : 1); }

}  // namespace

extern "C" int dfe_assemble(const dfe_mesh* m, const double* kappa, int kappa_mode, const double* f,
                            double* vals_full, double* F, void* stream) {
  int prev;
  int rc = enter(m, "dfe_assemble", &prev);
  if (rc) return rc;
  if (!kappa || !f || !vals_full || !F) {
    dfe::set_error("dfe_assemble: null argument");
    rc = DFE_ERR_INVALID;
  } else if (kappa_mode != DFE_KAPPA_SCALAR && kappa_mode != DFE_KAPPA_PER_ELEMENT) {
    dfe::set_error("dfe_assemble: kappa_mode must be SCALAR or PER_ELEMENT");
    rc = DFE_ERR_INVALID;
  } else {
    k_assemble<<<blocks(m->dev.n_nodes, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        m->dev, kappa, kappa_mode == DFE_KAPPA_PER_ELEMENT, f, vals_full, F);
    rc = check_launch("dfe_assemble");
  }
  leave(m, prev);
  return rc;
}

extern "C" int dfe_eliminate(const dfe_mesh* m, const double* vals_full, const double* F, double* vals_free,
                             double* sell_vals, double* F_free, double* dinv, void* stream) {
  int prev;
  int rc = enter(m, "dfe_eliminate", &prev);
  if (rc) return rc;
  if (!vals_full || !F || !F_free || !dinv) {
    dfe::set_error("dfe_eliminate: null argument");
    rc = DFE_ERR_INVALID;
  } else if (m->dev.n_free > 0) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    k_eliminate<<<blocks(m->dev.n_free, 128), 128, 0, st>>>(m->dev, vals_full, F, vals_free, F_free, dinv);
    if (sell_vals)
      k_fill_sell<<<blocks(m->dev.sell_nnz, 256) < 4096u ? blocks(m->dev.sell_nnz, 256) : 4096u, 256, 0, st>>>(
          m->dev, vals_full, sell_vals);
    rc = check_launch("dfe_eliminate");
  }
  leave(m, prev);
  return rc;
}

extern "C" int dfe_scatter(const dfe_mesh* m, const double* x_free, int zero_bc, double* u, void* stream) {
  int prev;
  int rc = enter(m, "dfe_scatter", &prev);
  if (rc) return rc;
  if (!u || (!x_free && m->dev.n_free > 0)) {
    dfe::set_error("dfe_scatter: null argument");
    rc = DFE_ERR_INVALID;
  } else {
    const int n = m->dev.n_free > m->dev.n_dir ? m->dev.n_free : m->dev.n_dir;
    k_scatter<<<blocks(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(m->dev, x_free, zero_bc, u);
    rc = check_launch("dfe_scatter");
  }
  leave(m, prev);
  return rc;
}

extern "C" int dfe_gather_free(const dfe_mesh* m, const double* v_full, double* v_free, void* stream) {
  int prev;
  int rc = enter(m, "dfe_gather_free", &prev);
  if (rc) return rc;
  if (!v_full || (!v_free && m->dev.n_free > 0)) {
    dfe::set_error("dfe_gather_free: null argument");
    rc = DFE_ERR_INVALID;
  } else if (m->dev.n_free > 0) {
    k_gather<<<blocks(m->dev.n_free, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(m->dev, v_full, v_free);
    rc = check_launch("dfe_gather_free");
  }
  leave(m, prev);
  return rc;
}

extern "C" size_t dfe_grad_workspace_bytes(const dfe_mesh* m) {
  return m ? (static_cast<size_t>(m->info.n_elements) + 32) * sizeof(double) : 0;
}

extern "C" int dfe_grad(const dfe_mesh* m, const double* lam_full, const double* u, const double* kappa,
                        int kappa_mode, double* gkappa, double* gf, void* ws, size_t ws_bytes, void* stream) {
  (void)kappa;  // K is linear in kappa: K_e = kappa_e K_e^0, the gradient does not need its value
  int prev;
  int rc = enter(m, "dfe_grad", &prev);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!lam_full || !u || !gkappa) {
    dfe::set_error("dfe_grad: null argument");
    rc = DFE_ERR_INVALID;
  } else if (kappa_mode == DFE_KAPPA_PER_ELEMENT) {
    if (m->dev.n_el > 0) k_grad_elem<<<blocks(m->dev.n_el, 256), 256, 0, st>>>(m->dev, lam_full, u, gkappa);
  } else if (kappa_mode == DFE_KAPPA_SCALAR) {
    if (!ws || ws_bytes < dfe_grad_workspace_bytes(m)) {
      dfe::set_error("dfe_grad: workspace too small");
      rc = DFE_ERR_WORKSPACE;
    } else {
      double* tmp = static_cast<double*>(ws);
      if (m->dev.n_el > 0) k_grad_elem<<<blocks(m->dev.n_el, 256), 256, 0, st>>>(m->dev, lam_full, u, tmp);
      k_sum<<<1, 1024, 0, st>>>(tmp, m->dev.n_el, gkappa);
    }
  } else {
    dfe::set_error("dfe_grad: kappa_mode must be SCALAR or PER_ELEMENT");
    rc = DFE_ERR_INVALID;
  }
  if (rc == DFE_OK && gf) k_grad_f<<<blocks(m->dev.n_nodes, 128), 128, 0, st>>>(m->dev, lam_full, gf);
  if (rc == DFE_OK) rc = check_launch("dfe_grad");
  leave(m, prev);
  return rc;
}
