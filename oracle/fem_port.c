/*
 * fem_port.c — plain-C restatement of the reference hot path, used ONLY as the timed CPU baseline of
 * bench.py ("cpu_baseline", kind "port") and cross-checked against oracle/oracle.py in tests.
 * TEST / MEASUREMENT INFRASTRUCTURE: nothing under difffe_physics_lab_b200/ links or calls this.
 *
 * Follows diffhe/solver.py of the reference:
 *   1-D element loop            solver.py:82-96   (h = xj-xi, k = kappa/h, lumped load h/2*f)
 *   Dirichlet lifting + scatter solver.py:162-181
 *   linear solve                solver.py:174     torch.linalg.solve (dense LU) -> here the Thomas
 *                               algorithm on the same tridiagonal float64 system (what a CPU port of
 *                               the path would use; the dense O(n^3) reference cannot run n = 1e5)
 *   backward                    autograd of the above in closed form
 *   2-D                         Jacobi-PCG on a CSR matrix (assembled by oracle.py), OpenMP SpMV
 * Parallelism: OpenMP over samples (1-D) / rows (2-D).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int port_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* One sample: forward u = K^-1 F, adjoint lam = K^-1 gbar, gk = dL/dkappa, gf = dL/df.
 * work: 4*nn doubles. gbar/gf may be NULL (forward only). */
static void solve1d_one(int nn, const double* x, int bcL, double gL, int bcR, double gR, double kappa,
                        const double* f, const double* gbar, double* u, double* gf, double* gk, double* work) {
  const int ne = nn - 1;
  double* d = work;            /* diagonal            */
  double* F = work + nn;       /* load / rhs          */
  double* cp = work + 2 * nn;  /* Thomas c'           */
  double* k = work + 3 * nn;   /* element stiffness   */
  memset(d, 0, sizeof(double) * nn);
  memset(F, 0, sizeof(double) * nn);
  for (int e = 0; e < ne; ++e) { /* solver.py:82-96 */
    const double h = x[e + 1] - x[e];
    const double ke = kappa / h;
    k[e] = ke;
    d[e] = d[e] + ke;
    d[e + 1] = d[e + 1] + ke;
    F[e] = F[e] + h / 2.0 * f[e];
    F[e + 1] = F[e + 1] + h / 2.0 * f[e + 1];
  }
  const int lo = bcL ? 1 : 0, hi = bcR ? nn - 2 : nn - 1; /* free rows lo..hi */
  if (bcL && lo <= hi) F[1] = F[1] - (-k[0]) * gL;          /* solver.py:166-169 */
  if (bcR && lo <= hi) F[nn - 2] = F[nn - 2] - (-k[ne - 1]) * gR;
  /* Thomas: sub/super diagonal of row i is -k[i-1] / -k[i] */
  for (int pass = 0; pass < 2; ++pass) {
    const double* rhs = pass == 0 ? F : gbar;
    double* out = pass == 0 ? u : cp + 0; /* adjoint result goes to `lam` = reuse F after pass 0 */
    if (pass == 1 && !gbar) break;
    double* sol = pass == 0 ? u : F;
    (void)out;
    double m = d[lo];
    double prev = rhs[lo] / m;
    cp[lo] = (lo < hi) ? -k[lo] / m : 0.0;
    sol[lo] = prev;
    for (int i = lo + 1; i <= hi; ++i) {
      const double a = -k[i - 1];
      m = d[i] - a * cp[i - 1];
      cp[i] = (i < hi) ? -k[i] / m : 0.0;
      prev = (rhs[i] - a * prev) / m;
      sol[i] = prev;
    }
    for (int i = hi - 1; i >= lo; --i) sol[i] = sol[i] - cp[i] * sol[i + 1];
    if (pass == 0) {
      if (bcL) u[0] = gL;       /* solver.py:177-181 */
      if (bcR) u[nn - 1] = gR;
    } else {
      if (bcL) F[0] = 0.0;
      if (bcR) F[nn - 1] = 0.0;
    }
  }
  if (gbar) {
    const double* lam = F;
    double s = 0.0;
    for (int e = 0; e < ne; ++e) s += -(lam[e + 1] - lam[e]) * (u[e + 1] - u[e]) / (x[e + 1] - x[e]);
    *gk = s;
    if (gf) {
      gf[0] = lam[0] * ((x[1] - x[0]) / 2.0);
      for (int i = 1; i < nn - 1; ++i) gf[i] = lam[i] * ((x[i] - x[i - 1]) / 2.0) + lam[i] * ((x[i + 1] - x[i]) / 2.0);
      gf[nn - 1] = lam[nn - 1] * ((x[nn - 1] - x[nn - 2]) / 2.0);
    }
  }
}

/* Batched 1-D forward (+ adjoint when gbar != NULL).  kappa: per-sample [B] if per_sample else [1]. */
int port_solve1d_batch(int nn, const double* x, int bcL, double gL, int bcR, double gR, long B, const double* f,
                       const double* kappa, int per_sample, const double* gbar, double* u, double* gf, double* gk,
                       int nthreads) {
  if (nn < 2 || (!bcL && !bcR)) return 1;
  int bad = 0;
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : port_max_threads())
  {
    double* work = (double*)malloc(sizeof(double) * 4 * (size_t)nn);
    if (!work) {
#pragma omp atomic write
      bad = 1;
    } else {
#pragma omp for schedule(dynamic, 1)
      for (long b = 0; b < B; ++b) {
        double g = 0.0;
        solve1d_one(nn, x, bcL, gL, bcR, gR, kappa[per_sample ? b : 0], f + b * (size_t)nn,
                    gbar ? gbar + b * (size_t)nn : NULL, u + b * (size_t)nn, gf ? gf + b * (size_t)nn : NULL, &g, work);
        if (gk) gk[b] = g;
      }
      free(work);
    }
  }
  return bad;
}

/* Jacobi-PCG on CSR (float64), recursive-residual stop; returns iterations. */
long port_pcg_csr(long n, const long* rowptr, const long* col, const double* val, const double* b, double* x,
                  double tol, long maxit, double* relres, int nthreads) {
  const int nt = nthreads > 0 ? nthreads : port_max_threads();
  double* r = (double*)malloc(sizeof(double) * 4 * (size_t)n);
  double *z = r + n, *p = r + 2 * n, *q = r + 3 * n;
  double* dinv = (double*)malloc(sizeof(double) * (size_t)n);
  double bb = 0.0, rz = 0.0;
#pragma omp parallel for num_threads(nt) reduction(+ : bb, rz)
  for (long i = 0; i < n; ++i) {
    double dg = 1.0;
    for (long k = rowptr[i]; k < rowptr[i + 1]; ++k)
      if (col[k] == i) dg = val[k];
    dinv[i] = 1.0 / dg;
    x[i] = 0.0;
    r[i] = b[i];
    z[i] = dinv[i] * r[i];
    p[i] = z[i];
    bb += b[i] * b[i];
    rz += r[i] * z[i];
  }
  long it = 0;
  double rr = bb;
  if (bb > 0.0) {
    const double bn = sqrt(bb);
    while (it < maxit) {
      double pq = 0.0;
#pragma omp parallel for num_threads(nt) reduction(+ : pq)
      for (long i = 0; i < n; ++i) {
        double s = 0.0;
        for (long k = rowptr[i]; k < rowptr[i + 1]; ++k) s += val[k] * p[col[k]];
        q[i] = s;
        pq += p[i] * s;
      }
      const double alpha = rz / pq;
      double rzn = 0.0;
      rr = 0.0;
#pragma omp parallel for num_threads(nt) reduction(+ : rzn, rr)
      for (long i = 0; i < n; ++i) {
        x[i] += alpha * p[i];
        r[i] -= alpha * q[i];
        z[i] = dinv[i] * r[i];
        rzn += r[i] * z[i];
        rr += r[i] * r[i];
      }
      ++it;
      if (sqrt(rr) <= tol * bn) break;
      const double beta = rzn / rz;
      rz = rzn;
#pragma omp parallel for num_threads(nt)
      for (long i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
    }
    if (relres) *relres = sqrt(rr) / bn;
  } else if (relres) {
    *relres = 0.0;
  }
  free(r);
  free(dinv);
  return it;
}

/* Many right-hand sides on ONE matrix (config 5b): OpenMP over the samples, every sample solved by the single-thread
   Jacobi-PCG above.  rhs, x are (B, n) row-major; returns the largest iteration count. */
long port_pcg_csr_batch(long n, const long* rowptr, const long* col, const double* val, long B, const double* rhs,
                        double* x, double tol, long maxit, int nthreads) {
  const int nt = nthreads > 0 ? nthreads : port_max_threads();
  long worst = 0;
#pragma omp parallel for num_threads(nt) schedule(dynamic, 1) reduction(max : worst)
  for (long b = 0; b < B; ++b) {
    double rel;
    const long it = port_pcg_csr(n, rowptr, col, val, rhs + b * (size_t)n, x + b * (size_t)n, tol, maxit, &rel, 1);
    if (it > worst) worst = it;
  }
  return worst;
}
