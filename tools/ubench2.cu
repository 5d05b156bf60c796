// Intrinsic cost of the three per-thread phases of the pipelined 1-D kernel (no synchronisation, no I/O):
// cycles per phase call for a given number of warps per CTA and CTAs per SM.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ double kdiv(double kaph, double h, double y) {
  const double q0 = kaph * y;
  return fma(fma(-h, q0, kaph), y, q0);
}
__device__ __forceinline__ double two_sum_err(double a, double b) {
  const double d = __dadd_rn(a, b);
  const double bb = __dsub_rn(d, a);
  return __dadd_rn(__dsub_rn(a, __dsub_rn(d, bb)), __dsub_rn(b, bb));
}
__device__ __forceinline__ void warp_scan2(double& is, double& im, double& es, double& em, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const double os = __shfl_up_sync(0xffffffffu, is, d), om = __shfl_up_sync(0xffffffffu, im, d);
    if (lane >= d) { is += os; im += om; }
  }
  es = __shfl_up_sync(0xffffffffu, is, 1);
  em = __shfl_up_sync(0xffffffffu, im, 1);
  if (lane == 0) { es = 0.0; em = 0.0; }
}
template <int R>
__device__ __forceinline__ void phase_a(double* buf, const double (&hs)[R + 1], double& S, double& Wc) {
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const double in = buf[j];
    double F = fma(in, hs[j + 1], in * hs[j]);
    buf[j] = Wc;
    S += F;
    Wc = fma(hs[j + 1], S, Wc);
  }
}
template <int R>
__device__ __forceinline__ void phase_b(double* buf, const double (&hs)[R + 1], const double (&rh)[R + 1],
                                        double X0, double c0, double kaph, double a0, double b0, double& S1, double& W1) {
  double kp = kdiv(kaph, hs[0], rh[0]);
  double t = fma(b0, X0, a0);
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const double Wt = buf[j];
    const double x0 = fma(-c0, Wt, t);
    t = fma(b0, hs[j + 1], t);
    const double ki = kdiv(kaph, hs[j + 1], rh[j + 1]);
    const double v1 = __dmul_rn(two_sum_err(kp, ki), x0);
    kp = ki;
    buf[j] = fma(-c0, W1, x0);
    S1 += v1;
    W1 = fma(hs[j + 1], S1, W1);
  }
}
template <int R>
__device__ __forceinline__ void phase_c(double* buf, const double (&hs)[R + 1], double X0, double a1, double b1) {
  double t = fma(b1, X0, a1);
#pragma unroll
  for (int j = 0; j < R; ++j) {
    buf[j] = buf[j] + t;
    t = fma(b1, hs[j + 1], t);
  }
}

template <int R, int WHICH>
__global__ void __launch_bounds__(512, 1) k_phase(const double* hsg, double* out, long long* cyc, int iters, int with_scan) {
  extern __shared__ double sm[];
  const int tid = threadIdx.x, lane = tid & 31;
  double hs[R + 1], rh[R + 1];
#pragma unroll
  for (int j = 0; j <= R; ++j) { hs[j] = hsg[tid * R + j]; rh[j] = 1.0 / hs[j]; }
  double* buf = sm + tid * R;
  for (int j = 0; j < R; ++j) buf[j] = 0.001 * (tid + j);
  const double X0 = tid * 0.01, Xe = X0 + 0.11;
  __syncthreads();
  double acc = 0.0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    double S = 0.0, Wv = 0.0;
    if (WHICH == 0) phase_a<R>(buf, hs, S, Wv);
    if (WHICH == 1) phase_b<R>(buf, hs, rh, X0, 1.0 + 1e-9 * it, 0.7, 0.1, 0.2, S, Wv);
    if (WHICH == 2) phase_c<R>(buf, hs, X0, 0.1 + 1e-9 * it, 0.2);
    if (WHICH == 3) {   // all three back to back
      phase_a<R>(buf, hs, S, Wv);
      double S1 = 0, W1 = 0;
      phase_b<R>(buf, hs, rh, X0, 1.0 + 1e-9 * it, 0.7, 0.1, 0.2, S1, W1);
      phase_c<R>(buf, hs, X0, 0.1 + 1e-9 * it, 0.2);
      S += S1; Wv += W1;
    }
    if (with_scan && WHICH != 2) {
      double is = S, im = fma(-S, Xe, Wv), es, em;
      warp_scan2(is, im, es, em, lane);
      acc += es + em;
    } else {
      acc += S + Wv;
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + tid] = acc + buf[0];
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int R, int WHICH>
void run(const char* name, int threads, int ctas_per_sm, int with_scan, const double* hsg, double* out, long long* cyc) {
  const int iters = 400;
  auto k = k_phase<R, WHICH>;
  const size_t smem = (size_t)threads * R * 8;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  k<<<148 * ctas_per_sm, threads, smem>>>(hsg, out, cyc, iters, with_scan);
  CK(cudaDeviceSynchronize());
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-8s R=%2d threads/CTA=%3d CTAs/SM=%d scan=%d: %7.0f cycles per call  (%.2f cycles per node per CTA)\n", name, R, threads, ctas_per_sm,
         with_scan, (double)c / iters, (double)c / iters / (threads * R) );
}

int main() {
  double* hsg; double* out; long long* cyc;
  const int n = 1 << 16;
  double* h = (double*)malloc(n * 8);
  for (int i = 0; i < n; ++i) h[i] = 5e-6 * (1.0 + 1e-12 * (i % 7));
  CK(cudaMalloc(&hsg, n * 8)); CK(cudaMemcpy(hsg, h, n * 8, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&out, 1 << 24)); CK(cudaMalloc(&cyc, 1 << 16));
  for (int cps = 1; cps <= 2; ++cps) {
    for (int th : {32, 128, 160, 256}) {
      run<11, 0>("A", th, cps, 1, hsg, out, cyc);
      run<11, 1>("B", th, cps, 1, hsg, out, cyc);
      run<11, 2>("C", th, cps, 0, hsg, out, cyc);
      run<11, 3>("A+B+C", th, cps, 1, hsg, out, cyc);
    }
  }
  run<11, 0>("A", 160, 2, 0, hsg, out, cyc);
  run<11, 1>("B", 160, 2, 0, hsg, out, cyc);
  run<7, 3>("A+B+C", 256, 2, 1, hsg, out, cyc);
  run<7, 3>("A+B+C", 512, 1, 1, hsg, out, cyc);
  run<5, 3>("A+B+C", 512, 1, 1, hsg, out, cyc);
  run<14, 3>("A+B+C", 128, 2, 1, hsg, out, cyc);
  return 0;
}
