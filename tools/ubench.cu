// Micro-benchmarks that size the 1-D pipelined kernel (sm_100a): FP64 dependent-issue latency and pipe
// throughput vs warps per SM, shuffle / shared-memory / named-barrier latency, and the publish -> poll
// round trip between two SMs through L2.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_bin/ubench tools/ubench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// ---- dependent chain of N ops, ILP independent chains per thread
template <int ILP, int KIND>
__global__ void k_chain(double* out, long long* cyc, double a, double b, int iters) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = a + i + threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        if (KIND == 0) x[i] = __dadd_rn(x[i], b);
        else if (KIND == 1) x[i] = fma(x[i], b, a);
        else x[i] = __dmul_rn(x[i], b);
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_shfl(double* out, long long* cyc, int iters) {
  double x = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) x = __shfl_up_sync(0xffffffffu, x, 1) ;
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void k_shfl_add(double* out, long long* cyc, int iters) {
  double x = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) x = __dadd_rn(x, __shfl_up_sync(0xffffffffu, x, 1));
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void k_lds(double* out, long long* cyc, int iters) {
  __shared__ int nxt[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) nxt[i] = (i + 33) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) p = nxt[p];
  }
  long long t1 = clock64();
  out[threadIdx.x] = p;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// named barrier ping: all warps of the block sync `iters` times
__global__ void k_bar(long long* cyc, int iters) {
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ping-pong between block 0 and block 1 (different SMs) through a global word: round trip latency
__global__ void k_pingpong(unsigned long long* flag, long long* cyc, int iters) {
  if (threadIdx.x != 0) return;
  long long t0 = clock64();
  for (int it = 1; it <= iters; ++it) {
    if (blockIdx.x == 0) {
      asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(flag), "l"((unsigned long long)(2 * it - 1)) : "memory");
      unsigned long long v;
      do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(flag + 16) : "memory"); } while (v < (unsigned long long)(2 * it));
    } else if (blockIdx.x == gridDim.x - 1) {
      unsigned long long v;
      do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory"); } while (v < (unsigned long long)(2 * it - 1));
      asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(flag + 16), "l"((unsigned long long)(2 * it)) : "memory");
    }
  }
  long long t1 = clock64();
  if (blockIdx.x == 0) cyc[0] = t1 - t0;
}

// L2 load latency (pointer chase over a 32 MB buffer, stride 4 KB + 128)
__global__ void k_l2chase(const int* nxt, double* out, long long* cyc, int iters) {
  int p = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(p) : "l"(nxt + p) : "memory");
  }
  long long t1 = clock64();
  out[0] = p;
  cyc[0] = t1 - t0;
}

template <int ILP, int KIND>
void run_chain(const char* name, int blocks, int threads, double* d_out, long long* d_cyc) {
  const int iters = 2000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_chain<ILP, KIND><<<blocks, threads>>>(d_out, d_cyc, 1.0, 1.0000001, 10);
  cudaEventRecord(e0);
  k_chain<ILP, KIND><<<blocks, threads>>>(d_out, d_cyc, 1.0, 1.0000001, iters);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c; cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
  const double ops = 8.0 * ILP * iters;
  printf("%-6s ILP=%d blocks=%4d threads=%4d: %.2f cycles/op/thread (dependent latency if ILP=1), warp-instr/clk/SM = %.3f, lane-ops/clk/SM = %.1f\n",
         name, ILP, blocks, threads, (double)c / (8.0 * iters), ops * (threads / 32.0) * (blocks > 148 ? blocks / 148.0 : 1.0) / c,
         ops * threads * (blocks > 148 ? blocks / 148.0 : 1.0) / c);
}

int main() {
  double* d_out; long long* d_cyc;
  CK(cudaMalloc(&d_out, 1 << 24));
  CK(cudaMalloc(&d_cyc, 1 << 16));
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  printf("device %s, %d SMs, clock %d kHz\n", pr.name, pr.multiProcessorCount, pr.clockRate);
  // latency: one warp
  run_chain<1, 0>("DADD", 1, 32, d_out, d_cyc);
  run_chain<1, 1>("DFMA", 1, 32, d_out, d_cyc);
  run_chain<1, 2>("DMUL", 1, 32, d_out, d_cyc);
  run_chain<2, 1>("DFMA", 1, 32, d_out, d_cyc);
  run_chain<4, 1>("DFMA", 1, 32, d_out, d_cyc);
  run_chain<8, 1>("DFMA", 1, 32, d_out, d_cyc);
  // one warp per SMSP
  run_chain<1, 1>("DFMA", 148, 128, d_out, d_cyc);
  run_chain<4, 1>("DFMA", 148, 128, d_out, d_cyc);
  run_chain<8, 1>("DFMA", 148, 128, d_out, d_cyc);
  // more warps per SM
  run_chain<1, 1>("DFMA", 148, 256, d_out, d_cyc);
  run_chain<1, 1>("DFMA", 148, 512, d_out, d_cyc);
  run_chain<1, 1>("DFMA", 148, 1024, d_out, d_cyc);
  run_chain<2, 1>("DFMA", 148, 512, d_out, d_cyc);
  run_chain<4, 1>("DFMA", 148, 512, d_out, d_cyc);
  run_chain<2, 0>("DADD", 148, 512, d_out, d_cyc);
  run_chain<4, 0>("DADD", 148, 1024, d_out, d_cyc);
  {
    const int iters = 2000;
    long long c;
    k_shfl<<<1, 32>>>(d_out, d_cyc, iters); CK(cudaDeviceSynchronize());
    cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
    printf("SHFL.UP f64 (2 x 32-bit) dependent: %.1f cycles\n", (double)c / (8.0 * iters));
    k_shfl_add<<<1, 32>>>(d_out, d_cyc, iters); CK(cudaDeviceSynchronize());
    cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
    printf("SHFL.UP f64 + DADD dependent: %.1f cycles\n", (double)c / (8.0 * iters));
    k_lds<<<1, 32>>>(d_out, d_cyc, iters); CK(cudaDeviceSynchronize());
    cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
    printf("LDS dependent (pointer chase): %.1f cycles\n", (double)c / (8.0 * iters));
    for (int th : {64, 256, 512}) {
      k_bar<<<1, th>>>(d_cyc, 4000); CK(cudaDeviceSynchronize());
      cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
      printf("__syncthreads, %d threads: %.1f cycles\n", th, (double)c / 4000.0);
    }
    unsigned long long* flag; CK(cudaMalloc(&flag, 4096)); CK(cudaMemset(flag, 0, 4096));
    for (int nb : {2, 75, 148}) {
      CK(cudaMemset(flag, 0, 4096));
      k_pingpong<<<nb, 32>>>(flag, d_cyc, 2000); CK(cudaDeviceSynchronize());
      cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
      printf("global ping-pong block 0 <-> block %d: %.0f cycles per round trip (two publish->observe hops)\n", nb - 1, (double)c / 2000.0);
    }
    // L2 chase
    const int n = 8 << 20;   // 32 MB of ints
    int* h = (int*)malloc(n * sizeof(int));
    const int stride = 1024 + 32;
    for (int i = 0; i < n; ++i) h[i] = 0;
    int p = 0; for (int i = 0; i < 7000; ++i) { int q = (p + stride) % n; h[p] = q; p = q; }
    int* dn; CK(cudaMalloc(&dn, n * sizeof(int))); CK(cudaMemcpy(dn, h, n * sizeof(int), cudaMemcpyHostToDevice));
    k_l2chase<<<1, 1>>>(dn, d_out, d_cyc, 7000); CK(cudaDeviceSynchronize());
    k_l2chase<<<1, 1>>>(dn, d_out, d_cyc, 7000); CK(cudaDeviceSynchronize());
    cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
    printf("L2-resident global load, dependent: %.0f cycles\n", (double)c / 7000.0);
  }
  return 0;
}
