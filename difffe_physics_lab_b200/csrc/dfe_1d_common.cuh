// Helpers shared by the fused 1-D kernels: the (S, X, W) prefix triple and its combine, warp-shuffle scans,
// and the PTX wrappers for mbarrier + 1-D TMA bulk copies.  Included inside an anonymous namespace.
#pragma once

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
// several barriers, one fence (the fence is not free: see k_band_solve_mma)
__device__ __forceinline__ void mbar_init_raw(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
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
__device__ __forceinline__ int ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// release-increment: orders this thread's earlier global stores before the counter update
__device__ __forceinline__ void red_release_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
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

