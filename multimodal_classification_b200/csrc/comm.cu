// Gradient exchange over NVLink 5 / NVSwitch without NCCL: an in-place mean all-reduce of a bf16 range of a SYMMETRIC buffer
// (the same virtual range mapped on every rank, plus one multicast mapping of all of them).
//
//   barrier  (1 CTA)   every rank's producers of the range have finished (flags in peer memory, CAS hand-shake)
//   reduce   (G CTAs)  rank r owns the r-th slice of the range: multimem.ld_reduce pulls the slice of ALL ranks through the
//                      switch, which adds them in fp32 (NVLS); x 1/N; multimem.st broadcasts the result into every rank's
//                      buffer -- in bf16 in place, or WIDENED TO FP32 straight into every rank's fp32 gradient buffer (a second
//                      symmetric buffer), which removes the separate bf16 -> fp32 pass (1.5 GB of HBM traffic per step) at
//                      the price of fp32 bytes on the otherwise idle links.
//   barrier  (1 CTA)   every rank's broadcasts have landed
//
// Three small launches on one stream (captured into the backward graph by the caller); the reduce CTAs are 256 threads with
// <= 32 registers so that they co-reside with the GEMM CTAs of the backward pass instead of displacing them.  Where the
// multicast mapping is not available the same kernel runs over the peer pointers (P2P loads of the slice from every rank, P2P
// stores to every rank).  Memory and address exchange: torch.distributed._symmetric_memory (plumbing); everything on the
// wire is issued by these kernels.  The reference has no multi-GPU code.
#include "common.cuh"
#include "../../include/vilbert_b200.h"

namespace vb {

__device__ __forceinline__ uint32_t cas_release_sys(uint32_t* addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.global.release.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ uint32_t cas_acquire_sys(uint32_t* addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}

// One thread per peer: raise my flag in the peer's pad (0 -> 1), then consume the peer's flag in mine (1 -> 0).  Slots are
// left at 0, so the same slots serve the next barrier; a rank that runs ahead spins in its put until the peer has consumed
// the previous one.  Bounded by wall clock (10 s) so that a lost rank traps instead of hanging the box.
__global__ void __launch_bounds__(32) rank_barrier_kernel(uint32_t* const* flag_ptrs, int rank, int world, int slot) {
  const int peer = threadIdx.x;
  if (peer >= world) return;
  __threadfence_system();
  uint32_t* theirs = flag_ptrs[peer] + slot * world + rank;
  uint32_t* mine = flag_ptrs[rank] + slot * world + peer;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  uint32_t spins = 0;
  while (cas_release_sys(theirs, 0u, 1u) != 0u) {
    if ((++spins & 255u) == 0) { unsigned long long now; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now)); if (now - t0 > 10000000000ull) __trap(); }
  }
  while (cas_acquire_sys(mine, 1u, 0u) != 1u) {
    if ((++spins & 255u) == 0) { unsigned long long now; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now)); if (now - t0 > 10000000000ull) __trap(); }
  }
  __threadfence_system();
}

__device__ __forceinline__ uint4 scale_bf16x8(uint4 v, float s) {
  uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = unpack_bf16x2(w[i]);
    w[i] = pack_bf16x2(f.x * s, f.y * s);
  }
  return v;
}

// units = 16-byte units of this rank's slice; `base` = multicast address of the slice (MULTICAST) or its offset in bytes
// from every peer base (P2P)
__device__ __forceinline__ void widen_bf16x8(const uint4& v, float s, uint4& lo, uint4& hi) {
  const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), d = unpack_bf16x2(v.w);
  lo = make_uint4(__float_as_uint(a.x * s), __float_as_uint(a.y * s), __float_as_uint(b.x * s), __float_as_uint(b.y * s));
  hi = make_uint4(__float_as_uint(c.x * s), __float_as_uint(c.y * s), __float_as_uint(d.x * s), __float_as_uint(d.y * s));
}

// OUT_F32: the result is broadcast as fp32 into `out_*` (element i of the bf16 range -> fp32 element i of the output range)
template <bool MULTICAST, int UNROLL, bool OUT_F32>
__global__ void __launch_bounds__(256) allreduce_slice_kernel(char* mc_slice, char* const* peer_bases, long long byte_off, long long units,
                                                              int world, float scale, char* out_mc_slice, char* const* out_peer_bases,
                                                              long long out_byte_off) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if constexpr (MULTICAST) {
    for (; i + (UNROLL - 1) * stride < units; i += UNROLL * stride) {
      uint4 v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
                     : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(mc_slice + (i + u * stride) * 16) : "memory");
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        if constexpr (OUT_F32) {
          uint4 lo, hi;
          widen_bf16x8(v[u], scale, lo, hi);
          char* o = out_mc_slice + (i + u * stride) * 32;
          asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w) : "memory");
          asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 16), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
        } else {
          const uint4 o = scale_bf16x8(v[u], scale);
          asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_slice + (i + u * stride) * 16), "r"(o.x), "r"(o.y),
                       "r"(o.z), "r"(o.w) : "memory");
        }
      }
    }
    for (; i < units; i += stride) {
      uint4 v;
      asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
                   : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(mc_slice + i * 16) : "memory");
      if constexpr (OUT_F32) {
        uint4 lo, hi;
        widen_bf16x8(v, scale, lo, hi);
        char* o = out_mc_slice + i * 32;
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w) : "memory");
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 16), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
      } else {
        const uint4 o = scale_bf16x8(v, scale);
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_slice + i * 16), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w)
                     : "memory");
      }
    }
  } else {
    for (; i < units; i += stride) {
      float acc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.f;
      for (int p = 0; p < world; ++p) {
        uint4 v;
        asm volatile("ld.volatile.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "l"(peer_bases[p] + byte_off + i * 16) : "memory");
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) { const float2 f = unpack_bf16x2(w[k]); acc[2 * k] += f.x; acc[2 * k + 1] += f.y; }
      }
      if constexpr (OUT_F32) {
        // the sum is rounded to bf16 first, exactly as the switch does, so that both transports give the same bits
        uint4 o, lo, hi;
        o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]); o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
        widen_bf16x8(o, scale, lo, hi);
        for (int p = 0; p < world; ++p) {
          *reinterpret_cast<uint4*>(out_peer_bases[p] + out_byte_off + i * 32) = lo;
          *reinterpret_cast<uint4*>(out_peer_bases[p] + out_byte_off + i * 32 + 16) = hi;
        }
      } else {
        uint4 o;
        o.x = pack_bf16x2(acc[0] * scale, acc[1] * scale); o.y = pack_bf16x2(acc[2] * scale, acc[3] * scale);
        o.z = pack_bf16x2(acc[4] * scale, acc[5] * scale); o.w = pack_bf16x2(acc[6] * scale, acc[7] * scale);
        for (int p = 0; p < world; ++p) *reinterpret_cast<uint4*>(peer_bases[p] + byte_off + i * 16) = o;
      }
    }
  }
}

}  // namespace vb

extern "C" int vb_rank_barrier(const vb_exchange_args* a, int32_t slot, void* stream) {
  using namespace vb;
  VB_REQUIRE(a != nullptr && a->flag_ptrs != nullptr, "null exchange arguments");
  VB_REQUIRE(a->world >= 1 && a->world <= 32 && a->rank >= 0 && a->rank < a->world, "rank / world");
  VB_REQUIRE(slot >= 0 && slot < a->flag_slots, "flag slot out of range");
  rank_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<uint32_t* const*>(a->flag_ptrs), a->rank, a->world, slot);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_allreduce_mean_bf16(const vb_exchange_args* a, int64_t lo, int64_t hi, int64_t out_lo, void* stream) {
  using namespace vb;
  VB_REQUIRE(a != nullptr && a->flag_ptrs != nullptr, "null exchange arguments");
  VB_REQUIRE(a->world >= 1 && a->world <= 32 && a->rank >= 0 && a->rank < a->world, "rank / world");
  VB_REQUIRE(a->mc_base != nullptr || a->peer_bases != nullptr, "neither a multicast mapping nor peer pointers");
  VB_REQUIRE(lo >= 0 && hi > lo && lo % 8 == 0 && hi % 8 == 0 && out_lo >= 0 && out_lo % 8 == 0, "the range must be 16-byte aligned");
  VB_REQUIRE(a->flag_slots >= 2, "two flag slots (of `world` words each) are needed");
  cudaStream_t s = (cudaStream_t)stream;
  // this rank's slice of the range, in 16-byte units
  const long long units_all = (hi - lo) / 8;
  const long long per = (units_all + a->world - 1) / a->world;
  const long long u0 = per * a->rank < units_all ? per * a->rank : units_all;
  const long long u1 = u0 + per < units_all ? u0 + per : units_all;
  rank_barrier_kernel<<<1, 32, 0, s>>>(reinterpret_cast<uint32_t* const*>(a->flag_ptrs), a->rank, a->world, 0);
  VB_CUDA_CHECK(cudaGetLastError());
  if (u1 > u0) {
    const long long units = u1 - u0;
    const long long byte_off = (lo / 8 + u0) * 16;
    // measured (tools/bench_exchange.py, 498 MB): 8 GPUs 1 054 us with 16 .. 148 CTAs alike (busbw 827 GB/s; NCCL 1 378 us), 2 GPUs
    // 2 495 / 1 383 / 1 251 us with 16 / 32 / 48 CTAs -- 32 keeps the SM footprint small without starving the 2-GPU case
    int ctas = a->ctas > 0 ? a->ctas : 32;
    const long long need = (units + 256 * 4 - 1) / (256 * 4);
    if (need < ctas) ctas = static_cast<int>(need);
    const float scale = 1.0f / static_cast<float>(a->world);
    const bool f32 = a->out_mc_base != nullptr || a->out_peer_bases != nullptr;
    const long long out_off = (out_lo + u0 * 8) * 4;        // fp32 bytes
    char* const* peers = reinterpret_cast<char* const*>(a->peer_bases);
    char* const* out_peers = reinterpret_cast<char* const*>(a->out_peer_bases);
    if (a->mc_base != nullptr && (!f32 || a->out_mc_base != nullptr)) {
      char* mc = static_cast<char*>(a->mc_base) + byte_off;
      if (f32) allreduce_slice_kernel<true, 4, true><<<ctas, 256, 0, s>>>(mc, nullptr, byte_off, units, a->world, scale,
                                                                         static_cast<char*>(a->out_mc_base) + out_off, nullptr, out_off);
      else     allreduce_slice_kernel<true, 4, false><<<ctas, 256, 0, s>>>(mc, nullptr, byte_off, units, a->world, scale, nullptr, nullptr, 0);
    } else {
      VB_REQUIRE(peers != nullptr && (!f32 || out_peers != nullptr), "peer pointers are needed when a multicast mapping is missing");
      if (f32) allreduce_slice_kernel<false, 1, true><<<ctas, 256, 0, s>>>(nullptr, peers, byte_off, units, a->world, scale, nullptr, out_peers, out_off);
      else     allreduce_slice_kernel<false, 1, false><<<ctas, 256, 0, s>>>(nullptr, peers, byte_off, units, a->world, scale, nullptr, nullptr, 0);
    }
    VB_CUDA_CHECK(cudaGetLastError());
  }
  rank_barrier_kernel<<<1, 32, 0, s>>>(reinterpret_cast<uint32_t* const*>(a->flag_ptrs), a->rank, a->world, 1);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}
