// Blackwell (sm_100a) primitives used by the tensor-core kernels: mbarrier, TMA, tcgen05 (UMMA) and TMEM.
// Inline PTX only; bit layouts follow the PTX ISA "tcgen05" chapter (shared-memory matrix descriptor,
// instruction descriptor) — cross-checked against cute/arch/mma_sm100_desc.hpp of CUTLASS.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gims {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a lost arrival must never hang the GPU.  On timeout the waiter records who/where in the
// device-global words below (the host turns a non-zero flag into an error) and gives up waiting; every later
// wait in the grid then falls through at once so that the kernel terminates.
// The wait is fully inline and CALL-free on purpose: a function call (printf, a noinline helper) inside the
// single-thread TMA / MMA loops makes the compiler keep loop-carried operands in vector registers, and every
// tcgen05.mma then pays an ELECT + R2UR.BROADCAST sequence per operand (measured: 78 instead of 47 clk per MMA).
static __device__ unsigned g_tc_timeout_flag;
static __device__ unsigned g_tc_timeout_info[4];     // blockIdx.x, blockIdx.y | blockIdx.z << 16, threadIdx.x, barrier smem address
__device__ __forceinline__ bool mbar_try_wait(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(addr), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  unsigned spins = 0;
  while (!mbar_try_wait(addr, parity)) {
    // try_wait suspends the thread for a hardware-defined slice (~microseconds) before it reports failure, so
    // a few hundred thousand failed probes are far beyond any legitimate wait in these kernels
    if (++spins > 400000u || ((spins & 255u) == 0 && *(volatile unsigned*)&g_tc_timeout_flag)) {
      if (atomicExch(&g_tc_timeout_flag, 1u) == 0u) {
        g_tc_timeout_info[0] = blockIdx.x; g_tc_timeout_info[1] = blockIdx.y | (blockIdx.z << 16);
        g_tc_timeout_info[2] = threadIdx.x; g_tc_timeout_info[3] = addr | (parity << 31);
      }
      return;
    }
  }
}

// ---- proxies / fences ---------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMA ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tiled load: coordinates {c0 = innermost (elements), c1 = row}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// ---- TMEM -----------------------------------------------------------------------------------------
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {   // one full warp
  static_assert(COLS == 32 || COLS == 64 || COLS == 128 || COLS == 256 || COLS == 512, "TMEM columns: power of 2 >= 32");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {        // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// 32 lanes x 32 columns of 32-bit: thread i of the warp receives row (lane base + i), columns [c, c+32)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- UMMA descriptors -----------------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major operand, SWIZZLE_128B, rows of exactly 128 bytes,
// 8-row swizzle atoms (1024 B) stacked contiguously along M/N:
//   [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (=64: 1024 B between atoms)
//   [46,48) version=1 | [49,52) base offset=0 (tile 1024-B aligned) | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)64 << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor, kind::tf32, fp32 accumulate, A and B K-major:
//   [4,6) D fmt = 1 (f32) | [7,10) A fmt = 2 (tf32) | [10,13) B fmt = 2 | [15] A major = 0 | [16] B major = 0
//   [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void umma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 columns of 32-bit, registers -> TMEM (thread i writes row lane base + i)
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
// 32 lanes x 16 columns of 32-bit, registers -> TMEM
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]^T
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// ---- 16-bit operands (kind::f16): fp16 or bf16, fp32 accumulate, K = 16 per instruction ---------------------------
// Instruction descriptor: same fields as the tf32 one; A / B format 0 = f16, 1 = bf16 (both operands must have the SAME
// format: a mixed f16 x bf16 instruction traps as illegal on sm_100a — tools/micro/f16_mix.cu).
__host__ __device__ constexpr uint32_t umma_idesc_f16(int m, int n, int fmt) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// D[tmem] (+)= A[tmem] * B[smem]^T, A packed two 16-bit values per 32-bit TMEM column (even k in the low half)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// two fp32 -> one 32-bit word of two 16-bit floats, `lo` in the low half; FMT 0 = f16, 1 = bf16 (round to nearest even)
template <int FMT>
__device__ __forceinline__ uint32_t pack16(float lo, float hi) {
  uint32_t d;
  if (FMT == 0) asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else          asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
template <int FMT>
__device__ __forceinline__ void unpack16(uint32_t d, float& lo, float& hi) {
  if (FMT == 0) {
    asm("{\n\t.reg .f16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(lo), "=f"(hi) : "r"(d));
  } else {
    lo = __uint_as_float(d << 16);
    hi = __uint_as_float(d & 0xffff0000u);
  }
}
// x0, x1 -> packed hi plane word and packed lo plane word: x = hi + lo with hi = rn16(x), lo = rn16(x - hi)
// (x - hi is exact in fp32).  fp16: |x - (hi + lo)| <= max(2^-22 |x|, 2^-25) for |x| < 65504.
template <int FMT>
__device__ __forceinline__ void split16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  hi = pack16<FMT>(x0, x1);
  float h0, h1;
  unpack16<FMT>(hi, h0, h1);
  lo = pack16<FMT>(x0 - h0, x1 - h1);
}

// Arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- tf32 split -----------------------------------------------------------------------------------
// x = hi + lo exactly; hi is x rounded to tf32 (10-bit mantissa, low 13 bits zero), lo the fp32 residual.
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  // Round-to-nearest, ties away from zero, on the sign-magnitude bit pattern (what cvt.rna.tf32.f32 computes, and
  // what gims_b200/packing.py does for the weights) — two full-rate integer ops; the cvt runs at 16 per clock
  // per SM and was the longest part of the A-split and of the softmax warps' P-split.
  const uint32_t h = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  hi = __uint_as_float(h);
  lo = x - hi;
}

// Host: driver entry point for cuTensorMapEncodeTiled without linking libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled();
// fp32 row-major [rows][cols] (row stride ld floats), box = box_rows x 32 floats, SWIZZLE_128B
int make_tmap_f32_k32(CUtensorMap* out, const float* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);
// 16-bit row-major [rows][cols] (row stride ld elements, ld % 8 == 0), box = box_rows x 64 elements (128 bytes), SWIZZLE_128B
int make_tmap_16_k64(CUtensorMap* out, const void* base, int fmt, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);

}  // namespace tc
}  // namespace gims
