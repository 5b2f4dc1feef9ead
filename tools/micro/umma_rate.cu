// How fast does ONE SM (or one SM pair) retire tcgen05.mma.kind::tf32 instructions when nothing else happens?
//
// Operands sit in shared memory (uninitialised bytes reinterpreted as tf32: zeros and denormal patterns cost the same), one
// elected thread issues `iters` K chunks of 4 MMAs (K = 8 each: one 128-byte SWIZZLE_128B row of 32 floats per chunk, the
// descriptor start advancing by 32 bytes per MMA exactly as conv_tc.cu does), cycling over `stages` stage buffers, then
// commits to an mbarrier and waits.  No TMA, no epilogue, no global memory traffic: the time per MMA is what the tensor
// core + its shared-memory operand fetch can do for that shape.  All 148 SMs run the same loop at the same time (power and
// clocks as in a real kernel).
//
// Prints clk / MMA, the TFLOP/s that corresponds to on 148 SMs at the measured clock and the operand bytes per clk the
// tensor core pulled from shared memory.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate.bin umma_rate.cu && ./umma_rate.bin
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .pred px;\nelect.sync _|px, 0xffffffff;\n@px mov.s32 %0, 1;\n}\n" : "+r"(pred));
  return pred;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n"
               :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t kmajor_sw128_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// MN-major SWIZZLE_128B_BASE32B (the weight-gradient operands: wgrad_tc.cu)
__device__ __forceinline__ uint64_t mnmajor_desc(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(1) << 61;
  return d;
}

// KIND: 0 = tf32 K-major, 1 = bf16 K-major (K = 16 per MMA: same operand bytes), 2 = tf32 MN-major,
//       3 = tf32 K-major with the A operand addressed as conv_tc3.cu does (rows of a 10-pixel-wide halo tile: 8-row groups
//           1280 bytes apart, start shifted by whole 128-byte rows)
// COPY: a second warp streams global memory into OTHER shared-memory stages with cp.async.bulk for as long as the MMAs run
//       (what the TMA producer of a real kernel does): do the bulk writes take shared-memory bandwidth from the operand reads?
template <int BN, int KIND, bool COPY = false>
__global__ void __launch_bounds__(128, 1)
umma1_kernel(int iters, int stages, long long* cycles, const float* __restrict__ gsrc = nullptr, long long* copied = nullptr) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint64_t cbar;
  __shared__ uint32_t slot;
  __shared__ volatile int done;
  constexpr int kA = KIND == 4 ? 41 * 1024 : KIND == 3 ? 24 * 1024 : 128 * 128, kB = BN * 128;
  constexpr uint32_t fmt = KIND == 1 ? 1u : 2u;
  constexpr uint32_t major = KIND == 2 ? ((1u << 15) | (1u << 16)) : 0u;
  constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | major | (uint32_t(BN >> 3) << 17) | (uint32_t(128 >> 4) << 24);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&cbar, 1); done = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (warp == 0) {
    const uint32_t base = smem_u32(smem);
    long long t0 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int c = 0; c < iters; ++c) {
        const int s = c % stages;
        const uint32_t a = base + s * (kA + kB), b = a + kA;
        if (KIND == 2) {
          // one chunk = 32 pixels (K) x 128 co / BN ci: 8 pixels (K = 8) per MMA = 1024 B further in both operands
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = mnmajor_desc(a + k * 1024, 4096), db = mnmajor_desc(b + k * 1024, 4096);
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                         :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"((c | k) ? 1u : 0u));
          }
        } else if (KIND == 4) {
          // conv_tc4.cu: 18-pixel-wide halo tile, tap (dy, dx) starts at row dy * 18 + dx, 8-row groups 2304 bytes apart, two row
          // blocks (patch halves, 1024 bytes apart) per weight tile with their own accumulators
          const int tap = c % 9;
          uint64_t da0 = kmajor_sw128_desc(a + ((tap / 3) * 18 + tap % 3) * 128);
          da0 = (da0 & ~(static_cast<uint64_t>(0x3FFF) << 32)) | (static_cast<uint64_t>(2304 >> 4) << 32);
          const uint64_t db0 = kmajor_sw128_desc(b);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                         :: "r"(tmem + (c & 1) * BN), "l"(da0 + (c & 1) * 64 + 2 * k), "l"(db0 + 2 * k), "r"(idesc), "r"(((c >> 1) | k) ? 1u : 0u));
        } else if (KIND == 3) {
          // tap (dy, dx) of a 3x3 kernel: start row = dy * 10 + dx of the halo tile
          const int tap = c % 9;
          uint64_t da0 = kmajor_sw128_desc(a + ((tap / 3) * 10 + tap % 3) * 128);
          da0 = (da0 & ~(static_cast<uint64_t>(0x3FFF) << 32)) | (static_cast<uint64_t>(1280 >> 4) << 32);
          const uint64_t db0 = kmajor_sw128_desc(b);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                         :: "r"(tmem), "l"(da0 + 2 * k), "l"(db0 + 2 * k), "r"(idesc), "r"((c | k) ? 1u : 0u));
        } else {
          const uint64_t da0 = kmajor_sw128_desc(a), db0 = kmajor_sw128_desc(b);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (KIND == 1)
              asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                           :: "r"(tmem), "l"(da0 + 2 * k), "l"(db0 + 2 * k), "r"(idesc), "r"((c | k) ? 1u : 0u));
            else
              asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                           :: "r"(tmem), "l"(da0 + 2 * k), "l"(db0 + 2 * k), "r"(idesc), "r"((c | k) ? 1u : 0u));
          }
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    if (elect_one()) { cycles[blockIdx.x] = clock64() - t0; done = 1; }   // the elected lane is the same one (lowest active)
  } else if (COPY && warp == 2) {
    // bulk copies of one stage (A + B bytes) at a time into the stages after the ones the MMAs read
    if (elect_one()) {
      const uint32_t bytes = kA + kB;
      const uint32_t dst0 = smem_u32(smem) + stages * (kA + kB);
      const char* src = reinterpret_cast<const char*>(gsrc) + static_cast<size_t>(blockIdx.x) * 8 * bytes;
      long long n = 0;
      uint32_t ph = 0;
      while (!done) {
        const uint32_t dst = dst0 + (n & 1) * bytes;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&cbar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(dst), "l"(src + (n & 7) * bytes), "r"(bytes), "r"(smem_u32(&cbar)) : "memory");
        mbar_wait(&cbar, ph); ph ^= 1;
        ++n;
      }
      copied[blockIdx.x] = n * bytes;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
}

// A faithful miniature of the convolution main loop: a producer warp refills stage s with ONE bulk copy (A + B bytes, L2-resident
// source) as soon as the MMAs that read it have retired (tcgen05.commit -> empty[s]); the MMA warp waits for full[s], issues
// MPS MMAs (4 = one 32-channel chunk; 8 = the two row blocks of conv_tc4.cu) and commits.  clk / MMA against the number of
// stages gives the refill latency: with S stages of one chunk each, clk per chunk = max(MMA time, latency / S).
template <int BN, int MPS, int COPYB = 0, bool AHEAD = false>
__global__ void __launch_bounds__(128, 1)
pipe_kernel(int iters, int stages, long long* cycles, const float* __restrict__ gsrc) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  __shared__ uint64_t full[16], empty[16], bar;
  __shared__ uint32_t slot;
  constexpr int kA = 128 * 128 * (MPS / 4), kB = BN * 128;
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(BN >> 3) << 17) | (uint32_t(128 >> 4) << 24);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  const uint32_t base = smem_u32(smem);
  const uint32_t bytes = kA + kB;
  const uint32_t cbytes = COPYB ? COPYB : bytes;          // COPYB: refill only that many bytes per stage (conv_tc4.cu: 8 KB of weights per 8 MMAs)
  if (warp == 0) {
    const char* src = reinterpret_cast<const char*>(gsrc) + static_cast<size_t>(blockIdx.x) * 8 * bytes;
    uint32_t s = 0, ph = 0;
    for (int c = 0; c < iters; ++c) {
      mbar_wait(&empty[s], ph ^ 1u);
      if (elect_one()) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&full[s])), "r"(cbytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(base + s * bytes), "l"(src + (c & 7) * bytes), "r"(cbytes), "r"(smem_u32(&full[s])) : "memory");
      }
      __syncwarp();
      if (++s == (uint32_t)stages) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    uint32_t s = 0, ph = 0;
    const long long t0 = clock64();
    if (!AHEAD) {
      for (int c = 0; c < iters; ++c) {
        mbar_wait(&full[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint32_t a = base + s * bytes, b = a + kA;
          const uint64_t da0 = kmajor_sw128_desc(a), db0 = kmajor_sw128_desc(b);
#pragma unroll
          for (int k = 0; k < MPS; ++k)
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                         :: "r"(tmem + (k / 4) * BN), "l"(da0 + (k / 4) * 1024 + 2 * (k % 4)), "l"(db0 + 2 * (k % 4)), "r"(idesc), "r"((c | (k % 4)) ? 1u : 0u));
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&empty[s])) : "memory");
        }
        __syncwarp();
        if (++s == (uint32_t)stages) { s = 0; ph ^= 1u; }
      }
      if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
      __syncwarp();
    } else if (elect_one()) {
      // ONE thread runs the whole loop (no per-stage elect / reconvergence), and the barrier of the NEXT stage is polled in the
      // middle of this stage's MMAs: the poll is a queued (MIO) operation like the MMAs, its result returns once the MMAs in
      // front of it have been handed to the tensor core, and the MMAs issued right after it keep the tensor core busy while the
      // thread commits, advances the ring and builds the next descriptors.
      mbar_wait(&full[0], 0);
      for (int c = 0; c < iters; ++c) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a = base + s * bytes, b = a + kA;
        const uint64_t da0 = kmajor_sw128_desc(a), db0 = kmajor_sw128_desc(b);
        uint32_t sn = s + 1, phn = ph;
        if (sn == (uint32_t)stages) { sn = 0; phn ^= 1u; }
#pragma unroll
        for (int k = 0; k < MPS; ++k) {
          if (k == MPS / 2 && c + 1 < iters) mbar_wait(&full[sn], phn);
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                       :: "r"(tmem + (k / 4) * BN), "l"(da0 + (k / 4) * 1024 + 2 * (k % 4)), "l"(db0 + 2 * (k % 4)), "r"(idesc), "r"((c | (k % 4)) ? 1u : 0u));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&empty[s])) : "memory");
        s = sn; ph = phn;
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    if (elect_one()) cycles[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
}

// SM pair: M = 256 (128 rows per CTA), each CTA holds BN/2 rows of B
template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
umma2_kernel(int iters, int stages, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  constexpr int kA = 128 * 128, kB = (BN / 2) * 128;
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(BN >> 3) << 17) | (uint32_t(256 >> 4) << 24);
  const int warp = threadIdx.x >> 5;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release;\nbarrier.cluster.wait.acquire;" ::: "memory");
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (warp == 0) {
    const uint32_t base = smem_u32(smem);
    long long t0 = 0;
    if (rank == 0 && elect_one()) {
      t0 = clock64();
      for (int c = 0; c < iters; ++c) {
        const int s = c % stages;
        const uint32_t a = base + s * (kA + kB), b = a + kA;
        const uint64_t da0 = kmajor_sw128_desc(a), db0 = kmajor_sw128_desc(b);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                       :: "r"(tmem), "l"(da0 + 2 * k), "l"(db0 + 2 * k), "r"(idesc), "r"((c | k) ? 1u : 0u));
      }
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                   :: "r"(smem_u32(&bar)), "h"(static_cast<uint16_t>(3)) : "memory");
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    if (rank == 0 && elect_one()) cycles[blockIdx.x >> 1] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release;\nbarrier.cluster.wait.acquire;" ::: "memory");
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
}

static int g_sms = 148;
static void launch1(void (*k)(int, int, long long*), int grid, size_t smem, int iters, int stages, long long* d) { k<<<grid, 128, smem>>>(iters, stages, d); }
static void launch1(void (*k)(int, int, long long*, const float*, long long*), int grid, size_t smem, int iters, int stages, long long* d) {
  k<<<grid, 128, smem>>>(iters, stages, d, nullptr, nullptr);
}

template <typename K>
static int run_copy(const char* name, K kernel, int grid, int bn, int a_bytes, int b_bytes, int stages) {
  const int iters = 4000;
  const size_t smem = static_cast<size_t>(stages + 2) * (a_bytes + b_bytes) + 1024;
  CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long *d = nullptr, *dc = nullptr;
  float* src = nullptr;
  CK(cudaMalloc(&d, sizeof(long long) * grid)); CK(cudaMalloc(&dc, sizeof(long long) * grid));
  CK(cudaMalloc(&src, static_cast<size_t>(grid) * 8 * (a_bytes + b_bytes)));
  CK(cudaMemset(src, 0, static_cast<size_t>(grid) * 8 * (a_bytes + b_bytes)));
  for (int rep = 0; rep < 3; ++rep) kernel<<<grid, 128, smem>>>(iters, stages, d, src, dc);
  CK(cudaDeviceSynchronize());
  std::vector<long long> h(grid), hc(grid);
  CK(cudaMemcpy(h.data(), d, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hc.data(), dc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
  std::vector<long long> hs = h; std::sort(hs.begin(), hs.end());
  const double clk = static_cast<double>(hs[grid / 2]) / (iters * 4.0);
  double cp = 0; for (int i = 0; i < grid; ++i) cp += static_cast<double>(hc[i]) / h[i];
  printf("%-34s N=%3d  %7.1f clk/MMA  %6.1f operand B/clk/SM  + %5.1f B/clk/SM of concurrent bulk copies into shared memory (L2-resident source)\n",
         name, bn, clk, (a_bytes + b_bytes) / 4.0 / clk, cp / grid);
  CK(cudaFree(d)); CK(cudaFree(dc)); CK(cudaFree(src));
  return 0;
}

template <typename K>
static int run_pipe(const char* name, K kernel, int grid, int bn, int mps, int stages) {
  const int iters = 4000;
  const int bytes = 128 * 128 * (mps / 4) + bn * 128;
  const size_t smem = static_cast<size_t>(stages) * bytes + 1024;
  CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long* d = nullptr;
  float* src = nullptr;
  CK(cudaMalloc(&d, sizeof(long long) * grid));
  CK(cudaMalloc(&src, static_cast<size_t>(grid) * 8 * bytes));
  CK(cudaMemset(src, 0, static_cast<size_t>(grid) * 8 * bytes));
  for (int rep = 0; rep < 3; ++rep) kernel<<<grid, 128, smem>>>(iters, stages, d, src);
  CK(cudaDeviceSynchronize());
  std::vector<long long> h(grid);
  CK(cudaMemcpy(h.data(), d, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
  std::sort(h.begin(), h.end());
  const double clk = static_cast<double>(h[grid / 2]) / (iters * (double)mps);
  printf("%-34s N=%3d  %2d stages x %5.1f KB, %d MMAs per stage: %7.1f clk/MMA  %6.1f B/clk/SM delivered, %6.0f clk per stage\n",
         name, bn, stages, bytes / 1024.0, mps, clk, bytes / (clk * mps), clk * mps);
  CK(cudaFree(d)); CK(cudaFree(src));
  return 0;
}

template <typename K>
static int run(const char* name, K kernel, int grid, int ncycles, int bn, int m, int kper, int a_bytes, int b_bytes, int stages) {
  const int iters = 4000;
  const size_t smem = static_cast<size_t>(stages) * (a_bytes + b_bytes) + 1024;
  CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long* d = nullptr;
  CK(cudaMalloc(&d, sizeof(long long) * grid));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 3; ++rep) launch1(kernel, grid, smem, iters, stages, d);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  const int reps = 10;
  for (int rep = 0; rep < reps; ++rep) launch1(kernel, grid, smem, iters, stages, d);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> h(ncycles);
  CK(cudaMemcpy(h.data(), d, sizeof(long long) * ncycles, cudaMemcpyDeviceToHost));
  std::sort(h.begin(), h.end());
  const double clk = static_cast<double>(h[ncycles / 2]) / (iters * 4.0);
  const double us = ms * 1e3 / reps;
  const double flops = 2.0 * m * bn * kper * iters * 4.0 * ncycles;            // per launch, all SMs (pairs)
  const double tflops = flops / (us * 1e-6) / 1e12;
  const double mhz = h[ncycles / 2] / us;                                       // SM clock implied by clock64 vs events (~)
  printf("%-34s N=%3d  %7.1f clk/MMA  %6.1f operand B/clk/SM  %7.1f TFLOP/s (whole GPU, launch time %.0f us, ~%.0f MHz)\n",
         name, bn, clk, (a_bytes + b_bytes) / 4.0 / clk, tflops, us, mhz);
  CK(cudaFree(d));
  return 0;
}

int main() {
  int dev = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
  printf("SMs: %d.  One CTA per SM (one cluster of 2 per SM pair), 4000 K chunks of 4 MMAs each, operands resident in shared memory.\n", g_sms);
  const int st = 6;
  int rc = 0;
  rc |= run("tf32 K-major, 1 CTA, M=128", umma1_kernel<32, 0>, g_sms, g_sms, 32, 128, 8, 128 * 128, 32 * 128, st);
  rc |= run("tf32 K-major, 1 CTA, M=128", umma1_kernel<64, 0>, g_sms, g_sms, 64, 128, 8, 128 * 128, 64 * 128, st);
  rc |= run("tf32 K-major, 1 CTA, M=128", umma1_kernel<128, 0>, g_sms, g_sms, 128, 128, 8, 128 * 128, 128 * 128, st);
  rc |= run("tf32 K-major, 1 CTA, M=128", umma1_kernel<256, 0>, g_sms, g_sms, 256, 128, 8, 128 * 128, 256 * 128, 4);
  rc |= run("bf16 K-major, 1 CTA, M=128", umma1_kernel<64, 1>, g_sms, g_sms, 64, 128, 16, 128 * 128, 64 * 128, st);
  rc |= run("bf16 K-major, 1 CTA, M=128", umma1_kernel<128, 1>, g_sms, g_sms, 128, 128, 16, 128 * 128, 128 * 128, st);
  rc |= run("bf16 K-major, 1 CTA, M=128", umma1_kernel<256, 1>, g_sms, g_sms, 256, 128, 16, 128 * 128, 256 * 128, 4);
  rc |= run("tf32 MN-major, 1 CTA, M=128", umma1_kernel<64, 2>, g_sms, g_sms, 64, 128, 8, 128 * 128, 64 * 128, st);
  rc |= run("tf32 MN-major, 1 CTA, M=128", umma1_kernel<128, 2>, g_sms, g_sms, 128, 128, 8, 128 * 128, 128 * 128, st);
  rc |= run("tf32 K-major, SM pair, M=256", umma2_kernel<64>, g_sms, g_sms / 2, 64, 256, 8, 128 * 128, 32 * 128, st);
  rc |= run("tf32 K-major, SM pair, M=256", umma2_kernel<128>, g_sms, g_sms / 2, 128, 256, 8, 128 * 128, 64 * 128, st);
  rc |= run("tf32 K-major, SM pair, M=256", umma2_kernel<256>, g_sms, g_sms / 2, 256, 256, 8, 128 * 128, 128 * 128, st);
  // one stage only: every MMA reads the same shared-memory bytes
  rc |= run("tf32 K-major, 1 CTA, 1 stage", umma1_kernel<64, 0>, g_sms, g_sms, 64, 128, 8, 128 * 128, 64 * 128, 1);
  // the halo-tile addressing of conv_tc3.cu
  rc |= run("tf32 K-major halo rows, 1 CTA", umma1_kernel<64, 3>, g_sms, g_sms, 64, 128, 8, 24 * 1024, 64 * 128, 4);
  rc |= run("tf32 K-major halo rows, 1 CTA", umma1_kernel<128, 3>, g_sms, g_sms, 128, 128, 8, 24 * 1024, 128 * 128, 4);
  rc |= run("tf32 K-major 18-wide halo rows", umma1_kernel<64, 4>, g_sms, g_sms, 64, 128, 8, 41 * 1024, 64 * 128, 2);
  rc |= run("tf32 K-major 18-wide halo rows", umma1_kernel<128, 4>, g_sms, g_sms, 128, 128, 8, 41 * 1024, 128 * 128, 2);
  // producer / consumer ring fed by bulk copies
  for (int st : {2, 4, 8}) rc |= run_pipe("ring: copy -> 4 MMAs -> commit", pipe_kernel<64, 4>, g_sms, 64, 4, st);
  for (int st : {2, 4, 6}) rc |= run_pipe("ring: copy -> 4 MMAs -> commit", pipe_kernel<128, 4>, g_sms, 128, 4, st);
  for (int st : {2, 4}) rc |= run_pipe("ring: copy -> 4 MMAs -> commit", pipe_kernel<256, 4>, g_sms, 256, 4, st);
  for (int st : {2, 4, 5}) rc |= run_pipe("ring: copy -> 8 MMAs -> commit", pipe_kernel<64, 8>, g_sms, 64, 8, st);
  rc |= run_pipe("ring: 8 KB copy -> 8 MMAs -> commit", pipe_kernel<64, 8, 8192>, g_sms, 64, 8, 5);
  rc |= run_pipe("ring: 8 KB copy -> 4 MMAs -> commit", pipe_kernel<64, 4, 8192>, g_sms, 64, 4, 8);
  rc |= run_pipe("ring: 16 KB copy -> 8 MMAs -> commit", pipe_kernel<128, 8, 16384>, g_sms, 128, 8, 4);
  rc |= run_pipe("ring: 16 KB copy -> 4 MMAs -> commit", pipe_kernel<128, 4, 16384>, g_sms, 128, 4, 6);
  rc |= run_pipe("ring: 8 KB copy -> 16 MMAs -> commit", pipe_kernel<64, 16, 8192>, g_sms, 64, 16, 3);
  // the same rings with one issuing thread for the whole loop and the next stage's barrier polled between the MMAs
  rc |= run_pipe("ring/1thread: 8 KB -> 8 MMAs", pipe_kernel<64, 8, 8192, true>, g_sms, 64, 8, 5);
  rc |= run_pipe("ring/1thread: 8 KB -> 4 MMAs", pipe_kernel<64, 4, 8192, true>, g_sms, 64, 4, 8);
  rc |= run_pipe("ring/1thread: 16 KB -> 8 MMAs", pipe_kernel<128, 8, 16384, true>, g_sms, 128, 8, 4);
  rc |= run_pipe("ring/1thread: 16 KB -> 4 MMAs", pipe_kernel<128, 4, 16384, true>, g_sms, 128, 4, 6);
  rc |= run_pipe("ring/1thread: full copy -> 4 MMAs", pipe_kernel<64, 4, 0, true>, g_sms, 64, 4, 8);
  rc |= run_pipe("ring/1thread: full copy -> 4 MMAs", pipe_kernel<128, 4, 0, true>, g_sms, 128, 4, 6);
  // MMAs + a producer writing shared memory at the same time
  rc |= run_copy("tf32 K-major + bulk copies", umma1_kernel<64, 0, true>, g_sms, 64, 128 * 128, 64 * 128, 4);
  rc |= run_copy("tf32 K-major + bulk copies", umma1_kernel<128, 0, true>, g_sms, 128, 128 * 128, 128 * 128, 4);
  rc |= run_copy("tf32 K-major + bulk copies", umma1_kernel<256, 0, true>, g_sms, 256, 128 * 128, 256 * 128, 2);
  return rc;
}
