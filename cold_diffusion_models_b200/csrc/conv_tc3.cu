// Halo-tile variant of the tcgen05 tap-list convolution for stride-1 convolutions whose taps lie in [-1, 1]^2 (every dense 3x3
// forward and data-gradient convolution of the network, plus the fused [3x3 | 1x1] pairs).
//
// Why: ncu of the round-2 kernel (profiles/ncu_conv_fwd_r02*.txt) shows the N = 64 / 128 layers bound by the L2 -> SM fabric, not by
// shared memory or by the issuing thread: conv_tc.cu fetches one 16 KB activation tile PER TAP AND CHANNEL CHUNK, i.e. every input
// pixel nine times (3.76 GB of L2 reads for a layer whose input is 0.27 GB; ~11 KB/clk chip-wide, tensor pipe 37 % busy).
//
// Here the 128 GEMM rows are a 16 x 8 pixel patch and ONE TMA box {32 ch, 10, 18, 1} brings its 18 x 10 halo patch per channel
// chunk (23 KB instead of 9 x 16 KB).  All nine taps read that patch through shifted shared-memory descriptors: a tap (dy, dx)
// starts ((1 + dy) * 10 + 1 + dx) rows of 128 bytes further and the 8-row core-matrix groups (= the 8 pixels of one image row of
// the patch) are 10 rows = 1280 bytes apart (descriptor stride-byte-offset).  The 128-byte swizzle is a function of the absolute
// shared-memory address for both the TMA write and the MMA read, so row-shifted descriptor starts address the right data
// (same property the weight-gradient kernel relies on, profiles/wgrad_modes_r01.txt).  Weights stream through their own
// ring of [tap][chunk] tiles.  Everything else (warp roles, double-buffered TMEM accumulators, fused epilogue) is conv_tc.cu's.
#include "tc_common.cuh"

namespace {

constexpr int kPH = 16, kPW = 8;                     // pixel patch = 128 GEMM rows
constexpr int kHH = kPH + 2, kHW = kPW + 2;          // halo patch
constexpr int kABytesTx = kHH * kHW * 128;           // 23040 bytes per TMA box
constexpr int kAStage = 24 * 1024;                   // stage pitch (1024-byte aligned)

struct Tc3Params {
  int B, H, W;
  int tiles_x, tiles_y, tiles_co, total_tiles;
  int Cout;
  int nsrc;
  int ntaps[2];
  int kchunks[2];
  int8_t dy[2][CD_MAX_TAPS];
  int8_t dx[2][CD_MAX_TAPS];
  float* out; int out_ld;
  const float* bias;
  const float* resid; int resid_ld;
  int act; int round_tf32;
  float* out2; int out2_ld;
  const float* aux; int aux_ld;
  int vec8;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
// K-major SWIZZLE_128B descriptor with an explicit stride between 8-row groups
__device__ __forceinline__ uint64_t make_desc_sbo(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// ASTAGES / BSTAGES: halo-patch and weight-tile rings; EPI: epilogue warps (16, or 8 with two CTAs per SM: conv_tc.cu explains)
template <int BN, int BSTAGES, int ASTAGES = 3, int EPI = 16>
__global__ void __launch_bounds__(64 + 32 * EPI, EPI == 16 ? 1 : 2)
conv_tc3_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                const __grid_constant__ CUtensorMap mapB0, const __grid_constant__ CUtensorMap mapB1,
                const Tc3Params p) {
  constexpr int kBBytes = BN * 128;
  constexpr uint32_t kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(BN >> 3) << 17) | (uint32_t(128 >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* smemA = smem;
  uint8_t* smemB = smem + ASTAGES * kAStage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smemB + BSTAGES * kBBytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + ASTAGES;
  uint64_t* b_full = bars + 2 * ASTAGES;
  uint64_t* b_empty = b_full + BSTAGES;
  uint64_t* tmem_full = b_empty + BSTAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint32_t* arow_tab = tmem_slot + 1;                    // [2][CD_MAX_TAPS + 1]: first patch row of every tap, in 16-byte units
  static_assert((2 * ASTAGES + 2 * BSTAGES + 4) * 8 + 4 + 4 * 2 * (CD_MAX_TAPS + 1) <= 512, "barrier block is 512 bytes");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < ASTAGES; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < BSTAGES; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    mbar_init(&tmem_full[0], 1); mbar_init(&tmem_full[1], 1);
    mbar_init(&tmem_empty[0], EPI); mbar_init(&tmem_empty[1], EPI);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 2 && lane < 2 * (CD_MAX_TAPS + 1)) {
    const int s = lane / (CD_MAX_TAPS + 1), t = lane % (CD_MAX_TAPS + 1);
    arow_tab[lane] = (t < CD_MAX_TAPS && t < p.ntaps[s]) ? static_cast<uint32_t>(((1 + p.dy[s][t]) * kHW + 1 + p.dx[s][t]) * 8) : 0u;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: one halo patch per channel chunk, one weight tile per (chunk, tap) =====================
    uint32_t sa = 0, pha = 0, sb = 0, phb = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int co_t = tile % p.tiles_co;
      int mt = tile / p.tiles_co;
      const int tx = mt % p.tiles_x; mt /= p.tiles_x;
      const int ty = mt % p.tiles_y;
      const int n = mt / p.tiles_y;
      const int x0 = tx * kPW - 1, y0 = ty * kPH - 1, co0 = co_t * BN;
      for (int s = 0; s < p.nsrc; ++s) {
        const CUtensorMap* mA = s ? &mapA1 : &mapA0;
        const CUtensorMap* mB = s ? &mapB1 : &mapB0;
        const int nt = p.ntaps[s];
        for (int kc = 0; kc < p.kchunks[s]; ++kc) {
          mbar_wait(&a_empty[sa], pha ^ 1u);
          if (elect_one()) {
            mbar_expect_tx(&a_full[sa], kABytesTx);
            tma_load_4d(smem_u32(smemA + sa * kAStage), mA, &a_full[sa], kc * 32, x0, y0, n);
          }
          __syncwarp();
          if (++sa == ASTAGES) { sa = 0; pha ^= 1u; }
          for (int tap = 0; tap < nt; ++tap) {
            mbar_wait(&b_empty[sb], phb ^ 1u);
            if (elect_one()) {
              mbar_expect_tx(&b_full[sb], kBBytes);
              tma_load_3d(smem_u32(smemB + sb * kBBytes), mB, &b_full[sb], kc * 32, co0, tap);
            }
            __syncwarp();
            if (++sb == BSTAGES) { sb = 0; phb ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Everything this warp executes per tap is on the critical path of a 128-clk (N = 64) MMA group: ring positions are
    // incremented (no division by the stage count), and the first patch row of every tap comes from a shared-memory table whose
    // next entry is fetched before the barrier wait (an indexed load of the kernel parameters costs a constant-cache round trip).
    uint32_t sa = 0, pha = 0, sb = 0, phb = 0, tcount = 0;
    const uint64_t descA0 = make_desc_sbo(smem_u32(smemA), kHW * 128);      // 8-row groups (one patch row) are 10 rows apart
    const uint64_t descB0 = make_desc_sbo(smem_u32(smemB), 1024);
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
      const uint32_t acc = tcount & 1u, accph = (tcount >> 1) & 1u;
      mbar_wait(&tmem_empty[acc], accph ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      uint32_t first = 0;
      for (int s = 0; s < p.nsrc; ++s) {
        const int nt = p.ntaps[s];
        const uint32_t* tab = arow_tab + s * (CD_MAX_TAPS + 1);
        for (int kc = 0; kc < p.kchunks[s]; ++kc) {
          mbar_wait(&a_full[sa], pha);
          const uint64_t da0 = descA0 + static_cast<uint64_t>(sa * uint32_t(kAStage >> 4));
          uint32_t arow8 = tab[0];
          for (int tap = 0; tap < nt; ++tap) {
            const uint32_t arow8_next = tab[tap + 1];                  // table has a spare entry
            mbar_wait(&b_full[sb], phb);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t da = da0 + static_cast<uint64_t>(arow8);
              const uint64_t db = descB0 + static_cast<uint64_t>(sb * uint32_t(kBBytes >> 4));
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                mma_tf32(tmem_d, da + uint64_t(kk * 2), db + uint64_t(kk * 2), kIdesc, (first | uint32_t(kk)) != 0 ? 1u : 0u);
              tc_commit(&b_empty[sb]);
              if (tap == nt - 1) tc_commit(&a_empty[sa]);              // the halo patch is free once its last tap retired
            }
            __syncwarp();
            first = 1;
            arow8 = arow8_next;
            if (++sb == BSTAGES) { sb = 0; phb ^= 1u; }
          }
          if (++sa == ASTAGES) { sa = 0; pha ^= 1u; }
        }
      }
      if (elect_one()) tc_commit(&tmem_full[acc]);
      __syncwarp();
    }
  } else {
    // ===================== epilogue (warps 2..17): same fused epilogue as conv_tc.cu, rows = 16 x 8 pixel patch =====================
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    const int m = q * 32 + lane;
    const int xx = m & (kPW - 1), yy = m >> 3;
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
      const int co_t = tile % p.tiles_co;
      int mt = tile / p.tiles_co;
      const int tx = mt % p.tiles_x; mt /= p.tiles_x;
      const int ty = mt % p.tiles_y;
      const int b = mt / p.tiles_y;
      const int gx = tx * kPW + xx, gy = ty * kPH + yy;
      const int co0 = co_t * BN;
      const long long pix = (static_cast<long long>(b) * p.H + gy) * p.W + gx;
      float* orow = p.out + pix * p.out_ld;
      const float* rrow = p.resid ? p.resid + pix * p.resid_ld : nullptr;
      float* o2row = p.out2 ? p.out2 + pix * p.out2_ld : nullptr;
      const float* arow = p.aux ? p.aux + pix * p.aux_ld : nullptr;

      const uint32_t acc = tcount & 1u, accph = (tcount >> 1) & 1u;
      mbar_wait(&tmem_full[acc], accph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int c = cg * 32; c < BN; c += 32 * (EPI / 4)) {
        uint32_t r[32];
        tmem_ld32(taddr + c, r);
        if (co0 + c < p.Cout) {
          const int nvalid = min(32, p.Cout - (co0 + c));
          if (nvalid == 32 && p.vec8) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[j + e]);
              if (p.bias) { float t[8]; ldg8(p.bias + co0 + c + j, t);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] += t[e]; }
              if (rrow) { float t[8]; ldg8(rrow + co0 + c + j, t);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] += t[e]; }
              if (o2row) stg8(o2row + co0 + c + j, v);
              if (p.act == CD_ACT_GELU) {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = cd_gelu(v[e]);
              } else if (p.act == CD_ACT_GELU_BWD) {
                float t[8]; ldg8(arow + co0 + c + j, t);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] *= cd_gelu_grad(t[e]);
              }
              if (p.round_tf32) {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = cd_round_tf32(v[e]);
              }
              stg8(orow + co0 + c + j, v);
            }
          } else {
            // fully unrolled with a predicate: a run-time index into r[] would move the whole accumulator chunk to local memory
            // (8 STL.128 + reloads per chunk on every path, also the vector one)
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (j >= nvalid) break;
              float v = __uint_as_float(r[j]);
              if (p.bias) v += p.bias[co0 + c + j];
              if (rrow) v += rrow[co0 + c + j];
              if (o2row) o2row[co0 + c + j] = v;
              if (p.act == CD_ACT_GELU) v = cd_gelu(v);
              else if (p.act == CD_ACT_GELU_BWD) v *= cd_gelu_grad(arow[co0 + c + j]);
              if (p.round_tf32) v = cd_round_tf32(v);
              orow[co0 + c + j] = v;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

int g_sms3 = 0;
}  // namespace
int cd_conv_tc_two_ctas_mask();        // conv_tc.cu (cd_conv_tc_set_two_ctas)
namespace {

template <int BN, int BSTAGES, int ASTAGES = 3, int EPI = 16>
int launch3(const CUtensorMap* maps, const Tc3Params& p, cudaStream_t st) {
  constexpr size_t smem = size_t(ASTAGES) * kAStage + size_t(BSTAGES) * BN * 128 + 1024 + 512;
  static_assert(smem <= 232448, "dynamic shared memory of one CTA (227 KB)");
  static_assert(EPI == 16 || 2 * (smem + 1024) <= 233472, "two CTAs per SM must fit the 228 KB of shared memory");
  static bool attr_done = false;
  if (!attr_done) {
    CD_CUDA(cudaFuncSetAttribute(conv_tc3_kernel<BN, BSTAGES, ASTAGES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  const int slots = g_sms3 * (EPI == 16 ? 1 : 2);
  const int grid = p.total_tiles < slots ? p.total_tiles : slots;
  conv_tc3_kernel<BN, BSTAGES, ASTAGES, EPI><<<grid, 64 + 32 * EPI, smem, st>>>(maps[0], maps[1], maps[2], maps[3], p);
  CD_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// returns 1 when the problem is not eligible (caller continues with conv_tc.cu), 0 on success, < 0 on error
int cd_conv_fwd_tc3(const CdConvDesc* d, cudaStream_t st) {
  if (d->nsrc < 1 || d->nsrc > 2) return 1;
  if (d->sy != 1 || d->sx != 1 || d->oys != 1 || d->oxs != 1 || d->oy0 != 0 || d->ox0 != 0) return 1;
  if (d->Hg % kPH != 0 || d->Wg % kPW != 0 || d->Ho != d->Hg || d->Wo != d->Wg) return 1;
  int ktotal = 0;
  bool any3x3 = false;
  for (int s = 0; s < d->nsrc; ++s) {
    const CdConvSrc& cs = d->s[s];
    if (cs.w_per_batch || cs.C % 32 != 0 || cs.C <= 0 || cs.H != d->Hg || cs.W != d->Wg || cs.ntaps < 1 || cs.ntaps > CD_MAX_TAPS) return 1;
    if ((reinterpret_cast<uintptr_t>(cs.src) & 15) || cs.ld % 4 || (reinterpret_cast<uintptr_t>(cs.w) & 15)) return 1;
    for (int t = 0; t < cs.ntaps; ++t) if (cs.dy[t] < -1 || cs.dy[t] > 1 || cs.dx[t] < -1 || cs.dx[t] > 1) return 1;
    if (cs.ntaps > 1) any3x3 = true;
    ktotal += cs.ntaps * (cs.C / 32);
  }
  if (!any3x3) return 1;                      // pure 1x1: the halo would only add traffic
  if ((reinterpret_cast<uintptr_t>(d->out) & 15) || d->out_ld % 4) return 1;
  if (d->resid && ((reinterpret_cast<uintptr_t>(d->resid) & 15) || d->resid_ld % 4)) return 1;
  if (d->out2 && ((reinterpret_cast<uintptr_t>(d->out2) & 15) || d->out2_ld % 4)) return 1;
  if (d->bias && (reinterpret_cast<uintptr_t>(d->bias) & 15)) return 1;
  if (d->act == CD_ACT_GELU_BWD && (!d->aux || (reinterpret_cast<uintptr_t>(d->aux) & 15) || d->aux_ld % 4)) return 1;
  EncodeTiledFn enc = get_encode();
  CD_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  if (!g_sms3) { int dev = 0; CD_CUDA(cudaGetDevice(&dev)); CD_CUDA(cudaDeviceGetAttribute(&g_sms3, cudaDevAttrMultiProcessorCount, dev)); }

  Tc3Params p{};
  p.B = d->B; p.H = d->Hg; p.W = d->Wg; p.Cout = d->Cout; p.nsrc = d->nsrc;
  p.tiles_x = d->Wg / kPW; p.tiles_y = d->Hg / kPH;
  const long long m_tiles = static_cast<long long>(p.tiles_x) * p.tiles_y * d->B;
  int BN = (d->Cout % 256 == 0) ? 256 : (d->Cout > 64 ? 128 : 64);
  if (BN == 256) {          // few M tiles (16x16, 32x32 levels): the narrower N tile when it saves whole waves
    const long long t256 = m_tiles * (d->Cout / 256), t128 = m_tiles * (d->Cout / 128);
    const double c256 = double((t256 + g_sms3 - 1) / g_sms3) * 256.0 / 0.92, c128 = double((t128 + g_sms3 - 1) / g_sms3) * 128.0 / 0.85;
    if (c128 < c256) BN = 128;
  }
  p.tiles_co = cd_cdiv(d->Cout, BN);
  p.total_tiles = static_cast<int>(m_tiles * p.tiles_co);
  p.out = d->out; p.out_ld = d->out_ld;
  p.bias = d->bias; p.resid = d->resid; p.resid_ld = d->resid_ld; p.act = d->act; p.round_tf32 = d->round_tf32;
  p.out2 = d->out2; p.out2_ld = d->out2_ld; p.aux = d->aux; p.aux_ld = d->aux_ld;
  {
    auto ok8 = [](const void* ptr, int ld) { return ptr == nullptr || ((reinterpret_cast<uintptr_t>(ptr) & 31) == 0 && ld % 8 == 0); };
    p.vec8 = ok8(d->out, d->out_ld) && ok8(d->out2, d->out2_ld) && ok8(d->resid, d->resid_ld) && ok8(d->aux, d->aux_ld) && ok8(d->bias, 8);
  }
  CUtensorMap maps[4];
  const CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_TFLOAT32;
  for (int s = 0; s < 2; ++s) {
    const CdConvSrc& cs = d->s[s < d->nsrc ? s : 0];
    p.ntaps[s] = cs.ntaps; p.kchunks[s] = cs.C / 32;
    for (int t = 0; t < cs.ntaps; ++t) { p.dy[s][t] = (int8_t)cs.dy[t]; p.dx[s][t] = (int8_t)cs.dx[t]; }
    {   // A: NHWC activations, dims {C, W, H, N}; box = one halo patch of one channel chunk
      cuuint64_t dims[4] = {(cuuint64_t)cs.C, (cuuint64_t)cs.W, (cuuint64_t)cs.H, (cuuint64_t)d->B};
      cuuint64_t strides[3] = {(cuuint64_t)cs.ld * 4, (cuuint64_t)cs.ld * 4 * cs.W, (cuuint64_t)cs.ld * 4 * cs.W * cs.H};
      cuuint32_t box[4] = {32, (cuuint32_t)kHW, (cuuint32_t)kHH, 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      CUresult r = enc(&maps[s], dt, 4, const_cast<float*>(cs.src), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      CD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(A%d, halo) failed: %d", s, (int)r);
    }
    {   // B: packed weights [tap][Cout][Cin]
      cuuint64_t dims[3] = {(cuuint64_t)cs.C, (cuuint64_t)d->Cout, (cuuint64_t)cs.ntaps};
      cuuint64_t strides[2] = {(cuuint64_t)cs.C * 4, (cuuint64_t)cs.C * 4 * d->Cout};
      cuuint32_t box[3] = {32, (cuuint32_t)BN, 1};
      cuuint32_t estr[3] = {1, 1, 1};
      CUresult r = enc(&maps[2 + s], dt, 3, const_cast<float*>(cs.w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      CD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(B%d) failed: %d", s, (int)r);
    }
  }
  (void)ktotal;
  const int two = cd_conv_tc_two_ctas_mask();
  if (BN == 256) return launch3<256, 4>(maps, p, st);
  if (BN == 128) return (two & 128) ? launch3<128, 3, 2, 8>(maps, p, st) : launch3<128, 8>(maps, p, st);
  return (two & 64) ? launch3<64, 6, 2, 8>(maps, p, st) : launch3<64, 12>(maps, p, st);
}
