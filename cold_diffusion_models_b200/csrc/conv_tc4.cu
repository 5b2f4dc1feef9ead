// Wide halo-tile variant of the tcgen05 tap-list convolution: 256 GEMM rows (a 16 x 16 pixel patch) per CTA for the stride-1
// 3x3 convolutions with Cout <= 128 (forward and data gradient; plus the fused [3x3 | 1x1] pairs and two-source concat inputs).
//
// Why (tools/micro/umma_rate.cu, profiles/umma_rate_r02.txt): with operands already in shared memory one SM retires a
// 128 x 64 x 8 tf32 MMA every 50 clk and a 128 x 128 x 8 one every 64 clk (707 / 1110 TFLOP/s chip-wide); conv_tc.cu reaches
// 102 / ~160 clk on those shapes because every MMA needs 6 / 8 KB of fresh operands from L2 (one 16 KB activation tile per
// tap and channel chunk: every input pixel is fetched nine times, and every 128-pixel tile re-streams the whole filter) and the
// L2 -> SM fabric delivers ~30-45 B/clk/SM.  The operand bytes per MMA have to come down:
//   * activations: ONE TMA box {32 ch, 18, 18, 1} brings the 18 x 18 halo patch of a channel chunk (41 KB) and all nine taps read
//     it through shifted shared-memory descriptors (start row (1 + dy) * 18 + 1 + dx, 8-row groups = 8 pixels of a patch row,
//     18 rows = 2304 bytes apart) -- 0.56 KB per MMA instead of 4 KB;
//   * weights: every weight tile (one tap, 32 input channels, BN outputs) feeds TWO row blocks -- the left and the right
//     16 x 8 half of the patch, each with its own TMEM accumulator -- so the filter is streamed once per 256 pixels;
//   * the issuing thread waits on one barrier and commits once per 8 MMAs instead of per 4.
// TMEM: 2 halves x BN columns, double-buffered across tiles (4 x BN <= 512 columns).  The 128-byte swizzle is a function of the
// absolute shared-memory address for both the TMA write and the MMA read, so row-shifted descriptor starts address the right
// data (conv_tc3.cu, wgrad_tc.cu rely on the same property; the unaligned 8-row groups cost ~20 % of the MMA rate:
// 60 instead of 50 clk at N = 64).  Warp roles and the fused epilogue are conv_tc.cu's.
#include "tc_common.cuh"

namespace {

constexpr int kPH = 16, kPW = 16;                    // pixel patch = 2 x 128 GEMM rows (left / right half)
constexpr int kHH = kPH + 2, kHW = kPW + 2;          // halo patch
constexpr int kABytesTx = kHH * kHW * 128;           // 41472 bytes per TMA box
constexpr int kAStage = 41 * 1024;                   // stage pitch (1024-byte aligned)

struct Tc4Params {
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
  int dbg;                        // timing experiments only (cd_conv_tc_set_debug): 1 = epilogue without global accesses and math, 2 = without TMEM loads either, 3 = math but no output stores, 4 = stores but no GELU
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

// ASTAGES / BSTAGES: halo-patch and weight-tile rings
template <int BN, int BSTAGES, int ASTAGES = 2, int EPI = 16>
__global__ void __launch_bounds__(64 + 32 * EPI, 1)
conv_tc4_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                const __grid_constant__ CUtensorMap mapB0, const __grid_constant__ CUtensorMap mapB1,
                const Tc4Params p) {
  constexpr int kBBytes = BN * 128;
  constexpr uint32_t kTmemCols = (4 * BN <= 256) ? 256 : 512;        // [2 tiles in flight][2 patch halves][BN]
  static_assert(4 * BN <= 512, "two double-buffered accumulator pairs must fit the 512 TMEM columns");
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
    // one elected thread runs the whole schedule
    if (elect_one()) {
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
            mbar_expect_tx(&a_full[sa], kABytesTx);
            tma_load_4d(smem_u32(smemA + sa * kAStage), mA, &a_full[sa], kc * 32, x0, y0, n);
            if (++sa == ASTAGES) { sa = 0; pha ^= 1u; }
            for (int tap = 0; tap < nt; ++tap) {
              mbar_wait(&b_empty[sb], phb ^ 1u);
              mbar_expect_tx(&b_full[sb], kBBytes);
              tma_load_3d(smem_u32(smemB + sb * kBBytes), mB, &b_full[sb], kc * 32, co0, tap);
              if (++sb == BSTAGES) { sb = 0; phb ^= 1u; }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer: one elected thread runs the whole loop =====================
    // Barrier polls, commits and MMAs pass through one in-order queue (tools/micro/umma_rate.cu): the barriers the NEXT tap needs
    // (its weight tile; the next halo patch when a chunk ends; the next accumulator pair when the tile ends) are polled between
    // the two row blocks of THIS tap, so the tensor core has four queued MMAs to execute while the thread commits, advances the
    // rings and builds the next descriptors.  The first patch row of every tap comes from a shared-memory table.
    if (elect_one()) {
      uint32_t sa = 0, pha = 0, sb = 0, phb = 0, tcount = 0;
      const uint64_t descA0 = make_desc_sbo(smem_u32(smemA), kHW * 128);      // 8-row groups (8 pixels of a patch row) are 18 rows apart
      const uint64_t descB0 = make_desc_sbo(smem_u32(smemB), 1024);
      if (static_cast<int>(blockIdx.x) < p.total_tiles) {
        mbar_wait(&tmem_empty[0], 1u);
        mbar_wait(&a_full[0], 0u);
        mbar_wait(&b_full[0], 0u);
      }
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
        const uint32_t acc = tcount & 1u;
        const uint32_t tmem_d = tmem_base + acc * (2 * BN);
        const bool more_tiles = tile + static_cast<int>(gridDim.x) < p.total_tiles;
        uint32_t first = 0;
        for (int s = 0; s < p.nsrc; ++s) {
          const int nt = p.ntaps[s];
          const bool last_src = s == p.nsrc - 1;
          // first patch row of every tap in REGISTERS: a shared-memory load in the loop would sit in the same in-order queue as the
          // MMAs and return only after the MMAs in front of it have been handed to the tensor core
          uint32_t arow[9];
#pragma unroll
          for (int t = 0; t < 9; ++t) arow[t] = arow_tab[s * (CD_MAX_TAPS + 1) + (t < nt ? t : 0)];
          for (int kc = 0; kc < p.kchunks[s]; ++kc) {
            const uint64_t da0 = descA0 + static_cast<uint64_t>(sa * uint32_t(kAStage >> 4));
            uint32_t san = sa + 1, phan = pha;
            if (san == ASTAGES) { san = 0; phan ^= 1u; }
            const bool last_chunk = last_src && kc == p.kchunks[s] - 1;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              if (tap < nt) {
                tc_fence_after();
                const uint64_t da = da0 + static_cast<uint64_t>(arow[tap]);
                const uint64_t db = descB0 + static_cast<uint64_t>(sb * uint32_t(kBBytes >> 4));
                uint32_t sbn = sb + 1, phbn = phb;
                if (sbn == BSTAGES) { sbn = 0; phbn ^= 1u; }
                const bool last_tap = tap == nt - 1;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)                             // left patch half
                  mma_tf32(tmem_d, da + uint64_t(kk * 2), db + uint64_t(kk * 2), kIdesc, (first | uint32_t(kk)) != 0 ? 1u : 0u);
                const bool more = !(last_tap && last_chunk) || more_tiles;
                uint32_t ready = 0;
                if (more) {                                                // one non-blocking poll of what the next tap needs
                  ready = mbar_test(&b_full[sbn], phbn);
                  if (last_tap) {
                    ready &= mbar_test(&a_full[san], phan);
                    if (last_chunk) ready &= mbar_test(&tmem_empty[acc ^ 1u], (((tcount + 1) >> 1) & 1u) ^ 1u);
                  }
                }
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)                             // right half: 8 pixels = 8 rows of 128 bytes further
                  mma_tf32(tmem_d + BN, da + uint64_t(64 + kk * 2), db + uint64_t(kk * 2), kIdesc, (first | uint32_t(kk)) != 0 ? 1u : 0u);
                tc_commit(&b_empty[sb]);
                if (last_tap) tc_commit(&a_empty[sa]);                    // the halo patch is free once its last tap retired
                if (more && !ready) {                                      // late data: wait in the open, after the commit
                  mbar_wait(&b_full[sbn], phbn);
                  if (last_tap) {
                    mbar_wait(&a_full[san], phan);
                    if (last_chunk) mbar_wait(&tmem_empty[acc ^ 1u], (((tcount + 1) >> 1) & 1u) ^ 1u);
                  }
                }
                first = 1;
                sb = sbn; phb = phbn;
              }
            }
            sa = san; pha = phan;
          }
        }
        tc_commit(&tmem_full[acc]);
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..17): same fused epilogue as conv_tc.cu, rows = 16 x 8 pixel patch =====================
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    const int m = q * 32 + lane;
    const int xx = m & 7, yy = m >> 3;                       // pixel of this TMEM lane inside a 16 x 8 patch half
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
      const int co_t = tile % p.tiles_co;
      int mt = tile / p.tiles_co;
      const int tx = mt % p.tiles_x; mt /= p.tiles_x;
      const int ty = mt % p.tiles_y;
      const int b = mt / p.tiles_y;
      const int co0 = co_t * BN;
      const long long pix0 = (static_cast<long long>(b) * p.H + ty * kPH + yy) * p.W + tx * kPW + xx;

      const uint32_t acc = tcount & 1u, accph = (tcount >> 1) & 1u;
      mbar_wait(&tmem_full[acc], accph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * (2 * BN) + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int item = cg; item < 2 * (BN / 32); item += EPI / 4) {   // (patch half, 32-column chunk) pairs over the column groups
        const int half = item / (BN / 32), c = (item % (BN / 32)) * 32;
        const long long pix = pix0 + half * 8;
        float* orow = p.out + pix * p.out_ld;
        const float* rrow = p.resid ? p.resid + pix * p.resid_ld : nullptr;
        float* o2row = p.out2 ? p.out2 + pix * p.out2_ld : nullptr;
        const float* arow = p.aux ? p.aux + pix * p.aux_ld : nullptr;
        uint32_t r[32];
        if (p.dbg == 2) continue;
        tmem_ld32(taddr + half * BN + c, r);
        if (p.dbg == 1) continue;
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
              if (p.act == CD_ACT_GELU && p.dbg != 4) {
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
              if (p.dbg != 3 || v[0] == 12345.678f) stg8(orow + co0 + c + j, v);
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

int g_sms4 = 0;
int g_dbg4 = 0;
int g_force4 = 0;                  // cd_conv_tc_set_halo(6): take every eligible problem (tests)
}  // namespace
namespace {

template <int BN, int BSTAGES, int ASTAGES = 2, int EPI = 16>
int launch4(const CUtensorMap* maps, const Tc4Params& p, cudaStream_t st) {
  constexpr size_t smem = size_t(ASTAGES) * kAStage + size_t(BSTAGES) * BN * 128 + 1024 + 512;
  static_assert(smem <= 232448, "dynamic shared memory of one CTA (227 KB)");
  static bool attr_done = false;
  if (!attr_done) {
    CD_CUDA(cudaFuncSetAttribute(conv_tc4_kernel<BN, BSTAGES, ASTAGES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  const int grid = p.total_tiles < g_sms4 ? p.total_tiles : g_sms4;
  conv_tc4_kernel<BN, BSTAGES, ASTAGES, EPI><<<grid, 64 + 32 * EPI, smem, st>>>(maps[0], maps[1], maps[2], maps[3], p);
  CD_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" int cd_conv_tc_set_debug(int mode) { g_dbg4 = mode; return 0; }
void cd_conv_tc4_force(int on) { g_force4 = on; }

// returns 1 when the problem is not eligible (caller continues with conv_tc.cu), 0 on success, < 0 on error
int cd_conv_fwd_tc4(const CdConvDesc* d, cudaStream_t st) {
  if (d->nsrc < 1 || d->nsrc > 2) return 1;
  if (d->sy != 1 || d->sx != 1 || d->oys != 1 || d->oxs != 1 || d->oy0 != 0 || d->ox0 != 0) return 1;
  if (d->Hg % kPH != 0 || d->Wg % kPW != 0 || d->Ho != d->Hg || d->Wo != d->Wg) return 1;
  int ktotal = 0;
  bool any3x3 = false;
  for (int s = 0; s < d->nsrc; ++s) {
    const CdConvSrc& cs = d->s[s];
    if (cs.w_per_batch || cs.C % 32 != 0 || cs.C <= 0 || cs.H != d->Hg || cs.W != d->Wg || cs.ntaps < 1 || cs.ntaps > 9) return 1;
    if ((reinterpret_cast<uintptr_t>(cs.src) & 15) || cs.ld % 4 || (reinterpret_cast<uintptr_t>(cs.w) & 15)) return 1;
    for (int t = 0; t < cs.ntaps; ++t) if (cs.dy[t] < -1 || cs.dy[t] > 1 || cs.dx[t] < -1 || cs.dx[t] > 1) return 1;
    if (cs.ntaps > 1) any3x3 = true;
    ktotal += cs.ntaps * (cs.C / 32);
  }
  if (!any3x3) return 1;                      // pure 1x1: the halo would only add traffic
  if (d->Cout > 128) return 1;                // 4 x BN TMEM columns; the wide layers run on SM pairs (conv_tc2.cu)
  if (!g_sms4) { int dev = 0; CD_CUDA(cudaGetDevice(&dev)); CD_CUDA(cudaDeviceGetAttribute(&g_sms4, cudaDevAttrMultiProcessorCount, dev)); }
  {
    // 256-pixel tiles need many of them: measured per shape (profiles/conv_shapes_r02f_final.txt) the kernel wins on the
    // 128 x 128 level at batch 32 (2048 tiles: N = 64 layers 144-162 us against 205-207, N = 128 K = 576 157 against 162) and is
    // mixed at 64 x 64 (512 tiles = 3.5 waves); short-K N = 128 layers (K = 288: 139 against 132) stay on the pair kernel
    const long long tiles256 = static_cast<long long>(d->B) * (d->Hg / kPH) * (d->Wg / kPW);
    if (!g_force4 && d->Cout > 64 && ktotal < 18) return 1;
    if (!g_force4 && tiles256 < 6LL * g_sms4) return 1;
  }
  if ((reinterpret_cast<uintptr_t>(d->out) & 15) || d->out_ld % 4) return 1;
  if (d->resid && ((reinterpret_cast<uintptr_t>(d->resid) & 15) || d->resid_ld % 4)) return 1;
  if (d->out2 && ((reinterpret_cast<uintptr_t>(d->out2) & 15) || d->out2_ld % 4)) return 1;
  if (d->bias && (reinterpret_cast<uintptr_t>(d->bias) & 15)) return 1;
  if (d->act == CD_ACT_GELU_BWD && (!d->aux || (reinterpret_cast<uintptr_t>(d->aux) & 15) || d->aux_ld % 4)) return 1;
  EncodeTiledFn enc = get_encode();
  CD_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  Tc4Params p{};
  p.B = d->B; p.H = d->Hg; p.W = d->Wg; p.Cout = d->Cout; p.nsrc = d->nsrc;
  p.tiles_x = d->Wg / kPW; p.tiles_y = d->Hg / kPH;
  p.dbg = g_dbg4;
  const long long m_tiles = static_cast<long long>(p.tiles_x) * p.tiles_y * d->B;
  const int BN = d->Cout > 64 ? 128 : 64;
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
  if (BN == 128) return launch4<128, 8>(maps, p, st);
  return launch4<64, 12, 3>(maps, p, st);
}
