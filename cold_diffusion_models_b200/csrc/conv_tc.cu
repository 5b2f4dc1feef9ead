// tcgen05 / TMA implicit-GEMM convolution for sm_100a (replaces the cuDNN calls behind nn.Conv2d /
// nn.ConvTranspose2d in the reference Unet: DB:105-109,149-154,173-174).
//
//   D[128 pixels x BN out-channels] (fp32, TMEM) += A[128 pixels x 32 ch] * B[BN x 32 ch]^T
//
// per (source, tap, 32-channel chunk).  A is an NHWC activation tile fetched by ONE 4-D TMA box
// {32 ch, TW, TH, TN} whose start coordinate carries the tap offset -- out-of-image pixels are
// zero-filled by the TMA unit, so padding costs nothing and no im2col buffer exists.  B is the
// packed weight slab [tap][Cout][Cin] (K-major), a 3-D TMA box {32, BN, 1}.  Both land in
// 128-byte-swizzled shared memory and feed tcgen05.mma.kind::tf32 straight from descriptors.
//
// Warp roles (576 threads, persistent over output tiles, static round-robin schedule):
//   warp 0 : TMA producer (one lane)            -- full/empty mbarrier ring, STAGES deep
//   warp 1 : TMEM allocator + MMA issuer (one lane), 2 accumulator buffers in TMEM
//   warps 2-17: epilogue -- tcgen05.ld 32 lanes x 32 columns, +bias, (+GELU), (+residual), (x GELU'),
//              128-byte row stores; 4 warps per TMEM lane quarter, each taking every 4th column chunk
// so the epilogue of tile i overlaps the mainloop of tile i+1.
#include "tc_common.cuh"
#include "conv_epilogue.cuh"

namespace {

constexpr int kEpiWarps = 16;                // 4 TMEM lane quarters x 4 column groups

constexpr int kTileM = 128;
constexpr int kChunkK = 32;                 // fp32 elements = 128 bytes = one swizzle row
constexpr int kABytes = kTileM * 128;       // 16 KiB per stage

struct TcParams {
  int B, Hg, Wg;
  int TW, TH, TN;
  int tiles_x, tiles_y, tiles_n, tiles_co, total_tiles;
  int sy, sx;
  int Cout;
  int nsrc;
  int ntaps[2];
  int kchunks[2];
  int wpb[2];
  int8_t dy[2][CD_MAX_TAPS];
  int8_t dx[2][CD_MAX_TAPS];
  float* out; int out_ld; int Ho, Wo; int oys, oxs, oy0, ox0;
  const float* bias;
  const float* resid; int resid_ld;
  int act; int round_tf32;
  float* out2; int out2_ld;
  const float* aux; int aux_ld;
  int vec8;                       // every epilogue pointer 32-byte aligned, every row stride a multiple of 8 floats
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// K-major, SWIZZLE_128B shared-memory operand descriptor (rows of 128 B, 8-row groups 1024 B apart)
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

// kind::f16 (FP16 operands, fp32 accumulate): same descriptors, K = 16 elements = 32 bytes per instruction
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

// STG: line-coalesced epilogue through a per-warp shared-memory tile (conv_epilogue.cuh; opt-in, fewer mainloop stages)
// F16: operand-format probe (cd_conv_fwd_f16_probe): sources and packed weights are FP16 arrays, one 128-byte swizzle row = 64 channels
// EPI: epilogue warps (16, or 8 for the two-CTAs-per-SM configuration: the single MMA-issuing thread needs ~8 clk per SASS
// instruction with nobody to hide its latencies, ~340 clk per K chunk against 128 clk of MMAs at N = 64 -- ncu source page,
// profiles/ncu_conv_fwd_r02b_*.txt; two co-resident CTAs interleave two such instruction streams on one tensor core)
template <int BN, int STAGES, bool STG = false, bool F16 = false, int EPI = 16>
__global__ void __launch_bounds__(64 + 32 * EPI, EPI == 16 ? 1 : 2)
conv_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               const __grid_constant__ CUtensorMap mapB0, const __grid_constant__ CUtensorMap mapB1,
               const TcParams p) {
  constexpr int kBBytes = BN * 128;
  constexpr int kStageBytes = kABytes + kBBytes;
  constexpr uint32_t kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  // instruction descriptor: D=f32, A=B=tf32, both K-major, N=BN, M=128
  constexpr uint32_t kFmt = F16 ? 0u : 2u;           // a_format / b_format: 0 = F16 (kind::f16), 2 = TF32 (kind::tf32)
  constexpr uint32_t kIdesc = (1u << 4) | (kFmt << 7) | (kFmt << 10) | (uint32_t(BN >> 3) << 17) | (uint32_t(kTileM >> 4) << 24);
  constexpr int kChunkElems = F16 ? 64 : kChunkK;   // elements per 128-byte row

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;                       // 1024-byte aligned (SWIZZLE_128B atoms)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  static_assert((2 * STAGES + 4) * 8 + 4 <= 256, "barrier block is 256 bytes");
  float* epi_stage = reinterpret_cast<float*>(smem + STAGES * kStageBytes + 256);     // STG: EPI x 32 x 36 floats

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&tmem_full[0], 1); mbar_init(&tmem_full[1], 1);
    mbar_init(&tmem_empty[0], EPI); mbar_init(&tmem_empty[1], EPI);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int kiters = p.ntaps[0] * p.kchunks[0] + (p.nsrc > 1 ? p.ntaps[1] * p.kchunks[1] : 0);

  if (warp == 0) {
    // ===================== TMA producer: ONE elected thread runs the whole schedule =====================
    if (elect_one()) {
      uint32_t stage = 0, ph = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int co_t = tile % p.tiles_co;
        int mt = tile / p.tiles_co;
        const int tx = mt % p.tiles_x; mt /= p.tiles_x;
        const int ty = mt % p.tiles_y;
        const int tn = mt / p.tiles_y;
        const int x0 = tx * p.TW * p.sx, y0 = ty * p.TH * p.sy, n0 = tn * p.TN, co0 = co_t * BN;
        for (int s = 0; s < p.nsrc; ++s) {
          const CUtensorMap* mA = s ? &mapA1 : &mapA0;
          const CUtensorMap* mB = s ? &mapB1 : &mapB0;
          const int wbase = p.wpb[s] ? n0 * p.ntaps[s] : 0;
          for (int tap = 0; tap < p.ntaps[s]; ++tap) {
            const int xin = x0 + p.dx[s][tap], yin = y0 + p.dy[s][tap];
            for (int kc = 0; kc < p.kchunks[s]; ++kc) {
              mbar_wait(&empty_bar[stage], ph ^ 1u);
              mbar_expect_tx(&full_bar[stage], kStageBytes);
              const uint32_t sa = smem_u32(smem + stage * kStageBytes);
              tma_load_4d(sa, mA, &full_bar[stage], kc * kChunkElems, xin, yin, n0);
              tma_load_3d(sa + kABytes, mB, &full_bar[stage], kc * kChunkElems, co0, wbase + tap);
              if (++stage == STAGES) { stage = 0; ph ^= 1u; }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer: ONE elected thread runs the whole loop =====================
    // tools/micro/umma_rate.cu (profiles/umma_rate_r02.txt): barrier polls, commits and MMAs go through one in-order queue.  A
    // loop of the form { wait full[s]; 4 MMAs; commit } leaves the tensor core idle from the moment the last MMA is handed over
    // until the next poll has returned and the next descriptors are built (77 clk per 128x64x8 MMA instead of 50; the per-chunk
    // elect / reconverge of a whole-warp loop adds to it).  Here the barrier of the NEXT chunk is polled in the MIDDLE of this
    // chunk's MMAs: its result comes back while the first two MMAs execute and the last two are queued behind it, so the tensor
    // core always has work while the thread commits, advances the ring and builds descriptors (53 clk / MMA in the same test).
    // The poll is a single non-blocking test: when the data is late (L2-bound layers) the chunk is finished and committed first.
    if (elect_one()) {
      uint32_t stage = 0, ph = 0, tcount = 0;
      const uint64_t desc0 = make_kmajor_sw128_desc(smem_u32(smem));
      if (static_cast<int>(blockIdx.x) < p.total_tiles) {
        mbar_wait(&tmem_empty[0], 1u);
        mbar_wait(&full_bar[0], 0u);
      }
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
        const uint32_t acc = tcount & 1u;
        const uint32_t tmem_d = tmem_base + acc * BN;
        const bool more_tiles = tile + static_cast<int>(gridDim.x) < p.total_tiles;
        for (int k = 0; k < kiters; ++k) {
          tc_fence_after();
          // descriptors of this stage = descriptor of stage 0 + a multiple of the stage size (the 14-bit address field cannot
          // carry: shared memory ends below 256 KB)
          const uint64_t da = desc0 + static_cast<uint64_t>(stage * uint32_t(kStageBytes >> 4));
          const uint64_t db = da + uint64_t(kABytes >> 4);
          uint32_t sn = stage + 1, phn = ph;
          if (sn == STAGES) { sn = 0; phn ^= 1u; }
          const bool last = k == kiters - 1;
          const bool more = !last || more_tiles;
          uint32_t ready = 0;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {          // 4 x (K = 8 tf32 = 32 bytes) per 128-byte row
            if (kk == 2 && more) {                  // ONE poll of the next chunk's barrier (of this tile or of the next one)
              ready = mbar_test(&full_bar[sn], phn);
              if (last) ready &= mbar_test(&tmem_empty[acc ^ 1u], (((tcount + 1) >> 1) & 1u) ^ 1u);   // next tile's accumulator drained
            }
            if constexpr (F16) mma_f16(tmem_d, da + uint64_t(kk * 2), db + uint64_t(kk * 2), kIdesc, (k | kk) != 0 ? 1u : 0u);
            else mma_tf32(tmem_d, da + uint64_t(kk * 2), db + uint64_t(kk * 2), kIdesc, (k | kk) != 0 ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);             // frees the smem slot once these MMAs retire
          if (more && !ready) {                     // data not there yet: nothing to overlap, wait in the open
            mbar_wait(&full_bar[sn], phn);
            if (last) mbar_wait(&tmem_empty[acc ^ 1u], (((tcount + 1) >> 1) & 1u) ^ 1u);
          }
          stage = sn; ph = phn;
        }
        tc_commit(&tmem_full[acc]);                 // accumulator complete -> epilogue
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..17) =====================
    // The epilogue (TMEM -> regs -> bias/GELU/residual -> global) costs about as many issue slots per tile as the
    // MMAs of a 3x3 tile take cycles, so it is spread over 16 warps: warp w serves TMEM lane quarter w % 4 and the
    // 32-column chunks cg, cg + 4, ... (ncu r01: 4 epilogue warps left the tensor pipe 15-30 % busy).
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int cg = (warp - 2) >> 2;                  // column group 0..3
    const int m = q * 32 + lane;                     // accumulator row == pixel within the tile
    const int xx = m % p.TW;
    const int yy = (m / p.TW) % p.TH;
    const int nn = m / (p.TW * p.TH);
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
      const int co_t = tile % p.tiles_co;
      int mt = tile / p.tiles_co;
      const int tx = mt % p.tiles_x; mt /= p.tiles_x;
      const int ty = mt % p.tiles_y;
      const int tn = mt / p.tiles_y;
      const int gx = tx * p.TW + xx, gy = ty * p.TH + yy, b = tn * p.TN + nn;
      const int co0 = co_t * BN;
      const bool valid = b < p.B;
      const long long pix = (static_cast<long long>(b) * p.Ho + (gy * p.oys + p.oy0)) * p.Wo + (gx * p.oxs + p.ox0);
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
        if constexpr (STG) {
          if (co0 + c + 32 <= p.Cout) {               // warp-uniform: full 32-column chunk
            cd_epilogue_staged32(r, epi_stage + (warp - 2) * kEpiStageFloats, lane, pix, valid, co0 + c, p);
            continue;
          }
        }
        if (valid && co0 + c < p.Cout) {
          const int nvalid = min(32, p.Cout - (co0 + c));
          if (nvalid == 32 && p.vec8) {
            // 256-bit accesses (STG.E.256 / LDG.E.256): one full 32-byte sector per lane and instruction -- with 128-bit stores
            // every sector of the output row is written in two half-sector pieces by different instructions
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
          } else if (nvalid == 32) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                     __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
              if (p.bias) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + co0 + c + j));
                v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
              }
              if (rrow) {
                const float4 rr = *reinterpret_cast<const float4*>(rrow + co0 + c + j);
                v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
              }
              if (o2row) *reinterpret_cast<float4*>(o2row + co0 + c + j) = v;
              if (p.act == CD_ACT_GELU) { v.x = cd_gelu(v.x); v.y = cd_gelu(v.y); v.z = cd_gelu(v.z); v.w = cd_gelu(v.w); }
              else if (p.act == CD_ACT_GELU_BWD) {
                const float4 a = *reinterpret_cast<const float4*>(arow + co0 + c + j);
                v.x *= cd_gelu_grad(a.x); v.y *= cd_gelu_grad(a.y); v.z *= cd_gelu_grad(a.z); v.w *= cd_gelu_grad(a.w);
              }
              if (p.round_tf32) { v.x = cd_round_tf32(v.x); v.y = cd_round_tf32(v.y); v.z = cd_round_tf32(v.z); v.w = cd_round_tf32(v.w); }
              *reinterpret_cast<float4*>(orow + co0 + c + j) = v;
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

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------

int g_num_sms = 0;
int g_tf32_map_dtype = 1;   // 1: TFLOAT32 tensor maps -- the TMA unit rounds fp32->tf32 (RN) on load (measured: profiles/tf32_probe_r01.txt); 0: FLOAT32 (MMA truncates)


template <int BN, int STAGES, bool STG = false, bool F16 = false, int EPI = 16>
int launch(const CUtensorMap* maps, const TcParams& p, cudaStream_t st) {
  constexpr size_t smem = size_t(STAGES) * (kABytes + BN * 128) + 1024 + 256 + (STG ? sizeof(float) * EPI * kEpiStageFloats : 0);
  static_assert(smem <= 232448, "dynamic shared memory of one CTA (227 KB)");
  static_assert(EPI == 16 || 2 * (smem + 1024) <= 233472, "two CTAs per SM must fit the 228 KB of shared memory");
  static bool attr_done = false;
  if (!attr_done) {
    CD_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, STAGES, STG, F16, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  const int slots = g_num_sms * (EPI == 16 ? 1 : 2);
  const int grid = p.total_tiles < slots ? p.total_tiles : slots;
  conv_tc_kernel<BN, STAGES, STG, F16, EPI><<<grid, 64 + 32 * EPI, smem, st>>>(maps[0], maps[1], maps[2], maps[3], p);
  CD_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" int cd_conv_tc_set_tf32_maps(int enable) { g_tf32_map_dtype = enable ? 1 : 0; return 0; }

int cd_conv_fwd_tc2(const CdConvDesc* d, cudaStream_t st, int BN);   // conv_tc2.cu: SM-pair (cta_group::2) variant
static int g_use_2cta = 1;
extern "C" int cd_conv_tc_set_2cta(int mode) { g_use_2cta = mode; return 0; }   // 0 off, 1 where the cost model prefers it, 2 wherever eligible
// halo-tile kernels for stride-1 convolutions with taps in [-1, 1]^2, bit mask: 1 = conv_tc3.cu (16 x 8 patch) wherever eligible,
// 2 = conv_tc4.cu (16 x 16 patch, two accumulators per weight tile) for Cout <= 128 where its tile count model says so (default),
// 4 (with 2) = conv_tc4.cu for every eligible problem (tests)
int cd_conv_fwd_tc3(const CdConvDesc* d, cudaStream_t st);
int cd_conv_fwd_tc4(const CdConvDesc* d, cudaStream_t st);
void cd_conv_tc4_force(int on);
static int g_use_halo = 2;
extern "C" int cd_conv_tc_set_halo(int enable) { g_use_halo = enable; cd_conv_tc4_force((enable & 4) ? 1 : 0); return 0; }
// two CTAs per SM (8 epilogue warps, half the stages each) for the 1-CTA kernels with N <= 128: bit mask of N tiles (128 | 64)
static int g_ctas2 = 0;
extern "C" int cd_conv_tc_set_two_ctas(int mask) { g_ctas2 = mask & (128 | 64); return 0; }
int cd_conv_tc_two_ctas_mask() { return g_ctas2; }
// narrower pair tiles: bit mask of the N tiles below 256 (128 | 64) that go to the SM-pair kernel when the problem is eligible
static int g_2cta_bn = 128;     // measured (profiles/conv_shapes_r02*.txt): the pair kernel wins at N = 128 (+3..6 %), loses at N = 64
extern "C" int cd_conv_tc_set_2cta_bn(int mask) { g_2cta_bn = mask & (128 | 64); return 0; }
// line-coalesced epilogue (conv_epilogue.cuh): 0 = off (default: not validated on a B200 yet), 1 = for the short-K launches that
// are bound by their output stores (at most kStagedMaxKIters 32-channel K chunks per tile), 2 = for every launch (tests),
// 3 = up to kStagedMidKIters chunks (where the row epilogue of a 128-pixel tile, ~6 us per 64 KB, still outlasts the tile's MMAs)
static int g_epi_staged = 0;
constexpr int kStagedMaxKIters = 16;
constexpr int kStagedMidKIters = 48;
extern "C" int cd_conv_tc_set_staged_epilogue(int mode) { g_epi_staged = mode; return 0; }

// Tile-shape choice.  Cost model: waves over the SMs x columns per tile / relative MMA issue rate of that tile shape.  The
// rates are the measured TFLOP/s of the long-K convolutions of the config-3 network (profiles/conv_shapes_2cta_r01.txt):
// SM-pair 256x256 ~ 660-690, 128x256 ~ 610-645, 128x128 ~ 490-505, 128x64 ~ 265.
static double tc_cost(long long m_tiles, int Cout, int BN, bool pair, int sms) {
  if (pair) {
    const long long tiles = ((m_tiles + 1) / 2) * (Cout / 256), slots = sms / 2;
    return double((tiles + slots - 1) / slots) * 256.0;
  }
  const long long tiles = m_tiles * ((Cout + BN - 1) / BN);
  const double rate = BN == 256 ? 0.92 : (BN == 128 ? 0.75 : 0.40);
  return double((tiles + sms - 1) / sms) * BN / rate;
}

static int conv_fwd_tc_impl(const CdConvDesc* d, cudaStream_t st, bool f16);
int cd_conv_fwd_tc(const CdConvDesc* d, cudaStream_t st) { return conv_fwd_tc_impl(d, st, false); }

// Operand-format probe (tools/conv_f16_probe.py; not used by the engine): the same convolution with FP16 sources and FP16 packed
// weights (src / w point to __half arrays, ld and the weight strides count elements), fp32 accumulate and the fp32 epilogue.
extern "C" int cd_conv_fwd_f16_probe(const CdConvDesc* d, void* stream) {
  CD_REQUIRE(d != nullptr && d->nsrc >= 1 && d->nsrc <= 2, "cd_conv_fwd_f16_probe: bad descriptor");
  return conv_fwd_tc_impl(d, static_cast<cudaStream_t>(stream), true);
}

static int conv_fwd_tc_impl(const CdConvDesc* d, cudaStream_t st, bool f16) {
  if (!f16 && (g_use_halo & 2) && g_tf32_map_dtype && g_epi_staged == 0) {
    const int r4 = cd_conv_fwd_tc4(d, st);
    if (r4 <= 0) return r4;
  }
  if (!f16 && (g_use_halo & 1) && g_tf32_map_dtype && g_epi_staged == 0) {
    const int r3 = cd_conv_fwd_tc3(d, st);
    if (r3 <= 0) return r3;                                   // 1 = not eligible: the per-tap kernels below
  }
  const int chunk_elems = f16 ? 64 : kChunkK;
  const int esz = f16 ? 2 : 4;
  EncodeTiledFn enc = get_encode();
  CD_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  if (!g_num_sms) {
    int dev = 0; CD_CUDA(cudaGetDevice(&dev));
    CD_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  CD_REQUIRE(d->nsrc >= 1 && d->nsrc <= 2, "conv_tc: nsrc must be 1 or 2");
  TcParams p{};
  p.B = d->B; p.Hg = d->Hg; p.Wg = d->Wg; p.sy = d->sy; p.sx = d->sx; p.Cout = d->Cout; p.nsrc = d->nsrc;
  // tile geometry: 128 GEMM rows = TW x TH x TN pixels
  if (d->Wg >= 128) {
    CD_REQUIRE(d->Wg % 128 == 0, "conv_tc: Wg >= 128 must be a multiple of 128 (got %d)", d->Wg);
    p.TW = 128; p.TH = 1; p.TN = 1;
  } else {
    CD_REQUIRE(is_pow2(d->Wg), "conv_tc: Wg < 128 must be a power of two (got %d)", d->Wg);
    p.TW = d->Wg;
    int th = 128 / p.TW;
    if (th > d->Hg) th = d->Hg;
    CD_REQUIRE(is_pow2(th) && d->Hg % th == 0, "conv_tc: unsupported Hg %d for Wg %d", d->Hg, d->Wg);
    p.TH = th; p.TN = 128 / (p.TW * p.TH);
  }
  CD_REQUIRE(p.TW * d->sx <= 256 && p.TH * d->sy <= 256, "conv_tc: strided box too large");
  p.tiles_x = d->Wg / p.TW; p.tiles_y = d->Hg / p.TH; p.tiles_n = cd_cdiv(d->B, p.TN);
  int BN = (d->Cout % 256 == 0) ? 256 : (d->Cout > 64 ? 128 : 64);
  int kiters_host = 0;
  for (int s = 0; s < d->nsrc; ++s) kiters_host += d->s[s].ntaps * (d->s[s].C / chunk_elems);
  const bool staged = !f16 &&
                      ( g_epi_staged == 2 || (g_epi_staged == 1 && kiters_host <= kStagedMaxKIters) ||
                      (g_epi_staged == 3 && kiters_host <= kStagedMidKIters));
  if (BN == 256 && staged) {
    // short K: the narrower tile only when it saves whole waves (same rule as below), never the SM-pair kernel
    const long long mt = static_cast<long long>(p.tiles_x) * p.tiles_y * p.tiles_n;
    if (tc_cost(mt, d->Cout, 128, false, g_num_sms) < tc_cost(mt, d->Cout, 256, false, g_num_sms)) BN = 128;
  } else if (BN == 256) {
    // small spatial sizes (16^2, 32^2) give few M tiles: take the narrower N tile only when it saves whole waves
    const long long mt = static_cast<long long>(p.tiles_x) * p.tiles_y * p.tiles_n;
    double best = tc_cost(mt, d->Cout, 256, false, g_num_sms);
    if (tc_cost(mt, d->Cout, 128, false, g_num_sms) < best) { BN = 128; best = tc_cost(mt, d->Cout, 128, false, g_num_sms); }
    if (!f16 && g_use_2cta && g_tf32_map_dtype && mt >= 2 && (g_use_2cta == 2 || tc_cost(mt, d->Cout, 256, true, g_num_sms) < best)) {
      const int r2 = cd_conv_fwd_tc2(d, st, 256);
      if (r2 <= 0) return r2;                                 // 1 = not eligible: stay on the 1-CTA kernel
    }
  } else if (!f16 && !staged && g_use_2cta && g_tf32_map_dtype && (g_2cta_bn & BN) && d->Cout % BN == 0 &&
             (g_use_2cta == 2 || static_cast<long long>(p.tiles_x) * p.tiles_y * p.tiles_n >= 2 * g_num_sms)) {
    const int r2 = cd_conv_fwd_tc2(d, st, BN);                // narrow pair tile: only with at least one pair tile per SM pair
    if (r2 <= 0) return r2;
  }
  p.tiles_co = cd_cdiv(d->Cout, BN);
  p.total_tiles = p.tiles_x * p.tiles_y * p.tiles_n * p.tiles_co;
  p.out = d->out; p.out_ld = d->out_ld; p.Ho = d->Ho; p.Wo = d->Wo;
  p.oys = d->oys; p.oxs = d->oxs; p.oy0 = d->oy0; p.ox0 = d->ox0;
  p.bias = d->bias; p.resid = d->resid; p.resid_ld = d->resid_ld; p.act = d->act; p.round_tf32 = d->round_tf32;
  p.out2 = d->out2; p.out2_ld = d->out2_ld; p.aux = d->aux; p.aux_ld = d->aux_ld;
  {
    auto ok8 = [](const void* ptr, int ld) { return ptr == nullptr || ((reinterpret_cast<uintptr_t>(ptr) & 31) == 0 && ld % 8 == 0); };
    p.vec8 = ok8(d->out, d->out_ld) && ok8(d->out2, d->out2_ld) && ok8(d->resid, d->resid_ld) && ok8(d->aux, d->aux_ld) && ok8(d->bias, 8);
  }
  CD_REQUIRE(d->act != CD_ACT_GELU_BWD || (d->aux && (reinterpret_cast<uintptr_t>(d->aux) & 15) == 0 && d->aux_ld % 4 == 0), "conv_tc: GELU_BWD needs an aligned aux");
  CD_REQUIRE((reinterpret_cast<uintptr_t>(d->out) & 15) == 0 && d->out_ld % 4 == 0, "conv_tc: out must be 16B aligned");
  CD_REQUIRE(!d->resid || ((reinterpret_cast<uintptr_t>(d->resid) & 15) == 0 && d->resid_ld % 4 == 0), "conv_tc: resid alignment");
  CD_REQUIRE(!d->out2 || ((reinterpret_cast<uintptr_t>(d->out2) & 15) == 0 && d->out2_ld % 4 == 0), "conv_tc: out2 alignment");
  CD_REQUIRE(!d->bias || (reinterpret_cast<uintptr_t>(d->bias) & 15) == 0, "conv_tc: bias alignment");

  CUtensorMap maps[4];
  const CUtensorMapDataType dt = f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : (g_tf32_map_dtype ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
  for (int s = 0; s < 2; ++s) {
    const CdConvSrc& cs = d->s[s < d->nsrc ? s : 0];
    CD_REQUIRE(cs.C % chunk_elems == 0 && cs.C > 0, "conv_tc: source channels %d not a multiple of %d", cs.C, chunk_elems);
    CD_REQUIRE(cs.ntaps >= 1 && cs.ntaps <= CD_MAX_TAPS, "conv_tc: bad ntaps");
    CD_REQUIRE((reinterpret_cast<uintptr_t>(cs.src) & 15) == 0 && (cs.ld * esz) % 16 == 0, "conv_tc: src must be 16B aligned");
    CD_REQUIRE((reinterpret_cast<uintptr_t>(cs.w) & 15) == 0, "conv_tc: weights must be 16B aligned");
    CD_REQUIRE(!cs.w_per_batch || p.TN == 1, "conv_tc: per-batch weights need >=128 pixels per image");
    p.ntaps[s] = cs.ntaps; p.kchunks[s] = cs.C / chunk_elems; p.wpb[s] = cs.w_per_batch;
    for (int t = 0; t < cs.ntaps; ++t) { p.dy[s][t] = (int8_t)cs.dy[t]; p.dx[s][t] = (int8_t)cs.dx[t]; }
    {   // A: NHWC activations, dims {C, W, H, N}
      cuuint64_t dims[4] = {(cuuint64_t)cs.C, (cuuint64_t)cs.W, (cuuint64_t)cs.H, (cuuint64_t)d->B};
      cuuint64_t strides[3] = {(cuuint64_t)cs.ld * esz, (cuuint64_t)cs.ld * esz * cs.W, (cuuint64_t)cs.ld * esz * cs.W * cs.H};
      cuuint32_t box[4] = {(cuuint32_t)chunk_elems, (cuuint32_t)(p.TW * d->sx), (cuuint32_t)(p.TH * d->sy), (cuuint32_t)p.TN};
      cuuint32_t estr[4] = {1, (cuuint32_t)d->sx, (cuuint32_t)d->sy, 1};
      CUresult r = enc(&maps[s], dt, 4, const_cast<float*>(cs.src), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      CD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(A%d) failed: %d", s, (int)r);
    }
    {   // B: packed weights, dims {Cin, Cout, ntaps * (per-batch ? B : 1)}
      const int BNl = BN;
      cuuint64_t dims[3] = {(cuuint64_t)cs.C, (cuuint64_t)d->Cout, (cuuint64_t)cs.ntaps * (cs.w_per_batch ? d->B : 1)};
      cuuint64_t strides[2] = {(cuuint64_t)cs.C * esz, (cuuint64_t)cs.C * esz * d->Cout};
      cuuint32_t box[3] = {(cuuint32_t)chunk_elems, (cuuint32_t)BNl, 1};
      cuuint32_t estr[3] = {1, 1, 1};
      CUresult r = enc(&maps[2 + s], dt, 3, const_cast<float*>(cs.w), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      CD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(B%d) failed: %d", s, (int)r);
    }
  }
  if (f16) {
    if (BN == 256) return launch<256, 4, false, true>(maps, p, st);
    if (BN == 128) return launch<128, 6, false, true>(maps, p, st);
    return launch<64, 8, false, true>(maps, p, st);
  }
  if (staged) {          // 72 KB of epilogue staging: 3 / 4 / 6 mainloop stages instead of 4 / 6 / 8
    if (BN == 256) return launch<256, 3, true>(maps, p, st);
    if (BN == 128) return launch<128, 4, true>(maps, p, st);
    return launch<64, 6, true>(maps, p, st);
  }
  if (BN == 256) return launch<256, 4>(maps, p, st);
  if (BN == 128) return (g_ctas2 & 128) ? launch<128, 3, false, false, 8>(maps, p, st) : launch<128, 6>(maps, p, st);
  return (g_ctas2 & 64) ? launch<64, 4, false, false, 8>(maps, p, st) : launch<64, 8>(maps, p, st);
}
