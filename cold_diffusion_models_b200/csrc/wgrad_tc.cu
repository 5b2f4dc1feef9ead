// tcgen05 weight-gradient of the tap-list convolution (replaces cuDNN wgrad behind loss.backward() for the
// dense convolutions of the reference Unet, DB:149-154,173-174,105-109):
//
//   dW[tap][co][ci] += sum_{pixels p} dY[p][co] * X[p (+) tap][ci]
//
// as a GEMM whose K dimension is the PIXEL axis: D[128 co x BN ci] (fp32, TMEM) += A^T B with both operands
// "MN-major" in shared memory -- NHWC rows are exactly that: one 128-byte swizzled row per pixel holding 32
// channels.  dY tiles ({32 co, CW px, R rows} TMA boxes) and X tiles land in SWIZZLE_128B smem and are
// consumed by tcgen05.mma.kind::tf32 eight pixels (K = 8) at a time.  One CTA owns (co tile, ci tile, tap
// group, pixel split) and keeps one TMEM accumulator per tap of its group (<= 4 x 128 columns); with
// `halo` mode the dx = -1/0/+1 taps of a 3x3 kernel share ONE X tile that carries a one-pixel halo and are
// addressed by shifting the descriptor start by whole 128-byte rows.  Partial sums of the pixel splits are
// combined with vector red.global.add.f32.
#include "tc_common.cuh"

namespace {

constexpr int kThreads = 192;
constexpr int kMaxTaps = 4;      // taps (accumulators) per CTA
constexpr int kMaxGroups = 4;

struct WgParams {
  int B, Hg, Wg;
  int CW, R, KR;                  // pixel chunk: CW x R = KR K-rows of A per stage
  int chunks_x, chunks_y, total_chunks, chunks_per_split, splits;
  int Cout, Cin;
  int tiles_co, tiles_ci;
  int ngroups;
  int ntaps[kMaxGroups];
  int nloads[kMaxGroups];
  int tap_index[kMaxGroups][kMaxTaps];
  int tap_load[kMaxGroups][kMaxTaps];
  int tap_shift[kMaxGroups][kMaxTaps];    // K-row shift inside the (halo) X tile
  int load_dy[kMaxGroups][kMaxTaps], load_dx[kMaxGroups][kMaxTaps];
  int halo;                       // extra pixels per row in the X box (0 or 2)
  int sy, sx;                     // X coordinate = g*s + d
  int oys, oxs, oy0, ox0;         // dY coordinate = g*os + o0
  int base_offset_mode;           // descriptor base_offset for row-shifted starts: 0 = none, 1 = (addr >> 7) & 7
  int per_batch, chunks_per_img, splits_per_img;   // per-batch weights: splits never cross images
  long long dw_batch_stride;
  float* dw;
  float* db;                      // BIAS kernels only: db[co] += sum over all pixels of dY (nullptr: not fused)
};

// MN-major operand for 32-bit (tf32) data: the only layout UMMA accepts is SWIZZLE_128B_BASE32B -- 128-byte rows (32
// channels of one pixel), 32-byte swizzle atoms, pattern period 4 rows (512 B) -- which is what a TMA tensor map with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B writes.  32-channel chunks are `lbo` bytes apart, 4-pixel groups 512 B apart.
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t base_off) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(base_off & 7u) << 49;
  d |= static_cast<uint64_t>(1) << 61;                          // SWIZZLE_128B_BASE32B
  return d;
}
// same instruction without the "memory" clobber: volatile asm statements keep their order among themselves (barrier waits,
// fences, commits), and ordinary loads of the constant descriptor table may then be scheduled across the MMAs
__device__ __forceinline__ void mma_tf32_nomem(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
               :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// BIAS = true (opt-in, cd_wgrad_tc_set_bias_fusion): the CTAs of the first ci tile and first tap group also accumulate the bias
// gradient db[co] = sum_pixels dY[p][co] with one extra 128 x 32 x 8 MMA per k-step against a tile of ones (TMEM columns after
// the tap accumulators), replacing the separate column-sum pass over dY.  BIAS = false is the production kernel.
template <int BN, bool BIAS>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap mapDY, const __grid_constant__ CUtensorMap mapX,
                const WgParams p, int stages, int a_bytes, int b_bytes, int b_tx_bytes) {
  constexpr int NCH = BN / 32;                    // ci chunks of the B operand
  // instruction descriptor: D=f32, A=B=tf32, A and B MN-major, N=BN, M=128
  constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                              (uint32_t(BN >> 3) << 17) | (uint32_t(128 >> 4) << 24);
  constexpr uint32_t kIdescBias = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                                  (uint32_t(32 >> 3) << 17) | (uint32_t(128 >> 4) << 24);
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint4 mma_tab[kMaxTaps * 8 + 8 + 1];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int g = blockIdx.z;
  const int nl = p.nloads[g], nt = p.ntaps[g];
  const int stage_bytes = a_bytes + nl * b_bytes;
  float* ones = reinterpret_cast<float*>(smem + stages * stage_bytes);          // BIAS: 8 pixels x 32 channels of 1.0f (1 KB)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes + (BIAS ? 1024 : 0));
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + stages;
  uint64_t* done_bar = bars + 2 * stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * stages + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = 512;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (BIAS) {
    if (warp >= 2) for (int i = threadIdx.x - 64; i < 256; i += kThreads - 64) ones[i] = 1.f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the tensor-core (async) proxy
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tile = blockIdx.x;
  const int co0 = (tile % p.tiles_co) * 128, ci0 = (tile / p.tiles_co) * BN;
  const int split = blockIdx.y;
  int c_beg = split * p.chunks_per_split;
  int c_end = c_beg + p.chunks_per_split; if (c_end > p.total_chunks) c_end = p.total_chunks;
  long long dw_off = 0;
  if (p.per_batch) {
    const int n = split / p.splits_per_img, sp = split % p.splits_per_img;
    c_beg = n * p.chunks_per_img + sp * p.chunks_per_split;
    c_end = c_beg + p.chunks_per_split;
    if (c_end > (n + 1) * p.chunks_per_img) c_end = (n + 1) * p.chunks_per_img;
    dw_off = n * p.dw_batch_stride;
  }
  const int nchunks = c_end > c_beg ? c_end - c_beg : 0;
  const bool bias_cta = BIAS && p.db != nullptr && (tile / p.tiles_co) == 0 && g == 0;

  if (warp == 0) {
    // TMA producer: the whole warp walks the chunks and waits on the barriers, one elected lane issues (elect_one, tc_common.cuh)
    for (int it = 0; it < nchunks; ++it) {
      const int c = c_beg + it;
      const int cx = c % p.chunks_x;
      const int cy = (c / p.chunks_x) % p.chunks_y;
      const int n = c / (p.chunks_x * p.chunks_y);
      const int gx0 = cx * p.CW, gy0 = cy * p.R;
      const uint32_t stage = it % stages, ph = (it / stages) & 1u;
      mbar_wait(&empty_bar[stage], ph ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(&full_bar[stage], a_bytes + nl * b_tx_bytes);
        const uint32_t sa = smem_u32(smem + stage * stage_bytes);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)              // dY: 4 chunks of 32 out-channels
          tma_load_4d(sa + ch * (a_bytes / 4), &mapDY, &full_bar[stage], co0 + ch * 32, gx0 * p.oxs + p.ox0, gy0 * p.oys + p.oy0, n);
        for (int l = 0; l < nl; ++l) {
          const uint32_t sb = sa + a_bytes + l * b_bytes;
          const int xin = gx0 * p.sx + p.load_dx[g][l], yin = gy0 * p.sy + p.load_dy[g][l];
          for (int ch = 0; ch < NCH; ++ch)
            tma_load_4d(sb + ch * (b_bytes / NCH), &mapX, &full_bar[stage], ci0 + ch * 32, xin, yin, n);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---- MMA issuer.  The operand offsets of the MMAs of one chunk do not depend on the chunk: tabulate them once
    // (one entry per MMA: A / B start offsets in 16-byte units, B descriptor high word, accumulator column, accumulate flag)
    // so that the single issuing thread spends one LDS.128 + two adds per tcgen05.mma -- a 128x128x8 TF32 MMA lasts ~64 clk,
    // a 128x64x8 one ~32 clk, and descriptor arithmetic in the loop was what bounded the issue rate.
    const uint32_t a_lbo = a_bytes / 4, b_lbo = b_bytes / NCH;
    const int ksteps_row = p.CW / 8;                // MMAs per image row of the chunk
    const int per_tap = p.R * ksteps_row;
    const int ntap_mma = nt * per_tap;              // <= kMaxTaps * 8
    const int nmma = ntap_mma + (bias_cta ? per_tap : 0);
    if (lane == 0) mma_tab[nmma] = make_uint4(0, 0, 0, 0);
    if (bias_cta && lane < per_tap) {               // bias entries: same A rows, B = the ones tile (flag bit 17), columns after the taps
      const int r = lane / ksteps_row, j = lane % ksteps_row;
      mma_tab[ntap_mma + lane] = make_uint4(((r * p.CW + j * 8) * 128) >> 4, 0, 0,
                                            static_cast<uint32_t>(nt * BN) | ((r | j) != 0 ? 0x10000u : 0u) | 0x20000u);
    }
    if (lane < ntap_mma) {
      const int t = lane / per_tap, r = (lane % per_tap) / ksteps_row, j = lane % ksteps_row;
      const uint32_t arow = r * p.CW + j * 8;
      const uint32_t brow = r * (p.CW + p.halo) + j * 8 + p.tap_shift[g][t];
      const uint32_t boffs = p.tap_load[g][t] * b_bytes + brow * 128;      // stage bases are 1024-byte aligned
      const uint32_t bo = p.base_offset_mode == 1 ? ((boffs >> 7) & 7u) : (p.base_offset_mode == 2 ? ((boffs >> 7) & 3u) : 0u);
      const uint64_t dbt = make_mnmajor_sw128_desc(0, b_lbo, bo);
      mma_tab[lane] = make_uint4((arow * 128) >> 4, boffs >> 4, static_cast<uint32_t>(dbt >> 32),
                                 static_cast<uint32_t>(t * BN) | ((r | j) != 0 ? 0x10000u : 0u));
    }
    __syncwarp();
    {
      const uint64_t da_t = make_mnmajor_sw128_desc(0, a_lbo, 0);
      const uint32_t da_hi = static_cast<uint32_t>(da_t >> 32), da_lo = static_cast<uint32_t>(da_t);
      const uint32_t db_lo = static_cast<uint32_t>(make_mnmajor_sw128_desc(0, b_lbo, 0));
      const uint64_t ones_desc = make_mnmajor_sw128_desc(smem_u32(ones), 1024, 0);
      for (int it = 0; it < nchunks; ++it) {
        const uint32_t stage = it % stages, ph = (it / stages) & 1u;
        mbar_wait(&full_bar[stage], ph);            // whole warp waits, one elected lane issues (elect_one, tc_common.cuh)
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa16 = smem_u32(smem + stage * stage_bytes) >> 4;     // < 2^14: no carry into the LBO field
          const uint32_t alo = da_lo + sa16, blo = db_lo + sa16 + (a_bytes >> 4);
          const uint32_t first = it == 0 ? 0u : 0x10000u;
          uint4 e = mma_tab[0];
#pragma unroll 4
          for (int i = 0; i < nmma; ++i) {
            const uint4 en = mma_tab[i + 1];          // table has one spare entry; prefetched so the LDS latency is off the issue path
            const uint64_t da = (static_cast<uint64_t>(da_hi) << 32) | (alo + e.x);
            uint64_t db = (static_cast<uint64_t>(e.z) << 32) | (blo + e.y);
            uint32_t idesc = kIdesc;
            if (BIAS && (e.w & 0x20000u)) { db = ones_desc; idesc = kIdescBias; }
            mma_tf32_nomem(tmem_base + (e.w & 0xFFFFu), da, db, idesc, (e.w | first) & 0x10000u);
            e = en;
          }
          tc_commit(&empty_bar[stage]);
        }
        __syncwarp();
      }
      if (elect_one()) tc_commit(done_bar);
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int co = co0 + q * 32 + lane;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    if (nchunks > 0) {
      for (int t = 0; t < nt; ++t) {
        float* wrow = p.dw + dw_off + (static_cast<long long>(p.tap_index[g][t]) * p.Cout + co) * p.Cin;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_base + t * BN + c + (static_cast<uint32_t>(q * 32) << 16), r);
          if (co < p.Cout && ci0 + c < p.Cin) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              red_add_v4(wrow + ci0 + c + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
          }
        }
      }
      if (BIAS && bias_cta) {                       // every column of the bias accumulator holds the same sum: take column 0
        uint32_t r[32];
        tmem_ld32(tmem_base + nt * BN + (static_cast<uint32_t>(q * 32) << 16), r);
        if (co < p.Cout) atomicAdd(p.db + co, __uint_as_float(r[0]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}


int g_wg_mode = 1;      // 0: one X tile per tap; 1 (default, verified on B200: profiles/wgrad_modes_r01.txt): dx taps share one halo
                        // tile via row-shifted descriptor starts, base_offset 0; 2/3: probes with base_offset=(addr>>7)&7 / &3 (wrong)
int g_sms = 0;

}  // namespace

extern "C" int cd_wgrad_tc_set_mode(int mode) { g_wg_mode = mode; return 0; }

// opt-in: fold the bias gradient (column sums of dY) into the tcgen05 weight gradient.  Off by default: not yet run on a B200.
static int g_wg_bias_fusion = 0;
extern "C" int cd_wgrad_tc_set_bias_fusion(int enable) { g_wg_bias_fusion = enable; return 0; }

// K-split policy.  One CTA owns (Cout tile, Cin tile, tap group, pixel range); it runs alone on its SM (the stages fill the
// shared memory) and its red.add epilogue is not overlapped, so the launch costs waves x (chunks_per_split * t_chunk + t_over).
//   0: 2 waves rounded up (first version)   1 / 2: at most 1 / 2 full waves   3: minimise the modelled cost
static int g_split_policy = 3;
static int g_split_over_clk = 12000;     // fixed cost per CTA in SM clocks (prologue + first-load latency + red.add epilogue)
extern "C" int cd_wgrad_tc_set_split(int policy, int over_clk) { g_split_policy = policy; if (over_clk > 0) g_split_over_clk = over_clk; return 0; }

static int choose_splits(int tg, int total_chunks, int max_splits, int sms, double chunk_clk) {
  if (max_splits < 1) max_splits = 1;
  int s;
  if (g_split_policy == 0) s = cd_cdiv(2 * sms, tg);
  else if (g_split_policy == 1) s = sms / tg;
  else if (g_split_policy == 2) s = 2 * sms / tg;
  else {
    double best = 1e30; s = 1;
    for (int c = 1; c <= max_splits && c * tg <= 4 * sms + tg; ++c) {
      const int cps = cd_cdiv(total_chunks, c);
      const int real = cd_cdiv(total_chunks, cps);
      const double cost = double(cd_cdiv(static_cast<long long>(tg) * real, sms)) * (cps * chunk_clk + g_split_over_clk);
      if (cost < best) { best = cost; s = c; }
    }
  }
  if (s > max_splits) s = max_splits;
  if (s < 1) s = 1;
  return s;
}

// returns 1 if the problem is not tensor-core shaped (caller falls back to the SIMT kernel), 0 on success, <0 on error
int cd_conv_wgrad_tc(const CdConvDesc* d, const float* dout, int dout_ld, float* dw, float* db, int* bias_done, cudaStream_t st) {
  if (bias_done) *bias_done = 0;
  const CdConvSrc& c = d->s[0];
  if (c.C % 32 != 0 || d->Cout % 32 != 0 || c.ld % 4 != 0 || dout_ld % 4 != 0) return 1;
  if (!is_pow2(d->Wg) || d->Wg < 8 || !is_pow2(d->Hg)) return 1;
  if ((reinterpret_cast<uintptr_t>(c.src) & 15) || (reinterpret_cast<uintptr_t>(dout) & 15) || (reinterpret_cast<uintptr_t>(dw) & 15)) return 1;
  EncodeTiledFn enc = get_encode();
  CD_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  if (!g_sms) { int dev = 0; CD_CUDA(cudaGetDevice(&dev)); CD_CUDA(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev)); }

  WgParams p{};
  p.B = d->B; p.Hg = d->Hg; p.Wg = d->Wg; p.Cout = d->Cout; p.Cin = c.C;
  p.sy = d->sy; p.sx = d->sx; p.oys = d->oys; p.oxs = d->oxs; p.oy0 = d->oy0; p.ox0 = d->ox0;
  p.dw = dw;
  const int BN = c.C >= 128 ? 128 : 64;
  // ---- tap groups ----
  const bool is3x3 = (c.ntaps == 9 && d->sy == 1 && d->sx == 1);
  bool shape3 = is3x3;
  if (is3x3) for (int t = 0; t < 9; ++t) if (c.dy[t] != t / 3 - 1 || c.dx[t] != t % 3 - 1) shape3 = false;
  const bool halo = shape3 && g_wg_mode != 0;
  p.halo = halo ? 2 : 0;
  p.base_offset_mode = (g_wg_mode == 2) ? 1 : (g_wg_mode == 3 ? 2 : 0);
  if (halo) {
    p.ngroups = 3;
    for (int gk = 0; gk < 3; ++gk) {
      p.ntaps[gk] = 3; p.nloads[gk] = 1; p.load_dy[gk][0] = gk - 1; p.load_dx[gk][0] = -1;
      for (int k = 0; k < 3; ++k) { p.tap_index[gk][k] = gk * 3 + k; p.tap_load[gk][k] = 0; p.tap_shift[gk][k] = k; }
    }
  } else {
    const int per = c.ntaps <= 4 ? c.ntaps : (c.ntaps == 9 ? 3 : 4);
    if (c.ntaps % per != 0 || c.ntaps / per > kMaxGroups) return 1;
    p.ngroups = c.ntaps / per;
    for (int gk = 0; gk < p.ngroups; ++gk) {
      p.ntaps[gk] = per; p.nloads[gk] = per;
      for (int k = 0; k < per; ++k) {
        const int t = gk * per + k;
        p.tap_index[gk][k] = t; p.tap_load[gk][k] = k; p.tap_shift[gk][k] = 0;
        p.load_dy[gk][k] = c.dy[t]; p.load_dx[gk][k] = c.dx[t];
      }
    }
  }
  const int max_taps = p.ntaps[0], max_loads = p.nloads[0];
  if (max_taps * BN > 512) return 1;
  // ---- pixel chunking: KR K-rows per stage ----
  int KR = (max_loads <= 1) ? 64 : 32;           // (KR = 32 everywhere measured 25 % slower: r01 bench_8)
  if (static_cast<long long>(d->Hg) * d->Wg < KR) KR = d->Hg * d->Wg;
  if (KR < 8) return 1;
  p.CW = d->Wg < KR ? d->Wg : KR;
  p.R = KR / p.CW;
  p.KR = KR;
  if (p.CW * d->sx > 256 || p.CW * d->oxs > 256) return 1;
  p.chunks_x = d->Wg / p.CW; p.chunks_y = d->Hg / p.R;
  p.total_chunks = d->B * p.chunks_x * p.chunks_y;
  p.tiles_co = cd_cdiv(d->Cout, 128); p.tiles_ci = cd_cdiv(c.C, BN);
  const int tiles = p.tiles_co * p.tiles_ci;
  // MMAs of one chunk: taps x KR/8 k-steps, 128 x BN x 8 each (~68 clk at BN = 128: 128 B/clk of shared-memory operand reads)
  const double chunk_clk = double(max_taps) * (KR / 8) * (BN == 128 ? 68.0 : 40.0);
  int splits;
  if (c.w_per_batch) {
    const int cpi = p.chunks_x * p.chunks_y;
    const int spi0 = choose_splits(tiles * p.ngroups * d->B, cpi, cpi, g_sms, chunk_clk);
    splits = spi0 * d->B;
  } else {
    splits = choose_splits(tiles * p.ngroups, p.total_chunks, cd_cdiv(p.total_chunks, 8), g_sms, chunk_clk);
  }
  p.chunks_per_split = cd_cdiv(p.total_chunks, splits);
  p.splits = cd_cdiv(p.total_chunks, p.chunks_per_split);
  if (c.w_per_batch) {
    p.per_batch = 1;
    p.chunks_per_img = p.chunks_x * p.chunks_y;
    int spi = cd_cdiv(splits, d->B); if (spi < 1) spi = 1; if (spi > p.chunks_per_img) spi = p.chunks_per_img;
    p.chunks_per_split = cd_cdiv(p.chunks_per_img, spi);
    p.splits_per_img = cd_cdiv(p.chunks_per_img, p.chunks_per_split);
    p.splits = p.splits_per_img * d->B;
    p.dw_batch_stride = static_cast<long long>(c.ntaps) * d->Cout * c.C;
  }
  const int a_bytes = 4 * KR * 128;
  const int b_rows = p.R * (p.CW + p.halo);
  const int b_rows_pad = (b_rows + 7) / 8 * 8;                 // keep every chunk base 1024-byte aligned
  const int b_bytes = (BN / 32) * b_rows_pad * 128;
  const int b_tx = (BN / 32) * b_rows * 128;                   // bytes the TMA unit actually writes per X tile
  const int stage_bytes = a_bytes + max_loads * b_bytes;
  int stages = (224 * 1024) / stage_bytes; if (stages > 6) stages = 6;   // 3 x 68 KB halo stages fit the 227 KB carve-out
  if (stages < 2) return 1;
  // bias fusion: dense output grid, one weight set, room for 32 more TMEM columns after the tap accumulators
  const bool fuse_bias = g_wg_bias_fusion && db != nullptr && !c.w_per_batch && d->oys == 1 && d->oxs == 1 && max_taps * BN + 32 <= 512;
  p.db = fuse_bias ? db : nullptr;
  const size_t smem = size_t(stages) * stage_bytes + 1024 + 256 + (fuse_bias ? 1024 : 0);

  CUtensorMap mapDY, mapX;
  const CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_TFLOAT32;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->Cout, (cuuint64_t)d->Wo, (cuuint64_t)d->Ho, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)dout_ld * 4, (cuuint64_t)dout_ld * 4 * d->Wo, (cuuint64_t)dout_ld * 4 * d->Wo * d->Ho};
    cuuint32_t box[4] = {32, (cuuint32_t)(p.CW * d->oxs), (cuuint32_t)(p.R * d->oys), 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)d->oxs, (cuuint32_t)d->oys, 1};
    CUresult r = enc(&mapDY, dt, 4, const_cast<float*>(dout), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(dY) failed: %d", (int)r);
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)c.C, (cuuint64_t)c.W, (cuuint64_t)c.H, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)c.ld * 4, (cuuint64_t)c.ld * 4 * c.W, (cuuint64_t)c.ld * 4 * c.W * c.H};
    cuuint32_t box[4] = {32, (cuuint32_t)((p.CW + p.halo) * d->sx), (cuuint32_t)(p.R * d->sy), 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)d->sx, (cuuint32_t)d->sy, 1};
    CUresult r = enc(&mapX, dt, 4, const_cast<float*>(c.src), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(X) failed: %d", (int)r);
  }
  dim3 grid(tiles, p.splits, p.ngroups);
  if (fuse_bias) {
    if (BN == 128) {
      CD_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      wgrad_tc_kernel<128, true><<<grid, kThreads, smem, st>>>(mapDY, mapX, p, stages, a_bytes, b_bytes, b_tx);
    } else {
      CD_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      wgrad_tc_kernel<64, true><<<grid, kThreads, smem, st>>>(mapDY, mapX, p, stages, a_bytes, b_bytes, b_tx);
    }
    if (bias_done) *bias_done = 1;
  } else if (BN == 128) {
    CD_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wgrad_tc_kernel<128, false><<<grid, kThreads, smem, st>>>(mapDY, mapX, p, stages, a_bytes, b_bytes, b_tx);
  } else {
    CD_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wgrad_tc_kernel<64, false><<<grid, kThreads, smem, st>>>(mapDY, mapX, p, stages, a_bytes, b_bytes, b_tx);
  }
  CD_LAUNCH_CHECK();
  return 0;
}
