// cta_group::2 variant of the tcgen05 tap-list convolution (same contract and epilogue as conv_tc.cu).
//
// Two CTAs of a cluster (an SM pair) cooperate on a 256-pixel x BN tile: each CTA stages ITS 128 pixels of the activation
// tile and HALF of the weight tile (BN/2 rows); one tcgen05.mma.cta_group::2 issued by the leader consumes both CTAs'
// shared memory (M = 256) and writes 128 accumulator rows into each CTA's TMEM.  Per SM this halves the weight bytes
// fetched over TMA and read from shared memory per MMA -- the two walls measured for the 1-CTA kernel (DESIGN.md section 8:
// a 128x128x8 TF32 MMA needs 128 B/clk of shared-memory reads; the pair needs 96 B/clk at N = 128 and 64 B/clk at N = 256).
//
// Protocol (after the CUTLASS sm100 2-SM collectives):
//  * TMA loads of both CTAs signal the LEADER's full barrier (cp.async.bulk.tensor...cta_group::2 with the peer bit of the
//    barrier address cleared); the leader arms it with the byte count of both CTAs.
//  * the leader's tcgen05.commit.cta_group::2...multicast::cluster releases the smem stage / publishes the accumulator in
//    BOTH CTAs; epilogue warps of both CTAs arrive on the leader's tmem_empty barrier (remote mbarrier.arrive).
//  * TMEM is allocated with tcgen05.alloc.cta_group::2 by the same warp of both CTAs.
#include "tc_common.cuh"

namespace {

constexpr int kEpiWarps = 16;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kTileM = 128;
constexpr int kChunkK = 32;
constexpr int kABytes = kTileM * 128;
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;      // shared::cluster address of the even CTA of the pair

struct Tc2Params {
  int B, Hg, Wg;
  int TW, TH, TN;
  int tiles_x, tiles_y, tiles_n, tiles_co, m_tiles, total_tiles;   // total_tiles = ceil(m_tiles/2) * tiles_co (pair tiles)
  int sy, sx;
  int Cout;
  int nsrc;
  int ntaps[2];
  int kchunks[2];
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

__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {          // arrive on the even CTA's copy of `bar`
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" :: "r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :: "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma2_load_3d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      :: "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {              // arrives on `bar` in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               :: "r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3)) : "memory");
}
__device__ __forceinline__ void mma2_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
               :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;\nbarrier.cluster.wait.acquire;" ::: "memory");
}

template <int BN, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                const __grid_constant__ CUtensorMap mapB0, const __grid_constant__ CUtensorMap mapB1,
                const Tc2Params p) {
  constexpr int kBBytes = (BN / 2) * 128;                 // this CTA's half of the weight tile
  constexpr int kStageBytes = kABytes + kBBytes;
  constexpr uint32_t kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  // instruction descriptor: D=f32, A=B=tf32, both K-major, N=BN, M=256 (two CTAs x 128 rows)
  constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(BN >> 3) << 17) | (uint32_t(256 >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&tmem_full[0], 1); mbar_init(&tmem_full[1], 1);
    mbar_init(&tmem_empty[0], 2 * kEpiWarps); mbar_init(&tmem_empty[1], 2 * kEpiWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync_all();                                     // both CTAs' barriers exist before any remote signal
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int kiters = p.ntaps[0] * p.kchunks[0] + (p.nsrc > 1 ? p.ntaps[1] * p.kchunks[1] : 0);
  const int ncl = gridDim.x >> 1, cl = blockIdx.x >> 1;   // clusters and this cluster's index

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; completion lands on the leader's full barrier) =====================
    // one elected thread runs the whole schedule
    if (elect_one()) {
      uint32_t stage = 0, ph = 0;
      for (int tile = cl; tile < p.total_tiles; tile += ncl) {
        const int co_t = tile % p.tiles_co;
        int mt = (tile / p.tiles_co) * 2 + static_cast<int>(rank);        // this CTA's 128-pixel tile (may be past the end: zero fill)
        const int tx = mt % p.tiles_x; mt /= p.tiles_x;
        const int ty = mt % p.tiles_y;
        const int tn = mt / p.tiles_y;
        const int x0 = tx * p.TW * p.sx, y0 = ty * p.TH * p.sy, n0 = tn * p.TN;
        const int co0 = co_t * BN + static_cast<int>(rank) * (BN / 2);     // this CTA's half of the weight rows
        for (int s = 0; s < p.nsrc; ++s) {
          const CUtensorMap* mA = s ? &mapA1 : &mapA0;
          const CUtensorMap* mB = s ? &mapB1 : &mapB0;
          for (int tap = 0; tap < p.ntaps[s]; ++tap) {
            const int xin = x0 + p.dx[s][tap], yin = y0 + p.dy[s][tap];
            for (int kc = 0; kc < p.kchunks[s]; ++kc) {
              mbar_wait(&empty_bar[stage], ph ^ 1u);
              if (leader) mbar_expect_tx(&full_bar[stage], 2 * kStageBytes);
              const uint32_t sa = smem_u32(smem + stage * kStageBytes);
              tma2_load_4d(sa, mA, &full_bar[stage], kc * kChunkK, xin, yin, n0);
              tma2_load_3d(sa + kABytes, mB, &full_bar[stage], kc * kChunkK, co0, tap);
              if (++stage == STAGES) { stage = 0; ph ^= 1u; }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (leader && elect_one()) {
      // ===================== MMA issuer (leader CTA only): one elected thread runs the whole loop =====================
      // the barrier of the NEXT chunk is polled between this chunk's MMAs (conv_tc.cu explains; tools/micro/umma_rate.cu)
      uint32_t stage = 0, ph = 0, tcount = 0;
      const uint64_t desc0 = make_kmajor_sw128_desc(smem_u32(smem));
      if (cl < p.total_tiles) {
        mbar_wait(&tmem_empty[0], 1u);
        mbar_wait(&full_bar[0], 0u);
      }
      for (int tile = cl; tile < p.total_tiles; tile += ncl, ++tcount) {
        const uint32_t acc = tcount & 1u;
        const uint32_t tmem_d = tmem_base + acc * BN;
        const bool more_tiles = tile + ncl < p.total_tiles;
        for (int k = 0; k < kiters; ++k) {
          tc_fence_after();
          const uint64_t da = desc0 + static_cast<uint64_t>(stage * uint32_t(kStageBytes >> 4));   // the 14-bit address field cannot carry
          const uint64_t db = da + uint64_t(kABytes >> 4);
          uint32_t sn = stage + 1, phn = ph;
          if (sn == STAGES) { sn = 0; phn ^= 1u; }
          const bool last = k == kiters - 1;
          const bool more = !last || more_tiles;
          uint32_t ready = 0;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            if (kk == 2 && more) {                  // one non-blocking poll
              ready = mbar_test(&full_bar[sn], phn);
              if (last) ready &= mbar_test(&tmem_empty[acc ^ 1u], (((tcount + 1) >> 1) & 1u) ^ 1u);
            }
            mma2_tf32(tmem_d, da + uint64_t(kk * 2), db + uint64_t(kk * 2), kIdesc, (k | kk) != 0 ? 1u : 0u);
          }
          tc_commit_pair(&empty_bar[stage]);
          if (more && !ready) {
            mbar_wait(&full_bar[sn], phn);
            if (last) mbar_wait(&tmem_empty[acc ^ 1u], (((tcount + 1) >> 1) & 1u) ^ 1u);
          }
          stage = sn; ph = phn;
        }
        tc_commit_pair(&tmem_full[acc]);
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..17 of both CTAs, each on its own 128 accumulator rows) =====================
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    const int m = q * 32 + lane;
    const int xx = m % p.TW;
    const int yy = (m / p.TW) % p.TH;
    const int nn = m / (p.TW * p.TH);
    uint32_t tcount = 0;
    for (int tile = cl; tile < p.total_tiles; tile += ncl, ++tcount) {
      const int co_t = tile % p.tiles_co;
      const int mtile = (tile / p.tiles_co) * 2 + static_cast<int>(rank);
      int mt = mtile;
      const int tx = mt % p.tiles_x; mt /= p.tiles_x;
      const int ty = mt % p.tiles_y;
      const int tn = mt / p.tiles_y;
      const int gx = tx * p.TW + xx, gy = ty * p.TH + yy, b = tn * p.TN + nn;
      const int co0 = co_t * BN;
      const bool valid = mtile < p.m_tiles && b < p.B;
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
      for (int c = cg * 32; c < BN; c += 32 * (kEpiWarps / 4)) {
        uint32_t r[32];
        tmem_ld32(taddr + c, r);
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
              float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
              if (p.bias) { const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + co0 + c + j)); v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w; }
              if (rrow) { const float4 rr = *reinterpret_cast<const float4*>(rrow + co0 + c + j); v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w; }
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
      if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);       // both CTAs release the accumulator to the leader's MMA warp
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                            // no CTA frees TMEM / exits while its peer may still signal it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

int g_sms2 = 0;

template <int BN, int STAGES>
int launch2(const CUtensorMap* maps, const Tc2Params& p, cudaStream_t st) {
  const size_t smem = size_t(STAGES) * (kABytes + (BN / 2) * 128) + 1024 + 256;
  static bool attr_done = false;
  if (!attr_done) {
    CD_CUDA(cudaFuncSetAttribute(conv_tc2_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  int clusters = g_sms2 / 2; if (clusters > p.total_tiles) clusters = p.total_tiles;
  conv_tc2_kernel<BN, STAGES><<<2 * clusters, kThreads, smem, st>>>(maps[0], maps[1], maps[2], maps[3], p);
  CD_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// returns 1 when the problem is not eligible for the SM-pair kernel (caller uses the 1-CTA kernel), 0 on success, <0 on error
int cd_conv_fwd_tc2(const CdConvDesc* d, cudaStream_t st, int BN) {
  if (d->nsrc < 1 || d->nsrc > 2) return 1;
  for (int s = 0; s < d->nsrc; ++s) if (d->s[s].w_per_batch) return 1;       // the pair shares ONE weight tile
  EncodeTiledFn enc = get_encode();
  CD_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  if (!g_sms2) { int dev = 0; CD_CUDA(cudaGetDevice(&dev)); CD_CUDA(cudaDeviceGetAttribute(&g_sms2, cudaDevAttrMultiProcessorCount, dev)); }
  Tc2Params p{};
  p.B = d->B; p.Hg = d->Hg; p.Wg = d->Wg; p.sy = d->sy; p.sx = d->sx; p.Cout = d->Cout; p.nsrc = d->nsrc;
  if (d->Wg >= 128) {
    if (d->Wg % 128 != 0) return 1;
    p.TW = 128; p.TH = 1; p.TN = 1;
  } else {
    if (!is_pow2(d->Wg)) return 1;
    p.TW = d->Wg;
    int th = 128 / p.TW; if (th > d->Hg) th = d->Hg;
    if (!is_pow2(th) || d->Hg % th != 0) return 1;
    p.TH = th; p.TN = 128 / (p.TW * p.TH);
  }
  if (p.TW * d->sx > 256 || p.TH * d->sy > 256) return 1;
  p.tiles_x = d->Wg / p.TW; p.tiles_y = d->Hg / p.TH; p.tiles_n = cd_cdiv(d->B, p.TN);
  p.m_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
  if (p.m_tiles < 2) return 1;
  // BN = 256 by default; 128 / 64 (each CTA stages 64 / 32 weight rows) are chosen by the caller's switch (cd_conv_tc_set_2cta_bn)
  if (BN != 256 && BN != 128 && BN != 64) return 1;
  if (d->Cout % BN != 0) return 1;
  p.tiles_co = d->Cout / BN;
  p.total_tiles = ((p.m_tiles + 1) / 2) * p.tiles_co;
  p.out = d->out; p.out_ld = d->out_ld; p.Ho = d->Ho; p.Wo = d->Wo;
  p.oys = d->oys; p.oxs = d->oxs; p.oy0 = d->oy0; p.ox0 = d->ox0;
  p.bias = d->bias; p.resid = d->resid; p.resid_ld = d->resid_ld; p.act = d->act; p.round_tf32 = d->round_tf32;
  p.out2 = d->out2; p.out2_ld = d->out2_ld; p.aux = d->aux; p.aux_ld = d->aux_ld;
  {
    auto ok8 = [](const void* ptr, int ld) { return ptr == nullptr || ((reinterpret_cast<uintptr_t>(ptr) & 31) == 0 && ld % 8 == 0); };
    p.vec8 = ok8(d->out, d->out_ld) && ok8(d->out2, d->out2_ld) && ok8(d->resid, d->resid_ld) && ok8(d->aux, d->aux_ld) && ok8(d->bias, 8);
  }
  if ((reinterpret_cast<uintptr_t>(d->out) & 15) || d->out_ld % 4) return 1;
  if (d->resid && ((reinterpret_cast<uintptr_t>(d->resid) & 15) || d->resid_ld % 4)) return 1;
  if (d->out2 && ((reinterpret_cast<uintptr_t>(d->out2) & 15) || d->out2_ld % 4)) return 1;
  if (d->bias && (reinterpret_cast<uintptr_t>(d->bias) & 15)) return 1;
  if (d->act == CD_ACT_GELU_BWD && (!d->aux || (reinterpret_cast<uintptr_t>(d->aux) & 15) || d->aux_ld % 4)) return 1;

  CUtensorMap maps[4];
  const CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_TFLOAT32;
  for (int s = 0; s < 2; ++s) {
    const CdConvSrc& cs = d->s[s < d->nsrc ? s : 0];
    if (cs.C % kChunkK != 0 || cs.C <= 0 || cs.ntaps < 1 || cs.ntaps > CD_MAX_TAPS) return 1;
    if ((reinterpret_cast<uintptr_t>(cs.src) & 15) || cs.ld % 4 || (reinterpret_cast<uintptr_t>(cs.w) & 15)) return 1;
    p.ntaps[s] = cs.ntaps; p.kchunks[s] = cs.C / kChunkK;
    for (int t = 0; t < cs.ntaps; ++t) { p.dy[s][t] = (int8_t)cs.dy[t]; p.dx[s][t] = (int8_t)cs.dx[t]; }
    {
      cuuint64_t dims[4] = {(cuuint64_t)cs.C, (cuuint64_t)cs.W, (cuuint64_t)cs.H, (cuuint64_t)d->B};
      cuuint64_t strides[3] = {(cuuint64_t)cs.ld * 4, (cuuint64_t)cs.ld * 4 * cs.W, (cuuint64_t)cs.ld * 4 * cs.W * cs.H};
      cuuint32_t box[4] = {(cuuint32_t)kChunkK, (cuuint32_t)(p.TW * d->sx), (cuuint32_t)(p.TH * d->sy), (cuuint32_t)p.TN};
      cuuint32_t estr[4] = {1, (cuuint32_t)d->sx, (cuuint32_t)d->sy, 1};
      CUresult r = enc(&maps[s], dt, 4, const_cast<float*>(cs.src), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      CD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(A%d) failed: %d", s, (int)r);
    }
    {
      cuuint64_t dims[3] = {(cuuint64_t)cs.C, (cuuint64_t)d->Cout, (cuuint64_t)cs.ntaps};
      cuuint64_t strides[2] = {(cuuint64_t)cs.C * 4, (cuuint64_t)cs.C * 4 * d->Cout};
      cuuint32_t box[3] = {(cuuint32_t)kChunkK, (cuuint32_t)(BN / 2), 1};
      cuuint32_t estr[3] = {1, 1, 1};
      CUresult r = enc(&maps[2 + s], dt, 3, const_cast<float*>(cs.w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      CD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(B%d) failed: %d", s, (int)r);
    }
  }
  if (BN == 256) return launch2<256, 7>(maps, p, st);
  if (BN == 128) return launch2<128, 8>(maps, p, st);
  return launch2<64, 8>(maps, p, st);
}
