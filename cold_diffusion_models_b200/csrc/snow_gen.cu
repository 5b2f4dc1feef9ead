// Snow-layer generation on the device (snowification/diffusion/forward_process_impl.py "FP": clipped_zoom FP:32-42,
// Snow.generate_snow_layer FP:252-355).  Upstream builds the layers on the host with scipy.ndimage.zoom and T (time steps)
// torch CPU convolutions -- every p_losses call when random_snow is set.  Here the host only draws the random numbers (the
// numpy / torch generators the reference draws from, so seeds keep their meaning); everything after that is two kernels:
//
//  1. snow_zoom_kernel: the order-1 spline zoom of scipy.ndimage.zoom (mode 'constant', grid_mode False) of the centre crop,
//     followed by the centre trim, evaluated directly at the H x W pixels that survive the trim.  Double precision, the same
//     operations in the same order as scipy's NI_ZoomShift (coordinate o * (n-1)/(m-1); weights 1 - frac and 1 - (1 - frac);
//     support points beyond the edge mirrored; ((v * wy) * wx) summed over (y0,x0), (y0,x1), (y1,x0), (y1,x1); a coordinate that
//     rounds past n - 1 yields the constant 0) -> bit-identical to scipy (tests/test_simt_cpu_kernels.py), so the threshold that
//     follows cuts exactly the same pixels.  The result is cast to fp32 as torch.Tensor(ndarray) does.
//  2. snow_blur_kernel: per time step t: zero below thres[t], clip to [0, 1], motion blur = k-tap 1-D cross-correlation with
//     zero 'same' padding along x (horizontal kernel, FP:333) or along y with the taps reversed (torch.rot90 of it, FP:335),
//     chosen per (step, snow sample); the single channel is written to the three channels conv2d's (3,1,k,k) kernel produces.
#include "cd_common.cuh"

namespace {

__global__ void snow_zoom_kernel(const double* __restrict__ noise, int SB, int ch, int trim, int H, double scale,
                                 float* __restrict__ base) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(SB) * H * H;
  if (idx >= total) return;
  const int x = static_cast<int>(idx % H), y = static_cast<int>((idx / H) % H), s = static_cast<int>(idx / (static_cast<long long>(H) * H));
  const double ccy = __dmul_rn(static_cast<double>(y + trim), scale), ccx = __dmul_rn(static_cast<double>(x + trim), scale);
  const double last = static_cast<double>(ch - 1);
  float r = 0.f;
  if (ccy <= last && ccx <= last) {
    const double fy = floor(ccy), fx = floor(ccx);
    const double w0y = __dsub_rn(1.0, __dsub_rn(ccy, fy)), w0x = __dsub_rn(1.0, __dsub_rn(ccx, fx));
    const double w1y = __dsub_rn(1.0, w0y), w1x = __dsub_rn(1.0, w0x);
    const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
    int y1 = y0 + 1, x1 = x0 + 1;
    if (y1 > ch - 1) y1 = 2 * (ch - 1) - y1;             // mirrored support point (its weight is 0 up to rounding)
    if (x1 > ch - 1) x1 = 2 * (ch - 1) - x1;
    const double* img = noise + static_cast<long long>(s) * ch * ch;
    const double a = __dmul_rn(__dmul_rn(img[y0 * ch + x0], w0y), w0x);
    const double b = __dmul_rn(__dmul_rn(img[y0 * ch + x1], w0y), w1x);
    const double c = __dmul_rn(__dmul_rn(img[y1 * ch + x0], w1y), w0x);
    const double d = __dmul_rn(__dmul_rn(img[y1 * ch + x1], w1y), w1x);
    r = static_cast<float>(__dadd_rn(__dadd_rn(__dadd_rn(a, b), c), d));
  }
  base[idx] = r;
}

__global__ void snow_blur_kernel(const float* __restrict__ base, int SB, int H, const float* __restrict__ thres,
                                 const float* __restrict__ taps, int k, const unsigned char* __restrict__ vertical, int T,
                                 float* __restrict__ snow) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long HW = static_cast<long long>(H) * H;
  const long long total = static_cast<long long>(T) * SB * HW;
  if (idx >= total) return;
  const int x = static_cast<int>(idx % H), y = static_cast<int>((idx / H) % H);
  const int s = static_cast<int>((idx / HW) % SB), t = static_cast<int>(idx / (HW * SB));
  const float th = thres[t];
  const float* img = base + static_cast<long long>(s) * HW;
  const float* w = taps + static_cast<long long>(t) * k;
  const bool vert = vertical[static_cast<long long>(t) * SB + s] != 0;
  const int half = k / 2;
  float acc = 0.f;
  for (int j = 0; j < k; ++j) {
    const int yy = vert ? y + j - half : y, xx = vert ? x : x + j - half;
    if (static_cast<unsigned>(yy) < static_cast<unsigned>(H) && static_cast<unsigned>(xx) < static_cast<unsigned>(H)) {
      float v = img[yy * H + xx];
      v = v < th ? 0.f : fminf(fmaxf(v, 0.f), 1.f);
      acc = fmaf(vert ? w[k - 1 - j] : w[j], v, acc);
    }
  }
  float* o = snow + (static_cast<long long>(t) * SB + s) * 3 * HW + static_cast<long long>(y) * H + x;
  o[0] = acc; o[HW] = acc; o[2 * HW] = acc;
}

}  // namespace

extern "C" int cd_snow_layers(const double* noise, int SB, int ch, int m, int trim, int H, const float* thres, const float* taps,
                              int k, const unsigned char* vertical, int T, float* base, float* snow, void* stream) {
  CD_REQUIRE(noise && thres && taps && vertical && base && snow, "cd_snow_layers: null pointer");
  CD_REQUIRE(SB >= 1 && T >= 1 && H >= 1 && ch >= 2 && m >= 2 && k >= 1 && (k & 1) == 1, "cd_snow_layers: bad sizes (SB=%d T=%d H=%d ch=%d m=%d k=%d)", SB, T, H, ch, m, k);
  CD_REQUIRE(trim >= 0 && trim + H <= m, "cd_snow_layers: the %d rows kept after the trim (%d) exceed the zoomed size %d", H, trim, m);
  const double scale = static_cast<double>(ch - 1) / static_cast<double>(m - 1);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n1 = static_cast<long long>(SB) * H * H;
  snow_zoom_kernel<<<cd_cdiv(n1, 256), 256, 0, st>>>(noise, SB, ch, trim, H, scale, base);
  CD_LAUNCH_CHECK();
  const long long n2 = static_cast<long long>(T) * n1;
  CD_REQUIRE(cd_cdiv(n2, 256) > 0, "cd_snow_layers: too many outputs");
  snow_blur_kernel<<<cd_cdiv(n2, 256), 256, 0, st>>>(base, SB, H, thres, taps, k, vertical, T, snow);
  CD_LAUNCH_CHECK();
  return 0;
}
