// rb_reducers.cu — focal-plane analysis reducers whose control flow must follow a plain C evaluation step for step.
// Compiled with -fmad=false: no multiply-add contraction, so every product, sum, quotient and square root rounds exactly
// as in the reference's scalar C++ (IEEE double), and the data-dependent search below takes the same path.
#include <cuda_runtime.h>
#include <math.h>
#include <stddef.h>

// ---- AGeoUtil::ContainmentRadius (reference src/AGeoUtil.cxx:18-42,198-308) on a device-resident histogram.
// One block per histogram; the search itself is sequential (every step depends on the previous sums), so the
// parallelism is inside SumInRadius: the block strides over the bins, reduces, and every thread then takes the same
// branch.  Products and sums that decide `d2 <= r2` are kept un-contracted (__dmul_rn/__dadd_rn) so that the bins
// selected — and therefore the whole search path — are the same as in a plain C evaluation.
#define CR_THREADS 1024
template <class BinT> struct CRHist {
  const BinT* c;
  int nx, ny;
  double xmin, wx, ymin, wy;
};
template <class BinT> __device__ double cr_sum(const CRHist<BinT>& h, double x, double y, double r, double* s_red) {
  const double r2 = __dmul_rn(r, r);
  double part = 0;
  const int nb = h.nx * h.ny;
  for (int b = threadIdx.x; b < nb; b += CR_THREADS) {
    double c = (double)h.c[b];
    if (c <= 0) continue;
    int ix = b % h.nx, iy = b / h.nx;
    double cx = __dadd_rn(__dadd_rn(h.xmin, __dmul_rn((double)ix, h.wx)), __dmul_rn(0.5, h.wx));  // TAxis::GetBinCenter
    double cy = __dadd_rn(__dadd_rn(h.ymin, __dmul_rn((double)iy, h.wy)), __dmul_rn(0.5, h.wy));
    double ddx = __dadd_rn(cx, -x), ddy = __dadd_rn(cy, -y);
    double d2 = __dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy));
    if (d2 <= r2) part += c;
  }
  for (int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
  __syncthreads();  // s_red may still be read from the previous call
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
  __syncthreads();
  double tot = 0;
#pragma unroll
  for (int w = 0; w < CR_THREADS / 32; w++) tot += s_red[w];  // same order in every thread: identical result everywhere
  return tot;
}
template <class BinT>
__global__ void __launch_bounds__(CR_THREADS) k_containment(const BinT* hist, size_t hist_stride, int nx, double xmin, double xmax, int ny, double ymin, double ymax,
                                                            const double* stats, size_t stats_stride, double fraction, double* out) {
  __shared__ double s_red[CR_THREADS / 32];
  CRHist<BinT> h;
  h.c = hist + blockIdx.x * hist_stride;
  h.nx = nx; h.ny = ny;
  h.xmin = xmin; h.wx = (xmax - xmin) / nx;
  h.ymin = ymin; h.wy = (ymax - ymin) / ny;
  const double* st = stats + blockIdx.x * stats_stride;
  double sw = st[0];
  double x = sw != 0 ? st[1] / sw : 0, y = sw != 0 ? st[2] / sw : 0;  // TH2::GetMean
  double sdx = sw != 0 ? sqrt(fabs(st[3] / sw - x * x)) : 0, sdy = sw != 0 ? sqrt(fabs(st[4] / sw - y * y)) : 0;  // TH2::GetStdDev
  double r = sqrt(sdx * sdx + sdy * sdy) * 1.5;
  double dr = 0.1 * r;
  double sum_goal = cr_sum(h, 0., 0., 1e300, s_red) * fraction;  // TH2::Integral
  int no_shift = 0, no_stable = 0;
  for (int i = 0; i < 100 && no_shift < 30; i++) {
    bool stable_r = false, stable_x = true, stable_y = true;
    double sum0 = cr_sum(h, x, y, r, s_red);
    double next_r = r;
    if (sum0 < sum_goal) {
      double sum1 = cr_sum(h, x, y, r + dr, s_red);
      if (sum1 == sum0) { dr *= 2.; continue; }
      next_r = r + dr * (sum_goal - sum0) / (sum1 - sum0);
    } else if (sum0 != sum_goal) {
      double sum1 = cr_sum(h, x, y, r - dr, s_red);
      if (sum1 == sum0) { dr *= 2.; continue; }
      next_r = r - dr * (sum0 - sum_goal) / (sum0 - sum1);
    }
    if (next_r < 0.) next_r = 0.5 * r;
    if (next_r < 0.5 * r) next_r = 0.5 * r;
    if (next_r > 2. * r) next_r = 2. * r;
    stable_r = fabs(next_r - r) < 0.0001 * r;
    r = next_r;
    {
      double sum1 = cr_sum(h, x, y, r, s_red);
      dr *= sum0 != sum_goal ? fabs((sum1 - sum_goal) / (sum0 - sum_goal)) : 0.5;
      if (dr > 0.5 * r) dr = 0.5 * r;
      if (dr < 0.0005 * r) dr = 0.0005 * r;
      no_shift++;
      for (double dx = 0.25 * r; dx > 0.1 * dr; dx *= 0.25) {
        double sum_x1 = cr_sum(h, x + dx, y, r, s_red), sum_x2 = cr_sum(h, x - dx, y, r, s_red);
        while (sum_x1 > sum1) {
          no_shift = 0;
          x += dx;
          sum_x2 = sum1;
          sum1 = sum_x1;
          sum_x1 = cr_sum(h, x + dx, y, r, s_red);
          stable_x = false;
        }
        while (sum_x2 > sum1) {
          no_shift = 0;
          x -= dx;
          sum_x1 = sum1;
          sum1 = sum_x2;
          sum_x2 = cr_sum(h, x - dx, y, r, s_red);
          stable_x = false;
        }
      }
    }
    for (double dy = 0.1 * r; dy > 0.1 * dr; dy *= 0.25) {
      double sum1 = cr_sum(h, x, y, r, s_red), sum_y1 = cr_sum(h, x, y + dy, r, s_red), sum_y2 = cr_sum(h, x, y - dy, r, s_red);
      while (sum_y1 > sum1) {
        no_shift = 0;
        y += dy;
        sum_y2 = sum1;
        sum1 = sum_y1;
        sum_y1 = cr_sum(h, x, y + dy, r, s_red);
        stable_y = false;
      }
      while (sum_y2 > sum1) {
        no_shift = 0;
        y -= dy;
        sum_y1 = sum1;
        sum1 = sum_y2;
        sum_y2 = cr_sum(h, x, y - dy, r, s_red);
        stable_y = false;
      }
    }
    if (stable_r && stable_x && stable_y) no_stable++;
    else no_stable = 0;
    if (no_stable >= 4) break;
  }
  if (threadIdx.x == 0) {
    out[3 * blockIdx.x] = r;
    out[3 * blockIdx.x + 1] = x;
    out[3 * blockIdx.x + 2] = y;
  }
}

int rb_launch_containment_u64(int nhist, const unsigned long long* hist, int nx, double xmin, double xmax, int ny, double ymin, double ymax, const double* stats,
                              double fraction, double* out, cudaStream_t st) {
  k_containment<unsigned long long><<<nhist, CR_THREADS, 0, st>>>(hist, (size_t)nx * ny, nx, xmin, xmax, ny, ymin, ymax, stats, 5, fraction, out);
  return (int)cudaGetLastError();
}
int rb_launch_containment_f64(int nhist, const double* hist, int nx, double xmin, double xmax, int ny, double ymin, double ymax, const double* stats,
                              double fraction, double* out, cudaStream_t st) {
  k_containment<double><<<nhist, CR_THREADS, 0, st>>>(hist, (size_t)nx * ny, nx, xmin, xmax, ny, ymin, ymax, stats, 5, fraction, out);
  return (int)cudaGetLastError();
}
