// rb_reducers.cu — focal-plane analysis reducers whose control flow must follow a plain C evaluation step for step.
// Compiled with -fmad=false: no multiply-add contraction, so every product, sum, quotient and square root rounds exactly
// as in the reference's scalar C++ (IEEE double), and the data-dependent search below takes the same path.
#include <cuda_runtime.h>
#include <math.h>
#include <stddef.h>

// ---- AGeoUtil::ContainmentRadius (reference src/AGeoUtil.cxx:18-42,198-308) on a device-resident histogram.
// One block per histogram; the search itself is sequential (every step depends on the previous sums), so the
// parallelism is inside SumInRadius: the block strides over the bins, reduces, and every thread then takes the same
// branch.  This file is compiled without multiply-add contraction, so the products and sums that decide `d2 <= r2` —
// and therefore the bins selected and the whole search path — are the same as in a plain C evaluation.
#define CR_THREADS 256
// The bins of one histogram row that satisfy the reference's test (cx-x)^2 + (cy-y)^2 <= r^2 form one contiguous run
// (the left side is monotone in |cx - x| also in floating point), so a row contributes prefix[hi+1] - prefix[lo] of the
// row's running sum of positive contents.  The run is located from the analytic chord and then corrected with the exact
// predicate, so the selected bins are precisely those of the reference's bin-by-bin loop; with integer contents the
// running sums are exact and the total is identical.  One thread per row instead of one pass over all nx*ny bins.
struct CRHist {
  const double* pre;  // [ny][nx + 1] running sums of max(content, 0) along x
  int nx, ny;
  double xmin, wx, ymin, wy;
};
__device__ inline bool cr_pred(const CRHist& h, int ix, double x, double dy2, double r2) {
  double cx = h.xmin + ix * h.wx + 0.5 * h.wx;  // TAxis::GetBinCenter
  return (cx - x) * (cx - x) + dy2 <= r2;
}
// SumInRadius: content of the bins whose centre lies within r of (x, y); block-wide, every thread gets the result
__device__ double cr_sum(const CRHist& h, double x, double y, double r, double* s_red) {
  const double r2 = r * r;
  double part = 0;
  for (int iy = threadIdx.x; iy < h.ny; iy += CR_THREADS) {
    double cy = h.ymin + iy * h.wy + 0.5 * h.wy, dy = cy - y, dy2 = dy * dy;
    if (!(dy2 <= r2)) continue;
    int lo = 0, hi = h.nx - 1;
    if (r2 < 1e300) {
      double half = sqrt(r2 - dy2), ul = (x - half - h.xmin) / h.wx - 0.5, uh = (x + half - h.xmin) / h.wx - 0.5;
      lo = ul <= 0. ? 0 : (ul >= (double)h.nx ? h.nx : (int)ceil(ul));
      hi = uh < 0. ? -1 : (uh >= (double)(h.nx - 1) ? h.nx - 1 : (int)floor(uh));
      while (lo > 0 && cr_pred(h, lo - 1, x, dy2, r2)) lo--;
      while (hi + 1 < h.nx && cr_pred(h, hi + 1, x, dy2, r2)) hi++;
      while (lo <= hi && !cr_pred(h, lo, x, dy2, r2)) lo++;
      while (hi >= lo && !cr_pred(h, hi, x, dy2, r2)) hi--;
    }
    if (lo <= hi) {
      const double* row = h.pre + (size_t)iy * (h.nx + 1);
      part += row[hi + 1] - row[lo];
    }
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
                                                            const double* stats, size_t stats_stride, double fraction, double* out, double* prefix) {
  __shared__ double s_red[CR_THREADS / 32];
  CRHist h;
  double* pre = prefix + blockIdx.x * (size_t)(nx + 1) * ny;
  h.pre = pre;
  h.nx = nx; h.ny = ny;
  h.xmin = xmin; h.wx = (xmax - xmin) / nx;
  h.ymin = ymin; h.wy = (ymax - ymin) / ny;
  {  // running sums along x, one thread per row; bins with content <= 0 do not count (the reference skips them)
    const BinT* c = hist + blockIdx.x * hist_stride;
    for (int iy = threadIdx.x; iy < ny; iy += CR_THREADS) {
      double acc = 0;
      double* row = pre + (size_t)iy * (nx + 1);
      row[0] = 0;
      for (int ix = 0; ix < nx; ix++) {
        double v = (double)c[ix + (size_t)nx * iy];
        if (v > 0) acc += v;
        row[ix + 1] = acc;
      }
    }
    __syncthreads();
  }
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

// `prefix` = device scratch of nhist * (nx + 1) * ny doubles
int rb_launch_containment_u64(int nhist, const unsigned long long* hist, int nx, double xmin, double xmax, int ny, double ymin, double ymax, const double* stats,
                              double fraction, double* out, double* prefix, cudaStream_t st) {
  k_containment<unsigned long long><<<nhist, CR_THREADS, 0, st>>>(hist, (size_t)nx * ny, nx, xmin, xmax, ny, ymin, ymax, stats, 5, fraction, out, prefix);
  return (int)cudaGetLastError();
}
int rb_launch_containment_f64(int nhist, const double* hist, int nx, double xmin, double xmax, int ny, double ymin, double ymax, const double* stats,
                              double fraction, double* out, double* prefix, cudaStream_t st) {
  k_containment<double><<<nhist, CR_THREADS, 0, st>>>(hist, (size_t)nx * ny, nx, xmin, xmax, ny, ymin, ymax, stats, 5, fraction, out, prefix);
  return (int)cudaGetLastError();
}
