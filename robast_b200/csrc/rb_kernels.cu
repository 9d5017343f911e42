// rb_kernels.cu — sm_100a kernels and the C ABI (include/robast_b200.h) of the B200 tracer.
//
//  k_trace<DEPTH>   one thread per live ray; runs `max_steps` iterations of the reference's
//                   while(ray->IsRunning()) loop (src/AOpticsManager.cxx:359-518) in registers.
//                   FP64-pipe bound.  Replaces TGeoNavigator::FindNextBoundaryAndStep + DoReflection
//                   + DoFresnel + GetFacetNormal and every TGeoShape::Dist*/ComputeNormal virtual call.
//  k_compact        stable stream compaction of surviving ray indices between bounces: warp scan +
//                   block scan + decoupled look-back across tiles (single pass, 4 B read + 4 B write
//                   per live ray).  HBM bound.  Replaces the survivor test of the reference's while loop.
//  k_shoot          ARayShooter generators on device (src/ARayShooter.cxx:122-460), 64 B/ray written.
//  k_hist2d/k_moments  focal-plane reducers (TH2D fill, GetMean/GetRMS of the tutorials).
//  k_tmm            AMultilayer::CoherentTMMMixed per (theta, lambda) pair.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <functional>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "rb_build.h"
#include "rb_variants.h"

// ------------------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};
static int g_profile = 0;
struct ProfEvt { cudaEvent_t a, b; int kind; };
static std::vector<ProfEvt> g_prof_events;
static std::mutex g_prof_mutex;
static double g_prof_ms[2] = {0, 0};
static long long g_prof_n[2] = {0, 0};

#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) throw std::runtime_error(std::string("cuda: ") + #call + ": " + cudaGetErrorString(e_)); \
  } while (0)

template <class F> static int guard(F f) {
  try {
    f();
    return RBG_OK;
  } catch (NotSupported& e) {
    g_err = e.what();
    return RBG_ENOTSUP;
  } catch (Invalid& e) {
    g_err = e.what();
    return RBG_EINVAL;
  } catch (std::bad_alloc&) {
    g_err = "out of host memory";
    return RBG_ENOMEM;
  } catch (std::runtime_error& e) {
    g_err = e.what();
    return g_err.find("cuda") == 0 ? RBG_ECUDA : RBG_EINTERNAL;
  } catch (std::exception& e) {
    g_err = e.what();
    return RBG_EINTERNAL;
  } catch (...) {
    g_err = "unknown exception";
    return RBG_EINTERNAL;
  }
}

struct ProfScope {
  cudaStream_t st;
  int kind;
  ProfEvt e;
  bool on;
  ProfScope(cudaStream_t s, int k) : st(s), kind(k), on(g_profile != 0) {
    if (on) {
      if (cudaEventCreate(&e.a) != cudaSuccess) { on = false; cudaGetLastError(); return; }
      if (cudaEventCreate(&e.b) != cudaSuccess) { on = false; cudaEventDestroy(e.a); cudaGetLastError(); return; }
      e.kind = kind;
      cudaEventRecord(e.a, st);
    }
  }
  ~ProfScope() {
    if (on) {
      cudaEventRecord(e.b, st);
      std::lock_guard<std::mutex> lk(g_prof_mutex);
      g_prof_events.push_back(e);
    }
  }
};

// ------------------------------------------------------------------------------------------------ kernels
// ---- stream compaction (single pass, decoupled look-back, stable)
// One tile = 256 threads x 16 rays, blocked so that every thread reads 64 contiguous bytes (4 x 128-bit loads).
// Tile prefixes are published as (flag << 32 | value) words; warp 0 looks back over 32 predecessor tiles at a time.
// The number of input entries comes from device memory (`count_in`, the previous compaction's output; null: n_max), so the
// host never reads a counter back; the grid covers n_max and the surplus tiles leave at once.  The survivor count goes to
// device memory for the next bounce and, as a fraction of the batch, to pinned host memory for the next call's launch plan.
#define CP_THREADS 256
#define CP_ITEMS 16
#define CP_TILE (CP_THREADS * CP_ITEMS)
__global__ void __launch_bounds__(CP_THREADS) k_compact(const int32_t* __restrict__ live_in, const int32_t* __restrict__ count_in, int n_max,
                                                       const int32_t* __restrict__ status, int32_t* __restrict__ live_out, int32_t* count_out,
                                                       float* frac_out, float inv_batch, unsigned long long* tile_state, int32_t* tile_counter, int vec_ok) {
  __shared__ int s_tile, s_prefix;
  __shared__ int s_warp[CP_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_in = count_in ? *count_in : n_max;
  const int ntiles = (n_in + CP_TILE - 1) / CP_TILE;
  if (tid == 0) s_tile = atomicAdd(tile_counter, 1);  // tiles are numbered in the order they start: no deadlock in the look-back
  __syncthreads();
  const int tile = s_tile;
  if (tile >= ntiles) {
    if (n_in == 0 && tile == 0 && tid == 0) {
      *count_out = 0;
      if (frac_out) *frac_out = 0.f;
    }
    return;
  }
  const int base = tile * CP_TILE + tid * CP_ITEMS;
  int idx[CP_ITEMS];
  unsigned alive = 0;
  if (vec_ok && base + CP_ITEMS <= n_in) {  // 128-bit loads need 16-byte aligned arrays
    int st[CP_ITEMS];
    if (live_in) {
#pragma unroll
      for (int k = 0; k < CP_ITEMS; k += 4) {
        int4 v = *reinterpret_cast<const int4*>(live_in + base + k);
        idx[k] = v.x; idx[k + 1] = v.y; idx[k + 2] = v.z; idx[k + 3] = v.w;
      }
#pragma unroll
      for (int k = 0; k < CP_ITEMS; k++) st[k] = status[idx[k]];
    } else {
#pragma unroll
      for (int k = 0; k < CP_ITEMS; k += 4) {
        int4 v = *reinterpret_cast<const int4*>(status + base + k);
        st[k] = v.x; st[k + 1] = v.y; st[k + 2] = v.z; st[k + 3] = v.w;
        idx[k] = base + k; idx[k + 1] = base + k + 1; idx[k + 2] = base + k + 2; idx[k + 3] = base + k + 3;
      }
    }
#pragma unroll
    for (int k = 0; k < CP_ITEMS; k++) alive |= (st[k] == RBG_RUN ? 1u : 0u) << k;
  } else {
#pragma unroll
    for (int k = 0; k < CP_ITEMS; k++) {
      int i = base + k;
      idx[k] = 0;
      if (i < n_in) {
        idx[k] = live_in ? live_in[i] : i;
        alive |= (status[idx[k]] == RBG_RUN ? 1u : 0u) << k;
      }
    }
  }
  const int cnt = __popc(alive);
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int v = lane < CP_THREADS / 32 ? s_warp[lane] : 0, w = v;
#pragma unroll
    for (int o = 1; o < CP_THREADS / 32; o <<= 1) {
      int u = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += u;
    }
    if (lane < CP_THREADS / 32) s_warp[lane] = w - v;  // exclusive warp offsets
    const int total = __shfl_sync(0xffffffffu, w, CP_THREADS / 32 - 1);
    if (lane == 0 && tile > 0) atomicExch(&tile_state[tile], (1ull << 32) | (unsigned)total);  // AGGREGATE
    int excl = 0;
    volatile unsigned long long* ts = tile_state;
    for (int j0 = tile - 1; j0 >= 0; j0 -= 32) {  // warp-parallel look-back, 32 predecessors per round
      int j = j0 - lane;
      unsigned long long stw = 0;
      unsigned flag;
      do {
        stw = j >= 0 ? ts[j] : (2ull << 32);
        flag = (unsigned)(stw >> 32);
      } while (__any_sync(0xffffffffu, flag == 0));
      unsigned incl_mask = __ballot_sync(0xffffffffu, flag == 2);
      int upto = incl_mask ? __ffs(incl_mask) - 1 : 31;  // nearest predecessor that already holds an inclusive prefix
      int contrib = lane <= upto ? (int)(unsigned)stw : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
      excl += contrib;
      if (incl_mask) break;
    }
    if (lane == 0) {
      atomicExch(&tile_state[tile], (2ull << 32) | (unsigned)(excl + total));  // INCLUSIVE PREFIX
      s_prefix = excl;
      if (tile == ntiles - 1) {
        *count_out = excl + total;
        if (frac_out) { *frac_out = (float)(excl + total) * inv_batch; __threadfence_system(); }
      }
    }
  }
  __syncthreads();
  int pos = s_prefix + s_warp[warp] + incl - cnt;
#pragma unroll
  for (int k = 0; k < CP_ITEMS; k++)
    if (alive & (1u << k)) live_out[pos++] = idx[k];
}

// ---- coherence binning of the first bounce's index list.
// A beam whose neighbouring rays are far apart (random shooters) makes every warp walk 32 different pieces of geometry.  The
// rays stay where they are; the wavefront visits them through an index list, and that list is put into the order of the cell
// (64 x 64 grid over the two long sides of the box around the top volume's daughters, Z-order) in which each ray enters the
// box: a counting sort in three small kernels (cell + histogram, scan of the 4096 counters, scatter).  Results do not depend
// on the order of the list (outputs stay in input order, Philox streams are keyed by the ray id).
#define SB_BINS 4096
#define SB_THREADS 256
struct BinFrame {
  float lo[3], hi[3];
  int ax0, ax1;  // the two axes the cells are laid over
};
__device__ inline uint32_t sb_interleave(uint32_t a, uint32_t b) {  // Z-order of two 6-bit cell coordinates
  uint32_t r = 0;
#pragma unroll
  for (int k = 0; k < 6; k++) r |= ((a >> k) & 1u) << (2 * k) | ((b >> k) & 1u) << (2 * k + 1);
  return r;
}
__device__ inline uint32_t sb_cell(const BinFrame& f, double x, double y, double z, double dx, double dy, double dz) {
  const float p[3] = {(float)x, (float)y, (float)z}, v[3] = {(float)dx, (float)dy, (float)dz};
  float t0 = 0.f, t1 = 3.0e38f;
  bool hit = true;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    if (v[a] != 0.f) {
      float inv = 1.f / v[a], ta = (f.lo[a] - p[a]) * inv, tb = (f.hi[a] - p[a]) * inv;
      t0 = fmaxf(t0, fminf(ta, tb));
      t1 = fminf(t1, fmaxf(ta, tb));
    } else hit = hit && p[a] >= f.lo[a] && p[a] <= f.hi[a];
  }
  if (!(hit && t0 <= t1)) return SB_BINS - 1;  // rays that miss the box go last
  uint32_t q[2];
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const int a = k == 0 ? f.ax0 : f.ax1;
    float u = (p[a] + t0 * v[a] - f.lo[a]) / fmaxf(f.hi[a] - f.lo[a], 1e-30f);
    q[k] = (uint32_t)fminf(fmaxf(u * 64.f, 0.f), 63.f);
  }
  return sb_interleave(q[0], q[1]);
}
// cell of every ray + global histogram (shared-memory privatised); with `cells` == null it is the sampling pass of the
// coherence probe: counts the pairs of neighbouring input rays that enter through the same or an adjacent cell
__global__ void __launch_bounds__(SB_THREADS) k_bin_count(long long n, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                                                          const double* __restrict__ dx, const double* __restrict__ dy, const double* __restrict__ dz, BinFrame f,
                                                          uint16_t* __restrict__ cells, uint32_t* __restrict__ hist, unsigned long long* probe, long long per_block,
                                                          long long block_stride) {
  __shared__ uint32_t sh[SB_BINS];
  if (cells) {
    for (int b = threadIdx.x; b < SB_BINS; b += SB_THREADS) sh[b] = 0;
    __syncthreads();
  }
  const long long first = (long long)blockIdx.x * block_stride * per_block, last = min(n, first + per_block);
  unsigned long long coherent = 0, pairs = 0;
  for (long long i0 = first; i0 < last; i0 += SB_THREADS) {
    long long i = i0 + threadIdx.x;
    uint32_t c = SB_BINS - 1;
    if (i < last) {
      c = sb_cell(f, x[i], y[i], z[i], dx[i], dy[i], dz[i]);
      if (cells) {
        cells[i] = (uint16_t)c;
        atomicAdd(&sh[c], 1u);
      }
    }
    if (probe) {
      uint32_t prev = __shfl_up_sync(0xffffffffu, c, 1);
      bool have = i < last && (threadIdx.x & 31) != 0;
      // Z-order cells: the same cell, or cells that differ in the lowest bit pair only, count as neighbours
      bool coh = have && (c >> 2) == (prev >> 2);
      coherent += __popc(__ballot_sync(0xffffffffu, coh));
      pairs += __popc(__ballot_sync(0xffffffffu, have));
    }
  }
  if (probe && (threadIdx.x & 31) == 0 && pairs) {
    atomicAdd(&probe[0], coherent);
    atomicAdd(&probe[1], pairs);
  }
  if (cells) {
    __syncthreads();
    for (int b = threadIdx.x; b < SB_BINS; b += SB_THREADS)
      if (sh[b]) atomicAdd(&hist[b], sh[b]);
  }
}
// exclusive scan of the 4096 cell counters (one block): hist[] becomes the write cursor of every cell
__global__ void __launch_bounds__(1024) k_bin_scan(uint32_t* hist) {
  __shared__ uint32_t s_warp[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t v[4], sum = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) { v[k] = hist[4 * tid + k]; sum += v[k]; }
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = s_warp[lane], iw = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t u = __shfl_up_sync(0xffffffffu, iw, o);
      if (lane >= o) iw += u;
    }
    s_warp[lane] = iw - w;
  }
  __syncthreads();
  uint32_t run = s_warp[warp] + incl - sum;
#pragma unroll
  for (int k = 0; k < 4; k++) { hist[4 * tid + k] = run; run += v[k]; }
}
// scatter: every block counts the cells of its slice, reserves a range per cell from the global cursors and writes the ray
// indices there (the order inside a cell is not defined; nothing depends on it)
__global__ void __launch_bounds__(SB_THREADS) k_bin_scatter(long long n, const uint16_t* __restrict__ cells, uint32_t* cursor, int32_t* __restrict__ out,
                                                            long long per_block) {
  __shared__ uint32_t sh[SB_BINS];
  for (int b = threadIdx.x; b < SB_BINS; b += SB_THREADS) sh[b] = 0;
  __syncthreads();
  const long long first = (long long)blockIdx.x * per_block, last = min(n, first + per_block);
  for (long long i = first + threadIdx.x; i < last; i += SB_THREADS) atomicAdd(&sh[cells[i]], 1u);
  __syncthreads();
  for (int b = threadIdx.x; b < SB_BINS; b += SB_THREADS) {
    uint32_t c = sh[b];
    sh[b] = c ? atomicAdd(&cursor[b], c) : 0u;
  }
  __syncthreads();
  for (long long i = first + threadIdx.x; i < last; i += SB_THREADS) out[atomicAdd(&sh[cells[i]], 1u)] = (int32_t)i;
}

// largest npoints of a batch, written to pinned host memory (how many rows of a polyline record hold data)
__global__ void k_max_npoints(const int32_t* __restrict__ np, long long n, int32_t* out) {
  int m = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) m = max(m, np[i]);
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

// copies a device counter into pinned (UVA-mapped) host memory with a plain store: no copy-engine queueing
__global__ void k_publish_probe(const unsigned long long* d, unsigned long long* h) {
  h[0] = d[0];
  h[1] = d[1];
  __threadfence_system();
}
__global__ void k_publish32(const int32_t* d, int32_t* h) {
  *h = *d;
  __threadfence_system();
}

// ---- ARayShooter on device
__global__ void k_shoot(rbg_shoot_desc s, long long first, long long n, double* x, double* y, double* z, double* t, double* dx, double* dy, double* dz,
                        double* lambda) {
  long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  unsigned long long id = (unsigned long long)(first + j);
  Philox g;
  g.k0 = (uint32_t)s.seed; g.k1 = (uint32_t)(s.seed >> 32);
  g.id0 = (uint32_t)id; g.id1 = (uint32_t)(id >> 32);
  g.ndraw = 0x40000000u;
  double px = 0, py = 0;
  if (s.kind >= 4) {  // point sources at tr: RandomCone / RandomSphere / RandomSphericalCone (src/ARayShooter.cxx:240-392)
    double vx, vy, vz;
    if (s.kind == 4) {  // aimed at a random point of the disc of radius dx at z = dy, rotated
      double r = s.dx;
      do {
        px = -r + 2 * r * rng_uniform(g);
        py = -r + 2 * r * rng_uniform(g);
      } while (px * px + py * py > r * r);
      vx = s.rot[0] * px + s.rot[1] * py + s.rot[2] * s.dy;
      vy = s.rot[3] * px + s.rot[4] * py + s.rot[5] * s.dy;
      vz = s.rot[6] * px + s.rot[7] * py + s.rot[8] * s.dy;
    } else if (s.kind == 5) {  // TRandom::Sphere (isotropic)
      double a, b, r2;
      do {
        a = rng_uniform(g) - 0.5;
        b = rng_uniform(g) - 0.5;
        r2 = a * a + b * b;
      } while (r2 > 0.25);
      double scale = 8.0 * sqrt(0.25 - r2);
      vx = a * scale; vy = b * scale; vz = -1. + 8.0 * r2;
    } else {  // uniform in solid angle within dx degrees of +z, rotated
      double c0 = cos(s.dx * RB_PI / 180.), ran = c0 + (1. - c0) * rng_uniform(g), th = rb_acos(ran), phi = 2 * RB_PI * rng_uniform(g);
      double lx = sin(th) * cos(phi), ly = sin(th) * sin(phi), lz = cos(th);
      vx = s.rot[0] * lx + s.rot[1] * ly + s.rot[2] * lz;
      vy = s.rot[3] * lx + s.rot[4] * ly + s.rot[5] * lz;
      vz = s.rot[6] * lx + s.rot[7] * ly + s.rot[8] * lz;
    }
    double mag = sqrt(vx * vx + vy * vy + vz * vz);  // ARay's constructor normalises the direction
    if (mag > 0) { vx /= mag; vy /= mag; vz /= mag; }
    x[j] = s.tr[0]; y[j] = s.tr[1]; z[j] = s.tr[2]; t[j] = 0;
    dx[j] = vx; dy[j] = vy; dz[j] = vz;
    lambda[j] = s.lambda_min == s.lambda_max ? s.lambda_min : s.lambda_min + (s.lambda_max - s.lambda_min) * rng_uniform(g);
    return;
  }
  if (s.kind == 0) {
    long long i = (long long)id / s.ny, k = (long long)id % s.ny;
    double deltax = s.nx == 1 ? s.dx / 2 : s.dx / (s.nx - 1), deltay = s.ny == 1 ? s.dy / 2 : s.dy / (s.ny - 1);
    px = i * deltax - s.dx / 2;
    py = k * deltay - s.dy / 2;
  } else if (s.kind == 1) {
    px = -s.dx / 2 + s.dx * rng_uniform(g);
    py = -s.dy / 2 + s.dy * rng_uniform(g);
  } else if (s.kind == 2) {
    double rmax = s.dx;
    do {
      px = -rmax + 2 * rmax * rng_uniform(g);
      py = -rmax + 2 * rmax * rng_uniform(g);
    } while (sqrt(px * px + py * py) > rmax);
  } else {
    long long idx = (long long)id;
    if (idx > 0) {
      long long i = 0, acc = 1;
      while (idx >= acc + (long long)s.ny * (i + 1)) { acc += (long long)s.ny * (i + 1); i++; }
      long long k = idx - acc;
      double rr = s.dx * (i + 1) / s.nx, phi = 2 * RB_PI / s.ny / (i + 1) * k;
      px = rr * cos(phi);
      py = rr * sin(phi);
    }
  }
  // rot * (px,py,0) + tr ; dir = rot * dir, normalised
  double qx = s.rot[0] * px + s.rot[1] * py, qy = s.rot[3] * px + s.rot[4] * py, qz = s.rot[6] * px + s.rot[7] * py;
  double ndx = s.rot[0] * s.dir[0] + s.rot[1] * s.dir[1] + s.rot[2] * s.dir[2], ndy = s.rot[3] * s.dir[0] + s.rot[4] * s.dir[1] + s.rot[5] * s.dir[2],
         ndz = s.rot[6] * s.dir[0] + s.rot[7] * s.dir[1] + s.rot[8] * s.dir[2];
  double mag = sqrt(ndx * ndx + ndy * ndy + ndz * ndz);
  if (mag > 0) { ndx /= mag; ndy /= mag; ndz /= mag; }
  x[j] = s.tr[0] + qx; y[j] = s.tr[1] + qy; z[j] = s.tr[2] + qz; t[j] = 0;
  dx[j] = ndx; dy[j] = ndy; dz[j] = ndz;
  lambda[j] = s.lambda_min == s.lambda_max ? s.lambda_min : s.lambda_min + (s.lambda_max - s.lambda_min) * rng_uniform(g);
}

// ---- reducers
#define HIST_SMEM_BINS 8192
__global__ void k_hist2d(long long n, const double* __restrict__ x, const double* __restrict__ y, const int32_t* __restrict__ status, int sel, int nx,
                         double xmin, double xmax, int ny, double ymin, double ymax, unsigned long long* hist, int use_smem, double* stats, double x0, double y0) {
  __shared__ unsigned int sh[HIST_SMEM_BINS];
  double st[5] = {0, 0, 0, 0, 0};  // TH2 statistics of the in-range fills: sum w, x, y, x^2, y^2
  int nb = nx * ny;
  if (use_smem) {
    for (int b = threadIdx.x; b < nb; b += blockDim.x) sh[b] = 0;
    __syncthreads();
  }
  double sx = nx / (xmax - xmin), sy = ny / (ymax - ymin);
  // warp-uniform trip count: the lanes of a warp (neighbouring rays of a beam, which land in the same few bins of a PSF) merge
  // their fills — one atomic per distinct bin and warp instead of one per ray
  const int lane = threadIdx.x & 31;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < n; i0 += (long long)gridDim.x * blockDim.x) {
    const long long i = i0 + lane;
    int bin = -1;
    if (i < n && status[i] == sel) {
      double vx = x[i] - x0, vy = y[i] - y0;
      if (!(vx < xmin || !(vx < xmax) || vy < ymin || !(vy < ymax))) {
        int bx = (int)((vx - xmin) * sx), by = (int)((vy - ymin) * sy);
        bx = bx >= nx ? nx - 1 : bx;
        by = by >= ny ? ny - 1 : by;
        bin = bx + nx * by;
        st[0] += 1; st[1] += vx; st[2] += vy; st[3] += vx * vx; st[4] += vy * vy;
      }
    }
    const unsigned peers = __match_any_sync(0xffffffffu, bin);
    if (bin >= 0 && lane == __ffs(peers) - 1) {
      if (use_smem) atomicAdd(&sh[bin], (unsigned)__popc(peers));
      else atomicAdd(&hist[bin], (unsigned long long)__popc(peers));
    }
  }
  if (stats) {
#pragma unroll
    for (int k = 0; k < 5; k++) {
      for (int o = 16; o > 0; o >>= 1) st[k] += __shfl_down_sync(0xffffffffu, st[k], o);
      if ((threadIdx.x & 31) == 0 && st[k] != 0) atomicAdd(&stats[k], st[k]);
    }
  }
  if (use_smem) {
    __syncthreads();
    for (int b = threadIdx.x; b < nb; b += blockDim.x)
      if (sh[b]) atomicAdd(&hist[b], (unsigned long long)sh[b]);
  }
}

__global__ void k_moments(long long n, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ t,
                          const int32_t* __restrict__ status, int sel, double* moments, unsigned long long* counts) {
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};
  unsigned int c[6] = {0, 0, 0, 0, 0, 0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int s = status[i];
    if (s >= 0 && s < 6) c[s]++;
    if (s != sel) continue;
    double vx = x[i], vy = y[i], vt = t[i];
    acc[0] += 1; acc[1] += vx; acc[2] += vy; acc[3] += vx * vx; acc[4] += vy * vy; acc[5] += vt; acc[6] += vt * vt;
  }
#pragma unroll
  for (int k = 0; k < 7; k++)
    for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_down_sync(0xffffffffu, acc[k], o);
#pragma unroll
  for (int k = 0; k < 6; k++)
    for (int o = 16; o > 0; o >>= 1) c[k] += __shfl_down_sync(0xffffffffu, c[k], o);
  if ((threadIdx.x & 31) == 0) {
    for (int k = 0; k < 7; k++)
      if (acc[k] != 0) atomicAdd(&moments[k], acc[k]);
    for (int k = 0; k < 6; k++)
      if (c[k]) atomicAdd(&counts[k], (unsigned long long)c[k]);
  }
}

// ---- AGeoUtil::ContainmentRadius: kernel in rb_reducers.cu (compiled with -fmad=false)
int rb_launch_containment_u64(int nhist, const unsigned long long* hist, int nx, double xmin, double xmax, int ny, double ymin, double ymax, const double* stats,
                              double fraction, double* out, double* prefix, cudaStream_t st);
int rb_launch_containment_f64(int nhist, const double* hist, int nx, double xmin, double xmax, int ny, double ymin, double ymax, const double* stats,
                              double fraction, double* out, double* prefix, cudaStream_t st);

__global__ void k_tmm(DScene sc, int ml, long long n, const double* __restrict__ theta, const double* __restrict__ lambda, double* R, double* T) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double r, t;
  tmm_mixed(sc, ml, theta[i], lambda[i], r, t);
  R[i] = r;
  T[i] = t;
}

// AMultilayer::CoherentTMM / IncoherentTMM for arrays of (complex theta, lambda); pol 2 = mean of s and p
__global__ void k_tmm_general(DScene sc, int ml, int mode, int pol, int reverse, long long n, const double* __restrict__ th_re, const double* __restrict__ th_im,
                              const double* __restrict__ lambda, double* R, double* T) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const rbg_multilayer M = sc.multilayers[ml];
  Cx th = cx(th_re[i], th_im[i]);
  double r = 0, t = 0;
  for (int p = (pol == 2 ? 0 : pol); p <= (pol == 2 ? 1 : pol); p++) {
    double rp, tp;
    if (mode == 0) tmm_coherent_sub(sc, M.first, 0, M.n - 1, reverse != 0, p, th, lambda[i], rp, tp);
    else tmm_incoherent(sc, ml, p, th, lambda[i], rp, tp);
    r += rp;
    t += tp;
  }
  if (pol == 2) { r *= 0.5; t *= 0.5; }
  R[i] = r;
  T[i] = t;
}

// ------------------------------------------------------------------------------------------------ host: scene
#define RB_HOST_STREAMS 8   // chunks in flight on the host-buffer path of rbg_trace (one worker thread + stream + stage each)
#define RB_MAX_ROUNDS 24    // wavefront bounces planned ahead at most; whatever survives them is finished by one run-to-the-end launch
struct rbg_scene {
  int device = 0;
  int depth = 0;
  const rb_variant* variant = nullptr;
  DScene d;
  std::vector<void*> allocs;
  std::vector<std::string> node_names;
  // host staging of the host-buffer path
  void* stage[RB_HOST_STREAMS] = {};
  size_t stage_bytes[RB_HOST_STREAMS] = {};
  cudaStream_t streams[RB_HOST_STREAMS] = {};
  cudaStream_t copy_in = nullptr, copy_out = nullptr;  // host-buffer pipeline: all H2D on one stream, all D2H on another
  cudaEvent_t ev_in[RB_HOST_STREAMS] = {}, ev_cmp[RB_HOST_STREAMS] = {}, ev_out[RB_HOST_STREAMS] = {};
  int32_t* d_count = nullptr;
  int32_t* h_count = nullptr;  // pinned
  // launch plan of the wavefront, learned from the calls before (pinned, written by the device without any host wait):
  // h_frac[b] = share of the batch still running after bounce b; h_probe = {coherent, sampled} neighbour pairs of the last beam
  float* h_frac = nullptr;
  unsigned long long* h_probe = nullptr;
  std::atomic<int> calls{0};
  float root_lo[3] = {0, 0, 0}, root_hi[3] = {0, 0, 0};  // bounding box of the top volume's daughters
  bool has_root = false;
  int top_daughters = 0;
  void* pin = nullptr;  // pinned host staging of the small-batch path
  size_t pin_bytes = 0;
};

template <class T> static const T* upload(rbg_scene* s, const std::vector<T>& v) {
  if (v.empty()) return nullptr;
  void* p = nullptr;
  CK(cudaMalloc(&p, v.size() * sizeof(T)));
  s->allocs.push_back(p);
  CK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return (const T*)p;
}
template <class T> static const T* upload(rbg_scene* s, const T* src, size_t n) {
  if (!n || !src) return nullptr;
  return upload(s, std::vector<T>(src, src + n));
}

static void scene_free(rbg_scene* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  cudaDeviceSynchronize();  // traces enqueued on caller streams may still read the tables
  for (void* p : s->allocs) cudaFree(p);
  for (int i = 0; i < RB_HOST_STREAMS; i++) {
    if (s->stage[i]) cudaFree(s->stage[i]);
    if (s->streams[i]) cudaStreamDestroy(s->streams[i]);
    if (s->ev_in[i]) cudaEventDestroy(s->ev_in[i]);
    if (s->ev_cmp[i]) cudaEventDestroy(s->ev_cmp[i]);
    if (s->ev_out[i]) cudaEventDestroy(s->ev_out[i]);
  }
  if (s->copy_in) cudaStreamDestroy(s->copy_in);
  if (s->copy_out) cudaStreamDestroy(s->copy_out);
  if (s->d_count) cudaFree(s->d_count);
  if (s->h_count) cudaFreeHost(s->h_count);
  if (s->h_frac) cudaFreeHost(s->h_frac);
  if (s->h_probe) cudaFreeHost(s->h_probe);
  if (s->pin) cudaFreeHost(s->pin);
  delete s;
}

// the stream-ordered pool must keep its memory between calls: with the default release threshold (0) every synchronisation
// hands it back to the driver and the next cudaMallocAsync pays a device allocation
static void keep_pool(int device) {
  static std::atomic<unsigned> pool_kept{0};
  if (device >= 0 && device < 32 && !(pool_kept.fetch_or(1u << device) & (1u << device))) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      uint64_t keep = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
  }
}

// Scratch of the reducers comes from a pool of its own.  Taken from the device's default pool on a second stream it shares
// address space with the wavefront's GB-sized scratch blocks of the main stream, whose stream-ordered frees are still pending
// when the host runs ahead: now and then the allocator had to rearrange the pool, which blocks the host until the device is
// idle — a step of 62 ms came out at 95-330 ms in one bench run out of four (profiles/r2_summary.md).
static cudaMemPool_t small_pool(int device) {
  static std::mutex mu;
  static cudaMemPool_t pools[32] = {};
  if (device < 0 || device >= 32) throw Invalid("bad device index");
  std::lock_guard<std::mutex> lk(mu);
  if (!pools[device]) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    CK(cudaMemPoolCreate(&pools[device], &props));
    uint64_t keep = UINT64_MAX;
    CK(cudaMemPoolSetAttribute(pools[device], cudaMemPoolAttrReleaseThreshold, &keep));
  }
  return pools[device];
}

// ------------------------------------------------------------------------------------------------ host: trace driver
int rb_coop_search() {
  static const int on = getenv("RB_COOP") ? atoi(getenv("RB_COOP")) : 1;
  return on;
}
// launch-bound tuning units (make TUNE=1, rb_trace_t_*.cu) register extra instantiations from static initialisers
static std::vector<const rb_variant*>& extra_variants() { static std::vector<const rb_variant*> v; return v; }
int rb_register_variant(const rb_variant* v) { extra_variants().push_back(v); return 0; }
static const rb_variant* pick_variant(int depth, unsigned shapes, unsigned phys) {
  const char* force = getenv("RB_FORCE_GENERIC");
  const rb_variant* best = nullptr;
  int best_cost = 1 << 30;
  if (const char* want = getenv("RB_VARIANT")) {  // experiments: force a named (compatible) instantiation
    for (const rb_variant* v : rb_variants)
      if (!strcmp(v->name, want) && v->depth >= depth && !(shapes & ~v->shapes) && !(phys & ~v->phys)) return v;
    for (const rb_variant* v : extra_variants())
      if (!strcmp(v->name, want) && v->depth >= depth && !(shapes & ~v->shapes) && !(phys & ~v->phys)) return v;
  }
  for (const rb_variant* v : rb_variants) {
    if (v->depth < depth || (shapes & ~v->shapes) || (phys & ~v->phys)) continue;
    if (force && force[0] == '1' && v->shapes != RB_SHAPES_ALL) continue;
    int cost = __builtin_popcount(v->shapes) + __builtin_popcount(v->phys) + 4 * (v->depth - depth);
    if (cost < best_cost) { best_cost = cost; best = v; }
  }
  if (!best) throw NotSupported("boolean composite nesting deeper than 6 is not supported");
  return best;
}
static void launch_phase(const rb_variant* v, int phase, const DScene& sc, const DTraceParams& tp, const DRays& R, const DNavOut& N, const int32_t* live,
                         const int32_t* count, long long n_grid, long long n_max, cudaStream_t st) {
  if (n_max <= 0) return;
  ProfScope ps(st, 0);
  int rc = v->launch_phase(phase, sc, tp, R, N, live, count, n_grid, n_max, st);
  g_launches++;
  if (rc != 0) throw std::runtime_error(std::string("cuda: bounce kernel launch: ") + cudaGetErrorString((cudaError_t)rc));
}
static void launch_trace(const rb_variant* v, const DScene& sc, const DTraceParams& tp, const DRays& R, const int32_t* live, const int32_t* count,
                         long long n_grid, long long n_max, int init, int keep, cudaStream_t st) {
  if (n_max <= 0) return;
  ProfScope ps(st, 0);
  int rc = v->launch(sc, tp, R, live, count, n_grid, n_max, init, keep, st);
  g_launches++;
  if (rc != 0) throw std::runtime_error(std::string("cuda: k_trace launch: ") + cudaGetErrorString((cudaError_t)rc));
}

static size_t wavefront_scratch_bytes(long long n, bool with_record) {
  size_t ntile = (size_t)((n + CP_TILE - 1) / CP_TILE);
  // cur, ndraw, liveA, liveB, cells (uint16) + bounce counters + tile state + tile counter + cell histogram; and for the split bounce
  // the record between k_nav and k_shade (6 x 16 B + 8 B per ray)
  size_t b = 4 * ((size_t)n * 4 + 256) + ((size_t)n * 2 + 256) + 256 + ntile * 8 + 256 + 256 + SB_BINS * 4 + 512;
  if (with_record) b += 6 * ((size_t)n * 16 + 256) + ((size_t)n * 8 + 256);
  return b;
}

// Device-resident trace of n rays (n < 2^31) enqueued on stream st; never waits for the device.
// Single launch: every ray runs to its terminal status in registers.  Wavefront (large batches): one boundary step per launch
// over the compacted list of survivors.  The survivor counts stay on the device (k_compact -> k_trace), so the sequence of
// launches is planned ahead: `rounds` bounces (how many is learned from the shares of survivors the earlier calls on this
// scene reported), then one run-to-the-end launch over whatever is left.  A plan that is too short only moves work into the
// last launch; one that is too long adds a few empty launches.
static void trace_device(rbg_scene* s, const rbg_trace_opts* o, DRays R, long long n, unsigned long long id_offset, cudaStream_t st) {
  if (n <= 0) return;
  if (n > 0x7fffff00LL) throw Invalid("at most 2^31-256 rays per call on the device path; shard the batch");
  DTraceParams tp;
  tp.limit = o->limit > 0 ? o->limit : 100;
  tp.disable_fresnel = o->disable_fresnel;
  tp.quirks = o->quirks;
  // steps_per_launch: > 0 as given; 0 = auto (wavefront with one boundary step per bounce kernel for large
  // batches, where compaction pays for itself; a single launch for small ones); < 0 = single launch
  bool auto_wavefront = o->steps_per_launch == 0 && n >= 262144;
  // An instantiation of a few plain shapes (fused_bounce == 2) leaves the wavefront once the scene has shown that its rays end
  // within two boundary steps (the survivor shares of an earlier call, pinned memory): compaction buys nothing then —
  // SimpleParabolicTelescope runs 8 % faster at 9e6 rays and 32 % faster at 1e6 in one launch.  A scene of the same shapes that
  // keeps its rays bouncing stays with the wavefront.
  if (auto_wavefront && s->variant->fused_bounce == 2 && !R.hist.x && s->calls.load() > 0) {
    const long long tail = std::max<long long>(4096, n / 512);
    const float f1 = s->h_frac[1];
    if (f1 >= 0.f && (double)f1 * (double)n <= (double)tail) auto_wavefront = false;
  }
  tp.max_steps = o->steps_per_launch > 0 ? o->steps_per_launch : (auto_wavefront ? 1 : 0);
  if (R.hist.x) tp.max_steps = 0;  // the polyline record is written by the single-launch mode only
  tp.seed = o->seed;
  tp.ray_id_offset = id_offset;
  if (tp.max_steps <= 0) {
    R.cur = nullptr;
    R.ndraw = nullptr;
    launch_trace(s->variant, s->d, tp, R, nullptr, nullptr, n, n, 1, 0, st);
    return;
  }
  keep_pool(s->device);
  const int call = s->calls.fetch_add(1);
  char* base = nullptr;
  // RB_FUSED_BOUNCE=1 / 0 overrides the instantiation's choice between one k_trace launch per bounce and k_nav + k_shade
  static const int fused_env = getenv("RB_FUSED_BOUNCE") ? atoi(getenv("RB_FUSED_BOUNCE")) : -1;
  const bool fused = fused_env >= 0 ? fused_env != 0 : s->variant->fused_bounce != 0;
  CK(cudaMallocAsync((void**)&base, wavefront_scratch_bytes(n, !fused), st));
  try {
    size_t ntile = (size_t)((n + CP_TILE - 1) / CP_TILE);
    size_t off = 0;
    auto take = [&](size_t bytes) { void* p = base + off; off += (bytes + 255) & ~size_t(255); return p; };
    R.cur = (int32_t*)take(n * 4);
    R.ndraw = (uint32_t*)take(n * 4);
    int32_t* liveA = (int32_t*)take(n * 4);
    int32_t* liveB = (int32_t*)take(n * 4);
    uint16_t* cells = (uint16_t*)take(n * 2);
    DNavOut N;
    memset(&N, 0, sizeof(N));
    if (!fused) {
      N.pxy = (double2*)take(n * 16); N.pzs = (double2*)take(n * 16);
      N.loc = (int4*)take(n * 16); N.hit = (int4*)take(n * 16); N.vis = (int4*)take(2 * n * 16);
      N.ent = (double*)take(n * 8);
    }
    N.n = n;
    int32_t* d_counts = (int32_t*)take((RB_MAX_ROUNDS + 2) * 4);
    unsigned long long* tile_state = (unsigned long long*)take(ntile * 8);
    int32_t* tile_counter = (int32_t*)take(4);
    uint32_t* cell_hist = (uint32_t*)take(SB_BINS * 4);
    unsigned long long* d_probe = (unsigned long long*)take(16);
    const int32_t* live = nullptr;
    const int32_t* count = nullptr;
    // ---- coherence binning of the index list (first bounce).  Only scenes with many daughters under the top volume can
    // scatter a warp over different geometry, and only beams whose input order is not already spatially coherent need it.
    // The verdict on the beam comes from a sampled probe; it is read from the call before (pinned memory, no wait) — the
    // first call on a scene waits for its own probe once.
    static const int sort_mode = getenv("RB_SORT") ? atoi(getenv("RB_SORT")) : -1;  // -1 auto, 0 never, 1 always
    if (s->has_root && sort_mode != 0 && (sort_mode == 1 || s->top_daughters >= 32)) {
      BinFrame f;
      int ax[3] = {0, 1, 2};
      std::sort(ax, ax + 3, [&](int a, int b) { return s->root_hi[a] - s->root_lo[a] > s->root_hi[b] - s->root_lo[b]; });
      for (int k = 0; k < 3; k++) { f.lo[k] = s->root_lo[k]; f.hi[k] = s->root_hi[k]; }
      f.ax0 = std::min(ax[0], ax[1]);
      f.ax1 = std::max(ax[0], ax[1]);
      bool do_sort = sort_mode == 1;
      if (!do_sort) {
        // sampled probe of this beam: every stride-th run of 256 rays
        long long runs = (n + SB_THREADS - 1) / SB_THREADS;
        long long stride = std::max<long long>(1, runs / 512);
        CK(cudaMemsetAsync(d_probe, 0, 16, st));
        k_bin_count<<<(unsigned)(runs / stride), SB_THREADS, 0, st>>>(n, R.x, R.y, R.z, R.dx, R.dy, R.dz, f, nullptr, nullptr, d_probe, SB_THREADS, stride);
        if (call == 0) {
          k_publish_probe<<<1, 1, 0, st>>>(d_probe, s->h_probe);
          CK(cudaStreamSynchronize(st));
        }
        unsigned long long coh = s->h_probe[0], pairs = s->h_probe[1];
        do_sort = pairs > 0 && (double)coh < 0.5 * (double)pairs;
        if (call != 0) k_publish_probe<<<1, 1, 0, st>>>(d_probe, s->h_probe);  // plain stores into pinned memory, for the next call
        g_launches += 2;
        CK(cudaGetLastError());
      }
      if (do_sort) {
        ProfScope ps(st, 1);
        long long blocks = std::min<long long>((n + 4095) / 4096, 148 * 8);
        long long per_block = ((n + blocks - 1) / blocks + SB_THREADS - 1) / SB_THREADS * SB_THREADS;
        blocks = (n + per_block - 1) / per_block;
        CK(cudaMemsetAsync(cell_hist, 0, SB_BINS * 4, st));
        k_bin_count<<<(unsigned)blocks, SB_THREADS, 0, st>>>(n, R.x, R.y, R.z, R.dx, R.dy, R.dz, f, cells, cell_hist, nullptr, per_block, 1);
        k_bin_scan<<<1, 1024, 0, st>>>(cell_hist);
        k_bin_scatter<<<(unsigned)blocks, SB_THREADS, 0, st>>>(n, cells, cell_hist, liveA, per_block);
        g_launches += 3;
        CK(cudaGetLastError());
        live = liveA;
      }
    }
    // ---- launch plan
    const long long tail = std::max<long long>(4096, n / 512);  // survivors below this finish in the run-to-the-end launch
    float frac[RB_MAX_ROUNDS];
    for (int b = 0; b < RB_MAX_ROUNDS; b++) frac[b] = call == 0 ? -1.f : s->h_frac[b];
    int rounds = RB_MAX_ROUNDS;
    for (int b = 0; b < RB_MAX_ROUNDS; b++)
      if (frac[b] >= 0.f && (double)frac[b] * (double)n <= (double)tail) { rounds = b + 1; break; }
    if (tp.limit - 1 < rounds) rounds = std::max(1, tp.limit - 1);  // a ray takes at most limit - 1 steps
    static const int max_rounds_env = getenv("RB_MAX_ROUNDS") ? atoi(getenv("RB_MAX_ROUNDS")) : RB_MAX_ROUNDS;
    rounds = std::min(rounds, std::max(1, max_rounds_env));
    auto estimate = [&](int b) -> long long {  // rays expected to be alive when bounce b starts
      if (b == 0) return n;
      float fr = frac[b - 1];
      if (fr < 0.f) return n;
      return std::min<long long>(n, (long long)((double)fr * 1.25 * (double)n) + 4096);
    };
    const float inv_batch = 1.f / (float)n;
    const int vec_ok = (reinterpret_cast<uintptr_t>(R.status) & 15) == 0;
    const int tiles = (int)ntile;
    for (int b = 0; b < rounds; b++) {
      if (fused) launch_trace(s->variant, s->d, tp, R, live, count, estimate(b), n, b == 0, 1, st);
      else {
        // first bounce: k_nav locates the start points (InitTrack), takes the step into the top volume for rays shot from outside
        // it, and reads the input arrays; k_shade starts from the input arrays too
        launch_phase(s->variant, b == 0 ? 3 : 1, s->d, tp, R, N, live, count, estimate(b), n, st);
        launch_phase(s->variant, b == 0 ? 4 : 2, s->d, tp, R, N, live, count, estimate(b), n, st);
      }
      CK(cudaMemsetAsync(tile_state, 0, (size_t)tiles * 8, st));
      CK(cudaMemsetAsync(tile_counter, 0, 4, st));
      int32_t* out = (live == liveA) ? liveB : liveA;
      {
        ProfScope ps(st, 1);
        // the grid has to cover the true count, which only the device knows: one tile per 4096 rays of the batch, the surplus
        // tiles leave at once
        k_compact<<<tiles, CP_THREADS, 0, st>>>(live, count, (int)n, R.status, out, d_counts + b, s->h_frac + b, inv_batch, tile_state, tile_counter, vec_ok);
        g_launches++;
        CK(cudaGetLastError());
      }
      live = out;
      count = d_counts + b;
    }
    DTraceParams tl = tp;
    tl.max_steps = 0;
    launch_trace(s->variant, s->d, tl, R, live, count, std::max<long long>(estimate(rounds), tail), n, 0, 1, st);
  } catch (...) {
    cudaFreeAsync(base, st);
    throw;
  }
  CK(cudaFreeAsync(base, st));
}

// ================================================================================================ several GPUs, one process
struct rbg_multi {
  std::vector<rbg_scene*> scenes;  // one replica per device
  std::vector<int> devices;
  bool peer = false;               // device 0 can be written by every other device
};
__global__ void k_add_u64(unsigned long long* dst, const unsigned long long* src, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] += src[i];
}

// contiguous ranges, the last device takes the remainder (src/AOpticsManager.cxx:533-541)
static void multi_range(long long n, int k, int G, long long& b, long long& e) {
  long long chunk = n / G;
  b = chunk * k;
  e = k == G - 1 ? n : chunk * (k + 1);
}
template <class F> static void multi_run(rbg_multi* m, F f) {
  const int G = (int)m->scenes.size();
  std::vector<std::string> errs(G);
  std::vector<int> rcs(G, RBG_OK);
  std::vector<std::thread> th;
  for (int k = 0; k < G; k++)
    th.emplace_back([&, k] {
      rcs[k] = guard([&] { f(k); });
      if (rcs[k] != RBG_OK) errs[k] = g_err;  // g_err is thread-local: carry it over
    });
  for (auto& t : th) t.join();
  for (int k = 0; k < G; k++)
    if (rcs[k] != RBG_OK) throw std::runtime_error(errs[k].rfind("cuda", 0) == 0 ? errs[k] : "device " + std::to_string(m->devices[k]) + ": " + errs[k]);
}


// ================================================================================================ C ABI
#pragma GCC visibility push(default)
extern "C" {

int rbg_abi_version(void) { return RBG_ABI_VERSION; }
const char* rbg_last_error(void) { return g_err.c_str(); }
int rbg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
int64_t rbg_launch_count(void) { return g_launches.load(); }
int rbg_profile_enable(int on) {
  g_profile = on;
  return RBG_OK;
}
int rbg_profile_read(double* bounce_ms, int64_t* bounce_launches, double* compact_ms, int64_t* compact_launches) {
  return guard([&] {
    std::vector<ProfEvt> evts;
    {
      std::lock_guard<std::mutex> lk(g_prof_mutex);
      evts.swap(g_prof_events);
    }
    for (auto& e : evts) {
      CK(cudaEventSynchronize(e.b));
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e.a, e.b));
      g_prof_ms[e.kind] += ms;
      g_prof_n[e.kind]++;
      cudaEventDestroy(e.a);
      cudaEventDestroy(e.b);
    }
    if (bounce_ms) *bounce_ms = g_prof_ms[0];
    if (bounce_launches) *bounce_launches = g_prof_n[0];
    if (compact_ms) *compact_ms = g_prof_ms[1];
    if (compact_launches) *compact_launches = g_prof_n[1];
    g_prof_ms[0] = g_prof_ms[1] = 0;
    g_prof_n[0] = g_prof_n[1] = 0;
  });
}

int rbg_scene_create(const rbg_scene_desc* D, int device, rbg_scene** out) {
  rbg_scene* s = nullptr;
  int rc = guard([&] {
    if (!out) throw Invalid("null output handle");
    *out = nullptr;
    validate_desc(D);
    int ndev = rbg_device_count();
    if (ndev <= 0) throw std::runtime_error("cuda: no CUDA device available — the tracer has no CPU fallback");
    if (device < 0 || device >= ndev) throw Invalid("bad device index");
    SceneBuilder B;
    B.D = D;
    B.build_shapes();
    if (D->top_volume >= 0) B.flatten(D->top_volume, mat_identity(), -1, 0, std::string(D->names + D->volumes[D->top_volume].name) + "_1");
    int depth = scene_depth_needed(B);
    if (depth > 6) throw NotSupported("boolean composite nesting deeper than 6 is not supported");
    CK(cudaSetDevice(device));
    s = new rbg_scene;
    s->device = device;
    s->depth = depth;
    unsigned need_shapes = 0, need_phys = 0;
    scene_features_needed(D, need_shapes, need_phys);
    s->variant = pick_variant(depth, need_shapes, need_phys);
    s->node_names = B.names;
    memset(&s->d, 0, sizeof(s->d));
    s->d.nodes = upload(s, B.nodes);
    s->d.bvh = upload(s, B.bvh);
    s->d.boxes = upload(s, B.boxes);
    s->d.shapes = upload(s, B.shapes);
    s->d.dpar = upload(s, B.dpar);
    s->d.mats = upload(s, B.mats);
    s->d.volumes = upload(s, D->volumes, D->nvolumes);
    s->d.borders = upload(s, D->borders, D->nborders);
    s->d.graphs = upload(s, D->graphs, D->ngraphs);
    s->d.gx = upload(s, D->gx, D->ngpts);
    s->d.gy = upload(s, D->gy, D->ngpts);
    s->d.th2 = upload(s, D->th2, D->nth2);
    s->d.th2v = upload(s, D->th2v, D->nth2v);
    s->d.indices = upload(s, D->indices, D->nindices);
    s->d.mirrors = upload(s, D->mirrors, D->nmirrors);
    s->d.focals = upload(s, D->focals, D->nfocals);
    s->d.multilayers = upload(s, D->multilayers, D->nmultilayers);
    s->d.layers = upload(s, D->layers, D->nlayers);
    s->d.graph2d = upload(s, D->graph2d, D->ngraph2d);
    s->d.tri = upload(s, D->tri, 3 * (size_t)D->ntri);
    s->d.g2x = upload(s, D->g2x, D->ng2pts);
    s->d.g2y = upload(s, D->g2y, D->ng2pts);
    s->d.g2z = upload(s, D->g2z, D->ng2pts);
    s->d.nnodes = (int)B.nodes.size();
    s->d.nbvh = (int)B.bvh.size();
    s->d.nshapes = (int)B.shapes.size();
    s->d.ndpar = (int)B.dpar.size();
    s->d.nmats = (int)B.mats.size();
    s->d.top_shape = D->top_volume >= 0 ? D->volumes[D->top_volume].shape : -1;
    s->d.top_leaf = s->d.top_shape >= 0 ? B.leaf_kind(s->d.top_shape) : RB_LEAF_GENERIC;
    s->d.has_many = (need_phys & RB_PH_OVERLAP) != 0;
    CK(cudaMalloc((void**)&s->d_count, RB_HOST_STREAMS * sizeof(int32_t)));
    CK(cudaMallocHost((void**)&s->h_count, RB_HOST_STREAMS * sizeof(int32_t)));
    CK(cudaMallocHost((void**)&s->h_frac, RB_MAX_ROUNDS * sizeof(float)));
    CK(cudaMallocHost((void**)&s->h_probe, 2 * sizeof(unsigned long long)));
    for (int b = 0; b < RB_MAX_ROUNDS; b++) s->h_frac[b] = -1.f;
    s->h_probe[0] = s->h_probe[1] = 0;
    if (!B.nodes.empty() && B.nodes[0].bvh_count > 0) {
      const DBvh& rb = B.bvh[B.nodes[0].bvh_first];
      for (int k = 0; k < 3; k++) { s->root_lo[k] = rb.lo[k]; s->root_hi[k] = rb.hi[k]; }
      s->has_root = true;
      for (const DNode& nd : B.nodes) s->top_daughters += nd.mother == 0 ? 1 : 0;
    }
    // the CSG call chain is not inlined: give it stack
    size_t want = 8192, have = 0;
    CK(cudaDeviceGetLimit(&have, cudaLimitStackSize));
    if (have < want) CK(cudaDeviceSetLimit(cudaLimitStackSize, want));
    *out = s;
    s = nullptr;
  });
  if (s) scene_free(s);
  return rc;
}

int rbg_scene_destroy(rbg_scene* s) {
  return guard([&] { scene_free(s); });
}
int rbg_scene_num_nodes(const rbg_scene* s) { return s ? (int)s->node_names.size() : 0; }
const char* rbg_scene_kernel_variant(const rbg_scene* s) { return s && s->variant ? s->variant->name : ""; }
const char* rbg_scene_node_name(const rbg_scene* s, int node) {
  if (!s || node < 0 || node >= (int)s->node_names.size()) return "";
  return s->node_names[node].c_str();
}

int rbg_trace(rbg_scene* s, const rbg_trace_opts* o, const rbg_rays* rays, void* stream) { return rbg_trace_history(s, o, rays, nullptr, stream); }

int rbg_trace_history(rbg_scene* s, const rbg_trace_opts* o, const rbg_rays* rays, const rbg_history* hist, void* stream) {
  return guard([&] {
    if (!s || !o || !rays) throw Invalid("null argument");
    const int hpts = hist ? hist->max_points : 0;
    if (hpts < 0) throw Invalid("negative history depth");
    if (hpts > 0 && (!hist->hx || !hist->hy || !hist->hz || !hist->ht || !hist->hnode)) throw Invalid("null history array");
    if (s->d.top_shape < 0) throw Invalid("scene has no top volume");
    if (rays->n <= 0) return;
    if (!rays->x || !rays->y || !rays->z || !rays->t || !rays->dx || !rays->dy || !rays->dz || !rays->lambda || !rays->ox || !rays->oy || !rays->oz ||
        !rays->ot || !rays->odx || !rays->ody || !rays->odz || !rays->status || !rays->last_node || !rays->npoints)
      throw Invalid("null ray array");
    CK(cudaSetDevice(s->device));
    if (rays->on_device) {
      cudaStream_t st = (cudaStream_t)stream;
      DRays R;
      R.x = rays->x; R.y = rays->y; R.z = rays->z; R.t = rays->t; R.dx = rays->dx; R.dy = rays->dy; R.dz = rays->dz; R.lambda = rays->lambda;
      R.ox = rays->ox; R.oy = rays->oy; R.oz = rays->oz; R.ot = rays->ot; R.odx = rays->odx; R.ody = rays->ody; R.odz = rays->odz;
      R.status = rays->status; R.last_node = rays->last_node; R.npoints = rays->npoints;
      R.cur = nullptr; R.ndraw = nullptr;
      memset(&R.hist, 0, sizeof(R.hist));
      if (hpts > 0) {
        R.hist.x = hist->hx; R.hist.y = hist->hy; R.hist.z = hist->hz; R.hist.t = hist->ht; R.hist.node = hist->hnode;
        R.hist.stride = rays->n;
        R.hist.max_points = hpts;
      }
      trace_device(s, o, R, rays->n, o->ray_id_offset, st);
      return;
    }
    // small batches (the tutorial / MINUIT-loop regime, thousands of calls of 1e3..1e5 rays): latency matters, not bandwidth.
    // Inputs are packed into one pinned staging block (1 H2D), results come back as one block (1 D2H), and of a polyline
    // record only the rows that hold data are fetched.  18 small pageable copies would cost more than the trace itself.
    if (rays->n < 262144) {
      const long long n = rays->n;
      if (!s->streams[0]) CK(cudaStreamCreateWithFlags(&s->streams[0], cudaStreamNonBlocking));
      cudaStream_t st = s->streams[0];
      const size_t in_b = (size_t)n * 64, out_b = (size_t)n * 68, hist_b = (size_t)n * hpts * 36;
      const size_t dev_b = in_b + out_b + 512 + hist_b + 1024;
      if (s->stage_bytes[0] < dev_b) {
        if (s->stage[0]) CK(cudaFree(s->stage[0]));
        s->stage[0] = nullptr;
        s->stage_bytes[0] = 0;
        CK(cudaMalloc(&s->stage[0], dev_b));
        s->stage_bytes[0] = dev_b;
      }
      const size_t pin_b = in_b + out_b + hist_b + 256;
      if (s->pin_bytes < pin_b) {
        if (s->pin) CK(cudaFreeHost(s->pin));
        s->pin = nullptr;
        s->pin_bytes = 0;
        CK(cudaMallocHost(&s->pin, pin_b + pin_b / 2));
        s->pin_bytes = pin_b + pin_b / 2;
      }
      char* base = (char*)s->stage[0];
      char* pin = (char*)s->pin;
      const double* hin[8] = {rays->x, rays->y, rays->z, rays->t, rays->dx, rays->dy, rays->dz, rays->lambda};
      double* hout[7] = {rays->ox, rays->oy, rays->oz, rays->ot, rays->odx, rays->ody, rays->odz};
      int32_t* hiout[3] = {rays->status, rays->last_node, rays->npoints};
      for (int a = 0; a < 8; a++) memcpy(pin + (size_t)a * n * 8, hin[a], (size_t)n * 8);
      CK(cudaMemcpyAsync(base, pin, in_b, cudaMemcpyHostToDevice, st));
      DRays R;
      double* din = (double*)base;
      double* dout = (double*)(base + in_b);
      int32_t* diout = (int32_t*)(base + in_b + (size_t)n * 56);
      R.x = din; R.y = din + n; R.z = din + 2 * n; R.t = din + 3 * n; R.dx = din + 4 * n; R.dy = din + 5 * n; R.dz = din + 6 * n; R.lambda = din + 7 * n;
      R.ox = dout; R.oy = dout + n; R.oz = dout + 2 * n; R.ot = dout + 3 * n; R.odx = dout + 4 * n; R.ody = dout + 5 * n; R.odz = dout + 6 * n;
      R.status = diout; R.last_node = diout + n; R.npoints = diout + 2 * n;
      R.cur = nullptr; R.ndraw = nullptr;
      memset(&R.hist, 0, sizeof(R.hist));
      char* hb = base + ((in_b + out_b + 511) & ~size_t(255));
      if (hpts > 0) {
        size_t plane = (size_t)n * hpts * 8;
        R.hist.x = (double*)hb; R.hist.y = (double*)(hb + plane); R.hist.z = (double*)(hb + 2 * plane); R.hist.t = (double*)(hb + 3 * plane);
        R.hist.node = (int32_t*)(hb + 4 * plane);
        R.hist.stride = n;
        R.hist.max_points = hpts;
      }
      trace_device(s, o, R, n, o->ray_id_offset, st);
      if (hpts > 0) {
        CK(cudaMemsetAsync(s->d_count + 1, 0, 4, st));
        k_max_npoints<<<(unsigned)std::min<long long>((n + 255) / 256, 296), 256, 0, st>>>(R.npoints, n, s->d_count + 1);
        k_publish32<<<1, 1, 0, st>>>(s->d_count + 1, s->h_count + 1);
        g_launches += 2;
      }
      CK(cudaMemcpyAsync(pin + in_b, dout, out_b, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      for (int a = 0; a < 7; a++) memcpy(hout[a], pin + in_b + (size_t)a * n * 8, (size_t)n * 8);
      for (int a = 0; a < 3; a++) memcpy(hiout[a], pin + in_b + (size_t)n * 56 + (size_t)a * n * 4, (size_t)n * 4);
      if (hpts > 0) {
        const int rows = std::min<int>(std::max<int>(s->h_count[1], 1), hpts);  // rows k < max npoints hold data
        char* ph = pin + in_b + out_b;
        size_t plane = (size_t)n * hpts * 8, used = (size_t)n * rows * 8;
        for (int a = 0; a < 4; a++) CK(cudaMemcpyAsync(ph + a * used, hb + a * plane, used, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(ph + 4 * used, hb + 4 * plane, used / 2, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        double* hd[4] = {hist->hx, hist->hy, hist->hz, hist->ht};
        for (int a = 0; a < 4; a++) memcpy(hd[a], ph + a * used, used);
        memcpy(hist->hnode, ph + 4 * used, used / 2);
      }
      return;
    }
    // host buffers: chunked H2D -> trace -> D2H.  Up to RB_HOST_STREAMS chunks are in flight, each on its own stream
    // with its own stage buffer, driven by its own host thread: the wavefront loop reads the survivor count back
    // after every bounce, and that wait must not keep the other chunks' copies from being enqueued.  In steady state
    // the H2D of chunk c+1, the bounce kernels of chunk c and the D2H of chunk c-1 overlap (PCIe is full duplex).
    // chunk size: small enough that a 1e7-ray call has a steady state (H2D, kernels and D2H of different chunks
    // overlapping), large enough that a bounce launch still fills the 148 SMs many times over
    bool pinned = hpts == 0 && getenv("RB_HOST_THREADS_PIPELINE") == nullptr;
    {
      const void* probe[4] = {rays->x, rays->lambda, rays->ox, rays->status};
      for (int a = 0; a < 4 && pinned; a++) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, probe[a]) != cudaSuccess || at.type != cudaMemoryTypeHost) pinned = false;
        cudaGetLastError();
      }
    }
    // chunk size: small enough that a 1e7-ray call has a steady state (H2D, kernels and D2H of different chunks overlapping) and
    // a short fill and drain, large enough that a bounce launch still fills the 148 SMs.  Measured on the nine 1.1e7-ray calls
    // of the bench (profiles/r2_summary.md): pinned pipeline 256 Ki rays x 8 stage buffers, no ramp; threaded pipeline 1 Mi x 6.
    static const long long CH_env = getenv("RB_HOST_CHUNK") ? std::max(4096LL, atoll(getenv("RB_HOST_CHUNK"))) : 0;
    const long long CH = CH_env ? CH_env : (pinned ? (1LL << 18) : (1LL << 20));
    long long chunk = std::min<long long>(rays->n, CH);
    size_t per_ray = 8 * 8 + 7 * 8 + 3 * 4;  // in + out
    size_t bytes = (size_t)chunk * per_ray + 4096 + (size_t)chunk * hpts * 36 + 1024;
    // chunk list: sizes ramp up from chunk/8 at the start and down again at the end, so that the un-overlapped
    // pipeline fill (first H2D) and drain (last D2H) of the call are short
    std::vector<std::pair<long long, long long>> chunks;  // (first ray, count)
    {
      long long head = 0, tail = rays->n;
      std::vector<std::pair<long long, long long>> back;
      static const int ramp_env = getenv("RB_HOST_RAMP") ? atoi(getenv("RB_HOST_RAMP")) : 0;  // first / last chunk = chunk / ramp
      const int ramp = ramp_env ? ramp_env : (pinned ? 1 : 8);
      for (long long sz = std::max<long long>(chunk / std::max(1, ramp), 65536); sz < chunk && tail - head > 4 * chunk; sz *= 2) {
        chunks.emplace_back(head, sz);
        head += sz;
        back.emplace_back(tail - sz, sz);
        tail -= sz;
      }
      while (head < tail) {
        long long m = std::min(chunk, tail - head);
        chunks.emplace_back(head, m);
        head += m;
      }
      chunks.insert(chunks.end(), back.rbegin(), back.rend());
    }
    long long nchunks = (long long)chunks.size();
    static const int streams_env = getenv("RB_HOST_NSTREAMS") ? std::min(RB_HOST_STREAMS, std::max(1, atoi(getenv("RB_HOST_NSTREAMS")))) : 0;
    const int max_streams = streams_env ? streams_env : (pinned ? 8 : 6);
    int nst = (int)std::min<long long>(nchunks, max_streams);
    for (int k = 0; k < nst; k++) {
      if (!s->streams[k]) CK(cudaStreamCreateWithFlags(&s->streams[k], cudaStreamNonBlocking));
      if (s->stage_bytes[k] < bytes) {
        if (s->stage[k]) CK(cudaFree(s->stage[k]));
        s->stage[k] = nullptr;
        s->stage_bytes[k] = 0;
        CK(cudaMalloc(&s->stage[k], bytes));
        s->stage_bytes[k] = bytes;
      }
    }
    // Pinned host arrays: three-stage pipeline driven by this thread alone (nothing below waits for the device until the end).
    // Every H2D copy goes to one stream, in chunk order, every D2H copy to another, the traces to one stream per stage
    // buffer; events chain copy-in -> trace -> copy-out per chunk and free a stage buffer for the chunk `nst` later.  Both copy
    // engines always have their next transfer queued, so the two PCIe directions stay busy together for the whole call instead
    // of idling whenever a stream that owns a chunk end to end (the scheme below) sits in one of its other two stages
    // (profiles/r2_summary.md: 59 -> 7x GB/s H2D + D2H per 1.1e7-ray call).
    if (pinned) {
      if (!s->copy_in) CK(cudaStreamCreateWithFlags(&s->copy_in, cudaStreamNonBlocking));
      if (!s->copy_out) CK(cudaStreamCreateWithFlags(&s->copy_out, cudaStreamNonBlocking));
      for (int k = 0; k < nst; k++) {
        if (!s->ev_in[k]) CK(cudaEventCreateWithFlags(&s->ev_in[k], cudaEventDisableTiming));
        if (!s->ev_cmp[k]) CK(cudaEventCreateWithFlags(&s->ev_cmp[k], cudaEventDisableTiming));
        if (!s->ev_out[k]) CK(cudaEventCreateWithFlags(&s->ev_out[k], cudaEventDisableTiming));
      }
      const double* hin[8] = {rays->x, rays->y, rays->z, rays->t, rays->dx, rays->dy, rays->dz, rays->lambda};
      double* hout[7] = {rays->ox, rays->oy, rays->oz, rays->ot, rays->odx, rays->ody, rays->odz};
      int32_t* hiout[3] = {rays->status, rays->last_node, rays->npoints};
      // columns laid out at one constant pitch (the rows of one matrix): one 2-D copy per chunk and direction instead of 8 + 10
      auto pitch_of = [](const char* const* p, int cnt, long long min_bytes) -> long long {
        const long long d = p[1] - p[0];
        if (d < min_bytes) return 0;
        for (int a = 2; a < cnt; a++)
          if (p[a] - p[a - 1] != d) return 0;
        return d;
      };
      const char* pin_[8]; const char* pout_[7]; const char* piout_[3];
      for (int a = 0; a < 8; a++) pin_[a] = (const char*)hin[a];
      for (int a = 0; a < 7; a++) pout_[a] = (const char*)hout[a];
      for (int a = 0; a < 3; a++) piout_[a] = (const char*)hiout[a];
      static const bool no2d = getenv("RB_HOST_NO2D") != nullptr;
      const long long in_pitch = no2d ? 0 : pitch_of(pin_, 8, rays->n * 8), out_pitch = no2d ? 0 : pitch_of(pout_, 7, rays->n * 8),
                      iout_pitch = no2d ? 0 : pitch_of(piout_, 3, rays->n * 4);
      for (long long ci = 0; ci < nchunks; ci++) {
        const int k = (int)(ci % nst);
        const long long b = chunks[ci].first, m = chunks[ci].second;
        char* base = (char*)s->stage[k];
        double* din[8];
        double* dout[7];
        int32_t* diout[3];
        size_t off = 0;
        for (int a = 0; a < 8; a++) { din[a] = (double*)(base + off); off += (size_t)chunk * 8; }
        for (int a = 0; a < 7; a++) { dout[a] = (double*)(base + off); off += (size_t)chunk * 8; }
        for (int a = 0; a < 3; a++) { diout[a] = (int32_t*)(base + off); off += (size_t)chunk * 4; }
        if (ci >= nst) CK(cudaStreamWaitEvent(s->copy_in, s->ev_out[k], 0));  // the stage buffer's previous chunk has left
        if (in_pitch > 0) CK(cudaMemcpy2DAsync(din[0], (size_t)chunk * 8, hin[0] + b, (size_t)in_pitch, (size_t)m * 8, 8, cudaMemcpyHostToDevice, s->copy_in));
        else
          for (int a = 0; a < 8; a++) CK(cudaMemcpyAsync(din[a], hin[a] + b, (size_t)m * 8, cudaMemcpyHostToDevice, s->copy_in));
        CK(cudaEventRecord(s->ev_in[k], s->copy_in));
        CK(cudaStreamWaitEvent(s->streams[k], s->ev_in[k], 0));
        DRays R;
        R.x = din[0]; R.y = din[1]; R.z = din[2]; R.t = din[3]; R.dx = din[4]; R.dy = din[5]; R.dz = din[6]; R.lambda = din[7];
        R.ox = dout[0]; R.oy = dout[1]; R.oz = dout[2]; R.ot = dout[3]; R.odx = dout[4]; R.ody = dout[5]; R.odz = dout[6];
        R.status = diout[0]; R.last_node = diout[1]; R.npoints = diout[2];
        R.cur = nullptr; R.ndraw = nullptr;
        memset(&R.hist, 0, sizeof(R.hist));
        trace_device(s, o, R, m, o->ray_id_offset + (unsigned long long)b, s->streams[k]);
        CK(cudaEventRecord(s->ev_cmp[k], s->streams[k]));
        CK(cudaStreamWaitEvent(s->copy_out, s->ev_cmp[k], 0));
        if (out_pitch > 0) CK(cudaMemcpy2DAsync(hout[0] + b, (size_t)out_pitch, dout[0], (size_t)chunk * 8, (size_t)m * 8, 7, cudaMemcpyDeviceToHost, s->copy_out));
        else
          for (int a = 0; a < 7; a++) CK(cudaMemcpyAsync(hout[a] + b, dout[a], (size_t)m * 8, cudaMemcpyDeviceToHost, s->copy_out));
        if (iout_pitch > 0) CK(cudaMemcpy2DAsync(hiout[0] + b, (size_t)iout_pitch, diout[0], (size_t)chunk * 4, (size_t)m * 4, 3, cudaMemcpyDeviceToHost, s->copy_out));
        else
          for (int a = 0; a < 3; a++) CK(cudaMemcpyAsync(hiout[a] + b, diout[a], (size_t)m * 4, cudaMemcpyDeviceToHost, s->copy_out));
        CK(cudaEventRecord(s->ev_out[k], s->copy_out));
      }
      CK(cudaStreamSynchronize(s->copy_out));
      return;
    }
    // Pageable host arrays (and polyline records): a cudaMemcpyAsync from pageable memory blocks its caller while the driver
    // stages the data, so every chunk in flight gets its own host thread, stream and stage buffer.
    static const bool time_host = getenv("RB_TIME_HOST") != nullptr;
    const auto t_call = std::chrono::steady_clock::now();
    auto since = [&](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
    double t_enq[RB_HOST_STREAMS] = {}, t_wait[RB_HOST_STREAMS] = {};
    auto work = [&](int k) {
      const auto t_w = std::chrono::steady_clock::now();
      CK(cudaSetDevice(s->device));
      cudaStream_t st = s->streams[k];
      char* base = (char*)s->stage[k];
      for (long long ci = k; ci < nchunks; ci += nst) {
        long long b = chunks[ci].first, m = chunks[ci].second;
        double* din[8];
        double* dout[7];
        int32_t* diout[3];
        size_t off = 0;
        for (int a = 0; a < 8; a++) { din[a] = (double*)(base + off); off += (size_t)chunk * 8; }
        for (int a = 0; a < 7; a++) { dout[a] = (double*)(base + off); off += (size_t)chunk * 8; }
        for (int a = 0; a < 3; a++) { diout[a] = (int32_t*)(base + off); off += (size_t)chunk * 4; }
        off = (off + 255) & ~size_t(255);
        const double* hin[8] = {rays->x, rays->y, rays->z, rays->t, rays->dx, rays->dy, rays->dz, rays->lambda};
        double* hout[7] = {rays->ox, rays->oy, rays->oz, rays->ot, rays->odx, rays->ody, rays->odz};
        int32_t* hiout[3] = {rays->status, rays->last_node, rays->npoints};
        for (int a = 0; a < 8; a++) CK(cudaMemcpyAsync(din[a], hin[a] + b, (size_t)m * 8, cudaMemcpyHostToDevice, st));
        DRays R;
        R.x = din[0]; R.y = din[1]; R.z = din[2]; R.t = din[3]; R.dx = din[4]; R.dy = din[5]; R.dz = din[6]; R.lambda = din[7];
        R.ox = dout[0]; R.oy = dout[1]; R.oz = dout[2]; R.ot = dout[3]; R.odx = dout[4]; R.ody = dout[5]; R.odz = dout[6];
        R.status = diout[0]; R.last_node = diout[1]; R.npoints = diout[2];
        R.cur = nullptr; R.ndraw = nullptr;
        memset(&R.hist, 0, sizeof(R.hist));
        if (hpts > 0) {  // history stage: 4 double planes + 1 int plane of hpts x chunk
          char* hb = base + off;
          size_t plane = (size_t)chunk * hpts * 8;
          R.hist.x = (double*)hb; R.hist.y = (double*)(hb + plane); R.hist.z = (double*)(hb + 2 * plane); R.hist.t = (double*)(hb + 3 * plane);
          R.hist.node = (int32_t*)(hb + 4 * plane);
          R.hist.stride = chunk;
          R.hist.max_points = hpts;
        }
        trace_device(s, o, R, m, o->ray_id_offset + (unsigned long long)b, st);
        if (hpts > 0) {
          double* hd[4] = {hist->hx, hist->hy, hist->hz, hist->ht};
          double* dd[4] = {R.hist.x, R.hist.y, R.hist.z, R.hist.t};
          for (int a = 0; a < 4; a++)
            CK(cudaMemcpy2DAsync(hd[a] + b, (size_t)rays->n * 8, dd[a], (size_t)chunk * 8, (size_t)m * 8, hpts, cudaMemcpyDeviceToHost, st));
          CK(cudaMemcpy2DAsync(hist->hnode + b, (size_t)rays->n * 4, R.hist.node, (size_t)chunk * 4, (size_t)m * 4, hpts, cudaMemcpyDeviceToHost, st));
        }
        // the stream is in-order: the next chunk's H2D into this stage buffer waits for these copies
        for (int a = 0; a < 7; a++) CK(cudaMemcpyAsync(hout[a] + b, dout[a], (size_t)m * 8, cudaMemcpyDeviceToHost, st));
        for (int a = 0; a < 3; a++) CK(cudaMemcpyAsync(hiout[a] + b, diout[a], (size_t)m * 4, cudaMemcpyDeviceToHost, st));
      }
      t_enq[k] = since(t_w);
      CK(cudaStreamSynchronize(st));
      t_wait[k] = since(t_w);
    };
    if (nst == 1) {
      work(0);
      return;
    }
    std::exception_ptr errs[RB_HOST_STREAMS];
    std::vector<std::thread> workers;
    for (int k = 0; k < nst; k++)
      workers.emplace_back([&, k] {
        try {
          work(k);
        } catch (...) {
          errs[k] = std::current_exception();
        }
      });
    for (auto& w : workers) w.join();
    if (time_host) {
      fprintf(stderr, "rbg_trace host path: n=%lld chunks=%lld streams=%d total %.2f ms; per worker enqueue/done:", (long long)rays->n, nchunks, nst, since(t_call));
      for (int k = 0; k < nst; k++) fprintf(stderr, " %.2f/%.2f", t_enq[k], t_wait[k]);
      fprintf(stderr, "\n");
    }
    for (int k = 0; k < nst; k++) {
      if (errs[k]) {
        for (int j = 0; j < nst; j++) cudaStreamSynchronize(s->streams[j]);
        std::rethrow_exception(errs[k]);
      }
    }
  });
}

// ---- several GPUs, one process
int rbg_multi_create(const rbg_scene_desc* D, int ndev, const int* devices, rbg_multi** out) {
  rbg_multi* m = nullptr;
  int rc = guard([&] {
    if (!out) throw Invalid("null output handle");
    *out = nullptr;
    int have = rbg_device_count();
    if (have <= 0) throw std::runtime_error("cuda: no CUDA device available — the tracer has no CPU fallback");
    if (ndev < 1 || ndev > have) throw Invalid("bad device count");
    m = new rbg_multi;
    for (int k = 0; k < ndev; k++) {
      int d = devices ? devices[k] : k;
      if (d < 0 || d >= have) throw Invalid("bad device index");
      for (int j : m->devices)
        if (j == d) throw Invalid("device listed twice");
      m->devices.push_back(d);
      rbg_scene* s = nullptr;
      if (rbg_scene_create(D, d, &s) != RBG_OK) throw std::runtime_error(g_err);
      m->scenes.push_back(s);
    }
    // peer access towards the first device (the reducers' home): remote atomics over NVLink
    m->peer = ndev > 1;
    for (int k = 1; k < ndev && m->peer; k++) {
      int can = 0;
      CK(cudaDeviceCanAccessPeer(&can, m->devices[k], m->devices[0]));
      if (!can) { m->peer = false; break; }
      CK(cudaSetDevice(m->devices[k]));
      cudaError_t e = cudaDeviceEnablePeerAccess(m->devices[0], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) m->peer = false;
      cudaGetLastError();
    }
    if (getenv("RB_NO_PEER")) m->peer = false;
    *out = m;
    m = nullptr;
  });
  if (m) {
    for (rbg_scene* s : m->scenes) scene_free(s);
    delete m;
  }
  return rc;
}
int rbg_multi_destroy(rbg_multi* m) {
  return guard([&] {
    if (!m) return;
    for (rbg_scene* s : m->scenes) scene_free(s);
    delete m;
  });
}
int rbg_multi_num_devices(const rbg_multi* m) { return m ? (int)m->scenes.size() : 0; }

int rbg_multi_trace(rbg_multi* m, const rbg_trace_opts* o, const rbg_rays* rays) {
  return guard([&] {
    if (!m || !o || !rays) throw Invalid("null argument");
    if (rays->on_device) throw Invalid("rbg_multi_trace takes host arrays (a device batch lives on one GPU: use rbg_trace)");
    if (rays->n <= 0) return;
    const int G = (int)m->scenes.size();
    multi_run(m, [&](int k) {
      long long b, e;
      multi_range(rays->n, k, G, b, e);
      if (e <= b) return;
      rbg_rays r = *rays;
      r.n = e - b;
      r.x += b; r.y += b; r.z += b; r.t += b; r.dx += b; r.dy += b; r.dz += b; r.lambda += b;
      r.ox += b; r.oy += b; r.oz += b; r.ot += b; r.odx += b; r.ody += b; r.odz += b;
      r.status += b; r.last_node += b; r.npoints += b;
      rbg_trace_opts ok = *o;
      ok.ray_id_offset = o->ray_id_offset + (unsigned long long)b;
      if (rbg_trace(m->scenes[k], &ok, &r, nullptr) != RBG_OK) throw std::runtime_error(g_err);
    });
  });
}

int rbg_multi_shoot_trace_reduce(rbg_multi* m, const rbg_trace_opts* o, const rbg_shoot_desc* shoot, int64_t n_total, int64_t batch, int32_t sel, int32_t nx,
                                 double xmin, double xmax, int32_t ny, double ymin, double ymax, unsigned long long* hist, double* moments, long long* counts) {
  return guard([&] {
    if (!m || !o || !shoot || !hist || !moments || !counts) throw Invalid("null argument");
    if (nx < 1 || ny < 1 || !(xmax > xmin) || !(ymax > ymin)) throw Invalid("bad histogram axes");
    if (n_total < 0 || batch < 1) throw Invalid("bad ray counts");
    const int G = (int)m->scenes.size();
    const size_t nb = (size_t)nx * ny;
    // the histogram's home: device 0
    unsigned long long* d_hist0 = nullptr;
    CK(cudaSetDevice(m->devices[0]));
    CK(cudaMalloc((void**)&d_hist0, nb * 8 * (m->peer ? 1 : G)));
    CK(cudaMemset(d_hist0, 0, nb * 8 * (m->peer ? 1 : G)));
    std::vector<double> mom((size_t)G * 8, 0.);
    std::vector<long long> cnt((size_t)G * 6, 0);
    try {
      multi_run(m, [&](int k) {
        long long b, e;
        multi_range(n_total, k, G, b, e);
        const int dev = m->devices[k];
        CK(cudaSetDevice(dev));
        rbg_scene* s = m->scenes[k];
        if (!s->streams[0]) CK(cudaStreamCreateWithFlags(&s->streams[0], cudaStreamNonBlocking));
        cudaStream_t st = s->streams[0];
        const long long cap = std::min<long long>(batch, std::max<long long>(e - b, 1));
        char* buf = nullptr;
        // rays (8 in + 7 out doubles, 3 ints) + local reducers
        const size_t bytes = (size_t)cap * (15 * 8 + 3 * 4) + 1024 + nb * 8 + 256;
        CK(cudaMalloc((void**)&buf, bytes));
        try {
          double* in = (double*)buf;
          double* out = in + 8 * cap;
          int32_t* io = (int32_t*)(out + 7 * cap);
          char* tail = (char*)(((uintptr_t)(io + 3 * cap) + 255) & ~uintptr_t(255));
          double* d_mom = (double*)tail;
          unsigned long long* d_cnt = (unsigned long long*)(tail + 64);
          unsigned long long* d_hist = (unsigned long long*)(tail + 256);  // used without peer access only
          CK(cudaMemsetAsync(tail, 0, 256 + (m->peer ? 0 : nb * 8), st));
          unsigned long long* target = (m->peer || k == 0) ? d_hist0 : d_hist;
          const int use_smem = (long long)nx * ny <= HIST_SMEM_BINS;
          for (long long first = b; first < e; first += cap) {
            const long long n = std::min<long long>(cap, e - first);
            k_shoot<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(*shoot, first, n, in, in + cap, in + 2 * cap, in + 3 * cap, in + 4 * cap, in + 5 * cap,
                                                                 in + 6 * cap, in + 7 * cap);
            DRays R;
            R.x = in; R.y = in + cap; R.z = in + 2 * cap; R.t = in + 3 * cap; R.dx = in + 4 * cap; R.dy = in + 5 * cap; R.dz = in + 6 * cap; R.lambda = in + 7 * cap;
            R.ox = out; R.oy = out + cap; R.oz = out + 2 * cap; R.ot = out + 3 * cap; R.odx = out + 4 * cap; R.ody = out + 5 * cap; R.odz = out + 6 * cap;
            R.status = io; R.last_node = io + cap; R.npoints = io + 2 * cap;
            R.cur = nullptr; R.ndraw = nullptr;
            memset(&R.hist, 0, sizeof(R.hist));
            trace_device(s, o, R, n, o->ray_id_offset + (unsigned long long)first, st);
            const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 8);
            // with peer access `target` is device 0's histogram: the flush of the block-private bins is the cross-GPU reduction
            k_hist2d<<<blocks, 256, 0, st>>>(n, out, out + cap, io, sel, nx, xmin, xmax, ny, ymin, ymax, target, use_smem, nullptr, 0., 0.);
            k_moments<<<blocks, 256, 0, st>>>(n, out, out + cap, out + 3 * cap, io, sel, d_mom, d_cnt);
            g_launches += 3;
            CK(cudaGetLastError());
          }
          CK(cudaMemcpyAsync(&mom[(size_t)k * 8], d_mom, 7 * 8, cudaMemcpyDeviceToHost, st));
          CK(cudaMemcpyAsync(&cnt[(size_t)k * 6], d_cnt, 6 * 8, cudaMemcpyDeviceToHost, st));
          if (!m->peer && k > 0) CK(cudaMemcpyPeerAsync(d_hist0 + (size_t)k * nb, m->devices[0], d_hist, dev, nb * 8, st));
          CK(cudaStreamSynchronize(st));
        } catch (...) {
          cudaFree(buf);
          throw;
        }
        CK(cudaFree(buf));
      });
      CK(cudaSetDevice(m->devices[0]));
      if (!m->peer)
        for (int k = 1; k < G; k++) {
          k_add_u64<<<(unsigned)std::min<size_t>((nb + 255) / 256, 1024), 256>>>(d_hist0, d_hist0 + (size_t)k * nb, (long long)nb);
          g_launches++;
        }
      CK(cudaMemcpy(hist, d_hist0, nb * 8, cudaMemcpyDeviceToHost));
    } catch (...) {
      cudaFree(d_hist0);
      throw;
    }
    CK(cudaFree(d_hist0));
    for (int a = 0; a < 8; a++) moments[a] = 0;
    for (int a = 0; a < 6; a++) counts[a] = 0;
    for (int k = 0; k < G; k++) {
      for (int a = 0; a < 7; a++) moments[a] += mom[(size_t)k * 8 + a];
      for (int a = 0; a < 6; a++) counts[a] += cnt[(size_t)k * 6 + a];
    }
  });
}

int rbg_shoot(const rbg_shoot_desc* d, int64_t first, int64_t n, double* x, double* y, double* z, double* t, double* dx, double* dy, double* dz,
              double* lambda, int device, void* stream) {
  return guard([&] {
    if (!d || n < 0) throw Invalid("bad argument");
    if (d->kind < 0 || d->kind > 6) throw Invalid("unknown shooter kind");
    if ((d->kind == 0 || d->kind == 3) && (d->nx < 1 || d->ny < 1)) throw Invalid("grid shooters need nx,ny >= 1");
    if (n == 0) return;
    CK(cudaSetDevice(device));
    long long blocks = (n + 255) / 256;
    k_shoot<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*d, first, n, x, y, z, t, dx, dy, dz, lambda);
    g_launches++;
    CK(cudaGetLastError());
  });
}

// ---- ACorsikaIACTFile::GetRayArray on device: one thread per ray, the bunch found by bisection of the ray-count prefix sums
static long long bunch_count(float photons) {  // rays of `for (j = 0; j < photons; j++)`
  if (!(photons > 0)) return 0;
  long long c = (long long)ceilf(photons);
  return c;
}
__global__ void k_shoot_bunches(long long nb, const long long* __restrict__ offs, const float* __restrict__ bx, const float* __restrict__ by,
                                const float* __restrict__ bt, const float* __restrict__ bcx, const float* __restrict__ bcy, const float* __restrict__ bcz,
                                const float* __restrict__ blam, double z, double telz, double refidx, double lmin, double lmax, unsigned long long seed,
                                long long first, long long n, double* x, double* y, double* zz, double* t, double* dx, double* dy, double* dz, double* lambda) {
  long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  long long ray = first + j;
  long long lo = 0, hi = nb;  // last bunch with offs[b] <= ray
  while (hi - lo > 1) {
    long long mid = (lo + hi) >> 1;
    if (offs[mid] <= ray) lo = mid; else hi = mid;
  }
  const double cm = 1., ns = 1e-9, nm = 1e-7, m = 100.;
  double cx = bcx[lo], cy = bcy[lo], cz = bcz[lo];
  double airmass = -1. / cz, tel_dist = (z - telz * cm) * airmass, speed = 2.99792458e8 * m / refidx;
  double lam = blam[lo];
  if (lam == 0) {
    Philox g;
    g.k0 = (uint32_t)seed; g.k1 = (uint32_t)(seed >> 32);
    g.id0 = (uint32_t)ray; g.id1 = (uint32_t)((unsigned long long)ray >> 32);
    g.ndraw = 0x40000000u;  // the shooters' counter range (k_shoot): a trace of the same rays with the same seed starts at 0
    lam = 1. / (1. / lmin - rng_uniform(g) * (1. / lmin - 1. / lmax));
  }
  x[j] = bx[lo] * cm - tel_dist * cx;
  y[j] = by[lo] * cm - tel_dist * cy;
  zz[j] = z;
  t[j] = bt[lo] * ns - tel_dist / speed;
  dx[j] = cx; dy[j] = cy; dz[j] = cz;
  lambda[j] = lam * nm;
}
static void check_bunches(const rbg_bunches* b) {
  if (!b || b->nbunches < 0) throw Invalid("bad bunch table");
  if (b->nbunches > 0 && (!b->x || !b->y || !b->time || !b->cx || !b->cy || !b->cz || !b->lambda || !b->photons)) throw Invalid("null bunch array");
}
int rbg_bunch_rays(const rbg_bunches* b, int64_t* nrays) {
  return guard([&] {
    check_bunches(b);
    if (!nrays) throw Invalid("null argument");
    long long tot = 0;
    for (int64_t i = 0; i < b->nbunches; i++) tot += bunch_count(b->photons[i]);
    *nrays = tot;
  });
}
int rbg_shoot_bunches(const rbg_bunches* b, int64_t first, int64_t n, double* x, double* y, double* z, double* t, double* dx, double* dy, double* dz,
                      double* lambda, int device, void* stream) {
  return guard([&] {
    check_bunches(b);
    if (first < 0 || n < 0) throw Invalid("bad ray range");
    if (n == 0) return;
    if (rbg_device_count() <= 0) throw std::runtime_error("cuda: no CUDA device available — the shooters have no CPU fallback");
    std::vector<long long> offs((size_t)b->nbunches + 1, 0);
    for (int64_t i = 0; i < b->nbunches; i++) offs[i + 1] = offs[i] + bunch_count(b->photons[i]);
    if (first + n > offs[b->nbunches]) throw Invalid("ray range beyond the bunches' photons");
    // only the bunches that hold rays [first, first+n) travel to the device
    int64_t b0 = std::upper_bound(offs.begin(), offs.end(), (long long)first) - offs.begin() - 1;
    int64_t b1 = std::lower_bound(offs.begin(), offs.end(), (long long)(first + n)) - offs.begin();
    int64_t nb = b1 - b0;
    CK(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    char* d = nullptr;
    size_t fb = ((size_t)nb * 4 + 255) & ~size_t(255), ob = ((size_t)(nb + 1) * 8 + 255) & ~size_t(255);
    CK(cudaMallocAsync((void**)&d, ob + 7 * fb, st));
    try {
      CK(cudaMemcpyAsync(d, offs.data() + b0, (size_t)(nb + 1) * 8, cudaMemcpyHostToDevice, st));
      const float* src[7] = {b->x, b->y, b->time, b->cx, b->cy, b->cz, b->lambda};
      for (int a = 0; a < 7; a++) CK(cudaMemcpyAsync(d + ob + a * fb, src[a] + b0, (size_t)nb * 4, cudaMemcpyHostToDevice, st));
      auto f = [&](int a) { return (const float*)(d + ob + a * fb); };
      long long blocks = (n + 255) / 256;
      k_shoot_bunches<<<(unsigned)blocks, 256, 0, st>>>(nb, (const long long*)d, f(0), f(1), f(2), f(3), f(4), f(5), f(6), b->z, b->telescope_z,
                                                         b->refractive_index, b->lambda_min_nm, b->lambda_max_nm, b->seed, first, n, x, y, z, t, dx, dy, dz, lambda);
      g_launches++;
      CK(cudaGetLastError());
      CK(cudaStreamSynchronize(st));  // offs (a host temporary) and the caller's bunch arrays may go away after the return
    } catch (...) {
      cudaFreeAsync(d, st);
      throw;
    }
    cudaFreeAsync(d, st);
  });
}

int rbg_hist2d(int64_t n, const double* x, const double* y, const int32_t* status, int32_t sel, int32_t nx, double xmin, double xmax, int32_t ny,
               double ymin, double ymax, unsigned long long* hist, int device, void* stream) {
  return rbg_hist2d_stats(n, x, y, status, sel, 0., 0., nx, xmin, xmax, ny, ymin, ymax, hist, nullptr, device, stream);
}
int rbg_hist2d_stats(int64_t n, const double* x, const double* y, const int32_t* status, int32_t sel, double x0, double y0, int32_t nx, double xmin,
                     double xmax, int32_t ny, double ymin, double ymax, unsigned long long* hist, double* stats, int device, void* stream) {
  return guard([&] {
    if (nx < 1 || ny < 1 || !(xmax > xmin) || !(ymax > ymin)) throw Invalid("bad histogram axes");
    if (n <= 0) return;
    CK(cudaSetDevice(device));
    int use_smem = (long long)nx * ny <= HIST_SMEM_BINS;
    int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 8);
    k_hist2d<<<blocks, 256, 0, (cudaStream_t)stream>>>(n, x, y, status, sel, nx, xmin, xmax, ny, ymin, ymax, hist, use_smem, stats, x0, y0);
    g_launches++;
    CK(cudaGetLastError());
  });
}

int rbg_moments(int64_t n, const double* x, const double* y, const double* t, const int32_t* status, int32_t sel, double* moments, long long* counts,
                int device, void* stream) {
  return guard([&] {
    if (n <= 0) return;
    CK(cudaSetDevice(device));
    int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 8);
    k_moments<<<blocks, 256, 0, (cudaStream_t)stream>>>(n, x, y, t, status, sel, moments, (unsigned long long*)counts);
    g_launches++;
    CK(cudaGetLastError());
  });
}

int rbg_containment_radius(int32_t nhist, const unsigned long long* hist, int32_t nx, double xmin, double xmax, int32_t ny, double ymin, double ymax,
                           const double* stats, double fraction, double* out, int device, void* stream) {
  return guard([&] {
    if (nhist < 0 || nx < 1 || ny < 1 || !(xmax > xmin) || !(ymax > ymin) || !hist || !stats || !out) throw Invalid("bad containment-radius arguments");
    if (nhist == 0) return;
    CK(cudaSetDevice(device));
    double* prefix = nullptr;  // stream-ordered scratch for the per-row running sums
    CK(cudaMallocFromPoolAsync((void**)&prefix, (size_t)nhist * (nx + 1) * ny * sizeof(double), small_pool(device), (cudaStream_t)stream));
    int rc = rb_launch_containment_u64(nhist, hist, nx, xmin, xmax, ny, ymin, ymax, stats, fraction, out, prefix, (cudaStream_t)stream);
    cudaFreeAsync(prefix, (cudaStream_t)stream);
    CK((cudaError_t)rc);
    g_launches++;
  });
}
int rbg_containment_radius_host(const double* bins, int32_t nx, double xmin, double xmax, int32_t ny, double ymin, double ymax, const double* stats,
                                double fraction, double* out, int device) {
  return guard([&] {
    if (nx < 1 || ny < 1 || !(xmax > xmin) || !(ymax > ymin) || !bins || !stats || !out) throw Invalid("bad containment-radius arguments");
    if (rbg_device_count() <= 0) throw std::runtime_error("cuda: no CUDA device available — the reducers have no CPU fallback");
    CK(cudaSetDevice(device));
    double* d = nullptr;
    size_t nb = (size_t)nx * ny;
    CK(cudaMalloc((void**)&d, (nb + 8 + (size_t)(nx + 1) * ny) * 8));
    try {
      CK(cudaMemcpy(d, bins, nb * 8, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(d + nb, stats, 5 * 8, cudaMemcpyHostToDevice));
      CK((cudaError_t)rb_launch_containment_f64(1, d, nx, xmin, xmax, ny, ymin, ymax, d + nb, fraction, d + nb + 5, d + nb + 8, nullptr));
      g_launches++;
      CK(cudaMemcpy(out, d + nb + 5, 3 * 8, cudaMemcpyDeviceToHost));
    } catch (...) {
      cudaFree(d);
      throw;
    }
    cudaFree(d);
  });
}

int rbg_tmm(rbg_scene* s, int ml, int64_t n, const double* theta, const double* lambda, double* refl, double* trans, void* stream) {
  return guard([&] {
    if (!s) throw Invalid("null scene");
    if (n <= 0) return;
    CK(cudaSetDevice(s->device));
    long long blocks = (n + 127) / 128;
    k_tmm<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(s->d, ml, n, theta, lambda, refl, trans);
    g_launches++;
    CK(cudaGetLastError());
  });
}

int rbg_tmm_general_host(rbg_scene* s, int ml, int mode, int pol, int reverse, int64_t n, const double* theta_re, const double* theta_im, const double* lambda,
                         double* refl, double* trans) {
  return guard([&] {
    if (!s) throw Invalid("null scene");
    if (ml < 0 || mode < 0 || mode > 1 || pol < 0 || pol > 2 || !theta_re || !theta_im || !lambda || !refl || !trans) throw Invalid("bad TMM arguments");
    if (n <= 0) return;
    CK(cudaSetDevice(s->device));
    double* d = nullptr;
    CK(cudaMalloc((void**)&d, (size_t)n * 8 * 5));
    try {
      CK(cudaMemcpy(d, theta_re, (size_t)n * 8, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(d + n, theta_im, (size_t)n * 8, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(d + 2 * n, lambda, (size_t)n * 8, cudaMemcpyHostToDevice));
      long long blocks = (n + 127) / 128;
      k_tmm_general<<<(unsigned)blocks, 128>>>(s->d, ml, mode, pol, reverse, n, d, d + n, d + 2 * n, d + 3 * n, d + 4 * n);
      g_launches++;
      CK(cudaGetLastError());
      CK(cudaMemcpy(refl, d + 3 * n, (size_t)n * 8, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(trans, d + 4 * n, (size_t)n * 8, cudaMemcpyDeviceToHost));
    } catch (...) {
      cudaFree(d);
      throw;
    }
    cudaFree(d);
  });
}

int rbg_tmm_host(rbg_scene* s, int ml, int64_t n, const double* theta, const double* lambda, double* refl, double* trans) {
  return guard([&] {
    if (!s) throw Invalid("null scene");
    if (n <= 0) return;
    CK(cudaSetDevice(s->device));
    double* d = nullptr;
    CK(cudaMalloc((void**)&d, (size_t)n * 8 * 4));
    try {
      CK(cudaMemcpy(d, theta, (size_t)n * 8, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(d + n, lambda, (size_t)n * 8, cudaMemcpyHostToDevice));
      long long blocks = (n + 127) / 128;
      k_tmm<<<(unsigned)blocks, 128>>>(s->d, ml, n, d, d + n, d + 2 * n, d + 3 * n);
      g_launches++;
      CK(cudaGetLastError());
      CK(cudaMemcpy(refl, d + 2 * n, (size_t)n * 8, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(trans, d + 3 * n, (size_t)n * 8, cudaMemcpyDeviceToHost));
    } catch (...) {
      cudaFree(d);
      throw;
    }
    cudaFree(d);
  });
}

}  // extern "C"
#pragma GCC visibility pop
