// Local-map post-processing for sm_100a: one thread block per voxel grid, SURVEY.md section 8(f) row 4.
//
// Replaces, per agent and per map update, the sequence of mapping_util/src/map_builder.cpp:207-216:
//   MapBuilder::SetUncertainToUnknown   map_builder.cpp:331-365            unknown space grows by the inflation cube
//   VoxelGrid::InflateObstacles         voxel_grid_util/src/voxel_grid.cpp:251-277
//   VoxelGrid::CreatePotentialField     voxel_grid.cpp:279-298
// (stencils from VoxelGrid::CreateMask, voxel_grid.cpp:192-226).  The output grid is what
// hdsm_corridor_batch and hdsm_reftraj_batch read.
//
// Design.  A local grid (66 x 66 x 20 int8 = 87 KB) fits twice into the 227 KB of shared memory of one SM: the
// block loads it once (one bulk asynchronous copy), runs the three passes ping-ponging between the two copies, and
// writes the result once - HBM sees exactly one read and one write per voxel.  The reference scatters from every
// occupied voxel; here every voxel gathers, which needs no atomics and is exact because none of the three passes
// feeds on its own output (the potential stencil's only value of 100 is its centre).  All passes work on voxel sets
// packed one bit per voxel along x.  The potential field is computed as an exact separable squared-distance transform
// (the stencil is a function of the distance: verified on the host), unknown-space growth and the default 3x3x3
// inflation as separable dilations of the bit rows; stencils of another shape keep the row / entry forms below.
// The stencils are computed on the host when the handle is created, with the reference's formula.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/hdsm.h"
#include "hdsm_common.h"

namespace hdsm_mp {

constexpr int kOcc = 100, kUnk = -1;
constexpr int kThreads = 512;

struct Args {
  int n_grids, cube, n_inf, n_pot;
  size_t stride;
  const int8_t* in;
  int8_t* out;
  const int32_t* dims;
  const int8_t* inf_off;  // [n_inf][4] (dx, dy, dz, -)
  const int8_t* pot_off;  // [n_pot][4] (dx, dy, dz, value), value descending
  // row form of the potential stencil: for every (dy, dz) the values at |dx| = 0..7, pairs sorted by their best value
  int n_pair, rn;
  int n_irow, rn_inf;                   // row form of the inflation stencil: (dy | dz << 8 | reach << 16) per row
  const int* irow;
  size_t bits_bytes;                    // shared memory set aside for the occupancy bit rows
  const int2* pair_yz;                  // (dy, dz) and best value packed: x = dy | dz << 8 | best << 16, y unused
  const unsigned long long* pair_vals;  // eight int8 values, |dx| = 0..7
  const int8_t* dist_tab;               // != null: the potential stencil is a function of the squared distance alone, [128] values (-128: none)
  int inf_cube;                         // 1: the inflation stencil is the full cube of half-width rn_inf
};

// ---- bulk asynchronous copies (TMA engine, sm_90+): one thread moves a whole grid between HBM and shared memory
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* mbar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(mbar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_addr(mbar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* mbar, unsigned parity) {
  unsigned done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(smem_addr(mbar)), "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void* gmem_dst, const void* smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_addr(smem_src)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// One axis of the exact squared-distance transform, in place: every entry of a column becomes the minimum over the
// entries within +-RN of (old entry + offset^2), capped at 127.  A thread owns whole columns and carries the 2 RN + 1
// old entries around its position in registers, so overwriting the column as it goes is safe.  The last axis (FINAL)
// writes the voxel's new value instead of the distance: the table value at that distance where it beats the value the
// voxel has in `cur` (unknown voxels keep theirs).
template <int RN, bool FINAL>
__device__ __forceinline__ void column_pass(uint8_t* q, int ncol, int inner, int outer_stride, int len, int stride, int tid,
                                            const int8_t* cur, const int8_t* tab) {
  constexpr int kFar = 1000;
  for (int c = tid; c < ncol; c += kThreads) {
    const size_t base = (c % inner) + (size_t)(c / inner) * outer_stride;
    uint8_t* col = q + base;
    int w[2 * RN + 1];
#pragma unroll
    for (int j = 0; j <= 2 * RN; ++j) w[j] = (j >= RN && j - RN < len) ? (int)col[(j - RN) * stride] : kFar;
    for (int pos = 0; pos < len; ++pos) {
      int best = kFar;
#pragma unroll
      for (int j = 0; j <= 2 * RN; ++j) best = min(best, w[j] + (j - RN) * (j - RN));
      best = min(best, 127);
      if (FINAL) {
        const int v = cur[base + pos * stride], val = tab[best];
        col[pos * stride] = (uint8_t)(int8_t)((v != kUnk && val > v) ? val : v);
      } else {
        col[pos * stride] = (uint8_t)best;
      }
#pragma unroll
      for (int j = 0; j < 2 * RN; ++j) w[j] = w[j + 1];
      w[2 * RN] = pos + RN + 1 < len ? (int)col[(pos + RN + 1) * stride] : kFar;
    }
  }
}
template <int RN>
__device__ __forceinline__ void distance_passes(uint8_t* q, int dx, int dy, int dz, int tid, const int8_t* cur, const int8_t* tab) {
  column_pass<RN, false>(q, dx * dz, dx, dx * dy, dy, dx, tid, cur, tab);  // along y: columns (x, z)
  __syncthreads();
  column_pass<RN, true>(q, dx * dy, dx * dy, 0, dz, dx * dy, tid, cur, tab);  // along z: columns (x, y), and the new values
}

__global__ void __launch_bounds__(kThreads) map_kernel(const Args A) {
  extern __shared__ __align__(128) int8_t smem[];
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ int8_t s_tab[128];
  const int g = blockIdx.x;
  if (g >= A.n_grids) return;
  const int dx = A.dims[3 * g], dy = A.dims[3 * g + 1], dz = A.dims[3 * g + 2];
  const int nvox = dx * dy * dz;
  const size_t half = (A.stride + 15) & ~size_t(15);
  int8_t* P = smem;         // ping
  int8_t* Q = smem + half;  // pong
  int* st_inf = reinterpret_cast<int*>(smem + 2 * half);  // stencils: one packed int (dx, dy, dz, value) per entry
  int* st_pot = st_inf + A.n_inf;
  const int8_t* src = A.in + (size_t)g * A.stride;
  int8_t* dst = A.out + (size_t)g * A.stride;
  const int tid = threadIdx.x;
  for (int m = tid; m < A.n_inf; m += kThreads) st_inf[m] = reinterpret_cast<const int*>(A.inf_off)[m];
  for (int m = tid; m < A.n_pot; m += kThreads) st_pot[m] = reinterpret_cast<const int*>(A.pot_off)[m];
  if (A.dist_tab && tid < 128) s_tab[tid] = A.dist_tab[tid];
  // ---- load: one bulk asynchronous copy of the whole grid (cp.async.bulk, completion on an mbarrier) where the grid
  // start allows it, issued by one thread while the others fetch the stencils; the last < 16 bytes by plain loads
  const bool bulk = (reinterpret_cast<size_t>(src) & 15) == 0 && (reinterpret_cast<size_t>(dst) & 15) == 0;
  const unsigned nbulk = bulk ? (unsigned)(nvox / 16) * 16u : 0u;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0 && nbulk) bulk_load(P, src, nbulk, &mbar);
  for (int i = (int)nbulk + tid; i < nvox; i += kThreads) P[i] = src[i];
  if (nbulk) mbar_wait(&mbar, 0);
  __syncthreads();
  // Occupancy / unknown flags are packed one bit per voxel along x (bit x + 8 of a row, so that a window never
  // starts below bit 0): "is there a set bit within +-r of x in row (y', z')" is then one 64-bit funnel shift and a
  // mask instead of 2r + 1 byte loads.  All three passes use it when the radii are at most 7 voxels and the rows fit.
  const int nw = (dx + 16 + 31) / 32 + 1;  // words per row, one spare
  int* pr_yz = st_inf + ((A.n_inf + A.n_pot + 1) & ~1);  // row tables in shared memory (8-byte aligned)
  unsigned long long* pr_val = reinterpret_cast<unsigned long long*>(pr_yz + ((A.n_pair + 1) & ~1));
  int* ir = reinterpret_cast<int*>(pr_val + A.n_pair);
  unsigned* bits = reinterpret_cast<unsigned*>(ir + ((A.n_irow + 3) & ~3));
  for (int m = tid; m < A.n_pair; m += kThreads) pr_yz[m] = A.pair_yz[m].x, pr_val[m] = A.pair_vals[m];
  for (int m = tid; m < A.n_irow; m += kThreads) ir[m] = A.irow[m];
  const bool fits = (size_t)nw * dy * dz * 4 <= A.bits_bytes;
  const int warp = tid >> 5, lane = tid & 31, nwarp = kThreads / 32;
  const auto window = [&](int yy, int zz, int x, int r) -> unsigned {  // the bits of row (yy, zz) at x - r .. x + r
    const int bp = x - r + 8;
    const unsigned* row = bits + (yy + zz * dy) * nw + (bp >> 5);
    return (unsigned)((((unsigned long long)row[1] << 32) | row[0]) >> (bp & 31)) & ((1u << (2 * r + 1)) - 1u);
  };
  // Dilation of a bit-packed voxel set by a cube of half-width r, one axis at a time (x inside the rows, y and z as
  // ORs of whole words of neighbouring rows): a few word operations per thread where the gather form walks (2 r + 1)^2
  // windows per voxel.  Needs a second set of bit rows; the result lands in bitsB.
  const int nbw = nw * dy * dz;
  const bool fits2 = (size_t)2 * nbw * 4 <= A.bits_bytes;
  unsigned* bitsB = bits + nbw;
  const auto dilate_cube = [&](int r) {
    for (int idx = tid; idx < nbw; idx += kThreads) {  // x: bits -> bitsB
      const int w = idx % nw;
      const unsigned cur = bits[idx], lo = w > 0 ? bits[idx - 1] : 0u, hi = w < nw - 1 ? bits[idx + 1] : 0u;
      unsigned o = cur;
      for (int sft = 1; sft <= r; ++sft) o |= (cur << sft) | (lo >> (32 - sft)) | (cur >> sft) | (hi << (32 - sft));
      bitsB[idx] = o;
    }
    __syncthreads();
    for (int idx = tid; idx < nbw; idx += kThreads) {  // y: bitsB -> bits
      const int w = idx % nw, row = idx / nw, y = row % dy, z = row / dy;
      unsigned o = 0;
      for (int yy = max(y - r, 0); yy <= min(y + r, dy - 1); ++yy) o |= bitsB[(yy + z * dy) * nw + w];
      bits[idx] = o;
    }
    __syncthreads();
    for (int idx = tid; idx < nbw; idx += kThreads) {  // z: bits -> bitsB
      const int w = idx % nw, row = idx / nw, y = row % dy, z = row / dy;
      unsigned o = 0;
      for (int zz = max(z - r, 0); zz <= min(z + r, dz - 1); ++zz) o |= bits[(y + zz * dy) * nw + w];
      bitsB[idx] = o;
    }
    __syncthreads();
  };
  const auto bit_of = [&](const unsigned* b, int row, int x) -> bool { return (b[row * nw + ((x + 8) >> 5)] >> ((x + 8) & 31)) & 1u; };
  // ---- SetUncertainToUnknown: P -> Q.  A voxel that is not occupied turns unknown when an unknown voxel of the
  // interior [c, dim - c) lies within the cube of half-width c around it.
  const int c = A.cube;
  if (fits2 && c >= 1 && c <= 7) {
    for (int r = warp; r < dy * dz; r += nwarp) {
      const int y = r % dy, z = r / dy;
      const bool row_in = y >= c && y < dy - c && z >= c && z < dz - c;
      for (int w = 0; w < nw; ++w) {
        const int x = w * 32 + lane - 8;
        const unsigned m = __ballot_sync(0xffffffffu, row_in && x >= c && x < dx - c && P[x + r * dx] == kUnk);
        if (lane == 0) bits[r * nw + w] = m;
      }
    }
    __syncthreads();
    dilate_cube(c);
    for (int i = tid; i < nvox; i += kThreads) {
      const int r = i / dx, x = i - r * dx;
      const int8_t v = P[i];
      Q[i] = (v != kOcc && bit_of(bitsB, r, x)) ? (int8_t)kUnk : v;
    }
  } else if (fits && c >= 1 && c <= 7) {
    for (int r = warp; r < dy * dz; r += nwarp) {
      const int y = r % dy, z = r / dy;
      const bool row_in = y >= c && y < dy - c && z >= c && z < dz - c;
      for (int w = 0; w < nw; ++w) {
        const int x = w * 32 + lane - 8;
        const unsigned m = __ballot_sync(0xffffffffu, row_in && x >= c && x < dx - c && P[x + r * dx] == kUnk);
        if (lane == 0) bits[r * nw + w] = m;
      }
    }
    __syncthreads();
    for (int i = tid; i < nvox; i += kThreads) {
      const int z = i / (dx * dy), r = i - z * dx * dy, y = r / dx, x = r - y * dx;
      int8_t v = P[i];
      if (v != kOcc && v != kUnk) {
        bool hit = false;
        for (int zz = max(z - c, 0); zz <= min(z + c, dz - 1) && !hit; ++zz)
          for (int yy = max(y - c, 0); yy <= min(y + c, dy - 1); ++yy)
            if (window(yy, zz, x, c)) {
              hit = true;
              break;
            }
        if (hit) v = kUnk;
      }
      Q[i] = v;
    }
  } else
  for (int i = tid; i < nvox; i += kThreads) {
    const int z = i / (dx * dy), r = i - z * dx * dy, y = r / dx, x = r - y * dx;
    int8_t v = P[i];
    if (v != kOcc && v != kUnk && c > 0) {
      bool hit = false;
      for (int zz = max(z - c, c); zz <= min(z + c, dz - c - 1) && !hit; ++zz)
        for (int yy = max(y - c, c); yy <= min(y + c, dy - c - 1) && !hit; ++yy)
          for (int xx = max(x - c, c); xx <= min(x + c, dx - c - 1); ++xx)
            if (P[xx + yy * dx + zz * dx * dy] == kUnk) {
              hit = true;
              break;
            }
      if (hit) v = kUnk;
    }
    Q[i] = v;
  }
  __syncthreads();
  // ---- InflateObstacles: Q -> P.  A voxel becomes occupied when an occupied voxel (before the pass) has it in
  // its stencil: gather over the mirrored stencil - row form (per stencil row the largest |dx| it reaches), or
  // entry by entry.
  const bool occ_bits_ready = fits2 && A.inf_cube && A.rn_inf >= 1 && A.rn_inf <= 7;  // bitsB will hold the occupancy after inflation
  if (occ_bits_ready) {
    for (int r = warp; r < dy * dz; r += nwarp)
      for (int w = 0; w < nw; ++w) {
        const int x = w * 32 + lane - 8;
        const unsigned m = __ballot_sync(0xffffffffu, x >= 0 && x < dx && Q[x + r * dx] == kOcc);
        if (lane == 0) bits[r * nw + w] = m;
      }
    __syncthreads();
    dilate_cube(A.rn_inf);
    for (int i = tid; i < nvox; i += kThreads) {
      const int r = i / dx, x = i - r * dx;
      P[i] = bit_of(bitsB, r, x) ? (int8_t)kOcc : Q[i];
    }
  } else if (fits && A.n_irow > 0 && A.rn_inf <= 7) {
    for (int r = warp; r < dy * dz; r += nwarp)
      for (int w = 0; w < nw; ++w) {
        const int x = w * 32 + lane - 8;
        const unsigned m = __ballot_sync(0xffffffffu, x >= 0 && x < dx && Q[x + r * dx] == kOcc);
        if (lane == 0) bits[r * nw + w] = m;
      }
    __syncthreads();
    for (int i = tid; i < nvox; i += kThreads) {
      const int z = i / (dx * dy), r = i - z * dx * dy, y = r / dx, x = r - y * dx;
      int8_t v = Q[i];
      if (v != kOcc) {
        for (int m = 0; m < A.n_irow; ++m) {
          const int e = ir[m];
          const int yy = y + (int)(signed char)(e & 0xff), zz = z + (int)(signed char)((e >> 8) & 0xff);
          if ((unsigned)yy < (unsigned)dy && (unsigned)zz < (unsigned)dz && window(yy, zz, x, e >> 16)) {
            v = kOcc;
            break;
          }
        }
      }
      P[i] = v;
    }
  } else
  for (int i = tid; i < nvox; i += kThreads) {
    const int z = i / (dx * dy), r = i - z * dx * dy, y = r / dx, x = r - y * dx;
    int8_t v = Q[i];
    if (v != kOcc) {
      for (int m = 0; m < A.n_inf; ++m) {
        const int e = st_inf[m];
        const int xx = x - (int)(signed char)(e & 0xff), yy = y - (int)(signed char)((e >> 8) & 0xff), zz = z - (int)(signed char)((e >> 16) & 0xff);
        if ((unsigned)xx < (unsigned)dx && (unsigned)yy < (unsigned)dy && (unsigned)zz < (unsigned)dz && Q[xx + yy * dx + zz * dx * dy] == kOcc) {
          v = kOcc;
          break;
        }
      }
    }
    P[i] = v;
  }
  __syncthreads();
  // ---- CreatePotentialField: P -> Q.  Known voxels take the largest stencil value any occupied voxel offers.
  // Row form: occupancy is packed one bit per voxel along x (bit x + 8 of the row, so that a window never starts
  // below bit 0); for a stencil row (dy, dz) the nearest occupied voxel along x is a count-leading / find-first on
  // an 11-bit window, and because the value only falls with |dx| that one voxel decides the row.  Rows are
  // visited in order of their best value and the walk stops once no row can beat what the voxel already has.
  const bool rows_ok = A.n_pair > 0 && A.rn <= 7 && fits;
  if (rows_ok && A.dist_tab) {
    // Distance form.  Every stencil the reference's CreateMask produces is a non-increasing function of the Euclidean
    // distance, cut off below (rn + 1) voxels (checked entry by entry when the handle is created): the largest value
    // any occupied voxel offers is the table value at the squared distance to the NEAREST one.  That distance
    // transform is separable and exact in integers: nearest occupied voxel along x from the bit rows (one window per
    // voxel), then min over dy of (.. + dy^2) and min over dz of (.. + dz^2) down the columns - 2 rn + 2 row steps per
    // voxel and axis instead of a walk over up to (2 rn + 1)^2 stencil rows.
    if (occ_bits_ready) {  // the dilation that inflated the obstacles left their bit rows behind: cut to the grid's x range
      for (int idx = tid; idx < nbw; idx += kThreads) {
        const int w = idx % nw, lo = max(8, 32 * w), hi = min(dx + 8, 32 * w + 32);
        const unsigned m = hi <= lo ? 0u : (hi - lo == 32 ? 0xffffffffu : ((1u << (hi - lo)) - 1u) << (lo - 32 * w));
        bits[idx] = bitsB[idx] & m;
      }
    } else {
      for (int r = warp; r < dy * dz; r += nwarp)
        for (int w = 0; w < nw; ++w) {
          const int x = w * 32 + lane - 8;
          const unsigned m = __ballot_sync(0xffffffffu, x >= 0 && x < dx && P[x + r * dx] == kOcc);
          if (lane == 0) bits[r * nw + w] = m;
        }
    }
    __syncthreads();
    const int rn = A.rn;
    const unsigned wmask = (1u << (2 * rn + 1)) - 1u, lmask = (1u << (rn + 1)) - 1u;
    uint8_t* D = reinterpret_cast<uint8_t*>(Q);
    for (int i = tid; i < nvox; i += kThreads) {
      const int r = i / dx, x = i - r * dx;
      const int bp = x - rn + 8;
      const unsigned* row = bits + r * nw + (bp >> 5);
      const unsigned win = (unsigned)((((unsigned long long)row[1] << 32) | row[0]) >> (bp & 31)) & wmask;
      int d = 99;
      const unsigned right = win >> rn, left = win & lmask;
      if (right) d = __ffs(right) - 1;
      if (left) d = min(d, rn - (31 - __clz(left)));
      D[i] = (uint8_t)(d <= rn ? d * d : 127);
    }
    __syncthreads();
    switch (rn) {
      case 1: distance_passes<1>(D, dx, dy, dz, tid, P, s_tab); break;
      case 2: distance_passes<2>(D, dx, dy, dz, tid, P, s_tab); break;
      case 3: distance_passes<3>(D, dx, dy, dz, tid, P, s_tab); break;
      case 4: distance_passes<4>(D, dx, dy, dz, tid, P, s_tab); break;
      case 5: distance_passes<5>(D, dx, dy, dz, tid, P, s_tab); break;
      case 6: distance_passes<6>(D, dx, dy, dz, tid, P, s_tab); break;
      default: distance_passes<7>(D, dx, dy, dz, tid, P, s_tab); break;
    }
  } else if (rows_ok) {
    for (int r = warp; r < dy * dz; r += nwarp)
      for (int w = 0; w < nw; ++w) {
        const int x = w * 32 + lane - 8;
        const unsigned m = __ballot_sync(0xffffffffu, x >= 0 && x < dx && P[x + r * dx] == kOcc);
        if (lane == 0) bits[r * nw + w] = m;
      }
    __syncthreads();
    const int rn = A.rn;
    const unsigned wmask = (1u << (2 * rn + 1)) - 1u, lmask = (1u << (rn + 1)) - 1u;
    for (int i = tid; i < nvox; i += kThreads) {
      const int z = i / (dx * dy), r = i - z * dx * dy, y = r / dx, x = r - y * dx;
      int v = P[i];
      if (v != kUnk && v != kOcc) {
        const int bp = x - rn + 8, wi = bp >> 5, sh = bp & 31;
        for (int m = 0; m < A.n_pair; ++m) {
          const int e = pr_yz[m];
          if (((e >> 16) & 0xff) <= v) break;
          const int yy = y + (int)(signed char)(e & 0xff), zz = z + (int)(signed char)((e >> 8) & 0xff);
          if ((unsigned)yy >= (unsigned)dy || (unsigned)zz >= (unsigned)dz) continue;
          const unsigned* row = bits + (yy + zz * dy) * nw + wi;
          const unsigned win = (unsigned)((((unsigned long long)row[1] << 32) | row[0]) >> sh) & wmask;
          if (!win) continue;
          const unsigned right = win >> rn, left = win & lmask;
          int d = 99;
          if (right) d = __ffs(right) - 1;
          if (left) d = min(d, rn - (31 - __clz(left)));
          const int val = (int)(signed char)((pr_val[m] >> (8 * d)) & 0xff);
          v = max(v, val);
        }
      }
      Q[i] = (int8_t)v;
    }
  } else {
    // scan form: the stencil is sorted by value, so the first occupied hit decides and values not above the own
    // one end the walk
    for (int i = tid; i < nvox; i += kThreads) {
      const int z = i / (dx * dy), r = i - z * dx * dy, y = r / dx, x = r - y * dx;
      int8_t v = P[i];
      if (v != kUnk && v != kOcc) {
        for (int m = 0; m < A.n_pot; ++m) {
          const int e = st_pot[m];
          const int8_t val = (int8_t)(e >> 24);
          if (val <= v) break;
          const int xx = x - (int)(signed char)(e & 0xff), yy = y - (int)(signed char)((e >> 8) & 0xff), zz = z - (int)(signed char)((e >> 16) & 0xff);
          if ((unsigned)xx < (unsigned)dx && (unsigned)yy < (unsigned)dy && (unsigned)zz < (unsigned)dz && P[xx + yy * dx + zz * dx * dy] == kOcc) {
            v = val;
            break;
          }
        }
      }
      Q[i] = v;
    }
  }
  __syncthreads();
  // ---- store: the finished grid leaves shared memory as one bulk asynchronous copy (the writes above were made
  // through the generic proxy: fenced for the async proxy before the barrier)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0 && nbulk) bulk_store(dst, Q, nbulk);
  for (int i = (int)nbulk + tid; i < nvox; i += kThreads) dst[i] = Q[i];
}

// VoxelGrid::CreateMask (voxel_grid.cpp:192-226) on the host
struct MaskEntry {
  int x, y, z;
  int8_t v;
};
inline std::vector<MaskEntry> create_mask(double vox, double mask_dist, double power) {
  std::vector<MaskEntry> m;
  if (!(mask_dist > 0)) return m;
  const int rn = (int)std::ceil(mask_dist / vox);
  for (int x = -rn; x <= rn; ++x)
    for (int y = -rn; y <= rn; ++y)
      for (int z = -rn; z <= rn; ++z) {
        const double d = std::hypot(std::hypot((double)x, (double)y), (double)z);
        if (std::abs(d - 1) * vox >= mask_dist) continue;
        const double h = 100.0 * std::pow((1 - (double)std::hypot(std::hypot((double)x, (double)y), (double)z) / (rn + 1)), power);
        if (h > 1e-3) m.push_back(MaskEntry{x, y, z, (int8_t)h});
      }
  return m;
}

// Distance form of a stencil: tab[d2] = the value of every entry at squared distance d2 (-128: none).  Usable (returns
// true) when the mask is a non-increasing function of x^2 + y^2 + z^2 that is complete inside the cube of half-width rn:
// equal values at equal squared distances, every offset of the cube with a listed squared distance in the mask, and no
// entry after the first squared distance of the cube that is missing - true for every mask CreateMask builds with
// rn <= 7, verified here for the handle's own parameters all the same.  Pure host code.
inline bool distance_table(const std::vector<MaskEntry>& mask, int rn, int8_t* tab) {
  for (int i = 0; i < 128; ++i) tab[i] = (int8_t)-128;
  if (rn < 1 || rn > 7 || mask.empty()) return false;
  for (const auto& e : mask) {
    const int d2 = e.x * e.x + e.y * e.y + e.z * e.z;
    if (std::abs(e.x) > rn || std::abs(e.y) > rn || std::abs(e.z) > rn) return false;
    if (d2 >= 127 || e.v == -128 || (tab[d2] != -128 && tab[d2] != e.v)) return false;  // 127 is the kernel's "farther than anything"
    tab[d2] = e.v;
  }
  std::vector<char> key(128, 0);
  for (int x = -rn; x <= rn; ++x)
    for (int y = -rn; y <= rn; ++y)
      for (int z = -rn; z <= rn; ++z) {
        const int d2 = x * x + y * y + z * z;
        if (d2 >= 128) continue;  // cannot be in the mask (checked above)
        key[d2] = 1;
        bool in = false;
        for (const auto& e : mask) in |= e.x == x && e.y == y && e.z == z;
        if (in != (tab[d2] != -128)) return false;
      }
  int last = 127;
  bool gone = false;
  for (int d2 = 0; d2 < 128; ++d2) {
    if (!key[d2]) continue;
    if (tab[d2] != -128) {
      if (gone || tab[d2] > last) return false;
      last = tab[d2];
    } else {
      gone = true;
    }
  }
  return true;
}

}  // namespace hdsm_mp

struct hdsm_map {
  hdsm_map_params prm{};
  int device = 0, max_grids = 0, cube = 0, n_inf = 0, n_pot = 0;
  size_t grid_stride = 0, smem = 0;
  cudaStream_t stream = nullptr;
  int8_t *d_inf = nullptr, *d_pot = nullptr;
  int2* d_pair = nullptr;
  int* d_irow = nullptr;
  int n_irow = 0, rn_inf = 0;
  unsigned long long* d_pvals = nullptr;
  int8_t* d_tab = nullptr;  // distance form of the potential stencil (null: the stencil is not a function of the distance)
  int inf_cube = 0;         // the inflation stencil is the full cube of half-width rn_inf
  int n_pair = 0, rn = 0;
  size_t bits_bytes = 0;
  unsigned char *d_buf = nullptr;
  size_t buf_cap = 0;
  int64_t launches = 0;
  std::string err;
};

namespace {
int mfail(hdsm_map* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}
#define MCU(call)                                                                              \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) return mfail(h, HDSM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)
}  // namespace

extern "C" {

void hdsm_map_destroy(hdsm_map* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  cudaFree(h->d_inf);
  cudaFree(h->d_pot);
  cudaFree(h->d_pair);
  cudaFree(h->d_irow);
  cudaFree(h->d_pvals);
  cudaFree(h->d_tab);
  cudaFree(h->d_buf);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int hdsm_map_create(const hdsm_map_params* p, int max_grids, size_t grid_stride, int device, hdsm_map** out) {
  if (!p || !out || max_grids < 1 || grid_stride < 1) return HDSM_ERR_INVALID;
  if (!(p->voxel_size > 0) || p->inflation_dist < 0 || p->potential_dist < 0 || p->potential_pow < 0) return HDSM_ERR_INVALID;
  size_t smem = 2 * ((grid_stride + 15) & ~size_t(15));
  if (smem > 220 * 1024) return HDSM_ERR_INVALID;  // both copies of a grid (and the stencils) must fit into one SM's shared memory
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return HDSM_ERR_CUDA;  // no CPU fallback
  hdsm_map* h = new (std::nothrow) hdsm_map();
  if (!h) return HDSM_ERR_INVALID;
  h->prm = *p, h->device = device, h->max_grids = max_grids, h->grid_stride = grid_stride, h->smem = smem;
  h->cube = (int)std::ceil(p->inflation_dist / p->voxel_size);
  // stencils: inflation (values unused), potential (entries of value 0 change nothing: dropped; sorted by value)
  std::vector<hdsm_mp::MaskEntry> inf = hdsm_mp::create_mask(p->voxel_size, p->inflation_dist, 1);
  std::vector<hdsm_mp::MaskEntry> pot = hdsm_mp::create_mask(p->voxel_size, p->potential_dist, (double)p->potential_pow);
  const std::vector<hdsm_mp::MaskEntry> pot_full = pot;  // with the entries of value 0: the distance form below checks the whole mask
  pot.erase(std::remove_if(pot.begin(), pot.end(), [](const hdsm_mp::MaskEntry& e) { return e.v <= 0; }), pot.end());
  std::stable_sort(pot.begin(), pot.end(), [](const hdsm_mp::MaskEntry& a, const hdsm_mp::MaskEntry& b) { return a.v > b.v; });
  for (const auto& e : pot)
    if (e.v >= 100 && (e.x || e.y || e.z)) {  // would make the pass feed on itself
      delete h;
      return HDSM_ERR_INVALID;
    }
  const auto pack = [](const std::vector<hdsm_mp::MaskEntry>& m) {
    std::vector<int8_t> v(4 * m.size() + 4, 0);
    for (size_t i = 0; i < m.size(); ++i) v[4 * i] = (int8_t)m[i].x, v[4 * i + 1] = (int8_t)m[i].y, v[4 * i + 2] = (int8_t)m[i].z, v[4 * i + 3] = m[i].v;
    return v;
  };
  const std::vector<int8_t> hi = pack(inf), hp = pack(pot);
  h->n_inf = (int)inf.size(), h->n_pot = (int)pot.size();
  smem += 4 * (size_t)(h->n_inf + h->n_pot);
  if (smem > 227 * 1024 - 256) {
    delete h;
    return HDSM_ERR_INVALID;
  }
  // row form of the inflation stencil: per (dy, dz) the largest |dx| the stencil reaches
  std::vector<int> irow;
  h->rn_inf = p->inflation_dist > 0 ? (int)std::ceil(p->inflation_dist / p->voxel_size) : 0;
  if (h->rn_inf >= 1 && h->rn_inf <= 7) {
    for (int dy = -h->rn_inf; dy <= h->rn_inf; ++dy)
      for (int dz = -h->rn_inf; dz <= h->rn_inf; ++dz) {
        int reach = -1;
        for (const auto& e : inf)
          if (e.y == dy && e.z == dz) reach = std::max(reach, std::abs(e.x));
        bool solid = reach >= 0;  // the row form needs every |dx| <= reach to be in the stencil (bar the centre)
        for (int x = 0; x <= reach && solid; ++x) {
          if (x == 0 && dy == 0 && dz == 0) continue;
          bool in = false;
          for (const auto& e : inf) in |= e.y == dy && e.z == dz && e.x == x;
          solid = in;
        }
        if (reach >= 0 && !solid) {
          irow.clear();
          dy = dz = 99;  // an irregular stencil: keep the entry-by-entry form
          break;
        }
        if (reach >= 0) irow.push_back((dy & 0xff) | ((dz & 0xff) << 8) | (reach << 16));
      }
    h->n_irow = (int)irow.size();
    // the full cube, with or without its centre (an occupied voxel stays occupied either way; CreateMask leaves the
    // centre out when the inflation distance equals the voxel size, the reference's default): dilation by bit rows
    const int side = 2 * h->rn_inf + 1;
    bool centre = false;
    for (const auto& e : inf) centre |= e.x == 0 && e.y == 0 && e.z == 0;
    h->inf_cube = (int)inf.size() + (centre ? 0 : 1) == side * side * side && !std::getenv("HDSM_MAP_NO_DIST");
  }
  smem += 4 * (size_t)((h->n_irow + 3) & ~3);
  // row form of the potential stencil (used when the radius is at most 7 voxels and the bit rows fit)
  std::vector<int2> pair_yz;
  std::vector<unsigned long long> pair_vals;
  h->rn = p->potential_dist > 0 ? (int)std::ceil(p->potential_dist / p->voxel_size) : 0;
  if (h->rn >= 1 && h->rn <= 7) {
    struct Row {
      int dy, dz, best;
      unsigned long long vals;
    };
    std::vector<Row> rows;
    for (int dy = -h->rn; dy <= h->rn; ++dy)
      for (int dz = -h->rn; dz <= h->rn; ++dz) {
        Row r{dy, dz, 0, 0ull};
        for (const auto& e : pot)
          if (e.y == dy && e.z == dz && e.x >= 0) {
            r.vals |= (unsigned long long)(unsigned char)e.v << (8 * e.x);
            r.best = std::max(r.best, (int)e.v);
          }
        if (r.best > 0) rows.push_back(r);
      }
    std::stable_sort(rows.begin(), rows.end(), [](const Row& a, const Row& b) { return a.best > b.best; });
    for (const Row& r : rows) {
      pair_yz.push_back(make_int2((r.dy & 0xff) | ((r.dz & 0xff) << 8) | (r.best << 16), 0));
      pair_vals.push_back(r.vals);
    }
    h->n_pair = (int)rows.size();
  }
  // distance form of the potential stencil (see hdsm_mp::distance_table)
  std::vector<int8_t> tab(128, (int8_t)-128);
  const bool dist_ok = !std::getenv("HDSM_MAP_NO_DIST") && hdsm_mp::distance_table(pot_full, h->rn, tab.data());
  smem += 4 + 4 * (size_t)((h->n_pair + 1) & ~1) + 8 * (size_t)h->n_pair;  // the row table (and its alignment slack)
  if (smem + 1024 < 227 * 1024) {  // whatever shared memory is left holds the bit rows
    h->bits_bytes = (227 * 1024 - 256 - smem) & ~size_t(15);  // 256 bytes stay free for the kernel's static shared memory (mbarrier)
    smem += h->bits_bytes;
  }
  h->smem = smem;
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_inf, hi.size());
  if (e == cudaSuccess) e = cudaMalloc(&h->d_pot, hp.size());
  if (e == cudaSuccess) e = cudaMemcpy(h->d_inf, hi.data(), hi.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->d_pot, hp.data(), hp.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && h->n_irow) e = cudaMalloc(&h->d_irow, irow.size() * sizeof(int));
  if (e == cudaSuccess && h->n_irow) e = cudaMemcpy(h->d_irow, irow.data(), irow.size() * sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && h->n_pair) e = cudaMalloc(&h->d_pair, pair_yz.size() * sizeof(int2));
  if (e == cudaSuccess && h->n_pair) e = cudaMalloc(&h->d_pvals, pair_vals.size() * 8);
  if (e == cudaSuccess && h->n_pair) e = cudaMemcpy(h->d_pair, pair_yz.data(), pair_yz.size() * sizeof(int2), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && h->n_pair) e = cudaMemcpy(h->d_pvals, pair_vals.data(), pair_vals.size() * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && dist_ok) e = cudaMalloc(&h->d_tab, tab.size());
  if (e == cudaSuccess && dist_ok) e = cudaMemcpy(h->d_tab, tab.data(), tab.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = hdsm::raise_smem_limit(hdsm_mp::map_kernel, device);
  if (e != cudaSuccess) {
    hdsm_map_destroy(h);
    return HDSM_ERR_CUDA;
  }
  *out = h;
  return HDSM_OK;
}

int hdsm_map_distance_table(const hdsm_map_params* p, int8_t* tab128) {
  if (!p || !tab128 || !(p->voxel_size > 0) || p->potential_dist < 0) return HDSM_ERR_INVALID;
  const int rn = p->potential_dist > 0 ? (int)std::ceil(p->potential_dist / p->voxel_size) : 0;
  return hdsm_mp::distance_table(hdsm_mp::create_mask(p->voxel_size, p->potential_dist, (double)p->potential_pow), rn, tab128) ? 1 : 0;
}

const char* hdsm_map_last_error(const hdsm_map* h) { return h ? h->err.c_str() : "null handle"; }
int64_t hdsm_map_launch_count(const hdsm_map* h) { return h ? h->launches : 0; }

int hdsm_map_batch_device(hdsm_map* h, int n_grids, const int8_t* grids_in, const int32_t* dims, int8_t* grids_out, void* stream) {
  if (!h) return HDSM_ERR_INVALID;
  if (n_grids < 0 || !grids_in || !dims || !grids_out) return mfail(h, HDSM_ERR_INVALID, "null argument");
  if (n_grids > h->max_grids) return mfail(h, HDSM_ERR_CAPACITY, "n_grids exceeds max_grids");
  if (n_grids == 0) return HDSM_OK;
  MCU(cudaSetDevice(h->device));
  hdsm_mp::Args a{};
  a.n_grids = n_grids, a.cube = h->cube, a.n_inf = h->n_inf, a.n_pot = h->n_pot, a.stride = h->grid_stride;
  a.in = grids_in, a.out = grids_out, a.dims = dims, a.inf_off = h->d_inf, a.pot_off = h->d_pot;
  a.n_pair = h->n_pair, a.rn = h->rn, a.pair_yz = h->d_pair, a.pair_vals = h->d_pvals, a.bits_bytes = h->bits_bytes;
  a.n_irow = h->n_irow, a.rn_inf = h->rn_inf, a.irow = h->d_irow;
  a.dist_tab = h->d_tab, a.inf_cube = h->inf_cube;
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : h->stream;
  hdsm_mp::map_kernel<<<n_grids, hdsm_mp::kThreads, h->smem, s>>>(a);
  h->launches += 1;
  MCU(cudaGetLastError());
  return HDSM_OK;
}

int hdsm_map_batch(hdsm_map* h, int n_grids, const int8_t* grids_in, const int32_t* dims, int8_t* grids_out) {
  if (!h) return HDSM_ERR_INVALID;
  if (n_grids < 0 || !grids_in || !dims || !grids_out) return mfail(h, HDSM_ERR_INVALID, "null argument");
  if (n_grids > h->max_grids) return mfail(h, HDSM_ERR_CAPACITY, "n_grids exceeds max_grids");
  if (n_grids == 0) return HDSM_OK;
  for (int i = 0; i < n_grids; ++i) {
    const int32_t* d = dims + 3 * i;
    if (d[0] < 1 || d[1] < 1 || d[2] < 1 || (size_t)d[0] * d[1] * d[2] > h->grid_stride)
      return mfail(h, HDSM_ERR_INVALID, "grid dimensions exceed grid_stride");
  }
  MCU(cudaSetDevice(h->device));
  const size_t gb = (size_t)n_grids * h->grid_stride, db = ((size_t)n_grids * 12 + 255) & ~size_t(255);
  const size_t need = 2 * ((gb + 255) & ~size_t(255)) + db;
  if (need > h->buf_cap) {
    cudaFree(h->d_buf);
    h->d_buf = nullptr, h->buf_cap = 0;
    MCU(cudaMalloc(&h->d_buf, need));
    h->buf_cap = need;
  }
  int8_t* d_in = reinterpret_cast<int8_t*>(h->d_buf);
  int8_t* d_out = d_in + ((gb + 255) & ~size_t(255));
  int32_t* d_dims = reinterpret_cast<int32_t*>(d_out + ((gb + 255) & ~size_t(255)));
  MCU(cudaMemcpyAsync(d_in, grids_in, gb, cudaMemcpyHostToDevice, h->stream));
  MCU(cudaMemcpyAsync(d_dims, dims, (size_t)n_grids * 12, cudaMemcpyHostToDevice, h->stream));
  const int rc = hdsm_map_batch_device(h, n_grids, d_in, d_dims, d_out, h->stream);
  if (rc != HDSM_OK) return rc;
  MCU(cudaMemcpyAsync(grids_out, d_out, gb, cudaMemcpyDeviceToHost, h->stream));
  MCU(cudaStreamSynchronize(h->stream));
  return HDSM_OK;
}

}  // extern "C"
