// Local-map acquisition for sm_100a: one thread block per agent, SURVEY.md section 8(f) row 4 (the part in front
// of csrc/hdsm_map.cu).
//
// Replaces, per agent and per map update (mapping_util/src/map_builder.cpp):
//   MapBuilder::EnvironmentVoxelGridCallback  :80-205   frame, crop of the environment grid, first-update grid
//   MapBuilder::RaycastAndClear               :280-329  one ray to every border voxel of the local grid
//   MapBuilder::ClearLine                     :367-432  field-of-view test, ray traversal, occupied / free writes
//   MapBuilder::MergeVoxelGrids               :242-278  unknown voxels keep what the previous grid knew
//   MapBuilder::ClearVoxelsCenter             :434-447
//   voxel_grid_util::Raycast                  voxel_grid_util/src/raycast.cpp:21-186 (via path_finding_util::IsLineClear)
// and writes voxel_grid_curr_ in the layout hdsm_map_batch_device reads.
//
// Design.  All agents of a swarm look at ONE environment grid (it stays in L2: 140 x 140 x 24 voxels = 470 KB), so
// the compulsory HBM traffic per agent is the kept grid in (87 KB) and the new grid out (87 KB).  A block first
// packs the occupancy of its crop into shared memory, one bit per voxel (11 KB for 66 x 66 x 20): that is all the
// ~14 000 rays need to read.  Rays are dealt to threads; each ray offers the key 2 (seq + 1) + is_free to every
// voxel it writes and the voxel keeps the maximum (atomicMax on a per-block u32 scratch that lives in L2), which
// reproduces the reference's sequential last-write-wins result in any execution order (hdsm_sense_core.h).  Rays
// run from the LAST of the reference's order to the first, so that most later offers lose against what is already
// there and are dropped after a plain load instead of an atomic.  The merge pass then turns keys into voxel
// values, fills unknown voxels from the kept grid, stores the new grid with coalesced writes and clears the
// scratch for the block's next agent.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdlib>
#include <new>
#include <string>

#include "../../include/hdsm.h"
#include "hdsm_common.h"
#include "hdsm_sense_core.h"

namespace hdsm_sn {

constexpr int kThreads = 512;

struct Args {
  double voxel, range[3], origin_env[3], cos_half_x, cos_half_y;
  int free_grid, limited_fov, n;
  int dim_env[3];
  size_t stride;
  const int8_t* env;
  const double *pos, *rot, *old_origin;
  const int8_t* old_grids;
  const uint8_t* have_old;
  int8_t* out;
  double* origin_out;
  uint32_t* keys;  // [gridDim.x][stride]
};

__global__ void __launch_bounds__(kThreads) sense_kernel(const Args A) {
  extern __shared__ uint32_t s_bits[];  // occupancy of the crop, one bit per voxel
  __shared__ Frame F;
  __shared__ int s_off[3], s_mid[3];
  uint32_t* keys = A.keys + (size_t)blockIdx.x * A.stride;
  const int tid = threadIdx.x;

  for (int a = blockIdx.x; a < A.n; a += gridDim.x) {
    if (tid == 0) {
      const double p[3] = {A.pos[3 * a], A.pos[3 * a + 1], A.pos[3 * a + 2]};
      make_frame(A.voxel, A.range, A.origin_env, p, F);
      const bool ho = A.have_old && A.have_old[a];
      int off[3] = {0, 0, 0};
      if (ho) {
        const double oo[3] = {A.old_origin[3 * a], A.old_origin[3 * a + 1], A.old_origin[3 * a + 2]};
        merge_offset(F.origin, oo, A.voxel, off);
      }
      for (int c = 0; c < 3; ++c) {
        s_off[c] = off[c], s_mid[c] = (int)floor(F.pos_local[c]);
        A.origin_out[3 * a + c] = F.origin[c];
      }
    }
    __syncthreads();
    const int dim[3] = {F.dim[0], F.dim[1], F.dim[2]};
    const int start[3] = {F.start[0], F.start[1], F.start[2]};
    const int cells = dim[0] * dim[1] * dim[2], plane = dim[0] * dim[1];
    int8_t* out = A.out + (size_t)a * A.stride;

    if (A.free_grid) {  // no sensing model: the crop itself, unknown -> free (:139-142, :201-204)
      for (int cell = tid; cell < cells; cell += kThreads) {
        const int z = cell / plane, r = cell - z * plane, y = r / dim[0], x = r - y * dim[0];
        out[cell] = crop_value(A.env, A.dim_env, start, true, x, y, z);
      }
      __syncthreads();
      continue;
    }

    // 1. occupancy bits of the crop (one warp ballot per 32 voxels: coalesced along x)
    const int words = (cells + 31) >> 5;
    for (int base = (tid >> 5) << 5; base < words * 32; base += kThreads) {
      const int cell = base + (tid & 31);
      bool o = false;
      if (cell < cells) {
        const int z = cell / plane, r = cell - z * plane, y = r / dim[0], x = r - y * dim[0];
        const int ie = x + start[0], je = y + start[1], ke = z + start[2];
        o = inside(A.dim_env, ie, je, ke) && __ldg(A.env + (size_t)ie + (size_t)A.dim_env[0] * ((size_t)je + (size_t)A.dim_env[1] * ke)) == 100;
      }
      const unsigned m = __ballot_sync(0xffffffffu, o);
      if ((tid & 31) == 0) s_bits[base >> 5] = m;
    }
    __syncthreads();

    // 2. rays, last of the reference's order first
    const double sp[3] = {F.pos_local[0], F.pos_local[1], F.pos_local[2]};
    const int nr = ray_count(dim);
    const double* rot = A.limited_fov ? A.rot + 9 * a : nullptr;
    for (int seq = nr - 1 - tid; seq >= 0; seq -= kThreads) {
      double end[3];
      ray_end(dim, seq, end);
      if (rot && !in_fov(rot, sp, end, A.cos_half_x, A.cos_half_y)) continue;
      clear_line(
          dim, sp, end, seq,
          [&](int x, int y, int z) {
            const int c = x + dim[0] * (y + dim[1] * z);
            return (s_bits[c >> 5] >> (c & 31)) & 1u;
          },
          [&](int cell, uint32_t key) {
            if (__ldcg(keys + cell) < key) atomicMax(keys + cell, key);
          });
    }
    __syncthreads();

    // 3. keys -> values, merge with the kept grid, clear the scratch
    const int8_t* old_grid = A.old_grids ? A.old_grids + (size_t)a * A.stride : nullptr;
    const bool ho = A.have_old && A.have_old[a] && old_grid;
    const int off[3] = {s_off[0], s_off[1], s_off[2]}, mid[3] = {s_mid[0], s_mid[1], s_mid[2]};
    for (int cell = tid; cell < cells; cell += kThreads) {
      const uint32_t key = __ldcg(keys + cell);
      __stcg(keys + cell, 0u);
      int8_t v = key_value(key);
      if (v == -1) {
        const int z = cell / plane, r = cell - z * plane, y = r / dim[0], x = r - y * dim[0];
        v = old_value(old_grid, ho, dim, off, mid, x, y, z);
      }
      out[cell] = v;
    }
    __syncthreads();
  }
}

// ---- second form of the kernel: the default (HDSM_SENSE_BITS=0 when the handle is created selects the first form, kept
// for A/B measurements).  Measured on B200: 33.6 ms per 4096 agents against 63.9 ms.  The profile of the first form (profiles/r1u_ncu_sense_summary.txt)
// has 27 % of its stall samples on the L2 round trip of the key load in front of the atomicMax.  Keys only matter for
// voxels that receive BOTH kinds of write (occupied from one ray, free from another) - a numerical corner case of the
// reference's nudged collision point.  So the ray loop here only sets one bit per voxel in shared memory (free-written /
// occupied-written); if the two bitmaps intersect anywhere, the agent's rays are traversed a second time and offer their
// keys to the intersecting voxels only.  Same result as the first form by construction: a voxel written by one kind of
// write has that kind's value whatever the order, and the others are decided by the same largest-key rule.
constexpr int kThreadsBits = 256;

__global__ void __launch_bounds__(kThreadsBits, 3) sense_kernel_bits(const Args A) {
  extern __shared__ uint32_t s_words[];  // [3][words]: occupancy of the crop, free-written, occupied-written
  __shared__ Frame F;
  __shared__ int s_off[3], s_mid[3];
  uint32_t* keys = A.keys + (size_t)blockIdx.x * A.stride;
  const int tid = threadIdx.x;

  for (int a = blockIdx.x; a < A.n; a += gridDim.x) {
    if (tid == 0) {
      const double p[3] = {A.pos[3 * a], A.pos[3 * a + 1], A.pos[3 * a + 2]};
      make_frame(A.voxel, A.range, A.origin_env, p, F);
      const bool ho = A.have_old && A.have_old[a];
      int off[3] = {0, 0, 0};
      if (ho) {
        const double oo[3] = {A.old_origin[3 * a], A.old_origin[3 * a + 1], A.old_origin[3 * a + 2]};
        merge_offset(F.origin, oo, A.voxel, off);
      }
      for (int c = 0; c < 3; ++c) {
        s_off[c] = off[c], s_mid[c] = (int)floor(F.pos_local[c]);
        A.origin_out[3 * a + c] = F.origin[c];
      }
    }
    __syncthreads();
    const int dim[3] = {F.dim[0], F.dim[1], F.dim[2]};
    const int start[3] = {F.start[0], F.start[1], F.start[2]};
    const int cells = dim[0] * dim[1] * dim[2], plane = dim[0] * dim[1];
    int8_t* out = A.out + (size_t)a * A.stride;

    if (A.free_grid) {
      for (int cell = tid; cell < cells; cell += kThreadsBits) {
        const int z = cell / plane, r = cell - z * plane, y = r / dim[0], x = r - y * dim[0];
        out[cell] = crop_value(A.env, A.dim_env, start, true, x, y, z);
      }
      __syncthreads();
      continue;
    }

    const int words = (cells + 31) >> 5;
    uint32_t *s_bits = s_words, *s_free = s_words + words, *s_occw = s_words + 2 * words;
    for (int base = (tid >> 5) << 5; base < words * 32; base += kThreadsBits) {
      const int cell = base + (tid & 31);
      bool o = false;
      if (cell < cells) {
        const int z = cell / plane, r = cell - z * plane, y = r / dim[0], x = r - y * dim[0];
        const int ie = x + start[0], je = y + start[1], ke = z + start[2];
        o = inside(A.dim_env, ie, je, ke) && __ldg(A.env + (size_t)ie + (size_t)A.dim_env[0] * ((size_t)je + (size_t)A.dim_env[1] * ke)) == 100;
      }
      const unsigned m = __ballot_sync(0xffffffffu, o);
      if ((tid & 31) == 0) s_bits[base >> 5] = m, s_free[base >> 5] = 0u, s_occw[base >> 5] = 0u;
    }
    __syncthreads();

    const double sp[3] = {F.pos_local[0], F.pos_local[1], F.pos_local[2]};
    const int nr = ray_count(dim);
    const double* rot = A.limited_fov ? A.rot + 9 * a : nullptr;
    const auto occ = [&](int x, int y, int z) {
      const int c = x + dim[0] * (y + dim[1] * z);
      return (s_bits[c >> 5] >> (c & 31)) & 1u;
    };
    // pass 1: which voxels are written free, which occupied
    for (int seq = nr - 1 - tid; seq >= 0; seq -= kThreadsBits) {
      double end[3];
      ray_end(dim, seq, end);
      if (rot && !in_fov(rot, sp, end, A.cos_half_x, A.cos_half_y)) continue;
      clear_line(dim, sp, end, seq, occ, [&](int cell, uint32_t key) {
        uint32_t* w = ((key & 1u) ? s_free : s_occw) + (cell >> 5);
        const uint32_t m = 1u << (cell & 31);
        if (!(*(volatile uint32_t*)w & m)) atomicOr(w, m);
      });
    }
    __syncthreads();
    int both = 0;
    for (int w = tid; w < words; w += kThreadsBits) both |= (s_free[w] & s_occw[w]) != 0u;
    // pass 2 (rare): the voxels with both kinds of write are decided by the largest key
    if (__syncthreads_or(both)) {
      for (int seq = nr - 1 - tid; seq >= 0; seq -= kThreadsBits) {
        double end[3];
        ray_end(dim, seq, end);
        if (rot && !in_fov(rot, sp, end, A.cos_half_x, A.cos_half_y)) continue;
        clear_line(dim, sp, end, seq, occ, [&](int cell, uint32_t key) {
          if (((s_free[cell >> 5] & s_occw[cell >> 5]) >> (cell & 31)) & 1u) atomicMax(keys + cell, key);
        });
      }
      __syncthreads();
    }

    // merge: four voxels per thread where the layout allows 32-bit stores
    const int8_t* old_grid = A.old_grids ? A.old_grids + (size_t)a * A.stride : nullptr;
    const bool ho = A.have_old && A.have_old[a] && old_grid;
    const int off[3] = {s_off[0], s_off[1], s_off[2]}, mid[3] = {s_mid[0], s_mid[1], s_mid[2]};
    const auto value_of = [&](int cell) -> int8_t {
      const uint32_t f = (s_free[cell >> 5] >> (cell & 31)) & 1u, o = (s_occw[cell >> 5] >> (cell & 31)) & 1u;
      int8_t v;
      if (f & o) {
        v = key_value(__ldcg(keys + cell));
        __stcg(keys + cell, 0u);
      } else {
        v = o ? (int8_t)100 : (f ? (int8_t)0 : (int8_t)-1);
      }
      if (v == -1) {
        const int z = cell / plane, r = cell - z * plane, y = r / dim[0], x = r - y * dim[0];
        v = old_value(old_grid, ho, dim, off, mid, x, y, z);
      }
      return v;
    };
    if ((cells & 3) == 0 && (A.stride & 3) == 0 && ((size_t)A.out & 3) == 0) {
      uint32_t* out4 = reinterpret_cast<uint32_t*>(out);
      for (int q = tid; q < (cells >> 2); q += kThreadsBits) {
        uint32_t pack = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) pack |= (uint32_t)(uint8_t)value_of(4 * q + j) << (8 * j);
        out4[q] = pack;
      }
    } else {
      for (int cell = tid; cell < cells; cell += kThreadsBits) out[cell] = value_of(cell);
    }
    __syncthreads();
  }
}

__global__ void zero_keys(uint32_t* k, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) k[i] = 0u;
}

}  // namespace hdsm_sn

struct hdsm_sense {
  hdsm_sense_params prm{};
  int device = 0, max_agents = 0, blocks = 0, dim[3] = {0, 0, 0}, bits = 0;
  size_t grid_stride = 0, env_cap = 0, smem = 0;
  double cos_half_x = 0, cos_half_y = 0;
  cudaStream_t stream = nullptr;
  uint32_t* d_keys = nullptr;
  unsigned char* d_buf = nullptr;
  size_t buf_cap = 0;
  int64_t launches = 0;
  std::string err;
};

namespace {
int sfail(hdsm_sense* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}
#define SCU(call)                                                                              \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) return sfail(h, HDSM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)
size_t up256(size_t x) { return (x + 255) & ~size_t(255); }
}  // namespace

extern "C" {

void hdsm_sense_destroy(hdsm_sense* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  cudaFree(h->d_keys);
  cudaFree(h->d_buf);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int hdsm_sense_grid_dims(const hdsm_sense_params* p, int32_t dim[3]) {
  if (!p || !dim || !(p->voxel_size > 0)) return HDSM_ERR_INVALID;
  for (int a = 0; a < 3; ++a) {
    if (!(p->range[a] > 0)) return HDSM_ERR_INVALID;
    dim[a] = (int32_t)std::floor(p->range[a] / p->voxel_size);  // map_builder.cpp:103-107
    if (dim[a] < 1) return HDSM_ERR_INVALID;
  }
  return HDSM_OK;
}

int hdsm_sense_create(const hdsm_sense_params* p, int max_agents, size_t grid_stride, int device, hdsm_sense** out) {
  if (!p || !out || max_agents < 1) return HDSM_ERR_INVALID;
  int32_t dim[3];
  if (hdsm_sense_grid_dims(p, dim) != HDSM_OK) return HDSM_ERR_INVALID;
  const size_t cells = (size_t)dim[0] * dim[1] * dim[2];
  // a ray visits at most dx + dy + dz + 1 voxels; the reference throws beyond 1500 (raycast.cpp:146-149)
  if (grid_stride < cells || dim[0] + dim[1] + dim[2] > 1400 || cells > (size_t)1 << 30) return HDSM_ERR_INVALID;
  if (p->limited_fov && !(p->fov_x > 0 && p->fov_y > 0)) return HDSM_ERR_INVALID;
  const char* form = std::getenv("HDSM_SENSE_BITS");  // "0" selects the first (key) form, for A/B measurements
  const int bits = !(form && form[0] == '0');
  const size_t smem = ((cells + 31) / 32) * 4 * (bits ? 3 : 1);
  if (smem > 200 * 1024) return HDSM_ERR_INVALID;  // the crop's bitmaps must fit into one SM's shared memory
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return HDSM_ERR_CUDA;  // no CPU fallback
  hdsm_sense* h = new (std::nothrow) hdsm_sense();
  if (!h) return HDSM_ERR_INVALID;
  h->prm = *p, h->device = device, h->max_agents = max_agents, h->grid_stride = grid_stride, h->smem = smem, h->bits = bits;
  h->dim[0] = dim[0], h->dim[1] = dim[1], h->dim[2] = dim[2];
  h->cos_half_x = std::cos(p->fov_x / 2), h->cos_half_y = std::cos(p->fov_y / 2);  // :390-391
  cudaError_t e = cudaSetDevice(device);
  int sms = 0, per_sm = 0;
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (e == cudaSuccess && !bits) e = hdsm::raise_smem_limit(hdsm_sn::sense_kernel, device);
  if (e == cudaSuccess && !bits) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hdsm_sn::sense_kernel, hdsm_sn::kThreads, smem);
  if (e == cudaSuccess && bits) e = hdsm::raise_smem_limit(hdsm_sn::sense_kernel_bits, device);
  if (e == cudaSuccess && bits)
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hdsm_sn::sense_kernel_bits, hdsm_sn::kThreadsBits, smem);
  if (e == cudaSuccess) {
    if (per_sm < 1) per_sm = 1;
    h->blocks = sms * per_sm;  // persistent: every resident block owns one key scratch
    if (h->blocks > max_agents) h->blocks = max_agents;
    e = cudaMalloc(&h->d_keys, (size_t)h->blocks * grid_stride * sizeof(uint32_t));
  }
  if (e == cudaSuccess) {
    hdsm_sn::zero_keys<<<256, 256, 0, h->stream>>>(h->d_keys, (size_t)h->blocks * grid_stride);
    e = cudaStreamSynchronize(h->stream);
  }
  if (e != cudaSuccess) {
    hdsm_sense_destroy(h);
    return HDSM_ERR_CUDA;
  }
  *out = h;
  return HDSM_OK;
}

const char* hdsm_sense_last_error(const hdsm_sense* h) { return h ? h->err.c_str() : "null handle"; }
int64_t hdsm_sense_launch_count(const hdsm_sense* h) { return h ? h->launches : 0; }

int hdsm_sense_batch_device(hdsm_sense* h, int n, const int8_t* env, const int32_t dim_env[3], const double origin_env[3], const double* pos,
                            const double* rot, const int8_t* old_grids, const double* old_origin, const uint8_t* have_old, int8_t* grids_out,
                            double* origin_out, void* stream) {
  if (!h) return HDSM_ERR_INVALID;
  if (n < 0 || !env || !dim_env || !origin_env || !pos || !grids_out || !origin_out) return sfail(h, HDSM_ERR_INVALID, "null argument");
  if (dim_env[0] < 1 || dim_env[1] < 1 || dim_env[2] < 1) return sfail(h, HDSM_ERR_INVALID, "environment dimensions");
  if (h->prm.limited_fov && !h->prm.free_grid && !rot) return sfail(h, HDSM_ERR_INVALID, "limited_fov needs the camera rotation");
  if ((old_grids != nullptr) != (old_origin != nullptr)) return sfail(h, HDSM_ERR_INVALID, "old_grids and old_origin go together");
  if (old_grids && !have_old) return sfail(h, HDSM_ERR_INVALID, "have_old is required with old_grids");
  if (grids_out == old_grids) return sfail(h, HDSM_ERR_INVALID, "grids_out may not alias old_grids");
  if (n > h->max_agents) return sfail(h, HDSM_ERR_CAPACITY, "n exceeds max_agents");
  if (n == 0) return HDSM_OK;
  SCU(cudaSetDevice(h->device));
  hdsm_sn::Args a{};
  a.voxel = h->prm.voxel_size, a.cos_half_x = h->cos_half_x, a.cos_half_y = h->cos_half_y;
  for (int c = 0; c < 3; ++c) a.range[c] = h->prm.range[c], a.origin_env[c] = origin_env[c], a.dim_env[c] = dim_env[c];
  a.free_grid = h->prm.free_grid != 0, a.limited_fov = h->prm.limited_fov != 0, a.n = n, a.stride = h->grid_stride;
  a.env = env, a.pos = pos, a.rot = rot, a.old_grids = old_grids, a.old_origin = old_origin;
  a.have_old = old_grids ? have_old : nullptr;  // without kept grids every agent is on its first update
  a.out = grids_out, a.origin_out = origin_out, a.keys = h->d_keys;
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : h->stream;
  const int blocks = n < h->blocks ? n : h->blocks;
  if (h->bits)
    hdsm_sn::sense_kernel_bits<<<blocks, hdsm_sn::kThreadsBits, h->smem, s>>>(a);
  else
    hdsm_sn::sense_kernel<<<blocks, hdsm_sn::kThreads, h->smem, s>>>(a);
  h->launches += 1;
  SCU(cudaGetLastError());
  return HDSM_OK;
}

int hdsm_sense_batch(hdsm_sense* h, int n, const int8_t* env, const int32_t dim_env[3], const double origin_env[3], const double* pos,
                     const double* rot, const int8_t* old_grids, const double* old_origin, const uint8_t* have_old, int8_t* grids_out,
                     double* origin_out) {
  if (!h) return HDSM_ERR_INVALID;
  if (n < 0 || !env || !dim_env || !origin_env || !pos || !grids_out || !origin_out) return sfail(h, HDSM_ERR_INVALID, "null argument");
  if (dim_env[0] < 1 || dim_env[1] < 1 || dim_env[2] < 1) return sfail(h, HDSM_ERR_INVALID, "environment dimensions");
  if ((old_grids != nullptr) != (old_origin != nullptr)) return sfail(h, HDSM_ERR_INVALID, "old_grids and old_origin go together");
  if (old_grids && !have_old) return sfail(h, HDSM_ERR_INVALID, "have_old is required with old_grids");
  if (n > h->max_agents) return sfail(h, HDSM_ERR_CAPACITY, "n exceeds max_agents");
  if (n == 0) return HDSM_OK;
  SCU(cudaSetDevice(h->device));
  const size_t eb = (size_t)dim_env[0] * dim_env[1] * dim_env[2], gb = (size_t)n * h->grid_stride;
  const size_t need = up256(eb) + 2 * up256(gb) + 3 * up256((size_t)n * 24) + up256((size_t)n * 72) + up256((size_t)n);
  if (need > h->buf_cap) {
    cudaFree(h->d_buf);
    h->d_buf = nullptr, h->buf_cap = 0;
    SCU(cudaMalloc(&h->d_buf, need));
    h->buf_cap = need;
  }
  unsigned char* q = h->d_buf;
  const auto take = [&](size_t bytes) {
    unsigned char* r = q;
    q += up256(bytes);
    return r;
  };
  int8_t* d_env = reinterpret_cast<int8_t*>(take(eb));
  int8_t* d_old = reinterpret_cast<int8_t*>(take(gb));
  int8_t* d_out = reinterpret_cast<int8_t*>(take(gb));
  double* d_pos = reinterpret_cast<double*>(take((size_t)n * 24));
  double* d_oorg = reinterpret_cast<double*>(take((size_t)n * 24));
  double* d_rot = reinterpret_cast<double*>(take((size_t)n * 72));
  uint8_t* d_have = reinterpret_cast<uint8_t*>(take((size_t)n));
  double* d_org = reinterpret_cast<double*>(take((size_t)n * 24));
  SCU(cudaMemcpyAsync(d_env, env, eb, cudaMemcpyHostToDevice, h->stream));
  SCU(cudaMemcpyAsync(d_pos, pos, (size_t)n * 24, cudaMemcpyHostToDevice, h->stream));
  if (rot) SCU(cudaMemcpyAsync(d_rot, rot, (size_t)n * 72, cudaMemcpyHostToDevice, h->stream));
  if (old_grids) {
    SCU(cudaMemcpyAsync(d_old, old_grids, gb, cudaMemcpyHostToDevice, h->stream));
    SCU(cudaMemcpyAsync(d_oorg, old_origin, (size_t)n * 24, cudaMemcpyHostToDevice, h->stream));
    SCU(cudaMemcpyAsync(d_have, have_old, (size_t)n, cudaMemcpyHostToDevice, h->stream));
  }
  const int rc = hdsm_sense_batch_device(h, n, d_env, dim_env, origin_env, d_pos, rot ? d_rot : nullptr, old_grids ? d_old : nullptr,
                                         old_grids ? d_oorg : nullptr, old_grids ? d_have : nullptr, d_out, d_org, h->stream);
  if (rc != HDSM_OK) return rc;
  SCU(cudaMemcpyAsync(grids_out, d_out, gb, cudaMemcpyDeviceToHost, h->stream));
  SCU(cudaMemcpyAsync(origin_out, d_org, (size_t)n * 24, cudaMemcpyDeviceToHost, h->stream));
  SCU(cudaStreamSynchronize(h->stream));
  return HDSM_OK;
}

}  // extern "C"
