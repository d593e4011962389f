// sense_emu.cpp - TEST INFRASTRUCTURE ONLY.  Runs the per-ray / per-voxel code of the acquisition kernel
// (multi_agent_pkgs_b200/csrc/hdsm_sense_core.h, the header csrc/hdsm_sense.cu is built from) on the CPU,
// following the kernel's structure: occupancy bits of the cropped grid, rays in an arbitrary (here: seeded,
// scrambled) order offering keys to voxels, largest key wins, then the merge pass (bits_form = 1: the kernel's second
// form - bitmaps of free / occupied writes first, keys only for voxels that received both).  tests/ compare it with the
// sequential restatement of the reference (sense_oracle.c): this is what shows, without a GPU, that the key scheme
// reproduces the reference's last-write-wins order and that the shared arithmetic matches.  Never shipped, never
// called by the product.  Compile with -ffp-contract=off.
#include <stdint.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "../multi_agent_pkgs_b200/csrc/hdsm_sense_core.h"

extern "C" int sense_emu_batch(double voxel, const double* range, int free_grid, int limited_fov, double cos_half_x, double cos_half_y, int n,
                               const int8_t* env, const int32_t* dim_env_in, const double* origin_env, const double* pos, const double* rot,
                               const int8_t* old_grids, const double* old_origin, const uint8_t* have_old, size_t stride, int8_t* out,
                               double* origin_out, unsigned seed, int bits_form, long long* conflicts) {
  using namespace hdsm_sn;
  const int dim_env[3] = {dim_env_in[0], dim_env_in[1], dim_env_in[2]};
  for (int a = 0; a < n; ++a) {
    Frame F;
    make_frame(voxel, range, origin_env, pos + 3 * a, F);
    const int cells = F.dim[0] * F.dim[1] * F.dim[2];
    if ((size_t)cells > stride) return -1;
    int8_t* o = out + stride * a;
    for (int c = 0; c < 3; ++c) origin_out[3 * a + c] = F.origin[c];
    if (free_grid) {
      for (int cell = 0; cell < cells; ++cell) {
        const int x = cell % F.dim[0], y = (cell / F.dim[0]) % F.dim[1], z = cell / (F.dim[0] * F.dim[1]);
        o[cell] = crop_value(env, dim_env, F.start, true, x, y, z);
      }
      continue;
    }
    std::vector<uint32_t> bits((cells + 31) / 32, 0u), keys(cells, 0u);
    for (int cell = 0; cell < cells; ++cell) {
      const int x = cell % F.dim[0], y = (cell / F.dim[0]) % F.dim[1], z = cell / (F.dim[0] * F.dim[1]);
      if (crop_value(env, dim_env, F.start, false, x, y, z) == 100) bits[cell >> 5] |= 1u << (cell & 31);
    }
    const int nr = ray_count(F.dim);
    std::vector<int> order(nr);
    for (int r = 0; r < nr; ++r) order[r] = r;
    unsigned st = seed * 2654435761u + 12345u + (unsigned)a;
    for (int r = nr - 1; r > 0; --r) {  // Fisher-Yates with an LCG
      st = st * 1664525u + 1013904223u;
      std::swap(order[r], order[(st >> 8) % (unsigned)(r + 1)]);
    }
    const int* dim = F.dim;
    const auto occ = [&](int x, int y, int z) { const int c = x + dim[0] * (y + dim[1] * z); return (bits[c >> 5] >> (c & 31)) & 1u; };
    std::vector<uint32_t> wfree(bits_form ? bits.size() : 0, 0u), wocc(bits_form ? bits.size() : 0, 0u);
    for (int pass = 0; pass < (bits_form ? 2 : 1); ++pass) {
      if (pass == 1) {  // the second form: keys only for voxels that received both kinds of write
        long long both = 0;
        for (size_t w = 0; w < bits.size(); ++w) both += __builtin_popcount(wfree[w] & wocc[w]);
        if (conflicts) conflicts[a] = both;
        if (!both) break;
      }
      for (int q = 0; q < nr; ++q) {
        const int seq = order[q];
        double end[3];
        ray_end(dim, seq, end);
        if (limited_fov && !in_fov(rot + 9 * a, F.pos_local, end, cos_half_x, cos_half_y)) continue;
        if (!bits_form)
          clear_line(dim, F.pos_local, end, seq, occ, [&](int cell, uint32_t key) { if (keys[cell] < key) keys[cell] = key; });
        else if (pass == 0)
          clear_line(dim, F.pos_local, end, seq, occ, [&](int cell, uint32_t key) { ((key & 1u) ? wfree : wocc)[cell >> 5] |= 1u << (cell & 31); });
        else
          clear_line(dim, F.pos_local, end, seq, occ, [&](int cell, uint32_t key) {
            if ((((wfree[cell >> 5] & wocc[cell >> 5]) >> (cell & 31)) & 1u) && keys[cell] < key) keys[cell] = key;
          });
      }
    }
    if (bits_form)  // single-kind voxels: the kind decides (expressed as a key so that the merge below is shared)
      for (int cell = 0; cell < cells; ++cell) {
        const uint32_t f = (wfree[cell >> 5] >> (cell & 31)) & 1u, o = (wocc[cell >> 5] >> (cell & 31)) & 1u;
        if (f != o) keys[cell] = f ? key_free(0) : key_occ(0);
      }
    const bool ho = have_old && have_old[a];
    int off[3] = {0, 0, 0}, mid[3];
    if (ho) merge_offset(F.origin, old_origin + 3 * a, voxel, off);
    for (int c = 0; c < 3; ++c) mid[c] = (int)floor(F.pos_local[c]);
    for (int cell = 0; cell < cells; ++cell) {
      const int x = cell % dim[0], y = (cell / dim[0]) % dim[1], z = cell / (dim[0] * dim[1]);
      int8_t v = key_value(keys[cell]);
      if (v == -1) v = old_value(ho ? old_grids + stride * a : nullptr, ho, dim, off, mid, x, y, z);
      o[cell] = v;
    }
  }
  return 0;
}
