// C ABI of libhdsm (see include/hdsm.h).  Host side only: argument checks, device memory,
// one pinned staging arena per direction, kernel dispatch on the horizon length, and the NCCL
// trajectory exchange (NCCL is resolved with dlopen so the library loads without it).
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/hdsm.h"
#include "hdsm_common.h"
#include "hdsm_kernel.cuh"

using namespace hdsm;

namespace {

struct Id128 {  // ncclUniqueId, passed by value
  char internal[128];
};
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, Id128, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.lib ? &api : nullptr;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) return nullptr;
  api.GetUniqueId = reinterpret_cast<int (*)(void*)>(dlsym(api.lib, "ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<int (*)(void**, int, Id128, int)>(dlsym(api.lib, "ncclCommInitRank"));
  api.AllGather =
      reinterpret_cast<int (*)(const void*, void*, size_t, int, void*, cudaStream_t)>(dlsym(api.lib, "ncclAllGather"));
  api.GroupStart = reinterpret_cast<int (*)()>(dlsym(api.lib, "ncclGroupStart"));
  api.GroupEnd = reinterpret_cast<int (*)()>(dlsym(api.lib, "ncclGroupEnd"));
  api.CommDestroy = reinterpret_cast<int (*)(void*)>(dlsym(api.lib, "ncclCommDestroy"));
  api.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(api.lib, "ncclGetErrorString"));
  if (!api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.CommDestroy) {
    api.lib = nullptr;
    return nullptr;
  }
  return &api;
}

}  // namespace

constexpr int kMaxChunks = 16;

struct hdsm_handle {
  hdsm_params prm{};
  Tables host_tables{};
  Tables* dev_tables = nullptr;
  int device = 0, max_agents = 0, max_neighbours = 0;
  // row-pool tiers: pass t re-solves the agents that overflowed pass t-1 with a larger pool (fewer
  // resident blocks per SM); the last tier is the worst case or what one SM can hold
  int n_tiers = 0, row_cap[3] = {0, 0, 0}, smem_bytes[3] = {0, 0, 0};
  cudaStream_t stream = nullptr, stream2 = nullptr;  // stream2: second lane of the chunked host-pointer pipeline
  cudaEvent_t ev_shared = nullptr, ev_chunk[kMaxChunks] = {}, ev_begin[kMaxChunks] = {};
  cudaStream_t stream_early = nullptr;  // head start of the agents expected to search long (see launch)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int early_div = 16;  // HDSM_EARLY_DIV: the first n_local / early_div agents of the dispatch order get the head start (0: none)
  // staging for the host-pointer entry point
  unsigned char *h_in = nullptr, *d_in = nullptr, *h_out = nullptr, *d_out = nullptr;
  size_t in_cap = 0, out_cap = 0;
  int64_t launches = 0;
  // longest-first dispatch: order[slot] of the previous call per pipeline chunk (valid while n matches)
  double* d_bounds = nullptr;  // [bounds_cap][4] plan spheres of the neighbour table, rebuilt every call
  int bounds_cap = 0;
  int32_t* d_order = nullptr;
  int order_n[kMaxChunks] = {};         // n_local the cached order of a slot was computed for (0: none)
  size_t order_off[kMaxChunks] = {};    // ... and its offset into d_order: both must match for the order to be reused
  bool use_order = true;
  bool smem_configured = false;
  long long* d_prof = nullptr;  // HDSM_PROFILE=1: per-agent phase cycle counters (host path prints a summary)
  int warps = 4;  // warps per agent (HDSM_WARPS=1 selects the single-warp kernel)
  int block_slots = 592;  // resident solver blocks of the device (SMs x HDSM_MINBLOCKS)
  int force_csize = 0;    // HDSM_CLUSTER: blocks per agent, overriding the batch-size rule (experiments)
  int first_rounds = 2;      // rounds an agent gets in the first pass of a large batch before it is left to the cluster pass (HDSM_FIRST_ROUNDS)
  bool single_pass = false;  // HDSM_SINGLE_PASS=1: no cluster pass for the long searches of a large batch (experiments)
  std::string err;
  void* comm = nullptr;
  int n_ranks = 1;
};

namespace {

int fail(hdsm_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}
int cuda_fail(hdsm_handle* h, cudaError_t e, const char* what) {
  return fail(h, HDSM_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(call)                                              \
  do {                                                        \
    cudaError_t e_ = (call);                                  \
    if (e_ != cudaSuccess) return cuda_fail(h, e_, #call);    \
  } while (0)

// Thread blocks per agent: the nodes of a search round are dealt to a cluster of up to `width` blocks when the batch
// is small enough that the extra blocks find room (a hard agent then finishes up to `width` times sooner and the
// launch is as long as its hardest agent); large batches keep one block per agent.  Results do not depend on this
// choice (see Solver::run).
int blocks_per_agent(const hdsm_handle* h, int n_local) {
  if (h->warps != 4) return 1;
  int csize = h->prm.search_width;
  while (csize > 1 && (long)n_local * csize > 4L * h->block_slots) csize >>= 1;
  if (h->force_csize > 0) csize = std::min(h->force_csize, h->prm.search_width);
  return csize;
}

template <int N, int W>
cudaError_t launch(hdsm_handle* h, KernelArgs a, cudaStream_t s) {
  if (!h->smem_configured) {  // raised to the device maximum, never to this handle's own need (see hdsm_common.h)
    cudaError_t e = raise_smem_limit(hdsm_solve_kernel<N, W>, h->device);
    if (e != cudaSuccess) return e;
    h->smem_configured = true;
  }
  const int csize = W == 4 ? blocks_per_agent(h, a.n_local) : 1;
  // When the batch is too large for a cluster per agent, the few agents whose search is long would still set
  // the length of the launch.  They are picked out instead: the first pass (one block per agent) gives up on an
  // agent after kFirstPassRounds rounds, and a cluster pass redoes those agents alone with `width` blocks each.
  // Rounds are deterministic, so the redone search is the same search.
  const int kFirstPassRounds = h->first_rounds;
  const bool two_pass = W == 4 && csize < a.width && !h->single_pass;
  auto go = [&](int cs, unsigned mask, int budget, int ovf, int tier, int slot0 = 0, int slot1 = -1, cudaStream_t on = nullptr) -> cudaError_t {
    KernelArgs k = a;
    k.csize = cs, k.only_mask = mask, k.round_budget = budget, k.overflow_status = ovf, k.row_cap = h->row_cap[tier];
    k.slot0 = slot0;
    if (slot1 >= 0) k.n_local = slot1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(k.n_local - slot0) * cs), cfg.blockDim = dim3(32 * W), cfg.dynamicSmemBytes = h->smem_bytes[tier], cfg.stream = on ? on : s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
    cfg.attrs = at, cfg.numAttrs = cs > 1 ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, hdsm_solve_kernel<N, W>, (const Tables*)h->dev_tables, k);
    h->launches += 1;
    return e;
  };
  // Head start: hardness persists from one replanning step to the next, and the dispatch order of this call is last
  // call's iteration count, longest first.  Its first few agents are the ones a first pass would give up on, and the
  // cluster pass that redoes them is as long as the longest search - after a first pass that was just as long.  So
  // they skip the first pass: a cluster launch for the head of the order starts on a second (high-priority) stream
  // while the first pass works through the rest, and the two overlap.  Agents the prediction missed are still picked
  // up by the cluster pass; which launch solves an agent never changes its result.
  const int n_early = (two_pass && a.order && h->early_div > 0 && h->stream_early) ? a.n_local / h->early_div : 0;
  for (int t = 0; t < h->n_tiers; ++t) {  // t > 0: only agents whose rows did not fit the previous pool
    const bool last = t == h->n_tiers - 1;
    cudaError_t e;
    if (t == 0 && n_early > 0) {
      if ((e = cudaEventRecord(h->ev_fork, s)) != cudaSuccess) return e;
      if ((e = cudaStreamWaitEvent(h->stream_early, h->ev_fork, 0)) != cudaSuccess) return e;
      if ((e = go(a.width, 0u, 0, HDSM_ROW_OVERFLOW, t, 0, n_early, h->stream_early)) != cudaSuccess) return e;
      if ((e = cudaEventRecord(h->ev_join, h->stream_early)) != cudaSuccess) return e;
      e = go(csize, 0u, kFirstPassRounds, HDSM_ROW_OVERFLOW, t, n_early);
      if (e != cudaSuccess) return e;
      if ((e = cudaStreamWaitEvent(s, h->ev_join, 0)) != cudaSuccess) return e;
    } else {
      e = go(csize, t == 0 ? 0u : 1u << HDSM_ROW_OVERFLOW, two_pass ? kFirstPassRounds : 0, HDSM_ROW_OVERFLOW, t);
      if (e != cudaSuccess) return e;
    }
    if (two_pass) {
      e = go(a.width, (1u << kDeferred) | (t > 0 ? 1u << kDeferredOverflow : 0u), 0, last ? HDSM_ROW_OVERFLOW : kDeferredOverflow, t);
      if (e != cudaSuccess) return e;
    }
  }
  return cudaGetLastError();
}

cudaError_t dispatch(hdsm_handle* h, const KernelArgs& a, cudaStream_t s) {
  switch (h->prm.n_hor) {
#define HDSM_CASE(n) \
  case n:            \
    return h->warps == 1 ? launch<n, 1>(h, a, s) : launch<n, 4>(h, a, s);
    HDSM_CASE(5)
    HDSM_CASE(6)
    HDSM_CASE(7)
    HDSM_CASE(8)
    HDSM_CASE(9)
    HDSM_CASE(10)
    HDSM_CASE(11)
    HDSM_CASE(12)
#undef HDSM_CASE
    default:
      return cudaErrorInvalidValue;
  }
}

size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

// true when the copy engine can read / write `p` directly (cudaMallocHost / cudaHostRegister memory): the staging
// memcpy into the handle's own pinned arena is then skipped
bool is_pinned(const void* p) {
  if (!p) return false;
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}

// memcpy between caller memory and the pinned arenas; large blocks are split over a few threads (one
// core moves ~10 GB/s, the 140 MB of a 40 960-agent batch would otherwise cost more than its H2D copy)
void staged_copy(void* dst, const void* src, size_t bytes) {
  constexpr size_t kPerThread = size_t(4) << 20;
  const int nt = (int)std::min<size_t>(6, bytes / kPerThread);
  if (nt < 2) {
    std::memcpy(dst, src, bytes);
    return;
  }
  std::thread th[6];
  const size_t part = ((bytes / nt) + 63) & ~size_t(63);
  for (int t = 1; t < nt; ++t) {
    const size_t o = (size_t)t * part, len = t == nt - 1 ? bytes - o : part;
    th[t] = std::thread([=] { std::memcpy((char*)dst + o, (const char*)src + o, len); });
  }
  std::memcpy(dst, src, part);
  for (int t = 1; t < nt; ++t) th[t].join();
}

}  // namespace

extern "C" {

int hdsm_version(void) { return HDSM_VERSION; }

int hdsm_create(const hdsm_params* params, int max_agents, int max_neighbours, int device, hdsm_handle** out) {
  if (!params || !out || max_agents < 1 || max_neighbours < 0) return HDSM_ERR_INVALID;
  *out = nullptr;
  hdsm_handle* h = new (std::nothrow) hdsm_handle();
  if (!h) return HDSM_ERR_INVALID;
  h->prm = *params;
  if (h->prm.max_iter <= 0) h->prm.max_iter = 60;
  if (h->prm.max_nodes <= 0) h->prm.max_nodes = 64;
  if (!(h->prm.tol > 0)) h->prm.tol = 1e-8;
  if (h->prm.n_hor < HDSM_MIN_HOR || h->prm.n_hor > HDSM_MAX_HOR) {
    std::fprintf(stderr, "hdsm_create: n_hor = %d is outside the supported range %d..%d\n", h->prm.n_hor, HDSM_MIN_HOR, HDSM_MAX_HOR);
    delete h;
    return HDSM_ERR_INVALID;
  }
  if (int rc = build_tables(h->prm, h->host_tables)) {
    std::fprintf(stderr, "hdsm_create: unsupported parameter set (build_tables code %d)\n", rc);
    delete h;
    return HDSM_ERR_INVALID;
  }
  h->device = device, h->max_agents = max_agents, h->max_neighbours = max_neighbours;
  if (const char* e = std::getenv("HDSM_WARPS")) h->warps = std::atoi(e) == 1 ? 1 : 4;
  if (const char* e = std::getenv("HDSM_NO_ORDER")) h->use_order = std::atoi(e) == 0;
  if (const char* e = std::getenv("HDSM_FIRST_ROUNDS")) h->first_rounds = std::max(1, std::atoi(e));
  if (const char* e = std::getenv("HDSM_SINGLE_PASS")) h->single_pass = std::atoi(e) != 0;
  if (const char* e = std::getenv("HDSM_CLUSTER")) h->force_csize = std::max(0, std::min(std::atoi(e), kMaxWidth));
  if (h->prm.search_width != 0 && h->prm.search_width != 1 && h->prm.search_width != 2 && h->prm.search_width != 4 &&
      h->prm.search_width != 8) {
    std::fprintf(stderr, "hdsm_create: search_width must be 0, 1, 2, 4 or 8\n");
    delete h;
    return HDSM_ERR_INVALID;
  }
  if (h->prm.search_width == 0) h->prm.search_width = 1;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || device < 0 || device >= ndev) {  // no CPU fallback: fail loudly
    delete h;
    return HDSM_ERR_CUDA;
  }
  auto bail = [&](cudaError_t err, const char* what) {
    std::fprintf(stderr, "hdsm_create: %s: %s\n", what, cudaGetErrorString(err));
    hdsm_destroy(h);
    return HDSM_ERR_CUDA;
  };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(e, "cudaSetDevice");
  {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) h->block_slots = sms * HDSM_MINBLOCKS;
  }
  if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "stream");
  if ((e = cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "stream2");
  if ((e = cudaEventCreateWithFlags(&h->ev_shared, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "event");
  {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if ((e = cudaStreamCreateWithPriority(&h->stream_early, cudaStreamNonBlocking, hi)) != cudaSuccess) return bail(e, "stream_early");
    if ((e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "event");
    if ((e = cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "event");
    if (const char* ev = std::getenv("HDSM_EARLY_DIV")) {  // 0: no head start; otherwise at least 2 (the first pass keeps half of the batch)
      const int v = std::atoi(ev);
      h->early_div = v <= 0 ? 0 : std::max(2, v);
    }
  }
  for (int c = 0; c < kMaxChunks; ++c)
    if ((e = cudaEventCreate(&h->ev_chunk[c])) != cudaSuccess || (e = cudaEventCreate(&h->ev_begin[c])) != cudaSuccess)
      return bail(e, "event");
  if ((e = cudaMalloc(&h->dev_tables, sizeof(Tables))) != cudaSuccess) return bail(e, "cudaMalloc tables");
  if ((e = cudaMemcpy(h->dev_tables, &h->host_tables, sizeof(Tables), cudaMemcpyHostToDevice)) != cudaSuccess)
    return bail(e, "copy tables");
  if ((e = cudaMalloc(&h->d_order, sizeof(int32_t) * (size_t)max_agents)) != cudaSuccess) return bail(e, "cudaMalloc order");
  {
    std::vector<int32_t> ident((size_t)max_agents);
    for (int i = 0; i < max_agents; ++i) ident[i] = i;
    if ((e = cudaMemcpy(h->d_order, ident.data(), ident.size() * sizeof(int32_t), cudaMemcpyHostToDevice)) != cudaSuccess)
      return bail(e, "init order");
  }
  // Shared-memory budget.  Worst case per agent: 2 planes per neighbour and variable position step plus two
  // polytopes' rows per step.  Exact pruning usually leaves a few dozen to a few hundred rows, so the first pass
  // runs with the largest row pool that still lets HDSM_MINBLOCKS blocks share an SM (the register allocation allows
  // no more than that anyway: a smaller pool would buy no occupancy, only a second launch for the crowded agents,
  // whose searches are the long ones); agents that overflow it are re-solved by a second launch with the worst-case
  // pool (or what one SM can hold).
  const int nkp = h->host_tables.nkp, rmax = h->prm.max_rows_per_poly, N = h->prm.n_hor, P = h->prm.poly_hor;
  const long worst = 2L * max_neighbours * nkp + 2L * rmax * nkp;
  int smem_max = 0, smem_sm = 0;
  cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device);
  const long fixed = (long)smem_doubles(N, P, rmax, 0) * 8;
  const long fit = std::max(16L, (long)(smem_max - fixed - 1024) / 48);
  // pool of HDSM_MINBLOCKS co-resident blocks inside the 228 KB carve-out.  (Round 1 stopped at 196 KB because the
  // parameter tables then lived in L1, and 28 KB of L1 cost 3 % on the 10-neighbour workload; the tables the
  // iterations walk are now copies in shared memory.  HDSM_CARVE_KB=196 restores the old split for A/B runs.)
  long carve_kb = 228;
  if (const char* e = std::getenv("HDSM_CARVE_KB")) carve_kb = std::max(64L, std::atol(e));
  const long carve = std::min<long>(smem_sm, carve_kb * 1024);
  const long shared4 = std::max(96L, (carve / HDSM_MINBLOCKS - 1024 - fixed) / 48 & ~7L);
  const long tiers[2] = {h->prm.prune ? shared4 : worst, worst};
  long prev = 0;
  for (long t : tiers) {
    const long cap = std::min(std::min(t, worst), fit);
    if (cap <= prev) continue;
    h->row_cap[h->n_tiers] = (int)cap;
    h->smem_bytes[h->n_tiers] = smem_doubles(N, P, rmax, (int)cap) * 8;
    ++h->n_tiers;
    prev = cap;
    if (cap >= std::min(worst, fit)) break;
  }
  *out = h;
  return HDSM_OK;
}

void hdsm_destroy(hdsm_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  hdsm_comm_destroy(h);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->stream2) cudaStreamSynchronize(h->stream2);
  if (h->stream_early) cudaStreamSynchronize(h->stream_early), cudaStreamDestroy(h->stream_early);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->ev_shared) cudaEventDestroy(h->ev_shared);
  for (int c = 0; c < kMaxChunks; ++c)
    if (h->ev_chunk[c]) cudaEventDestroy(h->ev_chunk[c]), cudaEventDestroy(h->ev_begin[c]);
  if (h->stream2) cudaStreamDestroy(h->stream2);
  cudaFree(h->dev_tables);
  cudaFree(h->d_order);
  cudaFree(h->d_bounds);
  cudaFree(h->d_prof);
  cudaFree(h->d_in);
  cudaFree(h->d_out);
  cudaFreeHost(h->h_in);
  cudaFreeHost(h->h_out);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

const char* hdsm_last_error(const hdsm_handle* h) { return h ? h->err.c_str() : "null handle"; }
int64_t hdsm_launch_count(const hdsm_handle* h) { return h ? h->launches : 0; }
int hdsm_smem_bytes(const hdsm_handle* h) { return h ? h->smem_bytes[0] : 0; }

// `slot` / `order_offset`: which pipeline chunk of the handle this call is (the public device entry point is
// slot 0); the dispatch order computed after the previous call of the same slot and size is used, then
// recomputed from this call's iteration counts.
// Plan spheres of the neighbour table for the solver's neighbour scan (hdsm_plan_bounds_kernel), enqueued on `s`;
// *out stays null for small tables, where the scan is cheap anyway.
static int build_bounds(hdsm_handle* h, int n_rob, const double* all_pos, const uint8_t* all_valid, cudaStream_t s, const double** out) {
  *out = nullptr;
  if (n_rob < 256 || !h->prm.prune || !all_pos || !all_valid) return HDSM_OK;
  if (n_rob > h->bounds_cap) {
    CU(cudaDeviceSynchronize());  // rare (first call / larger table): nothing may still read the old buffer
    cudaFree(h->d_bounds);
    h->d_bounds = nullptr, h->bounds_cap = 0;
    CU(cudaMalloc(&h->d_bounds, (size_t)n_rob * 32));
    h->bounds_cap = n_rob;
  }
  hdsm_plan_bounds_kernel<<<(n_rob + 127) / 128, 128, 0, s>>>(n_rob, h->prm.n_hor, all_pos, all_valid, h->d_bounds);
  h->launches += 1;
  CU(cudaGetLastError());
  *out = h->d_bounds;
  return HDSM_OK;
}

static int solve_device(hdsm_handle* h, int slot, size_t order_offset, const double* bounds, int n_local, const int32_t* global_id,
                        const int32_t* nbr_begin, const int32_t* nbr_end, const double* x0, const double* ref,
                        const double* poly_A, const double* poly_b, const int32_t* poly_rows, const double* prev_self_pos,
                        const double* all_pos, const uint8_t* all_valid, int n_rob, const int32_t* assign_in, double* traj,
                        double* ctrl, uint8_t* poly_used, int32_t* assign_out, hdsm_result* res, double* pos_out,
                        void* stream) {
  if (!h) return HDSM_ERR_INVALID;
  if (n_local == 0) return HDSM_OK;
  if (n_local < 0 || n_rob < 0 || !global_id || !x0 || !ref || !poly_A || !poly_b || !poly_rows || !prev_self_pos ||
      !traj || !ctrl || !poly_used || !assign_out || !res || (n_rob > 0 && (!all_pos || !all_valid)) ||
      ((nbr_begin == nullptr) != (nbr_end == nullptr)))
    return fail(h, HDSM_ERR_INVALID, "hdsm_solve_batch: null or inconsistent argument");
  if (n_local > h->max_agents || order_offset + n_local > (size_t)h->max_agents)
    return fail(h, HDSM_ERR_CAPACITY, "n_local exceeds max_agents of the handle");
  CU(cudaSetDevice(h->device));
  KernelArgs a{};
  a.n_local = n_local, a.n_rob = n_rob, a.rmax = h->prm.max_rows_per_poly, a.P = h->prm.poly_hor;
  a.global_id = global_id, a.nbr_begin = nbr_begin, a.nbr_end = nbr_end, a.poly_rows = poly_rows, a.assign_in = assign_in;
  a.x0 = x0, a.ref = ref, a.poly_A = poly_A, a.poly_b = poly_b, a.prev = prev_self_pos, a.all_pos = all_pos;
  a.all_valid = all_valid, a.traj = traj, a.ctrl = ctrl, a.pos_out = pos_out, a.poly_used = poly_used;
  a.assign_out = assign_out, a.res = res, a.prof = h->d_prof;
  a.max_iter = h->prm.max_iter, a.max_nodes = h->prm.max_nodes, a.prune = h->prm.prune, a.tol = h->prm.tol;
  a.width = h->prm.search_width, a.csize = 1, a.warm_start = h->prm.warm_start;
  if (const char* e = std::getenv("HDSM_DEBUG")) a.dbg = std::atoi(e);
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : h->stream;
  a.bounds = bounds;
  // longest-first dispatch once the blocks of the call do not all fit at once (below that the order cannot matter)
  const bool ordered = h->use_order && (n_local >= 1024 || (long)n_local * blocks_per_agent(h, n_local) > h->block_slots);
  a.order = ordered && h->order_n[slot] == n_local && h->order_off[slot] == order_offset ? h->d_order + order_offset : nullptr;
  CU(dispatch(h, a, s));
  if (ordered) {
    hdsm_order_kernel<<<1, 1024, 0, s>>>(res, n_local, h->d_order + order_offset);
    h->launches += 1;
    // this call's region of d_order now holds a permutation of [0, n_local): cached orders of other slots that
    // overlap it are stale
    for (int c = 0; c < kMaxChunks; ++c)
      if (c != slot && h->order_n[c] > 0 && h->order_off[c] < order_offset + (size_t)n_local &&
          order_offset < h->order_off[c] + (size_t)h->order_n[c])
        h->order_n[c] = 0;
    h->order_n[slot] = n_local, h->order_off[slot] = order_offset;
    CU(cudaGetLastError());
  }
  return HDSM_OK;
}

int hdsm_solve_batch_device(hdsm_handle* h, int n_local, const int32_t* global_id, const int32_t* nbr_begin,
                            const int32_t* nbr_end, const double* x0, const double* ref, const double* poly_A,
                            const double* poly_b, const int32_t* poly_rows, const double* prev_self_pos,
                            const double* all_pos, const uint8_t* all_valid, int n_rob, const int32_t* assign_in,
                            double* traj, double* ctrl, uint8_t* poly_used, int32_t* assign_out, hdsm_result* res,
                            double* pos_out, void* stream) {
  if (!h) return HDSM_ERR_INVALID;
  const double* bounds = nullptr;
  if (n_local > 0 && n_rob > 0) {
    CU(cudaSetDevice(h->device));
    if (int rc = build_bounds(h, n_rob, all_pos, all_valid, stream ? static_cast<cudaStream_t>(stream) : h->stream, &bounds)) return rc;
  }
  return solve_device(h, 0, 0, bounds, n_local, global_id, nbr_begin, nbr_end, x0, ref, poly_A, poly_b, poly_rows, prev_self_pos,
                      all_pos, all_valid, n_rob, assign_in, traj, ctrl, poly_used, assign_out, res, pos_out, stream);
}

int hdsm_solve_batch(hdsm_handle* h, int n_local, const int32_t* global_id, const int32_t* nbr_begin,
                     const int32_t* nbr_end, const double* x0, const double* ref, const double* poly_A,
                     const double* poly_b, const int32_t* poly_rows, const double* prev_self_pos,
                     const double* all_pos, const uint8_t* all_valid, int n_rob, const int32_t* assign_in, double* traj,
                     double* ctrl, uint8_t* poly_used, int32_t* assign_out, hdsm_result* res) {
  if (!h) return HDSM_ERR_INVALID;
  if (n_local == 0) return HDSM_OK;
  if (n_local < 0 || n_rob < 0 || !global_id || !x0 || !ref || !poly_A || !poly_b || !poly_rows || !prev_self_pos ||
      !traj || !ctrl || !poly_used || !assign_out || !res || (n_rob > 0 && (!all_pos || !all_valid)) ||
      ((nbr_begin == nullptr) != (nbr_end == nullptr)))
    return fail(h, HDSM_ERR_INVALID, "hdsm_solve_batch: null or inconsistent argument");
  if (n_local > h->max_agents) return fail(h, HDSM_ERR_CAPACITY, "n_local exceeds max_agents of the handle");
  CU(cudaSetDevice(h->device));
  const int N = h->prm.n_hor, P = h->prm.poly_hor, R = h->prm.max_rows_per_poly;
  const size_t n = n_local;
  // per-agent indices the kernel uses as array offsets (the device entry point clamps them instead)
  for (size_t i = 0; i < n; ++i) {
    if (n_rob > 0 && (global_id[i] < 0 || global_id[i] >= n_rob)) return fail(h, HDSM_ERR_INVALID, "global_id outside [0, n_rob)");
    if (nbr_begin && (nbr_begin[i] < 0 || nbr_begin[i] > nbr_end[i] || nbr_end[i] > n_rob))
      return fail(h, HDSM_ERR_INVALID, "neighbour range outside 0 <= nbr_begin <= nbr_end <= n_rob");
    for (int p = 0; p < P; ++p)
      if (poly_rows[i * P + p] < 0 || poly_rows[i * P + p] > R) return fail(h, HDSM_ERR_INVALID, "poly_rows outside [0, max_rows_per_poly]");
    if (assign_in)
      for (int k = 0; k < N; ++k)
        if (assign_in[i * N + k] < -1 || assign_in[i * N + k] >= P) return fail(h, HDSM_ERR_INVALID, "assign_in outside [-1, poly_hor)");
  }
  // ---- input arena: one pinned block, one H2D copy
  struct Seg {
    const void* src;
    size_t bytes, off;
  };
  Seg in[] = {
      {global_id, n * 4, 0},
      {nbr_begin, nbr_begin ? n * 4 : 0, 0},
      {nbr_end, nbr_end ? n * 4 : 0, 0},
      {poly_rows, n * P * 4, 0},
      {assign_in, assign_in ? n * N * 4 : 0, 0},
      {x0, n * 9 * 8, 0},
      {ref, n * N * 6 * 8, 0},
      {poly_A, n * P * R * 3 * 8, 0},
      {poly_b, n * P * R * 8, 0},
      {prev_self_pos, n * (N + 1) * 3 * 8, 0},
      {all_pos, (size_t)n_rob * (N + 1) * 3 * 8, 0},
      {all_valid, (size_t)n_rob, 0},
  };
  size_t in_bytes = 0;
  for (Seg& s : in) s.off = in_bytes, in_bytes += align256(s.bytes);
  const size_t o_traj = 0, o_ctrl = o_traj + align256(n * (N + 1) * 9 * 8), o_res = o_ctrl + align256(n * N * 3 * 8),
               o_asg = o_res + align256(n * sizeof(hdsm_result)), o_used = o_asg + align256(n * N * 4),
               out_bytes = o_used + align256(n * P);
  if (in_bytes > h->in_cap) {
    cudaFree(h->d_in), cudaFreeHost(h->h_in);
    h->d_in = h->h_in = nullptr, h->in_cap = 0;
    CU(cudaMalloc(&h->d_in, in_bytes));
    CU(cudaMallocHost(&h->h_in, in_bytes));
    h->in_cap = in_bytes;
  }
  if (out_bytes > h->out_cap) {
    cudaFree(h->d_out), cudaFreeHost(h->h_out);
    h->d_out = h->h_out = nullptr, h->out_cap = 0;
    CU(cudaMalloc(&h->d_out, out_bytes));
    CU(cudaMallocHost(&h->h_out, out_bytes));
    h->out_cap = out_bytes;
  }
  if (std::getenv("HDSM_PROFILE") && !h->d_prof) {
    CU(cudaMalloc(&h->d_prof, (size_t)h->max_agents * 16 * 8));
    CU(cudaMemset(h->d_prof, 0, (size_t)h->max_agents * 16 * 8));
  }
  // Chunked pipeline over two streams: while the GPU solves chunk c, the host stages chunk c+1 into the
  // pinned arena and its H2D copy runs on the copy engine; results stream back per chunk and are copied
  // out to the caller's buffers as soon as their chunk's event has fired.  The neighbour table is shared
  // by all chunks and goes first.  Small batches are one chunk.
  struct Out {
    void* dst;
    size_t off, stride;  // stride: bytes per agent
  };
  const Out outs[] = {{traj, o_traj, (size_t)(N + 1) * 9 * 8}, {ctrl, o_ctrl, (size_t)N * 3 * 8},
                      {res, o_res, sizeof(hdsm_result)},       {assign_out, o_asg, (size_t)N * 4},
                      {poly_used, o_used, (size_t)P}};
  const size_t in_stride[12] = {4, 4, 4, (size_t)P * 4, (size_t)N * 4, 72, (size_t)N * 48, (size_t)P * R * 24,
                                (size_t)P * R * 8, (size_t)(N + 1) * 24, 0, 0};  // per agent; 10, 11 are shared
  // Two chunks, not more: a few agents need 100x the median work (deep branch and bound), so every
  // kernel launch ends in a tail of nearly idle SMs; measured on 40 960 agents: 1 chunk 81 ms, 2 chunks
  // 72 ms, 4 chunks 74 ms, 10 chunks 102 ms per step.
  int n_chunks = n_local >= 8192 ? 2 : 1;
  if (const char* e = std::getenv("HDSM_CHUNKS")) n_chunks = std::max(1, std::min(std::atoi(e), kMaxChunks));
  if (h->d_prof) n_chunks = 1;
  const int per = (n_local + n_chunks - 1) / n_chunks;
  bool in_pinned[12], out_pinned[5];
  for (int i = 0; i < 12; ++i) in_pinned[i] = in[i].bytes && is_pinned(in[i].src);
  for (int i = 0; i < 5; ++i) out_pinned[i] = is_pinned(outs[i].dst);
  for (int i = 10; i < 12; ++i)
    if (in[i].bytes) {
      const void* src = in[i].src;
      if (!in_pinned[i]) {
        std::memcpy(h->h_in + in[i].off, in[i].src, in[i].bytes);
        src = h->h_in + in[i].off;
      }
      CU(cudaMemcpyAsync(h->d_in + in[i].off, src, in[i].bytes, cudaMemcpyHostToDevice, h->stream));
    }
  const double* bounds = nullptr;
  if (n_rob > 0)
    if (int rc = build_bounds(h, n_rob, (const double*)(h->d_in + in[10].off), (const uint8_t*)(h->d_in + in[11].off), h->stream, &bounds)) return rc;
  CU(cudaEventRecord(h->ev_shared, h->stream));
  CU(cudaStreamWaitEvent(h->stream2, h->ev_shared, 0));
  auto dp = [&](int i, size_t first) -> const void* { return in[i].bytes ? h->d_in + in[i].off + first * in_stride[i] : nullptr; };
  const bool trace = std::getenv("HDSM_TRACE") != nullptr;
  const auto t_begin = std::chrono::steady_clock::now();
  const auto ms_since = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
  double t_enq[kMaxChunks] = {}, t_done[kMaxChunks] = {};
  for (int c = 0; c < n_chunks; ++c) {
    const size_t first = (size_t)c * per, cnt = std::min<size_t>(per, n - first);
    cudaStream_t s = ((c & 1) && !std::getenv("HDSM_ONE_STREAM")) ? h->stream2 : h->stream;
    for (int i = 0; i < 10; ++i) {
      if (!in[i].bytes) continue;
      const size_t o = in[i].off + first * in_stride[i], bytes = cnt * in_stride[i];
      const char* src = (const char*)in[i].src + first * in_stride[i];
      if (!in_pinned[i]) {
        staged_copy(h->h_in + o, src, bytes);
        src = (const char*)h->h_in + o;
      }
      CU(cudaMemcpyAsync(h->d_in + o, src, bytes, cudaMemcpyHostToDevice, s));
    }
    if (trace) CU(cudaEventRecord(h->ev_begin[c], s));
    int rc = solve_device(
        h, c, first, bounds, (int)cnt, (const int32_t*)dp(0, first), (const int32_t*)dp(1, first), (const int32_t*)dp(2, first),
        (const double*)dp(5, first), (const double*)dp(6, first), (const double*)dp(7, first), (const double*)dp(8, first),
        (const int32_t*)dp(3, first), (const double*)dp(9, first), (const double*)dp(10, 0), (const uint8_t*)dp(11, 0), n_rob,
        (const int32_t*)dp(4, first), (double*)(h->d_out + o_traj + first * outs[0].stride),
        (double*)(h->d_out + o_ctrl + first * outs[1].stride), (uint8_t*)(h->d_out + o_used + first * outs[4].stride),
        (int32_t*)(h->d_out + o_asg + first * outs[3].stride), (hdsm_result*)(h->d_out + o_res + first * outs[2].stride),
        nullptr, s);
    if (rc != HDSM_OK) return rc;
    for (int i = 0; i < 5; ++i) {
      const Out& o = outs[i];
      void* dst = out_pinned[i] ? (void*)((char*)o.dst + first * o.stride) : (void*)(h->h_out + o.off + first * o.stride);
      CU(cudaMemcpyAsync(dst, h->d_out + o.off + first * o.stride, cnt * o.stride, cudaMemcpyDeviceToHost, s));
    }
    CU(cudaEventRecord(h->ev_chunk[c], s));
    t_enq[c] = ms_since();
  }
  for (int c = 0; c < n_chunks; ++c) {
    const size_t first = (size_t)c * per, cnt = std::min<size_t>(per, n - first);
    CU(cudaEventSynchronize(h->ev_chunk[c]));
    t_done[c] = ms_since();
    for (int i = 0; i < 5; ++i)
      if (!out_pinned[i]) {
        const Out& o = outs[i];
        staged_copy((char*)o.dst + first * o.stride, h->h_out + o.off + first * o.stride, cnt * o.stride);
      }
  }
  if (trace) {
    std::fprintf(stderr, "[hdsm trace] %d chunks, total %.2f ms\n", n_chunks, ms_since());
    for (int c = 0; c < n_chunks; ++c) {
      float k = 0;
      cudaEventElapsedTime(&k, h->ev_begin[c], h->ev_chunk[c]);
      std::fprintf(stderr, "  chunk %2d: enqueued at %7.2f ms, done seen at %7.2f ms, kernels+D2H %.2f ms\n", c, t_enq[c], t_done[c], k);
    }
  }
  if (h->d_prof) {
    std::vector<long long> hp((size_t)n_local * 16);
    CU(cudaMemcpy(hp.data(), h->d_prof, hp.size() * 8, cudaMemcpyDeviceToHost));
    double tot[16] = {0};
    for (int i = 0; i < n_local; ++i)
      for (int s = 0; s < 16; ++s) tot[s] += (double)hp[(size_t)i * 16 + s];
    const char* names[16] = {"dominance filter", "static rows", "qp init", "pass1", "residual/rhs", "kkt assembly", "factor",
                             "solves", "dense products", "other passes", "search", "load + index", "objective/boxes",
                             "neighbour scan", "neighbour rows", "root sets"};
#ifdef HDSM_PROF_PASS1
    names[11] = "p1 rows", names[12] = "p1 butterflies", names[13] = "p1 boxes", names[14] = "step-length passes", names[15] = "corrector rhs rows";
    names[3] = "p1 reduction", names[9] = "rhs product + update";
#endif
    double all = 0;
    for (double t : tot) all += t;
    std::fprintf(stderr, "[hdsm profile] warp-0 cycles per agent: %.0f\n", all / n_local);
    for (int s = 0; s < 16; ++s) std::fprintf(stderr, "  %-16s %6.1f%%  %10.0f cycles/agent\n", names[s], 100 * tot[s] / all, tot[s] / n_local);
  }
  return HDSM_OK;
}

int hdsm_comm_unique_id(uint8_t id_out[128]) {
  NcclApi* n = nccl_api();
  if (!n || !id_out) return HDSM_ERR_NCCL;
  Id128 id;
  if (n->GetUniqueId(&id) != 0) return HDSM_ERR_NCCL;
  std::memcpy(id_out, &id, 128);
  return HDSM_OK;
}

int hdsm_comm_init(hdsm_handle* h, int n_ranks, int rank, const uint8_t id[128]) {
  if (!h || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return HDSM_ERR_INVALID;
  NcclApi* n = nccl_api();
  if (!n) return fail(h, HDSM_ERR_NCCL, "libnccl.so.2 not found");
  CU(cudaSetDevice(h->device));
  Id128 uid;
  std::memcpy(&uid, id, 128);
  int rc = n->CommInitRank(&h->comm, n_ranks, uid, rank);
  if (rc != 0) return fail(h, HDSM_ERR_NCCL, std::string("ncclCommInitRank: ") + (n->GetErrorString ? n->GetErrorString(rc) : "?"));
  h->n_ranks = n_ranks;
  return HDSM_OK;
}

int hdsm_allgather_positions(hdsm_handle* h, const double* send, double* recv, int n_local, void* stream) {
  if (!h || !send || !recv || n_local < 0) return HDSM_ERR_INVALID;
  NcclApi* n = nccl_api();
  if (!n || !h->comm) return fail(h, HDSM_ERR_NCCL, "communicator not initialised");
  const size_t count = (size_t)n_local * (h->prm.n_hor + 1) * 3;
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : h->stream;
  int rc = n->AllGather(send, recv, count, /*ncclFloat64*/ 8, h->comm, s);
  if (rc != 0) return fail(h, HDSM_ERR_NCCL, std::string("ncclAllGather: ") + (n->GetErrorString ? n->GetErrorString(rc) : "?"));
  h->launches += 1;
  return HDSM_OK;
}

int hdsm_exchange_plans(hdsm_handle* h, const double* send_pos, double* recv_pos, const uint8_t* send_valid, uint8_t* recv_valid,
                        int n_local, void* stream) {
  if (!h || !send_pos || !recv_pos || !send_valid || !recv_valid || n_local < 0) return HDSM_ERR_INVALID;
  NcclApi* n = nccl_api();
  if (!n || !h->comm) return fail(h, HDSM_ERR_NCCL, "communicator not initialised");
  const size_t count = (size_t)n_local * (h->prm.n_hor + 1) * 3;
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : h->stream;
  // both gathers in one NCCL group: one launch, one synchronisation of the ranks per replanning step
  if (n->GroupStart && n->GroupEnd) n->GroupStart();
  int rc = n->AllGather(send_pos, recv_pos, count, /*ncclFloat64*/ 8, h->comm, s);
  int rc2 = rc == 0 ? n->AllGather(send_valid, recv_valid, (size_t)n_local, /*ncclUint8*/ 1, h->comm, s) : 0;
  int rc3 = (n->GroupStart && n->GroupEnd) ? n->GroupEnd() : 0;
  rc = rc ? rc : (rc2 ? rc2 : rc3);
  if (rc != 0) return fail(h, HDSM_ERR_NCCL, std::string("ncclAllGather: ") + (n->GetErrorString ? n->GetErrorString(rc) : "?"));
  h->launches += 1;
  return HDSM_OK;
}

int hdsm_advance_device(hdsm_handle* h, int n_local, const double* traj, const double* ctrl, const hdsm_result* res,
                        double* traj_curr, double* ctrl_curr, uint8_t* have_plan, double* x0, double* prev_self_pos, void* stream) {
  if (!h) return HDSM_ERR_INVALID;
  if (n_local == 0) return HDSM_OK;
  if (n_local < 0 || !traj || !ctrl || !res || !traj_curr || !ctrl_curr || !have_plan || !x0)
    return fail(h, HDSM_ERR_INVALID, "hdsm_advance_device: null argument");
  CU(cudaSetDevice(h->device));
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : h->stream;
  hdsm_advance_kernel<<<(n_local + 7) / 8, 256, 0, s>>>(n_local, h->prm.n_hor, traj, ctrl, res, traj_curr, ctrl_curr, have_plan, x0,
                                                        prev_self_pos);
  h->launches += 1;
  CU(cudaGetLastError());
  return HDSM_OK;
}

int hdsm_planes(hdsm_handle* h, int n, const double* own_pos, const double* other_pos, double* planes) {
  if (!h || n < 0 || (n > 0 && (!own_pos || !other_pos || !planes))) return HDSM_ERR_INVALID;
  if (n == 0) return HDSM_OK;
  CU(cudaSetDevice(h->device));
  double *d_in = nullptr, *d_out = nullptr;
  CU(cudaMalloc(&d_in, (size_t)n * 6 * 8));
  cudaError_t e = cudaMalloc(&d_out, (size_t)n * 4 * 8);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_in, own_pos, (size_t)n * 24, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_in + (size_t)n * 3, other_pos, (size_t)n * 24, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) {
    hdsm_planes_kernel<<<(n + 127) / 128, 128, 0, h->stream>>>(h->prm, n, d_in, d_in + (size_t)n * 3, d_out);
    h->launches += 1;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(planes, d_out, (size_t)n * 32, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d_in), cudaFree(d_out);
  if (e != cudaSuccess) return cuda_fail(h, e, "hdsm_planes");
  return HDSM_OK;
}

void hdsm_comm_destroy(hdsm_handle* h) {
  if (!h || !h->comm) return;
  NcclApi* n = nccl_api();
  if (n) n->CommDestroy(h->comm);
  h->comm = nullptr;
}

}  // extern "C"
