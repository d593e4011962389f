/* hdsm.h - C ABI of the B200 (sm_100a) batched trajectory-optimisation library.
 *
 * Drop-in boundary for ONE hot path of lis-epfl/multi_agent_pkgs: the per-agent receding-horizon
 * trajectory optimisation of multi_agent_planner.  The reference has no plugin / FFI interface for
 * it - the solver is three private member functions of class Agent plus private Gurobi members
 * (multi_agent_planner/include/agent_class.hpp:63-70, 100-104, 407-432) - so the seam is defined
 * here, one entry point per reference call site:
 *
 *   hdsm_create              replaces  Agent::InitializeGurobi + Agent::CreateGurobiModel +
 *                                      Agent::InitializePlannerParameters
 *                                      (agent_class.cpp:2063-2069, :2071-2153, :2169-2188)
 *   hdsm_solve_batch         replaces  Agent::GenerateTimeAwareSafeCorridor (:1086-1215) followed by
 *                                      Agent::SolveOptimizationProblem (:858-1023), i.e. the two
 *                                      calls at agent_class.cpp:168 and :174, incl. GRBModel::optimize (:959)
 *   hdsm_solve_batch_device  same, device pointers, stream ordered (synthetic-swarm harness)
 *   hdsm_comm_* / hdsm_allgather_positions
 *                            replaces  the ROS2 broadcast of Trajectory messages between agents
 *                                      (publisher :645-677, subscriber :629-643) inside the harness
 *   hdsm_destroy             replaces  ~GRBModel / ~GRBEnv (members at agent_class.hpp:409-410)
 *
 * Conventions: plain pointers and sizes, row-major contiguous arrays, IEEE double like the
 * reference (decomp_basis/data_type.h:50).  No exceptions cross the ABI: every function returns
 * HDSM_OK or a negative error code (hdsm_last_error gives text); per-agent solver outcomes are in
 * hdsm_result.status and the caller keeps the reference's fallback (:997-1019) for anything
 * that is not HDSM_OPTIMAL / HDSM_NODE_LIMIT-with-incumbent.  The library keeps no pointer after
 * a call returns.  One handle per calling thread; calls on one handle are not re-entrant.
 * There is no CPU fallback: without a CUDA device hdsm_create fails.
 */
#ifndef HDSM_H_
#define HDSM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HDSM_VERSION 100

/* library return codes */
#define HDSM_OK 0
#define HDSM_ERR_INVALID (-1)   /* bad argument / unsupported parameter value */
#define HDSM_ERR_CUDA (-2)      /* CUDA runtime error, see hdsm_last_error */
#define HDSM_ERR_CAPACITY (-3)  /* n_local / n_rob exceed what the handle was created for */
#define HDSM_ERR_NCCL (-4)      /* NCCL missing or failed */

/* per-agent solver status (hdsm_result.status) */
#define HDSM_OPTIMAL 0       /* optimum of the mixed-integer problem proven */
#define HDSM_INFEASIBLE 1    /* no assignment of polytopes admits a trajectory (GRB infeasible -> :988) */
#define HDSM_MAX_ITER 2      /* interior point hit max_iter without converging */
#define HDSM_NUMERICAL 3     /* factorisation broke down / non-finite iterate / degenerate plane */
#define HDSM_NODE_LIMIT 4    /* search stopped at max_nodes; traj holds the incumbent if obj is finite
                                (the reference accepts a TimeLimit incumbent the same way, :952-987) */
#define HDSM_ROW_OVERFLOW 5  /* more active constraint rows than the handle's shared-memory budget */

#define HDSM_MIN_HOR 5   /* kernels are instantiated for horizons 5..12 */
#define HDSM_MAX_HOR 12
#define HDSM_MAX_POLY 8

/* POD mirror of the ROS parameters that enter the optimisation (agent_class.cpp:2190-2248;
 * values e.g. multi_agent_planner/config/agent_agile_config.yaml). */
typedef struct hdsm_params {
  int32_t n_hor;             /* n_hor: horizon N, HDSM_MIN_HOR..HDSM_MAX_HOR (n_x = 9, n_u = 3 fixed: ModelODE :2155) */
  int32_t poly_hor;          /* poly_hor: polytopes per step P, 1..HDSM_MAX_POLY */
  int32_t max_rows_per_poly; /* row stride Rmax of poly_A / poly_b, <= 32 and poly_hor * Rmax <= 255 */
  int32_t rk4;               /* rk4: 0 = Euler (:2140-2151), 1 = RK4 (:2123-2139) */
  int32_t max_iter;          /* interior-point iterations per QP (0 -> 60) */
  int32_t max_nodes;         /* branch-and-bound nodes per agent (0 -> 64); plays the role of TimeLimit */
  int32_t prune;             /* 1 = drop rows that can never be active inside the reachable box (exact) */
  int32_t search_width;      /* open nodes the assignment search solves per round: 0 / 1 = depth-first search, 2, 4 or 8 =
                                rounds whose nodes are independent and run on a thread-block cluster when the batch is
                                small (lower latency of hard agents, some speculative work).  Results are a function of
                                this value only - not of the batch size, the GPU or the number of GPUs. */
  double dt;                 /* dt */
  double drag[3];            /* drag_coeff */
  double r_u;                /* r_u  (:2098) */
  double r_x[6];             /* r_x[0..5] (:878-881; acceleration weights are never applied) */
  double r_n[6];             /* r_n[0..5] (:873-876) */
  double max_vel;            /* :2181-2184 */
  double min_acc_xy, max_acc_xy, min_acc_z, max_acc_z;
  double max_jerk;           /* :2185-2186 */
  double drone_radius;       /* :1161-1163 */
  double drone_z_offset;
  double tilt;               /* var_tmp = 0.1 (:1180) */
  double tol;                /* KKT tolerance of a QP solve (0 -> 1e-8) */
  int32_t warm_start;        /* 1 = warm start of the assignment search (the reference keeps the polytopes the
                                last plan used for the same purpose, agent_class.cpp:1269-1282): once the root relaxation has
                                branched, the previous plan shifted by one step (prev_self_pos) names a cell per step, and that
                                assignment is solved in the next round, so that an incumbent exists early.  Changes which nodes
                                are explored, never the optimum */
  int32_t reserved;
} hdsm_params;

typedef struct hdsm_result {
  int32_t status; /* HDSM_OPTIMAL ... */
  int32_t iters;  /* interior-point iterations summed over nodes */
  int32_t nodes;  /* QP relaxations solved */
  int32_t rows;   /* largest number of position rows handed to one QP (after pruning) */
  double obj;     /* objective incl. the constant ref^2 terms, comparable with Gurobi ObjVal */
  double kkt_res; /* max scaled KKT residual of the returned solution */
} hdsm_result;

typedef struct hdsm_handle hdsm_handle;

int hdsm_version(void);

/* max_agents: largest n_local of a later call.  max_neighbours: largest number of neighbour
 * candidates (nbr_end - nbr_begin, or n_rob) of a later call; sizes the per-block row buffers. */
int hdsm_create(const hdsm_params* params, int max_agents, int max_neighbours, int device, hdsm_handle** out);
void hdsm_destroy(hdsm_handle* h);
const char* hdsm_last_error(const hdsm_handle* h);

/* One replanning step for n_local agents; HOST pointers; synchronous.
 *
 *  global_id     [n_local]            index of each agent in all_pos (own slot is skipped like the
 *                                     empty own entry of traj_other_agents_, :1132-1134)
 *  nbr_begin/end [n_local] or NULL    neighbour candidates = all_pos[nbr_begin, nbr_end) (NULL: 0..n_rob)
 *  x0            [n_local][9]         state_curr_ (:886-889)  (px py pz vx vy vz ax ay az)
 *  ref           [n_local][N][6]      traj_ref_curr_[0..N-1], pos+vel (:871-883)
 *  poly_A        [n_local][P][Rmax][3]  poly_const_vec_: A x <= b rows (polyhedron.h:98-147)
 *  poly_b        [n_local][P][Rmax]
 *  poly_rows     [n_local][P]         rows per polytope, 0 = absent (P_eff = leading non-zero count, :913)
 *  prev_self_pos [n_local][N+1][3]    own previous plan positions traj_curr_[k][0..2] (:1103-1110);
 *                                     state_ini_ repeated before the first solve
 *  all_pos       [n_rob][N+1][3]      last received plans of all agents (states[k].position, :1147-1149)
 *  all_valid     [n_rob]              plan received? (traj.states.size() != 0, :1134)
 *  assign_in     [n_local][N] or NULL per-step polytope index to force, -1 = search
 *  traj          [n_local][N+1][9]    traj_curr_ (:962-973)
 *  ctrl          [n_local][N][3]      control_curr_ (:975-978)
 *  poly_used     [n_local][P]         poly_used_idx_ (:981-985)
 *  assign_out    [n_local][N]         polytope chosen per step (the binaries b[k][p])
 *  res           [n_local]
 */
int hdsm_solve_batch(hdsm_handle* h, int n_local, const int32_t* global_id, const int32_t* nbr_begin,
                     const int32_t* nbr_end, const double* x0, const double* ref, const double* poly_A,
                     const double* poly_b, const int32_t* poly_rows, const double* prev_self_pos,
                     const double* all_pos, const uint8_t* all_valid, int n_rob, const int32_t* assign_in,
                     double* traj, double* ctrl, uint8_t* poly_used, int32_t* assign_out, hdsm_result* res);

/* Same with DEVICE pointers, enqueued on `stream` (a cudaStream_t; NULL = the handle's stream), no
 * host synchronisation.  pos_out [n_local][N+1][3] (may be NULL) receives the packed positions of
 * the new plans - the send buffer of the trajectory exchange, written by the solver's epilogue;
 * agents without a usable result get their previous plan shifted by one step (:1004-1013).
 * Calls on one handle must be ordered with respect to each other (same stream, or synchronised): the handle
 * keeps a device-side dispatch order (most expensive agents of the previous call first) between calls. */
int hdsm_solve_batch_device(hdsm_handle* h, int n_local, const int32_t* global_id, const int32_t* nbr_begin,
                            const int32_t* nbr_end, const double* x0, const double* ref, const double* poly_A,
                            const double* poly_b, const int32_t* poly_rows, const double* prev_self_pos,
                            const double* all_pos, const uint8_t* all_valid, int n_rob,
                            const int32_t* assign_in, double* traj, double* ctrl, uint8_t* poly_used,
                            int32_t* assign_out, hdsm_result* res, double* pos_out, void* stream);

/* K1 alone, HOST pointers, synchronous: the inter-agent plane of n point pairs exactly as the solver builds it
 * (agent_class.cpp:1152-1205 with pert = 0): own_pos / other_pos [n][3] -> planes [n][4] = (n_f, n_f . pt); NaN for
 * coincident points.  Lets a caller (and the parity tests) look at the rows GenerateTimeAwareSafeCorridor appends
 * to every polytope without solving anything. */
int hdsm_planes(hdsm_handle* h, int n, const double* own_pos, const double* other_pos, double* planes);

/* Number of kernels the library launched on this handle so far (bench.py's gpu_launches). */
int64_t hdsm_launch_count(const hdsm_handle* h);
/* Dynamic shared memory per thread block of the solver kernel, bytes. */
int hdsm_smem_bytes(const hdsm_handle* h);

/* ---- trajectory exchange between the GPUs of one box (one process per GPU) ------------------
 * hdsm_comm_unique_id fills a 128-byte NCCL id on rank 0; the caller broadcasts it by any means;
 * every rank then calls hdsm_comm_init.  hdsm_allgather_positions is one ncclAllGather of
 * n_local*(N+1)*3 doubles per rank on `stream` (device pointers; recv holds n_ranks such blocks). */
int hdsm_comm_unique_id(uint8_t id_out[128]);
int hdsm_comm_init(hdsm_handle* h, int n_ranks, int rank, const uint8_t id[128]);
int hdsm_allgather_positions(hdsm_handle* h, const double* send, double* recv, int n_local, void* stream);
void hdsm_comm_destroy(hdsm_handle* h);
/* The complete per-step exchange of the harness: plan positions as above plus the "plan received" flags
 * (traj.states.size() != 0 of traj_other_agents_, agent_class.cpp:1134) - send_valid [n_local] -> recv_valid
 * [n_ranks * n_local] - as one NCCL group (one launch). */
int hdsm_exchange_plans(hdsm_handle* h, const double* send_pos, double* recv_pos, const uint8_t* send_valid,
                        uint8_t* recv_valid, int n_local, void* stream);

/* Read-back and failure fallback of one replanning step for callers that keep the swarm state in HBM (DEVICE
 * pointers, enqueued on `stream`): what the reference does after optimize() with the solver's outputs
 * (agent_class.cpp:962-987 read-back; :997-1019 on failure the previous plan is shifted by one step and its last
 * element duplicated; :180-190 without a previous plan nothing is published) followed by the state advance
 * state_curr_ := traj_curr_[step_plan_] with step_plan_ = 1 (:233-238).
 *
 *  traj / ctrl / res   this step's outputs of hdsm_solve_batch_device
 *  traj_curr [n][N+1][9], ctrl_curr [n][N][3]   traj_curr_ / control_curr_, updated in place
 *  have_plan [n]       in: a plan exists (traj_curr_.size() > 0); out: same after this step
 *  x0 [n][9]           state_curr_, advanced in place for agents with a plan
 *  prev_self_pos [n][N+1][3] (may be NULL)  positions of traj_curr_ - the next call's prev_self_pos and this
 *                      step's all-gather payload; state_curr_ repeated while there is no plan (:1103-1110) */
int hdsm_advance_device(hdsm_handle* h, int n_local, const double* traj, const double* ctrl, const hdsm_result* res,
                        double* traj_curr, double* ctrl_curr, uint8_t* have_plan, double* x0, double* prev_self_pos,
                        void* stream);

/* ---- safe-corridor generation ----------------------------------
 * hdsm_corridor_batch replaces Agent::GenerateSafeCorridor (agent_class.cpp:1236-1447) together with
 * convex_decomp_lib::GetPolyOcta3D and GetPolyOcta3DNew (convex_decomp_util/src/convex_decomp.cpp:5-376,
 * :590-1162 with FindCorners :378-561), the call at
 * agent_class.cpp:165, for a batch of agents; its outputs are exactly hdsm_solve_batch's poly_A / poly_b
 * / poly_rows inputs (the conversion at :1428-1437).  One warp per agent; see csrc/hdsm_corridor.cu. */
#define HDSM_COR_SQUEEZED 1      /* informational: a seed voxel was squeezed between occupied voxels and its polytope
                                    was grown with GetPolyOcta3DNew, as the reference does (:1385-1395) */
#define HDSM_COR_ROW_OVERFLOW 2  /* a polytope had more rows than max_rows_per_poly: generation stopped */
#define HDSM_COR_SEED_OUTSIDE 4  /* a seed fell outside the voxel grid: generation stopped */
#define HDSM_COR_LIST_OVERFLOW 8 /* internal cell list overflow (cannot happen for n_it_decomp <= 90) */

typedef struct hdsm_corridor_params {
  int32_t poly_hor;          /* poly_hor: polytopes per agent P, 1..HDSM_MAX_POLY (:1302) */
  int32_t n_it_decomp;       /* n_it_decomp: face-growth iterations of GetPolyOcta3D, <= 90 */
  int32_t max_rows_per_poly; /* row stride Rmax of the polytope arrays, 18..32 (<= 12 chamfers + 6 faces) */
  int32_t n_traj;            /* points of the previous plan traj_curr_ (N + 1); 0 = none */
  int32_t max_path;          /* row stride of `path` */
  int32_t use_cvx_new;       /* use_cvx_new: 1 = always GetPolyOcta3DNew (:1383); 0 = only for squeezed seeds (:1385-1395) */
  double voxel_size;         /* VoxelGrid::GetVoxSize() */
} hdsm_corridor_params;

typedef struct hdsm_corridor hdsm_corridor;

/* max_grids voxel grids of at most grid_stride voxels each (int8, x fastest, then y, then z, as
 * voxel_grid_util lays them out; 0 free, 100 occupied, -1 unknown, 1..99 potential field = free). */
int hdsm_corridor_create(const hdsm_corridor_params* params, int max_agents, int max_grids, size_t grid_stride,
                         int device, hdsm_corridor** out);
void hdsm_corridor_destroy(hdsm_corridor* h);
const char* hdsm_corridor_last_error(const hdsm_corridor* h);
int64_t hdsm_corridor_launch_count(const hdsm_corridor* h);
int hdsm_corridor_smem_bytes(const hdsm_corridor* h);

/* One corridor update for n agents; HOST pointers; synchronous.
 *
 *  grids       [n_grids][grid_stride]   voxel_grid_ (copied and OccupyUnknown'ed on the fly, :1292-1301)
 *  grid_index  [n] or NULL              grid of each agent (NULL: agent i uses grid i)
 *  dims        [n][3]                   GetDim()  (x, y, z)
 *  origins     [n][3]                   GetOrigin()
 *  pos         [n][3]                   state_curr_[0..2] (:1288)
 *  path        [n][max_path][3]         path_curr_ (:1284-1287), n_path[n] points each
 *  prev_n      [n] or NULL              poly_const_vec_.size() of the previous step (NULL: first step)
 *  prev_A/b/rows/seeds                  previous poly_const_vec_ / poly_seeds_, layouts as the outputs
 *  prev_used   [n][P]                   poly_used_idx_ of the last optimisation (hdsm_solve_batch's poly_used)
 *  prev_traj   [n][n_traj][3]           positions of traj_curr_ (:1252-1259)
 *  poly_A      [n][P][Rmax][3], poly_b [n][P][Rmax], poly_rows [n][P] (0 = absent), seeds [n][P][3]
 *  flags       [n]                      HDSM_COR_* bits
 * Outputs must not alias the prev_* inputs. */
int hdsm_corridor_batch(hdsm_corridor* h, int n, int n_grids, const int8_t* grids, const int32_t* grid_index,
                        const int32_t* dims, const double* origins, const double* pos, const double* path,
                        const int32_t* n_path, const int32_t* prev_n, const double* prev_A, const double* prev_b,
                        const int32_t* prev_rows, const double* prev_seeds, const uint8_t* prev_used,
                        const double* prev_traj, double* poly_A, double* poly_b, int32_t* poly_rows, double* seeds,
                        int32_t* flags);

/* Same with DEVICE pointers, enqueued on `stream` (NULL = the handle's stream), no host synchronisation. */
int hdsm_corridor_batch_device(hdsm_corridor* h, int n, const int8_t* grids, const int32_t* grid_index,
                               const int32_t* dims, const double* origins, const double* pos, const double* path,
                               const int32_t* n_path, const int32_t* prev_n, const double* prev_A,
                               const double* prev_b, const int32_t* prev_rows, const double* prev_seeds,
                               const uint8_t* prev_used, const double* prev_traj, double* poly_A, double* poly_b,
                               int32_t* poly_rows, double* seeds, int32_t* flags, void* stream);

/* ---- reference-trajectory generation ----------------------------
 * hdsm_reftraj_batch replaces Agent::GenerateReferenceTrajectory (agent_class.cpp:1449-1553, the call at
 * :171) with SamplePath (:1591-1663), KeepOnlyFreeReference (:1665-1693), ComputePathVelocity (:1695-1803),
 * GetVelocityLimit (:1805-1817) and the ray casts of path_finding_util::IsLineClear (path_tools.cpp:148-180,
 * voxel_grid_util/src/raycast.cpp:21-186).  One warp per agent; see csrc/hdsm_reftraj.cu. */
typedef struct hdsm_reftraj_params {
  int32_t n_hor;            /* n_hor: N; the reference trajectory has N + 1 points */
  int32_t max_path;         /* row stride of `path`, <= 32 */
  int32_t n_traj;           /* points per plan in traj / all_pos (N + 1), 0 = no neighbour sweep */
  int32_t reserved;
  double dt;                /* dt */
  double path_vel_min, path_vel_max, path_vel_dec; /* agent_agile_config.yaml:16-20 */
  double sens_dist, sens_pot, sens_other_agents;   /* :1793, :1813-1815 */
  double voxel_size;        /* VoxelGrid::GetVoxSize() */
} hdsm_reftraj_params;

typedef struct hdsm_reftraj hdsm_reftraj;

int hdsm_reftraj_create(const hdsm_reftraj_params* params, int max_agents, int max_grids, size_t grid_stride, int device,
                        hdsm_reftraj** out);
void hdsm_reftraj_destroy(hdsm_reftraj* h);
const char* hdsm_reftraj_last_error(const hdsm_reftraj* h);
int64_t hdsm_reftraj_launch_count(const hdsm_reftraj* h);

/* One reference-trajectory update for n agents; HOST pointers; synchronous.
 *
 *  grids / grid_index / dims / origins   voxel_grid_ as for hdsm_corridor_batch (potential-field values 1..99 matter here)
 *  path        [n][max_path][3]          path_curr_ (:1451-1454), n_path[n] >= 1 points each
 *  prev_ref    [n][N+1][3]               positions of traj_ref_curr_ of the previous step (:1459-1470)
 *  have_prev   [n]                       traj_ref_curr_.size() > 0 && !reset_path_
 *  increment   [n]                       increment_traj_ref_
 *  traj        [n][n_traj][3]            positions of traj_curr_ (:1757-1760)
 *  global_id / nbr_begin / nbr_end       as for hdsm_solve_batch
 *  all_pos     [n_rob][n_traj][3]        last received plans of all agents (traj_other_agents_, :1769-1776)
 *  all_valid   [n_rob]
 *  ref         [n][N+1][6]               traj_ref_curr_: positions and the velocity reference (:1528-1546)
 *  path_vel    [n]                       path_vel_
 */
int hdsm_reftraj_batch(hdsm_reftraj* h, int n, int n_grids, const int8_t* grids, const int32_t* grid_index,
                       const int32_t* dims, const double* origins, const double* path, const int32_t* n_path,
                       const double* prev_ref, const uint8_t* have_prev, const uint8_t* increment, const double* traj,
                       const int32_t* global_id, const int32_t* nbr_begin, const int32_t* nbr_end, const double* all_pos,
                       const uint8_t* all_valid, int n_rob, double* ref, double* path_vel);

/* Same with DEVICE pointers on `stream`.  ref_solver [n][N][6] (may be NULL) receives the first N points in
 * the layout of hdsm_solve_batch_device's `ref` input, so that the two chain without a copy. */
int hdsm_reftraj_batch_device(hdsm_reftraj* h, int n, const int8_t* grids, const int32_t* grid_index, const int32_t* dims,
                              const double* origins, const double* path, const int32_t* n_path, const double* prev_ref,
                              const uint8_t* have_prev, const uint8_t* increment, const double* traj,
                              const int32_t* global_id, const int32_t* nbr_begin, const int32_t* nbr_end,
                              const double* all_pos, const uint8_t* all_valid, int n_rob, double* ref, double* ref_solver,
                              double* path_vel, void* stream);

/* ---- local-map post-processing -----------------------------------
 * hdsm_map_batch replaces the grid post-processing every map update runs before the grid is published to the
 * planner (mapping_util/src/map_builder.cpp:207-216): MapBuilder::SetUncertainToUnknown (:331-365),
 * VoxelGrid::InflateObstacles and VoxelGrid::CreatePotentialField (voxel_grid_util/src/voxel_grid.cpp:251-298,
 * stencils from CreateMask :192-226).  One thread block per grid, the grid held twice in shared memory; see
 * csrc/hdsm_map.cu.  The crop / merge / ray-cast clearing of map_builder.cpp:80-205 is hdsm_sense_batch below. */
typedef struct hdsm_map_params {
  double voxel_size;      /* VoxelGrid::GetVoxSize() */
  double inflation_dist;  /* inflation_dist (mapping_util/config: 0.3) */
  double potential_dist;  /* potential_dist (1.5); 0 = no potential field */
  int32_t potential_pow;  /* potential_pow (4) */
  int32_t reserved;
} hdsm_map_params;

typedef struct hdsm_map hdsm_map;

/* grid_stride: voxels per grid slot; two copies must fit into shared memory (grid_stride <= 116 000). */
int hdsm_map_create(const hdsm_map_params* params, int max_grids, size_t grid_stride, int device, hdsm_map** out);
/* Host only, no device needed: the form of the potential stencil (CreateMask, voxel_grid.cpp:192-226) the kernel uses when
 * the stencil is a function of the distance alone - tab128[d2] = value at squared voxel distance d2, -128 = not in the
 * stencil.  Returns 1 if that form applies to these parameters (the kernel then computes the potential field as a
 * squared-distance transform), 0 if the kernel keeps the stencil walk, < 0 on invalid arguments. */
int hdsm_map_distance_table(const hdsm_map_params* params, int8_t* tab128);
void hdsm_map_destroy(hdsm_map* h);
const char* hdsm_map_last_error(const hdsm_map* h);
int64_t hdsm_map_launch_count(const hdsm_map* h);

/* grids_in / grids_out [n_grids][grid_stride] int8 (x fastest; 0 free, 100 occupied, -1 unknown), dims [n_grids][3].
 * HOST pointers, synchronous; the _device twin takes device pointers and a stream.  in and out may not alias. */
int hdsm_map_batch(hdsm_map* h, int n_grids, const int8_t* grids_in, const int32_t* dims, int8_t* grids_out);
int hdsm_map_batch_device(hdsm_map* h, int n_grids, const int8_t* grids_in, const int32_t* dims, int8_t* grids_out,
                          void* stream);

/* ---- local-map acquisition (the part in front of hdsm_map_batch) ---------
 * hdsm_sense_batch replaces what a map update does before the post-processing (mapping_util/src/map_builder.cpp):
 * the frame and crop of the environment grid around the agent (:89-153), MapBuilder::RaycastAndClear (:280-329,
 * ClearLine :367-432 on voxel_grid_util::Raycast) with the 360 degree or the limited field of view, the first
 * update's ClearVoxelsCenter (:160-167, :434-447) and MapBuilder::MergeVoxelGrids (:242-278).  Its output is
 * voxel_grid_curr_: the caller keeps it for the next update and passes it to hdsm_map_batch.  One thread block
 * per agent; see csrc/hdsm_sense.cu. */
typedef struct hdsm_sense_params {
  double voxel_size;   /* voxel_size of the environment grid (:88) */
  double range[3];     /* voxel_grid_range (metres); the local grid has floor(range / voxel_size) voxels per axis */
  int32_t free_grid;   /* free_grid: 1 = known map (crop only, unknown -> free), 0 = ray-cast clearing and merge */
  int32_t limited_fov; /* limited_fov */
  double fov_x, fov_y; /* radians (fov_x_, fov_y_, map_builder.cpp:73-74); used with limited_fov */
} hdsm_sense_params;

typedef struct hdsm_sense hdsm_sense;

/* dim = floor(range / voxel_size) per axis (:103-107): the shape of every grid this handle reads and writes */
int hdsm_sense_grid_dims(const hdsm_sense_params* params, int32_t dim[3]);
/* grid_stride: voxels per grid slot (>= dim[0] dim[1] dim[2]; the same value hdsm_map_create takes) */
int hdsm_sense_create(const hdsm_sense_params* params, int max_agents, size_t grid_stride, int device, hdsm_sense** out);
void hdsm_sense_destroy(hdsm_sense* h);
const char* hdsm_sense_last_error(const hdsm_sense* h);
int64_t hdsm_sense_launch_count(const hdsm_sense* h);

/* env [dim_env z][y][x] int8: the environment grid all n agents share (0 free, 100 occupied); origin_env its origin.
 * pos [n][3] agent positions (pos_curr_); rot [n][9] row-major rot_mat_cam_ (only read with limited_fov, else NULL).
 * old_grids [n][grid_stride] / old_origin [n][3] / have_old [n]: voxel_grid_curr_ of the previous update and its
 * origin; have_old[a] = 0 (or old_grids = NULL) is an agent's first update.  grids_out [n][grid_stride],
 * origin_out [n][3]: the new voxel_grid_curr_.  HOST pointers, synchronous.  The _device twin takes device pointers
 * and a stream for everything except dim_env and origin_env, which stay host pointers.  grids_out may not alias
 * old_grids. */
int hdsm_sense_batch(hdsm_sense* h, int n, const int8_t* env, const int32_t dim_env[3], const double origin_env[3], const double* pos,
                     const double* rot, const int8_t* old_grids, const double* old_origin, const uint8_t* have_old, int8_t* grids_out,
                     double* origin_out);
int hdsm_sense_batch_device(hdsm_sense* h, int n, const int8_t* env, const int32_t dim_env[3], const double origin_env[3], const double* pos,
                            const double* rot, const int8_t* old_grids, const double* old_origin, const uint8_t* have_old, int8_t* grids_out,
                            double* origin_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HDSM_H_ */
