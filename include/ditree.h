/* ditree.h -- C ABI of libditree.so: the B200 (sm_100a) tree-expansion hot path of DiTree.
 *
 * Drop-in boundary.  Every entry point replaces one reference function (cited per declaration as
 * file:line of the upstream repository).  The reference is pure Python, so the binding a
 * maintainer adds is a ctypes stub (INTEGRATION.md shows it for each call).
 *
 * Conventions
 *  - All array arguments are DEVICE pointers owned by the caller unless the name ends in _host.
 *  - Functions only enqueue work on `stream` (a cudaStream_t passed as void*) and return an int
 *    status: 0 = OK, <0 = DT_E_*.  They never throw.  dt_last_error(ctx) returns a message.
 *  - One dt_ctx per device owns the staged occupancy grid, packed denoiser weights and scratch.
 *    A ctx is re-entrant across ctxs but not thread-safe within one (like the reference, which
 *    is single-threaded and synchronous).
 *  - Strided arrays: element (candidate b, component d[, step s]) lives at
 *        base[b*cand_stride + s*step_stride + d*comp_stride]      (strides in elements)
 *    so the same kernels take the reference's array-of-structs layouts ((B,D), (B,S,A)) and the
 *    coalesced struct-of-arrays layouts ((D,B), (S,A,B)) the device pipeline uses.
 */
#ifndef DITREE_H
#define DITREE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dt_ctx dt_ctx;

enum {
  DT_OK = 0,
  DT_E_CUDA = -1,       /* a CUDA runtime/driver call failed (message has the CUDA error) */
  DT_E_ARG = -2,        /* invalid argument (NULL pointer, size out of range, ...) */
  DT_E_NOMAP = -3,      /* dt_set_map has not been called */
  DT_E_NOMODEL = -4,    /* dt_load_denoiser has not been called */
  DT_E_INDEX = -5,      /* the reference would raise IndexError (tall map, map_utils.py:326) */
  DT_E_UNSUPPORTED = -6 /* shape not supported by the sm_100a kernels */
};

enum { DT_PROP_STOP_ON_COLLISION = 1 }; /* flags of dt_propagate_collide */
enum { DT_F32 = 0, DT_BF16 = 1 };

/* ---- context ------------------------------------------------------------------------------ */
int dt_ctx_create(int device, dt_ctx** out);
void dt_ctx_destroy(dt_ctx* ctx);
const char* dt_last_error(dt_ctx* ctx);
const char* dt_version(void);

/* Tuning switches.  "splitk" (default 1): conv GEMMs with at most 128 rows (sampler batches of 1-2 candidates,
 * the reference's own B = 1 loop) and flat few-tile GEMMs with K >= 4096 (the map encoder's last stage at a few
 * hundred candidates) split K over all SMs and reduce in a second kernel -- ~2x less latency, but a
 * candidate's bits then depend on whether it was sampled alone or in a batch (different fp32 summation order,
 * same 2e-2 tolerance).  0 restores batch-size independent results.  Changing the value synchronises the
 * device and drops the captured sampler graphs (they bake in the kernel selection).
 * "pdl" (default 1; env DITREE_PDL=0): the denoiser's kernels are launched with programmatic stream serialization --
 * a kernel's CTAs are scheduled and run their prologue while the previous kernel drains, griddepcontrol.wait orders
 * the memory traffic.  "fork" (default 1; env DITREE_FORK=0): at sampler batches <= 128 the 1 x 1 residual convs of
 * the U-Net and the downsample branches of the encoder run on a side stream beside the block's first conv.  Neither
 * changes a single bit of the results (same kernels, same arithmetic, different overlap). */
int dt_set_option(dt_ctx* ctx, const char* name, int value);

/* Occupancy grid upload.  Replaces RRT_Planner.update_maze / BasePlanner.maze
 * (planners/RRT.py:57-59, planners/base_planner.py:117).  grid_host: rows*cols floats, row-major,
 * cell value 1 = wall.  s_global = metres per cell (1 car, 4 ant). Synchronous on `stream`. */
int dt_set_map(dt_ctx* ctx, const float* grid_host, int rows, int cols, float s_global, void* stream);

/* Additional grids next to the main one, for passes that hold several scenarios on different mazes (the
 * device-resident planner below; dt_local_map_slots).  slot in 0..31; same grid format as dt_set_map.  Synchronous. */
int dt_set_map_slot(dt_ctx* ctx, int slot, const float* grid_host, int rows, int cols, float s_global, void* stream);

/* ---- geometry ----------------------------------------------------------------------------- */
/* is_colliding_car (common/map_utils.py:103-115 -> is_colliding_parallel :221-329) for B states;
 * x/y/theta strided by `stride` elements.  flags_out[b] in {0,1}.  Bit-exact vs float64 NumPy
 * when fed the same float32 values. */
int dt_collide_car(dt_ctx* ctx, const float* x, const float* y, const float* theta, int64_t stride, int64_t B,
                   uint8_t* flags_out, void* stream);

/* is_colliding_parallel itself (common/map_utils.py:221-329) for N points with cell size `scale`
 * and ball radius r, INCLUDING the whole-batch early return with only the out-of-bounds mask
 * (:255-259).  Returns DT_E_INDEX through dt_sync_status when the reference would raise. */
int dt_collide_points(dt_ctx* ctx, const float* x, const float* y, int64_t stride, int64_t N, double scale, double r,
                      uint8_t* flags_out, void* stream);

/* is_colliding_ant (common/map_utils.py:126-136 -> is_colliding_maze :139-218); states: B rows of
 * >= 7 floats (x, y, z, q0..q3) with row stride `row_stride`; uses the ctx map scale. */
int dt_collide_ant(dt_ctx* ctx, const float* states, int64_t row_stride, int64_t B, double radius, uint8_t* flags_out,
                   void* stream);

/* create_local_map (common/map_utils.py:391-459): B robot-centric N x N crops.
 * out_dtype DT_F32: out is (B,N,N) float32 in {0,1} (the reference's array);
 * out_dtype DT_BF16: out is (B,N,N) bf16 holding 2*m-1 (the sampler's rescale, fm_policy.py:152),
 * the encoder's input format. */
int dt_local_map(dt_ctx* ctx, const float* x, const float* y, const float* theta, int64_t stride, int64_t B, int N,
                 double scale, int out_dtype, void* out, void* stream);

/* create_local_map for candidates that come in groups of `group_size` consecutive poses (a multiple of 8 that divides
 * B), group g cropping from the grid staged in map slot slot_of_group[g] (device array of B / group_size int32): the
 * per-group map index of a multi-scenario pass.  Output: (B,N,N) bf16 holding 2m-1 (the encoder's input). */
int dt_local_map_slots(dt_ctx* ctx, const float* x, const float* y, const float* theta, int64_t stride, int64_t B, int N,
                       double scale, const int32_t* slot_of_group, int group_size, void* out_bf16, void* stream);

/* check_obstacle_ahead (planners/RRT.py:61-81). */
int dt_ray_probe(dt_ctx* ctx, const float* x, const float* y, const float* theta, int64_t stride, int64_t B,
                 uint8_t* flags_out, void* stream);

/* check_no_obstacles_in_path (run_scenarios_with_lidar_DiTree.py:158-181): index of the first of
 * n path points (x,y strided) whose cell == 1 in the ctx map, else -1, written to idx_out[0]. */
int dt_path_first_obstacle(dt_ctx* ctx, const float* x, const float* y, int64_t stride, int64_t n, int32_t* idx_out,
                           void* stream);

/* The same test against a caller-supplied grid: grid_u8 is a DEVICE array of rows*cols bytes, row-major, compared
 * with == 1 (the online driver's scanned map holds 0 = unknown, 1 = obstacle, 2 = seen free,
 * run_scenarios_with_lidar_DiTree.py:120-122).  The ctx map is neither needed nor touched.  Map centre and cell
 * size are the car's (cols/2, rows/2, 1 m). */
int dt_path_first_obstacle_grid(dt_ctx* ctx, const uint8_t* grid_u8, int rows, int cols, const float* x, const float* y,
                                int64_t stride, int64_t n, int32_t* idx_out, void* stream);

/* Lidar2DSim.scan (lidar_sim/lidar_2d_sim.py:18-98, noise_std = 0) for B poses in GRID coordinates
 * (x = col, y = row, yaw): 181 rays each.  dist_out (B,181) f64, end_out (B,181,2) f64,
 * visited_out nullable (B, rows*cols) u8 mask of cells the rays crossed before their hit. */
int dt_lidar_scan(dt_ctx* ctx, const float* pose, int64_t B, double* dist_out, double* end_out, uint8_t* visited_out,
                  void* stream);

/* ---- dynamics ----------------------------------------------------------------------------- */
/* Fused BasePlanner.propagate_action_sequence_env (planners/base_planner.py:257-320) over
 * CarEnv.step/_update_state/_check_done (car_env.py:240-282,341-396) + is_colliding_car, one
 * thread per candidate, all S Euler steps in registers.
 *   state0 / state_out : 6 components, strided (s_cand, s_comp)
 *   actions            : strided (a_cand, a_step, a_comp), 2 components
 *   traj_out (nullable): strided (t_cand, t_step, t_comp), 6 components; rows after the edge ends
 *                        are zero (the reference leaves them zero and strips them, RRT.py:196-199)
 *   first_coll[b] = first step whose state collides, -1 none;  done_step[b] = step at which the goal
 *   disc (0.5 m, car_env.py:349) was entered, -1 none.  With DT_PROP_STOP_ON_COLLISION the edge ends
 *   at the first collision (reference behaviour: the planner drops such edges, RRT.py:179-184). */
int dt_propagate_collide(dt_ctx* ctx, const float* state0, int64_t s_cand, int64_t s_comp, const float* actions,
                         int64_t a_cand, int64_t a_step, int64_t a_comp, int64_t B, int S, float goal_x, float goal_y,
                         float* traj_out, int64_t t_cand, int64_t t_step, int64_t t_comp, float* state_out,
                         int32_t* first_coll, int32_t* done_step, int flags, void* stream);

/* ---- sampler conditioning (policies/fm_policy.py:53-143) ------------------------------------ */
/* car: cond (B,7) f32 = [v_n, D_n, delta_n, a0_n, a1_n, tanh(R(-yaw)(goal-p)/map_size)].
 * state strided (s_cand, s_comp); prev_action (B,2) rows or NULL (zeros, un-normalised);
 * goal: goal_stride = 0 -> one (2,) goal for all, 2 -> (B,2).  norm: 16 doubles on the HOST:
 * obs_mean[6], obs_std[6], act_mean[2], act_std[2]. */
int dt_build_cond_car(dt_ctx* ctx, const float* state, int64_t s_cand, int64_t s_comp, const float* prev_action,
                      const float* goal, int goal_stride, int64_t B, const double* norm_host, double map_size,
                      float* cond_out, void* stream);

/* ant: obs_seq (B,h,29) f32 rows (x, y, 27 dims), h <= obs_history; cond (B, obs_history*29+8+2).
 * norm_host: obs_mean[27], obs_std[27], act_mean[8], act_std[8]. */
int dt_build_cond_ant(dt_ctx* ctx, const float* obs_seq, int h, int obs_history, const float* prev_action,
                      const float* goal, int goal_stride, int64_t B, const double* norm_host, double map_size,
                      float* cond_out, void* stream);

/* ---- reductions ----------------------------------------------------------------------------- */
/* RRT_Planner.nearest_node[_batch] (planners/RRT.py:49-55): 1-NN in (x,y) over n nodes for Q
 * queries, squared distance in float64, lowest index on ties. */
int dt_nearest(dt_ctx* ctx, const float* node_x, const float* node_y, int64_t n, const float* qx, const float* qy,
               int64_t q_stride, int64_t Q, int32_t* idx_out, void* stream);

/* k nearest nodes per query: kd_tree.query(sample, k) (planners/RRT.py:50 uses k = 1; SciPy's API for k > 1):
 * idx_out (Q, k) i32, ascending squared distance (float64), lowest index first on ties, n marks a missing
 * neighbour when k > n.  1 <= k <= 16. */
int dt_nearest_k(dt_ctx* ctx, const float* node_x, const float* node_y, int64_t n, const float* qx, const float* qy,
                 int64_t q_stride, int64_t Q, int k, int32_t* idx_out, void* stream);

/* Final node selection (planners/RRT.py:233-237): argmin_i ||p_i - goal|| + 1e4*ahead[i] (ahead
 * nullable), float64, first index on ties; idx_out[0]. */
int dt_goal_cost_argmin(dt_ctx* ctx, const float* node_x, const float* node_y, int64_t n, float goal_x, float goal_y,
                        const uint8_t* ahead, int32_t* idx_out, void* stream);

/* MPPI cost reduction (call sites run_scenarios_with_lidar_MPPI.py:339-341,422; PARITY UNPINNED:
 * the reference's MPPI module is not in its repository).  w = softmax(-(c - min c)/lambda);
 * u[t,a] += sum_k w_k noise[k,t,a]; argmin_out[0] = argmin c.  cost (K) f32, noise (K,TA) f32,
 * u_inout (TA) f32, weights_out nullable (K) f32. */
int dt_mppi_reduce(dt_ctx* ctx, const float* cost, const float* noise, int64_t K, int TA, float lambda, float* u_inout,
                   int32_t* argmin_out, float* weights_out, void* stream);

/* MPPI rollout cost for K rollouts of T steps from ONE start state (device, 6 floats) with controls u (T,2) + noise
 * (K,T,2): bicycle rollout with per-step collision / goal test fused with the cost
 *   cost[k] = ||p_T - target||^2 + collision_cost * collided + effort_cost * sum_t |u_t + noise_kt|^2,
 * target = ref_xy[min(argmin_j ||ref_j - p_0|| + lookahead, n_ref - 1)] (also written to target_out when not NULL).
 * PARITY UNPINNED like dt_mppi_reduce (the reference's MPPI module is absent); the call sites are
 * run_scenarios_with_lidar_MPPI.py:339-341,422. */
int dt_mppi_rollout_cost(dt_ctx* ctx, const float* state, const float* u, const float* noise, int64_t K, int T,
                         const float* ref_xy, int n_ref, int lookahead, float goal_x, float goal_y, float collision_cost,
                         float effort_cost, float* cost_out, float* target_out, void* stream);

/* End of an MPPI tick: action_out (A, nullable) = u[0]; u shifted left by one step, the last step repeated. */
int dt_mppi_shift(dt_ctx* ctx, float* u_inout, int T, int A, float* action_out, void* stream);

/* ---- probability-map state sampler (run_type >= 2) -------------------------------------------- */
/* CarEnv.prior (car_env.py:100-101, maze_map setter :117-121): exact Euclidean distance transform of the
 * free cells of the ctx map (scipy.ndimage.distance_transform_edt(1 - maze)) divided by its sum.
 * prior_out: rows*cols float64, device. */
int dt_edt_prior(dt_ctx* ctx, double* prior_out, void* stream);

/* CarEnv.prob_map of run_type >= 3 (car_env.py:106-110,124-128,130-137): gaussian_map(robot, goal, size =
 * (rows, cols)) (prob_sampling_utils.py:48-93) blended with the prior by combine_log_blend(prior, pdf, beta)
 * (:150-172, eps = 1e-12, its fallbacks included).  robot / goal are the (x, y) pairs the reference passes
 * (the pdf is zeroed at [int(robot_y), int(robot_x)]; DT_E_INDEX if that cell is outside the map, where
 * NumPy raises).  prior, prob_out, gauss_out: rows*cols float64, device; gauss_out receives CarEnv.gaussian_pdf. */
int dt_prob_map(dt_ctx* ctx, int rows, int cols, const double* prior, double robot_x, double robot_y, double goal_x,
                double goal_y, double beta, double* prob_out, double* gauss_out, void* stream);

/* combine_log_blend(prior, gauss, beta, obstacle_mask, eps) alone (prob_sampling_utils.py:150-172); obstacle_mask:
 * n bytes (non-zero = free) or NULL.  prior, gauss, prob_out: n float64, device. */
int dt_log_blend(dt_ctx* ctx, const double* prior, const double* gauss, const uint8_t* obstacle_mask, int n, double beta,
                 double eps, double* prob_out, void* stream);

/* BasePlanner.sample_row_col_from_probability_map (planners/base_planner.py:157-160) for B draws:
 * np.random.choice(n, p = prob) is cdf = cumsum(prob); cdf /= cdf[-1]; searchsorted(cdf, u, side='right') with
 * u ~ U[0,1) from the caller's generator; idx_out[b] is the flat cell index (bit-exact vs NumPy for the same u).
 * prob (n) f64, u (B) f64, idx_out (B) i32, all device. */
int dt_sample_cells(dt_ctx* ctx, const double* prob, int n, const double* u, int64_t B, int32_t* idx_out, void* stream);

/* ---- denoiser (local_map_encoder.py:78-122, conditional_unet1d.py:268-347, fm_policy.py:152-203) */
typedef struct {
  const char* name;   /* reference state_dict key, e.g. "unet.mid_modules.0.blocks.0.block.0.weight" */
  const float* data;  /* HOST pointer, float32, contiguous */
  int ndim;
  int64_t shape[4];
} dt_tensor_desc;

typedef struct {
  int action_dim;   /* A: 2 car, 8 ant */
  int horizon;      /* T = pred_horizon: 64 car, 16 ant */
  int cond_dim;     /* G: 7 car, 97 ant */
  int emb_dim;      /* local-map embedding: 400 */
  int map_size;     /* N: 20 car, 16 ant */
  int down_dims[3]; /* e.g. 512,1024,2048 */
  int max_batch;    /* scratch is sized for this many candidates */
} dt_model_cfg;

/* Pack a reference `noise_pred_net_state_dict` (run_scenarios.py:175-176) into bf16 GEMM operands. */
int dt_load_denoiser(dt_ctx* ctx, const dt_tensor_desc* tensors, int n_tensors, const dt_model_cfg* cfg, void* stream);

/* DiffusionSampler.forward's flow-matching loop (fm_policy.py:183-203): K Euler steps of the
 * denoiser from `noise`.  noise (B,T,A) f32, cond (B,G) f32, local_map (B,N,N) bf16 holding 2m-1
 * (dt_local_map DT_BF16).  actions_out (B,T,A) f32 = a*act_std + act_mean (norm_host: act_mean[A],
 * act_std[A]); pass norm_host = NULL for the normalised sample. */
int dt_fm_sample(dt_ctx* ctx, const float* noise, const float* cond, const void* local_map, int64_t B, int K,
                 double exp_scale, const double* norm_host, float* actions_out, void* stream);

/* Pieces of the denoiser, exported for the parity tests: encoder only -> emb_out (B,emb) f32;
 * one U-Net evaluation at `timestep` (already scaled by 20) -> vel_out (B,T,A) f32. */
int dt_encode_map(dt_ctx* ctx, const void* local_map, int64_t B, float* emb_out, void* stream);
int dt_unet_forward(dt_ctx* ctx, const float* sample, const float* emb, const float* cond, int64_t B, float timestep,
                    float* vel_out, void* stream);

/* ---- device-resident multi-scenario planner ---------------------------------------------------------------
 * RRT_Planner.plan (planners/RRT.py:113-257) for several (scenario, run) units of the benchmark loop
 * (run_scenarios.py:202-395) at once: unit_slots trees grow concurrently, edge_slots (= 256) edges each, one chunk
 * of action_horizon actions per device pass.  State sampling with goal bias (base_planner.py:162-207, run_type 0),
 * the conditioning-goal coin (RRT.py:154-157), nearest node (:49-55), local map -> sampler -> propagation ->
 * collision (:157-184), node insertion (:195-207), the goal / iteration-cap test, the final node selection
 * (:220-257) and the path back-trace (base_planner.py:342-363) all run on the device; units are popped from a
 * device-side queue, so a pass needs no host decision and no device->host copy.  Random numbers: Philox4x32-10
 * streams keyed by the unit's seed (a unit's result does not depend on which units run beside it). */
typedef struct dt_plan dt_plan;

typedef struct {
  int32_t unit_slots;        /* U: trees grown concurrently (1..64) */
  int32_t edge_slots;        /* S: edges in flight per tree; 256 */
  int32_t node_cap;          /* nodes per tree (root included) */
  int32_t action_horizon;    /* h: actions per chunk (8) */
  int32_t n_sched;           /* entries of sched_chunks (1..8) */
  int32_t sched_chunks[8];   /* chunks per edge by the parent's visit count: prop_duration[i] // action_horizon */
  int32_t iteration_cap;     /* a unit ends after this many chunk expansions (sampler calls) unless it reaches the goal */
  int32_t ode_steps;         /* K: planning_diffusion_iters */
  int32_t max_units;         /* units this plan can hold results for; unit ids are 0..max_units-1 */
  int32_t max_path;          /* rows reserved per unit for the final path */
  float goal_sample_rate;    /* 0.15 (RRT.py:23) */
  float goal_conditioning_bias; /* 0.85 */
  double local_map_scale;    /* 0.2 */
  double norm[16];           /* obs mean[6], obs std[6], action mean[2], action std[2] (metadata/carmaze.pt) */
  int32_t run_type;          /* 0: uniform sampler, goal-conditioning coin (the reference's runs); 1-3: the sample is the
                              * conditioning goal, obstacle-ahead probe per node + its penalty in the final selection
                              * (RRT.py:154-157,201-205,233-237); >= 2: cells drawn from the unit's probability map */
  int32_t reserved;
} dt_plan_cfg;

typedef struct {
  float start[6];            /* start state */
  float goal[2];             /* goal position */
  float half_w, half_h;      /* map_width / 2, map_length / 2: the uniform sampler's range (base_planner.py:186-187) */
  int32_t map_slot;          /* dt_set_map_slot slot holding this unit's maze */
  uint32_t seed;
  int32_t unit_id;           /* 0..max_units-1, unique */
  int32_t cdf_slot;          /* run_type >= 2: dt_plan_set_cdf slot of env.prob_map for this unit; else -1 */
} dt_plan_unit;

typedef struct {
  int32_t unit_id;
  int32_t goal_reached;      /* the goal disc was entered */
  int32_t has_path;          /* a path was written (goal, or the node closest to the goal when the cap ran out) */
  int32_t n_states, n_actions; /* rows of the path / action arrays */
  int32_t n_nodes;           /* len(node_list) */
  int32_t iterations;        /* chunk expansions spent */
  int32_t first_pass, last_pass; /* plan-wide pass indices the unit started / ended in */
  int32_t collisions;        /* chunks that ended in a collision */
  int32_t chunks;            /* chunk expansions booked (= iterations) */
  int32_t error;             /* 2 chain deeper than 1024, 4 path longer than max_path */
} dt_plan_result;

int dt_plan_create(dt_ctx* ctx, const dt_plan_cfg* cfg, dt_plan** out);   /* needs dt_load_denoiser first */
void dt_plan_destroy(dt_plan* plan);
/* Append n units (HOST array) to the device queue; idle unit slots start on them at once.  Synchronous. */
int dt_plan_push(dt_plan* plan, const dt_plan_unit* units_host, int n, void* stream);
/* Stage a probability map (HOST, n = rows * cols float64, row-major: CarEnv.prob_map, car_env.py:98-137) in slot
 * 0..63 as its normalised cumulative sum, what np.random.choice(size, p = prob_map.ravel()) searches
 * (base_planner.py:157-160).  Synchronous. */
int dt_plan_set_cdf(dt_plan* plan, int slot, const double* prob_host, int n, void* stream);
/* Enqueue one pass over all unit_slots x edge_slots edges (no synchronisation) and a snapshot of the counters. */
int dt_plan_pass(dt_plan* plan, void* stream);
/* Counters as of the end of pass `pass_index` (0-based, one of the last four enqueued): out5 = {units popped, units
 * pushed, units finished, passes done, error bits}.  wait = 0: returns 1 when that pass has not finished yet. */
int dt_plan_counters(dt_plan* plan, int64_t pass_index, int wait, int32_t* out5);
/* Result of a finished unit: header, and (when has_path) its path (n_states, 6) and actions (n_actions, 2) into HOST
 * arrays of cap_rows rows.  Synchronous. */
int dt_plan_fetch(dt_plan* plan, int unit_id, dt_plan_result* hdr_out, float* path_out, float* actions_out, int cap_rows,
                  void* stream);
/* Test hook: node count, unit id, positions (xy_out: x[cap] then y[cap]) and parents of the tree in unit slot u. */
int dt_plan_peek_tree(dt_plan* plan, int u, int32_t* n_nodes_out, int32_t* unit_id_out, float* xy_out,
                      int32_t* parent_out, int cap, void* stream);

/* Test hook: conditioning goals (unit_slots * edge_slots, 2) and parent node indices of all edge slots. */
int dt_plan_peek_slots(dt_plan* plan, float* goals_out, int32_t* parents_out, void* stream);

/* Test hook for the tcgen05 GEMM core: C[M,N] (f32) = A[M,K] (bf16, row-major) * W[N,K]^T (bf16). */
int dt_gemm_bf16(dt_ctx* ctx, const void* A, const void* W, int64_t M, int N, int K, float* C, void* stream);

/* Test hook for the encoder's fused launch (local_map_encoder.py:63-76,112-122: a ResNet conv followed by
 * GroupNorm(C / 16 groups) [+ identity] [+ ReLU]): in (B,H,W,Cin) bf16 channel-last with Cin % 64 == 0, w [N][k*k*Cin]
 * bf16 (tap-major), gamma / beta [N] f32, resid (B,OH,OW,N) bf16 or NULL, out (B,OH,OW,N) bf16; OH*OW <= 128. */
int dt_conv2d_gn_bf16(dt_ctx* ctx, const void* in, int64_t B, int H, int W, int Cin, const void* w, int N, int k, int stride,
                      int pad, const float* gamma, const float* beta, const void* resid, int relu, void* out, void* stream);

/* Per-launch device timing of the tensor-core GEMM kernels (bench.py's roofline): between begin and
 * end every k_conv_gemm launch is bracketed by CUDA events on its own stream.  dt_profile_end
 * synchronises the device and returns the summed kernel time (ms) and the number of launches. */
int dt_profile_begin(dt_ctx* ctx);
int dt_profile_end(dt_ctx* ctx, double* gemm_ms_out, int64_t* gemm_launches_out);
/* Write the per-launch records of the last dt_profile_begin/end window as CSV
 * (index, BN, epilogue: 0 bias / 1 GroupNorm+Mish / 2 the split-K reduction kernel, 10 * group width + CTAs per
 * tile, M rows, N, K, milliseconds, TFLOP/s, K slices).  Split-K launches are recorded like any other. */
int dt_profile_csv(dt_ctx* ctx, const char* path);

/* Number of kernels this library launched since the ctx was created (bench.py's gpu_launches). */
int64_t dt_launch_count(dt_ctx* ctx);

/* Asynchronous device-side status raised by earlier launches (e.g. DT_E_INDEX); synchronises
 * `stream`, returns and clears it. */
int dt_sync_status(dt_ctx* ctx, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DITREE_H */
