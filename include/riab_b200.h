/* riab_b200.h -- C ABI of the B200-native batched step engine for RatInABox's
 * per-step hot path (Agent.update + Neurons.update for PlaceCells, GridCells,
 * BoundaryVectorCells).
 *
 * The reference (RatInABox v1.15.3) is pure Python: it has no FFI.  Its
 * extension boundary is two overridable methods,
 *     Agent.update(dt, drift_velocity, drift_to_random_strength_ratio, **kw)   ratinabox/Agent.py:160
 *     Neurons.update(**kw) -> get_state(evaluate_at, **kw)                     ratinabox/Neurons.py:145,173
 * and that is what this library sits behind.  Each entry point below names the
 * reference function it replaces.  The Python host mirror (ratinabox_b200/)
 * binds these symbols with ctypes; INTEGRATION.md shows the stub a reference
 * maintainer would add.
 *
 * Conventions
 *   - plain C: pointers, sizes, POD structs; no torch / C++ types.
 *   - "dev" pointers are CUDA device pointers owned by the caller (the Python
 *     host allocates them with torch); "host" pointers are ordinary memory.
 *   - every launch goes to the CUDA stream passed as `stream` (a cudaStream_t
 *     cast to void*; NULL = legacy default stream).  No host sync inside.
 *   - return value: 0 on success, negative riab_status otherwise;
 *     riab_last_error() gives the message (thread local).
 *   - agent state is float64 (the reference's dtype, Agent.py:197-198); firing
 *     rates / history rows are float32.
 */
#ifndef RIAB_B200_H
#define RIAB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RIAB_ABI_VERSION 2

typedef enum {
  RIAB_OK = 0,
  RIAB_ERR_INVALID = -1,   /* bad argument (the reference would assert / raise ValueError) */
  RIAB_ERR_CUDA = -2,      /* CUDA runtime error */
  RIAB_ERR_UNSUPPORTED = -3 /* outside the supported path (see DESIGN.md "out of scope") */
} riab_status;

int riab_abi_version(void);
const char* riab_last_error(void);

/* RIAB_BOUNDARY_SOLID_BOX: strict 4-compare in-environment test and clamp (Environment.py:793-806, :880-889).
 * RIAB_BOUNDARY_PERIODIC_BOX: no boundary walls are built (:130-144); positions wrap (:877-879), displacement /
 *   distance vectors take the short way round (:670-675).
 * RIAB_BOUNDARY_POLYGON: polygon boundary and / or holes, solid.  In the environment <=> strictly inside the boundary
 *   polygon and not strictly inside a hole (:807-817, shapely `contains`), decided by an even-odd ray cast over the
 *   boundary / hole walls; a position outside after the bounce loop is re-drawn uniformly inside (:892-893, the
 *   reference draws from np.random there; here a Philox stream keyed by seed / step / agent). */
typedef enum { RIAB_BOUNDARY_SOLID_BOX = 0, RIAB_BOUNDARY_PERIODIC_BOX = 1, RIAB_BOUNDARY_POLYGON = 2 } riab_boundary_mode;

/* ---------------------------------------------------------------- Environment
 * walls: (n_walls,2,2) float64, boundary walls first (Environment.py:137-144),
 * then user walls (add_wall, :330-342), then the walls of the holes (:147-160).  2D: rectangular box (solid or
 * periodic), or a polygon boundary and / or holes (solid). */
typedef struct {
  const double* walls_dev;     /* device, n_walls*4 doubles */
  int32_t n_walls;
  int32_t n_boundary_walls;    /* the first n_boundary_walls walls are the closed boundary polygon (4 for the box) */
  double extent[4];            /* left,right,bottom,top  (Environment.py:171-173) */
  int32_t boundary_mode;       /* riab_boundary_mode */
  int32_t n_hole_walls;        /* RIAB_BOUNDARY_POLYGON: walls [hole_wall0, hole_wall0 + n_hole_walls) are the edges of the holes */
  double scale;                /* Environment.scale: the wrap threshold is scale/2 on both axes (:671) */
  int32_t hole_wall0;
  int32_t reserved;
} riab_env;

/* ---------------------------------------------------------------------- Agent
 * Structure-of-arrays batched state; every pointer is a device array with
 * n_agents rows.  Mirrors the attributes Agent.update mutates (Agent.py:193-201,
 * SURVEY 8(b)). */
typedef struct {
  int64_t n_agents;
  int64_t id_offset;             /* global id of row 0 (multi-GPU shards; keys the Philox stream) */
  double* pos;                   /* (A,2)  Agent.pos */
  double* velocity;              /* (A,2)  Agent.velocity */
  double* rotational_velocity;   /* (A)    Agent.rotational_velocity */
  double* measured_velocity;     /* (A,2)  Agent.measured_velocity */
  double* measured_rotational_velocity; /* (A) */
  double* head_direction;        /* (A,2) */
  double* distance_travelled;    /* (A) */
  double* distance_to_closest_wall; /* (A) */
} riab_agents;

/* Scalar parameters of one Agent.update call.  `*_kw` are the values after the
 * per-call kwargs override (Agent.py:280-285, :353-355); the others are the
 * attributes the reference reads directly. */
typedef struct {
  double dt;
  double speed_coherence_time_kw;               /* Agent.py:283 */
  double speed_mean_kw;                         /* Agent.py:284: Rayleigh sigma of the speed process */
  double speed_mean;                            /* attribute: wall repulsion (:370) and bounce speed (:439) */
  double speed_std;                             /* attribute: ==0 -> constant speed (:310) */
  double speed_coherence_time;                  /* attribute: drift update (:340) */
  double rotational_velocity_coherence_time_kw; /* :281 */
  double rotational_velocity_std_kw;            /* :280 */
  double rotational_velocity_drift_kw;          /* :282 */
  double head_direction_smoothing_timescale;    /* attribute (:489) */
  double thigmotaxis_kw;                        /* :355 */
  double wall_repel_distance_kw;                /* :354 */
  double wall_repel_strength_kw;                /* :353 */
  double drift_to_random_strength_ratio;        /* Agent.update arg */
} riab_motion_params;

/* Optional per-step inputs / parity taps (device pointers, may be NULL).  drift_velocity and pos_mirror may
 * also point to page-locked HOST memory (unified addressing): the motion step then reads the commands and posts the
 * new positions over the bus itself, without separate copies (1 MB each way at 65 536 agents). */
typedef struct {
  const double* drift_velocity;   /* (A,2) Agent.update(drift_velocity=...), Agent.py:324-341 */
  const double* xi;               /* (A,2) injected standard normals for the two OU draws
                                     (oracle mode A); NULL -> Philox4x32-10(seed, step, agent id) */
  uint64_t seed;
  uint64_t step;                  /* step counter (Philox counter word) */
  uint8_t* collision_mask;        /* (A, RIAB_MAX_REC_ITERS, n_walls) per-iteration wall_collisions
                                     of Environment.check_wall_collisions (Environment.py:820-841) */
  int32_t* first_hit;             /* (A, RIAB_MAX_REC_ITERS) lowest colliding wall index or -1 (Agent.py:437) */
  int32_t* n_iters;               /* (A) number of collision-loop iterations executed */
  float* history_row;             /* (A,8) float32: pos.xy, vel.xy (measured), head_direction.xy,
                                     rot_vel, distance_travelled -- Agent.save_to_history, Agent.py:509-521 */
  double* pos_mirror;             /* (A,2) second copy of the new positions (e.g. a pinned host buffer) or NULL */
} riab_step_io;

#define RIAB_MAX_REC_ITERS 4
#define RIAB_MAX_BOUNCE_ITERS 32

/* Agent.update (Agent.py:160-242, random-motion branch) for n_agents agents. */
int riab_agent_update(const riab_agents* agents, const riab_env* env,
                      const riab_motion_params* prm, const riab_step_io* io, void* stream);

/* ----------------------------------------------------------------- PlaceCells */
typedef enum { RIAB_PC_GAUSSIAN = 0, RIAB_PC_GAUSSIAN_THRESHOLD = 1, RIAB_PC_DIFF_OF_GAUSSIANS = 2,
               RIAB_PC_TOP_HAT = 3, RIAB_PC_ONE_HOT = 4 } riab_pc_description;   /* Neurons.py:959-976 */
typedef enum { RIAB_GEOM_EUCLIDEAN = 0, RIAB_GEOM_LINE_OF_SIGHT = 1, RIAB_GEOM_GEODESIC = 2 } riab_wall_geometry; /* Environment.py:707-774 */

typedef struct {
  int32_t n_cells;
  int32_t description;     /* riab_pc_description */
  int32_t wall_geometry;   /* riab_wall_geometry */
  int32_t n_inner_walls;   /* walls[4:] used by line_of_sight / geodesic */
  float min_fr, max_fr;    /* Neurons.py:978-980 */
  double top_hat_width;    /* the scalar `widths` param (Neurons.py:975-976) */
  const float* packed_dev; /* device block written by riab_place_pack (size riab_place_pack_floats) */
  const double* centres_dev; /* (N,2) float64 centres, used by the exact fall-back of the wall predicates */
  /* filled by riab_place_pack: */
  float eps[8];            /* relative uncertainty band of the float32 line-of-sight predicate, per inner wall */
  int32_t ep_valid;        /* geodesic: bit k set iff end k of walls[4] lies strictly inside the box (Environment.py:748) */
  int32_t n_pad;           /* n_cells rounded up to a multiple of 4 */
  float k_uniform;         /* log2(e)/(2 w^2) when every cell has the same width w, else 0 */
  float r2_max;            /* max squared distance of a centre or box corner from the box centre */
} riab_place_cells;

/* Host-side packing of PlaceCells parameters (place_cell_centres (N,2) f64,
 * place_cell_widths (N) f64, walls (W,2,2) f64) into the float32 block the
 * kernels read.  Returns number of floats written / needed. */
int64_t riab_place_pack_floats(int32_t n_cells, int32_t n_inner_walls);
int riab_place_pack(const double* centres_host, const double* widths_host, int32_t n_cells,
                    const double* walls_host, int32_t n_walls, int32_t n_boundary_walls,
                    const double* extent, int32_t wall_geometry, riab_place_cells* meta_out, float* out_host);

/* PlaceCells.get_state(evaluate_at=None, pos=P) (Neurons.py:936-981):
 * pos_dev (n_pos,2) f64 -> out_dev (n_pos, ld_out) f32, row = position, col = cell
 * (the transposed, coalesced view of the reference's (n_cells,n_pos)). */
int riab_place_rates(const double* pos_dev, int64_t n_pos, const riab_env* env,
                     const riab_place_cells* pc, float* out_dev, int64_t ld_out, void* stream);

/* ------------------------------------------------------------------ GridCells */
typedef enum { RIAB_GC_RECTIFIED_COSINES = 0, RIAB_GC_SHIFTED_COSINES = 1 } riab_gc_description; /* Neurons.py:1203-1218 */
typedef struct {
  int32_t n_cells;
  int32_t description;
  double width_ratio;      /* Neurons.py:1064 */
  float min_fr, max_fr;
  const float* packed_dev; /* riab_grid_pack output */
  int32_t n_pad;           /* filled by riab_grid_pack */
} riab_grid_cells;
int64_t riab_grid_pack_floats(int32_t n_cells);
int riab_grid_pack(const double* gridscales_host, const double* phase_offsets_host /* (N,2) */,
                   const double* w_host /* (N,3,2) */, int32_t n_cells, const double* extent,
                   riab_grid_cells* meta_out, float* out_host);
/* GridCells.get_state (2D), Neurons.py:1172-1236 */
int riab_grid_rates(const double* pos_dev, int64_t n_pos, const riab_env* env,
                    const riab_grid_cells* gc, float* out_dev, int64_t ld_out, void* stream);

/* -------------------------------------------------------- BoundaryVectorCells */
typedef struct {
  int32_t n_cells;
  int32_t n_test_angles;   /* T = int(360/dtheta), Neurons.py:1588 */
  float min_fr, max_fr;
  const float* packed_dev;       /* riab_bvc_pack output (per-cell tuning + von Mises table tiles) */
  const double* test_dirs_dev;   /* (T,2) f64 test_directions (Neurons.py:1584-1596, duplicated-0 quirk kept) */
  int32_t n_pad;                 /* filled by riab_bvc_pack: n_cells rounded up to the cell tile (64) */
  int32_t egocentric;            /* reference_frame == "egocentric" (Neurons.py:1693-1708, FieldOfViewBVCs :1847-1887):
                                    test angles are measured from utils.get_angle(head_direction) */
} riab_bvc_cells;
int64_t riab_bvc_pack_floats(int32_t n_cells, int32_t n_test_angles);
int riab_bvc_pack(const double* tuning_distances, const double* tuning_angles, const double* sigma_distances,
                  const double* sigma_angles, int32_t n_cells, const double* test_angles, int32_t n_test_angles,
                  riab_bvc_cells* meta_out, float* out_host);
/* Packed block (float32 unless noted), Np = n_cells rounded up to 64:  s[Np] | m[Np] | 1/norm[Np] (cell order) |
 * von Mises weights, peak 1, [Np/64][T][64] in SLOT order | egocentric extras 3 Np + 2 T | int32 perm[Np] (slot -> cell:
 * the cells sorted by tuning angle) | int32 (th0, len)[Np/32]: the angular window of each 32 slots outside which every
 * weight is < 2^-30 and the integrand terms are skipped (their sum is < T 2^-30 / norm of the peak rate). */
/* BoundaryVectorCells.get_state (allocentric), Neurons.py:1617-1778.
 * scratch_dev: riab_bvc_scratch_floats(n_pos, T) float32 workspace holding dist_to_first_wall in
 * [agent tile of 32][T][32] order; first_wall_dev optional (n_pos,T) int32 (argmax wall id, Neurons.py:1677-1679). */
int64_t riab_bvc_scratch_floats(int64_t n_pos, int32_t n_test_angles);
int riab_bvc_rates(const double* pos_dev, int64_t n_pos, const riab_env* env, const riab_bvc_cells* bvc,
                   float* scratch_dev, int32_t* first_wall_dev, const double* head_direction_dev /* (n_pos,2) f64 for
                   egocentric cells; NULL = [1,0] (the reference's default, Neurons.py:1703) */,
                   float* out_dev, int64_t ld_out, void* stream);

/* --------------------------------------------------------- ObjectVectorCells
 * Neurons.py:1892-2113.  Objects live in the Environment (Environment.add_object, Environment.py:366-395): up to
 * RIAB_MAX_OBJECTS positions with an integer type each.  Cell i fires for the objects whose type equals
 * tuning_types[i]: sum of gaussian(distance) * von_mises(bearing), both with peak 1 (Neurons.py:2090-2104); with
 * walls_occlude the distance is the `line_of_sight` one (1000 behind an inner wall, Environment.py:710-730);
 * egocentric cells measure bearings from the head direction (Neurons.py:2030-2047). */
#define RIAB_MAX_OBJECTS 9
typedef struct {
  int32_t n_cells;
  int32_t n_objects;
  double objects[2 * RIAB_MAX_OBJECTS];     /* Environment.objects["objects"], (n_objects,2) */
  int32_t object_types[RIAB_MAX_OBJECTS];   /* Environment.objects["object_types"] */
  int32_t walls_occlude;                    /* 1: wall_geometry "line_of_sight", 0: "euclidean" (Neurons.py:1937-1940) */
  int32_t egocentric;                       /* reference_frame == "egocentric" */
  float min_fr, max_fr;
  const float* packed_dev;                  /* device block written by riab_ovc_pack (riab_ovc_pack_floats floats) */
  int32_t n_pad;                            /* filled by riab_ovc_pack */
  int32_t reserved;
} riab_ovc_cells;
int64_t riab_ovc_pack_floats(int32_t n_cells);
/* tuning_angles / sigma_angles in radians (VectorCells attributes), tuning_types (n_cells) int32 */
int riab_ovc_pack(const double* tuning_distances, const double* tuning_angles, const double* sigma_distances,
                  const double* sigma_angles, const int32_t* tuning_types, int32_t n_cells, riab_ovc_cells* meta_out,
                  float* out_host);
/* ObjectVectorCells.get_state at given positions; head_direction_dev (n_pos,2) for egocentric cells, NULL = [1,0]. */
int riab_ovc_rates(const double* pos_dev, int64_t n_pos, const riab_env* env, const riab_ovc_cells* ovc,
                   const double* head_direction_dev, float* out_dev, int64_t ld_out, void* stream);

/* ------------------------------------------------------- Neurons.update extras
 * OU noise (Neurons.py:153-160) and spikes (Neurons.py:681-684) for a block of
 * rates already written to rates_dev.  noise_dev (A,N) f32 state (NULL when
 * noise_std == 0); spikes (A, 4*ceil(N/128)) uint32 words (layout: riab_rates_out.spikes_row), NULL to skip. */
typedef struct {
  float noise_std, noise_coherence_time, dt;
  uint64_t seed, step;
  int64_t id_offset;
  int32_t population_id;    /* distinguishes the Philox streams of populations of one Agent */
} riab_neuron_noise;

/* --------------------------------------------------------------- fused step
 * One launch = Agent.update for every agent + Neurons.update of ONE population
 * (motion -> rates [-> noise] [-> spikes] -> history row).  `cells_kind` selects
 * which of pc / gc / bvc / ovc is read. */
typedef enum { RIAB_CELLS_PLACE = 0, RIAB_CELLS_GRID = 1, RIAB_CELLS_BVC = 2, RIAB_CELLS_OVC = 3 } riab_cells_kind;
typedef struct {
  float* rates_row;        /* (A, ld) f32: firing rates of this step (doubles as the history row) */
  int64_t ld;
  uint32_t* spikes_row;    /* (A, 4*ceil(N/128)) uint32 words, 16-byte aligned, or NULL.  Bit L of word 4B+i =
                            * spike of cell 128B + 4L + i (a warp's ballot of its lanes' i-th cell). */
  float* noise_state;      /* (A, ld) f32 OU noise state or NULL (noise_std == 0) */
  float* bvc_scratch;      /* (A, T) f32, BVC only */
} riab_rates_out;

int riab_step_fused(const riab_agents* agents, const riab_env* env, const riab_motion_params* prm,
                    const riab_step_io* io, int32_t cells_kind, const void* cells /* riab_place_cells* etc. */,
                    const riab_neuron_noise* noise, const riab_rates_out* out, void* stream);

/* Neurons.update (Neurons.py:145-171) alone, for the agents' CURRENT positions
 * (agents->pos): rates [-> noise] [-> spikes].  Used for the 2nd, 3rd ... population
 * of an Agent after riab_step_fused / riab_agent_update moved it. */
int riab_neurons_update(const riab_agents* agents, const riab_env* env, int32_t cells_kind, const void* cells,
                        const riab_neuron_noise* noise, const riab_rates_out* out, void* stream);

/* ------------------------------------------------------------- multi-step run
 * `for _ in range(n_steps): Ag.update(); [Ns.update() for Ns in Ag.Neurons]`
 * (tests/test_advanced.py:21-23) without returning to the host between steps.
 * Population 0 is fused with the motion kernel, the others use riab_neurons_update.
 * History rows go to device rings: row (next + s) % rows for step s. */
typedef struct {
  int32_t kind;                 /* riab_cells_kind */
  const void* cells;            /* riab_place_cells* / riab_grid_cells* / riab_bvc_cells* / riab_ovc_cells* */
  riab_neuron_noise noise;      /* seed/step base; step is advanced per step */
  riab_rates_out out;           /* ld, noise_state, bvc_scratch; rates_row/spikes_row are set from the rings */
  float* rates_ring;            /* (rows, A, ld) f32 */
  uint32_t* spikes_ring;        /* (rows, A, 4*ceil(N/128)) or NULL */
  int32_t ring_rows;
  int32_t ring_next;            /* in: first row to write */
} riab_population;

typedef struct {
  float* ring;                  /* (rows, A, 8) f32 agent history rows, or NULL */
  int32_t ring_rows;
  int32_t ring_next;
} riab_agent_history;

int riab_run(const riab_agents* agents, const riab_env* env, const riab_motion_params* prm, const riab_step_io* io,
             const riab_population* pops, int32_t n_pops, const riab_agent_history* hist, int64_t n_steps,
             void* stream);

/* ------------------------------------------------------------ history analytics
 * utils.bin_data_for_histogramming (utils.py:544-589) over the device history rings, pooled over agents and steps:
 * count[ix, iy] = number of (step, agent) samples whose position falls into bin (ix, iy) -- np.histogram2d with the
 * explicit edges np.arange(extent[0], extent[1] + dx, dx) (right-most edge inclusive, outside samples dropped) --
 * and sum[(ix, iy), c] = the sum of cell c's rates over those samples.  rate map = sum / max(count, 1), laid out
 * `.T[::-1, :]` by the host like the reference (Neurons.py:483-490 plot_rate_map(method="history"),
 * Agent.py:956 plot_position_heatmap). */
typedef struct {
  const float* agent_ring;     /* (agent_ring_rows, A, 8) f32 history rows (pos.xy first), riab_agent_history.ring */
  int32_t agent_ring_rows;
  int32_t agent_row0;          /* ring row of the first step to use */
  const float* rates_ring;     /* (rates_ring_rows, A, ld) f32 or NULL (occupancy only) */
  int32_t rates_ring_rows;
  int32_t rates_row0;
  int64_t n_steps, n_agents, ld;
  int32_t n_cells;
  int32_t reserved;
} riab_history_view;
int riab_history_rate_maps(const riab_history_view* h, const double* edges_x_dev, int32_t n_edges_x,
                           const double* edges_y_dev, int32_t n_edges_y, float* sum_dev /* ((nx*ny), ld), zeroed here */,
                           float* count_dev /* (nx*ny), zeroed here */, void* stream);

/* Number of kernels launched by the library since load (bench.py "gpu_launches"). */
int64_t riab_launch_count(void);
/* cudaStreamSynchronize(stream): lets a host binding wait for its steps without another CUDA binding. */
int riab_stream_synchronize(void* stream);

/* -------------------------------------------------- host-buffer (e2e) entry
 * The reference-facing call with HOST buffers: copies drift (may be NULL) to the
 * device, runs riab_step_fused, copies pos (A,2 f64) back.  Buffers should be
 * pinned for the copies to be asynchronous.  staging_* are device scratch. */
int riab_step_fused_host(const riab_agents* agents, const riab_env* env, const riab_motion_params* prm,
                         riab_step_io* io, int32_t cells_kind, const void* cells,
                         const riab_neuron_noise* noise, const riab_rates_out* out,
                         const double* drift_host, double* drift_staging_dev,
                         double* pos_out_host, void* stream);

/* Agent.update with HOST buffers, for per-step control loops (the policy-control caller of
 * ratinabox/contribs/TaskEnvironment.py:399-408 passes one drift velocity per agent and reads the positions back):
 *   1. drift_host (A,2 f64, page-locked; may be NULL) is uploaded to drift_staging_dev by a copy engine on the
 *      library's side stream -- at once, i.e. while kernels queued earlier on `stream` (the previous step's rate
 *      kernels) still run;
 *   2. the motion kernel (riab_agent_update semantics; io->drift_velocity is set to the staging buffer) runs on
 *      `stream` after that copy;
 *   3. the new positions are copied to pos_out_host (page-locked; may be NULL) on the side stream, concurrently with
 *      whatever the caller queues on `stream` next (the rate kernels).
 * riab_positions_wait() blocks until step 3 of the LAST call on this device has finished (not the stream).  The next
 * kernel that overwrites agents->pos must be ordered after that copy: riab_positions_fence(stream) inserts the
 * dependency (riab_agent_update_host does it itself). */
int riab_agent_update_host(const riab_agents* agents, const riab_env* env, const riab_motion_params* prm,
                           riab_step_io* io, const double* drift_host, double* drift_staging_dev,
                           double* pos_out_host, void* stream);
int riab_positions_wait(void);
int riab_positions_fence(void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RIAB_B200_H */
