/*
 * brs.h -- C ABI of the batched robot stepping engine (libbrs.so).
 *
 * This is the drop-in boundary for the aitk.robots `World.step` hot path
 * (BASELINE.json north_star).  The reference exposes no
 * FFI / plugin interface at all -- /root/reference/setup.py:21-44 ships no
 * code and delegates to the un-vendored `aitk` distribution (setup.py:29) -- so
 * each entry point below names the upstream *Python* call it stands in for.
 * Upstream file names are from SURVEY.md section 1 (aitk/robots/...); line numbers
 * cannot be cited because the source is not mounted.
 *
 * Conventions
 *   - plain C types only: pointers, sizes, scalars.  No torch / C++ types.
 *   - every function returns 0 on success or a negative brs_status; the text of
 *     the last failure on the calling thread is brs_last_error().
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  All
 *     work is enqueued asynchronously; the caller synchronises.
 *   - the engine owns its device buffers (HBM).  brs_device_ptr() exposes them
 *     so Python can wrap them zero-copy (torch via __cuda_array_interface__);
 *     brs_bind_output() redirects an observation buffer to caller memory, e.g.
 *     a slice of an NCCL all-gather buffer.
 *   - there is no CPU fallback: without a CUDA device brs_create() fails.
 *   - an engine steps one launch at a time: brs_step calls on one engine must be
 *     ordered (same stream, or events between streams); different engines are
 *     independent.
 *
 * Angles are radians, lengths are world units (cm upstream), y grows downward.
 */
#ifndef BRS_H
#define BRS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BRS_ABI_VERSION 3
#define BRS_MAX_ROBOTS 8 /* robots per world */

typedef enum brs_status {
  BRS_OK = 0,
  BRS_ERR_INVALID = -1, /* bad argument / configuration */
  BRS_ERR_CUDA = -2,    /* CUDA runtime error, see brs_last_error() */
  BRS_ERR_NO_DEVICE = -3,
  BRS_ERR_UNSUPPORTED = -4
} brs_status;

/* Robot slot r of every world (upstream robot.py: Robot.__init__, bounding box). */
typedef struct brs_robot_desc {
  double bbox[4]; /* min_x, max_x, min_y, max_y in the robot frame */
  double height;  /* 0..1, camera strip height of this robot (Hit.height) */
  double vx_ramp, vy_ramp, va_ramp; /* max velocity change per second */
  uint8_t rgba[4]; /* colour seen by other robots' cameras */
  int32_t _pad;
} brs_robot_desc;

/* upstream devices/rangesensors.py: RangeSensor(position, a, max, width) */
typedef struct brs_range_desc {
  int32_t robot;  /* robot slot */
  int32_t n_rays; /* 1 (width == 0) or 3 (-w/2, 0, +w/2) */
  double dist_from_center, dir_from_center; /* mount point, polar, robot frame */
  double a;     /* heading offset, radians */
  double max;   /* maximum range */
  double width; /* beam width, radians */
} brs_range_desc;

/* upstream devices/cameras.py: Camera(width, height, a, ...) */
typedef struct brs_camera_desc {
  int32_t robot;
  int32_t width, height; /* image size; width % 4 == 0 */
  int32_t _pad;
  double dist_from_center, dir_from_center;
  double fov; /* radians */
  double max_range;
  double color_fade, size_fade; /* colorsFadeWithDistance, sizeFadeWithDistance */
} brs_camera_desc;

/* upstream devices/lightsensors.py / smellsensors.py: mount point only */
typedef struct brs_point_desc {
  int32_t robot;
  int32_t _pad;
  double dist_from_center, dir_from_center;
} brs_point_desc;

typedef struct brs_config {
  int32_t abi_version; /* BRS_ABI_VERSION */
  int32_t device;      /* CUDA device ordinal */
  int32_t n_worlds;    /* worlds in this engine (this rank's shard) */
  int32_t n_robots;    /* robots per world, 1..BRS_MAX_ROBOTS */
  int32_t seg_cap;     /* wall-segment capacity per world, multiple of 4 */
  int32_t n_bulbs;     /* bulbs per world (0 = none) */
  int32_t grid_w, grid_h; /* smell grid cells (0 = no grid) */
  int32_t n_range, n_camera, n_light, n_smell; /* devices per world */
  double smell_cell_size;
  double light_multiplier;
  /* camera sky / ground colour = base + grad * |j - H/2| / (H/2) */
  double sky_rgb[3], sky_grad[3], ground_rgb[3], ground_grad[3];
  const brs_robot_desc* robots;   /* [n_robots] */
  const brs_range_desc* ranges;   /* [n_range] */
  const brs_camera_desc* cameras; /* [n_camera], all the same width x height */
  const brs_point_desc* lights;   /* [n_light] */
  const brs_point_desc* smells;   /* [n_smell] */
} brs_config;

/* Device buffers, all leading dimension = world.                      dtype, shape        */
typedef enum brs_buffer {
  BRS_BUF_SEG = 0,      /* wall store, SoA rows x1|y1|x2|y2|rgba       u32/f32 [B][5][seg_cap]  */
  BRS_BUF_SEG_COUNT,    /* used segments per world                     i32 [B]                  */
  BRS_BUF_WORLD_SIZE,   /* width, height                               f32 [B][2]               */
  BRS_BUF_BULBS,        /* x, y, z, brightness                         f32 [B][n_bulbs][4]      */
  BRS_BUF_GRID,         /* smell field                                 f32 [B][grid_h][grid_w]  */
  BRS_BUF_POSE,         /* x, y, a                                     f64 [B][R][3]            */
  BRS_BUF_VEL,          /* actual vx, vy, va                           f64 [B][R][3]            */
  BRS_BUF_TARGET_VEL,   /* target vx, vy, va (Robot.move)              f64 [B][R][3]            */
  BRS_BUF_STALLED,      /* collision flag of the last step             u8  [B][R]               */
  BRS_BUF_RANGE,        /* RangeSensor.get_distance()                  f32 [B][n_range]         */
  BRS_BUF_CAMERA,       /* Camera.take_picture(), RGBA                 u8  [B][n_camera][H][W][4] */
  BRS_BUF_LIGHT,        /* LightSensor value                           f32 [B][n_light]         */
  BRS_BUF_SMELL,        /* SmellSensor value                           f32 [B][n_smell]         */
  BRS_BUF_COUNT_
} brs_buffer;

/* brs_step flags */
#define BRS_STEP_MOVE 1u    /* Robot.step for every robot (motion + collision) */
#define BRS_STEP_SENSE 2u   /* RangeSensor / LightSensor / SmellSensor update  */
#define BRS_STEP_CAMERA 4u  /* Camera rays + take_picture into BRS_BUF_CAMERA  */
#define BRS_STEP_ALL 7u

typedef struct brs_engine brs_engine;

int brs_abi_version(void);
const char* brs_last_error(void);

/* Number of CUDA devices visible, or a negative status. */
int brs_device_count(void);

/* World(...) + add_robot/add_device for every world of the batch
 * (upstream world.py: World.__init__, add_robot; robot.py: add_device). */
int brs_create(const brs_config* cfg, brs_engine** out);
int brs_destroy(brs_engine* e);

/* Device pointer / byte size of an engine buffer (zero-copy hand-off). */
int brs_device_ptr(brs_engine* e, int which, void** ptr, size_t* bytes);

/* Redirect an observation buffer (BRS_BUF_RANGE..BRS_BUF_SMELL) to caller-owned
 * device memory of at least the engine's size for it.  ptr == NULL restores
 * the engine's own buffer. */
int brs_bind_output(brs_engine* e, int which, void* device_ptr);

/* Host <-> device copies of a whole buffer or of worlds [first, first+count).
 * `host` may be pageable or pinned; copies are asynchronous on `stream` when
 * pinned.  (upstream: World.add_wall / add_bulb / Robot.set_pose / Robot.move
 * write these fields; Robot.get_pose / device getters read them.) */
int brs_upload(brs_engine* e, int which, const void* host, int first_world, int n_worlds, void* stream);
int brs_download(brs_engine* e, int which, void* host, int first_world, int n_worlds, void* stream);

/* World.step(time_step) over every world of the engine: one fused kernel
 * launch (upstream world.py: World.step -> robot.py: Robot.step, Robot.update
 * -> devices/*.update; cameras.py: Camera.take_picture when BRS_STEP_CAMERA).
 * RangeSensor observations are the float32 rounding of the float64 hit distance
 * (upstream keeps reading = d / max and returns reading * max; float32 cannot
 * see that round trip). */
int brs_step(brs_engine* e, double time_step, unsigned flags, void* stream);

/* World.steps(n): n consecutive brs_step calls with the inputs (target velocities) as they are
 * now (upstream world.py: World.steps(steps) / World.seconds(seconds) -- the loop config 1 of
 * BASELINE.json is written as: 1000 steps of one world).  Same results, bit for bit, as n
 * calls of brs_step; every step writes all of its observations, the buffers end up holding the
 * last step's.  ONE launch when the active plan supports it (the fused, solo, solo range and
 * mini range kernels: a world stays with one warp / team for the whole launch -- on the fused
 * kernel a tile's steps are ordered across CTAs by per-tile epoch words -- so the kernel's
 * pipeline runs on across the step boundaries instead of filling and draining once per step);
 * n launches when sensor noise is on, with BRS_STEPS_ONE_LAUNCH=0, or on the opt-in plans. */
int brs_steps(brs_engine* e, double time_step, unsigned flags, int n_steps, void* stream);
/* How many kernel launches brs_steps(flags, n_steps) issues on this engine as configured now. */
int brs_steps_launches(brs_engine* e, unsigned flags, int n_steps, int* launches);
/* The launch shape of a one-launch brs_steps: worlds per tile (the fused kernel takes larger tiles
 * than for a single step -- no per-step tail to balance), tiles, CTAs, dynamic shared memory. */
int brs_steps_info(brs_engine* e, int* worlds_per_tile, int* n_tiles, int* ctas, int* smem_bytes);

/* The fused picture gather.  With n_peers > 0 the kernel that paints BRS_BUF_CAMERA also stores
 * every picture at peer_ptrs[k] -- memory of another GPU mapped into this process (CUDA IPC,
 * NVLink peer access), laid out like the engine's own camera output.  Binding, on every rank,
 * that rank's slice of every other rank's gather buffer turns World.step into step + all-gather
 * of the pictures in one kernel; a stream barrier across the ranks is all that remains.
 * Supported for the solo launch plan (one robot, one camera); n_peers = 0 unbinds.  No upstream
 * counterpart (upstream has no batch to gather). */
int brs_bind_peer_outputs(brs_engine* e, int which, int n_peers, void* const* peer_ptrs);

/* Memory for that gather: brs_peer_alloc = cudaMalloc + the allocation's CUDA IPC handle (64
 * bytes; ship it to the other ranks with any transport), brs_peer_open = map a peer's handle
 * into this process with NVLink peer access from this engine's GPU (cudaIpcOpenMemHandle,
 * lazy peer access).  brs_peer_close / brs_peer_free undo them. */
int brs_peer_alloc(brs_engine* e, size_t bytes, void** dev_ptr, unsigned char* handle64);
int brs_peer_free(brs_engine* e, void* dev_ptr);
int brs_peer_open(brs_engine* e, const unsigned char* handle64, void** dev_ptr);
int brs_peer_close(brs_engine* e, void* dev_ptr);

/* Seeded sensor noise (upstream: World(seed=...) seeds the sensors' noise; its model is not
 * in the mounted tree, this one is the engine's own and is mirrored by the oracle's
 * Switches.SENSOR_NOISE).  Off until called with a non-zero deviation.  Every sample is
 * Philox4x32-10(counter = (first_world + world, kind << 24 | sensor, step), key = seed) ->
 * Box-Muller N(0,1): RangeSensor reading += range_std * z, clamped to [0, max]; LightSensor
 * value *= 1 + light_rel_std * z, clamped at 0.  `first_world`: global index of this engine's
 * world 0 (a multi-GPU shard passes its offset, so noise does not depend on the sharding).
 * The step counter restarts at 0 and advances once per brs_step / brs_step_host that senses. */
int brs_set_noise(brs_engine* e, unsigned long long seed, unsigned first_world, double range_std, double light_rel_std);

/* brs_step with the EXECUTED work counted on the device (for the roofline report; the step
 * is a real one, state advances): *ray_tests = ray-segment tests the search loops ran (table
 * entries that survived the range / cone culls x rays, plus direct-path segments x rays),
 * *edge_tests = box-edge-vs-wall tests the collision classifier ran (4 per wall inside a
 * robot's disc).  brs_count_tests() is the algorithmic count the culls are measured against.
 * Synchronises `stream`. */
int brs_step_counted(brs_engine* e, double time_step, unsigned flags, void* stream,
                     unsigned long long* ray_tests, unsigned long long* edge_tests);

/* Same step driven with HOST buffers (the end-to-end path): copies target
 * velocities host->device, steps, and copies every non-NULL observation back
 * device->host.  Large batches are cut into world ranges so that the kernel of one range
 * overlaps the device->host copies of the previous one (an internal copy stream); `stream`
 * waits for the last copy, so the call is ordered on `stream` like a single operation.
 * h_* should be pinned (brs_host_alloc).  Stands in for the caller-side loop upstream users
 * write around World.step: set Robot.move targets, step, read get_distance / take_picture. */
int brs_step_host(brs_engine* e, double time_step, unsigned flags,
                  const double* h_target_vel, /* [B][R][3] or NULL */
                  double* h_pose,             /* [B][R][3] or NULL */
                  uint8_t* h_stalled,         /* [B][R] or NULL */
                  float* h_range, uint8_t* h_camera, float* h_light, float* h_smell,
                  void* stream);

/* Pinned host memory for brs_step_host, allocated and first-touched on the NUMA node of the
 * engine's GPU (the calling thread is moved to the GPU's local CPUs for the allocation).
 * No upstream counterpart: upstream observations are Python objects. */
int brs_host_alloc(brs_engine* e, size_t bytes, void** host);
int brs_host_free(brs_engine* e, void* host);

/* Top-down pictures of worlds [first_world, first_world + n_worlds) in their CURRENT state
 * (upstream world.py: the World's picture of itself -- display / take_picture; upstream
 * devices/cameras.py GroundCamera: the downward camera under a robot.  The exact look is not in
 * the mounted tree; this is the engine's own definition, restated in oracle/topdown.py):
 * RGBA [n_worlds][img_h][img_w][4] into device memory `d_out`.
 *   robot < 0:  pixel (px, py) looks at world point ((px + .5) W / img_w, (py + .5) H / img_h);
 *   robot >= 0: a view_w x view_h window (world units) centred on that robot, row 0 towards its
 *               front, the robot itself not drawn.
 * Ground colour, bulbs as discs of bulb_px pixels radius, walls as lines line_px pixels thick in
 * list order, robots as filled boxes in list order (0 = defaults 1.5 / 2.5).  Not part of
 * brs_step: a separate launch on `stream`. */
int brs_render_world(brs_engine* e, int first_world, int n_worlds, int img_w, int img_h, int robot, double view_w,
                     double view_h, double line_px, double bulb_px, void* d_out, void* stream);

/* Block the calling host thread until `stream` has drained. */
int brs_synchronize(brs_engine* e, void* stream);

/* Ray-segment tests one brs_step(flags) performs per call over the whole
 * engine, counted algorithmically: needs the segment counts,
 * so it reads BRS_BUF_SEG_COUNT back (synchronises `stream`). */
int brs_count_tests(brs_engine* e, unsigned flags, void* stream, unsigned long long* tests);

/* Launch geometry chosen for this engine (for reports): threads per CTA,
 * CTAs in the persistent grid, dynamic shared memory bytes. */
int brs_launch_info(brs_engine* e, int* threads, int* ctas, int* smem_bytes, int* sm_count);

/* Tiling chosen for this engine (for reports): worlds stepped per CTA pass,
 * number of tiles, worker threads per CTA (the CTA has `threads` of
 * brs_launch_info = workers + one helper warp), camera columns per worker
 * thread, shared-origin ray groups per world. */
int brs_tile_info(brs_engine* e, int* worlds_per_tile, int* n_tiles, int* workers, int* cols_per_thread, int* n_groups);

/* Which of the two launch plans brs_create chose for this engine (for reports and tests):
 * *wide = 0: the fused kernel (one launch per step; Robot.step of tile n+1 overlaps the rays
 * of tile n inside a CTA), 1: the wide path (prep kernel + step kernel chained with
 * programmatic dependent launch; see csrc/brs_wide.cuh), 2: the fused kernel fed by the prep
 * kernel (its helper team skips the float64 pose proposal), 3: the solo path (one warp per
 * one-robot camera world, one launch; see csrc/brs_solo.cuh), 4: the solo range path (one warp
 * per world for a lone robot whose range sensors share one mount, no camera), 5: the solo multi
 * path (one warp per world for several robots without cameras).  All produce
 * identical bits (paths 0-3; path 4 searches every ray through the shared-origin table, which
 * the other paths do for mounts with >= BRS_GROUP_MIN rays).
 * *launches_per_step = kernels one brs_step enqueues. */
int brs_path_info(brs_engine* e, int* wide, int* launches_per_step);

/* FP32 FFMA throughput microbenchmark on the engine's device: runs `iters`
 * dependent-chain FFMA sweeps on every SM and returns TFLOP/s (2 flop / FFMA)
 * timed with CUDA events.  Used only for the roofline denominator. */
int brs_measure_fp32_peak(int device, int iters, double* tflops, double* sm_clock_mhz);

/* HBM write-bandwidth microbenchmark on `device`: streams `bytes` of 16-byte stores
 * (mode 0: grid-stride st.global.cs, 1: grid-stride default cache policy, 2: cudaMemsetAsync,
 * 3: CTA-contiguous 128 KB chunks of st.global.cs -- the shape of the PAINT phase), best of
 * `reps` launches timed with CUDA events, in GB/s.  Roofline context only. */
int brs_measure_write_bw(int device, size_t bytes, int mode, int reps, double* gbs);

#ifdef __cplusplus
}
#endif
#endif /* BRS_H */
