// brs_kernels.cuh -- sm_100a kernels of the batched World.step engine.
//
// One persistent CTA steps a TILE of TW consecutive worlds at a time; tiles are
// handed out by a global atomic counter (dynamic scheduling, self-resetting).  Every
// phase of World.step runs data-parallel over the whole tile, so barriers and
// per-world set-up are amortised over TW worlds and all threads work in every
// phase whatever the shape of a single world.  The CTA is warp-specialised:
//
//   HELPER team (threads >= NW): Robot.step of tile n+1
//     TMA     the tile's wall store (SoA rows x1|y1|x2|y2|rgba, TW*20*cap contiguous
//             bytes) lands in shared memory through ONE bulk copy (cp.async.bulk +
//             mbarrier, SASS UBLKCP), double buffered.  Worlds too large for that
//             ("big" shape) are read once from L2/HBM with 16-byte loads instead.
//     PREP    thread per (world, robot): FP64 pose proposal, committed and proposed
//             box corners; all helper threads: FP32 segment table;
//     COLLIDE threads per (world, robot, wall): conservative FP32 disc cull, FP32 edge
//             classifier with an explicit error bound, then the exact FP64 4-edge test
//             (IEEE ops, no FMA contraction: decisions equal the float64 reference's);
//             robot-vs-robot pairs in both orders so that the list-order dependence of
//             Robot.step reduces to bit masks;
//     COMMIT  thread per (world, robot): resolve the masks in list order, commit or
//             stall, write pose / velocity / stalled, publish the box as search
//             segments and the origins / cones of the robot's ray groups.
//
//   WORKER team (threads < NW): sensors and pictures of tile n
//     TABLE   rays that share an origin (all columns of a Camera, range sensors on one
//             mount) get a per-(origin, segment) table m = n/na, q = d/na that makes
//             the inner test division free:
//                 1/t = m . r      u/t = r x q      hit: 0 <= u/t <= 1/t, 1/t >= 1
//             entries no ray of the group can reach (range, cone) are culled and the
//             survivors compacted in index order;
//     SENSE   thread per ray (or SL lanes per ray when a tile has few rays): FP32
//             nearest-hit search (4 FP32 + 2 compares + 2 selects per test), warp-
//             shuffle reduction over the SL lanes, then ONE FP64 refinement of the
//             winning segment so distances and pixel rounding follow the float64
//             reference;
//     PAINT   Camera.take_picture: the tile's pictures are one contiguous HBM block,
//             written with coalesced 16-byte (4 x uchar4) streaming stores.
//
// The teams meet once per tile (__syncthreads); inside a team, phases are separated by
// named barriers.  Everything the helper hands to the workers is double buffered.
//
// Tensor cores are unused on purpose: the work is FP32 cross products with no
// dense contraction (BASELINE.json north_star).
//
// Behavioural source: upstream aitk/robots/{world,robot,utils}.py and
// devices/{rangesensors,cameras,lightsensors,smellsensors}.py as restated in
// oracle/aitk_oracle.py (the reference tree itself carries no code,
// /root/reference/setup.py:21-44).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef BRS_PAINT_UNROLL
#define BRS_PAINT_UNROLL 4  // rows in flight per thread in the PAINT fast path
#endif

#ifndef BRS_PAINT_ZONES
#define BRS_PAINT_ZONES 0  // 1: constant-quad fast rows in the fused PAINT (measured SLOWER on configs 2 and 4, profiles/r2: kept off)
#endif

namespace brs {

constexpr int kPaintUnroll = BRS_PAINT_UNROLL;

// ---- device-side descriptors (host-precomputed from the public brs_*_desc) ----
struct RobotDev {
  double cox[4], coy[4];  // bbox corner k in the a=0 frame: dist_k * cos/sin(atan2(-cx, cy) + pi/2)
  double height;
  double vx_ramp, vy_ramp, va_ramp;
  double radius;  // circumradius of the box (collision cull)
  uint32_t rgba;
  uint32_t _pad;
};
struct RangeDev {
  int robot, n_rays;
  int group, _pad;      // origin group (shared-origin table) or -1: direct path
  double mx, my;        // mount offset in the a=0 frame: dfc*cos(dir+pi/2), dfc*sin(dir+pi/2)
  double cr[3], sr[3];  // cos/sin(sensor.a - incr_k)
  double max;
};
struct CameraDev {
  int robot, width, height, group;
  double mx, my;
  double max_range, color_fade, size_fade;
};
struct PointDev {
  int robot, _pad;
  double mx, my;
};
struct GroupDev {  // rays sharing one origin: robot + mount offset, and what they can reach
  int robot, has_cone;
  double mx, my;
  // cone that contains every ray of the group, robot frame, widened by a margin:
  // unit vectors of its lower / upper boundary and of its axis (has_cone: width < 180 deg)
  float lo_c, lo_s, hi_c, hi_s, ax_c, ax_s;
  float range;  // longest ray of the group, widened by a margin
  int cam_rays;    // camera columns that search this group's table (walls only)
  int range_rays;  // range-sensor rays that search it (walls and robots)
  int _pad;
};
struct __align__(16) GroupCull {  // per (world, group), world frame
  float lo_x, lo_y, hi_x, hi_y, ax_x, ax_y, range2;
  int has_cone;
  float ox, oy;  // origin rounded to FP32 (search only; refinement uses the float64 origin)
  float _pad[2];
};

// n / d for n * d < 2^32 as one multiply-high (host precomputes ceil(2^32 / d))
struct FastDiv {
  uint32_t mul, one;
};
__host__ __device__ inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.one = d <= 1u;
  f.mul = d <= 1u ? 0u : (uint32_t)((0x100000000ull + d - 1) / d);
  return f;
}

struct SoloRay;
struct MultiRay;
// Seeded sensor noise (off unless brs_set_noise enabled it): one Philox4x32-10 block per
// (world, sensor, step) -- counter = (global world index, kind << 24 | sensor, step lo, step hi),
// key = the 64-bit seed -- so a reading's noise depends on nothing but those four numbers:
// not on the launch plan, the tiling, the shard a world lives in or the order of execution.
struct NoiseParams {
  uint32_t on, k0, k1, step_lo, step_hi, world0;
  float range_std;      // RangeSensor: reading += range_std * N(0,1), clamped to [0, max]
  float light_rel_std;  // LightSensor: value *= 1 + light_rel_std * N(0,1), clamped at 0
};
struct KParams {
  int n_worlds, n_robots, seg_cap, n_bulbs, grid_w, grid_h;
  int n_range, n_camera, n_light, n_smell, n_group;
  int cam_w, cam_h;
  int TW;            // worlds per tile
  int n_tiles;
  int GS;            // threads per (world, robot) in the wall-collision phase (power of two)
  int SLG, SLD, SLL; // lanes per ray task: group path, direct range path, light path (powers of two <= 32)
  int n_gr, n_dr;    // range sensors on the group path / on the direct path
  int paint_rs;      // row slots of the paint phase
  int direct_on_helper;  // 1: the helper team runs the direct-path sensors (camera worlds with few of them); 2: only the light and smell sensors (the workers keep the ranges)
  int fused_table;   // TABLE phase: one warp per (world, group) builds and compacts in one pass
  int flat_rows;     // sky and ground rows are one colour each (no vertical gradient)
  int NW;            // worker threads (ray tasks); the CTA has NW + 32*NH threads
  int NH;            // helper warps: Robot.step of the next tile while the workers sense and paint this one
  int prepped;       // fused kernel: the robot records come from brs_prep_kernel (TMA from p.wsg) instead of PREP on the helper
  int cq_max;        // capacity of the collision candidate queue (<= BRS_CQ_MAX; overflow is tested in place)
  FastDiv dR, dSC, dSQ, dTab, dNvis, dPair, dGT, dDT, dLT, dST, dCpb, dNG;
  unsigned flags;
  double dt, smell_cell, light_mult;
  float cull_margin;
  const uint32_t* seg;      // [B][5][seg_cap]
  const int* seg_count;     // [B]
  const float* world_size;  // [B][2]
  const float* bulbs;       // [B][n_bulbs][4]
  const float* grid;        // [B][grid_h][grid_w]
  double* pose;             // [B][R][3]
  double* vel;              // [B][R][3]
  const double* tvel;       // [B][R][3]
  uint8_t* stalled;         // [B][R]
  float* range_out;         // [B][n_range]
  uint8_t* cam_out;         // [B][n_camera][H][W][4]
  float* light_out;         // [B][n_light]
  float* smell_out;         // [B][n_smell]
  const RobotDev* robots;
  const RangeDev* ranges;
  const CameraDev* cameras;
  const PointDev* lights;
  const PointDev* smells;
  const GroupDev* groups;
  const int* gr_idx;        // [n_gr] range sensors on the group path
  const int* dr_idx;        // [n_dr] range sensors on the direct path
  const double2* cam_cs;    // [n_camera][cam_w] cos/sin of the column angle i/W*fov - fov/2
  const uint32_t* row_sky;  // [cam_h] rgba per image row
  const uint32_t* row_gnd;  // [cam_h]
  unsigned int* sched;      // [2] next tile, finished CTAs (self-resetting)
  unsigned long long* phase;  // [16] cycle sums per phase (builds with -DBRS_PHASE_TIMING), else unused
  double* wsg;              // [B][R][BRS_WS_STRIDE] robot records of the wide path (prep kernel -> step kernel)
  // [2] or NULL: EXECUTED work of this launch (brs_step_counted): [0] ray-segment tests the
  // search loops ran (table entries that survived the culls x rays, direct-path segments x
  // rays), [1] box-edge-vs-wall tests the collision classifier ran (4 per disc candidate)
  unsigned long long* counters;
  // the solo path's descriptors (one robot, one camera, one ray group), copied into the kernel
  // parameters so that the per-world code reads them from the constant bank instead of HBM / L1
  RobotDev s_robot;
  CameraDev s_camera;
  GroupDev s_group;
  uint32_t s_sky, s_gnd;  // flat sky / ground colours (row 0 of the row tables)
  int solo_interleave;    // warp gw takes worlds gw, gw + n_warps, ... (else: consecutive ranges)
  NoiseParams noise;
  // fused picture gather (brs_bind_peer_outputs): the solo kernel's PAINT also stores every
  // quad into the gather buffers of n_peers other GPUs (peer memory over NVLink), at these byte
  // offsets from cam_out
  int n_peers;
  long long peer_delta[7];
  // the solo range path (brs_scan_kernel): every ray of the robot's one mount
  const struct SoloRay* s_rays;  // [s_nrays]
  int s_nrays, s_SL;             // rays per world; lanes that share one ray (power of two)
  // the solo multi path (brs_multi_kernel): several robots, no camera
  const struct MultiRay* m_rays;  // [m_nrays] every range ray of the world, robot by robot
  int m_nrays, m_bw;              // rays per world; worlds per PREP batch
  float m_reach[8];               // per robot: how far from its centre its range rays can reach
  // fused / wide paths, direct-path range sensors: per (world, robot) lists of the segments within
  // m_reach of the robot (shared memory behind the kernel's own layout; 0 = search every segment)
  int use_rlist, rl_off, rc_off;
  int use_rmask;  // direct-path range sensors walk a per-(world, robot) reach mask built in the warp (nvis <= 32, SLD == 1)
  // brs_steps: this launch runs n_steps consecutive World.steps (0 / 1 = one).  A tile (fused
  // kernel) or a world (solo kernel) belongs to ONE CTA / warp for the whole launch -- static
  // assignment, no tile counter -- so step s+1 of a world only ever follows step s of the same
  // world in program order, and the pipelines run on across the step boundaries.
  int n_steps;
  int pdl;  // fused kernel: launched under programmatic dependent launch (see brs_step_kernel's prologue)
  // fused kernel under brs_steps: [n_tiles][2] epoch words (committed, painted), both equal to
  // epoch0 or less when the launch starts; the host advances epoch0 by n_steps per launch
  unsigned* epoch;
  unsigned epoch0;
};

// per-column paint record
struct __align__(16) ColInfo {
  uint32_t rgba;  // shaded wall colour
  int top;        // rows [0, top) sky; top < 0: no wall hit, column stays transparent
  int bot;        // rows [top, bot) wall; rows [bot, H) ground
  int n_robot;    // robot strips recorded for this column
};
struct __align__(16) Strip {
  uint32_t rgba;
  int lo, hi;  // inclusive row range
  float dist;
};

// per-(world, robot) working state, doubles (shared memory; the wide path also keeps it in
// HBM between its two kernels):
//   pose[3] meta[1] cpar[2] rcs[2] dbox[8] prop[16] pad[2] = 34
// cpar = float4 (proposed centre x, y, (circumradius + margin)^2, 0): the FP32 disc of the
// collision cull.  The stride is 34 doubles = 17 x 16 bytes: records stay 16-byte aligned
// (TMA bulk copies, the float4 at CPAR) and lane-per-robot accesses spread over the banks
// (a stride of 32 doubles put every robot's field into the same bank).
#define BRS_WS_POSE 0
#define BRS_WS_META 3  // robot 0 of a world, wide path: int2 (segment count, bits of the world extent)
#define BRS_WS_CPAR 4
#define BRS_WS_RCS 6
#define BRS_WS_DBOX 8
#define BRS_WS_PROP 16
#define BRS_WS_STRIDE 34
// prop: [0..8) proposed box corners, [8..11) proposed pose, [11..14) proposed velocity, [14..16) cos/sin

// shared memory carve-up, byte offsets; mirrored on the host for the launch size
struct SmemLayout {
  int raw[2], segtab[2], gtab, gidx, gcnt, gcull[2], ws[2], org[2], cflag[2], meta[2], colinfo, strips, camcs, rowsky, rowgnd, cq, cpar, mbar, sched, total;
};
__host__ __device__ inline int align_up(int x, int a) { return (x + a - 1) / a * a; }
__host__ __device__ inline SmemLayout make_layout(int TW, int seg_cap, int R, int n_group, int n_camera, int cam_w,
                                                  int cam_h, bool stage_walls) {
  SmemLayout L;
  const int nvis = seg_cap + 4 * R;
  int o = 0;
  // the TMA-staged wall store, double buffered (big worlds read walls from L2 / HBM instead)
  for (int b = 0; b < 2; ++b) { L.raw[b] = o; o += stage_walls ? TW * 5 * seg_cap * 4 : 0; o = align_up(o, 128); }
  for (int b = 0; b < 2; ++b) { L.segtab[b] = o; o += TW * nvis * 16; }
  L.gtab = o; o += TW * n_group * nvis * 16;
  for (int b = 0; b < 2; ++b) { L.gcull[b] = o; o += TW * n_group * (int)sizeof(GroupCull); }
  L.gcnt = o; o += align_up(TW * n_group * 8, 16);
  L.gidx = o; o += align_up(TW * n_group * nvis * 2, 16);
  for (int b = 0; b < 2; ++b) { L.ws[b] = o; o += TW * R * BRS_WS_STRIDE * 8; }
  for (int b = 0; b < 2; ++b) { L.org[b] = o; o += TW * (n_group > 0 ? n_group : 1) * 16; }
  for (int b = 0; b < 2; ++b) { L.cflag[b] = o; o += TW * R * 16; }  // wall hit, pair_old mask, pair_new mask, stalled
  for (int b = 0; b < 2; ++b) { L.meta[b] = o; o += align_up(TW * 8, 16); }  // int S; float world extent
  L.colinfo = o; o += TW * n_camera * cam_w * (int)sizeof(ColInfo);
  L.strips = o; o += TW * n_camera * cam_w * 4 * (R - 1) * (int)sizeof(Strip);
  o = align_up(o, 16);
  L.camcs = o; o += n_camera * cam_w * 16;
  L.rowsky = o; o += cam_h * 4;
  L.rowgnd = o; o += cam_h * 4;
  o = align_up(o, 16);
  L.cpar = o;  // (the collision disc moved into the ws record, BRS_WS_CPAR)
  L.cq = o; o += 4 * 256;  // collision candidate queue: count + BRS_CQ_MAX packed (robot item, wall)
  L.mbar = o; o += 2 * 8;
  L.sched = o; o += 32;  // [1..2] tile per buffer; brs_steps: [3..4] outputs free?, [5..6] else the epoch to wait for
  L.total = align_up(o, 16);
  return L;
}

#ifdef __CUDACC__

#define BRS_PI_OVER_2 1.5707963267948966

// Optional phase timing (nvcc -DBRS_PHASE_TIMING): thread 0 of the workers and thread 0 of
// the helper team add the clock64() cycles of each phase to KParams::phase; brs_destroy
// prints the shares.  This is how the helper team was found to be the critical path.
// Note: BAR.SYNC does not block at issue, so a barrier wait shows up in the NEXT phase.
#ifdef BRS_PHASE_TIMING
#define BRS_PH_DECL long long ph_t0 = clock64()
#define BRS_PH(cond, k)                                                         \
  do {                                                                          \
    const long long ph_t1 = clock64();                                          \
    if (cond) atomicAdd(&p.phase[k], (unsigned long long)(ph_t1 - ph_t0));      \
    ph_t0 = ph_t1;                                                              \
  } while (0)
#else
#define BRS_PH_DECL
#define BRS_PH(cond, k)
#endif

// inlining policy of the heavy FP64 helpers (code size vs call overhead), tunable at build time
#ifndef BRS_INL_REFINE
#define BRS_INL_REFINE __forceinline__
#endif
#ifndef BRS_INL_CAMCOL
#define BRS_INL_CAMCOL __forceinline__
#endif
#ifndef BRS_INL_DIRECT
#define BRS_INL_DIRECT __forceinline__
#endif
#ifndef BRS_INL_ROBOT
#define BRS_INL_ROBOT __forceinline__
#endif
// two launch shapes: 128 workers + 32 helpers at 4 CTAs/SM (<= 96 registers; the measured
// optimum for worlds that fit several CTAs per SM), or 256 + 32 at 2 CTAs/SM for worlds
// whose shared-memory footprint leaves room for one or two CTAs only
#define BRS_SMALL_THREADS 160
#ifndef BRS_SMALL_CTAS
#define BRS_SMALL_CTAS 4
#endif
#define BRS_BIG_THREADS 288
#define BRS_BIG_CTAS 2
// one-warp teams of the wide path: resident CTAs per SM the register cap is set for
#ifndef BRS_WIDE_WARP_CTAS
#define BRS_WIDE_WARP_CTAS 20
#endif

// ---- exact (never contracted) float64 helpers --------------------------------
__device__ __forceinline__ double dm(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double da(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double ds(double a, double b) { return __dsub_rn(a, b); }

// intersect_hit of the restatement, float64, same operation order; returns
// whether the parameters lie in [0,1] and the ray parameter ua.
__device__ __forceinline__ bool isect64(double x1, double y1, double x2, double y2, double x3, double y3, double x4,
                                        double y4, double* ua_out) {
  if ((x1 == x2 && y1 == y2) || (x3 == x4 && y3 == y4)) return false;
  const double d43y = ds(y4, y3), d43x = ds(x4, x3), d21x = ds(x2, x1), d21y = ds(y2, y1);
  const double den = ds(dm(d43y, d21x), dm(d43x, d21y));
  if (den == 0.0) return false;
  const double d13y = ds(y1, y3), d13x = ds(x1, x3);
  const double na = ds(dm(d43x, d13y), dm(d43y, d13x));
  const double nb = ds(dm(d21x, d13y), dm(d21y, d13x));
  const double ua = na / den, ub = nb / den;
  if (ua_out) *ua_out = ua;
  return !(ua < 0.0 || ua > 1.0 || ub < 0.0 || ub > 1.0);
}

// ray parameter only (refinement of a hit the search already established)
__device__ __forceinline__ double ua64(double x1, double y1, double x2, double y2, double x3, double y3, double x4,
                                       double y4) {
  const double d43y = ds(y4, y3), d43x = ds(x4, x3), d21x = ds(x2, x1), d21y = ds(y2, y1);
  const double den = ds(dm(d43y, d21x), dm(d43x, d21y));
  const double d13y = ds(y1, y3), d13x = ds(x1, x3);
  const double na = ds(dm(d43x, d13y), dm(d43y, d13x));
  return na / den;
}

// boolean-only variant without divisions: sign tests are equivalent to the
// quotient tests for IEEE round-to-nearest (|na| <= |den| <=> na/den <= 1).
__device__ __forceinline__ bool isect64_bool(double x1, double y1, double d21x, double d21y, double x3, double y3,
                                             double x4, double y4) {
  if ((d21x == 0.0 && d21y == 0.0) || (x3 == x4 && y3 == y4)) return false;
  const double d43y = ds(y4, y3), d43x = ds(x4, x3);
  const double den = ds(dm(d43y, d21x), dm(d43x, d21y));
  if (den == 0.0) return false;
  const double d13y = ds(y1, y3), d13x = ds(x1, x3);
  double na = ds(dm(d43x, d13y), dm(d43y, d13x));
  double nb = ds(dm(d21x, d13y), dm(d21y, d13x));
  double aden = den;
  if (den < 0.0) { na = -na; nb = -nb; aden = -den; }
  // -0.0 numerators compare equal to 0 and pass, as -0.0/den >= 0 does
  return !(na < 0.0 || na > aden || nb < 0.0 || nb > aden);
}

// ---- TMA bulk copy + mbarrier -----------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// epoch words between CTAs (brs_steps).  The writer publishes with a release store (prior
// writes of the CTA, ordered before it by the barrier, are visible at L2 first).  The reader
// polls with a RELAXED load and reads the state the epoch guards -- poses, velocities -- with
// L2 loads (ld_state: ld.relaxed.gpu, never served from the SM's L1), issued after a barrier
// that follows the poll: an acquire load would do the same job by invalidating the whole L1
// of the SM on every tile (SASS: CCTL.IVALL), which costs the other CTAs of the SM their
// cached descriptors and tables (measured: config 3 10 % slower).
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned* a) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* a, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a), "r"(v) : "memory");
}
__device__ __forceinline__ double ld_state(const double* a) {
  double v;
  asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(a) : "memory");
  return v;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// ---- seeded sensor noise: Philox4x32-10 (Salmon et al., SC'11), Box-Muller -----------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
#define BRS_NOISE_RANGE 1u
#define BRS_NOISE_LIGHT 2u
// N(0,1) of (world w of this launch, sensor kind / index, this step)
__device__ __forceinline__ float noise_normal(const KParams& p, int w, uint32_t kind, uint32_t index) {
  uint32_t x[4];
  philox4x32_10(p.noise.world0 + (uint32_t)w, (kind << 24) | index, p.noise.step_lo, p.noise.step_hi, p.noise.k0, p.noise.k1, x);
  const float u1 = ((float)(x[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0, 1), 24 bits
  const float u2 = ((float)(x[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}
__device__ __forceinline__ float range_reading(const KParams& p, int w, int si, float v, float vmax) {
  if (p.noise.on && p.noise.range_std != 0.f)
    v = fminf(fmaxf(v + p.noise.range_std * noise_normal(p, w, BRS_NOISE_RANGE, (uint32_t)si), 0.f), vmax);
  return v;
}
__device__ __forceinline__ float light_reading(const KParams& p, int w, int li, float v) {
  if (p.noise.on && p.noise.light_rel_std != 0.f)
    v = fmaxf(v * (1.f + p.noise.light_rel_std * noise_normal(p, w, BRS_NOISE_LIGHT, (uint32_t)li)), 0.f);
  return v;
}

// ---- FP32 nearest-hit searches ------------------------------------------------
// Both keep a key whose integer order is "nearer is better" and the index of
// the winning segment; ties go to the larger index (the reference keeps the
// last of equally distant hits).

// Shared-origin form.  tab[j] = (m.x, m.y, q.x, q.y); key = bits of 1/t, larger
// is nearer, starts at bits(1.0f): only hits with t <= 1 (inside the range)
// win.  Negative or NaN-free by construction (degenerate entries are zero).
#define BRS_GKEY_INIT 0x3F800000u
// Two columns per instruction: sm_100a has packed FP32 (PTX fma.rn.f32x2 / mul.rn.f32x2 ->
// SASS FFMA2 / FMUL2), and ptxas folds a table scalar broadcast to both halves into the
// operand (R.F32 against R.F32x2.HI_LO), so a pair of columns costs 4 FP instructions per
// table entry instead of 8.  Each half is the IEEE mul / fma of the scalar form: same bits.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
template <int NR, bool SL1>
__device__ __forceinline__ void search_group(const float4* __restrict__ tab, int j0, int j1, int g, int SLrt,
                                             const float (&rx)[NR], const float (&ry)[NR], uint32_t (&key)[NR],
                                             int (&idx)[NR]) {
  const int SL = SL1 ? 1 : SLrt;
  if constexpr (NR % 2 == 0) {
    unsigned long long rx2[NR / 2], ry2[NR / 2];
#pragma unroll
    for (int k = 0; k < NR / 2; ++k) {
      rx2[k] = pack2(rx[2 * k], rx[2 * k + 1]);
      ry2[k] = pack2(ry[2 * k], ry[2 * k + 1]);
    }
#pragma unroll 4
    for (int j = j0 + (SL1 ? 0 : g); j < j1; j += SL) {
      const float4 s = tab[j];
      const unsigned long long mx = pack2(s.x, s.x), my = pack2(s.y, s.y), nqx = pack2(-s.z, -s.z), qy = pack2(s.w, s.w);
#pragma unroll
      for (int k = 0; k < NR / 2; ++k) {
        const unsigned long long invt2 = fma2(mx, rx2[k], mul2(my, ry2[k]));   // 1/t = m . r
        const unsigned long long w2 = fma2(rx2[k], qy, mul2(ry2[k], nqx));     // u/t = r x q
        float invt[2], w[2];
        unpack2(invt2, invt[0], invt[1]);
        unpack2(w2, w[0], w[1]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          // 0 <= w <= invt as one unsigned compare (negative w has the top bit set)
          if (__float_as_uint(w[h]) <= __float_as_uint(invt[h]) && (int)__float_as_uint(invt[h]) >= (int)key[2 * k + h]) {
            key[2 * k + h] = __float_as_uint(invt[h]);
            idx[2 * k + h] = j;
          }
        }
      }
    }
  } else {
#pragma unroll 4
    for (int j = j0 + (SL1 ? 0 : g); j < j1; j += SL) {
      const float4 s = tab[j];
#pragma unroll
      for (int k = 0; k < NR; ++k) {
        const float invt = fmaf(s.x, rx[k], s.y * ry[k]);     // 1/t = m . r
        const float w = fmaf(rx[k], s.w, -(ry[k] * s.z));     // u/t = r x q
        // 0 <= w <= invt as one unsigned compare (negative w has the top bit set)
        if (__float_as_uint(w) <= __float_as_uint(invt) && (int)__float_as_uint(invt) >= (int)key[k]) {
          key[k] = __float_as_uint(invt);
          idx[k] = j;
        }
      }
    }
  }
}
// reduce over the SL lanes of a group: larger key wins, ties to the larger index
__device__ __forceinline__ void group_max(uint32_t& key, int& idx, int SL, uint32_t mask) {
  for (int o = SL >> 1; o > 0; o >>= 1) {
    const uint32_t k2 = __shfl_xor_sync(mask, key, o);
    const int i2 = __shfl_xor_sync(mask, idx, o);
    if ((int)k2 > (int)key || (k2 == key && i2 > idx)) {
      key = k2;
      idx = i2;
    }
  }
}

// General form, one ray against segtab[j] = (x1, y1, ex, ey).  key = bits of
// the ray parameter t (unsigned order == float order for t >= 0; negative / NaN
// / inf parameters compare above BRS_KEY_NOHIT), smaller is nearer.
#define BRS_KEY_NOHIT 0x3F800001u  // just above 1.0f: parameters <= 1 are hits
// the two ray parameters of one (ray, segment) pair as bit patterns: tb = t along the ray,
// ub = u along the segment (one definition, so that every path rounds alike)
__device__ __forceinline__ void direct_test(const float4 s, float ox, float oy, float rx, float ry, uint32_t& tb,
                                            uint32_t& ub) {
  const float dx = ox - s.x, dy = oy - s.y;
  const float den = s.w * rx - s.z * ry;
  const float na = s.z * dy - s.w * dx;
  const float nb = rx * dy - ry * dx;
  const float inv = rcp_approx(den);
  tb = __float_as_uint(na * inv);
  ub = __float_as_uint(nb * inv);
}
template <bool SL1>
__device__ BRS_INL_DIRECT void search_direct(const float4* __restrict__ segtab, int j0, int j1, int g, int SLrt,
                                              float ox, float oy, float rx, float ry, uint32_t& key, int& idx) {
  const int SL = SL1 ? 1 : SLrt;
#pragma unroll 4
  for (int j = j0 + (SL1 ? 0 : g); j < j1; j += SL) {
    uint32_t tb, ub;
    direct_test(segtab[j], ox, oy, rx, ry, tb, ub);
    if (ub <= 0x3F800000u && tb <= key) {
      key = tb;
      idx = j;
    }
  }
}
__device__ __forceinline__ void group_min(uint32_t& key, int& idx, int SL, uint32_t mask) {
  for (int o = SL >> 1; o > 0; o >>= 1) {
    const uint32_t k2 = __shfl_xor_sync(mask, key, o);
    const int i2 = __shfl_xor_sync(mask, idx, o);
    if (k2 < key || (k2 == key && i2 > idx)) {
      key = k2;
      idx = i2;
    }
  }
}

// segment j of a world in float64: static walls from the TMA-landed store
// (float32 values are exact in float64), robot boxes from dbox.
__device__ __forceinline__ void seg64(const uint32_t* raw, int SC, int S, const double* ws, int j, double& x3,
                                      double& y3, double& x4, double& y4) {
  if (j < S) {
    x3 = (double)__uint_as_float(raw[j]);
    y3 = (double)__uint_as_float(raw[SC + j]);
    x4 = (double)__uint_as_float(raw[2 * SC + j]);
    y4 = (double)__uint_as_float(raw[3 * SC + j]);
  } else {
    const int o = (j - S) >> 2, e = (j - S) & 3, e2 = (e + 1) & 3;
    const double* box = ws + o * BRS_WS_STRIDE + BRS_WS_DBOX;
    x3 = box[2 * e];
    y3 = box[2 * e + 1];
    x4 = box[2 * e2];
    y4 = box[2 * e2 + 1];
  }
}

// float64 distance of the hit of ray (x1,y1)->(x2,y2) of length `len` on segment j.
// cast_ray computes pos = intersect_hit(...), dist = distance(pos, origin), i.e.
// sqrt((ua*rx)^2 + (ua*ry)^2); that equals ua * len to within 2 ulp of float64 (|r| is
// len times a unit vector), which is 12 orders of magnitude inside the 1e-4 contract, so
// the square root and its operands are not recomputed per ray.
__device__ BRS_INL_REFINE double refine64(const uint32_t* raw, int SC, int S, const double* ws, int j, double x1,
                                           double y1, double x2, double y2, double len) {
  double x3, y3, x4, y4;
  seg64(raw, SC, S, ws, j, x3, y3, x4, y4);
  return dm(ua64(x1, y1, x2, y2, x3, y3, x4, y4), len);
}

__device__ __forceinline__ double clamp01(double v) { return fmax(fmin(v, 1.0), 0.0); }

__device__ __forceinline__ uint32_t shade(uint32_t rgba, double sc) {
  const uint32_t r = (uint32_t)(int)dm((double)(rgba & 0xff), sc);
  const uint32_t g = (uint32_t)(int)dm((double)((rgba >> 8) & 0xff), sc);
  const uint32_t b = (uint32_t)(int)dm((double)((rgba >> 16) & 0xff), sc);
  return r | (g << 8) | (b << 16) | 0xff000000u;
}

__device__ __forceinline__ double deltav(double tv, double v, double maxdv, double dt) {
  const double lim = dm(maxdv, dt);
  return da(v, fmin(fmax(ds(tv, v), -lim), lim));
}

// mount point of a device: robot centre + R(heading) * (mount offset in the a=0 frame)
__device__ __forceinline__ void mount_point(const double* wsr, double mx, double my, double& x1, double& y1,
                                            double& ca, double& sa) {
  ca = wsr[BRS_WS_RCS];
  sa = wsr[BRS_WS_RCS + 1];
  x1 = da(wsr[BRS_WS_POSE], ds(dm(ca, mx), dm(sa, my)));
  y1 = da(wsr[BRS_WS_POSE + 1], da(dm(sa, mx), dm(ca, my)));
}

// Paint rows r0, r0+rstep, ... of the 4 columns of quad q of one picture
// (upstream Camera.take_picture: sky above, shaded wall, ground below, then the
// strips of other robots back to front).
//
// Fast path (flat sky and ground colours, every column hit a wall, no robot in sight).
// `oct`: the 8 lanes of an aligned lane octet hold the 8 adjacent quads of one 128-byte
// line of the same rows (picture width a multiple of 32).  They reduce the wall tops and
// bottoms of their 32 columns with shuffles; rows above every wall top are one constant
// quad (sky), rows inside every wall and rows below every wall bottom likewise, and are
// written by tight store loops: ~5 issue slots per 16-byte store instead of the ~22 of
// the per-pixel form (2 compares + 2 selects per pixel), which only the few rows where the
// 32 columns disagree still take.  A line is always written by ONE store instruction.
// Everything else (gradients, blank columns, robot strips) takes the generic loop.
__device__ __forceinline__ uint4 quad_pixels(const ColInfo (&ci)[4], int j, uint32_t sky, uint32_t gnd) {
  uint32_t px[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) px[k] = j < ci[k].top ? sky : (j < ci[k].bot ? ci[k].rgba : gnd);
  return make_uint4(px[0], px[1], px[2], px[3]);
}
__device__ __forceinline__ void paint_rows(uint4* __restrict__ out, const ColInfo* __restrict__ cinf,
                                           const Strip* __restrict__ strips, int strips_per_col,
                                           const uint32_t* __restrict__ rowsky, const uint32_t* __restrict__ rowgnd,
                                           bool flat, bool oct, int lane, int q, int Q, int r0, int rstep, int H) {
  ColInfo ci[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) ci[k] = cinf[4 * q + k];
  bool slow = ((ci[0].n_robot | ci[1].n_robot | ci[2].n_robot | ci[3].n_robot) != 0) ||
              ((ci[0].top | ci[1].top | ci[2].top | ci[3].top) < 0) || !flat;
  uint4* o = out + (size_t)r0 * Q + q;
  const size_t ostep = (size_t)rstep * Q;
#if BRS_PAINT_ZONES
  if (oct) {
    const uint32_t gm = 0xFFu << (lane & 24);
    slow = (__ballot_sync(gm, slow) & gm) != 0u;  // the octet takes one path together
    if (!slow) {
      const uint32_t sky = rowsky[0], gnd = rowgnd[0];
      int t0 = min(min(ci[0].top, ci[1].top), min(ci[2].top, ci[3].top));
      int t1 = max(max(ci[0].top, ci[1].top), max(ci[2].top, ci[3].top));
      int b0 = min(min(ci[0].bot, ci[1].bot), min(ci[2].bot, ci[3].bot));
      int b1 = max(max(ci[0].bot, ci[1].bot), max(ci[2].bot, ci[3].bot));
#pragma unroll
      for (int x = 1; x < 8; x <<= 1) {
        t0 = min(t0, __shfl_xor_sync(gm, t0, x));
        t1 = max(t1, __shfl_xor_sync(gm, t1, x));
        b0 = min(b0, __shfl_xor_sync(gm, b0, x));
        b1 = max(b1, __shfl_xor_sync(gm, b1, x));
      }
      b1 = max(b1, t1);
      const uint4 skyq = make_uint4(sky, sky, sky, sky), gndq = make_uint4(gnd, gnd, gnd, gnd);
      const uint4 wallq = make_uint4(ci[0].rgba, ci[1].rgba, ci[2].rgba, ci[3].rgba);
      int j = r0;
      const int e1 = min(t0, H), e2 = min(t1, H), e3 = min(b0, H), e4 = min(b1, H);
      // (not unrolled: the compiler's remainder handling costs an integer division per loop,
      //  which at 16-32 rows per thread is more than the unrolling saves)
#pragma unroll 1
      for (; j < e1; j += rstep, o += ostep) __stcs(o, skyq);
#pragma unroll 1
      for (; j < e2; j += rstep, o += ostep) __stcs(o, quad_pixels(ci, j, sky, gnd));
#pragma unroll 1
      for (; j < e3; j += rstep, o += ostep) __stcs(o, wallq);
#pragma unroll 1
      for (; j < e4; j += rstep, o += ostep) __stcs(o, quad_pixels(ci, j, sky, gnd));
#pragma unroll 1
      for (; j < H; j += rstep, o += ostep) __stcs(o, gndq);
      return;
    }
  }
#endif
  if (!slow) {
    const uint32_t sky = rowsky[0], gnd = rowgnd[0];
#pragma unroll kPaintUnroll
    for (int j = r0; j < H; j += rstep, o += ostep) __stcs(o, quad_pixels(ci, j, sky, gnd));
    return;
  }
#pragma unroll 1
  for (int j = r0; j < H; j += rstep, o += ostep) {
    const uint32_t sky = rowsky[j], gnd = rowgnd[j];
    uint32_t px[4];
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
      const ColInfo c = cinf[4 * q + k];
      uint32_t v = j < c.top ? sky : (j < c.bot ? c.rgba : gnd);
      if (c.top < 0) v = 0u;  // no wall hit: the column keeps the blank RGBA (0,0,0,0)
      const Strip* st = strips + (size_t)(4 * q + k) * strips_per_col;
      float best = 3.0e38f;
      for (int n = 0; n < c.n_robot; ++n) {
        const Strip sp = st[n];
        // back-to-front painting == the nearest covering strip wins, the later
        // one among equally distant strips
        if (j >= sp.lo && j <= sp.hi && sp.dist <= best) {
          best = sp.dist;
          v = sp.rgba;
        }
      }
      px[k] = v;
    }
    __stcs(o, make_uint4(px[0], px[1], px[2], px[3]));
  }
}

// Wall column record + robot strips of one camera ray (upstream Camera._update
// and the per-column part of take_picture).
__device__ BRS_INL_CAMCOL void camera_column(const KParams& p, const CameraDev& cd, int r, int idx, const uint32_t* raw,
                                              int SC, int S, const double* ws, double wsize, double x1, double y1,
                                              double x2, double y2, ColInfo* ci_out, Strip* st) {
  const int R = p.n_robots;
  ColInfo ci;
  ci.rgba = 0; ci.top = -1; ci.bot = -1; ci.n_robot = 0;
  const double H = (double)cd.height;
  if (idx >= 0) {
    const double dist = refine64(raw, SC, S, ws, idx, x1, y1, x2, y2, cd.max_range);
    const double dw = dist / wsize;
    const double s = clamp01(ds(1.0, dm(dw, cd.size_fade)));
    const double sc = clamp01(ds(1.0, dm(dw, cd.color_fade)));
    ci.rgba = shade(raw[4 * SC + idx], sc);
    const double high = dm(ds(1.0, s), H);
    ci.top = (int)ceil(high / 2.0);
    ci.bot = (int)ceil(ds(H, high / 2.0));
  }
  // other robots: every hit is a strip (float64 tests, few segments)
  for (int o = 0; o < R; ++o) {
    if (o == r) continue;
    const RobotDev& od = p.robots[o];
    for (int e = 0; e < 4; ++e) {
      double x3, y3, x4, y4, ua;
      seg64(raw, SC, S, ws, S + 4 * o + e, x3, y3, x4, y4);
      if (!isect64(x1, y1, x2, y2, x3, y3, x4, y4, &ua)) continue;
      const double px = da(x1, dm(ua, ds(x2, x1))), py = da(y1, dm(ua, ds(y2, y1)));
      const double ex = ds(px, x1), ey = ds(py, y1);
      const double dist = sqrt(da(dm(ex, ex), dm(ey, ey)));
      const double s = clamp01(ds(1.0, dm(dist / wsize, cd.size_fade)));
      const double sc = clamp01(ds(1.0, dm(dist / wsize, cd.color_fade)));
      const int rd_to = (int)rint(dm(H / 2.0, ds(1.0, sc)));
      const int hgt = (int)rint(dm(dm(od.height, H) / 2.0, s));
      if (hgt <= 0) continue;
      Strip sp;
      sp.rgba = shade(od.rgba, sc);
      sp.hi = cd.height - 1 - rd_to;
      sp.lo = cd.height - hgt - rd_to;
      sp.dist = (float)dist;
      st[ci.n_robot++] = sp;
    }
  }
  *ci_out = ci;
}

// ------------------------------------------------------------------------------
// The fused World.step kernel.  NR = camera columns per thread.
//
// Warp specialisation: the HELPER team (threads >= NW) runs Robot.step -- PREP,
// COLLIDE, COMMIT -- of tile n+1 while the WORKERS (threads < NW) run the sensors
// and pictures -- TABLE, SENSE, PAINT -- of tile n.  The teams meet once per tile
// (__syncthreads); inside a team phases are separated by named barriers.
// ------------------------------------------------------------------------------
// One entry of the shared-origin table: m = (ey, -ex)/na, q = d/na with d = O - P,
// na = ex*dy - ey*dx; false when no ray of the group can hit the segment (degenerate,
// its line beyond the group's range, or entirely outside the cone of its rays).
__device__ __forceinline__ bool table_entry(const float4 s, const GroupCull& gc, float4& t) {
  const float dx = gc.ox - s.x, dy = gc.oy - s.y;
  const float na = fmaf(s.z, dy, -(s.w * dx));
  const float ee = fmaf(s.z, s.z, s.w * s.w);
  bool keep = fabsf(na) > 1e-30f && na * na <= gc.range2 * ee;  // distance to the wall's line
  if (gc.has_cone) {
    const float ax = -dx, ay = -dy, bx = s.z - dx, by = s.w - dy;  // end points seen from the origin
    const bool out_lo = (gc.lo_x * ay - gc.lo_y * ax < 0.f) && (gc.lo_x * by - gc.lo_y * bx < 0.f);
    const bool out_hi = (ax * gc.hi_y - ay * gc.hi_x < 0.f) && (bx * gc.hi_y - by * gc.hi_x < 0.f);
    const bool behind = (gc.ax_x * ax + gc.ax_y * ay < 0.f) && (gc.ax_x * bx + gc.ax_y * by < 0.f);
    keep = keep && !(out_lo || out_hi || behind);
  }
  const float inv = rcp_approx(na);
  t = make_float4(s.w * inv, -s.z * inv, dx * inv, dy * inv);
  return keep;
}

__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv f) { return f.one ? n : __umulhi(n, f.mul); }
__device__ __forceinline__ void team_sync(int id, int nthreads) {
  if (nthreads == 32)
    __syncwarp();
  else
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Proposed box of robot item i (= world wl, slot r) against wall j: FP32 classification of
// the 4 edges against the wall widened by an explicit error bound, then the exact float64
// 4-edge test for what is not a certain miss.
#define BRS_CQ_MAX 252
// FP32 classification of ONE box edge (ax,ay)->(bx2,by2) of a proposal against wall s =
// (x1, y1, ex, ey), widened by an explicit error bound: false = certain miss.
// (cx, cy: the proposed centre in FP32; radius: the box circumradius)
__device__ __forceinline__ bool edge_maybe_hits_at(float cx, float cy, float radius, float cull_margin, const float4 s,
                                                   double ax, double ay, double bx2, double by2) {
  const float rad = radius + cull_margin + 1e-5f * (fabsf(cx) + fabsf(cy));
  const float cscale = 4e-7f * (fabsf(cx) + fabsf(cy) + rad + 1.f);
  const float pcx = (float)ax, pcy = (float)ay;
  const float pex = (float)(bx2 - ax), pey = (float)(by2 - ay);
  const float bx = pcx - s.x, by = pcy - s.y;  // edge start relative to the wall start
  const float den = s.w * pex - s.z * pey;
  const float na = s.z * by - s.w * bx, nb = pex * by - pey * bx;
  // error bound: coordinates carry ~6e-8 * |coordinate| of FP32 rounding, which the
  // cross products scale by the edge / wall lengths; products add ~1e-7 relative
  const float ad = fabsf(den);
  const float tol = cscale * (fabsf(s.z) + fabsf(s.w) + fabsf(pex) + fabsf(pey)) + 1e-5f * (fabsf(na) + fabsf(nb) + ad);
  const float sa_ = den < 0.f ? -na : na, sb_ = den < 0.f ? -nb : nb;
  return !(sa_ < -tol || sa_ > ad + tol || sb_ < -tol || sb_ > ad + tol);
}
__device__ __forceinline__ bool edge_maybe_hits(const KParams& p, const double* __restrict__ pr, int r, const float4 s,
                                                double ax, double ay, double bx2, double by2) {
  return edge_maybe_hits_at((float)pr[8], (float)pr[9], (float)p.robots[r].radius, p.cull_margin, s, ax, ay, bx2, by2);
}
__device__ __forceinline__ bool box_hits_wall(const KParams& p, const double* __restrict__ wsb,
                                              const uint32_t* __restrict__ rawb, const float4* __restrict__ st, int i,
                                              int j, int wl, int r, int SC) {
  const double* pr = wsb + (size_t)i * BRS_WS_STRIDE + BRS_WS_PROP;
  const float4 s = st[j];
  bool maybe = false;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int e2 = (e + 1) & 3;
    maybe |= edge_maybe_hits(p, pr, r, s, pr[2 * e], pr[2 * e + 1], pr[2 * e2], pr[2 * e2 + 1]);
  }
  if (!maybe) return false;
  const uint32_t* raw = rawb + (size_t)wl * 5 * SC;
  const double x3 = (double)__uint_as_float(raw[j]), y3 = (double)__uint_as_float(raw[SC + j]);
  const double x4 = (double)__uint_as_float(raw[2 * SC + j]), y4 = (double)__uint_as_float(raw[3 * SC + j]);
  bool hit = false;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int e2 = (e + 1) & 3;
    hit |= isect64_bool(pr[2 * e], pr[2 * e + 1], ds(pr[2 * e2], pr[2 * e]), ds(pr[2 * e2 + 1], pr[2 * e + 1]), x3, y3,
                        x4, y4);
  }
  return hit;
}

// PREP of one robot half -- the committed pose (heading trig, box corners) or the proposal
// (velocity ramp, proposed pose, its trig and corners, the FP32 disc of the collision cull)
// -- written branch-free around ONE sincos call, so that a robot is two independent
// float64 chains on two lanes.  `wsr` is the robot's BRS_WS_STRIDE record (shared memory in
// the fused kernel, HBM in the wide path's prep kernel).  Robot.step / compute_boundingbox
// of the restatement (oracle/aitk_oracle.py Robot.step).
__device__ __forceinline__ void prep_robot_half(const KParams& p, int w, int r, bool prop, double* __restrict__ wsr) {
  const int R = p.n_robots;
  const RobotDev& rd = p.robots[r];
  const size_t o = ((size_t)w * R + r) * 3;
  // (the state a previous step wrote, possibly on another SM under brs_steps: L2 loads)
  const double x = ld_state(p.pose + o), y = ld_state(p.pose + o + 1), a = ld_state(p.pose + o + 2);
  const double v0 = ld_state(p.vel + o), v1 = ld_state(p.vel + o + 1), v2 = ld_state(p.vel + o + 2);
  double vx = v0, vy = v1, va = v2, pa = a;
  if (prop) {
    const double* tv = p.tvel + o;
    vx = deltav(tv[0], v0, rd.vx_ramp, p.dt);
    vy = deltav(tv[1], v1, rd.vy_ramp, p.dt);
    va = deltav(tv[2], v2, rd.va_ramp, p.dt);
    pa = ds(a, dm(va, p.dt));
  }
  double sn, cs;
  sincos(prop ? da(-pa, BRS_PI_OVER_2) : a, &sn, &cs);
  // heading cos / sin of this half: (cos a, sin a), or (cos pa, sin pa) = (sin, cos)(pi/2 - pa)
  const double c = prop ? sn : cs, sgn = prop ? cs : sn;
  double px = x, py = y;
  if (prop) {
    // vx*sin + vy*cos*dt ; vx*cos - vy*sin*dt  (dt binds to the vy term)
    px = da(x, da(dm(vx, sn), dm(dm(vy, cs), p.dt)));
    py = da(y, ds(dm(vx, cs), dm(dm(vy, sn), p.dt)));
  }
  double* box = wsr + (prop ? BRS_WS_PROP : BRS_WS_DBOX);
  // corner = centre + R(heading) * (corner offset in the a=0 frame)
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double ox = rd.cox[k], oy = rd.coy[k];
    box[2 * k] = da(px, ds(dm(c, ox), dm(sgn, oy)));
    box[2 * k + 1] = da(py, da(dm(sgn, ox), dm(c, oy)));
  }
  if (prop) {
    double* pr = wsr + BRS_WS_PROP;
    pr[8] = px; pr[9] = py; pr[10] = pa;
    pr[11] = vx; pr[12] = vy; pr[13] = va;
    pr[14] = c; pr[15] = sgn;
    // FP32 disc of the cull: proposed centre, (circumradius + margin)^2
    const float cxf = (float)px, cyf = (float)py;
    const float rad = (float)rd.radius + p.cull_margin + 1e-5f * (fabsf(cxf) + fabsf(cyf));
    *(float4*)(wsr + BRS_WS_CPAR) = make_float4(cxf, cyf, rad * rad, 0.f);
  } else {
    wsr[BRS_WS_POSE] = x; wsr[BRS_WS_POSE + 1] = y; wsr[BRS_WS_POSE + 2] = a;
    wsr[BRS_WS_RCS] = c;
    wsr[BRS_WS_RCS + 1] = sgn;
  }
}

// Per-world meta (segment count, world extent) and the FP32 table of the static walls
// (start point and delta) of one tile, by `nth` threads (index ht).  `rawb` is the tile's
// wall store: TMA-landed shared memory (the caller has waited for it) or L2 / HBM.
__device__ __forceinline__ void build_meta(const KParams& p, int w0, int nw, int ht, int nth, int2* __restrict__ meta) {
  const int SC = p.seg_cap;
  for (int wl = ht; wl < nw; wl += nth) {
    const int w = w0 + wl;
    meta[wl].x = min(max(p.seg_count[w], 0), SC);
    meta[wl].y = __float_as_int(fmaxf(p.world_size[2 * w], p.world_size[2 * w + 1]));
  }
}
__device__ __forceinline__ void build_segtab(const KParams& p, int nw, int ht, int nth,
                                             const uint32_t* __restrict__ rawb, float4* __restrict__ segtab) {
  const int R = p.n_robots, SC = p.seg_cap, nvis = SC + 4 * R;
  // four walls per thread and step: one 16-byte load per SoA row (x1|y1|x2|y2), so that four
  // independent loads are in flight -- the big shape reads them straight from L2 / HBM
  const int SQ = SC >> 2;
#pragma unroll 2
  for (int i = ht; i < nw * SQ; i += nth) {
    const int wl = (int)fdiv(i, p.dSQ), q = i - wl * SQ;
    const uint4* raw = (const uint4*)(rawb + (size_t)wl * 5 * SC);
    const uint4 X1 = raw[q], Y1 = raw[SQ + q], X2 = raw[2 * SQ + q], Y2 = raw[3 * SQ + q];
    float4* dst = segtab + (size_t)wl * nvis + 4 * q;
    const uint32_t x1[4] = {X1.x, X1.y, X1.z, X1.w}, y1[4] = {Y1.x, Y1.y, Y1.z, Y1.w};
    const uint32_t x2[4] = {X2.x, X2.y, X2.z, X2.w}, y2[4] = {Y2.x, Y2.y, Y2.z, Y2.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float ax = __uint_as_float(x1[k]), ay = __uint_as_float(y1[k]);
      dst[k] = make_float4(ax, ay, __uint_as_float(x2[k]) - ax, __uint_as_float(y2[k]) - ay);
    }
  }
}

// PREP of one tile, run by `nth` helper threads (index ht) of the fused kernel: robot state
// HBM -> shared (two lanes per robot), collision flags cleared, per-world meta and the
// FP32 table of the static walls from the TMA-landed store.
__device__ __forceinline__ void prepare_tile(const KParams& p, int w0, int nw, int ht, int nth, bool do_move,
                                             const uint32_t* __restrict__ rawb, float4* __restrict__ segtab,
                                             double* __restrict__ wsb, int4* __restrict__ cflag,
                                             int2* __restrict__ meta, uint32_t bar, uint32_t parity, bool staged) {
  const int R = p.n_robots;
  if (p.prepped) {
    // the robot records (and the per-world meta inside them) were computed by brs_prep_kernel
    // and arrive with the walls through the same TMA barrier
    for (int i = ht; i < nw * R; i += nth) cflag[i] = make_int4(0, 0, 0, 0);
    mbar_wait(bar, parity);
    for (int wl = ht; wl < nw; wl += nth) meta[wl] = *(const int2*)(wsb + (size_t)wl * R * BRS_WS_STRIDE + BRS_WS_META);
    build_segtab(p, nw, ht, nth, rawb, segtab);
    return;
  }
  // per-world meta first: its global loads overlap the float64 chain below
  build_meta(p, w0, nw, ht, nth, meta);
  const int nhalf = do_move ? 2 : 1;
  for (int h = ht; h < nw * R * nhalf; h += nth) {
    const int i = do_move ? (h >> 1) : h;
    const bool prop = do_move && (h & 1);
    const int wl = (int)fdiv(i, p.dR), r = i - wl * R;
    prep_robot_half(p, w0 + wl, r, prop, wsb + (size_t)i * BRS_WS_STRIDE);
    if (!prop) cflag[i] = make_int4(0, 0, 0, 0);
  }
  // the tile's wall store was requested before the robot state: by now it has usually landed
  if (staged) mbar_wait(bar, parity);
  build_segtab(p, nw, ht, nth, rawb, segtab);
}

// Direct-path sensors of one tile -- range sensors on their own mount, light sensors,
// smell sensors -- run by a team of `tn` threads (index tt).  One homogeneous loop per
// device kind: mixing kinds in a warp would serialise their code paths.  Either team can
// run this: the workers for sensor-heavy worlds, the otherwise idle helper team for
// camera worlds with a few extra sensors (right after it committed the tile's poses).
// `kinds`: which device kinds this call runs (1 range, 2 light, 4 smell) -- worlds whose helper
// team has time to spare hand it the light and smell sensors while the workers keep the ranges.
__device__ __forceinline__ void sense_direct(const KParams& p, int tt, int tn, int bid, int lane, int w0, int nw,
                                             const uint32_t* __restrict__ rawb, const float4* __restrict__ segtab,
                                             const double* __restrict__ wsb, const int2* __restrict__ meta, int kinds) {
  const int R = p.n_robots, SC = p.seg_cap, nvis = SC + 4 * R;
  extern __shared__ __align__(128) unsigned char brs_dyn_smem[];
  uint8_t* const rlist = (uint8_t*)(brs_dyn_smem + p.rl_off);
  int* const rcnt = (int*)(brs_dyn_smem + p.rc_off);
  const bool use_rlist = p.use_rlist && p.n_dr > 0 && (kinds & 1);
  if (use_rlist) {
    // Which segments can a robot's range rays reach at all?  One warp per (world, robot): FP32
    // distance of the final centre to every segment (walls and the other robots' box edges)
    // against the reach of the robot's sensors (mount offset + range, widened on the host),
    // survivors compacted in index order.  The searches below then run over these lists: in a
    // 300 x 300 room with 60-unit sensors two segments in three drop out, exactly.
    const int tw = tt >> 5, nwarp = tn >> 5;
    for (int pg = tw; pg < nw * R; pg += nwarp) {
      const int wl = (int)fdiv(pg, p.dR), r = pg - wl * R;
      const int S = meta[wl].x, nseg = S + 4 * R, own = S + 4 * r;
      const double* wsr = wsb + (size_t)pg * BRS_WS_STRIDE;
      const float cx = (float)wsr[BRS_WS_POSE], cy = (float)wsr[BRS_WS_POSE + 1];
      const float reach = p.m_reach[r], reach2 = reach * reach;
      const float4* st = segtab + (size_t)wl * nvis;
      uint8_t* li = rlist + (size_t)pg * nvis;
      int n = 0;
      for (int j0 = 0; j0 < nseg; j0 += 32) {
        const int j = j0 + lane;
        bool keep = false;
        if (j < nseg && (unsigned)(j - own) >= 4u) {
          const float4 sg = st[j];
          const float dx = cx - sg.x, dy = cy - sg.y;
          const float ee = fmaf(sg.z, sg.z, sg.w * sg.w);
          float h = fmaf(dx, sg.z, dy * sg.w) * rcp_approx(ee);
          h = fminf(fmaxf(h, 0.f), 1.f);
          const float vx = fmaf(-h, sg.z, dx), vy = fmaf(-h, sg.w, dy);
          keep = !(fmaf(vx, vx, vy * vy) > reach2);  // (NaN keeps the segment)
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (keep) li[n + __popc(bal & ((1u << lane) - 1u))] = (uint8_t)j;
        n += __popc(bal);
      }
      if (lane == 0) rcnt[pg] = n;
    }
    team_sync(bid, tn);
  }
  // range sensors on their own mount, reach-masked: a warp's 32 rays belong to a few (world,
  // robot) pairs (a robot's sensors are consecutive tasks); for each pair the warp tests that
  // robot's final centre against every segment of its world -- a lane per segment, the reach
  // predicate of the lists above -- and the ballot IS the robot's list: a 32-bit mask the rays
  // then walk in index order.  No shared memory, no barrier beyond the warp.  In a 300 x 300
  // room with 60-unit sensors two tests in three drop out; the rest is evaluated exactly as below.
  if ((kinds & 1) && p.n_dr > 0 && p.use_rmask) {
    const int ntask = p.n_dr, total = nw * ntask;
    for (int ib = tt & ~31; ib < total; ib += tn) {
      const int i = ib + lane;
      const bool act = i < total;
      const int wl = act ? (int)fdiv(i, p.dDT) : 0, task = i - wl * ntask;
      const int si = act ? p.dr_idx[task] : 0;
      const int r = act ? p.ranges[si].robot : 0;
      unsigned mymask = 0;
      unsigned todo = __ballot_sync(0xffffffffu, act);
      while (todo) {
        const int ld = __ffs(todo) - 1;
        const int pw = __shfl_sync(0xffffffffu, wl, ld), pr = __shfl_sync(0xffffffffu, r, ld);
        const bool same = act && wl == pw && r == pr;
        const int Sp = meta[pw].x, nsegp = Sp + 4 * R;
        bool keep = false;
        if (lane < nsegp && (unsigned)(lane - (Sp + 4 * pr)) >= 4u) {
          const double* wsr = wsb + ((size_t)pw * R + pr) * BRS_WS_STRIDE;
          const float cx = (float)wsr[BRS_WS_POSE], cy = (float)wsr[BRS_WS_POSE + 1];
          const float reach = p.m_reach[pr], reach2 = reach * reach;
          const float4 sg = segtab[(size_t)pw * nvis + lane];
          const float dx = cx - sg.x, dy = cy - sg.y;
          const float ee = fmaf(sg.z, sg.z, sg.w * sg.w);
          float h = fmaf(dx, sg.z, dy * sg.w) * rcp_approx(ee);
          h = fminf(fmaxf(h, 0.f), 1.f);
          const float vx = fmaf(-h, sg.z, dx), vy = fmaf(-h, sg.w, dy);
          keep = !(fmaf(vx, vx, vy * vy) > reach2);  // (NaN keeps the segment)
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (same) mymask = m;
        todo &= ~__ballot_sync(0xffffffffu, same);
      }
      if (!act) continue;
      const int w = w0 + wl;
      const int S = meta[wl].x, nseg = S + 4 * R;
      const uint32_t* raw = rawb + (size_t)wl * 5 * SC;
      const double* ws = wsb + (size_t)wl * R * BRS_WS_STRIDE;
      const float4* st = segtab + (size_t)wl * nvis;
      const RangeDev& sd = p.ranges[si];
      double x1, y1, ca, sa;
      mount_point(ws + r * BRS_WS_STRIDE, sd.mx, sd.my, x1, y1, ca, sa);
      const float ox = (float)x1, oy = (float)y1;
      double best = sd.max;
      for (int k = 0; k < sd.n_rays; ++k) {
        const double dxu = ds(dm(ca, sd.cr[k]), dm(sa, sd.sr[k]));
        const double dyu = da(dm(sa, sd.cr[k]), dm(ca, sd.sr[k]));
        const double x2 = da(dm(dxu, sd.max), x1), y2 = da(dm(dyu, sd.max), y1);
        const float rx = (float)ds(x2, x1), ry = (float)ds(y2, y1);
        uint32_t key = BRS_KEY_NOHIT;
        int idx = -1;
        for (unsigned m = mymask; m; m &= m - 1) {
          const int j = __ffs(m) - 1;
          uint32_t tb, ub;
          direct_test(st[j], ox, oy, rx, ry, tb, ub);
          if (ub <= 0x3F800000u && tb <= key) {
            key = tb;
            idx = j;
          }
        }
        if (idx >= 0) best = fmin(best, refine64(raw, SC, S, ws, idx, x1, y1, x2, y2, sd.max));
      }
      p.range_out[(size_t)w * p.n_range + si] = range_reading(p, w, si, (float)best, (float)sd.max);
      if (p.counters) atomicAdd(&p.counters[0], (unsigned long long)sd.n_rays * __popc(mymask));
    }
  } else if ((kinds & 1) && p.n_dr > 0) {
    // range sensors on their own mount ...
    const int SL = p.SLD, g = tt & (SL - 1), slot = tt / SL, nslot = tn / SL;
    const uint32_t gmask = (SL == 32) ? 0xffffffffu : (((1u << SL) - 1u) << (lane & ~(SL - 1)));
    const int ntask = p.n_dr;
    for (int i = slot; i < nw * ntask; i += nslot) {
      const int wl = (int)fdiv(i, p.dDT), task = i - wl * ntask, w = w0 + wl;
      const int S = meta[wl].x, nseg = S + 4 * R;
      const uint32_t* raw = rawb + (size_t)wl * 5 * SC;
      const double* ws = wsb + (size_t)wl * R * BRS_WS_STRIDE;
      const float4* st = segtab + (size_t)wl * nvis;
      // RangeSensor.update
      const int si = p.dr_idx[task];
      const RangeDev& sd = p.ranges[si];
      const int r = sd.robot;
      double x1, y1, ca, sa;
      mount_point(ws + r * BRS_WS_STRIDE, sd.mx, sd.my, x1, y1, ca, sa);
      const float ox = (float)x1, oy = (float)y1;
      const int own_lo = S + 4 * r, own_hi = own_lo + 4;
      double best = sd.max;  // the reading stays at max unless a ray hits nearer
      for (int k = 0; k < sd.n_rays; ++k) {
        const double dxu = ds(dm(ca, sd.cr[k]), dm(sa, sd.sr[k]));
        const double dyu = da(dm(sa, sd.cr[k]), dm(ca, sd.sr[k]));
        const double x2 = da(dm(dxu, sd.max), x1), y2 = da(dm(dyu, sd.max), y1);
        const float rx = (float)ds(x2, x1), ry = (float)ds(y2, y1);
        uint32_t key = BRS_KEY_NOHIT;
        int idx = -1;
        if (use_rlist) {
          const uint8_t* li = rlist + (size_t)(wl * R + r) * nvis;
          const int n = rcnt[wl * R + r];
#pragma unroll 4
          for (int q = g; q < n; q += SL) {
            const int j = li[q];
            uint32_t tb, ub;
            direct_test(st[j], ox, oy, rx, ry, tb, ub);
            if (ub <= 0x3F800000u && tb <= key) {
              key = tb;
              idx = j;
            }
          }
          if (SL > 1) group_min(key, idx, SL, gmask);
        } else if (SL == 1) {
          search_direct<true>(st, 0, own_lo, 0, 1, ox, oy, rx, ry, key, idx);
          search_direct<true>(st, own_hi, nseg, 0, 1, ox, oy, rx, ry, key, idx);
        } else {
          search_direct<false>(st, 0, own_lo, g, SL, ox, oy, rx, ry, key, idx);
          search_direct<false>(st, own_hi, nseg, g, SL, ox, oy, rx, ry, key, idx);
          group_min(key, idx, SL, gmask);
        }
        if (g == 0 && idx >= 0) best = fmin(best, refine64(raw, SC, S, ws, idx, x1, y1, x2, y2, sd.max));
      }
      // (upstream stores reading = d / max and returns reading * max: the float32
      //  observation cannot see that round trip)
      if (g == 0) {
        p.range_out[(size_t)w * p.n_range + si] = range_reading(p, w, si, (float)best, (float)sd.max);
        if (p.counters) atomicAdd(&p.counters[0], (unsigned long long)sd.n_rays * (nseg - 4));
      }
    }
  }
  // ... light sensors: line of sight to every bulb ...
  if ((kinds & 2) && p.n_light > 0) {
    const int SL = p.SLL, g = tt & (SL - 1), slot = tt / SL, nslot = tn / SL;
    const uint32_t gmask = (SL == 32) ? 0xffffffffu : (((1u << SL) - 1u) << (lane & ~(SL - 1)));
    const int ntask = p.n_light;
    for (int i = slot; i < nw * ntask; i += nslot) {
      const int wl = (int)fdiv(i, p.dLT), li = i - wl * ntask, w = w0 + wl;
      const int S = meta[wl].x, nseg = S + 4 * R;
      const double* ws = wsb + (size_t)wl * R * BRS_WS_STRIDE;
      const float4* st = segtab + (size_t)wl * nvis;
      const PointDev& sd = p.lights[li];
      const int r = sd.robot;
      double x1, y1, ca, sa;
      mount_point(ws + r * BRS_WS_STRIDE, sd.mx, sd.my, x1, y1, ca, sa);
      const float ox = (float)x1, oy = (float)y1;
      const int own_lo = S + 4 * r, own_hi = own_lo + 4;
      double value = 0.0;
      for (int b = 0; b < p.n_bulbs; ++b) {
        const float* bl = p.bulbs + ((size_t)w * p.n_bulbs + b) * 4;
        const double bx = (double)bl[0], by = (double)bl[1], br = (double)bl[3];
        const double ex = ds(bx, x1), ey = ds(by, y1);
        const double dist = sqrt(da(dm(ex, ex), dm(ey, ey)));
        const float rx = (float)ex, ry = (float)ey;
        uint32_t key = BRS_KEY_NOHIT;
        int idx = -1;
        if (SL == 1) {
          search_direct<true>(st, 0, own_lo, 0, 1, ox, oy, rx, ry, key, idx);
          search_direct<true>(st, own_hi, nseg, 0, 1, ox, oy, rx, ry, key, idx);
        } else {
          search_direct<false>(st, 0, own_lo, g, SL, ox, oy, rx, ry, key, idx);
          search_direct<false>(st, own_hi, nseg, g, SL, ox, oy, rx, ry, key, idx);
          group_min(key, idx, SL, gmask);
        }
        if (idx < 0 && dist > 0.0 && br != 0.0) value = da(value, dm(br, p.light_mult) / dm(dist, dist));
      }
      if (g == 0) {
        p.light_out[(size_t)w * p.n_light + li] = light_reading(p, w, li, (float)value);
        if (p.counters) atomicAdd(&p.counters[0], (unsigned long long)p.n_bulbs * (nseg - 4));
      }
    }
  }
  // ... smell sensors: grid lookup at the mount point
  if ((kinds & 4) && p.n_smell > 0) {
    const int ntask = p.n_smell;
    for (int i = tt; i < nw * ntask; i += tn) {
      const int wl = (int)fdiv(i, p.dST), si = i - wl * ntask, w = w0 + wl;
      const double* ws = wsb + (size_t)wl * R * BRS_WS_STRIDE;
      const PointDev& sd = p.smells[si];
      double x1, y1, ca, sa;
      mount_point(ws + sd.robot * BRS_WS_STRIDE, sd.mx, sd.my, x1, y1, ca, sa);
      float v = 0.f;
      if (p.grid_w > 0) {
        const int gx = (int)(x1 / p.smell_cell), gy = (int)(y1 / p.smell_cell);
        if (gx >= 0 && gx < p.grid_w && gy >= 0 && gy < p.grid_h)
          v = p.grid[((size_t)w * p.grid_h + gy) * p.grid_w + gx];
      }
      p.smell_out[(size_t)w * p.n_smell + si] = v;
    }
  }

}

// Robot.step of one tile after PREP, by a team of `nth` threads (index ht; barrier id
// `bid`): the collision tests, then commit / stall and everything that hangs off the final
// pose.  The fused kernel's helper team runs it a tile ahead of the workers; in the wide
// path every thread of the CTA runs it at the start of its tile.
__device__ BRS_INL_ROBOT void collide_commit_tile(const KParams& p, int w0, int nw, int ht, int nth, int bid,
                                                    bool do_move, bool do_rays, bool do_direct,
                                                    const uint32_t* __restrict__ rawb, float4* __restrict__ segtab,
                                                    double* __restrict__ wsb, int4* __restrict__ cflag,
                                                    int2* __restrict__ meta, double2* __restrict__ org,
                                                    GroupCull* __restrict__ gcull, int* __restrict__ cq) {
  const int R = p.n_robots, SC = p.seg_cap, NG = p.n_group, nvis = SC + 4 * R;
  BRS_PH_DECL;
  if (do_move) {
    // ---------------- COLLIDE ------------------------------------------------------
    // walls: GS threads per (world, robot) stride over its segments.  FP32 disc cull
    // (distance of the proposed centre to the segment vs the box circumradius plus a
    // margin far above FP32 error), an FP32 classification of the 4 edges, then the
    // exact FP64 4-edge test for what is not a certain miss.
      // pass 1: the cull, flat over (world, robot, wall) and four items per thread and step, so
      // that four independent dependency chains are in flight (the helper team is latency
      // bound); what falls inside a disc goes to a queue
      {
        const int total = nw * R * SC, lane = ht & 31;
        // warp-uniform trip count: the queue pushes below are warp-aggregated (one shared-memory
        // atomic per warp and step instead of one per candidate)
        for (int b0 = ht - lane; b0 < total; b0 += 4 * nth) {
          int it_i[4], it_j[4];
          bool inside[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int item = b0 + lane + k * nth;
            const bool valid = item < total;
            const int i = valid ? (int)fdiv(item, p.dSC) : 0, j = valid ? item - i * SC : 0;
            const int wl = (int)fdiv(i, p.dR);
            const float4 cp = *(const float4*)(wsb + (size_t)i * BRS_WS_STRIDE + BRS_WS_CPAR);
            const float4 s = segtab[(size_t)wl * nvis + j];
            const float ax = cp.x - s.x, ay = cp.y - s.y;
            const float ee = fmaf(s.z, s.z, s.w * s.w);
            float h = fmaf(ax, s.z, ay * s.w) * rcp_approx(ee);
            h = fminf(fmaxf(h, 0.f), 1.f);  // NaN (zero-length wall) clamps to an end point
            const float vx = fmaf(-h, s.z, ax), vy = fmaf(-h, s.w, ay);
            inside[k] = valid && j < meta[wl].x && fmaf(vx, vx, vy * vy) <= cp.z;
            it_i[k] = i;
            it_j[k] = j;
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const unsigned bal = __ballot_sync(0xffffffffu, inside[k]);
            if (bal == 0u) continue;
            int posb = 0;
            if (lane == 0) posb = atomicAdd(&cq[0], __popc(bal));
            posb = __shfl_sync(0xffffffffu, posb, 0);
            if (inside[k]) {
              const int i = it_i[k], j = it_j[k];
              const int pos = posb + __popc(bal & ((1u << lane) - 1u));
              if (pos < p.cq_max) {
                cq[4 + pos] = (i << 16) | j;
              } else {  // queue full: test in place
                const int wl = (int)fdiv(i, p.dR), r = i - wl * R;
                if (box_hits_wall(p, wsb, rawb, segtab + (size_t)wl * nvis, i, j, wl, r, SC)) cflag[i].x = 1;
              }
            }
          }
        }
      }
      team_sync(bid, nth);
      BRS_PH(ht == 0, 9);
      // pass 2: the queued (robot, wall) pairs, FOUR lanes per pair (one box edge each): FP32
      // edge classifier, then the exact FP64 test of all four edges if any edge is not a
      // certain miss -- the same tests as box_hits_wall, as a chain a quarter as long
      {
        const int n = min(cq[0], p.cq_max);
        const int e = ht & 3, e2 = (e + 1) & 3;
        const uint32_t qm = 0xFu << ((ht & 31) & ~3);
        for (int c = ht >> 2; c < n; c += nth >> 2) {
          const int ent = cq[4 + c], i = ent >> 16, j = ent & 0xffff;
          const int wl = (int)fdiv(i, p.dR), r = i - wl * R;
          const double* pr = wsb + (size_t)i * BRS_WS_STRIDE + BRS_WS_PROP;
          const double ax = pr[2 * e], ay = pr[2 * e + 1], bx2 = pr[2 * e2], by2 = pr[2 * e2 + 1];
          const bool maybe = edge_maybe_hits(p, pr, r, segtab[(size_t)wl * nvis + j], ax, ay, bx2, by2);
          if (p.counters && e == 0) atomicAdd(&p.counters[1], 4ull);
          if (__ballot_sync(qm, maybe) & qm) {
            const uint32_t* raw = rawb + (size_t)wl * 5 * SC;
            const double x3 = (double)__uint_as_float(raw[j]), y3 = (double)__uint_as_float(raw[SC + j]);
            const double x4 = (double)__uint_as_float(raw[2 * SC + j]), y4 = (double)__uint_as_float(raw[3 * SC + j]);
            const bool hit = isect64_bool(ax, ay, ds(bx2, ax), ds(by2, ay), x3, y3, x4, y4);
            if ((__ballot_sync(qm, hit) & qm) && e == 0) cflag[i].x = 1;
          }
        }
      }
    BRS_PH(ht == 0, 10);
    // robots: proposal i against the committed box of o (pair_old) and, for o < i,
    // against the proposal of o (pair_new): Robot.step's list order becomes masks
      if (R > 1) {
        const int per_world = R * R * 2;
        for (int i = ht; i < nw * per_world; i += nth) {
          const int wl = (int)fdiv(i, p.dPair), k = i - wl * per_world;
          const int ri = (int)fdiv(k >> 1, p.dR), ro = (k >> 1) - ri * R, v = k & 1;
          if (ro == ri || (v && ro > ri)) continue;
          const double* wi = wsb + (size_t)(wl * R + ri) * BRS_WS_STRIDE;
          const double* wo = wsb + (size_t)(wl * R + ro) * BRS_WS_STRIDE;
          const double* pr = wi + BRS_WS_PROP;
          const double* ob = v ? wo + BRS_WS_PROP : wo + BRS_WS_DBOX;
          const double ocx = v ? wo[BRS_WS_PROP + 8] : wo[BRS_WS_POSE], ocy = v ? wo[BRS_WS_PROP + 9] : wo[BRS_WS_POSE + 1];
          const double ddx = pr[8] - ocx, ddy = pr[9] - ocy;
          const double reach = p.robots[ri].radius + p.robots[ro].radius + (double)p.cull_margin;
          if (ddx * ddx + ddy * ddy > reach * reach) continue;
          bool hit = false;
          for (int e = 0; e < 4; ++e) {
            const int e2 = (e + 1) & 3;
            const double x1 = pr[2 * e], y1 = pr[2 * e + 1];
            const double ex = ds(pr[2 * e2], x1), ey = ds(pr[2 * e2 + 1], y1);
#pragma unroll
            for (int f = 0; f < 4; ++f) {
              const int f2 = (f + 1) & 3;
              hit |= isect64_bool(x1, y1, ex, ey, ob[2 * f], ob[2 * f + 1], ob[2 * f2], ob[2 * f2 + 1]);
            }
          }
          if (hit) atomicOr(v ? &cflag[wl * R + ri].z : &cflag[wl * R + ri].y, 1 << ro);
        }
      }
    team_sync(bid, nth);
    BRS_PH(ht == 0, 11);
  }
  // ---------------- COMMIT ---------------------------------------------------------
  // 8 lanes per (world, robot), all of which replay the list order on the masks (a few
  // bit operations), then split the rest: lane 0 commits or stalls (global pose / velocity
  // / stalled, shared POSE / RCS), lanes 1-4 move one box corner each and publish one box
  // edge as an FP32 search segment, lanes 5-7 publish the origins / cones of the robot's
  // ray groups.  Every lane reads the state it needs (old or proposed, by the decision)
  // before any lane writes: the phase is a short chain instead of a long serial one.
  // (Nobody reads DBOX / POSE / RCS of ANOTHER robot here: the replay only reads the masks.)
  // Tiles with many robots (more than two such passes) use one thread per robot instead.
  if (nw * R * 8 <= 2 * nth) {
    const int sub = ht & 7, nrob = nth >> 3;
    for (int i0 = 0; i0 < nw * R; i0 += nrob) {  // warp-uniform trip count (__syncwarp below)
      const int i = i0 + (ht >> 3);
      const bool live = i < nw * R;
      const int wl = live ? (int)fdiv(i, p.dR) : 0, r = live ? i - wl * R : 0, w = w0 + wl;
      double* wsr = wsb + (size_t)(live ? i : 0) * BRS_WS_STRIDE;
      const double* pr = wsr + BRS_WS_PROP;
      bool hit = true;  // !do_move: keep the committed state
      if (do_move && live) {
        unsigned stalled_mask = 0;
        for (int q = 0; q <= r; ++q) {
          const int4 f = cflag[wl * R + q];
          // robots before q: their proposal if they moved, their old box if they stalled;
          // robots after q have not moved yet
          const unsigned before = (1u << q) - 1u;
          const unsigned m = ((unsigned)f.z & before & ~stalled_mask) | ((unsigned)f.y & (~before | stalled_mask));
          hit = f.x || m;
          if (hit) stalled_mask |= 1u << q;
        }
      }
      // ---- reads (final state = proposal if the robot moves, committed state otherwise)
      const int e = (sub - 1) & 3, e2 = (e + 1) & 3;
      const double* fbox = hit ? wsr + BRS_WS_DBOX : pr;
      const double bx1 = fbox[2 * e], by1 = fbox[2 * e + 1], bx2 = fbox[2 * e2], by2 = fbox[2 * e2 + 1];
      const double fx = hit ? wsr[BRS_WS_POSE] : pr[8], fy = hit ? wsr[BRS_WS_POSE + 1] : pr[9];
      const double fa = hit ? wsr[BRS_WS_POSE + 2] : pr[10];
      const double fc = hit ? wsr[BRS_WS_RCS] : pr[14], fs = hit ? wsr[BRS_WS_RCS + 1] : pr[15];
      const double pv0 = pr[11], pv1 = pr[12], pv2 = pr[13];
      const int S = meta[wl].x;
      __syncwarp();
      // ---- writes
      if (live && sub == 0 && do_move) {
        const size_t o = ((size_t)w * R + r) * 3;
        if (!hit) {
          p.pose[o] = fx; p.pose[o + 1] = fy; p.pose[o + 2] = fa;
          p.vel[o] = pv0; p.vel[o + 1] = pv1; p.vel[o + 2] = pv2;
          wsr[BRS_WS_POSE] = fx; wsr[BRS_WS_POSE + 1] = fy; wsr[BRS_WS_POSE + 2] = fa;
          wsr[BRS_WS_RCS] = fc; wsr[BRS_WS_RCS + 1] = fs;
        } else {
          p.vel[o] = 0.0; p.vel[o + 1] = 0.0; p.vel[o + 2] = 0.0;
        }
        p.stalled[(size_t)w * R + r] = (uint8_t)(hit ? 1 : 0);
      } else if (live && sub >= 1 && sub <= 4) {
        if (do_move && !hit) { wsr[BRS_WS_DBOX + 2 * e] = bx1; wsr[BRS_WS_DBOX + 2 * e + 1] = by1; }
        if (do_rays) {
          const float x1 = (float)bx1, y1 = (float)by1;
          segtab[(size_t)wl * nvis + S + 4 * r + e] = make_float4(x1, y1, (float)bx2 - x1, (float)by2 - y1);
        }
      } else if (live && sub >= 5 && do_rays) {
        for (int g = sub - 5; g < NG; g += 3) {
          const GroupDev& gd = p.groups[g];
          if (gd.robot != r) continue;
          // mount point: robot centre + R(heading) * (mount offset in the a=0 frame)
          const double x1 = da(fx, ds(dm(fc, gd.mx), dm(fs, gd.my)));
          const double y1 = da(fy, da(dm(fs, gd.mx), dm(fc, gd.my)));
          org[wl * NG + g] = make_double2(x1, y1);
          const float cf = (float)fc, sf = (float)fs;
          GroupCull gc;
          gc.lo_x = cf * gd.lo_c - sf * gd.lo_s; gc.lo_y = sf * gd.lo_c + cf * gd.lo_s;
          gc.hi_x = cf * gd.hi_c - sf * gd.hi_s; gc.hi_y = sf * gd.hi_c + cf * gd.hi_s;
          gc.ax_x = cf * gd.ax_c - sf * gd.ax_s; gc.ax_y = sf * gd.ax_c + cf * gd.ax_s;
          gc.range2 = gd.range * gd.range;
          gc.has_cone = gd.has_cone;
          gc.ox = (float)x1;
          gc.oy = (float)y1;
          gc._pad[0] = gc._pad[1] = 0.f;
          gcull[wl * NG + g] = gc;
        }
      }
    }
  } else {
    for (int i = ht; i < nw * R; i += nth) {
      const int wl = (int)fdiv(i, p.dR), r = i - wl * R, w = w0 + wl;
      double* wsr = wsb + (size_t)i * BRS_WS_STRIDE;
      if (do_move) {
        unsigned stalled_mask = 0;
        bool hit = false;
        for (int q = 0; q <= r; ++q) {
          const int4 f = cflag[wl * R + q];
          // robots before q: their proposal if they moved, their old box if they stalled;
          // robots after q have not moved yet
          const unsigned before = (1u << q) - 1u;
          const unsigned m = ((unsigned)f.z & before & ~stalled_mask) | ((unsigned)f.y & (~before | stalled_mask));
          hit = f.x || m;
          if (hit) stalled_mask |= 1u << q;
        }
        const double* pr = wsr + BRS_WS_PROP;
        const size_t o = ((size_t)w * R + r) * 3;
        if (!hit) {
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            p.pose[o + k] = pr[8 + k];
            p.vel[o + k] = pr[11 + k];
            wsr[BRS_WS_POSE + k] = pr[8 + k];
          }
          wsr[BRS_WS_RCS] = pr[14];
          wsr[BRS_WS_RCS + 1] = pr[15];
#pragma unroll
          for (int k = 0; k < 8; ++k) wsr[BRS_WS_DBOX + k] = pr[k];
        } else {
#pragma unroll
          for (int k = 0; k < 3; ++k) p.vel[o + k] = 0.0;
        }
        p.stalled[(size_t)w * R + r] = (uint8_t)(hit ? 1 : 0);
      }
      if (do_rays) {
        const double* box = wsr + BRS_WS_DBOX;
        float4* dst = segtab + (size_t)wl * nvis + meta[wl].x + 4 * r;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int e2 = (e + 1) & 3;
          const float x1 = (float)box[2 * e], y1 = (float)box[2 * e + 1];
          const float x2 = (float)box[2 * e2], y2 = (float)box[2 * e2 + 1];
          dst[e] = make_float4(x1, y1, x2 - x1, y2 - y1);
        }
        for (int g = 0; g < NG; ++g) {
          const GroupDev& gd = p.groups[g];
          if (gd.robot != r) continue;
          double x1, y1, ca, sa;
          mount_point(wsr, gd.mx, gd.my, x1, y1, ca, sa);
          org[wl * NG + g] = make_double2(x1, y1);
          const float cf = (float)ca, sf = (float)sa;
          GroupCull gc;
          gc.lo_x = cf * gd.lo_c - sf * gd.lo_s; gc.lo_y = sf * gd.lo_c + cf * gd.lo_s;
          gc.hi_x = cf * gd.hi_c - sf * gd.hi_s; gc.hi_y = sf * gd.hi_c + cf * gd.hi_s;
          gc.ax_x = cf * gd.ax_c - sf * gd.ax_s; gc.ax_y = sf * gd.ax_c + cf * gd.ax_s;
          gc.range2 = gd.range * gd.range;
          gc.has_cone = gd.has_cone;
          gc.ox = (float)x1;
          gc.oy = (float)y1;
          gc._pad[0] = gc._pad[1] = 0.f;
          gcull[wl * NG + g] = gc;
        }
      }
    }
  }
  BRS_PH(ht == 0, 12);
  // camera worlds with a few extra sensors: the helper team has time to spare, the workers
  // do not -- run the direct-path sensors of this tile here, on the poses just committed
  if (do_direct) {
    team_sync(bid, nth);
    sense_direct(p, ht, nth, bid, threadIdx.x & 31, w0, nw, rawb, segtab, wsb, meta, p.direct_on_helper == 2 ? 6 : 7);
  }
  BRS_PH(ht == 0, 13);
}

// Robot.step of one tile by the fused kernel's helper team (thread ht of nth): PREP, then
// the collision tests and the commit.
__device__ BRS_INL_ROBOT void robot_step_tile(const KParams& p, int w0, int nw, int ht, int nth, bool do_move,
                                                bool do_rays, bool do_sense, const uint32_t* __restrict__ rawb,
                                                float4* __restrict__ segtab, double* __restrict__ wsb,
                                                int4* __restrict__ cflag, int2* __restrict__ meta,
                                                double2* __restrict__ org, GroupCull* __restrict__ gcull,
                                                int* __restrict__ cq, uint32_t bar, uint32_t parity, bool staged) {
  BRS_PH_DECL;
  if (ht == 0) cq[0] = 0;  // candidate queue of the collision phase
  prepare_tile(p, w0, nw, ht, nth, do_move, rawb, segtab, wsb, cflag, meta, bar, parity, staged);
  team_sync(2, nth);
  BRS_PH(ht == 0, 8);
  collide_commit_tile(p, w0, nw, ht, nth, 2, do_move, do_rays, do_sense && p.direct_on_helper, rawb, segtab, wsb, cflag,
                      meta, org, gcull, cq);
}

// TABLE, SENSE and PAINT of one tile by a team of NW threads (index tid, barrier id 1).
// `pm` is the thread's paint mapping.
struct PaintMap {
  int q, rslot, islot, nislot, RS;
};
__device__ __forceinline__ PaintMap make_paint_map(const KParams& p, int tid, int NW) {
  // thread -> (picture slot, row slot, column quad); pictures wider than 4 x NW: several quads per thread
  PaintMap m;
  const int pQ = max(p.cam_w >> 2, 1), pQc = min(pQ, NW);
  const int rest = tid / pQc;
  m.q = tid % pQc;
  m.RS = p.paint_rs;
  m.rslot = rest % m.RS;
  m.islot = rest / m.RS;
  m.nislot = (NW / pQc) / m.RS;
  return m;
}
template <int NR>
__device__ __forceinline__ void sense_paint_tile(const KParams& p, const int tid, const int NW, const int lane,
                                                 const int w0, const int nw, const bool do_sense, const bool do_cam,
                                                 const bool direct_here, const uint32_t* __restrict__ rawb,
                                                 const float4* __restrict__ segtab, const double* __restrict__ wsb,
                                                 const int2* __restrict__ meta, const double2* __restrict__ org,
                                                 const GroupCull* __restrict__ gcull, float4* __restrict__ gtab,
                                                 uint16_t* __restrict__ gidx, int2* __restrict__ gcnt,
                                                 ColInfo* __restrict__ colinfo, Strip* __restrict__ strips,
                                                 const double2* __restrict__ camcs, const uint32_t* __restrict__ rowsky,
                                                 const uint32_t* __restrict__ rowgnd, const PaintMap pm) {
  const int R = p.n_robots, SC = p.seg_cap, NG = p.n_group, nvis = SC + 4 * R;
  const int strips_per_col = 4 * (R - 1);
  const int p_q = pm.q, p_rslot = pm.rslot, p_islot = pm.islot, p_nislot = pm.nislot, p_RS = pm.RS;
  BRS_PH_DECL;
      // ---------------- TABLE ------------------------------------------------------------
      // per (world, group, visible segment): m = (ey, -ex)/na, q = d/na with d = O - P,
      // na = ex*dy - ey*dx.  Entries no ray of the group can hit are dropped: the own
      // robot's edges, degenerate ones (origin on the wall's line), walls whose line is
      // out of the group's range, walls entirely outside the cone of its rays (all three
      // conservative, with margins far above FP32 error).  Survivors are then compacted
      // in index order by one warp per (world, group), so the search loops over fewer
      // entries and ties still resolve as in the reference.
      if (NG > 0 && p.fused_table) {
        // enough (world, group) pairs to occupy the warps: build and compact in one pass
        const int warp = tid >> 5, nwarp = NW >> 5;
        for (int pg = warp; pg < nw * NG; pg += nwarp) {
          const int wl = (int)fdiv(pg, p.dNG), g = pg - wl * NG;
          const int S = meta[wl].x, nv = S + 4 * R, own = S + 4 * p.groups[g].robot;
          const GroupCull gc = gcull[pg];
          const float4* st = segtab + (size_t)wl * nvis;
          float4* tab = gtab + (size_t)pg * nvis;
          uint16_t* ix = gidx + (size_t)pg * nvis;
          int base = 0, cw = 0;
          for (int j0 = 0; j0 < nv; j0 += 32) {
            const int j = j0 + lane;
            float4 t;
            bool keep = false;
            if (j < nv && (unsigned)(j - own) >= 4u) keep = table_entry(st[j], gc, t);
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            const int pos = base + __popc(bal & ((1u << lane) - 1u));
            if (keep) {
              tab[pos] = t;
              ix[pos] = (uint16_t)j;
            }
            const int nwall = S - j0;  // lanes below this index hold walls
            cw += __popc(nwall >= 32 ? bal : (nwall <= 0 ? 0u : (bal & ((1u << nwall) - 1u))));
            base += __popc(bal);
          }
          if (lane == 0) {
            gcnt[pg] = make_int2(cw, base);
            if (p.counters)
              atomicAdd(&p.counters[0], (unsigned long long)cw * (do_cam ? p.groups[g].cam_rays : 0) +
                                            (unsigned long long)base * (do_sense ? p.groups[g].range_rays : 0));
          }
        }
        team_sync(1, NW);
      } else if (NG > 0) {
        // few pairs with many segments: all threads build, then one warp per pair compacts
        const int per_world = NG * nvis;
        for (int i = tid; i < nw * per_world; i += NW) {
          const int wl = (int)fdiv(i, p.dTab), k = i - wl * per_world;
          const int g = (int)fdiv(k, p.dNvis), j = k - g * nvis;
          const int S = meta[wl].x, nv = S + 4 * R, own = S + 4 * p.groups[g].robot;
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
          if (j < nv && (unsigned)(j - own) >= 4u) {
            float4 u;
            if (table_entry(segtab[(size_t)wl * nvis + j], gcull[wl * NG + g], u)) t = u;
          }
          gtab[(size_t)wl * per_world + k] = t;
        }
        team_sync(1, NW);
        // order-preserving in-place compaction: position <= index, chunks ascend
        {
          const int warp = tid >> 5, nwarp = NW >> 5;
          for (int pg = warp; pg < nw * NG; pg += nwarp) {
            const int wl = (int)fdiv(pg, p.dNG);
            const int S = meta[wl].x, nv = S + 4 * R;
            float4* tab = gtab + (size_t)pg * nvis;
            uint16_t* ix = gidx + (size_t)pg * nvis;
            int base = 0, cw = 0;
            for (int j0 = 0; j0 < nv; j0 += 32) {
              const int j = j0 + lane;
              float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
              if (j < nv) t = tab[j];
              const bool keep = (t.x != 0.f) || (t.y != 0.f);
              const unsigned bal = __ballot_sync(0xffffffffu, keep);
              const int pos = base + __popc(bal & ((1u << lane) - 1u));
              __syncwarp();
              if (keep) {
                tab[pos] = t;
                ix[pos] = (uint16_t)j;
              }
              const int nwall = S - j0;  // lanes below this index hold walls
              cw += __popc(nwall >= 32 ? bal : (nwall <= 0 ? 0u : (bal & ((1u << nwall) - 1u))));
              base += __popc(bal);
            }
            if (lane == 0) {
              gcnt[pg] = make_int2(cw, base);
              const GroupDev& gd = p.groups[pg - wl * NG];
              if (p.counters)
                atomicAdd(&p.counters[0], (unsigned long long)cw * (do_cam ? gd.cam_rays : 0) +
                                              (unsigned long long)base * (do_sense ? gd.range_rays : 0));
            }
          }
        }
        team_sync(1, NW);
      }

      BRS_PH(tid == 0, 1);
      // ---------------- SENSE --------------------------------------------------------------
      // 5a. group path: camera column blocks, then grouped range sensors
      {
        const int SL = p.SLG, g = tid & (SL - 1), slot = tid / SL, nslot = NW / SL;
        const uint32_t gmask = (SL == 32) ? 0xffffffffu : (((1u << SL) - 1u) << (lane & ~(SL - 1)));
        const int cblocks = do_cam ? p.n_camera * (p.cam_w / NR) : 0;
        const int ngr = do_sense ? p.n_gr : 0;
        const int ntask = cblocks + ngr;
        for (int i = slot; i < nw * ntask; i += nslot) {
          const int wl = (int)fdiv(i, p.dGT), task = i - wl * ntask, w = w0 + wl;
          const int S = meta[wl].x;
          const uint32_t* raw = rawb + (size_t)wl * 5 * SC;
          const double* ws = wsb + (size_t)wl * R * BRS_WS_STRIDE;
          if (task < cblocks) {
            const int cpb = p.cam_w / NR;  // column blocks per camera
            const int c = (int)fdiv(task, p.dCpb), col0 = (task - c * cpb) * NR;
            const CameraDev& cd = p.cameras[c];
            const int r = cd.robot;
            const double2 O = org[wl * NG + cd.group];
            const double ca = ws[r * BRS_WS_STRIDE + BRS_WS_RCS], sa = ws[r * BRS_WS_STRIDE + BRS_WS_RCS + 1];
            double x2[NR], y2[NR];
            float rx[NR], ry[NR];
            uint32_t key[NR];
            int idx[NR];
#pragma unroll
            for (int k = 0; k < NR; ++k) {
              const double2 cs = camcs[c * p.cam_w + col0 + k];
              const double dxu = ds(dm(ca, cs.x), dm(sa, cs.y));  // cos(a + theta_i)
              const double dyu = da(dm(sa, cs.x), dm(ca, cs.y));  // sin(a + theta_i)
              x2[k] = da(dm(dxu, cd.max_range), O.x);
              y2[k] = da(dm(dyu, cd.max_range), O.y);
              rx[k] = (float)ds(x2[k], O.x);
              ry[k] = (float)ds(y2[k], O.y);
              key[k] = BRS_GKEY_INIT;
              idx[k] = -1;
            }
            // walls only (height == 1); other robots become strips below
            const int pg = wl * NG + cd.group;
            const float4* tab = gtab + (size_t)pg * nvis;
            const int cw = gcnt[pg].x;
            if (SL == 1)
              search_group<NR, true>(tab, 0, cw, 0, 1, rx, ry, key, idx);
            else
              search_group<NR, false>(tab, 0, cw, g, SL, rx, ry, key, idx);
            const double wsize = (double)__int_as_float(meta[wl].y);
#pragma unroll
            for (int k = 0; k < NR; ++k) {
              if (SL > 1) group_max(key[k], idx[k], SL, gmask);
              if (g == 0) {
                const size_t col = ((size_t)wl * p.n_camera + c) * p.cam_w + col0 + k;
                if (idx[k] >= 0) idx[k] = gidx[(size_t)pg * nvis + idx[k]];
                camera_column(p, cd, r, idx[k], raw, SC, S, ws, wsize, O.x, O.y, x2[k], y2[k], colinfo + col,
                              strips + col * strips_per_col);
              }
            }
          } else {
            // RangeSensor.update on a shared-origin mount
            const int si = p.gr_idx[task - cblocks];
            const RangeDev& sd = p.ranges[si];
            const int r = sd.robot;
            const double2 O = org[wl * NG + sd.group];
            const double ca = ws[r * BRS_WS_STRIDE + BRS_WS_RCS], sa = ws[r * BRS_WS_STRIDE + BRS_WS_RCS + 1];
            const int pg = wl * NG + sd.group;
            const float4* tab = gtab + (size_t)pg * nvis;
            const int ct = gcnt[pg].y;
            double best = sd.max;  // the reading stays at max unless a ray hits nearer
            for (int k = 0; k < sd.n_rays; ++k) {
              const double dxu = ds(dm(ca, sd.cr[k]), dm(sa, sd.sr[k]));
              const double dyu = da(dm(sa, sd.cr[k]), dm(ca, sd.sr[k]));
              const double x2 = da(dm(dxu, sd.max), O.x), y2 = da(dm(dyu, sd.max), O.y);
              float rx[1] = {(float)ds(x2, O.x)}, ry[1] = {(float)ds(y2, O.y)};
              uint32_t key[1] = {BRS_GKEY_INIT};
              int idx[1] = {-1};
              if (SL == 1)
                search_group<1, true>(tab, 0, ct, 0, 1, rx, ry, key, idx);
              else
                search_group<1, false>(tab, 0, ct, g, SL, rx, ry, key, idx);
              if (SL > 1) group_max(key[0], idx[0], SL, gmask);
              if (g == 0 && idx[0] >= 0) best = fmin(best, refine64(raw, SC, S, ws, gidx[(size_t)pg * nvis + idx[0]], O.x, O.y, x2, y2, sd.max));
            }
            if (g == 0) p.range_out[(size_t)w * p.n_range + si] = range_reading(p, w, si, (float)best, (float)sd.max);
          }
        }
      }
      // 5b. direct path (unless the helper team already ran it for this tile)
      if (do_sense && direct_here) sense_direct(p, tid, NW, 1, lane, w0, nw, rawb, segtab, wsb, meta, p.direct_on_helper == 2 ? 1 : 7);

      BRS_PH(tid == 0, 2);
      // ---------------- PAINT ----------------------------------------------------------------
      if (do_cam) {
        team_sync(1, NW);
      const int W = p.cam_w, H = p.cam_h, Q = W >> 2;
      const int nimg = nw * p.n_camera;
      // aligned lane octets hold 8 adjacent quads of the same rows of one picture: the quad
      // index of lane l is (tid % min(Q, NW)), so this needs Q (and NW) to be multiples of 8
      const bool paint_oct = (Q & 7) == 0 && (NW & 7) == 0;
      uint4* const out0 = (uint4*)(p.cam_out + (size_t)w0 * p.n_camera * (size_t)H * W * 4);
      // worker thread -> (picture slot, row slot, column quad): the 4 column records stay
      // in registers over the rows, a warp writes whole 128-byte lines per store
      if (p_islot < p_nislot) {
        for (int im = p_islot; im < nimg; im += p_nislot)
          for (int q = p_q; q < Q; q += NW)
            paint_rows(out0 + (size_t)im * H * Q, colinfo + (size_t)im * W, strips + (size_t)im * W * strips_per_col,
                       strips_per_col, rowsky, rowgnd, p.flat_rows != 0, paint_oct, lane, q, Q, p_rslot, p_RS, H);
      }
      }
  BRS_PH(tid == 0, 3);
}

template <int NR, bool BIG>
__global__ void __launch_bounds__(BIG ? BRS_BIG_THREADS : BRS_SMALL_THREADS, BIG ? BRS_BIG_CTAS : BRS_SMALL_CTAS)
    brs_step_kernel(const __grid_constant__ KParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31;
  const int NW = p.NW;              // worker threads
  const int NHT = NT - NW;          // helper team (>= 32)
  const bool is_helper = tid >= NW;
  const int R = p.n_robots, SC = p.seg_cap, TW = p.TW, NG = p.n_group;
  const int nvis = SC + 4 * R;  // table stride per world: walls, then 4 edges per robot at S + 4*o + e
  const SmemLayout L = make_layout(TW, SC, R, NG, p.n_camera, p.cam_w, p.cam_h, !BIG);
  float4* const gtab = (float4*)(smem + L.gtab);
  uint16_t* const gidx = (uint16_t*)(smem + L.gidx);
  int2* const gcnt = (int2*)(smem + L.gcnt);
  ColInfo* const colinfo = (ColInfo*)(smem + L.colinfo);
  Strip* const strips = (Strip*)(smem + L.strips);
  double2* const camcs = (double2*)(smem + L.camcs);
  uint32_t* const rowsky = (uint32_t*)(smem + L.rowsky);
  uint32_t* const rowgnd = (uint32_t*)(smem + L.rowgnd);
  uint64_t* const mbar = (uint64_t*)(smem + L.mbar);
  volatile int* const sched = (volatile int*)(smem + L.sched);
  const uint32_t world_bytes = (uint32_t)(5 * SC * 4);
  const bool do_move = p.flags & 1u, do_sense = p.flags & 2u, do_cam = (p.flags & 4u) && p.n_camera > 0;
  const bool do_rays = do_sense || do_cam;

  // ---- prologue: the helper initialises the TMA barriers and goes straight to the first
  // tile; the workers meanwhile load the camera tables (iteration -1 gives them nothing
  // else to do), both meet at that iteration's barrier
  if (tid == NW) {
    mbar_init(smem_u32(&mbar[0]), 1);
    mbar_init(smem_u32(&mbar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (!is_helper) {
    for (int i = tid; i < p.n_camera * p.cam_w; i += NW) camcs[i] = p.cam_cs[i];
    for (int i = tid; i < p.cam_h; i += NW) {
      rowsky[i] = p.row_sky[i];
      rowgnd[i] = p.row_gnd[i];
    }
  }

  const PaintMap pmap = make_paint_map(p, tid, NW);
  // launched behind brs_prep_kernel (programmatic dependent launch): nothing below may pass it
  // p.pdl: launched with programmatic stream serialization behind whatever precedes it in the
  // stream (the previous step's kernel, a caller's kernel that wrote the actions): the prologue
  // above -- barrier init, the camera tables, static data only -- overlaps that kernel's tail and
  // this grid's own launch latency; everything below waits for its full completion.  The next
  // launch may be set up as soon as every CTA of this one is past this point.
  if (p.prepped || p.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (p.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  BRS_PH_DECL;
  // iteration -1 only lets the helper team step the CTA's first tile
  int tile = 0;
  int h_prev1 = -1, h_prev2 = -1;  // brs_steps (scheduler thread): items of the last two iterations, to be published
  for (int it = -1;; ++it) {
    const int buf = it & 1;
    const int w0 = tile * TW, nw = min(TW, p.n_worlds - w0);
    if (is_helper) {
      // ================= helper team: Robot.step of the NEXT tile =====================
      if (tid == NW) {
        // next tile index + TMA of its wall store into the other buffer (free since the
        // barrier that ended the previous iteration)
        // (the first tile of a CTA is its block index; the counter hands out the rest)
        int tn;
        if (p.n_steps > 1) {
          // brs_steps: the counter hands out n_steps * n_tiles items in (step, tile) order.  Step s
          // of a tile may run on another CTA than its step s-1, so two epoch words per tile order
          // them: `committed` (poses, velocities, stall flags of the tile are those after step s)
          // gates the helper, `painted` (every observation of step s is written) gates the
          // workers' stores.  Both are published here, by this thread, after the barrier that
          // ended the iteration which produced them.  A waiter's item is always later in item
          // order than what it waits for, and every CTA of the grid is resident: no deadlock.
          // (a batch of no more tiles than CTAs: CTA b keeps tile b for every step -- program
          // order does the rest, no counter, no epochs; config 1 as written is this case)
          unsigned* const ep = p.epoch;
          if (p.n_tiles <= (int)gridDim.x) {
            tn = ((int)blockIdx.x < p.n_tiles && it + 1 < p.n_steps) ? (int)blockIdx.x : p.n_tiles;
            sched[3 + (buf ^ 1)] = 1;
          } else {
            if (h_prev1 >= 0)
              st_release_gpu(ep + 2 * (h_prev1 % p.n_tiles), p.epoch0 + (unsigned)(h_prev1 / p.n_tiles) + 1u);
            if (h_prev2 >= 0)
              st_release_gpu(ep + 2 * (h_prev2 % p.n_tiles) + 1, p.epoch0 + (unsigned)(h_prev2 / p.n_tiles) + 1u);
            // (every item comes from the counter, the first one too: a CTA that is not resident yet owns nothing)
            const long long k = (long long)atomicAdd(&p.sched[0], 1u);
            int ok = 1;
            unsigned want = 0;
            h_prev2 = h_prev1;
            if (k < (long long)p.n_steps * p.n_tiles) {
              tn = (int)(k % p.n_tiles);
              const unsigned sk = (unsigned)(k / p.n_tiles);
              if (sk > 0) {
                want = p.epoch0 + sk;
                while ((int)(ld_relaxed_gpu(ep + 2 * tn) - want) < 0) __nanosleep(64);
                ok = (int)(ld_relaxed_gpu(ep + 2 * tn + 1) - want) >= 0;
              }
              h_prev1 = (int)k;
            } else {
              tn = p.n_tiles;
              h_prev1 = -1;
            }
            sched[3 + (buf ^ 1)] = ok;
            sched[5 + (buf ^ 1)] = (int)want;
          }
        } else {
          tn = it < 0 ? (int)blockIdx.x : (int)(gridDim.x + atomicAdd(&p.sched[0], 1u));
        }
        sched[1 + (buf ^ 1)] = tn;
        if ((!BIG || p.prepped) && tn < p.n_tiles) {
          const int wn = tn * TW, nn = min(TW, p.n_worlds - wn);
          const uint32_t robot_bytes = p.prepped ? (uint32_t)(nn * R * BRS_WS_STRIDE * 8) : 0u;
          mbar_expect_tx(smem_u32(&mbar[buf ^ 1]), (BIG ? 0u : nn * world_bytes) + robot_bytes);
          if (!BIG)
            tma_bulk_g2s(smem_u32(smem + L.raw[buf ^ 1]), p.seg + (size_t)wn * 5 * SC, nn * world_bytes,
                         smem_u32(&mbar[buf ^ 1]));
          if (p.prepped)
            tma_bulk_g2s(smem_u32(smem + L.ws[buf ^ 1]), p.wsg + (size_t)wn * R * BRS_WS_STRIDE, robot_bytes,
                         smem_u32(&mbar[buf ^ 1]));
        }
      }
      team_sync(2, NHT);
      const int tn = sched[1 + (buf ^ 1)];
      if (tn < p.n_tiles) {
        const int wn = tn * TW, nn = min(TW, p.n_worlds - wn);
        // big worlds (one or two CTAs per SM by their tables alone) do not stage the wall
        // store: the helper reads it once from L2 / HBM, coalesced, to build the FP32 table
        const uint32_t* rawn = BIG ? p.seg + (size_t)wn * 5 * SC : (const uint32_t*)(smem + L.raw[buf ^ 1]);
        robot_step_tile(p, wn, nn, tid - NW, NHT, do_move, do_rays, do_sense, rawn,
                        (float4*)(smem + L.segtab[buf ^ 1]), (double*)(smem + L.ws[buf ^ 1]),
                        (int4*)(smem + L.cflag[buf ^ 1]), (int2*)(smem + L.meta[buf ^ 1]),
                        (double2*)(smem + L.org[buf ^ 1]), (GroupCull*)(smem + L.gcull[buf ^ 1]),
                        (int*)(smem + L.cq), smem_u32(&mbar[buf ^ 1]), (uint32_t)((it + 1) >> 1) & 1u, !BIG);
      }
    } else if (do_rays && it >= 0) {
      // ================= workers: sensors and pictures of THIS tile ====================
      const uint32_t* const rawb = BIG ? p.seg + (size_t)w0 * 5 * SC : (const uint32_t*)(smem + L.raw[buf]);
      const float4* const segtab = (const float4*)(smem + L.segtab[buf]);
      const double* const wsb = (const double*)(smem + L.ws[buf]);
      const int2* const meta = (const int2*)(smem + L.meta[buf]);
      const double2* const org = (const double2*)(smem + L.org[buf]);
      const GroupCull* const gcull = (const GroupCull*)(smem + L.gcull[buf]);
      // every thread observes the phase of this tile's wall store itself before reading it
      if (!BIG || p.prepped) mbar_wait(smem_u32(&mbar[buf]), (uint32_t)(it >> 1) & 1u);
      if (p.n_steps > 1 && !sched[3 + buf]) {
        // brs_steps: the previous step's observations of this tile were still being written by
        // another CTA when the item was fetched (rare: the same tile comes up n_tiles items later)
        if (tid == 0) {
          const unsigned want = (unsigned)sched[5 + buf];
          while ((int)(ld_relaxed_gpu(p.epoch + 2 * tile + 1) - want) < 0) __nanosleep(64);
        }
        team_sync(1, NW);
      }
      BRS_PH(tid == 0, 0);

      sense_paint_tile<NR>(p, tid, NW, lane, w0, nw, do_sense, do_cam, p.direct_on_helper != 1, rawb, segtab, wsb, meta, org,
                           gcull, gtab, gidx, gcnt, colinfo, strips, camcs, rowsky, rowgnd, pmap);
    }
    BRS_PH(tid == NW, 14);
    __syncthreads();  // the teams meet: this tile is done, the next one is stepped
    tile = sched[1 + (buf ^ 1)];
    if (tile >= p.n_tiles) break;
    BRS_PH(tid == 0, 4);    // (the load of `tile` is the first consumer of the barrier: its wait lands here)
    BRS_PH(tid == NW, 15);
  }

  // brs_steps: the observations of the CTA's last item (its poses were published at the top of
  // the last iteration); the scheduler thread holds the item in h_prev2 after fetching the sentinel
  if (p.n_steps > 1 && tid == NW && h_prev2 >= 0 && p.n_tiles > (int)gridDim.x)
    st_release_gpu(p.epoch + 2 * (h_prev2 % p.n_tiles) + 1, p.epoch0 + (unsigned)(h_prev2 / p.n_tiles) + 1u);
  // self-resetting scheduler: the last CTA to leave zeroes both counters for the next launch
  if (tid == 0) {
    __threadfence();
    const unsigned done = atomicAdd(&p.sched[1], 1u);
    if (done == gridDim.x - 1) {
      p.sched[0] = 0u;
      p.sched[1] = 0u;
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------------------
// FP32 FFMA throughput microbenchmark (roofline denominator)
// ------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) brs_ffma_kernel(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f;
  float x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  const float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 123.456f) out[0] = s;  // keeps the chain alive, practically never taken
}

// ------------------------------------------------------------------------------
// HBM write-bandwidth microbenchmark: every thread streams 16-byte stores over a
// buffer, grid-stride (what PAINT does, without the work).  mode 0: st.global.cs
// (the kernel's policy), 1: default policy.  Roofline context for picture-bound worlds.
// ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) brs_fill_kernel(uint4* out, size_t n16, int mode) {
  const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 0xff00ff00u, 0xffffffffu);
  if (mode == 3) {
    // CTA-contiguous: a CTA streams whole 128 KB chunks (one config-2 picture), like PAINT does
    const size_t per = 8192, nchunk = n16 / per;
    for (size_t c = blockIdx.x; c < nchunk; c += gridDim.x) {
      uint4* o = out + c * per;
#pragma unroll 4
      for (int i = threadIdx.x; i < (int)per; i += 256) __stcs(o + i, v);
    }
    return;
  }
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    if (mode == 0)
      __stcs(out + i, v);
    else
      out[i] = v;
  }
}

#endif  // __CUDACC__
}  // namespace brs

// ------------------------------------------------------------------------------
// Top-down renderer (upstream: the World's own picture of itself --
// world.py display / take_picture -- and devices/cameras.py GroundCamera, the downward-looking
// camera under a robot).  Not on the step path: a separate launch that paints the CURRENT state
// of a range of worlds from above -- ground colour, bulbs as discs, walls as lines in list
// order, robots as filled boxes in list order -- into RGBA images.
//   pixel centre c = (px + 0.5, py + 0.5) looks at world point  o + c.x * u + c.y * v:
//     whole-world view:   o = 0, u = (W / img_w, 0), v = (0, H / img_h)
//     robot view (GroundCamera, robot >= 0): a view_w x view_h window centred on the robot,
//       row 0 towards its front, the robot itself not drawn
//   a wall covers the point when it is within `half` world units of the segment;
//   a robot covers it when the point is inside its (committed) bounding box.
// One CTA per (world, band of rows); the world's walls are staged in shared memory.  All pixel
// arithmetic is explicit IEEE float32 (no contraction), restated in oracle/topdown.py.
// ------------------------------------------------------------------------------
namespace brs {
#ifdef __CUDACC__
struct RenderParams {
  int first_world, n_worlds, img_w, img_h, bands, seg_cap, n_robots, n_bulbs;
  int robot;             // < 0: whole world; else the robot the window is centred on
  float view_w, view_h;  // robot view: window size in world units
  float half_px;         // half line thickness, in pixels
  float bulb_px;         // bulb disc radius, in pixels
  uint32_t ground_rgba, bulb_rgba;
  const uint32_t* seg;
  const int* seg_count;
  const float* world_size;
  const float* bulbs;
  const double* pose;
  const RobotDev* robots;
  uint32_t* out;
};
__global__ void __launch_bounds__(256) brs_render_kernel(const RenderParams q) {
  extern __shared__ __align__(16) unsigned char rsm[];
  float4* const walls = (float4*)rsm;                          // x1, y1, ex, ey
  uint32_t* const wcol = (uint32_t*)(rsm + (size_t)q.seg_cap * 16);
  float* const boxes = (float*)(wcol + q.seg_cap);             // per robot: 4 corners (x, y) + rgba bits
  const int wl = blockIdx.x / q.bands, band = blockIdx.x - wl * q.bands, w = q.first_world + wl;
  const int SC = q.seg_cap, S = min(max(q.seg_count[w], 0), SC);
  const uint32_t* segw = q.seg + (size_t)w * 5 * SC;
  for (int j = threadIdx.x; j < S; j += blockDim.x) {
    const float x1 = __uint_as_float(segw[j]), y1 = __uint_as_float(segw[SC + j]);
    walls[j] = make_float4(x1, y1, __fsub_rn(__uint_as_float(segw[2 * SC + j]), x1), __fsub_rn(__uint_as_float(segw[3 * SC + j]), y1));
    wcol[j] = segw[4 * SC + j] | 0xff000000u;
  }
  for (int r = threadIdx.x; r < q.n_robots; r += blockDim.x) {
    const double* ps = q.pose + ((size_t)w * q.n_robots + r) * 3;
    const RobotDev& rd = q.robots[r];
    double sn, cs;
    sincos(ps[2], &sn, &cs);
    for (int k = 0; k < 4; ++k) {
      boxes[r * 9 + 2 * k] = (float)__dadd_rn(ps[0], __dsub_rn(__dmul_rn(cs, rd.cox[k]), __dmul_rn(sn, rd.coy[k])));
      boxes[r * 9 + 2 * k + 1] = (float)__dadd_rn(ps[1], __dadd_rn(__dmul_rn(sn, rd.cox[k]), __dmul_rn(cs, rd.coy[k])));
    }
    boxes[r * 9 + 8] = __uint_as_float(rd.rgba | 0xff000000u);
  }
  __syncthreads();
  float ox = 0.f, oy = 0.f, ux, uy = 0.f, vx = 0.f, vy;
  if (q.robot < 0) {
    ux = __fdiv_rn(q.world_size[2 * w], (float)q.img_w);
    vy = __fdiv_rn(q.world_size[2 * w + 1], (float)q.img_h);
  } else {
    const double* ps = q.pose + ((size_t)w * q.n_robots + q.robot) * 3;
    double sn, cs;
    sincos(ps[2], &sn, &cs);
    const double pu = (double)q.view_w / (double)q.img_w, pv = (double)q.view_h / (double)q.img_h;
    const double fw = __dmul_rn(0.5 * (double)q.img_h, pv), lt = __dmul_rn(0.5 * (double)q.img_w, pu);
    ox = (float)__dadd_rn(ps[0], __dadd_rn(__dmul_rn(cs, fw), __dmul_rn(sn, lt)));
    oy = (float)__dadd_rn(ps[1], __dsub_rn(__dmul_rn(sn, fw), __dmul_rn(cs, lt)));
    ux = (float)(-__dmul_rn(sn, pu));
    uy = (float)__dmul_rn(cs, pu);
    vx = (float)(-__dmul_rn(cs, pv));
    vy = (float)(-__dmul_rn(sn, pv));
  }
  const float unit = q.robot < 0 ? fmaxf(ux, vy) : fmaxf(__fdiv_rn(q.view_w, (float)q.img_w), __fdiv_rn(q.view_h, (float)q.img_h));
  const float half = __fmul_rn(q.half_px, unit), half2 = __fmul_rn(half, half);
  const float br = __fmul_rn(q.bulb_px, unit), br2 = __fmul_rn(br, br);
  const int rows = (q.img_h + q.bands - 1) / q.bands, r0 = band * rows, r1 = min(r0 + rows, q.img_h);
  uint32_t* img = q.out + (size_t)wl * q.img_h * q.img_w;
  for (int i = r0 * q.img_w + threadIdx.x; i < r1 * q.img_w; i += blockDim.x) {
    const int py = i / q.img_w, px = i - py * q.img_w;
    const float cx = (float)px + 0.5f, cy = (float)py + 0.5f;
    const float x = __fadd_rn(ox, __fadd_rn(__fmul_rn(cx, ux), __fmul_rn(cy, vx)));
    const float y = __fadd_rn(oy, __fadd_rn(__fmul_rn(cx, uy), __fmul_rn(cy, vy)));
    uint32_t c = q.ground_rgba;
    for (int b = 0; b < q.n_bulbs; ++b) {
      const float* bl = q.bulbs + ((size_t)w * q.n_bulbs + b) * 4;
      const float dx = __fsub_rn(x, bl[0]), dy = __fsub_rn(y, bl[1]);
      if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) <= br2) c = q.bulb_rgba;
    }
    for (int j = 0; j < S; ++j) {
      const float4 s = walls[j];
      const float dx = __fsub_rn(x, s.x), dy = __fsub_rn(y, s.y);
      const float ee = __fadd_rn(__fmul_rn(s.z, s.z), __fmul_rn(s.w, s.w));
      float h = ee > 0.f ? __fdiv_rn(__fadd_rn(__fmul_rn(dx, s.z), __fmul_rn(dy, s.w)), ee) : 0.f;
      h = fminf(fmaxf(h, 0.f), 1.f);
      const float wx = __fsub_rn(dx, __fmul_rn(h, s.z)), wy = __fsub_rn(dy, __fmul_rn(h, s.w));
      if (__fadd_rn(__fmul_rn(wx, wx), __fmul_rn(wy, wy)) <= half2) c = wcol[j];
    }
    for (int r = 0; r < q.n_robots; ++r) {
      if (r == q.robot) continue;
      const float* bx = boxes + r * 9;
      bool pos = true, neg = true;
      for (int k = 0; k < 4; ++k) {
        const int k2 = (k + 1) & 3;
        const float ex = __fsub_rn(bx[2 * k2], bx[2 * k]), ey = __fsub_rn(bx[2 * k2 + 1], bx[2 * k + 1]);
        const float cr = __fsub_rn(__fmul_rn(ex, __fsub_rn(y, bx[2 * k + 1])), __fmul_rn(ey, __fsub_rn(x, bx[2 * k])));
        pos = pos && cr >= 0.f;
        neg = neg && cr <= 0.f;
      }
      if (pos || neg) c = __float_as_uint(bx[8]);
    }
    img[i] = c;
  }
}
#endif
}  // namespace brs
