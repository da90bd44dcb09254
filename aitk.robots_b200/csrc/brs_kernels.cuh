// brs_kernels.cuh -- sm_100a kernels of the batched World.step engine.
//
// One persistent CTA steps one world at a time (grid = SMs x resident CTAs,
// worlds strided over the grid).  The CTA is warp-specialised and software
// pipelined across worlds:
//
//   helper warp (first warp of pool B)         pool A warps (camera columns)
//   ------------------------------------       ------------------------------
//   range / light / smell sensors of world w   camera rays of world w
//   PREPARE world w+1:                         (named barrier, pool A only)
//     wait its TMA-landed wall store           paint world w's pictures
//     robot state HBM -> smem, sincos,
//     proposed poses, box corners (FP64),
//     FP32 search table
//   ------------------------------ __syncthreads ------------------------------
//   all warps: collision tests of world w+1 (FP64), commit / stall
//
//   0. wall stores (SoA rows x1|y1|x2|y2|rgba) arrive in shared memory through
//      TMA bulk copies (cp.async.bulk + mbarrier, SASS UBLKCP), double buffered
//      and issued one world ahead of the prepare step;
//   1. Robot.step: FP64 pose integration and 4-edge x segment collision tests
//      (exact IEEE ops, no FMA contraction: decisions equal the float64
//      reference's);
//   2. ray-cast sensors: FP32 nearest-hit search (lanes stride over segments,
//      warp-shuffle min), then one FP64 refinement of the winning segment per
//      ray so distances and pixel rounding match float64;
//   3. Camera.take_picture: columns painted with coalesced 16-byte (4 x uchar4)
//      streaming stores.
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

namespace brs {

// ---- device-side descriptors (host-precomputed from the public brs_*_desc) ----
struct RobotDev {
  double cox[4], coy[4];  // bbox corner k in the a=0 frame: dist_k * cos/sin(atan2(-cx, cy) + pi/2)
  double height;
  double vx_ramp, vy_ramp, va_ramp;
  uint32_t rgba;
  uint32_t _pad;
};
struct RangeDev {
  int robot, n_rays;
  double mx, my;        // mount offset in the a=0 frame: dfc*cos(dir+pi/2), dfc*sin(dir+pi/2)
  double cr[3], sr[3];  // cos/sin(sensor.a - incr_k)
  double max;
};
struct CameraDev {
  int robot, width, height, _pad;
  double mx, my;
  double max_range, color_fade, size_fade;
};
struct PointDev {
  int robot, _pad;
  double mx, my;
};

struct KParams {
  int n_worlds, n_robots, seg_cap, n_bulbs, grid_w, grid_h;
  int n_range, n_camera, n_light, n_smell;
  int cam_w, cam_h;
  int TA, TB, SL;  // camera-column threads, other-sensor threads, lanes per other-sensor ray
  unsigned flags;
  double dt, smell_cell, light_mult;
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
  const double2* cam_cs;    // [n_camera][cam_w] cos/sin of the column angle i/W*fov - fov/2
  const uint32_t* row_sky;  // [cam_h] rgba per image row
  const uint32_t* row_gnd;  // [cam_h]
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

// per-world working state in shared memory, double buffered (prepared one world ahead)
//   doubles per robot: pose[3] vel[3] rcs[2] dbox[8] prop[16] = 32
#define BRS_WS_POSE 0
#define BRS_WS_VEL 3
#define BRS_WS_RCS 6
#define BRS_WS_DBOX 8
#define BRS_WS_PROP 16
#define BRS_WS_STRIDE 32

// shared memory carve-up, byte offsets; mirrored on the host for the launch size
struct SmemLayout {
  int raw[2], segtab[2], ws[2], meta[2], mbar, colinfo, strips, camcs, rowsky, rowgnd, total;
};
__host__ __device__ inline int align_up(int x, int a) { return (x + a - 1) / a * a; }
__host__ __device__ inline SmemLayout make_layout(int seg_cap, int R, int n_camera, int cam_w, int cam_h) {
  SmemLayout L;
  int o = 0;
  for (int b = 0; b < 2; ++b) { L.raw[b] = o; o += 5 * seg_cap * 4; o = align_up(o, 128); }
  for (int b = 0; b < 2; ++b) { L.segtab[b] = o; o += (seg_cap + 4 * R) * 16; }
  for (int b = 0; b < 2; ++b) { L.ws[b] = o; o += R * BRS_WS_STRIDE * 8; }
  for (int b = 0; b < 2; ++b) { L.meta[b] = o; o += 16; }  // int S; float wsize
  L.mbar = o; o += 2 * 8;
  o = align_up(o, 16);
  L.colinfo = o; o += n_camera * cam_w * (int)sizeof(ColInfo);
  int strips_per_col = 4 * (R - 1);
  L.strips = o; o += n_camera * cam_w * strips_per_col * (int)sizeof(Strip);
  o = align_up(o, 16);
  L.camcs = o; o += n_camera * cam_w * 16;
  L.rowsky = o; o += cam_h * 4;
  L.rowgnd = o; o += cam_h * 4;
  L.total = align_up(o, 16);
  return L;
}

#ifdef __CUDACC__

#define BRS_PI_OVER_2 1.5707963267948966

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
  // ua = na/den in [0,1] and ub = nb/den in [0,1]; division kept for the ray
  // parameter (refinement) and for literal equality with the reference.
  const double ua = na / den, ub = nb / den;
  if (ua_out) *ua_out = ua;
  return !(ua < 0.0 || ua > 1.0 || ub < 0.0 || ub > 1.0);
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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

#define BRS_KEY_NOHIT 0x3F800001u  // just above 1.0f: parameters <= 1 are hits

// FP32 nearest-hit search of one ray against segtab[j0, j1), lane g of SL
// striding.  key = bits of the ray parameter t (unsigned order == float order
// for t >= 0; negative / NaN / inf parameters compare above BRS_KEY_NOHIT).
__device__ __forceinline__ void search32(const float4* __restrict__ segtab, int j0, int j1, int g, int SL, float ox,
                                         float oy, float rx, float ry, uint32_t& key, int& idx) {
#pragma unroll 4
  for (int j = j0 + g; j < j1; j += SL) {
    const float4 s = segtab[j];  // sx, sy, ex, ey
    const float dx = ox - s.x, dy = oy - s.y;
    const float den = s.w * rx - s.z * ry;
    const float na = s.z * dy - s.w * dx;
    const float nb = rx * dy - ry * dx;
    const float inv = rcp_approx(den);
    const uint32_t tb = __float_as_uint(na * inv);
    const uint32_t ub = __float_as_uint(nb * inv);
    if (ub <= 0x3F800000u && tb < key) {
      key = tb;
      idx = j;
    }
  }
}

// reduce (key, idx) over the SL lanes of a group (SL power of two, groups are
// aligned lane ranges, `mask` names exactly the group's lanes so that groups of
// one warp may sit in different branches); ties go to the larger index (the
// reference keeps the last of equally distant hits)
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

// segment j of the current world in float64: static walls from the TMA-landed
// store (float32 values are exact in float64), robot boxes from dbox.
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

// float64 distance of the hit of ray (x1,y1)->(x2,y2) on segment j, computed
// like cast_ray: pos = intersect_hit(...), dist = distance(pos, origin)
__device__ __forceinline__ double refine64(const uint32_t* raw, int SC, int S, const double* ws, int j, double x1,
                                           double y1, double x2, double y2) {
  double x3, y3, x4, y4, ua = 0.0;
  seg64(raw, SC, S, ws, j, x3, y3, x4, y4);
  isect64(x1, y1, x2, y2, x3, y3, x4, y4, &ua);
  const double px = da(x1, dm(ua, ds(x2, x1))), py = da(y1, dm(ua, ds(y2, y1)));
  const double ex = ds(px, x1), ey = ds(py, y1);
  return sqrt(da(dm(ex, ex), dm(ey, ey)));
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

// Paint rows r0, r0+rstep, ... of the 4 columns of quad q (upstream
// Camera.take_picture: sky above, shaded wall, ground below, then the strips of
// other robots back to front).
__device__ __forceinline__ void paint_rows(uint4* __restrict__ out, const ColInfo* __restrict__ cinf,
                                           const Strip* __restrict__ strips, int strips_per_col,
                                           const uint32_t* __restrict__ rowsky, const uint32_t* __restrict__ rowgnd,
                                           int q, int Q, int r0, int rstep, int H) {
  ColInfo ci[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) ci[k] = cinf[4 * q + k];
  // rare per-quad extras (robot strips, a column without any wall hit) take a slow path
  const bool slow = ((ci[0].n_robot | ci[1].n_robot | ci[2].n_robot | ci[3].n_robot) != 0) ||
                    ((ci[0].top | ci[1].top | ci[2].top | ci[3].top) < 0);
  uint4* o = out + (size_t)r0 * Q + q;
  const size_t ostep = (size_t)rstep * Q;
#pragma unroll 4
  for (int j = r0; j < H; j += rstep, o += ostep) {
    const uint32_t sky = rowsky[j], gnd = rowgnd[j];
    uint32_t px[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) px[k] = j < ci[k].top ? sky : (j < ci[k].bot ? ci[k].rgba : gnd);
    if (slow) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (ci[k].top < 0) px[k] = 0u;  // no wall hit: the column keeps the blank RGBA (0,0,0,0)
        const Strip* st = strips + (size_t)(4 * q + k) * strips_per_col;
        float best = 3.0e38f;
        for (int n = 0; n < ci[k].n_robot; ++n) {
          const Strip sp = st[n];
          // back-to-front painting == the nearest covering strip wins, the later
          // one among equally distant strips
          if (j >= sp.lo && j <= sp.hi && sp.dist <= best) {
            best = sp.dist;
            px[k] = sp.rgba;
          }
        }
      }
    }
    __stcs(o, make_uint4(px[0], px[1], px[2], px[3]));
  }
}

// PREPARE a world (run by the lanes of one warp): robot state HBM -> shared,
// heading trig, proposed poses and velocities, committed and proposed box
// corners (compute_boundingbox), FP32 search table of the static walls.  A
// robot's proposal depends on nothing but its own state, so all of this is off
// the collision-test critical path and runs one world ahead.
__device__ __forceinline__ void prepare_world(const KParams& p, int w, int lane, bool do_move, const uint32_t* raw,
                                              float4* segtab, double* ws, int* meta) {
  const int R = p.n_robots, SC = p.seg_cap;
  const int S = min(p.seg_count[w], SC);
  if (lane == 0) {
    meta[0] = S;
    ((float*)meta)[1] = fmaxf(p.world_size[2 * w], p.world_size[2 * w + 1]);
  }
  if (lane < R) {
    const int r = lane;
    const RobotDev& rd = p.robots[r];
    double* wsr = ws + r * BRS_WS_STRIDE;
    const size_t o = ((size_t)w * R + r) * 3;
    const double x = p.pose[o], y = p.pose[o + 1], a = p.pose[o + 2];
    const double v0 = p.vel[o], v1 = p.vel[o + 1], v2 = p.vel[o + 2];
    wsr[BRS_WS_POSE] = x; wsr[BRS_WS_POSE + 1] = y; wsr[BRS_WS_POSE + 2] = a;
    wsr[BRS_WS_VEL] = v0; wsr[BRS_WS_VEL + 1] = v1; wsr[BRS_WS_VEL + 2] = v2;
    double s0, c0;
    sincos(a, &s0, &c0);
    wsr[BRS_WS_RCS] = c0;
    wsr[BRS_WS_RCS + 1] = s0;
    double px = x, py = y, c1 = c0, s1 = s0;
    double* pr = wsr + BRS_WS_PROP;
    if (do_move) {
      const double* tv = p.tvel + o;
      const double vx = deltav(tv[0], v0, rd.vx_ramp, p.dt);
      const double vy = deltav(tv[1], v1, rd.vy_ramp, p.dt);
      const double va = deltav(tv[2], v2, rd.va_ramp, p.dt);
      const double pa = ds(a, dm(va, p.dt));
      const double ang = da(-pa, BRS_PI_OVER_2);
      double sn, cs;
      sincos(ang, &sn, &cs);
      // vx*sin + vy*cos*dt ; vx*cos - vy*sin*dt  (dt binds to the vy term)
      const double tvx = da(dm(vx, sn), dm(dm(vy, cs), p.dt));
      const double tvy = ds(dm(vx, cs), dm(dm(vy, sn), p.dt));
      px = da(x, tvx);
      py = da(y, tvy);
      c1 = sn;  // sin(pi/2 - pa) = cos(pa)
      s1 = cs;  // cos(pi/2 - pa) = sin(pa)
      pr[8] = px; pr[9] = py; pr[10] = pa;
      pr[11] = vx; pr[12] = vy; pr[13] = va;
      pr[14] = c1; pr[15] = s1;
    }
    // corner = centre + R(heading) * (corner offset in the a=0 frame)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double ox = rd.cox[k], oy = rd.coy[k];
      wsr[BRS_WS_DBOX + 2 * k] = da(x, ds(dm(c0, ox), dm(s0, oy)));
      wsr[BRS_WS_DBOX + 2 * k + 1] = da(y, da(dm(s0, ox), dm(c0, oy)));
      if (do_move) {
        pr[2 * k] = da(px, ds(dm(c1, ox), dm(s1, oy)));
        pr[2 * k + 1] = da(py, da(dm(s1, ox), dm(c1, oy)));
      }
    }
  }
  // derived FP32 search table for the static walls: start point and delta
  for (int j = lane; j < S; j += 32) {
    const float x1 = __uint_as_float(raw[j]), y1 = __uint_as_float(raw[SC + j]);
    const float x2 = __uint_as_float(raw[2 * SC + j]), y2 = __uint_as_float(raw[3 * SC + j]);
    segtab[j] = make_float4(x1, y1, x2 - x1, y2 - y1);
  }
  __syncwarp();
  if (!do_move && lane < 4 * R) {  // nobody moves: the boxes above are final
    const int r = lane >> 2, e = lane & 3, e2 = (e + 1) & 3;
    const double* box = ws + r * BRS_WS_STRIDE + BRS_WS_DBOX;
    const float x1 = (float)box[2 * e], y1 = (float)box[2 * e + 1];
    const float x2 = (float)box[2 * e2], y2 = (float)box[2 * e2 + 1];
    segtab[S + lane] = make_float4(x1, y1, x2 - x1, y2 - y1);
  }
}

// ------------------------------------------------------------------------------
// The fused World.step kernel
// ------------------------------------------------------------------------------
__global__ void __launch_bounds__(320, 3) brs_step_kernel(const __grid_constant__ KParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31;
  const int R = p.n_robots, SC = p.seg_cap;
  const SmemLayout L = make_layout(SC, R, p.n_camera, p.cam_w, p.cam_h);
  uint64_t* const mbar = (uint64_t*)(smem + L.mbar);
  ColInfo* const colinfo = (ColInfo*)(smem + L.colinfo);
  Strip* const strips = (Strip*)(smem + L.strips);
  double2* const camcs = (double2*)(smem + L.camcs);
  uint32_t* const rowsky = (uint32_t*)(smem + L.rowsky);
  uint32_t* const rowgnd = (uint32_t*)(smem + L.rowgnd);
  const int strips_per_col = 4 * (R - 1);
  const uint32_t rec_bytes = (uint32_t)(5 * SC * 4);
  const bool do_move = p.flags & 1u, do_sense = p.flags & 2u, do_cam = (p.flags & 4u) && p.n_camera > 0;
  const bool is_helper = (tid >= p.TA) && (tid < p.TA + 32);  // first warp of pool B
  const int G = gridDim.x;

  // ---- prologue: barriers, TMA for the first two worlds, camera tables, PREPARE world 0
  if (tid == 0) {
    mbar_init(smem_u32(&mbar[0]), 1);
    mbar_init(smem_u32(&mbar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int b = 0; b < 2; ++b) {
      const int wb = blockIdx.x + b * G;
      if (wb < p.n_worlds) {
        mbar_expect_tx(smem_u32(&mbar[b]), rec_bytes);
        tma_bulk_g2s(smem_u32(smem + L.raw[b]), p.seg + (size_t)wb * 5 * SC, rec_bytes, smem_u32(&mbar[b]));
      }
    }
  }
  for (int i = tid; i < p.n_camera * p.cam_w; i += T) camcs[i] = p.cam_cs[i];
  for (int i = tid; i < p.cam_h; i += T) {
    rowsky[i] = p.row_sky[i];
    rowgnd[i] = p.row_gnd[i];
  }
  __syncthreads();
  if (is_helper && (int)blockIdx.x < p.n_worlds) {
    mbar_wait(smem_u32(&mbar[0]), 0);
    prepare_world(p, blockIdx.x, lane, do_move, (const uint32_t*)(smem + L.raw[0]), (float4*)(smem + L.segtab[0]),
                  (double*)(smem + L.ws[0]), (int*)(smem + L.meta[0]));
  }
  __syncthreads();

  int it = 0;
  for (int w = blockIdx.x; w < p.n_worlds; w += G, ++it) {
    const int buf = it & 1;
    const uint32_t* raw = (const uint32_t*)(smem + L.raw[buf]);
    float4* const segtab = (float4*)(smem + L.segtab[buf]);
    double* const ws = (double*)(smem + L.ws[buf]);
    const int S = ((const int*)(smem + L.meta[buf]))[0];
    const double wsize = (double)((const float*)(smem + L.meta[buf]))[1];
    // the wall store of this world landed long ago (the helper waited on it);
    // every thread still observes the barrier phase itself before reading TMA data
    mbar_wait(smem_u32(&mbar[buf]), (uint32_t)(it >> 1) & 1u);
    // TMA for world it+1 into the other buffer (free since the end of iteration it-1);
    // iteration 0's was issued in the prologue
    if (tid == 0 && it >= 1) {
      const int wn = w + G;
      if (wn < p.n_worlds) {
        mbar_expect_tx(smem_u32(&mbar[buf ^ 1]), rec_bytes);
        tma_bulk_g2s(smem_u32(smem + L.raw[buf ^ 1]), p.seg + (size_t)wn * 5 * SC, rec_bytes,
                     smem_u32(&mbar[buf ^ 1]));
      }
    }

    // ---------------- phase 1: collision tests in list order ---------------------
    // robot i sees the already-moved boxes of the robots before it.  FP64,
    // decisions equal the float64 reference's.
    if (do_move) {
      for (int i = 0; i < R; ++i) {
        double* wsi = ws + i * BRS_WS_STRIDE;
        const double* pr = wsi + BRS_WS_PROP;
        int hit = 0;
        {
          double ex1[4], ey1[4], edx[4], edy[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int e2 = (e + 1) & 3;
            ex1[e] = pr[2 * e];
            ey1[e] = pr[2 * e + 1];
            edx[e] = ds(pr[2 * e2], pr[2 * e]);
            edy[e] = ds(pr[2 * e2 + 1], pr[2 * e + 1]);
          }
          const int nseg = S + 4 * R;
          for (int j = tid; j < nseg; j += T) {
            if (j >= S && ((j - S) >> 2) == i) continue;  // never collide with yourself
            double x3, y3, x4, y4;
            seg64(raw, SC, S, ws, j, x3, y3, x4, y4);
#pragma unroll
            for (int e = 0; e < 4; ++e) hit |= (int)isect64_bool(ex1[e], ey1[e], edx[e], edy[e], x3, y3, x4, y4);
          }
        }
        hit = __syncthreads_or(hit);
        // commit or stall; everything below reads stable sources (old box or proposal)
        if (tid < 4) {  // this robot's box as FP32 search segments for the sensors
          const int e = tid, e2 = (e + 1) & 3;
          const double* src = hit ? wsi + BRS_WS_DBOX : pr;
          const float x1 = (float)src[2 * e], y1 = (float)src[2 * e + 1];
          const float x2 = (float)src[2 * e2], y2 = (float)src[2 * e2 + 1];
          segtab[S + 4 * i + e] = make_float4(x1, y1, x2 - x1, y2 - y1);
        }
        if (tid < 3) {
          const size_t o = ((size_t)w * R + i) * 3 + tid;
          if (!hit) {
            p.pose[o] = pr[8 + tid];
            p.vel[o] = pr[11 + tid];
          } else {
            p.vel[o] = 0.0;
          }
        }
        if (tid == 0) p.stalled[(size_t)w * R + i] = (uint8_t)(hit ? 1 : 0);
        // (__syncthreads_or was a barrier: nobody is still testing against the old box)
        if (!hit) {
          if (tid < 8) wsi[BRS_WS_DBOX + tid] = pr[tid];
          if (tid < 3) wsi[BRS_WS_POSE + tid] = pr[8 + tid];
          if (tid == 8) { wsi[BRS_WS_RCS] = pr[14]; wsi[BRS_WS_RCS + 1] = pr[15]; }
        }
        __syncthreads();  // robot i+1's tests and the sensors read the boxes / poses
      }
    }

    // ---------------- phase 2: ray-cast sensors (+ PREPARE of the next world) -----
    if (tid < p.TA) {
      // ---- pool A: one camera column per thread (upstream Camera._update + the
      //      per-column part of take_picture), then the pictures
      if (do_cam) {
        const int ncols = p.n_camera * p.cam_w;
        for (int col = tid; col < ncols; col += p.TA) {
          const int c = col / p.cam_w;
          const CameraDev& cd = p.cameras[c];
          const int r = cd.robot;
          double x1, y1, ca, sa;
          mount_point(ws + r * BRS_WS_STRIDE, cd.mx, cd.my, x1, y1, ca, sa);
          const double2 cs = camcs[col];
          const double dxu = ds(dm(ca, cs.x), dm(sa, cs.y));  // cos(a + theta_i)
          const double dyu = da(dm(sa, cs.x), dm(ca, cs.y));  // sin(a + theta_i)
          const double x2 = da(dm(dxu, cd.max_range), x1), y2 = da(dm(dyu, cd.max_range), y1);
          const float ox = (float)x1, oy = (float)y1;
          const float rx = (float)ds(x2, x1), ry = (float)ds(y2, y1);
          uint32_t key = BRS_KEY_NOHIT;
          int idx = -1;
          search32(segtab, 0, S, 0, 1, ox, oy, rx, ry, key, idx);  // walls only (height == 1)
          ColInfo ci;
          ci.rgba = 0; ci.top = -1; ci.bot = -1; ci.n_robot = 0;
          const double H = (double)cd.height;
          if (idx >= 0) {
            const double dist = refine64(raw, SC, S, ws, idx, x1, y1, x2, y2);
            const double s = clamp01(ds(1.0, dm(dist / wsize, cd.size_fade)));
            const double sc = clamp01(ds(1.0, dm(dist / wsize, cd.color_fade)));
            ci.rgba = shade(raw[4 * SC + idx], sc);
            const double high = dm(ds(1.0, s), H);
            ci.top = (int)ceil(high / 2.0);
            ci.bot = (int)ceil(ds(H, high / 2.0));
          }
          // other robots: every hit is a strip (float64 tests, few segments)
          Strip* st = strips + (size_t)col * strips_per_col;
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
          colinfo[col] = ci;
        }
        // ---------------- phase 3: Camera.take_picture (pool A only) ---------------
        named_bar_sync(1, p.TA);
        const int W = p.cam_w, H = p.cam_h, Q = W >> 2;
        for (int c = 0; c < p.n_camera; ++c) {
          uint4* out = (uint4*)(p.cam_out + ((size_t)w * p.n_camera + c) * (size_t)H * W * 4);
          const ColInfo* cinf = colinfo + c * W;
          const Strip* cst = strips + (size_t)c * W * strips_per_col;
          // thread -> (column quad q, rows r0, r0+rstep, ...): the 4 column records stay
          // in registers, a warp writes 512 contiguous bytes per store instruction
          const int rstep = p.TA / Q;
          if (rstep > 0) {
            const int q = tid % Q, r0 = tid / Q;
            if (r0 < rstep) paint_rows(out, cinf, cst, strips_per_col, rowsky, rowgnd, q, Q, r0, rstep, H);
          } else {
            // image wider than 4 x threads: several quads per thread
            for (int q = tid; q < Q; q += p.TA) paint_rows(out, cinf, cst, strips_per_col, rowsky, rowgnd, q, Q, 0, 1, H);
          }
        }
      }
    } else if (tid < p.TA + p.TB) {
      // ---- pool B: range / light / smell sensors, SL lanes per ray
      if (do_sense) {
        const int l = tid - p.TA, SL = p.SL;
        const int g = l & (SL - 1), grp = l / SL, ngrp = p.TB / SL;
        const uint32_t gmask = (SL == 32) ? 0xffffffffu : (((1u << SL) - 1u) << (lane & ~(SL - 1)));
        const int ntask = p.n_range + p.n_light + p.n_smell;
        const int nseg = S + 4 * R;
        for (int task = grp; task < ntask; task += ngrp) {  // task is uniform over a lane group
          if (task < p.n_range) {
            // RangeSensor.update
            const RangeDev& sd = p.ranges[task];
            const int r = sd.robot;
            double x1, y1, ca, sa;
            mount_point(ws + r * BRS_WS_STRIDE, sd.mx, sd.my, x1, y1, ca, sa);
            const float ox = (float)x1, oy = (float)y1;
            const int own_lo = S + 4 * r, own_hi = own_lo + 4;
            double reading = 1.0;
            for (int k = 0; k < sd.n_rays; ++k) {
              const double dxu = ds(dm(ca, sd.cr[k]), dm(sa, sd.sr[k]));
              const double dyu = da(dm(sa, sd.cr[k]), dm(ca, sd.sr[k]));
              const double x2 = da(dm(dxu, sd.max), x1), y2 = da(dm(dyu, sd.max), y1);
              const float rx = (float)ds(x2, x1), ry = (float)ds(y2, y1);
              uint32_t key = BRS_KEY_NOHIT;
              int idx = -1;
              search32(segtab, 0, own_lo, g, SL, ox, oy, rx, ry, key, idx);
              search32(segtab, own_hi, nseg, g, SL, ox, oy, rx, ry, key, idx);
              group_min(key, idx, SL, gmask);
              if (g == 0 && idx >= 0) {
                const double d = refine64(raw, SC, S, ws, idx, x1, y1, x2, y2);
                if (d < dm(reading, sd.max)) reading = d / sd.max;
              }
            }
            if (g == 0) p.range_out[(size_t)w * p.n_range + task] = (float)dm(reading, sd.max);
          } else if (task < p.n_range + p.n_light) {
            // LightSensor.update: line of sight to every bulb
            const int li = task - p.n_range;
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
              search32(segtab, 0, own_lo, g, SL, ox, oy, rx, ry, key, idx);
              search32(segtab, own_hi, nseg, g, SL, ox, oy, rx, ry, key, idx);
              group_min(key, idx, SL, gmask);
              if (idx < 0 && dist > 0.0 && br != 0.0) value = da(value, dm(br, p.light_mult) / dm(dist, dist));
            }
            if (g == 0) p.light_out[(size_t)w * p.n_light + li] = (float)value;
          } else {
            // SmellSensor.update: grid lookup at the mount point
            const int si = task - p.n_range - p.n_light;
            const PointDev& sd = p.smells[si];
            const int r = sd.robot;
            double x1, y1, ca, sa;
            mount_point(ws + r * BRS_WS_STRIDE, sd.mx, sd.my, x1, y1, ca, sa);
            float v = 0.f;
            if (p.grid_w > 0) {
              const int gx = (int)(x1 / p.smell_cell), gy = (int)(y1 / p.smell_cell);
              if (gx >= 0 && gx < p.grid_w && gy >= 0 && gy < p.grid_h)
                v = p.grid[((size_t)w * p.grid_h + gy) * p.grid_w + gx];
            }
            if (g == 0) p.smell_out[(size_t)w * p.n_smell + si] = v;
          }
        }
      }
      // ---- helper warp: PREPARE the next world while pool A is still busy
      if (is_helper) {
        __syncwarp();
        const int wn = w + G;
        if (wn < p.n_worlds) {
          const int nb = buf ^ 1;
          mbar_wait(smem_u32(&mbar[nb]), (uint32_t)((it + 1) >> 1) & 1u);
          prepare_world(p, wn, lane, do_move, (const uint32_t*)(smem + L.raw[nb]), (float4*)(smem + L.segtab[nb]),
                        (double*)(smem + L.ws[nb]), (int*)(smem + L.meta[nb]));
        }
      }
    }
    __syncthreads();  // world w is done; world w+1 is prepared
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

#endif  // __CUDACC__
}  // namespace brs
