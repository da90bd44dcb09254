// brs_solo.cuh -- the SOLO path of World.step: one WARP owns a world from its wall store to
// its finished picture, for the largest class of batched worlds -- one robot carrying one
// camera (BASELINE.json config 4: a million such worlds).
//
// Why a third launch plan.  The fused kernel (brs_kernels.cuh) and the wide path
// (brs_wide.cuh) run every world shape through the same tile machinery: runtime robot /
// group / sensor counts, fast divisions to find (world, robot, task) from a flat index, a
// candidate queue with atomics, a list-order replay over bit masks, column records in shared
// memory, team barriers between phases.  For a one-robot camera world most of that is
// overhead: ncu counted 2,370 warp-instructions per world of which the FP32 search is 400.
// Here everything that the shape makes constant is a compile-time constant, a phase hands
// its results to the next one in registers (shuffles), and nothing ever waits for another
// warp:
//
//   PREP     two lanes per world, BRS_SOLO_BATCH worlds at a time: the float64 pose proposal
//            of Robot.step (velocity ramp, sincos) with every lane busy;
//   then per world, all 32 lanes:
//   TMA      the world's wall store (SoA, 20*cap bytes) by one bulk copy into the warp's own
//            buffer, requested while the previous picture is being written;
//   COLLIDE  FP32 disc cull over the walls (lane per wall), FP32 edge classifier and exact
//            float64 4-edge test on what falls inside (rare: warp-uniform branch);
//   TABLE    shared-origin table m, q per wall, range / cone cull, ballot compaction;
//   SENSE    NR adjacent columns per lane: FP32 nearest-hit search over the compacted table
//            (broadcast LDS.128), float64 refinement of the winner, column record in registers;
//   PAINT    lane -> (row slot, 4-column quad), records fetched by shuffles; rows above every
//            wall top / inside every wall / below every bottom of the quad are constant
//            16-byte stores, only the rows in between take per-pixel selects; a warp store
//            covers 512 contiguous bytes;
//   COMMIT   per batch: poses, velocities, stalled flags to HBM.
//
// Same device functions and operation order as the other two paths: the outputs are
// bit-identical (tests/test_parity_gpu.py::test_culling_and_tiling_never_change_results).
#pragma once
#include "brs_kernels.cuh"

namespace brs {

#ifndef BRS_SOLO_TCH
#define BRS_SOLO_TCH 1  // chunks of 32 walls per TABLE pass (independent chains in flight)
#endif
#ifndef BRS_SOLO_CCH
#define BRS_SOLO_CCH 1  // chunks of 32 walls per COLLIDE pass
#endif
#define BRS_SOLO_BATCH 16  // worlds per PREP pass of a warp (two lanes per world)
#define BRS_SOLO_REC 16    // doubles per world record
// record (doubles): committed x y a cos sin | proposed x y a cos sin vx vy va | int2 (S, bits of
// the world extent) | float4 collision disc (2 doubles)
#define BRS_SR_POSE 0
#define BRS_SR_PROP 5
#define BRS_SR_META 13
#define BRS_SR_CPAR 14

struct SoloLayout {
  int raw, tab, idx, rec, mbar, per_warp;
};
__host__ __device__ inline SoloLayout make_solo_layout(int seg_cap) {
  SoloLayout L;
  int o = 0;
  L.raw = o; o += align_up(5 * seg_cap * 4, 16);
  L.tab = o; o += seg_cap * 16;
  L.idx = o; o += align_up(seg_cap * 2, 16);
  L.rec = o; o += BRS_SOLO_BATCH * BRS_SOLO_REC * 8;
  L.mbar = o; o += 16;
  L.per_warp = align_up(o, 128);
  return L;
}

#ifdef __CUDACC__

// One half of Robot.step's proposal for the solo record: the same operations, in the same
// order, as prep_robot_half (brs_kernels.cuh) minus the box corners (a lone robot's box is
// only needed when a wall falls inside its disc; collide() rebuilds it there).
// (R robots per world, this is robot `ri`: the multi path shares the function)
__device__ __forceinline__ void solo_prep_half_at(const KParams& p, const RobotDev& rd, int w, int R, int ri, bool prop,
                                                  double* __restrict__ r) {
  const size_t o = ((size_t)w * R + ri) * 3;
  const double x = p.pose[o], y = p.pose[o + 1], a = p.pose[o + 2];
  const double v0 = p.vel[o], v1 = p.vel[o + 1], v2 = p.vel[o + 2];
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
  const double c = prop ? sn : cs, sgn = prop ? cs : sn;
  if (prop) {
    const double px = da(x, da(dm(vx, sn), dm(dm(vy, cs), p.dt)));
    const double py = da(y, ds(dm(vx, cs), dm(dm(vy, sn), p.dt)));
    r[BRS_SR_PROP] = px; r[BRS_SR_PROP + 1] = py; r[BRS_SR_PROP + 2] = pa;
    r[BRS_SR_PROP + 3] = c; r[BRS_SR_PROP + 4] = sgn;
    r[BRS_SR_PROP + 5] = vx; r[BRS_SR_PROP + 6] = vy; r[BRS_SR_PROP + 7] = va;
    const float cxf = (float)px, cyf = (float)py;
    const float rad = (float)rd.radius + p.cull_margin + 1e-5f * (fabsf(cxf) + fabsf(cyf));
    *(float4*)(r + BRS_SR_CPAR) = make_float4(cxf, cyf, rad * rad, 0.f);
  } else {
    r[BRS_SR_POSE] = x; r[BRS_SR_POSE + 1] = y; r[BRS_SR_POSE + 2] = a;
    r[BRS_SR_POSE + 3] = c; r[BRS_SR_POSE + 4] = sgn;
    const int S = min(max(p.seg_count[w], 0), p.seg_cap);
    const float ext = fmaxf(p.world_size[2 * w], p.world_size[2 * w + 1]);
    *(int2*)(r + BRS_SR_META) = make_int2(S, __float_as_int(ext));
  }
}
__device__ __forceinline__ void solo_prep_half(const KParams& p, const RobotDev& rd, int w, bool prop,
                                               double* __restrict__ r) {
  solo_prep_half_at(p, rd, w, 1, 0, prop, r);
}

// The warp's staged wall store seen from one lane: the four SoA rows at this lane's column, so
// that wall j0 + lane of chunk j0 is four loads at constant offsets from four registers.
struct LaneWalls {
  const uint32_t *x1, *y1, *x2, *y2;
};
__device__ __forceinline__ LaneWalls lane_walls(const uint32_t* __restrict__ raw, int SC, int lane) {
  LaneWalls lw;
  lw.x1 = raw + lane; lw.y1 = lw.x1 + SC; lw.x2 = lw.y1 + SC; lw.y2 = lw.x2 + SC;
  return lw;
}
// wall j0 + lane as the FP32 search segment (x1, y1, ex, ey)
__device__ __forceinline__ float4 solo_wall(const LaneWalls& lw, int j0) {
  const float ax = __uint_as_float(lw.x1[j0]), ay = __uint_as_float(lw.y1[j0]);
  return make_float4(ax, ay, __uint_as_float(lw.x2[j0]) - ax, __uint_as_float(lw.y2[j0]) - ay);
}
// wall j (any lane) the same way
__device__ __forceinline__ float4 solo_wall_at(const uint32_t* __restrict__ raw, int SC, int j) {
  const float ax = __uint_as_float(raw[j]), ay = __uint_as_float(raw[SC + j]);
  return make_float4(ax, ay, __uint_as_float(raw[2 * SC + j]) - ax, __uint_as_float(raw[3 * SC + j]) - ay);
}

// Robot-vs-wall collision of the proposal (collide_commit_tile's COLLIDE for one robot):
// warp-uniform result.
__device__ __forceinline__ bool solo_collide(const KParams& p, const RobotDev& rd, const double* __restrict__ r,
                                             const uint32_t* __restrict__ raw, const LaneWalls& lw, int SC, int S,
                                             int lane) {
  const float4 cp = *(const float4*)(r + BRS_SR_CPAR);
  bool have_box = false;
  double ax = 0, ay = 0, bx2 = 0, by2 = 0;
  unsigned long long n_cls = 0;
  bool hit = false;
  // four chunks of 32 walls per pass: their disc tests are independent chains the scheduler can
  // interleave (a warp alone is latency bound); the ballots come after all four
  for (int jb = 0; jb < S; jb += 32 * BRS_SOLO_CCH) {
    unsigned inside_bits = 0;
#pragma unroll
    for (int c = 0; c < BRS_SOLO_CCH; ++c) {
      const int j0 = jb + 32 * c;
      bool inside = false;
      if (j0 + lane < S) {
        const float4 s = solo_wall(lw, j0);
        const float dx = cp.x - s.x, dy = cp.y - s.y;
        const float ee = fmaf(s.z, s.z, s.w * s.w);
        float h = fmaf(dx, s.z, dy * s.w) * rcp_approx(ee);
        h = fminf(fmaxf(h, 0.f), 1.f);  // NaN (zero-length wall) clamps to an end point
        const float vx = fmaf(-h, s.z, dx), vy = fmaf(-h, s.w, dy);
        inside = fmaf(vx, vx, vy * vy) <= cp.z;
      }
      inside_bits |= (inside ? 1u : 0u) << c;
    }
    if (!__any_sync(0xffffffffu, inside_bits != 0u)) continue;
#pragma unroll 1
    for (int c = 0; c < BRS_SOLO_CCH && !hit; ++c) {
    const int j0 = jb + 32 * c;
    unsigned bal = __ballot_sync(0xffffffffu, (inside_bits >> c) & 1u);
    while (bal) {  // rare, warp-uniform: lanes 4k+e test box edge e (all octets redundantly)
      const int jj = j0 + __ffs(bal) - 1;
      bal &= bal - 1;
      if (!have_box) {
        have_box = true;
        const int e = lane & 3, e2 = (e + 1) & 3;
        const double px = r[BRS_SR_PROP], py = r[BRS_SR_PROP + 1], c = r[BRS_SR_PROP + 3], sgn = r[BRS_SR_PROP + 4];
        ax = da(px, ds(dm(c, rd.cox[e]), dm(sgn, rd.coy[e])));
        ay = da(py, da(dm(sgn, rd.cox[e]), dm(c, rd.coy[e])));
        bx2 = da(px, ds(dm(c, rd.cox[e2]), dm(sgn, rd.coy[e2])));
        by2 = da(py, da(dm(sgn, rd.cox[e2]), dm(c, rd.coy[e2])));
      }
      const float4 sj = solo_wall_at(raw, SC, jj);
      const bool maybe = edge_maybe_hits_at((float)r[BRS_SR_PROP], (float)r[BRS_SR_PROP + 1], (float)rd.radius,
                                            p.cull_margin, sj, ax, ay, bx2, by2);
      n_cls += 4;
      if (__ballot_sync(0xffffffffu, maybe)) {
        const double x3 = (double)__uint_as_float(raw[jj]), y3 = (double)__uint_as_float(raw[SC + jj]);
        const double x4 = (double)__uint_as_float(raw[2 * SC + jj]), y4 = (double)__uint_as_float(raw[3 * SC + jj]);
        const bool h = isect64_bool(ax, ay, ds(bx2, ax), ds(by2, ay), x3, y3, x4, y4);
        if (__ballot_sync(0xffffffffu, h)) hit = true;
      }
      if (hit) break;
    }
    }
    if (hit) break;
  }
  if (p.counters && lane == 0 && n_cls) atomicAdd(&p.counters[1], n_cls);
  return hit;
}

// PAINT of one picture by one warp.  Lane -> (row slot, quad): a row of W = 32*NR pixels is
// Q = 8*NR quads, 32 / Q rows per warp store, every store covers 512 contiguous bytes.
// rec_rgba / rec_tb: this lane's NR column records (tb = top | bot << 16; a column that hit
// nothing is a "wall" of colour 0 over every row, which is what the blank picture holds).
// One picture quad to HBM -- and, when the engine is bound to peer gather buffers (PEERS), the
// same quad into each peer GPU's copy over NVLink: the all-gather of the pictures happens
// inside PAINT, store by store, instead of as a collective after the step.
template <bool PEERS>
__device__ __forceinline__ void solo_store(const KParams& p, uint4* o, const uint4 v) {
  __stcs(o, v);
  if (PEERS) {
#pragma unroll 1
    for (int k = 0; k < p.n_peers; ++k) *(uint4*)((char*)o + p.peer_delta[k]) = v;
  }
}
template <int NR, bool PEERS>
__device__ __forceinline__ void solo_paint(const KParams& p, uint4* __restrict__ out, const uint32_t (&rec_rgba)[NR],
                                           const uint32_t (&rec_tb)[NR], const uint32_t* __restrict__ rowsky,
                                           const uint32_t* __restrict__ rowgnd, uint32_t sky0, uint32_t gnd0, bool flat,
                                           int lane, int H) {
  constexpr int Q = 8 * NR, RP = 32 / Q;
  const int q = lane % Q, r0 = lane / Q;
  uint32_t rgba[4];
  int top[4], bot[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    // column 4q + k lives in lane (4q + k) / NR, slot (4q + k) % NR == k % NR
    const int src = (4 * q + k) / NR;
    const uint32_t a = __shfl_sync(0xffffffffu, rec_rgba[k % NR], src);
    const uint32_t b = __shfl_sync(0xffffffffu, rec_tb[k % NR], src);
    rgba[k] = a;
    top[k] = (int)(b & 0xffffu);
    bot[k] = (int)(b >> 16);
  }
  uint4* o = out + (size_t)r0 * Q + q;
  if (flat) {
    const uint32_t sky = sky0, gnd = gnd0;
    const int t0 = min(min(top[0], top[1]), min(top[2], top[3]));
    const int t1 = max(max(top[0], top[1]), max(top[2], top[3]));
    const int b0 = min(min(bot[0], bot[1]), min(bot[2], bot[3]));
    const int b1 = max(max(max(bot[0], bot[1]), max(bot[2], bot[3])), t1);
    const int e1 = min(t0, H), e2 = min(t1, H), e3 = min(b0, H), e4 = min(b1, H);
    const uint4 skyq = make_uint4(sky, sky, sky, sky), gndq = make_uint4(gnd, gnd, gnd, gnd);
    const uint4 wallq = make_uint4(rgba[0], rgba[1], rgba[2], rgba[3]);
    int j = r0;
    // rows of one kind, four stores per loop step (the step's tail is predicated): the loop
    // bookkeeping is what a row of constant pixels costs
#define BRS_SOLO_RUN(END, VAL)                                   \
    {                                                            \
      const int n = (END - j + RP - 1) / RP;                     \
      if (n > 0) {                                               \
        _Pragma("unroll 1") for (int i = 0; i < n; i += 4) {     \
          solo_store<PEERS>(p, o + 32 * i, VAL);                               \
          if (i + 1 < n) solo_store<PEERS>(p, o + 32 * i + 32, VAL);           \
          if (i + 2 < n) solo_store<PEERS>(p, o + 32 * i + 64, VAL);           \
          if (i + 3 < n) solo_store<PEERS>(p, o + 32 * i + 96, VAL);           \
        }                                                        \
        j += n * RP;                                             \
        o += 32 * n;                                             \
      }                                                          \
    }
    BRS_SOLO_RUN(e1, skyq)
    // rows where the quad's walls begin: above every bottom, a pixel is sky or wall
    {
      const int e2a = min(e2, e3);
#pragma unroll 1
      for (; j < e2a; j += RP, o += 32)
        solo_store<PEERS>(p, o, make_uint4(j < top[0] ? sky : rgba[0], j < top[1] ? sky : rgba[1], j < top[2] ? sky : rgba[2],
                             j < top[3] ? sky : rgba[3]));
    }
    // (a wall of the quad ends above where another begins: all three kinds in one row)
#define BRS_SOLO_PX(k) (j < top[k] ? sky : (j < bot[k] ? rgba[k] : gnd))
#pragma unroll 1
    for (; j < e2; j += RP, o += 32) solo_store<PEERS>(p, o, make_uint4(BRS_SOLO_PX(0), BRS_SOLO_PX(1), BRS_SOLO_PX(2), BRS_SOLO_PX(3)));
#undef BRS_SOLO_PX
    BRS_SOLO_RUN(e3, wallq)
    // rows where the walls end: below every top, a pixel is wall or ground
#pragma unroll 1
    for (; j < e4; j += RP, o += 32)
      solo_store<PEERS>(p, o, make_uint4(j < bot[0] ? rgba[0] : gnd, j < bot[1] ? rgba[1] : gnd, j < bot[2] ? rgba[2] : gnd,
                           j < bot[3] ? rgba[3] : gnd));
    BRS_SOLO_RUN(H, gndq)
#undef BRS_SOLO_RUN
  } else {
#pragma unroll 1
    for (int j = r0; j < H; j += RP, o += 32) {
      const uint32_t sky = rowsky[j], gnd = rowgnd[j];
      uint32_t px[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) px[k] = j < top[k] ? sky : (j < bot[k] ? rgba[k] : gnd);
      solo_store<PEERS>(p, o, make_uint4(px[0], px[1], px[2], px[3]));
    }
  }
}

// ------------------------------------------------------------------------------
// The solo step kernel.  NR = camera columns per lane (camera width = 32 * NR); NT threads per
// CTA (independent warps: the CTA is only the unit of shared-memory allocation).
// Worlds are dealt to the warps as contiguous, equal ranges: every warp finishes at about the
// same time without a scheduler.
// ------------------------------------------------------------------------------
template <int NR, int NT, int MIN_CTAS>
__global__ void __launch_bounds__(NT, MIN_CTAS) brs_solo_kernel(const __grid_constant__ KParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int SC = p.seg_cap;
  const SoloLayout L = make_solo_layout(SC);
  unsigned char* const sm = smem + (size_t)wib * L.per_warp;
  const uint32_t* const raw = (const uint32_t*)(sm + L.raw);
  float4* const tab = (float4*)(sm + L.tab);
  uint16_t* const ix = (uint16_t*)(sm + L.idx);
  double* const rec = (double*)(sm + L.rec);
  const uint32_t bar = smem_u32(sm + L.mbar);
  const LaneWalls lw = lane_walls(raw, SC, lane);
  const bool do_move = p.flags & 1u, do_cam = (p.flags & 4u) && p.n_camera > 0;
  if (!do_move && !do_cam) return;

  const unsigned nwarps = gridDim.x * (NT / 32), gw = blockIdx.x * (NT / 32) + wib;
  // the warp's worlds: world(i) = wbeg + i * wstr, i < cnt.  Interleaved (default): warp gw
  // takes worlds gw, gw + nwarps, ... so that the pictures being written at any moment form one
  // dense, sliding window of HBM; contiguous: equal consecutive ranges.
  int wbeg, wstr, cnt;
  if (p.solo_interleave) {
    wbeg = (int)gw; wstr = (int)nwarps;
    cnt = (int)gw < p.n_worlds ? (p.n_worlds - 1 - (int)gw) / (int)nwarps + 1 : 0;
  } else {
    wbeg = (int)((unsigned long long)p.n_worlds * gw / nwarps); wstr = 1;
    cnt = (int)((unsigned long long)p.n_worlds * (gw + 1) / nwarps) - wbeg;
  }
  if (cnt <= 0) return;
#define BRS_SOLO_WORLD(i) (wbeg + (i) * wstr)

  const uint32_t world_bytes = (uint32_t)(5 * SC * 4);
  if (lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar, world_bytes);
    tma_bulk_g2s(smem_u32(raw), p.seg + (size_t)wbeg * 5 * SC, world_bytes, bar);
  }
  __syncwarp();

  const RobotDev& rd = p.s_robot;  // (kernel parameters: constant bank, not HBM)
  const CameraDev& cd = p.s_camera;
  const GroupDev& gd = p.s_group;
  const int H = p.cam_h;
  const double Hd = (double)cd.height;
  double2 ccs[NR];
#pragma unroll
  for (int k = 0; k < NR; ++k) ccs[k] = do_cam ? p.cam_cs[lane * NR + k] : make_double2(0.0, 0.0);
  uint32_t parity = 0;

  // brs_steps: the warp runs its worlds n_steps times over; step s+1 of a world follows its
  // step s in this warp's program order, the wall copies run on across the boundary
  const int nst = max(p.n_steps, 1);
  for (int st = 0; st < nst; ++st)
  for (int b0 = 0; b0 < cnt; b0 += BRS_SOLO_BATCH) {
    const int nb = min(BRS_SOLO_BATCH, cnt - b0);
    // ---------------- PREP ----------------------------------------------------------------
    {
      const int wl = lane >> 1;
      const bool prop = lane & 1;
      if (wl < nb && (!prop || do_move)) solo_prep_half(p, rd, BRS_SOLO_WORLD(b0 + wl), prop, rec + wl * BRS_SOLO_REC);
    }
    __syncwarp();
    // the next batch's robot state, segment counts and world sizes: start them towards L1 now, a
    // batch of pictures before PREP needs them
    if (b0 + BRS_SOLO_BATCH < cnt) {
      if (wstr == 1) {
        if (lane < 14) {
          const int wn = BRS_SOLO_WORLD(b0 + BRS_SOLO_BATCH), wlast = BRS_SOLO_WORLD(cnt - 1);
          const char* a;
          if (lane < 12) {
            const double* base = lane < 4 ? p.pose : (lane < 8 ? p.vel : p.tvel);
            a = (const char*)(base + (size_t)wn * 3) + 128 * (lane & 3);
            if (a >= (const char*)(base + ((size_t)wlast + 1) * 3)) a = nullptr;
          } else {
            a = lane == 12 ? (const char*)(p.seg_count + wn) : (const char*)(p.world_size + 2 * (size_t)wn);
          }
          if (a) asm volatile("prefetch.global.L1 [%0];" ::"l"(a));
        }
      } else if (b0 + BRS_SOLO_BATCH + (lane >> 1) < cnt) {
        // scattered worlds: lane pair -> world, its pose / velocity rows and its counts
        const size_t wn = (size_t)BRS_SOLO_WORLD(b0 + BRS_SOLO_BATCH + (lane >> 1));
        const char* a = (lane & 1) ? (const char*)(p.vel + wn * 3) : (const char*)(p.pose + wn * 3);
        const char* b = (lane & 1) ? (const char*)(p.tvel + wn * 3) : (const char*)(p.seg_count + wn);
        asm volatile("prefetch.global.L1 [%0];" ::"l"(a));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(b));
      }
    }
    unsigned stall_mask = 0;
    for (int k = 0; k < nb; ++k) {
      const int w = BRS_SOLO_WORLD(b0 + k);
      const double* r = rec + k * BRS_SOLO_REC;
      const int2 meta = *(const int2*)(r + BRS_SR_META);
      const int S = meta.x;
      mbar_wait(bar, parity);
      parity ^= 1u;
      // ---------------- COLLIDE ------------------------------------------------------------
      bool hit = true;  // !do_move: the committed state stays
      if (do_move) hit = solo_collide(p, rd, r, raw, lw, SC, S, lane);
      if (hit) stall_mask |= 1u << k;
      uint32_t rec_rgba[NR], rec_tb[NR];
      if (do_cam) {
        // final state of the robot: the proposal if it moves, the committed state otherwise
        const double* f = hit ? r + BRS_SR_POSE : r + BRS_SR_PROP;
        const double fx = f[0], fy = f[1], fc = f[3], fs = f[4];
        // ---------------- TABLE -------------------------------------------------------------
        // (the group's origin and cone in the world frame: COMMIT of the other paths)
        const double ox = da(fx, ds(dm(fc, gd.mx), dm(fs, gd.my)));
        const double oy = da(fy, da(dm(fs, gd.mx), dm(fc, gd.my)));
        GroupCull gc;
        {
          const float cf = (float)fc, sf = (float)fs;
          gc.lo_x = cf * gd.lo_c - sf * gd.lo_s; gc.lo_y = sf * gd.lo_c + cf * gd.lo_s;
          gc.hi_x = cf * gd.hi_c - sf * gd.hi_s; gc.hi_y = sf * gd.hi_c + cf * gd.hi_s;
          gc.ax_x = cf * gd.ax_c - sf * gd.ax_s; gc.ax_y = sf * gd.ax_c + cf * gd.ax_s;
          gc.range2 = gd.range * gd.range;
          gc.has_cone = gd.has_cone;
          gc.ox = (float)ox;
          gc.oy = (float)oy;
        }
        int cw = 0;
        for (int jb = 0; jb < S; jb += 32 * BRS_SOLO_TCH) {
          // (independent chunks per pass, as in COLLIDE; compaction keeps index order)
          float4 t[BRS_SOLO_TCH];
          unsigned keep_bits = 0;
#pragma unroll
          for (int c = 0; c < BRS_SOLO_TCH; ++c) {
            const int j0 = jb + 32 * c;
            bool keep = false;
            if (j0 + lane < S) keep = table_entry(solo_wall(lw, j0), gc, t[c]);
            keep_bits |= (keep ? 1u : 0u) << c;
          }
#pragma unroll
          for (int c = 0; c < BRS_SOLO_TCH; ++c) {
            const bool keep = (keep_bits >> c) & 1u;
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if (keep) {
              const int pos = cw + __popc(bal & ((1u << lane) - 1u));
              tab[pos] = t[c];
              ix[pos] = (uint16_t)(jb + 32 * c + lane);
            }
            cw += __popc(bal);
          }
        }
        __syncwarp();
        if (p.counters && lane == 0) atomicAdd(&p.counters[0], (unsigned long long)cw * (unsigned)p.cam_w);
        // ---------------- SENSE -------------------------------------------------------------
        double x2[NR], y2[NR];
        float rx[NR], ry[NR];
        uint32_t key[NR];
        int idx[NR];
#pragma unroll
        for (int c = 0; c < NR; ++c) {
          const double dxu = ds(dm(fc, ccs[c].x), dm(fs, ccs[c].y));  // cos(a + theta_i)
          const double dyu = da(dm(fs, ccs[c].x), dm(fc, ccs[c].y));  // sin(a + theta_i)
          x2[c] = da(dm(dxu, cd.max_range), ox);
          y2[c] = da(dm(dyu, cd.max_range), oy);
          rx[c] = (float)ds(x2[c], ox);
          ry[c] = (float)ds(y2[c], oy);
          key[c] = BRS_GKEY_INIT;
          idx[c] = -1;
        }
        search_group<NR, true>(tab, 0, cw, 0, 1, rx, ry, key, idx);
        const double wsize = (double)__int_as_float(meta.y);
        // Camera._update of the lane's columns (camera_column without other robots), written
        // branch-free so that the float64 chains of the columns interleave: a column that hit
        // nothing refines wall 0 and discards the result
        int jw[NR];
        double x3[NR], y3[NR], x4[NR], y4[NR];
        uint32_t wall_rgba[NR];
#pragma unroll
        for (int c = 0; c < NR; ++c) {
          jw[c] = idx[c] >= 0 ? (int)ix[idx[c]] : 0;
          x3[c] = (double)__uint_as_float(raw[jw[c]]);
          y3[c] = (double)__uint_as_float(raw[SC + jw[c]]);
          x4[c] = (double)__uint_as_float(raw[2 * SC + jw[c]]);
          y4[c] = (double)__uint_as_float(raw[3 * SC + jw[c]]);
          wall_rgba[c] = raw[4 * SC + jw[c]];
        }
        double ua[NR], dw[NR];
#pragma unroll
        for (int c = 0; c < NR; ++c) ua[c] = ua64(ox, oy, x2[c], y2[c], x3[c], y3[c], x4[c], y4[c]);
#pragma unroll
        for (int c = 0; c < NR; ++c) dw[c] = dm(ua[c], cd.max_range) / wsize;
#pragma unroll
        for (int c = 0; c < NR; ++c) {
          const double s = clamp01(ds(1.0, dm(dw[c], cd.size_fade)));
          const double sc = clamp01(ds(1.0, dm(dw[c], cd.color_fade)));
          const uint32_t rgba = shade(wall_rgba[c], sc);
          const double high = dm(ds(1.0, s), Hd);
          const int top = (int)ceil(high / 2.0), bot = (int)ceil(ds(Hd, high / 2.0));
          const bool hitc = idx[c] >= 0;
          rec_rgba[c] = hitc ? rgba : 0u;
          // 0 <= top <= bot <= H < 2^15; nothing hit: colour 0 over every row
          rec_tb[c] = hitc ? ((uint32_t)top | ((uint32_t)bot << 16)) : 0x7fff0000u;
        }
      }
      // every lane is done with this world's walls: request the next world's while the
      // picture is written
      __syncwarp();
      if (lane == 0 && (b0 + k + 1 < cnt || st + 1 < nst)) {
        const int wnext = b0 + k + 1 < cnt ? w + wstr : wbeg;
        mbar_expect_tx(bar, world_bytes);
        tma_bulk_g2s(smem_u32(raw), p.seg + (size_t)wnext * 5 * SC, world_bytes, bar);
      }
      // ---------------- PAINT -----------------------------------------------------------------
#ifdef BRS_SOLO_NOPAINT  // tuning experiment: everything but the picture stores (one store keeps the records alive)
      if (do_cam && lane == 0) *(uint2*)(p.cam_out + (size_t)w * (size_t)H * p.cam_w * 4) = make_uint2(rec_rgba[0] ^ rec_rgba[NR - 1], rec_tb[0] ^ rec_tb[NR - 1]);
#else
      if (do_cam) {
        uint4* const pic = (uint4*)(p.cam_out + (size_t)w * (size_t)H * p.cam_w * 4);
        if (p.n_peers)
          solo_paint<NR, true>(p, pic, rec_rgba, rec_tb, p.row_sky, p.row_gnd, p.s_sky, p.s_gnd, p.flat_rows != 0, lane, H);
        else
          solo_paint<NR, false>(p, pic, rec_rgba, rec_tb, p.row_sky, p.row_gnd, p.s_sky, p.s_gnd, p.flat_rows != 0, lane, H);
      }
#endif
    }
    // ---------------- COMMIT ------------------------------------------------------------------
    if (do_move) {
      const int wl = lane >> 1;
      if (wl < nb) {
        const bool hit = (stall_mask >> wl) & 1u;
        const double* r = rec + wl * BRS_SOLO_REC;
        const int wc = BRS_SOLO_WORLD(b0 + wl);
        const size_t o = (size_t)wc * 3;
        if (lane & 1) {
          p.vel[o] = hit ? 0.0 : r[BRS_SR_PROP + 5];
          p.vel[o + 1] = hit ? 0.0 : r[BRS_SR_PROP + 6];
          p.vel[o + 2] = hit ? 0.0 : r[BRS_SR_PROP + 7];
        } else {
          if (!hit) {
            p.pose[o] = r[BRS_SR_PROP];
            p.pose[o + 1] = r[BRS_SR_PROP + 1];
            p.pose[o + 2] = r[BRS_SR_PROP + 2];
          }
          p.stalled[wc] = (uint8_t)(hit ? 1 : 0);
        }
      }
    }
    __syncwarp();  // the records are overwritten by the next batch
  }
#undef BRS_SOLO_WORLD
}

#endif  // __CUDACC__

// ==============================================================================
// The solo RANGE kernel: one warp per world for one robot whose range sensors all sit on one
// mount (BASELINE.json config 5: 1..256 rays against 10..2000 walls; config 1), no camera.
//
// Same skeleton as brs_solo_kernel -- PREP two lanes per world, COLLIDE, shared-origin TABLE,
// FP32 search, float64 refinement of the winner, COMMIT per batch -- with three differences:
//   * the walls are not staged: lanes read them straight from L2 / HBM, a wall per lane and
//     chunk, coalesced (COLLIDE and TABLE each stream the world once; a 2000-wall world would
//     not fit a warp's share of shared memory, and a 10-wall world does not need staging);
//   * the table is built and searched in blocks of BRS_SCAN_TB walls, so shared memory per warp
//     is fixed whatever the wall count; a ray's best hit is carried across blocks;
//   * rays map to lanes by count: >= 32 rays, NRL rays per lane (two per FFMA2); fewer, SL lanes
//     share one ray, striding the table, reduced with shuffles (ties to the larger index).
// A sensor's reading is the minimum over its rays (fan of 3): shared-memory atomicMin on the
// FP32 bits, then one coalesced store of the world's readings.
// ==============================================================================
#define BRS_SCAN_TB 128  // walls per TABLE / search block
struct __align__(16) SoloRay {
  double cr, sr;  // cos / sin of (sensor.a - incr_k)
  double max;
  int sensor, _pad;
};
struct ScanLayout {
  int tab, idx, rec, sout, per_warp, rays, total;
};
__host__ __device__ inline ScanLayout make_scan_layout(int n_range, int n_rays, int warps) {
  ScanLayout L;
  int o = 0;
  L.tab = o; o += BRS_SCAN_TB * 16;
  L.idx = o; o += BRS_SCAN_TB * 2;
  L.rec = o; o += BRS_SOLO_BATCH * BRS_SOLO_REC * 8;
  L.sout = o; o += align_up(n_range * 4, 16);
  L.per_warp = align_up(o, 128);
  L.rays = L.per_warp * warps;  // one copy of the ray table per CTA
  L.total = L.rays + align_up(n_rays * (int)sizeof(SoloRay), 16);
  return L;
}

#ifdef __CUDACC__
// wall j of world w straight from the HBM / L2 wall store: raw end points (exact in float64)
__device__ __forceinline__ float4 scan_wall_raw(const uint32_t* __restrict__ segw, int SC, int j) {
  return make_float4(__uint_as_float(__ldg(segw + j)), __uint_as_float(__ldg(segw + SC + j)),
                     __uint_as_float(__ldg(segw + 2 * SC + j)), __uint_as_float(__ldg(segw + 3 * SC + j)));
}

template <int NRL, int NT, int MIN_CTAS>
__global__ void __launch_bounds__(NT, MIN_CTAS) brs_scan_kernel(const __grid_constant__ KParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int SC = p.seg_cap, K = p.s_nrays, SL = p.s_SL;
  const ScanLayout L = make_scan_layout(p.n_range, K, NT / 32);
  unsigned char* const sm = smem + (size_t)wib * L.per_warp;
  float4* const tab = (float4*)(sm + L.tab);
  uint16_t* const ix = (uint16_t*)(sm + L.idx);
  double* const rec = (double*)(sm + L.rec);
  uint32_t* const sout = (uint32_t*)(sm + L.sout);
  SoloRay* const rays = (SoloRay*)(smem + L.rays);
  const bool do_move = p.flags & 1u, do_sense = (p.flags & 2u) && K > 0;
  for (int i = threadIdx.x; i < K; i += NT) rays[i] = p.s_rays[i];
  __syncthreads();  // (the only CTA-wide barrier: the ray table; the warps are independent from here)
  if (!do_move && !do_sense) return;

  const unsigned nwarps = gridDim.x * (NT / 32), gw = blockIdx.x * (NT / 32) + wib;
  int wbeg, wstr, cnt;
  if (p.solo_interleave) {
    wbeg = (int)gw; wstr = (int)nwarps;
    cnt = (int)gw < p.n_worlds ? (p.n_worlds - 1 - (int)gw) / (int)nwarps + 1 : 0;
  } else {
    wbeg = (int)((unsigned long long)p.n_worlds * gw / nwarps); wstr = 1;
    cnt = (int)((unsigned long long)p.n_worlds * (gw + 1) / nwarps) - wbeg;
  }
  if (cnt <= 0) return;
#define BRS_SOLO_WORLD(i) (wbeg + (i) * wstr)
  const RobotDev& rd = p.s_robot;
  const GroupDev& gd = p.s_group;
  // lane -> rays: slot c of this lane is ray (lane / SL) + c * (32 / SL); g = lane % SL strides the table
  const int g = lane & (SL - 1), rbase = lane / SL, rstep = 32 / SL;
  const uint32_t gmask = (SL == 32) ? 0xffffffffu : (((1u << SL) - 1u) << (lane & ~(SL - 1)));

  const int nst = max(p.n_steps, 1);  // brs_steps: see brs_solo_kernel
  for (int st = 0; st < nst; ++st)
  for (int b0 = 0; b0 < cnt; b0 += BRS_SOLO_BATCH) {
    const int nb = min(BRS_SOLO_BATCH, cnt - b0);
    // ---------------- PREP ----------------------------------------------------------------
    {
      const int wl = lane >> 1;
      const bool prop = lane & 1;
      if (wl < nb && (!prop || do_move)) solo_prep_half(p, rd, BRS_SOLO_WORLD(b0 + wl), prop, rec + wl * BRS_SOLO_REC);
    }
    __syncwarp();
    unsigned stall_mask = 0;
    for (int k = 0; k < nb; ++k) {
      const int w = BRS_SOLO_WORLD(b0 + k);
      const double* r = rec + k * BRS_SOLO_REC;
      const int2 meta = *(const int2*)(r + BRS_SR_META);
      const int S = meta.x;
      const uint32_t* const segw = p.seg + (size_t)w * 5 * SC;
      // the next world's wall rows: start them towards L2 now
      if (b0 + k + 1 < cnt) {
        const char* nxt = (const char*)(p.seg + (size_t)(w + wstr) * 5 * SC);
        for (int o = lane * 128; o < 16 * SC; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + o));
      }
      // ---------------- COLLIDE ------------------------------------------------------------
      bool hit = true;  // !do_move: the committed state stays
      if (do_move) {
        hit = false;
        const float4 cp = *(const float4*)(r + BRS_SR_CPAR);
        bool have_box = false;
        double ax = 0, ay = 0, bx2 = 0, by2 = 0;
        unsigned long long n_cls = 0;
        for (int jb = 0; jb < S && !hit; jb += 128) {
          // four chunks of 32 walls: all their loads are issued before the first is used (the
          // walls come from L2 / HBM, a chunk at a time would expose that latency four times)
          float4 wr4[4];
          unsigned inside_bits = 0;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int j = jb + 32 * c + lane;
            wr4[c] = j < S ? scan_wall_raw(segw, SC, j) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int j = jb + 32 * c + lane;
            const float4 wr = wr4[c];
            const float4 s = make_float4(wr.x, wr.y, wr.z - wr.x, wr.w - wr.y);
            const float dx = cp.x - s.x, dy = cp.y - s.y;
            const float ee = fmaf(s.z, s.z, s.w * s.w);
            float h = fmaf(dx, s.z, dy * s.w) * rcp_approx(ee);
            h = fminf(fmaxf(h, 0.f), 1.f);  // NaN (zero-length wall) clamps to an end point
            const float vx = fmaf(-h, s.z, dx), vy = fmaf(-h, s.w, dy);
            if (j < S && fmaf(vx, vx, vy * vy) <= cp.z) inside_bits |= 1u << c;
          }
          if (!__any_sync(0xffffffffu, inside_bits != 0u)) continue;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
          const float4 wr = wr4[c];
          unsigned bal = __ballot_sync(0xffffffffu, (inside_bits >> c) & 1u);
          while (bal && !hit) {  // rare, warp-uniform: lanes 4k+e test box edge e
            const int src = __ffs(bal) - 1;
            bal &= bal - 1;
            if (!have_box) {
              have_box = true;
              const int e = lane & 3, e2 = (e + 1) & 3;
              const double px = r[BRS_SR_PROP], py = r[BRS_SR_PROP + 1], c = r[BRS_SR_PROP + 3], sgn = r[BRS_SR_PROP + 4];
              ax = da(px, ds(dm(c, rd.cox[e]), dm(sgn, rd.coy[e])));
              ay = da(py, da(dm(sgn, rd.cox[e]), dm(c, rd.coy[e])));
              bx2 = da(px, ds(dm(c, rd.cox[e2]), dm(sgn, rd.coy[e2])));
              by2 = da(py, da(dm(sgn, rd.cox[e2]), dm(c, rd.coy[e2])));
            }
            const float x1 = __shfl_sync(0xffffffffu, wr.x, src), y1 = __shfl_sync(0xffffffffu, wr.y, src);
            const float x2 = __shfl_sync(0xffffffffu, wr.z, src), y2 = __shfl_sync(0xffffffffu, wr.w, src);
            const float4 sj = make_float4(x1, y1, x2 - x1, y2 - y1);
            const bool maybe = edge_maybe_hits_at((float)r[BRS_SR_PROP], (float)r[BRS_SR_PROP + 1], (float)rd.radius,
                                                  p.cull_margin, sj, ax, ay, bx2, by2);
            n_cls += 4;
            if (__ballot_sync(0xffffffffu, maybe)) {
              const bool h = isect64_bool(ax, ay, ds(bx2, ax), ds(by2, ay), (double)x1, (double)y1, (double)x2, (double)y2);
              if (__ballot_sync(0xffffffffu, h)) hit = true;
            }
          }
          }
        }
        if (p.counters && lane == 0 && n_cls) atomicAdd(&p.counters[1], n_cls);
      }
      if (hit) stall_mask |= 1u << k;
      if (do_sense) {
        const double* f = hit ? r + BRS_SR_POSE : r + BRS_SR_PROP;
        const double fx = f[0], fy = f[1], fc = f[3], fs = f[4];
        const double ox = da(fx, ds(dm(fc, gd.mx), dm(fs, gd.my)));
        const double oy = da(fy, da(dm(fs, gd.mx), dm(fc, gd.my)));
        GroupCull gc;
        {
          const float cf = (float)fc, sf = (float)fs;
          gc.lo_x = cf * gd.lo_c - sf * gd.lo_s; gc.lo_y = sf * gd.lo_c + cf * gd.lo_s;
          gc.hi_x = cf * gd.hi_c - sf * gd.hi_s; gc.hi_y = sf * gd.hi_c + cf * gd.hi_s;
          gc.ax_x = cf * gd.ax_c - sf * gd.ax_s; gc.ax_y = sf * gd.ax_c + cf * gd.ax_s;
          gc.range2 = gd.range * gd.range;
          gc.has_cone = gd.has_cone;
          gc.ox = (float)ox;
          gc.oy = (float)oy;
        }
        // the readings start at each sensor's range
        for (int i = lane; i < p.n_range; i += 32) sout[i] = __float_as_uint((float)p.ranges[i].max);
        // this lane's rays, scaled to their range (FP32, search only)
        float rx[NRL], ry[NRL];
        uint32_t key[NRL];
        int widx[NRL];  // winning wall (world index), -1: nothing hit yet
#pragma unroll
        for (int c = 0; c < NRL; ++c) {
          const int ray = rbase + c * rstep;
          rx[c] = ry[c] = 0.f;  // (a slot past the last ray never hits: 1/t stays 0)
          if (ray < K) {
            const SoloRay sr = rays[ray];
            const double dxu = ds(dm(fc, sr.cr), dm(fs, sr.sr)), dyu = da(dm(fs, sr.cr), dm(fc, sr.sr));
            const double x2 = da(dm(dxu, sr.max), ox), y2 = da(dm(dyu, sr.max), oy);
            rx[c] = (float)ds(x2, ox);
            ry[c] = (float)ds(y2, oy);
          }
          key[c] = BRS_GKEY_INIT;
          widx[c] = -1;
        }
        // ---------------- TABLE + SENSE, block by block ---------------------------------------
        unsigned long long n_exec = 0;
        for (int jb = 0; jb < S; jb += BRS_SCAN_TB) {
          int cw = 0;
          float4 wr4[BRS_SCAN_TB / 32];
#pragma unroll
          for (int c = 0; c < BRS_SCAN_TB / 32; ++c) {
            const int j = jb + 32 * c + lane;
            wr4[c] = j < S ? scan_wall_raw(segw, SC, j) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int c = 0; c < BRS_SCAN_TB / 32; ++c) {
            const int j = jb + 32 * c + lane;
            float4 t;
            bool keep = false;
            if (j < S) {
              const float4 wr = wr4[c];
              keep = table_entry(make_float4(wr.x, wr.y, wr.z - wr.x, wr.w - wr.y), gc, t);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if (keep) {
              const int pos = cw + __popc(bal & ((1u << lane) - 1u));
              tab[pos] = t;
              ix[pos] = (uint16_t)j;
            }
            cw += __popc(bal);
          }
          __syncwarp();
          n_exec += (unsigned long long)cw;
          int pos[NRL];
#pragma unroll
          for (int c = 0; c < NRL; ++c) pos[c] = -1;
          if (SL == 1)
            search_group<NRL, true>(tab, 0, cw, 0, 1, rx, ry, key, pos);
          else
            search_group<NRL, false>(tab, 0, cw, g, SL, rx, ry, key, pos);
#pragma unroll
          for (int c = 0; c < NRL; ++c)
            if (pos[c] >= 0) widx[c] = ix[pos[c]];  // a nearer hit in this block (ties: the later wall)
          __syncwarp();  // the next block overwrites the table
        }
        if (p.counters && lane == 0) atomicAdd(&p.counters[0], n_exec * (unsigned)K);
        // ---------------- refinement, readings ------------------------------------------------
#pragma unroll
        for (int c = 0; c < NRL; ++c) {
          if (SL > 1) group_max(key[c], widx[c], SL, gmask);
          const int ray = rbase + c * rstep;
          if (g == 0 && ray < K && widx[c] >= 0) {
            const SoloRay sr = rays[ray];
            const double dxu = ds(dm(fc, sr.cr), dm(fs, sr.sr)), dyu = da(dm(fs, sr.cr), dm(fc, sr.sr));
            const double x2 = da(dm(dxu, sr.max), ox), y2 = da(dm(dyu, sr.max), oy);
            const float4 wr = scan_wall_raw(segw, SC, widx[c]);
            const double dist = dm(ua64(ox, oy, x2, y2, (double)wr.x, (double)wr.y, (double)wr.z, (double)wr.w), sr.max);
            const float rd32 = (float)fmin(sr.max, dist);
            // (+0.0 and -0.0 are the same reading; positive floats order like their bits)
            atomicMin(&sout[sr.sensor], __float_as_uint(rd32) & 0x7fffffffu);
          }
        }
        __syncwarp();
        for (int i = lane; i < p.n_range; i += 32)
          p.range_out[(size_t)w * p.n_range + i] = range_reading(p, w, i, __uint_as_float(sout[i]), (float)p.ranges[i].max);
        __syncwarp();
      }
    }
    // ---------------- COMMIT ------------------------------------------------------------------
    if (do_move) {
      const int wl = lane >> 1;
      if (wl < nb) {
        const bool hit = (stall_mask >> wl) & 1u;
        const double* r = rec + wl * BRS_SOLO_REC;
        const int wc = BRS_SOLO_WORLD(b0 + wl);
        const size_t o = (size_t)wc * 3;
        if (lane & 1) {
          p.vel[o] = hit ? 0.0 : r[BRS_SR_PROP + 5];
          p.vel[o + 1] = hit ? 0.0 : r[BRS_SR_PROP + 6];
          p.vel[o + 2] = hit ? 0.0 : r[BRS_SR_PROP + 7];
        } else {
          if (!hit) {
            p.pose[o] = r[BRS_SR_PROP];
            p.pose[o + 1] = r[BRS_SR_PROP + 1];
            p.pose[o + 2] = r[BRS_SR_PROP + 2];
          }
          p.stalled[wc] = (uint8_t)(hit ? 1 : 0);
        }
      }
    }
    __syncwarp();  // the records are overwritten by the next batch
  }
#undef BRS_SOLO_WORLD
}
#endif  // __CUDACC__
}  // namespace brs
