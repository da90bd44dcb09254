// -*- C++ -*-
// The mini range kernel: a TEAM of G lanes (2..16) per world, for small one-robot worlds whose
// range sensors share one mount (BASELINE.json config 1 batched; the few-wall / few-ray corner
// of config 5's sweep).  A warp per world (brs_scan_kernel) leaves most lanes idle when a world
// has ten walls and one ray, and the fused kernel pays tile barriers and a helper hand-off for
// a world that is a few hundred instructions of work; here a warp carries 32 / G worlds and
// nothing leaves registers:
//
//   PREP     lanes 0 / 1 of the team: the committed half / the proposal half of Robot.step
//   COLLIDE  the team strides the walls (lane g: walls g, g + G, ...): FP32 disc cull, and for a
//            wall inside the disc the FP32 edge classifier + exact float64 4-edge test, serially
//            in that lane (rare); the team votes
//   SENSE    every lane holds ALL rays of the world (NRAY slots in registers); lane g builds
//            the shared-origin table entry of each of its walls (table_entry, the same function
//            as every other plan) and tests it against every ray at once; the per-lane winners
//            are reduced over the team with xor-shuffles (larger key, ties to the later wall)
//   REFINE   ray c is refined in float64 by lane c mod G (ua64, as everywhere); a sensor's
//            reading is the minimum over its fan (shared-memory atomicMin on the FP32 bits)
//   COMMIT   lane 0: pose + stall flag, lane 1: velocities
//
// Same arithmetic per (ray, wall) pair as the solo range kernel and as the fused kernel's group
// path, no cull can change a result, and the winner of a ray is the maximum of (key, wall
// index) whatever the order of evaluation: bit-identical to both (tests/test_parity_gpu.py).
// Upstream: robot.py Robot.step / cast_ray, devices/rangesensors.py RangeSensor.update
// (oracle/aitk_oracle.py restates them).
#pragma once
#include "brs_solo.cuh"

namespace brs {

struct MiniLayout {
  int rays, team0, per_team, total;
};
__host__ __device__ inline MiniLayout make_mini_layout(int n_range, int n_rays, int teams) {
  MiniLayout L;
  L.rays = 0;
  L.team0 = align_up(n_rays * (int)sizeof(SoloRay), 16);
  L.per_team = BRS_SOLO_REC * 8 + align_up(n_range * 4, 16);
  L.total = L.team0 + L.per_team * teams;
  return L;
}

#ifdef __CUDACC__
template <int G, int NRAY, int NT, int MIN_CTAS>
__global__ void __launch_bounds__(NT, MIN_CTAS) brs_mini_kernel(const __grid_constant__ KParams p) {
  static_assert(G >= 2 && G <= 32 && (G & (G - 1)) == 0, "team size");
  constexpr int RPL = (NRAY + G - 1) / G;  // rays a lane prepares / refines
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, g = lane & (G - 1), team_l = threadIdx.x / G;
  const unsigned tmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
  const int SC = p.seg_cap, K = p.s_nrays;
  const MiniLayout L = make_mini_layout(p.n_range, K, NT / G);
  SoloRay* const rays = (SoloRay*)(smem + L.rays);
  double* const rec = (double*)(smem + L.team0 + (size_t)team_l * L.per_team);
  uint32_t* const sout = (uint32_t*)(rec + BRS_SOLO_REC);
  const bool do_move = p.flags & 1u, do_sense = (p.flags & 2u) && K > 0;
  for (int i = threadIdx.x; i < K; i += NT) rays[i] = p.s_rays[i];
  __syncthreads();  // (the only CTA-wide barrier; the teams are independent from here)
  if (!do_move && !do_sense) return;

  const int nteams = gridDim.x * (NT / G), team = blockIdx.x * (NT / G) + team_l;
  const RobotDev& rd = p.s_robot;
  const GroupDev& gd = p.s_group;
  const int nst = max(p.n_steps, 1);  // brs_steps: a team owns its worlds for the whole launch
  for (int st = 0; st < nst; ++st)
  for (int w = team; w < p.n_worlds; w += nteams) {
    // ---------------- PREP ----------------------------------------------------------------
    if (g == 0) solo_prep_half(p, rd, w, false, rec);
    if (g == 1 && do_move) solo_prep_half(p, rd, w, true, rec);
    __syncwarp(tmask);
    const double* r = rec;
    const int S = (*(const int2*)(r + BRS_SR_META)).x;
    const uint32_t* const segw = p.seg + (size_t)w * 5 * SC;
    // ---------------- COLLIDE ---------------------------------------------------------------
    bool hit = true;  // !do_move: the committed state stays
    unsigned long long n_cls = 0;
    if (do_move) {
      bool mine = false;
      const float4 cp = *(const float4*)(r + BRS_SR_CPAR);
      for (int j = g; j < S && !mine; j += G) {
        const float4 wr = scan_wall_raw(segw, SC, j);
        const float4 s = make_float4(wr.x, wr.y, wr.z - wr.x, wr.w - wr.y);
        const float dx = cp.x - s.x, dy = cp.y - s.y;
        const float ee = fmaf(s.z, s.z, s.w * s.w);
        float h = fmaf(dx, s.z, dy * s.w) * rcp_approx(ee);
        h = fminf(fmaxf(h, 0.f), 1.f);  // NaN (zero-length wall) clamps to an end point
        const float vx = fmaf(-h, s.z, dx), vy = fmaf(-h, s.w, dy);
        if (fmaf(vx, vx, vy * vy) <= cp.z) {
          // rare: the proposed box against this wall, edge by edge (classifier, then exact)
          const double px = r[BRS_SR_PROP], py = r[BRS_SR_PROP + 1], c = r[BRS_SR_PROP + 3], sgn = r[BRS_SR_PROP + 4];
          n_cls += 4;
#pragma unroll 1
          for (int e = 0; e < 4 && !mine; ++e) {
            const int e2 = (e + 1) & 3;
            const double ax = da(px, ds(dm(c, rd.cox[e]), dm(sgn, rd.coy[e])));
            const double ay = da(py, da(dm(sgn, rd.cox[e]), dm(c, rd.coy[e])));
            const double bx2 = da(px, ds(dm(c, rd.cox[e2]), dm(sgn, rd.coy[e2])));
            const double by2 = da(py, da(dm(sgn, rd.cox[e2]), dm(c, rd.coy[e2])));
            if (edge_maybe_hits_at((float)px, (float)py, (float)rd.radius, p.cull_margin, s, ax, ay, bx2, by2))
              mine = isect64_bool(ax, ay, ds(bx2, ax), ds(by2, ay), (double)wr.x, (double)wr.y, (double)wr.z, (double)wr.w);
          }
        }
      }
      hit = __ballot_sync(tmask, mine) != 0u;
    }
    // ---------------- SENSE -------------------------------------------------------------------
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
      for (int i = g; i < p.n_range; i += G) sout[i] = __float_as_uint((float)p.ranges[i].max);
      __syncwarp(tmask);
      // lane g prepares rays g, g + G, ... (float64, as RangeSensor.update does), the team shares them
      float mrx[RPL], mry[RPL];
#pragma unroll
      for (int i = 0; i < RPL; ++i) {
        const int ray = g + i * G;
        mrx[i] = mry[i] = 0.f;  // (a slot past the last ray never hits: 1/t stays 0)
        if (ray < K) {
          const SoloRay sr = rays[ray];
          const double dxu = ds(dm(fc, sr.cr), dm(fs, sr.sr)), dyu = da(dm(fs, sr.cr), dm(fc, sr.sr));
          const double x2 = da(dm(dxu, sr.max), ox), y2 = da(dm(dyu, sr.max), oy);
          mrx[i] = (float)ds(x2, ox);
          mry[i] = (float)ds(y2, oy);
        }
      }
      float rx[NRAY], ry[NRAY];
      uint32_t key[NRAY];
      int widx[NRAY];
#pragma unroll
      for (int c = 0; c < NRAY; ++c) {
        rx[c] = __shfl_sync(tmask, mrx[c / G], c % G, G);
        ry[c] = __shfl_sync(tmask, mry[c / G], c % G, G);
        key[c] = BRS_GKEY_INIT;
        widx[c] = -1;
      }
      unsigned long long n_exec = 0;
      for (int j = g; j < S; j += G) {
        const float4 wr = scan_wall_raw(segw, SC, j);
        float4 t;
        if (!table_entry(make_float4(wr.x, wr.y, wr.z - wr.x, wr.w - wr.y), gc, t)) continue;
        ++n_exec;
#pragma unroll
        for (int c = 0; c < NRAY; ++c) {
          const float invt = fmaf(t.x, rx[c], t.y * ry[c]);     // 1/t = m . r
          const float wq = fmaf(rx[c], t.w, -(ry[c] * t.z));    // u/t = r x q
          // 0 <= u/t <= 1/t as one unsigned compare; nearer or equal (later wall) wins
          if (__float_as_uint(wq) <= __float_as_uint(invt) && (int)__float_as_uint(invt) >= (int)key[c]) {
            key[c] = __float_as_uint(invt);
            widx[c] = j;
          }
        }
      }
      // winners over the team; then lane c mod G refines ray c
      int mywall[RPL];
#pragma unroll
      for (int i = 0; i < RPL; ++i) mywall[i] = -1;
#pragma unroll
      for (int c = 0; c < NRAY; ++c) {
        group_max(key[c], widx[c], G, tmask);
        if ((c % G) == g) mywall[c / G] = widx[c];
      }
#pragma unroll
      for (int i = 0; i < RPL; ++i) {
        const int ray = g + i * G;
        if (ray < K && mywall[i] >= 0) {
          const SoloRay sr = rays[ray];
          const double dxu = ds(dm(fc, sr.cr), dm(fs, sr.sr)), dyu = da(dm(fs, sr.cr), dm(fc, sr.sr));
          const double x2 = da(dm(dxu, sr.max), ox), y2 = da(dm(dyu, sr.max), oy);
          const float4 wr = scan_wall_raw(segw, SC, mywall[i]);
          const double dist = dm(ua64(ox, oy, x2, y2, (double)wr.x, (double)wr.y, (double)wr.z, (double)wr.w), sr.max);
          const float rd32 = (float)fmin(sr.max, dist);
          // (+0.0 and -0.0 are the same reading; positive floats order like their bits)
          atomicMin(&sout[sr.sensor], __float_as_uint(rd32) & 0x7fffffffu);
        }
      }
      if (p.counters) {
        n_exec += __shfl_xor_sync(tmask, n_exec, 1);  // (teams of 2+: fold pairs, then one atomic per lane pair)
        if ((g & 1) == 0 && n_exec) atomicAdd(&p.counters[0], n_exec * (unsigned)K);
      }
      __syncwarp(tmask);
      for (int i = g; i < p.n_range; i += G)
        p.range_out[(size_t)w * p.n_range + i] = range_reading(p, w, i, __uint_as_float(sout[i]), (float)p.ranges[i].max);
    }
    if (p.counters && n_cls) atomicAdd(&p.counters[1], n_cls);
    // ---------------- COMMIT ------------------------------------------------------------------
    if (do_move) {
      const size_t o = (size_t)w * 3;
      if (g == 1) {
        p.vel[o] = hit ? 0.0 : r[BRS_SR_PROP + 5];
        p.vel[o + 1] = hit ? 0.0 : r[BRS_SR_PROP + 6];
        p.vel[o + 2] = hit ? 0.0 : r[BRS_SR_PROP + 7];
      } else if (g == 0) {
        if (!hit) {
          p.pose[o] = r[BRS_SR_PROP];
          p.pose[o + 1] = r[BRS_SR_PROP + 1];
          p.pose[o + 2] = r[BRS_SR_PROP + 2];
        }
        p.stalled[w] = (uint8_t)(hit ? 1 : 0);
      }
    }
    __syncwarp(tmask);  // the record and the readings are overwritten by the team's next world
  }
}
#endif  // __CUDACC__
}  // namespace brs
