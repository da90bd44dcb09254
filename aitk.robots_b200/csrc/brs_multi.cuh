// brs_multi.cuh -- the solo MULTI path of World.step: one warp per world for worlds of several
// robots without cameras (BASELINE.json config 3: 4 robots, 16 single-ray range sensors each on
// its own mount, light sensors, robot-robot collisions).
//
// The third member of the warp-per-world family (brs_solo.cuh).  What differs:
//   * R robots: PREP packs (robot, half) pairs of several worlds into one pass; COLLIDE tests
//     every (robot, wall) pair and every ordered robot pair (proposal against the other's old
//     box and, for earlier robots, against its proposal); COMMIT replays Robot.step's list
//     order on the pair masks exactly as collide_commit_tile does;
//   * every sensor has its own origin (no shared-origin table): the FP32 search is the general
//     two-segment form (search_direct), but over a PER-ROBOT list of the segments within reach
//     of that robot's sensors (walls and other robots' box edges, index order kept), which in
//     a 300 x 300 room with 60-unit sensors removes two tests in three;
//   * light sensors: (sensor, bulb) rays spread over the lanes, four lanes per ray, any-hit; the
//     brightness sum then runs over the bulbs in order, in float64, as the reference does.
// Same decisions, same float64 arithmetic, same tie rules as the fused kernel: bit-identical
// outputs (tests/test_parity_gpu.py).
#pragma once
#include "brs_solo.cuh"

namespace brs {

#define BRS_MULTI_REC 16  // doubles per robot record (the solo layout, brs_solo.cuh)
struct __align__(16) MultiRay {
  double cr, sr;  // cos / sin of (sensor.a - incr_k)
  double max;
  double mx, my;  // mount offset in the a=0 frame
  int sensor, robot;
};
struct MultiLayout {
  int rec, box, segtab, list, lcnt, cflag, sout, lvis, per_warp;
};
// bw = worlds per PREP batch, nvis = seg_cap + 4 R
__host__ __device__ inline MultiLayout make_multi_layout(int R, int bw, int nvis, int n_range, int n_light, int n_bulbs) {
  MultiLayout L;
  int o = 0;
  L.rec = o; o += bw * R * BRS_MULTI_REC * 8;
  L.box = o; o += R * 16 * 8;            // per robot: final / committed box (8 doubles), proposed box (8 doubles)
  L.segtab = o; o += nvis * 16;
  L.list = o; o += align_up(R * nvis, 16);  // per robot: indices of the segments within reach (uint8 / nvis <= 256)
  L.lcnt = o; o += 16 * 4;               // per robot: list length (R <= 8), padded
  L.cflag = o; o += R * 16;
  L.sout = o; o += align_up((n_range > 0 ? n_range : 1) * 4, 16);
  L.lvis = o; o += align_up((n_light * n_bulbs > 0 ? n_light * n_bulbs : 1) * 4, 16);
  L.per_warp = align_up(o, 128);
  return L;
}

#ifdef __CUDACC__

template <int NT, int MIN_CTAS>
__global__ void __launch_bounds__(NT, MIN_CTAS) brs_multi_kernel(const __grid_constant__ KParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int R = p.n_robots, SC = p.seg_cap, nvis = SC + 4 * R, BW = p.m_bw, K = p.m_nrays;
  const MultiLayout L = make_multi_layout(R, BW, nvis, p.n_range, p.n_light, p.n_bulbs);
  unsigned char* const sm = smem + (size_t)wib * L.per_warp;
  double* const rec = (double*)(sm + L.rec);
  double* const box = (double*)(sm + L.box);
  float4* const segtab = (float4*)(sm + L.segtab);
  uint8_t* const list = (uint8_t*)(sm + L.list);
  int* const lcnt = (int*)(sm + L.lcnt);
  int4* const cflag = (int4*)(sm + L.cflag);
  uint32_t* const sout = (uint32_t*)(sm + L.sout);
  uint32_t* const lvis = (uint32_t*)(sm + L.lvis);
  const bool do_move = p.flags & 1u, do_sense = p.flags & 2u;
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
#define BRS_MW(i) (wbeg + (i) * wstr)

  for (int b0 = 0; b0 < cnt; b0 += BW) {
    const int nb = min(BW, cnt - b0);
    // ---------------- PREP: (world, robot, half) over the lanes -------------------------------
    for (int h = lane; h < nb * R * 2; h += 32) {
      const int i = h >> 1, wl = i / R, r = i - wl * R;
      const bool prop = h & 1;
      if (!prop || do_move) solo_prep_half_at(p, p.robots[r], BRS_MW(b0 + wl), R, r, prop, rec + (size_t)i * BRS_MULTI_REC);
    }
    __syncwarp();
    unsigned stall_mask = 0;  // bit wl * R + r
    for (int k = 0; k < nb; ++k) {
      const int w = BRS_MW(b0 + k);
      const double* wr = rec + (size_t)k * R * BRS_MULTI_REC;
      const int2 meta = *(const int2*)(wr + BRS_SR_META);
      const int S = meta.x, nseg = S + 4 * R;
      const uint32_t* const segw = p.seg + (size_t)w * 5 * SC;
      if (b0 + k + 1 < cnt) {  // the next world's wall rows: start them towards L2
        const char* nxt = (const char*)(p.seg + (size_t)(w + wstr) * 5 * SC);
        for (int o = lane * 128; o < 16 * SC; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + o));
      }
      // ---------------- boxes (committed and proposed), wall table, flags ---------------------
      for (int i = lane; i < R * 8; i += 32) {
        const int r = i >> 3, hf = (i >> 2) & 1, c = i & 3;  // robot, half (0 committed, 1 proposed), corner
        const double* rr = wr + r * BRS_MULTI_REC + (hf ? BRS_SR_PROP : BRS_SR_POSE);
        const RobotDev& rd = p.robots[r];
        const double cx = rr[0], cy = rr[1], cs = rr[3], sn = rr[4], ox = rd.cox[c], oy = rd.coy[c];
        double* bx = box + r * 16 + hf * 8 + 2 * c;
        if (hf == 0 || do_move) {
          bx[0] = da(cx, ds(dm(cs, ox), dm(sn, oy)));
          bx[1] = da(cy, da(dm(sn, ox), dm(cs, oy)));
        }
      }
      for (int j = lane; j < S; j += 32) {
        const float4 rw = scan_wall_raw(segw, SC, j);
        segtab[j] = make_float4(rw.x, rw.y, rw.z - rw.x, rw.w - rw.y);
      }
      if (lane < R) cflag[lane] = make_int4(0, 0, 0, 0);
      __syncwarp();
      if (do_move) {
        // ---------------- COLLIDE: walls ------------------------------------------------------
        unsigned long long n_cls = 0;
        for (int i0 = 0; i0 < R * S; i0 += 32) {
          const int it = i0 + lane, r = it < R * S ? it / S : 0, j = it - r * S;
          bool inside = false;
          if (it < R * S) {
            const float4 cp = *(const float4*)(wr + r * BRS_MULTI_REC + BRS_SR_CPAR);
            const float4 s = segtab[j];
            const float dx = cp.x - s.x, dy = cp.y - s.y;
            const float ee = fmaf(s.z, s.z, s.w * s.w);
            float h = fmaf(dx, s.z, dy * s.w) * rcp_approx(ee);
            h = fminf(fmaxf(h, 0.f), 1.f);
            const float vx = fmaf(-h, s.z, dx), vy = fmaf(-h, s.w, dy);
            inside = fmaf(vx, vx, vy * vy) <= cp.z;
          }
          unsigned bal = __ballot_sync(0xffffffffu, inside);
          while (bal) {  // rare, warp-uniform: lanes 4k+e test edge e of the proposal against the wall
            const int src = __ffs(bal) - 1;
            bal &= bal - 1;
            const int ci = i0 + src, cr = ci / S, cj = ci - cr * S;
            if (cflag[cr].x) continue;  // this robot already hit a wall
            const int e = lane & 3, e2 = (e + 1) & 3;
            const double* pb = box + cr * 16 + 8;
            const double ax = pb[2 * e], ay = pb[2 * e + 1], bx2 = pb[2 * e2], by2 = pb[2 * e2 + 1];
            const double* rr = wr + cr * BRS_MULTI_REC;
            const bool maybe = edge_maybe_hits_at((float)rr[BRS_SR_PROP], (float)rr[BRS_SR_PROP + 1], (float)p.robots[cr].radius,
                                                  p.cull_margin, segtab[cj], ax, ay, bx2, by2);
            n_cls += 4;
            if (__ballot_sync(0xffffffffu, maybe)) {
              const float4 rw = scan_wall_raw(segw, SC, cj);
              const bool h = isect64_bool(ax, ay, ds(bx2, ax), ds(by2, ay), (double)rw.x, (double)rw.y, (double)rw.z, (double)rw.w);
              if (__ballot_sync(0xffffffffu, h) && lane == 0) cflag[cr].x = 1;
              __syncwarp();
            }
          }
        }
        if (p.counters && lane == 0 && n_cls) atomicAdd(&p.counters[1], n_cls);
        // ---------------- COLLIDE: robot pairs (collide_commit_tile's pair loop) ----------------
        if (R > 1) {
          const int per_world = R * R * 2;
          for (int i = lane; i < per_world; i += 32) {
            const int ri = (i >> 1) / R, ro = (i >> 1) - ri * R, v = i & 1;
            if (ro == ri || (v && ro > ri)) continue;
            const double* wi = wr + ri * BRS_MULTI_REC;
            const double* wo = wr + ro * BRS_MULTI_REC;
            const double* pr = box + ri * 16 + 8;               // proposed box of ri
            const double* ob = box + ro * 16 + (v ? 8 : 0);     // proposal or committed box of ro
            const double ocx = v ? wo[BRS_SR_PROP] : wo[BRS_SR_POSE], ocy = v ? wo[BRS_SR_PROP + 1] : wo[BRS_SR_POSE + 1];
            const double ddx = wi[BRS_SR_PROP] - ocx, ddy = wi[BRS_SR_PROP + 1] - ocy;
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
            if (hit) atomicOr(v ? &cflag[ri].z : &cflag[ri].y, 1 << ro);
          }
        }
        __syncwarp();
        // ---------------- COMMIT: list-order replay on the masks --------------------------------
        if (lane < R) {
          const int r = lane;
          unsigned sm_ = 0;
          bool hit = false;
          for (int q = 0; q <= r; ++q) {
            const int4 f = cflag[q];
            const unsigned before = (1u << q) - 1u;
            const unsigned m = ((unsigned)f.z & before & ~sm_) | ((unsigned)f.y & (~before | sm_));
            hit = f.x || m;
            if (hit) sm_ |= 1u << q;
          }
          cflag[r].w = hit ? 1 : 0;
        }
        __syncwarp();
        unsigned hits = 0;
        for (int r = 0; r < R; ++r) hits |= (cflag[r].w ? 1u : 0u) << r;
        stall_mask |= hits << (k * R);
        // the final box of a robot that moves is its proposal
        for (int i = lane; i < R * 8; i += 32) {
          const int r = i >> 3;
          if (!((hits >> r) & 1u)) box[r * 16 + (i & 7)] = box[r * 16 + 8 + (i & 7)];
        }
        __syncwarp();
      }
      const unsigned hits_now = do_move ? (stall_mask >> (k * R)) & ((1u << R) - 1u) : ~0u;  // !do_move: committed state
      if (do_sense) {
        // ---------------- robot boxes as search segments ----------------------------------------
        if (lane < R * 4) {
          const int r = lane >> 2, e = lane & 3, e2 = (e + 1) & 3;
          const double* fb = box + r * 16;
          const float x1 = (float)fb[2 * e], y1 = (float)fb[2 * e + 1];
          segtab[S + lane] = make_float4(x1, y1, (float)fb[2 * e2] - x1, (float)fb[2 * e2 + 1] - y1);
        }
        for (int i = lane; i < p.n_range; i += 32) sout[i] = __float_as_uint((float)p.ranges[i].max);
        __syncwarp();
        // ---------------- per-robot lists of the segments within reach ----------------------------
        for (int r = 0; r < R; ++r) {
          const double* f = wr + r * BRS_MULTI_REC + (((hits_now >> r) & 1u) ? BRS_SR_POSE : BRS_SR_PROP);
          const float cx = (float)f[0], cy = (float)f[1];
          const float reach = p.m_reach[r], reach2 = reach * reach;
          int n = 0;
          for (int j0 = 0; j0 < nseg; j0 += 32) {
            const int j = j0 + lane;
            bool keep = false;
            if (j < nseg && (unsigned)(j - (S + 4 * r)) >= 4u) {
              const float4 s = segtab[j];
              const float dx = cx - s.x, dy = cy - s.y;
              const float ee = fmaf(s.z, s.z, s.w * s.w);
              float h = fmaf(dx, s.z, dy * s.w) * rcp_approx(ee);
              h = fminf(fmaxf(h, 0.f), 1.f);
              const float vx = fmaf(-h, s.z, dx), vy = fmaf(-h, s.w, dy);
              keep = !(fmaf(vx, vx, vy * vy) > reach2);  // (NaN keeps the segment)
            }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if (keep) list[r * nvis + n + __popc(bal & ((1u << lane) - 1u))] = (uint8_t)j;
            n += __popc(bal);
          }
          if (lane == 0) lcnt[r] = n;
        }
        __syncwarp();
        // ---------------- SENSE: range rays, a lane per ray ----------------------------------------
        unsigned long long n_exec = 0;
        for (int q0 = 0; q0 < K; q0 += 32) {
          const int q = q0 + lane;
          if (q < K) {
            const MultiRay mr = p.m_rays[q];
            const int r = mr.robot;
            const double* f = wr + r * BRS_MULTI_REC + (((hits_now >> r) & 1u) ? BRS_SR_POSE : BRS_SR_PROP);
            const double ca = f[3], sa = f[4];
            const double x1 = da(f[0], ds(dm(ca, mr.mx), dm(sa, mr.my)));
            const double y1 = da(f[1], da(dm(sa, mr.mx), dm(ca, mr.my)));
            const double dxu = ds(dm(ca, mr.cr), dm(sa, mr.sr)), dyu = da(dm(sa, mr.cr), dm(ca, mr.sr));
            const double x2 = da(dm(dxu, mr.max), x1), y2 = da(dm(dyu, mr.max), y1);
            const float ox = (float)x1, oy = (float)y1, rx = (float)ds(x2, x1), ry = (float)ds(y2, y1);
            uint32_t key = BRS_KEY_NOHIT;
            int idx = -1;
            const uint8_t* li = list + r * nvis;
            const int n = lcnt[r];
            n_exec += (unsigned)n;
#pragma unroll 2
            for (int i = 0; i < n; ++i) {
              const int j = li[i];
              uint32_t tb, ub;
              direct_test(segtab[j], ox, oy, rx, ry, tb, ub);
              if (ub <= 0x3F800000u && tb <= key) {
                key = tb;
                idx = j;
              }
            }
            if (idx >= 0) {
              double x3, y3, x4, y4;
              if (idx < S) {
                const float4 rw = scan_wall_raw(segw, SC, idx);
                x3 = (double)rw.x; y3 = (double)rw.y; x4 = (double)rw.z; y4 = (double)rw.w;
              } else {
                const int o = (idx - S) >> 2, e = (idx - S) & 3, e2 = (e + 1) & 3;
                const double* fb = box + o * 16;
                x3 = fb[2 * e]; y3 = fb[2 * e + 1]; x4 = fb[2 * e2]; y4 = fb[2 * e2 + 1];
              }
              const double dist = dm(ua64(x1, y1, x2, y2, x3, y3, x4, y4), mr.max);
              atomicMin(&sout[mr.sensor], __float_as_uint((float)fmin(mr.max, dist)) & 0x7fffffffu);
            }
          }
        }
        if (p.counters) {
          for (int o = 16; o > 0; o >>= 1) n_exec += __shfl_xor_sync(0xffffffffu, n_exec, o);
          if (lane == 0) atomicAdd(&p.counters[0], n_exec);
        }
        // ---------------- SENSE: lights: (sensor, bulb) rays, four lanes per ray, any hit ----------
        const int NLB = p.n_light * p.n_bulbs;
        for (int q0 = 0; q0 < NLB * 4; q0 += 32) {
          const int q = (q0 + lane) >> 2, g = lane & 3;
          bool blocked = false;
          if (q < NLB) {
            const int li = q / p.n_bulbs, b = q - li * p.n_bulbs;
            const PointDev& sd = p.lights[li];
            const int r = sd.robot;
            const double* f = wr + r * BRS_MULTI_REC + (((hits_now >> r) & 1u) ? BRS_SR_POSE : BRS_SR_PROP);
            const double ca = f[3], sa = f[4];
            const double x1 = da(f[0], ds(dm(ca, sd.mx), dm(sa, sd.my)));
            const double y1 = da(f[1], da(dm(sa, sd.mx), dm(ca, sd.my)));
            const float* bl = p.bulbs + ((size_t)w * p.n_bulbs + b) * 4;
            const float ox = (float)x1, oy = (float)y1;
            const float rx = (float)ds((double)bl[0], x1), ry = (float)ds((double)bl[1], y1);
            const int own = S + 4 * r;
            for (int j = g; j < nseg; j += 4) {
              if ((unsigned)(j - own) < 4u) continue;
              uint32_t tb, ub;
              direct_test(segtab[j], ox, oy, rx, ry, tb, ub);
              if (ub <= 0x3F800000u && tb <= BRS_KEY_NOHIT) blocked = true;
            }
          }
          const unsigned bal = __ballot_sync(0xffffffffu, blocked);
          if (q < NLB && g == 0) lvis[q] = ((bal >> (lane & ~3)) & 0xFu) ? 0u : 1u;
        }
        __syncwarp();
        for (int li = lane; li < p.n_light; li += 32) {
          const PointDev& sd = p.lights[li];
          const int r = sd.robot;
          const double* f = wr + r * BRS_MULTI_REC + (((hits_now >> r) & 1u) ? BRS_SR_POSE : BRS_SR_PROP);
          const double ca = f[3], sa = f[4];
          const double x1 = da(f[0], ds(dm(ca, sd.mx), dm(sa, sd.my)));
          const double y1 = da(f[1], da(dm(sa, sd.mx), dm(ca, sd.my)));
          double value = 0.0;
          for (int b = 0; b < p.n_bulbs; ++b) {
            const float* bl = p.bulbs + ((size_t)w * p.n_bulbs + b) * 4;
            const double ex = ds((double)bl[0], x1), ey = ds((double)bl[1], y1), br = (double)bl[3];
            const double dist = sqrt(da(dm(ex, ex), dm(ey, ey)));
            if (lvis[li * p.n_bulbs + b] && dist > 0.0 && br != 0.0) value = da(value, dm(br, p.light_mult) / dm(dist, dist));
          }
          p.light_out[(size_t)w * p.n_light + li] = light_reading(p, w, li, (float)value);
        }
        if (p.counters && lane == 0 && NLB) atomicAdd(&p.counters[0], (unsigned long long)NLB * (nseg - 4));
        // ---------------- smell sensors, readings out --------------------------------------------
        for (int si = lane; si < p.n_smell; si += 32) {
          const PointDev& sd = p.smells[si];
          const int r = sd.robot;
          const double* f = wr + r * BRS_MULTI_REC + (((hits_now >> r) & 1u) ? BRS_SR_POSE : BRS_SR_PROP);
          const double x1 = da(f[0], ds(dm(f[3], sd.mx), dm(f[4], sd.my)));
          const double y1 = da(f[1], da(dm(f[4], sd.mx), dm(f[3], sd.my)));
          float v = 0.f;
          if (p.grid_w > 0) {
            const int gx = (int)(x1 / p.smell_cell), gy = (int)(y1 / p.smell_cell);
            if (gx >= 0 && gx < p.grid_w && gy >= 0 && gy < p.grid_h) v = p.grid[((size_t)w * p.grid_h + gy) * p.grid_w + gx];
          }
          p.smell_out[(size_t)w * p.n_smell + si] = v;
        }
        __syncwarp();
        for (int i = lane; i < p.n_range; i += 32)
          p.range_out[(size_t)w * p.n_range + i] = range_reading(p, w, i, __uint_as_float(sout[i]), (float)p.ranges[i].max);
        __syncwarp();
      }
    }
    // ---------------- COMMIT to HBM, the whole batch ---------------------------------------------
    if (do_move) {
      for (int h = lane; h < nb * R * 2; h += 32) {
        const int i = h >> 1, wl = i / R, r = i - wl * R;
        const bool hit = (stall_mask >> i) & 1u;
        const double* rr = rec + (size_t)i * BRS_MULTI_REC;
        const int wc = BRS_MW(b0 + wl);
        const size_t o = ((size_t)wc * R + r) * 3;
        if (h & 1) {
          p.vel[o] = hit ? 0.0 : rr[BRS_SR_PROP + 5];
          p.vel[o + 1] = hit ? 0.0 : rr[BRS_SR_PROP + 6];
          p.vel[o + 2] = hit ? 0.0 : rr[BRS_SR_PROP + 7];
        } else {
          if (!hit) {
            p.pose[o] = rr[BRS_SR_PROP];
            p.pose[o + 1] = rr[BRS_SR_PROP + 1];
            p.pose[o + 2] = rr[BRS_SR_PROP + 2];
          }
          p.stalled[(size_t)wc * R + r] = (uint8_t)(hit ? 1 : 0);
        }
      }
    }
    __syncwarp();
  }
#undef BRS_MW
}
#endif  // __CUDACC__
}  // namespace brs
