// brs_wide.cuh -- the WIDE path of World.step: two kernels chained with programmatic
// dependent launch, for batches with many more worlds than resident CTAs.
//
//   brs_prep_kernel   thread per (world, robot, half): the float64 pose proposal of
//                     Robot.step -- velocity ramp, sincos, committed and proposed box
//                     corners, the FP32 disc of the collision cull -- written as one
//                     BRS_WS_STRIDE record per robot to HBM.  This is the only part of a
//                     step that is a long dependent chain per robot; run wide (every lane
//                     a different robot half) it costs a few microseconds for 10^5 robots
//                     where one helper warp per CTA made it the critical path.
//   brs_wide_kernel   a CTA (one warp for worlds whose rays fill a warp, up to 8 warps for
//                     worlds with thousands of walls) owns a tile of TW worlds from the wall
//                     store to the finished pictures, with NO warp specialisation:
//                       TMA     walls + the tile's robot records (cp.async.bulk, mbarrier,
//                               double buffered: tile n+1 lands while tile n is processed)
//                       COLLIDE disc cull over the walls, FP32 classifier, exact FP64 test
//                       COMMIT  list-order replay, commit / stall, poses to HBM
//                       TABLE / SENSE / PAINT  as in the fused kernel (same device code)
//                     A one-warp CTA synchronises with __syncwarp only: 20+ independent warps
//                     per SM sit in different phases (FP32 search, FP64 refinement, HBM
//                     stores), which is what keeps the issue slots busy.
//
// Both paths run the same device functions (brs_kernels.cuh) in the same order per world,
// so their outputs are bit-identical (tests/test_parity_gpu.py asserts it).
#pragma once
#include "brs_kernels.cuh"

namespace brs {

struct WideLayout {
  int raw[2], ws[2], segtab, gtab, gidx, gcnt, gcull, org, cflag, meta, colinfo, strips, camcs, rowsky, rowgnd, cq, mbar,
      sched, total;
};
// collision candidate queue of a wide tile: a robot's disc touches a handful of walls; what
// does not fit is tested in place (slower, same result)
__host__ __device__ inline int wide_cq_max(int TW, int R) {
  const int n = TW * R * 16;
  return n < 28 ? 28 : (n > BRS_CQ_MAX ? BRS_CQ_MAX : n);
}
// `tables_in_smem`: the camera's column table (cos / sin per column) and the row colours are
// copied to shared memory per CTA; one-warp teams read them through L1 instead (a kilobyte
// per CTA is a tenth of such a team's footprint)
__host__ __device__ inline WideLayout make_wide_layout(int TW, int seg_cap, int R, int n_group, int n_camera, int cam_w,
                                                       int cam_h, bool stage_walls, bool tables_in_smem) {
  WideLayout L;
  const int nvis = seg_cap + 4 * R;
  int o = 0;
  for (int b = 0; b < 2; ++b) { L.raw[b] = o; o += stage_walls ? TW * 5 * seg_cap * 4 : 0; o = align_up(o, 128); }
  for (int b = 0; b < 2; ++b) { L.ws[b] = o; o += TW * R * BRS_WS_STRIDE * 8; }
  L.segtab = o; o += TW * nvis * 16;
  L.gtab = o; o += TW * n_group * nvis * 16;
  L.gcull = o; o += TW * n_group * (int)sizeof(GroupCull);
  L.gcnt = o; o += align_up(TW * n_group * 8, 16);
  L.gidx = o; o += align_up(TW * n_group * nvis * 2, 16);
  L.org = o; o += TW * (n_group > 0 ? n_group : 1) * 16;
  L.cflag = o; o += TW * R * 16;
  L.meta = o; o += align_up(TW * 8, 16);
  L.colinfo = o; o += TW * n_camera * cam_w * (int)sizeof(ColInfo);
  L.strips = o; o += TW * n_camera * cam_w * 4 * (R - 1) * (int)sizeof(Strip);
  o = align_up(o, 16);
  L.camcs = o; o += tables_in_smem ? n_camera * cam_w * 16 : 0;
  L.rowsky = o; o += tables_in_smem ? cam_h * 4 : 0;
  L.rowgnd = o; o += tables_in_smem ? cam_h * 4 : 0;
  o = align_up(o, 16);
  L.cq = o; o += 16 + 4 * wide_cq_max(TW, R);
  L.mbar = o; o += 2 * 8;
  L.sched = o; o += 16;
  L.total = align_up(o, 16);
  return L;
}

#ifdef __CUDACC__

__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ------------------------------------------------------------------------------
// PREP: thread per (world, robot, half).  Launch: ceil(n_worlds * R * halves / 256) CTAs.
// ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) brs_prep_kernel(const __grid_constant__ KParams p) {
  // the dependent (brs_wide_kernel) may start its prologue -- barrier set-up, camera tables,
  // the first tile's walls -- while this grid runs; it waits (griddepcontrol.wait) before
  // it touches anything written here
  griddep_launch_dependents();
  const bool do_move = p.flags & 1u;
  const int R = p.n_robots;
  const unsigned total = (unsigned)p.n_worlds * (unsigned)R * (do_move ? 2u : 1u);
  const unsigned h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= total) return;
  const unsigned i = do_move ? (h >> 1) : h;
  const bool prop = do_move && (h & 1u);
  const int w = (int)fdiv(i, p.dR), r = (int)i - w * R;
  double* wsr = p.wsg + (size_t)i * BRS_WS_STRIDE;
  if (!prop && r == 0) {
    // per-world meta rides in robot 0's record, so that the step kernel's tile set-up has no
    // global load of its own left on its critical path
    const int S = min(max(p.seg_count[w], 0), p.seg_cap);
    const float ext = fmaxf(p.world_size[2 * w], p.world_size[2 * w + 1]);
    *(int2*)(wsr + BRS_WS_META) = make_int2(S, __float_as_int(ext));
  }
  prep_robot_half(p, w, r, prop, wsr);
}

// ------------------------------------------------------------------------------
// The wide step kernel.  NR = camera columns per thread; BIG = walls are read from L2 / HBM
// instead of being staged in shared memory; NT_MAX / MIN_CTAS = launch bounds.
// ------------------------------------------------------------------------------
template <int NR, bool BIG, int NT_MAX, int MIN_CTAS>
__global__ void __launch_bounds__(NT_MAX, MIN_CTAS) brs_wide_kernel(const __grid_constant__ KParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31;
  const int R = p.n_robots, SC = p.seg_cap, TW = p.TW, NG = p.n_group;
  constexpr bool TABS = NT_MAX > 32;  // camera tables in shared memory (wider teams) or through L1
  const WideLayout L = make_wide_layout(TW, SC, R, NG, p.n_camera, p.cam_w, p.cam_h, !BIG, TABS);
  float4* const segtab = (float4*)(smem + L.segtab);
  float4* const gtab = (float4*)(smem + L.gtab);
  uint16_t* const gidx = (uint16_t*)(smem + L.gidx);
  int2* const gcnt = (int2*)(smem + L.gcnt);
  GroupCull* const gcull = (GroupCull*)(smem + L.gcull);
  double2* const org = (double2*)(smem + L.org);
  int4* const cflag = (int4*)(smem + L.cflag);
  int2* const meta = (int2*)(smem + L.meta);
  ColInfo* const colinfo = (ColInfo*)(smem + L.colinfo);
  Strip* const strips = (Strip*)(smem + L.strips);
  double2* const camcs_s = (double2*)(smem + L.camcs);
  uint32_t* const rowsky_s = (uint32_t*)(smem + L.rowsky);
  uint32_t* const rowgnd_s = (uint32_t*)(smem + L.rowgnd);
  const double2* const camcs = TABS ? camcs_s : p.cam_cs;
  const uint32_t* const rowsky = TABS ? rowsky_s : p.row_sky;
  const uint32_t* const rowgnd = TABS ? rowgnd_s : p.row_gnd;
  int* const cq = (int*)(smem + L.cq);
  uint64_t* const mbar = (uint64_t*)(smem + L.mbar);
  volatile int* const sched = (volatile int*)(smem + L.sched);
  const uint32_t world_bytes = (uint32_t)(5 * SC * 4), robot_bytes = (uint32_t)(R * BRS_WS_STRIDE * 8);
  const bool do_move = p.flags & 1u, do_sense = p.flags & 2u, do_cam = (p.flags & 4u) && p.n_camera > 0;
  const bool do_rays = do_sense || do_cam;

  // one thread requests a tile: its walls (unless BIG) and its robot records, one mbarrier
  auto request_tile = [&](int tn, int b, bool first) {
    const int wn = tn * TW, nn = min(TW, p.n_worlds - wn);
    const uint32_t bar = smem_u32(&mbar[b]);
    mbar_expect_tx(bar, (BIG ? 0u : nn * world_bytes) + nn * robot_bytes);
    if (!BIG) tma_bulk_g2s(smem_u32(smem + L.raw[b]), p.seg + (size_t)wn * 5 * SC, nn * world_bytes, bar);
    if (first) griddep_wait();  // the robot records are what the prep kernel writes
    tma_bulk_g2s(smem_u32(smem + L.ws[b]), p.wsg + (size_t)wn * R * BRS_WS_STRIDE, nn * robot_bytes, bar);
  };

  // ---- prologue (overlaps the prep kernel's tail under programmatic dependent launch)
  if (tid == 0) {
    mbar_init(smem_u32(&mbar[0]), 1);
    mbar_init(smem_u32(&mbar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (TABS) {
    for (int i = tid; i < p.n_camera * p.cam_w; i += NT) camcs_s[i] = p.cam_cs[i];
    for (int i = tid; i < p.cam_h; i += NT) {
      rowsky_s[i] = p.row_sky[i];
      rowgnd_s[i] = p.row_gnd[i];
    }
  }
  const PaintMap pmap = make_paint_map(p, tid, NT);
  team_sync(0, NT);
  int tile = (int)blockIdx.x;  // the first tile of a CTA is its block index; the counter hands out the rest
  // Tickets run TWO tiles ahead: the atomic for tile it+2 is issued at the start of tile it and
  // consumed at the start of tile it+1, so its round trip to L2 never stalls the team.
  unsigned ticket = 0;
  if (tid == 0) {
    ticket = atomicAdd(&p.sched[0], 1u);
    if (tile < p.n_tiles) request_tile(tile, 0, true);
  }
  griddep_wait();  // every thread: nothing below may pass the prep kernel (it reads the poses COMMIT writes)

  for (int it = 0; tile < p.n_tiles; ++it) {
    const int buf = it & 1;
    const int w0 = tile * TW, nw = min(TW, p.n_worlds - w0);
    if (tid == 0) {
      // next tile: its walls and records go to the other buffer (free since the barrier that
      // ended the previous tile)
      const int tn = (int)(gridDim.x + ticket);
      sched[buf ^ 1] = tn;
      if (tn < p.n_tiles) request_tile(tn, buf ^ 1, false);
      ticket = atomicAdd(&p.sched[0], 1u);
    }
    const uint32_t* const rawb = BIG ? p.seg + (size_t)w0 * 5 * SC : (const uint32_t*)(smem + L.raw[buf]);
    double* const wsb = (double*)(smem + L.ws[buf]);
    // ---------------- BUILD: per-world meta, collision flags, FP32 wall table ----------
    for (int i = tid; i < nw * R; i += NT) cflag[i] = make_int4(0, 0, 0, 0);
    if (tid == 0) cq[0] = 0;
    mbar_wait(smem_u32(&mbar[buf]), (uint32_t)(it >> 1) & 1u);
    for (int wl = tid; wl < nw; wl += NT) meta[wl] = *(const int2*)(wsb + (size_t)wl * R * BRS_WS_STRIDE + BRS_WS_META);
    build_segtab(p, nw, tid, NT, rawb, segtab);
    team_sync(0, NT);
    // ---------------- COLLIDE + COMMIT ----------------------------------------------------
    collide_commit_tile(p, w0, nw, tid, NT, 0, do_move, do_rays, false, rawb, segtab, wsb, cflag, meta, org, gcull, cq);
    if (do_rays) {
      team_sync(0, NT);
      // ---------------- TABLE, SENSE, PAINT -------------------------------------------------
      sense_paint_tile<NR>(p, tid, NT, lane, w0, nw, do_sense, do_cam, true, rawb, segtab, wsb, meta, org, gcull, gtab,
                           gidx, gcnt, colinfo, strips, camcs, rowsky, rowgnd, pmap);
    }
    team_sync(0, NT);  // the tile is done: its buffers may be overwritten
    tile = sched[buf ^ 1];
  }

  // self-resetting scheduler: the last CTA to leave zeroes both counters for the next launch
  if (tid == 0) {
    // (the CTA's last ticket was never consumed: depend on it, so that its atomic has been
    //  performed before this CTA counts as finished and the counters can be reset)
    if (ticket == 0xFFFFFFFFu) p.sched[1] = 0u;
    __threadfence();
    const unsigned done = atomicAdd(&p.sched[1], 1u);
    if (done == gridDim.x - 1) {
      p.sched[0] = 0u;
      p.sched[1] = 0u;
      __threadfence();
    }
  }
}

#endif  // __CUDACC__
}  // namespace brs
