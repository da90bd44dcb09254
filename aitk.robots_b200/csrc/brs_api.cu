// brs_api.cu -- C ABI (include/brs.h) over the sm_100a kernels in brs_kernels.cuh.
//
// Host side only: configuration checks, descriptor precomputation in float64
// (with the same libm the float64 reference uses), HBM buffer ownership, launch
// geometry, host<->device copies.  There is no CPU compute path in here: every
// entry point that produces results launches brs_step_kernel.
#include "../../include/brs.h"
#include "brs_kernels.cuh"
#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: ranges cost nothing unless a tool is attached
#include "brs_wide.cuh"
#include "brs_solo.cuh"
#include "brs_multi.cuh"
#include "brs_mini.cuh"

#include <sched.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace brs;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(BRS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));          \
  } while (0)

struct brs_engine {
  brs_config cfg;
  std::vector<brs_robot_desc> robots;
  std::vector<brs_range_desc> ranges;
  std::vector<brs_camera_desc> cameras;
  std::vector<brs_point_desc> lights, smells;
  int sm_count = 0;
  int cam_w = 0, cam_h = 0;
  int rays_range = 0;  // total range rays per world
  void* buf[BRS_BUF_COUNT_] = {};
  size_t stride[BRS_BUF_COUNT_] = {};  // bytes per world
  void* bound[BRS_BUF_COUNT_] = {};
  RobotDev h_robot0;      // host copies of the first robot / camera / group descriptors and row colours
  CameraDev h_camera0;    // (the solo path passes them in the kernel parameters)
  GroupDev h_group0;
  uint32_t h_sky0 = 0, h_gnd0 = 0;
  RobotDev* d_robots = nullptr;
  RangeDev* d_ranges = nullptr;
  CameraDev* d_cameras = nullptr;
  PointDev* d_lights = nullptr;
  PointDev* d_smells = nullptr;
  GroupDev* d_groups = nullptr;
  int* d_gr_idx = nullptr;
  int* d_dr_idx = nullptr;
  double2* d_camcs = nullptr;
  uint32_t* d_rowsky = nullptr;
  uint32_t* d_rowgnd = nullptr;
  int flat_rows = 1;
  unsigned int* d_sched = nullptr;
  unsigned int* d_epoch = nullptr;  // [n_tiles][2] brs_steps on the fused kernel: per-tile epochs (committed, painted)
  unsigned epoch0 = 0;              // every epoch word is <= epoch0 between launches
  unsigned long long* d_phase = nullptr;  // BRS_PHASE_TIMING builds
  unsigned long long* d_counters = nullptr;  // [2] executed-work counters (brs_step_counted)
  bool counting = false;
  bool pdl_chain = false;  // fused kernel launches use programmatic dependent launch (BRS_PDL_CHAIN)
  bool steps_off = false;  // BRS_STEPS_ONE_LAUNCH=0: brs_steps always launches once per step
  int n_group = 0, n_gr = 0, n_dr = 0;
  // launch geometry: threads per CTA, worlds per tile, camera columns per thread, ...
  int T = 288, NW = 256, NH = 1, TW = 1, NR = 1, GS = 32, SLG = 1, SLD = 1, SLL = 1, paint_rs = 1, n_tiles = 1, grid = 1, smem = 0;
  // the same kernel's shape under brs_steps (one launch, many steps): no per-step fill / drain and
  // no tail to balance, so tiles can be larger (config 2 at 4,096 worlds: TW 1 -> 4, 5.9 -> 6.9 TB/s)
  int mTW = 1, mSLG = 1, mSLD = 1, mSLL = 1, m_paint_rs = 1, m_tiles = 1, m_grid = 1, m_smem = 0, m_fused_table = 0;
  int gtasks = 0, dtasks = 0;
  bool big = false, direct_on_helper = false;
  bool split_lights = false;  // helper team: light + smell sensors, workers: range sensors (KParams::direct_on_helper == 2)
  int fused_table = 0;
  const void* kfn = nullptr;
  // the wide path (brs_wide.cuh): prep kernel + one-team-per-tile step kernel
  bool wide = false;
  double* d_ws = nullptr;  // [B][R][BRS_WS_STRIDE] robot records between the two kernels
  int wNT = 32, wTW = 1, wSLG = 1, wSLD = 1, wSLL = 1, w_paint_rs = 1, w_tiles = 1, w_grid = 1, w_smem = 0, w_fused_table = 1;
  bool w_big = false, pdl = true;
  bool prep_fused = false;  // fused kernel fed by brs_prep_kernel (two launches)
  const void* w_kfn = nullptr;
  // the solo path (brs_solo.cuh): one warp per one-robot camera world, one launch
  bool solo = false;
  int sNT = 128, s_grid = 1, s_smem = 0, s_warps_per_sm = 0, solo_interleave = 1;
  const void* s_kfn = nullptr;
  int s_ctas_per_sm = 1, f_ctas_per_sm = 1, w_ctas_per_sm = 1;
  // the solo range path (brs_scan_kernel): one robot, every range sensor on one mount, no camera
  bool scan = false;
  GroupDev h_scan_group;
  SoloRay* d_scan_rays = nullptr;
  int scan_nrays = 0, scan_SL = 1, scan_NRL = 1;
  // the mini range path (brs_mini_kernel): a team of mini_G lanes per world, small range-sensor worlds
  bool mini = false;
  int mini_G = 4, mini_NRAY = 1;
  // the solo multi path (brs_multi_kernel): several robots (or several mounts), no camera
  bool use_rmask = true;   // fused / wide: warp-built reach masks for the direct-path range sensors (BRS_RMASK)
  bool use_rlist = false;  // fused / wide: per-robot lists of reachable segments for the direct-path range sensors
  bool multi = false;
  MultiRay* d_multi_rays = nullptr;
  int multi_nrays = 0, multi_bw = 1;
  float multi_reach[8] = {};
  // brs_step_host: chunked pipeline (kernel of chunk c+1 overlaps the device->host copies of chunk c)
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_chunk[8] = {}, ev_copied = nullptr;
  int host_chunks = 4;
  // fused picture gather: peer copies of the camera output (brs_bind_peer_outputs)
  int n_peers = 0;
  void* peer_cam[7] = {};
  // seeded sensor noise (brs_set_noise); the step counter advances once per sensing step
  NoiseParams noise = {};
  unsigned long long noise_step = 0;
  // kernel parameters cached per step-flags value (everything but dt and the rebindable outputs is fixed at create)
  KParams kp_cache;
  unsigned kp_flags = 0xffffffffu;
  bool kp_wide = false, kp_msteps = false;
};

static int pow2floor(int x) {
  int p = 1;
  while (p * 2 <= x) p *= 2;
  return p;
}
static int pow2ceil(int x) {
  int p = 1;
  while (p < x) p *= 2;
  return p;
}
// NVTX range around a host-side entry point (step / sensors / gather ranges)
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

extern "C" int brs_abi_version(void) { return BRS_ABI_VERSION; }
extern "C" const char* brs_last_error(void) { return g_err.c_str(); }

extern "C" int brs_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return fail(BRS_ERR_NO_DEVICE, cudaGetErrorString(e));
  return n;
}

template <class T>
static int to_device(T** dst, const std::vector<T>& src) {
  const size_t bytes = std::max<size_t>(src.size() * sizeof(T), 16);
  CK(cudaMalloc((void**)dst, bytes));
  CK(cudaMemset(*dst, 0, bytes));
  if (!src.empty()) CK(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
  return BRS_OK;
}

static uint32_t pack_rgb(double r, double g, double b) {
  return (uint32_t)(int)r | ((uint32_t)(int)g << 8) | ((uint32_t)(int)b << 16) | 0xff000000u;
}

extern "C" int brs_create(const brs_config* cfg, brs_engine** out) {
  if (!cfg || !out) return fail(BRS_ERR_INVALID, "brs_create: null argument");
  *out = nullptr;
  if (cfg->abi_version != BRS_ABI_VERSION) return fail(BRS_ERR_INVALID, "brs_create: ABI version mismatch");
  if (cfg->n_worlds < 1) return fail(BRS_ERR_INVALID, "brs_create: n_worlds < 1");
  if (cfg->n_robots < 1 || cfg->n_robots > BRS_MAX_ROBOTS)
    return fail(BRS_ERR_INVALID, "brs_create: n_robots out of range 1..BRS_MAX_ROBOTS");
  if (cfg->seg_cap < 4 || cfg->seg_cap % 4) return fail(BRS_ERR_INVALID, "brs_create: seg_cap must be a positive multiple of 4");
  if (cfg->n_range < 0 || cfg->n_camera < 0 || cfg->n_light < 0 || cfg->n_smell < 0 || cfg->n_bulbs < 0)
    return fail(BRS_ERR_INVALID, "brs_create: negative device count");
  if (cfg->grid_w < 0 || cfg->grid_h < 0 || ((cfg->grid_w > 0) != (cfg->grid_h > 0)))
    return fail(BRS_ERR_INVALID, "brs_create: smell grid must be grid_w x grid_h > 0, or 0 x 0 for none");
  if (cfg->grid_w > 0 && cfg->n_smell > 0 && !(cfg->smell_cell_size > 0))
    return fail(BRS_ERR_INVALID, "brs_create: smell_cell_size must be > 0 when a smell grid exists");
  if (!cfg->robots) return fail(BRS_ERR_INVALID, "brs_create: robots == NULL");
  if ((cfg->n_range && !cfg->ranges) || (cfg->n_camera && !cfg->cameras) || (cfg->n_light && !cfg->lights) ||
      (cfg->n_smell && !cfg->smells))
    return fail(BRS_ERR_INVALID, "brs_create: device table missing");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
    return fail(BRS_ERR_NO_DEVICE, "brs_create: no CUDA device (this engine has no CPU path)");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(BRS_ERR_INVALID, "brs_create: bad device ordinal");
  CK(cudaSetDevice(cfg->device));

  brs_engine* e = new brs_engine();
  e->cfg = *cfg;
  const int R = cfg->n_robots;
  e->robots.assign(cfg->robots, cfg->robots + R);
  if (cfg->n_range) e->ranges.assign(cfg->ranges, cfg->ranges + cfg->n_range);
  if (cfg->n_camera) e->cameras.assign(cfg->cameras, cfg->cameras + cfg->n_camera);
  if (cfg->n_light) e->lights.assign(cfg->lights, cfg->lights + cfg->n_light);
  if (cfg->n_smell) e->smells.assign(cfg->smells, cfg->smells + cfg->n_smell);
  e->cfg.robots = e->robots.data();
  e->cfg.ranges = e->ranges.data();
  e->cfg.cameras = e->cameras.data();
  e->cfg.lights = e->lights.data();
  e->cfg.smells = e->smells.data();
  auto bad = [&](const char* m) {
    delete e;
    return fail(BRS_ERR_INVALID, m);
  };
  for (auto& d : e->ranges)
    if (d.robot < 0 || d.robot >= R || !(d.max > 0)) return bad("brs_create: bad range sensor descriptor");
  for (auto& d : e->lights)
    if (d.robot < 0 || d.robot >= R) return bad("brs_create: bad light sensor descriptor");
  for (auto& d : e->smells)
    if (d.robot < 0 || d.robot >= R) return bad("brs_create: bad smell sensor descriptor");
  for (auto& d : e->cameras) {
    if (d.robot < 0 || d.robot >= R || d.width < 4 || d.width % 4 || d.height < 1)
      return bad("brs_create: bad camera descriptor (width must be a multiple of 4)");
    if (e->cam_w == 0) {
      e->cam_w = d.width;
      e->cam_h = d.height;
    } else if (e->cam_w != d.width || e->cam_h != d.height)
      return bad("brs_create: all cameras of an engine must share one image size");
  }
  if (cfg->n_light && !cfg->n_bulbs) { /* lights without bulbs read 0: allowed */ }

  cudaDeviceProp prop;
  {
    cudaError_t ce = cudaGetDeviceProperties(&prop, cfg->device);
    if (ce != cudaSuccess) {
      delete e;
      return fail(BRS_ERR_CUDA, std::string("brs_create: cudaGetDeviceProperties: ") + cudaGetErrorString(ce));
    }
  }
  e->sm_count = prop.multiProcessorCount;

  // ---- device descriptors, float64, same libm as the float64 reference -------
  std::vector<RobotDev> rdev(R);
  for (int r = 0; r < R; ++r) {
    const brs_robot_desc& d = e->robots[r];
    const double cx[4] = {d.bbox[0], d.bbox[0], d.bbox[1], d.bbox[1]};
    const double cy[4] = {d.bbox[3], d.bbox[2], d.bbox[2], d.bbox[3]};
    for (int k = 0; k < 4; ++k) {
      // compute_boundingbox: dist = distance(0,0,cx,cy), ang = atan2(-cx, cy), swung by ang + pi/2
      const double dist = std::sqrt((0 - cx[k]) * (0 - cx[k]) + (0 - cy[k]) * (0 - cy[k]));
      const double ang = std::atan2(-cx[k], cy[k]);
      rdev[r].cox[k] = dist * std::cos(ang + 1.5707963267948966);
      rdev[r].coy[k] = dist * std::sin(ang + 1.5707963267948966);
    }
    rdev[r].radius = 0;
    for (int k = 0; k < 4; ++k) rdev[r].radius = std::fmax(rdev[r].radius, std::hypot(cx[k], cy[k]));
    rdev[r].height = d.height;
    rdev[r].vx_ramp = d.vx_ramp;
    rdev[r].vy_ramp = d.vy_ramp;
    rdev[r].va_ramp = d.va_ramp;
    rdev[r].rgba = (uint32_t)d.rgba[0] | ((uint32_t)d.rgba[1] << 8) | ((uint32_t)d.rgba[2] << 16) | ((uint32_t)d.rgba[3] << 24);
  }
  const double PI2 = 1.5707963267948966;
  std::vector<RangeDev> gdev(cfg->n_range);
  e->rays_range = 0;
  for (int i = 0; i < cfg->n_range; ++i) {
    const brs_range_desc& d = e->ranges[i];
    RangeDev& g = gdev[i];
    memset(&g, 0, sizeof(g));
    g.robot = d.robot;
    g.mx = d.dist_from_center * std::cos(d.dir_from_center + PI2);
    g.my = d.dist_from_center * std::sin(d.dir_from_center + PI2);
    g.max = d.max;
    int n = 0;
    if (d.width != 0) {
      // inclusive float range -w/2, 0, +w/2 of the reference's fan
      double cur = -d.width / 2;
      while (cur <= d.width / 2 && n < 3) {
        g.cr[n] = std::cos(d.a - cur);
        g.sr[n] = std::sin(d.a - cur);
        ++n;
        cur += d.width / 2;
      }
    } else {
      g.cr[0] = std::cos(d.a);
      g.sr[0] = std::sin(d.a);
      n = 1;
    }
    g.n_rays = n;
    g.group = -1;
    e->ranges[i].n_rays = n;
    e->rays_range += n;
  }
  std::vector<CameraDev> cdev(cfg->n_camera);
  std::vector<double2> camcs((size_t)cfg->n_camera * e->cam_w);
  for (int i = 0; i < cfg->n_camera; ++i) {
    const brs_camera_desc& d = e->cameras[i];
    CameraDev& c = cdev[i];
    memset(&c, 0, sizeof(c));
    c.robot = d.robot;
    c.width = d.width;
    c.height = d.height;
    c.mx = d.dist_from_center * std::cos(d.dir_from_center + PI2);
    c.my = d.dist_from_center * std::sin(d.dir_from_center + PI2);
    c.max_range = d.max_range;
    c.color_fade = d.color_fade;
    c.size_fade = d.size_fade;
    for (int k = 0; k < d.width; ++k) {
      const double ang = (double)k / (double)d.width * d.fov - d.fov / 2;
      camcs[(size_t)i * e->cam_w + k] = make_double2(std::cos(ang), std::sin(ang));
    }
  }
  auto points = [&](const std::vector<brs_point_desc>& src) {
    std::vector<PointDev> v(src.size());
    for (size_t i = 0; i < src.size(); ++i) {
      v[i].robot = src[i].robot;
      v[i]._pad = 0;
      v[i].mx = src[i].dist_from_center * std::cos(src[i].dir_from_center + PI2);
      v[i].my = src[i].dist_from_center * std::sin(src[i].dir_from_center + PI2);
    }
    return v;
  };
  std::vector<PointDev> ldev = points(e->lights), sdev = points(e->smells);
  // ---- origin groups: rays that leave one mount point of one robot share a
  // per-(origin, segment) table.  Every camera gets one; range sensors join when
  // the mount carries a camera or at least BRS_GROUP_MIN rays in total.
  std::vector<GroupDev> groups;
  auto find_group = [&](int robot, double mx, double my) {
    for (size_t k = 0; k < groups.size(); ++k)
      if (groups[k].robot == robot && groups[k].mx == mx && groups[k].my == my) return (int)k;
    return -1;
  };
  auto add_group = [&](int robot, double mx, double my) {
    int k = find_group(robot, mx, my);
    if (k < 0) {
      GroupDev g;
      memset(&g, 0, sizeof(g));
      g.robot = robot; g.mx = mx; g.my = my;
      groups.push_back(g);
      k = (int)groups.size() - 1;
    }
    return k;
  };
  for (auto& c : cdev) c.group = add_group(c.robot, c.mx, c.my);
  {
    const int group_min = env_int("BRS_GROUP_MIN", 4);
    for (auto& g : gdev) {
      int k = find_group(g.robot, g.mx, g.my);
      if (k < 0) {
        int rays = 0;  // range rays leaving this mount
        for (auto& h : gdev)
          if (h.robot == g.robot && h.mx == g.mx && h.my == g.my) rays += h.n_rays;
        if (rays >= group_min) k = add_group(g.robot, g.mx, g.my);
      }
      g.group = k;
    }
  }
  // what a group's rays can reach: longest range (widened) and, when every direction fits an
  // arc narrower than 180 degrees, the cone of that arc (robot frame, widened by a margin)
  auto set_reach = [&](GroupDev& g, const std::vector<double>& ang, double range) {
    const double PI = 3.14159265358979323846;
    g.range = (float)(range * (1.0 + 1e-4) + 1e-3);
    g.has_cone = 0;
    // smallest arc containing all directions: try every direction as the arc's start
    double best_w = 1e9, best_lo = 0;
    for (double a0 : ang) {
      double w = 0;
      for (double a : ang) {
        double d = std::fmod(a - a0, 2 * PI);
        if (d < 0) d += 2 * PI;
        w = std::fmax(w, d);
      }
      if (w < best_w) { best_w = w; best_lo = a0; }
    }
    const double margin = 2e-3;
    if (!ang.empty() && best_w + 2 * margin < PI - 0.02 && !env_int("BRS_NO_CONE", 0)) {
      const double lo = best_lo - margin, hi = best_lo + best_w + margin, ax = 0.5 * (lo + hi);
      g.has_cone = 1;
      g.lo_c = (float)std::cos(lo); g.lo_s = (float)std::sin(lo);
      g.hi_c = (float)std::cos(hi); g.hi_s = (float)std::sin(hi);
      g.ax_c = (float)std::cos(ax); g.ax_s = (float)std::sin(ax);
    }
    if (env_int("BRS_NO_RANGE_CULL", 0)) g.range = 3.0e18f;
  };
  // what each group's rays can reach: longest range and the cone of directions (robot frame)
  {
    for (size_t k = 0; k < groups.size(); ++k) {
      std::vector<double> ang;
      double range = 0;
      for (int i = 0; i < cfg->n_camera; ++i)
        if (cdev[i].group == (int)k) {
          const brs_camera_desc& d = e->cameras[i];
          // every column direction: the fan is contiguous, its two ends alone would let a
          // wide fan (> 180 deg) pass for the short arc through the robot's back
          for (int q = 0; q < d.width; ++q) ang.push_back((double)q / (double)d.width * d.fov - d.fov / 2);
          range = std::fmax(range, d.max_range);
        }
      for (int i = 0; i < cfg->n_range; ++i)
        if (gdev[i].group == (int)k) {
          for (int q = 0; q < gdev[i].n_rays; ++q) ang.push_back(std::atan2(gdev[i].sr[q], gdev[i].cr[q]));
          range = std::fmax(range, gdev[i].max);
        }
      GroupDev& g = groups[k];
      g.cam_rays = g.range_rays = 0;
      for (int i = 0; i < cfg->n_camera; ++i)
        if (cdev[i].group == (int)k) g.cam_rays += e->cameras[i].width;
      for (int i = 0; i < cfg->n_range; ++i)
        if (gdev[i].group == (int)k) g.range_rays += gdev[i].n_rays;
      set_reach(g, ang, range);
    }
  }
  // the solo range path's own view: ALL range rays of a lone robot leave one mount (whatever
  // BRS_GROUP_MIN made of them above) -> one shared-origin group and a flat ray table
  std::vector<SoloRay> scan_rays;
  bool scan_shape = R == 1 && cfg->n_camera == 0 && cfg->n_light == 0 && cfg->n_smell == 0 && cfg->n_range >= 1 &&
                    cfg->seg_cap <= 65532;
  if (scan_shape) {
    std::vector<double> ang;
    double range = 0;
    for (int i = 0; i < cfg->n_range && scan_shape; ++i) {
      if (gdev[i].mx != gdev[0].mx || gdev[i].my != gdev[0].my) scan_shape = false;
      for (int q = 0; q < gdev[i].n_rays; ++q) {
        SoloRay sr;
        sr.cr = gdev[i].cr[q]; sr.sr = gdev[i].sr[q]; sr.max = gdev[i].max; sr.sensor = i; sr._pad = 0;
        scan_rays.push_back(sr);
        ang.push_back(std::atan2(gdev[i].sr[q], gdev[i].cr[q]));
      }
      range = std::fmax(range, gdev[i].max);
    }
    if (scan_rays.size() > 256) scan_shape = false;
    if (scan_shape) {
      memset(&e->h_scan_group, 0, sizeof(GroupDev));
      e->h_scan_group.robot = 0; e->h_scan_group.mx = gdev[0].mx; e->h_scan_group.my = gdev[0].my;
      e->h_scan_group.cam_rays = 0; e->h_scan_group.range_rays = (int)scan_rays.size();
      set_reach(e->h_scan_group, ang, range);
      e->scan_nrays = (int)scan_rays.size();
    }
  }
  // the solo multi path's view: every range ray of the world, robot by robot, and how far from
  // its robot's centre a ray can reach (mount offset + range, widened)
  std::vector<MultiRay> multi_rays;
  const bool multi_shape = cfg->n_camera == 0 && cfg->seg_cap + 4 * R <= 256 && cfg->n_light * cfg->n_bulbs <= 1024 &&
                           (cfg->n_range + cfg->n_light + cfg->n_smell) > 0;
  {
    for (int r = 0; r < R; ++r) {
      double reach = 0;
      for (int i = 0; i < cfg->n_range; ++i) {
        if (gdev[i].robot != r) continue;
        for (int q = 0; q < gdev[i].n_rays && multi_shape; ++q) {
          MultiRay mr;
          mr.cr = gdev[i].cr[q]; mr.sr = gdev[i].sr[q]; mr.max = gdev[i].max; mr.mx = gdev[i].mx; mr.my = gdev[i].my;
          mr.sensor = i; mr.robot = r;
          multi_rays.push_back(mr);
        }
        reach = std::fmax(reach, std::hypot(gdev[i].mx, gdev[i].my) + gdev[i].max);
      }
      e->multi_reach[r] = (float)(reach * (1.0 + 1e-3) + 0.1);
    }
    e->multi_nrays = (int)multi_rays.size();
  }
  std::vector<int> gr_idx, dr_idx;
  for (int i = 0; i < cfg->n_range; ++i) (gdev[i].group >= 0 ? gr_idx : dr_idx).push_back(i);
  e->n_group = (int)groups.size();
  e->n_gr = (int)gr_idx.size();
  e->n_dr = (int)dr_idx.size();
  std::vector<uint32_t> rowsky(std::max(e->cam_h, 1)), rowgnd(std::max(e->cam_h, 1));
  for (int j = 0; j < e->cam_h; ++j) {
    const double horizon = e->cam_h / 2.0;
    const double dist = std::fmax(std::fmin(std::fabs(j - horizon) / horizon, 1.0), 0.0);
    rowsky[j] = pack_rgb(cfg->sky_rgb[0] + cfg->sky_grad[0] * dist, cfg->sky_rgb[1] + cfg->sky_grad[1] * dist,
                         cfg->sky_rgb[2] + cfg->sky_grad[2] * dist);
    rowgnd[j] = pack_rgb(cfg->ground_rgb[0] + cfg->ground_grad[0] * dist, cfg->ground_rgb[1] + cfg->ground_grad[1] * dist,
                         cfg->ground_rgb[2] + cfg->ground_grad[2] * dist);
  }
  e->flat_rows = 1;
  for (int j = 1; j < e->cam_h; ++j)
    if (rowsky[j] != rowsky[0] || rowgnd[j] != rowgnd[0]) e->flat_rows = 0;
  int rc;
  memset(&e->h_robot0, 0, sizeof(RobotDev)); memset(&e->h_camera0, 0, sizeof(CameraDev)); memset(&e->h_group0, 0, sizeof(GroupDev));
  if (!rdev.empty()) e->h_robot0 = rdev[0];
  if (!cdev.empty()) e->h_camera0 = cdev[0];
  if (!groups.empty()) e->h_group0 = groups[0];
  e->h_sky0 = rowsky[0]; e->h_gnd0 = rowgnd[0];
  if ((rc = to_device(&e->d_robots, rdev)) || (rc = to_device(&e->d_ranges, gdev)) ||
      (rc = to_device(&e->d_cameras, cdev)) || (rc = to_device(&e->d_lights, ldev)) ||
      (rc = to_device(&e->d_smells, sdev)) || (rc = to_device(&e->d_camcs, camcs)) ||
      (rc = to_device(&e->d_groups, groups)) || (rc = to_device(&e->d_gr_idx, gr_idx)) ||
      (rc = to_device(&e->d_dr_idx, dr_idx)) ||
      (rc = to_device(&e->d_rowsky, rowsky)) || (rc = to_device(&e->d_rowgnd, rowgnd)) ||
      (scan_shape && (rc = to_device(&e->d_scan_rays, scan_rays))) ||
      (multi_shape && (rc = to_device(&e->d_multi_rays, multi_rays)))) {
    brs_destroy(e);
    return rc;
  }

  // ---- HBM buffers ----------------------------------------------------------
  size_t* st = e->stride;
  st[BRS_BUF_SEG] = (size_t)5 * cfg->seg_cap * 4;
  st[BRS_BUF_SEG_COUNT] = 4;
  st[BRS_BUF_WORLD_SIZE] = 8;
  st[BRS_BUF_BULBS] = (size_t)cfg->n_bulbs * 16;
  st[BRS_BUF_GRID] = (size_t)cfg->grid_w * cfg->grid_h * 4;
  st[BRS_BUF_POSE] = st[BRS_BUF_VEL] = st[BRS_BUF_TARGET_VEL] = (size_t)R * 24;
  st[BRS_BUF_STALLED] = (size_t)R;
  st[BRS_BUF_RANGE] = (size_t)cfg->n_range * 4;
  st[BRS_BUF_CAMERA] = (size_t)cfg->n_camera * e->cam_w * e->cam_h * 4;
  st[BRS_BUF_LIGHT] = (size_t)cfg->n_light * 4;
  st[BRS_BUF_SMELL] = (size_t)cfg->n_smell * 4;
  for (int b = 0; b < BRS_BUF_COUNT_; ++b) {
    const size_t bytes = std::max<size_t>(st[b] * (size_t)cfg->n_worlds, 16);
    cudaError_t ce = cudaMalloc(&e->buf[b], bytes);
    if (ce == cudaSuccess) ce = cudaMemset(e->buf[b], 0, bytes);
    if (ce != cudaSuccess) {
      brs_destroy(e);
      return fail(BRS_ERR_CUDA, std::string("brs_create: HBM allocation failed: ") + cudaGetErrorString(ce));
    }
  }

  // ---- launch geometry -------------------------------------------------------
  // Tiles of TW consecutive worlds per CTA pass; TW is sized so that a tile offers
  // about one ray task per thread within a shared-memory budget that keeps
  // several CTAs resident per SM (different CTAs sit in different phases, which
  // overlaps the FP32 search of one with the HBM stores of another).
  {
    cudaError_t ce = cudaMalloc((void**)&e->d_sched, 16);
    if (ce == cudaSuccess) ce = cudaMemset(e->d_sched, 0, 16);
    if (ce != cudaSuccess) {
      brs_destroy(e);
      return fail(BRS_ERR_CUDA, std::string("brs_create: scheduler allocation failed: ") + cudaGetErrorString(ce));
    }
  }
#ifdef BRS_PHASE_TIMING
  if (cudaMalloc((void**)&e->d_phase, 128) == cudaSuccess) cudaMemset(e->d_phase, 0, 128);
#endif
  const int ncols = cfg->n_camera * e->cam_w;
  e->NR = env_int("BRS_NR", ncols >= 64 ? 2 : 1);
  if (e->NR != 1 && e->NR != 2 && e->NR != 4) e->NR = 1;
  if (e->cam_w % std::max(e->NR, 1)) e->NR = 1;
  const int gtasks = ncols / e->NR + e->n_gr;
  const int dtasks = e->n_dr + cfg->n_light + cfg->n_smell;
  e->gtasks = gtasks;
  e->dtasks = dtasks;
  const int tasks = std::max(gtasks + dtasks, 1);
  // (direct-path range sensors search per-robot lists of the segments within reach: two byte
  //  arrays behind the kernel's own shared-memory layout)
  e->use_rlist = e->n_dr > 0 && cfg->seg_cap + 4 * R <= 256 && env_int("BRS_RLIST", 0) != 0;  // (measured on config 3: building the lists costs what the shorter searches save; opt-in)
  auto rl_extra = [&](int tw) {
    return e->use_rlist ? align_up(tw * R * (cfg->seg_cap + 4 * R), 16) + align_up(tw * R * 4, 16) : 0;
  };
  auto smem_shape = [&](int tw, bool big) {
    return make_layout(tw, cfg->seg_cap, R, e->n_group, cfg->n_camera, e->cam_w, e->cam_h, !big).total + rl_extra(tw);
  };
  // small shape (128 workers, 4 CTAs/SM, TMA-staged walls) unless one world's footprint leaves
  // room for <= 2 such CTAs; the big shape (256 workers) reads walls from L2 instead of staging
  const int smem_sm = (int)prop.sharedMemPerMultiprocessor - 4096;
  e->big = smem_shape(1, false) > smem_sm / 3;
  e->big = env_int("BRS_BIG", e->big ? 1 : 0) != 0;
  auto smem_for = [&](int tw) { return smem_shape(tw, e->big); };
  // split of the CTA between the helper team (Robot.step: wall table, collision cull over all
  // walls) and the workers (rays): in proportion to their work per world, in warps.  Worlds
  // with many walls and few rays get more helpers, camera worlds the minimum of one warp.
  const int T_total = e->big ? BRS_BIG_THREADS : BRS_SMALL_THREADS;
  {
    const double S = cfg->seg_cap + 4.0 * (R - 1);
    const double grays = (double)ncols, rrays = (double)e->rays_range, lrays = (double)cfg->n_light * cfg->n_bulbs;
    const double grp_rays = grays + (e->n_gr ? rrays : 0.0), dir_rays = (e->n_gr ? 0.0 : rrays) + lrays;
    const double helper_w = R * (cfg->seg_cap * 40.0 + 1500.0);
    const double worker_w = grp_rays * S * 6.0 + dir_rays * S * 17.0 + (grays + rrays + lrays) * 250.0 +
                            e->n_group * S * 45.0 + (double)ncols * e->cam_h * 1.5;
    // camera worlds with a few direct-path sensors: the (otherwise mostly idle) helper team
    // runs those sensors; per-warp it must stay clearly lighter than the workers
    const double direct_w = dir_rays * S * 17.0 + (cfg->n_light + cfg->n_smell + e->n_dr) * 250.0;
    e->direct_on_helper = ncols > 0 && dtasks > 0 &&
                          (helper_w + direct_w) * (T_total / 32 - 1) < 0.6 * (worker_w - direct_w);
    e->direct_on_helper = env_int("BRS_DIRECT_ON_HELPER", e->direct_on_helper ? 1 : 0) != 0 && dtasks > 0;
    int nh = (int)std::lround(T_total / 32.0 * helper_w / (helper_w + worker_w));
    nh = std::max(1, std::min(nh, T_total / 32 - 1));
    // multi-robot worlds whose range searches run over reach masks (sense_direct): the masks
    // halve the workers' share, and one helper warp stepping R robots per world (pair
    // collisions, list-order replay) then gates them -- measured on config 3: 0.196 -> 0.185 ms
    // per step, 0.200 -> 0.172 under brs_steps, with two
    if (R >= 2 && env_int("BRS_RMASK", 1) != 0 && !e->n_gr && cfg->seg_cap + 4 * R <= 32 && e->n_dr >= 8 * R && !e->big)
      nh = std::max(nh, 2);
    // ... and those helper warps still wait for the workers most of the time: they take the
    // light and smell sensors of the tile they just committed
    e->split_lights = R >= 2 && nh >= 2 && !e->direct_on_helper && (cfg->n_light + cfg->n_smell) > 0 && e->n_dr > 0 &&
                      env_int("BRS_SPLIT_LIGHTS", 1) != 0;
    nh = env_int("BRS_NH", nh);
    nh = std::max(1, std::min(nh, T_total / 32 - 1));
    e->NH = nh;
    e->NW = T_total - 32 * nh;
  }
  const int budget = env_int("BRS_SMEM_BUDGET", e->big ? smem_sm / BRS_BIG_CTAS : smem_sm / BRS_SMALL_CTAS);
  // worlds per tile: as many as the budget allows (amortises barriers, fills the lanes of
  // the per-robot phases) while keeping >= ~6 tiles per resident CTA for the dynamic scheduler; among
  // those prefer the count whose ray tasks fill whole passes of the workers
  // tiles per resident CTA slot: ~6 for worlds with real work (tail balance), ~2 for tiny
  // worlds whose tile time is a latency chain rather than work (estimate in lane-instructions)
  const int slots = e->sm_count * (e->big ? BRS_BIG_CTAS : BRS_SMALL_CTAS);
  const double rays = (double)ncols + e->rays_range + (double)cfg->n_light * cfg->n_bulbs;
  const double nseg = cfg->seg_cap + 4.0 * (R - 1);
  const double work = (4.0 * R + rays) * nseg * 12.0 + (double)ncols * e->cam_h * 1.5 + rays * 200.0 + R * 1500.0;
  const int per_slot = work < 20000.0 ? 2 : 6;
  int tw_max = std::min(64, std::max(1, cfg->n_worlds / (slots * per_slot)));
  while (tw_max > 1 && smem_for(tw_max) > budget) --tw_max;
  auto pick_tw = [&](int tmax) {
    int pick = tmax;
    double best = -1;
    for (int t = tmax; t >= 1; --t) {
      const int n = t * tasks, passes = (n + e->NW - 1) / e->NW;
      const double eff = (double)n / ((double)passes * e->NW);
      if (eff > best + 0.05) { best = eff; pick = t; }
    }
    return pick;
  };
  int tw = pick_tw(tw_max);
  tw = env_int("BRS_TW", tw);
  tw = std::max(1, std::min(tw, cfg->n_worlds));
  // brs_steps: ~1.5 tiles per resident CTA are enough (the two teams need two tiles in flight)
  int mtw_max = std::min(64, std::max(1, (int)(cfg->n_worlds / (slots * 1.5))));
  while (mtw_max > 1 && smem_for(mtw_max) > budget) --mtw_max;
  int mtw = std::max(tw, pick_tw(mtw_max));
  if (getenv("BRS_TW")) mtw = tw;
  mtw = env_int("BRS_STEPS_TW", mtw);
  mtw = std::max(1, std::min(mtw, cfg->n_worlds));
  if ((size_t)smem_for(mtw) > prop.sharedMemPerBlockOptin) mtw = tw;
  e->TW = tw;
  e->smem = smem_for(tw);
  if ((size_t)e->smem > prop.sharedMemPerBlockOptin) {
    brs_destroy(e);
    return fail(BRS_ERR_UNSUPPORTED, "brs_create: world too large for one CTA's shared memory (seg_cap / camera width / groups)");
  }
  e->n_tiles = (cfg->n_worlds + tw - 1) / tw;
  e->T = e->NW + 32 * e->NH;
  e->GS = std::min(32 * e->NH, std::max(8, pow2ceil(cfg->seg_cap)));  // helper threads per (world, robot)
  const int dteam = e->direct_on_helper ? 32 * e->NH : e->NW;  // threads that run the direct path
  auto lanes_for = [&](int t, int& slg, int& sld, int& sll, int& prs) {
    slg = gtasks ? std::min(32, pow2floor(std::max(e->NW / (gtasks * t), 1))) : 1;
    sld = e->n_dr ? std::min(32, pow2floor(std::max(dteam / (e->n_dr * t), 1))) : 1;
    sll = cfg->n_light ? std::min(32, pow2floor(std::max((e->split_lights ? 32 * e->NH : dteam) / (cfg->n_light * t), 1))) : 1;
    const int Q = std::max(e->cam_w / 4, 1);
    const int slots = std::max(e->NW / std::min(Q, e->NW), 1);
    const int nislot = std::max(1, std::min(t * std::max(cfg->n_camera, 1), slots));
    prs = std::max(1, slots / nislot);
  };
  lanes_for(tw, e->SLG, e->SLD, e->SLL, e->paint_rs);
  e->mTW = mtw;
  e->m_smem = smem_for(mtw);
  e->m_tiles = (cfg->n_worlds + mtw - 1) / mtw;
  lanes_for(mtw, e->mSLG, e->mSLD, e->mSLL, e->m_paint_rs);
  if (e->NR == 4) e->NR = 2;  // (4 columns per thread measured slower than 2)
  const void* kfn = e->big ? (e->NR == 2 ? (const void*)brs_step_kernel<2, true> : (const void*)brs_step_kernel<1, true>)
                           : (e->NR == 2 ? (const void*)brs_step_kernel<2, false> : (const void*)brs_step_kernel<1, false>);
  e->kfn = kfn;
  cudaError_t ce = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, std::max(e->smem, e->m_smem));
  int nb = 0, mnb = 0;
  if (ce == cudaSuccess) ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kfn, e->T, e->smem);
  if (ce == cudaSuccess) ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&mnb, kfn, e->T, e->m_smem);
  if (ce == cudaSuccess && mnb < 1) {  // (cannot happen within the budget; fall back to the single-step shape)
    e->mTW = e->TW; e->m_smem = e->smem; e->m_tiles = e->n_tiles; mnb = nb;
    e->mSLG = e->SLG; e->mSLD = e->SLD; e->mSLL = e->SLL; e->m_paint_rs = e->paint_rs;
  }
  if (ce != cudaSuccess || nb < 1) {
    brs_destroy(e);
    return fail(BRS_ERR_CUDA, std::string("brs_create: kernel does not fit: ") + cudaGetErrorString(ce));
  }
  nb = std::min(nb, env_int("BRS_MAX_CTAS_PER_SM", 32));
  e->f_ctas_per_sm = nb;
  e->grid = std::max(1, std::min(e->n_tiles, nb * e->sm_count));
  e->fused_table = env_int("BRS_FUSED_TABLE", e->TW * e->n_group >= e->NW / 32 ? 1 : 0);
  mnb = std::min(mnb, env_int("BRS_MAX_CTAS_PER_SM", 32));
  e->m_grid = std::max(1, std::min(e->m_tiles, mnb * e->sm_count));
  e->m_fused_table = env_int("BRS_FUSED_TABLE", e->mTW * e->n_group >= e->NW / 32 ? 1 : 0);
  {
    // per-tile epoch words of brs_steps (committed, painted): all zero = epoch0
    cudaError_t ee = cudaMalloc((void**)&e->d_epoch, (size_t)std::max(e->m_tiles, 1) * 8);
    if (ee == cudaSuccess) ee = cudaMemset(e->d_epoch, 0, (size_t)std::max(e->m_tiles, 1) * 8);
    if (ee != cudaSuccess) {
      brs_destroy(e);
      return fail(BRS_ERR_CUDA, std::string("brs_create: epoch words: ") + cudaGetErrorString(ee));
    }
  }

  // ---- the wide path: launch geometry ------------------------------------------
  // One team (CTA) per tile, no helper warps.  A one-warp team wherever a tile of worlds fits
  // ~1/20 of the SM's shared memory: 20+ independent warps per SM, synchronised with
  // __syncwarp only.  Worlds too large for that get fewer, wider teams (so that the SM still
  // holds ~20 warps) and, beyond half the SM's shared memory, read their walls from L2 / HBM.
  {
    // (one-warp teams read the camera tables through L1: see make_wide_layout)
    auto wsmem_nt = [&](int tw, bool big, int nt_) {
      return make_wide_layout(tw, cfg->seg_cap, R, e->n_group, cfg->n_camera, e->cam_w, e->cam_h, !big, nt_ > 32).total + rl_extra(tw);
    };
    auto wsmem = [&](int tw, bool big) { return wsmem_nt(tw, big, 32); };
    const int target_warps = env_int("BRS_WWARPS", 20);
    int nt = 32, wtw = 1;
    bool wbig = false;
    if (wsmem(1, false) <= smem_sm / 16) {
      // worlds per tile: fill the warp's lanes with ray tasks, within the budget, keeping >= 4
      // tiles per resident team for the dynamic scheduler
      const int wbudget = std::max(smem_sm / target_warps, wsmem(1, false));
      const int cap_tiles = std::max(1, cfg->n_worlds / (e->sm_count * target_warps * 4));
      int tw_max = std::min(32, cap_tiles);
      while (tw_max > 1 && wsmem(tw_max, false) > wbudget) --tw_max;
      double best = -1;
      for (int t = 1; t <= tw_max; ++t) {
        const int n = t * tasks, passes = (n + 31) / 32;
        const double eff = (double)n / (passes * 32.0);
        if (eff > best + 0.04) { best = eff; wtw = t; }
      }
    } else {
      wbig = wsmem(1, false) > smem_sm / 2;
      const int per_sm = std::max(1, smem_sm / std::max(wsmem_nt(1, wbig, 64), 1));
      nt = 32 * std::min(8, pow2ceil((target_warps + per_sm - 1) / per_sm));
    }
    wbig = env_int("BRS_WBIG", wbig ? 1 : 0) != 0;
    nt = env_int("BRS_WNT", nt);
    nt = std::max(32, std::min(256, nt / 32 * 32));
    wtw = env_int("BRS_WTW", wtw);
    wtw = std::max(1, std::min(wtw, cfg->n_worlds));
    while (wtw > 1 && (size_t)wsmem_nt(wtw, wbig, nt) > prop.sharedMemPerBlockOptin) --wtw;
    e->wNT = nt; e->wTW = wtw; e->w_big = wbig;
    e->w_smem = wsmem_nt(wtw, wbig, nt);
    e->w_tiles = (cfg->n_worlds + wtw - 1) / wtw;
    e->wSLG = gtasks ? std::min(32, pow2floor(std::max(nt / (gtasks * wtw), 1))) : 1;
    e->wSLD = e->n_dr ? std::min(32, pow2floor(std::max(nt / (e->n_dr * wtw), 1))) : 1;
    e->wSLL = cfg->n_light ? std::min(32, pow2floor(std::max(nt / (cfg->n_light * wtw), 1))) : 1;
    {
      const int Q = std::max(e->cam_w / 4, 1);
      const int slots = std::max(nt / std::min(Q, nt), 1);
      const int nislot = std::max(1, std::min(wtw * std::max(cfg->n_camera, 1), slots));
      e->w_paint_rs = std::max(1, slots / nislot);
    }
    e->w_fused_table = env_int("BRS_FUSED_TABLE", wtw * e->n_group >= nt / 32 ? 1 : 0);
    const bool nr2 = e->NR == 2;
    const void* wk;
    // register caps by team size: ~20 warps per SM for teams of 1, 2 and 4 warps, 16 for 8
#define BRS_WK(BIGF, NTM, CT) (nr2 ? (const void*)brs_wide_kernel<2, BIGF, NTM, CT> : (const void*)brs_wide_kernel<1, BIGF, NTM, CT>)
    if (wbig)
      wk = BRS_WK(true, 256, 2);
    else if (nt <= 32)
      wk = BRS_WK(false, 32, BRS_WIDE_WARP_CTAS);
    else if (nt <= 64)
      wk = BRS_WK(false, 64, 10);
    else if (nt <= 128)
      wk = BRS_WK(false, 128, 5);
    else
      wk = BRS_WK(false, 256, 2);
#undef BRS_WK
    e->w_kfn = wk;
    bool wide_ok = (size_t)e->w_smem <= prop.sharedMemPerBlockOptin;
    int wnb = 0;
    if (wide_ok) {
      ce = cudaFuncSetAttribute(wk, cudaFuncAttributeMaxDynamicSharedMemorySize, e->w_smem);
      if (ce == cudaSuccess) ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&wnb, wk, nt, e->w_smem);
      wide_ok = ce == cudaSuccess && wnb >= 1;
      if (!wide_ok) cudaGetLastError();
    }
    e->w_ctas_per_sm = std::max(wnb, 1);
    e->w_grid = std::max(1, std::min(e->w_tiles, std::max(wnb, 1) * e->sm_count));
    // which path: the fused kernel overlaps Robot.step with the rays inside a CTA, which pays
    // when a launch has few tiles per CTA (config 2: 4096 picture-bound worlds); the wide path
    // needs enough worlds to fill every SM with independent teams several times over
    // (measured, profiles/r2: the wide path never beat the fused kernel -- config 4 0.60 vs 0.53 ms,
    //  config 3 0.26 vs 0.20 ms -- so it is opt-in; the solo path below took over its role)
    e->wide = wide_ok && env_int("BRS_WIDE", 0) != 0;
    e->pdl = env_int("BRS_PDL", 1) != 0;
    e->prep_fused = wide_ok && !e->wide && env_int("BRS_PREP", 0) != 0;
    if (wide_ok) {
      const size_t bytes = (size_t)cfg->n_worlds * R * BRS_WS_STRIDE * 8 + 64;
      ce = cudaMalloc((void**)&e->d_ws, bytes);
      if (ce == cudaSuccess) ce = cudaMemset(e->d_ws, 0, bytes);
      if (ce != cudaSuccess) {
        brs_destroy(e);
        return fail(BRS_ERR_CUDA, std::string("brs_create: robot record allocation failed: ") + cudaGetErrorString(ce));
      }
    }
  }
  // ---- the solo path: launch geometry ------------------------------------------
  // One robot carrying one camera of 32 / 64 / 128 columns and nothing else, flat sky and
  // ground: a warp per world (brs_solo.cuh).  Needs enough worlds to give every resident warp
  // a few; smaller batches keep the fused kernel, which splits one world over a CTA.
  {
    const int w = e->cam_w, nr = w / 32;
    const bool shape_ok = R == 1 && cfg->n_camera == 1 && cfg->n_range == 0 && cfg->n_light == 0 && cfg->n_smell == 0 &&
                          e->n_group == 1 && w % 32 == 0 && (nr == 1 || nr == 2 || nr == 4) && e->cam_h <= 16384 &&
                          cfg->seg_cap <= 1024;
    if (shape_ok) {
      const SoloLayout SL = make_solo_layout(cfg->seg_cap);
      const int variant = env_int("BRS_SOLO_CTAS", 6);  // resident 128-thread CTAs per SM the register cap is set for
      const void* sk = nullptr;
#define BRS_SK(CT) (nr == 1 ? (const void*)brs_solo_kernel<1, 128, CT> : nr == 2 ? (const void*)brs_solo_kernel<2, 128, CT> : (const void*)brs_solo_kernel<4, 128, CT>)
      if (variant >= 8) sk = BRS_SK(8);
      else if (variant == 7) sk = BRS_SK(7);
      else if (variant == 5) sk = BRS_SK(5);
      else if (variant <= 4) sk = BRS_SK(4);
      else sk = BRS_SK(6);
#undef BRS_SK
      e->sNT = 128;
      e->s_smem = SL.per_warp * (e->sNT / 32);
      e->s_kfn = sk;
      bool ok = (size_t)e->s_smem <= prop.sharedMemPerBlockOptin;
      int snb = 0;
      if (ok) {
        ce = cudaFuncSetAttribute(sk, cudaFuncAttributeMaxDynamicSharedMemorySize, e->s_smem);
        if (ce == cudaSuccess) ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&snb, sk, e->sNT, e->s_smem);
        ok = ce == cudaSuccess && snb >= 1;
        if (!ok) cudaGetLastError();
      }
      if (ok) {
        snb = std::min(snb, env_int("BRS_SOLO_MAX_CTAS", 32));
        e->s_warps_per_sm = snb * (e->sNT / 32);
        e->s_ctas_per_sm = snb;
        const int want = (cfg->n_worlds + (e->sNT / 32) - 1) / (e->sNT / 32);  // at least one world per warp
        e->s_grid = std::max(1, std::min(want, snb * e->sm_count));
        const bool auto_solo = cfg->n_worlds >= env_int("BRS_SOLO_MIN_WORLDS", 4 * e->s_warps_per_sm * e->sm_count);
        e->solo = env_int("BRS_SOLO", auto_solo ? 1 : 0) != 0;
        e->solo_interleave = env_int("BRS_SOLO_INTERLEAVE", 1) != 0;
        if (e->solo) { e->wide = false; e->prep_fused = false; }
      }
    }
  }
  // ---- the solo range path: launch geometry -------------------------------------
  if (scan_shape && !e->solo) {
    const int K = e->scan_nrays;
    int nrl = 1, sl = 1;
    if (K >= 32) {
      const int per_lane = (K + 31) / 32;
      nrl = per_lane <= 1 ? 1 : per_lane <= 2 ? 2 : per_lane <= 4 ? 4 : 8;
    } else {
      sl = 32 / pow2ceil(K);
    }
    sl = std::max(1, std::min(32, pow2floor(env_int("BRS_SCAN_SL", sl))));
    if (sl > 1) nrl = std::max(nrl, (K * sl + 31) / 32 <= 1 ? 1 : (K * sl + 31) / 32 <= 2 ? 2 : (K * sl + 31) / 32 <= 4 ? 4 : 8);
    e->scan_SL = sl; e->scan_NRL = nrl;
    const int variant = env_int("BRS_SOLO_CTAS", nrl >= 8 ? 4 : 6);  // (8 rays per lane need > 80 registers)
    const void* sk = nullptr;
#define BRS_SCK(CT) (nrl == 1 ? (const void*)brs_scan_kernel<1, 128, CT> : nrl == 2 ? (const void*)brs_scan_kernel<2, 128, CT> : nrl == 4 ? (const void*)brs_scan_kernel<4, 128, CT> : (const void*)brs_scan_kernel<8, 128, CT>)
    if (variant >= 8) sk = BRS_SCK(8);
    else if (variant <= 4) sk = BRS_SCK(4);
    else if (variant == 5) sk = BRS_SCK(5);
    else sk = BRS_SCK(6);
#undef BRS_SCK
    const ScanLayout SL_ = make_scan_layout(cfg->n_range, K, 4);
    e->sNT = 128;
    e->s_smem = SL_.total;
    e->s_kfn = sk;
    bool ok = (size_t)e->s_smem <= prop.sharedMemPerBlockOptin && nrl * 32 >= K * sl;
    int snb = 0;
    if (ok) {
      ce = cudaFuncSetAttribute(sk, cudaFuncAttributeMaxDynamicSharedMemorySize, e->s_smem);
      if (ce == cudaSuccess) ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&snb, sk, e->sNT, e->s_smem);
      ok = ce == cudaSuccess && snb >= 1;
      if (!ok) cudaGetLastError();
    }
    if (ok) {
      snb = std::min(snb, env_int("BRS_SOLO_MAX_CTAS", 32));
      e->s_warps_per_sm = snb * 4;
      e->s_ctas_per_sm = snb;
      e->s_grid = std::max(1, std::min((cfg->n_worlds + 3) / 4, snb * e->sm_count));
      // measured over the 30 cells of the config 5 sweep (profiles/r2): a warp per world wins once
      // a world has a few hundred ray-wall pairs or walls; tiny worlds (10 walls x 4 rays) keep the
      // fused kernel, whose tiles fill the lanes of the per-robot phases with many worlds
      const bool enough_work = (long long)K * cfg->seg_cap >= 256 || cfg->seg_cap >= 256;
      const bool auto_scan = enough_work && cfg->n_worlds >= env_int("BRS_SOLO_MIN_WORLDS", 4 * e->s_warps_per_sm * e->sm_count);
      e->scan = env_int("BRS_SOLO", auto_scan ? 1 : 0) != 0;
      e->solo_interleave = env_int("BRS_SOLO_INTERLEAVE", 1) != 0;
      if (e->scan) { e->solo = true; e->wide = false; e->prep_fused = false; }  // (solo = "a warp-per-world plan is active")
    }
  }
  // ---- the mini range path: launch geometry -------------------------------------
  // (same shapes as the solo range path, at most 16 rays; wins where a world is too little work
  // for a whole warp -- measured over the config 5 sweep, profiles/r2)
  if (scan_shape && e->scan_nrays <= 16 && e->scan_nrays >= 1) {
    const int K = e->scan_nrays;
    const int nray = pow2ceil(K);
    // lanes per world: enough to cover a world's walls a few at a time and to fill the GPU's
    // thread slots with the batch
    // (measured over the low-ray cells of the config 5 sweep, profiles/r2/sweep_config5_mini.txt:
    // 4 lanes up to ~100 walls, 16 beyond)
    int G = cfg->seg_cap <= 128 ? 4 : 16;
    G = std::max(2, std::min(16, pow2floor(env_int("BRS_MINI_G", G))));
    const void* mk = nullptr;
#define BRS_MK1(GG, NN) (const void*)brs_mini_kernel<GG, NN, 128, (NN >= 16 ? 4 : NN >= 8 ? 6 : 8)>
#define BRS_MK(GG) (nray == 1 ? BRS_MK1(GG, 1) : nray == 2 ? BRS_MK1(GG, 2) : nray == 4 ? BRS_MK1(GG, 4) : nray == 8 ? BRS_MK1(GG, 8) : BRS_MK1(GG, 16))
    mk = G == 2 ? BRS_MK(2) : G == 4 ? BRS_MK(4) : G == 8 ? BRS_MK(8) : BRS_MK(16);
#undef BRS_MK
#undef BRS_MK1
    const MiniLayout ML = make_mini_layout(cfg->n_range, K, 128 / G);
    int mnb = 0;
    ce = cudaFuncSetAttribute(mk, cudaFuncAttributeMaxDynamicSharedMemorySize, ML.total);
    if (ce == cudaSuccess) ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&mnb, mk, 128, ML.total);
    const bool ok = ce == cudaSuccess && mnb >= 1 && (size_t)ML.total <= prop.sharedMemPerBlockOptin;
    if (!ok) cudaGetLastError();
    // auto: worlds whose (walls x rays) work is below what keeps a warp busy, in batches that
    // fill the machine with teams (measured: see DESIGN.md, the config 5 sweep)
    const bool small_world = (long long)cfg->seg_cap * (5 + K) < (long long)env_int("BRS_MINI_MAX_WORK", 3000);
    // (any batch size: one world x 1000 steps -- config 1 as written -- is 2.8 us per step here
    // against 3.5 on the fused kernel; an explicit BRS_SOLO keeps its meaning: fused or warp per world)
    const bool auto_mini = ok && small_world && !getenv("BRS_SOLO");
    if (ok && env_int("BRS_MINI", auto_mini ? 1 : 0) != 0 && env_int("BRS_SOLO", 1) != 0) {
      e->mini = true; e->mini_G = G; e->mini_NRAY = nray;
      e->scan = true; e->solo = true; e->wide = false; e->prep_fused = false;
      e->sNT = 128;
      e->s_smem = ML.total;
      e->s_kfn = mk;
      mnb = std::min(mnb, env_int("BRS_SOLO_MAX_CTAS", 32));
      e->s_ctas_per_sm = mnb;
      e->s_warps_per_sm = mnb * 4;
      const int wpc = 128 / G;
      e->s_grid = std::max(1, std::min((cfg->n_worlds + wpc - 1) / wpc, mnb * e->sm_count));
    }
  }
  // ---- the solo multi path: launch geometry -------------------------------------
  if (multi_shape && !e->solo) {
    const int bw = std::max(1, 16 / pow2ceil(R));
    e->multi_bw = bw;
    const MultiLayout ML = make_multi_layout(R, bw, cfg->seg_cap + 4 * R, cfg->n_range, cfg->n_light, cfg->n_bulbs);
    const int variant = env_int("BRS_SOLO_CTAS", 6);
    const void* mk = variant >= 8 ? (const void*)brs_multi_kernel<128, 8>
                     : variant <= 4 ? (const void*)brs_multi_kernel<128, 4>
                     : variant == 5 ? (const void*)brs_multi_kernel<128, 5> : (const void*)brs_multi_kernel<128, 6>;
    e->sNT = 128;
    e->s_smem = ML.per_warp * 4;
    e->s_kfn = mk;
    bool ok = (size_t)e->s_smem <= prop.sharedMemPerBlockOptin;
    int snb = 0;
    if (ok) {
      ce = cudaFuncSetAttribute(mk, cudaFuncAttributeMaxDynamicSharedMemorySize, e->s_smem);
      if (ce == cudaSuccess) ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&snb, mk, e->sNT, e->s_smem);
      ok = ce == cudaSuccess && snb >= 1;
      if (!ok) cudaGetLastError();
    }
    if (ok) {
      snb = std::min(snb, env_int("BRS_SOLO_MAX_CTAS", 32));
      e->s_warps_per_sm = snb * 4;
      e->s_ctas_per_sm = snb;
      e->s_grid = std::max(1, std::min((cfg->n_worlds + 3) / 4, snb * e->sm_count));
      // measured (profiles/r2, config 3): 0.22 ms against the fused kernel's 0.20 -- a warp alone
      // leaves most lanes idle in the per-robot phases that a 15-world tile fills -- so this plan
      // is opt-in (BRS_MULTI=1), kept bit-exact by the tests
      e->multi = env_int("BRS_MULTI", 0) != 0;
      e->solo_interleave = env_int("BRS_SOLO_INTERLEAVE", 1) != 0;
      if (e->multi) { e->solo = true; e->wide = false; e->prep_fused = false; }
    }
  }
  // ---- brs_step_host pipeline: a copy stream and the events that chain it to the caller's
  e->host_chunks = std::max(1, std::min(8, env_int("BRS_HOST_CHUNKS", cfg->n_worlds >= 8192 ? 4 : 1)));
  e->steps_off = env_int("BRS_STEPS_ONE_LAUNCH", 1) == 0;
  e->use_rmask = env_int("BRS_RMASK", 1) != 0;
  e->pdl_chain = env_int("BRS_PDL_CHAIN", 1) != 0;
  ce = cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking);
  for (int k = 0; k < 8 && ce == cudaSuccess; ++k) ce = cudaEventCreateWithFlags(&e->ev_chunk[k], cudaEventDisableTiming);
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&e->ev_copied, cudaEventDisableTiming);
  if (ce != cudaSuccess) {
    brs_destroy(e);
    return fail(BRS_ERR_CUDA, std::string("brs_create: stream / event creation failed: ") + cudaGetErrorString(ce));
  }
  *out = e;
  return BRS_OK;
}

extern "C" int brs_destroy(brs_engine* e) {
  if (!e) return BRS_OK;
  cudaSetDevice(e->cfg.device);
#ifdef BRS_PHASE_TIMING
  if (e->d_phase) {
    unsigned long long h[16];
    cudaDeviceSynchronize();
    cudaMemcpy(h, e->d_phase, 128, cudaMemcpyDeviceToHost);
    const char* nm[16] = {"w:tma-wait", "w:table", "w:sense", "w:paint", "w:WAIT-FOR-HELPER", "", "", "",
                          "h:prep", "h:cull", "h:queue", "h:pairs", "h:commit", "h:direct", "h:-", "h:WAIT-FOR-WORKERS"};
    double tw = 0, th = 0;
    for (int k = 0; k < 8; ++k) { tw += (double)h[k]; th += (double)h[8 + k]; }
    fprintf(stderr, "[brs phase] workers:");
    for (int k = 0; k < 8; ++k) if (h[k]) fprintf(stderr, " %s=%.1f%%", nm[k] + 2, 100.0 * h[k] / tw);
    fprintf(stderr, "\n[brs phase] helper :");
    for (int k = 8; k < 16; ++k) if (h[k]) fprintf(stderr, " %s=%.1f%%", nm[k] + 2, 100.0 * h[k] / th);
    fprintf(stderr, "\n");
    cudaFree(e->d_phase);
  }
#endif
  for (int b = 0; b < BRS_BUF_COUNT_; ++b)
    if (e->buf[b]) cudaFree(e->buf[b]);
  cudaFree(e->d_robots);
  cudaFree(e->d_ranges);
  cudaFree(e->d_cameras);
  cudaFree(e->d_lights);
  cudaFree(e->d_smells);
  cudaFree(e->d_groups);
  cudaFree(e->d_gr_idx);
  cudaFree(e->d_dr_idx);
  cudaFree(e->d_sched);
  cudaFree(e->d_epoch);
  cudaFree(e->d_camcs);
  cudaFree(e->d_rowsky);
  cudaFree(e->d_rowgnd);
  cudaFree(e->d_ws);
  cudaFree(e->d_counters);
  cudaFree(e->d_scan_rays);
  cudaFree(e->d_multi_rays);
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  for (int k = 0; k < 8; ++k)
    if (e->ev_chunk[k]) cudaEventDestroy(e->ev_chunk[k]);
  if (e->ev_copied) cudaEventDestroy(e->ev_copied);
  delete e;
  return BRS_OK;
}

static void* cur(brs_engine* e, int which) { return e->bound[which] ? e->bound[which] : e->buf[which]; }

extern "C" int brs_device_ptr(brs_engine* e, int which, void** ptr, size_t* bytes) {
  if (!e || which < 0 || which >= BRS_BUF_COUNT_) return fail(BRS_ERR_INVALID, "brs_device_ptr: bad argument");
  if (ptr) *ptr = cur(e, which);
  if (bytes) *bytes = e->stride[which] * (size_t)e->cfg.n_worlds;
  return BRS_OK;
}

extern "C" int brs_bind_output(brs_engine* e, int which, void* device_ptr) {
  if (!e || which < BRS_BUF_RANGE || which > BRS_BUF_SMELL)
    return fail(BRS_ERR_INVALID, "brs_bind_output: only observation buffers can be rebound");
  if (device_ptr && ((uintptr_t)device_ptr & 15)) return fail(BRS_ERR_INVALID, "brs_bind_output: pointer must be 16-byte aligned");
  e->bound[which] = device_ptr;
  return BRS_OK;
}

// The fused gather: observation `which` is ALSO stored, by the kernel that produces it, at
// peer_ptrs[k] for k < n_peers -- peer-GPU memory mapped into this process (CUDA IPC / peer
// access), laid out like the engine's own output.  With every rank's slice of every rank's
// gather buffer bound this way, the step leaves the complete training batch on every GPU and
// only a stream barrier across the ranks remains of the collective.
extern "C" int brs_bind_peer_outputs(brs_engine* e, int which, int n_peers, void* const* peer_ptrs) {
  if (!e) return fail(BRS_ERR_INVALID, "brs_bind_peer_outputs: null engine");
  if (n_peers < 0 || n_peers > 7 || (n_peers && !peer_ptrs)) return fail(BRS_ERR_INVALID, "brs_bind_peer_outputs: 0..7 peers");
  if (n_peers == 0) {
    e->n_peers = 0;
    return BRS_OK;
  }
  if (which != BRS_BUF_CAMERA || !e->solo || e->scan || e->multi)
    return fail(BRS_ERR_UNSUPPORTED, "brs_bind_peer_outputs: only the pictures of the solo launch plan are stored to peers");
  for (int k = 0; k < n_peers; ++k) {
    if (!peer_ptrs[k] || ((uintptr_t)peer_ptrs[k] & 15)) return fail(BRS_ERR_INVALID, "brs_bind_peer_outputs: pointers must be 16-byte aligned");
    e->peer_cam[k] = peer_ptrs[k];
  }
  e->n_peers = n_peers;
  return BRS_OK;
}

// Gather buffers that other processes can map: plain cudaMalloc memory plus its CUDA IPC handle
// (64 bytes, to be sent to the peers by whatever transport the host code has), and the peer
// side: map a handle into this process with peer access from this engine's GPU enabled.
extern "C" int brs_peer_alloc(brs_engine* e, size_t bytes, void** dev_ptr, unsigned char* handle64) {
  if (!e || !dev_ptr || !handle64) return fail(BRS_ERR_INVALID, "brs_peer_alloc: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  CK(cudaSetDevice(e->cfg.device));
  void* ptr = nullptr;
  CK(cudaMalloc(&ptr, std::max<size_t>(bytes, 256)));
  cudaIpcMemHandle_t h;
  cudaError_t ce = cudaIpcGetMemHandle(&h, ptr);
  if (ce != cudaSuccess) {
    cudaFree(ptr);
    return fail(BRS_ERR_CUDA, std::string("brs_peer_alloc: cudaIpcGetMemHandle: ") + cudaGetErrorString(ce));
  }
  memcpy(handle64, &h, 64);
  *dev_ptr = ptr;
  return BRS_OK;
}
extern "C" int brs_peer_free(brs_engine* e, void* dev_ptr) {
  if (e) cudaSetDevice(e->cfg.device);
  if (dev_ptr) CK(cudaFree(dev_ptr));
  return BRS_OK;
}
extern "C" int brs_peer_open(brs_engine* e, const unsigned char* handle64, void** dev_ptr) {
  if (!e || !dev_ptr || !handle64) return fail(BRS_ERR_INVALID, "brs_peer_open: null argument");
  CK(cudaSetDevice(e->cfg.device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* ptr = nullptr;
  CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
  *dev_ptr = ptr;
  return BRS_OK;
}
extern "C" int brs_peer_close(brs_engine* e, void* dev_ptr) {
  if (e) cudaSetDevice(e->cfg.device);
  if (dev_ptr) CK(cudaIpcCloseMemHandle(dev_ptr));
  return BRS_OK;
}

static int copy_range(brs_engine* e, int which, int first, int n, size_t* off, size_t* bytes) {
  if (!e || which < 0 || which >= BRS_BUF_COUNT_) return fail(BRS_ERR_INVALID, "copy: bad buffer id");
  if (n < 0) n = e->cfg.n_worlds - first;
  if (first < 0 || n < 0 || first + n > e->cfg.n_worlds) return fail(BRS_ERR_INVALID, "copy: world range out of bounds");
  *off = e->stride[which] * (size_t)first;
  *bytes = e->stride[which] * (size_t)n;
  return BRS_OK;
}

extern "C" int brs_upload(brs_engine* e, int which, const void* host, int first_world, int n_worlds, void* stream) {
  size_t off, bytes;
  int rc = copy_range(e, which, first_world, n_worlds, &off, &bytes);
  if (rc) return rc;
  if (!host) return fail(BRS_ERR_INVALID, "brs_upload: host == NULL");
  CK(cudaSetDevice(e->cfg.device));
  if (bytes) CK(cudaMemcpyAsync((char*)cur(e, which) + off, host, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return BRS_OK;
}

extern "C" int brs_download(brs_engine* e, int which, void* host, int first_world, int n_worlds, void* stream) {
  size_t off, bytes;
  int rc = copy_range(e, which, first_world, n_worlds, &off, &bytes);
  if (rc) return rc;
  if (!host) return fail(BRS_ERR_INVALID, "brs_download: host == NULL");
  CK(cudaSetDevice(e->cfg.device));
  if (bytes) CK(cudaMemcpyAsync(host, (const char*)cur(e, which) + off, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return BRS_OK;
}

// Kernel parameters of one step: everything but flags, dt and the (rebindable) output
// pointers is fixed at brs_create.
static void fill_params_uncached(brs_engine* e, KParams& p, double time_step, unsigned flags, bool wide, bool msteps);
static void fill_params(brs_engine* e, KParams& p, double time_step, unsigned flags, bool wide, bool msteps = false) {
  if (e->kp_flags != flags || e->kp_wide != wide || e->kp_msteps != msteps) {
    fill_params_uncached(e, e->kp_cache, time_step, flags, wide, msteps);
    e->kp_flags = flags;
    e->kp_wide = wide;
    e->kp_msteps = msteps;
  }
  p = e->kp_cache;
  p.dt = time_step;
  p.range_out = (float*)cur(e, BRS_BUF_RANGE);
  p.cam_out = (uint8_t*)cur(e, BRS_BUF_CAMERA);
  p.light_out = (float*)cur(e, BRS_BUF_LIGHT);
  p.smell_out = (float*)cur(e, BRS_BUF_SMELL);
  p.counters = e->counting ? e->d_counters : nullptr;
  p.n_peers = (e->solo && !e->scan && !e->multi) ? e->n_peers : 0;
  for (int k = 0; k < p.n_peers; ++k) p.peer_delta[k] = (long long)((char*)e->peer_cam[k] - (char*)p.cam_out);
  p.noise = e->noise;
  p.noise.step_lo = (uint32_t)(e->noise_step & 0xffffffffull);
  p.noise.step_hi = (uint32_t)(e->noise_step >> 32);
}
static void fill_params_uncached(brs_engine* e, KParams& p, double time_step, unsigned flags, bool wide, bool msteps) {
  const brs_config& c = e->cfg;
  memset(&p, 0, sizeof(p));
  p.n_worlds = c.n_worlds; p.n_robots = c.n_robots; p.seg_cap = c.seg_cap; p.n_bulbs = c.n_bulbs;
  p.grid_w = c.grid_w; p.grid_h = c.grid_h;
  p.n_range = c.n_range; p.n_camera = c.n_camera; p.n_light = c.n_light; p.n_smell = c.n_smell;
  p.cam_w = e->cam_w; p.cam_h = e->cam_h;
  p.n_group = e->n_group; p.n_gr = e->n_gr; p.n_dr = e->n_dr;
  p.GS = e->GS;
  p.flat_rows = e->flat_rows;
  if (wide) {
    p.TW = e->wTW; p.n_tiles = e->w_tiles; p.SLG = e->wSLG; p.SLD = e->wSLD; p.SLL = e->wSLL; p.paint_rs = e->w_paint_rs;
    p.direct_on_helper = 0;
    p.fused_table = e->w_fused_table;
    p.NW = e->wNT; p.NH = 0;
    p.cq_max = wide_cq_max(e->wTW, c.n_robots);
  } else {
    p.TW = e->TW; p.n_tiles = e->n_tiles; p.SLG = e->SLG; p.SLD = e->SLD; p.SLL = e->SLL; p.paint_rs = e->paint_rs;
    p.direct_on_helper = e->direct_on_helper ? 1 : (e->split_lights ? 2 : 0);
    p.fused_table = e->fused_table;
    if (msteps) {
      p.TW = e->mTW; p.n_tiles = e->m_tiles; p.SLG = e->mSLG; p.SLD = e->mSLD; p.SLL = e->mSLL; p.paint_rs = e->m_paint_rs;
      p.fused_table = e->m_fused_table;
    }
    p.NW = e->NW; p.NH = e->NH;
    p.cq_max = BRS_CQ_MAX;
    p.prepped = e->prep_fused ? 1 : 0;
  }
  {
    const int R = c.n_robots, nvis = c.seg_cap + 4 * R;
    p.dR = make_fastdiv(R); p.dSC = make_fastdiv(c.seg_cap); p.dSQ = make_fastdiv(c.seg_cap / 4); p.dTab = make_fastdiv(std::max(e->n_group * nvis, 1));
    p.dNvis = make_fastdiv(nvis); p.dPair = make_fastdiv(R * R * 2); p.dNG = make_fastdiv(std::max(e->n_group, 1));
    const bool cam = (flags & BRS_STEP_CAMERA) && c.n_camera > 0, sense = flags & BRS_STEP_SENSE;
    const int cblocks = cam ? c.n_camera * (e->cam_w / e->NR) : 0;
    p.dGT = make_fastdiv(std::max(cblocks + (sense ? e->n_gr : 0), 1));
    p.dDT = make_fastdiv(std::max(e->n_dr, 1));
    p.dLT = make_fastdiv(std::max(c.n_light, 1));
    p.dST = make_fastdiv(std::max(c.n_smell, 1));
    p.dCpb = make_fastdiv(std::max(e->cam_w / std::max(e->NR, 1), 1));
  }
  p.cull_margin = 0.05f;
  p.flags = flags;
  p.dt = time_step; p.smell_cell = c.smell_cell_size; p.light_mult = c.light_multiplier;
  p.seg = (const uint32_t*)e->buf[BRS_BUF_SEG];
  p.seg_count = (const int*)e->buf[BRS_BUF_SEG_COUNT];
  p.world_size = (const float*)e->buf[BRS_BUF_WORLD_SIZE];
  p.bulbs = (const float*)e->buf[BRS_BUF_BULBS];
  p.grid = (const float*)e->buf[BRS_BUF_GRID];
  p.pose = (double*)e->buf[BRS_BUF_POSE];
  p.vel = (double*)e->buf[BRS_BUF_VEL];
  p.tvel = (const double*)e->buf[BRS_BUF_TARGET_VEL];
  p.stalled = (uint8_t*)e->buf[BRS_BUF_STALLED];
  p.range_out = (float*)cur(e, BRS_BUF_RANGE);
  p.cam_out = (uint8_t*)cur(e, BRS_BUF_CAMERA);
  p.light_out = (float*)cur(e, BRS_BUF_LIGHT);
  p.smell_out = (float*)cur(e, BRS_BUF_SMELL);
  p.robots = e->d_robots; p.ranges = e->d_ranges; p.cameras = e->d_cameras;
  p.lights = e->d_lights; p.smells = e->d_smells;
  p.groups = e->d_groups; p.gr_idx = e->d_gr_idx; p.dr_idx = e->d_dr_idx;
  p.cam_cs = e->d_camcs; p.row_sky = e->d_rowsky; p.row_gnd = e->d_rowgnd;
  p.sched = e->d_sched;
  p.epoch = e->d_epoch;
  p.phase = e->d_phase;
  p.wsg = e->d_ws;
  p.counters = e->counting ? e->d_counters : nullptr;
  p.s_robot = e->h_robot0; p.s_camera = e->h_camera0; p.s_group = e->h_group0;
  p.s_sky = e->h_sky0; p.s_gnd = e->h_gnd0;
  p.solo_interleave = e->solo_interleave;
  for (int r = 0; r < 8; ++r) p.m_reach[r] = e->multi_reach[r];
  if (e->multi) {
    p.m_rays = e->d_multi_rays; p.m_nrays = e->multi_nrays; p.m_bw = e->multi_bw;
  }
  {
    // reach masks for the direct-path range sensors: one lane per ray, every segment of a world
    // in one ballot, and enough sensors per robot that a warp spans only a few robots
    const int sld = wide ? e->wSLD : (msteps ? e->mSLD : e->SLD);
    const int team = wide ? e->wNT : (e->direct_on_helper ? 32 * e->NH : e->NW);
    p.use_rmask = (e->use_rmask && !e->solo && !e->use_rlist && sld == 1 && team % 32 == 0 &&
                   c.seg_cap + 4 * c.n_robots <= 32 && e->n_dr >= 8 * c.n_robots) ? 1 : 0;
  }
  if (e->use_rlist && !e->solo) {
    const int tw = wide ? e->wTW : (msteps ? e->mTW : e->TW), R = c.n_robots, nvis = c.seg_cap + 4 * R;
    const int base = wide ? make_wide_layout(tw, c.seg_cap, R, e->n_group, c.n_camera, e->cam_w, e->cam_h, !e->w_big, e->wNT > 32).total
                          : make_layout(tw, c.seg_cap, R, e->n_group, c.n_camera, e->cam_w, e->cam_h, !e->big).total;
    p.use_rlist = 1;
    p.rl_off = base;
    p.rc_off = base + align_up(tw * R * nvis, 16);
  }
  if (e->scan) {
    p.s_group = e->h_scan_group;
    p.s_rays = e->d_scan_rays; p.s_nrays = e->scan_nrays; p.s_SL = e->scan_SL;
  }
}

// One World.step over worlds [first, first + n): every buffer has `world` as its leading
// dimension, so a sub-range is the same kernel on shifted base pointers.
static int launch_range(brs_engine* e, int first, int n, double time_step, unsigned flags, void* stream, int n_steps = 1) {
  KParams p;
  const bool msteps = n_steps > 1 && !e->solo && !e->wide;  // the fused kernel's brs_steps shape
  fill_params(e, p, time_step, flags, e->wide, msteps);
  p.n_steps = n_steps;
  p.epoch0 = e->epoch0;
  const brs_config& c = e->cfg;
  int grid = e->solo ? e->s_grid : e->wide ? e->w_grid : (msteps ? e->m_grid : e->grid);
  if (first != 0 || n != c.n_worlds) {
    const size_t R = c.n_robots, f = (size_t)first;
    p.seg += f * 5 * c.seg_cap; p.seg_count += f; p.world_size += 2 * f;
    p.bulbs += f * c.n_bulbs * 4; p.grid += f * c.grid_w * c.grid_h;
    p.pose += f * R * 3; p.vel += f * R * 3; p.tvel += f * R * 3; p.stalled += f * R;
    p.range_out += f * c.n_range; p.cam_out += f * e->stride[BRS_BUF_CAMERA];
    p.light_out += f * c.n_light; p.smell_out += f * c.n_smell;
    if (p.wsg) p.wsg += f * R * BRS_WS_STRIDE;
    p.n_worlds = n;
    p.noise.world0 += (uint32_t)first;
    p.n_tiles = (n + p.TW - 1) / p.TW;
    if (e->solo) {
      const int wpc = e->mini ? e->sNT / e->mini_G : e->sNT / 32;
      grid = std::max(1, std::min((n + wpc - 1) / wpc, e->s_ctas_per_sm * e->sm_count));
    } else {
      grid = std::max(1, std::min(p.n_tiles, (e->wide ? e->w_ctas_per_sm : e->f_ctas_per_sm) * e->sm_count));
    }
  }
  void* args[] = {(void*)&p};
  if (e->solo) {
    CK(cudaLaunchKernel(e->s_kfn, dim3(grid), dim3(e->sNT), args, (size_t)e->s_smem, (cudaStream_t)stream));
  } else if (!e->wide && !e->prep_fused) {
    if (e->pdl_chain) {
      p.pdl = 1;
      cudaLaunchConfig_t lc;
      memset(&lc, 0, sizeof(lc));
      lc.gridDim = dim3(grid);
      lc.blockDim = dim3(e->T);
      lc.dynamicSmemBytes = (size_t)(msteps ? e->m_smem : e->smem);
      lc.stream = (cudaStream_t)stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      lc.attrs = at;
      lc.numAttrs = 1;
      CK(cudaLaunchKernelExC(&lc, e->kfn, args));
    } else {
      CK(cudaLaunchKernel(e->kfn, dim3(grid), dim3(e->T), args, (size_t)(msteps ? e->m_smem : e->smem), (cudaStream_t)stream));
    }
  } else {
    // prep (thread per robot half), then the step kernel under programmatic dependent launch:
    // its prologue and first wall copies overlap the prep kernel, griddepcontrol.wait orders the rest
    const unsigned halves = (unsigned)n * c.n_robots * ((flags & BRS_STEP_MOVE) ? 2u : 1u);
    CK(cudaLaunchKernel((const void*)brs_prep_kernel, dim3((halves + 255u) / 256u), dim3(256), args, 0, (cudaStream_t)stream));
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof(lc));
    lc.gridDim = dim3(grid);
    lc.blockDim = dim3(e->wide ? e->wNT : e->T);
    lc.dynamicSmemBytes = (size_t)(e->wide ? e->w_smem : e->smem);
    lc.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = e->pdl ? 1 : 0;
    lc.attrs = at;
    lc.numAttrs = 1;
    CK(cudaLaunchKernelExC(&lc, e->wide ? e->w_kfn : e->kfn, args));
  }
  CK(cudaGetLastError());
  return BRS_OK;
}

extern "C" int brs_step(brs_engine* e, double time_step, unsigned flags, void* stream) {
  NvtxRange nvtx_("brs_step");
  if (!e) return fail(BRS_ERR_INVALID, "brs_step: null engine");
  if (!(flags & BRS_STEP_ALL)) return BRS_OK;
  CK(cudaSetDevice(e->cfg.device));
  const int rc = launch_range(e, 0, e->cfg.n_worlds, time_step, flags, stream);
  if (flags & BRS_STEP_SENSE) ++e->noise_step;
  return rc;
}

// World.steps(n): n consecutive steps with the inputs as they are now.  One launch when the
// active plan can carry its pipeline across step boundaries (fused, solo, solo range; every
// tile / world then belongs to one CTA / warp for the whole launch), else n launches.  Every
// step writes all of its observations; what is left in the buffers is the last step's.
static bool steps_one_launch(const brs_engine* e, unsigned flags, int n_steps) {
  return n_steps > 1 && (flags & BRS_STEP_MOVE) && !e->noise.on && !e->counting && !e->wide && !e->prep_fused &&
         !e->multi && !e->steps_off;
}
// (the fused kernel counts (step, tile) items in 32 bits: very long runs go in pieces)
static int steps_piece(const brs_engine* e, int n_steps) {
  const int tiles = e->solo ? 1 : std::max(e->m_tiles, 1);
  return std::max(2, (int)std::min<long long>(n_steps, 0x3fffffffll / tiles));
}
extern "C" int brs_steps_launches(brs_engine* e, unsigned flags, int n_steps, int* launches) {
  if (!e || !launches) return fail(BRS_ERR_INVALID, "brs_steps_launches: null argument");
  const int per = (e->wide || e->prep_fused) ? 2 : 1;
  if (!(flags & BRS_STEP_ALL) || n_steps <= 0) *launches = 0;
  else if (steps_one_launch(e, flags, n_steps)) *launches = (n_steps + steps_piece(e, n_steps) - 1) / steps_piece(e, n_steps);
  else *launches = ((flags & BRS_STEP_MOVE) ? n_steps : 1) * per;
  return BRS_OK;
}
// the launch shape brs_steps uses when it runs as one launch (the fused kernel picks larger tiles for it)
extern "C" int brs_steps_info(brs_engine* e, int* worlds_per_tile, int* n_tiles, int* ctas, int* smem_bytes) {
  if (!e) return fail(BRS_ERR_INVALID, "brs_steps_info: null engine");
  const bool fused = !e->solo && !e->wide;
  if (worlds_per_tile) *worlds_per_tile = e->solo ? 1 : fused ? e->mTW : e->wTW;
  if (n_tiles) *n_tiles = e->solo ? e->cfg.n_worlds : fused ? e->m_tiles : e->w_tiles;
  if (ctas) *ctas = e->solo ? e->s_grid : fused ? e->m_grid : e->w_grid;
  if (smem_bytes) *smem_bytes = e->solo ? e->s_smem : fused ? e->m_smem : e->w_smem;
  return BRS_OK;
}
extern "C" int brs_steps(brs_engine* e, double time_step, unsigned flags, int n_steps, void* stream) {
  NvtxRange nvtx_("brs_steps");
  if (!e) return fail(BRS_ERR_INVALID, "brs_steps: null engine");
  if (n_steps < 0) return fail(BRS_ERR_INVALID, "brs_steps: n_steps < 0");
  if (!(flags & BRS_STEP_ALL) || n_steps == 0) return BRS_OK;
  CK(cudaSetDevice(e->cfg.device));
  const bool one_launch = steps_one_launch(e, flags, n_steps);
  if (one_launch) {
    const int piece = steps_piece(e, n_steps);
    for (int done = 0; done < n_steps;) {
      const int n = std::min(piece, n_steps - done);
      const int rc = launch_range(e, 0, e->cfg.n_worlds, time_step, flags, stream, n);
      if (rc) return rc;
      if (!e->solo && n > 1) e->epoch0 += (unsigned)n;
      done += n;
    }
    return BRS_OK;
  }
  for (int s = 0; s < n_steps; ++s) {
    const int rc = launch_range(e, 0, e->cfg.n_worlds, time_step, flags, stream);
    if (flags & BRS_STEP_SENSE) ++e->noise_step;
    if (rc) return rc;
    if (!(flags & BRS_STEP_MOVE)) break;  // sensing a world that does not move twice is sensing it once
  }
  return BRS_OK;
}

// Seeded sensor noise.  Off by default (the parity contract of BASELINE.json is stated with
// noise disabled); a (seed, world, sensor, step) -> N(0,1) map that no launch detail can change.
extern "C" int brs_set_noise(brs_engine* e, unsigned long long seed, unsigned first_world, double range_std,
                             double light_rel_std) {
  if (!e) return fail(BRS_ERR_INVALID, "brs_set_noise: null engine");
  if (!(range_std >= 0) || !(light_rel_std >= 0)) return fail(BRS_ERR_INVALID, "brs_set_noise: negative standard deviation");
  memset(&e->noise, 0, sizeof(e->noise));
  e->noise.on = (range_std > 0 || light_rel_std > 0) ? 1u : 0u;
  e->noise.k0 = (uint32_t)(seed & 0xffffffffull);
  e->noise.k1 = (uint32_t)(seed >> 32);
  e->noise.world0 = first_world;
  e->noise.range_std = (float)range_std;
  e->noise.light_rel_std = (float)light_rel_std;
  e->noise_step = 0;
  return BRS_OK;
}

extern "C" int brs_step_counted(brs_engine* e, double time_step, unsigned flags, void* stream,
                                unsigned long long* ray_tests, unsigned long long* edge_tests) {
  NvtxRange nvtx_("brs_step_counted");
  if (!e) return fail(BRS_ERR_INVALID, "brs_step_counted: null engine");
  CK(cudaSetDevice(e->cfg.device));
  if (!e->d_counters) CK(cudaMalloc((void**)&e->d_counters, 16));
  CK(cudaMemsetAsync(e->d_counters, 0, 16, (cudaStream_t)stream));
  e->counting = true;
  const int rc = brs_step(e, time_step, flags, stream);
  e->counting = false;
  if (rc) return rc;
  unsigned long long h[2] = {0, 0};
  CK(cudaMemcpyAsync(h, e->d_counters, 16, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CK(cudaStreamSynchronize((cudaStream_t)stream));
  if (ray_tests) *ray_tests = h[0];
  if (edge_tests) *edge_tests = h[1];
  return BRS_OK;
}

// The end-to-end step.  The batch is cut into `host_chunks` world ranges: the kernel of range
// c+1 runs on the caller's stream while the observations of range c drain to the host on the
// engine's copy stream (one cudaMemcpyAsync per buffer and range); the caller's stream then
// waits for the last copy, so everything enqueued after this call sees the host buffers filled.
extern "C" int brs_step_host(brs_engine* e, double time_step, unsigned flags, const double* h_target_vel,
                             double* h_pose, uint8_t* h_stalled, float* h_range, uint8_t* h_camera, float* h_light,
                             float* h_smell, void* stream) {
  NvtxRange nvtx_("brs_step_host");
  if (!e) return fail(BRS_ERR_INVALID, "brs_step_host: null engine");
  CK(cudaSetDevice(e->cfg.device));
  const int B = e->cfg.n_worlds, C = std::max(1, std::min(e->host_chunks, B));
  cudaStream_t st = (cudaStream_t)stream;
  struct Out { int which; void* host; };
  const Out outs[] = {{BRS_BUF_POSE, h_pose}, {BRS_BUF_STALLED, h_stalled}, {BRS_BUF_RANGE, h_range},
                      {BRS_BUF_CAMERA, h_camera}, {BRS_BUF_LIGHT, h_light}, {BRS_BUF_SMELL, h_smell}};
  int rc;
  for (int c = 0; c < C; ++c) {
    const int first = (int)((long long)B * c / C), n = (int)((long long)B * (c + 1) / C) - first;
    if (n <= 0) continue;
    if (h_target_vel) {
      const size_t sb = e->stride[BRS_BUF_TARGET_VEL];
      CK(cudaMemcpyAsync((char*)e->buf[BRS_BUF_TARGET_VEL] + sb * first, (const char*)h_target_vel + sb * first, sb * n,
                         cudaMemcpyHostToDevice, st));
    }
    if ((flags & BRS_STEP_ALL) && (rc = launch_range(e, first, n, time_step, flags, stream))) return rc;
    cudaStream_t cs = C > 1 ? e->copy_stream : st;
    if (C > 1) {
      CK(cudaEventRecord(e->ev_chunk[c], st));
      CK(cudaStreamWaitEvent(cs, e->ev_chunk[c], 0));
    }
    for (const Out& o : outs) {
      const size_t sb = e->stride[o.which];
      if (o.host && sb)
        CK(cudaMemcpyAsync((char*)o.host + sb * first, (const char*)cur(e, o.which) + sb * first, sb * n,
                           cudaMemcpyDeviceToHost, cs));
    }
  }
  if (C > 1) {
    CK(cudaEventRecord(e->ev_copied, e->copy_stream));
    CK(cudaStreamWaitEvent(st, e->ev_copied, 0));
  }
  if (flags & BRS_STEP_SENSE) ++e->noise_step;
  return BRS_OK;
}

// Pinned host memory for brs_step_host, first-touched on the NUMA node of the engine's GPU:
// the calling thread is moved to the CPUs next to the GPU (sysfs local_cpulist of its PCI
// function) for the allocation and the touch, then back.  With one process per GPU this keeps
// each rank's observation traffic on its own socket's memory.
extern "C" int brs_host_alloc(brs_engine* e, size_t bytes, void** host) {
  if (!e || !host) return fail(BRS_ERR_INVALID, "brs_host_alloc: null argument");
  CK(cudaSetDevice(e->cfg.device));
  cpu_set_t old_set, new_set;
  bool moved = false;
  CPU_ZERO(&new_set);
  if (sched_getaffinity(0, sizeof(old_set), &old_set) == 0) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), e->cfg.device) == cudaSuccess) {
      for (char* q = bus; *q; ++q) *q = (char)tolower(*q);
      const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist";
      if (FILE* f = fopen(path.c_str(), "r")) {
        char line[4096] = {0};
        if (fgets(line, sizeof(line), f)) {
          int n_set = 0;
          for (char* tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
            int a = 0, b = 0;
            const int k = sscanf(tok, "%d-%d", &a, &b);
            if (k == 1) b = a;
            if (k >= 1)
              for (int cpu = a; cpu <= b && cpu < CPU_SETSIZE; ++cpu)
                if (CPU_ISSET(cpu, &old_set)) { CPU_SET(cpu, &new_set); ++n_set; }
          }
          moved = n_set > 0 && sched_setaffinity(0, sizeof(new_set), &new_set) == 0;
        }
        fclose(f);
      }
    }
  }
  void* ptr = nullptr;
  cudaError_t ce = cudaHostAlloc(&ptr, std::max<size_t>(bytes, 16), cudaHostAllocDefault);
  if (ce == cudaSuccess) memset(ptr, 0, std::max<size_t>(bytes, 16));  // first touch, on the GPU's node
  if (moved) sched_setaffinity(0, sizeof(old_set), &old_set);
  if (ce != cudaSuccess) return fail(BRS_ERR_CUDA, std::string("brs_host_alloc: ") + cudaGetErrorString(ce));
  *host = ptr;
  return BRS_OK;
}

extern "C" int brs_host_free(brs_engine* e, void* host) {
  if (e) cudaSetDevice(e->cfg.device);
  if (host) CK(cudaFreeHost(host));
  return BRS_OK;
}

// Top-down pictures of worlds [first, first + n): see brs_render_kernel (brs_kernels.cuh).
extern "C" int brs_render_world(brs_engine* e, int first_world, int n_worlds, int img_w, int img_h, int robot,
                                double view_w, double view_h, double line_px, double bulb_px, void* d_out, void* stream) {
  NvtxRange nvtx_("brs_render_world");
  if (!e || !d_out) return fail(BRS_ERR_INVALID, "brs_render_world: null argument");
  const brs_config& c = e->cfg;
  if (first_world < 0 || n_worlds < 1 || first_world + n_worlds > c.n_worlds || img_w < 1 || img_h < 1 || img_w > 8192 || img_h > 8192)
    return fail(BRS_ERR_INVALID, "brs_render_world: bad world range or image size");
  if (robot >= c.n_robots || (robot >= 0 && !(view_w > 0 && view_h > 0)))
    return fail(BRS_ERR_INVALID, "brs_render_world: bad robot index or view size");
  CK(cudaSetDevice(c.device));
  RenderParams q;
  memset(&q, 0, sizeof(q));
  q.first_world = first_world; q.n_worlds = n_worlds; q.img_w = img_w; q.img_h = img_h;
  q.bands = std::max(1, std::min(img_h, (4 * e->sm_count + n_worlds - 1) / n_worlds));
  q.seg_cap = c.seg_cap; q.n_robots = c.n_robots; q.n_bulbs = c.n_bulbs;
  q.robot = robot < 0 ? -1 : robot;
  q.view_w = (float)view_w; q.view_h = (float)view_h;
  q.half_px = (float)(0.5 * (line_px > 0 ? line_px : 1.5));
  q.bulb_px = (float)(bulb_px > 0 ? bulb_px : 2.5);
  q.ground_rgba = pack_rgb(c.ground_rgb[0], c.ground_rgb[1], c.ground_rgb[2]) | 0xff000000u;
  q.bulb_rgba = pack_rgb(255, 255, 0) | 0xff000000u;
  q.seg = (const uint32_t*)e->buf[BRS_BUF_SEG];
  q.seg_count = (const int*)e->buf[BRS_BUF_SEG_COUNT];
  q.world_size = (const float*)e->buf[BRS_BUF_WORLD_SIZE];
  q.bulbs = (const float*)e->buf[BRS_BUF_BULBS];
  q.pose = (const double*)e->buf[BRS_BUF_POSE];
  q.robots = e->d_robots;
  q.out = (uint32_t*)d_out;
  const size_t smem = (size_t)c.seg_cap * 20 + (size_t)c.n_robots * 36 + 16;
  if (smem > 200 * 1024) return fail(BRS_ERR_INVALID, "brs_render_world: too many walls per world for one CTA's shared memory");
  if (smem > 48 * 1024) CK(cudaFuncSetAttribute((const void*)brs_render_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  brs_render_kernel<<<n_worlds * q.bands, 256, smem, (cudaStream_t)stream>>>(q);
  CK(cudaGetLastError());
  return BRS_OK;
}

extern "C" int brs_synchronize(brs_engine* e, void* stream) {
  if (!e) return fail(BRS_ERR_INVALID, "brs_synchronize: null engine");
  CK(cudaSetDevice(e->cfg.device));
  CK(cudaStreamSynchronize((cudaStream_t)stream));
  return BRS_OK;
}

extern "C" int brs_count_tests(brs_engine* e, unsigned flags, void* stream, unsigned long long* tests) {
  if (!e || !tests) return fail(BRS_ERR_INVALID, "brs_count_tests: null argument");
  const brs_config& c = e->cfg;
  std::vector<int> cnt(c.n_worlds);
  CK(cudaSetDevice(c.device));
  CK(cudaMemcpyAsync(cnt.data(), e->buf[BRS_BUF_SEG_COUNT], (size_t)c.n_worlds * 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CK(cudaStreamSynchronize((cudaStream_t)stream));
  const unsigned long long R = c.n_robots;
  unsigned long long per_seg = 0;  // tests per visible segment
  if (flags & BRS_STEP_MOVE) per_seg += R * 4;
  if (flags & BRS_STEP_SENSE) per_seg += (unsigned long long)e->rays_range + (unsigned long long)c.n_light * c.n_bulbs;
  if (flags & BRS_STEP_CAMERA) per_seg += (unsigned long long)c.n_camera * e->cam_w;
  unsigned long long total = 0;
  for (int w = 0; w < c.n_worlds; ++w) {
    const unsigned long long S = (unsigned long long)std::min(std::max(cnt[w], 0), c.seg_cap);
    total += per_seg * (S + 4 * (R - 1));
  }
  *tests = total;
  return BRS_OK;
}

extern "C" int brs_launch_info(brs_engine* e, int* threads, int* ctas, int* smem_bytes, int* sm_count) {
  if (!e) return fail(BRS_ERR_INVALID, "brs_launch_info: null engine");
  if (threads) *threads = e->solo ? e->sNT : e->wide ? e->wNT : e->T;
  if (ctas) *ctas = e->solo ? e->s_grid : e->wide ? e->w_grid : e->grid;
  if (smem_bytes) *smem_bytes = e->solo ? e->s_smem : e->wide ? e->w_smem : e->smem;
  if (sm_count) *sm_count = e->sm_count;
  return BRS_OK;
}

extern "C" int brs_path_info(brs_engine* e, int* wide, int* launches_per_step) {
  if (!e) return fail(BRS_ERR_INVALID, "brs_path_info: null engine");
  if (wide) *wide = e->mini ? 6 : e->multi ? 5 : e->scan ? 4 : e->solo ? 3 : e->wide ? 1 : (e->prep_fused ? 2 : 0);
  if (launches_per_step) *launches_per_step = (!e->solo && (e->wide || e->prep_fused)) ? 2 : 1;
  return BRS_OK;
}

extern "C" int brs_tile_info(brs_engine* e, int* worlds_per_tile, int* n_tiles, int* workers, int* cols_per_thread,
                             int* n_groups) {
  if (!e) return fail(BRS_ERR_INVALID, "brs_tile_info: null engine");
  if (worlds_per_tile) *worlds_per_tile = e->solo ? 1 : e->wide ? e->wTW : e->TW;
  if (n_tiles) *n_tiles = e->solo ? e->cfg.n_worlds : e->wide ? e->w_tiles : e->n_tiles;
  if (workers) *workers = e->solo ? e->sNT : e->wide ? e->wNT : e->NW;
  if (cols_per_thread) *cols_per_thread = e->scan ? e->scan_NRL : e->solo ? e->cam_w / 32 : e->NR;
  if (n_groups) *n_groups = e->n_group;
  return BRS_OK;
}

extern "C" int brs_measure_fp32_peak(int device, int iters, double* tflops, double* sm_clock_mhz) {
  if (!tflops) return fail(BRS_ERR_INVALID, "brs_measure_fp32_peak: null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
    return fail(BRS_ERR_NO_DEVICE, "brs_measure_fp32_peak: no such CUDA device");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  float* d = nullptr;
  CK(cudaMalloc(&d, 16));
  const int threads = 1024, blocks = prop.multiProcessorCount * 2;
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  brs_ffma_kernel<<<blocks, threads>>>(d, iters / 4 + 1, 1.0000001f, 1e-7f);  // warm-up
  CK(cudaEventRecord(a));
  brs_ffma_kernel<<<blocks, threads>>>(d, iters, 1.0000001f, 1e-7f);
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, a, b));
  const double flops = 2.0 * 8 * 16 * (double)iters * threads * blocks;
  *tflops = flops / (ms * 1e-3) / 1e12;
  if (sm_clock_mhz) {
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
    *sm_clock_mhz = khz / 1000.0;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(d);
  return BRS_OK;
}

extern "C" int brs_measure_write_bw(int device, size_t bytes, int mode, int reps, double* gbs) {
  if (!gbs || bytes < 4096 || reps < 1) return fail(BRS_ERR_INVALID, "brs_measure_write_bw: bad argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
    return fail(BRS_ERR_NO_DEVICE, "brs_measure_write_bw: no such CUDA device");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  uint4* d = nullptr;
  CK(cudaMalloc((void**)&d, bytes));
  const size_t n16 = bytes / 16;
  const int blocks = prop.multiProcessorCount * 8;
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  // mode 0 / 1: grid-stride 16-byte stores (st.cs / default); 2: the driver's cudaMemsetAsync;
  // 3: CTA-contiguous 128 KB chunks of st.cs (the shape of PAINT)
  auto fill = [&]() {
    if (mode == 2) cudaMemsetAsync(d, 0x5a, n16 * 16, 0);
    else brs_fill_kernel<<<mode == 3 ? prop.multiProcessorCount * 4 : blocks, 256>>>(d, n16, mode);
  };
  fill();  // warm-up
  double best = 0;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(a));
    fill();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    best = std::max(best, (double)(n16 * 16) / (ms * 1e-3) / 1e9);
  }
  *gbs = best;
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(d);
  return BRS_OK;
}
