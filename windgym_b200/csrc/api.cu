// C-ABI of libwindgym_b200 (declared in include/windgym_b200.h): handle, state layout, launch orchestration.
// Host-side only; the kernels live in flow.cu and env.cu.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda.h>

#include "wg_internal.cuh"
#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define WG_CPU_RELAX() _mm_pause()
#else
#define WG_CPU_RELAX() ((void)0)
#endif

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char* what) {
  return fail(WG_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

struct Field {
  std::string name;
  size_t offset;
  int32_t dtype;  // 0 f32, 1 i32
  std::vector<int64_t> shape;
};

}  // namespace

struct wg_handle {
  wg_config cfg;
  wg::Dev dev;  // template: table pointers + dims set, state pointers filled per call from the state base
  std::vector<Field> fields;
  size_t state_bytes = 0;
  int hist_max = 0;
  uint64_t launches = 0;
  // optional per-kernel timing of wg_step (wg_profile_enable): event triples (before flow, after flow, after finish)
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events;
  size_t prof_used = 0;
  // device-resident immutable tables
  float *d_tab_ws = nullptr, *d_tab_p = nullptr, *d_tab_ct = nullptr;
  double *d_x = nullptr, *d_y = nullptr;
  int *d_ring_off = nullptr, *d_ring_chan = nullptr;
  wg::ObsDesc* d_desc = nullptr;
  wg::CopyField* d_copy = nullptr;
  int n_copy = 0;
  int n_active = 0;  // envs stepped by wg_step (prefix of the allocation; the rest is the spare pool)
  // longest-first launch order of wg_step (state field "order"): rebuilt when stale
  const void* order_state = nullptr;  // state tensor the current order was built for (null: rebuild)
  int order_n = 0;                    // ... and its active count
  int order_age = 0;                  // wg_step calls since the last rebuild
  bool use_order = true;              // WG_NO_ORDER=1 in the environment: plain index order (A/B measurements)
  // work table of single-step launches (wg_plan_kernel): resident CTA slots of the flow kernel variant in use
  // (queried once per attached turbulence set-up), table capacity, tail-split policy
  int slots = 0;
  int work_cap = 0;
  int n_work = 0;                     // entries of the current table
  int tail_units = 0, tail_parts = 1; // WG_TAIL_UNITS / WG_TAIL_PARTS in the environment (A/B measurements)
  bool use_split = true;              // WG_NO_SPLIT=1: never cut a farm into parts
  int max_part_tiles = 0;             // WG_MAX_PART_TILES (0: off): longest part of a large farm, in tiles.  Measured on
                                      // cfg 4 (8x8, 1024 envs): 48 -> 0.609 ms, 24 -> 0.620 ms, off -> 0.587 ms per launch
  float slot_share = 1.f;             // wg_set_slot_share / WG_SLOT_SHARE: fraction of the resident CTA slots the step plans for
  int fill_waves = 0;                 // WG_FILL_WAVES=k: below k waves of farms, equalised parts that fill the last wave
  bool two_wave = true;               // WG_NO_TWOWAVE=1: between 1 and 2 waves of farms, keep one CTA per farm
  bool pdl_next = true;               // WG_NO_PDL_NEXT=1: the next step's flow kernel waits for the whole finish kernel
  bool pdl_late = true;               // WG_NO_PDL_LATE=1: multi-wave grids launch the finish kernel in plain stream order
  bool after_swap = false;            // wg_pool_swap came last: its copy kernel releases its dependents at its start, so
                                      // the next flow launch waits at its top (FlowArgs::pdl_wait = 3) or is an ordinary one
  bool use_pdl = true;                // WG_NO_PDL=1: plain stream order between the flow and the finish kernel
  // wg_step_host, zero-copy path: completion word in mapped host memory + device arrival counter, step sequence
  // number, and the pinned host ranges already identified (host base, device alias, bytes)
  unsigned* flag_host = nullptr;
  unsigned* flag_dev = nullptr;       // device alias of flag_host
  unsigned* d_done_count = nullptr;
  unsigned seq = 0;
  bool use_zero_copy = true;          // WG_NO_ZEROCOPY=1: always stage through the copy engines
  struct PinnedRange { const char* host; char* dev; size_t bytes; };
  std::vector<PinnedRange> pinned;
  // L2 residency of the turbulence box: streams that already carry the access-policy window
  std::vector<cudaStream_t> policy_streams;
  size_t tb_lp_bytes = 0;
  std::vector<size_t> pool_off;       // ... and of the pool view (bind_pool)
  int* need_host = nullptr;           // mapped host word: spares waiting for a refill (wg_pool_need)
  int* need_dev = nullptr;            // its device alias
  std::vector<size_t> state_off;      // offsets of the device view's state pointers (WG_STATE_MEMBERS order)
  float4* d_lp8 = nullptr;            // library-owned brick copies of the low-pass box and of the raw box
  float4* d_raw8 = nullptr;
  bool use_bricks = true;             // WG_NO_BRICKS=1: always gather from the caller's lp layout
  bool force_bricks = false;          // WG_FORCE_BRICKS=1: bricks whatever the box size (tests)
};

namespace {

size_t add_field(wg_handle* h, const char* name, int32_t dtype, std::vector<int64_t> shape) {
  size_t n = 4;
  for (auto s : shape) n *= (size_t)s;
  size_t off = (h->state_bytes + 255) / 256 * 256;
  h->fields.push_back({name, off, dtype, shape});
  h->state_bytes = off + n;
  return off;
}

const Field* find_field(const wg_handle* h, const char* name) {
  for (auto& f : h->fields)
    if (f.name == name) return &f;
  return nullptr;
}

template <class Tp>
Tp* at(void* base, const wg_handle* h, const char* name) {
  return reinterpret_cast<Tp*>(reinterpret_cast<unsigned char*>(base) + find_field(h, name)->offset);
}

// The wake state streams through L2 once per step (GBs) and evicts the turbulence box, whose gathers are random
// 32-byte sector reads: pin the low-pass box (sampled once per station and step) as persisting L2 lines on the
// launching stream (measured on cfg 2 + 1024x128x32 box: 0.78 -> 0.64 ms per launch; per-load evict_last hints were
// slower than plain loads).  Best effort: failures leave the default policy.
void pin_turbulence_in_l2(wg_handle* h, cudaStream_t s) {
  if (!h->dev.tb_lp || h->dev.tb_lp8) return;  // bricks: the box is too large to stay in L2 anyway
  for (cudaStream_t t : h->policy_streams)
    if (t == s) return;
  h->policy_streams.push_back(s);
  int dev = 0, max_win = 0, max_persist = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, dev);
  cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
  if (max_win <= 0 || max_persist <= 0) { cudaGetLastError(); return; }
  const size_t want = h->tb_lp_bytes < (size_t)max_persist ? h->tb_lp_bytes : (size_t)max_persist;
  cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
  cudaStreamAttrValue v{};
  v.accessPolicyWindow.base_ptr = const_cast<float2*>(h->dev.tb_lp);
  v.accessPolicyWindow.num_bytes = h->tb_lp_bytes < (size_t)max_win ? h->tb_lp_bytes : (size_t)max_win;
  v.accessPolicyWindow.hitRatio = (float)((double)want / (double)v.accessPolicyWindow.num_bytes);
  if (v.accessPolicyWindow.hitRatio > 1.f) v.accessPolicyWindow.hitRatio = 1.f;
  v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &v);
  cudaGetLastError();
}

// State pointers of the device view: (member, type, field name).  The field offsets are looked up by name once per
// handle (wg_create); binding a state tensor per call is then plain pointer arithmetic -- this sits on the per-step
// host path of wg_step / wg_step_host.
#define WG_STATE_MEMBERS(X)                                                                                              \
  X(prof, float, "prof") X(pmut, float, "pmut") X(pcon, float, "pcon") X(head, int, "head") X(count, int, "count")      \
  X(retire, int, "retire") X(n_step, int, "n_step") X(load, int, "load") X(order, int, "order")                         \
  X(part_acc, int, "part_acc") X(part_keep, int, "part_keep") X(part_arrive, int, "part_arrive") X(work, int2, "work")  \
  X(yaw, float, "yaw") X(u, float, "u") X(v, float, "v") X(w, float, "w") X(power, float, "power") X(ct, float, "ct")   \
  X(derate, float, "derate") X(ws, float, "ws") X(ti, float, "ti") X(wd, float, "wd") X(rated, float, "rated_power")    \
  X(xmax, float, "xmax") X(knu1, float, "knu1") X(k_emit, int, "k_emit") X(time_max, int, "time_max")                   \
  X(timestep, int, "timestep") X(flags, int, "flags") X(n_push, int, "n_push") X(n_fp, int, "n_fp") X(n_bp, int, "n_bp") \
  X(spin, int, "spin") X(xr, float, "xr") X(yr, float, "yr") X(xs_sorted, float, "xs_sorted")                           \
  X(ord_sorted, int, "ord_sorted") X(meas, float, "meas") X(base_pow_mean, float, "base_pow_mean")                      \
  X(old_yaw, float, "old_yaw") X(rings, float, "rings") X(tb_off, float, "tb_off") X(tb_scale, float, "tb_scale")       \
  X(fp_ring, float, "farm_pow_ring") X(bp_ring, float, "base_pow_ring") X(fp_ring_lo, float, "farm_pow_ring_lo")

void resolve_state_offsets(wg_handle* h) {
  h->state_off.clear();
#define X(member, type, name) h->state_off.push_back(find_field(h, name)->offset);
  WG_STATE_MEMBERS(X)
#undef X
}

wg::Dev bind(const wg_handle* h, void* state) {
  wg::Dev d = h->dev;
  unsigned char* base = reinterpret_cast<unsigned char*>(state);
  const size_t* off = h->state_off.data();
#define X(member, type, name) d.member = reinterpret_cast<type*>(base + *off++);
  WG_STATE_MEMBERS(X)
#undef X
  return d;
}

template <class Tp>
cudaError_t upload(Tp** dst, const Tp* src, size_t n) {
  cudaError_t e = cudaMalloc(dst, sizeof(Tp) * (n ? n : 1));
  if (e != cudaSuccess) return e;
  return cudaMemcpy(*dst, src, sizeof(Tp) * n, cudaMemcpyHostToDevice);
}

// Observation descriptor list: order of farm_mes.get_measurements (MesClass.py:679-703) or, for the multi-agent
// layout, of WindFarmEnvMulti._get_obs_multi (WindEnvMulti.py:79-103).
void build_desc(const wg_config& cfg, std::vector<wg::ObsDesc>& out, int& obs_dim, int& obs_rows) {
  const wg_mes_config& m = cfg.mes;
  const int T = cfg.n_turb;
  const wg_mes_channel* ch[4] = {&m.ws, &m.wd, &m.yaw, &m.power};
  const double lo[4] = {m.ws_min, m.wd_min, m.yaw_min, 0.0};
  const double hi[4] = {m.ws_max, m.wd_max, m.yaw_max, m.power_max};
  auto emit_chan = [&](std::vector<wg::ObsDesc>& v, int c, int ring, int gate, double l, double h) {
    if (!gate) return;
    wg::ObsDesc d{};
    d.ring = ring; d.chan = c; d.lo = (float)l; d.span = (float)(h - l);
    if (ch[c]->current) { d.kind = 0; d.win = 0; v.push_back(d); }
    if (ch[c]->rolling_mean)
      for (int i = 0; i < ch[c]->history_N; ++i) { d.kind = 1; d.win = i; v.push_back(d); }
  };
  auto turb_block = [&](std::vector<wg::ObsDesc>& v, int t) {  // turb_mes.get_measurements order: ws|wd|yaw|TI|power
    emit_chan(v, 0, 0 * T + t, m.turb_ws, lo[0], hi[0]);
    emit_chan(v, 1, 1 * T + t, m.turb_wd, lo[1], hi[1]);
    emit_chan(v, 2, 2 * T + t, 1, lo[2], hi[2]);
    if (m.turb_TI) {
      wg::ObsDesc d{};
      d.kind = 2; d.ring = t; d.chan = 0; d.lo = (float)m.ti_min; d.span = (float)(m.ti_max - m.ti_min);
      v.push_back(d);
    }
    emit_chan(v, 3, 3 * T + t, m.turb_power, lo[3], hi[3]);
  };
  auto fill_channel = [&](std::vector<wg::ObsDesc>& v) {
    for (auto& d : v) { d.H = ch[d.chan]->history_length; d.N = ch[d.chan]->history_N; d.W = ch[d.chan]->window_length; }
  };
  const int fr = 4 * T;  // farm rings: ws, wd, power
  if (!m.multi_agent) {
    for (int t = 0; t < T; ++t) turb_block(out, t);
    emit_chan(out, 0, fr + 0, m.farm_ws, lo[0], hi[0]);
    emit_chan(out, 1, fr + 1, m.farm_wd, lo[1], hi[1]);
    if (m.farm_TI) {
      wg::ObsDesc d{};
      d.kind = 3; d.chan = 0; d.lo = (float)m.ti_min; d.span = (float)(m.ti_max - m.ti_min);
      out.push_back(d);
    }
    emit_chan(out, 3, fr + 2, m.farm_power, lo[3], hi[3] * T);
    fill_channel(out);
    obs_rows = 1;
    obs_dim = (int)out.size();
  } else {
    // farm block = farm-level turb_mes.get_measurements(scaled=True); its yaw history never receives data
    // (MesClass.py:588-591) so it contributes nothing; its TI is calc_TI of the farm ws ring.
    std::vector<wg::ObsDesc> farm;
    emit_chan(farm, 0, fr + 0, m.farm_ws, lo[0], hi[0]);
    emit_chan(farm, 1, fr + 1, m.farm_wd, lo[1], hi[1]);
    if (m.farm_TI) {
      wg::ObsDesc d{};
      d.kind = 2; d.ring = fr + 0; d.chan = 0; d.lo = (float)m.ti_min; d.span = (float)(m.ti_max - m.ti_min);
      farm.push_back(d);
    }
    emit_chan(farm, 3, fr + 2, m.farm_power, lo[3], hi[3] * T);
    for (int t = 0; t < T; ++t) {
      turb_block(out, t);
      out.insert(out.end(), farm.begin(), farm.end());
    }
    fill_channel(out);
    obs_rows = T;
    obs_dim = (int)out.size() / T;
  }
}

}  // namespace

extern "C" {

const char* wg_last_error(void) { return g_last_error.c_str(); }
int wg_version(void) { return WG_VERSION; }

int wg_create(const wg_config* cfg, wg_handle** out) {
  if (!cfg || !out) return fail(WG_ERR_INVALID, "wg_create: null argument");
  *out = nullptr;
  if (cfg->n_envs < 1) return fail(WG_ERR_INVALID, "n_envs must be >= 1");
  if (cfg->n_turb < 1 || cfg->n_turb > WG_MAX_T) return fail(WG_ERR_INVALID, "n_turb must be in 1..64");
  if (cfg->n_farms < 1 || cfg->n_farms > 2) return fail(WG_ERR_INVALID, "n_farms must be 1 or 2");
  if (cfg->p_cap < 8 || cfg->p_cap % 8) return fail(WG_ERR_INVALID, "p_cap must be a positive multiple of 8");
  if (cfg->substeps < 1) return fail(WG_ERR_INVALID, "dt_env must be a multiple of dt_sim");
  if (!(cfg->dt > 0.f) || !(cfg->diameter > 0.f)) return fail(WG_ERR_INVALID, "dt and diameter must be positive");
  if (cfg->n_tab < 2 || !cfg->tab_ws || !cfg->tab_power || !cfg->tab_ct || !cfg->x_pos || !cfg->y_pos)
    return fail(WG_ERR_INVALID, "turbine tables / layout missing");
  if (cfg->action_method == 2) return fail(WG_ERR_UNSUPPORTED, "The absolute method is not implemented yet");
  if (cfg->action_method < 0 || cfg->action_method > 2)
    return fail(WG_ERR_INVALID, "The ActionMethod must be yaw, wind or absolute");
  if (cfg->power_reward < 0 || cfg->power_reward > 3)
    return fail(WG_ERR_INVALID, "The Power_reward must be either Baseline, Power_avg, None or Power_diff");
  if (cfg->power_reward == 3 && cfg->power_avg < 40)
    return fail(WG_ERR_INVALID, "The Power_avg must be larger then 40 for the Power_diff reward.");
  if (cfg->power_reward == 1 && cfg->n_farms != 2) return fail(WG_ERR_INVALID, "Baseline reward needs n_farms == 2");
  if (cfg->n_farms == 2 && (cfg->base_controller < 0 || cfg->base_controller > 1))
    return fail(WG_ERR_INVALID, "The BaseController must be either Local or Global... For now");
  if (cfg->power_avg < 1) return fail(WG_ERR_INVALID, "Power_avg must be >= 1");
  if (cfg->act_var < 0 || cfg->act_var > 2) return fail(WG_ERR_INVALID, "act_var must be 1 (yaw) or 2 (yaw + induction)");
  if (cfg->act_var == 2 && !(cfg->derate_min > 0.f && cfg->derate_min <= 1.f))
    return fail(WG_ERR_INVALID, "derate_min must be in (0, 1]");
  const wg_mes_channel* ch[4] = {&cfg->mes.ws, &cfg->mes.wd, &cfg->mes.yaw, &cfg->mes.power};
  for (int c = 0; c < 4; ++c)
    if (ch[c]->history_length < 1 || ch[c]->window_length < 1 || ch[c]->history_N < 1)
      return fail(WG_ERR_INVALID, "measurement history_length / window_length / history_N must be >= 1");
  if (cfg->steps_on_reset < 1) return fail(WG_ERR_INVALID, "fill_window must be True or a non-negative integer");

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(WG_ERR_NO_DEVICE, "no CUDA device visible: windgym_b200 has no CPU fallback");
  int devid = 0;
  cudaGetDevice(&devid);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, devid);
  if (prop.major != 10)
    return fail(WG_ERR_NO_DEVICE, std::string("device '") + prop.name + "' is not sm_100 (Blackwell B200) class");

  wg_handle* h = new wg_handle();
  h->cfg = *cfg;
  const int B = cfg->n_envs, T = cfg->n_turb, F = cfg->n_farms, P = cfg->p_cap;
  cudaError_t e;
  if ((e = upload(&h->d_tab_ws, cfg->tab_ws, cfg->n_tab)) != cudaSuccess ||
      (e = upload(&h->d_tab_p, cfg->tab_power, cfg->n_tab)) != cudaSuccess ||
      (e = upload(&h->d_tab_ct, cfg->tab_ct, cfg->n_tab)) != cudaSuccess ||
      (e = upload(&h->d_x, cfg->x_pos, T)) != cudaSuccess || (e = upload(&h->d_y, cfg->y_pos, T)) != cudaSuccess) {
    wg_destroy(h);
    return cuda_fail(e, "wg_create upload");
  }
  h->cfg.tab_ws = h->cfg.tab_power = h->cfg.tab_ct = nullptr;
  h->cfg.x_pos = h->cfg.y_pos = nullptr;

  // ring table: turbine rings c*T+t (c = ws, wd, yaw, power), then farm ws / wd / power
  std::vector<int> ring_off, ring_chan;
  int off = 0;
  for (int c = 0; c < 4; ++c)
    for (int t = 0; t < T; ++t) { ring_off.push_back(off); ring_chan.push_back(c); off += ch[c]->history_length; }
  const int fch[3] = {0, 1, 3};
  for (int k = 0; k < 3; ++k) { ring_off.push_back(off); ring_chan.push_back(fch[k]); off += ch[fch[k]]->history_length; }
  off = (off + 3) / 4 * 4;  // rows of the ring block 16-byte aligned: the finish kernel stages them with 16-byte copies
  std::vector<wg::ObsDesc> desc;
  int obs_dim = 0, obs_rows = 1;
  build_desc(*cfg, desc, obs_dim, obs_rows);
  for (auto& od : desc) od.off = od.kind == 3 ? 0 : ring_off[od.ring];
  if ((e = upload(&h->d_ring_off, ring_off.data(), ring_off.size())) != cudaSuccess ||
      (e = upload(&h->d_ring_chan, ring_chan.data(), ring_chan.size())) != cudaSuccess ||
      (e = upload(&h->d_desc, desc.data(), desc.size())) != cudaSuccess) {
    wg_destroy(h);
    return cuda_fail(e, "wg_create upload");
  }
  // rotor quadrature: 4 equal-area rings x 4 azimuths (oracle/dwm_numpy.py:rotor_points)
  float qy[WG_NQ], qz[WG_NQ];
  for (int k = 0; k < 4; ++k)
    for (int m = 0; m < 4; ++m) {
      const double rho = std::sqrt((k + 0.5) / 4.0);
      const double phi = 2.0 * M_PI * (m + 0.5 * (k & 1)) / 4.0 + M_PI / 8.0;
      qy[4 * k + m] = (float)(rho * std::cos(phi));
      qz[4 * k + m] = (float)(rho * std::sin(phi));
    }
  wg::set_rotor_points(qy, qz);

  wg::Dev& d = h->dev;
  memset(&d, 0, sizeof(d));
  d.B = B; d.Bg = B; d.T = T; d.F = F; d.P = P; d.S = cfg->substeps; d.n_tab = cfg->n_tab;
  h->n_active = B;
  const char* no_order = getenv("WG_NO_ORDER");
  h->use_order = !(no_order && no_order[0] == '1');
  d.dt = cfg->dt; d.D = cfg->diameter; d.R = 0.5f * cfg->diameter; d.zh = cfg->hub_height; d.d_particle = cfg->d_particle;
  d.yaw_min = cfg->yaw_min; d.yaw_max = cfg->yaw_max; d.yaw_step = cfg->yaw_step;
  d.action_method = cfg->action_method; d.base_controller = cfg->base_controller;
  d.act_var = cfg->act_var == 2 ? 2 : 1; d.derate_min = cfg->derate_min;
  d.power_reward = cfg->power_reward; d.power_avg = cfg->power_avg; d.pen_type = cfg->action_penalty_type;
  d.power_scaling = cfg->power_scaling; d.action_penalty = cfg->action_penalty;
  d.n_rings = (int)ring_off.size(); d.ring_floats = off; d.obs_dim = obs_dim; d.obs_rows = obs_rows;
  for (int c = 0; c < 4; ++c) {
    d.ch_cur[c] = ch[c]->current; d.ch_roll[c] = ch[c]->rolling_mean; d.ch_N[c] = ch[c]->history_N;
    d.ch_H[c] = ch[c]->history_length; d.ch_W[c] = ch[c]->window_length;
    d.noise_std[c] = cfg->mes.noise_std[c];
  }
  d.noise = cfg->mes.noise; d.noise_seed = cfg->mes.noise_seed;
  bool ti_obs = false;
  for (const auto& od : desc) ti_obs |= od.kind >= 2;
  d.fin_lean = (!cfg->mes.noise && !ti_obs && cfg->power_reward != 3) ? 1 : 0;
  d.ti_lo = (float)cfg->mes.ti_min; d.ti_span = (float)(cfg->mes.ti_max - cfg->mes.ti_min);
  d.ring_off = h->d_ring_off; d.ring_chan = h->d_ring_chan; d.obs_desc = h->d_desc;
  for (int c = 0; c < 4; ++c) d.ch_base[c] = ring_off[c * T];
  for (int k = 0; k < 3; ++k) d.farm_off[k] = ring_off[4 * T + k];
  d.tab_ws = h->d_tab_ws; d.tab_p = h->d_tab_p; d.tab_ct = h->d_tab_ct; d.x_pos = h->d_x; d.y_pos = h->d_y;
  h->hist_max = std::max(std::max(cfg->mes.ws.history_length, cfg->mes.wd.history_length), cfg->mes.yaw.history_length);

  add_field(h, "prof", 0, {B, F, T, P, WG_NR});
  add_field(h, "pmut", 0, {2, B, F, T, P, 4});
  add_field(h, "pcon", 0, {B, F, T, P, 4});
  add_field(h, "head", 1, {B, F, T});
  add_field(h, "count", 1, {B, F, T});
  add_field(h, "retire", 1, {B, F, T});
  add_field(h, "n_step", 1, {B, F});
  add_field(h, "load", 1, {B, F});
  for (const char* n : {"yaw", "u", "v", "w", "power", "ct", "derate"}) add_field(h, n, 0, {B, F, T});
  for (const char* n : {"ws", "ti", "wd", "rated_power", "xmax", "knu1", "base_pow_mean"}) add_field(h, n, 0, {B});
  for (const char* n : {"k_emit", "time_max", "timestep", "flags", "n_push", "n_fp", "n_bp", "spin"})
    add_field(h, n, 1, {B});
  add_field(h, "xr", 0, {B, T});
  add_field(h, "yr", 0, {B, T});
  add_field(h, "xs_sorted", 0, {B, T});
  add_field(h, "ord_sorted", 1, {B, T});
  add_field(h, "meas", 0, {B, 4, T});
  add_field(h, "old_yaw", 0, {B, T});
  add_field(h, "tb_off", 0, {B, 3});
  add_field(h, "tb_scale", 0, {B});
  add_field(h, "rings", 0, {B, off});
  add_field(h, "farm_pow_ring", 0, {B, cfg->power_avg});
  add_field(h, "farm_pow_ring_lo", 0, {B, cfg->power_avg});
  add_field(h, "base_pow_ring", 0, {B, cfg->power_avg});
  // field table of the env-copy kernel: leading dimension B, or [2, B, ...] for the ping-pong buffer
  std::vector<wg::CopyField> cf;
  for (auto& f : h->fields) {
    size_t total = 4;
    for (auto sdim : f.shape) total *= (size_t)sdim;
    wg::CopyField c{};
    c.offset = f.offset;
    if (f.shape[0] == B) { c.n_rep = 1; c.rep_stride = 0; c.per_env = (unsigned)(total / B); }
    else { c.n_rep = (unsigned)f.shape[0]; c.rep_stride = total / f.shape[0]; c.per_env = (unsigned)(total / f.shape[0] / B); }
    cf.push_back(c);
  }
  h->n_copy = (int)cf.size();
  add_field(h, "order", 1, {B});  // a permutation of the active envs, not per-env state: outside the copy table
  // scratch of split farms (all zeros between launches) and the work table: not per-env state either
  add_field(h, "part_acc", 1, {B, F, 5, T});
  add_field(h, "part_keep", 1, {B, F, T});
  add_field(h, "part_arrive", 1, {B, F});
  h->work_cap = std::max(B * F * WG_MAX_PARTS, 4096);
  add_field(h, "work", 1, {h->work_cap, 2});
  // device-side spare pool (wg_pool_*): slot status, RNG generation, the step's swap list, counters, refill masks and
  // the reset arguments of the slots being refilled
  add_field(h, "pool_status", 1, {B});
  add_field(h, "pool_gen", 1, {B});
  add_field(h, "pool_swap", 1, {2 * WG_POOL_MAX_SWAP + 2});
  add_field(h, "pool_stats", 1, {16});
  add_field(h, "pool_masks", 1, {WG_POOL_MASKS, (B + 3) / 4});
  for (const char* n : {"pool_ws", "pool_ti", "pool_ti_flow", "pool_wd", "pool_rated", "pool_tb_scale"}) add_field(h, n, 0, {B});
  add_field(h, "pool_yaw0", 0, {B, T});
  add_field(h, "pool_tb_off", 0, {B, 3});
  for (const char* n : {"pool_k_emit", "pool_t_dev", "pool_time_max"}) add_field(h, n, 1, {B});
  const char* no_br = getenv("WG_NO_BRICKS");
  h->use_bricks = !(no_br && no_br[0] == '1');
  const char* f_br = getenv("WG_FORCE_BRICKS");
  h->force_bricks = f_br && f_br[0] == '1';
  if (const char* mt = getenv("WG_MAX_PART_TILES")) h->max_part_tiles = std::max(0, atoi(mt));
  if (const char* fw = getenv("WG_FILL_WAVES")) h->fill_waves = std::max(0, atoi(fw));
  if (const char* ss = getenv("WG_SLOT_SHARE")) h->slot_share = (float)atof(ss);
  const char* no_tw = getenv("WG_NO_TWOWAVE");
  h->two_wave = !(no_tw && no_tw[0] == '1');
  const char* no_pl = std::getenv("WG_NO_PDL_LATE");
  h->pdl_late = !(no_pl && no_pl[0] == '1');
  const char* no_pn = std::getenv("WG_NO_PDL_NEXT");
  h->pdl_next = !(no_pn && no_pn[0] == '1');
  const char* no_pdl = getenv("WG_NO_PDL");
  h->use_pdl = !(no_pdl && no_pdl[0] == '1');
  const char* no_zc = getenv("WG_NO_ZEROCOPY");
  h->use_zero_copy = !(no_zc && no_zc[0] == '1');
  const char* no_split = getenv("WG_NO_SPLIT");
  h->use_split = !(no_split && no_split[0] == '1');
  if (const char* tu = getenv("WG_TAIL_UNITS")) h->tail_units = std::max(0, atoi(tu));
  if (const char* tp = getenv("WG_TAIL_PARTS")) h->tail_parts = std::min(std::max(1, atoi(tp)), WG_MAX_PARTS);
  h->state_bytes = (h->state_bytes + 255) / 256 * 256;
  if ((e = upload(&h->d_copy, cf.data(), cf.size())) != cudaSuccess) {
    wg_destroy(h);
    return cuda_fail(e, "wg_create upload");
  }
  resolve_state_offsets(h);
  *out = h;
  return WG_OK;
}

void wg_destroy(wg_handle* h) {
  if (!h) return;
  for (cudaEvent_t e : h->prof_events) cudaEventDestroy(e);
  cudaFree(h->d_tab_ws); cudaFree(h->d_tab_p); cudaFree(h->d_tab_ct); cudaFree(h->d_x); cudaFree(h->d_y);
  cudaFree(h->d_ring_off); cudaFree(h->d_ring_chan); cudaFree(h->d_desc); cudaFree(h->d_copy);
  cudaFree(h->d_done_count);
  cudaFree(h->d_lp8);
  cudaFree(h->d_raw8);
  if (h->need_host) cudaFreeHost(h->need_host);
  if (h->flag_host) cudaFreeHost(h->flag_host);
  delete h;
}

int wg_state_bytes(const wg_handle* h, size_t* out) {
  if (!h || !out) return fail(WG_ERR_INVALID, "wg_state_bytes: null argument");
  *out = h->state_bytes;
  return WG_OK;
}

int wg_obs_dim(const wg_handle* h, int32_t* out) {
  if (!h || !out) return fail(WG_ERR_INVALID, "wg_obs_dim: null argument");
  *out = h->dev.obs_dim;
  return WG_OK;
}

int wg_state_field(const wg_handle* h, const char* name, size_t* offset, int32_t* dtype, int32_t* ndim, int64_t shape[8]) {
  if (!h || !name || !offset || !dtype || !ndim || !shape) return fail(WG_ERR_INVALID, "wg_state_field: null argument");
  const Field* f = find_field(h, name);
  if (!f) return fail(WG_ERR_INVALID, std::string("unknown state field '") + name + "'");
  *offset = f->offset; *dtype = f->dtype; *ndim = (int32_t)f->shape.size();
  for (size_t i = 0; i < f->shape.size(); ++i) shape[i] = f->shape[i];
  return WG_OK;
}

const char* wg_state_field_name(const wg_handle* h, int32_t index) {
  if (!h || index < 0 || index >= (int32_t)h->fields.size()) return nullptr;
  return h->fields[index].name.c_str();
}

int wg_launch_count(const wg_handle* h, uint64_t* out) {
  if (!h || !out) return fail(WG_ERR_INVALID, "wg_launch_count: null argument");
  *out = h->launches;
  return WG_OK;
}

#define WG_LAUNCH(expr, what)                                   \
  do {                                                          \
    cudaError_t e_ = (expr);                                    \
    ++h->launches;                                              \
    if (e_ != cudaSuccess) return cuda_fail(e_, what);          \
  } while (0)

#define WG_ORDER_PERIOD 64

int wg_flow_steps(wg_handle* h, void* state, int32_t n_steps, void* cuda_stream) {
  if (!h || !state) return fail(WG_ERR_INVALID, "wg_flow_steps: null argument");
  if (n_steps < 0) return fail(WG_ERR_INVALID, "n_steps must be >= 0");
  if (n_steps == 0) return WG_OK;
  wg::Dev d = bind(h, state);
  wg::FlowArgs fa{};
  fa.mode = wg::FLOW_FIXED; fa.n_fixed = n_steps; fa.farm_mask = (1 << d.F) - 1;
  WG_LAUNCH(wg::launch_flow(d, fa, (cudaStream_t)cuda_stream), "wg_flow_kernel");
  return WG_OK;
}

// WindFarmEnv.reset for the masked envs of the slot range [b0, b0 + nb)
static int reset_impl(wg_handle* h, void* state, const wg::ResetDevArgs& ra, float* obs, cudaStream_t s, int b0, int nb) {
  pin_turbulence_in_l2(h, s);
  wg::Dev d = bind(h, state);
  d.b0 = b0; d.Bg = nb;
  WG_LAUNCH(wg::launch_reset_init(d, ra, s), "wg_reset_init_kernel");
  // fs.run(t_developed) for the agent farm and the baseline farm (Wind_Farm_Env.py:734, :782)
  wg::FlowArgs spin{};
  spin.mode = wg::FLOW_SPIN; spin.mask = ra.mask; spin.farm_mask = (1 << d.F) - 1;
  WG_LAUNCH(wg::launch_flow(d, spin, s), "wg_flow_kernel(spin-up)");
  // measurement fill: steps_on_reset env steps for the agent farm (:737-766), hist_max for the baseline farm,
  // whose controller stays off during the fill (:784-796)
  const int n_agent = h->cfg.steps_on_reset, n_base = d.F > 1 ? h->hist_max : 0;
  for (int i = 0; i < std::max(n_agent, n_base); ++i) {
    const int fm = (i < n_agent ? 1 : 0) | (i < n_base ? 2 : 0);
    wg::FlowArgs fa{};
    fa.mode = wg::FLOW_STEP; fa.mask = ra.mask; fa.farm_mask = fm;
    WG_LAUNCH(wg::launch_flow(d, fa, s), "wg_flow_kernel(fill)");
    wg::FinishArgs fin{};
    fin.mask = ra.mask;
    fin.flags = ((fm & 1) ? (wg::FIN_PUSH_MES | wg::FIN_PUSH_FP) : 0) | ((fm & 2) ? wg::FIN_PUSH_BP : 0);
    WG_LAUNCH(wg::launch_finish(d, fin, s), "wg_finish_kernel(fill)");
  }
  wg::FinishArgs fin{};
  fin.mask = ra.mask; fin.flags = wg::FIN_OBS; fin.obs = obs;
  WG_LAUNCH(wg::launch_finish(d, fin, s), "wg_finish_kernel(obs)");
  return WG_OK;
}

int wg_reset(wg_handle* h, void* state, const wg_reset_args* args, float* obs, void* cuda_stream) {
  if (!h || !state || !args || !obs) return fail(WG_ERR_INVALID, "wg_reset: null argument");
  if (!args->ws || !args->ti_flow || !args->wd || !args->yaw0 || !args->rated_power || !args->k_emit ||
      !args->t_developed || !args->time_max)
    return fail(WG_ERR_INVALID, "wg_reset: every per-env input array is required");
  if (h->dev.tb_raw && (!args->tb_offset || !args->tb_scale))
    return fail(WG_ERR_INVALID, "wg_reset: a handle with a turbulence box needs tb_offset and tb_scale");
  h->order_state = nullptr;
  wg::ResetDevArgs ra{args->mask, args->ws, args->ti_flow, args->wd, args->yaw0, args->rated_power,
                      args->k_emit, args->t_developed, args->time_max, args->tb_offset, args->tb_scale};
  return reset_impl(h, state, ra, obs, (cudaStream_t)cuda_stream, 0, h->cfg.n_envs);
}

// host_out: optional mapped host copies of the results + completion flag (wg_step_host's zero-copy path)
static int step_impl(wg_handle* h, void* state, const float* actions, float* obs, float* reward, uint8_t* truncated,
                     void* cuda_stream, const wg::FinishArgs* host_out) {
  cudaStream_t s = (cudaStream_t)cuda_stream;
  pin_turbulence_in_l2(h, s);
  wg::Dev d = bind(h, state);
  d.Bg = h->n_active;
  // launch order / work table: rebuilt after a reset / env copy / change of the active set, and every WG_ORDER_PERIOD
  // steps (the live-station counts drift slowly).  Single-substep steps go by the work table (farms cut into parts
  // when the batch leaves CTA slots free, or at the tail of the grid); others by the per-env order.
  const bool by_table = h->use_split && d.S == 1;
  if ((h->use_order || by_table) && (h->order_state != state || h->order_n != d.Bg || ++h->order_age >= WG_ORDER_PERIOD)) {
    if (by_table) {
      if (h->slots <= 0) {
        h->slots = wg::flow_resident_ctas(d);
        // part of the machine is permanently taken by other work of the caller (background spin-ups of a spare pool:
        // wg_set_slot_share): plan the stepping grid for the share that is left, so that it stays ONE wave
        if (h->slot_share > 0.f && h->slot_share < 1.f) h->slots = std::max(1, (int)(h->slots * h->slot_share));
      }
      const int U = d.Bg * d.F;
      wg::PlanArgs pa{};
      pa.slots = h->slots;
      // large farms: parts of at most max_part_tiles tiles; the table then holds up to `fan` entries per farm
      int fan = 1;
      if (d.T > 16 && h->max_part_tiles > 0) {
        pa.max_tiles = h->max_part_tiles;
        fan = std::min(WG_MAX_PARTS, (d.T * d.P / WG_TILE + pa.max_tiles - 1) / pa.max_tiles);
      }
      if (U < h->slots) {
        pa.target = pa.n_work = h->slots;
      } else if (U < 2 * h->slots && h->two_wave) {
        pa.target = pa.n_work = 2 * h->slots;
      } else if (h->fill_waves > 0 && U < h->fill_waves * h->slots) {   // experiment: fill the last wave (WG_FILL_WAVES)
        pa.target = pa.n_work = (U + h->slots - 1) / h->slots * h->slots;
      } else {
        pa.tail_units = std::min(h->tail_units, U);
        pa.tail_parts = pa.tail_units > 0 ? h->tail_parts : 1;
        pa.n_work = U + pa.tail_units * (pa.tail_parts - 1);
      }
      if (fan > 1) pa.n_work += U * fan;  // sum of max(a, b) <= sum a + sum b: room for both rules
      pa.n_work = std::min(pa.n_work, h->work_cap);
      if (pa.n_work < U) return fail(WG_ERR_INVALID, "wg_step: work table too small");
      h->n_work = pa.n_work;
      WG_LAUNCH(wg::launch_plan(d, pa, s), "wg_plan_kernel");
    } else {
      WG_LAUNCH(wg::launch_order(d, s), "wg_order_kernel");
    }
    h->order_state = state; h->order_n = d.Bg; h->order_age = 0;
  }
  cudaEvent_t* ev = nullptr;
  if (h->profiling) {
    if (h->prof_used + 3 > h->prof_events.size()) {
      for (int i = 0; i < 3; ++i) {
        cudaEvent_t e;
        cudaError_t ce = cudaEventCreate(&e);
        if (ce != cudaSuccess) return cuda_fail(ce, "cudaEventCreate");
        h->prof_events.push_back(e);
      }
    }
    ev = &h->prof_events[h->prof_used];
    h->prof_used += 3;
    cudaEventRecord(ev[0], s);
  }
  wg::FlowArgs fa{};
  fa.mode = wg::FLOW_STEP; fa.actions = actions; fa.farm_mask = (1 << d.F) - 1; fa.controller_on = 1;
  // fewer farms than CTA slots: overlap the finish kernel's launch + staging with the flow grid's tail (PDL)
  const bool pdl = by_table && h->use_pdl && d.Bg * d.F < h->slots && !h->profiling;
  if (by_table) { fa.work = d.work; fa.n_work = h->n_work; fa.pdl_trigger = pdl ? 1 : 0; }
  else fa.order = h->use_order ? d.order : nullptr;
  // more farms than slots: the finish kernel is released behind the CTAs' tile loops
  const bool pdl_late = !pdl && h->use_pdl && h->pdl_late && !h->profiling;
  if (pdl_late) fa.pdl_trigger = 2;
  // the flow kernel of this step as programmatic dependent of the previous step's finish kernel (which triggers once
  // it holds the flow results): prologue + tile loop overlap it.  Behind any other kernel it is an ordinary launch.
  const bool pdl_chain = h->use_pdl && h->pdl_next && !h->profiling;
  // (a caller that waits for every step's results on the host -- wg_step_host -- has nothing to overlap: plain launch)
  fa.pdl_wait = (pdl_chain && !host_out) ? (h->after_swap ? 3 : 1) : 0;
  h->after_swap = false;
  WG_LAUNCH(wg::launch_flow(d, fa, s), "wg_flow_kernel(step)");
  if (ev) cudaEventRecord(ev[1], s);
  wg::FinishArgs fin{};
  fin.flags = wg::FIN_PUSH_MES | wg::FIN_PUSH_FP | (d.F > 1 ? wg::FIN_PUSH_BP : 0) | wg::FIN_OBS | wg::FIN_REWARD;
  fin.obs = obs; fin.reward = reward; fin.truncated = truncated;
  fin.pdl = (pdl || pdl_late) ? 1 : 0;
  fin.trigger = pdl_chain ? 1 : 0;
  if (host_out) {
    fin.obs_h = host_out->obs_h; fin.reward_h = host_out->reward_h; fin.truncated_h = host_out->truncated_h;
    fin.done_count = host_out->done_count; fin.done_flag = host_out->done_flag; fin.seq = host_out->seq;
  }
  WG_LAUNCH(wg::launch_finish(d, fin, s), "wg_finish_kernel(step)");
  if (ev) cudaEventRecord(ev[2], s);
  return WG_OK;
}

int wg_step(wg_handle* h, void* state, const float* actions, float* obs, float* reward, uint8_t* truncated,
            void* cuda_stream) {
  if (!h || !state || !actions || !obs || !reward || !truncated) return fail(WG_ERR_INVALID, "wg_step: null argument");
  return step_impl(h, state, actions, obs, reward, truncated, cuda_stream, nullptr);
}

int wg_result_bytes(const wg_handle* h, size_t* out) {
  if (!h || !out) return fail(WG_ERR_INVALID, "wg_result_bytes: null argument");
  const size_t B = (size_t)h->cfg.n_envs;
  *out = B * (size_t)h->dev.obs_rows * (size_t)h->dev.obs_dim * 4 + B * 4 + B;
  return WG_OK;
}

// Device alias of a host pointer if [p, p + bytes) lies in pinned (page-locked, mapped) host memory, else null.
// Ranges found once are remembered per handle (a lookup costs a driver call otherwise).
static char* pinned_alias(wg_handle* h, const void* p, size_t bytes) {
  const char* c = reinterpret_cast<const char*>(p);
  for (const auto& r : h->pinned)
    if (c >= r.host && c + bytes <= r.host + r.bytes) return r.dev + (c - r.host);
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess || at.type != cudaMemoryTypeHost || !at.devicePointer) {
    cudaGetLastError();
    return nullptr;
  }
  // the whole allocation, so that later sub-ranges (another row of the same pinned action buffer) hit the cache
  CUdeviceptr base = 0;
  size_t len = 0;
  char* dev = reinterpret_cast<char*>(at.devicePointer);
  // driver entry point resolved through the runtime: the library carries no link-time dependency on libcuda
  typedef CUresult (*range_fn)(CUdeviceptr*, size_t*, CUdeviceptr);
  static range_fn get_range = nullptr;
  static bool looked_up = false;
  if (!looked_up) {
    looked_up = true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      get_range = reinterpret_cast<range_fn>(fn);
    cudaGetLastError();
  }
  if (get_range && get_range(&base, &len, (CUdeviceptr)at.devicePointer) == CUDA_SUCCESS && base && len) {
    const size_t off = (size_t)((CUdeviceptr)at.devicePointer - base);
    if (off + bytes > len) return nullptr;
    if (h->pinned.size() < 64) h->pinned.push_back({c - off, dev - off, len});
  } else if (h->pinned.size() < 64) {
    h->pinned.push_back({c, dev, bytes});
  }
  return dev;
}

int wg_step_host(wg_handle* h, void* state, const float* actions_host, float* actions_dev, void* out_dev, void* out_host,
                 size_t out_bytes, void* cuda_stream) {
  if (!h || !state || !actions_host || !actions_dev || !out_dev || !out_host)
    return fail(WG_ERR_INVALID, "wg_step_host: null argument");
  size_t want = 0;
  wg_result_bytes(h, &want);
  if (out_bytes != want) return fail(WG_ERR_INVALID, "wg_step_host: out_bytes must equal wg_result_bytes");
  cudaStream_t s = (cudaStream_t)cuda_stream;
  const size_t B = (size_t)h->cfg.n_envs;
  const size_t act_bytes = (size_t)h->n_active * (size_t)h->cfg.n_turb * (size_t)h->dev.act_var * sizeof(float);
  unsigned char* o = reinterpret_cast<unsigned char*>(out_dev);
  const size_t obs_bytes = B * (size_t)h->dev.obs_rows * (size_t)h->dev.obs_dim * 4;
  cudaError_t e;

  // Zero-copy path (both host buffers pinned): the flow kernel reads the actions straight from mapped host memory,
  // the finish kernel stores obs | reward | truncated into the mapped result buffer and publishes the step number;
  // the host polls that word.  No copy engine, no stream synchronisation in the step.
  char* act_alias = h->use_zero_copy ? pinned_alias(h, actions_host, act_bytes) : nullptr;
  char* out_alias = act_alias ? pinned_alias(h, out_host, out_bytes) : nullptr;
  if (act_alias && out_alias) {
    if (!h->flag_host) {
      if ((e = cudaHostAlloc(reinterpret_cast<void**>(&h->flag_host), 64, cudaHostAllocMapped)) != cudaSuccess)
        return cuda_fail(e, "wg_step_host: cudaHostAlloc");
      h->flag_host[0] = 0;
      if ((e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->flag_dev), h->flag_host, 0)) != cudaSuccess)
        return cuda_fail(e, "wg_step_host: cudaHostGetDevicePointer");
      if ((e = cudaMalloc(&h->d_done_count, sizeof(unsigned))) != cudaSuccess ||
          (e = cudaMemset(h->d_done_count, 0, sizeof(unsigned))) != cudaSuccess)
        return cuda_fail(e, "wg_step_host: completion counter");
    }
    wg::FinishArgs ho{};
    ho.obs_h = reinterpret_cast<float*>(out_alias);
    ho.reward_h = reinterpret_cast<float*>(out_alias + obs_bytes);
    ho.truncated_h = reinterpret_cast<uint8_t*>(out_alias + obs_bytes + B * 4);
    ho.done_count = h->d_done_count;
    ho.done_flag = h->flag_dev;
    ho.seq = ++h->seq;
    if (ho.seq == 0) ho.seq = ++h->seq;  // 0 is the flag's rest value
    const int rc = step_impl(h, state, reinterpret_cast<const float*>(act_alias), reinterpret_cast<float*>(o),
                             reinterpret_cast<float*>(o + obs_bytes), o + obs_bytes + B * 4, cuda_stream, &ho);
    if (rc != WG_OK) return rc;
    volatile unsigned* flag = h->flag_host;
    unsigned spins = 0;
    const auto t0 = std::chrono::steady_clock::now();
    while (*flag != ho.seq) {
      WG_CPU_RELAX();
      if ((++spins & 0x3fffu) == 0) {  // every ~16k polls: did the stream die, or is this taking absurdly long?
        e = cudaStreamQuery(s);
        if (e != cudaSuccess && e != cudaErrorNotReady) return cuda_fail(e, "wg_step_host: stream failed");
        if (e == cudaSuccess && *flag != ho.seq)
          return fail(WG_ERR_CUDA, "wg_step_host: stream drained without publishing the step (lost completion flag)");
        if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(30))
          return fail(WG_ERR_CUDA, "wg_step_host: no completion after 30 s");
      }
    }
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
    return WG_OK;
  }

  // pageable host memory: stage through the copy engines and synchronise the stream
  if ((e = cudaMemcpyAsync(actions_dev, actions_host, act_bytes, cudaMemcpyHostToDevice, s)) != cudaSuccess)
    return cuda_fail(e, "wg_step_host: actions H2D");
  const int rc = wg_step(h, state, actions_dev, reinterpret_cast<float*>(o), reinterpret_cast<float*>(o + obs_bytes),
                         o + obs_bytes + B * 4, cuda_stream);
  if (rc != WG_OK) return rc;
  if ((e = cudaMemcpyAsync(out_host, out_dev, out_bytes, cudaMemcpyDeviceToHost, s)) != cudaSuccess)
    return cuda_fail(e, "wg_step_host: results D2H");
  if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return cuda_fail(e, "wg_step_host: cudaStreamSynchronize");
  return WG_OK;
}

int wg_set_turbulence(wg_handle* h, const float* raw_uvw0, const float* lp_vw, int32_t nx, int32_t ny, int32_t nz,
                      float dx, float dy, float dz) {
  if (!h) return fail(WG_ERR_INVALID, "wg_set_turbulence: null argument");
  wg::Dev& d = h->dev;
  if (!raw_uvw0 && !lp_vw) {  // back to uniform inflow
    d.tb_raw = nullptr; d.tb_lp = nullptr; d.tb2_raw = nullptr; d.tb_lp8 = nullptr; d.tb_raw8 = nullptr;
    cudaFree(h->d_lp8); h->d_lp8 = nullptr;
    cudaFree(h->d_raw8); h->d_raw8 = nullptr;
    h->slots = 0; h->order_state = nullptr;
    return WG_OK;
  }
  if (!raw_uvw0 || !lp_vw) return fail(WG_ERR_INVALID, "wg_set_turbulence: both box layouts are required");
  if (nx < 2 || ny < 2 || nz < 2 || !(dx > 0.f) || !(dy > 0.f) || !(dz > 0.f))
    return fail(WG_ERR_INVALID, "wg_set_turbulence: box needs >= 2 cells and positive spacing per axis");
  if (((uintptr_t)raw_uvw0 & 15) || ((uintptr_t)lp_vw & 7))
    return fail(WG_ERR_INVALID, "wg_set_turbulence: raw must be 16-byte and lp 8-byte aligned");
  d.tb_raw = reinterpret_cast<const float4*>(raw_uvw0);
  d.tb_lp = reinterpret_cast<const float2*>(lp_vw);
  d.tb_n[0] = nx; d.tb_n[1] = ny; d.tb_n[2] = nz;
  d.tb_inv_d[0] = 1.f / dx; d.tb_inv_d[1] = 1.f / dy; d.tb_inv_d[2] = 1.f / dz;
  d.tb_inv_n[0] = 1.f / nx; d.tb_inv_n[1] = 1.f / ny; d.tb_inv_n[2] = 1.f / nz;
  d.tb_len_x = (float)((double)nx * (double)dx);
  h->tb_lp_bytes = (size_t)nx * ny * nz * sizeof(float2);
  h->policy_streams.clear();
  h->slots = 0; h->order_state = nullptr;  // another kernel variant: resident-CTA count and work table are stale
  // The low-pass layout is re-laid as 64-byte bricks (8 trilinear corners per cell, 8x the memory -- HBM is
  // plentiful): one aligned read per wake-centre sample instead of 4-8 scattered 32-byte sectors.  Measured on cfg 2:
  // reference box 2048x512x64 1.39 -> 0.635 ms per launch; even the reduced test box (whose compact layout fits the
  // persisting-L2 window of pin_turbulence_in_l2) 0.677 -> 0.608 ms.  The compact layout is the fallback when the
  // bricks do not fit in memory (and WG_NO_BRICKS=1).
  cudaFree(h->d_lp8);
  cudaFree(h->d_raw8);
  h->d_lp8 = nullptr; d.tb_lp8 = nullptr;
  h->d_raw8 = nullptr; d.tb_raw8 = nullptr;
  int devid = 0, max_persist = 0;
  cudaGetDevice(&devid);
  cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, devid);
  (void)max_persist;
  if (h->use_bricks) {
    if (cudaMalloc(&h->d_lp8, h->tb_lp_bytes * 8) == cudaSuccess) {
      cudaError_t e = wg::launch_bricks(d.tb_lp, h->d_lp8, nx, ny, nz, 0);
      if (e == cudaSuccess) e = cudaStreamSynchronize(0);
      if (e != cudaSuccess) return cuda_fail(e, "wg_set_turbulence: brick layout");
      d.tb_lp8 = h->d_lp8;
    } else {
      cudaGetLastError();  // not enough memory for the bricks: keep gathering from the compact layout
    }
    // ... and the raw box (sampled at every rotor's 16 quadrature points each step) as 128-byte bricks of 8 x (u, v, w, 0)
    const size_t raw_bytes = (size_t)nx * ny * nz * sizeof(float4);
    if (cudaMalloc(&h->d_raw8, raw_bytes * 8) == cudaSuccess) {
      cudaError_t e = wg::launch_raw_bricks(d.tb_raw, h->d_raw8, nx, ny, nz, 0);
      if (e == cudaSuccess) e = cudaStreamSynchronize(0);
      if (e != cudaSuccess) return cuda_fail(e, "wg_set_turbulence: raw brick layout");
      d.tb_raw8 = h->d_raw8;
    } else {
      cudaGetLastError();
    }
  }
  return WG_OK;
}

int wg_set_added_turbulence(wg_handle* h, const float* iso_uvw0, int32_t nx, int32_t ny, int32_t nz, float dx, float dy,
                            float dz, float k_m1, float k_m2) {
  if (!h) return fail(WG_ERR_INVALID, "wg_set_added_turbulence: null argument");
  wg::Dev& d = h->dev;
  if (!iso_uvw0) {
    d.tb2_raw = nullptr;
    h->slots = 0; h->order_state = nullptr;
    return WG_OK;
  }
  if (!d.tb_raw) return fail(WG_ERR_INVALID, "wg_set_added_turbulence: attach the ambient box (wg_set_turbulence) first");
  if (nx < 2 || ny < 2 || nz < 2 || !(dx > 0.f) || !(dy > 0.f) || !(dz > 0.f))
    return fail(WG_ERR_INVALID, "wg_set_added_turbulence: box needs >= 2 cells and positive spacing per axis");
  if ((uintptr_t)iso_uvw0 & 15) return fail(WG_ERR_INVALID, "wg_set_added_turbulence: box must be 16-byte aligned");
  d.tb2_raw = reinterpret_cast<const float4*>(iso_uvw0);
  d.tb2_n[0] = nx; d.tb2_n[1] = ny; d.tb2_n[2] = nz;
  d.tb2_inv_d[0] = 1.f / dx; d.tb2_inv_d[1] = 1.f / dy; d.tb2_inv_d[2] = 1.f / dz;
  d.tb2_inv_n[0] = 1.f / nx; d.tb2_inv_n[1] = 1.f / ny; d.tb2_inv_n[2] = 1.f / nz;
  d.tb2_len_x = (float)((double)nx * (double)dx);
  d.k_m1 = k_m1; d.k_m2 = k_m2;
  h->slots = 0; h->order_state = nullptr;
  return WG_OK;
}

int wg_set_slot_share(wg_handle* h, float share) {
  if (!h) return fail(WG_ERR_INVALID, "wg_set_slot_share: null argument");
  if (!(share > 0.f && share <= 1.f)) return fail(WG_ERR_INVALID, "wg_set_slot_share: share must be in (0, 1]");
  h->slot_share = share;
  h->slots = 0; h->order_state = nullptr;   // re-plan at the next step
  return WG_OK;
}

int wg_set_active(wg_handle* h, int32_t n_active) {
  if (!h) return fail(WG_ERR_INVALID, "wg_set_active: null argument");
  if (n_active < 1 || n_active > h->cfg.n_envs) return fail(WG_ERR_INVALID, "n_active must be in 1..n_envs");
  h->n_active = n_active;
  return WG_OK;
}

int wg_copy_envs(wg_handle* h, void* state, const int32_t* src, const int32_t* dst, int32_t n, void* cuda_stream) {
  if (!h || !state || !src || !dst) return fail(WG_ERR_INVALID, "wg_copy_envs: null argument");
  if (n < 0) return fail(WG_ERR_INVALID, "n must be >= 0");
  if (n == 0) return WG_OK;
  h->order_state = nullptr;
  WG_LAUNCH(wg::launch_copy_envs(reinterpret_cast<unsigned char*>(state), h->d_copy, h->n_copy, src, dst, n,
                                 (cudaStream_t)cuda_stream),
            "wg_copy_envs_kernel");
  return WG_OK;
}

static wg::PoolDev bind_pool(wg_handle* h, void* state) {
  if (h->pool_off.empty())
    for (const char* n : {"pool_status", "pool_gen", "pool_swap", "pool_stats", "pool_masks", "pool_ws", "pool_ti",
                          "pool_ti_flow", "pool_wd", "pool_rated", "pool_tb_scale", "pool_yaw0", "pool_tb_off",
                          "pool_k_emit", "pool_t_dev", "pool_time_max"})
      h->pool_off.push_back(find_field(h, n)->offset);
  unsigned char* b = reinterpret_cast<unsigned char*>(state);
  const size_t* o = h->pool_off.data();
  wg::PoolDev p{};
  p.status = reinterpret_cast<int*>(b + o[0]); p.gen = reinterpret_cast<int*>(b + o[1]);
  p.swap = reinterpret_cast<int*>(b + o[2]); p.stats = reinterpret_cast<unsigned long long*>(b + o[3]);
  p.masks = reinterpret_cast<uint8_t*>(b + o[4]);
  p.ws = reinterpret_cast<float*>(b + o[5]); p.ti = reinterpret_cast<float*>(b + o[6]);
  p.ti_flow = reinterpret_cast<float*>(b + o[7]); p.wd = reinterpret_cast<float*>(b + o[8]);
  p.rated = reinterpret_cast<float*>(b + o[9]); p.tb_scale = reinterpret_cast<float*>(b + o[10]);
  p.yaw0 = reinterpret_cast<float*>(b + o[11]); p.tb_off = reinterpret_cast<float*>(b + o[12]);
  p.k_emit = reinterpret_cast<int*>(b + o[13]); p.t_dev = reinterpret_cast<int*>(b + o[14]);
  p.time_max = reinterpret_cast<int*>(b + o[15]);
  p.n_active = h->n_active; p.B = h->cfg.n_envs;
  p.need_host = h->need_dev;
  return p;
}

int wg_pool_init(wg_handle* h, void* state, int32_t n_active, void* cuda_stream) {
  if (!h || !state) return fail(WG_ERR_INVALID, "wg_pool_init: null argument");
  if (n_active < 1 || n_active >= h->cfg.n_envs)
    return fail(WG_ERR_INVALID, "wg_pool_init: n_active must leave at least one spare slot (1 <= n_active < n_envs)");
  h->n_active = n_active;
  if (!h->need_host && cudaHostAlloc(reinterpret_cast<void**>(&h->need_host), 64, cudaHostAllocMapped) == cudaSuccess) {
    if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->need_dev), h->need_host, 0) != cudaSuccess) h->need_dev = nullptr;
  }
  cudaGetLastError();
  if (h->need_host) *h->need_host = h->cfg.n_envs - n_active;
  wg::PoolDev p = bind_pool(h, state);
  std::vector<int> st((size_t)p.B, wg::POOL_ACTIVE);
  for (int b = n_active; b < p.B; ++b) st[b] = wg::POOL_NEED;
  cudaStream_t s = (cudaStream_t)cuda_stream;
  cudaError_t e;
  if ((e = cudaMemcpyAsync(p.status, st.data(), sizeof(int) * st.size(), cudaMemcpyHostToDevice, s)) != cudaSuccess ||
      (e = cudaMemsetAsync(p.gen, 0, sizeof(int) * (size_t)p.B, s)) != cudaSuccess ||
      (e = cudaMemsetAsync(p.swap, 0, sizeof(int) * (2 * WG_POOL_MAX_SWAP + 2), s)) != cudaSuccess ||
      (e = cudaMemsetAsync(p.stats, 0, sizeof(int) * 16, s)) != cudaSuccess ||
      (e = cudaStreamSynchronize(s)) != cudaSuccess)
    return cuda_fail(e, "wg_pool_init");
  return WG_OK;
}

int wg_pool_refill(wg_handle* h, void* state, const wg_pool_draw* draw, float* obs, int32_t mask_row, void* cuda_stream) {
  if (!h || !state || !draw || !obs) return fail(WG_ERR_INVALID, "wg_pool_refill: null argument");
  if (mask_row < 0 || mask_row >= WG_POOL_MASKS) return fail(WG_ERR_INVALID, "wg_pool_refill: mask_row out of range");
  if (h->n_active >= h->cfg.n_envs) return fail(WG_ERR_INVALID, "wg_pool_refill: no spare slots (wg_pool_init first)");
  cudaStream_t s = (cudaStream_t)cuda_stream;
  wg::Dev d = bind(h, state);
  wg::PoolDev p = bind_pool(h, state);
  wg::PoolDraw w{};
  w.ws_min = draw->ws_min; w.ws_max = draw->ws_max; w.ti_min = draw->ti_min; w.ti_max = draw->ti_max;
  w.wd_min = draw->wd_min; w.wd_max = draw->wd_max; w.yaw_start = draw->yaw_start; w.n_passthrough = draw->n_passthrough;
  w.yaw_const = draw->yaw_const; w.yaw_random = draw->yaw_random; w.eval_mode = draw->eval_mode; w.seed = draw->seed;
  if (h->dev.tb_raw) {
    if (!(draw->tb_std_u > 0.0)) return fail(WG_ERR_INVALID, "wg_pool_refill: tb_std_u (std of the box's u) is required with a turbulence box");
    for (int k = 0; k < 3; ++k) w.tb_len[k] = (double)h->dev.tb_n[k] / (double)h->dev.tb_inv_d[k];
    w.tb_inv_std = 1.0 / draw->tb_std_u;
  }
  WG_LAUNCH(wg::launch_pool_claim(d, p, w, mask_row, s), "wg_pool_claim_kernel");
  wg::ResetDevArgs ra{p.masks + (size_t)mask_row * p.B, p.ws, p.ti_flow, p.wd, p.yaw0, p.rated,
                      p.k_emit, p.t_dev, p.time_max, p.tb_off, p.tb_scale};
  const int rc = reset_impl(h, state, ra, obs, s, p.n_active, p.B - p.n_active);
  if (rc != WG_OK) return rc;
  WG_LAUNCH(wg::launch_pool_publish(p, mask_row, s), "wg_pool_publish_kernel");
  return WG_OK;
}

int wg_pool_swap(wg_handle* h, void* state, const uint8_t* truncated, float* obs, uint8_t* swapped, float* final_obs,
                 void* cuda_stream) {
  if (!h || !state || !truncated || !obs || !swapped) return fail(WG_ERR_INVALID, "wg_pool_swap: null argument");
  if (h->n_active >= h->cfg.n_envs) return fail(WG_ERR_INVALID, "wg_pool_swap: no spare slots (wg_pool_init first)");
  cudaStream_t s = (cudaStream_t)cuda_stream;
  wg::Dev d = bind(h, state);
  wg::PoolDev p = bind_pool(h, state);
  h->after_swap = true;
  WG_LAUNCH(wg::launch_pool_swap(d, p, truncated, swapped, s), "wg_pool_swap_kernel");
  WG_LAUNCH(wg::launch_pool_copy(reinterpret_cast<unsigned char*>(state), h->d_copy, h->n_copy, p, obs, final_obs,
                                 h->dev.obs_rows * h->dev.obs_dim, s),
            "wg_pool_copy_kernel");
  return WG_OK;
}

int wg_pool_need(wg_handle* h, int32_t* out) {
  if (!h || !out) return fail(WG_ERR_INVALID, "wg_pool_need: null argument");
  *out = h->need_host ? *reinterpret_cast<volatile int*>(h->need_host) : -1;
  return WG_OK;
}

int wg_pool_stats(wg_handle* h, void* state, uint64_t out[8], void* cuda_stream) {
  if (!h || !state || !out) return fail(WG_ERR_INVALID, "wg_pool_stats: null argument");
  wg::PoolDev p = bind_pool(h, state);
  cudaError_t e = cudaMemcpyAsync(out, p.stats, sizeof(uint64_t) * 8, cudaMemcpyDeviceToHost, (cudaStream_t)cuda_stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)cuda_stream);
  if (e != cudaSuccess) return cuda_fail(e, "wg_pool_stats");
  return WG_OK;
}

int wg_flow_field(wg_handle* h, void* state, int32_t env, int32_t farm, const float* x, const float* y,
                  int32_t n_points, float z, float* out_uvw, void* cuda_stream) {
  if (!h || !state || !x || !y || !out_uvw) return fail(WG_ERR_INVALID, "wg_flow_field: null argument");
  if (env < 0 || env >= h->cfg.n_envs || farm < 0 || farm >= h->cfg.n_farms)
    return fail(WG_ERR_INVALID, "wg_flow_field: env / farm index out of range");
  if (n_points < 0) return fail(WG_ERR_INVALID, "n_points must be >= 0");
  if (n_points == 0) return WG_OK;
  wg::Dev d = bind(h, state);
  WG_LAUNCH(wg::launch_flow_field(d, env, farm, x, y, n_points, z, out_uvw, (cudaStream_t)cuda_stream),
            "wg_flow_field_kernel");
  return WG_OK;
}

int wg_profile_enable(wg_handle* h, int32_t on) {
  if (!h) return fail(WG_ERR_INVALID, "wg_profile_enable: null argument");
  h->profiling = on != 0;
  h->prof_used = 0;
  return WG_OK;
}

int wg_profile_read(wg_handle* h, double* flow_ms, double* finish_ms, uint64_t* n_steps) {
  if (!h || !flow_ms || !finish_ms || !n_steps) return fail(WG_ERR_INVALID, "wg_profile_read: null argument");
  double a = 0.0, b = 0.0;
  const size_t n = h->prof_used / 3;
  for (size_t i = 0; i < n; ++i) {
    cudaEvent_t* ev = &h->prof_events[3 * i];
    cudaError_t e = cudaEventSynchronize(ev[2]);
    if (e != cudaSuccess) return cuda_fail(e, "cudaEventSynchronize");
    float t0 = 0.f, t1 = 0.f;
    if ((e = cudaEventElapsedTime(&t0, ev[0], ev[1])) != cudaSuccess) return cuda_fail(e, "cudaEventElapsedTime");
    if ((e = cudaEventElapsedTime(&t1, ev[1], ev[2])) != cudaSuccess) return cuda_fail(e, "cudaEventElapsedTime");
    a += t0; b += t1;
  }
  *flow_ms = a; *finish_ms = b; *n_steps = n;
  h->prof_used = 0;
  return WG_OK;
}

int wg_mes_push_extract(wg_handle* h, void* state, const float* ws, const float* wd, const float* yaw,
                        const float* power, float* obs, void* cuda_stream) {
  if (!h || !state || !ws || !wd || !yaw || !power || !obs)
    return fail(WG_ERR_INVALID, "wg_mes_push_extract: null argument");
  wg::Dev d = bind(h, state);
  wg::FinishArgs fin{};
  fin.flags = wg::FIN_PUSH_MES | wg::FIN_OBS | wg::FIN_MEAS_FROM_ARGS;
  fin.in_ws = ws; fin.in_wd = wd; fin.in_yaw = yaw; fin.in_power = power; fin.obs = obs;
  WG_LAUNCH(wg::launch_finish(d, fin, (cudaStream_t)cuda_stream), "wg_finish_kernel(mes)");
  return WG_OK;
}

}  // extern "C"
