// Env-layer kernels: MesClass ring buffers + observation extraction, rewards, truncation, reset bookkeeping.
// Reference: WindGym/MesClass.py:23-703 (Mes / turb_mes / farm_mes), Wind_Farm_Env.py:513-520 (_get_obs),
// :804-820 (_action_penalty), :866-918 (rewards), :972-1027 (step tail), :680-732 (reset head).
// One CTA per env; everything here is a few KB per env, i.e. ~1 % of the flow kernel's traffic.
#include "wg_internal.cuh"

namespace wg {

// numpy's pairwise summation (contiguous float64, numpy/core/src/umath/loops_utils.h.src) so that window means
// round exactly like the reference's np.mean on the same samples.
template <class Get>
__device__ __forceinline__ double pairwise_block(Get get, int lo, int n) {  // n <= 128: numpy's unrolled-by-8 leaf
  if (n < 8) {
    double res = 0.0;
    for (int i = 0; i < n; ++i) res += get(lo + i);
    return res;
  }
  double r[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) r[k] = get(lo + k);
  int i = 8;
  for (; i < n - (n % 8); i += 8) {
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] += get(lo + i + k);
  }
  double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
  for (; i < n; ++i) res += get(lo + i);
  return res;
}
template <class Get>
__device__ __noinline__ double pairwise_split(Get get, int lo, int n) {  // n > 128: numpy halves the range
  if (n <= 128) return pairwise_block(get, lo, n);
  int n2 = n / 2;
  n2 -= n2 % 8;
  return pairwise_split(get, lo, n2) + pairwise_split(get, lo + n2, n - n2);
}
template <class Get>
__device__ __forceinline__ double pairwise_sum(Get get, int lo, int n) {
  return n <= 128 ? pairwise_block(get, lo, n) : pairwise_split(get, lo, n);
}

// 2*(val-lo)/(hi-lo)-1 evaluated in float32 exactly like numpy does on a float32 array (MesClass.py:324-326)
__device__ __forceinline__ float scale_f32(float val, float lo, float span) {
  return __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fsub_rn(val, lo)), span), 1.0f);
}

__device__ __forceinline__ void cp_async4(float* dst_shared, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_shared)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void cp_async16(float* dst_shared, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_shared)), "l"(src)
               : "memory");
}

__device__ __forceinline__ unsigned long long splitmix(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ __noinline__ float normal_noise(unsigned long long seed, int b, int push, int chan, int t) {
  unsigned long long k = splitmix(seed ^ splitmix(((unsigned long long)b << 32) ^ (unsigned)push));
  k = splitmix(k ^ (((unsigned long long)chan << 32) | (unsigned)t));
  unsigned long long k2 = splitmix(k);
  double u1 = ((double)(k >> 11) + 1.0) * (1.0 / 9007199254740993.0);
  double u2 = (double)(k2 >> 11) * (1.0 / 9007199254740992.0);
  return (float)(sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2));
}

// window [lo, hi) of rolling value i over a history of length L (MesClass.py:85-116)
__device__ __forceinline__ void window_bounds(int L, int N, int W, int i, int& lo, int& hi) {
  if (i == 0) { lo = max(0, L - W); hi = L; return; }
  if (i == N - 1 && L >= W) { lo = 0; hi = W; return; }
  if (L < W) { lo = 0; hi = L; return; }
  int spacing = max(1, (L - W) / (N - 1));
  int pos = min(i * spacing, L - W);
  lo = pos; hi = pos + W;
}

// np.std(u - U) / U over a whole ring (turb_mes.calc_TI, MesClass.py:220-237), float64 like the reference
template <class Get>
__device__ __noinline__ float calc_ti(Get get, int L) {  // rare observation kind: kept out of the hot path's code
  double U = pairwise_sum(get, 0, L) / L;
  auto dev = [&](int k) { return get(k) - U; };
  double m2 = pairwise_sum(dev, 0, L) / L;
  auto sq = [&](int k) { double e = (get(k) - U) - m2; return e * e; };
  double var = pairwise_sum(sq, 0, L) / L;
  return (float)(sqrt(var) / U);
}

// One WARP per env (4 envs per CTA): every phase is a lane-strided loop, phases are separated by __syncwarp.
// STAGE = true: the env's measurement rings and power deques are staged in shared memory first (one coalesced
// read instead of dependent global loads inside the serial window sums); pushes go to both copies.
#define WG_FIN_WARPS 4
// LEAN = true: the common configuration -- no measurement noise, no TI observations, no Power_diff reward -- compiled
// without those paths (a third of the code: this 20 us single-wave kernel pays for every instruction line it has
// to fetch cold); the host picks the variant from the handle's configuration.
// PAIR = true (batches of at most one wave of envs): TWO warps per env.  The kernel is a serial chain of ~1500
// dependent warp instructions per env; the measurement chain (ring pushes, observation windows) and the power chain
// (power deques, reward, truncation) only share their inputs, so they run side by side on the two warps -- same
// arithmetic, same bits, ~40 % less latency.  The warps of an env meet at a named barrier (id 1 + env slot).
template <bool PAIR>
__device__ __forceinline__ void env_barrier(int slot) {
  if (PAIR) asm volatile("bar.sync %0, 64;" ::"r"(1 + slot) : "memory");
  else __syncwarp();
}
template <bool STAGE, bool LEAN, bool PAIR>
__global__ void __launch_bounds__(WG_FIN_WARPS * 32 * (PAIR ? 2 : 1), PAIR ? 4 : 8) wg_finish_kernel(const Dev d, const FinishArgs a) {
  const int warp_cta = threadIdx.x >> 5, T = d.T;
  const int warp = PAIR ? warp_cta >> 1 : warp_cta;        // env slot of the CTA
  const int role = PAIR ? warp_cta & 1 : 0;                // PAIR: 0 = measurement chain, 1 = power chain
  const int lane = threadIdx.x & 31;
  const int lane_e = PAIR ? (threadIdx.x & 63) : lane, n_e = PAIR ? 64 : 32;   // thread index / threads of the env
  const bool do_mes = !PAIR || role == 0, do_pow = !PAIR || role == 1;
  const int b = d.b0 + blockIdx.x * WG_FIN_WARPS + warp;
  if (b >= d.b0 + d.Bg) return;
  if (a.mask && !a.mask[b]) return;
  extern __shared__ __align__(16) float s_dyn[];
  __shared__ float s_vals[WG_FIN_WARPS][4][WG_MAX_T];
  __shared__ float s_noisy[PAIR ? WG_FIN_WARPS : 1][4][PAIR ? WG_MAX_T : 1];
  float (*s_val)[WG_MAX_T] = s_vals[warp];
  float* g_rings = d.rings + (size_t)b * d.ring_floats;
  float* g_fp = d.fp_ring + (size_t)b * d.power_avg;
  float* g_bp = d.bp_ring + (size_t)b * d.power_avg;
  const int per_env = (d.ring_floats + 2 * d.power_avg + 3) & ~3;  // shared-memory stride of an env: 16-byte aligned
  float* rings = STAGE ? s_dyn + (size_t)warp * per_env : g_rings;
  float* fp = STAGE ? rings + d.ring_floats : g_fp;
  float* bp = STAGE ? fp + d.power_avg : g_bp;
  if (STAGE) {  // asynchronous copies (LDGSTS): every load of the env is in flight before the first wait
    const bool v16 = ((d.ring_floats | per_env) & 3) == 0;  // rows 16-byte aligned in global and shared memory
    if (v16) {
      for (int i = lane_e * 4; i < d.ring_floats; i += 4 * n_e) cp_async16(rings + i, g_rings + i);
    } else {
      for (int i = lane_e; i < d.ring_floats; i += n_e) cp_async4(rings + i, g_rings + i);
    }
    for (int i = lane_e; i < d.power_avg; i += n_e) { cp_async4(fp + i, g_fp + i); cp_async4(bp + i, g_bp + i); }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  // scalars of the env, loaded up front (used by the pushes and the reward at the end)
  int np = d.n_push[b];
  int nfp_tot = d.n_fp[b], nbp_tot = d.n_bp[b];
  const float g_rated = d.rated[b];
  const int g_ts = d.timestep[b], g_tmax = d.time_max[b];
  // Programmatic dependent launch behind the step's flow kernel: everything above (ring staging, env scalars -- none
  // of it written by the flow kernel) ran while the flow grid was still draining; its results (substep means, baseline
  // power, yaws) are read from here on.  Without the launch attribute the wait returns at once.
  if (a.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  // The next step's flow kernel may start now (FlowArgs::pdl_wait): the flow grid of THIS step is complete (waited for
  // above, or ordinary stream order), and the next one touches nothing this kernel reads or writes before its own
  // griddepcontrol.wait.
  if (a.trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const float g_base_pow = d.base_pow_mean[b];
  if (a.flags & (FIN_PUSH_MES | FIN_PUSH_FP))
    for (int t = lane_e; t < T; t += n_e) {
      const float* src[4] = {a.in_ws, a.in_wd, a.in_yaw, a.in_power};
#pragma unroll
      for (int c = 0; c < 4; ++c)
        s_val[c][t] = (a.flags & FIN_MEAS_FROM_ARGS) ? src[c][b * T + t] : d.meas[(b * 4 + c) * T + t];
    }
  if (STAGE) asm volatile("cp.async.wait_group 0;" ::: "memory");
  env_barrier<PAIR>(warp);

  if (do_pow && (a.flags & FIN_PUSH_FP)) {  // farm_pow_deq.append(mean_power.sum()) (Wind_Farm_Env.py:975-977): noise-free means
    if (lane == 0) {
      auto gp = [&](int k) { return (double)s_val[3][k]; };
      const double vd = pairwise_sum(gp, 0, T);
      const float v = (float)vd;
      fp[nfp_tot % d.power_avg] = v;
      if (STAGE) g_fp[nfp_tot % d.power_avg] = v;
      if (!LEAN && d.power_reward == 3)  // Power_diff: keep what float32 drops
        d.fp_ring_lo[(size_t)b * d.power_avg + nfp_tot % d.power_avg] = (float)(vd - (double)v);
      d.n_fp[b] = nfp_tot + 1;
    }
    nfp_tot += 1;
  }
  if (do_pow && (a.flags & FIN_PUSH_BP)) {  // base_pow_deq.append(mean(baseline farm sums)) (:978-979)
    if (lane == 0) {
      const float v = g_base_pow;
      bp[nbp_tot % d.power_avg] = v;
      if (STAGE) g_bp[nbp_tot % d.power_avg] = v;
      d.n_bp[b] = nbp_tot + 1;
    }
    nbp_tot += 1;
  }
  __syncwarp();  // one warp per env: s_val is overwritten with the noisy values below
  float (*s_push)[WG_MAX_T] = s_val;   // PAIR: the measurement warp keeps its noisy copy apart from the power warp's input
  if (PAIR) s_push = reinterpret_cast<float (*)[WG_MAX_T]>(&s_noisy[warp][0][0]);

  if (do_mes && (a.flags & FIN_PUSH_MES)) {
    for (int t = lane; t < T; t += 32) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float v = s_val[c][t];
        if (!LEAN && d.noise && d.noise_std[c] > 0.f) v += d.noise_std[c] * normal_noise(d.noise_seed, b, np, c, t);
        s_push[c][t] = v;
        const int o = d.ch_base[c] + t * d.ch_H[c] + np % d.ch_H[c];
        rings[o] = v;  // deque.append (MesClass.py:66-68, :580-586)
        if (STAGE) g_rings[o] = v;
      }
    }
    __syncwarp();
    if (lane < 3) {  // farm-level rings: mean ws, mean wd, sum power (MesClass.py:589-591)
      const int c = lane == 2 ? 3 : lane;
      auto g = [&](int k) { return (double)s_push[c][k]; };
      const double sum = pairwise_sum(g, 0, T);
      const float v = (float)(lane == 2 ? sum : sum / T);
      const int o = d.farm_off[lane] + np % d.ch_H[c];
      rings[o] = v;
      if (STAGE) g_rings[o] = v;
    }
    if (lane == 0) d.n_push[b] = np + 1;
    np += 1;
  }
  __syncwarp();

  if (do_mes && (a.flags & FIN_OBS)) {  // farm_mes.get_measurements(scaled=True) + clip (MesClass.py:679-703, Wind_Farm_Env.py:513-520)
    const int n_out = d.obs_rows * d.obs_dim;
    for (int o = lane; o < n_out; o += 32) {
      const ObsDesc ds = d.obs_desc[o];
      float val = 0.f;
      if (np > 0) {
        if (!LEAN && ds.kind == 3) {  // farm TI = mean of the (individually scaled) turbine TIs (MesClass.py:670-673)
          const int H = d.ch_H[0], L = min(np, H);
          double acc = 0.0;
          for (int t = 0; t < T; ++t) {
            const float* rg = rings + d.ch_base[0] + t * d.ch_H[0];
            const int st = (np - L) % H;  // oldest sample of the deque; k < L <= H: one conditional wrap
            auto get = [&](int k) { int i = st + k; if (i >= H) i -= H; return (double)rg[i]; };
            acc += (double)scale_f32(calc_ti(get, L), d.ti_lo, d.ti_span);
          }
          val = (float)(acc / T);
        } else {
          const int H = ds.H, L = min(np, H);
          const float* rg = rings + ds.off;
          const int st = (np - L) % H;
          auto get = [&](int k) { int i = st + k; if (i >= H) i -= H; return (double)rg[i]; };
          float raw;
          if (ds.kind == 0) {
            raw = rg[(np - 1) % H];
          } else if (LEAN || ds.kind == 1) {
            int lo, hi;
            window_bounds(L, ds.N, ds.W, ds.win, lo, hi);
            raw = (float)(pairwise_sum(get, lo, hi - lo) / (hi - lo));
          } else {
            raw = calc_ti(get, L);
          }
          val = scale_f32(raw, ds.lo, ds.span);
        }
        val = fminf(fmaxf(val, -1.0f), 1.0f);
      }
      a.obs[(size_t)b * n_out + o] = val;
      if (a.obs_h) a.obs_h[(size_t)b * n_out + o] = val;
    }
  }

  if (do_pow && (a.flags & FIN_REWARD) && lane == 0) {
    const int PA = d.power_avg;
    const int nfp = min(nfp_tot, PA), nbp = min(nbp_tot, PA);
    bool nan_seen = false;
    double sfp = 0.0, sbp = 0.0;
    for (int k = 0; k < nfp; ++k) { sfp += fp[k]; nan_seen |= isnan(fp[k]); }
    for (int k = 0; k < nbp; ++k) sbp += bp[k];
    if (nan_seen) atomicOr(&d.flags[b], 1);  // raise Exception("NaN Power") (Wind_Farm_Env.py:980-981)
    double rew = 0.0;
    if (d.power_reward == 1) {
      rew = (sfp / nfp) / (sbp / nbp) - 1.0;
    } else if (d.power_reward == 2) {
      rew = (sfp / nfp) / T / (double)g_rated;
    } else if (!LEAN && d.power_reward == 3) {  // Power_diff over the logical (oldest -> newest) order of the deque
      const int ws_ = PA / 10, ntot = nfp_tot;
      const float* fplo = d.fp_ring_lo + (size_t)b * PA;
      auto lg = [&](int k) { const int i = (ntot - nfp + k) % PA; return (double)fp[i] + (double)fplo[i]; };
      double latest = 0.0, oldest = 0.0;
      int nl = 0, no = 0;
      for (int k = PA - ws_; k < PA && k < nfp; ++k) { latest += lg(k); ++nl; }
      for (int k = 0; k < ws_ && k < nfp; ++k) { oldest += lg(k); ++no; }
      rew = (latest / nl - oldest / no) / T;
    }
    rew *= (double)d.power_scaling;
    if (d.action_penalty >= 0.001f) {
      double pen = 0.0;
      const float* yaw = d.yaw + (size_t)b * d.F * T;  // farm 0
      for (int t = 0; t < T; ++t)
        pen += (d.pen_type == 0) ? fabs((double)d.old_yaw[b * T + t] - (double)yaw[t]) : fabs((double)yaw[t]);
      pen /= T;
      if (d.pen_type == 1) pen /= (double)d.yaw_max;
      rew -= (double)d.action_penalty * pen;
    }
    a.reward[b] = (float)rew;
    const int ts = g_ts;
    const uint8_t tr = ts >= g_tmax ? 1 : 0;  // Wind_Farm_Env.py:1003, :1027
    a.truncated[b] = tr;
    if (a.reward_h) { a.reward_h[b] = (float)rew; a.truncated_h[b] = tr; }
    d.timestep[b] = ts + 1;
  }
  if (a.done_flag) {
    // Host-visible completion.  Every warp orders its host stores at device scope (warp barrier, then lane 0's
    // __threadfence: cheap) and arrives on a device counter; the LAST warp to arrive issues the one system-scope fence
    // -- cumulative over everything it has observed through the counter, i.e. all warps' stores -- and publishes the
    // step's sequence number.  (A system fence per warp was the top stall of this kernel in the zero-copy path: ncu
    // profiles/r05n_finish512_ncu_summary.txt, membar 13.5 stall cycles per issued instruction.)
    __syncwarp();
    if (lane == 0) {
      __threadfence();
      const unsigned t = atomicAdd(a.done_count, 1u);
      if (t == (unsigned)d.Bg * (PAIR ? 2u : 1u) - 1u) {
        *a.done_count = 0u;
        __threadfence_system();
        *a.done_flag = a.seq;
      }
    }
  }
}

// WindFarmEnv.reset head (Wind_Farm_Env.py:689-732): wind conditions, rotated layout, clean flow + measurement state
__global__ void wg_reset_init_kernel(const Dev d, const ResetDevArgs a) {
  const int b = d.b0 + blockIdx.x, tid = threadIdx.x, T = d.T;
  if (a.mask && !a.mask[b]) return;
  const double wdv = (double)a.wd[b];
  const double th = (270.0 - wdv) * 0.017453292519943295;
  double mx = 0.0, my = 0.0;
  for (int t = 0; t < T; ++t) { mx += d.x_pos[t]; my += d.y_pos[t]; }
  mx /= T; my /= T;
  const double c = cos(th), s = sin(th);
  if (tid < T) {
    const double dx = d.x_pos[tid] - mx, dy = d.y_pos[tid] - my;
    d.xr[b * T + tid] = (float)(dx * c + dy * s);
    d.yr[b * T + tid] = (float)(-dx * s + dy * c);
    const float ws = a.ws[b];
    const float yaw0 = a.yaw0[b * T + tid];
    for (int f = 0; f < d.F; ++f) {
      const int i = (b * d.F + f) * T + tid;
      d.head[i] = 0; d.count[i] = 0; d.retire[i] = 0;
      d.yaw[i] = yaw0; d.u[i] = ws; d.v[i] = 0.f; d.w[i] = 0.f; d.power[i] = 0.f; d.ct[i] = 0.f; d.derate[i] = 1.f;
    }
    d.old_yaw[b * T + tid] = yaw0;
    for (int cc = 0; cc < 4; ++cc) d.meas[(b * 4 + cc) * T + tid] = 0.f;
  }
  __syncthreads();
  if (tid < T) {  // rotor planes sorted by x (ties by index): the flow kernel's bracket search runs on this order
    const float x = d.xr[b * T + tid];
    int rank = 0;
    for (int t = 0; t < T; ++t) {
      const float xt = d.xr[b * T + t];
      rank += (xt < x || (xt == x && t < tid)) ? 1 : 0;
    }
    d.xs_sorted[b * T + rank] = x;
    d.ord_sorted[b * T + rank] = tid;
  }
  if (tid == 0) {
    double xm = -1e300;
    for (int t = 0; t < T; ++t) xm = fmax(xm, (d.x_pos[t] - mx) * c + (d.y_pos[t] - my) * s);
    d.xmax[b] = (float)xm;
    d.ws[b] = a.ws[b]; d.ti[b] = a.ti[b]; d.wd[b] = a.wd[b]; d.rated[b] = a.rated[b];
    d.knu1[b] = a.ti[b] > 0.f ? K1 * powf(a.ti[b], 0.3f) : 0.f;
    d.k_emit[b] = a.k_emit[b]; d.time_max[b] = a.time_max[b]; d.spin[b] = a.t_dev[b];
    d.timestep[b] = 0; d.n_push[b] = 0; d.flags[b] = 0; d.base_pow_mean[b] = 0.f;
    for (int k = 0; k < 3; ++k) d.tb_off[b * 3 + k] = a.tb_off ? a.tb_off[b * 3 + k] : 0.f;
    d.tb_scale[b] = a.tb_scale ? a.tb_scale[b] : 0.f;
    for (int f = 0; f < d.F; ++f) d.n_step[b * d.F + f] = 0;
  }
}

template <bool LEAN, bool PAIR>
static cudaError_t launch_finish_as(const Dev& d, const FinishArgs& a, cudaStream_t s, size_t smem, int dev) {
  static size_t configured_dev[WG_MAX_DEVICES] = {};  // per device: the opt-in is a per-device function attribute
  size_t& configured = configured_dev[dev];
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(wg_finish_kernel<true, LEAN, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  const int grid = (d.Bg + WG_FIN_WARPS - 1) / WG_FIN_WARPS, block = WG_FIN_WARPS * 32 * (PAIR ? 2 : 1);
  if (a.pdl) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, wg_finish_kernel<true, LEAN, PAIR>, d, a);
  }
  wg_finish_kernel<true, LEAN, PAIR><<<grid, block, smem, s>>>(d, a);
  return cudaGetLastError();
}

// envs of one wave of the two-warps-per-env variant (4 CTAs of 4 envs per SM on 148 SMs): below this the kernel is pure
// latency and the variant pays; above it the one-warp variant keeps the whole batch in a single wave
#ifndef WG_FIN_PAIR_MAX
#define WG_FIN_PAIR_MAX 2048
#endif

cudaError_t launch_finish(const Dev& d, const FinishArgs& a, cudaStream_t s) {
  const size_t smem = sizeof(float) * WG_FIN_WARPS * (((size_t)d.ring_floats + 2 * (size_t)d.power_avg + 3) & ~(size_t)3);
  if (smem <= 100 * 1024) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= WG_MAX_DEVICES) dev = 0;
    const bool pair = d.Bg <= WG_FIN_PAIR_MAX && smem <= 48 * 1024;
    if (pair) return d.fin_lean ? launch_finish_as<true, true>(d, a, s, smem, dev) : launch_finish_as<false, true>(d, a, s, smem, dev);
    return d.fin_lean ? launch_finish_as<true, false>(d, a, s, smem, dev) : launch_finish_as<false, false>(d, a, s, smem, dev);
  }
  const int grid = (d.Bg + WG_FIN_WARPS - 1) / WG_FIN_WARPS;
  wg_finish_kernel<false, false, false><<<grid, WG_FIN_WARPS * 32, 0, s>>>(d, a);
  return cudaGetLastError();
}

cudaError_t launch_reset_init(const Dev& d, const ResetDevArgs& a, cudaStream_t s) {
  wg_reset_init_kernel<<<d.Bg, 64, 0, s>>>(d, a);
  return cudaGetLastError();
}

// Copy every per-env field of slot src[k] to slot dst[k] (swap-in of a pre-developed spare env on auto-reset).
__global__ void __launch_bounds__(256) wg_copy_envs_kernel(unsigned char* __restrict__ state,
                                                           const CopyField* __restrict__ fields, const int* __restrict__ src,
                                                           const int* __restrict__ dst, int n) {
  const CopyField f = fields[blockIdx.y];
  const int k = blockIdx.z;
  if (k >= n) return;
  const size_t so = (size_t)src[k] * f.per_env, dof = (size_t)dst[k] * f.per_env;
  for (unsigned r = 0; r < f.n_rep; ++r) {
    unsigned char* base = state + f.offset + (size_t)r * f.rep_stride;
    if ((f.per_env & 15u) == 0) {
      const uint4* s4 = reinterpret_cast<const uint4*>(base + so);
      uint4* d4 = reinterpret_cast<uint4*>(base + dof);
      for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < f.per_env / 16; i += gridDim.x * blockDim.x) d4[i] = s4[i];
    } else {
      const unsigned* s1 = reinterpret_cast<const unsigned*>(base + so);
      unsigned* d1 = reinterpret_cast<unsigned*>(base + dof);
      for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < f.per_env / 4; i += gridDim.x * blockDim.x) d1[i] = s1[i];
    }
  }
}

// Launch order of wg_step: the active envs sorted by descending work (live stations over the env's farms), so that
// the CTAs left over when the grid drains are the short ones (4096 CTAs on 888 resident slots: a greedy
// longest-first schedule ends within half a short CTA of the ideal, a random one within a long CTA).
// One CTA, counting sort over WG_ORDER_BINS load classes; the order inside a class is arbitrary (envs are
// independent: the launch order never changes a result).
#define WG_ORDER_BINS 1024
__global__ void __launch_bounds__(1024) wg_order_kernel(const int* __restrict__ load, int F, int Bg, int quantum,
                                                        int* __restrict__ order) {
  __shared__ int hist[WG_ORDER_BINS];
  __shared__ int wsum[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  hist[tid] = 0;
  __syncthreads();
  for (int b = tid; b < Bg; b += blockDim.x) {
    int n = 0;
    for (int f = 0; f < F; ++f) n += max(load[b * F + f], 0);
    atomicAdd(&hist[WG_ORDER_BINS - 1 - min(n / quantum, WG_ORDER_BINS - 1)], 1);  // bin 0 = heaviest class
  }
  __syncthreads();
  // exclusive scan of the 1024 bins: warp scan, then the 32 warp totals
  const int v = hist[tid];
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int w = wsum[lane];
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    wsum[lane] = winc - w;
  }
  __syncthreads();
  hist[tid] = wsum[warp] + inc - v;
  __syncthreads();
  for (int b = tid; b < Bg; b += blockDim.x) {
    int n = 0;
    for (int f = 0; f < F; ++f) n += max(load[b * F + f], 0);
    order[atomicAdd(&hist[WG_ORDER_BINS - 1 - min(n / quantum, WG_ORDER_BINS - 1)], 1)] = b;
  }
}

cudaError_t launch_order(const Dev& d, cudaStream_t s) {
  const int most = d.F * d.T * d.P;  // stations an env can hold
  wg_order_kernel<<<1, 1024, 0, s>>>(d.load, d.F, d.Bg, (most + WG_ORDER_BINS - 1) / WG_ORDER_BINS, d.order);
  return cudaGetLastError();
}

// Work table of a single-step wg_step launch: which CTA streams which part of which farm (FlowArgs::work).
//  * fewer farms than resident CTA slots (a GPU's share of a sharded batch, e.g. 512 envs): the farms are cut into
//    parts of at most q tiles, q the smallest value whose part count fits the slots -- one wave that fills the
//    machine, every CTA about equally long, instead of one CTA per farm with the launch as long as the heaviest farm;
//  * between one and two waves of farms (e.g. 1024 envs on 888 slots): parts of at most q tiles that make up TWO full
//    waves, instead of one full wave and a nearly empty one;
//  * more farms: one CTA per farm, heaviest first; optionally the lightest `tail_units` farms, launched last, in
//    `tail_parts` parts each, so that the grid drains on short CTAs;
//  * large farms (64-turbine variant: 100-250 tiles per farm): no part longer than `max_tiles` tiles in any regime --
//    a CTA of ~45 tiles runs ~100 us, the per-CTA overhead is negligible and many short CTAs pack and drain better.
// Any split gives bit-identical results (fixed-point rotor sums, flow.cu).  One CTA; counting sort by tiles per part.
#define WG_PLAN_BINS 1024
__device__ __forceinline__ int plan_block_sum(int v, int* red) {  // all 1024 threads
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  int t = red[lane];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}
__device__ __forceinline__ void plan_exclusive_scan(int* hist, int* wsum) {  // hist[1024] in place, all 1024 threads
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int v = hist[tid];
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int w = wsum[lane];
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    wsum[lane] = winc - w;
  }
  __syncthreads();
  hist[tid] = wsum[warp] + inc - v;
  __syncthreads();
}
__global__ void __launch_bounds__(1024) wg_plan_kernel(const int* __restrict__ load, int U, PlanArgs p,
                                                       int2* __restrict__ work, int* __restrict__ flags) {
  __shared__ int hist[WG_PLAN_BINS];
  __shared__ int wsum[32];
  const int tid = threadIdx.x;
  auto tiles = [&](int u) { return min(max((max(load[u], 0) + WG_TILE - 1) / WG_TILE, 1), WG_PLAN_BINS - 1); };
  int q = WG_PLAN_BINS;  // tiles per part; >= every farm's tiles: no split
  if (p.target > 0) {
    int lo = 1, hi = WG_PLAN_BINS - 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      int n = 0;
      for (int u = tid; u < U; u += blockDim.x) n += min((tiles(u) + mid - 1) / mid, WG_MAX_PARTS);
      n = plan_block_sum(n, wsum);
      if (n <= p.target) hi = mid; else lo = mid + 1;
    }
    q = lo;
  }
  // rank of every farm by weight (bin 0 = heaviest): the tail rule needs it
  hist[tid] = 0;
  __syncthreads();
  for (int u = tid; u < U; u += blockDim.x) atomicAdd(&hist[WG_PLAN_BINS - 1 - tiles(u)], 1);
  __syncthreads();
  plan_exclusive_scan(hist, wsum);
  auto parts_of = [&](int u, int t) {
    int n = 1;
    if (p.target > 0) n = (t + q - 1) / q;
    else if (p.tail_parts > 1 && hist[WG_PLAN_BINS - 1 - t] >= U - p.tail_units) n = p.tail_parts;
    if (p.max_tiles > 0) n = max(n, (t + p.max_tiles - 1) / p.max_tiles);
    return min(n, WG_MAX_PARTS);
  };
  int np_mine[8], t_mine[8];  // up to 8192 farms per handle go through registers; more are recomputed
  for (int k = 0, u = tid; u < U && k < 8; ++k, u += blockDim.x) { t_mine[k] = tiles(u); np_mine[k] = parts_of(u, t_mine[k]); }
  __syncthreads();
  hist[tid] = 0;
  __syncthreads();
  for (int k = 0, u = tid; u < U; ++k, u += blockDim.x) {
    const int t = k < 8 ? t_mine[k] : tiles(u);
    const int n = k < 8 ? np_mine[k] : 1;
    atomicAdd(&hist[WG_PLAN_BINS - 1 - (t + n - 1) / n], n);
  }
  __syncthreads();
  plan_exclusive_scan(hist, wsum);
  for (int k = 0, u = tid; u < U; ++k, u += blockDim.x) {
    const int t = k < 8 ? t_mine[k] : tiles(u);
    const int n = k < 8 ? np_mine[k] : 1;
    const int at = atomicAdd(&hist[WG_PLAN_BINS - 1 - (t + n - 1) / n], n);
    for (int i = 0; i < n; ++i)
      if (at + i < p.n_work) work[at + i] = make_int2(u, i | (n << 8));
  }
  __syncthreads();
  // after the scatter the last bin's counter is the number of entries written
  const int total = hist[WG_PLAN_BINS - 1];
  for (int i = total + tid; i < p.n_work; i += blockDim.x) work[i] = make_int2(-1, 0);
  if (tid == 0 && total > p.n_work) atomicOr(flags, 4);  // cannot happen by construction (api.cu sizes the table); loud if it does
}

cudaError_t launch_plan(const Dev& d, const PlanArgs& p, cudaStream_t s) {
  wg_plan_kernel<<<1, 1024, 0, s>>>(d.load, d.Bg * d.F, p, d.work, d.flags);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ device-side pool
// Auto-reset without the host in the loop (reference: the env is reset by its caller right after a truncated step,
// Wind_Farm_Env.py:1003-1025; reset = t_developed + fill flow steps, :722-766).  The envs [n_active, B) of the
// allocation are spares.  Every step, on the stepping stream: wg_pool_swap_kernel pairs finished episodes with READY
// spares and wg_pool_copy_kernel copies the spares' state over them.  Every few steps, on a background stream:
// wg_pool_claim_kernel takes the spares that were consumed, draws new wind conditions for them (counter-based RNG),
// the masked reset spins them up, wg_pool_publish_kernel marks them READY.  A spare travels
//   NEED -> (claim) REFILLING -> (publish) READY -> (swap) PENDING -> (next step's swap, i.e. after the copy) NEED.
__device__ __forceinline__ double pool_uniform(unsigned long long seed, int slot, int gen, int k) {
  unsigned long long x = splitmix(seed ^ splitmix(((unsigned long long)(unsigned)slot << 32) | (unsigned)gen));
  x = splitmix(x + 0x632BE59BD9B4E019ull * (unsigned long long)(k + 1));
  return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}

__global__ void __launch_bounds__(128) wg_pool_claim_kernel(const Dev d, const PoolDev p, const PoolDraw w, int mask_row) {
  const int b = p.n_active + blockIdx.x, tid = threadIdx.x, T = d.T;
  if (b >= p.B) return;
  uint8_t* mask = p.masks + (size_t)mask_row * p.B;
  __shared__ int s_take;
  if (tid == 0) {
    s_take = atomicCAS(&p.status[b], POOL_NEED, POOL_REFILLING) == POOL_NEED ? 1 : 0;
    mask[b] = (uint8_t)s_take;
  }
  __syncthreads();
  if (!s_take) return;
  const int gen = p.gen[b];
  // draw order of the reference: ws, ti, wd, then the yaw offsets (Wind_Farm_Env.py:564-568, :715)
  const double ws = w.ws_min + (w.ws_max - w.ws_min) * pool_uniform(w.seed, b, gen, 0);
  const double ti = w.ti_min + (w.ti_max - w.ti_min) * pool_uniform(w.seed, b, gen, 1);
  const double wd = w.wd_min + (w.wd_max - w.wd_min) * pool_uniform(w.seed, b, gen, 2);
  for (int t = tid; t < T; t += blockDim.x)
    p.yaw0[b * T + t] = w.yaw_random ? (float)(-w.yaw_start + 2.0 * w.yaw_start * pool_uniform(w.seed, b, gen, 8 + t))
                                     : w.yaw_const;
  if (tid == 0) {
    // integer decisions of the reset in fp64, as the reference makes them (:723-732; EnvConfig.reset_integers)
    const double th = (270.0 - wd) * 0.017453292519943295, c = cos(th), sn = sin(th);
    double mx = 0.0, my = 0.0;
    for (int t = 0; t < T; ++t) { mx += d.x_pos[t]; my += d.y_pos[t]; }
    mx /= T; my /= T;
    double lo = 1e300, hi = -1e300;
    for (int t = 0; t < T; ++t) {
      const double xr = (d.x_pos[t] - mx) * c + (d.y_pos[t] - my) * sn;
      lo = fmin(lo, xr); hi = fmax(hi, xr);
    }
    const double t_inflow = (hi - lo) / ws;
    const long long t_dev = (long long)(t_inflow * 2.0);
    p.t_dev[b] = (int)llrint((double)t_dev / (double)d.dt);
    p.time_max[b] = w.eval_mode ? 9999999 : (int)(t_inflow * w.n_passthrough);
    p.k_emit[b] = max(1, (int)ceil((double)d.d_particle * (double)d.D / (ws * (double)d.dt) - 1e-9));
    p.ws[b] = (float)ws; p.ti[b] = (float)ti; p.wd[b] = (float)wd;
    // turbine.power(ws): np.interp on the power table (Wind_Farm_Env.py:700)
    const int n = d.n_tab;
    double pw;
    if (ws <= (double)d.tab_ws[0]) pw = d.tab_p[0];
    else if (ws >= (double)d.tab_ws[n - 1]) pw = d.tab_p[n - 1];
    else {
      int k = 0;
      while (k + 2 < n && (double)d.tab_ws[k + 1] <= ws) ++k;
      const double x0 = d.tab_ws[k], x1 = d.tab_ws[k + 1];
      pw = ((double)d.tab_p[k + 1] - (double)d.tab_p[k]) / (x1 - x0) * (ws - x0) + (double)d.tab_p[k];
    }
    p.rated[b] = (float)pw;
    if (w.tb_inv_std > 0.0) {  // an env position inside the shared turbulence box, scale_TI factor TI U / std(u_box)
      for (int k = 0; k < 3; ++k) p.tb_off[b * 3 + k] = (float)(w.tb_len[k] * pool_uniform(w.seed, b, gen, 4 + k));
      p.tb_scale[b] = (float)(ti * ws * w.tb_inv_std);
      p.ti_flow[b] = (float)ti;
    } else {
      for (int k = 0; k < 3; ++k) p.tb_off[b * 3 + k] = 0.f;
      p.tb_scale[b] = 0.f;
      p.ti_flow[b] = 0.f;  // turbtype "None": RandomTurbulence(ti = 0) (:661-665)
    }
    p.gen[b] = gen + 1;
  }
}

__global__ void wg_pool_publish_kernel(const PoolDev p, int mask_row) {
  const int b = p.n_active + blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  if (p.masks[(size_t)mask_row * p.B + b]) {
    __threadfence();
    atomicCAS(&p.status[b], POOL_REFILLING, POOL_READY);
    atomicAdd(&p.stats[2], 1ull);
  }
}

// One small CTA (256 threads, a few thousand registers: it fits next to whatever is resident -- a 1024-thread CTA
// needs a whole SM's register file and waited tens of microseconds for one to drain while background spin-ups hold a
// CTA on every SM).  (1) spares whose copy was issued in the previous step are free to be refilled; (2) ready spares
// and finished episodes are collected IN INDEX ORDER (block scan, no atomics: the k-th ready spare goes to the k-th
// finished env, run after run) and paired; (3) swapped[b] = 1 for the envs that get a new episode in this step -- an
// episode that finds no ready spare simply runs on and is paired in a later step.  Every thread looks at 16
// consecutive envs per pass (4-byte loads of the flag bytes), so 4096 envs are one pass.
#define WG_SWAP_THREADS 256
__device__ __forceinline__ int swap_block_excl_scan(int v, int* wsum, int& total) {  // all WG_SWAP_THREADS threads
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  int before = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < WG_SWAP_THREADS / 32; ++w) {
    const int c = wsum[w];
    if (w < warp) before += c;
    tot += c;
  }
  total = tot;
  return before + inc - v;
}
__global__ void __launch_bounds__(WG_SWAP_THREADS) wg_pool_swap_kernel(const Dev d, const PoolDev p,
                                                                       const uint8_t* __restrict__ truncated,
                                                                       uint8_t* __restrict__ swapped) {
  __shared__ int wsum[WG_SWAP_THREADS / 32];
  __shared__ int s_src[WG_POOL_MAX_SWAP], s_dst[WG_POOL_MAX_SWAP];
  // Launched as programmatic dependent of the step's finish kernel (launch_pool_swap): only the launch latency
  // overlaps -- the kernel waits for the finish grid to be complete before its first read, and releases the copy
  // kernel's launch at once (that one waits for this grid in the same way).
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int tid = threadIdx.x;
  int nsrc = 0, ndst = 0, nneed = 0;
  // spares: 4 slots per thread and pass
  for (int base = p.n_active; base < p.B; base += 4 * WG_SWAP_THREADS) {
    int mine[4], cnt = 0, need = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int b = base + 4 * tid + k;
      mine[k] = -1;
      if (b < p.B) {
        int st = p.status[b];
        if (st == POOL_PENDING) { p.status[b] = POOL_NEED; st = POOL_NEED; }
        if (st == POOL_READY) { mine[k] = b; ++cnt; }
        need += st == POOL_NEED;
      }
    }
    int total, tneed;
    int at = nsrc + swap_block_excl_scan(cnt, wsum, total);
    swap_block_excl_scan(need, wsum, tneed);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (mine[k] >= 0) { if (at < WG_POOL_MAX_SWAP) s_src[at] = mine[k]; ++at; }
    nsrc += total;
    nneed += tneed;
  }
  // finished episodes: 16 envs per thread and pass, flags read four at a time
  const bool al4 = (((uintptr_t)truncated | (uintptr_t)swapped) & 3) == 0;
  for (int base = 0; base < p.n_active; base += 16 * WG_SWAP_THREADS) {
    const int b0 = base + 16 * tid;
    unsigned bits = 0;  // bit k: env b0 + k finished
    if (al4 && b0 + 16 <= p.n_active) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const unsigned w = reinterpret_cast<const unsigned*>(truncated + b0)[q];
#pragma unroll
        for (int k = 0; k < 4; ++k) bits |= ((w >> (8 * k)) & 0xffu) ? 1u << (4 * q + k) : 0u;
        reinterpret_cast<unsigned*>(swapped + b0)[q] = 0u;
      }
    } else {
      for (int k = 0; k < 16; ++k)
        if (b0 + k < p.n_active) { bits |= truncated[b0 + k] ? 1u << k : 0u; swapped[b0 + k] = 0; }
    }
    int total;
    int at = ndst + swap_block_excl_scan(__popc(bits), wsum, total);
    while (bits) {
      const int k = __ffs(bits) - 1;
      bits &= bits - 1;
      if (at < WG_POOL_MAX_SWAP) s_dst[at] = b0 + k;
      ++at;
    }
    ndst += total;
  }
  __syncthreads();
  const int n = min(min(nsrc, ndst), WG_POOL_MAX_SWAP);
  if (tid < n) {
    const int src = s_src[tid], dst = s_dst[tid];
    p.status[src] = POOL_PENDING;
    p.swap[tid] = src;
    p.swap[WG_POOL_MAX_SWAP + tid] = dst;
    swapped[dst] = 1;
  }
  if (tid == 0) {
    p.swap[2 * WG_POOL_MAX_SWAP] = n;
    if (n) atomicAdd(&p.stats[0], (unsigned long long)n);
    if (ndst > n) atomicAdd(&p.stats[1], (unsigned long long)(ndst - n));
    p.stats[3] = (unsigned long long)nneed;
    if (p.need_host) *p.need_host = nneed;
  }
}

// Copy the paired spares over the finished envs: grid (chunks, state fields + 1, pair lanes) -- one CTA per (chunk of
// a field, lane of pairs), so that the ~45 fields of an env move side by side (a CTA that walks the field list pays
// a dependent global load per field: measured 40 us per step); no pairs: every CTA reads n and leaves.  Field index
// n_fields is the observation row: the finished episode's last observation is kept in final_obs (the "terminal
// observation" of the vector-env APIs), the row then becomes the spare's reset observation.
__global__ void __launch_bounds__(256) wg_pool_copy_kernel(unsigned char* __restrict__ state, const CopyField* __restrict__ fields,
                                                           int n_fields, const PoolDev p, float* __restrict__ obs,
                                                           float* __restrict__ final_obs, int obs_floats) {
  asm volatile("griddepcontrol.wait;" ::: "memory");               // the swap kernel's list (see there)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // next: the flow kernel, waiting at its top
  const int n = p.swap[2 * WG_POOL_MAX_SWAP];
  if (n == 0) return;
  const unsigned nchunk = gridDim.x, chunk = blockIdx.x;
  const int fi = blockIdx.y;
  if (fi == n_fields) {
    if (chunk) return;
    for (int k = blockIdx.z; k < n; k += gridDim.z) {
      const int src = p.swap[k], dst = p.swap[WG_POOL_MAX_SWAP + k];
      for (int i = threadIdx.x; i < obs_floats; i += blockDim.x) {
        if (final_obs) final_obs[(size_t)dst * obs_floats + i] = obs[(size_t)dst * obs_floats + i];
        obs[(size_t)dst * obs_floats + i] = obs[(size_t)src * obs_floats + i];
      }
    }
    return;
  }
  const CopyField f = fields[fi];
  if ((size_t)chunk * blockDim.x * 16 >= f.per_env && chunk) return;  // small field: chunk 0 does it all
  for (int k = blockIdx.z; k < n; k += gridDim.z) {
    const int src = p.swap[k], dst = p.swap[WG_POOL_MAX_SWAP + k];
    const size_t so = (size_t)src * f.per_env, dof = (size_t)dst * f.per_env;
    for (unsigned r = 0; r < f.n_rep; ++r) {
      unsigned char* base = state + f.offset + (size_t)r * f.rep_stride;
      if ((f.per_env & 15u) == 0) {
        const uint4* s4 = reinterpret_cast<const uint4*>(base + so);
        uint4* d4 = reinterpret_cast<uint4*>(base + dof);
        for (unsigned i = chunk * blockDim.x + threadIdx.x; i < f.per_env / 16; i += nchunk * blockDim.x) d4[i] = s4[i];
      } else {
        const unsigned* s1 = reinterpret_cast<const unsigned*>(base + so);
        unsigned* d1 = reinterpret_cast<unsigned*>(base + dof);
        for (unsigned i = chunk * blockDim.x + threadIdx.x; i < f.per_env / 4; i += nchunk * blockDim.x) d1[i] = s1[i];
      }
    }
  }
}

cudaError_t launch_pool_claim(const Dev& d, const PoolDev& p, const PoolDraw& w, int mask_row, cudaStream_t s) {
  wg_pool_claim_kernel<<<p.B - p.n_active, 128, 0, s>>>(d, p, w, mask_row);
  return cudaGetLastError();
}
cudaError_t launch_pool_publish(const PoolDev& p, int mask_row, cudaStream_t s) {
  wg_pool_publish_kernel<<<(p.B - p.n_active + 255) / 256, 256, 0, s>>>(p, mask_row);
  return cudaGetLastError();
}
// launch with programmatic stream serialization: the grid may be scheduled while its predecessor in the stream is
// still running; the kernels wait (griddepcontrol.wait) before they touch memory
template <class... KArgs, class... Args>
static cudaError_t launch_programmatic(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}
cudaError_t launch_pool_swap(const Dev& d, const PoolDev& p, const uint8_t* truncated, uint8_t* swapped, cudaStream_t s) {
  return launch_programmatic(wg_pool_swap_kernel, dim3(1), dim3(WG_SWAP_THREADS), s, d, p, truncated, swapped);
}
cudaError_t launch_pool_copy(unsigned char* state, const CopyField* fields, int n_fields, const PoolDev& p, float* obs,
                             float* final_obs, int obs_floats, cudaStream_t s) {
  return launch_programmatic(wg_pool_copy_kernel, dim3(16, n_fields + 1, 4), dim3(256), s, state, fields, n_fields, p, obs,
                             final_obs, obs_floats);
}

cudaError_t launch_copy_envs(unsigned char* state, const CopyField* fields, int n_fields, const int* src, const int* dst,
                             int n, cudaStream_t s) {
  for (int k0 = 0; k0 < n; k0 += 65535) {
    const int nk = n - k0 < 65535 ? n - k0 : 65535;
    wg_copy_envs_kernel<<<dim3(24, n_fields, nk), 256, 0, s>>>(state, fields, src + k0, dst + k0, nk);
  }
  return cudaGetLastError();
}

}  // namespace wg
