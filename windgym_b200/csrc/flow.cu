// wg_flow_kernel -- the DWM flow step (dynamiks seam: DWMFlowSimulation.step, reference call sites
// WindGym/Wind_Farm_Env.py:734,:745,:945,:953) for thousands of independent farms in one launch.
//
// One CTA owns one (env, farm).  Its live wake stations (all turbine chains, ring-addressed) form one flat
// list that the CTA's warps stream in 32-station tiles:
//     cp.async.bulk (TMA 1-D bulk copy, mbarrier complete_tx)  HBM -> shared,   double buffered per warp
//     thread-per-station implicit Ainslie march (registers)                       r-stencil, Thomas solve
//     rotor-plane bracket detection + deficit sampling of the freshly marched rows  (superposition gather)
//     cp.async.bulk shared -> HBM
// then the per-turbine epilogue (rotor average, P/CT tables, particle release) runs in the same CTA, and the
// substep loop (dt_env/dt_sim, or a whole spin-up) repeats without leaving the kernel.
// Algorithmic traffic per station and step: 256 B profile + 16 B mutable + 16 B emission scalars read,
// 256 B + 16 B written = 560 B.  HBM/issue bound; no tensor cores (stencil + gather).
#include <math_constants.h>

#include "wg_internal.cuh"

namespace wg {

__constant__ float c_qy[WG_NQ];
__constant__ float c_qz[WG_NQ];

void set_rotor_points(const float* qy, const float* qz) {
  cudaMemcpyToSymbol(c_qy, qy, sizeof(float) * WG_NQ);
  cudaMemcpyToSymbol(c_qz, qz, sizeof(float) * WG_NQ);
}

struct __align__(16) FlowShared {
  unsigned long long mbar[8][2];
  float xr[WG_MAX_T], yr[WG_MAX_T], yaw[WG_MAX_T], u[WG_MAX_T], v[WG_MAX_T], w[WG_MAX_T], pw[WG_MAX_T],
      ct[WG_MAX_T], ind[WG_MAX_T];
  float sum_ws[WG_MAX_T], sum_wd[WG_MAX_T], sum_yaw[WG_MAX_T], sum_pw[WG_MAX_T];
  int head[WG_MAX_T], count[WG_MAX_T], pre[WG_MAX_T + 1], emit_slot[WG_MAX_T];
  float base_sum;
  int pad[3];
};

size_t flow_smem_bytes(int T, int n_warps) {
  size_t acc = ((size_t)2 * T * T * sizeof(float) + 127) / 128 * 128;
  return sizeof(FlowShared) + 128 + acc + (size_t)n_warps * 2 * WG_TILE * WG_ROW_BYTES;
}

// ---------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(void* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// TMA 1-D bulk copy shared -> global (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ float rcp_fast(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// ---------------------------------------------------------------------------------------------- physics
__device__ __forceinline__ float f1_filter(float xt) {
  if (xt >= 8.f) return 1.f;
  float q = fmaxf(xt, 0.f) * 0.125f;
  float s = q * sqrtf(q);
  return s - sinf(6.283185307179586f * s) * 0.15915494309189535f;
}
__device__ __forceinline__ float f2_filter(float xt) {
  float lin = 0.025f * xt - 0.0375f;
  if (xt < 4.f) return 0.0625f;
  if (xt < 12.f) return lin;
  if (xt < 20.f) {
    float e = xt - 12.f;
    return 0.00105f * e * e * e + lin;
  }
  return 1.f;
}

// np.interp semantics (clamped ends)
__device__ __forceinline__ float tab_interp(const float* __restrict__ xs, const float* __restrict__ ys, int n, float x) {
  if (x <= xs[0]) return ys[0];
  if (x >= xs[n - 1]) return ys[n - 1];
  int k = 0;
  while (k < n - 2 && x >= xs[k + 1]) ++k;
  return (ys[k + 1] - ys[k]) / (xs[k + 1] - xs[k]) * (x - xs[k]) + ys[k];
}

// new wake-centre position after one step (Hill-vortex self-induced velocity added to the ambient)
__device__ __forceinline__ void moved(const float4 pm, const float4 pc, float ws, float dt, float& xn, float& yn,
                                      float& zn, float& dx) {
  float kd = K_HILL * (1.f - pm.w) * pc.x;
  float vx = ws - kd * pc.z;
  float vy = kd * pc.w;
  dx = vx * dt;
  xn = pm.x + dx;
  yn = pm.y + vy * dt;
  zn = pm.z;
}

// Implicit Ainslie march of one profile row held in shared memory (chunk-swizzled with `key`).
// Registers: u[64] (profile -> rhs -> solution) and vh[64] (radial velocity per unit nu -> Thomas c').
__device__ __forceinline__ float march_row(float* __restrict__ row, int key, float dxt, float xt, float knu1) {
  float u[WG_NR], vh[WG_NR];
#pragma unroll
  for (int c = 0; c < WG_NR / 4; ++c) {
    float4 t = *reinterpret_cast<const float4*>(row + ((c ^ key) << 2));
    u[4 * c + 0] = t.x; u[4 * c + 1] = t.y; u[4 * c + 2] = t.z; u[4 * c + 3] = t.w;
  }
  constexpr float IDR2 = 1.f / (DR * DR);
  constexpr float HDR = 0.5f * DR;
  constexpr float I2DR = 0.5f / DR;
  // ---- pass A: Laplacian, continuity-consistent radial velocity per unit nu, integrals for nu
  float I = 0.f, rgp = 0.f, M = 0.f, umin = u[0];
#pragma unroll
  for (int j = 1; j < WG_NR - 1; ++j) {
    const float r = j * DR, rinv = 1.f / r;
    float up = (u[j + 1] - u[j - 1]) * I2DR;
    float upr = up * rinv;
    float lap = fmaf(fmaf(-2.f, u[j], u[j + 1] + u[j - 1]), IDR2, upr);
    float Ip = fmaf(HDR, rgp, I);
    float den = fmaf(-HDR, up, u[j]);
    float g = fmaf(upr, Ip, lap) * rcp_fast(den);
    float rg = r * g;
    I = fmaf(HDR, rg, Ip);
    vh[j] = -I * rinv;
    rgp = rg;
    M = fmaf(r, 1.f - u[j], M);
    umin = fminf(umin, u[j]);
  }
  umin = fminf(umin, u[WG_NR - 1]);
  const float lap0 = 4.f * (u[1] - u[0]) * IDR2;
  (void)lap0;
  M *= DR;
  const float nu = knu1 * f1_filter(xt) + K2 * f2_filter(xt) * sqrtf(fmaxf(2.f * M * (1.f - umin), 0.f));
  // ---- pass B: tridiagonal rows + Thomas forward sweep (unknowns 0..62, u[63] = 1 Dirichlet)
  const float idx = 1.f / fmaxf(dxt, DXT_MIN);
  const float nu2 = 2.f * nu * IDR2, nu8 = nu * I2DR;
  float cpm, dpm;
  {
    float m = rcp_fast(fmaf(u[0], idx, 2.f * nu2));
    cpm = -2.f * nu2 * m;
    dpm = u[0] * u[0] * idx * m;
    vh[0] = cpm;
    u[0] = dpm;
  }
#pragma unroll
  for (int j = 1; j < WG_NR - 1; ++j) {
    const float rinv = 1.f / (j * DR);
    const float am = (1.f - HDR * rinv) * IDR2, ap = (1.f + HDR * rinv) * IDR2;
    float Vd = nu8 * vh[j];
    float a = -fmaf(nu, am, Vd);
    float c = fmaf(-nu, ap, Vd);
    float bb = fmaf(u[j], idx, nu2);
    float dd = u[j] * u[j] * idx;
    if (j == WG_NR - 2) dd -= c;
    float m = rcp_fast(fmaf(-a, cpm, bb));
    cpm = c * m;
    dpm = fmaf(-a, dpm, dd) * m;
    vh[j] = cpm;
    u[j] = dpm;
  }
  // ---- pass C: back substitution
  u[WG_NR - 1] = 1.f;
#pragma unroll
  for (int j = WG_NR - 3; j >= 0; --j) u[j] = fmaf(-vh[j], u[j + 1], u[j]);
#pragma unroll
  for (int c = 0; c < WG_NR / 4; ++c) {
    float4 t = make_float4(u[4 * c + 0], u[4 * c + 1], u[4 * c + 2], u[4 * c + 3]);
    *reinterpret_cast<float4*>(row + ((c ^ key) << 2)) = t;
  }
  return u[0];
}

// rotor-averaged deficit of one marched row for a rotor whose centre sits (ry, rz) rotor radii off the wake centre
__device__ __forceinline__ float rotor_deficit(const float* __restrict__ row, int key, float ry, float rz) {
  float acc = 0.f;
#pragma unroll 4
  for (int q = 0; q < WG_NQ; ++q) {
    float dy = ry + c_qy[q], dz = rz + c_qz[q];
    float s = sqrtf(dy * dy + dz * dz) * (1.f / DR);
    int j0 = min((int)s, WG_NR - 2);
    float fr = s - (float)j0;
    int j1 = j0 + 1;
    float u0 = row[(((j0 >> 2) ^ key) << 2) | (j0 & 3)];
    float u1 = row[(((j1 >> 2) ^ key) << 2) | (j1 & 3)];
    float d = (1.f - u0) * (1.f - fr) + (1.f - u1) * fr;
    acc += (s >= (float)(WG_NR - 1)) ? 0.f : d;
  }
  return acc * (1.f / WG_NQ);
}

struct LaneLoc {
  int chain, slot, q, valid;
};

__device__ __forceinline__ LaneLoc locate(const FlowShared& sh, int tile, int lane, int T, int P, int ntot) {
  LaneLoc L;
  int fl = tile * WG_TILE + lane;
  L.valid = fl < ntot;
  if (!L.valid) fl = ntot - 1;
  int lo = 0, hi = T;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (sh.pre[mid] <= fl) lo = mid; else hi = mid;
  }
  L.chain = lo;
  L.q = fl - sh.pre[lo];
  int s = sh.head[lo] - sh.count[lo] + L.q;
  L.slot = s < 0 ? s + P : s;
  return L;
}

// segment bookkeeping for the bulk copies of one tile: a segment = run of lanes with consecutive slots of one chain
struct Seg {
  int start, len, nvalid;
};
__device__ __forceinline__ Seg segments(const LaneLoc& L, int lane) {
  const unsigned full = 0xffffffffu;
  int pc = __shfl_up_sync(full, L.chain, 1), ps = __shfl_up_sync(full, L.slot, 1);
  bool start = L.valid && (lane == 0 || pc != L.chain || ps + 1 != L.slot);
  unsigned sm = __ballot_sync(full, start), vm = __ballot_sync(full, L.valid);
  Seg s;
  s.nvalid = __popc(vm);
  s.start = start;
  unsigned higher = (lane == 31) ? 0u : (sm >> (lane + 1)) << (lane + 1);
  int next = higher ? (__ffs(higher) - 1) : s.nvalid;
  s.len = next - lane;
  return s;
}

__global__ void __launch_bounds__(128, 3) wg_flow_kernel(const Dev d, const FlowArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  FlowShared& sh = *reinterpret_cast<FlowShared*>(smem_raw);
  const int T = d.T, P = d.P, F = d.F;
  float* acc_du = reinterpret_cast<float*>(smem_raw + ((sizeof(FlowShared) + 127) / 128) * 128);
  float* acc_dv = acc_du + T * T;
  float* bufs = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(acc_du) +
                                         (((size_t)2 * T * T * sizeof(float) + 127) / 128) * 128);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, NW = blockDim.x >> 5;
  const int b = blockIdx.x / F, f = blockIdx.x % F;
  const int bf = b * F + f;
  if (a.mask && !a.mask[b]) return;
  if (!((a.farm_mask >> f) & 1)) return;
  int nsteps = (a.mode == FLOW_FIXED) ? a.n_fixed : (a.mode == FLOW_SPIN ? d.spin[b] : d.S);
  if (nsteps <= 0) return;

  const float ws = d.ws[b], wd = d.wd[b], dt = d.dt, R = d.R, xmax = d.xmax[b];
  const float ti = d.ti[b];
  const float knu1_env = ti > 0.f ? K1 * powf(ti, 0.3f) : 0.f;
  const int k_emit = max(d.k_emit[b], 1);
  float* __restrict__ prof = d.prof + (size_t)bf * T * P * WG_NR;
  float* __restrict__ pcon = d.pcon + (size_t)bf * T * P * 4;
  float* pmut0 = d.pmut + (size_t)bf * T * P * 4;
  float* pmut1 = pmut0 + (size_t)d.B * F * T * P * 4;
  float* my_buf = bufs + (size_t)warp * 2 * WG_TILE * WG_NR;

  if (tid < T) {
    sh.xr[tid] = d.xr[b * T + tid];
    sh.yr[tid] = d.yr[b * T + tid];
    float yaw = d.yaw[bf * T + tid];
    if (a.mode == FLOW_STEP && f == 0 && a.actions) {  // _adjust_yaws, Wind_Farm_Env.py:822-864
      d.old_yaw[b * T + tid] = yaw;
      float act = a.actions[b * T + tid];
      if (d.action_method == 0) {
        yaw = fminf(fmaxf(yaw + act * d.yaw_step, d.yaw_min), d.yaw_max);
      } else {
        float tgt = (act + 1.0f) / 2.0f * (d.yaw_max - d.yaw_min) + d.yaw_min;
        tgt = fminf(fmaxf(tgt, yaw - d.yaw_step), yaw + d.yaw_step);
        yaw = fminf(fmaxf(tgt, d.yaw_min), d.yaw_max);
      }
    }
    sh.yaw[tid] = yaw;
    sh.u[tid] = d.u[bf * T + tid];
    sh.v[tid] = d.v[bf * T + tid];
    sh.w[tid] = d.w[bf * T + tid];
    sh.pw[tid] = d.power[bf * T + tid];
    sh.ct[tid] = d.ct[bf * T + tid];
    sh.head[tid] = d.head[bf * T + tid];
    sh.count[tid] = d.count[bf * T + tid];
    sh.sum_ws[tid] = sh.sum_wd[tid] = sh.sum_yaw[tid] = sh.sum_pw[tid] = 0.f;
  }
  if (tid == 0) sh.base_sum = 0.f;
  if (lane == 0) {
    mbar_init(&sh.mbar[warp][0], 1);
    mbar_init(&sh.mbar[warp][1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  int nstep = d.n_step[bf];
  uint32_t phase = 0;
  __syncthreads();

  for (int sub = 0; sub < nsteps; ++sub) {
    const float* __restrict__ pm_old = (nstep & 1) ? pmut1 : pmut0;
    float* __restrict__ pm_new = (nstep & 1) ? pmut0 : pmut1;

    if (tid < T) {
      // baseline farm: greedy yaw controller before its flow step (BasicControllers.py:10-73, Wind_Farm_Env.py:949-952)
      if (a.controller_on && f == 1) {
        float yaw = sh.yaw[tid];
        if (d.base_controller == 0) {
          float off = atanf(sh.v[tid] / sh.u[tid]) * 57.29577951308232f - yaw;
          float st = fminf(fabsf(off), d.yaw_step);
          yaw += (off > 0.f ? st : (off < 0.f ? -st : 0.f));
        } else {
          float st = fminf(fabsf(yaw), d.yaw_step);
          yaw -= (yaw > 0.f ? st : (yaw < 0.f ? -st : 0.f));
        }
        sh.yaw[tid] = yaw;
      }
      // retire stations that will be past the farm (+margin) after this step's move
      int cnt = sh.count[tid];
      const int hd = sh.head[tid];
      while (cnt > 0) {
        int s = hd - cnt;
        if (s < 0) s += P;
        float4 pm = __ldcg(reinterpret_cast<const float4*>(pm_old + ((size_t)tid * P + s) * 4));
        float4 pc = __ldcg(reinterpret_cast<const float4*>(pcon + ((size_t)tid * P + s) * 4));
        float xn, yn, zn, dx;
        moved(pm, pc, ws, dt, xn, yn, zn, dx);
        if (xn > xmax + MARGIN_D * d.D) --cnt; else break;
      }
      sh.count[tid] = cnt;
    }
    for (int i = tid; i < 2 * T * T; i += blockDim.x) acc_du[i] = 0.f;
    __syncthreads();
    if (tid == 0) {
      int s = 0;
      for (int t = 0; t < T; ++t) { sh.pre[t] = s; s += sh.count[t]; }
      sh.pre[T] = s;
    }
    __syncthreads();
    const int ntot = sh.pre[T];
    const int ntiles = (ntot + WG_TILE - 1) / WG_TILE;

    // ------------------------------------------------------------------ warp-private tile pipeline
    LaneLoc Lc, Ln;
    float4 pmc, pcc, pmn, pcn;
    int stage = 0;
    Lc.valid = 0; Lc.chain = 0; Lc.slot = 0; Lc.q = 0;
    pmc = pcc = make_float4(0.f, 0.f, 0.f, 0.f);
    auto issue_load = [&](const LaneLoc& L, int stg) {
      Seg sg = segments(L, lane);
      void* bar = &sh.mbar[warp][stg];
      if (lane == 0) mbar_expect_tx(bar, (uint32_t)sg.nvalid * WG_ROW_BYTES);
      __syncwarp();
      if (sg.start)
        bulk_g2s(my_buf + ((size_t)stg * WG_TILE + lane) * WG_NR, prof + ((size_t)L.chain * P + L.slot) * WG_NR,
                 (uint32_t)sg.len * WG_ROW_BYTES, bar);
    };
    auto load_scalars = [&](const LaneLoc& L, float4& pm, float4& pc) {
      if (L.valid) {
        pm = __ldcg(reinterpret_cast<const float4*>(pm_old + ((size_t)L.chain * P + L.slot) * 4));
        pc = __ldcg(reinterpret_cast<const float4*>(pcon + ((size_t)L.chain * P + L.slot) * 4));
      }
    };
    if (warp < ntiles) {
      Lc = locate(sh, warp, lane, T, P, ntot);
      issue_load(Lc, 0);
      load_scalars(Lc, pmc, pcc);
    }
    for (int tile = warp; tile < ntiles; tile += NW) {
      const int nt = tile + NW;
      bulk_wait_read0();  // stores that used the other stage have drained their shared-memory reads
      __syncwarp();
      Ln.valid = 0;
      if (nt < ntiles) {
        Ln = locate(sh, nt, lane, T, P, ntot);
        issue_load(Ln, stage ^ 1);
        load_scalars(Ln, pmn, pcn);
      }
      mbar_wait(&sh.mbar[warp][stage], (phase >> stage) & 1u);
      phase ^= (1u << stage);

      float* row = my_buf + ((size_t)stage * WG_TILE + lane) * WG_NR;
      const int key = Lc.slot & 7;
      float xn = 0.f, yn = 0.f, zn = 0.f, dx = 0.f;
      if (Lc.valid) {
        moved(pmc, pcc, ws, dt, xn, yn, zn, dx);
        float xt_mid = (pmc.x + 0.5f * dx - sh.xr[Lc.chain]) / R;
        float ucn = march_row(row, key, dx / R, xt_mid, pcc.y);
        *reinterpret_cast<float4*>(pm_new + ((size_t)Lc.chain * P + Lc.slot) * 4) = make_float4(xn, yn, zn, ucn);
      }
      fence_async_smem();
      __syncwarp();
      {  // write the marched rows back (same segments as the load)
        Seg sg = segments(Lc, lane);
        if (sg.start) {
          bulk_s2g(prof + ((size_t)Lc.chain * P + Lc.slot) * WG_NR, row, (uint32_t)sg.len * WG_ROW_BYTES);
          bulk_commit();
        }
      }
      // ---- superposition: which rotor planes does this station bracket together with its age neighbours?
      {
        const unsigned full = 0xffffffffu;
        // older neighbour = flat index - 1 (same chain), younger = flat index + 1
        float xo = __shfl_up_sync(full, xn, 1), yo = __shfl_up_sync(full, yn, 1), zo = __shfl_up_sync(full, zn, 1);
        int co = __shfl_up_sync(full, Lc.chain, 1);
        float xy = __shfl_down_sync(full, xn, 1), yy = __shfl_down_sync(full, yn, 1), zy = __shfl_down_sync(full, zn, 1);
        int cy = __shfl_down_sync(full, Lc.chain, 1);
        int vy_ = __shfl_down_sync(full, Lc.valid, 1);
        bool has_o = Lc.valid && Lc.q > 0;
        bool has_y = Lc.valid && Lc.q < sh.count[Lc.chain] - 1;
        if (has_o && (lane == 0 || co != Lc.chain)) {
          int so = Lc.slot == 0 ? P - 1 : Lc.slot - 1;
          float4 pm = __ldcg(reinterpret_cast<const float4*>(pm_old + ((size_t)Lc.chain * P + so) * 4));
          float4 pc = __ldcg(reinterpret_cast<const float4*>(pcon + ((size_t)Lc.chain * P + so) * 4));
          float dxx;
          moved(pm, pc, ws, dt, xo, yo, zo, dxx);
        }
        if (has_y && (lane == 31 || !vy_ || cy != Lc.chain)) {
          int sy = Lc.slot == P - 1 ? 0 : Lc.slot + 1;
          float4 pm = __ldcg(reinterpret_cast<const float4*>(pm_old + ((size_t)Lc.chain * P + sy) * 4));
          float4 pc = __ldcg(reinterpret_cast<const float4*>(pcon + ((size_t)Lc.chain * P + sy) * 4));
          float dxx;
          moved(pm, pc, ws, dt, xy, yy, zy, dxx);
        }
        // x-range touched by this tile (warp-uniform early-out per turbine)
        float lo = Lc.valid ? xn : CUDART_INF_F, hi = Lc.valid ? xn : -CUDART_INF_F;
        if (has_o) { lo = fminf(lo, xo); hi = fmaxf(hi, xo); }
        if (has_y) { lo = fminf(lo, xy); hi = fmaxf(hi, xy); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          lo = fminf(lo, __shfl_xor_sync(full, lo, o));
          hi = fmaxf(hi, __shfl_xor_sync(full, hi, o));
        }
        const float u0cg = pcc.x * pcc.z, u0sg = pcc.x * pcc.w;
        for (int j = 0; j < T; ++j) {
          const float xj = sh.xr[j];
          if (xj < lo || xj >= hi) continue;
          if (!Lc.valid || j == Lc.chain) continue;
          float wsum_du = 0.f;  // signed interpolation weight * rotor-mean deficit, both intervals
          bool any = false;
          float wgt[2], ycs[2], zcs[2];
          int nh = 0;
          if (has_o) {  // interval (self = younger end, older neighbour)
            float sgn = (xn <= xj && xj < xo) ? 1.f : ((xo <= xj && xj < xn) ? -1.f : 0.f);
            if (sgn != 0.f) {
              float w = (xj - xn) / (xo - xn);
              wgt[nh] = sgn * (1.f - w); ycs[nh] = yn * (1.f - w) + yo * w; zcs[nh] = zn * (1.f - w) + zo * w;
              ++nh;
            }
          }
          if (has_y) {  // interval (younger neighbour, self = older end)
            float sgn = (xy <= xj && xj < xn) ? 1.f : ((xn <= xj && xj < xy) ? -1.f : 0.f);
            if (sgn != 0.f) {
              float w = (xj - xy) / (xn - xy);
              wgt[nh] = sgn * w; ycs[nh] = yy * (1.f - w) + yn * w; zcs[nh] = zy * (1.f - w) + zn * w;
              ++nh;
            }
          }
          for (int h = 0; h < nh; ++h) {
            float Dq = rotor_deficit(row, key, (sh.yr[j] - ycs[h]) / R, (d.zh - zcs[h]) / R);
            wsum_du += wgt[h] * Dq;
            any = true;
          }
          if (any) {
            atomicAdd(&acc_du[j * T + Lc.chain], wsum_du * u0cg);
            atomicAdd(&acc_dv[j * T + Lc.chain], wsum_du * u0sg);
          }
        }
      }
      __syncwarp();
      stage ^= 1;
      Lc = Ln; pmc = pmn; pcc = pcn;
    }
    bulk_wait_all0();
    fence_async_all();
    __syncthreads();

    // ------------------------------------------------------------------ turbine epilogue
    const bool emit = (nstep % k_emit) == 0;
    if (tid < T) {
      float du = 0.f, dv = 0.f;
      for (int i = 0; i < T; ++i) { du += acc_du[tid * T + i]; dv += acc_dv[tid * T + i]; }
      const float u = ws - du, v = dv, w = 0.f;
      const float yaw = sh.yaw[tid];
      float sg, cg;
      sincosf(yaw * 0.017453292519943295f, &sg, &cg);
      const float wse = u * cg;
      const float pw = tab_interp(d.tab_ws, d.tab_p, d.n_tab, wse);
      float ct = tab_interp(d.tab_ws, d.tab_ct, d.n_tab, wse) * cg * cg;
      ct = fminf(fmaxf(ct, 0.f), CT_MAX);
      sh.u[tid] = u; sh.v[tid] = v; sh.w[tid] = w; sh.pw[tid] = pw; sh.ct[tid] = ct;
      int slot = -1;
      if (emit) {
        const float ind = 0.5f * (1.f - sqrtf(1.f - ct));
        sh.ind[tid] = ind;
        slot = sh.head[tid];
        if (sh.count[tid] == P) atomicOr(&d.flags[b], 2); else sh.count[tid] += 1;
        sh.head[tid] = (slot + 1 == P) ? 0 : slot + 1;
        // inlet value at the centre node (cell [0, dr/2]) -- same formula as the row writer below
        const float fw = 1.f - 0.45f * ind * ind;
        const float rw2 = fw * fw * (1.f - ind) / (1.f - 2.f * ind);
        const float frac0 = fminf(fmaxf(rw2 / (0.25f * DR * DR), 0.f), 1.f);
        *reinterpret_cast<float4*>(pm_new + ((size_t)tid * P + slot) * 4) =
            make_float4(sh.xr[tid], sh.yr[tid], d.zh, 1.f - 2.f * ind * frac0);
        *reinterpret_cast<float4*>(pcon + ((size_t)tid * P + slot) * 4) = make_float4(u, knu1_env, cg, sg);
      }
      sh.emit_slot[tid] = slot;
      if (a.mode == FLOW_STEP) {
        if (f == 0) {  // _take_measurements, Wind_Farm_Env.py:480-495
          sh.sum_ws[tid] += sqrtf(u * u + v * v + w * w);
          sh.sum_wd[tid] += atanf(v / u) * 57.29577951308232f + wd;
          sh.sum_yaw[tid] += yaw;
          sh.sum_pw[tid] += pw;
        }
      }
    }
    __syncthreads();
    if (a.mode == FLOW_STEP && f == 1 && tid == 0) {
      float s = 0.f;
      for (int t = 0; t < T; ++t) s += sh.pw[t];
      sh.base_sum += s;
    }
    if (emit) {  // release one particle per turbine: cell-averaged top-hat inlet (IEC 61400-1 ed.4 Annex E)
      for (int idx = tid; idx < T * (WG_NR / 4); idx += blockDim.x) {
        const int t = idx >> 4, c = idx & 15;
        const int slot = sh.emit_slot[t];
        if (slot < 0) continue;
        const float ind = sh.ind[t];
        const float fw = 1.f - 0.45f * ind * ind;
        const float rw2 = fw * fw * (1.f - ind) / (1.f - 2.f * ind);
        float vals[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = 4 * c + e;
          const float rlo = fmaxf((float)j - 0.5f, 0.f) * DR, rhi = ((float)j + 0.5f) * DR;
          const float frac = fminf(fmaxf((rw2 - rlo * rlo) / (rhi * rhi - rlo * rlo), 0.f), 1.f);
          vals[e] = (j == WG_NR - 1) ? 1.f : 1.f - 2.f * ind * frac;
        }
        *reinterpret_cast<float4*>(prof + ((size_t)t * P + slot) * WG_NR + ((c ^ (slot & 7)) << 2)) =
            make_float4(vals[0], vals[1], vals[2], vals[3]);
      }
      fence_async_all();
    }
    ++nstep;
    __syncthreads();
  }

  if (tid < T) {
    d.yaw[bf * T + tid] = sh.yaw[tid];
    d.u[bf * T + tid] = sh.u[tid];
    d.v[bf * T + tid] = sh.v[tid];
    d.w[bf * T + tid] = sh.w[tid];
    d.power[bf * T + tid] = sh.pw[tid];
    d.ct[bf * T + tid] = sh.ct[tid];
    d.head[bf * T + tid] = sh.head[tid];
    d.count[bf * T + tid] = sh.count[tid];
    if (a.mode == FLOW_STEP && f == 0) {  // substep means (Wind_Farm_Env.py:965-969)
      const float inv = 1.f / (float)nsteps;
      d.meas[(b * 4 + 0) * T + tid] = sh.sum_ws[tid] * inv;
      d.meas[(b * 4 + 1) * T + tid] = sh.sum_wd[tid] * inv;
      d.meas[(b * 4 + 2) * T + tid] = sh.sum_yaw[tid] * inv;
      d.meas[(b * 4 + 3) * T + tid] = sh.sum_pw[tid] * inv;
    }
  }
  if (tid == 0) {
    d.n_step[bf] = nstep;
    if (a.mode == FLOW_STEP && f == 1) d.base_pow_mean[b] = sh.base_sum / (float)nsteps;
  }
}

cudaError_t launch_flow(const Dev& d, const FlowArgs& a, cudaStream_t s) {
  const int n_warps = 4;
  const size_t smem = flow_smem_bytes(d.T, n_warps);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(wg_flow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  wg_flow_kernel<<<d.B * d.F, n_warps * 32, smem, s>>>(d, a);
  return cudaGetLastError();
}

}  // namespace wg
