// wg_flow_kernel -- the DWM flow step (dynamiks seam: DWMFlowSimulation.step, reference call sites
// WindGym/Wind_Farm_Env.py:734,:745,:945,:953) for thousands of independent farms in one launch.
//
// One CTA owns one (env, farm).  Its live wake stations (all turbine chains, ring-addressed) form one flat
// list that the CTA's warps stream in 32-station tiles:
//     cp.async.bulk (TMA 1-D bulk copy, mbarrier complete_tx)  HBM -> shared
//     thread-per-station implicit Ainslie march: ONE fused forward sweep (continuity-consistent radial
//       velocity + tridiagonal rows + Thomas elimination; c' in registers, d' written in place into the
//       shared row) and one back substitution that also accumulates the shear-layer integrals of the NEW
//       profile for the next step's eddy viscosity
//     rotor-plane bracket detection -> per-warp hit list -> (hit x quadrature point) mapped onto full warps,
//       16-lane shuffle reduction for the rotor average                               (superposition gather)
//     cp.async.bulk shared -> HBM
// then the per-turbine epilogue (P/CT tables, particle release) runs in the same CTA, and the substep loop
// (dt_env/dt_sim, or a whole spin-up) repeats without leaving the kernel.
// Algorithmic traffic per station and step: 256 B profile + 16 B mutable + 16 B emission scalars read,
// 256 B + 16 B written = 560 B.  HBM/issue bound; no tensor cores (stencil + gather).
//
// Profile row layout (64 floats, 16-byte chunks XOR-swizzled with slot & 7): nodes 0..62 hold U(r_j); node 63 is
// the Dirichlet node (U = 1 always), so its slot carries bw = sqrt(2 M (1 - Umin)) of the row instead -- the
// shear-layer term of the eddy viscosity (oracle/dwm_numpy.py:113-115), computed when the row was last written.
#include <math_constants.h>

#include <cstdlib>

#include "wg_internal.cuh"

namespace wg {

__constant__ float c_qy[WG_NQ];
__constant__ float c_qz[WG_NQ];

void set_rotor_points(const float* qy, const float* qz) {
  cudaMemcpyToSymbol(c_qy, qy, sizeof(float) * WG_NQ);
  cudaMemcpyToSymbol(c_qz, qz, sizeof(float) * WG_NQ);
}

#define WG_NWARP 4        // warps per CTA
#define WG_HIT_CAP 96     // per-warp hit list entries (flushed when fewer than 64 free)

struct __align__(16) FlowShared {
  unsigned long long mbar[WG_NWARP][2];
  float xr[WG_MAX_T], yr[WG_MAX_T], yaw[WG_MAX_T], u[WG_MAX_T], v[WG_MAX_T], w[WG_MAX_T], pw[WG_MAX_T],
      ct[WG_MAX_T], ind[WG_MAX_T], bw0[WG_MAX_T];
  float xs[WG_MAX_T];                       // turbine x sorted ascending
  float sum_ws[WG_MAX_T], sum_wd[WG_MAX_T], sum_yaw[WG_MAX_T], sum_pw[WG_MAX_T];
  int ord[WG_MAX_T];                        // turbine index of xs[k]
  int head[WG_MAX_T], count[WG_MAX_T], pre[WG_MAX_T + 1], emit_slot[WG_MAX_T];
  float base_sum;
  int pad[3];
  float4 hit_a[WG_NWARP][WG_HIT_CAP];       // w*U0e*cos g0, w*U0e*sin g0, ry, rz
  int2 hit_b[WG_NWARP][WG_HIT_CAP];         // (row | key << 8), accumulator index j*T + chain
};

static size_t acc_bytes(int T) { return ((size_t)2 * T * T * sizeof(float) + 127) / 128 * 128; }
static size_t hdr_bytes() { return (sizeof(FlowShared) + 127) / 128 * 128; }

size_t flow_smem_bytes(int T, int n_stage) {
  return hdr_bytes() + acc_bytes(T) + (size_t)WG_NWARP * n_stage * WG_TILE * WG_ROW_BYTES;
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
__device__ __forceinline__ float sqrt_fast(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
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

__device__ __forceinline__ float4 ld_chunk(const float* row, int key, int c) {
  return *reinterpret_cast<const float4*>(row + ((c ^ key) << 2));
}
__device__ __forceinline__ void st_chunk(float* row, int key, int c, float4 v) {
  *reinterpret_cast<float4*>(row + ((c ^ key) << 2)) = v;
}

// Implicit Ainslie march of one profile row held in shared memory (oracle/dwm_numpy.py:ainslie_march).
// Forward sweep: per node j the continuity-consistent radial velocity (pass A of the oracle) feeds the
// tridiagonal row and its Thomas elimination (pass B) immediately; the eddy viscosity needs the row's shear
// integral bw, which rides in slot 63.  c' stays in registers, d' replaces U_j in the shared row.
// Back substitution (pass C) writes the new profile and accumulates its bw.  Returns the new centre value.
__device__ __forceinline__ float march_row(float* __restrict__ row, int key, float dxt, float xt, float knu1) {
  constexpr float IDR2 = 1.f / (DR * DR);
  constexpr float HDR = 0.5f * DR;
  constexpr float I2DR = 0.5f / DR;
  float cp[WG_NR - 1];
  float4 cur = ld_chunk(row, key, 0);
  const float bw = ld_chunk(row, key, WG_NR / 4 - 1).w;
  const float nu = knu1 * f1_filter(xt) + K2 * f2_filter(xt) * bw;
  const float idx = 1.f / fmaxf(dxt, DXT_MIN);
  const float nu2 = 2.f * nu * IDR2, nu8 = nu * I2DR;
  float I = 0.f, rgp = 0.f, cpm, dpm;
  float um = 0.f;  // U_{j-1}
#pragma unroll
  for (int c = 0; c < WG_NR / 4; ++c) {
    float4 nxt = cur;
    if (c + 1 < WG_NR / 4) nxt = ld_chunk(row, key, c + 1);
    if (c + 1 == WG_NR / 4 - 1) nxt.w = 1.f;  // Dirichlet node: the slot holds bw, the value is 1
    const float uu[5] = {cur.x, cur.y, cur.z, cur.w, nxt.x};
    float dout[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = 4 * c + e;
      const float uj = uu[e], up1 = uu[e + 1];
      if (j == 0) {
        const float m = rcp_fast(fmaf(uj, idx, 2.f * nu2));
        cpm = -2.f * nu2 * m;
        dpm = uj * uj * idx * m;
        cp[0] = cpm;
        dout[e] = dpm;
      } else if (j < WG_NR - 1) {
        const float r = j * DR, rinv = 1.f / r;
        const float am = (1.f - HDR * rinv) * IDR2, ap = (1.f + HDR * rinv) * IDR2;
        const float up = (up1 - um) * I2DR;
        const float upr = up * rinv;
        const float lap = fmaf(fmaf(-2.f, uj, up1 + um), IDR2, upr);
        const float Ip = fmaf(HDR, rgp, I);
        const float den = fmaf(-HDR, up, uj);
        const float g = fmaf(upr, Ip, lap) * rcp_fast(den);
        const float rg = r * g;
        I = fmaf(HDR, rg, Ip);
        rgp = rg;
        const float Vd = -nu8 * rinv * I;        // nu * Vh_j / (2 dr)
        const float a = fmaf(nu, am, Vd);        // -(sub-diagonal)
        const float cc = fmaf(-nu, ap, Vd);      // super-diagonal
        const float bb = fmaf(uj, idx, nu2);
        float dd = uj * uj * idx;
        if (j == WG_NR - 2) dd -= cc;            // Dirichlet U_63 = 1
        const float m = rcp_fast(fmaf(a, cpm, bb));
        cpm = cc * m;
        dpm = fmaf(a, dpm, dd) * m;
        cp[j] = cpm;
        dout[e] = dpm;
      } else {
        dout[e] = 0.f;
      }
      um = uj;
    }
    if (c < WG_NR / 4 - 1) st_chunk(row, key, c, make_float4(dout[0], dout[1], dout[2], dout[3]));
    else cur = make_float4(dout[0], dout[1], dout[2], 0.f);  // last chunk stays in registers: d'_60..62
    if (c < WG_NR / 4 - 1) cur = nxt;
  }
  // ---- back substitution + shear integrals of the new profile
  float un = 1.f, M = 0.f, umin = 1.f;
#pragma unroll
  for (int c = WG_NR / 4 - 1; c >= 0; --c) {
    float4 dq = (c == WG_NR / 4 - 1) ? cur : ld_chunk(row, key, c);
    float dv[4] = {dq.x, dq.y, dq.z, dq.w};
    float o[4];
#pragma unroll
    for (int e = 3; e >= 0; --e) {
      const int j = 4 * c + e;
      if (j == WG_NR - 1) { o[e] = 0.f; continue; }
      un = (j == WG_NR - 2) ? dv[e] : fmaf(-cp[j], un, dv[e]);
      o[e] = un;
      M = fmaf(j * DR, 1.f - un, M);
      umin = fminf(umin, un);
    }
    if (c < WG_NR / 4 - 1) st_chunk(row, key, c, make_float4(o[0], o[1], o[2], o[3]));
    else cur = make_float4(o[0], o[1], o[2], 0.f);
  }
  cur.w = sqrtf(fmaxf(2.f * (M * DR) * (1.f - umin), 0.f));
  st_chunk(row, key, WG_NR / 4 - 1, cur);
  return un;
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

// Evaluate the queued (station row, rotor) hits of one warp: two hits per pass, 16 quadrature points each on
// 16 lanes, shuffle-reduced to the rotor average, accumulated per (rotor, emitting chain).
__device__ __forceinline__ void flush_hits(const float* __restrict__ tile, const float4* __restrict__ ha,
                                           const int2* __restrict__ hb, int nh, float* acc_du, float* acc_dv,
                                           int lane, float qy, float qz) {
  const unsigned full = 0xffffffffu;
  const int half = lane >> 4;
  for (int h0 = 0; h0 < nh; h0 += 2) {
    const int h = h0 + half;
    const bool ok = h < nh;
    const float4 a = ha[ok ? h : h0];
    const int2 b = hb[ok ? h : h0];
    const float* row = tile + (b.x & 0xff) * WG_NR;
    const int key = b.x >> 8;
    const float dy = a.z + qy, dz = a.w + qz;
    const float s = sqrt_fast(fmaf(dy, dy, dz * dz)) * (1.f / DR);
    const int j0 = min((int)s, WG_NR - 2);
    const float fr = s - (float)j0;
    const int j1 = j0 + 1;
    const float u0 = row[(((j0 >> 2) ^ key) << 2) | (j0 & 3)];
    float u1 = row[(((j1 >> 2) ^ key) << 2) | (j1 & 3)];
    if (j1 == WG_NR - 1) u1 = 1.f;
    float d = fmaf(fr, u0 - u1, 1.f - u0);  // (1-u0)(1-fr) + (1-u1) fr
    if (s >= (float)(WG_NR - 1) || !ok) d = 0.f;
    d += __shfl_xor_sync(full, d, 8);
    d += __shfl_xor_sync(full, d, 4);
    d += __shfl_xor_sync(full, d, 2);
    d += __shfl_xor_sync(full, d, 1);
    if ((lane & 15) == 0 && ok) {
      d *= (1.f / WG_NQ);
      atomicAdd(&acc_du[b.y], a.x * d);
      atomicAdd(&acc_dv[b.y], a.y * d);
    }
  }
}

template <int NSTAGE>
__global__ void __launch_bounds__(WG_NWARP * 32, NSTAGE == 1 ? 4 : 3) wg_flow_kernel(const Dev d, const FlowArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  FlowShared& sh = *reinterpret_cast<FlowShared*>(smem_raw);
  const int T = d.T, P = d.P, F = d.F;
  float* acc_du = reinterpret_cast<float*>(smem_raw + ((sizeof(FlowShared) + 127) / 128) * 128);
  float* acc_dv = acc_du + T * T;
  float* bufs = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(acc_du) +
                                         (((size_t)2 * T * T * sizeof(float) + 127) / 128) * 128);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / F, f = blockIdx.x % F;
  const int bf = b * F + f;
  if (a.mask && !a.mask[b]) return;
  if (!((a.farm_mask >> f) & 1)) return;
  int nsteps = (a.mode == FLOW_FIXED) ? a.n_fixed : (a.mode == FLOW_SPIN ? d.spin[b] : d.S);
  if (nsteps <= 0) return;

  const float ws = d.ws[b], wd = d.wd[b], dt = d.dt, R = d.R, xmax = d.xmax[b];
  const float rR = 1.f / R;
  const float ti = d.ti[b];
  const float knu1_env = ti > 0.f ? K1 * powf(ti, 0.3f) : 0.f;
  const int k_emit = max(d.k_emit[b], 1);
  float* __restrict__ prof = d.prof + (size_t)bf * T * P * WG_NR;
  float* __restrict__ pcon = d.pcon + (size_t)bf * T * P * 4;
  float* pmut0 = d.pmut + (size_t)bf * T * P * 4;
  float* pmut1 = pmut0 + (size_t)d.B * F * T * P * 4;
  float* my_buf = bufs + (size_t)warp * NSTAGE * WG_TILE * WG_NR;
  const float qy = c_qy[lane & 15], qz = c_qz[lane & 15];

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
  if (tid < T) {  // rank sort of the rotor-plane x positions (ties broken by index): xs ascending, ord = turbine
    const float x = sh.xr[tid];
    int rank = 0;
    for (int t = 0; t < T; ++t) {
      const float xt = sh.xr[t];
      rank += (xt < x || (xt == x && t < tid)) ? 1 : 0;
    }
    sh.xs[rank] = x;
    sh.ord[rank] = tid;
  }
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
    pmc = pcc = pmn = pcn = make_float4(0.f, 0.f, 0.f, 0.f);
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
    for (int tile = warp; tile < ntiles; tile += WG_NWARP) {
      const int nt = tile + WG_NWARP;
      Ln.valid = 0;
      if (nt < ntiles) Ln = locate(sh, nt, lane, T, P, ntot);
      if (NSTAGE == 2) {
        bulk_wait_read0();  // stores that used the other stage have drained their shared-memory reads
        __syncwarp();
        if (nt < ntiles) issue_load(Ln, stage ^ 1);
      }
      if (nt < ntiles) load_scalars(Ln, pmn, pcn);
      mbar_wait(&sh.mbar[warp][stage], (phase >> stage) & 1u);
      phase ^= (1u << stage);

      float* tile_base = my_buf + (size_t)stage * WG_TILE * WG_NR;
      float* row = tile_base + lane * WG_NR;
      const int key = Lc.slot & 7;
      float xn = 0.f, yn = 0.f, zn = 0.f, dx = 0.f;
      if (Lc.valid) {
        moved(pmc, pcc, ws, dt, xn, yn, zn, dx);
        float xt_mid = (pmc.x + 0.5f * dx - sh.xr[Lc.chain]) * rR;
        float ucn = march_row(row, key, dx * rR, xt_mid, pcc.y);
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
        const bool has_o = Lc.valid && Lc.q > 0;
        const bool has_y = Lc.valid && Lc.q < sh.count[Lc.chain] - 1;
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
        // x-range touched by this tile (warp-uniform): only rotor planes inside it can be bracketed
        float lo = Lc.valid ? xn : CUDART_INF_F, hi = Lc.valid ? xn : -CUDART_INF_F;
        if (has_o) { lo = fminf(lo, xo); hi = fmaxf(hi, xo); }
        if (has_y) { lo = fminf(lo, xy); hi = fmaxf(hi, xy); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          lo = fminf(lo, __shfl_xor_sync(full, lo, o));
          hi = fmaxf(hi, __shfl_xor_sync(full, hi, o));
        }
        const float u0cg = pcc.x * pcc.z, u0sg = pcc.x * pcc.w;
        float4* ha = sh.hit_a[warp];
        int2* hb = sh.hit_b[warp];
        const unsigned lt = (1u << lane) - 1u;
        const int rowkey = lane | (key << 8);
        int nh = 0;
        int k = 0;
        while (k < T && sh.xs[k] < lo) ++k;
        for (; k < T; ++k) {
          const float xj = sh.xs[k];
          if (xj >= hi) break;
          const int j = sh.ord[k];
          const bool mine = Lc.valid && j != Lc.chain;
          // interval A: (self = younger end, older neighbour); interval B: (younger neighbour, self = older end)
          float sgA = 0.f, sgB = 0.f;
          if (mine && has_o) sgA = (xn <= xj && xj < xo) ? 1.f : ((xo <= xj && xj < xn) ? -1.f : 0.f);
          if (mine && has_y) sgB = (xy <= xj && xj < xn) ? 1.f : ((xn <= xj && xj < xy) ? -1.f : 0.f);
          const unsigned mA = __ballot_sync(full, sgA != 0.f), mB = __ballot_sync(full, sgB != 0.f);
          if ((mA | mB) == 0u) continue;
          const float yrj = sh.yr[j];
          if (sgA != 0.f) {
            const float w = (xj - xn) / (xo - xn);
            const float wg = sgA * (1.f - w);
            const float yc = yn * (1.f - w) + yo * w, zc = zn * (1.f - w) + zo * w;
            const int p = nh + __popc(mA & lt);
            ha[p] = make_float4(wg * u0cg, wg * u0sg, (yrj - yc) * rR, (d.zh - zc) * rR);
            hb[p] = make_int2(rowkey, j * T + Lc.chain);
          }
          nh += __popc(mA);
          if (sgB != 0.f) {
            const float w = (xj - xy) / (xn - xy);
            const float wg = sgB * w;
            const float yc = yy * (1.f - w) + yn * w, zc = zy * (1.f - w) + zn * w;
            const int p = nh + __popc(mB & lt);
            ha[p] = make_float4(wg * u0cg, wg * u0sg, (yrj - yc) * rR, (d.zh - zc) * rR);
            hb[p] = make_int2(rowkey, j * T + Lc.chain);
          }
          nh += __popc(mB);
          if (nh > WG_HIT_CAP - 64) {
            __syncwarp();
            flush_hits(tile_base, ha, hb, nh, acc_du, acc_dv, lane, qy, qz);
            __syncwarp();
            nh = 0;
          }
        }
        __syncwarp();
        flush_hits(tile_base, ha, hb, nh, acc_du, acc_dv, lane, qy, qz);
      }
      __syncwarp();
      if (NSTAGE == 1) {
        bulk_wait_read0();  // the store has drained its shared-memory reads: the buffer can be refilled
        __syncwarp();
        if (nt < ntiles) issue_load(Ln, 0);
      } else {
        stage ^= 1;
      }
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
        // cell-averaged top-hat inlet (same formula as the row writer below): centre value and shear integral
        const float fw = 1.f - 0.45f * ind * ind;
        const float rw2 = fw * fw * (1.f - ind) / (1.f - 2.f * ind);
        float M = 0.f, u0v = 1.f;
        for (int j = 0; j < WG_NR - 1; ++j) {
          const float rlo = fmaxf((float)j - 0.5f, 0.f) * DR, rhi = ((float)j + 0.5f) * DR;
          const float frac = fminf(fmaxf((rw2 - rlo * rlo) / (rhi * rhi - rlo * rlo), 0.f), 1.f);
          const float df = 2.f * ind * frac;
          if (j == 0) u0v = 1.f - df;
          M = fmaf((float)j * DR, df, M);
        }
        // the inlet is monotone non-decreasing in r, so Umin is the centre value
        sh.bw0[tid] = sqrtf(fmaxf(2.f * (M * DR) * (1.f - u0v), 0.f));
        *reinterpret_cast<float4*>(pm_new + ((size_t)tid * P + slot) * 4) =
            make_float4(sh.xr[tid], sh.yr[tid], d.zh, u0v);
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
          vals[e] = (j == WG_NR - 1) ? sh.bw0[t] : 1.f - 2.f * ind * frac;
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

// Tile buffers per warp: 1 (default; 4 CTAs/SM, latency hidden by the other warps) or 2 (WG_FLOW_STAGES=2; the next
// tile is prefetched while the current one is marched, 3 CTAs/SM).
static int flow_stages() {
  static int n = 0;
  if (n == 0) {
    const char* e = getenv("WG_FLOW_STAGES");
    n = (e && e[0] == '2') ? 2 : 1;
  }
  return n;
}

cudaError_t launch_flow(const Dev& d, const FlowArgs& a, cudaStream_t s) {
  const int ns = flow_stages();
  const size_t smem = flow_smem_bytes(d.T, ns);
  static size_t configured[3] = {0, 0, 0};
  if (smem > configured[ns]) {
    cudaError_t e = ns == 1
        ? cudaFuncSetAttribute(wg_flow_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
        : cudaFuncSetAttribute(wg_flow_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured[ns] = smem;
  }
  if (ns == 1) wg_flow_kernel<1><<<d.B * d.F, WG_NWARP * 32, smem, s>>>(d, a);
  else wg_flow_kernel<2><<<d.B * d.F, WG_NWARP * 32, smem, s>>>(d, a);
  return cudaGetLastError();
}

}  // namespace wg
