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


#define WG_HIT_CAP 64     // per-warp hit list entries (one detection pass adds at most 2 x 32)

// per-node constants of the radial grid r_j = j dr: {1/(2j), j/2}; entry 0 unused (the axis node has its own row)
__constant__ float2 c_node[WG_NR];

void set_rotor_points(const float* qy, const float* qz) {
  cudaMemcpyToSymbol(c_qy, qy, sizeof(float) * WG_NQ);
  cudaMemcpyToSymbol(c_qz, qz, sizeof(float) * WG_NQ);
  float2 nd[WG_NR];
  nd[0] = make_float2(0.f, 0.f);
  for (int j = 1; j < WG_NR; ++j) nd[j] = make_float2(1.f / (2.f * j), 0.5f * j);
  cudaMemcpyToSymbol(c_node, nd, sizeof(nd));
}

// Per-CTA bookkeeping in front of the tile buffers.  NW = warps per CTA, TC = turbine capacity of the tables
// (16 or WG_MAX_T: small farms leave the shared memory to more resident CTAs).
template <int NW, int TC>
struct __align__(16) FlowShared {
  unsigned long long mbar[NW];
  float xr[TC], yr[TC], yaw[TC], u[TC], v[TC], w[TC], pw[TC], ct[TC], ind[TC], cg[TC], sg[TC];
  float xs[2 * TC];                         // turbine x sorted ascending, padded with +inf
  float sum_ws[TC], sum_wd[TC], sum_yaw[TC], sum_pw[TC];
  float acc_du[NW][TC], acc_dv[NW][TC];     // per-warp superposed deficit per rotor
  int ord[TC];                              // turbine index of xs[k]
  int head[TC], count[TC], pre[TC + 1], emit_slot[TC];
  float base_sum;
  uint32_t tmem_base;                       // variant 2: TMEM allocation of the CTA
  int pad[1];
  float4 hit_a[NW][WG_HIT_CAP];             // w*U0e*cos g0, w*U0e*sin g0, ry, rz
  int2 hit_b[NW][WG_HIT_CAP];               // (row | key << 8), rotor index j
};

template <int NW, int TC>
__host__ __device__ constexpr size_t hdr_bytes() { return (sizeof(FlowShared<NW, TC>) + 127) / 128 * 128; }

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
  return s - __sinf(6.283185307179586f * s) * 0.15915494309189535f;  // argument in [0, 2 pi]
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
// Both sweeps are rolled over eight 8-node groups of the row (per-node grid constants come from c_node)
// so that the loop body stays inside the instruction cache; within a group everything is unrolled and c' is
// statically indexed.  With h = 1/(2j), N = nu/dr^2:
//   lap_j dr^2 = U_{j+1} + U_{j-1} - 2 U_j + h (U_{j+1} - U_{j-1}),  Vd_j = nu Vh_j / (2 dr) = -N h I_j,
//   sub-diagonal -a_j = -(N - N h (1 + I_j)), super-diagonal c_j = -N - N h (1 + I_j), diagonal U_j/dx + 2N.
// The sweep carries W = 1 + I: lap_j dr^2 + h dU (I + rgh) = (su - 2 U_j) + h dU (W + rgh).
// Node j of the forward sweep (shared by both march variants).  nd = {1/(2j), j/2}.
#define WG_NODE_FWD(uj, up1, um, nd, AXIS)                                              \
  {                                                                                     \
    const float ui = (uj) * idx;                                                        \
    float bb = ui + N2, dd = ui * (uj), a = 0.f, cc = -2.f * N2;                        \
    if (AXIS) {                                                                         \
      bb += N2; /* axis node: diagonal U_0/dx + 4N, super-diagonal -4N, no sub-diagonal */ \
    } else {                                                                            \
      const float du = (up1) - (um), su = (up1) + (um);                                 \
      const float hd = (nd).x * du;                                                     \
      const float t2 = fmaf(-2.f, (uj), su);      /* lap dr^2 - hd */                   \
      const float Wp = W + rgh;                   /* W = 1 + I */                       \
      const float den = fmaf(-0.25f, du, (uj));                                         \
      const float G = fmaf(hd, Wp, t2) * rcp_fast(den);                                 \
      rgh = (nd).y * G;                                                                 \
      W = Wp + rgh;                                                                     \
      const float q = (nd).x * W;                 /* h (1 + I) */                       \
      a = fmaf(-N, q, N);                                                               \
      cc = fmaf(-N, q, -N);                                                             \
    }                                                                                   \
    const float m = rcp_fast(fmaf(a, cpm, bb));                                         \
    cpm = cc * m;                                                                       \
    dpm = fmaf(a, dpm, dd) * m;                                                         \
  }

// Variant 0: everything unrolled, c' in 63 registers (grid constants become immediates).
__device__ __forceinline__ float march_row_regs(float* __restrict__ row, int key, float dxt, float xt, float knu1) {
  constexpr float IDR2 = 1.f / (DR * DR);
  float cp[WG_NR];
  float4 cur = ld_chunk(row, key, 0);
  const float bw = ld_chunk(row, key, WG_NR / 4 - 1).w;
  const float nu = knu1 * f1_filter(xt) + K2 * f2_filter(xt) * bw;
  const float idx = 1.f / fmaxf(dxt, DXT_MIN);
  const float N = nu * IDR2, N2 = 2.f * N;
  float W = 1.f, rgh = 0.f, cpm = 0.f, dpm = 0.f, um = 0.f;
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
      const float2 nd = make_float2(j ? 1.f / (2.f * j) : 0.f, 0.5f * j);
      WG_NODE_FWD(uu[e], uu[e + 1], um, nd, j == 0)
      cp[j] = cpm;
      dout[e] = dpm;
      um = uu[e];
    }
    st_chunk(row, key, c, make_float4(dout[0], dout[1], dout[2], dout[3]));
    cur = nxt;
  }
  // ---- back substitution (U_63 = 1) + shear integrals of the new profile; slot 63 is rewritten afterwards
  float un = 1.f, Mh = 0.f, umin = 1.f;
#pragma unroll
  for (int c = WG_NR / 4 - 1; c >= 0; --c) {
    const float4 dq = ld_chunk(row, key, c);
    const float dv[4] = {dq.x, dq.y, dq.z, dq.w};
    float o[4];
#pragma unroll
    for (int e = 3; e >= 0; --e) {
      const int j = 4 * c + e;
      if (j == WG_NR - 1) { o[e] = 0.f; continue; }  // node 63 carries no unknown
      un = fmaf(-cp[j], un, dv[e]);
      o[e] = un;
      Mh = fmaf(0.5f * j, 1.f - un, Mh);
      umin = fminf(umin, un);
    }
    st_chunk(row, key, c, make_float4(o[0], o[1], o[2], o[3]));
  }
  // M = dr^2 * sum j (1 - U_j) = 2 dr^2 Mh  ->  bw = sqrt(2 M (1 - Umin)) = 2 dr sqrt(Mh (1 - Umin))
  row[(((WG_NR / 4 - 1) ^ key) << 2) | 3] = 2.f * DR * sqrtf(fmaxf(Mh * (1.f - umin), 0.f));
  return un;
}

// Variant 1: both sweeps rolled (4 nodes per iteration), c' in a private shared-memory scratch row (same swizzle),
// per-node grid constants from c_node.  Small code (instruction-cache resident), few registers, twice the shared
// memory per station.
__device__ __forceinline__ float march_row_smem(float* __restrict__ row, float* __restrict__ cps, int key, float dxt,
                                                float xt, float knu1) {
  constexpr float IDR2 = 1.f / (DR * DR);
  float4 cur = ld_chunk(row, key, 0);
  const float bw = ld_chunk(row, key, WG_NR / 4 - 1).w;
  const float nu = knu1 * f1_filter(xt) + K2 * f2_filter(xt) * bw;
  const float idx = 1.f / fmaxf(dxt, DXT_MIN);
  const float N = nu * IDR2, N2 = 2.f * N;
  float W = 1.f, rgh = 0.f, cpm = 0.f, dpm = 0.f, um = 0.f;
#pragma unroll 1
  for (int c = 0; c < WG_NR / 4; ++c) {
    float4 nxt = ld_chunk(row, key, min(c + 1, WG_NR / 4 - 1));
    if (c + 1 == WG_NR / 4 - 1) nxt.w = 1.f;
    const float uu[5] = {cur.x, cur.y, cur.z, cur.w, nxt.x};
    float dout[4], cout[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 nd = c_node[4 * c + e];
      WG_NODE_FWD(uu[e], uu[e + 1], um, nd, (e == 0 && c == 0))
      cout[e] = cpm;
      dout[e] = dpm;
      um = uu[e];
    }
    st_chunk(row, key, c, make_float4(dout[0], dout[1], dout[2], dout[3]));
    st_chunk(cps, key, c, make_float4(cout[0], cout[1], cout[2], cout[3]));
    cur = nxt;
  }
  float un = 1.f, Mh = 0.f, umin = 1.f;
#pragma unroll 1
  for (int c = WG_NR / 4 - 1; c >= 0; --c) {
    const float4 dq = ld_chunk(row, key, c);
    const float4 cq = ld_chunk(cps, key, c);
    const float dv[4] = {dq.x, dq.y, dq.z, dq.w}, cv[4] = {cq.x, cq.y, cq.z, cq.w};
    float o[4];
#pragma unroll
    for (int e = 3; e >= 0; --e) {
      const float2 nd = c_node[4 * c + e];
      const float cand = fmaf(-cv[e], un, dv[e]);
      if (e == 3 && c == WG_NR / 4 - 1) {
        o[e] = 0.f;  // node 63 carries no unknown (its d', c' are dummies)
      } else {
        un = cand;
        o[e] = un;
        Mh = fmaf(nd.y, 1.f - un, Mh);
        umin = fminf(umin, un);
      }
    }
    st_chunk(row, key, c, make_float4(o[0], o[1], o[2], o[3]));
  }
  row[(((WG_NR / 4 - 1) ^ key) << 2) | 3] = 2.f * DR * sqrtf(fmaxf(Mh * (1.f - umin), 0.f));
  return un;
}

// ---------------------------------------------------------------------------------------------- TMEM scratch
// Variant 2 keeps the Thomas coefficients c' in Blackwell tensor memory instead of 63 registers: every thread owns
// one TMEM lane (warp w of the CTA addresses lanes 32 (w & 3) .. +31), node j lives in column j of the CTA's
// 64-column allocation.  tcgen05.st / tcgen05.ld with shape 32x32b move 4 consecutive columns of the thread's own
// lane per instruction (SASS STTM / LDTM), so TMEM acts as a software-managed per-thread scratch; nothing here
// touches the tensor cores.  The freed registers buy a fifth resident CTA per SM and let the march be a rolled
// loop that stays inside the instruction cache.
#define WG_TMEM_COLS 64
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "n"(WG_TMEM_COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(WG_TMEM_COLS) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, float a, float b, float c, float d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}
__device__ __forceinline__ float4 tmem_ld4(uint32_t taddr) {
  float4 v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(taddr)
               : "memory");
  return v;
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Variant 2: both sweeps rolled over the row's sixteen 4-node chunks, c' in the thread's TMEM lane (taddr = lane
// base + column 0).  Executed by all 32 lanes of the warp (tcgen05.ld/st are warp-collective).
__device__ __forceinline__ float march_row_tmem(float* __restrict__ row, uint32_t taddr, int key, float dxt, float xt,
                                                float knu1) {
  constexpr float IDR2 = 1.f / (DR * DR);
  constexpr int NC = WG_NR / 4;
  float4 cur = ld_chunk(row, key, 0);
  const float bw = ld_chunk(row, key, NC - 1).w;
  const float nu = knu1 * f1_filter(xt) + K2 * f2_filter(xt) * bw;
  const float idx = 1.f / fmaxf(dxt, DXT_MIN);
  const float N = nu * IDR2, N2 = 2.f * N;
  float W = 1.f, rgh = 0.f, cpm = 0.f, dpm = 0.f, um = 0.f;
  {  // chunk 0 holds the axis node
    const float4 nxt = ld_chunk(row, key, 1);
    const float uu[5] = {cur.x, cur.y, cur.z, cur.w, nxt.x};
    float dout[4], cout[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 nd = make_float2(e ? 1.f / (2.f * e) : 0.f, 0.5f * e);
      WG_NODE_FWD(uu[e], uu[e + 1], um, nd, e == 0)
      cout[e] = cpm;
      dout[e] = dpm;
      um = uu[e];
    }
    st_chunk(row, key, 0, make_float4(dout[0], dout[1], dout[2], dout[3]));
    tmem_st4(taddr, cout[0], cout[1], cout[2], cout[3]);
    cur = nxt;
  }
#pragma unroll 1
  for (int c = 1; c < NC; ++c) {
    float4 nxt = ld_chunk(row, key, min(c + 1, NC - 1));
    if (c + 1 == NC - 1) nxt.w = 1.f;  // Dirichlet node: the slot holds bw, the value is 1
    const float uu[5] = {cur.x, cur.y, cur.z, cur.w, nxt.x};
    float dout[4], cout[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 nd = c_node[4 * c + e];
      WG_NODE_FWD(uu[e], uu[e + 1], um, nd, false)
      cout[e] = cpm;
      dout[e] = dpm;
      um = uu[e];
    }
    st_chunk(row, key, c, make_float4(dout[0], dout[1], dout[2], dout[3]));
    tmem_st4(taddr + 4 * c, cout[0], cout[1], cout[2], cout[3]);
    cur = nxt;
  }
  tmem_wait_st();
  // ---- back substitution (U_63 = 1); chunk 15 carries the Dirichlet slot and is peeled
  float un = 1.f, Mh = 0.f, umin = 1.f;
  float4 cq = tmem_ld4(taddr + 4 * (NC - 1));
  tmem_wait_ld();
  {
    const float4 dq = ld_chunk(row, key, NC - 1);
    float4 cn = tmem_ld4(taddr + 4 * (NC - 2));
    float o[3];
    const float dv[3] = {dq.x, dq.y, dq.z}, cv[3] = {cq.x, cq.y, cq.z};
#pragma unroll
    for (int e = 2; e >= 0; --e) {
      un = fmaf(-cv[e], un, dv[e]);
      o[e] = un;
      Mh = fmaf(0.5f * (4 * (NC - 1) + e), 1.f - un, Mh);
      umin = fminf(umin, un);
    }
    st_chunk(row, key, NC - 1, make_float4(o[0], o[1], o[2], 0.f));
    tmem_wait_ld();
    cq = cn;
  }
#pragma unroll 1
  for (int c = NC - 2; c >= 0; --c) {
    const float4 dq = ld_chunk(row, key, c);
    float4 cn = tmem_ld4(taddr + 4 * max(c - 1, 0));
    const float dv[4] = {dq.x, dq.y, dq.z, dq.w}, cv[4] = {cq.x, cq.y, cq.z, cq.w};
    float o[4];
    const float jb = 2.f * (float)c;  // j/2 of the chunk's first node
#pragma unroll
    for (int e = 3; e >= 0; --e) {
      un = fmaf(-cv[e], un, dv[e]);
      o[e] = un;
      Mh = fmaf(jb + 0.5f * e, 1.f - un, Mh);
      umin = fminf(umin, un);
    }
    st_chunk(row, key, c, make_float4(o[0], o[1], o[2], o[3]));
    tmem_wait_ld();
    cq = cn;
  }
  row[(((NC - 1) ^ key) << 2) | 3] = 2.f * DR * sqrtf(fmaxf(Mh * (1.f - umin), 0.f));
  return un;
}

// #{k : xs[k] < x} over a sorted array padded with +inf to 2*top entries (top = power of two, 2*top-1 >= n)
__device__ __forceinline__ int count_below(const float* xs, int top, float x) {
  int c = 0;
  for (int s = top; s > 0; s >>= 1)
    if (xs[c + s - 1] < x) c += s;
  return c;
}

struct LaneLoc {
  int chain, slot, q, valid;
};

template <class SH>
__device__ __forceinline__ LaneLoc locate(const SH& sh, int tile, int lane, int T, int P, int ntot) {
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
// 16 lanes, shuffle-reduced to the rotor average, accumulated per rotor in the warp's private accumulators
// (fixed order -> bit-reproducible; the turbine epilogue adds the warps' partial sums).
__device__ __noinline__ void flush_hits(const float* __restrict__ tile, const float4* __restrict__ ha,
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
    d *= (1.f / WG_NQ);
    if (lane == 0) {
      acc_du[b.y] = fmaf(a.x, d, acc_du[b.y]);
      acc_dv[b.y] = fmaf(a.y, d, acc_dv[b.y]);
    }
    __syncwarp();
    if (lane == 16 && ok) {
      acc_du[b.y] = fmaf(a.x, d, acc_du[b.y]);
      acc_dv[b.y] = fmaf(a.y, d, acc_dv[b.y]);
    }
    __syncwarp();
  }
}

template <int VARIANT, int SYNC_ROUNDS, int WG_NWARP, int TC>
__global__ void __launch_bounds__(WG_NWARP * 32, (VARIANT == 0 ? 16 : (VARIANT == 1 ? 12 : 20)) / WG_NWARP)
    wg_flow_kernel(const Dev d, const FlowArgs a) {
  static_assert(VARIANT != 2 || WG_NWARP == 4, "the TMEM scratch maps one warp per 32-lane quarter");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  typedef FlowShared<WG_NWARP, TC> Shared;
  Shared& sh = *reinterpret_cast<Shared*>(smem_raw);
  const int T = d.T, P = d.P, F = d.F;
  float* bufs = reinterpret_cast<float*>(smem_raw + hdr_bytes<WG_NWARP, TC>());

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
  float* tile_base = bufs + (size_t)warp * (VARIANT == 1 ? 2 : 1) * WG_TILE * WG_NR;
  float* row = tile_base + lane * WG_NR;
  float* cps = row + WG_TILE * WG_NR;  // variant 1: c' scratch row behind the warp's tile
  (void)cps;
  const float qy = c_qy[lane & 15], qz = c_qz[lane & 15];
  void* bar = &sh.mbar[warp];

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
  if (VARIANT == 2 && warp == 0) tmem_alloc(&sh.tmem_base);
  for (int i = T + tid; i < 2 * TC; i += blockDim.x) sh.xs[i] = CUDART_INF_F;
  if (lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  int nstep = d.n_step[bf];
  uint32_t phase = 0;
  if (VARIANT == 2) asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  uint32_t taddr = 0;
  if (VARIANT == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    taddr = sh.tmem_base + ((uint32_t)(warp * 32) << 16);
  }
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
  int xs_top = 1;
  while (2 * xs_top - 1 < T) xs_top <<= 1;
  const float x_retire = xmax + MARGIN_D * d.D;

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
        if (xn > x_retire) --cnt; else break;
      }
      sh.count[tid] = cnt;
    }
    for (int i = tid; i < WG_NWARP * TC; i += blockDim.x) {
      (&sh.acc_du[0][0])[i] = 0.f;
      (&sh.acc_dv[0][0])[i] = 0.f;
    }
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
    float* acc_du = sh.acc_du[warp];
    float* acc_dv = sh.acc_dv[warp];
    float4* ha = sh.hit_a[warp];
    int2* hb = sh.hit_b[warp];
    const unsigned full = 0xffffffffu;
    const unsigned lt = (1u << lane) - 1u;
    int tile = warp;
    LaneLoc Ln;
    Ln.valid = 0; Ln.chain = 0; Ln.slot = 0; Ln.q = 0;
    if (tile < ntiles) Ln = locate(sh, tile, lane, T, P, ntot);
    const int nrounds = (ntiles + WG_NWARP - 1) / WG_NWARP;
    for (int rnd = 0; rnd < nrounds; ++rnd) {
      if (SYNC_ROUNDS) __syncthreads();  // the CTA's warps (one per SM sub-partition) enter the march together
      if (tile >= ntiles) continue;
      const LaneLoc Lc = Ln;
      const Seg sg = segments(Lc, lane);
      if (lane == 0) mbar_expect_tx(bar, (uint32_t)sg.nvalid * WG_ROW_BYTES);
      __syncwarp();
      float* gsrc = prof + ((size_t)Lc.chain * P + Lc.slot) * WG_NR;
      if (sg.start) bulk_g2s(row, gsrc, (uint32_t)sg.len * WG_ROW_BYTES, bar);
      float4 pmc = make_float4(0.f, 0.f, 0.f, 0.f), pcc = pmc;
      if (Lc.valid) {
        pmc = __ldcg(reinterpret_cast<const float4*>(pm_old + ((size_t)Lc.chain * P + Lc.slot) * 4));
        pcc = __ldcg(reinterpret_cast<const float4*>(pcon + ((size_t)Lc.chain * P + Lc.slot) * 4));
      }
      // while the tile is in flight: find the next one and pull its rows towards L2
      const int nt = tile + WG_NWARP;
      if (nt < ntiles) {
        Ln = locate(sh, nt, lane, T, P, ntot);
        if (Ln.valid) {
          const size_t st = (size_t)Ln.chain * P + Ln.slot;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(prof + st * WG_NR));
          if (lane == 0 || (Ln.slot & 7) == 0) {  // the station scalars: one 128-byte line per 8 slots
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pm_old + st * 4));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pcon + st * 4));
          }
        }
      }
      const int key = Lc.slot & 7;
      float xn = 0.f, yn = 0.f, zn = 0.f, dx = 0.f;
      if (Lc.valid) moved(pmc, pcc, ws, dt, xn, yn, zn, dx);
      mbar_wait(bar, phase);
      phase ^= 1u;
      if (VARIANT == 2) {  // warp-collective TMEM traffic: idle lanes march their (stale) row too
        const float xt_mid = (pmc.x + 0.5f * dx - sh.xr[Lc.chain]) * rR;
        const float ucn = march_row_tmem(row, taddr, key, dx * rR, xt_mid, pcc.y);
        if (Lc.valid)
          *reinterpret_cast<float4*>(pm_new + ((size_t)Lc.chain * P + Lc.slot) * 4) = make_float4(xn, yn, zn, ucn);
      } else if (Lc.valid) {
        const float xt_mid = (pmc.x + 0.5f * dx - sh.xr[Lc.chain]) * rR;
        const float ucn = VARIANT == 0 ? march_row_regs(row, key, dx * rR, xt_mid, pcc.y)
                                       : march_row_smem(row, cps, key, dx * rR, xt_mid, pcc.y);
        *reinterpret_cast<float4*>(pm_new + ((size_t)Lc.chain * P + Lc.slot) * 4) = make_float4(xn, yn, zn, ucn);
      }
      fence_async_smem();
      __syncwarp();
      if (sg.start) {  // write the marched rows back (same segments as the load)
        bulk_s2g(gsrc, row, (uint32_t)sg.len * WG_ROW_BYTES);
        bulk_commit();
      }
      // ---- superposition: which rotor planes does this station bracket together with its age neighbours?
      {
        // older neighbour = flat index - 1 (same chain), younger = flat index + 1
        float xo = __shfl_up_sync(full, xn, 1), yo = __shfl_up_sync(full, yn, 1), zo = __shfl_up_sync(full, zn, 1);
        const int ch_o = __shfl_up_sync(full, Lc.chain, 1);
        float xy = __shfl_down_sync(full, xn, 1), yy = __shfl_down_sync(full, yn, 1), zy = __shfl_down_sync(full, zn, 1);
        const int ch_y = __shfl_down_sync(full, Lc.chain, 1);
        const int vy_ = __shfl_down_sync(full, Lc.valid, 1);
        const bool has_o = Lc.valid && Lc.q > 0;
        const bool has_y = Lc.valid && Lc.q < sh.count[Lc.chain] - 1;
        if (has_o && (lane == 0 || ch_o != Lc.chain)) {
          int so = Lc.slot == 0 ? P - 1 : Lc.slot - 1;
          float4 pm = __ldcg(reinterpret_cast<const float4*>(pm_old + ((size_t)Lc.chain * P + so) * 4));
          float4 pc = __ldcg(reinterpret_cast<const float4*>(pcon + ((size_t)Lc.chain * P + so) * 4));
          float dxx;
          moved(pm, pc, ws, dt, xo, yo, zo, dxx);
        }
        if (has_y && (lane == 31 || !vy_ || ch_y != Lc.chain)) {
          int sy = Lc.slot == P - 1 ? 0 : Lc.slot + 1;
          float4 pm = __ldcg(reinterpret_cast<const float4*>(pm_old + ((size_t)Lc.chain * P + sy) * 4));
          float4 pc = __ldcg(reinterpret_cast<const float4*>(pcon + ((size_t)Lc.chain * P + sy) * 4));
          float dxx;
          moved(pm, pc, ws, dt, xy, yy, zy, dxx);
        }
        // Rotor planes bracketed by this station and its age neighbours: with c(x) = #{k : xs[k] < x} over the sorted
        // plane positions, plane k lies in [x1, x2) iff c(x1) <= k < c(x2).  Interval A = [self (younger end), older
        // neighbour), interval B = [younger neighbour, self (older end)); an inverted pair counts with sign -1
        // (oracle/dwm_numpy.py:353-356).  Every lane handles its own row for both of its intervals.
        const int cn = count_below(sh.xs, xs_top, xn);
        const int co = has_o ? count_below(sh.xs, xs_top, xo) : cn;
        const int cy = has_y ? count_below(sh.xs, xs_top, xy) : cn;
        const int a_lo = min(cn, co), nA = abs(cn - co), b_lo = min(cn, cy), nB = abs(cn - cy);
        const int nmax = __reduce_max_sync(full, max(nA, nB));
        const float u0cg = pcc.x * pcc.z, u0sg = pcc.x * pcc.w;
        const int rowkey = lane | (key << 8);
        const float rdA = rcp_fast(xo - xn), rdB = rcp_fast(xn - xy);
        int nh = 0;
        for (int it = 0; it < nmax; ++it) {
          const int kA = a_lo + it, kB = b_lo + it;
          const int jA = it < nA ? sh.ord[kA] : Lc.chain, jB = it < nB ? sh.ord[kB] : Lc.chain;
          const bool hitA = jA != Lc.chain, hitB = jB != Lc.chain;
          const unsigned mA = __ballot_sync(full, hitA), mB = __ballot_sync(full, hitB);
          if (hitA) {
            const float w = (sh.xs[kA] - xn) * rdA;
            const float wg = kA >= cn ? 1.f - w : w - 1.f;
            const float yc = fmaf(w, yo - yn, yn), zc = fmaf(w, zo - zn, zn);
            const int p = nh + __popc(mA & lt);
            ha[p] = make_float4(wg * u0cg, wg * u0sg, (sh.yr[jA] - yc) * rR, (d.zh - zc) * rR);
            hb[p] = make_int2(rowkey, jA);
          }
          nh += __popc(mA);
          if (hitB) {
            const float w = (sh.xs[kB] - xy) * rdB;
            const float wg = kB >= cy ? w : -w;
            const float yc = fmaf(w, yn - yy, yy), zc = fmaf(w, zn - zy, zy);
            const int p = nh + __popc(mB & lt);
            ha[p] = make_float4(wg * u0cg, wg * u0sg, (sh.yr[jB] - yc) * rR, (d.zh - zc) * rR);
            hb[p] = make_int2(rowkey, jB);
          }
          nh += __popc(mB);
          if (nh > WG_HIT_CAP - 64) {
            __syncwarp();
            flush_hits(tile_base, ha, hb, nh, acc_du, acc_dv, lane, qy, qz);
            nh = 0;
          }
        }
        __syncwarp();
        if (nh > 0) flush_hits(tile_base, ha, hb, nh, acc_du, acc_dv, lane, qy, qz);
      }
      __syncwarp();
      bulk_wait_read0();  // the store has drained its shared-memory reads: the buffer can be refilled
      __syncwarp();
      tile = nt;
    }
    bulk_wait_all0();
    fence_async_all();
    __syncthreads();

    // ------------------------------------------------------------------ turbine epilogue
    const bool emit = (nstep % k_emit) == 0;
    if (tid < T) {
      float du = 0.f, dv = 0.f;
#pragma unroll
      for (int wi = 0; wi < WG_NWARP; ++wi) { du += sh.acc_du[wi][tid]; dv += sh.acc_dv[wi][tid]; }
      const float u = ws - du, v = dv, w = 0.f;
      const float yaw = sh.yaw[tid];
      float sg, cg;
      sincosf(yaw * 0.017453292519943295f, &sg, &cg);
      const float wse = u * cg;
      const float pw = tab_interp(d.tab_ws, d.tab_p, d.n_tab, wse);
      float ct = tab_interp(d.tab_ws, d.tab_ct, d.n_tab, wse) * cg * cg;
      ct = fminf(fmaxf(ct, 0.f), CT_MAX);
      sh.u[tid] = u; sh.v[tid] = v; sh.w[tid] = w; sh.pw[tid] = pw; sh.ct[tid] = ct;
      sh.cg[tid] = cg; sh.sg[tid] = sg;
      int slot = -1;
      if (emit) {
        sh.ind[tid] = 0.5f * (1.f - sqrtf(1.f - ct));
        slot = sh.head[tid];
        if (sh.count[tid] == P) atomicOr(&d.flags[b], 2); else sh.count[tid] += 1;
        sh.head[tid] = (slot + 1 == P) ? 0 : slot + 1;
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
    if (emit) {
      // release one particle per turbine: cell-averaged top-hat inlet (IEC 61400-1 ed.4 Annex E).  16 threads per
      // turbine write one 16-byte chunk each and shuffle-reduce the row's shear integral into slot 63.
      for (int idx = tid; idx < T * (WG_NR / 4); idx += blockDim.x) {  // T*16 is a multiple of 16: half-warps stay whole
        const int t = idx >> 4, c = idx & 15;
        const int slot = sh.emit_slot[t];
        const float ind = sh.ind[t];
        const float fw = 1.f - 0.45f * ind * ind;
        const float rw2 = fw * fw * (1.f - ind) / (1.f - 2.f * ind);
        float vals[4], Mh = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = 4 * c + e;
          const float rlo = fmaxf((float)j - 0.5f, 0.f) * DR, rhi = ((float)j + 0.5f) * DR;
          const float frac = fminf(fmaxf((rw2 - rlo * rlo) / (rhi * rhi - rlo * rlo), 0.f), 1.f);
          const float df = (j == WG_NR - 1) ? 0.f : 2.f * ind * frac;
          vals[e] = 1.f - df;
          Mh = fmaf(0.5f * (float)j, df, Mh);
        }
        const unsigned hm = 0xffffu << (lane & 16);
        Mh += __shfl_xor_sync(hm, Mh, 8);
        Mh += __shfl_xor_sync(hm, Mh, 4);
        Mh += __shfl_xor_sync(hm, Mh, 2);
        Mh += __shfl_xor_sync(hm, Mh, 1);
        const float u0v = __shfl_sync(hm, vals[0], lane & 16);  // centre value = Umin of the monotone inlet
        if (slot < 0) continue;
        if (c == WG_NR / 4 - 1) vals[3] = 2.f * DR * sqrtf(fmaxf(Mh * (1.f - u0v), 0.f));
        *reinterpret_cast<float4*>(prof + ((size_t)t * P + slot) * WG_NR + ((c ^ (slot & 7)) << 2)) =
            make_float4(vals[0], vals[1], vals[2], vals[3]);
        if (c == 0) {
          *reinterpret_cast<float4*>(pm_new + ((size_t)t * P + slot) * 4) = make_float4(sh.xr[t], sh.yr[t], d.zh, u0v);
          *reinterpret_cast<float4*>(pcon + ((size_t)t * P + slot) * 4) =
              make_float4(sh.u[t], knu1_env, sh.cg[t], sh.sg[t]);
        }
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
  if (VARIANT == 2 && warp == 0) {  // every warp's last TMEM access precedes the substep loop's closing barrier
    __syncwarp();
    tmem_dealloc(sh.tmem_base);
  }
}

// WG_FLOW_VARIANT: 0 = c' in registers, unrolled march; 1 = c' in shared memory, rolled march; 2 = c' in TMEM,
// rolled march.  WG_FLOW_SYNC=1: the CTA's warps start every tile round together (instruction-cache sharing).
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && e[0] >= '0' && e[0] <= '9') ? atoi(e) : dflt;
}

typedef void (*flow_fn)(const Dev, const FlowArgs);

template <int V, int SY, int NW, int TC>
static cudaError_t launch_as(const Dev& d, const FlowArgs& a, cudaStream_t s) {
  const flow_fn fn = wg_flow_kernel<V, SY, NW, TC>;
  const size_t smem = hdr_bytes<NW, TC>() + (size_t)NW * (V == 1 ? 2 : 1) * WG_TILE * WG_ROW_BYTES;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  fn<<<d.B * d.F, NW * 32, smem, s>>>(d, a);
  return cudaGetLastError();
}

template <int V, int SY>
static cudaError_t launch_tc(const Dev& d, const FlowArgs& a, cudaStream_t s) {
  return d.T <= 16 ? launch_as<V, SY, 4, 16>(d, a, s) : launch_as<V, SY, 4, WG_MAX_T>(d, a, s);
}

cudaError_t launch_flow(const Dev& d, const FlowArgs& a, cudaStream_t s) {
  static const int v = env_int("WG_FLOW_VARIANT", 2);
  static const int sync = env_int("WG_FLOW_SYNC", 1) ? 1 : 0;
  if (v == 0) return sync ? launch_tc<0, 1>(d, a, s) : launch_tc<0, 0>(d, a, s);
  if (v == 1) return sync ? launch_tc<1, 1>(d, a, s) : launch_tc<1, 0>(d, a, s);
  return sync ? launch_tc<2, 1>(d, a, s) : launch_tc<2, 0>(d, a, s);
}

}  // namespace wg
