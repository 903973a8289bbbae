// wg_flow_kernel -- the DWM flow step (dynamiks seam: DWMFlowSimulation.step, reference call sites
// WindGym/Wind_Farm_Env.py:734,:745,:945,:953) for thousands of independent farms in one launch.
//
// One CTA (4 warps) owns one (env, farm).  Its live wake stations (all turbine chains, ring-addressed) form one
// flat list that the CTA's warps stream in 32-station tiles:
//     cp.async.bulk (TMA 1-D bulk copy, mbarrier complete_tx)  HBM -> shared
//     while the tile is in flight: station scalars (own + the one age neighbour outside the tile), the move of the
//       wake centres, their positions among the sorted rotor planes, L2 prefetch of the next tile
//     thread-per-station implicit Ainslie march: ONE fused forward sweep (continuity-consistent radial
//       velocity + tridiagonal rows + Thomas elimination; d' written in place into the shared row, c' parked in
//       the thread's own TMEM lane with tcgen05.st) and one back substitution (tcgen05.ld) that also accumulates
//       the shear-layer integrals of the NEW profile for the next step's eddy viscosity
//     rotor-plane bracket detection (plane index ranges from the searches above; hits whose wake cannot reach the
//       rotor are dropped: exact zeros) -> per-warp hit list
//       -> (hit x quadrature point) mapped onto full warps, 16-lane shuffle reduction for the rotor average,
//       per-rotor partial sums in registers                                             (superposition gather)
//     cp.async.bulk shared -> HBM
// then the per-turbine epilogue (P/CT tables, particle release) runs in the same CTA, and the substep loop
// (dt_env/dt_sim, or a whole spin-up) repeats without leaving the kernel.  wg_step launches the envs longest first
// (FlowArgs::order) or, for single-substep steps, through a work table that cuts farms into parts when the batch
// leaves CTA slots free (FlowArgs::work); the stations a step retires were found by the lanes that marched them in
// the previous one.  Launch edges: the step's finish kernel is a programmatic dependent of this kernel below one
// wave of CTAs (pdl_trigger), and this kernel is a programmatic dependent of the PREVIOUS step's finish kernel
// (pdl_wait: everything up to the turbine epilogue touches the wake state only and overlaps it).
// Algorithmic traffic per station and step: 256 B profile + 16 B mutable + 16 B emission scalars read,
// 256 B + 16 B written = 560 B.  HBM/issue bound; no tensor cores (stencil + gather).
//
// Profile row layout (64 floats, 16-byte chunks XOR-swizzled with slot & 7): nodes 0..62 hold U(r_j); node 63 is
// the Dirichlet node (U = 1 always), so its slot carries bw = sqrt(2 M (1 - Umin)) of the row instead -- the
// shear-layer term of the eddy viscosity (oracle/dwm_numpy.py:113-115), computed when the row was last written.
//
// TMEM as scratch: the 63 Thomas coefficients c' of a row would cost 63 registers per thread.  Every thread owns
// one TMEM lane instead (warp w of the CTA addresses lanes 32 w .. 32 w + 31 of the CTA's 64-column allocation,
// node j = column j); tcgen05.st / tcgen05.ld with shape 32x32b move 4 consecutive columns of the thread's own
// lane per instruction (SASS STTM / LDTM).  Nothing here touches the tensor cores.  The freed registers buy a
// fifth resident CTA per SM and let the march be a rolled loop that stays inside the instruction cache.
#include <math_constants.h>

#include <algorithm>
#include <cstdlib>

#include "wg_internal.cuh"

namespace wg {

#define WG_NWARP 4        // warps per CTA (one per 32-lane TMEM quarter)
#define WG_HIT_CAP 32     // per-warp hit list entries (one detection pass of one interval kind adds at most 32)
#define WG_TAB_CAP 32     // P/CT table knots staged in shared memory (longer tables are read from global)
#define WG_TMEM_COLS 64
// squared centre distance (rotor radii) beyond which a wake profile cannot reach any rotor quadrature point
#define R_CULL2 (((WG_NR - 1) * DR + 1.01f) * ((WG_NR - 1) * DR + 1.01f))
#define WG_RETIRE_CAP 8   // stations per chain and step that can retire (physically one, two when the wake compresses)

__constant__ float c_qy[WG_NQ];
__constant__ float c_qz[WG_NQ];
// rho_j = j / (j + 1) of the radial grid r_j = j dr, four nodes per entry (see WG_NODE_FWD)
__constant__ float4 c_rho4[WG_NR / 4];

void set_rotor_points(const float* qy, const float* qz) {
  cudaMemcpyToSymbol(c_qy, qy, sizeof(float) * WG_NQ);
  cudaMemcpyToSymbol(c_qz, qz, sizeof(float) * WG_NQ);
  float4 nd[WG_NR / 4];
  for (int j = 0; j < WG_NR; j += 4)
    nd[j / 4] = make_float4(j / (j + 1.f), (j + 1) / (j + 2.f), (j + 2) / (j + 3.f), (j + 3) / (j + 4.f));
  cudaMemcpyToSymbol(c_rho4, nd, sizeof(nd));
}

// Per-CTA bookkeeping in front of the tile buffers.  TC = turbine capacity of the tables (16 or WG_MAX_T: small
// farms leave the shared memory to more resident CTAs -- with TC = 16 the CTA needs 37 KB, six fit on an SM).
template <int TC, int TURB>
struct __align__(16) FlowShared {
  unsigned long long mbar[WG_NWARP];
  float xr[TC], yr[TC], yaw[TC], u[TC], v[TC], w[TC], pw[TC], ct[TC], ind[TC], cg[TC], sg[TC];
  float xs[2 * TC];                         // turbine x sorted ascending, padded with +inf
  float sum_ws[TC], sum_wd[TC], sum_yaw[TC], sum_pw[TC];
  float tu[TURB ? TC : 1], tv[TURB ? TC : 1], tw[TURB ? TC : 1];  // rotor-averaged ambient fluctuation
  // TURB == 2 (wake-added turbulence): isotropic box sampled at every rotor's quadrature points, per-warp sums
  float iso[TURB == 2 ? 3 * TC * WG_NQ : 1];
  // superposed deficit per rotor (and the wake-added turbulence, TURB == 2) as 32-bit fixed point, 2^-24 m/s per unit:
  // integer addition is associative, so every warp (and every CTA of a split farm, see FlowArgs::work) adds its
  // hits with shared / global atomics in ANY order and the sum is bit-identical -- alone, inside a batch, split or not
  int acc_du[TC], acc_dv[TC];
  int acc_ad[TURB == 2 ? 3 * TC : 1];
  int ticket;                               // split farms: arrival ticket of this CTA (the last one runs the epilogue)
  int ord[TC];                              // turbine index of xs[k]
  int head[TC], count[TC], pre[TC + 1];
  // tile loop: index (from the oldest) of the chain's first station that stays in the farm after the NEXT step's
  // move (atomicMin by the lanes that march them) -> the stations to retire next step; release: slot of the new particle
  int keep_emit[TC];
  float base_sum;
  uint32_t tmem_base;                       // TMEM allocation of the CTA
  float tab_ws[WG_TAB_CAP], tab_p[WG_TAB_CAP], tab_ct[WG_TAB_CAP];
  float4 hit_a[WG_NWARP][WG_HIT_CAP];       // w*U0e*cos g0, w*U0e*sin g0, ry, rz
  uint32_t hit_b[WG_NWARP][WG_HIT_CAP];     // (shared address of the row ^ (key << 4)) | rotor index j << 20
};

template <int TC, int TURB>
__host__ __device__ constexpr size_t hdr_bytes() { return (sizeof(FlowShared<TC, TURB>) + 255) / 256 * 256; }
// six CTAs per SM: 6 x (header + 32 KB of tiles + 1 KB reserved) must fit the 228 KB of the SM
static_assert(hdr_bytes<16, 0>() <= 5120 && hdr_bytes<16, 1>() <= 5120, "small-farm header outgrew 5 KB");

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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// TMA 1-D bulk copy shared -> global (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  // (an L2 evict_first hint on these copies: 0.334 vs 0.331 ms at 4096 envs, -1.5 % at 512: not worth a variant)
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ float rcp_fast(float x) {
  // (a correctly rounded reciprocal / square root / sine here change the deviation from the fp64 oracle by < 3 % --
  // it is float32 accumulation in the carried wake state, not these -- and cost 50 % of the kernel: measured)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float sqrt_fast(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// shared-memory accesses by 32-bit shared address (the swizzled chunk address is one XOR away from the row's)
__device__ __forceinline__ float4 lds4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float lds1(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts1(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// ---------------------------------------------------------------------------------------------- TMEM scratch
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
// the loaded registers are operands of the wait so that no consumer can be scheduled ahead of it
__device__ __forceinline__ void tmem_wait_ld(float4& a, float4& b) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+f"(a.x), "+f"(a.y), "+f"(a.z), "+f"(a.w), "+f"(b.x), "+f"(b.y), "+f"(b.z), "+f"(b.w)
               :
               : "memory");
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

// np.interp semantics (clamped ends) for power and CT on the same abscissa: one bracket search, two interpolations
__device__ __forceinline__ void tab_interp2(const float* __restrict__ xs, const float* __restrict__ y1,
                                            const float* __restrict__ y2, int n, float x, float& o1, float& o2) {
  if (x <= xs[0]) { o1 = y1[0]; o2 = y2[0]; return; }
  if (x >= xs[n - 1]) { o1 = y1[n - 1]; o2 = y2[n - 1]; return; }
  int lo = 0, hi = n - 1;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (xs[mid] <= x) lo = mid; else hi = mid;
  }
  const float x0 = xs[lo], dxk = xs[lo + 1] - x0;
  o1 = (y1[lo + 1] - y1[lo]) / dxk * (x - x0) + y1[lo];
  o2 = (y2[lo + 1] - y2[lo]) / dxk * (x - x0) + y2[lo];
}

// new wake-centre position after one step (Hill-vortex self-induced velocity added to the ambient)
// tv = low-pass ambient (v', w') at the centre (meandering; zero for uniform inflow)
__device__ __forceinline__ void moved(const float4 pm, const float4 pc, float ws, float dt, const float2 tv, float& xn,
                                      float& yn, float& zn, float& dx) {
  float kd = K_HILL * (1.f - pm.w) * pc.x;
  float vx = ws - kd * pc.z;
  float vy = kd * pc.w + tv.x;
  dx = vx * dt;
  xn = pm.x + dx;
  yn = pm.y + vy * dt;
  zn = pm.z + tv.y * dt;
}

// Implicit Ainslie march of one profile row held in shared memory (oracle/dwm_numpy.py:ainslie_march).
// Forward sweep: per node j the continuity-consistent radial velocity (pass A of the oracle) feeds the
// tridiagonal row and its Thomas elimination (pass B) immediately; the eddy viscosity needs the row's shear
// integral bw, which rides in slot 63.  d' replaces U_j in the shared row, c' goes to the thread's TMEM lane.
// Back substitution (pass C) writes the new profile and accumulates its bw.  Returns the new centre value.
// With h = 1/(2j), N = nu/dr^2:
//   lap_j dr^2 = U_{j+1} + U_{j-1} - 2 U_j + h (U_{j+1} - U_{j-1}),  Vd_j = nu Vh_j / (2 dr) = -N h I_j,
//   sub-diagonal -a_j = -(N - N h (1 + I_j)), super-diagonal c_j = -N - N h (1 + I_j), diagonal U_j/dx + 2N
//   (the code multiplies each row by dx: diagonal U_j + 2 N dx, right-hand side U_j^2).
// The sweep carries Z = h (1 + I) instead of I: with h_j j/2 = 1/4 and rho_j = h_{j+1}/h_j = j/(j+1),
//   G_j = (lap_j dr^2 + h dU (I + rgh)) / den = ((su - 2 U_j) + dU Zp_j) / den,   Zp_j = h_j (1 + I_{j-1} + rgh_{j-1}),
//   Z_j = Zp_j + G_j/4,  Zp_{j+1} = rho_j (Z_j + G_j/4),  a_j = N - N Z_j,  c_j = -N - N Z_j   (Zp_1 = 1/2),
// so the only per-node grid constant is rho_j.
#define WG_NODE_FWD(uj, up1, um, rho, AXIS)                                             \
  {                                                                                     \
    float bb = (uj) + N2, dd = (uj) * (uj), a = 0.f, cc = -2.f * N2;                    \
    if (AXIS) {                                                                         \
      bb += N2; /* axis node: diagonal U_0/dx + 4N, super-diagonal -4N, no sub-diagonal */ \
    } else {                                                                            \
      const float du = (up1) - (um), su = (up1) + (um);                                 \
      const float t2 = fmaf(-2.f, (uj), su);                                            \
      const float den = fmaf(-0.25f, du, (uj));                                         \
      const float G = fmaf(du, Zp, t2) * rcp_fast(den);                                 \
      const float Z = fmaf(0.25f, G, Zp);                                               \
      a = fmaf(-N, Z, N);                                                               \
      cc = fmaf(-N, Z, -N);                                                             \
      Zp = (rho) * fmaf(0.25f, G, Z);                                                   \
    }                                                                                   \
    const float m = rcp_fast(fmaf(a, cpm, bb));                                         \
    cpm = cc * m;                                                                       \
    dpm = fmaf(a, dpm, dd) * m;                                                         \
  }

// rowk = shared address of the row ^ (key << 4): chunk c of the swizzled row sits at rowk ^ (c << 4).
// Both sweeps are rolled (8 nodes per iteration) with the end chunks peeled: chunk 0 holds the axis node, chunk
// 15 the Dirichlet slot.  Executed by all 32 lanes of the warp (tcgen05.ld/st are warp-collective).
__device__ __forceinline__ float march_row_tmem(uint32_t rowk, uint32_t taddr, float dxt, float xt, float knu1) {
  constexpr float IDR2 = 1.f / (DR * DR);
  constexpr int NC = WG_NR / 4;
  float4 cur = lds4(rowk);
  float4 nxt = lds4(rowk ^ 16u);
  const float bw = lds1((rowk ^ ((NC - 1) << 4)) + 12u);
  const float nu = knu1 * f1_filter(xt) + K2 * f2_filter(xt) * bw;
  // every row of the system is scaled by dx (Thomas' c', d' are invariant under row scaling): the diagonal is
  // U_j + 2 N dx and the right-hand side U_j^2, with N = nu dx / dr^2 -- one multiplication per node less
  const float N = nu * IDR2 * fmaxf(dxt, DXT_MIN), N2 = 2.f * N;
  float Zp = 0.5f, cpm = 0.f, dpm = 0.f, um = 0.f;
  {  // chunk 0
    const float uu[5] = {cur.x, cur.y, cur.z, cur.w, nxt.x};
    float dout[4], cout[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      WG_NODE_FWD(uu[e], uu[e + 1], um, e / (e + 1.f), e == 0)
      cout[e] = cpm;
      dout[e] = dpm;
      um = uu[e];
    }
    sts4(rowk, dout[0], dout[1], dout[2], dout[3]);
    tmem_st4(taddr, cout[0], cout[1], cout[2], cout[3]);
    cur = nxt;
  }
#pragma unroll 1
  for (int c = 1; c < NC - 1; c += 2) {  // chunks (1,2) .. (13,14)
    const uint32_t a0 = rowk ^ ((uint32_t)c << 4), a1 = rowk ^ ((uint32_t)(c + 1) << 4);
    const float4 n1 = lds4(a1);
    const float4 n2 = lds4(rowk ^ ((uint32_t)(c + 2) << 4));
    const float uu[9] = {cur.x, cur.y, cur.z, cur.w, n1.x, n1.y, n1.z, n1.w, n2.x};
    float dout[8], cout[8];
    const float4 r0 = c_rho4[c], r1 = c_rho4[c + 1];
    const float rho[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      WG_NODE_FWD(uu[e], uu[e + 1], um, rho[e], false)
      cout[e] = cpm;
      dout[e] = dpm;
      um = uu[e];
    }
    sts4(a0, dout[0], dout[1], dout[2], dout[3]);
    sts4(a1, dout[4], dout[5], dout[6], dout[7]);
    tmem_st4(taddr + 4 * c, cout[0], cout[1], cout[2], cout[3]);
    tmem_st4(taddr + 4 * c + 4, cout[4], cout[5], cout[6], cout[7]);
    cur = n2;
  }
  {  // chunk 15: nodes 60, 61, 62; node 63 is the Dirichlet node (value 1, its slot holds bw)
    const float uu[4] = {cur.x, cur.y, cur.z, 1.f};
    float dout[3], cout[3];
#pragma unroll
    for (int e = 0; e < 3; ++e) {
      const int j = 4 * (NC - 1) + e;
      WG_NODE_FWD(uu[e], uu[e + 1], um, j / (j + 1.f), false)
      cout[e] = cpm;
      dout[e] = dpm;
      um = uu[e];
    }
    // ---- back substitution starts in registers (U_63 = 1): nodes 62, 61, 60
    float un = 1.f, Mh = 0.f, umin = 1.f, o[3];
#pragma unroll
    for (int e = 2; e >= 0; --e) {
      un = fmaf(-cout[e], un, dout[e]);
      o[e] = un;
      Mh = fmaf(0.5f * (4 * (NC - 1) + e), 1.f - un, Mh);
      umin = fminf(umin, un);
    }
    const uint32_t aL = rowk ^ ((uint32_t)(NC - 1) << 4);
    sts4(aL, o[0], o[1], o[2], 0.f);
    tmem_wait_st();
    float4 cq1 = tmem_ld4(taddr + 4 * (NC - 2)), cq0 = tmem_ld4(taddr + 4 * (NC - 3));
#pragma unroll 1
    for (int c = NC - 2; c >= 2; c -= 2) {  // chunks (14,13) .. (2,1)
      const uint32_t a1 = rowk ^ ((uint32_t)c << 4), a0 = rowk ^ ((uint32_t)(c - 1) << 4);
      const float4 d1 = lds4(a1), d0 = lds4(a0);
      tmem_wait_ld(cq0, cq1);
      const float cv[8] = {cq0.x, cq0.y, cq0.z, cq0.w, cq1.x, cq1.y, cq1.z, cq1.w};
      const float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
      // next pair (or, last time round, chunk 0 twice: only cq1 is used then)
      cq1 = tmem_ld4(taddr + 4 * max(c - 2, 0));
      cq0 = tmem_ld4(taddr + 4 * max(c - 3, 0));
      float ov[8];
      // sum_e (jb + e/2)(1 - u_e) = jb (8 - A) + (14 - Bq) with A = sum u_e, Bq = sum (e/2) u_e: two instructions per
      // node instead of three (jb = j/2 of the pair's first node)
      float A = 0.f, Bq = 0.f;
#pragma unroll
      for (int e = 7; e >= 0; --e) {
        un = fmaf(-cv[e], un, dv[e]);
        ov[e] = un;
        A += un;
        Bq = fmaf(0.5f * e, un, Bq);
        umin = fminf(umin, un);
      }
      Mh += fmaf(2.f * (float)(c - 1), 8.f - A, 14.f - Bq);
      sts4(a1, ov[4], ov[5], ov[6], ov[7]);
      sts4(a0, ov[0], ov[1], ov[2], ov[3]);
    }
    {  // chunk 0
      const float4 d0 = lds4(rowk);
      tmem_wait_ld(cq0, cq1);
      const float cv[4] = {cq1.x, cq1.y, cq1.z, cq1.w}, dv[4] = {d0.x, d0.y, d0.z, d0.w};
      float ov[4];
#pragma unroll
      for (int e = 3; e >= 0; --e) {
        un = fmaf(-cv[e], un, dv[e]);
        ov[e] = un;
        Mh = fmaf(0.5f * e, 1.f - un, Mh);
        umin = fminf(umin, un);
      }
      sts4(rowk, ov[0], ov[1], ov[2], ov[3]);
    }
    // M = dr^2 * sum j (1 - U_j) = 2 dr^2 Mh  ->  bw = sqrt(2 M (1 - Umin)) = 2 dr sqrt(Mh (1 - Umin))
    sts1(aL + 12u, 2.f * DR * sqrtf(fmaxf(Mh * (1.f - umin), 0.f)));
    return un;
  }
}

// #{k : xs[k] < x} over a sorted array padded with +inf to 2*top entries (top = power of two, 2*top-1 >= n)
__device__ __forceinline__ int count_below(const float* xs, int top, float x) {
  int c = 0;
  for (int s = top; s > 0; s >>= 1)
    if (xs[c + s - 1] < x) c += s;
  return c;
}

// two independent searches in one loop (their shared-memory loads overlap)
__device__ __forceinline__ void count_below2(const float* xs, int top, float x1, float x2, int& c1, int& c2) {
  c1 = 0; c2 = 0;
  for (int s = top; s > 0; s >>= 1) {
    const float a = xs[c1 + s - 1], b = xs[c2 + s - 1];
    if (a < x1) c1 += s;
    if (b < x2) c2 += s;
  }
}

struct LaneLoc {
  int chain, slot, q, valid;
};

// Which chain / ring slot does flat station tile * 32 + lane belong to?  Farms of up to 32 chains: lane k reads the
// inclusive prefix pre[k + 1]; two ballots give the chains of the tile's first and last station, the (rare) chain
// boundaries inside the tile come by shuffle -- no dependent shared-memory search.  Larger farms: binary search.
template <int TC, class SH>
__device__ __forceinline__ LaneLoc locate(const SH& sh, int tile, int lane, int T, int P, int ntot) {
  LaneLoc L;
  const int f0 = tile * WG_TILE;
  int fl = f0 + lane;
  L.valid = fl < ntot;
  if (!L.valid) fl = ntot - 1;
  int lo;
  if (TC <= 32) {
    const unsigned full = 0xffffffffu;
    const int pre_hi = lane < T ? sh.pre[lane + 1] : 0x7fffffff;
    const int c0 = __popc(__ballot_sync(full, pre_hi <= f0));
    // chain ends inside the tile, as a bit mask over the tile's lanes: lane k (a chain) sets bit pre[k+1] - f0 when that
    // lies in (0, 31]; a station at lane L has passed every end at a position <= L.  One warp-wide OR instead of a
    // data-dependent shuffle loop (which the compiler turned into ~400 instructions per tile).  Empty chains make two
    // ends coincide on one bit: then (rare: the first steps of a spin-up) the ends are counted one by one.
    const unsigned rel = (unsigned)(pre_hi - f0);
    const bool inside = rel - 1u < (unsigned)(WG_TILE - 1) && pre_hi <= ntot - 1;   // 1 <= rel <= 31, a station follows
    const unsigned ends = __reduce_or_sync(full, inside ? 1u << rel : 0u);
    const int n_in = __popc(__ballot_sync(full, inside));
    lo = c0 + __popc(ends & (0xffffffffu >> (31 - lane)));
    if (n_in != __popc(ends)) {  // coinciding ends (warp-uniform branch)
      lo = c0;
      for (int c = c0; c < c0 + n_in; ++c) lo += (fl >= __shfl_sync(full, pre_hi, c)) ? 1 : 0;
    }
    const int base = __shfl_sync(full, pre_hi, max(lo - 1, 0));
    L.q = fl - (lo > 0 ? base : 0);
  } else {
    int hi = T;
    lo = 0;
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (sh.pre[mid] <= fl) lo = mid; else hi = mid;
    }
    L.q = fl - sh.pre[lo];
  }
  L.chain = lo;
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

// Fixed-point scale of the per-rotor sums: 2^24 units per m/s (resolution 6e-8 m/s, 16x finer than the float32 ulp of
// a 10 m/s rotor speed; range +-128 m/s)
#define WG_FX_SCALE 16777216.f
#define WG_FX_INV (1.f / 16777216.f)

// Evaluate the queued (station row, rotor) hits of one warp: two hits per pass, 16 quadrature points each on
// 16 lanes, shuffle-reduced to the rotor average; lanes 0 and 16 add the pass's two results to the farm's per-rotor
// sums (shared-memory integer atomics: order-free, hence bit-reproducible however the tiles are distributed).
template <int TC, int TURB>
__device__ __forceinline__ void flush_hits(const float4* __restrict__ ha, const uint32_t* __restrict__ hb, int nh,
                                           int* __restrict__ acc_du, int* __restrict__ acc_dv, int* __restrict__ acc_ad,
                                           int lane, float qy, float qz, const float* __restrict__ iso, float k_m1,
                                           float k_m2) {
  const unsigned full = 0xffffffffu;
  const int half = lane >> 4;
  for (int h0 = 0; h0 < nh; h0 += 2) {
    const bool ok = h0 + half < nh;
    const int h = ok ? h0 + half : h0;
    const float4 a = ha[h];
    const uint32_t bp = hb[h], brow = bp & 0xfffffu;
    const int bj = (int)(bp >> 20);
    const float dy = a.z + qy, dz = a.w + qz;
    const float s = sqrt_fast(fmaf(dy, dy, dz * dz)) * (1.f / DR);
    const int j0 = min((int)s, WG_NR - 2);
    const float fr = s - (float)j0;
    const float u0 = lds1(brow ^ ((uint32_t)j0 << 2));
    float u1 = lds1(brow ^ ((uint32_t)(j0 + 1) << 2));
    if (j0 == WG_NR - 2) u1 = 1.f;
    float d = fmaf(fr, u0 - u1, 1.f - u0);  // (1-u0)(1-fr) + (1-u1) fr
    const bool out = s >= (float)(WG_NR - 1);
    if (out) d = 0.f;
    const bool owner = ok && (lane & 15) == 0;  // a padded second hit adds nothing
    if (TURB == 2) {
      // wake-added turbulence: k_mt = k_m1 |1 - U| + k_m2 |dU/dr| at this point, times the station's weight w U0e
      // (|a.xy| with the sign of a.x: cos g0 > 0) and the isotropic box at (rotor bj, point lane & 15)
      const float kq = out ? 0.f : fmaf(k_m2, fabsf(u1 - u0) * (1.f / DR), k_m1 * d);
      const float c = copysignf(sqrtf(fmaf(a.x, a.x, a.y * a.y)), a.x) * kq;
      const int ip = bj * WG_NQ + (lane & 15);
      float au = c * iso[ip], av = c * iso[TC * WG_NQ + ip], aw = c * iso[2 * TC * WG_NQ + ip];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        au += __shfl_xor_sync(full, au, o);
        av += __shfl_xor_sync(full, av, o);
        aw += __shfl_xor_sync(full, aw, o);
      }
      if (owner) {
        atomicAdd(&acc_ad[bj], __float2int_rn(au * (WG_FX_SCALE / WG_NQ)));
        atomicAdd(&acc_ad[TC + bj], __float2int_rn(av * (WG_FX_SCALE / WG_NQ)));
        atomicAdd(&acc_ad[2 * TC + bj], __float2int_rn(aw * (WG_FX_SCALE / WG_NQ)));
      }
    }
    d += __shfl_xor_sync(full, d, 8);
    d += __shfl_xor_sync(full, d, 4);
    d += __shfl_xor_sync(full, d, 2);
    d += __shfl_xor_sync(full, d, 1);
    if (owner) {
      d *= (WG_FX_SCALE / WG_NQ);
      atomicAdd(&acc_du[bj], __float2int_rn(a.x * d));
      atomicAdd(&acc_dv[bj], __float2int_rn(a.y * d));
    }
  }
}

#ifdef WG_TRACE
// diagnostic build (scripts/gpu_trace.sh): per-CTA start / end time and SM of the last FLOW_STEP launch
__device__ unsigned long long g_trace[8 * 65536];
__device__ unsigned long long g_phase[8 * 4 * 16384];  // SM cycles per (CTA, warp, phase): see WG_PHASE call sites
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#endif

template <int TC, int TURB>
#ifndef WG_TURB_CTAS
#define WG_TURB_CTAS 5
#endif
__global__ void __launch_bounds__(WG_NWARP * 32, TC <= 16 ? (TURB ? WG_TURB_CTAS : 6) : 4)
    wg_flow_kernel(const Dev d, const FlowArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];  // rows must be 256-byte aligned (XOR addressing)
  typedef FlowShared<TC, TURB> Shared;
  Shared& sh = *reinterpret_cast<Shared*>(smem_raw);
  const int T = d.T, P = d.P, F = d.F;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // wg_step launches the envs longest first (a.order): the CTAs that drain the grid are then the short ones
  // ... or, with a work table (wg_plan_kernel), one PART of a farm: CTA p of n streams the tiles 4 p + warp, + 4 n, ...
  // of the farm's station list; the parts meet in global fixed-point sums and the last one to arrive runs the
  // turbine epilogue (only for launches of a single flow step: nothing but the epilogue follows the tile loop)
  int bf, part = 0, nparts = 1;
  if (a.work) {
    const int2 wk = __ldg(a.work + blockIdx.x);  // consumed after the TMEM allocation below (wk.x < 0: unused entry)
    bf = wk.x; part = wk.y & 0xff; nparts = wk.y >> 8;
  } else {
    const int bi = F == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, f_blk = F == 2 ? (int)(blockIdx.x & 1) : 0;
    if (!((a.farm_mask >> f_blk) & 1)) return;  // needs no load: before the TMEM allocation
    bf = (d.b0 + (a.order ? a.order[bi] : bi)) * F + f_blk;
  }
  const int b = F == 2 ? bf >> 1 : bf, f = F == 2 ? bf & 1 : 0;  // F is 1 or 2
#ifdef WG_TRACE
  const unsigned long long t_start = gtimer();
  unsigned long long t_p1 = 0, t_p2 = 0, t_p3 = 0, t_p4 = 0;
#define WG_STAMP(v) v = gtimer()
  long long ph_last = clock64();
#define WG_PHASE(k)                                                         \
  if (lane == 0) {                                                          \
    const long long now_ = clock64();                                       \
    atomicAdd(&g_phase[((blockIdx.x & 16383) * 4 + warp) * 8 + k], (unsigned long long)(now_ - ph_last));           \
    ph_last = now_;                                                         \
  }
#define WG_TRACE_WRITE()                                                              \
  if (tid == 0 && a.mode == FLOW_STEP && blockIdx.x < 65536) {                        \
    unsigned smid;                                                                    \
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));                                 \
    unsigned long long* tr = g_trace + 8 * blockIdx.x;                                \
    tr[0] = t_start; tr[1] = gtimer(); tr[2] = smid | (nparts << 16) | (part << 24);  \
    tr[3] = ((unsigned long long)b << 32) | (unsigned)sh.pre[T];                      \
    tr[4] = t_p1; tr[5] = t_p2; tr[6] = t_p3; tr[7] = t_p4;                           \
  }
#else
#define WG_STAMP(v)
#define WG_PHASE(k)
#define WG_TRACE_WRITE()
#endif
  // TMEM first: the SM starts the next CTA of a tcgen05-allocating kernel only after this one has given up its
  // allocation permit (measured, scripts/micro/cta_launch.cu: starts on an SM are spaced by the time to the
  // relinquish -- 0.5 us when it is the first instruction, 2.9 us behind the prologue's loads).  Nothing loaded from
  // memory may be needed before this point (the work-table entry is first used below).
  if (warp == 0) tmem_alloc(&sh.tmem_base);
  if (a.pdl_trigger == 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // behind the device pool's copy kernel (which rewrites whole envs): only the launch overlaps, wait before any load
  if (a.pdl_wait == 3) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (bf < 0) {  // unused entry of the work table (CTA-uniform): give the columns back
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 0) tmem_dealloc(sh.tmem_base);
    return;
  }
  // ---- prologue: every global load of the CTA is issued before the first use of any of them (one round trip to
  // L2 / HBM instead of a chain of them: each dependent group costs 0.6 - 0.8 us while the CTA holds its slot)
  const bool is_t = tid < T;
  const int it = bf * T + (is_t ? tid : 0), jt = b * T + (is_t ? tid : 0);
  const bool take_action = a.mode == FLOW_STEP && f == 0 && a.actions != nullptr;
  const bool tab_sh = d.n_tab <= WG_TAB_CAP;
  const bool tab_mine = tab_sh && tid < d.n_tab;  // tables of up to 32 knots: one knot per thread
  const uint8_t g_mask = a.mask ? a.mask[b] : (uint8_t)1;
  const int g_spin = a.mode == FLOW_SPIN ? d.spin[b] : 0;
  const float ws = d.ws[b], wd = d.wd[b], xmax = d.xmax[b];
  const float knu1_env = d.knu1[b];  // K1 TI^0.3 of the episode (ambient term of the eddy viscosity), set at reset
  const int g_kemit = d.k_emit[b];
  int nstep = d.n_step[bf];
  float g_xr = 0.f, g_yr = 0.f, g_xs = 0.f, g_yaw = 0.f, g_der = 1.f, g_u = 0.f, g_v = 0.f, g_w = 0.f, g_pw = 0.f, g_ct = 0.f;
  float g_act = 0.f, g_act2 = 0.f, g_tws = 0.f, g_tp = 0.f, g_tct = 0.f;
  int g_ord = 0, g_head = 0, g_count = 0, g_retire = 0;
  if (is_t) {
    g_xr = d.xr[jt]; g_yr = d.yr[jt]; g_xs = d.xs_sorted[jt]; g_ord = d.ord_sorted[jt];
    g_yaw = d.yaw[it]; g_der = d.derate[it];
    g_u = d.u[it]; g_v = d.v[it]; g_w = d.w[it]; g_pw = d.power[it]; g_ct = d.ct[it];
    g_head = d.head[it]; g_count = d.count[it]; g_retire = d.retire[it];
    if (take_action) {  // issued HERE (volatile: not sunk to the first use in the epilogue), consumed behind the tile loop
      asm volatile("ld.global.f32 %0, [%1];" : "=f"(g_act) : "l"(a.actions + b * T * d.act_var + tid));
      if (d.act_var == 2) asm volatile("ld.global.f32 %0, [%1];" : "=f"(g_act2) : "l"(a.actions + b * T * 2 + T + tid));
    }
  }
  if (tab_mine) { g_tws = d.tab_ws[tid]; g_tp = d.tab_p[tid]; g_tct = d.tab_ct[tid]; }
  const int nsteps = (a.mode == FLOW_FIXED) ? a.n_fixed : (a.mode == FLOW_SPIN ? g_spin : d.S);
  if (!g_mask || nsteps <= 0) {  // nothing to do for this env (CTA-uniform): give the TMEM columns back
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 0) tmem_dealloc(sh.tmem_base);
    return;
  }

  const float dt = d.dt, R = d.R;
  const float rR = 1.f / R;
  const int k_emit = max(g_kemit, 1);
  float* __restrict__ prof = d.prof + (size_t)bf * T * P * WG_NR;
  float* __restrict__ pcon = d.pcon + (size_t)bf * T * P * 4;
  float* pmut0 = d.pmut + (size_t)bf * T * P * 4;
  float* pmut1 = pmut0 + (size_t)d.B * F * T * P * 4;
  // this warp's tile buffer (32 rows x 256 B) and this lane's row, as shared addresses
  const uint32_t tile_a = smem_u32(smem_raw + hdr_bytes<TC, TURB>()) + (uint32_t)warp * WG_TILE * WG_ROW_BYTES;
  const uint32_t row_a = tile_a + (uint32_t)lane * WG_ROW_BYTES;
  const float qy = c_qy[lane & 15], qz = c_qz[lane & 15];
  void* bar = &sh.mbar[warp];

  float der_r = 1.f;  // induction scale of turbine tid (act_var = 2 extension); lives in its thread
  int retire_r = 0;   // oldest stations of chain tid to drop at the head of the next flow step (found in this one)
  if (is_t) {
    sh.xr[tid] = g_xr;
    sh.yr[tid] = g_yr;
    sh.xs[tid] = g_xs;
    sh.ord[tid] = g_ord;
    float yaw = g_yaw;
    der_r = g_der;
    sh.yaw[tid] = yaw;
    sh.u[tid] = g_u;
    sh.v[tid] = g_v;
    sh.w[tid] = g_w;
    sh.pw[tid] = g_pw;
    sh.ct[tid] = g_ct;
    sh.head[tid] = g_head;
    sh.count[tid] = g_count;
    retire_r = g_retire;
    sh.sum_ws[tid] = sh.sum_wd[tid] = sh.sum_yaw[tid] = sh.sum_pw[tid] = 0.f;
  }
  if (tab_mine) {
    sh.tab_ws[tid] = g_tws;
    sh.tab_p[tid] = g_tp;
    sh.tab_ct[tid] = g_tct;
  }
  if (tid == 0) sh.base_sum = 0.f;
  for (int i = T + tid; i < 2 * TC; i += blockDim.x) sh.xs[i] = CUDART_INF_F;
  if (lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t phase = 0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  WG_STAMP(t_p1);
  const uint32_t taddr = sh.tmem_base + ((uint32_t)(warp * 32) << 16);
  const float* tws = tab_sh ? sh.tab_ws : d.tab_ws;
  const float* tpw = tab_sh ? sh.tab_p : d.tab_p;
  const float* tct = tab_sh ? sh.tab_ct : d.tab_ct;
  int xs_top = 1;
  while (2 * xs_top - 1 < T) xs_top <<= 1;
  const float x_retire = xmax + MARGIN_D * d.D;
  const float2 tv0 = make_float2(0.f, 0.f);
  float tb_xo = 0.f, tb_yo = 0.f, tb_zo = 0.f, tb_sc = 0.f;
  if (TURB) {
    tb_xo = d.tb_off[b * 3 + 0]; tb_yo = d.tb_off[b * 3 + 1]; tb_zo = d.tb_off[b * 3 + 2];
    tb_sc = d.tb_scale[b];
  }
  const unsigned full = 0xffffffffu;
  const unsigned lt = (1u << lane) - 1u;
  float4* ha = sh.hit_a[warp];
  uint32_t* hb = sh.hit_b[warp];

  for (int sub = 0; sub < nsteps; ++sub) {
    const float* __restrict__ pm_old = (nstep & 1) ? pmut1 : pmut0;
    float* __restrict__ pm_new = (nstep & 1) ? pmut0 : pmut1;
    // Taylor shift of the turbulence box at the old time level (particle motion) and the new one (rotor inflow)
    const float xs_t = TURB ? taylor_shift(d, ws, nstep, tb_xo, d.tb_len_x) : 0.f;
    const float xs_t1 = TURB ? taylor_shift(d, ws, nstep + 1, tb_xo, d.tb_len_x) : 0.f;
    if (TURB == 2) {  // the isotropic box at every rotor point, new time level (read by flush_hits after the barrier)
      const float xs2 = taylor_shift(d, ws, nstep + 1, tb_xo, d.tb2_len_x);
      for (int idx = tid; idx < T * WG_NQ; idx += blockDim.x) {
        const int t = idx >> 4;
        const float4 q = sample_iso(d, sh.xr[t], fmaf(c_qy[idx & 15], R, sh.yr[t]), fmaf(c_qz[idx & 15], R, d.zh), xs2,
                                    tb_yo, tb_zo);
        sh.iso[idx] = q.x; sh.iso[TC * WG_NQ + idx] = q.y; sh.iso[2 * TC * WG_NQ + idx] = q.z;
      }
    }

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
      // Retire the stations that are past the farm (+margin) after this step's move.  The oracle tests the moved
      // position of the oldest stations at the head of the step (oracle/dwm_numpy.py:step, retire loop); the same
      // test was evaluated by the lanes that marched those stations in the previous step (keep_emit below), so the
      // head of the step needs no dependent loads.
      const int c = sh.count[tid] - min(retire_r, sh.count[tid]);
      sh.count[tid] = c;
      sh.keep_emit[tid] = min(c, WG_RETIRE_CAP);
      sh.acc_du[tid] = 0; sh.acc_dv[tid] = 0;
      if (TURB == 2) { sh.acc_ad[tid] = 0; sh.acc_ad[TC + tid] = 0; sh.acc_ad[2 * TC + tid] = 0; }
    }
    if (TC > 32) __syncthreads(); else __syncwarp();
    if (warp == 0) {  // chain offsets in the flat station list: warp scan of the counts
      const int c0 = lane < T ? sh.count[lane] : 0;
      const int c1 = (TC > 32 && lane + 32 < T) ? sh.count[lane + 32] : 0;
      int i0 = c0, i1 = c1;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t0 = __shfl_up_sync(full, i0, o), t1 = __shfl_up_sync(full, i1, o);
        if (lane >= o) { i0 += t0; i1 += t1; }
      }
      const int tot0 = __shfl_sync(full, i0, 31);
      if (lane == 0) sh.pre[0] = 0;
      if (lane < T) sh.pre[lane + 1] = i0;
      if (TC > 32 && lane + 32 < T) sh.pre[lane + 33] = tot0 + i1;
    }
    __syncthreads();
    const int ntot = sh.pre[T];
    const int ntiles = (ntot + WG_TILE - 1) / WG_TILE;
    WG_STAMP(t_p2);

    // ------------------------------------------------------------------ warp-private tile pipeline
    LaneLoc Ln;
    Ln.valid = 0; Ln.chain = 0; Ln.slot = 0; Ln.q = 0;
    const int tile0 = part * WG_NWARP + warp, tstride = nparts * WG_NWARP;
    if (tile0 < ntiles) Ln = locate<TC>(sh, tile0, lane, T, P, ntot);
    WG_PHASE(0)  // outside the tile loop (prologue, step head, barriers, epilogue)
    for (int tile = tile0; tile < ntiles; tile += tstride) {
      const LaneLoc Lc = Ln;
      const Seg sg = segments(Lc, lane);
      if (lane == 0) mbar_expect_tx(bar, (uint32_t)sg.nvalid * WG_ROW_BYTES);
      __syncwarp();
      const unsigned st_c = (unsigned)(Lc.chain * P + Lc.slot);  // < T * P: 32-bit offsets inside the farm's block
      if (sg.start) bulk_g2s(row_a, prof + st_c * (unsigned)WG_NR, (uint32_t)sg.len * WG_ROW_BYTES, bar);
      // station scalars: in flight together with the tile, consumed after the mbarrier wait
      float4 pmc = make_float4(0.f, 0.f, 0.f, 0.f), pcc = pmc;
      if (Lc.valid) {
        pmc = __ldcg(reinterpret_cast<const float4*>(pm_old + st_c * 4u));
        pcc = __ldcg(reinterpret_cast<const float4*>(pcon + st_c * 4u));
      }
      // age neighbours outside this tile (older = flat index - 1 of the same chain, younger = + 1): their scalars
      // are loaded now, in flight together with the tile, and reduced to the moved position before the march (three
      // registers across the march instead of a dependent load chain in front of the bracket search)
      const int ch_o = __shfl_up_sync(full, Lc.chain, 1), ch_y = __shfl_down_sync(full, Lc.chain, 1);
      const int vy_ = __shfl_down_sync(full, Lc.valid, 1);
      const bool has_o = Lc.valid && Lc.q > 0;
      const bool has_y = Lc.valid && Lc.q < sh.count[Lc.chain] - 1;
      const bool need_o = has_o && (lane == 0 || ch_o != Lc.chain);
      const bool need_y = has_y && (lane == 31 || !vy_ || ch_y != Lc.chain);
      float4 pmx = make_float4(0.f, 0.f, 0.f, 0.f), pcx = pmx;
      if (need_o || need_y) {
        const int sx = need_o ? (Lc.slot == 0 ? P - 1 : Lc.slot - 1) : (Lc.slot == P - 1 ? 0 : Lc.slot + 1);
        const unsigned st_x = (unsigned)(Lc.chain * P + sx);
        pmx = __ldcg(reinterpret_cast<const float4*>(pm_old + st_x * 4u));
        pcx = __ldcg(reinterpret_cast<const float4*>(pcon + st_x * 4u));
      }
      // while the tile is in flight: find the next tile and pull its station scalars towards L2
      if (tile + tstride < ntiles) {
        Ln = locate<TC>(sh, tile + tstride, lane, T, P, ntot);
        if (Ln.valid) {
          const unsigned st = (unsigned)(Ln.chain * P + Ln.slot);
          // The station scalars only (one 128-byte line per 8 slots).  The profile rows are NOT prefetched any more: a
          // prefetch pulls one 128-byte line of a 256-byte row, the TMA copy then fetches the other half on its own --
          // measured 0.339 -> 0.331 ms per launch without it (0.347 with both lines prefetched); dropping the scalar
          // prefetch as well is neutral on the 4x4 farm and costs 2 % on the 8x8 one.
          if (lane == 0 || (Ln.slot & 7) == 0) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pm_old + st * 4u));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pcon + st * 4u));
          }
        }
      }
      const uint32_t rowk = row_a ^ ((uint32_t)(Lc.slot & 7) << 4);
      // Everything that needs only the station scalars -- the move of this station and of its out-of-tile
      // neighbour, their positions among the sorted rotor planes -- runs while the tile is still in flight.  With a
      // turbulence box the moves wait for gathers that miss L2; there the block runs behind the tile wait (measured:
      // 0.671 vs 0.693 ms per launch for the Mann variant).
      float xn = 0.f, yn = 0.f, zn = 0.f, dx = 0.f;
      float xe = 0.f, ye = 0.f, ze = 0.f;  // moved position of the out-of-tile neighbour (the older one if both)
      int cn = 0, ce = 0;  // #{rotor planes upstream of} this station / its out-of-tile neighbour, after the move
      auto move_and_search = [&]() {
        if (Lc.valid) {
          const float2 tvc = TURB ? sample_lp(d, pmc.x, pmc.y, pmc.z, xs_t, tb_yo, tb_zo, tb_sc) : tv0;
          moved(pmc, pcc, ws, dt, tvc, xn, yn, zn, dx);
        }
        if (need_o || need_y) {
          const float2 tvx = TURB ? sample_lp(d, pmx.x, pmx.y, pmx.z, xs_t, tb_yo, tb_zo, tb_sc) : tv0;
          float dxx;
          moved(pmx, pcx, ws, dt, tvx, xe, ye, ze, dxx);
        }
        count_below2(sh.xs, xs_top, xn, xe, cn, ce);
      };
      // (with a box, running the block before the wait -- the later tiles' bricks are L2 hits thanks to the round-ahead
      // prefetch -- was measured again with the brick layout: 0.546 vs 0.538 ms, still slower)
      const bool early = !TURB;
      if (early) move_and_search();
      if (TURB && tile == tile0 && Lc.valid) prefetch_lp(d, pmc.x, pmc.y, pmc.z, xs_t, tb_yo, tb_zo);  // later tiles: a round ahead
      WG_PHASE(1)  // tile set-up: segments, load issue, scalars, prefetches, moves, plane searches
      mbar_wait(bar, phase);
      phase ^= 1u;
      WG_PHASE(2)  // waiting for the tile
      if (!early) move_and_search();
      const float u0cg = pcc.x * pcc.z, u0sg = pcc.x * pcc.w;
      {  // warp-collective TMEM traffic: idle lanes march their (stale) row too
        const float xt_mid = (pmc.x + 0.5f * dx - sh.xr[Lc.chain]) * rR;
#ifdef WG_EXP_NOMARCH  // timing experiment only: rows pass through unchanged
        const float ucn = pmc.w + 0.f * (xt_mid + (float)rowk + (float)taddr);
#else
        const float ucn = march_row_tmem(rowk, taddr, dx * rR, xt_mid, pcc.y);
#endif
        if (Lc.valid) {
          *reinterpret_cast<float4*>(pm_new + st_c * 4u) = make_float4(xn, yn, zn, ucn);
          if (Lc.q < WG_RETIRE_CAP) {  // does this station leave the farm with the next step's move?
            float x2, y2, z2, dx2;
            moved(make_float4(xn, yn, zn, ucn), pcc, ws, dt, tv0, x2, y2, z2, dx2);
            if (!(x2 > x_retire)) atomicMin(&sh.keep_emit[Lc.chain], Lc.q);
          }
        }
      }
      WG_PHASE(3)  // move + march
      fence_async_smem();
      __syncwarp();
      if (sg.start) {  // write the marched rows back (same segments as the load)
        bulk_s2g(prof + st_c * (unsigned)WG_NR, row_a, (uint32_t)sg.len * WG_ROW_BYTES);
        bulk_commit();
      }
      // ---- superposition: which rotor planes does this station bracket together with its age neighbours?
      {
        float xo = __shfl_up_sync(full, xn, 1), yo = __shfl_up_sync(full, yn, 1), zo = __shfl_up_sync(full, zn, 1);
        float xy = __shfl_down_sync(full, xn, 1), yy = __shfl_down_sync(full, yn, 1), zy = __shfl_down_sync(full, zn, 1);
        if (need_o) { xo = xe; yo = ye; zo = ze; }  // one out-of-tile neighbour per lane (the older one if both)
        else if (need_y) { xy = xe; yy = ye; zy = ze; }
        if (need_o && need_y) {  // rare: a one-station segment needs both neighbours from outside the tile
          const unsigned sx = (unsigned)(Lc.chain * P + (Lc.slot == P - 1 ? 0 : Lc.slot + 1));
          const float4 pm = __ldcg(reinterpret_cast<const float4*>(pm_old + sx * 4u));
          const float4 pc = __ldcg(reinterpret_cast<const float4*>(pcon + sx * 4u));
          const float2 tvx = TURB ? sample_lp(d, pm.x, pm.y, pm.z, xs_t, tb_yo, tb_zo, tb_sc) : tv0;
          float dxx;
          moved(pm, pc, ws, dt, tvx, xy, yy, zy, dxx);
        }
        // Rotor planes bracketed by this station and its age neighbours: with c(x) = #{k : xs[k] < x} over the sorted
        // plane positions, plane k lies in [x1, x2) iff c(x1) <= k < c(x2).  Interval A = [self (younger end), older
        // neighbour), interval B = [younger neighbour, self (older end)); an inverted pair counts with sign -1
        // (oracle/dwm_numpy.py:353-356).  Every lane handles its own row for both of its intervals.  In-tile
        // neighbours pass their count by shuffle; the out-of-tile ones share one extra search.
        int co = __shfl_up_sync(full, cn, 1), cy = __shfl_down_sync(full, cn, 1);
        if (need_o) co = ce; else if (need_y) cy = ce;
        if (__any_sync(full, need_o && need_y)) {
          const int ce = count_below(sh.xs, xs_top, xy);
          if (need_o && need_y) cy = ce;
        }
        if (!has_o) co = cn;
        if (!has_y) cy = cn;
        const int a_lo = min(cn, co), nA = abs(cn - co), b_lo = min(cn, cy), nB = abs(cn - cy);
#ifdef WG_EXP_NOSUPER  // timing experiment only: no rotor-plane hits
        const int nmax = 0;
#else
        const int nmax = __reduce_max_sync(full, max(nA, nB));
#endif
        if (nmax > 0) {
          const float rdA = rcp_fast(xo - xn), rdB = rcp_fast(xn - xy);
          for (int it = 0; it < nmax; ++it) {
            const int kA = a_lo + it, kB = b_lo + it;
            const int jA = it < nA ? sh.ord[kA] : Lc.chain, jB = it < nB ? sh.ord[kB] : Lc.chain;
            // A wake whose centre passes the rotor centre at >= (WG_NR - 1) dr + 1 rotor radii puts every quadrature
            // point (|q| < 1) outside its profile: such a hit adds exactly zero and is dropped here.
            bool hitA = jA != Lc.chain, hitB = jB != Lc.chain;
            float4 recA, recB;
            if (hitA) {
              const float w = (sh.xs[kA] - xn) * rdA;
              const float wg = kA >= cn ? 1.f - w : w - 1.f;
              const float yc = fmaf(w, yo - yn, yn), zc = fmaf(w, zo - zn, zn);
              recA = make_float4(wg * u0cg, wg * u0sg, (sh.yr[jA] - yc) * rR, (d.zh - zc) * rR);
              hitA = fmaf(recA.z, recA.z, recA.w * recA.w) < R_CULL2;
            }
            if (hitB) {
              const float w = (sh.xs[kB] - xy) * rdB;
              const float wg = kB >= cy ? w : -w;
              const float yc = fmaf(w, yn - yy, yy), zc = fmaf(w, zn - zy, zy);
              recB = make_float4(wg * u0cg, wg * u0sg, (sh.yr[jB] - yc) * rR, (d.zh - zc) * rR);
              hitB = fmaf(recB.z, recB.z, recB.w * recB.w) < R_CULL2;
            }
            const unsigned mA = __ballot_sync(full, hitA), mB = __ballot_sync(full, hitB);
            const int nhA = __popc(mA), nhB = __popc(mB);
            if (nhA + nhB == 0) continue;
            const bool split = nhA + nhB > WG_HIT_CAP;  // rare: evaluate the two interval kinds one after the other
            if (hitA) {
              const int p = __popc(mA & lt);
              ha[p] = recA;
              hb[p] = rowk | ((uint32_t)jA << 20);
            }
            if (split) {
              __syncwarp();
              flush_hits<TC, TURB>(ha, hb, nhA, sh.acc_du, sh.acc_dv, sh.acc_ad, lane, qy, qz, sh.iso, d.k_m1, d.k_m2);
              __syncwarp();
            }
            if (hitB) {
              const int p = (split ? 0 : nhA) + __popc(mB & lt);
              ha[p] = recB;
              hb[p] = rowk | ((uint32_t)jB << 20);
            }
            const int nh = split ? nhB : nhA + nhB;
            __syncwarp();
            if (nh > 0) flush_hits<TC, TURB>(ha, hb, nh, sh.acc_du, sh.acc_dv, sh.acc_ad, lane, qy, qz, sh.iso, d.k_m1, d.k_m2);
            __syncwarp();
          }
        }
      }
      __syncwarp();
      WG_PHASE(4)  // store issue + superposition
      // Turbulence box: the next tile's wake-centre bricks are pulled towards L2 one round ahead.  The next tile's
      // positions were prefetched to L2 at the top of this round; reading them is an L2 hit that overlaps the wait for
      // the store, and the random DRAM read of the brick then has the whole tile set-up of the next round to land.
      float4 pnx = make_float4(0.f, 0.f, 0.f, 0.f);
      const bool pf_next = TURB && tile + tstride < ntiles && Ln.valid;
      if (pf_next) pnx = __ldcg(reinterpret_cast<const float4*>(pm_old + (unsigned)(Ln.chain * P + Ln.slot) * 4u));
      bulk_wait_read0();  // the store has drained its shared-memory reads: the buffer can be refilled
      if (pf_next) prefetch_lp(d, pnx.x, pnx.y, pnx.z, xs_t, tb_yo, tb_zo);
      __syncwarp();
      WG_PHASE(5)  // waiting for the store to release the buffer
    }
    WG_STAMP(t_p3);
    // multi-wave grids release the finish kernel here, behind the tile loop: it is launched while the last CTAs run
    // their epilogues (its launch latency and ring staging are off the step's critical path) without taking CTA
    // slots from the waves that still have to start (a trigger at CTA start does: measured slower)
    if (a.pdl_trigger == 2 && sub + 1 == nsteps) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // last substep
    // Programmatic dependent of the previous step's finish kernel: from here on the CTA writes what that kernel reads
    // (substep means, yaws, powers, baseline power) -- wait for it to be complete.  Everything above touched the wake
    // state only.  (Without the launch attribute, or behind a kernel that never triggers, the wait returns at once.)
    if (a.pdl_wait == 1) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (nparts > 1) {
      // Split farm: add this CTA's sums to the farm's global ones and take an arrival ticket.  Every part but the last
      // to arrive is done (its rows and station scalars are on their way to HBM); the last one collects the sums --
      // leaving the scratch zeroed for the next step -- and runs the epilogue.  The retire prefix travels as
      // WG_RETIRE_CAP - keep through atomicMax, so that the scratch's rest state is all zeros.
      __syncthreads();
      if (tid < T) {
        int* ga = d.part_acc + (size_t)bf * (5 * T);
        if (sh.acc_du[tid]) atomicAdd(ga + tid, sh.acc_du[tid]);
        if (sh.acc_dv[tid]) atomicAdd(ga + T + tid, sh.acc_dv[tid]);
        if (TURB == 2) {
#pragma unroll
          for (int k = 0; k < 3; ++k)
            if (sh.acc_ad[k * TC + tid]) atomicAdd(ga + (2 + k) * T + tid, sh.acc_ad[k * TC + tid]);
        }
        const int kp = WG_RETIRE_CAP - sh.keep_emit[tid];
        if (kp > 0) atomicMax(d.part_keep + (size_t)bf * T + tid, kp);
        __threadfence();
      }
      __syncthreads();
      if (tid == 0) sh.ticket = atomicAdd(d.part_arrive + bf, 1);
      __syncthreads();
      if (sh.ticket != nparts - 1) {  // not the last part (CTA-uniform)
        WG_TRACE_WRITE();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (warp == 0) tmem_dealloc(sh.tmem_base);
        return;
      }
      __threadfence();
      if (tid < T) {
        int* ga = d.part_acc + (size_t)bf * (5 * T);
        sh.acc_du[tid] = atomicExch(ga + tid, 0);
        sh.acc_dv[tid] = atomicExch(ga + T + tid, 0);
        if (TURB == 2) {
#pragma unroll
          for (int k = 0; k < 3; ++k) sh.acc_ad[k * TC + tid] = atomicExch(ga + (2 + k) * T + tid, 0);
        }
        sh.keep_emit[tid] = WG_RETIRE_CAP - atomicExch(d.part_keep + (size_t)bf * T + tid, 0);
      }
      if (tid == 0) d.part_arrive[bf] = 0;
    }
    if (TURB) {
      // Ambient fluctuation averaged over every rotor's quadrature points at the new time level.  (Sampling it at the
      // head of the step, together with the isotropic box, was measured slower -- 0.566 vs 0.549 ms per launch: there
      // every warp waits for the random reads before its first tile, here the warps that finish early absorb them.)
      for (int idx = tid; idx < T * WG_NQ; idx += blockDim.x) {  // T*16: half-warps stay whole
        const int t = idx >> 4;
        const float4 q = sample_raw(d, sh.xr[t], fmaf(qy, R, sh.yr[t]), fmaf(qz, R, d.zh), xs_t1, tb_yo, tb_zo, tb_sc);
        float qu = q.x, qv = q.y, qw = q.z;
        const unsigned hm = 0xffffu << (lane & 16);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
          qu += __shfl_xor_sync(hm, qu, o);
          qv += __shfl_xor_sync(hm, qv, o);
          qw += __shfl_xor_sync(hm, qw, o);
        }
        if ((lane & 15) == 0) {
          sh.tu[t] = qu * (1.f / WG_NQ); sh.tv[t] = qv * (1.f / WG_NQ); sh.tw[t] = qw * (1.f / WG_NQ);
        }
      }
    }
    // Another substep re-reads the stored rows through TMA: wait for the bulk stores to land and order the generic
    // writes (pm_new, released particles) before them.  The last substep only needs the stores' shared-memory reads
    // to be over (bulk_wait_read0 per tile); the writes complete with the kernel.
    const bool more = sub + 1 < nsteps;
    if (more) {
      bulk_wait_all0();
      fence_async_all();
    }
    __syncthreads();

    WG_STAMP(t_p4);
    // ------------------------------------------------------------------ turbine epilogue
    const bool emit = (nstep % k_emit) == 0;
    if (tid < T) {
      const float du = (float)sh.acc_du[tid] * WG_FX_INV, dv = (float)sh.acc_dv[tid] * WG_FX_INV;
      float u = ws - du + (TURB ? sh.tu[tid] : 0.f), v = dv + (TURB ? sh.tv[tid] : 0.f), w = TURB ? sh.tw[tid] : 0.f;
      if (TURB == 2) {
        u += (float)sh.acc_ad[tid] * WG_FX_INV;
        v += (float)sh.acc_ad[TC + tid] * WG_FX_INV;
        w += (float)sh.acc_ad[2 * TC + tid] * WG_FX_INV;
      }
      // _adjust_yaws (Wind_Farm_Env.py:822-864), once per env step.  The actions were requested in the prologue and
      // are first needed here, behind the tile loop: when they come from mapped host memory (wg_step_host) the PCIe
      // read latency is off every CTA's critical path.  (Nothing in the tile loop reads the yaw.)
      if (sub == 0 && take_action) {
        float yaw0 = sh.yaw[tid];
        d.old_yaw[b * T + tid] = yaw0;
        if (d.act_var == 2)  // extension: induction (derating) action, applied as set point
          der_r = fminf(fmaxf(d.derate_min + 0.5f * (g_act2 + 1.f) * (1.f - d.derate_min), d.derate_min), 1.f);
        if (d.action_method == 0) {
          yaw0 = fminf(fmaxf(yaw0 + g_act * d.yaw_step, d.yaw_min), d.yaw_max);
        } else {
          float tgt = (g_act + 1.0f) / 2.0f * (d.yaw_max - d.yaw_min) + d.yaw_min;
          tgt = fminf(fmaxf(tgt, yaw0 - d.yaw_step), yaw0 + d.yaw_step);
          yaw0 = fminf(fmaxf(tgt, d.yaw_min), d.yaw_max);
        }
        sh.yaw[tid] = yaw0;
      }
      const float yaw = sh.yaw[tid];
      float sg, cg;
      sincosf(yaw * 0.017453292519943295f, &sg, &cg);
      const float wse = u * cg;
      float pw, ct;
      tab_interp2(tws, tpw, tct, d.n_tab, wse, pw, ct);
      const float der = der_r;
      if (der != 1.f) {  // derated rotor: a = delta a_tab (actuator disc), see include/windgym_b200.h
        const float ctt = fminf(fmaxf(ct, 0.f), CT_MAX);
        const float a0 = 0.5f * (1.f - sqrtf(1.f - ctt)), a1 = der * a0;
        ct = 4.f * a1 * (1.f - a1);
        if (a0 > 0.f) pw *= (a1 * (1.f - a1) * (1.f - a1)) / (a0 * (1.f - a0) * (1.f - a0));
      }
      ct *= cg * cg;
      ct = fminf(fmaxf(ct, 0.f), CT_MAX);
      sh.u[tid] = u; sh.v[tid] = v; sh.w[tid] = w; sh.pw[tid] = pw; sh.ct[tid] = ct;
      sh.cg[tid] = cg; sh.sg[tid] = sg;
      retire_r = sh.keep_emit[tid];  // length of the chain's prefix (oldest first) that retires next step
      int slot = -1;
      if (emit) {
        sh.ind[tid] = 0.5f * (1.f - sqrtf(1.f - ct));
        slot = sh.head[tid];
        if (sh.count[tid] == P) atomicOr(&d.flags[b], 2); else sh.count[tid] += 1;
        sh.head[tid] = (slot + 1 == P) ? 0 : slot + 1;
      }
      sh.keep_emit[tid] = slot;
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
        const int slot = sh.keep_emit[t];
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
      if (more) fence_async_all();
    }
    ++nstep;
    __syncthreads();
  }

  if (tid < T) {
    d.yaw[bf * T + tid] = sh.yaw[tid];
    d.derate[bf * T + tid] = der_r;
    d.retire[bf * T + tid] = retire_r;
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
    WG_TRACE_WRITE();
    d.n_step[bf] = nstep;
    d.load[bf] = sh.pre[T];
    if (a.mode == FLOW_STEP && f == 1) d.base_pow_mean[b] = sh.base_sum / (float)nsteps;
  }
  WG_PHASE(6)  // after the tile loop: barrier wait, turbine epilogue, particle release, write-back
  if (warp == 0) {  // every warp's last TMEM access precedes the substep loop's closing barrier
    __syncwarp();
    tmem_dealloc(sh.tmem_base);
  }
}

template <int TC, int TURB>
constexpr size_t flow_smem() { return hdr_bytes<TC, TURB>() + (size_t)WG_NWARP * WG_TILE * WG_ROW_BYTES; }

template <int TC, int TURB>
static cudaError_t launch_as(const Dev& d, const FlowArgs& a, cudaStream_t s) {
  const size_t smem = flow_smem<TC, TURB>();
  // the opt-in to > 48 KB of dynamic shared memory is a per-DEVICE function attribute: one flag per device, so that a
  // process driving several GPUs configures each of them
  static bool configured[WG_MAX_DEVICES] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= WG_MAX_DEVICES || !configured[dev]) {
    cudaError_t e =
        cudaFuncSetAttribute(wg_flow_kernel<TC, TURB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < WG_MAX_DEVICES) configured[dev] = true;
  }
  const int grid = a.work ? a.n_work : d.Bg * d.F;
  if (a.pdl_wait) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(WG_NWARP * 32); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, wg_flow_kernel<TC, TURB>, d, a);
  }
  wg_flow_kernel<TC, TURB><<<grid, WG_NWARP * 32, smem, s>>>(d, a);
  return cudaGetLastError();
}

// Resident CTAs per SM x SMs, from the kernel's own resource use.  (cudaOccupancyMaxActiveBlocksPerMultiprocessor
// answers 1 for this kernel -- it seems to charge a tcgen05-allocating kernel the whole tensor memory; measured
// per-CTA timelines show 6 / 5 / 4 resident CTAs, i.e. registers and shared memory decide, as computed here.)
template <int TC, int TURB>
static int slots_as() {
  int dev = 0, sms = 0, regs_sm = 0, smem_sm = 0, smem_rsv = 0;
  cudaGetDevice(&dev);
  cudaFuncAttributes fa{};
  if (cudaFuncGetAttributes(&fa, wg_flow_kernel<TC, TURB>) != cudaSuccess) { cudaGetLastError(); return 0; }
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev);
  cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
  cudaDeviceGetAttribute(&smem_rsv, cudaDevAttrReservedSharedMemoryPerBlock, dev);
  cudaGetLastError();
  const int regs_warp = ((fa.numRegs + 7) / 8 * 8) * 32;                    // allocated per warp in units of 8 per thread
  const int by_regs = regs_sm / (regs_warp * WG_NWARP);
  const int by_smem = smem_sm / (int)(flow_smem<TC, TURB>() + fa.sharedSizeBytes + smem_rsv);
  const int by_tmem = 512 / WG_TMEM_COLS;
  const int per_sm = std::max(1, std::min(std::min(by_regs, by_smem), std::min(by_tmem, 32)));
  return per_sm * sms;
}

// CTAs of the flow kernel variant this handle launches that are resident on the device at once
int flow_resident_ctas(const Dev& d) {
  if (d.tb_raw && d.tb2_raw) return d.T <= 16 ? slots_as<16, 2>() : slots_as<WG_MAX_T, 2>();
  if (d.tb_raw) return d.T <= 16 ? slots_as<16, 1>() : slots_as<WG_MAX_T, 1>();
  return d.T <= 16 ? slots_as<16, 0>() : slots_as<WG_MAX_T, 0>();
}

#ifdef WG_TRACE
extern "C" int wg_debug_phase_read(unsigned long long* out, int reset) {
  static unsigned long long host[8 * 4 * 16384];
  cudaError_t e = cudaMemcpyFromSymbol(host, g_phase, sizeof(host));
  for (int k = 0; k < 8; ++k) out[k] = 0;
  for (size_t i = 0; i < 8 * 4 * 16384; ++i) out[i & 7] += host[i];
  if (reset) {
    void* p = nullptr;
    cudaGetSymbolAddress(&p, g_phase);
    cudaMemset(p, 0, sizeof(host));
  }
  return (int)e;
}
extern "C" int wg_debug_trace_read(unsigned long long* out, int n_cta) {
  return (int)cudaMemcpyFromSymbol(out, g_trace, sizeof(unsigned long long) * 8 * n_cta);
}
#endif

cudaError_t launch_flow(const Dev& d, const FlowArgs& a, cudaStream_t s) {
  if (d.tb_raw && d.tb2_raw) return d.T <= 16 ? launch_as<16, 2>(d, a, s) : launch_as<WG_MAX_T, 2>(d, a, s);
  if (d.tb_raw) return d.T <= 16 ? launch_as<16, 1>(d, a, s) : launch_as<WG_MAX_T, 1>(d, a, s);
  return d.T <= 16 ? launch_as<16, 0>(d, a, s) : launch_as<WG_MAX_T, 0>(d, a, s);
}

}  // namespace wg
