// Internal declarations shared by the sm_100a kernels and the C-ABI host code of libwindgym_b200.
// Model constants mirror oracle/dwm_numpy.py (the frozen specification); see DESIGN.md.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/windgym_b200.h"

#define WG_NR 64          // radial nodes per wake profile
#define WG_NQ 16          // rotor quadrature points
#define WG_MAX_T 64       // turbines per farm supported by the fixed-size shared tables
#define WG_TILE 32        // stations per warp tile
#define WG_ROW_BYTES (WG_NR * 4)
#define WG_MAX_DEVICES 64  // per-device one-time kernel attribute set-up
#define WG_MAX_PARTS 8     // CTAs one farm's station list can be split over (wg_plan_kernel)

namespace wg {

struct Dev;

constexpr float DR = 1.0f / 16.0f;
constexpr float K_HILL = 0.4f;
constexpr float K1 = 0.023f;
constexpr float K2 = 0.016f;
constexpr float CT_MAX = 0.96f;
constexpr float MARGIN_D = 2.0f;
constexpr float DXT_MIN = 1e-6f;

enum FlowMode { FLOW_FIXED = 0, FLOW_SPIN = 1, FLOW_STEP = 2 };

// finish-kernel work flags
enum FinishFlags {
  FIN_PUSH_MES = 1, FIN_PUSH_FP = 2, FIN_PUSH_BP = 4, FIN_OBS = 8, FIN_REWARD = 16, FIN_MEAS_FROM_ARGS = 32
};

// One output element of the observation vector (built on the host from wg_mes_config).
struct ObsDesc {
  int32_t kind;   // 0 current, 1 window mean, 2 TI of one ring, 3 mean of turbine TIs (scaled individually)
  int32_t ring;   // ring id (kind 0,1,2)
  int32_t chan;   // 0 ws, 1 wd, 2 yaw, 3 power
  int32_t win;    // window index i (kind 1)
  float lo, span; // scaling: 2*(v-lo)/span-1, span = float32(double(hi)-double(lo))
  int32_t H, N, W; // history_length / history_N / window_length of the channel (copied here: no dynamic indexing
                   // of the kernel-parameter arrays in the observation loop)
  int32_t off;     // offset of the ring inside the env's ring block (kind 0, 1, 2)
};

struct Dev {
  // dims
  int B, T, F, P, S, n_tab;
  int Bg;  // envs [b0, b0 + Bg) are launched: the whole allocation B, the active prefix for wg_step (wg_set_active),
  int b0;  //   or the spare range of a background refill (wg_pool_refill)
  float dt, D, R, zh, d_particle;
  float yaw_min, yaw_max, yaw_step;
  int action_method, base_controller;
  int act_var;            // 1: yaw actions; 2: + one induction (derating) action per turbine (extension)
  float derate_min;
  int power_reward, power_avg, pen_type;
  float power_scaling, action_penalty;
  // mes
  int n_rings, ring_floats, obs_dim, obs_rows;  // obs_rows = 1 (single agent) or T (multi agent)
  int ch_cur[4], ch_roll[4], ch_N[4], ch_H[4], ch_W[4];
  int noise;
  int fin_lean;           // no noise, no TI observations, no Power_diff reward: the lean finish-kernel variant applies
  float noise_std[4];
  unsigned long long noise_seed;
  float ti_lo, ti_span;
  int ch_base[4];         // offset of channel c's first turbine ring (ring of turbine t: ch_base[c] + t ch_H[c])
  int farm_off[3];        // offsets of the farm-level rings ws, wd, power
  const int* ring_off;    // [n_rings]
  const int* ring_chan;   // [n_rings]
  const ObsDesc* obs_desc;  // [obs_rows * obs_dim]
  // tables
  const float* tab_ws; const float* tab_p; const float* tab_ct;
  const double* x_pos; const double* y_pos;
  // particle state
  float* prof;   // [B,F,T,P,64]   16-byte chunks XOR-swizzled with (slot & 7)
  float* pmut;   // [2,B,F,T,P,4]  x, y, z, uc ; ping-pong on the farm's step parity
  float* pcon;   // [B,F,T,P,4]    U0e, knu1, cos g0, sin g0
  int* head; int* count;  // [B,F,T]  ring buffer per chain; count includes the stations about to retire:
  int* retire;            // [B,F,T]  oldest stations of the chain that the next flow step drops before it moves
  int* n_step;            // [B,F]
  int* load;              // [B,F] live stations of the farm after its last flow step (work estimate of its CTA)
  int* order;             // [B]   launch order of wg_step: envs [0, Bg) sorted by descending load (see wg_order_kernel)
  // split farms (FlowArgs::work): scratch the parts of a farm meet in; all zeros between launches
  int* part_acc;          // [B,F,5,T] fixed-point rotor sums du, dv (+ wake-added u, v, w)
  int* part_keep;         // [B,F,T]   WG_RETIRE_CAP - (retire prefix of the chain)
  int* part_arrive;       // [B,F]     arrival counter
  int2* work;             // [n_work]  work table of wg_step: (farm b*F+f | -1, part | nparts << 8), longest part first
  float *yaw, *u, *v, *w, *power, *ct;  // [B,F,T]
  float *derate;          // [B,F,T] induction scale delta in [derate_min, 1] (1 = the reference's turbine)
  // env state
  float *ws, *ti, *wd, *rated, *xmax;   // [B]
  float *knu1;            // [B] K1 TI^0.3: ambient-turbulence term of the eddy viscosity (emission scalar of new particles)
  int *k_emit, *time_max, *timestep, *flags, *n_push, *n_fp, *n_bp, *spin;  // [B]
  float *xr, *yr;         // [B,T]
  float *xs_sorted;       // [B,T] rotor-plane x ascending (ties by turbine index); int *ord_sorted: turbine of each
  int *ord_sorted;        // [B,T]
  float *meas;            // [B,4,T] substep means ws, wd, yaw, power
  float *base_pow_mean;   // [B]
  float *old_yaw;         // [B,T]
  float *rings;           // [B,ring_floats]
  float *fp_ring, *bp_ring;  // [B,power_avg]
  float *fp_ring_lo;         // [B,power_avg] float32 remainder of the farm power sums (value = fp_ring + fp_ring_lo): the
                             // Power_diff reward is a DIFFERENCE of window means of ~1e7 W values (Wind_Farm_Env.py:904-918)
  // ambient turbulence box shared by all envs of the handle (null: uniform inflow); per-env offset and scale
  const float4* tb_raw;   // [Nx,Ny,Nz] (u, v, w, 0)
  const float2* tb_lp;    // [Nx,Ny,Nz] (v, w) low-pass filtered in y, z: moves the wake centres
  const float4* tb_lp8;   // [Nx,Ny,Nz][4] the same field as 64-byte BRICKS: cell (i,j,k) holds its 8 trilinear corners
                          // (i+a, j+b, k+c), periodic -- one aligned 64-byte read per sample instead of 8 scattered
                          // sectors; built by the library at wg_set_turbulence (null: gather from tb_lp)
  const float4* tb_raw8;  // [Nx,Ny,Nz][8] the raw box as 128-byte bricks (8 corners x (u, v, w, 0)): one line per rotor-point
                          // sample; null: gather from tb_raw
  int tb_n[3];
  float tb_inv_d[3], tb_inv_n[3], tb_len_x;
  float *tb_off;          // [B,3] position of the env inside the box [m]
  float *tb_scale;        // [B]   scale_TI factor
  // wake-added turbulence: isotropic unit-variance box (u, v, w, 0), advected with the ambient one; null = off
  const float4* tb2_raw;
  int tb2_n[3];
  float tb2_inv_d[3], tb2_inv_n[3], tb2_len_x;
  float k_m1, k_m2;       // k_mt(r) = k_m1 |1 - U| + k_m2 |dU/dr|
};

struct FlowArgs {
  int mode;             // FlowMode
  int n_fixed;          // FLOW_FIXED: steps for every env
  const uint8_t* mask;  // optional env mask
  const float* actions; // FLOW_STEP: [B,T] or null (reset fill)
  int farm_mask;        // bit f set: farm f advances
  int controller_on;    // baseline farm applies its greedy controller each substep
  const int* order;     // optional permutation of [0, Bg): CTA group i works on env order[i] (longest first)
  const int2* work;     // optional work table (single-step launches only): CTA i works on part work[i]; replaces order
  int n_work;           // entries of the table = CTAs of the launch
  int pdl_trigger;      // 1: every CTA releases the dependent launch (the step's finish kernel) at its start: the batch
                        // leaves CTA slots free, so the finish grid's launch and staging overlap the flow grid's tail;
                        // 2: ... behind its tile loop (multi-wave grids)
  int pdl_wait;         // 1: launched as programmatic dependent of the PREVIOUS step's finish kernel: prologue and tile
                        // loop (wake state only: nothing the finish kernel touches) overlap it; griddepcontrol.wait
                        // in front of the turbine epilogue (substep means, yaws, powers: what the finish kernel reads);
                        // 3: ... of the device pool's copy kernel: griddepcontrol.wait before the first load
};

// how wg_plan_kernel cuts the farms of a step into CTAs
struct PlanArgs {
  int n_work;      // table entries (= CTAs launched); unused ones are marked -1
  int slots;       // resident CTAs of the flow kernel on this device
  int target;      // > 0: cut the farms into about `target` parts of equal size (one or two full waves of CTAs)
  int max_tiles;   // > 0: no part longer than this many tiles (large farms: many short CTAs pack and drain better)
  int tail_units;  // more farms than slots: the lightest tail_units farms (launched last) are cut into
  int tail_parts;  //   tail_parts parts each, so that the grid drains on short CTAs (1: no tail split)
};

struct FinishArgs {
  int flags;
  const uint8_t* mask;
  const float* in_ws; const float* in_wd; const float* in_yaw; const float* in_power;  // FIN_MEAS_FROM_ARGS
  float* obs; float* reward; uint8_t* truncated;
  // wg_step_host with pinned host buffers: the results are ALSO stored straight into mapped host memory (no copy
  // engine in the step) and the last warp to finish publishes the step's sequence number for the polling host
  float* obs_h; float* reward_h; uint8_t* truncated_h;
  unsigned* done_count;          // device counter of finished warps (back to 0 when the flag is written)
  volatile unsigned* done_flag;  // mapped host word
  unsigned seq;
  int pdl;                       // launched as programmatic dependent of the flow kernel (griddepcontrol.wait inside)
  int trigger;                   // releases the NEXT step's flow kernel (FlowArgs::pdl_wait) once the flow results are in
};

// Device-side spare pool (wg_pool_*): status of every env slot
enum PoolStatus { POOL_ACTIVE = 0, POOL_NEED = 1, POOL_REFILLING = 2, POOL_READY = 3, POOL_PENDING = 4 };
#define WG_POOL_MAX_SWAP 64     // swaps per step (more finished episodes wait for the next step)
#define WG_POOL_MASKS 8         // refills in flight (one mask row each)

struct PoolDev {
  int* status;          // [B]   PoolStatus
  int* gen;             // [B]   refills of the slot so far (RNG counter)
  int* swap;            // [2 * WG_POOL_MAX_SWAP + 1]  src[], dst[], n of the current step
  unsigned long long* stats;  // [8] swapped, deferred (finished episodes that found no ready spare), refilled
  uint8_t* masks;       // [WG_POOL_MASKS, B]
  volatile int* need_host;  // mapped host word: spares waiting for a refill, published by every swap (host polls it
                            // without synchronising to decide when a refill is worth its ~60 launches); may be null
  // reset arguments of the slots being refilled (the arrays wg_reset takes, drawn on the device)
  float *ws, *ti, *ti_flow, *wd, *yaw0, *rated, *tb_off, *tb_scale;
  int *k_emit, *t_dev, *time_max;
  int n_active, B;
};

struct PoolDraw {       // condition sampling of wg_pool_refill (EnvConfig on the host side)
  double ws_min, ws_max, ti_min, ti_max, wd_min, wd_max, yaw_start, n_passthrough;
  double tb_len[3], tb_inv_std;   // turbulence box extent [m] and 1 / std(u_box); tb_inv_std = 0: no box
  float yaw_const;
  int yaw_random, eval_mode;
  unsigned long long seed;
};

struct ResetDevArgs {
  const uint8_t* mask;
  const float* ws; const float* ti; const float* wd; const float* yaw0; const float* rated;
  const int* k_emit; const int* t_dev; const int* time_max;
  const float* tb_off; const float* tb_scale;
};

#ifdef __CUDACC__
// Periodic trilinear sample of the frozen turbulence box at box coordinates (X, Y, Z) in cells.
struct BoxIdx {
  int i[2], j[2], k[2];
  float fx, fy, fz;
};
// cell index and fraction of a periodic coordinate without integer division: X - n floor(X / n), then floor
__device__ __forceinline__ void wrap_cell(float X, int n, float inv_n, int& i0, int& i1, float& fr) {
  const float Xw = fmaf(-(float)n, floorf(X * inv_n), X);
  const float f0 = floorf(Xw);
  fr = Xw - f0;
  i0 = min(max((int)f0, 0), n - 1);  // rounding at the seam can land on -0 / n
  i1 = i0 + 1 == n ? 0 : i0 + 1;
}
__device__ __forceinline__ BoxIdx box_index(const int* n, const float* inv_n, float X, float Y, float Z) {
  BoxIdx b;
  wrap_cell(X, n[0], inv_n[0], b.i[0], b.i[1], b.fx);
  wrap_cell(Y, n[1], inv_n[1], b.j[0], b.j[1], b.fy);
  wrap_cell(Z, n[2], inv_n[2], b.k[0], b.k[1], b.fz);
  return b;
}
__device__ __forceinline__ BoxIdx box_index(const Dev& d, float X, float Y, float Z) {
  return box_index(d.tb_n, d.tb_inv_n, X, Y, Z);
}
// trilinear (u, v, w) of a float4 box at cell coordinates
__device__ __forceinline__ float4 gather4(const float4* __restrict__ raw, const int* n, const BoxIdx& b) {
  float u = 0.f, v = 0.f, w = 0.f;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int bb = 0; bb < 2; ++bb) {
      const float wxy = (a ? b.fx : 1.f - b.fx) * (bb ? b.fy : 1.f - b.fy);
      const float4* rowp = raw + ((size_t)b.i[a] * n[1] + b.j[bb]) * n[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float wt = wxy * (c ? b.fz : 1.f - b.fz);
        const float4 q = __ldg(rowp + b.k[c]);
        u = fmaf(wt, q.x, u);
        v = fmaf(wt, q.y, v);
        w = fmaf(wt, q.z, w);
      }
    }
  return make_float4(u, v, w, 0.f);
}
// low-pass (v, w) at a wake centre; xs = Taylor shift U t - x_off (so that the box x is x - xs)
__device__ __forceinline__ float2 sample_lp(const Dev& d, float x, float y, float z, float xs, float yo, float zo,
                                            float scale) {
  const BoxIdx b = box_index(d, (x - xs) * d.tb_inv_d[0], (y + yo) * d.tb_inv_d[1], (z + zo) * d.tb_inv_d[2]);
  float v = 0.f, w = 0.f;
  if (d.tb_lp8) {  // brick of cell (i0, j0, k0): corners in the order of the loops below, two per float4
    const float4* p = d.tb_lp8 + (((size_t)b.i[0] * d.tb_n[1] + b.j[0]) * d.tb_n[2] + b.k[0]) * 4;
    const float4 q[4] = {__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3)};
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int bb = 0; bb < 2; ++bb) {
        const float wxy = (a ? b.fx : 1.f - b.fx) * (bb ? b.fy : 1.f - b.fy);
        const float4 c01 = q[a * 2 + bb];
        const float w0 = wxy * (1.f - b.fz), w1 = wxy * b.fz;
        v = fmaf(w0, c01.x, v); w = fmaf(w0, c01.y, w);
        v = fmaf(w1, c01.z, v); w = fmaf(w1, c01.w, w);
      }
    return make_float2(v * scale, w * scale);
  }
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int bb = 0; bb < 2; ++bb) {
      const float wxy = (a ? b.fx : 1.f - b.fx) * (bb ? b.fy : 1.f - b.fy);
      const float2* rowp = d.tb_lp + ((size_t)b.i[a] * d.tb_n[1] + b.j[bb]) * d.tb_n[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float wt = wxy * (c ? b.fz : 1.f - b.fz);
        const float2 q = __ldg(rowp + b.k[c]);
        v = fmaf(wt, q.x, v);
        w = fmaf(wt, q.y, w);
      }
    }
  return make_float2(v * scale, w * scale);
}
// pull the brick a wake-centre sample will read towards L2 (issued as soon as the station's position is known, consumed
// after the tile wait: the DRAM latency of the random read hides behind the tile's flight)
__device__ __forceinline__ void prefetch_lp(const Dev& d, float x, float y, float z, float xs, float yo, float zo) {
  if (!d.tb_lp8) return;
  const BoxIdx b = box_index(d, (x - xs) * d.tb_inv_d[0], (y + yo) * d.tb_inv_d[1], (z + zo) * d.tb_inv_d[2]);
  asm volatile("prefetch.global.L2 [%0];" ::"l"(d.tb_lp8 + (((size_t)b.i[0] * d.tb_n[1] + b.j[0]) * d.tb_n[2] + b.k[0]) * 4));
}
__device__ __forceinline__ float4 sample_raw(const Dev& d, float x, float y, float z, float xs, float yo, float zo,
                                             float scale) {
  const BoxIdx b = box_index(d, (x - xs) * d.tb_inv_d[0], (y + yo) * d.tb_inv_d[1], (z + zo) * d.tb_inv_d[2]);
  if (d.tb_raw8) {  // brick of cell (i0, j0, k0): the 8 corners in gather4's loop order, same arithmetic -> same bits
    const float4* p = d.tb_raw8 + (((size_t)b.i[0] * d.tb_n[1] + b.j[0]) * d.tb_n[2] + b.k[0]) * 8;
    float u = 0.f, v = 0.f, w = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int bb = 0; bb < 2; ++bb) {
        const float wxy = (a ? b.fx : 1.f - b.fx) * (bb ? b.fy : 1.f - b.fy);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float wt = wxy * (c ? b.fz : 1.f - b.fz);
          const float4 q = __ldg(p + (a * 4 + bb * 2 + c));
          u = fmaf(wt, q.x, u);
          v = fmaf(wt, q.y, v);
          w = fmaf(wt, q.z, w);
        }
      }
    return make_float4(u * scale, v * scale, w * scale, 0.f);
  }
  const float4 q = gather4(d.tb_raw, d.tb_n, b);
  return make_float4(q.x * scale, q.y * scale, q.z * scale, 0.f);
}
// unit-variance isotropic box of the wake-added turbulence (same offsets, its own periodic length)
__device__ __forceinline__ float4 sample_iso(const Dev& d, float x, float y, float z, float xs2, float yo, float zo) {
  const BoxIdx b = box_index(d.tb2_n, d.tb2_inv_n, (x - xs2) * d.tb2_inv_d[0], (y + yo) * d.tb2_inv_d[1],
                             (z + zo) * d.tb2_inv_d[2]);
  return gather4(d.tb2_raw, d.tb2_n, b);
}
// Taylor shift of the box at flow time t = n dt, reduced modulo the box length in double so that the float
// coordinate stays small: box x = x - (U t - x_off)
__device__ __forceinline__ float taylor_shift(const Dev& d, float ws, int n_step, float x_off, float len_x) {
  const double s = fmod((double)ws * (double)n_step * (double)d.dt - (double)x_off, (double)len_x);
  return (float)s;
}
#endif

void set_rotor_points(const float* qy, const float* qz);
cudaError_t launch_flow(const Dev& d, const FlowArgs& a, cudaStream_t s);
cudaError_t launch_finish(const Dev& d, const FinishArgs& a, cudaStream_t s);
cudaError_t launch_order(const Dev& d, cudaStream_t s);
cudaError_t launch_plan(const Dev& d, const PlanArgs& p, cudaStream_t s);
cudaError_t launch_pool_claim(const Dev& d, const PoolDev& p, const PoolDraw& w, int mask_row, cudaStream_t s);
cudaError_t launch_pool_publish(const PoolDev& p, int mask_row, cudaStream_t s);
cudaError_t launch_pool_swap(const Dev& d, const PoolDev& p, const uint8_t* truncated, uint8_t* swapped, cudaStream_t s);
cudaError_t launch_bricks(const float2* lp, float4* lp8, int nx, int ny, int nz, cudaStream_t s);
cudaError_t launch_raw_bricks(const float4* raw, float4* raw8, int nx, int ny, int nz, cudaStream_t s);
int flow_resident_ctas(const Dev& d);
cudaError_t launch_reset_init(const Dev& d, const ResetDevArgs& a, cudaStream_t s);
// one state field as the env-copy kernel sees it: n_rep blocks of B envs, per_env bytes each
struct CopyField {
  unsigned long long offset, rep_stride;
  unsigned per_env, n_rep;
};
cudaError_t launch_copy_envs(unsigned char* state, const CopyField* fields, int n_fields, const int* src, const int* dst,
                             int n, cudaStream_t s);
cudaError_t launch_pool_copy(unsigned char* state, const CopyField* fields, int n_fields, const PoolDev& p, float* obs,
                             float* final_obs, int obs_floats, cudaStream_t s);
cudaError_t launch_flow_field(const Dev& d, int b, int f, const float* px, const float* py, int n, float z, float* out,
                              cudaStream_t s);

}  // namespace wg
