/*
 * windgym_b200.h -- C-ABI of the B200-native batched wind-farm environment hot path.
 *
 * Drop-in boundary for the per-step hot path of DTUWindEnergy/WindGym.  The reference is pure Python and has
 * no FFI of its own; every entry point below cites the reference interface it replaces.
 * All pointers named "device" are CUDA device pointers owned by the caller (torch tensors in the Python host
 * layer); the library allocates nothing per step, never synchronises the host, and is not re-entrant per
 * handle (one handle per GPU / stream).  Every function returns 0 on success or a negative wg_status; the
 * message of the last failure on the calling thread is available from wg_last_error().
 */
#ifndef WINDGYM_B200_H
#define WINDGYM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WG_VERSION 100

typedef enum {
  WG_OK = 0,
  WG_ERR_INVALID = -1,   /* bad argument / configuration (reference raises ValueError) */
  WG_ERR_UNSUPPORTED = -2, /* reference raises NotImplementedError (e.g. ActionMethod "absolute") */
  WG_ERR_CUDA = -3,      /* CUDA runtime failure, message has the cudaError string */
  WG_ERR_NO_DEVICE = -4  /* no sm_100 device: the product path never falls back to the CPU */
} wg_status;

/* One scalar history of MesClass.Mes (WindGym/MesClass.py:34-52). */
typedef struct {
  int32_t current;        /* return the latest sample                      */
  int32_t rolling_mean;   /* return history_N window means                 */
  int32_t history_N;
  int32_t history_length; /* deque maxlen                                  */
  int32_t window_length;
} wg_mes_channel;

/* farm_mes constructor arguments (WindGym/MesClass.py:360-401, built at Wind_Farm_Env.py:409-451). */
typedef struct {
  wg_mes_channel ws, wd, yaw, power;
  int32_t turb_ws, turb_wd, turb_TI, turb_power, farm_ws, farm_wd, farm_TI, farm_power;
  /* scaling ranges of _scale_val (MesClass.py:324-326); double so that float32(hi - lo) rounds like numpy */
  double ws_min, ws_max, wd_min, wd_max, yaw_min, yaw_max, ti_min, ti_max, power_max;
  int32_t noise;          /* 0 "None", 1 "Normal" (MesClass.py:441-444)     */
  float noise_std[4];     /* ws, wd, yaw, power (MesClass.py:436-439)       */
  uint64_t noise_seed;    /* the reference's noise RNG is unseeded; ours is counter based */
  int32_t multi_agent;    /* 0: WindFarmEnv obs vector (MesClass.py:679-703);
                             1: WindFarmEnvMulti per-agent rows (WindEnvMulti.py:79-103) */
} wg_mes_config;

/* Everything WindFarmEnv.__init__ / load_config fixes for the lifetime of an env (Wind_Farm_Env.py:50-261,:349-399). */
typedef struct {
  int32_t n_envs;         /* B: independent farm instances in the batch      */
  int32_t n_turb;         /* T = nx*ny                                       */
  int32_t n_farms;        /* 1, or 2 when Baseline_comp (Wind_Farm_Env.py:217-220) */
  int32_t p_cap;          /* wake-particle slots per turbine chain (multiple of 8) */
  int32_t substeps;       /* sim_steps_per_env_step = dt_env/dt_sim (:105)   */
  float dt;               /* dt_sim [s]                                      */
  float diameter;         /* turbine.diameter()  (:244)                      */
  float hub_height;       /* turbine.hub_height() (:475)                     */
  float d_particle;       /* 0.2 (:116)                                      */
  int32_t n_tab;          /* power/CT table knots (py_wake PowerCtTabular)   */
  const float* tab_ws;    /* host pointers, copied by wg_create              */
  const float* tab_power; /* [W]                                             */
  const float* tab_ct;
  const double* x_pos;    /* host, layout frame, [T] (:246-252)              */
  const double* y_pos;
  int32_t action_method;  /* 0 "yaw", 1 "wind" (:828-858); "absolute" -> WG_ERR_UNSUPPORTED (:861) */
  float yaw_min, yaw_max, yaw_step;
  int32_t base_controller; /* 0 "Local", 1 "Global" (BasicControllers.py:10,:49) */
  int32_t power_reward;   /* 0 "None", 1 "Baseline", 2 "Power_avg", 3 "Power_diff" (:173-194) */
  int32_t power_avg;      /* deque maxlen (:142-143)                         */
  float power_scaling;
  float action_penalty;
  int32_t action_penalty_type; /* 0 "Change", 1 "Total" (:804-820)           */
  int32_t steps_on_reset; /* (:229-240)                                      */
  wg_mes_config mes;
  /* Extension (no reference counterpart: act_var = 1 there, TODO at Wind_Farm_Env.py:43,:97-99; BASELINE.json cfg 4
   * "yaw + induction actions"): act_var = 2 appends one derating action per turbine, actions [B, 2T] =
   * [yaw actions | induction actions]; u in [-1, 1] sets the induction scale delta = derate_min + (u+1)/2 (1 - derate_min):
   * a = delta a_tab, CT = 4a(1-a) cos^2(yaw), P = P_tab a(1-a)^2 / (a_tab (1-a_tab)^2).  act_var = 0 means 1. */
  int32_t act_var;
  float derate_min;
} wg_config;

/* Per-env inputs of WindFarmEnv.reset (Wind_Farm_Env.py:680-802).  All device pointers.  The integer fields are
 * computed by the host in fp64 exactly as the reference does (:727-732), so that no discrete decision depends
 * on device rounding. */
typedef struct {
  const uint8_t* mask;       /* [B] 1 = reset this env; NULL = all                      */
  const float* ws;           /* [B] _set_windconditions (:557-585)                      */
  const float* ti_flow;      /* [B] TI seen by the flow solver (0 for turbtype "None")  */
  const float* wd;           /* [B]                                                     */
  const float* yaw0;         /* [B,T] initial yaw offsets (:715-720)                    */
  const float* rated_power;  /* [B] turbine.power(ws) (:700)                            */
  const int32_t* k_emit;     /* [B] particle release cadence in flow steps              */
  const int32_t* t_developed;/* [B] int(2*dist/ws)/dt spin-up flow steps (:729,:734)    */
  const int32_t* time_max;   /* [B] int(t_inflow*n_passthrough) (:732); 9999999 for FarmEval */
  const float* tb_offset;    /* [B,3] env position inside the shared turbulence box [m]; NULL without a box */
  const float* tb_scale;     /* [B] scale_TI(TI, U) factor TI*U/std(u_box) (:617,:638,:657); NULL without a box */
} wg_reset_args;

typedef struct wg_handle wg_handle;

/* WindFarmEnv.__init__: copies the immutable tables to the current CUDA device. */
int wg_create(const wg_config* cfg, wg_handle** out);
void wg_destroy(wg_handle* h);
const char* wg_last_error(void);
int wg_version(void);

/* Size of the caller-owned state buffer (one per handle; zero-initialised by the caller). */
int wg_state_bytes(const wg_handle* h, size_t* out);
/* observed_variables() (MesClass.py:610-618); per agent when mes.multi_agent. */
int wg_obs_dim(const wg_handle* h, int32_t* out);
/* Introspection of the state layout so the host can expose fs.windTurbines.{yaw,power(),rotor_avg_windspeed},
 * fs.time ... as tensor views (AgentEval.py:193-209).  dtype: 0 = f32, 1 = i32.  Returns WG_ERR_INVALID for an
 * unknown name; names are listed by wg_state_field_name(i). */
int wg_state_field(const wg_handle* h, const char* name, size_t* offset, int32_t* dtype, int32_t* ndim,
                   int64_t shape[8]);
const char* wg_state_field_name(const wg_handle* h, int32_t index);

/* WindFarmEnv.reset for the masked envs: spin-up fs.run(t_developed), measurement fill, first observation.
 * obs: device [B, obs_dim] (single agent) or [B, T, obs_dim] (multi agent); rows of unmasked envs untouched. */
int wg_reset(wg_handle* h, void* state, const wg_reset_args* args, float* obs, void* cuda_stream);

/* WindFarmEnv.step (Wind_Farm_Env.py:920-1034) for all envs.  actions: device [B,T] in [-1,1].
 * reward: device [B]; truncated: device [B] (terminated is always False in the reference, :1029).
 * Asynchronous: the kernels are enqueued on cuda_stream and the results are complete, in stream order, when the
 * step's last kernel is.  Consecutive steps on one stream overlap where their data allow (programmatic dependent
 * launch: the next step's wake advection runs beside this step's observation kernel); results do not depend on it. */
int wg_step(wg_handle* h, void* state, const float* actions, float* obs, float* reward, uint8_t* truncated,
            void* cuda_stream);

/* WindFarmEnv.step with HOST buffers -- the call a host-side user of the reference makes (numpy action in, numpy
 * obs / reward / truncated out, Wind_Farm_Env.py:920-1034), in one entry point.  actions_host: float32
 * [n_active, T*act_var]; out_host: packed results obs f32 [B, obs] | reward f32 [B] | truncated u8 [B] with
 * B = n_envs of the handle, out_bytes = wg_result_bytes.  actions_dev / out_dev: caller-owned device buffers of the
 * same sizes (out_dev always receives the results too: device-side consumers keep working).
 *  - both host buffers pinned (cudaHostAlloc / cudaHostRegister, e.g. torch pin_memory()): ZERO-COPY -- the flow
 *    kernel reads the actions from the mapped host buffer, the finish kernel stores the results into out_host and
 *    publishes the step's sequence number in a mapped word the call polls; no copy engine, no stream synchronise;
 *  - pageable buffers: H2D copy into actions_dev, wg_step, D2H copy of out_dev, cudaStreamSynchronize.
 * Either way the call returns when out_host holds the step's results (the only entry point besides
 * wg_profile_read that waits for the device). */
int wg_result_bytes(const wg_handle* h, size_t* out);
int wg_step_host(wg_handle* h, void* state, const float* actions_host, float* actions_dev, void* out_dev,
                 void* out_host, size_t out_bytes, void* cuda_stream);

/* DWMFlowSimulation.step()/run() alone (dynamiks seam, call sites Wind_Farm_Env.py:734,:745,:945): advance the
 * flow of every env by n_steps with the current yaws; no measurement bookkeeping. */
int wg_flow_steps(wg_handle* h, void* state, int32_t n_steps, void* cuda_stream);

/* farm_mes.add_measurements + get_measurements(scaled=True) + clip (MesClass.py:568-591,:679-703;
 * Wind_Farm_Env.py:513-520) on the state's ring buffers.  ws/wd/yaw/power: device [B,T]. */
int wg_mes_push_extract(wg_handle* h, void* state, const float* ws, const float* wd, const float* yaw,
                        const float* power, float* obs, void* cuda_stream);

/* Auto-reset without a stall (the reference resets inside step(): Wind_Farm_Env.py:1003-1025 tears the episode
 * down, callers such as SB3's VecEnv reset immediately; reset = t_developed + fill flow steps, :722-766).
 * wg_set_active: wg_step advances only the first n_active envs of the allocation; the remaining slots are a pool
 * of spare envs the caller spins up in the background (wg_reset with a mask, on another stream).
 * wg_copy_envs: copy the complete per-env state of slot src[k] into slot dst[k] (src, dst: device int32 [n]) --
 * swap a pre-developed spare env in for an env whose episode just ended. */
int wg_set_active(wg_handle* h, int32_t n_active);
/* Fraction (0, 1] of the device's resident CTA slots wg_step plans its single-wave decomposition for (default 1).  A
 * caller that keeps other work resident next to the stepping grid -- the background spin-ups of a spare pool hold one
 * CTA per env for ~20 ms -- lowers it so that a small batch, which wg_step cuts into parts that fill the machine,
 * still fits ONE wave of what is left. */
int wg_set_slot_share(wg_handle* h, float share);
int wg_copy_envs(wg_handle* h, void* state, const int32_t* src, const int32_t* dst, int32_t n, void* cuda_stream);

/* Device-side auto-reset: the host is not in the loop.  wg_pool_init makes the slots [n_active, n_envs) a pool of
 * spare envs (wg_step advances the first n_active, like wg_set_active).  Per step, on the stepping stream:
 * wg_pool_swap pairs the envs whose episode just ended (truncated[b] != 0, as written by wg_step) with READY spares
 * and copies the spares' complete state over them -- swapped[b] = 1 marks the envs that start a new episode in this
 * step (their obs row becomes the spare's reset observation, the finished episode's last observation is kept in
 * final_obs when given); an episode that finds no ready spare runs on and is swapped in a later step.  Every few
 * steps, on a background stream: wg_pool_refill takes the spares consumed so far, draws their wind conditions on
 * the device (counter-based generator keyed by seed, slot and refill count; the reference draws ws, ti, wd, yaw in
 * that order from np_random, Wind_Farm_Env.py:557-568,:715), runs WindFarmEnv.reset's spin-up + measurement fill on
 * them (:722-766) and marks them READY.  mask_row in [0, 8): one per refill in flight (a row may be reused once the
 * stream that carried its previous refill has drained it, e.g. row = stream index).  wg_pool_stats (synchronises):
 * out[0] swaps, out[1] finished episodes deferred for lack of a ready spare (summed over steps), out[2] refills,
 * out[3] spares waiting for a refill at the latest swap. */
typedef struct {
  double ws_min, ws_max, ti_min, ti_max, wd_min, wd_max; /* wind: ws_min ... (YAML "wind" section)              */
  double yaw_start;        /* yaw_init "Random": uniform(-yaw_start, yaw_start) (:715)                          */
  double n_passthrough;    /* time_max = int(t_inflow * n_passthrough) (:732)                                   */
  double tb_std_u;         /* std(u) of the attached turbulence box (scale_TI, :617); ignored without a box     */
  float yaw_const;         /* initial yaw offset when not random ("Zeros": 0)                                   */
  int32_t yaw_random;
  int32_t eval_mode;       /* FarmEval: time_max = 9999999                                                      */
  uint64_t seed;
} wg_pool_draw;
int wg_pool_init(wg_handle* h, void* state, int32_t n_active, void* cuda_stream);
int wg_pool_refill(wg_handle* h, void* state, const wg_pool_draw* draw, float* obs, int32_t mask_row, void* cuda_stream);
int wg_pool_swap(wg_handle* h, void* state, const uint8_t* truncated, float* obs, uint8_t* swapped, float* final_obs,
                 void* cuda_stream);
int wg_pool_stats(wg_handle* h, void* state, uint64_t out[8], void* cuda_stream);
/* Spares waiting for a refill as of the latest wg_pool_swap the device has finished (a word in mapped host memory the
 * swap kernel publishes; reading it does not synchronise; -1: not available).  The caller uses it to issue
 * wg_pool_refill when enough spares have been consumed instead of on a fixed step count. */
int wg_pool_need(wg_handle* h, int32_t* out);

/* TurbulenceFieldSite over a MannTurbulenceField (_def_site, Wind_Farm_Env.py:598-678): attach one periodic
 * turbulence box, shared read-only by every env of the handle, in the two layouts the flow kernel samples:
 * raw_uvw0 device [nx,ny,nz,4] f32 (u, v, w, 0) and lp_vw device [nx,ny,nz,2] f32 ((v, w) low-pass filtered in
 * y, z -- the scales that meander the wake centres).  Caller-owned, must outlive the handle's launches.
 * Both NULL: back to uniform inflow (turbtype "None").  Takes effect at the next wg_reset. */
int wg_set_turbulence(wg_handle* h, const float* raw_uvw0, const float* lp_vw, int32_t nx, int32_t ny, int32_t nz,
                      float dx, float dy, float dz);

/* addedTurbulenceModel = SynchronizedAutoScalingIsotropicMannTurbulence() (Wind_Farm_Env.py:618,:639,:658): a
 * unit-variance isotropic box iso_uvw0 (device [nx,ny,nz,4] f32), advected with the ambient box; inside wakes the
 * rotor inflow gains k_mt(r) * U0 * iso with k_mt = k_m1 |1 - U(r)| + k_m2 |dU/dr| (Madsen et al. 2010).  Needs
 * wg_set_turbulence first; NULL switches it off. */
int wg_set_added_turbulence(wg_handle* h, const float* iso_uvw0, int32_t nx, int32_t ny, int32_t nz, float dx, float dy,
                            float dz, float k_m1, float k_m2);

/* DWMFlowSimulation.get_windspeed(view, include_wakes=True) (render path, Wind_Farm_Env.py:1056; view :470-476):
 * wake-superposed (u, v, w) of farm `farm` of env `env` at n_points points (x[i], y[i], z) of the wind-aligned
 * frame of positions_xyz.  x, y: device [n_points]; out_uvw: device [3, n_points]. */
int wg_flow_field(wg_handle* h, void* state, int32_t env, int32_t farm, const float* x, const float* y,
                  int32_t n_points, float z, float* out_uvw, void* cuda_stream);

/* Number of kernel launches issued through this handle so far (bench.py "gpu_launches"). */
int wg_launch_count(const wg_handle* h, uint64_t* out);

/* Measurement aid (bench.py "roofline"): when enabled, wg_step records CUDA events on its stream before the flow
 * kernel, between the two kernels and after the finish kernel.  wg_profile_read waits for the recorded steps,
 * returns the summed device time of each kernel [ms] and the number of steps, and clears the record.
 * This is the only entry point that synchronises the host. */
int wg_profile_enable(wg_handle* h, int32_t on);
int wg_profile_read(wg_handle* h, double* flow_ms, double* finish_ms, uint64_t* n_steps);

#ifdef __cplusplus
}
#endif
#endif /* WINDGYM_B200_H */
