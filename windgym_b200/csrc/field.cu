// wg_flow_field_kernel -- DWMFlowSimulation.get_windspeed(view, include_wakes=True) on a set of points of one
// env's farm (render path, reference call site WindGym/Wind_Farm_Env.py:1056; view built at :470-476 as a
// 250 x 250 XYView at hub height).  Second consumer of the superposition rule of the flow kernel: for every
// point, every chain is scanned pair by pair (consecutive ages); a pair that brackets the point's x contributes
// the two stations' deficits, interpolated linearly in x, evaluated at the point's distance from the interpolated
// wake centre (oracle/dwm_numpy.py:wind_at_points).  One thread per point; the pair scan reads the same station
// in every lane (broadcast loads), the profile rows come through L2.
#include "wg_internal.cuh"

namespace wg {

__device__ __forceinline__ float row_deficit(const float* __restrict__ row, int key, float s) {
  if (s >= (float)(WG_NR - 1)) return 0.f;
  const int j0 = min((int)s, WG_NR - 2), j1 = j0 + 1;
  const float fr = s - (float)j0;
  const float u0 = row[(((j0 >> 2) ^ key) << 2) | (j0 & 3)];
  const float u1 = (j1 == WG_NR - 1) ? 1.f : row[(((j1 >> 2) ^ key) << 2) | (j1 & 3)];
  return (1.f - u0) * (1.f - fr) + (1.f - u1) * fr;
}

__global__ void __launch_bounds__(128) wg_flow_field_kernel(const Dev d, int b, int f, const float* __restrict__ px,
                                                            const float* __restrict__ py, int n, float z,
                                                            float* __restrict__ out) {
  const int T = d.T, P = d.P;
  const int bf = b * d.F + f;
  const float ws = d.ws[b], rR = 1.f / d.R;
  const int par = d.n_step[bf] & 1;
  const float* __restrict__ prof = d.prof + (size_t)bf * T * P * WG_NR;
  const float* __restrict__ pcon = d.pcon + (size_t)bf * T * P * 4;
  const float* __restrict__ pm = d.pmut + ((size_t)par * d.B * d.F + bf) * T * P * 4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float x = px[i], y = py[i];
    float du = 0.f, dv = 0.f;
    for (int t = 0; t < T; ++t) {
      const int cnt = d.count[bf * T + t], head = d.head[bf * T + t];
      if (cnt < 2) continue;
      // ages: a = 0 youngest (slot head-1) ... cnt-1 oldest
      int s0 = head - 1;
      if (s0 < 0) s0 += P;
      float4 p0 = *reinterpret_cast<const float4*>(pm + ((size_t)t * P + s0) * 4);
      for (int a = 1; a < cnt; ++a) {
        int s1 = s0 - 1;
        if (s1 < 0) s1 += P;
        const float4 p1 = *reinterpret_cast<const float4*>(pm + ((size_t)t * P + s1) * 4);
        // younger p0, older p1: up = x0 <= x < x1 (+1), dn = x1 <= x < x0 (-1)
        const bool up = p0.x <= x && x < p1.x, dn = p1.x <= x && x < p0.x;
        if (up || dn) {
          const float sgn = up ? 1.f : -1.f;
          const float w = (x - p0.x) / (p1.x - p0.x);
          const float yc = p0.y * (1.f - w) + p1.y * w, zc = p0.z * (1.f - w) + p1.z * w;
          const float ry = (y - yc) * rR, rz = (z - zc) * rR;
          const float s = sqrtf(ry * ry + rz * rz) * (1.f / DR);
          const float4 c0 = *reinterpret_cast<const float4*>(pcon + ((size_t)t * P + s0) * 4);
          const float4 c1 = *reinterpret_cast<const float4*>(pcon + ((size_t)t * P + s1) * 4);
          const float D0 = row_deficit(prof + ((size_t)t * P + s0) * WG_NR, s0 & 7, s);
          const float D1 = row_deficit(prof + ((size_t)t * P + s1) * WG_NR, s1 & 7, s);
          du += sgn * ((1.f - w) * c0.x * c0.z * D0 + w * c1.x * c1.z * D1);
          dv += sgn * ((1.f - w) * c0.x * c0.w * D0 + w * c1.x * c1.w * D1);
        }
        s0 = s1;
        p0 = p1;
      }
    }
    float4 amb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (d.tb_raw)
      amb = sample_raw(d, x, y, z, taylor_shift(d, ws, d.n_step[bf], d.tb_off[b * 3], d.tb_len_x), d.tb_off[b * 3 + 1],
                       d.tb_off[b * 3 + 2], d.tb_scale[b]);
    out[i] = ws + amb.x - du;
    out[n + i] = amb.y + dv;
    out[2 * n + i] = amb.z;
  }
}

cudaError_t launch_flow_field(const Dev& d, int b, int f, const float* px, const float* py, int n, float z, float* out,
                              cudaStream_t s) {
  const int grid = min((n + 127) / 128, 148 * 8);
  wg_flow_field_kernel<<<grid, 128, 0, s>>>(d, b, f, px, py, n, z, out);
  return cudaGetLastError();
}

// 64-byte bricks of the low-pass box: cell (i, j, k) <- its 8 periodic trilinear corners, in sample_lp's order
// ((a, b) pairs as one float4 = corners c = 0, 1).  Run once when a box is attached.
__global__ void __launch_bounds__(256) wg_brick_kernel(const float2* __restrict__ lp, float4* __restrict__ lp8, int nx,
                                                       int ny, int nz) {
  const size_t n = (size_t)nx * ny * nz;
  for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(c % nz), j = (int)((c / nz) % ny), i = (int)(c / ((size_t)nz * ny));
    const int i1 = i + 1 == nx ? 0 : i + 1, j1 = j + 1 == ny ? 0 : j + 1, k1 = k + 1 == nz ? 0 : k + 1;
    const int ii[2] = {i, i1}, jj[2] = {j, j1};
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const float2* row = lp + ((size_t)ii[a] * ny + jj[b]) * nz;
        const float2 q0 = row[k], q1 = row[k1];
        lp8[c * 4 + a * 2 + b] = make_float4(q0.x, q0.y, q1.x, q1.y);
      }
  }
}

// 128-byte bricks of the raw box: cell (i, j, k) <- its 8 periodic corners (u, v, w, 0), in gather4's loop order
__global__ void __launch_bounds__(256) wg_raw_brick_kernel(const float4* __restrict__ raw, float4* __restrict__ raw8, int nx,
                                                           int ny, int nz) {
  const size_t n = (size_t)nx * ny * nz * 8;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const size_t c = e >> 3;
    const int corner = (int)(e & 7), a = corner >> 2, b = (corner >> 1) & 1, cc = corner & 1;
    const int k = (int)(c % nz), j = (int)((c / nz) % ny), i = (int)(c / ((size_t)nz * ny));
    const int ii = a ? (i + 1 == nx ? 0 : i + 1) : i, jj = b ? (j + 1 == ny ? 0 : j + 1) : j, kk = cc ? (k + 1 == nz ? 0 : k + 1) : k;
    raw8[e] = raw[((size_t)ii * ny + jj) * nz + kk];
  }
}

cudaError_t launch_raw_bricks(const float4* raw, float4* raw8, int nx, int ny, int nz, cudaStream_t s) {
  wg_raw_brick_kernel<<<148 * 16, 256, 0, s>>>(raw, raw8, nx, ny, nz);
  return cudaGetLastError();
}

cudaError_t launch_bricks(const float2* lp, float4* lp8, int nx, int ny, int nz, cudaStream_t s) {
  wg_brick_kernel<<<148 * 16, 256, 0, s>>>(lp, lp8, nx, ny, nz);
  return cudaGetLastError();
}

}  // namespace wg
