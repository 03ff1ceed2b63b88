// Stochastic duration predictor (reverse) tail, length regulator and prior sampling.
//
// Reference: StochasticDurationPredictor.forward(reverse=True) (models.py:108-117), ConvFlow.forward
// (modules.py:496-527), the inverse rational-quadratic spline (transforms.py:50-98,101-191) restated
// branch-free, ElementwiseAffine reverse (modules.py:408), and the length
// regulator of SynthesizerTrn.infer (models.py:702-718) with generate_path (commons.py:116-129)
// restated as a searchsorted gather instead of the [T',T] one-hot matmul.
#include "kernels.cuh"
#include "launch.cuh"

namespace pb200 {
void count_launch();

namespace {

// ---- counter-based normal noise (Philox-4x32-10 + Box-Muller).  The reference graph's two
// RandomNormalLike nodes are unseeded, so any N(0,1) stream is a valid stand-in.
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ float philox_normal(unsigned long long seed, uint32_t stream, uint32_t b, uint32_t c, uint32_t t) {
  uint32_t ctr[4] = {t, c, b, stream};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(ctr, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  const float u1 = ((float)(ctr[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0,1)
  const float u2 = ((float)(ctr[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  return sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
}

// one warp per (row, item): dot product over gin with a shuffle reduction
__global__ void __launch_bounds__(256) speaker_cond_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                                           const float* __restrict__ emb_g, const int* __restrict__ sid,
                                                           float* __restrict__ cond, int rows, int gin) {
  pdl_launch_dependents();   // programmatic dependent launch (launch.cuh): the next grid may be scheduled now;
  pdl_wait();                // nothing below runs before the previous grid has completed
  const int b = blockIdx.y;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* e = emb_g + (long long)sid[b] * gin;
  const float* wr = w + (long long)row * gin;
  float acc = 0.f;
  for (int k = lane; k < gin; k += 32) acc = fmaf(__ldg(wr + k), __ldg(e + k), acc);
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) cond[(long long)b * rows + row] = acc + __ldg(bias + row);
}

__global__ void dp_noise_kernel(View z, const float* __restrict__ eps, const long long* __restrict__ eps_off,
                                const CallParams* __restrict__ cp, const int* __restrict__ len) {
  pdl_launch_dependents();   // programmatic dependent launch (launch.cuh): the next grid may be scheduled now;
  pdl_wait();                // nothing below runs before the previous grid has completed
  const unsigned long long seed = cp->seed;
  const float noise_w = cp->noise_w;
  const int b = blockIdx.z, ch = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int T = len[b];
  if (t >= T) return;
  float e;
  if (eps != nullptr) e = eps[eps_off[b] + (long long)ch * T + t];
  else e = philox_normal(seed, 1u, (uint32_t)b, (uint32_t)ch, (uint32_t)t);
  z.p[(long long)b * z.bs + (long long)ch * z.cs + t] = e * noise_w;
}

__global__ void cf_pre_kernel(View z, int x0_ch, const float* __restrict__ w, const float* __restrict__ bias, View g,
                              View h, int C, const int* __restrict__ len) {
  pdl_launch_dependents();   // programmatic dependent launch (launch.cuh): the next grid may be scheduled now;
  pdl_wait();                // nothing below runs before the previous grid has completed
  const int b = blockIdx.z;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= len[b]) return;
  const float x0 = z.p[(long long)b * z.bs + (long long)x0_ch * z.cs + t];
  for (int c = blockIdx.y; c < C; c += gridDim.y)
    h.p[(long long)b * h.bs + (long long)c * h.cs + t] =
        fmaf(__ldg(w + c), x0, __ldg(bias + c)) + g.p[(long long)b * g.bs + (long long)c * g.cs + t];
}

constexpr int MAX_BINS = 16;

__device__ __forceinline__ float softplus_t(float x) { return x > 20.f ? x : log1pf(expf(x)); }

// One thread per (b, t).  Mirrors rational_quadratic_spline(inverse=True) with tails="linear".
__global__ void spline_inverse_kernel(View z, int x1_ch, View h, int nb, float inv_sqrt_c, float bound,
                                      const int* __restrict__ len) {
  pdl_launch_dependents();   // programmatic dependent launch (launch.cuh): the next grid may be scheduled now;
  pdl_wait();                // nothing below runs before the previous grid has completed
  const int b = blockIdx.z;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= len[b]) return;
  float* xp = z.p + (long long)b * z.bs + (long long)x1_ch * z.cs + t;
  const float x = *xp;
  if (!(x >= -bound && x <= bound)) return;  // identity outside the tails (and for NaN)
  const float* hb = h.p + (long long)b * h.bs + t;
  float cw[MAX_BINS + 1], chh[MAX_BINS + 1], der[MAX_BINS + 1];
  const float min_w = 1e-3f, min_h = 1e-3f, min_d = 1e-3f;
  // Python evaluates (1 - min_bin_width * num_bins) in double before it meets the fp32 tensor.
  const float shrink = (float)(1.0 - 1e-3 * (double)nb);
  const float span2 = 2.f * bound;
  // widths
  {
    float u[MAX_BINS], mx = -INFINITY, sum = 0.f;
    for (int k = 0; k < nb; ++k) { u[k] = hb[(long long)k * h.cs] * inv_sqrt_c; mx = fmaxf(mx, u[k]); }
    for (int k = 0; k < nb; ++k) { u[k] = expf(u[k] - mx); sum += u[k]; }
    float run = 0.f;
    cw[0] = -bound;
    for (int k = 0; k < nb; ++k) {
      const float wk = __fadd_rn(min_w, __fmul_rn(shrink, u[k] / sum));
      run += wk;
      cw[k + 1] = __fadd_rn(__fmul_rn(span2, run), -bound);   // two roundings, like torch's mul then add
    }
    cw[nb] = bound;
  }
  // heights
  {
    float u[MAX_BINS], mx = -INFINITY, sum = 0.f;
    for (int k = 0; k < nb; ++k) { u[k] = hb[(long long)(nb + k) * h.cs] * inv_sqrt_c; mx = fmaxf(mx, u[k]); }
    for (int k = 0; k < nb; ++k) { u[k] = expf(u[k] - mx); sum += u[k]; }
    float run = 0.f;
    chh[0] = -bound;
    for (int k = 0; k < nb; ++k) {
      const float hk = __fadd_rn(min_h, __fmul_rn(shrink, u[k] / sum));
      run += hk;
      chh[k + 1] = __fadd_rn(__fmul_rn(span2, run), -bound);
    }
    chh[nb] = bound;
  }
  // derivatives: interior from the network, both ends pinned so that min_d + softplus(const) == 1
  {
    const float cst = (float)log(exp(1.0 - 1e-3) - 1.0);   // np.log(np.exp(1 - min_derivative) - 1), float64 then cast
    der[0] = min_d + softplus_t(cst);
    der[nb] = der[0];
    for (int k = 1; k < nb; ++k) der[k] = min_d + softplus_t(hb[(long long)(2 * nb + k - 1) * h.cs]);
  }
  // bin = #(x >= cumheights_k) - 1 with the last edge nudged by 1e-6 (transforms.py:44-47)
  int bin = -1;
  for (int k = 0; k <= nb; ++k) {
    const float edge = k == nb ? chh[k] + 1e-6f : chh[k];
    bin += (x >= edge) ? 1 : 0;
  }
  bin = min(max(bin, 0), nb - 1);
  const float in_cw = cw[bin], in_w = cw[bin + 1] - cw[bin];
  const float in_ch = chh[bin], in_h = chh[bin + 1] - chh[bin];
  const float delta = in_h / in_w;
  const float d0 = der[bin], d1 = der[bin + 1];
  const float tt = (x - in_ch) * (d0 + d1 - 2.f * delta);
  const float qa = tt + in_h * (delta - d0);
  const float qb = in_h * d0 - tt;
  const float qc = -delta * (x - in_ch);
  const float disc = qb * qb - 4.f * qa * qc;
  const float root = (2.f * qc) / (-qb - sqrtf(disc));
  *xp = root * in_w + in_cw;
}

// One CTA per utterance: durations, inclusive scan, output length.
__global__ void __launch_bounds__(256) durations_kernel(View z, float ea_m, float ea_scale, const CallParams* __restrict__ cp,
                                                        const int* __restrict__ w_override, int w_override_pitch,
                                                        int* __restrict__ cum, int cum_pitch, int* __restrict__ y_len,
                                                        float* __restrict__ logw_out, const int* __restrict__ len) {
  pdl_launch_dependents();   // programmatic dependent launch (launch.cuh): the next grid may be scheduled now;
  pdl_wait();                // nothing below runs before the previous grid has completed
  __shared__ int warp_tot[8];
  __shared__ int carry_s;
  const float length_scale = cp->length_scale;
  const int b = blockIdx.x;
  const int T = len[b];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const float* z0 = z.p + (long long)b * z.bs;
  for (int base = 0; base < T; base += 256) {
    const int t = base + threadIdx.x;
    int wc = 0;
    if (t < T) {
      const float logw = (z0[t] - ea_m) * ea_scale;   // ElementwiseAffine reverse, channel 0 (modules.py:408)
      if (logw_out) logw_out[(long long)b * cum_pitch + t] = logw;
      const float w = expf(logw) * length_scale;      // models.py:702
      float c = ceilf(w);
      if (!(c >= 0.f)) c = 0.f;
      if (c > 1048576.f) c = 1048576.f;
      wc = (int)c;
      if (w_override) wc = min(max(w_override[(long long)b * w_override_pitch + t], 0), 1048576);
    }
    int v = wc;
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += n;
    }
    if (lane == 31) warp_tot[warp] = v;
    __syncthreads();
    // a block adds at most 256 * 2^20 = 2^28; the running total saturates at 2^30 so it can never wrap: an absurd
    // length stays monotonic and far above the engine's frame limit, where plan_back() rejects it
    long long prefix = carry_s;
    for (int w2 = 0; w2 < warp; ++w2) prefix += warp_tot[w2];
    const int tot = (int)min(prefix + v, 1LL << 30);
    if (t < T) cum[(long long)b * cum_pitch + t] = tot;
    __syncthreads();
    if (threadIdx.x == 255) carry_s = tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) y_len[b] = max(carry_s, 1);   // clamp_min(sum, 1)  (models.py:704)
}

__global__ void __launch_bounds__(256) expand_kernel(View stats, int inter, const int* __restrict__ cum, int cum_pitch,
                                                     const int* __restrict__ len, const int* __restrict__ y_len, View zp,
                                                     const float* __restrict__ eps, long long eps_bs, int eps_cs,
                                                     const CallParams* __restrict__ cp) {
  pdl_launch_dependents();   // programmatic dependent launch (launch.cuh): the next grid may be scheduled now;
  pdl_wait();                // nothing below runs before the previous grid has completed
  const unsigned long long seed = cp->seed;
  const float noise_scale = cp->noise_scale;
  const int b = blockIdx.z;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int F = y_len[b];
  if (j >= F) return;
  const int T = len[b];
  const int* cb = cum + (long long)b * cum_pitch;
  // i(j) = #{ i : cum[i] <= j }  (searchsorted right); frames past the total duration carry m = logs = 0
  int lo = 0, hi = T;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cb[mid] <= j) lo = mid + 1; else hi = mid;
  }
  const bool valid = lo < T;
  const float* sb = stats.p + (long long)b * stats.bs;
  float* zb = zp.p + (long long)b * zp.bs;
  for (int c = blockIdx.y; c < inter; c += gridDim.y) {
    float m = 0.f, logs = 0.f;
    if (valid) {
      m = sb[(long long)c * stats.cs + lo];
      logs = sb[(long long)(inter + c) * stats.cs + lo];
    }
    float e;
    if (eps != nullptr) e = eps[(long long)b * eps_bs + (long long)c * eps_cs + j];
    else e = philox_normal(seed, 2u, (uint32_t)b, (uint32_t)c, (uint32_t)j);
    zb[(long long)c * zp.cs + j] = m + e * expf(logs) * noise_scale;
  }
}

}  // namespace

void launch_speaker_cond(const float* w, const float* bias, const float* emb_g, const int* sid, float* cond, int rows,
                         int gin, int B, cudaStream_t st) {
  if (B <= 0 || rows <= 0) return;
  dim3 grid((rows + 7) / 8, B);
  launch_k(speaker_cond_kernel, dim3(grid), dim3(256), 0, st, w, bias, emb_g, sid, cond, rows, gin);
  count_launch();
}

void launch_dp_noise(View z, const float* eps, const long long* eps_off, const CallParams* cp,
                     const int* len, int B, int Tmax, cudaStream_t st) {
  if (B <= 0 || Tmax <= 0) return;
  dim3 grid((Tmax + 127) / 128, 2, B);
  launch_k(dp_noise_kernel, dim3(grid), dim3(128), 0, st, z, eps, eps_off, cp, len);
  count_launch();
}

void launch_cf_pre(View z, int x0_ch, const float* w, const float* b, View g, View h, int C, const int* len, int B,
                   int Tmax, cudaStream_t st) {
  if (B <= 0 || Tmax <= 0) return;
  dim3 grid((Tmax + 127) / 128, 8, B);
  launch_k(cf_pre_kernel, dim3(grid), dim3(128), 0, st, z, x0_ch, w, b, g, h, C, len);
  count_launch();
}

void launch_spline_inverse(View z, int x1_ch, View h, int bins, float inv_sqrt_c, float bound, const int* len, int B,
                           int Tmax, cudaStream_t st) {
  if (B <= 0 || Tmax <= 0) return;
  dim3 grid((Tmax + 63) / 64, 1, B);
  launch_k(spline_inverse_kernel, dim3(grid), dim3(64), 0, st, z, x1_ch, h, bins, inv_sqrt_c, bound, len);
  count_launch();
}

void launch_durations(View z, float ea_m, float ea_scale, const CallParams* cp, const int* w_override,
                      int w_override_pitch, int* cum, int cum_pitch, int* y_len, float* logw_out, const int* len,
                      int B, int Tmax, cudaStream_t st) {
  if (B <= 0) return;
  (void)Tmax;
  launch_k(durations_kernel, dim3(B), dim3(256), 0, st, z, ea_m, ea_scale, cp, w_override, w_override_pitch, cum, cum_pitch,
                                      y_len, logw_out, len);
  count_launch();
}

void launch_expand(View stats, int inter, const int* cum, int cum_pitch, const int* len, const int* y_len, View zp,
                   const float* eps, long long eps_bs, int eps_cs, const CallParams* cp, int B,
                   int Fmax, cudaStream_t st) {
  if (B <= 0 || Fmax <= 0) return;
  dim3 grid((Fmax + 255) / 256, 8, B);
  launch_k(expand_kernel, dim3(grid), dim3(256), 0, st, stats, inter, cum, cum_pitch, len, y_len, zp, eps, eps_bs, eps_cs, cp);
  count_launch();
}

}  // namespace pb200
