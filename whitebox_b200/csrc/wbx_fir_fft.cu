// wbx_fir_fft.cu — convolution reverb by uniformly partitioned overlap-save (BASELINE cfg 5; extension, see wbx.h).
//
// y[n, s] = sum_{k < L} h[k] * x[n - k, s]. The impulse response is cut into NP partitions of P taps, the signal into
// windows of N = 2P frames that advance by P; with X_q = FFT_N(window q) and H_p = FFT_N(partition p, zero-padded),
//     Y_b = sum_{p < NP} H_p (.) X_{b-p},        y[bP .. bP+P) = the last P samples of IFFT_N(Y_b)
// — O(log) work per output sample instead of O(L): at 65536 taps 80 x fewer multiply-adds than the direct form the
// tensor-core path (wbx_fir_tc.cu) evaluates, and they are f32 FMAs, so no split-precision products are needed either.
//
// h is real, so the convolution is a real-linear map: the two bus channels of a track ride through the transforms as ONE
// complex signal z = L + iR (trackbuf's interleaved (L, R) frames are read and written as complex numbers as they are),
// with full N-point complex transforms and no real-FFT packing / unpacking pass. A mono bus is a signal with zero
// imaginary part.
//
// Three kernels per render (+ one per impulse response):
//   fft_windows_kernel  one CTA per (window q, track): Stockham radix-4 FFT in shared memory -> Z[q][track][f]
//   fft_mac_kernel      W[b][track][f] = sum_p H[p][f] * Z[b - p][track][f]: one thread per (f, track, 16 blocks b), the
//                       Z values of consecutive b at consecutive p form a sliding window kept in registers: one 8-byte
//                       load of Z and one of H per 64 FMAs
//   ifft_blocks_kernel  one CTA per (block b, track): inverse FFT, the last P samples -> trackbuf (the mix kernel's input)
// History: the time-domain tail [n_tracks][2][L-1] shared with the other two reverb paths (fir_gather / fir_save), so the
// paths can be switched between renders and a render may have any length; the history windows are re-transformed each
// render (for a 64-callback render that triples the forward-transform count, which is ~1/4 of the stage).
//
// Accuracy: f32 throughout, twiddles rounded from f64; measured against the f64 spec in tests/ (1e-5 of the block peak is
// the bound north_star states; this path sits near 1e-6 like the tensor-core path).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "wbx_device.cuh"

namespace wbx {

namespace {

constexpr int FFT_THREADS = 256;
constexpr int MAC_BG = 16;  // output blocks per thread of the partition sum (sliding window length)

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y)); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(__fmaf_rn(a.x, b.x, -__fmul_rn(a.y, b.y)), __fmaf_rn(a.x, b.y, __fmul_rn(a.y, b.x)));
}
// acc += h * z
__device__ __forceinline__ void cmac(float2& acc, float2 h, float2 z) {
  acc.x = __fmaf_rn(h.x, z.x, acc.x);
  acc.x = __fmaf_rn(-h.y, z.y, acc.x);
  acc.y = __fmaf_rn(h.x, z.y, acc.y);
  acc.y = __fmaf_rn(h.y, z.x, acc.y);
}

// N-point complex FFT (N = 4^S) of the data in buf0, Stockham autosort radix 4, ping-pong between buf0 and buf1; returns
// the buffer holding the result in natural order. tw[k] = exp(-2 pi i k / N). INV conjugates the twiddles (no 1/N).
template <int N, bool INV>
__device__ __forceinline__ float2* fft_shared(float2* buf0, float2* buf1, const float2* __restrict__ tw, int tid) {
  float2* in = buf0;
  float2* out = buf1;
#pragma unroll 1
  for (int Ns = 1; Ns < N; Ns *= 4) {
    const int tstep = N / (Ns * 4);
#pragma unroll
    for (int u = 0; u < N / 4 / FFT_THREADS; u++) {
      const int j = tid + u * FFT_THREADS;
      const int k = j & (Ns - 1);
      float2 v0 = in[j], v1 = in[j + N / 4], v2 = in[j + N / 2], v3 = in[j + 3 * N / 4];
      if (Ns > 1) {
        float2 w1 = __ldg(tw + k * tstep), w2 = __ldg(tw + 2 * k * tstep), w3 = __ldg(tw + 3 * k * tstep);
        if (INV) w1.y = -w1.y, w2.y = -w2.y, w3.y = -w3.y;
        v1 = cmul(v1, w1);
        v2 = cmul(v2, w2);
        v3 = cmul(v3, w3);
      }
      const float2 a = cadd(v0, v2), b = csub(v0, v2), c = cadd(v1, v3);
      float2 d = csub(v1, v3);
      d = INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);  // * (+-i)
      const int j0 = ((j - k) << 2) + k;
      out[j0] = cadd(a, c);
      out[j0 + Ns] = cadd(b, d);
      out[j0 + 2 * Ns] = csub(a, c);
      out[j0 + 3 * Ns] = csub(b, d);
    }
    __syncthreads();
    float2* t = in;
    in = out;
    out = t;
  }
  return in;
}

// Z[q][e][f] = FFT of the frames [(q - NP) P, (q - NP + 2) P) of track e's signal (frame 0 = the first of this render),
// read from the planar gather buffer xin[e * C + c][H + T] (H history frames first); frames outside are zero.
template <int N>
__global__ void __launch_bounds__(FFT_THREADS) fft_windows_kernel(const DFx* __restrict__ fx, uint32_t C, uint64_t H, uint64_t T,
                                                                  uint32_t NP, const float* __restrict__ xin,
                                                                  const float2* __restrict__ tw, float2* __restrict__ Z,
                                                                  uint32_t n_fx) {
  extern __shared__ __align__(16) float2 fsm[];
  const uint32_t q = blockIdx.x, e = blockIdx.y;
  if (!fx[e].reverb_on) return;
  const int tid = threadIdx.x;
  constexpr int P = N / 2;
  const int64_t i0 = ((int64_t)q - (int64_t)NP) * P + (int64_t)H;  // index into the gather buffer of window sample 0
  const float* xl = xin + (size_t)(e * C) * (H + T);
  const float* xr = xl + (H + T);
  const int64_t len = (int64_t)(H + T);
  for (int n = tid; n < N; n += FFT_THREADS) {
    const int64_t i = i0 + n;
    float2 v = make_float2(0.0f, 0.0f);
    if (i >= 0 && i < len) {
      v.x = __ldg(xl + i);
      if (C == 2) v.y = __ldg(xr + i);
    }
    fsm[n] = v;
  }
  __syncthreads();
  const float2* r = fft_shared<N, false>(fsm, fsm + N, tw, tid);
  float2* dst = Z + ((size_t)q * n_fx + e) * N;
  for (int n = tid; n < N; n += FFT_THREADS) dst[n] = r[n];
}

// Hs[p][f] = FFT of partition p of the impulse response (P taps, zero-padded to N) * 1/N (the inverse transform's scale)
template <int N>
__global__ void __launch_bounds__(FFT_THREADS) fft_ir_kernel(const float* __restrict__ ir, uint32_t L, const float2* __restrict__ tw,
                                                             float2* __restrict__ Hs) {
  extern __shared__ __align__(16) float2 fsm[];
  const uint32_t p = blockIdx.x;
  const int tid = threadIdx.x;
  constexpr int P = N / 2;
  for (int n = tid; n < N; n += FFT_THREADS) {
    const uint64_t k = (uint64_t)p * P + n;
    fsm[n] = make_float2((n < P && k < L) ? __ldg(ir + k) : 0.0f, 0.0f);
  }
  __syncthreads();
  const float2* r = fft_shared<N, false>(fsm, fsm + N, tw, tid);
  const float s = 1.0f / (float)N;  // a power of two: exact
  for (int n = tid; n < N; n += FFT_THREADS) Hs[(size_t)p * N + n] = make_float2(r[n].x * s, r[n].y * s);
}

// W[b][e][f] = sum_{p < NP} Hs[p][f] * Z[b - p + NP - 1 + 1 ...] — with the window numbering of fft_windows_kernel, output
// block b (frames [bP, bP + P)) takes the window ending at (b - p + 1) P, which is window q = b - p + NP - 1.
// Thread = (f, e, group of MAC_BG blocks): at step p it needs q = Q0 - p + j for its blocks j = 0 .. MAC_BG-1 — a window
// that slides down by one per step, kept in registers (slot (p - j) mod MAC_BG, static under the unroll).
template <int N>
__global__ void __launch_bounds__(128) fft_mac_kernel(const DFx* __restrict__ fx, uint32_t n_fx, uint32_t NP, uint32_t NB,
                                                      const float2* __restrict__ Hs, const float2* __restrict__ Z,
                                                      float2* __restrict__ W) {
  const uint32_t e = blockIdx.y;
  if (!fx[e].reverb_on) return;
  const uint32_t f = blockIdx.x * 128 + threadIdx.x;
  const uint32_t b0 = blockIdx.z * MAC_BG;
  const int64_t NQ = (int64_t)NB + NP - 1;
  const int64_t Q0 = (int64_t)b0 + NP - 1;
  const size_t qstride = (size_t)n_fx * N;
  const float2* zp = Z + (size_t)e * N + f;
  float2 win[MAC_BG], acc[MAC_BG];
#pragma unroll
  for (int j = 0; j < MAC_BG; j++) {
    acc[j] = make_float2(0.0f, 0.0f);
    win[j] = make_float2(0.0f, 0.0f);
  }
#pragma unroll
  for (int j = 1; j < MAC_BG; j++) {  // q = Q0 + j sits in slot (0 - j) mod MAC_BG when the loop starts
    const int64_t q = Q0 + j;
    if (q < NQ) win[(MAC_BG - j) % MAC_BG] = __ldg(zp + (size_t)q * qstride);
  }
  for (uint32_t p0 = 0; p0 < NP; p0 += MAC_BG) {
#pragma unroll
    for (int pp = 0; pp < MAC_BG; pp++) {
      const uint32_t p = p0 + pp;
      const int64_t q = Q0 - (int64_t)p;
      float2 h = make_float2(0.0f, 0.0f), z = make_float2(0.0f, 0.0f);
      if (p < NP) {
        h = __ldg(Hs + (size_t)p * N + f);
        if (q >= 0 && q < NQ) z = __ldg(zp + (size_t)q * qstride);
      }
      win[pp] = z;
#pragma unroll
      for (int j = 0; j < MAC_BG; j++) cmac(acc[j], h, win[(pp - j + MAC_BG) % MAC_BG]);
    }
  }
#pragma unroll
  for (int j = 0; j < MAC_BG; j++)
    if (b0 + j < NB) W[((size_t)(b0 + j) * n_fx + e) * N + f] = acc[j];
}

// frames [bP, bP + P) of track e = the last P samples of IFFT(W[b][e]) -> trackbuf[e][frame] = (L, R)
template <int N>
__global__ void __launch_bounds__(FFT_THREADS) ifft_blocks_kernel(const DFx* __restrict__ fx, uint32_t n_fx, uint32_t C, uint64_t T,
                                                                  const float2* __restrict__ W, const float2* __restrict__ tw,
                                                                  float* __restrict__ trackbuf, uint64_t tbs) {
  extern __shared__ __align__(16) float2 fsm[];
  const uint32_t b = blockIdx.x, e = blockIdx.y;
  if (!fx[e].reverb_on) return;
  const int tid = threadIdx.x;
  constexpr int P = N / 2;
  const float2* src = W + ((size_t)b * n_fx + e) * N;
  for (int n = tid; n < N; n += FFT_THREADS) fsm[n] = src[n];
  __syncthreads();
  const float2* r = fft_shared<N, true>(fsm, fsm + N, tw, tid);
  float2* dst = reinterpret_cast<float2*>(trackbuf) + (size_t)e * tbs;
  for (int n = tid; n < P; n += FFT_THREADS) {
    const uint64_t t = (uint64_t)b * P + n;
    if (t < T) {
      const float2 v = r[P + n];
      if (C == 2)
        dst[t] = v;
      else
        dst[t].x = v.x;
    }
  }
}

template <int N>
cudaError_t set_smem(const void* k) {
  return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * N * (int)sizeof(float2));
}

}  // namespace

// partition size: 2048 taps (4096-point transforms) unless WBX_FFT_P=512 asks for 1024-point ones; read when an impulse
// response is set and carried with the engine's reverb mode from there on
uint32_t fir_fft_partition() {
  const char* env = getenv("WBX_FFT_P");
  return (env && atoi(env) == 512) ? 512u : 2048u;
}

// bytes of the twiddle table tw[N] followed by the partition spectra
size_t fir_fft_ir_bytes(uint32_t L, uint32_t P) {
  const uint32_t N = 2 * P, NP = (L + P - 1) / P;
  return (size_t)N * sizeof(float2) + (size_t)NP * N * sizeof(float2);
}

size_t fir_fft_scratch_bytes(uint64_t T, uint32_t L, uint32_t n_fx, uint32_t P) {
  const uint32_t N = 2 * P, NP = (L + P - 1) / P;
  const uint64_t NB = (T + P - 1) / P, NQ = NB + NP - 1;
  return (size_t)(NQ + NB) * n_fx * N * sizeof(float2);
}

// twiddles (f64 -> f32 on the host, copied synchronously) + the partition spectra of the impulse response
cudaError_t launch_fir_fft_prepare(const float* ir, uint32_t L, void* ir_spectra, uint32_t P, cudaStream_t stream) {
  const uint32_t N = 2 * P, NP = (L + P - 1) / P;
  std::vector<float2> tw(N);
  for (uint32_t k = 0; k < N; k++) {
    const double a = -2.0 * 3.14159265358979323846 * (double)k / (double)N;
    tw[k] = make_float2((float)cos(a), (float)sin(a));
  }
  cudaError_t err = cudaStreamSynchronize(stream);
  if (err != cudaSuccess) return err;
  err = cudaMemcpy(ir_spectra, tw.data(), (size_t)N * sizeof(float2), cudaMemcpyHostToDevice);
  if (err != cudaSuccess) return err;
  const float2* twd = reinterpret_cast<const float2*>(ir_spectra);
  float2* Hs = reinterpret_cast<float2*>(ir_spectra) + N;
  if (N == 1024) {
    if ((err = set_smem<1024>((const void*)fft_ir_kernel<1024>)) != cudaSuccess) return err;
    fft_ir_kernel<1024><<<NP, FFT_THREADS, 2 * 1024 * sizeof(float2), stream>>>(ir, L, twd, Hs);
  } else {
    if ((err = set_smem<4096>((const void*)fft_ir_kernel<4096>)) != cudaSuccess) return err;
    fft_ir_kernel<4096><<<NP, FFT_THREADS, 2 * 4096 * sizeof(float2), stream>>>(ir, L, twd, Hs);
  }
  return cudaGetLastError();
}

template <int N>
static cudaError_t launch_fir_fft_t(const DFx* fx, uint32_t n_fx, uint32_t C, uint64_t H, uint64_t T, uint32_t L,
                                    const void* ir_spectra, const float* xin, void* scratch, float* trackbuf, uint64_t tbs,
                                    cudaStream_t stream) {
  constexpr uint32_t P = N / 2;
  const uint32_t NP = (L + P - 1) / P;
  const uint32_t NB = (uint32_t)((T + P - 1) / P), NQ = NB + NP - 1;
  const float2* tw = reinterpret_cast<const float2*>(ir_spectra);
  const float2* Hs = tw + N;
  float2* Z = reinterpret_cast<float2*>(scratch);
  float2* W = Z + (size_t)NQ * n_fx * N;
  const size_t smem = 2 * N * sizeof(float2);
  cudaError_t err;
  if ((err = set_smem<N>((const void*)fft_windows_kernel<N>)) != cudaSuccess) return err;
  if ((err = set_smem<N>((const void*)ifft_blocks_kernel<N>)) != cudaSuccess) return err;
  fft_windows_kernel<N><<<dim3(NQ, n_fx), FFT_THREADS, smem, stream>>>(fx, C, H, T, NP, xin, tw, Z, n_fx);
  fft_mac_kernel<N><<<dim3(N / 128, n_fx, (NB + MAC_BG - 1) / MAC_BG), 128, 0, stream>>>(fx, n_fx, NP, NB, Hs, Z, W);
  ifft_blocks_kernel<N><<<dim3(NB, n_fx), FFT_THREADS, smem, stream>>>(fx, n_fx, C, T, W, tw, trackbuf, tbs);
  return cudaGetLastError();
}

cudaError_t launch_fir_fft(const DFx* fx, uint32_t n_fx, uint32_t C, uint64_t H, uint64_t T, uint32_t L, const void* ir_spectra,
                           const float* xin, void* scratch, float* trackbuf, uint64_t tbs, uint32_t P, cudaStream_t stream) {
  if (P == 512)
    return launch_fir_fft_t<1024>(fx, n_fx, C, H, T, L, ir_spectra, xin, scratch, trackbuf, tbs, stream);
  return launch_fir_fft_t<4096>(fx, n_fx, C, H, T, L, ir_spectra, xin, scratch, trackbuf, tbs, stream);
}

}  // namespace wbx
