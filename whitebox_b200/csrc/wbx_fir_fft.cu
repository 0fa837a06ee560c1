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
// Kernels per render (+ one per impulse response and partition size):
//   fft_windows_kernel  one CTA per (window q, track): in-place radix-16 FFT in shared memory -> Z[slot(q)][track][f];
//                       reads the time-domain history ring and this render's chain output (trackbuf) directly
//   fft_mac_kernel      W[b][track][f] = sum_p H[p][f] * Z[b - p][track][f]: one thread per (f, track, 16 blocks b), the
//                       Z values of consecutive b at consecutive p form a sliding window kept in registers: one 8-byte
//                       load of Z and one of H per 64 FMAs
//   ifft_blocks_kernel  one CTA per (block b, track): inverse FFT, the last P samples -> trackbuf (the mix kernel's input)
// State. The canonical state is the time-domain tail [n_tracks][2][L-1] shared with the other two reverb paths (a ring:
// every render appends its min(T, L-1) newest frames), so the paths can be switched between renders and a render may
// have any length. On top of it the window spectra Z live in a ring over q that persists across renders: when the
// previous render was a whole number m of partitions long and nothing else changed (same chains, response, partition
// size), this render's history windows ARE the previous render's windows q + m, so only the NB windows that contain new
// frames are transformed (for 64 callbacks under a 65536-tap response that is 16 of 47). Otherwise all windows are
// rebuilt from the time-domain tail.
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


__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y)); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(__fmaf_rn(a.x, b.x, -__fmul_rn(a.y, b.y)), __fmaf_rn(a.x, b.y, __fmul_rn(a.y, b.x)));
}
template <bool INV>
__device__ __forceinline__ float2 cmulc(float2 a, float2 w) {  // a * w (forward) or a * conj(w) (inverse)
  if (INV) w.y = -w.y;
  return cmul(a, w);
}
// acc += h * z
__device__ __forceinline__ void cmac(float2& acc, float2 h, float2 z) {
  acc.x = __fmaf_rn(h.x, z.x, acc.x);
  acc.x = __fmaf_rn(-h.y, z.y, acc.x);
  acc.y = __fmaf_rn(h.x, z.y, acc.y);
  acc.y = __fmaf_rn(h.y, z.x, acc.y);
}

// 4-point DFT in place, outputs in natural order (INV: the conjugate transform)
template <bool INV>
__device__ __forceinline__ void dft4(float2& x0, float2& x1, float2& x2, float2& x3) {
  const float2 a = cadd(x0, x2), b = csub(x0, x2), c = cadd(x1, x3);
  float2 d = csub(x1, x3);
  d = INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);  // * (+-i)
  x0 = cadd(a, c);
  x1 = cadd(b, d);
  x2 = csub(a, c);
  x3 = csub(b, d);
}

// 16-point DFT of v[0..15] in registers: index r = 4a + b, output k = c + 4d:
//   W16^((4a+b)(c+4d)) = W4^(ac) W16^(bc) W4^(bd)  ->  DFT4 over a, twiddle W16^(bc), DFT4 over b
template <bool INV>
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
#pragma unroll
  for (int b = 0; b < 4; b++) dft4<INV>(v[b], v[4 + b], v[8 + b], v[12 + b]);  // v[4c + b] = y_b[c]
  constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, R2 = 0.70710678118654752f;
  // W16^m = (cos(m pi/8), -sin(m pi/8)), m = b * c
  v[4 * 1 + 1] = cmulc<INV>(v[4 * 1 + 1], make_float2(C1, -S1));   // m = 1
  v[4 * 1 + 2] = cmulc<INV>(v[4 * 1 + 2], make_float2(R2, -R2));   // m = 2
  v[4 * 1 + 3] = cmulc<INV>(v[4 * 1 + 3], make_float2(S1, -C1));   // m = 3
  v[4 * 2 + 1] = cmulc<INV>(v[4 * 2 + 1], make_float2(R2, -R2));   // m = 2
  v[4 * 2 + 2] = cmulc<INV>(v[4 * 2 + 2], make_float2(0.0f, -1.0f));  // m = 4
  v[4 * 2 + 3] = cmulc<INV>(v[4 * 2 + 3], make_float2(-R2, -R2));  // m = 6
  v[4 * 3 + 1] = cmulc<INV>(v[4 * 3 + 1], make_float2(S1, -C1));   // m = 3
  v[4 * 3 + 2] = cmulc<INV>(v[4 * 3 + 2], make_float2(-R2, -R2));  // m = 6
  v[4 * 3 + 3] = cmulc<INV>(v[4 * 3 + 3], make_float2(-C1, S1));   // m = 9
#pragma unroll
  for (int c = 0; c < 4; c++) dft4<INV>(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);  // v[4c + d] = X[c + 4d]
}

// shared-memory index with one pad element per 16: the stride-16 accesses of the radix-16 passes hit distinct banks
__device__ __forceinline__ int pad16(int i) { return i + (i >> 4); }
template <int N>
struct FftShape {
  static constexpr int THREADS = N / 16;
  static constexpr int SMEM = (N + N / 16) * (int)sizeof(float2);
};

// Twiddle tables, one per pass, laid out [r - 1][thread] so that a warp's loads are one contiguous run:
//   radix-16 pass with sub-transform length Ns (> 1): T[r - 1][tid] = exp(-2 pi i r k / (16 Ns)), k = tid mod Ns, r = 1..15
//   1024 only, final radix-4 pass (Ns = 256):          T[r - 1][j]   = exp(-2 pi i r j / 1024),   j < 256,        r = 1..3
template <int N>
struct FftTables {
  static constexpr int TH = N / 16;
  static constexpr int PASS = 15 * TH;                              // float2 per radix-16 pass table
  static constexpr int N16 = N == 4096 ? 2 : 1;                     // radix-16 passes with twiddles (Ns = 16, 256)
  static constexpr int TOTAL = N16 * PASS + (N == 1024 ? 3 * 256 : 0);
};

// N-point complex FFT (N = 1024 or 4096) of buf (padded layout, natural order in and out), in place, N/16 threads:
// Stockham-ordered radix-16 passes (each thread: 16 loads, twiddles, a 16-point DFT in registers, 16 stores); 1024 ends
// with a radix-4 pass. INV conjugates every twiddle (no 1/N).
template <int N, bool INV>
__device__ __forceinline__ void fft_inplace(float2* buf, const float2* __restrict__ tw, int tid) {
  constexpr int TH = N / 16;
  int pass = 0;
#pragma unroll 1
  for (int Ns = 1; Ns * 16 <= N; Ns *= 16) {
    const int k = tid & (Ns - 1);
    float2 v[16];
#pragma unroll
    for (int r = 0; r < 16; r++) v[r] = buf[pad16(tid + r * TH)];
    if (Ns > 1) {
      const float2* t = tw + (pass - 1) * FftTables<N>::PASS + tid;
#pragma unroll
      for (int r = 1; r < 16; r++) v[r] = cmulc<INV>(v[r], __ldg(t + (r - 1) * TH));
    }
    pass++;
    dft16<INV>(v);
    __syncthreads();
    const int j0 = ((tid - k) << 4) + k;
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
      for (int d = 0; d < 4; d++) buf[pad16(j0 + (c + 4 * d) * Ns)] = v[4 * c + d];
    __syncthreads();
  }
  if (N == 1024) {  // 1024 = 16 * 16 * 4: the last pass is radix 4 (Ns = 256), four butterflies per thread
    constexpr int Ns = 256;
    const float2* t4 = tw + FftTables<N>::N16 * FftTables<N>::PASS;
    float2 v[4][4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int j = tid + u * TH;  // j < 256 = Ns: k = j
#pragma unroll
      for (int r = 0; r < 4; r++) v[u][r] = buf[pad16(j + r * (N / 4))];
#pragma unroll
      for (int r = 1; r < 4; r++) v[u][r] = cmulc<INV>(v[u][r], __ldg(t4 + (r - 1) * 256 + j));
      dft4<INV>(v[u][0], v[u][1], v[u][2], v[u][3]);
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int j = tid + u * TH;
#pragma unroll
      for (int r = 0; r < 4; r++) buf[pad16(j + r * Ns)] = v[u][r];
    }
    __syncthreads();
  }
}

struct FftRing {  // spectra ring over the window index q: slot(q) = (base + q) mod cap
  uint32_t base, cap;
  __device__ __forceinline__ uint32_t slot(uint32_t q) const {
    const uint32_t s = base + q;
    return s >= cap ? s - cap : s;
  }
};

// Z[slot(q)][e][f] = FFT of the frames [(q - NP) P, (q - NP + 2) P) of track e's signal (frame 0 = the first of this
// render): history frames from the time-domain ring hist[track][2][H] (logical index i -> (hist_pos + i) mod H, oldest
// first), frames >= 0 from trackbuf; frames outside [-H, T) are zero.
template <int N>
__global__ void __launch_bounds__(FftShape<N>::THREADS, N == 4096 ? 4 : 16) fft_windows_kernel(const DFx* __restrict__ fx, uint32_t n_fx, uint32_t C,
                                                                           uint64_t H, uint64_t T, uint32_t NP, uint32_t q0,
                                                                           const float* __restrict__ hist, uint64_t hist_pos,
                                                                           const float* __restrict__ trackbuf, uint64_t tbs,
                                                                           const float2* __restrict__ tw, float2* __restrict__ Z,
                                                                           FftRing ring) {
  extern __shared__ __align__(16) float2 fsm[];
  const uint32_t q = q0 + blockIdx.x, e = blockIdx.y;
  if (!fx[e].reverb_on) return;
  const int tid = threadIdx.x;
  constexpr int P = N / 2, TH = FftShape<N>::THREADS;
  const int64_t t0 = ((int64_t)q - (int64_t)NP) * P;  // frame of window sample 0
  const float* hl = hist + (size_t)fx[e].track * 2 * H;
  const float* hr = hl + H;
  const float2* tb = reinterpret_cast<const float2*>(trackbuf) + (size_t)e * tbs;
  if (t0 >= 0 && t0 + N <= (int64_t)T && C == 2) {  // the whole window is this render's chain output: 16 loads in flight
    float2 v[16];
#pragma unroll
    for (int u = 0; u < 16; u++) v[u] = tb[t0 + tid + u * TH];
#pragma unroll
    for (int u = 0; u < 16; u++) fsm[pad16(tid + u * TH)] = v[u];
  } else {
    for (int n = tid; n < N; n += TH) {
      const int64_t t = t0 + n;
      float2 v = make_float2(0.0f, 0.0f);
      if (t >= 0) {
        if (t < (int64_t)T) {
          v = tb[t];
          if (C != 2) v.y = 0.0f;
        }
      } else if (t >= -(int64_t)H) {
        uint64_t i = hist_pos + (uint64_t)(t + (int64_t)H);
        if (i >= H) i -= H;
        v.x = __ldg(hl + i);
        if (C == 2) v.y = __ldg(hr + i);
      }
      fsm[pad16(n)] = v;
    }
  }
  __syncthreads();
  fft_inplace<N, false>(fsm, tw, tid);
  float2* dst = Z + ((size_t)ring.slot(q) * n_fx + e) * N;
#pragma unroll
  for (int u = 0; u < 16; u++) dst[tid + u * TH] = fsm[pad16(tid + u * TH)];
}

// Hs[p][f] = FFT of partition p of the impulse response (P taps, zero-padded to N) * 1/N (the inverse transform's scale)
template <int N>
__global__ void __launch_bounds__(FftShape<N>::THREADS) fft_ir_kernel(const float* __restrict__ ir, uint32_t L,
                                                                      const float2* __restrict__ tw, float2* __restrict__ Hs) {
  extern __shared__ __align__(16) float2 fsm[];
  const uint32_t p = blockIdx.x;
  const int tid = threadIdx.x;
  constexpr int P = N / 2, TH = FftShape<N>::THREADS;
  for (int n = tid; n < N; n += TH) {
    const uint64_t k = (uint64_t)p * P + n;
    fsm[pad16(n)] = make_float2((n < P && k < L) ? __ldg(ir + k) : 0.0f, 0.0f);
  }
  __syncthreads();
  fft_inplace<N, false>(fsm, tw, tid);
  const float s = 1.0f / (float)N;  // a power of two: exact
  for (int n = tid; n < N; n += TH) {
    const float2 r = fsm[pad16(n)];
    Hs[(size_t)p * N + n] = make_float2(r.x * s, r.y * s);
  }
}

// W[b][e][f] = sum_{p < NP} Hs[p][f] * Z[slot(b - p + NP - 1)][e][f]: output block b (frames [bP, bP + P)) takes, for
// partition p, the window ending at (b - p + 1) P, which is window q = b - p + NP - 1.
// Thread = (f, e, group of MAC_BG blocks): at step p it needs q = Q0 - p + j for its blocks j = 0 .. MAC_BG-1 — a window
// that slides down by one per step, kept in registers (slot (p - j) mod MAC_BG, static under the unroll).
// MAC_BG = output blocks per thread (the sliding window's length): 16 for long renders, 4 / 1 for renders of a few blocks
// (a realtime callback is one block: a longer window would only multiply zeros).
// (Measured and dropped: the same sum with H and Z staged through a shared-memory ring by 1 KiB cp.async.bulk copies —
// 186 us against 174 us for this version at cfg 5: two bulk copies per step per CTA run into the TMA issue rate of
// small copies, ~88 cycles each per SM, the same limit the mix kernel met with 2 KiB windows in round 1.)
template <int N, int MAC_BG>
__global__ void __launch_bounds__(128, MAC_BG == 16 ? 5 : 8) fft_mac_kernel(const DFx* __restrict__ fx, uint32_t n_fx, uint32_t NP, uint32_t NB,
                                                      const float2* __restrict__ Hs, const float2* __restrict__ Z, FftRing ring,
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
    if (q < NQ) win[(MAC_BG - j) % MAC_BG] = __ldg(zp + (size_t)ring.slot((uint32_t)q) * qstride);
  }
  // Inside the loop q = Q0 - p is always a valid window (0 <= q <= Q0 < NQ); its ring slot runs down from slot(Q0) and
  // wraps at most once. Steps go in groups of four: a group that neither crosses the wrap nor the end of the partitions
  // takes its eight loads from two base pointers with fixed strides, all in flight together.
  const int64_t S0 = (int64_t)ring.slot((uint32_t)Q0);
  const float2* hp0 = Hs + f;
  constexpr int UNR = MAC_BG < 4 ? 4 : MAC_BG;  // steps per unrolled loop body (a multiple of the window length)
  for (uint32_t p0 = 0; p0 < NP; p0 += UNR) {
#pragma unroll
    for (int g = 0; g < UNR / 4; g++) {
      const uint32_t pg = p0 + 4 * g;
      if (pg >= NP) break;
      int64_t s = S0 - (int64_t)pg;
      if (s < 0) s += ring.cap;
      float2 hv[4], zv[4];
      if (pg + 4 <= NP && s >= 3) {
        const float2* zq = zp + (size_t)s * qstride;
        const float2* hp = hp0 + (size_t)pg * N;
#pragma unroll
        for (int u = 0; u < 4; u++) {
          hv[u] = __ldg(hp + u * N);
          zv[u] = __ldg(zq - (size_t)u * qstride);
        }
      } else {
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const uint32_t p = pg + u;
          hv[u] = make_float2(0.0f, 0.0f);
          zv[u] = make_float2(0.0f, 0.0f);
          if (p < NP) {
            hv[u] = __ldg(hp0 + (size_t)p * N);
            zv[u] = __ldg(zp + (size_t)ring.slot((uint32_t)(Q0 - p)) * qstride);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int pp = (4 * g + u) % MAC_BG;
        win[pp] = zv[u];
#pragma unroll
        for (int j = 0; j < MAC_BG; j++) cmac(acc[j], hv[u], win[(pp - j + MAC_BG) % MAC_BG]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < MAC_BG; j++)
    if (b0 + j < NB) W[((size_t)(b0 + j) * n_fx + e) * N + f] = acc[j];
}

// frames [bP, bP + P) of track e = the last P samples of IFFT(W[b][e]) -> trackbuf[e][frame] = (L, R)
template <int N>
__global__ void __launch_bounds__(FftShape<N>::THREADS) ifft_blocks_kernel(const DFx* __restrict__ fx, uint32_t n_fx, uint32_t C,
                                                                           uint64_t T, const float2* __restrict__ W,
                                                                           const float2* __restrict__ tw,
                                                                           float* __restrict__ trackbuf, uint64_t tbs) {
  extern __shared__ __align__(16) float2 fsm[];
  const uint32_t b = blockIdx.x, e = blockIdx.y;
  if (!fx[e].reverb_on) return;
  const int tid = threadIdx.x;
  constexpr int P = N / 2, TH = FftShape<N>::THREADS;
  const float2* src = W + ((size_t)b * n_fx + e) * N;
  {
    float2 v[16];
#pragma unroll
    for (int u = 0; u < 16; u++) v[u] = src[tid + u * TH];
#pragma unroll
    for (int u = 0; u < 16; u++) fsm[pad16(tid + u * TH)] = v[u];
  }
  __syncthreads();
  fft_inplace<N, true>(fsm, tw, tid);
  float2* dst = reinterpret_cast<float2*>(trackbuf) + (size_t)e * tbs;
  for (int n = tid; n < P; n += TH) {
    const uint64_t t = (uint64_t)b * P + n;
    if (t < T) {
      const float2 v = fsm[pad16(P + n)];
      if (C == 2)
        dst[t] = v;
      else
        dst[t].x = v.x;
    }
  }
}

// the time-domain history ring takes this render's newest min(T, H) frames (read from trackbuf BEFORE the inverse
// transforms overwrite it): logical index i in [H - n, H) of the NEW history lives at (new_pos + i) mod H
__global__ void fft_save_kernel(const DFx* __restrict__ fx, uint32_t C, uint64_t H, uint64_t T, const float* __restrict__ trackbuf,
                                uint64_t tbs, float* __restrict__ hist, uint64_t new_pos) {
  const uint32_t e = blockIdx.y;
  if (!fx[e].reverb_on) return;
  const uint64_t n = T < H ? T : H;
  float* hl = hist + (size_t)fx[e].track * 2 * H;
  float* hr = hl + H;
  const float2* tb = reinterpret_cast<const float2*>(trackbuf) + (size_t)e * tbs;
  for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < n; u += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = H - n + u;  // logical index in the new history
    uint64_t s = new_pos + i;
    if (s >= H) s -= H;
    const float2 v = tb[T - n + u];
    hl[s] = v.x;
    if (C == 2) hr[s] = v.y;
  }
}

template <typename K>
cudaError_t set_smem(K k, int bytes) {
  return cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

}  // namespace

// partition size for a render of T frames: 2048 taps (4096-point transforms) for long renders, 512 for short ones (a
// realtime callback is then a whole number of partitions, which keeps the spectra ring warm); WBX_FFT_P=512|2048 pins it
uint32_t fir_fft_partition(uint64_t T) {
  if (const char* env = getenv("WBX_FFT_P")) {
    const int v = atoi(env);
    if (v == 512 || v == 2048) return (uint32_t)v;
  }
  return T >= 8192 ? 2048u : 512u;
}

static uint32_t fft_table_len(uint32_t N) { return N == 1024 ? FftTables<1024>::TOTAL : FftTables<4096>::TOTAL; }

// bytes of the twiddle tables followed by the partition spectra
size_t fir_fft_ir_bytes(uint32_t L, uint32_t P) {
  const uint32_t N = 2 * P, NP = (L + P - 1) / P;
  return (size_t)fft_table_len(N) * sizeof(float2) + (size_t)NP * N * sizeof(float2);
}

uint32_t fir_fft_windows(uint64_t T, uint32_t L, uint32_t P) {  // NQ = NB + NP - 1
  return (uint32_t)((T + P - 1) / P) + (L + P - 1) / P - 1;
}
size_t fir_fft_ring_bytes(uint32_t cap, uint32_t n_fx, uint32_t P) { return (size_t)cap * n_fx * 2 * P * sizeof(float2); }
size_t fir_fft_scratch_bytes(uint64_t T, uint32_t n_fx, uint32_t P) {  // W[NB][n_fx][N]
  return (size_t)((T + P - 1) / P) * n_fx * 2 * P * sizeof(float2);
}

// twiddles (f64 -> f32 on the host, copied synchronously) + the partition spectra of the impulse response
cudaError_t launch_fir_fft_prepare(const float* ir, uint32_t L, void* ir_spectra, uint32_t P, cudaStream_t stream) {
  const uint32_t N = 2 * P, NP = (L + P - 1) / P;
  const uint32_t TH = N / 16, TL = fft_table_len(N);
  std::vector<float2> tw(TL);
  const double PI2 = -2.0 * 3.14159265358979323846;
  uint32_t o = 0;
  for (uint32_t Ns = 16; Ns * 16 <= N; Ns *= 16)  // radix-16 passes with twiddles
    for (uint32_t r = 1; r < 16; r++)
      for (uint32_t t = 0; t < TH; t++) {
        const double a = PI2 * (double)r * (double)(t & (Ns - 1)) / (16.0 * (double)Ns);
        tw[o++] = make_float2((float)cos(a), (float)sin(a));
      }
  if (N == 1024)
    for (uint32_t r = 1; r < 4; r++)
      for (uint32_t j = 0; j < 256; j++) {
        const double a = PI2 * (double)r * (double)j / 1024.0;
        tw[o++] = make_float2((float)cos(a), (float)sin(a));
      }
  cudaError_t err = cudaStreamSynchronize(stream);
  if (err != cudaSuccess) return err;
  err = cudaMemcpy(ir_spectra, tw.data(), (size_t)TL * sizeof(float2), cudaMemcpyHostToDevice);
  if (err != cudaSuccess) return err;
  const float2* twd = reinterpret_cast<const float2*>(ir_spectra);
  float2* Hs = reinterpret_cast<float2*>(ir_spectra) + TL;
  if (N == 1024) {
    if ((err = set_smem(fft_ir_kernel<1024>, FftShape<1024>::SMEM)) != cudaSuccess) return err;
    fft_ir_kernel<1024><<<NP, FftShape<1024>::THREADS, FftShape<1024>::SMEM, stream>>>(ir, L, twd, Hs);
  } else {
    if ((err = set_smem(fft_ir_kernel<4096>, FftShape<4096>::SMEM)) != cudaSuccess) return err;
    fft_ir_kernel<4096><<<NP, FftShape<4096>::THREADS, FftShape<4096>::SMEM, stream>>>(ir, L, twd, Hs);
  }
  return cudaGetLastError();
}

template <int N>
static cudaError_t launch_fir_fft_t(const DFx* fx, uint32_t n_fx, uint32_t C, uint64_t T, const FirLaunch& a, float* trackbuf,
                                    uint64_t tbs, cudaStream_t stream) {
  constexpr uint32_t P = N / 2;
  const uint64_t H = a.L - 1;
  const uint32_t NP = (a.L + P - 1) / P;
  const uint32_t NB = (uint32_t)((T + P - 1) / P), NQ = NB + NP - 1;
  const float2* tw = reinterpret_cast<const float2*>(a.ir_aux);
  const float2* Hs = tw + FftTables<N>::TOTAL;
  float2* Z = reinterpret_cast<float2*>(a.fft_ring);
  float2* W = reinterpret_cast<float2*>(a.scratch);
  FftRing ring;
  ring.base = a.fft_ring_base;
  ring.cap = a.fft_ring_cap;
  using SH = FftShape<N>;
  cudaError_t err;
  if ((err = set_smem(fft_windows_kernel<N>, SH::SMEM)) != cudaSuccess) return err;
  if ((err = set_smem(ifft_blocks_kernel<N>, SH::SMEM)) != cudaSuccess) return err;
  const uint32_t q0 = a.fft_first_q < NQ ? a.fft_first_q : 0;  // windows below q0 are already in the ring
  fft_windows_kernel<N><<<dim3(NQ - q0, n_fx), SH::THREADS, SH::SMEM, stream>>>(fx, n_fx, C, H, T, NP, q0, a.hist, a.hist_pos, trackbuf,
                                                                                tbs, tw, Z, ring);
  if (H) {
    const uint64_t n = T < H ? T : H;
    uint64_t new_pos = (a.hist_pos + T) % H;
    fft_save_kernel<<<dim3((unsigned)((n + 255) / 256 < 256 ? (n + 255) / 256 : 256), n_fx), 256, 0, stream>>>(fx, C, H, T, trackbuf,
                                                                                                             tbs, a.hist, new_pos);
  }
  if (NB > 4)
    fft_mac_kernel<N, 16><<<dim3(N / 128, n_fx, (NB + 15) / 16), 128, 0, stream>>>(fx, n_fx, NP, NB, Hs, Z, ring, W);
  else if (NB > 1)
    fft_mac_kernel<N, 4><<<dim3(N / 128, n_fx, 1), 128, 0, stream>>>(fx, n_fx, NP, NB, Hs, Z, ring, W);
  else
    fft_mac_kernel<N, 1><<<dim3(N / 128, n_fx, 1), 128, 0, stream>>>(fx, n_fx, NP, NB, Hs, Z, ring, W);
  ifft_blocks_kernel<N><<<dim3(NB, n_fx), SH::THREADS, SH::SMEM, stream>>>(fx, n_fx, C, T, W, tw, trackbuf, tbs);
  return cudaGetLastError();
}

cudaError_t launch_fir_fft(const DFx* fx, uint32_t n_fx, uint32_t C, uint64_t T, const FirLaunch& a, float* trackbuf, uint64_t tbs,
                           cudaStream_t stream) {
  if (a.fft_p == 512) return launch_fir_fft_t<1024>(fx, n_fx, C, T, a, trackbuf, tbs, stream);
  return launch_fir_fft_t<4096>(fx, n_fx, C, T, a, trackbuf, tbs, stream);
}

}  // namespace wbx
