// wbx_fir_tc.cu — convolution reverb on the 5th-generation tensor cores (BASELINE cfg 5; extension, see wbx.h).
//
// y[n, s] = sum_k h[k] * x[n - k, s] for S signals (track x channel) as a GEMM with a Toeplitz operand:
//     D[128 output times, N signals] += A_i[128, 64] * B_i[64, N]        for tap chunks i = 0 .. n_chunks-1,  N = 128 | 64
//     A_i[m][j] = h[64 i - 64 + m - j]      (depends only on the impulse response: expanded once per IR, L2-resident)
//     B_i[j][s] = x[n0 + 64 - 64 i + j, s]  (a plain K-major slice of the signal planes, fetched by TMA; out-of-range
//                                            times are zero-filled by the tensor map)
// tcgen05.mma (kind::f16, fp16 inputs, f32 accumulators in TMEM) issued by one thread; operands staged by
// cp.async.bulk.tensor into 128B-swizzled shared memory through an mbarrier pipeline; accumulators read back with
// tcgen05.ld.
// f32 accuracy from a 2-term fp16 split of both operands: v * 2^e = v1 + v2 to 22 bits (fp16 carries 11), with the
// power-of-two scales chosen so that the largest |h| and the largest |x| of the render land in [2^13, 2^14) — the
// residual terms then stay inside fp16's normal range — and undone exactly in the epilogue. Three products are kept:
// h1x1 (order 1), h1x2 and h2x1 (order 2^-11); the dropped h2x2 is of order 2^-22. (Round 1 used a 3-term bf16 split, which
// needs six products for the same accuracy: bf16 carries 8 bits.)
// The tensor core's f32 accumulate truncates, so a long accumulation chain drifts (measured 7e-6 of peak after
// 336 accumulate steps): the leading product h1x1 and the two small ones go to SEPARATE TMEM accumulators, and
// every TC_G chunks the epilogue warps drain both into f32 registers (round-to-nearest adds) while the MMA warp
// continues in a second TMEM buffer (2 buffers x 2 accumulators x N columns).
// When there are too few (time tile, signal tile) CTAs to fill the GPU (strong scaling: a rank of an 8-GPU session holds
// 64 signals) the tap loop is split over gridDim.z CTAs whose f32 partial results a small kernel adds in fixed order.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "wbx_device.cuh"

namespace wbx {

namespace {

constexpr int TC_M = 128;               // output times per CTA (UMMA M, TMEM lanes)
constexpr int TC_K = 64;                // input times per chunk = one 128-byte swizzle row of fp16
constexpr int TC_G = 8;                 // chunks accumulated in TMEM before a drain (32 full-magnitude MMA steps)
constexpr int TC_TILE_BYTES = 128 * TC_K * 2;        // 16 KiB: [128 rows][64 fp16], SWIZZLE_128B
constexpr int TC_THREADS = 192;                      // warp 0 TMA, warp 1 MMA (+TMEM alloc), warps 2-5 epilogue
constexpr int TC_HEADER = 256;                       // bytes in front of the IR tiles / the signal planes (scales)
constexpr int TC_MAX_SPLITS = 4;
// signals per CTA (UMMA N, TMEM columns): 128, or 64 when the rank holds no more than 64 signals
template <int N>
struct TcShape {
  static constexpr int B_TILE_BYTES = N * TC_K * 2;
  static constexpr int STAGE_BYTES = 2 * TC_TILE_BYTES + 2 * B_TILE_BYTES;  // A1 A2 B1 B2
  static constexpr int STAGES = N == 128 ? 3 : 4;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*alignment*/ + 256 /*barriers*/;
  static constexpr uint32_t TMEM_COLS = 4 * N;  // 2 buffers x (leading + small products)
};

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// K-major, 128-byte-swizzled operand tile [rows][64 bf16]: 8-row groups are 1024 B apart (SBO), rows 128 B.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);  // start address, 16-byte units
  d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset
  d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
  return d;
}

// kind::f16 instruction descriptor: D = f32 (bits 4-5 = 1), A = B = fp16 (format fields 7-9, 10-12 = 0), both K-major, M x N
__device__ __forceinline__ uint32_t umma_idesc(uint32_t M, uint32_t N) {
  return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {  // arrives on `bar` when all prior MMAs have completed
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

struct FirTcParams {
  const DFx* fx;
  float* trackbuf;  // [n_fx][tbs][2]
  uint64_t tbs;     // frames per track in trackbuf (>= T, even)
  uint32_t C;       // bus channels: signal index = e * C + c
  uint32_t n_signals;
  uint64_t H, T;    // plane column of input time 0 (history rounded up to 8 columns), frames in this render
  uint32_t n_chunks;
  const float* ir_header;     // [0] = the power-of-two scale the IR tiles were expanded with
  const uint32_t* x_header;   // [0] = bits of max |x| over this render's inputs (the planes' scale follows from it)
  float* partials;            // gridDim.z > 1: [z][signal][T] f32 partial results, else unused
};

// power-of-two scale that puts `max_abs` into [2^13, 2^14) (1 for an all-zero input); exact to apply and to undo
__device__ __forceinline__ float tc_scale_for(float max_abs) {
  if (!(max_abs > 0.0f) || !(max_abs < 3.0e38f)) return 1.0f;
  int ex;
  frexpf(max_abs, &ex);  // max_abs = m * 2^ex, m in [0.5, 1)
  return ldexpf(1.0f, 14 - ex);
}

}  // namespace

// x * scale -> x1 + x2 with fp16 terms (22 bits); planes are [signals][W]: `lead` zero columns (so that the TMA box start
// — 16-byte aligned — lands on a multiple of 8 columns), then the len inputs, then zero padding
__global__ void split_f16_kernel(const float* __restrict__ x, uint64_t len, uint64_t lead, uint64_t W, uint32_t n_signals,
                                 const uint32_t* __restrict__ x_header, __half* __restrict__ p1, __half* __restrict__ p2) {
  const uint32_t s = blockIdx.y;
  const float scale = tc_scale_for(__uint_as_float(x_header[0]));
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < W; i += (uint64_t)gridDim.x * blockDim.x) {
    const float v = (i >= lead && i - lead < len) ? __fmul_rn(x[(size_t)s * len + (i - lead)], scale) : 0.0f;
    const __half a = __float2half_rn(v);
    const float r1 = __fsub_rn(v, __half2float(a));  // exact
    p1[(size_t)s * W + i] = a;
    p2[(size_t)s * W + i] = __float2half_rn(r1);
  }
}

// Toeplitz expansion of the impulse response: tiles[i][m][j] = h[64 i - 64 + m - j] * scale, two fp16 terms
__global__ void toeplitz_kernel(const float* __restrict__ h, uint32_t L, uint32_t n_chunks, float scale, float* __restrict__ header,
                                __half* __restrict__ a1, __half* __restrict__ a2) {
  if (blockIdx.x == 0 && threadIdx.x == 0) header[0] = scale;
  const uint64_t total = (uint64_t)n_chunks * TC_M * TC_K;
  for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t j = (uint32_t)(o % TC_K), m = (uint32_t)((o / TC_K) % TC_M), i = (uint32_t)(o / (TC_K * TC_M));
    const int64_t k = 64 * (int64_t)i - 64 + (int64_t)m - (int64_t)j;
    const float v = (k >= 0 && k < (int64_t)L) ? __fmul_rn(h[k], scale) : 0.0f;
    const __half a = __float2half_rn(v);
    const float r1 = __fsub_rn(v, __half2float(a));
    a1[o] = a;
    a2[o] = __float2half_rn(r1);
  }
}

template <int N>
__global__ void __launch_bounds__(TC_THREADS, 1)
fir_tc_kernel(const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapA2,
              const __grid_constant__ CUtensorMap mapX1, const __grid_constant__ CUtensorMap mapX2, const FirTcParams p) {
  using SH = TcShape<N>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // operand tiles need 1024-byte alignment (128B swizzle atoms)
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + SH::STAGES * SH::STAGE_BYTES);
  const uint32_t full0 = s32(&bars[0]), empty0 = s32(&bars[SH::STAGES]);
  const uint32_t tfull0 = s32(&bars[2 * SH::STAGES]), tempty0 = s32(&bars[2 * SH::STAGES + 2]);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(&bars[2 * SH::STAGES + 4]);

  const int64_t n0 = (int64_t)blockIdx.x * TC_M;
  const int s0 = (int)blockIdx.y * N;
  // this CTA's share of the tap chunks (gridDim.z > 1: strong scaling, the partial results are added afterwards)
  const uint32_t per = (p.n_chunks + gridDim.z - 1) / gridDim.z;
  const uint32_t c_lo = blockIdx.z * per;
  const uint32_t c_hi = c_lo + per < p.n_chunks ? c_lo + per : p.n_chunks;
  const uint32_t my_chunks = c_hi > c_lo ? c_hi - c_lo : 0;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < SH::STAGES; s++) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int b = 0; b < 2; b++) {
      mbar_init(tfull0 + 8 * b, 1);   // tcgen05.commit of the group's last MMA
      mbar_init(tempty0 + 8 * b, 4);  // one arrival per epilogue warp once it has drained the buffer
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: 2 buffers x (leading-product + small-products) accumulators x N f32 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(SH::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (uint32_t ii = 0; ii < my_chunks; ii++) {
        const uint32_t i = c_lo + ii;
        const uint32_t st = ii % SH::STAGES;
        mbar_wait(empty0 + 8 * st, ((ii / SH::STAGES) & 1u) ^ 1u);
        const uint32_t base = s32(tiles + (size_t)st * SH::STAGE_BYTES);
        const uint32_t bar = full0 + 8 * st;
        mbar_expect_tx(bar, SH::STAGE_BYTES);
        const int arow = (int)(i * TC_M);
        tma_load_2d(base + 0 * TC_TILE_BYTES, &mapA1, 0, arow, bar);
        tma_load_2d(base + 1 * TC_TILE_BYTES, &mapA2, 0, arow, bar);
        const int t = (int)((int64_t)p.H + n0 + 64 - 64 * (int64_t)i);  // plane column of input time n0 + 64 - 64 i
        tma_load_2d(base + 2 * TC_TILE_BYTES, &mapX1, t, s0, bar);
        tma_load_2d(base + 2 * TC_TILE_BYTES + SH::B_TILE_BYTES, &mapX2, t, s0, bar);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    const uint32_t idesc = umma_idesc(TC_M, N);
    for (uint32_t i = 0; i < my_chunks; i++) {
      const uint32_t st = i % SH::STAGES;
      const uint32_t g = i / TC_G, b = g & 1u, ig = i % TC_G;
      if (ig == 0) mbar_wait(tempty0 + 8 * b, ((g >> 1) & 1u) ^ 1u);  // epilogue drained this TMEM buffer
      mbar_wait(full0 + 8 * st, (i / SH::STAGES) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t base = s32(tiles + (size_t)st * SH::STAGE_BYTES);
        const uint32_t d_hi = tmem_base + b * (2u * N), d_lo = d_hi + N;
        // (h term, x term): q == 0 is the leading product h1 x1; h1 x2 and h2 x1 are of order 2^-11
        const int ha[3] = {0, 0, 1}, xb[3] = {0, 1, 0};
#pragma unroll
        for (int q = 0; q < 3; q++) {
          const uint64_t da = umma_desc(base + ha[q] * TC_TILE_BYTES);
          const uint64_t db = umma_desc(base + 2 * TC_TILE_BYTES + xb[q] * SH::B_TILE_BYTES);
#pragma unroll
          for (int ks = 0; ks < TC_K / 16; ks++) {  // UMMA K = 16 fp16 = 32 bytes = +2 in 16-byte address units
            const uint32_t first = (q == 0) ? (ig | ks) : (ig | (q - 1) | ks);  // 0 on the accumulator's first MMA of the group
            umma_f16(q == 0 ? d_hi : d_lo, da + 2 * ks, db + 2 * ks, idesc, first ? 1u : 0u);
          }
        }
        umma_commit(empty0 + 8 * st);                                       // frees this stage's shared memory
        if (ig == TC_G - 1 || i + 1 == my_chunks) umma_commit(tfull0 + 8 * b);  // this group's accumulators are complete
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue: drain TMEM groups into f32 registers, then registers -> track buffer (or this split's partials) =====
    const uint32_t quarter = warp & 3u;  // a warp may only touch TMEM lanes [32 * (warp % 4), +32)
    const int64_t n = n0 + quarter * 32 + lane;
    float acc[N];
#pragma unroll
    for (int q = 0; q < N; q++) acc[q] = 0.0f;
    const uint32_t n_groups = (my_chunks + TC_G - 1) / TC_G;
    for (uint32_t g = 0; g < n_groups; g++) {
      const uint32_t b = g & 1u;
      mbar_wait(tfull0 + 8 * b, (g >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t trow = tmem_base + ((quarter * 32u) << 16) + b * (2u * N);
#pragma unroll
      for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16], w[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(trow + (uint32_t)c0));
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]),
              "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
            : "r"(trow + (uint32_t)N + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int q = 0; q < 16; q++) acc[c0 + q] = __fadd_rn(acc[c0 + q], __fadd_rn(__uint_as_float(v[q]), __uint_as_float(w[q])));
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty0 + 8 * b) : "memory");
    }
    // undo the operand scales (powers of two: exact)
    const float unscale = 1.0f / (p.ir_header[0] * tc_scale_for(__uint_as_float(p.x_header[0])));
    if (n < (int64_t)p.T) {
#pragma unroll
      for (int q = 0; q < N; q++) {
        const uint32_t sig = (uint32_t)(s0 + q);
        if (sig < p.n_signals) {
          const float y = __fmul_rn(acc[q], unscale);
          if (gridDim.z > 1) {
            p.partials[((size_t)blockIdx.z * p.n_signals + sig) * p.T + (size_t)n] = y;
          } else {
            const uint32_t e = sig / p.C, c = sig % p.C;
            if (p.fx[e].reverb_on) p.trackbuf[((size_t)e * p.tbs + (size_t)n) * 2 + c] = y;
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(SH::TMEM_COLS) : "memory");
  }
}

// tap-split renders: y = sum over the splits in index order (f32, round to nearest), into the track buffer
__global__ void fir_tc_combine_kernel(const FirTcParams p, uint32_t splits) {
  const uint32_t sig = blockIdx.y;
  const uint32_t e = sig / p.C, c = sig % p.C;
  if (!p.fx[e].reverb_on) return;
  for (uint64_t n = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; n < p.T; n += (uint64_t)gridDim.x * blockDim.x) {
    float y = p.partials[(size_t)sig * p.T + n];
    for (uint32_t z = 1; z < splits; z++) y = __fadd_rn(y, p.partials[((size_t)z * p.n_signals + sig) * p.T + n]);
    p.trackbuf[((size_t)e * p.tbs + n) * 2 + c] = y;
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D fp16 tensor [rows][cols] (cols contiguous), box = 64 cols x box_rows rows, 128-byte swizzle, zero fill
static bool make_map(CUtensorMap* m, void* base, uint64_t cols, uint64_t rows, uint64_t pitch_elems, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {pitch_elems * 2};
  const cuuint32_t box[2] = {TC_K, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

uint32_t fir_tc_chunks(uint32_t L) { return (L - 1 + 64 + 63) / 64 + 1; }
size_t fir_tc_tiles_bytes(uint32_t L) { return TC_HEADER + (size_t)2 * fir_tc_chunks(L) * TC_TILE_BYTES; }
static uint64_t fir_tc_origin(uint64_t H) { return (H + 7) & ~(uint64_t)7; }  // plane column of input time 0
uint64_t fir_tc_plane_width(uint64_t H, uint64_t T) { return (fir_tc_origin(H) + T + 7) & ~(uint64_t)7; }
// scratch of one render: header (max |x| word) | two fp16 planes [S][W] | f32 partials of a tap-split render
size_t fir_tc_scratch_bytes(uint64_t H, uint64_t T, uint32_t S) {
  const uint64_t W = fir_tc_plane_width(H, T);
  return TC_HEADER + (((size_t)2 * S * W * 2 + 255) & ~(size_t)255) + (size_t)TC_MAX_SPLITS * S * T * sizeof(float) + 256;
}
// the word fir_gather_kernel folds max |x| into (zeroed by launch_effects before the gather)
uint32_t* fir_tc_max_word(void* scratch) { return (uint32_t*)scratch; }
int fir_tc_split_factor() { return 3; }  // fp16 products per tap (bench.py reports the issued tensor work with it)

// once per impulse response: header (scale) + the two Toeplitz term planes, each [n_chunks * 128][64] fp16.
// h_scale = the power of two that puts max |h| into [2^13, 2^14) (computed by the caller, which holds h on the host).
cudaError_t launch_fir_tc_prepare(const float* ir, uint32_t L, void* tiles, float h_scale, cudaStream_t stream) {
  const uint32_t nc = fir_tc_chunks(L);
  __half* a = (__half*)((uint8_t*)tiles + TC_HEADER);
  const size_t plane = (size_t)nc * TC_M * TC_K;
  toeplitz_kernel<<<1024, 256, 0, stream>>>(ir, L, nc, h_scale, (float*)tiles, a, a + plane);
  return cudaGetLastError();
}

template <int N>
static cudaError_t launch_fir_tc_n(const CUtensorMap* mA, __half* x1, uint64_t W, uint32_t S, const FirTcParams& p, uint32_t splits,
                                   cudaStream_t stream) {
  CUtensorMap mX[2];
  for (int t = 0; t < 2; t++)
    if (!make_map(&mX[t], x1 + (size_t)t * S * W, W, S, W, N)) return cudaErrorNotSupported;
  auto kfn = fir_tc_kernel<N>;
  cudaError_t err = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, TcShape<N>::SMEM);
  if (err != cudaSuccess) return err;
  kfn<<<dim3((unsigned)((p.T + TC_M - 1) / TC_M), (S + N - 1) / N, splits), TC_THREADS, TcShape<N>::SMEM, stream>>>(mA[0], mA[1], mX[0],
                                                                                                             mX[1], p);
  return cudaGetLastError();
}

// xin: f32 [n_signals][H + T]; scratch: fir_tc_scratch_bytes, its max word already holds max |xin|
cudaError_t launch_fir_tc(const DFx* fx, uint32_t n_fx, uint32_t C, uint64_t H, uint64_t T, uint32_t L, void* tiles,
                          const float* xin, void* scratch, float* trackbuf, uint64_t tbs, int n_sm, cudaStream_t stream) {
  const uint32_t S = n_fx * C;
  if (S == 0 || T == 0) return cudaSuccess;
  const uint64_t W = fir_tc_plane_width(H, T);
  const uint32_t nc = fir_tc_chunks(L);
  __half* x1 = (__half*)((uint8_t*)scratch + TC_HEADER);
  __half* x2 = x1 + (size_t)S * W;
  float* partials = (float*)((uint8_t*)scratch + TC_HEADER + (((size_t)2 * S * W * 2 + 255) & ~(size_t)255));
  const uint64_t origin = fir_tc_origin(H);
  split_f16_kernel<<<dim3((unsigned)((W + 255) / 256 < 2048 ? (W + 255) / 256 : 2048), S), 256, 0, stream>>>(
      xin, H + T, origin - H, W, S, (const uint32_t*)scratch, x1, x2);
  CUtensorMap mA[2];
  __half* a = (__half*)((uint8_t*)tiles + TC_HEADER);
  const size_t aplane = (size_t)nc * TC_M * TC_K;
  for (int t = 0; t < 2; t++)
    if (!make_map(&mA[t], a + t * aplane, TC_K, (uint64_t)nc * TC_M, TC_K, 128)) return cudaErrorNotSupported;
  // tile width: 64 signals per CTA when that wastes nothing; tap split: aim at >= 4 CTAs per SM's worth of work units
  const int N = (S <= 64 || (S % 128 != 0 && S % 128 <= 64 && S < 4 * 128)) ? 64 : 128;
  const uint64_t ctas = ((T + TC_M - 1) / TC_M) * ((S + N - 1) / N);
  uint32_t splits = 1;
  while (splits < TC_MAX_SPLITS && ctas * splits < (uint64_t)n_sm * 4 && nc / (splits * 2) >= 64) splits *= 2;
  if (const char* env = getenv("WBX_FIR_SPLITS")) {
    const int v = atoi(env);
    if (v == 1 || v == 2 || v == 4) splits = (uint32_t)v;
  }
  FirTcParams p;
  p.fx = fx;
  p.trackbuf = trackbuf;
  p.tbs = tbs;
  p.C = C;
  p.n_signals = S;
  p.H = origin;  // the kernel only needs the plane column of input time 0
  p.T = T;
  p.n_chunks = nc;
  p.ir_header = (const float*)tiles;
  p.x_header = (const uint32_t*)scratch;
  p.partials = partials;
  cudaError_t err = N == 64 ? launch_fir_tc_n<64>(mA, x1, W, S, p, splits, stream) : launch_fir_tc_n<128>(mA, x1, W, S, p, splits, stream);
  if (err != cudaSuccess) return err;
  if (splits > 1) {
    fir_tc_combine_kernel<<<dim3((unsigned)((T + 255) / 256 < 1024 ? (T + 255) / 256 : 1024), S), 256, 0, stream>>>(p, splits);
    err = cudaGetLastError();
  }
  return err;
}

}  // namespace wbx
