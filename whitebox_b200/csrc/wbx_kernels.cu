// wbx_kernels.cu — sm_100a kernels of the whitebox mixing hot path.
//
//   expand_schedule   one thread per wbx_segment: replays the f64 recurrence of Sampler::sample_offset_
//                     (dsp/sampler.cpp:99-104,209) and writes one 16-B DCell per callback of the run.
//   mix_kernel<FPL>   the fused render: Sampler::stream (unity + 2-tap linear, all source formats,
//                     dsp/sampler.cpp:34-59,106-158) -> clip gain -> volume*pan (dsp/dsp_ops.h:27-31) -> VU
//                     block peak (engine/vu_meter.h:20-30) -> bus sum (core/audio_buffer.h:73-82) -> clamp
//                     (engine/engine.cpp:1627-1636). Persistent warps pull (block, frame-tile, track-group)
//                     items from an atomic queue; each warp owns a private shared-memory ring that it fills
//                     itself with cp.async.bulk (TMA 1-D bulk copies, SASS UBLKCP) completing on mbarriers,
//                     so the only global loads in the loop are the 16-B cells / L2-resident span records.
//                     A warp owns every output sample of its tile and adds tracks in index order, so with
//                     one track group the f32 sum is bit-identical to the reference's sequential mix.
//   clamp_kernel      engine.cpp:1627-1636 alone, for the bus after a cross-GPU reduce.
//   interleave_kernel core/audio_format_conv.cpp:5-106, planar f32 bus -> interleaved device format.
//
// Built with -fmad=false: every f32/f64 multiply and add is separately rounded like the reference's ISO C++
// x86-64 build (no FMA contraction); the arithmetic that decides parity also uses explicit _rn intrinsics.
#include <cuda_runtime.h>

#include <type_traits>

#include <cstdint>
#include <cstdlib>

#include "wbx_device.cuh"

namespace wbx {

// ---------------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk async copy (TMA) into shared memory
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// global -> shared bulk copy, 16-B aligned addresses, size a multiple of 16; completes `bytes` on `bar`.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// schedule expansion
// ---------------------------------------------------------------------------------------------------------
// num_actual_samples = min(num_samples, (uint32_t)ceil((count - offset) / speed))   (sampler.cpp:102-104)
__device__ __forceinline__ uint32_t clipped_length(double cnt, double pos, double speed, uint32_t length, double safe) {
  const double rem = __dsub_rn(cnt, pos);
  if (rem >= safe) return length;  // rem >= (length + 1) * speed  =>  ceil(rem / speed) >= length
  const double m = ceil(__ddiv_rn(rem, speed));
  const uint32_t mm = m >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)m;
  return mm < length ? mm : length;
}

// One WARP per segment. Unity-speed runs that start on an integer frame advance by exact integer additions, so
// pos_b = pos0 + b * length in closed form and the lanes take callbacks b = lane, lane + 32, ... Any other run
// replays the reference's recurrence pos += (double)num_samples * speed, one rounding per callback
// (sampler.cpp:103,209) — split over the lanes by callback range, each lane reaching its range's start with the exact
// per-binade closed form of that recurrence.
__device__ __forceinline__ bool span_closed_form(const DSpan& s) {
  return s.speed == 1.0 && floor(s.pos0) == s.pos0 && s.pos0 + (double)s.n_blocks * (double)s.length < 4.0e15;
}

// Closed-form runs of a long render, one LANE per span: consecutive spans are consecutive tracks in the common case, so
// the 32 cells a warp writes for one callback are one coalesced 512-byte store (the warp-per-span form below writes
// cells n_tracks * slots * 16 B apart). blockIdx.y selects a range of 64 callbacks.
__global__ void expand_closed_kernel(const DSpan* __restrict__ spans, uint32_t n_spans, DCell* __restrict__ cells,
                                     uint32_t n_tracks, uint32_t slots, uint32_t n_blocks) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_spans) return;
  const DSpan s = spans[i];
  if (!span_closed_form(s)) return;
  const double cnt = (double)s.count;
  const double len = (double)s.length;
  const double safe = __dmul_rn(len + 1.0, s.speed);
  const size_t stride = (size_t)n_tracks * slots;
  DCell* out = cells + (size_t)s.track * slots + s.slot;
  const uint32_t lo = blockIdx.y * 64u > s.block0 ? blockIdx.y * 64u : s.block0;
  uint32_t hi = blockIdx.y * 64u + 64u;
  if (hi > n_blocks) hi = n_blocks;
  if (hi > s.block0 + s.n_blocks) hi = s.block0 + s.n_blocks;
  for (uint32_t k = lo; k < hi; k++) {
    const double pos = s.pos0 + (double)(k - s.block0) * len;  // exact
    if (pos >= cnt) break;  // finished streaming; sample_offset_ no longer advances (sampler.cpp:99-100)
    DCell c;
    c.pos = pos;
    c.span = i;
    c.n_act = clipped_length(cnt, pos, s.speed, s.length, safe);
    out[(size_t)k * stride] = c;
  }
}

__global__ void expand_schedule(const DSpan* __restrict__ spans, uint32_t n_spans, DCell* __restrict__ cells,
                                uint32_t n_tracks, uint32_t slots, uint32_t skip_closed) {
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (w >= n_spans) return;
  const DSpan s = spans[w];
  const double cnt = (double)s.count;
  const double len = (double)s.length;
  const double adv = __dmul_rn(len, s.speed);                  // (double)num_samples * playback_speed_
  const double safe = __dmul_rn(len + 1.0, s.speed);
  DCell* out = cells + ((size_t)s.block0 * n_tracks + s.track) * slots + s.slot;
  const size_t stride = (size_t)n_tracks * slots;
  const bool closed = span_closed_form(s);
  if (closed && skip_closed) return;  // written by expand_closed_kernel
  if (closed) {
    for (uint32_t b = lane; b < s.n_blocks; b += 32) {
      const double pos = s.pos0 + (double)b * len;  // exact
      if (pos >= cnt) break;  // finished streaming; sample_offset_ no longer advances (sampler.cpp:99-100)
      DCell c;
      c.pos = pos;
      c.span = w;
      c.n_act = clipped_length(cnt, pos, s.speed, s.length, safe);
      out[(size_t)b * stride] = c;
    }
  } else {
    // lane L owns callbacks [L*chunk, (L+1)*chunk): it jumps to its first one with the exact closed form of the
    // recurrence (advance_rounded_impl: same values as stepping from callback 0) and steps through its own 1/32
    const uint32_t chunk = (s.n_blocks + 31) / 32;
    const uint32_t b0 = lane * chunk;
    if (b0 < s.n_blocks) {
      double pos = s.pos0;
      const uint32_t done = advance_rounded_impl(&pos, adv, b0, cnt);
      if (done == b0) {  // otherwise the sample was exhausted before this lane's range: its cells stay silent
        const uint32_t b1 = b0 + chunk < s.n_blocks ? b0 + chunk : s.n_blocks;
        for (uint32_t b = b0; b < b1; b++) {
          if (pos >= cnt) break;  // finished streaming; sample_offset_ no longer advances (sampler.cpp:99-100)
          DCell c;
          c.pos = pos;
          c.span = w;
          c.n_act = clipped_length(cnt, pos, s.speed, s.length, safe);
          out[(size_t)b * stride] = c;
          pos = __dadd_rn(pos, adv);  // next_sample_offset (sampler.cpp:103,209)
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// element access with the reference's normalisation rules
// ---------------------------------------------------------------------------------------------------------
enum : uint32_t { F_I16 = 3, F_I24 = 5, F_I32 = 7, F_F32 = 9 };

__device__ __forceinline__ float clampf_ref(float x, float lo, float hi) {  // math::clamp, core_math.h:34-38
  float m = x < hi ? x : hi;
  return m > lo ? m : lo;
}
__device__ __forceinline__ double clampd_ref(double x, double lo, double hi) {
  double m = x < hi ? x : hi;
  return m > lo ? m : lo;
}

// unity-speed branch: normalise + clamp to [-1, 1] (sampler.cpp:109-156)
template <uint32_t FMT>
__device__ __forceinline__ float load_unity(const void* row, int64_t idx) {
  if (FMT == F_F32) {
    return ((const float*)row)[idx];
  } else if (FMT == F_I16) {
    const float norm = 1.0f / 32767.0f;  // i16_pcm_normalizer, sampler.cpp:95
    float v = __fmul_rn((float)((const int16_t*)row)[idx], norm);
    return clampf_ref(v, -1.0f, 1.0f);
  } else if (FMT == F_I24) {
    const double norm = 1.0 / 8388607.0;  // sampler.cpp:96
    double v = __dmul_rn((double)((const int32_t*)row)[idx], norm);
    return (float)clampd_ref(v, -1.0, 1.0);
  } else {
    const double norm = 1.0 / 2147483647.0;  // sampler.cpp:97
    double v = __dmul_rn((double)((const int32_t*)row)[idx], norm);
    return (float)clampd_ref(v, -1.0, 1.0);
  }
}

// linear branch: normalise, no clamp (sample_linear, sampler.cpp:53-54; get_pcm_sample_normalizer :7-18)
template <uint32_t FMT>
__device__ __forceinline__ float load_lin(const void* row, int64_t idx) {
  if (FMT == F_F32) {
    return ((const float*)row)[idx];
  } else if (FMT == F_I16) {
    const float norm = (float)(1.0 / 32767.0);
    return __fmul_rn(norm, (float)((const int16_t*)row)[idx]);
  } else if (FMT == F_I24) {
    const double norm = 1.0 / 8388607.0;
    return (float)__dmul_rn(norm, (double)((const int32_t*)row)[idx]);
  } else {
    const double norm = 1.0 / 2147483647.0;
    return (float)__dmul_rn(norm, (double)((const int32_t*)row)[idx]);
  }
}

// run-time format dispatch for the generic path (keeps its code size down: it is the rare path)
__device__ __forceinline__ float load_unity_rt(uint32_t fmt, const void* row, int64_t idx) {
  switch (fmt) {
    case F_I16: return load_unity<F_I16>(row, idx);
    case F_I24: return load_unity<F_I24>(row, idx);
    case F_I32: return load_unity<F_I32>(row, idx);
    default: return load_unity<F_F32>(row, idx);
  }
}
__device__ __forceinline__ float load_lin_rt(uint32_t fmt, const void* row, int64_t idx) {
  switch (fmt) {
    case F_I16: return load_lin<F_I16>(row, idx);
    case F_I24: return load_lin<F_I24>(row, idx);
    case F_I32: return load_lin<F_I32>(row, idx);
    default: return load_lin<F_F32>(row, idx);
  }
}

// term = (sample * clip_gain) * track_gain, bus += term, peak = max(peak, |term|)
__device__ __forceinline__ void accumulate(float s, float gain, float tg, float& acc, float& pk) {
  const float term = __fmul_rn(__fmul_rn(s, gain), tg);  // sampler.cpp:154 then dsp_ops.h:29
  acc = __fadd_rn(acc, term);                            // audio_buffer.h:79
  pk = fmaxf(pk, fabsf(term));                           // vu_meter.h:24
}

// Lane <-> frame mapping of a tile of T = 32*FPL frames: lane owns frame pairs (2q, 2q+1), q = lane + 32*i,
// i < FPL/2 — for an interleaved stereo f32 window that pair is exactly one 16-byte shared-memory word.
// acc[i*2+e] is the (L, R) bus accumulator of frame 2*(lane+32*i)+e.

// Generic per-frame path: any format, unity or linear, staged window (shared) or direct (global) rows.
// `row` points at window frame 0 (frame-interleaved, NCH channels); d.base is that frame's sample index.
// Polyphase extension (include/wbx.h): 16 taps of phase floor(frac * 128) on source frames ix - 7 .. ix + 8, f32 fused
// multiply-adds in tap order (oracle/wb_oracle.c sample_polyphase). `row` addresses frame 0 of a frame-interleaved
// stereo f32 buffer; frames before the sample and after its end read the zero padding of the device layout.
__device__ __forceinline__ float2 poly_frame(const float* __restrict__ table, const float2* row, int64_t ix, double fd) {
  const int ph = __double2int_rz(__dmul_rn(fd, 128.0));
  const float4* h4 = reinterpret_cast<const float4*>(table + ph * 16);
  float2 acc = make_float2(0.0f, 0.0f);
  const float2* p = row + (ix - 7);
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const float4 h = __ldg(h4 + q);
    const float2 s0 = p[4 * q + 0], s1 = p[4 * q + 1], s2 = p[4 * q + 2], s3 = p[4 * q + 3];
    acc.x = __fmaf_rn(h.x, s0.x, acc.x);
    acc.y = __fmaf_rn(h.x, s0.y, acc.y);
    acc.x = __fmaf_rn(h.y, s1.x, acc.x);
    acc.y = __fmaf_rn(h.y, s1.y, acc.y);
    acc.x = __fmaf_rn(h.z, s2.x, acc.x);
    acc.y = __fmaf_rn(h.z, s2.y, acc.y);
    acc.x = __fmaf_rn(h.w, s3.x, acc.x);
    acc.y = __fmaf_rn(h.w, s3.y, acc.y);
  }
  return acc;
}

// Fade extension (include/wbx.h): envelope of clip-relative output frame n.
struct FadeEnv {
  double n0;  // clip frame of segment-relative frame 0
  double fin, fout, len;
  __device__ __forceinline__ float at(int32_t jj) const {
    const double n = __dadd_rn(n0, (double)jj);
    double e = 1.0;
    if (fin > 0.0) {
      const double r = __ddiv_rn(n, fin);
      e = r < 1.0 ? r : 1.0;
    }
    if (fout > 0.0) {
      double r = __ddiv_rn(__dsub_rn(len, n), fout);
      r = r > 0.0 ? r : 0.0;
      r = r < 1.0 ? r : 1.0;
      e = __dmul_rn(e, r);
    }
    return __double2float_rn(e);
  }
};

template <int FPL, bool UNITY, bool FADE, bool POLY>
__device__ __forceinline__ void consume_gen_t(const Desc& d, const void* row, float2 (&acc)[FPL], float& pkL,
                                              float& pkR, int lane, bool two, const FadeEnv& fe, const float* poly) {
  const int64_t ip = (int64_t)(uint32_t)(int64_t)d.pos;  // (uint32_t)sample_offset_, sampler.cpp:107
  const uint32_t FMT = d.fmt & 0x3fu;
  const int NCH = (d.fmt & 0x80u) ? 1 : 2;
  const bool use_poly = POLY && !UNITY && (d.fmt & 0x40u) && poly != nullptr;  // 0x40: stereo f32 -> stereo bus only
#pragma unroll
  for (int i = 0; i < FPL / 2; i++) {
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int fr = 2 * (lane + 32 * i) + e;
      if (fr >= (int)d.lo && fr < (int)d.hi) {
        const int32_t jj = fr + d.jrel0;
        float sL, sR = 0.0f;
        if (UNITY) {
          const int64_t idx = (ip + jj - d.base) * NCH;
          sL = load_unity_rt(FMT, row, idx);
          if (two) sR = (NCH == 2) ? load_unity_rt(FMT, row, idx + 1) : sL;
        } else {
          const double x = __dadd_rn(d.pos, __dmul_rn((double)jj, d.speed));  // sampler.cpp:50
          const int64_t ix = __double2ll_rz(x);                                // :51
          const float fx = __double2float_rn(__dsub_rn(x, __ll2double_rn(ix)));  // :52
          const int64_t idx = (ix - d.base) * NCH;
          if (use_poly) {
            const float2 pv = poly_frame(poly, reinterpret_cast<const float2*>(row) - d.base, ix,
                                         __dsub_rn(x, __ll2double_rn(ix)));
            sL = pv.x;
            sR = pv.y;
          } else {
          const float a = load_lin_rt(FMT, row, idx), b = load_lin_rt(FMT, row, idx + NCH);
          sL = __fadd_rn(a, __fmul_rn(fx, __fsub_rn(b, a)));  // :55
          if (two) {
            if (NCH == 2) {
              const float a2 = load_lin_rt(FMT, row, idx + 1), b2 = load_lin_rt(FMT, row, idx + 3);
              sR = __fadd_rn(a2, __fmul_rn(fx, __fsub_rn(b2, a2)));
            } else {
              sR = sL;
            }
          }
          }
        }
        if (FADE) {  // (src * gain) * env, then the track gain as usual
          const float env = fe.at(jj);
          const float mL = __fmul_rn(__fmul_rn(sL, d.gain), env);
          const float tL = __fmul_rn(mL, d.tg[0]);
          acc[i * 2 + e].x = __fadd_rn(acc[i * 2 + e].x, tL);
          pkL = fmaxf(pkL, fabsf(tL));
          if (two) {
            const float mR = __fmul_rn(__fmul_rn(sR, d.gain), env);
            const float tR = __fmul_rn(mR, d.tg[1]);
            acc[i * 2 + e].y = __fadd_rn(acc[i * 2 + e].y, tR);
            pkR = fmaxf(pkR, fabsf(tR));
          }
        } else {
          accumulate(sL, d.gain, d.tg[0], acc[i * 2 + e].x, pkL);
          if (two) accumulate(sR, d.gain, d.tg[1], acc[i * 2 + e].y, pkR);
        }
      }
    }
  }
}

template <int FPL, bool FADE, bool POLY>
__device__ __forceinline__ void consume_gen_f(const Desc& d, const void* row, float2 (&acc)[FPL], float& pkL,
                                              float& pkR, int lane, bool two, const FadeEnv& fe, const float* poly) {
  if (d.speed == 1.0)  // playback_speed_ == 1.0, sampler.cpp:106
    consume_gen_t<FPL, true, FADE, false>(d, row, acc, pkL, pkR, lane, two, fe, poly);
  else
    consume_gen_t<FPL, false, FADE, POLY>(d, row, acc, pkL, pkR, lane, two, fe, poly);
}

// EXT == false is the lean build used when no segment of the render carries an extension flag (fade, polyphase):
// the rarely used paths then cost neither registers nor instruction-cache space on the reference-parity path.
template <int FPL, bool EXT>
__device__ __forceinline__ void consume_gen(const Desc& d, const void* row, float2 (&acc)[FPL], float& pkL,
                                            float& pkR, int lane, bool two, const DSpan* spans, const float* poly) {
  FadeEnv fe;
  fe.n0 = 0.0;
  fe.fin = 0.0;
  fe.fout = 0.0;
  fe.len = 0.0;
  if (EXT && (d.kind == K_FADE || d.kind == K_DIRECT_FADE)) {
    const DSpan* sp = spans + d.span;
    fe.fin = __ldg(&sp->fade_in);
    fe.fout = __ldg(&sp->fade_out);
    fe.len = __ldg(&sp->clip_len);
    // clip frame of segment-relative frame 0 in this callback (exact: integers)
    fe.n0 = __ldg(&sp->clip_frame) + (double)d.block_in_run * (double)__ldg(&sp->length);
    consume_gen_f<FPL, true, EXT>(d, row, acc, pkL, pkR, lane, two, fe, poly);
  } else {
    consume_gen_f<FPL, false, EXT>(d, row, acc, pkL, pkR, lane, two, fe, poly);
  }
}

// Fast path: stereo f32, unity speed, the whole tile, 16-B aligned window. One 128-bit shared load = two
// (L, R) frames; gain, pan and the bus add run as packed f32x2 (each component separately rounded, rn).
template <int FPL>
__device__ __forceinline__ void consume_fast(const uint8_t* row, float gain, float tgL, float tgR,
                                             float2 (&acc)[FPL], float& pkL, float& pkR, int lane) {
  const float4* r4 = reinterpret_cast<const float4*>(row);
  const float2 g2 = make_float2(gain, gain);
  const float2 t2 = make_float2(tgL, tgR);
  float4 v[FPL / 2];
#pragma unroll
  for (int i = 0; i < FPL / 2; i++) v[i] = r4[lane + 32 * i];
#pragma unroll
  for (int i = 0; i < FPL / 2; i++) {
    const float2 a = __fmul2_rn(__fmul2_rn(make_float2(v[i].x, v[i].y), g2), t2);  // frame 2q:   (L, R)
    const float2 b = __fmul2_rn(__fmul2_rn(make_float2(v[i].z, v[i].w), g2), t2);  // frame 2q+1: (L, R)
    acc[i * 2 + 0] = __fadd2_rn(acc[i * 2 + 0], a);
    acc[i * 2 + 1] = __fadd2_rn(acc[i * 2 + 1], b);
    pkL = fmaxf(fmaxf(pkL, fabsf(a.x)), fabsf(b.x));
    pkR = fmaxf(fmaxf(pkR, fabsf(a.y)), fabsf(b.y));
  }
}

// packed (L, R) version of `accumulate` for one frame
__device__ __forceinline__ void accumulate2(float2 s, float2 g2, float2 t2, float2& acc, float& pkL, float& pkR) {
  const float2 term = __fmul2_rn(__fmul2_rn(s, g2), t2);
  acc = __fadd2_rn(acc, term);
  pkL = fmaxf(pkL, fabsf(term.x));
  pkR = fmaxf(pkR, fabsf(term.y));
}

// Stereo f32, unity speed, window not aligned to the tile (odd start frame, partial coverage): 64-bit shared
// loads of (L, R) frames, packed math.
template <int FPL, bool FULL>
__device__ __forceinline__ void consume_uni_t(const Desc& d, const uint8_t* row, float2 (&acc)[FPL], float& pkL,
                                              float& pkR, int lane) {
  const float2* r2 = reinterpret_cast<const float2*>(row);
  const int lo = d.lo, hi = d.hi;
  const int shift = (int)((int64_t)(uint32_t)(int64_t)d.pos + d.jrel0 - d.base);  // window index of tile frame 0
  const float2 g2 = make_float2(d.gain, d.gain);
  const float2 t2 = make_float2(d.tg[0], d.tg[1]);
#pragma unroll
  for (int i = 0; i < FPL / 2; i++) {
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int fr = 2 * (lane + 32 * i) + e;
      if (FULL || (fr >= lo && fr < hi)) accumulate2(r2[fr + shift], g2, t2, acc[i * 2 + e], pkL, pkR);
    }
  }
}

template <int FPL>
__device__ __forceinline__ void consume_uni(const Desc& d, const uint8_t* row, float2 (&acc)[FPL], float& pkL,
                                            float& pkR, int lane) {
  if (d.lo == 0 && d.hi == 32 * FPL)  // whole tile: no per-frame range checks
    consume_uni_t<FPL, true>(d, row, acc, pkL, pkR, lane);
  else
    consume_uni_t<FPL, false>(d, row, acc, pkL, pkR, lane);
}

// Stereo f32, 2-tap linear resample (sample_linear<float, F32>, dsp/sampler.cpp:34-59) from the staged window.
// The position split avoids the slow f64<->int conversions: for 0 <= x < 2^31, t = x + 2^52 rounded TOWARDS
// -INF is exactly floor(x) + 2^52 (the ulp there is 1), so its low mantissa word is (int64_t)x and
// x - (t - 2^52) is x - (double)ix — both subtractions exact — as in sampler.cpp:51-52.
template <int FPL, bool FULL>
__device__ __forceinline__ void consume_lin_t(const Desc& d, const uint8_t* row, float2 (&acc)[FPL], float& pkL,
                                              float& pkR, int lane) {
  const float2* r2 = reinterpret_cast<const float2*>(row);
  const int lo = d.lo, hi = d.hi;
  const float2* rb = r2 - d.base;
  const double pos = d.pos, speed = d.speed;
  const double M = 4503599627370496.0;  // 2^52
  const double jj0 = (double)(d.jrel0 + 2 * lane);
  const float2 g2 = make_float2(d.gain, d.gain);
  const float2 t2 = make_float2(d.tg[0], d.tg[1]);
  const float2 neg1 = make_float2(-1.0f, -1.0f);
#pragma unroll
  for (int i = 0; i < FPL / 2; i++) {
    float2 term[2];
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int fr = 2 * (lane + 32 * i) + e;
      term[e] = make_float2(0.0f, 0.0f);
      if (FULL || (fr >= lo && fr < hi)) {
        const double jj = __dadd_rn(jj0, (double)(64 * i + e));  // exact small integers == (double)j
        const double x = __dadd_rn(pos, __dmul_rn(jj, speed));   // sampler.cpp:50
        const double t = __dadd_rd(x, M);                        // floor(x) + 2^52
        const int ix = __double2loint(t);                        // (int64_t)x, :51
        const float fx = __double2float_rn(__dsub_rn(x, __dsub_rn(t, M)));  // (float)(x - (double)ix), :52
        const float2 a = rb[ix], b = rb[ix + 1];
        const float2 df = __ffma2_rn(a, neg1, b);  // b - a (a * -1 is exact: one rounding)
        // a + fx * (b - a), :55 — scalar _rn ops on purpose: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into
        // FFMA2 when the product has no other use, even under --fmad false (tests/test_host_cpu.py lints SASS)
        float2 sv;
        sv.x = __fadd_rn(a.x, __fmul_rn(fx, df.x));
        sv.y = __fadd_rn(a.y, __fmul_rn(fx, df.y));
        term[e] = __fmul2_rn(__fmul2_rn(sv, g2), t2);
        acc[i * 2 + e] = __fadd2_rn(acc[i * 2 + e], term[e]);
      }
    }
    pkL = fmaxf(fmaxf(pkL, fabsf(term[0].x)), fabsf(term[1].x));
    pkR = fmaxf(fmaxf(pkR, fabsf(term[0].y)), fabsf(term[1].y));
  }
}

template <int FPL>
__device__ __forceinline__ void consume_lin(const Desc& d, const uint8_t* row, float2 (&acc)[FPL], float& pkL,
                                            float& pkR, int lane) {
  if (d.lo == 0 && d.hi == 32 * FPL)  // whole tile: no per-frame range checks
    consume_lin_t<FPL, true>(d, row, acc, pkL, pkR, lane);
  else
    consume_lin_t<FPL, false>(d, row, acc, pkL, pkR, lane);
}

// Stereo f32, polyphase windowed-sinc resample (extension) from the staged window: the lin path's position split,
// then 16 taps per frame.
template <int FPL>
__device__ __forceinline__ void consume_poly(const Desc& d, const uint8_t* row, const float* __restrict__ poly,
                                             float2 (&acc)[FPL], float& pkL, float& pkR, int lane) {
  const float2* rb = reinterpret_cast<const float2*>(row) - d.base;
  const int lo = d.lo, hi = d.hi;
  const double pos = d.pos, speed = d.speed;
  const double M = 4503599627370496.0;  // 2^52
  const double jj0 = (double)(d.jrel0 + 2 * lane);
  const float2 g2 = make_float2(d.gain, d.gain);
  const float2 t2 = make_float2(d.tg[0], d.tg[1]);
#pragma unroll
  for (int i = 0; i < FPL / 2; i++) {
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int fr = 2 * (lane + 32 * i) + e;
      if (fr >= lo && fr < hi) {
        const double jj = __dadd_rn(jj0, (double)(64 * i + e));
        const double x = __dadd_rn(pos, __dmul_rn(jj, speed));
        const double t = __dadd_rd(x, M);  // floor(x) + 2^52 (x >= 0)
        const int ix = __double2loint(t);
        const double fd = __dsub_rn(x, __dsub_rn(t, M));
        const float2 sv = poly_frame(poly, rb, (int64_t)ix, fd);
        accumulate2(sv, g2, t2, acc[i * 2 + e], pkL, pkR);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// the mix kernel
// ---------------------------------------------------------------------------------------------------------
template <int FPL, int STAGES>
struct MixLayout {
  static constexpr int T = 32 * FPL;                  // frames per tile
  static constexpr int STAGE_BYTES = (T + 32) * 8;    // staged window: T frames of stereo f32 + taps/alignment
  static constexpr int BATCH = 16;                    // cells resolved at a time
  static constexpr int RING = 2 * BATCH;              // descriptor ring entries
  static constexpr int OFF_DESC = STAGES * STAGE_BYTES;
  static constexpr int OFF_BAR = OFF_DESC + RING * (int)sizeof(Desc);
  static constexpr int WARP_BYTES = ((OFF_BAR + STAGES * 8) + 127) & ~127;
};

__device__ __forceinline__ DCell load_cell(const DCell* cells, uint32_t ci, uint32_t n_cells) {
  DCell c;
  c.pos = 0.0;
  c.span = kSilent;
  c.n_act = 0;
  if (ci < n_cells) {
    const int4 v = __ldg(reinterpret_cast<const int4*>(cells + ci));
    c.pos = __hiloint2double(v.y, v.x);
    c.span = (uint32_t)v.z;
    c.n_act = (uint32_t)v.w;
  }
  return c;
}

__device__ __forceinline__ DSpan load_span(const DSpan* spans, const DCell& c) {
  DSpan s;
  int4* q = reinterpret_cast<int4*>(&s);
  if (c.span != kSilent) {
    const int4* p = reinterpret_cast<const int4*>(spans + c.span);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(DSpan) / 16); i++) q[i] = __ldg(p + i);
  } else {
#pragma unroll
    for (int i = 0; i < (int)(sizeof(DSpan) / 16); i++) q[i] = make_int4(0, 0, 0, 0);
  }
  return s;
}

// Turn (cell, span) into the per-tile descriptor: which frames, which source window, which code path.
template <int STAGE_BYTES, int T>
__device__ __forceinline__ void resolve_store(const DCell& c, const DSpan& s, const float* __restrict__ gains,
                                              int f0, int tile_len, bool two, uint32_t k, Desc* out) {
  Desc d;
  d.src = nullptr;
  d.pos = c.pos;
  d.speed = s.speed;
  d.gain = s.gain;
  d.tg[0] = 0.f;
  d.tg[1] = 0.f;
  d.track = s.track;
  d.base = 0;
  d.jrel0 = 0;
  d.lo = 0;
  d.hi = 0;
  d.bytes = 0;
  d.kind = K_SILENT;
  d.fmt = (uint8_t)(s.fmt | (s.nch == 1 ? 0x80u : 0u));
  d.span = c.span;
  d.block_in_run = k - s.block0;
  if (c.span != kSilent) {
    const int seg_lo = (int)s.dst_off, seg_hi = (int)(s.dst_off + c.n_act);
    const int lo = (seg_lo > f0 ? seg_lo : f0) - f0;
    const int hi = (seg_hi < f0 + tile_len ? seg_hi : f0 + tile_len) - f0;
    if (hi > lo) {
      d.tg[0] = __ldg(gains + 2 * s.track);
      d.tg[1] = __ldg(gains + 2 * s.track + 1);
      d.jrel0 = f0 - seg_lo;
      d.lo = (uint16_t)lo;
      d.hi = (uint16_t)hi;
      const int fbytes = (int)s.nch * ((s.fmt == F_I16) ? 2 : 4);  // bytes per frame on the device
      const int64_t jj_lo = lo + d.jrel0, jj_hi = hi - 1 + d.jrel0;
      const bool unity = (s.speed == 1.0);
      // polyphase quality mode applies to stereo f32 sources on a stereo bus at speed != 1 (include/wbx.h)
      const bool poly = !unity && (s.fade & 2u) && two && s.fmt == F_F32 && s.nch == 2;
      if (poly) d.fmt |= 0x40u;
      int64_t first, last;  // first / last source frame the item can touch
      if (unity) {
        const int64_t ip = (int64_t)(uint32_t)(int64_t)c.pos;
        first = ip + jj_lo;
        last = ip + jj_hi;
      } else {  // conservative superset of [floor(x_lo), floor(x_hi) + 1] (+ the 16-tap reach in polyphase mode)
        first = (int64_t)(c.pos + (double)jj_lo * s.speed) - (poly ? 8 : 1);
        last = (int64_t)(c.pos + (double)jj_hi * s.speed) + (poly ? 9 : 2);
      }
      if (first < 0 && !poly) first = 0;  // polyphase taps may reach into the zero frames before the sample
      const int64_t align = 16 / fbytes;  // frames per 16 bytes: 2 (stereo f32), 4 (mono f32 / stereo i16), 8
      const int64_t a = first & ~(align - 1);
      const int64_t end = (last + align) & ~(align - 1);
      const int64_t bytes = (end - a) * fbytes;
      // fade extension: does a ramp overlap the frames of this item?
      bool fading = false;
      if (s.fade & 1u) {
        const double n_lo = s.clip_frame + (double)d.block_in_run * (double)s.length + (double)jj_lo;
        const double n_hi = s.clip_frame + (double)d.block_in_run * (double)s.length + (double)jj_hi;
        fading = (s.fade_in > 0.0 && n_lo < s.fade_in) || (s.fade_out > 0.0 && s.clip_len - n_hi < s.fade_out);
      }
      if (bytes <= STAGE_BYTES) {
        d.src = (const uint8_t*)s.base + a * fbytes;
        d.base = (int32_t)a;
        d.bytes = (uint16_t)bytes;
        const bool st32 = two && s.fmt == F_F32 && s.nch == 2;  // stereo f32 source into a stereo bus
        if (fading)
          d.kind = K_FADE;
        else if (poly)
          d.kind = K_POLY;
        else if (st32 && unity)
          d.kind = (lo == 0 && hi == T && first == a) ? K_FAST : K_UNI;
        else if (st32 && last < (int64_t)0x3fffffff)
          d.kind = K_LIN;
        else
          d.kind = K_GEN;
      } else {  // window larger than a stage (speed well above 1): read the source straight from global
        d.src = s.base;
        d.kind = fading ? K_DIRECT_FADE : K_DIRECT;
      }
    }
  }
  const int4* q = reinterpret_cast<const int4*>(&d);
  int4* o = reinterpret_cast<int4*>(out);
  o[0] = q[0];
  o[1] = q[1];
  o[2] = q[2];
  o[3] = q[3];
}

template <int FPL, int STAGES, int WARPS, bool EXT>
__global__ void __launch_bounds__(WARPS * 32) mix_kernel(const MixParams p) {
  using L = MixLayout<FPL, STAGES>;
  constexpr int BATCH = L::BATCH;
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* wbase = smem + (size_t)warp * L::WARP_BYTES;
  Desc* ring = reinterpret_cast<Desc*>(wbase + L::OFF_DESC);
  const uint32_t rows_s = smem_u32(wbase);
  const uint32_t bars_s = smem_u32(wbase + L::OFF_BAR);

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; s++) mbar_init(bars_s + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();

  const bool two = (p.C == 2);
  const uint32_t N = p.n_tracks, S = p.slots;
  uint32_t n_issued = 0, n_consumed = 0;  // staged items, monotonic over the kernel: stage = n % STAGES

  for (;;) {
    uint32_t w = 0;
    if (lane == 0) w = atomicAdd(&p.counters[0], 1u);
    w = __shfl_sync(0xffffffffu, w, 0);
    if (w >= p.n_items) break;
    const uint32_t g = w % p.groups;
    const uint32_t f = (w / p.groups) % p.n_tiles;
    const uint32_t k = w / (p.groups * p.n_tiles);
    const uint32_t tb = g * p.tracks_per_group;
    const uint32_t te = (tb + p.tracks_per_group < N) ? tb + p.tracks_per_group : N;
    const uint32_t n_cells = (te - tb) * S;
    const DCell* cells = p.cells + ((size_t)k * N + tb) * S;
    const int f0 = (int)(f * L::T);
    const int tile_len = ((int)p.B - f0 < L::T) ? (int)p.B - f0 : L::T;
    float* peaks_k = p.peaks + (size_t)k * N * 2;

    float2 acc[FPL];
#pragma unroll
    for (int i = 0; i < FPL; i++) acc[i] = make_float2(0.0f, 0.0f);
    float pkL = 0.0f, pkR = 0.0f;
    bool active = false;
    uint32_t cur_track = 0;
    uint32_t slot_ctr = 0;  // position of the current cell within its track's `slots` cells

    // descriptors of the first batch (lanes >= BATCH idle)
    if (lane < BATCH) {
      const DCell c0 = load_cell(cells, lane, n_cells);
      const DSpan s0 = load_span(p.spans, c0);
      resolve_store<L::STAGE_BYTES, L::T>(c0, s0, p.gains, f0, tile_len, two, k, &ring[lane]);
    }
    __syncwarp();
    uint32_t limit = BATCH;  // cells [0, limit) have descriptors
    uint32_t ip = 0;         // next cell to consider for staging
    const uint32_t nb = (n_cells + BATCH - 1) / BATCH;
    DCell cN;
    DSpan sN;
    cN.pos = 0.0;
    cN.span = kSilent;
    cN.n_act = 0;

    // stage the window of cell `ip` if it needs one (lane 0 issues the bulk copy)
    auto produce = [&]() {
      const uint32_t lim = limit < n_cells ? limit : n_cells;
      while (ip < lim && (n_issued - n_consumed) < (uint32_t)STAGES) {
        const Desc* dd = &ring[ip & (L::RING - 1)];
        const uint32_t kind = dd->kind;
        if (kind != K_SILENT && kind != K_DIRECT && kind != K_DIRECT_FADE) {
          if (lane == 0) {
            const uint32_t st = n_issued % STAGES;
            const uint32_t bar = bars_s + 8 * st;
            const uint32_t bytes = dd->bytes;
            mbar_expect_tx(bar, bytes);
            bulk_g2s(rows_s + st * L::STAGE_BYTES, dd->src, bytes, bar);
          }
          n_issued++;
        }
        ip++;
      }
    };

    for (uint32_t b = 0; b < nb; b++) {
      const bool more = (b + 1 < nb);
      if (more && lane < BATCH) cN = load_cell(cells, (b + 1) * BATCH + lane, n_cells);
#pragma unroll 1
      for (int i = 0; i < BATCH; i++) {
        const uint32_t ci = b * BATCH + i;
        if (ci >= n_cells) break;
        if (more && i == BATCH / 4 && lane < BATCH) sN = load_span(p.spans, cN);
        if (more && i == BATCH / 2) {
          if (lane < BATCH)
            resolve_store<L::STAGE_BYTES, L::T>(cN, sN, p.gains, f0, tile_len, two, k, &ring[((b + 1) & 1) * BATCH + lane]);
          __syncwarp();
          limit += BATCH;
        }
        produce();
        // ---- consumer role ------------------------------------------------------------------------------
        const Desc* dp = &ring[ci & (L::RING - 1)];
        const uint32_t kind = dp->kind;
        if (kind != K_SILENT) {
          active = true;
          cur_track = dp->track;
          const uint32_t st = n_consumed % STAGES;
          const bool staged = (kind != K_DIRECT && kind != K_DIRECT_FADE);
          if (staged) {
            const uint32_t par = (n_consumed / STAGES) & 1u;
            while (!mbar_try_wait(bars_s + 8 * st, par)) {
            }
          }
          const uint8_t* row = wbase + (size_t)st * L::STAGE_BYTES;
          if (kind == K_FAST) {
            consume_fast<FPL>(row, dp->gain, dp->tg[0], dp->tg[1], acc, pkL, pkR, lane);
          } else if (kind == K_UNI) {
            consume_uni<FPL>(*dp, row, acc, pkL, pkR, lane);
          } else if (kind == K_LIN) {
            consume_lin<FPL>(*dp, row, acc, pkL, pkR, lane);
          } else if (EXT && kind == K_POLY) {
            consume_poly<FPL>(*dp, row, p.poly, acc, pkL, pkR, lane);
          } else {
            const Desc d = *dp;
            consume_gen<FPL, EXT>(d, staged ? (const void*)row : d.src, acc, pkL, pkR, lane, two, p.spans, p.poly);
          }
          if (staged) {
            __syncwarp();  // every lane is done reading the stage before lane 0 may refill it
            n_consumed++;
          }
        }
        // ---- VU block peak once the track's last slot is done (vu_meter.h:20-30) ----------------------
        if (++slot_ctr == S) {
          slot_ctr = 0;
          if (active) {
            // lanes 0-15 reduce L, lanes 16-31 reduce R: one exchange, then four butterfly steps
            const float send = (lane < 16) ? pkR : pkL;
            const float recv = __shfl_xor_sync(0xffffffffu, send, 16);
            float m = fmaxf((lane < 16) ? pkL : pkR, recv);
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
            if ((lane & 15) == 0 && (lane == 0 || two)) {
              float* dst = peaks_k + (size_t)cur_track * 2 + (lane >> 4);
              if (p.n_tiles == 1)
                *dst = m;
              else if (m > 0.0f)
                atomicMax(reinterpret_cast<unsigned int*>(dst), __float_as_uint(m));  // m >= 0: uint order == float order
            }
            pkL = 0.0f;
            pkR = 0.0f;
            active = false;
          }
        }
      }
    }

    // ---- bus write ------------------------------------------------------------------------------------
    // destination of channel c of this tile: the planar device bus, or — sharded render — this rank's plane in the
    // exchange buffer of the rank that owns callback k (a peer-memory store over NVLink, overlapped with the mix)
    float* outp[2];
    if (p.shard_blocks) {
      const uint32_t owner = k / p.shard_blocks;
      const size_t plane = (size_t)p.shard_blocks * p.B;
      float* xb = p.xchg[owner] + (size_t)p.shard_rank * p.C * plane + (size_t)(k - owner * p.shard_blocks) * p.B + f0;
      outp[0] = xb;
      outp[1] = xb + plane;
    } else {
      const size_t chan_stride = (size_t)p.n_blocks * p.B;
      outp[0] = p.bus + (size_t)k * p.B + f0;
      outp[1] = outp[0] + chan_stride;
    }
    const size_t out_off = (size_t)k * p.B + f0;
    const bool vec_ok = (p.B & 1u) == 0;  // frame pairs are 8-byte aligned in the planar bus
    if (p.groups == 1) {
#pragma unroll
      for (int i = 0; i < FPL / 2; i++) {
        const int fr = 2 * (lane + 32 * i);
        float v[2][2] = {{acc[i * 2].x, acc[i * 2 + 1].x}, {acc[i * 2].y, acc[i * 2 + 1].y}};
#pragma unroll
        for (int c = 0; c < 2; c++) {
          if (c == 1 && !two) break;
          float* out = outp[c];
          float x0 = v[c][0], x1 = v[c][1];
          if (p.clamp) {  // engine.cpp:1627-1636 (NaN passes)
            x0 = x0 > 1.0f ? 1.0f : (x0 < -1.0f ? -1.0f : x0);
            x1 = x1 > 1.0f ? 1.0f : (x1 < -1.0f ? -1.0f : x1);
          }
          float* mir = p.mirror[c] ? p.mirror[c] + out_off : nullptr;
          if (vec_ok && fr + 1 < tile_len) {
            *reinterpret_cast<float2*>(out + fr) = make_float2(x0, x1);
            if (mir) *reinterpret_cast<float2*>(mir + fr) = make_float2(x0, x1);
          } else {
            if (fr < tile_len) out[fr] = x0;
            if (fr + 1 < tile_len) out[fr + 1] = x1;
            if (mir) {
              if (fr < tile_len) mir[fr] = x0;
              if (fr + 1 < tile_len) mir[fr + 1] = x1;
            }
          }
        }
      }
    } else {
      // tree mode: publish this group's partial, the last group to arrive adds them in group order
      const uint32_t tile_id = k * p.n_tiles + f;
      float2* part = reinterpret_cast<float2*>(p.ws) + ((size_t)tile_id * p.groups + g) * 2 * (L::T / 2);
#pragma unroll
      for (int i = 0; i < FPL / 2; i++) {
        const int q = lane + 32 * i;
        part[q] = make_float2(acc[i * 2].x, acc[i * 2 + 1].x);               // channel 0, frames 2q, 2q+1
        part[(L::T / 2) + q] = make_float2(acc[i * 2].y, acc[i * 2 + 1].y);  // channel 1
      }
      __threadfence();
      __syncwarp();
      uint32_t prev = 0;
      if (lane == 0) prev = atomicAdd(&p.counters[1 + tile_id], 1u);
      prev = __shfl_sync(0xffffffffu, prev, 0);
      if (prev == p.groups - 1) {
        __threadfence();
        const float2* base = reinterpret_cast<const float2*>(p.ws) + (size_t)tile_id * p.groups * 2 * (L::T / 2);
        // This warp is the tail of the whole render, and every load below is an L2 round trip: keep many in flight. All
        // (channel, frame-pair) combinations of GB groups are loaded together, then added per output element in group
        // order (the association is unchanged: partial 0, 1, 2, ... for every element).
        // (512-frame tiles keep one group per batch: tree order at that tile size is a corner case and the hot exact-order
        // path of that instantiation must not lose registers to it)
        constexpr int NP = FPL / 2;                      // frame pairs per lane and channel
        constexpr int GB = FPL <= 8 ? 32 / (2 * NP) : 1;  // groups per batch: <= 32 float2 in flight
        const int nc = two ? 2 : 1;
        float2 sum[2][NP];
#pragma unroll
        for (int c = 0; c < 2; c++)
#pragma unroll
          for (int i = 0; i < NP; i++) sum[c][i] = make_float2(0.f, 0.f);
        for (uint32_t g0 = 0; g0 < p.groups; g0 += GB) {
          float2 v[GB][2][NP];
#pragma unroll
          for (int u = 0; u < GB; u++)
#pragma unroll
            for (int c = 0; c < 2; c++)
#pragma unroll
              for (int i = 0; i < NP; i++)
                v[u][c][i] = (g0 + u < p.groups && c < nc)
                                 ? __ldcg(base + ((size_t)(g0 + u) * 2 + c) * (L::T / 2) + lane + 32 * i)
                                 : make_float2(0.f, 0.f);
#pragma unroll
          for (int u = 0; u < GB; u++) {
            if (g0 + u < p.groups) {
#pragma unroll
              for (int c = 0; c < 2; c++)
#pragma unroll
                for (int i = 0; i < NP; i++) {
                  sum[c][i].x = __fadd_rn(sum[c][i].x, v[u][c][i].x);
                  sum[c][i].y = __fadd_rn(sum[c][i].y, v[u][c][i].y);
                }
            }
          }
        }
#pragma unroll
        for (int c = 0; c < 2; c++) {
          if (c >= nc) break;
          float* out = outp[c];
          float* mir = p.mirror[c] ? p.mirror[c] + out_off : nullptr;
#pragma unroll
          for (int i = 0; i < NP; i++) {
            const int fr = 2 * (lane + 32 * i);
            float2 r = sum[c][i];
            if (p.clamp) {
              r.x = r.x > 1.0f ? 1.0f : (r.x < -1.0f ? -1.0f : r.x);
              r.y = r.y > 1.0f ? 1.0f : (r.y < -1.0f ? -1.0f : r.y);
            }
            if (vec_ok && fr + 1 < tile_len) {
              *reinterpret_cast<float2*>(out + fr) = r;
              if (mir) *reinterpret_cast<float2*>(mir + fr) = r;
            } else {
              if (fr < tile_len) out[fr] = r.x;
              if (fr + 1 < tile_len) out[fr + 1] = r.y;
              if (mir) {
                if (fr < tile_len) mir[fr] = r.x;
                if (fr + 1 < tile_len) mir[fr + 1] = r.y;
              }
            }
          }
        }
      }
    }
  }
}

// (frames, channels) planar -> frame-interleaved device sample layout, keeping the first `nch` channels
__global__ void interleave_sample_kernel(const uint8_t* __restrict__ planar, size_t plane_bytes, uint64_t frames,
                                         uint32_t nch, uint32_t esize, uint8_t* __restrict__ dst) {
  const uint64_t n = frames * nch;
  for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t f = o / nch;
    const uint32_t c = (uint32_t)(o % nch);
    if (esize == 4)
      reinterpret_cast<uint32_t*>(dst)[o] = reinterpret_cast<const uint32_t*>(planar + c * plane_bytes)[f];
    else
      reinterpret_cast<uint16_t*>(dst)[o] = reinterpret_cast<const uint16_t*>(planar + c * plane_bytes)[f];
  }
}

// ---------------------------------------------------------------------------------------------------------
// effect chain (extension): render the tracks that carry a chain into a per-track buffer, run the chain
// sequentially in time, then let the mix kernel read that buffer like any other resident sample
// ---------------------------------------------------------------------------------------------------------

// value one Sampler::stream call adds at segment-relative frame jj for output channel c (generic, from global)
__device__ __forceinline__ float stream_value(const DSpan& sp, const DCell& cell, int32_t jj, uint32_t c,
                                              uint32_t block_in_run, const float* poly, bool two) {
  const uint32_t nch = sp.nch;
  const uint32_t ch = c % nch;
  float sv;
  if (sp.speed == 1.0) {
    const int64_t ip = (int64_t)(uint32_t)(int64_t)cell.pos;
    sv = load_unity_rt(sp.fmt, sp.base, (ip + jj) * nch + ch);
  } else {
    const double x = __dadd_rn(cell.pos, __dmul_rn((double)jj, sp.speed));
    const int64_t ix = __double2ll_rz(x);
    if ((sp.fade & 2u) && two && sp.fmt == F_F32 && nch == 2 && poly) {
      const float2 pv = poly_frame(poly, reinterpret_cast<const float2*>(sp.base), ix, __dsub_rn(x, __ll2double_rn(ix)));
      sv = c ? pv.y : pv.x;
    } else {
      const float fx = __double2float_rn(__dsub_rn(x, __ll2double_rn(ix)));
      const float a = load_lin_rt(sp.fmt, sp.base, ix * nch + ch), b = load_lin_rt(sp.fmt, sp.base, (ix + 1) * nch + ch);
      sv = __fadd_rn(a, __fmul_rn(fx, __fsub_rn(b, a)));
    }
  }
  float m = __fmul_rn(sv, sp.gain);
  if (sp.fade & 1u) {
    FadeEnv fe;
    fe.n0 = sp.clip_frame + (double)block_in_run * (double)sp.length;
    fe.fin = sp.fade_in;
    fe.fout = sp.fade_out;
    fe.len = sp.clip_len;
    m = __fmul_rn(m, fe.at(jj));
  }
  return m;
}

// one warp per (effect track e, callback k): the track's mixing buffer before effects, frame-interleaved stereo
__global__ void render_tracks_kernel(const DSpan* __restrict__ spans, const DCell* __restrict__ cells,
                                     const DFx* __restrict__ fx, uint32_t n_fx, uint32_t N, uint32_t S, uint32_t K,
                                     uint32_t B, uint32_t C, const float* __restrict__ poly, float* __restrict__ trackbuf) {
  const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (w >= (uint64_t)n_fx * K) return;
  const uint32_t e = (uint32_t)(w / K), k = (uint32_t)(w % K);
  const uint32_t t = fx[e].track;
  float2* out = reinterpret_cast<float2*>(trackbuf) + ((size_t)e * K + k) * B;
  if (S == 1) {  // one Sampler::stream call per callback (the steady state): cell and span read once per warp
    const DCell cell = cells[(size_t)k * N + t];
    if (cell.span == kSilent) {
      for (uint32_t j = lane; j < B; j += 32) out[j] = make_float2(0.0f, 0.0f);
      return;
    }
    const DSpan sp = spans[cell.span];
    const uint32_t lo = sp.dst_off, hi = sp.dst_off + cell.n_act;
    if (sp.fmt == F_F32 && sp.nch == 2 && sp.speed == 1.0 && sp.fade == 0 && C == 2) {
      // unity-speed stereo f32 clip: a scaled copy — the same rounded operations as stream_value (src * gain, then the
      // add into the cleared mixing buffer: 0 + m)
      const float2* src = reinterpret_cast<const float2*>(sp.base) + (int64_t)(uint32_t)(int64_t)cell.pos;
      for (uint32_t j = lane; j < B; j += 32) {
        float2 v = make_float2(0.0f, 0.0f);
        if (j >= lo && j < hi) {
          const float2 x = src[j - lo];
          v.x = __fadd_rn(0.0f, __fmul_rn(x.x, sp.gain));
          v.y = __fadd_rn(0.0f, __fmul_rn(x.y, sp.gain));
        }
        out[j] = v;
      }
      return;
    }
    for (uint32_t j = lane; j < B; j += 32) {
      float2 v = make_float2(0.0f, 0.0f);
      if (j >= lo && j < hi) {
        const int32_t jj = (int32_t)(j - lo);
        v.x = __fadd_rn(v.x, stream_value(sp, cell, jj, 0, k - sp.block0, poly, C == 2));
        if (C == 2) v.y = __fadd_rn(v.y, stream_value(sp, cell, jj, 1, k - sp.block0, poly, true));
      }
      out[j] = v;
    }
    return;
  }
  for (uint32_t j = lane; j < B; j += 32) {
    float2 v = make_float2(0.0f, 0.0f);
    for (uint32_t s = 0; s < S; s++) {
      const DCell cell = cells[((size_t)k * N + t) * S + s];
      if (cell.span == kSilent) continue;
      const DSpan sp = spans[cell.span];
      if (j >= sp.dst_off && j < sp.dst_off + cell.n_act) {
        const int32_t jj = (int32_t)(j - sp.dst_off);
        v.x = __fadd_rn(v.x, stream_value(sp, cell, jj, 0, k - sp.block0, poly, C == 2));  // dst += ... on a cleared buffer
        if (C == 2) v.y = __fadd_rn(v.y, stream_value(sp, cell, jj, 1, k - sp.block0, poly, true));
      }
    }
    out[j] = v;
  }
}

// The chain is a recurrence in time: one THREAD per (effect track, channel) carries the state in registers and
// walks the whole render, 16 frames at a time so the loads of a chunk are in flight together. Same operations, same order as oracle/wb_oracle.c apply_effects (every op a single IEEE rn op;
// __fmaf_rn = fmaf). Time-parallel (scan) evaluation of the biquads would re-associate and is left for later.
struct FxChannel {
  float s1[4], s2[4], env;
};

__device__ __forceinline__ float fx_sample(float x, FxChannel& st, const float (&b0)[4], const float (&b1)[4],
                                           const float (&b2)[4], const float (&a1)[4], const float (&a2)[4], bool eq_on,
                                           bool comp_on, float thr, float att, float rel, float makeup, uint32_t code) {
  if (eq_on) {
#pragma unroll
    for (int b = 0; b < 4; b++) {  // transposed direct form II
      const float y = __fmaf_rn(b0[b], x, st.s1[b]);
      st.s1[b] = __fmaf_rn(b1[b], x, __fmaf_rn(-a1[b], y, st.s2[b]));
      st.s2[b] = __fmaf_rn(b2[b], x, -__fmul_rn(a2[b], y));
      x = y;
    }
  }
  if (comp_on) {
    const float xa = fabsf(x);
    st.env = xa > st.env ? __fmaf_rn(att, __fsub_rn(st.env, xa), xa) : __fmaf_rn(rel, __fsub_rn(st.env, xa), xa);
    float g = 1.0f;
    if (st.env > thr) {
      const float r = __fdiv_rn(thr, st.env);
      const float r2 = __fsqrt_rn(r);
      switch (code) {
        case 1: g = r2; break;
        case 2: g = __fmul_rn(r2, __fsqrt_rn(r2)); break;
        case 3: g = __fmul_rn(__fmul_rn(r2, __fsqrt_rn(r2)), __fsqrt_rn(__fsqrt_rn(r2))); break;
        default: g = r; break;
      }
    }
    x = __fmul_rn(__fmul_rn(x, g), makeup);
  }
  return x;
}

// The memoryless part of the compressor: gain from the envelope, applied with the make-up gain.
__device__ __forceinline__ float fx_gain(float x, float env, float thr, float makeup, uint32_t code) {
  float g = 1.0f;
  if (env > thr) {
    const float r = __fdiv_rn(thr, env);
    const float r2 = __fsqrt_rn(r);
    switch (code) {
      case 1: g = r2; break;
      case 2: g = __fmul_rn(r2, __fsqrt_rn(r2)); break;
      case 3: g = __fmul_rn(__fmul_rn(r2, __fsqrt_rn(r2)), __fsqrt_rn(__fsqrt_rn(r2))); break;
      default: g = r; break;
    }
  }
  return __fmul_rn(__fmul_rn(x, g), makeup);
}

// The chain is five recurrences in series (4 biquads, the envelope follower) and one memoryless map (gain computer).
// A thread that walks them sample by sample is bound by the dependent-FMA latency of all five in a row while the GPU
// idles. Here one WARP carries P (track, channel) pairs as a software pipeline over 32-frame chunks held in shared
// memory: lane p*4+s runs biquad s of pair p on chunk i-s, lane p also runs pair p's envelope follower on chunk i-4 —
// both recurrences sit in the same instruction stream, so their latencies overlap — and all 32 lanes then evaluate the
// gain computer of chunk i-5 one frame per lane and store it. Every frame still sees exactly the operations of
// fx_sample in the same order (a stage's input is the previous stage's rounded output either way): bit-identical to
// the one-thread walk and to oracle/wb_oracle.c apply_effects.
template <int P>
struct FxLayout {
  static constexpr int DEPTH = 4;                      // chunk slots per stage buffer (writer and readers 2 apart)
  // 32 frames + pad. P <= 4: 36 — rows stay 16-B aligned for 128-bit accesses and the rows a quarter-warp touches
  // start 4 banks apart (measured 16.0 -> 14.5 ms at 512 tracks). P = 8: 33 — with all 32 lanes active the 36-float
  // stride is a 4-way bank conflict (measured 11.7 -> 13.2 ms at 4096 tracks), so that size keeps scalar accesses with
  // one row per bank.
  static constexpr int ROW = P <= 4 ? 36 : 33;
  static constexpr bool VEC = ROW % 4 == 0;
  static constexpr int ROWS = (5 * DEPTH) * P + DEPTH * P;  // X[0..4], E
  static constexpr int WARP_FLOATS = ROWS * ROW;
  __device__ static int x_row(int stage, int slot, int pair) { return (stage * DEPTH + slot) * P + pair; }
  __device__ static int e_row(int slot, int pair) { return (5 * DEPTH + slot) * P + pair; }
  static constexpr int IDLE = 0;  // row idle lanes point at (never dereferenced)
};

// T > 1: the P pairs are carried by a TEAM of T warps on different schedulers, meeting at a named barrier once per chunk;
// the chunk slots each role touches within an iteration are disjoint (same schedule as the one-warp form).
//   T = 2: warp 0 runs the recurrences, warp 1 the input loads, the gain computer and the stores;
//   T = 3: the envelope followers get their own warp (ptxas schedules the two recurrences of one warp mostly one after
//          the other instead of interleaving them, and moving the followers to the I/O warp just moves the long pole).
// Used while the session is too small to give every scheduler of the GPU a warp otherwise: the step is then bound by
// per-warp latency, not by issue slots.
template <int P, int T>
__global__ void __launch_bounds__(T == 3 ? 192 : 128) effects_kernel(DFx* __restrict__ fx, uint32_t n_fx, uint32_t C, uint64_t frames,
                                                       float* __restrict__ trackbuf) {
  using L = FxLayout<P>;
  extern __shared__ __align__(16) float fx_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr bool DUO = T > 1;  // (kept as a name for "several warps per team")
  const int team = warp / T, role = warp % T;
  const int teams_per_cta = (int)(blockDim.x >> 5) / T;
  const bool do_rec = T == 1 || role == 0;                       // this warp runs the biquads
  const bool do_io = T == 1 || role == 1;                        // this warp loads, computes gains and stores
  const bool do_env = T == 1 || (T == 2 ? role == 0 : role == 2);  // this warp runs the envelope followers
  float* sm = fx_smem + (size_t)team * L::WARP_FLOATS;
  const uint32_t n_pairs = n_fx * C;
  const uint32_t pair0 = (blockIdx.x * teams_per_cta + team) * P;
  if (pair0 >= n_pairs) return;  // both warps of a team leave together
  auto team_sync = [&]() {
    if (DUO)
      asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(32 * T) : "memory");
    else
      __syncwarp();
  };

  // biquad role: lane = p*4 + s
  const int bp = lane >> 2, bs = lane & 3;
  const bool bq_lane = bp < P && pair0 + bp < n_pairs;
  float b0 = 0.f, b1 = 0.f, b2 = 0.f, a1 = 0.f, a2 = 0.f, s1 = 0.f, s2 = 0.f;
  bool eq_on = false;
  if (bq_lane) {
    const uint32_t g = pair0 + bp;
    const DFx* f = fx + g / C;
    const uint32_t c = g % C;
    eq_on = f->eq_on != 0;
    b0 = f->b0[bs], b1 = f->b1[bs], b2 = f->b2[bs], a1 = f->a1[bs], a2 = f->a2[bs];
    s1 = f->s1[c][bs], s2 = f->s2[c][bs];
  }
  // envelope role: lane = p
  bool env_lane = lane < P && pair0 + lane < n_pairs;
  float env = 0.f, att = 0.f, rel = 0.f;
  if (env_lane) {
    const uint32_t g = pair0 + lane;
    const DFx* f = fx + g / C;
    env = f->env[g % C];
    att = f->att, rel = f->rel;
    env_lane = f->comp_on != 0;  // without a compressor the envelope state is left alone (fx_sample)
  }
  // gain role: every lane, pair r in turn — per-pair constants and the pair's channel in the interleaved track buffer
  float thr[P], makeup[P];
  uint32_t code[P];
  bool comp_on[P];
  float* buf[P];
#pragma unroll
  for (int r = 0; r < P; r++) {
    const uint32_t g = pair0 + r < n_pairs ? pair0 + r : n_pairs - 1;
    const DFx* f = fx + g / C;
    thr[r] = f->thr, makeup[r] = f->makeup, code[r] = f->ratio_code, comp_on[r] = f->comp_on != 0;
    buf[r] = trackbuf + (size_t)(g / C) * frames * 2 + (g % C);
  }

  // chunk counters are 32-bit: 2^32 chunks of 32 frames is 33 days of 48 kHz audio in one render
  const uint32_t n_chunks = (uint32_t)((frames + 31) / 32);
  const int last_len = (int)(frames - (uint64_t)(n_chunks ? n_chunks - 1 : 0) * 32);
  auto chunk_len = [&](int32_t c) -> int {
    if (c < 0 || (uint32_t)c >= n_chunks) return 0;
    return (uint32_t)c + 1 == n_chunks ? last_len : 32;
  };
  // input chunk c (one frame per lane) of pair r; the track buffer is larger than L2, so every chunk is an HBM round trip
  auto load_chunk = [&](uint32_t c, int r) -> float {
    const uint64_t f = (uint64_t)c * 32 + lane;
    return (do_io && pair0 + r < n_pairs && f < frames) ? buf[r][f * 2] : 0.0f;
  };
  // chunk 0 into X[0][0]; chunks 1 and 2 on their way (the loop keeps three chunks in flight in registers)
  float nxt0[P], nxt1[P];
  if (do_io) {
#pragma unroll
    for (int r = 0; r < P; r++)
      if (pair0 + r < n_pairs && (uint64_t)lane < frames) sm[L::x_row(0, 0, r) * L::ROW + lane] = buf[r][(size_t)lane * 2];
  }
#pragma unroll
  for (int r = 0; r < P; r++) {
    nxt0[r] = load_chunk(1, r);
    nxt1[r] = load_chunk(2, r);
  }
  team_sync();

  for (uint32_t i = 0; i < n_chunks + 5; i++) {
    // input of chunk i+3: issued now, stored into its slot two iterations from now — an HBM latency is longer than one
    // iteration of the recurrences
    float nxt2[P];
#pragma unroll
    for (int r = 0; r < P; r++) nxt2[r] = load_chunk(i + 3, r);

    // ---- recurrences: biquad s on chunk i-s, envelope on chunk i-4 -------------------------------------------
    const bool steady = i >= 5 && i + 1 < n_chunks;  // every role has a full chunk this iteration: no bounds work
    const int32_t cb = (int32_t)i - bs, ce = (int32_t)i - 4;
    const int nb = (bq_lane && do_rec) ? (steady ? 32 : chunk_len(cb)) : 0;
    const int ne = (env_lane && do_env) ? (steady ? 32 : chunk_len(ce)) : 0;
    const float* xin = sm + (nb ? L::x_row(bs, (int)(cb & 3), bp) : L::IDLE) * L::ROW;
    float* xout = sm + (nb ? L::x_row(bs + 1, (int)(cb & 3), bp) : L::IDLE) * L::ROW;
    const float* ein = sm + (ne ? L::x_row(4, (int)(ce & 3), lane) : L::IDLE) * L::ROW;
    float* eout = sm + (ne ? L::e_row((int)(ce & 3), lane) : L::IDLE) * L::ROW;
    // One steady-state pass over a full chunk for the roles this warp holds. Idle lanes run the same arithmetic on zeros
    // (their state is never stored): only the shared-memory accesses are predicated on the lane's role, the recurrences
    // carry no predicate. The chunk is staged in registers: a shared-memory load between dependent FMAs (the compiler
    // cannot move it above the previous frame's store) would put its latency into every step of the recurrence.
    auto steady_pass = [&](auto bq_tag, auto en_tag) {
      constexpr bool BQ = decltype(bq_tag)::value, EN = decltype(en_tag)::value;
      const bool bq = nb != 0, en = ne != 0;
      float xv[BQ ? 32 : 1], ev[EN ? 32 : 1];
      if constexpr (L::VEC) {
#pragma unroll
        for (int q = 0; q < 32; q += 4) {  // 128-bit shared-memory accesses: 8 loads instead of 32 per role
          if constexpr (BQ) {
            const float4 a = bq ? *reinterpret_cast<const float4*>(xin + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            xv[q] = a.x, xv[q + 1] = a.y, xv[q + 2] = a.z, xv[q + 3] = a.w;
          }
          if constexpr (EN) {
            const float4 e4 = en ? *reinterpret_cast<const float4*>(ein + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            ev[q] = e4.x, ev[q + 1] = e4.y, ev[q + 2] = e4.z, ev[q + 3] = e4.w;
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < 32; q++) {
          if constexpr (BQ) xv[q] = bq ? xin[q] : 0.0f;
          if constexpr (EN) ev[q] = en ? ein[q] : 0.0f;
        }
      }
#pragma unroll
      for (int q = 0; q < 32; q++) {
        if constexpr (BQ) {
          const float x = xv[q];
          const float y = __fmaf_rn(b0, x, s1);  // transposed direct form II (fx_sample)
          s1 = __fmaf_rn(b1, x, __fmaf_rn(-a1, y, s2));
          s2 = __fmaf_rn(b2, x, -__fmul_rn(a2, y));
          xv[q] = eq_on ? y : x;
        }
        if constexpr (EN) {
          const float xa = fabsf(ev[q]);
          const float d = __fsub_rn(env, xa);
          env = xa > env ? __fmaf_rn(att, d, xa) : __fmaf_rn(rel, d, xa);
          ev[q] = env;
        }
      }
      if constexpr (L::VEC) {
#pragma unroll
        for (int q = 0; q < 32; q += 4) {
          if constexpr (BQ)
            if (bq) *reinterpret_cast<float4*>(xout + q) = make_float4(xv[q], xv[q + 1], xv[q + 2], xv[q + 3]);
          if constexpr (EN)
            if (en) *reinterpret_cast<float4*>(eout + q) = make_float4(ev[q], ev[q + 1], ev[q + 2], ev[q + 3]);
        }
      } else {
#pragma unroll
        for (int q = 0; q < 32; q++) {
          if constexpr (BQ)
            if (bq) xout[q] = xv[q];
          if constexpr (EN)
            if (en) eout[q] = ev[q];
        }
      }
    };
    if (steady) {
      if (do_rec && do_env)
        steady_pass(std::true_type{}, std::true_type{});
      else if (do_rec)
        steady_pass(std::true_type{}, std::false_type{});
      else if (do_env)
        steady_pass(std::false_type{}, std::true_type{});
    } else {
      for (int q = 0; q < 32; q++) {
        if (q < nb) {
          const float x = xin[q];
          const float y = __fmaf_rn(b0, x, s1);
          s1 = __fmaf_rn(b1, x, __fmaf_rn(-a1, y, s2));
          s2 = __fmaf_rn(b2, x, -__fmul_rn(a2, y));
          xout[q] = eq_on ? y : x;
        }
        if (q < ne) {
          const float xa = fabsf(ein[q]);
          const float d = __fsub_rn(env, xa);
          env = xa > env ? __fmaf_rn(att, d, xa) : __fmaf_rn(rel, d, xa);
          eout[q] = env;
        }
      }
    }
    // ---- gain computer + store of chunk i-5, one frame per lane ------------------------------------------------
    const int32_t cg = (int32_t)i - 5;
    const int ng = do_io ? (steady ? 32 : chunk_len(cg)) : 0;
    if (lane < ng) {
#pragma unroll
      for (int r = 0; r < P; r++) {
        if (pair0 + r >= n_pairs) break;
        const float x = sm[L::x_row(4, (int)(cg & 3), r) * L::ROW + lane];
        const float e = sm[L::e_row((int)(cg & 3), r) * L::ROW + lane];
        buf[r][((size_t)(uint32_t)cg * 32 + lane) * 2] = comp_on[r] ? fx_gain(x, e, thr[r], makeup[r], code[r]) : x;
      }
    }
    // ---- stage the next input chunk -------------------------------------------------------------------------------
    if (do_io) {
#pragma unroll
      for (int r = 0; r < P; r++) sm[L::x_row(0, (int)((i + 1) & 3), r) * L::ROW + lane] = nxt0[r];
    }
#pragma unroll
    for (int r = 0; r < P; r++) {
      nxt0[r] = nxt1[r];
      nxt1[r] = nxt2[r];
    }
    team_sync();
  }

  if (do_rec && bq_lane && eq_on) {
    const uint32_t g = pair0 + bp;
    DFx* f = fx + g / C;
    f->s1[g % C][bs] = s1;
    f->s2[g % C][bs] = s2;
  }
  if (do_env && env_lane) {
    const uint32_t g = pair0 + lane;
    fx[g / C].env[g % C] = env;
  }
}

template <int P, int T>
static cudaError_t launch_effects_p(DFx* fx, uint32_t n_fx, uint32_t C, uint64_t frames, float* trackbuf, cudaStream_t stream) {
  constexpr int TEAMS = T == 1 ? 4 : 2;  // teams (= shared-memory regions) per CTA
  constexpr int WARPS = TEAMS * T;
  const size_t smem = (size_t)FxLayout<P>::WARP_FLOATS * TEAMS * sizeof(float);
  auto kfn = effects_kernel<P, T>;
  cudaError_t err = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  const uint32_t teams = (n_fx * C + P - 1) / P;
  kfn<<<(teams + TEAMS - 1) / TEAMS, WARPS * 32, smem, stream>>>(fx, n_fx, C, frames, trackbuf);
  return cudaGetLastError();
}

// Shape of the chain kernel. Small and medium sessions cannot fill the GPU's 4 schedulers per SM with one warp per P
// pairs, and are bound by per-warp latency: they run as three-warp teams with the fewest pairs per team that keep the
// total at ~1.5 warps per scheduler (measured at 512 tracks = 1024 pairs: P = 4 / 768 warps 11.3 ms, P = 2 / 1536 warps
// 12.5 ms, P = 1 / 3072 warps 16.9 ms; single warps 17.2 ms). Large sessions are instruction-issue-bound: one warp per 8
// pairs (measured at 4096 tracks: 10.0 ms, three-warp teams 13.1 ms).
static cudaError_t launch_effects_chain(DFx* fx, uint32_t n_fx, uint32_t C, uint64_t frames, float* trackbuf, int n_sm,
                                        cudaStream_t stream) {
  const uint32_t pairs = n_fx * C;
  const uint32_t slots = (uint32_t)n_sm * 4;  // warp schedulers of the GPU
  auto teams = [&](uint32_t p) { return (pairs + p - 1) / p; };
  int P = 4, T = 3;
  if (3 * teams(1) <= slots + slots / 2) P = 1;
  else if (3 * teams(2) <= slots + slots / 2) P = 2;
  if (3 * teams(4) > 3 * slots) {  // even 4 pairs per team would put more than ~3 warps on every scheduler
    T = 1;
    P = pairs <= 4 * slots ? 4 : 8;
  }
  if (const char* env = getenv("WBX_FX_PAIRS")) {
    const int v = atoi(env);
    if (v == 1 || v == 2 || v == 4 || v == 8) P = v;
  }
  if (const char* env = getenv("WBX_FX_TEAM")) {
    const int v = atoi(env);
    if (v >= 1 && v <= 3) T = v;
  }
#define WBX_FX_CASE(PP)                                                                                \
  case PP:                                                                                             \
    return T == 3   ? launch_effects_p<PP, 3>(fx, n_fx, C, frames, trackbuf, stream)                   \
           : T == 2 ? launch_effects_p<PP, 2>(fx, n_fx, C, frames, trackbuf, stream)                   \
                    : launch_effects_p<PP, 1>(fx, n_fx, C, frames, trackbuf, stream);
  switch (P) {
    WBX_FX_CASE(1)
    WBX_FX_CASE(2)
    WBX_FX_CASE(4)
    default: return T == 3   ? launch_effects_p<8, 3>(fx, n_fx, C, frames, trackbuf, stream)
                    : T == 2 ? launch_effects_p<8, 2>(fx, n_fx, C, frames, trackbuf, stream)
                             : launch_effects_p<8, 1>(fx, n_fx, C, frames, trackbuf, stream);
  }
#undef WBX_FX_CASE
}

// ---- convolution reverb (extension, cfg 5): direct form on the CUDA cores --------------------------------
// xin  [n_fx][C][H + T] planar: H = taps - 1 history frames (oldest first) followed by this render's T chain outputs
// hist [n_tracks][2][H] persists across renders (indexed by track, so it survives chain list rebuilds)
__global__ void fir_gather_kernel(const DFx* __restrict__ fx, uint32_t n_fx, uint32_t C, uint64_t H, uint64_t T,
                                  const float* __restrict__ hist, const float* __restrict__ trackbuf,
                                  float* __restrict__ xin) {
  const uint32_t ec = blockIdx.y;
  const uint32_t e = ec / C, c = ec % C;
  if (!fx[e].reverb_on) return;
  const float* h = hist + ((size_t)fx[e].track * 2 + c) * H;
  const float* tb = trackbuf + (size_t)e * T * 2 + c;
  float* x = xin + (size_t)ec * (H + T);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < H + T; i += (uint64_t)gridDim.x * blockDim.x)
    x[i] = i < H ? h[i] : tb[(i - H) * 2];
}

// 256 consecutive outputs of one (track, channel) per CTA; taps in tiles of 256 staged in shared memory
__global__ void __launch_bounds__(256) fir_kernel(const DFx* __restrict__ fx, uint32_t C, uint64_t H, uint64_t T,
                                                   const float* __restrict__ ir, uint32_t L,
                                                   const float* __restrict__ xin, float* __restrict__ trackbuf) {
  __shared__ __align__(16) float hs[256];
  __shared__ float xs[512];
  const uint32_t ec = blockIdx.y;
  const uint32_t e = ec / C, c = ec % C;
  if (!fx[e].reverb_on) return;
  const uint32_t tid = threadIdx.x;
  const int64_t n0 = (int64_t)blockIdx.x * 256;
  const float* x = xin + (size_t)ec * (H + T) + H;  // x[n - k], n - k >= -H
  double total = 0.0;
  for (uint32_t k0 = 0; k0 < L; k0 += 256) {
    __syncthreads();
    hs[tid] = (k0 + tid < L) ? __ldg(ir + k0 + tid) : 0.0f;
    for (uint32_t i = tid; i < 511; i += 256) {  // xs[i] = x[n0 - k0 - 255 + i]
      const int64_t p = n0 - (int64_t)k0 - 255 + (int64_t)i;
      xs[i] = (p >= -(int64_t)H && p < (int64_t)T) ? x[p] : 0.0f;
    }
    __syncthreads();
    float part = 0.0f;
#pragma unroll 8
    for (uint32_t kk = 0; kk < 256; kk += 4) {
      const float4 h4 = *reinterpret_cast<const float4*>(&hs[kk]);
      part = __fmaf_rn(h4.x, xs[tid + 255 - kk], part);
      part = __fmaf_rn(h4.y, xs[tid + 254 - kk], part);
      part = __fmaf_rn(h4.z, xs[tid + 253 - kk], part);
      part = __fmaf_rn(h4.w, xs[tid + 252 - kk], part);
    }
    total += (double)part;
  }
  const int64_t n = n0 + tid;
  if (n < (int64_t)T) trackbuf[((size_t)e * T + n) * 2 + c] = (float)total;
}

__global__ void fir_save_kernel(const DFx* __restrict__ fx, uint32_t n_fx, uint32_t C, uint64_t H, uint64_t T,
                                const float* __restrict__ xin, float* __restrict__ hist) {
  const uint32_t ec = blockIdx.y;
  const uint32_t e = ec / C, c = ec % C;
  if (!fx[e].reverb_on) return;
  float* h = hist + ((size_t)fx[e].track * 2 + c) * H;
  const float* x = xin + (size_t)ec * (H + T) + T;  // the last H entries
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < H; i += (uint64_t)gridDim.x * blockDim.x) h[i] = x[i];
}

// point the cells of effect tracks at their processed buffer: one whole-block unity call per callback
__global__ void patch_fx_cells_kernel(const DFx* __restrict__ fx, uint32_t n_fx, uint32_t N, uint32_t S, uint32_t K,
                                      uint32_t B, uint32_t first_fx_span, DCell* __restrict__ cells) {
  const uint64_t id = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= (uint64_t)n_fx * K) return;
  const uint32_t e = (uint32_t)(id / K), k = (uint32_t)(id % K);
  DCell* c = cells + ((size_t)k * N + fx[e].track) * S;
  DCell v;
  v.pos = (double)k * (double)B;
  v.span = first_fx_span + e;
  v.n_act = B;
  c[0] = v;
  v.pos = 0.0;
  v.span = kSilent;
  v.n_act = 0;
  for (uint32_t s = 1; s < S; s++) c[s] = v;
}

// ---------------------------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------------------------
__global__ void clamp_kernel(float* __restrict__ x, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    if (v > 1.0f)
      x[i] = 1.0f;
    else if (v < -1.0f)
      x[i] = -1.0f;
  }
}

// ---- sharded render: exchange step over peer memory ------------------------------------------
// Cross-GPU barrier in two halves. signal: lane j publishes `epoch` into rank j's arrival word for this rank (a peer
// store, after a system-scope fence so this rank's earlier peer stores — the mix kernel's tiles, the reduced slices —
// are visible first). wait: lane j spins until rank j's word here reaches `epoch`. Epochs only grow and a rank cannot
// run two barriers ahead of a peer, so one word per pair suffices. A peer that never arrives (crashed process) must
// not hang the GPU: after `timeout_ns` the wait gives up and raises *status (page-locked host word the API checks
// after the next synchronise).
__global__ void shard_signal_kernel(ShardPeers peers, uint32_t rank, uint32_t world, uint32_t epoch) {
  __threadfence_system();
  const uint32_t j = threadIdx.x;
  if (j < world) *(volatile uint32_t*)(peers.flags[j] + rank) = epoch;
}

__global__ void shard_wait_kernel(ShardPeers peers, uint32_t rank, uint32_t world, uint32_t epoch,
                                  unsigned long long timeout_ns, volatile uint32_t* status) {
  const uint32_t j = threadIdx.x;
  if (j < world) {
    volatile uint32_t* mine = peers.flags[rank] + j;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while ((int32_t)(*mine - epoch) < 0) {
      __nanosleep(100);
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > timeout_ns) {
        *status = 1u;
        break;
      }
    }
  }
  __threadfence_system();
}

// Owner rank's reduce of its slice of callbacks: out = clamp(sum over source ranks 0..W-1, in rank order) — the bus sum
// of AudioBuffer::mix continued across shards, then the clamp that must follow it (engine.cpp:1600-1617, 1627-1636).
// xchg [W][C][plane] is this rank's exchange buffer (filled by every rank's mix kernel, read past L1: the lines were
// written by peers); the result goes to dst[d] + c * chan_stride + dst_off (rank 0's master bus, a peer store).
template <int V>
__global__ void shard_reduce_kernel(const float* __restrict__ xchg, uint32_t W, uint32_t C, uint64_t plane, uint64_t valid,
                                    ShardPeers peers, uint64_t chan_stride, uint64_t dst_off) {
  const uint64_t n = valid / V;
  for (uint32_t c = 0; c < C; c++) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
      float acc[V];
#pragma unroll
      for (int q = 0; q < V; q++) acc[q] = 0.0f;
      for (uint32_t s = 0; s < W; s++) {
        const float* src = xchg + ((size_t)s * C + c) * plane;
        float v[V];
        if constexpr (V == 4) {
          const float4 t = __ldcg(reinterpret_cast<const float4*>(src) + i);
          v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
        } else {
          v[0] = __ldcg(src + i);
        }
#pragma unroll
        for (int q = 0; q < V; q++) acc[q] = s == 0 ? v[q] : __fadd_rn(acc[q], v[q]);
      }
#pragma unroll
      for (int q = 0; q < V; q++) acc[q] = acc[q] > 1.0f ? 1.0f : (acc[q] < -1.0f ? -1.0f : acc[q]);  // NaN passes
      for (uint32_t d = 0; d < peers.n_dst; d++) {
        float* out = peers.dst[d] + (size_t)c * chan_stride + dst_off;
        if constexpr (V == 4)
          reinterpret_cast<float4*>(out)[i] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        else
          out[i] = acc[0];
      }
      if (peers.host_dst[c]) {  // posted stores over this rank's own PCIe link
        float* out = peers.host_dst[c] + dst_off;
        if constexpr (V == 4)
          reinterpret_cast<float4*>(out)[i] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        else
          out[i] = acc[0];
      }
    }
  }
}

// Realtime callback (one-callback render): the submitted table (spans | gains | cells) is pulled from page-locked host
// memory by this kernel and the mix's zero region is cleared by it too, so the callback is kernels only — no hop between
// the copy engine and the SMs (each costs several microseconds of the ~60 a callback takes).
__global__ void ingest_kernel(const uint4* __restrict__ host_table, uint4* __restrict__ table, size_t n_table,
                              uint4* __restrict__ zero, size_t n_zero) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_table; i += stride) table[i] = host_table[i];
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_zero; i += stride) zero[i] = make_uint4(0u, 0u, 0u, 0u);
}

// level_kernel for renders of few callbacks: one thread per (track, channel), result stored straight into page-locked
// host memory (no atomics, no copy afterwards)
__global__ void level_direct_kernel(const float* __restrict__ peaks, uint32_t K, uint32_t NC, float* __restrict__ levels_host) {
  const uint32_t tc = blockIdx.x * blockDim.x + threadIdx.x;
  if (tc >= NC) return;
  float m = 0.0f;
  for (uint32_t k = 0; k < K; k++) m = fmaxf(m, __ldcg(peaks + (size_t)k * NC + tc));
  levels_host[tc] = m;
}

// VUMeter::level semantics over a whole render: max over callbacks of the block peaks (vu_meter.h:25-29).
// peaks [K][NC] (NC = n_tracks*2, all >= 0), levels [NC] pre-zeroed; each thread folds `chunk` callbacks.
__global__ void level_kernel(const float* __restrict__ peaks, uint32_t K, uint32_t NC, uint32_t chunk,
                             float* __restrict__ levels) {
  const uint64_t id = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t tc = (uint32_t)(id % NC);
  const uint32_t k0 = (uint32_t)(id / NC) * chunk;
  if (k0 >= K) return;
  const uint32_t k1 = (k0 + chunk < K) ? k0 + chunk : K;
  float m = 0.0f;
  for (uint32_t k = k0; k < k1; k++) m = fmaxf(m, __ldg(peaks + (size_t)k * NC + tc));
  if (m > 0.0f) atomicMax(reinterpret_cast<unsigned int*>(levels + tc), __float_as_uint(m));
}

// core/audio_format_conv.cpp:5-106. One thread per (frame, channel).
__global__ void interleave_kernel(const float* __restrict__ bus, uint64_t frames, uint32_t channels, int fmt,
                                  void* __restrict__ dst) {
  const uint64_t n = frames * channels;
  for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = o / channels;
    const uint32_t c = (uint32_t)(o % channels);
    const float v = bus[(uint64_t)c * frames + i];
    switch (fmt) {
      case 3:  // I16 (:5-20): positive * 32767, else * 32768, truncating cast
        ((int16_t*)dst)[o] = (int16_t)__float2int_rz(v > 0.0f ? __fmul_rn(v, 32767.0f) : __fmul_rn(v, 32768.0f));
        break;
      case 5: {  // I24 packed (:22-43). The reference writes every channel at dst[3*i .. 3*i+2] (no channel
                 // stride), so the last channel wins; reproduced as written: only the last channel stores.
        if (c == channels - 1) {
          const int32_t q = __float2int_rz(v > 0.0f ? __fmul_rn(v, 8388607.0f) : __fmul_rn(v, 8388608.0f));
          uint8_t* p = (uint8_t*)dst + i * 3;
          p[0] = (uint8_t)q;
          p[1] = (uint8_t)(q >> 8);
          p[2] = (uint8_t)(q >> 16);
        }
        break;
      }
      case 6: {  // I24_X8 (:45-59)
        const int32_t q = __float2int_rz(v > 0.0f ? __fmul_rn(v, 8388607.0f) : __fmul_rn(v, 8388608.0f));
        ((int32_t*)dst)[o] = q & 0xFFFFFF;
        break;
      }
      case 7:  // I32 (:61-74), in double
        ((int32_t*)dst)[o] =
            __double2int_rz(v > 0.0f ? __dmul_rn((double)v, 2147483647.0) : __dmul_rn((double)v, 2147483648.0));
        break;
      default: ((float*)dst)[o] = v; break;  // F32 (:76-88)
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// waveform peak mip-maps (gfx/waveform_visual.cpp:9-173): per chunk of `chunk` frames the converted min and max
// with their first occurrences, stored in order of occurrence. A second HBM-bound scan over the resident samples.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int mip_convert(uint32_t fmt, const void* base, uint64_t elem, bool high) {
  const int tmin = high ? -32768 : -128, tmax = high ? 32767 : 127;
  int q;
  if (fmt == F_F32) {  // :143-151
    const float v = ((const float*)base)[elem];
    q = __float2int_rz(__fmul_rn(v, v >= 0.0f ? (float)tmax : (float)(-tmin)));
  } else if (fmt == F_I16) {  // :66-76
    const int16_t v = ((const int16_t*)base)[elem];
    const float dmin = high ? (-32768.0f / -32768.0f) : (-128.0f / -32768.0f);
    const float dmax = high ? (32767.0f / 32767.0f) : (127.0f / 32767.0f);
    q = __float2int_rz(__fmul_rn((float)v, v >= 0 ? dmax : dmin));
  } else {  // I32 (:104-114), in double
    const int32_t v = ((const int32_t*)base)[elem];
    const double dmin = high ? (-32768.0 / -2147483648.0) : (-128.0 / -2147483648.0);
    const double dmax = high ? (32767.0 / 2147483647.0) : (127.0 / 2147483647.0);
    q = __double2int_rz(__dmul_rn((double)v, v >= 0 ? dmax : dmin));
  }
  return high ? (int)(int16_t)q : (int)(int8_t)q;  // (T)conv
}

// Level 0 (chunks of 2 frames): for a full chunk (v0, v1) the first-occurrence rule always yields (v0, v1) itself —
// v1 < v0 makes v0 the earlier max, v1 > v0 makes v0 the earlier min — and a 1-frame tail chunk yields (v0, v0).
// So level 0 is one streaming pass: read the sample once, write the converted values. One thread per frame pair,
// both channels (one 128-bit load for stereo f32).
__global__ void mip_level0_kernel(const void* __restrict__ base, uint32_t fmt, uint32_t nch, uint64_t count, uint64_t mdc,
                                  int high, void* __restrict__ out) {
  const uint64_t pairs = mdc / 2;
  for (uint64_t pi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; pi < pairs; pi += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t f0 = 2 * pi;
    const bool two_frames = f0 + 1 < count;
    for (uint32_t c = 0; c < nch; c++) {
      const int v0 = mip_convert(fmt, base, f0 * nch + c, high != 0);
      const int v1 = two_frames ? mip_convert(fmt, base, (f0 + 1) * nch + c, high != 0) : v0;
      const uint64_t o = mdc * c + f0;
      if (high) {
        reinterpret_cast<int16_t*>(out)[o] = (int16_t)v0;
        reinterpret_cast<int16_t*>(out)[o + 1] = (int16_t)v1;
      } else {
        reinterpret_cast<int8_t*>(out)[o] = (int8_t)v0;
        reinterpret_cast<int8_t*>(out)[o + 1] = (int8_t)v1;
      }
    }
  }
}

// Level l+1 from level l: a chunk of level l+1 is four consecutive chunks of level l. Each child pair (first, second)
// carries its min, its max and which came first; the parent's min / max are the first child (in order) attaining
// them, and their relative order follows from child order or, inside one child, from that child's pair order —
// exactly the first-occurrence indices summarize_for_mipmaps_impl tracks over the raw samples
// (gfx/waveform_visual.cpp:33-52), because the conversion to int8/int16 happens before the comparisons there too.
template <typename T>
__global__ void mip_merge_kernel(const T* __restrict__ child, uint64_t child_mdc, T* __restrict__ parent,
                                 uint64_t parent_mdc, uint32_t nch, int tmin, int tmax) {
  const uint64_t ppairs = parent_mdc / 2, cpairs = child_mdc / 2;
  const uint64_t total = ppairs * nch;
  for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < total; u += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t c = (uint32_t)(u / ppairs);
    const uint64_t P = u % ppairs;
    int gmin = tmax, gmax = tmin;  // numeric_limits<T>::max() / ::min(), strict compares below
    uint32_t min_pos = 0, max_pos = 0;
#pragma unroll
    for (uint32_t k = 0; k < 4; k++) {
      const uint64_t cp = 4 * P + k;
      if (cp < cpairs) {
        const int f = child[child_mdc * c + 2 * cp], s2 = child[child_mdc * c + 2 * cp + 1];
        const int cmin = f < s2 ? f : s2, cmax = f < s2 ? s2 : f;
        const bool min_first = (f == cmin);
        const uint32_t pmin = 2 * k + ((min_first || cmin == cmax) ? 0u : 1u);
        const uint32_t pmax = 2 * k + ((min_first && cmin != cmax) ? 1u : 0u);
        if (cmin < gmin) {
          gmin = cmin;
          min_pos = pmin;
        }
        if (cmax > gmax) {
          gmax = cmax;
          max_pos = pmax;
        }
      }
    }
    const bool max_first = max_pos < min_pos;
    parent[parent_mdc * c + 2 * P] = (T)(max_first ? gmax : gmin);
    parent[parent_mdc * c + 2 * P + 1] = (T)(max_first ? gmin : gmax);
  }
}

cudaError_t launch_mip_level0(const void* base, uint32_t fmt, uint32_t nch, uint64_t count, uint64_t mdc, int high, void* out,
                              int n_sm, cudaStream_t stream) {
  const uint64_t pairs = mdc / 2;
  if (pairs == 0) return cudaSuccess;
  uint64_t blocks = (pairs + 255) / 256;
  if (blocks > (uint64_t)n_sm * 32) blocks = (uint64_t)n_sm * 32;
  mip_level0_kernel<<<(unsigned)blocks, 256, 0, stream>>>(base, fmt, nch, count, mdc, high, out);
  return cudaGetLastError();
}

cudaError_t launch_mip_merge(const void* child, uint64_t child_mdc, void* parent, uint64_t parent_mdc, uint32_t nch, int high,
                             int n_sm, cudaStream_t stream) {
  const uint64_t total = (parent_mdc / 2) * nch;
  if (total == 0) return cudaSuccess;
  uint64_t blocks = (total + 255) / 256;
  if (blocks > (uint64_t)n_sm * 32) blocks = (uint64_t)n_sm * 32;
  if (high)
    mip_merge_kernel<int16_t><<<(unsigned)blocks, 256, 0, stream>>>((const int16_t*)child, child_mdc, (int16_t*)parent, parent_mdc,
                                                                   nch, -32768, 32767);
  else
    mip_merge_kernel<int8_t><<<(unsigned)blocks, 256, 0, stream>>>((const int8_t*)child, child_mdc, (int8_t*)parent, parent_mdc, nch,
                                                                  -128, 127);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// host-callable launchers (used by wbx_api.cu)
// ---------------------------------------------------------------------------------------------------------
struct MixVariant {
  int fpl, stages, warps;
};

template <int FPL, int STAGES, int WARPS, bool EXT>
static cudaError_t launch_mix_e(const MixParams& p, int n_sm, cudaStream_t stream, int* ctas_out) {
  using L = MixLayout<FPL, STAGES>;
  auto kfn = mix_kernel<FPL, STAGES, WARPS, EXT>;
  const int smem = L::WARP_BYTES * WARPS;
  // attribute + occupancy are properties of the instantiation (per device of the same kind): queried once, not on the
  // realtime callback's path
  static int per_sm_cached[64] = {0};
  int dev = 0;
  cudaError_t err = cudaGetDevice(&dev);
  if (err != cudaSuccess) return err;
  int per_sm = (dev >= 0 && dev < 64) ? per_sm_cached[dev] : 0;
  if (per_sm == 0) {
    err = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess) return err;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, WARPS * 32, smem);
    if (err != cudaSuccess) return err;
    if (per_sm < 1) per_sm = 1;
    if (dev >= 0 && dev < 64) per_sm_cached[dev] = per_sm;
  }
  long ctas = (long)n_sm * per_sm;
  // a render with few work items (the realtime callback) is spread over the SMs with fewer warps per CTA: the bulk-copy
  // issue rate is a per-SM resource, so 8 warps on each of 16 SMs stage their windows 8x slower than 1 warp on each of 128
  int wpc = WARPS;
  if ((long)p.n_items < (long)n_sm * WARPS) {
    wpc = (int)(((long)p.n_items + n_sm - 1) / n_sm);
    if (wpc < 1) wpc = 1;
    if (wpc > WARPS) wpc = WARPS;
  }
  const long need = ((long)p.n_items + wpc - 1) / wpc;
  if (ctas > need) ctas = need;
  if (ctas < 1) ctas = 1;
  if (ctas_out) *ctas_out = (int)ctas;
  kfn<<<(unsigned)ctas, wpc * 32, (size_t)L::WARP_BYTES * wpc, stream>>>(p);
  return cudaGetLastError();
}

// 512-frame tiles: 2 stages x 8 warps per CTA (16 warps/SM, the register-file limit at 128 regs) measured faster
// than 3 stages x 7 warps (14 warps/SM) on both cfg 2 (6.84 vs 6.42 TB/s) and cfg 3 (4.50 vs 4.42 TB/s);
// WBX_VARIANT=a selects the latter for experiments.
template <int FPL, int STAGES, int WARPS>
static cudaError_t launch_mix_t(const MixParams& p, int n_sm, cudaStream_t stream, int* ctas_out) {
  return p.ext ? launch_mix_e<FPL, STAGES, WARPS, true>(p, n_sm, stream, ctas_out)
               : launch_mix_e<FPL, STAGES, WARPS, false>(p, n_sm, stream, ctas_out);
}

static int variant_b() {
  const char* v = getenv("WBX_VARIANT");
  return !(v && v[0] == 'a');
}

cudaError_t launch_mix(const MixParams& p, int fpl, int n_sm, cudaStream_t stream, int* ctas_out) {
  if (fpl == 16 && variant_b()) return launch_mix_t<16, 2, 8>(p, n_sm, stream, ctas_out);
  switch (fpl) {
    case 16: return launch_mix_t<16, 3, 7>(p, n_sm, stream, ctas_out);
    case 8: return launch_mix_t<8, 3, 8>(p, n_sm, stream, ctas_out);
    default: return launch_mix_t<4, 4, 8>(p, n_sm, stream, ctas_out);
  }
}

// resident warps per SM of each variant (shared-memory bound), for the host's work-shape heuristics
int mix_warps_per_sm(int fpl) {
  auto per_sm = [](int warp_bytes, int warps) {
    int ctas = (227 * 1024) / (warp_bytes * warps + 1024);
    return (ctas < 1 ? 1 : ctas) * warps;
  };
  if (fpl == 16 && variant_b()) return per_sm(MixLayout<16, 2>::WARP_BYTES, 8);
  switch (fpl) {
    case 16: return per_sm(MixLayout<16, 3>::WARP_BYTES, 7);
    case 8: return per_sm(MixLayout<8, 3>::WARP_BYTES, 8);
    default: return per_sm(MixLayout<4, 4>::WARP_BYTES, 8);
  }
}

cudaError_t launch_interleave_sample(const void* planar, size_t plane_bytes, uint64_t frames, uint32_t nch,
                                     uint32_t esize, void* dst, int n_sm, cudaStream_t stream) {
  const uint64_t n = frames * nch;
  if (n == 0) return cudaSuccess;
  uint64_t blocks = (n + 255) / 256;
  if (blocks > (uint64_t)n_sm * 16) blocks = (uint64_t)n_sm * 16;
  interleave_sample_kernel<<<(unsigned)blocks, 256, 0, stream>>>((const uint8_t*)planar, plane_bytes, frames, nch, esize,
                                                                 (uint8_t*)dst);
  return cudaGetLastError();
}

cudaError_t launch_expand(const DSpan* spans, uint32_t n_spans, DCell* cells, uint32_t n_tracks, uint32_t slots,
                          uint32_t n_blocks, cudaStream_t stream) {
  if (n_spans == 0) return cudaSuccess;
  // long renders: closed-form runs lane-per-span (coalesced cell stores), everything else warp-per-span
  const uint32_t split = n_blocks >= 64 && n_spans >= 32 ? 1u : 0u;
  if (split) {
    const dim3 grid((n_spans + 127) / 128, (n_blocks + 63) / 64);
    expand_closed_kernel<<<grid, 128, 0, stream>>>(spans, n_spans, cells, n_tracks, slots, n_blocks);
  }
  expand_schedule<<<(n_spans + 3) / 4, 128, 0, stream>>>(spans, n_spans, cells, n_tracks, slots, split);  // warp per span
  return cudaGetLastError();
}

cudaError_t launch_fir_tc(const DFx* fx, uint32_t n_fx, uint32_t C, uint64_t H, uint64_t T, uint32_t L, void* tiles,
                          const float* xin, void* planes, float* trackbuf, cudaStream_t stream);  // wbx_fir_tc.cu

cudaError_t launch_effects(const DSpan* spans, DCell* cells, DFx* fx, uint32_t n_fx, uint32_t N, uint32_t S, uint32_t K,
                           uint32_t B, uint32_t C, uint32_t first_fx_span, float* trackbuf, const float* ir, uint32_t L,
                           float* fir_hist, float* fir_in, void* tc_tiles, void* tc_planes, const float* poly,
                           cudaStream_t stream) {
  if (n_fx == 0) return cudaSuccess;
  const uint64_t warps = (uint64_t)n_fx * K;
  render_tracks_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, stream>>>(spans, cells, fx, n_fx, N, S, K, B, C, poly, trackbuf);
  {
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t err = launch_effects_chain(fx, n_fx, C, (uint64_t)K * B, trackbuf, n_sm, stream);
    if (err != cudaSuccess) return err;
  }
  if (L && ir && fir_hist && fir_in) {  // convolution reverb as the chain's last stage
    const uint64_t T = (uint64_t)K * B, H = L - 1;
    const dim3 gcopy((unsigned)(((H + T) + 255) / 256 < 4096 ? ((H + T) + 255) / 256 : 4096), n_fx * C);
    fir_gather_kernel<<<gcopy, 256, 0, stream>>>(fx, n_fx, C, H, T, fir_hist, trackbuf, fir_in);
    if (tc_tiles && tc_planes) {  // tensor-core path (wbx_fir_tc.cu)
      cudaError_t err = launch_fir_tc(fx, n_fx, C, H, T, L, tc_tiles, fir_in, tc_planes, trackbuf, stream);
      if (err != cudaSuccess) return err;
    } else {
      fir_kernel<<<dim3((unsigned)((T + 255) / 256), n_fx * C), 256, 0, stream>>>(fx, C, H, T, ir, L, fir_in, trackbuf);
    }
    if (H) fir_save_kernel<<<dim3((unsigned)((H + 255) / 256 < 1024 ? (H + 255) / 256 : 1024), n_fx * C), 256, 0, stream>>>(
        fx, n_fx, C, H, T, fir_in, fir_hist);
  }
  patch_fx_cells_kernel<<<(unsigned)((warps + 127) / 128), 128, 0, stream>>>(fx, n_fx, N, S, K, B, first_fx_span, cells);
  return cudaGetLastError();
}

cudaError_t launch_levels(const float* peaks, uint32_t K, uint32_t NC, float* levels, cudaStream_t stream) {
  if (K == 0 || NC == 0) return cudaSuccess;
  const uint32_t chunk = 32;
  const uint64_t threads = (uint64_t)((K + chunk - 1) / chunk) * NC;
  level_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(peaks, K, NC, chunk, levels);
  return cudaGetLastError();
}

cudaError_t launch_shard_signal(const ShardPeers& peers, uint32_t rank, uint32_t world, uint32_t epoch, cudaStream_t stream) {
  shard_signal_kernel<<<1, 32, 0, stream>>>(peers, rank, world, epoch);
  return cudaGetLastError();
}

cudaError_t launch_shard_wait(const ShardPeers& peers, uint32_t rank, uint32_t world, uint32_t epoch,
                              unsigned long long timeout_ns, uint32_t* status, cudaStream_t stream) {
  shard_wait_kernel<<<1, 32, 0, stream>>>(peers, rank, world, epoch, timeout_ns, status);
  return cudaGetLastError();
}

cudaError_t launch_shard_reduce(const float* xchg, uint32_t W, uint32_t C, uint64_t plane, uint64_t valid,
                                const ShardPeers& peers, uint64_t chan_stride, uint64_t dst_off, int n_sm,
                                cudaStream_t stream) {
  if (valid == 0) return cudaSuccess;
  bool v4 = (plane % 4 == 0) && (valid % 4 == 0) && (chan_stride % 4 == 0) && (dst_off % 4 == 0);
  for (int c = 0; c < 2; c++)
    if (peers.host_dst[c] && ((uintptr_t)peers.host_dst[c] & 15u)) v4 = false;
  const uint64_t n = v4 ? valid / 4 : valid;
  uint64_t blocks = (n + 255) / 256;
  if (blocks > (uint64_t)n_sm * 8) blocks = (uint64_t)n_sm * 8;
  if (v4)
    shard_reduce_kernel<4><<<(unsigned)blocks, 256, 0, stream>>>(xchg, W, C, plane, valid, peers, chan_stride, dst_off);
  else
    shard_reduce_kernel<1><<<(unsigned)blocks, 256, 0, stream>>>(xchg, W, C, plane, valid, peers, chan_stride, dst_off);
  return cudaGetLastError();
}

cudaError_t launch_ingest(const void* host_table, void* table, size_t table_bytes, void* zero, size_t zero_bytes,
                          cudaStream_t stream) {
  const size_t n = (table_bytes / 16 > zero_bytes / 16 ? table_bytes : zero_bytes) / 16;
  unsigned blocks = (unsigned)((n + 255) / 256);
  if (blocks > 128) blocks = 128;
  if (blocks < 1) blocks = 1;
  ingest_kernel<<<blocks, 256, 0, stream>>>((const uint4*)host_table, (uint4*)table, table_bytes / 16, (uint4*)zero,
                                            zero_bytes / 16);
  return cudaGetLastError();
}

cudaError_t launch_levels_direct(const float* peaks, uint32_t K, uint32_t NC, float* levels_host, cudaStream_t stream) {
  if (NC == 0) return cudaSuccess;
  level_direct_kernel<<<(NC + 127) / 128, 128, 0, stream>>>(peaks, K, NC, levels_host);
  return cudaGetLastError();
}

cudaError_t launch_clamp(float* x, uint64_t n, int n_sm, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  uint64_t blocks = (n + 255) / 256;
  if (blocks > (uint64_t)n_sm * 8) blocks = (uint64_t)n_sm * 8;
  clamp_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, n);
  return cudaGetLastError();
}

cudaError_t launch_interleave(const float* bus, uint64_t frames, uint32_t channels, int fmt, void* dst, int n_sm,
                              cudaStream_t stream) {
  const uint64_t n = frames * channels;
  if (n == 0) return cudaSuccess;
  uint64_t blocks = (n + 255) / 256;
  if (blocks > (uint64_t)n_sm * 8) blocks = (uint64_t)n_sm * 8;
  interleave_kernel<<<(unsigned)blocks, 256, 0, stream>>>(bus, frames, channels, fmt, dst);
  return cudaGetLastError();
}

}  // namespace wbx
