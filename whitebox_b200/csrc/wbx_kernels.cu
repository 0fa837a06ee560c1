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

#include <cstddef>
#include <type_traits>

#include <cstdint>
#include <cstdio>
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
// global -> shared bulk copy (16-B aligned addresses, size a multiple of 16; completes `bytes` on `bar`):
// one elected lane of a CONVERGED warp announces `bytes` on `bar` and issues the bulk copy. Written as one predicated
// block around elect.sync: ptxas then knows a single lane is active and moves the operands to uniform registers with
// plain R2UR — an `if (lane == 0)` around the same two instructions compiles to a divergent branch plus a
// first-active-lane loop (ELECT / R2UR.BROADCAST / BRA.U.ANY) around UBLKCP, ~20 more instructions per copy.
__device__ __forceinline__ void bulk_g2s_elect(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%3], %2;\n\t"
      "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}" ::"r"(dst),
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

// Lane <-> frame mapping of a tile of T = 32*FPL frames: lane owns the frames lane + 32*m, m < FPL; acc[m] is the
// (L, R) bus accumulator of frame lane + 32*m. Consecutive lanes read consecutive source frames: an interleaved stereo
// f32 frame is one 8-byte shared-memory word, so a warp-wide load of one frame per lane is conflict-free at unity speed
// and stays within two 128-byte rows for resampled clips (round 1 gave a lane the frame PAIR (2q, 2q+1) for 128-bit
// loads on the unity path; the resampled paths then read every other word and paid two wavefronts per load).
// The loops below still run (i, e) with m = 2*i + e.

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

// The mix kernel's copy of the same evaluation: the coefficient table staged in shared memory (rows padded to
// POLY_ROW floats: 32 lanes gather 32 different rows per tap group, which through L1 costs one tag lookup per lane and
// bounded the first version at 5.6e10 track-frames/s), (L, R) accumulated as one packed FFMA2 per tap (per-component rn:
// the same bits as the two scalar fused multiply-adds of the specification).
constexpr int POLY_ROW = 20;                       // floats per phase row in shared memory (16 taps + 4 pad: 80 B, 16-B aligned)
constexpr int POLY_SMEM_BYTES = 128 * POLY_ROW * 4;
__device__ __forceinline__ float2 poly_frame_s(const float* tab_s, uint32_t rb_s, uint32_t ix, double fd) {
  const int ph = __double2int_rz(__dmul_rn(fd, 128.0));
  const float4* h4 = reinterpret_cast<const float4*>(tab_s + ph * POLY_ROW);
  float2 acc = make_float2(0.0f, 0.0f);
  const uint32_t a0 = rb_s + (ix - 7u) * 8u;  // shared address of source frame ix - 7
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const float4 h = h4[q];
    float2 s0, s1, s2, s3;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(s0.x), "=f"(s0.y) : "r"(a0 + (4 * q + 0) * 8u));
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(s1.x), "=f"(s1.y) : "r"(a0 + (4 * q + 1) * 8u));
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(s2.x), "=f"(s2.y) : "r"(a0 + (4 * q + 2) * 8u));
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(s3.x), "=f"(s3.y) : "r"(a0 + (4 * q + 3) * 8u));
    acc = __ffma2_rn(make_float2(h.x, h.x), s0, acc);
    acc = __ffma2_rn(make_float2(h.y, h.y), s1, acc);
    acc = __ffma2_rn(make_float2(h.z, h.z), s2, acc);
    acc = __ffma2_rn(make_float2(h.w, h.w), s3, acc);
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
      const int fr = lane + 32 * (2 * i + e);
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

// Fast path: stereo f32, unity speed, the whole tile. One 64-bit shared load = one (L, R) frame; gain, pan and the bus
// add run as packed f32x2 (each component separately rounded, rn).
template <int FPL>
__device__ __forceinline__ void consume_fast(const uint8_t* row, float gain, float tgL, float tgR,
                                             float2 (&acc)[FPL], float& pkL, float& pkR, int lane) {
  const float2* r2 = reinterpret_cast<const float2*>(row);
  const float2 g2 = make_float2(gain, gain);
  const float2 t2 = make_float2(tgL, tgR);
  float2 v[FPL];
#pragma unroll
  for (int m = 0; m < FPL; m++) v[m] = r2[lane + 32 * m];
#pragma unroll
  for (int m = 0; m < FPL; m += 2) {
    const float2 a = __fmul2_rn(__fmul2_rn(v[m], g2), t2);      // frame lane + 32 m:       (L, R)
    const float2 b = __fmul2_rn(__fmul2_rn(v[m + 1], g2), t2);  // frame lane + 32 (m + 1): (L, R)
    acc[m] = __fadd2_rn(acc[m], a);
    acc[m + 1] = __fadd2_rn(acc[m + 1], b);
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
      const int fr = lane + 32 * (2 * i + e);
      if (FULL || (fr >= lo && fr < hi)) accumulate2(r2[fr + shift], g2, t2, acc[i * 2 + e], pkL, pkR);
    }
  }
}


// Stereo f32, 2-tap linear resample (sample_linear<float, F32>, dsp/sampler.cpp:34-59) from the staged window.
// The position split avoids the slow f64<->int conversions: for 0 <= x < 2^31, t = x + 2^52 rounded TOWARDS
// -INF is exactly floor(x) + 2^52 (the ulp there is 1), so its low mantissa word is (int64_t)x and
// x - (t - 2^52) is x - (double)ix — both subtractions exact — as in sampler.cpp:51-52.
// 64-bit shared-memory load by 32-bit shared address: keeps the per-frame address to one integer multiply-add
__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}

// `one` is 1.0f read from the kernel parameters — a value ptxas cannot see: fma(prod, one, a) is add_rn(prod, a)
// exactly (prod * 1 is exact, one rounding), but unlike mul.rn.f32x2 + add.rn.f32x2 it cannot be contracted into a
// fused multiply-add of the product's own factors (ptxas does that even under --fmad false when the product has a
// single use; tests/test_host_cpu.py lints the SASS). It makes the lerp three packed instructions instead of five.
#ifndef WBX_LIN_GROUP
#define WBX_LIN_GROUP 8
#endif
// Frames are taken WBX_LIN_GROUP at a time in three phases — the f64 position chains of the group, its shared-memory
// loads, the packed lerp / gain / bus add — so that several dependent chains (7 f64-pipe operations deep each) are in
// flight per warp; every frame has its own accumulator, so the grouping does not touch the arithmetic.
template <int FPL, bool FULL>
__device__ __forceinline__ void consume_lin_t(const Desc& d, const uint8_t* row, float2 (&acc)[FPL], float& pkL,
                                              float& pkR, int lane, float one) {
  constexpr int G = (FULL && FPL % WBX_LIN_GROUP == 0) ? WBX_LIN_GROUP : 2;
  const int lo = d.lo, hi = d.hi;
  const uint32_t rb = smem_u32(row) - (uint32_t)d.base * 8u;  // shared address of source frame 0 (wraps; only sums are used)
  const double pos = d.pos, speed = d.speed;
  const double M = 4503599627370496.0;  // 2^52
  const double jj0 = (double)(d.jrel0 + lane);
  const float2 g2 = make_float2(d.gain, d.gain);
  const float2 t2 = make_float2(d.tg[0], d.tg[1]);
  const float2 neg1 = make_float2(-1.0f, -1.0f);
  const float2 one2 = make_float2(one, one);
#pragma unroll
  for (int m0 = 0; m0 < FPL; m0 += G) {
    uint32_t addr[G];
    float fx[G];
    bool on[G];
#pragma unroll
    for (int u = 0; u < G; u++) {
      const int fr = lane + 32 * (m0 + u);
      on[u] = FULL || (fr >= lo && fr < hi);
      addr[u] = 0u;
      fx[u] = 0.0f;
      if (on[u]) {
        const double jj = __dadd_rn(jj0, (double)(32 * (m0 + u)));  // exact small integers == (double)j
        const double x = __dadd_rn(pos, __dmul_rn(jj, speed));    // sampler.cpp:50
        const double t = __dadd_rd(x, M);                         // floor(x) + 2^52
        const uint32_t ix = (uint32_t)__double2loint(t);          // (int64_t)x, :51
        fx[u] = __double2float_rn(__dsub_rn(x, __dsub_rn(t, M)));  // (float)(x - (double)ix), :52
        addr[u] = rb + ix * 8u;
      }
    }
    float2 a[G], b[G];
#pragma unroll
    for (int u = 0; u < G; u++) {
      a[u] = make_float2(0.0f, 0.0f), b[u] = a[u];
      if (on[u]) a[u] = lds_f2(addr[u]), b[u] = lds_f2(addr[u] + 8u);
    }
    float2 term[G];
#pragma unroll
    for (int u = 0; u < G; u++) {
      term[u] = make_float2(0.0f, 0.0f);
      if (on[u]) {
        const float2 df = __ffma2_rn(a[u], neg1, b[u]);               // b - a (a * -1 is exact: one rounding)
        const float2 pr = __fmul2_rn(make_float2(fx[u], fx[u]), df);  // fx * (b - a)
        const float2 sv = __ffma2_rn(pr, one2, a[u]);                 // a + fx * (b - a), :55 (see `one` above)
        term[u] = __fmul2_rn(__fmul2_rn(sv, g2), t2);
        acc[m0 + u] = __fadd2_rn(acc[m0 + u], term[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < G; u += 2) {
      pkL = fmaxf(fmaxf(pkL, fabsf(term[u].x)), fabsf(term[u + 1].x));
      pkR = fmaxf(fmaxf(pkR, fabsf(term[u].y)), fabsf(term[u + 1].y));
    }
  }
}


// Stereo f32, polyphase windowed-sinc resample (extension) from the staged window: the lin path's position split,
// then 16 taps per frame.
template <int FPL>
__device__ __forceinline__ void consume_poly(const Desc& d, const uint8_t* row, const float* tab_s, float2 (&acc)[FPL], float& pkL,
                                             float& pkR, int lane) {
  const uint32_t rb_s = smem_u32(row) - (uint32_t)d.base * 8u;  // shared address of source frame 0 (wraps; only sums are used)
  const int lo = d.lo, hi = d.hi;
  const double pos = d.pos, speed = d.speed;
  const double M = 4503599627370496.0;  // 2^52
  const double jj0 = (double)(d.jrel0 + lane);
  const float2 g2 = make_float2(d.gain, d.gain);
  const float2 t2 = make_float2(d.tg[0], d.tg[1]);
#pragma unroll
  for (int i = 0; i < FPL / 2; i++) {
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int fr = lane + 32 * (2 * i + e);
      if (fr >= lo && fr < hi) {
        const double jj = __dadd_rn(jj0, (double)(32 * (2 * i + e)));
        const double x = __dadd_rn(pos, __dmul_rn(jj, speed));
        const double t = __dadd_rd(x, M);  // floor(x) + 2^52 (x >= 0)
        const uint32_t ix = (uint32_t)__double2loint(t);
        const double fd = __dsub_rn(x, __dsub_rn(t, M));
        const float2 sv = poly_frame_s(tab_s, rb_s, ix, fd);
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
__device__ __forceinline__ uint32_t resolve_store(const DCell& c, const DSpan& s, const float* __restrict__ gains,
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
          d.kind = (lo == 0 && hi == T) ? (first == a ? K_FAST : K_UNI) : K_UNI_P;
        else if (st32 && last < (int64_t)0x3fffffff)
          d.kind = (lo == 0 && hi == T) ? K_LIN : K_LIN_P;
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
  return d.kind;
}

// A batch of descriptors is LEAN when its BATCH cells all resolved to the same whole-tile stereo-f32 kind (the shape of
// BASELINE cfg 2 / cfg 3: every track plays one clip across the tile). Such a batch runs through a loop that knows the
// kind, knows every cell stages a window and (one slot per track) ends a track: no per-cell dispatch, no `bytes` test in
// front of the bulk copy, no slot / activity bookkeeping. Returns the kind, 0 when the batch is mixed. All lanes call.
__device__ __forceinline__ uint32_t lean_batch_kind(uint32_t my_kind, int lane, int batch) {
  const uint32_t k0 = __shfl_sync(0xffffffffu, my_kind, 0);
  const bool same = __all_sync(0xffffffffu, lane >= batch || my_kind == k0);
  return (same && (k0 == K_FAST || k0 == K_LIN || k0 == K_UNI)) ? k0 : 0u;
}

template <int FPL, int STAGES, int WARPS, bool EXT>
__global__ void __launch_bounds__(WARPS * 32, 2) mix_kernel(const MixParams p) {
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

  // EXT build: the polyphase coefficient table behind the warps' regions (launch_mix_e sizes the allocation)
  const float* poly_s = nullptr;
  if (EXT && p.poly != nullptr) {
    float* tab = reinterpret_cast<float*>(smem + (size_t)(blockDim.x >> 5) * L::WARP_BYTES);
    for (int i = threadIdx.x; i < 128 * 16; i += blockDim.x) tab[(i >> 4) * POLY_ROW + (i & 15)] = __ldg(p.poly + i);
    poly_s = tab;
    __syncthreads();
  }

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

    float2 acc[FPL];
#pragma unroll
    for (int i = 0; i < FPL; i++) acc[i] = make_float2(0.0f, 0.0f);
    float pkL = 0.0f, pkR = 0.0f;
    bool active = false;
    uint32_t cur_track = 0;
    uint32_t slot_ctr = 0;  // position of the current cell within its track's `slots` cells

    // descriptors of the first batch (lanes >= BATCH idle); the barrier orders the previous item's last descriptor reads
    // before these writes (the shuffle above synchronises execution, not memory)
    __syncwarp();
    uint32_t mk0 = K_SILENT;
    if (lane < BATCH) {
      const DCell c0 = load_cell(cells, lane, n_cells);
      const DSpan s0 = load_span(p.spans, c0);
      mk0 = resolve_store<L::STAGE_BYTES, L::T>(c0, s0, p.gains, f0, tile_len, two, k, &ring[lane]);
    }
    __syncwarp();
    uint32_t lean = (S == 1u) ? lean_batch_kind(mk0, lane, BATCH) : 0u;  // kind of the current batch when it is lean
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
        const uint32_t bytes = dd->bytes;  // non-zero exactly for the kinds that stage a window
        if (bytes) {
          const uint32_t st = n_issued % STAGES;
          bulk_g2s_elect(rows_s + st * L::STAGE_BYTES, dd->src, bytes, bars_s + 8 * st);
          n_issued++;
        }
        ip++;
      }
    };
    // VU block peak of a finished track (vu_meter.h:20-30). Peaks are >= 0 and never NaN (fmaxf drops NaNs), so uint
    // order == float order: one warp-wide integer max per channel (REDUX) instead of a five-step shuffle butterfly
    auto flush_peaks = [&](uint32_t track) {
      const uint32_t mL = __reduce_max_sync(0xffffffffu, __float_as_uint(pkL));
      const uint32_t mR = __reduce_max_sync(0xffffffffu, __float_as_uint(pkR));
      if (lane < 2 && (lane == 0 || two)) {
        const uint32_t m = lane ? mR : mL;
        uint32_t* dst = reinterpret_cast<uint32_t*>(p.peaks) + ((k * N + track) * 2u + (uint32_t)lane);  // < 2^32 (launch_mix)
        if (p.n_tiles == 1)
          *dst = m;
        else if (m != 0u)
          atomicMax(dst, m);
      }
      pkL = 0.0f;
      pkR = 0.0f;
    };

    for (uint32_t b = 0; b < nb; b++) {
      const bool more = (b + 1 < nb);
      if (more && lane < BATCH) cN = load_cell(cells, (b + 1) * BATCH + lane, n_cells);
      uint32_t lean_next = 0u;
      // descriptors of the next batch, half a batch ahead
      auto resolve_next = [&]() {
        __syncwarp();  // every lane is done reading the ring half written next (the previous batch's descriptors)
        uint32_t mk = K_SILENT;
        if (lane < BATCH)
          mk = resolve_store<L::STAGE_BYTES, L::T>(cN, sN, p.gains, f0, tile_len, two, k, &ring[((b + 1) & 1) * BATCH + lane]);
        __syncwarp();
        limit += BATCH;
        if (S == 1u) lean_next = lean_batch_kind(mk, lane, BATCH);
      };
      // ---- lean batch: BATCH whole-tile cells of one kind, each staging a window and ending its track ------------------
      // The pipeline is brought to depth STAGES - 1 up front, so the producer is at most one step per cell (a mixed batch
      // before this one may have left it full: the depth test stays); descriptors are walked by byte offset; the 16 block
      // peaks are parked in lanes 0..15 and stored once per batch (tracks of a batch are neighbours: one 128-byte run).
      auto lean_batch = [&](auto kind_c) {
        constexpr uint32_t KIND = decltype(kind_c)::value;
        constexpr uint32_t RING_BYTES = L::RING * (uint32_t)sizeof(Desc);
        uint32_t stage_lim = (b + 1) * BATCH;  // cells below it are known to stage a window
        const uint8_t* ring_b = reinterpret_cast<const uint8_t*>(ring);
        auto stage_one = [&]() {
          const Desc* dd = reinterpret_cast<const Desc*>(ring_b + (ip * (uint32_t)sizeof(Desc)) % RING_BYTES);
          const uint32_t st = n_issued % STAGES;
          bulk_g2s_elect(rows_s + st * L::STAGE_BYTES, dd->src, dd->bytes, bars_s + 8 * st);
          n_issued++;
          ip++;
        };
        while (ip < stage_lim && (n_issued - n_consumed) < (uint32_t)(STAGES - 1)) stage_one();
        uint32_t keepL = 0u, keepR = 0u;  // lane i: block peaks of the batch's cell i
        uint32_t doff = (b & 1u) * BATCH * (uint32_t)sizeof(Desc);
        int cib = 0;  // cell within the batch
#pragma unroll 1
        for (int q = 0; q < 4; q++) {
          if (more) {
            if (q == 1 && lane < BATCH) sN = load_span(p.spans, cN);
            if (q == 2) {
              resolve_next();
              if (lean_next) stage_lim += BATCH;
            }
          }
#pragma unroll 1
          for (int j = 0; j < 4; j++) {
            if (ip < stage_lim && (n_issued - n_consumed) < (uint32_t)STAGES) stage_one();
            const Desc* dp = reinterpret_cast<const Desc*>(ring_b + doff);
            const uint32_t st = n_consumed % STAGES;
            const uint32_t par = (n_consumed / STAGES) & 1u;
            while (!mbar_try_wait(bars_s + 8 * st, par)) {
            }
            const uint8_t* row = wbase + (size_t)st * L::STAGE_BYTES;
            if (KIND == K_FAST)
              consume_fast<FPL>(row, dp->gain, dp->tg[0], dp->tg[1], acc, pkL, pkR, lane);
            else if (KIND == K_LIN)
              consume_lin_t<FPL, true>(*dp, row, acc, pkL, pkR, lane, p.one);
            else
              consume_uni_t<FPL, true>(*dp, row, acc, pkL, pkR, lane);
            __syncwarp();  // every lane is done reading the stage before the elected lane may refill it
            n_consumed++;
            const uint32_t mL = __reduce_max_sync(0xffffffffu, __float_as_uint(pkL));
            const uint32_t mR = __reduce_max_sync(0xffffffffu, __float_as_uint(pkR));
            if (lane == cib) keepL = mL, keepR = mR;
            pkL = 0.0f;
            pkR = 0.0f;
            cib++;
            doff += (uint32_t)sizeof(Desc);
          }
        }
        if (lane < BATCH) {  // VU block peaks of the batch (vu_meter.h:20-30), as flush_peaks stores them
          const uint32_t track = ring[(b & 1u) * BATCH + lane].track;
          uint32_t* dst = reinterpret_cast<uint32_t*>(p.peaks) + (k * N + track) * 2u;  // < 2^32 (launch_mix)
          if (p.n_tiles == 1) {
            if (two)
              *reinterpret_cast<uint2*>(dst) = make_uint2(keepL, keepR);
            else
              dst[0] = keepL;
          } else {
            if (keepL != 0u) atomicMax(dst, keepL);
            if (two && keepR != 0u) atomicMax(dst + 1, keepR);
          }
        }
      };
      if (lean == K_LIN) {
        lean_batch(std::integral_constant<uint32_t, K_LIN>());
        lean = lean_next;
        continue;
      }
      if (lean == K_FAST) {
        lean_batch(std::integral_constant<uint32_t, K_FAST>());
        lean = lean_next;
        continue;
      }
      if (lean == K_UNI) {
        lean_batch(std::integral_constant<uint32_t, K_UNI>());
        lean = lean_next;
        continue;
      }
#pragma unroll 1
      for (int i = 0; i < BATCH; i++) {
        const uint32_t ci = b * BATCH + i;
        if (ci >= n_cells) break;
        if (more && i == BATCH / 4 && lane < BATCH) sN = load_span(p.spans, cN);
        if (more && i == BATCH / 2) resolve_next();
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
          } else if (kind == K_LIN) {
            consume_lin_t<FPL, true>(*dp, row, acc, pkL, pkR, lane, p.one);
          } else if (kind == K_UNI) {
            consume_uni_t<FPL, true>(*dp, row, acc, pkL, pkR, lane);
          } else if (kind == K_LIN_P) {
            consume_lin_t<FPL, false>(*dp, row, acc, pkL, pkR, lane, p.one);
          } else if (kind == K_UNI_P) {
            consume_uni_t<FPL, false>(*dp, row, acc, pkL, pkR, lane);
          } else if (EXT && kind == K_POLY) {
            consume_poly<FPL>(*dp, row, poly_s, acc, pkL, pkR, lane);
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
            flush_peaks(cur_track);
            active = false;
          }
        }
      }
      lean = lean_next;
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
    const bool vec_ok = (p.B & 1u) == 0;  // frame pairs are 8-byte aligned in the planar bus (tree mode's final stores)
    if (p.groups == 1) {
#pragma unroll
      for (int m = 0; m < FPL; m++) {
        const int fr = lane + 32 * m;  // one frame per lane: every store is one contiguous 128-byte run per warp
        if (fr < tile_len) {
#pragma unroll
          for (int c = 0; c < 2; c++) {
            if (c == 1 && !two) break;
            float x0 = c ? acc[m].y : acc[m].x;
            if (p.clamp) x0 = x0 > 1.0f ? 1.0f : (x0 < -1.0f ? -1.0f : x0);  // engine.cpp:1627-1636 (NaN passes)
            outp[c][fr] = x0;
            if (p.mirror[c]) p.mirror[c][out_off + fr] = x0;
          }
        }
      }
    } else {
      // tree mode: publish this group's partial, the last group to arrive adds them in group order
      const uint32_t tile_id = k * p.n_tiles + f;
      float* part = p.ws + ((size_t)tile_id * p.groups + g) * 2 * L::T;  // [channel][T] planar
#pragma unroll
      for (int m = 0; m < FPL; m++) {
        part[lane + 32 * m] = acc[m].x;         // channel 0
        part[L::T + lane + 32 * m] = acc[m].y;  // channel 1
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
                                     uint32_t B, uint32_t C, const float* __restrict__ poly, float* __restrict__ trackbuf, uint64_t tbs) {
  const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (w >= (uint64_t)n_fx * K) return;
  const uint32_t e = (uint32_t)(w / K), k = (uint32_t)(w % K);
  if (fx[e].eq_on || fx[e].comp_on) return;  // rendered inside fx_chain_kernel
  const uint32_t t = fx[e].track;
  float2* out = reinterpret_cast<float2*>(trackbuf) + (size_t)e * tbs + (size_t)k * B;
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

// The memoryless part of the compressor: gain from the envelope, applied with the make-up gain
// (oracle/wb_oracle.c fx_comp_stage, last loop).
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

// ---------------------------------------------------------------------------------------------------------
// fx_chain_kernel — track render + 4-band EQ + compressor, TIME-PARALLEL (extension, BASELINE cfg 4)
//
// Specification: oracle/wb_oracle.c apply_effects (fx_eq_stage / fx_comp_stage) — the chain evaluated in a blocked
// association chosen so that almost nothing is serial in time; this kernel performs exactly those IEEE operations
// (explicit _rn intrinsics; packed f32x2 forms are per-component rn), so CUDA == the C spec bit for bit.
//
// One CTA carries TPC tracks (1, 2 or 4: enough to give every SM one CTA), software-pipelined over chunks of <= 512
// frames of a callback, one __syncthreads per iteration, chunk c living in ring slot X[c % 5]:
//   iteration i:   output warps       loads of the clips of chunk i in flight (Sampler::stream + clip gain; the chunk's cell
//                                     was pulled into L1 two iterations earlier), then
//                                     chunk i-2: the intercepts Q4 of the follower's 4-step look-ahead per block of 4 frames
//                                     (with two EQ warps per track the second one builds them instead),
//                                     chunk i-4: envelope inside each block of 4 frames (1/2/3-step look-ahead from the
//                                     block start), gain computer, make-up, store to the track buffer the mix kernel
//                                     reads; then chunk i -> X[i % 5]
//                  EQ warp            (per track) the four biquads of chunk i-1 in place: per biquad a zero-state pass
//                                     over the lane's 16-frame segment, a Kogge-Stone scan of the 32 segment end states
//                                     with A^(16 * 2^j), the zero-input correction; L and R share coefficients -> packed
//                                     f32x2 math
//                  serial warp        chunk i-3: env <- mm_i(S4[i] * env + Q4[i]) per block — the ONLY recurrence that
//                                     is walked serially (5 independent FMAs + 2 three-input max per 4 frames); lane =
//                                     (track, channel), so the TPC tracks of the CTA share one instruction stream
// ---------------------------------------------------------------------------------------------------------
constexpr int FXC_SEG_STRIDE = 36;                // floats per 16-frame segment of packed (L, R) frames: 32 + 4 pad
constexpr int FXC_XSLOT = 32 * FXC_SEG_STRIDE;    // one chunk of 512 frames
constexpr int FXC_BLOCKS = 128;                   // 4-frame blocks per chunk
constexpr int FXC_QPAIR = 5 * FXC_BLOCKS + 4;     // floats per (parity, pair) of intercepts [5][128]; +4: pairs 16 B apart mod 128
constexpr int FXC_EPAIR = FXC_BLOCKS + 4;         // floats per (parity, pair) of block-end envelopes: [3] = incoming, [4 + g]

struct FxcTrack {
  float X[5][FXC_XSLOT];  // ring: render | EQ | intercepts | (held while the serial warp runs) | output
  float TP[2][2][3][2];  // tail frames of a chunk (n % 4): (pa, pr) per channel and frame, signed domain
  float ET[2][2][4];     // envelope at the tail frames, signed domain
  float P[4][17][4];     // A^m per biquad (row-major 2x2)
  float S[4][5][4];      // A^(16 * 2^j)
  float cf[4][8];        // b0, -a1, -a2, B1, B2 per biquad
  float2 st[4][2];       // biquad states (s1, s2) as (L, R), carried from chunk to chunk
};
template <int TPC>
struct FxcSmem {
  FxcTrack tr[TPC];
  float Q4[2][2 * TPC][FXC_QPAIR];  // [parity][pair = track * 2 + channel][candidate i][block], signed domain
  float E[2][2 * TPC][FXC_EPAIR];   // envelope at the end of every block
};

__device__ __forceinline__ float2 f2(float v) { return make_float2(v, v); }
__device__ __forceinline__ float2 shfl_up2(float2 v, unsigned d) {
  return make_float2(__shfl_up_sync(0xffffffffu, v.x, d), __shfl_up_sync(0xffffffffu, v.y, d));
}
__device__ __forceinline__ float2 fmax2(float2 a, float2 b) { return make_float2(fmaxf(a.x, b.x), fmaxf(a.y, b.y)); }
// smem address of frame f of a chunk slot (packed (L, R) float2)
__device__ __forceinline__ int fxc_addr(int f) { return (f >> 4) * FXC_SEG_STRIDE + ((f & 15) << 1); }

// what the clips of (callback k, track t) put at frame j of the cleared mixing buffer (generic path, from global)
__device__ __noinline__ float2 fxc_render_frame(const DSpan* __restrict__ spans, const DCell* __restrict__ cells, uint32_t k,
                                                uint32_t t, uint32_t N, uint32_t S, uint32_t j, uint32_t C,
                                                const float* __restrict__ poly) {
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
  return v;
}

// intercepts of the follower's look-ahead for one block of 4 frames, both channels packed; signed domain (negated when
// the follower takes the smaller candidate, so that mm is always max). pa / pr = (1 - att)|x|, (1 - rel)|x|.
struct FxcQ {
  float2 q1[2], q2[3], q3[4], q4[5];
};
template <int DEPTH>
__device__ __forceinline__ void fxc_intercepts(const float2 (&x)[4], float a, float r, float a1m_s, float r1m_s, FxcQ& q) {
  float2 pa[4], pr[4];
#pragma unroll
  for (int m = 0; m < 4; m++) {
    const float2 xa = make_float2(fabsf(x[m].x), fabsf(x[m].y));
    pa[m] = __fmul2_rn(f2(a1m_s), xa);
    pr[m] = __fmul2_rn(f2(r1m_s), xa);
  }
  const float2 a2v = f2(a), r2v = f2(r);
  q.q1[0] = pr[0], q.q1[1] = pa[0];
  q.q2[0] = __ffma2_rn(r2v, q.q1[0], pr[1]);
  q.q2[1] = fmax2(__ffma2_rn(r2v, q.q1[1], pr[1]), __ffma2_rn(a2v, q.q1[0], pa[1]));
  q.q2[2] = __ffma2_rn(a2v, q.q1[1], pa[1]);
  q.q3[0] = __ffma2_rn(r2v, q.q2[0], pr[2]);
#pragma unroll
  for (int i = 1; i < 3; i++) q.q3[i] = fmax2(__ffma2_rn(r2v, q.q2[i], pr[2]), __ffma2_rn(a2v, q.q2[i - 1], pa[2]));
  q.q3[3] = __ffma2_rn(a2v, q.q2[2], pa[2]);
  if (DEPTH >= 4) {
    q.q4[0] = __ffma2_rn(r2v, q.q3[0], pr[3]);
#pragma unroll
    for (int i = 1; i < 4; i++) q.q4[i] = fmax2(__ffma2_rn(r2v, q.q3[i], pr[3]), __ffma2_rn(a2v, q.q3[i - 1], pa[3]));
    q.q4[4] = __ffma2_rn(a2v, q.q3[3], pa[3]);
  }
}

// one step of the serial recurrence: env after a block of 4 frames
__device__ __forceinline__ float fxc_step(float ev, const float (&s4)[5], float q0, float q1, float q2, float q3, float q4) {
  const float t0 = __fmaf_rn(s4[0], ev, q0), t1 = __fmaf_rn(s4[1], ev, q1), t2 = __fmaf_rn(s4[2], ev, q2);
  const float t3 = __fmaf_rn(s4[3], ev, q3), t4 = __fmaf_rn(s4[4], ev, q4);
  return fmaxf(fmaxf(fmaxf(t0, t1), t2), fmaxf(t3, t4));
}

template <int TPC, int OW_, int EQW_>
struct FxcShape {
  static constexpr int OW = OW_;    // output warps per track
  static constexpr int EQW = EQW_;  // EQ warps per track: 1 = all four biquads in one warp, 2 = a / b pipeline
  static constexpr int WARPS = 1 + EQW * TPC + TPC * OW;
  // warp slots of the CTA; the 4-track shape gets three unused ones so that its roles can be placed per sub-partition
  // (registers are allocated four warps at a time: 13 warps already cost 16)
  static constexpr int SLOTS = (TPC == 4 && OW_ == 2 && EQW_ == 1) ? 16 : WARPS;
  static constexpr int THREADS = SLOTS * 32;
  static constexpr int OLANES = OW * 32;       // lanes rendering / finishing one track
};

// Which role the warp in each slot of the CTA plays. A warp runs on sub-partition (slot % 4) of its SM, and the roles are
// very different instruction streams — the serial warp and the EQ warps are dependent chains that need an issue slot the
// cycle they become ready, the output warps are long independent streams — so WHERE a role sits decides how long it takes:
// with the slots simply in role order (serial, EQ..., output...) the serial warp of the 4-track CTA shared a sub-partition
// with an EQ warp and two output warps and took 86 % of the iteration (45 cycles per step of a 12-cycle chain), the EQ warp
// next to it 83 % against 63 % for its three siblings (per-role clock64 counters, profiles/r02c_fx_roles.log).
// code: 0 = serial warp, 1 + tl = EQ warp a of track tl, 9 + tl = EQ warp b, 32 + tl * OW + ow = output warp, 255 = unused
struct FxcRoles {
  uint8_t code[32];
};

template <int TPC, int OW_, int EQW_, int MINB>
__global__ void __launch_bounds__((FxcShape<TPC, OW_, EQW_>::THREADS), MINB)
fx_chain_kernel(const DSpan* __restrict__ spans, const DCell* __restrict__ cells, DFx* __restrict__ fx, uint32_t n_fx, uint32_t N,
                uint32_t S, uint32_t K, uint32_t B, uint32_t C, const float* __restrict__ poly, float* __restrict__ trackbuf,
                uint64_t tbs, const FxcRoles roles) {
  using SH = FxcShape<TPC, OW_, EQW_>;
  extern __shared__ __align__(16) unsigned char fxc_raw[];
  FxcSmem<TPC>& sm = *reinterpret_cast<FxcSmem<TPC>*>(fxc_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // roles by slot (FxcRoles): the serial warp (lane = pair), EQ a / EQ b (one of each per track), output warps
  const int rc = roles.code[warp & 31];
  if (rc == 255) return;  // unused slot: the CTA's barriers count the live warps only
  const bool w_serial = rc == 0, w_eq = rc >= 1 && rc < 32;
  const bool w_eqb = SH::EQW == 1 ? w_eq : (w_eq && rc >= 9);  // the warp that finishes the EQ and builds the intercepts
  const bool w_eqa = SH::EQW == 1 ? w_eq : (w_eq && rc < 9);
  const int oidx = rc - 32;
  const int tl = w_serial ? (lane >> 1) : (w_eq ? (rc - 1) & 7 : oidx / SH::OW);  // local track of this thread
  const int ow = w_serial || w_eq ? 0 : oidx % SH::OW;
  const uint32_t e = blockIdx.x * TPC + (uint32_t)(tl < TPC ? tl : 0);
  const bool live_lane = tl < TPC && e < n_fx;
  DFx* __restrict__ f = fx + (live_lane ? e : 0);
  const bool eq_on = live_lane && f->eq_on != 0, comp_on = live_lane && f->comp_on != 0;
  const bool active = eq_on || comp_on;  // reverb-only chains are rendered by render_tracks_kernel
  const uint32_t t = f->track;
  FxcTrack& tr = sm.tr[tl < TPC ? tl : 0];

  // ---- tables and states into shared memory ----------------------------------------------------------------------
  if (w_eqa && active) {
    for (int i = lane; i < 4 * 17 * 4; i += 32) (&tr.P[0][0][0])[i] = (&f->P[0][0][0])[i];
    for (int i = lane; i < 4 * 5 * 4; i += 32) (&tr.S[0][0][0])[i] = (&f->S[0][0][0])[i];
    if (lane < 4) {
      tr.cf[lane][0] = f->b0[lane];
      tr.cf[lane][1] = -f->a1[lane];
      tr.cf[lane][2] = -f->a2[lane];
      tr.cf[lane][3] = f->B1[lane];
      tr.cf[lane][4] = f->B2[lane];
      tr.st[lane][0] = make_float2(f->s1[0][lane], f->s1[1][lane]);
      tr.st[lane][1] = make_float2(f->s2[0][lane], f->s2[1][lane]);
    }
  }
  __syncthreads();

  const float att = f->att, rel = f->rel;
  const bool sel = f->sel != 0;
  const float sgn = sel ? 1.0f : -1.0f;
  const float a1m_s = sel ? f->a1m : -f->a1m, r1m_s = sel ? f->r1m : -f->r1m;
  const float thr = f->thr, makeup = f->makeup;
  const uint32_t code = f->ratio_code;

  const uint32_t spc = (B + 511u) / 512u;  // chunks per callback
  const uint32_t NC = K * spc;
  auto chunk_shape = [&](uint32_t i, uint32_t& k, uint32_t& f0, uint32_t& n) {
    if (spc == 1u) {  // the usual case (block <= 512 frames): no integer divisions on the per-chunk path
      k = i, f0 = 0u, n = B;
    } else {
      k = i / spc;
      f0 = (i % spc) * 512u;
      n = B - f0 < 512u ? B - f0 : 512u;
    }
  };
  float2* const tb = reinterpret_cast<float2*>(trackbuf) + (size_t)e * tbs;
  const bool tb_vec = (B & 1u) == 0;  // 4-frame groups of the track buffer are 16-byte aligned

  const int pair = lane;  // serial warp: lane = track * 2 + channel
  const bool ser_lane = w_serial && lane < 2 * TPC && comp_on;
  float env_s = ser_lane ? sgn * f->env[lane & 1] : 0.0f;  // follower state, signed domain
  float s4[5], s123[9];
#pragma unroll
  for (int i = 0; i < 5; i++) s4[i] = w_serial ? f->sl[9 + i] : 0.0f;
#pragma unroll
  for (int i = 0; i < 9; i++) s123[i] = (!w_serial && !w_eq) ? f->sl[i] : 0.0f;

  // chunks behind the render: EQ 1 (2 for the second EQ warp), the follower's intercepts 2 (built by the second EQ warp, or —
  // with one EQ warp per track — by the track's output warps, which have the slack), the serial warp 3, the output 4
  constexpr uint32_t LAG_Q = 2, LAG_S = 3, LAG_O = 4;
  // output warps, one slot per (callback, track): the cell of a chunk is pulled into L1 two iterations before it is read
  // (a prefetch instruction holds no register across the output stage; an early LOAD did, was spilled, and the spill store
  // waited for the data — 12 % of the kernel's stall samples), so the cell -> span -> source chain starts from an L1 hit
  auto prefetch_cell = [&](uint32_t chunk) {
    if (chunk < NC) {
      uint32_t pk, pf0, pn;
      chunk_shape(chunk, pk, pf0, pn);
      asm volatile("prefetch.global.L1 [%0];" ::"l"(cells + (size_t)pk * N + t));
    }
  };
  if (!w_serial && !w_eq && active && S == 1 && lane == 0 && ow == 0) {
    prefetch_cell(0);
    prefetch_cell(1);
  }
  for (uint32_t it = 0; it < NC + LAG_O; it++) {
    if (w_eq) {
      // ---- EQ: biquads 0, 1 of chunk it-1 (warp a) / biquads 2, 3 of chunk it-2, then the follower's intercepts (b) ----
      const uint32_t lag = (SH::EQW == 2 && w_eqb) ? 2u : 1u;  // with one EQ warp the slot after it is simply held
      if (active && it >= lag && it - lag < NC) {
        uint32_t k, f0, n;
        chunk_shape(it - lag, k, f0, n);
        float* X = tr.X[(it - lag) % 5u];
        const int par = (int)((it - lag) & 1);
        if (eq_on) {
          const int len = (int)n - 16 * lane < 0 ? 0 : ((int)n - 16 * lane > 16 ? 16 : (int)n - 16 * lane);
          const int last = ((int)n - 1) >> 4;  // lane holding the chunk's last frame
          float2 xs[16];
          float* seg = X + lane * FXC_SEG_STRIDE;
#pragma unroll
          for (int m = 0; m < 16; m += 2) {
            const float4 v = (m < len) ? *reinterpret_cast<const float4*>(seg + 2 * m) : make_float4(0.f, 0.f, 0.f, 0.f);
            xs[m] = make_float2(v.x, v.y);
            xs[m + 1] = make_float2(v.z, v.w);
          }
#pragma unroll 1
          for (int b = (SH::EQW == 2 && w_eqb) ? 2 : 0; b < ((SH::EQW == 2 && w_eqa) ? 2 : 4); b++) {
            const float2 b0 = f2(tr.cf[b][0]), na1 = f2(tr.cf[b][1]), na2 = f2(tr.cf[b][2]), B1 = f2(tr.cf[b][3]), B2 = f2(tr.cf[b][4]);
            // everything the scan and the state update read from shared memory is fetched HERE, under the zero-state pass:
            // inside the scan each of these loads sat on the critical path (load -> 2 dependent FMAs -> shuffle), and the
            // EQ warp is the long pole of the CTA's pipeline. The scan steps are selects, not branches, for the same reason.
            float4 sj[5];
#pragma unroll
            for (int j = 0; j < 5; j++) sj[j] = *reinterpret_cast<const float4*>(&tr.S[b][j][0]);  // A^(16 * 2^j)
            const float2 in1 = tr.st[b][0], in2 = tr.st[b][1];  // state entering the chunk (L, R)
            const float4 pl = *reinterpret_cast<const float4*>(&tr.P[b][len][0]);
            float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
            if (n == 512u) {  // every segment is full: no per-frame guards
#pragma unroll
              for (int m = 0; m < 16; m++) {
                const float2 xi = xs[m];
                xs[m] = __ffma2_rn(b0, xi, s1);
                const float2 tt = __ffma2_rn(B1, xi, s2);
                s2 = __ffma2_rn(na2, s1, __fmul2_rn(B2, xi));
                s1 = __ffma2_rn(na1, s1, tt);
              }
            } else {
#pragma unroll
              for (int m = 0; m < 16; m++) {
                if (m < len) {
                  const float2 xi = xs[m];
                  xs[m] = __ffma2_rn(b0, xi, s1);
                  const float2 tt = __ffma2_rn(B1, xi, s2);
                  s2 = __ffma2_rn(na2, s1, __fmul2_rn(B2, xi));
                  s1 = __ffma2_rn(na1, s1, tt);
                }
              }
            }
            const float2 e1 = s1, e2 = s2;
            float2 v1 = e1, v2 = e2;
            {  // the incoming state enters through segment 0
              const float2 w1 = __ffma2_rn(f2(sj[0].x), in1, __ffma2_rn(f2(sj[0].y), in2, e1));
              const float2 w2 = __ffma2_rn(f2(sj[0].z), in1, __ffma2_rn(f2(sj[0].w), in2, e2));
              if (lane == 0) v1 = w1, v2 = w2;
            }
#pragma unroll
            for (int j = 0; j < 5; j++) {  // Kogge-Stone: v[l] += A^(16 d) v[l - d]
              const unsigned d = 1u << j;
              const float2 u1 = shfl_up2(v1, d), u2 = shfl_up2(v2, d);
              const float2 n1 = __ffma2_rn(f2(sj[j].x), u1, __ffma2_rn(f2(sj[j].y), u2, v1));
              const float2 n2 = __ffma2_rn(f2(sj[j].z), u1, __ffma2_rn(f2(sj[j].w), u2, v2));
              if (lane >= (int)d) v1 = n1, v2 = n2;  // lanes below d computed on their own values: dropped
            }
            float2 st1 = shfl_up2(v1, 1), st2 = shfl_up2(v2, 1);
            if (lane == 0) st1 = in1, st2 = in2;
            // zero-input response of the segment's true start state
#pragma unroll
            for (int m = 0; m < 16; m++) {
              const float2 pm = *reinterpret_cast<const float2*>(&tr.P[b][m][0]);
              xs[m] = __ffma2_rn(f2(pm.y), st2, __ffma2_rn(f2(pm.x), st1, xs[m]));
            }
            // state after the chunk's last frame, from the lane that holds it (every lane evaluates it, one stores)
            const float2 ns1 = __ffma2_rn(f2(pl.x), st1, __ffma2_rn(f2(pl.y), st2, e1));
            const float2 ns2 = __ffma2_rn(f2(pl.z), st1, __ffma2_rn(f2(pl.w), st2, e2));
            __syncwarp();
            if (lane == last) {
              tr.st[b][0] = ns1;
              tr.st[b][1] = ns2;
            }
          }
#pragma unroll
          for (int m = 0; m < 16; m += 2)
            if (m < len) *reinterpret_cast<float4*>(seg + 2 * m) = make_float4(xs[m].x, xs[m].y, xs[m + 1].x, xs[m + 1].y);
          __syncwarp();
        }
        if (SH::EQW == 2 && comp_on && w_eqb) {
          const int nb = (int)n >> 2;
          float* qL = sm.Q4[par][2 * tl];
          float* qR = sm.Q4[par][2 * tl + 1];
          for (int g = lane; g < nb; g += 32) {
            const float* src = X + fxc_addr(4 * g);
            const float4 v0 = *reinterpret_cast<const float4*>(src), v1 = *reinterpret_cast<const float4*>(src + 4);
            const float2 x4[4] = {make_float2(v0.x, v0.y), make_float2(v0.z, v0.w), make_float2(v1.x, v1.y), make_float2(v1.z, v1.w)};
            FxcQ q;
            fxc_intercepts<4>(x4, att, rel, a1m_s, r1m_s, q);
#pragma unroll
            for (int i = 0; i < 5; i++) {
              qL[i * FXC_BLOCKS + g] = q.q4[i].x;
              qR[i * FXC_BLOCKS + g] = q.q4[i].y;
            }
          }
          const int tail = (int)n & 3;
          if (lane < tail) {  // the chunk's last 1..3 frames: (pa, pr) for the one-step form
            const float2 xv = *reinterpret_cast<const float2*>(X + fxc_addr(4 * nb + lane));
            tr.TP[par][0][lane][0] = __fmul_rn(a1m_s, fabsf(xv.x));
            tr.TP[par][0][lane][1] = __fmul_rn(r1m_s, fabsf(xv.x));
            tr.TP[par][1][lane][0] = __fmul_rn(a1m_s, fabsf(xv.y));
            tr.TP[par][1][lane][1] = __fmul_rn(r1m_s, fabsf(xv.y));
          }
        }
      }
    } else if (w_serial) {
      // ---- the follower's serial recurrence over chunk it-LAG_S: one step per block of 4 frames ---------------------
      if (ser_lane && it >= LAG_S && it - LAG_S < NC) {
        uint32_t k, f0, n;
        chunk_shape(it - LAG_S, k, f0, n);
        const int par = (int)((it - LAG_S) & 1);
        const int nb = (int)n >> 2;
        const float* Q = sm.Q4[par][pair];
        float* E = sm.E[par][pair];
        float ev = env_s;
        E[3] = ev;  // the envelope entering the chunk
        const int ng = nb >> 2;  // groups of 4 blocks: 5 x 128-bit loads, 1 x 128-bit store each
        float4 qa[5], qb[5];
        if (ng > 0) {
#pragma unroll
          for (int i = 0; i < 5; i++) qa[i] = *reinterpret_cast<const float4*>(Q + i * FXC_BLOCKS);
        }
        for (int j = 0; j < ng; j += 2) {
          if (j + 1 < ng) {
#pragma unroll
            for (int i = 0; i < 5; i++) qb[i] = *reinterpret_cast<const float4*>(Q + i * FXC_BLOCKS + 4 * (j + 1));
          }
          {
            float4 o;
            o.x = ev = fxc_step(ev, s4, qa[0].x, qa[1].x, qa[2].x, qa[3].x, qa[4].x);
            o.y = ev = fxc_step(ev, s4, qa[0].y, qa[1].y, qa[2].y, qa[3].y, qa[4].y);
            o.z = ev = fxc_step(ev, s4, qa[0].z, qa[1].z, qa[2].z, qa[3].z, qa[4].z);
            o.w = ev = fxc_step(ev, s4, qa[0].w, qa[1].w, qa[2].w, qa[3].w, qa[4].w);
            *reinterpret_cast<float4*>(E + 4 + 4 * j) = o;
          }
          if (j + 1 < ng) {
            if (j + 2 < ng) {
#pragma unroll
              for (int i = 0; i < 5; i++) qa[i] = *reinterpret_cast<const float4*>(Q + i * FXC_BLOCKS + 4 * (j + 2));
            }
            float4 o;
            o.x = ev = fxc_step(ev, s4, qb[0].x, qb[1].x, qb[2].x, qb[3].x, qb[4].x);
            o.y = ev = fxc_step(ev, s4, qb[0].y, qb[1].y, qb[2].y, qb[3].y, qb[4].y);
            o.z = ev = fxc_step(ev, s4, qb[0].z, qb[1].z, qb[2].z, qb[3].z, qb[4].z);
            o.w = ev = fxc_step(ev, s4, qb[0].w, qb[1].w, qb[2].w, qb[3].w, qb[4].w);
            *reinterpret_cast<float4*>(E + 4 + 4 * (j + 1)) = o;
          }
        }
        for (int g = 4 * ng; g < nb; g++) {  // up to 3 blocks left over
          ev = fxc_step(ev, s4, Q[g], Q[FXC_BLOCKS + g], Q[2 * FXC_BLOCKS + g], Q[3 * FXC_BLOCKS + g], Q[4 * FXC_BLOCKS + g]);
          E[4 + g] = ev;
        }
        const int tail = (int)n & 3;
        FxcTrack& mt = sm.tr[pair >> 1];
        for (int j = 0; j < tail; j++) {
          ev = fmaxf(__fmaf_rn(rel, ev, mt.TP[par][pair & 1][j][1]), __fmaf_rn(att, ev, mt.TP[par][pair & 1][j][0]));
          mt.ET[par][pair & 1][j] = ev;
        }
        env_s = ev;
      }
    } else {
      // ---- output warps: (1) loads of chunk `it` in flight, (2) output of chunk it-3, (3) chunk `it` -> X -----------
      constexpr int RQ = 512 / SH::OLANES;  // frames per lane of a chunk
      const int ol = ow * 32 + lane;        // lane among the track's output lanes
      uint32_t rk = 0, rf0 = 0, rn = 0;
      float2 rv[RQ];
      int rmode = 0;  // 0 nothing, 1 scaled copy of rv, 2 generic per-frame path, 3 scaled copy of a whole chunk
      uint32_t rlo = 0, rhi = 0;
      float rgain = 0.0f;
      if (active && it < NC) {
        chunk_shape(it, rk, rf0, rn);
        rmode = 2;
        if (S == 1) {
          if (lane == 0 && ow == 0) prefetch_cell(it + 2);
          const DCell cell = cells[(size_t)rk * N + t];  // in L1 since two iterations ago
          if (cell.span == kSilent) {
            rmode = 1;  // rlo == rhi: zeros
#pragma unroll
            for (int q = 0; q < RQ; q++) rv[q] = make_float2(0.0f, 0.0f);
          } else {
            // the fields of the span that decide the path, fetched together (four independent 16-byte loads, one round trip):
            // tested one after the other through the pointer, each && was a dependent load + branch on the path that
            // issues this chunk's clip loads
            static_assert(offsetof(DSpan, base) == 0 && offsetof(DSpan, speed) == 16 && offsetof(DSpan, gain) == 32 &&
                              offsetof(DSpan, dst_off) == 48 && offsetof(DSpan, fmt) == 56 && offsetof(DSpan, nch) == 64 &&
                              offsetof(DSpan, fade) == 68,
                          "DSpan layout");
            const int4* spq = reinterpret_cast<const int4*>(spans + cell.span);
            const int4 q0 = __ldg(spq), q1 = __ldg(spq + 1), q2 = __ldg(spq + 2), q3 = __ldg(spq + 3);
            const int4 q4 = __ldg(spq + 4);
            const void* sp_base = reinterpret_cast<const void*>(((uint64_t)(uint32_t)q0.y << 32) | (uint32_t)q0.x);
            const double sp_speed = __hiloint2double(q1.y, q1.x);
            const float sp_gain = __int_as_float(q2.x);
            const uint32_t sp_dst_off = (uint32_t)q3.x, sp_fmt = (uint32_t)q3.z, sp_nch = (uint32_t)q4.x, sp_fade = (uint32_t)q4.y;
            if (sp_fmt == F_F32 && sp_nch == 2 && sp_speed == 1.0 && sp_fade == 0 && C == 2) {
              // unity-speed stereo f32 clip: a scaled copy (src * gain, then the add into the cleared buffer: 0 + m)
              rmode = 1;
              rlo = sp_dst_off, rhi = sp_dst_off + cell.n_act;
              rgain = sp_gain;
              const float2* src = reinterpret_cast<const float2*>(sp_base) + (int64_t)(uint32_t)(int64_t)cell.pos;
              if (rn == 512u && rlo <= rf0 && rhi >= rf0 + 512u) {  // a whole chunk inside the clip: no per-frame range checks
                rmode = 3;
                const float2* s0 = src + (rf0 - rlo) + ol;
#pragma unroll
                for (int q = 0; q < RQ; q++) rv[q] = __ldg(s0 + q * SH::OLANES);
              } else {
#pragma unroll
                for (int q = 0; q < RQ; q++) {
                  const uint32_t fr = q * SH::OLANES + ol, j = rf0 + fr;
                  rv[q] = (fr < rn && j >= rlo && j < rhi) ? __ldg(src + (j - rlo)) : make_float2(0.0f, 0.0f);
                }
              }
            }
          }
        }
      }
      if (SH::EQW == 1 && comp_on && it >= LAG_Q && it - LAG_Q < NC) {
        // the follower's intercepts of chunk it-2 (its EQ finished last iteration): Q4 per block of 4 frames
        uint32_t k, f0, n;
        chunk_shape(it - LAG_Q, k, f0, n);
        const float* X = tr.X[(it - LAG_Q) % 5u];
        const int par = (int)((it - LAG_Q) & 1);
        const int nb = (int)n >> 2;
        float* qL = sm.Q4[par][2 * tl];
        float* qR = sm.Q4[par][2 * tl + 1];
        for (int g = ol; g < nb; g += SH::OLANES) {
          const float* src = X + fxc_addr(4 * g);
          const float4 v0 = *reinterpret_cast<const float4*>(src), v1 = *reinterpret_cast<const float4*>(src + 4);
          const float2 x4[4] = {make_float2(v0.x, v0.y), make_float2(v0.z, v0.w), make_float2(v1.x, v1.y), make_float2(v1.z, v1.w)};
          FxcQ q;
          fxc_intercepts<4>(x4, att, rel, a1m_s, r1m_s, q);
#pragma unroll
          for (int i = 0; i < 5; i++) {
            qL[i * FXC_BLOCKS + g] = q.q4[i].x;
            qR[i * FXC_BLOCKS + g] = q.q4[i].y;
          }
        }
        const int tail = (int)n & 3;
        if (ol < tail) {  // the chunk's last 1..3 frames: (pa, pr) for the one-step form
          const float2 xv = *reinterpret_cast<const float2*>(X + fxc_addr(4 * nb + ol));
          tr.TP[par][0][ol][0] = __fmul_rn(a1m_s, fabsf(xv.x));
          tr.TP[par][0][ol][1] = __fmul_rn(r1m_s, fabsf(xv.x));
          tr.TP[par][1][ol][0] = __fmul_rn(a1m_s, fabsf(xv.y));
          tr.TP[par][1][ol][1] = __fmul_rn(r1m_s, fabsf(xv.y));
        }
      }
      if (active && it >= LAG_O) {
        uint32_t k, f0, n;
        chunk_shape(it - LAG_O, k, f0, n);
        const float* X = tr.X[(it - LAG_O) % 5u];
        const int par = (int)((it - LAG_O) & 1);
        const int nb = (int)n >> 2;
        float2* dst = tb + (size_t)k * B + f0;
        const float* EL = sm.E[par][2 * tl];
        const float* ER = sm.E[par][2 * tl + 1];
#pragma unroll
        for (int u = 0; u < FXC_BLOCKS / SH::OLANES; u++) {
          const int g = ol + u * SH::OLANES;  // block of 4 frames
          if (g < nb) {
            const float* src = X + fxc_addr(4 * g);
            const float4 v0 = *reinterpret_cast<const float4*>(src), v1 = *reinterpret_cast<const float4*>(src + 4);
            float2 y[4] = {make_float2(v0.x, v0.y), make_float2(v0.z, v0.w), make_float2(v1.x, v1.y), make_float2(v1.z, v1.w)};
            if (comp_on) {
              FxcQ q;
              fxc_intercepts<3>(y, att, rel, a1m_s, r1m_s, q);
              const float2 e0 = make_float2(EL[3 + g], ER[3 + g]);
              float2 ev[4];
              ev[0] = fmax2(__ffma2_rn(f2(s123[0]), e0, q.q1[0]), __ffma2_rn(f2(s123[1]), e0, q.q1[1]));
              float2 tt = __ffma2_rn(f2(s123[2]), e0, q.q2[0]);
#pragma unroll
              for (int i = 1; i < 3; i++) tt = fmax2(tt, __ffma2_rn(f2(s123[2 + i]), e0, q.q2[i]));
              ev[1] = tt;
              tt = __ffma2_rn(f2(s123[5]), e0, q.q3[0]);
#pragma unroll
              for (int i = 1; i < 4; i++) tt = fmax2(tt, __ffma2_rn(f2(s123[5 + i]), e0, q.q3[i]));
              ev[2] = tt;
              ev[3] = make_float2(EL[4 + g], ER[4 + g]);
              // the gain computer only does work above the threshold: one branch per block instead of one per sample
              // (below it fx_gain is (x * 1) * makeup = x * makeup exactly)
              bool hot = false;
#pragma unroll
              for (int m = 0; m < 4; m++) {
                ev[m] = __fmul2_rn(f2(sgn), ev[m]);  // back from the signed domain (exact)
                hot = hot || ev[m].x > thr || ev[m].y > thr;
              }
              if (hot) {
#pragma unroll
                for (int m = 0; m < 4; m++) {
                  y[m].x = fx_gain(y[m].x, ev[m].x, thr, makeup, code);
                  y[m].y = fx_gain(y[m].y, ev[m].y, thr, makeup, code);
                }
              } else {
#pragma unroll
                for (int m = 0; m < 4; m++) y[m] = __fmul2_rn(y[m], f2(makeup));
              }
            }
            if (tb_vec) {
              reinterpret_cast<float4*>(dst + 4 * g)[0] = make_float4(y[0].x, y[0].y, y[1].x, y[1].y);
              reinterpret_cast<float4*>(dst + 4 * g)[1] = make_float4(y[2].x, y[2].y, y[3].x, y[3].y);
            } else {
#pragma unroll
              for (int m = 0; m < 4; m++) dst[4 * g + m] = y[m];
            }
          }
        }
        if (ol == 0) {  // the chunk's last 1..3 frames
          const int tail = (int)n & 3;
          for (int j = 0; j < tail; j++) {
            float2 yv = *reinterpret_cast<const float2*>(X + fxc_addr(4 * nb + j));
            if (comp_on) {
              yv.x = fx_gain(yv.x, sgn * tr.ET[par][0][j], thr, makeup, code);
              yv.y = fx_gain(yv.y, sgn * tr.ET[par][1][j], thr, makeup, code);
            }
            dst[4 * nb + j] = yv;
          }
        }
      }
      if (rmode == 3) {
        float* X = tr.X[it % 5u] + fxc_addr(ol);  // frame q * OLANES + ol sits q * (OLANES / 16) segments further on
#pragma unroll
        for (int q = 0; q < RQ; q++) {
          float2 ov;
          ov.x = __fadd_rn(0.0f, __fmul_rn(rv[q].x, rgain));
          ov.y = __fadd_rn(0.0f, __fmul_rn(rv[q].y, rgain));
          *reinterpret_cast<float2*>(X + q * (SH::OLANES / 16) * FXC_SEG_STRIDE) = ov;
        }
      } else if (rmode) {
        float* X = tr.X[it % 5u];
#pragma unroll
        for (int q = 0; q < RQ; q++) {
          const uint32_t fr = q * SH::OLANES + ol, j = rf0 + fr;
          if (fr < rn) {
            float2 ov = make_float2(0.0f, 0.0f);
            if (rmode == 1) {
              if (j >= rlo && j < rhi) {
                ov.x = __fadd_rn(0.0f, __fmul_rn(rv[q].x, rgain));
                ov.y = __fadd_rn(0.0f, __fmul_rn(rv[q].y, rgain));
              }
            } else {
              ov = fxc_render_frame(spans, cells, rk, t, N, S, j, C, poly);
            }
            *reinterpret_cast<float2*>(X + fxc_addr((int)fr)) = ov;
          }
        }
      }
    }
    __syncthreads();
  }

  // ---- state back ---------------------------------------------------------------------------------------------------
  if (w_eq && eq_on && lane < 4) {
    f->s1[0][lane] = tr.st[lane][0].x, f->s2[0][lane] = tr.st[lane][1].x;
    if (C == 2) f->s1[1][lane] = tr.st[lane][0].y, f->s2[1][lane] = tr.st[lane][1].y;
  }
  if (ser_lane && (lane & 1) < (int)C) f->env[lane & 1] = sgn * env_s;
}

template <int TPC, int OW, int EQW, int MINB>
static cudaError_t launch_effects_chain_t(const DSpan* spans, const DCell* cells, DFx* fx, uint32_t n_fx, uint32_t N, uint32_t S,
                                          uint32_t K, uint32_t B, uint32_t C, const float* poly, float* trackbuf, uint64_t tbs,
                                          cudaStream_t stream) {
  using SH = FxcShape<TPC, OW, EQW>;
  auto kfn = fx_chain_kernel<TPC, OW, EQW, MINB>;
  const size_t smem = sizeof(FxcSmem<TPC>);
  cudaError_t err = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  // slots in role order: serial, EQ a per track, EQ b per track, output warps
  FxcRoles roles;
  for (int i = 0; i < 32; i++) roles.code[i] = 255;
  uint8_t want[32];  // every role of the shape, in role order: serial, EQ a per track, EQ b per track, output warps
  int nw = 0;
  want[nw++] = 0;
  for (int t = 0; t < TPC; t++) want[nw++] = (uint8_t)(1 + t);
  if (EQW == 2)
    for (int t = 0; t < TPC; t++) want[nw++] = (uint8_t)(9 + t);
  for (int o = 0; o < TPC * OW; o++) want[nw++] = (uint8_t)(32 + o);
  for (int i = 0; i < nw; i++) roles.code[i] = want[i];
  if (SH::SLOTS == 16 && nw == 13) {
    // the 4-track CTA (13 warps in 16 slots) over the four sub-partitions (slot % 4): the serial warp next to two output
    // warps (which spend much of an iteration waiting for memory), the EQ warps in pairs with one output warp, the other
    // output warps pooled. Measured against role order and three other placements: profiles/r02c_fx_roles.log.
    //   0: serial out out -   1: EQ0 EQ1 out -   2: EQ2 EQ3 out -   3: out out out out
    static const uint8_t bal[16] = {0, 1, 3, 32, 33, 2, 4, 34, 35, 36, 37, 38, 255, 255, 255, 39};
    for (int i = 0; i < 16; i++) roles.code[i] = bal[i];
    if (const char* env = getenv("WBX_FX_LAYOUT")) {  // experiments: 16 comma-separated codes, every role exactly once
      int v[16], got = 0;
      for (const char* q = env; got < 16 && *q;) {
        v[got++] = atoi(q);
        while (*q && *q != ',') q++;
        if (*q == ',') q++;
      }
      bool ok = got == 16;
      for (int j = 0; j < nw && ok; j++) {
        int c = 0;
        for (int i = 0; i < 16; i++) c += v[i] == want[j];
        ok = c == 1;
      }
      int live = 0;
      for (int i = 0; i < 16 && ok; i++) live += v[i] != 255;
      if (!ok || live != nw) return cudaErrorInvalidValue;
      for (int i = 0; i < 16; i++) roles.code[i] = (uint8_t)v[i];
    }
  }
  kfn<<<(n_fx + TPC - 1) / TPC, SH::THREADS, smem, stream>>>(spans, cells, fx, n_fx, N, S, K, B, C, poly, trackbuf, tbs, roles);
  return cudaGetLastError();
}

// Shape = (tracks per CTA, output warps per track, EQ warps per track, CTAs per SM). The serial warp walks one instruction
// stream for all pairs of its CTA, so a CTA should carry as many tracks as it takes to give every SM its share; several
// small CTAs per SM decouple their per-chunk barriers. WBX_FX_SHAPE="tpc,ow,eqw,ctas" picks a compiled shape by hand.
static cudaError_t launch_effects_chain(const DSpan* spans, const DCell* cells, DFx* fx, uint32_t n_fx, uint32_t N, uint32_t S,
                                        uint32_t K, uint32_t B, uint32_t C, const float* poly, float* trackbuf,
                                        uint64_t tbs, int n_sm, cudaStream_t stream) {
  static_assert(sizeof(FxcSmem<4>) <= 227 * 1024, "fx_chain_kernel shared memory");
  static_assert(2 * sizeof(FxcSmem<2>) + 2048 <= 227 * 1024 && 4 * sizeof(FxcSmem<1>) + 4096 <= 227 * 1024, "CTAs per SM");
  int shape = n_fx <= (uint32_t)n_sm ? 1410 : (n_fx <= 2u * (uint32_t)n_sm ? 2410 : 4210);
  if (const char* env = getenv("WBX_FX_SHAPE")) {
    int a = 0, b = 0, c = 0, d = 0;
    if (sscanf(env, "%d,%d,%d,%d", &a, &b, &c, &d) == 4) shape = a * 1000 + b * 100 + c * 10 + (d - 1);
  }
#define WBX_FX_CASE(TPC, OW, EQW, MINB) \
  case TPC * 1000 + OW * 100 + EQW * 10 + (MINB - 1): \
    return launch_effects_chain_t<TPC, OW, EQW, MINB>(spans, cells, fx, n_fx, N, S, K, B, C, poly, trackbuf, tbs, stream)
  switch (shape) {
    WBX_FX_CASE(1, 4, 1, 1);
    WBX_FX_CASE(2, 4, 1, 1);
    WBX_FX_CASE(4, 2, 1, 1);
    WBX_FX_CASE(4, 2, 2, 1);
    WBX_FX_CASE(2, 2, 1, 2);
    WBX_FX_CASE(2, 2, 2, 2);
    WBX_FX_CASE(1, 2, 1, 4);
    WBX_FX_CASE(1, 2, 2, 4);
    WBX_FX_CASE(1, 4, 2, 2);
    default: return cudaErrorInvalidValue;
  }
#undef WBX_FX_CASE
}

// ---- convolution reverb (extension, cfg 5): direct form on the CUDA cores --------------------------------
// xin  [n_fx][C][H + T] planar: H = taps - 1 history frames (oldest first) followed by this render's T chain outputs
// hist [n_tracks][2][H] persists across renders (indexed by track, so it survives chain list rebuilds)
__global__ void fir_gather_kernel(const DFx* __restrict__ fx, uint32_t n_fx, uint32_t C, uint64_t H, uint64_t T,
                                  const float* __restrict__ hist, uint64_t hist_pos, const float* __restrict__ trackbuf,
                                  uint64_t tbs, float* __restrict__ xin, uint32_t* __restrict__ max_word) {
  const uint32_t ec = blockIdx.y;
  const uint32_t e = ec / C, c = ec % C;
  if (!fx[e].reverb_on) return;
  const float* h = hist + ((size_t)fx[e].track * 2 + c) * H;
  const float* tb = trackbuf + (size_t)e * tbs * 2 + c;
  float* x = xin + (size_t)ec * (H + T);
  float m = 0.0f;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < H + T; i += (uint64_t)gridDim.x * blockDim.x) {
    float v;
    if (i < H) {  // the history is a ring: logical index i lives at (hist_pos + i) mod H
      uint64_t r = hist_pos + i;
      if (r >= H) r -= H;
      v = h[r];
    } else {
      v = tb[(i - H) * 2];
    }
    x[i] = v;
    m = fmaxf(m, fabsf(v));
  }
  if (max_word) {  // the tensor-core path scales its fp16 operands by the largest input magnitude of the render
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(max_word, __float_as_uint(m));
  }
}

// 256 consecutive outputs of one (track, channel) per CTA; taps in tiles of 256 staged in shared memory
__global__ void __launch_bounds__(256) fir_kernel(const DFx* __restrict__ fx, uint32_t C, uint64_t H, uint64_t T,
                                                   const float* __restrict__ ir, uint32_t L,
                                                   const float* __restrict__ xin, float* __restrict__ trackbuf, uint64_t tbs) {
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
  if (n < (int64_t)T) trackbuf[((size_t)e * tbs + n) * 2 + c] = (float)total;
}

// the history ring takes the render's newest min(T, H) inputs (from the gather buffer: the outputs overwrote trackbuf):
// logical index i in [H - n, H) of the NEW history lives at (new_pos + i) mod H
__global__ void fir_save_kernel(const DFx* __restrict__ fx, uint32_t n_fx, uint32_t C, uint64_t H, uint64_t T,
                                const float* __restrict__ xin, float* __restrict__ hist, uint64_t new_pos) {
  const uint32_t ec = blockIdx.y;
  const uint32_t e = ec / C, c = ec % C;
  if (!fx[e].reverb_on) return;
  float* h = hist + ((size_t)fx[e].track * 2 + c) * H;
  const uint64_t n = T < H ? T : H;
  const float* x = xin + (size_t)ec * (H + T) + (H + T - n);  // the last n entries
  for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < n; u += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t r = new_pos + (H - n + u);
    if (r >= H) r -= H;
    h[r] = x[u];
  }
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

// The whole exchange after a rank's mix in ONE launch (one process or thread per GPU: wbx_mix_sharded): signal this rank's
// arrival, wait for every rank's, reduce this rank's slice of callbacks in rank order + clamp into the master bus (and the
// host output), then — the last block to finish — signal completion and wait for every rank's. Five launches' worth of
// latency become one. Blocks only ever spin on REMOTE flags, so no co-residency of the grid is assumed.
template <int V>
__global__ void __launch_bounds__(256)
shard_exchange_kernel(const float* __restrict__ xchg, uint32_t W, uint32_t C, uint64_t plane, uint64_t valid, ShardPeers peers,
                      uint64_t chan_stride, uint64_t dst_off, uint32_t rank, uint32_t epoch_arrive, uint32_t epoch_done,
                      unsigned long long timeout_ns, volatile uint32_t* status, uint32_t* __restrict__ done_counter) {
  __shared__ uint32_t is_last;
  const uint32_t tid = threadIdx.x;
  auto wait_all = [&](uint32_t epoch) {
    if (tid < W) {
      volatile uint32_t* mine = peers.flags[rank] + tid;
      unsigned long long t0, t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      while ((int32_t)(*mine - epoch) < 0) {
        __nanosleep(64);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > timeout_ns) {
          *status = 1u;
          break;
        }
      }
    }
    __threadfence_system();
    __syncthreads();
  };
  if (blockIdx.x == 0) {  // this rank's tiles are all on their way (the mix kernel before this launch has completed)
    __threadfence_system();
    if (tid < W) *(volatile uint32_t*)(peers.flags[tid] + rank) = epoch_arrive;
  }
  wait_all(epoch_arrive);
  const uint64_t n = valid / V;
  for (uint32_t c = 0; c < C; c++) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + tid; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
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
  // the last block to get here publishes this rank's completion and waits for everybody's
  __threadfence_system();
  __syncthreads();
  if (tid == 0) is_last = atomicAdd(done_counter, 1u) == gridDim.x - 1 ? 1u : 0u;
  __syncthreads();
  if (is_last) {
    __threadfence_system();
    if (tid < W) *(volatile uint32_t*)(peers.flags[tid] + rank) = epoch_done;
    wait_all(epoch_done);
    if (tid == 0) *done_counter = 0u;
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
  const int extra = EXT ? POLY_SMEM_BYTES : 0;  // the polyphase table behind the warps' regions
  const int smem = L::WARP_BYTES * WARPS + extra;
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
  kfn<<<(unsigned)ctas, wpc * 32, (size_t)L::WARP_BYTES * wpc + extra, stream>>>(p);
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
  if ((uint64_t)p.n_blocks * p.n_tracks * 2u >= (1ull << 32)) return cudaErrorInvalidValue;  // the kernel indexes peaks in 32 bits
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
                          const float* xin, void* scratch, float* trackbuf, uint64_t tbs, int n_sm, cudaStream_t stream);  // wbx_fir_tc.cu
uint32_t* fir_tc_max_word(void* scratch);
cudaError_t launch_fir_fft(const DFx* fx, uint32_t n_fx, uint32_t C, uint64_t T, const FirLaunch& a, float* trackbuf, uint64_t tbs,
                           cudaStream_t stream);  // wbx_fir_fft.cu

cudaError_t launch_effects(const DSpan* spans, DCell* cells, DFx* fx, uint32_t n_fx, uint32_t N, uint32_t S, uint32_t K,
                           uint32_t B, uint32_t C, uint32_t first_fx_span, float* trackbuf, const FirLaunch& fir, const float* poly,
                           uint32_t fx_flags, uint32_t* sm_arrivals, uint64_t tbs, cudaStream_t stream) {
  if (n_fx == 0) return cudaSuccess;
  const uint64_t warps = (uint64_t)n_fx * K;
  // chains with an EQ or a compressor: render + chain fused and time-parallel (fx_chain_kernel); chains that only end in
  // the reverb: the plain render of the track
  if (fx_flags & 2u)
    render_tracks_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, stream>>>(spans, cells, fx, n_fx, N, S, K, B, C, poly, trackbuf, tbs);
  if (fx_flags & 1u) {
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t err = launch_effects_chain(spans, cells, fx, n_fx, N, S, K, B, C, poly, trackbuf, tbs, n_sm, stream);
    if (err != cudaSuccess) return err;
  }
  const uint32_t L = fir.L;
  if (L && fir.ir && fir.hist && fir.mode == 2 && fir.ir_aux && fir.scratch && fir.fft_ring) {
    // convolution reverb as the chain's last stage: partitioned FFT convolution (wbx_fir_fft.cu), straight from the
    // history ring and trackbuf
    cudaError_t err = launch_fir_fft(fx, n_fx, C, (uint64_t)K * B, fir, trackbuf, tbs, stream);
    if (err != cudaSuccess) return err;
  } else if (L && fir.ir && fir.hist && fir.xin) {  // direct form: CUDA cores or tensor cores, from the gather buffer
    const uint64_t T = (uint64_t)K * B, H = L - 1;
    const dim3 gcopy((unsigned)(((H + T) + 255) / 256 < 4096 ? ((H + T) + 255) / 256 : 4096), n_fx * C);
    const bool tc = fir.mode == 1 && fir.ir_aux && fir.scratch;
    uint32_t* max_word = tc ? fir_tc_max_word(fir.scratch) : nullptr;
    if (max_word) {
      cudaError_t err = cudaMemsetAsync(max_word, 0, sizeof(uint32_t), stream);
      if (err != cudaSuccess) return err;
    }
    fir_gather_kernel<<<gcopy, 256, 0, stream>>>(fx, n_fx, C, H, T, fir.hist, fir.hist_pos, trackbuf, tbs, fir.xin, max_word);
    if (tc) {  // tensor-core path (wbx_fir_tc.cu)
      int dev = 0, n_sm = 148;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
      cudaError_t err = launch_fir_tc(fx, n_fx, C, H, T, L, fir.ir_aux, fir.xin, fir.scratch, trackbuf, tbs, n_sm, stream);
      if (err != cudaSuccess) return err;
    } else {
      fir_kernel<<<dim3((unsigned)((T + 255) / 256), n_fx * C), 256, 0, stream>>>(fx, C, H, T, fir.ir, L, fir.xin, trackbuf, tbs);
    }
    if (H) {
      const uint64_t n = T < H ? T : H;
      fir_save_kernel<<<dim3((unsigned)((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024), n_fx * C), 256, 0, stream>>>(
          fx, n_fx, C, H, T, fir.xin, fir.hist, (fir.hist_pos + T) % H);
    }
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

cudaError_t launch_shard_exchange(const float* xchg, uint32_t W, uint32_t C, uint64_t plane, uint64_t valid, const ShardPeers& peers,
                                  uint64_t chan_stride, uint64_t dst_off, uint32_t rank, uint32_t epoch_arrive, uint32_t epoch_done,
                                  unsigned long long timeout_ns, uint32_t* status, uint32_t* done_counter, int n_sm,
                                  cudaStream_t stream) {
  bool v4 = (plane % 4 == 0) && (valid % 4 == 0) && (chan_stride % 4 == 0) && (dst_off % 4 == 0);
  for (int c = 0; c < 2; c++)
    if (peers.host_dst[c] && ((uintptr_t)peers.host_dst[c] & 15u)) v4 = false;
  const uint64_t n = v4 ? valid / 4 : valid;
  uint64_t blocks = (n + 255) / 256;
  if (blocks > (uint64_t)n_sm * 4) blocks = (uint64_t)n_sm * 4;
  if (blocks < 1) blocks = 1;
  if (v4)
    shard_exchange_kernel<4><<<(unsigned)blocks, 256, 0, stream>>>(xchg, W, C, plane, valid, peers, chan_stride, dst_off, rank,
                                                                   epoch_arrive, epoch_done, timeout_ns, status, done_counter);
  else
    shard_exchange_kernel<1><<<(unsigned)blocks, 256, 0, stream>>>(xchg, W, C, plane, valid, peers, chan_stride, dst_off, rank,
                                                                   epoch_arrive, epoch_done, timeout_ns, status, done_counter);
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
