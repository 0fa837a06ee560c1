// wbx_device.cuh — device-side data model of the mixing hot path (internal; the ABI is include/wbx.h).
//
// HBM layout
//   samples   one allocation per Sample (dsp/sample.h:18-28). The host hands planar channels over the ABI;
//             on the device a sample is stored FRAME-INTERLEAVED (L0 R0 L1 R1 ...; mono stays as is) so that
//             the window one callback needs from one stereo track (512 frames = 4 KiB of f32) is ONE
//             contiguous, 16-byte-aligned run = one TMA bulk copy, and a 128-bit shared-memory load yields two
//             (L, R) pairs for packed f32x2 math. Only the first min(channels, 2) source channels are kept
//             (output channel c reads source channel c % channels and the bus has at most 2 channels,
//             dsp/sampler.cpp:111, engine/track.h:50). Base 256-B aligned, >= 16 zero frames after `frames`
//             (dsp/sample.cpp:127,140) plus slack so aligned windows never leave the allocation.
//   spans     one DSpan per wbx_segment (L2-resident; reused by every block of a run).
//   cells     [n_blocks][n_tracks][slots] DCell, 16 B each: the state of one Sampler::stream call
//             (position in f64, clipped length) written by the schedule-expansion kernel. 0.4 % of the
//             audio bytes it describes.
//   bus       [out_channels][n_blocks * block_frames] f32 (AudioBuffer layout, core/audio_buffer.h:19-23)
//   peaks     [n_blocks][n_tracks][2] f32 (VUMeter block peaks, engine/vu_meter.h:20-30)
#pragma once
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define WBX_HD __host__ __device__
#else
#define WBX_HD
#endif

namespace wbx {

constexpr uint32_t kSilent = 0xFFFFFFFFu;

// The sampler's position recurrence `off = fl(off + adv)` (dsp/sampler.cpp:103,209: one rounding per callback), run for
// up to n callbacks while off < limit; returns the number of steps taken — exactly what the step-by-step loop
//     while (steps < n && off < limit) { off = off + adv; steps++; }
// produces, bit for bit, in O(binades crossed) instead of O(n). Inside one binade [2^e, 2^(e+1)) every value is a
// multiple of u = ulp and fl(x + adv) = x + A*u with A = adv/u rounded to nearest — a fixed integer unless adv/u is
// exactly halfway (then round-to-even depends on x and the steps are taken for real) — so m steps are x + m*A*u, exact
// in 64-bit integers. The jump stops a safe margin before the binade's end and before `limit`; the steps in between are
// real additions. Shared by the host scheduler (wbx_host.cpp) and expand_schedule; both are built without FMA contraction.
WBX_HD inline uint32_t advance_rounded_impl(double* off_io, double adv, uint32_t n, double limit) {
  double x = *off_io;
  uint32_t steps = 0;
  double no_jump_below = 0.0;  // end of the binade in which a jump was last refused: real steps until then
  while (steps < n && x < limit) {
    bool jumped = false;
    if (n - steps > 4 && x >= no_jump_below && x > 0.0 && adv > 0.0) {
      uint64_t xb;
      memcpy(&xb, &x, 8);
      const uint64_t ex = (xb >> 52) & 0x7FFu;  // biased exponent: x in [2^(ex-1023), 2^(ex-1022))
      if (ex > 54 && ex < 0x7FEu) {
        const uint64_t ub = (ex - 52) << 52, hb = (ex + 1) << 52, ib = (2046 - (ex - 52)) << 52;
        double u, hi, inv_u;
        memcpy(&u, &ub, 8);      // ulp of the binade
        memcpy(&hi, &hb, 8);     // binade end
        memcpy(&inv_u, &ib, 8);  // 1 / u (a power of two as well)
        const double qa = adv * inv_u;  // scaling by a power of two: exact
        if (qa >= 16.0 && qa < 4503599627370496.0) {  // 16 <= adv/ulp < 2^52
          const uint64_t fl = (uint64_t)qa;            // floor (qa > 0)
          const double frac = qa - (double)fl;         // exact
          if (frac != 0.5) {
            const uint64_t A = frac > 0.5 ? fl + 1 : fl;  // adv/u rounded to nearest
            const double au = (double)A * u;              // exact
            // every closed-form step must keep the exact sum x_k + adv below the binade end and x_k below limit:
            // x + (m-1)*au + adv < end  <=  m <= (end - x - adv - au) / au (one step size of slack: au >= 16 ulps,
            // the rounding slop of this expression and of the division below is under 2)
            const double room = (hi < limit ? hi : limit) - x - adv - au;
            if (room > 0.0) {
              double m = room / au;
              const double left = (double)(n - steps);
              if (m > left) m = left;
              const uint64_t mi = (uint64_t)m;
              if (mi >= 1) {
                const uint64_t X = (uint64_t)(x * inv_u);  // < 2^53, exact
                x = (double)(X + mi * A) * u;          // <= 2^53 ulps, exact
                steps += (uint32_t)mi;
                jumped = true;
              }
            }
          }
        }
        no_jump_below = hi;  // refused, or taken as far as this binade allows: real steps up to the binade's end
      }
    }
    if (!jumped) {
      x = x + adv;
      steps++;
    }
  }
  *off_io = x;
  return steps;
}

struct __align__(16) DSpan {
  const void* base;   // frame-interleaved sample data
  double pos0;        // Sampler::sample_offset_ at the first call
  double speed;       // Sampler::playback_speed_
  uint64_t count;     // Sample::count
  float gain;         // AudioClip::gain
  uint32_t track;
  uint32_t block0, n_blocks;
  uint32_t dst_off, length;
  uint32_t fmt;   // wbx_format
  uint32_t slot;  // which of the `slots` cells of (block, track) this span writes
  uint32_t nch;   // channels stored on the device (1 or 2)
  uint32_t fade;  // segment flags: bit 0 WBX_SEG_FADE (the envelope fields below apply), bit 1 WBX_SEG_POLYPHASE
  double clip_frame, fade_in, fade_out, clip_len;  // fade extension (include/wbx.h), in output frames
};
static_assert(sizeof(DSpan) == 112, "DSpan layout");

struct __align__(16) DCell {
  double pos;      // sample_offset_ when the call is made
  uint32_t span;   // index into spans, kSilent when nothing plays
  uint32_t n_act;  // num_actual_samples (dsp/sampler.cpp:104)
};
static_assert(sizeof(DCell) == 16, "DCell layout");

// item kinds after resolve (per cell and frame tile)
enum : uint32_t {
  K_SILENT = 0,
  K_FAST = 1,    // stereo f32, unity speed, whole tile, window starting on tile frame 0: packed f32x2 math, no index math
  K_GEN = 2,     // anything else whose window fits a stage: per-frame path on the staged window
  K_DIRECT = 3,  // window larger than a stage (speed well above 1): per-frame path straight from global
  K_UNI = 4,     // stereo f32, unity speed, odd start frame (whole tile): 64-bit loads + packed math
  K_LIN = 5,     // stereo f32, 2-tap linear resample from the staged window, conversion-free position split
  K_FADE = 6,        // a fade ramp overlaps this tile: per-frame path times the envelope (staged window)
  K_DIRECT_FADE = 7, // K_DIRECT with a fade ramp
  K_POLY = 8,        // stereo f32, polyphase windowed-sinc resample (extension) from the staged window
  K_UNI_P = 9,       // K_UNI / K_LIN on a partial tile (per-frame range checks); the plain kinds cover the whole tile
  K_LIN_P = 10
};

// Resolved per-(cell, tile) descriptor, lives in shared memory (64 B).
struct __align__(16) Desc {
  const void* src;   // staged kinds: 16-B aligned global address of the window; K_DIRECT: sample base
  double pos;        // segment position (f64) of segment-relative frame 0
  double speed;
  float gain;
  float tg[2];     // (mute ? 0 : volume) * pan_coeffs[c]
  uint32_t track;
  int32_t base;    // frame index (relative to the sample start) of window frame 0; 0 for K_DIRECT
  int32_t jrel0;   // segment-relative index of tile frame 0 (jj = frame_in_tile + jrel0)
  uint16_t lo, hi; // tile-relative frame range [lo, hi) this item covers
  uint16_t bytes;  // bytes to stage (multiple of 16)
  uint8_t kind;
  uint8_t fmt;     // wbx_format | 0x80 when the device copy has one channel | 0x40 polyphase quality mode
  uint32_t span;   // span index (K_FADE reads the envelope parameters from it)
  uint32_t block_in_run;  // callbacks since the run's first one (clip_frame advances by length per callback)
};
static_assert(sizeof(Desc) == 64, "Desc layout");

// effect chain of one track (extension, include/wbx.h): coefficients + running state
struct DFx {
  uint32_t track, eq_on, comp_on, ratio_code;
  float b0[4], b1[4], b2[4], a1[4], a2[4];
  float thr, att, rel, makeup;
  float s1[2][4], s2[2][4], env[2];  // state, persists across renders
  uint32_t reverb_on;
  uint32_t sel;  // att <= rel: the follower's branch is the larger candidate
  // time-parallel form (oracle/wb_oracle.c fx_design_tables): state-space input vector, A^m (m = 0..16) and A^(16 * 2^j)
  // (j = 0..4) per biquad, row-major 2x2; follower constants 1 - att, 1 - rel and look-ahead slopes S1 | S2 | S3 | S4
  float B1[4], B2[4], P[4][17][4], S[4][5][4];
  float a1m, r1m, sl[14];
};
static_assert(sizeof(DFx) == 16 + 80 + 16 + 72 + 8 + 32 + 1088 + 320 + 64, "DFx layout");

constexpr uint32_t kMaxPeers = 16;  // ranks of one sharded render (one NVSwitch box has 8)

struct MixParams {
  const DSpan* spans;
  const DCell* cells;
  const float* gains;   // [n_tracks][2]
  const float* poly;    // polyphase coefficient table [128][16] (extension)
  float* bus;           // [C][n_blocks*B]
  float* peaks;         // [n_blocks][n_tracks][2], pre-zeroed
  float* ws;            // tree mode: [tiles][groups][2][T] partial sums
  uint32_t* counters;   // [0] = dynamic work counter, [1 + (k*n_tiles+f)] = arrivals per output tile
  uint32_t n_tracks, n_blocks, slots;
  uint32_t B, C;
  uint32_t n_tiles;     // frame tiles per block
  uint32_t groups;      // track groups (1 = exact order)
  uint32_t tracks_per_group;
  uint32_t n_items;     // n_blocks * n_tiles * groups
  uint32_t clamp;       // apply the [-1, 1] clamp
  uint32_t ext;         // some segment carries an extension flag (fade / polyphase): use the full kernel build
  float one;            // 1.0f, opaque to ptxas (see consume_lin_t)
  // optional second destination of every bus tile: the caller's page-locked AudioBuffer channels, written from the
  // kernel over PCIe (posted stores) so no device-to-host copy follows the mix. nullptr = off.
  float* mirror[2];
  // sharded render (tracks split over ranks): callback k belongs to owner rank k / shard_blocks and this
  // rank's UNCLAMPED tile goes straight into the owner's exchange buffer (peer memory over NVLink),
  // xchg[owner] laid out [src_rank][C][shard_blocks * B]. shard_blocks == 0 = off (tile goes to `bus`).
  float* xchg[kMaxPeers];
  uint32_t shard_blocks, shard_rank;
};

// convolution reverb stage of launch_effects (extension, BASELINE cfg 5): the impulse response, the per-track time-domain
// history ring and what the chosen path needs on top of it
struct FirLaunch {
  const float* ir = nullptr;  // [L] taps on the device; L == 0: no reverb stage
  uint32_t L = 0;
  float* hist = nullptr;      // [n_tracks][2][L - 1] ring: logical index i (oldest first) lives at (hist_pos + i) mod (L - 1)
  uint64_t hist_pos = 0;
  float* xin = nullptr;       // direct / tensor-core paths: gather buffer [n_fx * C][L - 1 + T]
  int mode = 0;               // 0 direct form, 1 tensor cores, 2 partitioned FFT
  void* ir_aux = nullptr;     // 1: Toeplitz tiles of the response; 2: twiddles + partition spectra
  void* scratch = nullptr;    // 1: split-precision signal planes; 2: block spectra W
  void* fft_ring = nullptr;   // 2: window spectra Z[cap][n_fx][2P], a ring over the window index
  uint32_t fft_ring_base = 0, fft_ring_cap = 0;
  uint32_t fft_first_q = 0;   // 2: windows below this index are already in the ring (previous render)
  uint32_t fft_p = 0;         // 2: partition size (512 or 2048)
};

// ranks' views of one another for a sharded render (all pointers valid on this rank's device)
struct ShardPeers {
  uint32_t* flags[kMaxPeers];  // flags[j] = rank j's arrival words [kMaxPeers]
  float* dst[kMaxPeers];       // master-bus copies the reduced slices are written to (dst[0..n_dst))
  uint32_t n_dst;
  float* host_dst[2];          // optional: the host output channels (page-locked, mapped on this device) — the owner also
                               // stores its slice there, so no rank copies the whole bus to the host afterwards
};

}  // namespace wbx
