// wbx_device.cuh — device-side data model of the mixing hot path (internal; the ABI is include/wbx.h).
//
// HBM layout
//   samples   one allocation per Sample (dsp/sample.h:18-28): channels x stride elements, planar, channel
//             base 256-B aligned, stride a multiple of 32 elements, >= 16 zero frames after `frames`
//             (dsp/sample.cpp:127,140) plus slack so 16-B aligned bulk windows never leave the allocation.
//   spans     one DSpan per wbx_segment (L2-resident; reused by every block of a run).
//   cells     [n_blocks][n_tracks][slots] DCell, 16 B each: the state of one Sampler::stream call
//             (position in f64, clipped length) written by the schedule-expansion kernel. 0.4 % of the
//             audio bytes it describes.
//   bus       [out_channels][n_blocks * block_frames] f32 (AudioBuffer layout, core/audio_buffer.h:19-23)
//   peaks     [n_blocks][n_tracks][2] f32 (VUMeter block peaks, engine/vu_meter.h:20-30)
#pragma once
#include <cstdint>

namespace wbx {

constexpr uint32_t kSilent = 0xFFFFFFFFu;

struct __align__(16) DSpan {
  const void* ch[2];  // source channel base for output channel 0 / 1 (c % sample_channels resolved)
  double pos0;        // Sampler::sample_offset_ at the first call
  double speed;       // Sampler::playback_speed_
  uint64_t count;     // Sample::count
  float gain;         // AudioClip::gain
  uint32_t track;
  uint32_t block0, n_blocks;
  uint32_t dst_off, length;
  uint32_t fmt;   // wbx_format
  uint32_t slot;  // which of the `slots` cells of (block, track) this span writes
  uint32_t mono;  // ch[1] == ch[0]
  uint32_t pad;
};
static_assert(sizeof(DSpan) == 80, "DSpan layout");

struct __align__(16) DCell {
  double pos;      // sample_offset_ when the call is made
  uint32_t span;   // index into spans, kSilent when nothing plays
  uint32_t n_act;  // num_actual_samples (dsp/sampler.cpp:104)
};
static_assert(sizeof(DCell) == 16, "DCell layout");

// item kinds after resolve (per cell and frame tile)
enum : uint32_t { K_SILENT = 0, K_VEC = 1, K_UNI = 2, K_GEN = 3, K_DIRECT = 4 };

// Resolved per-(cell, tile) descriptor, lives in shared memory (64 B).
struct __align__(16) Desc {
  const void* src[2];  // staged kinds: 16-B aligned global address of the window; K_DIRECT: channel base
  double pos;          // segment position (f64) of segment-relative frame 0
  double speed;
  float gain;
  float tg[2];     // (mute ? 0 : volume) * pan_coeffs[c]
  uint32_t track;
  int32_t base;    // element index (relative to the channel base) of window element 0; 0 for K_DIRECT
  int32_t jrel0;   // segment-relative index of tile frame 0 (jj = frame_in_tile + jrel0)
  uint16_t lo, hi; // tile-relative frame range [lo, hi) this item covers
  uint16_t bytes;  // bytes per channel to stage (multiple of 16)
  uint8_t kind;
  uint8_t fmt;     // wbx_format | 0x80 when both output channels read the same source channel (mono)
};
static_assert(sizeof(Desc) == 64, "Desc layout");

struct MixParams {
  const DSpan* spans;
  const DCell* cells;
  const float* gains;   // [n_tracks][2]
  float* bus;           // [C][n_blocks*B]
  float* peaks;         // [n_blocks][n_tracks][2], pre-zeroed
  float* ws;            // tree mode: [items][C][tile] partial sums
  uint32_t* counters;   // [0] = dynamic work counter, [1 + (k*n_tiles+f)] = arrivals per output tile
  uint32_t n_tracks, n_blocks, slots;
  uint32_t B, C;
  uint32_t n_tiles;     // frame tiles per block
  uint32_t groups;      // track groups (1 = exact order)
  uint32_t tracks_per_group;
  uint32_t n_items;     // n_blocks * n_tiles * groups
  uint32_t clamp;       // apply the [-1, 1] clamp
};

}  // namespace wbx
