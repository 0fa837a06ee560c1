/*
 * wbx_host.h — C exports of the host-side engine (include/wbx_engine.hpp) for FFI callers (ctypes, cgo-style
 * bindings). Same call shape as the reference's editing / transport API:
 *   Engine::add_track (engine/engine.cpp:199), Track::set_volume/set_pan/set_mute (engine/track.cpp:47-79),
 *   Engine::add_audio_clip (engine.cpp:293), Engine::set_playhead_position (:32), play (:68), stop (:82),
 *   Engine::process (:1576) — here wbxh_render(n_blocks = 1) — and the batched n_blocks > 1 form.
 * All sample work runs in the CUDA engine (wbx.h); there is no CPU render path.
 * Threads: one audio thread (wbxh_render*, wbxh_schedule) and one UI thread (everything else) may run concurrently —
 * edits and renders serialise on the engine's editor lock, parameter setters and wbxh_level are lock-free
 * (include/wbx_engine.hpp, "Threading contract").
 */
#ifndef WBX_HOST_H
#define WBX_HOST_H
#include <stdint.h>

#include "wbx.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wbxh_engine wbxh_engine;

/* device_ordinal < 0: scheduling-only engine (wbxh_schedule works, wbxh_render returns WBX_ERR_NO_DEVICE). */
int wbxh_create(wbxh_engine** out, int device_ordinal, uint32_t out_channels, uint32_t block_frames,
                uint32_t sample_rate, double bpm);
void wbxh_destroy(wbxh_engine* h);
const char* wbxh_last_error(wbxh_engine* h);
wbx_engine* wbxh_device(wbxh_engine* h); /* the underlying device engine (wbx.h) */

int wbxh_add_track(wbxh_engine* h, float volume_db, float pan, int mute); /* returns the track index */
/* Track::set_volume / set_pan / set_mute (engine/track.cpp:47-79): lock-free messages to the audio thread, may be called
 * from the UI thread while another thread renders; WBX_ERR_INVALID for a bad track index. */
int wbxh_set_volume(wbxh_engine* h, int track, float db);
int wbxh_set_pan(wbxh_engine* h, int track, float pan);
int wbxh_set_mute(wbxh_engine* h, int track, int mute);
/* A plugin in the track's slot (engine/track.h:124): the reference then renders the clips into the plugin's effect buffer
 * and never mixes them (track.cpp:600,645-724), so the track contributes only the plugin's own output — silence here. */
int wbxh_set_plugin(wbxh_engine* h, int track, int present);
/* Engine::set_audio_channel_config again (config.cpp:198-232: device change / removal): new block size / rate / channel
 * count; tracks, clips, resident samples and the transport persist. */
int wbxh_configure(wbxh_engine* h, uint32_t out_channels, uint32_t block_frames, uint32_t sample_rate);
/* Engine load (core/timing.h:54-67): EMA of render wall time / rendered audio time, clamped to [0, 1]. */
double wbxh_cpu_usage(wbxh_engine* h);
/* returns the sample id (>= 0) or a negative wbx_status */
int wbxh_add_sample(wbxh_engine* h, int format, uint32_t channels, uint64_t frames, uint32_t sample_rate,
                    const void* const* planar);
int wbxh_add_clip(wbxh_engine* h, int track, int sample, double min_beat, double max_beat, double start_offset,
                  double speed, float gain);
/* as wbxh_add_clip plus AudioClip::fade_start / fade_end in beats (engine/clip.h:41-42) — the fade EXTENSION
 * specified in wbx.h; (0, 0) is the reference path. */
int wbxh_add_clip_fade(wbxh_engine* h, int track, int sample, double min_beat, double max_beat, double start_offset,
                       double speed, float gain, double fade_start, double fade_end);
/* Clip editing, `clip` = index in the track's clip list (ordered by min_beat) at the time of the call: Engine::move_clip
 * (engine/engine.cpp:346), resize_clip (:365), delete_clip (:400), duplicate_clip (:336). 0 or a negative wbx_status. */
/* Engine::delete_track (engine/engine.cpp:209), move_track (:228), solo_track (:245), set_clip_gain (:1460) */
int wbxh_delete_track(wbxh_engine* h, int track);
int wbxh_move_track(wbxh_engine* h, int from_slot, int to_slot);
int wbxh_solo_track(wbxh_engine* h, int track);
int wbxh_set_clip_gain(wbxh_engine* h, int track, int clip, float gain);
int wbxh_clip_count(wbxh_engine* h, int track);
int wbxh_clip_range(wbxh_engine* h, int track, int clip, double* min_beat, double* max_beat);
int wbxh_move_clip(wbxh_engine* h, int track, int clip, double relative_pos);
int wbxh_resize_clip(wbxh_engine* h, int track, int clip, double relative_pos, double resize_limit, double min_length,
                     int left_side, int shift, int stretch);
int wbxh_delete_clip(wbxh_engine* h, int track, int clip);
int wbxh_duplicate_clip(wbxh_engine* h, int track, int clip, double min_beat, double max_beat);
int wbxh_delete_region(wbxh_engine* h, int track, double min_beat, double max_beat); /* Engine::delete_region (:463) */
/* attach (params != NULL) or remove the built-in EQ + compressor chain of a track (extension, see wbx.h) */
int wbxh_set_effects(wbxh_engine* h, int track, const wbx_effect_params* params);
int wbxh_set_impulse_response(wbxh_engine* h, const float* ir, uint32_t n_taps); /* convolution reverb IR (wbx.h) */
void wbxh_set_resampler(wbxh_engine* h, int mode); /* 0 linear (reference), 1 polyphase (extension, wbx.h) */
void wbxh_set_bpm(wbxh_engine* h, double bpm); /* Engine::set_bpm (engine/engine.cpp:24-30) */
void wbxh_set_playhead(wbxh_engine* h, double beat);
void wbxh_play(wbxh_engine* h);
void wbxh_stop(wbxh_engine* h);
void wbxh_set_fast_forward(wbxh_engine* h, int on);

/* n_blocks consecutive Engine::process callbacks: out_channels[c] -> n_blocks*block_frames f32 (clamped bus),
 * peaks (optional) [n_blocks][n_tracks][2]. */
int wbxh_render(wbxh_engine* h, uint32_t n_blocks, float* const* out_channels, float* peaks);
/* Offline bounce / export of [start_beat, end_beat) from a stopped transport (wbx::Engine::bounce, wbx_engine.hpp): the
 * clamped bus in dst_format (WBX_FMT_I16 / I24_X8 / I32 / F32, interleaved as core/audio_format_conv.cpp writes them),
 * rendered in chunks of chunk_blocks callbacks (0 = 256) whose copy-out overlaps the next chunk's mix. wbxh_bounce fills
 * dst (cap_bytes; too small = WBX_ERR_INVALID); wbxh_bounce_wav writes a RIFF/WAVE file (I24_X8 becomes 24-bit PCM).
 * *frames_out = frames delivered. */
int wbxh_bounce(wbxh_engine* h, double start_beat, double end_beat, int dst_format, uint32_t chunk_blocks, void* dst,
                uint64_t cap_bytes, uint64_t* frames_out);
int wbxh_bounce_wav(wbxh_engine* h, double start_beat, double end_beat, int dst_format, uint32_t chunk_blocks, const char* path,
                    uint64_t* frames_out);
/* wbxh_render in two halves for one thread driving several engines of a sharded setup (wbx.h "sharded render"):
 * begin = host schedule + wbx_submit; then wbx_mix_sharded_phase(wbxh_device(h), 0..2) in lock step over all engines;
 * end = bus (rank 0 only, others pass NULL) / peaks / levels back. */
int wbxh_render_begin(wbxh_engine* h, uint32_t n_blocks);
int wbxh_render_end(wbxh_engine* h, float* const* out_channels, float* peaks);
/* host scheduling only: builds the wbx_segment table + track gains for n_blocks callbacks and advances the
 * transport; pointers stay valid until the next call on this engine. */
int wbxh_schedule(wbxh_engine* h, uint32_t n_blocks, const wbx_segment** segs, uint32_t* n_segs,
                  const float** gains);

double wbxh_sampler_offset(wbxh_engine* h, int track); /* Track::sampler.sample_offset_ */
double wbxh_sample_position(wbxh_engine* h);           /* Engine::sample_position */
double wbxh_playhead(wbxh_engine* h);                  /* Engine::playhead */
float wbxh_level(wbxh_engine* h, int track, int channel, int reset); /* VUMeter::level (+ exchange(0)) */
/* wbx::advance_rounded (wbx_engine.hpp): the sampler position recurrence over n callbacks, exact, closed form per binade */
uint32_t wbxh_advance_rounded(double* off, double adv, uint32_t n, double limit);
void wbxh_panning_coefs(float pan, float* left, float* right);
float wbxh_db_to_linear(float db);

#ifdef __cplusplus
}
#endif
#endif
