/* TEST INFRASTRUCTURE ONLY — never linked into, imported by or executed from the product path.
 *
 * wbo.h — the scenario API shared by the two CPU checkers of the whitebox mixing hot path:
 *
 *   oracle/_ref/libwbref.so  the reference's OWN Engine::process, compiled unmodified from
 *                            /root/reference/src (engine/engine.cpp:1576-1654, engine/track.cpp:587-736,
 *                            dsp/sampler.cpp:88-210 ...) by oracle/Makefile, driven by ref_harness.cpp.
 *   oracle/liboracle.so      a plain-C restatement of the same algorithm (wb_oracle.c), pinned against
 *                            libwbref.so and against tests/golden/ vectors produced by libwbref.so.
 *
 * Both export exactly these symbols so one ctypes wrapper (tests/oracle_api.py) drives either.
 * Calls mirror the reference's editing API (Engine::add_track / add_audio_clip / play / process,
 * Track::set_volume / set_pan / set_mute) so a scenario reads like a reference session.
 */
#ifndef WBO_H
#define WBO_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* wb::AudioFormat values (src/core/audio_format.h:7-20) used by Sample (src/dsp/sample.h:18-28). */
enum { WBO_FMT_I16 = 3, WBO_FMT_I24 = 5, WBO_FMT_I32 = 7, WBO_FMT_F32 = 9 };

typedef struct wbo_session wbo_session;

/* Engine::set_audio_channel_config(0, out_channels, block, rate) + set_bpm(bpm). */
wbo_session* wbo_create(uint32_t out_channels, uint32_t block_frames, uint32_t sample_rate, double bpm);
void wbo_destroy(wbo_session*);
const char* wbo_kind(void); /* "reference" or "port" */

/* Engine::set_bpm (engine.cpp:24-30): takes effect from the next callback's transport math (engine.cpp:1578-1585). */
void wbo_set_bpm(wbo_session*, double bpm);

/* Engine::add_track + Track::set_volume/set_pan/set_mute. Returns the track index. */
int wbo_add_track(wbo_session*, float volume_db, float pan, int mute);
void wbo_set_volume(wbo_session*, int track, float db);
void wbo_set_pan(wbo_session*, int track, float pan);
void wbo_set_mute(wbo_session*, int track, int mute);

/* A resident Sample (planar channels, `frames` elements each + 16 zero frames of padding,
 * sample.cpp:127,140). I24 data is passed already widened to int32 (sample.cpp:20). Returns id. */
int wbo_add_sample(wbo_session*, int format, uint32_t channels, uint64_t frames, uint32_t sample_rate,
                   const void* const* planar);

/* Engine::add_audio_clip(track, name, min_beat, max_beat, start_offset_frames, {asset, speed, gain}). */
int wbo_add_clip(wbo_session*, int track, int sample, double min_beat, double max_beat, double start_offset,
                 double speed, float gain);

/* Clip editing (audio clips). `clip` = index in the track's clip list (ordered by min_time) at the time of the call:
 * Engine::move_clip (engine.cpp:346-363), resize_clip (:365-398, calc_resize_clip clip_edit.h:18-126), delete_clip
 * (:400-407), duplicate_clip (:336-344). Return 0, or -1 for a bad index. */
int wbo_clip_count(wbo_session*, int track);
int wbo_clip_range(wbo_session*, int track, int clip, double* min_beat, double* max_beat);
int wbo_move_clip(wbo_session*, int track, int clip, double relative_pos);
int wbo_resize_clip(wbo_session*, int track, int clip, double relative_pos, double resize_limit, double min_length,
                    int left_side, int shift, int stretch);
int wbo_delete_clip(wbo_session*, int track, int clip);
int wbo_duplicate_clip(wbo_session*, int track, int clip, double min_beat, double max_beat);
/* Engine::delete_region(track, min, max) (engine.cpp:463-473): erase a time range — clips inside it are trimmed / split / deleted */
int wbo_delete_region(wbo_session*, int track, double min_beat, double max_beat);

/* Mixer-side calls that change what the path sums: Engine::set_clip_gain (engine.cpp:1460-1464: the playing voice reads
 * the clip gain every callback, track.cpp:676,716), solo_track (:245-262, through set_mute), move_track (:228-243: the bus
 * summation order), delete_track (:209-217). */
int wbo_set_clip_gain(wbo_session*, int track, int clip, float gain);
void wbo_solo_track(wbo_session*, int track);
void wbo_move_track(wbo_session*, int from_slot, int to_slot);
void wbo_delete_track(wbo_session*, int track);

/* As wbo_add_clip, also setting AudioClip::fade_start / fade_end (beats, engine/clip.h:41-42).
 * EXTENSION — PARITY UNPINNED w.r.t. whitebox: the reference stores these fields but no audio code reads them,
 * so libwbref.so renders such a clip WITHOUT a fade; the port implements the builder's specification
 * (wb_oracle.c fade_env): for clip-relative output frame n,
 *   env(n) = (float)(min(1, n / Fin) * min(1, max(0, (L - n) / Fout))),  a factor is 1 when its length <= 0,
 *   Fin/Fout/L = beat_to_samples(fade_start / fade_end / max_time - min_time); frame value (src * gain) * env. */
int wbo_add_clip_fade(wbo_session*, int track, int sample, double min_beat, double max_beat, double start_offset,
                      double speed, float gain, double fade_start, double fade_end);

/* Per-track effect chain — EXTENSION, PARITY UNPINNED w.r.t. whitebox (the reference has no DSP effects; its
 * only per-track effect slot is a third-party VST3 instance, engine/track.h:124). Builder's specification
 * (wb_oracle.c apply_effects), placed where a native PluginInterface::process would sit: on the track's mixing
 * buffer after the clips are rendered and before dsp::apply_gain / the VU meter (engine/track.cpp:726-733).
 *   4-band EQ: RBJ biquads (0 low shelf, 1/2 peaking, 3 high shelf), coefficients designed in f64, stored f32,
 *              transposed direct form II in f32 with fused multiply-adds, per channel.
 *   compressor: per channel peak follower (attack/release one-poles), hard knee, ratio 2/4/8/limiter evaluated
 *              with IEEE divide + square roots only (bit-reproducible), then make-up gain.
 * eq_gain_db[b] == 0 for all b and ratio_code == 0 disables the respective stage. */
typedef struct wbo_effects {
  float eq_freq[4], eq_gain_db[4], eq_q[4];
  float comp_threshold_db, comp_attack_ms, comp_release_ms, comp_makeup_db;
  int comp_ratio_code; /* 0 off, 1 = 2:1, 2 = 4:1, 3 = 8:1, 4 = limiter */
  int reverb_on;       /* 1: convolve the chain's output with the session's impulse response (BASELINE cfg 5) */
} wbo_effects;
int wbo_set_effects(wbo_session*, int track, const wbo_effects* fx); /* port only; libwbref.so returns -1 */
/* A plugin in the track's slot (Engine::add_plugin_to_track / delete_plugin_from_track, engine/engine.cpp:1466-1551).
 * libwbref.so attaches a plugin that does nothing: what is pinned is the effect of a plugin's PRESENCE on the path —
 * Track::process then renders the clips into effect_buffer and never mixes them (track.cpp:600,645-724). */
int wbo_set_plugin(wbo_session*, int track, int present);
/* Engine::set_audio_channel_config again on a live session (config.cpp:198-232: device change / removal). */
int wbo_reconfigure(wbo_session*, uint32_t out_channels, uint32_t block_frames, uint32_t sample_rate);
/* port only: 0 = the chain's specification (time-parallel association), 1 = textbook f32, 2 = textbook f64 */
void wbo_set_fx_textbook(int mode);
/* Convolution reverb — EXTENSION, PARITY UNPINNED: one impulse response h[0..n_taps) per session, applied as the
 * last stage of a track's chain: y[n] = (float) sum_k (double)h[k] * (double)x[n - k], k ascending, accumulated in
 * f64 (the ground truth a tensor-core or f32 implementation is held to within 1e-5 of the block peak); x before
 * the chain was attached is 0 and the history persists across callbacks. */
int wbo_set_impulse_response(wbo_session*, const float* h, uint32_t n_taps); /* port only */

/* Resampler quality — EXTENSION, PARITY UNPINNED: the reference only has the 2-tap linear resampler
 * (dsp/sampler.h:8-11; Track::process hard-codes ResamplerType::Linear, track.cpp:692-697). mode 1 selects the
 * builder's polyphase windowed-sinc resampler (wb_oracle.c sample_polyphase: 128 phases x 16 taps, Blackman window,
 * f32 fused multiply-add accumulation in tap order) for stereo f32 sources on a stereo bus at speed != 1; every
 * other case keeps the reference's linear path. mode 0 (default) is the reference path. */
void wbo_set_resampler(wbo_session*, int mode);

void wbo_set_playhead(wbo_session*, double beat); /* Engine::set_playhead_position */
void wbo_play(wbo_session*);                      /* Engine::play */
void wbo_stop(wbo_session*);                      /* Engine::stop */

/* n_blocks consecutive Engine::process callbacks.
 *   out   [n_blocks][out_channels][block_frames] f32 — the clamped master bus of each callback
 *   peaks [n_blocks][n_tracks][2] f32 — VUMeter block peak of each track/channel, i.e. the value
 *         push_samples (vu_meter.h:20-30) offers to `level` for that callback (may be NULL). */
int wbo_process(wbo_session*, uint32_t n_blocks, float* out, float* peaks);

double wbo_sampler_offset(wbo_session*, int track); /* Track::sampler.sample_offset_ */
double wbo_sample_position(wbo_session*);           /* Engine::sample_position */
double wbo_playhead(wbo_session*);                  /* Engine::playhead */

/* Host-side scalar math feeding the kernel (core/panning_law.cpp:9-32, core/core_math.h:83-89). */
void wbo_panning_coefs(float pan, float* left, float* right);
float wbo_db_to_linear(float db);

/* planar f32 -> interleaved device format (core/audio_format_conv.cpp:5-106). fmt: WBO_FMT_I16,
 * WBO_FMT_I24 (packed 3 bytes), 6 = I24_X8, WBO_FMT_I32, WBO_FMT_F32. */
void wbo_interleave(void* dst, const float* const* src, uint32_t offset, uint32_t frames, uint32_t channels,
                    int fmt);

/* Waveform peak mip-maps of a resident sample (gfx/waveform_visual.cpp:9-173 summarize_for_mipmaps_impl,
 * :181-248 WaveformVisual::create). quality 0 = Low (int8), 1 = High (int16). Level l uses chunks of
 * 2^(2l+1) frames; its data is [channels][count] elements, each chunk giving a (first, second) pair = (min, max)
 * ordered by which occurs first. Copies level `level` into out (capacity cap_elems elements) and returns the
 * number of levels, or a negative value; *count = elements per channel of that level. */
int wbo_mipmap(wbo_session*, int sample, int quality, int level, void* out, uint64_t cap_elems, uint32_t* count);

/* CPU timing of the same loop (bench.py cpu_baseline / --impl reference only): runs n_blocks callbacks
 * without copying results out and returns elapsed seconds (steady clock). */
double wbo_time_process(wbo_session*, uint32_t n_blocks);

#ifdef __cplusplus
}
#endif
#endif
