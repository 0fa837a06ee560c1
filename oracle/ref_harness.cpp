// TEST INFRASTRUCTURE ONLY (oracle build).
//
// Drives the reference's OWN wb::Engine (compiled unmodified from /root/reference/src) behind the wbo.h
// scenario API. Nothing in this file computes a sample: it only calls the reference's public editing API
// (engine.h:66-239, track.h:139-143) and copies the buffers Engine::process filled.
#include <chrono>
#include <cstring>
#include <string>
#include <vector>

#include "core/audio_format_conv.h"
#include "core/panning_law.h"
#include "engine/assets_table.h"
#include "engine/engine.h"
#include "engine/track.h"
#include "gfx/renderer.h"
#include "gfx/waveform_visual.h"
#include "wbo.h"
#ifdef WBO_GPU
#include "wbx_gpu_hooks.h"
#endif

using namespace wb;

namespace wb {
void* wbref_buffer_memory(GPUBuffer* b);  // ref_stubs.cpp: host memory behind a buffer of the stand-in renderer
void wbref_enable_null_plugin(bool on);   // ref_stubs.cpp: pm_open_plugin then returns a plugin that does nothing
}

struct wbo_session {
  Engine engine;
  std::vector<SampleAsset*> samples;
  uint32_t out_channels, block, rate;
  uint64_t uid;
  AudioBuffer<float> in;
  AudioBuffer<float> out;
  wbo_session(uint32_t c, uint32_t b) : out(b, c) {}
};

static uint64_t g_next_uid = 1;

extern "C" {

#ifndef WBO_KIND
#define WBO_KIND "reference"
#endif
const char* wbo_kind(void) { return WBO_KIND; }

wbo_session* wbo_create(uint32_t out_channels, uint32_t block_frames, uint32_t sample_rate, double bpm) {
  auto* s = new wbo_session(out_channels, block_frames);
  s->out_channels = out_channels;
  s->block = block_frames;
  s->rate = sample_rate;
  s->uid = g_next_uid++;
  s->engine.set_audio_channel_config(0, out_channels, block_frames, sample_rate);
  s->engine.set_bpm(bpm);
  return s;
}

void wbo_destroy(wbo_session* s) {
  if (!s) return;
  s->engine.clear_all();  // deletes tracks -> clips release their SampleAsset refs
#ifdef WBO_GPU
  wbx_gpu::release(&s->engine);
#endif
  delete s;
}

int wbo_add_track(wbo_session* s, float volume_db, float pan, int mute) {
  Track* t = s->engine.add_track("t" + std::to_string(s->engine.tracks.size()));
  t->set_volume(volume_db);
  t->set_pan(pan);
  t->set_mute(mute != 0);
  return (int)s->engine.tracks.size() - 1;
}

void wbo_set_bpm(wbo_session* s, double bpm) { s->engine.set_bpm(bpm); }
// Engine::add_plugin_to_track / delete_plugin_from_track (engine.cpp:1466-1551) with the no-op plugin of ref_stubs.cpp
int wbo_set_plugin(wbo_session* s, int track, int present) {
  Track* t = s->engine.tracks[track];
  if (present) {
    wbref_enable_null_plugin(true);
    PluginUID uid = {};
    PluginInterface* p = s->engine.add_plugin_to_track(t, uid);
    wbref_enable_null_plugin(false);
    return p ? 0 : -1;
  }
  s->engine.delete_plugin_from_track(t);
  return 0;
}
// Engine::set_audio_channel_config again (config.cpp:198-232: the audio device changed under a running session)
int wbo_reconfigure(wbo_session* s, uint32_t out_channels, uint32_t block, uint32_t rate) {
  s->engine.set_audio_channel_config(0, out_channels, block, rate);
  s->out_channels = out_channels;
  s->block = block;
  s->rate = rate;
  s->out.resize(block);
  s->out.resize_channel(out_channels);
  return 0;
}
void wbo_set_volume(wbo_session* s, int track, float db) { s->engine.tracks[track]->set_volume(db); }
void wbo_set_pan(wbo_session* s, int track, float pan) { s->engine.tracks[track]->set_pan(pan); }
void wbo_set_mute(wbo_session* s, int track, int mute) { s->engine.tracks[track]->set_mute(mute != 0); }

int wbo_add_sample(wbo_session* s, int format, uint32_t channels, uint64_t frames, uint32_t sample_rate,
                   const void* const* planar) {
  Sample smp((AudioFormat)format, sample_rate);
  // The asset table keys on XXH64(path) (assets_table.cpp:25-26): unique path per sample and session.
  smp.name = "s" + std::to_string(s->uid) + "_" + std::to_string(s->samples.size());
  smp.path = smp.name;
  smp.resize(frames, channels);
  size_t elem = (AudioFormat)format == AudioFormat::I24 ? 4 : get_audio_format_size((AudioFormat)format);
  size_t bytes = frames * elem;  // I24 = int32 container, see ref_stubs.cpp Sample::resize
  for (uint32_t c = 0; c < channels; c++)
    std::memcpy(smp.sample_data[c], planar[c], bytes);
  SampleAsset* asset = g_sample_table.create_from_existing_sample(std::move(smp));
  if (!asset) return -1;
  asset->keep_alive = false;
  s->samples.push_back(asset);
  return (int)s->samples.size() - 1;
}

int wbo_add_clip(wbo_session* s, int track, int sample, double min_beat, double max_beat, double start_offset,
                 double speed, float gain) {
  SampleAsset* asset = s->samples[sample];
  asset->add_ref();  // the clip owns one reference (released by ~Clip, clip.h:129-141)
  s->engine.add_audio_clip(
      s->engine.tracks[track], "c", min_beat, max_beat, start_offset,
      AudioClip{ .asset = asset, .fade_start = 0.0, .fade_end = 0.0, .speed = speed, .gain = gain });
  return 0;
}

int wbo_add_clip_fade(wbo_session* s, int track, int sample, double min_beat, double max_beat, double start_offset,
                      double speed, float gain, double fade_start, double fade_end) {
  SampleAsset* asset = s->samples[sample];
  asset->add_ref();
  // the fields are stored (clip.h:41-42) — and ignored by every audio path of the reference
  s->engine.add_audio_clip(
      s->engine.tracks[track], "c", min_beat, max_beat, start_offset,
      AudioClip{ .asset = asset, .fade_start = fade_start, .fade_end = fade_end, .speed = speed, .gain = gain });
  return 0;
}

int wbo_clip_count(wbo_session* s, int track) { return (int)s->engine.tracks[track]->clips.size(); }

static Clip* clip_at(wbo_session* s, int track, int clip) {
  auto& clips = s->engine.tracks[track]->clips;
  return (clip < 0 || (size_t)clip >= clips.size()) ? nullptr : clips[clip];
}

int wbo_clip_range(wbo_session* s, int track, int clip, double* min_beat, double* max_beat) {
  Clip* c = clip_at(s, track, clip);
  if (!c) return -1;
  *min_beat = c->min_time;
  *max_beat = c->max_time;
  return 0;
}

int wbo_move_clip(wbo_session* s, int track, int clip, double relative_pos) {
  Clip* c = clip_at(s, track, clip);
  if (!c) return -1;
  s->engine.move_clip(s->engine.tracks[track], c, relative_pos);
  return 0;
}

int wbo_resize_clip(wbo_session* s, int track, int clip, double relative_pos, double resize_limit, double min_length,
                    int left_side, int shift, int stretch) {
  Clip* c = clip_at(s, track, clip);
  if (!c) return -1;
  s->engine.resize_clip(s->engine.tracks[track], c, relative_pos, resize_limit, min_length, left_side != 0, shift != 0,
                        stretch != 0);
  return 0;
}

int wbo_delete_clip(wbo_session* s, int track, int clip) {
  Clip* c = clip_at(s, track, clip);
  if (!c) return -1;
  s->engine.delete_clip(s->engine.tracks[track], c);
  return 0;
}

int wbo_duplicate_clip(wbo_session* s, int track, int clip, double min_beat, double max_beat) {
  Clip* c = clip_at(s, track, clip);
  if (!c) return -1;
  s->engine.duplicate_clip(s->engine.tracks[track], c, min_beat, max_beat);
  return 0;
}

int wbo_delete_region(wbo_session* s, int track, double min_beat, double max_beat) {
  // Engine::delete_region trims through the GLOBAL g_engine (engine.cpp:467-470); in the application that is the only
  // engine. Mirror this session's tempo and playhead there so the call sees what the application's engine would hold.
  g_engine.beat_duration.store(s->engine.beat_duration.load(std::memory_order_relaxed), std::memory_order_relaxed);
  g_engine.playhead = s->engine.playhead;
  s->engine.delete_region(s->engine.tracks[track], min_beat, max_beat);
  return 0;
}

int wbo_set_clip_gain(wbo_session* s, int track, int clip, float gain) {
  if (!clip_at(s, track, clip)) return -1;
  s->engine.set_clip_gain(s->engine.tracks[track], (uint32_t)clip, gain);
  return 0;
}
void wbo_solo_track(wbo_session* s, int track) { s->engine.solo_track((uint32_t)track); }
void wbo_move_track(wbo_session* s, int from_slot, int to_slot) { s->engine.move_track((uint32_t)from_slot, (uint32_t)to_slot); }
void wbo_delete_track(wbo_session* s, int track) { s->engine.delete_track((uint32_t)track); }

int wbo_set_effects(wbo_session*, int, const wbo_effects*) { return -1; }  // the reference has no effects
int wbo_set_impulse_response(wbo_session*, const float*, uint32_t) { return -1; }
void wbo_set_resampler(wbo_session*, int) {}  // the reference has only the linear resampler

void wbo_set_playhead(wbo_session* s, double beat) { s->engine.set_playhead_position(beat); }
void wbo_play(wbo_session* s) { s->engine.play(); }
void wbo_stop(wbo_session* s) { s->engine.stop(); }

int wbo_process(wbo_session* s, uint32_t n_blocks, float* out, float* peaks) {
  const uint32_t C = s->out_channels, B = s->block;
  const size_t n_tracks = s->engine.tracks.size();
  for (uint32_t k = 0; k < n_blocks; k++) {
    s->engine.process(s->in, s->out, (double)s->rate);
    if (out)
      for (uint32_t c = 0; c < C; c++)
        std::memcpy(out + ((size_t)k * C + c) * B, s->out.channel_buffers[c], B * sizeof(float));
    for (size_t t = 0; t < n_tracks; t++)
      for (uint32_t c = 0; c < 2; c++) {
        // level only ever rises inside process (vu_meter.h:26-29); reading-and-zeroing it after each
        // callback (what VUMeter::update does, vu_meter.h:33) yields that callback's block peak.
        float v = s->engine.tracks[t]->level_meter[c].level.exchange(0.0f);
        if (peaks) peaks[((size_t)k * n_tracks + t) * 2 + c] = v;
      }
  }
  return 0;
}

double wbo_time_process(wbo_session* s, uint32_t n_blocks) {
  auto t0 = std::chrono::steady_clock::now();
  for (uint32_t k = 0; k < n_blocks; k++)
    s->engine.process(s->in, s->out, (double)s->rate);
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

int wbo_mipmap(wbo_session* s, int sample, int quality, int level, void* out, uint64_t cap_elems, uint32_t* count) {
  Sample* smp = &s->samples[sample]->sample_instance;
  WaveformVisual* v = WaveformVisual::create(smp, quality ? WaveformVisualQuality::High : WaveformVisualQuality::Low);
  if (!v) return -1;
  const int n_levels = (int)v->mipmaps.size();
  if (level >= 0 && level < n_levels) {
    const WaveformMipmap& m = v->mipmaps[level];
    const uint64_t elems = (uint64_t)m.count * smp->channels;
    if (count) *count = m.count;
    if (out && elems <= cap_elems) std::memcpy(out, wbref_buffer_memory(m.data), elems * (quality ? 2 : 1));
  }
  delete v;
  return n_levels;
}

double wbo_sampler_offset(wbo_session* s, int track) { return s->engine.tracks[track]->sampler.sample_offset_; }
double wbo_sample_position(wbo_session* s) { return s->engine.sample_position; }
double wbo_playhead(wbo_session* s) { return s->engine.playhead; }

void wbo_panning_coefs(float pan, float* left, float* right) {
  PanningCoefficient c = calculate_panning_coefs(pan, PanningLaw::ConstantPower_3db);
  *left = c.left;
  *right = c.right;
}

float wbo_db_to_linear(float db) { return math::db_to_linear(db); }

void wbo_interleave(void* dst, const float* const* src, uint32_t offset, uint32_t frames, uint32_t channels,
                    int fmt) {
  switch ((AudioFormat)fmt) {
    case AudioFormat::I16: convert_f32_to_interleaved_i16((int16_t*)dst, src, offset, frames, channels); break;
    case AudioFormat::I24: convert_f32_to_interleaved_i24((std::byte*)dst, src, offset, frames, channels); break;
    case AudioFormat::I24_X8: convert_f32_to_interleaved_i24_x8((int32_t*)dst, src, offset, frames, channels); break;
    case AudioFormat::I32: convert_f32_to_interleaved_i32((int32_t*)dst, src, offset, frames, channels); break;
    case AudioFormat::F32: convert_to_interleaved_f32((float*)dst, src, offset, frames, channels); break;
    default: break;
  }
}

}  // extern "C"
