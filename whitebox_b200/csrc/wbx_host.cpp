// wbx_host.cpp — host side of the mixing path (see include/wbx_engine.hpp): transport, clip scheduling and
// parameter bookkeeping in doubles, mirroring the reference's audio-thread host logic so that the segment
// table handed to the device describes exactly the Sampler::stream calls the reference would make.
// No sample value is read or written here.
//
// Built with -ffp-contract=off: every double operation below must round like the reference's x86-64 build,
// because event offsets come from truncating those doubles (engine/track.cpp:359-361,378-379,423-425).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "../../include/wbx_engine.hpp"
#include "../../include/wbx_host.h"
#include "wbx_device.cuh"

namespace wbx {

// ---- scalar math ------------------------------------------------------------------------------------------

// core/core_math.h:209-212
static inline double beat_to_samples(double beat, double sample_rate, double beat_duration) {
  double sec = beat * beat_duration;
  return sec * sample_rate;
}

// core/core_math.h:83-89
float db_to_linear(float db) {
  if (db <= -72.0f) return 0.0f;
  return std::pow(10.0f, (float)((double)db * 0.05));
}

// core/panning_law.cpp:9-32, PanningLaw::ConstantPower_3db (the law Track::process hard-codes, track.cpp:627)
PanningCoefficient calculate_panning_coefs(float p) {
  constexpr double pi = 3.141592653589793238462643383279502884;
  const double x = 0.5 * ((double)p + 1.0);
  const double left = std::sin(0.5 * pi * (1.0 - x));
  const double right = std::sin(0.5 * pi * x);
  const double boost = std::sqrt(2.0);
  return {(float)(left * boost), (float)(right * boost)};
}

// ---- Track ------------------------------------------------------------------------------------------------

enum : uint32_t { kParamVolume = 0, kParamPan = 1, kParamMute = 2 };  // TrackParameter, engine/track.h:29-34

Track::~Track() {
  for (auto* c : graveyard) delete c;
  for (auto* c : clips) delete c;
}
void Track::set_volume(float db) {
  ui_parameter_state.volume_db = db;
  ui_parameter_state.volume = db_to_linear(db);
  track_msg_queue.push({kParamVolume, (double)ui_parameter_state.volume});
}
void Track::set_pan(float pan) {
  ui_parameter_state.pan = pan;
  track_msg_queue.push({kParamPan, (double)pan});
}
void Track::set_effects(const wbx_effect_params* params) {
  effects_on = params != nullptr;
  if (params) effect_params = *params;
  effects_dirty = true;
}
void Track::set_mute(bool mute) {
  ui_parameter_state.mute = mute;
  track_msg_queue.push({kParamMute, (double)mute});
}

// ---- clip lookup (Track::find_next_clip, engine/track.cpp:182-213, incl. find_lower_bound's n-1 start) ----

static bool find_next_clip(const Track& t, double time_pos, uint32_t* idx) {
  const size_t n = t.clips.size();
  if (n == 0) return false;
  if (t.clips.back()->max_time < time_pos) return false;
  int64_t left = 0, right = (int64_t)n - 1;
  while (left < right) {
    const int64_t middle = (left + right) >> 1;
    if (t.clips[middle]->max_time <= time_pos)
      left = middle + 1;
    else
      right = middle;
  }
  *idx = (uint32_t)right;
  return true;
}

// Track::reset_playback_state, engine/track.cpp:220-232
static void reset_playback_state(Track& t, double time_pos, bool refresh_voices) {
  if (!refresh_voices) {
    uint32_t idx = 0;
    t.has_clip_idx = find_next_clip(t, time_pos, &idx);
    t.clip_idx = idx;
    t.partially_ended = false;
  }
  t.refresh_voice = refresh_voices;
}

// ---- Engine -----------------------------------------------------------------------------------------------

Engine::Engine(int device_ordinal) {
  if (device_ordinal < 0) {  // scheduling-only engine: builds segment tables, cannot render (no CPU path)
    host_only_ = true;
    return;
  }
  int rc = wbx_create(&dev_, device_ordinal);
  if (rc != WBX_OK) {
    dev_ = nullptr;
    err_ = "wbx_create failed (no sm_100 CUDA device? there is no CPU path), status " + std::to_string(rc);
  }
}

Engine::~Engine() {
  for (auto* t : tracks) delete t;
  if (dev_) wbx_destroy(dev_);
}

const char* Engine::last_error() const { return dev_ ? wbx_last_error(dev_) : err_.c_str(); }

int Engine::set_audio_channel_config(uint32_t, uint32_t output_channels, uint32_t buffer_size, uint32_t sample_rate) {
  SpinGuard edit(editor_lock);
  if (!dev_ && !host_only_) return WBX_ERR_NO_DEVICE;
  if (output_channels < 1 || output_channels > 2 || buffer_size == 0 || buffer_size > 65535) return WBX_ERR_INVALID;
  if (dev_) {
    int rc = wbx_configure(dev_, output_channels, buffer_size, sample_rate);
    if (rc) return rc;
  }
  out_channels_ = output_channels;
  buffer_size_ = buffer_size;
  sample_rate_ = sample_rate;
  return WBX_OK;
}

void Engine::set_bpm(double bpm) {  // engine.cpp:24-30 (an atomic store there; the same lock as every other edit here)
  SpinGuard edit(editor_lock);
  beat_duration_ = 60.0 / bpm;
}

int Engine::set_track_effects(Track* track, const wbx_effect_params* params) {
  if (!track) return WBX_ERR_INVALID;
  SpinGuard edit(editor_lock);
  track->set_effects(params);
  return WBX_OK;
}

int Engine::set_track_plugin(Track* track, bool present) {
  if (!track) return WBX_ERR_INVALID;
  SpinGuard edit(editor_lock);
  track->has_plugin = present;
  return WBX_OK;
}

void Engine::set_playhead_position(double beat) {
  SpinGuard edit(editor_lock);
  playhead_start = beat;
  playhead = beat;
}

Track* Engine::add_track(const std::string& name) {
  SpinGuard edit(editor_lock);
  Track* t = new Track();
  t->name = name;
  // Track::Track() queues the defaults first (engine/track.cpp:22-27)
  t->set_volume(0.0f);
  t->set_pan(0.0f);
  t->set_mute(false);
  tracks.push_back(t);
  return t;
}

// Engine::delete_track (engine/engine.cpp:209-217)
int Engine::delete_track(uint32_t slot) {
  SpinGuard edit(editor_lock);
  if (slot >= tracks.size()) return WBX_ERR_INVALID;
  Track* t = tracks[slot];
  tracks.erase(tracks.begin() + slot);
  delete t;
  reindex_effects();
  return WBX_OK;
}

// The device addresses effect chains (extension) by track index: after the track list was reordered every chain is handed
// over again at the next render (its filter state restarts), and slots that now hold a chain-less track are cleared.
void Engine::reindex_effects() {
  for (Track* tr : tracks) tr->effects_dirty = true;
}

// Engine::move_track (engine/engine.cpp:228-243): the track order is the bus summation order
int Engine::move_track(uint32_t from_slot, uint32_t to_slot) {
  SpinGuard edit(editor_lock);
  if (from_slot >= tracks.size() || to_slot >= tracks.size()) return WBX_ERR_INVALID;
  if (from_slot == to_slot) return WBX_OK;
  Track* tmp = tracks[from_slot];
  if (from_slot < to_slot)
    for (uint32_t i = from_slot; i < to_slot; i++) tracks[i] = tracks[i + 1];
  else
    for (uint32_t i = from_slot; i > to_slot; i--) tracks[i] = tracks[i - 1];
  tracks[to_slot] = tmp;
  reindex_effects();
  return WBX_OK;
}

// Engine::solo_track (engine/engine.cpp:245-262): toggles the slot's solo flag and mutes / unmutes every other track
int Engine::solo_track(uint32_t slot) {
  SpinGuard edit(editor_lock);
  if (slot >= tracks.size()) return WBX_ERR_INVALID;
  bool mute = false;
  if (tracks[slot]->ui_parameter_state.solo) {
    tracks[slot]->ui_parameter_state.solo = false;
  } else {
    tracks[slot]->ui_parameter_state.solo = true;
    tracks[slot]->set_mute(false);
    mute = true;
  }
  for (uint32_t i = 0; i < tracks.size(); i++) {
    if (i == slot) continue;
    if (tracks[i]->ui_parameter_state.solo) tracks[i]->ui_parameter_state.solo = false;
    tracks[i]->set_mute(mute);
  }
  return WBX_OK;
}

// Engine::set_clip_gain (engine/engine.cpp:1460-1464)
int Engine::set_clip_gain(Track* track, uint32_t clip_id, float gain) {
  SpinGuard edit(editor_lock);
  if (!track || clip_id >= track->clips.size()) return WBX_ERR_INVALID;
  track->clips[clip_id]->gain = gain;
  return WBX_OK;
}

int Engine::add_sample(int format, uint32_t channels, uint64_t frames, uint32_t sample_rate, const void* const* planar) {
  SpinGuard edit(editor_lock);
  if (!dev_ && !host_only_) return WBX_ERR_NO_DEVICE;
  uint32_t id = (uint32_t)samples_.size();
  if (dev_) {
    int rc = wbx_sample_upload(dev_, format, channels, frames, sample_rate, planar, &id);
    if (rc) return rc;
  }
  if (samples_.size() <= id) samples_.resize(id + 1);
  samples_[id] = {frames, sample_rate};
  return (int)id;
}

// core/core_math.h:199-212
static inline double samples_to_beat(double samples, double sample_rate, double beat_duration) {
  const double sec = samples / sample_rate;
  return sec / beat_duration;
}

// shift_clip_content + calc_clip_shift for audio clips (engine/clip_edit.h:128-150): the clip's content start after its
// left edge moved by -relative_pos beats; sample_rate is the ASSET's rate.
static double shift_clip_content(const AudioClip* clip, double relative_pos, double beat_duration) {
  const double sample_rate = (double)clip->sample_rate;
  relative_pos *= clip->speed;
  const double offset_in_beat = samples_to_beat(clip->start_offset, sample_rate, beat_duration);
  const double shifted = offset_in_beat - relative_pos;
  return beat_to_samples(shifted > 0.0 ? shifted : 0.0, sample_rate, beat_duration);
}

// wb::find_lower_bound (core/algorithm.h:25-40): NOT std::lower_bound — the search ends at the last element
template <class Pred>
static uint32_t find_lower_bound_idx(const std::vector<AudioClip*>& clips, double value, Pred pred) {
  int64_t left = 0, right = (int64_t)clips.size() - 1;
  while (left < right) {
    const int64_t middle = (left + right) >> 1;
    if (pred(clips[(size_t)middle], value))
      left = middle + 1;
    else
      right = middle;
  }
  return (uint32_t)right;
}

// Track::query_clip_by_range (engine/track.cpp:112-157): first / last clip touched by [min, max]
static bool query_clip_by_range(const Track& t, double min, double max, uint32_t* first_out, uint32_t* last_out) {
  const auto& clips = t.clips;
  if (clips.empty()) return false;
  if (max <= clips.front()->min_time) return false;
  if (min >= clips.back()->max_time) return false;
  auto ends_before = [](const AudioClip* c, double time) { return c->max_time <= time; };
  uint32_t first_clip = find_lower_bound_idx(clips, min, ends_before);
  uint32_t last_clip = find_lower_bound_idx(clips, max, ends_before);
  const AudioClip* first = clips[first_clip];
  const AudioClip* last = clips[last_clip];
  if (first_clip == last_clip && (max <= first->min_time || min >= last->max_time)) return false;
  if (min > first->max_time) first_clip++;
  if (!(max > last->min_time)) last_clip--;
  *first_out = first_clip;
  *last_out = last_clip;
  return true;
}

// Track::mark_clip_deleted + update_clip_ordering (engine/track.cpp:107-110,159-180). Removed clips are parked in the
// track's graveyard instead of being destroyed: a voice that is still playing one keeps a valid pointer until the
// refresh at the next callback stops it.
static void update_clip_ordering(Track& t) {
  std::vector<AudioClip*> kept;
  for (AudioClip* c : t.clips) {
    if (c->deleted)
      t.graveyard.push_back(c);
    else
      kept.push_back(c);
  }
  t.clips.swap(kept);
  std::sort(t.clips.begin(), t.clips.end(), [](const AudioClip* a, const AudioClip* b) { return a->min_time < b->min_time; });
}

// Engine::reserve_track_region (engine/engine.cpp:478-569): make room for [min, max] by trimming, splitting or deleting
// the clips first_clip..last_clip (except ignore_clip, the clip being moved / resized).
void Engine::reserve_track_region(Track& t, uint32_t first_clip, uint32_t last_clip, double min, double max,
                                  const AudioClip* ignore_clip) {
  auto& clips = t.clips;
  if (clips.empty()) return;
  const double current_beat_duration = beat_duration_;
  if (first_clip == last_clip) {
    AudioClip* clip = clips[first_clip];
    if (clip == ignore_clip) return;
    if (min > clip->min_time && max < clip->max_time) {  // split the clip into two parts
      AudioClip* right = new AudioClip(*clip);
      right->deleted = right->internal_state_changed = false;  // Clip(const Clip&) does not copy them (clip.h:92-112)
      right->min_time = max;
      right->start_offset = shift_clip_content(right, clip->min_time - max, current_beat_duration);
      clip->max_time = min;
      clips.push_back(right);
    } else if (min > clip->min_time) {
      clip->max_time = min;
    } else if (max < clip->max_time) {
      clip->start_offset = shift_clip_content(clip, clip->min_time - max, current_beat_duration);
      clip->min_time = max;
    } else {
      clip->deleted = true;
    }
    return;
  }
  AudioClip* first = clips[first_clip];
  AudioClip* last = clips[last_clip];
  if (first != ignore_clip && min > first->min_time) {
    first->max_time = min;
    first_clip++;
  }
  if (last != ignore_clip && max < last->max_time) {
    last->start_offset = shift_clip_content(last, last->min_time - max, current_beat_duration);
    last->min_time = max;
    last_clip--;
  }
  if (first_clip <= last_clip && last_clip < clips.size())
    for (uint32_t i = first_clip; i <= last_clip; i++)
      if (clips[i] != ignore_clip) clips[i]->deleted = true;
}

// Engine::add_to_cliplist (engine/engine.cpp:409-461): a clip that overlaps existing ones trims / splits / deletes them
// first (reserve_track_region).
void Engine::add_to_cliplist(Track* track, AudioClip* clip) {
  auto& clips = track->clips;
  if (clips.empty() || clips.back()->max_time < clip->min_time) {  // first clip, or add to the back
    clips.push_back(clip);
  } else if (clips.front()->min_time > clip->max_time) {  // add to the front
    clips.insert(clips.begin(), clip);
  } else {
    uint32_t first = 0, last = 0;
    if (query_clip_by_range(*track, clip->min_time, clip->max_time, &first, &last))
      reserve_track_region(*track, first, last, clip->min_time, clip->max_time, nullptr);  // reserve space for the clip
    clips.push_back(clip);
    update_clip_ordering(*track);
  }
  reset_playback_state(*track, playhead, true);
}

// Engine::add_audio_clip (engine/engine.cpp:293-309)
int Engine::add_audio_clip(Track* track, double min_time, double max_time, double start_offset, uint32_t sample_id,
                           double speed, float gain, double fade_start, double fade_end) {
  SpinGuard edit(editor_lock);
  if (!track || sample_id >= samples_.size() || !(max_time >= min_time)) return WBX_ERR_INVALID;
  AudioClip* clip = new AudioClip();
  clip->min_time = min_time;
  clip->max_time = max_time;
  clip->start_offset = start_offset;
  clip->speed = speed;
  clip->gain = gain;
  clip->fade_start = fade_start;
  clip->fade_end = fade_end;
  clip->sample_id = sample_id;
  clip->sample_rate = samples_[sample_id].rate;
  add_to_cliplist(track, clip);
  return WBX_OK;
}

static bool owns_clip(const Track* track, const AudioClip* clip) {
  return track && clip && std::find(track->clips.begin(), track->clips.end(), clip) != track->clips.end();
}

// Engine::duplicate_clip (engine/engine.cpp:336-344)
int Engine::duplicate_clip(Track* track, const AudioClip* clip_to_duplicate, double min_time, double max_time) {
  SpinGuard edit(editor_lock);
  if (!owns_clip(track, clip_to_duplicate) || !(max_time >= min_time)) return WBX_ERR_INVALID;
  AudioClip* clip = new AudioClip(*clip_to_duplicate);
  clip->deleted = clip->internal_state_changed = false;  // Clip(const Clip&) does not copy them (clip.h:92-112)
  clip->min_time = min_time;
  clip->max_time = max_time;
  add_to_cliplist(track, clip);
  return WBX_OK;
}

// Engine::move_clip (engine/engine.cpp:346-363) + calc_move_clip (engine/clip_edit.h:10-16). A clip moved while it plays
// is stopped and restarted at its new content offset by the next callback (internal_state_changed, track.cpp:394-419).
int Engine::move_clip(Track* track, AudioClip* clip, double relative_pos) {
  SpinGuard edit(editor_lock);
  if (!owns_clip(track, clip)) return WBX_ERR_INVALID;
  if (relative_pos == 0.0) return WBX_OK;
  double new_pos = clip->min_time + relative_pos;
  if (!(new_pos > 0.0)) new_pos = 0.0;  // math::max(.., min_move = 0.0)
  const double min_time = new_pos, max_time = new_pos + (clip->max_time - clip->min_time);
  uint32_t first = 0, last = 0;
  if (query_clip_by_range(*track, min_time, max_time, &first, &last))
    reserve_track_region(*track, first, last, min_time, max_time, clip);
  clip->min_time = min_time;
  clip->max_time = max_time;
  clip->internal_state_changed = true;
  update_clip_ordering(*track);
  reset_playback_state(*track, playhead, true);
  return WBX_OK;
}

// Engine::resize_clip (engine/engine.cpp:365-398) + calc_resize_clip (engine/clip_edit.h:18-126, clamp_at_resize_pos off):
// drag the left or right edge by relative_pos beats; `shift` keeps the content in place instead of the edge, `stretch`
// changes the clip speed so that the same content fills the new length.
int Engine::resize_clip(Track* track, AudioClip* clip, double relative_pos, double resize_limit, double min_length,
                        bool left_side, bool shift, bool stretch) {
  SpinGuard edit(editor_lock);
  if (!owns_clip(track, clip)) return WBX_ERR_INVALID;
  if (relative_pos == 0.0) return WBX_OK;
  const double beat_duration = beat_duration_;
  const double asset_rate = (double)clip->sample_rate;
  const double sample_count = (double)samples_[clip->sample_id].count;
  double min_time, max_time, start_offset = clip->start_offset, new_speed = 1.0;
  if (!left_side) {
    const double old_max = clip->max_time;
    const double actual_min_length = resize_limit + min_length - clip->min_time;
    double new_max = clip->max_time + relative_pos;
    if (!(new_max > 0.0)) new_max = 0.0;
    const double length = new_max - clip->min_time;
    if (length < actual_min_length) new_max = clip->min_time + actual_min_length;
    if (shift) {
      const double mult = clip->speed;
      start_offset = samples_to_beat(start_offset, asset_rate, beat_duration);
      if (old_max < new_max)
        start_offset -= (new_max - old_max) * mult;
      else
        start_offset += (old_max - new_max) * mult;
      if (!(start_offset > 0.0)) start_offset = 0.0;
      if (!(start_offset < sample_count)) start_offset = sample_count;  // math::min(start_offset, count)
      start_offset = beat_to_samples(start_offset, asset_rate, beat_duration);
    }
    if (stretch) {
      const double old_length = sample_count / clip->speed;
      const double num_samples = beat_to_samples(relative_pos, asset_rate, beat_duration);
      new_speed = sample_count / (old_length + num_samples);
    }
    min_time = clip->min_time;
    max_time = new_max;
  } else {
    const double old_min = clip->min_time;
    const double actual_min_length = clip->max_time - resize_limit + min_length;
    double new_min = clip->min_time + relative_pos;
    if (!(new_min > 0.0)) new_min = 0.0;
    const double length = clip->max_time - new_min;
    if (length < actual_min_length) new_min = clip->max_time - actual_min_length;
    if (!shift) {
      start_offset = samples_to_beat(start_offset, asset_rate, beat_duration);
      if (old_min < new_min)
        start_offset -= old_min - new_min;
      else
        start_offset += new_min - old_min;
      if (start_offset < 0.0) new_min = new_min - start_offset;
      if (!(start_offset > 0.0)) start_offset = 0.0;
      start_offset = beat_to_samples(start_offset, asset_rate, beat_duration);
    }
    if (stretch) {
      const double old_length = sample_count / clip->speed;
      const double num_samples = beat_to_samples(old_min - new_min, asset_rate, beat_duration);
      new_speed = sample_count / (old_length + num_samples);
    }
    min_time = new_min;
    max_time = clip->max_time;
  }
  uint32_t first = 0, last = 0;
  if (query_clip_by_range(*track, min_time, max_time, &first, &last))
    reserve_track_region(*track, first, last, min_time, max_time, clip);
  if (left_side)
    clip->min_time = min_time;
  else
    clip->max_time = max_time;
  clip->start_offset = start_offset;
  if (stretch) clip->speed = new_speed;
  clip->internal_state_changed = shift || stretch;
  update_clip_ordering(*track);
  reset_playback_state(*track, playhead, true);
  return WBX_OK;
}

// Engine::delete_region(track, min, max) (engine/engine.cpp:463-473): erase a time range — the clips it touches are
// trimmed, split or deleted.
int Engine::delete_region(Track* track, double min, double max) {
  SpinGuard edit(editor_lock);
  if (!track || !(max >= min)) return WBX_ERR_INVALID;
  uint32_t first = 0, last = 0;
  if (!query_clip_by_range(*track, min, max, &first, &last)) return WBX_OK;
  reserve_track_region(*track, first, last, min, max, nullptr);
  update_clip_ordering(*track);
  reset_playback_state(*track, playhead, true);
  return WBX_OK;
}

// Engine::delete_clip (engine/engine.cpp:400-407). The clip is parked in the track's graveyard (see update_clip_ordering).
int Engine::delete_clip(Track* track, AudioClip* clip) {
  SpinGuard edit(editor_lock);
  if (!owns_clip(track, clip)) return WBX_ERR_INVALID;
  clip->deleted = true;
  update_clip_ordering(*track);
  reset_playback_state(*track, playhead, true);
  return WBX_OK;
}

int Engine::set_impulse_response(const float* h, uint32_t n_taps) {
  SpinGuard edit(editor_lock);
  if (!dev_) return WBX_ERR_NO_DEVICE;
  return wbx_set_impulse_response(dev_, h, n_taps);
}

void Engine::play() {
  SpinGuard edit(editor_lock);
  for (auto* t : tracks) reset_playback_state(*t, playhead_start, false);
  sample_position = 0;
  playing = true;
}

void Engine::stop() {
  SpinGuard edit(editor_lock);
  playing = false;
  playhead = playhead_start;
  for (auto* t : tracks) {  // Track::stop, engine/track.cpp:248-256
    t->current_audio_event = AudioEvent();
    t->audio_event_buffer.clear();
  }
}

// Track::process_event (engine/track.cpp:258-451), audio clips.
void Engine::process_event(Track& t, double start_time, double end_time, double sample_position_,
                           double beat_duration, double sample_rate, uint32_t buffer_size) {
  auto push_stop = [&](uint32_t off, double time) {
    AudioEvent e;
    e.type = EventType::StopSample;
    e.buffer_offset = off;
    e.time = time;
    t.audio_event_buffer.push_back(e);
  };
  auto push_play = [&](uint32_t off, double time, const AudioClip* clip, uint64_t sample_offset,
                       uint64_t clip_frame = 0) {
    AudioEvent e;
    e.type = EventType::PlaySample;
    e.buffer_offset = off;
    e.time = time;
    e.speed = clip->speed;
    e.sample_offset = sample_offset;
    e.clip_frame = clip_frame;
    e.clip = clip;
    t.audio_event_buffer.push_back(e);
  };

  if (t.clips.empty()) {
    if (t.refresh_voice) {
      push_stop(0, start_time);
      t.has_clip_idx = false;
      t.refresh_voice = false;
    }
    return;
  }
  const uint32_t num_clips = (uint32_t)t.clips.size();
  if (t.refresh_voice) {
    uint32_t at = 0;
    if (find_next_clip(t, start_time, &at)) {
      if (t.has_clip_idx) {
        const uint32_t idx = t.clip_idx;
        if (idx < num_clips) {
          const AudioClip* clip = t.clips[at];
          const AudioClip* current_clip = t.clips[idx];
          const bool inside = start_time >= clip->min_time && start_time <= clip->max_time;
          if ((clip != current_clip && inside) || (clip == current_clip && !inside)) {
            push_stop(0, start_time);
            t.clip_idx = at;
            t.partially_ended = false;
          }
        }
      } else {
        t.has_clip_idx = true;
        t.clip_idx = at;
      }
    } else {
      push_stop(0, start_time);
      t.has_clip_idx = false;
    }
    t.refresh_voice = false;
  }
  if (!t.has_clip_idx) return;

  uint32_t next_clip = t.clip_idx;
  while (next_clip < num_clips) {
    AudioClip* clip = t.clips[next_clip];
    const double min_time = clip->min_time, max_time = clip->max_time;
    if (min_time > end_time) break;

    if (min_time >= start_time) {  // the clip starts inside this callback
      const double offset_from_start = beat_to_samples(min_time - start_time, sample_rate, beat_duration);
      const double sample_offset = sample_position_ + offset_from_start;
      const uint32_t buffer_offset = (uint32_t)((uint64_t)sample_offset % (uint64_t)buffer_size);
      push_play(buffer_offset, min_time, clip, (uint64_t)clip->start_offset);
      clip->internal_state_changed = false;
    } else if (start_time > min_time && !t.partially_ended) {  // playback starts in the middle of the clip
      const double sample_pos = beat_to_samples(start_time - min_time, sample_rate, beat_duration);
      const uint64_t sample_offset = (uint64_t)(clip->start_offset + (sample_pos * clip->speed));
      push_play(0, start_time, clip, sample_offset, (uint64_t)sample_pos);
      clip->internal_state_changed = false;
    } else if (clip->internal_state_changed && t.partially_ended) {  // clip edited while it plays
      const double sample_pos = beat_to_samples(start_time - min_time, sample_rate, beat_duration);
      const uint64_t sample_offset = (uint64_t)(clip->start_offset + (sample_pos * clip->speed));
      push_stop(0, start_time);
      push_play(0, start_time, clip, sample_offset, (uint64_t)sample_pos);
      clip->internal_state_changed = false;
    }

    if (max_time <= end_time) {  // the clip ends inside this callback
      const double offset_from_start = beat_to_samples(max_time - start_time, sample_rate, beat_duration);
      const double sample_offset = sample_position_ + offset_from_start;
      const uint32_t buffer_offset = (uint32_t)((uint64_t)sample_offset % (uint64_t)buffer_size);
      push_stop(buffer_offset, max_time);
      t.partially_ended = false;
    } else {
      t.partially_ended = true;
      break;
    }
    next_clip++;
  }
  t.clip_idx = next_clip;
}

// Fade extension (include/wbx.h): ramp lengths of the clip in output frames.
void Engine::fill_fade(wbx_segment& s, const AudioClip* clip, uint64_t clip_frame) const {
  s.flags = (resampler_mode == 1 && s.speed != 1.0) ? WBX_SEG_POLYPHASE : 0u;
  s.clip_frame = s.fade_in_frames = s.fade_out_frames = s.clip_len_frames = 0.0;
  if (clip->fade_start > 0.0 || clip->fade_end > 0.0) {
    s.flags |= WBX_SEG_FADE;
    s.clip_frame = (double)clip_frame;
    s.fade_in_frames = beat_to_samples(clip->fade_start, cur_sample_rate_, beat_duration_);
    s.fade_out_frames = beat_to_samples(clip->fade_end, cur_sample_rate_, beat_duration_);
    s.clip_len_frames = beat_to_samples(clip->max_time - clip->min_time, cur_sample_rate_, beat_duration_);
  }
}

// One dsp::Sampler::stream call (dsp/sampler.cpp:88-210) becomes one wbx_segment — or extends the track's
// open run when it is the whole-block continuation of the previous callback's call. Host keeps only the
// position bookkeeping (sampler.cpp:99-103,209); the device clips to the sample's end itself.
void Engine::stream(Track& t, uint32_t track_index, uint32_t block, uint32_t num_samples, uint32_t buffer_offset) {
  const AudioClip* clip = t.current_audio_event.clip;
  const double count = (double)samples_[clip->sample_id].count;
  const uint64_t clip_frame = t.clip_frame;
  t.clip_frame += num_samples;  // the clip's own timeline advances whether or not the sample still has data
  if (t.sample_offset >= count) return;  // has finished streaming: position no longer advances
  // a plugin in the slot: the clip is rendered into the plugin's effect_buffer and never mixed (track.cpp:600,645-724)
  // — no segment reaches the device, the sampler's bookkeeping below runs all the same
  if (num_samples != 0 && !t.has_plugin) {
    bool extended = false;
    if (t.open_run >= 0 && buffer_offset == 0 && num_samples == buffer_size_) {
      wbx_segment& r = segs_[t.open_run];
      if (r.block + r.n_blocks == block && r.length == buffer_size_ && r.dst_offset == 0) {
        r.n_blocks++;
        extended = true;
      }
    }
    if (!extended) {
      wbx_segment s;
      s.track = track_index;
      s.block = block;
      s.n_blocks = 1;
      s.dst_offset = buffer_offset;
      s.length = num_samples;
      s.sample_id = clip->sample_id;
      s.src_pos = t.sample_offset;
      s.speed = t.playback_speed;
      s.gain = clip->gain;
      fill_fade(s, clip, clip_frame);
      segs_.push_back(s);
      // only a whole-block call can be continued by the next callback's whole-block call
      t.open_run = (buffer_offset == 0 && num_samples == buffer_size_) ? (int32_t)segs_.size() - 1 : -1;
    }
  }
  t.sample_offset = t.sample_offset + ((double)num_samples * t.playback_speed);
}

// Track::process (engine/track.cpp:587-736) minus the sample loops.
void Engine::track_block(Track& t, uint32_t track_index, uint32_t block, double sample_rate, double beat_duration,
                         double start_time, double end_time, double block_sample_position, bool currently_playing) {
  t.audio_event_buffer.clear();  // engine.cpp:1591
  if (currently_playing)
    process_event(t, start_time, end_time, block_sample_position, beat_duration, sample_rate, buffer_size_);

  Track::Msg m;
  while (t.track_msg_queue.pop(m)) {  // process_track_messages + parameter application, :773-779, :618-643
    switch (m.id) {
      case kParamVolume: t.parameter_state.volume = (float)m.value; break;
      case kParamPan: {
        t.parameter_state.pan = (float)m.value;
        const PanningCoefficient pc = calculate_panning_coefs(t.parameter_state.pan);
        t.parameter_state.pan_coeffs[0] = pc.left;
        t.parameter_state.pan_coeffs[1] = pc.right;
        break;
      }
      case kParamMute: t.parameter_state.mute = m.value > 0.0; break;
    }
  }

  if (!currently_playing) {
    t.open_run = -1;
    return;
  }
  // walk the callback's events, splitting the block at each buffer_offset (:664-724)
  const uint32_t B = buffer_size_;
  size_t next = 0;
  uint32_t start_sample = 0;
  if (!t.audio_event_buffer.empty()) t.open_run = -1;  // any event ends the run this track was extending
  while (start_sample < B) {
    if (next != t.audio_event_buffer.size()) {
      const AudioEvent& ne = t.audio_event_buffer[next];
      uint32_t event_length = ne.buffer_offset - start_sample;
      if (ne.buffer_offset < start_sample || ne.buffer_offset > B) {
        // Events out of offset order: undefined behaviour in the reference (event_length wraps, :670, and
        // Sampler::stream writes past the buffer; happens when a clip ends exactly on the callback's end,
        // :423-425). Defined here as: render nothing for this event and silence the voice.
        event_length = 0;
        if (t.current_audio_event.type == EventType::PlaySample) t.current_audio_event.type = EventType::None;
      }
      if (t.current_audio_event.type == EventType::PlaySample)
        stream(t, track_index, block, event_length, start_sample);
      if (ne.type == EventType::PlaySample) {  // Sampler::reset_state, dsp/sampler.h:18-27
        t.playback_speed = ((double)ne.clip->sample_rate / sample_rate) * ne.speed;
        t.sample_offset = (double)ne.sample_offset;
        t.clip_frame = ne.clip_frame;
      }
      t.current_audio_event = ne;
      start_sample += event_length;
      next++;
    } else {
      const uint32_t event_length = B - start_sample;
      if (t.current_audio_event.type == EventType::PlaySample)
        stream(t, track_index, block, event_length, start_sample);
      start_sample = B;
    }
  }
}

// How many of the callbacks [k, K) are provably event-free for this track: Track::process_event would push
// no event and leave its state untouched (engine/track.cpp:347-446), and no parameter message is pending.
uint32_t Engine::quiet_blocks(const Track& t, uint32_t k, uint32_t K) const {
  if (!t.track_msg_queue.empty()) return 0;
  if (!playing) return K - k;
  if (t.refresh_voice) return 0;
  if (t.clips.empty() || !t.has_clip_idx || t.clip_idx >= t.clips.size()) return K - k;
  const AudioClip* clip = t.clips[t.clip_idx];
  if (clip->internal_state_changed) return 0;
  double key;
  if (t.partially_ended) {  // inside the clip: nothing happens until `max_time <= end_time` (:423)
    if (!(blk_start_[k] > clip->min_time)) return 0;
    key = clip->max_time;
  } else {  // waiting for the clip: `min_time > end_time` breaks out immediately (:353)
    key = clip->min_time;
  }
  // first callback whose end_time reaches `key` (end times are non-decreasing)
  const double* first = blk_end_.data() + k;
  const double* last = blk_end_.data() + K;
  const double* it = std::lower_bound(first, last, key);  // first end_time >= key
  return (uint32_t)(it - first);
}

// `while (steps < n && off < limit) { off = off + adv; steps++; }` in O(binades): see advance_rounded_impl (wbx_device.cuh)
uint32_t advance_rounded(double* off_io, double adv, uint32_t n, double limit) {
  return advance_rounded_impl(off_io, adv, n, limit);
}

// q consecutive event-free callbacks: each one is `stream(whole block)` when a sample is playing (:713-719).
void Engine::stream_run(Track& t, uint32_t track_index, uint32_t block, uint32_t q) {
  if (t.current_audio_event.type != EventType::PlaySample || q == 0) return;
  const AudioClip* clip = t.current_audio_event.clip;
  const double count = (double)samples_[clip->sample_id].count;
  const uint32_t B = buffer_size_;
  if (t.sample_offset >= count) {
    t.clip_frame += (uint64_t)q * B;
    return;
  }
  const double adv = (double)B * t.playback_speed;
  double off = t.sample_offset;
  uint32_t streamed = 0;
  if (t.playback_speed == 1.0 && off + (double)q * (double)B < 4.0e15) {
    // integers: the per-callback additions are exact, so the recurrence has a closed form
    const double calls = std::ceil((count - off) / (double)B);  // calls made before the sample is exhausted
    streamed = calls < (double)q ? (uint32_t)calls : q;
    off = off + (double)streamed * (double)B;
  } else {
    streamed = advance_rounded(&off, adv, q, count);  // the reference's own recurrence, one rounding per callback
  }
  bool extended = t.has_plugin || streamed == 0;  // (a plugin in the slot: nothing of the clip reaches the device)
  if (!extended && t.open_run >= 0) {
    wbx_segment& r = segs_[t.open_run];
    if (r.block + r.n_blocks == block && r.length == B && r.dst_offset == 0) {
      r.n_blocks += streamed;
      extended = true;
    }
  }
  if (!extended) {
    wbx_segment s;
    s.track = track_index;
    s.block = block;
    s.n_blocks = streamed;
    s.dst_offset = 0;
    s.length = B;
    s.sample_id = clip->sample_id;
    s.src_pos = t.sample_offset;
    s.speed = t.playback_speed;
    s.gain = clip->gain;
    fill_fade(s, clip, t.clip_frame);
    segs_.push_back(s);
    t.open_run = (int32_t)segs_.size() - 1;
  }
  t.clip_frame += (uint64_t)q * B;
  t.sample_offset = off;
}

int Engine::schedule(uint32_t n_blocks, double sample_rate) {
  SpinGuard edit(editor_lock);
  return schedule_locked(n_blocks, sample_rate);
}

int Engine::schedule_locked(uint32_t n_blocks, double sample_rate) {
  if (sample_rate == 0.0) sample_rate = (double)sample_rate_;
  cur_sample_rate_ = sample_rate;
  segs_.clear();
  const uint32_t N = (uint32_t)tracks.size();
  const uint32_t K = n_blocks;
  // transport arithmetic of Engine::process (engine.cpp:1578-1585, 1619-1623), one entry per callback
  blk_start_.resize(K);
  blk_end_.resize(K);
  blk_spos_.resize(K);
  const bool currently_playing = playing;
  const double current_beat_duration = beat_duration_;
  {
    double ph = playhead, sp = sample_position;
    for (uint32_t k = 0; k < K; k++) {
      const double buffer_duration = (double)buffer_size_ / sample_rate;
      const double buffer_duration_in_beats = buffer_duration / current_beat_duration;
      const double next_playhead_pos = ph + buffer_duration_in_beats;
      blk_start_[k] = ph;
      blk_end_[k] = next_playhead_pos;
      blk_spos_[k] = sp;
      if (currently_playing) {
        sp += beat_to_samples(buffer_duration_in_beats, sample_rate, current_beat_duration);
        ph = next_playhead_pos;
      }
    }
    playhead = ph;
    sample_position = sp;
  }
  // Tracks are independent until the bus sum, so walk each track through all callbacks in turn; callbacks
  // that provably hold no event for the track are skipped in closed form (fast_forward).
  for (uint32_t i = 0; i < N; i++) {
    if (i + 2 < N) __builtin_prefetch(tracks[i + 2]);
    Track& t = *tracks[i];
    t.open_run = -1;
    uint32_t k = 0;
    while (k < K) {
      // event-free callbacks (the steady state of every playing or idle track): Track::process reduces to one
      // whole-block Sampler::stream call per callback, taken in closed form without walking process_event
      const uint32_t q = fast_forward ? quiet_blocks(t, k, K) : 0;
      if (q) {
        if (currently_playing) stream_run(t, i, k, q);
        k += q;
        continue;
      }
      track_block(t, i, k, sample_rate, current_beat_duration, blk_start_[k], blk_end_[k], blk_spos_[k],
                  currently_playing);
      k++;
    }
  }
  // gain used this render = (mute ? 0 : volume) * pan_coeffs[ch]   (track.cpp:728-731)
  gains_.resize((size_t)N * 2);
  for (uint32_t i = 0; i < N; i++) {
    const TrackParameterState& ps = tracks[i]->parameter_state;
    const float volume = ps.mute ? 0.0f : ps.volume;
    gains_[2 * i + 0] = volume * ps.pan_coeffs[0];
    gains_[2 * i + 1] = volume * ps.pan_coeffs[1];
  }
  return WBX_OK;
}

// track count + edited effect chains -> device, then the host schedule of the next n_blocks callbacks
int Engine::prepare(uint32_t n_blocks, double sample_rate) {
  if (!dev_) return WBX_ERR_NO_DEVICE;
  if (n_blocks == 0) return WBX_ERR_INVALID;
  const uint32_t N = (uint32_t)tracks.size();
  int rc = wbx_set_track_count(dev_, N);
  if (rc) return rc;
  for (uint32_t i = 0; i < N; i++) {  // effect chains edited since the last render
    Track& t = *tracks[i];
    if (!t.effects_dirty) continue;
    wbx_effects fx;
    if (t.effects_on) {
      if ((rc = wbx_effects_design(&t.effect_params, sample_rate_, &fx))) return rc;
      rc = wbx_set_track_effects(dev_, i, &fx);
    } else {
      rc = wbx_set_track_effects(dev_, i, nullptr);
    }
    if (rc) return rc;
    t.effects_dirty = false;
  }
  return schedule_locked(n_blocks, sample_rate);
}

// VUMeter::push_samples: level only rises until the UI reads it (vu_meter.h:25-29)
void Engine::merge_levels() {
  const uint32_t N = (uint32_t)tracks.size();
  for (uint32_t i = 0; i < N; i++)
    for (uint32_t c = 0; c < 2; c++)
      tracks[i]->level[c].push(levels_[2 * i + c]);
}

// ScopedPerformanceCounter + perf_measurer.update (engine.cpp:1577,1653): wall time of the call against the audio it made
void Engine::meter(std::chrono::steady_clock::time_point t0, uint32_t n_blocks) {
  const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  const double buffer_ms = 1000.0 * (double)buffer_size_ / cur_sample_rate_ * (double)n_blocks;
  perf_measurer.update(ms, buffer_ms);
}

int Engine::render(uint32_t n_blocks, float* const* out_channels, float* peaks, double sample_rate) {
  const auto t0 = std::chrono::steady_clock::now();
  SpinGuard edit(editor_lock);  // held for the whole callback, like Engine::process (engine.cpp:1587,1651)
  int rc = prepare(n_blocks, sample_rate);
  if (rc) return rc;
  const uint32_t N = (uint32_t)tracks.size();
  // submit + mix + bus/peaks/levels back under one synchronise; on an engine that is part of a sharded setup
  // (wbx_shard_*) the mix is the sharded one and only rank 0 receives the master bus
  if (N) levels_.resize((size_t)N * 2);
  rc = wbx_render_levels(dev_, segs_.data(), (uint32_t)segs_.size(), gains_.data(), n_blocks, out_channels, peaks,
                         N ? levels_.data() : nullptr);
  if (rc) return rc;
  merge_levels();
  meter(t0, n_blocks);
  return WBX_OK;
}

// Offline bounce: the batched render as an export driver. Chunk i's conversion runs right behind its mix, its copy to the
// host on the device engine's second stream and the sink's write on this thread, all under chunk i+1's schedule + mix.
int Engine::bounce(double start_beat, double end_beat, int dst_format, BounceSink& sink, uint32_t chunk_blocks,
                   uint64_t* frames_out) {
  if (frames_out) *frames_out = 0;
  if (!dev_) return WBX_ERR_NO_DEVICE;
  if (!(end_beat > start_beat) || chunk_blocks == 0) return WBX_ERR_INVALID;
  size_t es;
  switch (dst_format) {
    case WBX_FMT_I16: es = 2; break;
    case WBX_FMT_I24_X8:
    case WBX_FMT_I32:
    case WBX_FMT_F32: es = 4; break;
    default: return WBX_ERR_UNSUPPORTED;  // (the reference's packed-I24 writer has no channel stride: not a file format)
  }
  stop();
  set_playhead_position(start_beat);
  play();
  const auto t0 = std::chrono::steady_clock::now();
  SpinGuard edit(editor_lock);  // an export owns the engine like a (long) callback does
  const double total = std::ceil(beat_to_samples(end_beat - start_beat, (double)sample_rate_, beat_duration_));
  const uint64_t total_frames = total > 0.0 ? (uint64_t)total : 0;
  const uint64_t total_blocks = (total_frames + buffer_size_ - 1) / buffer_size_;
  int rc = wbx_bounce_begin(dev_, dst_format);
  uint64_t done_blocks = 0, in_flight = 0, written = 0;
  auto pop_one = [&]() -> int {
    const void* data = nullptr;
    size_t bytes = 0;
    int r = wbx_bounce_pop(dev_, &data, &bytes);
    if (r) return r;
    const size_t frame_bytes = es * out_channels_;
    uint64_t frames = bytes / frame_bytes;
    if (written + frames > total_frames) frames = total_frames - written;  // the last callback is cut at end_beat
    written += frames;
    in_flight--;
    return frames ? sink.write(data, (size_t)frames * frame_bytes) : WBX_OK;
  };
  while (!rc && done_blocks < total_blocks) {
    const uint32_t n = (uint32_t)std::min<uint64_t>(chunk_blocks, total_blocks - done_blocks);
    if ((rc = prepare(n, 0.0))) break;
    if ((rc = wbx_submit(dev_, segs_.data(), (uint32_t)segs_.size(), gains_.data(), n))) break;
    if ((rc = wbx_mix(dev_, 0))) break;
    if ((rc = wbx_bounce_push(dev_))) break;
    in_flight++;
    done_blocks += n;
    if (in_flight == 2) rc = pop_one();  // the older chunk: its copy-out ran under the mix just enqueued
  }
  while (!rc && in_flight) rc = pop_one();
  if (!rc && tracks.size()) {  // VU levels of the last chunk, as after any render
    levels_.resize(tracks.size() * 2);
    if (!(rc = wbx_fetch_levels(dev_, levels_.data()))) merge_levels();
  }
  if (!rc) rc = wbx_synchronize(dev_);
  meter(t0, (uint32_t)std::min<uint64_t>(total_blocks, 0xFFFFFFFFu));
  // Engine::stop (engine.cpp:82-92) without re-taking the lock
  playing = false;
  playhead = playhead_start;
  for (Track* t : tracks) {
    t->current_audio_event = AudioEvent();
    t->audio_event_buffer.clear();
  }
  if (frames_out) *frames_out = written;
  return rc;
}

// render() in two halves for a thread that drives several sharded engines (wbx_sharded.hpp): render_begin on every
// engine, the phases of wbx_mix_sharded_phase on every engine's device(), render_end on every engine.
int Engine::render_begin(uint32_t n_blocks, double sample_rate) {
  SpinGuard edit(editor_lock);
  int rc = prepare(n_blocks, sample_rate);
  if (rc) return rc;
  return wbx_submit(dev_, segs_.data(), (uint32_t)segs_.size(), gains_.data(), n_blocks);
}

int Engine::render_end(float* const* out_channels, float* peaks) {
  if (!dev_) return WBX_ERR_NO_DEVICE;
  SpinGuard edit(editor_lock);
  const uint32_t N = (uint32_t)tracks.size();
  int rc = wbx_fetch(dev_, out_channels, peaks);
  if (rc) return rc;
  if (N) {
    levels_.resize((size_t)N * 2);
    if ((rc = wbx_fetch_levels(dev_, levels_.data()))) return rc;
    merge_levels();
  }
  return WBX_OK;
}

}  // namespace wbx

// ---- C exports of the host engine (include/wbx_host.h), for ctypes / other FFI -----------------------------

using wbx::Engine;

extern "C" {

struct wbxh_engine {
  Engine eng;
  explicit wbxh_engine(int dev) : eng(dev) {}
};

int wbxh_create(wbxh_engine** out, int device_ordinal, uint32_t out_channels, uint32_t block_frames,
                uint32_t sample_rate, double bpm) {
  if (!out) return WBX_ERR_INVALID;
  *out = nullptr;
  wbxh_engine* h = new wbxh_engine(device_ordinal);
  if (!h->eng.ok() && device_ordinal >= 0) {
    delete h;
    return WBX_ERR_NO_DEVICE;
  }
  int rc = h->eng.set_audio_channel_config(0, out_channels, block_frames, sample_rate);
  if (rc) {
    delete h;
    return rc;
  }
  h->eng.set_bpm(bpm);
  *out = h;
  return WBX_OK;
}
void wbxh_destroy(wbxh_engine* h) { delete h; }
const char* wbxh_last_error(wbxh_engine* h) { return h ? h->eng.last_error() : "null"; }
wbx_engine* wbxh_device(wbxh_engine* h) { return h ? h->eng.device() : nullptr; }

int wbxh_add_track(wbxh_engine* h, float volume_db, float pan, int mute) {
  wbx::Track* t = h->eng.add_track("t" + std::to_string(h->eng.tracks.size()));
  t->set_volume(volume_db);
  t->set_pan(pan);
  t->set_mute(mute != 0);
  return (int)h->eng.tracks.size() - 1;
}
static wbx::Track* track_at(wbxh_engine* h, int track) {
  return (!h || track < 0 || (size_t)track >= h->eng.tracks.size()) ? nullptr : h->eng.tracks[track];
}
int wbxh_set_volume(wbxh_engine* h, int track, float db) {
  wbx::Track* t = track_at(h, track);
  if (!t) return WBX_ERR_INVALID;
  t->set_volume(db);
  return WBX_OK;
}
int wbxh_set_pan(wbxh_engine* h, int track, float pan) {
  wbx::Track* t = track_at(h, track);
  if (!t) return WBX_ERR_INVALID;
  t->set_pan(pan);
  return WBX_OK;
}
int wbxh_set_mute(wbxh_engine* h, int track, int mute) {
  wbx::Track* t = track_at(h, track);
  if (!t) return WBX_ERR_INVALID;
  t->set_mute(mute != 0);
  return WBX_OK;
}
int wbxh_set_plugin(wbxh_engine* h, int track, int present) {
  wbx::Track* t = track_at(h, track);
  return t ? h->eng.set_track_plugin(t, present != 0) : WBX_ERR_INVALID;
}
int wbxh_configure(wbxh_engine* h, uint32_t out_channels, uint32_t block_frames, uint32_t sample_rate) {
  return h ? h->eng.set_audio_channel_config(0, out_channels, block_frames, sample_rate) : WBX_ERR_INVALID;
}
double wbxh_cpu_usage(wbxh_engine* h) { return h ? h->eng.cpu_usage() : 0.0; }
int wbxh_add_sample(wbxh_engine* h, int format, uint32_t channels, uint64_t frames, uint32_t sample_rate,
                    const void* const* planar) {
  return h->eng.add_sample(format, channels, frames, sample_rate, planar);
}
int wbxh_add_clip(wbxh_engine* h, int track, int sample, double min_beat, double max_beat, double start_offset,
                  double speed, float gain) {
  if (track < 0 || (size_t)track >= h->eng.tracks.size() || sample < 0) return WBX_ERR_INVALID;
  return h->eng.add_audio_clip(h->eng.tracks[track], min_beat, max_beat, start_offset, (uint32_t)sample, speed, gain);
}
int wbxh_add_clip_fade(wbxh_engine* h, int track, int sample, double min_beat, double max_beat, double start_offset,
                       double speed, float gain, double fade_start, double fade_end) {
  if (track < 0 || (size_t)track >= h->eng.tracks.size() || sample < 0) return WBX_ERR_INVALID;
  return h->eng.add_audio_clip(h->eng.tracks[track], min_beat, max_beat, start_offset, (uint32_t)sample, speed, gain,
                               fade_start, fade_end);
}
static wbx::AudioClip* clip_at(wbxh_engine* h, int track, int clip) {
  if (track < 0 || (size_t)track >= h->eng.tracks.size()) return nullptr;
  auto& clips = h->eng.tracks[track]->clips;
  return (clip < 0 || (size_t)clip >= clips.size()) ? nullptr : clips[clip];
}
int wbxh_delete_track(wbxh_engine* h, int track) { return track < 0 ? WBX_ERR_INVALID : h->eng.delete_track((uint32_t)track); }
int wbxh_move_track(wbxh_engine* h, int from_slot, int to_slot) {
  return (from_slot < 0 || to_slot < 0) ? WBX_ERR_INVALID : h->eng.move_track((uint32_t)from_slot, (uint32_t)to_slot);
}
int wbxh_solo_track(wbxh_engine* h, int track) { return track < 0 ? WBX_ERR_INVALID : h->eng.solo_track((uint32_t)track); }
int wbxh_set_clip_gain(wbxh_engine* h, int track, int clip, float gain) {
  if (track < 0 || (size_t)track >= h->eng.tracks.size() || clip < 0) return WBX_ERR_INVALID;
  return h->eng.set_clip_gain(h->eng.tracks[track], (uint32_t)clip, gain);
}
int wbxh_clip_count(wbxh_engine* h, int track) {
  return (track < 0 || (size_t)track >= h->eng.tracks.size()) ? WBX_ERR_INVALID : (int)h->eng.tracks[track]->clips.size();
}
int wbxh_clip_range(wbxh_engine* h, int track, int clip, double* min_beat, double* max_beat) {
  wbx::AudioClip* c = clip_at(h, track, clip);
  if (!c || !min_beat || !max_beat) return WBX_ERR_INVALID;
  *min_beat = c->min_time;
  *max_beat = c->max_time;
  return WBX_OK;
}
int wbxh_move_clip(wbxh_engine* h, int track, int clip, double relative_pos) {
  wbx::AudioClip* c = clip_at(h, track, clip);
  return c ? h->eng.move_clip(h->eng.tracks[track], c, relative_pos) : WBX_ERR_INVALID;
}
int wbxh_resize_clip(wbxh_engine* h, int track, int clip, double relative_pos, double resize_limit, double min_length,
                     int left_side, int shift, int stretch) {
  wbx::AudioClip* c = clip_at(h, track, clip);
  return c ? h->eng.resize_clip(h->eng.tracks[track], c, relative_pos, resize_limit, min_length, left_side != 0, shift != 0,
                                stretch != 0)
           : WBX_ERR_INVALID;
}
int wbxh_delete_clip(wbxh_engine* h, int track, int clip) {
  wbx::AudioClip* c = clip_at(h, track, clip);
  return c ? h->eng.delete_clip(h->eng.tracks[track], c) : WBX_ERR_INVALID;
}
int wbxh_duplicate_clip(wbxh_engine* h, int track, int clip, double min_beat, double max_beat) {
  wbx::AudioClip* c = clip_at(h, track, clip);
  return c ? h->eng.duplicate_clip(h->eng.tracks[track], c, min_beat, max_beat) : WBX_ERR_INVALID;
}

int wbxh_delete_region(wbxh_engine* h, int track, double min_beat, double max_beat) {
  if (track < 0 || (size_t)track >= h->eng.tracks.size()) return WBX_ERR_INVALID;
  return h->eng.delete_region(h->eng.tracks[track], min_beat, max_beat);
}

int wbxh_set_effects(wbxh_engine* h, int track, const wbx_effect_params* params) {
  if (track < 0 || (size_t)track >= h->eng.tracks.size()) return WBX_ERR_INVALID;
  return h->eng.set_track_effects(h->eng.tracks[track], params);
}
int wbxh_set_impulse_response(wbxh_engine* h, const float* ir, uint32_t n_taps) {
  return h->eng.set_impulse_response(ir, n_taps);
}
void wbxh_set_resampler(wbxh_engine* h, int mode) { h->eng.resampler_mode = mode; }
void wbxh_set_bpm(wbxh_engine* h, double bpm) { h->eng.set_bpm(bpm); }
void wbxh_set_playhead(wbxh_engine* h, double beat) { h->eng.set_playhead_position(beat); }
void wbxh_play(wbxh_engine* h) { h->eng.play(); }
void wbxh_stop(wbxh_engine* h) { h->eng.stop(); }
void wbxh_set_fast_forward(wbxh_engine* h, int on) { h->eng.fast_forward = on != 0; }

namespace {
struct MemorySink final : Engine::BounceSink {
  uint8_t* dst;
  size_t cap, used = 0;
  MemorySink(void* d, size_t c) : dst((uint8_t*)d), cap(c) {}
  int write(const void* data, size_t bytes) override {
    if (used + bytes > cap) return WBX_ERR_INVALID;
    memcpy(dst + used, data, bytes);
    used += bytes;
    return WBX_OK;
  }
};
// Minimal RIFF/WAVE writer: PCM 16 / 24 / 32 bit or IEEE float 32. 24-bit files are fed I24_X8 words (the 24-bit value in
// the low three bytes, core/audio_format_conv.cpp:45-59) and packed to three bytes here.
struct WavFileSink final : Engine::BounceSink {
  FILE* f = nullptr;
  int fmt;
  uint32_t channels, rate;
  uint64_t data_bytes = 0;
  std::vector<uint8_t> pack;
  WavFileSink(const char* path, int format, uint32_t ch, uint32_t sr) : fmt(format), channels(ch), rate(sr) {
    f = fopen(path, "wb");
    if (f) header();
  }
  ~WavFileSink() override { close(); }
  void header() {
    const uint32_t bits = fmt == WBX_FMT_I16 ? 16 : (fmt == WBX_FMT_I24_X8 ? 24 : 32);
    const uint16_t tag = fmt == WBX_FMT_F32 ? 3 : 1, ch = (uint16_t)channels, align = (uint16_t)(channels * bits / 8), bps = (uint16_t)bits;
    const uint32_t byte_rate = rate * align, fmt_len = 16, data_len = (uint32_t)data_bytes, riff_len = 36 + data_len;
    fseek(f, 0, SEEK_SET);
    fwrite("RIFF", 1, 4, f), fwrite(&riff_len, 4, 1, f), fwrite("WAVEfmt ", 1, 8, f), fwrite(&fmt_len, 4, 1, f);
    fwrite(&tag, 2, 1, f), fwrite(&ch, 2, 1, f), fwrite(&rate, 4, 1, f), fwrite(&byte_rate, 4, 1, f), fwrite(&align, 2, 1, f);
    fwrite(&bps, 2, 1, f), fwrite("data", 1, 4, f), fwrite(&data_len, 4, 1, f);
  }
  int write(const void* data, size_t bytes) override {
    if (!f) return WBX_ERR_INVALID;
    if (fmt == WBX_FMT_I24_X8) {
      const size_t n = bytes / 4;
      pack.resize(n * 3);
      const uint8_t* s = (const uint8_t*)data;
      for (size_t i = 0; i < n; i++) pack[3 * i] = s[4 * i], pack[3 * i + 1] = s[4 * i + 1], pack[3 * i + 2] = s[4 * i + 2];
      data = pack.data();
      bytes = n * 3;
    }
    if (fwrite(data, 1, bytes, f) != bytes) return WBX_ERR_INVALID;
    data_bytes += bytes;
    return WBX_OK;
  }
  void close() {
    if (!f) return;
    header();  // sizes are known now
    fclose(f);
    f = nullptr;
  }
};
}  // namespace

int wbxh_bounce(wbxh_engine* h, double start_beat, double end_beat, int dst_format, uint32_t chunk_blocks, void* dst,
                uint64_t cap_bytes, uint64_t* frames_out) {
  if (!h || !dst) return WBX_ERR_INVALID;
  MemorySink sink(dst, (size_t)cap_bytes);
  return h->eng.bounce(start_beat, end_beat, dst_format, sink, chunk_blocks ? chunk_blocks : 256, frames_out);
}
int wbxh_bounce_wav(wbxh_engine* h, double start_beat, double end_beat, int dst_format, uint32_t chunk_blocks, const char* path,
                    uint64_t* frames_out) {
  if (!h || !path) return WBX_ERR_INVALID;
  WavFileSink sink(path, dst_format, h->eng.out_channels(), h->eng.sample_rate());
  if (!sink.f) return WBX_ERR_INVALID;
  return h->eng.bounce(start_beat, end_beat, dst_format, sink, chunk_blocks ? chunk_blocks : 256, frames_out);
}

int wbxh_render_begin(wbxh_engine* h, uint32_t n_blocks) { return h->eng.render_begin(n_blocks); }
int wbxh_render_end(wbxh_engine* h, float* const* out_channels, float* peaks) { return h->eng.render_end(out_channels, peaks); }

int wbxh_render(wbxh_engine* h, uint32_t n_blocks, float* const* out_channels, float* peaks) {
  return h->eng.render(n_blocks, out_channels, peaks);
}

// Scheduling only (no device work): fills the engine's segment table for n_blocks callbacks.
int wbxh_schedule(wbxh_engine* h, uint32_t n_blocks, const wbx_segment** segs, uint32_t* n_segs, const float** gains) {
  int rc = h->eng.schedule(n_blocks);
  if (segs) *segs = h->eng.segments().data();
  if (n_segs) *n_segs = (uint32_t)h->eng.segments().size();
  if (gains) *gains = h->eng.track_gains().data();
  return rc;
}

double wbxh_sampler_offset(wbxh_engine* h, int track) {
  wbx::Track* t = track_at(h, track);
  return t ? t->sample_offset : 0.0;
}
double wbxh_sample_position(wbxh_engine* h) { return h->eng.sample_position; }
double wbxh_playhead(wbxh_engine* h) { return h->eng.playhead; }
float wbxh_level(wbxh_engine* h, int track, int channel, int reset) {
  wbx::Track* t = track_at(h, track);
  if (!t || channel < 0 || channel > 1) return 0.0f;
  return reset ? t->level[channel].take() : t->level[channel].peek();  // VUMeter::update's exchange(0) (vu_meter.h:32-40)
}
uint32_t wbxh_advance_rounded(double* off, double adv, uint32_t n, double limit) {
  return wbx::advance_rounded(off, adv, n, limit);
}

void wbxh_panning_coefs(float pan, float* left, float* right) {
  wbx::PanningCoefficient c = wbx::calculate_panning_coefs(pan);
  *left = c.left;
  *right = c.right;
}
float wbxh_db_to_linear(float db) { return wbx::db_to_linear(db); }

}  // extern "C"
