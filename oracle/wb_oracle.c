/* TEST INFRASTRUCTURE ONLY — never linked into, imported by or executed from the product path.
 *
 * wb_oracle.c — plain-C restatement ("port") of whitebox's mixing hot path, exporting the wbo.h scenario
 * API. Each function cites the reference file:line (relative to /root/reference) it follows.
 *
 * Parity status: PINNED. This file is checked bit-for-bit against the reference's own Engine::process
 * (oracle/_ref/libwbref.so, compiled unmodified from /root/reference/src) by tests/test_oracle.py on seeded
 * scenarios whenever that library is present, and against the committed vectors under tests/golden/ that
 * libwbref.so produced (tests/golden/make_golden.py) everywhere else. The reference's own tests hold no
 * golden vectors for this path.
 *
 * Build: gcc -std=c11 -O3 -ffp-contract=off (no -march, no -ffast-math) so every f32/f64 operation is
 * separately rounded exactly as in the reference's ISO-C++ x86-64 build.
 *
 * Scope notes (documented deviations, none reachable through wbo.h calls used by the tests):
 *  - a clip removed by a later overlapping clip (Engine::reserve_track_region, engine.cpp:478-569) is parked in a
 *    per-track graveyard instead of being destroyed at once (the reference destroys it in update_clip_ordering while a
 *    playing voice may still point at it).
 *  - the per-track message ring (capacity 64, track.cpp:23) is an unbounded list here.
 *  - MIDI clips, plugins and recording are not restated.
 */
#define _POSIX_C_SOURCE 199309L
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "wbo.h"

#define SAMPLE_PADDING 16 /* Sample::sample_padding, dsp/sample.h:19 */

typedef struct {
  int fmt;
  uint32_t channels, rate;
  size_t count;
  void** data;
} o_sample;

typedef struct {
  double min_time, max_time, start_offset; /* clip.h:68-70 */
  double speed;                            /* AudioClip::speed clip.h:43 */
  float gain;                              /* AudioClip::gain  clip.h:44 */
  double fade_start, fade_end;             /* AudioClip::fade_start/fade_end clip.h:41-42 (EXTENSION, see fade_env) */
  o_sample* sample;
  int internal_state_changed; /* clip.h:62 — only UI edits set it */
  int deleted;                /* Clip::mark_deleted */
} o_clip;

enum { EV_NONE, EV_STOP, EV_PLAY }; /* EventType, event.h:11-15 */

typedef struct { /* AudioEvent, event.h:66-74 */
  int type;
  uint32_t buffer_offset;
  double time, speed;
  size_t sample_offset;
  uint64_t clip_frame; /* EXTENSION: output frames since the clip's start when the event fires */
  o_clip* clip;
  o_sample* sample;
} o_event;

typedef struct {
  uint32_t id; /* TrackParameter, track.h:29-34: 0 volume, 1 pan, 2 mute */
  double value;
} o_msg;

/* EXTENSION: designed effect chain + its running state (see apply_effects) */
typedef struct {
  int eq_on, comp_on;
  float b0[4], b1[4], b2[4], a1[4], a2[4];
  float thr, att, rel, makeup;
  int ratio_code;
  float s1[2][4], s2[2][4], env[2];
  int reverb_on;
  float* hist[2]; /* last n_taps - 1 chain outputs per channel (ring), zero at attach time */
  uint64_t hist_pos;
  /* time-parallel form (fx_design_tables): per biquad the state-space input vector, A^m (m = 0..16) and A^(16*2^j)
   * (j = 0..4), row-major 2x2; envelope follower constants and look-ahead slopes */
  float B1[4], B2[4], P[4][17][4], S[4][5][4];
  float a1m, r1m, sl[14]; /* sl: S1[2] | S2[3] | S3[4] | S4[5] */
  int sel;                /* att <= rel: the branch of the follower is the larger candidate, else the smaller */
} o_fx;

typedef struct {
  o_clip** clips;
  uint32_t n_clips, cap_clips;
  int ui_solo;        /* TrackParameterState::solo of ui_parameter_state (track.h:52, UI only) */
  int has_plugin;     /* Track::plugin_instance != nullptr (track.h:124), for a plugin that writes nothing */
  o_clip** graveyard; /* clips trimmed away by later edits */
  uint32_t n_grave, cap_grave;
  o_fx fx;
  /* TrackEventState, track.h:36-44 */
  int has_clip_idx;
  uint32_t clip_idx;
  int refresh_voice, partially_ended;
  o_event* events;
  uint32_t n_events, cap_events;
  o_event current; /* Track::current_audio_event, track.h:118 */
  /* dsp::Sampler, sampler.h:14-16 */
  double playback_speed, sample_offset;
  uint64_t clip_frame; /* EXTENSION: output frames since the playing clip's start */
  /* parameter_state, track.h:46-53 */
  float volume, pan, pan_coeffs[2];
  int mute;
  o_msg* msgs;
  uint32_t n_msgs, cap_msgs;
  float level[2]; /* VUMeter::level, vu_meter.h:17 */
} o_track;

struct wbo_session {
  uint32_t C, B, rate;
  double beat_duration, ppq;
  double playhead, playhead_start, sample_position;
  int playing;
  o_track** tracks;
  uint32_t n_tracks, cap_tracks;
  o_sample** samples;
  uint32_t n_samples, cap_samples;
  float** mixing; /* Engine::mixing_buffer, engine.h:57 */
  float** out;
  float* ir; /* EXTENSION: session impulse response */
  uint32_t ir_taps;
  int resampler;          /* EXTENSION: 0 linear (reference), 1 polyphase */
  float poly[128 * 16];   /* polyphase coefficient table */
};

const char* wbo_kind(void) { return "port"; }
static void fx_free(o_fx* f);
static void fx_design_tables(o_fx* f);
static void fx64_forget(const o_fx* f);

static uint64_t g_ub_count = 0;
/* port only: how many times a scenario drove the reference algorithm into undefined behaviour */
uint64_t wbo_ub_count(void) { return g_ub_count; }

/* ---- scalar math ------------------------------------------------------------------------------------- */

/* core/core_math.h:209-212 */
static double beat_to_samples(double beat, double sample_rate, double beat_duration) {
  double sec = beat * beat_duration;
  return sec * sample_rate;
}

/* core/core_math.h:83-89: db <= -72 -> 0, else powf(10, (float)((double)db * 0.05)) */
float wbo_db_to_linear(float db) {
  if (db <= -72.0f) return 0.0f;
  return powf(10.0f, (float)((double)db * 0.05));
}

/* core/panning_law.cpp:9-32, ConstantPower_3db branch */
void wbo_panning_coefs(float p, float* left, float* right) {
  const double pi = 3.141592653589793238462643383279502884; /* std::numbers::pi */
  double x = 0.5 * ((double)p + 1.0);
  double l = sin(0.5 * pi * (1.0 - x));
  double r = sin(0.5 * pi * x);
  double boost = sqrt(2.0);
  *left = (float)(l * boost);
  *right = (float)(r * boost);
}

/* ---- session / editing API ------------------------------------------------------------------------- */

static float** alloc_planar(uint32_t C, uint32_t B) {
  float** p = (float**)calloc(C ? C : 1, sizeof(float*));
  for (uint32_t c = 0; c < C; c++) p[c] = (float*)calloc(B ? B : 1, sizeof(float));
  return p;
}

/* Engine::set_audio_channel_config (engine.cpp:43-57) + Engine::set_bpm (engine.cpp:24-30) */
wbo_session* wbo_create(uint32_t out_channels, uint32_t block_frames, uint32_t sample_rate, double bpm) {
  wbo_session* s = (wbo_session*)calloc(1, sizeof(*s));
  s->C = out_channels;
  s->B = block_frames;
  s->rate = sample_rate;
  s->beat_duration = 60.0 / bpm;
  s->ppq = 96.0; /* engine.h:43 */
  s->mixing = alloc_planar(out_channels, block_frames);
  s->out = alloc_planar(out_channels, block_frames);
  return s;
}

/* Engine::set_audio_channel_config again (engine.cpp:43-57): only the buffer shapes and the numbers process() is called
 * with change; tracks, clips, transport and the running voices' sampler state persist. */
int wbo_reconfigure(wbo_session* s, uint32_t out_channels, uint32_t block_frames, uint32_t sample_rate) {
  for (uint32_t c = 0; c < s->C; c++) {
    free(s->mixing[c]);
    free(s->out[c]);
  }
  free(s->mixing);
  free(s->out);
  s->C = out_channels;
  s->B = block_frames;
  s->rate = sample_rate;
  s->mixing = alloc_planar(out_channels, block_frames);
  s->out = alloc_planar(out_channels, block_frames);
  return 0;
}

static void track_free(o_track* tr) {
  for (uint32_t i = 0; i < tr->n_clips; i++) free(tr->clips[i]);
  free(tr->clips);
  for (uint32_t i = 0; i < tr->n_grave; i++) free(tr->graveyard[i]);
  free(tr->graveyard);
  fx_free(&tr->fx);
  free(tr->events);
  free(tr->msgs);
  free(tr);
}

void wbo_destroy(wbo_session* s) {
  if (!s) return;
  for (uint32_t t = 0; t < s->n_tracks; t++) {
    o_track* tr = s->tracks[t];
    track_free(tr);
  }
  free(s->tracks);
  for (uint32_t i = 0; i < s->n_samples; i++) {
    for (uint32_t c = 0; c < s->samples[i]->channels; c++) free(s->samples[i]->data[c]);
    free(s->samples[i]->data);
    free(s->samples[i]);
  }
  free(s->samples);
  for (uint32_t c = 0; c < s->C; c++) {
    free(s->mixing[c]);
    free(s->out[c]);
  }
  free(s->mixing);
  free(s->out);
  free(s->ir);
  free(s);
}

static void push_msg(o_track* tr, uint32_t id, double value) {
  if (tr->n_msgs == tr->cap_msgs) {
    tr->cap_msgs = tr->cap_msgs ? tr->cap_msgs * 2 : 8;
    tr->msgs = (o_msg*)realloc(tr->msgs, tr->cap_msgs * sizeof(o_msg));
  }
  tr->msgs[tr->n_msgs].id = id;
  tr->msgs[tr->n_msgs].value = value;
  tr->n_msgs++;
}

/* Track::set_volume / set_pan / set_mute (track.cpp:47-79): the UI thread converts dB -> linear in f32
 * and queues the value as a double. */
void wbo_set_volume(wbo_session* s, int t, float db) { push_msg(s->tracks[t], 0, (double)wbo_db_to_linear(db)); }
void wbo_set_pan(wbo_session* s, int t, float pan) { push_msg(s->tracks[t], 1, (double)pan); }
void wbo_set_mute(wbo_session* s, int t, int mute) { push_msg(s->tracks[t], 2, (double)(mute != 0)); }

/* Engine::add_track (engine.cpp:199-207) + Track::Track() (track.cpp:22-27), which queues volume 0 dB,
 * pan 0, mute off; the harness then applies the caller's values the same way. */
int wbo_add_track(wbo_session* s, float volume_db, float pan, int mute) {
  o_track* tr = (o_track*)calloc(1, sizeof(*tr));
  if (s->n_tracks == s->cap_tracks) {
    s->cap_tracks = s->cap_tracks ? s->cap_tracks * 2 : 16;
    s->tracks = (o_track**)realloc(s->tracks, s->cap_tracks * sizeof(o_track*));
  }
  s->tracks[s->n_tracks++] = tr;
  int t = (int)s->n_tracks - 1;
  wbo_set_volume(s, t, 0.0f);
  wbo_set_pan(s, t, 0.0f);
  wbo_set_mute(s, t, 0);
  wbo_set_volume(s, t, volume_db);
  wbo_set_pan(s, t, pan);
  wbo_set_mute(s, t, mute);
  return t;
}

static size_t fmt_size(int fmt) { return fmt == WBO_FMT_I16 ? 2 : 4; }

/* Sample layout (dsp/sample.h:18-28): one array per channel, `count` frames + 16 zero frames of padding
 * (dsp/sample.cpp:127,140). */
int wbo_add_sample(wbo_session* s, int format, uint32_t channels, uint64_t frames, uint32_t sample_rate,
                   const void* const* planar) {
  o_sample* sm = (o_sample*)calloc(1, sizeof(*sm));
  sm->fmt = format;
  sm->channels = channels;
  sm->rate = sample_rate;
  sm->count = (size_t)frames;
  sm->data = (void**)calloc(channels, sizeof(void*));
  for (uint32_t c = 0; c < channels; c++) {
    sm->data[c] = calloc(frames + SAMPLE_PADDING, fmt_size(format));
    memcpy(sm->data[c], planar[c], frames * fmt_size(format));
  }
  if (s->n_samples == s->cap_samples) {
    s->cap_samples = s->cap_samples ? s->cap_samples * 2 : 16;
    s->samples = (o_sample**)realloc(s->samples, s->cap_samples * sizeof(o_sample*));
  }
  s->samples[s->n_samples++] = sm;
  return (int)s->n_samples - 1;
}

/* wb::find_lower_bound (core/algorithm.h:26-42) with the comparator of Track::find_next_clip
 * (track.cpp:207): note `right` starts at n-1, so the result is never one-past-the-end. */
static uint32_t lower_bound_max_time(o_clip** clips, uint32_t n, double time_pos) {
  int64_t left = 0, right = (int64_t)n - 1;
  while (left < right) {
    int64_t middle = (left + right) >> 1;
    if (clips[middle]->max_time <= time_pos)
      left = middle + 1;
    else
      right = middle;
  }
  return (uint32_t)right;
}

/* Track::find_next_clip (track.cpp:182-213). Returns 1 and *idx when a clip is found. */
static int find_next_clip(o_track* tr, double time_pos, uint32_t* idx) {
  if (tr->n_clips == 0) return 0;
  if (tr->clips[tr->n_clips - 1]->max_time < time_pos) return 0;
  *idx = lower_bound_max_time(tr->clips, tr->n_clips, time_pos); /* ids == positions after ordering */
  return 1;
}

/* Track::reset_playback_state (track.cpp:220-232) */
static void reset_playback_state(o_track* tr, double time_pos, int refresh_voices) {
  if (!refresh_voices) {
    uint32_t idx = 0;
    tr->has_clip_idx = find_next_clip(tr, time_pos, &idx);
    tr->clip_idx = idx;
    tr->partially_ended = 0;
  }
  tr->refresh_voice = refresh_voices;
}

/* Engine::add_audio_clip (engine.cpp:293-309) + add_to_cliplist (engine.cpp:409-461) for the
 * non-overlapping cases: append / prepend / insert + sort by min_time, then
 * reset_playback_state(playhead, true). */
int wbo_add_clip(wbo_session* s, int track, int sample, double min_beat, double max_beat, double start_offset,
                 double speed, float gain) {
  return wbo_add_clip_fade(s, track, sample, min_beat, max_beat, start_offset, speed, gain, 0.0, 0.0);
}

/* core/core_math.h:204-207 */
static double samples_to_beat(double samples, double sample_rate, double beat_duration) {
  double sec = samples / sample_rate;
  return sec / beat_duration;
}

/* shift_clip_content + calc_clip_shift, audio clips (engine/clip_edit.h:128-150); sample_rate = the asset's */
static double shift_clip_content(const o_clip* clip, double relative_pos, double beat_duration) {
  double sample_rate = (double)clip->sample->rate;
  relative_pos *= clip->speed;
  double offset_in_beat = samples_to_beat(clip->start_offset, sample_rate, beat_duration);
  double shifted = offset_in_beat - relative_pos;
  return beat_to_samples(shifted > 0.0 ? shifted : 0.0, sample_rate, beat_duration);
}

static void clips_push(o_track* tr, o_clip* c) {
  if (tr->n_clips == tr->cap_clips) {
    tr->cap_clips = tr->cap_clips ? tr->cap_clips * 2 : 4;
    tr->clips = (o_clip**)realloc(tr->clips, tr->cap_clips * sizeof(o_clip*));
  }
  tr->clips[tr->n_clips++] = c;
}

/* Track::query_clip_by_range (track.cpp:112-157) with wb::find_lower_bound (core/algorithm.h:25-40) */
static int query_clip_by_range(const o_track* tr, double min, double max, uint32_t* first_out, uint32_t* last_out) {
  if (tr->n_clips == 0) return 0;
  if (max <= tr->clips[0]->min_time) return 0;
  if (min >= tr->clips[tr->n_clips - 1]->max_time) return 0;
  uint32_t first_clip = lower_bound_max_time(tr->clips, tr->n_clips, min);
  uint32_t last_clip = lower_bound_max_time(tr->clips, tr->n_clips, max);
  const o_clip* first = tr->clips[first_clip];
  const o_clip* last = tr->clips[last_clip];
  if (first_clip == last_clip && (max <= first->min_time || min >= last->max_time)) return 0;
  if (min > first->max_time) first_clip++;
  if (!(max > last->min_time)) last_clip--;
  *first_out = first_clip;
  *last_out = last_clip;
  return 1;
}

static int clip_min_time_less(const void* a, const void* b) {
  const o_clip* x = *(o_clip* const*)a;
  const o_clip* y = *(o_clip* const*)b;
  return (x->min_time > y->min_time) - (x->min_time < y->min_time);
}

/* Track::update_clip_ordering (track.cpp:159-180): drop deleted clips, sort by min_time */
static void update_clip_ordering(o_track* tr) {
  uint32_t n = 0;
  for (uint32_t i = 0; i < tr->n_clips; i++) {
    o_clip* c = tr->clips[i];
    if (c->deleted) {
      if (tr->n_grave == tr->cap_grave) {
        tr->cap_grave = tr->cap_grave ? tr->cap_grave * 2 : 4;
        tr->graveyard = (o_clip**)realloc(tr->graveyard, tr->cap_grave * sizeof(o_clip*));
      }
      tr->graveyard[tr->n_grave++] = c;
    } else {
      tr->clips[n++] = c;
    }
  }
  tr->n_clips = n;
  qsort(tr->clips, tr->n_clips, sizeof(o_clip*), clip_min_time_less);
}

/* Track::mark_clip_deleted. If a voice is playing this very clip, the reference goes on reading
 * current_audio_event.clip->audio.gain (track.cpp:676,716) after update_clip_ordering destroyed the clip and its pool
 * slot was handed to the next new clip: a use-after-free with no defined answer. Counted like the other UB case so
 * that comparisons against the reference skip such sessions; port and product keep the clip (and its gain) alive. */
static void mark_deleted(o_track* tr, o_clip* clip) {
  clip->deleted = 1;
  if (tr->current.type == EV_PLAY && tr->current.clip == clip) g_ub_count++;
}

/* Engine::reserve_track_region (engine.cpp:478-569), ignore_clip == nullptr */
static void reserve_track_region(wbo_session* s, o_track* tr, uint32_t first_clip, uint32_t last_clip, double min, double max,
                                 const o_clip* ignore_clip) {
  if (tr->n_clips == 0) return;
  double current_beat_duration = s->beat_duration;
  if (first_clip == last_clip) {
    o_clip* clip = tr->clips[first_clip];
    if (clip == ignore_clip) return;
    if (min > clip->min_time && max < clip->max_time) { /* split into two parts */
      o_clip* right = (o_clip*)malloc(sizeof(*right));
      *right = *clip; /* Clip(const Clip&), clip.h:92-112: `deleted` and `internal_state_changed` are NOT copied */
      right->deleted = 0;
      right->internal_state_changed = 0;
      right->min_time = max;
      right->start_offset = shift_clip_content(right, clip->min_time - max, current_beat_duration);
      clip->max_time = min;
      clips_push(tr, right);
    } else if (min > clip->min_time) {
      clip->max_time = min;
    } else if (max < clip->max_time) {
      clip->start_offset = shift_clip_content(clip, clip->min_time - max, current_beat_duration);
      clip->min_time = max;
    } else {
      mark_deleted(tr, clip);
    }
    return;
  }
  o_clip* first = tr->clips[first_clip];
  o_clip* last = tr->clips[last_clip];
  if (first != ignore_clip && min > first->min_time) {
    first->max_time = min;
    first_clip++;
  }
  if (last != ignore_clip && max < last->max_time) {
    last->start_offset = shift_clip_content(last, last->min_time - max, current_beat_duration);
    last->min_time = max;
    last_clip--;
  }
  if (first_clip <= last_clip && last_clip < tr->n_clips)
    for (uint32_t i = first_clip; i <= last_clip; i++)
      if (tr->clips[i] != ignore_clip) mark_deleted(tr, tr->clips[i]);
}

/* Engine::add_audio_clip (engine.cpp:293-309) + add_to_cliplist (:409-461) */
int wbo_add_clip_fade(wbo_session* s, int track, int sample, double min_beat, double max_beat, double start_offset,
                      double speed, float gain, double fade_start, double fade_end) {
  o_track* tr = s->tracks[track];
  o_clip* c = (o_clip*)calloc(1, sizeof(*c));
  c->min_time = min_beat;
  c->max_time = max_beat;
  c->start_offset = start_offset;
  c->speed = speed;
  c->gain = gain;
  c->fade_start = fade_start;
  c->fade_end = fade_end;
  c->sample = s->samples[sample];
  if (tr->n_clips == 0 || tr->clips[tr->n_clips - 1]->max_time < c->min_time) { /* first clip / add to the back */
    clips_push(tr, c);
  } else if (tr->clips[0]->min_time > c->max_time) { /* add to the front */
    clips_push(tr, c);
    for (uint32_t i = tr->n_clips - 1; i > 0; i--) tr->clips[i] = tr->clips[i - 1];
    tr->clips[0] = c;
  } else {
    uint32_t first = 0, last = 0;
    if (query_clip_by_range(tr, c->min_time, c->max_time, &first, &last))
      reserve_track_region(s, tr, first, last, c->min_time, c->max_time, NULL);
    clips_push(tr, c);
    update_clip_ordering(tr);
  }
  reset_playback_state(tr, s->playhead, 1);
  return 0;
}

/* Engine::add_to_cliplist (engine.cpp:409-461) for an already constructed clip */
static void add_to_cliplist(wbo_session* s, o_track* tr, o_clip* c) {
  if (tr->n_clips == 0 || tr->clips[tr->n_clips - 1]->max_time < c->min_time) {
    clips_push(tr, c);
  } else if (tr->clips[0]->min_time > c->max_time) {
    clips_push(tr, c);
    for (uint32_t i = tr->n_clips - 1; i > 0; i--) tr->clips[i] = tr->clips[i - 1];
    tr->clips[0] = c;
  } else {
    uint32_t first = 0, last = 0;
    if (query_clip_by_range(tr, c->min_time, c->max_time, &first, &last))
      reserve_track_region(s, tr, first, last, c->min_time, c->max_time, NULL);
    clips_push(tr, c);
    update_clip_ordering(tr);
  }
  reset_playback_state(tr, s->playhead, 1);
}

void wbo_set_bpm(wbo_session* s, double bpm) { s->beat_duration = 60.0 / bpm; } /* engine.cpp:24-30 */

int wbo_set_plugin(wbo_session* s, int track, int present) { /* engine.cpp:1466-1551 */
  s->tracks[track]->has_plugin = present != 0;
  return 0;
}

/* Engine::set_clip_gain (engine.cpp:1460-1464) */
int wbo_set_clip_gain(wbo_session* s, int track, int clip, float gain) {
  o_track* tr = s->tracks[track];
  if (clip < 0 || (uint32_t)clip >= tr->n_clips) return -1;
  tr->clips[clip]->gain = gain;
  return 0;
}

/* Engine::solo_track (engine.cpp:245-262) */
void wbo_solo_track(wbo_session* s, int slot) {
  int mute = 0;
  if (s->tracks[slot]->ui_solo) {
    s->tracks[slot]->ui_solo = 0;
  } else {
    s->tracks[slot]->ui_solo = 1;
    wbo_set_mute(s, slot, 0);
    mute = 1;
  }
  for (uint32_t i = 0; i < s->n_tracks; i++) {
    if ((int)i == slot) continue;
    if (s->tracks[i]->ui_solo) s->tracks[i]->ui_solo = 0;
    wbo_set_mute(s, (int)i, mute);
  }
}

/* Engine::move_track (engine.cpp:228-243) */
void wbo_move_track(wbo_session* s, int from_slot, int to_slot) {
  if (from_slot == to_slot) return;
  o_track* tmp = s->tracks[from_slot];
  if (from_slot < to_slot)
    for (int i = from_slot; i < to_slot; i++) s->tracks[i] = s->tracks[i + 1];
  else
    for (int i = from_slot; i > to_slot; i--) s->tracks[i] = s->tracks[i - 1];
  s->tracks[to_slot] = tmp;
}

/* Engine::delete_track (engine.cpp:209-217) */
void wbo_delete_track(wbo_session* s, int slot) {
  o_track* tr = s->tracks[slot];
  for (uint32_t i = (uint32_t)slot; i + 1 < s->n_tracks; i++) s->tracks[i] = s->tracks[i + 1];
  s->n_tracks--;
  track_free(tr);
}

int wbo_clip_count(wbo_session* s, int track) { return (int)s->tracks[track]->n_clips; }

static o_clip* clip_at(wbo_session* s, int track, int clip) {
  o_track* tr = s->tracks[track];
  return (clip < 0 || (uint32_t)clip >= tr->n_clips) ? NULL : tr->clips[clip];
}

int wbo_clip_range(wbo_session* s, int track, int clip, double* min_beat, double* max_beat) {
  o_clip* c = clip_at(s, track, clip);
  if (!c) return -1;
  *min_beat = c->min_time;
  *max_beat = c->max_time;
  return 0;
}

/* Engine::move_clip (engine.cpp:346-363) + calc_move_clip (clip_edit.h:10-16) */
int wbo_move_clip(wbo_session* s, int track, int clip, double relative_pos) {
  o_track* tr = s->tracks[track];
  o_clip* c = clip_at(s, track, clip);
  if (!c) return -1;
  if (relative_pos == 0.0) return 0;
  double new_pos = c->min_time + relative_pos;
  if (!(new_pos > 0.0)) new_pos = 0.0; /* math::max(.., min_move = 0.0) */
  double min_time = new_pos, max_time = new_pos + (c->max_time - c->min_time);
  uint32_t first = 0, last = 0;
  if (query_clip_by_range(tr, min_time, max_time, &first, &last))
    reserve_track_region(s, tr, first, last, min_time, max_time, c);
  c->min_time = min_time;
  c->max_time = max_time;
  c->internal_state_changed = 1;
  update_clip_ordering(tr);
  reset_playback_state(tr, s->playhead, 1);
  return 0;
}

/* Engine::resize_clip (engine.cpp:365-398) + calc_resize_clip (clip_edit.h:18-126), audio clips,
 * clamp_at_resize_pos = false */
int wbo_resize_clip(wbo_session* s, int track, int clip, double relative_pos, double resize_limit, double min_length,
                    int left_side, int shift, int stretch) {
  o_track* tr = s->tracks[track];
  o_clip* c = clip_at(s, track, clip);
  if (!c) return -1;
  if (relative_pos == 0.0) return 0;
  const double beat_duration = s->beat_duration;
  const double asset_rate = (double)c->sample->rate;
  const double sample_count = (double)c->sample->count;
  double min_time, max_time, start_offset = c->start_offset, new_speed = 1.0;
  if (!left_side) {
    const double old_max = c->max_time;
    const double actual_min_length = resize_limit + min_length - c->min_time;
    double new_max = c->max_time + relative_pos;
    if (!(new_max > 0.0)) new_max = 0.0;
    double length = new_max - c->min_time;
    if (length < actual_min_length) new_max = c->min_time + actual_min_length;
    if (shift) {
      double mult = c->speed;
      start_offset = samples_to_beat(start_offset, asset_rate, beat_duration);
      if (old_max < new_max)
        start_offset -= (new_max - old_max) * mult;
      else
        start_offset += (old_max - new_max) * mult;
      if (!(start_offset > 0.0)) start_offset = 0.0;
      if (!(start_offset < sample_count)) start_offset = sample_count; /* math::min(start_offset, count) */
      start_offset = beat_to_samples(start_offset, asset_rate, beat_duration);
    }
    if (stretch) {
      double old_length = sample_count / c->speed;
      double num_samples = beat_to_samples(relative_pos, asset_rate, beat_duration);
      new_speed = sample_count / (old_length + num_samples);
    }
    min_time = c->min_time;
    max_time = new_max;
  } else {
    const double old_min = c->min_time;
    const double actual_min_length = c->max_time - resize_limit + min_length;
    double new_min = c->min_time + relative_pos;
    if (!(new_min > 0.0)) new_min = 0.0;
    double length = c->max_time - new_min;
    if (length < actual_min_length) new_min = c->max_time - actual_min_length;
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
      double old_length = sample_count / c->speed;
      double num_samples = beat_to_samples(old_min - new_min, asset_rate, beat_duration);
      new_speed = sample_count / (old_length + num_samples);
    }
    min_time = new_min;
    max_time = c->max_time;
  }
  uint32_t first = 0, last = 0;
  if (query_clip_by_range(tr, min_time, max_time, &first, &last))
    reserve_track_region(s, tr, first, last, min_time, max_time, c);
  if (left_side)
    c->min_time = min_time;
  else
    c->max_time = max_time;
  c->start_offset = start_offset;
  if (stretch) c->speed = new_speed;
  c->internal_state_changed = (shift || stretch) ? 1 : 0;
  update_clip_ordering(tr);
  reset_playback_state(tr, s->playhead, 1);
  return 0;
}

/* Engine::delete_clip (engine.cpp:400-407) */
int wbo_delete_clip(wbo_session* s, int track, int clip) {
  o_track* tr = s->tracks[track];
  o_clip* c = clip_at(s, track, clip);
  if (!c) return -1;
  mark_deleted(tr, c);
  update_clip_ordering(tr);
  reset_playback_state(tr, s->playhead, 1);
  return 0;
}

/* Engine::delete_region(track, min, max) (engine.cpp:463-473) */
int wbo_delete_region(wbo_session* s, int track, double min_beat, double max_beat) {
  o_track* tr = s->tracks[track];
  uint32_t first = 0, last = 0;
  if (!query_clip_by_range(tr, min_beat, max_beat, &first, &last)) return 0;
  reserve_track_region(s, tr, first, last, min_beat, max_beat, NULL);
  update_clip_ordering(tr);
  reset_playback_state(tr, s->playhead, 1);
  return 0;
}

/* Engine::duplicate_clip (engine.cpp:336-344) */
int wbo_duplicate_clip(wbo_session* s, int track, int clip, double min_beat, double max_beat) {
  o_track* tr = s->tracks[track];
  o_clip* src = clip_at(s, track, clip);
  if (!src) return -1;
  o_clip* c = (o_clip*)malloc(sizeof(*c));
  *c = *src; /* Clip(const Clip&): flags start cleared, see reserve_track_region */
  c->deleted = 0;
  c->internal_state_changed = 0;
  c->min_time = min_beat;
  c->max_time = max_beat;
  add_to_cliplist(s, tr, c);
  return 0;
}

/* Engine::set_playhead_position (engine.cpp:32-41) */
void wbo_set_playhead(wbo_session* s, double beat) {
  s->playhead_start = beat;
  s->playhead = beat;
}

/* Engine::play (engine.cpp:68-80) */
void wbo_play(wbo_session* s) {
  for (uint32_t t = 0; t < s->n_tracks; t++) reset_playback_state(s->tracks[t], s->playhead_start, 0);
  s->sample_position = 0;
  s->playing = 1;
}

/* Engine::stop (engine.cpp:82-92) + Track::stop (track.cpp:248-256) */
void wbo_stop(wbo_session* s) {
  s->playing = 0;
  s->playhead = s->playhead_start;
  for (uint32_t t = 0; t < s->n_tracks; t++) {
    memset(&s->tracks[t]->current, 0, sizeof(o_event));
    s->tracks[t]->n_events = 0;
  }
}

/* ---- event scheduling ------------------------------------------------------------------------------ */

static void push_event(o_track* tr, o_event e) {
  if (tr->n_events == tr->cap_events) {
    tr->cap_events = tr->cap_events ? tr->cap_events * 2 : 8;
    tr->events = (o_event*)realloc(tr->events, tr->cap_events * sizeof(o_event));
  }
  tr->events[tr->n_events++] = e;
}

static void push_stop(o_track* tr, uint32_t buffer_offset, double time) {
  o_event e;
  memset(&e, 0, sizeof(e));
  e.type = EV_STOP;
  e.buffer_offset = buffer_offset;
  e.time = time;
  push_event(tr, e);
}

static void push_play(o_track* tr, uint32_t buffer_offset, double time, o_clip* clip, size_t sample_offset,
                      uint64_t clip_frame) {
  o_event e;
  memset(&e, 0, sizeof(e));
  e.clip_frame = clip_frame;
  e.type = EV_PLAY;
  e.buffer_offset = buffer_offset;
  e.time = time;
  e.speed = clip->speed;
  e.sample_offset = sample_offset;
  e.clip = clip;
  e.sample = clip->sample;
  push_event(tr, e);
}

/* Track::process_event (track.cpp:258-451), audio clips only. */
static void process_event(o_track* tr, double start_time, double end_time, double sample_position,
                          double beat_duration, double sample_rate, uint32_t buffer_size) {
  if (tr->n_clips == 0) { /* :268-284 */
    if (tr->refresh_voice) {
      push_stop(tr, 0, start_time);
      tr->has_clip_idx = 0;
      tr->refresh_voice = 0;
    }
    return;
  }

  uint32_t num_clips = tr->n_clips;
  if (tr->refresh_voice) { /* :287-340 */
    uint32_t at = 0;
    if (find_next_clip(tr, start_time, &at)) {
      if (tr->has_clip_idx) {
        uint32_t idx = tr->clip_idx;
        if (idx < num_clips) {
          o_clip* clip = tr->clips[at];
          o_clip* current_clip = tr->clips[idx];
          if (clip != current_clip && start_time >= clip->min_time && start_time <= clip->max_time) {
            push_stop(tr, 0, start_time);
            tr->clip_idx = at;
            tr->partially_ended = 0;
          } else if (clip == current_clip && (start_time < clip->min_time || start_time > clip->max_time)) {
            push_stop(tr, 0, start_time);
            tr->clip_idx = at;
            tr->partially_ended = 0;
          }
        }
      } else {
        tr->has_clip_idx = 1;
        tr->clip_idx = at;
      }
    } else {
      push_stop(tr, 0, start_time);
      tr->has_clip_idx = 0;
    }
    tr->refresh_voice = 0;
  }

  if (!tr->has_clip_idx) return; /* :342-346 */

  uint32_t next_clip = tr->clip_idx;
  while (next_clip < num_clips) { /* :348-446 */
    o_clip* clip = tr->clips[next_clip];
    double min_time = clip->min_time;
    double max_time = clip->max_time;

    if (min_time > end_time) break;

    if (min_time >= start_time) { /* started from the beginning, :357-375 */
      double offset_from_start = beat_to_samples(min_time - start_time, sample_rate, beat_duration);
      double sample_offset = sample_position + offset_from_start;
      uint32_t buffer_offset = (uint32_t)((uint64_t)sample_offset % (uint64_t)buffer_size);
      push_play(tr, buffer_offset, min_time, clip, (size_t)clip->start_offset, 0);
      clip->internal_state_changed = 0;
    } else if (start_time > min_time && !tr->partially_ended) { /* started in the middle, :376-395 */
      double relative_start_time = start_time - min_time;
      double sample_pos = beat_to_samples(relative_start_time, sample_rate, beat_duration);
      size_t sample_offset = (size_t)(clip->start_offset + (sample_pos * clip->speed));
      push_play(tr, 0, start_time, clip, sample_offset, (uint64_t)sample_pos);
      clip->internal_state_changed = 0;
    } else if (clip->internal_state_changed && tr->partially_ended) { /* :396-421 */
      double relative_start_time = start_time - min_time;
      double sample_pos = beat_to_samples(relative_start_time, sample_rate, beat_duration);
      size_t sample_offset = (size_t)(clip->start_offset + (sample_pos * clip->speed));
      push_stop(tr, 0, start_time);
      push_play(tr, 0, start_time, clip, sample_offset, (uint64_t)sample_pos);
      clip->internal_state_changed = 0;
    }

    if (max_time <= end_time) { /* reaching the end of the clip, :423-437 */
      double offset_from_start = beat_to_samples(max_time - start_time, sample_rate, beat_duration);
      double sample_offset = sample_position + offset_from_start;
      uint32_t buffer_offset = (uint32_t)((uint64_t)sample_offset % (uint64_t)buffer_size);
      push_stop(tr, buffer_offset, max_time);
      tr->partially_ended = 0;
    } else { /* :438-444 */
      tr->partially_ended = 1;
      break;
    }
    next_clip++;
  }
  tr->clip_idx = next_clip; /* :450 */
}

/* ---- EXTENSION (parity unpinned w.r.t. whitebox): polyphase resampler ------------------------------------ */
#define POLY_PHASES 128
#define POLY_TAPS 16
/* h[ph][k] = sinc(u) * blackman(u), u = (k - 7) - ph / 128, normalised to unit DC gain per phase; f64 -> f32. */
static void design_polyphase(float* table) {
  const double pi = 3.141592653589793238462643383279502884;
  for (int ph = 0; ph < POLY_PHASES; ph++) {
    double h[POLY_TAPS], sum = 0.0;
    for (int k = 0; k < POLY_TAPS; k++) {
      const double u = (double)(k - 7) - (double)ph / (double)POLY_PHASES;
      const double sinc = u == 0.0 ? 1.0 : sin(pi * u) / (pi * u);
      const double w = 0.42 + 0.5 * cos(pi * u / 8.0) + 0.08 * cos(2.0 * pi * u / 8.0);
      h[k] = sinc * w;
      sum += h[k];
    }
    for (int k = 0; k < POLY_TAPS; k++) table[ph * POLY_TAPS + k] = (float)(h[k] / sum);
  }
}

void wbo_set_resampler(wbo_session* s, int mode) {
  s->resampler = mode;
  if (mode == 1) design_polyphase(s->poly);
}

/* Same position arithmetic as sample_linear (sampler.cpp:50-52); the fractional part picks a phase, the 16 taps sit
 * on source frames ix - 7 .. ix + 8 (zero outside the sample), f32 fused multiply-adds in tap order. */
static float sample_polyphase(const float* table, const float* src, size_t count, double x) {
  const int64_t ix = (int64_t)x;
  const double fd = x - (double)ix;
  const int ph = (int)(fd * (double)POLY_PHASES);
  const float* h = table + ph * POLY_TAPS;
  float acc = 0.0f;
  for (int k = 0; k < POLY_TAPS; k++) {
    const int64_t i = ix - 7 + k;
    const float v = (i >= 0 && i < (int64_t)(count + SAMPLE_PADDING)) ? src[i] : 0.0f;
    acc = fmaf(h[k], v, acc);
  }
  return acc;
}

/* ---- sampler --------------------------------------------------------------------------------------- */

static float clampf(float x, float lo, float hi) { /* math::clamp, core_math.h:34-38 */
  float m = x < hi ? x : hi;
  return m > lo ? m : lo;
}
static double clampd(double x, double lo, double hi) {
  double m = x < hi ? x : hi;
  return m > lo ? m : lo;
}

/* dsp::Sampler::reset_state (dsp/sampler.h:18-27) */
static void sampler_reset(o_track* tr, double sample_offset, double speed, double src_rate, double dst_rate) {
  tr->playback_speed = (src_rate / dst_rate) * speed;
  tr->sample_offset = sample_offset;
}

/* dsp::Sampler::stream (dsp/sampler.cpp:88-210) incl. sample_linear<T,Fmt> (dsp/sampler.cpp:34-59). */
static const float* g_poly_table = NULL; /* non-NULL while a session in polyphase mode is rendering */

static void sampler_stream(o_track* tr, o_sample* sm, uint32_t num_channels, uint32_t num_samples,
                           uint32_t buffer_offset, float gain, float** dst) {
  const float i16_norm = 1.0f / (float)INT16_MAX;              /* :95 */
  const double i24_norm = 1.0 / (double)((1 << 23) - 1);       /* :96 */
  const double i32_norm = 1.0 / (double)INT32_MAX;             /* :97 */
  const float i16_norm_lin = (float)(1.0 / (double)INT16_MAX); /* get_pcm_sample_normalizer, :7-18 */

  if (tr->sample_offset >= (double)sm->count) return; /* :99-100 */

  double stream_max_length = ((double)sm->count - tr->sample_offset) / tr->playback_speed;
  double next_sample_offset = tr->sample_offset + ((double)num_samples * tr->playback_speed);
  uint32_t ceil_len = (uint32_t)ceil(stream_max_length);
  uint32_t n = num_samples < ceil_len ? num_samples : ceil_len; /* :104 */

  if (tr->playback_speed == 1.0) { /* :106-158 */
    uint32_t off = (uint32_t)tr->sample_offset;
    for (uint32_t i = 0; i < num_channels; i++) {
      uint32_t c = i % sm->channels;
      float* out = dst[i] + buffer_offset;
      switch (sm->fmt) {
        case WBO_FMT_I16: {
          const int16_t* d = (const int16_t*)sm->data[c];
          for (uint32_t j = 0; j < n; j++) {
            float v = (float)d[off + j] * i16_norm;
            out[j] += clampf(v, -1.0f, 1.0f) * gain;
          }
          break;
        }
        case WBO_FMT_I24: {
          const int32_t* d = (const int32_t*)sm->data[c];
          for (uint32_t j = 0; j < n; j++) {
            double v = (double)d[off + j] * i24_norm;
            out[j] += (float)clampd(v, -1.0, 1.0) * gain;
          }
          break;
        }
        case WBO_FMT_I32: {
          const int32_t* d = (const int32_t*)sm->data[c];
          for (uint32_t j = 0; j < n; j++) {
            double v = (double)d[off + j] * i32_norm;
            out[j] += (float)clampd(v, -1.0, 1.0) * gain;
          }
          break;
        }
        default: {
          const float* d = (const float*)sm->data[c];
          for (uint32_t j = 0; j < n; j++) out[j] += d[off + j] * gain;
          break;
        }
      }
    }
  } else if (g_poly_table && sm->fmt == WBO_FMT_F32 && sm->channels >= 2 && num_channels == 2) {
    /* EXTENSION: polyphase quality mode (stereo f32 sources on a stereo bus) */
    for (uint32_t i = 0; i < num_channels; i++) {
      float* out = dst[i] + buffer_offset;
      for (uint32_t j = 0; j < n; j++) {
        const double x = tr->sample_offset + ((double)j * tr->playback_speed);
        out[j] += sample_polyphase(g_poly_table, (const float*)sm->data[i], sm->count, x) * gain;
      }
    }
  } else { /* sample_linear, :34-59. The reference indexes src_channels[i] without `% channels`
              (:47, undefined for mono sources); fixtures never combine mono with speed != 1. */
    for (uint32_t i = 0; i < num_channels; i++) {
      uint32_t c = i % sm->channels;
      float* out = dst[i] + buffer_offset;
      for (uint32_t j = 0; j < n; j++) {
        const double x = tr->sample_offset + ((double)j * tr->playback_speed);
        const int64_t ix = (int64_t)x;
        const float fx = (float)(x - (double)ix);
        float a, b;
        switch (sm->fmt) {
          case WBO_FMT_I16:
            a = (float)(i16_norm_lin * (float)((const int16_t*)sm->data[c])[ix]);
            b = (float)(i16_norm_lin * (float)((const int16_t*)sm->data[c])[ix + 1]);
            break;
          case WBO_FMT_I24:
            a = (float)(i24_norm * (double)((const int32_t*)sm->data[c])[ix]);
            b = (float)(i24_norm * (double)((const int32_t*)sm->data[c])[ix + 1]);
            break;
          case WBO_FMT_I32:
            a = (float)(i32_norm * (double)((const int32_t*)sm->data[c])[ix]);
            b = (float)(i32_norm * (double)((const int32_t*)sm->data[c])[ix + 1]);
            break;
          default:
            a = ((const float*)sm->data[c])[ix];
            b = ((const float*)sm->data[c])[ix + 1];
            break;
        }
        const float sv = a + fx * (b - a);
        out[j] += sv * gain;
      }
    }
  }
  tr->sample_offset = next_sample_offset; /* :209 */
}

/* ---- EXTENSION (parity unpinned w.r.t. whitebox): clip fade envelope ------------------------------------- */
/* The reference stores AudioClip::fade_start / fade_end (clip.h:41-42, project.cpp:189-190) and draws the
 * handles, but no audio code reads them. Builder's specification, shared with include/wbx.h:
 *   n     clip-relative OUTPUT frame (0 at the clip's first rendered frame; a mid-clip start begins at
 *         (uint64)beat_to_samples(start_time - min_time), like the sample offset of track.cpp:378-379)
 *   Fin   beat_to_samples(fade_start), Fout = beat_to_samples(fade_end), L = beat_to_samples(max_time - min_time)
 *   env   (float)( (Fin > 0 ? min(1, n / Fin) : 1) * (Fout > 0 ? min(1, max(0, (L - n) / Fout)) : 1) )   [f64 math]
 *   frame (src * gain) * env     — env == 1.0f leaves the reference's value untouched, so (0, 0) fades are the
 *                                  reference path bit for bit. */
static float fade_env(double n, double fin, double fout, double len) {
  double e = 1.0;
  if (fin > 0.0) {
    double r = n / fin;
    e = r < 1.0 ? r : 1.0;
  }
  if (fout > 0.0) {
    double r = (len - n) / fout;
    r = r > 0.0 ? r : 0.0;
    r = r < 1.0 ? r : 1.0;
    e = e * r;
  }
  return (float)e;
}

/* One Sampler::stream call as Track::process issues it, plus the fade extension. The mixing buffer holds 0 in
 * the frames a call writes (one voice per track), so scaling what the call just added is exactly
 * (0 + src * gain) * env. */
static void track_stream(wbo_session* s, o_track* tr, uint32_t num_samples, uint32_t buffer_offset, float** out) {
  o_clip* clip = tr->current.clip;
  o_sample* sm = tr->current.sample;
  const uint64_t clip_frame = tr->clip_frame;
  tr->clip_frame += num_samples;
  if (!(clip->fade_start > 0.0 || clip->fade_end > 0.0)) {
    sampler_stream(tr, sm, s->C, num_samples, buffer_offset, clip->gain, out);
    return;
  }
  if (tr->sample_offset >= (double)sm->count) return;
  double stream_max_length = ((double)sm->count - tr->sample_offset) / tr->playback_speed;
  uint32_t ceil_len = (uint32_t)ceil(stream_max_length);
  uint32_t n = num_samples < ceil_len ? num_samples : ceil_len;
  sampler_stream(tr, sm, s->C, num_samples, buffer_offset, clip->gain, out);
  const double rate = (double)s->rate;
  const double fin = beat_to_samples(clip->fade_start, rate, s->beat_duration);
  const double fout = beat_to_samples(clip->fade_end, rate, s->beat_duration);
  const double len = beat_to_samples(clip->max_time - clip->min_time, rate, s->beat_duration);
  for (uint32_t j = 0; j < n; j++) {
    const float env = fade_env((double)clip_frame + (double)j, fin, fout, len);
    for (uint32_t c = 0; c < s->C; c++) out[c][buffer_offset + j] = out[c][buffer_offset + j] * env;
  }
}

/* ---- EXTENSION (parity unpinned w.r.t. whitebox): per-track EQ + compressor ----------------------------- */
/* RBJ "Audio EQ Cookbook" biquads, designed in f64, normalised by a0, stored as f32. */
static void design_band(int band, double freq, double gain_db, double q, double rate, float* b0, float* b1, float* b2,
                        float* a1, float* a2) {
  const double pi = 3.141592653589793238462643383279502884;
  const double A = pow(10.0, gain_db / 40.0);
  const double w0 = 2.0 * pi * freq / rate;
  const double cw = cos(w0), sw = sin(w0);
  const double alpha = sw / (2.0 * q);
  double B0, B1, B2, A0, A1, A2;
  if (band == 0) { /* low shelf */
    const double sq = 2.0 * sqrt(A) * alpha;
    B0 = A * ((A + 1.0) - (A - 1.0) * cw + sq);
    B1 = 2.0 * A * ((A - 1.0) - (A + 1.0) * cw);
    B2 = A * ((A + 1.0) - (A - 1.0) * cw - sq);
    A0 = (A + 1.0) + (A - 1.0) * cw + sq;
    A1 = -2.0 * ((A - 1.0) + (A + 1.0) * cw);
    A2 = (A + 1.0) + (A - 1.0) * cw - sq;
  } else if (band == 3) { /* high shelf */
    const double sq = 2.0 * sqrt(A) * alpha;
    B0 = A * ((A + 1.0) + (A - 1.0) * cw + sq);
    B1 = -2.0 * A * ((A - 1.0) + (A + 1.0) * cw);
    B2 = A * ((A + 1.0) + (A - 1.0) * cw - sq);
    A0 = (A + 1.0) - (A - 1.0) * cw + sq;
    A1 = 2.0 * ((A - 1.0) - (A + 1.0) * cw);
    A2 = (A + 1.0) - (A - 1.0) * cw - sq;
  } else { /* peaking */
    B0 = 1.0 + alpha * A;
    B1 = -2.0 * cw;
    B2 = 1.0 - alpha * A;
    A0 = 1.0 + alpha / A;
    A1 = -2.0 * cw;
    A2 = 1.0 - alpha / A;
  }
  *b0 = (float)(B0 / A0);
  *b1 = (float)(B1 / A0);
  *b2 = (float)(B2 / A0);
  *a1 = (float)(A1 / A0);
  *a2 = (float)(A2 / A0);
}

static void fx_free(o_fx* f) {
  free(f->hist[0]);
  free(f->hist[1]);
  f->hist[0] = f->hist[1] = NULL;
}

int wbo_set_impulse_response(wbo_session* s, const float* h, uint32_t n_taps) {
  free(s->ir);
  s->ir = NULL;
  s->ir_taps = 0;
  if (h && n_taps) {
    s->ir = (float*)malloc(n_taps * sizeof(float));
    memcpy(s->ir, h, n_taps * sizeof(float));
    s->ir_taps = n_taps;
  }
  for (uint32_t t = 0; t < s->n_tracks; t++) { /* a new response starts from silence */
    o_fx* f = &s->tracks[t]->fx;
    fx_free(f);
    f->hist_pos = 0;
  }
  return 0;
}

int wbo_set_effects(wbo_session* s, int track, const wbo_effects* fx) {
  o_fx* f = &s->tracks[track]->fx;
  fx_free(f);
  fx64_forget(f);
  memset(f, 0, sizeof(*f));
  if (!fx) return 0;
  f->reverb_on = fx->reverb_on != 0;
  for (int b = 0; b < 4; b++) {
    if (fx->eq_gain_db[b] != 0.0f) f->eq_on = 1;
    design_band(b, fx->eq_freq[b], fx->eq_gain_db[b], fx->eq_q[b], (double)s->rate, &f->b0[b], &f->b1[b], &f->b2[b],
                &f->a1[b], &f->a2[b]);
  }
  f->ratio_code = fx->comp_ratio_code;
  f->comp_on = fx->comp_ratio_code != 0;
  f->thr = (float)pow(10.0, (double)fx->comp_threshold_db / 20.0);
  f->makeup = (float)pow(10.0, (double)fx->comp_makeup_db / 20.0);
  f->att = (float)exp(-1.0 / ((double)fx->comp_attack_ms * 0.001 * (double)s->rate));
  f->rel = (float)exp(-1.0 / ((double)fx->comp_release_ms * 0.001 * (double)s->rate));
  fx_design_tables(f);
  return 0;
}

/* Convolution reverb, one channel, one callback, in place: y[n] = (float) sum_k h[k] * x[n-k] in f64, k ascending. */
static void apply_reverb(wbo_session* s, o_fx* f, int c, float* buf, uint32_t n, uint64_t pos0) {
  const uint32_t L = s->ir_taps;
  if (L == 0) return;
  const uint32_t H = L > 1 ? L - 1 : 1;
  if (!f->hist[c]) f->hist[c] = (float*)calloc(H, sizeof(float));
  float* x = (float*)malloc(((size_t)H + n) * sizeof(float)); /* [history | this callback's input] */
  for (uint32_t i = 0; i < H; i++) x[i] = L > 1 ? f->hist[c][(pos0 + i) % H] : 0.0f; /* oldest first */
  memcpy(x + H, buf, n * sizeof(float));
  for (uint32_t j = 0; j < n; j++) {
    double acc = 0.0;
    for (uint32_t k = 0; k < L; k++) acc += (double)s->ir[k] * (double)x[H + j - k];
    buf[j] = (float)acc;
  }
  if (L > 1)
    for (uint32_t j = 0; j < n; j++) f->hist[c][(pos0 + j) % H] = x[H + j]; /* ring: overwrite the oldest */
  free(x);
}

/* The chain in its textbook sample-by-sample form (transposed direct form II biquads, branching peak follower): what
 * apply_effects computes up to rounding. Kept as the accuracy yardstick of the time-parallel specification below
 * (tests/test_oracle.py holds the two within 1e-5 of the block peak) and selectable with wbo_set_fx_textbook(1). */
static void apply_effects_textbook(o_fx* f, int c, float* buf, uint32_t n) {
  for (uint32_t j = 0; j < n; j++) {
    float x = buf[j];
    if (f->eq_on) {
      for (int b = 0; b < 4; b++) { /* transposed direct form II */
        const float y = fmaf(f->b0[b], x, f->s1[c][b]);
        f->s1[c][b] = fmaf(f->b1[b], x, fmaf(-f->a1[b], y, f->s2[c][b]));
        f->s2[c][b] = fmaf(f->b2[b], x, -(f->a2[b] * y));
        x = y;
      }
    }
    if (f->comp_on) {
      const float xa = fabsf(x);
      float env = f->env[c];
      env = xa > env ? fmaf(f->att, env - xa, xa) : fmaf(f->rel, env - xa, xa); /* peak follower */
      f->env[c] = env;
      float g = 1.0f;
      if (env > f->thr) {
        const float r = f->thr / env; /* (thr/env)^(1 - 1/ratio) with divide and square roots only */
        const float r2 = sqrtf(r);
        switch (f->ratio_code) {
          case 1: g = r2; break;                                /* 2:1 -> r^(1/2) */
          case 2: g = r2 * sqrtf(r2); break;                    /* 4:1 -> r^(3/4) */
          case 3: g = (r2 * sqrtf(r2)) * sqrtf(sqrtf(r2)); break; /* 8:1 -> r^(7/8) */
          default: g = r; break;                                /* limiter */
        }
      }
      x = (x * g) * f->makeup;
    }
    buf[j] = x;
  }
}


/* The same chain with f64 state and arithmetic (coefficients as designed, i.e. the f32 values): ground truth for the
 * accuracy of both f32 evaluations. State lives in a side table keyed by the o_fx address (test use only). */
typedef struct {
  double s1[2][4], s2[2][4], env[2];
} o_fx64;
static struct {
  const o_fx* key;
  o_fx64 st;
} g_fx64[64];
static o_fx64* fx64_state(const o_fx* f) {
  for (int i = 0; i < 64; i++)
    if (g_fx64[i].key == f) return &g_fx64[i].st;
  for (int i = 0; i < 64; i++)
    if (!g_fx64[i].key) {
      g_fx64[i].key = f;
      memset(&g_fx64[i].st, 0, sizeof(o_fx64));
      return &g_fx64[i].st;
    }
  return NULL;
}
static void fx64_forget(const o_fx* f) {
  for (int i = 0; i < 64; i++)
    if (g_fx64[i].key == f) g_fx64[i].key = NULL;
}
static void apply_effects_f64(o_fx* f, int c, float* buf, uint32_t n) {
  o_fx64* st = fx64_state(f);
  if (!st) return;
  for (uint32_t j = 0; j < n; j++) {
    double x = (double)buf[j];
    if (f->eq_on) {
      for (int b = 0; b < 4; b++) {
        const double y = (double)f->b0[b] * x + st->s1[c][b];
        st->s1[c][b] = (double)f->b1[b] * x - (double)f->a1[b] * y + st->s2[c][b];
        st->s2[c][b] = (double)f->b2[b] * x - (double)f->a2[b] * y;
        x = y;
      }
    }
    if (f->comp_on) {
      const double xa = fabs(x);
      double env = st->env[c];
      env = xa > env ? (double)f->att * (env - xa) + xa : (double)f->rel * (env - xa) + xa;
      st->env[c] = env;
      double g = 1.0;
      if (env > (double)f->thr) {
        const double q = (double)f->thr / env;
        switch (f->ratio_code) {
          case 1: g = pow(q, 0.5); break;
          case 2: g = pow(q, 0.75); break;
          case 3: g = pow(q, 0.875); break;
          default: g = q; break;
        }
      }
      x = (x * g) * (double)f->makeup;
    }
    buf[j] = (float)x;
  }
}

static int g_fx_textbook = 0; /* 0: the specification; 1: textbook f32; 2: textbook f64 */
void wbo_set_fx_textbook(int mode) { g_fx_textbook = mode; }

/* ---- the chain's SPECIFICATION: the same filters evaluated in a time-parallel association -------------------
 * A biquad and the follower are recurrences in time; evaluated sample by sample a GPU is bound by the latency of the
 * dependent operations. The specification therefore fixes a blocked evaluation order that exposes the parallelism —
 * the CUDA kernel (wbx_kernels.cu fx_chain_kernel) performs exactly these IEEE operations, so CUDA == this bit for bit:
 *  - a callback is processed in chunks of up to 512 frames = 32 segments ("lanes") of 16 frames;
 *  - biquad b in state-space form s' = A s + B x, y = s1 + b0 x with A = [[-a1, 1], [-a2, 0]], B = [b1 - a1 b0,
 *    b2 - a2 b0]: every segment is run from a ZERO state (y_zs, end state e), the 32 end states are combined into the
 *    true state at every segment boundary by a Kogge-Stone scan with the matrices A^(16 * 2^j), and every frame gets
 *    the zero-input response of its segment's true start state added: y[m] = y_zs[m] + (A^m)[0][.] . s_start;
 *  - the peak follower env' = (|x| > env ? att : rel) * (env - |x|) + |x| equals mm(att * env + (1 - att)|x|,
 *    rel * env + (1 - rel)|x|) with mm = max when att <= rel (min otherwise); mm-of-affine maps compose, so four steps
 *    are env4 = mm_i(S4[i] * env0 + Q4[i]), i = number of attack steps taken, S4[i] = att^i rel^(4-i), with the
 *    intercepts Q built by a small dynamic program from the inputs alone: only ONE fused multiply-add + mm per block
 *    of 4 frames is serial in time; the envelope inside a block is the 1-, 2-, 3-step look-ahead from the block start;
 *  - the gain computer and make-up gain are memoryless and unchanged.
 * All tables are designed in f64 from the f32 coefficients and rounded once. */
#define FX_SEG 16u
#define FX_LANES 32u
#define FX_CHUNK (FX_SEG * FX_LANES)

static void mat2_mul(const double a[4], const double b[4], double out[4]) {
  const double o0 = a[0] * b[0] + a[1] * b[2], o1 = a[0] * b[1] + a[1] * b[3];
  const double o2 = a[2] * b[0] + a[3] * b[2], o3 = a[2] * b[1] + a[3] * b[3];
  out[0] = o0, out[1] = o1, out[2] = o2, out[3] = o3;
}

static void fx_design_tables(o_fx* f) {
  for (int b = 0; b < 4; b++) {
    const double a1 = (double)f->a1[b], a2 = (double)f->a2[b], b0 = (double)f->b0[b];
    f->B1[b] = (float)((double)f->b1[b] - a1 * b0);
    f->B2[b] = (float)((double)f->b2[b] - a2 * b0);
    const double A[4] = {-a1, 1.0, -a2, 0.0};
    double pw[4] = {1.0, 0.0, 0.0, 1.0};
    for (int m = 0; m <= 16; m++) {
      for (int q = 0; q < 4; q++) f->P[b][m][q] = (float)pw[q];
      if (m < 16) mat2_mul(pw, A, pw);
    }
    double sq[4] = {pw[0], pw[1], pw[2], pw[3]}; /* A^16 */
    for (int j = 0; j < 5; j++) {
      for (int q = 0; q < 4; q++) f->S[b][j][q] = (float)sq[q];
      mat2_mul(sq, sq, sq);
    }
  }
  const float a = f->att, r = f->rel;
  f->sel = a <= r;
  f->a1m = 1.0f - a;
  f->r1m = 1.0f - r;
  const float r2 = r * r, a2 = a * a, r3 = r2 * r, a3 = a2 * a;
  float* sl = f->sl;
  sl[0] = r, sl[1] = a;
  sl[2] = r2, sl[3] = a * r, sl[4] = a2;
  sl[5] = r3, sl[6] = a * r2, sl[7] = a2 * r, sl[8] = a3;
  sl[9] = r2 * r2, sl[10] = a * r3, sl[11] = a2 * r2, sl[12] = a3 * r, sl[13] = a2 * a2;
}

static float fx_mm(int sel, float x, float y) { return sel ? fmaxf(x, y) : fminf(x, y); }

/* one biquad over one chunk of n <= 512 frames, in place */
static void fx_eq_stage(o_fx* f, int c, int b, float* x, uint32_t n) {
  const float b0 = f->b0[b], na1 = -f->a1[b], na2 = -f->a2[b], B1 = f->B1[b], B2 = f->B2[b];
  float e1[FX_LANES], e2[FX_LANES], v1[FX_LANES], v2[FX_LANES], yz[FX_CHUNK];
  for (uint32_t l = 0; l < FX_LANES; l++) { /* zero-state pass of every segment */
    const uint32_t f0 = l * FX_SEG;
    const uint32_t len = f0 >= n ? 0u : (n - f0 < FX_SEG ? n - f0 : FX_SEG);
    float s1 = 0.0f, s2 = 0.0f;
    for (uint32_t m = 0; m < len; m++) {
      const float xi = x[f0 + m];
      yz[f0 + m] = fmaf(b0, xi, s1);
      const float t = fmaf(B1, xi, s2);
      s2 = fmaf(na2, s1, B2 * xi);
      s1 = fmaf(na1, s1, t);
    }
    e1[l] = s1, e2[l] = s2;
  }
  const float in1 = f->s1[c][b], in2 = f->s2[c][b];
  memcpy(v1, e1, sizeof(v1));
  memcpy(v2, e2, sizeof(v2));
  { /* the incoming state enters through segment 0 */
    const float* S = f->S[b][0];
    v1[0] = fmaf(S[0], in1, fmaf(S[1], in2, e1[0]));
    v2[0] = fmaf(S[2], in1, fmaf(S[3], in2, e2[0]));
  }
  for (int j = 0; j < 5; j++) { /* Kogge-Stone: v[l] += A^(16 d) v[l - d] */
    const uint32_t d = 1u << j;
    const float* S = f->S[b][j];
    float n1[FX_LANES], n2[FX_LANES];
    for (uint32_t l = 0; l < FX_LANES; l++) {
      if (l >= d) {
        n1[l] = fmaf(S[0], v1[l - d], fmaf(S[1], v2[l - d], v1[l]));
        n2[l] = fmaf(S[2], v1[l - d], fmaf(S[3], v2[l - d], v2[l]));
      } else {
        n1[l] = v1[l], n2[l] = v2[l];
      }
    }
    memcpy(v1, n1, sizeof(v1));
    memcpy(v2, n2, sizeof(v2));
  }
  for (uint32_t l = 0; l < FX_LANES; l++) { /* zero-input response of the segment's true start state */
    const uint32_t f0 = l * FX_SEG;
    if (f0 >= n) break;
    const uint32_t len = n - f0 < FX_SEG ? n - f0 : FX_SEG;
    const float st1 = l ? v1[l - 1] : in1, st2 = l ? v2[l - 1] : in2;
    for (uint32_t m = 0; m < len; m++) x[f0 + m] = fmaf(f->P[b][m][1], st2, fmaf(f->P[b][m][0], st1, yz[f0 + m]));
    if (f0 + len == n) { /* state after the chunk's last frame */
      f->s1[c][b] = fmaf(f->P[b][len][0], st1, fmaf(f->P[b][len][1], st2, e1[l]));
      f->s2[c][b] = fmaf(f->P[b][len][2], st1, fmaf(f->P[b][len][3], st2, e2[l]));
    }
  }
}

/* compressor over one chunk, in place */
static void fx_comp_stage(o_fx* f, int c, float* x, uint32_t n) {
  const float a = f->att, r = f->rel, a1m = f->a1m, r1m = f->r1m;
  const float* sl = f->sl;
  const int sel = f->sel;
  float env[FX_CHUNK];
  float e = f->env[c];
  const uint32_t nb = n / 4;
  for (uint32_t k = 0; k < nb; k++) {
    float pa[4], pr[4];
    for (int m = 0; m < 4; m++) {
      const float xa = fabsf(x[4 * k + m]);
      pa[m] = a1m * xa;
      pr[m] = r1m * xa;
    }
    /* intercepts: Q_m[i] = best intercept after m steps of which i took the attack branch */
    float q1[2], q2[3], q3[4], q4[5];
    q1[0] = pr[0], q1[1] = pa[0];
    q2[0] = fmaf(r, q1[0], pr[1]);
    q2[1] = fx_mm(sel, fmaf(r, q1[1], pr[1]), fmaf(a, q1[0], pa[1]));
    q2[2] = fmaf(a, q1[1], pa[1]);
    q3[0] = fmaf(r, q2[0], pr[2]);
    for (int i = 1; i < 3; i++) q3[i] = fx_mm(sel, fmaf(r, q2[i], pr[2]), fmaf(a, q2[i - 1], pa[2]));
    q3[3] = fmaf(a, q2[2], pa[2]);
    q4[0] = fmaf(r, q3[0], pr[3]);
    for (int i = 1; i < 4; i++) q4[i] = fx_mm(sel, fmaf(r, q3[i], pr[3]), fmaf(a, q3[i - 1], pa[3]));
    q4[4] = fmaf(a, q3[3], pa[3]);
    float t = fmaf(sl[0], e, q1[0]);
    env[4 * k] = fx_mm(sel, t, fmaf(sl[1], e, q1[1]));
    t = fmaf(sl[2], e, q2[0]);
    for (int i = 1; i < 3; i++) t = fx_mm(sel, t, fmaf(sl[2 + i], e, q2[i]));
    env[4 * k + 1] = t;
    t = fmaf(sl[5], e, q3[0]);
    for (int i = 1; i < 4; i++) t = fx_mm(sel, t, fmaf(sl[5 + i], e, q3[i]));
    env[4 * k + 2] = t;
    t = fmaf(sl[9], e, q4[0]);
    for (int i = 1; i < 5; i++) t = fx_mm(sel, t, fmaf(sl[9 + i], e, q4[i]));
    env[4 * k + 3] = t;
    e = t; /* the only step that is serial in time */
  }
  for (uint32_t j = 4 * nb; j < n; j++) { /* a chunk's last 1..3 frames: one step at a time */
    const float xa = fabsf(x[j]);
    e = fx_mm(sel, fmaf(r, e, r1m * xa), fmaf(a, e, a1m * xa));
    env[j] = e;
  }
  f->env[c] = e;
  for (uint32_t j = 0; j < n; j++) {
    float g = 1.0f;
    if (env[j] > f->thr) {
      const float q = f->thr / env[j]; /* (thr/env)^(1 - 1/ratio) with divide and square roots only */
      const float q2 = sqrtf(q);
      switch (f->ratio_code) {
        case 1: g = q2; break;                                  /* 2:1 -> q^(1/2) */
        case 2: g = q2 * sqrtf(q2); break;                      /* 4:1 -> q^(3/4) */
        case 3: g = (q2 * sqrtf(q2)) * sqrtf(sqrtf(q2)); break; /* 8:1 -> q^(7/8) */
        default: g = q; break;                                  /* limiter */
      }
    }
    x[j] = (x[j] * g) * f->makeup;
  }
}

/* One channel, one callback, in place. */
static void apply_effects(o_fx* f, int c, float* buf, uint32_t n) {
  if (g_fx_textbook) {
    if (g_fx_textbook == 2)
      apply_effects_f64(f, c, buf, n);
    else
      apply_effects_textbook(f, c, buf, n);
    return;
  }
  for (uint32_t off = 0; off < n; off += FX_CHUNK) {
    const uint32_t len = n - off < FX_CHUNK ? n - off : FX_CHUNK;
    if (f->eq_on)
      for (int b = 0; b < 4; b++) fx_eq_stage(f, c, b, buf + off, len);
    if (f->comp_on) fx_comp_stage(f, c, buf + off, len);
  }
}

/* ---- Track::process (engine/track.cpp:587-736) ---------------------------------------------------------- */

static void track_process(wbo_session* s, o_track* tr, float** out, double sample_rate, double beat_duration,
                          double sample_position, double start_time, double end_time, int playing) {
  const uint32_t B = s->B;

  /* process_track_messages (:602, :773-779) + parameter application (:618-643) */
  if (playing) process_event(tr, start_time, end_time, sample_position, beat_duration, sample_rate, B);
  for (uint32_t i = 0; i < tr->n_msgs; i++) {
    switch (tr->msgs[i].id) {
      case 0: tr->volume = (float)tr->msgs[i].value; break;
      case 1:
        tr->pan = (float)tr->msgs[i].value;
        wbo_panning_coefs(tr->pan, &tr->pan_coeffs[0], &tr->pan_coeffs[1]);
        break;
      case 2: tr->mute = tr->msgs[i].value > 0.0; break;
    }
  }
  tr->n_msgs = 0;

  if (playing) { /* :664-724 */
    uint32_t next = 0;
    uint32_t start_sample = 0;
    while (start_sample < B) {
      if (next != tr->n_events) {
        o_event* ne = &tr->events[next];
        uint32_t event_length = ne->buffer_offset - start_sample; /* unsigned, as in the reference (:670) */
        if (ne->buffer_offset < start_sample || ne->buffer_offset > B) {
          /* Reference UB: events out of offset order (e.g. a clip that ends EXACTLY on the block's end gets
           * StopSample offset `B % B == 0`, track.cpp:425) make event_length wrap and Sampler::stream
           * write past the mixing buffer. No defined answer exists; count it so fixtures can avoid it, and
           * render nothing for this event instead of corrupting memory. */
          g_ub_count++;
          event_length = 0;
          if (tr->current.type == EV_PLAY) tr->current.type = EV_NONE;
        }
        if (tr->current.type == EV_PLAY) track_stream(s, tr, event_length, start_sample, out);
        if (ne->type == EV_PLAY) {
          sampler_reset(tr, (double)ne->sample_offset, ne->speed, (double)ne->sample->rate, sample_rate);
          tr->clip_frame = ne->clip_frame;
        }
        tr->current = *ne;
        start_sample += event_length;
        next++;
      } else {
        uint32_t event_length = B - start_sample;
        if (tr->current.type == EV_PLAY) track_stream(s, tr, event_length, start_sample, out);
        start_sample = B;
      }
    }
  }

  /* A plugin in the slot: write_buffer was effect_buffer (:600), i.e. the clips above never reach output_buffer, which
   * holds what the plugin wrote (:645-662) — nothing, for the no-op plugin the reference is pinned with. */
  if (tr->has_plugin)
    for (uint32_t c = 0; c < s->C; c++) memset(out[c], 0, B * sizeof(float));

  /* dsp::apply_gain (dsp/dsp_ops.h:27-31) + VUMeter::push_samples (engine/vu_meter.h:20-30), :728-733.
   * pan_coeffs has two entries, so C <= 2 (track.h:50). */
  if (tr->fx.eq_on || tr->fx.comp_on) /* EXTENSION: where a native PluginInterface::process would run */
    for (uint32_t c = 0; c < s->C; c++) apply_effects(&tr->fx, (int)c, out[c], B);
  if (tr->fx.reverb_on && s->ir_taps) {
    for (uint32_t c = 0; c < s->C; c++) apply_reverb(s, &tr->fx, (int)c, out[c], B, tr->fx.hist_pos);
    tr->fx.hist_pos += B;
  }

  float volume = tr->mute ? 0.0f : tr->volume;
  for (uint32_t c = 0; c < s->C; c++) {
    float* buf = out[c];
    float g = volume * tr->pan_coeffs[c];
    for (uint32_t j = 0; j < B; j++) buf[j] *= g;
    float new_level = 0.0f;
    for (uint32_t j = 0; j < B; j++) {
      float v = buf[j];
      float a = v < 0 ? -v : v;               /* math::abs, core_math.h:20-22 */
      new_level = a < new_level ? new_level : a; /* math::max(new_level, a), core_math.h:29-32 */
    }
    if (tr->level[c] < new_level) tr->level[c] = new_level;
  }
}

/* ---- Engine::process (engine/engine.cpp:1576-1654) ------------------------------------------------------ */

static void engine_process(wbo_session* s) {
  const uint32_t B = s->B, C = s->C;
  g_poly_table = s->resampler == 1 ? s->poly : NULL;
  const double sample_rate = (double)s->rate;
  double buffer_duration = (double)B / sample_rate;
  double current_beat_duration = s->beat_duration;
  double current_playhead_position = s->playhead;
  double buffer_duration_in_beats = buffer_duration / current_beat_duration;
  double next_playhead_pos = s->playhead + buffer_duration_in_beats;
  int currently_playing = s->playing;

  for (uint32_t t = 0; t < s->n_tracks; t++) s->tracks[t]->n_events = 0; /* :1589-1596 */
  for (uint32_t c = 0; c < C; c++) memset(s->out[c], 0, B * sizeof(float)); /* :1598 */

  for (uint32_t t = 0; t < s->n_tracks; t++) { /* :1600-1617 */
    for (uint32_t c = 0; c < C; c++) memset(s->mixing[c], 0, B * sizeof(float));
    track_process(s, s->tracks[t], s->mixing, sample_rate, current_beat_duration, s->sample_position,
                  current_playhead_position, next_playhead_pos, currently_playing);
    for (uint32_t c = 0; c < C; c++) { /* AudioBuffer::mix, core/audio_buffer.h:73-82 */
      const float* o = s->mixing[c];
      float* b = s->out[c];
      for (uint32_t j = 0; j < B; j++) b[j] += o[j];
    }
  }

  if (currently_playing) { /* :1619-1623 */
    s->sample_position += beat_to_samples(buffer_duration_in_beats, sample_rate, current_beat_duration);
    s->playhead = next_playhead_pos;
  }

  for (uint32_t c = 0; c < C; c++) { /* :1627-1636 — NaN passes through */
    float* ch = s->out[c];
    for (uint32_t j = 0; j < B; j++) {
      if (ch[j] > 1.0)
        ch[j] = 1.0f;
      else if (ch[j] < -1.0)
        ch[j] = -1.0f;
    }
  }
}

int wbo_process(wbo_session* s, uint32_t n_blocks, float* out, float* peaks) {
  for (uint32_t k = 0; k < n_blocks; k++) {
    engine_process(s);
    if (out)
      for (uint32_t c = 0; c < s->C; c++)
        memcpy(out + ((size_t)k * s->C + c) * s->B, s->out[c], s->B * sizeof(float));
    for (uint32_t t = 0; t < s->n_tracks; t++)
      for (uint32_t c = 0; c < 2; c++) {
        if (peaks) peaks[((size_t)k * s->n_tracks + t) * 2 + c] = s->tracks[t]->level[c];
        s->tracks[t]->level[c] = 0.0f; /* VUMeter::update's exchange(0), vu_meter.h:33 */
      }
  }
  return 0;
}

double wbo_time_process(wbo_session* s, uint32_t n_blocks) {
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (uint32_t k = 0; k < n_blocks; k++) engine_process(s);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

double wbo_sampler_offset(wbo_session* s, int track) { return s->tracks[track]->sample_offset; }
double wbo_sample_position(wbo_session* s) { return s->sample_position; }
double wbo_playhead(wbo_session* s) { return s->playhead; }

/* ---- gfx/waveform_visual.cpp: waveform peak mip-maps ------------------------------------------------------- */
/* summarize_for_mipmaps_impl<T> (:9-173) for one channel. T = int16 (quality High) or int8 (Low); per chunk the
 * converted min and max with their FIRST occurrences, emitted in order of occurrence. */
static void summarize(int fmt, size_t sample_count, const void* data, size_t chunk_count, size_t block_count,
                      size_t output_count, int high, void* output) {
  const long tmin = high ? -32768 : -128, tmax = high ? 32767 : 127;
  for (size_t i = 0; i < output_count; i += 2) {
    size_t idx = i * block_count;
    size_t rem = sample_count - idx;
    size_t chunk_length = chunk_count < rem ? chunk_count : rem;
    long min_val = tmax, max_val = tmin; /* numeric_limits<T>::max() / ::min() */
    size_t min_idx = 0, max_idx = 0;
    for (size_t j = 0; j < chunk_length; j++) {
      long value;
      if (fmt == WBO_FMT_F32) { /* :143-151 */
        float v = ((const float*)data)[idx + j];
        float conv = v * (v >= 0.0f ? (float)tmax : (float)(-tmin));
        value = high ? (long)(int16_t)conv : (long)(int8_t)conv;
      } else if (fmt == WBO_FMT_I16) { /* :66-76 */
        const float dmin = (float)tmin / (float)INT16_MIN, dmax = (float)tmax / (float)INT16_MAX;
        int16_t v = ((const int16_t*)data)[idx + j];
        float conv = (float)v * (v >= 0 ? dmax : dmin);
        value = high ? (long)(int16_t)conv : (long)(int8_t)conv;
      } else { /* I32, :104-114, in double */
        const double dmin = (double)tmin / (double)INT32_MIN, dmax = (double)tmax / (double)INT32_MAX;
        int32_t v = ((const int32_t*)data)[idx + j];
        double conv = (double)v * (v >= 0 ? dmax : dmin);
        value = high ? (long)(int16_t)conv : (long)(int8_t)conv;
      }
      if (value < min_val) {
        min_val = value;
        min_idx = j;
      }
      if (value > max_val) {
        max_val = value;
        max_idx = j;
      }
    }
    long first = max_idx < min_idx ? max_val : min_val, second = max_idx < min_idx ? min_val : max_val;
    if (high) {
      ((int16_t*)output)[i] = (int16_t)first;
      ((int16_t*)output)[i + 1] = (int16_t)second;
    } else {
      ((int8_t*)output)[i] = (int8_t)first;
      ((int8_t*)output)[i + 1] = (int8_t)second;
    }
  }
}

/* WaveformVisual::create (:181-248): levels current_mip = 1, 3, 5, ... while sample_count > 64 (/= 4 per level). */
int wbo_mipmap(wbo_session* s, int sample, int quality, int level, void* out, uint64_t cap_elems, uint32_t* count) {
  o_sample* sm = s->samples[sample];
  if (sm->fmt == WBO_FMT_I24) return 0; /* `default: break` — the loader never produces this tag */
  size_t sample_count = sm->count;
  uint32_t current_mip = 1;
  int n_levels = 0;
  const size_t esz = quality ? 2 : 1;
  while (sample_count > 64) {
    size_t chunk_count = (size_t)1 << current_mip;
    size_t block_count = (size_t)1 << (current_mip - 1);
    size_t mip_data_count = sm->count / block_count;
    mip_data_count += mip_data_count % 2;
    if (n_levels == level) {
      if (count) *count = (uint32_t)mip_data_count;
      if (out && mip_data_count * sm->channels <= cap_elems)
        for (uint32_t c = 0; c < sm->channels; c++)
          summarize(sm->fmt, sm->count, sm->data[c], chunk_count, block_count, mip_data_count, quality,
                    (char*)out + mip_data_count * c * esz);
    }
    n_levels++;
    sample_count /= 4;
    current_mip += 2;
  }
  return n_levels;
}

/* ---- core/audio_format_conv.cpp:5-106: planar f32 -> interleaved device format -------------------------- */

void wbo_interleave(void* dst, const float* const* src, uint32_t offset, uint32_t frames, uint32_t channels,
                    int fmt) {
  for (uint32_t c = 0; c < channels; c++) {
    const float* ch = src[c] + offset;
    for (uint32_t i = 0; i < frames; i++) {
      float v = ch[i];
      size_t o = (size_t)i * channels + c;
      switch (fmt) {
        case WBO_FMT_I16: /* :5-20: positive * 32767, else * 32768, truncating cast */
          ((int16_t*)dst)[o] = (int16_t)(v > 0.0f ? v * 32767.0f : v * 32768.0f);
          break;
        case WBO_FMT_I24: { /* :22-43 */
          int32_t q = v > 0.0f ? (int32_t)(v * 8388607.0f) : (int32_t)(v * 8388608.0f);
          /* the reference writes channel c's bytes at dst[3*i .. 3*i+2] for EVERY channel (no channel
           * stride, :31-41): later channels overwrite earlier ones. Restated as written. */
          uint8_t* p = (uint8_t*)dst + (size_t)i * 3;
          p[0] = (uint8_t)q;
          p[1] = (uint8_t)(q >> 8);
          p[2] = (uint8_t)(q >> 16);
          break;
        }
        case 6: { /* I24_X8, :45-59 */
          int32_t q = v > 0.0f ? (int32_t)(v * 8388607.0f) : (int32_t)(v * 8388608.0f);
          ((int32_t*)dst)[o] = q & 0xFFFFFF;
          break;
        }
        case WBO_FMT_I32: /* :61-74, in double */
          ((int32_t*)dst)[o] = (int32_t)(v > 0.0f ? (double)v * 2147483647.0 : (double)v * 2147483648.0);
          break;
        default: ((float*)dst)[o] = v; break; /* :76-88 */
      }
    }
  }
}
