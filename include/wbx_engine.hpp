// wbx_engine.hpp — C++ host side of the B200 mixing path: keeps the reference engine's call shape
// (wb::Engine, engine/engine.h:26-271; wb::Track, engine/track.h:90-273) for everything the mix needs, does
// the transport / clip scheduling / parameter bookkeeping on the host in doubles exactly as the reference
// does (engine/track.cpp:258-451, 587-736), and hands the sample work to the CUDA engine through the C ABI
// in wbx.h. It never touches a sample value itself and has no CPU render path.
//
// Drop-in use: where audio_io_* calls `engine->process(input_buffer_, output_buffer_, sample_rate)`
// (engine/audio_io_pulseaudio.cpp:411, engine/audio_io_wasapi.cpp:708) a wbx::Engine can be called with the
// same wb::AudioBuffer<float> objects: process() is a template over any buffer type that exposes
// `n_samples`, `n_channels` and `channel_buffers` (core/audio_buffer.h:19-23).
//
// Threading contract — the reference's, kept:
//   * ONE audio thread calls process() / render(); it holds Engine::editor_lock — a yield-spinning lock
//     (core/thread.h:11-35) — for the whole callback (engine/engine.cpp:1587,1651);
//   * ONE UI thread edits: every Engine method that changes tracks, clips, transport or configuration takes the same
//     lock (engine.h:89-94; e.g. engine.cpp:204-206, 301), so an edit lands between two callbacks;
//   * Track::set_volume / set_pan / set_mute bypass the lock: they push a message into the track's lock-free
//     single-producer / single-consumer ring (core/queue.h:142-196, capacity 64, engine/track.cpp:23) that the audio
//     thread drains at the start of the next callback (track.cpp:773-779); a full ring makes the producer yield;
//   * VU levels go back through std::atomic<float>: the audio thread raises them with a CAS-max, the UI takes them
//     with exchange(0) (engine/vu_meter.h:17-40);
//   * the load meter (core/timing.h:54-67, engine.cpp:1577,1653) is an atomic<double> EMA the UI may read at any time.
#pragma once
#include <atomic>
#include <chrono>
#include <cstdint>
#include <string>
#include <thread>
#include <vector>

#include "wbx.h"

namespace wbx {

// core/thread.h:11-35
struct Spinlock {
  std::atomic<bool> lock_{false};
  bool try_lock() noexcept { return !lock_.load(std::memory_order_relaxed) && !lock_.exchange(true, std::memory_order_acquire); }
  void lock() noexcept {
    for (;;) {
      if (!lock_.exchange(true, std::memory_order_acquire)) return;
      while (lock_.load(std::memory_order_relaxed)) std::this_thread::yield();
    }
  }
  void unlock() noexcept { lock_.store(false, std::memory_order_release); }
};
struct SpinGuard {
  Spinlock& l;
  explicit SpinGuard(Spinlock& lock) : l(lock) { l.lock(); }
  ~SpinGuard() { l.unlock(); }
  SpinGuard(const SpinGuard&) = delete;
  SpinGuard& operator=(const SpinGuard&) = delete;
};

// core/queue.h:142-196 (ConcurrentRingBuffer): one producer thread, one consumer thread, no lock. One slot stays empty,
// so CAP - 1 messages fit; push() yields while the ring is full, like the reference's.
template <class T, uint32_t CAP>
struct SpscRing {
  static_assert(CAP >= 2, "capacity");
  alignas(64) std::atomic<uint32_t> write_pos_{0};
  alignas(64) std::atomic<uint32_t> read_pos_{0};
  T data_[CAP];
  bool try_push(const T& v) noexcept {
    const uint32_t w = write_pos_.load(std::memory_order_relaxed);
    const uint32_t r = read_pos_.load(std::memory_order_acquire);
    const uint32_t nw = (w + 1) % CAP;
    if (nw == r) return false;
    data_[w] = v;
    write_pos_.store(nw, std::memory_order_release);
    return true;
  }
  void push(const T& v) noexcept {
    while (!try_push(v)) std::this_thread::yield();
  }
  bool pop(T& v) noexcept {
    const uint32_t w = write_pos_.load(std::memory_order_acquire);
    const uint32_t r = read_pos_.load(std::memory_order_relaxed);
    if (w == r) return false;
    v = data_[r];
    read_pos_.store((r + 1) % CAP, std::memory_order_release);
    return true;
  }
  bool empty() const noexcept {  // consumer side
    return write_pos_.load(std::memory_order_acquire) == read_pos_.load(std::memory_order_relaxed);
  }
};

// engine/vu_meter.h:17-40: the level only rises until the UI takes it
struct VULevel {
  std::atomic<float> level{0.0f};
  void push(float new_level) noexcept {  // audio thread
    float old_level = level.load(std::memory_order_relaxed);
    while (old_level < new_level &&
           !level.compare_exchange_weak(old_level, new_level, std::memory_order_release, std::memory_order_relaxed)) {
    }
  }
  float take() noexcept { return level.exchange(0.0f, std::memory_order_acq_rel); }  // UI thread (VUMeter::update)
  float peek() const noexcept { return level.load(std::memory_order_acquire); }
};

// core/timing.h:54-67: EMA (alpha = 0.25) of callback time / buffer time, shown by the UI as the engine load
struct PerformanceMeasurer {
  std::atomic<double> usage{0.0};
  void update(double duration, double target_duration) noexcept {
    const double percentage = duration / target_duration;
    const double old_usage = usage.load(std::memory_order_relaxed);
    usage.store(old_usage + 0.25 * (percentage - old_usage), std::memory_order_release);
  }
  double get_usage() const noexcept {
    const double u = usage.load(std::memory_order_acquire);
    return u < 0.0 ? 0.0 : (u > 1.0 ? 1.0 : u);
  }
};

// (mute ? 0 : volume) and pan law, engine/track.h:46-53
struct TrackParameterState {
  float volume_db = 0.0f;
  float volume = 0.0f;
  float pan = 0.0f;
  float pan_coeffs[2] = {0.0f, 0.0f};
  bool mute = false;
  bool solo = false;  // UI only (engine/track.h:52), kept by Engine::solo_track
};

struct PanningCoefficient {
  float left, right;
};
// core/panning_law.cpp:9-32 (ConstantPower_3db) and core/core_math.h:83-89 — host-only scalar math
PanningCoefficient calculate_panning_coefs(float pan);
// `while (steps < n && off < limit) { off = off + adv; steps++; }` — the position recurrence of dsp::Sampler::stream
// (dsp/sampler.cpp:103,209) over n event-free callbacks — in O(binades) exact integer steps; returns steps
uint32_t advance_rounded(double* off, double adv, uint32_t n, double limit);
float db_to_linear(float db);

struct AudioClip {  // engine/clip.h:39-45 + Clip time placement (:68-70)
  double min_time = 0, max_time = 0;  // beats
  double start_offset = 0;            // source frames
  double speed = 1.0;
  float gain = 1.0f;
  double fade_start = 0, fade_end = 0;  // beats (engine/clip.h:41-42); extension: the reference never reads them
  uint32_t sample_id = 0;
  uint32_t sample_rate = 0;
  bool internal_state_changed = false;
  bool deleted = false;  // Clip::mark_deleted (engine/clip.h): removed at the next update_clip_ordering
};

enum class EventType : uint8_t { None, StopSample, PlaySample };  // engine/event.h:11-15

struct AudioEvent {  // engine/event.h:66-74
  EventType type = EventType::None;
  uint32_t buffer_offset = 0;
  double time = 0, speed = 0;
  uint64_t sample_offset = 0;
  uint64_t clip_frame = 0;  // output frames since the clip's start (fade extension)
  const AudioClip* clip = nullptr;
};

struct Track {
  std::string name;
  std::vector<AudioClip*> clips;  // sorted by min_time, never overlapping
  std::vector<AudioClip*> graveyard;  // clips trimmed away by later edits, kept alive for voices still playing them
  // scheduler state (TrackEventState, engine/track.h:36-44)
  bool has_clip_idx = false;
  uint32_t clip_idx = 0;
  bool refresh_voice = false;
  bool partially_ended = false;
  std::vector<AudioEvent> audio_event_buffer;
  AudioEvent current_audio_event;
  // dsp::Sampler state (dsp/sampler.h:14-16); the device replays the same recurrence
  double playback_speed = 0, sample_offset = 0;
  uint64_t clip_frame = 0;  // output frames since the playing clip's start (fade extension)
  TrackParameterState ui_parameter_state, parameter_state;
  struct Msg {
    uint32_t id;
    double value;
  };
  SpscRing<Msg, 64> track_msg_queue;  // UI -> audio parameter messages (engine/track.h:131, capacity track.cpp:23)
  VULevel level[2];                   // VUMeter::level (engine/vu_meter.h:17): max since the UI last took it
  int32_t open_run = -1;             // index of this track's extendable run in the segment list
  // built-in effect chain (extension, see wbx.h): applied on the device between the clips and volume/pan
  wbx_effect_params effect_params{};
  bool effects_on = false, effects_dirty = false;
  // nullptr removes the chain. Not a message: call it under Engine::edit_lock() (Engine::set_track_effects does).
  void set_effects(const wbx_effect_params* params);
  // A plugin in the track's slot (engine/track.h:124). In the reference Track::process then points its write buffer at
  // the plugin's effect_buffer (track.cpp:600): the plugin runs BEFORE the clips are rendered, on an empty input, the
  // clips are rendered into effect_buffer afterwards and never reach the mix (track.cpp:645-724) — a track with a plugin
  // contributes only what the plugin itself writes. The slot here holds no third-party code: has_plugin reproduces the
  // reference's behaviour for a plugin that outputs silence (the sampler state still advances, the clips are dropped).
  bool has_plugin = false;
  ~Track();
  // UI thread, lock-free: messages the audio thread applies at its next callback
  void set_volume(float db);  // engine/track.cpp:47-57
  void set_pan(float pan);    // :59-68
  void set_mute(bool mute);   // :70-79
};

class Engine {
 public:
  // device_ordinal < 0 makes a scheduling-only engine (segment tables, no rendering: there is no CPU path).
  explicit Engine(int device_ordinal = 0);
  ~Engine();
  Engine(const Engine&) = delete;
  Engine& operator=(const Engine&) = delete;
  bool ok() const { return dev_ != nullptr; }
  const char* last_error() const;

  // Held by the audio thread for a whole callback and by every edit below (engine.h:89-94); UI code that touches
  // tracks / clips directly brackets it with edit_lock() / edit_unlock().
  Spinlock editor_lock;
  void edit_lock() { editor_lock.lock(); }
  void edit_unlock() { editor_lock.unlock(); }
  // engine load as the reference's control bar shows it (core/timing.h:54-67): callback time / buffer time, EMA
  PerformanceMeasurer perf_measurer;
  double cpu_usage() const { return perf_measurer.get_usage(); }

  // engine/engine.cpp:43-57, :24-30, :32-41
  int set_audio_channel_config(uint32_t input_channels, uint32_t output_channels, uint32_t buffer_size,
                               uint32_t sample_rate);
  void set_bpm(double bpm);
  void set_playhead_position(double beat);

  Track* add_track(const std::string& name);  // engine/engine.cpp:199-207
  // engine/engine.cpp:209-217, 228-243, 245-262, 1460-1464: mixer-side calls that change what the path sums — a track
  // leaves, the bus summation order changes, every other track is muted / unmuted, a clip's gain changes (the playing
  // voice reads it every callback, track.cpp:676,716)
  int delete_track(uint32_t slot);
  int set_track_effects(Track* track, const wbx_effect_params* params);  // Track::set_effects under the editor lock
  // Engine::add_plugin_to_track / delete_plugin_from_track (engine.cpp:1466-1538) reduced to the slot's effect on the
  // mix (see Track::has_plugin)
  int set_track_plugin(Track* track, bool present);
  int move_track(uint32_t from_slot, uint32_t to_slot);
  int solo_track(uint32_t slot);
  int set_clip_gain(Track* track, uint32_t clip_id, float gain);
  // Resident Sample (dsp/sample.h:18-28); returns the id clips refer to, or a negative wbx_status.
  int add_sample(int format, uint32_t channels, uint64_t frames, uint32_t sample_rate, const void* const* planar);
  // engine/engine.cpp:293-309 + add_to_cliplist (:409-461): overlapping clips are trimmed / split / deleted first
  // (Engine::reserve_track_region, :478-569).
  int add_audio_clip(Track* track, double min_time, double max_time, double start_offset, uint32_t sample_id,
                     double speed, float gain, double fade_start = 0.0, double fade_end = 0.0);
  // Clip editing that feeds the scheduler (engine/engine.cpp:336-407, engine/clip_edit.h): the clip must belong to the
  // track. Neighbours the new extent overlaps are trimmed / split / deleted (reserve_track_region); a clip moved or
  // shift/stretch-resized while it plays restarts at its new content offset on the next callback.
  int move_clip(Track* track, AudioClip* clip, double relative_pos);
  int resize_clip(Track* track, AudioClip* clip, double relative_pos, double resize_limit, double min_length, bool left_side,
                  bool shift = false, bool stretch = false);
  int delete_clip(Track* track, AudioClip* clip);
  int duplicate_clip(Track* track, const AudioClip* clip_to_duplicate, double min_time, double max_time);
  int delete_region(Track* track, double min, double max);  // engine.cpp:463-473: erase a time range of the track
  // convolution reverb (extension, see wbx.h): one impulse response per engine, used by chains with reverb_on
  int set_impulse_response(const float* h, uint32_t n_taps);
  void play();  // engine/engine.cpp:68-80
  void stop();  // :82-92

  // One audio callback: Engine::process(input_buffer, output_buffer, sample_rate), engine.cpp:1576-1654.
  template <class Buffer>
  int process(const Buffer& /*input_buffer*/, Buffer& output_buffer, double sample_rate) {
    if (output_buffer.n_samples != buffer_size_ || output_buffer.n_channels != out_channels_) return WBX_ERR_INVALID;
    return render(1, output_buffer.channel_buffers, nullptr, sample_rate);
  }

  // n_blocks consecutive callbacks in one device launch (offline bounce / throughput mode). out_channels[c]
  // receives n_blocks * buffer_size frames; peaks (optional) [n_blocks][n_tracks][2].
  int render(uint32_t n_blocks, float* const* out_channels, float* peaks, double sample_rate = 0.0);

  // Offline bounce / export (the reference has the dialog, ui/export_audio_dlg.cpp:44-200, and the
  // properties, engine/export_prop.h:14-45, but no driver): renders [start_beat, end_beat) from a stopped transport as
  // Engine::process would — play() at start_beat, one callback after the other — in chunks of chunk_blocks callbacks,
  // the clamped bus converted on the device to dst_format (core/audio_format_conv.cpp:5-106: WBX_FMT_I16 / I24_X8 / I32 /
  // F32 interleaved) and handed to `sink` in order; chunk i's copy-out and sink.write overlap chunk i+1's mix. The last
  // callback is cut at end_beat. Leaves the transport stopped at start_beat. Returns the frames written (>= 0) through
  // frames_out.
  struct BounceSink {
    virtual ~BounceSink() {}
    virtual int write(const void* data, size_t bytes) = 0;  // 0 or a negative wbx_status
  };
  int bounce(double start_beat, double end_beat, int dst_format, BounceSink& sink, uint32_t chunk_blocks = 256,
             uint64_t* frames_out = nullptr);

  // render() in two halves, for one thread driving several engines of a sharded setup (include/wbx_sharded.hpp):
  // render_begin = host schedule + wbx_submit; the caller then runs wbx_mix_sharded_phase(device(), 0..2) in lock step
  // over all engines; render_end = bus (rank 0; others pass nullptr) / peaks / levels back.
  int render_begin(uint32_t n_blocks, double sample_rate = 0.0);
  int render_end(float* const* out_channels, float* peaks);

  // Build the segment table for n_blocks callbacks and advance the transport, without touching the device
  // (render = schedule + wbx_render). Exposed for tests, multi-GPU sharding and benchmarks.
  int schedule(uint32_t n_blocks, double sample_rate = 0.0);
  const std::vector<wbx_segment>& segments() const { return segs_; }
  const std::vector<float>& track_gains() const { return gains_; }

  wbx_engine* device() const { return dev_; }
  uint32_t buffer_size() const { return buffer_size_; }
  uint32_t out_channels() const { return out_channels_; }
  uint32_t sample_rate() const { return sample_rate_; }
  std::vector<Track*> tracks;
  double ppq = 96.0;
  double playhead = 0, playhead_start = 0, sample_position = 0;
  bool playing = false;
  bool fast_forward = true;  // skip event-free callbacks in closed form while scheduling (same results)
  // dsp::ResamplerType (dsp/sampler.h:8-11): 0 = Linear, what Track::process hard-codes (track.cpp:692-697);
  // 1 = polyphase windowed sinc (extension, see WBX_SEG_POLYPHASE in wbx.h)
  int resampler_mode = 0;

 private:
  void process_event(Track& t, double start_time, double end_time, double sample_position_, double beat_duration,
                     double sample_rate, uint32_t buffer_size);
  void stream(Track& t, uint32_t track_index, uint32_t block, uint32_t num_samples, uint32_t buffer_offset);
  void track_block(Track& t, uint32_t track_index, uint32_t block, double sample_rate, double beat_duration,
                   double start_time, double end_time, double block_sample_position, bool currently_playing);
  int prepare(uint32_t n_blocks, double sample_rate);
  int schedule_locked(uint32_t n_blocks, double sample_rate);
  void meter(std::chrono::steady_clock::time_point t0, uint32_t n_blocks);
  void reindex_effects();
  void merge_levels();
  uint32_t quiet_blocks(const Track& t, uint32_t k, uint32_t K) const;
  void fill_fade(wbx_segment& s, const AudioClip* clip, uint64_t clip_frame) const;
  void reserve_track_region(Track& t, uint32_t first_clip, uint32_t last_clip, double min, double max,
                            const AudioClip* ignore_clip);
  void add_to_cliplist(Track* track, AudioClip* clip);
  void stream_run(Track& t, uint32_t track_index, uint32_t block, uint32_t q);
  std::vector<double> blk_start_, blk_end_, blk_spos_;  // per-callback transport of the current schedule()
  wbx_engine* dev_ = nullptr;
  bool host_only_ = false;
  uint32_t out_channels_ = 2, buffer_size_ = 512, sample_rate_ = 48000;
  double cur_sample_rate_ = 48000.0;  // sample rate of the schedule() in progress
  double beat_duration_ = 0.5;
  struct SampleInfo {
    uint64_t count;
    uint32_t rate;
  };
  std::vector<SampleInfo> samples_;
  std::vector<wbx_segment> segs_;
  std::vector<float> gains_;
  std::vector<float> levels_;
  std::string err_;
};

}  // namespace wbx
