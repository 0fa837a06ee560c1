// wbx_sharded.hpp — one process, several GPUs: the reference's engine call shape in front of W wbx::Engine shards.
//
// Tracks are independent until AudioBuffer::mix adds them into the bus (engine/engine.cpp:1600-1617), so track i lives on
// shard i % W together with the samples its clips use; every callback each shard schedules and mixes its own tracks and
// the bus sum + clamp (engine.cpp:1627-1636) run as the peer-memory exchange of include/wbx.h ("sharded render"), driven
// here phase by phase from the one audio thread. Drop-in use is the same as wbx::Engine: where audio_io_* calls
// engine->process(input_buffer, output_buffer, sample_rate) (audio_io_pulseaudio.cpp:411, audio_io_wasapi.cpp:708).
// Header-only on top of wbx_engine.hpp / wbx.h.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "wbx_engine.hpp"

namespace wbx {

class ShardedEngine {
 public:
  struct TrackRef {
    uint32_t shard = 0;
    Track* track = nullptr;
  };

  // One shard per entry of device_ordinals (an ordinal may repeat: several shards on one GPU, which is how the
  // single-GPU tests exercise the exchange).
  explicit ShardedEngine(const std::vector<int>& device_ordinals) {
    for (int d : device_ordinals) shards_.emplace_back(new Engine(d));
  }
  bool ok() const {
    for (auto& s : shards_)
      if (!s->ok()) return false;
    return !shards_.empty();
  }
  uint32_t world() const { return (uint32_t)shards_.size(); }
  Engine& shard(uint32_t i) { return *shards_[i]; }
  const char* last_error() const { return err_shard_ < shards_.size() ? shards_[err_shard_]->last_error() : ""; }

  // Engine::set_audio_channel_config + the exchange set-up; max_blocks = the largest n_blocks render() will be given.
  int set_audio_channel_config(uint32_t in_channels, uint32_t out_channels, uint32_t buffer_size, uint32_t sample_rate,
                               uint32_t max_blocks = 1) {
    std::vector<wbx_engine*> devs;
    for (uint32_t r = 0; r < world(); r++) {
      if (int rc = fail_on(r, shards_[r]->set_audio_channel_config(in_channels, out_channels, buffer_size, sample_rate))) return rc;
      if (int rc = fail_on(r, wbx_shard_init(shards_[r]->device(), r, world(), max_blocks, nullptr))) return rc;
      devs.push_back(shards_[r]->device());
    }
    for (uint32_t r = 0; r < world(); r++)
      if (int rc = fail_on(r, wbx_shard_connect_local(shards_[r]->device(), devs.data()))) return rc;
    max_blocks_ = max_blocks;
    return WBX_OK;
  }
  void set_bpm(double bpm) {
    for (auto& s : shards_) s->set_bpm(bpm);
  }
  void set_playhead_position(double beat) {
    for (auto& s : shards_) s->set_playhead_position(beat);
  }
  void play() {
    for (auto& s : shards_) s->play();
  }
  void stop() {
    for (auto& s : shards_) s->stop();
  }

  // Engine::add_track: track i of the session lives on shard i % W
  TrackRef add_track(const std::string& name) {
    TrackRef t;
    t.shard = n_tracks_++ % world();
    t.track = shards_[t.shard]->add_track(name);
    return t;
  }
  // A Sample is resident on the shard whose tracks play it; returns that shard's sample id (or a negative wbx_status).
  int add_sample(uint32_t shard, int format, uint32_t channels, uint64_t frames, uint32_t sample_rate,
                 const void* const* planar) {
    return shards_[shard]->add_sample(format, channels, frames, sample_rate, planar);
  }
  int add_audio_clip(const TrackRef& t, double min_time, double max_time, double start_offset, uint32_t sample_id,
                     double speed, float gain, double fade_start = 0.0, double fade_end = 0.0) {
    return shards_[t.shard]->add_audio_clip(t.track, min_time, max_time, start_offset, sample_id, speed, gain, fade_start, fade_end);
  }

  // One audio callback: Engine::process(input_buffer, output_buffer, sample_rate), engine.cpp:1576-1654.
  template <class Buffer>
  int process(const Buffer& /*input_buffer*/, Buffer& output_buffer, double sample_rate) {
    if (output_buffer.n_samples != shards_[0]->buffer_size() || output_buffer.n_channels != shards_[0]->out_channels())
      return WBX_ERR_INVALID;
    return render(1, output_buffer.channel_buffers, sample_rate);
  }

  // n_blocks consecutive callbacks; out_channels[c] receives n_blocks * buffer_size clamped frames of the master bus.
  int render(uint32_t n_blocks, float* const* out_channels, double sample_rate = 0.0) {
    if (n_blocks > max_blocks_) return WBX_ERR_INVALID;
    // page-locked output channels (wbx_host_alloc): every owner shard stores its slice of the master bus into them itself
    const uint64_t frames = (uint64_t)n_blocks * shards_[0]->buffer_size();
    if (out_channels && (out_channels[0] != host_out_[0] || out_channels[1 % shards_[0]->out_channels()] != host_out_[1] ||
                         frames > host_frames_)) {
      bool ok = true;
      for (uint32_t r = 0; r < world(); r++) ok = wbx_shard_set_host_output(shards_[r]->device(), out_channels, frames) == WBX_OK && ok;
      if (!ok)  // pageable channels: rank 0 copies the master bus instead
        for (uint32_t r = 0; r < world(); r++) wbx_shard_set_host_output(shards_[r]->device(), nullptr, 0);
      host_out_[0] = out_channels[0];
      host_out_[1] = out_channels[1 % shards_[0]->out_channels()];
      host_frames_ = frames;
    }
    for (uint32_t r = 0; r < world(); r++)
      if (int rc = fail_on(r, shards_[r]->render_begin(n_blocks, sample_rate))) return rc;
    for (int phase = 0; phase < 3; phase++)  // every engine finishes a phase's enqueue before any starts the next
      for (uint32_t r = 0; r < world(); r++)
        if (int rc = fail_on(r, wbx_mix_sharded_phase(shards_[r]->device(), phase))) return rc;
    for (uint32_t r = 0; r < world(); r++)
      if (int rc = fail_on(r, shards_[r]->render_end(r == 0 ? out_channels : nullptr, nullptr))) return rc;
    return WBX_OK;
  }

 private:
  int fail_on(uint32_t r, int rc) {
    if (rc) err_shard_ = r;
    return rc;
  }
  std::vector<std::unique_ptr<Engine>> shards_;
  uint32_t n_tracks_ = 0, max_blocks_ = 0, err_shard_ = 0;
  float* host_out_[2] = {nullptr, nullptr};  // output channels the shards were last pointed at
  uint64_t host_frames_ = 0;
};

}  // namespace wbx
