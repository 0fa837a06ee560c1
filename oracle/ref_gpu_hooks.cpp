// TEST / INTEGRATION INFRASTRUCTURE (oracle build). The reference-side binding of INTEGRATION.md section A: what the
// patched reference (oracle/patch_ref_gpu.py) calls instead of its sample loops, implemented with nothing but the C ABI of
// include/wbx.h. Together with the reference's own, otherwise unmodified engine this is libwbref_gpu.so — whitebox's
// Engine::process with the mixing hot path running on the GPU — which tests/test_gpu_parity.py holds to the golden
// vectors of the CPU reference bit for bit.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <unordered_map>
#include <vector>

#include "../include/wbx.h"
#include "dsp/sample.h"
#include "dsp/sampler.h"
#include "engine/engine.h"
#include "engine/track.h"
#include "wbx_gpu_hooks.h"

namespace {

struct Binding {  // one per wb::Engine
  wbx_engine* dev = nullptr;
  uint32_t channels = 0, block = 0, rate = 0;
  std::unordered_map<const wb::Sample*, uint32_t> samples;  // resident copies of the reference's Sample objects
  std::unordered_map<const wb::Track*, uint32_t> index;     // Track* -> position in Engine::tracks, per callback
  std::vector<wbx_segment> segs;
  std::vector<float> gains, peaks;
};

std::unordered_map<const wb::Engine*, Binding*> g_bindings;
Binding* g_current = nullptr;  // Engine::process runs on one audio thread

[[noreturn]] void die(const Binding* b, const char* what, int rc) {
  std::fprintf(stderr, "wbx_gpu: %s failed with status %d: %s\n", what, rc, b && b->dev ? wbx_last_error(b->dev) : "");
  std::abort();  // Engine::process has no error path; a broken device must not produce silent garbage
}

Binding* binding_of(const wb::Engine* engine) {
  auto it = g_bindings.find(engine);
  if (it != g_bindings.end()) return it->second;
  Binding* b = new Binding();
  const char* ord = std::getenv("WBX_GPU_DEVICE");
  int rc = wbx_create(&b->dev, ord ? std::atoi(ord) : 0);
  if (rc) die(nullptr, "wbx_create", rc);
  if ((rc = wbx_set_sum_mode(b->dev, WBX_SUM_EXACT))) die(b, "wbx_set_sum_mode", rc);  // AudioBuffer::mix order
  g_bindings[engine] = b;
  return b;
}

}  // namespace

namespace wbx_gpu {

void begin(wb::Engine* engine, wb::AudioBuffer<float>& out, double sample_rate) {
  Binding* b = binding_of(engine);
  if (b->channels != out.n_channels || b->block != out.n_samples || b->rate != (uint32_t)sample_rate) {
    // first callback, or Engine::set_audio_channel_config ran again (config.cpp:198-232)
    const int rc = wbx_configure(b->dev, out.n_channels, out.n_samples, (uint32_t)sample_rate);
    if (rc) die(b, "wbx_configure", rc);
    b->channels = out.n_channels, b->block = out.n_samples, b->rate = (uint32_t)sample_rate;
  }
  const uint32_t n = (uint32_t)engine->tracks.size();
  b->index.clear();
  for (uint32_t i = 0; i < n; i++) b->index[engine->tracks[i]] = i;
  b->segs.clear();
  b->gains.assign((size_t)n * 2, 0.0f);
  g_current = b;
}

void stream(wb::Track* track, wb::dsp::Sampler& sampler, wb::Sample* sample, uint32_t num_samples, uint32_t buffer_offset,
            float gain, bool dropped) {
  Binding* b = g_current;
  if (!dropped && num_samples != 0 && sampler.sample_offset_ < (double)sample->count) {
    auto it = b->samples.find(sample);
    if (it == b->samples.end()) {  // first use: make the Sample resident (planar channel pointers, dsp/sample.h:18-28)
      uint32_t id = 0;
      std::vector<const void*> planes(sample->channels);
      for (uint32_t c = 0; c < sample->channels; c++) planes[c] = sample->sample_data[c];
      const int rc = wbx_sample_upload(b->dev, (int)sample->format, sample->channels, sample->count, sample->sample_rate,
                                       planes.data(), &id);
      if (rc) die(b, "wbx_sample_upload", rc);
      it = b->samples.emplace(sample, id).first;
    }
    wbx_segment s{};
    s.track = b->index.at(track);
    s.block = 0;
    s.n_blocks = 1;
    s.dst_offset = buffer_offset;
    s.length = num_samples;
    s.sample_id = it->second;
    s.src_pos = sampler.sample_offset_;
    s.speed = sampler.playback_speed_;
    s.gain = gain;
    b->segs.push_back(s);
  }
  // sample_offset_ bookkeeping (sampler.cpp:99-104,209) by the reference's own code, on zero output channels
  sampler.stream(sample, 0, num_samples, buffer_offset, gain, nullptr);
}

void track_gains(wb::Track* track, float gain_left, float gain_right) {
  Binding* b = g_current;
  const uint32_t i = b->index.at(track);
  b->gains[2 * i] = gain_left;
  b->gains[2 * i + 1] = gain_right;
}

void render(wb::Engine* engine, wb::AudioBuffer<float>& out) {
  Binding* b = g_current;
  const uint32_t n = (uint32_t)engine->tracks.size();
  int rc = wbx_set_track_count(b->dev, n);
  if (rc) die(b, "wbx_set_track_count", rc);
  b->peaks.assign((size_t)n * 2 + 2, 0.0f);
  rc = wbx_render_levels(b->dev, b->segs.data(), (uint32_t)b->segs.size(), b->gains.data(), 1, out.channel_buffers,
                         n ? b->peaks.data() : nullptr, nullptr);
  if (rc) die(b, "wbx_render_levels", rc);
  for (uint32_t i = 0; i < n; i++)  // VUMeter::push_samples' CAS-max (vu_meter.h:25-29) with the device's block peaks
    for (uint32_t c = 0; c < 2; c++) {
      std::atomic<float>& level = engine->tracks[i]->level_meter[c].level;
      const float new_level = b->peaks[2 * i + c];
      float old_level = level.load(std::memory_order_relaxed);
      while (old_level < new_level &&
             !level.compare_exchange_weak(old_level, new_level, std::memory_order_release, std::memory_order_relaxed)) {
      }
    }
}

void release(wb::Engine* engine) {
  auto it = g_bindings.find(engine);
  if (it == g_bindings.end()) return;
  if (g_current == it->second) g_current = nullptr;
  wbx_destroy(it->second->dev);
  delete it->second;
  g_bindings.erase(it);
}

}  // namespace wbx_gpu
