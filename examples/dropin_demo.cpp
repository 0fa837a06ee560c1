// dropin_demo.cpp — the C++ side of the boundary without any Python: a stand-in for whitebox's audio I/O thread
// (engine/audio_io_pulseaudio.cpp:396-466) that owns two AudioBuffer<float>-shaped buffers and calls
//     engine.process(input_buffer, output_buffer, sample_rate)
// once per callback, exactly as the reference does, then bounces the same session offline in one launch.
//
// build: g++ -std=c++17 -Iinclude examples/dropin_demo.cpp -Lwhitebox_b200 -lwbx -Wl,-rpath,$PWD/whitebox_b200 -o dropin_demo
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "wbx_engine.hpp"

// Same members as wb::AudioBuffer<float> (core/audio_buffer.h:19-23); the real class can be passed instead.
struct AudioBufferF {
  uint32_t n_samples{};
  uint32_t n_channels{};
  float* internal_channel_buffers[16]{};
  float** channel_buffers{internal_channel_buffers};
  AudioBufferF(uint32_t samples, uint32_t channels) : n_samples(samples), n_channels(channels) {
    for (uint32_t c = 0; c < channels; c++) channel_buffers[c] = (float*)wbx_host_alloc(samples * sizeof(float));
  }
  ~AudioBufferF() {
    for (uint32_t c = 0; c < n_channels; c++) wbx_host_free(channel_buffers[c]);
  }
};

int main(int argc, char** argv) {
  const uint32_t n_tracks = argc > 1 ? (uint32_t)atoi(argv[1]) : 64;
  const uint32_t n_blocks = argc > 2 ? (uint32_t)atoi(argv[2]) : 32;
  const uint32_t B = 512, rate = 48000;

  wbx::Engine engine(0);
  if (!engine.ok()) {
    std::fprintf(stderr, "no sm_100 device: %s (there is no CPU path)\n", engine.last_error());
    return 2;
  }
  engine.set_audio_channel_config(0, 2, B, rate);  // start_audio_engine(), config.cpp:224
  engine.set_bpm(120.0);
  wbx_set_sum_mode(engine.device(), WBX_SUM_EXACT);  // reference summation order: realtime == bounce bit for bit

  const size_t frames = (size_t)(n_blocks + 2) * B;
  std::vector<float> l(frames), r(frames);
  for (uint32_t t = 0; t < n_tracks; t++) {
    for (size_t i = 0; i < frames; i++) {
      l[i] = 0.02f * std::sin(0.001f * (float)(i * (t + 1)));
      r[i] = 0.02f * std::cos(0.0013f * (float)(i * (t + 1)));
    }
    const void* planes[2] = {l.data(), r.data()};
    const int sid = engine.add_sample(WBX_FMT_F32, 2, frames, rate, planes);
    wbx::Track* track = engine.add_track("track");
    track->set_volume(-6.0f - (float)(t % 7));
    track->set_pan(-1.0f + 0.2f * (float)(t % 11));
    engine.add_audio_clip(track, 0.0, 1.0e6, 0.0, (uint32_t)sid, 1.0, 0.8f);
  }

  // realtime: one Engine::process per callback
  AudioBufferF input(B, 2), output(B, 2);
  std::vector<float> realtime[2];
  engine.play();
  for (uint32_t k = 0; k < n_blocks; k++) {
    const int rc = engine.process(input, output, (double)rate);
    if (rc != WBX_OK) {
      std::fprintf(stderr, "process failed: %d %s\n", rc, engine.last_error());
      return 1;
    }
    for (int c = 0; c < 2; c++) realtime[c].insert(realtime[c].end(), output.channel_buffers[c], output.channel_buffers[c] + B);
  }
  const float level0 = engine.tracks[0]->level[0].peek();

  // offline bounce of the same range: one launch for all callbacks
  engine.stop();
  engine.play();
  AudioBufferF bounce(n_blocks * B, 2);
  const int rc = engine.render(n_blocks, bounce.channel_buffers, nullptr);
  if (rc != WBX_OK) {
    std::fprintf(stderr, "render failed: %d %s\n", rc, engine.last_error());
    return 1;
  }
  size_t diff = 0;
  float peak = 0.f;
  for (int c = 0; c < 2; c++)
    for (size_t i = 0; i < (size_t)n_blocks * B; i++) {
      diff += std::memcmp(&realtime[c][i], &bounce.channel_buffers[c][i], 4) != 0;
      peak = std::fmax(peak, std::fabs(bounce.channel_buffers[c][i]));
    }
  std::printf("dropin_demo: %u tracks x %u callbacks, bus peak %.4f, track 0 VU %.4f, realtime vs bounce: %zu samples differ\n",
              n_tracks, n_blocks, peak, level0, diff);
  return diff == 0 && peak > 0.f ? 0 : 1;
}
