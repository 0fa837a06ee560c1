// sharded_demo.cpp — one process, several shards: wbx::ShardedEngine (include/wbx_sharded.hpp) behind the same
// engine.process(input_buffer, output_buffer, sample_rate) call as the reference's audio thread
// (engine/audio_io_pulseaudio.cpp:411), checked against one wbx::Engine holding every track.
// Shards use GPUs 0..n-1 when the box has several, otherwise they share GPU 0 (the exchange is the same code).
//
// build: g++ -std=c++17 -Iinclude examples/sharded_demo.cpp -Lwhitebox_b200 -lwbx -Wl,-rpath,$PWD/whitebox_b200 -o sharded_demo
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "wbx_sharded.hpp"

struct AudioBufferF {  // members of wb::AudioBuffer<float> (core/audio_buffer.h:19-23)
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
  const uint32_t n_tracks = argc > 1 ? (uint32_t)atoi(argv[1]) : 48;
  const uint32_t n_blocks = argc > 2 ? (uint32_t)atoi(argv[2]) : 12;
  const int gpus = argc > 3 ? atoi(argv[3]) : 1;
  const uint32_t B = 512, rate = 48000, W = 2;
  std::vector<int> devices;
  for (uint32_t r = 0; r < W; r++) devices.push_back(gpus > 1 ? (int)(r % gpus) : 0);

  wbx::ShardedEngine sharded(devices);
  wbx::Engine single(0);
  if (!sharded.ok() || !single.ok()) {
    std::fprintf(stderr, "no sm_100 device (there is no CPU path)\n");
    return 2;
  }
  if (sharded.set_audio_channel_config(0, 2, B, rate, 1) || single.set_audio_channel_config(0, 2, B, rate)) return 3;
  sharded.set_bpm(120.0);
  single.set_bpm(120.0);

  const uint64_t frames = (uint64_t)(n_blocks + 4) * B;
  std::vector<float> l(frames), r(frames);
  for (uint32_t t = 0; t < n_tracks; t++) {
    for (uint64_t i = 0; i < frames; i++) {
      l[i] = 0.2f * std::sin(0.001f * (float)(t + 1) * (float)i);
      r[i] = 0.2f * std::cos(0.0013f * (float)(t + 1) * (float)i);
    }
    const void* planes[2] = {l.data(), r.data()};
    auto ref = sharded.add_track("t");
    ref.track->set_volume(-6.0f - (float)(t % 5));
    ref.track->set_pan(-1.0f + 0.25f * (float)(t % 9));
    const int sid = sharded.add_sample(ref.shard, WBX_FMT_F32, 2, frames, rate, planes);
    if (sid < 0 || sharded.add_audio_clip(ref, 0.0, 1e9, 0.0, (uint32_t)sid, 1.0, 0.8f)) return 4;
    wbx::Track* st = single.add_track("t");
    st->set_volume(-6.0f - (float)(t % 5));
    st->set_pan(-1.0f + 0.25f * (float)(t % 9));
    const int sid1 = single.add_sample(WBX_FMT_F32, 2, frames, rate, planes);
    if (sid1 < 0 || single.add_audio_clip(st, 0.0, 1e9, 0.0, (uint32_t)sid1, 1.0, 0.8f)) return 4;
  }
  sharded.play();
  single.play();

  AudioBufferF in(B, 2), out_a(B, 2), out_b(B, 2);
  double worst = 0.0, peak = 0.0;
  for (uint32_t k = 0; k < n_blocks; k++) {  // the audio thread's loop
    if (int rc = sharded.process(in, out_a, (double)rate)) {
      std::fprintf(stderr, "sharded process failed: %d %s\n", rc, sharded.last_error());
      return 5;
    }
    if (int rc = single.process(in, out_b, (double)rate)) {
      std::fprintf(stderr, "single process failed: %d %s\n", rc, single.last_error());
      return 5;
    }
    for (uint32_t c = 0; c < 2; c++)
      for (uint32_t i = 0; i < B; i++) {
        const double d = std::fabs((double)out_a.channel_buffers[c][i] - (double)out_b.channel_buffers[c][i]);
        if (d > worst) worst = d;
        if (std::fabs((double)out_b.channel_buffers[c][i]) > peak) peak = std::fabs((double)out_b.channel_buffers[c][i]);
      }
  }
  std::printf("%u tracks over %u shards, %u callbacks: max |sharded - single| = %.3g (peak %.3g)\n", n_tracks, W, n_blocks, worst, peak);
  if (!(peak > 0.05) || worst > 1e-5 * peak) {
    std::printf("MISMATCH\n");
    return 1;
  }
  std::printf("sharded == single within tolerance\n");
  return 0;
}
