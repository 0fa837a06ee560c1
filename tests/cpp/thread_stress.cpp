// Two-thread stress test of the host engine's threading contract (include/wbx_engine.hpp): a UI thread hammers
// Track::set_volume / set_pan / set_mute (lock-free SPSC ring), Engine::add_audio_clip / set_bpm / set_track_effects
// (editor lock) and takes VU levels, while the audio thread runs callbacks. Built by tests/test_host_cpu.py with
// -fsanitize=thread against whitebox_b200/csrc/wbx_host.cpp alone: the engine is scheduling-only (device < 0), so the
// device ABI is never reached — the stubs below only satisfy the linker. Every callback must equal SOME serialised
// schedule: the volume a callback uses is one the UI wrote, and later callbacks never see an older one.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "../../include/wbx_engine.hpp"

extern "C" {
int wbx_create(wbx_engine**, int) { return WBX_ERR_NO_DEVICE; }
int wbx_destroy(wbx_engine*) { return WBX_OK; }
const char* wbx_last_error(const wbx_engine*) { return ""; }
int wbx_configure(wbx_engine*, uint32_t, uint32_t, uint32_t) { return WBX_ERR_NO_DEVICE; }
int wbx_set_track_count(wbx_engine*, uint32_t) { return WBX_ERR_NO_DEVICE; }
int wbx_sample_upload(wbx_engine*, int, uint32_t, uint64_t, uint32_t, const void* const*, uint32_t*) { return WBX_ERR_NO_DEVICE; }
int wbx_effects_design(const wbx_effect_params*, uint32_t, wbx_effects*) { return WBX_ERR_NO_DEVICE; }
int wbx_set_impulse_response(wbx_engine*, const float*, uint32_t) { return WBX_ERR_NO_DEVICE; }
int wbx_set_track_effects(wbx_engine*, uint32_t, const wbx_effects*) { return WBX_ERR_NO_DEVICE; }
int wbx_render_levels(wbx_engine*, const wbx_segment*, uint32_t, const float*, uint32_t, float* const*, float*, float*) { return WBX_ERR_NO_DEVICE; }
int wbx_submit(wbx_engine*, const wbx_segment*, uint32_t, const float*, uint32_t) { return WBX_ERR_NO_DEVICE; }
int wbx_fetch(wbx_engine*, float* const*, float*) { return WBX_ERR_NO_DEVICE; }
int wbx_fetch_levels(wbx_engine*, float*) { return WBX_ERR_NO_DEVICE; }
int wbx_mix(wbx_engine*, uint32_t) { return WBX_ERR_NO_DEVICE; }
int wbx_bounce_begin(wbx_engine*, int) { return WBX_ERR_NO_DEVICE; }
int wbx_bounce_push(wbx_engine*) { return WBX_ERR_NO_DEVICE; }
int wbx_bounce_pop(wbx_engine*, const void**, size_t*) { return WBX_ERR_NO_DEVICE; }
int wbx_synchronize(wbx_engine*) { return WBX_ERR_NO_DEVICE; }
}

int main(int argc, char** argv) {
  const int callbacks = argc > 1 ? atoi(argv[1]) : 10000;
  wbx::Engine eng(-1);
  if (eng.set_audio_channel_config(0, 2, 128, 48000) != WBX_OK) return 2;
  eng.set_bpm(120.0);
  const int n_tracks = 8;
  for (int t = 0; t < n_tracks; t++) {
    wbx::Track* tr = eng.add_track("t");
    const int sid = eng.add_sample(WBX_FMT_F32, 2, 1u << 22, 48000, nullptr);
    if (sid < 0) return 3;
    if (eng.add_audio_clip(tr, 0.0, 1.0e6, 0.0, (uint32_t)sid, 1.0, 1.0f) != WBX_OK) return 4;
  }
  eng.play();

  std::atomic<bool> stop{false};
  std::atomic<int> written{0};          // index of the last volume the UI thread wrote to track 0
  std::vector<float> volumes(1 << 20);  // volumes[i] = linear volume of write i (monotonically increasing)
  for (size_t i = 0; i < volumes.size(); i++) volumes[i] = wbx::db_to_linear(-60.0f + 60.0f * (float)i / (float)volumes.size());

  std::thread ui([&] {
    int i = 0;
    uint32_t edits = 0;
    while (!stop.load(std::memory_order_acquire) && i + 1 < (int)volumes.size()) {
      i++;
      eng.tracks[0]->set_volume(-60.0f + 60.0f * (float)i / (float)volumes.size());
      written.store(i, std::memory_order_release);
      eng.tracks[1 + i % (n_tracks - 1)]->set_pan(-1.0f + 2.0f * (float)(i % 97) / 96.0f);
      eng.tracks[1 + (i / 3) % (n_tracks - 1)]->set_mute((i & 64) != 0);
      if (i % 50 == 0) {  // edits under the editor lock: a clip on top of the playing one, tempo, a chain, levels
        wbx::Track* tr = eng.tracks[1 + (edits % (n_tracks - 1))];
        const double at = 0.5 + 0.25 * (double)(edits % 200);
        eng.add_audio_clip(tr, at, at + 0.2, 0.0, edits % n_tracks, 1.0, 0.7f);
        if (edits % 7 == 0) eng.set_bpm(100.0 + (double)(edits % 40));
        if (edits % 11 == 0) {
          wbx_effect_params fx{};
          for (int b = 0; b < 4; b++) fx.eq_freq[b] = 100.0f * (float)(b + 1), fx.eq_q[b] = 0.7f, fx.eq_gain_db[b] = 1.0f;
          eng.set_track_effects(tr, (edits % 22 == 0) ? &fx : nullptr);
        }
        for (int t = 0; t < n_tracks; t++) (void)eng.tracks[t]->level[t & 1].take();
        (void)eng.cpu_usage();
        edits++;
      }
    }
  });

  int errors = 0, last_index = 0;
  size_t total_segments = 0;
  const int min_writes = argc > 2 ? atoi(argv[2]) : 20000;
  const auto t_end = std::chrono::steady_clock::now() + std::chrono::seconds(60);
  int k = 0;
  for (; (k < callbacks || written.load(std::memory_order_acquire) < min_writes) && std::chrono::steady_clock::now() < t_end; k++) {
    if (eng.schedule(1) != WBX_OK) {
      errors++;
      break;
    }
    total_segments += eng.segments().size();
    // the gain this callback uses for track 0 (pan 0: both coefficients 1) is volume * 1
    const int hi = written.load(std::memory_order_acquire);
    const float g = eng.track_gains()[0];
    // find which write it is: volumes[] is increasing, so search at or after the last one seen
    int idx = -1;
    if (g == wbx::db_to_linear(0.0f) && last_index == 0) {
      idx = 0;  // the constructor's default (0 dB) before the first message was consumed
    } else {
      for (int j = last_index; j <= hi + 1 && j < (int)volumes.size(); j++)
        if (volumes[j] == g) {
          idx = j;
          break;
        }
    }
    if (idx < 0) {
      if (errors < 5) fprintf(stderr, "callback %d: gain %.9g is not a volume the UI wrote in [%d, %d]\n", k, g, last_index, hi);
      errors++;
    } else {
      last_index = idx;
    }
    for (int t = 0; t < n_tracks; t++) eng.tracks[t]->level[t & 1].push((float)(k % 100) * 0.01f);  // VUMeter::push_samples
    eng.perf_measurer.update(0.1, 2.67);
  }
  stop.store(true, std::memory_order_release);
  ui.join();
  printf("callbacks %d, segments %zu, UI volume writes %d (last seen by the audio thread: %d), errors %d\n", k,
         total_segments, written.load(), last_index, errors);
  if (k < callbacks || written.load() < min_writes) {
    fprintf(stderr, "timed out before %d callbacks and %d UI writes\n", callbacks, min_writes);
    return 5;
  }
  return errors ? 1 : 0;
}
