// wbx_gpu_hooks.h — the four calls the patched reference makes instead of its sample loops (INTEGRATION.md section A,
// oracle/patch_ref_gpu.py). TEST / INTEGRATION INFRASTRUCTURE: included only by the patched copies of the reference's
// engine.cpp / track.cpp under oracle/_ref/patched/ and implemented by oracle/ref_gpu_hooks.cpp on top of include/wbx.h.
#pragma once
#include <cstdint>

#include "core/audio_buffer.h"  // the reference's own headers (include path of the oracle build)
#include "dsp/sample.h"
#include "dsp/sampler.h"

namespace wb {
struct Engine;
struct Track;
}  // namespace wb

namespace wbx_gpu {
// Engine::process, where output_buffer.clear() stood (engine.cpp:1598): a new callback begins
void begin(wb::Engine* engine, wb::AudioBuffer<float>& output_buffer, double sample_rate);
// Track::process, where dsp::Sampler::stream was called (track.cpp:678,718): one wbx_segment; `dropped` = the track has a
// plugin, i.e. the reference would render this call into effect_buffer and never mix it. The sampler's position
// bookkeeping is still done by the reference's own Sampler::stream (on zero channels: it touches no sample).
void stream(wb::Track* track, wb::dsp::Sampler& sampler, wb::Sample* sample, uint32_t num_samples, uint32_t buffer_offset,
            float gain, bool dropped);
// Track::process, where the apply_gain + VUMeter::push_samples loop stood (track.cpp:728-733)
void track_gains(wb::Track* track, float gain_left, float gain_right);
// Engine::process, where the clamp loop stood (engine.cpp:1627-1636): one wbx_render_levels = Sampler::stream, clip gain,
// volume * pan, VU peaks, bus sum, clamp for the whole callback; fills output_buffer and the tracks' level meters
void render(wb::Engine* engine, wb::AudioBuffer<float>& output_buffer);
// the engine is going away: drop its device engine and resident samples
void release(wb::Engine* engine);
}  // namespace wbx_gpu
