// TEST INFRASTRUCTURE ONLY (oracle build).
//
// The eleven non-arithmetic symbols the reference's engine translation units reference but whose real
// definitions live in files that need libsndfile / Vulkan / VST3 / leveldb / midi-parser, none of which is
// on the mixing hot path. Everything that touches a sample VALUE is compiled from the
// reference's own sources; these stubs only allocate memory or return "absent".
#include <cstdlib>
#include <cstring>
#include <utility>

#include "core/midi_file.h"
#include "dsp/sample.h"
#include "gfx/renderer.h"
#include "gfx/waveform_visual.h"
#include "plughost/plugin_manager.h"

namespace wb {

// --- dsp/sample.cpp stand-ins (real file needs libsndfile/vorbis/dr_mp3 for decoding only) -------------
Sample::Sample(AudioFormat format, uint32_t sample_rate) : format(format), sample_rate(sample_rate) {}

Sample::Sample(Sample&& other) noexcept
    : name(std::move(other.name)),
      path(std::move(other.path)),
      format(std::exchange(other.format, AudioFormat::Unknown)),
      channels(std::exchange(other.channels, 0)),
      sample_rate(std::exchange(other.sample_rate, 0)),
      count(std::exchange(other.count, 0)),
      sample_data(std::move(other.sample_data)) {}

Sample::~Sample() {
  for (auto p : sample_data)
    std::free(p);
}

// Fresh zeroed channels of n frames + Sample::sample_padding zero frames — the shape load_file produces
// (dsp/sample.cpp:127,140). The harness fills the first n frames afterwards.
void Sample::resize(size_t n, uint32_t new_channels, bool) {
  for (auto p : sample_data)
    std::free(p);
  sample_data.resize(new_channels);
  // The sampler reads AudioFormat::I24 sources through int32_t pointers (dsp/sampler.cpp:121-132,171-181),
  // i.e. 24-bit data widened to a 4-byte container (dsp/sample.cpp:20), so size I24 channels as 4 bytes.
  size_t elem = format == AudioFormat::I24 ? 4 : get_audio_format_size(format);
  size_t bytes = (n + sample_padding) * elem;
  for (uint32_t c = 0; c < new_channels; c++)
    sample_data[c] = (std::byte*)std::calloc(1, bytes);
  channels = new_channels;
  count = n;
}

std::optional<Sample> Sample::load_file(const std::filesystem::path&) noexcept { return {}; }

// --- gfx: the mip-map builder (gfx/waveform_visual.cpp) is compiled from the reference unmodified; it uploads
// its results through the abstract GPURenderer (gfx/renderer.h:165-209). The Vulkan renderer is replaced by a
// host-memory one: create_buffer mallocs, begin_upload_data hands the memory out. No arithmetic here.
struct HostBuffer : GPUBuffer {
  void* mem = nullptr;
};
struct HostRenderer : GPURenderer {
  GPUBuffer* create_buffer(GPUBufferUsageFlags usage, size_t buffer_size, bool, size_t, const void*) override {
    HostBuffer* b = new HostBuffer();
    b->usage = usage;
    b->size = buffer_size;
    b->mem = std::calloc(1, buffer_size ? buffer_size : 1);
    return b;
  }
  GPUTexture* create_texture(GPUTextureUsageFlags, GPUFormat, uint32_t, uint32_t, bool, uint32_t, uint32_t, const void*) override { return nullptr; }
  GPUPipeline* create_pipeline(const GPUPipelineDesc&) override { return nullptr; }
  void destroy_buffer(GPUBuffer* buffer) override {
    HostBuffer* b = static_cast<HostBuffer*>(buffer);
    if (b) std::free(b->mem);
    delete b;
  }
  void destroy_texture(GPUTexture*) override {}
  void destroy_pipeline(GPUPipeline*) override {}
  void add_viewport(ImGuiViewport*) override {}
  void remove_viewport(ImGuiViewport*) override {}
  void resize_viewport(ImGuiViewport*, ImVec2) override {}
  void end_frame() override {}
  void present() override {}
  void* map_buffer(GPUBuffer* buffer) override { return static_cast<HostBuffer*>(buffer)->mem; }
  void unmap_buffer(GPUBuffer*) override {}
  void* begin_upload_data(GPUBuffer* buffer, size_t) override { return static_cast<HostBuffer*>(buffer)->mem; }
  void end_upload_data() override {}
  void begin_render(GPUTexture*, const ImVec4&) override {}
  void end_render() override {}
  void set_shader_parameter(size_t, const void*) override {}
  void flush_state() override {}
};
// the three non-pure virtuals of GPURenderer live in gfx/renderer.cpp (ImGui/SDL code): empty stand-ins
bool GPURenderer::init(SDL_Window*) { return true; }
void GPURenderer::shutdown() {}
void GPURenderer::begin_frame() {}
static HostRenderer g_host_renderer;
GPURenderer* g_renderer = &g_host_renderer;
void* wbref_buffer_memory(GPUBuffer* b) { return static_cast<HostBuffer*>(b)->mem; }

// --- plughost / midi file stand-ins ----------------------------------------------------------------------
// A plugin that is valid and does nothing: process() leaves the track's mixing buffer as Engine::process cleared it, which
// is all it takes to pin what a plugin's PRESENCE does to the path (Track::process, track.cpp:600,645-724: the clips are
// rendered into effect_buffer and never mixed). pm_open_plugin returns one once wbref_enable_null_plugin(1) was called.
PluginInterface::PluginInterface(uint64_t module_hash, PluginFormat format) : module_hash(module_hash), format(format) {}
PluginResult PluginInterface::render_ui() { return PluginResult::Ok; }
namespace {
struct NullPlugin final : PluginInterface {
  NullPlugin() : PluginInterface(0, PluginFormat::Native) { is_plugin_valid = true; }
  PluginResult init() override { return PluginResult::Ok; }
  PluginResult shutdown() override { return PluginResult::Ok; }
  uint32_t get_param_count() const override { return 0; }
  uint32_t get_audio_bus_count(bool) const override { return 0; }
  uint32_t get_event_bus_count(bool) const override { return 0; }
  uint32_t get_latency_samples() const override { return 0; }
  uint32_t get_tail_samples() const override { return 0; }
  const char* get_name() const override { return "null"; }
  PluginResult get_plugin_param_info(uint32_t, PluginParamInfo*) const override { return PluginResult::Unsupported; }
  PluginResult get_audio_bus_info(bool, uint32_t, PluginAudioBusInfo*) const override { return PluginResult::Unsupported; }
  PluginResult get_event_bus_info(bool, uint32_t, PluginEventBusInfo*) const override { return PluginResult::Unsupported; }
  PluginResult activate_audio_bus(bool, uint32_t, bool) override { return PluginResult::Ok; }
  PluginResult activate_event_bus(bool, uint32_t, bool) override { return PluginResult::Ok; }
  PluginResult init_processing(PluginProcessingMode, uint32_t, double) override { return PluginResult::Ok; }
  PluginResult start_processing() override { return PluginResult::Ok; }
  PluginResult stop_processing() override { return PluginResult::Ok; }
  void transfer_param(uint32_t, double) override {}
  PluginResult process(PluginProcessInfo&) override { return PluginResult::Ok; }
  bool has_view() const override { return false; }
  bool has_window_attached() const override { return false; }
  PluginResult get_view_size(uint32_t*, uint32_t*) const override { return PluginResult::Unsupported; }
  PluginResult attach_window(SDL_Window*) override { return PluginResult::Unsupported; }
  PluginResult detach_window() override { return PluginResult::Ok; }
};
bool g_null_plugin = false;
}  // namespace
void wbref_enable_null_plugin(bool on) { g_null_plugin = on; }
PluginInterface* pm_open_plugin(PluginUID) { return g_null_plugin ? new NullPlugin() : nullptr; }
void pm_close_plugin(PluginInterface* p) { delete p; }
bool load_notes_from_file(MidiNoteBuffer&, const std::filesystem::path&) { return false; }

}  // namespace wb
