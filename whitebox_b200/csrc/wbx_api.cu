// wbx_api.cu — the extern "C" ABI of include/wbx.h: engine object, resident samples, schedule upload,
// kernel launches, result copies. No torch, no CPU fallback: every entry point fails loudly without a device.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/wbx.h"
#include "wbx_device.cuh"

namespace wbx {
cudaError_t launch_mix(const MixParams& p, int fpl, int n_sm, cudaStream_t stream, int* ctas_out);
cudaError_t launch_expand(const DSpan* spans, uint32_t n_spans, DCell* cells, uint32_t n_tracks, uint32_t slots,
                          uint32_t n_blocks, cudaStream_t stream);
cudaError_t launch_clamp(float* x, uint64_t n, int n_sm, cudaStream_t stream);
cudaError_t launch_levels(const float* peaks, uint32_t K, uint32_t NC, float* levels, cudaStream_t stream);
cudaError_t launch_interleave(const float* bus, uint64_t frames, uint32_t channels, int fmt, void* dst, int n_sm,
                              cudaStream_t stream);
cudaError_t launch_interleave_sample(const void* planar, size_t plane_bytes, uint64_t frames, uint32_t nch,
                                     uint32_t esize, void* dst, int n_sm, cudaStream_t stream);
int mix_warps_per_sm(int fpl);
cudaError_t launch_mip_level0(const void* base, uint32_t fmt, uint32_t nch, uint64_t count, uint64_t mdc, int high, void* out,
                              int n_sm, cudaStream_t stream);
cudaError_t launch_mip_merge(const void* child, uint64_t child_mdc, void* parent, uint64_t parent_mdc, uint32_t nch, int high,
                             int n_sm, cudaStream_t stream);
cudaError_t launch_effects(const DSpan* spans, DCell* cells, DFx* fx, uint32_t n_fx, uint32_t N, uint32_t S, uint32_t K,
                           uint32_t B, uint32_t C, uint32_t first_fx_span, float* trackbuf, const FirLaunch& fir, const float* poly,
                           uint32_t fx_flags, uint32_t* sm_arrivals, uint64_t tbs, cudaStream_t stream);
cudaError_t launch_shard_signal(const ShardPeers& peers, uint32_t rank, uint32_t world, uint32_t epoch, cudaStream_t stream);
cudaError_t launch_shard_wait(const ShardPeers& peers, uint32_t rank, uint32_t world, uint32_t epoch,
                              unsigned long long timeout_ns, uint32_t* status, cudaStream_t stream);
cudaError_t launch_shard_reduce(const float* xchg, uint32_t W, uint32_t C, uint64_t plane, uint64_t valid,
                                const ShardPeers& peers, uint64_t chan_stride, uint64_t dst_off, int n_sm,
                                cudaStream_t stream);
cudaError_t launch_ingest(const void* host_table, void* table, size_t table_bytes, void* zero, size_t zero_bytes,
                          cudaStream_t stream);
cudaError_t launch_shard_exchange(const float* xchg, uint32_t W, uint32_t C, uint64_t plane, uint64_t valid, const ShardPeers& peers,
                                  uint64_t chan_stride, uint64_t dst_off, uint32_t rank, uint32_t epoch_arrive, uint32_t epoch_done,
                                  unsigned long long timeout_ns, uint32_t* status, uint32_t* done_counter, int n_sm,
                                  cudaStream_t stream);
cudaError_t launch_levels_direct(const float* peaks, uint32_t K, uint32_t NC, float* levels_host, cudaStream_t stream);
size_t fir_tc_tiles_bytes(uint32_t L);
uint64_t fir_tc_plane_width(uint64_t H, uint64_t T);
cudaError_t launch_fir_tc_prepare(const float* ir, uint32_t L, void* tiles, float h_scale, cudaStream_t stream);
size_t fir_tc_scratch_bytes(uint64_t H, uint64_t T, uint32_t S);
int fir_tc_split_factor();
uint32_t fir_fft_partition(uint64_t T);
size_t fir_fft_ir_bytes(uint32_t L, uint32_t P);
uint32_t fir_fft_windows(uint64_t T, uint32_t L, uint32_t P);
size_t fir_fft_ring_bytes(uint32_t cap, uint32_t n_fx, uint32_t P);
size_t fir_fft_scratch_bytes(uint64_t T, uint32_t n_fx, uint32_t P);
cudaError_t launch_fir_fft_prepare(const float* ir, uint32_t L, void* ir_spectra, uint32_t P, cudaStream_t stream);
}  // namespace wbx

using namespace wbx;

namespace {

constexpr size_t kSampleHeadBytes = 256;  // zero bytes in front of frame 0 (polyphase taps reach 7 frames back)

struct SampleRec {
  void* d_alloc = nullptr;  // the allocation: kSampleHeadBytes of zeros, then the frames
  void* d_base = nullptr;   // frame 0: frame-interleaved, nch channels
  uint32_t channels = 0;    // channels of the Sample as uploaded
  uint32_t nch = 0;         // channels kept on the device: min(channels, 2)
  uint32_t rate = 0, fmt = 0, esize = 0;
  uint64_t frames = 0;
  bool live = false;
  // waveform mip-maps (WaveformVisual::create): all levels of one quality, built on first request, kept resident
  void* d_mip[2] = {nullptr, nullptr};
  std::vector<uint64_t> mip_count[2];   // elements per channel of each level
  std::vector<uint64_t> mip_offset[2];  // byte offset of each level in d_mip[q]
};

// grow-only device / pinned-host buffers: no allocation in steady state
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};
struct HostBuf {
  void* p = nullptr;
  size_t cap = 0;
};

// Sharded render state: one device block per rank = [arrival words | exchange buffer | master bus],
// visible to every peer rank (CUDA IPC across processes, peer access within one process).
constexpr size_t kShardHeaderBytes = 4096;
struct Shard {
  bool on = false, connected = false;
  uint32_t rank = 0, world = 1, max_blocks = 0, B = 0, C = 0;
  void* block = nullptr;
  size_t xchg_off = 0, bus_off = 0, bytes = 0;
  void* peer_block[kMaxPeers] = {nullptr};
  bool peer_ipc[kMaxPeers] = {false};
  uint32_t epoch = 0;
  float* host_out[2] = {nullptr, nullptr};       // device view of the host output channels (wbx_shard_set_host_output)
  float* host_out_host[2] = {nullptr, nullptr};  // ... and the host pointers they were given as
  uint64_t host_out_frames = 0;
  uint32_t* status = nullptr;  // page-locked host word raised by a barrier that timed out
  uint32_t* done_counter = nullptr;  // device word: blocks of the fused exchange kernel that have finished their reduce
  unsigned long long timeout_ns = 10ull * 1000 * 1000 * 1000;
};

}  // namespace

struct wbx_engine {
  int device = 0;
  int n_sm = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  uint32_t C = 2, B = 512, rate = 48000, n_tracks = 0;
  int sum_mode = WBX_SUM_AUTO;
  std::vector<SampleRec> samples;
  DevBuf d_smarr;  // per-SM arrival counters of fx_chain_kernel (role rotation)
  uint32_t fx_flags = 0;  // bit 0: some chain has an EQ or a compressor, bit 1: some chain is reverb-only
  DevBuf d_spans, d_cells, d_bus, d_zero, d_ws, d_conv, d_upload, d_fx, d_trackbuf, d_ir, d_firhist, d_firin, d_irtiles, d_firplanes, d_fftz, d_poly;
  HostBuf h_spans, h_bus, h_peaks, h_conv, h_levels, h_fx;
  // views into d_spans (the submitted table: spans | gains | cells of a one-callback render) and d_zero (the region one
  // memset clears before a mix: work counters | peaks | levels)
  float* gains_ptr = nullptr;
  DCell* cells_ptr = nullptr;
  uint32_t* counters_ptr = nullptr;
  float* peaks_ptr = nullptr;
  float* levels_ptr = nullptr;
  // one-callback render: table bytes the next mix's ingest kernel still has to pull from h_spans (0 = already on the device)
  size_t ingest_bytes = 0;
  // the staging table (h_spans) is read asynchronously by the H2D copy / ingest kernel: the next submit waits for that
  // read before it overwrites the table (matters for callers that submit again without fetching in between)
  cudaEvent_t staging_read = nullptr;
  bool staging_busy = false;
  std::vector<uint32_t> slot_busy;  // [track][slot] -> first free block
  uint32_t slot_cap = 0;
  // last submit
  uint32_t n_blocks = 0, n_spans = 0, slots = 1;
  bool submitted = false, mixed = false;
  uint32_t seg_flags = 0;  // OR of the submitted segments' flags
  uint64_t launches = 0;
  uint32_t upload_flip = 0;
  // effect chains (extension): host copy of the designed coefficients per track, device array with state
  std::vector<wbx_effects> fx;       // indexed by track
  std::vector<uint8_t> fx_on;        // track has a chain
  std::vector<uint8_t> fx_reset;     // clear the track's state at the next submit
  bool fx_dirty = false;
  uint32_t n_fx = 0;                 // chains resident in d_fx
  uint32_t ir_taps = 0;              // convolution reverb: taps of the impulse response in d_ir
  uint32_t firhist_tracks = 0;       // tracks d_firhist is sized (and zeroed) for
  int fir_mode = 0;                  // reverb path: 0 direct form, 1 tensor cores (d_irtiles = Toeplitz tiles), 2 partitioned FFT
                                     // (d_irtiles = twiddles + partition spectra)
  uint32_t fir_fft_p = 0;            // partition size of the FFT path the spectra in d_irtiles were built for (0: none yet)
  uint64_t firhist_pos = 0;          // the history ring's origin: logical index i lives at (firhist_pos + i) mod (taps - 1)
  uint64_t fx_gen = 0;               // bumped whenever the chain list in d_fx is rebuilt (or a chain's state reset)
  // FFT path: the window spectra ring in d_fftz persists across renders (wbx_fir_fft.cu)
  bool fftz_valid = false;
  uint32_t fftz_cap = 0, fftz_base = 0, fftz_nfx = 0, fftz_p = 0;
  uint64_t fftz_prev_frames = 0, fftz_gen = 0;
  float* mirror[2] = {nullptr, nullptr};       // device view of page-locked caller channels the running render also writes
  float* mirror_host[2] = {nullptr, nullptr};  // ... and the caller's pointers they belong to (wbx_render)
  bool levels_queued = false;              // level reduce + copy into h_levels already enqueued for this mix
  // offline bounce (wbx_bounce_*): two chunk slots, converted on the main stream, copied out on a second stream
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t bounce_ready[2] = {nullptr, nullptr}, bounce_done[2] = {nullptr, nullptr};
  DevBuf d_bounce[2];
  HostBuf h_bounce[2];
  size_t bounce_bytes[2] = {0, 0};
  int bounce_fmt = 0;
  uint64_t bounce_pushed = 0, bounce_popped = 0;
  bool bounce_on = false;
  Shard shard;
  bool shard_result = false;               // the last mix was sharded: the master bus is shard.block + bus_off (rank 0)
  int shard_phase = 0;                     // next phase of the running sharded mix (0 = none running)
  char err[256] = {0};
  char kernel_name[64] = {0};
};

namespace {

int fail(wbx_engine* e, int code, const char* fmt, ...) {
  if (e) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(e->err, sizeof(e->err), fmt, ap);
    va_end(ap);
  }
  return code;
}

#define CU(e, call)                                                                            \
  do {                                                                                         \
    cudaError_t _err = (call);                                                                 \
    if (_err != cudaSuccess)                                                                   \
      return fail((e), WBX_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_err), \
                  __FILE__, __LINE__);                                                         \
  } while (0)

int dev_reserve(wbx_engine* e, DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return WBX_OK;
  if (b.p) {
    CU(e, cudaStreamSynchronize(e->stream));
    CU(e, cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
  }
  size_t cap = bytes + bytes / 4 + 256;
  cudaError_t err = cudaMalloc(&b.p, cap);
  if (err != cudaSuccess) return fail(e, WBX_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", cap, cudaGetErrorString(err));
  b.cap = cap;
  return WBX_OK;
}

int host_reserve(wbx_engine* e, HostBuf& b, size_t bytes) {
  if (bytes <= b.cap) return WBX_OK;
  if (b.p) {
    CU(e, cudaStreamSynchronize(e->stream));
    CU(e, cudaFreeHost(b.p));
    b.p = nullptr;
    b.cap = 0;
  }
  size_t cap = bytes + bytes / 4 + 256;
  cudaError_t err = cudaMallocHost(&b.p, cap);
  if (err != cudaSuccess)
    return fail(e, WBX_ERR_NOMEM, "cudaMallocHost(%zu) failed: %s", cap, cudaGetErrorString(err));
  b.cap = cap;
  return WBX_OK;
}

// grow a page-locked buffer while keeping its first keep_bytes
int host_reserve_keep(wbx_engine* e, HostBuf& b, size_t bytes, size_t keep_bytes) {
  if (bytes <= b.cap) return WBX_OK;
  void* np = nullptr;
  const size_t cap = bytes + bytes / 4 + 256;
  if (cudaMallocHost(&np, cap) != cudaSuccess) return fail(e, WBX_ERR_NOMEM, "cudaMallocHost(%zu) failed", cap);
  if (b.p) {
    CU(e, cudaStreamSynchronize(e->stream));
    memcpy(np, b.p, keep_bytes < b.cap ? keep_bytes : b.cap);
    CU(e, cudaFreeHost(b.p));
  }
  b.p = np;
  b.cap = cap;
  return WBX_OK;
}

// num_actual_samples of one Sampler::stream call — the host twin of clipped_length in wbx_kernels.cu (same IEEE
// operations; this file is built with -ffp-contract=off): min(n, (uint32)ceil((count - offset) / speed)), sampler.cpp:102-104
uint32_t host_clipped_length(double cnt, double pos, double speed, uint32_t length) {
  const double safe = ((double)length + 1.0) * speed;
  const double rem = cnt - pos;
  if (rem >= safe) return length;
  const double m = std::ceil(rem / speed);
  const uint32_t mm = m >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)m;
  return mm < length ? mm : length;
}

uint32_t esize_of(int fmt) {
  switch (fmt) {
    case WBX_FMT_I16: return 2;
    case WBX_FMT_I24:
    case WBX_FMT_I32:
    case WBX_FMT_F32: return 4;
    default: return 0;
  }
}

// frames per lane (tile = 32 * fpl frames) and track groups for a render of n_blocks callbacks
void choose_shape(const wbx_engine* e, uint32_t n_blocks, int* fpl_out, uint32_t* groups_out) {
  const uint32_t B = e->B, N = e->n_tracks;
  const long sm = e->n_sm;
  int fpl = 16;
  if (B <= 128)
    fpl = 4;
  else if (B <= 256)
    fpl = 8;
  auto items_for = [&](int f) { return (long)n_blocks * ((B + 32 * f - 1) / (32 * f)); };
  auto warps_for = [&](int f) { return sm * wbx::mix_warps_per_sm(f); };
  const char* env = getenv("WBX_FPL");
  if (env && (atoi(env) == 4 || atoi(env) == 8 || atoi(env) == 16)) {
    fpl = atoi(env);
  } else {
    // prefer big tiles (4 KiB bulk copies): bulk-copy issue is a per-SM resource, so halving the tile costs more than
    // idle warp slots do (measured, 1024 callbacks: 512-frame tiles on 43 % of the warp slots 0.75 ms, 128-frame tiles on
    // all of them 1.8 ms). Split tiles only while the items cannot give every SM four busy warps.
    while (fpl > 4 && items_for(fpl) < sm * 4) fpl >>= 1;
  }
  uint32_t groups = 1;
  const long items = items_for(fpl), warps = warps_for(fpl);
  bool tree = e->sum_mode == WBX_SUM_TREE || (e->sum_mode == WBX_SUM_AUTO && items < warps);
  if (tree && N > 32) {
    long want = (2 * warps + items - 1) / items;  // ~2 items per resident warp
    long max_groups = (N + 15) / 16;              // at least 16 tracks per group
    if (want > max_groups) want = max_groups;
    if (want < 1) want = 1;
    groups = (uint32_t)want;
  }
  const char* genv = getenv("WBX_GROUPS");
  if (genv && atoi(genv) > 0 && e->sum_mode != WBX_SUM_EXACT) groups = (uint32_t)atoi(genv);
  if (groups > N && N > 0) groups = N;
  if (groups < 1) groups = 1;
  *fpl_out = fpl;
  *groups_out = groups;
}

}  // namespace

extern "C" {

int wbx_abi_version(void) { return WBX_ABI_VERSION; }

int wbx_create(wbx_engine** out, int device_ordinal) {
  if (!out) return WBX_ERR_INVALID;
  *out = nullptr;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0) return WBX_ERR_NO_DEVICE;
  if (device_ordinal < 0 || device_ordinal >= n_dev) return WBX_ERR_NO_DEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device_ordinal) != cudaSuccess) return WBX_ERR_NO_DEVICE;
  if (prop.major != 10) return WBX_ERR_NO_DEVICE;  // sm_100a code only
  wbx_engine* e = new (std::nothrow) wbx_engine();
  if (!e) return WBX_ERR_NOMEM;
  e->device = device_ordinal;
  e->n_sm = prop.multiProcessorCount;
  if (cudaSetDevice(device_ordinal) != cudaSuccess ||
      cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete e;
    return WBX_ERR_CUDA;
  }
  e->stream = e->own_stream;
  if (cudaEventCreateWithFlags(&e->staging_read, cudaEventDisableTiming) != cudaSuccess) {
    cudaStreamDestroy(e->own_stream);
    delete e;
    return WBX_ERR_CUDA;
  }
  // polyphase coefficient table (extension, include/wbx.h): h[ph][k] = sinc(u) * blackman(u), u = (k - 7) - ph / 128,
  // unit DC gain per phase, designed in f64 (same formulas as oracle/wb_oracle.c design_polyphase)
  {
    std::vector<float> table(128 * 16);
    const double pi = 3.141592653589793238462643383279502884;
    for (int ph = 0; ph < 128; ph++) {
      double h[16], sum = 0.0;
      for (int k = 0; k < 16; k++) {
        const double u = (double)(k - 7) - (double)ph / 128.0;
        const double sinc = u == 0.0 ? 1.0 : std::sin(pi * u) / (pi * u);
        const double w = 0.42 + 0.5 * std::cos(pi * u / 8.0) + 0.08 * std::cos(2.0 * pi * u / 8.0);
        h[k] = sinc * w;
        sum += h[k];
      }
      for (int k = 0; k < 16; k++) table[ph * 16 + k] = (float)(h[k] / sum);
    }
    if (cudaMalloc(&e->d_poly.p, table.size() * sizeof(float)) != cudaSuccess ||
        cudaMemcpy(e->d_poly.p, table.data(), table.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
      cudaStreamDestroy(e->own_stream);
      delete e;
      return WBX_ERR_CUDA;
    }
    e->d_poly.cap = table.size() * sizeof(float);
  }
  *out = e;
  return WBX_OK;
}

int wbx_destroy(wbx_engine* e) {
  if (!e) return WBX_OK;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  wbx_shard_close(e);
  for (auto& s : e->samples)
    if (s.live) {
      cudaFree(s.d_alloc);
      for (int q = 0; q < 2; q++)
        if (s.d_mip[q]) cudaFree(s.d_mip[q]);
    }
  for (DevBuf* b : {&e->d_smarr, &e->d_spans, &e->d_cells, &e->d_bus, &e->d_zero, &e->d_ws, &e->d_conv,
                    &e->d_upload, &e->d_fx, &e->d_trackbuf, &e->d_ir, &e->d_firhist, &e->d_firin, &e->d_irtiles, &e->d_firplanes, &e->d_fftz, &e->d_poly})
    if (b->p) cudaFree(b->p);
  for (HostBuf* b : {&e->h_spans, &e->h_bus, &e->h_peaks, &e->h_conv, &e->h_levels, &e->h_fx})
    if (b->p) cudaFreeHost(b->p);
  if (e->staging_read) cudaEventDestroy(e->staging_read);
  for (int i = 0; i < 2; i++) {
    if (e->bounce_ready[i]) cudaEventDestroy(e->bounce_ready[i]);
    if (e->bounce_done[i]) cudaEventDestroy(e->bounce_done[i]);
    if (e->d_bounce[i].p) cudaFree(e->d_bounce[i].p);
    if (e->h_bounce[i].p) cudaFreeHost(e->h_bounce[i].p);
  }
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  cudaStreamDestroy(e->own_stream);
  delete e;
  return WBX_OK;
}

const char* wbx_last_error(const wbx_engine* e) { return e ? e->err : "null engine"; }

int wbx_configure(wbx_engine* e, uint32_t out_channels, uint32_t block_frames, uint32_t sample_rate) {
  if (!e) return WBX_ERR_INVALID;
  if (out_channels < 1 || out_channels > 2)
    return fail(e, WBX_ERR_UNSUPPORTED, "out_channels must be 1 or 2 (pan_coeffs[2], engine/track.h:50)");
  if (block_frames == 0 || block_frames > 65535) return fail(e, WBX_ERR_INVALID, "block_frames out of range");
  e->C = out_channels;
  e->B = block_frames;
  e->rate = sample_rate;
  e->submitted = e->mixed = false;
  return WBX_OK;
}

int wbx_set_track_count(wbx_engine* e, uint32_t n_tracks) {
  if (!e) return WBX_ERR_INVALID;
  if (n_tracks != e->n_tracks && !e->fx.empty()) {  // chains are attached to track indices
    e->fx.resize(n_tracks);
    e->fx_on.resize(n_tracks, 0);
    e->fx_reset.resize(n_tracks, 1);
    e->fx_dirty = true;
  }
  e->n_tracks = n_tracks;
  e->submitted = e->mixed = false;
  return WBX_OK;
}

int wbx_set_sum_mode(wbx_engine* e, int mode) {
  if (!e || mode < WBX_SUM_AUTO || mode > WBX_SUM_TREE) return WBX_ERR_INVALID;
  e->sum_mode = mode;
  return WBX_OK;
}

int wbx_set_stream(wbx_engine* e, void* cuda_stream) {
  if (!e) return WBX_ERR_INVALID;
  CU(e, cudaSetDevice(e->device));
  CU(e, cudaStreamSynchronize(e->stream));
  e->stream = cuda_stream ? (cudaStream_t)cuda_stream : e->own_stream;
  return WBX_OK;
}

int wbx_sample_upload(wbx_engine* e, int format, uint32_t channels, uint64_t frames, uint32_t sample_rate,
                      const void* const* planar, uint32_t* out_id) {
  if (!e || !planar || !out_id || channels == 0) return fail(e, WBX_ERR_INVALID, "bad sample arguments");
  const uint32_t es = esize_of(format);
  if (!es) return fail(e, WBX_ERR_UNSUPPORTED, "sample format %d not supported by the sampler", format);
  CU(e, cudaSetDevice(e->device));
  SampleRec r;
  r.fmt = (uint32_t)format;
  r.esize = es;
  r.channels = channels;
  r.rate = sample_rate;
  r.frames = frames;
  r.nch = channels < 2 ? 1 : 2;  // output channel c reads source channel c % channels, c < 2 (sampler.cpp:111)
  // frames + 16 zero frames (dsp/sample.cpp:127,140) + slack for 16-B aligned windows and the 2-tap reach
  const size_t alloc_frames = ((size_t)frames + 16 + 32 + 31) & ~(size_t)31;
  const size_t bytes = alloc_frames * es * r.nch;
  const size_t plane_bytes = (((size_t)frames * es) + 255) & ~(size_t)255;
  int rc = dev_reserve(e, e->d_upload, plane_bytes * r.nch);
  if (rc) return rc;
  cudaError_t err = cudaMalloc(&r.d_alloc, bytes + kSampleHeadBytes);
  if (err != cudaSuccess) return fail(e, WBX_ERR_NOMEM, "cudaMalloc(%zu) for sample failed", bytes);
  r.d_base = (uint8_t*)r.d_alloc + kSampleHeadBytes;
  CU(e, cudaMemsetAsync(r.d_alloc, 0, bytes + kSampleHeadBytes, e->stream));
  for (uint32_t c = 0; c < r.nch; c++)
    CU(e, cudaMemcpyAsync((uint8_t*)e->d_upload.p + c * plane_bytes, planar[c], (size_t)frames * es,
                          cudaMemcpyHostToDevice, e->stream));
  CU(e, launch_interleave_sample(e->d_upload.p, plane_bytes, frames, r.nch, es, r.d_base, e->n_sm, e->stream));
  e->launches++;
  CU(e, cudaStreamSynchronize(e->stream));  // the caller's host arrays may go away after return
  r.live = true;
  uint32_t id = 0;
  for (; id < e->samples.size(); id++)
    if (!e->samples[id].live) break;
  if (id == e->samples.size())
    e->samples.push_back(r);
  else
    e->samples[id] = r;
  *out_id = id;
  return WBX_OK;
}

// true when p is page-locked host memory the device can copy to/from directly (cudaMallocHost / cudaHostRegister)
static bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

int wbx_sample_update(wbx_engine* e, uint32_t id, const void* const* planar) {
  if (!e || !planar || id >= e->samples.size() || !e->samples[id].live) return fail(e, WBX_ERR_INVALID, "bad sample id");
  CU(e, cudaSetDevice(e->device));
  SampleRec& r = e->samples[id];
  for (int q = 0; q < 2; q++) {  // the data changes: cached mip-maps are stale
    if (r.d_mip[q]) {
      CU(e, cudaStreamSynchronize(e->stream));
      CU(e, cudaFree(r.d_mip[q]));
      r.d_mip[q] = nullptr;
    }
    r.mip_count[q].clear();
    r.mip_offset[q].clear();
  }
  const size_t plane_bytes = (((size_t)r.frames * r.esize) + 255) & ~(size_t)255;
  // two staging halves alternate so the next sample's H2D overlaps this one's interleave kernel
  int rc = dev_reserve(e, e->d_upload, 2 * plane_bytes * r.nch);
  if (rc) return rc;
  uint8_t* stage = (uint8_t*)e->d_upload.p + (size_t)(e->upload_flip & 1u) * plane_bytes * r.nch;
  e->upload_flip++;
  bool pinned = true;
  for (uint32_t c = 0; c < r.nch; c++) {
    pinned = pinned && is_pinned(planar[c]);
    CU(e, cudaMemcpyAsync(stage + c * plane_bytes, planar[c], (size_t)r.frames * r.esize, cudaMemcpyHostToDevice,
                          e->stream));
  }
  CU(e, launch_interleave_sample(stage, plane_bytes, r.frames, r.nch, r.esize, r.d_base, e->n_sm, e->stream));
  e->launches++;
  if (!pinned) CU(e, cudaStreamSynchronize(e->stream));
  return WBX_OK;
}

int wbx_sample_mipmap(wbx_engine* e, uint32_t id, int quality, int level, void* out, uint64_t cap_elems,
                      uint32_t* count) {
  if (!e || id >= e->samples.size() || !e->samples[id].live) return fail(e, WBX_ERR_INVALID, "bad sample id");
  SampleRec& r = e->samples[id];
  if (r.channels > 2) return fail(e, WBX_ERR_UNSUPPORTED, "mip-maps: only the 2 resident channels are available");
  if (r.fmt == WBX_FMT_I24) return 0;  // the reference's switch has no case for this tag (`default: break`)
  CU(e, cudaSetDevice(e->device));
  const int q = quality ? 1 : 0;
  const size_t esz = q ? 2 : 1;
  if (r.mip_count[q].empty() && r.frames > 64) {
    // WaveformVisual::create (gfx/waveform_visual.cpp:181-248): levels with chunk 2, 8, 32, ... while count/4^l > 64
    uint64_t sample_count = r.frames, total = 0;
    uint32_t current_mip = 1;
    while (sample_count > 64) {
      const uint64_t block = 1ull << (current_mip - 1);
      uint64_t mdc = r.frames / block;
      mdc += mdc % 2;
      r.mip_count[q].push_back(mdc);
      r.mip_offset[q].push_back(total);
      total += ((mdc * r.nch * esz) + 255) & ~(uint64_t)255;
      sample_count /= 4;
      current_mip += 2;
    }
    cudaError_t err = cudaMalloc(&r.d_mip[q], total);
    if (err != cudaSuccess) {
      r.mip_count[q].clear();
      r.mip_offset[q].clear();
      return fail(e, WBX_ERR_NOMEM, "cudaMalloc(%llu) for mip-maps failed", (unsigned long long)total);
    }
    uint8_t* base = (uint8_t*)r.d_mip[q];
    CU(e, launch_mip_level0(r.d_base, r.fmt, r.nch, r.frames, r.mip_count[q][0], q, base, e->n_sm, e->stream));
    e->launches++;
    for (size_t l = 1; l < r.mip_count[q].size(); l++) {  // each level from the previous one (4 chunks -> 1)
      CU(e, launch_mip_merge(base + r.mip_offset[q][l - 1], r.mip_count[q][l - 1], base + r.mip_offset[q][l],
                             r.mip_count[q][l], r.nch, q, e->n_sm, e->stream));
      e->launches++;
    }
  }
  const int n_levels = (int)r.mip_count[q].size();
  if (level >= 0 && level < n_levels) {
    const uint64_t mdc = r.mip_count[q][level];
    if (count) *count = (uint32_t)mdc;
    const uint64_t elems = mdc * r.nch;
    if (out && elems <= cap_elems) {
      int rc;
      if ((rc = host_reserve(e, e->h_conv, elems * esz))) return rc;
      CU(e, cudaMemcpyAsync(e->h_conv.p, (uint8_t*)r.d_mip[q] + r.mip_offset[q][level], elems * esz, cudaMemcpyDeviceToHost,
                            e->stream));
      CU(e, cudaStreamSynchronize(e->stream));
      memcpy(out, e->h_conv.p, elems * esz);
    }
  }
  return n_levels;
}

int wbx_sample_release(wbx_engine* e, uint32_t id) {
  if (!e || id >= e->samples.size() || !e->samples[id].live) return fail(e, WBX_ERR_INVALID, "bad sample id");
  CU(e, cudaSetDevice(e->device));
  CU(e, cudaStreamSynchronize(e->stream));
  CU(e, cudaFree(e->samples[id].d_alloc));
  for (int q = 0; q < 2; q++)
    if (e->samples[id].d_mip[q]) CU(e, cudaFree(e->samples[id].d_mip[q]));
  e->samples[id] = SampleRec();
  e->submitted = e->mixed = false;  // the submitted span table may hold this sample's device pointer
  return WBX_OK;
}

// RBJ "Audio EQ Cookbook" biquad, designed in f64, normalised by a0, stored f32 (same formulas as the C port).
static void design_band(int band, double freq, double gain_db, double q, double rate, float* b0, float* b1, float* b2,
                        float* a1, float* a2) {
  const double pi = 3.141592653589793238462643383279502884;
  const double A = std::pow(10.0, gain_db / 40.0);
  const double w0 = 2.0 * pi * freq / rate;
  const double cw = std::cos(w0), sw = std::sin(w0);
  const double alpha = sw / (2.0 * q);
  double B0, B1, B2, A0, A1, A2;
  if (band == 0) {  // low shelf
    const double sq = 2.0 * std::sqrt(A) * alpha;
    B0 = A * ((A + 1.0) - (A - 1.0) * cw + sq);
    B1 = 2.0 * A * ((A - 1.0) - (A + 1.0) * cw);
    B2 = A * ((A + 1.0) - (A - 1.0) * cw - sq);
    A0 = (A + 1.0) + (A - 1.0) * cw + sq;
    A1 = -2.0 * ((A - 1.0) + (A + 1.0) * cw);
    A2 = (A + 1.0) + (A - 1.0) * cw - sq;
  } else if (band == 3) {  // high shelf
    const double sq = 2.0 * std::sqrt(A) * alpha;
    B0 = A * ((A + 1.0) + (A - 1.0) * cw + sq);
    B1 = -2.0 * A * ((A - 1.0) + (A + 1.0) * cw);
    B2 = A * ((A + 1.0) + (A - 1.0) * cw - sq);
    A0 = (A + 1.0) - (A - 1.0) * cw + sq;
    A1 = 2.0 * ((A - 1.0) - (A + 1.0) * cw);
    A2 = (A + 1.0) - (A - 1.0) * cw - sq;
  } else {  // peaking
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

int wbx_effects_design(const wbx_effect_params* p, uint32_t sample_rate, wbx_effects* out) {
  if (!p || !out || sample_rate == 0) return WBX_ERR_INVALID;
  memset(out, 0, sizeof(*out));
  for (int b = 0; b < 4; b++) {
    if (p->eq_gain_db[b] != 0.0f) out->eq_on = 1;
    if (!(p->eq_freq[b] > 0.0f) || !(p->eq_q[b] > 0.0f)) return WBX_ERR_INVALID;
    design_band(b, p->eq_freq[b], p->eq_gain_db[b], p->eq_q[b], (double)sample_rate, &out->b0[b], &out->b1[b],
                &out->b2[b], &out->a1[b], &out->a2[b]);
  }
  if (p->comp_ratio_code < 0 || p->comp_ratio_code > 4) return WBX_ERR_INVALID;
  out->comp_ratio_code = (uint32_t)p->comp_ratio_code;
  out->comp_on = p->comp_ratio_code != 0;
  out->comp_threshold = (float)std::pow(10.0, (double)p->comp_threshold_db / 20.0);
  out->comp_makeup = (float)std::pow(10.0, (double)p->comp_makeup_db / 20.0);
  out->comp_attack = (float)std::exp(-1.0 / ((double)p->comp_attack_ms * 0.001 * (double)sample_rate));
  out->comp_release = (float)std::exp(-1.0 / ((double)p->comp_release_ms * 0.001 * (double)sample_rate));
  out->reverb_on = p->reverb_on != 0;
  return WBX_OK;
}

int wbx_set_impulse_response(wbx_engine* e, const float* h, uint32_t n_taps) {
  if (!e) return WBX_ERR_INVALID;
  CU(e, cudaSetDevice(e->device));
  CU(e, cudaStreamSynchronize(e->stream));
  e->ir_taps = 0;
  e->firhist_tracks = 0;  // histories are re-created (zeroed) at the next submit
  e->firhist_pos = 0;
  e->fftz_valid = false;
  e->fir_fft_p = 0;
  if (!h || n_taps == 0) return WBX_OK;
  int rc = dev_reserve(e, e->d_ir, (size_t)n_taps * sizeof(float));
  if (rc) return rc;
  CU(e, cudaMemcpyAsync(e->d_ir.p, h, (size_t)n_taps * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  // long responses run as a partitioned FFT convolution, short ones in direct form; WBX_FIR=direct / tc / fft picks the path
  // by hand (tc = the direct form as a Toeplitz GEMM on the tensor cores)
  const char* mode = getenv("WBX_FIR");
  e->fir_mode = mode ? (mode[0] == 't' ? 1 : (mode[0] == 'f' ? 2 : 0)) : (n_taps >= 1024 ? 2 : 0);
  if (e->fir_mode == 1) {
    if ((rc = dev_reserve(e, e->d_irtiles, fir_tc_tiles_bytes(n_taps)))) return rc;
    // fp16 operands: scale the response so that its largest tap lands in [2^13, 2^14) (undone exactly in the epilogue)
    float hmax = 0.0f;
    for (uint32_t i = 0; i < n_taps; i++) hmax = std::fmax(hmax, std::fabs(h[i]));
    float h_scale = 1.0f;
    if (hmax > 0.0f && std::isfinite(hmax)) {
      int ex;
      std::frexp(hmax, &ex);
      h_scale = std::ldexp(1.0f, 14 - ex);
    }
    CU(e, launch_fir_tc_prepare((const float*)e->d_ir.p, n_taps, e->d_irtiles.p, h_scale, e->stream));
    e->launches++;
  } else if (e->fir_mode == 2) {
    // the partition spectra depend on the partition size, which is chosen per render: built at the first submit
  }
  CU(e, cudaStreamSynchronize(e->stream));
  e->ir_taps = n_taps;
  return WBX_OK;
}

int wbx_set_track_effects(wbx_engine* e, uint32_t track, const wbx_effects* fx) {
  if (!e) return WBX_ERR_INVALID;
  if (track >= e->n_tracks) return fail(e, WBX_ERR_INVALID, "effects: track %u >= %u", track, e->n_tracks);
  if (e->fx.size() < e->n_tracks) {
    e->fx.resize(e->n_tracks);
    e->fx_on.resize(e->n_tracks, 0);
    e->fx_reset.resize(e->n_tracks, 0);
  }
  const bool on = fx && (fx->eq_on || fx->comp_on || fx->reverb_on);
  if (on) e->fx[track] = *fx;
  e->fx_on[track] = on ? 1 : 0;
  e->fx_reset[track] = 1;
  e->fx_dirty = true;
  return WBX_OK;
}

// Tables of the chain's time-parallel form — the same f64 design as oracle/wb_oracle.c fx_design_tables (this file is
// built with -ffp-contract=off: every product and sum below is separately rounded, as there).
static void mat2_mul(const double a[4], const double b[4], double out[4]) {
  const double o0 = a[0] * b[0] + a[1] * b[2], o1 = a[0] * b[1] + a[1] * b[3];
  const double o2 = a[2] * b[0] + a[3] * b[2], o3 = a[2] * b[1] + a[3] * b[3];
  out[0] = o0, out[1] = o1, out[2] = o2, out[3] = o3;
}
static void design_fx_tables(DFx& d) {
  for (int b = 0; b < 4; b++) {
    const double a1 = (double)d.a1[b], a2 = (double)d.a2[b], b0 = (double)d.b0[b];
    d.B1[b] = (float)((double)d.b1[b] - a1 * b0);
    d.B2[b] = (float)((double)d.b2[b] - a2 * b0);
    const double A[4] = {-a1, 1.0, -a2, 0.0};
    double pw[4] = {1.0, 0.0, 0.0, 1.0};
    for (int m = 0; m <= 16; m++) {
      for (int q = 0; q < 4; q++) d.P[b][m][q] = (float)pw[q];
      if (m < 16) mat2_mul(pw, A, pw);
    }
    double sq[4] = {pw[0], pw[1], pw[2], pw[3]};  // A^16
    for (int j = 0; j < 5; j++) {
      for (int q = 0; q < 4; q++) d.S[b][j][q] = (float)sq[q];
      mat2_mul(sq, sq, sq);
    }
  }
  const float a = d.att, r = d.rel;
  d.sel = a <= r ? 1u : 0u;
  d.a1m = 1.0f - a;
  d.r1m = 1.0f - r;
  const float r2 = r * r, a2 = a * a, r3 = r2 * r, a3 = a2 * a;
  float* sl = d.sl;
  sl[0] = r, sl[1] = a;
  sl[2] = r2, sl[3] = a * r, sl[4] = a2;
  sl[5] = r3, sl[6] = a * r2, sl[7] = a2 * r, sl[8] = a3;
  sl[9] = r2 * r2, sl[10] = a * r3, sl[11] = a2 * r2, sl[12] = a3 * r, sl[13] = a2 * a2;
}

// (re)build the compact device array of chains, keeping the running state of tracks whose chain did not change
static int sync_effects(wbx_engine* e) {
  if (e->ir_taps > 1 && e->firhist_tracks != e->n_tracks && !e->fx_on.empty()) e->fx_dirty = true;
  if (!e->fx_dirty) return WBX_OK;
  std::vector<DFx> old;
  if (e->n_fx) {
    old.resize(e->n_fx);
    CU(e, cudaMemcpyAsync(old.data(), e->d_fx.p, e->n_fx * sizeof(DFx), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
  }
  std::vector<DFx> cur;
  for (uint32_t t = 0; t < e->n_tracks && t < e->fx_on.size(); t++) {
    if (!e->fx_on[t]) continue;
    DFx d;
    memset(&d, 0, sizeof(d));
    const wbx_effects& f = e->fx[t];
    d.track = t;
    d.eq_on = f.eq_on;
    d.comp_on = f.comp_on;
    d.ratio_code = f.comp_ratio_code;
    for (int b = 0; b < 4; b++) {
      d.b0[b] = f.b0[b];
      d.b1[b] = f.b1[b];
      d.b2[b] = f.b2[b];
      d.a1[b] = f.a1[b];
      d.a2[b] = f.a2[b];
    }
    d.reverb_on = f.reverb_on;
    d.thr = f.comp_threshold;
    d.att = f.comp_attack;
    d.rel = f.comp_release;
    d.makeup = f.comp_makeup;
    design_fx_tables(d);
    if (!e->fx_reset[t])
      for (const DFx& o : old)
        if (o.track == t) {
          memcpy(d.s1, o.s1, sizeof(d.s1));
          memcpy(d.s2, o.s2, sizeof(d.s2));
          memcpy(d.env, o.env, sizeof(d.env));
        }
    cur.push_back(d);
  }
  // reverb histories: [n_tracks][2][taps-1]; (re)attached chains start from silence
  if (e->ir_taps > 1) {
    const size_t H = e->ir_taps - 1;
    if (e->firhist_tracks != e->n_tracks) {
      int rc = dev_reserve(e, e->d_firhist, (size_t)e->n_tracks * 2 * H * sizeof(float));
      if (rc) return rc;
      CU(e, cudaMemsetAsync(e->d_firhist.p, 0, (size_t)e->n_tracks * 2 * H * sizeof(float), e->stream));
      e->firhist_tracks = e->n_tracks;
    } else {
      for (uint32_t t = 0; t < e->n_tracks && t < e->fx_reset.size(); t++)
        if (e->fx_reset[t])
          CU(e, cudaMemsetAsync((float*)e->d_firhist.p + (size_t)t * 2 * H, 0, 2 * H * sizeof(float), e->stream));
    }
  }
  std::fill(e->fx_reset.begin(), e->fx_reset.end(), 0);
  e->n_fx = (uint32_t)cur.size();
  e->fx_flags = 0;
  for (const DFx& d : cur) e->fx_flags |= (d.eq_on || d.comp_on) ? 1u : 2u;
  {
    int rc = dev_reserve(e, e->d_smarr, 1024 * sizeof(uint32_t));
    if (rc) return rc;
  }
  if (e->n_fx) {
    int rc = dev_reserve(e, e->d_fx, cur.size() * sizeof(DFx));
    if (rc) return rc;
    if ((rc = host_reserve(e, e->h_fx, cur.size() * sizeof(DFx)))) return rc;
    memcpy(e->h_fx.p, cur.data(), cur.size() * sizeof(DFx));
    CU(e, cudaMemcpyAsync(e->d_fx.p, e->h_fx.p, cur.size() * sizeof(DFx), cudaMemcpyHostToDevice, e->stream));
  }
  e->fx_dirty = false;
  e->fx_gen++;
  return WBX_OK;
}

int wbx_submit(wbx_engine* e, const wbx_segment* segs, uint32_t n_segs, const float* track_gains,
               uint32_t n_blocks) {
  if (!e) return WBX_ERR_INVALID;
  if (n_blocks == 0) return fail(e, WBX_ERR_INVALID, "n_blocks must be >= 1");
  if (n_segs && !segs) return fail(e, WBX_ERR_INVALID, "segs is null");
  if (e->n_tracks && !track_gains) return fail(e, WBX_ERR_INVALID, "track_gains is null");
  CU(e, cudaSetDevice(e->device));
  const uint32_t N = e->n_tracks, B = e->B, C = e->C;
  int rc;

  // ---- validate + resolve segments into spans (pinned staging), assign cell slots --------------------
  if (e->staging_busy) {  // the previous table may still be being read by its copy / ingest kernel
    CU(e, cudaEventSynchronize(e->staging_read));
    e->staging_busy = false;
  }
  if ((rc = sync_effects(e))) return rc;
  const uint32_t n_fx = e->n_fx;
  // one page-locked table, one H2D copy: spans | track gains | (one-callback render: the cells, written here on the
  // host — a callback is one Sampler::stream call per segment, nothing to replay — so no cell memset / expand kernel)
  const size_t span_bytes = (size_t)(n_segs + n_fx + 1) * sizeof(DSpan);
  const size_t gain_bytes = (((size_t)N * 2 * sizeof(float)) + 15) & ~(size_t)15;
  const bool host_cells = n_blocks == 1;
  if (e->slot_cap == 0) e->slot_cap = 2;
  if ((rc = host_reserve(e, e->h_spans, span_bytes + gain_bytes + (host_cells ? (size_t)N * e->slot_cap * sizeof(DCell) : 0))))
    return rc;
  DSpan* hs = (DSpan*)e->h_spans.p;
  uint32_t slots = 1;
  uint32_t seg_flags = 0;
  if (e->slot_busy.size() < (size_t)N * e->slot_cap) e->slot_busy.resize((size_t)N * e->slot_cap);
  std::fill(e->slot_busy.begin(), e->slot_busy.end(), 0u);
  for (uint32_t i = 0; i < n_segs; i++) {
    const wbx_segment& sg = segs[i];
    if (sg.track >= N) return fail(e, WBX_ERR_INVALID, "segment %u: track %u >= %u", i, sg.track, N);
    if (sg.n_blocks == 0 || sg.block >= n_blocks || sg.n_blocks > n_blocks - sg.block)
      return fail(e, WBX_ERR_INVALID, "segment %u: blocks [%u,+%u) outside render of %u", i, sg.block, sg.n_blocks,
                  n_blocks);
    if (sg.dst_offset > B || sg.length > B - sg.dst_offset)
      return fail(e, WBX_ERR_INVALID, "segment %u: frames [%u,+%u) outside block of %u", i, sg.dst_offset, sg.length, B);
    if (sg.sample_id >= e->samples.size() || !e->samples[sg.sample_id].live)
      return fail(e, WBX_ERR_INVALID, "segment %u: unknown sample %u", i, sg.sample_id);
    if (!(sg.speed > 0.0) || !(sg.src_pos >= 0.0) || !(sg.speed < 1e6))
      return fail(e, WBX_ERR_INVALID, "segment %u: speed/src_pos must be positive and finite", i);
    if (sg.flags & ~(WBX_SEG_FADE | WBX_SEG_POLYPHASE))
      return fail(e, WBX_ERR_INVALID, "segment %u: unknown flags 0x%x", i, sg.flags);
    if ((sg.flags & WBX_SEG_FADE) && !(sg.clip_frame >= 0.0 && sg.clip_frame < 9.0e15 && sg.clip_len_frames >= 0.0))
      return fail(e, WBX_ERR_INVALID, "segment %u: bad fade parameters", i);
    const SampleRec& sm = e->samples[sg.sample_id];
    // slot = first one free at sg.block for this track (segments of a track arrive in block order)
    uint32_t slot = 0;
    for (;; slot++) {
      if (slot == e->slot_cap) {  // grow the per-track slot table (rare: > 2 calls in one block)
        const uint32_t nc = e->slot_cap * 2;
        std::vector<uint32_t> nb((size_t)N * nc, 0u);
        for (uint32_t t = 0; t < N; t++)
          for (uint32_t s = 0; s < e->slot_cap; s++) nb[(size_t)t * nc + s] = e->slot_busy[(size_t)t * e->slot_cap + s];
        e->slot_busy.swap(nb);
        e->slot_cap = nc;
      }
      if (e->slot_busy[(size_t)sg.track * e->slot_cap + slot] <= sg.block) break;
    }
    e->slot_busy[(size_t)sg.track * e->slot_cap + slot] = sg.block + sg.n_blocks;
    if (slot + 1 > slots) slots = slot + 1;
    DSpan& d = hs[i];
    d.base = sm.d_base;
    d.pos0 = sg.src_pos;
    d.speed = sg.speed;
    d.count = sm.frames;
    d.gain = sg.gain;
    d.track = sg.track;
    d.block0 = sg.block;
    d.n_blocks = sg.n_blocks;
    d.dst_off = sg.dst_offset;
    d.length = sg.length;
    d.fmt = sm.fmt;
    d.slot = slot;
    d.nch = sm.nch;
    seg_flags |= sg.flags;
    const bool fade = (sg.flags & WBX_SEG_FADE) != 0;
    d.fade = sg.flags & (WBX_SEG_FADE | WBX_SEG_POLYPHASE);
    d.clip_frame = fade ? sg.clip_frame : 0.0;
    d.fade_in = fade ? sg.fade_in_frames : 0.0;
    d.fade_out = fade ? sg.fade_out_frames : 0.0;
    d.clip_len = fade ? sg.clip_len_frames : 0.0;
  }
  // ---- device buffers ------------------------------------------------------------------------------
  const size_t n_cells = (size_t)n_blocks * N * slots;
  const size_t cell_bytes = host_cells ? n_cells * sizeof(DCell) : 0;
  if ((rc = host_reserve_keep(e, e->h_spans, span_bytes + gain_bytes + cell_bytes, span_bytes))) return rc;
  hs = (DSpan*)e->h_spans.p;
  if (N) memcpy((uint8_t*)hs + span_bytes, track_gains, (size_t)N * 2 * sizeof(float));
  if (host_cells && n_cells) {
    DCell* hc = (DCell*)((uint8_t*)hs + span_bytes + gain_bytes);
    memset(hc, 0xFF, cell_bytes);  // span = kSilent
    for (uint32_t i = 0; i < n_segs; i++) {
      const DSpan& d = hs[i];
      if (d.pos0 >= (double)d.count) continue;  // has finished streaming (sampler.cpp:99-100)
      DCell& c = hc[(size_t)d.track * slots + d.slot];
      c.pos = d.pos0;
      c.span = i;
      c.n_act = host_clipped_length((double)d.count, d.pos0, d.speed, d.length);
    }
  }
  const size_t bus_floats = (size_t)C * n_blocks * B;
  if ((rc = dev_reserve(e, e->d_spans, span_bytes + gain_bytes + cell_bytes))) return rc;
  if (n_fx) {
    // tracks with an effect chain are rendered to a per-track buffer (frame-interleaved stereo f32) that the mix
    // kernel then reads as one whole-block unity-speed call per callback with clip gain 1.0
    // (per-track stride rounded up to an even number of frames: every track's buffer starts 16-byte aligned)
    const size_t tbs = ((size_t)n_blocks * B + 1) & ~(size_t)1;
    const size_t tb_floats = (size_t)n_fx * tbs * 2;
    if ((rc = dev_reserve(e, e->d_trackbuf, tb_floats * sizeof(float) + 256))) return rc;
    const DFx* hf = (const DFx*)e->h_fx.p;
    for (uint32_t i = 0; i < n_fx; i++) {
      DSpan& d = hs[n_segs + i];
      memset(&d, 0, sizeof(d));
      d.base = (const float*)e->d_trackbuf.p + (size_t)i * tbs * 2;
      d.pos0 = 0.0;
      d.speed = 1.0;
      d.count = (uint64_t)n_blocks * B;
      d.gain = 1.0f;
      d.track = hf[i].track;
      d.block0 = 0;
      d.n_blocks = n_blocks;
      d.dst_off = 0;
      d.length = B;
      d.fmt = WBX_FMT_F32;
      d.slot = 0;
      d.nch = 2;
    }
  }
  if (!host_cells && (rc = dev_reserve(e, e->d_cells, (n_cells ? n_cells : 1) * sizeof(DCell)))) return rc;
  if ((rc = dev_reserve(e, e->d_bus, bus_floats * sizeof(float)))) return rc;
  e->gains_ptr = (float*)((uint8_t*)e->d_spans.p + span_bytes);
  e->cells_ptr = host_cells ? (DCell*)((uint8_t*)e->d_spans.p + span_bytes + gain_bytes) : (DCell*)e->d_cells.p;

  // a one-callback render without effect chains leaves the table to the mix's ingest kernel (kernels only, no copy engine)
  e->ingest_bytes = 0;
  if (host_cells && !n_fx && !getenv("WBX_NO_INGEST"))
    e->ingest_bytes = (span_bytes + gain_bytes + cell_bytes + 15) & ~(size_t)15;
  else {
    CU(e, cudaMemcpyAsync(e->d_spans.p, hs, span_bytes + gain_bytes + cell_bytes, cudaMemcpyHostToDevice, e->stream));
    CU(e, cudaEventRecord(e->staging_read, e->stream));
    e->staging_busy = true;
  }
  if (!host_cells) {
    if (n_cells) CU(e, cudaMemsetAsync(e->d_cells.p, 0xFF, n_cells * sizeof(DCell), e->stream));  // span = kSilent
    if (n_segs && N) {
      CU(e, launch_expand((const DSpan*)e->d_spans.p, n_segs, e->cells_ptr, N, slots, n_blocks, e->stream));
      e->launches += (n_blocks >= 64 && n_segs >= 32) ? 2 : 1;
    }
  }
  if (n_fx && N) {
    const bool reverb = e->ir_taps > 0 && e->firhist_tracks == N;
    if (e->ir_taps > 0) {
      const size_t H = e->ir_taps - 1;
      if ((rc = dev_reserve(e, e->d_firin, (size_t)n_fx * C * (H + (size_t)n_blocks * B) * sizeof(float) + 256))) return rc;
      if (H == 0 && (rc = dev_reserve(e, e->d_firhist, 256))) return rc;
    }
    const bool rv = e->ir_taps > 0 && (reverb || e->ir_taps == 1);
    FirLaunch fir;
    if (rv) {
      const uint64_t T = (uint64_t)n_blocks * B, H = e->ir_taps - 1;
      fir.ir = (const float*)e->d_ir.p;
      fir.L = e->ir_taps;
      fir.hist = (float*)e->d_firhist.p;
      fir.hist_pos = e->firhist_pos;
      fir.xin = (float*)e->d_firin.p;
      fir.mode = e->fir_mode;
      if (fir.mode == 1) {
        if ((rc = dev_reserve(e, e->d_firplanes, fir_tc_scratch_bytes(H, T, n_fx * C)))) return rc;
        fir.scratch = e->d_firplanes.p;
      } else if (fir.mode == 2) {
        const uint32_t P = fir_fft_partition(T);
        if (P != e->fir_fft_p) {  // partition spectra of the response for this partition size
          if ((rc = dev_reserve(e, e->d_irtiles, fir_fft_ir_bytes(e->ir_taps, P)))) return rc;
          CU(e, launch_fir_fft_prepare(fir.ir, e->ir_taps, e->d_irtiles.p, P, e->stream));
          e->launches++;
          e->fir_fft_p = P;
          e->fftz_valid = false;
        }
        if ((rc = dev_reserve(e, e->d_firplanes, fir_fft_scratch_bytes(T, n_fx, P)))) return rc;
        const uint32_t NQ = fir_fft_windows(T, e->ir_taps, P), NP = (e->ir_taps + P - 1) / P;
        // The previous render's windows are this render's history windows when it was a whole number of partitions long
        // and nothing else changed; otherwise every window is rebuilt from the time-domain history.
        bool warm = e->fftz_valid && e->fftz_p == P && e->fftz_nfx == n_fx && e->fftz_gen == e->fx_gen &&
                    e->fftz_prev_frames % P == 0 && e->fftz_cap >= NQ && NP > 1 && !getenv("WBX_FFT_COLD");
        if (!warm) {
          const size_t need = fir_fft_ring_bytes(NQ, n_fx, P);
          if (need > e->d_fftz.cap || e->fftz_nfx != n_fx || e->fftz_p != P || e->fftz_cap < NQ) {
            if ((rc = dev_reserve(e, e->d_fftz, need))) return rc;
            e->fftz_cap = NQ;
          }
          e->fftz_base = 0;
          fir.fft_first_q = 0;
        } else {
          e->fftz_base = (uint32_t)((e->fftz_base + e->fftz_prev_frames / P) % e->fftz_cap);
          fir.fft_first_q = NP - 1;
        }
        fir.scratch = e->d_firplanes.p;
        fir.fft_ring = e->d_fftz.p;
        fir.fft_ring_base = e->fftz_base;
        fir.fft_ring_cap = e->fftz_cap;
        fir.fft_p = P;
        e->fftz_valid = true;
        e->fftz_p = P;
        e->fftz_nfx = n_fx;
        e->fftz_gen = e->fx_gen;
        e->fftz_prev_frames = T;
      }
      if (fir.mode != 2) e->fftz_valid = false;
      fir.ir_aux = fir.mode ? e->d_irtiles.p : nullptr;
      if (H) e->firhist_pos = (e->firhist_pos + T) % H;
    }
    CU(e, launch_effects((const DSpan*)e->d_spans.p, e->cells_ptr, (DFx*)e->d_fx.p, n_fx, N, slots, n_blocks, B, C,
                         n_segs, (float*)e->d_trackbuf.p, fir, (const float*)e->d_poly.p, e->fx_flags, (uint32_t*)e->d_smarr.p,
                         (((uint64_t)n_blocks * B + 1) & ~(uint64_t)1), e->stream));
    e->launches += (rv ? (fir.mode == 1 ? 5 : 4) : 1) + ((e->fx_flags & 1u) ? 1 : 0) + ((e->fx_flags & 2u) ? 1 : 0);
  }
  e->n_blocks = n_blocks;
  e->n_spans = n_segs;
  e->slots = slots;
  e->seg_flags = seg_flags;
  e->submitted = true;
  e->mixed = false;
  return WBX_OK;
}

static ShardPeers shard_peers(const Shard& sh) {
  ShardPeers peers;
  for (uint32_t j = 0; j < kMaxPeers; j++) {
    peers.flags[j] = j < sh.world ? (uint32_t*)sh.peer_block[j] : nullptr;
    peers.dst[j] = nullptr;
  }
  peers.dst[0] = (float*)((uint8_t*)sh.peer_block[0] + sh.bus_off);  // rank 0 holds the master bus
  peers.n_dst = 1;
  peers.host_dst[0] = sh.host_out[0];
  peers.host_dst[1] = sh.host_out[1];
  return peers;
}

static int do_mix(wbx_engine* e, uint32_t flags, bool sharded, bool defer_signal = false) {
  CU(e, cudaSetDevice(e->device));
  const uint32_t N = e->n_tracks, B = e->B, C = e->C, K = e->n_blocks;
  const size_t bus_floats = (size_t)C * K * B;
  const size_t peak_floats = (size_t)K * N * 2;
  e->levels_queued = false;
  e->shard_result = false;
  if (N == 0 && !sharded) {  // Engine::process with no tracks: output_buffer.clear() (engine.cpp:1598)
    CU(e, cudaMemsetAsync(e->d_bus.p, 0, bus_floats * sizeof(float), e->stream));
    for (uint32_t c = 0; c < C; c++)
      if (e->mirror[c]) CU(e, cudaMemsetAsync(e->mirror[c], 0, (size_t)K * B * sizeof(float), e->stream));
    snprintf(e->kernel_name, sizeof(e->kernel_name), "clear");
    e->mixed = true;
    return WBX_OK;
  }
  int fpl;
  uint32_t groups;
  choose_shape(e, K, &fpl, &groups);
  const uint32_t T = 32u * fpl;
  const uint32_t n_tiles = (B + T - 1) / T;
  const uint32_t tpg = N ? (N + groups - 1) / groups : 0;  // N == 0 (a sharded rank without tracks): silent tiles
  groups = N ? (N + tpg - 1) / tpg : 1;
  int rc;
  // one zero-filled region per mix: work / arrival counters | block peaks | levels
  const size_t n_counters = 1 + (size_t)K * n_tiles;
  const size_t counter_bytes = ((n_counters * sizeof(uint32_t)) + 255) & ~(size_t)255;
  const size_t peak_bytes = ((peak_floats * sizeof(float)) + 255) & ~(size_t)255;
  const size_t level_bytes = (size_t)N * 2 * sizeof(float);
  if ((rc = dev_reserve(e, e->d_zero, counter_bytes + peak_bytes + level_bytes + 256))) return rc;
  if (groups > 1)
    if ((rc = dev_reserve(e, e->d_ws, (size_t)K * n_tiles * groups * 2 * T * sizeof(float)))) return rc;
  e->counters_ptr = (uint32_t*)e->d_zero.p;
  e->peaks_ptr = (float*)((uint8_t*)e->d_zero.p + counter_bytes);
  e->levels_ptr = (float*)((uint8_t*)e->d_zero.p + counter_bytes + peak_bytes);
  const size_t zero_bytes = (counter_bytes + peak_bytes + level_bytes + 15) & ~(size_t)15;
  void* host_view = nullptr;
  if (e->ingest_bytes && cudaHostGetDevicePointer(&host_view, e->h_spans.p, 0) == cudaSuccess && host_view) {
    CU(e, launch_ingest(host_view, e->d_spans.p, e->ingest_bytes, e->d_zero.p, zero_bytes, e->stream));
    e->launches++;
    CU(e, cudaEventRecord(e->staging_read, e->stream));
    e->staging_busy = true;
  } else {
    if (e->ingest_bytes) {
      CU(e, cudaMemcpyAsync(e->d_spans.p, e->h_spans.p, e->ingest_bytes, cudaMemcpyHostToDevice, e->stream));
      CU(e, cudaEventRecord(e->staging_read, e->stream));
      e->staging_busy = true;
    }
    CU(e, cudaMemsetAsync(e->d_zero.p, 0, zero_bytes, e->stream));
  }
  e->ingest_bytes = 0;

  MixParams p;
  p.spans = (const DSpan*)e->d_spans.p;
  p.cells = e->cells_ptr;
  p.gains = e->gains_ptr;
  p.poly = (const float*)e->d_poly.p;
  p.bus = (float*)e->d_bus.p;
  p.peaks = e->peaks_ptr;
  p.ws = (float*)e->d_ws.p;
  p.counters = e->counters_ptr;
  p.n_tracks = N;
  p.n_blocks = K;
  p.slots = e->slots;
  p.B = B;
  p.C = C;
  p.n_tiles = n_tiles;
  p.groups = groups;
  p.tracks_per_group = tpg;
  p.n_items = K * n_tiles * groups;
  p.clamp = (flags & WBX_MIX_NO_CLAMP) ? 0u : 1u;
  p.ext = e->seg_flags ? 1u : 0u;
  p.one = 1.0f;
  p.mirror[0] = e->mirror[0];
  p.mirror[1] = e->mirror[1];
  for (uint32_t j = 0; j < kMaxPeers; j++) p.xchg[j] = nullptr;
  p.shard_blocks = 0;
  p.shard_rank = 0;
  const Shard& sh = e->shard;
  const uint32_t Ks = sharded ? (K + sh.world - 1) / sh.world : 0;
  if (sharded) {
    // every rank's unclamped tiles go straight into the owner's exchange buffer; the clamp follows the reduce
    for (uint32_t j = 0; j < sh.world; j++) p.xchg[j] = (float*)((uint8_t*)sh.peer_block[j] + sh.xchg_off);
    p.shard_blocks = Ks;
    p.shard_rank = sh.rank;
    p.clamp = 0;
    p.mirror[0] = p.mirror[1] = nullptr;
  }
  int ctas = 0;
  CU(e, launch_mix(p, fpl, e->n_sm, e->stream, &ctas));
  e->launches++;
  snprintf(e->kernel_name, sizeof(e->kernel_name), "%s/fpl%d/g%u/ctas%d%s", groups == 1 ? "exact" : "tree", fpl, groups, ctas,
           sharded ? "/peer-reduce" : "");
  if (sharded) {
    // all of this rank's tiles are on their way into the owners' exchange buffers: tell every rank (phase 0 ends)
    if (!defer_signal) {  // (the fused exchange kernel signals the arrival itself)
      const ShardPeers peers = shard_peers(e->shard);
      CU(e, launch_shard_signal(peers, sh.rank, sh.world, ++e->shard.epoch, e->stream));
      e->launches++;
    }
    e->shard_phase = 1;
    e->shard_result = true;
  }
  e->mixed = true;
  return WBX_OK;
}

// phase 1 of a sharded mix: wait until every rank's tiles have landed here, reduce this rank's slice of callbacks into
// rank 0's master bus (rank order, then the clamp), tell every rank; phase 2: wait until every slice is in the master
// bus — which also keeps a fast rank from overwriting an exchange buffer that is still being read.
static int shard_phase(wbx_engine* e, int phase) {
  Shard& sh = e->shard;
  if (e->shard_phase != phase) return fail(e, WBX_ERR_INVALID, "sharded mix: phase %d called in phase %d", phase, e->shard_phase);
  CU(e, cudaSetDevice(e->device));
  ShardPeers peers = shard_peers(sh);
  if ((uint64_t)e->n_blocks * e->B > sh.host_out_frames) peers.host_dst[0] = peers.host_dst[1] = nullptr;
  CU(e, launch_shard_wait(peers, sh.rank, sh.world, sh.epoch, sh.timeout_ns, sh.status, e->stream));
  e->launches++;
  if (phase == 1) {
    const uint32_t K = e->n_blocks, B = e->B;
    const uint32_t Ks = (K + sh.world - 1) / sh.world;
    const uint64_t k0 = (uint64_t)sh.rank * Ks;
    const uint64_t valid = k0 < K ? ((k0 + Ks < K ? Ks : K - k0) * (uint64_t)B) : 0;
    CU(e, launch_shard_reduce((const float*)((uint8_t*)sh.block + sh.xchg_off), sh.world, e->C, (uint64_t)Ks * B, valid, peers,
                              (uint64_t)K * B, k0 * B, e->n_sm, e->stream));
    CU(e, launch_shard_signal(peers, sh.rank, sh.world, ++sh.epoch, e->stream));
    e->launches += valid ? 2 : 1;
    e->shard_phase = 2;
  } else {
    e->shard_phase = 0;
  }
  return WBX_OK;
}

int wbx_mix(wbx_engine* e, uint32_t flags) {
  if (!e) return WBX_ERR_INVALID;
  if (!e->submitted) return fail(e, WBX_ERR_INVALID, "wbx_mix before wbx_submit");
  return do_mix(e, flags, false);
}

int wbx_mix_sharded_phase(wbx_engine* e, int phase) {
  if (!e) return WBX_ERR_INVALID;
  if (phase == 1 || phase == 2) {
    const int rc = shard_phase(e, phase);
    if (rc && rc != WBX_ERR_INVALID) e->shard_phase = 0;  // a failed launch ends this collective (see wbx_shard_reset)
    return rc;
  }
  if (phase != 0 && phase != 3) return fail(e, WBX_ERR_INVALID, "wbx_mix_sharded_phase: phase %d", phase);
  if (!e->submitted) return fail(e, WBX_ERR_INVALID, "wbx_mix_sharded before wbx_submit");
  const Shard& sh = e->shard;
  if (!sh.on || !sh.connected) return fail(e, WBX_ERR_INVALID, "wbx_mix_sharded: call wbx_shard_init and wbx_shard_connect_* first");
  if (sh.B != e->B || sh.C != e->C) return fail(e, WBX_ERR_INVALID, "wbx_mix_sharded: engine reconfigured since wbx_shard_init");
  if (e->n_blocks > sh.max_blocks)
    return fail(e, WBX_ERR_INVALID, "wbx_mix_sharded: %u callbacks > max_blocks %u of wbx_shard_init", e->n_blocks, sh.max_blocks);
  if (e->shard_phase != 0) return fail(e, WBX_ERR_INVALID, "wbx_mix_sharded: the previous sharded mix stopped in phase %d", e->shard_phase);
  int rc = do_mix(e, 0, true, phase == 3);
  if (!rc && phase == 3) {  // phases 1 + 2 (and phase 0's signal) in one launch
    Shard& s2 = e->shard;
    ShardPeers peers = shard_peers(s2);
    if ((uint64_t)e->n_blocks * e->B > s2.host_out_frames) peers.host_dst[0] = peers.host_dst[1] = nullptr;
    const uint32_t K = e->n_blocks, B = e->B;
    const uint32_t Ks = (K + s2.world - 1) / s2.world;
    const uint64_t k0 = (uint64_t)s2.rank * Ks;
    const uint64_t valid = k0 < K ? ((k0 + Ks < K ? Ks : K - k0) * (uint64_t)B) : 0;
    const uint32_t e1 = ++s2.epoch, e2 = ++s2.epoch;
    cudaError_t err = launch_shard_exchange((const float*)((uint8_t*)s2.block + s2.xchg_off), s2.world, e->C, (uint64_t)Ks * B, valid,
                                            peers, (uint64_t)K * B, k0 * B, s2.rank, e1, e2, s2.timeout_ns, s2.status, s2.done_counter,
                                            e->n_sm, e->stream);
    if (err != cudaSuccess) rc = fail(e, WBX_ERR_CUDA, "shard exchange launch failed: %s", cudaGetErrorString(err));
    e->launches++;
    e->shard_phase = 0;
  }
  if (rc) e->shard_phase = 0;
  return rc;
}

// One process or thread per GPU: the mix, then the WHOLE exchange (arrival signal, wait, owner reduce + clamp, completion
// signal, wait) as one kernel launch. WBX_SHARD_FUSED=0 keeps the five separate launches of the phased form.
int wbx_mix_sharded(wbx_engine* e) {
  static const bool fused = !(getenv("WBX_SHARD_FUSED") && atoi(getenv("WBX_SHARD_FUSED")) == 0);
  if (!fused) {
    int rc = wbx_mix_sharded_phase(e, 0);
    if (!rc) rc = wbx_mix_sharded_phase(e, 1);
    if (!rc) rc = wbx_mix_sharded_phase(e, 2);
    return rc;
  }
  return wbx_mix_sharded_phase(e, 3);
}

int wbx_shard_reset(wbx_engine* e) {
  if (!e) return WBX_ERR_INVALID;
  Shard& sh = e->shard;
  if (!sh.on) return fail(e, WBX_ERR_INVALID, "wbx_shard_reset before wbx_shard_init");
  CU(e, cudaSetDevice(e->device));
  cudaStreamSynchronize(e->stream);  // a failed launch may have left a sticky-free error behind: drain, then clear
  cudaGetLastError();
  CU(e, cudaMemset(sh.block, 0, kShardHeaderBytes));  // this rank's arrival words
  sh.epoch = 0;
  if (sh.status) *sh.status = 0;
  e->shard_phase = 0;
  e->shard_result = false;
  e->mixed = false;
  return WBX_OK;
}

// the bus the last mix produced: the engine's own, or — after a sharded mix — the master bus (rank 0 only)
static float* result_bus(wbx_engine* e) {
  if (!e->shard_result) return (float*)e->d_bus.p;
  return e->shard.rank == 0 ? (float*)((uint8_t*)e->shard.block + e->shard.bus_off) : nullptr;
}

static int check_shard_status(wbx_engine* e) {
  if (e->shard.on && e->shard.status && *e->shard.status) {
    *e->shard.status = 0;
    e->shard_phase = 0;
    return fail(e, WBX_ERR_CUDA, "sharded render: a peer rank did not reach the bus exchange within %.1f s",
                (double)e->shard.timeout_ns * 1e-9);
  }
  return WBX_OK;
}

// enqueue the VUMeter::level reduce of the last mix and its copy into page-locked h_levels (no synchronise)
static int queue_levels(wbx_engine* e) {
  const uint32_t NC = e->n_tracks * 2;
  if (NC == 0 || e->levels_queued) return WBX_OK;
  int rc;
  if ((rc = host_reserve(e, e->h_levels, NC * sizeof(float)))) return rc;
  void* host_view = nullptr;
  if (e->n_blocks <= 32 && cudaHostGetDevicePointer(&host_view, e->h_levels.p, 0) == cudaSuccess && host_view) {
    // few callbacks: one thread per (track, channel) stores its level straight into the page-locked result
    CU(e, launch_levels_direct(e->peaks_ptr, e->n_blocks, NC, (float*)host_view, e->stream));
    e->launches++;
  } else {
    // levels were zeroed with the rest of the mix's zero region; the reduce is a max, so running it twice is harmless
    CU(e, launch_levels(e->peaks_ptr, e->n_blocks, NC, e->levels_ptr, e->stream));
    e->launches++;
    CU(e, cudaMemcpyAsync(e->h_levels.p, e->levels_ptr, NC * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  }
  e->levels_queued = true;
  return WBX_OK;
}

int wbx_fetch(wbx_engine* e, float* const* out_channels, float* peaks) {
  if (!e) return WBX_ERR_INVALID;
  if (!e->mixed) return fail(e, WBX_ERR_INVALID, "wbx_fetch before wbx_mix");
  if (e->shard_phase != 0) return fail(e, WBX_ERR_INVALID, "wbx_fetch: the sharded mix stopped in phase %d", e->shard_phase);
  CU(e, cudaSetDevice(e->device));
  const size_t chan_floats = (size_t)e->n_blocks * e->B;
  const size_t bus_floats = chan_floats * e->C;
  const size_t peak_floats = (size_t)e->n_blocks * e->n_tracks * 2;
  int rc;
  bool staged_bus = false, staged_peaks = false;
  if (out_channels) {
    const float* bus = result_bus(e);
    if (!bus) return fail(e, WBX_ERR_INVALID, "wbx_fetch: after a sharded mix only rank 0 holds the master bus");
    bool direct = true;  // page-locked caller buffers (wbx_host_alloc) take the D2H copy directly
    bool written = true;  // ... unless the mix kernel already wrote them (wbx_render's mirror)
    const bool by_owners = e->shard_result && chan_floats <= e->shard.host_out_frames;
    for (uint32_t c = 0; c < e->C; c++) {
      direct = direct && out_channels[c] && is_pinned(out_channels[c]);
      written = written && out_channels[c] &&
                ((e->mirror[c] && e->mirror_host[c] == out_channels[c]) ||
                 (by_owners && e->shard.host_out[c] && e->shard.host_out_host[c] == out_channels[c]));
    }
    if (written) {
    } else if (direct) {
      for (uint32_t c = 0; c < e->C; c++)
        CU(e, cudaMemcpyAsync(out_channels[c], bus + c * chan_floats, chan_floats * sizeof(float),
                              cudaMemcpyDeviceToHost, e->stream));
    } else {
      if ((rc = host_reserve(e, e->h_bus, bus_floats * sizeof(float)))) return rc;
      CU(e, cudaMemcpyAsync(e->h_bus.p, bus, bus_floats * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
      staged_bus = true;
    }
  }
  if (peaks && peak_floats) {
    if (is_pinned(peaks)) {
      CU(e, cudaMemcpyAsync(peaks, e->peaks_ptr, peak_floats * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    } else {
      if ((rc = host_reserve(e, e->h_peaks, peak_floats * sizeof(float)))) return rc;
      CU(e, cudaMemcpyAsync(e->h_peaks.p, e->peaks_ptr, peak_floats * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
      staged_peaks = true;
    }
  }
  CU(e, cudaStreamSynchronize(e->stream));
  if ((rc = check_shard_status(e))) return rc;
  if (staged_bus)
    for (uint32_t c = 0; c < e->C; c++)
      if (out_channels[c]) memcpy(out_channels[c], (const float*)e->h_bus.p + c * chan_floats, chan_floats * sizeof(float));
  if (staged_peaks) memcpy(peaks, e->h_peaks.p, peak_floats * sizeof(float));
  return WBX_OK;
}

int wbx_fetch_levels(wbx_engine* e, float* levels) {
  if (!e || !levels) return WBX_ERR_INVALID;
  if (!e->mixed) return fail(e, WBX_ERR_INVALID, "wbx_fetch_levels before wbx_mix");
  const uint32_t NC = e->n_tracks * 2;
  if (NC == 0) return WBX_OK;
  CU(e, cudaSetDevice(e->device));
  int rc;
  if ((rc = queue_levels(e))) return rc;  // no-op when wbx_render_levels / an earlier call already queued it
  CU(e, cudaStreamSynchronize(e->stream));
  memcpy(levels, e->h_levels.p, NC * sizeof(float));
  return WBX_OK;
}

void* wbx_host_alloc(size_t bytes) {
  void* p = nullptr;
  // portable + mapped: every device of the process can copy to / store into it (sharded renders mirror into it)
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

void wbx_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int wbx_fetch_interleaved(wbx_engine* e, void* dst, int dst_format) {
  if (!e || !dst) return WBX_ERR_INVALID;
  if (!e->mixed) return fail(e, WBX_ERR_INVALID, "wbx_fetch_interleaved before wbx_mix");
  size_t es;
  switch (dst_format) {
    case WBX_FMT_I16: es = 2; break;
    case WBX_FMT_I24: es = 3; break;
    case WBX_FMT_I24_X8:
    case WBX_FMT_I32:
    case WBX_FMT_F32: es = 4; break;
    default: return fail(e, WBX_ERR_UNSUPPORTED, "device format %d", dst_format);
  }
  CU(e, cudaSetDevice(e->device));
  const uint64_t frames = (uint64_t)e->n_blocks * e->B;
  // the reference's packed-I24 writer has no channel stride (audio_format_conv.cpp:31-41): frames*3 bytes
  const size_t bytes = dst_format == WBX_FMT_I24 ? frames * 3 : frames * e->C * es;
  int rc;
  if ((rc = dev_reserve(e, e->d_conv, bytes))) return rc;
  if ((rc = host_reserve(e, e->h_conv, bytes))) return rc;
  const float* bus = result_bus(e);
  if (!bus) return fail(e, WBX_ERR_INVALID, "wbx_fetch_interleaved: after a sharded mix only rank 0 holds the master bus");
  CU(e, launch_interleave(bus, frames, e->C, dst_format, e->d_conv.p, e->n_sm, e->stream));
  e->launches++;
  CU(e, cudaMemcpyAsync(e->h_conv.p, e->d_conv.p, bytes, cudaMemcpyDeviceToHost, e->stream));
  CU(e, cudaStreamSynchronize(e->stream));
  memcpy(dst, e->h_conv.p, bytes);
  return WBX_OK;
}

// ---- offline bounce: chunks in the device format, their copy-out overlapping the next chunk's mix ---------------------
static size_t conv_elem_size(int fmt) {
  switch (fmt) {
    case WBX_FMT_I16: return 2;
    case WBX_FMT_I24: return 3;
    case WBX_FMT_I24_X8:
    case WBX_FMT_I32:
    case WBX_FMT_F32: return 4;
    default: return 0;
  }
}

int wbx_bounce_begin(wbx_engine* e, int dst_format) {
  if (!e) return WBX_ERR_INVALID;
  if (!conv_elem_size(dst_format)) return fail(e, WBX_ERR_UNSUPPORTED, "device format %d", dst_format);
  CU(e, cudaSetDevice(e->device));
  if (!e->copy_stream) CU(e, cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; i++) {
    if (!e->bounce_ready[i]) CU(e, cudaEventCreateWithFlags(&e->bounce_ready[i], cudaEventDisableTiming));
    if (!e->bounce_done[i]) CU(e, cudaEventCreateWithFlags(&e->bounce_done[i], cudaEventDisableTiming));
  }
  CU(e, cudaStreamSynchronize(e->stream));
  CU(e, cudaStreamSynchronize(e->copy_stream));
  e->bounce_fmt = dst_format;
  e->bounce_pushed = e->bounce_popped = 0;
  e->bounce_on = true;
  return WBX_OK;
}

int wbx_bounce_push(wbx_engine* e) {
  if (!e) return WBX_ERR_INVALID;
  if (!e->bounce_on) return fail(e, WBX_ERR_INVALID, "wbx_bounce_push before wbx_bounce_begin");
  if (!e->mixed) return fail(e, WBX_ERR_INVALID, "wbx_bounce_push before wbx_mix");
  if (e->bounce_pushed - e->bounce_popped >= 2) return fail(e, WBX_ERR_INVALID, "wbx_bounce_push: two chunks are already in flight, pop one");
  CU(e, cudaSetDevice(e->device));
  const float* bus = result_bus(e);
  if (!bus) return fail(e, WBX_ERR_INVALID, "wbx_bounce_push: after a sharded mix only rank 0 holds the master bus");
  const int slot = (int)(e->bounce_pushed & 1);
  const uint64_t frames = (uint64_t)e->n_blocks * e->B;
  const size_t bytes = e->bounce_fmt == WBX_FMT_I24 ? frames * 3 : frames * e->C * conv_elem_size(e->bounce_fmt);
  if (bytes > e->d_bounce[slot].cap || bytes > e->h_bounce[slot].cap) {  // (re)size this slot: its previous copy-out is long popped
    CU(e, cudaStreamSynchronize(e->copy_stream));
    int rc;
    if ((rc = dev_reserve(e, e->d_bounce[slot], bytes))) return rc;
    if ((rc = host_reserve(e, e->h_bounce[slot], bytes))) return rc;
  }
  // the slot's previous copy-out (two chunks ago) must be over before the conversion overwrites the device buffer
  if (e->bounce_pushed >= 2) CU(e, cudaStreamWaitEvent(e->stream, e->bounce_done[slot], 0));
  CU(e, launch_interleave(bus, frames, e->C, e->bounce_fmt, e->d_bounce[slot].p, e->n_sm, e->stream));
  e->launches++;
  CU(e, cudaEventRecord(e->bounce_ready[slot], e->stream));
  // ... and the copy to the host runs on the second stream, under the next chunk's submit + mix
  CU(e, cudaStreamWaitEvent(e->copy_stream, e->bounce_ready[slot], 0));
  CU(e, cudaMemcpyAsync(e->h_bounce[slot].p, e->d_bounce[slot].p, bytes, cudaMemcpyDeviceToHost, e->copy_stream));
  CU(e, cudaEventRecord(e->bounce_done[slot], e->copy_stream));
  e->bounce_bytes[slot] = bytes;
  e->bounce_pushed++;
  return WBX_OK;
}

int wbx_bounce_pop(wbx_engine* e, const void** data, size_t* bytes) {
  if (!e || !data || !bytes) return WBX_ERR_INVALID;
  if (!e->bounce_on || e->bounce_popped == e->bounce_pushed) return fail(e, WBX_ERR_INVALID, "wbx_bounce_pop: no chunk in flight");
  CU(e, cudaSetDevice(e->device));
  const int slot = (int)(e->bounce_popped & 1);
  CU(e, cudaEventSynchronize(e->bounce_done[slot]));
  *data = e->h_bounce[slot].p;
  *bytes = e->bounce_bytes[slot];
  e->bounce_popped++;
  return WBX_OK;
}

int wbx_render_levels(wbx_engine* e, const wbx_segment* segs, uint32_t n_segs, const float* track_gains,
                      uint32_t n_blocks, float* const* out_channels, float* peaks, float* levels) {
  int rc = wbx_submit(e, segs, n_segs, track_gains, n_blocks);
  if (rc) return rc;
  const bool sharded = e->shard.on && e->shard.connected;
  // Page-locked caller channels (wbx_host_alloc) are written by the mix kernel itself, tile by tile, next to the
  // device bus: no device-to-host copy of the bus follows the kernel.
  bool direct = out_channels != nullptr && !sharded;
  float* dview[2] = {nullptr, nullptr};
  for (uint32_t c = 0; direct && c < e->C; c++) {
    // the kernel stores bus tiles as float2: a channel that is only 4-byte aligned takes the copy path instead
    direct = out_channels[c] && ((uintptr_t)out_channels[c] & 7u) == 0 && is_pinned(out_channels[c]) &&
             cudaHostGetDevicePointer((void**)&dview[c], out_channels[c], 0) == cudaSuccess && dview[c] &&
             ((uintptr_t)dview[c] & 7u) == 0;
    if (!direct) cudaGetLastError();
  }
  for (uint32_t c = 0; c < 2; c++) {
    e->mirror[c] = (direct && c < e->C) ? dview[c] : nullptr;
    e->mirror_host[c] = (direct && c < e->C) ? out_channels[c] : nullptr;
  }
  rc = sharded ? wbx_mix_sharded(e) : wbx_mix(e, 0);
  if (!rc && levels) rc = queue_levels(e);
  if (!rc) rc = wbx_fetch(e, (sharded && e->shard.rank != 0) ? nullptr : out_channels, peaks);  // the one synchronise
  e->mirror[0] = e->mirror[1] = nullptr;
  e->mirror_host[0] = e->mirror_host[1] = nullptr;
  if (rc) return rc;
  if (levels && e->n_tracks) memcpy(levels, e->h_levels.p, (size_t)e->n_tracks * 2 * sizeof(float));
  return WBX_OK;
}

int wbx_render(wbx_engine* e, const wbx_segment* segs, uint32_t n_segs, const float* track_gains,
               uint32_t n_blocks, float* const* out_channels, float* peaks) {
  return wbx_render_levels(e, segs, n_segs, track_gains, n_blocks, out_channels, peaks, nullptr);
}

// ---- sharded render (tracks split over the GPUs of one box) -------------------------------------------------------
int wbx_shard_init(wbx_engine* e, uint32_t rank, uint32_t world, uint32_t max_blocks, void* ipc_handle_out) {
  if (!e) return WBX_ERR_INVALID;
  if (world < 1 || world > kMaxPeers || rank >= world || max_blocks == 0)
    return fail(e, WBX_ERR_INVALID, "wbx_shard_init: rank %u / world %u (max %u) / max_blocks %u", rank, world, kMaxPeers,
                max_blocks);
  CU(e, cudaSetDevice(e->device));
  int rc = wbx_shard_close(e);
  if (rc) return rc;
  Shard& sh = e->shard;
  const size_t Ks = (max_blocks + world - 1) / world;
  const size_t xchg_bytes = (((size_t)world * e->C * Ks * e->B * sizeof(float)) + 255) & ~(size_t)255;
  const size_t bus_bytes = (((size_t)e->C * max_blocks * e->B * sizeof(float)) + 255) & ~(size_t)255;
  sh.xchg_off = kShardHeaderBytes;
  sh.bus_off = kShardHeaderBytes + xchg_bytes;
  sh.bytes = sh.bus_off + bus_bytes;
  cudaError_t err = cudaMalloc(&sh.block, sh.bytes);
  if (err != cudaSuccess) return fail(e, WBX_ERR_NOMEM, "cudaMalloc(%zu) for the shard block failed: %s", sh.bytes, cudaGetErrorString(err));
  CU(e, cudaMemset(sh.block, 0, sh.bytes));
  if (cudaHostAlloc((void**)&sh.status, sizeof(uint32_t), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
    cudaFree(sh.block);
    sh.block = nullptr;
    return fail(e, WBX_ERR_NOMEM, "cudaHostAlloc for the shard status word failed");
  }
  *sh.status = 0;
  if (cudaMalloc((void**)&sh.done_counter, 256) != cudaSuccess || cudaMemset(sh.done_counter, 0, 256) != cudaSuccess) {
    cudaFree(sh.block);
    cudaFreeHost(sh.status);
    sh.block = nullptr;
    sh.status = nullptr;
    return fail(e, WBX_ERR_NOMEM, "cudaMalloc for the shard exchange counter failed");
  }
  if (const char* t = getenv("WBX_SHARD_TIMEOUT_MS"))
    if (atof(t) > 0) sh.timeout_ns = (unsigned long long)(atof(t) * 1e6);
  if (ipc_handle_out) {
    cudaIpcMemHandle_t h;
    static_assert(sizeof(h) == WBX_IPC_HANDLE_BYTES, "IPC handle size");
    err = cudaIpcGetMemHandle(&h, sh.block);
    if (err != cudaSuccess) {
      cudaGetLastError();
      memset(ipc_handle_out, 0, WBX_IPC_HANDLE_BYTES);  // same-process use (wbx_shard_connect_local) still works
    } else {
      memcpy(ipc_handle_out, &h, WBX_IPC_HANDLE_BYTES);
    }
  }
  sh.rank = rank;
  sh.world = world;
  sh.max_blocks = max_blocks;
  sh.B = e->B;
  sh.C = e->C;
  sh.epoch = 0;
  sh.on = true;
  sh.connected = false;
  return WBX_OK;
}

int wbx_shard_connect_ipc(wbx_engine* e, const void* handles) {
  if (!e || !handles) return WBX_ERR_INVALID;
  Shard& sh = e->shard;
  if (!sh.on) return fail(e, WBX_ERR_INVALID, "wbx_shard_connect_ipc before wbx_shard_init");
  CU(e, cudaSetDevice(e->device));
  for (uint32_t j = 0; j < sh.world; j++) {
    if (j == sh.rank) {
      sh.peer_block[j] = sh.block;
      sh.peer_ipc[j] = false;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const uint8_t*)handles + (size_t)j * WBX_IPC_HANDLE_BYTES, sizeof(h));
    cudaError_t err = cudaIpcOpenMemHandle(&sh.peer_block[j], h, cudaIpcMemLazyEnablePeerAccess);
    if (err != cudaSuccess) {
      cudaGetLastError();
      return fail(e, WBX_ERR_CUDA, "cudaIpcOpenMemHandle(rank %u) failed: %s", j, cudaGetErrorString(err));
    }
    sh.peer_ipc[j] = true;
  }
  sh.connected = true;
  return WBX_OK;
}

int wbx_shard_connect_local(wbx_engine* e, wbx_engine* const* engines) {
  if (!e || !engines) return WBX_ERR_INVALID;
  Shard& sh = e->shard;
  if (!sh.on) return fail(e, WBX_ERR_INVALID, "wbx_shard_connect_local before wbx_shard_init");
  CU(e, cudaSetDevice(e->device));
  for (uint32_t j = 0; j < sh.world; j++) {
    const wbx_engine* pe = engines[j];
    if (!pe || !pe->shard.on || pe->shard.world != sh.world || pe->shard.rank != j || pe->shard.bytes != sh.bytes)
      return fail(e, WBX_ERR_INVALID, "wbx_shard_connect_local: engine %u is not rank %u of the same sharded setup", j, j);
    if (pe->device != e->device) {
      int can = 0;
      CU(e, cudaDeviceCanAccessPeer(&can, e->device, pe->device));
      if (!can) return fail(e, WBX_ERR_UNSUPPORTED, "device %d cannot access device %d's memory", e->device, pe->device);
      cudaError_t err = cudaDeviceEnablePeerAccess(pe->device, 0);
      if (err != cudaSuccess && err != cudaErrorPeerAccessAlreadyEnabled)
        return fail(e, WBX_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d) failed: %s", pe->device, cudaGetErrorString(err));
      cudaGetLastError();
    }
    sh.peer_block[j] = pe->shard.block;
    sh.peer_ipc[j] = false;
  }
  sh.connected = true;
  return WBX_OK;
}

int wbx_shard_set_host_output(wbx_engine* e, float* const* channels, uint64_t frames_per_channel) {
  if (!e) return WBX_ERR_INVALID;
  Shard& sh = e->shard;
  if (!sh.on) return fail(e, WBX_ERR_INVALID, "wbx_shard_set_host_output before wbx_shard_init");
  CU(e, cudaSetDevice(e->device));
  for (int c = 0; c < 2; c++) sh.host_out[c] = sh.host_out_host[c] = nullptr;
  sh.host_out_frames = 0;
  if (!channels || frames_per_channel == 0) return WBX_OK;
  for (uint32_t c = 0; c < e->C; c++) {
    void* dv = nullptr;
    if (!channels[c] || !is_pinned(channels[c]) || cudaHostGetDevicePointer(&dv, channels[c], 0) != cudaSuccess || !dv) {
      cudaGetLastError();
      for (int q = 0; q < 2; q++) sh.host_out[q] = sh.host_out_host[q] = nullptr;
      return fail(e, WBX_ERR_INVALID, "wbx_shard_set_host_output: channel %u is not page-locked memory mapped on device %d", c, e->device);
    }
    sh.host_out[c] = (float*)dv;
    sh.host_out_host[c] = channels[c];
  }
  sh.host_out_frames = frames_per_channel;
  return WBX_OK;
}

int wbx_host_register(void* p, size_t bytes) {
  if (!p || !bytes) return WBX_ERR_INVALID;
  if (cudaHostRegister(p, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped) != cudaSuccess) {
    cudaGetLastError();
    return WBX_ERR_CUDA;
  }
  return WBX_OK;
}

int wbx_host_unregister(void* p) {
  if (!p) return WBX_ERR_INVALID;
  if (cudaHostUnregister(p) != cudaSuccess) {
    cudaGetLastError();
    return WBX_ERR_CUDA;
  }
  return WBX_OK;
}

int wbx_shard_close(wbx_engine* e) {
  if (!e) return WBX_ERR_INVALID;
  Shard& sh = e->shard;
  if (!sh.on) return WBX_OK;
  CU(e, cudaSetDevice(e->device));
  CU(e, cudaStreamSynchronize(e->stream));
  for (uint32_t j = 0; j < kMaxPeers; j++) {
    if (sh.peer_ipc[j] && sh.peer_block[j]) cudaIpcCloseMemHandle(sh.peer_block[j]);
    sh.peer_block[j] = nullptr;
    sh.peer_ipc[j] = false;
  }
  if (sh.block) cudaFree(sh.block);
  if (sh.status) cudaFreeHost(sh.status);
  if (sh.done_counter) cudaFree(sh.done_counter);
  sh = Shard();
  e->shard_result = false;
  e->shard_phase = 0;
  return WBX_OK;
}

int wbx_shard_info(const wbx_engine* e, uint32_t* rank, uint32_t* world) {
  if (!e) return WBX_ERR_INVALID;
  const bool on = e->shard.on && e->shard.connected;
  if (rank) *rank = on ? e->shard.rank : 0;
  if (world) *world = on ? e->shard.world : 1;
  return WBX_OK;
}

int wbx_device_bus(wbx_engine* e, float** d_bus, uint64_t* n_floats) {
  if (!e || !e->submitted) return WBX_ERR_INVALID;
  if (d_bus) *d_bus = e->mixed ? result_bus(e) : (float*)e->d_bus.p;
  if (n_floats) *n_floats = (uint64_t)e->C * e->n_blocks * e->B;
  return WBX_OK;
}

int wbx_device_peaks(wbx_engine* e, float** d_peaks, uint64_t* n_floats) {
  if (!e || !e->mixed) return WBX_ERR_INVALID;
  if (d_peaks) *d_peaks = e->peaks_ptr;
  if (n_floats) *n_floats = (uint64_t)e->n_blocks * e->n_tracks * 2;
  return WBX_OK;
}

int wbx_clamp_device(wbx_engine* e, float* d_bus, uint64_t n_floats) {
  if (!e || (!d_bus && n_floats)) return WBX_ERR_INVALID;
  CU(e, cudaSetDevice(e->device));
  CU(e, launch_clamp(d_bus, n_floats, e->n_sm, e->stream));
  e->launches++;
  return WBX_OK;
}

int wbx_synchronize(wbx_engine* e) {
  if (!e) return WBX_ERR_INVALID;
  CU(e, cudaSetDevice(e->device));
  CU(e, cudaStreamSynchronize(e->stream));
  return check_shard_status(e);
}

uint64_t wbx_launch_count(const wbx_engine* e) { return e ? e->launches : 0; }
int wbx_fir_split_factor(void) { return fir_tc_split_factor(); }

int wbx_fir_path(const wbx_engine* e) { return e ? e->fir_mode : WBX_ERR_INVALID; }
const char* wbx_last_kernel(const wbx_engine* e) { return e ? e->kernel_name : ""; }

}  // extern "C"
