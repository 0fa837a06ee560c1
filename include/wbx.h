/*
 * wbx.h — C ABI of the B200-native mixing hot path of native-m/whitebox.
 *
 * This is the drop-in boundary: everything in the reference from dsp::Sampler::stream down to the master
 * bus clamp runs behind these calls as hand-written sm_100a CUDA; everything above it (transport, clip
 * scheduling in doubles, parameter messages) stays host code. All file:line citations are relative to the
 * reference tree (/root/reference/src).
 *
 *   reference                                                       this ABI
 *   ---------------------------------------------------------------------------------------------------
 *   Engine::set_audio_channel_config      engine/engine.cpp:43-57   wbx_configure
 *   Sample (planar channels + 16 pad)     dsp/sample.h:18-28        wbx_sample_upload / wbx_sample_release
 *   Engine::tracks.size()                 engine/engine.h:40        wbx_set_track_count
 *   AudioEvent + dsp::Sampler state       engine/event.h:66-74,     wbx_segment (one per Sampler::stream call,
 *                                         dsp/sampler.h:14-16       or one per run of identical calls)
 *   volume * pan_coeffs[c]                engine/track.cpp:728-731  track_gains[n_tracks][2]
 *   Sampler::stream                       dsp/sampler.cpp:88-210    \
 *   dsp::apply_gain                       dsp/dsp_ops.h:27-31        |
 *   VUMeter::push_samples                 engine/vu_meter.h:20-30    |  wbx_render (= wbx_submit + wbx_mix
 *   AudioBuffer::clear / ::mix            core/audio_buffer.h:67-82  |              + wbx_fetch)
 *   output clamp                          engine/engine.cpp:1627-36 /
 *   convert_f32_to_interleaved_*          core/audio_format_conv.cpp wbx_fetch_interleaved
 *
 * Conventions: every function returns 0 (WBX_OK) or a negative wbx_status; nothing throws; one thread
 * drives an engine at a time (the reference's audio thread); no allocation happens in steady state (device
 * and pinned staging buffers grow on first use of a larger size and are then reused). There is NO CPU
 * fallback: without a CUDA device wbx_create fails with WBX_ERR_NO_DEVICE.
 */
#ifndef WBX_H
#define WBX_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WBX_ABI_VERSION 3

typedef struct wbx_engine wbx_engine;

typedef enum wbx_status {
  WBX_OK = 0,
  WBX_ERR_INVALID = -1,     /* bad argument / call order */
  WBX_ERR_CUDA = -2,        /* a CUDA runtime call failed; see wbx_last_error */
  WBX_ERR_NOMEM = -3,       /* host or device allocation failed */
  WBX_ERR_UNSUPPORTED = -4, /* e.g. more than 2 output channels (pan_coeffs[2], engine/track.h:50) */
  WBX_ERR_NO_DEVICE = -5    /* no usable sm_100 device: there is no CPU path */
} wbx_status;

/* Source sample element types — numeric values of wb::AudioFormat (core/audio_format.h:7-20).
 * I24 means 24-bit data widened to an int32 container, which is how the sampler reads it
 * (dsp/sampler.cpp:121-132) and how the loader stores it (dsp/sample.cpp:20). */
typedef enum wbx_format { WBX_FMT_I16 = 3, WBX_FMT_I24 = 5, WBX_FMT_I24_X8 = 6, WBX_FMT_I32 = 7, WBX_FMT_F32 = 9 } wbx_format;

/* One dsp::Sampler::stream call (dsp/sampler.cpp:88-210) as Track::process issues it
 * (engine/track.cpp:678,718), or `n_blocks` consecutive identical calls, one per callback.
 *
 * For block b in [block, block + n_blocks) the device renders
 *     dst[c][dst_offset + j] += resample(sample, pos_b + j * speed) * gain        j < min(length, remaining)
 * with pos_block = src_pos and pos_{b+1} = pos_b + (double)length * speed — the exact f64 recurrence of
 * Sampler::sample_offset_ (sampler.cpp:103,209). speed == 1.0 takes the unity-copy branch
 * (sampler.cpp:106-158; position truncated to an integer), anything else the 2-tap linear branch
 * (sampler.cpp:34-59). Output channel c reads source channel c % sample_channels (sampler.cpp:111). */
typedef struct wbx_segment {
  uint32_t track;      /* index into Engine::tracks — also the bus summation order */
  uint32_t block;      /* first callback (0-based within this render) the call belongs to */
  uint32_t n_blocks;   /* run length in callbacks, >= 1 */
  uint32_t dst_offset; /* `buffer_offset` argument: first frame written within the block */
  uint32_t length;     /* `num_samples` argument (before clipping to the sample's end) */
  uint32_t sample_id;  /* from wbx_sample_upload */
  double src_pos;      /* Sampler::sample_offset_ at the first call */
  double speed;        /* Sampler::playback_speed_ = src_rate / dst_rate * clip speed (sampler.h:24), > 0 */
  float gain;          /* AudioClip::gain (engine/clip.h:44) */
  uint32_t flags;      /* WBX_SEG_FADE | WBX_SEG_POLYPHASE or 0 */
  /* Fade envelope — EXTENSION, not in the reference: AudioClip::fade_start / fade_end (engine/clip.h:41-42)
   * are stored and drawn by whitebox but no audio code reads them. Spec (oracle/wb_oracle.c, "parity unpinned"):
   * for clip-relative output frame n = clip_frame + j,
   *     env(n) = (float)( min(1, n / fade_in_frames) * min(1, max(0, (clip_len_frames - n) / fade_out_frames)) )
   * (a factor is 1 when its length is <= 0) and the frame contributes (src * gain) * env. Ignored unless
   * WBX_SEG_FADE is set; with both lengths 0 the path is the reference's, bit for bit. For a run,
   * clip_frame advances by `length` per callback. */
  double clip_frame;       /* output frames since the clip's start at this call's first frame (integer value) */
  double fade_in_frames;   /* beat_to_samples(fade_start) */
  double fade_out_frames;  /* beat_to_samples(fade_end) */
  double clip_len_frames;  /* beat_to_samples(max_time - min_time) */
} wbx_segment;

#define WBX_SEG_FADE 1u
/* Polyphase quality mode — EXTENSION, not in the reference (its only resampler is the 2-tap linear one,
 * dsp/sampler.h:8-11): a call with speed != 1 on a stereo f32 sample into a stereo bus is resampled with a 128-phase
 * x 16-tap Blackman-windowed sinc (oracle/wb_oracle.c sample_polyphase, "parity unpinned") instead of the lerp; the
 * position arithmetic is the reference's. Every other case ignores the flag. */
#define WBX_SEG_POLYPHASE 2u

/* Bus summation order (wbx_set_sum_mode). */
typedef enum wbx_sum_mode {
  WBX_SUM_AUTO = 0,  /* exact when the render is large enough to fill the GPU that way, else tree */
  WBX_SUM_EXACT = 1, /* tracks added 0..N-1 sequentially in f32 per output sample: bit-identical to
                        AudioBuffer::mix order (core/audio_buffer.h:73-82, engine.cpp:1600-1617) */
  WBX_SUM_TREE = 2   /* track groups summed in parallel, group partials added in fixed group order:
                        deterministic, re-associated (within 1e-5 of block peak of the reference) */
} wbx_sum_mode;

/* wbx_mix flags */
#define WBX_MIX_NO_CLAMP 1u /* leave the bus unclamped (partial bus of a track shard, clamp after the reduce) */

/* ---- lifetime / configuration ---------------------------------------------------------------------- */
int wbx_abi_version(void);
int wbx_create(wbx_engine** out, int device_ordinal);
int wbx_destroy(wbx_engine* e);
const char* wbx_last_error(const wbx_engine* e);
/* Engine::set_audio_channel_config(_, out_channels, block_frames, sample_rate). May be called again at
 * any time (device change, engine/config.cpp:198-232); resident samples survive. out_channels is 1 or 2. */
int wbx_configure(wbx_engine* e, uint32_t out_channels, uint32_t block_frames, uint32_t sample_rate);
int wbx_set_track_count(wbx_engine* e, uint32_t n_tracks);
int wbx_set_sum_mode(wbx_engine* e, int mode);
/* Run all work of this engine on an existing CUDA stream (cudaStream_t as void*); NULL = engine's own. */
int wbx_set_stream(wbx_engine* e, void* cuda_stream);

/* ---- resident samples (dsp/sample.h:18-28) -------------------------------------------------------- */
/* planar[c] points to `frames` elements of `format` (host memory). The engine keeps a device copy with
 * >= 16 zero frames of tail padding (dsp/sample.cpp:127,140). */
int wbx_sample_upload(wbx_engine* e, int format, uint32_t channels, uint64_t frames, uint32_t sample_rate,
                      const void* const* planar, uint32_t* out_id);
int wbx_sample_release(wbx_engine* e, uint32_t id);
/* Overwrite the data of an existing sample (same format / channels / frames) — streaming sources from host
 * memory. Asynchronous on the engine's stream when planar[c] are page-locked (wbx_host_alloc): the arrays
 * must then stay valid until the next wbx_synchronize / wbx_fetch; pageable arrays are copied before return. */
int wbx_sample_update(wbx_engine* e, uint32_t id, const void* const* planar);

/* Waveform peak mip-maps of a resident sample — WaveformVisual::create + summarize_for_mipmaps_impl
 * (gfx/waveform_visual.cpp:9-173,181-248), which the reference runs on every sample import
 * (engine/assets_table.cpp:34,56). quality 0 = Low (int8), 1 = High (int16); level l summarises chunks of
 * 2^(2l+1) frames into (first, second) = (min, max) ordered by first occurrence; layout [channels][*count].
 * Copies level `level` to host memory `out` when it fits cap_elems elements (out may be NULL to query *count)
 * and returns the number of levels (>= 0) or a negative wbx_status. */
int wbx_sample_mipmap(wbx_engine* e, uint32_t id, int quality, int level, void* out, uint64_t cap_elems,
                      uint32_t* count);

/* ---- per-track effect chain — EXTENSION, not in the reference ---------------------------------------------- */
/* whitebox has no DSP effects (its per-track slot is a third-party VST3 instance, engine/track.h:124,
 * engine/track.cpp:645-662). This is the built-in chain a PluginFormat::Native effect would provide
 * (plughost/plugin_interface.h:31-35), specified by the builder (oracle/wb_oracle.c apply_effects — "parity
 * unpinned"): 4-band RBJ biquad EQ (band 0 low shelf, 1/2 peaking, 3 high shelf; transposed direct form II in
 * f32 with fused multiply-adds) then a per-channel peak compressor (attack/release one-poles, hard knee, ratio
 * 2/4/8/limiter evaluated with IEEE divide and square roots only), applied to the track's signal after its
 * clips are rendered and before volume/pan and the VU meter. Filter state persists across renders. */
typedef struct wbx_effect_params { /* musical parameters */
  float eq_freq[4], eq_gain_db[4], eq_q[4];
  float comp_threshold_db, comp_attack_ms, comp_release_ms, comp_makeup_db;
  int32_t comp_ratio_code; /* 0 off, 1 = 2:1, 2 = 4:1, 3 = 8:1, 4 = limiter */
  int32_t reverb_on;       /* 1: last stage = convolution with the engine's impulse response (BASELINE cfg 5) */
} wbx_effect_params;
typedef struct wbx_effects { /* designed coefficients */
  uint32_t eq_on, comp_on;
  float b0[4], b1[4], b2[4], a1[4], a2[4]; /* normalised by a0 */
  float comp_threshold, comp_attack, comp_release, comp_makeup; /* linear / one-pole coefficients */
  uint32_t comp_ratio_code;
  uint32_t reverb_on;
} wbx_effects;
/* f64 coefficient design (host only). */
int wbx_effects_design(const wbx_effect_params* params, uint32_t sample_rate, wbx_effects* out);
/* Convolution reverb (extension, BASELINE cfg 5): one impulse response h[0..n_taps) per engine, applied as the last
 * stage of every chain with reverb_on: y[n] = sum_k h[k] * x[n-k] with the history carried across renders.
 * Specification = f64 accumulation (oracle/wb_oracle.c apply_reverb). Three implementations, all held to 1e-5 of the block
 * peak: responses of >= 1024 taps run as a uniformly partitioned FFT convolution (overlap-save, f32, a track's two channels
 * as one complex signal; partitions of 2048 taps for long renders, 512 for short ones; the window spectra persist across
 * renders), shorter ones as a direct form on the CUDA cores (f32 fused multiply-adds per 256-tap tile, tiles summed in
 * f64); the direct form on the tensor cores (the convolution as a Toeplitz GEMM, tcgen05.mma on a 2-term fp16 split of
 * both operands, f32 accumulation in TMEM drained periodically) is kept selectable. WBX_FIR=direct|tc|fft picks by hand
 * (read here); wbx_fir_path tells which one an engine uses.
 * h == NULL or n_taps == 0 removes it. Changing it clears every track's reverb history. */
int wbx_set_impulse_response(wbx_engine* e, const float* h, uint32_t n_taps);
/* Attach (fx != NULL) or remove (NULL) a track's chain and clear its state. Tracks without a chain take the
 * reference path untouched. */
int wbx_set_track_effects(wbx_engine* e, uint32_t track, const wbx_effects* fx);

/* ---- render ----------------------------------------------------------------------------------------- */
/* One call = n_blocks consecutive Engine::process callbacks (n_blocks = 1 is the realtime callback).
 *   track_gains  [n_tracks][2] f32: (mute ? 0 : volume) * pan_coeffs[c]            (track.cpp:728-731)
 *   out_channels out_channels pointers, each to n_blocks*block_frames f32 (AudioBuffer layout, one
 *                channel = one contiguous array, core/audio_buffer.h:19-23); callback b is frames
 *                [b*block_frames, (b+1)*block_frames). Clamped to [-1, 1] (NaN passes, engine.cpp:1627-36).
 *   peaks        [n_blocks][n_tracks][2] f32 block peak per track/channel = what VUMeter::push_samples
 *                offers to `level` in that callback (vu_meter.h:20-30); may be NULL. */
int wbx_render(wbx_engine* e, const wbx_segment* segs, uint32_t n_segs, const float* track_gains,
               uint32_t n_blocks, float* const* out_channels, float* peaks);
/* wbx_render that also returns levels[n_tracks][2] (see wbx_fetch_levels; may be NULL) under the same single
 * synchronisation. When every out_channels[c] is page-locked (wbx_host_alloc) the mix kernel stores each finished bus
 * tile into the caller's channels itself (posted writes over PCIe, next to the device copy of the bus), so no
 * device-to-host copy of the bus follows the kernel. On an engine in a connected sharded setup (wbx_shard_*) the mix is
 * the sharded one: out_channels is only written on rank 0. */
int wbx_render_levels(wbx_engine* e, const wbx_segment* segs, uint32_t n_segs, const float* track_gains,
                      uint32_t n_blocks, float* const* out_channels, float* peaks, float* levels);

/* The three stages of wbx_render, for callers that keep data on the device (sharded multi-GPU mixing,
 * benchmarking the kernel alone):
 *   wbx_submit  H2D of the segment table + gains, device-side schedule expansion (asynchronous)
 *   wbx_mix     the fused sampler/gain/pan/peak/bus-sum/clamp kernel over the submitted schedule
 *   wbx_fetch   D2H of bus and peaks into AudioBuffer-style host channels, then stream synchronise */
int wbx_submit(wbx_engine* e, const wbx_segment* segs, uint32_t n_segs, const float* track_gains,
               uint32_t n_blocks);
int wbx_mix(wbx_engine* e, uint32_t flags);
int wbx_fetch(wbx_engine* e, float* const* out_channels, float* peaks);
/* VUMeter::level over the whole render: levels[n_tracks][2] = max over callbacks of the block peaks
 * (level only rises until the UI reads it, engine/vu_meter.h:25-29). Reduced on the device: 8 bytes per
 * track cross PCIe instead of the full peaks array. */
int wbx_fetch_levels(wbx_engine* e, float* levels);
/* Page-locked host memory: channel / peak buffers allocated here receive the device copy directly instead
 * of through the engine's staging buffer (AudioBuffer allocations, core/audio_buffer.h:34, may use it). */
void* wbx_host_alloc(size_t bytes);
void wbx_host_free(void* p);
/* planar f32 bus -> interleaved device format (core/audio_format_conv.cpp:5-106) fused after the clamp;
 * dst_format: WBX_FMT_I16 / I24 (3 packed bytes) / I24_X8 / I32 / F32. dst is host memory of
 * n_blocks*block_frames*out_channels elements. */
int wbx_fetch_interleaved(wbx_engine* e, void* dst, int dst_format);

/* Offline bounce (the export driver the reference's export dialog lacks, ui/export_audio_dlg.cpp:44-200,
 * engine/export_prop.h:14-45): a long render is cut into chunks of callbacks; each chunk is converted to the device format
 * on the engine's stream right after its mix and copied to page-locked host memory on a second stream, under the next
 * chunk's submit + mix. Per chunk:  wbx_submit, wbx_mix, wbx_bounce_push;  wbx_bounce_pop hands out the oldest pushed
 * chunk (frames * out_channels elements of dst_format, interleaved as wbx_fetch_interleaved) — the pointer stays valid
 * until the next wbx_bounce_push reuses the slot, i.e. until the pop after next. At most two chunks are in flight. */
int wbx_bounce_begin(wbx_engine* e, int dst_format);
int wbx_bounce_push(wbx_engine* e);
int wbx_bounce_pop(wbx_engine* e, const void** data, size_t* bytes);

/* Device-side views of the last wbx_mix result (valid until the next wbx_submit):
 * bus [out_channels][n_blocks*block_frames] f32, peaks [n_blocks][n_tracks][2] f32. */
int wbx_device_bus(wbx_engine* e, float** d_bus, uint64_t* n_floats);
int wbx_device_peaks(wbx_engine* e, float** d_peaks, uint64_t* n_floats);
/* Clamp n floats at a device pointer to [-1, 1] on the engine's stream (engine.cpp:1627-1636) — the step
 * that must follow the cross-GPU bus reduce when tracks are sharded. */
int wbx_clamp_device(wbx_engine* e, float* d_bus, uint64_t n_floats);
int wbx_synchronize(wbx_engine* e);

/* ---- sharded render: tracks split over the GPUs of one box ---------------------------------------
 * Tracks are independent until AudioBuffer::mix adds them into the bus (engine/engine.cpp:1600-1617), so every rank
 * (one engine per GPU; one process per GPU or several engines in one process) owns a subset of the tracks and its
 * samples, and the only exchange is the bus sum, followed by the clamp (engine.cpp:1627-1636). That exchange runs
 * over peer memory (NVLink / NVSwitch), fused into the mix: callback k belongs to owner rank k / ceil(n_blocks /
 * world); each rank's mix kernel stores its UNCLAMPED bus tiles directly into the owner's exchange buffer while it
 * mixes, a flag barrier between the GPUs follows, every owner adds the world partial buses of its callbacks in rank
 * order 0..world-1 (deterministic; re-associated relative to one engine holding all tracks, like WBX_SUM_TREE),
 * clamps, and stores the slice into rank 0's master bus; a second barrier completes the step.
 *   1. wbx_shard_init on every rank (after wbx_configure): allocates the rank's exchange block and returns its CUDA
 *      IPC handle (WBX_IPC_HANDLE_BYTES bytes; ipc_handle_out may be NULL for same-process use);
 *   2. ranks exchange the handles by any means (bench.py: torch.distributed all_gather) and call
 *      wbx_shard_connect_ipc(handles[world][WBX_IPC_HANDLE_BYTES]) — or, engines of one process,
 *      wbx_shard_connect_local(engines[world]) once every engine is initialised;
 *   3. per render: wbx_submit, then wbx_mix_sharded on EVERY rank, the same number of times and with the same
 *      n_blocks <= max_blocks (it is a collective). wbx_fetch / wbx_device_bus / wbx_fetch_interleaved on rank 0 then
 *      see the clamped master bus; peaks and levels stay per rank (they are per track).
 * A rank that never arrives makes the barrier give up after 10 s (WBX_SHARD_TIMEOUT_MS) and the next synchronising
 * call return WBX_ERR_CUDA instead of hanging the GPU. */
#define WBX_IPC_HANDLE_BYTES 64
#define WBX_MAX_SHARD_RANKS 16
int wbx_shard_init(wbx_engine* e, uint32_t rank, uint32_t world, uint32_t max_blocks, void* ipc_handle_out);
int wbx_shard_connect_ipc(wbx_engine* e, const void* handles);
int wbx_shard_connect_local(wbx_engine* e, wbx_engine* const* engines);
int wbx_mix_sharded(wbx_engine* e);
/* The same collective in its three phases, for ONE thread driving several engines (it must issue phase p on every
 * engine before phase p+1 on any, so that no engine's stream waits for work that has not been enqueued yet):
 *   0  mix into the owners' exchange buffers + arrival signal      1  wait, reduce own slice + clamp into the master
 *   bus, signal      2  wait. wbx_mix_sharded(e) (one process or thread per GPU) = phase 0's mix followed by the rest of the
 *   exchange fused into ONE kernel launch (signal, wait, reduce + clamp, signal, wait); `phase` 3 selects that form. */
int wbx_mix_sharded_phase(wbx_engine* e, int phase);
/* Recovery. A phase that fails (a launch error, or a peer that never arrived: WBX_ERR_CUDA from the next synchronising
 * call) ends the running collective on this rank — the engine accepts a new wbx_mix_sharded afterwards — but the ranks'
 * barrier epochs may then disagree. To resume, EVERY rank calls wbx_shard_reset (drains its stream, clears its arrival
 * words, restarts its epoch at 0), the caller makes sure all ranks have done so (any host-side barrier), and rendering
 * continues with the connections and buffers as they were. */
int wbx_shard_reset(wbx_engine* e);
/* Optional: the host buffer the master bus is wanted in (channels[c] -> frames_per_channel f32, page-locked and mapped on
 * this engine's device: wbx_host_alloc memory within one process, or a shared-memory segment every process registered with
 * wbx_host_register). Set on EVERY rank, each with its own mapping of the SAME buffer: an owner then also stores its reduced
 * slice there (over its own PCIe link) and rank 0's wbx_fetch / wbx_render into exactly these channels copies nothing.
 * channels == NULL clears it. Renders longer than frames_per_channel fall back to the copy from rank 0's master bus. */
int wbx_shard_set_host_output(wbx_engine* e, float* const* channels, uint64_t frames_per_channel);
/* cudaHostRegister(portable | mapped) / cudaHostUnregister for caller-owned host memory (e.g. a shared-memory segment) */
int wbx_host_register(void* p, size_t bytes);
int wbx_host_unregister(void* p);
int wbx_shard_close(wbx_engine* e);
/* rank / world of a connected sharded setup, (0, 1) otherwise */
int wbx_shard_info(const wbx_engine* e, uint32_t* rank, uint32_t* world);

/* ---- introspection --------------------------------------------------------------------------------- */
/* Number of CUDA kernels this engine has launched since creation (bench.py's gpu_launches). */
uint64_t wbx_launch_count(const wbx_engine* e);
/* Reduced-precision tensor-core products issued per tap by the convolution reverb (split-precision factor; 3 = 2-term fp16
 * split of both operands). */
int wbx_fir_split_factor(void);
/* The path the convolution reverb of this engine's impulse response takes: 0 direct form (CUDA cores), 1 direct form as a
 * Toeplitz GEMM on the tensor cores, 2 partitioned FFT convolution. Chosen at wbx_set_impulse_response (>= 1024 taps: 2). */
int wbx_fir_path(const wbx_engine* e);
/* Name of the mix kernel variant the last wbx_mix used ("exact/vec16", "tree/g8", ...). */
const char* wbx_last_kernel(const wbx_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* WBX_H */
