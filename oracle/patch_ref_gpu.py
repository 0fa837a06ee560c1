#!/usr/bin/env python
"""TEST / INTEGRATION INFRASTRUCTURE. Produces the reference-side binding of INTEGRATION.md section A as real code:
patched COPIES of the reference's engine/engine.cpp and engine/track.cpp (written to oracle/_ref/patched/, which is
git-ignored — no reference source enters the repository) in which the sample loops of the hot path are replaced by calls
into the wbx C ABI through oracle/wbx_gpu_hooks.h:

  Track::process   the two dsp::Sampler::stream calls (track.cpp:678,718)      -> wbx_gpu::stream   (records a wbx_segment)
                   the dsp::apply_gain + VUMeter::push_samples loop (:728-733) -> wbx_gpu::track_gains
  Engine::process  output_buffer.clear / mixing_buffer.clear / ::mix           -> wbx_gpu::begin
                   (engine.cpp:1598,1602,1616) and the clamp loop (:1627-1636) -> wbx_gpu::render   (one wbx_render_levels)

Everything else of both files — transport, editor lock, process_event, parameter messages, plugin slot — stays the
reference's own code. Every substitution asserts how many times its anchor matched, so a changed reference fails loudly.

    python oracle/patch_ref_gpu.py /root/reference/src oracle/_ref/patched
"""
import os
import re
import sys


def sub(text, pattern, repl, count, what, flags=re.S):
    new, n = re.subn(pattern, repl, text, flags=flags)
    if n != count:
        raise SystemExit("patch_ref_gpu: %s: expected %d match(es), found %d" % (what, count, n))
    return new


def main(src, out):
    os.makedirs(os.path.join(out, "engine"), exist_ok=True)
    inc = '#include "wbx_gpu_hooks.h"\n'

    t = open(os.path.join(src, "engine", "track.cpp")).read()
    t = sub(t, r"sampler\.stream\(sample, output_buffer\.n_channels, event_length, start_sample, gain, write_buffer\.channel_buffers\);",
            "wbx_gpu::stream(this, sampler, sample, event_length, start_sample, gain, plugin_instance != nullptr);", 2,
            "Sampler::stream calls in Track::process")
    t = sub(t, r"for \(uint32_t i = 0; i < output_buffer\.n_channels; i\+\+\) \{\s*float\* buf = output_buffer\.channel_buffers\[i\];\s*"
               r"dsp::apply_gain\(buf, output_buffer\.n_samples, volume \* parameter_state\.pan_coeffs\[i\]\);\s*"
               r"level_meter\[i\]\.push_samples\(output_buffer, i\);\s*\}",
            "wbx_gpu::track_gains(this, volume * parameter_state.pan_coeffs[0], volume * parameter_state.pan_coeffs[1]);", 1,
            "apply_gain + VU loop in Track::process")
    open(os.path.join(out, "engine", "track.cpp"), "w").write(inc + t)

    e = open(os.path.join(src, "engine", "engine.cpp")).read()
    e = sub(e, r"output_buffer\.clear\(\);(\s*for \(uint32_t i = 0; i < tracks\.size\(\); i\+\+\) \{\s*auto track = tracks\[i\];)\s*mixing_buffer\.clear\(\);",
            r"wbx_gpu::begin(this, output_buffer, sample_rate);\1", 1, "bus / mixing buffer clears in Engine::process")
    e = sub(e, r"(currently_playing\);)\s*output_buffer\.mix\(mixing_buffer\);", r"\1", 1, "AudioBuffer::mix in Engine::process")
    e = sub(e, r"for \(uint32_t i = 0; i < output_buffer\.n_channels; i\+\+\) \{\s*float\* channel = output_buffer\.get_write_pointer\(i\);\s*"
               r"for \(uint32_t j = 0; j < output_buffer\.n_samples; j\+\+\) \{.*?channel\[j\] = -1\.0;\s*\}\s*\}\s*\}",
            "wbx_gpu::render(this, output_buffer);", 1, "clamp loop in Engine::process")
    open(os.path.join(out, "engine", "engine.cpp"), "w").write(inc + e)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
