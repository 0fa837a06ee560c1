// TEST INFRASTRUCTURE ONLY (oracle build). No-op stand-in for the two fmt entry points the
// reference's engine TUs name (engine.cpp:120 format_to; track.cpp:497,540,563 format_to_n). Both are
// on recording / MIDI-debug paths the oracle never executes.
#pragma once
#include <cstddef>
namespace fmt {
template<typename Out, typename... A>
inline Out format_to(Out out, const char*, A&&...) { return out; }
template<typename Out>
struct format_to_n_result { Out out; std::size_t size; };
template<typename Out, typename... A>
inline format_to_n_result<Out> format_to_n(Out out, std::size_t, const char*, A&&...) { return {out, 0}; }
}  // namespace fmt
