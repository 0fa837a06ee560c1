// TEST INFRASTRUCTURE ONLY (oracle build). No-op stand-in for <spdlog/spdlog.h>.
// The reference's logging wrapper (src/core/debug.h:11-52) only needs these names to exist; the
// oracle never logs. Written from scratch for this repo; nothing here is spdlog code.
#pragma once
#include <initializer_list>
#include <memory>
#include <string>
#include <string_view>
#include <type_traits>
#include <utility>

namespace spdlog {
namespace level {
enum level_enum { trace, debug, info, warn, err, critical, off };
}
namespace sinks {
struct sink {};
struct stdout_color_sink_mt : sink {};
}  // namespace sinks
using sink_ptr = std::shared_ptr<sinks::sink>;
using sinks_init_list = std::initializer_list<sink_ptr>;

template<typename... Args>
struct basic_format_string {
  template<typename S>
  consteval basic_format_string(const S&) {}
};
template<typename... Args>
using format_string_t = basic_format_string<std::type_identity_t<Args>...>;

class logger {
 public:
  logger(std::string, sinks_init_list) {}
  void set_level(level::level_enum) {}
  template<typename... A> void trace(format_string_t<A...>, A&&...) {}
  template<typename... A> void debug(format_string_t<A...>, A&&...) {}
  template<typename... A> void info(format_string_t<A...>, A&&...) {}
  template<typename... A> void warn(format_string_t<A...>, A&&...) {}
  template<typename... A> void error(format_string_t<A...>, A&&...) {}
  template<typename... A> void critical(format_string_t<A...>, A&&...) {}
};
}  // namespace spdlog
