#pragma once
namespace spdlog {
template <class... A> inline void info(const A &...) {}
template <class... A> inline void warn(const A &...) {}
template <class... A> inline void error(const A &...) {}
template <class... A> inline void debug(const A &...) {}
template <class... A> inline void critical(const A &...) {}
}  // namespace spdlog
#define SPDLOG_INFO(...) do { } while (0)
#define SPDLOG_WARN(...) do { } while (0)
#define SPDLOG_ERROR(...) do { } while (0)
#define SPDLOG_DEBUG(...) do { } while (0)
#define SPDLOG_CRITICAL(...) do { } while (0)
