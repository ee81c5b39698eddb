// ref_shim stub: utility::LogError throws (utility/Logging.h:44-53); the rest is silent.
#pragma once
#include <stdexcept>
#include <string>
namespace open3d {
namespace utility {
template <typename... Args>
[[noreturn]] inline void LogError(const char* fmt, Args&&...) { throw std::runtime_error(fmt); }
template <typename... Args>
inline void LogWarning(const char*, Args&&...) {}
template <typename... Args>
inline void LogInfo(const char*, Args&&...) {}
template <typename... Args>
inline void LogDebug(const char*, Args&&...) {}
}  // namespace utility
}  // namespace open3d
