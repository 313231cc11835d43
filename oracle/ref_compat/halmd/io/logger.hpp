// TEST INFRASTRUCTURE ONLY: stand-in for halmd/io/logger.hpp (Boost.Log is absent): log statements compile to nothing.
#ifndef HALMD_IO_LOGGER_HPP
#define HALMD_IO_LOGGER_HPP
#include <memory>
#include <sstream>
#include <string>
namespace halmd {
class logger { public: logger() {} explicit logger(std::string const&) {} };
} // namespace halmd
#define HALMD_REF_COMPAT_LOG(expr) do { if (false) { std::ostringstream halmd_ref_compat_s; halmd_ref_compat_s << expr; } } while (0)
#define LOG(expr) HALMD_REF_COMPAT_LOG(expr)
#define LOG_ERROR(expr) HALMD_REF_COMPAT_LOG(expr)
#define LOG_WARNING(expr) HALMD_REF_COMPAT_LOG(expr)
#define LOG_INFO(expr) HALMD_REF_COMPAT_LOG(expr)
#define LOG_DEBUG(expr) HALMD_REF_COMPAT_LOG(expr)
#define LOG_TRACE(expr) HALMD_REF_COMPAT_LOG(expr)
#define LOG_ONCE(expr) HALMD_REF_COMPAT_LOG(expr)
#define LOG_WARNING_ONCE(expr) HALMD_REF_COMPAT_LOG(expr)
#endif
