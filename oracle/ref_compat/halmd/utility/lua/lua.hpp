// TEST INFRASTRUCTURE ONLY: stand-in for halmd/utility/lua/lua.hpp (Lua and luaponte are absent).  The luaopen() bodies of
// the reference's host classes compile against a catch-all binding DSL that accepts every expression and does nothing, so
// the .cpp files can be compiled in place for their C++ logic.
#ifndef HALMD_UTILITY_LUA_LUA_HPP
#define HALMD_UTILITY_LUA_LUA_HPP
#include <lua.hpp>
#include <memory>
#include <halmd/config.hpp>
#ifndef HALMD_LUA_API
# define HALMD_LUA_API extern "C"
#endif
namespace luaponte {
struct sink;
struct scope_proxy { template <typename A> sink operator[](A&&) const; };
struct sink {
    sink() {}
    template <typename... A> explicit sink(A&&...) {}
    template <typename... A> sink& def(A&&...) { return *this; }
    template <typename... A> sink& property(A&&...) { return *this; }
    template <typename... A> sink& def_readonly(A&&...) { return *this; }
    template <typename... A> sink& def_readwrite(A&&...) { return *this; }
    template <typename... A> sink& enum_(A&&...) { return *this; }
    template <typename A> sink& operator[](A&&) { return *this; }
    scope_proxy scope;              // `.scope[ ... ]`
};
template <typename A> sink scope_proxy::operator[](A&&) const { return sink(); }
inline sink operator,(sink const& a, sink const&) { return a; }
template <typename... T> struct class_ : sink { template <typename... A> class_(A&&...) {} };
template <typename... T> struct constructor {};
template <typename... A> sink module(A&&...) { return sink(); }
template <typename... A> sink namespace_(A&&...) { return sink(); }
template <typename... A> sink def(A&&...) { return sink(); }
template <typename... A> sink value(A&&...) { return sink(); }
struct no_bases {};
template <typename... T> struct bases {};
} // namespace luaponte
#endif
