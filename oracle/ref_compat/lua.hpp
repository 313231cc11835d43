// TEST INFRASTRUCTURE ONLY: stand-in for the Lua C API header (Lua is absent); the reference's host classes only name
// lua_State in their luaopen() declarations.
#ifndef HALMD_REF_COMPAT_LUA_HPP
#define HALMD_REF_COMPAT_LUA_HPP
struct lua_State;
#endif
