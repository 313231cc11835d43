// TEST INFRASTRUCTURE ONLY: empty stand-in (Boost is absent; utility/scoped_timer.hpp is replaced by ref_compat's profiler stub)
