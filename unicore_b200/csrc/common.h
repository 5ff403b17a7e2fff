// Shared host-side helpers: error plumbing for the C ABI (thread-local last error, never throw
// across the boundary) and CUDA call checking.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

namespace p5 {

// error codes returned through the C ABI (include/prostt5_b200.h)
enum : int {
    P5_OK = 0,
    P5_ERR_ARG = 1,
    P5_ERR_IO = 2,
    P5_ERR_FORMAT = 3,
    P5_ERR_CUDA = 4,
    P5_ERR_NOMEM = 5,
    P5_ERR_UNSUPPORTED = 6,
};

// Experiment knobs (environment variables) exist only in the debug library (libprostt5_b200_debug.so, built with
// -DP5_DEBUG_BUILD): the product library always takes the default, so no environment can make it skip work.
#ifdef P5_DEBUG_BUILD
inline int env_knob(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}
inline bool env_flag(const char* name) { return getenv(name) != nullptr; }
#else
inline int env_knob(const char*, int dflt) { return dflt; }
inline bool env_flag(const char*) { return false; }
#endif

struct Error : public std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

inline std::string strf(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    return std::string(buf);
}

void set_last_error(const std::string& msg);

#define P5_CUDA(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            throw ::p5::Error(::p5::P5_ERR_CUDA, ::p5::strf("%s failed: %s (%s:%d)", #expr,             \
                                                            cudaGetErrorString(_e), __FILE__, __LINE__)); \
    } while (0)

#define P5_REQUIRE(cond, code, ...)                                    \
    do {                                                               \
        if (!(cond)) throw ::p5::Error((code), ::p5::strf(__VA_ARGS__)); \
    } while (0)

// Run `body` and translate exceptions to a C error code + thread-local message.
template <class F>
int guarded(F&& body) noexcept {
    try {
        body();
        return P5_OK;
    } catch (const Error& e) {
        set_last_error(e.what());
        return e.code;
    } catch (const std::bad_alloc&) {
        set_last_error("out of host memory");
        return P5_ERR_NOMEM;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return P5_ERR_ARG;
    } catch (...) {
        set_last_error("unknown error");
        return P5_ERR_ARG;
    }
}

}  // namespace p5
