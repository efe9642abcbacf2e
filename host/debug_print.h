// debug_print.h -- the print / stopwatch macros the reference's callers use (names and printed text of
// src/debug_print.h:5-13: DEBUG_PRINT_ERR, DEBUG_PRINT_OUT, START_TIME(name) ... STOP_TIME(name) printing
// "<name> execution time : <seconds> sec").  Implemented on std::chrono; the arguments of the print macros are
// stream expressions ("a << b"), the stopwatch macros are statements.
#pragma once
#include <chrono>
#include <iostream>

#include "opencv2/core/utility.hpp"

namespace erp_host {
typedef std::chrono::steady_clock::time_point stopwatch_t;
inline stopwatch_t stopwatch_now() { return std::chrono::steady_clock::now(); }
inline double stopwatch_seconds(const stopwatch_t& since)
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - since).count();
}
} // namespace erp_host

// define ERP_HOST_QUIET to compile the messages out
#ifndef ERP_HOST_QUIET
#define DEBUG_PRINT
#endif

#ifdef DEBUG_PRINT
#define ERP_HOST_EMIT(stream, expr) do { (stream) << expr << std::endl; } while (0)
#define DEBUG_PRINT_OUT(x) ERP_HOST_EMIT(std::cout, x)
#define DEBUG_PRINT_ERR(x) ERP_HOST_EMIT(std::cerr, x)
#define START_TIME(name) erp_host::stopwatch_t name = erp_host::stopwatch_now()
#define STOP_TIME(name) DEBUG_PRINT_OUT(#name " execution time : " << erp_host::stopwatch_seconds(name) << " sec")
#else
#define DEBUG_PRINT_OUT(x) do { } while (0)
#define DEBUG_PRINT_ERR(x) do { } while (0)
#define START_TIME(name) do { } while (0)
#define STOP_TIME(name) do { } while (0)
#endif
