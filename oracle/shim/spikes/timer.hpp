// TEST INFRASTRUCTURE (oracle): stand-in for <spikes/timer.hpp>.
// tic() returns milliseconds since construction (cumulative), which is what
// reference test/navier_stokes_2d_p2_p1.cpp:168-172 relies on.
#ifndef ORACLE_SHIM_SPIKES_TIMER_HPP
#define ORACLE_SHIM_SPIKES_TIMER_HPP

#include <chrono>

class timer {
public:
  timer() : t0_(std::chrono::steady_clock::now()) {}
  double tic() const {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0_).count();
  }
private:
  std::chrono::steady_clock::time_point t0_;
};

#endif
