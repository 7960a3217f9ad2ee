// Exception convention of the reference at the Python boundary: every failure surfaces as
// Pennylane::Util::LightningException -> Python `PLException` (bindings/Bindings.cpp:1734 of the
// reference).  Here every non-zero status of the C ABI is rethrown with qsv_last_error()'s text.
#pragma once
#include <exception>
#include <string>

#include "qsv_b200.h"

namespace Pennylane::Util {

class LightningException : public std::exception {
  public:
    explicit LightningException(std::string msg) : msg_(std::move(msg)) {}
    const char *what() const noexcept override { return msg_.c_str(); }

  private:
    std::string msg_;
};

[[noreturn]] inline void Abort(const std::string &message, const char *file, int line, const char *func) {
    throw LightningException(std::string("[") + file + "][Line:" + std::to_string(line) + "][Method:" + func +
                             "]: Error in PennyLane Lightning: " + message);
}

inline void check(int status) {
    if (status != 0) throw LightningException(qsv_last_error());
}

}  // namespace Pennylane::Util

#define PL_ABORT(message) ::Pennylane::Util::Abort(message, __FILE__, __LINE__, __func__)
#define PL_ABORT_IF(cond, message)                                                                 \
    do {                                                                                           \
        if (cond) PL_ABORT(message);                                                               \
    } while (0)
#define PL_ABORT_IF_NOT(cond, message) PL_ABORT_IF(!(cond), message)
#define PL_ASSERT(cond) PL_ABORT_IF_NOT(cond, "Assertion failed: " #cond)
