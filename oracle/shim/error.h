#ifndef SHIM_ERROR_H
#define SHIM_ERROR_H
#include "pointers.h"
#include <stdexcept>
namespace LAMMPS_NS {
class Error {
 public:
  [[noreturn]] void all(const char *file, int line, const std::string &msg) {
    throw std::runtime_error(std::string("ERROR: ") + msg + " (" + file + ":" + std::to_string(line) + ")");
  }
  [[noreturn]] void one(const char *file, int line, const std::string &msg) { all(file, line, msg); }
  void warning(const char *, int, const std::string &msg) { fprintf(stderr, "WARNING: %s\n", msg.c_str()); }
};
}
#endif
