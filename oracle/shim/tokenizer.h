#ifndef SHIM_TOKENIZER_H
#define SHIM_TOKENIZER_H
#include <exception>
#include <string>
namespace LAMMPS_NS {
class TokenizerException : public std::exception {
  std::string message;
 public:
  explicit TokenizerException(const std::string &msg, const std::string & = "") : message(msg) {}
  const char *what() const noexcept override { return message.c_str(); }
};
}
#endif
