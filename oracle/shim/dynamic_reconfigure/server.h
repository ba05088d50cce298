// oracle shim (test infrastructure only): dynamic_reconfigure::Server surface used by the reference's main.cpp.
#pragma once
#include <functional>
#include <cstdint>
namespace dynamic_reconfigure {
template <class ConfigT> class Server {
 public:
  typedef std::function<void(ConfigT&, uint32_t)> CallbackType;
  void setCallback(const CallbackType& cb) { ConfigT c; cb(c, 0xffffffffu); }
};
}
