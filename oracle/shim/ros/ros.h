// Shim of <ros/ros.h> for building the reference hot path without ROS.
// TEST INFRASTRUCTURE ONLY (oracle/): lets /root/reference/src/*.cpp compile verbatim, in place.
// Publishers capture what the node would publish into shim::capture() so a C entry can read it back.
#pragma once
#include <cstdint>
#include <cstdio>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include <functional>

namespace shim {
struct Capture;          // defined in shim_capture.h (needs pcl + marker types)
Capture& capture();
template <class M> void capture_publish(const std::string& topic, const M& msg);
}  // namespace shim

namespace ros {
struct Time { double t = 0; Time() = default; static Time now() { return Time(); } };
struct Duration { double d = 0; Duration() = default; explicit Duration(double s) : d(s) {} };

class Publisher {
 public:
  Publisher() = default;
  explicit Publisher(std::string t) : topic_(std::move(t)) {}
  template <class M> void publish(const M& msg) const { shim::capture_publish(topic_, msg); }
  const std::string& getTopic() const { return topic_; }
 private:
  std::string topic_;
};
class Subscriber {};

class NodeHandle {
 public:
  template <class M, class T>
  Subscriber subscribe(const std::string&, uint32_t, void (T::*)(M), T*) { return Subscriber(); }
  template <class M> Publisher advertise(const std::string& topic, uint32_t) { return Publisher(topic); }
};
namespace this_node { inline const std::string& getName() { static std::string n = "urban_road_filt"; return n; } }
inline void init(int&, char**, const std::string&) {}
inline void spin() {}
}  // namespace ros

#ifndef ROS_INFO
#define ROS_INFO(...) do { } while (0)
#define ROS_ERROR(...) do { std::fprintf(stderr, "[ROS_ERROR] " __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)
#define ROS_FATAL(...) do { std::fprintf(stderr, "[ROS_FATAL] " __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)
#define ROS_ERROR_THROTTLE(period, ...) do { std::fprintf(stderr, "[ROS_ERROR] " __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)
#define ROS_WARN_THROTTLE(period, ...) do { std::fprintf(stderr, "[ROS_WARN] " __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)
#endif
