// TEST INFRASTRUCTURE - stand-in for rclcpp, just large enough for the reference's map_builder node
// (mapping_util/src/map_builder.cpp) to compile UNMODIFIED into oracle/_ref/libref_map.so: a Node that keeps its
// parameters in a map (with overrides the test wrapper sets before construction), publishers that remember the last
// message, subscriptions that do nothing.  No ROS2 code; nothing of the product includes this.
#ifndef HDSM_REF_SHIM_RCLCPP_HPP_
#define HDSM_REF_SHIM_RCLCPP_HPP_
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <future>
#include <thread>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace builtin_interfaces { namespace msg { struct Time { int32_t sec = 0; uint32_t nanosec = 0; }; } }
namespace std_msgs { namespace msg { struct Header { builtin_interfaces::msg::Time stamp; std::string frame_id; }; } }
namespace tf2_msgs { namespace msg { struct TFMessage; } }

namespace rclcpp {
struct Time {
  int64_t ns = 0;
  int64_t nanoseconds() const { return ns; }
  operator builtin_interfaces::msg::Time() const { return builtin_interfaces::msg::Time(); }
};
struct Clock { typedef std::shared_ptr<Clock> SharedPtr; };
class Parameter {
 public:
  std::string s; int64_t i = 0; double d = 0; bool b = false; std::vector<double> dv;
  Parameter() {}
  Parameter(const char* v) : s(v) {}
  Parameter(const std::string& v) : s(v) {}
  Parameter(int v) : i(v) {}
  Parameter(double v) : d(v) {}
  Parameter(bool v) : b(v) {}
  Parameter(const std::vector<double>& v) : dv(v) {}
  std::string as_string() const { return s; }
  int64_t as_int() const { return i; }
  double as_double() const { return d; }
  bool as_bool() const { return b; }
  std::vector<double> as_double_array() const { return dv; }
};
inline std::map<std::string, Parameter>& parameter_overrides() { static std::map<std::string, Parameter> m; return m; }
template <class T> struct Publisher {
  typedef std::shared_ptr<Publisher<T>> SharedPtr;
  T last; int count = 0;
  void publish(const T& m) { last = m; ++count; }
};
template <class T> struct Subscription { typedef std::shared_ptr<Subscription<T>> SharedPtr; };
struct TimerBase { typedef std::shared_ptr<TimerBase> SharedPtr; };
struct Logger {};
inline bool& ok_flag() { static bool f = false; return f; }  // the node's worker loops (while (rclcpp::ok())) end at once
inline bool ok() { return ok_flag(); }
template <class S> struct Client {
  typedef std::shared_ptr<Client<S>> SharedPtr;
  typedef std::shared_future<std::shared_ptr<typename S::Response>> SharedFuture;
  template <class D> bool wait_for_service(D) { return true; }
  template <class F> SharedFuture async_send_request(std::shared_ptr<typename S::Request>, F&&) {
    std::promise<std::shared_ptr<typename S::Response>> p;
    p.set_value(std::make_shared<typename S::Response>());
    return p.get_future().share();
  }
};
class Node {
 public:
  explicit Node(const std::string&) {}
  virtual ~Node() {}
  template <class T> void declare_parameter(const std::string& name, const T& def) {
    auto it = parameter_overrides().find(name);
    params_[name] = it != parameter_overrides().end() ? it->second : Parameter(def);
  }
  Parameter get_parameter(const std::string& name) const { return params_.at(name); }
  template <class F> void on_shutdown(F&&) {}
  template <class T, class F> typename Subscription<T>::SharedPtr create_subscription(const std::string&, int, F&&) { return std::make_shared<Subscription<T>>(); }
  template <class T> typename Publisher<T>::SharedPtr create_publisher(const std::string&, int) { return std::make_shared<Publisher<T>>(); }
  template <class S> typename Client<S>::SharedPtr create_client(const std::string&) { return std::make_shared<Client<S>>(); }
  template <class D, class F> TimerBase::SharedPtr create_wall_timer(D, F&&) { return std::make_shared<TimerBase>(); }
  Logger get_logger() const { return Logger(); }
  Time now() const { return Time(); }
  Clock::SharedPtr get_clock() { return std::make_shared<Clock>(); }
 private:
  std::map<std::string, Parameter> params_;
};
}  // namespace rclcpp
#define RCLCPP_INFO(logger, ...) do { (void)(logger); } while (0)
#define RCLCPP_ERROR(logger, ...) do { (void)(logger); } while (0)
#define RCLCPP_WARN(logger, ...) do { (void)(logger); } while (0)
#endif
