// ORACLE — TEST INFRASTRUCTURE ONLY. Stand-in for the Boost headers the reference includes (system dependency, absent from this image): the
// handful of names the hot-path headers mention, mapped onto their C++17 standard-library equivalents. No arithmetic lives here.
#pragma once
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
namespace boost {
using std::shared_ptr;
using std::make_shared;
using std::function;
using std::mutex;
using std::unique_lock;
using std::lock_guard;
using std::condition_variable;
// boost::thread's cooperative interruption (DepthFilter::stopThread / updateSeedsLoop, src/depth_filter.cpp:121,244): the tests drive the
// per-seed functions directly and never start the mapping thread, so interrupt() has nothing to do and no interruption is ever pending
class thread : public std::thread {
 public:
  using std::thread::thread;
  void interrupt() {}
};
using std::ref;
using std::cref;
class noncopyable {
 protected:
  noncopyable() = default;
  ~noncopyable() = default;
  noncopyable(const noncopyable&) = delete;
  noncopyable& operator=(const noncopyable&) = delete;
};
template <class... A> auto bind(A&&... a) -> decltype(std::bind(std::forward<A>(a)...)) { return std::bind(std::forward<A>(a)...); }
namespace this_thread { using std::this_thread::yield; using std::this_thread::sleep_for; inline bool interruption_requested() { return false; } }
}  // namespace boost
// boost::bind expressions can be compared (boost/bind/bind.hpp: relational operators build a new bind expression); Reprojector::reprojectMap sorts
// its close keyframes with `boost::bind(&pair::second, _1) < boost::bind(&pair::second, _2)` (src/reprojector.cpp:165). Declared in namespace std so
// that argument-dependent lookup finds it for std::bind's result types (test infrastructure only).
namespace std {
template <class A, class B, typename std::enable_if<std::is_bind_expression<A>::value && std::is_bind_expression<B>::value, int>::type = 0>
auto operator<(A a, B b) {
  return [a, b](auto&&... args) mutable { return a(args...) < b(args...); };
}
}  // namespace std
// boost/bind.hpp puts _1.._9 into the global namespace
using namespace std::placeholders;
