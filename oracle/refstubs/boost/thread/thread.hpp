#pragma once
// boost::thread over the C++11 standard library (stand-in for the reference's libcore/{thread,lock,sched}.cpp).
#include <thread>
#include <mutex>
#include <condition_variable>
#include <chrono>
#include <stdexcept>
#include <utility>
namespace boost {
struct thread_interrupted {};
struct thread_resource_error : std::runtime_error { thread_resource_error() : std::runtime_error("thread_resource_error") {} };
namespace posix_time {
typedef std::chrono::steady_clock::time_point ptime;
inline std::chrono::milliseconds milliseconds(long ms) { return std::chrono::milliseconds(ms); }
}
typedef std::chrono::steady_clock::time_point system_time;
inline system_time get_system_time() { return std::chrono::steady_clock::now(); }
using std::lock_guard;
using std::unique_lock;
class mutex : public std::mutex { public: typedef std::unique_lock<mutex> scoped_lock; };
class recursive_mutex : public std::recursive_mutex { public: typedef std::unique_lock<recursive_mutex> scoped_lock; };
class timed_mutex : public std::timed_mutex {
public:
    typedef std::unique_lock<timed_mutex> scoped_lock;
    bool timed_lock(const system_time &t) { return try_lock_until(t); }
};
class recursive_timed_mutex : public std::recursive_timed_mutex {
public:
    typedef std::unique_lock<recursive_timed_mutex> scoped_lock;
    bool timed_lock(const system_time &t) { return try_lock_until(t); }
};
class condition_variable_any : public std::condition_variable_any {
public:
    template <typename L> bool timed_wait(L &lock, const system_time &t) { return wait_until(lock, t) == std::cv_status::no_timeout; }
};
typedef condition_variable_any condition_variable;
class thread {
    std::thread t;
public:
    thread() {}
    template <typename F, typename A> thread(F f, A a) : t(f, a) {}
    template <typename F> explicit thread(F f) : t(f) {}
    thread(thread &&o) : t(std::move(o.t)) {}
    thread &operator=(thread &&o) { t = std::move(o.t); return *this; }
    void join() { if (t.joinable()) t.join(); }
    void detach() { if (t.joinable()) t.detach(); }
    bool joinable() const { return t.joinable(); }
    std::thread::native_handle_type native_handle() { return t.native_handle(); }
    static unsigned hardware_concurrency() { return std::thread::hardware_concurrency(); }
};
namespace this_thread {
template <typename D> inline void sleep(const D &d) { std::this_thread::sleep_for(d); }
inline void yield() { std::this_thread::yield(); }
}
}
