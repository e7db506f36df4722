// Shim: boost::thread -> std::thread (the reference copies a functor into each
// thread and joins; common/kernel_matrix.cpp:536-554).
#pragma once
#include <mutex>
#include <thread>
namespace boost {
using thread = std::thread;
}
