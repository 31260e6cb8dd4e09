// Test infrastructure only (oracle/_ref build). The reference pins moodycamel concurrentqueue
// 1.0.4 through Conan (/root/reference/conanfile.py); it is not vendored. An identical copy lives
// in this image inside dmlc-core under namespace dmlc::moodycamel; the Makefile adds its include
// directory. Only include/executor/task_queue.hpp uses it (enqueue / try_dequeue).
#pragma once
#include <dmlc/concurrentqueue.h>
namespace moodycamel = dmlc::moodycamel;
