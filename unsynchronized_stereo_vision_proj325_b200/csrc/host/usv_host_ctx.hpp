// One GPU context per (host thread, device): the reference's matching/distance functions are
// called concurrently from four threads and share no state, so the replacements
// keep their contexts thread-local.
#pragma once
#include <cstdio>
#include <string>
#include <vector>

#include "../../../include/usv_b200.h"

namespace usv {
struct ThreadContexts {
  std::vector<usv_ctx*> by_device;
  std::string last_error;
  ~ThreadContexts() {
    for (usv_ctx* c : by_device)
      if (c) usv_destroy(c);
  }
  usv_ctx* get(int device) {
    if (device < 0) return nullptr;
    if ((size_t)device >= by_device.size()) by_device.resize(device + 1, nullptr);
    if (!by_device[device]) {
      int rc = usv_create(device, &by_device[device]);
      if (rc != USV_OK) {
        by_device[device] = nullptr;
        last_error = "usv_create failed (" + std::to_string(rc) + "): no usable sm_100 CUDA device; there is no CPU fallback";
      }
    }
    return by_device[device];
  }
  bool check(usv_ctx* c, int rc, const char* what) {
    if (rc == USV_OK) { last_error.clear(); return true; }
    last_error = std::string(what) + " failed (" + std::to_string(rc) + "): " + (c ? usv_last_error(c) : "no context");
    return false;
  }
};
inline ThreadContexts& thread_contexts() {
  static thread_local ThreadContexts t;
  return t;
}
}  // namespace usv
