// jit.hpp -- panel-specialised scan kernel (NVRTC).  Placeholder: filled in below.
#pragma once
#include <string>
#include "plan.hpp"
namespace ipcr {
struct JitKernel {
    void *library = nullptr;
    void *kernel = nullptr;
    bool ok = false;
    std::string log;
};
inline std::string jit_generate_source(const Panel &, int) { return std::string(); }
}  // namespace ipcr
